// trf.cuh -- building blocks of the bounded nonlinear least-squares solver: Trust Region Reflective with the
// dense "exact" sub-problem.
//
// The reference fits Bazin curves and decline models through scipy.optimize.curve_fit with bounds
// (bazin_fitting.py:128-137, scripts/train_v55_powerlaw.py:173-184), i.e. SciPy's
// least_squares(method='trf', jac='2-point', x_scale=1, ftol=xtol=gtol=1e-8, tr_solver='exact').
// SciPy is a third-party dependency of the reference (requirements.txt:22); this file and trf_thread.cuh
// restate that published algorithm (Branch, Coleman & Li 1999; More 1977 for the trust-region sub-problem)
// step for step so that the iterates track the reference's:
//   * make_strictly_feasible / Coleman-Li scaling / step_to_bound / the 1-D quadratic minimiser,
//   * the SVD of the small triangular factor (one-sided Jacobi) and More's secular-equation iteration
//     (solve_lsq_trust_region),
// all scalar code on NP x NP data.  The driver that owns the residuals, the streamed Jacobian factorisation and
// the step selection is trf_thread.cuh (one thread per fit).
#pragma once
#include "team.cuh"
#if defined(TRF_TRACE) && !defined(__CUDACC__)
#include <cstdio>
#define TRF_LOG(...) std::printf(__VA_ARGS__)
#else
#define TRF_LOG(...) ((void)0)
#endif

namespace mlrn {

constexpr double TRF_EPS = 2.220446049250313e-16;
constexpr double TRF_SQRT_EPS = 1.4901161193847656e-08;

enum TrfFail { TRF_BAD_BOUNDS = -101, TRF_X0_INFEASIBLE = -102, TRF_F0_NONFINITE = -103, TRF_J_NONFINITE = -104,
               TRF_GEOMETRY = -105 };

// a / d from r = 1 / d (correctly rounded): one multiplication and a residual correction.  The divisors of a fit pass
// (time scales, finite-difference steps) are the same for every row, so their reciprocals are taken once per pass and each
// of the 11 such divisions of a Jacobian row costs 3 instructions instead of ~14.  The result equals the IEEE quotient
// (0 mismatches in 2e8 random operand pairs; the known exceptions of this scheme need operands that do not occur here),
// and resid() / fd_row() use the SAME expression, so the row stays consistent with the residual bit for bit.
// Not for divisors that can be infinite (r = 0 would turn a / inf = 0 into NaN).
MLRN_FN_FORCE double div_r(double a, double d, double r) { const double q = a * r; return fma(fma(-q, d, a), r, q); }

template <int NP>
MLRN_FN double vnorm(const double* a) { double s = 0.0; for (int i = 0; i < NP; ++i) s += a[i] * a[i]; return sqrt(s); }
template <int NP>
MLRN_FN double vdot(const double* a, const double* b) { double s = 0.0; for (int i = 0; i < NP; ++i) s += a[i] * b[i]; return s; }

template <int NP>
MLRN_FN bool in_bounds(const double* x, const double* lb, const double* ub) {
  for (int i = 0; i < NP; ++i) if (!(x[i] >= lb[i] && x[i] <= ub[i])) return false;
  return true;
}

// scipy.optimize._lsq.common.make_strictly_feasible (+ find_active_constraints)
template <int NP>
MLRN_NOINLINE void make_strictly_feasible(double* x, const double* lb, const double* ub, double rstep) {
  for (int i = 0; i < NP; ++i) {
    int active = 0;
    if (rstep == 0.0) {
      if (x[i] <= lb[i]) active = -1;
      if (x[i] >= ub[i]) active = 1;
    } else {
      double ld = x[i] - lb[i], ud = ub[i] - x[i];
      double lt = rstep * fmax(1.0, fabs(lb[i])), ut = rstep * fmax(1.0, fabs(ub[i]));
      if (is_finite(lb[i]) && ld <= fmin(ud, lt)) active = -1;
      if (is_finite(ub[i]) && ud <= fmin(ld, ut)) active = 1;
    }
    double xn = x[i];
    if (active == -1) xn = (rstep == 0.0) ? nextafter(lb[i], ub[i]) : lb[i] + rstep * fmax(1.0, fabs(lb[i]));
    if (active == 1) xn = (rstep == 0.0) ? nextafter(ub[i], lb[i]) : ub[i] - rstep * fmax(1.0, fabs(ub[i]));
    if (xn < lb[i] || xn > ub[i]) xn = 0.5 * (lb[i] + ub[i]);
    x[i] = xn;
  }
}

template <int NP>
MLRN_FN void cl_scaling(const double* x, const double* g, const double* lb, const double* ub, double* v, double* dv) {
  for (int i = 0; i < NP; ++i) {
    v[i] = 1.0; dv[i] = 0.0;
    if (g[i] < 0.0 && is_finite(ub[i])) { v[i] = ub[i] - x[i]; dv[i] = -1.0; }
    if (g[i] > 0.0 && is_finite(lb[i])) { v[i] = x[i] - lb[i]; dv[i] = 1.0; }
  }
}

// min step to a bound along s, and which variables hit (scipy step_size_to_bound)
template <int NP>
MLRN_NOINLINE double step_to_bound(const double* x, const double* s, const double* lb, const double* ub, int* hits) {
  double steps[NP]; double mn = pinf();
  for (int i = 0; i < NP; ++i) {
    steps[i] = pinf();
    if (s[i] != 0.0) steps[i] = fmax((lb[i] - x[i]) / s[i], (ub[i] - x[i]) / s[i]);
    mn = fmin(mn, steps[i]);
  }
  if (hits) for (int i = 0; i < NP; ++i) {
    int sg = (s[i] > 0.0) ? 1 : ((s[i] < 0.0) ? -1 : 0);
    hits[i] = (steps[i] == mn) ? sg : 0;
  }
  return mn;
}

MLRN_NOINLINE void minimize_quadratic_1d(double a, double b, double lo, double hi, double c, double& t_best, double& y_best) {
  double t[3] = {lo, hi, 0.0}; int nt = 2;
  if (a != 0.0) { double ex = -0.5 * b / a; if (lo < ex && ex < hi) { t[2] = ex; nt = 3; } }
  t_best = t[0]; y_best = t[0] * (a * t[0] + b) + c;
  for (int k = 1; k < nt; ++k) { double y = t[k] * (a * t[k] + b) + c; if (y < y_best) { y_best = y; t_best = t[k]; } }
}

// One-sided Jacobi SVD of an NP x NP matrix A (column-major a[c*NP + r]); returns singular values
// s (descending), right vectors V (v[c*NP + r]) and w = U^T b for a given b.
template <int NP>
MLRN_NOINLINE void small_svd(double* a, const double* b, double* s, double* V, double* w) {
  for (int i = 0; i < NP * NP; ++i) V[i] = 0.0;
  for (int i = 0; i < NP; ++i) V[i * NP + i] = 1.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll 1
    for (int p = 0; p < NP - 1; ++p)
#pragma unroll 1
      for (int q = p + 1; q < NP; ++q) {
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int r = 0; r < NP; ++r) { double x = a[p * NP + r], y = a[q * NP + r]; al += x * x; be += y * y; ga += x * y; }
        if (ga == 0.0 || fabs(ga) <= 1.1e-16 * sqrt(al * be)) continue;   // columns orthogonal to working precision
        rotated = true;
        double zeta = (be - al) / (2.0 * ga);
        double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int r = 0; r < NP; ++r) {
          double x = a[p * NP + r], y = a[q * NP + r];
          a[p * NP + r] = c * x - sn * y; a[q * NP + r] = sn * x + c * y;
          double vx = V[p * NP + r], vy = V[q * NP + r];
          V[p * NP + r] = c * vx - sn * vy; V[q * NP + r] = sn * vx + c * vy;
        }
      }
    if (!rotated) break;
  }
  double ub[NP];
  for (int c = 0; c < NP; ++c) {
    double nn = 0.0, d = 0.0;
    for (int r = 0; r < NP; ++r) { nn += a[c * NP + r] * a[c * NP + r]; d += a[c * NP + r] * b[r]; }
    s[c] = sqrt(nn);
    ub[c] = (s[c] > 0.0) ? d / s[c] : 0.0;
  }
  // selection sort, descending
  for (int i = 0; i < NP; ++i) {
    int k = i;
    for (int j = i + 1; j < NP; ++j) if (s[j] > s[k]) k = j;
    if (k != i) {
      double ts = s[i]; s[i] = s[k]; s[k] = ts;
      double tu = ub[i]; ub[i] = ub[k]; ub[k] = tu;
      for (int r = 0; r < NP; ++r) { double tv = V[i * NP + r]; V[i * NP + r] = V[k * NP + r]; V[k * NP + r] = tv; }
    }
  }
  for (int c = 0; c < NP; ++c) w[c] = ub[c];
}

// scipy solve_lsq_trust_region (More's iteration on the secular equation)
template <int NP>
MLRN_NOINLINE void solve_lsq_tr(int m, const double* uf, const double* s, const double* V, double Delta, double& alpha, double* p) {
  double suf[NP];
  for (int i = 0; i < NP; ++i) suf[i] = s[i] * uf[i];
  bool full_rank = false;
  if (m >= NP) full_rank = s[NP - 1] > TRF_EPS * (double)m * s[0];
  if (full_rank) {
    double q[NP];
    for (int i = 0; i < NP; ++i) q[i] = uf[i] / s[i];
    for (int r = 0; r < NP; ++r) { double a = 0.0; for (int c = 0; c < NP; ++c) a += V[c * NP + r] * q[c]; p[r] = -a; }
    if (vnorm<NP>(p) <= Delta) { alpha = 0.0; return; }
  }
  double alpha_upper = vnorm<NP>(suf) / Delta;
  double alpha_lower = 0.0;
  if (full_rank) {
    double pn = 0.0, dd = 0.0;
    for (int i = 0; i < NP; ++i) { double den = s[i] * s[i]; double r = suf[i] / den; pn += r * r; dd += suf[i] * suf[i] / (den * den * den); }
    pn = sqrt(pn);
    double phi = pn - Delta, phi_prime = -dd / pn;
    alpha_lower = -phi / phi_prime;
  }
  if ((!full_rank && alpha == 0.0))
    alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
  for (int it = 0; it < 10; ++it) {
    if (alpha < alpha_lower || alpha > alpha_upper) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    double pn = 0.0, dd = 0.0;
    for (int i = 0; i < NP; ++i) { double den = s[i] * s[i] + alpha; double r = suf[i] / den; pn += r * r; dd += suf[i] * suf[i] / (den * den * den); }
    pn = sqrt(pn);
    double phi = pn - Delta, phi_prime = -dd / pn;
    if (phi < 0.0) alpha_upper = alpha;
    double ratio = phi / phi_prime;
    alpha_lower = fmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (fabs(phi) < 0.01 * Delta) break;
  }
  double q[NP];
  for (int i = 0; i < NP; ++i) q[i] = suf[i] / (s[i] * s[i] + alpha);
  for (int r = 0; r < NP; ++r) { double a = 0.0; for (int c = 0; c < NP; ++c) a += V[c * NP + r] * q[c]; p[r] = -a; }
  double sc = Delta / vnorm<NP>(p);
  for (int r = 0; r < NP; ++r) p[r] *= sc;
}

}  // namespace mlrn
