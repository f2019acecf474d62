// trf.cuh -- bounded nonlinear least squares: Trust Region Reflective, dense "exact" sub-problem.
//
// The reference fits Bazin curves and decline models through scipy.optimize.curve_fit with
// bounds (bazin_fitting.py:128-137, scripts/train_v55_powerlaw.py:173-184), i.e. SciPy's
// least_squares(method='trf', jac='2-point', x_scale=1, ftol=xtol=gtol=1e-8, tr_solver='exact').
// SciPy is a third-party dependency of the reference (requirements.txt:22); this file restates
// that published algorithm (Branch, Coleman & Li 1999; More 1977 for the trust-region
// sub-problem) step for step so that the iterates -- and therefore the local minimum reached
// and the termination point -- track the reference's to rounding:
//   * forward-difference Jacobian with h = sqrt(eps) * sign(x) * max(1, |x|), flipped or
//     shortened at the bounds;
//   * Coleman-Li scaling d = sqrt(v), augmented Jacobian [J d; diag(sqrt(g dv))];
//   * trust-region step from the SVD of the augmented Jacobian (here: CGS2 QR across the team,
//     then a one-sided Jacobi SVD of the small R factor, redundantly on every member);
//   * reflected / constrained-Cauchy step selection, radius update, ftol/xtol/gtol tests.
// One team (a warp) owns one fit: residual / Jacobian rows are lane-strided, the NP x NP algebra
// is uniform across lanes.
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

template <int NP>
struct TrfWork {
  double* f;      // [m]   residuals at x
  double* fn;     // [m]   residuals at the trial point
  double* J;      // [NP][ld]  Jacobian, column-major
  double* Q;      // [NP][ld + NP]  augmented scaled Jacobian -> orthonormal factor
  int ld;
};

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
MLRN_FN void make_strictly_feasible(double* x, const double* lb, const double* ub, double rstep) {
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
MLRN_FN double step_to_bound(const double* x, const double* s, const double* lb, const double* ub, int* hits) {
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

MLRN_FN void minimize_quadratic_1d(double a, double b, double lo, double hi, double c, double& t_best, double& y_best) {
  double t[3] = {lo, hi, 0.0}; int nt = 2;
  if (a != 0.0) { double ex = -0.5 * b / a; if (lo < ex && ex < hi) { t[2] = ex; nt = 3; } }
  t_best = t[0]; y_best = t[0] * (a * t[0] + b) + c;
  for (int k = 1; k < nt; ++k) { double y = t[k] * (a * t[k] + b) + c; if (y < y_best) { y_best = y; t_best = t[k]; } }
}

// One-sided Jacobi SVD of an NP x NP matrix A (column-major a[c*NP + r]); returns singular values
// s (descending), right vectors V (v[c*NP + r]) and w = U^T b for a given b.
template <int NP>
MLRN_FN void small_svd(double* a, const double* b, double* s, double* V, double* w) {
  for (int i = 0; i < NP * NP; ++i) V[i] = 0.0;
  for (int i = 0; i < NP; ++i) V[i * NP + i] = 1.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < NP - 1; ++p)
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
MLRN_FN void solve_lsq_tr(int m, const double* uf, const double* s, const double* V, double Delta, double& alpha, double* p) {
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

// (J_h a) . (J_h b) with J_h = J diag(d), rows distributed over the team
template <int NP, class Team>
MLRN_FN double jh_dot(const Team& tm, const double* J, int ld, int m, const double* d, const double* a, const double* b) {
  double da[NP], db[NP];
  for (int c = 0; c < NP; ++c) { da[c] = d[c] * a[c]; db[c] = d[c] * b[c]; }
  double acc = 0.0;
  for (int i = tm.rank(); i < m; i += Team::size) {
    double u = 0.0, w = 0.0;
    for (int c = 0; c < NP; ++c) { double jv = J[c * ld + i]; u += jv * da[c]; w += jv * db[c]; }
    acc += u * w;
  }
  return tm.sum(acc);
}

template <int NP, class Team>
MLRN_FN double eval_quadratic(const Team& tm, const double* J, int ld, int m, const double* d, const double* diag_h,
                              const double* g_h, const double* s) {
  double q = jh_dot<NP>(tm, J, ld, m, d, s, s);
  for (int c = 0; c < NP; ++c) q += s[c] * diag_h[c] * s[c];
  return 0.5 * q + vdot<NP>(s, g_h);
}

// residuals at x -> out[0..m); returns 0.5 * ||out||^2, `finite` = all residuals finite
template <int NP, class Team, class Model>
MLRN_FN double eval_residuals(const Team& tm, const Model& mdl, int m, const double* x, double* out, bool& finite) {
  double acc = 0.0; bool bad = false;
  for (int i = tm.rank(); i < m; i += Team::size) {
    double r = mdl.resid(i, x);
    out[i] = r;
    acc += r * r;
    bad = bad || !is_finite(r);
  }
  finite = !tm.any(bad);
  tm.sync();
  return 0.5 * tm.sum(acc);
}

// forward-difference Jacobian with SciPy's bound-aware step (approx_derivative, '2-point')
template <int NP, class Team, class Model>
MLRN_FN bool fd_jacobian(const Team& tm, const Model& mdl, int m, const double* x, const double* lb, const double* ub,
                         const double* f0, double* J, int ld) {
  bool bad = false;
  for (int c = 0; c < NP; ++c) {
    double sgn = (x[c] >= 0.0) ? 1.0 : -1.0;
    double h = TRF_SQRT_EPS * sgn * fmax(1.0, fabs(x[c]));
    double lower = x[c] - lb[c], upper = ub[c] - x[c];
    double x1 = x[c] + h;
    bool violated = (x1 < lb[c]) || (x1 > ub[c]);
    bool fitting = fabs(h) <= fmax(lower, upper);
    if (violated && fitting) h = -h;
    if (!fitting) h = (upper >= lower) ? upper : -lower;
    double xp[NP];
    for (int k = 0; k < NP; ++k) xp[k] = x[k];
    xp[c] = x[c] + h;
    double dx = xp[c] - x[c];
    for (int i = tm.rank(); i < m; i += Team::size) {
      double jv = (mdl.resid(i, xp) - f0[i]) / dx;
      J[c * ld + i] = jv;
      bad = bad || !is_finite(jv);
    }
  }
  bool anybad = tm.any(bad);
  tm.sync();
  return !anybad;
}

// Returns SciPy's termination status (1 gtol, 2 ftol, 3 xtol, 4 both, 0 max_nfev) or a TrfFail code
// where curve_fit would have raised.  x: in = p0, out = solution.
template <int NP, class Team, class Model>
MLRN_FN int trf_bounded(const Team& tm, const Model& mdl, int m, double* x, const double* lb, const double* ub,
                        int max_nfev, const TrfWork<NP>& w, int& nfev, double& cost_out) {
  const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8;
  nfev = 0; cost_out = qnan();
  for (int i = 0; i < NP; ++i) if (!(lb[i] < ub[i])) return TRF_BAD_BOUNDS;
  if (!in_bounds<NP>(x, lb, ub)) return TRF_X0_INFEASIBLE;
  make_strictly_feasible<NP>(x, lb, ub, 1e-10);
  bool finite;
  double cost = eval_residuals<NP>(tm, mdl, m, x, w.f, finite);
  nfev = 1;
  if (!finite) return TRF_F0_NONFINITE;
  if (!fd_jacobian<NP>(tm, mdl, m, x, lb, ub, w.f, w.J, w.ld)) return TRF_J_NONFINITE;
  const int ldq = w.ld + NP;
  double g[NP], v[NP], dv[NP], d[NP], diag_h[NP], g_h[NP];
  auto grad = [&]() {
    for (int c = 0; c < NP; ++c) {
      double a = 0.0;
      for (int i = tm.rank(); i < m; i += Team::size) a += w.J[c * w.ld + i] * w.f[i];
      g[c] = tm.sum(a);
    }
  };
  grad();
  cl_scaling<NP>(x, g, lb, ub, v, dv);
  double Delta;
  { double a = 0.0; for (int i = 0; i < NP; ++i) { double q = x[i] / sqrt(v[i]); a += q * q; } Delta = sqrt(a); }
  if (Delta == 0.0) Delta = 1.0;
  double alpha = 0.0;
  int status = -1;                      // -1 = "None"
  double x_new[NP], step[NP], step_h[NP];
  for (;;) {
    cl_scaling<NP>(x, g, lb, ub, v, dv);
    double g_norm = 0.0;
    for (int i = 0; i < NP; ++i) g_norm = fmax(g_norm, fabs(g[i] * v[i]));
    if (g_norm < gtol) status = 1;
    TRF_LOG("iter nfev %d cost %.10e gnorm %.6e Delta %.6e x0 %.17g x1 %.17g\n", nfev, cost, g_norm, Delta, x[0], x[1]);
    if (status != -1 || nfev == max_nfev) break;
    for (int i = 0; i < NP; ++i) { d[i] = sqrt(v[i]); diag_h[i] = g[i] * dv[i]; g_h[i] = d[i] * g[i]; }
    // ---- augmented scaled Jacobian -> Q (CGS2), R; then SVD of R
    for (int c = 0; c < NP; ++c) {
      for (int i = tm.rank(); i < m; i += Team::size) w.Q[c * ldq + i] = w.J[c * w.ld + i] * d[c];
      for (int k = tm.rank(); k < NP; k += Team::size) w.Q[c * ldq + m + k] = (k == c) ? sqrt(diag_h[c]) : 0.0;
    }
    tm.sync();
    const int ma = m + NP;
    double R[NP * NP];                  // column-major R[c*NP + r], upper triangular
    for (int i = 0; i < NP * NP; ++i) R[i] = 0.0;
    for (int c = 0; c < NP; ++c) {
      for (int pass = 0; pass < 2; ++pass) {
        double proj[NP];
        for (int k = 0; k < c; ++k) {
          double a = 0.0;
          for (int i = tm.rank(); i < ma; i += Team::size) a += w.Q[k * ldq + i] * w.Q[c * ldq + i];
          proj[k] = tm.sum(a);
        }
        for (int i = tm.rank(); i < ma; i += Team::size) {
          double val = w.Q[c * ldq + i];
          for (int k = 0; k < c; ++k) val -= proj[k] * w.Q[k * ldq + i];
          w.Q[c * ldq + i] = val;
        }
        for (int k = 0; k < c; ++k) R[c * NP + k] += proj[k];
        tm.sync();
      }
      double a = 0.0;
      for (int i = tm.rank(); i < ma; i += Team::size) a += w.Q[c * ldq + i] * w.Q[c * ldq + i];
      double nn = sqrt(tm.sum(a));
      R[c * NP + c] = nn;
      double inv = (nn > 0.0) ? 1.0 / nn : 0.0;
      for (int i = tm.rank(); i < ma; i += Team::size) w.Q[c * ldq + i] *= inv;
      tm.sync();
    }
    double qtf[NP];
    for (int c = 0; c < NP; ++c) {
      double a = 0.0;
      for (int i = tm.rank(); i < m; i += Team::size) a += w.Q[c * ldq + i] * w.f[i];
      qtf[c] = tm.sum(a);
    }
    double sv[NP], V[NP * NP], uf[NP];
    small_svd<NP>(R, qtf, sv, V, uf);
    for (int i = 0; i < NP; ++i) if (!is_finite(sv[i])) return TRF_J_NONFINITE;

    const double theta = fmax(0.995, 1.0 - g_norm);
    double actual_reduction = -1.0, cost_new = cost, step_norm = 0.0;
    while (actual_reduction <= 0.0 && nfev < max_nfev) {
      double p_h[NP], p[NP];
      solve_lsq_tr<NP>(m, uf, sv, V, Delta, alpha, p_h);
      for (int i = 0; i < NP; ++i) p[i] = d[i] * p_h[i];
      // ---- select_step
      double predicted;
      {
        double xp[NP];
        for (int i = 0; i < NP; ++i) xp[i] = x[i] + p[i];
        if (in_bounds<NP>(xp, lb, ub)) {
          predicted = -eval_quadratic<NP>(tm, w.J, w.ld, m, d, diag_h, g_h, p_h);
          for (int i = 0; i < NP; ++i) { step[i] = p[i]; step_h[i] = p_h[i]; }
        } else {
          int hits[NP];
          double p_stride = step_to_bound<NP>(x, p, lb, ub, hits);
          double r_h[NP], r[NP], xb[NP];
          for (int i = 0; i < NP; ++i) { r_h[i] = hits[i] ? -p_h[i] : p_h[i]; r[i] = d[i] * r_h[i]; }
          for (int i = 0; i < NP; ++i) { p[i] *= p_stride; p_h[i] *= p_stride; xb[i] = x[i] + p[i]; }
          // intersect_trust_region(p_h, r_h, Delta): positive root
          double ia = vdot<NP>(r_h, r_h), ib = vdot<NP>(p_h, r_h), ic = vdot<NP>(p_h, p_h) - Delta * Delta;
          if (ia == 0.0 || ic > 0.0) return TRF_GEOMETRY;
          double disc = sqrt(ib * ib - ia * ic);
          double qq = -(ib + ((ib >= 0.0 && !(ib == 0.0 && 1.0 / ib < 0.0)) ? disc : -disc));   // copysign(disc, ib)
          double t1 = qq / ia, t2 = ic / qq;
          double to_tr = fmax(t1, t2);
          double to_bound = step_to_bound<NP>(xb, r, lb, ub, nullptr);
          double r_stride = fmin(to_bound, to_tr);
          double r_lo, r_hi;
          if (r_stride > 0.0) {
            r_lo = (1.0 - theta) * p_stride / r_stride;
            r_hi = (r_stride == to_bound) ? theta * to_bound : to_tr;
          } else { r_lo = 0.0; r_hi = -1.0; }
          double r_value = pinf();
          if (r_lo <= r_hi) {
            // build_quadratic_1d(J_h, g_h, r_h, s0 = p_h, diag = diag_h)
            double vv = jh_dot<NP>(tm, w.J, w.ld, m, d, r_h, r_h);
            double uv = jh_dot<NP>(tm, w.J, w.ld, m, d, p_h, r_h);
            double uu = jh_dot<NP>(tm, w.J, w.ld, m, d, p_h, p_h);
            double qa = vv, qb = vdot<NP>(g_h, r_h) + uv, qc = 0.5 * uu + vdot<NP>(g_h, p_h);
            for (int i = 0; i < NP; ++i) { qa += r_h[i] * diag_h[i] * r_h[i]; qb += p_h[i] * diag_h[i] * r_h[i]; qc += 0.5 * p_h[i] * diag_h[i] * p_h[i]; }
            qa *= 0.5;
            double ts, ys;
            minimize_quadratic_1d(qa, qb, r_lo, r_hi, qc, ts, ys);
            r_value = ys;
            for (int i = 0; i < NP; ++i) { r_h[i] = r_h[i] * ts + p_h[i]; r[i] = r_h[i] * d[i]; }
          }
          for (int i = 0; i < NP; ++i) { p[i] *= theta; p_h[i] *= theta; }
          double p_value = eval_quadratic<NP>(tm, w.J, w.ld, m, d, diag_h, g_h, p_h);
          double ag_h[NP], ag[NP];
          for (int i = 0; i < NP; ++i) { ag_h[i] = -g_h[i]; ag[i] = d[i] * ag_h[i]; }
          double to_tr2 = Delta / vnorm<NP>(ag_h);
          double to_bound2 = step_to_bound<NP>(x, ag, lb, ub, nullptr);
          double ag_stride = (to_bound2 < to_tr2) ? theta * to_bound2 : to_tr2;
          double aa = jh_dot<NP>(tm, w.J, w.ld, m, d, ag_h, ag_h);
          for (int i = 0; i < NP; ++i) aa += ag_h[i] * diag_h[i] * ag_h[i];
          aa *= 0.5;
          double ab = vdot<NP>(g_h, ag_h);
          double ts, ag_value;
          minimize_quadratic_1d(aa, ab, 0.0, ag_stride, 0.0, ts, ag_value);
          for (int i = 0; i < NP; ++i) { ag_h[i] *= ts; ag[i] *= ts; }
          if (p_value < r_value && p_value < ag_value) {
            predicted = -p_value; for (int i = 0; i < NP; ++i) { step[i] = p[i]; step_h[i] = p_h[i]; }
          } else if (r_value < p_value && r_value < ag_value) {
            predicted = -r_value; for (int i = 0; i < NP; ++i) { step[i] = r[i]; step_h[i] = r_h[i]; }
          } else {
            predicted = -ag_value; for (int i = 0; i < NP; ++i) { step[i] = ag[i]; step_h[i] = ag_h[i]; }
          }
        }
      }
      for (int i = 0; i < NP; ++i) x_new[i] = x[i] + step[i];
      make_strictly_feasible<NP>(x_new, lb, ub, 0.0);
      cost_new = eval_residuals<NP>(tm, mdl, m, x_new, w.fn, finite);
      ++nfev;
      double step_h_norm = vnorm<NP>(step_h);
      if (!finite) { Delta = 0.25 * step_h_norm; continue; }
      actual_reduction = cost - cost_new;
      TRF_LOG("   trial nfev %d cost_new %.10e ared %.6e pred %.6e stepn %.6e alpha %.6e\n", nfev, cost_new, actual_reduction, predicted, step_h_norm, alpha);
      // update_tr_radius
      double ratio;
      if (predicted > 0.0) ratio = actual_reduction / predicted;
      else if (predicted == 0.0 && actual_reduction == 0.0) ratio = 1.0;
      else ratio = 0.0;
      double Delta_new = Delta;
      if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
      else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
      step_norm = vnorm<NP>(step);
      // check_termination
      bool f_ok = (actual_reduction < ftol * cost) && (ratio > 0.25);
      bool x_ok = step_norm < xtol * (xtol + vnorm<NP>(x));
      if (f_ok && x_ok) status = 4; else if (f_ok) status = 2; else if (x_ok) status = 3;
      if (status != -1) break;
      alpha *= Delta / Delta_new;
      Delta = Delta_new;
    }
    if (actual_reduction > 0.0) {
      for (int i = 0; i < NP; ++i) x[i] = x_new[i];
      for (int i = tm.rank(); i < m; i += Team::size) w.f[i] = w.fn[i];
      tm.sync();
      cost = cost_new;
      if (!fd_jacobian<NP>(tm, mdl, m, x, lb, ub, w.f, w.J, w.ld)) return TRF_J_NONFINITE;
      grad();
    }
  }
  if (status == -1) status = 0;
  cost_out = cost;
  return status;
}

}  // namespace mlrn
