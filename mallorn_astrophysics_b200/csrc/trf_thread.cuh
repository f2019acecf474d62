// trf_thread.cuh -- the same bounded Trust-Region-Reflective solver as trf.cuh (SciPy least_squares
// 'trf', 2-point Jacobian, exact sub-problem), laid out for ONE THREAD PER FIT with no per-fit arrays:
//
//   * residuals and the forward-difference Jacobian are produced row by row and folded into the NP x NP
//     triangular factor immediately (Givens rotations, "row-streaming QR"), together with Q^T f; J and f
//     are never stored.  J = Q R0  =>  J_h = J D = Q (R0 D), so Coleman-Li scaling is a column scaling of
//     R0 applied afterwards, the augmented rows diag(sqrt(g dv)) are rotated in, and every later product
//     with J_h ((J_h a).(J_h b) in the step selection) is a product with the small factor;
//   * the solver is a resumable state machine: step() performs at most one Jacobian pass and exactly one
//     trial evaluation (one `nfev`), so the lanes of a warp -- each on its own fit -- stay in lock step
//     and a lane that finishes can be handed the next fit at once (k_*_thread kernels in fits.cu).
//
// Step control (radius update, reflected / Cauchy step selection, ftol/xtol/gtol tests, max_nfev) is the
// code of trf.cuh, called on the small factor.  Scalar code: it also compiles for tests/hostsim.
#pragma once
#include "trf.cuh"

namespace mlrn {

template <int NP>
struct SmallFactor {            // R0: upper triangular, column-major r[c*NP + k] (k <= c); q = Q^T f
  double r[NP * NP];
  double q[NP];
  MLRN_FN void clear() { for (int i = 0; i < NP * NP; ++i) r[i] = 0.0; for (int i = 0; i < NP; ++i) q[i] = 0.0; }
  // fold the row (v, rhs) in; v is destroyed
  MLRN_NOINLINE void add_row(double* v, double rhs) { fold(v, rhs); }
  // The same, inlined at the call site: with a factor that is a local object of the caller every index is a compile-time
  // constant and the 15 + NP live values stay in registers across the row loop (jacobian_pass).  The multiply-adds are
  // written out as fma() so that the rounding does not depend on the context the compiler contracts them in (inlined
  // and out of line, device and host give the same bits).
  MLRN_FN_FORCE void fold(double* v, double rhs) {
#pragma unroll
    for (int c = 0; c < NP; ++c) {
      const double b = v[c];
      if (b != 0.0) {
        const double a = r[c * NP + c];
        const double h2 = fma(a, a, b * b);
        const double rh = rsqrt(h2);                 // one MUFU + Newton step instead of sqrt and two divisions
        const double cs = a * rh, sn = b * rh;
        r[c * NP + c] = h2 * rh;
#pragma unroll
        for (int k = c + 1; k < NP; ++k) {
          const double rk = r[k * NP + c], vk = v[k];
          r[k * NP + c] = fma(cs, rk, sn * vk);
          v[k] = fma(cs, vk, -(sn * rk));
        }
        const double qc = q[c];
        q[c] = fma(cs, qc, sn * rhs);
        rhs = fma(cs, rhs, -(sn * qc));
      }
    }
  }
};

enum { TRF_RUNNING = 100 };

// forward-difference row of a model without a specialised fd_row: one full evaluation per parameter.  The per-vector
// preparations of the model (reciprocals of its divisors, Model::Prep1) are taken once per pass for x and for each of the
// NP perturbed vectors.
template <int NP, class Model> struct GenericPrep { typename Model::Prep1 base, pert[NP]; };
template <int NP, class Model>
MLRN_FN GenericPrep<NP, Model> generic_prepare(const Model& mdl, const double* x, const double* xpd) {
  GenericPrep<NP, Model> g;
  g.base = mdl.prepare1(x);
  double xp[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) xp[k] = x[k];
#pragma unroll
  for (int c = 0; c < NP; ++c) { xp[c] = xpd[c]; g.pert[c] = mdl.prepare1(xp); xp[c] = x[c]; }
  return g;
}
template <int NP, class Model>
MLRN_FN double generic_fd_row(const Model& mdl, int i, const double* x, const double* xpd, const double* dx, const double* rdx,
                              const GenericPrep<NP, Model>& p, double* row) {
  const double r0 = mdl.resid(i, x, p.base);
  double xp[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) xp[k] = x[k];
#pragma unroll
  for (int c = 0; c < NP; ++c) {
    xp[c] = xpd[c];
    row[c] = div_r(mdl.resid(i, xp, p.pert[c]) - r0, dx[c], rdx[c]);
    xp[c] = x[c];
  }
  return r0;
}

// COOP = false: one thread per fit.  COOP = true (device only): the 32 lanes of a warp hold the SAME fit -- identical
// state in every lane, uniform control flow -- and share out the only part that is per-row and expensive, the model
// evaluations: lane l evaluates rows l, l+32, ...; the values are then broadcast and folded into the cost / the triangular
// factor by every lane in row order, i.e. with exactly the arithmetic of the one-thread solver (bit-identical results).
// Used for the stragglers of the staged fit kernels (fits.cu), whose trips are dominated by 40-70 rows x 6 exp().
template <int NP, class Model, bool COOP = false>
struct TrfThread {
  // problem
  Model mdl; int m, max_nfev;
  double x[NP], lb[NP], ub[NP];
  // solver state
  double cost, Delta, alpha, g[NP];
  SmallFactor<NP> F0;                 // factor of the unscaled Jacobian at x
  double d[NP], diag_h[NP], g_h[NP], sv[NP], V[NP * NP], uf[NP], theta, g_norm;
  int nfev, status;
  bool need_jac;

  MLRN_NOINLINE double residual_cost(const double* xx, bool& finite) const {
    double acc = 0.0; bool bad = false;
    const typename Model::Prep1 pr = mdl.prepare1(xx);
#ifdef __CUDA_ARCH__
    if (COOP) {
      const int lane = threadIdx.x & 31;
      for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        const double mine = (i < m) ? mdl.resid(i, xx, pr) : 0.0;
        const int cnt = (m - base < 32) ? m - base : 32;
        for (int j = 0; j < cnt; ++j) { const double r = __shfl_sync(0xffffffffu, mine, j); acc += r * r; bad = bad || !is_finite(r); }
      }
      finite = !bad;
      return 0.5 * acc;
    }
#endif
    for (int i = 0; i < m; ++i) { const double r = mdl.resid(i, xx, pr); acc += r * r; bad = bad || !is_finite(r); }
    finite = !bad;
    return 0.5 * acc;
  }

  // (J_h a) . (J_h b) with J_h = Q R0 D
  MLRN_NOINLINE double jh_dot(const double* a, const double* b) const {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      double u = 0.0, w = 0.0;
#pragma unroll
      for (int c = k; c < NP; ++c) { const double rc = F0.r[c * NP + k]; u += rc * (d[c] * a[c]); w += rc * (d[c] * b[c]); }
      acc += u * w;
    }
    return acc;
  }
  MLRN_NOINLINE double eval_quadratic(const double* s) const {
    double q = jh_dot(s, s);
    for (int c = 0; c < NP; ++c) q += s[c] * diag_h[c] * s[c];
    return 0.5 * q + vdot<NP>(s, g_h);
  }

  // Returns TRF_RUNNING or a final status / TrfFail code.
  // cost_in: the cost at the (strictly feasible) start computed earlier by the prologue pass of fits.cu -- NaN stands for
  // a non-finite residual -- instead of a pass over the rows here; cost_out: where that pass leaves it
  MLRN_FN int begin(const Model& model, int rows, const double* x0, const double* lo, const double* hi, int maxfev,
                    const double* cost_in = nullptr, double* cost_out = nullptr) {
    mdl = model; m = rows; max_nfev = maxfev; nfev = 0; status = -1; cost = qnan();
    for (int i = 0; i < NP; ++i) { x[i] = x0[i]; lb[i] = lo[i]; ub[i] = hi[i]; }
    for (int i = 0; i < NP; ++i) if (!(lb[i] < ub[i])) return TRF_BAD_BOUNDS;
    if (!in_bounds<NP>(x, lb, ub)) return TRF_X0_INFEASIBLE;
    make_strictly_feasible<NP>(x, lb, ub, 1e-10);
    bool finite;
    if (cost_in) { cost = *cost_in; finite = !is_nan(cost); }
    else cost = residual_cost(x, finite);
    if (cost_out) *cost_out = finite ? cost : qnan();
    nfev = 1;
    if (!finite) return TRF_F0_NONFINITE;
    need_jac = true; alpha = 0.0; Delta = -1.0;     // Delta < 0: initialise after the first gradient
    return TRF_RUNNING;
  }

  // residual + forward-difference Jacobian at x, streamed into F0; g = J^T f
  // forward-difference steps at x (scipy approx_derivative, '2-point' with bounds): the perturbed value of parameter c and
  // the exactly representable step
  MLRN_FN void perturbation(double* xpd, double* dx) const {
    for (int c = 0; c < NP; ++c) {
      const double sgn = (x[c] >= 0.0) ? 1.0 : -1.0;
      double h = TRF_SQRT_EPS * sgn * fmax(1.0, fabs(x[c]));
      const double lower = x[c] - lb[c], upper = ub[c] - x[c];
      const double x1 = x[c] + h;
      const bool violated = (x1 < lb[c]) || (x1 > ub[c]);
      const bool fitting = fabs(h) <= fmax(lower, upper);
      if (violated && fitting) h = -h;
      if (!fitting) h = (upper >= lower) ? upper : -lower;
      xpd[c] = x[c] + h;
      dx[c] = xpd[c] - x[c];
    }
  }
  MLRN_FN void store_factor(const SmallFactor<NP>& F) {
#pragma unroll
    for (int c = 0; c < NP; ++c) {
#pragma unroll
      for (int k = 0; k < NP; ++k) F0.r[c * NP + k] = (k <= c) ? F.r[c * NP + k] : 0.0;
      F0.q[c] = F.q[c];
    }
  }
  MLRN_FN void gradient() {
#pragma unroll
    for (int c = 0; c < NP; ++c) { double a = 0.0; for (int k = 0; k <= c; ++k) a += F0.r[c * NP + k] * F0.q[k]; g[c] = a; }
  }

  MLRN_NOINLINE bool jacobian_pass() {
    double xpd[NP], dx[NP], rdx[NP];
    perturbation(xpd, dx);
    for (int c = 0; c < NP; ++c) rdx[c] = 1.0 / dx[c];          // the steps are the same for every row: see div_r (fits.cuh)
    const typename Model::Prep prep = mdl.prepare(x, xpd);
#if !defined(TRF_LOCAL_FACTOR) && defined(__CUDA_ARCH__)
    // The factor lives in registers while the rows stream in (the noinline add_row pinned it in local memory: 40 local
    // loads/stores per row, most of the pass's memory traffic: Bazin 282 k -> 391 k objects/s at 10^6 objects together with
    // the work order of fits.cu); stored to the solver state once, after the last row.
    SmallFactor<NP> F;
    F.clear();
#define TRF_FOLD(row, rhs) F.fold(row, rhs)
#else
    SmallFactor<NP>& F = F0;
    F.clear();
#define TRF_FOLD(row, rhs) F.add_row(row, rhs)
#endif
    bool bad = false;
#ifdef __CUDA_ARCH__
    if (COOP) {
      const int lane = threadIdx.x & 31;
      for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        double mine[NP], r_mine = 0.0;
#pragma unroll
        for (int c = 0; c < NP; ++c) mine[c] = 0.0;
        if (i < m) r_mine = mdl.fd_row(i, x, xpd, dx, rdx, prep, mine);
        const int cnt = (m - base < 32) ? m - base : 32;
        for (int j = 0; j < cnt; ++j) {
          double row[NP];
#pragma unroll
          for (int c = 0; c < NP; ++c) { row[c] = __shfl_sync(0xffffffffu, mine[c], j); bad = bad || !is_finite(row[c]); }
          TRF_FOLD(row, __shfl_sync(0xffffffffu, r_mine, j));
        }
      }
    } else
#endif
    for (int i = 0; i < m; ++i) {
      double row[NP];
      const double r0 = mdl.fd_row(i, x, xpd, dx, rdx, prep, row);
#pragma unroll
      for (int c = 0; c < NP; ++c) bad = bad || !is_finite(row[c]);
      TRF_FOLD(row, r0);
    }
#undef TRF_FOLD
#if !defined(TRF_LOCAL_FACTOR) && defined(__CUDA_ARCH__)
    store_factor(F);
#endif
    if (bad) return false;
    gradient();
    return true;
  }

  // One trip: (Jacobian + factorisation if a step was accepted) + one trial step.  Returns TRF_RUNNING or final.
  MLRN_FN int step() {
    const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8;
    double v[NP], dv[NP];
    if (need_jac) {
      if (!jacobian_pass()) return TRF_J_NONFINITE;
      cl_scaling<NP>(x, g, lb, ub, v, dv);
      if (Delta < 0.0) {
        double a = 0.0; for (int i = 0; i < NP; ++i) { const double q = x[i] / sqrt(v[i]); a += q * q; }
        Delta = sqrt(a); if (Delta == 0.0) Delta = 1.0;
      }
      g_norm = 0.0;
      for (int i = 0; i < NP; ++i) g_norm = fmax(g_norm, fabs(g[i] * v[i]));
      if (g_norm < gtol) return 1;
      if (nfev == max_nfev) return 0;
      for (int i = 0; i < NP; ++i) { d[i] = sqrt(v[i]); diag_h[i] = g[i] * dv[i]; g_h[i] = d[i] * g[i]; }
      // augmented factor: [R0 D; diag(sqrt(diag_h))] -> R, Q^T [f; 0]
      SmallFactor<NP> Fa;
      for (int c = 0; c < NP; ++c) for (int k = 0; k < NP; ++k) Fa.r[c * NP + k] = (k <= c) ? F0.r[c * NP + k] * d[c] : 0.0;
      for (int k = 0; k < NP; ++k) Fa.q[k] = F0.q[k];
      for (int c = 0; c < NP; ++c) if (diag_h[c] != 0.0) {
        double row[NP];
        for (int k = 0; k < NP; ++k) row[k] = (k == c) ? sqrt(diag_h[c]) : 0.0;
        Fa.add_row(row, 0.0);
      }
      small_svd<NP>(Fa.r, Fa.q, sv, V, uf);
      for (int i = 0; i < NP; ++i) if (!is_finite(sv[i])) return TRF_J_NONFINITE;
      theta = fmax(0.995, 1.0 - g_norm);
      need_jac = false;
    } else if (nfev >= max_nfev) {
      return 0;
    }
    // ---- one trial
    double p_h[NP], p[NP], step_[NP], step_h[NP], x_new[NP], predicted;
    solve_lsq_tr<NP>(m, uf, sv, V, Delta, alpha, p_h);
    for (int i = 0; i < NP; ++i) p[i] = d[i] * p_h[i];
    {
      double xq[NP];
      for (int i = 0; i < NP; ++i) xq[i] = x[i] + p[i];
      if (in_bounds<NP>(xq, lb, ub)) {
        predicted = -eval_quadratic(p_h);
        for (int i = 0; i < NP; ++i) { step_[i] = p[i]; step_h[i] = p_h[i]; }
      } else {
        int hits[NP];
        const double p_stride = step_to_bound<NP>(x, p, lb, ub, hits);
        double r_h[NP], r[NP], xb[NP];
        for (int i = 0; i < NP; ++i) { r_h[i] = hits[i] ? -p_h[i] : p_h[i]; r[i] = d[i] * r_h[i]; }
        for (int i = 0; i < NP; ++i) { p[i] *= p_stride; p_h[i] *= p_stride; xb[i] = x[i] + p[i]; }
        const double ia = vdot<NP>(r_h, r_h), ib = vdot<NP>(p_h, r_h), ic = vdot<NP>(p_h, p_h) - Delta * Delta;
        if (ia == 0.0 || ic > 0.0) return TRF_GEOMETRY;
        const double disc = sqrt(ib * ib - ia * ic);
        const double qq = -(ib + ((ib >= 0.0 && !(ib == 0.0 && 1.0 / ib < 0.0)) ? disc : -disc));
        const double t1 = qq / ia, t2 = ic / qq;
        const double to_tr = fmax(t1, t2);
        const double to_bound = step_to_bound<NP>(xb, r, lb, ub, nullptr);
        const double r_stride = fmin(to_bound, to_tr);
        double r_lo, r_hi;
        if (r_stride > 0.0) { r_lo = (1.0 - theta) * p_stride / r_stride; r_hi = (r_stride == to_bound) ? theta * to_bound : to_tr; }
        else { r_lo = 0.0; r_hi = -1.0; }
        double r_value = pinf();
        if (r_lo <= r_hi) {
          const double vv = jh_dot(r_h, r_h), uv = jh_dot(p_h, r_h), uu = jh_dot(p_h, p_h);
          double qa = vv, qb = vdot<NP>(g_h, r_h) + uv, qc = 0.5 * uu + vdot<NP>(g_h, p_h);
          for (int i = 0; i < NP; ++i) { qa += r_h[i] * diag_h[i] * r_h[i]; qb += p_h[i] * diag_h[i] * r_h[i]; qc += 0.5 * p_h[i] * diag_h[i] * p_h[i]; }
          qa *= 0.5;
          double ts, ys;
          minimize_quadratic_1d(qa, qb, r_lo, r_hi, qc, ts, ys);
          r_value = ys;
          for (int i = 0; i < NP; ++i) { r_h[i] = r_h[i] * ts + p_h[i]; r[i] = r_h[i] * d[i]; }
        }
        for (int i = 0; i < NP; ++i) { p[i] *= theta; p_h[i] *= theta; }
        const double p_value = eval_quadratic(p_h);
        double ag_h[NP], ag[NP];
        for (int i = 0; i < NP; ++i) { ag_h[i] = -g_h[i]; ag[i] = d[i] * ag_h[i]; }
        const double to_tr2 = Delta / vnorm<NP>(ag_h);
        const double to_bound2 = step_to_bound<NP>(x, ag, lb, ub, nullptr);
        const double ag_stride = (to_bound2 < to_tr2) ? theta * to_bound2 : to_tr2;
        double aa = jh_dot(ag_h, ag_h);
        for (int i = 0; i < NP; ++i) aa += ag_h[i] * diag_h[i] * ag_h[i];
        aa *= 0.5;
        const double ab = vdot<NP>(g_h, ag_h);
        double ts, ag_value;
        minimize_quadratic_1d(aa, ab, 0.0, ag_stride, 0.0, ts, ag_value);
        for (int i = 0; i < NP; ++i) { ag_h[i] *= ts; ag[i] *= ts; }
        if (p_value < r_value && p_value < ag_value) { predicted = -p_value; for (int i = 0; i < NP; ++i) { step_[i] = p[i]; step_h[i] = p_h[i]; } }
        else if (r_value < p_value && r_value < ag_value) { predicted = -r_value; for (int i = 0; i < NP; ++i) { step_[i] = r[i]; step_h[i] = r_h[i]; } }
        else { predicted = -ag_value; for (int i = 0; i < NP; ++i) { step_[i] = ag[i]; step_h[i] = ag_h[i]; } }
      }
    }
    for (int i = 0; i < NP; ++i) x_new[i] = x[i] + step_[i];
    make_strictly_feasible<NP>(x_new, lb, ub, 0.0);
    bool finite;
    const double cost_new = residual_cost(x_new, finite);
    ++nfev;
    const double step_h_norm = vnorm<NP>(step_h);
    if (!finite) { Delta = 0.25 * step_h_norm; return TRF_RUNNING; }
    const double actual_reduction = cost - cost_new;
    double ratio;
    if (predicted > 0.0) ratio = actual_reduction / predicted;
    else if (predicted == 0.0 && actual_reduction == 0.0) ratio = 1.0;
    else ratio = 0.0;
    double Delta_new = Delta;
    if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
    else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
    const double step_norm = vnorm<NP>(step_);
    const bool f_ok = (actual_reduction < ftol * cost) && (ratio > 0.25);
    const bool x_ok = step_norm < xtol * (xtol + vnorm<NP>(x));
    int term = -1;
    if (f_ok && x_ok) term = 4; else if (f_ok) term = 2; else if (x_ok) term = 3;
    if (term == -1) { alpha *= Delta / Delta_new; Delta = Delta_new; }
    if (actual_reduction > 0.0) {
      for (int i = 0; i < NP; ++i) x[i] = x_new[i];
      cost = cost_new;
      need_jac = true;
      if (term != -1) {
        // SciPy evaluates the Jacobian at the accepted point before it returns; curve_fit then takes its SVD
        // and raises on a non-finite entry
        if (!jacobian_pass()) return TRF_J_NONFINITE;
      }
    }
    if (term != -1) return term;
    return TRF_RUNNING;
  }
};

}  // namespace mlrn
