// gp1d.cuh -- per-band 1-D Gaussian-process features (reference: src/features/gaussian_process.py:29-248,
// served by sklearn's GaussianProcessRegressor there).  One team (a warp on the GPU) fits one band at a time.
//
// Model (gaussian_process.py:88-104): times scaled to [0, 1], fluxes standardised,
//   K = c exp(-(x_i - x_j)^2 / (2 l^2)) + (s + alpha_i) delta_ij,   alpha_i = max((err_i / f_std)^2, 1e-10),
//   theta = ln [c, l, s] inside ln [(0.01, 100), (0.01, 2), (1e-5, 10)],
// maximised by four L-BFGS-B starts (lbfgsb.cuh): the kernel's own theta and three fixed restart points
// (sklearn draws them from RandomState(42), re-seeded in every fit, so they are constants).
//
// Evaluation of -ln L and its gradient: the (n+1) x (n+1) symmetric matrix [[K, y], [y', 0]] is held as a
// packed lower triangle in team scratch and its first n pivots are swept in place (symmetric Gauss-Jordan
// "sweep" operator): afterwards the leading block is -K^-1, the last row is a = K^-1 y, the corner is
// -y'K^-1 y and the product of the pivots is |K|.  One rank-1 update of the whole triangle per pivot,
// the same work for every lane, no triangular solves; n^3/2 FMA like Cholesky + inverse.  The gradient
// 1/2 tr((a a' - K^-1) dK/dtheta_j) is then one more pass over the triangle with the kernel re-evaluated.
#pragma once
#include "team.cuh"
#include "objview.cuh"
#include "lbfgsb.cuh"

namespace mlrn {

constexpr int GP1D_COLS = 21;
constexpr int GP1D_STARTS = 4;
constexpr int GP1D_STATUS = 12;        // per object: per band {evaluations (all starts), iterations, winner start}

MLRN_FN_FORCE int tri_size(int n) { return n * (n + 1) / 2; }          // packed lower triangle of an n x n matrix
// scratch doubles one team needs for a band of n valid points: matrix of order n+1, x, y, alpha, v, w, pivots
MLRN_HD long long gp1d_scratch_doubles(int n) { return (long long)(n + 1) * (n + 2) / 2 + 6LL * (n + 1); }
// Longest band the per-band GP fits: the packed triangle is indexed with 32-bit integers (tri_size(n + 1) must fit) and
// a 32,768-point band already needs a 4.3 GB matrix (sklearn's dense one would be 8.6 GB).  Longer bands are reported
// like bands with fewer than 5 points: not fitted, their columns NaN (documented deviation; survey bands have ~25 points).
constexpr int GP1D_MAX_POINTS = 32768;

struct Gp1dScratch {
  double* A;        // packed lower triangle, order n+1
  double* x; double* y; double* al;     // [n]
  double* v; double* w;                 // [n+1]
  double* piv;                          // [n] pivots of the sweep
};
MLRN_FN Gp1dScratch gp1d_carve(double* base, int ncap) {
  Gp1dScratch s;
  s.x = base; s.y = s.x + (ncap + 1); s.al = s.y + (ncap + 1); s.v = s.al + (ncap + 1); s.w = s.v + (ncap + 1);
  s.piv = s.w + (ncap + 1);
  s.A = s.piv + (ncap + 1);
  return s;
}

// lane-strided walk over the packed triangle: element e = i(i+1)/2 + j, j <= i
struct TriWalk {
  int i, j, e;
  MLRN_FN_FORCE void begin(int first) { e = first; i = 0; j = first; fix(); }
  MLRN_FN_FORCE void step(int stride) { e += stride; j += stride; fix(); }
  MLRN_FN_FORCE void fix() { while (j > i) { j -= i + 1; ++i; } }
};

// -ln L and -gradient at theta (sklearn _gpr.py log_marginal_likelihood).  Returns false if a pivot is not positive
// (sklearn: LinAlgError -> objective +inf, zero gradient).
template <class Team>
MLRN_FN bool gp1d_objective(const Team& tm, int n, const Gp1dScratch& s, const double* theta, double& fval, double* grad) {
  const double c = exp(theta[0]), l = exp(theta[1]), sn = exp(theta[2]);
  const double il2 = 1.0 / (l * l);
  const int n1 = n + 1, total = tri_size(n1);
  double* __restrict__ A = s.A;
  double* __restrict__ v = s.v;
  double* __restrict__ w = s.w;
  double* __restrict__ piv = s.piv;
  const double* __restrict__ x = s.x;
  TriWalk wk;
  for (wk.begin(tm.rank()); wk.e < total; wk.step(Team::size)) {
    double val;
    if (wk.i == n) val = (wk.j == n) ? 0.0 : s.y[wk.j];
    else {
      const double d = x[wk.i] - x[wk.j];
      val = c * exp(-0.5 * (d * d * il2));
      if (wk.i == wk.j) val += sn + s.al[wk.i];
    }
    A[wk.e] = val;
  }
  tm.sync();
  // Sweep.  Lane = column: in row i the lanes j <= i update A(i, j) -= v_i w_j (consecutive addresses, v_i broadcast,
  // w_j in a register; predicated, no branches).  The lane that owns the pivot column computes throw-away values there;
  // a short lane-parallel pass then writes that column (w_i) and collects the next pivot column.
  double* __restrict__ vc = v;
  double* __restrict__ vn = w;
  for (int i = tm.rank(); i < n1; i += Team::size) vc[i] = A[tri_size(i)];        // column 0
  tm.sync();
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    const double pv = vc[k];
    if (!(pv > 0.0)) { ok = false; break; }
    const double ip = 1.0 / pv;
    if (tm.rank() == 0) piv[k] = pv;
    const int kn = k + 1;
    for (int jb = 0; jb < n1; jb += Team::size) {
      const int j = jb + tm.rank();
      if (j < n1) {
        const double wj = vc[j] * ip;
        int e = tri_size(jb) + j, i = jb;
#pragma unroll 4
        for (; i < k; ++i) { if (j <= i) A[e] = A[e] - vc[i] * wj; e += i + 1; }
        if (i == k) { if (j <= i) A[e] = (j == k) ? -ip : wj; e += i + 1; ++i; }        // pivot row: w, and -1/pivot
#pragma unroll 4
        for (; i < n1; ++i) { if (j <= i) A[e] = A[e] - vc[i] * wj; e += i + 1; }
      }
    }
    tm.sync();
    for (int i = tm.rank(); i < n1; i += Team::size) {
      if (i > k) A[tri_size(i) + k] = vc[i] * ip;                                     // pivot column below the diagonal: w_i
      // next pivot column: row k+1 up to the diagonal, column k+1 below it; its entry in column k is the w just defined
      vn[i] = (i == k) ? vc[kn] * ip : ((i <= kn) ? A[tri_size(kn) + i] : A[tri_size(i) + kn]);
    }
    tm.sync();
    double* tmp = vc; vc = vn; vn = tmp;
  }
  if (!ok) { fval = pinf(); grad[0] = grad[1] = grad[2] = 0.0; tm.sync(); return false; }
  const double* __restrict__ a = A + tri_size(n);              // row n: K^-1 y
  const double yKy = -a[n];
  double g0 = 0.0, g1 = 0.0, g2 = 0.0, logdet = 0.0;
  for (int i = tm.rank(); i < n; i += Team::size) logdet += log(piv[i]);
  const int tn = tri_size(n);
  for (wk.begin(tm.rank()); wk.e < tn; wk.step(Team::size)) {
    const double W = a[wk.i] * a[wk.j] + A[wk.e];          // (a a' - K^-1)_ij
    if (wk.i == wk.j) { g0 += W * c; g2 += W; }
    else {
      const double d = x[wk.i] - x[wk.j];
      const double q = d * d * il2;
      const double C = c * exp(-0.5 * q);
      g0 += 2.0 * W * C; g1 += 2.0 * W * C * q;
    }
  }
  g0 = tm.sum(g0); g1 = tm.sum(g1); g2 = tm.sum(g2); logdet = tm.sum(logdet);
  const double lml = -0.5 * yKy - 0.5 * logdet - 0.5 * (double)n * 1.8378770664093453;   // ln(2 pi)
  fval = -lml;
  grad[0] = -0.5 * g0; grad[1] = -0.5 * g1; grad[2] = -0.5 * sn * g2;
  tm.sync();
  return true;
}

struct Gp1dBandResult { double length_scale, amplitude, noise, loglik; int nfev, nit, winner; };

// Band preparation (gaussian_process.py:54-85): validity mask, time scaling to [0, 1], flux standardisation, alpha.
// gt/gf/ge: the band's time-sorted segment (n_raw rows, global memory).  Fills x, y, al [n]; returns n (0 = no fit).
template <class Team>
MLRN_FN int gp1d_prepare(const Team& tm, const double* gt, const double* gf, const double* ge, int n_raw,
                         double* sx, double* sy, double* sal, double& t_range, double& f_std) {
  if (n_raw < 5 || n_raw > GP1D_MAX_POINTS) return 0;
  // keep rows with finite flux, finite error, error > 0 (:58); order preserved
  int n = 0;
  for (int base = 0; base < n_raw; base += Team::size) {
    const int i = base + tm.rank();
    double tt = 0.0, ff = 0.0, ee = 0.0; bool keep = false;
    if (i < n_raw) { tt = gt[i]; ff = gf[i]; ee = ge[i]; keep = !(is_nan(ff) || is_nan(ee) || ee <= 0.0); }
    const int slot = tm.compact_slot(keep, n);
    if (keep) { sx[slot] = tt; sy[slot] = ff; sal[slot] = ee; }
  }
  tm.sync();
  if (n < 5) return 0;
  const double t_min = team_min(tm, sx, n), t_max = team_max(tm, sx, n);
  t_range = t_max - t_min;
  if (t_range == 0.0 || is_nan(t_range)) return 0;
  const double f_mean = team_mean(tm, sy, n);
  f_std = sqrt(team_var(tm, sy, n, f_mean));
  if (f_std == 0.0) f_std = 1.0;
  for (int i = tm.rank(); i < n; i += Team::size) {
    sx[i] = (sx[i] - t_min) / t_range;
    sy[i] = (sy[i] - f_mean) / f_std;
    const double en = sal[i] / f_std;
    sal[i] = fmax(en * en, 1e-10);
  }
  tm.sync();
  return n;
}

// The four optimiser runs of one prepared band and the reference's four numbers (gaussian_process.py:88-146).
// objective(theta, f, g): -ln L and its gradient, identical on every member; opt: optimiser state owned by rank 0;
// ctl: 8 doubles of team-visible scratch.
template <class Team, class Objective>
MLRN_FN void gp1d_optimise(const Team& tm, Objective&& objective, double t_range, double f_std, LbfgsB<3>* opt, double* ctl,
                           Gp1dBandResult& out) {
  const double LO[3] = {-4.605170185988091, -4.605170185988091, -11.512925464970229};      // ln 0.01, ln 0.01, ln 1e-5
  const double HI[3] = {4.605170185988092, 0.6931471805599453, 2.302585092994046};         // ln 100, ln 2, ln 10
  const double START[GP1D_STARTS][3] = {
      {0.0, -1.6094379124341003, -2.3025850929940455},                                      // ln [1, 0.2, 0.1]
      {-0x1.27d0b258dec3ep+0, 0x1.ba6262947e1f0p-2, -0x1.666a086b81fa0p+0},                 // RandomState(42).uniform(lo, hi) x 3
      {0x1.d13e458b4d080p-1, -0x1.e3a6ffd15b375p+1, -0x1.2b72f23c29e62p+3},
      {-0x1.047e29999566cp+2, -0x1.04689013a2800p-6, -0x1.9aa6c74700fb8p+1}};
  double best_f = pinf(), best_x[3] = {qnan(), qnan(), qnan()};
  for (int st = 0; st < GP1D_STARTS; ++st) {
    if (tm.rank() == 0) { opt->start(START[st], LO, HI); ctl[0] = opt->x[0]; ctl[1] = opt->x[1]; ctl[2] = opt->x[2]; }
    tm.sync();
    // SciPy's ScalarFunction neither re-evaluates nor counts a trial point equal to the previous one (the line search
    // falls back to its best step when the bracket collapses); skipping it saves a whole sweep.
    double prev[3] = {qnan(), qnan(), qnan()}, fv = 0.0, g[3] = {0.0, 0.0, 0.0};
    int evals = 0;
    for (;;) {
      const double theta[3] = {ctl[0], ctl[1], ctl[2]};
      if (!(theta[0] == prev[0] && theta[1] == prev[1] && theta[2] == prev[2])) {
        objective(theta, fv, g);
        ++evals;
        prev[0] = theta[0]; prev[1] = theta[1]; prev[2] = theta[2];
      }
      if (tm.rank() == 0) {
        const bool more = is_finite(fv) ? opt->advance(fv, g) : false;       // non-PD matrix: give up on this start
        if (!is_finite(fv)) opt->f = pinf();
        ctl[0] = opt->x[0]; ctl[1] = opt->x[1]; ctl[2] = opt->x[2]; ctl[3] = more ? 1.0 : 0.0;
        ctl[4] = opt->f; ctl[6] = (double)opt->nit;
      }
      tm.sync();
      const bool more = ctl[3] != 0.0;
      if (!more) break;
      tm.sync();
    }
    ctl[5] = (double)evals;     // same value from every member
    const double ff = ctl[4];
    out.nfev += (int)ctl[5]; out.nit += (int)ctl[6];
    if (ff < best_f) { best_f = ff; best_x[0] = ctl[0]; best_x[1] = ctl[1]; best_x[2] = ctl[2]; out.winner = st; }
    tm.sync();
  }
  if (out.winner < 0) return;
  out.length_scale = exp(best_x[1]) * t_range;
  out.amplitude = sqrt(exp(best_x[0])) * f_std;
  out.noise = sqrt(exp(best_x[2])) * f_std;
  out.loglik = -best_f;
}

MLRN_FN void gp1d_clear(Gp1dBandResult& out) {
  out.length_scale = out.amplitude = out.noise = out.loglik = qnan();
  out.nfev = 0; out.nit = 0; out.winner = -1;
}

// One band with the packed-triangle sweep in team scratch `s` (gaussian_process.py:29-146).
template <class Team>
MLRN_FN void gp1d_band(const Team& tm, const double* gt, const double* gf, const double* ge, int n_raw,
                       const Gp1dScratch& s, LbfgsB<3>* opt, double* ctl, Gp1dBandResult& out) {
  gp1d_clear(out);
  double t_range = 0.0, f_std = 1.0;
  const int n = gp1d_prepare(tm, gt, gf, ge, n_raw, s.x, s.y, s.al, t_range, f_std);
  if (n == 0) return;
  gp1d_optimise(tm, [&](const double* theta, double& fv, double* g) { gp1d_objective(tm, n, s, theta, fv, g); },
                t_range, f_std, opt, ctl, out);
}

// Cross-band block of extract_gp_features_single (gaussian_process.py:214-246); scalar, every member computes it.
MLRN_FN void gp1d_tail(const Gp1dBandResult* b /* g r i z */, double* row) {
  for (int k = 0; k < 4; ++k) {
    row[4 * k + 0] = b[k].length_scale; row[4 * k + 1] = b[k].amplitude;
    row[4 * k + 2] = b[k].noise; row[4 * k + 3] = b[k].loglik;
  }
  for (int k = 0; k < 2; ++k) {
    const double l1 = b[k].length_scale, l2 = b[k + 1].length_scale;
    row[16 + k] = (!is_nan(l1) && !is_nan(l2) && l2 > 0.0) ? l1 / l2 : qnan();
  }
  double ls[3] = {b[0].length_scale, b[1].length_scale, b[2].length_scale};
  double am[3] = {b[0].amplitude, b[1].amplitude, b[2].amplitude};
  int cnt; double mean, sd;
  small_stats(ls, 3, cnt, mean, sd);
  row[18] = cnt ? mean : qnan();
  row[19] = cnt ? (cnt > 1 ? sd : 0.0) : qnan();
  small_stats(am, 3, cnt, mean, sd);
  row[20] = cnt ? mean : qnan();
}

}  // namespace mlrn
