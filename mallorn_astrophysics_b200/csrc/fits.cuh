// fits.cuh -- K4: the bounded nonlinear fits of the feature table, on top of trf.cuh.
//
//   bazin_row      52 columns (schema.BAZIN_COLUMNS), reference src/features/bazin_fitting.py:63-251:
//                  5-parameter Bazin curve per band, weighted by Flux_err, raw MJD time axis.
//   powerlaw_row   27 columns (schema.POWERLAW_COLUMNS), reference scripts/train_v55_powerlaw.py:108-202:
//                  nine post-peak decline models per band g, r, i, unweighted; the feature is R^2.
// One team (warp) owns one object and walks its bands; every fit is one trf_bounded() call.
// Status convention (optional int32 output): {TRF status or TrfFail code, nfev} per fit;
// FIT_SKIPPED where the reference never calls curve_fit (too few points).
#pragma once
#include "objview.cuh"
#include "trf.cuh"

namespace mlrn {

constexpr int BAZIN_COLS = 52;
constexpr int POWERLAW_COLS = 27;
constexpr int FIT_SKIPPED = -200;       // fewer points than the reference requires: no fit attempted
constexpr int FIT_BAD_INPUT = -106;     // NaN/inf in the data: curve_fit(check_finite=True) raises ValueError

template <int NP>
MLRN_HD size_t trf_work_doubles(int cap) { return (size_t)cap * 2 + (size_t)NP * cap + (size_t)NP * (cap + NP); }

template <int NP>
MLRN_FN TrfWork<NP> carve_work(double* base, int cap) {
  TrfWork<NP> w;
  w.f = base; w.fn = base + cap; w.J = base + 2 * cap; w.Q = w.J + (size_t)NP * cap; w.ld = cap;
  return w;
}

MLRN_FN_FORCE double clip(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }   // NaN propagates like numpy.clip

// ------------------------------------------------------------------------------------------
// Bazin
// ------------------------------------------------------------------------------------------
MLRN_FN_FORCE double bazin_value(double t, const double* x) {
  // bazin_fitting.py:54-60, same association: A * num / den + B
  double num = exp(-(t - x[1]) / x[3]);
  double den = 1.0 + exp(-(t - x[1]) / x[2]);
  return x[0] * num / den + x[4];
}

struct BazinModel {
  const double* t; const double* f; const double* e;
  // curve_fit's wrapped residual: transform * (func(xdata, *params) - ydata), transform = 1 / sigma
  MLRN_FN_FORCE double resid(int i, const double* x) const {
    double sg = (e[i] > 0.0) ? e[i] : 1.0;                    // bazin_fitting.py:126
    return (1.0 / sg) * (bazin_value(t[i], x) - f[i]);
  }
};

// One band, time-sorted (t, f, e)[0..n).  out: 8 features.  srt: scratch >= next_pow2(n) for the median.
template <class Team>
MLRN_FN void bazin_band(const Team& tm, const double* t, const double* f, const double* e, int n, double* srt,
                        double* work, int cap, double* out, int* status, int* nfev_out) {
  for (int k = tm.rank(); k < 8; k += Team::size) out[k] = qnan();
  *status = FIT_SKIPPED; *nfev_out = 0;
  tm.sync();
  if (n < 5) return;                                          // :76
  bool bad = false;
  for (int i = tm.rank(); i < n; i += Team::size) bad = bad || !is_finite(t[i]) || !is_finite(f[i]);
  if (tm.any(bad)) { *status = FIT_BAD_INPUT; return; }
  // p0 and box (:97-118)
  const int pk = team_argmax_first(tm, f, n);
  for (int i = tm.rank(); i < n; i += Team::size) srt[i] = f[i];
  tm.sync();
  tm.sort(srt, n);
  const double med = sorted_median(srt, n);
  tm.sync();
  const double mx = f[pk], dur = t[n - 1] - t[0];
  double x[5] = {f[pk] - med, t[pk], dur * 0.2, dur * 0.3, med};
  const double lb[5] = {0.0, t[0], 0.1, 0.1, -mx};
  const double ub[5] = {3.0 * mx, t[n - 1], dur, dur, 2.0 * mx};
  BazinModel mdl{t, f, e};
  TrfWork<5> w = carve_work<5>(work, cap);
  int nfev; double cost;
  int st = trf_bounded<5>(tm, mdl, n, x, lb, ub, 2000, w, nfev, cost);
  *status = st; *nfev_out = nfev;
  if (st <= 0) return;                                        // curve_fit raises unless res.success (status > 0)
  // clips and derived features (:141-160)
  const double A = clip(x[0], -1e6, 1e6), tr = clip(x[2], 0.1, 1e4), tf = clip(x[3], 0.1, 1e4), B = clip(x[4], -1e6, 1e6);
  const double xc[5] = {A, x[1], tr, tf, B};
  double chi = 0.0;
  for (int i = tm.rank(); i < n; i += Team::size) {
    double sg = (e[i] > 0.0) ? e[i] : 1.0;
    double r = (f[i] - bazin_value(t[i], xc)) / sg;
    chi += r * r;
  }
  chi = tm.sum(chi);
  if (tm.rank() == 0) {
    out[0] = A; out[1] = x[1]; out[2] = tr; out[3] = tf; out[4] = B;
    out[5] = clip(chi / (double)(n - 5), 0.0, 1e6);            // n == 5: x/0 -> inf -> 1e6 (0/0 -> NaN)
    out[6] = clip(tr / (tf + 1e-6), 0.0, 100.0);
    out[7] = clip(A + B, -1e6, 1e6);
  }
  tm.sync();
}

// cross-band block (bazin_fitting.py:217-249) on a row whose 6 x 8 band features are filled
template <class Team>
MLRN_FN void bazin_tail(const Team& tm, double* row) {
  tm.sync();
  if (tm.rank() == 0) {
    double rise[3] = {row[8 * BG + 2], row[8 * BR + 2], row[8 * BI + 2]};
    double fall[3] = {row[8 * BG + 3], row[8 * BR + 3], row[8 * BI + 3]};
    int c; double m, s;
    small_stats(rise, 3, c, m, s); row[48] = (c >= 2) ? s / m : qnan();
    small_stats(fall, 3, c, m, s); row[49] = (c >= 2) ? s / m : qnan();
    double chi[6];
    for (int b = 0; b < 6; ++b) chi[b] = row[8 * b + 5];
    small_stats(chi, 6, c, m, s); row[50] = c ? m : qnan(); row[51] = c ? s : qnan();
  }
  tm.sync();
}

// srt: >= next_pow2(longest band); work: trf_work_doubles<5>(cap); status: [6][2] or nullptr
template <class Team>
MLRN_FN void bazin_row(const Team& tm, const ObjView& v, double* srt, double* work, int cap, double* row, int* status) {
  for (int b = 0; b < N_BAND; ++b) {
    const int lo = v.off[b];
    int st, nf;
    bazin_band(tm, v.t + lo, v.f + lo, v.e + lo, v.cnt(b), srt, work, cap, row + 8 * b, &st, &nf);
    if (status && tm.rank() == 0) { status[2 * b] = st; status[2 * b + 1] = nf; }
  }
  bazin_tail(tm, row);
}

// ------------------------------------------------------------------------------------------
// Decline models (train_v55_powerlaw.py:108-133)
// ------------------------------------------------------------------------------------------
struct PowerModel {       // A * max(x - t0, 0.1)^k
  const double* x; const double* y; double k;
  MLRN_FN_FORCE double value(int i, const double* p) const { return p[0] * pow(fmax(x[i] - p[1], 0.1), k); }
  MLRN_FN_FORCE double resid(int i, const double* p) const { return value(i, p) - y[i]; }
};
struct ExpModel {         // A * exp(-max(x - t0, 0) / tau), p = (A, tau, t0)
  const double* x; const double* y;
  MLRN_FN_FORCE double value(int i, const double* p) const { return p[0] * exp(-fmax(x[i] - p[2], 0.0) / p[1]); }
  MLRN_FN_FORCE double resid(int i, const double* p) const { return value(i, p) - y[i]; }
};
struct LinModel {         // A - b * max(x - t0, 0), p = (A, b, t0)
  const double* x; const double* y;
  MLRN_FN_FORCE double value(int i, const double* p) const { return p[0] - p[1] * fmax(x[i] - p[2], 0.0); }
  MLRN_FN_FORCE double resid(int i, const double* p) const { return value(i, p) - y[i]; }
};

template <int NP, class Team, class Model>
MLRN_FN double decline_r2(const Team& tm, const Model& mdl, int m, double* p, const double* lb, const double* ub,
                          double ss_tot, double* work, int cap, int* status, int* nfev_out) {
  TrfWork<NP> w = carve_work<NP>(work, cap);
  int nfev; double cost;
  int st = trf_bounded<NP>(tm, mdl, m, p, lb, ub, 1000, w, nfev, cost);
  *status = st; *nfev_out = nfev;
  if (st <= 0) return qnan();
  double ss = 0.0;
  for (int i = tm.rank(); i < m; i += Team::size) { double r = mdl.y[i] - mdl.value(i, p); ss += r * r; }
  ss = tm.sum(ss);
  return (ss_tot > 0.0) ? 1.0 - ss / ss_tot : 0.0;            // :186-189
}

// One band (time-sorted).  xs, ys: scratch >= n.  out: 9 R^2 values.  status: [9][2] or nullptr.
template <class Team>
MLRN_FN void decline_band(const Team& tm, const double* t, const double* f, int n, double* xs, double* ys,
                          double* work, int cap, double* out, int* status) {
  for (int k = tm.rank(); k < 9; k += Team::size) out[k] = qnan();
  if (status) for (int k = tm.rank(); k < 9; k += Team::size) { status[2 * k] = FIT_SKIPPED; status[2 * k + 1] = 0; }
  tm.sync();
  if (n < 5) return;                                          // :150
  const int pk = team_argmax_first(tm, f, n);
  const double pt = t[pk], pf = f[pk];
  int m = 0;
  for (int base = 0; base < n; base += Team::size) {          // post-peak points, t > t_peak (:158-160)
    int i = base + tm.rank();
    bool keep = (i < n) && (t[i] > pt);
    int slot = tm.compact_slot(keep, m);
    if (keep) { xs[slot] = t[i] - pt; ys[slot] = f[i]; }
  }
  tm.sync();
  if (m < 3) return;                                          // :162
  bool bad = false;
  for (int i = tm.rank(); i < m; i += Team::size) bad = bad || !is_finite(xs[i]) || !is_finite(ys[i]);
  if (tm.any(bad)) {
    if (status) for (int k = tm.rank(); k < 9; k += Team::size) status[2 * k] = FIT_BAD_INPUT;
    return;
  }
  const double ymean = team_mean(tm, ys, m);
  double ss_tot = 0.0;
  for (int i = tm.rank(); i < m; i += Team::size) { double d = ys[i] - ymean; ss_tot += d * d; }
  ss_tot = tm.sum(ss_tot);
  const double kk[7] = {-5.0 / 3.0, -1.0, -1.5, -2.0, -2.5, -3.0, -0.5};     // MODELS order (:135-145)
  int st, nf;
  for (int q = 0; q < 7; ++q) {
    double p[2] = {pf, 0.0};
    const double lb[2] = {0.0, -10.0}, ub[2] = {1e6, 10.0};
    PowerModel mdl{xs, ys, kk[q]};
    double r2 = decline_r2<2>(tm, mdl, m, p, lb, ub, ss_tot, work, cap, &st, &nf);
    if (tm.rank() == 0) { out[q] = r2; if (status) { status[2 * q] = st; status[2 * q + 1] = nf; } }
  }
  {
    double p[3] = {pf, 30.0, 0.0};
    const double lb[3] = {0.0, 1.0, -10.0}, ub[3] = {1e6, 500.0, 10.0};
    ExpModel mdl{xs, ys};
    double r2 = decline_r2<3>(tm, mdl, m, p, lb, ub, ss_tot, work, cap, &st, &nf);
    if (tm.rank() == 0) { out[7] = r2; if (status) { status[14] = st; status[15] = nf; } }
  }
  {
    double p[3] = {pf, 1.0, 0.0};
    const double lb[3] = {0.0, 0.0, -10.0}, ub[3] = {1e6, 100.0, 10.0};
    LinModel mdl{xs, ys};
    double r2 = decline_r2<3>(tm, mdl, m, p, lb, ub, ss_tot, work, cap, &st, &nf);
    if (tm.rank() == 0) { out[8] = r2; if (status) { status[16] = st; status[17] = nf; } }
  }
  tm.sync();
}

// work: trf_work_doubles<3>(cap); status: [27][2] or nullptr
template <class Team>
MLRN_FN void powerlaw_row(const Team& tm, const ObjView& v, double* xs, double* ys, double* work, int cap,
                          double* row, int* status) {
  const int bands[3] = {BG, BR, BI};
  for (int j = 0; j < 3; ++j) {
    const int lo = v.off[bands[j]];
    decline_band(tm, v.t + lo, v.f + lo, v.cnt(bands[j]), xs, ys, work, cap, row + 9 * j, status ? status + 18 * j : nullptr);
  }
  tm.sync();
}

}  // namespace mlrn
