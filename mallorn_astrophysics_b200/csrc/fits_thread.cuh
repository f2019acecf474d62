// fits_thread.cuh -- K4 with one THREAD per fit (trf_thread.cuh): the Bazin fit of one (object, band) and
// the decline-model fit of one (object, band, model) as resumable jobs.  Same reference behaviour as
// fits.cuh (bazin_fitting.py:63-179, scripts/train_v55_powerlaw.py:147-194); the band is read in place
// from the packed store (time-sorted), nothing is staged.
#pragma once
#include "fits.cuh"
#include "trf_thread.cuh"

namespace mlrn {

// numpy.median of a[0..n) without scratch: one rank-counting sweep (ties broken by index) that picks up both middle
// order statistics at once; O(n^2) compares, n is a band length
MLRN_FN double median_unsorted(const double* a, int n) {
  const int k_hi = n >> 1, k_lo = (n & 1) ? k_hi : k_hi - 1;
  double lo = qnan(), hi = qnan();
  int found = 0;
  for (int i = 0; i < n && found < 2; ++i) {
    const double v = a[i];
    int rk = 0;
    for (int j = 0; j < n; ++j) { const double u = a[j]; rk += (u < v) || (u == v && j < i); }
    if (rk == k_lo) { lo = v; ++found; if (k_hi == k_lo) { hi = v; ++found; } }
    else if (rk == k_hi) { hi = v; ++found; }
  }
  return (n & 1) ? hi : 0.5 * (lo + hi);
}

template <bool COOP>
struct BazinJobT {
  TrfThread<5, BazinModel, COOP> trf;
  const double *t, *f, *e; int n;

  // Returns true if the fit is running; otherwise `code` is final (skipped / bad input / infeasible start).
  // pre_in / pre_out: {median of the fluxes, cost at the start}, the two per-fit quantities that cost a pass (O(n^2) and
  // n model evaluations) -- computed by the prologue pass of fits.cu with every lane busy, read here (PRE doubles per fit)
  static constexpr int PRE = 2;
  MLRN_FN bool init(const double* t_, const double* f_, const double* e_, int n_, int& code,
                    const double* pre_in = nullptr, double* pre_out = nullptr) {
    t = t_; f = f_; e = e_; n = n_;
    if (n < 5) { code = FIT_SKIPPED; return false; }                          // :76
    bool bad = false; int pk = 0;
    for (int i = 0; i < n; ++i) { bad = bad || !is_finite(t[i]) || !is_finite(f[i]); if (f[i] > f[pk]) pk = i; }
    if (bad) { code = FIT_BAD_INPUT; return false; }
    const double med = pre_in ? pre_in[0] : median_unsorted(f, n), mx = f[pk], dur = t[n - 1] - t[0];
    if (pre_out) pre_out[0] = med;
    const double x0[5] = {f[pk] - med, t[pk], dur * 0.2, dur * 0.3, med};     // :97-105
    const double lo[5] = {0.0, t[0], 0.1, 0.1, -mx};                          // :114-118
    const double hi[5] = {3.0 * mx, t[n - 1], dur, dur, 2.0 * mx};
    code = trf.begin(BazinModel{t, f, e}, n, x0, lo, hi, 2000, pre_in ? pre_in + 1 : nullptr, pre_out ? pre_out + 1 : nullptr);
    return code == TRF_RUNNING;
  }
  MLRN_FN int step() { return trf.step(); }
  // out: 8 features (NaN unless status > 0)
  MLRN_FN void finish(int status, double* out) const {
    for (int k = 0; k < 8; ++k) out[k] = qnan();
    if (status <= 0) return;
    const double* x = trf.x;
    const double A = clip(x[0], -1e6, 1e6), tr = clip(x[2], 0.1, 1e4), tf = clip(x[3], 0.1, 1e4), B = clip(x[4], -1e6, 1e6);
    const double xc[5] = {A, x[1], tr, tf, B};
    // chi^2 of the clipped curve (:160-166).  The clips are inactive for any fit inside its bounds with |A|, |B| <= 1e6,
    // and then the sum is the solver's own final cost (0.5 sum r^2 over the same rows, r = (model - f) / sigma): no second
    // pass over the band by a single lane of the warp
    double chi = 2.0 * trf.cost;
    if (!(A == x[0] && tr == x[2] && tf == x[3] && B == x[4])) {
      chi = 0.0;
      for (int i = 0; i < n; ++i) { const double sg = (e[i] > 0.0) ? e[i] : 1.0; const double r = (f[i] - bazin_value(t[i], xc)) / sg; chi += r * r; }
    }
    out[0] = A; out[1] = x[1]; out[2] = tr; out[3] = tf; out[4] = B;
    out[5] = clip(chi / (double)(n - 5), 0.0, 1e6);
    out[6] = clip(tr / (tf + 1e-6), 0.0, 100.0);
    out[7] = clip(A + B, -1e6, 1e6);
  }
};

using BazinJob = BazinJobT<false>;

// decline models on the post-peak suffix of a time-sorted band: x_i = t_i - t_peak, y_i = f_i
struct NoPrep {};
struct PowerModelT {
  const double* t; const double* y; double pt, k;
  using Prep1 = NoPrep; using Prep = GenericPrep<2, PowerModelT>;
  MLRN_FN_FORCE Prep1 prepare1(const double*) const { return Prep1{}; }
  MLRN_FN_FORCE Prep prepare(const double* x, const double* xpd) const { return generic_prepare<2>(*this, x, xpd); }
  MLRN_FN double fd_row(int i, const double* x, const double* xpd, const double* dx, const double* rdx, const Prep& p, double* row) const { return generic_fd_row<2>(*this, i, x, xpd, dx, rdx, p, row); }
  // x^k as exp(k ln x): the base is clipped to >= 0.1, so none of pow()'s special cases can occur and the general routine
  // (about twice the instructions) is not needed; |k ln x| <= 3 * 12, i.e. the result is good to ~2e-15 relative, far inside
  // what a two-parameter R^2 can see (asserted at 1e-3)
  MLRN_FN_FORCE double value(int i, const double* p) const { return p[0] * exp(k * log(fmax((t[i] - pt) - p[1], 0.1))); }
  MLRN_FN_FORCE double resid(int i, const double* p, const Prep1&) const { return value(i, p) - y[i]; }
};
struct ExpModelT {
  const double* t; const double* y; double pt, k;
  struct Prep1 { double r1; };                            // 1 / tau
  using Prep = GenericPrep<3, ExpModelT>;
  MLRN_FN_FORCE Prep1 prepare1(const double* p) const { Prep1 q; q.r1 = 1.0 / p[1]; return q; }
  MLRN_FN_FORCE Prep prepare(const double* x, const double* xpd) const { return generic_prepare<3>(*this, x, xpd); }
  MLRN_FN double fd_row(int i, const double* x, const double* xpd, const double* dx, const double* rdx, const Prep& p, double* row) const { return generic_fd_row<3>(*this, i, x, xpd, dx, rdx, p, row); }
  MLRN_FN_FORCE double value(int i, const double* p) const { return p[0] * exp(-fmax((t[i] - pt) - p[2], 0.0) / p[1]); }
  MLRN_FN_FORCE double resid(int i, const double* p, const Prep1& q) const { return p[0] * exp(div_r(-fmax((t[i] - pt) - p[2], 0.0), p[1], q.r1)) - y[i]; }
};
struct LinModelT {
  const double* t; const double* y; double pt, k;
  using Prep1 = NoPrep; using Prep = GenericPrep<3, LinModelT>;
  MLRN_FN_FORCE Prep1 prepare1(const double*) const { return Prep1{}; }
  MLRN_FN_FORCE Prep prepare(const double* x, const double* xpd) const { return generic_prepare<3>(*this, x, xpd); }
  MLRN_FN double fd_row(int i, const double* x, const double* xpd, const double* dx, const double* rdx, const Prep& p, double* row) const { return generic_fd_row<3>(*this, i, x, xpd, dx, rdx, p, row); }
  MLRN_FN_FORCE double value(int i, const double* p) const { return p[0] - p[1] * fmax((t[i] - pt) - p[2], 0.0); }
  MLRN_FN_FORCE double resid(int i, const double* p, const Prep1&) const { return value(i, p) - y[i]; }
};

// KIND 0: power law with fixed index k (NP = 2); 1: exponential (NP = 3); 2: linear (NP = 3)
template <int KIND> struct DeclineTraits;
template <> struct DeclineTraits<0> { static constexpr int NP = 2; using Model = PowerModelT; };
template <> struct DeclineTraits<1> { static constexpr int NP = 3; using Model = ExpModelT; };
template <> struct DeclineTraits<2> { static constexpr int NP = 3; using Model = LinModelT; };

template <int KIND, bool COOP>
struct DeclineJobT {
  static constexpr int NP = DeclineTraits<KIND>::NP;
  using Model = typename DeclineTraits<KIND>::Model;
  TrfThread<NP, Model, COOP> trf;
  Model mdl; int m; double ss_tot;

  static constexpr int PRE = 0;         // no prologue pass: measured, one cost per fit up front does not pay for 27 fits per object (709 -> 719 ms)
  MLRN_FN bool init(const double* t, const double* f, int n, double kpow, int& code,
                    const double* pre_in = nullptr, double* pre_out = nullptr) {
    if (n < 5) { code = FIT_SKIPPED; return false; }                          // :150
    int pk = 0;
    for (int i = 1; i < n; ++i) if (f[i] > f[pk]) pk = i;
    const double pt = t[pk], pf = f[pk];
    int start = n;
    for (int i = n - 1; i >= 0 && t[i] > pt; --i) start = i;                  // time-sorted: t > t_peak is a suffix
    m = n - start;
    if (m < 3) { code = FIT_SKIPPED; return false; }                          // :162
    const double* ts = t + start; const double* ys = f + start;
    bool bad = false; double sum = 0.0;
    for (int i = 0; i < m; ++i) { bad = bad || !is_finite(ts[i] - pt) || !is_finite(ys[i]); sum += ys[i]; }
    if (bad) { code = FIT_BAD_INPUT; return false; }
    const double ymean = sum / (double)m;
    ss_tot = 0.0;
    for (int i = 0; i < m; ++i) { const double d = ys[i] - ymean; ss_tot += d * d; }
    mdl = Model{ts, ys, pt, kpow};
    double x0[3], lo[3], hi[3];
    if (KIND == 0) { x0[0] = pf; x0[1] = 0.0; lo[0] = 0.0; lo[1] = -10.0; hi[0] = 1e6; hi[1] = 10.0; }
    else if (KIND == 1) { x0[0] = pf; x0[1] = 30.0; x0[2] = 0.0; lo[0] = 0.0; lo[1] = 1.0; lo[2] = -10.0; hi[0] = 1e6; hi[1] = 500.0; hi[2] = 10.0; }
    else { x0[0] = pf; x0[1] = 1.0; x0[2] = 0.0; lo[0] = 0.0; lo[1] = 0.0; lo[2] = -10.0; hi[0] = 1e6; hi[1] = 100.0; hi[2] = 10.0; }
    code = trf.begin(mdl, m, x0, lo, hi, 1000, pre_in, pre_out);
    return code == TRF_RUNNING;
  }
  MLRN_FN int step() { return trf.step(); }
  MLRN_FN double finish(int status) const {
    if (status <= 0) return qnan();
    const double ss = 2.0 * trf.cost;          // sum (y - model)^2 at the final point: the solver's cost is half of exactly that sum
    return (ss_tot > 0.0) ? 1.0 - ss / ss_tot : 0.0;                           // :186-189
  }
};

template <int KIND> using DeclineJob = DeclineJobT<KIND, false>;

MLRN_FN double decline_index(int q) {                                          // MODELS order (:135-145)
  const double kk[7] = {-5.0 / 3.0, -1.0, -1.5, -2.0, -2.5, -3.0, -0.5};
  return kk[q];
}

}  // namespace mlrn
