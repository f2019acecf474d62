// fits.cuh -- K4: shared pieces of the bounded nonlinear fits (the jobs themselves are in fits_thread.cuh).
//
//   Bazin           52 columns (schema.BAZIN_COLUMNS), reference src/features/bazin_fitting.py:63-251:
//                   5-parameter Bazin curve per band, weighted by Flux_err, raw MJD time axis.
//   decline models  27 columns (schema.POWERLAW_COLUMNS), reference scripts/train_v55_powerlaw.py:108-202:
//                   nine post-peak decline models per band g, r, i, unweighted; the feature is R^2.
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
  // Forward-difference Jacobian row with the exponentials shared between the six evaluations: perturbing A
  // or B needs no new exp, tau_rise / tau_fall one each, t0 two -- 6 instead of 12.  Every value is computed by
  // the same operations on the same operands as resid() would, so the row is bit-identical to the generic one.
  MLRN_FN double fd_row(int i, const double* x, const double* xpd, const double* dx, double* row) const {
    const double sg = (e[i] > 0.0) ? e[i] : 1.0, w = 1.0 / sg, ti = t[i], fi = f[i];
    const double u = ti - x[1];
    const double en = exp(-u / x[3]), den = 1.0 + exp(-u / x[2]);
    const double core = x[0] * en / den;
    const double r0 = w * (core + x[4] - fi);
    const double u1 = ti - xpd[1];
    const double v0 = w * (xpd[0] * en / den + x[4] - fi);
    const double v1 = w * (x[0] * exp(-u1 / x[3]) / (1.0 + exp(-u1 / x[2])) + x[4] - fi);
    const double v2 = w * (x[0] * en / (1.0 + exp(-u / xpd[2])) + x[4] - fi);
    const double v3 = w * (x[0] * exp(-u / xpd[3]) / den + x[4] - fi);
    const double v4 = w * (core + xpd[4] - fi);
    row[0] = (v0 - r0) / dx[0]; row[1] = (v1 - r0) / dx[1]; row[2] = (v2 - r0) / dx[2];
    row[3] = (v3 - r0) / dx[3]; row[4] = (v4 - r0) / dx[4];
    return r0;
  }
};

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

}  // namespace mlrn
