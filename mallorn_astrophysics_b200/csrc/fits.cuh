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

// a / den for a >= 0 and den in [1, +inf], the two shapes the Bazin curve produces.  Far from the peak the numerator
// underflows to 0 or the denominator overflows to +inf, and for such operands the compiler's division leaves its fast path
// for a ~50-instruction subroutine that one or two lanes of the warp execute alone (ncu: 7.6 % of all warp instructions of
// the fit kernel at 2.2 threads each).  Both cases have the exact IEEE quotient 0 and are answered directly.
MLRN_FN_FORCE double div_curve(double a, double den) {
  if (a == 0.0) return a;
  if (den == pinf() && a < pinf()) return 0.0;
  return a / den;
}

struct BazinModel {
  const double* t; const double* f; const double* e;
  struct Prep1 { double r2, r3; };                        // 1 / tau_rise, 1 / tau_fall of one parameter vector
  struct Prep { double r2, r3, rp2, rp3; };               // ... and of the two perturbed time scales
  MLRN_FN_FORCE Prep1 prepare1(const double* x) const { Prep1 p; p.r2 = 1.0 / x[2]; p.r3 = 1.0 / x[3]; return p; }
  MLRN_FN_FORCE Prep prepare(const double* x, const double* xpd) const {
    Prep p; p.r2 = 1.0 / x[2]; p.r3 = 1.0 / x[3]; p.rp2 = 1.0 / xpd[2]; p.rp3 = 1.0 / xpd[3]; return p;
  }
  // curve_fit's wrapped residual: transform * (func(xdata, *params) - ydata), transform = 1 / sigma
  // (bazin_fitting.py:54-60, same association: A * num / den + B)
  MLRN_FN_FORCE double resid(int i, const double* x, const Prep1& p) const {
    double sg = (e[i] > 0.0) ? e[i] : 1.0;                    // bazin_fitting.py:126
    const double u = t[i] - x[1];
    const double num = exp(div_r(-u, x[3], p.r3)), den = 1.0 + exp(div_r(-u, x[2], p.r2));
    return (1.0 / sg) * (div_curve(x[0] * num, den) + x[4] - f[i]);
  }
  // Forward-difference Jacobian row with the exponentials shared between the six evaluations: perturbing A
  // or B needs no new exp, tau_rise / tau_fall one each, t0 two -- 6 instead of 12.  Every value is computed by
  // the same operations on the same operands as resid() would, so the row is bit-identical to the generic one.
  MLRN_FN double fd_row(int i, const double* x, const double* xpd, const double* dx, const double* rdx, const Prep& p, double* row) const {
    const double sg = (e[i] > 0.0) ? e[i] : 1.0, w = 1.0 / sg, ti = t[i], fi = f[i];
    const double u = ti - x[1];
    const double en = exp(div_r(-u, x[3], p.r3)), den = 1.0 + exp(div_r(-u, x[2], p.r2));
    const double core = div_curve(x[0] * en, den);
    const double r0 = w * (core + x[4] - fi);
    const double u1 = ti - xpd[1];
    const double v0 = w * (div_curve(xpd[0] * en, den) + x[4] - fi);
    const double v1 = w * (div_curve(x[0] * exp(div_r(-u1, x[3], p.r3)), 1.0 + exp(div_r(-u1, x[2], p.r2))) + x[4] - fi);
    const double v2 = w * (div_curve(x[0] * en, 1.0 + exp(div_r(-u, xpd[2], p.rp2))) + x[4] - fi);
    const double v3 = w * (div_curve(x[0] * exp(div_r(-u, xpd[3], p.rp3)), den) + x[4] - fi);
    const double v4 = w * (core + xpd[4] - fi);
    row[0] = div_r(v0 - r0, dx[0], rdx[0]); row[1] = div_r(v1 - r0, dx[1], rdx[1]); row[2] = div_r(v2 - r0, dx[2], rdx[2]);
    row[3] = div_r(v3 - r0, dx[3], rdx[3]); row[4] = div_r(v4 - r0, dx[4], rdx[4]);
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
