// shape.cuh -- light-curve shape features (65 columns, schema.SHAPE_COLUMNS).
// Behaviour follows reference src/features/lightcurve_shape.py:177-332.  Per-band quantities use
// the time-sorted band (>= 3 points); the all-band block uses every row in FILE order, which
// matters for the "first pre-peak observation above threshold" of compute_rise_time (:34-65).
#pragma once
#include "objview.cuh"
#include "stat.cuh"

namespace mlrn {

constexpr int SHAPE_COLS = 65;

// Log-log decay fit shared by the band and all-band blocks (lightcurve_shape.py:107-144):
// points with t > t_peak + 5 and f > 0, need >= 5; x = log10(max(dt, 1)), y = log10(max(f, 1e-10)).
template <class Team>
MLRN_TEAM_FN void loglog_decay(const Team& tm, const double* t, const double* f, int n, double pt, double pf,
                          double* sx, double* sy, double& alpha, double& resid) {
  alpha = qnan(); resid = qnan();
  if (is_nan(pt) || is_nan(pf)) return;
  int cnt = 0;
  for (int base = 0; base < n; base += Team::size) {
    int i = base + tm.rank();
    bool keep = (i < n) && (t[i] > pt + 5.0) && (f[i] > 0.0);
    int slot = tm.compact_slot(keep, cnt);
    if (keep) { sx[slot] = log10(fmax(t[i] - pt, 1.0)); sy[slot] = log10(fmax(f[i], 1e-10)); }
  }
  tm.sync();
  if (cnt >= 5) {
    Line l = team_polyfit1(tm, sx, sy, cnt);
    double ss = 0.0;
    for (int i = tm.rank(); i < cnt; i += Team::size) { double r = sy[i] - (l.slope * sx[i] + l.icpt); ss += r * r; }
    ss = tm.sum(ss);
    alpha = l.slope; resid = sqrt(ss / (double)cnt);
  }
  tm.sync();
}

// sa, sb: scratch with capacity >= next_pow2(v.n)
template <class Team>
MLRN_FN void shape_row(const Team& tm, const ObjView& v, double* sa, double* sb, double* row) {
  for (int k = tm.rank(); k < SHAPE_COLS; k += Team::size) row[k] = qnan();
  tm.sync();
  double peak_t[N_BAND];
  for (int b = 0; b < N_BAND; ++b) {
    peak_t[b] = qnan();
    const int n = v.cnt(b);
    if (n < 3) continue;                                           // :193
    const double* t = v.t + v.off[b];
    const double* f = v.f + v.off[b];
    const int k = team_argmax_first(tm, f, n);                     // find_peak_info, time order
    const double pt = t[k], pf = f[k];
    peak_t[b] = pt;
    // rise: first pre-peak point above 0.1 * peak, else the first pre-peak point (:47-65)
    // fade: first post-peak point below frac * peak, else the last post-peak point (:81-104)
    // duration: span of points above frac * max (:147-161)
    const double thr = 0.1 * pf, t50 = 0.5 * pf, t25 = 0.25 * pf;
    int first_pre = 0x7fffffff, first_above = 0x7fffffff, fade50 = 0x7fffffff, fade25 = 0x7fffffff, last_post = -1;
    double d50lo = pinf(), d50hi = -pinf(), d25lo = pinf(), d25hi = -pinf();
    for (int i = tm.rank(); i < n; i += Team::size) {
      double ti = t[i], fi = f[i];
      if (ti < pt) { if (i < first_pre) first_pre = i; if (fi > thr && i < first_above) first_above = i; }
      if (ti > pt) {
        if (i > last_post) last_post = i;
        if (fi < t50 && i < fade50) fade50 = i;
        if (fi < t25 && i < fade25) fade25 = i;
      }
      if (fi > t50) { d50lo = fmin(d50lo, ti); d50hi = fmax(d50hi, ti); }
      if (fi > t25) { d25lo = fmin(d25lo, ti); d25hi = fmax(d25hi, ti); }
    }
    first_pre = tm.imin(first_pre); first_above = tm.imin(first_above);
    fade50 = tm.imin(fade50); fade25 = tm.imin(fade25); last_post = tm.imax(last_post);
    d50lo = tm.min(d50lo); d50hi = tm.max(d50hi); d25lo = tm.min(d25lo); d25hi = tm.max(d25hi);
    double rise = qnan(), f50 = qnan(), f25 = qnan();
    if (!is_nan(pt) && !is_nan(pf)) {
      if (first_pre != 0x7fffffff) rise = pt - t[first_above != 0x7fffffff ? first_above : first_pre];
      if (last_post >= 0) {
        f50 = t[fade50 != 0x7fffffff ? fade50 : last_post] - pt;
        f25 = t[fade25 != 0x7fffffff ? fade25 : last_post] - pt;
      }
    }
    double alpha, resid;
    loglog_decay(tm, t, f, n, pt, pf, sa, sb, alpha, resid);
    if (tm.rank() == 0) {
      double* o = row + 8 * b;
      o[0] = rise; o[1] = f50; o[2] = f25;
      o[3] = (!is_nan(rise) && !is_nan(f50) && f50 > 0.0) ? rise / f50 : qnan();     // :235-238
      o[4] = (d50hi >= d50lo) ? d50hi - d50lo : 0.0;
      o[5] = (d25hi >= d25lo) ? d25hi - d25lo : 0.0;
      o[6] = alpha; o[7] = resid;
    }
    tm.sync();
  }
  // ---- all-band block on every row in file order (:287-319)
  double a_rise = qnan(), a_fade = qnan(), a_alpha = qnan(), a_resid = qnan(), conc = qnan();
  double pct[4] = {qnan(), qnan(), qnan(), qnan()};
  const int n = v.n;
  if (n >= 5) {
    const int k = argmax_file_order(tm, v, 0, n);
    const double pt = v.t[k], pf = v.f[k];
    const double thr = 0.1 * pf, t50 = 0.5 * pf;
    int pre_rank = 0x7fffffff, above_rank = 0x7fffffff;     // (file rank << 11) | index
    double fade_t = pinf(), last_t = -pinf(), fsum = 0.0;
    for (int i = tm.rank(); i < n; i += Team::size) {
      double ti = v.t[i], fi = v.f[i];
      fsum += fi;
      int key = ((int)v.rank[i] << 11) | i;
      if (ti < pt) { if (key < pre_rank) pre_rank = key; if (fi > thr && key < above_rank) above_rank = key; }
      if (ti > pt) { last_t = fmax(last_t, ti); if (fi < t50) fade_t = fmin(fade_t, ti); }
    }
    pre_rank = tm.imin(pre_rank); above_rank = tm.imin(above_rank);
    fade_t = tm.min(fade_t); last_t = tm.max(last_t); fsum = tm.sum(fsum);
    if (!is_nan(pt) && !is_nan(pf)) {
      if (pre_rank != 0x7fffffff) a_rise = pt - v.t[(above_rank != 0x7fffffff ? above_rank : pre_rank) & 2047];
      if (last_t > -pinf()) a_fade = ((fade_t < pinf()) ? fade_t : last_t) - pt;
    }
    loglog_decay(tm, v.t, v.f, n, pt, pf, sa, sb, a_alpha, a_resid);
    for (int i = tm.rank(); i < n; i += Team::size) sa[i] = v.f[i];
    tm.sync();
    tm.sort(sa, n);
    pct[0] = sorted_percentile(sa, n, 10.0); pct[1] = sorted_percentile(sa, n, 25.0);
    pct[2] = sorted_percentile(sa, n, 75.0); pct[3] = sorted_percentile(sa, n, 90.0);
    if (pf > 0.0) conc = pf / (fsum + 1e-6);
    tm.sync();
  }
  if (tm.rank() == 0) {
    // cross-band (:252-284)
    int np = 0; double lo = pinf(), hi = -pinf();
    for (int b = 0; b < N_BAND; ++b) if (!is_nan(peak_t[b])) { ++np; lo = fmin(lo, peak_t[b]); hi = fmax(hi, peak_t[b]); }
    if (np >= 2) {
      int cnt; double mean, sd;
      small_stats(peak_t, N_BAND, cnt, mean, sd);
      row[48] = hi - lo; row[49] = sd;
    }
    const int opt[3] = {BG, BR, BI};
    double rises[3], fades[3], alphas[3];
    for (int j = 0; j < 3; ++j) { rises[j] = row[8 * opt[j]]; fades[j] = row[8 * opt[j] + 1]; alphas[j] = row[8 * opt[j] + 6]; }
    int cr, cf, ca; double mr, mf, ma, sr, sf, sal;
    small_stats(rises, 3, cr, mr, sr);
    small_stats(fades, 3, cf, mf, sf);
    small_stats(alphas, 3, ca, ma, sal);
    row[50] = cr ? mr : qnan(); row[51] = cf ? mf : qnan(); row[52] = ca ? ma : qnan();
    if (cr >= 2) row[53] = sr / (mr + 1e-6);
    if (cf >= 2) row[54] = sf / (mf + 1e-6);
    row[55] = a_rise; row[56] = a_fade;
    row[57] = (!is_nan(a_rise) && !is_nan(a_fade) && a_fade > 0.0) ? a_rise / a_fade : qnan();
    row[58] = a_alpha; row[59] = a_resid;
    row[60] = pct[0]; row[61] = pct[1]; row[62] = pct[2]; row[63] = pct[3];
    row[64] = conc;
  }
  tm.sync();
}

}  // namespace mlrn
