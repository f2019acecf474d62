// stat.cuh -- the moment pass, team-generic statement: 17 statistics x (6 bands + all) + 3 flux ratios + peak_band.
// Mirrors the behaviour of reference src/features/statistical.py:41-132 (compute_band_statistics)
// and :185-222 (cross-band tail); output order = schema.STAT_COLUMNS.
// The GPU kernel that ships is stat_fast.cu (register sorts, window-search MAD, merge of the band runs); this
// straightforward version defines the column layout constants both share and is what tests/hostsim instantiates
// on the CPU to check the semantics line by line against the reference goldens.
#pragma once
#include "objview.cuh"

namespace mlrn {

constexpr int N_STAT = 17;
constexpr int STAT_COLS = 7 * N_STAT + 4;   // 123

template <class Team>
MLRN_FN void put(const Team& tm, double* row, int k, double v) { if (tm.rank() == 0) row[k] = v; }

// One segment.  (t, f, e) are n observations; if !time_sorted the routine sorts a copy by time
// for max_slope (statistical.py:99-104).  sa / sb: scratch of capacity next_pow2(n).
template <class Team>
MLRN_FN void segment_stats(const Team& tm, const double* t, const double* f, const double* e, int n,
                           bool time_sorted, double* sa, double* sb, double* out) {
  if (n == 0) {                                              // statistical.py:56-66
    put(tm, out, 0, 0.0);
    for (int k = 1 + tm.rank(); k < N_STAT; k += Team::size) out[k] = qnan();
    tm.sync();
    return;
  }
  const double dn = (double)n;
  // ---- pass 1: sum, min, max, snr, time extent
  double s = 0.0, mn = pinf(), mx = -pinf(), tmin = pinf(), tmax = -pinf(), snr = 0.0;
  int nsnr = 0;
  for (int i = tm.rank(); i < n; i += Team::size) {
    double x = f[i], er = e[i], tt = t[i];
    s += x; mn = fmin(mn, x); mx = fmax(mx, x);
    tmin = fmin(tmin, tt); tmax = fmax(tmax, tt);
    if (er > 0.0) { snr += fabs(x) / er; ++nsnr; }
  }
  s = tm.sum(s); mn = tm.min(mn); mx = tm.max(mx); tmin = tm.min(tmin); tmax = tm.max(tmax);
  snr = tm.sum(snr); nsnr = tm.isum(nsnr);
  const double mean = s / dn;
  // ---- pass 2: central moments about the mean (numpy.std is two-pass as well)
  double m2 = 0.0;
  for (int i = tm.rank(); i < n; i += Team::size) { double d = f[i] - mean; m2 += d * d; }
  m2 = tm.sum(m2);
  const double sd = (n > 1) ? sqrt(m2 / dn) : 0.0;           // :71
  double skew = 0.0, kurt = 0.0;
  int b1 = 0, b2 = 0;
  if (sd > 0.0) {
    double z3 = 0.0, z4 = 0.0;
    for (int i = tm.rank(); i < n; i += Team::size) {
      double z = (f[i] - mean) / sd;
      double z2 = z * z;
      z3 += z2 * z; z4 += z2 * z2;
      double za = fabs(f[i] - mean) / sd;                    // :91-93, strict >
      b1 += (za > 1.0); b2 += (za > 2.0);
    }
    z3 = tm.sum(z3); z4 = tm.sum(z4); b1 = tm.isum(b1); b2 = tm.isum(b2);
    if (n > 2) skew = z3 / dn;                               // gates of :14-35 and :77
    if (n > 3) kurt = z4 / dn - 3.0;
  }
  // ---- max |df/dt| over consecutive time-sorted pairs with dt > 0, and mean cadence
  double slope = 0.0, cad = 0.0;
  if (n > 1) {
    const double* ts = t; const double* fs = f;
    if (!time_sorted) {
      for (int i = tm.rank(); i < n; i += Team::size) { sa[i] = t[i]; sb[i] = f[i]; }
      tm.sync();
      tm.sort_by_key(sa, sb, n);
      ts = sa; fs = sb;
    }
    for (int i = tm.rank(); i < n - 1; i += Team::size) {
      double dt = ts[i + 1] - ts[i];
      cad += dt;
      if (dt > 0.0) slope = fmax(slope, fabs((fs[i + 1] - fs[i]) / dt));
    }
    slope = tm.max(slope);
    cad = tm.sum(cad) / (double)(n - 1);
    tm.sync();
  }
  // ---- order statistics: median, IQR, MAD via sorted copies
  for (int i = tm.rank(); i < n; i += Team::size) sa[i] = f[i];
  tm.sync();
  tm.sort(sa, n);
  const double med = sorted_median(sa, n);
  const double iqr = (n > 1) ? sorted_percentile(sa, n, 75.0) - sorted_percentile(sa, n, 25.0) : 0.0;
  tm.sync();
  for (int i = tm.rank(); i < n; i += Team::size) sb[i] = fabs(sa[i] - med);
  tm.sync();
  tm.sort(sb, n);
  const double mad = sorted_median(sb, n);
  tm.sync();
  if (tm.rank() == 0) {
    out[0] = dn; out[1] = mean; out[2] = sd; out[3] = mn; out[4] = mx; out[5] = med;
    out[6] = skew; out[7] = kurt; out[8] = mx - mn; out[9] = mad; out[10] = iqr;
    out[11] = (sd > 0.0) ? (double)b1 / dn : 0.0;
    out[12] = (sd > 0.0) ? (double)b2 / dn : 0.0;
    out[13] = slope;
    out[14] = (nsnr > 0) ? snr / (double)nsnr : qnan();      // :116-120
    out[15] = (n > 1) ? tmax - tmin : 0.0;
    out[16] = cad;
  }
  tm.sync();
}

// Whole 123-column row for one staged object.  row: team-visible scratch of STAT_COLS doubles.
template <class Team>
MLRN_FN void stat_row(const Team& tm, const ObjView& v, double* sa, double* sb, double* row) {
  for (int b = 0; b < N_BAND; ++b) {
    int lo = v.off[b];
    segment_stats(tm, v.t + lo, v.f + lo, v.e + lo, v.cnt(b), true, sa, sb, row + b * N_STAT);
  }
  segment_stats(tm, v.t, v.f, v.e, v.n, false, sa, sb, row + 6 * N_STAT);   // every row, :186-190
  if (tm.rank() == 0) {
    const int k = 7 * N_STAT;
    const int pa[3] = {BG, BR, BI}, pb[3] = {BR, BI, BZ};
    for (int j = 0; j < 3; ++j) {                                           // :201-214
      double a = row[pa[j] * N_STAT + 1], c = row[pb[j] * N_STAT + 1];
      row[k + j] = (!is_nan(a) && c > 0.0) ? a / c : qnan();
    }
    int pk = -1; double best = 0.0;                                         // :217-222
    for (int b = 0; b < N_BAND; ++b) {
      double m = row[b * N_STAT + 4];
      if (!is_nan(m) && (pk < 0 || m > best)) { best = m; pk = b; }
    }
    row[k + 3] = (double)pk;
  }
  tm.sync();
}

}  // namespace mlrn
