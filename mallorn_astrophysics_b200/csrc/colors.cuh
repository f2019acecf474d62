// colors.cuh -- colour-evolution features (83 columns, schema.COLOR_COLUMNS).
// Behaviour follows reference src/features/colors.py:108-344: reference peak from the first flux
// maximum of r (else g, else i) in FILE order, linear interpolation of every band at ten epochs
// (bracketing gap <= 50 d), magnitude colours of four band pairs, slopes, variability (gap <= 5 d),
// peak fluxes / ratios / lags, curvature, late stability and the piecewise temperature proxy.
#pragma once
#include "objview.cuh"
#include "stat.cuh"

namespace mlrn {

constexpr int COLOR_COLS = 83;
constexpr int N_EPOCH = 10;
constexpr int N_PAIR = 4;

MLRN_FN_FORCE double mag_color(double f1, double f2) {            // colors.py:92-105
  if (is_nan(f1) || is_nan(f2) || f1 <= 0.0 || f2 <= 0.0) return qnan();
  return -2.5 * log10(f1 / f2);
}

// Linear interpolation on a time-sorted band (colors.py:47-89; numpy.searchsorted side='left').
MLRN_TEAM_FN double lerp_flux(const double* ts, const double* fs, int n, double target, double max_gap) {
  if (n < 2) return qnan();
  if (target < ts[0] || target > ts[n - 1]) return qnan();
  if (is_nan(target)) return fs[n - 1];                           // searchsorted puts NaN at the end
  int lo = 0, hi = n;                                             // first k with ts[k] >= target
  while (lo < hi) { int mid = (lo + hi) >> 1; if (ts[mid] < target) lo = mid + 1; else hi = mid; }
  if (lo == 0) return fs[0];
  if (lo == n) return fs[n - 1];
  double t1 = ts[lo - 1], t2 = ts[lo];
  if (t2 - t1 > max_gap) return qnan();
  double w = (target - t1) / (t2 - t1);
  return fs[lo - 1] + w * (fs[lo] - fs[lo - 1]);
}

MLRN_FN_FORCE double temp_from_gr(double c) {                      // colors.py:297-308
  if (is_nan(c)) return qnan();
  if (c < -0.5) return 50000.0;
  if (c > 2.0) return 3000.0;
  return 7000.0 / (c + 0.6);
}

// sa, sb: scratch (capacity >= max(64, n)); row: COLOR_COLS doubles
template <class Team>
MLRN_FN void color_row(const Team& tm, const ObjView& v, double* sa, double* sb, double* row) {
  const int pa[N_PAIR] = {BG, BR, BU, BI}, pb[N_PAIR] = {BR, BI, BG, BZ};
  const double epochs[N_EPOCH] = {0.0, 10.0, 20.0, 30.0, 50.0, 75.0, 100.0, 150.0, -10.0, -20.0};
  for (int k = tm.rank(); k < COLOR_COLS; k += Team::size) row[k] = qnan();
  tm.sync();
  // reference peak time: first argmax in file order of r, g, i (colors.py:131-148)
  double peak_t[N_BAND];
  for (int b = 0; b < N_BAND; ++b) peak_t[b] = qnan();
  const int pref[3] = {BR, BG, BI};
  double ref = qnan();
  for (int j = 2; j >= 0; --j) {
    int b = pref[j];
    if (v.cnt(b) > 0) {
      int k = argmax_file_order(tm, v, v.off[b], v.off[b + 1]);
      if (k >= 0) peak_t[b] = v.t[k];
      if (!is_nan(peak_t[b])) ref = peak_t[b];
    }
  }
  // interpolated flux of every band at every epoch -> sa[ei*6 + b]   (:168-182)
  for (int q = tm.rank(); q < N_EPOCH * N_BAND; q += Team::size) {
    int ei = q / N_BAND, b = q % N_BAND;
    double target = is_nan(ref) ? qnan() : ref + epochs[ei];
    int lo = v.off[b], n = v.cnt(b);
    sa[q] = (n > 0) ? lerp_flux(v.t + lo, v.f + lo, n, target, 50.0) : qnan();
  }
  tm.sync();
  for (int q = tm.rank(); q < N_EPOCH * N_PAIR; q += Team::size) {   // :185-189
    int ei = q / N_PAIR, p = q % N_PAIR;
    row[1 + q] = mag_color(sa[ei * N_BAND + pa[p]], sa[ei * N_BAND + pb[p]]);
  }
  tm.sync();
  // colour variability: band-1 epochs, band-2 interpolated with gap <= 5 d, need >= 3 (:211-232)
  for (int p = 0; p < N_PAIR; ++p) {
    int a = pa[p], c = pb[p];
    int na = v.cnt(a), nc = v.cnt(c);
    if (na == 0 || nc == 0) continue;
    int cnt = 0;
    for (int base = 0; base < na; base += Team::size) {
      int i = base + tm.rank();
      double col = qnan();
      if (i < na) col = mag_color(v.f[v.off[a] + i], lerp_flux(v.t + v.off[c], v.f + v.off[c], nc, v.t[v.off[a] + i], 5.0));
      bool keep = !is_nan(col);
      int slot = tm.compact_slot(keep, cnt);
      if (keep) sb[slot] = col;
    }
    tm.sync();
    if (cnt >= 3) {
      double mean = team_mean(tm, sb, cnt);
      double sd = sqrt(team_var(tm, sb, cnt, mean));
      double mx = team_max(tm, sb, cnt), mn = team_min(tm, sb, cnt);
      put(tm, row, 49 + 2 * p, sd);
      put(tm, row, 50 + 2 * p, mx - mn);
    }
    tm.sync();
  }
  // peak fluxes (:235-239)
  double pk[N_BAND];
  for (int b = 0; b < N_BAND; ++b) pk[b] = (v.cnt(b) > 0) ? team_max(tm, v.f + v.off[b], v.cnt(b)) : qnan();
  if (tm.rank() == 0) {
    row[0] = ref;
    const double* col = row + 1;                                   // col[ei*4 + p]
    for (int p = 0; p < N_PAIR; ++p) {                              // slopes (:192-207)
      double c0 = col[0 * 4 + p], c50 = col[4 * 4 + p], c100 = col[6 * 4 + p];
      row[41 + 2 * p] = (!is_nan(c0) && !is_nan(c50)) ? (c50 - c0) / 50.0 : qnan();
      row[42 + 2 * p] = (!is_nan(c0) && !is_nan(c100)) ? (c100 - c0) / 100.0 : qnan();
    }
    for (int b = 0; b < N_BAND; ++b) row[57 + b] = pk[b];
    for (int p = 0; p < N_PAIR; ++p) {                              // :242-248
      double f1 = pk[pa[p]], f2 = pk[pb[p]];
      row[63 + p] = (!is_nan(f1) && !is_nan(f2) && f2 > 0.0) ? f1 / f2 : qnan();
    }
    row[67] = (v.cnt(BG) > 0 && v.cnt(BR) > 0) ? peak_t[BG] - peak_t[BR] : qnan();   // :252-257
    row[68] = (v.cnt(BR) > 0 && v.cnt(BI) > 0) ? peak_t[BR] - peak_t[BI] : qnan();
    for (int p = 0; p < 2; ++p) {                                   // curvature (:263-275)
      double c0 = col[0 * 4 + p], c30 = col[3 * 4 + p], c75 = col[5 * 4 + p];
      if (!is_nan(c0) && !is_nan(c30) && !is_nan(c75))
        row[69 + p] = ((c75 - c30) / 45.0 - (c30 - c0) / 30.0) / 37.5;
    }
    for (int p = 0; p < 2; ++p) {                                   // late stability (:279-293)
      double late[4] = {col[4 * 4 + p], col[5 * 4 + p], col[6 * 4 + p], col[7 * 4 + p]};
      int cnt; double mean, sd;
      small_stats(late, 4, cnt, mean, sd);
      if (cnt >= 2) { row[71 + 2 * p] = sd; row[72 + 2 * p] = mean; }
    }
    double tp[4] = {temp_from_gr(col[0]), temp_from_gr(col[3 * 4]), temp_from_gr(col[5 * 4]), temp_from_gr(col[7 * 4])};
    for (int k = 0; k < 4; ++k) row[75 + k] = tp[k];                // :310-313
    if (!is_nan(tp[0]) && !is_nan(tp[1])) row[79] = (tp[1] - tp[0]) / 30.0;   // :321-334
    if (!is_nan(tp[1]) && !is_nan(tp[2])) row[80] = (tp[2] - tp[1]) / 45.0;
    if (!is_nan(tp[2]) && !is_nan(tp[3])) row[81] = (tp[3] - tp[2]) / 75.0;
    int cnt; double mean, sd;
    small_stats(tp, 4, cnt, mean, sd);
    if (cnt >= 2) row[82] = sd / mean;                              // :337-342
  }
  tm.sync();
}

}  // namespace mlrn
