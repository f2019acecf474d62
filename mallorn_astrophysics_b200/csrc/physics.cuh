// physics.cuh -- physics-based features (32 columns, schema.PHYSICS_COLUMNS).
// Behaviour follows reference src/features/physics_based.py:31-458: Stetson J (nearest pair within
// 0.5 d, both errors > 0) and K, the r-band structure function over all pairs with overlapping
// tau bins, rest-frame durations, the clipped temperature proxy, the "Bazin-like" heuristics
// (sequential latch/break loops reproduced literally) and SNR / excess variance over every row.
#pragma once
#include "objview.cuh"
#include "stat.cuh"
#include "tde.cuh"

namespace mlrn {

constexpr int PHYSICS_COLS = 32;

template <class Team>
MLRN_TEAM_FN double stetson_j(const Team& tm, const ObjView& v, int b1, int b2) {     // physics_based.py:49-82
  const int n1 = v.cnt(b1), n2 = v.cnt(b2);
  const double *t1 = v.t + v.off[b1], *f1 = v.f + v.off[b1], *e1 = v.e + v.off[b1];
  const double *t2 = v.t + v.off[b2], *f2 = v.f + v.off[b2], *e2 = v.e + v.off[b2];
  double m1 = team_mean(tm, f1, n1), m2 = team_mean(tm, f2, n2);
  double s1 = sqrt(team_var(tm, f1, n1, m1)), s2 = sqrt(team_var(tm, f2, n2, m2));
  if (s1 == 0.0 || s2 == 0.0) return 0.0;
  double acc = 0.0; int np = 0;
  for (int i = tm.rank(); i < n1; i += Team::size) {
    double d; int j = nearest_first(t2, n2, t1[i], d);
    if (d <= 0.5 && e1[i] > 0.0 && e2[j] > 0.0) {
      double p = ((f1[i] - m1) / e1[i]) * ((f2[j] - m2) / e2[j]);
      double sg = (p > 0.0) ? 1.0 : ((p < 0.0) ? -1.0 : 0.0);
      acc += sg * sqrt(fabs(p));
      ++np;
    }
  }
  acc = tm.sum(acc); np = tm.isum(np);
  return np ? acc / (double)np : qnan();
}

template <class Team>
MLRN_TEAM_FN double stetson_k(const Team& tm, const double* f, const double* e, int n) {   // :92-107
  if (n < 4) return qnan();
  double mean = team_mean(tm, f, n);
  double s1 = 0.0, s2 = 0.0; int nv = 0;
  for (int i = tm.rank(); i < n; i += Team::size)
    if (e[i] > 0.0) { double d = fabs(f[i] - mean) / e[i]; s1 += d; s2 += d * d; ++nv; }
  s1 = tm.sum(s1); s2 = tm.sum(s2); nv = tm.isum(nv);
  if (nv < 4) return qnan();
  return s1 / sqrt(s2) / sqrt((double)n);
}

MLRN_TEAM_FN double phys_temperature(double g, double r, double i) {                  // :179-199
  if (is_nan(g) || is_nan(r) || is_nan(i) || g <= 0.0 || r <= 0.0 || i <= 0.0) return qnan();
  double c = -2.5 * log10(g / r);
  double temp = (c < -0.5) ? 50000.0 : ((c > 2.0) ? 3000.0 : 7000.0 / (c + 0.6));
  return fmin(fmax(temp, 3000.0), 100000.0);
}

// sa, sb: scratch with capacity >= next_pow2(v.n)
template <class Team>
MLRN_FN void physics_row(const Team& tm, const ObjView& v, double redshift, double* sa, double* sb, double* row) {
  for (int k = tm.rank(); k < PHYSICS_COLS; k += Team::size) row[k] = qnan();
  tm.sync();
  const int gri[3] = {BG, BR, BI};
  bool has[N_BAND];
  for (int b = 0; b < N_BAND; ++b) has[b] = v.cnt(b) >= 3;                      // :307-314
  const int ja[3] = {BG, BR, BG}, jb[3] = {BR, BI, BI};
  for (int p = 0; p < 3; ++p)
    if (has[ja[p]] && has[jb[p]]) { double j = stetson_j(tm, v, ja[p], jb[p]); put(tm, row, p, j); }
  for (int p = 0; p < 3; ++p)
    if (has[gri[p]]) { double k = stetson_k(tm, v.f + v.off[gri[p]], v.e + v.off[gri[p]], v.cnt(gri[p])); put(tm, row, 3 + p, k); }
  // structure function of the r band (:128-168)
  if (has[BR] && v.cnt(BR) >= 5) {
    const int n = v.cnt(BR);
    const double* t = v.t + v.off[BR]; const double* f = v.f + v.off[BR];
    const double taus[5] = {1.0, 5.0, 10.0, 30.0, 100.0};
    double acc[5] = {0, 0, 0, 0, 0}; int cnt[5] = {0, 0, 0, 0, 0};
    for (int i = tm.rank(); i < n; i += Team::size)
      for (int j = i + 1; j < n; ++j) {
        double dt = fabs(t[j] - t[i]), d = f[j] - f[i];
        double d2 = d * d;
#pragma unroll
        for (int q = 0; q < 5; ++q) if (dt >= 0.5 * taus[q] && dt <= 1.5 * taus[q]) { acc[q] += d2; ++cnt[q]; }
      }
    double sfv[5]; double xs[5], ys[5]; int nv = 0;
    for (int q = 0; q < 5; ++q) {
      double a = tm.sum(acc[q]); int c = tm.isum(cnt[q]);
      sfv[q] = (c >= 3) ? sqrt(a / (double)c) : qnan();
      if (!is_nan(sfv[q]) && sfv[q] > 0.0) { xs[nv] = log10(taus[q]); ys[nv] = log10(sfv[q]); ++nv; }
    }
    if (tm.rank() == 0) {
      for (int q = 0; q < 5; ++q) row[6 + q] = sfv[q];
      if (nv >= 3) {
        double mx = 0, my = 0;
        for (int q = 0; q < nv; ++q) { mx += xs[q]; my += ys[q]; }
        mx /= nv; my /= nv;
        double sxx = 0, sxy = 0;
        for (int q = 0; q < nv; ++q) { sxx += (xs[q] - mx) * (xs[q] - mx); sxy += (xs[q] - mx) * (ys[q] - my); }
        row[11] = sxy / sxx;
      }
    }
  }
  tm.sync();
  // rest-frame timescales (:348-379)
  const double zz = is_nan(redshift) ? 0.0 : redshift;
  int pk_idx[3] = {-1, -1, -1};
  for (int p = 0; p < 3; ++p) {
    const int b = gri[p];
    if (!has[b]) continue;
    const int n = v.cnt(b);
    const double* t = v.t + v.off[b];
    const int k = team_argmax_first(tm, v.f + v.off[b], n);
    pk_idx[p] = k;
    if (tm.rank() == 0) {
      row[12 + 3 * p] = (t[n - 1] - t[0]) / (1.0 + zz);
      row[13 + 3 * p] = (k > 0) ? (t[k] - t[0]) / (1.0 + zz) : qnan();
      row[14 + 3 * p] = (k < n - 1) ? (t[n - 1] - t[k]) / (1.0 + zz) : qnan();
    }
  }
  // temperatures (:383-423)
  if (has[BG] && has[BR] && has[BI]) {
    double pk[3], late[3];
    const double target = v.t[v.off[BR] + pk_idx[1]] + 50.0;
    for (int p = 0; p < 3; ++p) {
      const int b = gri[p], n = v.cnt(b);
      pk[p] = v.f[v.off[b] + pk_idx[p]];
      double dbest = 0.0; int jb2 = -1;
      for (int i = tm.rank(); i < n; i += Team::size) {
        double d = fabs(v.t[v.off[b] + i] - target);
        if (jb2 < 0 || d < dbest) { dbest = d; jb2 = i; }
      }
      int j = tm.argmin_first(dbest, jb2);
      double dj = fabs(v.t[v.off[b] + j] - target);
      late[p] = (dj < 20.0) ? v.f[v.off[b] + j] : qnan();
    }
    if (tm.rank() == 0) {
      double t_pk = phys_temperature(pk[0], pk[1], pk[2]);
      double t_late = phys_temperature(late[0], late[1], late[2]);
      row[21] = t_pk; row[22] = t_late;
      row[23] = (!is_nan(t_pk) && !is_nan(t_late)) ? (t_late - t_pk) / 50.0 : qnan();
    }
  }
  // "Bazin-like" heuristics on the r band (:210-289); sequential logic kept literal
  if (has[BR] && v.cnt(BR) >= 5 && tm.rank() == 0) {
    const int n = v.cnt(BR), k = pk_idx[1];
    const double* t = v.t + v.off[BR]; const double* f = v.f + v.off[BR];
    const double pt = t[k], pf = f[k];
    row[24] = pf; row[25] = pt;
    if (k + 1 >= 2) {
      double t10 = t[0], t90 = pt;
      for (int i = 0; i <= k; ++i) {
        if (f[i] >= 0.1 * pf && t10 == t[0]) t10 = t[i];
        if (f[i] >= 0.9 * pf) { t90 = t[i]; break; }
      }
      row[26] = t90 - t10;
    }
    const int npost = n - k;
    if (npost >= 3) {
      const double target = pf / 2.718281828459045;
      double fall = qnan();
      for (int i = k; i < n; ++i) if (f[i] <= target) { fall = t[i] - pt; break; }
      if (is_nan(fall)) fall = (t[n - 1] - pt) * pf / (pf - f[n - 1] + 1e-6);
      row[27] = fall;
    }
    if (npost >= 5) {
      const int mid = npost / 2;
      double early = 0.0, late = 0.0;
      for (int i = 0; i < mid; ++i) early += f[k + i];
      for (int i = mid; i < npost; ++i) late += f[k + i];
      early /= (double)mid; late /= (double)(npost - mid);
      if (early > 0.0) row[28] = late / early;
    }
  }
  tm.sync();
  // SNR statistics and excess variance over every row with err > 0 and flux > 0 (:437-456)
  {
    int cnt = 0;
    for (int base = 0; base < v.n; base += Team::size) {
      int i = base + tm.rank();
      bool keep = (i < v.n) && (v.e[i] > 0.0) && (v.f[i] > 0.0);
      int slot = tm.compact_slot(keep, cnt);
      if (keep) { sa[slot] = v.f[i] / v.e[i]; sb[slot] = v.f[i]; }
    }
    tm.sync();
    if (cnt > 0) {
      double msnr = team_mean(tm, sa, cnt);
      double mf = team_mean(tm, sb, cnt);
      double vf = team_var(tm, sb, cnt, mf);
      double noise = 0.0;
      for (int i = tm.rank(); i < v.n; i += Team::size)
        if (v.e[i] > 0.0 && v.f[i] > 0.0) noise += v.e[i] * v.e[i];
      noise = tm.sum(noise) / (double)cnt;
      tm.sync();
      tm.sort(sa, cnt);
      double med = sorted_median(sa, cnt);
      double ev = (vf - noise) / (mf * mf);
      put(tm, row, 29, msnr); put(tm, row, 30, med); put(tm, row, 31, (ev > 0.0) ? ev : 0.0);   // max(0, ev): NaN -> 0
    }
    tm.sync();
  }
}

}  // namespace mlrn
