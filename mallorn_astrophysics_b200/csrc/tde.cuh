// tde.cuh -- TDE-physics features (25 columns, schema.TDE_COLUMNS).
// Behaviour follows reference src/features/tde_physics.py:25-374: nearest-neighbour colour
// pairing (|dt| < 5 d, first argmin), late-time log-log slopes / flux ratios / rebrightening,
// rise shape and rate, g-r temperature stability (|dt| < 3 d) and the r-band decay power law.
// All band views are time-sorted.
#pragma once
#include "objview.cuh"
#include "stat.cuh"

namespace mlrn {

constexpr int TDE_COLS = 25;

// index of the band-2 point nearest in time (first minimum of |t2 - a|), as numpy.argmin
MLRN_TEAM_FN int nearest_first(const double* t2, int n2, double a, double& dmin) {
  int best = 0; dmin = fabs(t2[0] - a);
  for (int j = 1; j < n2; ++j) { double d = fabs(t2[j] - a); if (d < dmin) { dmin = d; best = j; } }
  return best;
}

// Collects, in band-1 time order, (value, time) of the pairs that pass; value from `map`.
// MODE 0: magnitude colour (tde_physics.py:52-66); MODE 1: temperature proxy (:234-258).
template <int MODE, class Team>
MLRN_TEAM_FN int paired_series(const Team& tm, const ObjView& v, int b1, int b2, double window, double* vals, double* times) {
  const int n1 = v.cnt(b1), n2 = v.cnt(b2);
  const double* t1 = v.t + v.off[b1]; const double* f1 = v.f + v.off[b1];
  const double* t2 = v.t + v.off[b2]; const double* f2 = v.f + v.off[b2];
  int cnt = 0;
  for (int base = 0; base < n1; base += Team::size) {
    int i = base + tm.rank();
    bool keep = false; double val = 0.0;
    if (i < n1) {
      double d; int j = nearest_first(t2, n2, t1[i], d);
      if (d < window && f1[i] > 0.0 && f2[j] > 0.0) {
        double c = -2.5 * team_log10(f1[i] / f2[j]);
        if (MODE == 0) val = c;
        else val = (c < -0.5) ? 40000.0 : ((c > 1.5) ? 5000.0 : 7000.0 / (c + 0.5));
        keep = true;
      }
    }
    int slot = tm.compact_slot(keep, cnt);
    if (keep) { vals[slot] = val; times[slot] = t1[i]; }
  }
  tm.sync();
  return cnt;
}

// sa..sd: four scratch arrays with capacity >= next_pow2(v.n)
template <class Team>
MLRN_FN void tde_row(const Team& tm, const ObjView& v, double* sa, double* sb, double* sc, double* sd, double* row) {
  for (int k = tm.rank(); k < TDE_COLS; k += Team::size) row[k] = qnan();
  tm.sync();
  // (i) colour variance / range / trend (:25-90)
  const int cp1[2] = {BG, BR}, cp2[2] = {BR, BI};
  for (int p = 0; p < 2; ++p) {
    if (v.cnt(cp1[p]) < 3 || v.cnt(cp2[p]) < 3) continue;
    int cnt = paired_series<0>(tm, v, cp1[p], cp2[p], 5.0, sa, sb);
    if (cnt >= 3) {
      double mean = team_mean(tm, sa, cnt);
      double var = team_var(tm, sa, cnt, mean);
      double rng = team_max(tm, sa, cnt) - team_min(tm, sa, cnt);
      double t0 = sb[0];
      tm.sync();
      for (int i = tm.rank(); i < cnt; i += Team::size) sb[i] -= t0;
      tm.sync();
      Line l = team_polyfit1(tm, sb, sa, cnt);
      put(tm, row, 3 * p, var); put(tm, row, 3 * p + 1, rng); put(tm, row, 3 * p + 2, l.slope * 100.0);
    }
    tm.sync();
  }
  // (ii) late-time behaviour of g, r, i (:93-155)
  const int gri[3] = {BG, BR, BI};
  for (int j = 0; j < 3; ++j) {
    const int b = gri[j], n = v.cnt(b);
    if (n < 5) continue;
    const double* t = v.t + v.off[b]; const double* f = v.f + v.off[b];
    const int k = team_argmax_first(tm, f, n);
    const double pt = t[k], pf = f[k];
    int cnt = 0;
    for (int base = 0; base < n; base += Team::size) {
      int i = base + tm.rank();
      bool keep = (i < n) && (t[i] > pt + 50.0);
      int slot = tm.compact_slot(keep, cnt);
      if (keep) { sa[slot] = team_log10(t[i] - pt + 1.0); sb[slot] = team_log10(fmax(f[i], 1e-10)); sc[slot] = f[i]; }
    }
    tm.sync();
    if (cnt >= 3 && pf > 0.0) {
      double mx = team_mean(tm, sa, cnt);
      double slope = qnan();
      if (sqrt(team_var(tm, sa, cnt, mx)) > 0.0) slope = team_polyfit1(tm, sa, sb, cnt).slope;
      double lmean = team_mean(tm, sc, cnt), lmax = team_max(tm, sc, cnt);
      put(tm, row, 6 + 3 * j, slope);
      put(tm, row, 7 + 3 * j, lmean / pf);
      put(tm, row, 8 + 3 * j, (lmean > 0.0) ? lmax / lmean : 1.0);
    }
    tm.sync();
  }
  // (iii) rise shape and rate of g, r (:158-206)
  for (int j = 0; j < 2; ++j) {
    const int b = gri[j], n = v.cnt(b);
    if (n < 5) continue;
    const double* t = v.t + v.off[b]; const double* f = v.f + v.off[b];
    const int k = team_argmax_first(tm, f, n);
    const double pf = f[k];
    const int nr = k + 1;
    if (nr >= 3 && pf > 0.0) {
      const double t0 = t[0], span = t[k] - t0;
      double sf = 0.0, stn = 0.0;
      for (int i = tm.rank(); i < nr; i += Team::size) { sf += f[i] / pf; stn += (t[i] - t0) / (span + 1e-6); }
      sf = tm.sum(sf) / (double)nr; stn = tm.sum(stn) / (double)nr;
      put(tm, row, 15 + 2 * j, (stn > 0.0) ? sf / stn : 1.0);
      put(tm, row, 16 + 2 * j, (t[k] > t0) ? pf / span : qnan());
    }
    tm.sync();
  }
  // (iv) temperature stability / trend / late-vs-peak from g-r pairs within 3 d (:209-284)
  if (v.cnt(BG) >= 3 && v.cnt(BR) >= 3) {
    int cnt = paired_series<1>(tm, v, BG, BR, 3.0, sa, sb);
    if (cnt >= 3) {
      double mean = team_mean(tm, sa, cnt);
      double sdv = sqrt(team_var(tm, sa, cnt, mean));
      double lvp = qnan();
      if (cnt > 4) {
        int npk = cnt / 4; if (npk < 2) npk = 2;
        double a = team_sum(tm, sa, npk) / (double)npk;
        double l = team_sum(tm, sa + (cnt - 3), 3) / 3.0;
        lvp = l / a;
      }
      double t0 = sb[0];
      tm.sync();
      for (int i = tm.rank(); i < cnt; i += Team::size) sb[i] -= t0;
      tm.sync();
      Line l = team_polyfit1(tm, sb, sa, cnt);
      put(tm, row, 19, sdv / mean); put(tm, row, 20, l.slope * 100.0); put(tm, row, 21, lvp);
    }
    tm.sync();
  }
  // (v) r-band decay power law (:287-352)
  {
    const int b = BR, n = v.cnt(b);
    if (n >= 5) {
      const double* t = v.t + v.off[b]; const double* f = v.f + v.off[b];
      const int k = team_argmax_first(tm, f, n);
      const double pt = t[k], pf = f[k];
      int npost = 0, cnt = 0, cl = 0;
      for (int base = 0; base < n; base += Team::size) {
        int i = base + tm.rank();
        bool post = (i < n) && (t[i] > pt);
        bool keep = post && (f[i] > 0.0);
        double dt = post ? fmax(t[i] - pt, 1.0) : 0.0;
        bool late = keep && (dt > 50.0);
        npost += post ? 1 : 0;
        int slot = tm.compact_slot(keep, cnt);
        double ldt = 0.0, lf = 0.0;
        if (keep) { ldt = team_log10(dt); lf = team_log10(f[i]); sa[slot] = ldt; sb[slot] = lf; }
        int sl = tm.compact_slot(late, cl);
        if (late) { sc[sl] = ldt; sd[sl] = lf; }                    // late implies keep: the same two logarithms
      }
      npost = tm.isum(npost);
      tm.sync();
      if (npost >= 4 && pf > 0.0 && cnt >= 3) {
        Line l = team_polyfit1(tm, sa, sb, cnt);
        // residual spread = numpy.std of (y - fit)
        for (int i = tm.rank(); i < cnt; i += Team::size) sb[i] -= (l.slope * sa[i] + l.icpt);
        tm.sync();
        double rm = team_mean(tm, sb, cnt);
        double rs = sqrt(team_var(tm, sb, cnt, rm));
        double al = qnan();
        if (cl >= 3) al = team_polyfit1(tm, sc, sd, cl).slope;
        put(tm, row, 22, l.slope); put(tm, row, 23, rs); put(tm, row, 24, al);
      }
      tm.sync();
    }
  }
  tm.sync();
}

}  // namespace mlrn
