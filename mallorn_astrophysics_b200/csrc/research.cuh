// research.cuh -- "research" feature block (40 columns, schema.RESEARCH_COLUMNS): SURVEY 8(f) row 3.
// Behaviour follows reference src/features/research_features.py:44-560: log-log power-law decay quality per band
// (g, r, i; :44-156), nuclear-position proxies from the r band (:163-247), colour at peak and its drift (file-order
// argmax / argmin, :254-327), Mexican-hat power spectra of the r band interpolated on a 1-day grid at scales
// 10 / 30 / 100 d (:334-425) and luminosity features from the redshift (:432-527).
// Band views are time-sorted (ties by file rank); where the reference works in file order the stored rank is used.
// STATUS: template verified against the reference on the CPU (tests/hostsim); the CUDA pass that instantiates it is
// not exported yet.
#pragma once
#include "objview.cuh"
#include "stat.cuh"

namespace mlrn {

constexpr int RESEARCH_COLS = 40;
constexpr int MHPS_WAVELET_CAP = 512;           // 5 x 100 d is the longest wavelet (:397)

// population std (numpy.std) of a[0..n)
template <class Team>
MLRN_FN double team_std(const Team& tm, const double* a, int n) { return sqrt(team_var(tm, a, n, team_mean(tm, a, n))); }

// research_features.py:44-117 on one time-sorted band; out[6] = alpha, |alpha+5/3|, |alpha+5/12|, chi2, residual std, success
template <class Team>
MLRN_FN void research_power_law(const Team& tm, const double* t, const double* f, const double* e, int n,
                                double* sa, double* sb, double* sc, double* out) {
  for (int k = 0; k < 5; ++k) out[k] = qnan();
  out[5] = 0.0;
  const int pk = team_argmax_first(tm, f, n);
  const double pt = t[pk];
  int cnt = 0;
  for (int base = 0; base < n; base += Team::size) {
    const int i = base + tm.rank();
    const bool keep = i < n && t[i] > pt + 10.0 && f[i] > 0.0;
    const int slot = tm.compact_slot(keep, cnt);
    if (keep) {
      sa[slot] = log10(t[i] - pt); sb[slot] = log10(f[i]);
      sc[slot] = fmin(fmax(e[i] / (f[i] * 2.302585092994046 + 1e-10), 0.01), 1.0);        // clipped error in log space
    }
  }
  tm.sync();
  if (cnt < 4) return;
  // np.polyfit(..., cov=True) raises (and the reference swallows it) when the abscissae do not span a line
  const double mx = team_mean(tm, sa, cnt);
  double sxx = 0.0;
  for (int i = tm.rank(); i < cnt; i += Team::size) { const double d = sa[i] - mx; sxx += d * d; }
  sxx = tm.sum(sxx);
  if (!(sxx > 0.0)) return;
  const Line ln = team_polyfit1(tm, sa, sb, cnt);
  if (!is_finite(ln.slope)) return;
  double chi = 0.0;
  for (int i = tm.rank(); i < cnt; i += Team::size) {
    const double r = sb[i] - (ln.slope * sa[i] + ln.icpt);
    sb[i] = r;
    const double q = r / sc[i];
    chi += q * q;
  }
  chi = tm.sum(chi);
  tm.sync();
  out[0] = ln.slope; out[1] = fabs(ln.slope + 5.0 / 3.0); out[2] = fabs(ln.slope + 5.0 / 12.0);
  out[3] = chi / (double)((cnt - 2) > 1 ? (cnt - 2) : 1);
  out[4] = team_std(tm, sb, cnt);
  out[5] = 1.0;
  tm.sync();
}

// index of the band point nearest to `at` among those with |t - at| < window, first in FILE order on ties; -1 if none
template <class Team>
MLRN_FN int nearest_in_window_file_order(const Team& tm, const ObjView& v, int b, double at, double window) {
  const int lo = v.off[b], hi = v.off[b + 1];
  double dmin = pinf();
  for (int i = lo + tm.rank(); i < hi; i += Team::size) { const double d = fabs(v.t[i] - at); if (d < window) dmin = fmin(dmin, d); }
  dmin = tm.min(dmin);
  if (!(dmin < window)) return -1;
  int best = 0x7fffffff;
  for (int i = lo + tm.rank(); i < hi; i += Team::size)
    if (fabs(v.t[i] - at) == dmin) { const int key = ((int)v.rank[i] << 11) | (i - lo); if (key < best) best = key; }
  best = tm.imin(best);
  return lo + (best & 2047);
}

// sa, sb, sc: scratch with capacity >= next_pow2(v.n); grid: capacity gcap doubles (1-day grid of the r band);
// wav: MHPS_WAVELET_CAP doubles.  z: redshift (NaN = unknown).
template <class Team>
MLRN_FN void research_row(const Team& tm, const ObjView& v, double z, double* sa, double* sb, double* sc,
                          double* grid, int gcap, double* wav, double* row) {
  for (int k = tm.rank(); k < RESEARCH_COLS; k += Team::size) row[k] = qnan();
  tm.sync();
  // ---- 1. power-law decay quality (:120-156)
  {
    const int bands[3] = {BG, BR, BI};
    double alphas[3]; int na = 0;
    for (int p = 0; p < 3; ++p) {
      const int b = bands[p], n = v.cnt(b);
      if (n < 5) continue;                                           // six NaN
      double out[6];
      research_power_law(tm, v.t + v.off[b], v.f + v.off[b], v.e + v.off[b], n, sa, sb, sc, out);
      for (int k = 0; k < 6; ++k) put(tm, row, 6 * p + k, out[k]);
      if (!is_nan(out[0])) alphas[na++] = out[0];
    }
    if (na >= 2) {
      int c; double m, sd;
      small_stats(alphas, na, c, m, sd);
      double dev = 0.0;
      for (int k = 0; k < na; ++k) dev += fabs(alphas[k] + 5.0 / 3.0);
      put(tm, row, 18, m); put(tm, row, 19, sd); put(tm, row, 20, dev / na);
    } else if (na == 1) {
      put(tm, row, 18, alphas[0]); put(tm, row, 20, fabs(alphas[0] + 5.0 / 3.0));
    }
  }
  // ---- 2. nuclear-position proxies from the r band (:163-247)
  {
    const int n = v.cnt(BR);
    const double* t = v.t + v.off[BR]; const double* f = v.f + v.off[BR]; const double* e = v.e + v.off[BR];
    if (n >= 10) {
      double smooth = qnan(), conc = qnan(), ratio = qnan();
      for (int i = tm.rank(); i < n - 1; i += Team::size) sa[i] = fabs(f[i + 1] - f[i]) / ((t[i + 1] - t[i]) + 0.1);
      for (int i = tm.rank(); i < n; i += Team::size) sb[i] = e[i];
      tm.sync();
      tm.sort(sa, n - 1); tm.sort(sb, n);
      tm.sync();
      const double med_rate = sorted_median(sa, n - 1), med_err = sorted_median(sb, n);
      if (med_err > 0.0) smooth = 1.0 / (1.0 + med_rate / med_err);
      tm.sync();
      for (int i = tm.rank(); i < n; i += Team::size) { sa[i] = f[i]; sb[i] = fabs(f[i]) + 1.0; }
      tm.sync();
      tm.sort(sa, n);
      tm.sync();
      const double peak = sa[n - 1], base = sorted_percentile(sa, n, 10.0);
      if (base > 0.0) conc = peak / base;
      else if (peak > 0.0) { tm.sort(sb, n); tm.sync(); conc = peak / sorted_median(sb, n); }
      tm.sync();
      if (n >= 20) {
        double acc = 0.0; int windows = 0;
        for (int i = tm.rank(); i < n - 5; i += Team::size)
          if (t[i + 5] - t[i] < 15.0) {
            double m = 0.0;
            for (int k = 0; k < 5; ++k) m += f[i + k];
            m /= 5.0;
            double s2 = 0.0;
            for (int k = 0; k < 5; ++k) { const double d = f[i + k] - m; s2 += d * d; }
            acc += sqrt(s2 / 5.0); ++windows;
          }
        acc = tm.sum(acc); windows = tm.isum(windows);
        const double long_var = team_std(tm, f, n);
        if (windows > 0 && long_var > 0.0) ratio = (acc / windows) / long_var;
      }
      double s = 0.0; int c = 0;
      if (!is_nan(smooth)) { s += smooth; ++c; }
      if (!is_nan(conc)) { s += fmin(1.0, conc / 100.0); ++c; }
      if (!is_nan(ratio)) { s += 1.0 - fmin(1.0, ratio); ++c; }
      put(tm, row, 21, smooth); put(tm, row, 22, conc); put(tm, row, 23, ratio);
      if (c) put(tm, row, 24, s / c);
    }
  }
  // ---- 3. colour at peak and its drift (:254-327), file order
  {
    int pb = -1;
    if (v.cnt(BR) >= 3) pb = BR; else if (v.cnt(BG) >= 3) pb = BG;
    if (pb >= 0) {
      const double peak_time = v.t[argmax_file_order(tm, v, v.off[pb], v.off[pb + 1])];
      const int p1[2] = {BG, BR}, p2[2] = {BR, BI};
      for (int p = 0; p < 2; ++p) {
        const int b1 = p1[p], b2 = p2[p];
        if (v.cnt(b1) < 2 || v.cnt(b2) < 2) continue;
        const int i1 = nearest_in_window_file_order(tm, v, b1, peak_time, 10.0);
        const int i2 = nearest_in_window_file_order(tm, v, b2, peak_time, 10.0);
        if (i1 < 0 || i2 < 0) continue;
        const double fa = v.f[i1], fb = v.f[i2];
        if (!(fa > 0.0 && fb > 0.0)) continue;
        const double col = -2.5 * log10(fa / fb);
        put(tm, row, 25 + 2 * p, col);
        // late colours: every band-1 point later than peak + 50 d paired with the nearest late band-2 point
        double acc = 0.0; int cnt = 0;
        const int lo1 = v.off[b1], hi1 = v.off[b1 + 1], lo2 = v.off[b2], hi2 = v.off[b2 + 1];
        for (int i = lo1 + tm.rank(); i < hi1; i += Team::size) {
          if (!(v.t[i] > peak_time + 50.0)) continue;
          double dmin = pinf(); int best = -1; int brank = 0x7fffffff;
          for (int j = lo2; j < hi2; ++j) {
            if (!(v.t[j] > peak_time + 50.0)) continue;
            const double d = fabs(v.t[j] - v.t[i]);
            if (d < dmin || (d == dmin && (int)v.rank[j] < brank)) { dmin = d; best = j; brank = (int)v.rank[j]; }
          }
          if (best >= 0 && dmin < 5.0 && v.f[i] > 0.0 && v.f[best] > 0.0) { acc += -2.5 * log10(v.f[i] / v.f[best]); ++cnt; }
        }
        acc = tm.sum(acc); cnt = tm.isum(cnt);
        if (cnt > 0) put(tm, row, 26 + 2 * p, acc / cnt - col);
      }
    }
  }
  // ---- 4. Mexican-hat power spectra of the r band on a 1-day grid (:349-425)
  {
    const int n = v.cnt(BR);
    const double* t = v.t + v.off[BR]; const double* f = v.f + v.off[BR];
    if (n >= 20 && t[n - 1] - t[0] >= 50.0) {
      const double span = t[n - 1] - t[0];
      const int L = (int)ceil(span);                                  // len(np.arange(t0, t_last, 1.0))
      if (L <= gcap) {
        double s = 0.0;
        for (int m = tm.rank(); m < L; m += Team::size) {
          const double x = t[0] + (double)m;
          int lo = 0, hi = n - 1;                                     // largest j with t[j] <= x (x < t[n-1] on the grid)
          while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (t[mid] <= x) lo = mid; else hi = mid; }
          const double slope = (f[lo + 1] - f[lo]) / (t[lo + 1] - t[lo]);
          const double val = slope * (x - t[lo]) + f[lo];
          grid[m] = val; s += val;
        }
        s = tm.sum(s);
        tm.sync();
        const double mean = s / (double)L;
        for (int m = tm.rank(); m < L; m += Team::size) grid[m] -= mean;
        tm.sync();
        const double scales[3] = {10.0, 30.0, 100.0};
        double power[3] = {qnan(), qnan(), qnan()};
        for (int q = 0; q < 3; ++q) {
          int wl = (int)(5.0 * scales[q]); if (L / 2 < wl) wl = L / 2;
          if (wl < 5) continue;
          const double a = -(double)((wl + 1) / 2), b = (double)(wl / 2);        // -wl // 2 and wl // 2
          const double step = (b - a) / (double)(wl - 1);
          double nrm = 0.0;
          for (int k = tm.rank(); k < wl; k += Team::size) {
            const double x = ((k == wl - 1) ? b : (double)k * step + a) / scales[q];
            const double w = (1.0 - x * x) * exp(-(x * x) / 2.0);
            wav[k] = w; nrm += w * w;
          }
          nrm = sqrt(tm.sum(nrm));
          tm.sync();
          // 'same' convolution: out[i] = sum_k w[k] grid[i + (wl-1)/2 - k]
          const int off = (wl - 1) / 2;
          double acc = 0.0;
          for (int i = tm.rank(); i < L; i += Team::size) {
            const int c0 = i + off;
            int k0 = c0 - (L - 1); if (k0 < 0) k0 = 0;
            int k1 = c0; if (k1 > wl - 1) k1 = wl - 1;
            double cv = 0.0;
            for (int k = k0; k <= k1; ++k) cv += wav[k] * grid[c0 - k];
            cv /= nrm;
            acc += cv * cv;
          }
          power[q] = tm.sum(acc) / (double)L;
          put(tm, row, 29 + q, power[q]);
          tm.sync();
        }
        if (!is_nan(power[0]) && !is_nan(power[2]) && power[2] > 0.0) put(tm, row, 32, power[0] / power[2]);
        if (!is_nan(power[1]) && !is_nan(power[2]) && power[2] > 0.0) put(tm, row, 33, power[1] / power[2]);
        int dom = -1;
        for (int q = 0; q < 3; ++q) if (!is_nan(power[q]) && (dom < 0 || power[q] > power[dom])) dom = q;
        if (dom >= 0) put(tm, row, 34, scales[dom]);
      }
    }
  }
  // ---- 5. luminosity (:432-527), only for Z > 0 (:556)
  if (z > 0.0) {
    const double ch = 299792.458 / 70.0;
    const double dl = (z < 0.1) ? ch * z * (1.0 + z / 2.0) : ch * z * (1.0 + 0.5 * (1.0 - (0.5 * 0.3 - 0.7)) * z);
    put(tm, row, 35, dl);
    const int lo = v.off[BG], hi = v.off[BI + 1];                      // g, r, i are adjacent segments
    const int n = hi - lo;
    if (n >= 5) {
      const double d2 = dl * dl;
      for (int i = tm.rank(); i < n; i += Team::size) { sa[i] = v.f[lo + i] * d2; sb[i] = sa[i]; }
      tm.sync();
      const double mean = team_mean(tm, sa, n);
      // peak = first maximum in time order: largest value, then earliest time
      const double mx = team_max(tm, sa, n);
      double tp = pinf();
      for (int i = tm.rank(); i < n; i += Team::size) if (sa[i] == mx) tp = fmin(tp, v.t[lo + i]);
      tp = tm.min(tp);
      tm.sort(sb, n);
      tm.sync();
      put(tm, row, 36, mx); put(tm, row, 37, mx - sorted_percentile(sb, n, 10.0)); put(tm, row, 38, mean);
      tm.sync();
      // post-peak points (the peak included), time offsets and log luminosities
      int cnt = 0; int nonpos = 0;
      for (int base = 0; base < n; base += Team::size) {
        const int i = base + tm.rank();
        const bool keep = i < n && v.t[lo + i] >= tp;
        const int slot = tm.compact_slot(keep, cnt);
        if (keep) { sb[slot] = v.t[lo + i] - tp; sc[slot] = log10(sa[i]); if (!(sa[i] > 0.0)) ++nonpos; }
      }
      nonpos = tm.isum(nonpos);
      tm.sync();
      if (cnt > 5 && nonpos == 0) {                                   // peak_idx < n - 5, min(post) > 0
        const double mt = team_mean(tm, sb, cnt);
        if (team_var(tm, sb, cnt, mt) > 0.0) put(tm, row, 39, team_polyfit1(tm, sb, sc, cnt).slope * 100.0);
      }
    }
  }
  tm.sync();
}

}  // namespace mlrn
