// stat_warp.cuh -- K2, the moment pass for objects of at most 255 observations (the survey's ordinary light curves):
// reference src/features/statistical.py:41-226, 123 columns (schema.STAT_COLUMNS).  One warp per object, 3.3 k warp
// instructions per object (ncu, 150 observations) where the first two builds of this pass (per-band bitonic networks, one
// band at a time) needed 12-15 k -- the pass is instruction-bound, so the instruction count IS its distance from the HBM
// roofline: 25.8 -> 4.9 ms per 10^6 objects, 0.029 -> 0.154 of the measured HBM peak.
//
// What makes it short:
//   * every per-observation quantity that costs a division (|f|/err, |df/dt| to the next observation of the band and to the
//     next observation of the object) is produced once, in the FLAT phase, where lane l owns observations l, l+32, ... and
//     all 32 lanes are busy whatever the band lengths are;
//   * the per-band sums run in the GROUP phase: four lanes per segment (u g r i z y, the unknown-filter segment, and a
//     spare group that carries the all-band block), two-step butterflies instead of seven five-step ones per quantity; the
//     all-band central moments are assembled from the band moments (binomial shift to the common mean), not re-summed;
//   * ONE sort per object instead of seven: the fluxes are quantised to 24 bits over the object's range and packed with
//     the observation index into a 32-bit key, so a compare-exchange of the register bitonic network (eight keys per
//     lane) is a shuffle and a min/max.  Keys that tie in 24 bits may be misordered; they are adjacent, and odd-even
//     transposition passes on the full float64 values (usually none, rarely one) repair them, so the order statistics are
//     those of the exact sort.  The seven per-band sorted runs then fall out of a stable partition of the sorted sequence
//     by band (packed 8-bit counters, one warp scan);
//   * medians, quartiles and the MAD (closest-window rule, binary search) are read off the sorted runs by the group
//     leaders, eight segments at a time; reciprocals that only scale an output are a MUFU seed with two Newton steps, mean
//     and standard deviation -- which feed strict comparisons -- use IEEE division and square root.
//
// Anything unusual -- more than 255 observations, a non-finite time or flux -- is appended to a list and handled by the
// general kernel of stat_fast.cu (k_stat) in a second launch, which defines the semantics there.
#pragma once
#include <cuda_runtime.h>
#include "stat.cuh"

namespace mlrn {
namespace statw {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CAP = 256;                 // observations per object on this path: at most 255 (8-bit index, 8-bit band counters)
constexpr int MRG = CAP + CAP / 8;       // merged run, skewed by one slot per 8 (conflict-free blocked stores)
constexpr int ROW = 128;

struct WarpSmem {
  double st[CAP];          // times, store order (band-major, time-sorted inside a band); after the group phase: the output row
  double sf[CAP];          // fluxes, store order
  double sa[MRG];          // flat phase: |f|/err or -1; after the sort: all fluxes ascending (skewed index)
  double sb[CAP];          // flat phase: |df/dt| to the next observation of the same band or 0; after the sort: the band runs
  unsigned char inv[CAP];  // store index of the observation with a given time rank
  unsigned char sband[CAP];
  int soff[8];
};

__device__ __forceinline__ double frcp(double x) {            // 1/x: MUFU seed (>= 20 bits) + two Newton steps
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ unsigned fkey(float x) {           // order-preserving map float -> uint
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float funkey(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---- register bitonic network over 256 keys, position E = lane * 8 + a; every comparator puts the smaller key at the lower
// position (the first step of each merge pairs E with E ^ (k - 1), the others E with E ^ j)
__device__ __forceinline__ void cx(unsigned& lo, unsigned& hi) { const unsigned a = min(lo, hi), b = max(lo, hi); lo = a; hi = b; }
template <int J>
__device__ __forceinline__ void local_stage(unsigned (&v)[8]) {
#pragma unroll
  for (int a = 0; a < 8; ++a) if ((a & J) == 0 && (a ^ J) < 8) cx(v[a], v[a ^ J]);
}
template <int M>
__device__ __forceinline__ void local_flip(unsigned (&v)[8]) {       // partner a ^ M, M = 1, 3, 7
#pragma unroll
  for (int a = 0; a < 8; ++a) if ((a ^ M) > a) cx(v[a], v[a ^ M]);
}
// cross-lane step: partner lane = lane ^ m; `flip`: the partner's register is a ^ 7
__device__ __forceinline__ void cross_stage(unsigned (&v)[8], int m, bool flip, bool lower) {
  unsigned o[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) o[a] = __shfl_xor_sync(FULL, v[flip ? (a ^ 7) : a], m);
#pragma unroll
  for (int a = 0; a < 8; ++a) v[a] = lower ? min(v[a], o[a]) : max(v[a], o[a]);
}
__device__ __forceinline__ void sort256(unsigned (&v)[8], int lane) {
  local_flip<1>(v);
  local_flip<3>(v); local_stage<1>(v);
  local_flip<7>(v); local_stage<2>(v); local_stage<1>(v);
#pragma unroll
  for (int t = 1; t <= 5; ++t) {                      // merges of 16, 32, 64, 128, 256 keys
    const int m = (1 << t) - 1;
    cross_stage(v, m, true, (lane & (1 << (t - 1))) == 0);
#pragma unroll
    for (int s = t - 2; s >= 0; --s) cross_stage(v, 1 << s, false, (lane & (1 << s)) == 0);
    local_stage<4>(v); local_stage<2>(v); local_stage<1>(v);
  }
}

// One object.  `w`: the warp's shared-memory slice.  Returns false when the object has to go to the general kernel.
__device__ __forceinline__ bool stat_object(WarpSmem& w, const long long* __restrict__ seg_off, const double* __restrict__ gt,
                                            const double* __restrict__ gf, const double* __restrict__ ge,
                                            const unsigned short* __restrict__ gtrank, long long o, int lane,
                                            double* __restrict__ out) {
  // ---- offsets
  long long myoff = 0;
  if (lane < 8) myoff = seg_off[o * N_SEG + lane];
  const long long base = __shfl_sync(FULL, myoff, 0);
  int off[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) off[b] = (int)(__shfl_sync(FULL, myoff, b) - base);
  const int n = off[7];
  if (n > CAP - 1) return false;
  double* dst = out + o * STAT_COLS;
  if (n == 0) {                                          // statistical.py:56-66 for every segment, no ratios, no peak band
    for (int c = lane; c < STAT_COLS; c += 32)
      dst[c] = (c < 7 * N_STAT) ? ((c % N_STAT == 0) ? 0.0 : qnan()) : ((c == STAT_COLS - 1) ? -1.0 : qnan());
    return true;
  }
  if (lane < 8) w.soff[lane] = off[lane];
  const int iters = (n + 31) >> 5;
  float w_fmn, w_fmx; bool w_bad;
  // ---- flat phase 1: stage, |f|/err, range of the fluxes.  All global loads of the object are issued before the first
  // use (one DRAM latency per object instead of one per 32 observations)
  double fr[8]; int trk[8];
  {
    double tt[8], ee[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      fr[it] = 0.0; trk[it] = 0; tt[it] = 0.0; ee[it] = 0.0;
      const int e = it * 32 + lane;
      if (e < n) { tt[it] = gt[base + e]; fr[it] = gf[base + e]; ee[it] = ge[base + e]; trk[it] = gtrank[base + e]; }
    }
    float fmn = __int_as_float(0x7f800000), fmx = -fmn;
    bool bad = false;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (it < iters) {
        const int e = it * 32 + lane;
        if (e < n) {
          const double t = tt[it], f = fr[it], er = ee[it];
          w.st[e] = t; w.sf[e] = f; w.inv[trk[it]] = (unsigned char)e;
          w.sa[e] = (er > 0.0) ? fabs(f) * frcp(er) : -1.0;              // :116-120
          bad = bad || !(fabs(f) < pinf()) || !(fabs(t) < pinf());
          const float ff = (float)f;
          fmn = fminf(fmn, ff); fmx = fmaxf(fmx, ff);
        }
      }
    }
    w_fmn = fmn; w_fmx = fmx; w_bad = bad;
  }
  const float lo_f = funkey(__reduce_min_sync(FULL, fkey(w_fmn))), hi_f = funkey(__reduce_max_sync(FULL, fkey(w_fmx)));
  // bounds of the fluxes good to float precision, pushed outwards by two float ulps: lo <= every f <= hi
  const double lo = (double)lo_f - 2.4e-7 * fabs((double)lo_f) - 1e-300, hi = (double)hi_f + 2.4e-7 * fabs((double)hi_f) + 1e-300;
  const double range = hi - lo;
  if (__any_sync(FULL, w_bad || !(range < pinf()))) return false;
  const double scale = 16777215.0 * frcp(range) * (1.0 - 1e-9);      // (f - lo) * scale < 2^24
  __syncwarp();
  // ---- flat phase 2: band of every observation, slope terms, sort keys
  unsigned key[8];
  double amax = 0.0;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    key[it] = 0xffffffffu;
    if (it < iters) {
      const int e = it * 32 + lane;
      if (e < n) {
        int b = 0;
#pragma unroll
        for (int k = 1; k < 7; ++k) b += (e >= off[k]);
        w.sband[e] = (unsigned char)b;
        const double t = w.st[e], f = fr[it];
        double sl = 0.0;
        if (e + 1 < w.soff[b + 1]) {                                   // next observation of the band (time order), :99-113
          const double dt = w.st[e + 1] - t;
          if (dt > 0.0) sl = fabs((w.sf[e + 1] - f) * frcp(dt));
        }
        w.sb[e] = sl;
        if (trk[it] + 1 < n) {                                         // next observation of the object (all rows, :186-190)
          const int s = w.inv[trk[it] + 1];
          const double dt = w.st[s] - t;
          if (dt > 0.0) { const double x = fabs((w.sf[s] - f) * frcp(dt)); amax = (x > amax) ? x : amax; }
        }
        const unsigned q = (unsigned)((f - lo) * scale);
        key[it] = (q << 8) | (unsigned)e;
      }
    }
  }
  // all-band maximum slope: a maximum of non-negative values, reduced as float bit patterns (1 REDUX); the result is one
  // of the lanes' values rounded to float (6e-8 relative, tolerance of the pass 1e-5) unless that lane's double is re-read
  const unsigned amax_bits = __reduce_max_sync(FULL, __float_as_uint((float)amax));
  double slope_all = 0.0;
  if (amax_bits != 0u) {                                              // recover the exact double: the lane(s) holding the maximum
    const unsigned who = __ballot_sync(FULL, __float_as_uint((float)amax) == amax_bits);
    double cand = (__float_as_uint((float)amax) == amax_bits) ? amax : 0.0;
    // several doubles may round to the same float: take the largest of them over the (few) candidate lanes
    unsigned rest = who;
    double best = 0.0;
    while (rest) { const int src = __ffs(rest) - 1; rest &= rest - 1; best = fmax(best, __shfl_sync(FULL, cand, src)); }
    slope_all = best;
  }
  const double t_first = w.st[w.inv[0]], t_last = w.st[w.inv[n - 1]];
  __syncwarp();
  // ---- group phase: lanes 4g .. 4g+3 hold segment g (0..5 bands, 6 unknown filter); group 7 carries the all-band block
  const int gq = lane >> 2, p = lane & 3;
  const int glo = (gq < 7) ? w.soff[gq] : 0;
  const int nb = (gq < 7) ? w.soff[gq + 1] - glo : 0;
  const double tb0 = (nb > 0) ? w.st[glo] : 0.0, tb1 = (nb > 0) ? w.st[glo + nb - 1] : 0.0;      // the band is time-sorted
  double s1 = 0.0, snr = 0.0, slope = 0.0; int nsnr = 0;
  for (int i = p; i < nb; i += 4) {
    s1 += w.sf[glo + i];
    const double r = w.sa[glo + i];
    if (r >= 0.0) { snr += r; ++nsnr; }
    { const double x = w.sb[glo + i]; slope = (x > slope) ? x : slope; }
  }
#pragma unroll
  for (int x = 1; x <= 2; x <<= 1) {
    s1 += __shfl_xor_sync(FULL, s1, x); snr += __shfl_xor_sync(FULL, snr, x);
    { const double y = __shfl_xor_sync(FULL, slope, x); slope = (y > slope) ? y : slope; } nsnr += __shfl_xor_sync(FULL, nsnr, x);
  }
  double S1 = s1, SNR = snr; int NSNR = nsnr;            // totals over the groups (group 7 contributes zeros)
#pragma unroll
  for (int x = 4; x <= 16; x <<= 1) { S1 += __shfl_xor_sync(FULL, S1, x); SNR += __shfl_xor_sync(FULL, SNR, x); NSNR += __shfl_xor_sync(FULL, NSNR, x); }
  const double dnb = (double)nb, dn = (double)n;
  const double rnb = (nb > 0) ? frcp(dnb) : 0.0, rn = frcp(dn);
  // mean and standard deviation with IEEE division and square root: they feed the strict comparisons of :90-96, where one
  // ulp flips a count (n = 2: both observations are exactly one standard deviation from the mean)
  const double mean_b = (nb > 0) ? s1 / dnb : 0.0, mean_all = S1 / dn;
  double m2 = 0.0, m3 = 0.0, m4 = 0.0;
  for (int i = p; i < nb; i += 4) { const double d = w.sf[glo + i] - mean_b, d2 = d * d; m2 += d2; m3 += d2 * d; m4 += d2 * d2; }
#pragma unroll
  for (int x = 1; x <= 2; x <<= 1) { m2 += __shfl_xor_sync(FULL, m2, x); m3 += __shfl_xor_sync(FULL, m3, x); m4 += __shfl_xor_sync(FULL, m4, x); }
  // all-band central moments from the band moments: shift every band to the common mean (exact binomial identities)
  double M2, M3, M4;
  {
    const double dl = mean_b - mean_all, dl2 = dl * dl;
    double c2 = 0.0, c3 = 0.0, c4 = 0.0;
    if (nb > 0 && p == 0) {
      c2 = m2 + dnb * dl2;
      c3 = m3 + 3.0 * dl * m2 + dnb * dl2 * dl;
      c4 = m4 + 4.0 * dl * m3 + 6.0 * dl2 * m2 + dnb * dl2 * dl2;
    }
#pragma unroll
    for (int x = 4; x <= 16; x <<= 1) { c2 += __shfl_xor_sync(FULL, c2, x); c3 += __shfl_xor_sync(FULL, c3, x); c4 += __shfl_xor_sync(FULL, c4, x); }
    M2 = __shfl_sync(FULL, c2, 0); M3 = __shfl_sync(FULL, c3, 0); M4 = __shfl_sync(FULL, c4, 0);
  }
  const double sd_b = (nb > 1) ? sqrt(m2 / dnb) : 0.0;          // :71 (ddof = 0)
  const double sd_all = (n > 1) ? sqrt(M2 / dn) : 0.0;
  int b1 = 0, b2 = 0;
  if (sd_b > 0.0)
    for (int i = p; i < nb; i += 4) { const double ad = fabs(w.sf[glo + i] - mean_b); b1 += (ad > sd_b); b2 += (ad > 2.0 * sd_b); }
  b1 += __shfl_xor_sync(FULL, b1, 1); b2 += __shfl_xor_sync(FULL, b2, 1);
  b1 += __shfl_xor_sync(FULL, b1, 2); b2 += __shfl_xor_sync(FULL, b2, 2);
  int B1 = 0, B2 = 0;                                     // all rows: flat over the registers
  if (sd_all > 0.0) {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      if (it < iters && it * 32 + lane < n) { const double ad = fabs(fr[it] - mean_all); B1 += (ad > sd_all); B2 += (ad > 2.0 * sd_all); }
  }
  B1 = __reduce_add_sync(FULL, B1); B2 = __reduce_add_sync(FULL, B2);
  // ---- the leaders write the columns that need no order statistic (group 7: the all-band block)
  const bool all = (gq == 7);
  const int nn = all ? n : nb;
  const bool leader = (p == 0) && (gq != 6);
  static_assert(ROW <= CAP && STAT_COLS <= ROW, "the row is assembled in the storage of the staged times");
  double* const row = w.st;
  double* rw = row + (all ? 6 : gq) * N_STAT;
  __syncwarp();                                            // every lane has read its times: the row takes their place
  if (leader) {
    if (nn == 0) {
      rw[0] = 0.0;
#pragma unroll
      for (int k = 1; k < N_STAT; ++k) rw[k] = qnan();
    } else {
      const double dnn = all ? dn : dnb, rnn = all ? rn : rnb;
      const double mean = all ? mean_all : mean_b, sd = all ? sd_all : sd_b;
      const double mm3 = all ? M3 : m3, mm4 = all ? M4 : m4;
      const int c1 = all ? B1 : b1, c2 = all ? B2 : b2, ns = all ? NSNR : nsnr;
      const double tf0 = all ? t_first : tb0, tf1 = all ? t_last : tb1;
      double skew = 0.0, kurt = 0.0;
      if (sd > 0.0) {
        const double isd = frcp(sd), isd2 = isd * isd;
        if (nn > 2) skew = (mm3 * rnn) * (isd2 * isd);
        if (nn > 3) kurt = (mm4 * rnn) * (isd2 * isd2) - 3.0;
      }
      rw[0] = dnn; rw[1] = mean; rw[2] = sd; rw[6] = skew; rw[7] = kurt;
      rw[11] = (sd > 0.0) ? (double)c1 * rnn : 0.0;
      rw[12] = (sd > 0.0) ? (double)c2 * rnn : 0.0;
      rw[13] = all ? slope_all : slope;
      rw[14] = (ns > 0) ? (all ? SNR : snr) * frcp((double)ns) : qnan();
      const double span = (nn > 1) ? tf1 - tf0 : 0.0;
      rw[15] = span;
      rw[16] = (nn > 1) ? span * frcp((double)(nn - 1)) : 0.0;       // mean of the consecutive differences (telescoped)
    }
  }
  __syncwarp();                                            // sa / sb are free from here on
  // ---- one sort for the object
  sort256(key, lane);
  double g[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) g[a] = (lane * 8 + a < n) ? w.sf[key[a] & 255u] : pinf();
  for (;;) {                                               // repair 24-bit ties: odd-even transposition on the exact values
    bool inv = false;
#pragma unroll
    for (int a = 0; a < 7; ++a) inv = inv || (g[a] > g[a + 1]);
    const double nx = __shfl_down_sync(FULL, g[0], 1);
    inv = inv || (lane < 31 && g[7] > nx);
    if (!__any_sync(FULL, inv)) break;
#pragma unroll
    for (int a = 0; a < 8; a += 2)
      if (g[a] > g[a + 1]) { const double tg = g[a]; g[a] = g[a + 1]; g[a + 1] = tg; const unsigned tk = key[a]; key[a] = key[a + 1]; key[a + 1] = tk; }
#pragma unroll
    for (int a = 1; a < 7; a += 2)
      if (g[a] > g[a + 1]) { const double tg = g[a]; g[a] = g[a + 1]; g[a + 1] = tg; const unsigned tk = key[a]; key[a] = key[a + 1]; key[a + 1] = tk; }
    const double g0n = __shfl_down_sync(FULL, g[0], 1), g7p = __shfl_up_sync(FULL, g[7], 1);
    const unsigned k0n = __shfl_down_sync(FULL, key[0], 1), k7p = __shfl_up_sync(FULL, key[7], 1);
    const bool sw_hi = lane < 31 && g[7] > g0n, sw_lo = lane > 0 && g7p > g[0];
    if (sw_hi) { g[7] = g0n; key[7] = k0n; }
    if (sw_lo) { g[0] = g7p; key[0] = k7p; }
  }
  // merged run (skewed index E + E/8 = lane * 9 + a) and the stable partition into the band runs
  unsigned long long cnt = 0ULL;
  int lr[8], bnd[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    lr[a] = 0; bnd[a] = 0;
    if (lane * 8 + a < n) {
      w.sa[lane * 9 + a] = g[a];
      const int b = w.sband[key[a] & 255u];
      bnd[a] = b;
      lr[a] = (int)((cnt >> (8 * b)) & 255ULL);
      cnt += 1ULL << (8 * b);
    }
  }
  unsigned long long pre = cnt;                            // inclusive scan over the lanes (fields cannot overflow: n <= 255)
#pragma unroll
  for (int x = 1; x < 32; x <<= 1) { const unsigned long long up = __shfl_up_sync(FULL, pre, x); if (lane >= x) pre += up; }
  pre -= cnt;
#pragma unroll
  for (int a = 0; a < 8; ++a)
    if (lane * 8 + a < n) w.sb[w.soff[bnd[a]] + (int)((pre >> (8 * bnd[a])) & 255ULL) + lr[a]] = g[a];
  __syncwarp();
  // ---- order statistics, eight segments at a time
  if (leader && nn > 0) {
    const double* s = all ? w.sa : w.sb + glo;
    const int skew_mask = all ? ~0 : 0;
    auto S = [&](int i) { return s[i + ((i >> 3) & skew_mask)]; };
    const double mn = S(0), mx = S(nn - 1);
    const int h = nn >> 1;
    const double med = (nn & 1) ? S(h) : 0.5 * (S(h - 1) + S(h));               // numpy.median
    double iqr = 0.0;
    if (nn > 1) {                                                            // numpy.percentile, linear (_lerp)
      double q[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double pos = (k ? 0.75 : 0.25) * (double)(nn - 1);
        const int l = (int)pos;
        if (l >= nn - 1) { q[k] = mx; continue; }
        const double gg = pos - (double)l, a = S(l), b = S(l + 1), d = b - a;
        q[k] = (gg >= 0.5) ? (b - d * (1.0 - gg)) : (a + d * gg);
      }
      iqr = q[1] - q[0];
    }
    // MAD: the h+1 values closest to the median are a contiguous window of the run; smallest start c at which the window
    // need not move right (monotone predicate -> binary search)
    const int wdt = h + 1;
    int c0 = 0, c1 = nn - wdt;
    while (c0 < c1) {
      const int mid = (c0 + c1) >> 1;
      if (med - S(mid) > S(mid + wdt) - med) c0 = mid + 1; else c1 = mid;       // mid + wdt <= nn - 1 here
    }
    const double dl = med - S(c0), dr = S(c0 + wdt - 1) - med;
    const double dh = fmax(dl, dr);                          // order statistic h of the deviations
    double mad = dh;
    if (!(nn & 1)) {
      double dprev;                                          // order statistic h-1: drop the farther end
      if (wdt == 1) dprev = dh;
      else if (dl >= dr) dprev = fmax(med - S(c0 + 1), dr);
      else dprev = fmax(dl, S(c0 + wdt - 2) - med);
      mad = 0.5 * (dprev + dh);
    }
    rw[3] = mn; rw[4] = mx; rw[5] = med; rw[8] = mx - mn; rw[9] = mad; rw[10] = iqr;
  }
  __syncwarp();
  if (lane == 0) {                                          // cross-band tail (:201-222)
    const int k = 7 * N_STAT;
    const int pa[3] = {BG, BR, BI}, pb[3] = {BR, BI, BZ};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double a = row[pa[j] * N_STAT + 1], c = row[pb[j] * N_STAT + 1];
      row[k + j] = (!is_nan(a) && c > 0.0) ? a / c : qnan();
    }
    int pk = -1; double best = 0.0;
#pragma unroll
    for (int b = 0; b < N_BAND; ++b) {
      const double m = row[b * N_STAT + 4];
      if (!is_nan(m) && (pk < 0 || m > best)) { best = m; pk = b; }
    }
    row[k + 3] = (double)pk;
  }
  __syncwarp();
  for (int c = lane; c < STAT_COLS; c += 32) dst[c] = row[c];
  __syncwarp();
  return true;
}

}  // namespace statw
}  // namespace mlrn
