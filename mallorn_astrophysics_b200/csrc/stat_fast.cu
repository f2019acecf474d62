// stat_fast.cu -- K2, the moment pass (123 columns, schema.STAT_COLUMNS), reference
// src/features/statistical.py:41-226.  One warp per object.
//
// HBM traffic is the algorithmic minimum: every observation (t, flux, err float64 + the uint16 time rank)
// is read once, coalesced, into the warp's shared-memory slice, and the 123-value row is written once.
// Everything else stays on chip:
//   * per band (time-sorted in the store): sums by xor-butterfly shuffles, two-pass central moments,
//     max |df/dt| over consecutive observations, the flux sort as a bitonic network over REGISTERS
//     (one or two keys per lane, compare-exchange by shuffle), quartiles by direct indexing, and the MAD
//     without a second sort: the h+1 values closest to the median form a contiguous window of the sorted
//     run, found by one ballot search;
//   * the all-band segment re-uses the seven sorted runs (k-way merge by binary-search ranks), takes its
//     time order from the stored time rank (no sort), and its sums from one pass over the staged fluxes.
// Bands longer than 64 points fall back to a shared-memory bitonic sort (same results).
#include <cuda_runtime.h>
#include "store.cuh"
#include "stat.cuh"
#include "stat_warp.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SF_WARPS_MAX = 8;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// ascending bitonic sort of NR*32 keys, element e = a*32 + lane lives in v[a] of `lane`.  The (k, j) loops stay
// rolled on purpose: fully unrolled, the network of the two instantiations pushed the kernel past the
// instruction cache (ncu: 33 % of stall samples were "no instruction").
template <int NR>
__device__ __forceinline__ void reg_bitonic_sort(double (&v)[NR], int lane) {
#pragma unroll 1
  for (int k = 2; k <= 32 * NR; k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {                       // only for NR == 2: exchange between the two registers of a lane
        const bool up = (lane & k) == 0;   // element index of v[0] is `lane` (< 32), k is 64 here: always ascending
        const double lo = fmin(v[0], v[NR - 1]), hi = fmax(v[0], v[NR - 1]);
        v[0] = up ? lo : hi; v[NR - 1] = up ? hi : lo;
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int a = 0; a < NR; ++a) {
          const double other = __shfl_xor_sync(FULL, v[a], j);
          const bool up = (((a << 5) | lane) & k) == 0;
          const bool take_other = (lower == up) ? (other < v[a]) : (other > v[a]);
          v[a] = take_other ? other : v[a];
        }
      }
    }
  }
}

// median of |s - med| over an ascending run s[0..n): the h+1 elements closest to med are a contiguous window
__device__ double mad_of_sorted(const double* s, int n, double med, int lane) {
  const int h = n >> 1, w = h + 1;
  // smallest lo in [0, n-w] for which element lo is NOT farther from med than element lo+w
  int lo = n - w;
  for (int base = 0; base <= n - w; base += 32) {
    const int c = base + lane;
    bool shift = false;                                   // true: window must move right past c
    if (c <= n - w) shift = (c + w < n) && (med - s[c] > s[c + w] - med);
    else shift = false;
    const unsigned keep = __ballot_sync(FULL, !shift);    // candidates at which the window may stay
    if (keep) { lo = base + __ffs(keep) - 1; break; }
  }
  if (lo > n - w) lo = n - w;
  const double dl = med - s[lo], dr = s[lo + w - 1] - med;
  const double dh = fmax(dl, dr);                         // order statistic h of the deviations
  if (n & 1) return dh;
  double dprev;                                           // order statistic h-1: drop the farther end
  if (w == 1) dprev = dh;
  else if (dl >= dr) dprev = fmax(med - s[lo + 1], dr);
  else dprev = fmax(dl, s[lo + w - 2] - med);
  return 0.5 * (dprev + dh);
}

// index of the first element of s[0..n) that is >= v (lower) / > v (upper); trip count depends on n only
__device__ __forceinline__ int bound(const double* s, int n, double v, bool upper) {
  int pos = 0;
  for (int step = 1 << (31 - __clz(n)); step > 0; step >>= 1) {
    const int idx = pos + step;
    if (idx <= n) { const double x = s[idx - 1]; if (upper ? (x <= v) : (x < v)) pos = idx; }
  }
  return pos;
}

struct Sums { double s1, snr; int nsnr; };

// One band, n >= 1, time-sorted (t, f, e) in shared memory.  Writes the sorted fluxes to run[0..n) and the 17
// statistics to out (lane 0).  srt_tmp: scratch >= next_pow2(n) for the n > 64 fallback.
__device__ void band_stats(const double* t, const double* f, const double* e, int n, double* run, double* srt_tmp,
                           double* out, Sums& acc, int lane) {
  const double dn = (double)n;
  // ---- sums
  double s1 = 0.0, snr = 0.0; int nsnr = 0;
  for (int i = lane; i < n; i += 32) {
    const double x = f[i], er = e[i];
    s1 += x;
    if (er > 0.0) { snr += fabs(x) / er; ++nsnr; }
  }
  s1 = wsum(s1); snr = wsum(snr); nsnr = __reduce_add_sync(FULL, nsnr);
  acc.s1 += s1; acc.snr += snr; acc.nsnr += nsnr;
  const double mean = s1 / dn;
  double m2 = 0.0, m3 = 0.0, m4 = 0.0, slope = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double d = f[i] - mean, d2 = d * d;
    m2 += d2; m3 += d2 * d; m4 += d2 * d2;
    if (i + 1 < n) {
      const double dt = t[i + 1] - t[i];
      if (dt > 0.0) slope = fmax(slope, fabs((f[i + 1] - f[i]) / dt));
    }
  }
  m2 = wsum(m2); m3 = wsum(m3); m4 = wsum(m4); slope = wmax(slope);
  const double sd = (n > 1) ? sqrt(m2 / dn) : 0.0;
  int b1 = 0, b2 = 0;
  if (sd > 0.0) {
    for (int i = lane; i < n; i += 32) {
      const double ad = fabs(f[i] - mean);
      b1 += (ad > sd); b2 += (ad > 2.0 * sd);
    }
    b1 = __reduce_add_sync(FULL, b1); b2 = __reduce_add_sync(FULL, b2);
  }
  // ---- sorted run
  if (n <= 32) {
    double v[1] = {lane < n ? f[lane] : pinf()};
    reg_bitonic_sort<1>(v, lane);
    if (lane < n) run[lane] = v[0];
  } else if (n <= 64) {
    double v[2] = {f[lane], (32 + lane < n) ? f[32 + lane] : pinf()};
    reg_bitonic_sort<2>(v, lane);
    run[lane] = v[0];
    if (32 + lane < n) run[32 + lane] = v[1];
  } else {
    WarpTeam tm;
    for (int i = lane; i < n; i += 32) srt_tmp[i] = f[i];
    __syncwarp();
    tm.sort(srt_tmp, n);
    for (int i = lane; i < n; i += 32) run[i] = srt_tmp[i];
  }
  __syncwarp();
  const double med = sorted_median(run, n);
  const double iqr = (n > 1) ? sorted_percentile(run, n, 75.0) - sorted_percentile(run, n, 25.0) : 0.0;
  const double mad = mad_of_sorted(run, n, med, lane);
  if (lane == 0) {
    const double sd2 = sd * sd;
    out[0] = dn; out[1] = mean; out[2] = sd; out[3] = run[0]; out[4] = run[n - 1]; out[5] = med;
    out[6] = (sd > 0.0 && n > 2) ? (m3 / dn) / (sd2 * sd) : 0.0;
    out[7] = (sd > 0.0 && n > 3) ? (m4 / dn) / (sd2 * sd2) - 3.0 : 0.0;
    out[8] = run[n - 1] - run[0]; out[9] = mad; out[10] = iqr;
    out[11] = (sd > 0.0) ? (double)b1 / dn : 0.0;
    out[12] = (sd > 0.0) ? (double)b2 / dn : 0.0;
    out[13] = slope;
    out[14] = (nsnr > 0) ? snr / (double)nsnr : qnan();
    out[15] = (n > 1) ? t[n - 1] - t[0] : 0.0;
    double cad = 0.0;
    if (n > 1) cad = (t[n - 1] - t[0]) / (double)(n - 1);      // mean of the consecutive differences (telescoped)
    out[16] = cad;
  }
}

constexpr int SF_ROW = 160;   // 123-value row + a dump slot for the unknown-filter segment's statistics
__host__ __device__ inline size_t sf_warp_doubles(int cap) { return 5 * (size_t)cap + SF_ROW; }

__global__ void __launch_bounds__(SF_WARPS_MAX * 32)
k_stat(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
       const double* __restrict__ ge, const unsigned short* __restrict__ gtrank, ObjClass cls, const int* __restrict__ list_count,
       int cap, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (list_count) cls.count = *list_count;            // a list made on the device by the launch before (k_stat_warp)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double* row = cls.scratch ? (double*)cls.scratch + ((size_t)blockIdx.x * nwarps + warp) * sf_warp_doubles(cap)
                            : (double*)smem + (size_t)warp * sf_warp_doubles(cap);
  double* st = row + SF_ROW;
  double* sf = st + cap;
  double* se = sf + cap;        // errors during the band phase, time-ordered times afterwards
  double* run = se + cap;       // seven sorted runs, back to back (same offsets as the segments)
  double* tf = run + cap;       // time-ordered fluxes; scratch for the sort of long bands before that
  double* mrg = st;             // merged sorted fluxes: the staged times are dead once the time-ordered copy exists
  for (long long slot = (long long)blockIdx.x * nwarps + warp; slot < cls.count; slot += (long long)gridDim.x * nwarps) {
    const long long o = cls.list ? cls.list[slot] : slot;
    const long long* so = seg_off + o * N_SEG;
    const long long base = so[0];
    int off[N_SEG + 1];
#pragma unroll
    for (int b = 0; b <= N_SEG; ++b) off[b] = (int)(so[b] - base);
    const int n = off[N_SEG];
    if (!cls.list && n > cls.n_hi) continue;          // a long object: the second launch's
    unsigned short trk[8];                                  // time ranks of this lane's observations (n <= 256 fast path)
    for (int i = lane, q = 0; i < n; i += 32, ++q) {
      st[i] = gt[base + i]; sf[i] = gf[base + i]; se[i] = ge[base + i];
      if (q < 8) trk[q] = gtrank[base + i];
    }
    __syncwarp();
    Sums acc{0.0, 0.0, 0};
    // ---- six bands (+ the unknown-filter segment, which only feeds the all-band block)
    for (int b = 0; b < N_SEG; ++b) {
      const int lo = off[b], nb = off[b + 1] - lo;
      if (nb == 0) {
        if (b < N_BAND) { if (lane == 0) row[b * N_STAT] = 0.0; if (lane >= 1 && lane < N_STAT) row[b * N_STAT + lane] = qnan(); }
        continue;
      }
      band_stats(st + lo, sf + lo, se + lo, nb, run + lo, tf, (b < N_BAND) ? row + b * N_STAT : row + 140, acc, lane);
      __syncwarp();
    }
    // ---- all rows (statistical.py:186-190)
    double* ra = row + 6 * N_STAT;
    if (n == 0) {
      if (lane == 0) ra[0] = 0.0;
      if (lane >= 1 && lane < N_STAT) ra[lane] = qnan();
    } else {
      const double dn = (double)n;
      // time order from the stored time rank; the error array is dead by now
      double* tt = se;
      if (n <= 256) {
        for (int i = lane, q = 0; i < n; i += 32, ++q) { tt[trk[q]] = st[i]; tf[trk[q]] = sf[i]; }
      } else {
        for (int i = lane; i < n; i += 32) { const int k = gtrank[base + i]; tt[k] = st[i]; tf[k] = sf[i]; }
      }
      __syncwarp();
      // merged sorted fluxes: rank of every run element among all runs (stable: earlier runs win ties)
      for (int i = lane; i < n; i += 32) {
        int q = 0;
        while (q < N_SEG - 1 && i >= off[q + 1]) ++q;
        const double v = run[i];
        int pos = i - off[q];
#pragma unroll
        for (int r = 0; r < N_SEG; ++r) {
          const int nr = off[r + 1] - off[r];
          if (r != q && nr > 0) pos += bound(run + off[r], nr, v, r < q);
        }
        mrg[pos] = v;
      }
      __syncwarp();
      const double mean = acc.s1 / dn;
      double m2 = 0.0, m3 = 0.0, m4 = 0.0, slope = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double d = sf[i] - mean, d2 = d * d;
        m2 += d2; m3 += d2 * d; m4 += d2 * d2;
        if (i + 1 < n) {
          const double dt = tt[i + 1] - tt[i];
          if (dt > 0.0) slope = fmax(slope, fabs((tf[i + 1] - tf[i]) / dt));
        }
      }
      m2 = wsum(m2); m3 = wsum(m3); m4 = wsum(m4); slope = wmax(slope);
      const double sd = (n > 1) ? sqrt(m2 / dn) : 0.0;
      int b1 = 0, b2 = 0;
      if (sd > 0.0) {
        for (int i = lane; i < n; i += 32) { const double ad = fabs(sf[i] - mean); b1 += (ad > sd); b2 += (ad > 2.0 * sd); }
        b1 = __reduce_add_sync(FULL, b1); b2 = __reduce_add_sync(FULL, b2);
      }
      const double med = sorted_median(mrg, n);
      const double iqr = (n > 1) ? sorted_percentile(mrg, n, 75.0) - sorted_percentile(mrg, n, 25.0) : 0.0;
      const double mad = mad_of_sorted(mrg, n, med, lane);
      if (lane == 0) {
        const double sd2 = sd * sd;
        ra[0] = dn; ra[1] = mean; ra[2] = sd; ra[3] = mrg[0]; ra[4] = mrg[n - 1]; ra[5] = med;
        ra[6] = (sd > 0.0 && n > 2) ? (m3 / dn) / (sd2 * sd) : 0.0;
        ra[7] = (sd > 0.0 && n > 3) ? (m4 / dn) / (sd2 * sd2) - 3.0 : 0.0;
        ra[8] = mrg[n - 1] - mrg[0]; ra[9] = mad; ra[10] = iqr;
        ra[11] = (sd > 0.0) ? (double)b1 / dn : 0.0;
        ra[12] = (sd > 0.0) ? (double)b2 / dn : 0.0;
        ra[13] = slope;
        ra[14] = (acc.nsnr > 0) ? acc.snr / (double)acc.nsnr : qnan();
        ra[15] = (n > 1) ? tt[n - 1] - tt[0] : 0.0;
        ra[16] = (n > 1) ? (tt[n - 1] - tt[0]) / (double)(n - 1) : 0.0;
      }
    }
    __syncwarp();
    if (lane == 0) {                                        // cross-band tail (:201-222)
      const int k = 7 * N_STAT;
      const int pa[3] = {BG, BR, BI}, pb[3] = {BR, BI, BZ};
      for (int j = 0; j < 3; ++j) {
        const double a = row[pa[j] * N_STAT + 1], c = row[pb[j] * N_STAT + 1];
        row[k + j] = (!is_nan(a) && c > 0.0) ? a / c : qnan();
      }
      int pk = -1; double best = 0.0;
      for (int b = 0; b < N_BAND; ++b) {
        const double m = row[b * N_STAT + 4];
        if (!is_nan(m) && (pk < 0 || m > best)) { best = m; pk = b; }
      }
      row[k + 3] = (double)pk;
    }
    __syncwarp();
    double* dst = out + o * STAT_COLS;
    for (int c = lane; c < STAT_COLS; c += 32) dst[c] = row[c];
    __syncwarp();
  }
}

// The ordinary objects (at most 255 observations): stat_warp.cuh.  What it declines -- longer short objects, non-finite
// values -- goes on `fb_list` for k_stat above; objects beyond SHORT_LIMIT are the long-object launch's.
constexpr int SW_WARPS = 4;
#ifndef SW_MIN_BLOCKS
#define SW_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(SW_WARPS * 32, SW_MIN_BLOCKS)
k_stat_warp(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
            const double* __restrict__ ge, const unsigned short* __restrict__ gtrank, int n_obj, double* __restrict__ out,
            int* __restrict__ fb_list, int* __restrict__ fb_count) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  statw::WarpSmem& w = reinterpret_cast<statw::WarpSmem*>(smem)[warp];
  for (long long o = (long long)blockIdx.x * SW_WARPS + warp; o < n_obj; o += (long long)gridDim.x * SW_WARPS) {
    if (!statw::stat_object(w, seg_off, gt, gf, ge, gtrank, o, lane, out)) {
      if (lane == 0 && (int)(seg_off[(o + 1) * N_SEG] - seg_off[o * N_SEG]) <= SHORT_LIMIT) fb_list[atomicAdd(fb_count, 1)] = (int)o;
    }
    __syncwarp();
  }
}

}  // namespace

// staging capacity per warp for objects of at most `longest` observations with segments of at most `seg`: the object, and
// room for the power-of-two padding of the long-band sort
static int stat_cap(int longest, int seg) {
  int cap = (longest + 7) & ~7;
  if (cap < 32) cap = 32;
  if (seg > 64 && cap < next_pow2(seg)) cap = next_pow2(seg);
  return cap;
}

int run_stat(const mlrn_store* s, double* out_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  // ---- launch 1: every object of at most 255 observations (stat_warp.cuh); the others up to SHORT_LIMIT are listed
  int* fb = nullptr;                                  // [0] = count, [1..] = the listed objects
  MLRN_CUDA(cudaMallocAsync(&fb, sizeof(int) * ((size_t)s->n_obj + 1), st));
  auto fail = [&](cudaError_t e, const char* what) { cudaFreeAsync(fb, st); return cuda_fail(e, what); };
  cudaError_t e = cudaMemsetAsync(fb, 0, sizeof(int), st);
  if (e != cudaSuccess) return fail(e, "cudaMemsetAsync (k_stat_warp list)");
  {
    const size_t smem = sizeof(statw::WarpSmem) * SW_WARPS;
    e = cudaFuncSetAttribute(k_stat_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute (k_stat_warp)");
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stat_warp, SW_WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long want = ((long long)s->n_obj + SW_WARPS - 1) / SW_WARPS, capg = (long long)s->sm_count * per_sm;
    k_stat_warp<<<(int)(want < capg ? want : capg), SW_WARPS * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->err, s->trank, s->n_obj, out_dev, fb + 1, fb);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "k_stat_warp launch");
    count_launch();
  }
  // ---- launch 2: the listed objects with the general kernel, shared-memory staging sized by the longest short object (not
  // by the longest object of the batch: one 3,000-point object used to push every warp's slice to the point where a single
  // warp fitted on an SM).  The list is normally empty (the count is read on the device): a small grid then.
  {
    const int cap = stat_cap(s->max_short, s->max_seg < s->max_short ? s->max_seg : s->max_short);
    int warps = 4;
    while (warps > 1 && sf_warp_doubles(cap) * sizeof(double) * warps > 56 * 1024) warps >>= 1;
    const size_t smem = sf_warp_doubles(cap) * sizeof(double) * warps;
    e = cudaFuncSetAttribute(k_stat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute (k_stat)");
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stat, warps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long want = ((long long)s->n_obj + warps - 1) / warps;
    long long capg = (long long)s->sm_count * (s->max_short >= statw::CAP ? per_sm : 1);
    const int grid = (int)(want < capg ? want : capg);
    k_stat<<<grid, warps * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->err, s->trank, ObjClass{fb + 1, 0, SHORT_LIMIT, nullptr}, fb, cap, out_dev);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "k_stat launch");
    count_launch();
  }
  cudaFreeAsync(fb, st);
  if (s->n_long > 0) {                                // long objects: staging in a per-warp global scratch
    const int capl = stat_cap(s->max_obs, s->max_seg);
    int gl = (s->n_long + 3) / 4;
    if (gl > s->sm_count * 2) gl = s->sm_count * 2;
    unsigned char* scratch = nullptr;
    MLRN_CUDA(cudaMallocAsync(&scratch, sf_warp_doubles(capl) * sizeof(double) * 4 * (size_t)gl, st));
    k_stat<<<gl, 4 * 32, 0, st>>>(s->seg_off, s->t, s->flux, s->err, s->trank, ObjClass{s->long_list, s->n_long, MAX_OBS, scratch}, nullptr, capl, out_dev);
    e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return cuda_fail(e, "k_stat (long objects) launch");
    count_launch();
  }
  return MLRN_OK;
}

}  // namespace mlrn
