// team.cuh -- execution "teams": the cooperative-thread abstraction every feature routine is
// written against.
//
// A feature routine is a template over Team and uses only
//     tm.rank(), Team::size        lane-strided loops: for (i = tm.rank(); i < n; i += Team::size)
//     tm.sum/min/max/...           all-reduce, result identical on every member (uniform control flow)
//     tm.sync()                    make shared-scratch writes visible to the team
//     tm.sort / sort_by_key        in-scratch ascending sort
// On the B200 the team is one warp (WarpTeam: shuffles, __syncwarp, bitonic network in shared
// memory) or one CTA (BlockTeam, gp.cuh).  The same source also instantiates with a
// single-member HostTeam when compiled by g++ for tests/hostsim -- a *debugging aid* to exercise
// the arithmetic without a GPU; it is not part of the shipped library (libmallorn_b200.so is
// built by nvcc only and has no host compute path).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MLRN_FN __device__ inline
#define MLRN_FN_FORCE __device__ __forceinline__
#define MLRN_HD __host__ __device__ inline
#define MLRN_ROLL _Pragma("unroll 1")     // keep the NP-sized loops of the solver rolled (code size)
#define MLRN_NOINLINE __device__ __noinline__     // one copy per kernel: keeps the thread-per-fit kernels inside the i-cache
// Team helpers (reductions, sorts, pairings): inlined by default; a translation unit that defines MLRN_SHARED_HELPERS
// before including this header gets ONE copy of each per kernel instead.  passes.cu does: the K3 rows call them 40-75 times
// each, and inlined the kernels were 6-13 k instructions (100-213 KB of code against a 128 KB instruction cache).
#ifdef MLRN_SHARED_HELPERS
#define MLRN_TEAM_FN __device__ __noinline__
#define MLRN_TEAM_MEMBER __device__ __noinline__
#else
#define MLRN_TEAM_FN __device__ inline
#define MLRN_TEAM_MEMBER __device__
#endif
#else
#define MLRN_TEAM_FN inline
#define MLRN_TEAM_MEMBER
#define MLRN_FN inline
#define MLRN_FN_FORCE inline
#define MLRN_HD inline
#define MLRN_NOINLINE inline
#define MLRN_ROLL
#include <algorithm>
#include <limits>
#include <utility>
#include <vector>
#endif

namespace mlrn {

MLRN_FN_FORCE double qnan() {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return __builtin_nan("");
#endif
}
MLRN_FN_FORCE double pinf() {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(0x7ff0000000000000LL);
#else
  return __builtin_inf();
#endif
}
MLRN_FN_FORCE bool is_nan(double x) { return x != x; }
// orderings in which a NaN is the extreme value, as numpy.argmax / argmin treat it (the first NaN wins)
MLRN_FN_FORCE bool nan_gt(double a, double b) { return a > b || (a != a && b == b); }
MLRN_FN_FORCE bool nan_lt(double a, double b) { return a < b || (a != a && b == b); }
MLRN_FN_FORCE bool nan_eq(double a, double b) { return a == b || (a != a && b != b); }
MLRN_FN_FORCE bool is_finite(double x) { return fabs(x) < pinf(); }
#ifndef __CUDACC__
inline double rsqrt(double x) { return 1.0 / sqrt(x); }   // host instantiation only (tests/hostsim)
#endif
MLRN_HD int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// WarpTeam: 32 lanes, reductions by xor-butterfly (bit-identical result on every lane).
// ----------------------------------------------------------------------------------------------
struct WarpTeam {
  static constexpr int size = 32;
  static constexpr unsigned FULL = 0xffffffffu;
  int lane_;
  __device__ WarpTeam() : lane_(threadIdx.x & 31) {}
  __device__ __forceinline__ int rank() const { return lane_; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }

  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }
  // two independent sums through one butterfly: the shuffles of the second chain issue under the latency of the first
  // (each chain adds in exactly the order of sum(): bit-identical to two calls)
  __device__ __forceinline__ void sum2(double& a, double& b) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oa = __shfl_xor_sync(FULL, a, o), ob = __shfl_xor_sync(FULL, b, o);
      a += oa; b += ob;
    }
  }
  __device__ __forceinline__ int isum(int v) const { return __reduce_add_sync(FULL, v); }
  __device__ __forceinline__ int imin(int v) const { return __reduce_min_sync(FULL, v); }
  __device__ __forceinline__ int imax(int v) const { return __reduce_max_sync(FULL, v); }
  __device__ __forceinline__ double max(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
  }
  __device__ __forceinline__ double min(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
  }
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(FULL, p) != 0; }
  __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(FULL, v, src); }
  __device__ __forceinline__ int ibcast(int v, int src) const { return __shfl_sync(FULL, v, src); }

  // index of the largest value; ties -> smallest index; idx < 0 marks "no candidate".  A NaN counts as larger than every
  // number and equal to another NaN (numpy.argmax returns the first NaN): the relation is a total order, so every lane ends
  // with the SAME index -- with plain comparisons a NaN made the lanes disagree, and the team-collective code that follows
  // (compactions, sorts) ran with lane-dependent sizes.
  __device__ __forceinline__ int argmax_first(double v, int idx) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(FULL, v, o);
      int oi = __shfl_xor_sync(FULL, idx, o);
      bool take = (oi >= 0) && (idx < 0 || nan_gt(ov, v) || (nan_eq(ov, v) && oi < idx));
      if (take) { v = ov; idx = oi; }
    }
    return idx;
  }
  // index of the smallest value; ties -> smallest index (a NaN counts as smaller than every number: numpy.argmin)
  __device__ __forceinline__ int argmin_first(double v, int idx) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(FULL, v, o);
      int oi = __shfl_xor_sync(FULL, idx, o);
      bool take = (oi >= 0) && (idx < 0 || nan_lt(ov, v) || (nan_eq(ov, v) && oi < idx));
      if (take) { v = ov; idx = oi; }
    }
    return idx;
  }
  // Order-preserving compaction over a lane-strided loop: call once per trip with this lane's
  // predicate; returns the slot of this lane's element (valid only if keep) and advances count.
  __device__ __forceinline__ int compact_slot(bool keep, int& count) const {
    unsigned m = __ballot_sync(FULL, keep);
    int slot = count + __popc(m & ((1u << lane_) - 1u));
    count += __popc(m);
    return slot;
  }

  // Ascending bitonic sort of a[0..n) in shared scratch; capacity must be >= next_pow2(n).
  MLRN_TEAM_MEMBER void sort(double* a, int n) const {
    int np = next_pow2(n);
    for (int i = n + lane_; i < np; i += 32) a[i] = pinf();
    __syncwarp();
    for (int k = 2; k <= np; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane_; i < np; i += 32) {
          int p = i ^ j;
          if (p > i) {
            double x = a[i], y = a[p];
            bool up = ((i & k) == 0);
            if ((x > y) == up) { a[i] = y; a[p] = x; }
          }
        }
        __syncwarp();
      }
    }
  }
  // Ascending sort by key with a payload; ties broken by payload order is NOT guaranteed.
  MLRN_TEAM_MEMBER void sort_by_key(double* key, double* val, int n) const {
    int np = next_pow2(n);
    for (int i = n + lane_; i < np; i += 32) { key[i] = pinf(); val[i] = 0.0; }
    __syncwarp();
    for (int k = 2; k <= np; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane_; i < np; i += 32) {
          int p = i ^ j;
          if (p > i) {
            double x = key[i], y = key[p];
            bool up = ((i & k) == 0);
            if ((x > y) == up) {
              key[i] = y; key[p] = x;
              double vx = val[i]; val[i] = val[p]; val[p] = vx;
            }
          }
        }
        __syncwarp();
      }
    }
  }
};
#else
// ----------------------------------------------------------------------------------------------
// HostTeam: one member.  tests/hostsim only.
// ----------------------------------------------------------------------------------------------
struct HostTeam {
  static constexpr int size = 1;
  int rank() const { return 0; }
  void sync() const {}
  double sum(double v) const { return v; }
  void sum2(double&, double&) const {}
  int isum(int v) const { return v; }
  int imin(int v) const { return v; }
  int imax(int v) const { return v; }
  double max(double v) const { return v; }
  double min(double v) const { return v; }
  bool any(bool p) const { return p; }
  double bcast(double v, int) const { return v; }
  int ibcast(int v, int) const { return v; }
  int argmax_first(double, int idx) const { return idx; }
  int argmin_first(double, int idx) const { return idx; }
  int compact_slot(bool keep, int& count) const { int s = count; count += keep ? 1 : 0; return s; }
  void sort(double* a, int n) const { std::sort(a, a + n); }
  void sort_by_key(double* key, double* val, int n) const {
    std::vector<std::pair<double, double>> v(n);
    for (int i = 0; i < n; ++i) v[i] = {key[i], val[i]};
    std::stable_sort(v.begin(), v.end(), [](const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first < b.first; });
    for (int i = 0; i < n; ++i) { key[i] = v[i].first; val[i] = v[i].second; }
  }
};
#endif

// ----------------------------------------------------------------------------------------------
// Small team-wide reductions over scratch arrays (all members get the same answer).
// ----------------------------------------------------------------------------------------------
template <class Team>
MLRN_TEAM_FN double team_sum(const Team& tm, const double* a, int n) {
  double s = 0.0;
  for (int i = tm.rank(); i < n; i += Team::size) s += a[i];
  return tm.sum(s);
}
template <class Team>
MLRN_FN double team_mean(const Team& tm, const double* a, int n) { return team_sum(tm, a, n) / (double)n; }

// population variance, two-pass (numpy.var / numpy.std, ddof = 0)
template <class Team>
MLRN_TEAM_FN double team_var(const Team& tm, const double* a, int n, double mean) {
  double s = 0.0;
  for (int i = tm.rank(); i < n; i += Team::size) { double d = a[i] - mean; s += d * d; }
  return tm.sum(s) / (double)n;
}
template <class Team>
MLRN_TEAM_FN double team_max(const Team& tm, const double* a, int n) {
  double m = -pinf();
  for (int i = tm.rank(); i < n; i += Team::size) m = fmax(m, a[i]);
  return tm.max(m);
}
template <class Team>
MLRN_TEAM_FN double team_min(const Team& tm, const double* a, int n) {
  double m = pinf();
  for (int i = tm.rank(); i < n; i += Team::size) m = fmin(m, a[i]);
  return tm.min(m);
}
// first index of the maximum of a[0..n) (numpy.argmax); -1 if n == 0
template <class Team>
MLRN_TEAM_FN int team_argmax_first(const Team& tm, const double* a, int n) {
  double bv = 0.0; int bi = -1;
  for (int i = tm.rank(); i < n; i += Team::size) {
    double v = a[i];
    if (bi < 0 || nan_gt(v, bv)) { bv = v; bi = i; }
  }
  return tm.argmax_first(bv, bi);
}
// first index of the minimum of a[0..n) (numpy.argmin); -1 if n == 0
template <class Team>
MLRN_TEAM_FN int team_argmin_first(const Team& tm, const double* a, int n) {
  double bv = 0.0; int bi = -1;
  for (int i = tm.rank(); i < n; i += Team::size) {
    double v = a[i];
    if (bi < 0 || nan_lt(v, bv)) { bv = v; bi = i; }
  }
  return tm.argmin_first(bv, bi);
}

// log10 through one shared copy where helpers are shared (13 call sites in the TDE row alone)
MLRN_TEAM_FN double team_log10(double x) { return log10(x); }

// numpy.percentile(method='linear') on an ascending-sorted array
MLRN_FN double sorted_percentile(const double* s, int n, double q) {
  double pos = (q / 100.0) * (double)(n - 1);
  int lo = (int)floor(pos);
  if (lo >= n - 1) return s[n - 1];
  double g = pos - (double)lo;
  double a = s[lo], b = s[lo + 1], d = b - a;
  return (g >= 0.5) ? (b - d * (1.0 - g)) : (a + d * g);   // numpy's _lerp
}
MLRN_FN double sorted_median(const double* s, int n) {
  return (n & 1) ? s[n >> 1] : 0.5 * (s[(n >> 1) - 1] + s[n >> 1]);   // numpy.median = mean of the two middles
}

// Ordinary least squares y = a x + b (numpy.polyfit(x, y, 1) for full-rank data); centred sums.
struct Line { double slope, icpt; };
template <class Team>
MLRN_TEAM_FN Line team_polyfit1(const Team& tm, const double* x, const double* y, int n) {
  double mx = 0.0, my = 0.0;                         // the two means, then the two centred sums: each pair in one butterfly
  for (int i = tm.rank(); i < n; i += Team::size) { mx += x[i]; my += y[i]; }
  tm.sum2(mx, my);
  mx /= (double)n; my /= (double)n;
  double sxx = 0.0, sxy = 0.0;
  for (int i = tm.rank(); i < n; i += Team::size) {
    double dx = x[i] - mx;
    sxx += dx * dx; sxy += dx * (y[i] - my);
  }
  tm.sum2(sxx, sxy);
  Line l; l.slope = sxy / sxx; l.icpt = my - l.slope * mx;
  return l;
}

// population std / mean over the non-NaN entries of a tiny list (rank-0 scalar code)
MLRN_FN void small_stats(const double* x, int n, int& cnt, double& mean, double& sd) {
  cnt = 0; double s = 0.0;
  for (int i = 0; i < n; ++i) if (!is_nan(x[i])) { s += x[i]; ++cnt; }
  mean = cnt ? s / cnt : qnan();
  double v = 0.0;
  for (int i = 0; i < n; ++i) if (!is_nan(x[i])) { double d = x[i] - mean; v += d * d; }
  sd = cnt ? sqrt(v / cnt) : qnan();
}

}  // namespace mlrn
