// pack.cu -- K1: flat long-format rows in file order -> CSR (object, band) segments in HBM.
//
// Replaces the `groupby('object_id')` dict every reference extractor rebuilds (e.g.
// src/features/statistical.py:155) plus the per-band boolean mask + `sort_values('Time (MJD)')`
// (e.g. src/features/lightcurve_shape.py:192).  Integer outputs (offsets, counts, peak
// positions, peak_band) are bit-exact by construction: no floating-point arithmetic is involved,
// only comparisons.
//
// Pipeline (all on one stream):
//   k_count        histogram of rows per (object, band)                       atomics, HBM-bound
//   k_obj_partial / k_scan_partials / k_obj_offsets   exclusive scan -> seg_off
//   k_scatter      row ids into their object's slot range (arbitrary order)   atomics
//   k_arrange      one warp per object: sort row ids (recovers FILE order), sort by (time, file
//                  rank), stable split by band, gather t/flux/err, emit file ranks + segment meta
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include "store.cuh"
#include "objview.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;   // objects per block

__global__ void k_count(const int32_t* __restrict__ obj, const uint8_t* __restrict__ band, int64_t n_rows,
                        int n_obj, int* __restrict__ seg_cnt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x) {
    int o = obj[i];
    if (o >= 0 && o < n_obj) {
      int b = band[i];
      b = b > 6 ? 6 : b;
      atomicAdd(&seg_cnt[(int64_t)o * N_SEG + b], 1);
    }
  }
}

__device__ __forceinline__ int obj_total(const int* seg_cnt, int64_t o) {
  const int* c = seg_cnt + o * N_SEG;
  return c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6];
}

// block sum of per-object totals -> partial[blockIdx]; also the largest object
// flags: [0] largest object, [1] error flag of k_arrange, [2] largest segment, [4] largest SHORT object, [5] number of long
// objects (their indices go to long_list)
__global__ void k_obj_partial(const int* __restrict__ seg_cnt, int n_obj, long long* __restrict__ partial,
                              int* __restrict__ max_obs, int* __restrict__ long_list) {
  __shared__ long long ws[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  long long s = 0; int mx = 0, ms = 0, mshort = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t o = base + k;
    if (o < n_obj) {
      int c = obj_total(seg_cnt, o); s += c; mx = max(mx, c);
      if (c <= SHORT_LIMIT) mshort = max(mshort, c);
      else long_list[atomicAdd(max_obs + 5, 1)] = (int)o;
      for (int b = 0; b < N_SEG; ++b) ms = max(ms, seg_cnt[o * N_SEG + b]);
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, d); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    ms = max(ms, __shfl_xor_sync(0xffffffffu, ms, d)); mshort = max(mshort, __shfl_xor_sync(0xffffffffu, mshort, d));
  }
  if ((threadIdx.x & 31) == 0) {
    ws[threadIdx.x >> 5] = s; atomicMax(max_obs, mx); atomicMax(max_obs + 2, ms); atomicMax(max_obs + 4, mshort);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) tot += ws[w];
    partial[blockIdx.x] = tot;
  }
}

// exclusive scan of the block partials, in place; partial[n_blocks] receives the grand total.
__global__ void k_scan_partials(long long* partial, int n_blocks) {
  __shared__ long long sh[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    int i = base + threadIdx.x;
    long long v = (i < n_blocks) ? partial[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      long long add = (threadIdx.x >= d) ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += add;
      __syncthreads();
    }
    long long incl = sh[threadIdx.x];
    if (i < n_blocks) partial[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[n_blocks] = carry;
}

__global__ void k_obj_offsets(const int* __restrict__ seg_cnt, int n_obj, const long long* __restrict__ partial,
                              long long* __restrict__ seg_off) {
  __shared__ long long ws[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  int tot[SCAN_ITEMS];
  long long s = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t o = base + k;
    tot[k] = (o < n_obj) ? obj_total(seg_cnt, o) : 0;
    s += tot[k];
  }
  // exclusive scan of s across the block
  long long incl = s;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int d = 1; d < 32; d <<= 1) { long long u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
  if (lane == 31) ws[w] = incl;
  __syncthreads();
  long long wpre = 0;
  for (int k = 0; k < w; ++k) wpre += ws[k];
  long long run = partial[blockIdx.x] + wpre + incl - s;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t o = base + k;
    if (o < n_obj) {
      const int* c = seg_cnt + o * N_SEG;
      long long r = run;
      for (int b = 0; b < N_SEG; ++b) { seg_off[o * N_SEG + b] = r; r += c[b]; }
      run += tot[k];
      if (o == n_obj - 1) seg_off[(int64_t)n_obj * N_SEG] = r;
    }
  }
}

__global__ void k_scatter(const int32_t* __restrict__ obj, int64_t n_rows, int n_obj,
                          const long long* __restrict__ seg_off, int* __restrict__ cursor, int* __restrict__ rowid) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x) {
    int o = obj[i];
    if (o >= 0 && o < n_obj) {
      int p = atomicAdd(&cursor[o], 1);
      rowid[seg_off[(int64_t)o * N_SEG] + p] = (int)i;
    }
  }
}

// ---- warp-level bitonic sorts in shared memory -------------------------------------------------
__device__ void warp_sort_int(int* a, int np, int lane) {
  for (int k = 2; k <= np; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < np; i += 32) {
        int p = i ^ j;
        if (p > i) {
          int x = a[i], y = a[p];
          bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[p] = x; }
        }
      }
      __syncwarp();
    }
}
// sort the permutation perm[0..np) by (tt[perm], perm); entries >= n act as +infinity
__device__ void warp_sort_perm(unsigned short* perm, const double* tt, int n, int np, int lane) {
  for (int k = 2; k <= np; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < np; i += 32) {
        int p = i ^ j;
        if (p > i) {
          int x = perm[i], y = perm[p];
          bool gt;                       // x sorts after y ?
          if (x >= n || y >= n) gt = (x > y);
          else { double tx = tt[x], ty = tt[y]; gt = (tx > ty) || (tx == ty && x > y); }
          bool up = ((i & k) == 0);
          if (gt == up) { perm[i] = (unsigned short)y; perm[p] = (unsigned short)x; }
        }
      }
      __syncwarp();
    }
}

constexpr int ARR_WARPS = 4;
__host__ __device__ inline size_t arrange_bytes_per_warp(int np) {
  return (size_t)np * (8 + 8 + 8 + 4 + 2 + 2);   // tt, ff, fs, ids, perm, bb(+pad to 2)
}

__global__ void __launch_bounds__(ARR_WARPS * 32)
k_arrange(const uint8_t* __restrict__ band_in, const double* __restrict__ t_in, const double* __restrict__ f_in,
          const double* __restrict__ e_in, const int* __restrict__ rowid, const long long* __restrict__ seg_off,
          ObjClass cls, int np, double* __restrict__ t_out, double* __restrict__ f_out, double* __restrict__ e_out,
          unsigned short* __restrict__ rank_out, unsigned short* __restrict__ trank_out, int* __restrict__ seg_meta, int* __restrict__ err_flag) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* base = cls.scratch ? cls.scratch + ((size_t)blockIdx.x * ARR_WARPS + warp) * arrange_bytes_per_warp(np)
                                    : smem + (size_t)warp * arrange_bytes_per_warp(np);
  double* tt = (double*)base;
  double* ff = tt + np;
  double* fs = ff + np;
  int* ids = (int*)(fs + np);
  unsigned short* perm = (unsigned short*)(ids + np);
  unsigned short* bb = perm + np;

  for (int64_t slot = (int64_t)blockIdx.x * ARR_WARPS + warp; slot < cls.count; slot += (int64_t)gridDim.x * ARR_WARPS) {
    const int64_t o = cls.list ? cls.list[slot] : slot;
    const long long* so = seg_off + o * N_SEG;
    const long long gbase = so[0];
    const int n = (int)(so[N_SEG] - gbase);
    int* meta = seg_meta + o * MLRN_SEGMETA_COLS;
    if (!cls.list && n > cls.n_hi) continue;          // a long object: the second launch's
    if (n > MAX_OBS) { if (lane == 0) atomicMax(err_flag, n); continue; }
    int off[N_SEG + 1];
    for (int b = 0; b <= N_SEG; ++b) off[b] = (int)(so[b] - gbase);
    if (n == 0) {
      for (int k = lane; k < MLRN_SEGMETA_COLS; k += 32) meta[k] = (k % 3 == 0 && k < 21) ? 0 : -1;
      continue;
    }
    const int npo = next_pow2(n);
    // (a) row ids ascending = file order
    for (int i = lane; i < npo; i += 32) ids[i] = (i < n) ? rowid[gbase + i] : INT_MAX;
    __syncwarp();
    warp_sort_int(ids, npo, lane);
    // (b) gather keys
    for (int i = lane; i < npo; i += 32) {
      if (i < n) {
        int r = ids[i];
        tt[i] = t_in[r]; ff[i] = f_in[r];
        int b = band_in[r]; bb[i] = (unsigned short)(b > 6 ? 6 : b);
      }
      perm[i] = (unsigned short)i;
    }
    __syncwarp();
    // (c) order by (time, file rank)
    warp_sort_perm(perm, tt, n, npo, lane);
    // (d) stable split by band + gather payload
    int run[N_SEG];
    for (int b = 0; b < N_SEG; ++b) run[b] = off[b];
    for (int s = 0; s < n; s += 32) {
      int i = s + lane;
      bool valid = i < n;
      int src = valid ? perm[i] : 0;
      int b = valid ? bb[src] : -1;
      int dest = 0;
#pragma unroll
      for (int q = 0; q < N_SEG; ++q) {
        unsigned m = __ballot_sync(0xffffffffu, b == q);
        if (b == q) dest = run[q] + __popc(m & ((1u << lane) - 1u));
        run[q] += __popc(m);
      }
      if (valid) {
        int r = ids[src];
        t_out[gbase + dest] = tt[src];
        f_out[gbase + dest] = ff[src];
        e_out[gbase + dest] = e_in[r];
        rank_out[gbase + dest] = (unsigned short)src;
        trank_out[gbase + dest] = (unsigned short)i;
        fs[dest] = ff[src];
      }
    }
    __syncwarp();
    // (e) segment meta: count, first-argmax position in file order / in time order, peak_band
    int peak_band = -1; double peak_val = 0.0;
    for (int b = 0; b < N_SEG; ++b) {
      int cnt = off[b + 1] - off[b];
      int pos_file = -1, pos_time = -1;
      if (cnt > 0) {
        double m = -pinf();
        for (int i = off[b] + lane; i < off[b + 1]; i += 32) m = fmax(m, fs[i]);
        for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
        int bt = INT_MAX;                               // first max in time order
        for (int i = off[b] + lane; i < off[b + 1]; i += 32) if (fs[i] == m) { bt = min(bt, i - off[b]); }
        bt = __reduce_min_sync(0xffffffffu, bt);
        int bf = INT_MAX;                               // smallest file rank attaining the max
        for (int i = lane; i < n; i += 32) if (bb[i] == b && ff[i] == m) bf = min(bf, i);
        bf = __reduce_min_sync(0xffffffffu, bf);
        int before = 0;
        for (int i = lane; i < n; i += 32) before += (bb[i] == b && i < bf);
        before = __reduce_add_sync(0xffffffffu, before);
        if (bt != INT_MAX) { pos_time = bt; pos_file = before; }   // all-NaN band keeps -1
        if (b < N_BAND && m == m && (peak_band < 0 || m > peak_val)) { peak_val = m; peak_band = b; }
      }
      if (lane == 0) { meta[3 * b] = cnt; meta[3 * b + 1] = pos_file; meta[3 * b + 2] = pos_time; }
    }
    if (lane == 0) meta[21] = peak_band;
    __syncwarp();
  }
}

}  // namespace

// Builds the packed arrays of `s` from device-resident flat columns.  Temporary buffers are
// stream-ordered allocations.
int pack_device(mlrn_store* s, int64_t n_rows, const int32_t* obj, const uint8_t* band, const double* t,
                const double* flux, const double* err, cudaStream_t st) {
  const int n_obj = s->n_obj;
  s->stream = st;                       // the store's arrays are stream-ordered allocations (cached by the pool across batches)
  const int64_t n_seg = (int64_t)n_obj * N_SEG;
  const int n_blocks = (n_obj + SCAN_CHUNK - 1) / SCAN_CHUNK;
  int *seg_cnt = nullptr, *cursor = nullptr, *rowid = nullptr, *flags = nullptr;
  long long* partial = nullptr;
  MLRN_CUDA(cudaMallocAsync(&seg_cnt, sizeof(int) * (size_t)(n_seg > 0 ? n_seg : 1), st));
  MLRN_CUDA(cudaMallocAsync(&cursor, sizeof(int) * (size_t)(n_obj > 0 ? n_obj : 1), st));
  MLRN_CUDA(cudaMallocAsync(&partial, sizeof(long long) * (size_t)(n_blocks + 1), st));
  MLRN_CUDA(cudaMallocAsync(&flags, sizeof(int) * 8, st));
  MLRN_CUDA(cudaMallocAsync(&s->long_list, sizeof(int) * (size_t)(n_obj > 0 ? n_obj : 1), st));
  MLRN_CUDA(cudaMemsetAsync(seg_cnt, 0, sizeof(int) * (size_t)(n_seg > 0 ? n_seg : 1), st));
  MLRN_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)(n_obj > 0 ? n_obj : 1), st));
  MLRN_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 8, st));
  MLRN_CUDA(cudaMallocAsync(&s->seg_off, sizeof(long long) * (size_t)(n_seg + 1), st));
  MLRN_CUDA(cudaMallocAsync(&s->seg_meta, sizeof(int32_t) * (size_t)(n_obj > 0 ? n_obj : 1) * MLRN_SEGMETA_COLS, st));
  MLRN_CUDA(cudaMallocAsync(&s->ticket, sizeof(unsigned long long) * 8, st));

  const int row_grid = (int)((n_rows + 255) / 256 < (int64_t)s->sm_count * 16 ? (n_rows + 255) / 256 : (int64_t)s->sm_count * 16);
  if (n_rows > 0) { k_count<<<row_grid > 0 ? row_grid : 1, 256, 0, st>>>(obj, band, n_rows, n_obj, seg_cnt); count_launch(); }
  if (n_obj > 0) {
    k_obj_partial<<<n_blocks, SCAN_THREADS, 0, st>>>(seg_cnt, n_obj, partial, flags, s->long_list);
    k_scan_partials<<<1, 1024, 0, st>>>(partial, n_blocks);
    k_obj_offsets<<<n_blocks, SCAN_THREADS, 0, st>>>(seg_cnt, n_obj, partial, s->seg_off);
    count_launch(3);
  } else {
    MLRN_CUDA(cudaMemsetAsync(s->seg_off, 0, sizeof(long long), st));
  }
  MLRN_CUDA(cudaGetLastError());
  long long total = 0; int h_flags[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  MLRN_CUDA(cudaMemcpyAsync(&total, s->seg_off + n_seg, sizeof(long long), cudaMemcpyDeviceToHost, st));
  MLRN_CUDA(cudaMemcpyAsync(h_flags, flags, sizeof(int) * 8, cudaMemcpyDeviceToHost, st));
  MLRN_CUDA(cudaStreamSynchronize(st));
  s->n_rows = total;
  s->max_obs = h_flags[0];
  s->max_seg = h_flags[2];
  s->max_short = h_flags[4];
  s->n_long = h_flags[5];
  if (s->max_obs > MAX_OBS) {
    cudaFreeAsync(seg_cnt, st); cudaFreeAsync(cursor, st); cudaFreeAsync(partial, st); cudaFreeAsync(flags, st);
    set_error("an object has " + std::to_string(s->max_obs) + " observations; the per-object cap is " + std::to_string(MAX_OBS));
    return MLRN_ETOOLONG;
  }
  const size_t nr = (size_t)(total > 0 ? total : 1);
  MLRN_CUDA(cudaMallocAsync(&s->t, sizeof(double) * nr, st));
  MLRN_CUDA(cudaMallocAsync(&s->flux, sizeof(double) * nr, st));
  MLRN_CUDA(cudaMallocAsync(&s->err, sizeof(double) * nr, st));
  MLRN_CUDA(cudaMallocAsync(&s->rank, sizeof(unsigned short) * nr, st));
  MLRN_CUDA(cudaMallocAsync(&s->trank, sizeof(unsigned short) * nr, st));
  MLRN_CUDA(cudaMallocAsync(&rowid, sizeof(int) * nr, st));
  if (n_obj > 0) {
    // k_arrange is the only writer of seg_meta, so it also runs for a batch without rows (every object then gets
    // the count-0 / position -1 pattern and no input column is touched)
    if (n_rows > 0) k_scatter<<<row_grid > 0 ? row_grid : 1, 256, 0, st>>>(obj, n_rows, n_obj, s->seg_off, cursor, rowid);
    // short objects: staging in shared memory sized by the longest of them
    const int np = next_pow2(s->max_short > 1 ? s->max_short : 1);
    const size_t smem = arrange_bytes_per_warp(np) * ARR_WARPS;
    MLRN_CUDA(cudaFuncSetAttribute(k_arrange, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (n_obj + ARR_WARPS - 1) / ARR_WARPS;
    const int cap = s->sm_count * 8;
    if (grid > cap) grid = cap;
    k_arrange<<<grid, ARR_WARPS * 32, smem, st>>>(band, t, flux, err, rowid, s->seg_off, ObjClass{nullptr, n_obj, SHORT_LIMIT, nullptr}, np,
                                                   s->t, s->flux, s->err, s->rank, s->trank, s->seg_meta, flags + 1);
    if (s->n_long > 0) {                              // long objects: the same code, staging in a global scratch
      const int npl = next_pow2(s->max_obs);
      int gl = (s->n_long + ARR_WARPS - 1) / ARR_WARPS;
      if (gl > s->sm_count * 2) gl = s->sm_count * 2;
      unsigned char* scratch = nullptr;
      MLRN_CUDA(cudaMallocAsync(&scratch, arrange_bytes_per_warp(npl) * ARR_WARPS * (size_t)gl, st));
      k_arrange<<<gl, ARR_WARPS * 32, 0, st>>>(band, t, flux, err, rowid, s->seg_off, ObjClass{s->long_list, s->n_long, MAX_OBS, scratch}, npl,
                                                s->t, s->flux, s->err, s->rank, s->trank, s->seg_meta, flags + 1);
      cudaFreeAsync(scratch, st);
      count_launch();
    }
    count_launch(2);
  }
  MLRN_CUDA(cudaGetLastError());
  cudaFreeAsync(seg_cnt, st); cudaFreeAsync(cursor, st); cudaFreeAsync(partial, st); cudaFreeAsync(flags, st);
  cudaFreeAsync(rowid, st);
  return MLRN_OK;
}

}  // namespace mlrn
