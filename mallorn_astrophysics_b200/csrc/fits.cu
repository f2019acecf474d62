// fits.cu -- K4 kernels: Bazin (52 columns) and decline-model (27 columns) passes.
// One warp per object; each band is staged from the packed store into the warp's shared-memory
// slice (coalesced) and fitted there by trf_bounded() (trf.cuh).  Grid = resident CTAs x SM count,
// warps take objects from a global ticket counter (fit cost varies by 100x between objects).
#include <cuda_runtime.h>
#include "store.cuh"
#include "fits.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

constexpr int ROW_PAD = 64;       // row scratch per warp (>= 52)

__host__ __device__ inline size_t bazin_warp_doubles(int cap) { return ROW_PAD + 3 * (size_t)cap + trf_work_doubles<5>(cap); }
__host__ __device__ inline size_t decline_warp_doubles(int cap) { return ROW_PAD + 4 * (size_t)cap + trf_work_doubles<3>(cap); }

__device__ __forceinline__ long long next_ticket(unsigned long long* counter, int lane) {
  unsigned long long v = 0;
  if (lane == 0) v = atomicAdd(counter, 1ULL);
  return (long long)__shfl_sync(0xffffffffu, v, 0);
}

__global__ void __launch_bounds__(256, 2)
k_bazin(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
        const double* __restrict__ ge, int n_obj, int cap, unsigned long long* __restrict__ ticket,
        double* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  WarpTeam tm;
  const int warp = threadIdx.x >> 5;
  double* row = (double*)smem + (size_t)warp * bazin_warp_doubles(cap);
  double* st = row + ROW_PAD;
  double* sf = st + cap;
  double* se = sf + cap;
  double* work = se + cap;
  for (long long o = next_ticket(ticket, tm.rank()); o < n_obj; o = next_ticket(ticket, tm.rank())) {
    const long long* so = seg_off + o * N_SEG;
    for (int b = 0; b < N_BAND; ++b) {
      const long long lo = so[b];
      const int n = (int)(so[b + 1] - lo);
      for (int i = tm.rank(); i < n; i += 32) { st[i] = gt[lo + i]; sf[i] = gf[lo + i]; se[i] = ge[lo + i]; }
      tm.sync();
      int code, nfev;
      bazin_band(tm, st, sf, se, n, work, work, cap, row + 8 * b, &code, &nfev);
      if (status && tm.rank() == 0) { status[o * 12 + 2 * b] = code; status[o * 12 + 2 * b + 1] = nfev; }
      tm.sync();
    }
    bazin_tail(tm, row);
    for (int k = tm.rank(); k < BAZIN_COLS; k += 32) out[o * BAZIN_COLS + k] = row[k];
    tm.sync();
  }
}

__global__ void __launch_bounds__(256, 2)
k_powerlaw(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
           int n_obj, int cap, unsigned long long* __restrict__ ticket, double* __restrict__ out,
           int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  WarpTeam tm;
  const int warp = threadIdx.x >> 5;
  double* row = (double*)smem + (size_t)warp * decline_warp_doubles(cap);
  double* st = row + ROW_PAD;
  double* sf = st + cap;
  double* xs = sf + cap;
  double* ys = xs + cap;
  double* work = ys + cap;
  const int bands[3] = {BG, BR, BI};
  for (long long o = next_ticket(ticket, tm.rank()); o < n_obj; o = next_ticket(ticket, tm.rank())) {
    const long long* so = seg_off + o * N_SEG;
    for (int j = 0; j < 3; ++j) {
      const long long lo = so[bands[j]];
      const int n = (int)(so[bands[j] + 1] - lo);
      for (int i = tm.rank(); i < n; i += 32) { st[i] = gt[lo + i]; sf[i] = gf[lo + i]; }
      tm.sync();
      decline_band(tm, st, sf, n, xs, ys, work, cap, row + 9 * j, status ? status + o * 54 + 18 * j : nullptr);
      tm.sync();
    }
    for (int k = tm.rank(); k < POWERLAW_COLS; k += 32) out[o * POWERLAW_COLS + k] = row[k];
    tm.sync();
  }
}

template <class K>
int fit_geometry(K kernel, size_t warp_bytes, const mlrn_store* s, int& warps, size_t& smem, int& grid) {
  warps = 8;
  while (warps > 1 && warp_bytes * warps > 100 * 1024) warps >>= 1;     // keep >= 2 CTAs per SM where possible
  smem = warp_bytes * warps;
  if (smem > 227 * 1024) { set_error("band too long for the shared-memory fit workspace"); return MLRN_ETOOLONG; }
  MLRN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MLRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
  if (per_sm < 1) per_sm = 1;
  long long want = ((long long)s->n_obj + warps - 1) / warps;
  long long cap = (long long)s->sm_count * per_sm;
  grid = (int)(want < cap ? want : cap);
  return MLRN_OK;
}

}  // namespace

int run_bazin(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  const int cap = (s->max_seg + 1) & ~1;
  int warps, grid; size_t smem;
  int rc = fit_geometry(k_bazin, bazin_warp_doubles(cap) * sizeof(double), s, warps, smem, grid);
  if (rc != MLRN_OK) return rc;
  MLRN_CUDA(cudaMemsetAsync(s->ticket, 0, sizeof(unsigned long long), st));
  k_bazin<<<grid, warps * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->err, s->n_obj, cap, s->ticket, out_dev, status_dev);
  MLRN_CUDA(cudaGetLastError());
  count_launch();
  return MLRN_OK;
}

int run_powerlaw(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  const int cap = (s->max_seg + 1) & ~1;
  int warps, grid; size_t smem;
  int rc = fit_geometry(k_powerlaw, decline_warp_doubles(cap) * sizeof(double), s, warps, smem, grid);
  if (rc != MLRN_OK) return rc;
  MLRN_CUDA(cudaMemsetAsync(s->ticket + 1, 0, sizeof(unsigned long long), st));
  k_powerlaw<<<grid, warps * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->n_obj, cap, s->ticket + 1, out_dev, status_dev);
  MLRN_CUDA(cudaGetLastError());
  count_launch();
  return MLRN_OK;
}

}  // namespace mlrn
