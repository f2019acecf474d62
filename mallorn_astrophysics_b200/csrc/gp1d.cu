// gp1d.cu -- K7: per-band 1-D GP features (extract_gp_features, src/features/gaussian_process.py:251), 21 columns.
//
// One WARP per object, persistent warps fed from a ticket counter; the warp fits the g, r, i, z bands one after
// the other (gp1d.cuh: four L-BFGS-B starts per band, the symmetric sweep over a packed triangle in shared
// memory per evaluation) and writes the row.  Shared memory per warp = the optimiser state (lbfgsb.cuh, owned by
// lane 0) + a scratch sized for bands of up to `cap` valid points; the rare longer band runs the same code with
// its triangle in a global scratch (L2-resident): one slot per warp while that fits in 2 GB, otherwise a pool of slots a
// warp takes for the duration of one band (a 3,000-point band needs 36 MB).  `cap` trades occupancy against how many
// bands take the slow path and is chosen by the launcher from the store's longest segment.
// (A register-tiled variant of the sweep -- 4 x 4 tiles of the triangle held in registers -- was prepared in round 1 and
// measured in round 2: 15.4 k objects/s against 66.5 k for this kernel at 100,000 objects; it was removed.)
#include <cuda_runtime.h>
#include <cstdlib>
#include "store.cuh"
#include "gp1d.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

MLRN_HD size_t gp1d_warp_bytes(int cap) {
  size_t b = sizeof(LbfgsB<3>) + 8 * sizeof(double) + sizeof(double) * (size_t)gp1d_scratch_doubles(cap);
  return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(256, 2) k_gp1d(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
                       const double* __restrict__ ge, int n_obj, int cap, int cap_global, double* __restrict__ gscratch,
                       int n_slots, int* __restrict__ slot_busy,
                       unsigned long long* __restrict__ ticket, double* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpTeam tm;
  const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  unsigned char* mine = smem_raw + (size_t)warp * gp1d_warp_bytes(cap);
  LbfgsB<3>* opt = reinterpret_cast<LbfgsB<3>*>(mine);
  double* ctl = reinterpret_cast<double*>(mine + sizeof(LbfgsB<3>));
  double* sbase = ctl + 8;
  const int my_warp = blockIdx.x * warps + warp;
  const bool own_slot = n_slots >= (int)(gridDim.x * warps);          // enough scratch for every warp: no pool
  const int bands[4] = {BG, BR, BI, BZ};
  for (;;) {
    long long o = 0;
    if (tm.rank() == 0) o = (long long)atomicAdd(ticket, 1ULL);
    o = __shfl_sync(WarpTeam::FULL, o, 0);
    if (o >= n_obj) break;
    Gp1dBandResult res[4];
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const long long lo = seg_off[o * N_SEG + bands[j]];
      const int n = (int)(seg_off[o * N_SEG + bands[j] + 1] - lo);
      // two instantiations so the common case keeps shared-memory addressing (LDS/STS instead of generic loads)
      if (n <= cap) gp1d_band(tm, gt + lo, gf + lo, ge + lo, n, gp1d_carve(sbase, cap), opt, ctl, res[j]);
      else {
        int slot = my_warp;
        if (!own_slot) {                              // take a slot of the pool for this band
          if (tm.rank() == 0) {
            slot = my_warp % n_slots;
            while (atomicCAS(slot_busy + slot, 0, 1) != 0) { slot = (slot + 1) % n_slots; __nanosleep(200); }
          }
          slot = __shfl_sync(WarpTeam::FULL, slot, 0);
        }
        gp1d_band(tm, gt + lo, gf + lo, ge + lo, n, gp1d_carve(gscratch + (size_t)slot * gp1d_scratch_doubles(cap_global), cap_global), opt, ctl, res[j]);
        if (!own_slot) {
          __threadfence();
          tm.sync();
          if (tm.rank() == 0) atomicExch(slot_busy + slot, 0);
        }
      }
      if (status && tm.rank() == 0) {
        status[o * GP1D_STATUS + 3 * j] = res[j].nfev; status[o * GP1D_STATUS + 3 * j + 1] = res[j].nit;
        status[o * GP1D_STATUS + 3 * j + 2] = res[j].winner;
      }
    }
    if (tm.rank() == 0) gp1d_tail(res, out + o * GP1D_COLS);
  }
}

}  // namespace

int run_gp1d(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  const int longest = s->max_seg > GP1D_MAX_POINTS ? GP1D_MAX_POINTS : (s->max_seg > 5 ? s->max_seg : 5);      // longer bands are not fitted (gp1d.cuh)
  int cap = longest < 52 ? longest : 52;                      // 52 points: 15 KB per warp
  if (const char* e = getenv("MLRN_GP1D_CAP")) { int v = atoi(e); if (v >= 5 && v <= 160) cap = v < longest ? v : longest; }
  const size_t per_warp = gp1d_warp_bytes(cap);
  int dev_smem = 0;
  MLRN_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
  if ((size_t)dev_smem < per_warp) { set_error("gp1d: shared memory too small for one warp"); return MLRN_EINVAL; }
  auto kernel = k_gp1d;
  const int max_warps = 8;                                       // launch bounds of the kernel
  // CTA shape = the warps-per-CTA that keeps most warps resident on an SM; small CTAs also mean a warp that draws a long
  // object does not hold many others back at the end of the grid
  int per_cta = 1, occ = 1, best = 0;
  for (int w = max_warps; w >= 1; --w) {
    const size_t smem_w = per_warp * w;
    if (smem_w > (size_t)dev_smem) continue;
    MLRN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    int o = 0;
    MLRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, w * 32, smem_w));
    if (o * w > best) { best = o * w; per_cta = w; occ = o; }
  }
  if (best < 1) { set_error("gp1d: kernel does not fit on an SM"); return MLRN_EINVAL; }
  const size_t smem = per_warp * per_cta;
  MLRN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long want = ((long long)s->n_obj + per_cta - 1) / per_cta, capg = (long long)s->sm_count * occ;
  const int grid = (int)(want < capg ? want : capg);
  double* gscratch = nullptr;
  int* slot_busy = nullptr;
  int n_slots = 0;
  if (longest > cap) {
    const size_t slot_bytes = sizeof(double) * (size_t)gp1d_scratch_doubles(longest);
    const long long all = (long long)per_cta * grid;
    long long fit = (long long)((2ULL << 30) / slot_bytes);     // at most 2 GB of scratch: beyond that the warps share a pool of slots
    if (fit < 4) fit = 4;
    n_slots = (int)(fit < all ? fit : all);
    MLRN_CUDA(cudaMallocAsync(&gscratch, slot_bytes * (size_t)n_slots, st));
    MLRN_CUDA(cudaMallocAsync(&slot_busy, sizeof(int) * (size_t)n_slots, st));
    MLRN_CUDA(cudaMemsetAsync(slot_busy, 0, sizeof(int) * (size_t)n_slots, st));
  }
  MLRN_CUDA(cudaMemsetAsync(s->ticket + 7, 0, sizeof(unsigned long long), st));
  kernel<<<grid, per_cta * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->err, s->n_obj, cap, longest, gscratch, n_slots, slot_busy,
                                           s->ticket + 7, out_dev, status_dev);
  cudaError_t e = cudaGetLastError();
  if (gscratch) cudaFreeAsync(gscratch, st);
  if (slot_busy) cudaFreeAsync(slot_busy, st);
  if (e != cudaSuccess) return cuda_fail(e, "k_gp1d launch");
  count_launch();
  return MLRN_OK;
}

}  // namespace mlrn
