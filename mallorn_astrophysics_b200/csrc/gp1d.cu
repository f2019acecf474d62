// gp1d.cu -- K7: per-band 1-D GP features (extract_gp_features, src/features/gaussian_process.py:251), 21 columns.
//
// One WARP per object, persistent warps fed from a ticket counter; the warp fits the g, r, i, z bands one after
// the other (gp1d.cuh: four L-BFGS-B starts per band, the symmetric sweep over a packed triangle in shared
// memory per evaluation) and writes the row.  Shared memory per warp = the optimiser state (lbfgsb.cuh, owned by
// lane 0) + a scratch sized for bands of up to `cap` valid points; the rare longer band runs the same code with
// its triangle in a per-warp global scratch (L2-resident).  `cap` trades occupancy against how many bands take
// the slow path and is chosen by the launcher from the store's longest segment.
#include <cuda_runtime.h>
#include <cstdlib>
#include "store.cuh"
#include "gp1d.cuh"
#include "gp1d_tiled.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

MLRN_HD size_t gp1d_warp_bytes(int cap) {
  size_t b = sizeof(LbfgsB<3>) + 8 * sizeof(double) + sizeof(double) * (size_t)gp1d_scratch_doubles(cap);
  return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(256, 2) k_gp1d(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
                       const double* __restrict__ ge, int n_obj, int cap, int cap_global, double* __restrict__ gscratch,
                       unsigned long long* __restrict__ ticket, double* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpTeam tm;
  const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  unsigned char* mine = smem_raw + (size_t)warp * gp1d_warp_bytes(cap);
  LbfgsB<3>* opt = reinterpret_cast<LbfgsB<3>*>(mine);
  double* ctl = reinterpret_cast<double*>(mine + sizeof(LbfgsB<3>));
  double* sbase = ctl + 8;
  double* gbase = gscratch ? gscratch + ((size_t)blockIdx.x * warps + warp) * (size_t)gp1d_scratch_doubles(cap_global) : nullptr;
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
      else gp1d_band(tm, gt + lo, gf + lo, ge + lo, n, gp1d_carve(gbase, cap_global), opt, ctl, res[j]);
      if (status && tm.rank() == 0) {
        status[o * GP1D_STATUS + 3 * j] = res[j].nfev; status[o * GP1D_STATUS + 3 * j + 1] = res[j].nit;
        status[o * GP1D_STATUS + 3 * j + 2] = res[j].winner;
      }
    }
    if (tm.rank() == 0) gp1d_tail(res, out + o * GP1D_COLS);
  }
}

// EXPERIMENTAL (MLRN_GP1D_TILED=1): same mapping, but bands of up to 51 valid points evaluate the objective with the
// matrix in registers (gp1d_tiled.cuh); longer bands fall back to the shared-memory / global sweep above.
__global__ void __launch_bounds__(128, 3) k_gp1d_tiled(const long long* __restrict__ seg_off, const double* __restrict__ gt,
                                                        const double* __restrict__ gf, const double* __restrict__ ge, int n_obj, int cap,
                                                        int cap_global, double* __restrict__ gscratch,
                                                        unsigned long long* __restrict__ ticket, double* __restrict__ out,
                                                        int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpTeam tm;
  const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  unsigned char* mine = smem_raw + (size_t)warp * gp1d_warp_bytes(cap);
  LbfgsB<3>* opt = reinterpret_cast<LbfgsB<3>*>(mine);
  double* ctl = reinterpret_cast<double*>(mine + sizeof(LbfgsB<3>));
  double* sbase = ctl + 8;
  double* gbase = gscratch ? gscratch + ((size_t)blockIdx.x * warps + warp) * (size_t)gp1d_scratch_doubles(cap_global) : nullptr;
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
      const int nraw = (int)(seg_off[o * N_SEG + bands[j] + 1] - lo);
      if (nraw <= cap && nraw + 1 <= gp1d_tiled_capacity(3)) {
        gp1d_clear(res[j]);
        const Gp1dTiledScratch ts = gp1d_carve_tiled(sbase, cap);
        double t_range = 0.0, f_std = 1.0;
        const int n = gp1d_prepare(tm, gt + lo, gf + lo, ge + lo, nraw, ts.x, ts.y, ts.al, t_range, f_std);
        if (n > 0) {
          if (n + 1 <= gp1d_tiled_capacity(2)) {
            WarpLanes<TileRegs<2>> ex;
            gp1d_optimise(tm, [&](const double* th, double& fv, double* g) {
              gp1d_objective_tiled<2>(ex, n, ts.x, ts.y, ts.al, ts.vb0, ts.vb1, ts.piv, th, fv, g); }, t_range, f_std, opt, ctl, res[j]);
          } else {
            WarpLanes<TileRegs<3>> ex;
            gp1d_optimise(tm, [&](const double* th, double& fv, double* g) {
              gp1d_objective_tiled<3>(ex, n, ts.x, ts.y, ts.al, ts.vb0, ts.vb1, ts.piv, th, fv, g); }, t_range, f_std, opt, ctl, res[j]);
          }
        }
      } else if (nraw <= cap) gp1d_band(tm, gt + lo, gf + lo, ge + lo, nraw, gp1d_carve(sbase, cap), opt, ctl, res[j]);
      else gp1d_band(tm, gt + lo, gf + lo, ge + lo, nraw, gp1d_carve(gbase, cap_global), opt, ctl, res[j]);
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
  const int longest = s->max_seg > 5 ? s->max_seg : 5;
  int cap = longest < 52 ? longest : 52;                      // 52 points: 15 KB per warp
  if (const char* e = getenv("MLRN_GP1D_CAP")) { int v = atoi(e); if (v >= 5 && v <= 160) cap = v < longest ? v : longest; }
  const size_t per_warp = gp1d_warp_bytes(cap);
  int dev_smem = 0;
  MLRN_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
  if ((size_t)dev_smem < per_warp) { set_error("gp1d: shared memory too small for one warp"); return MLRN_EINVAL; }
  const char* tiled_env = getenv("MLRN_GP1D_TILED");
  const bool tiled = tiled_env && tiled_env[0] == '1';          // experimental register-tiled objective (gp1d_tiled.cuh)
  auto kernel = tiled ? k_gp1d_tiled : k_gp1d;
  const int max_warps = tiled ? 4 : 8;                           // launch bounds of the two kernels
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
  if (longest > cap)
    MLRN_CUDA(cudaMallocAsync(&gscratch, sizeof(double) * (size_t)gp1d_scratch_doubles(longest) * per_cta * grid, st));
  MLRN_CUDA(cudaMemsetAsync(s->ticket + 7, 0, sizeof(unsigned long long), st));
  kernel<<<grid, per_cta * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->err, s->n_obj, cap, longest, gscratch, s->ticket + 7,
                                           out_dev, status_dev);
  cudaError_t e = cudaGetLastError();
  if (gscratch) cudaFreeAsync(gscratch, st);
  if (e != cudaSuccess) return cuda_fail(e, "k_gp1d launch");
  count_launch();
  return MLRN_OK;
}

}  // namespace mlrn
