// research.cu -- K8: "research" feature block (extract_research_features, src/features/research_features.py:566; 40 columns).
// EXPERIMENTAL: the arithmetic (research.cuh) is verified against the reference on the CPU; this CUDA instantiation has not
// run on a GPU yet (the round's GPU budget was spent), so it is exported but not used by smoke / bench / the default
// pipeline, and its GPU tests are gated behind MLRN_TEST_EXPERIMENTAL=1.
//
// Warp per object, as the K3 passes: the object is staged once into the warp's shared-memory slice; three scratch arrays
// and the wavelet live beside it; the 1-day grid of the r band (up to RESEARCH_GRID_CAP days) is a per-warp slice of a
// global scratch buffer that stays in L2 (64 KB per warp).  The Mexican-hat convolutions are the only heavy part:
// ~L x 5 x scale FMA per scale, lanes over output samples.
#include <cuda_runtime.h>
#include "store.cuh"
#include "research.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

constexpr int RESEARCH_WARPS = 4;
constexpr int RESEARCH_ROW_CAP = 64;
constexpr int RESEARCH_GRID_CAP = 8192;        // days of r-band coverage the Mexican-hat block handles (22 years)

__host__ __device__ inline size_t research_bytes_per_warp(int np) {
  return (size_t)np * (8 * 3 + 8 * 3 + 2) + (RESEARCH_ROW_CAP + MHPS_WAVELET_CAP) * 8;
}

__global__ void __launch_bounds__(RESEARCH_WARPS * 32)
k_research(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
           const double* __restrict__ ge, const unsigned short* __restrict__ grank, const double* __restrict__ z,
           ObjClass cls, int np, double* __restrict__ gridbuf, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  WarpTeam tm;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned char* base = cls.scratch ? cls.scratch + ((size_t)blockIdx.x * nwarps + warp) * research_bytes_per_warp(np)
                                    : smem + (size_t)warp * research_bytes_per_warp(np);
  double* row = (double*)base;
  double* wav = row + RESEARCH_ROW_CAP;
  double* st = wav + MHPS_WAVELET_CAP;
  double* sf = st + np;
  double* se = sf + np;
  double* scr = se + np;
  unsigned short* srank = (unsigned short*)(scr + (size_t)3 * np);
  double* grid = gridbuf + ((size_t)blockIdx.x * nwarps + warp) * RESEARCH_GRID_CAP;
  for (long long slot = (long long)blockIdx.x * nwarps + warp; slot < cls.count; slot += (long long)gridDim.x * nwarps) {
    const long long o = cls.list ? cls.list[slot] : slot;
    if (!cls.list && (int)(seg_off[(o + 1) * N_SEG] - seg_off[o * N_SEG]) > cls.n_hi) continue;     // a long object: the second launch's
    ObjView v;
    stage_object(tm, seg_off, gt, gf, ge, grank, o, st, sf, se, srank, v);
    research_row(tm, v, z ? z[o] : qnan(), scr, scr + np, scr + 2 * np, grid, RESEARCH_GRID_CAP, wav, row);
    tm.sync();
    double* dst = out + o * RESEARCH_COLS;
    for (int k = tm.rank(); k < RESEARCH_COLS; k += 32) dst[k] = row[k];
    tm.sync();
  }
}

}  // namespace

int run_research(const mlrn_store* s, const double* z_dev, double* out_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  const int np = next_pow2(s->max_short > 64 ? s->max_short : 64);
  int warps = RESEARCH_WARPS;
  while (warps > 1 && research_bytes_per_warp(np) * warps > 200 * 1024) warps >>= 1;
  const size_t smem = research_bytes_per_warp(np) * warps;
  MLRN_CUDA(cudaFuncSetAttribute(k_research, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MLRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_research, warps * 32, smem));
  if (per_sm < 1) per_sm = 1;
  long long want = ((long long)s->n_obj + warps - 1) / warps, cap = (long long)s->sm_count * per_sm;
  const int grid = (int)(want < cap ? want : cap);
  int gl = 0, npl = 0;
  if (s->n_long > 0) {
    npl = next_pow2(s->max_obs);
    gl = (s->n_long + RESEARCH_WARPS - 1) / RESEARCH_WARPS;
    if (gl > s->sm_count * 2) gl = s->sm_count * 2;
  }
  const size_t grid_warps = (size_t)warps * grid > (size_t)RESEARCH_WARPS * gl ? (size_t)warps * grid : (size_t)RESEARCH_WARPS * gl;
  double* gridbuf = nullptr;
  MLRN_CUDA(cudaMallocAsync(&gridbuf, sizeof(double) * (size_t)RESEARCH_GRID_CAP * grid_warps, st));
  k_research<<<grid, warps * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->err, s->rank, z_dev, ObjClass{nullptr, s->n_obj, SHORT_LIMIT, nullptr}, np, gridbuf, out_dev);
  cudaError_t e = cudaGetLastError();
  count_launch();
  if (e == cudaSuccess && gl > 0) {                   // long objects: staging in a per-warp global scratch
    unsigned char* scratch = nullptr;
    e = cudaMallocAsync(&scratch, research_bytes_per_warp(npl) * RESEARCH_WARPS * (size_t)gl, st);
    if (e == cudaSuccess) {
      k_research<<<gl, RESEARCH_WARPS * 32, 0, st>>>(s->seg_off, s->t, s->flux, s->err, s->rank, z_dev, ObjClass{s->long_list, s->n_long, MAX_OBS, scratch}, npl, gridbuf, out_dev);
      e = cudaGetLastError();
      cudaFreeAsync(scratch, st);
      count_launch();
    }
  }
  cudaFreeAsync(gridbuf, st);
  if (e != cudaSuccess) return cuda_fail(e, "k_research launch");
  return MLRN_OK;
}

}  // namespace mlrn
