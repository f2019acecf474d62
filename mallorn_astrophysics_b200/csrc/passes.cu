// passes.cu -- warp-per-object K3 scalar feature passes over the packed store (the K2 moment pass is
// stat_fast.cu).  One warp stages one object's observations into its shared-memory slice with
// coalesced loads, computes the module's row there, and writes the row back coalesced.
// Grid = a multiple of the SM count; warps stride over objects.
#include <cuda_runtime.h>
#include <stdlib.h>
#ifndef MLRN_K3_INLINE             // developer build: the helpers inlined, as before
#define MLRN_SHARED_HELPERS      // team.cuh: one copy of every team helper per kernel (instruction-cache footprint of the rows)
#endif
#include "store.cuh"
#include "stat.cuh"
#include "colors.cuh"
#include "shape.cuh"
#include "tde.cuh"
#include "physics.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

constexpr int PASS_WARPS = 4;
constexpr int ROW_CAP = 128;      // >= widest row (123)

enum PassId { PASS_STAT = 0, PASS_COLOR, PASS_SHAPE, PASS_TDE, PASS_PHYSICS };

template <int PASS> struct PassTraits;
// scratch: arrays of the row routine; sorts: the routine runs the bitonic network in its scratch (capacity a power of two)
template <> struct PassTraits<PASS_COLOR> { static constexpr int cols = COLOR_COLS; static constexpr int scratch = 2; static constexpr bool sorts = false; };
template <> struct PassTraits<PASS_SHAPE> { static constexpr int cols = SHAPE_COLS; static constexpr int scratch = 2; static constexpr bool sorts = true; };
template <> struct PassTraits<PASS_TDE> { static constexpr int cols = TDE_COLS; static constexpr int scratch = 4; static constexpr bool sorts = false; };
template <> struct PassTraits<PASS_PHYSICS> { static constexpr int cols = PHYSICS_COLS; static constexpr int scratch = 2; static constexpr bool sorts = true; };

// One warp's staging slice: the row, the staged (t, flux, err) of `nst` observations, `n_scratch` scratch arrays of `nsc`
// doubles and the file ranks.  The K3 rows are latency-bound chains of small warp reductions, so what they need is resident
// warps, and shared memory is what limits those: nst is the longest short object of the batch rounded up to 4 (not to a
// power of two) and only the passes that sort get power-of-two scratch.
__host__ __device__ inline size_t pass_bytes_per_warp(int nst, int nsc, int n_scratch) {
  return (((size_t)nst * (8 * 3 + 2) + (size_t)nsc * 8 * n_scratch + ROW_CAP * 8) + 15) & ~(size_t)15;
}

template <int PASS>
__global__ void __launch_bounds__(PASS_WARPS * 32)
k_pass(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
       const double* __restrict__ ge, const unsigned short* __restrict__ grank, const double* __restrict__ z,
       ObjClass cls, int nst, int nsc, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NS = PassTraits<PASS>::scratch;
  constexpr int COLS = PassTraits<PASS>::cols;
  WarpTeam tm;
  const int warp = threadIdx.x >> 5;
  unsigned char* base = cls.scratch ? cls.scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * pass_bytes_per_warp(nst, nsc, NS)
                                    : smem + (size_t)warp * pass_bytes_per_warp(nst, nsc, NS);
  double* row = (double*)base;
  double* st = row + ROW_CAP;
  double* sf = st + nst;
  double* se = sf + nst;
  double* scr = se + nst;
  unsigned short* srank = (unsigned short*)(scr + (size_t)NS * nsc);
  const int np = nsc;

  const int nwarps = blockDim.x >> 5;     // 4, or fewer when long objects need more scratch per warp
  for (long long slot = (long long)blockIdx.x * nwarps + warp; slot < cls.count; slot += (long long)gridDim.x * nwarps) {
    const long long o = cls.list ? cls.list[slot] : slot;
    if (!cls.list && (int)(seg_off[(o + 1) * N_SEG] - seg_off[o * N_SEG]) > cls.n_hi) continue;     // a long object: the second launch's
    ObjView v;
    stage_object(tm, seg_off, gt, gf, ge, grank, o, st, sf, se, srank, v);
    if (PASS == PASS_COLOR) color_row(tm, v, scr, scr + np, row);
    if (PASS == PASS_SHAPE) shape_row(tm, v, scr, scr + np, row);
    if (PASS == PASS_TDE) tde_row(tm, v, scr, scr + np, scr + 2 * np, scr + 3 * np, row);
    if (PASS == PASS_PHYSICS) physics_row(tm, v, z ? z[o] : qnan(), scr, scr + np, row);
    tm.sync();
    double* dst = out + o * COLS;
    for (int k = tm.rank(); k < COLS; k += 32) dst[k] = row[k];
    tm.sync();
  }
}

template <int PASS>
int launch_pass(const mlrn_store* s, const double* z, double* out, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  // short objects: shared-memory staging sized by the longest of them
  constexpr int NSC = PassTraits<PASS>::scratch;
  static const bool tight = [] { const char* e = getenv("MLRN_K3_TIGHT"); return !(e && e[0] == '0'); }();      // developer knob: 0 = power-of-two staging
  const int nmax = s->max_short > 64 ? s->max_short : 64;           // scratch also holds small fixed tables
  const int nsc = (tight && !PassTraits<PASS>::sorts) ? ((nmax + 3) & ~3) : next_pow2(nmax);
  const int nst = tight ? ((nmax + 3) & ~3) : next_pow2(nmax);
  // CTA shape: 4 warps, or 2 when that leaves more warps resident (shared memory is granted per CTA)
  int warps = PASS_WARPS, per_sm = 0;
  while (warps > 1 && pass_bytes_per_warp(nst, nsc, NSC) * warps > 200 * 1024) warps >>= 1;
  size_t smem = pass_bytes_per_warp(nst, nsc, NSC) * warps;
  MLRN_CUDA(cudaFuncSetAttribute(k_pass<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MLRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pass<PASS>, warps * 32, smem));
  if (tight && warps == PASS_WARPS) {
    int per_sm2 = 0;
    const size_t smem2 = pass_bytes_per_warp(nst, nsc, NSC) * 2;
    MLRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_pass<PASS>, 64, smem2));
    if (per_sm2 * 2 > per_sm * warps) { warps = 2; per_sm = per_sm2; smem = smem2; }
  }
  if (per_sm < 1) per_sm = 1;
  long long want = ((long long)s->n_obj + warps - 1) / warps;
  long long cap = (long long)s->sm_count * per_sm;
  int grid = (int)(want < cap ? want : cap);
  k_pass<PASS><<<grid, warps * 32, smem, st>>>(s->seg_off, s->t, s->flux, s->err, s->rank, z, ObjClass{nullptr, s->n_obj, SHORT_LIMIT, nullptr}, nst, nsc, out);
  MLRN_CUDA(cudaGetLastError());
  count_launch();
  if (s->n_long > 0) {                                // long objects: the same code, staging in a per-warp global scratch
    const int npl = next_pow2(s->max_obs);
    int gl = (s->n_long + PASS_WARPS - 1) / PASS_WARPS;
    if (gl > s->sm_count * 2) gl = s->sm_count * 2;
    unsigned char* scratch = nullptr;
    MLRN_CUDA(cudaMallocAsync(&scratch, pass_bytes_per_warp(npl, npl, NSC) * PASS_WARPS * (size_t)gl, st));
    k_pass<PASS><<<gl, PASS_WARPS * 32, 0, st>>>(s->seg_off, s->t, s->flux, s->err, s->rank, z, ObjClass{s->long_list, s->n_long, MAX_OBS, scratch}, npl, npl, out);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return cuda_fail(e, "k_pass (long objects) launch");
    count_launch();
  }
  return MLRN_OK;
}

}  // namespace

int run_color(const mlrn_store* s, double* out_dev, cudaStream_t st) { return launch_pass<PASS_COLOR>(s, nullptr, out_dev, st); }
int run_shape(const mlrn_store* s, double* out_dev, cudaStream_t st) { return launch_pass<PASS_SHAPE>(s, nullptr, out_dev, st); }
int run_tde(const mlrn_store* s, double* out_dev, cudaStream_t st) { return launch_pass<PASS_TDE>(s, nullptr, out_dev, st); }
int run_physics(const mlrn_store* s, const double* z_dev, double* out_dev, cudaStream_t st) {
  return launch_pass<PASS_PHYSICS>(s, z_dev, out_dev, st);
}

}  // namespace mlrn
