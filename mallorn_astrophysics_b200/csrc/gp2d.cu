// gp2d.cu -- K5: the 2-D (time x wavelength) Matern-3/2 GP features, one CTA per object.
//
// Reference: src/features/multiband_gp.py:34-344 (george GP + scipy L-BFGS-B); semantics of the absent
// third-party package `george` as documented in oracle/gp2d.py (assumption A-george).  27 columns,
// schema.GP2D_COLUMNS.
//
// Per object (N valid observations, N >= 10) the CTA holds the symmetric N x N matrix in shared memory as
// lower-triangular 16 x 16 tiles (column-major inside a tile) and evaluates -ln L and its gradient with
// an FP64 *tile sweep*: for each pivot tile k
//     L = chol(A_kk)              (warp 0; ln|K| and the forward-substituted residual come from here, so
//                                  -ln L has Cholesky accuracy)
//     V_i = A_ik L^-T             (all other tile rows, column k read in place)
//     A_ij -= V_i V_j^T           (every other tile: rank-16 update, 2 x 4 register blocks per lane)
//     A_ik <- V_i L^-1, A_kk <- -(L L^T)^-1, b_i -= V_i w, b_k <- L^-T w      (w = L^-1 b_k)
// which leaves -K^-1 in place and alpha = K^-1 r in b (symmetric block Gauss-Jordan with an augmented
// right-hand side); the gradient 1/2 tr((alpha alpha^T - K^-1) dK/dtheta) is then an element-wise
// contraction with the kernel derivatives recomputed on the fly.  N^3/2 FMA per evaluation, all in
// shared memory; HBM traffic is one read of the object's observations and one 27-value row.
// The hyper-parameter search is lbfgs4.cuh run redundantly by every thread on broadcast (f, g).
// Objects with more than 16 * T_cap valid points (shared memory full) run in a second launch of the
// same kernel with the tiles in a per-CTA global scratch (L2-resident).
#include <cuda_runtime.h>
#include "store.cuh"
#include "objview.cuh"
#include "lbfgs4.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

#ifndef GP_THREADS_N
#define GP_THREADS_N 512
#endif
#ifndef GP_T_SMEM_MAX
#define GP_T_SMEM_MAX 12
#endif
constexpr int GP_THREADS = GP_THREADS_N;
constexpr int GP_WARPS = GP_THREADS / 32;
// Warp roles during a sweep step.  A scheduler partition owns the warps with the same (warp id % 4) and ONE FP64 pipe:
// a DMMA occupies it for 16 cycles, so every dependent DFMA of the serial pivot chain queued behind the three bulk warps
// of its partition (measured: 4.1 k cycles per pivot tile alone, 7.2-8.6 k beside them -- longer than the whole bulk
// phase).  The pivot warp therefore gets its partition to itself while it factors: the other warps of partition 0 run the
// small jobs and then wait for it, and the rank-16 tile updates go to the warps of partitions 1..3.
constexpr int PIV = 0;
static_assert(GP_WARPS >= 8 && GP_WARPS % 4 == 0, "the role split needs at least two warps per scheduler partition");
constexpr int N_BULK = GP_WARPS - GP_WARPS / 4;
constexpr int TB = 16;                 // tile edge
constexpr int TT = TB * TB;            // doubles per tile
constexpr int GP_COLS = 27;
constexpr int T_SMEM_MAX = GP_T_SMEM_MAX;         // 78 tiles + V buffer + vectors < 227 KB  (N <= 192 in shared memory)
constexpr int VLD = 20;                // k-row stride of a V tile: 8-word skew per row keeps the DMMA fragment loads conflict-free
constexpr int VT = TB * VLD;           // doubles per V tile
constexpr double GP_TINY = 1.25e-12;   // george's frozen white-noise floor
constexpr double LOG_2PI = 1.8378770664093453;
constexpr unsigned FULLMASK = 0xffffffffu;

__device__ __constant__ double c_wave[6] = {3670.0, 4825.0, 6222.0, 7545.0, 8691.0, 9710.0};

__host__ __device__ inline int n_tiles(int T) { return T * (T + 1) / 2; }
__host__ __device__ inline size_t gp_fixed_doubles(int np) {
  // vectors t, lam, y, noise, b (np each) + Lc + 2 x (Li, LiT, w) + reduction scratch + broadcast slots
  return 5 * (size_t)np + TT + 2 * (TT + VT + TB) + GP_WARPS * 4 + 16;
}
__host__ __device__ inline size_t gp_tile_doubles(int T) { return (size_t)n_tiles(T) * TT + (size_t)T * VT; }
__host__ inline size_t gp_smem_bytes(int T, bool tiles_in_smem) {
  size_t d = gp_fixed_doubles(T * TB);
  if (tiles_in_smem) d += (size_t)n_tiles(T) * TT + (size_t)T * VT;      // tiles and V buffer; otherwise both live in the global scratch
  return d * sizeof(double) + 2 * n_tiles(T) + 64;      // + (I, J) byte table of the tiles
}
// Longest object the GP handles: 192 tile rows (the (I, J) tables are bytes and the five per-point vectors have to fit in
// shared memory).  The reference has no such limit in principle (george would factor any N x N matrix); above it the 27
// columns are NaN and status[2] = GP_TOO_LONG.
constexpr int GP_T_MAX = 192;
constexpr int GP_TOO_LONG = -2;

struct GpShared {
  double *tiles, *vbuf;                 // [n_tiles][256], [T][VT]
  double *st, *sl, *sy, *sn, *sb;       // [np]
  double *Lc, *Li, *LiT, *w;              // pivot factor; its inverse row-major / transposed (stride VLD); L^-1 b_k
  double *red, *bc;                     // reduction scratch, broadcast slots
  unsigned char *tI, *tJ;
  int* flag;                            // = k+1 once pivot tile k+1 of the current evaluation is factored
};

__device__ __forceinline__ double* tile_ptr(const GpShared& S, int I, int J) { return S.tiles + (size_t)(I * (I + 1) / 2 + J) * TT; }

// deterministic block-wide sum of 4 values per thread; every thread gets the totals
__device__ void block_sum4(const GpShared& S, double* v) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(FULLMASK, v[q], o);
  __syncthreads();
  if (lane == 0) for (int q = 0; q < 4; ++q) S.red[warp * 4 + q] = v[q];
  __syncthreads();
  for (int q = 0; q < 4; ++q) { double s = 0.0; for (int w = 0; w < GP_WARPS; ++w) s += S.red[w * 4 + q]; v[q] = s; }
}

struct KernelPars { double mean, amp, im0, im1; };
__device__ __forceinline__ KernelPars unpack(const double* p) {
  KernelPars k; k.mean = p[0]; k.amp = 2.0 * exp(p[1]); k.im0 = exp(-p[2]); k.im1 = exp(-p[3]); return k;   // amp = ndim * exp(log_constant)
}

// Tile element (r, c).  A tile is stored in the FRAGMENT ORDER of the FP64 tensor-core accumulator (mma.m8n8k4: lane l
// of a warp holds C[l/4][2(l%4)] and C[l/4][2(l%4)+1] of an 8x8 quadrant): quadrant (r/8, c/8), then lane, then the pair.
// The rank-16 update -- 45 of the 63 tile products of a sweep step -- then reads and writes its accumulators as four
// conflict-free 128-bit accesses per lane; the column-major layout of round 1 cost eight 64-bit accesses with a 2-way
// bank conflict each, and the sweep is bound by shared-memory bandwidth (8 KB of traffic per 4096-FMA tile product).
// On top of that, the two halves (columns 0-3 / 4-7 of an 8x8 quadrant) of rows 2, 3, 6, 7 are exchanged (GP_SWZ): rows are
// 8 doubles = 16 banks apart, so without the exchange rows r and r + 2 of a quadrant fall on the same banks, and every
// access that walks DOWN a column -- the A-operand fragments of tile_scale, the pivot warp's row loads, the transposed
// stores of tile_finalise -- was a 4-way bank conflict (ncu: 17.5 % of the kernel's shared-memory wavefronts excessive, all
// in these three places).  With the exchange the first two are conflict-free and the third is 2-way; the accumulator
// accesses stay conflict-free (a lane reads its pair at lane ^ ((lane >> 2) & 2), see frag_lane()).
#ifndef GP_NO_SWZ
#define GP_SWZ(r) ((((r) >> 1) & 1) << 2)
#else
#define GP_SWZ(r) 0
#endif
__device__ __forceinline__ int tix(int r, int c) {
  return ((((r >> 3) * 2 + (c >> 3)) * 32 + (r & 7) * 4 + (((c & 7) ^ GP_SWZ(r)) >> 1)) << 1) | (c & 1);
}
// inverse: storage index -> (r, c); consecutive threads then touch consecutive words
__device__ __forceinline__ void tix_decode(int idx, int& r, int& c) {
  r = ((idx >> 7) << 3) | ((idx >> 3) & 7);
  c = (((idx >> 6) & 1) << 3) | ((((((idx >> 1) & 3) << 1) | (idx & 1))) ^ GP_SWZ(r));
}
// the double2 slot (within a 32-slot quadrant) that holds lane's accumulator pair C[lane/4][2(lane%4) .. +1]
__device__ __forceinline__ int frag_lane() {
  const int lane = threadIdx.x & 31;
#ifndef GP_NO_SWZ
  return lane ^ ((lane >> 2) & 2);
#else
  return lane;
#endif
}

// running product of the pivots as mantissa * 2^exponent (ln|K| is taken once per evaluation, off the critical path)
struct LogDet {
  double m; int e;
  __device__ __forceinline__ void mul(double x) { int ex; m *= frexp(x, &ex); e += ex; m = frexp(m, &ex); e += ex; }
  __device__ __forceinline__ double value() const { return log(m) + (double)e * 0.6931471805599453; }
};

// Pivot tile k, one warp.  Lanes 0..15 hold one row of A_kk each and turn it into L = chol(A_kk) (right-
// looking, the column being eliminated is broadcast by shuffle); lanes 16..31 hold one column of L^-1 each and
// ride on the SAME broadcasts: forward substitution of column j needs exactly L[c2][c] x_c, the term the
// Cholesky update multiplies by L[r][c].  The right-hand side b_k is carried as one more value per row lane,
// so w = L^-1 b_k also falls out of the same loop.  Results in S.Li / S.LiT / S.w [buf].
// (A square-root-free LDL^T form with the reciprocal off the dependency chain was measured in round 2: 5.5 k cycles
// per tile against 4.1 k for this one, both running alone -- the chain is not what bounds it.)
__device__ __forceinline__ void pivot_warp(const GpShared& S, int k, int buf, double& quad, LogDet& logdet, int& fail) {
  const int lane = threadIdx.x & 31, r = lane & 15;
  const bool lower = lane < TB;
  const double* A = tile_ptr(S, k, k);
  double val[TB];
#pragma unroll
  for (int c = 0; c < TB; ++c) val[c] = lower ? A[tix(r, c)] : ((c == r) ? 1.0 : 0.0);
  double bw = lower ? S.sb[k * TB + r] : 0.0;      // rhs of this row; becomes w_r at step r
  double wr = 0.0, prod = 1.0; bool bad = false;
  // (Round 2, measured and rejected: forcing the two shuffles on the recurrence -- next column's diagonal update and its
  // broadcast -- ahead of the step's ~28 other shuffles through an ordering dependency: 117.6 k objects/s against 123.4 k.)
#pragma unroll
  for (int c = 0; c < TB; ++c) {
    const double piv = __shfl_sync(FULLMASK, val[c], c);
    bad = bad || !(piv > 0.0);
    const double ri = rsqrt(piv);
    prod *= piv;
    const double mult = val[c] * ri;       // L[r][c] on the lower lanes, (L^-1)[c][j] on the upper lanes
    val[c] = mult;
    const double wc = __shfl_sync(FULLMASK, bw, c) * ri;       // w_c = (b_c - sum_{m<c} L[c][m] w_m) / L[c][c]
    if (lane == c) wr = wc;
    bw -= mult * wc;                       // rows below c (other lanes: harmless)
#pragma unroll
    for (int c2 = c + 1; c2 < TB; ++c2) val[c2] -= mult * __shfl_sync(FULLMASK, mult, c2);
  }
  if (!lower) {
#pragma unroll
    for (int q = 0; q < TB; ++q) { S.Li[buf * TT + q * TB + r] = val[q]; S.LiT[buf * VT + r * VLD + q] = val[q]; }   // Li[q][j], j = r
  } else {
    S.w[buf * TB + r] = wr;
  }
  double wq = lower ? wr * wr : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wq += __shfl_xor_sync(FULLMASK, wq, o);
  quad += wq; logdet.mul(prod);
  if (bad || !(prod == prod)) fail = 1;
  __syncwarp();
}

// ---- warp-level 16x16x16 FP64 products on the FP64 tensor-core path (mma.sync m8n8k4, "DMMA"):
// per k-step of 4 a lane loads ONE double per operand fragment instead of the 6 a register-blocked FMA
// loop needs, which takes the shared-memory pipe off the critical path.  Fragment layout (PTX ISA):
//   A (8x4, row)  lane l holds A[l/4][l%4]     B (4x8, col)  lane l holds B[l%4][l/4]
//   C/D (8x8)     lane l holds C[l/4][2*(l%4)], C[l/4][2*(l%4)+1]
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// acc[ri][ci][0..1] += sum_m A(8ri + l/4, m) * B(m, 8ci + 2(l%4) + {0,1}) over m = 0..15, where the operands
// are given K-MAJOR: Ak[m*lda + row], Bk[m*ldb + col]
__device__ __forceinline__ void dmma_16x16x16(double (&acc)[2][2][2], const double* __restrict__ Ak, int lda,
                                              const double* __restrict__ Bk, int ldb, bool negate_a) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int m = 4 * ks + lk;
    double a0 = Ak[m * lda + lr], a1 = Ak[m * lda + 8 + lr];
    const double b0 = Bk[m * ldb + lr], b1 = Bk[m * ldb + 8 + lr];
    if (negate_a) { a0 = -a0; a1 = -a1; }
    dmma(acc[0][0][0], acc[0][0][1], a0, b0);
    dmma(acc[0][1][0], acc[0][1][1], a0, b1);
    dmma(acc[1][0][0], acc[1][0][1], a1, b0);
    dmma(acc[1][1][0], acc[1][1][1], a1, b1);
  }
}

// C(i, j) -= Vi Vj^T        (C in fragment order, V tiles k-major with stride VLD)
__device__ __forceinline__ void tile_update(double* __restrict__ C, const double* __restrict__ Vi, const double* __restrict__ Vj) {
  const int lane = frag_lane();
  double2* __restrict__ C2 = reinterpret_cast<double2*>(C);
  double acc[2][2][2];
#pragma unroll
  for (int ri = 0; ri < 2; ++ri)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) { const double2 v = C2[(ri * 2 + ci) * 32 + lane]; acc[ri][ci][0] = v.x; acc[ri][ci][1] = v.y; }
  // (measured and rejected: loading the fragments once when Vi == Vj, the diagonal tiles: 128.5 k against 130.7 k objects/s)
  dmma_16x16x16(acc, Vi, VLD, Vj, VLD, true);
#pragma unroll
  for (int ri = 0; ri < 2; ++ri)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) C2[(ri * 2 + ci) * 32 + lane] = make_double2(acc[ri][ci][0], acc[ri][ci][1]);
}

// V = X Li^T, X = block (i, k) of the symmetric matrix: tile(i, k) if i > k, tile(k, i)^T if i < k.
// V[r][c] = sum_m X[r][m] Li[c][m]: A k-major = X[m][r] (column m of X), B k-major = LiT[m*VLD + c].
__device__ __forceinline__ void tile_scale(double* __restrict__ V, const double* __restrict__ X, bool transposed,
                                           const double* __restrict__ LiT) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lk = lane & 3, lc = 2 * lk;
  double acc[2][2][2] = {};
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int m = 4 * ks + lk;
    double a0, a1;
    if (!transposed) { a0 = X[tix(lr, m)]; a1 = X[tix(8 + lr, m)]; }          // element (r, m) of tile(i, k)
    else { a0 = X[tix(m, lr)]; a1 = X[tix(m, 8 + lr)]; }                       // element (m, r) of tile(k, i)
    const double b0 = LiT[m * VLD + lr], b1 = LiT[m * VLD + 8 + lr];
    dmma(acc[0][0][0], acc[0][0][1], a0, b0);
    dmma(acc[0][1][0], acc[0][1][1], a0, b1);
    dmma(acc[1][0][0], acc[1][0][1], a1, b0);
    dmma(acc[1][1][0], acc[1][1][1], a1, b1);
  }
#pragma unroll
  for (int ri = 0; ri < 2; ++ri)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
      for (int e = 0; e < 2; ++e) V[(8 * ci + lc + e) * VLD + 8 * ri + lr] = acc[ri][ci][e];
}

// block (i, k) <- V Li, written to tile(i, k) or transposed into tile(k, i).
// Z[r][c] = sum_m V[r][m] Li[m][c]: A k-major = V[m*VLD + r], B k-major = Li[m*16 + c] (row-major Li).
__device__ __forceinline__ void tile_finalise(double* __restrict__ Tl, bool transposed, const double* __restrict__ V,
                                              const double* __restrict__ Li) {
  const int lane = threadIdx.x & 31, lr = lane >> 2, lc = 2 * (lane & 3);
  double acc[2][2][2] = {};
  dmma_16x16x16(acc, V, VLD, Li, TB, false);
  if (!transposed) {
    double2* T2 = reinterpret_cast<double2*>(Tl);
#pragma unroll
    for (int ri = 0; ri < 2; ++ri)
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) T2[(ri * 2 + ci) * 32 + frag_lane()] = make_double2(acc[ri][ci][0], acc[ri][ci][1]);
  } else {
#pragma unroll
    for (int ri = 0; ri < 2; ++ri)
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int e = 0; e < 2; ++e) Tl[tix(8 * ci + lc + e, 8 * ri + lr)] = acc[ri][ci][e];
  }
}

// One bulk task of sweep step k: a rank-16 update of a tile outside row/column k, or column k <- V_i Li.
__device__ __forceinline__ void bulk_task(const GpShared& S, int task, int k, int nupd, const double* __restrict__ Li) {
  if (task < nupd) {
    const int ii = S.tI[task], jj = S.tJ[task];
    const int i = ii + (ii >= k), j = jj + (jj >= k);
    if (i == k + 1 && j == k + 1) return;             // the pivot warp's
    tile_update(tile_ptr(S, i, j), S.vbuf + i * VT, S.vbuf + j * VT);
  } else {
    const int q = task - nupd, i = q + (q >= k);
    if (i > k) tile_finalise(tile_ptr(S, i, k), false, S.vbuf + i * VT, Li);
    else tile_finalise(tile_ptr(S, k, i), true, S.vbuf + i * VT, Li);
  }
}

// -ln L and its gradient at p.  Leaves alpha in S.sb[0..n).  Result broadcast to every thread.
// cache: per-CTA global (L2-resident) scratch [2][nt * 256] holding s and exp(-s) of every matrix element
// between the build and the gradient contraction.
#ifdef GP_PHASE_TIMING
__device__ unsigned long long g_gp_phase[12];     // developer build only: cycles per phase summed over CTA 0's evaluations
#define GP_TICK(slot) do { if (blockIdx.x == 0 && threadIdx.x == PIV * 32) { const long long now_ = clock64(); atomicAdd(&g_gp_phase[slot], (unsigned long long)(now_ - t_mark_)); t_mark_ = now_; } } while (0)
// per-warp split of the sweep (lane 0 of every warp of CTA 0): 0 task acquire, 1 tile updates, 2 column finalise, 3 small jobs,
// 4 wait at the end-of-step barrier, 5 number of tile updates, 6 pivot-warp work (update + b + factor), 7 scale phase + its barrier
__device__ unsigned long long g_gp_warp[GP_WARPS][8];
#define GPW_ADD(kind) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { const long long n_ = clock64(); atomicAdd(&g_gp_warp[threadIdx.x >> 5][kind], (unsigned long long)(n_ - gw_t_)); gw_t_ = n_; } } while (0)
#define GPW_COUNT(kind) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_gp_warp[threadIdx.x >> 5][kind], 1ULL); } while (0)
#else
#define GP_TICK(slot) do { } while (0)
#define GPW_ADD(kind) do { } while (0)
#define GPW_COUNT(kind) do { } while (0)
#endif

__device__ __forceinline__ void gp_eval(const GpShared& S, double* __restrict__ cache, int n, int T, int T_cap,
                                        const double* p, double& f, double* grad) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef GP_PHASE_TIMING
  long long t_mark_ = clock64();
#endif
  const KernelPars kp = unpack(p);
  const int nt = n_tiles(T);
  double* cache_s = cache;
  double* cache_e = cache + (size_t)n_tiles(T_cap) * TT;
  // ---- (a) kernel matrix, b = y - mean
  // e = storage index: tile, then fragment order.  Several elements per trip: the sqrt and exp chains of
  // the pair interleave (the loop is bound by FP64 latency, not by its 4 shared-memory loads).
  auto element = [&](int e) {
    const int tl = e >> 8;
    int r, c;
    tix_decode(e & 255, r, c);
    const int gi = S.tI[tl] * TB + r, gj = S.tJ[tl] * TB + c;
    const bool ok = gi < n && gj < n;
    if (!ok) { S.tiles[e] = (gi == gj) ? 1.0 : 0.0; return; }      // padding of the last tile row / column (12 % of the elements at N = 150): no kernel value to compute (+2 %)
    // (measured and rejected: computing only the lower half of the diagonal tiles and mirroring it: 125.4 k against 130.5 k objects/s)
    const double dt = S.st[gi] - S.st[gj], dl = S.sl[gi] - S.sl[gj];
    const double s = sqrt(3.0 * (dt * dt * kp.im0 + dl * dl * kp.im1));
    const double ex = exp(-s);
    double v = kp.amp * (1.0 + s) * ex;
    if (gi == gj) v += S.sn[gi];
    cache_s[e] = s; cache_e[e] = ex;
    S.tiles[e] = v;
  };
  {
    int e = tid;
#ifndef GP_BUILD_ILP2      // four elements per trip: 121.7 k -> 123.1 k objects/s at 100,000 objects (same sums in the same order)
    for (; e + 3 * GP_THREADS < nt * TT; e += 4 * GP_THREADS) { element(e); element(e + GP_THREADS); element(e + 2 * GP_THREADS); element(e + 3 * GP_THREADS); }
#endif
    for (; e + GP_THREADS < nt * TT; e += 2 * GP_THREADS) { element(e); element(e + GP_THREADS); }
    if (e < nt * TT) element(e);
  }
  for (int i = tid; i < T * TB; i += GP_THREADS) S.sb[i] = (i < n) ? S.sy[i] - kp.mean : 0.0;
  double quad = 0.0; LogDet logdet{1.0, 0}; int fail = 0;     // accumulated by the pivot warp (uniform across its lanes)
  __syncthreads();
  GP_TICK(0);
#ifdef GP_PHASE_TIMING
  if (blockIdx.x == 0 && tid == PIV * 32) { atomicAdd(&g_gp_phase[10], 1ULL); atomicAdd(&g_gp_phase[11], (unsigned long long)T); }
#endif
  // ---- (b) tile sweep with look-ahead: warp 0 owns the critical path (next pivot tile), warps 1.. the bulk
  if (warp == PIV) pivot_warp(S, 0, 0, quad, logdet, fail);
  __syncthreads();
  GP_TICK(1);
#ifdef GP_PHASE_TIMING
  long long gw_t_ = clock64();
#endif
  if (tid == 0) *S.flag = 0;                           // (published by the barrier above the first step's helpers)
  for (int k = 0; k < T; ++k) {
    const int cur = k & 1, nxt = cur ^ 1;
    const double* Li = S.Li + cur * TT;
    const double* LiT = S.LiT + cur * VT;
    const double* wv = S.w + cur * TB;
    const bool has_next = k + 1 < T;
    const int Tm = T - 1, nupd = Tm * (Tm + 1) / 2, ntask = nupd + Tm;
    // Bulk tasks (rank-16 tile updates outside row/column k, then column k <- V_i Li) are dealt round-robin to the warps of
    // partitions 1..3; the tail of the last, partly filled round goes to the other warps of partition 0, which run it once
    // the next pivot tile is factored.
    // (only when that saves the bulk warps a round: otherwise the helpers' late start would be the end of the step.
    // Also measured in round 2 and rejected: letting those warps start on bulk tasks at once -- the pivot chain slows
    // from 4.5 k to 5.1 k cycles, 118.6 k objects/s against 121.7 k; and scaling the next step's V inside this phase behind
    // flags instead of the scale barrier -- the scales then wait for the pivot chain, 104.9 k objects/s.)
    // (Round 2, final build, measured and rejected: the spare warps taking TWO tasks each from the start whenever that saves
    // the bulk warps a round (54 tasks = 4 x 12 + 6 at T = 10): the pivot slows from 4.2 k to 5.0 k cycles and the bulk tasks
    // from 1.18 k to 1.37 k each -- the phase is bound by shared-memory bandwidth, not by who issues it; 126.5 k against
    // 130.6 k objects/s.)
    const int n_help = (has_next && (ntask % N_BULK) >= 1 && (ntask % N_BULK) <= GP_WARPS / 4 - 1) ? (ntask % N_BULK) : 0;
    // ---- phase A: V_i = X_i Li^T for i != k, one warp per tile row.  The pivot warp takes row k+1 and then runs AHEAD
    // of the barrier (it only arrives): the update of tile (k+1, k+1), of b_{k+1} and the factorisation need V_{k+1} only.
    if (warp == PIV) {
      if (has_next) tile_scale(S.vbuf + (k + 1) * VT, tile_ptr(S, k + 1, k), false, LiT);
      __threadfence_block();
      asm volatile("bar.arrive 1, %0;" ::"r"(GP_THREADS) : "memory");
      GP_TICK(2);
      GPW_ADD(7);
      if (has_next) {
        const double* V = S.vbuf + (k + 1) * VT;
        __syncwarp();
        tile_update(tile_ptr(S, k + 1, k + 1), V, V);
        {                                             // b_{k+1} -= V_{k+1} w_k, the 16-term sums split over the two half-warps
          const int rr = lane & 15, m0 = (lane >> 4) * 8;
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < 8; ++m) s += V[(m0 + m) * VLD + rr] * wv[m0 + m];
          s += __shfl_xor_sync(FULLMASK, s, 16);
          if (lane < TB) S.sb[(k + 1) * TB + lane] -= s;
        }
        __syncwarp();
        GP_TICK(3);
        pivot_warp(S, k + 1, nxt, quad, logdet, fail);
        __threadfence_block();
        if (lane == 0) *(volatile int*)S.flag = k + 1;
        GP_TICK(6);
        GPW_ADD(6);
      }
    } else {
      for (int q = warp - 1; q < T - 1; q += GP_WARPS - 1) {
        if (q == k) continue;                         // row k+1: the pivot warp's
        const int i = q + (q >= k);
        if (i > k) tile_scale(S.vbuf + i * VT, tile_ptr(S, i, k), false, LiT);
        else tile_scale(S.vbuf + i * VT, tile_ptr(S, k, i), true, LiT);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(GP_THREADS) : "memory");
      GPW_ADD(7);
    }
    // ---- phase B
    if ((warp & 3) != 0) {
      const int bulk_id = (warp >> 2) * 3 + (warp & 3) - 1;
      // (Round 2, measured and rejected: dealing a last round that fills at most half of the bulk warps as HALF tasks, two
      // warps per tile -- 4.5 rounds instead of 5 at T = 10: 126.4 k objects/s against 128.4 k.)
      for (int task = bulk_id; task < ntask - n_help; task += N_BULK) bulk_task(S, task, k, nupd, Li);
      GPW_ADD(1);
    } else {
      if (warp == 4) {
        // pivot tile <- -(Li^T Li) as one 16x16x16 product; b_k <- Li^T w (nobody reads b_k or w_k any more in this
        // step); b_i -= V_i w for the rows outside tiles k and k+1
        double acc[2][2][2] = {};
        dmma_16x16x16(acc, Li, TB, Li, TB, false);
        double2* P2 = reinterpret_cast<double2*>(tile_ptr(S, k, k));
#pragma unroll
        for (int ri = 0; ri < 2; ++ri)
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) P2[(ri * 2 + ci) * 32 + frag_lane()] = make_double2(-acc[ri][ci][0], -acc[ri][ci][1]);
        if (lane < TB) {
          double s = 0.0;
          for (int m = lane; m < TB; ++m) s += Li[m * TB + lane] * wv[m];
          S.sb[k * TB + lane] = s;
        }
        double wreg[TB];
#pragma unroll
        for (int m = 0; m < TB; ++m) wreg[m] = wv[m];
        for (int i = lane; i < T * TB; i += 32) {
          const int ti = i >> 4, r = i & 15;
          if (ti != k && ti != k + 1) {
            const double* V = S.vbuf + ti * VT;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int m = 0; m < TB; m += 2) { s0 += V[m * VLD + r] * wreg[m]; s1 += V[(m + 1) * VLD + r] * wreg[m + 1]; }
            S.sb[i] -= s0 + s1;
          }
        }
        GPW_ADD(3);
      }
      const int h = (warp >> 2) - 1;                  // the pivot warp (-1) is the critical path: it takes none
      if (h >= 0 && h < n_help) {
        if (lane == 0) while (*(volatile int*)S.flag != k + 1) __nanosleep(64);
        __syncwarp();
        GPW_ADD(0);
        bulk_task(S, ntask - n_help + h, k, nupd, Li);
        GPW_ADD(2);
      }
    }
    __syncthreads();
    GP_TICK(4);
    GPW_ADD(4);
  }
  __syncthreads();
  // ---- (c) gradient contraction: 1/2 sum (alpha_i alpha_j - Kinv_ij) dK_ij ; tiles hold -Kinv
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = tid; i < n; i += GP_THREADS) acc[0] += S.sb[i];
  // two elements per trip with the (L2-resident) cache loads of both issued first
  auto contract = [&](int e, double s, double ex) {
    const int tl = e >> 8;
    int r, c;
    tix_decode(e & 255, r, c);
    const int I = S.tI[tl], J = S.tJ[tl];
    const int gi = I * TB + r, gj = J * TB + c;
    if (gi < n && gj < n) {
      const double wgt = (I == J) ? 1.0 : 2.0;
      const double Aij = wgt * (S.sb[gi] * S.sb[gj] + S.tiles[e]);
      const double dt = S.st[gi] - S.st[gj], dl = S.sl[gi] - S.sl[gj];
      const double a0 = dt * dt * kp.im0, a1 = dl * dl * kp.im1;
      const double dk = 1.5 * kp.amp * ex;
      acc[1] += Aij * kp.amp * (1.0 + s) * ex; acc[2] += Aij * dk * a0; acc[3] += Aij * dk * a1;
    }
  };
  {
    int e = tid;
#ifndef GP_BUILD_ILP2      // four elements per trip: 121.7 k -> 123.1 k objects/s at 100,000 objects (same sums in the same order)
    for (; e + 3 * GP_THREADS < nt * TT; e += 4 * GP_THREADS) {
      const double s0 = cache_s[e], x0 = cache_e[e], s1 = cache_s[e + GP_THREADS], x1 = cache_e[e + GP_THREADS];
      const double s2 = cache_s[e + 2 * GP_THREADS], x2 = cache_e[e + 2 * GP_THREADS], s3 = cache_s[e + 3 * GP_THREADS], x3 = cache_e[e + 3 * GP_THREADS];
      contract(e, s0, x0); contract(e + GP_THREADS, s1, x1); contract(e + 2 * GP_THREADS, s2, x2); contract(e + 3 * GP_THREADS, s3, x3);
    }
#endif
    for (; e + GP_THREADS < nt * TT; e += 2 * GP_THREADS) {
      const double s0 = cache_s[e], x0 = cache_e[e], s1 = cache_s[e + GP_THREADS], x1 = cache_e[e + GP_THREADS];
      contract(e, s0, x0); contract(e + GP_THREADS, s1, x1);
    }
    if (e < nt * TT) contract(e, cache_s[e], cache_e[e]);
  }
  GP_TICK(5);
  if (tid == PIV * 32) { S.bc[0] = quad; S.bc[1] = logdet.value(); S.bc[2] = (double)fail; }
  block_sum4(S, acc);                                 // contains the barriers that publish bc[]
  const double ll = -0.5 * S.bc[0] - 0.5 * S.bc[1] - 0.5 * (double)n * LOG_2PI;
  if (S.bc[2] != 0.0 || !is_finite(ll)) {             // multiband_gp.py:141-154: 1e25 and a zero gradient
    f = 1e25; grad[0] = grad[1] = grad[2] = grad[3] = 0.0;
  } else {
    f = -ll; grad[0] = -acc[0]; grad[1] = -0.5 * acc[1]; grad[2] = -0.5 * acc[2]; grad[3] = -0.5 * acc[3];
  }
  __syncthreads();
}

template <bool SMEM_TILES>
__global__ void __launch_bounds__(GP_THREADS, (GP_THREADS <= 256) ? 2 : 1)
k_gp2d(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
       const double* __restrict__ ge, const unsigned short* __restrict__ grank, int n_obj, int T_cap, int n_lo,
       int n_hi, double* __restrict__ gscratch, double* __restrict__ gcache, unsigned long long* __restrict__ ticket, double* __restrict__ out,
       int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long s_obj;
  __shared__ int s_cnt[GP_WARPS + 1];
  __shared__ int s_pk;
  __shared__ int s_go;
  __shared__ Lbfgs4 s_opt;              // optimiser state: advanced by thread 0, x broadcast to the CTA
  __shared__ int s_flag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np_cap = T_cap * TB;
  GpShared S;
  {
    double* d = (double*)smem_raw;
    S.st = d; d += np_cap; S.sl = d; d += np_cap; S.sy = d; d += np_cap; S.sn = d; d += np_cap; S.sb = d; d += np_cap;
    S.Lc = d; d += TT;
    S.Li = d; d += 2 * TT; S.LiT = d; d += 2 * VT; S.w = d; d += 2 * TB;      // double-buffered (look-ahead)
    S.red = d; d += GP_WARPS * 4; S.bc = d; d += 16;
    if (SMEM_TILES) { S.vbuf = d; d += (size_t)T_cap * VT; S.tiles = d; d += (size_t)n_tiles(T_cap) * TT; }
    else {
      S.tiles = gscratch + (size_t)blockIdx.x * gp_tile_doubles(T_cap);
      S.vbuf = S.tiles + (size_t)n_tiles(T_cap) * TT;
    }
    S.tI = (unsigned char*)d; S.tJ = S.tI + n_tiles(T_cap);
    S.flag = &s_flag;
    // address-space hints: everything carved from the dynamic shared array is shared memory
    __builtin_assume(__isShared(S.st)); __builtin_assume(__isShared(S.sl)); __builtin_assume(__isShared(S.sy));
    __builtin_assume(__isShared(S.sn)); __builtin_assume(__isShared(S.sb)); __builtin_assume(__isShared(S.Lc));
    __builtin_assume(__isShared(S.Li));
    __builtin_assume(__isShared(S.LiT));
    __builtin_assume(__isShared(S.w));
    __builtin_assume(__isShared(S.red)); __builtin_assume(__isShared(S.bc));
    __builtin_assume(__isShared(S.tI)); __builtin_assume(__isShared(S.tJ));
    if (SMEM_TILES) { __builtin_assume(__isShared(S.vbuf)); __builtin_assume(__isShared(S.tiles)); }
  }
  for (int tl = tid; tl < n_tiles(T_cap); tl += GP_THREADS) {
    int I = (int)((sqrt(8.0 * tl + 1.0) - 1.0) * 0.5);
    while (I * (I + 1) / 2 > tl) --I;
    while ((I + 1) * (I + 2) / 2 <= tl) ++I;
    S.tI[tl] = (unsigned char)I; S.tJ[tl] = (unsigned char)(tl - I * (I + 1) / 2);
  }
  for (;;) {
    __syncthreads();
    if (tid == 0) s_obj = (long long)atomicAdd(ticket, 1ULL);
    __syncthreads();
    const long long o = s_obj;
    if (o >= n_obj) break;
    const long long* so = seg_off + o * N_SEG;
    const long long base = so[0];
    const int ntot = (int)(so[N_SEG] - base);
    int off[N_SEG + 1];
#pragma unroll
    for (int b = 0; b <= N_SEG; ++b) off[b] = (int)(so[b] - base);
    // ---- valid rows (multiband_gp.py:49-63), ordered compaction into shared memory
    int n = 0;
    double tmin_all = pinf();
    for (int chunk = 0; chunk < ntot; chunk += GP_THREADS) {
      const int i = chunk + tid;
      bool keep = false; double t = 0.0, f = 0.0, e = 0.0; int b = 0;
      if (i < ntot) {
        t = gt[base + i]; f = gf[base + i]; e = ge[base + i];
        while (b < N_SEG - 1 && i >= off[b + 1]) ++b;
        tmin_all = fmin(tmin_all, t);
        keep = (b < N_BAND) && !is_nan(f) && !is_nan(e) && (e > 0.0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_cnt[warp] = __popc(m);
      __syncthreads();
      int before = n;
      for (int w = 0; w < warp; ++w) before += s_cnt[w];
      int total = 0;
      for (int w = 0; w < GP_WARPS; ++w) total += s_cnt[w];
      const int slot = before + __popc(m & ((1u << lane) - 1u));
      if (keep && slot < np_cap) { S.st[slot] = t; S.sl[slot] = c_wave[b]; S.sy[slot] = f; S.sn[slot] = e; }
      n += total;
      __syncthreads();
    }
    double* row = out + o * GP_COLS;
    const bool mine = (n >= n_lo && n <= n_hi);         // this launch's share (n < 10 rows belong to the first launch)
    if (!mine) continue;
    if (status && tid < 4) status[o * 4 + tid] = (tid == 3) ? n : 0;
    if (n < 10) { if (tid < GP_COLS) row[tid] = qnan(); continue; }       // :66
    if (n > np_cap) {                                   // beyond GP_T_MAX tile rows
      if (tid < GP_COLS) row[tid] = qnan();
      if (status && tid == 0) status[o * 4 + 2] = GP_TOO_LONG;
      continue;
    }
    // ---- t -= min(t); flux scale = median |f| over f != 0 (rank counting); y, noise; mean / var
    double red[4] = {0.0, 0.0, 0.0, 0.0};
    {
      double tm = pinf();
      for (int i = tid; i < n; i += GP_THREADS) tm = fmin(tm, S.st[i]);
      for (int o2 = 16; o2 > 0; o2 >>= 1) { tm = fmin(tm, __shfl_xor_sync(0xffffffffu, tm, o2)); tmin_all = fmin(tmin_all, __shfl_xor_sync(0xffffffffu, tmin_all, o2)); }
      __syncthreads();
      if (lane == 0) { S.red[warp * 4] = tm; S.red[warp * 4 + 1] = tmin_all; }
      __syncthreads();
      tm = pinf(); tmin_all = pinf();
      for (int w = 0; w < GP_WARPS; ++w) { tm = fmin(tm, S.red[w * 4]); tmin_all = fmin(tmin_all, S.red[w * 4 + 1]); }
      __syncthreads();
      for (int i = tid; i < n; i += GP_THREADS) S.st[i] -= tm;
    }
    int nz = 0;
    for (int i = 0; i < n; ++i) nz += (S.sy[i] != 0.0);               // uniform; n is small
    if (tid == 0) { S.bc[4] = qnan(); S.bc[5] = qnan(); }
    __syncthreads();
    for (int i = tid; i < n; i += GP_THREADS) {
      const double vi = fabs(S.sy[i]);
      if (S.sy[i] == 0.0) continue;
      int rk = 0;
      for (int j = 0; j < n; ++j) { const double vj = fabs(S.sy[j]); rk += (S.sy[j] != 0.0) && (vj < vi || (vj == vi && j < i)); }
      if (rk == (nz >> 1)) S.bc[4] = vi;
      if (rk == (nz >> 1) - 1) S.bc[5] = vi;
    }
    __syncthreads();
    double scale = (nz == 0) ? qnan() : ((nz & 1) ? S.bc[4] : 0.5 * (S.bc[4] + S.bc[5]));   // numpy.median
    if (scale == 0.0) scale = 1.0;
    __syncthreads();
    for (int i = tid; i < n; i += GP_THREADS) {
      const double y = S.sy[i] / scale, ye = S.sn[i] / scale;
      S.sy[i] = y; S.sn[i] = ye * ye + GP_TINY;
      red[0] += y;
    }
    block_sum4(S, red);
    const double mean = red[0] / (double)n;
    red[0] = red[1] = red[2] = red[3] = 0.0;
    for (int i = tid; i < n; i += GP_THREADS) { const double d = S.sy[i] - mean; red[0] += d * d; }
    block_sum4(S, red);
    const double var = red[0] / (double)n;
    const int T = (n + TB - 1) / TB;
    // ---- hyper-parameter search (multiband_gp.py:129-164)
    const double p0[4] = {mean, log(var / 2.0), log(100.0 * 100.0), log(6000.0 * 6000.0)};
    if (tid == 0) s_opt.start(p0);
    __syncthreads();
    double* cache = gcache + (size_t)blockIdx.x * 2 * n_tiles(T_cap) * TT;
    double f_fin = 0.0;
    for (int searching = 1;;) {
      double f, g[4];
      const double xk[4] = {s_opt.x[0], s_opt.x[1], s_opt.x[2], s_opt.x[3]};
      gp_eval(S, cache, n, T, T_cap, xk, f, g);        // ends with a barrier: everyone has read x
      if (!searching) { f_fin = f; break; }            // that was the evaluation at the optimum (alpha for predict)
#ifdef GP_PHASE_TIMING
      const long long adv_t0 = clock64();
#endif
      if (tid == 0) s_go = s_opt.advance(f, g) ? 1 : 0;
      __syncthreads();
#ifdef GP_PHASE_TIMING
      if (blockIdx.x == 0 && tid == 0) atomicAdd(&g_gp_phase[7], (unsigned long long)(clock64() - adv_t0));      // optimiser step (thread 0) + barrier
#endif
      if (!s_go) {
        // The search normally ends at the point it evaluated last (a successful line search): alpha of the optimum is
        // still in S.sb and the evaluation the reference repeats for predict() would reproduce it bit for bit -- skipped
        // (one evaluation in twenty).  Only after an abnormal end (x restored to an earlier iterate) is it run.
        if (s_opt.x[0] == xk[0] && s_opt.x[1] == xk[1] && s_opt.x[2] == xk[2] && s_opt.x[3] == xk[3]) { f_fin = f; break; }
        searching = 0;
      }
    }
    const double p[4] = {s_opt.x[0], s_opt.x[1], s_opt.x[2], s_opt.x[3]};
    const double f_opt = s_opt.f;
    if (status && tid == 0) { status[o * 4] = s_opt.nit; status[o * 4 + 1] = s_opt.nfev; status[o * 4 + 2] = s_opt.why; }
    // ---- features (multiband_gp.py:171-188): positions 0, 1, 2 of the parameter vector, read literally
    const double amp_f = exp(p[0]), ts_f = sqrt(exp(p[1])), ws_f = sqrt(exp(p[2]));
    if (tid < GP_COLS) row[tid] = qnan();
    __syncthreads();
    if (tid == 0) { row[0] = amp_f; row[1] = ts_f; row[2] = ws_f; row[3] = -f_opt; row[4] = ts_f / (ws_f / 1000.0); }
    if (f_fin >= 1e25) continue;
    // peak time: first flux maximum of the r band in FILE order, else over all rows (:331-338)
    if (warp == 0) {
      int lo = off[BR], hi = off[BR + 1];
      if (hi == lo) { lo = 0; hi = ntot; }
      double best = -pinf(); int brank = 0x7fffffff, bi = -1;
      for (int i = lo + lane; i < hi; i += 32) {
        const double f = gf[base + i];
        if (is_nan(f)) continue;
        const int rk = grank[base + i];
        if (f > best || (f == best && rk < brank)) { best = f; brank = rk; bi = i; }
      }
      for (int o2 = 16; o2 > 0; o2 >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o2);
        const int ork = __shfl_xor_sync(0xffffffffu, brank, o2), oi = __shfl_xor_sync(0xffffffffu, bi, o2);
        if (oi >= 0 && (bi < 0 || ob > best || (ob == best && ork < brank))) { best = ob; brank = ork; bi = oi; }
      }
      if (lane == 0) s_pk = bi;
    }
    __syncthreads();
    if (s_pk < 0) continue;
    const double pt = gt[base + s_pk] - tmin_all;
    const KernelPars kp = unpack(p);
    for (int pw = warp; pw < 12; pw += GP_WARPS) {     // 12 predictions, one warp each (:236-262)
      const int ei = pw / 3, j = pw % 3;
      const double ep = (ei == 0) ? 0.0 : (ei == 1 ? 20.0 : (ei == 2 ? 50.0 : 100.0));
      const double tq = pt + ep, lq = c_wave[1 + j];
      double acc = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double dt = tq - S.st[i], dl = lq - S.sl[i];
        const double s = sqrt(3.0 * (dt * dt * kp.im0 + dl * dl * kp.im1));
        acc += kp.amp * (1.0 + s) * exp(-s) * S.sb[i];
      }
      for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
      if (lane == 0) S.bc[pw] = (acc + kp.mean) * scale;
    }
    __syncthreads();
    if (tid == 0) {
      double col[4][2];
      for (int ei = 0; ei < 4; ++ei) {
        const double fg = S.bc[3 * ei], fr = S.bc[3 * ei + 1], fi = S.bc[3 * ei + 2];
        row[5 + 5 * ei] = fg; row[6 + 5 * ei] = fr; row[7 + 5 * ei] = fi;
        col[ei][0] = (fg > 0.0 && fr > 0.0) ? -2.5 * log10(fg / fr) : qnan();
        col[ei][1] = (fr > 0.0 && fi > 0.0) ? -2.5 * log10(fr / fi) : qnan();
        row[8 + 5 * ei] = col[ei][0]; row[9 + 5 * ei] = col[ei][1];
      }
      if (!is_nan(col[0][0]) && !is_nan(col[2][0])) row[25] = (col[2][0] - col[0][0]) / 50.0;     // :265-277
      if (!is_nan(col[0][0]) && !is_nan(col[3][0])) row[26] = (col[3][0] - col[0][0]) / 100.0;
    }
  }
}

}  // namespace

#ifdef GP_PHASE_TIMING
extern "C" void mlrn_gp_warp_cycles(unsigned long long* out, int reset) {       // [GP_WARPS][8]
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_gp_warp, sizeof(unsigned long long) * GP_WARPS * 8);
  if (reset) { unsigned long long z[GP_WARPS * 8] = {0}; cudaMemcpyToSymbol(g_gp_warp, z, sizeof(z)); }
}
extern "C" void mlrn_gp_phase_cycles(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_gp_phase, sizeof(unsigned long long) * 12);
  if (reset) { unsigned long long z[12] = {0}; cudaMemcpyToSymbol(g_gp_phase, z, sizeof(z)); }
}
#endif

int run_gp2d(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  int T_need = (s->max_obs + TB - 1) / TB > 1 ? (s->max_obs + TB - 1) / TB : 1;
  const bool any_beyond = T_need > T_SMEM_MAX;
  if (T_need > GP_T_MAX) T_need = GP_T_MAX;
  const int T1 = T_need < T_SMEM_MAX ? T_need : T_SMEM_MAX;
  const size_t smem1 = gp_smem_bytes(T1, true), smem2 = gp_smem_bytes(T_need, false);
  MLRN_CUDA(cudaFuncSetAttribute(k_gp2d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  MLRN_CUDA(cudaFuncSetAttribute(k_gp2d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  int occ1 = 1, occ2 = 1;
  MLRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k_gp2d<true>, GP_THREADS, smem1));
  MLRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_gp2d<false>, GP_THREADS, smem2));
  if (occ1 < 1) occ1 = 1;
  if (occ2 < 1) occ2 = 1;
  auto grid_for = [&](int occ) { long long g = (long long)s->sm_count * occ; return (int)(s->n_obj < g ? s->n_obj : g); };
  const int grid1 = grid_for(occ1);
  int grid2 = 0;
  const size_t cache1 = sizeof(double) * 2 * (size_t)n_tiles(T1) * TT, cache2 = sizeof(double) * 2 * (size_t)n_tiles(T_need) * TT;
  const size_t scratch2 = sizeof(double) * gp_tile_doubles(T_need);
  if (any_beyond) {                                     // objects beyond the shared-memory capacity: tiles in an L2-resident scratch
    grid2 = grid_for(occ2);
    const long long fit = (long long)((6ULL << 30) / (cache2 + scratch2));     // at most 6 GB of per-CTA scratch (a 3,000-point object needs 109 MB)
    if (grid2 > fit) grid2 = (int)(fit > 1 ? fit : 1);
  }
  const size_t cache_bytes = cache1 * grid1 > cache2 * grid2 ? cache1 * grid1 : cache2 * grid2;
  double* cache = nullptr;                              // s / exp(-s) of the current evaluation, per CTA (L2-resident)
  MLRN_CUDA(cudaMallocAsync(&cache, cache_bytes, st));
  MLRN_CUDA(cudaMemsetAsync(s->ticket + 2, 0, 2 * sizeof(unsigned long long), st));
  k_gp2d<true><<<grid1, GP_THREADS, smem1, st>>>(s->seg_off, s->t, s->flux, s->err, s->rank, s->n_obj, T1, 0, T1 * TB,
                                                 nullptr, cache, s->ticket + 2, out_dev, status_dev);
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFreeAsync(cache, st); return cuda_fail(e, "k_gp2d launch"); }
  }
  count_launch();
  if (any_beyond) {
    double* scratch = nullptr;
    MLRN_CUDA(cudaMallocAsync(&scratch, scratch2 * grid2, st));
    k_gp2d<false><<<grid2, GP_THREADS, smem2, st>>>(s->seg_off, s->t, s->flux, s->err, s->rank, s->n_obj, T_need,
                                                   T1 * TB + 1, 0x7fffffff, scratch, cache, s->ticket + 3, out_dev, status_dev);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) { cudaFreeAsync(cache, st); return cuda_fail(e, "k_gp2d<global> launch"); }
    count_launch();
  }
  cudaFreeAsync(cache, st);
  return MLRN_OK;
}

}  // namespace mlrn
