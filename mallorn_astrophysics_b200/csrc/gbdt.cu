// gbdt.cu -- K10: histogram gradient-boosted trees trained on a DEVICE-RESIDENT feature table.
//
// Reference: the downstream drivers train XGBoost with tree_method='hist' on the merged feature table --
// non_successful_tests/scripts/train_v92_focal_adversarial.py:157-169,232-272 (v92d: binary:logistic, max_depth 5,
// learning_rate 0.025, subsample 0.8, colsample_bytree 0.8, min_child_weight 3, reg_alpha 0.2, reg_lambda 1.5,
// scale_pos_weight, adversarial sample weights, 500 rounds with early stopping) and adversarial_validation.py:120-160
// (the train-vs-test classifier those weights come from).  xgboost is a third-party dependency that is not installed
// here; this file restates its published `hist` algorithm (Chen & Guestrin 2016; depth-wise growth on quantised features,
// second-order gain with L1/L2 regularisation, learned default direction for missing values) for the binary-logistic
// objective.  The CPU restatement of the SAME algorithm is oracle/gbdt.py and the two agree bit for bit:
//   * gradients are quantised once to multiples of 2^-30 and histograms are accumulated in int64, so sums do not depend
//     on the order of the atomics;
//   * the sigmoid uses the fixed polynomial exp of gbdt_exp() and the gain uses explicitly rounded operations (no FMA
//     contraction), so both sides execute the same IEEE operations;
//   * row and column sub-sampling draw from a counter-based hash of (seed, round, index), not from a stream.
// One call grows one tree: gradients -> per level {histograms, best split per (node, feature), best split per node, row
// re-assignment} -> leaf values -> margins, all on one stream without host synchronisation.
#include <cuda_runtime.h>
#include <new>
#include "store.cuh"
#include "../../include/mallorn_b200.h"

struct mlrn_gbdt {
  int device = 0;
  int64_t n = 0;
  int32_t F = 0;
  mlrn_gbdt_params p{};
  const uint8_t* bins = nullptr;          // [n, F] caller-owned, row-major; bin 255 = missing
  const double* y = nullptr;              // [n] 0 / 1, caller-owned
  const double* w = nullptr;              // [n] sample weights or NULL, caller-owned
  double* margin = nullptr;               // [n] raw scores of the training rows
  long long* gq = nullptr;                // [n] quantised gradient, 0 for rows outside this round's sub-sample
  long long* hq = nullptr;
  int* node = nullptr;                    // [n] node of the complete binary tree the row is in
  long long* hist = nullptr;              // [16 nodes][F][256][2]
  double* cand = nullptr;                 // [16][F][4]: gain, bin, default_left, (unused)
  long long* sums = nullptr;              // [63][2] G, H of every node
  mlrn_gbdt_tree* trees = nullptr;        // [rounds_cap]
  uint8_t* colmask = nullptr;             // [F]
  int32_t rounds_cap = 0, rounds = 0;
};

namespace mlrn {
namespace {

constexpr int GB_BINS = 256, GB_MISSING = 255, GB_NODES = 63, GB_MAX_DEPTH = 5, GB_LEVEL_NODES = 16;
constexpr double GB_Q = 1073741824.0;     // 2^30: gradient quantum

// exp(x) as a fixed sequence of correctly rounded operations (oracle/gbdt.py::gbdt_exp executes the same ones):
// x = k ln2 + r with the two-constant Cody-Waite split, a degree-12 Taylor polynomial in r, scaling by 2^k.
__host__ __device__ inline double gbdt_exp(double x) {
  if (x > 700.0) x = 700.0;
  if (x < -700.0) x = -700.0;
#ifdef __CUDA_ARCH__
  const double kf = rint(__dmul_rn(x, 1.4426950408889634));
  const double r = __dadd_rn(__dadd_rn(x, -__dmul_rn(kf, 0.693147180369123816490)), -__dmul_rn(kf, 1.90821492927058770002e-10));
  double p = 1.0 / 479001600.0;
  const double c[12] = {1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
                        1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
#pragma unroll
  for (int i = 0; i < 12; ++i) p = __dadd_rn(__dmul_rn(p, r), c[i]);
  return ldexp(p, (int)kf);
#else
  return x;     // host side never evaluates it
#endif
}

__device__ __forceinline__ unsigned long long gb_hash(unsigned long long seed, unsigned long long a, unsigned long long b) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (a + 1) + 0xBF58476D1CE4E5B9ULL * (b + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// gradients of the weighted logistic loss, quantised; rows outside the sub-sample get zeros; every row restarts at the root
__global__ void k_gb_grad(const double* __restrict__ margin, const double* __restrict__ y, const double* __restrict__ w, long long n,
                          double scale_pos_weight, double subsample, unsigned long long seed, int round, long long* __restrict__ gq,
                          long long* __restrict__ hq, int* __restrict__ node) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    node[i] = 0;
    const double u = (double)(gb_hash(seed, (unsigned long long)round, (unsigned long long)i) >> 11) * (1.0 / 9007199254740992.0);
    if (!(u < subsample)) { gq[i] = 0; hq[i] = 0; continue; }
    const double p = __ddiv_rn(1.0, __dadd_rn(1.0, gbdt_exp(-margin[i])));
    double wt = w ? w[i] : 1.0;
    if (y[i] > 0.5) wt = __dmul_rn(wt, scale_pos_weight);
    const double g = __dmul_rn(__dadd_rn(p, -y[i]), wt);
    double h = __dmul_rn(__dmul_rn(p, __dadd_rn(1.0, -p)), wt);
    if (h < 1e-16) h = 1e-16;
    gq[i] = llrint(__dmul_rn(g, GB_Q));
    hq[i] = llrint(__dmul_rn(h, GB_Q));
  }
}

// root sums
__global__ void k_gb_root(const long long* __restrict__ gq, const long long* __restrict__ hq, long long n, long long* __restrict__ sums) {
  long long g = 0, h = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { g += gq[i]; h += hq[i]; }
  for (int o = 16; o > 0; o >>= 1) { g += __shfl_xor_sync(0xffffffffu, g, o); h += __shfl_xor_sync(0xffffffffu, h, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd((unsigned long long*)&sums[0], (unsigned long long)g); atomicAdd((unsigned long long*)&sums[1], (unsigned long long)h); }
}

// histograms of one level: every (row, feature) adds its gradient pair to hist[node - first][feature][bin]
__global__ void k_gb_hist(const uint8_t* __restrict__ bins, const long long* __restrict__ gq, const long long* __restrict__ hq,
                          const int* __restrict__ node, const mlrn_gbdt_tree* __restrict__ tree, const uint8_t* __restrict__ colmask,
                          long long n, int F, int first, long long* __restrict__ hist) {
  const long long total = n * F;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / F;
    const int f = (int)(idx - i * F);
    const int nd = node[i];
    if (nd < first || !colmask[f]) continue;          // row already in a leaf above this level / column not sampled
    const long long g = gq[i], h = hq[i];
    if (g == 0 && h == 0) continue;
    long long* cell = hist + ((((long long)(nd - first) * F + f) * GB_BINS) + bins[idx]) * 2;
    atomicAdd((unsigned long long*)cell, (unsigned long long)g);
    atomicAdd((unsigned long long*)(cell + 1), (unsigned long long)h);
  }
}

__device__ __forceinline__ double gb_score(double G, double H, double alpha, double lambda) {
  const double a = fabs(G) - alpha;
  const double T = a > 0.0 ? a : 0.0;
  return __ddiv_rn(__dmul_rn(T, T), __dadd_rn(H, lambda));
}

// best split of every (node of the level, feature): one warp scans the 255 value bins with the missing bin sent left or right
__global__ void k_gb_split(const long long* __restrict__ hist, const long long* __restrict__ sums, const uint8_t* __restrict__ colmask,
                           int F, int first, int n_nodes, mlrn_gbdt_params p, double* __restrict__ cand) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_nodes * F) return;
  const int nd = warp / F, f = warp - nd * F;
  double* out = cand + ((long long)nd * F + f) * 4;
  const long long Gt = sums[(first + nd) * 2], Ht = sums[(first + nd) * 2 + 1];
  if (!colmask[f] || Ht == 0) { if (lane == 0) { out[0] = 0.0; out[1] = -1.0; out[2] = 0.0; } return; }
  const long long* hb = hist + (((long long)nd * F + f) * GB_BINS) * 2;
  // lane owns bins 8*lane .. 8*lane+7; exclusive prefix over lanes
  long long g[8], h[8], gs = 0, hs = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { g[k] = hb[(8 * lane + k) * 2]; h[k] = hb[(8 * lane + k) * 2 + 1]; }
  const long long gm = hb[GB_MISSING * 2], hm = hb[GB_MISSING * 2 + 1];
  if (lane == 31) { g[7] = 0; h[7] = 0; }             // bin 255 is the missing bin, not a value bin
#pragma unroll
  for (int k = 0; k < 8; ++k) { gs += g[k]; hs += h[k]; }
  long long gp = gs, hp = hs;
  for (int o = 1; o < 32; o <<= 1) {
    const long long a = __shfl_up_sync(0xffffffffu, gp, o), b = __shfl_up_sync(0xffffffffu, hp, o);
    if (lane >= o) { gp += a; hp += b; }
  }
  gp -= gs; hp -= hs;                                 // exclusive
  const double parent = gb_score((double)Gt / GB_Q, (double)Ht / GB_Q, p.reg_alpha, p.reg_lambda);
  double best = 0.0; int best_bin = -1, best_left = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    gp += g[k]; hp += h[k];                           // left = value bins 0 .. 8*lane+k
    const int bin = 8 * lane + k;
    if (bin >= GB_MISSING - 1) continue;              // a split after the last value bin separates nothing but the missing rows
#pragma unroll
    for (int miss_left = 0; miss_left < 2; ++miss_left) {
      const long long GL = gp + (miss_left ? gm : 0), HL = hp + (miss_left ? hm : 0);
      const long long GR = Gt - GL, HR = Ht - HL;
      const double hl = (double)HL / GB_Q, hr = (double)HR / GB_Q;
      if (hl < p.min_child_weight || hr < p.min_child_weight) continue;
      const double gain = __dadd_rn(__dadd_rn(gb_score((double)GL / GB_Q, hl, p.reg_alpha, p.reg_lambda),
                                              gb_score((double)GR / GB_Q, hr, p.reg_alpha, p.reg_lambda)), -parent);
      if (gain > best) { best = gain; best_bin = bin; best_left = miss_left; }      // first (lowest bin, missing right) wins ties
    }
  }
  // warp argmax: larger gain, then lower bin, then missing-right
  for (int o = 16; o > 0; o >>= 1) {
    const double og = __shfl_xor_sync(0xffffffffu, best, o);
    const int ob = __shfl_xor_sync(0xffffffffu, best_bin, o), ol = __shfl_xor_sync(0xffffffffu, best_left, o);
    const bool take = ob >= 0 && (best_bin < 0 || og > best || (og == best && (ob < best_bin || (ob == best_bin && ol < best_left))));
    if (take) { best = og; best_bin = ob; best_left = ol; }
  }
  if (lane == 0) { out[0] = best; out[1] = (double)best_bin; out[2] = (double)best_left; }
}

// best feature per node (lowest feature index wins ties); writes the tree node and the children's sums
__global__ void k_gb_best(const double* __restrict__ cand, const long long* __restrict__ hist, int F, int first, int n_nodes, int depth,
                          mlrn_gbdt_params p, mlrn_gbdt_tree* __restrict__ tree, long long* __restrict__ sums) {
  const int nd = blockIdx.x;
  if (nd >= n_nodes) return;
  __shared__ double s_gain[128];
  __shared__ int s_f[128];
  double best = 0.0; int bf = -1;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const double* c = cand + ((long long)nd * F + f) * 4;
    if (c[1] >= 0.0 && (bf < 0 || c[0] > best)) { best = c[0]; bf = f; }             // ascending f: first of equal gains stays
  }
  s_gain[threadIdx.x] = best; s_f[threadIdx.x] = bf;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < blockDim.x; ++t) {
      const int of = s_f[t]; const double og = s_gain[t];
      if (of >= 0 && (bf < 0 || og > best || (og == best && of < bf))) { best = og; bf = of; }
    }
    const int id = first + nd;
    const long long Gt = sums[id * 2], Ht = sums[id * 2 + 1];
    const bool split = depth < p.max_depth && bf >= 0 && best > 1e-6 && Ht != 0;     // xgboost's kRtEps
    if (!split) {
      tree->feature[id] = -1; tree->bin[id] = 0; tree->default_left[id] = 0;
      double val = 0.0;
      if (Ht != 0) {
        const double G = (double)Gt / GB_Q, H = (double)Ht / GB_Q;
        const double a = fabs(G) - p.reg_alpha, T = a > 0.0 ? (G > 0.0 ? a : -a) : 0.0;
        val = __dmul_rn(-__ddiv_rn(T, __dadd_rn(H, p.reg_lambda)), p.learning_rate);
      }
      tree->value[id] = val;
    } else {
      const double* c = cand + ((long long)nd * F + bf) * 4;
      const int bin = (int)c[1], left = (int)c[2];
      tree->feature[id] = bf; tree->bin[id] = bin; tree->default_left[id] = left; tree->value[id] = 0.0;
      const long long* hb = hist + (((long long)nd * F + bf) * GB_BINS) * 2;
      long long GL = 0, HL = 0;
      for (int b = 0; b <= bin; ++b) { GL += hb[b * 2]; HL += hb[b * 2 + 1]; }
      if (left) { GL += hb[GB_MISSING * 2]; HL += hb[GB_MISSING * 2 + 1]; }
      sums[(2 * id + 1) * 2] = GL; sums[(2 * id + 1) * 2 + 1] = HL;
      sums[(2 * id + 2) * 2] = Gt - GL; sums[(2 * id + 2) * 2 + 1] = Ht - HL;
    }
  }
}

// rows of split nodes move to a child
__global__ void k_gb_assign(const uint8_t* __restrict__ bins, const mlrn_gbdt_tree* __restrict__ tree, long long n, int F, int first,
                            int* __restrict__ node) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int nd = node[i];
    if (nd < first) continue;
    const int f = tree->feature[nd];
    if (f < 0) continue;
    const int b = bins[i * F + f];
    const bool left = (b == GB_MISSING) ? (tree->default_left[nd] != 0) : (b <= tree->bin[nd]);
    node[i] = 2 * nd + (left ? 1 : 2);
  }
}

__global__ void k_gb_update(const int* __restrict__ node, const mlrn_gbdt_tree* __restrict__ tree, long long n, double* __restrict__ margin) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    margin[i] = __dadd_rn(margin[i], tree->value[node[i]]);
}

// margins of other rows under trees [from, to)
__global__ void k_gb_predict(const uint8_t* __restrict__ bins, long long n, int F, const mlrn_gbdt_tree* __restrict__ trees, int from, int to,
                             double* __restrict__ margin) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double m = margin[i];
    for (int t = from; t < to; ++t) {
      const mlrn_gbdt_tree* tr = trees + t;
      int nd = 0;
      for (;;) {
        const int f = tr->feature[nd];
        if (f < 0) break;
        const int b = bins[i * F + f];
        nd = 2 * nd + (((b == GB_MISSING) ? (tr->default_left[nd] != 0) : (b <= tr->bin[nd])) ? 1 : 2);
      }
      m = __dadd_rn(m, tr->value[nd]);
    }
    margin[i] = m;
  }
}

// value -> bin: number of cuts below or equal ... bin b holds cuts[b-1] < v <= cuts[b]; NaN -> missing
__global__ void k_gb_bin(const double* __restrict__ X, long long n, int F, int ld, const double* __restrict__ cuts, const int* __restrict__ ncuts,
                         uint8_t* __restrict__ bins) {
  const long long total = n * F;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / F;
    const int f = (int)(idx - i * F);
    const double v = X[i * ld + f];
    int b = GB_MISSING;
    if (v == v) {
      const double* c = cuts + (long long)f * (GB_BINS - 1);
      int lo = 0, hi = ncuts[f];                      // first cut >= v
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (c[mid] < v) lo = mid + 1; else hi = mid; }
      b = lo < GB_MISSING - 1 ? lo : GB_MISSING - 1;
    }
    bins[idx] = (uint8_t)b;
  }
}

inline int grid_for(long long work, int sms) {
  long long want = (work + 255) / 256;
  const long long cap = (long long)sms * 16;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace
}  // namespace mlrn

using namespace mlrn;

extern "C" {

int mlrn_gbdt_bin(int device, const double* X, int64_t n, int32_t F, int32_t ld, const double* cuts, const int32_t* ncuts, uint8_t* bins,
                  void* stream) {
  if (n < 0 || F <= 0 || ld < F || (n > 0 && (!X || !cuts || !ncuts || !bins))) { set_error("mlrn_gbdt_bin: bad argument"); return MLRN_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device visible; libmallorn_b200 has no CPU fallback"); return MLRN_ENODEVICE; }
  if (n == 0) return MLRN_OK;
  cudaSetDevice(device);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  k_gb_bin<<<grid_for(n * F, sms), 256, 0, (cudaStream_t)stream>>>(X, n, F, ld, cuts, ncuts, bins);
  MLRN_CUDA(cudaGetLastError());
  count_launch();
  return MLRN_OK;
}

int mlrn_gbdt_create(int device, const uint8_t* bins, int64_t n, int32_t F, const double* y, const double* weight,
                     const mlrn_gbdt_params* params, int32_t max_rounds, void* stream, mlrn_gbdt** out) {
  if (!out) { set_error("out is null"); return MLRN_EINVAL; }
  *out = nullptr;
  if (!bins || !y || !params || n <= 0 || F <= 0 || max_rounds <= 0 || params->max_depth < 1 || params->max_depth > GB_MAX_DEPTH) {
    set_error("mlrn_gbdt_create: bad argument (max_depth must be 1..5)"); return MLRN_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device visible; libmallorn_b200 has no CPU fallback"); return MLRN_ENODEVICE; }
  cudaSetDevice(device);
  mlrn_gbdt* g = new (std::nothrow) mlrn_gbdt();
  if (!g) { set_error("out of host memory"); return MLRN_EINVAL; }
  g->device = device; g->n = n; g->F = F; g->p = *params; g->bins = bins; g->y = y; g->w = weight; g->rounds_cap = max_rounds;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMalloc(&g->margin, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMalloc(&g->gq, sizeof(long long) * n);
  if (e == cudaSuccess) e = cudaMalloc(&g->hq, sizeof(long long) * n);
  if (e == cudaSuccess) e = cudaMalloc(&g->node, sizeof(int) * n);
  if (e == cudaSuccess) e = cudaMalloc(&g->hist, sizeof(long long) * (size_t)GB_LEVEL_NODES * F * GB_BINS * 2);
  if (e == cudaSuccess) e = cudaMalloc(&g->cand, sizeof(double) * (size_t)GB_LEVEL_NODES * F * 4);
  if (e == cudaSuccess) e = cudaMalloc(&g->sums, sizeof(long long) * GB_NODES * 2);
  if (e == cudaSuccess) e = cudaMalloc(&g->trees, sizeof(mlrn_gbdt_tree) * (size_t)max_rounds);
  if (e == cudaSuccess) e = cudaMalloc(&g->colmask, (size_t)F);
  if (e == cudaSuccess) e = cudaMemsetAsync(g->margin, 0, sizeof(double) * n, st);      // base_score 0.5 -> margin 0
  if (e != cudaSuccess) { mlrn_gbdt_destroy(g); return cuda_fail(e, "mlrn_gbdt_create"); }
  *out = g;
  return MLRN_OK;
}

int mlrn_gbdt_destroy(mlrn_gbdt* g) {
  if (!g) return MLRN_OK;
  cudaSetDevice(g->device);
  cudaFree(g->margin); cudaFree(g->gq); cudaFree(g->hq); cudaFree(g->node); cudaFree(g->hist); cudaFree(g->cand); cudaFree(g->sums);
  cudaFree(g->trees); cudaFree(g->colmask);
  delete g;
  return MLRN_OK;
}

int mlrn_gbdt_round(mlrn_gbdt* g, const uint8_t* colmask_host, void* stream) {
  if (!g || !colmask_host) { set_error("mlrn_gbdt_round: null argument"); return MLRN_EINVAL; }
  if (g->rounds >= g->rounds_cap) { set_error("mlrn_gbdt_round: more rounds than the handle was created for"); return MLRN_EINVAL; }
  cudaSetDevice(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device);
  const int round = g->rounds;
  mlrn_gbdt_tree* tree = g->trees + round;
  MLRN_CUDA(cudaMemcpyAsync(g->colmask, colmask_host, (size_t)g->F, cudaMemcpyHostToDevice, st));
  MLRN_CUDA(cudaMemsetAsync(g->sums, 0, sizeof(long long) * GB_NODES * 2, st));
  MLRN_CUDA(cudaMemsetAsync(tree, 0xff, sizeof(mlrn_gbdt_tree), st));                 // feature -1 everywhere; values rewritten below
  k_gb_grad<<<grid_for(g->n, sms), 256, 0, st>>>(g->margin, g->y, g->w, g->n, g->p.scale_pos_weight, g->p.subsample, g->p.seed, round,
                                                 g->gq, g->hq, g->node);
  k_gb_root<<<grid_for(g->n, sms), 256, 0, st>>>(g->gq, g->hq, g->n, g->sums);
  int launches = 2;
  for (int depth = 0; depth <= g->p.max_depth; ++depth) {
    const int first = (1 << depth) - 1, n_nodes = 1 << depth;
    if (depth < g->p.max_depth) {
      MLRN_CUDA(cudaMemsetAsync(g->hist, 0, sizeof(long long) * (size_t)n_nodes * g->F * GB_BINS * 2, st));
      k_gb_hist<<<grid_for(g->n * g->F, sms), 256, 0, st>>>(g->bins, g->gq, g->hq, g->node, tree, g->colmask, g->n, g->F, first, g->hist);
      k_gb_split<<<(n_nodes * g->F * 32 + 255) / 256, 256, 0, st>>>(g->hist, g->sums, g->colmask, g->F, first, n_nodes, g->p, g->cand);
      launches += 2;
    }
    k_gb_best<<<n_nodes, 128, 0, st>>>(g->cand, g->hist, g->F, first, n_nodes, depth, g->p, tree, g->sums);
    ++launches;
    if (depth < g->p.max_depth) {
      k_gb_assign<<<grid_for(g->n, sms), 256, 0, st>>>(g->bins, tree, g->n, g->F, first, g->node);
      ++launches;
    }
  }
  k_gb_update<<<grid_for(g->n, sms), 256, 0, st>>>(g->node, tree, g->n, g->margin);
  MLRN_CUDA(cudaGetLastError());
  count_launch(launches + 1);
  ++g->rounds;
  return MLRN_OK;
}

int mlrn_gbdt_predict(const mlrn_gbdt* g, const uint8_t* bins, int64_t n, int32_t tree_from, int32_t tree_to, double* margin, void* stream) {
  if (!g || (n > 0 && (!bins || !margin)) || tree_from < 0 || tree_to > g->rounds || tree_from > tree_to) { set_error("mlrn_gbdt_predict: bad argument"); return MLRN_EINVAL; }
  if (n == 0 || tree_from == tree_to) return MLRN_OK;
  cudaSetDevice(g->device);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device);
  k_gb_predict<<<grid_for(n, sms), 256, 0, (cudaStream_t)stream>>>(bins, n, g->F, g->trees, tree_from, tree_to, margin);
  MLRN_CUDA(cudaGetLastError());
  count_launch();
  return MLRN_OK;
}

int mlrn_gbdt_export(const mlrn_gbdt* g, mlrn_gbdt_tree* trees_host, double* train_margin_host) {
  if (!g) { set_error("null handle"); return MLRN_EINVAL; }
  cudaSetDevice(g->device);
  MLRN_CUDA(cudaDeviceSynchronize());
  if (trees_host && g->rounds) MLRN_CUDA(cudaMemcpy(trees_host, g->trees, sizeof(mlrn_gbdt_tree) * (size_t)g->rounds, cudaMemcpyDeviceToHost));
  if (train_margin_host) MLRN_CUDA(cudaMemcpy(train_margin_host, g->margin, sizeof(double) * (size_t)g->n, cudaMemcpyDeviceToHost));
  return MLRN_OK;
}

int32_t mlrn_gbdt_rounds(const mlrn_gbdt* g) { return g ? g->rounds : 0; }

}  // extern "C"
