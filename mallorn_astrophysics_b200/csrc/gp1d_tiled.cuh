// gp1d_tiled.cuh -- register-tiled evaluation of the per-band GP objective (EXPERIMENTAL: built into the library,
// selected only with MLRN_GP1D_TILED=1 until it has been validated and timed on the GPU; the default path is
// gp1d_objective in gp1d.cuh).
//
// Why: the shared-memory sweep of gp1d.cuh spends two loads and a store per FMA (25 k warp instructions per
// evaluation for ~15 k FMA; profiles/r01_gp1d_v2_full.txt).  Here the symmetric (n+1) x (n+1) matrix is cut into
// 4 x 4 tiles; the tiles of the lower triangle (diagonal tiles kept as full symmetric blocks) are dealt round-robin
// to the 32 lanes and stay in REGISTERS for the whole sweep.  Per pivot k every lane fetches, for each of its tiles,
// four entries of the pivot column for its rows and four for its columns (shared memory, written by the lanes that
// own column k) and does 16 FMA; the pivot row / column fix-ups touch only the tiles in tile-row or tile-column k/4.
// One barrier per pivot (the pivot column is double-buffered).  Same arithmetic as gp1d_objective up to the order of
// the final reductions.
//
// The per-lane code is written against a "lanes" executor so that the same source runs on the GPU (one real lane)
// and in tests/hostsim (32 emulated lanes, phases separated at the barriers) -- the tile bookkeeping is checked
// against the reference goldens on the CPU.
#pragma once
#include "gp1d.cuh"

namespace mlrn {

constexpr int GT = 4;                         // tile edge
MLRN_HD int gp1d_tiled_capacity(int tpl) {    // largest matrix order (n + 1) that fits tpl tiles per lane
  int nt = 0;
  while ((nt + 1) * (nt + 2) / 2 <= 32 * tpl) ++nt;
  return GT * nt;
}

template <int TPL>
struct TileRegs {
  double a[TPL][GT][GT];
  int bi[TPL], bj[TPL];                       // tile coordinates, bi >= bj; bi < 0: slot unused
  double g0, g1, g2, ld;                      // partial sums of the gradient pass
};

#ifdef __CUDACC__
template <class State>
struct WarpLanes {
  State st;
  int lane;
  __device__ WarpLanes() : lane(threadIdx.x & 31) {}
  template <class F> __device__ __forceinline__ void each(F&& f) { f(lane, st); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  template <class G> __device__ __forceinline__ double sum(G&& get) {
    double v = get(st);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
};
#else
template <class State>
struct HostLanes {                            // tests/hostsim only: 32 emulated lanes
  State st[32];
  template <class F> void each(F&& f) { for (int l = 0; l < 32; ++l) f(l, st[l]); }
  void sync() const {}
  template <class G> double sum(G&& get) { double s = 0.0; for (int l = 0; l < 32; ++l) s += get(st[l]); return s; }
};
#endif

MLRN_FN_FORCE double sel4(double v0, double v1, double v2, double v3, int q) { return q == 0 ? v0 : (q == 1 ? v1 : (q == 2 ? v2 : v3)); }

// x, y, al: the prepared band [n]; vb0, vb1: two buffers of GT * ceil((n+1)/GT) doubles; piv: [n].
template <int TPL, class Lanes>
MLRN_FN bool gp1d_objective_tiled(Lanes& ex, int n, const double* __restrict__ x, const double* __restrict__ y,
                                  const double* __restrict__ al, double* __restrict__ vb0, double* __restrict__ vb1,
                                  double* __restrict__ piv, const double* theta, double& fval, double* grad) {
  using Regs = TileRegs<TPL>;
  const double c = exp(theta[0]), l = exp(theta[1]), sn = exp(theta[2]);
  const double il2 = 1.0 / (l * l);
  const int n1 = n + 1, nt = (n1 + GT - 1) / GT, ntiles = nt * (nt + 1) / 2;
  // build: each lane fills its tiles
  ex.each([&](int lane, Regs& R) {
#pragma unroll
    for (int s = 0; s < TPL; ++s) {
      const int tau = lane + 32 * s;
      int bi = -1, bj = 0;
      if (tau < ntiles) { bi = 0; int rem = tau; while (rem > bi) { rem -= bi + 1; ++bi; } bj = rem; }
      R.bi[s] = bi; R.bj[s] = bj;
#pragma unroll
      for (int r = 0; r < GT; ++r)
#pragma unroll
        for (int q = 0; q < GT; ++q) {
          double val = 0.0;
          const int i = GT * bi + r, j = GT * bj + q;
          if (bi >= 0 && i <= n && j <= n) {
            if (i == n || j == n) val = (i == j) ? 0.0 : y[i == n ? j : i];          // right-hand side row / column
            else {
              const double d = x[i] - x[j];
              val = c * exp(-0.5 * (d * d * il2));
              if (i == j) val += sn + al[i];
            }
          }
          R.a[s][r][q] = val;
        }
    }
  });
  // column k of the symmetric matrix -> vb (the lanes that own tile column k/4, and tile row k/4 left of the diagonal)
  auto publish = [&](int k, double* __restrict__ vb) {
    ex.each([&](int, Regs& R) {
      const int kb = k >> 2, q = k & 3;
#pragma unroll
      for (int s = 0; s < TPL; ++s) {
        if (R.bi[s] < 0) continue;
        if (R.bj[s] == kb) {
#pragma unroll
          for (int r = 0; r < GT; ++r) vb[GT * R.bi[s] + r] = sel4(R.a[s][r][0], R.a[s][r][1], R.a[s][r][2], R.a[s][r][3], q);
        } else if (R.bi[s] == kb) {
#pragma unroll
          for (int p = 0; p < GT; ++p) vb[GT * R.bj[s] + p] = sel4(R.a[s][0][p], R.a[s][1][p], R.a[s][2][p], R.a[s][3][p], q);
        }
      }
    });
  };
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    double* __restrict__ vb = (k & 1) ? vb1 : vb0;
    publish(k, vb);
    ex.sync();
    const double pv = vb[k];
    if (!(pv > 0.0)) { ok = false; break; }
    const double ip = 1.0 / pv;
    ex.each([&](int lane, Regs& R) {
      if (lane == 0) piv[k] = pv;
      const int kb = k >> 2, q = k & 3;
#pragma unroll
      for (int s = 0; s < TPL; ++s) {
        if (R.bi[s] < 0) continue;
        double vr[GT], wc[GT];
#pragma unroll
        for (int r = 0; r < GT; ++r) { vr[r] = vb[GT * R.bi[s] + r]; wc[r] = vb[GT * R.bj[s] + r] * ip; }
#pragma unroll
        for (int r = 0; r < GT; ++r)
#pragma unroll
          for (int p = 0; p < GT; ++p) R.a[s][r][p] = R.a[s][r][p] - vr[r] * wc[p];
        if (R.bj[s] == kb) {                                   // pivot column: w_i
#pragma unroll
          for (int r = 0; r < GT; ++r) {
            const double wv = vr[r] * ip;
#pragma unroll
            for (int p = 0; p < GT; ++p) R.a[s][r][p] = (p == q) ? wv : R.a[s][r][p];
          }
        }
        if (R.bi[s] == kb) {                                   // pivot row: w_j, and -1/pivot on the diagonal
#pragma unroll
          for (int r = 0; r < GT; ++r)
#pragma unroll
            for (int p = 0; p < GT; ++p) R.a[s][r][p] = (r == q) ? ((R.bj[s] == kb && p == q) ? -ip : wc[p]) : R.a[s][r][p];
        }
      }
    });
  }
  if (!ok) { fval = pinf(); grad[0] = grad[1] = grad[2] = 0.0; ex.sync(); return false; }
  // row n = K^-1 y (and -y'K^-1 y in the corner), published like a pivot column
  double* __restrict__ av = (n & 1) ? vb1 : vb0;
  publish(n, av);
  ex.sync();
  ex.each([&](int lane, Regs& R) {
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, ld = 0.0;
#pragma unroll
    for (int s = 0; s < TPL; ++s) {
      if (R.bi[s] < 0) continue;
#pragma unroll
      for (int r = 0; r < GT; ++r)
#pragma unroll
        for (int p = 0; p < GT; ++p) {
          const int i = GT * R.bi[s] + r, j = GT * R.bj[s] + p;
          if (i < n && j <= i) {
            const double W = av[i] * av[j] + R.a[s][r][p];     // (a a' - K^-1)_ij
            if (i == j) { g0 += W * c; g2 += W; }
            else {
              const double d = x[i] - x[j];
              const double qq = d * d * il2;
              const double C = c * exp(-0.5 * qq);
              g0 += 2.0 * W * C; g1 += 2.0 * W * C * qq;
            }
          }
        }
    }
    for (int i = lane; i < n; i += 32) ld += log(piv[i]);
    R.g0 = g0; R.g1 = g1; R.g2 = g2; R.ld = ld;
  });
  const double g0 = ex.sum([](const Regs& R) { return R.g0; });
  const double g1 = ex.sum([](const Regs& R) { return R.g1; });
  const double g2 = ex.sum([](const Regs& R) { return R.g2; });
  const double logdet = ex.sum([](const Regs& R) { return R.ld; });
  const double yKy = -av[n];
  const double lml = -0.5 * yKy - 0.5 * logdet - 0.5 * (double)n * 1.8378770664093453;   // ln(2 pi)
  fval = -lml;
  grad[0] = -0.5 * g0; grad[1] = -0.5 * g1; grad[2] = -0.5 * sn * g2;
  ex.sync();
  return true;
}

// scratch layout of the tiled path inside the same per-warp slice: x, y, al [cap+1], two pivot-column buffers, pivots
struct Gp1dTiledScratch { double *x, *y, *al, *vb0, *vb1, *piv; };
MLRN_HD int gp1d_tiled_scratch_doubles(int cap) { return 4 * (cap + 1) + 2 * (cap + 1 + GT); }
MLRN_FN Gp1dTiledScratch gp1d_carve_tiled(double* base, int cap) {
  Gp1dTiledScratch t;
  t.x = base; t.y = t.x + (cap + 1); t.al = t.y + (cap + 1); t.piv = t.al + (cap + 1);
  t.vb0 = t.piv + (cap + 1); t.vb1 = t.vb0 + (cap + 1 + GT);
  return t;
}

}  // namespace mlrn
