// fits.cu -- K4 kernels: Bazin (52 columns) and decline-model (27 columns) passes.
// One warp per object; each band is staged from the packed store into the warp's shared-memory
// slice (coalesced) and fitted there by trf_bounded() (trf.cuh).  Grid = resident CTAs x SM count,
// warps take objects from a global ticket counter (fit cost varies by 100x between objects).
#include <cuda_runtime.h>
#include "store.cuh"
#include "fits.cuh"
#include "fits_thread.cuh"
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


// ------------------------------------------------------------------------------------------------
// One THREAD per fit (trf_thread.cuh): every lane of a warp advances its own fit by one trial
// evaluation per trip, and a lane whose fit ends takes the next (object, band[, model]) from the
// ticket counter, so the 100x spread in iteration counts does not idle the warp.
// ------------------------------------------------------------------------------------------------
constexpr int FIT_THREADS = 128;

__global__ void __launch_bounds__(FIT_THREADS, 4)
k_bazin_thread(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
               const double* __restrict__ ge, int n_obj, unsigned long long* __restrict__ ticket,
               double* __restrict__ out, int* __restrict__ status) {
  const long long total = (long long)n_obj * N_BAND;
  BazinJob job;
  bool active = false, exhausted = false;
  long long fit = 0;
  for (;;) {
    for (int tries = 0; tries < 4 && !active && !exhausted; ++tries) {
      fit = (long long)atomicAdd(ticket, 1ULL);
      if (fit >= total) { exhausted = true; break; }
      const long long o = fit / N_BAND; const int b = (int)(fit % N_BAND);
      const long long lo = seg_off[o * N_SEG + b];
      const int n = (int)(seg_off[o * N_SEG + b + 1] - lo);
      int code;
      active = job.init(gt + lo, gf + lo, ge + lo, n, code);
      if (!active) {
        job.finish(code, out + o * BAZIN_COLS + 8 * b);
        if (status) { status[fit * 2] = code; status[fit * 2 + 1] = 0; }
      }
    }
    if (!__any_sync(0xffffffffu, active || !exhausted)) break;
    if (active) {
      const int rc = job.step();
      if (rc != TRF_RUNNING) {
        const long long o = fit / N_BAND; const int b = (int)(fit % N_BAND);
        job.finish(rc, out + o * BAZIN_COLS + 8 * b);
        if (status) { status[fit * 2] = rc; status[fit * 2 + 1] = job.trf.nfev; }
        active = false;
      }
    }
  }
}

// cross-band columns 48..51 (bazin_fitting.py:217-249), one thread per object
__global__ void k_bazin_tail(int n_obj, double* __restrict__ out) {
  const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (o >= n_obj) return;
  double* row = out + o * BAZIN_COLS;
  double rise[3] = {row[8 * BG + 2], row[8 * BR + 2], row[8 * BI + 2]};
  double fall[3] = {row[8 * BG + 3], row[8 * BR + 3], row[8 * BI + 3]};
  int c; double m, sd;
  small_stats(rise, 3, c, m, sd); row[48] = (c >= 2) ? sd / m : qnan();
  small_stats(fall, 3, c, m, sd); row[49] = (c >= 2) ? sd / m : qnan();
  double chi[6];
  for (int b = 0; b < 6; ++b) chi[b] = row[8 * b + 5];
  small_stats(chi, 6, c, m, sd); row[50] = c ? m : qnan(); row[51] = c ? sd : qnan();
}

// KIND 0: the seven fixed-index power laws (21 fits per object); 1: exponential; 2: linear (3 per object each)
template <int KIND>
__global__ void __launch_bounds__(FIT_THREADS, 4)
k_decline_thread(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
                 int n_obj, unsigned long long* __restrict__ ticket, double* __restrict__ out, int* __restrict__ status) {
  constexpr int PER_BAND = (KIND == 0) ? 7 : 1;
  const long long total = (long long)n_obj * 3 * PER_BAND;
  DeclineJob<KIND> job;
  bool active = false, exhausted = false;
  long long fit = 0;
  auto slot = [&](long long ft) {                     // column of this fit inside the 27-wide row
    const long long o = ft / (3 * PER_BAND); const int rem = (int)(ft % (3 * PER_BAND));
    const int j = rem / PER_BAND, q = (KIND == 0) ? rem % PER_BAND : (KIND == 1 ? 7 : 8);
    return o * POWERLAW_COLS + 9 * j + q;
  };
  for (;;) {
    for (int tries = 0; tries < 4 && !active && !exhausted; ++tries) {
      fit = (long long)atomicAdd(ticket, 1ULL);
      if (fit >= total) { exhausted = true; break; }
      const long long o = fit / (3 * PER_BAND); const int rem = (int)(fit % (3 * PER_BAND));
      const int j = rem / PER_BAND, q = rem % PER_BAND;
      const int band = (j == 0) ? BG : (j == 1 ? BR : BI);
      const long long lo = seg_off[o * N_SEG + band];
      const int n = (int)(seg_off[o * N_SEG + band + 1] - lo);
      int code;
      active = job.init(gt + lo, gf + lo, n, (KIND == 0) ? decline_index(q) : 0.0, code);
      if (!active) {
        const long long s = slot(fit);
        out[s] = qnan();
        if (status) { status[s * 2] = code; status[s * 2 + 1] = 0; }
      }
    }
    if (!__any_sync(0xffffffffu, active || !exhausted)) break;
    if (active) {
      const int rc = job.step();
      if (rc != TRF_RUNNING) {
        const long long s = slot(fit);
        out[s] = job.finish(rc);
        if (status) { status[s * 2] = rc; status[s * 2 + 1] = job.trf.nfev; }
        active = false;
      }
    }
  }
}

template <class K>
int thread_fit_grid(K kernel, const mlrn_store* s, long long fits) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FIT_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long want = (fits + FIT_THREADS - 1) / FIT_THREADS;
  long long cap = (long long)s->sm_count * per_sm;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
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
  MLRN_CUDA(cudaMemsetAsync(s->ticket, 0, sizeof(unsigned long long), st));
  const int grid = thread_fit_grid(k_bazin_thread, s, (long long)s->n_obj * N_BAND);
  k_bazin_thread<<<grid, FIT_THREADS, 0, st>>>(s->seg_off, s->t, s->flux, s->err, s->n_obj, s->ticket, out_dev, status_dev);
  k_bazin_tail<<<(s->n_obj + 255) / 256, 256, 0, st>>>(s->n_obj, out_dev);
  MLRN_CUDA(cudaGetLastError());
  count_launch(2);
  return MLRN_OK;
}

int run_powerlaw(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  MLRN_CUDA(cudaMemsetAsync(s->ticket + 4, 0, 3 * sizeof(unsigned long long), st));
  k_decline_thread<0><<<thread_fit_grid(k_decline_thread<0>, s, (long long)s->n_obj * 21), FIT_THREADS, 0, st>>>(
      s->seg_off, s->t, s->flux, s->n_obj, s->ticket + 4, out_dev, status_dev);
  k_decline_thread<1><<<thread_fit_grid(k_decline_thread<1>, s, (long long)s->n_obj * 3), FIT_THREADS, 0, st>>>(
      s->seg_off, s->t, s->flux, s->n_obj, s->ticket + 5, out_dev, status_dev);
  k_decline_thread<2><<<thread_fit_grid(k_decline_thread<2>, s, (long long)s->n_obj * 3), FIT_THREADS, 0, st>>>(
      s->seg_off, s->t, s->flux, s->n_obj, s->ticket + 6, out_dev, status_dev);
  MLRN_CUDA(cudaGetLastError());
  count_launch(3);
  return MLRN_OK;
}

}  // namespace mlrn
