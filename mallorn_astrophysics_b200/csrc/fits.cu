// fits.cu -- K4 kernels: Bazin (52 columns) and decline-model (27 columns) passes, one thread per fit.
#include <cuda_runtime.h>
#include "store.cuh"
#include "fits.cuh"
#include "fits_thread.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

// ------------------------------------------------------------------------------------------------
// One THREAD per fit (trf_thread.cuh): every lane of a warp advances its own fit by one trial
// evaluation per trip, and a lane whose fit ends takes the next (object, band[, model]) from the
// ticket counter, so the 100x spread in iteration counts does not idle the warp.
// ------------------------------------------------------------------------------------------------
#ifndef FIT_THREADS_N
#define FIT_THREADS_N 512
#endif
#ifndef FIT_MIN_BLOCKS
#define FIT_MIN_BLOCKS 1
#endif
constexpr int FIT_THREADS = FIT_THREADS_N;     // warps of one CTA start every trip together (see the barrier in the loops)
// true while any lane of the group (CTA, or warp in the barrier-free developer build) still has work
__device__ __forceinline__ bool group_any(bool pred) {
#ifdef FIT_NO_BARRIER
  return __any_sync(0xffffffffu, pred);
#else
  return __syncthreads_or(pred);
#endif
}

__global__ void __launch_bounds__(FIT_THREADS, FIT_MIN_BLOCKS)
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
      const int b = (int)(fit / n_obj); const long long o = fit % n_obj;    // band-major: lanes of a warp get bands of similar length
      const long long lo = seg_off[o * N_SEG + b];
      const int n = (int)(seg_off[o * N_SEG + b + 1] - lo);
      int code;
      active = job.init(gt + lo, gf + lo, ge + lo, n, code);
      if (!active) {
        job.finish(code, out + o * BAZIN_COLS + 8 * b);
        if (status) { status[(o * N_BAND + b) * 2] = code; status[(o * N_BAND + b) * 2 + 1] = 0; }
      }
    }
    if (!group_any(active || !exhausted)) break;      // also lines the warps up: they fetch the same code
    if (active) {
      const int rc = job.step();
      if (rc != TRF_RUNNING) {
        const int b = (int)(fit / n_obj); const long long o = fit % n_obj;
        job.finish(rc, out + o * BAZIN_COLS + 8 * b);
        if (status) { status[(o * N_BAND + b) * 2] = rc; status[(o * N_BAND + b) * 2 + 1] = job.trf.nfev; }
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
__global__ void __launch_bounds__(FIT_THREADS, FIT_MIN_BLOCKS)
k_decline_thread(const long long* __restrict__ seg_off, const double* __restrict__ gt, const double* __restrict__ gf,
                 int n_obj, unsigned long long* __restrict__ ticket, double* __restrict__ out, int* __restrict__ status) {
  constexpr int PER_BAND = (KIND == 0) ? 7 : 1;
  const long long total = (long long)n_obj * 3 * PER_BAND;
  DeclineJob<KIND> job;
  bool active = false, exhausted = false;
  long long fit = 0;
  // (band, model)-major ticket order: the lanes of a warp fit the same model to the same band of different objects
  auto slot = [&](long long ft) {                     // column of this fit inside the 27-wide row
    const int rem = (int)(ft / n_obj); const long long o = ft % n_obj;
    const int j = rem / PER_BAND, q = (KIND == 0) ? rem % PER_BAND : (KIND == 1 ? 7 : 8);
    return o * POWERLAW_COLS + 9 * j + q;
  };
  for (;;) {
    for (int tries = 0; tries < 4 && !active && !exhausted; ++tries) {
      fit = (long long)atomicAdd(ticket, 1ULL);
      if (fit >= total) { exhausted = true; break; }
      const int rem = (int)(fit / n_obj); const long long o = fit % n_obj;
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
    if (!group_any(active || !exhausted)) break;
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
