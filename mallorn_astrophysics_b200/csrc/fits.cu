// fits.cu -- K4 kernels: Bazin (52 columns) and decline-model (27 columns) passes, one thread per fit, run in STAGES.
//
// One THREAD per fit (trf_thread.cuh): every lane of a warp advances its own fit by one trial evaluation per trip, and a
// lane whose fit ends takes the next (object, band[, model]) from a ticket counter.  Function-evaluation counts spread
// 100x (median 34, mean 58, 1 % above 480, max_nfev = 2000), and a trip of a warp costs the union of its lanes' paths,
// so a fit that runs long after the ticket counter is exhausted keeps a whole warp -- and, through the CTA-wide trip
// barrier, a whole SM -- busy for one useful lane.  Round 1 measured that tail at 0.44 s per launch whatever the batch
// size (11 % of a 10^6-object launch, 90 % of a 20,000-object one).  A stage therefore only runs a fit up to a cap of
// function evaluations; a fit that is still running at the cap is written out whole (the solver is a resumable state
// machine, so the job struct is the complete state) and picked up by the next stage, which has a higher cap, far fewer
// fits, and -- once the survivors are fewer than the lanes -- spreads them one per warp, so that the stragglers no longer
// pay for each other's divergence.  Caps double from stage to stage; the last stage has none.
//
// Stragglers.  The last stages hold the fits whose every trip is a Jacobian pass over a long band (40-70 rows x 6 exp):
// ncu showed them at 1.04 threads per instruction and 224 us per trip -- the latency of ONE thread.  When a resumed stage
// finds fewer survivors than COOP_FITS_PER_WARP per warp of the grid, it runs them in the COOPERATIVE instantiation of
// the same solver instead (trf_thread.cuh, COOP = true): one fit per warp, every lane holding the same state, the model
// evaluations of the rows shared out over the lanes and folded in row order -- the same arithmetic, so which instantiation
// ran a fit cannot be seen in the result.
//
// Work order.  A trip of a warp costs the LONGEST of its lanes' row loops (Jacobian pass and trial evaluation both run over
// the rows of the fit), and in (band, object) order the lanes of a warp hold bands of 10-50 rows side by side -- decline
// fits, whose rows are the post-peak suffix, anything from 3 rows to the whole band.  Stage 0 therefore takes its fits from
// a list sorted by row count, longest first (counting sort: histogram, one-block scan, scatter): lanes that work side by
// side have the same loop lengths, and the long fits do not start last.  The order inside a bucket is whatever the atomics
// give; the result of a fit does not depend on when or where it runs.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdlib.h>
#include "store.cuh"
#include "fits.cuh"
#include "fits_thread.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

#ifndef FIT_THREADS_N
#define FIT_THREADS_N 512
#endif
#ifndef FIT_MIN_BLOCKS
#define FIT_MIN_BLOCKS 1
#endif
constexpr int FIT_THREADS = FIT_THREADS_N;     // warps of one CTA start every trip together (see the barrier in the loop)
constexpr int MAX_STAGES = 6;
#ifndef COOP_FITS_N
#define COOP_FITS_N 4
#endif
constexpr int COOP_FITS_PER_WARP = COOP_FITS_N;          // a resumed stage with at most this many survivors per warp runs them cooperatively

// a fit that outlived its stage: the whole resumable job and which fit it is
template <class Job> struct FitCarry { Job job; long long fit; };

struct StageIo {
  const double* pre;                // stage 0: [ticket][Job::PRE] start values from the prologue pass; nullptr: computed in place
  const unsigned* order;            // stage 0: ticket -> fit (sorted by row count, longest first); nullptr: ticket = fit
  unsigned long long* ticket;        // this stage's work counter
  int cap;                           // function evaluations after which a running fit moves on (INT_MAX: never)
  const void* in; const unsigned* in_count; unsigned in_capacity;       // carried fits to resume (nullptr: fresh fits)
  void* out; unsigned* out_count; unsigned out_capacity;                // where survivors go (nullptr: last stage)
};

// ---- Bazin: fit = band * n_obj + object (band-major: the lanes of a warp get bands of similar length)
template <bool COOP>
struct BazinPolicyT {
  using Job = BazinJobT<COOP>;
  using Twin = BazinPolicyT<!COOP>;
  const long long* seg_off; const double *gt, *gf, *ge; int n_obj; double* out; int* status;
  __device__ long long total() const { return (long long)n_obj * N_BAND; }
  __device__ int rows(long long fit) const {                            // rows of the fit (work-order key)
    const int b = (int)(fit / n_obj); const long long o = fit % n_obj;
    return (int)(seg_off[o * N_SEG + b + 1] - seg_off[o * N_SEG + b]);
  }
  __device__ void prologue(Job& job, long long fit, double* pre_out) const {      // the start values of the fit, nothing else
    const int b = (int)(fit / n_obj); const long long o = fit % n_obj;
    const long long lo = seg_off[o * N_SEG + b];
    int code;
    job.init(gt + lo, gf + lo, ge + lo, (int)(seg_off[o * N_SEG + b + 1] - lo), code, nullptr, pre_out);
  }
  __device__ bool start(Job& job, long long fit, const double* pre) const {
    const int b = (int)(fit / n_obj); const long long o = fit % n_obj;
    const long long lo = seg_off[o * N_SEG + b];
    const int n = (int)(seg_off[o * N_SEG + b + 1] - lo);
    int code;
    if (job.init(gt + lo, gf + lo, ge + lo, n, code, pre)) return true;
    job.finish(code, out + o * BAZIN_COLS + 8 * b);
    if (status) { status[(o * N_BAND + b) * 2] = code; status[(o * N_BAND + b) * 2 + 1] = 0; }
    return false;
  }
  __device__ void done(const Job& job, long long fit, int rc) const {
    const int b = (int)(fit / n_obj); const long long o = fit % n_obj;
    job.finish(rc, out + o * BAZIN_COLS + 8 * b);
    if (status) { status[(o * N_BAND + b) * 2] = rc; status[(o * N_BAND + b) * 2 + 1] = job.trf.nfev; }
  }
};

// ---- decline models.  KIND 0: the seven fixed-index power laws (21 fits per object); 1: exponential; 2: linear (3 per
// object each); (band, model)-major fit order: the lanes of a warp fit the same model to the same band of different objects
template <int KIND, bool COOP>
struct DeclinePolicyT {
  using Job = DeclineJobT<KIND, COOP>;
  using Twin = DeclinePolicyT<KIND, !COOP>;
  static constexpr int PER_BAND = (KIND == 0) ? 7 : 1;
  const long long* seg_off; const double *gt, *gf; int n_obj; double* out; int* status; const int* meta;
  __device__ long long total() const { return (long long)n_obj * 3 * PER_BAND; }
  __device__ int rows(long long fit) const {                            // observations after the peak (estimate: ties in time ignored)
    const int rem = (int)(fit / n_obj); const long long o = fit % n_obj;
    const int j = rem / PER_BAND;
    const int band = (j == 0) ? BG : (j == 1 ? BR : BI);
    const int n = meta[o * 22 + 3 * band], pk = meta[o * 22 + 3 * band + 2];
    return (n < 5 || pk < 0) ? 0 : n - pk - 1;
  }
  __device__ long long slot(long long fit) const {                      // column of this fit inside the 27-wide row
    const int rem = (int)(fit / n_obj); const long long o = fit % n_obj;
    const int j = rem / PER_BAND, q = (KIND == 0) ? rem % PER_BAND : (KIND == 1 ? 7 : 8);
    return o * POWERLAW_COLS + 9 * j + q;
  }
  __device__ void prologue(Job& job, long long fit, double* pre_out) const {
    const int rem = (int)(fit / n_obj); const long long o = fit % n_obj;
    const int j = rem / PER_BAND, q = rem % PER_BAND;
    const int band = (j == 0) ? BG : (j == 1 ? BR : BI);
    const long long lo = seg_off[o * N_SEG + band];
    int code;
    job.init(gt + lo, gf + lo, (int)(seg_off[o * N_SEG + band + 1] - lo), (KIND == 0) ? decline_index(q) : 0.0, code, nullptr, pre_out);
  }
  __device__ bool start(Job& job, long long fit, const double* pre) const {
    const int rem = (int)(fit / n_obj); const long long o = fit % n_obj;
    const int j = rem / PER_BAND, q = rem % PER_BAND;
    const int band = (j == 0) ? BG : (j == 1 ? BR : BI);
    const long long lo = seg_off[o * N_SEG + band];
    const int n = (int)(seg_off[o * N_SEG + band + 1] - lo);
    int code;
    if (job.init(gt + lo, gf + lo, n, (KIND == 0) ? decline_index(q) : 0.0, code, pre)) return true;
    const long long s = slot(fit);
    out[s] = qnan();
    if (status) { status[s * 2] = code; status[s * 2 + 1] = 0; }
    return false;
  }
  __device__ void done(const Job& job, long long fit, int rc) const {
    const long long s = slot(fit);
    out[s] = job.finish(rc);
    if (status) { status[s * 2] = rc; status[s * 2 + 1] = job.trf.nfev; }
  }
};

template <bool C> BazinPolicyT<!C> twin(const BazinPolicyT<C>& p) { return {p.seg_off, p.gt, p.gf, p.ge, p.n_obj, p.out, p.status}; }
template <int K, bool C> DeclinePolicyT<K, !C> twin(const DeclinePolicyT<K, C>& p) { return {p.seg_off, p.gt, p.gf, p.n_obj, p.out, p.status, p.meta}; }

// ---- work order: counting sort of the fits by row count, longest first
constexpr int ORDER_BINS = 1024;
template <class P>
__global__ void __launch_bounds__(256) k_fit_hist(P pol, unsigned* __restrict__ hist) {
  __shared__ unsigned sh[ORDER_BINS];
  for (int i = threadIdx.x; i < ORDER_BINS; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long long total = pol.total();
  for (long long fit = blockIdx.x * (long long)blockDim.x + threadIdx.x; fit < total; fit += (long long)gridDim.x * blockDim.x) {
    const int r = pol.rows(fit);
    atomicAdd(&sh[r < ORDER_BINS ? r : ORDER_BINS - 1], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ORDER_BINS; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// cursor[k] = number of fits with more rows than k (one block of ORDER_BINS threads)
__global__ void __launch_bounds__(ORDER_BINS) k_fit_scan(const unsigned* __restrict__ hist, unsigned* __restrict__ cursor) {
  __shared__ unsigned sh[ORDER_BINS];
  const int k = threadIdx.x;
  sh[k] = hist[ORDER_BINS - 1 - k];                    // descending keys
  __syncthreads();
  for (int d = 1; d < ORDER_BINS; d <<= 1) {
    const unsigned v = (k >= d) ? sh[k - d] : 0u;
    __syncthreads();
    sh[k] += v;
    __syncthreads();
  }
  cursor[ORDER_BINS - 1 - k] = sh[k] - hist[ORDER_BINS - 1 - k];
}
template <class P>
__global__ void __launch_bounds__(256) k_fit_scatter(P pol, unsigned* __restrict__ cursor, unsigned* __restrict__ order) {
  const long long total = pol.total();
  for (long long fit = blockIdx.x * (long long)blockDim.x + threadIdx.x; fit < total; fit += (long long)gridDim.x * blockDim.x) {
    const int r = pol.rows(fit);
    order[atomicAdd(&cursor[r < ORDER_BINS ? r : ORDER_BINS - 1], 1u)] = (unsigned)fit;
  }
}

// COOP = false: one thread per fit.  COOP = true: one WARP per fit (resumed stages only); `coop_limit`: the largest survivor
// count the cooperative launch takes -- the two launches of a resumed stage read the same count and exactly one of them works.
// PROLOGUE.  A fit starts with two passes over its band that only the lane holding it can run while the other 31 wait in
// the middle of their own fits -- the median of the fluxes (rank counting, O(n^2)) and the cost at the start (n model
// evaluations): 5 % of the warp instructions of the stage-0 kernel at ONE thread per instruction.  Here they are computed
// for every fit up front, one thread per ticket (a warp's lanes have the same row count in the work order), and stage 0
// reads Job::PRE doubles per ticket instead.  Same functions on the same data: the values are bit-identical.  Bazin at 10^6
// objects 2.33 -> 2.26 s; not used for the decline models (Job::PRE = 0: their start costs one short pass).
template <class P>
__global__ void __launch_bounds__(128) k_fit_prologue(P pol, const unsigned* __restrict__ order, double* __restrict__ pre) {
  const long long total = pol.total();
  for (long long tk = blockIdx.x * (long long)blockDim.x + threadIdx.x; tk < total; tk += (long long)gridDim.x * blockDim.x) {
    typename P::Job job;
    pol.prologue(job, order ? (long long)order[tk] : tk, pre + tk * P::Job::PRE);
  }
}

#ifdef FIT_STATS
__device__ unsigned long long g_fit_stats[8];   // warp trips, active lanes, lanes with a Jacobian pass, rows (sum, max), rows of the passes (sum, max), trips with a pass
#endif
template <class P, bool COOP>
__global__ void __launch_bounds__(FIT_THREADS, FIT_MIN_BLOCKS) k_fit_stage(P pol, StageIo io, long long coop_limit) {
  using Job = typename P::Job;
  using Carry = FitCarry<Job>;
  const Carry* in = (const Carry*)io.in;
  Carry* outc = (Carry*)io.out;
  const int lane = threadIdx.x & 31;
  long long total;
  int lanes = 32;                                     // lanes of a warp that take work (thread mode)
  if (in) {
    const unsigned c = *io.in_count;
    total = c < io.in_capacity ? c : io.in_capacity;
    if (COOP ? (total > coop_limit) : (total <= coop_limit)) return;      // the other instantiation's stage
    const long long warps = (long long)gridDim.x * (FIT_THREADS / 32);
    const long long per_warp = (total + warps - 1) / warps;      // fewer survivors than lanes: spread them over the warps
    lanes = per_warp < 1 ? 1 : (per_warp > 32 ? 32 : (int)per_warp);
  } else {
    total = pol.total();
  }
  Job job;
  bool active = false, exhausted = !COOP && lane >= lanes, may_export = outc != nullptr;
  long long fit = 0;
  for (;;) {
    for (int tries = 0; tries < 4 && !active && !exhausted; ++tries) {
      long long tk = 0;
      if (!COOP || lane == 0) tk = (long long)atomicAdd(io.ticket, 1ULL);
      if (COOP) tk = __shfl_sync(0xffffffffu, tk, 0);               // one ticket per warp
      if (tk >= total) { exhausted = true; break; }
      if (in) { job = in[tk].job; fit = in[tk].fit; active = true; }
      else { fit = io.order ? (long long)io.order[tk] : tk; active = pol.start(job, fit, (Job::PRE > 0 && io.pre) ? io.pre + tk * Job::PRE : nullptr); }
    }
    if (!__syncthreads_or(active || !exhausted)) break;      // also lines the warps up: they fetch the same code
#ifdef FIT_STATS
    if (!COOP) {                                       // developer build: where do the lanes of a trip stand?
      const bool nj = active && job.trf.need_jac;
      const int mm = active ? job.trf.m : 0;
      const int n_act = __popc(__ballot_sync(0xffffffffu, active)), n_jac = __popc(__ballot_sync(0xffffffffu, nj));
      const int sum_m = __reduce_add_sync(0xffffffffu, mm), max_m = __reduce_max_sync(0xffffffffu, mm);
      const int sum_mj = __reduce_add_sync(0xffffffffu, nj ? mm : 0), max_mj = __reduce_max_sync(0xffffffffu, nj ? mm : 0);
      if (lane == 0 && n_act > 0) {
        atomicAdd(&g_fit_stats[0], 1ULL); atomicAdd(&g_fit_stats[1], (unsigned long long)n_act); atomicAdd(&g_fit_stats[2], (unsigned long long)n_jac);
        atomicAdd(&g_fit_stats[3], (unsigned long long)sum_m); atomicAdd(&g_fit_stats[4], (unsigned long long)max_m);
        atomicAdd(&g_fit_stats[5], (unsigned long long)sum_mj); atomicAdd(&g_fit_stats[6], (unsigned long long)max_mj);
        if (n_jac > 0) atomicAdd(&g_fit_stats[7], 1ULL);
      }
    }
#endif
    if (active) {
      const int rc = job.step();                      // COOP: every lane advances the same fit (uniform control flow)
      if (rc != TRF_RUNNING) {
        if (!COOP || lane == 0) pol.done(job, fit, rc);
        active = false;
      } else if (may_export && job.trf.nfev >= io.cap) {
        unsigned slot = 0;
        if (!COOP || lane == 0) slot = atomicAdd(io.out_count, 1u);
        if (COOP) slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot < io.out_capacity) {
          if (!COOP || lane == 0) { outc[slot].job = job; outc[slot].fit = fit; }
          active = false;
        } else may_export = false;                    // carry buffer full: finish here
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

// Runs every fit of `pol` through the stages.  caps[0..n_caps) are the evaluation caps of the capped stages; one more,
// uncapped stage follows.  carry_frac: size of the carry buffers as a fraction of the fits (a full buffer is not an error:
// the fits that do not find room finish in the stage they are in).
template <class P>
int run_staged(const mlrn_store* s, const P& pol, long long fits, const int* caps, int n_caps, double carry_frac, cudaStream_t st) {
  using Carry = FitCarry<typename P::Job>;
  using PC = typename P::Twin;                         // the cooperative instantiation: same members, same carry layout
  static_assert(sizeof(FitCarry<typename PC::Job>) == sizeof(Carry), "carry records of the two instantiations must be interchangeable");
  const PC polc = twin(pol);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fit_stage<P, false>, FIT_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  const long long full = (long long)s->sm_count * per_sm;
  long long want = (fits + FIT_THREADS - 1) / FIT_THREADS;
  const int grid0 = (int)(want < full ? (want > 0 ? want : 1) : full);
  long long cap_l = (long long)(carry_frac * (double)fits);
  if (cap_l < 8192) cap_l = 8192;
  if (cap_l > (4LL << 20)) cap_l = 4LL << 20;         // at most 4 Mi records per buffer (~4 GB for the Bazin job)
  if (cap_l > fits) cap_l = fits;
  const unsigned capacity = (unsigned)(cap_l < 0xffffffffLL ? cap_l : 0xffffffffLL);
  unsigned long long* ctr = nullptr;                  // [MAX_STAGES] tickets, then [MAX_STAGES] carry counts (as unsigned)
  Carry* buf[2] = {nullptr, nullptr};
  unsigned* order = nullptr;                          // [fits] the work order of stage 0, then [ORDER_BINS] histogram + [ORDER_BINS] cursors
  MLRN_CUDA(cudaMallocAsync(&ctr, sizeof(unsigned long long) * 2 * MAX_STAGES, st));
  cudaError_t e = cudaMemsetAsync(ctr, 0, sizeof(unsigned long long) * 2 * MAX_STAGES, st);
  if (e == cudaSuccess && fits > 64 && fits < 0xffffffffLL && !getenv("MLRN_FIT_UNORDERED")) {
    e = cudaMallocAsync(&order, sizeof(unsigned) * ((size_t)fits + 2 * ORDER_BINS), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(order + fits, 0, sizeof(unsigned) * 2 * ORDER_BINS, st);
    if (e == cudaSuccess) {
      long long blocks = (fits + 256 * 8 - 1) / (256 * 8);
      if (blocks > s->sm_count * 8) blocks = s->sm_count * 8;
      k_fit_hist<P><<<(int)blocks, 256, 0, st>>>(pol, order + fits);
      k_fit_scan<<<1, ORDER_BINS, 0, st>>>(order + fits, order + fits + ORDER_BINS);
      k_fit_scatter<P><<<(int)blocks, 256, 0, st>>>(pol, order + fits + ORDER_BINS, order);
      count_launch(3);
      e = cudaGetLastError();
    }
  }
  double* pre = nullptr;                              // [fits][Job::PRE] start values of every fit (prologue pass)
  if (e == cudaSuccess && P::Job::PRE > 0 && fits > 64 && !getenv("MLRN_FIT_NO_PROLOGUE")) {
    e = cudaMallocAsync(&pre, sizeof(double) * (size_t)fits * P::Job::PRE, st);
    if (e == cudaSuccess) {
      long long blocks = (fits + 127) / 128;
      if (blocks > (long long)s->sm_count * 16) blocks = (long long)s->sm_count * 16;
      k_fit_prologue<P><<<(int)blocks, 128, 0, st>>>(pol, order, pre);
      count_launch();
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess && n_caps > 0) e = cudaMallocAsync(&buf[0], sizeof(Carry) * (size_t)capacity, st);
  if (e == cudaSuccess && n_caps > 1) e = cudaMallocAsync(&buf[1], sizeof(Carry) * (size_t)capacity, st);
  if (e == cudaSuccess) {
    unsigned* counts = (unsigned*)(ctr + MAX_STAGES);
    for (int stage = 0; stage <= n_caps; ++stage) {
      StageIo io;
      io.order = stage == 0 ? order : nullptr;
      io.pre = stage == 0 ? pre : nullptr;
      io.ticket = ctr + stage;
      io.cap = stage < n_caps ? caps[stage] : INT_MAX;
      io.in = stage > 0 ? buf[(stage - 1) & 1] : nullptr;
      io.in_count = stage > 0 ? counts + 2 * (stage - 1) : nullptr;
      io.in_capacity = capacity;
      io.out = stage < n_caps ? buf[stage & 1] : nullptr;
      io.out_count = stage < n_caps ? counts + 2 * stage : nullptr;
      io.out_capacity = capacity;
      // resumed stages: the survivor count is only known on the device, so they launch the full persistent grid (every
      // SM: the survivors are spread over as many warps as there are, a warp's trip costs the union of its lanes' paths)
      // and every warp reads the count; a warp without work leaves at once
      const long long grid = stage == 0 ? grid0 : full;
      const long long coop_limit = stage == 0 ? -1 : full * (FIT_THREADS / 32) * COOP_FITS_PER_WARP;
      k_fit_stage<P, false><<<(int)grid, FIT_THREADS, 0, st>>>(pol, io, coop_limit);
      count_launch();
      if (stage > 0) {
        k_fit_stage<PC, true><<<(int)full, FIT_THREADS, 0, st>>>(polc, io, coop_limit);
        count_launch();
      }
    }
    e = cudaGetLastError();
  }
  if (buf[0]) cudaFreeAsync(buf[0], st);
  if (buf[1]) cudaFreeAsync(buf[1], st);
  if (order) cudaFreeAsync(order, st);
  if (pre) cudaFreeAsync(pre, st);
  cudaFreeAsync(ctr, st);
  if (e != cudaSuccess) return cuda_fail(e, "staged fit launch");
  return MLRN_OK;
}

#ifdef FIT_STATS
}  // namespace
extern "C" void mlrn_fit_stats(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_fit_stats, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_fit_stats, z, sizeof(z)); }
}
namespace {
#endif

#ifndef FIT_STAGE_CAPS
// swept on the final kernels at 10^6 / 10,178 objects (ms): {128,256,512,1024} 2276 / 170, {128,512} 2229 / 166,
// {64,256,1024} 2426 / 407, {256,1024} 2286 / 173, {96,192,384,768} 2212 / 171
#define FIT_STAGE_CAPS {96, 192, 384, 768}
#endif
#ifndef FIT_DECLINE_CAPS
#if defined(FIT_DECLINE_SET) && FIT_DECLINE_SET == 1      // developer sweeps (nvcc splits -D values at commas)
#define FIT_DECLINE_CAPS {48, 192}
#elif defined(FIT_DECLINE_SET) && FIT_DECLINE_SET == 2
#define FIT_DECLINE_CAPS {96, 384}
#elif defined(FIT_DECLINE_SET) && FIT_DECLINE_SET == 3
#define FIT_DECLINE_CAPS {48, 96, 192, 384}
#else
#define FIT_DECLINE_CAPS {64, 256}
#endif
#endif

}  // namespace

int run_bazin(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  const int caps[] = FIT_STAGE_CAPS;
  BazinPolicyT<false> pol{s->seg_off, s->t, s->flux, s->err, s->n_obj, out_dev, status_dev};
  const int rc = run_staged(s, pol, (long long)s->n_obj * N_BAND, caps, (int)(sizeof(caps) / sizeof(int)), 0.15, st);
  if (rc != MLRN_OK) return rc;
  k_bazin_tail<<<(s->n_obj + 255) / 256, 256, 0, st>>>(s->n_obj, out_dev);
  MLRN_CUDA(cudaGetLastError());
  count_launch();
  return MLRN_OK;
}

int run_powerlaw(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st) {
  if (s->n_obj == 0) return MLRN_OK;
  const int caps[] = FIT_DECLINE_CAPS;
  const int n_caps = (int)(sizeof(caps) / sizeof(int));
  int rc = run_staged(s, DeclinePolicyT<0, false>{s->seg_off, s->t, s->flux, s->n_obj, out_dev, status_dev, s->seg_meta}, (long long)s->n_obj * 21, caps, n_caps, 0.10, st);
  if (rc == MLRN_OK) rc = run_staged(s, DeclinePolicyT<1, false>{s->seg_off, s->t, s->flux, s->n_obj, out_dev, status_dev, s->seg_meta}, (long long)s->n_obj * 3, caps, n_caps, 0.10, st);
  if (rc == MLRN_OK) rc = run_staged(s, DeclinePolicyT<2, false>{s->seg_off, s->t, s->flux, s->n_obj, out_dev, status_dev, s->seg_meta}, (long long)s->n_obj * 3, caps, n_caps, 0.10, st);
  return rc;
}

}  // namespace mlrn
