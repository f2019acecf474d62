// api.cu -- the C-ABI of libmallorn_b200.so (include/mallorn_b200.h).  Thin: argument checks,
// host<->device staging, stream plumbing.  All arithmetic lives in the kernels.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges per C-ABI entry (tracing), visible in nsys / ncu timelines
#include <atomic>
#include <string>
#include <new>
#include "store.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return MLRN_ECUDA;
}
int pack_device(mlrn_store* s, int64_t n_rows, const int32_t* obj, const uint8_t* band, const double* t,
                const double* flux, const double* err, cudaStream_t st);
int run_stat(const mlrn_store* s, double* out_dev, cudaStream_t st);
int run_color(const mlrn_store* s, double* out_dev, cudaStream_t st);
int run_shape(const mlrn_store* s, double* out_dev, cudaStream_t st);
int run_tde(const mlrn_store* s, double* out_dev, cudaStream_t st);
int run_physics(const mlrn_store* s, const double* z_dev, double* out_dev, cudaStream_t st);
int run_bazin(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st);
int run_powerlaw(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st);
int run_gp2d(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st);
int run_gp1d(const mlrn_store* s, double* out_dev, int* status_dev, cudaStream_t st);
int run_research(const mlrn_store* s, const double* z_dev, double* out_dev, cudaStream_t st);
}  // namespace mlrn

using mlrn::set_error;

namespace {
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Runs `pass(out_dev)`; if the caller's buffer is host memory, stages through a device buffer.
template <class F>
int with_output(const mlrn_store* s, double* out, int out_on_device, int cols, cudaStream_t st, F pass) {
  if (!s || !out) { set_error("null store or output"); return MLRN_EINVAL; }
  DevGuard g(s->device);
  const size_t bytes = sizeof(double) * (size_t)s->n_obj * cols;
  if (out_on_device) return pass(out);
  if (bytes == 0) return MLRN_OK;
  double* d = nullptr;
  MLRN_CUDA(cudaMallocAsync(&d, bytes, st));
  int rc = pass(d);
  if (rc == MLRN_OK) {
    cudaError_t e = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = mlrn::cuda_fail(e, "copy of feature rows to host");
  }
  cudaFreeAsync(d, st);
  return rc;
}

// Same for a pass that also fills an optional int32 status table of `scols` values per object.
template <class F>
int with_status(const mlrn_store* s, double* out, int32_t* status, int out_on_device, int cols, int scols,
                cudaStream_t st, F pass) {
  if (!s) { set_error("null store"); return MLRN_EINVAL; }
  if (out_on_device || !status)
    return with_output(s, out, out_on_device, cols, st, [&](double* d) { return pass(d, status); });
  DevGuard g(s->device);
  const size_t sbytes = sizeof(int32_t) * (size_t)s->n_obj * scols;
  if (sbytes == 0) return MLRN_OK;
  int32_t* ds = nullptr;
  MLRN_CUDA(cudaMallocAsync(&ds, sbytes, st));
  int rc = with_output(s, out, 0, cols, st, [&](double* d) { return pass(d, ds); });
  if (rc == MLRN_OK) {
    cudaError_t e = cudaMemcpyAsync(status, ds, sbytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = mlrn::cuda_fail(e, "copy of the status table to host");
  }
  cudaFreeAsync(ds, st);
  return rc;
}

}  // namespace

extern "C" {

const char* mlrn_last_error(void) { return mlrn::g_err.c_str(); }
int mlrn_version(void) { return 100; }
long long mlrn_launch_count(void) { return mlrn::g_launches.load(std::memory_order_relaxed); }
int mlrn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mlrn_store_create(int device, int64_t n_rows, int32_t n_obj, const int32_t* obj, const uint8_t* band,
                      const double* t, const double* flux, const double* err, int inputs_on_device,
                      void* stream, mlrn_store** out) {
  NvtxRange nvtx_("mlrn_store_create");
  if (!out) { set_error("out is null"); return MLRN_EINVAL; }
  *out = nullptr;
  if (n_rows < 0 || n_obj < 0 || (n_rows > 0 && (!obj || !band || !t || !flux || !err))) {
    set_error("bad sizes or null column"); return MLRN_EINVAL;
  }
  if (n_rows > 2147483647LL) { set_error("more than 2^31-1 rows in one batch"); return MLRN_EINVAL; }
  int ndev = mlrn_device_count();
  if (ndev == 0) { set_error("no CUDA device visible; libmallorn_b200 has no CPU fallback"); return MLRN_ENODEVICE; }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return MLRN_EINVAL; }
  DevGuard g(device);
  cudaStream_t st = (cudaStream_t)stream;
  {
    // Temporaries (input staging, per-pass output staging, pack scratch) are stream-ordered allocations; keep the
    // pool's memory cached across synchronisation points instead of returning it to the driver after every call.
    static std::atomic<unsigned> pool_ready{0};
    if (device < 32 && !(pool_ready.load() & (1u << device))) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      pool_ready.fetch_or(1u << device);
    }
  }
  mlrn_store* s = new (std::nothrow) mlrn_store();
  if (!s) { set_error("out of host memory"); return MLRN_EINVAL; }
  s->device = device; s->n_obj = n_obj;
  cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
  const int32_t* d_obj = obj; const uint8_t* d_band = band; const double *d_t = t, *d_f = flux, *d_e = err;
  void* tmp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int rc = MLRN_OK;
  if (!inputs_on_device && n_rows > 0) {
    const size_t n = (size_t)n_rows;
    const size_t sz[5] = {n * 4, n, n * 8, n * 8, n * 8};
    const void* src[5] = {obj, band, t, flux, err};
    for (int k = 0; k < 5 && rc == MLRN_OK; ++k) {
      cudaError_t e = cudaMallocAsync(&tmp[k], sz[k], st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(tmp[k], src[k], sz[k], cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) rc = mlrn::cuda_fail(e, "host->device copy of the input columns");
    }
    d_obj = (const int32_t*)tmp[0]; d_band = (const uint8_t*)tmp[1];
    d_t = (const double*)tmp[2]; d_f = (const double*)tmp[3]; d_e = (const double*)tmp[4];
  }
  if (rc == MLRN_OK) rc = mlrn::pack_device(s, n_rows, d_obj, d_band, d_t, d_f, d_e, st);
  for (int k = 0; k < 5; ++k) if (tmp[k]) cudaFreeAsync(tmp[k], st);
  if (rc != MLRN_OK) { mlrn_store_destroy(s); return rc; }
  *out = s;
  return MLRN_OK;
}

int mlrn_store_destroy(mlrn_store* s) {
  if (!s) return MLRN_OK;
  DevGuard g(s->device);
  // Passes may have been issued on other streams than the one the store was packed on: wait for the device (what cudaFree
  // did implicitly), then hand the arrays back to the pool, which keeps them for the next batch.
  cudaDeviceSynchronize();
  void* arrays[9] = {s->seg_off, s->t, s->flux, s->err, s->rank, s->trank, s->seg_meta, s->ticket, s->long_list};
  for (void* a : arrays) if (a) cudaFreeAsync(a, s->stream);
  cudaGetLastError();
  delete s;
  return MLRN_OK;
}

int mlrn_store_info(const mlrn_store* s, int64_t* n_rows, int32_t* n_obj, int32_t* max_obs) {
  if (!s) { set_error("null store"); return MLRN_EINVAL; }
  if (n_rows) *n_rows = s->n_rows;
  if (n_obj) *n_obj = s->n_obj;
  if (max_obs) *max_obs = s->max_obs;
  return MLRN_OK;
}

int mlrn_store_segments(const mlrn_store* s, int32_t* out_host) {
  if (!s || !out_host) { set_error("null argument"); return MLRN_EINVAL; }
  DevGuard g(s->device);
  MLRN_CUDA(cudaDeviceSynchronize());
  MLRN_CUDA(cudaMemcpy(out_host, s->seg_meta, sizeof(int32_t) * (size_t)s->n_obj * MLRN_SEGMETA_COLS, cudaMemcpyDeviceToHost));
  return MLRN_OK;
}

int mlrn_store_export(const mlrn_store* s, int64_t* seg_off, double* t, double* flux, double* err, uint16_t* rank,
                      uint16_t* trank) {
  if (!s) { set_error("null store"); return MLRN_EINVAL; }
  DevGuard g(s->device);
  MLRN_CUDA(cudaDeviceSynchronize());
  const size_t n = (size_t)s->n_rows;
  if (seg_off) MLRN_CUDA(cudaMemcpy(seg_off, s->seg_off, sizeof(int64_t) * ((size_t)s->n_obj * MLRN_N_SEG + 1), cudaMemcpyDeviceToHost));
  if (t && n) MLRN_CUDA(cudaMemcpy(t, s->t, 8 * n, cudaMemcpyDeviceToHost));
  if (flux && n) MLRN_CUDA(cudaMemcpy(flux, s->flux, 8 * n, cudaMemcpyDeviceToHost));
  if (err && n) MLRN_CUDA(cudaMemcpy(err, s->err, 8 * n, cudaMemcpyDeviceToHost));
  if (rank && n) MLRN_CUDA(cudaMemcpy(rank, s->rank, 2 * n, cudaMemcpyDeviceToHost));
  if (trank && n) MLRN_CUDA(cudaMemcpy(trank, s->trank, 2 * n, cudaMemcpyDeviceToHost));
  return MLRN_OK;
}

int mlrn_stat(const mlrn_store* s, double* out, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_stat");
  cudaStream_t st = (cudaStream_t)stream;
  return with_output(s, out, out_on_device, MLRN_STAT_COLS, st, [&](double* d) { return mlrn::run_stat(s, d, st); });
}

int mlrn_color(const mlrn_store* s, double* out, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_color");
  cudaStream_t st = (cudaStream_t)stream;
  return with_output(s, out, out_on_device, MLRN_COLOR_COLS, st, [&](double* d) { return mlrn::run_color(s, d, st); });
}
int mlrn_shape(const mlrn_store* s, double* out, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_shape");
  cudaStream_t st = (cudaStream_t)stream;
  return with_output(s, out, out_on_device, MLRN_SHAPE_COLS, st, [&](double* d) { return mlrn::run_shape(s, d, st); });
}
int mlrn_tde(const mlrn_store* s, double* out, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_tde");
  cudaStream_t st = (cudaStream_t)stream;
  return with_output(s, out, out_on_device, MLRN_TDE_COLS, st, [&](double* d) { return mlrn::run_tde(s, d, st); });
}
int mlrn_physics(const mlrn_store* s, const double* z, int z_on_device, double* out, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_physics");
  if (!s || !z) { set_error("null store or redshift column"); return MLRN_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  DevGuard g(s->device);
  const double* zd = z;
  double* ztmp = nullptr;
  if (!z_on_device && s->n_obj > 0) {
    MLRN_CUDA(cudaMallocAsync(&ztmp, sizeof(double) * (size_t)s->n_obj, st));
    const cudaError_t e = cudaMemcpyAsync(ztmp, z, sizeof(double) * (size_t)s->n_obj, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFreeAsync(ztmp, st); return mlrn::cuda_fail(e, "host->device copy of the redshift column"); }
    zd = ztmp;
  }
  int rc = with_output(s, out, out_on_device, MLRN_PHYSICS_COLS, st, [&](double* d) { return mlrn::run_physics(s, zd, d, st); });
  if (ztmp) cudaFreeAsync(ztmp, st);
  return rc;
}

int mlrn_research(const mlrn_store* s, const double* z, int z_on_device, double* out, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_research");
  if (!s) { set_error("null store"); return MLRN_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  DevGuard g(s->device);
  const double* zd = z;                                   // NULL = no redshifts: the luminosity columns are NaN
  double* ztmp = nullptr;
  if (z && !z_on_device && s->n_obj > 0) {
    MLRN_CUDA(cudaMallocAsync(&ztmp, sizeof(double) * (size_t)s->n_obj, st));
    const cudaError_t e = cudaMemcpyAsync(ztmp, z, sizeof(double) * (size_t)s->n_obj, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFreeAsync(ztmp, st); return mlrn::cuda_fail(e, "host->device copy of the redshift column"); }
    zd = ztmp;
  }
  int rc = with_output(s, out, out_on_device, MLRN_RESEARCH_COLS, st, [&](double* d) { return mlrn::run_research(s, zd, d, st); });
  if (ztmp) cudaFreeAsync(ztmp, st);
  return rc;
}

int mlrn_bazin(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_bazin");
  cudaStream_t st = (cudaStream_t)stream;
  return with_status(s, out, status, out_on_device, MLRN_BAZIN_COLS, 12, st,
                     [&](double* d, int32_t* ds) { return mlrn::run_bazin(s, d, ds, st); });
}
int mlrn_powerlaw(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_powerlaw");
  cudaStream_t st = (cudaStream_t)stream;
  return with_status(s, out, status, out_on_device, MLRN_POWERLAW_COLS, 54, st,
                     [&](double* d, int32_t* ds) { return mlrn::run_powerlaw(s, d, ds, st); });
}
int mlrn_fits(const mlrn_store* s, double* bazin_out, int32_t* bazin_status, double* powerlaw_out,
              int32_t* powerlaw_status, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_fits");
  if (!s || !bazin_out || !powerlaw_out) { set_error("null store or output"); return MLRN_EINVAL; }
  DevGuard g(s->device);
  cudaStream_t st = (cudaStream_t)stream, side = nullptr;
  if (!out_on_device) {                              // host buffers: device staging for all four tables, then the same call
    const size_t n = (size_t)s->n_obj;
    if (n == 0) return MLRN_OK;
    const size_t nb = sizeof(double) * n * MLRN_BAZIN_COLS, np = sizeof(double) * n * MLRN_POWERLAW_COLS;
    const size_t sb = bazin_status ? sizeof(int32_t) * n * 12 : 0, sp = powerlaw_status ? sizeof(int32_t) * n * 54 : 0;
    unsigned char* d = nullptr;
    const size_t off_p = nb, off_sb = nb + np, off_sp = nb + np + sb;       // doubles first: every table stays aligned
    MLRN_CUDA(cudaMallocAsync(&d, nb + np + sb + sp, st));
    int rc = mlrn_fits(s, (double*)d, sb ? (int32_t*)(d + off_sb) : nullptr, (double*)(d + off_p),
                       sp ? (int32_t*)(d + off_sp) : nullptr, 1, stream);
    if (rc == MLRN_OK) {
      cudaError_t e = cudaMemcpyAsync(bazin_out, d, nb, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(powerlaw_out, d + off_p, np, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && sb) e = cudaMemcpyAsync(bazin_status, d + off_sb, sb, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && sp) e = cudaMemcpyAsync(powerlaw_status, d + off_sp, sp, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = mlrn::cuda_fail(e, "mlrn_fits: copy of the rows to host");
    }
    cudaFreeAsync(d, st);
    return rc;
  }
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(fork, st);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(side, fork, 0);
  int rc = e == cudaSuccess ? MLRN_OK : mlrn::cuda_fail(e, "mlrn_fits: side stream");
  if (rc == MLRN_OK) {
    // Bazin first: its first stage is the longest kernel of the two passes and takes the SMs; the decline models fill
    // them as its stage-0 CTAs retire and its few stragglers go on alone
    rc = mlrn::run_bazin(s, bazin_out, bazin_status, st);
    const int rc2 = mlrn::run_powerlaw(s, powerlaw_out, powerlaw_status, side);
    if (rc == MLRN_OK) rc = rc2;
    e = cudaEventRecord(join, side);                 // joined even after an error: nothing of the side stream may outlive the call's ordering
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, join, 0);
    if (e != cudaSuccess && rc == MLRN_OK) rc = mlrn::cuda_fail(e, "mlrn_fits: join");
  }
  if (fork) cudaEventDestroy(fork);
  if (join) cudaEventDestroy(join);
  if (side) cudaStreamDestroy(side);                 // released by the runtime once its work has completed
  return rc;
}
int mlrn_gp2d(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_gp2d");
  cudaStream_t st = (cudaStream_t)stream;
  return with_status(s, out, status, out_on_device, MLRN_GP2D_COLS, 4, st,
                     [&](double* d, int32_t* ds) { return mlrn::run_gp2d(s, d, ds, st); });
}
int mlrn_gp1d(const mlrn_store* s, double* out, int32_t* status, int out_on_device, void* stream) {
  NvtxRange nvtx_("mlrn_gp1d");
  cudaStream_t st = (cudaStream_t)stream;
  return with_status(s, out, status, out_on_device, MLRN_GP1D_COLS, 12, st,
                     [&](double* d, int32_t* ds) { return mlrn::run_gp1d(s, d, ds, st); });
}

}  // extern "C"
