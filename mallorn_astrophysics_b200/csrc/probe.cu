// probe.cu -- mlrn_measure_fp64: the FP64 roofline denominator, measured on the device the bench runs on.
//
// SURVEY.md 8(d) asks for the GP pass to be reported against an FP64 peak "measured with a DFMA micro-benchmark at bench
// time", not a datasheet number (MEASURED_PEAKS.json holds only HBM and bf16 figures).  Two register-resident loops, one of
// dependent-chain-free DFMAs and one of mma.sync.m8n8k4.f64 (the DMMA the GP sweep issues), 16 warps per SM on every SM,
// timed with CUDA events; the result is 2 * FMA / s.
#include <cuda_runtime.h>
#include "store.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

__global__ void __launch_bounds__(512) k_probe_dfma(double* out, int iters) {
  double a[8], x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) k_probe_dmma(double* out, int iters) {
  double c[8][2], a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace
}  // namespace mlrn

extern "C" int mlrn_measure_fp64(int device, double* dfma_tflops, double* dmma_tflops) {
  using namespace mlrn;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device visible"); return MLRN_ENODEVICE; }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return MLRN_EINVAL; }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double* out = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(double) * (size_t)sms * 512);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  const int iters = 20000;
  double res[2] = {0.0, 0.0};
  for (int which = 0; which < 2 && e == cudaSuccess; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 3 && e == cudaSuccess; ++rep) {       // first repetition warms up; best of the rest
      cudaEventRecord(e0);
      if (which == 0) k_probe_dfma<<<sms, 512>>>(out, iters); else k_probe_dmma<<<sms, 512>>>(out, iters);
      cudaEventRecord(e1);
      e = cudaEventSynchronize(e1);
      float ms = 0.f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    count_launch(3);
    const double fma = which == 0 ? 512.0 * iters * 8 : 16.0 * iters * 8 * 256;     // per SM
    res[which] = 2.0 * fma * sms / (best * 1e-3) / 1e12;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (out) cudaFree(out);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return cuda_fail(e, "mlrn_measure_fp64");
  if (dfma_tflops) *dfma_tflops = res[0];
  if (dmma_tflops) *dmma_tflops = res[1];
  return MLRN_OK;
}
