// Developer micro-benchmark (GPU box): FP64 FMA-pipe rate vs mma.sync.m8n8k4.f64 (DMMA) rate per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_rates.cu -o fp64_rates && ./fp64_rates
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
  double a[8], x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
  double c[8][2], a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int which = 0; which < 2; ++which) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_dfma<<<sms, warps * 32>>>(out, iters); else k_dmma<<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      // FMAs per SM: dfma: warps*32 threads * iters * 8 ; dmma: warps * iters * 8 * 256
      double fma_per_sm = (which == 0) ? (double)warps * 32 * iters * 8 : (double)warps * iters * 8 * 256;
      double cycles = ms * 1e-3 * clk * 1e3;
      printf("%s warps/SM %2d: %.3f ms, %.1f FMA/clk/SM (at %d MHz nominal), %.1f TFLOP/s chip\n", which ? "DMMA" : "DFMA", warps, ms,
             fma_per_sm / cycles, clk / 1000, 2.0 * fma_per_sm * sms / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
