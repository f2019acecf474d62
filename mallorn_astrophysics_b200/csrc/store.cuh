// store.cuh -- device-resident packed light-curve store (opaque `mlrn_store` of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

struct mlrn_store {
  int device = 0;
  int64_t n_rows = 0;       // rows kept (obj >= 0)
  int32_t n_obj = 0;
  int32_t max_obs = 0;      // largest per-object count
  int32_t max_seg = 0;      // largest (object, band) segment
  int32_t max_short = 0;    // largest object among those with at most SHORT_LIMIT observations
  int32_t n_long = 0;       // objects above SHORT_LIMIT ...
  int* long_list = nullptr; // ... and their indices [n_long] (order unspecified)
  int32_t sm_count = 148;
  long long* seg_off = nullptr;     // [n_obj*7 + 1]
  double* t = nullptr;              // [n_rows]
  double* flux = nullptr;
  double* err = nullptr;
  unsigned short* rank = nullptr;   // [n_rows] position of the row inside its object in FILE order
  unsigned short* trank = nullptr;  // [n_rows] position inside its object in (time, file rank) order over ALL bands
  int32_t* seg_meta = nullptr;      // [n_obj*22]
  unsigned long long* ticket = nullptr;   // [8] work counters of the GP passes (one per launch)
  cudaStream_t stream = nullptr;          // the stream the store was packed on: its arrays are stream-ordered allocations of it
};

namespace mlrn {
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);   // every kernel launch of the library is counted (mlrn_launch_count)
}  // namespace mlrn

#define MLRN_CUDA(call)                                                   \
  do {                                                                    \
    cudaError_t _e = (call);                                              \
    if (_e != cudaSuccess) return mlrn::cuda_fail(_e, #call);             \
  } while (0)
