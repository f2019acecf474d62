// interactions.cu -- K9: the physics-motivated feature interactions of src/features/interactions.py:23-196
// (create_physics_interactions) as an element-wise pass over a DEVICE-RESIDENT feature table.
//
// The reference adds up to 41 columns to the merged feature frame, each a product / ratio / power of one to three existing
// columns, and only those whose inputs are present.  Which columns exist is decided on the host from the column names
// (features/interactions.py builds the op list); the kernel evaluates the ops row by row: one thread per (row, op), ops
// fastest, so the threads of a warp read the same table row (3.5 KB for the 434-column table, L1-resident after the first
// touch) and write consecutive outputs.  HBM-bound: one read of the rows that hold inputs, one write of the new columns.
#include <cuda_runtime.h>
#include "store.cuh"
#include "../../include/mallorn_b200.h"

namespace mlrn {
namespace {

__global__ void k_interactions(const double* __restrict__ table, long long n_rows, int ld, const mlrn_inter_op* __restrict__ ops,
                               int n_ops, double* __restrict__ out, int ld_out) {
  const long long total = n_rows * n_ops;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / n_ops;
    const int k = (int)(idx - row * n_ops);
    const mlrn_inter_op op = ops[k];
    const double* r = table + row * ld;
    const double a = r[op.a], b = op.b >= 0 ? r[op.b] : 0.0, c = op.c >= 0 ? r[op.c] : 0.0;
    double v;
    switch (op.kind) {
      case MLRN_IOP_MUL: v = a * b; break;                                   // a * b
      case MLRN_IOP_DIV_OFFSET: v = a / (b + op.k); break;                   // a / (b + k)
      case MLRN_IOP_DIFF_SCALED: v = (a - b) / op.k; break;                  // (a - b) / k
      case MLRN_IOP_CLIP_POW: v = pow(fmin(fmax(a, 0.0), op.k), 0.25); break;  // clip(a, 0, k) ** 0.25 (NaN propagates like numpy.clip)
      case MLRN_IOP_SUB: v = a - b; break;                                   // a - b
      case MLRN_IOP_SQUARE: v = a * a; break;                                // a ** 2
      case MLRN_IOP_ASYM_COLOR: v = (a / (b + 1.0)) * c; break;              // (rise / (fade + 1)) * colour
      default: v = a; break;
    }
    if (op.kind == MLRN_IOP_CLIP_POW && a != a) v = a;
    out[row * ld_out + k] = v;
  }
}

}  // namespace
}  // namespace mlrn

extern "C" int mlrn_interactions(int device, const double* table, int64_t n_rows, int32_t ld, const mlrn_inter_op* ops,
                                 int32_t n_ops, double* out, int32_t ld_out, void* stream) {
  using namespace mlrn;
  if (n_rows < 0 || n_ops < 0 || ld <= 0 || ld_out < n_ops || (n_rows > 0 && n_ops > 0 && (!table || !ops || !out))) {
    set_error("mlrn_interactions: bad sizes or null pointer"); return MLRN_EINVAL;
  }
  if (n_rows == 0 || n_ops == 0) return MLRN_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device visible; libmallorn_b200 has no CPU fallback"); return MLRN_ENODEVICE; }
  for (int k = 0; k < n_ops; ++k)
    if (ops[k].a < 0 || ops[k].a >= ld || ops[k].b >= ld || ops[k].c >= ld) { set_error("mlrn_interactions: column index out of range"); return MLRN_EINVAL; }
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  cudaStream_t st = (cudaStream_t)stream;
  mlrn_inter_op* dops = nullptr;
  cudaError_t e = cudaMallocAsync(&dops, sizeof(mlrn_inter_op) * (size_t)n_ops, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dops, ops, sizeof(mlrn_inter_op) * (size_t)n_ops, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const long long total = (long long)n_rows * n_ops;
    long long want = (total + 255) / 256;
    const int grid = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);
    k_interactions<<<grid, 256, 0, st>>>(table, n_rows, ld, dops, n_ops, out, ld_out);
    e = cudaGetLastError();
    count_launch();
  }
  if (dops) cudaFreeAsync(dops, st);
  if (prev != device) cudaSetDevice(prev);
  if (e != cudaSuccess) return cuda_fail(e, "mlrn_interactions");
  return MLRN_OK;
}
