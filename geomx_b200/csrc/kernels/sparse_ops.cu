// Row-sparse support kernels: row-id dedup (sort + unique), row gather / scatter / scatter-add.
//
// Parity: src/kvstore/kvstore_utils.cu:42-108 UniqueImplGPU (CUB radix sort + DeviceSelect::Unique, count read back by the caller),
// sparse_retain / row_sparse pull helpers (src/operator/tensor/sparse_retain-inl.h, kvstore_local.h:357-417).  One CTA row per 128-bit
// lane group: rows of the embedding-style tables these paths serve are >= 16 bytes, so every access is a float4 when aligned.
#include <cub/cub.cuh>

#include "common.cuh"

namespace gx {

__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ ids, float* __restrict__ out, int n, int L) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x;
  if (row >= n) return;
  const float* s = src + ids[row] * (long long)L;
  float* d = out + (long long)row * L;
  if ((L & 3) == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
    for (int i = threadIdx.x; i < L / 4; i += blockDim.x) reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
  } else {
    for (int i = threadIdx.x; i < L; i += blockDim.x) d[i] = s[i];
  }
}

// dst[ids[r]] (+)= rows[r]; `add` uses atomics so duplicate ids accumulate (row_sparse gradient push)
__global__ void scatter_rows_kernel(float* __restrict__ dst, const long long* __restrict__ ids, const float* __restrict__ rows, int n, int L, int add) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x;
  if (row >= n) return;
  float* d = dst + ids[row] * (long long)L;
  const float* s = rows + (long long)row * L;
  if (add) { for (int i = threadIdx.x; i < L; i += blockDim.x) atomicAdd(d + i, s[i]); }
  else { for (int i = threadIdx.x; i < L; i += blockDim.x) d[i] = s[i]; }
}

}  // namespace gx

using namespace gx;

// workspace bytes for gx_unique_i64 on n ids
GX_API long long gx_unique_i64_workspace(int n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const long long*)nullptr, (long long*)nullptr, n);
  cub::DeviceSelect::Unique(nullptr, b, (const long long*)nullptr, (long long*)nullptr, (int*)nullptr, n);
  return (long long)(a > b ? a : b);
}
// sorted_tmp, out: n ids each; count_dev: one int (number of unique ids, read back by the caller as the reference does)
GX_API int gx_unique_i64(const long long* in, int n, long long* sorted_tmp, long long* out, int* count_dev, void* ws, long long ws_bytes, cudaStream_t s) {
  size_t bytes = (size_t)ws_bytes;
  if (cub::DeviceRadixSort::SortKeys(ws, bytes, in, sorted_tmp, n, 0, 64, s) != cudaSuccess) return -2;
  bytes = (size_t)ws_bytes;
  if (cub::DeviceSelect::Unique(ws, bytes, sorted_tmp, out, count_dev, n, s) != cudaSuccess) return -3;
  return GX_CHECK_LAUNCH();
}
GX_API int gx_gather_rows(const float* src, const long long* ids, float* out, int n, int L, cudaStream_t s) {
  if (n <= 0) return 0;
  launch_pdl(gather_rows_kernel, dim3(n), dim3(L >= 1024 ? 256 : 64), 0, s, src, ids, out, n, L);
  return GX_CHECK_LAUNCH();
}
GX_API int gx_scatter_rows(float* dst, const long long* ids, const float* rows, int n, int L, int add, cudaStream_t s) {
  if (n <= 0) return 0;
  launch_pdl(scatter_rows_kernel, dim3(n), dim3(L >= 1024 ? 256 : 64), 0, s, dst, ids, rows, n, L, add);
  return GX_CHECK_LAUNCH();
}
