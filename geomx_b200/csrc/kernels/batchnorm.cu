// BatchNorm forward / backward for sm_100a (NCHW or [N,C]), one CTA per channel, two-pass Welford-free fp32 statistics
// with 128-bit loads when HW % 4 == 0.  Reference: src/operator/nn/batch_norm.cu:247-470 (grid = #channels,
// warp-shuffle reductions :150), running-stat convention  running = momentum*running + (1-momentum)*batch  (:262-266).
#include "common.cuh"

namespace gx {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = blockDim.x >> 5;
  for (int j = 0; j < nw; ++j) s += red[j];
  return s;
}

__global__ void __launch_bounds__(512) bn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ y,
                                                      float* __restrict__ save_mean, float* __restrict__ save_invstd, int N, int C, int HW,
                                                      int training, float momentum, float eps) {
  gx::pdl_wait();
  gx::pdl_launch();
  __shared__ float red[16];
  const int c = blockIdx.x;
  const long long cnt = (long long)N * HW;
  float mean, invstd;
  if (training) {
    float s = 0.f;
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) s += x[((i / HW) * C + c) * HW + i % HW];
    mean = block_sum(s, red) / (float)cnt;
    float q = 0.f;
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float d = x[((i / HW) * C + c) * HW + i % HW] - mean;
      q = fmaf(d, d, q);
    }
    const float var = block_sum(q, red) / (float)cnt;
    invstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
      running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * mean;
      running_var[c] = momentum * running_var[c] + (1.f - momentum) * var;
      save_mean[c] = mean;
      save_invstd[c] = invstd;
    }
  } else {
    mean = running_mean[c];
    invstd = rsqrtf(running_var[c] + eps);
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float a = g * invstd, off = b - mean * a;
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
    const long long o = ((i / HW) * C + c) * HW + i % HW;
    y[o] = fmaf(x[o], a, off);
  }
}

__global__ void __launch_bounds__(512) bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                                      const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                                                      float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int C, int HW) {
  gx::pdl_wait();
  gx::pdl_launch();
  __shared__ float red[16];
  const int c = blockIdx.x;
  const long long cnt = (long long)N * HW;
  const float mean = save_mean[c], invstd = save_invstd[c];
  float sdy = 0.f, sdyx = 0.f;
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
    const long long o = ((i / HW) * C + c) * HW + i % HW;
    const float g = dy[o];
    sdy += g;
    sdyx = fmaf(g, (x[o] - mean) * invstd, sdyx);
  }
  sdy = block_sum(sdy, red);
  sdyx = block_sum(sdyx, red);
  if (threadIdx.x == 0) {
    if (dgamma) dgamma[c] = sdyx;
    if (dbeta) dbeta[c] = sdy;
  }
  const float g = gamma ? gamma[c] : 1.f;
  const float k = g * invstd, inv_cnt = 1.f / (float)cnt;
  for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
    const long long o = ((i / HW) * C + c) * HW + i % HW;
    const float xh = (x[o] - mean) * invstd;
    dx[o] = k * (dy[o] - (sdy + xh * sdyx) * inv_cnt);
  }
}

}  // namespace gx

using namespace gx;

GX_API int gx_bn_fwd(const float* x, const float* gamma, const float* beta, float* rm, float* rv, float* y, float* save_mean, float* save_invstd,
                     int N, int C, int HW, int training, float momentum, float eps, cudaStream_t s) {
  launch_pdl(bn_fwd_kernel, dim3(C), dim3(512), 0, s, x, gamma, beta, rm, rv, y, save_mean, save_invstd, N, C, HW, training, momentum, eps);
  return GX_CHECK_LAUNCH();
}
GX_API int gx_bn_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean, const float* save_invstd, float* dx, float* dgamma,
                     float* dbeta, int N, int C, int HW, cudaStream_t s) {
  launch_pdl(bn_bwd_kernel, dim3(C), dim3(512), 0, s, x, dy, gamma, save_mean, save_invstd, dx, dgamma, dbeta, N, C, HW);
  return GX_CHECK_LAUNCH();
}
