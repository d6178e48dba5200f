// Depthwise 2-D convolution (groups == channels, multiplier 1) for sm_100a: forward, input gradient and filter gradient.
//
// Reference: DepthwiseConv2d{Forward,BackwardData,BackwardFilter}Kernel, src/operator/nn/depthwise_convolution_tf.cuh:76-754 (Kepler-era
// one-output-per-thread kernels with shuffle reductions, __launch_bounds__(1024, 2)).  Design here: one CTA per (image, channel) plane —
// the plane (with its halo) and the K x K filter are staged ONCE in shared memory, every thread then produces several outputs from shared
// memory only; the filter gradient is a per-plane K x K reduction (warp shuffles) accumulated over the batch with one atomic per tap and plane.
// A depthwise layer does K*K MACs per element, so it is bandwidth-bound: each input element is read from HBM exactly once per pass.
#include "common.cuh"

namespace gx {

struct DwParams {
  int N, C, H, W, KH, KW, SH, SW, PH, PW, OH, OW;
};

// y[n,c,oh,ow] = bias[c] + sum_{kh,kw} x[n,c,oh*SH-PH+kh,ow*SW-PW+kw] * w[c,kh,kw]
__global__ void __launch_bounds__(256) depthwise_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ y, const DwParams p, int relu) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  float* sx = sm;                         // [H][W]
  float* sw = sx + p.H * p.W;             // [KH][KW]
  const int n = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const float* xp = x + ((long long)n * p.C + c) * p.H * p.W;
  for (int i = threadIdx.x; i < p.H * p.W; i += blockDim.x) sx[i] = xp[i];
  for (int i = threadIdx.x; i < p.KH * p.KW; i += blockDim.x) sw[i] = w[(long long)c * p.KH * p.KW + i];
  __syncthreads();
  const float b = bias != nullptr ? bias[c] : 0.f;
  float* yp = y + ((long long)n * p.C + c) * p.OH * p.OW;
  for (int o = threadIdx.x; o < p.OH * p.OW; o += blockDim.x) {
    const int oh = o / p.OW, ow = o - oh * p.OW;
    const int h0 = oh * p.SH - p.PH, w0 = ow * p.SW - p.PW;
    float acc = b;
    for (int kh = 0; kh < p.KH; ++kh) {
      const int ih = h0 + kh;
      if (ih < 0 || ih >= p.H) continue;
      for (int kw = 0; kw < p.KW; ++kw) {
        const int iw = w0 + kw;
        if (iw >= 0 && iw < p.W) acc = fmaf(sx[ih * p.W + iw], sw[kh * p.KW + kw], acc);
      }
    }
    yp[o] = relu ? fmaxf(acc, 0.f) : acc;
  }
}

// dx[n,c,ih,iw] = sum_{kh,kw : (ih+PH-kh) % SH == 0 ...} dy[n,c,(ih+PH-kh)/SH,(iw+PW-kw)/SW] * w[c,kh,kw]
__global__ void __launch_bounds__(256) depthwise_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                               const DwParams p) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  float* sdy = sm;                        // [OH][OW]
  float* sw = sdy + p.OH * p.OW;
  const int n = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const float* dyp = dy + ((long long)n * p.C + c) * p.OH * p.OW;
  for (int i = threadIdx.x; i < p.OH * p.OW; i += blockDim.x) sdy[i] = dyp[i];
  for (int i = threadIdx.x; i < p.KH * p.KW; i += blockDim.x) sw[i] = w[(long long)c * p.KH * p.KW + i];
  __syncthreads();
  float* dxp = dx + ((long long)n * p.C + c) * p.H * p.W;
  for (int e = threadIdx.x; e < p.H * p.W; e += blockDim.x) {
    const int ih = e / p.W, iw = e - ih * p.W;
    float acc = 0.f;
    for (int kh = 0; kh < p.KH; ++kh) {
      const int th = ih + p.PH - kh;
      if (th < 0 || th % p.SH) continue;
      const int oh = th / p.SH;
      if (oh >= p.OH) continue;
      for (int kw = 0; kw < p.KW; ++kw) {
        const int tw = iw + p.PW - kw;
        if (tw < 0 || tw % p.SW) continue;
        const int ow = tw / p.SW;
        if (ow < p.OW) acc = fmaf(sdy[oh * p.OW + ow], sw[kh * p.KW + kw], acc);
      }
    }
    dxp[e] = acc;
  }
}

// dw[c,kh,kw] += sum_{n,oh,ow} dy[n,c,oh,ow] * x[n,c,oh*SH-PH+kh,ow*SW-PW+kw];  db[c] += sum dy      (atomics over the batch: zero dw/db first)
__global__ void __launch_bounds__(256) depthwise_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                               float* __restrict__ db, const DwParams p) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];
  float* sx = sm;                         // [H][W]
  float* sdy = sx + p.H * p.W;            // [OH][OW]
  __shared__ float red[8];
  const int n = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const float* xp = x + ((long long)n * p.C + c) * p.H * p.W;
  const float* dyp = dy + ((long long)n * p.C + c) * p.OH * p.OW;
  for (int i = threadIdx.x; i < p.H * p.W; i += blockDim.x) sx[i] = xp[i];
  for (int i = threadIdx.x; i < p.OH * p.OW; i += blockDim.x) sdy[i] = dyp[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int tap = 0; tap <= p.KH * p.KW; ++tap) {      // tap == KH*KW: the bias gradient
    const int kh = tap / p.KW, kw = tap - kh * p.KW;
    float acc = 0.f;
    for (int o = threadIdx.x; o < p.OH * p.OW; o += blockDim.x) {
      if (tap == p.KH * p.KW) { acc += sdy[o]; continue; }
      const int oh = o / p.OW, ow = o - oh * p.OW;
      const int ih = oh * p.SH - p.PH + kh, iw = ow * p.SW - p.PW + kw;
      if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) acc = fmaf(sdy[o], sx[ih * p.W + iw], acc);
    }
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < nwarp; ++i) s += red[i];
      if (tap < p.KH * p.KW) atomicAdd(dw + (long long)c * p.KH * p.KW + tap, s);
      else if (db != nullptr) atomicAdd(db + c, s);
    }
  }
}

}  // namespace gx

using namespace gx;

static DwParams mk(int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW) {
  DwParams p; p.N = N; p.C = C; p.H = H; p.W = W; p.KH = KH; p.KW = KW; p.SH = SH; p.SW = SW; p.PH = PH; p.PW = PW;
  p.OH = (H + 2 * PH - KH) / SH + 1; p.OW = (W + 2 * PW - KW) / SW + 1;
  return p;
}
template <typename K>
static int ensure_smem(K kern, size_t bytes) {
  if (bytes > 200 * 1024) return -1;        // plane does not fit: caller falls back
  if (bytes > 48 * 1024) { if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1; }
  return 0;
}
GX_API int gx_depthwise_fwd(const float* x, const float* w, const float* bias, float* y, int N, int C, int H, int W, int KH, int KW, int SH, int SW,
                            int PH, int PW, int relu, cudaStream_t s) {
  const DwParams p = mk(N, C, H, W, KH, KW, SH, SW, PH, PW);
  const size_t smem = ((size_t)H * W + (size_t)KH * KW) * 4;
  if (p.OH < 1 || p.OW < 1 || ensure_smem(depthwise_fwd_kernel, smem)) return -1;
  launch_pdl(depthwise_fwd_kernel, dim3(N * C), dim3(256), smem, s, x, w, bias, y, p, relu);
  return GX_CHECK_LAUNCH();
}
GX_API int gx_depthwise_dgrad(const float* dy, const float* w, float* dx, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW,
                              cudaStream_t s) {
  const DwParams p = mk(N, C, H, W, KH, KW, SH, SW, PH, PW);
  const size_t smem = ((size_t)p.OH * p.OW + (size_t)KH * KW) * 4;
  if (p.OH < 1 || p.OW < 1 || ensure_smem(depthwise_dgrad_kernel, smem)) return -1;
  launch_pdl(depthwise_dgrad_kernel, dim3(N * C), dim3(256), smem, s, dy, w, dx, p);
  return GX_CHECK_LAUNCH();
}
GX_API int gx_depthwise_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH,
                              int PW, cudaStream_t s) {
  const DwParams p = mk(N, C, H, W, KH, KW, SH, SW, PH, PW);
  const size_t smem = ((size_t)H * W + (size_t)p.OH * p.OW) * 4;
  if (p.OH < 1 || p.OW < 1 || ensure_smem(depthwise_wgrad_kernel, smem)) return -1;
  launch_pdl(depthwise_wgrad_kernel, dim3(N * C), dim3(256), smem, s, x, dy, dw, db, p);
  return GX_CHECK_LAUNCH();
}
