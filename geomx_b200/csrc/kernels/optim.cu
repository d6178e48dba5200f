// Fused optimizer / elementwise kernels for sm_100a.
//
// Reference: SGD / SGD-momentum / Adam / DCASGD are 1-3 mshadow MapPlanKernel passes *per parameter*
// (src/operator/optimizer_op-inl.h:86-103,305-326,840-873; DCASGD python/mxnet/optimizer/optimizer.py:872-925), driven by a
// Python loop.  Here:
//   * arena_* : ONE launch updates a whole flat parameter arena (all keys), 128-bit accesses, with the step counter and
//     Adam bias correction kept ON THE DEVICE so the launch is CUDA-graph replayable (no host scalar changes per step).
//     Per-key lr/wd multipliers come from a per-tile table (tile = 1024 floats; keys are tile-aligned in the arena).
//   * multi_tensor_* : one launch over a pointer table (Trainer path for tensors that do not live in an arena).
//   * n-ary sum (CommDevice reduce), scale+cast (the script-level `grad / num_samples` and `.astype('float16')`).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace gx {

enum OptKind { OPT_SGD = 0, OPT_ADAM = 1, OPT_DCASGD = 2 };

struct OptHyper {
  float lr, wd, rescale, clip;      // clip < 0 -> off
  float momentum;                   // sgd / dcasgd
  float beta1, beta2, eps;          // adam
  float lamda;                      // dcasgd
};

constexpr int ARENA_TILE = 1024;

__device__ __forceinline__ float prep_grad(float g, float w, const OptHyper& h, float wd) {
  g *= h.rescale;
  if (h.clip >= 0.f) g = fminf(fmaxf(g, -h.clip), h.clip);
  return fmaf(wd, w, g);
}

// one element of each optimizer (s0/s1 = state slots: sgd: s0=mom; adam: s0=m, s1=v; dcasgd: s0=mom, s1=prev_weight)
template <int KIND>
__device__ __forceinline__ void opt_elem(float& w, float g, float& s0, float& s1, const OptHyper& h, float lr, float wd) {
  if (KIND == OPT_SGD) {
    g = prep_grad(g, w, h, wd);
    if (h.momentum != 0.f) { s0 = h.momentum * s0 - lr * g; w += s0; }
    else w -= lr * g;
  } else if (KIND == OPT_ADAM) {
    // adam_update clips the REGULARISED gradient: rescale, + wd*w, then clip (src/operator/optimizer_op-inl.h:840-873)
    g = fmaf(wd, w, g * h.rescale);
    if (h.clip >= 0.f) g = fminf(fmaxf(g, -h.clip), h.clip);
    s0 = h.beta1 * s0 + (1.f - h.beta1) * g;
    s1 = h.beta2 * s1 + (1.f - h.beta2) * g * g;
    w -= lr * s0 / (sqrtf(s1) + h.eps);
  } else {
    g *= h.rescale;
    if (h.clip >= 0.f) g = fminf(fmaxf(g, -h.clip), h.clip);
    const float upd = g + wd * w + h.lamda * g * g * (w - s1);
    const float prev = w;
    if (h.momentum != 0.f) { s0 = h.momentum * s0 - lr * upd; w += s0; }
    else w -= lr * upd;
    s1 = prev;
  }
}

// step_state[0] = number of completed updates t, step_state[1] = CTA completion counter (self-resetting)
template <int KIND>
__global__ void __launch_bounds__(256) arena_opt_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ s0, float* __restrict__ s1,
                                                         long long n, const float2* __restrict__ tile_mult, OptHyper h, int* __restrict__ step_state,
                                                         float* __restrict__ g_zero) {
  gx::pdl_wait();
  gx::pdl_launch();
  const int t = step_state ? (*reinterpret_cast<volatile int*>(step_state)) + 1 : 1;
  float lr_t = h.lr;
  if (KIND == OPT_ADAM) lr_t = h.lr * sqrtf(1.f - powf(h.beta2, (float)t)) / (1.f - powf(h.beta1, (float)t));
  const long long n4 = n >> 2;  // arenas are tile-padded, n % 1024 == 0
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float lr = lr_t, wd = h.wd;
    if (tile_mult) { const float2 m = __ldg(tile_mult + (i * 4) / ARENA_TILE); lr *= m.x; wd *= m.y; }
    float4 W = reinterpret_cast<float4*>(w)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 A = s0 ? reinterpret_cast<float4*>(s0)[i] : make_float4(0, 0, 0, 0);
    float4 Bv = s1 ? reinterpret_cast<float4*>(s1)[i] : make_float4(0, 0, 0, 0);
    opt_elem<KIND>(W.x, G.x, A.x, Bv.x, h, lr, wd);
    opt_elem<KIND>(W.y, G.y, A.y, Bv.y, h, lr, wd);
    opt_elem<KIND>(W.z, G.z, A.z, Bv.z, h, lr, wd);
    opt_elem<KIND>(W.w, G.w, A.w, Bv.w, h, lr, wd);
    reinterpret_cast<float4*>(w)[i] = W;
    if (s0) reinterpret_cast<float4*>(s0)[i] = A;
    if (s1) reinterpret_cast<float4*>(s1)[i] = Bv;
    if (g_zero) reinterpret_cast<float4*>(g_zero)[i] = make_float4(0, 0, 0, 0);  // fused zero_grad for the next step
  }
  if (step_state) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int done = atomicAdd(step_state + 1, 1);
      if (done == (int)gridDim.x - 1) {  // every other CTA has already read t
        step_state[1] = 0;
        step_state[0] = t;
        __threadfence();
      }
    }
  }
}

struct TensorEntry { float* w; const float* g; float* s0; float* s1; long long n; float lr_mult, wd_mult; };

template <int KIND>
__global__ void __launch_bounds__(256) multi_tensor_opt_kernel(const TensorEntry* __restrict__ table, OptHyper h, float lr_t) {
  gx::pdl_wait();
  gx::pdl_launch();
  const TensorEntry e = table[blockIdx.y];
  const float lr = lr_t * e.lr_mult, wd = h.wd * e.wd_mult;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e.n; i += (long long)gridDim.x * blockDim.x) {
    float W = e.w[i], A = e.s0 ? e.s0[i] : 0.f, Bv = e.s1 ? e.s1[i] : 0.f;
    opt_elem<KIND>(W, e.g[i], A, Bv, h, lr, wd);
    e.w[i] = W;
    if (e.s0) e.s0[i] = A;
    if (e.s1) e.s1[i] = Bv;
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) single_opt_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ s0, float* __restrict__ s1,
                                                          long long n, OptHyper h, float lr_t) {
  gx::pdl_wait();
  gx::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float W = w[i], A = s0 ? s0[i] : 0.f, Bv = s1 ? s1[i] : 0.f;
    opt_elem<KIND>(W, g[i], A, Bv, h, lr_t, h.wd);
    w[i] = W;
    if (s0) s0[i] = A;
    if (s1) s1[i] = Bv;
  }
}

struct PtrList8 { const float* p[8]; };
__global__ void __launch_bounds__(256) nary_sum_kernel(float* __restrict__ out, PtrList8 in, int cnt, long long n) {
  gx::pdl_wait();
  gx::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < cnt) acc += in.p[j][i];
    out[i] = acc;
  }
}

template <typename T> __device__ __forceinline__ T cvt(float v);
template <> __device__ __forceinline__ float cvt<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T> __device__ __forceinline__ float up(T v);
template <> __device__ __forceinline__ float up<float>(float v) { return v; }
template <> __device__ __forceinline__ float up<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float up<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) scale_cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, float scale, long long n) {
  gx::pdl_wait();
  gx::pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = cvt<TO>(up<TI>(x[i]) * scale);
}

static inline int grid_for(long long n, int per_thread = 1) {
  long long b = (n / per_thread + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

}  // namespace gx

using namespace gx;

static OptHyper mk(float lr, float wd, float rescale, float clip, float momentum, float b1, float b2, float eps, float lamda) {
  OptHyper h; h.lr = lr; h.wd = wd; h.rescale = rescale; h.clip = clip; h.momentum = momentum; h.beta1 = b1; h.beta2 = b2; h.eps = eps; h.lamda = lamda;
  return h;
}

// kind: 0 sgd, 1 adam, 2 dcasgd.  n must be a multiple of 4 (arenas are 1024-padded).  step_state: int[2] device or null.
GX_API int gx_arena_opt(int kind, float* w, const float* g, float* s0, float* s1, long long n, const void* tile_mult, float lr, float wd,
                        float rescale, float clip, float momentum, float b1, float b2, float eps, float lamda, int* step_state, float* g_zero,
                        cudaStream_t s) {
  const OptHyper h = mk(lr, wd, rescale, clip, momentum, b1, b2, eps, lamda);
  const int grid = grid_for(n, 4);
  const float2* tm = reinterpret_cast<const float2*>(tile_mult);
  if (kind == OPT_SGD) launch_pdl(arena_opt_kernel<OPT_SGD>, dim3(grid), dim3(256), 0, s, w, g, s0, s1, n, tm, h, step_state, g_zero);
  else if (kind == OPT_ADAM) launch_pdl(arena_opt_kernel<OPT_ADAM>, dim3(grid), dim3(256), 0, s, w, g, s0, s1, n, tm, h, step_state, g_zero);
  else launch_pdl(arena_opt_kernel<OPT_DCASGD>, dim3(grid), dim3(256), 0, s, w, g, s0, s1, n, tm, h, step_state, g_zero);
  return GX_CHECK_LAUNCH();
}

GX_API int gx_multi_tensor_opt(int kind, const void* table, int num_tensors, long long max_n, float lr_t, float wd, float rescale, float clip,
                               float momentum, float b1, float b2, float eps, float lamda, cudaStream_t s) {
  const OptHyper h = mk(lr_t, wd, rescale, clip, momentum, b1, b2, eps, lamda);
  int gx_ = grid_for(max_n); if (gx_ > 64) gx_ = 64;
  dim3 grid(gx_, num_tensors);
  const TensorEntry* t = reinterpret_cast<const TensorEntry*>(table);
  if (kind == OPT_SGD) launch_pdl(multi_tensor_opt_kernel<OPT_SGD>, dim3(grid), dim3(256), 0, s, t, h, lr_t);
  else if (kind == OPT_ADAM) launch_pdl(multi_tensor_opt_kernel<OPT_ADAM>, dim3(grid), dim3(256), 0, s, t, h, lr_t);
  else launch_pdl(multi_tensor_opt_kernel<OPT_DCASGD>, dim3(grid), dim3(256), 0, s, t, h, lr_t);
  return GX_CHECK_LAUNCH();
}

GX_API int gx_single_opt(int kind, float* w, const float* g, float* s0, float* s1, long long n, float lr_t, float wd, float rescale, float clip,
                         float momentum, float b1, float b2, float eps, float lamda, cudaStream_t s) {
  const OptHyper h = mk(lr_t, wd, rescale, clip, momentum, b1, b2, eps, lamda);
  const int grid = grid_for(n);
  if (kind == OPT_SGD) launch_pdl(single_opt_kernel<OPT_SGD>, dim3(grid), dim3(256), 0, s, w, g, s0, s1, n, h, lr_t);
  else if (kind == OPT_ADAM) launch_pdl(single_opt_kernel<OPT_ADAM>, dim3(grid), dim3(256), 0, s, w, g, s0, s1, n, h, lr_t);
  else launch_pdl(single_opt_kernel<OPT_DCASGD>, dim3(grid), dim3(256), 0, s, w, g, s0, s1, n, h, lr_t);
  return GX_CHECK_LAUNCH();
}

GX_API int gx_nary_sum(float* out, const float* const* inputs, int cnt, long long n, cudaStream_t s) {
  if (cnt < 1 || cnt > 8) return -1;
  PtrList8 l;
  for (int i = 0; i < 8; ++i) l.p[i] = i < cnt ? inputs[i] : nullptr;
  launch_pdl(nary_sum_kernel, dim3(grid_for(n)), dim3(256), 0, s, out, l, cnt, n);
  return GX_CHECK_LAUNCH();
}

// dtype codes: 0 fp32, 1 fp16, 2 bf16
GX_API int gx_scale_cast(const void* x, int in_dt, void* y, int out_dt, float scale, long long n, cudaStream_t s) {
  const int grid = grid_for(n);
#define GX_SC(TI, TO) launch_pdl(scale_cast_kernel<TI, TO>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const TI*>(x), reinterpret_cast<TO*>(y), scale, n)
  if (in_dt == 0 && out_dt == 0) GX_SC(float, float);
  else if (in_dt == 0 && out_dt == 1) GX_SC(float, __half);
  else if (in_dt == 0 && out_dt == 2) GX_SC(float, __nv_bfloat16);
  else if (in_dt == 1 && out_dt == 0) GX_SC(__half, float);
  else if (in_dt == 2 && out_dt == 0) GX_SC(__nv_bfloat16, float);
  else if (in_dt == 1 && out_dt == 1) GX_SC(__half, __half);
  else if (in_dt == 2 && out_dt == 2) GX_SC(__nv_bfloat16, __nv_bfloat16);
  else return -1;
#undef GX_SC
  return GX_CHECK_LAUNCH();
}
