// Direct (fp32 FMA) convolution kernels for the small-batch regime of the reference's demo CNN
//   Conv2D(16,k5,relu) -> MaxPool2 -> Conv2D(32,k5,relu) -> MaxPool2      (examples/cnn.py:56-60, input 1x28x28)
// forward and backward, TWO launches in total:
//   cnn_fwd_kernel      x -> a1 (+argmax) -> a2 (+argmax)          both convolutions, bias, ReLU and both max-pools; a1 stays in shared memory
//   cnn_bwd_all_kernel  one heterogeneous grid: CTAs [0, 4B)  da2 -> da1 (shared memory only) -> dW0, db0   (conv1 data gradient + pool/ReLU
//                       backward of both layers + conv0 weight gradient), CTAs [4B, 4B+128)  da2, a1 -> dW1, db1  (conv1 weight gradient,
//                       sparse: one non-zero per pooling window).  Both bodies resolve the programmatic grid dependency only AFTER the staging
//                       that does not depend on the preceding kernel, so the CTAs become resident and stage while that kernel still runs.
//                       (cnn_bwd_kernel / cnn_wgrad1_kernel: the same bodies as separate launches; cnn_bwd_exchange_kernel: the same grid
//                       whose last CTAs also perform the conv keys' exchange — measured, no gain, opt-in.)
//
// Why not the tcgen05 implicit GEMM here: with a per-worker batch of 32 the conv1 products are M = 2048 x N = 32 x K = 400 — 16 tensor-core
// tiles.  Measured on B200 (tools/kernel_times.py, profiles/): the tcgen05 path needs 8 us (TF32) / 16 us (3xTF32) for the forward GEMM alone
// and 30-50 us for forward + dgrad + wgrad + im2col/col2im traffic, all of it launch / TMA / TMEM latency on 16 SMs, while the same 78 MFLOP
// spread over 128 CTAs of plain FMAs finish in a few microseconds, in exact fp32 (the reference's cublasSgemmEx precision).  The max-pool makes
// the backward pass 4x sparse (one non-zero per 2x2 window), which the weight-gradient kernel exploits.  Larger batches / other geometries use
// the tcgen05 GEMM path (ops/functional.py conv2d).
//
// Reference kernels replaced: im2col_gpu_kernel / col2im_gpu_kernel (src/operator/nn/im2col.cuh:79-187), per-image cublasSgemmEx
// (convolution-inl.h:165-284), bias broadcast, ReLU (activation-inl.h:89-121), pool_max_2d / unpool (pool.cuh:128-165, 379-432).
#include "common.cuh"
#include "hips_ll.cuh"

namespace gx {

// geometry (compile time)
constexpr int CD_H = 28, CD_K = 5;
constexpr int CD_C1 = 16, CD_P1 = 12;            // conv0: 16 channels, 24x24 -> pooled 12x12
constexpr int CD_C2 = 32, CD_P2 = 4;             // conv1: 32 channels,  8x8  -> pooled 4x4
constexpr int CD_W1PAD = 28;                     // 25 taps padded to 28 floats (16-byte rows)

// asynchronous global -> shared copies: all of a CTA's operand loads are in flight at once instead of one dependent round trip per element
__device__ __forceinline__ void cpa4(float* sdst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa16(float* sdst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// 2x2 output window of a 5x5 valid convolution from a 6x6 input patch: acc[dy*2+dx] += sum_{r,s} patch[dy+r][dx+s] * w[r*5+s]
__device__ __forceinline__ void window_fma(const float (&pt)[6][6], const float (&w)[28], float (&acc)[4]) {
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const float wv = w[r * 5 + s];
      acc[0] = fmaf(wv, pt[r][s], acc[0]);
      acc[1] = fmaf(wv, pt[r][s + 1], acc[1]);
      acc[2] = fmaf(wv, pt[r + 1][s], acc[2]);
      acc[3] = fmaf(wv, pt[r + 1][s + 1], acc[3]);
    }
}
// 6x6 patch whose top-left corner is `p` (8-byte aligned, row stride `ld` floats, ld even)
__device__ __forceinline__ void load_patch(const float* p, int ld, float (&pt)[6][6]) {
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float2 v = *reinterpret_cast<const float2*>(p + r * ld + 2 * q);
      pt[r][2 * q] = v.x; pt[r][2 * q + 1] = v.y;
    }
}
__device__ __forceinline__ void load_w28(const float* p, float (&w)[28]) {
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
  }
}
// max-pool of a window's four pre-bias values (first maximum wins, like maxpool2x2_fwd_kernel), then bias + ReLU
__device__ __forceinline__ float pool4(const float (&a)[4], float bias, int& bi) {
  float best = a[0]; bi = 0;
  if (a[1] > best) { best = a[1]; bi = 1; }
  if (a[2] > best) { best = a[2]; bi = 2; }
  if (a[3] > best) { best = a[3]; bi = 3; }
  return fmaxf(best + bias, 0.f);
}

// ------------------------------------------------------------------------------------------------------------------ forward
// grid (B, 4), 256 threads.  CTA (b, g): conv0 of image b (all 16 channels, kept in shared memory) and conv1 for output channels [8g, 8g+8).
__global__ void __launch_bounds__(256) cnn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w0, const float* __restrict__ b0,
                                                       const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ a1,
                                                       unsigned char* __restrict__ idx1, float* __restrict__ a2, unsigned char* __restrict__ idx2,
                                                       const float* __restrict__ carry_src, float* __restrict__ carry_dst, int carry_n, float* __restrict__ x_keep,
                                                       unsigned long long* dbg) {
  __shared__ __align__(16) float sx[CD_H * CD_H];
  __shared__ __align__(16) float sw0[CD_C1 * CD_W1PAD];
  __shared__ float sb0[CD_C1];
  __shared__ __align__(16) float sa1[CD_C1 * CD_P1 * CD_P1];
  __shared__ __align__(16) float sw1[8 * CD_C1 * CD_W1PAD];
  __shared__ float sb1[8];
  __shared__ __align__(16) float spart[4 * 64 * 8];
  const bool dbg_on = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  auto stamp = [&](int slot) { if (dbg_on) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[slot] = t; } };
  stamp(0);
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  // look-ahead steps: this launch is the tail of the previous batch's step; it hands the labels that arrived with its image batch over to
  // the buffer the classifier head of the NEXT launch reads (a 128-byte side job instead of a copy node in the step graph)
  if (carry_n > 0 && b == 0 && g == 0) for (int i = tid; i < carry_n; i += 256) carry_dst[i] = carry_src[i];
  for (int i = tid; i < CD_H * CD_H / 4; i += 256) cpa16(sx + 4 * i, x + (long long)b * CD_H * CD_H + 4 * i);
  for (int i = tid; i < CD_C1 * 25; i += 256) { const int ch = i / 25, t = i - ch * 25; cpa4(sw0 + ch * CD_W1PAD + t, w0 + i); }
  for (int i = tid; i < CD_C1 * 3; i += 256) sw0[(i / 3) * CD_W1PAD + 25 + i % 3] = 0.f;          // pad taps 25..27
  if (tid < CD_C1) sb0[tid] = b0[tid];
  {
    const float* wsrc = w1 + (long long)(8 * g) * CD_C1 * 25;      // the 8 output channels of this CTA: 3200 contiguous floats
    for (int i = tid; i < 8 * CD_C1 * 25; i += 256) { const int r = i / 25, t = i - r * 25; cpa4(sw1 + r * CD_W1PAD + t, wsrc + i); }
    for (int i = tid; i < 8 * CD_C1 * 3; i += 256) sw1[(i / 3) * CD_W1PAD + 25 + i % 3] = 0.f;
  }
  if (tid < 8) sb1[tid] = b1[8 * g + tid];
  cpa_wait_all();
  __syncthreads();
  stamp(1);
  // look-ahead steps: the backward pass of this batch runs in the NEXT launch, after the following batch has overwritten x — keep a copy
  if (x_keep != nullptr && g == 0) for (int i = tid; i < CD_H * CD_H / 4; i += 256)
    reinterpret_cast<float4*>(x_keep + (long long)b * CD_H * CD_H)[i] = reinterpret_cast<const float4*>(sx)[i];
  // ---- conv0 + bias + ReLU + pool: thread -> channel tid/16, windows (tid%16) + 16 j
  {
    const int ch = tid >> 4, sub = tid & 15;
    float w[28];
    load_w28(sw0 + ch * CD_W1PAD, w);
    const float bias = sb0[ch];
#pragma unroll 1
    for (int win = sub; win < CD_P1 * CD_P1; win += 16) {
      const int ph = win / CD_P1, pw = win - ph * CD_P1;
      float pt[6][6];
      load_patch(sx + (2 * ph) * CD_H + 2 * pw, CD_H, pt);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      window_fma(pt, w, acc);
      int bi;
      const float v = pool4(acc, bias, bi);
      sa1[ch * CD_P1 * CD_P1 + win] = v;
      if (g == 0) {
        const long long o = ((long long)b * CD_C1 + ch) * (CD_P1 * CD_P1) + win;
        a1[o] = v; idx1[o] = (unsigned char)bi;
      }
    }
  }
  __syncthreads();
  stamp(2);
  // ---- conv1: thread -> input-channel quarter cq, output-channel pair op, pooling window win; 2 oc x 4 pixels accumulators
  {
    const int cq = tid >> 6, t64 = tid & 63, op = t64 >> 4, win = t64 & 15, ph = win >> 2, pw = win & 3;
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int c = 4 * cq + cc;
      float pt[6][6], w[28];
      load_patch(sa1 + c * (CD_P1 * CD_P1) + (2 * ph) * CD_P1 + 2 * pw, CD_P1, pt);
      load_w28(sw1 + ((2 * op) * CD_C1 + c) * CD_W1PAD, w);
      window_fma(pt, w, acc0);
      load_w28(sw1 + ((2 * op + 1) * CD_C1 + c) * CD_W1PAD, w);
      window_fma(pt, w, acc1);
    }
    float4* dst = reinterpret_cast<float4*>(spart + (cq * 64 + t64) * 8);
    dst[0] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
    dst[1] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
  }
  __syncthreads();
  stamp(3);
  if (tid < 128) {
    const int i64 = tid >> 1, which = tid & 1, op = i64 >> 4, win = i64 & 15, ol = 2 * op + which;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int cq = 0; cq < 4; ++cq) {
      const float4 v = *reinterpret_cast<const float4*>(spart + (cq * 64 + i64) * 8 + which * 4);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
    int bi;
    const float v = pool4(acc, sb1[ol], bi);
    const long long o = ((long long)b * CD_C2 + 8 * g + ol) * (CD_P2 * CD_P2) + win;
    a2[o] = v; idx2[o] = (unsigned char)bi;
  }
  stamp(4);
}

// ------------------------------------------------------------------------------------------------------------------ backward (data + conv0)
// grid (B, 4), 288 threads.  CTA (b, cg): conv1-input channels [4cg, 4cg+4) of image b.
//   dz2 (32 x 8 x 8, one non-zero per pooling window where a2 > 0) is scattered into a zero-padded 32 x 16 x 16 plane in shared memory;
//   da1 = full correlation with the flipped filters = the forward window routine on the padded plane; da1 never leaves shared memory:
//   it is max-pool / ReLU routed straight into conv0's weight and bias gradient (the first layer needs no input gradient).
struct CnnBwdSmem {
  static constexpr int DZLD = 22;                                 // row stride of the padded dz2 plane: 16 + 6 floats keeps the 6x6 patch loads of a
                                                                 // warp (6 windows per row pair, rows 2*ld apart) on distinct banks
  static constexpr int DZPL = 16 * DZLD;                         // floats per plane
  static constexpr int DZ = 0;                                   // [32][16][DZLD]
  static constexpr int WT = DZ + CD_C2 * DZPL;                   // [4 cl][32 oc][28] flipped filters
  static constexpr int X = WT + 4 * CD_C2 * CD_W1PAD;            // [28][28]
  static constexpr int A1 = X + CD_H * CD_H;                     // [4][144] pooled conv0 activations (ReLU mask)
  static constexpr int DA1 = A1 + 4 * 144;                       // [4][144] -> routed gradient g
  static constexpr int PART = DA1 + 4 * 144;                     // [4 ocq][72][8]
  static constexpr int POS = PART + 4 * 72 * 8;                  // [4][144] int: top-left of the arg-max pixel in the 24x24 map (as x offset)
  static constexpr int TOTAL = POS + 4 * 144;
  static constexpr int BYTES = TOTAL * 4;
};

__device__ __forceinline__ void cnn_bwd_body(float* sm, const int b, const int cg, const float* __restrict__ x, const float* __restrict__ w1,
                                             const float* __restrict__ a1, const unsigned char* __restrict__ idx1, const float* __restrict__ a2,
                                             const unsigned char* __restrict__ idx2, const float* __restrict__ da2, float* __restrict__ dw0,
                                             float* __restrict__ db0, unsigned long long* dbg) {
  using L = CnnBwdSmem;
  float* sdz = sm + L::DZ; float* swt = sm + L::WT; float* sx = sm + L::X; float* sa1 = sm + L::A1; float* sda1 = sm + L::DA1;
  float* spart = sm + L::PART; int* spos = reinterpret_cast<int*>(sm + L::POS);
  const bool dbg_on = dbg != nullptr && b == 0 && cg == 0 && threadIdx.x == 0;
  auto stamp = [&](int slot) { if (dbg_on) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[slot] = t; } };
  stamp(8);
  const int tid = threadIdx.x;
  for (int i = tid; i < CD_C2 * L::DZPL / 4; i += 288) reinterpret_cast<float4*>(sdz)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 4 * CD_C2 * 25; i += 288) {
    // swt[(cl*32 + oc)*28 + t] = W1[oc][4cg+cl][4 - t/5][4 - t%5]   (t = flipped tap); source index runs over (oc, cl, tap) = 100 floats per oc
    const int oc = i / 100, rem = i - oc * 100, cl = rem / 25, tap = rem - cl * 25;
    cpa4(swt + (cl * CD_C2 + oc) * CD_W1PAD + (24 - tap), w1 + ((long long)oc * CD_C1 + 4 * cg) * 25 + rem);
  }
  for (int i = tid; i < 4 * CD_C2 * 3; i += 288) swt[(i / 3) * CD_W1PAD + 25 + i % 3] = 0.f;
  for (int i = tid; i < CD_H * CD_H / 4; i += 288) cpa16(sx + 4 * i, x + (long long)b * CD_H * CD_H + 4 * i);
  for (int i = tid; i < 4 * 144 / 4; i += 288) cpa16(sa1 + 4 * i, a1 + ((long long)b * CD_C1 + 4 * cg) * 144 + 4 * i);
  for (int i = tid; i < 4 * 144; i += 288) {
    const long long o = ((long long)b * CD_C1 + 4 * cg) * 144 + i;
    const int win = i % 144, id = idx1[o];
    spos[i] = (2 * (win / CD_P1) + (id >> 1)) * CD_H + 2 * (win % CD_P1) + (id & 1);
  }
  cpa_wait_all();
  __syncthreads();
  stamp(9);
  // Everything staged so far was produced by the forward kernel (complete long ago); da2 comes from the kernel right before this one.
  // The grid dependency is resolved HERE, so that with programmatic dependent launch these CTAs become resident on the SMs the 16-CTA
  // dense-chain cluster leaves idle and finish their staging while it is still running.
  pdl_wait();
  pdl_launch();
  // scatter the non-zeros of dz2 (pool + ReLU backward of conv1's output)
  for (int e = tid; e < CD_C2 * 16; e += 288) {
    const long long o = (long long)b * CD_C2 * 16 + e;
    const int oc = e >> 4, win = e & 15, id = idx2[o];
    const float gv = a2[o] > 0.f ? da2[o] : 0.f;
    const int oh = 2 * (win >> 2) + (id >> 1), ow = 2 * (win & 3) + (id & 1);
    sdz[oc * L::DZPL + (oh + 4) * L::DZLD + ow + 4] = gv;
  }
  __syncthreads();
  stamp(10);
  // ---- conv1 data gradient: thread -> oc quarter, channel pair, 2x2 output window of the 12x12 map
  {
    const int ocq = tid / 72, t72 = tid % 72, cp = t72 / 36, win = t72 % 36, ph = win / 6, pw = win % 6;
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int o8 = 0; o8 < 8; ++o8) {
      const int oc = 8 * ocq + o8;
      float pt[6][6], w[28];
      load_patch(sdz + oc * L::DZPL + (2 * ph) * L::DZLD + 2 * pw, L::DZLD, pt);
      load_w28(swt + ((2 * cp) * CD_C2 + oc) * CD_W1PAD, w);
      window_fma(pt, w, acc0);
      load_w28(swt + ((2 * cp + 1) * CD_C2 + oc) * CD_W1PAD, w);
      window_fma(pt, w, acc1);
    }
    float4* dst = reinterpret_cast<float4*>(spart + (ocq * 72 + t72) * 8);
    dst[0] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
    dst[1] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
  }
  __syncthreads();
  stamp(11);
  if (tid < 144) {
    const int i72 = tid >> 1, which = tid & 1, cp = i72 / 36, win = i72 % 36, ph = win / 6, pw = win % 6, cl = 2 * cp + which;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(spart + (q * 72 + i72) * 8 + which * 4);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
    // route through conv0's ReLU (+ pool: the value belongs to the arg-max pixel recorded in spos)
    float* d = sda1 + cl * 144 + (2 * ph) * CD_P1 + 2 * pw;
    const float* m = sa1 + cl * 144 + (2 * ph) * CD_P1 + 2 * pw;
    d[0] = m[0] > 0.f ? acc[0] : 0.f; d[1] = m[1] > 0.f ? acc[1] : 0.f;
    d[CD_P1] = m[CD_P1] > 0.f ? acc[2] : 0.f; d[CD_P1 + 1] = m[CD_P1 + 1] > 0.f ? acc[3] : 0.f;
  }
  __syncthreads();
  stamp(12);
  // ---- conv0 weight / bias gradient of this image (4 channels): dW0[ch][kh][kw] += sum_windows g * x[oh+kh][ow+kw]
  // 200 threads: (tap, channel) x 2 halves of the 144 windows; four independent accumulators break the FMA dependency chain
  if (tid < 200) {
    const int half = tid / 100, t = tid - half * 100, cl = t / 25, tap = t % 25, koff = (tap / 5) * CD_H + tap % 5;
    const float* gp = sda1 + cl * 144 + half * 72;
    const int* pp = spos + cl * 144 + half * 72;
    float a0 = 0.f, a1_ = 0.f, a2_ = 0.f, a3 = 0.f;
#pragma unroll 2
    for (int p = 0; p < 72; p += 4) {
      a0 = fmaf(gp[p], sx[pp[p] + koff], a0);
      a1_ = fmaf(gp[p + 1], sx[pp[p + 1] + koff], a1_);
      a2_ = fmaf(gp[p + 2], sx[pp[p + 2] + koff], a2_);
      a3 = fmaf(gp[p + 3], sx[pp[p + 3] + koff], a3);
    }
    atomicAdd(dw0 + (4 * cg + cl) * 25 + tap, (a0 + a1_) + (a2_ + a3));
  } else if (tid >= 224 && tid < 228) {
    const int cl = tid - 224;
    float acc = 0.f;
    for (int p = 0; p < 144; ++p) acc += sda1[cl * 144 + p];
    atomicAdd(db0 + 4 * cg + cl, acc);
  }
  stamp(13);
}

// ------------------------------------------------------------------------------------------------------------------ conv1 weight gradient
// grid (32 oc, 4 cg), 256 threads.  CTA owns dW1[oc][4cg..4cg+3][5][5] (100 outputs, written once: no atomics) and sums over all images:
//   dW1[oc][c][kh][kw] = sum_{b, window} g[b][oc][window] * a1[b][c][oh+kh][ow+kw]     (g != 0 only at the window's arg-max where a2 > 0)
__device__ __forceinline__ void cnn_wgrad1_body(float* sm, const int oc, const int cg, const float* __restrict__ a1, const float* __restrict__ a2,
                                                const unsigned char* __restrict__ idx2, const float* __restrict__ da2, float* __restrict__ dw1,
                                                float* __restrict__ db1, int B, unsigned long long* dbg) {
  float* sa = sm;                               // [B][4][144] the four input-channel planes of every image
  float* sg = sa + (size_t)B * 576;             // [B*16] routed gradients of this output channel
  int* sp = reinterpret_cast<int*>(sg + B * 16);  // [B*16] arg-max position as offset into a 12x12 plane
  float* sacc = reinterpret_cast<float*>(sp + B * 16);   // [2][128]
  const bool dbg_on = dbg != nullptr && oc == 0 && cg == 0 && threadIdx.x == 0;
  auto stamp = [&](int slot) { if (dbg_on) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[slot] = t; } };
  stamp(16);
  const int tid = threadIdx.x;
  for (int i = tid; i < B * 144; i += 256) {          // 144 float4 per image: 4 contiguous planes
    const int bb = i / 144, r = i % 144;
    cpa16(sa + 4 * i, a1 + ((long long)bb * CD_C1 + 4 * cg) * 144 + 4 * r);
  }
  for (int e = tid; e < B * 16; e += 256) {
    const int bb = e >> 4, win = e & 15;
    const int id = idx2[((long long)bb * CD_C2 + oc) * 16 + win];
    sp[e] = (2 * (win >> 2) + (id >> 1)) * CD_P1 + 2 * (win & 3) + (id & 1);
  }
  pdl_wait();            // (see cnn_bwd_body: only da2 depends on the preceding kernel)
  pdl_launch();
  for (int e = tid; e < B * 16; e += 256) {
    const long long o = ((long long)(e >> 4) * CD_C2 + oc) * 16 + (e & 15);
    sg[e] = a2[o] > 0.f ? da2[o] : 0.f;
  }
  cpa_wait_all();
  __syncthreads();
  stamp(17);
  const int half = tid >> 7, t = tid & 127;
  float acc = 0.f;
  if (t < 100) {
    const int cl = t / 25, tap = t % 25, koff = (tap / 5) * CD_P1 + tap % 5;
    const int e0 = half * (B * 8), e1 = e0 + B * 8;      // images [half*B/2, (half+1)*B/2)
    // four independent accumulators (the chain of B*8 dependent FMAs was the critical path of this phase)
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    const float* sak = sa + cl * 144 + koff;
#pragma unroll 2
    for (int e = e0; e < e1; e += 4) {       // B*8 is a multiple of 16; e .. e+3 belong to the same image
      const float* img = sak + (e >> 4) * 576;
      c0 = fmaf(sg[e], img[sp[e]], c0);
      c1 = fmaf(sg[e + 1], img[sp[e + 1]], c1);
      c2 = fmaf(sg[e + 2], img[sp[e + 2]], c2);
      c3 = fmaf(sg[e + 3], img[sp[e + 3]], c3);
    }
    acc = (c0 + c1) + (c2 + c3);
  }
  sacc[half * 128 + t] = acc;
  __syncthreads();
  if (tid < 100) dw1[((long long)oc * CD_C1 + 4 * cg + tid / 25) * 25 + tid % 25] = sacc[tid] + sacc[128 + tid];
  if (cg == 0 && tid < 32) {                 // db1[oc] = sum of all routed gradients of this channel
    float s = 0.f;
    for (int e = tid; e < B * 16; e += 32) s += sg[e];
    s = warp_sum(s);
    if (tid == 0) db1[oc] = s;
  }
  stamp(18);
}

__global__ void __launch_bounds__(288) cnn_bwd_kernel(const float* x, const float* w1, const float* a1, const unsigned char* idx1, const float* a2,
                                                       const unsigned char* idx2, const float* da2, float* dw0, float* db0, unsigned long long* dbg) {
  extern __shared__ __align__(16) float sm[];
  cnn_bwd_body(sm, blockIdx.x, blockIdx.y, x, w1, a1, idx1, a2, idx2, da2, dw0, db0, dbg);
}
__global__ void __launch_bounds__(256) cnn_wgrad1_kernel(const float* a1, const float* a2, const unsigned char* idx2, const float* da2, float* dw1,
                                                          float* db1, int B, unsigned long long* dbg) {
  extern __shared__ __align__(16) float sm[];
  cnn_wgrad1_body(sm, blockIdx.x, blockIdx.y, a1, a2, idx2, da2, dw1, db1, B, dbg);
}
// The whole convolution backward pass as ONE launch: CTAs [0, 4B) run the data-gradient / conv0 body, CTAs [4B, 4B + 128) the conv1
// weight-gradient body (two independent jobs that read the same inputs; a single heterogeneous grid saves a stream fork / join in the step graph).
__global__ void __launch_bounds__(288) cnn_bwd_all_kernel(const float* x, const float* w1, const float* a1, const unsigned char* idx1, const float* a2,
                                                           const unsigned char* idx2, const float* da2, float* dw0, float* db0, float* dw1, float* db1,
                                                           int B, unsigned long long* dbg) {
  extern __shared__ __align__(16) float sm[];
  const int i = blockIdx.x;      // (the grid dependency is resolved inside the bodies, after their independent staging)
  // debug: every CTA records (start, end, SM id) at dbg[64 + 3 i ..] so that tools/step_timeline.py can show how the grid packs onto the SMs
  unsigned long long t0 = 0;
  if (dbg != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  if (i < 4 * B) {
    cnn_bwd_body(sm, i >> 2, i & 3, x, w1, a1, idx1, a2, idx2, da2, dw0, db0, dbg);
  } else if (threadIdx.x < 256) {            // (no thread of this CTA takes the other branch, so its __syncthreads pair up among these 256)
    const int j = i - 4 * B;
    cnn_wgrad1_body(sm, j >> 2, j & 3, a1, a2, idx2, da2, dw1, db1, B, dbg);
  }
  if (dbg != nullptr && threadIdx.x == 0) {
    unsigned long long t1; unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    dbg[64 + 3 * i] = t0; dbg[65 + 3 * i] = t1; dbg[66 + 3 * i] = smid;
  }
}

// The convolution backward pass AND the exchange of the four conv keys (push -> both server tiers -> optimizer -> pull, the replicated one-hop
// mode of hips_fabric.cu's direct protocol) as ONE launch.  Grid and bodies are those of cnn_bwd_all_kernel.  Every CTA takes a ticket when its
// gradients are written; the CTAs that draw the last `n_active` tickets wait until the ticket counter shows that the whole grid has finished
// and then each serve one 1024-float tile of the conv keys: gradient -> LL packets into every rank's slot, poll the packets of all ranks, sum
// in the hierarchy's order (party sums, push_scale, sum over parties), optimizer on this rank's replica of the server state, fresh weights
// into the parameter arena.  Compared with a separate exchange launch this removes the grid drain + launch between the last gradient store and
// the first packet (4-5 us of the step's critical path on B200): the tail CTAs are already running when the last gradient lands.
//   state words of the channel (FabricParams::state): [0] round, [1] tail completion counter, [2] optimizer step, [5] protocol error, [6] tickets.
// All 256 CTAs are co-resident (2 per SM), so spinning on the ticket counter cannot starve the CTAs it waits for.
__global__ void __launch_bounds__(288) cnn_bwd_exchange_kernel(const float* x, const float* w1, const float* a1, const unsigned char* idx1, const float* a2,
                                                                const unsigned char* idx2, const float* da2, float* dw0, float* db0, float* dw1, float* db1,
                                                                int B, unsigned long long* dbg, const __grid_constant__ FabricParams p,
                                                                const int* __restrict__ tile_list, int n_active) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_ticket;
  const int i = blockIdx.x;
  if (i < 4 * B) {
    cnn_bwd_body(sm, i >> 2, i & 3, x, w1, a1, idx1, a2, idx2, da2, dw0, db0, dbg);
  } else if (threadIdx.x < 256) {
    const int j = i - 4 * B;
    cnn_wgrad1_body(sm, j >> 2, j & 3, a1, a2, idx2, da2, dw1, db1, B, dbg);
  }
  if (threadIdx.x >= FAB_THREADS) return;                       // the exchange tail is written for 256 threads (one float4 of a tile each)
  // the channel's round / optimizer step (they can only advance after every CTA of this launch has drawn its ticket below)
  const uint32_t round = (uint32_t)(*reinterpret_cast<volatile int*>(p.state)) + 1u;
  const int opt_t = (*reinterpret_cast<volatile int*>(p.state + 2)) + 1;
  auto bar = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
  bar();                                                         // every gradient store / atomic of this CTA has been issued
  if (threadIdx.x == 0) { __threadfence(); s_ticket = atomicAdd(p.state + 6, 1); }
  bar();
  const int first = (int)gridDim.x - n_active;
  if (s_ticket < first) return;
  const bool dbg_on = dbg != nullptr && s_ticket == first && threadIdx.x == 0;
  auto stamp = [&](int slot) { if (dbg_on) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[slot] = t; } };
  stamp(20);
  if (threadIdx.x == 0) {
    while (*reinterpret_cast<volatile int*>(p.state + 6) < (int)gridDim.x) {}
    __threadfence();
  }
  bar();
  stamp(21);
  const int t = tile_list[s_ticket - first];
  const uint32_t epoch = (round << 3) | (uint32_t)(p.channel_id & 7);
  const int S = p.party_size, P = p.num_parties;
  const long long n2 = 2 * p.n;
  int* err = p.state + 5;
  int f = p.tile_fmt ? (int)p.tile_fmt[t] : FMT_F32;
  if (f != FMT_F16 && f != FMT_F8) f = FMT_F32;
  const long long off = (long long)t * TILE + threadIdx.x * 4;
  float* g = p.grad[p.rank] + off;
  const float4 v = __ldcg(reinterpret_cast<const float4*>(g));  // written by other CTAs of this launch: read at L2
  if (p.zero_grad) *reinterpret_cast<float4*>(g) = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool wire = p.ll_d[p.rank] != nullptr;                   // false: single rank without loop-back — nothing to exchange
  if (wire) {
    const long long slot = (long long)p.rank * n2 + 2 * off;
    if (p.ll_d_mc != nullptr) ll_send_dense_mc(p.ll_d_mc + slot, v, f, epoch);
    else for (int r = 0; r < p.world; ++r) ll_send_dense(p.ll_d[r] + slot, v, f, epoch);
  }
  stamp(22);
  float lr = adam_lr(p.h, opt_t), wd = p.h.wd;
  if (p.tile_mult) { const float2 mm = __ldg(p.tile_mult + t); lr *= mm.x; wd *= mm.y; }
  float4 W = *reinterpret_cast<float4*>(p.w + off);             // state loads overlap the flight of the packets
  float4 A = p.s0 ? *reinterpret_cast<float4*>(p.s0 + off) : make_float4(0, 0, 0, 0);
  float4 Bm = p.s1 ? *reinterpret_cast<float4*>(p.s1 + off) : make_float4(0, 0, 0, 0);
  float4 agg = f4_scale(v, p.push_scale);
  if (wire) {
    const float* in = p.ll_d[p.rank] + 2 * off;
    for (int gq = 0; gq < P; ++gq) {
      float4 ps = ll_recv_dense(in + (long long)(gq * S) * n2, f, epoch, err);
      for (int j = 1; j < S; ++j) ps = f4_add(ps, ll_recv_dense(in + (long long)(gq * S + j) * n2, f, epoch, err));
      ps = f4_scale(ps, p.push_scale);
      agg = gq == 0 ? ps : f4_add(agg, ps);
    }
  }
  stamp(23);
  opt_apply(W.x, agg.x, A.x, Bm.x, p.h, lr, wd);
  opt_apply(W.y, agg.y, A.y, Bm.y, p.h, lr, wd);
  opt_apply(W.z, agg.z, A.z, Bm.z, p.h, lr, wd);
  opt_apply(W.w, agg.w, A.w, Bm.w, p.h, lr, wd);
  *reinterpret_cast<float4*>(p.w + off) = W;
  if (p.s0) *reinterpret_cast<float4*>(p.s0 + off) = A;
  if (p.s1) *reinterpret_cast<float4*>(p.s1 + off) = Bm;
  *reinterpret_cast<float4*>(p.param[p.rank] + off) = W;        // the pull is a local store: every rank holds the fresh replica
  if (threadIdx.x == 0) {
    const int done = atomicAdd(p.state + 1, 1);
    if (done == n_active - 1) {                                  // every tail CTA is past the ticket spin: the counters can be re-armed
      p.state[1] = 0;
      p.state[6] = 0;
      p.state[2] = opt_t;
      p.state[0] = (int)round;
    }
  }
  stamp(24);
}

}  // namespace gx

using namespace gx;

static unsigned long long* g_cnn_dbg = nullptr;      // optional %globaltimer stamps of CTA (0,0): fwd 0-4, bwd 8-13, wgrad1 16-18
GX_API int gx_cnn_set_debug(unsigned long long* p) { g_cnn_dbg = p; return 0; }

// x [B,1,28,28]; w0 [16,1,5,5]; w1 [32,16,5,5]; writes a1 [B,16,12,12] + idx1, a2 [B,32,4,4] + idx2
GX_API int gx_cnn_fwd(const float* x, const float* w0, const float* b0, const float* w1, const float* b1, float* a1, unsigned char* idx1, float* a2,
                      unsigned char* idx2, const float* carry_src, float* carry_dst, int carry_n, float* x_keep, int B, cudaStream_t s) {
  if (B < 1) return 0;
  launch_pdl(cnn_fwd_kernel, dim3(B, 4), dim3(256), 0, s, x, w0, b0, w1, b1, a1, idx1, a2, idx2, carry_src, carry_dst, carry_n, x_keep, g_cnn_dbg);
  return GX_CHECK_LAUNCH();
}
// accumulates into dw0 [16,25] / db0 [16] (atomics: zero them first)
GX_API int gx_cnn_bwd(const float* x, const float* w1, const float* a1, const unsigned char* idx1, const float* a2, const unsigned char* idx2,
                      const float* da2, float* dw0, float* db0, int B, cudaStream_t s) {
  if (B < 1) return 0;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(cnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CnnBwdSmem::BYTES);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  launch_pdl(cnn_bwd_kernel, dim3(B, 4), dim3(288), (size_t)CnnBwdSmem::BYTES, s, x, w1, a1, idx1, a2, idx2, da2, dw0, db0, g_cnn_dbg);
  return GX_CHECK_LAUNCH();
}
// overwrites dw1 [32,16,5,5] and db1 [32].  B even, B <= 64 (shared-memory planes of all images).
GX_API int gx_cnn_wgrad1(const float* a1, const float* a2, const unsigned char* idx2, const float* da2, float* dw1, float* db1, int B, cudaStream_t s) {
  if (B < 2 || (B & 1) || B > 64) return -1;
  const size_t smem = ((size_t)B * 576 + (size_t)B * 16 * 2 + 256) * 4;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(cnn_wgrad1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = smem;
  }
  launch_pdl(cnn_wgrad1_kernel, dim3(CD_C2, 4), dim3(256), smem, s, a1, a2, idx2, da2, dw1, db1, B, g_cnn_dbg);
  return GX_CHECK_LAUNCH();
}

// conv backward in one launch (see cnn_bwd_all_kernel): accumulates dw0 / db0 (zero them first), overwrites dw1 / db1.  B even, <= 64.
GX_API int gx_cnn_bwd_all(const float* x, const float* w1, const float* a1, const unsigned char* idx1, const float* a2, const unsigned char* idx2,
                          const float* da2, float* dw0, float* db0, float* dw1, float* db1, int B, cudaStream_t s) {
  if (B < 2 || (B & 1) || B > 64) return -1;
  const size_t smem_w = ((size_t)B * 576 + (size_t)B * 16 * 2 + 256) * 4;
  const size_t smem = smem_w > (size_t)CnnBwdSmem::BYTES ? smem_w : (size_t)CnnBwdSmem::BYTES;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(cnn_bwd_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = smem;
  }
  launch_pdl(cnn_bwd_all_kernel, dim3(4 * B + 4 * CD_C2), dim3(288), smem, s, x, w1, a1, idx1, a2, idx2, da2, dw0, db0, dw1, db1, B, g_cnn_dbg);
  return GX_CHECK_LAUNCH();
}

// conv backward + exchange of the conv keys in one launch (see cnn_bwd_exchange_kernel).  `params`: FabricParams block of the channel
// (replicated mode), `tile_list` / `n_active`: the channel's tiles.  Returns -1 for shapes the kernel does not cover.
GX_API int gx_cnn_bwd_exchange(const float* x, const float* w1, const float* a1, const unsigned char* idx1, const float* a2, const unsigned char* idx2,
                               const float* da2, float* dw0, float* db0, float* dw1, float* db1, int B, const void* params, const int* tile_list,
                               int n_active, cudaStream_t s) {
  if (B < 2 || (B & 1) || B > 64) return -1;
  const int grid = 4 * B + 4 * CD_C2;
  if (n_active < 1 || n_active > grid) return -1;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (grid > 2 * sms) return -1;                                  // the ticket wait needs the whole grid resident (2 CTAs per SM)
  const size_t smem_w = ((size_t)B * 576 + (size_t)B * 16 * 2 + 256) * 4;
  const size_t smem = smem_w > (size_t)CnnBwdSmem::BYTES ? smem_w : (size_t)CnnBwdSmem::BYTES;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(cnn_bwd_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = smem;
  }
  const FabricParams p = *reinterpret_cast<const FabricParams*>(params);
  launch_pdl(cnn_bwd_exchange_kernel, dim3(grid), dim3(288), smem, s, x, w1, a1, idx1, a2, idx2, da2, dw0, db0, dw1, db1, B, g_cnn_dbg, p, tile_list,
             n_active);
  return GX_CHECK_LAUNCH();
}
