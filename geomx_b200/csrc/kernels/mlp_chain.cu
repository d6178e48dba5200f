// Fused small-batch MLP head: dense0 -> dense1 -> classifier -> softmax-CE, forward AND backward, in ONE thread-block-cluster launch.
//
// Reference path (examples/cnn.py:60-64 Dense(256,relu)/Dense(128,relu)/Dense(10) + SoftmaxCrossEntropyLoss, run as FullyConnected
// forward/backward ops src/operator/nn/fully_connected-inl.h:71-173 -> cublasSgemmEx linalg_impl.h:196-214, log_softmax softmax-inl.h:166-250,
// pick broadcast_reduce_op_index.cu:39-45): 3 forward GEMMs + 6 backward GEMMs + bias / activation / softmax kernels, each a separate engine op.
//
// With a per-worker batch of 32 every one of those GEMMs has M = 32: a tcgen05 tile would be 3/4 padding and each launch is pure latency
// (profiles/round1/prof_step_ncu_raw.md: tensor pipe 1-5 %, 6-10 us per launch).  Here the whole chain is one kernel on a cluster of CS CTAs
// (16 — a non-portable cluster size, opted into at launch — or 8 when the device cannot co-schedule 16; GEOMX_MLP_CLUSTER):
//   * every CTA owns a 1/CS column slice of each layer; weights are staged in shared memory ONCE (cp.async, issued before griddepcontrol.wait
//     so the 150 KB of weight traffic overlaps the tail of the convolution kernels) and reused by forward and backward;
//   * products run on the fp32 FMA pipes (exact fp32, like the reference's SGEMM): 4x4 register tiles, K split over warps, operands read with
//     conflict-free 128-bit shared loads from XOR-swizzled tiles, partial sums reduced through shared memory;
//   * layer outputs are broadcast to all CTAs of the cluster with distributed-shared-memory stores (st.shared::cluster through
//     cluster.map_shared_rank) followed by one hardware cluster barrier — no global round trip between layers;
//   * epilogues fused: bias, ReLU, ReLU masks of the backward pass, softmax / loss / dlogits, bias gradients (column sums).
//   * 512 threads: the two warp-groups split K in the wide forward layer and run data-gradient and weight-gradient products side by side
//     in the backward phases.
// A handful of cluster barriers replace what used to be seven kernel launches (dense0, dense1, head, dW1, dz3, dW0, da2); measured 24.5 us
// for the whole chain forward + backward on B200 (profiles/kernel_times.txt, with per-phase %globaltimer stamps).
//
// Shapes are compile-time (D0 -> D1 -> D2 -> C<=16, batch <= 32): the demo CNN's 512 -> 256 -> 128 -> 10.  Other shapes use the tcgen05 GEMMs.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace gx {

constexpr int MC_THREADS = 512;   // 16 warps: two warp-groups of 256 that either split K (P1) or run data- and weight-gradient side by side (P4, P5)
constexpr int MC_HALF = 256;
constexpr int MC_B = 32;      // batch rows held (rows >= B are zero)
constexpr int MC_CMAX = 16;   // classes (padded)

struct MlpChainParams {
  const float* x;       // [B][D0]   input activations (flattened pooled conv output)
  const float* w0; const float* b0;   // [D1][D0], [D1]
  const float* w1; const float* b1;   // [D2][D1], [D2]
  const float* w2; const float* b2;   // [C][D2],  [C]
  const float* label;   // [B] class index as float
  float* loss;          // [B]
  float* logits;        // [B][C] or nullptr
  float* dw0; float* db0; float* dw1; float* db1; float* dw2; float* db2;
  float* dx;            // [B][D0]   gradient w.r.t. x
  int B, C;
  unsigned long long* dbg;   // optional %globaltimer stamps of cluster CTA 0 (tools/kernel_times.py)
};

// 16-byte chunk swizzle of a K-major [rows][K] tile: chunk (k>>2) is XORed with (row>>2)&7, so the 8 row-groups a warp touches in one
// 128-bit load land in 8 different bank groups
__device__ __forceinline__ int sw_off(int row, int k, int K) { return row * K + ((((k >> 2) ^ ((row >> 2) & 7))) << 2) + (k & 3); }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage `rows` x K floats (global row stride ld) into a swizzled K-major tile
__device__ __forceinline__ void stage_swizzled_async(float* dst, const float* src, int rows, int K, long long ld) {
  const int chunks = rows * (K >> 2);
  for (int c = threadIdx.x; c < chunks; c += MC_THREADS) {
    const int r = c / (K >> 2), q = c - r * (K >> 2);
    cp_async16(dst + sw_off(r, q << 2, K), src + (long long)r * ld + (q << 2));
  }
}
// stage `rows` x `cols` floats into a plain row-major tile [rows][cols]
__device__ __forceinline__ void stage_plain_async(float* dst, const float* src, int rows, int cols, long long ld) {
  const int chunks = rows * (cols >> 2);
  for (int c = threadIdx.x; c < chunks; c += MC_THREADS) {
    const int r = c / (cols >> 2), q = c - r * (cols >> 2);
    cp_async16(dst + r * cols + (q << 2), src + (long long)r * ld + (q << 2));
  }
}

// ---- register-tile products.  Thread (tb, tn, ks): rows 4tb..4tb+3 of X, 4 output columns, K chunks ks, ks+KS, ...
// NT: W is a swizzled K-major tile [NC][K]
template <int K, int KS>
__device__ __forceinline__ void prod_nt(const float* __restrict__ sX, const float* __restrict__ sW, int tb, int tn, int ks, float (&acc)[4][4]) {
#pragma unroll 2
  for (int kq = ks; kq < K / 4; kq += KS) {
    float4 x[4], w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(sX + (4 * tb + i) * K + ((kq ^ (tb & 7)) << 2));
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(sW + (4 * tn + j) * K + ((kq ^ (tn & 7)) << 2));
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(x[i].x, w[j].x, acc[i][j]);
        acc[i][j] = fmaf(x[i].y, w[j].y, acc[i][j]);
        acc[i][j] = fmaf(x[i].z, w[j].z, acc[i][j]);
        acc[i][j] = fmaf(x[i].w, w[j].w, acc[i][j]);
      }
  }
}
// NN: W is a plain tile [K][NC] (output columns contiguous): out[b][n] = sum_k X[b][k] * W[k][n]
template <int K, int NC, int KS>
__device__ __forceinline__ void prod_nn(const float* __restrict__ sX, const float* __restrict__ sW, int tb, int tn, int ks, float (&acc)[4][4]) {
#pragma unroll 2
  for (int kq = ks; kq < K / 4; kq += KS) {
    float4 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(sX + (4 * tb + i) * K + ((kq ^ (tb & 7)) << 2));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(sW + (4 * kq + kk) * NC + 4 * tn);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xv = kk == 0 ? x[i].x : kk == 1 ? x[i].y : kk == 2 ? x[i].z : x[i].w;
        acc[i][0] = fmaf(xv, w.x, acc[i][0]);
        acc[i][1] = fmaf(xv, w.y, acc[i][1]);
        acc[i][2] = fmaf(xv, w.z, acc[i][2]);
        acc[i][3] = fmaf(xv, w.w, acc[i][3]);
      }
    }
  }
}
// write this thread's 4x4 partial tile into scratch [KS][32][NC]
template <int NC>
__device__ __forceinline__ void put_partial(float* scratch, int tb, int tn, int ks, const float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(scratch + (ks * MC_B + 4 * tb + i) * NC + 4 * tn) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}
template <int NC, int KS>
__device__ __forceinline__ float4 sum_partials(const float* scratch, int b, int n4) {
  float4 s = *reinterpret_cast<const float4*>(scratch + b * NC + 4 * n4);
#pragma unroll
  for (int k = 1; k < KS; ++k) {
    const float4 t = *reinterpret_cast<const float4*>(scratch + (k * MC_B + b) * NC + 4 * n4);
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  return s;
}
// wgrad: dW[n][k] = sum_b DZ[b][nbase+n] * X[b][k]   (DZ: swizzled [32][ND], X: swizzled [32][KD]); thread tile TN x 4, coalesced global store
template <int ND, int KD, int TN>
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ sDZ, const float* __restrict__ sX, int nbase, int n0, int tk, float* __restrict__ out_row0) {
  float acc[TN][4];
#pragma unroll
  for (int i = 0; i < TN; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll 4
  for (int b = 0; b < MC_B; ++b) {
    const int key = (b >> 2) & 7;
    const float4 x = *reinterpret_cast<const float4*>(sX + b * KD + ((tk ^ key) << 2));
    float dz[TN];
#pragma unroll
    for (int q = 0; q < TN / 4; ++q) {
      const float4 d = *reinterpret_cast<const float4*>(sDZ + b * ND + ((((nbase + n0) >> 2) + q) ^ key) * 4);
      dz[4 * q] = d.x; dz[4 * q + 1] = d.y; dz[4 * q + 2] = d.z; dz[4 * q + 3] = d.w;
    }
#pragma unroll
    for (int i = 0; i < TN; ++i) {
      acc[i][0] = fmaf(dz[i], x.x, acc[i][0]);
      acc[i][1] = fmaf(dz[i], x.y, acc[i][1]);
      acc[i][2] = fmaf(dz[i], x.z, acc[i][2]);
      acc[i][3] = fmaf(dz[i], x.w, acc[i][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < TN; ++i)
    *reinterpret_cast<float4*>(out_row0 + (long long)(n0 + i) * KD + 4 * tk) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

template <int D0, int D1, int D2, int CS>
struct MlpSmem {
  static constexpr int NS1 = D1 / CS, NS2 = D2 / CS, KS0 = D0 / CS;
  // float offsets
  static constexpr int X = 0;                                  // [32][D0] swizzled
  static constexpr int W0 = X + MC_B * D0;                     // fwd: [NS1][D0] swizzled;  bwd: [D1][KS0] plain
  static constexpr int W0_SZ = (NS1 * D0 > D1 * KS0) ? NS1 * D0 : D1 * KS0;
  static constexpr int W1 = W0 + W0_SZ;                        // fwd: [NS2][D1] swizzled;  bwd: [D2][NS1] plain
  static constexpr int W1_SZ = (NS2 * D1 > D2 * NS1) ? NS2 * D1 : D2 * NS1;
  static constexpr int A3 = W1 + W1_SZ;                        // [32][D1] swizzled  (later: dz3)
  static constexpr int A4 = A3 + MC_B * D1;                    // [32][D2] swizzled  (later: scratch)
  static constexpr int DZ4 = A4 + MC_B * D2;                   // [32][D2] swizzled  (earlier: scratch)
  static constexpr int W2 = DZ4 + MC_B * D2;                   // [16][D2] plain
  static constexpr int DL = W2 + MC_CMAX * D2;                 // [32][16] dlogits
  static constexpr int MISC = DL + MC_B * MC_CMAX;             // b0 slice [NS1] | b1 slice [NS2] | b2 [16] | label [32]
  static constexpr int TOTAL = MISC + NS1 + NS2 + 16 + 32;
  static constexpr int BYTES = TOTAL * 4;
};

// CS = CTAs per cluster: 8 (portable) or 16 (non-portable size, one GPC): every layer is cut into CS column slices
template <int D0, int D1, int D2, int CS>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(MC_THREADS, 1) mlp_chain_kernel(const MlpChainParams p) {
  using L = MlpSmem<D0, D1, D2, CS>;
  constexpr int NS1 = L::NS1, NS2 = L::NS2, KS0 = L::KS0;
  constexpr int MC_CLUSTER = CS;
  static_assert(D0 % 64 == 0 && D1 % 64 == 0 && D2 % 64 == 0, "layer widths must be multiples of 64");
  static_assert(NS1 % 8 == 0 && NS2 % 4 == 0 && KS0 % 4 == 0, "slices must hold whole register tiles");
  static_assert(MC_THREADS % (2 * NS1) == 0 && MC_HALF % (2 * NS2) == 0 && MC_HALF % (2 * NS1) == 0 && MC_HALF % (2 * KS0) == 0, "K splits");
  static_assert(MC_THREADS / (2 * NS1) * MC_B * NS1 <= 2 * MC_B * D2 && MC_HALF / (2 * KS0) * MC_B * KS0 <= MC_B * D2, "scratch sizes");
  extern __shared__ __align__(16) float sm[];
  float* sX = sm + L::X; float* sW0 = sm + L::W0; float* sW1 = sm + L::W1; float* sA3 = sm + L::A3; float* sA4 = sm + L::A4;
  float* sDZ4 = sm + L::DZ4; float* sW2 = sm + L::W2; float* sDL = sm + L::DL;
  float* sB0 = sm + L::MISC; float* sB1 = sB0 + NS1; float* sB2 = sB1 + NS2; float* sLab = sB2 + 16;
  cg::cluster_group cluster = cg::this_cluster();
  const int cr = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = p.B, C = p.C;
  const bool dbg_on = p.dbg != nullptr && cr == 0 && tid == 0;
  auto stamp = [&](int slot) { if (dbg_on) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[slot] = t; } };
  stamp(0);

  // ---------------- prologue: weights do not depend on the preceding kernel (they were written by the previous step's exchange, which a
  // kernel further up the stream has already waited for) -> stage them before griddepcontrol.wait
  stage_swizzled_async(sW0, p.w0 + (long long)cr * NS1 * D0, NS1, D0, D0);
  stage_swizzled_async(sW1, p.w1 + (long long)cr * NS2 * D1, NS2, D1, D1);
  for (int c = tid; c < MC_CMAX * (D2 >> 2); c += MC_THREADS) {
    const int r = c / (D2 >> 2), q = c - r * (D2 >> 2);
    if (r < C) cp_async16(sW2 + r * D2 + (q << 2), p.w2 + (long long)r * D2 + (q << 2));
    else *reinterpret_cast<float4*>(sW2 + r * D2 + (q << 2)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid < NS1) sB0[tid] = p.b0[cr * NS1 + tid];
  if (tid < NS2) sB1[tid] = p.b1[cr * NS2 + tid];
  if (tid < 16) sB2[tid] = tid < C ? p.b2[tid] : 0.f;
  cp_async_commit();
  stamp(1);
  pdl_wait();
  pdl_launch();
  stamp(2);
  // input activations (full copy per CTA) + labels
  for (int c = tid; c < MC_B * (D0 >> 2); c += MC_THREADS) {
    const int r = c / (D0 >> 2), q = c - r * (D0 >> 2);
    if (r < B) cp_async16(sX + sw_off(r, q << 2, D0), p.x + (long long)r * D0 + (q << 2));
    else *reinterpret_cast<float4*>(sX + sw_off(r, q << 2, D0)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid < MC_B) sLab[tid] = tid < B ? p.label[tid] : 0.f;
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  cluster.sync();      // every CTA of the cluster is running (DSMEM stores below need the target CTA's shared memory to exist)
  stamp(3);

  // tile coordinates shared by the 32-row products: 8 row groups x (NC/4) column groups per K split
  // ---------------- P1: a3[:, slice] = relu(x * W0[slice]^T + b0)          NC = 32, K = D0, KS = 4
  {
    constexpr int PER = 2 * NS1, KS = MC_THREADS / PER;  // 8 row groups x NS1/4 column groups per split; K split over all 16 warps
    const int ks = tid / PER, tile = tid % PER, tb = tile & 7, tn = tile >> 3;
    float acc[4][4] = {};
    prod_nt<D0, KS>(sX, sW0, tb, tn, ks, acc);
    stamp(10);
    float* scratch = sA4;                                // [sA4 | sDZ4] = 32 KB contiguous, both unused until the end of P2 / P3
    put_partial<NS1>(scratch, tb, tn, ks, acc);
    __syncthreads();
    // W0's row slice is no longer needed: start fetching the COLUMN slice W0[:, cr*KS0 ...] for the backward pass (da2) underneath P2..P4
    stamp(11);
    stage_plain_async(sW0, p.w0 + cr * KS0, D1, KS0, D0);
    cp_async_commit();
    stamp(12);
    if (tid < MC_B * NS1 / 4) {
      const int b = tid / (NS1 / 4), n4 = tid % (NS1 / 4);   // 32 x NS1/4 float4 outputs
      float4 v = sum_partials<NS1, KS>(scratch, b, n4);
      v.x = fmaxf(v.x + sB0[4 * n4], 0.f); v.y = fmaxf(v.y + sB0[4 * n4 + 1], 0.f);
      v.z = fmaxf(v.z + sB0[4 * n4 + 2], 0.f); v.w = fmaxf(v.w + sB0[4 * n4 + 3], 0.f);
      if (b >= B) v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int o = sw_off(b, cr * NS1 + 4 * n4, D1);
#pragma unroll
      for (int r = 0; r < MC_CLUSTER; ++r) *reinterpret_cast<float4*>(cluster.map_shared_rank(sA3, r) + o) = v;
    }
    stamp(13);
  }
  cluster.sync();
  stamp(4);
  // ---------------- P2: a4[:, slice] = relu(a3 * W1[slice]^T + b1)         NC = 16, K = D1, KS = 8
  {
    constexpr int PER = 2 * NS2, KS = MC_HALF / PER;     // this layer is small: warp-group 0 only
    if (tid < MC_HALF) {
      const int ks = tid / PER, tile = tid % PER, tb = tile & 7, tn = tile >> 3;
      float acc[4][4] = {};
      prod_nt<D1, KS>(sA3, sW1, tb, tn, ks, acc);
      put_partial<NS2>(sDZ4, tb, tn, ks, acc);
    }
    __syncthreads();
    stage_plain_async(sW1, p.w1 + cr * NS1, D2, NS1, D1);   // W1[:, slice of D1] for dz3
    cp_async_commit();
    if (tid < MC_B * NS2 / 4) {
      const int b = tid / (NS2 / 4), n4 = tid % (NS2 / 4);
      float4 v = sum_partials<NS2, KS>(sDZ4, b, n4);
      v.x = fmaxf(v.x + sB1[4 * n4], 0.f); v.y = fmaxf(v.y + sB1[4 * n4 + 1], 0.f);
      v.z = fmaxf(v.z + sB1[4 * n4 + 2], 0.f); v.w = fmaxf(v.w + sB1[4 * n4 + 3], 0.f);
      if (b >= B) v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int o = sw_off(b, cr * NS2 + 4 * n4, D2);
#pragma unroll
      for (int r = 0; r < MC_CLUSTER; ++r) *reinterpret_cast<float4*>(cluster.map_shared_rank(sA4, r) + o) = v;
    }
  }
  cluster.sync();
  stamp(5);
  // ---------------- P3: classifier + softmax-CE (every CTA redundantly: 32 x C logits), dz4 slice, dW2 slice, db2
  {
    for (int b = warp; b < MC_B; b += MC_THREADS / 32) {
      // all class dot products of this row at once: 16 independent partial sums per lane, then one butterfly over the whole vector
      // (a shuffle reduction per class would be a chain of 5 dependent ~25-cycle shuffles, ten times over)
      float part[MC_CMAX];
#pragma unroll
      for (int c = 0; c < MC_CMAX; ++c) part[c] = 0.f;
#pragma unroll
      for (int k = lane; k < D2; k += 32) {
        const float av = sA4[sw_off(b, k, D2)];
#pragma unroll
        for (int c = 0; c < MC_CMAX; ++c) part[c] = fmaf(av, sW2[c * D2 + k], part[c]);      // rows >= C of sW2 are zero
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < MC_CMAX; ++c) part[c] += __shfl_xor_sync(0xffffffffu, part[c], o);
      float mylogit = 0.f;
#pragma unroll
      for (int c = 0; c < MC_CMAX; ++c) if (lane == c) mylogit = part[c] + sB2[c];
      float mx = lane < C ? mylogit : -INFINITY;
      mx = warp_max(mx);
      const float e = lane < C ? __expf(mylogit - mx) : 0.f;
      const float s = warp_sum(e);
      const int l = (int)sLab[b];
      if (lane < MC_CMAX) sDL[b * MC_CMAX + lane] = (lane < C && b < B) ? e / s - (lane == l ? 1.f : 0.f) : 0.f;
      const float picked = __shfl_sync(0xffffffffu, mylogit, l & 31);
      if (cr == 0 && b < B) {
        if (lane == 0) p.loss[b] = -(picked - mx - __logf(s));
        if (p.logits != nullptr && lane < C) p.logits[(long long)b * C + lane] = mylogit;
      }
    }
    __syncthreads();
    // dz4[b][m] = relu'(a4[b][m]) * sum_c dl[b][c] * W2[c][m]   for m in this CTA's slice of D2; broadcast to the cluster
    for (int e = tid; e < MC_B * NS2; e += MC_THREADS) {
      const int b = e / NS2, m = cr * NS2 + e % NS2;
      float s = 0.f;
      for (int c = 0; c < C; ++c) s = fmaf(sDL[b * MC_CMAX + c], sW2[c * D2 + m], s);
      const int o = sw_off(b, m, D2);
      if (!(sA4[o] > 0.f)) s = 0.f;
#pragma unroll
      for (int r = 0; r < MC_CLUSTER; ++r) cluster.map_shared_rank(sDZ4, r)[o] = s;
    }
    // dW2[c][m] = sum_b dl[b][c] * a4[b][m] (this CTA's m slice), db2 (CTA 0)
    for (int e = tid; e < C * NS2; e += MC_THREADS) {
      const int c = e / NS2, m = cr * NS2 + e % NS2;
      float s = 0.f;
      for (int b = 0; b < MC_B; ++b) s = fmaf(sDL[b * MC_CMAX + c], sA4[sw_off(b, m, D2)], s);
      p.dw2[(long long)c * D2 + m] = s;
    }
    if (cr == 0 && tid < C) {
      float s = 0.f;
      for (int b = 0; b < MC_B; ++b) s += sDL[b * MC_CMAX + tid];
      p.db2[tid] = s;
    }
  }
  cluster.sync();
  stamp(6);
  // ---------------- P4: dz3[:, slice] = relu'(a3) * (dz4 * W1[:, slice]);  dW1[slice of D2 rows] = dz4[:, rows]^T * a3;  db1
  float4 dz3v;   // this thread's float4 of the dz3 slice (kept in registers across the barrier that protects a3)
  {
    cp_async_wait<0>();          // W1 column slice (and W0 column slice) landed
    __syncthreads();
    if (tid < MC_HALF) {
      // warp-group 0: the data gradient dz3 (critical path)
      constexpr int PER = 2 * NS1, KS = MC_HALF / PER;   // K = D2
      const int ks = tid / PER, tile = tid % PER, tb = tile & 7, tn = tile >> 3;
      float acc[4][4] = {};
      prod_nn<D2, NS1, KS>(sDZ4, sW1, tb, tn, ks, acc);
      put_partial<NS1>(sA4, tb, tn, ks, acc);           // scratch = a4 buffer (dead after P3)
      named_bar_sync(1, MC_HALF);
      if (tid < MC_B * NS1 / 4) {
        const int b = tid / (NS1 / 4), n4 = tid % (NS1 / 4);
        dz3v = sum_partials<NS1, KS>(sA4, b, n4);
        const float4 a3 = *reinterpret_cast<const float4*>(sA3 + sw_off(b, cr * NS1 + 4 * n4, D1));
        dz3v.x = a3.x > 0.f ? dz3v.x : 0.f; dz3v.y = a3.y > 0.f ? dz3v.y : 0.f; dz3v.z = a3.z > 0.f ? dz3v.z : 0.f; dz3v.w = a3.w > 0.f ? dz3v.w : 0.f;
      }
    } else {
      // warp-group 1, concurrently: dW1 rows [cr*NS2, +NS2) as 4 x 4 tiles (4 x D1/4 = 256 of them) and the db1 slice
      const int t = tid - MC_HALF;
      if (t < (NS2 / 4) * (D1 / 4)) {
        const int tnw = t / (D1 / 4), tk = t % (D1 / 4);
        wgrad_tile<D2, D1, 4>(sDZ4, sA3, cr * NS2, 4 * tnw, tk, p.dw1 + (long long)cr * NS2 * D1);
      }
      if (t < NS2) {
        float s = 0.f;
        for (int bb = 0; bb < MC_B; ++bb) s += sDZ4[sw_off(bb, cr * NS2 + t, D2)];
        p.db1[cr * NS2 + t] = s;
      }
    }
  }
  stamp(7);
  cluster.sync();      // everybody is done reading a3 -> its buffer becomes dz3
  if (tid < MC_B * NS1 / 4) {
    const int b = tid / (NS1 / 4), n4 = tid % (NS1 / 4);
    const int o = sw_off(b, cr * NS1 + 4 * n4, D1);
#pragma unroll
    for (int r = 0; r < MC_CLUSTER; ++r) *reinterpret_cast<float4*>(cluster.map_shared_rank(sA3, r) + o) = dz3v;
  }
  cluster.sync();
  stamp(8);
  // ---------------- P5: dx[:, slice of D0] = dz3 * W0[:, slice];  dW0[slice of D1 rows] = dz3[:, rows]^T * x;  db0
  {
    float* sDZ3 = sA3;
    if (tid < MC_HALF) {
      // warp-group 0: the input gradient (what the convolution backward pass is waiting for)
      constexpr int PER = 2 * KS0, KS = MC_HALF / PER;   // K = D1
      const int ks = tid / PER, tile = tid % PER, tb = tile & 7, tn = tile >> 3;
      float acc[4][4] = {};
      prod_nn<D1, KS0, KS>(sDZ3, sW0, tb, tn, ks, acc);
      stamp(14);
      put_partial<KS0>(sDZ4, tb, tn, ks, acc);          // scratch = dz4 buffer (dead after P4): KS x 32 x KS0 floats = 16 KB
      named_bar_sync(1, MC_HALF);
      for (int e = tid; e < MC_B * KS0 / 4; e += MC_HALF) {
        const int b = e / (KS0 / 4), n4 = e % (KS0 / 4);
        if (b < B) *reinterpret_cast<float4*>(p.dx + (long long)b * D0 + cr * KS0 + 4 * n4) = sum_partials<KS0, KS>(sDZ4, b, n4);
      }
    } else {
      // warp-group 1, concurrently: dW0 rows [cr*NS1, +NS1) as 8 x 4 tiles (4 x D0/4 = 512 tiles, two per thread) and the db0 slice
      const int t0 = tid - MC_HALF;
      for (int t = t0; t < (NS1 / 8) * (D0 / 4); t += MC_HALF) {
        const int tnw = t / (D0 / 4), tk = t % (D0 / 4);
        wgrad_tile<D1, D0, 8>(sDZ3, sX, cr * NS1, 8 * tnw, tk, p.dw0 + (long long)cr * NS1 * D0);
      }
      if (t0 < NS1) {
        float s = 0.f;
        for (int bb = 0; bb < MC_B; ++bb) s += sDZ3[sw_off(bb, cr * NS1 + t0, D1)];
        p.db0[cr * NS1 + t0] = s;
      }
    }
  }
  stamp(9);
}

}  // namespace gx

using namespace gx;

// 512 -> 256 -> 128 -> C (C <= 16), B <= 32.  Returns -1 for unsupported shapes (caller uses the per-layer kernels).
template <int CS>
static int mlp_launch(const MlpChainParams& p, cudaStream_t stream) {
  using L = MlpSmem<512, 256, 128, CS>;
  auto kern = mlp_chain_kernel<512, 256, 128, CS>;
  static int ready = 0;      // 0 = not configured, 1 = ok, -1 = this cluster size cannot be scheduled on this device
  if (ready == 0) {
    ready = -1;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES) == cudaSuccess &&
        (CS <= 8 || cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess)) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CS); cfg.blockDim = dim3(MC_THREADS); cfg.dynamicSmemBytes = L::BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters >= 1) ready = 1;
    }
    cudaGetLastError();
  }
  if (ready < 0) return -2;
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("GEOMX_MLP_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  if (use_pdl) launch_pdl(kern, dim3(CS), dim3(MC_THREADS), (size_t)L::BYTES, stream, p);
  else kern<<<dim3(CS), dim3(MC_THREADS), (size_t)L::BYTES, stream>>>(p);
  return GX_CHECK_LAUNCH();
}

static unsigned long long* g_mlp_dbg = nullptr;
GX_API int gx_mlp_chain_set_debug(unsigned long long* p) { g_mlp_dbg = p; return 0; }

GX_API int gx_mlp_chain_fwd_bwd(const float* x, const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* label, float* loss, float* logits, float* dw0, float* db0, float* dw1, float* db1, float* dw2,
                                float* db2, float* dx, int B, int D0, int D1, int D2, int C, cudaStream_t stream) {
  if (D0 != 512 || D1 != 256 || D2 != 128 || C < 1 || C > MC_CMAX || B < 1 || B > MC_B) return -1;
  MlpChainParams p;
  p.x = x; p.w0 = w0; p.b0 = b0; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.label = label; p.loss = loss; p.logits = logits;
  p.dw0 = dw0; p.db0 = db0; p.dw1 = dw1; p.db1 = db1; p.dw2 = dw2; p.db2 = db2; p.dx = dx; p.B = B; p.C = C; p.dbg = g_mlp_dbg;
  // cluster of 16 CTAs (one GPC) halves every slice; fall back to the portable size 8 when the device cannot co-schedule 16
  static int want = -1;
  if (want < 0) { const char* e = getenv("GEOMX_MLP_CLUSTER"); want = e ? atoi(e) : 16; }
  if (want >= 16) {
    const int rc = mlp_launch<16>(p, stream);
    if (rc != -2) return rc;
    want = 8;
  }
  return mlp_launch<8>(p, stream);
}
GX_API int gx_mlp_chain_smem_bytes() { return MlpSmem<512, 256, 128, 8>::BYTES; }
