// tcgen05 / TMEM / TMA GEMM for sm_100a:   D[M,N] = epilogue( alpha * A[M,K] * B[N,K]^T )
//
//  * operands are fp32 in global memory, multiplied on the 5th-gen tensor cores as TF32 (kind::tf32) with fp32
//    accumulation in TMEM — the closest tensor-core analogue of the reference's cublasSgemmEx path
//    (reference src/operator/linalg_impl.h:196-214, FullyConnected src/operator/nn/fully_connected-inl.h:71-173,
//    Convolution im2col+GEMM src/operator/nn/convolution-inl.h:165-284).
//  * each operand may be K-major (row-major [rows][K]) or MN-major (row-major [K][rows]) so that forward, dgrad
//    (dY*W) and wgrad (dY^T*X) all read the tensors where they already live — no transposes are materialised.
//  * warp-specialised, one 128 x BLOCK_N tile per CTA: warp0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B,
//    4-stage mbarrier ring), warp1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2-5 = epilogue
//    (tcgen05.ld 32x32b -> registers -> fused bias / ReLU / ReLU-mask / column-sum (bias gradient) / NCHW scatter /
//    split-K atomic accumulate).
//
// Precision: the default is **3xTF32** (fp32-accurate): every operand tile that TMA lands in shared memory is split in place by the four
// (otherwise idle) epilogue warps into hi = rn_tf32(x) and lo = rn_tf32(x - hi); the MMA thread then issues lo*hi + hi*lo + hi*hi into the
// same fp32 TMEM accumulator.  The dropped lo*lo term and the rounding of lo are ~2^-22 relative per product, i.e. the result matches an
// fp32 SGEMM (the reference's cublasSgemmEx with CUDA_R_32F, linalg_impl.h:196-214) to fp32 rounding.  `gx_gemm_set_precision(1)` /
// GEOMX_GEMM_PRECISION=tf32 selects plain TF32 (one MMA per K step, no split stage).
//
// Tile shapes: UMMA M=128, N=BLOCK_N in {32,64,128}, K=8 per instruction (32 B of tf32), BLOCK_K = 32 fp32 = one
// 128-byte swizzle atom per row.  Out-of-bounds rows/cols/k are zero-filled by TMA, masked in the epilogue.
#include <cuda.h>

#include "common.cuh"

namespace gx {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;
constexpr int UMMA_K = 8;
constexpr int GEMM_THREADS = 192;         // TF32 mode: warp0 TMA, warp1 MMA, warps 2-5 epilogue
constexpr int GEMM_THREADS_SPLIT = 320;   // 3xTF32 mode: + warps 6-9 = hi/lo split pass (warps 2-5 drain the accumulator chunks)
constexpr int ACC_CHUNK_KB = 4;           // 3xTF32: K blocks accumulated in one TMEM buffer before it is drained into registers
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KiB either major
constexpr int MN_BOX_BYTES = 32 * BLOCK_K * 4;       // one 32(MN) x 32(K) box = 4 KiB

struct GemmParams {
  int M, N, K;
  int kb_per_split;
  float* D;
  long long ldd;
  const float* bias;
  const float* mask;
  long long ldmask;
  float* colsum;
  int relu, accumulate, store_mode, hw;
  int pool_w;              // store_mode 2: width of the (pre-pool) feature map; rows are (image, oh, ow)
  unsigned char* pool_idx; // store_mode 2: arg-max position (0..3) of every pooled element, same layout as D
  float alpha;
  int a_boxes, b_boxes;    // MN-major operands: number of 32-wide TMA boxes that are (partly) in bounds for this problem
  int a_bytes, b_bytes;    // bytes of the A / B tile that TMA actually fills per stage (what the 3xTF32 split pass has to touch)
  int tx_bytes;            // bytes that land per stage (A box(es) + B box(es)); out-of-range rows/boxes are never requested
  int stages;              // TMA ring depth actually used (<= SmemLayout::STAGES); smaller rings need less smem -> cheaper launch
  unsigned long long* dbg; // optional: %globaltimer stamps of CTA (0,0,0) phases (tools/kernel_times.py)
};

constexpr int MAX_STAGES = 8;
template <int BLOCK_N, bool SPLIT>
struct SmemLayout {
  // deep TMA ring: the problems this framework sees are latency-bound (few CTAs, cold operands), so as many K blocks as fit are kept in
  // flight: 8 stages for N<=64 (160/192 KiB), 6 for N=128 (192 KiB); the 3xTF32 variant keeps a lo copy of every tile => half as many
  static constexpr int STAGES = SPLIT ? (BLOCK_N <= 64 ? 4 : 3) : (BLOCK_N <= 64 ? 8 : 6);
  static_assert(BLOCK_N >= 16, "UMMA N >= 16 for M = 128");
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 4;
  static constexpr int OPS_BYTES = A_TILE_BYTES + B_TILE_BYTES;          // [A | B] as landed by TMA (= the hi parts after the split pass)
  static constexpr int STAGE_BYTES = SPLIT ? 2 * OPS_BYTES : OPS_BYTES;  // 3xTF32: [A_hi | B_hi | A_lo | B_lo]
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 512 + 1024;  // + barriers + bias tile + alignment slack
};

// hi = rn_tf32(x) (low 13 mantissa bits zero), lo = rn_tf32(x - hi): x = hi + lo up to ~2^-22 |x|
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}
// 128 threads split `bytes` of a tile in place (hi) and into the lo copy `lo_off` bytes further; element-wise => layout/swizzle agnostic
__device__ __forceinline__ void split_region(uint8_t* base, int bytes, int lo_off, int tid) {
  for (int off = tid * 16; off < bytes; off += 128 * 16) {
    const float4 v = *reinterpret_cast<const float4*>(base + off);
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(base + off) = h;
    *reinterpret_cast<float4*>(base + off + lo_off) = l;
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? GEMM_THREADS_SPLIT : GEMM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using L = SmemLayout<BLOCK_N, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * L::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* ready_bar = empty_bar + MAX_STAGES;     // 3xTF32: split pass done (128 arrivals) -> MMA may read the stage
  uint64_t* tmem_full_bar = ready_bar + MAX_STAGES;
  uint64_t* acc_full_bar = tmem_full_bar + 1;       // [2] 3xTF32: chunk accumulated in TMEM buffer b -> drain warps
  uint64_t* acc_empty_bar = acc_full_bar + 2;       // [2] buffer b drained (128 arrivals) -> MMA may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + 2);

  pdl_launch();  // let the next kernel begin its own prologue right away
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  auto stamp = [&](int slot) { if (dbg_on) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[slot] = t; } };
  if (threadIdx.x == 0) stamp(0);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N;
  const int m0 = blockIdx.y * BLOCK_M;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
  // The tensor core's fp32 accumulation is not round-to-nearest: its error is biased and grows linearly with the number of MMAs folded into
  // one TMEM accumulator (measured: ~2e-8 relative per MMA, 2.9e-5 at K = 4096).  The 3xTF32 mode therefore accumulates at most
  // ACC_CHUNK_KB K blocks (48 MMAs) per TMEM buffer, alternates between two buffers and lets the epilogue warps add every finished chunk
  // into fp32 registers with IEEE adds while the tensor core works on the next one.
  constexpr uint32_t ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr uint32_t TMEM_COLS = SPLIT ? 2 * ACC_COLS : ACC_COLS;
  const int nchunks = SPLIT ? (kb_end - kb_begin + ACC_CHUNK_KB - 1) / ACC_CHUNK_KB : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
        mbar_init(&ready_bar[s], 128);
      }
      mbar_init(tmem_full_bar, 1);
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full_bar[b], 1); mbar_init(&acc_empty_bar[b], 128); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(1);
  pdl_wait();    // everything above overlapped the previous kernel's tail; from here on we touch its outputs
  if (threadIdx.x == 0) stamp(2);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sA = smem + stage * L::STAGE_BYTES;
        uint8_t* sB = sA + A_TILE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], p.tx_bytes);
        const int k0 = kb * BLOCK_K;
        if constexpr (!A_MN) {
          tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 32; ++j)
            if (j < p.a_boxes) tma_load_2d(sA + j * MN_BOX_BYTES, &tmA, &full_bar[stage], m0 + 32 * j, k0);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 32; ++j)
            if (j < p.b_boxes) tma_load_2d(sB + j * MN_BOX_BYTES, &tmB, &full_bar[stage], n0 + 32 * j, k0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (one thread) ================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(/*TF32*/ 2u, A_MN ? 1u : 0u, B_MN ? 1u : 0u, BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        uint32_t tmem_d = tmem_base;
        bool chunk_first = (kb == kb_begin), chunk_last = (kb == kb_end - 1);
        int ci = 0;
        if constexpr (SPLIT) {
          ci = (kb - kb_begin) / ACC_CHUNK_KB;
          const int within = (kb - kb_begin) - ci * ACC_CHUNK_KB;
          chunk_first = within == 0;
          chunk_last = within == ACC_CHUNK_KB - 1 || kb == kb_end - 1;
          tmem_d = tmem_base + static_cast<uint32_t>(ci & 1) * ACC_COLS;
          if (chunk_first && ci >= 2) {      // the buffer still holds chunk ci-2: wait until the drain warps have taken it
            mbar_wait(&acc_empty_bar[ci & 1], static_cast<uint32_t>(((ci >> 1) - 1) & 1));
            tc_fence_after();
          }
        }
        mbar_wait(SPLIT ? &ready_bar[stage] : &full_bar[stage], phase);
        if (kb == kb_begin) stamp(3);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t sB = sA + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major : SWIZZLE_128B, rows 128 B apart, 8-row groups 1024 B apart (SBO); advance 32 B per K step inside the atom.
          // MN-major: tf32 operands must use SWIZZLE_128B_BASE32B ("128B swizzle, 32B atomicity", Swizzle<2,5,2>): 32-wide MN atoms
          //           LBO apart (one TMA box each), K in groups of 4 rows = 512 B (SBO); 8 K-rows = 1024 B per MMA K step.
          const uint32_t oa = A_MN ? k * 1024 : k * UMMA_K * 4, ob = B_MN ? k * 1024 : k * UMMA_K * 4;
          const uint64_t da = A_MN ? umma_desc(sA + oa, MN_BOX_BYTES, 512, 1) : umma_desc(sA + oa, 16, 1024, 2);
          const uint64_t db = B_MN ? umma_desc(sB + ob, MN_BOX_BYTES, 512, 1) : umma_desc(sB + ob, 16, 1024, 2);
          const uint32_t acc = (!chunk_first || k > 0) ? 1u : 0u;
          if constexpr (SPLIT) {
            // small terms first: a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, all into the same fp32 TMEM accumulator
            const uint64_t da_lo = A_MN ? umma_desc(sA + L::OPS_BYTES + oa, MN_BOX_BYTES, 512, 1) : umma_desc(sA + L::OPS_BYTES + oa, 16, 1024, 2);
            const uint64_t db_lo = B_MN ? umma_desc(sB + L::OPS_BYTES + ob, MN_BOX_BYTES, 512, 1) : umma_desc(sB + L::OPS_BYTES + ob, 16, 1024, 2);
            umma_tf32(tmem_d, da_lo, db, idesc, acc);
            umma_tf32(tmem_d, da, db_lo, idesc, 1u);
            umma_tf32(tmem_d, da, db, idesc, 1u);
          } else {
            umma_tf32(tmem_d, da, db, idesc, acc);
          }
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (SPLIT && chunk_last && ci < nchunks - 1) umma_commit(&acc_full_bar[ci & 1]);   // chunk complete -> drain (the last chunk goes to the epilogue)
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full_bar);  // accumulator complete -> epilogue
      stamp(4);
    }
  } else if (SPLIT && warp >= 6) {
    // ================================ 3xTF32 split pass (warps 6-9): every landed stage -> hi (in place) + lo copy ================================
    const int et = threadIdx.x - 192;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      uint8_t* base = smem + stage * L::STAGE_BYTES;
      split_region(base, p.a_bytes, L::OPS_BYTES, et);
      split_region(base + A_TILE_BYTES, p.b_bytes, L::OPS_BYTES, et);
      fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
      mbar_arrive(&ready_bar[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ================================ epilogue (4 warps = 128 TMEM lanes) ================================
    constexpr int CH = BLOCK_N < 32 ? BLOCK_N : 32;   // columns per TMEM load
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    const bool warp_has_rows = (m0 + q * 32) < p.M;   // warp-uniform
    // stage the bias slice of this tile in smem while the main loop runs (one coalesced load instead of per-element L2 round trips)
    float* s_bias = reinterpret_cast<float*>(smem + p.stages * L::STAGE_BYTES + 256);
    {
      const int et = threadIdx.x - 64;
      for (int j = et; j < BLOCK_N; j += 128) s_bias[j] = (p.bias != nullptr && n0 + j < p.N && blockIdx.z == 0) ? __ldg(p.bias + n0 + j) : 0.f;
      named_bar_sync(1, 128);
    }
    // 3xTF32: running fp32 sums of the drained accumulator chunks (IEEE adds in registers)
    float sum[SPLIT ? BLOCK_N : 1];
#pragma unroll
    for (int j = 0; j < (SPLIT ? BLOCK_N : 1); ++j) sum[j] = 0.f;
    if constexpr (SPLIT) {
      for (int ci = 0; ci < nchunks - 1; ++ci) {
        const int b = ci & 1;
        mbar_wait(&acc_full_bar[b], static_cast<uint32_t>((ci >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
          uint32_t r[32];
          const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(b) * ACC_COLS + static_cast<uint32_t>(c0);
          if constexpr (CH == 32) tmem_ld_32x32(ta, r); else tmem_ld_32x16(ta, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) sum[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty_bar[b]);
      }
    }
    mbar_wait(tmem_full_bar, 0);
    if (warp == 2 && lane == 0) stamp(5);
    tc_fence_after();
    const bool have_acc = kb_end > kb_begin;
    const uint32_t fin = SPLIT ? static_cast<uint32_t>((nchunks - 1) & 1) * ACC_COLS : 0u;   // TMEM buffer of the last chunk
    if (warp_has_rows) {
#pragma unroll
      for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        if constexpr (CH == 32) tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + fin + static_cast<uint32_t>(c0), r);
        else tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + fin + static_cast<uint32_t>(c0), r);
        // issue the (optional) ReLU-mask row loads before waiting on TMEM so that both latencies overlap
        float mk[CH];
        const bool full = n0 + c0 + CH <= p.N;
        if (p.mask != nullptr) {
          const float* mrow = p.mask + (long long)m * p.ldmask + n0 + c0;
          if (row_ok && full && ((p.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(mrow + j)); mk[j] = t.x; mk[j + 1] = t.y; mk[j + 2] = t.z; mk[j + 3] = t.w; }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) mk[j] = (row_ok && n0 + c0 + j < p.N) ? __ldg(mrow + j) : 0.f;
          }
        }
        tmem_ld_wait();
        float v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          float x = have_acc ? (__uint_as_float(r[j]) + (SPLIT ? sum[c0 + j] : 0.f)) * p.alpha : 0.f;
          x += s_bias[c0 + j];
          if (p.relu) x = fmaxf(x, 0.f);
          if (p.mask != nullptr) x = mk[j] > 0.f ? x : 0.f;
          v[j] = (row_ok && n0 + c0 + j < p.N) ? x : 0.f;
        }
        if (p.colsum != nullptr) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float s = warp_sum(v[j]);
            if (lane == 0 && n0 + c0 + j < p.N) atomicAdd(p.colsum + n0 + c0 + j, s);
          }
        }
        if (p.store_mode == 2) {
          // fused 2x2/2 max-pool of an NCHW feature map (ReLU already applied): row m = (image, oh, ow); the three partners of a window's
          // top-left element sit 1, W and W+1 lanes further in the same warp (host guarantees 32 % (2W) == 0 and hw % 32 == 0).
          // First maximum wins, like maxpool2x2_fwd_kernel; only pooled values + arg-max leave the SM.
          const int W = p.pool_w, pix = m % p.hw, oh = pix / W, ow = pix - oh * W, img = m / p.hw;
          const bool base = row_ok && ((oh & 1) == 0) && ((ow & 1) == 0);
          const int PW = W >> 1, PHW = p.hw >> 2;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float a01 = __shfl_down_sync(0xffffffffu, v[j], 1), a10 = __shfl_down_sync(0xffffffffu, v[j], W),
                        a11 = __shfl_down_sync(0xffffffffu, v[j], W + 1);
            const int n = n0 + c0 + j;
            if (base && n < p.N) {
              float best = v[j]; int bi = 0;
              if (a01 > best) { best = a01; bi = 1; }
              if (a10 > best) { best = a10; bi = 2; }
              if (a11 > best) { best = a11; bi = 3; }
              const long long o = ((long long)img * p.N + n) * PHW + (oh >> 1) * PW + (ow >> 1);
              p.D[o] = best;
              p.pool_idx[o] = (unsigned char)bi;
            }
          }
        } else
        if (row_ok) {
          if (p.store_mode == 0) {
            float* dst = p.D + (long long)m * p.ldd + n0 + c0;
            const bool vec = ((p.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && full && !p.accumulate;
            if (vec) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                if (n0 + c0 + j < p.N) {
                  if (p.accumulate) atomicAdd(dst + j, v[j]);
                  else dst[j] = v[j];
                }
              }
            }
          } else {
            // NCHW scatter: row m = (image, pixel), column n = channel
            const int img = m / p.hw, pix = m - img * p.hw;
            float* dst = p.D + ((long long)img * p.N) * p.hw + pix;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              const int n = n0 + c0 + j;
              if (n < p.N) {
                if (p.accumulate) atomicAdd(dst + (long long)n * p.hw, v[j]);
                else dst[(long long)n * p.hw] = v[j];
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    if (warp == 2 && lane == 0) stamp(6);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// inner = contiguous dimension (elements), outer = strided dimension, ld = elements between outer rows
static int make_tmap(CUtensorMap* tm, const float* base, long long inner, long long outer, long long ld, int box_inner, int box_outer, bool mn_major) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return -2;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  // driver entry points need a current context on the calling thread (autograd worker threads may never have made a runtime call)
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(1000 + (int)r);
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool SPLIT>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, dim3 grid, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, SPLIT>;
  static int env_stages = -1;
  if (env_stages < 0) { const char* e = getenv("GEOMX_GEMM_STAGES"); env_stages = e ? atoi(e) : 0; }
  int stages = L::STAGES;
  if (p.kb_per_split < stages) stages = p.kb_per_split;   // never more ring slots than K blocks
  if (!SPLIT && BLOCK_N == 128 && stages > 3) {
    // Throughput regime (at least two tiles per SM): a 3-deep ring is 98 KB, so TWO CTAs share an SM and one tile's epilogue (TMEM -> registers
    // -> 64 KB of stores) overlaps the other tile's MMA main loop.  Measured at 8192 x 4096 x 4096: 473 -> 574 TFLOP/s (profiles/gemm_anchor.txt).
    static int sms = 0;
    if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    if ((long long)grid.x * grid.y * grid.z >= 2LL * sms) stages = 3;
  }
  if (env_stages > 0 && env_stages < stages) stages = env_stages;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const int smem_bytes = stages * L::STAGE_BYTES + 256 + 512 + 1024;
  static bool attr_set = false;
  auto kern = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN, SPLIT>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_pdl(kern, dim3(grid), dim3(SPLIT ? GEMM_THREADS_SPLIT : GEMM_THREADS), smem_bytes, stream, ta, tb, p);
  return (int)cudaGetLastError();
}

template <int BLOCK_N, bool SPLIT>
static int dispatch_major2(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BLOCK_N, false, false, SPLIT>(ta, tb, p, grid, s);
  if (!a_mn && b_mn) return launch<BLOCK_N, false, true, SPLIT>(ta, tb, p, grid, s);
  if (a_mn && !b_mn) return launch<BLOCK_N, true, false, SPLIT>(ta, tb, p, grid, s);
  return launch<BLOCK_N, true, true, SPLIT>(ta, tb, p, grid, s);
}
template <int BLOCK_N>
static int dispatch_major(bool split, bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid, cudaStream_t s) {
  return split ? dispatch_major2<BLOCK_N, true>(a_mn, b_mn, ta, tb, p, grid, s) : dispatch_major2<BLOCK_N, false>(a_mn, b_mn, ta, tb, p, grid, s);
}

}  // namespace gx

static unsigned long long* g_gemm_dbg = nullptr;
// 3 = 3xTF32 (fp32-accurate, default), 1 = TF32.  Process-wide; GEOMX_GEMM_PRECISION=tf32|3xtf32 sets the initial value.
static int g_gemm_prec = -1;
static int gemm_precision() {
  if (g_gemm_prec < 0) { const char* e = getenv("GEOMX_GEMM_PRECISION"); g_gemm_prec = (e && (e[0] == 't' || e[0] == '1')) ? 1 : 3; }
  return g_gemm_prec;
}
GX_API int gx_gemm_set_precision(int prec) { g_gemm_prec = (prec == 1) ? 1 : 3; return g_gemm_prec; }
GX_API int gx_gemm_get_precision() { return gemm_precision(); }
GX_API int gx_gemm_set_debug(unsigned long long* p) { g_gemm_dbg = p; return 0; }

// A: K-major -> [M][K] with row stride lda; MN-major -> [K][M] with row stride lda.  Same for B with N.
// returns 0 on success, -1 if the operands do not satisfy TMA alignment (caller falls back to gx_gemm_simt).
static int gemm_tf32_impl(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, int M, int N, int K, float* D,
                          long long ldd, const float* bias, const float* mask, long long ldmask, float* colsum, int relu, int accumulate,
                          int store_mode, int hw, float alpha, int split_k,
                          int pool_w, unsigned char* pool_idx, cudaStream_t stream) {
  using namespace gx;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (store_mode == 2) {   // fused pooling: whole windows must live inside one warp's 32 rows, no split-K / accumulate
    if (pool_w < 2 || (pool_w & 1) || (32 % (2 * pool_w)) != 0 || hw % 32 != 0 || (hw % pool_w) != 0 || ((hw / pool_w) & 1) || M % 32 != 0 ||
        split_k > 1 || accumulate || pool_idx == nullptr) return -1;
  }
  if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return -1;
  // tile width: these problems are bound by how fast ONE SM can pull its operand panel out of L2 (~100 GB/s per SM, measured with
  // tools/gemm_phases.py), so small problems are cut into narrow tiles to spread the B panel over many SMs; large ones use 128-wide tiles.
  int block_n = 16;
  const long long mt = ceil_div(M, BLOCK_M);
  for (int bn : {128, 64, 32, 16}) {
    if (mt * ceil_div(N, bn) >= 24 || bn == 16) { block_n = bn; break; }
  }
  if (mt * ceil_div(N, 128) >= 148) block_n = 128;
  if (b_mn && block_n < 32) block_n = 32;   // an MN-major B tile is made of 32-wide TMA boxes
  CUtensorMap ta, tb;
  int rc;
  // TMA cost is per requested row (measured: ~1.5 ns per box row, in or out of bounds), so boxes are trimmed to the rows that exist when the
  // whole problem fits one tile in that dimension; the untouched smem rows feed accumulator lanes that the epilogue never stores.
  const int a_rows = (mt == 1) ? (int)((M + 7) / 8 * 8) : BLOCK_M;
  const int nt = (int)ceil_div(N, block_n);
  const int b_rows = (nt == 1) ? (int)((N + 7) / 8 * 8 < block_n ? (N + 7) / 8 * 8 : block_n) : block_n;
  const int a_boxes = (mt == 1) ? (int)ceil_div(M, 32) : BLOCK_M / 32;
  const int b_boxes = (nt == 1) ? (int)ceil_div(N, 32) : block_n / 32;
  if (!a_mn) rc = make_tmap(&ta, A, K, M, lda, BLOCK_K, a_rows, false);
  else rc = make_tmap(&ta, A, M, K, lda, 32, BLOCK_K, true);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap(&tb, B, K, N, ldb, BLOCK_K, b_rows, false);
  else rc = make_tmap(&tb, B, N, K, ldb, 32, BLOCK_K, true);
  if (rc) return rc;
  const int num_kb = (int)ceil_div(K, BLOCK_K);
  if (split_k < 1) split_k = 1;
  if (split_k > num_kb) split_k = num_kb;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.kb_per_split = (int)ceil_div(num_kb, split_k);
  split_k = (int)ceil_div(num_kb, p.kb_per_split);
  p.D = D; p.ldd = ldd; p.bias = bias; p.mask = mask; p.ldmask = ldmask; p.colsum = colsum;
  p.relu = relu; p.accumulate = (accumulate || split_k > 1) ? 1 : 0; p.store_mode = store_mode; p.hw = hw > 0 ? hw : 1; p.alpha = alpha;
  p.stages = 0; p.dbg = g_gemm_dbg;
  p.pool_w = pool_w; p.pool_idx = pool_idx;
  p.a_boxes = a_boxes < 1 ? 1 : a_boxes; p.b_boxes = b_boxes < 1 ? 1 : b_boxes;
  p.a_bytes = a_mn ? p.a_boxes * MN_BOX_BYTES : a_rows * BLOCK_K * 4;
  p.b_bytes = b_mn ? p.b_boxes * MN_BOX_BYTES : b_rows * BLOCK_K * 4;
  p.tx_bytes = p.a_bytes + p.b_bytes;
  const bool split = gemm_precision() == 3;
  dim3 grid((unsigned)ceil_div(N, block_n), (unsigned)mt, (unsigned)split_k);
  switch (block_n) {
    case 16: return dispatch_major<16>(split, a_mn, b_mn, ta, tb, p, grid, stream);
    case 32: return dispatch_major<32>(split, a_mn, b_mn, ta, tb, p, grid, stream);
    case 64: return dispatch_major<64>(split, a_mn, b_mn, ta, tb, p, grid, stream);
    default: return dispatch_major<128>(split, a_mn, b_mn, ta, tb, p, grid, stream);
  }
}

GX_API int gx_gemm_tf32(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, int M, int N, int K, float* D,
                        long long ldd, const float* bias, const float* mask, long long ldmask, float* colsum, int relu, int accumulate,
                        int store_mode, int hw, float alpha, int split_k, cudaStream_t stream) {
  return gemm_tf32_impl(A, lda, a_mn, B, ldb, b_mn, M, N, K, D, ldd, bias, mask, ldmask, colsum, relu, accumulate, store_mode == 2 ? 1 : store_mode, hw,
                        alpha, split_k, 0, nullptr, stream);
}
// Convolution-as-GEMM with the 2x2/2 max-pool fused into the epilogue: rows of A are (image, oh, ow) with ow fastest and `hw` = OH*OW,
// `pool_w` = OW.  D receives the POOLED NCHW map [images][N][OH/2][OW/2], `pool_idx` the arg-max position of every pooled element.
// returns -1 when the geometry does not fit the in-warp pooling (caller runs GEMM + pool kernels instead).
GX_API int gx_gemm_tf32_pool(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, int M, int N, int K, float* D,
                             const float* bias, int relu, int hw, int pool_w, unsigned char* pool_idx, float alpha, cudaStream_t stream) {
  return gemm_tf32_impl(A, lda, a_mn, B, ldb, b_mn, M, N, K, D, N, bias, nullptr, 0, nullptr, relu, 0, 2, hw, alpha, 1, pool_w,
                        pool_idx, stream);
}

// ------------------------------------------------------------------------------------------------
// CUDA-core fallback for operands that violate TMA alignment (odd leading dimensions, e.g. N=10 classifier heads).
// Same contract as gx_gemm_tf32 (fp32 FMA).  64x64 tile, 16-deep smem panels.
namespace gx {
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, long long lda, int a_mn, const float* __restrict__ B,
                                                         long long ldb, int b_mn, int M, int N, int K, GemmParams p) {
  gx::pdl_wait();
  gx::pdl_launch();
  __shared__ float sA[16][64 + 1];
  __shared__ float sB[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i / 64, r = i % 64;
      const int k = k0 + kk;
      float a = 0.f, b = 0.f;
      if (k < K && m0 + r < M) a = a_mn ? A[(long long)k * lda + m0 + r] : A[(long long)(m0 + r) * lda + k];
      if (k < K && n0 + r < N) b = b_mn ? B[(long long)k * ldb + n0 + r] : B[(long long)(n0 + r) * ldb + k];
      sA[kk][r] = a;
      sB[kk][r] = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (m >= M || n >= N) continue;
      float x = acc[i][j] * p.alpha;
      if (p.bias) x += p.bias[n];
      if (p.relu) x = fmaxf(x, 0.f);
      if (p.mask) x = p.mask[(long long)m * p.ldmask + n] > 0.f ? x : 0.f;
      if (p.colsum) atomicAdd(p.colsum + n, x);
      float* dst = p.store_mode == 0 ? p.D + (long long)m * p.ldd + n
                                     : p.D + ((long long)(m / p.hw) * N + n) * p.hw + (m % p.hw);
      if (p.accumulate) atomicAdd(dst, x);
      else *dst = x;
    }
  }
}
}  // namespace gx

GX_API int gx_gemm_simt(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, int M, int N, int K, float* D,
                        long long ldd, const float* bias, const float* mask, long long ldmask, float* colsum, int relu, int accumulate,
                        int store_mode, int hw, float alpha, cudaStream_t stream) {
  using namespace gx;
  if (M <= 0 || N <= 0) return 0;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.kb_per_split = 0; p.D = D; p.ldd = ldd; p.bias = bias; p.mask = mask; p.ldmask = ldmask; p.colsum = colsum;
  p.relu = relu; p.accumulate = accumulate; p.store_mode = store_mode; p.hw = hw > 0 ? hw : 1; p.alpha = alpha; p.stages = 0; p.dbg = nullptr; p.a_boxes = p.b_boxes = 0; p.tx_bytes = 0; p.a_bytes = p.b_bytes = 0;
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(M, 64));
  launch_pdl(gemm_simt_kernel, dim3(grid), dim3(256), 0, stream, A, lda, a_mn, B, ldb, b_mn, M, N, K, p);
  return (int)cudaGetLastError();
}
