// tcgen05 / TMEM flash attention, head_dim 128, non-causal: the FLUX.1 joint [text ; image] attention.
//
// Replaces dispatch_attention_fn(query, key, value) at DF/models/transformers/transformer_flux.py:118-125 together with the
// torch.cat of text/image q,k,v (110-112) and the head (un)flattening (92-94, 126): q, k, v are read straight out of a fused
// token-major [B, S, 3D] buffer (RMSNorm and RoPE already applied to q and k by the QKV GEMM epilogue) with 3-D TMA boxes.
//
// Same design as attention.cu (see there): thread = query row = TMEM lane; P in its own TMEM columns; Q K^T of tile j+1 issued as
// soon as the softmax warps hold S(j) in registers; P V accumulates O in TMEM; no per-tile row maximum (softmax.cuh).  Differences:
//   * two 128-row sub-tiles per CTA (TMEM: per sub-tile S 64 + P 32 + O 128 columns = 448 of 512), 12 warps:
//     warps 0-3 / 4-7 softmax of sub-tile 0 / 1, warp 8 TMA, warps 9 / 10 MMA issuers;
//   * Q / K / V tiles are two 64-column SWIZZLE_128B panels each (a TMA box cannot be wider than the 128-byte swizzle span):
//     Q K^T walks 8 K16 steps over the two panels, P V is one M128 N128 K16 MMA per step whose MN-major V operand spans both
//     panels (leading-dimension byte offset = the panel stride).
#include <type_traits>
#include "common.cuh"
#include "kernels.h"
#include "softmax.cuh"

namespace ffb {

constexpr int A128_BM = 128;
constexpr int A128_NSUB = 2;
constexpr int A128_QB = A128_NSUB * A128_BM;
constexpr int A128_BN = 64;
constexpr int A128_D = 128;
constexpr int A128_STAGES = 4;
constexpr int A128_THREADS = 384;
constexpr int A128_QPANEL = 128 * 64 * 2;           // 16 KB: one 64-column panel of a Q sub-tile
constexpr int A128_KVPANEL = A128_BN * 64 * 2;      //  8 KB: one 64-column panel of a K or V tile
constexpr int A128_SMEM = A128_NSUB * 2 * A128_QPANEL + 2 * A128_STAGES * 2 * A128_KVPANEL + 1024;   // 193 KB
constexpr int A128_TMEM_COLS = 512;
constexpr int A128_TMEM_S = 0;        // S_x at columns x*64
constexpr int A128_TMEM_P = 128;      // P_x at columns 128 + x*32
constexpr int A128_TMEM_O = 256;      // O_x at columns 256 + x*128

// CROSS = false: self / joint attention over one fused [B, S, 3D] buffer (the validated FLUX.1 / Qwen-Image kernel, unchanged).
// CROSS = true : queries and keys / values come from different tensors with different lengths (Wan cross-attention to the text
//                tokens, DF/models/transformers/transformer_wan.py:78-162 with encoder_hidden_states); first GPU run pending.
template <bool CROSS>
__global__ void __launch_bounds__(A128_THREADS, 1)
attention_d128_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                          // [2 sub-tiles][2 panels][128][64]
  uint8_t* sK = sQ + A128_NSUB * 2 * A128_QPANEL;              // [stages][2 panels][64][64]
  uint8_t* sV = sK + A128_STAGES * 2 * A128_KVPANEL;           // [stages][2 panels][64][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A128_STAGES * 2 * A128_KVPANEL);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + A128_STAGES;
  uint64_t* v_full = k_empty + A128_STAGES;
  uint64_t* v_empty = v_full + A128_STAGES;
  uint64_t* s_full = v_empty + A128_STAGES;    // [2]
  uint64_t* s_free = s_full + A128_NSUB;       // [2]
  uint64_t* p_full = s_free + A128_NSUB;       // [2]
  uint64_t* p_free = p_full + A128_NSUB;       // [2]
  uint64_t* o_full = p_free + A128_NSUB;       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + A128_NSUB);

#ifndef FFB_ATT_NO_UWARP  // warp-uniform warp index: see attention.cu
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
#else
  const int warp = threadIdx.x >> 5;
#endif
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) mbar_timeout(0xA12);
  const int q0 = blockIdx.x * A128_QB;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int S = p.seq_len;
  const int Skv = CROSS ? p.kv_len : S;
  const int n_tiles = (Skv + A128_BN - 1) / A128_BN;
  const int n_sub = min(A128_NSUB, (S - q0 + A128_BM - 1) / A128_BM);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < A128_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], n_sub);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], n_sub);
    }
    for (int i = 0; i < A128_NSUB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_ptr_smem, A128_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 8) {
    setmaxnreg_dec<64>();
    if (warp == 8) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        const int cq = (CROSS ? p.q_col : 0) + head * A128_D, ck = (CROSS ? p.k_col : p.inner_dim) + head * A128_D,
                  cv = (CROSS ? p.v_col : 2 * p.inner_dim) + head * A128_D;
        mbar_arrive_expect_tx(q_full, n_sub * 2 * A128_QPANEL);
        for (int x = 0; x < n_sub; ++x)
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sQ + (x * 2 + h) * A128_QPANEL, &p.tmQKV, q_full, cq + h * 64, q0 + x * A128_BM, b);
        for (int j = 0; j < n_tiles; ++j) {
          const int st = j % A128_STAGES;
          const uint32_t ph = (j / A128_STAGES) & 1;
          mbar_wait_relaxed(&k_empty[st], ph ^ 1, 0x40);
          mbar_arrive_expect_tx(&k_full[st], 2 * A128_KVPANEL);
          for (int h = 0; h < 2; ++h) tma_load_3d(sK + (st * 2 + h) * A128_KVPANEL, &p.tmKV, &k_full[st], ck + h * 64, j * A128_BN, b);
          mbar_wait_relaxed(&v_empty[st], ph ^ 1, 0x41);
          mbar_arrive_expect_tx(&v_full[st], 2 * A128_KVPANEL);
          for (int h = 0; h < 2; ++h) tma_load_3d(sV + (st * 2 + h) * A128_KVPANEL, &p.tmKV, &v_full[st], cv + h * 64, j * A128_BN, b);
        }
      }
    } else if (warp - 9 < n_sub) {
      // ===================== MMA issuers: warp 9 + x -> sub-tile x =====================
      constexpr uint32_t idesc_s = make_idesc_bf16(A128_BM, A128_BN, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(A128_BM, A128_D, 0, 1);    // P (TMEM) x V (MN-major, N = 128 over two panels)
      const int x = warp - 9;
      const uint32_t q_addr = smem_u32(sQ) + x * 2 * A128_QPANEL, sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
      const uint32_t tSx = tmem_base + A128_TMEM_S + x * A128_BN, tPx = tmem_base + A128_TMEM_P + x * (A128_BN / 2),
                     tOx = tmem_base + A128_TMEM_O + x * A128_D;
      auto issue_qk = [&](int j) {
        const int st = j % A128_STAGES;
        mbar_wait(&k_full[st], (j / A128_STAGES) & 1, 0x50);
        tc_fence_after();
        const uint32_t k_addr = sK_addr + st * 2 * A128_KVPANEL;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < A128_D / 16; ++kk)   // K16 steps 0-3 in panel 0, 4-7 in panel 1
            umma_bf16(tSx, desc_kmajor_sw128(q_addr + (kk >> 2) * A128_QPANEL + (kk & 3) * 32),
                      desc_kmajor_sw128(k_addr + (kk >> 2) * A128_KVPANEL + (kk & 3) * 32), idesc_s, kk != 0 ? 1u : 0u);
          umma_commit(&s_full[x]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0, 0x52);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % A128_STAGES;
        if (j + 1 < n_tiles) {
          mbar_wait(&s_free[x], j & 1, 0x51);
          issue_qk(j + 1);
        }
        mbar_wait(&v_full[st], (j / A128_STAGES) & 1, 0x53);
        mbar_wait(&p_full[x], j & 1, 0x54);
        tc_fence_after();
        const uint32_t v_addr = sV_addr + st * 2 * A128_KVPANEL;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < A128_BN / 16; ++kk) {
            // A = P from TMEM (8 columns per K16 step).  B = V, MN-major: 16 kv rows = 2048 B per step inside a panel, the two
            // 64-wide panels of the N = 128 extent are A128_KVPANEL apart (leading-dimension byte offset).
            const uint64_t db = desc_mnmajor_sw128(v_addr + kk * 2048, A128_KVPANEL);
            umma_bf16_ts(tOx, tPx + kk * 8, db, idesc_o, (j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[st]);
          umma_commit(&p_free[x]);
          if (j == n_tiles - 1) umma_commit(&o_full[x]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax: warps 4x .. 4x+3 -> sub-tile x =====================
    setmaxnreg_inc<200>();   // pool: 384 x 168 regs at launch >= 8 x 32 x 200 + 4 x 32 x 64
    const int x = warp >> 2;
    if (x < n_sub) {
      const int wq = warp & 3;
      const int r = wq * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tSx = tmem_base + lane_off + A128_TMEM_S + x * A128_BN;
      const uint32_t tPx = tmem_base + lane_off + A128_TMEM_P + x * (A128_BN / 2);
      const uint32_t tOx = tmem_base + lane_off + A128_TMEM_O + x * A128_D;
      const float sc = p.scale_log2;
      const bool pre = p.k_prescaled != 0;              // the scores already are base-2 exponents (softmax.cuh)
      SoftmaxState sm;
      const int mask_hi = p.kv_mask_lo ? p.kv_mask_hi : 0;
      const int mask_lo = p.kv_mask_lo ? max(__ldg(p.kv_mask_lo + b), 1) : 0;   // key 0 always stays (keeps the running max finite)
      // P(j) is stored per 32-key half and released at once (see attention.cu for the measured alternatives).
      uint32_t s0[32], s1[32];
      // One KV tile; kFirst / kLast are compile-time as in attention.cu: the steady-state tiles carry neither the first-tile maximum nor the
      // tail mask, and their common case (scores are the exponents, polynomial slots in range, nothing to rescale) is one branch body.
      // (Round 2 history: peeling only the first tile's S load already made ptxas schedule the loop ~10 % faster, 1228-1252 vs 1115-1156.)
      auto tile = [&](const int j, auto first_c, auto last_c) {
        constexpr bool kFirst = decltype(first_c)::value, kLast = decltype(last_c)::value;
        mbar_wait(&s_full[x], j & 1, 0x60);
        tc_fence_after();
        tmem_ld32(tSx + 0, s0);
        tmem_ld32(tSx + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        ATT_TILE_SYNCWARP();
#ifndef FFB_ATT_NO_ELECT
        if (elect_one()) mbar_arrive(&s_free[x]);
#else
        if (lane == 0) mbar_arrive(&s_free[x]);        // Q K^T of the next tile may overwrite S_x now
#endif
        if (mask_lo < mask_hi) {                        // key-padding mask: only the first few KV tiles overlap the text rows
          const int k0 = j * A128_BN;
          if (k0 < mask_hi && k0 + A128_BN > mask_lo) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if (k0 + c >= mask_lo && k0 + c < mask_hi) s0[c] = 0xFF800000u;
              if (k0 + 32 + c >= mask_lo && k0 + 32 + c < mask_hi) s1[c] = 0xFF800000u;
            }
          }
        }
        SoftmaxTile t;
        softmax_begin<PolyD128, PolyD128G>(s0, s1, kLast ? Skv - j * A128_BN : A128_BN, sc, pre, kFirst, sm, t);
        uint32_t pk[16];                               // P(j), one half at a time, as packed bf16 pairs
        auto wait_p_free = [&]() {                     // P V of tile j-1 (released at the end of that tile) retired: P_x free, O_x quiescent
          if (!kFirst) {
            mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
            tc_fence_after();
          }
        };
        if (!kFirst && !kLast && t.fast && t.poly && !t.rescale) {
          softmax_exp32<true, true, true, PolyD128>(s0, t.sc2, t.mneg2, t.sums2, pk);
          wait_p_free();
          tmem_st16(tPx, pk);                          // P_x(j) columns [0, 16): keys 0-31
          softmax_exp32<true, true, true, PolyD128>(s1, t.sc2, t.mneg2, t.sums2, pk);
          tmem_st16(tPx + 16, pk);                     // columns [16, 32): keys 32-63
        } else if (!kFirst && !kLast && !t.fast && t.poly && !t.rescale && PolyD128G::num > 0) {   // the same for unscaled keys (op-level entry, hooks)
          softmax_exp32<false, true, true, PolyD128G>(s0, t.sc2, t.mneg2, t.sums2, pk);
          wait_p_free();
          tmem_st16(tPx, pk);
          softmax_exp32<false, true, true, PolyD128G>(s1, t.sc2, t.mneg2, t.sums2, pk);
          tmem_st16(tPx + 16, pk);
        } else {
          softmax_half<PolyD128, PolyD128G>(s0, t, pk);
          wait_p_free();
          if (!kFirst && t.rescale) {                  // rare: O_x *= alpha in TMEM (128 columns, 32 at a time)
            uint32_t o0[32];
#pragma unroll 1
            for (int c = 0; c < A128_D; c += 32) {
              tmem_ld32(tOx + c, o0);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * t.alpha);
              tmem_st32(tOx + c, o0);
            }
          }
          tmem_st16(tPx, pk);
          softmax_half<PolyD128, PolyD128G>(s1, t, pk);
          tmem_st16(tPx + 16, pk);
        }
        softmax_end(sm, t);
        tmem_st_wait();                                // P(j) is in TMEM
        tc_fence_before();
        ATT_TILE_SYNCWARP();
#ifndef FFB_ATT_NO_ELECT
        if (elect_one()) mbar_arrive(&p_full[x]);
#else
        if (lane == 0) mbar_arrive(&p_full[x]);        // P V (j) may start: a whole tile of slack before its P_x / O_x are needed again
#endif
      };
      using T_ = std::true_type; using F_ = std::false_type;
      tile(0, T_{}, T_{});                             // first tile: exact maximum, tail mask when it is also the last
#pragma unroll 1
      for (int j = 1; j < n_tiles - 1; ++j) tile(j, F_{}, F_{});
      if (n_tiles > 1) tile(n_tiles - 1, F_{}, T_{});  // ragged last tile: keys beyond the sequence masked
      mbar_wait(&o_full[x], 0, 0x69);
      tc_fence_after();
      const int q = q0 + x * A128_BM + r;
      if (q < S) softmax_final_check(sm.l_run);
      const float inv = 1.0f / sm.l_run;
      bf16* dst = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(q) * p.out_row_stride + head * A128_D;
#pragma unroll 1
      for (int c = 0; c < A128_D; c += 32) {
        uint32_t o0[32];
        tmem_ld32(tOx + c, o0);
        tmem_ld_wait();
        if (q < S) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(o0[g * 8 + 0]) * inv, __uint_as_float(o0[g * 8 + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(o0[g * 8 + 2]) * inv, __uint_as_float(o0[g * 8 + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(o0[g * 8 + 4]) * inv, __uint_as_float(o0[g * 8 + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(o0[g * 8 + 6]) * inv, __uint_as_float(o0[g * 8 + 7]) * inv);
            reinterpret_cast<uint4*>(dst + c)[g] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A128_TMEM_COLS);
  }
}

cudaError_t launch_attention_d128(const AttnParams& p, cudaStream_t stream) {
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] { return cudaFuncSetAttribute(attention_d128_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A128_SMEM); });
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq_len + A128_QB - 1) / A128_QB, p.num_heads, p.batch);
  attention_d128_kernel<false><<<grid, A128_THREADS, A128_SMEM, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_attention_d128_cross(const AttnParams& p, cudaStream_t stream) {
  if (p.kv_len <= 0 || p.kv_mask_lo != nullptr) return cudaErrorInvalidValue;
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] { return cudaFuncSetAttribute(attention_d128_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A128_SMEM); });
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq_len + A128_QB - 1) / A128_QB, p.num_heads, p.batch);
  attention_d128_kernel<true><<<grid, A128_THREADS, A128_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ffb
