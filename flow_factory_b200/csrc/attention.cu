// tcgen05 / TMEM flash attention for the MMDiT joint (image+text) attention, head_dim 64, non-causal.
//
// Replaces F.scaled_dot_product_attention at DF/models/attention_processor.py:1484 together with the
// torch.cat of image/text q,k,v (1480-1482) and the head transposes (1451-1452, 1485): q, k, v are read
// straight out of the fused-QKV GEMM's token-major [B, S, 3D] buffer with 3-D TMA boxes (one per head),
// so there is no concat copy and no [B,H,S,d] transpose in HBM.
//
// One CTA per (384 query rows = three 128-row sub-tiles, head, batch), 16 warps, one CTA per SM:
//   warps 0-3 / 4-7 / 8-11 : softmax of sub-tile 0 / 1 / 2; thread == query row == TMEM lane (no shuffles).  Three softmax warps
//                            per SM sub-partition, each from a different sub-tile, cover each other's latencies.
//   warp  12               : TMA producer (Q once; ring of 64x64 K and V tiles shared by the three sub-tiles)
//   warps 13 / 14 / 15     : MMA issuers, one per sub-tile (S = Q K^T : M128 N64 K64 ; O += P V : M128 N64 K64, V as MN-major
//                            operand).  One issuer per sub-tile keeps the softmax groups out of lockstep.
//   setmaxnreg: warps 12-15 keep 64 registers each, the softmax warps get 144.
// TMEM (all 512 columns): per sub-tile S (64 columns, fp32), P (32 columns, bf16 pairs) and the output accumulator O (64, fp32).
//   * P has its OWN columns, so S is free again as soon as the softmax warps hold S(j) in registers: Q K^T of tile j+1 is issued
//     right then and runs on the tensor core WHILE the softmax of tile j is computed - the softmax -> tensor core -> softmax
//     round trip is off the critical path (with P aliased onto S it cost ~45 % of the tile time).
//   * P is written with tcgen05.st and consumed by the P V MMA as its TMEM A operand (TS mode): no shared-memory round trip,
//     no generic->async proxy fence.
//   * O accumulates in TMEM across KV tiles and is read once at the end; it is rescaled in TMEM only when the softmax reference moves
//     (softmax.cuh: no per-tile row maximum - the reference is 0, or the first tile's maximum, and moves by an exact power of two when
//     the running sum passes 2^64), which for RMS-normed q / k never happens.
// The softmax itself (softmax.cuh) is the cost of this kernel: at d = 64 the tensor core needs 8 cycles per SM for 256 scores, the MUFU
// unit alone 16.  What is left per score on the product path (keys pre-scaled by softmax_scale * log2(e) in the QKV GEMM epilogue,
// reference 0): one MUFU.EX2 or, for 3 of every 8 pairs, a polynomial exp2 on the FMA pipe; half a packed add for the row sum; half a
// bf16 pack - and a steady-state tile whose skeleton is as small as it gets: compile-time first / last tiles, one common-path body, elected-lane
// arrives, warp-uniform addresses, and no range check of the polynomial slots when the RMSNorm weights prove the range (AttnParams::bound_wq).
// Measured history (778 -> 1018 TFLOP/s in round 2), rejected alternatives and the ncu picture: profiles/r02_attention_experiments.md.
#include <type_traits>
#include "common.cuh"
#include "kernels.h"
#include "softmax.cuh"

namespace ffb {

constexpr int ATT_BM = 128;     // query rows per sub-tile
constexpr int ATT_NSUB = 3;     // sub-tiles per CTA
constexpr int ATT_QB = ATT_NSUB * ATT_BM;   // query rows per CTA
constexpr int ATT_BN = 64;      // kv rows per tile
constexpr int ATT_D = 64;
constexpr int ATT_STAGES = 6;
constexpr int ATT_THREADS = 512;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;      // 16 KB: a Q sub-tile (128 rows x 64 bf16)
constexpr int ATT_KV_BYTES = ATT_BN * 64 * 2;     // 8 KB: a K or V tile
constexpr int ATT_SMEM = ATT_NSUB * ATT_TILE_BYTES /*Q*/ + 2 * ATT_STAGES * ATT_KV_BYTES /*K,V*/ + 1024;   // 145 KB
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_TMEM_S = 0;       // S_x at columns x*64
constexpr int ATT_TMEM_P = 192;     // P_x at columns 192 + x*32
constexpr int ATT_TMEM_O = 320;     // O_x at columns 320 + x*64 (last column used: 511)

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];    // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* sQ = smem;                                   // [3 sub-tiles][128][64]
  uint8_t* sK = sQ + ATT_NSUB * ATT_TILE_BYTES;         // [stages][64][64]
  uint8_t* sV = sK + ATT_STAGES * ATT_KV_BYTES;         // [stages][64][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_STAGES * ATT_KV_BYTES);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // [ST]
  uint64_t* k_empty = k_full + ATT_STAGES;       // [ST]
  uint64_t* v_full = k_empty + ATT_STAGES;       // [ST]
  uint64_t* v_empty = v_full + ATT_STAGES;       // [ST]
  uint64_t* s_full = v_empty + ATT_STAGES;       // [3]  S_x(j) = Q K_j^T is in TMEM
  uint64_t* s_free = s_full + ATT_NSUB;          // [3]  the four softmax warps hold S_x(j) in registers: S_x may be overwritten
  uint64_t* p_full = s_free + ATT_NSUB;          // [3]  P_x(j) written to TMEM (and any rescale of O_x done)
  uint64_t* p_free = p_full + ATT_NSUB;          // [3]  P V of tile j retired: P_x may be overwritten, O_x is quiescent
  uint64_t* o_full = p_free + ATT_NSUB;          // [3]  final O_x complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + ATT_NSUB);
  uint32_t* bounded_smem = tmem_ptr_smem + 1;         // 1: |exponent| <= 126 proven for this launch (see AttnParams::bound_wq)

  // warp index through a shuffle: ptxas then knows it is warp-uniform and keeps everything derived from it (TMEM addresses, barrier
  // addresses, role tests) in uniform registers - no R2UR in front of every LDTM / STTM / SYNCS of the softmax loop (round 2, call 19:
  // 950 -> 981 TFLOP/s at head_dim 64, 1290 -> 1360 at head_dim 128; -DFFB_ATT_NO_UWARP for the A/B)
#ifndef FFB_ATT_NO_UWARP
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
#else
  const int warp = threadIdx.x >> 5;
#endif
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) mbar_timeout(0xA11);   // swizzled tiles would be silently misread
  const int q0 = blockIdx.x * ATT_QB;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int S = p.seq_len;
  const int n_tiles = (S + ATT_BN - 1) / ATT_BN;
  const int n_sub = min(ATT_NSUB, (S - q0 + ATT_BM - 1) / ATT_BM);   // sub-tiles holding at least one valid row (>= 1)

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], n_sub);   // one MMA issuer warp per sub-tile releases the slot
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], n_sub);
    }
    for (int i = 0; i < ATT_NSUB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 13) tmem_alloc(tmem_ptr_smem, ATT_TMEM_COLS);
  if (warp == 14) {
    // Range proof for the polynomial exp2 slots (softmax.cuh): with keys pre-scaled by scale_log2 and both q and k RMS-normed per head,
    // |q . k'| <= ||q|| ||k'|| <= (8 max|w_q|) (8 max|w_k| scale_log2); 1.016 covers the bf16 roundings of the epilogue that produced them.
    float wq = 0.f, wk = 0.f;
    bool known = p.k_prescaled != 0 && p.bound_wq[0] != nullptr && p.bound_wk[0] != nullptr;
    if (known) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (p.bound_wq[i]) wq = fmaxf(wq, fmaxf(fabsf(__bfloat162float(p.bound_wq[i][lane])), fabsf(__bfloat162float(p.bound_wq[i][lane + 32]))));
        if (p.bound_wk[i]) wk = fmaxf(wk, fmaxf(fabsf(__bfloat162float(p.bound_wk[i][lane])), fabsf(__bfloat162float(p.bound_wk[i][lane + 32]))));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wq = fmaxf(wq, __shfl_xor_sync(0xffffffffu, wq, o));
        wk = fmaxf(wk, __shfl_xor_sync(0xffffffffu, wk, o));
      }
    }
    const float bound = 64.0f * 1.016f * wq * wk * p.scale_log2;
    if (lane == 0) *bounded_smem = (known && bound <= 120.0f) ? 1u : 0u;     // NaN / inf weights fail the comparison: guard stays
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 12) {
    setmaxnreg_dec<64>();
    if (warp == 12) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        const int cq = head * ATT_D, ck = p.inner_dim + head * ATT_D, cv = 2 * p.inner_dim + head * ATT_D;
        mbar_arrive_expect_tx(q_full, n_sub * ATT_TILE_BYTES);
        for (int x = 0; x < n_sub; ++x) tma_load_3d(sQ + x * ATT_TILE_BYTES, &p.tmQKV, q_full, cq, q0 + x * ATT_BM, b);
        for (int j = 0; j < n_tiles; ++j) {
          const int st = j % ATT_STAGES;
          const uint32_t ph = (j / ATT_STAGES) & 1;
          mbar_wait_relaxed(&k_empty[st], ph ^ 1, 0x40);   // off the critical path: do not steal issue slots
          mbar_arrive_expect_tx(&k_full[st], ATT_KV_BYTES);
          tma_load_3d(sK + st * ATT_KV_BYTES, &p.tmKV, &k_full[st], ck, j * ATT_BN, b);
          mbar_wait_relaxed(&v_empty[st], ph ^ 1, 0x41);
          mbar_arrive_expect_tx(&v_full[st], ATT_KV_BYTES);
          tma_load_3d(sV + st * ATT_KV_BYTES, &p.tmKV, &v_full[st], cv, j * ATT_BN, b);
        }
      }
    } else if (warp - 13 < n_sub) {
      // ===================== MMA issuers: warp 13 + x -> sub-tile x =====================
      // The whole warp walks the loop (warp-uniform state -> uniform registers feed UTCHMMA), one elected lane issues.
      // Tensor-core order per sub-tile:  QK(0) ; { QK(j+1) ; PV(j) } for j = 0..  -  QK(j+1) starts when the softmax warps have
      // S(j) in registers and runs during their exp2 work; PV(j) starts when they have written P(j).
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);   // P (TMEM) x V (MN-major)
      const int x = warp - 13;
      const uint32_t q_addr = smem_u32(sQ) + x * ATT_TILE_BYTES, sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
      const uint32_t tSx = tmem_base + ATT_TMEM_S + x * ATT_BN, tPx = tmem_base + ATT_TMEM_P + x * (ATT_BN / 2),
                     tOx = tmem_base + ATT_TMEM_O + x * ATT_D;
      auto issue_qk = [&](int j) {
        const int st = j % ATT_STAGES;
        mbar_wait(&k_full[st], (j / ATT_STAGES) & 1, 0x50);
        tc_fence_after();
        const uint32_t k_addr = sK_addr + st * ATT_KV_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_bf16(tSx, desc_kmajor_sw128(q_addr + k * 32), desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[x]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0, 0x52);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % ATT_STAGES;
        if (j + 1 < n_tiles) {
          mbar_wait(&s_free[x], j & 1, 0x51);    // S_x(j) is in the softmax warps' registers
          issue_qk(j + 1);
        }
        mbar_wait(&v_full[st], (j / ATT_STAGES) & 1, 0x53);
        mbar_wait(&p_full[x], j & 1, 0x54);      // P_x(j) is in TMEM (and any rescale of O_x done)
        tc_fence_after();
        const uint32_t v_addr = sV_addr + st * ATT_KV_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k) {
            // A = P from TMEM: 16 bf16 of K per step = 8 columns.  B = V (MN-major): 16 kv rows = 2048 B per step.
            const uint64_t db = desc_mnmajor_sw128(v_addr + k * 2048, ATT_KV_BYTES);
            umma_bf16_ts(tOx, tPx + k * 8, db, idesc_o, (j | k) != 0 ? 1u : 0u);   // O_x accumulates across KV tiles
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
    setmaxnreg_inc<144>();   // pool: 512 x 128 regs at launch >= 12 x 32 x 144 + 4 x 32 x 64 (the softmax uses ~106; 24 for the helpers
                             // spilled the MMA issuers' descriptors and the producer's ring state to local memory)
    const int x = warp >> 2;                          // sub-tile
    if (x < n_sub) {
      const int wq = warp & 3;                        // TMEM lane quadrant
      const int r = wq * 32 + lane;                   // query row in the sub-tile == TMEM lane
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tSx = tmem_base + lane_off + ATT_TMEM_S + x * ATT_BN;
      const uint32_t tPx = tmem_base + lane_off + ATT_TMEM_P + x * (ATT_BN / 2);
      const uint32_t tOx = tmem_base + lane_off + ATT_TMEM_O + x * ATT_D;
      const float sc = p.scale_log2;
      const bool pre = p.k_prescaled != 0;              // the scores already are base-2 exponents (softmax.cuh)
      const bool bounded = *bounded_smem != 0;          // warp-uniform: the range guard of the polynomial slots is proven unnecessary
      SoftmaxState sm;
      const long long pc0 = prof_begin();
      long long lap = prof_begin();
      // P(j) is stored per 32-key half (16 packed registers each: the 32 consecutive registers of a whole-tile store were what made the
      // allocator spill) and released at once.  Measured and rejected (round 2, profiles/r02_attention_experiments.md): loading S(j+1)
      // under the exp2 work of tile j (-12 % at head_dim 64, -26 % at 128: the mid-tile wait lands before the tensor core delivers) and
      // deferring the P(j) release to the next tile's top (-10 .. -28 %: P V then sits on the p_free critical path).
      uint32_t s0[32], s1[32];
      // One KV tile.  kFirst / kLast are compile-time so that the steady-state body (tiles 1 .. n-2) carries neither the first-tile maximum
      // nor the key mask of the ragged last tile, and its common case - scores are the exponents (pre-scaled keys, reference 0), the
      // polynomial slots in range, nothing to rescale - is ONE branch body instead of a dispatch per 32-key half.
      auto tile = [&](const int j, auto first_c, auto last_c, auto par_c) {
        constexpr bool kFirst = decltype(first_c)::value, kLast = decltype(last_c)::value;
        constexpr int kPar = decltype(par_c)::value;           // barrier parity of tile j when known at compile time, else -1
        const uint32_t par = kPar >= 0 ? static_cast<uint32_t>(kPar) : static_cast<uint32_t>(j & 1);
        prof_lap(&lap, 0x67);                          // loop overhead
        mbar_wait(&s_full[x], par, 0x60);
        tc_fence_after();
        tmem_ld32(tSx + 0, s0);
        tmem_ld32(tSx + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        ATT_TILE_SYNCWARP();
#ifndef FFB_ATT_NO_ELECT  // elected lane instead of the lane index (which ptxas keeps in local memory across setmaxnreg): +1 % (call 19)
        if (elect_one()) mbar_arrive(&s_free[x]);
#else
        if (lane == 0) mbar_arrive(&s_free[x]);        // Q K^T of the next tile may overwrite S_x now
#endif
        prof_lap(&lap, 0x62);                          // TMEM waits, arrives

        SoftmaxTile t;
        // the range proof covers real keys only: the ragged last tile carries -inf for the keys beyond the sequence, so it keeps the guard
        softmax_begin<PolyD64, PolyD64G>(s0, s1, kLast ? S - j * ATT_BN : ATT_BN, sc, pre, kFirst, sm, t, bounded && !kLast);
        uint32_t pk[16];                               // P(j), one half at a time, as packed bf16 pairs
        auto wait_p_free = [&]() {                     // P V of tile j-1 (released at the end of that tile) retired: P_x free, O_x quiescent
          if (!kFirst) {
            mbar_wait(&p_free[x], par ^ 1u, 0x61);
            tc_fence_after();
          }
        };
        if (!kFirst && !kLast && t.fast && t.poly && !t.rescale) {
          softmax_exp32<true, true, true, PolyD64>(s0, t.sc2, t.mneg2, t.sums2, pk);
          wait_p_free();
          tmem_st16(tPx, pk);                          // P_x(j) columns [0, 16): keys 0-31
          softmax_exp32<true, true, true, PolyD64>(s1, t.sc2, t.mneg2, t.sums2, pk);
          tmem_st16(tPx + 16, pk);                     // columns [16, 32): keys 32-63
        } else if (!kFirst && !kLast && !t.fast && t.poly && !t.rescale && PolyD64G::num > 0) {   // the same for unscaled keys (op-level entry, hooks)
          softmax_exp32<false, true, true, PolyD64G>(s0, t.sc2, t.mneg2, t.sums2, pk);
          wait_p_free();
          tmem_st16(tPx, pk);
          softmax_exp32<false, true, true, PolyD64G>(s1, t.sc2, t.mneg2, t.sums2, pk);
          tmem_st16(tPx + 16, pk);
        } else {
          softmax_half<PolyD64, PolyD64G>(s0, t, pk);
          wait_p_free();
          if (!kFirst && t.rescale) {                  // rare: O_x *= alpha in TMEM
            uint32_t o0[32];
#pragma unroll 1
            for (int c = 0; c < ATT_D; c += 32) {
              tmem_ld32(tOx + c, o0);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * t.alpha);
              tmem_st32(tOx + c, o0);
            }
          }
          tmem_st16(tPx, pk);
          softmax_half<PolyD64, PolyD64G>(s1, t, pk);
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
      using PR_ = std::integral_constant<int, -1>; using P0_ = std::integral_constant<int, 0>; using P1_ = std::integral_constant<int, 1>;
      tile(0, T_{}, T_{}, P0_{});                      // first tile: exact maximum, key mask when it is also the last (S <= 64)
#ifdef FFB_ATT_UNROLL2     // A/B: steady-state tiles two at a time, barrier parities compile-time
      int j = 1;
#pragma unroll 1
      for (; j + 1 < n_tiles - 1; j += 2) { tile(j, F_{}, F_{}, P1_{}); tile(j + 1, F_{}, F_{}, P0_{}); }
      if (j < n_tiles - 1) tile(j, F_{}, F_{}, P1_{});
#else
#pragma unroll 1
      for (int j = 1; j < n_tiles - 1; ++j) tile(j, F_{}, F_{}, PR_{});
#endif
      if (n_tiles > 1) tile(n_tiles - 1, F_{}, T_{}, PR_{});  // ragged last tile: keys beyond the sequence masked
      // final output: O_x / l
      mbar_wait(&o_full[x], 0, 0x69);
      tc_fence_after();
      float o_acc[ATT_D];
      {
        uint32_t o0[32], o1[32];
        tmem_ld32(tOx, o0);
        tmem_ld32(tOx + 32, o1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { o_acc[i] = __uint_as_float(o0[i]); o_acc[32 + i] = __uint_as_float(o1[i]); }
      }
      prof_end(pc0, 0x70 + warp);
      // O_x and sm.l_run are both relative to the final reference.
      const int q = q0 + x * ATT_BM + r;
      if (q < S) {
        softmax_final_check(sm.l_run);
        const float inv = 1.0f / sm.l_run;
        bf16* dst = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(q) * p.inner_dim + head * ATT_D;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 o;
          o.x = pack_bf16x2(o_acc[c * 8 + 0] * inv, o_acc[c * 8 + 1] * inv);
          o.y = pack_bf16x2(o_acc[c * 8 + 2] * inv, o_acc[c * 8 + 3] * inv);
          o.z = pack_bf16x2(o_acc[c * 8 + 4] * inv, o_acc[c * 8 + 5] * inv);
          o.w = pack_bf16x2(o_acc[c * 8 + 6] * inv, o_acc[c * 8 + 7] * inv);
          reinterpret_cast<uint4*>(dst)[c] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] { return cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM); });
    if (e != cudaSuccess) return e;
  }
  dim3 grid((p.seq_len + ATT_QB - 1) / ATT_QB, p.num_heads, p.batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ffb
