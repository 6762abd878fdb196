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
//   setmaxnreg moves registers from warps 12-15 (24 each) to the softmax warps (152 each).
// TMEM (all 512 columns): per sub-tile S (64 columns, fp32), P (32 columns, bf16 pairs) and the output accumulator O (64, fp32).
//   * P has its OWN columns, so S is free again as soon as the softmax warps hold S(j) in registers: Q K^T of tile j+1 is issued
//     right then and runs on the tensor core WHILE the softmax of tile j is computed - the softmax -> tensor core -> softmax
//     round trip is off the critical path (with P aliased onto S it cost ~45 % of the tile time).
//   * P is written with tcgen05.st and consumed by the P V MMA as its TMEM A operand (TS mode): no shared-memory round trip,
//     no generic->async proxy fence.
//   * O accumulates in TMEM across KV tiles; the running row max is adopted lazily (only when some row of the warp grew by more
//     than 2^8; otherwise the stale max is kept and P may exceed 1 - harmless in fp32/bf16), so O is rescaled in TMEM only in
//     the first few tiles and read once at the end.
// The MUFU unit (16 ex2/clk/SM) is the scarcest pipe at d = 64: 3 of every 8 element pairs go through a polynomial exp2 on the
// FMA pipe; scale / sum use packed fp32x2 arithmetic; the row max uses 3-input max in independent chains.
#include <cstdlib>
#include <cstring>
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

  const int warp = threadIdx.x >> 5;
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 12) {
    setmaxnreg_dec<24>();
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
    setmaxnreg_inc<160>();   // pool: 512 x 128 regs at launch = 12 x 32 x 160 + 4 x 32 x 24 (+ 1024 spare)
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
      SoftmaxState sm;
      const long long pc0 = prof_begin();
      long long lap = prof_begin();
      // P(j) is stored per 32-key half (16 packed registers each: the 32 consecutive registers of a whole-tile store were what made the
      // allocator spill) and released at once.  Measured and rejected (round 2, profiles/r02_attention_experiments.md): loading S(j+1)
      // under the exp2 work of tile j (-12 % at head_dim 64, -26 % at 128: the mid-tile wait lands before the tensor core delivers) and
      // deferring the P(j) release to the next tile's top (-10 .. -28 %: P V then sits on the p_free critical path).
      uint32_t s0[32], s1[32];
      for (int j = 0; j < n_tiles; ++j) {
        prof_lap(&lap, 0x67);                          // loop overhead
        mbar_wait(&s_full[x], j & 1, 0x60);
        tc_fence_after();
        tmem_ld32(tSx + 0, s0);
        tmem_ld32(tSx + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);        // Q K^T of the next tile may overwrite S_x now
        prof_lap(&lap, 0x62);                          // TMEM waits, arrives

        SoftmaxTile t;
        softmax_begin(s0, s1, S - j * ATT_BN, sc, pre, j == 0, sm, t);
        uint32_t pk[16];                               // P(j), one half at a time, as packed bf16 pairs
        softmax_half(s0, t, pk);
        if (j > 0) {                                   // P V of tile j-1 (released at the end of that tile) retired: P_x free, O_x quiescent
          mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
          tc_fence_after();
        }
        if (j > 0 && t.rescale) {                      // rare: O_x *= alpha in TMEM 
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
        tmem_st16(tPx, pk);                            // P_x(j) columns [0, 16): keys 0-31
        softmax_half(s1, t, pk);
        tmem_st16(tPx + 16, pk);                       // columns [16, 32): keys 32-63
        softmax_end(sm, t);
        tmem_st_wait();                                // P(j) is in TMEM
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);        // P V (j) may start: a whole tile of slack before its P_x / O_x are needed again
      }
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

// =====================================================================================================================================
// Column-split variant (round 2): TWO softmax warps per query row, each owning 32 of the 64 keys of a tile.
//
// The kernel above keeps every pipe of the SM below 60 % (ncu, profiles/r02_ncu_attention_d64*.md: XU 56 %, issue slots 46 %, FMA 28 %,
// tensor 34 %): three softmax warps per SM sub-partition cannot cover each other's dependent-issue and MUFU-queue stalls.  Without a
// per-tile row maximum (softmax.cuh) the two halves of a row have nothing to tell each other per tile, so a row can be split between two
// warps for free: 16 softmax warps (4 per sub-partition) over two 128-row sub-tiles per CTA, and - the sub-tiles now need only 320 of the
// 512 TMEM columns - the score tile S is DOUBLE-buffered, so Q K^T runs two tiles ahead of the softmax and never waits for it.
//   warps 0-15 : softmax.  warp w: lane quadrant w % 4 (TMEM lanes), sub-tile (w / 4) % 2, key half w / 8.
//   then       : TMA producer, the MMA issuers of sub-tile 0 / 1, one idle warp.
// Agreement between the two halves of a row (the reference of online softmax must be the same in both, they feed one accumulator):
//   * first tile: the half maxima are exchanged through shared memory (one named barrier per CTA lifetime) - both halves take the same
//     decision (reference 0 or the row maximum);
//   * growth: each half watches its PARTIAL sum; "mine passed 2^64" is published in shared memory at the end of tile j, read by the partner
//     in tile j+1 (ordered by the p_full -> P V -> p_free chain both halves are on), and BOTH move the reference by exactly 2^64 at the start
//     of tile j+2.  A sum beyond 2^96 (a jump of > 2^32 inside those two tiles) fails loudly (0x6F), as in the row-per-thread kernel.
// =====================================================================================================================================
constexpr int AT2_NSUB = 2;
constexpr int AT2_QB = AT2_NSUB * ATT_BM;               // 256 query rows per CTA
constexpr int AT2_STAGES = 6;
constexpr int AT2_THREADS = 640;
constexpr int AT2_XCH_BYTES = AT2_NSUB * 2 * ATT_BM * 4;         // float exchange (first-tile maxima, final sums): [sub][half][row]
constexpr int AT2_FLAG_BYTES = AT2_NSUB * 2 * 2 * ATT_BM;        // growth flags: [sub][parity][half][row] bytes
constexpr int AT2_SMEM = AT2_NSUB * ATT_TILE_BYTES + 2 * AT2_STAGES * ATT_KV_BYTES + AT2_XCH_BYTES + AT2_FLAG_BYTES + 1024;
constexpr int AT2_TMEM_S = 0;        // S_x[buf] at (x * 2 + buf) * 64
constexpr int AT2_TMEM_P = 256;      // P_x at 256 + x * 32
constexpr int AT2_TMEM_O = 320;      // O_x at 320 + x * 64
constexpr float AT2_SHIFT = 64.0f;   // the reference moves by exactly 2^64

// One warp per 32 query rows, all 64 keys of a tile (kHalves == 1 of attention_split_kernel): the row-per-thread softmax of attention_kernel
// on the double-buffered S of the split layout.
__device__ __forceinline__ void attention_rowpair_softmax(const AttnParams& p, uint32_t tmem_base, int wq, int x, int lane, int q0, int head,
                                                          int b, int S, int n_tiles, uint64_t* s_full, uint64_t* s_free, uint64_t* p_full,
                                                          uint64_t* p_free, uint64_t* o_full) {
  const int r = wq * 32 + lane;
  const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
  const uint32_t tSx = tmem_base + lane_off + 0 /*AT2_TMEM_S*/ + x * 2 * ATT_BN;
  const uint32_t tPx = tmem_base + lane_off + 256 /*AT2_TMEM_P*/ + x * (ATT_BN / 2);
  const uint32_t tOx = tmem_base + lane_off + 320 /*AT2_TMEM_O*/ + x * ATT_D;
  const float sc = p.scale_log2;
  const bool pre = p.k_prescaled != 0;
  SoftmaxState sm;
  uint32_t s0[32], s1[32];
  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    mbar_wait(&s_full[x * 2 + buf], (j >> 1) & 1, 0x60);
    tc_fence_after();
    tmem_ld32(tSx + buf * ATT_BN, s0);
    tmem_ld32(tSx + buf * ATT_BN + 32, s1);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_free[x * 2 + buf]);
    SoftmaxTile t;
    softmax_begin(s0, s1, S - j * ATT_BN, sc, pre, j == 0, sm, t);
    uint32_t pk[16];
    softmax_half(s0, t, pk);
    if (j > 0) {
      mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
      tc_fence_after();
    }
    if (j > 0 && t.rescale) {
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
    softmax_half(s1, t, pk);
    tmem_st16(tPx + 16, pk);
    softmax_end(sm, t);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&p_full[x]);
  }
  mbar_wait(&o_full[x], 0, 0x69);
  tc_fence_after();
  const int q = q0 + x * ATT_BM + r;
  if (q < S) softmax_final_check(sm.l_run);
  const float inv = 1.0f / sm.l_run;
  bf16* dst = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(q) * p.inner_dim + head * ATT_D;
#pragma unroll 1
  for (int c = 0; c < ATT_D; c += 32) {
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

template <int kHalves>      // 2: two warps per row (16 softmax warps); 1: one warp per row (8 softmax warps), same double-buffered S
__global__ void __launch_bounds__((8 * kHalves + 4) * 32, 1)
attention_split_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                            // [2 sub-tiles][128][64]
  uint8_t* sK = sQ + AT2_NSUB * ATT_TILE_BYTES;                  // [stages][64][64]
  uint8_t* sV = sK + AT2_STAGES * ATT_KV_BYTES;
  float* xch = reinterpret_cast<float*>(sV + AT2_STAGES * ATT_KV_BYTES);      // [sub][half][row]
  uint8_t* flags = reinterpret_cast<uint8_t*>(xch) + AT2_XCH_BYTES;           // [sub][parity][half][row]
  uint64_t* bars = reinterpret_cast<uint64_t*>(flags + AT2_FLAG_BYTES);
  uint64_t* q_full = bars;                          // 1
  uint64_t* k_full = bars + 1;                      // [ST]
  uint64_t* k_empty = k_full + AT2_STAGES;
  uint64_t* v_full = k_empty + AT2_STAGES;
  uint64_t* v_empty = v_full + AT2_STAGES;
  uint64_t* s_full = v_empty + AT2_STAGES;          // [sub][buf]
  uint64_t* s_free = s_full + 2 * AT2_NSUB;         // [sub][buf], 8 arrivals
  uint64_t* p_full = s_free + 2 * AT2_NSUB;         // [sub], 8 arrivals
  uint64_t* p_free = p_full + AT2_NSUB;             // [sub]
  uint64_t* o_full = p_free + AT2_NSUB;             // [sub]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + AT2_NSUB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) mbar_timeout(0xA11);
  const int q0 = blockIdx.x * AT2_QB, head = blockIdx.y, b = blockIdx.z;
  const int S = p.seq_len;
  const int n_tiles = (S + ATT_BN - 1) / ATT_BN;
  const int n_sub = min(AT2_NSUB, (S - q0 + ATT_BM - 1) / ATT_BM);

  constexpr int kTma = 8 * kHalves;                 // warp roles after the softmax warps: TMA producer, two MMA issuers, one idle
  if (warp == kTma && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT2_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], n_sub);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], n_sub);
    }
    for (int i = 0; i < 2 * AT2_NSUB; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4 * kHalves); }
    for (int i = 0; i < AT2_NSUB; ++i) { mbar_init(&p_full[i], 4 * kHalves); mbar_init(&p_free[i], 1); mbar_init(&o_full[i], 1); }
    fence_barrier_init();
  }
  if (warp == kTma + 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kTma) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int cq = head * ATT_D, ck = p.inner_dim + head * ATT_D, cv = 2 * p.inner_dim + head * ATT_D;
      mbar_arrive_expect_tx(q_full, n_sub * ATT_TILE_BYTES);
      for (int x = 0; x < n_sub; ++x) tma_load_3d(sQ + x * ATT_TILE_BYTES, &p.tmQKV, q_full, cq, q0 + x * ATT_BM, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % AT2_STAGES;
        const uint32_t ph = (j / AT2_STAGES) & 1;
        mbar_wait_relaxed(&k_empty[st], ph ^ 1, 0x40);
        mbar_arrive_expect_tx(&k_full[st], ATT_KV_BYTES);
        tma_load_3d(sK + st * ATT_KV_BYTES, &p.tmKV, &k_full[st], ck, j * ATT_BN, b);
        mbar_wait_relaxed(&v_empty[st], ph ^ 1, 0x41);
        mbar_arrive_expect_tx(&v_full[st], ATT_KV_BYTES);
        tma_load_3d(sV + st * ATT_KV_BYTES, &p.tmKV, &v_full[st], cv, j * ATT_BN, b);
      }
    }
  } else if (warp == kTma + 1 || warp == kTma + 2) {
    // ===================== MMA issuers: warp kTma + 1 + x -> sub-tile x =====================
    const int x = warp - (kTma + 1);
    if (x < n_sub) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);   // P (TMEM) x V (MN-major)
      const uint32_t q_addr = smem_u32(sQ) + x * ATT_TILE_BYTES, sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
      const uint32_t tPx = tmem_base + AT2_TMEM_P + x * (ATT_BN / 2), tOx = tmem_base + AT2_TMEM_O + x * ATT_D;
      auto issue_qk = [&](int t) {                                       // S_x[t & 1] = Q K_t^T
        const int st = t % AT2_STAGES, buf = t & 1;
        mbar_wait(&k_full[st], (t / AT2_STAGES) & 1, 0x50);
        tc_fence_after();
        const uint32_t k_addr = sK_addr + st * ATT_KV_BYTES;
        const uint32_t tS = tmem_base + AT2_TMEM_S + (x * 2 + buf) * ATT_BN;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_bf16(tS, desc_kmajor_sw128(q_addr + k * 32), desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[x * 2 + buf]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0, 0x52);
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % AT2_STAGES;
        mbar_wait(&v_full[st], (j / AT2_STAGES) & 1, 0x53);
        mbar_wait(&p_full[x], j & 1, 0x54);                              // both halves of P_x(j) are in TMEM (and any rescale of O_x done)
        tc_fence_after();
        const uint32_t v_addr = sV_addr + st * ATT_KV_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k) {
            const uint64_t db = desc_mnmajor_sw128(v_addr + k * 2048, ATT_KV_BYTES);
            umma_bf16_ts(tOx, tPx + k * 8, db, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[st]);
          umma_commit(&p_free[x]);
          if (j == n_tiles - 1) umma_commit(&o_full[x]);
        }
        __syncwarp();
        if (j + 2 < n_tiles) {                                            // S_x[j & 1] is free once both halves hold S(j) in registers
          mbar_wait(&s_free[x * 2 + (j & 1)], (j >> 1) & 1, 0x51);
          issue_qk(j + 2);
        }
      }
    }
  } else if (warp < 8 * kHalves) {
    // ===================== softmax: 32 keys of 32 query rows per warp and tile =====================
    const int wq = warp & 3, x = (warp >> 2) & 1, h = warp >> 3;
    if (kHalves == 1) {
      if (x < n_sub) attention_rowpair_softmax(p, tmem_base, wq, x, lane, q0, head, b, S, n_tiles, s_full, s_free, p_full, p_free, o_full);
    } else if (x < n_sub) {
      const int r = wq * 32 + lane;                                       // query row in the sub-tile == TMEM lane
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tSx = tmem_base + lane_off + AT2_TMEM_S + x * 2 * ATT_BN + 32 * h;
      const uint32_t tPx = tmem_base + lane_off + AT2_TMEM_P + x * (ATT_BN / 2) + 16 * h;
      const uint32_t tOx = tmem_base + lane_off + AT2_TMEM_O + x * ATT_D + 32 * h;      // this half's 32 output columns
      float* xmine = xch + (x * 2 + h) * ATT_BM + r;
      const float* xpeer = xch + (x * 2 + (1 - h)) * ATT_BM + r;
      const bool pre = p.k_prescaled != 0;
      const float sce = pre ? 1.0f : p.scale_log2;
      const uint64_t sc2 = pack_f32x2(sce, sce);
      float m_run = 0.f, l_half = 0.f;
      bool zero_ref = false, g_prev = false, pending = false;
      for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        mbar_wait(&s_full[x * 2 + buf], (j >> 1) & 1, 0x60);
        tc_fence_after();
        uint32_t s[32];
        tmem_ld32(tSx + buf * ATT_BN, s);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x * 2 + buf]);
        const int kv_valid = S - j * ATT_BN - 32 * h;                     // keys of this half that exist
        if (kv_valid < 32) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c >= kv_valid) s[c] = 0xFF800000u;
        }
        // ---- reference policy
        float alpha = 1.0f;
        bool rescale = false;
        if (j == 0) {
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 32; c += 2) mx[(c >> 1) & 3] = fmax3(mx[(c >> 1) & 3], __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
          *xmine = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          asm volatile("bar.sync %0, 256;" ::"r"(1 + x) : "memory");       // the 8 warps of this sub-tile
          const float mt = fmaxf(*xmine, *xpeer) * sce;                   // the row maximum, identical in both halves
          zero_ref = __all_sync(0xffffffffu, fabsf(mt) <= ATT_REF_ZERO_BAND);
          m_run = zero_ref ? 0.0f : mt;
        } else if (__any_sync(0xffffffffu, pending)) {                    // decided two tiles ago by either half of some row of this warp
          if (pending) { alpha = 5.421010862427522e-20f /* 2^-64 */; m_run += AT2_SHIFT; l_half *= alpha; }
          zero_ref = false;
          rescale = true;
          pending = false;
        }
        const bool fast = pre && zero_ref;
        bool poly = ATT_POLY_NUM > 0;
        if (poly) {
          float am[2] = {0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (att_poly_slot(c)) am[c & 1] = fmax3(am[c & 1], fabsf(__uint_as_float(s[2 * c])), fabsf(__uint_as_float(s[2 * c + 1])));
          const float a = fmaxf(am[0], am[1]);
          poly = __all_sync(0xffffffffu, (fast ? a : fmaf(a, sce, fabsf(m_run))) <= 126.0f);
        }
        // ---- exp2, partial sums, bf16 pack
        uint32_t pk[16];
        uint64_t sums2[2] = {0ull, 0ull};
        const uint64_t mneg2 = pack_f32x2(-m_run, -m_run);
        if (fast) {
          if (poly) softmax_exp32<true, true, true>(s, sc2, mneg2, sums2, pk);
          else softmax_exp32<true, false, true>(s, sc2, mneg2, sums2, pk);
        } else {
          if (poly) softmax_exp32<false, true, true>(s, sc2, mneg2, sums2, pk);
          else softmax_exp32<false, false, true>(s, sc2, mneg2, sums2, pk);
        }
        {
          float sa, sb, sc_, sd;
          unpack_f32x2(sums2[0], sa, sb);
          unpack_f32x2(sums2[1], sc_, sd);
          l_half += (sa + sb) + (sc_ + sd);
        }
        if (!(l_half < ATT_FAIL_AT)) mbar_timeout(0x6F);
        // ---- P V of tile j-1 retired: P_x free, O_x quiescent; the partner's growth flag of tile j-1 is visible
        if (j > 0) {
          mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
          tc_fence_after();
          pending = g_prev || flags[((x * 2 + ((j - 1) & 1)) * 2 + (1 - h)) * ATT_BM + r] != 0;
        }
        // my partial sum, as it will stand after a shift already scheduled for the next tile, passed 2^64 (or is not finite)
        const bool g = !(l_half * (pending ? 5.421010862427522e-20f : 1.0f) <= ATT_SHIFT_AT);
        if (rescale) {                                                    // rare: this half's 32 columns of O_x *= alpha
          uint32_t o0[32];
          tmem_ld32(tOx, o0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
          tmem_st32(tOx, o0);
        }
        tmem_st16(tPx, pk);                                               // P_x(j), keys [32 h, 32 h + 32): 16 columns
        flags[((x * 2 + buf) * 2 + h) * ATT_BM + r] = g ? 1 : 0;
        g_prev = g;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);
      }
      // ---- final: row sum = both halves' partial sums (same reference); this half normalises and stores 32 of the 64 output columns
      mbar_wait(&o_full[x], 0, 0x69);
      tc_fence_after();
      *xmine = l_half;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + x) : "memory");
      const float l_row = l_half + *xpeer;
      uint32_t o0[32];
      tmem_ld32(tOx, o0);
      tmem_ld_wait();
      const int q = q0 + x * ATT_BM + r;
      if (q < S) {
        if (!(l_row < ATT_FAIL_AT) || !(l_row > 0.f)) mbar_timeout(0x6F);
        const float inv = 1.0f / l_row;
        bf16* dst = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(q) * p.inner_dim + head * ATT_D + 32 * h;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(o0[c * 8 + 0]) * inv, __uint_as_float(o0[c * 8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(o0[c * 8 + 2]) * inv, __uint_as_float(o0[c * 8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(o0[c * 8 + 4]) * inv, __uint_as_float(o0[c * 8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(o0[c * 8 + 6]) * inv, __uint_as_float(o0[c * 8 + 7]) * inv);
          reinterpret_cast<uint4*>(dst)[c] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTma + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// FFB200_ATT_VARIANT selects the head_dim-64 kernel (A/B measurements; read once): "row3" = three sub-tiles, one warp per 32 rows,
// single-buffered S (attention_kernel); "row2" = two sub-tiles, one warp per 32 rows, double-buffered S; "split2" = two sub-tiles, two warps
// per row, double-buffered S.
static int attention_variant() {
  static const int v = [] {
    const char* e = getenv("FFB200_ATT_VARIANT");
    if (e == nullptr) return 0;
    if (strcmp(e, "row2") == 0) return 1;
    if (strcmp(e, "split2") == 0) return 2;
    return 0;
  }();
  return v;
}

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] {
      cudaError_t e1 = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
      if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attention_split_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM);
      if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(attention_split_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM);
      return e1;
    });
    if (e != cudaSuccess) return e;
  }
  const int v = attention_variant();
  if (v == 0) {
    dim3 grid((p.seq_len + ATT_QB - 1) / ATT_QB, p.num_heads, p.batch);
    attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(p);
  } else {
    dim3 grid((p.seq_len + AT2_QB - 1) / AT2_QB, p.num_heads, p.batch);
    if (v == 1) attention_split_kernel<1><<<grid, 12 * 32, AT2_SMEM, stream>>>(p);
    else attention_split_kernel<2><<<grid, 20 * 32, AT2_SMEM, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace ffb
