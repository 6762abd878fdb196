// EXPERIMENT - compiled only with -DFFB_ATT_SUMMMA (see ffb200.cu); the product kernel is ../attention.cu.  NOT yet run on a GPU.
//
// Head-dim-64 flash attention whose ROW SUM comes from the tensor core: every V stage is followed in shared memory by a panel of
// bf16 ones, the P V MMA runs with N = 80 (V's 64 columns + 16 columns of ones through the MN-major leading-dimension offset), so
// accumulator column 64 of O_x is  l = sum_j P_ij  - accumulated in fp32 from the SAME bf16-rounded P that multiplies V - and the
// softmax warps drop their 32 packed adds per tile (~16 % of the FMA-pipe work the kernel is bound by, profiles/r01_attention_whatif.md).
// TMEM: O_x needs 80 columns, so 3 x (S 64 + P 32 + O 80) = 528 does not fit; P_x is aliased onto the first 32 columns of S_x
// (3 x (64 + 80) = 432) and the tensor-core order per sub-tile becomes QK(j) ; PV(j) ; QK(j+1): with three sub-tiles per SM
// sub-partition the other two sub-tiles' softmax work covers the round trip.  -DFFB_ATT_SUMMMA_NOWAIT additionally issues QK(j+1)
// right behind PV(j) without waiting for its completion (relies on in-order execution of one thread's MMAs).
// Differences from ../attention.cu are marked SUMMMA.
#include "../common.cuh"
#include "../kernels.h"
#include "../softmax.cuh"
#include "tmem_narrow.cuh"

// Combined with -DFFB_ATT_MAXFREE (softmax.cuh) the reference shift needs the running sum as its magnitude signal: the softmax warps
// then read accumulator column 64 back once per tile, right after the wait that guarantees P V (j-1) has retired, and use it - two tiles
// late - for the decision of tile j+1 (threshold 2^24 against an fp32 / bf16 range of 2^127: the lag is harmless, an overflow still traps).

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
constexpr int ATT_V_STAGE = 2 * ATT_KV_BYTES;     // SUMMMA: a V tile followed by its 8 KB panel of ones
constexpr int ATT_ON = ATT_D + 16;                // SUMMMA: accumulator width (64 output columns + the row sum, N must be a multiple of 16)
constexpr int ATT_SMEM = ATT_NSUB * ATT_TILE_BYTES /*Q*/ + ATT_STAGES * ATT_KV_BYTES /*K*/ + ATT_STAGES * ATT_V_STAGE /*V|1*/ + 1024;   // 193 KB
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_TMEM_S = 0;       // S_x at columns x*64
constexpr int ATT_TMEM_O = 192;     // SUMMMA: O_x at columns 192 + x*80 (64 output columns, then 16 equal copies of the row sum); P_x aliases S_x

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];    // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* sQ = smem;                                   // [3 sub-tiles][128][64]
  uint8_t* sK = sQ + ATT_NSUB * ATT_TILE_BYTES;         // [stages][64][64]
  uint8_t* sV = sK + ATT_STAGES * ATT_KV_BYTES;         // [stages][64][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_STAGES * ATT_V_STAGE);
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
  // SUMMMA: the panels of ones behind every V stage (all elements equal, so the swizzled layout does not matter); written through the
  // generic proxy, read by the tensor core through the async proxy -> proxy fence before the barrier
  for (int i = threadIdx.x; i < ATT_STAGES * (ATT_KV_BYTES / 16); i += ATT_THREADS) {
    const int st = i / (ATT_KV_BYTES / 16), o = i % (ATT_KV_BYTES / 16);
    st_shared_v4(smem_u32(sV) + st * ATT_V_STAGE + ATT_KV_BYTES + o * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  }
  fence_proxy_async_smem();
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
          tma_load_3d(sV + st * ATT_V_STAGE, &p.tmKV, &v_full[st], cv, j * ATT_BN, b);
        }
      }
    } else if (warp - 13 < n_sub) {
      // ===================== MMA issuers: warp 13 + x -> sub-tile x =====================
      // The whole warp walks the loop (warp-uniform state -> uniform registers feed UTCHMMA), one elected lane issues.
      // Tensor-core order per sub-tile:  QK(0) ; { QK(j+1) ; PV(j) } for j = 0..  -  QK(j+1) starts when the softmax warps have
      // S(j) in registers and runs during their exp2 work; PV(j) starts when they have written P(j).
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_ON, 0, 1);  // SUMMMA: P (TMEM) x [V | 1] (MN-major, N = 80)
      const int x = warp - 13;
      const uint32_t q_addr = smem_u32(sQ) + x * ATT_TILE_BYTES, sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
      const uint32_t tSx = tmem_base + ATT_TMEM_S + x * ATT_BN, tPx = tSx /* SUMMMA: aliased */,
                     tOx = tmem_base + ATT_TMEM_O + x * ATT_ON;
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
#ifdef FFB_ATT_STAGGER
      // EXPERIMENT (tools/gpu_maxfree.sh): start sub-tile x a fraction of a tile late, to test whether the three identical softmax
      // groups of an SM sub-partition run in lockstep (synchronised stalls) - costs x * FFB_ATT_STAGGER cycles once per CTA.
      if (x > 0) { const long long t0 = clock64(); while (clock64() - t0 < static_cast<long long>(x) * FFB_ATT_STAGGER) {} }
#endif
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % ATT_STAGES;
        mbar_wait(&v_full[st], (j / ATT_STAGES) & 1, 0x53);
        mbar_wait(&p_full[x], j & 1, 0x54);      // P_x(j) is in TMEM (and any rescale of O_x done)
        tc_fence_after();
        const uint32_t v_addr = sV_addr + st * ATT_V_STAGE;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k) {
            // A = P from TMEM: 16 bf16 of K per step = 8 columns.  B = [V | 1] (MN-major): 16 kv rows = 2048 B per step; the second
            // 64-wide MN atom (of which N = 80 uses 16 columns) is the panel of ones ATT_KV_BYTES behind the V tile (SUMMMA).
            const uint64_t db = desc_mnmajor_sw128(v_addr + k * 2048, ATT_KV_BYTES);
            umma_bf16_ts(tOx, tPx + k * 8, db, idesc_o, (j | k) != 0 ? 1u : 0u);   // O_x (and the row sum) accumulate across KV tiles
          }
          umma_commit(&v_empty[st]);
          umma_commit(&p_free[x]);
          if (j == n_tiles - 1) umma_commit(&o_full[x]);
        }
        __syncwarp();
        if (j + 1 < n_tiles) {                   // SUMMMA: S_x(j+1) overwrites the columns P_x(j) lives in
#ifndef FFB_ATT_SUMMMA_NOWAIT
          mbar_wait(&p_free[x], j & 1, 0x51);    // P V of tile j has retired
          tc_fence_after();
#endif
          issue_qk(j + 1);
        }
      }
    }
  } else {
    // ===================== softmax: warps 4x .. 4x+3 -> sub-tile x =====================
    setmaxnreg_inc<152>();   // pool: 512 x 128 regs at launch = 12 x 32 x 152 + 4 x 32 x 24 (+ 4096 spare)
    const int x = warp >> 2;                          // sub-tile
    if (x < n_sub) {
      const int wq = warp & 3;                        // TMEM lane quadrant
      const int r = wq * 32 + lane;                   // query row in the sub-tile == TMEM lane
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tSx = tmem_base + lane_off + ATT_TMEM_S + x * ATT_BN;
      const uint32_t tPx = tSx;                         // SUMMMA: P_x aliases the first 32 columns of S_x
      const uint32_t tOx = tmem_base + lane_off + ATT_TMEM_O + x * ATT_ON;
      const float sc = p.scale_log2;
      float m_run = -INFINITY, l_run = 0.f;
      const long long pc0 = prof_begin();
      long long lap = prof_begin();
      for (int j = 0; j < n_tiles; ++j) {
        prof_lap(&lap, 0x67);                          // loop overhead / previous arrive
        mbar_wait(&s_full[x], j & 1, 0x60);
        tc_fence_after();
        prof_lap(&lap, 0x68);                          // wait s_full
        uint32_t s0[32], s1[32];
        tmem_ld32(tSx + 0, s0);
        tmem_ld32(tSx + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);        // Q K^T of the next tile may overwrite S_x now
        prof_lap(&lap, 0x62);                          // TMEM load of S

        uint32_t pk[32];                               // P(j) as packed bf16 pairs
        float alpha;
        const bool rescale = softmax_block64<false>(s0, s1, S - j * ATT_BN, sc, m_run, l_run, pk, alpha);   // SUMMMA: no row sum here
        prof_lap(&lap, 0x64);                          // max + exp2 + sum + pack

        if (j > 0) {                                   // P V of tile j-1 retired (issued a whole softmax ago): P_x free, O_x quiescent
          mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
          tc_fence_after();
#ifdef FFB_ATT_MAXFREE
          if (!rescale) {                              // running sum through tile j-1 (same reference as this tile: no shift pending)
            uint32_t lr;
            tmem_ld1(tOx + 64, lr);
            tmem_ld_wait();
            l_run = __uint_as_float(lr);
          }
#endif
        }
        if (j > 0 && rescale) {                          // rare: O_x *= alpha in TMEM
          uint32_t o0[32], o1[32];
          tmem_ld32(tOx, o0);
          tmem_ld32(tOx + 32, o1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
            o1[i] = __float_as_uint(__uint_as_float(o1[i]) * alpha);
          }
          tmem_st32(tOx, o0);
          tmem_st32(tOx + 32, o1);
          uint32_t o2[16];                               // SUMMMA: the row-sum columns scale with O
          tmem_ld16(tOx + 64, o2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o2[i] = __float_as_uint(__uint_as_float(o2[i]) * alpha);
          tmem_st16(tOx + 64, o2);
#ifdef FFB_ATT_MAXFREE
          l_run = __uint_as_float(o2[0]);              // the rescaled running sum through tile j-1
#endif
        }
        prof_lap(&lap, 0x65);                          // wait p_free, rare O rescale
        tmem_st32(tPx, pk);                            // P_x(j): 64 bf16 per row = 32 columns
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);
        prof_lap(&lap, 0x66);                          // P -> TMEM, arrive
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
        uint32_t o2[16];                                 // SUMMMA: l = accumulator column 64
        tmem_ld16(tOx + 64, o2);
        tmem_ld_wait();
        l_run = __uint_as_float(o2[0]);
      }
      prof_end(pc0, 0x70 + warp);
      // O_x and l_run are both relative to the final running max m_run.
      const int q = q0 + x * ATT_BM + r;
      if (q < S) {
        softmax_final_check(l_run);
        const float inv = 1.0f / l_run;
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
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int qb = ATT_QB;
  dim3 grid((p.seq_len + qb - 1) / qb, p.num_heads, p.batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ffb
