// EXPERIMENT - compiled only with -DFFB_ATT_BN128 (see ffb200.cu); the product kernel is ../attention.cu.
// Result (r01, tools/gpu_bn128.sh): passes the attention + engine parity tests, but B=8 S=4429 H=24 takes 1.357 ms against 1.236 ms for
// the product kernel (+10 %): losing the third softmax warp per SM sub-partition costs more than the halved skeleton saves.
//
// Head-dim-64 flash attention with a KV tile of 128 (one Q K^T MMA of N = 128 and one P V of K = 128 per tile) whose two 64-column
// halves are processed back to back by the same softmax warp: the per-tile synchronisation skeleton (s_full wait, s_free / p_full
// arrives, p_free wait, loop) is paid once per 128 columns instead of once per 64.  Price: S 128 + P 64 + O 64 TMEM columns per
// sub-tile, i.e. only TWO 128-row sub-tiles per CTA (two softmax warps per SM sub-partition instead of three).
// profiles/r01_attention_whatif.md (T(n) ~ 940 + 400 n cycles per 64-column tile for n softmax warps per sub-partition) says
// this only wins if halving the skeleton outweighs losing the third warp - it does not.
#include "../common.cuh"
#include "../kernels.h"
#include "../softmax.cuh"

namespace ffb {

constexpr int ATT_BM = 128;
constexpr int ATT_NSUB = 2;
constexpr int ATT_QB = ATT_NSUB * ATT_BM;
constexpr int ATT_BN = 128;     // kv rows per tile (engine.cu sizes the K/V TMA box from this)
constexpr int ATT_D = 64;
constexpr int ATT_STAGES = 4;
constexpr int ATT_THREADS = 384;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;      // 16 KB: a Q sub-tile, a K tile or a V tile
constexpr int ATT_SMEM = ATT_NSUB * ATT_TILE_BYTES + 2 * ATT_STAGES * ATT_TILE_BYTES + 1024;   // 161 KB
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_TMEM_S = 0;       // S_x at columns x*128
constexpr int ATT_TMEM_P = 256;     // P_x at columns 256 + x*64
constexpr int ATT_TMEM_O = 384;     // O_x at columns 384 + x*64

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_NSUB * ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + ATT_STAGES;
  uint64_t* v_full = k_empty + ATT_STAGES;
  uint64_t* v_empty = v_full + ATT_STAGES;
  uint64_t* s_full = v_empty + ATT_STAGES;
  uint64_t* s_free = s_full + ATT_NSUB;
  uint64_t* p_full = s_free + ATT_NSUB;
  uint64_t* p_free = p_full + ATT_NSUB;
  uint64_t* o_full = p_free + ATT_NSUB;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + ATT_NSUB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) mbar_timeout(0xA13);
  const int q0 = blockIdx.x * ATT_QB;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int S = p.seq_len;
  const int n_tiles = (S + ATT_BN - 1) / ATT_BN;
  const int n_sub = min(ATT_NSUB, (S - q0 + ATT_BM - 1) / ATT_BM);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], n_sub);
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
  if (warp == 9) tmem_alloc(tmem_ptr_smem, ATT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 8) {
    setmaxnreg_dec<24>();
    if (warp == 8) {
      if (lane == 0) {   // TMA producer: Q, K and V tiles are all 128-row x 64-column boxes of the token-major qkv buffer
        const int cq = head * ATT_D, ck = p.inner_dim + head * ATT_D, cv = 2 * p.inner_dim + head * ATT_D;
        mbar_arrive_expect_tx(q_full, n_sub * ATT_TILE_BYTES);
        for (int x = 0; x < n_sub; ++x) tma_load_3d(sQ + x * ATT_TILE_BYTES, &p.tmQKV, q_full, cq, q0 + x * ATT_BM, b);
        for (int j = 0; j < n_tiles; ++j) {
          const int st = j % ATT_STAGES;
          const uint32_t ph = (j / ATT_STAGES) & 1;
          mbar_wait_relaxed(&k_empty[st], ph ^ 1, 0x40);
          mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
          tma_load_3d(sK + st * ATT_TILE_BYTES, &p.tmQKV, &k_full[st], ck, j * ATT_BN, b);
          mbar_wait_relaxed(&v_empty[st], ph ^ 1, 0x41);
          mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
          tma_load_3d(sV + st * ATT_TILE_BYTES, &p.tmQKV, &v_full[st], cv, j * ATT_BN, b);
        }
      }
    } else if (warp - 9 < n_sub) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);   // M128 N128
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);    // M128 N64, V MN-major
      const int x = warp - 9;
      const uint32_t q_addr = smem_u32(sQ) + x * ATT_TILE_BYTES, sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
      const uint32_t tSx = tmem_base + ATT_TMEM_S + x * ATT_BN, tPx = tmem_base + ATT_TMEM_P + x * (ATT_BN / 2),
                     tOx = tmem_base + ATT_TMEM_O + x * ATT_D;
      auto issue_qk = [&](int j) {
        const int st = j % ATT_STAGES;
        mbar_wait(&k_full[st], (j / ATT_STAGES) & 1, 0x50);
        tc_fence_after();
        const uint32_t k_addr = sK_addr + st * ATT_TILE_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < ATT_D / 16; ++kk)
            umma_bf16(tSx, desc_kmajor_sw128(q_addr + kk * 32), desc_kmajor_sw128(k_addr + kk * 32), idesc_s, kk != 0 ? 1u : 0u);
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
          mbar_wait(&s_free[x], j & 1, 0x51);
          issue_qk(j + 1);
        }
        mbar_wait(&v_full[st], (j / ATT_STAGES) & 1, 0x53);
        mbar_wait(&p_full[x], j & 1, 0x54);
        tc_fence_after();
        const uint32_t v_addr = sV_addr + st * ATT_TILE_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < ATT_BN / 16; ++kk) {   // 8 K16 steps over the 128 kv rows
            const uint64_t db = desc_mnmajor_sw128(v_addr + kk * 2048, ATT_TILE_BYTES);
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
    setmaxnreg_inc<200>();   // 8 x 32 x 200 + 4 x 32 x 24 <= 384 x 168
    const int x = warp >> 2;
    if (x < n_sub) {
      const int wq = warp & 3;
      const int r = wq * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tSx = tmem_base + lane_off + ATT_TMEM_S + x * ATT_BN;
      const uint32_t tPx = tmem_base + lane_off + ATT_TMEM_P + x * (ATT_BN / 2);
      const uint32_t tOx = tmem_base + lane_off + ATT_TMEM_O + x * ATT_D;
      const float sc = p.scale_log2;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&s_full[x], j & 1, 0x60);
        tc_fence_after();
        // ---- half A (columns 0-63)
        uint32_t s0[32], s1[32], pkA[32], pkB[32];
        tmem_ld32(tSx + 0, s0);
        tmem_ld32(tSx + 32, s1);
        tmem_ld_wait();
        float alphaA, alphaB;
        const bool rescaleA = softmax_block64(s0, s1, S - j * ATT_BN, sc, m_run, l_run, pkA, alphaA);
        // ---- half B (columns 64-127); afterwards the whole S tile is in registers / consumed
        tmem_ld32(tSx + 64, s0);
        tmem_ld32(tSx + 96, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);
        const bool rescaleB = softmax_block64(s0, s1, S - j * ATT_BN - 64, sc, m_run, l_run, pkB, alphaB);
        if (rescaleB) {   // rare: half A's P was computed against the older running max
#pragma unroll
          for (int c = 0; c < 32; ++c) pkA[c] = pack_bf16x2(bf16_lo(pkA[c]) * alphaB, bf16_hi(pkA[c]) * alphaB);
        }
        if (j > 0) {
          mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
          tc_fence_after();
        }
        if (j > 0 && (rescaleA || rescaleB)) {   // rare: O_x *= alphaA * alphaB in TMEM
          const float alpha = alphaA * alphaB;
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
        }
        tmem_st32(tPx, pkA);        // P_x(j): 128 bf16 per row = 64 columns
        tmem_st32(tPx + 32, pkB);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);
      }
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
      const int q = q0 + x * ATT_BM + r;
      if (q < S) {
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
  if (warp == 9) {
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
  dim3 grid((p.seq_len + ATT_QB - 1) / ATT_QB, p.num_heads, p.batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ffb
