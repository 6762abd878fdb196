// EXPERIMENT - compiled only with -DFFB_ATT_SPLIT (see ffb200.cu); the product kernel is ../attention_d128.cu.  NOT yet run on a GPU.
//
// Head-dim-128 flash attention with the score tile's COLUMNS split between two softmax warps.  TMEM allows only two 128-row sub-tiles
// at head_dim 128 (S 64 + P 32 + O 128 columns each), i.e. two softmax warps per SM sub-partition - the latency-bound end of the
// T(n) ~ 940 + 400 n table in profiles/r01_attention_whatif.md.  The two other experiments remove every per-row reduction from the
// softmax warps - no row maximum after tile 0 (max-free reference, softmax.cuh) and no row sum (the tensor core produces it as
// accumulator column 128, attention_d128_summma.cu) - so nothing forces one thread to own a whole row any more: warps w and w + 4
// address the same TMEM lane quadrant (lanes 32 (w % 4) ..) and take columns 0-31 / 32-63 of S, write their half of P, and never talk
// to each other after tile 0 (one named-barrier exchange of the exact maximum).  A reference shift is decided from the SAME value -
// the row sum read back from TMEM - by the same arithmetic in both halves, so it needs no handshake either; each half rescales its own
// 64 output columns.  16 softmax warps = four per sub-partition at unchanged TMEM (2 x (64 + 32 + 144) = 480 columns).
//   warps 0-15 : softmax; warp w -> sub-tile w >> 3, column half (w >> 2) & 1, lane quadrant w & 3
//   warp 16    : TMA producer      warps 17 / 18 : MMA issuers (as in the product kernel)      warp 19 : idle
#include "../common.cuh"
#include "../kernels.h"
#include "../softmax.cuh"
#include "tmem_narrow.cuh"

namespace ffb {

constexpr int A128_BM = 128;
constexpr int A128_NSUB = 2;
constexpr int A128_QB = A128_NSUB * A128_BM;
constexpr int A128_BN = 64;
constexpr int A128_D = 128;
constexpr int A128_STAGES = 4;
constexpr int A128_THREADS = 640;
constexpr int A128_QPANEL = 128 * 64 * 2;           // 16 KB: one 64-column panel of a Q sub-tile
constexpr int A128_KVPANEL = A128_BN * 64 * 2;      //  8 KB: one 64-column panel of a K or V tile
constexpr int A128_VSTAGE = 3 * A128_KVPANEL;       // two V panels + the panel of ones
constexpr int A128_ON = A128_D + 16;                // accumulator width (128 output columns + the row sum)
constexpr int A128_TILES = A128_NSUB * 2 * A128_QPANEL + A128_STAGES * 2 * A128_KVPANEL + A128_STAGES * A128_VSTAGE;   // 224 KB
constexpr int A128_XCH = 2 * 4 * 2 * 32 * 4;        // tile-0 maximum exchange: [sub-tile][quadrant][half][32 rows] floats
constexpr int A128_SMEM = A128_TILES + A128_XCH + 512;   // + barriers; 226.5 KB of the 227 KB a CTA may use
constexpr int A128_TMEM_COLS = 512;
constexpr int A128_TMEM_S = 0;        // S_x at columns x*64
constexpr int A128_TMEM_P = 128;      // P_x at columns 128 + x*32
constexpr int A128_TMEM_O = 192;      // O_x at columns 192 + x*144 (128 output columns, then 16 copies of the row sum)

// exp2 of 32 scores (one half of a tile row) against the reference m: x = s * sc - m * sc, P as 16 packed bf16 pairs; no max, no sum
__device__ __forceinline__ void exp_block32(const uint32_t (&s)[32], float sc, float m_run, uint32_t (&pk)[16]) {
  const uint64_t sc2 = pack_f32x2(sc, sc), mneg2 = pack_f32x2(-m_run * sc, -m_run * sc);
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const uint64_t x2 = ffma2(pack_f32x2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), sc2, mneg2);
    float e0, e1;
    if ((c % ATT_POLY_PERIOD) < ATT_POLY_NUM) {
      exp2_poly_pair(x2, e0, e1);
    } else {
      float t0, t1;
      unpack_f32x2(x2, t0, t1);
      e0 = ex2_approx(t0); e1 = ex2_approx(t1);
    }
    pk[c] = pack_bf16x2(e0, e1);
  }
}

template <bool CROSS>
__global__ void __launch_bounds__(A128_THREADS, 1)
attention_d128_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                          // [2 sub-tiles][2 panels][128][64]
  uint8_t* sK = sQ + A128_NSUB * 2 * A128_QPANEL;              // [stages][2 panels][64][64]
  uint8_t* sV = sK + A128_STAGES * 2 * A128_KVPANEL;           // [stages][V panel 0 | V panel 1 | ones]
  float* xch = reinterpret_cast<float*>(sV + A128_STAGES * A128_VSTAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xch) + A128_XCH);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + A128_STAGES;
  uint64_t* v_full = k_empty + A128_STAGES;
  uint64_t* v_empty = v_full + A128_STAGES;
  uint64_t* s_full = v_empty + A128_STAGES;    // [2]
  uint64_t* s_free = s_full + A128_NSUB;       // [2]  all EIGHT softmax warps of the sub-tile hold their half of S_x(j)
  uint64_t* p_full = s_free + A128_NSUB;       // [2]  all eight have written their half of P_x(j)
  uint64_t* p_free = p_full + A128_NSUB;       // [2]
  uint64_t* o_full = p_free + A128_NSUB;       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + A128_NSUB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) mbar_timeout(0xA13);
  const int q0 = blockIdx.x * A128_QB;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int S = p.seq_len;
  const int Skv = CROSS ? p.kv_len : S;
  const int n_tiles = (Skv + A128_BN - 1) / A128_BN;
  const int n_sub = min(A128_NSUB, (S - q0 + A128_BM - 1) / A128_BM);

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < A128_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], n_sub);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], n_sub);
    }
    for (int i = 0; i < A128_NSUB; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_full[i], 8);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_ptr_smem, A128_TMEM_COLS);
  // the panel of ones behind the two V panels of every stage (generic-proxy stores, read by the tensor core: proxy fence)
  for (int i = threadIdx.x; i < A128_STAGES * (A128_KVPANEL / 16); i += A128_THREADS) {
    const int st = i / (A128_KVPANEL / 16), o = i % (A128_KVPANEL / 16);
    st_shared_v4(smem_u32(sV) + st * A128_VSTAGE + 2 * A128_KVPANEL + o * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 16) {
    setmaxnreg_dec<24>();
    if (warp == 16) {
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
          for (int h = 0; h < 2; ++h) tma_load_3d(sV + st * A128_VSTAGE + h * A128_KVPANEL, &p.tmKV, &v_full[st], cv + h * 64, j * A128_BN, b);
        }
      }
    } else if (warp < 19 && warp - 17 < n_sub) {
      // ===================== MMA issuers: warp 17 + x -> sub-tile x (unchanged protocol) =====================
      constexpr uint32_t idesc_s = make_idesc_bf16(A128_BM, A128_BN, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(A128_BM, A128_ON, 0, 1);   // P (TMEM) x [V | 1] (MN-major, N = 144 over three panels)
      const int x = warp - 17;
      const uint32_t q_addr = smem_u32(sQ) + x * 2 * A128_QPANEL, sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
      const uint32_t tSx = tmem_base + A128_TMEM_S + x * A128_BN, tPx = tmem_base + A128_TMEM_P + x * (A128_BN / 2),
                     tOx = tmem_base + A128_TMEM_O + x * A128_ON;
      auto issue_qk = [&](int j) {
        const int st = j % A128_STAGES;
        mbar_wait(&k_full[st], (j / A128_STAGES) & 1, 0x50);
        tc_fence_after();
        const uint32_t k_addr = sK_addr + st * 2 * A128_KVPANEL;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < A128_D / 16; ++kk)
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
        const uint32_t v_addr = sV_addr + st * A128_VSTAGE;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < A128_BN / 16; ++kk) {
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
    // ===================== softmax: warp w -> sub-tile w >> 3, column half (w >> 2) & 1, TMEM lane quadrant w & 3 =====================
    setmaxnreg_inc<112>();   // pool: 640 x 96 regs at launch >= 16 x 32 x 112 + 4 x 32 x 24
    const int x = warp >> 3;
    if (x < n_sub) {
      const int hf = (warp >> 2) & 1;
      const int wq = warp & 3;
      const int r = wq * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tSx = tmem_base + lane_off + A128_TMEM_S + x * A128_BN + hf * 32;
      const uint32_t tPx = tmem_base + lane_off + A128_TMEM_P + x * (A128_BN / 2) + hf * 16;
      const uint32_t tOx = tmem_base + lane_off + A128_TMEM_O + x * A128_ON;
      const float sc = p.scale_log2;
      float m_run = 0.f, l_seen = 0.f;
      const int mask_hi = p.kv_mask_lo ? p.kv_mask_hi : 0;
      const int mask_lo = p.kv_mask_lo ? max(__ldg(p.kv_mask_lo + b), 1) : 0;   // key 0 always stays (keeps the tile-0 maximum finite)
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&s_full[x], j & 1, 0x60);
        tc_fence_after();
        uint32_t s[32];
        tmem_ld32(tSx, s);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);
        const int k0 = j * A128_BN + hf * 32;                     // first key of this warp's columns
        if (Skv - k0 < 32 || (mask_lo < mask_hi && k0 < mask_hi && k0 + 32 > mask_lo)) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (k0 + c >= Skv || (k0 + c >= mask_lo && k0 + c < mask_hi)) s[c] = 0xFF800000u;
        }
        float alpha = 1.0f;
        bool rescale = false;
        if (j == 0) {
          // exact row maximum of tile 0, exchanged once between the two halves of the row
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 32; c += 2) mx[(c >> 1) & 3] = fmax3(mx[(c >> 1) & 3], __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
          const float mine = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          float* slot = xch + ((x * 4 + wq) * 2) * 32;
          slot[hf * 32 + lane] = mine;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + x * 4 + wq) : "memory");
          m_run = fmaxf(mine, slot[(hf ^ 1) * 32 + lane]);
        } else {
          const bool grow = !(l_seen <= 16777216.0f);             // also true for inf / NaN
          rescale = __any_sync(0xffffffffu, grow);
          if (rescale) {
            if (!(l_seen < 3.0e38f)) mbar_timeout(0x6E);          // overflow inside one tile: fail loudly
            if (grow) {
              const int e = static_cast<int>((__float_as_uint(l_seen) >> 23) & 0xFF) - 127;
              alpha = __uint_as_float(static_cast<uint32_t>(127 - e) << 23);     // 2^-e, exact
              m_run += __fdividef(static_cast<float>(e), sc);
              l_seen *= alpha;
            }
          }
        }
        uint32_t pk[16];
        exp_block32(s, sc, m_run, pk);
        if (j > 0) {
          mbar_wait(&p_free[x], (j - 1) & 1, 0x61);              // P V of tile j-1 retired: P_x free, O_x quiescent
          tc_fence_after();
          if (!rescale) {                                         // running sum through tile j-1: identical in both halves of the row
            uint32_t lr;
            tmem_ld1(tOx + A128_D, lr);
            tmem_ld_wait();
            l_seen = __uint_as_float(lr);
          } else {                                                // rare: this half's 64 output columns (half 0 also the row sum) *= alpha
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
              uint32_t o0[32];
              tmem_ld32(tOx + hf * 64 + c, o0);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
              tmem_st32(tOx + hf * 64 + c, o0);
            }
            if (hf == 0) {
              uint32_t o2[16];
              tmem_ld16(tOx + A128_D, o2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o2[i] = __float_as_uint(__uint_as_float(o2[i]) * alpha);
              tmem_st16(tOx + A128_D, o2);
            }
          }
        }
        tmem_st16(tPx, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);
      }
      mbar_wait(&o_full[x], 0, 0x69);
      tc_fence_after();
      float l_run;
      {
        uint32_t lr;
        tmem_ld1(tOx + A128_D, lr);
        tmem_ld_wait();
        l_run = __uint_as_float(lr);
      }
      const int q = q0 + x * A128_BM + r;
      if (q < S && !(l_run < 3.0e38f)) mbar_timeout(0x6E);
      const float inv = 1.0f / l_run;
      bf16* dst = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(q) * p.out_row_stride + head * A128_D + hf * 64;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {                          // this half's 64 output columns
        uint32_t o0[32];
        tmem_ld32(tOx + hf * 64 + c, o0);
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
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A128_TMEM_COLS);
  }
}

cudaError_t launch_attention_d128(const AttnParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_d128_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, A128_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.seq_len + A128_QB - 1) / A128_QB, p.num_heads, p.batch);
  attention_d128_kernel<false><<<grid, A128_THREADS, A128_SMEM, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_attention_d128_cross(const AttnParams& p, cudaStream_t stream) {
  if (p.kv_len <= 0 || p.kv_mask_lo != nullptr) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_d128_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, A128_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.seq_len + A128_QB - 1) / A128_QB, p.num_heads, p.batch);
  attention_d128_kernel<true><<<grid, A128_THREADS, A128_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ffb
