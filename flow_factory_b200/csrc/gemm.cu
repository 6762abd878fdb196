// tcgen05 / TMEM / TMA bf16 GEMM for the MMDiT linears:  out = epilogue( A[M,K] * W[N,K]^T )
//
// Replaces every nn.Linear / Conv2d(k=s=patch) call site on the hot path
// (DF/models/attention_processor.py:1443-1445,1461-1463,1495-1498; attention.py:1731; activations.py:87-90;
//  transformers/transformer_sd3.py:293,327; embeddings.py:559) with one persistent warp-specialised kernel:
//   warp 0    : TMA producer   (A tile 128x64, half W tile (BN/2)x64, SWIZZLE_128B, mbarrier ring)
//   warp 1    : MMA issuer     (one elected thread, tcgen05.mma cta_group::2 kind::f16, M=256 N=BN K=16, fp32 accum in TMEM)
//   warp 2    : TMEM allocator
//   warps 4-11: epilogue       (tcgen05.ld 32x32b -> registers; thread == accumulator row; fused bias /
//                               GELU-tanh / adaLN gate + residual / per-head RMSNorm(q,k) / row-table add).
//                               Two warps per TMEM lane quadrant (w and w+4), each taking half of the tile's columns: with one warp
//                               per quadrant the GELU / gate-residual epilogues of a K = 1536 tile took longer than its MMAs
//                               (ncu r01: tensor pipe 85 % / 81 % on MLP-up / attention-out against 97 % on the QKV projection).
//   setmaxnreg moves registers from warps 0-3 (40 each) to the epilogue warps (232 each).
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// CTAs run as pairs (cluster of 2, tcgen05 cta_group::2): one MMA instruction spans both SMs (M = 256 x BN): each CTA
// stages its own 128 A rows and HALF of the W tile, the tensor cores read the other half from the peer's shared memory.
// Per CTA that is 32 KB per k-block instead of 48 KB: deeper TMA pipeline (6 stages) for the same shared memory and
// 128 FLOP/B of L2 -> SM traffic instead of 85.
// The epilogue goes through a per-warp shared-memory staging tile so that every global load/store instruction moves
// whole 128-byte lines (a thread-per-row store touches 32 different lines per instruction); bias / gate slices are read with
// warp-uniform 16-byte __ldg loads (one L1 line serves the whole warp and stays resident over the consecutive tiles of an n-column),
// the q/k RMSNorm weights sit in shared memory once per kernel.
//
// A is addressed as a 3-D tensor [batch][rows_per_batch][K] so that token sub-ranges of a joint
// [B, S, D] buffer (image rows / text rows) are separate GEMM problems with zero-filled ragged tails.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace ffb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;

template <int BN> struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 6 : 8;
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = (BN / 2) * GEMM_BK * 2;   // this CTA's half of the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kStagingBytes = GEMM_EPI_WARPS * 32 * 128;   // per epilogue warp: 32 rows x 64 bf16
  static constexpr int kVecBytes = 2 * 128 * 2;              // norm_q[<=128], norm_k[<=128]: one copy per CTA
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kVecBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget of one CTA per SM");
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5*x*(1+tanh(u)) == x * sigmoid(2u), u = sqrt(2/pi)*(x + 0.044715 x^3)   (F.gelu(approximate='tanh'))
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return __fdividef(x, 1.0f + __expf(-2.0f * u));
}

__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
// Work unit -> (m-pair, n-tile).  Units are rasterised in bands of `band` m-pairs, m fastest inside a band, then n, then the
// next band: a band's A rows (band*256*K*2 B, sized by the host to <= 64 MB) stay L2-resident for the whole sweep over N instead
// of being streamed from HBM once per n-tile (at M = 65536 the activation matrix is 200 MB at K = 1536 and 805 MB at K = 6144,
// far larger than the 126 MB L2).
__device__ __forceinline__ void unit_to_tile(int unit, int pairs_m, int tiles_n, int band, int& mp, int& tn) {
  const int per_band = band * tiles_n;
  const int b = unit / per_band;
  const int rem = unit - b * per_band;
  const int pb = min(band, pairs_m - b * band);   // pairs in this band (the last band may be short)
  tn = rem / pb;
  mp = b * band + rem - tn * pb;
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + Cfg::kStages * Cfg::kABytes;
  const uint32_t smem_stage = smem_base + Cfg::kStages * Cfg::kStageBytes;   // epilogue staging tiles
  const uint32_t smem_vec = smem_stage + Cfg::kStagingBytes;                   // epilogue vectors
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes + Cfg::kVecBytes);
  uint64_t* full_bar = bars;                          // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;          // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;      // [2]
  uint64_t* tmem_empty = bars + 2 * Cfg::kStages + 2; // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

  // warp index through a shuffle: warp-uniform for ptxas (see attention.cu); +2 % on the MLP-up shape in the bench (round 2, call 20)
#ifndef FFB_GEMM_NO_UWARP
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
#else
  const int warp = threadIdx.x >> 5;
#endif
  const int lane = threadIdx.x & 31;
  // column units of a tile the epilogue splits between its two warp groups: 64-column chunks, or whole 128-column heads for the
  // head_dim-128 q/k epilogue; a tile with a single unit is handled by group 0 alone
  const int epi_units = p.epi == EPI_QKV_RMSNORM_ROPE128 ? BN / 128 : BN / 64;
  const int epi_groups = (epi_units >= 2 && p.epi_split == 2) ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    // full: the LEADER's barrier collects the bytes of both CTAs' loads; empty / tmem_full: the leader's commits arrive in
    // both CTAs; tmem_empty: the leader's barrier collects the 4 + 4 epilogue warps of the pair
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8 * epi_groups); }
    fence_barrier_init();
  }
  cluster_sync_all();   // both CTAs resident + barrier inits visible cluster-wide before the pair-wide TMEM allocation
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_m = p.num_batch * p.tiles_m_per_batch;
  const int pairs_m = (tiles_m + 1) >> 1;
  const int tiles_n = p.N / BN;
  const int num_units = pairs_m * tiles_n;      // unit = two vertically adjacent 128 x BN tiles sharing one W tile
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  // register pool: 384 x 168 at launch = 128 x 40 (warps 0-3) + 256 x 232 (epilogue warps)
  if (warp < 4) {
  setmaxnreg_dec<40>();
  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    const long long pc0 = prof_begin();
    int stage = 0; uint32_t phase = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      int mp, tn;
      unit_to_tile(unit, pairs_m, tiles_n, p.band, mp, tn);
      const int tm = 2 * mp + static_cast<int>(cta_rank);
      // a ghost tile (odd tile count) still feeds its half of W to the pair; its A rows are out of bounds -> zeros
      const int b = tm < tiles_m ? tm / p.tiles_m_per_batch : p.num_batch;
      const int row0 = tm < tiles_m ? (tm % p.tiles_m_per_batch) * GEMM_BM : 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 0x10);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          tma_load_3d_2sm(smem + stage * Cfg::kABytes, &p.tmA, &full_bar[stage], kb * GEMM_BK, row0, b);
          tma_load_2d_2sm(smem + Cfg::kStages * Cfg::kABytes + stage * Cfg::kBBytes, &p.tmB, &full_bar[stage], kb * GEMM_BK,
                          tn * BN + static_cast<int>(cta_rank) * (BN / 2));
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
    prof_end(pc0, 0x70);
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; M = 256 across the pair) =====================
    // The whole warp walks the loop so that stage / phase / descriptors stay warp-uniform (uniform registers feed
    // UTCHMMA directly; a lane-0-only loop pays a vector->uniform move per operand per MMA).
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, 0, 0);
      const long long pc0 = prof_begin();
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 0x20);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 0x21);
          tc_fence_after();
          const uint32_t a_addr = smem_a + stage * Cfg::kABytes;
          const uint32_t b_addr = smem_b + stage * Cfg::kBBytes;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              umma_bf16_2sm(d_tmem, desc_kmajor_sw128(a_addr + k * 32), desc_kmajor_sw128(b_addr + k * 32), idesc,
                            (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 0x3);                       // slot free in both CTAs once these MMAs retire
            if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc], 0x3);   // accumulators (both halves) complete
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
      prof_end(pc0, 0x71);
    }
  }
  } else {
    setmaxnreg_inc<232>();
    // ===================== epilogue =====================
    const int ew = warp & 3;                 // TMEM lanes [32*ew, 32*ew+32): a warp reaches the quadrant warp % 4
    const int grp = (warp - 4) >> 2;         // column half of the tile this warp takes
    const uint32_t stg = smem_stage + (warp - 4) * (32 * 128);            // this warp's 32 x 128 B staging tile
    const uint32_t vec_nq = smem_vec, vec_nk = smem_vec + 256;            // norm_q[<=128] | norm_k[<=128], one copy per CTA
    const int coop_row = lane >> 3, coop_c = lane & 7;                    // cooperative (coalesced) access: 8 lanes per 128-B row
    const long long pc0 = prof_begin();
    if (p.epi == EPI_QKV_RMSNORM || p.epi == EPI_QKV_RMSNORM_ROPE128) {
      if (warp == 4) {
        if (p.epi == EPI_QKV_RMSNORM) {
          if (lane < 8) st_shared_v4x(vec_nq + lane * 16, __ldg(reinterpret_cast<const uint4*>(p.norm_q) + lane));
          else if (lane < 16) st_shared_v4x(vec_nk + (lane - 8) * 16, __ldg(reinterpret_cast<const uint4*>(p.norm_k) + (lane - 8)));
        } else {
          if (lane < 16) st_shared_v4x(vec_nq + lane * 16, __ldg(reinterpret_cast<const uint4*>(p.norm_q) + lane));
          else st_shared_v4x(vec_nk + (lane - 16) * 16, __ldg(reinterpret_cast<const uint4*>(p.norm_k) + (lane - 16)));
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GEMM_EPI_WARPS * 32) : "memory");   // the eight epilogue warps only
    }
    const int u_begin = grp * epi_units / epi_groups, u_end = (grp + 1) * epi_units / epi_groups;   // this group's column units
    if (grp < epi_groups) {
    int it = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int mp, tn;
      unit_to_tile(unit, pairs_m, tiles_n, p.band, mp, tn);
      const int tm = 2 * mp + static_cast<int>(cta_rank);
      const bool tile_ok = tm < tiles_m;
      const int b = tile_ok ? tm / p.tiles_m_per_batch : 0;
      const int row_base = (tm % p.tiles_m_per_batch) * GEMM_BM + ew * 32;   // first row (within the batch) of this warp
      bf16* out_base = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(p.out_row_offset + row_base) * p.ldo + tn * BN;
      // bias / gate slices of this tile: warp-uniform 16-byte loads through L1 (the same lines serve the consecutive tiles of an n-column)
      const uint4* bias_v = p.bias ? reinterpret_cast<const uint4*>(p.bias + tn * BN) : nullptr;      // [BN / 8] groups of 8 columns
      const uint4* gate_v = p.epi == EPI_GATE_RESIDUAL
                                ? reinterpret_cast<const uint4*>(p.gate + static_cast<long>(b) * p.gate_batch_stride + tn * BN) : nullptr;
      auto bias8 = [&](int col8) { return bias_v ? __ldg(bias_v + col8) : make_uint4(0, 0, 0, 0); };
      // residual mode: the h chunk (whole-line loads, 8 lanes per 128-B row) is fetched one chunk AHEAD into registers - chunk 0
      // before the accumulator wait, chunk c+1 while chunk c is processed - so its HBM latency never sits on the epilogue's
      // critical path (at K = 1536 the epilogue, not the MMA, paces the attention out-projection).
      uint4 hv[8];
      auto load_residual = [&](int c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + coop_row;
          hv[i] = make_uint4(0, 0, 0, 0);
          if (tile_ok && row_base + rr < p.rows_per_batch)
            hv[i] = *reinterpret_cast<const uint4*>(out_base + static_cast<long>(rr) * p.ldo + c * 64 + coop_c * 8);
        }
      };
      if (p.epi == EPI_GATE_RESIDUAL) load_residual(u_begin);
      mbar_wait(&tmem_full[acc], acc_phase, 0x30);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;

      // bf16 chunk of 64 columns -> own row of the swizzled staging tile -> whole-line stores (8 lanes per 128-B row, 4 rows per instruction)
      auto flush_chunk = [&](int c, const float (&v)[64]) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st_shared_v4(stg + lane * 128 + ((q ^ (lane & 7)) << 4), pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]),
                       pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]), pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]),
                       pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
        __syncwarp();
        // (3) whole-line stores: 8 lanes per 128-B row, 4 rows per instruction
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + coop_row;
          if (tile_ok && row_base + rr < p.rows_per_batch)
            *reinterpret_cast<uint4*>(out_base + static_cast<long>(rr) * p.ldo + c * 64 + coop_c * 8) =
                ld_shared_v4(stg + rr * 128 + ((coop_c ^ (rr & 7)) << 4));
        }
        __syncwarp();
          };
      if (p.epi == EPI_QKV_RMSNORM_ROPE128) {
        // fused q|k|v projection, head_dim 128 (FLUX.1): torch.nn.RMSNorm on the q / k heads (transformer_flux.py:96-97, 104-105:
        // bf16((x * rsqrt(mean x^2 + eps)) * w), fp32 inside) then apply_rotary_emb with interleaved pairs (embeddings.py:1222-1231:
        // bf16(x * cos + [-x1, x0] * sin), fp32 inside).  A head is two 64-column chunks: pass 1 re-reads the accumulator for the
        // sum of squares (TMEM reads are cheap), pass 2 normalises, rotates and stores.  thread == token row.
        const bool row_ok = tile_ok && row_base + lane < p.rows_per_batch;
        const long token = static_cast<long>(p.rope_row_offset + row_base + lane);
        for (int hc = u_begin; hc < u_end; ++hc) {
          const int which = (tn * BN + hc * 128) / p.qk_dim;   // 0 = q, 1 = k, 2 = v
          float rs = 1.0f;
          if (which < 2) {
            float ss = 0.f;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              uint32_t r0[32], r1[32];
              tmem_ld32(t_row + hc * 128 + half * 64, r0);
              tmem_ld32(t_row + hc * 128 + half * 64 + 32, r1);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float bf[8];
                unpack8_bf16(bias8(hc * 16 + half * 8 + q), bf);
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  const int j = q * 8 + e;
                  float a = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]) + bf[e];
                  float b2 = __uint_as_float(j + 1 < 32 ? r0[j + 1] : r1[j + 1 - 32]) + bf[e + 1];
                  bf16_round2(a, b2);
                  ss += a * a + b2 * b2;
                }
              }
            }
            rs = rsqrtf(ss * (1.0f / 128.0f) + p.eps);
          }
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            const int c = hc * 2 + half;
            uint32_t r0[32], r1[32];
            tmem_ld32(t_row + c * 64, r0);
            tmem_ld32(t_row + c * 64 + 32, r1);
            tmem_ld_wait();
            float v[64];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float bf[8];
              unpack8_bf16(bias8(c * 8 + q), bf);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int j = q * 8 + e;
                v[j] = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]) + bf[e];
              }
#pragma unroll
              for (int e = 0; e < 8; e += 2) bf16_round2(v[q * 8 + e], v[q * 8 + e + 1]);   // nn.Linear output is a bf16 tensor
            }
            if (which < 2) {
              const uint32_t wn = (which == 0 ? vec_nq : vec_nk) + half * 128;
              const float* cs = p.rope_cos + token * 128 + half * 64;
              const float* sn = p.rope_sin + token * 128 + half * 64;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float wf[8];
                unpack8_bf16(ld_shared_v4(wn + q * 16), wf);
                float4 c0 = make_float4(1.f, 1.f, 1.f, 1.f), c1 = c0, s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
                if (row_ok) {
                  c0 = *reinterpret_cast<const float4*>(cs + q * 8); c1 = *reinterpret_cast<const float4*>(cs + q * 8 + 4);
                  s0 = *reinterpret_cast<const float4*>(sn + q * 8); s1 = *reinterpret_cast<const float4*>(sn + q * 8 + 4);
                }
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  float a = __fmul_rn(v[q * 8 + e], rs), b2 = __fmul_rn(v[q * 8 + e + 1], rs);
                  if (p.rms_round_first) bf16_round2(a, b2);                     // diffusers RMSNorm: .to(weight.dtype) before the weight multiply
                  a = __fmul_rn(a, wf[e]); b2 = __fmul_rn(b2, wf[e + 1]);
                  bf16_round2(a, b2);                                            // RMSNorm output is a bf16 tensor
                  v[q * 8 + e] = __fadd_rn(__fmul_rn(a, cc[e]), __fmul_rn(-b2, sv[e]));
                  v[q * 8 + e + 1] = __fadd_rn(__fmul_rn(b2, cc[e + 1]), __fmul_rn(a, sv[e + 1]));
                }
              }
              if (which == 1 && p.k_scale != 0.f) {       // keys carry softmax_scale * log2(e) (softmax.cuh): folded before the ONE bf16 store
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] *= p.k_scale;
              }
            }
            flush_chunk(c, v);
          }
        }
      } else
      for (int c = u_begin; c < u_end; ++c) {
        const int n0 = tn * BN + c * 64;       // global column of this 64-wide chunk
        // (1) residual mode: prefetched h chunk -> staging tile (16-B chunks XOR-swizzled by row); start the next chunk's loads
        if (p.epi == EPI_GATE_RESIDUAL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + coop_row;
            st_shared_v4x(stg + rr * 128 + ((coop_c ^ (rr & 7)) << 4), hv[i]);
          }
          if (c + 1 < u_end) load_residual(c + 1);
          __syncwarp();
        }
        // (2) thread == row: accumulator chunk -> registers -> fused math -> bf16 -> own row of the staging tile
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + c * 64, r0);
        tmem_ld32(t_row + c * 64 + 32, r1);
        tmem_ld_wait();
        float v[64];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float bf[8];
          unpack8_bf16(bias8(c * 8 + q), bf);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = q * 8 + e;
            v[j] = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]) + bf[e];
          }
#pragma unroll
          for (int e = 0; e < 8; e += 2) bf16_round2(v[q * 8 + e], v[q * 8 + e + 1]);   // nn.Linear output is a bf16 tensor
        }
        if (p.epi == EPI_QKV_RMSNORM) {
          // 64 columns == one attention head; q/k heads get RMSNorm (normalization.py:553-561)
          const int which = n0 / p.qk_dim;   // 0 = q, 1 = k, 2 = v
          if (which < 2) {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) ss += v[j] * v[j];
            const float rs = rsqrtf(ss * (1.0f / 64.0f) + p.eps);
            const uint32_t wn = which == 0 ? vec_nq : vec_nk;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float wf[8];
              unpack8_bf16(ld_shared_v4(wn + q * 16), wf);
#pragma unroll
              for (int e = 0; e < 8; e += 2) {
                float a = v[q * 8 + e] * rs, b = v[q * 8 + e + 1] * rs;
                bf16_round2(a, b);                                     // .to(weight.dtype) before the weight multiply
                v[q * 8 + e] = a * wf[e];
                v[q * 8 + e + 1] = b * wf[e + 1];
              }
            }
            if (which == 1 && p.k_scale != 0.f) {         // keys carry softmax_scale * log2(e) (softmax.cuh): folded before the ONE bf16 store
#pragma unroll
              for (int j = 0; j < 64; ++j) v[j] *= p.k_scale;
            }
          }
        } else if (p.epi == EPI_BIAS_GELU) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = gelu_tanh_f(v[j]);
        } else if (p.epi == EPI_GATE_RESIDUAL) {
          // h = h + gate[b,:] * y   in bf16 steps (attention.py:711-712, 726-728)
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float gf[8], hf[8];
            unpack8_bf16(__ldg(gate_v + c * 8 + q), gf);
            unpack8_bf16(ld_shared_v4(stg + lane * 128 + ((q ^ (lane & 7)) << 4)), hf);
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              float a = gf[e] * v[q * 8 + e], b = gf[e + 1] * v[q * 8 + e + 1];
              bf16_round2(a, b);                                       // gate * y is a bf16 tensor
              v[q * 8 + e] = hf[e] + a;
              v[q * 8 + e + 1] = hf[e + 1] + b;
            }
          }
        } else if (p.epi == EPI_BIAS_ADD_ROWTABLE) {
          // (latent + pos_embed).to(latent.dtype)   (embeddings.py:583)
          if (tile_ok && row_base + lane < p.rows_per_batch) {
            const float* t = p.row_table + static_cast<long>(row_base + lane) * p.N + n0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float4 t4 = *reinterpret_cast<const float4*>(t + q * 4);
              v[q * 4 + 0] += t4.x; v[q * 4 + 1] += t4.y; v[q * 4 + 2] += t4.z; v[q * 4 + 3] += t4.w;
            }
          }
        }
        flush_chunk(c, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);   // the leader's MMA warp waits for all 8 * epi_groups epilogue warps
    }
    }
    prof_end(pc0, 0x72 + ew);
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA may exit while its peer can still read its smem / arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static cudaError_t launch_bn(const GemmParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] { return cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); });
    if (e != cudaSuccess) return e;
  }
  const int units = ((p.num_batch * p.tiles_m_per_batch + 1) / 2) * (p.N / BN);
  const int max_clusters = num_sms / 2;
  const int grid = 2 * (units < max_clusters ? units : max_clusters);
  gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

int gemm_pick_bn(int N) { return (N % 256 == 0) ? 256 : ((N % 128 == 0) ? 128 : 64); }

// Column split of the epilogue between its two warp groups: on by default for every epilogue.  Measured on B200 (round 2, call 15,
// profiles/r02_kernel_bench_gemm.jsonl, TFLOP/s split / single group / the previous four-warp kernel): MLP-up M = 65536 1457 / 1181 / 1403,
// M = 8192 1424 / 1180 / 1348; attention out-projection M = 65536 1438 / 1118 / 1398; QKV M = 65536 1578 / 1546 / 1578 (M = 8192: 1452 /
// 1381 / 1489, the one shape that lost 2.5 %).  FFB200_GEMM_EPI_SPLIT = 1 | 2 forces a mode for A/B runs.
static int gemm_epi_split() {
  static const int forced = [] { const char* e = getenv("FFB200_GEMM_EPI_SPLIT"); return e ? atoi(e) : 0; }();
  return forced == 1 ? 1 : 2;
}

cudaError_t launch_gemm(const GemmParams& p0, int num_sms, cudaStream_t stream) {
  GemmParams p = p0;
  p.epi_split = gemm_epi_split();
  switch (p.bn) {
    case 256: return launch_bn<256>(p, num_sms, stream);
    case 128: return launch_bn<128>(p, num_sms, stream);
    case 64: return launch_bn<64>(p, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ffb
