// tcgen05 implicit-GEMM convolution for the VAE decoder that follows the rollout:
//   nn.Conv2d(k=3, s=1, p=1) / nn.Conv2d(k=1) / nn.Linear over pixel rows of AutoencoderKL's Decoder
//   (DF/models/autoencoders/vae.py:279-316, DF/models/resnet.py:319-377, DF/models/upsampling.py, attention_processor.py AttnProcessor2_0)
//
// Same persistent CTA-pair skeleton as gemm.cu (warp 0 TMA, warp 1 tcgen05.mma cta_group::2, warps 4-7 epilogue, two TMEM accumulator
// stages), with the A operand addressed as a 4-D NHWC tensor: the M tile is a th x tw patch of output pixels and the K loop walks
// taps x 64-channel blocks, every tap loading the SAME box shifted by (dy, dx).  The TMA unit zero-fills whatever falls outside the
// image (the convolution's padding), outside the channel range (Cin < 64) or outside the batch (ghost tile of an odd tile count),
// so there is no im2col buffer, no halo exchange and no bounds logic on the load side.  Weights are pre-packed [N][tap][Cin_pad].
//
// Validated on B200 in round 2 (tests/test_gpu_vae.py; the TMA zero-fill assumptions - boxes wider than the channel extent, negative and
// beyond-extent coordinates - hold on sm_100a).
#include "common.cuh"
#include "kernels.h"

namespace ffb {

__device__ __forceinline__ void tma_load_4d_2sm(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// m tile index -> (batch, first pixel row, first pixel column); a ghost tile maps to batch == B (all zeros)
__device__ __forceinline__ void conv_tile_origin(const ConvParams& p, int tm, int tiles_m, int& b, int& h0, int& w0) {
  if (tm >= tiles_m) { b = p.B; h0 = 0; w0 = 0; return; }
  const int per_img = p.tiles_h * p.tiles_w;
  b = tm / per_img;
  const int rem = tm - b * per_img;
  const int ty = rem / p.tiles_w;
  h0 = ty * (GEMM_BM >> p.tw_log2);
  w0 = (rem - ty * p.tiles_w) << p.tw_log2;
}

// Shared-memory layout of the convolution kernel: the GEMM kernel's pipeline stages with FOUR epilogue warps (one per TMEM lane quadrant)
// and per-warp bias vectors - the layout gemm.cu had before its epilogue was split over eight warps (the convolution epilogues are light).
constexpr int CONV_THREADS = 256;
template <int BN> struct ConvCfg {
  static constexpr int kStages = GemmCfg<BN>::kStages;
  static constexpr int kABytes = GemmCfg<BN>::kABytes;
  static constexpr int kBBytes = GemmCfg<BN>::kBBytes;
  static constexpr int kStageBytes = GemmCfg<BN>::kStageBytes;
  static constexpr int kTmemCols = GemmCfg<BN>::kTmemCols;
  static constexpr int kStagingBytes = 4 * 32 * 128;        // per epilogue warp: 32 rows x 64 bf16
  static constexpr int kVecBytes = 4 * (2 * BN * 2 + 2 * 128 * 2);  // per warp: bias[BN] (+ the slots the GEMM epilogues used)
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kVecBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget of one CTA per SM");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1)
conv_bf16_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + Cfg::kStages * Cfg::kABytes;
  const uint32_t smem_stage = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t smem_vec = smem_stage + Cfg::kStagingBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes + Cfg::kVecBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty = bars + 2 * Cfg::kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_m = p.B * p.tiles_h * p.tiles_w;
  const int pairs_m = (tiles_m + 1) >> 1;
  const int tiles_n = p.N / BN;
  const int num_units = pairs_m * tiles_n;
  const int num_kb = p.taps * p.kc;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      int mp, tn;
      unit_to_tile(unit, pairs_m, tiles_n, p.band, mp, tn);
      int b, h0, w0;
      conv_tile_origin(p, 2 * mp + static_cast<int>(cta_rank), tiles_m, b, h0, w0);
      int tap = 0, kcb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 0x40);
        if (elect_one()) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
          const int dx = p.taps == 9 ? tap - (tap / 3) * 3 - 1 : 0;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          tma_load_4d_2sm(smem + stage * Cfg::kABytes, &p.tmA, &full_bar[stage], kcb * GEMM_BK, w0 + dx, h0 + dy, b);
          tma_load_2d_2sm(smem + Cfg::kStages * Cfg::kABytes + stage * Cfg::kBBytes, &p.tmB, &full_bar[stage], kb * GEMM_BK,
                          tn * BN + static_cast<int>(cta_rank) * (BN / 2));
        }
        __syncwarp();
        if (++kcb == p.kc) { kcb = 0; ++tap; }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; M = 256 across the pair) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, 0, 0);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 0x41);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 0x42);
          tc_fence_after();
          const uint32_t a_addr = smem_a + stage * Cfg::kABytes;
          const uint32_t b_addr = smem_b + stage * Cfg::kBBytes;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              umma_bf16_2sm(d_tmem, desc_kmajor_sw128(a_addr + k * 32), desc_kmajor_sw128(b_addr + k * 32), idesc,
                            (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 0x3);
            if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc], 0x3);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread == output pixel =====================
    const int ew = warp - 4;
    const uint32_t stg = smem_stage + ew * (32 * 128);
    const uint32_t vec_bias = smem_vec + ew * (2 * BN * 2 + 512);
    const int coop_row = lane >> 3, coop_c = lane & 7;
    const int tw_mask = (1 << p.tw_log2) - 1;
    int it = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int mp, tn;
      unit_to_tile(unit, pairs_m, tiles_n, p.band, mp, tn);
      const int tm = 2 * mp + static_cast<int>(cta_rank);
      const bool tile_ok = tm < tiles_m;
      int b, h0, w0;
      conv_tile_origin(p, tm, tiles_m, b, h0, w0);
      // pixel (row of the implicit GEMM) handled by tile row r: valid flag and pixel index inside the [B, H, W] grid
      auto pixel_of = [&](int r, long& pix) -> bool {
        const int h = h0 + (r >> p.tw_log2), w = w0 + (r & tw_mask);
        pix = (static_cast<long>(b) * p.H + h) * p.W + w;
        return tile_ok && h < p.H && w < p.W;
      };
      if (lane * 8 < BN) {
        uint4 bv = make_uint4(0, 0, 0, 0);
        if (p.bias && tn * BN + lane * 8 < ((p.n_store + 7) & ~7))   // bias vectors are padded to a multiple of 8 entries, not to N
          bv = __ldg(reinterpret_cast<const uint4*>(p.bias + tn * BN) + lane);
        st_shared_v4x(vec_bias + lane * 16, bv);
      }
      __syncwarp();
      mbar_wait(&tmem_full[acc], acc_phase, 0x43);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
      long my_pix;
      const bool my_ok = pixel_of(ew * 32 + lane, my_pix);

      for (int c = 0; c < BN / 64; ++c) {
        const int n0 = tn * BN + c * 64;
        if (p.epi == EPI_CONV_RESIDUAL) {
          // residual chunk -> staging tile with whole-line loads (8 lanes per 128-B row, 4 rows per instruction)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + coop_row;
            long pix;
            uint4 hv = make_uint4(0, 0, 0, 0);
            if (pixel_of(ew * 32 + rr, pix) && n0 + coop_c * 8 < p.n_store)
              hv = *reinterpret_cast<const uint4*>(p.residual + pix * p.ldr + n0 + coop_c * 8);
            st_shared_v4x(stg + rr * 128 + ((coop_c ^ (rr & 7)) << 4), hv);
          }
          __syncwarp();
        }
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + c * 64, r0);
        tmem_ld32(t_row + c * 64 + 32, r1);
        tmem_ld_wait();
        float v[64];
        if (p.epi == EPI_CONV_F32) {
          // fp32 scores: two half chunks of 32 columns through the same 32 x 128 B staging tile
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                w4[e] = __float_as_uint(__uint_as_float(half == 0 ? r0[q * 4 + e] : r1[q * 4 + e]) * p.out_scale);
              st_shared_v4(stg + lane * 128 + ((q ^ (lane & 7)) << 4), w4[0], w4[1], w4[2], w4[3]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = i * 4 + coop_row;
              long pix;
              if (pixel_of(ew * 32 + rr, pix) && n0 + half * 32 + coop_c * 4 < p.n_store)
                *reinterpret_cast<uint4*>(p.out_f32 + pix * p.ldo_f32 + n0 + half * 32 + coop_c * 4) =
                    ld_shared_v4(stg + rr * 128 + ((coop_c ^ (rr & 7)) << 4));
            }
            __syncwarp();
          }
          continue;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float bf[8];
          unpack8_bf16(ld_shared_v4(vec_bias + (c * 64 + q * 8) * 2), bf);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = q * 8 + e;
            v[j] = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]) + bf[e];
          }
#pragma unroll
          for (int e = 0; e < 8; e += 2) bf16_round2(v[q * 8 + e], v[q * 8 + e + 1]);   // the convolution's output is a bf16 tensor
        }
        if (p.epi == EPI_CONV_NCHW) {
          // planar image store: for a fixed channel consecutive lanes are consecutive pixels of an image row
          if (my_ok && c == 0) {
            const long plane = static_cast<long>(p.H) * p.W;
            const long in_img = my_pix - static_cast<long>(b) * plane;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              if (ch < p.n_store) p.out[(static_cast<long>(b) * p.n_store + ch) * plane + in_img] = __float2bfloat16_rn(v[ch]);
          }
          continue;
        }
        if (p.epi == EPI_CONV_RESIDUAL) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float hf[8];
            unpack8_bf16(ld_shared_v4(stg + lane * 128 + ((q ^ (lane & 7)) << 4)), hf);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q * 8 + e] += hf[e];
          }
          __syncwarp();
        }
        // bf16 chunk -> own row of the swizzled staging tile -> whole-line stores
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st_shared_v4(stg + lane * 128 + ((q ^ (lane & 7)) << 4), pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]),
                       pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]), pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]),
                       pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + coop_row;
          long pix;
          if (pixel_of(ew * 32 + rr, pix) && n0 + coop_c * 8 < p.n_store)
            *reinterpret_cast<uint4*>(p.out + pix * p.ldo + n0 + coop_c * 8) = ld_shared_v4(stg + rr * 128 + ((coop_c ^ (rr & 7)) << 4));
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static cudaError_t launch_conv_bn(const ConvParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] { return cudaFuncSetAttribute(conv_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); });
    if (e != cudaSuccess) return e;
  }
  const int units = ((p.B * p.tiles_h * p.tiles_w + 1) / 2) * (p.N / BN);
  const int max_clusters = num_sms / 2;
  const int grid = 2 * (units < max_clusters ? units : max_clusters);
  conv_bf16_kernel<BN><<<grid, CONV_THREADS, Cfg::kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_conv(const ConvParams& p, int num_sms, cudaStream_t stream) {
  if (p.N % p.bn != 0 || (p.taps != 1 && p.taps != 9) || p.tw_log2 < 3 || p.tw_log2 > 7) return cudaErrorInvalidValue;
  if (p.epi == EPI_CONV_NCHW && p.n_store > 8) return cudaErrorInvalidValue;
  switch (p.bn) {
    case 256: return launch_conv_bn<256>(p, num_sms, stream);
    case 128: return launch_conv_bn<128>(p, num_sms, stream);
    case 64: return launch_conv_bn<64>(p, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ffb
