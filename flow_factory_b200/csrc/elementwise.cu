// HBM-bound pieces of the denoise step (coalesced 16-byte vector access, warp-shuffle reductions):
//   ln_modulate      : LayerNorm(no affine) + adaLN-Zero modulate            (normalization.py:120-126,167-169,348-351)
//   small_linear     : skinny-batch Linear (timestep/pooled MLPs, all-layer adaLN GEMV)  (embeddings.py:1294-1306,2213-2217)
//   timestep_proj    : 256-ch sinusoid                                        (embeddings.py:26-77)
//   patchify         : im2col for the 2x2 stride-2 patch-embed conv           (embeddings.py:559)
//   sde_step         : CFG combine + Euler/SDE update + noise + fp16 round-trip + Gaussian log-prob
//                      (sd3_5.py:431-433 ; flow_match_euler_discrete.py:309-420 ; abc.py:172-182)
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>

namespace ffb {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm + modulate.  One warp per token row; the row stays in registers as the raw 16-byte bf16 chunks it was loaded as
// (<= 48 registers instead of 96 fp32 values: twice the resident warps, which is what hides the HBM latency of this pure
// streaming kernel) and is unpacked on the fly in each of the three passes (mean, variance, output).  D <= 3072, D % 8 == 0.
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXC = 12;  // 16-byte chunks per lane (D = 1536: 6, D = 3072: 12)

__global__ void __launch_bounds__(256, 4) ln_modulate_kernel(const LnModParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int rows = p.rows_per_batch * p.num_batch;
  if (warp >= rows) return;
  const int b = warp / p.rows_per_batch;
  const int nchunk = p.D >> 3;
  const int rib = warp - b * p.rows_per_batch;   // row inside the batch
  const long dense = static_cast<long>(p.rows_per_batch) * p.D;
  const bf16* xr = p.x + static_cast<long>(b) * (p.x_batch_stride ? p.x_batch_stride : dense) + static_cast<long>(rib) * p.D;
  const long orow = static_cast<long>(b) * (p.out_batch_stride ? p.out_batch_stride : dense) + static_cast<long>(rib) * p.D;
  uint4 raw[LN_MAXC];
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    const int c = lane + 32 * i;
    raw[i] = make_uint4(0, 0, 0, 0);
    if (c < nchunk) raw[i] = *reinterpret_cast<const uint4*>(xr + c * 8);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    if (lane + 32 * i < nchunk) {
      float v[8];
      unpack8(raw[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[e];
    }
  }
  const float mean = warp_sum(sum) / static_cast<float>(p.D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    if (lane + 32 * i < nchunk) {
      float v[8];
      unpack8(raw[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[e] - mean; sq += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(p.D) + p.eps);
  const long mo = static_cast<long>(b) * p.mod_batch_stride;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunk) {
      float y[8], sc[8], sh[8], o[8];
      unpack8(raw[i], y);
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = (y[e] - mean) * rstd;
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.scale1 + mo + c * 8)), sc);
      unpack8(__ldg(reinterpret_cast<const uint4*>(p.shift1 + mo + c * 8)), sh);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = __fadd_rn(__fmul_rn(y[e], bf16_round(1.0f + sc[e])), sh[e]);
      *reinterpret_cast<uint4*>(p.out1 + orow + c * 8) = pack8(o);
      if (p.out2 != nullptr) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(p.scale2 + mo + c * 8)), sc);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p.shift2 + mo + c * 8)), sh);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __fadd_rn(__fmul_rn(y[e], bf16_round(1.0f + sc[e])), sh[e]);
        *reinterpret_cast<uint4*>(p.out2 + orow + c * 8) = pack8(o);
      }
    }
  }
}

// Measured and NOT adopted (round 2, call 22, tools/ln_bench.py -> profiles/r02_ln_modulate_experiment.md): a persistent grid whose warps
// prefetch their next row (two rows in registers, streaming loads / stores that bypass L1) - 0.113 / 0.184 ms (single / dual output, 16 x 4096
// x 1536) against 0.106 / 0.158 ms for this one-row-per-warp kernel: the 16 warps per SM that fit with two rows in registers carry fewer
// bytes in flight than 32 short-lived warps.
cudaError_t launch_ln_modulate(const LnModParams& p, cudaStream_t stream) {
  if (p.D % 8 != 0 || p.D > LN_MAXC * 32 * 8) return cudaErrorInvalidValue;
  const long rows = static_cast<long>(p.rows_per_batch) * p.num_batch;
  const int wpb = 8;
  const int grid = static_cast<int>((rows + wpb - 1) / wpb);
  ln_modulate_kernel<<<grid, wpb * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// small_linear: out[b, n] = bf16(bf16(acc + bias) + addend), batch <= 32 handled 8 rows at a time;
// one warp per output feature n; weights streamed once per batch chunk with 16-byte loads.
// ------------------------------------------------------------------------------------------------
constexpr int SL_BCHUNK = 16;  // batch rows per pass: the whole CFG batch of the bench (2 x 8) streams the weights once
constexpr int SL_WARPS = 8;

__global__ void __launch_bounds__(SL_WARPS * 32) small_linear_kernel(const SmallLinearParams p) {
  extern __shared__ __align__(16) uint8_t sl_smem[];
  bf16* xin = reinterpret_cast<bf16*>(sl_smem);  // [SL_BCHUNK][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * SL_BCHUNK;
  const int nb = min(SL_BCHUNK, p.batch - b0);
  const int kc = p.K >> 3;
  for (int i = threadIdx.x; i < nb * kc; i += blockDim.x) {
    const int bb = i / kc, c = i % kc;
    uint4 u = *reinterpret_cast<const uint4*>(p.in + static_cast<long>(b0 + bb) * p.in_stride + c * 8);
    if (p.silu_input) {
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __fdividef(f[e], 1.0f + __expf(-f[e]));  // F.silu, then bf16
      u = pack8(f);
    }
    *reinterpret_cast<uint4*>(xin + bb * p.K + c * 8) = u;
  }
  __syncthreads();
  // persistent over groups of SL_WARPS output features: the staged inputs are reused for every group this CTA handles
  for (int n = blockIdx.x * SL_WARPS + warp; n < p.N; n += gridDim.x * SL_WARPS) {
    float acc[SL_BCHUNK];
#pragma unroll
    for (int i = 0; i < SL_BCHUNK; ++i) acc[i] = 0.f;
    const bf16* wr = p.W + static_cast<long>(n) * p.K;
    for (int c = lane; c < kc; c += 32) {
      float wf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(wr + c * 8)), wf);
#pragma unroll
      for (int bb = 0; bb < SL_BCHUNK; ++bb) {
        if (bb < nb) {
          float xf[8];
          unpack8(*reinterpret_cast<const uint4*>(xin + bb * p.K + c * 8), xf);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[bb] = fmaf(wf[e], xf[e], acc[bb]);
        }
      }
    }
#pragma unroll
    for (int bb = 0; bb < SL_BCHUNK; ++bb) acc[bb] = warp_sum(acc[bb]);
    if (lane == 0) {
      const float bias = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
      for (int bb = 0; bb < nb; ++bb) {
        float y = bf16_round(acc[bb] + bias);
        if (p.addend) y = y + __bfloat162float(p.addend[static_cast<long>(b0 + bb) * p.addend_stride + n]);
        if (p.addend2) y = bf16_round(y) + __bfloat162float(p.addend2[static_cast<long>(b0 + bb) * p.addend_stride + n]);
        p.out[static_cast<long>(b0 + bb) * p.out_stride + n] = __float2bfloat16_rn(y);
      }
    }
  }
}

// Tensor-core variant (mma.sync m16n8k16, bf16 -> fp32): the batch (<= 16 rows per pass, zero padded) is the M dimension, eight
// output features the N dimension.  The stacked adaLN matrix is 1.5 GB (SD3.5-medium) / 6.5 GB (FLUX.1-dev) per denoise step and
// the SIMT kernel above spends ~20 instructions per 16 weight bytes and batch row (0.9 TB/s); here a warp issues one 16-byte
// weight load, two 16-byte shared-memory loads and two MMAs per 32 k, so the kernel streams the weights at HBM speed.
// The k index inside a 32-wide block is permuted identically for A and B (a dot product does not care): lane q = lane % 4 owns
// k = kb + 8q .. 8q+7 of its weight row (one 16-byte load) and feeds k-slots {2q, 2q+1, 2q+8, 2q+9} of two consecutive MMAs.
constexpr int SLM_WARPS = 8;
// staged input rows are padded so that the row stride is 64 bytes modulo 128: the 16-byte A loads of a quarter warp (two batch rows x
// four k-chunks) then cover all 32 banks exactly once
__host__ __device__ constexpr int slm_ld(int K) { return K + ((K & 63) == 0 ? 32 : 0); }

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(SLM_WARPS * 32) small_linear_mma_kernel(const SmallLinearParams p) {
  extern __shared__ __align__(16) uint8_t sl_smem[];
  bf16* xin = reinterpret_cast<bf16*>(sl_smem);  // [16][slm_ld(K)], rows >= nb are zero
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * 16;
  const int nb = min(16, p.batch - b0);
  const int kc = p.K >> 3, ldx = slm_ld(p.K);
  for (int i = threadIdx.x; i < 16 * kc; i += blockDim.x) {
    const int bb = i / kc, c = i % kc;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (bb < nb) {
      u = *reinterpret_cast<const uint4*>(p.in + static_cast<long>(b0 + bb) * p.in_stride + c * 8);
      if (p.silu_input) {
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __fdividef(f[e], 1.0f + __expf(-f[e]));  // F.silu, then bf16
        u = pack8(f);
      }
    }
    *reinterpret_cast<uint4*>(xin + bb * ldx + c * 8) = u;
  }
  __syncthreads();
  const int g = lane >> 2, q = lane & 3;
  const bf16* xa = xin + g * ldx + q * 8;          // batch row g
  const bf16* xb = xin + (g + 8) * ldx + q * 8;    // batch row g + 8
  const int tiles = (p.N + 7) >> 3;
  for (int t = blockIdx.x * SLM_WARPS + warp; t < tiles; t += gridDim.x * SLM_WARPS) {
    const int n = t * 8 + g;                                            // this lane's weight row (B-fragment column)
    const bf16* wr = p.W + static_cast<long>(n < p.N ? n : p.N - 1) * p.K + q * 8;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int kb = 0; kb < p.K; kb += 32) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(wr + kb));
      const uint4 a = *reinterpret_cast<const uint4*>(xa + kb);
      const uint4 b = *reinterpret_cast<const uint4*>(xb + kb);
      mma_bf16_16816(c, a.x, b.x, a.y, b.y, w.x, w.y);
      mma_bf16_16816(c, a.z, b.z, a.w, b.w, w.z, w.w);
    }
    // c[0], c[1]: batch row g, features t*8 + 2q, +1 ; c[2], c[3]: batch row g + 8
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int bb = g + 8 * h;
      if (bb >= nb) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int nn = t * 8 + 2 * q + e;
        if (nn >= p.N) continue;
        float y = bf16_round(c[2 * h + e] + (p.bias ? __bfloat162float(p.bias[nn]) : 0.f));
        if (p.addend) y = y + __bfloat162float(p.addend[static_cast<long>(b0 + bb) * p.addend_stride + nn]);
        if (p.addend2) y = bf16_round(y) + __bfloat162float(p.addend2[static_cast<long>(b0 + bb) * p.addend_stride + nn]);
        p.out[static_cast<long>(b0 + bb) * p.out_stride + nn] = __float2bfloat16_rn(y);
      }
    }
  }
}

cudaError_t launch_small_linear(const SmallLinearParams& p, cudaStream_t stream) {
  if (p.K % 8 != 0) return cudaErrorInvalidValue;
  if (p.K % 32 == 0 && p.in_stride % 8 == 0) {
    const int smem = 16 * slm_ld(p.K) * 2;
    static int max_smem_mma = 48 * 1024;
    if (smem > max_smem_mma) {
      cudaError_t e = cudaFuncSetAttribute(small_linear_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
      max_smem_mma = smem;
    }
    const int groups = ((p.N + 7) / 8 + SLM_WARPS - 1) / SLM_WARPS;
    dim3 grid(std::min(groups, 148 * 4), (p.batch + 15) / 16);
    small_linear_mma_kernel<<<grid, SLM_WARPS * 32, smem, stream>>>(p);
    return cudaGetLastError();
  }
  const int smem = SL_BCHUNK * p.K * 2;
  static int max_smem = 48 * 1024;
  if (smem > max_smem) {
    cudaError_t e = cudaFuncSetAttribute(small_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    max_smem = smem;
  }
  const int groups = (p.N + SL_WARPS - 1) / SL_WARPS;
  dim3 grid(std::min(groups, 148 * 4), (p.batch + SL_BCHUNK - 1) / SL_BCHUNK);
  small_linear_kernel<<<grid, SL_WARPS * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// timestep sinusoid: out[b, 0:128] = cos(t * f_i), out[b, 128:256] = sin(t * f_i), f_i = exp(-ln(1e4) * i / 128)
// ------------------------------------------------------------------------------------------------
__global__ void timestep_proj_kernel(const StepCoef* table, const int* step_ptr, int index, int batch, bf16* out, float post_scale) {
  const int i = threadIdx.x;  // 0..127
  const float t = table[step_ptr ? *step_ptr : index].t_model;
  const float exponent = (-9.210340371976184f * static_cast<float>(i)) / 128.0f;
  const float f = expf(exponent);
  const float a = __fmul_rn(post_scale, __fmul_rn(t, f));   // get_timestep_embedding: emb = scale * (t * f)  (embeddings.py:62-65)
  const bf16 c = __float2bfloat16_rn(cosf(a)), s = __float2bfloat16_rn(sinf(a));
  for (int b = 0; b < batch; ++b) {
    out[b * 256 + i] = c;
    out[b * 256 + 128 + i] = s;
  }
}
cudaError_t launch_timestep_proj(const StepCoef* table, const int* step_ptr, int index, int batch, bf16* out, cudaStream_t stream,
                                 float post_scale) {
  timestep_proj_kernel<<<1, 128, 0, stream>>>(table, step_ptr, index, batch, out, post_scale);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// latent storage dtype (cast_latents, FF/models/abc.py:172-182): fp16 (default; +-65504 clamp + sticky flag), bf16 or fp32
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lat_load(const void* base, long i, int st) {
  if (st == LAT_F16) return __half2float(static_cast<const __half*>(base)[i]);
  if (st == LAT_BF16) return __bfloat162float(static_cast<const bf16*>(base)[i]);
  return static_cast<const float*>(base)[i];
}
// the storage round trip of a freshly sampled value (flow_match...py:359-362: `.to(input dtype).float()`), before any clamp
__device__ __forceinline__ float lat_round(float v, int st) {
  if (st == LAT_F16) return __half2float(__float2half_rn(v));
  if (st == LAT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  return v;
}
// store one value (already through lat_round, or teacher-forced / ODE mean): fp16 clamps to the finite range and raises the flag
__device__ __forceinline__ void lat_store(void* base, long i, int st, float v, int* overflow_flag) {
  if (st == LAT_F16) {
    if (fabsf(v) > 65504.0f) { v = copysignf(65504.0f, v); if (overflow_flag) *overflow_flag = 1; }
    static_cast<__half*>(base)[i] = __float2half_rn(v);
  } else if (st == LAT_BF16) {
    static_cast<bf16*>(base)[i] = __float2bfloat16_rn(v);
  } else {
    static_cast<float*>(base)[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// patchify (im2col): out[(r*B + b)*Ni + i*wp + j][c*p*p + py*p + px] = bf16(x[b, c, i*p+py, j*p+px])
// ------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const void* x, int storage, int B, int reps, int C, int H, int W, int patch, bf16* out) {
  const int hp = H / patch, wp = W / patch;
  const int KK = C * patch * patch;
  const long total = static_cast<long>(B) * hp * wp * KK;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % KK);
    const long tok = idx / KK;
    const int j = static_cast<int>(tok % wp);
    const int i = static_cast<int>((tok / wp) % hp);
    const int b = static_cast<int>(tok / (static_cast<long>(wp) * hp));
    const int px = k % patch, py = (k / patch) % patch, c = k / (patch * patch);
    const float v = lat_load(x, ((static_cast<long>(b) * C + c) * H + i * patch + py) * W + j * patch + px, storage);
    const bf16 o = __float2bfloat16_rn(v);
    for (int r = 0; r < reps; ++r)
      out[((static_cast<long>(r) * B + b) * hp * wp + static_cast<long>(i) * wp + j) * KK + k] = o;
  }
}
cudaError_t launch_patchify(const void* x, int storage, int B, int reps, int C, int H, int W, int patch, bf16* out, cudaStream_t stream) {
  const long total = static_cast<long>(B) * (H / patch) * (W / patch) * C * patch * patch;
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148 * 8));
  patchify_kernel<<<grid, 256, 0, stream>>>(x, storage, B, reps, C, H, W, patch, out);
  return cudaGetLastError();
}

__global__ void cast_f32_bf16_kernel(const float* in, bf16* out, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
// ------------------------------------------------------------------------------------------------
// diffusers RMSNorm over whole rows (QwenImage txt_norm, DF/models/normalization.py:553-567): variance in fp32, x * rsqrt(var + eps)
// in fp32 -> bf16 -> x bf16 weight.  One warp per row, two passes over the (L1-resident) row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rms_norm_rows_kernel(const bf16* x, const bf16* weight, bf16* out, long rows, int K, float eps) {
  const long row = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* xr = x + row * K;
  const int nchunk = K >> 3;
  float ss = 0.f;
  for (int c = lane; c < nchunk; c += 32) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + c * 8), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
  }
  const float rs = rsqrtf(warp_sum(ss) / static_cast<float>(K) + eps);
  for (int c = lane; c < nchunk; c += 32) {
    float v[8], wv[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + c * 8), v);
    unpack8(__ldg(reinterpret_cast<const uint4*>(weight + c * 8)), wv);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = __fmul_rn(bf16_round(__fmul_rn(v[e], rs)), wv[e]);
    *reinterpret_cast<uint4*>(out + row * K + c * 8) = pack8(o);
  }
}
cudaError_t launch_rms_norm_rows(const bf16* x, const bf16* weight, bf16* out, long rows, int K, float eps, cudaStream_t stream) {
  if (K % 8 != 0) return cudaErrorInvalidValue;
  const int grid = static_cast<int>((rows + 7) / 8);
  rms_norm_rows_kernel<<<grid, 256, 0, stream>>>(x, weight, out, rows, K, eps);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Qwen-Image true CFG with per-token norm rescale (FF/models/qwen_image/qwen_image.py:580-587), all in bf16 tensor steps:
//   comb = neg + g * (pos - neg) ; pred = comb * (||pos|| / ||comb||)   with the norms over the last (64-wide) dim.
// v: bf16 [2B, Ni, 64] (negative half first); out: bf16 [B, Ni, 64].  One thread per token (a 128-byte row).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cfg_norm_rescale_kernel(const bf16* v, bf16* out, long tokens_per_half, float g) {
  const long tok = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (tok >= tokens_per_half) return;
  const uint4* neg = reinterpret_cast<const uint4*>(v + tok * 64);
  const uint4* pos = reinterpret_cast<const uint4*>(v + (tokens_per_half + tok) * 64);
  float comb[64];
  float sp = 0.f, sc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float n[8], p[8];
    unpack8(neg[c], n);
    unpack8(pos[c], p);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = bf16_round(p[e] - n[e]);
      const float gd = bf16_round(g * d);          // python scalar x bf16 tensor: fp32 math, bf16 result
      const float cb = bf16_round(n[e] + gd);
      comb[c * 8 + e] = cb;
      sp += p[e] * p[e];
      sc += cb * cb;
    }
  }
  const float ratio = bf16_round(bf16_round(sqrtf(sp)) / bf16_round(sqrtf(sc)));
  uint4* o = reinterpret_cast<uint4*>(out + tok * 64);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = comb[c * 8 + e] * ratio;
    o[c] = pack8(r);
  }
}
cudaError_t launch_cfg_norm_rescale(const bf16* v, bf16* out, long tokens_per_half, float g, cudaStream_t stream) {
  const int grid = static_cast<int>((tokens_per_half + 127) / 128);
  cfg_norm_rescale_kernel<<<grid, 128, 0, stream>>>(v, out, tokens_per_half, g);
  return cudaGetLastError();
}

__global__ void cast_f16_bf16_kernel(const __half* in, bf16* out, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(__half2float(in[i]));
}
cudaError_t launch_cast_f16_to_bf16(const __half* in, bf16* out, long n, cudaStream_t stream) {
  const int grid = static_cast<int>(std::min<long>((n + 255) / 256, 148 * 16));
  cast_f16_bf16_kernel<<<grid, 256, 0, stream>>>(in, out, n);
  return cudaGetLastError();
}
cudaError_t launch_cast_f32_to_bf16(const float* in, bf16* out, long n, cudaStream_t stream) {
  const int grid = static_cast<int>(std::min<long>((n + 255) / 256, 148 * 16));
  cast_f32_bf16_kernel<<<grid, 256, 0, stream>>>(in, out, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (counter = element-quad index, step; key = seed)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void box_muller(uint32_t u0, uint32_t u1, float* z0, float* z1) {
  const float a = (static_cast<float>(u0) + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  const float b = static_cast<float>(u1) * 2.3283064365386963e-10f;           // [0, 1)
  const float rad = sqrtf(-2.0f * logf(a));
  float s, c;
  sincosf(6.283185307179586f * b, &s, &c);
  *z0 = rad * c; *z1 = rad * s;
}

// ------------------------------------------------------------------------------------------------
// Fused CFG + Euler/SDE step + log-prob.  Each thread owns 4 consecutive pixels of one (b, c, y) row.
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Per-element pieces of scheduler.step, shared by the standalone kernel and the fused proj_out epilogue.
// fp32, reference operation order, no FMA contraction.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cfg_combine_bf16(float vu, float vt, float g) {   // sd3_5.py:431-433 on bf16 tensors
  return bf16_round(vu + bf16_round(g * bf16_round(vt - vu)));
}
__device__ __forceinline__ float sde_mean(const StepCoef& k, float x, float v) {
  if (k.dynamics == DYN_ODE) return __fadd_rn(x, __fmul_rn(v, k.dt));
  if (k.dynamics == DYN_FLOW_SDE) return __fadd_rn(__fmul_rn(x, k.c_x), __fmul_rn(__fmul_rn(v, k.c_v), k.dt));
  if (k.dynamics == DYN_DANCE_SDE) {
    const float x0p = __fsub_rn(x, __fmul_rn(k.sigma, v));
    const float num = __fmul_rn(k.c_x /* 0.5*eta^2 */, __fsub_rn(x, __fmul_rn(x0p, k.c_v /* 1-sigma */)));
    const float lt = __fdiv_rn(num, __fmul_rn(k.sigma, k.sigma));
    return __fadd_rn(x, __fmul_rn(__fadd_rn(v, lt), k.dt));
  }
  const float x0p = __fsub_rn(x, __fmul_rn(k.sigma, v));                 // CPS
  const float x1p = __fadd_rn(x, __fmul_rn(v, k.c_v /* 1-sigma */));
  return __fadd_rn(__fmul_rn(x0p, k.cps_a), __fmul_rn(x1p, k.cps_b));
}
// mean + scale*z, rounded through the storage dtype (flow_match...py:359-362)
__device__ __forceinline__ float sde_sample(const StepCoef& k, float mean, float z, int storage) {
  return lat_round(__fadd_rn(mean, __fmul_rn(k.noise_scale, z)), storage);
}
__device__ __forceinline__ float sde_logp_term(const StepCoef& k, float nxt, float mean) {
  const float d = __fsub_rn(nxt, mean);
  const float d2 = __fmul_rn(d, d);
  return k.dynamics == DYN_CPS ? -d2 : __fdiv_rn(-d2, k.two_var);
}

constexpr int SDE_THREADS = 256;

__global__ void __launch_bounds__(SDE_THREADS) sde_step_kernel(const SdeStepParams p) {
  const int sidx = p.step_ptr ? *p.step_ptr : p.coef_index;
  const StepCoef k = p.coef_table[sidx];
  const int b = blockIdx.y;
  const float* noise = p.noise ? p.noise + static_cast<long>(sidx) * p.noise_step_stride : nullptr;
  const int CHW = p.C * p.H * p.W;
  const int st = p.storage;
  const long traj_off = static_cast<long>(b) * p.traj_batch_stride + static_cast<long>(k.store_slot) * CHW;   // elements
  const bool has_traj = p.traj != nullptr && k.store_slot >= 0;
  const int quads = CHW >> 2;
  float part = 0.f;
  const int hp = p.H / p.patch, wp = p.W / p.patch;
  const int ntok = hp * wp;
  const int vch = p.patch * p.patch * p.C;
  for (int qd = blockIdx.x * SDE_THREADS + threadIdx.x; qd < quads; qd += gridDim.x * SDE_THREADS) {
    const int e0 = qd << 2;
    const int x0 = e0 % p.W;
    const int y = (e0 / p.W) % p.H;
    const int c = e0 / (p.W * p.H);
    const long base = static_cast<long>(b) * CHW + e0;
    // ---- noise prediction (with CFG) ----
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float vc;
      if (p.v_tokens != nullptr) {
        const int xx = x0 + i;
        const int tok = (y / p.patch) * wp + xx / p.patch;
        const int n = ((y % p.patch) * p.patch + (xx % p.patch)) * p.C + c;   // "nhwpqc"
        if (p.cfg) {
          const float vu = __bfloat162float(p.v_tokens[(static_cast<long>(b) * ntok + tok) * vch + n]);
          const float vt = __bfloat162float(p.v_tokens[(static_cast<long>(p.B + b) * ntok + tok) * vch + n]);
          vc = cfg_combine_bf16(vu, vt, p.guidance);
        } else {
          vc = __bfloat162float(p.v_tokens[(static_cast<long>(b) * ntok + tok) * vch + n]);
        }
      } else {
        vc = __bfloat162float(p.v_direct[base + i]);
      }
      v[i] = vc;
    }
    // ---- current latents ----
    float xs[4];
    if (st == LAT_F16) {
      const uint2 xr = *reinterpret_cast<const uint2*>(static_cast<const __half*>(p.x) + base);
      const __half2 xa = *reinterpret_cast<const __half2*>(&xr.x), xb = *reinterpret_cast<const __half2*>(&xr.y);
      xs[0] = __low2float(xa); xs[1] = __high2float(xa); xs[2] = __low2float(xb); xs[3] = __high2float(xb);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) xs[i] = lat_load(p.x, base + i, st);
    }
    // ---- mean ----
    float mean[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) mean[i] = sde_mean(k, xs[i], v[i]);
    // ---- next sample ----
    float nxt[4];
    if (p.next_given != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) nxt[i] = lat_load(p.next_given, base + i, st);
    } else if (k.dynamics == DYN_ODE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) nxt[i] = mean[i];
    } else {
      float z[4];
      if (noise != nullptr) {
        const float4 nz = *reinterpret_cast<const float4*>(noise + base);
        z[0] = nz.x; z[1] = nz.y; z[2] = nz.z; z[3] = nz.w;
      } else {
        uint32_t r[4];
        philox4x32_10(static_cast<uint32_t>(qd), static_cast<uint32_t>(b), static_cast<uint32_t>(sidx), 0x5DEu,
                      static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32), r);
        box_muller(r[0], r[1], &z[0], &z[1]);
        box_muller(r[2], r[3], &z[2], &z[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) nxt[i] = sde_sample(k, mean[i], z[i], st);
    }
    // ---- store (cast_latents: fp16 clamps to +-65504 on overflow) ----
    if (p.x_next != nullptr || has_traj) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (p.x_next) lat_store(p.x_next, base + i, st, nxt[i], p.overflow_flag);
        if (has_traj) lat_store(p.traj, traj_off + e0 + i, st, nxt[i], p.overflow_flag);
      }
    }
    if (p.mean_out) *reinterpret_cast<float4*>(p.mean_out + base) = make_float4(mean[0], mean[1], mean[2], mean[3]);
    if (p.v_out) {
      uint2 vo;
      vo.x = pack_bf16x2(v[0], v[1]); vo.y = pack_bf16x2(v[2], v[3]);
      *reinterpret_cast<uint2*>(p.v_out + base) = vo;
    }
    // ---- log-prob terms ----
    if (k.compute_log_prob && k.dynamics != DYN_ODE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) part += sde_logp_term(k, nxt[i], mean[i]);
    }
  }
  // block reduce -> logp_partial[b, blockIdx.x]
  if (k.compute_log_prob && p.logp_partial != nullptr) {
    __shared__ float red[SDE_THREADS / 32];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < SDE_THREADS / 32; ++i) t += red[i];
      p.logp_partial[b * gridDim.x + blockIdx.x] = t;
    }
  }
}

// One block: per-sample deterministic reduction of the block partials, normaliser, slot stores, step advance.
__global__ void sde_finalize_kernel(const SdeStepParams p, int nblk) {
  const int sidx = p.step_ptr ? *p.step_ptr : p.coef_index;
  const StepCoef k = p.coef_table[sidx];
  const int chw = p.C * p.H * p.W;
  if (k.compute_log_prob) {
    for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
      float lp = 0.f;
      if (k.dynamics != DYN_ODE) {
        float t = 0.f;
        for (int i = 0; i < nblk; ++i) t += p.logp_partial[b * nblk + i];
        lp = t / static_cast<float>(chw) - k.log_norm;
      }
      if (p.log_prob) p.log_prob[b] = lp;
      if (p.logp_traj && k.logp_slot >= 0) p.logp_traj[b * p.logp_batch_stride + k.logp_slot] = lp;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && p.step_ptr) *p.step_ptr = sidx + 1;
}

constexpr int SDE_MAX_BLOCKS = 64;

cudaError_t launch_sde_step(const SdeStepParams& p, cudaStream_t stream) {
  if (p.W % 4 != 0) return cudaErrorInvalidValue;
  const int quads = p.C * p.H * p.W / 4;
  int nblk = (quads + SDE_THREADS - 1) / SDE_THREADS;
  if (nblk > SDE_MAX_BLOCKS) nblk = SDE_MAX_BLOCKS;
  dim3 grid(nblk, p.B);
  sde_step_kernel<<<grid, SDE_THREADS, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sde_finalize_kernel<<<1, 32, 0, stream>>>(p, nblk);
  e = cudaGetLastError();
  return e;
}

}  // namespace ffb
