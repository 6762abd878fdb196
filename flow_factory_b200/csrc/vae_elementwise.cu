// HBM-bound pieces of the VAE decode, all on NHWC bf16 activations with 16-byte vector access:
//   group_norm_stats / apply : nn.GroupNorm (+ SiLU) of ResnetBlock2D / the mid-block attention / conv_norm_out
//                              (DF/models/resnet.py:345-366, attention_processor.py group_norm, autoencoders/vae.py:309-311);
//                              fp32 statistics and affine - CUDA autocast runs group_norm in fp32 - rounded to bf16 once, at the
//                              input of the next convolution
//   upsample2x_nhwc          : F.interpolate(scale_factor=2, mode="nearest")   (DF/models/upsampling.py Upsample2D.forward)
//   vae_prep_latents         : latents.to(vae.dtype) / scaling_factor + shift_factor   (FF/models/stable_diffusion/sd3_5.py:166-167)
//   softmax_rows_inplace     : softmax of the single-head mid-block attention scores (F.scaled_dot_product_attention, head_dim = C)
//
// STATUS: written after round 1's GPU budget was spent - compiled for sm_100a, NOT yet run on a GPU (see vae_conv.cu).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace ffb {

__device__ __forceinline__ void vae_unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

constexpr int GN_PIXELS_PER_BLOCK = 2048;

// thread -> (channel octet o = tid % oct, pixel lane pl = tid / oct); blockDim.x = oct * lanes (host-chosen, <= 1024)
__global__ void group_norm_stats_kernel(const GroupNormParams p) {
  extern __shared__ float gn_sm[];   // [2][C]
  const int oct = p.C >> 3, lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct, pl = threadIdx.x / oct;
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < 2 * p.C; c += blockDim.x) gn_sm[c] = 0.f;
  __syncthreads();
  const long p0 = static_cast<long>(blockIdx.x) * GN_PIXELS_PER_BLOCK;
  const long p1 = min(p.P, p0 + GN_PIXELS_PER_BLOCK);
  const bf16* xb = p.x + static_cast<long>(b) * p.P * p.C + o * 8;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  for (long pix = p0 + pl; pix < p1; pix += lanes) {
    float v[8];
    vae_unpack8(*reinterpret_cast<const uint4*>(xb + pix * p.C), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += v[e]; q[e] = fmaf(v[e], v[e], q[e]); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&gn_sm[o * 8 + e], s[e]);
    atomicAdd(&gn_sm[p.C + o * 8 + e], q[e]);
  }
  __syncthreads();
  double* st = p.stats + static_cast<long>(b) * p.C * 2;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    atomicAdd(&st[2 * c], static_cast<double>(gn_sm[c]));
    atomicAdd(&st[2 * c + 1], static_cast<double>(gn_sm[p.C + c]));
  }
}

__global__ void group_norm_apply_kernel(const GroupNormParams p) {
  extern __shared__ float gn_sm[];   // a[C] | d[C]:  y = x * a + d
  const int oct = p.C >> 3, lanes = blockDim.x / oct;
  const int o = threadIdx.x % oct, pl = threadIdx.x / oct;
  const int b = blockIdx.y;
  const int cpg = p.C / p.groups;
  const double* st = p.stats + static_cast<long>(b) * p.C * 2;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    double s = 0.0, q = 0.0;
    for (int j = g0; j < g0 + cpg; ++j) { s += st[2 * j]; q += st[2 * j + 1]; }
    const double n = static_cast<double>(p.P) * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
    const float a = rstd * __bfloat162float(p.gamma[c]);
    gn_sm[c] = a;
    gn_sm[p.C + c] = __bfloat162float(p.beta[c]) - static_cast<float>(mean) * a;
  }
  __syncthreads();
  float a[8], d[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { a[e] = gn_sm[o * 8 + e]; d[e] = gn_sm[p.C + o * 8 + e]; }
  const long p0 = static_cast<long>(blockIdx.x) * GN_PIXELS_PER_BLOCK;
  const long p1 = min(p.P, p0 + GN_PIXELS_PER_BLOCK);
  const long base = static_cast<long>(b) * p.P * p.C + o * 8;
  for (long pix = p0 + pl; pix < p1; pix += lanes) {
    float v[8];
    vae_unpack8(*reinterpret_cast<const uint4*>(p.x + base + pix * p.C), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = fmaf(v[e], a[e], d[e]);
      if (p.silu) y = __fdividef(y, 1.0f + __expf(-y));
      v[e] = y;
    }
    uint4 out;
    out.x = pack_bf16x2(v[0], v[1]); out.y = pack_bf16x2(v[2], v[3]);
    out.z = pack_bf16x2(v[4], v[5]); out.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p.out + base + pix * p.C) = out;
  }
}

static bool gn_geometry(const GroupNormParams& p, int& threads) {
  if (p.C % 8 != 0 || p.groups <= 0 || p.C % p.groups != 0 || p.C > 4096) return false;
  const int oct = p.C / 8;
  if (oct > 512) return false;
  int lanes = 256 / oct;
  if (lanes < 1) lanes = 1;
  threads = oct * lanes;
  return true;
}
cudaError_t launch_group_norm_stats(const GroupNormParams& p, cudaStream_t stream) {
  int threads;
  if (!gn_geometry(p, threads)) return cudaErrorInvalidValue;
  const dim3 grid(static_cast<unsigned>((p.P + GN_PIXELS_PER_BLOCK - 1) / GN_PIXELS_PER_BLOCK), p.B);
  group_norm_stats_kernel<<<grid, threads, 2 * p.C * sizeof(float), stream>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_group_norm_apply(const GroupNormParams& p, cudaStream_t stream) {
  int threads;
  if (!gn_geometry(p, threads)) return cudaErrorInvalidValue;
  const dim3 grid(static_cast<unsigned>((p.P + GN_PIXELS_PER_BLOCK - 1) / GN_PIXELS_PER_BLOCK), p.B);
  group_norm_apply_kernel<<<grid, threads, 2 * p.C * sizeof(float), stream>>>(p);
  return cudaGetLastError();
}

// one thread per (input pixel, channel octet): one 16-byte load, four 16-byte stores
__global__ void upsample2x_nhwc_kernel(const bf16* x, bf16* out, int B, int H, int W, int C) {
  const int oct = C >> 3;
  const long total = static_cast<long>(B) * H * W * oct;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % oct);
    long pix = i / oct;
    const int w = static_cast<int>(pix % W); pix /= W;
    const int h = static_cast<int>(pix % H);
    const int b = static_cast<int>(pix / H);
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((static_cast<long>(b) * H + h) * W + w) * C + o * 8);
    bf16* ob = out + ((static_cast<long>(b) * 2 * H + 2 * h) * (2 * W) + 2 * w) * C + o * 8;
    const long row = static_cast<long>(2 * W) * C;
    *reinterpret_cast<uint4*>(ob) = v;
    *reinterpret_cast<uint4*>(ob + C) = v;
    *reinterpret_cast<uint4*>(ob + row) = v;
    *reinterpret_cast<uint4*>(ob + row + C) = v;
  }
}
cudaError_t launch_upsample2x_nhwc(const bf16* x, bf16* out, int B, int H, int W, int C, cudaStream_t stream) {
  if (C % 8 != 0) return cudaErrorInvalidValue;
  const long total = static_cast<long>(B) * H * W * (C / 8);
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 32));
  upsample2x_nhwc_kernel<<<grid, 256, 0, stream>>>(x, out, B, H, W, C);
  return cudaGetLastError();
}

// one thread per pixel: planar fp16 reads are coalesced across the warp for every channel, the NHWC row is written by its owner
__global__ void vae_prep_latents_kernel(const __half* x, bf16* out, int B, int C, int H, int W, int Cp, float scaling, float shift) {
  const long plane = static_cast<long>(H) * W;
  const long total = static_cast<long>(B) * plane;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = i / plane, pix = i - b * plane;
    bf16* o = out + i * Cp;
    for (int c = 0; c < Cp; ++c) {
      float v = 0.f;
      if (c < C) {
        v = bf16_round(__half2float(x[(b * C + c) * plane + pix]));   // latents.to(vae.dtype)
        v = bf16_round(__fdiv_rn(v, scaling));                        // latents / scaling_factor  (a bf16 tensor)
        v = bf16_round(__fadd_rn(v, shift));                          // ... + shift_factor
      }
      o[c] = __float2bfloat16_rn(v);
    }
  }
}
cudaError_t launch_vae_prep_latents(const __half* x, bf16* out, int B, int C, int H, int W, int Cp, float scaling, float shift,
                                    cudaStream_t stream) {
  if (Cp < C || Cp % 8 != 0) return cudaErrorInvalidValue;
  const long total = static_cast<long>(B) * H * W;
  const int grid = static_cast<int>(std::min<long>((total + 127) / 128, 148L * 16));
  vae_prep_latents_kernel<<<grid, 128, 0, stream>>>(x, out, B, C, H, W, Cp, scaling, shift);
  return cudaGetLastError();
}

// One block per row.  The row (n <= 16384 fp32 scores) is staged in shared memory, so the bf16 probabilities can overwrite the start of
// the same row: P[row][j] lives at bf16 offset 2 * row * pitch + j, i.e. a bf16 matrix with row pitch 2 * pitch.
constexpr int SMR_THREADS = 512;
__global__ void __launch_bounds__(SMR_THREADS) softmax_rows_inplace_kernel(float* scores, int n, long pitch) {
  extern __shared__ float smr_row[];
  __shared__ float red[SMR_THREADS / 32];
  float* row = scores + static_cast<long>(blockIdx.x) * pitch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < n; j += SMR_THREADS) {
    const float v = row[j];
    smr_row[j] = v;
    m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < SMR_THREADS / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += SMR_THREADS) {
    const float e = exp2f((smr_row[j] - m) * 1.4426950408889634f);
    smr_row[j] = e;
    s += e;
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < SMR_THREADS / 32; ++w) s += red[w];
  const float inv = 1.0f / s;
  bf16* prow = reinterpret_cast<bf16*>(row);
  for (int j = threadIdx.x; j < n; j += SMR_THREADS) prow[j] = __float2bfloat16_rn(smr_row[j] * inv);
}
cudaError_t launch_softmax_rows_inplace(float* scores, long rows, int n, long pitch, cudaStream_t stream) {
  if (n <= 0 || n > 16384 || rows <= 0 || rows > 0x7FFFFFFFL) return cudaErrorInvalidValue;
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] { return cudaFuncSetAttribute(softmax_rows_inplace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4); });
    if (e != cudaSuccess) return e;
  }
  softmax_rows_inplace_kernel<<<static_cast<unsigned>(rows), SMR_THREADS, static_cast<size_t>(n) * sizeof(float), stream>>>(scores, n, pitch);
  return cudaGetLastError();
}

}  // namespace ffb
