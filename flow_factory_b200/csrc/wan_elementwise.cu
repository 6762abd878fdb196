// HBM-bound pieces of the Wan2.1 T2V denoise step, 16-byte vector access, one warp per token row:
//   wan_patchify        : im2col of the Conv3d patch embedding, kernel = stride = (pt, ph, pw)      (transformer_wan.py:598, 668-669)
//   wan_mod_vectors     : scale_shift_table + temb (fp32) for every block in one launch               (transformer_wan.py:474-483)
//   ln_mod_f32          : FP32LayerNorm + fp32 modulate, or FP32LayerNorm with affine weights (norm2)  (transformer_wan.py:486, 493, 500)
//   gate_residual_f32   : hs = (hs.float() + y * gate).type_as(hs)                                    (transformer_wan.py:489, 503)
//   rms_rope_rows       : torch.nn.RMSNorm ACROSS heads + interleaved-pair RoPE in the tensor dtype   (transformer_wan.py:96-117)
//
// Validated on B200 in round 2 (tests/test_gpu_wan.py).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace ffb {

__device__ __forceinline__ void wan_unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 wan_pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// ------------------------------------------------------------------------------------------------ patch embedding im2col
// x fp16 [B, C, F, H, W] -> bf16 rows [(rep, b, f', h', w')][c, dt, dh, dw]  (the flattening order of the Conv3d weight [D, C, pt, ph, pw])
__global__ void wan_patchify_kernel(const __half* x, int B, int reps, int C, int F, int H, int W, int pt, int ph, int pw, bf16* out) {
  const int K = C * pt * ph * pw;
  const int fp = F / pt, hp = H / ph, wp = W / pw;
  const long tokens = static_cast<long>(B) * fp * hp * wp;
  const long total = tokens * K;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    long tok = i / K;
    const int dw = k % pw, dh = (k / pw) % ph, dt = (k / (pw * ph)) % pt, c = k / (pw * ph * pt);
    const int w = static_cast<int>(tok % wp); tok /= wp;
    const int h = static_cast<int>(tok % hp); tok /= hp;
    const int f = static_cast<int>(tok % fp);
    const int b = static_cast<int>(tok / fp);
    const float v = __half2float(x[(((static_cast<long>(b) * C + c) * F + f * pt + dt) * H + h * ph + dh) * W + w * pw + dw]);
    const bf16 o = __float2bfloat16_rn(v);                        // latents.to(transformer dtype)   (wan2_t2v.py:497)
    for (int r = 0; r < reps; ++r) out[r * total + i] = o;        // CFG: the same latents under the negative and the positive prompt
  }
}
cudaError_t launch_wan_patchify(const __half* x, int B, int reps, int C, int F, int H, int W, int pt, int ph, int pw, bf16* out,
                                cudaStream_t stream) {
  if (F % pt || H % ph || W % pw) return cudaErrorInvalidValue;
  const long total = static_cast<long>(B) * C * F * H * W;
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 32));
  wan_patchify_kernel<<<grid, 256, 0, stream>>>(x, B, reps, C, F, H, W, pt, ph, pw, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ modulation vectors
// mod[l][b][j] = float(table_l[j]) + float(temb6[b][j]),  j < 6 D ; tables: L pointers to bf16 [6 D] (the parameter in the module dtype)
__global__ void wan_mod_vectors_kernel(const bf16* const* tables, const bf16* temb6, float* mod, int L, int Bp, int n6) {
  const long total = static_cast<long>(L) * Bp * n6;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % n6);
    const int b = static_cast<int>((i / n6) % Bp);
    const int l = static_cast<int>(i / (static_cast<long>(n6) * Bp));
    mod[i] = __bfloat162float(tables[l][j]) + __bfloat162float(temb6[static_cast<long>(b) * n6 + j]);
  }
}
cudaError_t launch_wan_mod_vectors(const bf16* const* tables, const bf16* temb6, float* mod, int L, int Bp, int n6, cudaStream_t stream) {
  const long total = static_cast<long>(L) * Bp * n6;
  const int grid = static_cast<int>(std::min<long>((total + 255) / 256, 148L * 16));
  wan_mod_vectors_kernel<<<grid, 256, 0, stream>>>(tables, temb6, mod, L, Bp, n6);
  return cudaGetLastError();
}
// final modulation (transformer_wan.py:715-728): (scale_shift_table + temb.unsqueeze(1)) stays in the module dtype -> bf16 [Bp][2][D]
__global__ void wan_final_mod_kernel(const bf16* table, const bf16* temb, bf16* out, int Bp, int D) {
  const int total = Bp * 2 * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % D, j = (i / D) % 2, b = i / (2 * D);
    out[i] = __float2bfloat16_rn(__bfloat162float(table[j * D + n]) + __bfloat162float(temb[b * D + n]));
  }
}
cudaError_t launch_wan_final_mod(const bf16* table, const bf16* temb, bf16* out, int Bp, int D, cudaStream_t stream) {
  wan_final_mod_kernel<<<(Bp * 2 * D + 255) / 256, 256, 0, stream>>>(table, temb, out, Bp, D);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ FP32LayerNorm + modulate / affine
// mode 0: out = bf16( LN(x) * (1 + scale[b]) + shift[b] )   scale / shift fp32 vectors per batch row (mod_batch_stride apart)
// mode 1: out = bf16( LN(x) * weight + bias )                weight / bias bf16 [D]                  (norm2, elementwise_affine)
constexpr int WLN_MAXC = 12;
__global__ void __launch_bounds__(256, 3) wan_ln_kernel(const WanLnParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long rows = static_cast<long>(p.rows_per_batch) * p.num_batch;
  if (warp >= rows) return;
  const int b = warp / p.rows_per_batch;
  const int nchunk = p.D >> 3;
  const bf16* xr = p.x + static_cast<long>(warp) * p.D;
  uint4 raw[WLN_MAXC];
#pragma unroll
  for (int i = 0; i < WLN_MAXC; ++i) {
    const int c = lane + 32 * i;
    raw[i] = make_uint4(0, 0, 0, 0);
    if (c < nchunk) raw[i] = *reinterpret_cast<const uint4*>(xr + c * 8);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < WLN_MAXC; ++i)
    if (lane + 32 * i < nchunk) {
      float v[8];
      wan_unpack8(raw[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[e];
    }
  const float mean = warp_sum(sum) / static_cast<float>(p.D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < WLN_MAXC; ++i)
    if (lane + 32 * i < nchunk) {
      float v[8];
      wan_unpack8(raw[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[e] - mean; sq += d * d; }
    }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(p.D) + p.eps);
#pragma unroll
  for (int i = 0; i < WLN_MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunk) {
      float y[8], o[8];
      wan_unpack8(raw[i], y);
      if (p.mode == 0) {
        const float* sc = p.scale + static_cast<long>(b) * p.mod_batch_stride + c * 8;
        const float* sh = p.shift + static_cast<long>(b) * p.mod_batch_stride + c * 8;
        const float4 s0 = *reinterpret_cast<const float4*>(sc), s1 = *reinterpret_cast<const float4*>(sc + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(sh), h1 = *reinterpret_cast<const float4*>(sh + 4);
        const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float shv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __fadd_rn(__fmul_rn((y[e] - mean) * rstd, __fadd_rn(1.0f, scv[e])), shv[e]);
      } else {
        float wv[8], bv[8];
        wan_unpack8(__ldg(reinterpret_cast<const uint4*>(p.weight + c * 8)), wv);
        wan_unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias + c * 8)), bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __fadd_rn(__fmul_rn((y[e] - mean) * rstd, wv[e]), bv[e]);
      }
      *reinterpret_cast<uint4*>(p.out + static_cast<long>(warp) * p.D + c * 8) = wan_pack8(o);
    }
  }
}
cudaError_t launch_wan_ln(const WanLnParams& p, cudaStream_t stream) {
  if (p.D % 8 != 0 || p.D > WLN_MAXC * 32 * 8) return cudaErrorInvalidValue;
  const long rows = static_cast<long>(p.rows_per_batch) * p.num_batch;
  wan_ln_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ gated residual in fp32
// h[b, s, :] = bf16( float(h) + float(y) * gate[b, :] ),  gate fp32 (mod_batch_stride apart)
__global__ void wan_gate_residual_kernel(bf16* h, const bf16* y, const float* gate, long gate_batch_stride, long rows_per_batch, int D, long total8) {
  const int oct = D >> 3;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % oct);
    const long row = i / oct;
    const long b = row / rows_per_batch;
    float hv[8], yv[8];
    wan_unpack8(*reinterpret_cast<const uint4*>(h + i * 8), hv);
    wan_unpack8(*reinterpret_cast<const uint4*>(y + i * 8), yv);
    const float* g = gate + b * gate_batch_stride + o * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) hv[e] = __fadd_rn(hv[e], __fmul_rn(yv[e], gv[e]));
    *reinterpret_cast<uint4*>(h + i * 8) = wan_pack8(hv);
  }
}
cudaError_t launch_wan_gate_residual(bf16* h, const bf16* y, const float* gate, long gate_batch_stride, int num_batch, long rows_per_batch,
                                     int D, cudaStream_t stream) {
  if (D % 8 != 0) return cudaErrorInvalidValue;
  const long total8 = static_cast<long>(num_batch) * rows_per_batch * (D / 8);
  const int grid = static_cast<int>(std::min<long>((total8 + 255) / 256, 148L * 32));
  wan_gate_residual_kernel<<<grid, 256, 0, stream>>>(h, y, gate, gate_batch_stride, rows_per_batch, D, total8);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ RMSNorm across heads (+ RoPE)
// In place on a column block of width D inside rows of pitch ld:  x = bf16( x * rsqrt(mean_D x^2 + eps) * w )   (torch.nn.RMSNorm: fp32
// inside, one rounding), then - if tables are given - per head of 128 the interleaved-pair rotation evaluated in bf16 like the
// reference's tensor arithmetic (x1 * cos - x2 * sin, x1 * sin + x2 * cos: every product and the sum are bf16 tensors).
// cos / sin: fp32 [tokens_per_batch][128] holding the (bf16-rounded, pairwise repeated) table values.
// out_scale != 1: the result is multiplied by it before the final bf16 store (the engine's key pre-scaling, GemmParams::k_scale's twin).
constexpr int WRR_MAXC = 12;
__global__ void __launch_bounds__(256, 3) wan_rms_rope_kernel(bf16* x, long rows, int rows_per_batch, int ld, int D, const bf16* weight,
                                                              float eps, const float* cos_t, const float* sin_t, float out_scale) {
  const long warp = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nchunk = D >> 3;
  bf16* xr = x + warp * ld;
  uint4 raw[WRR_MAXC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < WRR_MAXC; ++i) {
    const int c = lane + 32 * i;
    raw[i] = make_uint4(0, 0, 0, 0);
    if (c < nchunk) {
      raw[i] = *reinterpret_cast<const uint4*>(xr + c * 8);
      float v[8];
      wan_unpack8(raw[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(v[e], v[e], ss);
    }
  }
  const float rs = rsqrtf(warp_sum(ss) / static_cast<float>(D) + eps);
  const long tok = warp % rows_per_batch;
#pragma unroll
  for (int i = 0; i < WRR_MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunk) {
      float v[8], wv[8];
      wan_unpack8(raw[i], v);
      wan_unpack8(__ldg(reinterpret_cast<const uint4*>(weight + c * 8)), wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = bf16_round(__fmul_rn(__fmul_rn(v[e], rs), wv[e]));
      if (cos_t != nullptr) {
        const int hc = (c * 8) & 127;                               // column inside the 128-wide head
        const float* cs = cos_t + tok * 128 + hc;
        const float* sn = sin_t + tok * 128 + hc;
        const float4 c0 = *reinterpret_cast<const float4*>(cs), c1 = *reinterpret_cast<const float4*>(cs + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(sn), s1 = *reinterpret_cast<const float4*>(sn + 4);
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float x1 = v[e], x2 = v[e + 1];
          // cos[..., 0::2] / sin[..., 1::2] (transformer_wan.py:106-107): the even entry's cos, the odd entry's sin
          const float a = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x1, cc[e])), -bf16_round(__fmul_rn(x2, sv[e + 1]))));
          const float b2 = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x1, sv[e + 1])), bf16_round(__fmul_rn(x2, cc[e]))));
          v[e] = a; v[e + 1] = b2;
        }
      }
      if (out_scale != 1.0f) {                                      // keys: softmax_scale * log2(e) folded in (softmax.cuh, pre-scaled keys)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= out_scale;
      }
      *reinterpret_cast<uint4*>(xr + c * 8) = wan_pack8(v);
    }
  }
}
cudaError_t launch_wan_rms_rope(bf16* x, long rows, int rows_per_batch, int ld, int D, const bf16* weight, float eps, const float* cos_t,
                                const float* sin_t, cudaStream_t stream, float out_scale) {
  if (D % 8 != 0 || D > WRR_MAXC * 32 * 8 || ld % 8 != 0 || (cos_t != nullptr && D % 128 != 0)) return cudaErrorInvalidValue;
  wan_rms_rope_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, rows, rows_per_batch, ld, D, weight, eps, cos_t, sin_t, out_scale);
  return cudaGetLastError();
}

}  // namespace ffb
