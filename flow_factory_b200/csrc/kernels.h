// Internal launch interfaces shared by the .cu files and the C-ABI layer (capi.cu / engine.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace ffb {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ GEMM
enum GemmEpilogue : int {
  EPI_BIAS = 0,               // out = bf16(acc + bias)
  EPI_BIAS_GELU = 1,          // out = bf16(gelu_tanh(bf16(acc + bias)))
  EPI_GATE_RESIDUAL = 2,      // out = bf16(out + bf16(gate[b,:] * bf16(acc + bias)))      (in place on the residual stream)
  EPI_QKV_RMSNORM = 3,        // fused q|k|v projection: per-head RMSNorm on the q and k column blocks
  EPI_BIAS_ADD_ROWTABLE = 4,  // out = bf16(bf16(acc + bias) + table[row, :])             (patch-embed + pos-embed)
  EPI_QKV_RMSNORM_ROPE128 = 5,  // fused q|k|v projection, head_dim 128: RMSNorm + interleaved-pair RoPE on the q and k heads (FLUX.1)
};

struct GemmParams {
  CUtensorMap tmA;  // 3-D {K, rows_per_batch, num_batch}, box {64, 128, 1}, SWIZZLE_128B
  CUtensorMap tmB;  // 2-D {K, N},                          box {64, bn/2},  SWIZZLE_128B (one half per CTA of the cluster pair)
  int rows_per_batch, num_batch, N, K;
  int tiles_m_per_batch;
  int bn;   // N tile: 256 / 128 / 64 (must divide N)
  int band; // m-pairs per rasterisation band (L2 residency of the A rows)
  int epi;
  const bf16* bias;  // [N] or null
  bf16* out;         // out[b*out_batch_stride + (out_row_offset + row)*ldo + n]
  long out_batch_stride;
  int out_row_offset;
  int ldo;
  const bf16* gate;  // EPI_GATE_RESIDUAL: gate[b*gate_batch_stride + n]
  long gate_batch_stride;
  const bf16* norm_q;  // EPI_QKV_RMSNORM: [64] weights for the q / k column blocks
  const bf16* norm_k;
  int qk_dim;        // width of each of the q|k|v column blocks (= inner_dim)
  float eps;
  const float* row_table;  // EPI_BIAS_ADD_ROWTABLE: fp32 [rows_per_batch, N]
  const float* rope_cos;   // EPI_QKV_RMSNORM_ROPE128: fp32 [tokens, 128] (values repeated pairwise, FluxPosEmbed), row = rope_row_offset + row
  const float* rope_sin;
  int rope_row_offset;
  float k_scale;           // EPI_QKV_RMSNORM[_ROPE128]: factor folded into the k heads before their bf16 store (0 = 1): softmax_scale * log2(e)
  int rms_round_first;     // 0: torch.nn.RMSNorm bf16((x*rs)*w) (FLUX.1) ; 1: diffusers RMSNorm bf16(bf16(x*rs)*w) (Qwen-Image)
  int epi_split;           // set by launch_gemm: 1 = one epilogue warp group takes the whole tile, 2 = the two groups take half the columns each
};

int gemm_pick_bn(int N);
cudaError_t launch_gemm(const GemmParams& p, int num_sms, cudaStream_t stream);

// ------------------------------------------------------------------ attention
struct AttnParams {
  CUtensorMap tmQKV;  // 3-D {3*D, S, B}, box {64, 128, 1}, SWIZZLE_128B over the joint qkv buffer [B, S, 3D] (Q sub-tiles)
  CUtensorMap tmKV;   // same tensor, box {64, 64, 1} (K / V tiles)
  int seq_len;        // S (q and kv length)
  int num_heads;      // H
  int inner_dim;      // D = 64*H
  int batch;
  bf16* out;          // [B, S, out_row_stride >= D]
  long out_batch_stride;
  int out_row_stride; // elements between consecutive token rows of `out` (D, or wider when the output is a column block)
  float scale_log2;   // (1/sqrt(head_dim)) * log2(e)
  int k_prescaled;    // 1: the producer of K folded scale_log2 into the keys (GemmParams::k_scale): scores are base-2 exponents as they are
  // optional key-padding mask (head_dim 128 kernel only): keys [kv_mask_lo[b], kv_mask_hi) of batch b are excluded - the padded tail
  // of the text rows of a joint [text ; image] sequence (Qwen-Image attention_mask, transformer_qwenimage.py:952-958); null = none
  const int* kv_mask_lo;
  int kv_mask_hi;
  // cross-attention (head_dim 128 kernel, launch_attention_d128_cross only): tmQKV addresses the query tensor [B, seq_len, >= D]
  // (head h at column q_col + 128 h), tmKV the key / value tensor [B, kv_len, ...] (columns k_col + 128 h / v_col + 128 h)
  int kv_len, q_col, k_col, v_col;
  // optional (head_dim 64 kernel, pre-scaled keys): the per-head RMSNorm weights [64] that PRODUCED q and k (up to two sets: image and text
  // stream of a joint attention; unused slots null).  RMS-normed heads have ||q|| <= 8 max|w_q| and ||k'|| <= 8 max|w_k| scale_log2, so
  // |exponent| <= 64 max|w_q| max|w_k| scale_log2 (Cauchy-Schwarz): when that proves |exponent| <= 126 for the whole launch, the kernel
  // skips the per-tile range guard of its polynomial exp2 slots.  All null: no proof, the guard stays (op-level entry, hooks).
  const bf16* bound_wq[2];
  const bf16* bound_wk[2];
};
cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream);        // head_dim 64  (SD3.x)
cudaError_t launch_attention_d128(const AttnParams& p, cudaStream_t stream);   // head_dim 128 (FLUX.1)
cudaError_t launch_attention_d128_cross(const AttnParams& p, cudaStream_t stream);   // head_dim 128, separate q / kv tensors and lengths (Wan)

// ------------------------------------------------------------------ per-step scalars
enum Dynamics : int { DYN_FLOW_SDE = 0, DYN_DANCE_SDE = 1, DYN_CPS = 2, DYN_ODE = 3 };

// Per-step scalars, computed on the host in fp32 exactly as the reference's (B,1,1,1) tensors are.
struct StepCoef {
  float t_model;      // fp16-rounded timestep fed to the sinusoid (sd3_5.py:394)
  float sigma, sigma_prev, dt, noise_level, std_dev_t;
  float c_x;          // Flow-SDE: 1 + std^2/(2 sigma) * dt
  float c_v;          // Flow-SDE: 1 + std^2 (1-sigma)/(2 sigma)
  float noise_scale;  // std_dev_t * sqrt(-dt)   (CPS: std_dev_t)
  float two_var;      // 2 * noise_scale^2 (per-element divisor of the Gaussian log-density)
  float log_norm;     // log(noise_scale) + log(sqrt(2 pi))   (0 for CPS)
  float cps_a, cps_b; // CPS: (1 - sigma_prev), sqrt(sigma_prev^2 - std^2)
  int dynamics;
  int compute_log_prob;
  int store_slot;     // index into all_latents for the *result* of this step, or -1
  int logp_slot;      // index into log_probs for this step, or -1
};

// ------------------------------------------------------------------ elementwise / small ops
// LayerNorm (no affine, eps) + adaLN modulate: out = bf16( LN(x) * bf16(1 + scale[b]) + shift[b] ), up to two outputs.
struct LnModParams {
  const bf16* x;    // [num_batch * rows_per_batch, D]
  int rows_per_batch, num_batch, D;
  float eps;
  const bf16* shift1; const bf16* scale1; bf16* out1;  // vectors: ptr + b*mod_batch_stride
  const bf16* shift2; const bf16* scale2; bf16* out2;  // optional second modulation (SD35AdaLayerNormZeroX)
  long mod_batch_stride;
  long x_batch_stride;    // elements between batches of x (0 = rows_per_batch * D): row ranges of a joint [B, S, D] buffer
  long out_batch_stride;  // same for out1 / out2
};
cudaError_t launch_ln_modulate(const LnModParams& p, cudaStream_t stream);

// out[b, n] = bf16( bf16(sum_k in'[b,k] * W[n,k] + bias[n]) (+ addend[b,n]) ),  in' = silu(in) if silu_input
struct SmallLinearParams {
  const bf16* in; int batch; int K; long in_stride;
  const bf16* W; const bf16* bias; int N;
  bf16* out; long out_stride;
  const bf16* addend; long addend_stride;  // optional
  int silu_input;
  const bf16* addend2;   // optional second bf16 addend, added after the first with its own rounding (same stride)
};
cudaError_t launch_small_linear(const SmallLinearParams& p, cudaStream_t stream);

// sinusoidal timestep projection (embeddings.py:26-77, flip_sin_to_cos, shift 0) -> bf16 [batch, 256]
cudaError_t launch_timestep_proj(const StepCoef* table, const int* step_ptr, int index, int batch, bf16* out, cudaStream_t stream,
                                 float post_scale = 1.0f);   // post_scale: Timesteps(scale=...) of Qwen-Image (1000)

// im2col for the 2x2/stride-2 patch-embed conv: fp16 latents [B,C,H,W] -> bf16 [Bp*Ni, C*p*p] (Bp = B*reps)
cudaError_t launch_patchify(const void* x, int storage, int B, int reps, int C, int H, int W, int patch, bf16* out, cudaStream_t stream);

// cast helpers
cudaError_t launch_cast_f32_to_bf16(const float* in, bf16* out, long n, cudaStream_t stream);
cudaError_t launch_cast_f16_to_bf16(const __half* in, bf16* out, long n, cudaStream_t stream);
// diffusers RMSNorm over rows of width K (text-stream input norm of Qwen-Image)
cudaError_t launch_rms_norm_rows(const bf16* x, const bf16* weight, bf16* out, long rows, int K, float eps, cudaStream_t stream);
// Qwen-Image true CFG + per-token norm rescale: v bf16 [2B, Ni, 64] (negative half first) -> out bf16 [B, Ni, 64]
cudaError_t launch_cfg_norm_rescale(const bf16* v, bf16* out, long tokens_per_half, float guidance, cudaStream_t stream);

// ------------------------------------------------------------------ VAE decode: NHWC bf16 convolutions
enum ConvEpilogue : int {
  EPI_CONV_BIAS = 0,      // out = bf16(acc + bias)                                  NHWC
  EPI_CONV_RESIDUAL = 1,  // out = bf16(bf16(acc + bias) + residual)                 NHWC   (ResnetBlock2D: x + h, resnet.py:375)
  EPI_CONV_NCHW = 2,      // first n_store channels of bf16(acc + bias) as planar [B, n_store, H, W]   (conv_out -> image)
  EPI_CONV_F32 = 3,       // out_f32 = acc * out_scale, row pitch ldo_f32                              (attention scores of the mid block)
};

// Stride-1 "same" convolution (3x3 pad 1, or 1x1 == plain GEMM over pixel rows) as an implicit GEMM on the tensor cores:
//   out[b, h, w, n] = sum_{tap, c} x[b, h + dy(tap), w + dx(tap), c] * Wp[n, tap * kc * 64 + c]
// The A operand of tap (dy, dx) is the SAME 4-D TMA box of x moved by (dy, dx): out-of-range pixels (the padding) and channels beyond
// Cin are zero-filled by the TMA unit, so there is no im2col buffer and no halo logic in the kernel.
struct ConvParams {
  CUtensorMap tmA;  // 4-D {Cin, W, H, B} over NHWC x (pixel pitch lda), box {64, tw, th, 1}, SWIZZLE_128B
  CUtensorMap tmB;  // 2-D {taps * kc * 64, N} packed weights [N][tap][Cin padded to 64], box {64, bn/2}
  int B, H, W;
  int tw_log2;      // pixel tile = th x tw with tw = 1 << tw_log2, th = 128 / tw
  int tiles_w, tiles_h;
  int kc, taps;     // 64-channel blocks per tap; 9 or 1
  int N, bn, band;
  int epi;
  const bf16* bias;      // [N] or null
  bf16* out; int ldo;    // NHWC (pixel pitch ldo) or planar (EPI_CONV_NCHW)
  int n_store;           // output channels actually stored (<= N; multiple of 8 for the NHWC modes)
  const bf16* residual; int ldr;
  float* out_f32; long ldo_f32; float out_scale;
};
cudaError_t launch_conv(const ConvParams& p, int num_sms, cudaStream_t stream);

// GroupNorm over NHWC bf16 [B, P pixels, C]: fp32 statistics (CUDA autocast keeps group_norm in fp32), affine, optional SiLU, bf16 out.
// stats: double [B][C][2] (sum, sum of squares), zeroed by the caller; one pass accumulates, the apply pass finishes per group.
struct GroupNormParams {
  const bf16* x; bf16* out;
  int B; long P; int C, groups;
  float eps; int silu;
  const bf16* gamma; const bf16* beta;
  double* stats;
};
cudaError_t launch_group_norm_stats(const GroupNormParams& p, cudaStream_t stream);
cudaError_t launch_group_norm_apply(const GroupNormParams& p, cudaStream_t stream);
// nearest-neighbour 2x upsample, NHWC bf16 [B, H, W, C] -> [B, 2H, 2W, C]   (Upsample2D, upsampling.py)
cudaError_t launch_upsample2x_nhwc(const bf16* x, bf16* out, int B, int H, int W, int C, cudaStream_t stream);
// decode_latents prologue (sd3_5.py:166-167): fp16 NCHW latents -> bf16 NHWC z = bf16(bf16(bf16(x) / scaling) + shift), channels padded to Cp
cudaError_t launch_vae_prep_latents(const __half* x, bf16* out, int B, int C, int H, int W, int Cp, float scaling, float shift, cudaStream_t stream);
// row softmax of fp32 scores [rows][pitch] (first n columns) -> bf16 probabilities written IN PLACE at the start of each row
cudaError_t launch_softmax_rows_inplace(float* scores, long rows, int n, long pitch, cudaStream_t stream);

// ------------------------------------------------------------------ Wan2.1 T2V: elementwise pieces (wan_elementwise.cu)
cudaError_t launch_wan_patchify(const __half* x, int B, int reps, int C, int F, int H, int W, int pt, int ph, int pw, bf16* out, cudaStream_t stream);
cudaError_t launch_wan_mod_vectors(const bf16* const* tables, const bf16* temb6, float* mod, int L, int Bp, int n6, cudaStream_t stream);
cudaError_t launch_wan_final_mod(const bf16* table, const bf16* temb, bf16* out, int Bp, int D, cudaStream_t stream);
struct WanLnParams {
  const bf16* x; bf16* out;          // [num_batch * rows_per_batch, D], dense
  int rows_per_batch, num_batch, D;
  float eps;
  int mode;                          // 0: fp32 (1 + scale[b]) / shift[b] modulation   1: affine weight / bias (bf16 [D])
  const float* scale; const float* shift; long mod_batch_stride;
  const bf16* weight; const bf16* bias;
};
cudaError_t launch_wan_ln(const WanLnParams& p, cudaStream_t stream);
cudaError_t launch_wan_gate_residual(bf16* h, const bf16* y, const float* gate, long gate_batch_stride, int num_batch, long rows_per_batch,
                                     int D, cudaStream_t stream);
cudaError_t launch_wan_rms_rope(bf16* x, long rows, int rows_per_batch, int ld, int D, const bf16* weight, float eps, const float* cos_t,
                                const float* sin_t, cudaStream_t stream, float out_scale = 1.0f);

// latent storage dtypes (element type of the x / next / trajectory buffers)
enum { LAT_F16 = 0, LAT_BF16 = 1, LAT_F32 = 2 };
inline int lat_elem_bytes(int storage) { return storage == LAT_F32 ? 4 : 2; }

// ------------------------------------------------------------------ fused Euler/SDE step + log-prob (K14)

struct SdeStepParams {
  const bf16* v_tokens;   // [Bp*Ni, p*p*C] proj_out rows (token-major, "nhwpqc"), uncond half first when cfg
  const bf16* v_direct;   // alternatively an NCHW bf16 noise_pred [B,C,H,W] (op-level entry); one of the two
  int B, C, H, W, patch;
  int cfg;                // 1: v = vu + g*(vc - vu) in bf16 steps (sd3_5.py:431-433)
  float guidance;
  const void* x;          // [B,C,H,W] current latents in the storage dtype (`storage`)
  const float* noise;     // fp32 N(0,1) [(steps,) B,C,H,W] or null -> in-kernel Philox4x32-10 + Box-Muller
  long noise_step_stride; // elements between steps (0 for a single step)
  unsigned long long seed;
  const void* next_given; // teacher-forced next latents (storage dtype) or null
  void* x_next;           // [B,C,H,W] next_latents rounded to the storage dtype (fp16: +-65504 clamp); may alias x
  void* traj;             // all_latents base [B, n_slots, C,H,W] (storage dtype) or null; slot = coef.store_slot
  long traj_batch_stride; // elements
  float* mean_out;        // optional fp32 next_latents_mean or null
  bf16* v_out;            // optional bf16 [B,C,H,W] noise_pred after CFG or null
  float* logp_partial;    // [B, SDE_MAX_BLOCKS] scratch
  float* log_prob;        // direct [B] output or null
  float* logp_traj;       // log_probs base [B, n_logp_slots] or null; slot = coef.logp_slot
  int logp_batch_stride;
  int* overflow_flag;     // sticky fp16-overflow flag (cast_latents, abc.py:172-182)
  const StepCoef* coef_table;
  int* step_ptr;          // device step counter (entry = *step_ptr, incremented by the finalize kernel) or null
  int coef_index;         // used when step_ptr == null
  int storage;            // latent_storage_dtype (FF/hparams/training_args.py:245-252): LAT_F16 (default) | LAT_BF16 | LAT_F32
};
cudaError_t launch_sde_step(const SdeStepParams& p, cudaStream_t stream);

// proj_out GEMM whose epilogue is the whole denoise step (final_step.cu); `sde.v_tokens/v_direct` are unused.
struct FinalStepParams {
  CUtensorMap tmA;   // 3-D {K, Ni, Bp} over the norm_out-modulated hidden states, box {64, 128, 1}
  CUtensorMap tmW;   // 2-D {K, 64} proj_out weight, box {64, 64}
  int K;
  const bf16* bias;  // [64]
  SdeStepParams sde;
  unsigned int* done_counter;  // device word, zero between launches
};
cudaError_t launch_final_step(const FinalStepParams& p, cudaStream_t stream);

}  // namespace ffb
