/* ffb200 - C ABI of the B200-native rollout engine for Flow-Factory's SD3.5 path.
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  Every entry point returns 0 on success,
 * a positive cudaError_t, or a negative engine code; ffb200_last_error() returns the message.
 * All device pointers are borrowed for the duration of the call and used on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream).
 *
 * Each function names the reference interface it replaces (paths relative to /root/reference;
 * FF = src/flow_factory, DF = diffusers/src/diffusers).  The reference is pure Python, so the
 * "FFI" a maintainer adds is a ctypes binding - see INTEGRATION.md and flow_factory_b200/_lib.py.
 */
#ifndef FFB200_H
#define FFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFB200_ABI_VERSION 3   /* 2: + ffb200_attention_scaled, ffb200_plan_set_latent_dtype, ffb200_sde_step_ex; 3: + ffb200_attention_normed (additions only) */

/* ---------------------------------------------------------------- errors */
const char* ffb200_last_error(void);
/* Reads (and clears) the device-side error word written by a kernel before it trapped. */
int ffb200_device_error(unsigned int out[4]);
int ffb200_abi_version(void);
/* Developer aid: per-mbarrier-tag wait cycles of CTA 0 (out[tag], counts at out[128+tag]); all zero unless the library was built
 * with -DFFB_PROFILE (tools/prof_waits.py). Reading clears the counters. */
int ffb200_debug_read_prof(unsigned long long* out, int n);

/* ---------------------------------------------------------------- model description
 * DF/models/transformers/transformer_sd3.py:117-141 (register_to_config arguments). head_dim must be 64. */
typedef struct ffb200_model_config {
  int num_layers;
  int num_heads;           /* inner_dim D = 64 * num_heads */
  int patch_size;          /* 2 */
  int in_channels;         /* 16 (== out_channels) */
  int joint_attention_dim; /* 4096 */
  int pooled_projection_dim; /* 2048 */
  int pos_embed_max_size;  /* 384 */
  int num_dual_layers;     /* layers [0, num_dual_layers) carry attn2 (SD3.5: 13) */
} ffb200_model_config;

/* Per-layer weights, bf16, row-major [out_features, in_features] exactly as nn.Linear stores them.
 * q|k|v projections are concatenated along out_features (host side: torch.cat at load / refresh time). */
typedef struct ffb200_layer_weights {
  const void *qkv_w, *qkv_b, *norm_q, *norm_k;             /* attn.to_{q,k,v}, attn.norm_{q,k}            */
  const void *add_qkv_w, *add_qkv_b, *norm_added_q, *norm_added_k; /* attn.add_{q,k,v}_proj, norm_added_* */
  const void *out_w, *out_b;                               /* attn.to_out.0                               */
  const void *add_out_w, *add_out_b;                       /* attn.to_add_out (NULL on the last layer)    */
  const void *qkv2_w, *qkv2_b, *norm_q2, *norm_k2, *out2_w, *out2_b; /* attn2.* (NULL if not dual)        */
  const void *ff1_w, *ff1_b, *ff2_w, *ff2_b;               /* ff.net.0.proj, ff.net.2                     */
  const void *cff1_w, *cff1_b, *cff2_w, *cff2_b;           /* ff_context.* (NULL on the last layer)       */
} ffb200_layer_weights;

typedef struct ffb200_weights {
  const void *pe_w, *pe_b;      /* pos_embed.proj as [D, C*p*p], [D]                                       */
  const float* pos_embed;       /* fp32 [pos_embed_max_size^2, D]                                         */
  const void *t1_w, *t1_b, *t2_w, *t2_b; /* time_text_embed.timestep_embedder.linear_{1,2}               */
  const void *p1_w, *p1_b, *p2_w, *p2_b; /* time_text_embed.text_embedder.linear_{1,2}                   */
  const void *ctx_w, *ctx_b;    /* context_embedder                                                       */
  /* all adaLN projections stacked row-wise: for each layer [norm1.linear ; norm1_context.linear], then norm_out.linear */
  const void *mod_w, *mod_b;
  const void *proj_w, *proj_b;  /* proj_out                                                               */
  const ffb200_layer_weights* layers; /* [num_layers] */
} ffb200_weights;

typedef struct ffb200_engine ffb200_engine;
typedef struct ffb200_plan ffb200_plan;

/* Replaces BaseAdapter.load_pipeline()'s transformer for the rollout path (FF/models/abc.py:185-188;
 * FF/models/stable_diffusion/sd3_5.py:60-120).  Weight pointers are borrowed and may be refreshed with
 * ffb200_engine_set_weights() after an optimizer / EMA / LoRA-merge step. */
int ffb200_engine_create(const ffb200_model_config* cfg, const ffb200_weights* w, ffb200_engine** out);
int ffb200_engine_set_weights(ffb200_engine* e, const ffb200_weights* w);
void ffb200_engine_destroy(ffb200_engine* e);
/* total rows of the stacked adaLN matrix (for the host-side packer) */
int ffb200_engine_mod_rows(const ffb200_engine* e);

/* A plan fixes the geometry (batch, CFG on/off, latent H x W, text tokens) and owns the workspace + TMA descriptors. */
int ffb200_plan_create(ffb200_engine* e, int batch, int cfg, int lat_h, int lat_w, int n_text, ffb200_plan** out);
void ffb200_plan_destroy(ffb200_plan* p);
long long ffb200_plan_workspace_bytes(const ffb200_plan* p);

/* latent_storage_dtype of this plan's latents buffers (FF/hparams/training_args.py:245-252, BaseAdapter.cast_latents, FF/models/abc.py:172-182):
 * 0 = fp16 (default; values beyond +-65504 are clamped and the overflow flag raised), 1 = bf16, 2 = fp32.  It is the element type of every
 * `latents` / `next_latents` / `all_latents` / `final_latents` pointer of ffb200_step / ffb200_rollout[_host] and the dtype freshly sampled
 * next_latents are rounded through before the log-prob (FF/scheduler/flow_match_euler_discrete.py:359-362). */
#define FFB200_LAT_F16 0
#define FFB200_LAT_BF16 1
#define FFB200_LAT_F32 2
int ffb200_plan_set_latent_dtype(ffb200_plan* p, int dtype);
/* Prompt conditioning for the next forward()/rollout() calls: bf16 [Bp, n_text, joint_dim] and [Bp, pooled_dim]
 * with Bp = batch * (cfg ? 2 : 1), negative (unconditional) half FIRST (sd3_5.py:409-413).  Runs context_embedder
 * and the pooled-text MLP once (they do not depend on the timestep). */
int ffb200_plan_set_prompts(ffb200_plan* p, const void* prompt_embeds_bf16, const void* pooled_bf16, void* stream);

/* Per-step scalars computed on the host in fp32 exactly as FlowMatchEulerDiscreteSDEScheduler.step builds its
 * (B,1,1,1) tensors (FF/scheduler/flow_match_euler_discrete.py:299-420).  Filled by flow_factory_b200.scheduler. */
typedef struct ffb200_step_coef {
  float t_model;
  float sigma, sigma_prev, dt, noise_level, std_dev_t;
  float c_x, c_v, noise_scale, two_var, log_norm, cps_a, cps_b;
  int dynamics;          /* 0 Flow-SDE, 1 Dance-SDE, 2 CPS, 3 ODE */
  int compute_log_prob;
  int store_slot;        /* slot of all_latents receiving this step's result, or -1 */
  int logp_slot;         /* slot of log_probs receiving this step's log-prob, or -1 */
} ffb200_step_coef;

/* SD3Transformer2DModel.forward + CFG batching (DF/models/transformers/transformer_sd3.py:249-345;
 * sd3_5.py:409-428): latents fp16 [B,C,H,W] -> proj_out rows bf16 [Bp*Ni, p*p*C] (token-major, pre-unpatchify).
 * `noise_pred_nchw` (optional, bf16 [Bp,C,H,W]) receives the unpatchified prediction for inspection. */
int ffb200_transformer_forward(ffb200_plan* p, const void* latents_fp16, float t_model, void* noise_pred_nchw, void* stream);

/* SD3_5Adapter.forward under no_grad (sd3_5.py:352-448): one denoise step = transformer + CFG + scheduler.step. */
typedef struct ffb200_step_args {
  const void* latents;        /* fp16 [B,C,H,W] */
  ffb200_step_coef coef;
  float guidance_scale;
  const float* noise;         /* fp32 [B,C,H,W] or NULL (in-kernel Philox) */
  unsigned long long seed;
  int step_index;
  const void* next_latents;   /* teacher-forced fp16 [B,C,H,W] or NULL */
  void* out_next_latents;     /* fp16 [B,C,H,W] or NULL */
  float* out_mean;            /* fp32 [B,C,H,W] or NULL */
  float* out_log_prob;        /* fp32 [B] or NULL */
  void* out_noise_pred;       /* bf16 [B,C,H,W] (after CFG) or NULL */
  int* overflow_flag;         /* device int or NULL */
} ffb200_step_args;
int ffb200_step(ffb200_plan* p, const ffb200_step_args* a, void* stream);

/* SD3_5Adapter.inference's denoising loop (sd3_5.py:266-304): T steps, no host sync inside.
 * all_latents: fp16 [B, n_latent_slots, C,H,W]; log_probs: fp32 [B, n_logp_slots]. */
typedef struct ffb200_rollout_args {
  int num_steps;
  const ffb200_step_coef* coefs;  /* host array [num_steps] */
  float guidance_scale;
  const void* x0;                 /* fp16 [B,C,H,W] (already cast to the storage dtype) */
  const float* noise;             /* fp32 [num_steps,B,C,H,W] or NULL -> in-kernel Philox keyed by (seed, step) */
  unsigned long long seed;
  void* all_latents; int n_latent_slots; int store_initial_slot; /* slot for position 0 or -1 */
  float* log_probs; int n_logp_slots;
  void* final_latents;            /* fp16 [B,C,H,W] */
  int* overflow_flag;
  int use_graph;                  /* 1: replay one captured per-step CUDA graph */
} ffb200_rollout_args;
int ffb200_rollout(ffb200_plan* p, const ffb200_rollout_args* a, void* stream);
/* Same, but every pointer in `a` (and the prompts) is a HOST buffer: copies in/out are part of the call. */
int ffb200_rollout_host(ffb200_plan* p, const ffb200_rollout_args* a, const void* prompt_embeds_bf16,
                        const void* pooled_bf16, void* stream);
/* number of kernels the last ffb200_rollout / ffb200_step / ffb200_transformer_forward launched */
long long ffb200_last_launch_count(void);

/* ---------------------------------------------------------------- op-level entries (one reference call site each) */
/* nn.Linear + fused epilogue. A: bf16 [num_batch][rows_per_batch][K] (row stride lda, batch stride a_batch_stride, in
 * elements); W: bf16 [N,K]; epilogue: 0 bias, 1 bias+GELU-tanh, 2 gate*y + residual (in place), 3 fused qkv + RMSNorm,
 * 4 bias + fp32 row table. */
int ffb200_linear(const void* A, int num_batch, int rows_per_batch, long long a_batch_stride, int lda, int K,
                  const void* W, int N, const void* bias, void* out, long long out_batch_stride, int out_row_offset,
                  int ldo, int epilogue, const void* gate, long long gate_batch_stride, const void* norm_q,
                  const void* norm_k, int qk_dim, float eps, const float* row_table, void* stream);
/* FLUX.1 fused q|k|v projection (head_dim 128): nn.Linear + torch.nn.RMSNorm on the q and k heads + apply_rotary_emb with
 * interleaved pairs, written token-major into a joint [B, S, 3*qk_dim] buffer at row `out_row_offset` (text rows first, then image:
 * DF/models/transformers/transformer_flux.py:87-117; DF/models/embeddings.py:1207-1233).  rope_cos / rope_sin: fp32 [tokens, 128]
 * as FluxPosEmbed returns them (transformer_flux.py:500-522); the table row of GEMM row r is rope_row_offset + r. */
int ffb200_linear_qkv_rope(const void* A, int num_batch, int rows_per_batch, long long a_batch_stride, int lda, int K,
                           const void* W, int N, const void* bias, void* out, long long out_batch_stride, int out_row_offset,
                           int ldo, const void* norm_q, const void* norm_k, int qk_dim, float eps, const float* rope_cos,
                           const float* rope_sin, int rope_row_offset, void* stream);
/* F.scaled_dot_product_attention over a fused token-major qkv buffer bf16 [B, S, 3*64*H] -> out bf16 [B, S, 64*H]
 * (DF/models/attention_processor.py:1480-1486). */
int ffb200_attention(const void* qkv, int batch, int seq_len, int num_heads, void* out, void* stream);
/* Same with head_dim 64 or 128 and an output row stride (elements; 0 = dense 64/128*H): head_dim 128 is the FLUX.1 joint
 * [text ; image] attention, dispatch_attention_fn at DF/models/transformers/transformer_flux.py:118-125 (q, k already RMS-normed
 * and rotated); a wider row stride lets FluxSingleTransformerBlock's torch.cat([attn_output, mlp_hidden_states], dim=2)
 * (transformer_flux.py:400) be the attention kernel's own store.  head_dim 64 accepts dense rows only. */
int ffb200_attention_ex(const void* qkv, int batch, int seq_len, int num_heads, int head_dim, void* out, int out_row_stride,
                        void* stream);
/* Same with an explicit softmax scale (`scale=` of F.scaled_dot_product_attention / dispatch_attention_fn, DF/models/attention_dispatch.py:
 * 2930-2945; <= 0 selects 1/sqrt(head_dim)), and `k_prescaled`: 1 = the caller already multiplied the keys by softmax_scale * log2(e)
 * (what the engines' QKV projection epilogue does), so the kernel takes the q.k scores as base-2 exponents and skips the per-score multiply. */
int ffb200_attention_scaled(const void* qkv, int batch, int seq_len, int num_heads, int head_dim, void* out, int out_row_stride,
                            float softmax_scale, int k_prescaled, void* stream);
/* head_dim 64, keys pre-scaled, q and k produced by a per-head RMSNorm (DF/models/attention_processor.py:1456-1459, 1470-1473: `norm_q`,
 * `norm_k`, `norm_added_q`, `norm_added_k`) whose weights [64] the caller names: wq0 / wk0 (image stream) and optionally wq1 / wk1 (text
 * stream of a joint attention; may be null).  RMS-normed heads bound every score (Cauchy-Schwarz: |q.k'| <= 64 max|wq| max|wk| scale*log2e),
 * which lets the kernel skip the per-tile range check of its polynomial exp2 (valid for |exponent| <= 126) when that bound stays below 120;
 * results are bit-identical to ffb200_attention_scaled(k_prescaled = 1).  This is the call the SD3.5 engine makes internally; passing weights that did NOT produce q / k
 * voids the bound (then use ffb200_attention_scaled). */
int ffb200_attention_normed(const void* qkv, int batch, int seq_len, int num_heads, void* out, const void* wq0, const void* wk0,
                            const void* wq1, const void* wk1, void* stream);
/* LayerNorm(no affine) * (1 + scale) + shift  (DF/models/normalization.py:120-126). vectors: [num_batch, *] with stride. */
int ffb200_ln_modulate(const void* x, int num_batch, int rows_per_batch, int D, float eps, const void* shift1,
                       const void* scale1, void* out1, const void* shift2, const void* scale2, void* out2,
                       long long mod_batch_stride, void* stream);
/* skinny-batch nn.Linear (optionally SiLU on the input, optional bf16 addend). */
int ffb200_small_linear(const void* in, int batch, int K, long long in_stride, const void* W, const void* bias, int N,
                        void* out, long long out_stride, const void* addend, long long addend_stride, int silu_input,
                        void* stream);
/* FlowMatchEulerDiscreteSDEScheduler.step on an NCHW bf16 noise_pred (FF/scheduler/flow_match_euler_discrete.py:243-438). */
int ffb200_sde_step(const void* noise_pred_bf16, const void* latents_fp16, int B, int C, int H, int W,
                    const ffb200_step_coef* coef, const float* noise, unsigned long long seed, int step_index,
                    const void* next_latents_fp16, void* out_next_fp16, float* out_mean, float* out_log_prob,
                    int* overflow_flag, void* stream);
/* Same with the latents' storage dtype (FFB200_LAT_*): element type of latents / next_latents / out_next, and the round trip of a sample. */
int ffb200_sde_step_ex(const void* noise_pred_bf16, const void* latents, int B, int C, int H, int W, const ffb200_step_coef* coef,
                       const float* noise, unsigned long long seed, int step_index, const void* next_latents, void* out_next,
                       float* out_mean, float* out_log_prob, int* overflow_flag, int storage_dtype, void* stream);

/* ================================================================ FLUX.1 (BASELINE config 3)
 * Same boundary for the FLUX.1 rollout path: FluxTransformer2DModel.forward (DF/models/transformers/transformer_flux.py:676-778)
 * behind Flux1Adapter.inference / forward (FF/models/flux/flux1.py:152-292, 296-349).  No CFG batch: guidance is an embedded
 * scalar.  Latents are the PACKED [B, Ni, 64] tensors of the reference (FluxPipeline._pack_latents), stored fp16. */
typedef struct ffb200_flux_config {
  int num_layers;            /* dual-stream FluxTransformerBlocks (19) */
  int num_single_layers;     /* FluxSingleTransformerBlocks (38) */
  int num_heads;             /* inner_dim D = 128 * num_heads (head_dim is 128) */
  int in_channels;           /* 64 */
  int joint_attention_dim;   /* 4096 */
  int pooled_projection_dim; /* 768 */
  int guidance_embeds;       /* 1 for FLUX.1-dev */
  int variant;               /* 0 = FLUX.1 ; 1 = Qwen-Image (dual blocks only, see below) */
} ffb200_flux_config;

/* bf16, nn.Linear layout [out, in]; q|k|v concatenated along out_features by the host packer */
typedef struct ffb200_flux_dual_weights {
  const void *qkv_w, *qkv_b, *norm_q, *norm_k;                     /* attn.to_{q,k,v}, attn.norm_{q,k} [128]              */
  const void *add_qkv_w, *add_qkv_b, *norm_added_q, *norm_added_k; /* attn.add_{q,k,v}_proj, attn.norm_added_{q,k}        */
  const void *out_w, *out_b, *add_out_w, *add_out_b;               /* attn.to_out.0, attn.to_add_out                      */
  const void *ff1_w, *ff1_b, *ff2_w, *ff2_b;                       /* ff.net.0.proj, ff.net.2                             */
  const void *cff1_w, *cff1_b, *cff2_w, *cff2_b;                   /* ff_context.*                                        */
} ffb200_flux_dual_weights;
typedef struct ffb200_flux_single_weights {
  const void *qkv_w, *qkv_b, *norm_q, *norm_k;   /* attn.to_{q,k,v}, attn.norm_{q,k}   */
  const void *mlp_w, *mlp_b;                     /* proj_mlp [4D, D]                   */
  const void *out_w, *out_b;                     /* proj_out [D, 5D] over [attn | mlp] */
} ffb200_flux_single_weights;
typedef struct ffb200_flux_weights {
  const void *x_w, *x_b;                 /* x_embedder [D, 64]                                                       */
  const void *ctx_w, *ctx_b;             /* context_embedder                                                         */
  const void *t1_w, *t1_b, *t2_w, *t2_b; /* time_text_embed.timestep_embedder.linear_{1,2}                           */
  const void *g1_w, *g1_b, *g2_w, *g2_b; /* time_text_embed.guidance_embedder.linear_{1,2} (NULL unless guidance_embeds) */
  const void *p1_w, *p1_b, *p2_w, *p2_b; /* time_text_embed.text_embedder.linear_{1,2}                               */
  /* all adaLN projections stacked row-wise: per dual block [norm1.linear (6D) ; norm1_context.linear (6D)], then per single
   * block norm.linear (3D), then norm_out.linear (2D) */
  const void *mod_w, *mod_b;
  const void *proj_w, *proj_b;           /* proj_out [64, D]                                                         */
  const void *ctxn_w;                    /* Qwen-Image only: txt_norm.weight [joint_dim] (NULL for FLUX.1)           */
  const ffb200_flux_dual_weights* dual;     /* [num_layers]        */
  const ffb200_flux_single_weights* single; /* [num_single_layers] */
} ffb200_flux_weights;

typedef struct ffb200_flux_engine ffb200_flux_engine;
typedef struct ffb200_flux_plan ffb200_flux_plan;

int ffb200_flux_engine_create(const ffb200_flux_config* cfg, const ffb200_flux_weights* w, ffb200_flux_engine** out);
int ffb200_flux_engine_set_weights(ffb200_flux_engine* e, const ffb200_flux_weights* w);
void ffb200_flux_engine_destroy(ffb200_flux_engine* e);
int ffb200_flux_engine_mod_rows(const ffb200_flux_engine* e);
/* Geometry: batch, image tokens Ni = (h/16)(w/16), text tokens; rope_cos / rope_sin: fp32 [Nt + Ni, 128] for ids = cat(txt_ids,
 * img_ids) exactly as FluxPosEmbed returns them (transformer_flux.py:500-522, float64 frequencies) - host or device memory, copied. */
int ffb200_flux_plan_create(ffb200_flux_engine* e, int batch, int n_img_tokens, int n_text, const float* rope_cos,
                            const float* rope_sin, ffb200_flux_plan** out);
/* variant 1 = Qwen-Image (QwenImageTransformer2DModel.forward, DF/models/transformers/transformer_qwenimage.py:878-993,
 * behind QwenImageAdapter.forward, FF/models/qwen_image/qwen_image.py:476-600): the same dual-stream engine with
 * x_w/ctx_w := img_in/txt_in, norm1(.context).linear := img_mod.1/txt_mod.1, ff(.context) := img_mlp/txt_mlp, the text RMSNorm, the
 * timestep-only conditioning (t1/t2; p*, g* unused), diffusers-style q/k RMSNorm, rope tables that also rotate the text rows, and
 * - with cfg = 1 - a forward batch of 2B (negative prompts first) combined by the per-token norm-rescaled true CFG
 * (qwen_image.py:580-587); the CFG scale is the `guidance_model` argument of ffb200_flux_set_prompts, `pooled_bf16` is NULL.
 * Prompts of one call must share one (unpadded) length: key-padding masks are not implemented. */
int ffb200_flux_plan_create_ex(ffb200_flux_engine* e, int batch, int cfg, int n_img_tokens, int n_text, const float* rope_cos,
                               const float* rope_sin, ffb200_flux_plan** out);
/* Qwen-Image key-padding mask (encoder_hidden_states_mask -> attention_mask, transformer_qwenimage.py:941-958): lengths[i] = number of
 * valid (leading) text tokens of forward-batch row i (2B rows with cfg, negative prompts first); keys [lengths[i], n_text) are masked.
 * Default after plan creation: no padding. */
int ffb200_flux_set_text_lengths(ffb200_flux_plan* p, const int* lengths_host, void* stream);
void ffb200_flux_plan_destroy(ffb200_flux_plan* p);
long long ffb200_flux_plan_workspace_bytes(const ffb200_flux_plan* p);
/* prompt_embeds bf16 [B, Nt, joint_dim], pooled bf16 [B, pooled_dim]; guidance_model = float(bf16(bf16(guidance_scale) * 1000)),
 * the value transformer_flux.py:681-682 feeds to the sinusoid (flux1.py:318-319 builds the tensor in the latents' dtype). */
int ffb200_flux_set_prompts(ffb200_flux_plan* p, const void* prompt_embeds_bf16, const void* pooled_bf16, float guidance_model,
                            void* stream);
/* FluxTransformer2DModel.forward: packed latents fp16 [B, Ni, 64] -> noise prediction bf16 [B, Ni, 64].
 * t_model = float(bf16(bf16(t / 1000) * 1000)) (flux1.py:325 ; transformer_flux.py:679). */
int ffb200_flux_forward(ffb200_flux_plan* p, const void* latents_fp16, float t_model, void* noise_pred_bf16, void* stream);
/* Flux1Adapter.forward under no_grad / the denoise loop of Flux1Adapter.inference: same argument blocks as ffb200_step /
 * ffb200_rollout with [C,H,W] := [Ni*64] packed elements per sample; guidance_scale in the blocks is ignored (set_prompts). */
int ffb200_flux_step(ffb200_flux_plan* p, const ffb200_step_args* a, void* stream);
int ffb200_flux_rollout(ffb200_flux_plan* p, const ffb200_rollout_args* a, void* stream);

/* ================================================================ Wan2.1 T2V (SURVEY.md section 8f row 4 / BASELINE config 4)
 * STATUS: added at the end of round 1 after the GPU budget was spent: compiles for sm_100a, host logic unit-tested on CPU, first GPU
 * run pending (tests/test_gpu_wan.py, gated on FFB200_PENDING=1).
 *
 * WanTransformer3DModel.forward (DF/models/transformers/transformer_wan.py:629-740; blocks 462-505, attention processor 78-162, rotary
 * embedding 354-417, condition embedder 330-351) behind Wan2_T2V_Adapter.forward / .inference (FF/models/wan/wan2_t2v.py:235-543):
 * true CFG as a batch of 2B (negative prompts first), u + g (c - u) in bf16, then the Euler / SDE step of UniPCMultistepSDEScheduler
 * (FF/scheduler/unipc_multistep.py:290-421 - the same arithmetic as FlowMatchEulerDiscreteSDEScheduler.step).
 * Latents fp16 [B, 16, F, H, W]; patch (1, 2, 2); head_dim 128; T2V only (no image conditioning, one scalar timestep per sample). */
typedef struct ffb200_wan_config {
  int num_layers;     /* 30 (1.3 B) / 40 (14 B) */
  int num_heads;      /* inner_dim D = 128 * num_heads: 12 / 40 */
  int in_channels;    /* 16 (== out_channels) */
  int text_dim;       /* 4096 */
  int freq_dim;       /* 256 */
  int ffn_dim;        /* 8960 / 13824 */
  int patch_t, patch_h, patch_w; /* 1, 2, 2 (patch_t must be 1) */
  float eps;          /* 1e-6 */
} ffb200_wan_config;

/* bf16, nn.Linear layout [out, in]; q|k|v (attn1) and k|v (attn2) concatenated along out_features by the host packer */
typedef struct ffb200_wan_layer_weights {
  const void *table;                                  /* scale_shift_table [6, D] (module dtype)            */
  const void *qkv_w, *qkv_b, *norm_q, *norm_k;        /* attn1.to_{q,k,v}, attn1.norm_{q,k} [D]             */
  const void *out_w, *out_b;                          /* attn1.to_out.0                                     */
  const void *norm2_w, *norm2_b;                      /* norm2 (FP32LayerNorm, elementwise_affine) [D]      */
  const void *q2_w, *q2_b, *kv2_w, *kv2_b;            /* attn2.to_q ; attn2.to_{k,v} [2D, D]                */
  const void *norm_q2, *norm_k2;                      /* attn2.norm_{q,k} [D]                               */
  const void *out2_w, *out2_b;                        /* attn2.to_out.0                                     */
  const void *ff1_w, *ff1_b, *ff2_w, *ff2_b;          /* ffn.net.0.proj [ffn, D], ffn.net.2 [D, ffn]        */
} ffb200_wan_layer_weights;
typedef struct ffb200_wan_weights {
  const void *pe_w, *pe_b;                 /* patch_embedding as [D, C*pt*ph*pw]                                  */
  const void *t1_w, *t1_b, *t2_w, *t2_b;   /* condition_embedder.time_embedder.linear_{1,2}                       */
  const void *tp_w, *tp_b;                 /* condition_embedder.time_proj [6D, D]                                */
  const void *x1_w, *x1_b, *x2_w, *x2_b;   /* condition_embedder.text_embedder.linear_{1,2}                       */
  const void *table;                       /* scale_shift_table [2, D]                                            */
  const void *proj_w, *proj_b;             /* proj_out [C*pt*ph*pw, D]                                            */
  const ffb200_wan_layer_weights* layers;  /* [num_layers]                                                        */
} ffb200_wan_weights;

typedef struct ffb200_wan_engine ffb200_wan_engine;
typedef struct ffb200_wan_plan ffb200_wan_plan;

int ffb200_wan_engine_create(const ffb200_wan_config* cfg, const ffb200_wan_weights* w, ffb200_wan_engine** out);
int ffb200_wan_engine_set_weights(ffb200_wan_engine* e, const ffb200_wan_weights* w);
void ffb200_wan_engine_destroy(ffb200_wan_engine* e);
/* Geometry: batch, cfg (1 = forward batch 2B), latent frames / height / width, text tokens; rope_cos / rope_sin: fp32 [S, 128] with
 * S = (F/pt)(H/ph)(W/pw), the values WanRotaryPosEmbed returns AFTER the cast to the module dtype (its buffers follow `.to(bf16)`) -
 * host or device memory, copied. */
int ffb200_wan_plan_create(ffb200_wan_engine* e, int batch, int cfg, int frames, int height, int width, int n_text, const float* rope_cos,
                           const float* rope_sin, ffb200_wan_plan** out);
void ffb200_wan_plan_destroy(ffb200_wan_plan* p);
long long ffb200_wan_plan_workspace_bytes(const ffb200_wan_plan* p);
/* prompt_embeds bf16 [Bp, Nt, text_dim] (Bp = 2B with cfg: negative prompts first): text embedder + the cross-attention keys / values
 * of every block, cached for the whole rollout (they do not depend on the timestep). */
int ffb200_wan_set_prompts(ffb200_wan_plan* p, const void* prompt_embeds_bf16, void* stream);
/* WanTransformer3DModel.forward + CFG combine: latents fp16 [B, C, F, H, W] -> noise prediction bf16 [B, C, F, H, W].
 * t_model = the fp32 timestep on the 0..1000 scale (wan2_t2v.py:499-506). */
int ffb200_wan_forward(ffb200_wan_plan* p, const void* latents_fp16, float t_model, float guidance_scale, void* noise_pred_bf16, void* stream);
/* Wan2_T2V_Adapter.forward under no_grad / the denoise loop of .inference: same argument blocks as ffb200_step / ffb200_rollout with
 * [C, H, W] := [C, F*H, W]. */
int ffb200_wan_step(ffb200_wan_plan* p, const ffb200_step_args* a, void* stream);
int ffb200_wan_rollout(ffb200_wan_plan* p, const ffb200_rollout_args* a, void* stream);

/* Op-level entries (one per reference call site), used by the parity tests to isolate a kernel:
 * torch.nn.RMSNorm over the full width D of rows with pitch ld, in place, + optional interleaved-pair RoPE per 128-wide head evaluated in
 * bf16 like the reference's tensor arithmetic (transformer_wan.py:96-117); cos / sin: fp32 [rows_per_batch, 128] or NULL. */
int ffb200_wan_rms_rope(void* x_bf16, long long rows, int rows_per_batch, int ld, int D, const void* weight_bf16, float eps,
                        const float* cos, const float* sin, void* stream);
/* FP32LayerNorm (no affine) + fp32 modulate: out = bf16(LN(x) * (1 + scale[b]) + shift[b]) with fp32 vectors `mod_batch_stride` apart
 * (mode 0, transformer_wan.py:486, 500), or FP32LayerNorm with bf16 affine weight / bias [D] (mode 1, norm2, 493). */
int ffb200_wan_layer_norm(const void* x_bf16, void* out_bf16, int num_batch, int rows_per_batch, int D, float eps, int mode,
                          const float* scale, const float* shift, long long mod_batch_stride, const void* weight_bf16,
                          const void* bias_bf16, void* stream);
/* hs = (hs.float() + y * gate[b]).type_as(hs) (transformer_wan.py:489, 503); gate fp32, `gate_batch_stride` apart. */
int ffb200_wan_gate_residual(void* h_bf16, const void* y_bf16, const float* gate, long long gate_batch_stride, int num_batch,
                             long long rows_per_batch, int D, void* stream);
/* im2col of the Conv3d patch embedding: fp16 [B, C, F, H, W] -> bf16 [reps * B * tokens, C*pt*ph*pw]. */
int ffb200_wan_patchify(const void* x_f16, int B, int reps, int C, int F, int H, int W, int pt, int ph, int pw, void* out_bf16, void* stream);
/* Cross-attention, head_dim 128: q bf16 [B, Sq, q_ld] (head h at column 128 h), kv bf16 [B, Skv, kv_ld] (k at column 128 h, v at
 * v_col + 128 h) -> out bf16 [B, Sq, 128 H]   (WanAttnProcessor with encoder_hidden_states, transformer_wan.py:78-162). */
int ffb200_attention_cross(const void* q_bf16, int q_ld, const void* kv_bf16, int kv_ld, int v_col, int B, int Sq, int Skv, int num_heads,
                           void* out_bf16, void* stream);

/* ================================================================ VAE decode
 * STATUS: added at the end of round 1 after the GPU budget was spent: compiles for sm_100a, host logic unit-tested on CPU, first GPU
 * run pending (tests/test_gpu_vae.py, gated on FFB200_PENDING=1).
 *
 * Replaces SD3_5Adapter.decode_latents (FF/models/stable_diffusion/sd3_5.py:161-172) = AutoencoderKL.decode ->
 * Decoder.forward (DF/models/autoencoders/vae.py:279-316) under the trainer's bf16 autocast: conv_in, UNetMidBlock2D
 * (ResnetBlock2D, single-head attention, ResnetBlock2D), UpDecoderBlock2D x n (ResnetBlock2D x (layers_per_block + 1), nearest 2x
 * upsample + conv), GroupNorm, SiLU, conv_out.  Activations are NHWC bf16; every convolution / linear runs on the tensor cores. */
typedef struct ffb200_vae_config {
  int latent_channels;        /* 16 */
  int out_channels;           /* 3 */
  int num_blocks;             /* len(block_out_channels), <= 8 */
  int block_out_channels[8];  /* encoder order, e.g. 128 256 512 512 (the decoder walks it reversed); multiples of 8, the last one of 64 */
  int layers_per_block;       /* 2 (the decoder uses layers_per_block + 1 resnets per up block) */
  int norm_num_groups;        /* 32 */
  float scaling_factor;       /* 1.5305 */
  float shift_factor;         /* 0.0609 */
} ffb200_vae_config;

typedef struct ffb200_vae_decoder ffb200_vae_decoder;

/* Number of weight pointers ffb200_vae_decoder_create expects for `cfg` (the order is documented in flow_factory_b200/vae.py:
 * conv_in, mid resnet 0, mid attention, mid resnet 1, up blocks, conv_norm_out, conv_out; bf16; 3x3 kernels packed [Cout][tap][Cin
 * padded to 64]; biases padded to a multiple of 8).  Negative on an invalid config. */
int ffb200_vae_weight_count(const ffb200_vae_config* cfg);
/* Builds the launch list and the workspace for decoding `batch` latents of lat_h x lat_w at a time.  Weight pointers are borrowed
 * and must stay valid while the decoder lives. */
int ffb200_vae_decoder_create(const ffb200_vae_config* cfg, const void* const* weights, int n_weights, int batch, int lat_h, int lat_w,
                              ffb200_vae_decoder** out);
void ffb200_vae_decoder_destroy(ffb200_vae_decoder* d);
long long ffb200_vae_decoder_workspace_bytes(const ffb200_vae_decoder* d);
/* latents: fp16 [batch, latent_channels, lat_h, lat_w] (the rollout's final latents); image: bf16 [batch, out_channels, 8 lat_h, 8 lat_w]
 * (AutoencoderKL.decode(...)[0], before image_processor.postprocess). */
int ffb200_vae_decode(ffb200_vae_decoder* d, const void* latents_f16, void* image_bf16, void* stream);

/* Op-level entries (one per reference call site), used by the parity tests:
 * nn.Conv2d(k=3, p=1) / nn.Conv2d(k=1) on NHWC bf16 x [B, H, W, Cin] -> out [B, H, W, Cout]; w packed as above (taps = 9) or the plain
 * [Cout, Cin] matrix (taps = 1); residual (optional, NHWC [B, H, W, Cout]) is added after the bias with its own bf16 rounding. */
int ffb200_conv2d_nhwc(const void* x, const void* w_packed, const void* bias, const void* residual, void* out, int B, int H, int W,
                       int Cin, int Cout, int taps, void* stream);
/* nn.GroupNorm(groups, C, eps, affine) (+ SiLU) on NHWC bf16 [B, P, C]; workspace: >= B * C * 16 bytes (zeroed by the call). */
int ffb200_group_norm_nhwc(const void* x, const void* gamma, const void* beta, void* out, int B, long long P, int C, int groups,
                           float eps, int silu, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FFB200_H */
