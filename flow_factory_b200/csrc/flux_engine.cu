// Host side of the FLUX.1 rollout path (SURVEY.md 8f row 2 / BASELINE config 3): FluxTransformer2DModel.forward + Flux1Adapter's
// denoise loop as one launch list per step, built from the same tcgen05 GEMM / attention / elementwise kernels as the SD3.5 engine.
//
//   FluxTransformer2DModel.forward        DF/models/transformers/transformer_flux.py:676-778
//   FluxTransformerBlock (dual stream)    transformer_flux.py:438-492      FluxSingleTransformerBlock   378-407
//   FluxAttnProcessor                     transformer_flux.py:83-139       FluxPosEmbed / apply_rotary_emb  500-522 / embeddings.py:1207-1233
//   CombinedTimestepGuidanceTextProjEmbeddings   embeddings.py:1603-1624
//   Flux1Adapter.inference / forward      FF/models/flux/flux1.py:152-292 / 296-349  (no CFG: guidance is an embedded scalar)
//
// Layout: the residual stream is ONE joint buffer h [B, S = Nt + Ni, D] with the text rows first (transformer_flux.py:110-112, 382):
// the dual-stream blocks address its two row ranges as separate GEMM problems, the single-stream blocks use it whole, so the
// torch.cat / split pairs of the reference (382, 406) disappear.  q|k|v are projected by one GEMM per stream whose epilogue applies
// the per-head RMSNorm and the rotary embedding and writes token-major into the joint qkv buffer; the single block's
// torch.cat([attn_output, mlp_hidden_states]) (400) is the attention kernel / MLP GEMM writing column blocks of one [B, S, 5D] buffer.
#include "common.cuh"
#include "kernels.h"
#include "../../include/ffb200.h"

struct FluxOffsets { std::vector<int> n1, n1c, ns; int out, rows; };   // row offsets into the stacked adaLN matrix

struct ffb200_flux_engine {
  ffb200_flux_config cfg;
  ffb200_flux_weights w;
  std::vector<ffb200_flux_dual_weights> dual;
  std::vector<ffb200_flux_single_weights> single;
  int D;
  FluxOffsets off;
};

struct ffb200_flux_plan {
  ffb200_flux_engine* e;
  int B, Bp, cfg, Ni, Nt, S, D;   // Bp = forward batch (2B with Qwen-Image's true CFG, negative half first)
  float cfg_scale;
  std::vector<void*> allocs;
  long long ws_bytes;
  bf16 *xin, *c0, *ctxn, *tproj, *gproj, *ta, *ga, *gemb, *pa, *pemb, *temb, *mod, *h, *a1, *qkv, *att, *ff, *cat, *vout, *vcfg;
  float *rope_cos, *rope_sin;
  __half* x_cur;
  float* logp_partial;
  int* d_step;
  StepCoef* d_coefs; int coef_cap;
  StepCoef* d_gcoef;
  int* d_txt_len;   // [Bp] valid text tokens per forward-batch row (Qwen-Image key-padding mask); Nt = no padding
  std::vector<Op> fwd_ops;
  bool prompts_set;
  cudaGraphExec_t graph_exec; SdeStepParams graph_sde; bool graph_valid; long long graph_launches;
};

static int fplan_alloc(ffb200_flux_plan* p, void** ptr, size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaMalloc(flux workspace)");
  p->allocs.push_back(*ptr);
  p->ws_bytes += static_cast<long long>(bytes);
  return 0;
}
static int fadd_gemm(ffb200_flux_plan* p, const GemmSpec& s) {
  GemmParams gp;
  int r = build_gemm(s, &gp);
  if (r) return r;
  const int sms = num_sms();
  p->fwd_ops.push_back([gp, sms](cudaStream_t st) { ++g_launch_count; return launch_gemm(gp, sms, st); });
  return 0;
}
static void fadd_lnmod(ffb200_flux_plan* p, const bf16* x, long x_bs, int rows_per_batch, const bf16* shift, const bf16* scale, bf16* out,
                       long out_bs) {
  LnModParams lp{};
  lp.x = x; lp.rows_per_batch = rows_per_batch; lp.num_batch = p->Bp; lp.D = p->D; lp.eps = 1e-6f;
  lp.shift1 = shift; lp.scale1 = scale; lp.out1 = out;
  lp.mod_batch_stride = p->e->off.rows; lp.x_batch_stride = x_bs; lp.out_batch_stride = out_bs;
  p->fwd_ops.push_back([lp](cudaStream_t st) { ++g_launch_count; return launch_ln_modulate(lp, st); });
}
static Op small_op(int batch, const bf16* in, int K, const void* W, const void* bias, int N, bf16* out, const bf16* addend,
                   const bf16* addend2, int silu) {
  SmallLinearParams sp{};
  sp.in = in; sp.batch = batch; sp.K = K; sp.in_stride = K; sp.W = static_cast<const bf16*>(W);
  sp.bias = static_cast<const bf16*>(bias); sp.N = N; sp.out = out; sp.out_stride = N;
  sp.addend = addend; sp.addend2 = addend2; sp.addend_stride = N; sp.silu_input = silu;
  return [sp](cudaStream_t st) { ++g_launch_count; return launch_small_linear(sp, st); };
}

extern "C" {

int ffb200_flux_engine_set_weights(ffb200_flux_engine* e, const ffb200_flux_weights* w) {
  FFB_CHECK(e && w && w->dual && w->single, "null engine/weights");
  e->w = *w;
  e->dual.assign(w->dual, w->dual + e->cfg.num_layers);
  e->single.assign(w->single, w->single + e->cfg.num_single_layers);
  e->w.dual = e->dual.data();
  e->w.single = e->single.data();
  return 0;
}

int ffb200_flux_engine_create(const ffb200_flux_config* cfg, const ffb200_flux_weights* w, ffb200_flux_engine** out) {
  FFB_CHECK(cfg && w && out, "null argument");
  FFB_CHECK(cfg->num_layers >= 0 && cfg->num_single_layers >= 0 && cfg->num_heads > 0, "bad config");
  FFB_CHECK(cfg->in_channels == 64, "packed latents: in_channels must be 64");
  FFB_CHECK(cfg->variant == 0 || cfg->variant == 1, "variant: 0 = FLUX.1, 1 = Qwen-Image");
  if (cfg->variant == 1) FFB_CHECK(cfg->num_single_layers == 0 && cfg->guidance_embeds == 0 && w->ctxn_w, "Qwen-Image: dual blocks only, no guidance embedding, txt_norm weight required");
  FFB_CHECK(cfg->joint_attention_dim % 8 == 0 && cfg->pooled_projection_dim % 8 == 0, "joint/pooled dims must be multiples of 8");
  int dev = 0, major = 0;
  FFB_CUDA(cudaGetDevice(&dev));
  FFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FFB_CHECK(major == 10, "ffb200 kernels are sm_100a only (no fallback path)");
  ffb200_flux_engine* e = new ffb200_flux_engine();
  e->cfg = *cfg;
  e->D = 128 * cfg->num_heads;
  FFB_CHECK(e->D <= 3072, "inner dim above 3072 not supported by ln_modulate");
  int off = 0;
  for (int i = 0; i < cfg->num_layers; ++i) {   // stacked adaLN rows: [norm1 (6D) ; norm1_context (6D)] per dual block
    e->off.n1.push_back(off); off += 6 * e->D;
    e->off.n1c.push_back(off); off += 6 * e->D;
  }
  for (int i = 0; i < cfg->num_single_layers; ++i) { e->off.ns.push_back(off); off += 3 * e->D; }   // norm (3D) per single block
  e->off.out = off; off += 2 * e->D;                                                                 // norm_out (2D)
  e->off.rows = off;
  int r = ffb200_flux_engine_set_weights(e, w);
  if (r) { delete e; return r; }
  *out = e;
  return 0;
}
void ffb200_flux_engine_destroy(ffb200_flux_engine* e) { delete e; }
int ffb200_flux_engine_mod_rows(const ffb200_flux_engine* e) { return e ? e->off.rows : -1; }

void ffb200_flux_plan_destroy(ffb200_flux_plan* p) {
  if (!p) return;
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (void* a : p->allocs) cudaFree(a);
  delete p;
}
long long ffb200_flux_plan_workspace_bytes(const ffb200_flux_plan* p) { return p ? p->ws_bytes : 0; }

int ffb200_flux_plan_create_ex(ffb200_flux_engine* e, int batch, int cfg, int n_img_tokens, int n_text, const float* rope_cos,
                               const float* rope_sin, ffb200_flux_plan** out) {
  FFB_CHECK(e && out && rope_cos && rope_sin, "null argument");
  FFB_CHECK(batch > 0 && batch * (cfg ? 2 : 1) <= 64 && n_img_tokens > 0 && n_text > 0, "batch / token counts out of range");
  FFB_CHECK(!cfg || e->cfg.variant == 1, "a CFG batch exists only for Qwen-Image (FLUX.1 embeds the guidance scale)");
  const ffb200_flux_config& mc = e->cfg;
  ffb200_flux_plan* p = new ffb200_flux_plan();
  p->e = e; p->B = batch; p->cfg = cfg ? 1 : 0; p->Bp = batch * (cfg ? 2 : 1); p->cfg_scale = 1.0f;
  p->Ni = n_img_tokens; p->Nt = n_text; p->S = n_img_tokens + n_text; p->D = e->D;
  p->ws_bytes = 0; p->graph_exec = nullptr; p->graph_valid = false; p->prompts_set = false; p->coef_cap = 0; p->d_coefs = nullptr;
  const int D = p->D, B = p->Bp, Ni = p->Ni, Nt = p->Nt, S = p->S, R = e->off.rows;   // B: FORWARD batch from here on
  const size_t BS = static_cast<size_t>(B) * S;
  int r = 0;
#define ALLOC(field, count, type) if (!r) r = fplan_alloc(p, reinterpret_cast<void**>(&p->field), static_cast<size_t>(count) * sizeof(type))
  ALLOC(xin, static_cast<size_t>(B) * Ni * 64, bf16);
  ALLOC(c0, static_cast<size_t>(B) * Nt * D, bf16);
  if (mc.variant == 1) ALLOC(ctxn, static_cast<size_t>(B) * Nt * mc.joint_attention_dim, bf16);
  ALLOC(tproj, B * 256, bf16); ALLOC(gproj, B * 256, bf16);
  ALLOC(ta, B * D, bf16); ALLOC(ga, B * D, bf16); ALLOC(gemb, B * D, bf16); ALLOC(pa, B * D, bf16); ALLOC(pemb, B * D, bf16);
  ALLOC(temb, B * D, bf16);
  ALLOC(mod, static_cast<size_t>(B) * R, bf16);
  ALLOC(h, BS * D, bf16);
  ALLOC(a1, BS * D, bf16);
  ALLOC(qkv, BS * 3 * D, bf16);
  ALLOC(att, BS * D, bf16);
  ALLOC(ff, BS * 4 * D, bf16);
  if (mc.num_single_layers > 0) ALLOC(cat, BS * 5 * D, bf16);
  ALLOC(vout, static_cast<size_t>(B) * Ni * 64, bf16);
  if (p->cfg) ALLOC(vcfg, static_cast<size_t>(p->B) * Ni * 64, bf16);
  ALLOC(rope_cos, static_cast<size_t>(S) * 128, float);
  ALLOC(rope_sin, static_cast<size_t>(S) * 128, float);
  ALLOC(x_cur, static_cast<size_t>(p->B) * Ni * 64, __half);
  ALLOC(logp_partial, static_cast<size_t>(p->B) * 64, float);
  ALLOC(d_step, 1, int);
  ALLOC(d_gcoef, 1, StepCoef);
  ALLOC(d_txt_len, B, int);
#undef ALLOC
  if (r) { ffb200_flux_plan_destroy(p); return r; }
  {
    std::vector<int> full(B, Nt);
    cudaMemcpy(p->d_txt_len, full.data(), sizeof(int) * B, cudaMemcpyHostToDevice);
  }
  cudaMemcpy(p->rope_cos, rope_cos, static_cast<size_t>(S) * 128 * 4, cudaMemcpyDefault);
  cudaMemcpy(p->rope_sin, rope_sin, static_cast<size_t>(S) * 128 * 4, cudaMemcpyDefault);

  std::vector<Op>& ops = p->fwd_ops;
  const ffb200_flux_weights& w = e->w;
  ffb200_flux_plan* pp = p;
  // ---- timestep embedding: temb = bf16(bf16(t_emb + g_emb) + pooled_emb) (embeddings.py:1612-1624), then every adaLN projection of
  //      the model in one GEMV over the stacked matrix (normalization.py:167, 199, 348)
  const float t_post = mc.variant == 1 ? 1000.0f : 1.0f;   // Qwen-Image: Timesteps(scale=1000) on t/1000 (transformer_qwenimage.py:180)
  ops.push_back([pp, B, t_post](cudaStream_t st) { ++g_launch_count; return launch_timestep_proj(pp->d_coefs, pp->d_step, 0, B, pp->tproj, st, t_post); });
  ops.push_back(small_op(B, p->tproj, 256, w.t1_w, w.t1_b, D, p->ta, nullptr, nullptr, 0));
  if (mc.variant == 1)   // QwenTimestepProjEmbeddings: the timestep embedding alone (176-193)
    ops.push_back(small_op(B, p->ta, D, w.t2_w, w.t2_b, D, p->temb, nullptr, nullptr, 1));
  else
    ops.push_back(small_op(B, p->ta, D, w.t2_w, w.t2_b, D, p->temb, mc.guidance_embeds ? p->gemb : p->pemb,
                           mc.guidance_embeds ? p->pemb : nullptr, 1));
  ops.push_back(small_op(B, p->temb, D, w.mod_w, w.mod_b, R, p->mod, nullptr, nullptr, 1));
  // ---- x_embedder on the packed latents (676), context rows copied from the cached context_embedder output
  {
    const long n = static_cast<long>(p->B) * Ni * 64;
    const int reps = p->cfg ? 2 : 1;   // CFG: both halves of the forward batch see the same latents
    ops.push_back([pp, n, reps](cudaStream_t st) {
      cudaError_t ce = cudaSuccess;
      for (int r2 = 0; r2 < reps && ce == cudaSuccess; ++r2) { ++g_launch_count; ce = launch_cast_f16_to_bf16(pp->x_cur, pp->xin + r2 * n, n, st); }
      return ce;
    });
    GemmSpec s = {p->xin, B, Ni, 0, 64, 64, w.x_w, D, w.x_b, p->h, static_cast<long>(S) * D, Nt, D, EPI_BIAS};
    if ((r = fadd_gemm(p, s))) { ffb200_flux_plan_destroy(p); return r; }
    const size_t row_bytes = static_cast<size_t>(Nt) * D * 2;
    ops.push_back([pp, row_bytes](cudaStream_t st) {
      return cudaMemcpy2DAsync(pp->h, static_cast<size_t>(pp->S) * pp->D * 2, pp->c0, row_bytes, row_bytes, pp->Bp, cudaMemcpyDeviceToDevice, st);
    });
  }
  const long hS = static_cast<long>(S) * D;            // batch stride of the joint buffers
  bf16* h_ctx = p->h; bf16* h_img = p->h + static_cast<size_t>(Nt) * D;
  bf16* a_ctx = p->a1; bf16* a_img = p->a1 + static_cast<size_t>(Nt) * D;
  for (int i = 0; i < mc.num_layers && !r; ++i) {      // ---- dual-stream blocks (438-492)
    const ffb200_flux_dual_weights& L = e->dual[i];
    const bf16* m1 = p->mod + e->off.n1[i];    // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const bf16* mc1 = p->mod + e->off.n1c[i];
    fadd_lnmod(p, h_img, hS, Ni, m1 + 0 * D, m1 + 1 * D, a_img, hS);
    fadd_lnmod(p, h_ctx, hS, Nt, mc1 + 0 * D, mc1 + 1 * D, a_ctx, hS);
    GemmSpec sc = {a_ctx, B, Nt, hS, D, D, L.add_qkv_w, 3 * D, L.add_qkv_b, p->qkv, 3 * hS, 0, 3 * D, EPI_QKV_RMSNORM_ROPE128,
                   nullptr, 0, L.norm_added_q, L.norm_added_k, D, 1e-6f, nullptr, p->rope_cos, p->rope_sin, 0, mc.variant == 1};
    sc.k_scale = engine_prescale() ? 0.08838834764831845f * 1.4426950408889634f : 0.f;
    if ((r = fadd_gemm(p, sc))) break;
    GemmSpec sq = {a_img, B, Ni, hS, D, D, L.qkv_w, 3 * D, L.qkv_b, p->qkv, 3 * hS, Nt, 3 * D, EPI_QKV_RMSNORM_ROPE128,
                   nullptr, 0, L.norm_q, L.norm_k, D, 1e-6f, nullptr, p->rope_cos, p->rope_sin, Nt, mc.variant == 1};
    sq.k_scale = engine_prescale() ? 0.08838834764831845f * 1.4426950408889634f : 0.f;
    if ((r = fadd_gemm(p, sq))) break;
    {
      AttnParams ap;
      if ((r = build_attn(p->qkv, B, S, mc.num_heads, p->att, &ap, 128, 0))) break;
      ap.k_prescaled = engine_prescale() ? 1 : 0;
      if (mc.variant == 1) { ap.kv_mask_lo = p->d_txt_len; ap.kv_mask_hi = Nt; }   // padded text keys (set_text_lengths; default: none)
      ops.push_back([ap](cudaStream_t st) { ++g_launch_count; return launch_attention_d128(ap, st); });
    }
    GemmSpec so = {p->att + static_cast<size_t>(Nt) * D, B, Ni, hS, D, D, L.out_w, D, L.out_b, p->h, hS, Nt, D, EPI_GATE_RESIDUAL,
                   m1 + 2 * D, R};
    if ((r = fadd_gemm(p, so))) break;
    GemmSpec sa = {p->att, B, Nt, hS, D, D, L.add_out_w, D, L.add_out_b, p->h, hS, 0, D, EPI_GATE_RESIDUAL, mc1 + 2 * D, R};
    if ((r = fadd_gemm(p, sa))) break;
    // image MLP
    fadd_lnmod(p, h_img, hS, Ni, m1 + 3 * D, m1 + 4 * D, a_img, hS);
    GemmSpec f1 = {a_img, B, Ni, hS, D, D, L.ff1_w, 4 * D, L.ff1_b, p->ff, static_cast<long>(Ni) * 4 * D, 0, 4 * D, EPI_BIAS_GELU};
    if ((r = fadd_gemm(p, f1))) break;
    GemmSpec f2 = {p->ff, B, Ni, 0, 4 * D, 4 * D, L.ff2_w, D, L.ff2_b, p->h, hS, Nt, D, EPI_GATE_RESIDUAL, m1 + 5 * D, R};
    if ((r = fadd_gemm(p, f2))) break;
    // context MLP
    fadd_lnmod(p, h_ctx, hS, Nt, mc1 + 3 * D, mc1 + 4 * D, a_ctx, hS);
    GemmSpec c1 = {a_ctx, B, Nt, hS, D, D, L.cff1_w, 4 * D, L.cff1_b, p->ff, static_cast<long>(Nt) * 4 * D, 0, 4 * D, EPI_BIAS_GELU};
    if ((r = fadd_gemm(p, c1))) break;
    GemmSpec c2 = {p->ff, B, Nt, 0, 4 * D, 4 * D, L.cff2_w, D, L.cff2_b, p->h, hS, 0, D, EPI_GATE_RESIDUAL, mc1 + 5 * D, R};
    if ((r = fadd_gemm(p, c2))) break;
  }
  for (int i = 0; i < mc.num_single_layers && !r; ++i) {   // ---- single-stream blocks (378-407) on the joint [text ; image] rows
    const ffb200_flux_single_weights& L = e->single[i];
    const bf16* ms = p->mod + e->off.ns[i];   // shift_msa, scale_msa, gate
    fadd_lnmod(p, p->h, 0, S, ms + 0 * D, ms + 1 * D, p->a1, 0);
    GemmSpec sq = {p->a1, B, S, 0, D, D, L.qkv_w, 3 * D, L.qkv_b, p->qkv, 3 * hS, 0, 3 * D, EPI_QKV_RMSNORM_ROPE128,
                   nullptr, 0, L.norm_q, L.norm_k, D, 1e-6f, nullptr, p->rope_cos, p->rope_sin, 0};
    sq.k_scale = engine_prescale() ? 0.08838834764831845f * 1.4426950408889634f : 0.f;
    if ((r = fadd_gemm(p, sq))) break;
    // proj_mlp + GELU -> columns [D, 5D) of the cat buffer ; attention -> columns [0, D)
    GemmSpec sm = {p->a1, B, S, 0, D, D, L.mlp_w, 4 * D, L.mlp_b, p->cat + D, 5 * hS, 0, 5 * D, EPI_BIAS_GELU};
    if ((r = fadd_gemm(p, sm))) break;
    {
      AttnParams ap;
      if ((r = build_attn(p->qkv, B, S, mc.num_heads, p->cat, &ap, 128, 5 * D))) break;
      ap.k_prescaled = engine_prescale() ? 1 : 0;
      ops.push_back([ap](cudaStream_t st) { ++g_launch_count; return launch_attention_d128(ap, st); });
    }
    GemmSpec so = {p->cat, B, S, 0, 5 * D, 5 * D, L.out_w, D, L.out_b, p->h, hS, 0, D, EPI_GATE_RESIDUAL, ms + 2 * D, R};
    if ((r = fadd_gemm(p, so))) break;
  }
  if (!r) {   // ---- norm_out (AdaLayerNormContinuous: scale, shift) on the image rows + proj_out (770-771)
    const bf16* mo = p->mod + e->off.out;
    fadd_lnmod(p, h_img, hS, Ni, mo + 1 * D, mo + 0 * D, p->a1, 0);
    GemmSpec po = {p->a1, B, Ni, 0, D, D, w.proj_w, 64, w.proj_b, p->vout, static_cast<long>(Ni) * 64, 0, 64, EPI_BIAS};
    r = fadd_gemm(p, po);
    if (!r && p->cfg) {   // true CFG + per-token norm rescale (FF/models/qwen_image/qwen_image.py:580-587)
      const long toks = static_cast<long>(p->B) * Ni;
      ops.push_back([pp, toks](cudaStream_t st) { ++g_launch_count; return launch_cfg_norm_rescale(pp->vout, pp->vcfg, toks, pp->cfg_scale, st); });
    }
  }
  if (r) { ffb200_flux_plan_destroy(p); return r; }
  *out = p;
  return 0;
}

int ffb200_flux_plan_create(ffb200_flux_engine* e, int batch, int n_img_tokens, int n_text, const float* rope_cos,
                            const float* rope_sin, ffb200_flux_plan** out) {
  return ffb200_flux_plan_create_ex(e, batch, 0, n_img_tokens, n_text, rope_cos, rope_sin, out);
}

static int fensure_coefs(ffb200_flux_plan* p, int n) {
  if (n <= p->coef_cap) return 0;
  StepCoef* d = nullptr;
  int r = fplan_alloc(p, reinterpret_cast<void**>(&d), static_cast<size_t>(n) * sizeof(StepCoef));
  if (r) return r;
  p->d_coefs = d; p->coef_cap = n;
  p->graph_valid = false;
  return 0;
}

int ffb200_flux_set_prompts(ffb200_flux_plan* p, const void* prompt_embeds_bf16, const void* pooled_bf16, float guidance_model,
                            void* stream) {
  FFB_CHECK(p && prompt_embeds_bf16, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ffb200_flux_engine* e = p->e;
  const int J = e->cfg.joint_attention_dim, P = e->cfg.pooled_projection_dim, D = p->D, B = p->Bp;
  g_launch_count = 0;
  const void* ctx_in = prompt_embeds_bf16;
  if (e->cfg.variant == 1) {   // Qwen-Image: txt_norm (diffusers RMSNorm, eps 1e-6) before txt_in (transformer_qwenimage.py:936-937)
    ++g_launch_count;
    FFB_CUDA(launch_rms_norm_rows(static_cast<const bf16*>(prompt_embeds_bf16), static_cast<const bf16*>(e->w.ctxn_w), p->ctxn,
                                  static_cast<long>(B) * p->Nt, J, 1e-6f, st));
    ctx_in = p->ctxn;
    p->cfg_scale = guidance_model;   // the true-CFG scale of the rollout (no embedded guidance in this model)
  }
  // context_embedder / txt_in (transformer_flux.py:686) - timestep independent, cached for the whole rollout
  GemmSpec s = {ctx_in, B, p->Nt, 0, J, J, e->w.ctx_w, D, e->w.ctx_b, p->c0, static_cast<long>(p->Nt) * D, 0, D, EPI_BIAS};
  GemmParams gp;
  int r = build_gemm(s, &gp);
  if (r) return r;
  ++g_launch_count;
  FFB_CUDA(launch_gemm(gp, num_sms(), st));
  if (e->cfg.variant == 0) {
    FFB_CHECK(pooled_bf16, "FLUX.1 needs the pooled prompt embedding");
    // pooled-text MLP and (dev checkpoints) the guidance embedding: both constant over the rollout (embeddings.py:1616-1621)
    FFB_CUDA(small_op(B, static_cast<const bf16*>(pooled_bf16), P, e->w.p1_w, e->w.p1_b, D, p->pa, nullptr, nullptr, 0)(st));
    FFB_CUDA(small_op(B, p->pa, D, e->w.p2_w, e->w.p2_b, D, p->pemb, nullptr, nullptr, 1)(st));
    if (e->cfg.guidance_embeds) {
      StepCoef c; memset(&c, 0, sizeof(c)); c.t_model = guidance_model;   // bf16(guidance) * 1000 in bf16, computed by the host
      FFB_CUDA(cudaMemcpyAsync(p->d_gcoef, &c, sizeof(c), cudaMemcpyHostToDevice, st));
      ++g_launch_count;
      FFB_CUDA(launch_timestep_proj(p->d_gcoef, nullptr, 0, B, p->gproj, st));
      FFB_CUDA(small_op(B, p->gproj, 256, e->w.g1_w, e->w.g1_b, D, p->ga, nullptr, nullptr, 0)(st));
      FFB_CUDA(small_op(B, p->ga, D, e->w.g2_w, e->w.g2_b, D, p->gemb, nullptr, nullptr, 1)(st));
    }
  }
  p->prompts_set = true;
  p->graph_valid = p->graph_valid && e->cfg.variant == 0;   // the CFG scale is baked into the captured graph
  return 0;
}

int ffb200_flux_set_text_lengths(ffb200_flux_plan* p, const int* lengths_host, void* stream) {
  FFB_CHECK(p && lengths_host, "null argument");
  FFB_CHECK(p->e->cfg.variant == 1, "key-padding masks exist only on the Qwen-Image path");
  for (int i = 0; i < p->Bp; ++i) FFB_CHECK(lengths_host[i] >= 1 && lengths_host[i] <= p->Nt, "text length out of range");
  FFB_CUDA(cudaMemcpyAsync(p->d_txt_len, lengths_host, sizeof(int) * p->Bp, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  FFB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));   // lengths_host may be a temporary
  return 0;
}

static int frun_forward(ffb200_flux_plan* p, cudaStream_t st) {
  for (auto& op : p->fwd_ops) {
    cudaError_t e = op(st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "flux forward launch");
  }
  return 0;
}

static void ffill_sde(const ffb200_flux_plan* p, SdeStepParams* sp) {
  memset(sp, 0, sizeof(*sp));
  // the packed latents [B, Ni, 64] are one flat "image" per sample for the elementwise scheduler step
  sp->B = p->B; sp->C = 1; sp->H = p->Ni; sp->W = 64; sp->patch = 1; sp->cfg = 0; sp->guidance = 1.0f;
  sp->v_direct = p->cfg ? p->vcfg : p->vout; sp->x = p->x_cur; sp->logp_partial = p->logp_partial; sp->coef_table = p->d_coefs;
}

int ffb200_flux_forward(ffb200_flux_plan* p, const void* latents_fp16, float t_model, void* noise_pred_bf16, void* stream) {
  FFB_CHECK(p && latents_fp16, "null argument");
  FFB_CHECK(p->prompts_set, "ffb200_flux_set_prompts must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  int r = fensure_coefs(p, 1);
  if (r) return r;
  StepCoef c; memset(&c, 0, sizeof(c)); c.t_model = t_model;
  const size_t lat_bytes = static_cast<size_t>(p->B) * p->Ni * 64 * 2;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, &c, sizeof(c), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, latents_fp16, lat_bytes, cudaMemcpyDeviceToDevice, st));
  if ((r = frun_forward(p, st))) return r;
  if (noise_pred_bf16) FFB_CUDA(cudaMemcpyAsync(noise_pred_bf16, p->cfg ? p->vcfg : p->vout, lat_bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int ffb200_flux_step(ffb200_flux_plan* p, const ffb200_step_args* a, void* stream) {
  FFB_CHECK(p && a && a->latents, "null argument");
  FFB_CHECK(p->prompts_set, "ffb200_flux_set_prompts must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  int r = fensure_coefs(p, 1);
  if (r) return r;
  const size_t lat_bytes = static_cast<size_t>(p->B) * p->Ni * 64 * 2;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, &a->coef, sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, a->latents, lat_bytes, cudaMemcpyDeviceToDevice, st));
  if ((r = frun_forward(p, st))) return r;
  SdeStepParams sp; ffill_sde(p, &sp);
  sp.noise = a->noise; sp.seed = a->seed; sp.coef_index = 0;
  sp.next_given = static_cast<const __half*>(a->next_latents);
  sp.x_next = static_cast<__half*>(a->out_next_latents);
  sp.mean_out = a->out_mean; sp.log_prob = a->out_log_prob; sp.v_out = static_cast<bf16*>(a->out_noise_pred);
  sp.overflow_flag = a->overflow_flag;
  g_launch_count += 2;
  FFB_CUDA(launch_sde_step(sp, st));
  return 0;
}

int ffb200_flux_rollout(ffb200_flux_plan* p, const ffb200_rollout_args* a, void* stream) {
  FFB_CHECK(p && a && a->coefs && a->x0 && a->num_steps > 0, "bad rollout arguments");
  FFB_CHECK(p->prompts_set, "ffb200_flux_set_prompts must be called first");
  FFB_CHECK(!(a->use_graph && stream == nullptr), "use_graph needs a non-default stream (stream capture)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  const int T = a->num_steps;
  const int chw = p->Ni * 64;
  const size_t lat_elems = static_cast<size_t>(p->B) * chw;
  int r = fensure_coefs(p, T);
  if (r) return r;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, a->coefs, static_cast<size_t>(T) * sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, a->x0, lat_elems * 2, cudaMemcpyDeviceToDevice, st));
  if (a->all_latents && a->store_initial_slot >= 0) {
    FFB_CUDA(cudaMemcpy2DAsync(static_cast<__half*>(a->all_latents) + static_cast<size_t>(a->store_initial_slot) * chw,
                               static_cast<size_t>(a->n_latent_slots) * chw * 2, a->x0, static_cast<size_t>(chw) * 2,
                               static_cast<size_t>(chw) * 2, p->B, cudaMemcpyDeviceToDevice, st));
  }
  SdeStepParams sp; ffill_sde(p, &sp);
  sp.noise = a->noise; sp.noise_step_stride = static_cast<long>(lat_elems); sp.seed = a->seed;
  sp.x_next = p->x_cur;
  sp.traj = static_cast<__half*>(a->all_latents); sp.traj_batch_stride = static_cast<long>(a->n_latent_slots) * chw;
  sp.logp_traj = a->log_probs; sp.logp_batch_stride = a->n_logp_slots;
  sp.overflow_flag = a->overflow_flag; sp.step_ptr = p->d_step;
  auto one_step = [&](cudaStream_t s) -> int {
    int rr = frun_forward(p, s);
    if (rr) return rr;
    g_launch_count += 2;
    cudaError_t e = launch_sde_step(sp, s);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "sde_step launch");
    return 0;
  };
  if (a->use_graph) {
    const bool same = p->graph_valid && memcmp(&p->graph_sde, &sp, sizeof(sp)) == 0;
    if (!same) {
      const long long before = g_launch_count;
      const int rr = graph_capture_or_update(&p->graph_exec, st, one_step);
      const long long per_step = g_launch_count - before;
      g_launch_count = before;
      if (rr) { p->graph_valid = false; return rr; }
      p->graph_sde = sp; p->graph_valid = true; p->graph_launches = per_step;
    }
    for (int i = 0; i < T; ++i) FFB_CUDA(cudaGraphLaunch(p->graph_exec, st));
    g_launch_count += p->graph_launches * T;
  } else {
    for (int i = 0; i < T; ++i) {
      int rr = one_step(st);
      if (rr) return rr;
    }
  }
  if (a->final_latents) FFB_CUDA(cudaMemcpyAsync(a->final_latents, p->x_cur, lat_elems * 2, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
