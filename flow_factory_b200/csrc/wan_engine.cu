// Host side of the Wan2.1 T2V rollout path (SURVEY.md 8f row 4 / BASELINE config 4): WanTransformer3DModel.forward + the denoise loop of
// Wan2_T2V_Adapter as one launch list per step, built from the tcgen05 GEMM / head-dim-128 attention kernels of the other engines plus
// the elementwise kernels of wan_elementwise.cu.
//
//   WanTransformer3DModel.forward     DF/models/transformers/transformer_wan.py:629-740
//   WanTransformerBlock               462-505      WanAttnProcessor 78-162      WanRotaryPosEmbed 354-417
//   WanTimeTextImageEmbedding         330-351
//   Wan2_T2V_Adapter.forward          FF/models/wan/wan2_t2v.py:425-543 (true CFG: two forwards -> here one batch of 2B, negatives first)
//   UniPCMultistepSDEScheduler.step   FF/scheduler/unipc_multistep.py:290-421 (the Euler / SDE arithmetic of the other models)
//
// Per block: fp32 LayerNorm + fp32 modulate -> fused q|k|v GEMM -> RMSNorm across heads + 3-D RoPE in place on the q / k column blocks
// -> self-attention over the S = F'H'W' video tokens -> out-projection GEMM + fp32 gated residual; affine fp32 LayerNorm -> q GEMM +
// RMSNorm -> cross-attention to the text tokens (keys / values of all blocks are computed once per prompt set: they do not depend on
// the timestep) -> out-projection with the residual in the GEMM epilogue; fp32 LayerNorm + modulate -> GELU-tanh FFN -> gated residual.
//
// Validated on B200 in round 2 (tests/test_gpu_wan.py).
#include "common.cuh"
#include "kernels.h"
#include "../../include/ffb200.h"

struct ffb200_wan_engine {
  ffb200_wan_config cfg;
  ffb200_wan_weights w;
  std::vector<ffb200_wan_layer_weights> layers;
  int D;
};

struct ffb200_wan_plan {
  ffb200_wan_engine* e;
  int B, Bp, cfg, F, H, W, fp, hp, wp, S, Nt, D, Kpe;
  std::vector<void*> allocs;
  long long ws_bytes;
  bf16 *peA, *h, *a1, *qkv, *att, *y, *q2, *ff, *vout, *ctx_a, *ctx, *kv2, *tproj, *ta, *temb, *temb6, *fin, *ones;
  float *mod, *rope_cos, *rope_sin;
  const bf16** d_tables;
  __half* x_cur;
  float* logp_partial;
  int* d_step;
  StepCoef* d_coefs; int coef_cap;
  std::vector<Op> fwd_ops;
  bool prompts_set;
  cudaGraphExec_t graph_exec; SdeStepParams graph_sde; bool graph_valid; long long graph_launches;
};

static int wplan_alloc(ffb200_wan_plan* p, void** ptr, size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaMalloc(wan workspace)");
  p->allocs.push_back(*ptr);
  p->ws_bytes += static_cast<long long>(bytes);
  return 0;
}
static int wadd_gemm(std::vector<Op>& ops, const GemmSpec& s) {
  GemmParams gp;
  int r = build_gemm(s, &gp);
  if (r) return r;
  const int sms = num_sms();
  ops.push_back([gp, sms](cudaStream_t st) { ++g_launch_count; return launch_gemm(gp, sms, st); });
  return 0;
}
// self-attention over the fused [Bp, S, 3D] buffer, or (kv != nullptr) cross-attention of q [Bp, S, D] against kv [Bp, Nt, 2D]
static int wbuild_attn(const ffb200_wan_plan* p, const bf16* q, int q_ld, const bf16* kv, int kv_ld, int kv_len, bf16* out, AttnParams* ap) {
  memset(ap, 0, sizeof(*ap));
  const int D = p->D, heads = D / 128;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(q_ld), static_cast<uint64_t>(p->S), static_cast<uint64_t>(p->Bp)};
    const uint64_t str[2] = {static_cast<uint64_t>(q_ld), static_cast<uint64_t>(p->S) * q_ld};
    const uint32_t box[3] = {64, 128, 1};
    int r = make_tmap(&ap->tmQKV, q, 3, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(kv_ld), static_cast<uint64_t>(kv_len), static_cast<uint64_t>(p->Bp)};
    const uint64_t str[2] = {static_cast<uint64_t>(kv_ld), static_cast<uint64_t>(kv_len) * kv_ld};
    const uint32_t box[3] = {64, 64, 1};      // A128_BN kv rows per tile
    int r = make_tmap(&ap->tmKV, kv, 3, dims, str, box);
    if (r) return r;
  }
  ap->seq_len = p->S; ap->num_heads = heads; ap->inner_dim = D; ap->batch = p->Bp;
  ap->out = out; ap->out_row_stride = D; ap->out_batch_stride = static_cast<long>(p->S) * D;
  ap->scale_log2 = (1.0f / sqrtf(128.0f)) * 1.4426950408889634f;
  ap->kv_len = kv_len; ap->q_col = 0; ap->k_col = 0; ap->v_col = D;
  return 0;
}

static int wan_build_forward(ffb200_wan_plan* p) {
  const ffb200_wan_engine* e = p->e;
  const ffb200_wan_config& c = e->cfg;
  const int D = p->D, Bp = p->Bp, S = p->S, Nt = p->Nt, L = c.num_layers, ffn = c.ffn_dim;
  auto& ops = p->fwd_ops;
  ops.clear();
  int r;
  ffb200_wan_plan* pp = p;
  // ---- patch embedding (Conv3d kernel = stride = patch -> im2col + GEMM, transformer_wan.py:668-669) ----
  ops.push_back([pp](cudaStream_t st) {
    ++g_launch_count;
    const ffb200_wan_config& cc = pp->e->cfg;
    return launch_wan_patchify(pp->x_cur, pp->B, pp->Bp / pp->B, cc.in_channels, pp->F, pp->H, pp->W, cc.patch_t, cc.patch_h, cc.patch_w, pp->peA, st);
  });
  {
    GemmSpec s = {p->peA, Bp, S, 0, p->Kpe, p->Kpe, e->w.pe_w, D, e->w.pe_b, p->h, static_cast<long>(S) * D, 0, D, EPI_BIAS};
    if ((r = wadd_gemm(ops, s))) return r;
  }
  // ---- condition embedder (330-351): timestep sinusoid -> MLP -> temb ; time_proj(silu(temb)) -> 6 D ----
  ops.push_back([pp](cudaStream_t st) { ++g_launch_count; return launch_timestep_proj(pp->d_coefs, pp->d_step, 0, pp->Bp, pp->tproj, st); });
  ops.push_back(small_op(Bp, p->tproj, c.freq_dim, e->w.t1_w, e->w.t1_b, D, p->ta, nullptr, nullptr, 0));
  ops.push_back(small_op(Bp, p->ta, D, e->w.t2_w, e->w.t2_b, D, p->temb, nullptr, nullptr, 1));
  ops.push_back(small_op(Bp, p->temb, D, e->w.tp_w, e->w.tp_b, 6 * D, p->temb6, nullptr, nullptr, 1));
  ops.push_back([pp, L](cudaStream_t st) { ++g_launch_count; return launch_wan_mod_vectors(pp->d_tables, pp->temb6, pp->mod, L, pp->Bp, 6 * pp->D, st); });
  ops.push_back([pp](cudaStream_t st) {
    ++g_launch_count;
    return launch_wan_final_mod(static_cast<const bf16*>(pp->e->w.table), pp->temb, pp->fin, pp->Bp, pp->D, st);
  });
  const long rows = static_cast<long>(Bp) * S;
  const long mod_bs = 6L * D;                 // fp32 elements between the batch rows of one block's modulation
  auto ln_mod = [&](int l, int shift_j, int scale_j) {
    WanLnParams lp{};
    lp.x = p->h; lp.out = p->a1; lp.rows_per_batch = S; lp.num_batch = Bp; lp.D = D; lp.eps = c.eps; lp.mode = 0;
    lp.shift = p->mod + (static_cast<long>(l) * Bp) * mod_bs + static_cast<long>(shift_j) * D;
    lp.scale = p->mod + (static_cast<long>(l) * Bp) * mod_bs + static_cast<long>(scale_j) * D;
    lp.mod_batch_stride = mod_bs;
    ops.push_back([lp](cudaStream_t st) { ++g_launch_count; return launch_wan_ln(lp, st); });
  };
  auto gate_res = [&](int l, int gate_j) {
    bf16* h = p->h; const bf16* y = p->y;
    const float* gate = p->mod + (static_cast<long>(l) * Bp) * mod_bs + static_cast<long>(gate_j) * D;
    ops.push_back([h, y, gate, mod_bs, Bp, S, D](cudaStream_t st) { ++g_launch_count; return launch_wan_gate_residual(h, y, gate, mod_bs, Bp, S, D, st); });
  };
  for (int l = 0; l < L; ++l) {
    const ffb200_wan_layer_weights& w = e->layers[l];
    // ---- self-attention (474-489) ----
    ln_mod(l, 0, 1);
    {
      GemmSpec s = {p->a1, Bp, S, 0, D, D, w.qkv_w, 3 * D, w.qkv_b, p->qkv, static_cast<long>(S) * 3 * D, 0, 3 * D, EPI_BIAS};
      if ((r = wadd_gemm(ops, s))) return r;
    }
    for (int which = 0; which < 2; ++which) {
      bf16* x = p->qkv + which * D;
      const bf16* nw = static_cast<const bf16*>(which == 0 ? w.norm_q : w.norm_k);
      const float *cs = p->rope_cos, *sn = p->rope_sin; const float eps = c.eps;
      const float ks = (which == 1 && engine_prescale()) ? 0.08838834764831845f * 1.4426950408889634f : 1.0f;      // keys carry the softmax scale
      ops.push_back([x, rows, S, D, nw, eps, cs, sn, ks](cudaStream_t st) { ++g_launch_count; return launch_wan_rms_rope(x, rows, S, 3 * D, D, nw, eps, cs, sn, st, ks); });
    }
    {
      AttnParams ap;
      if ((r = build_attn(p->qkv, Bp, S, D / 128, p->att, &ap, 128, 0))) return r;
      ap.k_prescaled = engine_prescale() ? 1 : 0;
      ops.push_back([ap](cudaStream_t st) { ++g_launch_count; return launch_attention_d128(ap, st); });
    }
    {
      GemmSpec s = {p->att, Bp, S, 0, D, D, w.out_w, D, w.out_b, p->y, static_cast<long>(S) * D, 0, D, EPI_BIAS};
      if ((r = wadd_gemm(ops, s))) return r;
    }
    gate_res(l, 2);
    // ---- cross-attention (491-497) ----
    {
      WanLnParams lp{};
      lp.x = p->h; lp.out = p->a1; lp.rows_per_batch = S; lp.num_batch = Bp; lp.D = D; lp.eps = c.eps; lp.mode = 1;
      lp.weight = static_cast<const bf16*>(w.norm2_w); lp.bias = static_cast<const bf16*>(w.norm2_b);
      ops.push_back([lp](cudaStream_t st) { ++g_launch_count; return launch_wan_ln(lp, st); });
    }
    {
      GemmSpec s = {p->a1, Bp, S, 0, D, D, w.q2_w, D, w.q2_b, p->q2, static_cast<long>(S) * D, 0, D, EPI_BIAS};
      if ((r = wadd_gemm(ops, s))) return r;
    }
    {
      bf16* x = p->q2; const bf16* nw = static_cast<const bf16*>(w.norm_q2); const float eps = c.eps;
      ops.push_back([x, rows, S, D, nw, eps](cudaStream_t st) { ++g_launch_count; return launch_wan_rms_rope(x, rows, S, D, D, nw, eps, nullptr, nullptr, st); });
    }
    {
      AttnParams ap;
      const bf16* kv = p->kv2 + static_cast<long>(l) * Bp * Nt * 2 * D;
      if ((r = wbuild_attn(p, p->q2, D, kv, 2 * D, Nt, p->att, &ap))) return r;
      ap.k_prescaled = engine_prescale() ? 1 : 0;
      ops.push_back([ap](cudaStream_t st) { ++g_launch_count; return launch_attention_d128_cross(ap, st); });
    }
    {
      // hidden_states = hidden_states + attn_output (497): the bf16 residual of the GEMM epilogue with a gate of ones
      GemmSpec s = {p->att, Bp, S, 0, D, D, w.out2_w, D, w.out2_b, p->h, static_cast<long>(S) * D, 0, D, EPI_GATE_RESIDUAL, p->ones, 0};
      if ((r = wadd_gemm(ops, s))) return r;
    }
    // ---- feed-forward (499-503) ----
    ln_mod(l, 3, 4);
    {
      GemmSpec s = {p->a1, Bp, S, 0, D, D, w.ff1_w, ffn, w.ff1_b, p->ff, static_cast<long>(S) * ffn, 0, ffn, EPI_BIAS_GELU};
      if ((r = wadd_gemm(ops, s))) return r;
    }
    {
      GemmSpec s = {p->ff, Bp, S, 0, ffn, ffn, w.ff2_w, D, w.ff2_b, p->y, static_cast<long>(S) * D, 0, D, EPI_BIAS};
      if ((r = wadd_gemm(ops, s))) return r;
    }
    gate_res(l, 5);
  }
  // ---- output norm + projection (715-731): shift / scale stay in the module dtype -> the bf16 modulate kernel of the other engines ----
  {
    LnModParams lp{};
    lp.x = p->h; lp.rows_per_batch = S; lp.num_batch = Bp; lp.D = D; lp.eps = c.eps;
    lp.shift1 = p->fin; lp.scale1 = p->fin + D; lp.out1 = p->a1; lp.mod_batch_stride = 2L * D;
    ops.push_back([lp](cudaStream_t st) { ++g_launch_count; return launch_ln_modulate(lp, st); });
  }
  {
    GemmSpec s = {p->a1, Bp, S, 0, D, D, e->w.proj_w, p->Kpe, e->w.proj_b, p->vout, static_cast<long>(S) * p->Kpe, 0, p->Kpe, EPI_BIAS};
    if ((r = wadd_gemm(ops, s))) return r;
  }
  return 0;
}

static int wensure_coefs(ffb200_wan_plan* p, int n) {
  if (n <= p->coef_cap) return 0;
  StepCoef* d = nullptr;
  int r = wplan_alloc(p, reinterpret_cast<void**>(&d), static_cast<size_t>(n) * sizeof(StepCoef));
  if (r) return r;
  p->d_coefs = d; p->coef_cap = n;
  p->graph_valid = false;
  return 0;
}
static int wrun_forward(ffb200_wan_plan* p, cudaStream_t st) {
  for (auto& op : p->fwd_ops) {
    cudaError_t e = op(st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "wan forward launch");
  }
  return 0;
}
static void wfill_sde(const ffb200_wan_plan* p, float guidance, SdeStepParams* sp) {
  memset(sp, 0, sizeof(*sp));
  // [C, F, H, W] is one [C, F*H, W] image for the elementwise step; token (f', h', w') = row f'*H' + h' of the token grid (patch_t = 1)
  sp->B = p->B; sp->C = p->e->cfg.in_channels; sp->H = p->F * p->H; sp->W = p->W; sp->patch = p->e->cfg.patch_h;
  sp->cfg = p->cfg; sp->guidance = guidance;
  sp->v_tokens = p->vout; sp->x = p->x_cur; sp->logp_partial = p->logp_partial; sp->coef_table = p->d_coefs;
}

extern "C" {

int ffb200_wan_engine_set_weights(ffb200_wan_engine* e, const ffb200_wan_weights* w) {
  FFB_CHECK(e && w && w->layers, "null argument");
  e->w = *w;
  e->layers.assign(w->layers, w->layers + e->cfg.num_layers);
  e->w.layers = e->layers.data();
  const void* const* g = reinterpret_cast<const void* const*>(&e->w);
  for (size_t i = 0; i < (sizeof(ffb200_wan_weights) - sizeof(void*)) / sizeof(void*); ++i) FFB_CHECK(g[i] != nullptr, "null global weight pointer");
  for (const auto& l : e->layers) {
    const void* const* q = reinterpret_cast<const void* const*>(&l);
    for (size_t i = 0; i < sizeof(ffb200_wan_layer_weights) / sizeof(void*); ++i) FFB_CHECK(q[i] != nullptr, "null layer weight pointer");
  }
  return 0;
}

int ffb200_wan_engine_create(const ffb200_wan_config* cfg, const ffb200_wan_weights* w, ffb200_wan_engine** out) {
  FFB_CHECK(cfg && w && out, "null argument");
  FFB_CHECK(cfg->num_layers >= 1 && cfg->num_heads >= 1, "wan: num_layers / num_heads");
  FFB_CHECK(cfg->patch_t == 1 && cfg->patch_h == 2 && cfg->patch_w == 2, "wan: patch size must be (1, 2, 2)");
  FFB_CHECK((cfg->in_channels * 4) % 64 == 0, "wan: in_channels * patch volume must be a multiple of 64");
  FFB_CHECK(cfg->num_heads * 128 <= 3072, "wan: inner_dim up to 3072");
  FFB_CHECK(cfg->freq_dim == 256 && cfg->text_dim % 8 == 0 && cfg->ffn_dim % 64 == 0, "wan: freq_dim 256, text_dim % 8, ffn_dim % 64");
  {
    int dev = 0, major = 0;
    FFB_CUDA(cudaGetDevice(&dev));
    FFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    FFB_CHECK(major == 10, "this library only runs on compute capability 10.x (B200)");
  }
  ffb200_wan_engine* e = new ffb200_wan_engine();
  e->cfg = *cfg; e->D = cfg->num_heads * 128;
  int r = ffb200_wan_engine_set_weights(e, w);
  if (r) { delete e; return r; }
  *out = e;
  return 0;
}
void ffb200_wan_engine_destroy(ffb200_wan_engine* e) { delete e; }

void ffb200_wan_plan_destroy(ffb200_wan_plan* p) {
  if (!p) return;
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (void* q : p->allocs) cudaFree(q);
  delete p;
}
long long ffb200_wan_plan_workspace_bytes(const ffb200_wan_plan* p) { return p ? p->ws_bytes : 0; }

int ffb200_wan_plan_create(ffb200_wan_engine* e, int batch, int cfg, int frames, int height, int width, int n_text, const float* rope_cos,
                           const float* rope_sin, ffb200_wan_plan** out) {
  FFB_CHECK(e && out && rope_cos && rope_sin, "null argument");
  FFB_CHECK(batch >= 1 && frames >= 1 && n_text >= 1 && (cfg == 0 || cfg == 1), "wan plan: bad geometry");
  FFB_CHECK(height % 2 == 0 && width % 4 == 0, "wan plan: latent height must be even and width a multiple of 4");
  ffb200_wan_plan* p = new ffb200_wan_plan();
  p->e = e; p->B = batch; p->cfg = cfg; p->Bp = batch * (cfg ? 2 : 1);
  p->F = frames; p->H = height; p->W = width; p->fp = frames; p->hp = height / 2; p->wp = width / 2;
  p->S = p->fp * p->hp * p->wp; p->Nt = n_text; p->D = e->D; p->Kpe = e->cfg.in_channels * 4;
  const int D = p->D, Bp = p->Bp, S = p->S, L = e->cfg.num_layers, ffn = e->cfg.ffn_dim;
  auto bail = [&](int code) { ffb200_wan_plan_destroy(p); return code; };
  int r;
#define WALLOC(field, elems, type)                                                                      \
  if ((r = wplan_alloc(p, reinterpret_cast<void**>(&p->field), static_cast<size_t>(elems) * sizeof(type)))) return bail(r)
  const size_t rows = static_cast<size_t>(Bp) * S;
  WALLOC(peA, rows * p->Kpe, bf16); WALLOC(h, rows * D, bf16); WALLOC(a1, rows * D, bf16); WALLOC(qkv, rows * 3 * D, bf16);
  WALLOC(att, rows * D, bf16); WALLOC(y, rows * D, bf16); WALLOC(q2, rows * D, bf16); WALLOC(ff, rows * ffn, bf16);
  WALLOC(vout, rows * p->Kpe, bf16);
  WALLOC(ctx_a, static_cast<size_t>(Bp) * n_text * D, bf16); WALLOC(ctx, static_cast<size_t>(Bp) * n_text * D, bf16);
  WALLOC(kv2, static_cast<size_t>(L) * Bp * n_text * 2 * D, bf16);
  WALLOC(tproj, static_cast<size_t>(Bp) * 256, bf16); WALLOC(ta, static_cast<size_t>(Bp) * D, bf16); WALLOC(temb, static_cast<size_t>(Bp) * D, bf16);
  WALLOC(temb6, static_cast<size_t>(Bp) * 6 * D, bf16); WALLOC(fin, static_cast<size_t>(Bp) * 2 * D, bf16); WALLOC(ones, D, bf16);
  WALLOC(mod, static_cast<size_t>(L) * Bp * 6 * D, float);
  WALLOC(rope_cos, static_cast<size_t>(S) * 128, float); WALLOC(rope_sin, static_cast<size_t>(S) * 128, float);
  WALLOC(d_tables, L, const bf16*);
  WALLOC(x_cur, static_cast<size_t>(batch) * e->cfg.in_channels * frames * height * width, __half);
  WALLOC(logp_partial, static_cast<size_t>(batch) * 64, float);
  WALLOC(d_step, 1, int);
#undef WALLOC
  if ((r = wensure_coefs(p, 64))) return bail(r);
  {
    cudaError_t ce = cudaMemcpy(p->rope_cos, rope_cos, static_cast<size_t>(S) * 128 * 4, cudaMemcpyDefault);
    if (ce == cudaSuccess) ce = cudaMemcpy(p->rope_sin, rope_sin, static_cast<size_t>(S) * 128 * 4, cudaMemcpyDefault);
    std::vector<const bf16*> tabs(L);
    for (int l = 0; l < L; ++l) tabs[l] = static_cast<const bf16*>(e->layers[l].table);
    if (ce == cudaSuccess) ce = cudaMemcpy(p->d_tables, tabs.data(), sizeof(const bf16*) * L, cudaMemcpyHostToDevice);
    std::vector<uint16_t> one(D, 0x3F80);     // bf16 1.0
    if (ce == cudaSuccess) ce = cudaMemcpy(p->ones, one.data(), static_cast<size_t>(D) * 2, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemset(p->d_step, 0, sizeof(int));
    if (ce != cudaSuccess) return bail(fail(static_cast<int>(ce), "wan plan: table upload"));
  }
  if ((r = wan_build_forward(p))) return bail(r);
  *out = p;
  return 0;
}

int ffb200_wan_set_prompts(ffb200_wan_plan* p, const void* prompt_embeds_bf16, void* stream) {
  FFB_CHECK(p && prompt_embeds_bf16, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ffb200_wan_engine* e = p->e;
  const int D = p->D, Bp = p->Bp, Nt = p->Nt, T = e->cfg.text_dim, L = e->cfg.num_layers;
  g_launch_count = 0;
  std::vector<Op> ops;
  int r;
  // text embedder (PixArtAlphaTextProjection, gelu_tanh): linear_1 -> GELU -> linear_2   (transformer_wan.py:346)
  {
    GemmSpec s = {prompt_embeds_bf16, Bp, Nt, 0, T, T, e->w.x1_w, D, e->w.x1_b, p->ctx_a, static_cast<long>(Nt) * D, 0, D, EPI_BIAS_GELU};
    if ((r = wadd_gemm(ops, s))) return r;
  }
  {
    GemmSpec s = {p->ctx_a, Bp, Nt, 0, D, D, e->w.x2_w, D, e->w.x2_b, p->ctx, static_cast<long>(Nt) * D, 0, D, EPI_BIAS};
    if ((r = wadd_gemm(ops, s))) return r;
  }
  // cross-attention keys / values of every block: to_k | to_v of the text states, RMSNorm across heads on k (no RoPE)
  for (int l = 0; l < L; ++l) {
    const ffb200_wan_layer_weights& w = e->layers[l];
    bf16* kv = p->kv2 + static_cast<long>(l) * Bp * Nt * 2 * D;
    GemmSpec s = {p->ctx, Bp, Nt, 0, D, D, w.kv2_w, 2 * D, w.kv2_b, kv, static_cast<long>(Nt) * 2 * D, 0, 2 * D, EPI_BIAS};
    if ((r = wadd_gemm(ops, s))) return r;
    const bf16* nw = static_cast<const bf16*>(w.norm_k2); const float eps = e->cfg.eps; const long rows = static_cast<long>(Bp) * Nt;
    const float ks = engine_prescale() ? 0.08838834764831845f * 1.4426950408889634f : 1.0f;                         // cross-attention keys carry the softmax scale
    ops.push_back([kv, rows, Nt, D, nw, eps, ks](cudaStream_t s2) { ++g_launch_count; return launch_wan_rms_rope(kv, rows, Nt, 2 * D, D, nw, eps, nullptr, nullptr, s2, ks); });
  }
  for (auto& op : ops) FFB_CUDA(op(st));
  p->prompts_set = true;
  return 0;
}

int ffb200_wan_forward(ffb200_wan_plan* p, const void* latents_fp16, float t_model, float guidance_scale, void* noise_pred_bf16, void* stream) {
  FFB_CHECK(p && latents_fp16 && noise_pred_bf16, "null argument");
  FFB_CHECK(p->prompts_set, "ffb200_wan_set_prompts must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  int r = wensure_coefs(p, 1);
  if (r) return r;
  StepCoef c; memset(&c, 0, sizeof(c)); c.t_model = t_model; c.dynamics = DYN_ODE; c.store_slot = -1; c.logp_slot = -1;   // dt = 0: x' = x
  const size_t lat_bytes = static_cast<size_t>(p->B) * p->e->cfg.in_channels * p->F * p->H * p->W * 2;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, &c, sizeof(c), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, latents_fp16, lat_bytes, cudaMemcpyDeviceToDevice, st));
  if ((r = wrun_forward(p, st))) return r;
  SdeStepParams sp; wfill_sde(p, guidance_scale, &sp);
  sp.coef_index = 0; sp.x_next = p->x_cur; sp.v_out = static_cast<bf16*>(noise_pred_bf16);
  g_launch_count += 2;
  FFB_CUDA(launch_sde_step(sp, st));
  return 0;
}

int ffb200_wan_step(ffb200_wan_plan* p, const ffb200_step_args* a, void* stream) {
  FFB_CHECK(p && a && a->latents, "null argument");
  FFB_CHECK(p->prompts_set, "ffb200_wan_set_prompts must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  int r = wensure_coefs(p, 1);
  if (r) return r;
  const size_t lat_bytes = static_cast<size_t>(p->B) * p->e->cfg.in_channels * p->F * p->H * p->W * 2;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, &a->coef, sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, a->latents, lat_bytes, cudaMemcpyDeviceToDevice, st));
  if ((r = wrun_forward(p, st))) return r;
  SdeStepParams sp; wfill_sde(p, a->guidance_scale, &sp);
  sp.noise = a->noise; sp.seed = a->seed; sp.coef_index = 0;
  sp.next_given = static_cast<const __half*>(a->next_latents);
  sp.x_next = static_cast<__half*>(a->out_next_latents);
  sp.mean_out = a->out_mean; sp.log_prob = a->out_log_prob; sp.v_out = static_cast<bf16*>(a->out_noise_pred);
  sp.overflow_flag = a->overflow_flag;
  g_launch_count += 2;
  FFB_CUDA(launch_sde_step(sp, st));
  return 0;
}

int ffb200_wan_rollout(ffb200_wan_plan* p, const ffb200_rollout_args* a, void* stream) {
  FFB_CHECK(p && a && a->coefs && a->x0 && a->num_steps > 0, "bad rollout arguments");
  FFB_CHECK(p->prompts_set, "ffb200_wan_set_prompts must be called first");
  FFB_CHECK(!(a->use_graph && stream == nullptr), "use_graph needs a non-default stream (stream capture)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  const int T = a->num_steps;
  const int chw = p->e->cfg.in_channels * p->F * p->H * p->W;
  const size_t lat_elems = static_cast<size_t>(p->B) * chw;
  int r = wensure_coefs(p, T);
  if (r) return r;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, a->coefs, static_cast<size_t>(T) * sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, a->x0, lat_elems * 2, cudaMemcpyDeviceToDevice, st));
  if (a->all_latents && a->store_initial_slot >= 0) {
    FFB_CUDA(cudaMemcpy2DAsync(static_cast<__half*>(a->all_latents) + static_cast<size_t>(a->store_initial_slot) * chw,
                               static_cast<size_t>(a->n_latent_slots) * chw * 2, a->x0, static_cast<size_t>(chw) * 2,
                               static_cast<size_t>(chw) * 2, p->B, cudaMemcpyDeviceToDevice, st));
  }
  SdeStepParams sp; wfill_sde(p, a->guidance_scale, &sp);
  sp.noise = a->noise; sp.noise_step_stride = static_cast<long>(lat_elems); sp.seed = a->seed;
  sp.x_next = p->x_cur;
  sp.traj = static_cast<__half*>(a->all_latents); sp.traj_batch_stride = static_cast<long>(a->n_latent_slots) * chw;
  sp.logp_traj = a->log_probs; sp.logp_batch_stride = a->n_logp_slots;
  sp.overflow_flag = a->overflow_flag; sp.step_ptr = p->d_step;
  auto one_step = [&](cudaStream_t s) -> int {
    int rr = wrun_forward(p, s);
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

// ---------------------------------------------------------------- op-level entries (parity tests)
int ffb200_wan_rms_rope(void* x_bf16, long long rows, int rows_per_batch, int ld, int D, const void* weight_bf16, float eps,
                        const float* cos, const float* sin, void* stream) {
  FFB_CHECK(x_bf16 && weight_bf16 && rows > 0 && rows_per_batch > 0, "wan_rms_rope: bad argument");
  FFB_CHECK((cos == nullptr) == (sin == nullptr), "wan_rms_rope: cos and sin come together");
  g_launch_count = 1;
  FFB_CUDA(launch_wan_rms_rope(static_cast<bf16*>(x_bf16), static_cast<long>(rows), rows_per_batch, ld, D, static_cast<const bf16*>(weight_bf16),
                               eps, cos, sin, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_wan_layer_norm(const void* x_bf16, void* out_bf16, int num_batch, int rows_per_batch, int D, float eps, int mode,
                          const float* scale, const float* shift, long long mod_batch_stride, const void* weight_bf16,
                          const void* bias_bf16, void* stream) {
  FFB_CHECK(x_bf16 && out_bf16, "wan_layer_norm: null argument");
  FFB_CHECK(mode == 0 ? (scale && shift) : (mode == 1 && weight_bf16 && bias_bf16), "wan_layer_norm: mode 0 needs scale / shift, mode 1 weight / bias");
  WanLnParams lp{};
  lp.x = static_cast<const bf16*>(x_bf16); lp.out = static_cast<bf16*>(out_bf16); lp.rows_per_batch = rows_per_batch; lp.num_batch = num_batch;
  lp.D = D; lp.eps = eps; lp.mode = mode; lp.scale = scale; lp.shift = shift; lp.mod_batch_stride = static_cast<long>(mod_batch_stride);
  lp.weight = static_cast<const bf16*>(weight_bf16); lp.bias = static_cast<const bf16*>(bias_bf16);
  g_launch_count = 1;
  FFB_CUDA(launch_wan_ln(lp, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_wan_gate_residual(void* h_bf16, const void* y_bf16, const float* gate, long long gate_batch_stride, int num_batch,
                             long long rows_per_batch, int D, void* stream) {
  FFB_CHECK(h_bf16 && y_bf16 && gate, "wan_gate_residual: null argument");
  g_launch_count = 1;
  FFB_CUDA(launch_wan_gate_residual(static_cast<bf16*>(h_bf16), static_cast<const bf16*>(y_bf16), gate, static_cast<long>(gate_batch_stride),
                                    num_batch, static_cast<long>(rows_per_batch), D, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_wan_patchify(const void* x_f16, int B, int reps, int C, int F, int H, int W, int pt, int ph, int pw, void* out_bf16, void* stream) {
  FFB_CHECK(x_f16 && out_bf16 && reps >= 1, "wan_patchify: bad argument");
  g_launch_count = 1;
  FFB_CUDA(launch_wan_patchify(static_cast<const __half*>(x_f16), B, reps, C, F, H, W, pt, ph, pw, static_cast<bf16*>(out_bf16),
                               static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_attention_cross(const void* q_bf16, int q_ld, const void* kv_bf16, int kv_ld, int v_col, int B, int Sq, int Skv, int num_heads,
                           void* out_bf16, void* stream) {
  FFB_CHECK(q_bf16 && kv_bf16 && out_bf16 && B > 0 && Sq > 0 && Skv > 0 && num_heads > 0, "attention_cross: bad argument");
  FFB_CHECK(q_ld % 8 == 0 && kv_ld % 8 == 0 && v_col % 8 == 0, "attention_cross: pitches / column offsets must be multiples of 8");
  ffb200_wan_plan geo{};                       // only the geometry fields wbuild_attn reads
  geo.S = Sq; geo.Bp = B; geo.D = num_heads * 128;
  AttnParams ap;
  int r = wbuild_attn(&geo, static_cast<const bf16*>(q_bf16), q_ld, static_cast<const bf16*>(kv_bf16), kv_ld, Skv, static_cast<bf16*>(out_bf16), &ap);
  if (r) return r;
  ap.v_col = v_col;
  g_launch_count = 1;
  FFB_CUDA(launch_attention_d128_cross(ap, static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
