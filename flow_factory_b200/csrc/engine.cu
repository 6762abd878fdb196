// Host-side runtime of the rollout engine + the extern "C" boundary declared in include/ffb200.h.
//
// A `plan` pre-builds, for one geometry (batch, CFG, latent size, text length), every TMA descriptor and kernel
// parameter block of the SD3.5 MMDiT forward (DF/models/transformers/transformer_sd3.py:288-345) and the fused
// CFG + Euler/SDE + log-prob step (FF/models/stable_diffusion/sd3_5.py:431-446).  One denoise step is a fixed list
// of launches with no host synchronisation: the step index lives in device memory (bumped by the last kernel of
// the step) so the same list - or one captured CUDA graph of it - is replayed T times by ffb200_rollout.
#include "common.cuh"
#include "kernels.h"
#include "../../include/ffb200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

namespace ffb {

static thread_local std::string g_last_error;
static thread_local long long g_launch_count = 0;   // launches of the LAST entry-point call made by this thread (one driver thread per rank)

static int fail(int code, const char* what) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s (code %d%s%s)", what, code, code > 0 ? ": " : "",
           code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "");
  g_last_error = buf;
  return code;
}

#define FFB_CUDA_EARLY(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(static_cast<int>(_e), #expr);                               \
  } while (0)

// Captures `body` on `st` and makes *exec replay it.  The first time (or when the topology changed) the graph is instantiated; afterwards
// - caller-owned output / noise pointers differ from rollout to rollout, the launch list does not - the instantiated graph is updated IN
// PLACE (cudaGraphExecUpdate: new kernel arguments, no re-instantiation).
static int graph_capture_or_update(cudaGraphExec_t* exec, cudaStream_t st, const std::function<int(cudaStream_t)>& body) {
  cudaGraph_t graph = nullptr;
  FFB_CUDA_EARLY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rr = body(st);
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rr) { if (graph) cudaGraphDestroy(graph); return rr; }
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "cudaStreamEndCapture");
  if (*exec) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(*exec, graph, &info) != cudaSuccess) {
      (void)cudaGetLastError();                       // topology changed: fall back to a fresh instantiation
      cudaGraphExecDestroy(*exec); *exec = nullptr;
    }
  }
  if (!*exec) {
    const cudaError_t ie = cudaGraphInstantiate(exec, graph, 0);
    if (ie != cudaSuccess) { cudaGraphDestroy(graph); *exec = nullptr; return fail(static_cast<int>(ie), "cudaGraphInstantiate"); }
  }
  cudaGraphDestroy(graph);
  return 0;
}
#define FFB_CUDA(expr)                                             \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return fail(static_cast<int>(_e), #expr); \
  } while (0)
#define FFB_CHECK(cond, msg)            \
  do {                                  \
    if (!(cond)) return fail(-1, msg);  \
  } while (0)

// ---------------------------------------------------------------- TMA descriptor encoding (driver entry point via cudart)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// bf16 tensor, innermost dim contiguous; dims/strides innermost first; strides in ELEMENTS for dims 1..rank-1.
static int make_tmap(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_el,
                     const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(-2, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_el[i] * 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(-3, "TMA base pointer not 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16 != 0) return fail(-3, "TMA stride not a multiple of 16 bytes");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-4, "cuTensorMapEncodeTiled failed");
  return 0;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Build the parameter block of one GEMM problem.
struct GemmSpec {
  const void* A; int num_batch, rows_per_batch; long a_batch_stride; int lda, K;
  const void* W; int N; const void* bias;
  void* out; long out_batch_stride; int out_row_offset, ldo;
  int epi; const void* gate; long gate_batch_stride; const void* norm_q; const void* norm_k; int qk_dim; float eps;
  const float* row_table;
  const float* rope_cos; const float* rope_sin; int rope_row_offset;   // EPI_QKV_RMSNORM_ROPE128
  int rms_round_first;
  float k_scale;           // see GemmParams::k_scale (0 = none)
};
static int build_gemm(const GemmSpec& s, GemmParams* p) {
  memset(p, 0, sizeof(*p));
  FFB_CHECK(s.N % 64 == 0, "GEMM N must be a multiple of 64");
  FFB_CHECK(s.K % 8 == 0 && s.lda % 8 == 0 && s.ldo % 8 == 0, "GEMM K / lda / ldo must be multiples of 8");
  p->bn = gemm_pick_bn(s.N);
  if (s.epi == EPI_QKV_RMSNORM) FFB_CHECK(s.qk_dim % 64 == 0, "qk_dim must be a multiple of 64");
  if (s.epi == EPI_QKV_RMSNORM_ROPE128)
    FFB_CHECK(s.qk_dim % 128 == 0 && p->bn >= 128 && s.rope_cos && s.rope_sin && s.norm_q && s.norm_k,
              "qkv+rope epilogue: head_dim 128 blocks, N multiple of 128, rope tables and norm weights required");
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(s.K), static_cast<uint64_t>(s.rows_per_batch), static_cast<uint64_t>(s.num_batch)};
    const uint64_t str[2] = {static_cast<uint64_t>(s.lda), static_cast<uint64_t>(s.a_batch_stride > 0 ? s.a_batch_stride : static_cast<long>(s.rows_per_batch) * s.lda)};
    const uint32_t box[3] = {64, 128, 1};
    int r = make_tmap(&p->tmA, s.A, 3, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(s.K), static_cast<uint64_t>(s.N)};
    const uint64_t str[1] = {static_cast<uint64_t>(s.K)};
    const uint32_t box[2] = {64, static_cast<uint32_t>(p->bn / 2)};   // each CTA of the pair fetches half of the W tile
    int r = make_tmap(&p->tmB, s.W, 2, dims, str, box);
    if (r) return r;
  }
  p->rows_per_batch = s.rows_per_batch; p->num_batch = s.num_batch; p->N = s.N; p->K = s.K;
  p->tiles_m_per_batch = (s.rows_per_batch + 127) / 128;
  {
    const long band = (64l << 20) / (512l * s.K);   // A rows of one band <= 64 MB (half of the L2)
    const int clusters = num_sms() / 2;
    p->band = static_cast<int>(band < 1 ? 1 : (band > clusters ? clusters : band));
  }
  p->epi = s.epi;
  p->bias = static_cast<const bf16*>(s.bias);
  p->out = static_cast<bf16*>(s.out);
  p->out_batch_stride = s.out_batch_stride; p->out_row_offset = s.out_row_offset; p->ldo = s.ldo;
  p->gate = static_cast<const bf16*>(s.gate); p->gate_batch_stride = s.gate_batch_stride;
  p->norm_q = static_cast<const bf16*>(s.norm_q); p->norm_k = static_cast<const bf16*>(s.norm_k);
  p->qk_dim = s.qk_dim; p->eps = s.eps; p->row_table = s.row_table;
  p->rope_cos = s.rope_cos; p->rope_sin = s.rope_sin; p->rope_row_offset = s.rope_row_offset; p->rms_round_first = s.rms_round_first;
  p->k_scale = s.k_scale;
  return 0;
}
// The engines fold softmax_scale * log2(e) into the key heads in the QKV GEMM epilogue (GemmParams::k_scale) and tell the attention kernel
// (AttnParams::k_prescaled): the per-score FFMA of the softmax disappears (softmax.cuh).  FFB200_NO_PRESCALE=1 keeps the general path
// (A/B measurements only; read once).
static bool engine_prescale() {
  static const bool on = getenv("FFB200_NO_PRESCALE") == nullptr;
  return on;
}
static int build_attn(const void* qkv, int batch, int seq, int heads, void* out, AttnParams* p, int head_dim = 64,
                      int out_row_stride = 0) {
  memset(p, 0, sizeof(*p));
  const int D = heads * head_dim;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * D), static_cast<uint64_t>(seq), static_cast<uint64_t>(batch)};
  const uint64_t str[2] = {static_cast<uint64_t>(3 * D), static_cast<uint64_t>(seq) * 3 * D};
  const uint32_t box[3] = {64, 128, 1};      // one 64-column (= 128-byte swizzle span) panel of a 128-row Q sub-tile
  int r = make_tmap(&p->tmQKV, qkv, 3, dims, str, box);
  if (r) return r;
  const uint32_t box_kv[3] = {64, ATT_BN, 1};   // kv rows per tile (attention.cu / attention_d128.cu, same translation unit)
  r = make_tmap(&p->tmKV, qkv, 3, dims, str, box_kv);
  if (r) return r;
  p->seq_len = seq; p->num_heads = heads; p->inner_dim = D; p->batch = batch;
  p->out_row_stride = out_row_stride > 0 ? out_row_stride : D;
  p->out = static_cast<bf16*>(out); p->out_batch_stride = static_cast<long>(seq) * p->out_row_stride;
  p->scale_log2 = (1.0f / sqrtf(static_cast<float>(head_dim))) * 1.4426950408889634f;
  return 0;
}

}  // namespace ffb

using namespace ffb;

// ================================================================================================
// engine / plan objects
// ================================================================================================
struct ffb200_engine {
  ffb200_model_config cfg;
  ffb200_weights w;
  std::vector<ffb200_layer_weights> layers;
  int D;
  std::vector<int> off_n1, off_n1c;  // row offsets into the stacked adaLN matrix
  int off_out, mod_rows;
};

typedef std::function<cudaError_t(cudaStream_t)> Op;

struct ffb200_plan {
  ffb200_engine* e;
  int B, cfg, Bp, C, H, W, hp, wp, Ni, Nt, S, D;
  std::vector<void*> allocs;
  long long ws_bytes;
  // buffers
  bf16 *c0, *temb_p, *tp_a, *tproj, *ta, *temb, *mod, *peA, *h_img, *h_ctx, *a1, *a2, *ac, *qkv, *qkv2, *att, *att2, *ff, *ffc, *vout;
  float* pos_crop;
  __half* x_cur;                // current latents, `storage` dtype (allocated for 4-byte elements)
  int storage;                  // LAT_F16 (default) | LAT_BF16 | LAT_F32
  float* logp_partial;
  int* d_step;
  StepCoef* d_coefs; int coef_cap;
  std::vector<Op> fwd_ops;      // transformer forward up to norm_out (reads x_cur, leaves LN-modulated hidden states in a1)
  Op proj_op;                   // plain proj_out GEMM -> vout (inspection path: ffb200_transformer_forward)
  FinalStepParams fs;           // fused proj_out + CFG + unpatchify + scheduler.step (step / rollout path)
  unsigned int* d_done;
  bool prompts_set;
  const void* prompt_ptr;
  // per-step graph
  cudaGraphExec_t graph_exec; SdeStepParams graph_sde; bool graph_valid;
  // host staging for ffb200_rollout_host
  void *d_prompt, *d_pooled, *d_x0, *d_traj, *d_final; float* d_logp; float* d_noise; int* d_flag;
  long long staged_traj_bytes, staged_logp_bytes, staged_noise_bytes;
};

static int plan_alloc(ffb200_plan* p, void** ptr, size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaMalloc(workspace)");
  p->allocs.push_back(*ptr);
  p->ws_bytes += static_cast<long long>(bytes);
  return 0;
}

extern "C" {

const char* ffb200_last_error(void) { return g_last_error.c_str(); }
int ffb200_abi_version(void) { return FFB200_ABI_VERSION; }
long long ffb200_last_launch_count(void) { return g_launch_count; }

int ffb200_device_error(unsigned int out[4]) {
  unsigned int zero[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpyFromSymbol(out, g_dev_error, sizeof(zero));
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaMemcpyToSymbol(g_dev_error, zero, sizeof(zero));
  return 0;
}

int ffb200_debug_read_prof(unsigned long long* out, int n) {
  unsigned long long buf[256];
  cudaError_t e = cudaMemcpyFromSymbol(buf, g_prof, sizeof(buf));
  if (e != cudaSuccess) return static_cast<int>(e);
  for (int i = 0; i < n && i < 256; ++i) out[i] = buf[i];
  memset(buf, 0, sizeof(buf));
  cudaMemcpyToSymbol(g_prof, buf, sizeof(buf));
  return 0;
}

// ------------------------------------------------------------------------------------------------
int ffb200_engine_set_weights(ffb200_engine* e, const ffb200_weights* w) {
  FFB_CHECK(e && w && w->layers, "null engine/weights");
  e->w = *w;
  e->layers.assign(w->layers, w->layers + e->cfg.num_layers);
  e->w.layers = e->layers.data();
  return 0;
}

int ffb200_engine_create(const ffb200_model_config* cfg, const ffb200_weights* w, ffb200_engine** out) {
  FFB_CHECK(cfg && w && out, "null argument");
  FFB_CHECK(cfg->num_layers > 0 && cfg->num_heads > 0, "bad config");
  FFB_CHECK(cfg->patch_size == 2 && cfg->in_channels == 16, "patch_size 2 / 16 latent channels (SD3 family) required");
  FFB_CHECK(cfg->joint_attention_dim % 8 == 0 && cfg->pooled_projection_dim % 8 == 0, "joint/pooled dims must be multiples of 8");
  int dev = 0, major = 0;
  FFB_CUDA(cudaGetDevice(&dev));
  FFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FFB_CHECK(major == 10, "ffb200 kernels are sm_100a only (no fallback path)");
  ffb200_engine* e = new ffb200_engine();
  e->cfg = *cfg;
  e->D = 64 * cfg->num_heads;
  int off = 0;
  for (int i = 0; i < cfg->num_layers; ++i) {
    const bool dual = i < cfg->num_dual_layers, last = i == cfg->num_layers - 1;
    e->off_n1.push_back(off); off += (dual ? 9 : 6) * e->D;
    e->off_n1c.push_back(off); off += (last ? 2 : 6) * e->D;
  }
  e->off_out = off; off += 2 * e->D;
  e->mod_rows = off;
  int r = ffb200_engine_set_weights(e, w);
  if (r) { delete e; return r; }
  *out = e;
  return 0;
}
void ffb200_engine_destroy(ffb200_engine* e) { delete e; }
int ffb200_engine_mod_rows(const ffb200_engine* e) { return e ? e->mod_rows : -1; }

// ------------------------------------------------------------------------------------------------
void ffb200_plan_destroy(ffb200_plan* p) {
  if (!p) return;
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (void* a : p->allocs) cudaFree(a);
  delete p;
}
long long ffb200_plan_workspace_bytes(const ffb200_plan* p) { return p ? p->ws_bytes : 0; }

static int add_gemm(ffb200_plan* p, const GemmSpec& s) {
  GemmParams gp;
  int r = build_gemm(s, &gp);
  if (r) return r;
  const int sms = num_sms();
  p->fwd_ops.push_back([gp, sms](cudaStream_t st) { ++g_launch_count; return launch_gemm(gp, sms, st); });
  return 0;
}
// wq / wk: the per-head RMSNorm weights of the QKV epilogues that write this attention's q and k (image stream, text stream or null):
// the kernel derives a bound on every exponent from them and drops the range guard of its polynomial slots when the bound allows
// (AttnParams::bound_wq; FFB200_NO_SCORE_BOUND=1 withholds them for A/B runs).
static int add_attn(ffb200_plan* p, const void* qkv, int seq, void* out, const void* wq0, const void* wk0, const void* wq1 = nullptr,
                    const void* wk1 = nullptr) {
  AttnParams ap;
  int r = build_attn(qkv, p->Bp, seq, p->e->cfg.num_heads, out, &ap);
  if (r) return r;
  ap.k_prescaled = engine_prescale() ? 1 : 0;       // every QKV GEMM of this engine folds the scale into k (below)
  static const bool with_bound = getenv("FFB200_NO_SCORE_BOUND") == nullptr;
  if (with_bound && ap.k_prescaled) {
    ap.bound_wq[0] = static_cast<const bf16*>(wq0); ap.bound_wk[0] = static_cast<const bf16*>(wk0);
    ap.bound_wq[1] = static_cast<const bf16*>(wq1); ap.bound_wk[1] = static_cast<const bf16*>(wk1);
  }
  p->fwd_ops.push_back([ap](cudaStream_t st) { ++g_launch_count; return launch_attention(ap, st); });
  return 0;
}
static void add_lnmod(ffb200_plan* p, const bf16* x, int rows_per_batch, const bf16* shift1, const bf16* scale1, bf16* out1,
                      const bf16* shift2, const bf16* scale2, bf16* out2) {
  LnModParams lp{};
  lp.x = x; lp.rows_per_batch = rows_per_batch; lp.num_batch = p->Bp; lp.D = p->D; lp.eps = 1e-6f;
  lp.shift1 = shift1; lp.scale1 = scale1; lp.out1 = out1; lp.shift2 = shift2; lp.scale2 = scale2; lp.out2 = out2;
  lp.mod_batch_stride = p->e->mod_rows;
  p->fwd_ops.push_back([lp](cudaStream_t st) { ++g_launch_count; return launch_ln_modulate(lp, st); });
}
static void add_small(ffb200_plan* p, std::vector<Op>& ops, const bf16* in, int K, const void* W, const void* bias, int N,
                      bf16* out, const bf16* addend, int silu) {
  SmallLinearParams sp{};
  sp.in = in; sp.batch = p->Bp; sp.K = K; sp.in_stride = K; sp.W = static_cast<const bf16*>(W);
  sp.bias = static_cast<const bf16*>(bias); sp.N = N; sp.out = out; sp.out_stride = N;
  sp.addend = addend; sp.addend_stride = N; sp.silu_input = silu;
  ops.push_back([sp](cudaStream_t st) { ++g_launch_count; return launch_small_linear(sp, st); });
}

int ffb200_plan_create(ffb200_engine* e, int batch, int cfg, int lat_h, int lat_w, int n_text, ffb200_plan** out) {
  FFB_CHECK(e && out, "null argument");
  const ffb200_model_config& mc = e->cfg;
  FFB_CHECK(batch > 0 && batch * (cfg ? 2 : 1) <= 64, "batch out of range");
  FFB_CHECK(lat_h % mc.patch_size == 0 && lat_w % mc.patch_size == 0 && lat_w % 4 == 0, "latent size must be patch aligned, W % 4 == 0");
  FFB_CHECK(lat_h / mc.patch_size <= mc.pos_embed_max_size && lat_w / mc.patch_size <= mc.pos_embed_max_size, "latent larger than pos_embed_max_size");
  FFB_CHECK(n_text > 0, "n_text must be positive");
  ffb200_plan* p = new ffb200_plan();
  p->e = e; p->B = batch; p->cfg = cfg ? 1 : 0; p->Bp = batch * (cfg ? 2 : 1);
  p->C = mc.in_channels; p->H = lat_h; p->W = lat_w; p->hp = lat_h / mc.patch_size; p->wp = lat_w / mc.patch_size;
  p->Ni = p->hp * p->wp; p->Nt = n_text; p->S = p->Ni + p->Nt; p->D = e->D;
  p->storage = LAT_F16;
  p->ws_bytes = 0; p->graph_exec = nullptr; p->graph_valid = false; p->prompts_set = false; p->coef_cap = 0; p->d_coefs = nullptr;
  p->d_prompt = p->d_pooled = p->d_x0 = p->d_traj = p->d_final = nullptr; p->d_logp = nullptr; p->d_noise = nullptr; p->d_flag = nullptr;
  p->staged_traj_bytes = p->staged_logp_bytes = p->staged_noise_bytes = 0;
  const int D = p->D, Bp = p->Bp, Ni = p->Ni, Nt = p->Nt, S = p->S, R = e->mod_rows;
  const bool any_dual = mc.num_dual_layers > 0;
  int r = 0;
#define ALLOC(field, count, type) if (!r) r = plan_alloc(p, reinterpret_cast<void**>(&p->field), static_cast<size_t>(count) * sizeof(type))
  ALLOC(c0, static_cast<size_t>(Bp) * Nt * D, bf16);
  ALLOC(temb_p, static_cast<size_t>(Bp) * D, bf16);
  ALLOC(tp_a, static_cast<size_t>(Bp) * D, bf16);
  ALLOC(tproj, static_cast<size_t>(Bp) * 256, bf16);
  ALLOC(ta, static_cast<size_t>(Bp) * D, bf16);
  ALLOC(temb, static_cast<size_t>(Bp) * D, bf16);
  ALLOC(mod, static_cast<size_t>(Bp) * R, bf16);
  ALLOC(peA, static_cast<size_t>(Bp) * Ni * 64, bf16);
  ALLOC(pos_crop, static_cast<size_t>(Ni) * D, float);
  ALLOC(h_img, static_cast<size_t>(Bp) * Ni * D, bf16);
  ALLOC(h_ctx, static_cast<size_t>(Bp) * Nt * D, bf16);
  ALLOC(a1, static_cast<size_t>(Bp) * Ni * D, bf16);
  if (any_dual) ALLOC(a2, static_cast<size_t>(Bp) * Ni * D, bf16);
  ALLOC(ac, static_cast<size_t>(Bp) * Nt * D, bf16);
  ALLOC(qkv, static_cast<size_t>(Bp) * S * 3 * D, bf16);
  if (any_dual) ALLOC(qkv2, static_cast<size_t>(Bp) * Ni * 3 * D, bf16);
  ALLOC(att, static_cast<size_t>(Bp) * S * D, bf16);
  if (any_dual) ALLOC(att2, static_cast<size_t>(Bp) * Ni * D, bf16);
  ALLOC(ff, static_cast<size_t>(Bp) * Ni * 4 * D, bf16);
  ALLOC(ffc, static_cast<size_t>(Bp) * Nt * 4 * D, bf16);
  ALLOC(vout, static_cast<size_t>(Bp) * Ni * 64, bf16);
  ALLOC(x_cur, static_cast<size_t>(p->B) * p->C * p->H * p->W * 2, __half);   // room for fp32 storage
  ALLOC(logp_partial, static_cast<size_t>(p->B) * std::max(64, (Ni + 127) / 128), float);
  ALLOC(d_step, 1, int);
  ALLOC(d_done, 1, unsigned int);
  if (!r) cudaMemset(p->d_done, 0, sizeof(unsigned int));
#undef ALLOC
  if (r) { ffb200_plan_destroy(p); return r; }
  // cropped positional table (DF/models/embeddings.py:531-552)
  {
    const int top = (mc.pos_embed_max_size - p->hp) / 2, left = (mc.pos_embed_max_size - p->wp) / 2;
    const float* src = e->w.pos_embed + (static_cast<size_t>(top) * mc.pos_embed_max_size + left) * D;
    cudaError_t ce = cudaMemcpy2D(p->pos_crop, static_cast<size_t>(p->wp) * D * 4, src, static_cast<size_t>(mc.pos_embed_max_size) * D * 4,
                                  static_cast<size_t>(p->wp) * D * 4, p->hp, cudaMemcpyDeviceToDevice);
    if (ce != cudaSuccess) { ffb200_plan_destroy(p); return fail(static_cast<int>(ce), "cudaMemcpy2D(pos_embed crop)"); }
  }

  // ---------------- forward op list ----------------
  std::vector<Op>& ops = p->fwd_ops;
  const ffb200_weights& w = e->w;
  {  // timestep embedding + all-layer adaLN GEMV (embeddings.py:1592-1600; normalization.py:120,167,348)
    ffb200_plan* pp = p;
    ops.push_back([pp](cudaStream_t st) { ++g_launch_count; return launch_timestep_proj(pp->d_coefs, pp->d_step, 0, pp->Bp, pp->tproj, st); });
    add_small(p, ops, p->tproj, 256, w.t1_w, w.t1_b, D, p->ta, nullptr, 0);
    add_small(p, ops, p->ta, D, w.t2_w, w.t2_b, D, p->temb, p->temb_p, 1);
    add_small(p, ops, p->temb, D, w.mod_w, w.mod_b, R, p->mod, nullptr, 1);
  }
  {  // patch embed (embeddings.py:554-583)
    ffb200_plan* pp = p;
    const int reps = p->cfg ? 2 : 1, patch = mc.patch_size;
    ops.push_back([pp, reps, patch](cudaStream_t st) { ++g_launch_count; return launch_patchify(pp->x_cur, pp->storage, pp->B, reps, pp->C, pp->H, pp->W, patch, pp->peA, st); });
    GemmSpec s = {p->peA, Bp, Ni, 0, 64, 64, w.pe_w, D, w.pe_b, p->h_img, static_cast<long>(Ni) * D, 0, D,
                  EPI_BIAS_ADD_ROWTABLE, nullptr, 0, nullptr, nullptr, 0, 0.f, p->pos_crop};
    if ((r = add_gemm(p, s))) { ffb200_plan_destroy(p); return r; }
    const size_t ctx_bytes = static_cast<size_t>(Bp) * Nt * D * 2;
    ops.push_back([pp, ctx_bytes](cudaStream_t st) { return cudaMemcpyAsync(pp->h_ctx, pp->c0, ctx_bytes, cudaMemcpyDeviceToDevice, st); });
  }
  for (int i = 0; i < mc.num_layers && !r; ++i) {
    const ffb200_layer_weights& L = e->layers[i];
    const bool dual = i < mc.num_dual_layers, last = i == mc.num_layers - 1;
    const bf16* m1 = p->mod + e->off_n1[i];    // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp, (shift2, scale2, gate2)
    const bf16* mc1 = p->mod + e->off_n1c[i];  // same 6 for the context stream, or (scale, shift) on the last layer
    // adaLN-Zero norm1 (+ second modulation for attn2)
    add_lnmod(p, p->h_img, Ni, m1 + 0 * D, m1 + 1 * D, p->a1, dual ? m1 + 6 * D : nullptr, dual ? m1 + 7 * D : nullptr, dual ? p->a2 : nullptr);
    if (last) add_lnmod(p, p->h_ctx, Nt, mc1 + 1 * D, mc1 + 0 * D, p->ac, nullptr, nullptr, nullptr);
    else add_lnmod(p, p->h_ctx, Nt, mc1 + 0 * D, mc1 + 1 * D, p->ac, nullptr, nullptr, nullptr);
    // fused q|k|v projections + per-head RMSNorm, written token-major into the joint [Bp, S, 3D] buffer (image rows first)
    GemmSpec sq = {p->a1, Bp, Ni, 0, D, D, L.qkv_w, 3 * D, L.qkv_b, p->qkv, static_cast<long>(S) * 3 * D, 0, 3 * D,
                   EPI_QKV_RMSNORM, nullptr, 0, L.norm_q, L.norm_k, D, 1e-6f, nullptr};
    sq.k_scale = engine_prescale() ? 0.125f * 1.4426950408889634f : 0.f;
    if ((r = add_gemm(p, sq))) break;
    GemmSpec sc = {p->ac, Bp, Nt, 0, D, D, L.add_qkv_w, 3 * D, L.add_qkv_b, p->qkv, static_cast<long>(S) * 3 * D, Ni, 3 * D,
                   EPI_QKV_RMSNORM, nullptr, 0, L.norm_added_q, L.norm_added_k, D, 1e-6f, nullptr};
    sc.k_scale = engine_prescale() ? 0.125f * 1.4426950408889634f : 0.f;
    if ((r = add_gemm(p, sc))) break;
    if ((r = add_attn(p, p->qkv, S, p->att, L.norm_q, L.norm_k, L.norm_added_q, L.norm_added_k))) break;
    // to_out + gate_msa residual (image rows of the joint attention output)
    GemmSpec so = {p->att, Bp, Ni, static_cast<long>(S) * D, D, D, L.out_w, D, L.out_b, p->h_img, static_cast<long>(Ni) * D, 0, D,
                   EPI_GATE_RESIDUAL, m1 + 2 * D, R, nullptr, nullptr, 0, 0.f, nullptr};
    if ((r = add_gemm(p, so))) break;
    if (!last) {
      GemmSpec sa = {p->att + static_cast<size_t>(Ni) * D, Bp, Nt, static_cast<long>(S) * D, D, D, L.add_out_w, D, L.add_out_b, p->h_ctx,
                     static_cast<long>(Nt) * D, 0, D, EPI_GATE_RESIDUAL, mc1 + 2 * D, R, nullptr, nullptr, 0, 0.f, nullptr};
      if ((r = add_gemm(p, sa))) break;
    }
    if (dual) {  // attn2: image-only self attention (attention.py:714-717)
      GemmSpec s2 = {p->a2, Bp, Ni, 0, D, D, L.qkv2_w, 3 * D, L.qkv2_b, p->qkv2, static_cast<long>(Ni) * 3 * D, 0, 3 * D,
                     EPI_QKV_RMSNORM, nullptr, 0, L.norm_q2, L.norm_k2, D, 1e-6f, nullptr};
      s2.k_scale = engine_prescale() ? 0.125f * 1.4426950408889634f : 0.f;
      if ((r = add_gemm(p, s2))) break;
      if ((r = add_attn(p, p->qkv2, Ni, p->att2, L.norm_q2, L.norm_k2))) break;
      GemmSpec so2 = {p->att2, Bp, Ni, 0, D, D, L.out2_w, D, L.out2_b, p->h_img, static_cast<long>(Ni) * D, 0, D,
                      EPI_GATE_RESIDUAL, m1 + 8 * D, R, nullptr, nullptr, 0, 0.f, nullptr};
      if ((r = add_gemm(p, so2))) break;
    }
    // image MLP
    add_lnmod(p, p->h_img, Ni, m1 + 3 * D, m1 + 4 * D, p->a1, nullptr, nullptr, nullptr);
    GemmSpec f1 = {p->a1, Bp, Ni, 0, D, D, L.ff1_w, 4 * D, L.ff1_b, p->ff, static_cast<long>(Ni) * 4 * D, 0, 4 * D,
                   EPI_BIAS_GELU, nullptr, 0, nullptr, nullptr, 0, 0.f, nullptr};
    if ((r = add_gemm(p, f1))) break;
    GemmSpec f2 = {p->ff, Bp, Ni, 0, 4 * D, 4 * D, L.ff2_w, D, L.ff2_b, p->h_img, static_cast<long>(Ni) * D, 0, D,
                   EPI_GATE_RESIDUAL, m1 + 5 * D, R, nullptr, nullptr, 0, 0.f, nullptr};
    if ((r = add_gemm(p, f2))) break;
    if (!last) {  // context MLP
      add_lnmod(p, p->h_ctx, Nt, mc1 + 3 * D, mc1 + 4 * D, p->ac, nullptr, nullptr, nullptr);
      GemmSpec c1 = {p->ac, Bp, Nt, 0, D, D, L.cff1_w, 4 * D, L.cff1_b, p->ffc, static_cast<long>(Nt) * 4 * D, 0, 4 * D,
                     EPI_BIAS_GELU, nullptr, 0, nullptr, nullptr, 0, 0.f, nullptr};
      if ((r = add_gemm(p, c1))) break;
      GemmSpec c2 = {p->ffc, Bp, Nt, 0, 4 * D, 4 * D, L.cff2_w, D, L.cff2_b, p->h_ctx, static_cast<long>(Nt) * D, 0, D,
                     EPI_GATE_RESIDUAL, mc1 + 5 * D, R, nullptr, nullptr, 0, 0.f, nullptr};
      if ((r = add_gemm(p, c2))) break;
    }
  }
  if (!r) {  // norm_out (AdaLayerNormContinuous: scale, shift) + proj_out  (transformer_sd3.py:326-327)
    const bf16* mo = p->mod + e->off_out;
    add_lnmod(p, p->h_img, Ni, mo + 1 * D, mo + 0 * D, p->a1, nullptr, nullptr, nullptr);
    GemmSpec po = {p->a1, Bp, Ni, 0, D, D, w.proj_w, 64, w.proj_b, p->vout, static_cast<long>(Ni) * 64, 0, 64,
                   EPI_BIAS, nullptr, 0, nullptr, nullptr, 0, 0.f, nullptr};
    GemmParams gp;
    r = build_gemm(po, &gp);
    if (!r) {
      const int sms = num_sms();
      p->proj_op = [gp, sms](cudaStream_t st) { ++g_launch_count; return launch_gemm(gp, sms, st); };
      memset(&p->fs, 0, sizeof(p->fs));
      p->fs.tmA = gp.tmA;
      const uint64_t dims[2] = {static_cast<uint64_t>(D), 64};
      const uint64_t str[1] = {static_cast<uint64_t>(D)};
      const uint32_t box[2] = {64, 64};
      r = make_tmap(&p->fs.tmW, w.proj_w, 2, dims, str, box);
      p->fs.K = D; p->fs.bias = static_cast<const bf16*>(w.proj_b); p->fs.done_counter = p->d_done;
    }
  }
  if (r) { ffb200_plan_destroy(p); return r; }
  *out = p;
  return 0;
}

static int ensure_coefs(ffb200_plan* p, int n) {
  if (n <= p->coef_cap) return 0;
  StepCoef* d = nullptr;
  int r = plan_alloc(p, reinterpret_cast<void**>(&d), static_cast<size_t>(n) * sizeof(StepCoef));
  if (r) return r;
  p->d_coefs = d; p->coef_cap = n;
  p->graph_valid = false;
  return 0;
}

int ffb200_plan_set_prompts(ffb200_plan* p, const void* prompt_embeds_bf16, const void* pooled_bf16, void* stream) {
  FFB_CHECK(p && prompt_embeds_bf16 && pooled_bf16, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ffb200_engine* e = p->e;
  const int J = e->cfg.joint_attention_dim, P = e->cfg.pooled_projection_dim, D = p->D;
  // context_embedder (transformer_sd3.py:293) - timestep independent, cached for the whole rollout
  GemmSpec s = {prompt_embeds_bf16, p->Bp, p->Nt, 0, J, J, e->w.ctx_w, D, e->w.ctx_b, p->c0, static_cast<long>(p->Nt) * D, 0, D,
                EPI_BIAS, nullptr, 0, nullptr, nullptr, 0, 0.f, nullptr};
  GemmParams gp;
  int r = build_gemm(s, &gp);
  if (r) return r;
  ++g_launch_count;
  FFB_CUDA(launch_gemm(gp, num_sms(), st));
  // pooled-text MLP (embeddings.py:2213-2217)
  std::vector<Op> ops;
  add_small(p, ops, static_cast<const bf16*>(pooled_bf16), P, e->w.p1_w, e->w.p1_b, D, p->tp_a, nullptr, 0);
  add_small(p, ops, p->tp_a, D, e->w.p2_w, e->w.p2_b, D, p->temb_p, nullptr, 1);
  for (auto& op : ops) FFB_CUDA(op(st));
  p->prompts_set = true;
  return 0;
}

static int run_forward(ffb200_plan* p, cudaStream_t st) {
  for (auto& op : p->fwd_ops) {
    cudaError_t e = op(st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "forward launch");
  }
  return 0;
}

// unpatchify kernel for inspection output
__global__ void unpatchify_kernel(const bf16* v_tokens, int Bp, int C, int H, int W, int patch, bf16* out) {
  const int wp = W / patch, hp = H / patch;
  const long total = static_cast<long>(Bp) * C * H * W;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W), y = static_cast<int>((idx / W) % H), c = static_cast<int>((idx / (static_cast<long>(W) * H)) % C);
    const int b = static_cast<int>(idx / (static_cast<long>(W) * H * C));
    const int tok = (y / patch) * wp + x / patch;
    const int n = ((y % patch) * patch + (x % patch)) * C + c;
    out[idx] = v_tokens[(static_cast<long>(b) * hp * wp + tok) * (patch * patch * C) + n];
  }
}

static void fill_sde(const ffb200_plan* p, SdeStepParams* sp) {
  memset(sp, 0, sizeof(*sp));
  sp->B = p->B; sp->C = p->C; sp->H = p->H; sp->W = p->W; sp->patch = p->e->cfg.patch_size;
  sp->cfg = p->cfg; sp->x = p->x_cur; sp->logp_partial = p->logp_partial; sp->coef_table = p->d_coefs;
  sp->storage = p->storage;
}

int ffb200_plan_set_latent_dtype(ffb200_plan* p, int dtype) {
  FFB_CHECK(p, "null plan");
  FFB_CHECK(dtype == LAT_F16 || dtype == LAT_BF16 || dtype == LAT_F32, "latent dtype: 0 = fp16, 1 = bf16, 2 = fp32");
  if (p->storage != dtype) { p->storage = dtype; p->graph_valid = false; }
  return 0;
}

int ffb200_transformer_forward(ffb200_plan* p, const void* latents_fp16, float t_model, void* noise_pred_nchw, void* stream) {
  FFB_CHECK(p && latents_fp16, "null argument");
  FFB_CHECK(p->prompts_set, "ffb200_plan_set_prompts must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  int r = ensure_coefs(p, 1);
  if (r) return r;
  StepCoef c; memset(&c, 0, sizeof(c)); c.t_model = t_model;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, &c, sizeof(c), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, latents_fp16, static_cast<size_t>(p->B) * p->C * p->H * p->W * lat_elem_bytes(p->storage), cudaMemcpyDeviceToDevice, st));
  if ((r = run_forward(p, st))) return r;
  FFB_CUDA(p->proj_op(st));
  if (noise_pred_nchw) {
    ++g_launch_count;
    unpatchify_kernel<<<148 * 4, 256, 0, st>>>(p->vout, p->Bp, p->C, p->H, p->W, p->e->cfg.patch_size, static_cast<bf16*>(noise_pred_nchw));
    FFB_CUDA(cudaGetLastError());
  }
  return 0;
}

int ffb200_step(ffb200_plan* p, const ffb200_step_args* a, void* stream) {
  FFB_CHECK(p && a && a->latents, "null argument");
  FFB_CHECK(p->prompts_set, "ffb200_plan_set_prompts must be called first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 0;
  int r = ensure_coefs(p, 1);
  if (r) return r;
  static_assert(sizeof(StepCoef) == sizeof(ffb200_step_coef), "StepCoef layout");
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, &a->coef, sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, a->latents, static_cast<size_t>(p->B) * p->C * p->H * p->W * lat_elem_bytes(p->storage), cudaMemcpyDeviceToDevice, st));
  if ((r = run_forward(p, st))) return r;
  SdeStepParams sp; fill_sde(p, &sp);
  sp.guidance = a->guidance_scale; sp.noise = a->noise; sp.seed = a->seed; sp.coef_index = 0;
  sp.next_given = a->next_latents;
  sp.x_next = a->out_next_latents;
  sp.mean_out = a->out_mean; sp.log_prob = a->out_log_prob; sp.v_out = static_cast<bf16*>(a->out_noise_pred);
  sp.overflow_flag = a->overflow_flag;
  FinalStepParams fs = p->fs;
  fs.sde = sp;
  ++g_launch_count;
  FFB_CUDA(launch_final_step(fs, st));
  return 0;
}

static int rollout_impl(ffb200_plan* p, const ffb200_rollout_args* a, cudaStream_t st) {
  const int T = a->num_steps;
  const size_t lat_elems = static_cast<size_t>(p->B) * p->C * p->H * p->W;
  const int chw = p->C * p->H * p->W;
  int r = ensure_coefs(p, T);
  if (r) return r;
  FFB_CUDA(cudaMemcpyAsync(p->d_coefs, a->coefs, static_cast<size_t>(T) * sizeof(StepCoef), cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  const size_t eb = static_cast<size_t>(lat_elem_bytes(p->storage));
  FFB_CUDA(cudaMemcpyAsync(p->x_cur, a->x0, lat_elems * eb, cudaMemcpyDeviceToDevice, st));
  if (a->all_latents && a->store_initial_slot >= 0) {
    FFB_CUDA(cudaMemcpy2DAsync(static_cast<char*>(a->all_latents) + static_cast<size_t>(a->store_initial_slot) * chw * eb,
                               static_cast<size_t>(a->n_latent_slots) * chw * eb, a->x0, static_cast<size_t>(chw) * eb,
                               static_cast<size_t>(chw) * eb, p->B, cudaMemcpyDeviceToDevice, st));
  }
  SdeStepParams sp; fill_sde(p, &sp);
  sp.guidance = a->guidance_scale; sp.noise = a->noise; sp.noise_step_stride = static_cast<long>(lat_elems); sp.seed = a->seed;
  sp.x_next = p->x_cur;  // in place: every thread reads its own 4 pixels before writing them
  sp.traj = a->all_latents; sp.traj_batch_stride = static_cast<long>(a->n_latent_slots) * chw;
  sp.logp_traj = a->log_probs; sp.logp_batch_stride = a->n_logp_slots;
  sp.overflow_flag = a->overflow_flag; sp.step_ptr = p->d_step;

  FinalStepParams fs = p->fs;
  fs.sde = sp;
  auto one_step = [&](cudaStream_t s) -> int {
    int rr = run_forward(p, s);
    if (rr) return rr;
    ++g_launch_count;
    cudaError_t e = launch_final_step(fs, s);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "final_step launch");
    return 0;
  };

  if (a->use_graph) {
    const bool same = p->graph_valid && memcmp(&p->graph_sde, &sp, sizeof(sp)) == 0;
    if (!same) {
      const long long before = g_launch_count;
      const int rr = graph_capture_or_update(&p->graph_exec, st, one_step);
      g_launch_count = before;
      if (rr) { p->graph_valid = false; return rr; }
      p->graph_sde = sp; p->graph_valid = true;
    }
    const long long per_step = static_cast<long long>(p->fwd_ops.size()) - 1 /*memcpy*/ + 1;
    for (int i = 0; i < T; ++i) FFB_CUDA(cudaGraphLaunch(p->graph_exec, st));
    g_launch_count += per_step * T;
  } else {
    for (int i = 0; i < T; ++i) {
      int rr = one_step(st);
      if (rr) return rr;
    }
  }
  if (a->final_latents) FFB_CUDA(cudaMemcpyAsync(a->final_latents, p->x_cur, lat_elems * eb, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int ffb200_rollout(ffb200_plan* p, const ffb200_rollout_args* a, void* stream) {
  FFB_CHECK(p && a && a->coefs && a->x0 && a->num_steps > 0, "bad rollout arguments");
  FFB_CHECK(p->prompts_set, "ffb200_plan_set_prompts must be called first");
  FFB_CHECK(!(a->use_graph && stream == nullptr), "use_graph needs a non-default stream (stream capture)");
  g_launch_count = 0;
  return rollout_impl(p, a, static_cast<cudaStream_t>(stream));
}

int ffb200_rollout_host(ffb200_plan* p, const ffb200_rollout_args* a, const void* prompt_embeds_bf16, const void* pooled_bf16, void* stream) {
  FFB_CHECK(p && a && a->coefs && a->x0 && a->num_steps > 0 && prompt_embeds_bf16 && pooled_bf16, "bad rollout arguments");
  FFB_CHECK(!(a->use_graph && stream == nullptr), "use_graph needs a non-default stream (stream capture)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ffb200_engine* e = p->e;
  const size_t lat_elems = static_cast<size_t>(p->B) * p->C * p->H * p->W;
  const size_t prompt_bytes = static_cast<size_t>(p->Bp) * p->Nt * e->cfg.joint_attention_dim * 2;
  const size_t pooled_bytes = static_cast<size_t>(p->Bp) * e->cfg.pooled_projection_dim * 2;
  int r = 0;
  if (!p->d_prompt) {
    if ((r = plan_alloc(p, &p->d_prompt, prompt_bytes))) return r;
    if ((r = plan_alloc(p, &p->d_pooled, pooled_bytes))) return r;
    if ((r = plan_alloc(p, &p->d_x0, lat_elems * 4))) return r;       // sized for fp32 storage
    if ((r = plan_alloc(p, &p->d_final, lat_elems * 4))) return r;
    if ((r = plan_alloc(p, reinterpret_cast<void**>(&p->d_flag), sizeof(int)))) return r;
  }
  const long long eb = lat_elem_bytes(p->storage);
  const long long traj_bytes = a->all_latents ? static_cast<long long>(lat_elems) * a->n_latent_slots * eb : 0;
  const long long logp_bytes = a->log_probs ? static_cast<long long>(p->B) * a->n_logp_slots * 4 : 0;
  const long long noise_bytes = a->noise ? static_cast<long long>(lat_elems) * a->num_steps * 4 : 0;
  if (traj_bytes > p->staged_traj_bytes) { if ((r = plan_alloc(p, &p->d_traj, traj_bytes))) return r; p->staged_traj_bytes = traj_bytes; }
  if (logp_bytes > p->staged_logp_bytes) { if ((r = plan_alloc(p, reinterpret_cast<void**>(&p->d_logp), logp_bytes))) return r; p->staged_logp_bytes = logp_bytes; }
  if (noise_bytes > p->staged_noise_bytes) { if ((r = plan_alloc(p, reinterpret_cast<void**>(&p->d_noise), noise_bytes))) return r; p->staged_noise_bytes = noise_bytes; }
  FFB_CUDA(cudaMemcpyAsync(p->d_prompt, prompt_embeds_bf16, prompt_bytes, cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemcpyAsync(p->d_pooled, pooled_bf16, pooled_bytes, cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemcpyAsync(p->d_x0, a->x0, lat_elems * eb, cudaMemcpyHostToDevice, st));
  if (a->noise) FFB_CUDA(cudaMemcpyAsync(p->d_noise, a->noise, noise_bytes, cudaMemcpyHostToDevice, st));
  FFB_CUDA(cudaMemsetAsync(p->d_flag, 0, sizeof(int), st));
  g_launch_count = 0;
  if ((r = ffb200_plan_set_prompts(p, p->d_prompt, p->d_pooled, st))) return r;
  ffb200_rollout_args d = *a;
  d.x0 = p->d_x0; d.noise = a->noise ? p->d_noise : nullptr;
  d.all_latents = a->all_latents ? p->d_traj : nullptr;
  d.log_probs = a->log_probs ? p->d_logp : nullptr;
  d.final_latents = p->d_final; d.overflow_flag = p->d_flag;
  if ((r = rollout_impl(p, &d, st))) return r;
  if (a->all_latents) FFB_CUDA(cudaMemcpyAsync(a->all_latents, p->d_traj, traj_bytes, cudaMemcpyDeviceToHost, st));
  if (a->log_probs) FFB_CUDA(cudaMemcpyAsync(a->log_probs, p->d_logp, logp_bytes, cudaMemcpyDeviceToHost, st));
  if (a->final_latents) FFB_CUDA(cudaMemcpyAsync(a->final_latents, p->d_final, lat_elems * eb, cudaMemcpyDeviceToHost, st));
  if (a->overflow_flag) FFB_CUDA(cudaMemcpyAsync(a->overflow_flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  FFB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ================================================================================================
// op-level entries
// ================================================================================================
int ffb200_linear(const void* A, int num_batch, int rows_per_batch, long long a_batch_stride, int lda, int K, const void* W, int N,
                  const void* bias, void* out, long long out_batch_stride, int out_row_offset, int ldo, int epilogue,
                  const void* gate, long long gate_batch_stride, const void* norm_q, const void* norm_k, int qk_dim, float eps,
                  const float* row_table, void* stream) {
  FFB_CHECK(A && W && out, "null argument");
  GemmSpec s = {A, num_batch, rows_per_batch, static_cast<long>(a_batch_stride), lda, K, W, N, bias, out,
                static_cast<long>(out_batch_stride), out_row_offset, ldo, epilogue, gate, static_cast<long>(gate_batch_stride),
                norm_q, norm_k, qk_dim, eps, row_table};
  GemmParams gp;
  int r = build_gemm(s, &gp);
  if (r) return r;
  g_launch_count = 1;
  FFB_CUDA(launch_gemm(gp, num_sms(), static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_linear_qkv_rope(const void* A, int num_batch, int rows_per_batch, long long a_batch_stride, int lda, int K, const void* W,
                           int N, const void* bias, void* out, long long out_batch_stride, int out_row_offset, int ldo,
                           const void* norm_q, const void* norm_k, int qk_dim, float eps, const float* rope_cos,
                           const float* rope_sin, int rope_row_offset, void* stream) {
  FFB_CHECK(A && W && out, "null argument");
  GemmSpec s = {A, num_batch, rows_per_batch, static_cast<long>(a_batch_stride), lda, K, W, N, bias, out,
                static_cast<long>(out_batch_stride), out_row_offset, ldo, EPI_QKV_RMSNORM_ROPE128, nullptr, 0, norm_q, norm_k, qk_dim, eps,
                nullptr, rope_cos, rope_sin, rope_row_offset};
  GemmParams gp;
  int r = build_gemm(s, &gp);
  if (r) return r;
  g_launch_count = 1;
  FFB_CUDA(launch_gemm(gp, num_sms(), static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_attention(const void* qkv, int batch, int seq_len, int num_heads, void* out, void* stream) {
  FFB_CHECK(qkv && out && batch > 0 && seq_len > 0 && num_heads > 0, "bad argument");
  AttnParams ap;
  int r = build_attn(qkv, batch, seq_len, num_heads, out, &ap);
  if (r) return r;
  g_launch_count = 1;
  FFB_CUDA(launch_attention(ap, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_attention_ex(const void* qkv, int batch, int seq_len, int num_heads, int head_dim, void* out, int out_row_stride,
                        void* stream) {
  FFB_CHECK(qkv && out && batch > 0 && seq_len > 0 && num_heads > 0, "bad argument");
  if (head_dim != 64 && head_dim != 128) return fail(-2, "ffb200_attention_ex: head_dim must be 64 or 128");
  if (out_row_stride != 0 && out_row_stride < num_heads * head_dim) return fail(-2, "ffb200_attention_ex: out_row_stride < inner dim");
  AttnParams ap;
  int r = build_attn(qkv, batch, seq_len, num_heads, out, &ap, head_dim, out_row_stride);
  if (r) return r;
  if (head_dim == 64 && ap.out_row_stride != ap.inner_dim) return fail(-2, "ffb200_attention_ex: head_dim 64 writes dense rows only");
  g_launch_count = 1;
  FFB_CUDA(head_dim == 64 ? launch_attention(ap, static_cast<cudaStream_t>(stream)) : launch_attention_d128(ap, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_attention_scaled(const void* qkv, int batch, int seq_len, int num_heads, int head_dim, void* out, int out_row_stride,
                            float softmax_scale, int k_prescaled, void* stream) {
  FFB_CHECK(qkv && out && batch > 0 && seq_len > 0 && num_heads > 0, "bad argument");
  if (head_dim != 64 && head_dim != 128) return fail(-2, "ffb200_attention_scaled: head_dim must be 64 or 128");
  if (out_row_stride != 0 && out_row_stride < num_heads * head_dim) return fail(-2, "ffb200_attention_scaled: out_row_stride < inner dim");
  AttnParams ap;
  int r = build_attn(qkv, batch, seq_len, num_heads, out, &ap, head_dim, out_row_stride);
  if (r) return r;
  if (head_dim == 64 && ap.out_row_stride != ap.inner_dim) return fail(-2, "ffb200_attention_scaled: head_dim 64 writes dense rows only");
  if (softmax_scale > 0.f) ap.scale_log2 = softmax_scale * 1.4426950408889634f;
  ap.k_prescaled = k_prescaled ? 1 : 0;
  g_launch_count = 1;
  FFB_CUDA(head_dim == 64 ? launch_attention(ap, static_cast<cudaStream_t>(stream)) : launch_attention_d128(ap, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_attention_normed(const void* qkv, int batch, int seq_len, int num_heads, void* out, const void* wq0, const void* wk0,
                            const void* wq1, const void* wk1, void* stream) {
  FFB_CHECK(qkv && out && wq0 && wk0 && batch > 0 && seq_len > 0 && num_heads > 0, "bad argument");
  AttnParams ap;
  int r = build_attn(qkv, batch, seq_len, num_heads, out, &ap);
  if (r) return r;
  ap.k_prescaled = 1;
  ap.bound_wq[0] = static_cast<const bf16*>(wq0); ap.bound_wk[0] = static_cast<const bf16*>(wk0);
  ap.bound_wq[1] = static_cast<const bf16*>(wq1); ap.bound_wk[1] = static_cast<const bf16*>(wk1);
  g_launch_count = 1;
  FFB_CUDA(launch_attention(ap, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_ln_modulate(const void* x, int num_batch, int rows_per_batch, int D, float eps, const void* shift1, const void* scale1,
                       void* out1, const void* shift2, const void* scale2, void* out2, long long mod_batch_stride, void* stream) {
  FFB_CHECK(x && shift1 && scale1 && out1, "null argument");
  LnModParams lp{};
  lp.x = static_cast<const bf16*>(x); lp.rows_per_batch = rows_per_batch; lp.num_batch = num_batch; lp.D = D; lp.eps = eps;
  lp.shift1 = static_cast<const bf16*>(shift1); lp.scale1 = static_cast<const bf16*>(scale1); lp.out1 = static_cast<bf16*>(out1);
  lp.shift2 = static_cast<const bf16*>(shift2); lp.scale2 = static_cast<const bf16*>(scale2); lp.out2 = static_cast<bf16*>(out2);
  lp.mod_batch_stride = static_cast<long>(mod_batch_stride);
  g_launch_count = 1;
  FFB_CUDA(launch_ln_modulate(lp, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_small_linear(const void* in, int batch, int K, long long in_stride, const void* W, const void* bias, int N, void* out,
                        long long out_stride, const void* addend, long long addend_stride, int silu_input, void* stream) {
  FFB_CHECK(in && W && out, "null argument");
  SmallLinearParams sp{};
  sp.in = static_cast<const bf16*>(in); sp.batch = batch; sp.K = K; sp.in_stride = static_cast<long>(in_stride);
  sp.W = static_cast<const bf16*>(W); sp.bias = static_cast<const bf16*>(bias); sp.N = N; sp.out = static_cast<bf16*>(out);
  sp.out_stride = static_cast<long>(out_stride); sp.addend = static_cast<const bf16*>(addend);
  sp.addend_stride = static_cast<long>(addend_stride); sp.silu_input = silu_input;
  g_launch_count = 1;
  FFB_CUDA(launch_small_linear(sp, static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_sde_step_ex(const void* noise_pred_bf16, const void* latents_fp16, int B, int C, int H, int W, const ffb200_step_coef* coef,
                    const float* noise, unsigned long long seed, int step_index, const void* next_latents_fp16, void* out_next_fp16,
                    float* out_mean, float* out_log_prob, int* overflow_flag, int storage_dtype, void* stream) {
  FFB_CHECK(storage_dtype == LAT_F16 || storage_dtype == LAT_BF16 || storage_dtype == LAT_F32, "latent dtype: 0 = fp16, 1 = bf16, 2 = fp32");
  FFB_CHECK(noise_pred_bf16 && latents_fp16 && coef, "null argument");
  FFB_CHECK(W % 4 == 0, "W must be a multiple of 4");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // scratch: coefficient entry + block partials (stream-ordered allocation keeps the call re-entrant)
  StepCoef* d_c = nullptr; float* d_part = nullptr;
  FFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_c), sizeof(StepCoef) * (step_index + 1), st));
  {
    cudaError_t ae = cudaMallocAsync(reinterpret_cast<void**>(&d_part), static_cast<size_t>(B) * 64 * sizeof(float), st);
    if (ae == cudaSuccess) ae = cudaMemcpyAsync(d_c + step_index, coef, sizeof(StepCoef), cudaMemcpyHostToDevice, st);
    if (ae != cudaSuccess) {                       // no leak on the error path
      cudaFreeAsync(d_c, st);
      if (d_part) cudaFreeAsync(d_part, st);
      FFB_CUDA(ae);
    }
  }
  SdeStepParams sp; memset(&sp, 0, sizeof(sp));
  sp.v_direct = static_cast<const bf16*>(noise_pred_bf16); sp.B = B; sp.C = C; sp.H = H; sp.W = W; sp.patch = 1;
  sp.storage = storage_dtype;
  sp.x = latents_fp16; sp.noise = noise; sp.noise_step_stride = 0; sp.seed = seed;
  sp.next_given = next_latents_fp16; sp.x_next = out_next_fp16;
  sp.mean_out = out_mean; sp.log_prob = out_log_prob; sp.logp_partial = d_part; sp.overflow_flag = overflow_flag;
  sp.coef_table = d_c; sp.coef_index = step_index;
  if (noise) sp.noise = noise - 0;  // single step: no per-step offset (stride 0)
  g_launch_count = 2;
  cudaError_t le = launch_sde_step(sp, st);
  cudaFreeAsync(d_c, st);
  cudaFreeAsync(d_part, st);
  FFB_CUDA(le);
  return 0;
}

int ffb200_sde_step(const void* noise_pred_bf16, const void* latents_fp16, int B, int C, int H, int W, const ffb200_step_coef* coef,
                    const float* noise, unsigned long long seed, int step_index, const void* next_latents_fp16, void* out_next_fp16,
                    float* out_mean, float* out_log_prob, int* overflow_flag, void* stream) {
  return ffb200_sde_step_ex(noise_pred_bf16, latents_fp16, B, C, H, W, coef, noise, seed, step_index, next_latents_fp16, out_next_fp16, out_mean,
                            out_log_prob, overflow_flag, LAT_F16, stream);
}

}  // extern "C"
