// Host side of the VAE decode that follows the rollout + its extern "C" entries (include/ffb200.h).
//
//   SD3_5Adapter.decode_latents      FF/models/stable_diffusion/sd3_5.py:161-172
//   AutoencoderKL.decode / Decoder   DF/models/autoencoders/autoencoder_kl.py, vae.py:279-316
//   ResnetBlock2D                    DF/models/resnet.py:319-377          UNetMidBlock2D / UpDecoderBlock2D   unets/unet_2d_blocks.py
//   Attention + AttnProcessor2_0     DF/models/attention_processor.py (heads = 1, group norm, residual connection)
//   Upsample2D                       DF/models/upsampling.py (nearest 2x + 3x3 conv)
//
// A decoder is, like the rollout plans, a fixed launch list for one geometry (batch, latent size): NHWC bf16 activations in four
// rotating workspace buffers, every convolution / linear an implicit GEMM on the tensor cores (vae_conv.cu), GroupNorm as a
// statistics pass + an apply pass that also performs the SiLU and the bf16 rounding the next convolution's input gets anyway.
// The single-head mid-block attention (head_dim = C = 512 does not fit the TMEM-resident attention kernels) runs per image as
// three GEMMs around an in-place row softmax: S = Q K^T (fp32), P = softmax(S) (bf16, over the same rows), O = P V + b_v with
// V^T produced directly by a GEMM whose A operand is W_v (the bias of V commutes with the row-stochastic P).
//
// STATUS: written after round 1's GPU budget was spent - compiled for sm_100a, NOT yet run on a GPU (see vae_conv.cu).
#include "common.cuh"
#include "kernels.h"
#include "../../include/ffb200.h"

struct VaeConvSpec {
  const void* x; int B, H, W, Cin, lda;
  const void* w; int ldb; long k_total; int Cout; int taps;
  const void* bias; int epi;
  void* out; int ldo; int n_store;
  const void* residual; int ldr;
  float* out_f32; long ldo_f32; float out_scale;
};

// pixel tile th x tw (tw * th = 128) that wastes the fewest out-of-image pixels; ties go to the wider tile (longer contiguous rows)
static int vae_pick_tw_log2(int H, int W) {
  int best = 7; long best_area = -1;
  for (int l = 7; l >= 3; --l) {
    const int tw = 1 << l, th = 128 >> l;
    const long area = static_cast<long>((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th;
    if (best_area < 0 || area < best_area) { best_area = area; best = l; }
  }
  return best;
}

static int build_conv(const VaeConvSpec& s, ConvParams* p) {
  memset(p, 0, sizeof(*p));
  FFB_CHECK(s.taps == 1 || s.taps == 9, "conv: taps must be 1 or 9");
  FFB_CHECK(s.lda % 8 == 0 && s.ldb % 8 == 0 && (s.taps == 1 || s.Cin % 8 == 0), "conv: pixel pitch / weight pitch (and Cin of a 3x3) must be multiples of 8");
  FFB_CHECK(s.B > 0 && s.H > 0 && s.W > 0 && s.Cout > 0, "conv: empty problem");
  p->kc = (s.Cin + 63) / 64;
  p->taps = s.taps;
  FFB_CHECK(s.taps == 1 || s.k_total == 9L * p->kc * 64, "conv: packed 3x3 weights must be [Cout][9][Cin padded to 64]");
  p->N = (s.Cout + 63) / 64 * 64;
  p->bn = gemm_pick_bn(p->N);
  p->tw_log2 = vae_pick_tw_log2(s.H, s.W);
  const int tw = 1 << p->tw_log2, th = 128 >> p->tw_log2;
  p->tiles_w = (s.W + tw - 1) / tw; p->tiles_h = (s.H + th - 1) / th;
  p->B = s.B; p->H = s.H; p->W = s.W;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(s.Cin), static_cast<uint64_t>(s.W), static_cast<uint64_t>(s.H), static_cast<uint64_t>(s.B)};
    const uint64_t str[3] = {static_cast<uint64_t>(s.lda), static_cast<uint64_t>(s.W) * s.lda, static_cast<uint64_t>(s.H) * s.W * s.lda};
    const uint32_t box[4] = {64, static_cast<uint32_t>(tw), static_cast<uint32_t>(th), 1};
    int r = make_tmap(&p->tmA, s.x, 4, dims, str, box);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(s.k_total), static_cast<uint64_t>(s.Cout)};
    const uint64_t str[1] = {static_cast<uint64_t>(s.ldb)};
    const uint32_t box[2] = {64, static_cast<uint32_t>(p->bn / 2)};
    int r = make_tmap(&p->tmB, s.w, 2, dims, str, box);
    if (r) return r;
  }
  p->band = num_sms() / 2;
  p->epi = s.epi;
  p->bias = static_cast<const bf16*>(s.bias);
  p->out = static_cast<bf16*>(s.out); p->ldo = s.ldo; p->n_store = s.n_store;
  p->residual = static_cast<const bf16*>(s.residual); p->ldr = s.ldr;
  p->out_f32 = s.out_f32; p->ldo_f32 = s.ldo_f32; p->out_scale = s.out_scale;
  if (s.epi == EPI_CONV_BIAS || s.epi == EPI_CONV_RESIDUAL) FFB_CHECK(s.n_store % 8 == 0 && s.ldo % 8 == 0, "conv: NHWC output channels must be a multiple of 8");
  if (s.epi == EPI_CONV_RESIDUAL) FFB_CHECK(s.residual != nullptr && s.ldr % 8 == 0, "conv: residual missing");
  if (s.epi == EPI_CONV_NCHW) FFB_CHECK(s.n_store >= 1 && s.n_store <= 8, "conv: planar output supports up to 8 channels");
  if (s.epi == EPI_CONV_F32) FFB_CHECK(s.out_f32 != nullptr && s.ldo_f32 % 4 == 0 && s.n_store % 4 == 0, "conv: fp32 output pitch");
  return 0;
}

struct ffb200_vae_decoder {
  ffb200_vae_config cfg;
  int B, h, w;
  std::vector<void*> allocs;
  long long ws_bytes;
  bf16* buf[4];
  bf16* z;            // NHWC latents, channels padded to 64
  float* scores;      // one image's attention scores [S][S_pad] fp32 (overwritten in place by the bf16 probabilities)
  double* stats;      // GroupNorm accumulators, one [B][Cmax][2] slot per norm layer
  size_t stats_bytes;
  std::vector<Op> ops;        // everything between the latent prologue and conv_out
  ConvParams final_conv;      // conv_out (planar image store; `out` is set per call)
};

namespace {

struct VaeBuilder {
  ffb200_vae_decoder* d;
  const void* const* w; int n_w; int wi;
  bool dry;
  long need;          // largest activation tensor (elements) seen by the walk
  long score_elems;   // fp32 elements of the attention score buffer
  int gn_layers; int cmax;
  bf16 *X, *T1, *T2, *T3;
  int err;

  const void* next() { const void* p = (dry || wi >= n_w) ? nullptr : w[wi]; ++wi; return p; }
  void track(long elems) { if (elems > need) need = elems; }

  void gn(const bf16* x, bf16* out, long P, int C, int silu, const void* gamma, const void* beta) {
    const int slot = gn_layers++;
    if (C > cmax) cmax = C;
    track(static_cast<long>(d->B) * P * C);
    if (dry || err) return;
    GroupNormParams gp{};
    gp.x = x; gp.out = out; gp.B = d->B; gp.P = P; gp.C = C; gp.groups = d->cfg.norm_num_groups; gp.eps = 1e-6f; gp.silu = silu;
    gp.gamma = static_cast<const bf16*>(gamma); gp.beta = static_cast<const bf16*>(beta);
    gp.stats = d->stats + static_cast<long>(slot) * d->B * cmax_alloc * 2;
    d->ops.push_back([gp](cudaStream_t st) { ++g_launch_count; return launch_group_norm_stats(gp, st); });
    d->ops.push_back([gp](cudaStream_t st) { ++g_launch_count; return launch_group_norm_apply(gp, st); });
  }
  int cmax_alloc;

  void conv(const VaeConvSpec& s) {
    if (dry || err) return;
    ConvParams cp;
    err = build_conv(s, &cp);
    if (err) return;
    const int sms = num_sms();
    d->ops.push_back([cp, sms](cudaStream_t st) { ++g_launch_count; return launch_conv(cp, sms, st); });
  }
  // 3x3 / 1x1 convolution over the whole batch, NHWC in -> NHWC out
  void conv_nhwc(const bf16* x, int H, int W, int Cin, const void* wt, const void* bias, int Cout, int taps, bf16* out, const bf16* residual) {
    track(static_cast<long>(d->B) * H * W * Cout);
    VaeConvSpec s{};
    s.x = x; s.B = d->B; s.H = H; s.W = W; s.Cin = Cin; s.lda = Cin;
    s.w = wt; s.taps = taps; s.Cout = Cout;
    if (taps == 9) { s.k_total = 9L * ((Cin + 63) / 64) * 64; s.ldb = static_cast<int>(s.k_total); }
    else { s.k_total = Cin; s.ldb = Cin; }
    s.bias = bias; s.epi = residual ? EPI_CONV_RESIDUAL : EPI_CONV_BIAS;
    s.out = out; s.ldo = Cout; s.n_store = Cout; s.residual = residual; s.ldr = Cout;
    conv(s);
  }

  void resnet(int H, int W, int cin, int cout) {
    const void *g1 = next(), *b1 = next(), *c1w = next(), *c1b = next(), *g2 = next(), *b2 = next(), *c2w = next(), *c2b = next();
    const void *sw = nullptr, *sb = nullptr;
    if (cin != cout) { sw = next(); sb = next(); }
    const long P = static_cast<long>(H) * W;
    gn(X, T1, P, cin, 1, g1, b1);
    conv_nhwc(T1, H, W, cin, c1w, c1b, cout, 9, T2, nullptr);
    gn(T2, T1, P, cout, 1, g2, b2);
    const bf16* res = X;
    if (cin != cout) { conv_nhwc(X, H, W, cin, sw, sb, cout, 1, T3, nullptr); res = T3; }
    conv_nhwc(T1, H, W, cout, c2w, c2b, cout, 9, T2, res);
    std::swap(X, T2);
  }

  void attention(int H, int W, int C) {
    const void *gg = next(), *gb = next(), *qkw = next(), *qkb = next(), *vw = next(), *vb = next(), *ow = next(), *ob = next();
    const int S = H * W, S_pad = (S + 63) / 64 * 64, B = d->B;
    track(static_cast<long>(B) * S * 2 * C);
    track(static_cast<long>(B) * C * S_pad);
    if (static_cast<long>(S) * S_pad > score_elems) score_elems = static_cast<long>(S) * S_pad;
    gn(X, T1, S, C, 0, gg, gb);
    conv_nhwc(T1, H, W, C, qkw, qkb, 2 * C, 1, T2, nullptr);                       // [B, S, 2C] = q | k
    for (int b = 0; b < B && !dry && !err; ++b) {                                  // V^T[b] = W_v x_b^T  -> [C][S_pad]
      VaeConvSpec s{};
      s.x = vw; s.B = 1; s.H = 1; s.W = C; s.Cin = C; s.lda = C;
      s.w = T1 + static_cast<long>(b) * S * C; s.ldb = C; s.k_total = C; s.Cout = S; s.taps = 1;
      s.epi = EPI_CONV_BIAS; s.out = T3 + static_cast<long>(b) * C * S_pad; s.ldo = S_pad; s.n_store = S_pad;
      conv(s);
    }
    for (int b = 0; b < B && !dry && !err; ++b) {
      const bf16* qk = T2 + static_cast<long>(b) * S * 2 * C;
      VaeConvSpec s{};                                                              // scores = q k^T / sqrt(C), fp32
      s.x = qk; s.B = 1; s.H = 1; s.W = S; s.Cin = C; s.lda = 2 * C;
      s.w = qk + C; s.ldb = 2 * C; s.k_total = C; s.Cout = S; s.taps = 1;
      s.epi = EPI_CONV_F32; s.out_f32 = d->scores; s.ldo_f32 = S_pad; s.n_store = S_pad; s.out_scale = 1.0f / sqrtf(static_cast<float>(C));
      conv(s);
      float* sc = d->scores;
      d->ops.push_back([sc, S, S_pad](cudaStream_t st) { ++g_launch_count; return launch_softmax_rows_inplace(sc, S, S, S_pad, st); });
      VaeConvSpec o{};                                                              // O = P V + b_v
      // K extent = S exactly: the bytes behind a row's S probabilities are stale fp32 score halves (possibly NaN patterns) and must
      // come back from the TMA unit as zeros, not be multiplied by the zero-filled V^T columns
      o.x = d->scores; o.B = 1; o.H = 1; o.W = S; o.Cin = S; o.lda = 2 * S_pad;
      o.w = T3 + static_cast<long>(b) * C * S_pad; o.ldb = S_pad; o.k_total = S; o.Cout = C; o.taps = 1;
      o.bias = vb; o.epi = EPI_CONV_BIAS; o.out = T1 + static_cast<long>(b) * S * C; o.ldo = C; o.n_store = C;
      conv(o);
    }
    conv_nhwc(T1, H, W, C, ow, ob, C, 1, T2, X);                                    // to_out + residual
    std::swap(X, T2);
  }

  void walk() {
    const ffb200_vae_config& c = d->cfg;
    const int nb = c.num_blocks, top = c.block_out_channels[nb - 1];
    int H = d->h, W = d->w;
    const void *ciw = next(), *cib = next();
    conv_nhwc(d->z, H, W, 64, ciw, cib, top, 9, X, nullptr);
    resnet(H, W, top, top);
    attention(H, W, top);
    resnet(H, W, top, top);
    int prev = top;
    for (int i = 0; i < nb; ++i) {
      const int ch = c.block_out_channels[nb - 1 - i];
      for (int j = 0; j <= c.layers_per_block; ++j) resnet(H, W, j == 0 ? prev : ch, ch);
      if (i != nb - 1) {
        const void *uw = next(), *ub = next();
        track(static_cast<long>(d->B) * 4 * H * W * ch);
        if (!dry && !err) {
          const bf16* src = X; bf16* dst = T1; const int B = d->B, h0 = H, w0 = W;
          d->ops.push_back([src, dst, B, h0, w0, ch](cudaStream_t st) { ++g_launch_count; return launch_upsample2x_nhwc(src, dst, B, h0, w0, ch, st); });
        }
        H *= 2; W *= 2;
        conv_nhwc(T1, H, W, ch, uw, ub, ch, 9, T2, nullptr);
        std::swap(X, T2);
      }
      prev = ch;
    }
    const void *ng = next(), *nbeta = next(), *cow = next(), *cob = next();
    gn(X, T1, static_cast<long>(H) * W, prev, 1, ng, nbeta);
    if (!dry && !err) {
      VaeConvSpec s{};
      s.x = T1; s.B = d->B; s.H = H; s.W = W; s.Cin = prev; s.lda = prev;
      s.w = cow; s.taps = 9; s.k_total = 9L * ((prev + 63) / 64) * 64; s.ldb = static_cast<int>(s.k_total); s.Cout = c.out_channels;
      s.bias = cob; s.epi = EPI_CONV_NCHW; s.out = T2 /* replaced per call */; s.ldo = 0; s.n_store = c.out_channels;
      err = build_conv(s, &d->final_conv);
    }
  }
};

int vae_check_config(const ffb200_vae_config* c) {
  FFB_CHECK(c != nullptr, "vae: null config");
  FFB_CHECK(c->num_blocks >= 1 && c->num_blocks <= 8, "vae: num_blocks must be 1..8");
  FFB_CHECK(c->latent_channels >= 1 && c->latent_channels <= 64, "vae: latent_channels must be 1..64");
  FFB_CHECK(c->out_channels >= 1 && c->out_channels <= 8, "vae: out_channels must be 1..8");
  FFB_CHECK(c->layers_per_block >= 0 && c->norm_num_groups >= 1, "vae: layers_per_block / norm_num_groups");
  for (int i = 0; i < c->num_blocks; ++i)
    FFB_CHECK(c->block_out_channels[i] >= 8 && c->block_out_channels[i] % 8 == 0 && c->block_out_channels[i] % c->norm_num_groups == 0,
              "vae: block_out_channels must be multiples of 8 and of norm_num_groups");
  FFB_CHECK(c->block_out_channels[c->num_blocks - 1] % 64 == 0, "vae: the mid-block width must be a multiple of 64");
  return 0;
}

}  // namespace

extern "C" {

int ffb200_vae_weight_count(const ffb200_vae_config* cfg) {
  if (vae_check_config(cfg)) return -1;
  ffb200_vae_decoder tmp{};
  tmp.cfg = *cfg; tmp.B = 1; tmp.h = 8; tmp.w = 8;
  VaeBuilder vb{};
  vb.d = &tmp; vb.dry = true;
  vb.walk();
  return vb.wi;
}

int ffb200_vae_decoder_create(const ffb200_vae_config* cfg, const void* const* weights, int n_weights, int batch, int lat_h, int lat_w,
                              ffb200_vae_decoder** out) {
  int r = vae_check_config(cfg);
  if (r) return r;
  FFB_CHECK(weights != nullptr && out != nullptr, "vae: null argument");
  FFB_CHECK(batch >= 1 && lat_h >= 1 && lat_w >= 1, "vae: batch and latent size must be positive");
  {
    int dev = 0, major = 0;
    FFB_CUDA(cudaGetDevice(&dev));
    FFB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    FFB_CHECK(major == 10, "vae: this library only runs on compute capability 10.x (B200)");
  }
  FFB_CHECK(static_cast<long>(lat_h) * lat_w <= 16384, "vae: mid-block attention supports up to 16384 latent pixels (1024^2 images)");
  ffb200_vae_decoder* d = new ffb200_vae_decoder();
  d->cfg = *cfg; d->B = batch; d->h = lat_h; d->w = lat_w;
  auto cleanup = [&](int code) { ffb200_vae_decoder_destroy(d); return code; };
  VaeBuilder dry{};
  dry.d = d; dry.dry = true;
  dry.walk();
  if (dry.wi != n_weights) return cleanup(fail(-1, "vae: wrong number of weight pointers (see ffb200_vae_weight_count)"));
  for (int i = 0; i < n_weights; ++i)
    if (weights[i] == nullptr) return cleanup(fail(-1, "vae: null weight pointer"));
  auto alloc = [&](void** p, size_t bytes) -> int {
    bytes = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaMalloc(vae workspace)");
    d->allocs.push_back(*p);
    d->ws_bytes += static_cast<long long>(bytes);
    return 0;
  };
  for (int i = 0; i < 4; ++i)
    if ((r = alloc(reinterpret_cast<void**>(&d->buf[i]), static_cast<size_t>(dry.need) * 2))) return cleanup(r);
  if ((r = alloc(reinterpret_cast<void**>(&d->z), static_cast<size_t>(batch) * lat_h * lat_w * 64 * 2))) return cleanup(r);
  if ((r = alloc(reinterpret_cast<void**>(&d->scores), static_cast<size_t>(dry.score_elems) * 4))) return cleanup(r);
  d->stats_bytes = static_cast<size_t>(dry.gn_layers) * batch * dry.cmax * 2 * sizeof(double);
  if ((r = alloc(reinterpret_cast<void**>(&d->stats), d->stats_bytes))) return cleanup(r);
  VaeBuilder vb{};
  vb.d = d; vb.w = weights; vb.n_w = n_weights; vb.dry = false; vb.cmax_alloc = dry.cmax;
  vb.X = d->buf[0]; vb.T1 = d->buf[1]; vb.T2 = d->buf[2]; vb.T3 = d->buf[3];
  vb.walk();
  if (vb.err) return cleanup(vb.err);
  *out = d;
  return 0;
}

void ffb200_vae_decoder_destroy(ffb200_vae_decoder* d) {
  if (!d) return;
  for (void* p : d->allocs) cudaFree(p);
  delete d;
}

long long ffb200_vae_decoder_workspace_bytes(const ffb200_vae_decoder* d) { return d ? d->ws_bytes : 0; }

int ffb200_vae_decode(ffb200_vae_decoder* d, const void* latents_f16, void* image_bf16, void* stream) {
  FFB_CHECK(d != nullptr && latents_f16 != nullptr && image_bf16 != nullptr, "vae_decode: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launch_count = 1;
  FFB_CUDA(cudaMemsetAsync(d->stats, 0, d->stats_bytes, st));
  FFB_CUDA(launch_vae_prep_latents(static_cast<const __half*>(latents_f16), d->z, d->B, d->cfg.latent_channels, d->h, d->w, 64,
                                   d->cfg.scaling_factor, d->cfg.shift_factor, st));
  for (auto& op : d->ops) FFB_CUDA(op(st));
  ConvParams fc = d->final_conv;
  fc.out = static_cast<bf16*>(image_bf16);
  ++g_launch_count;
  FFB_CUDA(launch_conv(fc, num_sms(), st));
  return 0;
}

int ffb200_conv2d_nhwc(const void* x, const void* w_packed, const void* bias, const void* residual, void* out, int B, int H, int W,
                       int Cin, int Cout, int taps, void* stream) {
  FFB_CHECK(x && w_packed && out, "conv2d_nhwc: null argument");
  FFB_CHECK(Cout % 8 == 0, "conv2d_nhwc: Cout must be a multiple of 8");
  VaeConvSpec s{};
  s.x = x; s.B = B; s.H = H; s.W = W; s.Cin = Cin; s.lda = Cin;
  s.w = w_packed; s.taps = taps; s.Cout = Cout;
  if (taps == 9) { s.k_total = 9L * ((Cin + 63) / 64) * 64; s.ldb = static_cast<int>(s.k_total); }
  else { s.k_total = Cin; s.ldb = Cin; }
  s.bias = bias; s.epi = residual ? EPI_CONV_RESIDUAL : EPI_CONV_BIAS;
  s.out = out; s.ldo = Cout; s.n_store = Cout; s.residual = residual; s.ldr = Cout;
  ConvParams cp;
  int r = build_conv(s, &cp);
  if (r) return r;
  g_launch_count = 1;
  FFB_CUDA(launch_conv(cp, num_sms(), static_cast<cudaStream_t>(stream)));
  return 0;
}

int ffb200_group_norm_nhwc(const void* x, const void* gamma, const void* beta, void* out, int B, long long P, int C, int groups,
                           float eps, int silu, void* workspace, void* stream) {
  FFB_CHECK(x && gamma && beta && out && workspace, "group_norm_nhwc: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GroupNormParams gp{};
  gp.x = static_cast<const bf16*>(x); gp.out = static_cast<bf16*>(out); gp.B = B; gp.P = static_cast<long>(P); gp.C = C; gp.groups = groups;
  gp.eps = eps; gp.silu = silu; gp.gamma = static_cast<const bf16*>(gamma); gp.beta = static_cast<const bf16*>(beta);
  gp.stats = static_cast<double*>(workspace);
  FFB_CUDA(cudaMemsetAsync(workspace, 0, static_cast<size_t>(B) * C * 2 * sizeof(double), st));
  g_launch_count = 2;
  FFB_CUDA(launch_group_norm_stats(gp, st));
  FFB_CUDA(launch_group_norm_apply(gp, st));
  return 0;
}

}  // extern "C"
