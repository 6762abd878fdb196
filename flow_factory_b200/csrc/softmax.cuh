// Online-softmax building blocks shared by the attention kernels (attention.cu: head_dim 64, attention_d128.cu: head_dim 128):
// thread = query row = TMEM lane; scores arrive as fp32 register blocks of 64 columns.
//
// What the SIMT side has to do per score is the whole cost of these kernels (at head_dim 64 the tensor core needs 8 cycles per SM for 256
// scores, the MUFU unit alone 16), so the routine below is built around doing as little as possible per element:
//   * no per-tile row maximum.  Online softmax needs A reference that keeps 2^(x - ref) inside the fp32 / bf16 exponent range (8 bits
//     both), not THE maximum.  Only the first tile of a row block looks at its maximum (to choose the reference); afterwards the reference
//     moves by an exact power of two when the running sum - an upper bound of every P so far - passes 2^64.  (round 2 measurement of this
//     alone: +5 % at head_dim 64, +4.5 % at 128.)
//   * reference ZERO whenever the first tile's scores are within 2^+-32 (always, for RMS-normed q / k): x - 0 needs no instruction.
//   * scores that already ARE base-2 exponents (`pre`): the engines fold softmax_scale * log2(e) into the key's RMSNorm weight multiply in
//     the QKV GEMM epilogue (one rounding to bf16, as the reference has), so the usual per-element FFMA (s * scale - max * scale) disappears:
//     MUFU-slot elements go from the TMEM load straight into ex2.
//   * part of the elements take a polynomial exp2 on the FMA pipe (the MUFU unit, 16 ex2/clk/SM, is the scarcest pipe), spread evenly
//     between the MUFU slots; the row sum uses packed f32x2 adds.
// Anything that could leave the representable range fails loudly (device error tag 0x6F), never silently.
#pragma once
#include "common.cuh"

namespace ffb {

// Which pairs of a 32-score block take the polynomial: kNum of every 8, evenly spread (3 -> {0,3,6}, 2 -> {0,4}, 4 -> {0,2,4,6}) or clustered
// (the first kNum of every 8).  A per-kernel choice - what ptxas makes of the mix differs between the two attention kernels (round 2
// measurements, profiles/r02_attention_experiments.md): with the lean steady-state loop of call 17 both kernels run best with 3 of 8 clustered
// (head_dim 64, call 18: 2 of 8 spread 881, 1 of 8 864, 3 of 8 spread 935, 3 of 8 clustered 947 TFLOP/s; before that loop 2 of 8 spread won);
// re-checked on the final loops in call 20: head_dim 64 2 / 3 / 4 clustered 965 / 990 / 955, head_dim 128 1379 / 1368 / 1343 -> 3 and 2.
constexpr int ATT_POLY_PERIOD = 8;
template <int kNum, bool kCluster>
struct PolyPolicy {
  static constexpr int num = kNum;
  __host__ __device__ static constexpr bool slot(int c) {
    return kCluster ? (c % ATT_POLY_PERIOD) < kNum : ((c % ATT_POLY_PERIOD) * kNum) % ATT_POLY_PERIOD < kNum;
  }
};
#ifdef FFB_ATT_POLY_NUM                 // A/B builds only: one policy for both kernels
#ifdef FFB_ATT_POLY_CLUSTER
using PolyD64 = PolyPolicy<FFB_ATT_POLY_NUM, true>;
#else
using PolyD64 = PolyPolicy<FFB_ATT_POLY_NUM, false>;
#endif
using PolyD128 = PolyD64;
#else
using PolyD64 = PolyPolicy<3, true>;
using PolyD128 = PolyPolicy<2, true>;
#endif
// general path (unscaled keys: op-level entry, diffusers backend hook): the per-pair FFMA2 of the exponent already loads the FMA pipe
#ifdef FFB_ATT_POLY_NUM
using PolyD64G = PolyD64;
using PolyD128G = PolyD128;
#else
using PolyD64G = PolyPolicy<2, false>;
using PolyD128G = PolyPolicy<2, true>;
#endif

constexpr float ATT_REF_ZERO_BAND = 32.0f;        // first-tile |max exponent| up to which the reference stays 0
constexpr float ATT_SHIFT_AT = 18446744073709551616.0f;        // 2^64: running sum at which the reference moves
constexpr float ATT_FAIL_AT = 7.922816251426434e28f;           // 2^96: more than 2^90 of growth inside ONE tile - refuse (tag 0x6F)

// 2^x for a PAIR on the FMA / ALU pipes: 2^x = 2^round(x) * p(x - round(x)), p = degree-3 minimax of 2^f on [-0.5, 0.5] (max rel. error
// 7.5e-5, far below the bf16 rounding of P); round() through the 1.5*2^23 magic-number add, exponent inserted with one shift-add.
// Valid for |x| <= 126 ONLY (the exponent insert wraps beyond): the caller checks the polynomial slots of a tile with one 3-input max
// per pair (softmax_poly_range_ok) and sends an out-of-range tile through MUFU instead - cheaper than clamping every element.
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& e0, float& e1) {
  const uint64_t xc = pack_f32x2(x0, x1);
  const uint64_t xr = fadd2(xc, pack_f32x2(12582912.0f, 12582912.0f));    // round(x) sits in the low mantissa bits
  const uint64_t r = fadd2(xr, pack_f32x2(-12582912.0f, -12582912.0f));   // round(x) as float
  const uint64_t f = ffma2(r, pack_f32x2(-1.0f, -1.0f), xc);              // x - round(x) in [-0.5, 0.5]
  uint64_t p = ffma2(f, pack_f32x2(0.05517132207751274f, 0.05517132207751274f), pack_f32x2(0.24261054396629333f, 0.24261054396629333f));
  p = ffma2(p, f, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  p = ffma2(p, f, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
  float p0, p1, r0, r1;
  unpack_f32x2(p, p0, p1);
  unpack_f32x2(xr, r0, r1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}

// exp2 of the 32 scores of `a` into 16 packed bf16 pairs, partial row sums into sums2.  kPre: the scores are the exponents (pre-scaled keys,
// reference 0); else exponent = s * sc - ref.  kPoly: the polynomial slots are in use (false: every element through MUFU).
template <bool kPre, bool kPoly, bool kSum, class Poly>
__device__ __forceinline__ void softmax_exp32(const uint32_t (&a)[32], uint64_t sc2, uint64_t mneg2, uint64_t (&sums2)[2],
                                              uint32_t (&pk)[16]) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float t0 = __uint_as_float(a[2 * c]), t1 = __uint_as_float(a[2 * c + 1]);
    if (!kPre) unpack_f32x2(ffma2(pack_f32x2(t0, t1), sc2, mneg2), t0, t1);
    float e0, e1;
    if (kPoly && Poly::slot(c)) {
      exp2_poly_pair(t0, t1, e0, e1);
    } else {
      e0 = ex2_approx(t0); e1 = ex2_approx(t1);
    }
    if (kSum) sums2[c & 1] = fadd2(sums2[c & 1], pack_f32x2(e0, e1));
    pk[c] = pack_bf16x2(e0, e1);
  }
}

// max |exponent| over the polynomial slots of a tile (one 3-input max per pair; |.| is an operand modifier): the polynomial is used only
// if it is <= 126.  -inf (masked keys), inf and NaN fail the test and take the MUFU path, which handles them.
template <bool kPre, class Poly>
__device__ __forceinline__ float softmax_poly_absmax(const uint32_t (&s0)[32], const uint32_t (&s1)[32], float sce, float m_run) {
  float mx[2] = {0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (Poly::slot(c)) {
      mx[0] = fmax3(mx[0], fabsf(__uint_as_float(s0[2 * c])), fabsf(__uint_as_float(s0[2 * c + 1])));
      mx[1] = fmax3(mx[1], fabsf(__uint_as_float(s1[2 * c])), fabsf(__uint_as_float(s1[2 * c + 1])));
    }
  }
  const float a = fmaxf(mx[0], mx[1]);
  return kPre ? a : fmaf(a, sce, fabsf(m_run));          // |s * sc - ref| <= |s| sc + |ref|
}

struct SoftmaxState {
  float m_run = -INFINITY;   // the reference, in EXPONENT units (score * sc); set by the first tile
  float l_run = 0.f;         // running row sum relative to it
  bool zero_ref = false;     // warp-uniform: every row of the warp still has reference 0
};

// Per-tile decisions of one thread (= one query row), taken once the 64 scores of the tile are in registers.
struct SoftmaxTile {
  bool rescale;        // warp-uniform: the accumulator has to be scaled by `alpha` (first tile: alpha = 0, nothing accumulated yet)
  float alpha;
  bool fast;           // warp-uniform: the scores are the exponents (pre-scaled keys, reference 0)
  bool poly;           // warp-uniform: the polynomial slots may be used on this tile (all their exponents within +-126)
  uint64_t sc2, mneg2; // general path: exponent = s * sc - ref, as packed pairs
  uint64_t sums2[2];   // 4 partial row sums of this tile as two packed pairs
};

// Online-softmax step of one thread over a 128 x 64 block of scores held in registers (s0: columns 0-31, s1: 32-63), in three parts so
// that the attention kernels can put the TMEM load of the NEXT tile's scores between the two halves (the registers of a half are dead as
// soon as its exponentials are taken):
//   softmax_begin : masking of keys beyond the sequence, reference policy (first tile / growth), polynomial range check
//   softmax_half  : exp2 + partial sums + bf16 packing of 32 columns (16 packed registers: stored to TMEM per half)
//   softmax_end   : running-sum update
//   sc    : softmax_scale * log2(e); ignored when `pre`.
//   pre   : warp-uniform; the scores already are base-2 exponents (keys pre-scaled by the producer).
//   first : warp-uniform; first KV tile of the row block.
//   bounded : warp-uniform; every exponent of this launch is proven to lie within +-126 as long as the reference is 0 (fast path).
template <class Poly, class PolyG = Poly>
__device__ __forceinline__ void softmax_begin(uint32_t (&s0)[32], uint32_t (&s1)[32], int kv_valid, float sc, bool pre, bool first,
                                              SoftmaxState& st, SoftmaxTile& t, bool bounded = false) {
  if (kv_valid < 64) {
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (c >= kv_valid) s0[c] = 0xFF800000u;      // -inf: key beyond the sequence
      if (32 + c >= kv_valid) s1[c] = 0xFF800000u;
    }
  }
  const float sce = pre ? 1.0f : sc;
  t.alpha = 1.0f;
  t.rescale = false;
  t.poly = Poly::num > 0 || PolyG::num > 0;
  t.sums2[0] = t.sums2[1] = 0ull;
  t.sc2 = pack_f32x2(sce, sce);
  if (first) {
    // first tile of the row block: its exact maximum (8 independent 3-input chains) chooses the reference
    float mxs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mxs[i] = -INFINITY;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      mxs[(c >> 1) & 7] = fmax3(mxs[(c >> 1) & 7], __uint_as_float(s0[c]), __uint_as_float(s0[c + 1]));
      mxs[(c >> 1) & 7] = fmax3(mxs[(c >> 1) & 7], __uint_as_float(s1[c]), __uint_as_float(s1[c + 1]));
    }
    const float mt = fmax3(fmax3(mxs[0], mxs[1], mxs[2]), fmax3(mxs[3], mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])) * sce;
    st.zero_ref = __all_sync(0xffffffffu, fabsf(mt) <= ATT_REF_ZERO_BAND);     // also false for a NaN / inf maximum
    st.m_run = st.zero_ref ? 0.0f : mt;
    st.l_run = 0.f;
    t.alpha = 0.0f;                                 // nothing accumulated yet
    t.rescale = true;
  }
  t.fast = pre && st.zero_ref;
  // ONE vote per tile covers both rare events - "my running sum passed 2^64" and "a polynomial slot of this tile is outside +-126" -
  // and only a warp that saw one of them sorts out which (the common tile pays a single VOTE + branch for its bookkeeping)
  const bool grow = !first && !(st.l_run <= ATT_SHIFT_AT);       // also true for inf / NaN
  bool over = false;
  // `bounded` (warp-uniform): the launch PROVED |exponent| <= 126 for pre-scaled keys with reference 0 (AttnParams::bound_w*): no range guard
  if ((Poly::num > 0 || PolyG::num > 0) && !(bounded && t.fast)) {   // general path: its own slot policy (one more FFMA2 per pair there)
    const float amax = t.fast ? softmax_poly_absmax<true, Poly>(s0, s1, sce, 0.f) : softmax_poly_absmax<false, PolyG>(s0, s1, sce, st.m_run);
    over = !(amax <= 126.0f);
  }
  if (__any_sync(0xffffffffu, grow || over)) {
    if (__any_sync(0xffffffffu, grow)) {
      if (!(st.l_run < ATT_FAIL_AT)) mbar_timeout(0x6F);   // a score jumped by more than 2^90 within one tile (or is not finite): fail loudly
      if (grow) {
        const int e = static_cast<int>((__float_as_uint(st.l_run) >> 23) & 0xFF) - 127;   // floor(log2(l_run)) >= 64
        t.alpha = __uint_as_float(static_cast<uint32_t>(127 - e) << 23);                  // 2^-e, exact
        st.m_run += static_cast<float>(e);                                                 // reference up by e exponent units, exact
      }
      st.zero_ref = false;
      t.fast = false;
      t.rescale = true;
      // the range check was made against the old reference: re-evaluate it (general path) against the new one
      over = PolyG::num > 0 && !(softmax_poly_absmax<false, PolyG>(s0, s1, sce, st.m_run) <= 126.0f);
    }
    t.poly = (Poly::num > 0 || PolyG::num > 0) && !__any_sync(0xffffffffu, over);
  }
  t.mneg2 = pack_f32x2(-st.m_run, -st.m_run);
}

template <class Poly, class PolyG = Poly, bool kSum = true>
__device__ __forceinline__ void softmax_half(const uint32_t (&a)[32], SoftmaxTile& t, uint32_t (&pk)[16]) {
  if (t.fast) {
    if (t.poly && Poly::num > 0) softmax_exp32<true, true, kSum, Poly>(a, t.sc2, t.mneg2, t.sums2, pk);
    else softmax_exp32<true, false, kSum, Poly>(a, t.sc2, t.mneg2, t.sums2, pk);
  } else {
    if (t.poly && PolyG::num > 0) softmax_exp32<false, true, kSum, PolyG>(a, t.sc2, t.mneg2, t.sums2, pk);
    else softmax_exp32<false, false, kSum, PolyG>(a, t.sc2, t.mneg2, t.sums2, pk);
  }
}

__device__ __forceinline__ void softmax_end(SoftmaxState& st, const SoftmaxTile& t) {
  float sa, sb, sc_, sd;
  unpack_f32x2(t.sums2[0], sa, sb);
  unpack_f32x2(t.sums2[1], sc_, sd);
  st.l_run = st.l_run * t.alpha + ((sa + sb) + (sc_ + sd));
}

// End-of-row check: a sum that left the supported range during the LAST tile (no later tile to notice it) fails loudly too.
__device__ __forceinline__ void softmax_final_check(float l_run) {
  if (!(l_run < ATT_FAIL_AT)) mbar_timeout(0x6F);
}


// Warp barrier between tcgen05.wait::ld / wait::st and the elected lane's arrive in the softmax tiles.  Both waits are .sync.aligned (the
// whole warp passes them together), so the barrier is arguably redundant; -DFFB_ATT_NO_SYNCWARP drops it (A/B only, not the product).
#ifdef FFB_ATT_NO_SYNCWARP
#define ATT_TILE_SYNCWARP() ((void)0)
#else
#define ATT_TILE_SYNCWARP() __syncwarp()
#endif

}  // namespace ffb
