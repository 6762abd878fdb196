// Online-softmax building blocks shared by the attention kernels (attention.cu: head_dim 64, attention_d128.cu: head_dim 128):
// thread = query row = TMEM lane; scores arrive as fp32 register blocks of 64 columns.
#pragma once
#include "common.cuh"

namespace ffb {

// exp2 on the FMA/ALU pipes for part of the elements (the MUFU unit, 16 ex2/clk/SM, is the binding resource at d = 64):
// 2^x = 2^round(x) * p(x - round(x)), p = degree-3 minimax of 2^f on [-0.5, 0.5] (max rel. error 7.5e-5, far below the
// bf16 rounding of P); round() through the 1.5*2^23 magic-number add, exponent inserted with one shift-add.
// Operates on a packed pair.  x must be <= ~+100; clamped below at -126.
__device__ __forceinline__ void exp2_poly_pair(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack_f32x2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const uint64_t xc = pack_f32x2(x0, x1);
  const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);
  const uint64_t xr = fadd2(xc, magic);                                   // round(x) sits in the low mantissa bits
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
constexpr int ATT_POLY_PERIOD = 8;   // of every ATT_POLY_PERIOD element pairs ...
#ifndef FFB_ATT_POLY_NUM
#define FFB_ATT_POLY_NUM 3           // product value; -DFFB_ATT_POLY_NUM=n only for the A/B builds of tools/gpu_maxfree.sh
#endif
constexpr int ATT_POLY_NUM = FFB_ATT_POLY_NUM;   // ... this many go through exp2_poly_pair, the rest through MUFU.EX2

// Online-softmax step of one thread (= one query row) over a 128 x 64 block of scores held in registers (s0: columns 0-31,
// s1: 32-63).  Updates the running max / sum, returns P as 32 packed bf16 pairs, the factor `alpha` by which the accumulator has
// to be scaled if the (warp-uniform) return value is true.  Shared by the d = 64 and d = 128 kernels.
// kSum = false (experimental/attention_summma.cu only): the row sum is NOT accumulated here - the tensor core produces it as an extra
// accumulator column (V extended by a panel of ones) - and l_run is left untouched.
template <bool kSum = true>
__device__ __forceinline__ bool softmax_block64(uint32_t (&s0)[32], uint32_t (&s1)[32], int kv_valid, float sc, float& m_run,
                                                float& l_run, uint32_t (&pk)[32], float& alpha) {
  // row max of this tile: 8 independent chains (a single serial fmax chain is 128 x 4 cycles of pure latency)
  float mxs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mxs[i] = -INFINITY;
  auto max32 = [&](uint32_t(&a)[32], int base) {
    if (kv_valid < 64) {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (base + c >= kv_valid) a[c] = 0xFF800000u;  // -inf: key beyond the sequence
    }
#pragma unroll
    for (int c = 0; c < 32; c += 2) mxs[(c >> 1) & 7] = fmax3(mxs[(c >> 1) & 7], __uint_as_float(a[c]), __uint_as_float(a[c + 1]));
  };
#ifdef FFB_ATT_MAXFREE
  // EXPERIMENT (not the product build; tools/gpu_maxfree.sh): the row-max pass costs ~14 % of the kernel (profiles/
  // r01_attention_whatif.md) and online softmax does not need the MAX as its reference, only A reference that keeps 2^((s - m) sc)
  // inside the fp32 / bf16 exponent range (both have 8 exponent bits).  So only the first tile of a row takes its exact maximum;
  // later tiles reuse the reference and move it by a power of two whenever the running sum - an upper bound of every P so far -
  // has grown past 2^24.  P may exceed 1 (by at most the growth of one tile); a score more than 127 / sc above the reference within
  // a single tile would overflow: that is detected (l_run = inf) and trapped by softmax_final_check / the next tile, never silent.
  bool rescale;
  alpha = 1.0f;
  if (__any_sync(0xffffffffu, m_run == -INFINITY)) {     // first tile of the row block: exact maximum
    max32(s0, 0); max32(s1, 32);
    const float mt = fmax3(fmax3(mxs[0], mxs[1], mxs[2]), fmax3(mxs[3], mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7]));
    const float mnew = fmaxf(m_run, mt);                 // per lane: a row that already has state keeps it consistent
    alpha = ex2_approx((m_run - mnew) * sc);             // 0 on the first tile
    m_run = mnew;
    rescale = true;
  } else {
    if (kv_valid < 64) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (c >= kv_valid) s0[c] = 0xFF800000u;
        if (32 + c >= kv_valid) s1[c] = 0xFF800000u;
      }
    }
    const bool grow = !(l_run <= 16777216.0f);            // also true for inf / NaN
    rescale = __any_sync(0xffffffffu, grow);
    if (rescale) {
      if (!(l_run < 3.0e38f)) mbar_timeout(0x6F);         // overflow inside one tile: fail loudly (tag 0x6F)
      if (grow) {
        const int e = static_cast<int>((__float_as_uint(l_run) >> 23) & 0xFF) - 127;   // floor(log2(l_run)) >= 24
        alpha = __uint_as_float(static_cast<uint32_t>(127 - e) << 23);                 // 2^-e, exact
        m_run += __fdividef(static_cast<float>(e), sc);                                 // reference up by e in exponent units
      }
    }
  }
#else
  max32(s0, 0); max32(s1, 32);
  const float mt = fmax3(fmax3(mxs[0], mxs[1], mxs[2]), fmax3(mxs[3], mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7]));
  // lazy rescale: adopt the new max only if some row of the warp grew by more than 2^8 (warp-uniform decision)
  const bool grow = (mt - m_run) * sc > 8.0f;           // true on the first tile (m_run = -inf)
  const bool rescale = __any_sync(0xffffffffu, grow);
  alpha = 1.0f;
  if (rescale) {
    const float mnew = fmaxf(m_run, mt);
    alpha = ex2_approx((m_run - mnew) * sc);             // 0 on the first tile
    m_run = mnew;
  }
#endif
  const uint64_t sc2 = pack_f32x2(sc, sc), mneg2 = pack_f32x2(-m_run * sc, -m_run * sc);
  uint64_t sums2[2] = {0ull, 0ull};              // 4 partial row sums as two packed pairs
  auto exp32 = [&](uint32_t(&a)[32], int quarter) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint64_t x2 = ffma2(pack_f32x2(__uint_as_float(a[2 * c]), __uint_as_float(a[2 * c + 1])), sc2, mneg2);
      float e0, e1;
      if ((c % ATT_POLY_PERIOD) < ATT_POLY_NUM) {
        exp2_poly_pair(x2, e0, e1);
      } else {
        float t0, t1;
        unpack_f32x2(x2, t0, t1);
        e0 = ex2_approx(t0); e1 = ex2_approx(t1);
      }
      if (kSum) sums2[c & 1] = fadd2(sums2[c & 1], pack_f32x2(e0, e1));
      pk[quarter * 16 + c] = pack_bf16x2(e0, e1);
    }
  };
  exp32(s0, 0); exp32(s1, 1);
  if (kSum) {
    float sa, sb, sc_, sd;
    unpack_f32x2(sums2[0], sa, sb);
    unpack_f32x2(sums2[1], sc_, sd);
    l_run = l_run * alpha + ((sa + sb) + (sc_ + sd));
  }
  return rescale;
}

// End-of-row check of the max-free experiment (no code in the product build).
__device__ __forceinline__ void softmax_final_check(float l_run) {
#ifdef FFB_ATT_MAXFREE
  if (!(l_run < 3.0e38f)) mbar_timeout(0x6F);
#else
  (void)l_run;
#endif
}

template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

}  // namespace ffb
