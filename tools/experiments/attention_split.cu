// NOT BUILT - measured alternative layouts of the head_dim-64 attention kernel (round 2, calls 6 / 7), kept for the record.
// It was compiled into libffb200.so behind FFB200_ATT_VARIANT={split2,row2}, passed tests/test_gpu_attention.py + tests/test_gpu_engine.py
// (78 / 47 tests) and measured on the bench shape (B=8, S=4429, H=24, pre-scaled keys):
//     row3   (product: 3 sub-tiles, one warp per 32 rows, single-buffered S)       833-838 TFLOP/s
//     split2 (2 sub-tiles, TWO warps per row = 4 per sub-partition, double-buffered S)   671 TFLOP/s
//     row2   (2 sub-tiles, one warp per 32 rows, double-buffered S)                      605 TFLOP/s
// Reading: neither more warps per sub-partition (split2) nor a score tile that never has to be waited for (row2, split2) pays for the
// third sub-tile they cost in TMEM; the per-tile skeleton (two mbarrier round trips, TMEM load / store) is per WARP, so halving the
// scores per warp doubles its share.  To rebuild: paste between the row kernel and launch_attention() in csrc/attention.cu.
#if 0
// =====================================================================================================================================
// Column-split variant (round 2): TWO softmax warps per query row, each owning 32 of the 64 keys of a tile.
//
// The kernel above keeps every pipe of the SM below 60 % (ncu, profiles/r02_ncu_attention_d64*.md: XU 56 %, issue slots 46 %, FMA 28 %,
// tensor 34 %): three softmax warps per SM sub-partition cannot cover each other's dependent-issue and MUFU-queue stalls.  Without a
// per-tile row maximum (softmax.cuh) the two halves of a row have nothing to tell each other per tile, so a row can be split between two
// warps for free: 16 softmax warps (4 per sub-partition) over two 128-row sub-tiles per CTA, and - the sub-tiles now need only 320 of the
// 512 TMEM columns - the score tile S is DOUBLE-buffered, so Q K^T runs two tiles ahead of the softmax and never waits for it.
//   warps 0-15 : softmax.  warp w: lane quadrant w % 4 (TMEM lanes), sub-tile (w / 4) % 2, key half w / 8.
//   then       : TMA producer, the MMA issuers of sub-tile 0 / 1, one idle warp.
// Agreement between the two halves of a row (the reference of online softmax must be the same in both, they feed one accumulator):
//   * first tile: the half maxima are exchanged through shared memory (one named barrier per CTA lifetime) - both halves take the same
//     decision (reference 0 or the row maximum);
//   * growth: each half watches its PARTIAL sum; "mine passed 2^64" is published in shared memory at the end of tile j, read by the partner
//     in tile j+1 (ordered by the p_full -> P V -> p_free chain both halves are on), and BOTH move the reference by exactly 2^64 at the start
//     of tile j+2.  A sum beyond 2^96 (a jump of > 2^32 inside those two tiles) fails loudly (0x6F), as in the row-per-thread kernel.
// =====================================================================================================================================
constexpr int AT2_NSUB = 2;
constexpr int AT2_QB = AT2_NSUB * ATT_BM;               // 256 query rows per CTA
constexpr int AT2_STAGES = 6;
constexpr int AT2_THREADS = 640;
constexpr int AT2_XCH_BYTES = AT2_NSUB * 2 * ATT_BM * 4;         // float exchange (first-tile maxima, final sums): [sub][half][row]
constexpr int AT2_FLAG_BYTES = AT2_NSUB * 2 * 2 * ATT_BM;        // growth flags: [sub][parity][half][row] bytes
constexpr int AT2_SMEM = AT2_NSUB * ATT_TILE_BYTES + 2 * AT2_STAGES * ATT_KV_BYTES + AT2_XCH_BYTES + AT2_FLAG_BYTES + 1024;
constexpr int AT2_TMEM_S = 0;        // S_x[buf] at (x * 2 + buf) * 64
constexpr int AT2_TMEM_P = 256;      // P_x at 256 + x * 32
constexpr int AT2_TMEM_O = 320;      // O_x at 320 + x * 64
constexpr float AT2_SHIFT = 64.0f;   // the reference moves by exactly 2^64

// One warp per 32 query rows, all 64 keys of a tile (kHalves == 1 of attention_split_kernel): the row-per-thread softmax of attention_kernel
// on the double-buffered S of the split layout.
__device__ __forceinline__ void attention_rowpair_softmax(const AttnParams& p, uint32_t tmem_base, int wq, int x, int lane, int q0, int head,
                                                          int b, int S, int n_tiles, uint64_t* s_full, uint64_t* s_free, uint64_t* p_full,
                                                          uint64_t* p_free, uint64_t* o_full) {
  const int r = wq * 32 + lane;
  const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
  const uint32_t tSx = tmem_base + lane_off + 0 /*AT2_TMEM_S*/ + x * 2 * ATT_BN;
  const uint32_t tPx = tmem_base + lane_off + 256 /*AT2_TMEM_P*/ + x * (ATT_BN / 2);
  const uint32_t tOx = tmem_base + lane_off + 320 /*AT2_TMEM_O*/ + x * ATT_D;
  const float sc = p.scale_log2;
  const bool pre = p.k_prescaled != 0;
  SoftmaxState sm;
  uint32_t s0[32], s1[32];
  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    mbar_wait(&s_full[x * 2 + buf], (j >> 1) & 1, 0x60);
    tc_fence_after();
    tmem_ld32(tSx + buf * ATT_BN, s0);
    tmem_ld32(tSx + buf * ATT_BN + 32, s1);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_free[x * 2 + buf]);
    SoftmaxTile t;
    softmax_begin<PolyD64>(s0, s1, S - j * ATT_BN, sc, pre, j == 0, sm, t);
    uint32_t pk[16];
    softmax_half<PolyD64>(s0, t, pk);
    if (j > 0) {
      mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
      tc_fence_after();
    }
    if (j > 0 && t.rescale) {
      uint32_t o0[32];
#pragma unroll 1
      for (int c = 0; c < ATT_D; c += 32) {
        tmem_ld32(tOx + c, o0);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * t.alpha);
        tmem_st32(tOx + c, o0);
      }
    }
    tmem_st16(tPx, pk);
    softmax_half<PolyD64>(s1, t, pk);
    tmem_st16(tPx + 16, pk);
    softmax_end(sm, t);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&p_full[x]);
  }
  mbar_wait(&o_full[x], 0, 0x69);
  tc_fence_after();
  const int q = q0 + x * ATT_BM + r;
  if (q < S) softmax_final_check(sm.l_run);
  const float inv = 1.0f / sm.l_run;
  bf16* dst = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(q) * p.inner_dim + head * ATT_D;
#pragma unroll 1
  for (int c = 0; c < ATT_D; c += 32) {
    uint32_t o0[32];
    tmem_ld32(tOx + c, o0);
    tmem_ld_wait();
    if (q < S) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(o0[g * 8 + 0]) * inv, __uint_as_float(o0[g * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(o0[g * 8 + 2]) * inv, __uint_as_float(o0[g * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(o0[g * 8 + 4]) * inv, __uint_as_float(o0[g * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(o0[g * 8 + 6]) * inv, __uint_as_float(o0[g * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst + c)[g] = o;
      }
    }
  }
}

template <int kHalves>      // 2: two warps per row (16 softmax warps); 1: one warp per row (8 softmax warps), same double-buffered S
__global__ void __launch_bounds__((8 * kHalves + 4) * 32, 1)
attention_split_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                            // [2 sub-tiles][128][64]
  uint8_t* sK = sQ + AT2_NSUB * ATT_TILE_BYTES;                  // [stages][64][64]
  uint8_t* sV = sK + AT2_STAGES * ATT_KV_BYTES;
  float* xch = reinterpret_cast<float*>(sV + AT2_STAGES * ATT_KV_BYTES);      // [sub][half][row]
  uint8_t* flags = reinterpret_cast<uint8_t*>(xch) + AT2_XCH_BYTES;           // [sub][parity][half][row]
  uint64_t* bars = reinterpret_cast<uint64_t*>(flags + AT2_FLAG_BYTES);
  uint64_t* q_full = bars;                          // 1
  uint64_t* k_full = bars + 1;                      // [ST]
  uint64_t* k_empty = k_full + AT2_STAGES;
  uint64_t* v_full = k_empty + AT2_STAGES;
  uint64_t* v_empty = v_full + AT2_STAGES;
  uint64_t* s_full = v_empty + AT2_STAGES;          // [sub][buf]
  uint64_t* s_free = s_full + 2 * AT2_NSUB;         // [sub][buf], 8 arrivals
  uint64_t* p_full = s_free + 2 * AT2_NSUB;         // [sub], 8 arrivals
  uint64_t* p_free = p_full + AT2_NSUB;             // [sub]
  uint64_t* o_full = p_free + AT2_NSUB;             // [sub]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + AT2_NSUB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) mbar_timeout(0xA11);
  const int q0 = blockIdx.x * AT2_QB, head = blockIdx.y, b = blockIdx.z;
  const int S = p.seq_len;
  const int n_tiles = (S + ATT_BN - 1) / ATT_BN;
  const int n_sub = min(AT2_NSUB, (S - q0 + ATT_BM - 1) / ATT_BM);

  constexpr int kTma = 8 * kHalves;                 // warp roles after the softmax warps: TMA producer, two MMA issuers, one idle
  if (warp == kTma && lane == 0) {
    tma_prefetch_desc(&p.tmQKV);
    tma_prefetch_desc(&p.tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT2_STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], n_sub);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], n_sub);
    }
    for (int i = 0; i < 2 * AT2_NSUB; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4 * kHalves); }
    for (int i = 0; i < AT2_NSUB; ++i) { mbar_init(&p_full[i], 4 * kHalves); mbar_init(&p_free[i], 1); mbar_init(&o_full[i], 1); }
    fence_barrier_init();
  }
  if (warp == kTma + 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kTma) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int cq = head * ATT_D, ck = p.inner_dim + head * ATT_D, cv = 2 * p.inner_dim + head * ATT_D;
      mbar_arrive_expect_tx(q_full, n_sub * ATT_TILE_BYTES);
      for (int x = 0; x < n_sub; ++x) tma_load_3d(sQ + x * ATT_TILE_BYTES, &p.tmQKV, q_full, cq, q0 + x * ATT_BM, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % AT2_STAGES;
        const uint32_t ph = (j / AT2_STAGES) & 1;
        mbar_wait_relaxed(&k_empty[st], ph ^ 1, 0x40);
        mbar_arrive_expect_tx(&k_full[st], ATT_KV_BYTES);
        tma_load_3d(sK + st * ATT_KV_BYTES, &p.tmKV, &k_full[st], ck, j * ATT_BN, b);
        mbar_wait_relaxed(&v_empty[st], ph ^ 1, 0x41);
        mbar_arrive_expect_tx(&v_full[st], ATT_KV_BYTES);
        tma_load_3d(sV + st * ATT_KV_BYTES, &p.tmKV, &v_full[st], cv, j * ATT_BN, b);
      }
    }
  } else if (warp == kTma + 1 || warp == kTma + 2) {
    // ===================== MMA issuers: warp kTma + 1 + x -> sub-tile x =====================
    const int x = warp - (kTma + 1);
    if (x < n_sub) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);   // P (TMEM) x V (MN-major)
      const uint32_t q_addr = smem_u32(sQ) + x * ATT_TILE_BYTES, sK_addr = smem_u32(sK), sV_addr = smem_u32(sV);
      const uint32_t tPx = tmem_base + AT2_TMEM_P + x * (ATT_BN / 2), tOx = tmem_base + AT2_TMEM_O + x * ATT_D;
      auto issue_qk = [&](int t) {                                       // S_x[t & 1] = Q K_t^T
        const int st = t % AT2_STAGES, buf = t & 1;
        mbar_wait(&k_full[st], (t / AT2_STAGES) & 1, 0x50);
        tc_fence_after();
        const uint32_t k_addr = sK_addr + st * ATT_KV_BYTES;
        const uint32_t tS = tmem_base + AT2_TMEM_S + (x * 2 + buf) * ATT_BN;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_bf16(tS, desc_kmajor_sw128(q_addr + k * 32), desc_kmajor_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[x * 2 + buf]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0, 0x52);
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % AT2_STAGES;
        mbar_wait(&v_full[st], (j / AT2_STAGES) & 1, 0x53);
        mbar_wait(&p_full[x], j & 1, 0x54);                              // both halves of P_x(j) are in TMEM (and any rescale of O_x done)
        tc_fence_after();
        const uint32_t v_addr = sV_addr + st * ATT_KV_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k) {
            const uint64_t db = desc_mnmajor_sw128(v_addr + k * 2048, ATT_KV_BYTES);
            umma_bf16_ts(tOx, tPx + k * 8, db, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[st]);
          umma_commit(&p_free[x]);
          if (j == n_tiles - 1) umma_commit(&o_full[x]);
        }
        __syncwarp();
        if (j + 2 < n_tiles) {                                            // S_x[j & 1] is free once both halves hold S(j) in registers
          mbar_wait(&s_free[x * 2 + (j & 1)], (j >> 1) & 1, 0x51);
          issue_qk(j + 2);
        }
      }
    }
  } else if (warp < 8 * kHalves) {
    // ===================== softmax: 32 keys of 32 query rows per warp and tile =====================
    const int wq = warp & 3, x = (warp >> 2) & 1, h = warp >> 3;
    if (kHalves == 1) {
      if (x < n_sub) attention_rowpair_softmax(p, tmem_base, wq, x, lane, q0, head, b, S, n_tiles, s_full, s_free, p_full, p_free, o_full);
    } else if (x < n_sub) {
      const int r = wq * 32 + lane;                                       // query row in the sub-tile == TMEM lane
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tSx = tmem_base + lane_off + AT2_TMEM_S + x * 2 * ATT_BN + 32 * h;
      const uint32_t tPx = tmem_base + lane_off + AT2_TMEM_P + x * (ATT_BN / 2) + 16 * h;
      const uint32_t tOx = tmem_base + lane_off + AT2_TMEM_O + x * ATT_D + 32 * h;      // this half's 32 output columns
      float* xmine = xch + (x * 2 + h) * ATT_BM + r;
      const float* xpeer = xch + (x * 2 + (1 - h)) * ATT_BM + r;
      const bool pre = p.k_prescaled != 0;
      const float sce = pre ? 1.0f : p.scale_log2;
      const uint64_t sc2 = pack_f32x2(sce, sce);
      float m_run = 0.f, l_half = 0.f;
      bool zero_ref = false, g_prev = false, pending = false;
      for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        mbar_wait(&s_full[x * 2 + buf], (j >> 1) & 1, 0x60);
        tc_fence_after();
        uint32_t s[32];
        tmem_ld32(tSx + buf * ATT_BN, s);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x * 2 + buf]);
        const int kv_valid = S - j * ATT_BN - 32 * h;                     // keys of this half that exist
        if (kv_valid < 32) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c >= kv_valid) s[c] = 0xFF800000u;
        }
        // ---- reference policy
        float alpha = 1.0f;
        bool rescale = false;
        if (j == 0) {
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 32; c += 2) mx[(c >> 1) & 3] = fmax3(mx[(c >> 1) & 3], __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
          *xmine = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          asm volatile("bar.sync %0, 256;" ::"r"(1 + x) : "memory");       // the 8 warps of this sub-tile
          const float mt = fmaxf(*xmine, *xpeer) * sce;                   // the row maximum, identical in both halves
          zero_ref = __all_sync(0xffffffffu, fabsf(mt) <= ATT_REF_ZERO_BAND);
          m_run = zero_ref ? 0.0f : mt;
        } else if (__any_sync(0xffffffffu, pending)) {                    // decided two tiles ago by either half of some row of this warp
          if (pending) { alpha = 5.421010862427522e-20f /* 2^-64 */; m_run += AT2_SHIFT; l_half *= alpha; }
          zero_ref = false;
          rescale = true;
          pending = false;
        }
        const bool fast = pre && zero_ref;
        bool poly = PolyD64::num > 0;
        if (poly) {
          float am[2] = {0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (PolyD64::slot(c)) am[c & 1] = fmax3(am[c & 1], fabsf(__uint_as_float(s[2 * c])), fabsf(__uint_as_float(s[2 * c + 1])));
          const float a = fmaxf(am[0], am[1]);
          poly = __all_sync(0xffffffffu, (fast ? a : fmaf(a, sce, fabsf(m_run))) <= 126.0f);
        }
        // ---- exp2, partial sums, bf16 pack
        uint32_t pk[16];
        uint64_t sums2[2] = {0ull, 0ull};
        const uint64_t mneg2 = pack_f32x2(-m_run, -m_run);
        if (fast) {
          if (poly) softmax_exp32<true, true, true, PolyD64>(s, sc2, mneg2, sums2, pk);
          else softmax_exp32<true, false, true, PolyD64>(s, sc2, mneg2, sums2, pk);
        } else {
          if (poly) softmax_exp32<false, true, true, PolyD64>(s, sc2, mneg2, sums2, pk);
          else softmax_exp32<false, false, true, PolyD64>(s, sc2, mneg2, sums2, pk);
        }
        {
          float sa, sb, sc_, sd;
          unpack_f32x2(sums2[0], sa, sb);
          unpack_f32x2(sums2[1], sc_, sd);
          l_half += (sa + sb) + (sc_ + sd);
        }
        if (!(l_half < ATT_FAIL_AT)) mbar_timeout(0x6F);
        // ---- P V of tile j-1 retired: P_x free, O_x quiescent; the partner's growth flag of tile j-1 is visible
        if (j > 0) {
          mbar_wait(&p_free[x], (j - 1) & 1, 0x61);
          tc_fence_after();
          pending = g_prev || flags[((x * 2 + ((j - 1) & 1)) * 2 + (1 - h)) * ATT_BM + r] != 0;
        }
        // my partial sum, as it will stand after a shift already scheduled for the next tile, passed 2^64 (or is not finite)
        const bool g = !(l_half * (pending ? 5.421010862427522e-20f : 1.0f) <= ATT_SHIFT_AT);
        if (rescale) {                                                    // rare: this half's 32 columns of O_x *= alpha
          uint32_t o0[32];
          tmem_ld32(tOx, o0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * alpha);
          tmem_st32(tOx, o0);
        }
        tmem_st16(tPx, pk);                                               // P_x(j), keys [32 h, 32 h + 32): 16 columns
        flags[((x * 2 + buf) * 2 + h) * ATT_BM + r] = g ? 1 : 0;
        g_prev = g;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);
      }
      // ---- final: row sum = both halves' partial sums (same reference); this half normalises and stores 32 of the 64 output columns
      mbar_wait(&o_full[x], 0, 0x69);
      tc_fence_after();
      *xmine = l_half;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + x) : "memory");
      const float l_row = l_half + *xpeer;
      uint32_t o0[32];
      tmem_ld32(tOx, o0);
      tmem_ld_wait();
      const int q = q0 + x * ATT_BM + r;
      if (q < S) {
        if (!(l_row < ATT_FAIL_AT) || !(l_row > 0.f)) mbar_timeout(0x6F);
        const float inv = 1.0f / l_row;
        bf16* dst = p.out + static_cast<long>(b) * p.out_batch_stride + static_cast<long>(q) * p.inner_dim + head * ATT_D + 32 * h;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(o0[c * 8 + 0]) * inv, __uint_as_float(o0[c * 8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(o0[c * 8 + 2]) * inv, __uint_as_float(o0[c * 8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(o0[c * 8 + 4]) * inv, __uint_as_float(o0[c * 8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(o0[c * 8 + 6]) * inv, __uint_as_float(o0[c * 8 + 7]) * inv);
          reinterpret_cast<uint4*>(dst)[c] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTma + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// FFB200_ATT_VARIANT selects the head_dim-64 kernel (A/B measurements; read once): "row3" = three sub-tiles, one warp per 32 rows,
// single-buffered S (attention_kernel); "row2" = two sub-tiles, one warp per 32 rows, double-buffered S; "split2" = two sub-tiles, two warps
// per row, double-buffered S.
static int attention_variant() {
  static const int v = [] {
    const char* e = getenv("FFB200_ATT_VARIANT");
    if (e == nullptr) return 0;
    if (strcmp(e, "row2") == 0) return 1;
    if (strcmp(e, "split2") == 0) return 2;
    return 0;
  }();
  return v;
}

#endif
