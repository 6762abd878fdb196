// proj_out + unpatchify + CFG + Euler/SDE update + noise + fp16 round trip + Gaussian log-prob, fused.
//
// The last GEMM of the transformer (DF/models/transformers/transformer_sd3.py:327, N = p*p*C = 64) keeps its result in
// TMEM; its epilogue IS the denoise step: it combines the unconditional / conditional accumulators of the same token
// (sd3_5.py:431-433), scatters through the unpatchify mapping (transformer_sd3.py:334-340, "nhwpqc->nchpwq"), applies
// scheduler.step (FF/scheduler/flow_match_euler_discrete.py:309-420), draws the noise (Philox) or reads it, writes the next
// latents (fp16, in place + the trajectory slot) and reduces the log-prob - no separate elementwise pass, and the noise
// prediction never goes to HBM.  The last CTA to finish reduces the per-CTA partials in a fixed order (deterministic) and
// advances the device-side step counter.
//
// One CTA per (128 tokens, sample): warp 0 TMA (A_uncond, A_cond, W tiles), warp 1 MMA (two M128 x N64 accumulators),
// warp 2 TMEM alloc, warps 4-7 epilogue (thread = token).
#include "common.cuh"
#include "kernels.h"

namespace ffb {

constexpr int FS_STAGES = 4;
constexpr int FS_A_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int FS_W_BYTES = 64 * 64 * 2;    // 8 KB
constexpr int FS_STAGE_BYTES = 2 * FS_A_BYTES + FS_W_BYTES;
constexpr int FS_SMEM = FS_STAGES * FS_STAGE_BYTES + 1024 + 256;
constexpr int FS_THREADS = 256;

__global__ void __launch_bounds__(FS_THREADS, 1)
final_step_kernel(const __grid_constant__ FinalStepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FS_STAGES * FS_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + FS_STAGES;
  uint64_t* acc_full = bars + 2 * FS_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * FS_STAGES + 1);
  __shared__ float red[4];
  __shared__ int is_last;

  const SdeStepParams& s = p.sde;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, b = blockIdx.y;
  const int n_acc = s.cfg ? 2 : 1;
  const int num_kb = (p.K + 63) / 64;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA); tma_prefetch_desc(&p.tmW); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < FS_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 0x70);
        uint8_t* st = smem + stage * FS_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], n_acc * FS_A_BYTES + FS_W_BYTES);
        tma_load_3d(st, &p.tmA, &full_bar[stage], kb * 64, tile * 128, b);                          // uncond (or only) half
        if (n_acc == 2) tma_load_3d(st + FS_A_BYTES, &p.tmA, &full_bar[stage], kb * 64, tile * 128, s.B + b);  // cond half
        tma_load_2d(st + 2 * FS_A_BYTES, &p.tmW, &full_bar[stage], kb * 64, 0);
        if (++stage == FS_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase, 0x71);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + stage * FS_STAGE_BYTES);
        const uint32_t w0 = a0 + 2 * FS_A_BYTES;
        for (int h = 0; h < n_acc; ++h) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + h * 64, desc_kmajor_sw128(a0 + h * FS_A_BYTES + k * 32), desc_kmajor_sw128(w0 + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(acc_full);
        if (++stage == FS_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = token =====================
    const int sidx = s.step_ptr ? *s.step_ptr : s.coef_index;
    const StepCoef k = s.coef_table[sidx];
    const int ew = warp - 4;
    const int tok = tile * 128 + ew * 32 + lane;
    const int wp = s.W / s.patch, hp = s.H / s.patch;
    const bool tok_ok = tok < hp * wp;
    const int ti = tok / wp, tj = tok % wp;
    const int CHW = s.C * s.H * s.W;
    const float* noise = s.noise ? s.noise + static_cast<long>(sidx) * s.noise_step_stride + static_cast<long>(b) * CHW : nullptr;
    const int st = s.storage;
    const long xoff = static_cast<long>(b) * CHW;                                     // element offsets into the storage-dtype buffers
    const bool has_traj = s.traj != nullptr && k.store_slot >= 0;
    const long toff = static_cast<long>(b) * s.traj_batch_stride + static_cast<long>(k.store_slot) * CHW;
    float* mean_out = s.mean_out ? s.mean_out + static_cast<long>(b) * CHW : nullptr;
    bf16* v_out = s.v_out ? s.v_out + static_cast<long>(b) * CHW : nullptr;

    mbar_wait(acc_full, 0, 0x72);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    float part = 0.f;
    // n = (py*2 + px)*C + c with C = 16: a 32-column chunk is one py (both px, all c)
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      uint32_t ru[32], rc[32];
      tmem_ld32(t_row + py * 32, ru);
      if (n_acc == 2) tmem_ld32(t_row + 64 + py * 32, rc);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 bu = __ldg(reinterpret_cast<const uint4*>(p.bias + py * 32 + q * 8));
        const float bf[8] = {bf16_lo(bu.x), bf16_hi(bu.x), bf16_lo(bu.y), bf16_hi(bu.y), bf16_lo(bu.z), bf16_hi(bu.z), bf16_lo(bu.w), bf16_hi(bu.w)};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int n = q * 8 + e;
          const float vu = bf16_round(__uint_as_float(ru[n]) + bf[e]);
          v[n] = vu;
          if (n_acc == 2) v[n] = cfg_combine_bf16(vu, bf16_round(__uint_as_float(rc[n]) + bf[e]), s.guidance);
        }
      }
      if (tok_ok) {
        const int y = ti * 2 + py;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int e0 = (c * s.H + y) * s.W + tj * 2;          // element index of the px = 0 pixel
          const float vv[2] = {v[c], v[16 + c]};                // px = 0, 1
          float xs[2];
          if (st == LAT_F16) {
            const __half2 xh = *reinterpret_cast<const __half2*>(static_cast<const __half*>(s.x) + xoff + e0);
            xs[0] = __low2float(xh); xs[1] = __high2float(xh);
          } else {
            xs[0] = lat_load(s.x, xoff + e0, st); xs[1] = lat_load(s.x, xoff + e0 + 1, st);
          }
          float mean[2], nxt[2];
          mean[0] = sde_mean(k, xs[0], vv[0]);
          mean[1] = sde_mean(k, xs[1], vv[1]);
          if (s.next_given != nullptr) {
            nxt[0] = lat_load(s.next_given, xoff + e0, st); nxt[1] = lat_load(s.next_given, xoff + e0 + 1, st);
          } else if (k.dynamics == DYN_ODE) {
            nxt[0] = mean[0]; nxt[1] = mean[1];
          } else {
            float z[2];
            if (noise != nullptr) {
              const float2 nz = *reinterpret_cast<const float2*>(noise + e0);
              z[0] = nz.x; z[1] = nz.y;
            } else {   // identical stream to sde_step_kernel: counter = (element quad, sample, step)
              uint32_t r[4];
              philox4x32_10(static_cast<uint32_t>(e0 >> 2), static_cast<uint32_t>(b), static_cast<uint32_t>(sidx), 0x5DEu,
                            static_cast<uint32_t>(s.seed), static_cast<uint32_t>(s.seed >> 32), r);
              if (e0 & 2) box_muller(r[2], r[3], &z[0], &z[1]);
              else box_muller(r[0], r[1], &z[0], &z[1]);
            }
            nxt[0] = sde_sample(k, mean[0], z[0], st);
            nxt[1] = sde_sample(k, mean[1], z[1], st);
          }
          if (s.x_next != nullptr) { lat_store(s.x_next, xoff + e0, st, nxt[0], s.overflow_flag); lat_store(s.x_next, xoff + e0 + 1, st, nxt[1], s.overflow_flag); }
          if (has_traj) { lat_store(s.traj, toff + e0, st, nxt[0], s.overflow_flag); lat_store(s.traj, toff + e0 + 1, st, nxt[1], s.overflow_flag); }
          if (mean_out) *reinterpret_cast<float2*>(mean_out + e0) = make_float2(mean[0], mean[1]);
          if (v_out) *reinterpret_cast<uint32_t*>(v_out + e0) = pack_bf16x2(vv[0], vv[1]);
          if (k.compute_log_prob && k.dynamics != DYN_ODE) part += sde_logp_term(k, nxt[0], mean[0]) + sde_logp_term(k, nxt[1], mean[1]);
        }
      }
    }
    part = warp_sum(part);
    if (lane == 0) red[ew] = part;
    // epilogue-only barrier (named barrier 1, 128 threads)
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (ew == 0 && lane == 0) {
      const int nblk = gridDim.x;
      s.logp_partial[b * nblk + tile] = red[0] + red[1] + red[2] + red[3];
      __threadfence();
      const unsigned int done = atomicAdd(p.done_counter, 1u);
      is_last = (done == gridDim.x * gridDim.y - 1);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (is_last) {
      // last CTA: fixed-order reduction per sample, normaliser, slot stores, step advance (sde_finalize_kernel's job)
      __threadfence();
      const int nblk = gridDim.x;
      const int t = threadIdx.x - 128;
      if (k.compute_log_prob) {
        for (int bb = t; bb < s.B; bb += 128) {
          float lp = 0.f;
          if (k.dynamics != DYN_ODE) {
            float acc = 0.f;
            for (int i = 0; i < nblk; ++i) acc += __ldcg(&s.logp_partial[bb * nblk + i]);
            lp = acc / static_cast<float>(CHW) - k.log_norm;
          }
          if (s.log_prob) s.log_prob[bb] = lp;
          if (s.logp_traj && k.logp_slot >= 0) s.logp_traj[bb * s.logp_batch_stride + k.logp_slot] = lp;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (t == 0) {
        *p.done_counter = 0;
        if (s.step_ptr) *s.step_ptr = sidx + 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

cudaError_t launch_final_step(const FinalStepParams& p, cudaStream_t stream) {
  static DeviceOnce once;                              // dynamic shared memory opt-in, once per device
  {
    cudaError_t e = once.run([] { return cudaFuncSetAttribute(final_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM); });
    if (e != cudaSuccess) return e;
  }
  const int ntok = (p.sde.H / p.sde.patch) * (p.sde.W / p.sde.patch);
  dim3 grid((ntok + 127) / 128, p.sde.B);
  final_step_kernel<<<grid, FS_THREADS, FS_SMEM, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ffb
