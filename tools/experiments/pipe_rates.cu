// Micro-benchmark: issue rate (cycles per warp-instruction per SM sub-partition) of the SIMT instructions the attention
// softmax is built from, on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, float seed) {
  float a[CHAINS], b[CHAINS];
  uint64_t pa[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i;
    b[i] = seed * 0.5f + i * 0.25f;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pa[i]) : "f"(a[i]), "f"(b[i]));
  }
  float c1 = seed * 1.0001f, c2 = seed * 0.9999f;
  uint64_t pc1, pc2;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc1) : "f"(c1), "f"(c2));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc2) : "f"(c2), "f"(c1));
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));                 // 3-reg FFMA
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a[i]) : "f"(c1));                   // FFMA reg,reg,imm
      if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(pc1), "l"(pc2));             // FFMA2
      if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pc1));                           // FADD2
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(b[i]));                   // FMNMX3
      if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));                                  // FMNMX
      if (OP == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                                     // MUFU.EX2
      if (OP == 7) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(b[i])); a[i] = __uint_as_float(r); }  // F2FP
      if (OP == 8) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));                               // FADD
      if (OP == 9) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pc1));                           // FMUL2
      if (OP == 10) { uint32_t r = __float_as_uint(a[i]); asm volatile("{.reg .b32 t; shl.b32 t, %1, 23; add.s32 %0, %0, t;}" : "+r"(r) : "r"(__float_as_uint(b[i]))); a[i] = __uint_as_float(r); }
      if (OP == 11) {   // MUFU + 3 FFMA-imm interleaved: do the pipes overlap?
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(b[i]) : "f"(c1));
        asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(b[i]) : "f"(c2));
      }
      if (OP == 12) {   // FFMA2 + FMNMX3 interleaved (fma pipe + alu pipe)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(pc1), "l"(pc2));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(b[i]));
      }
      if (OP == 13) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(a[(i + 1) % CHAINS]));  // 3 distinct varying regs
      if (OP == 14) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(pa[(i + 3) % CHAINS]), "l"(pa[(i + 5) % CHAINS]));
      // round 2: 16-bit packed transcendental forms (does one MUFU slot deliver two results?) and the conversions around them
      if (OP == 15) { uint32_t r = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
      if (OP == 16) { uint32_t r = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
      if (OP == 17) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 18) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(b[i])); a[i] = __uint_as_float(r); }   // F2FP.F16
      if (OP == 19) { uint32_t r = __float_as_uint(a[i]); asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r) : "r"(__float_as_uint(c1))); a[i] = __uint_as_float(r); }  // HADD2
      if (OP == 20) { uint32_t r = __float_as_uint(a[i]); asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(b[i]) : "r"(r)); a[i] += b[i]; }  // f16 -> f32 + FADD
      if (OP == 21) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                                      // MUFU.RCP
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pa[i]));
    s += a[i] + b[i] + lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, int warps_per_smsp) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int threads = warps_per_smsp * 4 * 32;
  k<OP><<<148, threads>>>(out, cyc, 1.0f);
  k<OP><<<148, threads>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double insts_per_warp = double(ITERS) * CHAINS * per_iter;
  printf("{\"op\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_warp_inst_per_smsp\": %.3f}\n", name, warps_per_smsp,
         avg / (insts_per_warp * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<0>("ffma_rrr_const", 1, w);
    run<13>("ffma_rrr_varying", 1, w);
    run<1>("ffma_rr_imm", 1, w);
    run<2>("ffma2_const", 1, w);
    run<14>("ffma2_varying", 1, w);
    run<3>("fadd2", 1, w);
    run<9>("fmul2", 1, w);
    run<8>("fadd", 1, w);
    run<4>("fmnmx3", 1, w);
    run<5>("fmnmx", 1, w);
    run<6>("mufu_ex2", 1, w);
    run<7>("f2fp_bf16x2", 1, w);
    run<10>("shl_add", 2, w);
    run<11>("mufu+2ffma_imm (per inst, 3 per group)", 3, w);
    run<12>("ffma2+fmnmx3 (per inst, 2 per group)", 2, w);
    run<15>("mufu_ex2_f16x2", 1, w);
    run<16>("mufu_ex2_bf16x2", 1, w);
    run<17>("mufu_tanh_f32", 1, w);
    run<18>("f2fp_f16x2", 1, w);
    run<19>("hadd2_f16x2", 1, w);
    run<20>("cvt_f32_f16+fadd (per inst, 2 per group)", 2, w);
    run<21>("mufu_rcp", 1, w);
  }
  return 0;
}
