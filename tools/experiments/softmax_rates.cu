// Micro-benchmark of the attention softmax inner routine in isolation (no TMEM, no barriers, no tensor core): cycles per warp and KV tile
// (64 scores per thread) for the product instruction mix and for variants with one ingredient removed or replaced, at 1-4 warps per SM
// sub-partition.  Tells apart what the SIMT side costs by itself from what the synchronisation skeleton of the kernels adds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I flow_factory_b200/csrc -o gpurun_out/softmax_rates tools/experiments/softmax_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace ffb;

#define ITERS 256

__device__ __forceinline__ bool slot(int c, int num, bool cluster) { return cluster ? (c % 8) < num : ((c % 8) * num) % 8 < num; }

__device__ __forceinline__ void poly_packed(float x0, float x1, float& e0, float& e1, bool deg2) {
  const uint64_t xc = pack_f32x2(x0, x1);
  const uint64_t xr = fadd2(xc, pack_f32x2(12582912.0f, 12582912.0f));
  const uint64_t r = fadd2(xr, pack_f32x2(-12582912.0f, -12582912.0f));
  const uint64_t f = ffma2(r, pack_f32x2(-1.0f, -1.0f), xc);
  uint64_t p;
  if (deg2) {
    p = ffma2(f, pack_f32x2(0.2402265f, 0.2402265f), pack_f32x2(0.6931472f, 0.6931472f));
    p = ffma2(p, f, pack_f32x2(1.0017248f, 1.0017248f));
  } else {
    p = ffma2(f, pack_f32x2(0.05517132207751274f, 0.05517132207751274f), pack_f32x2(0.24261054396629333f, 0.24261054396629333f));
    p = ffma2(p, f, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
    p = ffma2(p, f, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
  }
  float p0, p1, r0, r1;
  unpack_f32x2(p, p0, p1);
  unpack_f32x2(xr, r0, r1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(r0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(r1) << 23));
}
__device__ __forceinline__ float poly_scalar(float x) {
  const float xr = x + 12582912.0f;
  const float r = xr - 12582912.0f;
  const float f = x - r;
  float p = fmaf(f, 0.05517132207751274f, 0.24261054396629333f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999281167984009f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(xr) << 23));
}

struct Cfg { int num; bool cluster, sum, pack, scalar_sum, scalar_poly, scale, deg2, sum4; };

template <int V> __device__ __forceinline__ constexpr Cfg cfg() {
  //            num clus  sum   pack  ssum  spoly scale deg2  sum4
  if (V == 0) return {2, false, true, true, false, false, false, false, false};    // product (pre-scaled keys, reference 0)
  if (V == 1) return {1, false, true, true, false, false, false, false, false};
  if (V == 2) return {0, false, true, true, false, false, false, false, false};    // all MUFU
  if (V == 3) return {3, false, true, true, false, false, false, false, false};
  if (V == 4) return {2, false, false, true, false, false, false, false, false};   // no row sum
  if (V == 5) return {2, false, true, false, false, false, false, false, false};   // no bf16 pack
  if (V == 6) return {2, false, false, false, false, false, false, false, false};  // neither
  if (V == 7) return {0, false, false, false, false, false, false, false, false};  // 64 MUFU only
  if (V == 8) return {2, false, true, true, true, false, false, false, false};     // scalar row sum
  if (V == 9) return {2, false, true, true, false, true, false, false, false};     // scalar polynomial
  if (V == 10) return {2, false, true, true, false, false, true, false, false};    // general path: + FFMA2 scale per pair
  if (V == 11) return {2, false, true, true, false, false, false, true, false};    // degree-2 polynomial
  if (V == 12) return {2, false, true, true, false, false, false, false, true};    // 4 packed sum accumulators
  if (V == 13) return {3, true, true, true, false, false, false, false, false};    // 3 of 8 clustered
  if (V == 14) return {0, false, true, false, false, false, false, false, false};  // all MUFU + sum, no pack
  if (V == 15) return {0, false, false, true, false, false, false, false, false};  // all MUFU + pack, no sum
  if (V == 16) return {4, false, true, true, false, false, false, false, false};
  if (V == 17) return {1, false, false, true, false, false, false, false, false};  // poly 1, no sum (tensor-core row sum what-if)
  return {2, false, true, true, false, false, false, false, false};
}

template <int V>
__global__ void __launch_bounds__(512, 1) k(const float* in, uint32_t* out, long long* cyc) {
  constexpr Cfg c = cfg<V>();
  // the scores of every "tile" are re-read from shared memory (volatile 128-bit loads, 16 per thread and tile - the stand-in for the two
  // TMEM loads of the kernels), so nothing can be hoisted out of the loop
  extern __shared__ uint4 sm_scores[];             // [16 score + 8 P][blockDim.x]
  for (int i = 0; i < 16; ++i) {
    uint4 v;
    v.x = __float_as_uint(in[(threadIdx.x * 64 + 4 * i + 0) & 4095]); v.y = __float_as_uint(in[(threadIdx.x * 64 + 4 * i + 1) & 4095]);
    v.z = __float_as_uint(in[(threadIdx.x * 64 + 4 * i + 2) & 4095]); v.w = __float_as_uint(in[(threadIdx.x * 64 + 4 * i + 3) & 4095]);
    sm_scores[i * blockDim.x + threadIdx.x] = v;
  }
  uint32_t s[64];
  uint32_t pk[32];
  float l = 0.f;
  const uint64_t sc2 = pack_f32x2(1.0009f, 1.0009f), mneg2 = pack_f32x2(-0.001f, -0.001f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint4 v;
      asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(&sm_scores[i * blockDim.x + threadIdx.x])));
      s[4 * i] = v.x; s[4 * i + 1] = v.y; s[4 * i + 2] = v.z; s[4 * i + 3] = v.w;
    }
    uint64_t sums2[4] = {0ull, 0ull, 0ull, 0ull};
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      float t0_ = __uint_as_float(s[2 * q]), t1_ = __uint_as_float(s[2 * q + 1]);
      if (c.scale) unpack_f32x2(ffma2(pack_f32x2(t0_, t1_), sc2, mneg2), t0_, t1_);
      float e0, e1;
      if (slot(q % 16, c.num, c.cluster)) {
        if (c.scalar_poly) { e0 = poly_scalar(t0_); e1 = poly_scalar(t1_); }
        else poly_packed(t0_, t1_, e0, e1, c.deg2);
      } else {
        e0 = ex2_approx(t0_); e1 = ex2_approx(t1_);
      }
      if (c.sum) {
        if (c.scalar_sum) { ss[q & 1] += e0; ss[2 + (q & 1)] += e1; }
        else if (c.sum4) sums2[q & 3] = fadd2(sums2[q & 3], pack_f32x2(e0, e1));
        else sums2[q & 1] = fadd2(sums2[q & 1], pack_f32x2(e0, e1));
      }
      if (c.pack) pk[q] = pack_bf16x2(e0, e1);
      else { pk[q] = __float_as_uint(e0) ^ __float_as_uint(e1); }
    }
    if (c.sum) {
      float a, b, cc, d;
      if (c.scalar_sum) l += (ss[0] + ss[1]) + (ss[2] + ss[3]);
      else {
        unpack_f32x2(fadd2(sums2[0], sums2[2]), a, b);
        unpack_f32x2(fadd2(sums2[1], sums2[3]), cc, d);
        l += (a + b) + (cc + d);
      }
    }
    // P "leaves": 8 volatile 128-bit stores into a separate region (stand-in for the TMEM store; in the no-pack variants the stored words
    // are the XOR of the two exponentials, one LOP3 instead of one F2FP per pair, so that the work stays alive)
#pragma unroll
    for (int q = 0; q < 8; ++q)
      asm volatile("st.volatile.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(&sm_scores[(16 + q) * blockDim.x + threadIdx.x])), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(l) ^ pk[threadIdx.x & 31];
}

template <int V>
void run(const char* name, const float* in, uint32_t* out, long long* cyc, int sms) {
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 384);
  for (int wps = 1; wps <= 4; ++wps) {
    k<V><<<sms, wps * 128, wps * 128 * 384>>>(in, out, cyc);
    cudaDeviceSynchronize();
    k<V><<<sms, wps * 128, wps * 128 * 384>>>(in, out, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"variant\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    long long h[256];
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per_round = static_cast<double>(mx) / ITERS;                  // cycles per round of `wps` warp-tiles per sub-partition
    printf("{\"variant\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_round\": %.1f, \"cycles_per_warp_tile\": %.1f, \"scores_per_clk_per_sm\": %.2f}\n",
           name, wps, per_round, per_round / wps, 4.0 * wps * 32 * 64 / per_round);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, sms * 512 * 4); cudaMalloc(&cyc, sms * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = -8.0f + 0.004f * i;                    // exponents in [-8, 8]
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("product: poly 2/8 spread, packed sum, pack", in, out, cyc, sms);
  run<1>("poly 1/8", in, out, cyc, sms);
  run<2>("poly 0 (all MUFU)", in, out, cyc, sms);
  run<3>("poly 3/8 spread", in, out, cyc, sms);
  run<13>("poly 3/8 clustered", in, out, cyc, sms);
  run<16>("poly 4/8", in, out, cyc, sms);
  run<4>("poly 2/8, NO row sum", in, out, cyc, sms);
  run<5>("poly 2/8, NO bf16 pack", in, out, cyc, sms);
  run<6>("poly 2/8, no sum, no pack", in, out, cyc, sms);
  run<7>("64 MUFU only", in, out, cyc, sms);
  run<14>("all MUFU + sum", in, out, cyc, sms);
  run<15>("all MUFU + pack", in, out, cyc, sms);
  run<17>("poly 1/8, NO row sum", in, out, cyc, sms);
  run<8>("poly 2/8, scalar row sum", in, out, cyc, sms);
  run<12>("poly 2/8, 4 packed sum accumulators", in, out, cyc, sms);
  run<9>("poly 2/8, scalar polynomial", in, out, cyc, sms);
  run<11>("poly 2/8, degree-2 polynomial", in, out, cyc, sms);
  run<10>("general path: + FFMA2 scale", in, out, cyc, sms);
  return 0;
}
