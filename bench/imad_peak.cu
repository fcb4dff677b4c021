// Integer-pipe peak micro-benchmark for the roofline denominator ("measure the
// achievable IMAD peak with a micro-benchmark first").  Independent accumulators, no memory
// traffic, all SMs busy.  Prints one JSON object: lane-ops per second and per clock per SM for
//   imad_lo   mad.lo.u32          (IMAD)
//   imad_hi   mad.hi.u32          (IMAD.HI)
//   imad_wide mad.wide.u32        (IMAD.WIDE.U32, 64-bit accumulate)
//   imad_wide_x  mad.lo.cc/madc.hi.cc carry chain (IMAD.WIDE.U32.X as used by the Montgomery product)
//   iadd3     add.cc/addc chain   (IADD3[.X], ALU pipe)
//   ffma      fma.rn.f32          (FFMA, for reference)
//   fp_mul    full 381-bit Montgomery products/s with the library's own fp mul (dependent chain/thread)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../dvt_circuits_b200/csrc/field.cuh"

#define ITERS 2048
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t lo[ILP], hi[ILP];
  float f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { lo[i] = a + i; hi[i] = b + i; f[i] = (float)i; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a), "r"(b));
        if (MODE == 1) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a), "r"(b));
        if (MODE == 2) {
          unsigned long long acc = ((unsigned long long)hi[i] << 32) | lo[i];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
          lo[i] = (uint32_t)acc; hi[i] = (uint32_t)(acc >> 32);
        }
        if (MODE == 5) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
      }
      if (MODE == 3) {  // one carry chain through ILP wide products (like one row of the Montgomery product)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a), "r"(b));
#pragma unroll
        for (int i = 1; i < ILP; i++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a), "r"(b));
      }
      if (MODE == 4) {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo[0]) : "r"(a));
#pragma unroll
        for (int i = 1; i < ILP; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(b));
      }
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r ^= lo[i] ^ hi[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void __launch_bounds__(256) k_fpmul(uint32_t* out, uint32_t seed, long long* cyc, int iters) {
  using namespace dkgv;
  Fp a, b;
#pragma unroll
  for (int i = 0; i < 12; i++) { a.l[i] = seed + threadIdx.x * 12 + i; b.l[i] = seed * 7 + blockIdx.x + i; }
  a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) { a = mul(a, b); b = mul(b, a); }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) r ^= a.l[i] ^ b.l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
static void run(const char* name, int sms, int blocks_per_sm, uint32_t* d_out, long long* d_cyc, bool last) {
  int grid = sms * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_pipe<MODE><<<grid, 256>>>(d_out, 1234, d_cyc);  // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_pipe<MODE><<<grid, 256>>>(d_out, 1234 + rep, d_cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  double ops_per_thread = (double)ITERS * 4 * ILP;  // wide products (MODE 3: each lo/hi pair = one wide op)
  double total = ops_per_thread * grid * 256.0;
  double per_clk_sm = ops_per_thread * blocks_per_sm * 256.0 / (double)cyc;
  printf("  \"%s\": {\"lane_ops_per_s\": %.4e, \"ms\": %.4f, \"block0_cycles\": %lld, \"lane_ops_per_clk_per_sm\": %.2f}%s\n", name,
         total / (best * 1e-3), best, cyc, per_clk_sm, last ? "" : ",");
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  uint32_t* d_out; long long* d_cyc;
  cudaMalloc(&d_out, (size_t)sms * 8 * 256 * 4); cudaMalloc(&d_cyc, 8);
  printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.clockRate);
  run<0>("imad_lo", sms, 8, d_out, d_cyc, false);
  run<1>("imad_hi", sms, 8, d_out, d_cyc, false);
  run<2>("imad_wide", sms, 8, d_out, d_cyc, false);
  run<3>("imad_wide_x", sms, 8, d_out, d_cyc, false);
  run<4>("iadd3_x", sms, 8, d_out, d_cyc, false);
  run<5>("ffma", sms, 8, d_out, d_cyc, false);
  // Montgomery products: sweep resident warps per SM (blocks of 256 threads)
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_fpmul<<<grid, 256>>>(d_out, 99, d_cyc, 64); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0); k_fpmul<<<grid, 256>>>(d_out, 99 + rep, d_cyc, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    double muls = 2.0 * iters * grid * 256.0;
    printf("  \"fp_mul_%dwarps_per_sm\": {\"modmul_per_s\": %.4e, \"ms\": %.4f, \"modmul_per_clk_per_sm\": %.4f},\n", bps * 8,
           muls / (best * 1e-3), best, 2.0 * iters * bps * 256.0 / (double)cyc);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("  \"cuda_status\": \"%s\"\n}\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}
