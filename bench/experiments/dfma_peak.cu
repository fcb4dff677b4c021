// Micro-benchmark: FP64 DFMA issue rate on B200 and whether it overlaps the integer multiplier
// (IMAD.WIDE.U32) - input for the "another multiplier pipe" question in DESIGN.md (known gaps).
// Prints one JSON object.  Not part of the product.
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITERS = 4096;

template <int MODE>  // 0: DFMA only, 1: IMAD.WIDE only, 2: both interleaved 1:1, 3: DFMA + IADD3 (ALU) 1:2
__global__ void __launch_bounds__(256) k_mix(double* outd, unsigned long long* outi, double a0, unsigned b0) {
  double a = a0 + threadIdx.x, d[8];
  unsigned long long acc[8];
  unsigned m[8];
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d[i] = 1.0 + i;
    acc[i] = i;
    m[i] = b0 + i + threadIdx.x;
    x[i] = i;
  }
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || MODE == 2 || MODE == 3) d[i] = fma(d[i], a, 1.0);
      if (MODE == 1 || MODE == 2) acc[i] += (unsigned long long)m[i] * b0;
      if (MODE == 3) {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(m[i]));
        asm volatile("addc.u32 %0, %0, %1;" : "+r"(m[i]) : "r"(x[i]));
      }
    }
  }
  double s = 0;
  unsigned long long t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    s += d[i];
    t += acc[i] + x[i] + m[i];
  }
  outd[blockIdx.x * blockDim.x + threadIdx.x] = s;
  outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
static double run(int blocks_per_sm, int sms, double* od, unsigned long long* oi) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int grid = sms * blocks_per_sm;
  k_mix<MODE><<<grid, 256>>>(od, oi, 1.000001, 12345u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_mix<MODE><<<grid, 256>>>(od, oi, 1.000001, 12345u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double)grid * 256 * ITERS * 8 / (ms * 1e-3);  // loop-body "slots" per second (per op kind)
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* od;
  unsigned long long* oi;
  cudaMalloc(&od, (size_t)sms * 8 * 256 * 8);
  cudaMalloc(&oi, (size_t)sms * 8 * 256 * 8);
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  for (int bps : {2, 4, 8}) {
    printf(", \"blocks_per_sm_%d\": {\"dfma_per_s\": %.4g, \"imad_wide_per_s\": %.4g, \"both_pairs_per_s\": %.4g, \"dfma_plus_2alu_per_s\": %.4g}", bps,
           run<0>(bps, sms, od, oi), run<1>(bps, sms, od, oi), run<2>(bps, sms, od, oi), run<3>(bps, sms, od, oi));
  }
  printf("}\n");
  return 0;
}
