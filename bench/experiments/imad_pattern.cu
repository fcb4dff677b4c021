// Micro-benchmark: IMAD.WIDE.U32 throughput with REALISTIC operand patterns (distinct register
// multiplicands, 64-bit accumulators), to decide what the true integer-pipe ceiling of a big-number
// product is on B200.  Variants:
//   prod13   13x13 column-accumulated product (fp30.cuh's first phase), 169 IMAD.WIDE / iteration
//   prod13i  same but one multiplicand array is compile-time constants (immediates)
//   same2    every IMAD.WIDE uses the SAME two multiplicand registers (what bench/imad_peak measures)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a[13], b[13];
#pragma unroll
  for (int i = 0; i < 13; i++) { a[i] = (seed + threadIdx.x * 13 + i) & 0x3fffffff; b[i] = (seed * 7 + blockIdx.x + i) & 0x3fffffff; }
  uint64_t t[26];
#pragma unroll
  for (int k2 = 0; k2 < 26; k2++) t[k2] = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 13; i++) {
#pragma unroll
      for (int j = 0; j < 13; j++) {
        if (MODE == 0) t[i + j] += (uint64_t)a[i] * b[j];
        if (MODE == 1) t[i + j] += (uint64_t)a[i] * (uint32_t)(0x2affffacu + 0x01010101u * j);
        if (MODE == 2) t[i + j] += (uint64_t)a[0] * b[0];
      }
    }
    a[it % 13 == 0 ? 0 : 1] ^= (uint32_t)(t[12] >> 34) & 1;
  }
  uint32_t r = 0;
#pragma unroll
  for (int k2 = 0; k2 < 26; k2++) r ^= (uint32_t)t[k2] ^ (uint32_t)(t[k2] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* nm, int bps, uint32_t* d) {
  int grid = 148 * bps, iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(d, 1, 8); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, 2 + r, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  printf("\"%s_%dw\": %.4e,\n", nm, bps * 8, 169.0 * iters * grid * 256.0 / (best * 1e-3));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  printf("{\n");
  for (int bps = 1; bps <= 4; bps *= 2) { run<0>("prod13_lane_macs_per_s", bps, d); run<1>("prod13imm_lane_macs_per_s", bps, d); run<2>("same2_lane_macs_per_s", bps, d); }
  printf("\"status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
