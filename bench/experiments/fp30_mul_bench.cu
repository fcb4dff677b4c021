// Micro-benchmark: dependent chains of full field multiplications per thread, no shared memory,
// swept over resident warps per SM:  fp30 (13x30-bit carry-free, fp30.cuh)  vs  fp32 (12x32-bit
// carry chains, field.cuh), and fp30 with two independent chains interleaved (ILP 2).
#include <cuda_runtime.h>
#include <cstdio>
#include "fp30.cuh"
using namespace dkgv;
template <int MODE>
__global__ void __launch_bounds__(128) k(uint32_t* out, uint32_t seed, int iters) {
  uint32_t r = 0;
  if (MODE == 0) {
    Fp a, b;
#pragma unroll
    for (int i = 0; i < 12; i++) { a.l[i] = seed + threadIdx.x * 12 + i; b.l[i] = seed * 7 + blockIdx.x + i; }
    a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
#pragma unroll 1
    for (int it = 0; it < iters; it++) { a = mul(a, b); b = mul(b, a); }
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= a.l[i] ^ b.l[i];
  } else {
    Fp30 a, b, c, d;
#pragma unroll
    for (int i = 0; i < 13; i++) { a.l[i] = (seed + threadIdx.x * 13 + i) & M30; b.l[i] = (seed * 7 + blockIdx.x + i) & M30; c.l[i] = (a.l[i] * 3) & M30; d.l[i] = (b.l[i] * 5) & M30; }
    a.l[12] &= 0xfffff; b.l[12] &= 0xfffff; c.l[12] &= 0xfffff; d.l[12] &= 0xfffff;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
      if (MODE == 1) { a = fp30_mul(a, b); b = fp30_mul(b, a); }
      if (MODE == 2) { a = fp30_mul(a, b); c = fp30_mul(c, d); b = fp30_mul(b, a); d = fp30_mul(d, c); }
    }
#pragma unroll
    for (int i = 0; i < 13; i++) r ^= a.l[i] ^ b.l[i] ^ c.l[i] ^ d.l[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* nm, int wps, uint32_t* d) {
  int grid = 148 * wps / 4, iters = 1024;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 128>>>(d, 1, 8); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<MODE><<<grid, 128>>>(d, 2 + r, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  double muls = (MODE == 2 ? 4.0 : 2.0) * iters * grid * 128.0;
  printf("\"%s_%dwarps_per_sm\": %.4e,\n", nm, wps, muls / (best * 1e-3));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 16 * 128 * 4);
  printf("{\n");
  for (int wps = 4; wps <= 32; wps *= 2) { run<0>("fp32_modmul_per_s", wps, d); run<1>("fp30_modmul_per_s", wps, d); run<2>("fp30_ilp2_modmul_per_s", wps, d); }
  printf("\"status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
