// Micro-benchmark + bit-exactness check of the FP64-pipe Montgomery product (csrc/field_dfma.cuh) against
// the IMAD carry-chain product (csrc/field.cuh), alone and with the SM's warps split between the two.
#include <cuda_runtime.h>
#include <cstdio>
#include "experiments/field_dfma.cuh"
using namespace dkgv;

__device__ __forceinline__ uint32_t xs(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// random (and edge) operands < p: mismatches between the two products
__global__ void k_check(unsigned long long* bad, int iters) {
  uint32_t s = 0x9e3779b9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long nb = 0;
  for (int it = 0; it < iters; it++) {
    Fp a, b;
    for (int i = 0; i < 12; i++) { a.l[i] = xs(s); b.l[i] = xs(s); }
    a.l[11] &= 0x0fffffffu; b.l[11] &= 0x0fffffffu;  // < 2^380 < p
    int e = (it + threadIdx.x) & 15;
    if (e == 0) a = zero<FpParams>();
    if (e == 1) for (int i = 0; i < 12; i++) a.l[i] = FpParams::mod(i) - (i == 0);  // p - 1
    if (e == 2) { for (int i = 0; i < 12; i++) b.l[i] = FpParams::mod(i) - (i == 0); a = b; }
    if (e == 3) a = one<FpParams>();
    if (e == 4) for (int i = 0; i < 12; i++) a.l[i] = (i == 11) ? 0x1a0111eau : 0xffffffffu & FpParams::mod(i);  // = p: not < p but still a*b < pR
    if (e == 5) { for (int i = 0; i < 12; i++) a.l[i] = 0xffffu << 16; a.l[11] = 0x0fff0000u; }
    Fp r0 = mul(a, b), r1 = mul_dfma(a, b);
    Fp c0 = mul(r0, r0), c1 = mul_dfma(r1, r1);
    if (!eq(r0, r1) || !eq(c0, c1)) nb++;
  }
  if (nb) atomicAdd(bad, nb);
}

// mul2add(a,b,c,d) against add(mul(a,b), mul(c,d)) on random and extreme operands
__global__ void k_check2(unsigned long long* bad, int iters) {
  uint32_t s = 0x85ebca6bu * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long nb = 0;
  Fp pm1;
  for (int i = 0; i < 12; i++) pm1.l[i] = FpParams::mod(i) - (i == 0);
  for (int it = 0; it < iters; it++) {
    Fp v[4];
    for (int q = 0; q < 4; q++) {
      for (int i = 0; i < 12; i++) v[q].l[i] = xs(s);
      v[q].l[11] &= 0x0fffffffu;
      uint32_t e = xs(s) & 15;
      if (e == 0) v[q] = pm1;                 // p - 1
      if (e == 1) v[q] = zero<FpParams>();
      if (e == 2) { for (int i = 0; i < 11; i++) v[q].l[i] = 0xffffffffu; v[q].l[11] = 0x1a0111e9u; }  // largest value with all-ones low limbs < p
      if (e == 3) v[q] = one<FpParams>();
    }
    if ((it & 63) == 0) { v[0] = pm1; v[1] = pm1; v[2] = pm1; v[3] = pm1; }
    Fp r0 = add(mul(v[0], v[1]), mul(v[2], v[3])), r1 = mul2add(v[0], v[1], v[2], v[3]);
    if (!eq(r0, r1)) nb++;
    Fp r2 = sub(mul(v[0], v[1]), mul(v[2], v[3])), r3 = mul2add(v[0], v[1], v[2], neg(v[3]));
    if (!eq(r2, r3)) nb++;
  }
  if (nb) atomicAdd(bad, nb);
}
// speed: mode 0: two products + add, mode 1: mul2add
template <int MODE>
__global__ void __launch_bounds__(128) k_speed2(uint32_t* out, uint32_t seed, int iters) {
  Fp a, b, c, d;
#pragma unroll
  for (int i = 0; i < 12; i++) { a.l[i] = seed + threadIdx.x * 12 + i; b.l[i] = seed * 7 + blockIdx.x + i; c.l[i] = a.l[i] * 3; d.l[i] = b.l[i] * 5; }
  a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff; c.l[11] &= 0x0fffffff; d.l[11] &= 0x0fffffff;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) { a = add(mul(a, b), mul(c, d)); c = add(mul(c, a), mul(b, d)); }
    else { a = mul2add(a, b, c, d); c = mul2add(c, a, b, d); }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) r ^= a.l[i] ^ c.l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> static double run2(int wps, uint32_t* d) {
  int grid = 148 * wps / 4, iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_speed2<MODE><<<grid, 128>>>(d, 1, 8); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k_speed2<MODE><<<grid, 128>>>(d, 2 + r, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return 2.0 * iters * grid * 128.0 / (best * 1e-3);  // sum-of-two-products per second
}

// frac_dfma_of_16: blocks with (blockIdx.x & 15) < frac use the DFMA product
__global__ void __launch_bounds__(128) k_speed(uint32_t* out, uint32_t seed, int iters, int frac) {
  Fp a, b;
#pragma unroll
  for (int i = 0; i < 12; i++) { a.l[i] = seed + threadIdx.x * 12 + i; b.l[i] = seed * 7 + blockIdx.x + i; }
  a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
  if ((int)(blockIdx.x & 15) < frac) {
#pragma unroll 1
    for (int it = 0; it < iters; it++) { a = mul_dfma(a, b); b = mul_dfma(b, a); }
  } else {
#pragma unroll 1
    for (int it = 0; it < iters; it++) { a = mul(a, b); b = mul(b, a); }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) r ^= a.l[i] ^ b.l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// work-conserving version: blocks pull chunks of 32 iterations from a global counter until the total is done,
// so a mix of fast and slow blocks finishes together (as resident blocks of a real kernel do)
__global__ void __launch_bounds__(128) k_pull(uint32_t* out, uint32_t seed, unsigned* ctr, unsigned nchunks, int frac, unsigned* done_by_kind) {
  __shared__ unsigned chunk;
  Fp a, b;
#pragma unroll
  for (int i = 0; i < 12; i++) { a.l[i] = seed + threadIdx.x * 12 + i; b.l[i] = seed * 7 + blockIdx.x + i; }
  a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;
  bool use_dfma = (int)(blockIdx.x & 15) < frac;
  unsigned mine = 0;
  while (true) {
    if (threadIdx.x == 0) chunk = atomicAdd(ctr, 1u);
    __syncthreads();
    unsigned c = chunk;
    __syncthreads();
    if (c >= nchunks) break;
    mine++;
    if (use_dfma) {
#pragma unroll 1
      for (int it = 0; it < 32; it++) { a = mul_dfma(a, b); b = mul_dfma(b, a); }
    } else {
#pragma unroll 1
      for (int it = 0; it < 32; it++) { a = mul(a, b); b = mul(b, a); }
    }
  }
  if (threadIdx.x == 0) atomicAdd(&done_by_kind[use_dfma ? 1 : 0], mine);
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) r ^= a.l[i] ^ b.l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static double run_pull(int wps, int frac, uint32_t* d, unsigned* ctr, double* dfma_share) {
  int grid = 148 * wps / 4;
  unsigned nchunks = (unsigned)grid * 16;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  unsigned kinds[2] = {0, 0};
  for (int r = 0; r < 3; r++) {
    cudaMemset(ctr, 0, 12);
    cudaEventRecord(e0); k_pull<<<grid, 128>>>(d, 2 + r, ctr, nchunks, frac, ctr + 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) { best = ms; cudaMemcpy(kinds, ctr + 1, 8, cudaMemcpyDeviceToHost); }
  }
  *dfma_share = (double)kinds[1] / (kinds[0] + kinds[1]);
  return 2.0 * 32 * (double)nchunks * 128.0 / (best * 1e-3);
}

static double run(int wps, int frac, uint32_t* d) {
  int grid = 148 * wps / 4, iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_speed<<<grid, 128>>>(d, 1, 8, frac); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k_speed<<<grid, 128>>>(d, 2 + r, iters, frac); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return 2.0 * iters * grid * 128.0 / (best * 1e-3);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 16 * 128 * 4);
  unsigned long long* bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
  k_check<<<296, 128>>>(bad, 256);
  unsigned long long hb = ~0ull; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
  printf("{\"checked_pairs\": %d, \"mismatches\": %llu,\n", 296 * 128 * 256 * 2, hb);
  cudaMemset(bad, 0, 8);
  k_check2<<<296, 128>>>(bad, 256);
  cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
  printf("\"mul2add_checked\": %d, \"mul2add_mismatches\": %llu,\n", 296 * 128 * 256 * 2, hb);
  for (int wps : {8, 16, 32})
    printf("\"sum2prod_per_s_%dwarps_separate\": %.4e, \"sum2prod_per_s_%dwarps_mul2add\": %.4e,\n", wps, run2<0>(wps, d), wps, run2<1>(wps, d));
  for (int wps : {8, 16, 32})
    for (int frac : {0, 16, 4, 6, 8, 10})
      printf("\"modmul_per_s_%dwarps_dfma%dof16\": %.4e,\n", wps, frac, run(wps, frac, d));
  unsigned* ctr; cudaMalloc(&ctr, 12);
  for (int wps : {8, 16, 32})
    for (int frac : {0, 16, 4, 6, 8, 10, 12}) {
      double share;
      double v = run_pull(wps, frac, d, ctr, &share);
      printf("\"pull_modmul_per_s_%dwarps_dfma%dof16\": %.4e, \"pull_dfma_work_share_%dwarps_%dof16\": %.3f,\n", wps, frac, v, wps, frac, share);
    }
  printf("\"status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
