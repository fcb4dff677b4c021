// Micro-benchmark: does IMAD.WIDE.U32 with carry-OUT only (no carry-in) issue at the full IMAD rate,
// and can the carry be absorbed by an IADD3.X on the ALU pipe in its shadow?
//   mode 0: mad.wide.u32 acc64 += a*b                      (no carries at all; reference)
//   mode 1: mad.lo.cc + madc.hi.cc (carry-out), addc cnt   (carry-save accumulation)
//   mode 2: mad.lo.cc + madc.hi.cc chained through 8 accumulators (.X chain; what field.cuh uses)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define ILP 8
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a[ILP], b[ILP], lo[ILP], hi[ILP], cnt[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { a[i] = seed + threadIdx.x + i; b[i] = seed * 3 + blockIdx.x + 7 * i; lo[i] = i; hi[i] = 2 * i; cnt[i] = 0; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (MODE == 2) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a[0]), "r"(b[u]));
#pragma unroll
        for (int i = 1; i < ILP; i++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(b[(i + u) % ILP]));
      } else {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
          if (MODE == 0) {
            unsigned long long acc = ((unsigned long long)hi[i] << 32) | lo[i];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[i]), "r"(b[(i + u) % ILP]));
            lo[i] = (uint32_t)acc; hi[i] = (uint32_t)(acc >> 32);
          }
          if (MODE == 1)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                         : "+r"(lo[i]), "+r"(hi[i]), "+r"(cnt[i]) : "r"(a[i]), "r"(b[(i + u) % ILP]));
        }
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r ^= lo[i] ^ hi[i] ^ cnt[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* nm, uint32_t* d) {
  int grid = 148 * 4, iters = 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(d, 1, 8); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<MODE><<<grid, 256>>>(d, 2 + r, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  printf("\"%s\": %.4e,\n", nm, 4.0 * ILP * iters * grid * 256.0 / (best * 1e-3));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  printf("{\n");
  run<0>("wide_nocarry_lane_macs_per_s", d); run<1>("wide_carryout_plus_addc_lane_macs_per_s", d); run<2>("wide_x_chain_lane_macs_per_s", d);
  printf("\"status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
