// Montgomery product of BLS12-381 Fp on the FP64 pipe (DFMA), bit-identical to mul() of field.cuh.
//
// Why: the 12 x 32-bit product of field.cuh is 300 carry-chained IMAD.WIDE.U32.X, which issue at half
// rate on B200 - the multiplier pipe is the bound of every hot kernel.  The FP64
// pipe issues DFMA at the full 16 lanes/clk/SMSP (bench/dfma_peak.cu) and sits idle.  A warp that runs
// THIS product leaves the integer multiplier to the warps that run the other one, so an SM whose
// resident warps are split between the two implementations uses both pipes at once.  Same Montgomery
// radix R = 2^384 and fully reduced results, so the two are interchangeable call by call.
//
// Scheme (8 limbs of 48 bits held as exact doubles; cf. Emmart-Zheng-Weems, ARITH 2018, for the split):
//   hi chain   H' = fma_rz(a, b, H)          H in [2^100, 2^101) has ulp 2^48, so the truncated sum keeps
//                                            adding floor(ab / 2^48): H = 2^100 + 2^48 * sum(hi parts)
//   lo part    L  = fma(a, b, H - H')        = ab mod 2^48 exactly (H - H' = -2^48 floor(ab / 2^48))
//   lo sum     LC += L                       at most 16 terms < 2^48: exact in 53 bits
// i.e. 4 FP64 instructions per 48x48-bit limb product and no integer work for the accumulation.
// Reduction: 8 rounds q = (column * -p^-1) mod 2^48 (integer, 3 IMAD), then q * p through the same chains.
#pragma once
#include "../../dvt_circuits_b200/csrc/field.cuh"

namespace dkgv {

#if defined(__CUDACC__)
namespace dfma {
constexpr double TWO52 = 4503599627370496.0;            // 2^52
constexpr double TWO100 = 1267650600228229401496703205376.0;  // 2^100
constexpr unsigned long long M48 = 0xFFFFFFFFFFFFull, M52 = 0xFFFFFFFFFFFFFull;
constexpr unsigned long long PINV48 = 0xfffcfffcfffdull;  // -p^-1 mod 2^48
__device__ __forceinline__ constexpr unsigned long long p48(int i) {
  constexpr unsigned long long v[8] = {0xffffffffaaabull, 0xb153ffffb9feull, 0xf6241eabfffeull, 0x6730d2a0f6b0ull,
                                       0x4b84f38512bfull, 0x434bacd76477ull, 0xe69a4b1ba7b6ull, 0x1a0111ea397full};
  return v[i];
}
// exact integer < 2^52 -> double
__device__ __forceinline__ double to_double(unsigned long long x) { return __longlong_as_double((long long)(0x4330000000000000ull | x)) - TWO52; }
// double holding an exact integer in [0, 2^52) -> integer
__device__ __forceinline__ unsigned long long to_int(double d) { return (unsigned long long)__double_as_longlong(d + TWO52) & M52; }
// 12 x 32-bit limbs -> 8 x 48-bit limbs
__device__ __forceinline__ void split48(const uint32_t* l, double* d) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    unsigned long long lo = (unsigned long long)l[3 * k] | ((unsigned long long)(l[3 * k + 1] & 0xffffu) << 32);
    unsigned long long hi = (unsigned long long)(l[3 * k + 1] >> 16) | ((unsigned long long)l[3 * k + 2] << 16);
    d[2 * k] = to_double(lo);
    d[2 * k + 1] = to_double(hi);
  }
}
}  // namespace dfma

// a * b * 2^-384 mod p, fully reduced; a, b < p in Montgomery form (any values < 2^384 with a*b < p*2^384 work)
__device__ __forceinline__ Fp mul_dfma(const Fp& a, const Fp& b) {
  using namespace dfma;
  double A[8], B[8];
  split48(a.l, A);
  split48(b.l, B);
  // column k (weight 2^(48k)): LC[k] = sum of lo parts, HC[k] = 2^100 + 2^48 * sum of hi parts of column k-1's products
  double LC[16], HC[17];
#pragma unroll
  for (int k = 0; k < 16; k++) LC[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 17; k++) HC[k] = TWO100;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      double h = __fma_rz(A[i], B[j], HC[i + j + 1]);
      double l = __fma_rn(A[i], B[j], HC[i + j + 1] - h);
      HC[i + j + 1] = h;
      LC[i + j] += l;
    }
  // Montgomery reduction, radix 2^48
  unsigned long long carry = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    // column i is complete: value = LC[i] + hi-sum(HC[i]) + carry
    unsigned long long v = to_int(LC[i]) + ((unsigned long long)__double_as_longlong(HC[i]) & M52) + carry;
    unsigned long long q = ((v & M48) * PINV48) & M48;
    double qd = to_double(q);
    // j = 0: v + lo(q p_0) is a multiple of 2^48
    {
      double pj = (double)p48(0);
      double h = __fma_rz(qd, pj, HC[i + 1]);
      double l = __fma_rn(qd, pj, HC[i + 1] - h);
      HC[i + 1] = h;
      carry = (v + to_int(l)) >> 48;
    }
#pragma unroll
    for (int j = 1; j < 8; j++) {
      double pj = (double)p48(j);
      double h = __fma_rz(qd, pj, HC[i + j + 1]);
      double l = __fma_rn(qd, pj, HC[i + j + 1] - h);
      HC[i + j + 1] = h;
      LC[i + j] += l;
    }
  }
  // result digits: columns 8..15 (column 16 is zero because the result is < 2p < 2^384)
  unsigned long long dgt[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    unsigned long long v = to_int(LC[8 + k]) + ((unsigned long long)__double_as_longlong(HC[8 + k]) & M52) + carry;
    dgt[k] = v & M48;
    carry = v >> 48;
  }
  Fp r;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    unsigned long long lo = dgt[2 * k], hi = dgt[2 * k + 1];
    r.l[3 * k] = (uint32_t)lo;
    r.l[3 * k + 1] = (uint32_t)(lo >> 32) | (uint32_t)(hi << 16);
    r.l[3 * k + 2] = (uint32_t)(hi >> 16);
  }
  cond_sub_mod<FpParams>(r.l, 0);
  return r;
}
#endif

}  // namespace dkgv
