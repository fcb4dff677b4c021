// Fp arithmetic on 13 unsaturated 30-bit limbs - the field backend of the hot kernel.
//
// Why: bench/imad_peak (profiles/r1_imad_peak_first.json) shows that on B200 the carry-chained
// IMAD.WIDE.U32.X issues at HALF the rate of a plain IMAD.WIDE.U32 (0.90e13 vs 1.84e13 lane-ops/s),
// and the 12 x 32-bit Montgomery product of field.cuh is exactly bound by the former.  With 30-bit
// limbs a 13 x 13 product needs no carries at all: every partial product is < 2^60, a column sums at
// most 13 of them (< 2^64), so the whole product is 351 independent IMAD.WIDE.U32 with 64-bit
// accumulation plus shift/mask work that runs on the ALU pipe in the shadow of the multiplier.
//
// Representation: value = sum l[i] * 2^(30 i), Montgomery form with R = 2^390.  Values are LAZY:
//   * limbs are "weakly normalised": l[i] < 2^30 + 64 for i < 12 (top limb takes the rest),
//   * the value is only bounded by BND * p with BND < 600 (2^390 / p = 630), not reduced mod p.
// mul accepts any such operands and returns a value < (BNDa * BNDb / 630 + 1) * p with strictly
// normalised limbs; add/sub/mul_small just add limb-wise and do one parallel carry pass.
// sub adds a biased multiple K*p (every limb >= 2^30 + 64) so no limb ever goes negative; the caller
// picks K > bound of the subtrahend.  The point formulas in vm30.cuh are written so that all
// coordinate bounds reach a fixed point <= 16 (checked dynamically in host builds with
// -DDKGV_BOUND_CHECK, see tests/hostemu).
// Plain C on both host and device: nvcc maps `acc += (uint64_t)a * b` to IMAD.WIDE.U32.
#pragma once
#include "../../dvt_circuits_b200/csrc/field.cuh"

#if defined(DKGV_BOUND_CHECK) && !defined(__CUDA_ARCH__)
#include <cassert>
#define FP30_BD(x) x
#else
#define FP30_BD(x)
#endif

namespace dkgv {

constexpr uint32_t M30 = (1u << 30) - 1;

struct Fp30 {
  uint32_t l[13];
#if defined(DKGV_BOUND_CHECK) && !defined(__CUDA_ARCH__)
  double bd = 1.0;  // value < bd * p (static worst case, independent of the data)
#endif
};

#if defined(DKGV_BOUND_CHECK) && !defined(__CUDA_ARCH__)
inline void fp30_check(const Fp30& a) {
  for (int i = 0; i < 12; i++) assert(a.l[i] < (1u << 30) + 64);
  assert(a.l[12] < (1u << 30));
  assert(a.bd < 600.0);
}
#endif

DKGV_HD Fp30 fp30_mul(const Fp30& a, const Fp30& b) {
  FP30_BD(fp30_check(a); fp30_check(b);)
  uint64_t t[26];
#pragma unroll
  for (int k = 0; k < 26; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 13; i++) {
#pragma unroll
    for (int j = 0; j < 13; j++) t[i + j] += (uint64_t)a.l[i] * b.l[j];
  }
  // parallel carry pass: column k keeps its low 30 bits and receives bits 30..59 of column k-1 and
  // bits 60..63 of column k-2  ->  every column < 2^31 + 16, no serial dependency
  uint64_t acc[26];
#pragma unroll
  for (int k = 0; k < 26; k++) {
    uint32_t v = (uint32_t)t[k] & M30;
    if (k >= 1) v += (uint32_t)(t[k - 1] >> 30) & M30;
    if (k >= 2) v += (uint32_t)(t[k - 2] >> 60);
    acc[k] = v;
  }
  // Montgomery reduction, one 30-bit digit per step
  uint64_t carry = 0;
#pragma unroll
  for (int i = 0; i < 13; i++) {
    uint64_t c = acc[i] + carry;
    uint32_t m = ((uint32_t)c * consts::c30::NINV) & M30;
    c += (uint64_t)m * consts::c30::P(0);
    carry = c >> 30;
#pragma unroll
    for (int j = 1; j < 13; j++) acc[i + j] += (uint64_t)m * consts::c30::P(j);
  }
  Fp30 r;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    uint64_t c = acc[13 + k] + carry;
    r.l[k] = (uint32_t)c & M30;
    carry = c >> 30;
  }
  r.l[12] = (uint32_t)(acc[25] + carry);
  FP30_BD(r.bd = a.bd * b.bd / 630.0 + 1.0; fp30_check(r);)
  return r;
}

// one parallel carry pass over limb sums s[i] < 2^32 (top limb absorbs its carry-in)
DKGV_HD void fp30_carry(Fp30& r, const uint32_t* s) {
  r.l[0] = s[0] & M30;
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = (s[i] & M30) + (s[i - 1] >> 30);
  r.l[12] = s[12] + (s[11] >> 30);
}

DKGV_HD Fp30 fp30_add(const Fp30& a, const Fp30& b) {
  uint32_t s[13];
#pragma unroll
  for (int i = 0; i < 13; i++) s[i] = a.l[i] + b.l[i];
  Fp30 r;
  fp30_carry(r, s);
  FP30_BD(r.bd = a.bd + b.bd; fp30_check(r);)
  return r;
}

// a - b + K*p with K in {4, 8, 32, 64}; requires bound(b) < K
template <int K>
DKGV_HD uint32_t fp30_kp(int i) {
  return K == 4 ? consts::c30::KP4(i) : K == 8 ? consts::c30::KP8(i) : K == 32 ? consts::c30::KP32(i) : consts::c30::KP64(i);
}
template <int K>
DKGV_HD Fp30 fp30_sub(const Fp30& a, const Fp30& b) {
  FP30_BD(assert(b.bd < (double)K);)
  uint32_t s[13];
#pragma unroll
  for (int i = 0; i < 13; i++) s[i] = a.l[i] + fp30_kp<K>(i) - b.l[i];
  Fp30 r;
  fp30_carry(r, s);
  FP30_BD(r.bd = a.bd + (double)K; fp30_check(r);)
  return r;
}

// k * a for a small constant k <= 16
template <int KS>
DKGV_HD Fp30 fp30_mul_small(const Fp30& a) {
  uint64_t s[13];
#pragma unroll
  for (int i = 0; i < 13; i++) s[i] = (uint64_t)a.l[i] * (uint32_t)KS;
  Fp30 r;
  r.l[0] = (uint32_t)s[0] & M30;
#pragma unroll
  for (int i = 1; i < 12; i++) r.l[i] = ((uint32_t)s[i] & M30) + (uint32_t)(s[i - 1] >> 30);
  r.l[12] = (uint32_t)s[12] + (uint32_t)(s[11] >> 30);
  FP30_BD(r.bd = a.bd * KS; fp30_check(r);)
  return r;
}

DKGV_HD Fp30 fp30_const(uint32_t (*f)(int)) {
  Fp30 r;
#pragma unroll
  for (int i = 0; i < 13; i++) r.l[i] = f(i);
  return r;
}
DKGV_HD Fp30 fp30_zero() {
  Fp30 r;
#pragma unroll
  for (int i = 0; i < 13; i++) r.l[i] = 0;
  return r;
}
DKGV_HD Fp30 fp30_one() {
  Fp30 r;
#pragma unroll
  for (int i = 0; i < 13; i++) r.l[i] = consts::c30::ONE(i);
  return r;
}
DKGV_HD Fp30 fp30_twelve() {
  Fp30 r;
#pragma unroll
  for (int i = 0; i < 13; i++) r.l[i] = consts::c30::TWELVE(i);
  return r;
}

// canonical integer given as 12 x 32-bit limbs (< p)  ->  Fp30 Montgomery form
DKGV_HD Fp30 fp30_from_canonical(const uint32_t* c12) {
  Fp30 raw;
#pragma unroll
  for (int i = 0; i < 13; i++) {
    int bit = 30 * i, w = bit >> 5, sh = bit & 31;
    uint64_t two = (uint64_t)c12[w] | ((w + 1 < 12) ? ((uint64_t)c12[w + 1] << 32) : 0);
    raw.l[i] = (uint32_t)(two >> sh) & M30;
  }
  Fp30 r2;
#pragma unroll
  for (int i = 0; i < 13; i++) r2.l[i] = consts::c30::R2(i);
  return fp30_mul(raw, r2);
}
// Fp (12 x 32 Montgomery, field.cuh) -> Fp30
DKGV_HD Fp30 fp30_from_fp(const Fp& a) {
  Fp c = from_mont(a);
  return fp30_from_canonical(c.l);
}

// Fp30 (any lazy value) -> canonical 12 x 32-bit limbs in [0, p)
DKGV_HD void fp30_to_canonical(const Fp30& a, uint32_t* c12) {
  Fp30 o = fp30_zero();
  o.l[0] = 1;
  Fp30 v = fp30_mul(a, o);  // a / R, value < 2p, strictly normalised limbs
  // pack to 32-bit limbs
  uint32_t w[13];
#pragma unroll
  for (int i = 0; i < 13; i++) w[i] = 0;
#pragma unroll
  for (int i = 0; i < 13; i++) {
    int bit = 30 * i, k = bit >> 5, sh = bit & 31;
    uint64_t x = (uint64_t)v.l[i] << sh;
    w[k] |= (uint32_t)x;
    if (k + 1 < 13) w[k + 1] |= (uint32_t)(x >> 32);
  }
  // w < 2p < 2^384: one conditional subtraction
#pragma unroll
  for (int i = 0; i < 12; i++) c12[i] = w[i];
  cond_sub_mod<FpParams>(c12, 0);
}

DKGV_HD bool fp30_eq_mod_p(const Fp30& a, const Fp30& b) {
  uint32_t x[12], y[12];
  fp30_to_canonical(a, x);
  fp30_to_canonical(b, y);
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) d |= x[i] ^ y[i];
  return d == 0;
}
DKGV_HD bool fp30_is_zero_mod_p(const Fp30& a) {
  uint32_t x[12];
  fp30_to_canonical(a, x);
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) d |= x[i];
  return d == 0;
}

}  // namespace dkgv
