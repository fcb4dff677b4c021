// Hot-kernel execution model, v2: the operand-file design of vm.cuh on the carry-free 13 x 30-bit
// field backend of fp30.cuh.  Every routine exists once (noinline); operands live in shared memory.
//
// Slot layout: one Fp30 = 13 words, stored as four 16-byte chunks (words 13..15 are padding; word 13
// carries the static bound in DKGV_BOUND_CHECK host builds):  file[(s*4 + c) * NT + t].
//
// HBM layouts consumed here (produced by k_decompress_vv30 / k_build_gtab30 in dkgv.cu):
//   vv30   limbs[(k*26 + w) * n_pad + d]   w = 0..12 x, 13..25 y   (Fp30 Montgomery form, canonical value)
//          inf[k * n_pad + d]
//   gtab30 [(w*256 + b) * 26 ..]           offset fixed-base table, see feldman.cuh
//
// Bounds (multiples of p, see fp30.cuh): point coordinates entering add / dbl / madd are <= 16 and
// every formula returns coordinates <= 16; each subtraction states the K it needs.
#pragma once
#include "fp30.cuh"
#include "../../dvt_circuits_b200/csrc/vm.cuh"

namespace dkgv {

struct OpFile30 {
  U4* base;
  uint32_t stride;
};

DKGV_HD Fp30 o30_load(const OpFile30& f, int s) {
  Fp30 r;
  uint32_t w[16];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    U4 v = f.base[(size_t)(s * 4 + c) * f.stride];
    w[4 * c] = v.x;
    w[4 * c + 1] = v.y;
    w[4 * c + 2] = v.z;
    w[4 * c + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < 13; i++) r.l[i] = w[i];
#if defined(DKGV_BOUND_CHECK) && !defined(__CUDA_ARCH__)
  float b;
  memcpy(&b, &w[13], 4);
  r.bd = b;
#endif
  return r;
}
DKGV_HD void o30_store(const OpFile30& f, int s, const Fp30& a) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 13; i++) w[i] = a.l[i];
  w[13] = w[14] = w[15] = 0;
#if defined(DKGV_BOUND_CHECK) && !defined(__CUDA_ARCH__)
  float b = (float)(a.bd * 1.0001);
  memcpy(&w[13], &b, 4);
#endif
#pragma unroll
  for (int c = 0; c < 4; c++) {
    U4 v;
    v.x = w[4 * c];
    v.y = w[4 * c + 1];
    v.z = w[4 * c + 2];
    v.w = w[4 * c + 3];
    f.base[(size_t)(s * 4 + c) * f.stride] = v;
  }
}

DKGV_NI void v30_mul(OpFile30 f, int d, int a, int b) { o30_store(f, d, fp30_mul(o30_load(f, a), o30_load(f, b))); }
DKGV_NI void v30_add(OpFile30 f, int d, int a, int b) { o30_store(f, d, fp30_add(o30_load(f, a), o30_load(f, b))); }
DKGV_NI void v30_sub8(OpFile30 f, int d, int a, int b) { o30_store(f, d, fp30_sub<8>(o30_load(f, a), o30_load(f, b))); }
DKGV_NI void v30_sub32(OpFile30 f, int d, int a, int b) { o30_store(f, d, fp30_sub<32>(o30_load(f, a), o30_load(f, b))); }
DKGV_NI void v30_sub64(OpFile30 f, int d, int a, int b) { o30_store(f, d, fp30_sub<64>(o30_load(f, a), o30_load(f, b))); }
DKGV_NI void v30_mul12(OpFile30 f, int d, int a) { o30_store(f, d, fp30_mul_small<12>(o30_load(f, a))); }
DKGV_NI void v30_mul8(OpFile30 f, int d, int a) { o30_store(f, d, fp30_mul_small<8>(o30_load(f, a))); }
DKGV_NI void v30_mul3(OpFile30 f, int d, int a) { o30_store(f, d, fp30_mul_small<3>(o30_load(f, a))); }
DKGV_NI void v30_mul2(OpFile30 f, int d, int a) { o30_store(f, d, fp30_mul_small<2>(o30_load(f, a))); }
// d = a * (12 in Montgomery form): a reduced 12a (bound ~1) where 12 limb-wise would overshoot
DKGV_NI void v30_mul12r(OpFile30 f, int d, int a) { o30_store(f, d, fp30_mul(o30_load(f, a), fp30_twelve())); }
// d = (a1 + a2) * (b1 + b2)
DKGV_NI void v30_addmul(OpFile30 f, int d, int a1, int a2, int b1, int b2) {
  o30_store(f, d, fp30_mul(fp30_add(o30_load(f, a1), o30_load(f, a2)), fp30_add(o30_load(f, b1), o30_load(f, b2))));
}
DKGV_NI void v30_copy3(OpFile30 f, int d, int a) {
#pragma unroll
  for (int i = 0; i < 12; i++) f.base[(size_t)(d * 4 + i) * f.stride] = f.base[(size_t)(a * 4 + i) * f.stride];
}

DKGV_HD void v30_set_point(const OpFile30& f, int s, const Fp30& x, const Fp30& y, const Fp30& z) {
  o30_store(f, s, x);
  o30_store(f, s + 1, y);
  o30_store(f, s + 2, z);
}
DKGV_HD void v30_set_identity(const OpFile30& f, int s) { v30_set_point(f, s, fp30_zero(), fp30_one(), fp30_zero()); }

// A <- A + B (RCB Alg. 7).  in: all coordinates <= 16; out <= 16
DKGV_NI void v30_g1_add(OpFile30 f) {
  v30_mul(f, T0, AX, BX);
  v30_mul(f, T1, AY, BY);
  v30_mul(f, T2, AZ, BZ);
  v30_addmul(f, T3, AX, AY, BX, BY);
  v30_add(f, T4, T0, T1);
  v30_sub8(f, T3, T3, T4);
  v30_addmul(f, T4, AY, AZ, BY, BZ);
  v30_add(f, T5, T1, T2);
  v30_sub8(f, T4, T4, T5);
  v30_addmul(f, T5, AX, AZ, BX, BZ);
  v30_add(f, T6, T0, T2);
  v30_sub8(f, T5, T5, T6);  // "Y3" of the paper; A dead from here on
  v30_mul3(f, T0, T0);
  v30_mul12(f, T2, T2);
  v30_add(f, AZ, T1, T2);
  v30_sub32(f, T1, T1, T2);
  v30_mul12(f, T5, T5);
  v30_mul(f, AX, T4, T5);
  v30_mul(f, T2, T3, T1);
  v30_sub8(f, AX, T2, AX);
  v30_mul(f, T5, T5, T0);
  v30_mul(f, T1, T1, AZ);
  v30_add(f, AY, T1, T5);
  v30_mul(f, T0, T0, T3);
  v30_mul(f, AZ, AZ, T4);
  v30_add(f, AZ, AZ, T0);
}

// A <- 2A (RCB Alg. 9).  in <= 16, out <= 16
DKGV_NI void v30_g1_dbl(OpFile30 f) {
  v30_mul(f, T0, AY, AY);
  v30_mul8(f, T3, T0);  // Z3' = 8 Y^2
  v30_mul(f, T1, AY, AZ);
  v30_mul(f, T2, AZ, AZ);
  v30_mul12(f, T2, T2);
  v30_mul(f, T4, T2, T3);  // X3'
  v30_add(f, T5, T0, T2);  // Y3'
  v30_mul(f, AZ, T1, T3);
  v30_mul3(f, T2, T2);
  v30_sub64(f, T0, T0, T2);
  v30_mul(f, T5, T0, T5);
  v30_mul(f, T1, AX, AY);
  v30_add(f, AY, T4, T5);
  v30_mul(f, T4, T0, T1);
  v30_mul2(f, AX, T4);
}

// P <- P + Q, P at slots (p..p+2) with coordinates <= 16, Q = (x, y) affine canonical at (T5, T6),
// never the identity (RCB Alg. 8; 12*Z1 is taken through a Montgomery product so it stays reduced)
DKGV_NI void v30_g1_madd(OpFile30 f, int p) {
  const int X1 = p, Y1 = p + 1, Z1 = p + 2, QX = T5, QY = T6;
  v30_mul(f, T0, X1, QX);
  v30_mul(f, T1, Y1, QY);
  v30_addmul(f, T3, QX, QY, X1, Y1);
  v30_add(f, T4, T0, T1);
  v30_sub8(f, T3, T3, T4);
  v30_mul(f, T4, QY, Z1);
  v30_add(f, T4, T4, Y1);
  v30_mul(f, T2, QX, Z1);
  v30_add(f, T2, T2, X1);  // "Y3" pre; Q, X1, Y1 dead
  v30_mul12r(f, T5, Z1);   // t2 = 12 Z1 (reduced)
  v30_mul3(f, T0, T0);
  v30_add(f, Z1, T1, T5);
  v30_sub8(f, T1, T1, T5);
  v30_mul12(f, T2, T2);
  v30_mul(f, X1, T4, T2);
  v30_mul(f, T5, T3, T1);
  v30_sub8(f, X1, T5, X1);
  v30_mul(f, T2, T2, T0);
  v30_mul(f, T1, T1, Z1);
  v30_add(f, Y1, T1, T2);
  v30_mul(f, T0, T0, T3);
  v30_mul(f, Z1, Z1, T4);
  v30_add(f, Z1, Z1, T0);
}

DKGV_HD void v30_g1_mul_small(const OpFile30& f, uint32_t k) {
  if (k == 0) {
    v30_set_identity(f, AX);
    return;
  }
  if ((k & (k - 1)) != 0) v30_copy3(f, BX, AX);
  int top = 31;
  while (!((k >> top) & 1)) top--;
#pragma unroll 1
  for (int b = top - 1; b >= 0; b--) {
    v30_g1_dbl(f);
    if ((k >> b) & 1) v30_g1_add(f);
  }
}

struct VV30View {
  const uint32_t* limbs;
  const uint8_t* inf;
  uint32_t n_pad;
};

DKGV_HD void vv30_store(uint32_t* limbs, uint8_t* inf, uint32_t n_pad, uint32_t k, uint32_t d, const Fp30& x, const Fp30& y, bool is_inf) {
  uint32_t* base = limbs + (size_t)k * 26 * n_pad + d;
#pragma unroll
  for (int w = 0; w < 13; w++) {
    base[(size_t)w * n_pad] = x.l[w];
    base[(size_t)(w + 13) * n_pad] = y.l[w];
  }
  inf[(size_t)k * n_pad + d] = is_inf ? 1 : 0;
}

// coefficient k of dealer d -> projective point at (s..s+2); identity -> (0 : 1 : 0)
DKGV_HD void v30_load_coeff(const OpFile30& f, int s, const VV30View& v, uint32_t k, uint32_t d) {
  const uint32_t* base = v.limbs + (size_t)k * 26 * v.n_pad + d;
  Fp30 x, y;
#pragma unroll
  for (int w = 0; w < 13; w++) {
    x.l[w] = base[(size_t)w * v.n_pad];
    y.l[w] = base[(size_t)(w + 13) * v.n_pad];
  }
  bool inf = v.inf[(size_t)k * v.n_pad + d] != 0;
  Fp30 one = fp30_one(), zero = fp30_zero();
#pragma unroll
  for (int w = 0; w < 13; w++) {
    x.l[w] = inf ? 0u : x.l[w];
    y.l[w] = inf ? one.l[w] : y.l[w];
    one.l[w] = inf ? 0u : one.l[w];
  }
  (void)zero;
  v30_set_point(f, s, x, y, one);
}

// Horner in the exponent (crates/dkg/src/dkg_math.rs:160-174): A <- sum_k C_k id^k
DKGV_HD void v30_feldman_eval(const OpFile30& f, const VV30View& v, uint32_t t, uint32_t d, uint32_t id) {
  if (t == 0) {
    v30_set_identity(f, AX);
    return;
  }
  v30_load_coeff(f, AX, v, t - 1, d);
#pragma unroll 1
  for (int k = (int)t - 2; k >= 0; k--) {
    v30_g1_mul_small(f, id);
    v30_load_coeff(f, BX, v, (uint32_t)k, d);
    v30_g1_add(f);
  }
}

constexpr size_t GTAB30_WORDS = (size_t)GTAB_ENTRIES * 26;

// B <- G * s through the offset fixed-base table (33 mixed additions, no identity cases)
DKGV_HD void v30_fixed_base_mul(const OpFile30& f, const uint32_t* gtab30, const uint32_t* s_raw) {
  v30_set_identity(f, BX);
#pragma unroll 1
  for (int w = 0; w <= GTAB_WINDOWS; w++) {
    const uint32_t* e = gtab30 + (size_t)gtab_index(s_raw, w) * 26;
    Fp30 x, y;
#pragma unroll
    for (int i = 0; i < 13; i++) {
      x.l[i] = e[i];
      y.l[i] = e[13 + i];
    }
    o30_store(f, T5, x);
    o30_store(f, T6, y);
    v30_g1_madd(f, BX);
  }
}

// canonical comparison helpers on slots holding values already divided by R (v30_mul by raw 1):
// such values are < 2p with strictly normalised limbs, so one conditional subtraction finishes them
DKGV_HD void fp30_pack_reduce(const Fp30& v, uint32_t* c12) {
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
#pragma unroll
  for (int i = 0; i < 12; i++) c12[i] = w[i];
  cond_sub_mod<FpParams>(c12, 0);
}
DKGV_NI bool v30_canon_eq(OpFile30 f, int a, int b) {
  uint32_t x[12], y[12];
  fp30_pack_reduce(o30_load(f, a), x);
  fp30_pack_reduce(o30_load(f, b), y);
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) d |= x[i] ^ y[i];
  return d == 0;
}
DKGV_NI bool v30_canon_is_zero(OpFile30 f, int a) {
  uint32_t x[12];
  fp30_pack_reduce(o30_load(f, a), x);
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) d |= x[i];
  return d == 0;
}

// A == B as projective points (all Montgomery products go through the single v30_mul)
DKGV_HD bool v30_g1_eq_ab(const OpFile30& f) {
  Fp30 raw1 = fp30_zero();
  raw1.l[0] = 1;
  o30_store(f, T6, raw1);
  v30_mul(f, T4, AZ, T6);
  v30_mul(f, T5, BZ, T6);
  bool ia = v30_canon_is_zero(f, T4), ib = v30_canon_is_zero(f, T5);
  v30_mul(f, T0, AX, BZ);
  v30_mul(f, T1, BX, AZ);
  v30_mul(f, T2, AY, BZ);
  v30_mul(f, T3, BY, AZ);
  v30_mul(f, T0, T0, T6);
  v30_mul(f, T1, T1, T6);
  v30_mul(f, T2, T2, T6);
  v30_mul(f, T3, T3, T6);
  bool e = v30_canon_eq(f, T0, T1) && v30_canon_eq(f, T2, T3);
  return (ia || ib) ? (ia && ib) : e;
}

// one share: same contract as share_check (feldman.cuh) / vm_share_check (vm.cuh)
DKGV_HD uint8_t v30_share_check(const OpFile30& f, const VV30View& vv, uint32_t t, uint32_t d, uint32_t id, const uint8_t* secret_be,
                                const uint32_t* gtab30, bool dealer_bad) {
  v30_feldman_eval(f, vv, t, d, id);
  uint32_t s[8];
  bool in_range = fr_raw_from_be32(s, secret_be);
  v30_fixed_base_mul(f, gtab30, s);
  uint8_t st = v30_g1_eq_ab(f) ? DKGV_OK : DKGV_SLASHABLE_SHARE_MISMATCH;
  if (dealer_bad) st = DKGV_PANIC_BAD_G1;
  if (!in_range) st = DKGV_SLASHABLE_SECRET_RANGE;
  return st;
}

}  // namespace dkgv
