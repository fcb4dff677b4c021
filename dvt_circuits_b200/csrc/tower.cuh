// Fp2 / Fp6 / Fp12 tower, G2 (twist) arithmetic, optimal-ate pairing and RFC 9380 hash-to-G2 for the
// BLS partial-signature checks:
//   bls_verify_precomputed_hash / bls_verify   crates/dkg/src/crypto/bls_common.rs:26-40
//   hash_message_to_g2                          crates/dkg/src/crypto/bls_common.rs:11-24
//   BlsSignature::from_bytes(_safe)             crates/dkg/src/crypto/bls_keys.rs:165-185
// Tower: Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3-(1+u)), Fp12 = Fp6[w]/(w^2-v)  (SURVEY App. A).
//
// Execution model: one thread per pairing check.  An Fp12 is 144 words, so values live in local
// memory and every non-trivial routine is noinline with pointer arguments (one copy of each in the
// instruction cache - the lesson of profiles/r1_share_verify_v0_inlined.md).  Same code runs on the
// host for tests/hostemu.
#pragma once
#include "g1.cuh"

#if defined(__CUDACC__)
#define DKGV_NI2 __host__ __device__ __noinline__
#else
#define DKGV_NI2 static
#endif

namespace dkgv {

DKGV_NI2 void fpm(Fp* r, const Fp* a, const Fp* b) { *r = mul(*a, *b); }
DKGV_NI2 void fpa(Fp* r, const Fp* a, const Fp* b) { *r = add(*a, *b); }
DKGV_NI2 void fps(Fp* r, const Fp* a, const Fp* b) { *r = sub(*a, *b); }
struct ExpPm1Half { DKGV_HD uint32_t operator()(int i) const { return consts::P_MINUS_1_DIV2(i); } };
// a^e for the three public exponents of this file through the ONE product routine fpm (square & multiply, MSB first):
// which = 0: p - 2 (inverse, 0 -> 0), 1: (p + 1) / 4 (square-root candidate), 2: (p - 1) / 2 (Euler criterion)
DKGV_NI2 void fp_pow_ni(Fp* r, const Fp* a, int which) {
  Fp acc = one<FpParams>(), base = *a;
  bool started = false;
#pragma unroll 1
  for (int i = 11; i >= 0; i--) {
    uint32_t w = which == 0 ? consts::P_MINUS_2(i) : which == 1 ? consts::P_PLUS_1_DIV4(i) : consts::P_MINUS_1_DIV2(i);
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      if (started) fpm(&acc, &acc, &acc);
      if ((w >> b) & 1) {
        if (started)
          fpm(&acc, &acc, &base);
        else
          acc = base;
        started = true;
      }
    }
  }
  *r = acc;
}
DKGV_NI2 void fp_inv_ni(Fp* r, const Fp* a) { *r = fp_inv_bgcd(*a); }  // binary extended Euclid (field.cuh): ~8x fewer instructions than a^(p-2)
DKGV_NI2 void fp_sqrt_ni(Fp* r, const Fp* a) { fp_pow_ni(r, a, 1); }
DKGV_NI2 bool fp_is_square(const Fp* a) {
  if (is_zero(*a)) return true;
  Fp e;
  fp_pow_ni(&e, a, 2);
  return eq(e, one<FpParams>());
}
template <uint32_t (*F)(int)>
DKGV_HD Fp fp_const() {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = F(i);
  return r;
}

// ------------------------------------------------------------------------------------ Fp2
struct Fp2 {
  Fp c0, c1;
};
DKGV_HD Fp2 fp2_zero() { return Fp2{zero<FpParams>(), zero<FpParams>()}; }
DKGV_HD Fp2 fp2_one() { return Fp2{one<FpParams>(), zero<FpParams>()}; }
DKGV_HD bool fp2_is_zero(const Fp2& a) { return is_zero(a.c0) && is_zero(a.c1); }
DKGV_HD bool fp2_eq(const Fp2& a, const Fp2& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
template <uint32_t (*F0)(int), uint32_t (*F1)(int)>
DKGV_HD Fp2 fp2_const() {
  return Fp2{fp_const<F0>(), fp_const<F1>()};
}
// The Fp2 layer computes in registers: each routine loads its operands once, runs the (independent)
// Fp products with the inlined carry-chain code - three of them interleave in fp2_mul - and stores
// once.  Only Fp2 values travel through local memory; profiles/r1_bls_verify.md has the before/after.
DKGV_NI2 void fp2_add(Fp2* r, const Fp2* a, const Fp2* b) {
  Fp x = add(a->c0, b->c0), y = add(a->c1, b->c1);
  r->c0 = x;
  r->c1 = y;
}
DKGV_NI2 void fp2_sub(Fp2* r, const Fp2* a, const Fp2* b) {
  Fp x = sub(a->c0, b->c0), y = sub(a->c1, b->c1);
  r->c0 = x;
  r->c1 = y;
}
DKGV_NI2 void fp2_neg(Fp2* r, const Fp2* a) {
  Fp x = neg(a->c0), y = neg(a->c1);
  r->c0 = x;
  r->c1 = y;
}
DKGV_NI2 void fp2_dbl(Fp2* r, const Fp2* a) {
  Fp x = dbl(a->c0), y = dbl(a->c1);
  r->c0 = x;
  r->c1 = y;
}
// One copy of each product routine (DKGV_TOWER_INLINE_MUL restores the fully inlined Fp2 layer): the pairing
// kernel's instruction working set must stay near the instruction cache's ~3.4 k instructions
// (profiles/r1_fd_kernels.md addendum, profiles/r1_bls_verify.md).
DKGV_NI2 void fpm2a(Fp* r, const Fp* a, const Fp* b, const Fp* c, const Fp* d) { *r = mul2add(*a, *b, *c, *d); }
#ifndef DKGV_TOWER_INLINE_MUL
DKGV_NI2 void fp2_mul(Fp2* r, const Fp2* a, const Fp2* b) {  // c0 = a0 b0 - a1 b1, c1 = a0 b1 + a1 b0: two fused pairs
  Fp nb1 = neg(b->c1), x, y;
  fpm2a(&x, &a->c0, &b->c0, &a->c1, &nb1);
  fpm2a(&y, &a->c0, &b->c1, &a->c1, &b->c0);
  r->c0 = x;
  r->c1 = y;
}
DKGV_NI2 void fp2_sqr(Fp2* r, const Fp2* a) {  // (c0+c1)(c0-c1), 2 c0 c1
  Fp s = add(a->c0, a->c1), d = sub(a->c0, a->c1), x, m;
  fpm(&m, &a->c0, &a->c1);
  fpm(&x, &s, &d);
  r->c0 = x;
  r->c1 = dbl(m);
}
DKGV_NI2 void fp2_scale(Fp2* r, const Fp2* a, const Fp* s) {
  Fp k = *s, x, y;
  fpm(&x, &a->c0, &k);
  fpm(&y, &a->c1, &k);
  r->c0 = x;
  r->c1 = y;
}
#else
DKGV_NI2 void fp2_mul(Fp2* r, const Fp2* a, const Fp2* b) {  // Karatsuba, 3 M
  Fp a0 = a->c0, a1 = a->c1, b0 = b->c0, b1 = b->c1;
  Fp t0 = mul(a0, b0), t1 = mul(a1, b1), t2 = mul(add(a0, a1), add(b0, b1));
  r->c0 = sub(t0, t1);
  r->c1 = sub(sub(t2, t0), t1);
}
DKGV_NI2 void fp2_sqr(Fp2* r, const Fp2* a) {  // (c0+c1)(c0-c1), 2 c0 c1
  Fp a0 = a->c0, a1 = a->c1;
  Fp x = mul(add(a0, a1), sub(a0, a1)), m = mul(a0, a1);
  r->c0 = x;
  r->c1 = dbl(m);
}
DKGV_NI2 void fp2_scale(Fp2* r, const Fp2* a, const Fp* s) {
  Fp k = *s;
  Fp x = mul(a->c0, k), y = mul(a->c1, k);
  r->c0 = x;
  r->c1 = y;
}
#endif
DKGV_NI2 void fp2_conj(Fp2* r, const Fp2* a) {
  Fp x = a->c0, y = neg(a->c1);
  r->c0 = x;
  r->c1 = y;
}
DKGV_NI2 void fp2_mul_xi(Fp2* r, const Fp2* a) {  // * (1 + u)
  Fp a0 = a->c0, a1 = a->c1;
  r->c0 = sub(a0, a1);
  r->c1 = add(a0, a1);
}
DKGV_NI2 void fp2_inv(Fp2* r, const Fp2* a) {
  Fp n, t, z = zero<FpParams>();
  fpm(&n, &a->c0, &a->c0);
  fpm(&t, &a->c1, &a->c1);
  fpa(&n, &n, &t);
  fp_inv_ni(&n, &n);
  fpm(&r->c0, &a->c0, &n);
  fpm(&t, &a->c1, &n);
  fps(&r->c1, &z, &t);
}
DKGV_NI2 bool fp2_is_square(const Fp2* a) {
  Fp n, t;
  fpm(&n, &a->c0, &a->c0);
  fpm(&t, &a->c1, &a->c1);
  fpa(&n, &n, &t);
  return fp_is_square(&n);
}
// any square root (complex method); false when a is a non-residue
DKGV_NI2 bool fp2_sqrt(Fp2* out, const Fp2* a) {
  if (fp2_is_zero(*a)) {
    *out = fp2_zero();
    return true;
  }
  Fp s, t, z = zero<FpParams>();
  if (is_zero(a->c1)) {
    fp_sqrt_ni(&s, &a->c0);
    fpm(&t, &s, &s);
    if (eq(t, a->c0)) {
      out->c0 = s;
      out->c1 = z;
      return true;
    }
    Fp na;
    fps(&na, &z, &a->c0);
    fp_sqrt_ni(&s, &na);
    fpm(&t, &s, &s);
    out->c0 = z;
    out->c1 = s;
    return eq(t, na);
  }
  Fp n, alpha, delta, x0, x1, inv2 = fp_const<consts::INV2_M>();
  fpm(&n, &a->c0, &a->c0);
  fpm(&t, &a->c1, &a->c1);
  fpa(&n, &n, &t);
  fp_sqrt_ni(&alpha, &n);
  fpm(&t, &alpha, &alpha);
  if (!eq(t, n)) return false;
  fpa(&delta, &a->c0, &alpha);
  fpm(&delta, &delta, &inv2);
  fp_sqrt_ni(&x0, &delta);
  fpm(&t, &x0, &x0);
  if (!eq(t, delta)) {
    fps(&delta, &a->c0, &alpha);
    fpm(&delta, &delta, &inv2);
    fp_sqrt_ni(&x0, &delta);
    fpm(&t, &x0, &x0);
    if (!eq(t, delta)) return false;
  }
  fpa(&t, &x0, &x0);
  fp_inv_ni(&t, &t);
  fpm(&x1, &a->c1, &t);
  out->c0 = x0;
  out->c1 = x1;
  Fp2 chk;
  fp2_sqr(&chk, out);
  return fp2_eq(chk, *a);
}
DKGV_HD bool fp2_lex_largest(const Fp2& y) { return is_zero(y.c1) ? fp_lex_largest(y.c0) : fp_lex_largest(y.c1); }
DKGV_HD int fp2_sgn0(const Fp2& a) {
  Fp r0 = from_mont(a.c0), r1 = from_mont(a.c1);
  return (int)((r0.l[0] & 1) | ((is_zero(a.c0) ? 1u : 0u) & (r1.l[0] & 1)));
}

// ------------------------------------------------------------------------------------ Fp6
struct Fp6 {
  Fp2 c0, c1, c2;
};
DKGV_HD Fp6 fp6_zero() { return Fp6{fp2_zero(), fp2_zero(), fp2_zero()}; }
DKGV_HD Fp6 fp6_one() { return Fp6{fp2_one(), fp2_zero(), fp2_zero()}; }
DKGV_NI2 void fp6_add(Fp6* r, const Fp6* a, const Fp6* b) {
  fp2_add(&r->c0, &a->c0, &b->c0);
  fp2_add(&r->c1, &a->c1, &b->c1);
  fp2_add(&r->c2, &a->c2, &b->c2);
}
DKGV_NI2 void fp6_sub(Fp6* r, const Fp6* a, const Fp6* b) {
  fp2_sub(&r->c0, &a->c0, &b->c0);
  fp2_sub(&r->c1, &a->c1, &b->c1);
  fp2_sub(&r->c2, &a->c2, &b->c2);
}
DKGV_NI2 void fp6_neg(Fp6* r, const Fp6* a) {
  fp2_neg(&r->c0, &a->c0);
  fp2_neg(&r->c1, &a->c1);
  fp2_neg(&r->c2, &a->c2);
}
DKGV_NI2 void fp6_mul(Fp6* r, const Fp6* a, const Fp6* b) {
  Fp2 t0, t1, t2, s0, s1, u;
  Fp6 o;
  fp2_mul(&t0, &a->c0, &b->c0);
  fp2_mul(&t1, &a->c1, &b->c1);
  fp2_mul(&t2, &a->c2, &b->c2);
  fp2_add(&s0, &a->c1, &a->c2);
  fp2_add(&s1, &b->c1, &b->c2);
  fp2_mul(&u, &s0, &s1);
  fp2_sub(&u, &u, &t1);
  fp2_sub(&u, &u, &t2);
  fp2_mul_xi(&u, &u);
  fp2_add(&o.c0, &u, &t0);
  fp2_add(&s0, &a->c0, &a->c1);
  fp2_add(&s1, &b->c0, &b->c1);
  fp2_mul(&u, &s0, &s1);
  fp2_sub(&u, &u, &t0);
  fp2_sub(&u, &u, &t1);
  fp2_mul_xi(&s0, &t2);
  fp2_add(&o.c1, &u, &s0);
  fp2_add(&s0, &a->c0, &a->c2);
  fp2_add(&s1, &b->c0, &b->c2);
  fp2_mul(&u, &s0, &s1);
  fp2_sub(&u, &u, &t0);
  fp2_sub(&u, &u, &t2);
  fp2_add(&o.c2, &u, &t1);
  *r = o;
}
DKGV_NI2 void fp6_mul_v(Fp6* r, const Fp6* a) {
  Fp2 t;
  fp2_mul_xi(&t, &a->c2);
  Fp2 a0 = a->c0, a1 = a->c1;
  r->c0 = t;
  r->c1 = a0;
  r->c2 = a1;
}
DKGV_NI2 void fp6_scale(Fp6* r, const Fp6* a, const Fp2* s) {
  fp2_mul(&r->c0, &a->c0, s);
  fp2_mul(&r->c1, &a->c1, s);
  fp2_mul(&r->c2, &a->c2, s);
}
// a * (c0 + c1 v): 5 Fp2 products
DKGV_NI2 void fp6_mul_by_01(Fp6* r, const Fp6* a, const Fp2* c0, const Fp2* c1) {
  Fp2 aa, bb, t1, t2, t3, s, cs;
  fp2_mul(&aa, &a->c0, c0);
  fp2_mul(&bb, &a->c1, c1);
  fp2_add(&s, &a->c1, &a->c2);
  fp2_mul(&t1, &s, c1);
  fp2_sub(&t1, &t1, &bb);
  fp2_mul_xi(&t1, &t1);
  fp2_add(&t1, &t1, &aa);  // a0 c0 + xi a2 c1
  fp2_add(&s, &a->c0, &a->c2);
  fp2_mul(&t3, &s, c0);
  fp2_sub(&t3, &t3, &aa);
  fp2_add(&t3, &t3, &bb);  // a2 c0 + a1 c1
  fp2_add(&s, &a->c0, &a->c1);
  fp2_add(&cs, c0, c1);
  fp2_mul(&t2, &s, &cs);
  fp2_sub(&t2, &t2, &aa);
  fp2_sub(&t2, &t2, &bb);  // a0 c1 + a1 c0
  r->c0 = t1;
  r->c1 = t2;
  r->c2 = t3;
}
// a * (c1 v): 3 Fp2 products
DKGV_NI2 void fp6_mul_by_1(Fp6* r, const Fp6* a, const Fp2* c1) {
  Fp2 t0, t1, t2;
  fp2_mul(&t0, &a->c2, c1);
  fp2_mul_xi(&t0, &t0);
  fp2_mul(&t1, &a->c0, c1);
  fp2_mul(&t2, &a->c1, c1);
  r->c0 = t0;
  r->c1 = t1;
  r->c2 = t2;
}
DKGV_NI2 void fp6_inv(Fp6* r, const Fp6* a) {
  Fp2 t0, t1, t2, u, d;
  fp2_sqr(&t0, &a->c0);
  fp2_mul(&u, &a->c1, &a->c2);
  fp2_mul_xi(&u, &u);
  fp2_sub(&t0, &t0, &u);
  fp2_sqr(&t1, &a->c2);
  fp2_mul_xi(&t1, &t1);
  fp2_mul(&u, &a->c0, &a->c1);
  fp2_sub(&t1, &t1, &u);
  fp2_sqr(&t2, &a->c1);
  fp2_mul(&u, &a->c0, &a->c2);
  fp2_sub(&t2, &t2, &u);
  fp2_mul(&d, &a->c2, &t1);
  fp2_mul(&u, &a->c1, &t2);
  fp2_add(&d, &d, &u);
  fp2_mul_xi(&d, &d);
  fp2_mul(&u, &a->c0, &t0);
  fp2_add(&d, &d, &u);
  fp2_inv(&d, &d);
  fp2_mul(&r->c0, &t0, &d);
  fp2_mul(&r->c1, &t1, &d);
  fp2_mul(&r->c2, &t2, &d);
}
DKGV_NI2 void fp6_frob(Fp6* r, const Fp6* a) {
  Fp2 k1 = fp2_const<consts::FROB6_C1_C0, consts::FROB6_C1_C1>(), k2 = fp2_const<consts::FROB6_C2_C0, consts::FROB6_C2_C1>(), t;
  fp2_conj(&r->c0, &a->c0);
  fp2_conj(&t, &a->c1);
  fp2_mul(&r->c1, &t, &k1);
  fp2_conj(&t, &a->c2);
  fp2_mul(&r->c2, &t, &k2);
}

// ------------------------------------------------------------------------------------ Fp12
struct Fp12 {
  Fp6 c0, c1;
};
DKGV_HD Fp12 fp12_one() { return Fp12{fp6_one(), fp6_zero()}; }
DKGV_HD bool fp12_eq(const Fp12& a, const Fp12& b) {
  const Fp* x = &a.c0.c0.c0;
  const Fp* y = &b.c0.c0.c0;
  bool e = true;
  for (int i = 0; i < 12; i++) e = e && eq(x[i], y[i]);
  return e;
}
DKGV_NI2 void fp12_mul(Fp12* r, const Fp12* a, const Fp12* b) {
  Fp6 t0, t1, s0, s1, u;
  Fp12 o;
  fp6_mul(&t0, &a->c0, &b->c0);
  fp6_mul(&t1, &a->c1, &b->c1);
  fp6_add(&s0, &a->c0, &a->c1);
  fp6_add(&s1, &b->c0, &b->c1);
  fp6_mul(&u, &s0, &s1);
  fp6_sub(&u, &u, &t0);
  fp6_sub(&o.c1, &u, &t1);
  fp6_mul_v(&t1, &t1);
  fp6_add(&o.c0, &t0, &t1);
  *r = o;
}
DKGV_NI2 void fp12_sqr(Fp12* r, const Fp12* a) {
  Fp6 ab, s0, s1, t;
  Fp12 o;
  fp6_mul(&ab, &a->c0, &a->c1);
  fp6_add(&s0, &a->c0, &a->c1);
  fp6_mul_v(&s1, &a->c1);
  fp6_add(&s1, &s1, &a->c0);
  fp6_mul(&t, &s0, &s1);
  fp6_sub(&t, &t, &ab);
  fp6_mul_v(&s0, &ab);
  fp6_sub(&o.c0, &t, &s0);
  fp6_add(&o.c1, &ab, &ab);
  *r = o;
}
DKGV_NI2 void fp12_conj(Fp12* r, const Fp12* a) {
  r->c0 = a->c0;
  fp6_neg(&r->c1, &a->c1);
}
DKGV_NI2 void fp12_inv(Fp12* r, const Fp12* a) {
  Fp6 d, t;
  fp6_mul(&d, &a->c0, &a->c0);
  fp6_mul(&t, &a->c1, &a->c1);
  fp6_mul_v(&t, &t);
  fp6_sub(&d, &d, &t);
  fp6_inv(&d, &d);
  fp6_mul(&r->c0, &a->c0, &d);
  fp6_mul(&t, &a->c1, &d);
  fp6_neg(&r->c1, &t);
}
DKGV_NI2 void fp12_frob(Fp12* r, const Fp12* a) {
  Fp2 k = fp2_const<consts::FROB12_C1_C0, consts::FROB12_C1_C1>();
  Fp6 t;
  fp6_frob(&r->c0, &a->c0);
  fp6_frob(&t, &a->c1);
  fp6_scale(&r->c1, &t, &k);
}
// f *= (a + b v + (c v) w) : the sparse shape of a line function (coefficient slots 0, 1, 4);
// 13 Fp2 products instead of the 18 of a general product
DKGV_NI2 void fp12_mul_by_014(Fp12* f, const Fp2* a, const Fp2* b, const Fp2* c) {
  Fp6 aa, bb, s;
  Fp2 o;
  fp6_mul_by_01(&aa, &f->c0, a, b);
  fp6_mul_by_1(&bb, &f->c1, c);
  fp2_add(&o, b, c);
  fp6_add(&s, &f->c1, &f->c0);
  fp6_mul_by_01(&s, &s, a, &o);
  fp6_sub(&s, &s, &aa);
  fp6_sub(&f->c1, &s, &bb);
  fp6_mul_v(&bb, &bb);
  fp6_add(&f->c0, &bb, &aa);
}
// (a + b s)^2 in Fp4 = Fp2[s]/(s^2 - xi): c0 = a^2 + xi b^2, c1 = 2ab
DKGV_NI2 void fp4_sqr(Fp2* c0, Fp2* c1, const Fp2* a, const Fp2* b) {
  Fp2 t0, t1, t2;
  fp2_sqr(&t0, a);
  fp2_sqr(&t1, b);
  fp2_add(&t2, a, b);
  fp2_sqr(&t2, &t2);
  fp2_sub(&t2, &t2, &t0);
  fp2_sub(c1, &t2, &t1);
  fp2_mul_xi(&t1, &t1);
  fp2_add(c0, &t1, &t0);
}
// squaring in the cyclotomic subgroup (Granger-Scott, eprint 2009/565): 9 Fp2 squarings.  Valid for
// every value after the easy part of the final exponentiation.
DKGV_NI2 void fp12_cyc_sqr(Fp12* r, const Fp12* f) {
  Fp2 z0 = f->c0.c0, z4 = f->c0.c1, z3 = f->c0.c2, z2 = f->c1.c0, z1 = f->c1.c1, z5 = f->c1.c2;
  Fp2 t0, t1, t2, t3;
  fp4_sqr(&t0, &t1, &z0, &z1);
  fp2_sub(&z0, &t0, &z0);
  fp2_dbl(&z0, &z0);
  fp2_add(&z0, &z0, &t0);
  fp2_add(&z1, &t1, &z1);
  fp2_dbl(&z1, &z1);
  fp2_add(&z1, &z1, &t1);
  fp4_sqr(&t0, &t1, &z2, &z3);
  fp4_sqr(&t2, &t3, &z4, &z5);
  fp2_sub(&z4, &t0, &z4);
  fp2_dbl(&z4, &z4);
  fp2_add(&z4, &z4, &t0);
  fp2_add(&z5, &t1, &z5);
  fp2_dbl(&z5, &z5);
  fp2_add(&z5, &z5, &t1);
  fp2_mul_xi(&t0, &t3);
  fp2_add(&z2, &t0, &z2);
  fp2_dbl(&z2, &z2);
  fp2_add(&z2, &z2, &t0);
  fp2_sub(&z3, &t2, &z3);
  fp2_dbl(&z3, &z3);
  fp2_add(&z3, &z3, &t2);
  r->c0.c0 = z0;
  r->c0.c1 = z4;
  r->c0.c2 = z3;
  r->c1.c0 = z2;
  r->c1.c1 = z1;
  r->c1.c2 = z5;
}
DKGV_NI2 void fp12_pow_x(Fp12* r, const Fp12* a) {  // a^x for cyclotomic a (x < 0): conj(a^|x|)
  Fp12 acc = *a;
#pragma unroll 1
  for (int b = 62; b >= 0; b--) {
    fp12_cyc_sqr(&acc, &acc);
    if ((consts::X_ABS >> b) & 1) fp12_mul(&acc, &acc, a);
  }
  fp12_conj(r, &acc);
}

// ------------------------------------------------------------------------------------ G2 on the twist y^2 = x^3 + 4(1+u)
struct G2Aff {
  Fp2 x, y;
  uint32_t inf;
};
struct G2Proj {
  Fp2 x, y, z;
};
DKGV_NI2 void fp2_mul_b3(Fp2* r, const Fp2* a) {  // * 12 (1 + u)
  Fp2 t2, t4, t8;
  fp2_dbl(&t2, a);
  fp2_dbl(&t4, &t2);
  fp2_dbl(&t8, &t4);
  fp2_add(&t8, &t8, &t4);
  fp2_mul_xi(r, &t8);
}
DKGV_HD G2Proj g2_identity() { return G2Proj{fp2_zero(), fp2_one(), fp2_zero()}; }
DKGV_HD G2Proj g2_from_affine(const G2Aff& a) {
  if (a.inf) return g2_identity();
  return G2Proj{a.x, a.y, fp2_one()};
}
// RCB Alg. 7 over Fp2
DKGV_NI2 void g2_add(G2Proj* r, const G2Proj* p, const G2Proj* q) {
  Fp2 t0, t1, t2, t3, t4, x3, y3, z3, s0, s1;
  fp2_mul(&t0, &p->x, &q->x);
  fp2_mul(&t1, &p->y, &q->y);
  fp2_mul(&t2, &p->z, &q->z);
  fp2_add(&s0, &p->x, &p->y);
  fp2_add(&s1, &q->x, &q->y);
  fp2_mul(&t3, &s0, &s1);
  fp2_add(&s0, &t0, &t1);
  fp2_sub(&t3, &t3, &s0);
  fp2_add(&s0, &p->y, &p->z);
  fp2_add(&s1, &q->y, &q->z);
  fp2_mul(&t4, &s0, &s1);
  fp2_add(&s0, &t1, &t2);
  fp2_sub(&t4, &t4, &s0);
  fp2_add(&s0, &p->x, &p->z);
  fp2_add(&s1, &q->x, &q->z);
  fp2_mul(&y3, &s0, &s1);
  fp2_add(&s0, &t0, &t2);
  fp2_sub(&y3, &y3, &s0);
  fp2_dbl(&x3, &t0);
  fp2_add(&t0, &x3, &t0);
  fp2_mul_b3(&t2, &t2);
  fp2_add(&z3, &t1, &t2);
  fp2_sub(&t1, &t1, &t2);
  fp2_mul_b3(&y3, &y3);
  fp2_mul(&x3, &t4, &y3);
  fp2_mul(&t2, &t3, &t1);
  fp2_sub(&r->x, &t2, &x3);
  fp2_mul(&y3, &y3, &t0);
  fp2_mul(&t1, &t1, &z3);
  fp2_add(&r->y, &t1, &y3);
  fp2_mul(&t0, &t0, &t3);
  fp2_mul(&z3, &z3, &t4);
  fp2_add(&r->z, &z3, &t0);
}
// RCB Alg. 9 over Fp2
DKGV_NI2 void g2_dbl(G2Proj* r, const G2Proj* p) {
  Fp2 t0, t1, t2, x3, y3, z3;
  fp2_sqr(&t0, &p->y);
  fp2_dbl(&z3, &t0);
  fp2_dbl(&z3, &z3);
  fp2_dbl(&z3, &z3);
  fp2_mul(&t1, &p->y, &p->z);
  fp2_sqr(&t2, &p->z);
  fp2_mul_b3(&t2, &t2);
  fp2_mul(&x3, &t2, &z3);
  fp2_add(&y3, &t0, &t2);
  Fp2 px = p->x, py = p->y;
  fp2_mul(&r->z, &t1, &z3);
  fp2_dbl(&t1, &t2);
  fp2_add(&t2, &t1, &t2);
  fp2_sub(&t0, &t0, &t2);
  fp2_mul(&y3, &t0, &y3);
  fp2_add(&r->y, &x3, &y3);
  fp2_mul(&t1, &px, &py);
  fp2_mul(&x3, &t0, &t1);
  fp2_dbl(&r->x, &x3);
}
DKGV_NI2 bool g2_eq(const G2Proj* a, const G2Proj* b) {
  bool ia = fp2_is_zero(a->z), ib = fp2_is_zero(b->z);
  Fp2 l, r;
  fp2_mul(&l, &a->x, &b->z);
  fp2_mul(&r, &b->x, &a->z);
  bool e = fp2_eq(l, r);
  fp2_mul(&l, &a->y, &b->z);
  fp2_mul(&r, &b->y, &a->z);
  e = e && fp2_eq(l, r);
  return (ia || ib) ? (ia && ib) : e;
}
// [k]P for a public multi-limb scalar (little-endian 32-bit limbs), MSB first
template <class LimbFn>
DKGV_HD void g2_mul_public(G2Proj* r, const G2Proj* p, LimbFn limb, int nlimbs) {
  G2Proj acc = g2_identity();
  bool started = false;
#pragma unroll 1
  for (int i = nlimbs - 1; i >= 0; i--) {
    uint32_t w = limb(i);
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      if (started) g2_dbl(&acc, &acc);
      if ((w >> b) & 1) {
        if (started) g2_add(&acc, &acc, p);
        else acc = *p;
        started = true;
      }
    }
  }
  *r = acc;
}
struct LimbXAbs { DKGV_HD uint32_t operator()(int i) const { return (uint32_t)(consts::X_ABS >> (32 * i)); } };
struct LimbHEff { DKGV_HD uint32_t operator()(int i) const { return consts::H_EFF(i); } };
DKGV_NI2 void g2_to_affine(G2Aff* r, const G2Proj* p) {
  if (fp2_is_zero(p->z)) {
    r->x = fp2_zero();
    r->y = fp2_one();
    r->inf = 1;
    return;
  }
  Fp2 zi;
  fp2_inv(&zi, &p->z);
  fp2_mul(&r->x, &p->x, &zi);
  fp2_mul(&r->y, &p->y, &zi);
  r->inf = 0;
}
// psi(x, y) = (conj(x) PSI_X, conj(y) PSI_Y);  Q in G2  <=>  psi(Q) == [x]Q  (x < 0)
DKGV_NI2 bool g2_in_subgroup(const G2Aff* a) {
  if (a->inf) return true;
  Fp2 kx = fp2_const<consts::PSI_X_C0, consts::PSI_X_C1>(), ky = fp2_const<consts::PSI_Y_C0, consts::PSI_Y_C1>(), t;
  G2Proj psi, q = g2_from_affine(*a), xq;
  fp2_conj(&t, &a->x);
  fp2_mul(&psi.x, &t, &kx);
  fp2_conj(&t, &a->y);
  fp2_mul(&psi.y, &t, &ky);
  psi.z = fp2_one();
  g2_mul_public(&xq, &q, LimbXAbs(), 2);
  fp2_neg(&xq.y, &xq.y);  // [x]Q = -[|x|]Q
  return g2_eq(&psi, &xq);
}
// 96-byte compressed encoding (x.c1 || x.c0) -> affine; status as g1_decompress
DKGV_NI2 uint32_t g2_decompress(const uint8_t* in, G2Aff* out, bool check_subgroup) {
  uint8_t b[96];
  for (int i = 0; i < 96; i++) b[i] = in[i];
  bool fc = (b[0] >> 7) & 1, fi = (b[0] >> 6) & 1, fs = (b[0] >> 5) & 1;
  b[0] &= 0x1f;
  Fp x1r, x0r;
  fp_raw_from_be48(x1r.l, b);
  fp_raw_from_be48(x0r.l, b + 48);
  out->x = fp2_zero();
  out->y = fp2_one();
  out->inf = 1;
  if (!fc) return G1_DEC_BAD_FLAGS;
  if (!raw_lt_mod<FpParams>(x1r.l) || !raw_lt_mod<FpParams>(x0r.l)) return G1_DEC_X_RANGE;
  if (fi) return (fs || !is_zero(x1r) || !is_zero(x0r)) ? G1_DEC_BAD_FLAGS : G1_DEC_OK;
  Fp2 x{to_mont(x0r), to_mont(x1r)}, rhs, y;
  Fp four = dbl(dbl(one<FpParams>()));
  Fp2 bb{four, four};
  fp2_sqr(&rhs, &x);
  fp2_mul(&rhs, &rhs, &x);
  fp2_add(&rhs, &rhs, &bb);
  if (!fp2_sqrt(&y, &rhs)) return G1_DEC_NOT_ON_CURVE;
  if (fp2_lex_largest(y) != fs) fp2_neg(&y, &y);
  out->x = x;
  out->y = y;
  out->inf = 0;
  if (check_subgroup && !g2_in_subgroup(out)) {
    out->inf = 1;
    return G1_DEC_NOT_IN_SUBGROUP;
  }
  return G1_DEC_OK;
}
DKGV_NI2 void g2_compress(const G2Aff* a, uint8_t* out) {
  if (a->inf) {
    for (int i = 0; i < 96; i++) out[i] = 0;
    out[0] = 0xc0;
    return;
  }
  Fp c1 = from_mont(a->x.c1), c0 = from_mont(a->x.c0);
  fp_raw_to_be48(out, c1.l);
  fp_raw_to_be48(out + 48, c0.l);
  out[0] |= 0x80;
  if (fp2_lex_largest(a->y)) out[0] |= 0x20;
}

// ------------------------------------------------------------------------------------ pairing
// Line functions through twist points, evaluated at P = (xp, yp); scaling by Fp2/Fp4 factors is
// killed by the final exponentiation:
//   tangent at T = (X:Y:Z):  c00 = Y^2 - 3b'Z^2, c01 = -3X^2 xp, c11 = 2YZ yp       then T <- 2T
//   chord T,Q:  N = yQ Z - Y, D = xQ Z - X: c00 = N xQ - D yQ, c01 = -N xp, c11 = D yp   then T <- T+Q
// A G2Line holds the P-independent part (c01, c11 before the scaling by xp, yp), so the lines of a G2
// point shared by many checks - the hashed message - are computed once (g2_prepare) and every check
// only pays the two scalings and the sparse product.
struct G2Line {
  Fp2 c00, c01, c11;
};
constexpr int G2_PREP_LINES = 68;  // 63 tangents + 5 chords over the bits of |x| below the top one
DKGV_NI2 void g2_line_dbl(G2Line* l, G2Proj* t) {
  Fp2 u;
  fp2_sqr(&l->c00, &t->y);
  fp2_sqr(&u, &t->z);
  fp2_mul_b3(&u, &u);
  fp2_sub(&l->c00, &l->c00, &u);
  fp2_sqr(&u, &t->x);
  fp2_dbl(&l->c01, &u);
  fp2_add(&l->c01, &l->c01, &u);
  fp2_neg(&l->c01, &l->c01);
  fp2_mul(&l->c11, &t->y, &t->z);
  fp2_dbl(&l->c11, &l->c11);
  g2_dbl(t, t);
}
DKGV_NI2 void g2_line_add(G2Line* l, G2Proj* t, const G2Aff* q) {
  Fp2 n, d, u;
  fp2_mul(&n, &q->y, &t->z);
  fp2_sub(&n, &n, &t->y);
  fp2_mul(&d, &q->x, &t->z);
  fp2_sub(&d, &d, &t->x);
  fp2_mul(&l->c00, &n, &q->x);
  fp2_mul(&u, &d, &q->y);
  fp2_sub(&l->c00, &l->c00, &u);
  fp2_neg(&l->c01, &n);
  l->c11 = d;
  G2Proj qq = g2_from_affine(*q);
  g2_add(t, t, &qq);
}
DKGV_NI2 void fp12_mul_line(Fp12* f, const G2Line* l, const Fp* xp, const Fp* yp) {
  Fp2 c01, c11;
  fp2_scale(&c01, &l->c01, xp);
  fp2_scale(&c11, &l->c11, yp);
  fp12_mul_by_014(f, &l->c00, &c01, &c11);
}
// the G2_PREP_LINES lines of q in Miller-loop order (q must not be the identity)
DKGV_NI2 void g2_prepare(G2Line* out, const G2Aff* q) {
  G2Proj t = g2_from_affine(*q);
  int k = 0;
#pragma unroll 1
  for (int b = 62; b >= 0; b--) {
    g2_line_dbl(&out[k++], &t);
    if ((consts::X_ABS >> b) & 1) g2_line_add(&out[k++], &t, q);
  }
}
// f <- prod_i f_{|x|,Q_i}(P_i) over up to two pairs with ONE shared chain of Fp12 squarings (not yet
// conjugated).  Pair 1 may come with prepared lines (prep1 != nullptr).  A pair with an identity
// argument contributes the Gt identity (bls12_381::pairing semantics) and is skipped.
DKGV_NI2 void miller_loop_2(Fp12* f, const G1Aff* p1, const G2Aff* q1, const G2Line* prep1, const G1Aff* p2, const G2Aff* q2) {
  bool on1 = p1 && !(p1->inf || q1->inf), on2 = p2 && !(p2->inf || q2->inf);
  *f = fp12_one();
  if (!on1 && !on2) return;
  G2Proj t1, t2;
  if (on1 && !prep1) t1 = g2_from_affine(*q1);
  if (on2) t2 = g2_from_affine(*q2);
  G2Line l;
  int k = 0;
#pragma unroll 1
  for (int b = 62; b >= 0; b--) {
    if (b != 62) fp12_sqr(f, f);
    bool bit = (consts::X_ABS >> b) & 1;
    if (on1) {
      if (prep1) {
        fp12_mul_line(f, &prep1[k], &p1->x, &p1->y);
      } else {
        g2_line_dbl(&l, &t1);
        fp12_mul_line(f, &l, &p1->x, &p1->y);
      }
    }
    if (on2) {
      g2_line_dbl(&l, &t2);
      fp12_mul_line(f, &l, &p2->x, &p2->y);
    }
    k++;
    if (bit) {
      if (on1) {
        if (prep1) {
          fp12_mul_line(f, &prep1[k], &p1->x, &p1->y);
        } else {
          g2_line_add(&l, &t1, q1);
          fp12_mul_line(f, &l, &p1->x, &p1->y);
        }
      }
      if (on2) {
        g2_line_add(&l, &t2, q2);
        fp12_mul_line(f, &l, &p2->x, &p2->y);
      }
      k++;
    }
  }
}
// f <- f * f_{|x|,Q}(P)  (not yet conjugated); identity arguments as above
DKGV_NI2 void miller_loop_acc(Fp12* f, const G1Aff* p, const G2Aff* q) {
  if (p->inf || q->inf) return;
  Fp12 g;
  miller_loop_2(&g, p, q, nullptr, nullptr, nullptr);
  fp12_mul(f, f, &g);
}
// f^(3 (p^12 - 1) / r) via 3(p^4-p^2+1)/r = (x-1)^2 (x+p)(x^2+p^2-1) + 3; input already conjugated
DKGV_NI2 void final_exponentiation(Fp12* r, const Fp12* f0) {
  Fp12 f, t0, t1, t2, t3, u;
  fp12_conj(&t0, f0);
  fp12_inv(&t1, f0);
  fp12_mul(&f, &t0, &t1);  // f^(p^6-1)
  fp12_frob(&t0, &f);
  fp12_frob(&t0, &t0);
  fp12_mul(&f, &t0, &f);  // ^(p^2+1)
  fp12_pow_x(&t0, &f);
  fp12_conj(&u, &f);
  fp12_mul(&t0, &t0, &u);  // f^(x-1)
  fp12_pow_x(&t1, &t0);
  fp12_conj(&u, &t0);
  fp12_mul(&t1, &t1, &u);  // ^(x-1)
  fp12_pow_x(&t2, &t1);
  fp12_frob(&u, &t1);
  fp12_mul(&t2, &t2, &u);  // ^(x+p)
  fp12_pow_x(&t3, &t2);
  fp12_pow_x(&t3, &t3);
  fp12_frob(&u, &t2);
  fp12_frob(&u, &u);
  fp12_mul(&t3, &t3, &u);
  fp12_conj(&u, &t2);
  fp12_mul(&t3, &t3, &u);  // ^(x^2+p^2-1)
  fp12_sqr(&u, &f);
  fp12_mul(&u, &u, &f);  // f^3
  fp12_mul(r, &t3, &u);
}
// e(pk, hm) == e(G1, sig)  (bls_common.rs:26-35) as ONE product of two Miller loops (shared squarings)
// and one final exponentiation: e(pk, hm) * e(-G1, sig) == 1.  Same boolean as the reference's two
// pairings.  hm_lines: the prepared lines of hm, or nullptr to compute them here.
DKGV_NI2 bool bls_verify_prepared(const G1Aff* pk, const G2Aff* sig, const G2Aff* hm, const G2Line* hm_lines) {
  Fp12 f, e;
  G1Aff ng = g1_generator();
  ng.y = neg(ng.y);
  miller_loop_2(&f, pk, hm, hm_lines, &ng, sig);
  fp12_conj(&f, &f);  // x < 0
  final_exponentiation(&e, &f);
  return fp12_eq(e, fp12_one());
}
DKGV_NI2 bool bls_verify_precomputed(const G1Aff* pk, const G2Aff* sig, const G2Aff* hm) {
  return bls_verify_prepared(pk, sig, hm, nullptr);
}

}  // namespace dkgv
