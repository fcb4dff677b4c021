// BLS12-381 G1 group arithmetic: homogeneous projective coordinates with the complete formulas of
// Renes-Costello-Batina (eprint 2015/1060, Algorithms 7-9 for a = 0, b3 = 12) - the same formula
// family the reference's dependency uses, so the operation counts of SURVEY.md 8(d) apply:
// add 12 M, mixed add 11 M, double 6 M + 2 S.  No exceptional cases: identity is (0 : y : 0).
//
// Replaces BlsG1 / TPoint (crates/dkg/src/dkg_math.rs:11-13,104-128) and G1 decoding
// (crates/dkg/src/crypto/bls_common.rs:108-112).  Encodings per SURVEY.md App. B.
#pragma once
#include "field.cuh"

namespace dkgv {

struct G1Proj {
  Fp x, y, z;
};
struct G1Aff {  // Montgomery-form affine point; inf != 0 means the identity (x, y ignored)
  Fp x, y;
  uint32_t inf;
};

DKGV_HD Fp fp_mul12(const Fp& a) {  // b3 * a, additions only
  Fp t2 = dbl(a), t4 = dbl(t2), t8 = dbl(t4);
  return add(t8, t4);
}

DKGV_HD G1Proj g1_identity() {
  G1Proj r;
  r.x = zero<FpParams>();
  r.y = one<FpParams>();
  r.z = zero<FpParams>();
  return r;
}
DKGV_HD bool g1_is_identity(const G1Proj& p) { return is_zero(p.z); }

DKGV_HD G1Proj g1_from_affine(const G1Aff& a) {
  G1Proj r;
  bool inf = a.inf != 0;
  r.x = select(a.x, zero<FpParams>(), inf);
  r.y = select(a.y, one<FpParams>(), inf);
  r.z = select(one<FpParams>(), zero<FpParams>(), inf);
  return r;
}

DKGV_HD G1Aff g1_generator() {
  G1Aff g;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    g.x.l[i] = consts::G1X_M(i);
    g.y.l[i] = consts::G1Y_M(i);
  }
  g.inf = 0;
  return g;
}

// RCB Algorithm 7
DKGV_HD G1Proj g1_add(const G1Proj& p, const G1Proj& q) {
  Fp t0 = mul(p.x, q.x), t1 = mul(p.y, q.y), t2 = mul(p.z, q.z);
  Fp t3 = mul(add(p.x, p.y), add(q.x, q.y));
  t3 = sub(t3, add(t0, t1));
  Fp t4 = mul(add(p.y, p.z), add(q.y, q.z));
  t4 = sub(t4, add(t1, t2));
  Fp y3 = mul(add(p.x, p.z), add(q.x, q.z));
  y3 = sub(y3, add(t0, t2));
  Fp x3 = dbl(t0);
  t0 = add(x3, t0);
  t2 = fp_mul12(t2);
  Fp z3 = add(t1, t2);
  t1 = sub(t1, t2);
  y3 = fp_mul12(y3);
  x3 = mul(t4, y3);
  t2 = mul(t3, t1);
  G1Proj r;
  r.x = sub(t2, x3);
  y3 = mul(y3, t0);
  t1 = mul(t1, z3);
  r.y = add(t1, y3);
  t0 = mul(t0, t3);
  z3 = mul(z3, t4);
  r.z = add(z3, t0);
  return r;
}

// RCB Algorithm 8 (q affine, not the identity)
DKGV_HD G1Proj g1_add_mixed_nz(const G1Proj& p, const Fp& qx, const Fp& qy) {
  Fp t0 = mul(p.x, qx), t1 = mul(p.y, qy);
  Fp t3 = mul(add(qx, qy), add(p.x, p.y));
  t3 = sub(t3, add(t0, t1));
  Fp t4 = add(mul(qy, p.z), p.y);
  Fp y3 = add(mul(qx, p.z), p.x);
  Fp x3 = dbl(t0);
  t0 = add(x3, t0);
  Fp t2 = fp_mul12(p.z);
  Fp z3 = add(t1, t2);
  t1 = sub(t1, t2);
  y3 = fp_mul12(y3);
  x3 = mul(t4, y3);
  t2 = mul(t3, t1);
  G1Proj r;
  r.x = sub(t2, x3);
  y3 = mul(y3, t0);
  t1 = mul(t1, z3);
  r.y = add(t1, y3);
  t0 = mul(t0, t3);
  z3 = mul(z3, t4);
  r.z = add(z3, t0);
  return r;
}
// mixed add where q may be the identity (branch-free select)
DKGV_HD G1Proj g1_add_mixed(const G1Proj& p, const G1Aff& q) {
  G1Proj s = g1_add_mixed_nz(p, q.x, q.y);
  bool inf = q.inf != 0;
  G1Proj r;
  r.x = select(s.x, p.x, inf);
  r.y = select(s.y, p.y, inf);
  r.z = select(s.z, p.z, inf);
  return r;
}

// RCB Algorithm 9
DKGV_HD G1Proj g1_dbl(const G1Proj& p) {
  Fp t0 = sqr(p.y);
  Fp z3 = dbl(dbl(dbl(t0)));
  Fp t1 = mul(p.y, p.z);
  Fp t2 = fp_mul12(sqr(p.z));
  Fp x3 = mul(t2, z3);
  Fp y3 = add(t0, t2);
  G1Proj r;
  r.z = mul(t1, z3);
  t1 = dbl(t2);
  t2 = add(t1, t2);
  t0 = sub(t0, t2);
  y3 = mul(t0, y3);
  r.y = add(x3, y3);
  t1 = mul(p.x, p.y);
  x3 = mul(t0, t1);
  r.x = dbl(x3);
  return r;
}

DKGV_HD G1Proj g1_neg(const G1Proj& p) {
  G1Proj r = p;
  r.y = neg(p.y);
  return r;
}

// projective equality (== equality of the compressed encodings the reference compares,
// crates/dkg/src/verification.rs:140,301,320)
DKGV_HD bool g1_eq(const G1Proj& a, const G1Proj& b) {
  bool ia = is_zero(a.z), ib = is_zero(b.z);
  bool e = eq(mul(a.x, b.z), mul(b.x, a.z)) && eq(mul(a.y, b.z), mul(b.y, a.z));
  return (ia || ib) ? (ia && ib) : e;
}

// [k]P, k a 64-bit public scalar, MSB-first double-and-add (control flow depends on k only)
DKGV_HD G1Proj g1_mul_u64(const G1Proj& p, unsigned long long k) {
  G1Proj acc = g1_identity();
  bool started = false;
#pragma unroll 1
  for (int b = 63; b >= 0; b--) {
    if (started) acc = g1_dbl(acc);
    if ((k >> b) & 1) {
      acc = started ? g1_add(acc, p) : p;
      started = true;
    }
  }
  return acc;
}

// subgroup membership: phi(P) == -[x^2]P with phi(x,y) = (beta x, y)  (Scott, eprint 2021/1130;
// same accept set as [r]P == O)
DKGV_HD bool g1_in_subgroup(const G1Aff& a) {
  if (a.inf) return true;
  G1Proj p = g1_from_affine(a);
  G1Proj q = g1_mul_u64(g1_mul_u64(p, consts::X_ABS), consts::X_ABS);
  Fp beta;
#pragma unroll
  for (int i = 0; i < 12; i++) beta.l[i] = consts::BETA_M(i);
  G1Proj e;
  e.x = mul(a.x, beta);
  e.y = neg(a.y);
  e.z = one<FpParams>();
  return g1_eq(q, e);  // [x^2]P == -phi(P)
}

DKGV_HD G1Aff g1_to_affine(const G1Proj& p) {
  G1Aff r;
  Fp zi = fp_inv_bgcd(p.z);
  r.x = mul(p.x, zi);
  r.y = mul(p.y, zi);
  r.inf = is_zero(p.z) ? 1u : 0u;
  return r;
}

// decode status
enum : uint32_t { G1_DEC_OK = 0, G1_DEC_BAD_FLAGS = 1, G1_DEC_X_RANGE = 2, G1_DEC_NOT_ON_CURVE = 3, G1_DEC_NOT_IN_SUBGROUP = 4 };

// 48-byte compressed encoding -> affine Montgomery point
DKGV_HD uint32_t g1_decompress(const uint8_t* in, G1Aff* out, bool check_subgroup) {
  uint8_t b[48];
#pragma unroll
  for (int i = 0; i < 48; i++) b[i] = in[i];
  bool fc = (b[0] >> 7) & 1, fi = (b[0] >> 6) & 1, fs = (b[0] >> 5) & 1;
  b[0] &= 0x1f;
  Fp xr;
  fp_raw_from_be48(xr.l, b);
  out->x = zero<FpParams>();
  out->y = zero<FpParams>();
  out->inf = 1;
  if (!fc) return G1_DEC_BAD_FLAGS;
  if (!raw_lt_mod<FpParams>(xr.l)) return G1_DEC_X_RANGE;
  if (fi) {
    if (fs || !is_zero(xr)) return G1_DEC_BAD_FLAGS;
    return G1_DEC_OK;
  }
  Fp x = to_mont(xr);
  Fp four = dbl(dbl(one<FpParams>()));
  Fp rhs = add(mul(sqr(x), x), four);
  Fp y = fp_sqrt_candidate(rhs);
  if (!eq(sqr(y), rhs)) return G1_DEC_NOT_ON_CURVE;
  if (fp_lex_largest(y) != fs) y = neg(y);
  out->x = x;
  out->y = y;
  out->inf = 0;
  if (check_subgroup && !g1_in_subgroup(*out)) {
    out->inf = 1;
    return G1_DEC_NOT_IN_SUBGROUP;
  }
  return G1_DEC_OK;
}

DKGV_HD void g1_compress(const G1Aff& a, uint8_t* out) {
  if (a.inf) {
#pragma unroll
    for (int i = 0; i < 48; i++) out[i] = 0;
    out[0] = 0xc0;
    return;
  }
  Fp xr = from_mont(a.x);
  fp_raw_to_be48(out, xr.l);
  out[0] |= 0x80;
  if (fp_lex_largest(a.y)) out[0] |= 0x20;
}

// 32-byte big-endian scalar -> raw little-endian limbs; returns false when >= r
// (crates/dkg/src/crypto/bls_keys.rs:98-114)
DKGV_HD bool fr_raw_from_be32(uint32_t* l, const uint8_t* b) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = b + 28 - 4 * i;
    l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
  return raw_lt_mod<FrParams>(l);
}

}  // namespace dkgv
