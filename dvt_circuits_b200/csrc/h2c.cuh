// RFC 9380 hash_to_curve for G2, suite BLS12381G2_XMD:SHA-256_SSWU_RO_ with the reference's DST
// "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_" (crates/dkg/src/crypto/bls_common.rs:11-24), plus
// the SHA-256 it needs.  One thread per message (messages are few: one per finalization, one per
// bad-partial-key item); host/device code so tests/hostemu covers it.
#pragma once
#include "tower.cuh"

namespace dkgv {

struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len;
  uint32_t fill;
};
DKGV_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
DKGV_HD uint32_t sha_k(int i) {
  constexpr uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
      0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
      0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
      0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
      0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
      0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  return K[i];
}
DKGV_NI2 void sha_block(uint32_t* h, const uint8_t* p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = sha_rotr(w[i - 15], 7) ^ sha_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = sha_rotr(w[i - 2], 17) ^ sha_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = hh + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g)) + sha_k(i) + w[i];
    uint32_t t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
DKGV_HD void sha_init(Sha256* s) {
  s->h[0] = 0x6a09e667; s->h[1] = 0xbb67ae85; s->h[2] = 0x3c6ef372; s->h[3] = 0xa54ff53a;
  s->h[4] = 0x510e527f; s->h[5] = 0x9b05688c; s->h[6] = 0x1f83d9ab; s->h[7] = 0x5be0cd19;
  s->len = 0;
  s->fill = 0;
}
DKGV_NI2 void sha_update(Sha256* s, const uint8_t* p, size_t n) {
  s->len += n;
  for (size_t i = 0; i < n; i++) {
    s->buf[s->fill++] = p[i];
    if (s->fill == 64) {
      sha_block(s->h, s->buf);
      s->fill = 0;
    }
  }
}
DKGV_NI2 void sha_finish(Sha256* s, uint8_t* out) {
  uint64_t bits = s->len * 8;
  uint8_t pad = 0x80, z = 0;
  sha_update(s, &pad, 1);
  while (s->fill != 56) sha_update(s, &z, 1);
  uint8_t lb[8];
  for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
  sha_update(s, lb, 8);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(s->h[i] >> 24);
    out[4 * i + 1] = (uint8_t)(s->h[i] >> 16);
    out[4 * i + 2] = (uint8_t)(s->h[i] >> 8);
    out[4 * i + 3] = (uint8_t)s->h[i];
  }
}

constexpr int H2C_DST_LEN = 43;
DKGV_HD uint8_t h2c_dst(int i) {
  constexpr char D[] = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";
  return (uint8_t)D[i];
}

// expand_message_xmd(msg, DST, 256) (RFC 9380 5.3.1), DST <= 255 bytes
DKGV_NI2 void expand_message_xmd_256(const uint8_t* msg, size_t len, uint8_t* out /*256*/) {
  uint8_t dstp[H2C_DST_LEN + 1];
  for (int i = 0; i < H2C_DST_LEN; i++) dstp[i] = h2c_dst(i);
  dstp[H2C_DST_LEN] = (uint8_t)H2C_DST_LEN;
  uint8_t b0[32], bi[32], zpad[64];
  for (int i = 0; i < 64; i++) zpad[i] = 0;
  Sha256 s;
  sha_init(&s);
  sha_update(&s, zpad, 64);
  sha_update(&s, msg, len);
  uint8_t l2[3] = {1, 0, 0};  // l_i_b_str = 256 = 0x0100, then I2OSP(0, 1)
  sha_update(&s, l2, 3);
  sha_update(&s, dstp, H2C_DST_LEN + 1);
  sha_finish(&s, b0);
  sha_init(&s);
  sha_update(&s, b0, 32);
  uint8_t one = 1;
  sha_update(&s, &one, 1);
  sha_update(&s, dstp, H2C_DST_LEN + 1);
  sha_finish(&s, bi);
  for (int i = 1; i <= 8; i++) {
    for (int j = 0; j < 32; j++) out[32 * (i - 1) + j] = bi[j];
    if (i == 8) break;
    uint8_t x[32];
    for (int j = 0; j < 32; j++) x[j] = b0[j] ^ bi[j];
    sha_init(&s);
    sha_update(&s, x, 32);
    uint8_t idx = (uint8_t)(i + 1);
    sha_update(&s, &idx, 1);
    sha_update(&s, dstp, H2C_DST_LEN + 1);
    sha_finish(&s, bi);
  }
}

// 64 big-endian bytes -> Fp (mod p), Montgomery form:  hi * 2^256 + lo
DKGV_NI2 void fp_from_be64_reduce(Fp* r, const uint8_t* b) {
  Fp hi = zero<FpParams>(), lo = zero<FpParams>();
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = b + 28 - 4 * i;
    hi.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    const uint8_t* q2 = b + 60 - 4 * i;
    lo.l[i] = ((uint32_t)q2[0] << 24) | ((uint32_t)q2[1] << 16) | ((uint32_t)q2[2] << 8) | (uint32_t)q2[3];
  }
  hi = to_mont(hi);
  lo = to_mont(lo);
  Fp k = fp_const<consts::TWO256_M>();
  *r = add(mul(hi, k), lo);
}

// simplified SWU onto E2' (RFC 9380 6.6.2, straight-line form with the exceptional case)
DKGV_NI2 void sswu_g2(G2Aff* r, const Fp2* u) {
  Fp2 Z = fp2_const<consts::SSWU_Z_C0, consts::SSWU_Z_C1>(), A = fp2_const<consts::SSWU_A_C0, consts::SSWU_A_C1>(),
      Bc = fp2_const<consts::SSWU_B_C0, consts::SSWU_B_C1>();
  Fp2 zu2, tv1, x1, gx, t, y, one = fp2_one();
  fp2_sqr(&zu2, u);
  fp2_mul(&zu2, &zu2, &Z);
  fp2_sqr(&tv1, &zu2);
  fp2_add(&tv1, &tv1, &zu2);
  if (fp2_is_zero(tv1)) {
    x1 = fp2_const<consts::SSWU_B_OVER_ZA_C0, consts::SSWU_B_OVER_ZA_C1>();
  } else {
    Fp2 k = fp2_const<consts::SSWU_NEG_B_OVER_A_C0, consts::SSWU_NEG_B_OVER_A_C1>();
    fp2_inv(&t, &tv1);
    fp2_add(&t, &t, &one);
    fp2_mul(&x1, &k, &t);
  }
  fp2_sqr(&gx, &x1);
  fp2_add(&gx, &gx, &A);
  fp2_mul(&gx, &gx, &x1);
  fp2_add(&gx, &gx, &Bc);
  Fp2 x = x1;
  if (fp2_is_square(&gx)) {
    fp2_sqrt(&y, &gx);
  } else {
    fp2_mul(&x, &zu2, &x1);
    fp2_sqr(&gx, &x);
    fp2_add(&gx, &gx, &A);
    fp2_mul(&gx, &gx, &x);
    fp2_add(&gx, &gx, &Bc);
    fp2_sqrt(&y, &gx);
  }
  if (fp2_sgn0(*u) != fp2_sgn0(y)) fp2_neg(&y, &y);
  r->x = x;
  r->y = y;
  r->inf = 0;
}

#define DKGV_ISO_COEF(NAME, K) fp2_const<consts::NAME##K##_C0, consts::NAME##K##_C1>()
// 3-isogeny E2' -> E2 (RFC 9380 appendix E.3)
DKGV_NI2 void iso3_g2(G2Aff* r, const G2Aff* p) {
  Fp2 xn, xd, yn, yd, c;
  // Horner, high -> low
  xn = DKGV_ISO_COEF(ISO_XNUM, 3);
  c = DKGV_ISO_COEF(ISO_XNUM, 2); fp2_mul(&xn, &xn, &p->x); fp2_add(&xn, &xn, &c);
  c = DKGV_ISO_COEF(ISO_XNUM, 1); fp2_mul(&xn, &xn, &p->x); fp2_add(&xn, &xn, &c);
  c = DKGV_ISO_COEF(ISO_XNUM, 0); fp2_mul(&xn, &xn, &p->x); fp2_add(&xn, &xn, &c);
  xd = DKGV_ISO_COEF(ISO_XDEN, 2);
  c = DKGV_ISO_COEF(ISO_XDEN, 1); fp2_mul(&xd, &xd, &p->x); fp2_add(&xd, &xd, &c);
  c = DKGV_ISO_COEF(ISO_XDEN, 0); fp2_mul(&xd, &xd, &p->x); fp2_add(&xd, &xd, &c);
  yn = DKGV_ISO_COEF(ISO_YNUM, 3);
  c = DKGV_ISO_COEF(ISO_YNUM, 2); fp2_mul(&yn, &yn, &p->x); fp2_add(&yn, &yn, &c);
  c = DKGV_ISO_COEF(ISO_YNUM, 1); fp2_mul(&yn, &yn, &p->x); fp2_add(&yn, &yn, &c);
  c = DKGV_ISO_COEF(ISO_YNUM, 0); fp2_mul(&yn, &yn, &p->x); fp2_add(&yn, &yn, &c);
  yd = DKGV_ISO_COEF(ISO_YDEN, 3);
  c = DKGV_ISO_COEF(ISO_YDEN, 2); fp2_mul(&yd, &yd, &p->x); fp2_add(&yd, &yd, &c);
  c = DKGV_ISO_COEF(ISO_YDEN, 1); fp2_mul(&yd, &yd, &p->x); fp2_add(&yd, &yd, &c);
  c = DKGV_ISO_COEF(ISO_YDEN, 0); fp2_mul(&yd, &yd, &p->x); fp2_add(&yd, &yd, &c);
  if (fp2_is_zero(xd) || fp2_is_zero(yd)) {
    r->x = fp2_zero();
    r->y = fp2_one();
    r->inf = 1;
    return;
  }
  fp2_inv(&xd, &xd);
  fp2_inv(&yd, &yd);
  fp2_mul(&r->x, &xn, &xd);
  fp2_mul(&yn, &yn, &yd);
  fp2_mul(&r->y, &p->y, &yn);
  r->inf = 0;
}

// hash_to_curve(msg) on G2 with the reference's DST; result affine
DKGV_NI2 void hash_to_g2(G2Aff* out, const uint8_t* msg, size_t len) {
  uint8_t uni[256];
  expand_message_xmd_256(msg, len, uni);
  Fp2 u0, u1;
  fp_from_be64_reduce(&u0.c0, uni);
  fp_from_be64_reduce(&u0.c1, uni + 64);
  fp_from_be64_reduce(&u1.c0, uni + 128);
  fp_from_be64_reduce(&u1.c1, uni + 192);
  G2Aff q0, q1;
  sswu_g2(&q0, &u0);
  iso3_g2(&q0, &q0);
  sswu_g2(&q1, &u1);
  iso3_g2(&q1, &q1);
  G2Proj a = g2_from_affine(q0), b = g2_from_affine(q1), s;
  g2_add(&s, &a, &b);
  g2_mul_public(&a, &s, LimbHEff(), 20);
  g2_to_affine(out, &a);
}

}  // namespace dkgv
