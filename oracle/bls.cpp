// CPU ORACLE - TEST INFRASTRUCTURE ONLY (see bls.hpp).
#include "bls.hpp"

#include <mutex>

namespace orc {

struct Consts {
  u64 p_plus1_div4[6], p_minus1_div2[6], r_limbs[4];
  Fp2 frob6_c1, frob6_c2, frob12_c1;  // xi^((p-1)/3), xi^(2(p-1)/3), xi^((p-1)/6)
  G1Aff g1;
  Fp2 sswu_a, sswu_b, sswu_z, sswu_neg_b_over_a, sswu_b_over_za;
  Fp2 iso_xnum[4], iso_xden[3], iso_ynum[4], iso_yden[4];
  std::vector<u64> h_eff;
};
static Consts g_k;
static std::once_flag g_once;
const Consts& K() { return g_k; }
const char* DST_POP = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";

static void div_small(u64* q, const u64* a, int n, u64 d) {
  u128 rem = 0;
  for (int i = n - 1; i >= 0; i--) {
    u128 cur = (rem << 64) | a[i];
    q[i] = (u64)(cur / d);
    rem = cur % d;
  }
}
static Fp2 fp2_hex(const char* a, const char* b) { return Fp2{Fp::from_hex_str(a), Fp::from_hex_str(b)}; }

void init() {
  std::call_once(g_once, [] {
    Fp::init(P_HEX);
    Fr::init(R_HEX);
    u64 one6[6] = {1}, t[6];
    add_n<6>(t, Fp::MOD, one6);
    div_small(g_k.p_plus1_div4, t, 6, 4);
    sub_n<6>(t, Fp::MOD, one6);
    div_small(g_k.p_minus1_div2, t, 6, 2);
    memcpy(g_k.r_limbs, Fr::MOD, sizeof g_k.r_limbs);
    u64 e[6];
    Fp2 xi{Fp::one(), Fp::one()};
    div_small(e, t, 6, 6);  // (p-1)/6
    g_k.frob12_c1 = xi.pow(e, 6);
    div_small(e, t, 6, 3);  // (p-1)/3
    g_k.frob6_c1 = xi.pow(e, 6);
    g_k.frob6_c2 = g_k.frob6_c1.sqr();
    g_k.g1 = {Fp::from_hex_str("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"),
              Fp::from_hex_str("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1"), false};
    // RFC 9380 8.8.2: E2' : y^2 = x^3 + 240 u x + 1012 (1 + u),  Z = -(2 + u)
    g_k.sswu_a = {Fp::zero(), Fp::from_u64(240)};
    g_k.sswu_b = {Fp::from_u64(1012), Fp::from_u64(1012)};
    g_k.sswu_z = {-Fp::from_u64(2), -Fp::from_u64(1)};
    g_k.sswu_neg_b_over_a = (-g_k.sswu_b) * g_k.sswu_a.inv();
    g_k.sswu_b_over_za = g_k.sswu_b * (g_k.sswu_z * g_k.sswu_a).inv();
    // RFC 9380 appendix E.3 3-isogeny coefficients (low -> high degree)
    const char* Z = "0";
    g_k.iso_xnum[0] = fp2_hex("5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6",
                              "5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6");
    g_k.iso_xnum[1] = fp2_hex(Z, "11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71a");
    g_k.iso_xnum[2] = fp2_hex("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71e",
                              "8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38d");
    g_k.iso_xnum[3] = fp2_hex("171d6541fa38ccfaed6dea691f5fb614cb14b4e7f4e810aa22d6108f142b85757098e38d0f671c7188e2aaaaaaaa5ed1", Z);
    g_k.iso_xden[0] = fp2_hex(Z, "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa63");
    g_k.iso_xden[1] = fp2_hex("c", "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa9f");
    g_k.iso_xden[2] = fp2_hex("1", Z);
    g_k.iso_ynum[0] = fp2_hex("1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706",
                              "1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706");
    g_k.iso_ynum[1] = fp2_hex(Z, "5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97be");
    g_k.iso_ynum[2] = fp2_hex("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71c",
                              "8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38f");
    g_k.iso_ynum[3] = fp2_hex("124c9ad43b6cf79bfbf7043de3811ad0761b0f37a1e26286b0e977c69aa274524e79097a56dc4bd9e1b371c71c718b10", Z);
    g_k.iso_yden[0] = fp2_hex("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa8fb",
                              "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa8fb");
    g_k.iso_yden[1] = fp2_hex(Z, "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa9d3");
    g_k.iso_yden[2] = fp2_hex("12", "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa99");
    g_k.iso_yden[3] = fp2_hex("1", Z);
    const char* heff =
        "bc69f08f2ee75b3584c6a0ea91b352888e2a8e9145ad7689986ff031508ffe1329c2f178731db956d82bf015d1212b02ec0ec69d7477c1ae954cbc06689f6a359894c0adebbf6b4e8020005aaa95551";
    g_k.h_eff.assign(10, 0);
    from_hex<10>(g_k.h_eff.data(), heff);
  });
}

Fp fp_sqrt_cand(const Fp& a) { return a.pow(g_k.p_plus1_div4, 6); }
bool fp_is_square(const Fp& a) { return a.is_zero() || a.pow(g_k.p_minus1_div2, 6) == Fp::one(); }
bool fp_lex_largest(const Fp& a) {
  u64 raw[6];
  a.to_raw(raw);
  return cmp<6>(raw, g_k.p_minus1_div2) > 0;
}

bool Fp2::sqrt(Fp2* out) const {
  // complex method: sqrt(a0 + a1 u) via alpha = sqrt(a0^2 + a1^2) in Fp
  if (is_zero()) {
    *out = zero();
    return true;
  }
  Fp inv2 = Fp::from_u64(2).inv();
  if (c1.is_zero()) {
    Fp s = fp_sqrt_cand(c0);
    if (s.sqr() == c0) {
      *out = {s, Fp::zero()};
      return true;
    }
    s = fp_sqrt_cand(-c0);  // (s u)^2 = -s^2
    *out = {Fp::zero(), s};
    return out->sqr() == *this;
  }
  Fp n = c0.sqr() + c1.sqr();
  Fp alpha = fp_sqrt_cand(n);
  if (alpha.sqr() != n) return false;
  Fp delta = (c0 + alpha) * inv2;
  Fp x0 = fp_sqrt_cand(delta);
  if (x0.sqr() != delta) {
    delta = (c0 - alpha) * inv2;
    x0 = fp_sqrt_cand(delta);
    if (x0.sqr() != delta) return false;
  }
  Fp x1 = c1 * x0.dbl().inv();
  *out = {x0, x1};
  return out->sqr() == *this;
}

Fp6 Fp6::frob() const { return {c0.conj(), c1.conj() * g_k.frob6_c1, c2.conj() * g_k.frob6_c2}; }
Fp12 Fp12::frob() const {
  Fp6 a = c0.frob(), b = c1.frob();
  return {a, b.scale(g_k.frob12_c1)};
}

// ------------------------------------------------------------------------------------ curves
G1Aff g1_generator() { return g_k.g1; }
bool g1_on_curve(const G1Aff& a) { return a.inf || a.y.sqr() == a.x.sqr() * a.x + B3<Fp>::b(); }
bool g1_in_subgroup(const G1Aff& a) { return a.inf || G1::from_affine(a).mul_vartime(g_k.r_limbs, 4).is_identity(); }
bool g2_in_subgroup(const G2Aff& a) { return a.inf || G2::from_affine(a).mul_vartime(g_k.r_limbs, 4).is_identity(); }

int g1_decompress(const uint8_t* in, G1Aff* out) {
  uint8_t b[48];
  memcpy(b, in, 48);
  bool fc = b[0] >> 7 & 1, fi = b[0] >> 6 & 1, fs = b[0] >> 5 & 1;
  b[0] &= 0x1f;
  *out = {Fp::zero(), Fp::one(), true};
  if (!fc) return DEC_BAD_FLAGS;
  Fp x;
  if (!Fp::from_be(&x, b, 48)) return DEC_X_RANGE;
  if (fi) return (fs || !x.is_zero()) ? DEC_BAD_FLAGS : DEC_OK;
  Fp rhs = x.sqr() * x + B3<Fp>::b();
  Fp y = fp_sqrt_cand(rhs);
  if (y.sqr() != rhs) return DEC_NOT_ON_CURVE;
  if (fp_lex_largest(y) != fs) y = -y;
  G1Aff p{x, y, false};
  if (!g1_in_subgroup(p)) return DEC_NOT_IN_SUBGROUP;
  *out = p;
  return DEC_OK;
}
void g1_compress(const G1Aff& a, uint8_t* out) {
  if (a.inf) {
    memset(out, 0, 48);
    out[0] = 0xc0;
    return;
  }
  a.x.to_be(out, 48);
  out[0] |= 0x80;
  if (fp_lex_largest(a.y)) out[0] |= 0x20;
}
int g2_decompress(const uint8_t* in, G2Aff* out) {
  uint8_t b[96];
  memcpy(b, in, 96);
  bool fc = b[0] >> 7 & 1, fi = b[0] >> 6 & 1, fs = b[0] >> 5 & 1;
  b[0] &= 0x1f;
  *out = {Fp2::zero(), Fp2::one(), true};
  if (!fc) return DEC_BAD_FLAGS;
  Fp2 x;
  if (!Fp::from_be(&x.c1, b, 48) || !Fp::from_be(&x.c0, b + 48, 48)) return DEC_X_RANGE;
  if (fi) return (fs || !x.is_zero()) ? DEC_BAD_FLAGS : DEC_OK;
  Fp2 rhs = x.sqr() * x + B3<Fp2>::b();
  Fp2 y;
  if (!rhs.sqrt(&y)) return DEC_NOT_ON_CURVE;
  if (y.lex_largest() != fs) y = -y;
  G2Aff p{x, y, false};
  if (!g2_in_subgroup(p)) return DEC_NOT_IN_SUBGROUP;
  *out = p;
  return DEC_OK;
}
void g2_compress(const G2Aff& a, uint8_t* out) {
  if (a.inf) {
    memset(out, 0, 96);
    out[0] = 0xc0;
    return;
  }
  a.x.c1.to_be(out, 48);
  a.x.c0.to_be(out + 48, 48);
  out[0] |= 0x80;
  if (a.y.lex_largest()) out[0] |= 0x20;
}

// ------------------------------------------------------------------------------------ pairing
// Line through the twist point(s), evaluated at P = (xP, yP) and scaled by Fp2/Fp4 factors that the
// final exponentiation kills:  l = c00 + c01 * v + (c11 * v) * w   with
//   tangent at T = (X:Y:Z):   c00 = Y^2 - 3 b' Z^2,  c01 = -3 X^2 * xP,  c11 = 2 Y Z * yP
//   chord through T and Q:    N = yQ Z - Y, D = xQ Z - X:  c00 = N xQ - D yQ, c01 = -N xP, c11 = D yP
Fp12 miller_loop(const G1Aff& p, const G2Aff& q) {
  if (p.inf || q.inf) return Fp12::one();
  Fp12 f = Fp12::one();
  G2 t = G2::from_affine(q);
  for (int b = 62; b >= 0; b--) {
    f = f.sqr();
    {
      Fp2 c00 = t.y.sqr() - B3<Fp2>::mul(t.z.sqr());
      Fp2 x2 = t.x.sqr();
      Fp2 c01 = -(x2.dbl() + x2).scale(p.x);
      Fp2 c11 = (t.y * t.z).dbl().scale(p.y);
      f = f.mul_by_014(c00, c01, c11);
      t = t.dbl();
    }
    if ((X_ABS >> b) & 1) {
      Fp2 n = q.y * t.z - t.y, d = q.x * t.z - t.x;
      Fp2 c00 = n * q.x - d * q.y;
      Fp2 c01 = -n.scale(p.x);
      Fp2 c11 = d.scale(p.y);
      f = f.mul_by_014(c00, c01, c11);
      t = t.add_mixed(q);
    }
  }
  return f.conj();  // x < 0
}

static Fp12 pow_x(const Fp12& a) { return a.pow_x_abs().conj(); }  // a^x for unitary a (x < 0)

Fp12 final_exponentiation(const Fp12& f0) {
  // easy part: f^((p^6 - 1)(p^2 + 1))
  Fp12 f = f0.conj() * f0.inv();
  f = f.frob().frob() * f;
  // hard part times 3: f^((x-1)^2 (x+p) (x^2+p^2-1) + 3); inverse of a unitary element = conjugate
  Fp12 t0 = pow_x(f) * f.conj();      // f^(x-1)
  Fp12 t1 = pow_x(t0) * t0.conj();    // ^(x-1)
  Fp12 t2 = pow_x(t1) * t1.frob();    // ^(x+p)
  Fp12 t3 = pow_x(pow_x(t2)) * t2.frob().frob() * t2.conj();  // ^(x^2+p^2-1)
  return t3 * f.sqr() * f;
}

Fp12 pairing(const G1Aff& p, const G2Aff& q) {
  if (p.inf || q.inf) return Fp12::one();
  return final_exponentiation(miller_loop(p, q));
}

// ------------------------------------------------------------------------------------ SHA-256 (FIPS 180-4)
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_block(uint32_t* h, const uint8_t* p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
    uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
Sha256::Sha256() : len(0), fill(0) {
  static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  memcpy(h, iv, sizeof h);
}
void Sha256::update(const uint8_t* p, size_t n) {
  len += n;
  while (n) {
    size_t k = 64 - fill < n ? 64 - fill : n;
    memcpy(buf + fill, p, k);
    fill += k;
    p += k;
    n -= k;
    if (fill == 64) {
      sha_block(h, buf);
      fill = 0;
    }
  }
}
void Sha256::finish(uint8_t out[32]) {
  uint64_t bits = len * 8;
  uint8_t pad = 0x80;
  update(&pad, 1);
  uint8_t z = 0;
  while (fill != 56) update(&z, 1);
  uint8_t lb[8];
  for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
  update(lb, 8);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i];
  }
}
void sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  Sha256 s;
  s.update(msg, len);
  s.finish(out);
}

// ------------------------------------------------------------------------------------ RFC 9380 hash_to_curve (G2, XMD:SHA-256, SSWU, RO)
static void expand_message_xmd(const uint8_t* msg, size_t len, const uint8_t* dst, size_t dst_len, uint8_t* out, size_t n) {
  uint8_t dstp[256 + 1];
  uint8_t hashed[32];
  if (dst_len > 255) {
    Sha256 s;
    s.update((const uint8_t*)"H2C-OVERSIZE-DST-", 17);
    s.update(dst, dst_len);
    s.finish(hashed);
    dst = hashed;
    dst_len = 32;
  }
  memcpy(dstp, dst, dst_len);
  dstp[dst_len] = (uint8_t)dst_len;
  size_t ell = (n + 31) / 32;
  uint8_t b0[32], bi[32];
  {
    Sha256 s;
    uint8_t zpad[64] = {0};
    s.update(zpad, 64);
    s.update(msg, len);
    uint8_t l2[3] = {(uint8_t)(n >> 8), (uint8_t)n, 0};
    s.update(l2, 3);
    s.update(dstp, dst_len + 1);
    s.finish(b0);
  }
  {
    Sha256 s;
    s.update(b0, 32);
    uint8_t one = 1;
    s.update(&one, 1);
    s.update(dstp, dst_len + 1);
    s.finish(bi);
  }
  size_t off = 0;
  for (size_t i = 1; i <= ell; i++) {
    size_t k = n - off < 32 ? n - off : 32;
    memcpy(out + off, bi, k);
    off += k;
    if (i == ell) break;
    uint8_t x[32];
    for (int j = 0; j < 32; j++) x[j] = b0[j] ^ bi[j];
    Sha256 s;
    s.update(x, 32);
    uint8_t idx = (uint8_t)(i + 1);
    s.update(&idx, 1);
    s.update(dstp, dst_len + 1);
    s.finish(bi);
  }
}

static Fp fp_from_be64_reduce(const uint8_t* b) {  // 64 big-endian bytes mod p
  // value = hi * 2^256 + lo with hi, lo of 32 bytes each; 2^256 < p so both halves are canonical
  u64 hi[6] = {0}, lo[6] = {0};
  for (int i = 0; i < 32; i++) {
    hi[i / 8] |= (u64)b[31 - i] << (8 * (i % 8));
    lo[i / 8] |= (u64)b[63 - i] << (8 * (i % 8));
  }
  u64 two256[6] = {0, 0, 0, 0, 1, 0};
  return Fp::from_raw(hi) * Fp::from_raw(two256) + Fp::from_raw(lo);
}

static G2Aff sswu_g2(const Fp2& u) {
  const Consts& k = g_k;
  Fp2 zu2 = k.sswu_z * u.sqr();
  Fp2 tv1 = zu2.sqr() + zu2;
  Fp2 x1 = tv1.is_zero() ? k.sswu_b_over_za : k.sswu_neg_b_over_a * (Fp2::one() + tv1.inv());
  Fp2 gx1 = (x1.sqr() + k.sswu_a) * x1 + k.sswu_b;
  Fp2 x, y;
  if (gx1.is_square()) {
    x = x1;
    gx1.sqrt(&y);
  } else {
    x = zu2 * x1;
    Fp2 gx2 = (x.sqr() + k.sswu_a) * x + k.sswu_b;
    gx2.sqrt(&y);
  }
  if (u.sgn0() != y.sgn0()) y = -y;
  return {x, y, false};
}
static Fp2 horner(const Fp2* c, int n, const Fp2& x) {
  Fp2 acc = c[n - 1];
  for (int i = n - 2; i >= 0; i--) acc = acc * x + c[i];
  return acc;
}
static G2Aff iso3(const G2Aff& p) {
  const Consts& k = g_k;
  Fp2 xn = horner(k.iso_xnum, 4, p.x), xd = horner(k.iso_xden, 3, p.x);
  Fp2 yn = horner(k.iso_ynum, 4, p.x), yd = horner(k.iso_yden, 4, p.x);
  if (xd.is_zero() || yd.is_zero()) return {Fp2::zero(), Fp2::one(), true};
  return {xn * xd.inv(), p.y * yn * yd.inv(), false};
}

G2Aff hash_to_g2(const uint8_t* msg, size_t len, const uint8_t* dst, size_t dst_len) {
  uint8_t uni[256];
  expand_message_xmd(msg, len, dst, dst_len, uni, 256);
  Fp2 u0{fp_from_be64_reduce(uni), fp_from_be64_reduce(uni + 64)};
  Fp2 u1{fp_from_be64_reduce(uni + 128), fp_from_be64_reduce(uni + 192)};
  G2 q = G2::from_affine(iso3(sswu_g2(u0))).add(G2::from_affine(iso3(sswu_g2(u1))));
  return q.mul_vartime(g_k.h_eff.data(), 10).to_affine();
}

}  // namespace orc
