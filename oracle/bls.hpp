// CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Nothing under dvt_circuits_b200/ may include, link or
// execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it, and only as the checker / reported baseline.
//
// C++17 restatement (64-bit limbs, unsigned __int128) of the BLS12-381 arithmetic that the
// reference's hot path calls in the un-vendored `bls12_381` crate (sp1-patches fork of
// zkcrypto/bls12_381 0.8.0, crates/dkg/Cargo.toml:25; call sites crates/dkg/src/dkg_math.rs:43-128,
// crates/dkg/src/crypto/bls_common.rs:11-47,108-112, crates/dkg/src/crypto/bls_keys.rs:15-40,
// 98-137,165-190).  The algorithms are restated from their published descriptions:
//   Montgomery CIOS; Renes-Costello-Batina complete formulas (eprint 2015/1060, Alg. 7-9);
//   ZCash compressed encodings; RFC 9380 (expand_message_xmd, simplified SWU, 3-isogeny, h_eff);
//   optimal-ate Miller loop with homogeneous twist arithmetic; final exponentiation through
//   3*(p^4-p^2+1)/r = (x-1)^2 (x+p)(x^2+p^2-1) + 3.
// Parity pin: the reference's own KATs (crates/dkg/src/dkg_math.rs:259-375), all 67 non-encrypted
// test_vectors/ outcomes and the examples/ (via tests/golden/vector_outcomes.json) and the
// independent Python restatement oracle/pyref.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace orc {
typedef unsigned __int128 u128;
typedef uint64_t u64;

// ------------------------------------------------------------------------------------ big helpers
template <int N>
struct Limbs {
  u64 v[N];
};

template <int N>
static inline int cmp(const u64* a, const u64* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  }
  return 0;
}
template <int N>
static inline u64 add_n(u64* r, const u64* a, const u64* b) {
  u128 c = 0;
  for (int i = 0; i < N; i++) {
    c += (u128)a[i] + b[i];
    r[i] = (u64)c;
    c >>= 64;
  }
  return (u64)c;
}
template <int N>
static inline u64 sub_n(u64* r, const u64* a, const u64* b) {
  u64 br = 0;
  for (int i = 0; i < N; i++) {
    u128 d = (u128)a[i] - b[i] - br;
    r[i] = (u64)d;
    br = (u64)(d >> 64) & 1;
  }
  return br;
}
static inline int hexval(char c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; }
template <int N>
static inline void from_hex(u64* r, const char* h) {
  for (int i = 0; i < N; i++) r[i] = 0;
  size_t n = strlen(h);
  for (size_t i = 0; i < n; i++) {
    int d = hexval(h[n - 1 - i]);
    r[i / 16] |= (u64)d << (4 * (i % 16));
  }
}

// ------------------------------------------------------------------------------------ Montgomery field
template <int N, class Tag>
struct Mont {
  u64 v[N];
  static u64 MOD[N], ONE[N], R2[N], INV;
  static void init(const char* modhex) {
    from_hex<N>(MOD, modhex);
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - MOD[0] * inv;  // Newton: mod^-1 mod 2^64
    INV = (u64)0 - inv;
    u64 t[N] = {1};
    for (int i = 1; i < N; i++) t[i] = 0;
    for (int i = 0; i < 128 * N; i++) {  // t = 2^(128N) mod p by doubling
      u64 c = add_n<N>(t, t, t);
      if (c || cmp<N>(t, MOD) >= 0) sub_n<N>(t, t, MOD);
      if (i == 64 * N - 1) memcpy(ONE, t, sizeof t);
    }
    memcpy(R2, t, sizeof t);
  }
  static Mont zero() {
    Mont r;
    memset(r.v, 0, sizeof r.v);
    return r;
  }
  static Mont one() {
    Mont r;
    memcpy(r.v, ONE, sizeof r.v);
    return r;
  }
  bool is_zero() const {
    u64 x = 0;
    for (int i = 0; i < N; i++) x |= v[i];
    return x == 0;
  }
  bool operator==(const Mont& o) const { return memcmp(v, o.v, sizeof v) == 0; }
  bool operator!=(const Mont& o) const { return !(*this == o); }
  Mont operator+(const Mont& o) const {
    Mont r;
    u64 c = add_n<N>(r.v, v, o.v);
    if (c || cmp<N>(r.v, MOD) >= 0) sub_n<N>(r.v, r.v, MOD);
    return r;
  }
  Mont operator-(const Mont& o) const {
    Mont r;
    if (sub_n<N>(r.v, v, o.v)) add_n<N>(r.v, r.v, MOD);
    return r;
  }
  Mont operator-() const { return zero() - *this; }
  Mont dbl() const { return *this + *this; }
  static thread_local u64 MULS;  // multiplication counter (algorithmic work)
  Mont operator*(const Mont& o) const {
    // CIOS with the two inner loops fused ("no-carry" variant: valid because both moduli leave the
    // top bit of the top limb clear, so the running sum never needs an (N+1)-th limb)
    MULS++;
    u64 t[N];
#pragma GCC unroll 8
    for (int j = 0; j < N; j++) t[j] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) {
      u128 a = (u128)v[0] * o.v[i] + t[0];
      u64 m = (u64)a * INV;
      u128 c = (u128)m * MOD[0] + (u64)a;
#pragma GCC unroll 8
      for (int j = 1; j < N; j++) {
        a = (u128)v[j] * o.v[i] + t[j] + (u64)(a >> 64);
        c = (u128)m * MOD[j] + (u64)a + (u64)(c >> 64);
        t[j - 1] = (u64)c;
      }
      t[N - 1] = (u64)(c >> 64) + (u64)(a >> 64);
    }
    Mont r;
    if (cmp<N>(t, MOD) >= 0) sub_n<N>(r.v, t, MOD);
    else memcpy(r.v, t, sizeof r.v);
    return r;
  }
  Mont sqr() const { return *this * *this; }
  // canonical little-endian limbs <-> Montgomery
  static Mont from_raw(const u64* raw) {
    Mont a, r2;
    memcpy(a.v, raw, sizeof a.v);
    memcpy(r2.v, R2, sizeof r2.v);
    return a * r2;
  }
  void to_raw(u64* raw) const {
    Mont o = zero();
    o.v[0] = 1;
    Mont c = *this * o;
    memcpy(raw, c.v, sizeof c.v);
  }
  static Mont from_u64(u64 x) {
    u64 raw[N] = {x};
    return from_raw(raw);
  }
  static Mont from_hex_str(const char* h) {
    u64 raw[N];
    from_hex<N>(raw, h);
    return from_raw(raw);
  }
  // exponent = little-endian limbs
  Mont pow(const u64* e, int n) const {
    Mont r = one();
    bool started = false;
    for (int i = n - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        if (started) r = r.sqr();
        if ((e[i] >> b) & 1) {
          r = started ? r * *this : *this;
          started = true;
        }
      }
    return r;
  }
  Mont inv() const {  // a^(p-2); 0 -> 0
    u64 e[N], two[N] = {2};
    sub_n<N>(e, MOD, two);
    return pow(e, N);
  }
  // big-endian byte codec of the canonical value; false when >= modulus
  static bool from_be(Mont* out, const uint8_t* b, int nbytes) {
    u64 raw[N];
    memset(raw, 0, sizeof raw);
    for (int i = 0; i < nbytes; i++) raw[i / 8] |= (u64)b[nbytes - 1 - i] << (8 * (i % 8));
    if (cmp<N>(raw, MOD) >= 0) return false;
    *out = from_raw(raw);
    return true;
  }
  void to_be(uint8_t* b, int nbytes) const {
    u64 raw[N];
    to_raw(raw);
    for (int i = 0; i < nbytes; i++) b[nbytes - 1 - i] = (uint8_t)(raw[i / 8] >> (8 * (i % 8)));
  }
};
template <int N, class Tag> u64 Mont<N, Tag>::MOD[N];
template <int N, class Tag> u64 Mont<N, Tag>::ONE[N];
template <int N, class Tag> u64 Mont<N, Tag>::R2[N];
template <int N, class Tag> u64 Mont<N, Tag>::INV;
template <int N, class Tag> thread_local u64 Mont<N, Tag>::MULS;

struct FpTag {};
struct FrTag {};
typedef Mont<6, FpTag> Fp;
typedef Mont<4, FrTag> Fr;

[[maybe_unused]] static const char* P_HEX = "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab";
[[maybe_unused]] static const char* R_HEX = "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001";
static const u64 X_ABS = 0xd201000000010000ull;  // |x|, x < 0

struct Consts;
const Consts& K();

// Fp helpers needing p-derived exponents
Fp fp_sqrt_cand(const Fp& a);   // a^((p+1)/4)
bool fp_is_square(const Fp& a);  // Euler
bool fp_lex_largest(const Fp& a);  // canonical(a) > (p-1)/2

// ------------------------------------------------------------------------------------ Fp2 = Fp[u]/(u^2+1)
struct Fp2 {
  Fp c0, c1;
  static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
  static Fp2 one() { return {Fp::one(), Fp::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fp2& o) const { return !(*this == o); }
  Fp2 operator+(const Fp2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fp2 operator-(const Fp2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fp2 operator-() const { return {-c0, -c1}; }
  Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fp2 operator*(const Fp2& o) const {  // Karatsuba, 3 M
    Fp a = c0 * o.c0, b = c1 * o.c1;
    Fp c = (c0 + c1) * (o.c0 + o.c1);
    return {a - b, c - a - b};
  }
  Fp2 sqr() const {  // (c0+c1)(c0-c1), 2 c0 c1
    Fp a = (c0 + c1) * (c0 - c1), b = c0 * c1;
    return {a, b.dbl()};
  }
  Fp2 scale(const Fp& s) const { return {c0 * s, c1 * s}; }
  Fp2 conj() const { return {c0, -c1}; }
  Fp2 mul_xi() const { return {c0 - c1, c0 + c1}; }  // * (1 + u)
  Fp2 inv() const {
    Fp n = (c0.sqr() + c1.sqr()).inv();
    return {c0 * n, -(c1 * n)};
  }
  Fp2 pow(const u64* e, int n) const {
    Fp2 r = one();
    for (int i = n - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        r = r.sqr();
        if ((e[i] >> b) & 1) r = r * *this;
      }
    return r;
  }
  bool is_square() const {
    Fp n = c0.sqr() + c1.sqr();
    return n.is_zero() || fp_is_square(n);
  }
  bool sqrt(Fp2* out) const;
  bool lex_largest() const { return c1.is_zero() ? fp_lex_largest(c0) : fp_lex_largest(c1); }
  int sgn0() const {
    u64 r0[6], r1[6];
    c0.to_raw(r0);
    c1.to_raw(r1);
    return (int)((r0[0] & 1) | ((u64)c0.is_zero() & (r1[0] & 1)));
  }
};

// ------------------------------------------------------------------------------------ Fp6 = Fp2[v]/(v^3 - xi)
struct Fp6 {
  Fp2 c0, c1, c2;
  static Fp6 zero() { return {Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
  static Fp6 one() { return {Fp2::one(), Fp2::zero(), Fp2::zero()}; }
  bool operator==(const Fp6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  Fp6 operator+(const Fp6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  Fp6 operator-(const Fp6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  Fp6 operator-() const { return {-c0, -c1, -c2}; }
  Fp6 operator*(const Fp6& o) const {
    Fp2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
    Fp2 r0 = ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_xi() + t0;
    Fp2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_xi();
    Fp2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  Fp6 sqr() const { return *this * *this; }
  Fp6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
  Fp6 scale(const Fp2& s) const { return {c0 * s, c1 * s, c2 * s}; }
  Fp6 inv() const {
    Fp2 t0 = c0.sqr() - (c1 * c2).mul_xi();
    Fp2 t1 = c2.sqr().mul_xi() - c0 * c1;
    Fp2 t2 = c1.sqr() - c0 * c2;
    Fp2 d = (c0 * t0 + (c2 * t1 + c1 * t2).mul_xi()).inv();
    return {t0 * d, t1 * d, t2 * d};
  }
  Fp6 frob() const;  // x -> x^p
};

// ------------------------------------------------------------------------------------ Fp12 = Fp6[w]/(w^2 - v)
struct Fp12 {
  Fp6 c0, c1;
  static Fp12 one() { return {Fp6::one(), Fp6::zero()}; }
  bool operator==(const Fp12& o) const { return c0 == o.c0 && c1 == o.c1; }
  Fp12 operator*(const Fp12& o) const {
    Fp6 t0 = c0 * o.c0, t1 = c1 * o.c1;
    return {t0 + t1.mul_v(), (c0 + c1) * (o.c0 + o.c1) - t0 - t1};
  }
  Fp12 sqr() const {
    Fp6 ab = c0 * c1;
    Fp6 t = (c0 + c1) * (c0 + c1.mul_v()) - ab - ab.mul_v();
    return {t, ab + ab};
  }
  Fp12 conj() const { return {c0, -c1}; }
  Fp12 inv() const {
    Fp6 d = (c0.sqr() - c1.sqr().mul_v()).inv();
    return {c0 * d, -(c1 * d)};
  }
  Fp12 frob() const;
  // sparse product with  a + b*v + (c*v)*w   (line function shape, coefficient slots 0, 1, 4)
  Fp12 mul_by_014(const Fp2& a, const Fp2& b, const Fp2& c) const {
    Fp12 l = {{a, b, Fp2::zero()}, {Fp2::zero(), c, Fp2::zero()}};
    return *this * l;
  }
  Fp12 pow_x_abs() const {  // this^|x| (square and multiply over the 64-bit |x|)
    Fp12 r = *this;
    for (int b = 62; b >= 0; b--) {
      r = r.sqr();
      if ((X_ABS >> b) & 1) r = r * *this;
    }
    return r;
  }
};

// ------------------------------------------------------------------------------------ curves
// generic homogeneous-projective point with RCB complete formulas; F = Fp (b3 = 12) or Fp2 (b3 = 12(1+u))
template <class F>
struct B3;
template <>
struct B3<Fp> {
  static Fp mul(const Fp& a) {
    Fp t4 = a.dbl().dbl();
    return t4.dbl() + t4;
  }
  static Fp b() { return Fp::from_u64(4); }
};
template <>
struct B3<Fp2> {
  static Fp2 mul(const Fp2& a) {
    Fp2 t4 = a.dbl().dbl();
    return (t4.dbl() + t4).mul_xi();
  }
  static Fp2 b() { return Fp2{Fp::from_u64(4), Fp::from_u64(4)}; }
};

template <class F>
struct Affine {
  F x, y;
  bool inf;
};

template <class F>
struct Proj {
  F x, y, z;
  static Proj identity() { return {F::zero(), F::one(), F::zero()}; }
  static Proj from_affine(const Affine<F>& a) { return a.inf ? identity() : Proj{a.x, a.y, F::one()}; }
  bool is_identity() const { return z.is_zero(); }
  Proj add(const Proj& q) const {  // RCB Alg. 7
    F t0 = x * q.x, t1 = y * q.y, t2 = z * q.z;
    F t3 = (x + y) * (q.x + q.y) - (t0 + t1);
    F t4 = (y + z) * (q.y + q.z) - (t1 + t2);
    F y3 = (x + z) * (q.x + q.z) - (t0 + t2);
    F x3 = t0.dbl();
    t0 = x3 + t0;
    t2 = B3<F>::mul(t2);
    F z3 = t1 + t2;
    t1 = t1 - t2;
    y3 = B3<F>::mul(y3);
    x3 = t4 * y3;
    t2 = t3 * t1;
    Proj r;
    r.x = t2 - x3;
    y3 = y3 * t0;
    t1 = t1 * z3;
    r.y = t1 + y3;
    t0 = t0 * t3;
    z3 = z3 * t4;
    r.z = z3 + t0;
    return r;
  }
  Proj add_mixed(const Affine<F>& q) const {  // RCB Alg. 8 (+ identity select)
    if (q.inf) return *this;
    F t0 = x * q.x, t1 = y * q.y;
    F t3 = (q.x + q.y) * (x + y) - (t0 + t1);
    F t4 = q.y * z + y;
    F y3 = q.x * z + x;
    F x3 = t0.dbl();
    t0 = x3 + t0;
    F t2 = B3<F>::mul(z);
    F z3 = t1 + t2;
    t1 = t1 - t2;
    y3 = B3<F>::mul(y3);
    x3 = t4 * y3;
    t2 = t3 * t1;
    Proj r;
    r.x = t2 - x3;
    y3 = y3 * t0;
    t1 = t1 * z3;
    r.y = t1 + y3;
    t0 = t0 * t3;
    z3 = z3 * t4;
    r.z = z3 + t0;
    return r;
  }
  Proj dbl() const {  // RCB Alg. 9
    F t0 = y.sqr();
    F z3 = t0.dbl().dbl().dbl();
    F t1 = y * z;
    F t2 = B3<F>::mul(z.sqr());
    F x3 = t2 * z3;
    F y3 = t0 + t2;
    Proj r;
    r.z = t1 * z3;
    t1 = t2.dbl();
    t2 = t1 + t2;
    t0 = t0 - t2;
    y3 = t0 * y3;
    r.y = x3 + y3;
    t1 = x * y;
    x3 = t0 * t1;
    r.x = x3.dbl();
    return r;
  }
  Proj neg() const { return {x, -y, z}; }
  Affine<F> to_affine() const {
    if (z.is_zero()) return {F::zero(), F::one(), true};
    F zi = z.inv();
    return {x * zi, y * zi, false};
  }
  bool eq(const Proj& o) const {
    bool ia = is_identity(), ib = o.is_identity();
    if (ia || ib) return ia && ib;
    return x * o.z == o.x * z && y * o.z == o.y * z;
  }
  // variable-time MSB-first double-and-add over little-endian limbs (skips leading zeros)
  Proj mul_vartime(const u64* k, int n) const {
    Proj acc = identity();
    bool started = false;
    for (int i = n - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        if (started) acc = acc.dbl();
        if ((k[i] >> b) & 1) {
          acc = started ? acc.add(*this) : *this;
          started = true;
        }
      }
    return acc;
  }
  // the reference's `G1Projective * Scalar`: 255 doublings and 255 full additions for EVERY scalar
  // (constant time) - used by the "faithful" CPU-baseline mode only
  Proj mul_consttime_256(const u64* k /*4 limbs*/) const {
    Proj acc = identity();
    for (int i = 254; i >= 0; i--) {  // top bit (255) skipped: scalars are < 2^255
      acc = acc.dbl();
      Proj s = acc.add(*this);
      if ((k[i / 64] >> (i % 64)) & 1) acc = s;
    }
    return acc;
  }
};
typedef Affine<Fp> G1Aff;
typedef Proj<Fp> G1;
typedef Affine<Fp2> G2Aff;
typedef Proj<Fp2> G2;

G1Aff g1_generator();
bool g1_on_curve(const G1Aff& a);
bool g1_in_subgroup(const G1Aff& a);  // [r]P == O
bool g2_in_subgroup(const G2Aff& a);

// decode status == include/dkgv.h dkgv_decode
enum { DEC_OK = 0, DEC_BAD_FLAGS = 1, DEC_X_RANGE = 2, DEC_NOT_ON_CURVE = 3, DEC_NOT_IN_SUBGROUP = 4 };
int g1_decompress(const uint8_t* in48, G1Aff* out);
void g1_compress(const G1Aff& a, uint8_t* out48);
int g2_decompress(const uint8_t* in96, G2Aff* out);
void g2_compress(const G2Aff& a, uint8_t* out96);

// ------------------------------------------------------------------------------------ pairing
Fp12 miller_loop(const G1Aff& p, const G2Aff& q);
Fp12 final_exponentiation(const Fp12& f);  // value = e(P,Q)^3 of the canonical optimal-ate pairing
Fp12 pairing(const G1Aff& p, const G2Aff& q);  // identity argument -> one

// ------------------------------------------------------------------------------------ hashing
void sha256(const uint8_t* msg, size_t len, uint8_t out[32]);
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len;
  size_t fill;
  Sha256();
  void update(const uint8_t* p, size_t n);
  void finish(uint8_t out[32]);
};
G2Aff hash_to_g2(const uint8_t* msg, size_t len, const uint8_t* dst, size_t dst_len);
extern const char* DST_POP;  // bls_common.rs:12

void init();  // idempotent; call before anything else
}  // namespace orc
