// Host-side utilities of the dkg-compatible front end: a minimal JSON reader (serde_json subset:
// objects, arrays, strings, integers, booleans, null; unknown fields are ignored as serde does,
// crates/dkg/src/types.rs has no deny_unknown_fields), SHA-256 for the commitment hashes
// (crates/dkg/src/verification.rs:151-175, 29-48, 333-362) and secp256k1 ECDSA verification for the
// identity signatures of BlsDkgWithSecp256kCommitment (crates/dkg/src/crypto/secp256k1_keys.rs:51-64).
// None of this is on the data-parallel hot path: one hash / one ECDSA check per *item*.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dkgh {

// ------------------------------------------------------------------------------------ JSON
struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  bool num_is_int = false;
  long long inum = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;
  const Json* get(const std::string& k) const {
    const Json* r = nullptr;
    for (auto& kv : obj)
      if (kv.first == k) r = &kv.second;  // last duplicate wins, as serde_json
    return r;
  }
  const Json& at(const std::string& k) const {
    const Json* r = get(k);
    if (!r) throw std::runtime_error("missing field `" + k + "`");
    return *r;
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string& s) : s_(s) {}
  Json parse() {
    Json v = value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m + " at " + std::to_string(i_)); }
  void ws() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) i_++;
  }
  Json value() {
    ws();
    if (i_ >= s_.size()) fail("unexpected end");
    char c = s_[i_];
    Json v;
    if (c == '{') {
      v.kind = Json::Obj;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == '}') {
        i_++;
        return v;
      }
      for (;;) {
        ws();
        Json k = value();
        if (k.kind != Json::Str) fail("object key must be a string");
        ws();
        if (i_ >= s_.size() || s_[i_] != ':') fail("expected ':'");
        i_++;
        Json val = value();
        v.obj.emplace_back(k.str, std::move(val));
        ws();
        if (i_ < s_.size() && s_[i_] == ',') {
          i_++;
          continue;
        }
        if (i_ < s_.size() && s_[i_] == '}') {
          i_++;
          return v;
        }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      v.kind = Json::Arr;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == ']') {
        i_++;
        return v;
      }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (i_ < s_.size() && s_[i_] == ',') {
          i_++;
          continue;
        }
        if (i_ < s_.size() && s_[i_] == ']') {
          i_++;
          return v;
        }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') {
      v.kind = Json::Str;
      i_++;
      while (i_ < s_.size() && s_[i_] != '"') {
        if (s_[i_] == '\\') {
          i_++;
          if (i_ >= s_.size()) fail("bad escape");
          char e = s_[i_++];
          switch (e) {
            case 'n': v.str += '\n'; break;
            case 't': v.str += '\t'; break;
            case 'r': v.str += '\r'; break;
            case 'b': v.str += '\b'; break;
            case 'f': v.str += '\f'; break;
            case '/': v.str += '/'; break;
            case '\\': v.str += '\\'; break;
            case '"': v.str += '"'; break;
            case 'u': {
              if (i_ + 4 > s_.size()) fail("bad \\u escape");
              unsigned cp = (unsigned)std::stoul(s_.substr(i_, 4), nullptr, 16);
              i_ += 4;
              if (cp < 0x80) v.str += (char)cp;
              else if (cp < 0x800) {
                v.str += (char)(0xC0 | (cp >> 6));
                v.str += (char)(0x80 | (cp & 0x3F));
              } else {
                v.str += (char)(0xE0 | (cp >> 12));
                v.str += (char)(0x80 | ((cp >> 6) & 0x3F));
                v.str += (char)(0x80 | (cp & 0x3F));
              }
              break;
            }
            default: fail("bad escape");
          }
        } else {
          v.str += s_[i_++];
        }
      }
      if (i_ >= s_.size()) fail("unterminated string");
      i_++;
      return v;
    }
    if (s_.compare(i_, 4, "true") == 0) {
      i_ += 4;
      v.kind = Json::Bool;
      v.b = true;
      return v;
    }
    if (s_.compare(i_, 5, "false") == 0) {
      i_ += 5;
      v.kind = Json::Bool;
      return v;
    }
    if (s_.compare(i_, 4, "null") == 0) {
      i_ += 4;
      return v;
    }
    size_t j = i_;
    if (j < s_.size() && (s_[j] == '-' || s_[j] == '+')) j++;
    bool is_int = true;
    while (j < s_.size() && (isdigit((unsigned char)s_[j]) || s_[j] == '.' || s_[j] == 'e' || s_[j] == 'E' || s_[j] == '-' || s_[j] == '+')) {
      if (!isdigit((unsigned char)s_[j])) is_int = false;
      j++;
    }
    if (j == i_) fail("unexpected character");
    v.kind = Json::Num;
    std::string t = s_.substr(i_, j - i_);
    v.num = std::stod(t);
    v.num_is_int = is_int || (t[0] == '-' && t.find_first_not_of("0123456789", 1) == std::string::npos);
    if (v.num_is_int) v.inum = std::stoll(t);
    i_ = j;
    return v;
  }
};

inline int hexv(char c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}
// hex string of exactly n bytes (the define_raw_type! newtypes, types.rs:346-358); throws otherwise
inline std::vector<uint8_t> hex_fixed(const Json& j, size_t n, const char* what) {
  if (j.kind != Json::Str) throw std::runtime_error(std::string(what) + ": expected a hex string");
  const std::string& s = j.str;
  if (s.size() != 2 * n) throw std::runtime_error(std::string(what) + ": invalid length");
  std::vector<uint8_t> out(n);
  for (size_t i = 0; i < n; i++) {
    int a = hexv(s[2 * i]), b = hexv(s[2 * i + 1]);
    if (a < 0 || b < 0) throw std::runtime_error(std::string(what) + ": invalid hex");
    out[i] = (uint8_t)(a * 16 + b);
  }
  return out;
}
inline uint8_t json_u8(const Json& j, const char* what) {
  if (j.kind != Json::Num || !j.num_is_int || j.inum < 0 || j.inum > 255) throw std::runtime_error(std::string(what) + ": expected u8");
  return (uint8_t)j.inum;
}

// ------------------------------------------------------------------------------------ SHA-256
class Sha256 {
 public:
  Sha256() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h_, iv, sizeof h_);
  }
  void update(const uint8_t* p, size_t n) {
    len_ += n;
    while (n) {
      size_t k = 64 - fill_ < n ? 64 - fill_ : n;
      memcpy(buf_ + fill_, p, k);
      fill_ += k;
      p += k;
      n -= k;
      if (fill_ == 64) {
        block(buf_);
        fill_ = 0;
      }
    }
  }
  void update(const std::vector<uint8_t>& v) { update(v.data(), v.size()); }
  std::vector<uint8_t> finish() {
    uint64_t bits = len_ * 8;
    uint8_t pad[72] = {0x80};
    size_t padlen = (fill_ < 56) ? 56 - fill_ : 120 - fill_;
    update(pad, padlen);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    std::vector<uint8_t> out(32);
    for (int i = 0; i < 8; i++) {
      out[4 * i] = (uint8_t)(h_[i] >> 24);
      out[4 * i + 1] = (uint8_t)(h_[i] >> 16);
      out[4 * i + 2] = (uint8_t)(h_[i] >> 8);
      out[4 * i + 3] = (uint8_t)h_[i];
    }
    return out;
  }

 private:
  uint32_t h_[8];
  uint8_t buf_[64];
  uint64_t len_ = 0;
  size_t fill_ = 0;
  static uint32_t rr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++)
      w[i] = w[i - 16] + (rr(w[i - 15], 7) ^ rr(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] + (rr(w[i - 2], 17) ^ rr(w[i - 2], 19) ^ (w[i - 2] >> 10));
    uint32_t a = h_[0], b = h_[1], c = h_[2], d = h_[3], e = h_[4], f = h_[5], g = h_[6], hh = h_[7];
    for (int i = 0; i < 64; i++) {
      uint32_t t1 = hh + (rr(e, 6) ^ rr(e, 11) ^ rr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      uint32_t t2 = (rr(a, 2) ^ rr(a, 13) ^ rr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h_[0] += a; h_[1] += b; h_[2] += c; h_[3] += d; h_[4] += e; h_[5] += f; h_[6] += g; h_[7] += hh;
  }
};

// ------------------------------------------------------------------------------------ secp256k1 ECDSA verify
// Plain 256-bit modular arithmetic (binary shift-and-add products; one verification costs a few ms,
// and there is one per *item*, never per share).
struct U256 {
  uint64_t w[4] = {0, 0, 0, 0};
  static U256 from_be(const uint8_t* b) {
    U256 r;
    for (int i = 0; i < 32; i++) r.w[i / 8] |= (uint64_t)b[31 - i] << (8 * (i % 8));
    return r;
  }
  bool is_zero() const { return !(w[0] | w[1] | w[2] | w[3]); }
  bool bit(int i) const { return (w[i / 64] >> (i % 64)) & 1; }
};
inline int u256_cmp(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; i--)
    if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
  return 0;
}
inline uint64_t u256_add(U256& r, const U256& a, const U256& b) {
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (unsigned __int128)a.w[i] + b.w[i];
    r.w[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
inline uint64_t u256_sub(U256& r, const U256& a, const U256& b) {
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) {
    unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - br;
    r.w[i] = (uint64_t)d;
    br = (uint64_t)(d >> 64) & 1;
  }
  return br;
}
struct ModArith {
  U256 m;
  U256 add(const U256& a, const U256& b) const {
    U256 r;
    uint64_t c = u256_add(r, a, b);
    if (c || u256_cmp(r, m) >= 0) u256_sub(r, r, m);
    return r;
  }
  U256 sub(const U256& a, const U256& b) const {
    U256 r;
    if (u256_sub(r, a, b)) u256_add(r, r, m);
    return r;
  }
  U256 mul(const U256& a, const U256& b) const {
    U256 r;
    for (int i = 255; i >= 0; i--) {
      r = add(r, r);
      if (b.bit(i)) r = add(r, a);
    }
    return r;
  }
  U256 pow(const U256& a, const U256& e) const {
    U256 r;
    r.w[0] = 1;
    for (int i = 255; i >= 0; i--) {
      r = mul(r, r);
      if (e.bit(i)) r = mul(r, a);
    }
    return r;
  }
  U256 inv(const U256& a) const {
    U256 e, two;
    two.w[0] = 2;
    u256_sub(e, m, two);
    return pow(a, e);
  }
};
struct SecpPoint {
  U256 x, y;
  bool inf = true;
};
class Secp256k1 {
 public:
  Secp256k1() {
    static const uint8_t P[32] = {0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff,
                                  0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xfe, 0xff, 0xff, 0xfc, 0x2f};
    static const uint8_t N[32] = {0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xfe,
                                  0xba, 0xae, 0xdc, 0xe6, 0xaf, 0x48, 0xa0, 0x3b, 0xbf, 0xd2, 0x5e, 0x8c, 0xd0, 0x36, 0x41, 0x41};
    static const uint8_t GX[32] = {0x79, 0xbe, 0x66, 0x7e, 0xf9, 0xdc, 0xbb, 0xac, 0x55, 0xa0, 0x62, 0x95, 0xce, 0x87, 0x0b, 0x07,
                                   0x02, 0x9b, 0xfc, 0xdb, 0x2d, 0xce, 0x28, 0xd9, 0x59, 0xf2, 0x81, 0x5b, 0x16, 0xf8, 0x17, 0x98};
    static const uint8_t GY[32] = {0x48, 0x3a, 0xda, 0x77, 0x26, 0xa3, 0xc4, 0x65, 0x5d, 0xa4, 0xfb, 0xfc, 0x0e, 0x11, 0x08, 0xa8,
                                   0xfd, 0x17, 0xb4, 0x48, 0xa6, 0x85, 0x54, 0x19, 0x9c, 0x47, 0xd0, 0x8f, 0xfb, 0x10, 0xd4, 0xb8};
    fp.m = U256::from_be(P);
    fn.m = U256::from_be(N);
    g.x = U256::from_be(GX);
    g.y = U256::from_be(GY);
    g.inf = false;
  }
  // secp256k1::PublicKey::from_slice on a 33-byte SEC1 compressed key
  bool parse_pubkey(const uint8_t* b, SecpPoint* out) const {
    if (b[0] != 2 && b[0] != 3) return false;
    U256 x = U256::from_be(b + 1);
    if (u256_cmp(x, fp.m) >= 0) return false;
    U256 seven;
    seven.w[0] = 7;
    U256 y2 = fp.add(fp.mul(fp.mul(x, x), x), seven);
    U256 e, one;  // (p + 1) / 4
    one.w[0] = 1;
    u256_add(e, fp.m, one);
    for (int i = 0; i < 4; i++) e.w[i] = (e.w[i] >> 2) | (i < 3 ? e.w[i + 1] << 62 : 0);  // p + 1 < 2^256: no carry
    U256 y = fp.pow(y2, e);
    if (u256_cmp(fp.mul(y, y), y2) != 0) return false;
    if ((y.w[0] & 1) != (uint64_t)(b[0] & 1)) y = fp.sub(U256(), y);
    out->x = x;
    out->y = y;
    out->inf = false;
    return true;
  }
  // ecdsa::Signature::from_compact: fails when r or s is not below the group order
  bool parse_sig(const uint8_t* b, U256* r, U256* s) const {
    *r = U256::from_be(b);
    *s = U256::from_be(b + 32);
    return u256_cmp(*r, fn.m) < 0 && u256_cmp(*s, fn.m) < 0;
  }
  // verify_ecdsa on a 32-byte digest; libsecp256k1 rejects zero r/s and high-S signatures
  bool verify(const SecpPoint& pk, const uint8_t* digest32, const U256& r, const U256& s) const {
    if (r.is_zero() || s.is_zero()) return false;
    U256 half = fn.m;  // n / 2
    for (int i = 0; i < 4; i++) half.w[i] = (half.w[i] >> 1) | (i < 3 ? fn.m.w[i + 1] << 63 : 0);
    if (u256_cmp(s, half) > 0) return false;
    U256 z = U256::from_be(digest32);
    while (u256_cmp(z, fn.m) >= 0) u256_sub(z, z, fn.m);
    U256 w = fn.inv(s);
    SecpPoint p = padd(pmul(g, fn.mul(z, w)), pmul(pk, fn.mul(r, w)));
    if (p.inf) return false;
    U256 xr = p.x;
    while (u256_cmp(xr, fn.m) >= 0) u256_sub(xr, xr, fn.m);
    return u256_cmp(xr, r) == 0;
  }

 private:
  ModArith fp, fn;
  SecpPoint g;
  SecpPoint padd(const SecpPoint& a, const SecpPoint& b) const {
    if (a.inf) return b;
    if (b.inf) return a;
    U256 lam;
    if (u256_cmp(a.x, b.x) == 0) {
      if (fp.add(a.y, b.y).is_zero()) return SecpPoint();
      U256 three;
      three.w[0] = 3;
      lam = fp.mul(fp.mul(three, fp.mul(a.x, a.x)), fp.inv(fp.add(a.y, a.y)));
    } else {
      lam = fp.mul(fp.sub(b.y, a.y), fp.inv(fp.sub(b.x, a.x)));
    }
    SecpPoint r;
    r.x = fp.sub(fp.sub(fp.mul(lam, lam), a.x), b.x);
    r.y = fp.sub(fp.mul(lam, fp.sub(a.x, r.x)), a.y);
    r.inf = false;
    return r;
  }
  // Jacobian double-and-add (one inversion at the end); a = 0 curve
  SecpPoint pmul(const SecpPoint& p, const U256& k) const {
    if (p.inf) return p;
    U256 X, Y, Z;  // Z == 0: infinity
    for (int i = 255; i >= 0; i--) {
      if (!Z.is_zero()) {  // dbl-2009-l
        U256 A = fp.mul(X, X), B = fp.mul(Y, Y), C = fp.mul(B, B);
        U256 t = fp.add(X, B);
        U256 D = fp.sub(fp.sub(fp.mul(t, t), A), C);
        D = fp.add(D, D);
        U256 E = fp.add(fp.add(A, A), A), F = fp.mul(E, E);
        U256 X3 = fp.sub(F, fp.add(D, D));
        U256 C8 = fp.add(C, C);
        C8 = fp.add(C8, C8);
        C8 = fp.add(C8, C8);
        U256 Y3 = fp.sub(fp.mul(E, fp.sub(D, X3)), C8);
        U256 Z3 = fp.mul(fp.add(Y, Y), Z);
        X = X3;
        Y = Y3;
        Z = Z3;
      }
      if (k.bit(i)) {
        if (Z.is_zero()) {
          X = p.x;
          Y = p.y;
          Z = U256();
          Z.w[0] = 1;
        } else {  // madd-2007-bl with the exceptional cases handled through the affine law
          U256 Z1Z1 = fp.mul(Z, Z), U2 = fp.mul(p.x, Z1Z1), S2 = fp.mul(fp.mul(p.y, Z), Z1Z1);
          U256 Hh = fp.sub(U2, X), rr = fp.sub(S2, Y);
          if (Hh.is_zero()) {
            SecpPoint cur = to_affine(X, Y, Z);
            SecpPoint sum = padd(cur, p);
            if (sum.inf) {
              Z = U256();
            } else {
              X = sum.x;
              Y = sum.y;
              Z = U256();
              Z.w[0] = 1;
            }
          } else {
            U256 HH = fp.mul(Hh, Hh), HHH = fp.mul(Hh, HH), V = fp.mul(X, HH);
            U256 X3 = fp.sub(fp.sub(fp.mul(rr, rr), HHH), fp.add(V, V));
            U256 Y3 = fp.sub(fp.mul(rr, fp.sub(V, X3)), fp.mul(Y, HHH));
            Z = fp.mul(Z, Hh);
            X = X3;
            Y = Y3;
          }
        }
      }
    }
    return to_affine(X, Y, Z);
  }
  SecpPoint to_affine(const U256& X, const U256& Y, const U256& Z) const {
    SecpPoint r;
    if (Z.is_zero()) return r;
    U256 zi = fp.inv(Z), zi2 = fp.mul(zi, zi);
    r.x = fp.mul(X, zi2);
    r.y = fp.mul(Y, fp.mul(zi2, zi));
    r.inf = false;
    return r;
  }
};

}  // namespace dkgh
