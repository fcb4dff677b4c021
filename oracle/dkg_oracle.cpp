// CPU ORACLE - TEST INFRASTRUCTURE ONLY (see bls.hpp).  C entry points for ctypes (tests/,
// bench.py cpu_baseline) restating the math-level functions of the reference's hot path:
//   evaluate_polynomial   crates/dkg/src/dkg_math.rs:160-174
//   lagrange_interpolation crates/dkg/src/dkg_math.rs:178-227
//   agg_coefficients      crates/dkg/src/dkg_math.rs:230-248
//   BlsG1::add/mul_scalar crates/dkg/src/dkg_math.rs:114-127   (mode 0 keeps their per-op affine
//                         round trip and the constant-time 255-step scalar multiplication)
//   to_public_key         crates/dkg/src/crypto/bls_keys.rs:133-137
//   share comparison      crates/dkg/src/verification.rs:92-99,129-146
//   bls_verify*           crates/dkg/src/crypto/bls_common.rs:26-40
// mode 0 = "faithful" (the reference's operation sequence; this is the CPU baseline),
// mode 1 = "fast"     (projective small-scalar Horner: the canonical algorithm of SURVEY 8(d)).
// Status codes are those of include/dkgv.h.
#include <thread>

#include "../include/dkgv.h"
#include "bls.hpp"

using namespace orc;

namespace {
// ---- reference-faithful TPoint ops on affine values (dkg_math.rs:114-127)
G1Aff tpoint_add(const G1Aff& a, const G1Aff& b) { return G1::from_affine(a).add(G1::from_affine(b)).to_affine(); }
G1Aff tpoint_mul_scalar(const G1Aff& a, const u64* k4) { return G1::from_affine(a).mul_consttime_256(k4).to_affine(); }

G1Aff eval_poly_faithful(const std::vector<G1Aff>& cfs, const u64* x4) {
  size_t n = cfs.size();
  if (n == 0) return {Fp::zero(), Fp::one(), true};
  if (n == 1) return cfs[0];
  G1Aff y = cfs[n - 1];
  for (size_t i = 2; i <= n; i++) {
    y = tpoint_mul_scalar(y, x4);
    y = tpoint_add(y, cfs[n - i]);
  }
  return y;
}
G1 mul_small(const G1& p, uint32_t k) {
  if (k == 0) return G1::identity();
  u64 kk[1] = {k};
  return p.mul_vartime(kk, 1);
}
G1Aff eval_poly_fast(const std::vector<G1Aff>& cfs, uint32_t id) {
  size_t n = cfs.size();
  if (n == 0) return {Fp::zero(), Fp::one(), true};
  G1 acc = G1::from_affine(cfs[n - 1]);
  for (size_t i = 2; i <= n; i++) acc = mul_small(acc, id).add_mixed(cfs[n - i]);
  return acc.to_affine();
}
G1Aff eval_poly(const std::vector<G1Aff>& cfs, uint32_t id, int mode) {
  if (mode == 0) {
    u64 x4[4] = {id, 0, 0, 0};
    return eval_poly_faithful(cfs, x4);
  }
  return eval_poly_fast(cfs, id);
}
bool decode_vv(const uint8_t* vv, uint32_t t, std::vector<G1Aff>* out) {
  out->resize(t);
  bool ok = true;
  for (uint32_t k = 0; k < t; k++)
    if (g1_decompress(vv + (size_t)k * 48, &(*out)[k]) != DEC_OK) ok = false;
  return ok;
}
bool scalar_from_be(const uint8_t* b, u64* raw4) {
  memset(raw4, 0, 32);
  for (int i = 0; i < 32; i++) raw4[i / 8] |= (u64)b[31 - i] << (8 * (i % 8));
  return cmp<4>(raw4, Fr::MOD) < 0;
}
G1Aff g_times(const u64* s4, int mode) {
  G1 g = G1::from_affine(g1_generator());
  return (mode == 0 ? g.mul_consttime_256(s4) : g.mul_vartime(s4, 4)).to_affine();
}
uint8_t share_status(const std::vector<G1Aff>& cfs, bool vv_ok, uint32_t id, const uint8_t* secret, int mode) {
  u64 s4[4];
  if (!scalar_from_be(secret, s4)) return DKGV_SLASHABLE_SECRET_RANGE;
  if (!vv_ok) return DKGV_PANIC_BAD_G1;
  uint8_t a[48], b[48];
  g1_compress(eval_poly(cfs, id, mode), a);
  g1_compress(g_times(s4, mode), b);
  return memcmp(a, b, 48) ? DKGV_SLASHABLE_SHARE_MISMATCH : DKGV_OK;
}
}  // namespace

// The SAME exact shortcut the GPU library's default share path takes (dvt_circuits_b200/csrc/share_fd.cu), on the CPU, so that a
// bench can separate what the algorithm buys from what the hardware buys: all n shares of a dealer are valid iff each is < r, their
// t-th forward differences vanish, and compress(G * p_k) == C_k for every coefficient of the interpolated polynomial p.  ids must be
// 1..n_r in order.  status[d][j] = OK for a dealer that meets the conditions; dealers that do not are evaluated share by share (fast
// mode).  Test infrastructure (the CPU leg of bench.py), never part of the product.  fixed-base G * s: 8-bit windows, built once.
static std::vector<G1Aff> g_gwin;  // [32][256]: (b * 2^(8w)) * G, b = 0..255 (b = 0: identity)
static void build_gwin() {
  if (!g_gwin.empty()) return;
  std::vector<G1Aff> tab(32 * 256);
  G1 base = G1::from_affine(g1_generator());
  for (int w = 0; w < 32; w++) {
    G1 acc = G1::identity();
    for (int b = 0; b < 256; b++) {
      tab[w * 256 + b] = acc.to_affine();
      acc = acc.add(base);
    }
    base = acc;  // 256 * previous base
  }
  g_gwin.swap(tab);
}
static G1Aff g_times_windowed(const u64* s4) {
  G1 acc = G1::identity();
  for (int w = 0; w < 32; w++) {
    unsigned b = (unsigned)(s4[w / 8] >> (8 * (w % 8))) & 0xff;
    if (b) acc = acc.add_mixed(g_gwin[w * 256 + b]);
  }
  return acc.to_affine();
}
extern "C" void orc_share_matrix_shortcut(uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* vv, const uint8_t* shares, uint8_t* status,
                                          int threads, uint32_t* n_fallback) {
  init();
  build_gwin();
  if (threads < 1) threads = 1;
  std::vector<uint32_t> fb(threads, 0);
  auto work = [&](int tid) {
    std::vector<Fr> d(n_r), c(t);
    for (uint32_t dl = tid; dl < n_d; dl += threads) {
      bool ok = n_r > t && t >= 1;
      for (uint32_t j = 0; ok && j < n_r; j++) ok = Fr::from_be(&d[j], shares + ((size_t)dl * n_r + j) * 32, 32);
      if (ok) {  // t rounds of differences: d[k] = Delta^k s(1) for k < t, the t-th differences beyond
        for (uint32_t r = 1; r <= t; r++)
          for (uint32_t k = n_r - 1; k >= r; k--) d[k] = d[k] - d[k - 1];
        for (uint32_t k = t; ok && k < n_r; k++) ok = d[k].is_zero();
      }
      if (ok) {  // Newton basis at the nodes 1, 2, ... -> monomial coefficients:  P <- P (x - j) + Delta^(j-1) s(1) / (j-1)!
        Fr fact = Fr::one();
        std::vector<Fr> E(t);
        for (uint32_t k = 0; k < t; k++) {
          if (k) fact = fact * Fr::from_u64(k);
          E[k] = d[k] * fact.inv();
        }
        for (uint32_t k = 0; k < t; k++) c[k] = Fr::zero();
        c[0] = E[t - 1];
        for (uint32_t j = t - 1; j >= 1; j--) {
          Fr fj = Fr::from_u64(j);
          for (uint32_t k = t - j; k >= 1; k--) c[k] = c[k - 1] - fj * c[k];
          c[0] = E[j - 1] - fj * c[0];
        }
        for (uint32_t k = 0; ok && k < t; k++) {  // compress(G * p_k) == C_k, the commitment never decoded
          u64 raw[4];
          c[k].to_raw(raw);
          uint8_t enc[48];
          g1_compress(g_times_windowed(raw), enc);
          ok = memcmp(enc, vv + ((size_t)dl * t + k) * 48, 48) == 0;
        }
      }
      if (ok) {
        memset(status + (size_t)dl * n_r, DKGV_OK, n_r);
      } else {
        fb[tid]++;
        std::vector<G1Aff> cfs;
        bool vok = decode_vv(vv + (size_t)dl * t * 48, t, &cfs);
        for (uint32_t j = 0; j < n_r; j++) status[(size_t)dl * n_r + j] = share_status(cfs, vok, j + 1, shares + ((size_t)dl * n_r + j) * 32, 1);
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; i++) th.emplace_back(work, i);
  work(0);
  for (auto& x : th) x.join();
  if (n_fallback) {
    *n_fallback = 0;
    for (uint32_t f : fb) *n_fallback += f;
  }
}

extern "C" {
int orc_init() {
  init();
  return 0;
}
uint64_t orc_fp_mul_count(int reset) {
  uint64_t c = Fp::MULS;
  if (reset) Fp::MULS = 0;
  return c;
}
int orc_g1_decompress(const uint8_t* in48, uint8_t* out48) {
  init();
  G1Aff a;
  int st = g1_decompress(in48, &a);
  if (st == DEC_OK && out48) g1_compress(a, out48);
  return st;
}
int orc_g2_decompress(const uint8_t* in96, uint8_t* out96) {
  init();
  G2Aff a;
  int st = g2_decompress(in96, &a);
  if (st == DEC_OK && out96) g2_compress(a, out96);
  return st;
}
int orc_g1_fixed_base(const uint8_t* s32, int mode, uint8_t* out48) {
  init();
  u64 s4[4];
  if (!scalar_from_be(s32, s4)) return DKGV_SLASHABLE_SECRET_RANGE;
  g1_compress(g_times(s4, mode), out48);
  return DKGV_OK;
}
int orc_evaluate_polynomial(const uint8_t* vv, uint32_t t, uint32_t id, int mode, uint8_t* out48) {
  init();
  std::vector<G1Aff> cfs;
  if (!decode_vv(vv, t, &cfs)) return DKGV_PANIC_BAD_G1;
  g1_compress(eval_poly(cfs, id, mode), out48);
  return DKGV_OK;
}
int orc_share_verify(const uint8_t* vv, uint32_t t, uint32_t id, const uint8_t* secret32, int mode) {
  init();
  std::vector<G1Aff> cfs;
  bool ok = decode_vv(vv, t, &cfs);
  return share_status(cfs, ok, id, secret32, mode);
}
// full (dealer x recipient) matrix, rows split over `threads` host threads.
// sample_stride > 1 checks only every sample_stride-th share (row-major index) and writes 0xff elsewhere.
void orc_share_matrix(uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* vv, const uint32_t* ids, const uint8_t* shares,
                      uint8_t* status, int mode, int threads, uint32_t sample_stride) {
  init();
  if (threads < 1) threads = 1;
  if (sample_stride < 1) sample_stride = 1;
  auto work = [&](int tid) {
    for (uint32_t d = tid; d < n_d; d += threads) {
      std::vector<G1Aff> cfs;
      bool any = false;
      for (uint32_t j = 0; j < n_r; j++) any |= (((size_t)d * n_r + j) % sample_stride) == 0;
      bool ok = true;
      if (any) ok = decode_vv(vv + (size_t)d * t * 48, t, &cfs);
      for (uint32_t j = 0; j < n_r; j++) {
        size_t idx = (size_t)d * n_r + j;
        status[idx] = (idx % sample_stride) ? 0xff : share_status(cfs, ok, ids[j], shares + idx * 32, mode);
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; i++) th.emplace_back(work, i);
  work(0);
  for (auto& x : th) x.join();
}

// agg_coefficients: vv [n][t][48] (t = len of vv[0]; ragged inputs are rejected by the caller),
// coeff_out [t][48] column sums, keys_out [n_ids][48] = evaluate_polynomial(coeffs, ids[j])
int orc_agg_coefficients(uint32_t n, uint32_t t, const uint8_t* vv, const uint32_t* ids, uint32_t n_ids, int mode, uint8_t* coeff_out,
                         uint8_t* keys_out) {
  init();
  std::vector<std::vector<G1Aff>> pts(n);
  for (uint32_t i = 0; i < n; i++)
    if (!decode_vv(vv + (size_t)i * t * 48, t, &pts[i])) return DKGV_PANIC_BAD_G1;
  std::vector<G1Aff> cfs(t);
  for (uint32_t k = 0; k < t; k++) {
    if (mode == 0) {
      G1Aff s{Fp::zero(), Fp::one(), true};
      for (uint32_t i = 0; i < n; i++) s = tpoint_add(s, pts[i][k]);
      cfs[k] = s;
    } else {
      G1 s = G1::identity();
      for (uint32_t i = 0; i < n; i++) s = s.add_mixed(pts[i][k]);
      cfs[k] = s.to_affine();
    }
    if (coeff_out) g1_compress(cfs[k], coeff_out + (size_t)k * 48);
  }
  for (uint32_t j = 0; j < n_ids; j++) g1_compress(eval_poly(cfs, ids[j], mode), keys_out + (size_t)j * 48);
  return DKGV_OK;
}

// lagrange_interpolation at 0 (dkg_math.rs:178-227); ids are u32 -> Scalar (bls_common.rs:42-47)
int orc_lagrange(uint32_t k, const uint8_t* pts48, const uint32_t* ids, int mode, uint8_t* out48) {
  init();
  if (k == 0) return DKGV_ERR_LEN;
  std::vector<G1Aff> ys(k);
  for (uint32_t i = 0; i < k; i++)
    if (g1_decompress(pts48 + (size_t)i * 48, &ys[i]) != DEC_OK) return DKGV_PANIC_BAD_G1;
  if (k == 1) {
    g1_compress(ys[0], out48);
    return DKGV_OK;
  }
  std::vector<Fr> xs(k);
  for (uint32_t i = 0; i < k; i++) xs[i] = Fr::from_u64(ids[i]);
  Fr a = xs[0];
  for (uint32_t i = 1; i < k; i++) a = a * xs[i];
  if (a.is_zero()) return DKGV_ERR_ZERO_ID;
  G1 r = G1::identity();
  G1Aff ra{Fp::zero(), Fp::one(), true};
  for (uint32_t i = 0; i < k; i++) {
    Fr b = xs[i];
    for (uint32_t j = 0; j < k; j++)
      if (j != i) {
        Fr v = xs[j] - xs[i];
        if (v.is_zero()) return DKGV_ERR_DUP_ID;
        b = b * v;
      }
    Fr li0 = a * b.inv();
    u64 raw[4];
    li0.to_raw(raw);
    if (mode == 0) {
      ra = tpoint_add(ra, tpoint_mul_scalar(ys[i], raw));
    } else {
      r = r.add(G1::from_affine(ys[i]).mul_vartime(raw, 4));
    }
  }
  g1_compress(mode == 0 ? ra : r.to_affine(), out48);
  return DKGV_OK;
}

void orc_hash_to_g2(const uint8_t* msg, size_t len, uint8_t* out96) {
  init();
  g2_compress(hash_to_g2(msg, len, (const uint8_t*)DST_POP, strlen(DST_POP)), out96);
}
// bls_verify_precomputed_hash (bls_common.rs:26-35): 1 = valid, 0 = invalid, <0: -48 bad pk / -49 bad sig or hm
int orc_bls_verify_hm(const uint8_t* pk48, const uint8_t* sig96, const uint8_t* hm96) {
  init();
  G1Aff pk;
  G2Aff sig, hm;
  if (g1_decompress(pk48, &pk) != DEC_OK) return -DKGV_PANIC_BAD_G1;
  if (g2_decompress(sig96, &sig) != DEC_OK) return -DKGV_PANIC_BAD_G2;
  if (g2_decompress(hm96, &hm) != DEC_OK) return -DKGV_PANIC_BAD_G2;
  return pairing(pk, hm) == pairing(g1_generator(), sig) ? 1 : 0;
}
int orc_bls_verify(const uint8_t* pk48, const uint8_t* sig96, const uint8_t* msg, size_t len) {
  init();
  uint8_t hm[96];
  orc_hash_to_g2(msg, len, hm);
  return orc_bls_verify_hm(pk48, sig96, hm);
}
// batch of independent checks against one hashed message, split over host threads
void orc_bls_verify_batch(uint32_t m, const uint8_t* pk48, const uint8_t* sig96, const uint8_t* hm96, int8_t* out, int threads) {
  init();
  if (threads < 1) threads = 1;
  auto work = [&](int tid) {
    for (uint32_t i = tid; i < m; i += threads) out[i] = (int8_t)orc_bls_verify_hm(pk48 + (size_t)i * 48, sig96 + (size_t)i * 96, hm96);
  };
  std::vector<std::thread> th;
  for (int i = 1; i < threads; i++) th.emplace_back(work, i);
  work(0);
  for (auto& x : th) x.join();
}
void orc_sha256(const uint8_t* msg, size_t len, uint8_t* out32) { sha256(msg, len, out32); }
// raw pairing value e(P,Q)^3 as 576 canonical big-endian bytes (c0.c0.c0, c0.c0.c1, ... ) for cross-checks
int orc_pairing_bytes(const uint8_t* p48, const uint8_t* q96, uint8_t* out576) {
  init();
  G1Aff p;
  G2Aff q;
  if (g1_decompress(p48, &p) != DEC_OK || g2_decompress(q96, &q) != DEC_OK) return -1;
  Fp12 e = pairing(p, q);
  const Fp* c = &e.c0.c0.c0;
  for (int i = 0; i < 12; i++) c[i].to_be(out576 + 48 * i, 48);
  return 0;
}
}
