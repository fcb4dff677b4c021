// Per-thread building blocks of the Feldman share check (host/device so the very same code is
// exercised by the CPU-side emulation tests):
//   evaluate_polynomial  crates/dkg/src/dkg_math.rs:160-174  (Horner in the exponent)
//   G * s                crates/dkg/src/crypto/bls_keys.rs:133-137
//   compare              crates/dkg/src/verification.rs:140-146
//
// HBM layout of a session's verification vectors ("coefficient-major, limb-planar"):
//   limbs[(k*24 + w) * n_pad + d]   w-th 32-bit limb (x: 0..11, y: 12..23, Montgomery form) of
//                                   coefficient k of dealer d;  n_pad = dealers rounded up to 32
//   inf[k * n_pad + d]              1 when the coefficient is the identity
// so that the 32 lanes of a warp (32 consecutive dealers, one recipient id) read 128 contiguous
// bytes per limb.
#pragma once
#include "../../include/dkgv.h"
#include "g1.cuh"
#include "gtab.hpp"

namespace dkgv {

struct VVView {
  const uint32_t* limbs;
  const uint8_t* inf;
  uint32_t n_pad;
};

DKGV_HD G1Aff vv_load(const VVView& v, uint32_t k, uint32_t d) {
  G1Aff a;
  const uint32_t* base = v.limbs + (size_t)k * 24 * v.n_pad + d;
#pragma unroll
  for (int w = 0; w < 12; w++) {
    a.x.l[w] = base[(size_t)w * v.n_pad];
    a.y.l[w] = base[(size_t)(w + 12) * v.n_pad];
  }
  a.inf = v.inf[(size_t)k * v.n_pad + d];
  return a;
}

DKGV_HD void vv_store(uint32_t* limbs, uint8_t* inf, uint32_t n_pad, uint32_t k, uint32_t d, const G1Aff& a) {
  uint32_t* base = limbs + (size_t)k * 24 * n_pad + d;
#pragma unroll
  for (int w = 0; w < 12; w++) {
    base[(size_t)w * n_pad] = a.x.l[w];
    base[(size_t)(w + 12) * n_pad] = a.y.l[w];
  }
  inf[(size_t)k * n_pad + d] = (uint8_t)(a.inf != 0);
}

// [k]P for a small public scalar (recipient id): left-to-right binary, top bit free.
// Control flow depends on k only - warp-uniform when all lanes share the id.
DKGV_HD G1Proj g1_mul_small(const G1Proj& p, uint32_t k) {
  if (k == 0) return g1_identity();
  int top = 31;
  while (!((k >> top) & 1)) top--;
  G1Proj acc = p;
#pragma unroll 1
  for (int b = top - 1; b >= 0; b--) {
    acc = g1_dbl(acc);
    if ((k >> b) & 1) acc = g1_add(acc, p);
  }
  return acc;
}

// sum_k C_k * id^k by Horner from the top coefficient (dkg_math.rs:160-174)
DKGV_HD G1Proj feldman_eval(const VVView& v, uint32_t t, uint32_t d, uint32_t id) {
  if (t == 0) return g1_identity();
  G1Proj acc = g1_from_affine(vv_load(v, t - 1, d));
#pragma unroll 1
  for (int k = (int)t - 2; k >= 0; k--) {
    acc = g1_mul_small(acc, id);
    acc = g1_add_mixed(acc, vv_load(v, (uint32_t)k, d));
  }
  return acc;
}

// Fixed-base table for the generator with SIGNED ODD digits, B-bit windows.  The scalar is made odd first (u = s or
// s + r: G has order r and 2r < 2^256), and an odd u < 2^256 has exactly one representation u = sum_w d_w 2^(B w) with every
// d_w odd, |d_w| < 2^B (d_w = [bits B w .. B w + B of u, bit B w forced to 1] - 2^B for w < W - 1, the last digit is what remains):
// no digit is zero - the mixed addition never meets the identity and needs no select, no correction point either - and only the
// odd positive multiples are stored, a negative digit negates y:
//   gtab[(w * 2^(B-1) + m) * 24 ..] = ((2m + 1) * 2^(B w)) * G      m < 2^(B-1), w < W = ceil(256 / B)   (affine Montgomery)
// G*s = sum_w +-gtab[w][(|d_w| - 1) / 2].  B = 16: 16 windows, the first entry initialises the sum, 15 mixed additions per
// multiplication (21 with the unsigned 13-bit offset windows of the same idea before, 33 with byte windows); 524 288 entries = 50 MB,
// in the 126 MB L2 (the default share path does t fixed-base multiplications per dealer - 699 392 per (1024, 683) ceremony - and
// its x-half kernel is more than half of the step).
// The window width B is chosen per ctx (GTAB_BITS_MIN..GTAB_BITS_MAX): the table walk is one random 96-byte read per window, which the
// HBM serves as well as the L2 does, so a wider window trades memory for mixed additions - B = 16: 50 MB / 15 additions, B = 22
// (default): 2.4 GB / 11, B = 26: 32 GB / 9.
constexpr uint32_t GTAB_BITS_MIN = 8, GTAB_BITS_MAX = 26, GTAB_BITS_DEFAULT = 22;
DKGV_HD uint32_t gtab_windows(uint32_t bits) { return (256 + bits - 1) / bits; }
DKGV_HD uint32_t gtab_entries(uint32_t bits) { return gtab_windows(bits) << (bits - 1); }
DKGV_HD size_t gtab_words(uint32_t bits) { return (size_t)gtab_entries(bits) * 24; }

// entry idx of the table with `bits`-bit windows, the slow way (doublings + a small multiplication + one inversion): the reference the
// table builder (dkgv.cu k_gtab_*) is tested against, and what the host emulation fills its sparse table with
DKGV_HD G1Aff gtab_entry(uint32_t bits, uint32_t idx) {
  uint32_t w = idx >> (bits - 1), m = idx & ((1u << (bits - 1)) - 1u);
  G1Proj p = g1_from_affine(g1_generator());
#pragma unroll 1
  for (uint32_t i = 0; i < bits * w; i++) p = g1_dbl(p);
  return g1_to_affine(g1_mul_small(p, 2 * m + 1));
}

// ---- building the table (dkgv.cu k_gtab_bases / k_gtab_fill; the same routines run on the host in tests/hostemu) --------------
// base[w * 48 ..]: affine x, y of 2^(B w) G (24 words), then of 2 * 2^(B w) G (24 words)
DKGV_HD void gtab_base(uint32_t bits, uint32_t w, uint32_t* base) {
  G1Proj p = g1_from_affine(g1_generator());
#pragma unroll 1
  for (uint32_t i = 0; i < bits * w; i++) p = g1_dbl(p);
  G1Aff a = g1_to_affine(p), d = g1_to_affine(g1_dbl(p));
#pragma unroll
  for (int i = 0; i < 12; i++) {
    base[(size_t)w * 48 + i] = a.x.l[i];
    base[(size_t)w * 48 + 12 + i] = a.y.l[i];
    base[(size_t)w * 48 + 24 + i] = d.x.l[i];
    base[(size_t)w * 48 + 36 + i] = d.y.l[i];
  }
}
// GTAB_RUN consecutive entries m0 .. m0 + GTAB_RUN - 1 of window w: one small multiplication for the first, then a walk by
// 2 * 2^(B w) G with one mixed addition per entry, and ONE inversion for the whole run (Montgomery's trick over the Z's; the
// projective X, Y wait in their own table slots).  ~28 products per entry instead of the ~3 000 of gtab_entry.
constexpr uint32_t GTAB_RUN = 32;
DKGV_HD void gtab_fill_run(uint32_t bits, const uint32_t* base, uint32_t w, uint32_t m0, uint32_t* gtab) {
  G1Aff g;
  Fp dx, dy;
  g.inf = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    g.x.l[i] = base[(size_t)w * 48 + i];
    g.y.l[i] = base[(size_t)w * 48 + 12 + i];
    dx.l[i] = base[(size_t)w * 48 + 24 + i];
    dy.l[i] = base[(size_t)w * 48 + 36 + i];
  }
  G1Proj p = g1_mul_small(g1_from_affine(g), 2 * m0 + 1);
  uint32_t* e0 = gtab + ((size_t)(w << (bits - 1)) + m0) * 24;
  Fp zs[GTAB_RUN], pre[GTAB_RUN];  // Z_i and Z_0 ... Z_i (no multiple of the generator in the table is the identity: Z_i != 0)
#pragma unroll 1
  for (uint32_t i = 0; i < GTAB_RUN; i++) {
    uint32_t* e = e0 + (size_t)i * 24;
#pragma unroll
    for (int k = 0; k < 12; k++) {
      e[k] = p.x.l[k];
      e[12 + k] = p.y.l[k];
    }
    zs[i] = p.z;
    pre[i] = i ? mul(pre[i - 1], p.z) : p.z;
    if (i + 1 < GTAB_RUN) p = g1_add_mixed_nz(p, dx, dy);
  }
  Fp inv = fp_inv_bgcd(pre[GTAB_RUN - 1]);
#pragma unroll 1
  for (int i = (int)GTAB_RUN - 1; i >= 0; i--) {
    Fp zi = i ? mul(inv, pre[i - 1]) : inv;
    inv = mul(inv, zs[i]);
    uint32_t* e = e0 + (size_t)i * 24;
    Fp x, y;
#pragma unroll
    for (int k = 0; k < 12; k++) {
      x.l[k] = e[k];
      y.l[k] = e[12 + k];
    }
    x = mul(x, zi);
    y = mul(y, zi);
#pragma unroll
    for (int k = 0; k < 12; k++) {
      e[k] = x.l[k];
      e[12 + k] = y.l[k];
    }
  }
}

// u = the odd representative of s mod r (s < r: the callers range-check; for other 256-bit values the sum may wrap and the point is
// unspecified - still a memory-safe walk through the table)
DKGV_HD void gtab_scalar(uint32_t* u, const uint32_t* s_raw /*8 limbs*/) {
  const uint32_t m = (s_raw[0] & 1u) ? 0u : 0xffffffffu;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)s_raw[i] + (FrParams::mod(i) & m);
    u[i] = (uint32_t)c;
    c >>= 32;
  }
}
// index of the table entry for window w of the odd scalar u (gtab_scalar), *neg: the digit is negative
DKGV_HD uint32_t gtab_index(uint32_t bits, uint32_t windows, const uint32_t* u, uint32_t w, bool* neg) {
  const uint32_t bit = w * bits, limb = bit >> 5, sh = bit & 31;
  uint32_t v = u[limb] >> sh;
  if (sh + bits + 1 > 32 && limb + 1 < 8) v |= sh ? u[limb + 1] << (32 - sh) : 0u;
  uint32_t mag;
  if (w == windows - 1) {  // the last digit is what remains: positive (fewer than bits + 1 bits are left of u < 2^256)
    mag = (v & ((1u << bits) - 1u)) | 1u;
    *neg = false;
  } else {
    const uint32_t x = (v & ((2u << bits) - 1u)) | 1u, low = x & ((1u << bits) - 1u);
    const bool pos = (x >> bits) & 1u;
    mag = pos ? low : (1u << bits) - low;
    *neg = !pos;
  }
  return (w << (bits - 1)) + (mag >> 1);
}
// the entry of window w for u: x, and y with the digit's sign
DKGV_HD void gtab_lookup(const GTab& g, const uint32_t* u, uint32_t w, Fp* x, Fp* y) {
  bool ng;
  const uint32_t* e = g.p + (size_t)gtab_index(g.bits, g.windows, u, w, &ng) * 24;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x->l[i] = e[i];
    y->l[i] = e[12 + i];
  }
  if (ng) *y = neg(*y);
}

DKGV_HD G1Proj fixed_base_mul(const GTab& gtab, const uint32_t* s_raw /*8 limbs*/) {
  uint32_t u[8];
  gtab_scalar(u, s_raw);
  G1Aff first;
  first.inf = 0;
  gtab_lookup(gtab, u, 0, &first.x, &first.y);
  G1Proj acc = g1_from_affine(first);
#pragma unroll 1
  for (uint32_t w = 1; w < gtab.windows; w++) {
    Fp x, y;
    gtab_lookup(gtab, u, w, &x, &y);
    acc = g1_add_mixed_nz(acc, x, y);
  }
  return acc;
}

// One share: status of (dealer d, recipient id) - the body of verify_seed_exchange_commitment
// after the hash / lookup checks (crates/dkg/src/verification.rs:92-99,129-146).
DKGV_HD uint8_t share_check(const VVView& vv, uint32_t t, uint32_t d, uint32_t id, const uint8_t* secret_be,
                            GTab gtab, bool dealer_bad) {
  G1Proj ev = feldman_eval(vv, t, d, id);
  uint32_t s[8];
  bool in_range = fr_raw_from_be32(s, secret_be);
  G1Proj gs = fixed_base_mul(gtab, s);
  uint8_t st = g1_eq(ev, gs) ? DKGV_OK : DKGV_SLASHABLE_SHARE_MISMATCH;
  if (dealer_bad) st = DKGV_PANIC_BAD_G1;
  if (!in_range) st = DKGV_SLASHABLE_SECRET_RANGE;
  return st;
}

}  // namespace dkgv
