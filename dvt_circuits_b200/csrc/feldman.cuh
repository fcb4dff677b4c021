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

// Offset fixed-base table for the generator (no zero digits, so the mixed addition never meets the
// identity and needs no select), GTAB_BITS-bit windows:
//   gtab[(w * 2^B + b) * 24 ..] = ((b + 1) * 2^(B w)) * G      b < 2^B, w < GTAB_WINDOWS   (affine Montgomery)
//   gtab[GTAB_WINDOWS * 2^B * 24 ..] = -(sum_w 2^(B w)) * G    (correction point)
// G*s = sum_w gtab[w][digit_w(s)] + correction.  B = 13: 20 windows, 21 mixed additions per multiplication instead of the 33 of byte
// windows; 163 841 entries = 15.7 MB, resident in the 126 MB L2 (the default share path does t fixed-base multiplications per
// dealer - 699 392 per (1024, 683) ceremony - and its x-half kernel is 2/3 of the step).
constexpr int GTAB_BITS = 13;
constexpr int GTAB_WINDOWS = (256 + GTAB_BITS - 1) / GTAB_BITS;
constexpr uint32_t GTAB_ENTRIES = GTAB_WINDOWS * (1u << GTAB_BITS) + 1;
constexpr size_t GTAB_WORDS = (size_t)GTAB_ENTRIES * 24;

DKGV_HD G1Aff gtab_entry(uint32_t idx) {
  if (idx < GTAB_WINDOWS * (1u << GTAB_BITS)) {
    uint32_t w = idx >> GTAB_BITS, b = idx & ((1u << GTAB_BITS) - 1u);
    G1Proj p = g1_from_affine(g1_generator());
#pragma unroll 1
    for (uint32_t i = 0; i < GTAB_BITS * w; i++) p = g1_dbl(p);
    return g1_to_affine(g1_mul_small(p, b + 1));
  }
  G1Proj g = g1_from_affine(g1_generator()), acc = g1_identity();
#pragma unroll 1
  for (uint32_t w = 0; w < GTAB_WINDOWS; w++) {
    acc = g1_add(acc, g);
#pragma unroll 1
    for (int i = 0; i < GTAB_BITS; i++) g = g1_dbl(g);
  }
  return g1_to_affine(g1_neg(acc));
}

// index of the table entry for window w of the 256-bit scalar s_raw (8 little-endian limbs); w == GTAB_WINDOWS: the correction point
DKGV_HD uint32_t gtab_index(const uint32_t* s_raw, int w) {
  if (w >= GTAB_WINDOWS) return (uint32_t)GTAB_WINDOWS << GTAB_BITS;
  const uint32_t bit = (uint32_t)w * GTAB_BITS, limb = bit >> 5, sh = bit & 31;
  uint32_t v = s_raw[limb] >> sh;
  if (sh + GTAB_BITS > 32 && limb + 1 < 8) v |= s_raw[limb + 1] << (32 - sh);
  return ((uint32_t)w << GTAB_BITS) + (v & ((1u << GTAB_BITS) - 1u));
}

DKGV_HD G1Proj fixed_base_mul(const uint32_t* gtab, const uint32_t* s_raw /*8 limbs*/) {
  G1Proj acc = g1_identity();
#pragma unroll 1
  for (int w = 0; w <= GTAB_WINDOWS; w++) {
    const uint32_t* e = gtab + (size_t)gtab_index(s_raw, w) * 24;
    Fp x, y;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      x.l[i] = e[i];
      y.l[i] = e[12 + i];
    }
    acc = g1_add_mixed_nz(acc, x, y);
  }
  return acc;
}

// One share: status of (dealer d, recipient id) - the body of verify_seed_exchange_commitment
// after the hash / lookup checks (crates/dkg/src/verification.rs:92-99,129-146).
DKGV_HD uint8_t share_check(const VVView& vv, uint32_t t, uint32_t d, uint32_t id, const uint8_t* secret_be,
                            const uint32_t* gtab, bool dealer_bad) {
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
