// Finite-difference evaluation of the Feldman commitments for a whole row of recipients.
//
// evaluate_polynomial (crates/dkg/src/dkg_math.rs:160-174) costs (t-1) * ~120 field products per
// (dealer, recipient) pair by Horner.  The recipient ids of a ceremony are always the ranks 1..n of
// the sorted commitment hashes (crates/dkg/src/verification.rs:50-66,129), i.e. CONSECUTIVE integers,
// and f_d(x) = sum_k C_{d,k} x^k has degree t-1 < n.  So only t values per dealer need Horner; every
// further value follows from the backward differences of the last seed point with t-1 point
// additions (12 products each) instead of a (t-1)-step Horner chain:
//
//   seeds       v[i] = f(lo + i), i = 0..t-1          (Horner, cheapest window of t integers around 0;
//                                                      [-x]P = -[x]P so negative points cost the same)
//   differences D[r] = nabla^r f(hi), r = 0..t-1      (t-1 rounds of pairwise subtractions, t^2/2 total)
//   extension   E_s[k] = nabla^k f(hi + s):  E_s[k] = E_{s-1}[k] + E_s[k+1],  E_s[t-1] = D[t-1]
//               run as a wavefront: at tick tau thread k computes step s = tau - (t-2-k)
//
// Splitting.  The seeds cost t Horner chains of t-1 steps each - quadratic in t.  Writing
//   f(x) = sum_{i<m} y^i f_i(x),   y = x^h mod r,   f_i(x) = sum_{k<h} C_{ih+k} x^k,   h = ceil(t/m)
// turns one dealer into m "virtual dealers" of degree h-1: m*h = t seeds of h-1 steps (m times less
// work), the same t-1 additions per extended point, and one joint double-and-add (Straus with GLV halves
// and width-4 signed digits of the public scalars y^i, shared by every dealer of the warp) per share.  [y]P = [x^h]P
// because every decoded commitment lies in the order-r subgroup (the decoder's subgroup check).
//
// Exact group arithmetic throughout (same complete RCB formulas), so every evaluation is the same
// projective-equivalent point Horner would give and the verdicts are bit-identical.
//
// Point planes in HBM: plane[(e * 36 + w) * n_pad + d] = w-th limb (X 0..11, Y 12..23, Z 24..35,
// Montgomery form, projective) of entry e for dealer d - a warp (32 consecutive dealers, one entry)
// moves 128 contiguous bytes per limb.
#pragma once
#include "vm.cuh"

namespace dkgv {

DKGV_HD const uint32_t* fd_entry(const uint32_t* plane, uint32_t n_pad, size_t e, uint32_t d) {
  return plane + e * 36 * (size_t)n_pad + d;
}
DKGV_HD uint32_t* fd_entry(uint32_t* plane, uint32_t n_pad, size_t e, uint32_t d) {
  return plane + e * 36 * (size_t)n_pad + d;
}

// entry -> the projective point at slots (s, s+1, s+2)
DKGV_HD void fd_load(const OpFile& f, int s, const uint32_t* ent, uint32_t n_pad) {
#pragma unroll
  for (int c = 0; c < 9; c++) {
    U4 v;
    v.x = ent[(size_t)(4 * c) * n_pad];
    v.y = ent[(size_t)(4 * c + 1) * n_pad];
    v.z = ent[(size_t)(4 * c + 2) * n_pad];
    v.w = ent[(size_t)(4 * c + 3) * n_pad];
    f.base[(size_t)(s * 3 + c) * f.stride] = v;
  }
}
DKGV_HD void fd_store(const OpFile& f, int s, uint32_t* ent, uint32_t n_pad) {
#pragma unroll
  for (int c = 0; c < 9; c++) {
    U4 v = f.base[(size_t)(s * 3 + c) * f.stride];
    ent[(size_t)(4 * c) * n_pad] = v.x;
    ent[(size_t)(4 * c + 1) * n_pad] = v.y;
    ent[(size_t)(4 * c + 2) * n_pad] = v.z;
    ent[(size_t)(4 * c + 3) * n_pad] = v.w;
  }
}

// coefficient k of dealer d; beyond the vector's end the padding coefficient is the identity
DKGV_HD void fd_load_coeff(const OpFile& f, int s, const VVView& v, uint32_t t, uint32_t k, uint32_t d) {
  if (k < t)
    vm_load_coeff(f, s, v, k, d);
  else
    vm_set_point(f, s, g1_identity());
}

// A <- sum_{k<h} C_{k0+k} x^k for a small signed x (Horner with the signed-digit chain of |x|;
// [-x]P = -([x]P)).  k0 = 0, h = t evaluates the whole polynomial.
DKGV_HD void fd_seed_eval(const OpFile& f, const VVView& v, uint32_t t, uint32_t d, int32_t x, uint32_t k0, uint32_t h) {
  if (h == 0) {
    vm_set_point(f, AX, g1_identity());
    return;
  }
  if (x == 0) {
    fd_load_coeff(f, AX, v, t, k0, d);
    return;
  }
  bool negx = x < 0;
  SmallChain chain = make_small_chain(negx ? (uint32_t)(-(int64_t)x) : (uint32_t)x);
  fd_load_coeff(f, AX, v, t, k0 + h - 1, d);
#pragma unroll 1
  for (int k = (int)h - 2; k >= 0; k--) {
    vm_g1_mul_chain(f, chain);
    if (negx) vm_neg(f, AY, AY);
    fd_load_coeff(f, BX, v, t, k0 + (uint32_t)k, d);
    vm_g1_add(f);
  }
}

// one item of difference round r >= 1:  dst[i] = src[i+1] - src[i]; the last item of the round
// (i == t-1-r) is nabla^r f(hi) and is also frozen into both copies of the extension state
DKGV_HD void fd_init_item(const OpFile& f, const uint32_t* src, uint32_t* dst, uint32_t* da, uint32_t* db, uint32_t n_pad,
                          uint32_t t, uint32_t r, uint32_t i, uint32_t d) {
  fd_load(f, AX, fd_entry(src, n_pad, i + 1, d), n_pad);
  fd_load(f, BX, fd_entry(src, n_pad, i, d), n_pad);
  vm_neg(f, BY, BY);
  vm_g1_add(f);
  fd_store(f, AX, fd_entry(dst, n_pad, i, d), n_pad);
  if (i == t - 1 - r) {
    fd_store(f, AX, fd_entry(da, n_pad, r, d), n_pad);
    fd_store(f, AX, fd_entry(db, n_pad, r, d), n_pad);
  }
}

// wavefront tick: item k (0 <= k <= t-2) advances to step s = tick - (t-2-k):
//   cur[k] = old[k] + old[k+1];  k == 0 yields f(hi + s), written to evals entry e_hi + s
DKGV_HD void fd_ext_item(const OpFile& f, const uint32_t* old, uint32_t* cur, uint32_t* evals, uint32_t n_pad, uint32_t t,
                         uint32_t tick, uint32_t k, size_t e_hi, uint32_t d) {
  fd_load(f, AX, fd_entry(old, n_pad, k, d), n_pad);
  fd_load(f, BX, fd_entry(old, n_pad, k + 1, d), n_pad);
  vm_g1_add(f);
  fd_store(f, AX, fd_entry(cur, n_pad, k, d), n_pad);
  if (k == 0) fd_store(f, AX, fd_entry(evals, n_pad, e_hi + (tick - (t - 2)), d), n_pad);
}

// band of active items at a tick (inclusive); empty when lo > hi
DKGV_HD void fd_ext_band(uint32_t t, uint32_t steps, uint32_t tick, int32_t* k_lo, int32_t* k_hi) {
  int32_t lo = (int32_t)t - 1 - (int32_t)tick, hi = (int32_t)t - 2 - (int32_t)tick + (int32_t)steps;
  *k_lo = lo < 0 ? 0 : lo;
  *k_hi = hi > (int32_t)t - 2 ? (int32_t)t - 2 : hi;
}

// ---- recombination of the m parts ---------------------------------------------------------------
// sum_i [k_i] P_i with k_i = y^i mod r (public, the same for every dealer) and P_i = f_i(x):
//  * GLV: z^2 P = -phi(P) = (beta X : -Y : Z) on G1 (z the curve parameter, r = z^4 - z^2 + 1; the same
//    endomorphism the subgroup check of g1.cuh uses), so k = k1 + k2 z^2 with k2 = k div z^2, k1 = k mod z^2
//    (both < 2^128, no lattice rounding needed) halves the doublings: [k]P = [k1]P + [k2](-phi(P));
//  * width-4 signed windows: digits in {+-1, +-3, +-5, +-7}, one non-zero digit in five on average, against
//    a per-share table {P, 3P, 5P, 7P} (phi is applied to the table entry on the fly: one product).
constexpr uint32_t FD_MAX_PARTS = 16;
constexpr uint32_t FD_GLV_DIGITS = 132;   // positions per half scalar (at most 130 used)
constexpr uint32_t FD_TAB_SLOTS = 4;      // per point: 2P (scratch), 3P, 5P, 7P
DKGV_HD size_t fd_dig_bytes(uint32_t m) { return (size_t)(m > 1 ? m - 1 : 1) * 2 * FD_GLV_DIGITS; }

// width-4 NAF of a value of at most 129 bits (5 limbs, destroyed); returns the top non-zero position or -1
DKGV_HD int fd_wnaf4(uint32_t* k, int8_t* out) {
  int top = -1;
  for (uint32_t b = 0; b < FD_GLV_DIGITS; b++) {
    int dgt = 0;
    if (k[0] & 1) {
      dgt = (int)(k[0] & 15);
      if (dgt >= 8) dgt -= 16;
      // k -= dgt
      uint64_t c;
      if (dgt > 0) {
        uint64_t br = (uint64_t)dgt;
        for (int w = 0; w < 5; w++) {
          uint64_t v = (uint64_t)k[w] - br;
          k[w] = (uint32_t)v;
          br = (v >> 32) & 1;
        }
      } else {
        c = (uint64_t)(-dgt);
        for (int w = 0; w < 5; w++) {
          c += k[w];
          k[w] = (uint32_t)c;
          c >>= 32;
        }
      }
      top = (int)b;
    }
    out[b] = (int8_t)dgt;
    for (int w = 0; w < 4; w++) k[w] = (k[w] >> 1) | (k[w + 1] << 31);
    k[4] >>= 1;
  }
  return top;
}

// k (canonical, 8 limbs, < r) -> k1 = k mod z^2, k2 = k div z^2 (5-limb buffers, top limb 0)
DKGV_HD void fd_glv_split(const uint32_t* k, uint32_t* k1, uint32_t* k2) {
  const uint32_t Z2[5] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u, 0u};  // z^2, z = -0xd201000000010000
  uint32_t rem[5] = {0, 0, 0, 0, 0};
  for (int w = 0; w < 5; w++) k2[w] = 0;
  for (int b = 255; b >= 0; b--) {
    for (int w = 4; w > 0; w--) rem[w] = (rem[w] << 1) | (rem[w - 1] >> 31);
    rem[0] = (rem[0] << 1) | ((k[b >> 5] >> (b & 31)) & 1);
    uint32_t tmp[5];
    uint64_t br = 0;
    for (int w = 0; w < 5; w++) {
      uint64_t v = (uint64_t)rem[w] - Z2[w] - br;
      tmp[w] = (uint32_t)v;
      br = (v >> 32) & 1;
    }
    if (!br) {
      for (int w = 0; w < 5; w++) rem[w] = tmp[w];
      if (b < 160) k2[b >> 5] |= 1u << (b & 31);  // b < 128 whenever the quotient bit is set (k < r < 2^255, z^2 > 2^127)
    }
  }
  for (int w = 0; w < 5; w++) k1[w] = rem[w];
}

// Signed digits of the recombination scalars y^i (i = 1..m-1), y = x^h mod r:
//   dig[((i-1)*2 + half) * FD_GLV_DIGITS + b], half 0 = k1 (acts on P_i), half 1 = k2 (acts on -phi(P_i)).
// Returns the highest non-zero position over all of them (-1 if none).
DKGV_HD int fd_comb_digits(uint32_t x, uint32_t h, uint32_t m, int8_t* dig) {
  Fr xm = zero<FrParams>();
  xm.l[0] = x;
  xm = to_mont(xm);
  Fr y = one<FrParams>();
  for (int b = 31; b >= 0; b--) {
    y = mul(y, y);
    if ((h >> b) & 1) y = mul(y, xm);
  }
  Fr p = y;
  int top = -1;
  for (uint32_t i = 1; i < m; i++) {
    Fr s = from_mont(p);
    uint32_t k1[5], k2[5];
    fd_glv_split(s.l, k1, k2);
    int t1 = fd_wnaf4(k1, dig + (size_t)((i - 1) * 2) * FD_GLV_DIGITS);
    int t2 = fd_wnaf4(k2, dig + (size_t)((i - 1) * 2 + 1) * FD_GLV_DIGITS);
    if (t1 > top) top = t1;
    if (t2 > top) top = t2;
    p = mul(p, y);
  }
  return top;
}

// slot <- beta (the cube root of unity of phi); a constant store instead of a third copy of the product routine
DKGV_NI void vm_set_beta(OpFile f, int d) {
  Fp beta;
#pragma unroll
  for (int i = 0; i < 12; i++) beta.l[i] = consts::BETA_M(i);
  of_store(f, d, beta);
}

// A <- sum_i [y^i] f_i(x): entries of the m virtual dealers (part i of dealer d sits at column
// i * n_pad + d of a plane n_padv = m * n_pad wide).  tab: this share's table plane, entry
// (i-1) * FD_TAB_SLOTS + slot for column d of a plane n_pad wide.  Control flow depends on (x, h, m)
// only - warp-uniform when all lanes share the recipient.
DKGV_HD void fd_combine_eval(const OpFile& f, const uint32_t* evals, uint32_t n_padv, uint32_t n_pad, uint32_t m, size_t e, uint32_t d,
                             const int8_t* dig, int top, uint32_t* tab) {
  // odd multiples 3P, 5P, 7P of every point that has a non-trivial scalar
  if (top >= 0) {
#pragma unroll 1
    for (uint32_t i = 1; i < m; i++) {
      const uint32_t* pi = fd_entry(evals, n_padv, e, i * n_pad + d);
      uint32_t* t2 = fd_entry(tab, n_pad, (size_t)(i - 1) * FD_TAB_SLOTS, d);
      fd_load(f, AX, pi, n_padv);
      vm_g1_dbl(f);
      fd_store(f, AX, t2, n_pad);
      fd_load(f, BX, pi, n_padv);
#pragma unroll 1
      for (uint32_t sl = 1; sl < FD_TAB_SLOTS; sl++) {
        vm_g1_add(f);
        fd_store(f, AX, fd_entry(tab, n_pad, (size_t)(i - 1) * FD_TAB_SLOTS + sl, d), n_pad);
        if (sl + 1 < FD_TAB_SLOTS) fd_load(f, BX, t2, n_pad);
      }
    }
  }
  bool started = false;
#pragma unroll 1
  for (int b = top; b >= 0; b--) {
    if (started) vm_g1_dbl(f);
#pragma unroll 1
    for (uint32_t ih = 0; ih < 2 * (m - 1); ih++) {
      int dg = dig[(size_t)ih * FD_GLV_DIGITS + b];
      if (dg == 0) continue;
      uint32_t i = 1 + (ih >> 1);
      bool half = ih & 1, ng = dg < 0;
      uint32_t a = (uint32_t)(ng ? -dg : dg);
      if (a == 1)
        fd_load(f, BX, fd_entry(evals, n_padv, e, i * n_pad + d), n_padv);
      else
        fd_load(f, BX, fd_entry(tab, n_pad, (size_t)(i - 1) * FD_TAB_SLOTS + (a >> 1), d), n_pad);
      if (half) {  // -phi(P) = (beta X : -Y : Z)
        vm_set_beta(f, T6);
        vm_mul(f, BX, BX, T6);
        ng = !ng;
      }
      if (ng) vm_neg(f, BY, BY);
      if (started) {
        vm_g1_add(f);
      } else {
        vm_copy3(f, AX, BX);
        started = true;
      }
    }
  }
  fd_load(f, BX, fd_entry(evals, n_padv, e, d), n_padv);
  if (started)
    vm_g1_add(f);
  else
    vm_copy3(f, AX, BX);
}

// recombine, then compare with G * s: the tail of verify_seed_exchange_commitment
// (crates/dkg/src/verification.rs:92-99,138-146), same status contract as vm_share_check
DKGV_HD uint8_t fd_combine_compare_item(const OpFile& f, const uint32_t* evals, uint32_t n_padv, uint32_t n_pad, uint32_t m, size_t e,
                                        uint32_t d, const int8_t* dig, int top, uint32_t* tab, const uint8_t* secret_be,
                                        GTab gtab, bool dealer_bad) {
  fd_combine_eval(f, evals, n_padv, n_pad, m, e, d, dig, top, tab);
  uint32_t s[8];
  bool in_range = fr_raw_from_be32(s, secret_be);
  vm_fixed_base_mul(f, gtab, s);
  uint8_t st = vm_g1_eq_ab(f) ? DKGV_OK : DKGV_SLASHABLE_SHARE_MISMATCH;
  if (dealer_bad) st = DKGV_PANIC_BAD_G1;
  if (!in_range) st = DKGV_SLASHABLE_SECRET_RANGE;
  return st;
}

// ---- plan (host side) ------------------------------------------------------------------------
// Field products executed by one Horner evaluation at |x| (chain + one full addition per step).
inline uint64_t fd_horner_cost(uint32_t t, uint32_t ax) {
  if (t < 2 || ax == 0) return 0;
  SmallChain c = make_small_chain(ax);
  return (uint64_t)(t - 1) * (uint64_t)(chain_cost(c.pos, c.neg, c.top) + 12);
}

// field products of one recombination: 128 shared doublings, per point a table (1 doubling + 3 additions),
// 2 * 128/5 additions and 128/5 products by beta; one more addition for part 0
inline uint64_t fd_comb_cost(uint32_t m) { return m > 1 ? 128 * 8 + (uint64_t)(m - 1) * (44 + 52 * 12 + 26) + 12 : 0; }

struct FdPlan {
  bool use;          // finite differences pay off for this shape
  uint32_t m, h;     // parts per dealer and coefficients per part (m * h >= t)
  int32_t lo, hi;    // seed points lo..hi (hi - lo + 1 == h, lo <= 1 <= hi)
  uint32_t steps;    // extension steps: n_r - hi
  uint64_t cost_fd, cost_horner;  // field products per dealer, both ways (excluding G*s)
};

// cheapest window of h consecutive integers containing 1 for polynomials of h coefficients that are then
// extended up to id n_r
inline uint64_t fd_best_window(uint32_t h, uint32_t n_r, int32_t* lo_out) {
  uint64_t* pre = new uint64_t[h + 1];  // pre[a] = sum_{x=1..a} cost(x)
  pre[0] = 0;
  for (uint32_t a = 1; a <= h; a++) pre[a] = pre[a - 1] + fd_horner_cost(h, a);
  uint64_t best = ~0ull;
  for (int32_t lo = 2 - (int32_t)h; lo <= 1; lo++) {
    int32_t hi = lo + (int32_t)h - 1;
    uint64_t c = pre[hi] + (lo < 0 ? pre[-lo] : 0);
    c += (uint64_t)h * (h - 1) / 2 * 12 + (uint64_t)(n_r > (uint32_t)hi ? n_r - (uint32_t)hi : 0) * (h - 1) * 12;
    if (c < best) {
      best = c;
      *lo_out = lo;
    }
  }
  delete[] pre;
  return best;
}

// ids must already be known to be a permutation of 1..n_r.  m_force != 0 fixes the number of parts.  n_opt: the
// number of ids the cost model assumes are evaluated in the group (0 = all n_r; t when the consistency shortcut of
// share_fd.cu is expected to settle the ids beyond t) - it shapes the choice of m and of the seed window only.
inline FdPlan fd_make_plan(uint32_t t, uint32_t n_r, uint32_t m_force = 0, uint32_t n_opt = 0) {
  if (n_opt == 0 || n_opt > n_r) n_opt = n_r;
  FdPlan p{};
  p.use = false;
  p.m = 1;
  p.h = t;
  p.cost_fd = ~0ull;
  for (uint32_t j = 1; j <= n_r; j++) p.cost_horner += fd_horner_cost(t, j);
  for (uint32_t m = 1; m <= FD_MAX_PARTS; m++) {
    if (m_force && m != m_force) continue;
    uint32_t h = (t + m - 1) / m;
    if (h < 2 || n_r <= h || (m > 1 && (uint64_t)(m - 1) * h >= t)) continue;  // every part must hold a coefficient
    int32_t lo = 1;
    uint64_t c = (uint64_t)m * fd_best_window(h, n_opt, &lo);
    if (m > 1) c += (uint64_t)n_opt * fd_comb_cost(m);
    if (c < p.cost_fd) {
      p.cost_fd = c;
      p.m = m;
      p.h = h;
      p.lo = lo;
      p.hi = lo + (int32_t)h - 1;
    }
  }
  if (p.cost_fd == ~0ull) return p;
  p.steps = n_r - (uint32_t)p.hi;
  p.use = p.cost_fd * 10 < p.cost_horner * 9;
  return p;
}

// ---- difference table of the consistency shortcut (share_fd.cu k_fd_difftab) ------------------------------------
// All values are CANONICAL residues mod r (no Montgomery form): the table needs subtractions and products by the small
// integers j < 2^10 only.
//
// prev - j * a mod r for canonical prev, a < r and j < 2^10: v = prev + j * (r - a) < 2^10 r < 2^265, quotient estimate
// q = floor(floor(v / 2^234) * floor(2^286 / r) / 2^52) in {floor(v / r) - 1, floor(v / r)} (both floors lose < 2^-20),
// v - q r < 2r by one chain against 2^256 - r, one conditional subtraction.
DKGV_HD Fr fr_submul_small(const Fr& prev, const Fr& a, uint32_t j) {
  constexpr uint32_t NEGR[8] = {0xffffffffu, 0x00000000u, 0x0001a401u, 0xac425bfdu, 0xf65e27fau, 0xccc627f7u, 0xd66282b7u, 0x8c1258acu};  // 2^256 - r
  constexpr uint32_t M = 0x8d54253bu;  // floor(2^286 / r)
  uint32_t na[8], v[8];
  uint64_t c = 0;
#pragma unroll
  for (int l = 0; l < 8; l++) {  // na = r - a in [1, r]
    uint64_t d = (uint64_t)FrParams::mod(l) - a.l[l] - c;
    na[l] = (uint32_t)d;
    c = (d >> 32) & 1;
  }
  c = 0;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    c += (uint64_t)na[l] * j + prev.l[l];
    v[l] = (uint32_t)c;
    c >>= 32;
  }
  uint32_t top = ((uint32_t)c << 22) | (v[7] >> 10);  // floor(v / 2^234) < 2^31
  uint32_t q = (uint32_t)(((uint64_t)top * M) >> 52);
  Fr w;
  c = 0;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    c += (uint64_t)NEGR[l] * q + v[l];
    w.l[l] = (uint32_t)c;
    c >>= 32;
  }
  cond_sub_mod<FrParams>(w.l, 0);
  return w;
}

// ---- lazy residues for the difference table ---------------------------------------------------------------------
// The table only ever subtracts neighbours (phase 1) and adds a small multiple (phase 2), so the entries need not be canonical
// after every step: a 9-limb (288-bit) value v with v == e (mod r) is carried instead, two's complement in phase 1 (|v| doubles
// per round at most), non-negative in phase 2 (v grows by a factor <= t per round), and brought back to [0, r) only every
// DT1_PERIOD rounds / every dt2_period(t) rounds - 9 subtract-with-borrow (or 9 multiply-adds) per entry and round instead of
// a modular subtraction (~30 instructions) / fr_submul_small (~80).
struct Lz {
  uint32_t l[9];
};
constexpr uint32_t DT1_PERIOD = 30;  // |v| < 2^255 * 2^30 = 2^285 before a reduction
// rounds between reductions in phase 2: v < 2^255 * t^period <= 2^285 (the round with factor j multiplies the bound by j + 1 <= t)
DKGV_HD uint32_t dt2_period(uint32_t t) {
  uint32_t lg = 1;
  while ((1u << lg) < t) lg++;
  uint32_t p = 30 / lg;
  return p ? p : 1;
}
DKGV_HD Lz lz_from(const Fr& a) {
  Lz r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = a.l[i];
  r.l[8] = 0;
  return r;
}
DKGV_HD Fr lz_low(const Lz& a) {  // a canonical (after lz_reduce)
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = a.l[i];
  return r;
}
DKGV_HD Lz lz_zero() {
  Lz r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = 0;
  return r;
}
DKGV_HD bool lz_is_zero(const Lz& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) o |= a.l[i];
  return o == 0;
}
// a - b mod 2^288
DKGV_HD Lz lz_sub(const Lz& a, const Lz& b) {
  Lz r;
#if defined(__CUDA_ARCH__)
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r.l[0]) : "r"(a.l[0]), "r"(b.l[0]));
#pragma unroll
  for (int i = 1; i < 8; i++) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r.l[i]) : "r"(a.l[i]), "r"(b.l[i]));
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r.l[8]) : "r"(a.l[8]), "r"(b.l[8]));
#else
  uint64_t br = 0;
  for (int i = 0; i < 9; i++) {
    uint64_t d = (uint64_t)a.l[i] - b.l[i] - br;
    r.l[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
#endif
  return r;
}
// prev + j * a mod 2^288 (j < 2^11).  Device: the products of the even limbs on one carry chain over prev, the products of the
// odd limbs on a second chain over that (the pattern of field.cuh's rows: lo / hi pairs become one wide multiply-add each).
DKGV_HD Lz lz_muladd_small(const Lz& prev, const Lz& a, uint32_t j) {
  Lz r;
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = prev.l[i];
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r.l[0]), "+r"(r.l[1]) : "r"(a.l[0]), "r"(j));
#pragma unroll
  for (int i = 2; i < 8; i += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r.l[i]), "+r"(r.l[i + 1]) : "r"(a.l[i]), "r"(j));
  asm volatile("madc.lo.u32 %0, %1, %2, %0;" : "+r"(r.l[8]) : "r"(a.l[8]), "r"(j));
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r.l[1]), "+r"(r.l[2]) : "r"(a.l[1]), "r"(j));
#pragma unroll
  for (int i = 3; i < 8; i += 2)
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r.l[i]), "+r"(r.l[i + 1]) : "r"(a.l[i]), "r"(j));
#else
  uint64_t c = 0;
  for (int i = 0; i < 9; i++) {
    c += (uint64_t)a.l[i] * j + prev.l[i];
    r.l[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return r;
}
// v -> the canonical residue of v mod r.  is_signed: v is two's complement with |v| <= 2^285; else 0 <= v <= 2^286.
// w = v + 2^31 r >= 0 (signed case), q = floor(floor(w / 2^224) * floor(2^286 / r) / 2^62) in [floor(w / r) - 2, floor(w / r)]
// (the two truncations lose < w / 2^286 + 2^-30 < 1.4), w - q r < 3r, two conditional subtractions.
DKGV_HD void lz_reduce(Lz& v, bool is_signed) {
  constexpr uint32_t R31[9] = {0x80000000u, 0x80000000u, 0x7fffffffu, 0x7fff2dffu, 0xa9ded201u, 0x04d0ec02u, 0x199cec04u, 0x94cebea4u, 0x39f6d3a9u};  // r << 31
  constexpr uint32_t M = 0x8d54253bu;  // floor(2^286 / r)
  if (is_signed) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      c += (uint64_t)v.l[i] + R31[i];
      v.l[i] = (uint32_t)c;
      c >>= 32;
    }
  }
  uint64_t lo = (uint64_t)v.l[7] * M, hi = (uint64_t)v.l[8] * M;
  uint32_t q = (uint32_t)((hi + (lo >> 32)) >> 30);
  uint32_t qr[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)FrParams::mod(i) * q;
    qr[i] = (uint32_t)c;
    c >>= 32;
  }
  qr[8] = (uint32_t)c;
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint64_t d = (uint64_t)v.l[i] - qr[i] - br;
    v.l[i] = (uint32_t)d;
    br = (d >> 32) & 1;
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    uint32_t t[9];
    br = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      uint64_t d = (uint64_t)v.l[i] - (i < 8 ? FrParams::mod(i) : 0u) - br;
      t[i] = (uint32_t)d;
      br = (d >> 32) & 1;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) v.l[i] = br ? v.l[i] : t[i];
  }
}
// r - a for canonical a (0 stays 0)
DKGV_HD Fr fr_neg_canonical(const Fr& a) {
  Fr r;
  uint64_t br = 0;
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)FrParams::mod(i) - a.l[i] - br;
    r.l[i] = (uint32_t)d;
    br = (d >> 32) & 1;
    nz |= a.l[i];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = nz ? r.l[i] : 0u;
  return r;
}
// the published copy of an entry: 9 words of thread i in three planes of a block-wide array (two 16-byte chunks + one word,
// conflict-free vector accesses); `pub` has 9 * nt words
DKGV_HD void lz_publish(uint32_t* pub, uint32_t nt, uint32_t i, const Lz& v) {
#if defined(__CUDA_ARCH__)
  ((uint4*)pub)[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  ((uint4*)(pub + 4 * (size_t)nt))[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
#else
  for (int k = 0; k < 4; k++) pub[4 * (size_t)i + k] = v.l[k];
  for (int k = 0; k < 4; k++) pub[4 * (size_t)nt + 4 * (size_t)i + k] = v.l[4 + k];
#endif
  pub[8 * (size_t)nt + i] = v.l[8];
}
DKGV_HD Lz lz_published(const uint32_t* pub, uint32_t nt, uint32_t i) {
  Lz v;
#if defined(__CUDA_ARCH__)
  uint4 x = ((const uint4*)pub)[i], y = ((const uint4*)(pub + 4 * (size_t)nt))[i];
  v.l[0] = x.x; v.l[1] = x.y; v.l[2] = x.z; v.l[3] = x.w;
  v.l[4] = y.x; v.l[5] = y.y; v.l[6] = y.z; v.l[7] = y.w;
#else
  for (int k = 0; k < 4; k++) v.l[k] = pub[4 * (size_t)i + k];
  for (int k = 0; k < 4; k++) v.l[4 + k] = pub[4 * (size_t)nt + 4 * (size_t)i + k];
#endif
  v.l[8] = pub[8 * (size_t)nt + i];
  return v;
}

// One thread of the table owns the adjacent entries a = e[2i], b = e[2i+1]; per round it publishes b, waits for the block,
// and reads its left neighbour's b.
struct DtPair {
  Lz a, b;
};
// phase 1, round r = 1..t: e[k] <- e[k] - e[k-1] for k >= r (lazy, two's complement; canonical again after every DT1_PERIOD-th
// round and after dt1_finish).  After t rounds e[k] = Delta^k s(1) for k < t and Delta^t s(k - t + 1) for k >= t (zero for every
// such k <=> the n shares lie on a polynomial of degree < t).
DKGV_HD bool dt1_active(uint32_t i, uint32_t r) { return 2 * i + 1 >= r; }
DKGV_HD bool dt1_publishes(uint32_t i, uint32_t r) { return 2 * i + 2 >= r; }  // the right neighbour still updates its a
// `reduce`: r is a multiple of DT1_PERIOD (the caller counts); left = the left neighbour's published b (read only when 2i >= r, i > 0)
DKGV_HD void dt1_step_with(DtPair& p, uint32_t i, uint32_t r, bool reduce, const Lz& left) {
  Lz nb = lz_sub(p.b, p.a);
  if (2 * i >= r && i) p.a = lz_sub(p.a, left);  // (i = 0: e[0] never changes; 2i >= r >= 1 excludes it anyway)
  p.b = nb;  // dt1_active(i, r) holds
  if (reduce) {
    lz_reduce(p.a, true);
    lz_reduce(p.b, true);
  }
}
DKGV_HD bool dt1_needs_left(uint32_t i, uint32_t r) { return 2 * i >= r && i; }
DKGV_HD void dt1_step(DtPair& p, uint32_t i, uint32_t r, bool reduce, const uint32_t* pub, uint32_t nt) {
  dt1_step_with(p, i, r, reduce, dt1_needs_left(i, r) ? lz_published(pub, nt, i - 1) : lz_zero());
}
DKGV_HD void dt1_finish(DtPair& p) {  // entries that dropped out between two reductions are still lazy
  lz_reduce(p.a, true);
  lz_reduce(p.b, true);
}
// phase 2, round j = t-1 .. 1 on the coefficient vector of P <- P (x - j) + E_{j-1} (E_k = Delta^k s(1) / k!, deg P = t-1-j
// before the round): c[k] <- c[k-1] - j c[k], c[-1] := E_{j-1}.  Entries k > t - j are still zero.  Carried with alternating
// signs, d[k] = (-1)^(k + rounds done) c[k], the step is d[k] <- d[k-1] + j d[k] - additions only, so the lazy values stay
// non-negative; the injected value is dt2_injected(E, t, j - 1) and dt2_finish undoes the sign.
DKGV_HD bool dt2_active(uint32_t i, uint32_t j, uint32_t t) { return 2 * i <= t - j; }
DKGV_HD Fr dt2_signed_e(const Fr& e, uint32_t t, uint32_t k) { return ((t - 1 - k) & 1) ? fr_neg_canonical(e) : e; }  // what to store as E'[k]
// `reduce`: this is the dt2_period(t)-th round since the last reduction (the caller counts; the same for every thread);
// left = d[2i - 1] before the round: the left neighbour's published b, E'[j - 1] for thread 0
DKGV_HD void dt2_step_with(DtPair& p, uint32_t j, bool reduce, const Lz& left) {
  Lz nb = lz_muladd_small(p.a, p.b, j);
  p.a = lz_muladd_small(left, p.a, j);
  p.b = nb;
  if (reduce) {
    lz_reduce(p.a, false);
    lz_reduce(p.b, false);
  }
}
DKGV_HD void dt2_step(DtPair& p, uint32_t i, uint32_t j, bool reduce, const uint32_t* pub, uint32_t nt, const Fr* Es) {
  dt2_step_with(p, j, reduce, i ? lz_published(pub, nt, i - 1) : lz_from(Es[j - 1]));
}
DKGV_HD void dt2_finish(DtPair& p, uint32_t i, uint32_t t) {  // after the t - 1 rounds: c[k] = (-1)^(k + t - 1) d[k]
  lz_reduce(p.a, false);
  lz_reduce(p.b, false);
  if ((2 * i + t - 1) & 1) p.a = lz_from(fr_neg_canonical(lz_low(p.a)));
  if ((2 * i + 1 + t - 1) & 1) p.b = lz_from(fr_neg_canonical(lz_low(p.b)));
}

// ---- condition (3) against the COMPRESSED commitment (share_fd.cu k_fd_coefpoint / k_fd_coefsign) -----------------
// compress(G * p_k) == C_k without decompressing C_k (no square root): the encoding of a non-identity point is its canonical
// x plus the sign of y, so the bytes agree iff the flags are well-formed, x_C < p, x_C * Z == X and lex_largest(Y / Z) equals
// the sign flag.  An encoding that is not a point of the subgroup can never equal the encoding of G * p_k, so an agreeing
// commitment needs no curve or subgroup test either.  The x half runs right after the fixed-base multiplication (projectively);
// the sign half needs 1/Z and is batched over FD_SIGN_K coefficients per thread (one inversion per batch).
//
// B <- G * sc; true when flags and x of c48 agree with B.  y_out / z_out: what the sign half consumes - (0, 1) for an agreeing
// identity (its sign flag must be 0 = lex_largest(0)) and whenever the answer is already false (keeps the batch invertible).
DKGV_HD bool fd_coef_point(const OpFile& f, GTab gtab, const uint32_t* sc, const uint8_t* c48, Fp* y_out, Fp* z_out) {
  vm_fixed_base_mul(f, gtab, sc);
  uint8_t b[48];
#pragma unroll
  for (int i = 0; i < 48; i++) b[i] = c48[i];
  const bool fc = (b[0] >> 7) & 1, fi = (b[0] >> 6) & 1, fs = (b[0] >> 5) & 1;
  b[0] &= 0x1f;
  Fp xr;
  fp_raw_from_be48(xr.l, b);
  Fp bz = of_load(f, BZ);
  const bool binf = is_zero(bz);
  *y_out = zero<FpParams>();
  *z_out = one<FpParams>();
  if (!fc || !raw_lt_mod<FpParams>(xr.l)) return false;
  if (fi) return !fs && is_zero(xr) && binf;
  if (binf) return false;
  Fp r2;
#pragma unroll
  for (int i = 0; i < 12; i++) r2.l[i] = FpParams::r2(i);
  of_store(f, T0, xr);
  of_store(f, T1, r2);
  vm_mul(f, T0, T0, T1);  // x_C in Montgomery form
  vm_mul(f, T2, T0, BZ);
  if (!vm_eq(f, T2, BX)) return false;
  *y_out = of_load(f, BY);
  *z_out = bz;
  return true;
}

// sign halves of cnt <= K points: lex_largest(y[i] / z[i]) == fs[i] for every i; z[i] != 0 (Montgomery's simultaneous inversion)
constexpr int FD_SIGN_K = 8;
template <int K>
DKGV_HD bool fd_coef_signs(const Fp* z, const Fp* y, const uint8_t* fs, int cnt) {
  Fp pref[K];
  Fp acc = z[0];
  pref[0] = acc;
#pragma unroll 1
  for (int i = 1; i < cnt; i++) {
    acc = mul(acc, z[i]);
    pref[i] = acc;
  }
  Fp inv = fp_inv_bgcd(acc);
  bool ok = true;
#pragma unroll 1
  for (int i = cnt - 1; i >= 0; i--) {
    Fp zi = inv;
    if (i) {
      zi = mul(inv, pref[i - 1]);
      inv = mul(inv, z[i]);
    }
    ok &= fp_lex_largest(mul(y[i], zi)) == (fs[i] != 0);
  }
  return ok;
}

}  // namespace dkgv
