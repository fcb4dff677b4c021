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
// work), the same t-1 additions per extended point, and one joint double-and-add (Straus, NAF digits
// of the public scalars y^i shared by every dealer of the warp) per share to recombine.  [y]P = [x^h]P
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
constexpr uint32_t FD_MAX_PARTS = 16;
constexpr uint32_t FD_DIG_WORDS = 16;  // per (x, i): 8 words of +1 digits, 8 words of -1 digits (NAF, 256 positions)

// NAF digit masks of the public scalars y^i (i = 1..m-1), y = x^h mod r; returns the highest digit
// position over all i (-1 if every scalar is zero).  dig: (m-1) * FD_DIG_WORDS words.
DKGV_HD int fd_comb_digits(uint32_t x, uint32_t h, uint32_t m, uint32_t* dig) {
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
    uint32_t s3[9];
    uint64_t c = 0;
    for (int w = 0; w < 8; w++) {
      c += (uint64_t)s.l[w] * 3;
      s3[w] = (uint32_t)c;
      c >>= 32;
    }
    s3[8] = (uint32_t)c;
    uint32_t* pos = dig + (size_t)(i - 1) * FD_DIG_WORDS;
    uint32_t* neg = pos + 8;
    for (int w = 0; w < 8; w++) {
      uint32_t sw = s.l[w], sn = w < 7 ? s.l[w + 1] : 0u;
      uint32_t plo = s3[w] & ~sw, phi = s3[w + 1] & ~sn;  // bits of (3s & ~s), this word and the next
      uint32_t nlo = ~s3[w] & sw, nhi = ~s3[w + 1] & sn;
      pos[w] = (plo >> 1) | (phi << 31);
      neg[w] = (nlo >> 1) | (nhi << 31);
      uint32_t any = pos[w] | neg[w];
      for (int b = 31; b >= 0; b--)
        if ((any >> b) & 1) {
          if (32 * w + b > top) top = 32 * w + b;
          break;
        }
    }
    p = mul(p, y);
  }
  return top;
}

// A <- sum_i [y^i] f_i(x): entries of the m virtual dealers (part i of dealer d sits at column
// i * n_pad + d of a plane n_padv = m * n_pad wide), joint double-and-add over the NAF digits.
// Control flow depends on (x, h, m) only - warp-uniform when all lanes share the recipient.
DKGV_HD void fd_combine_eval(const OpFile& f, const uint32_t* evals, uint32_t n_padv, uint32_t n_pad, uint32_t m, size_t e,
                             uint32_t d, const uint32_t* dig, int top) {
  bool started = false;
#pragma unroll 1
  for (int b = top; b >= 0; b--) {
    if (started) vm_g1_dbl(f);
#pragma unroll 1
    for (uint32_t i = 1; i < m; i++) {
      const uint32_t* pos = dig + (size_t)(i - 1) * FD_DIG_WORDS;
      bool p = (pos[b >> 5] >> (b & 31)) & 1, n = (pos[8 + (b >> 5)] >> (b & 31)) & 1;
      if (!(p || n)) continue;
      fd_load(f, BX, fd_entry(evals, n_padv, e, i * n_pad + d), n_padv);
      if (n) vm_neg(f, BY, BY);
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
                                        uint32_t d, const uint32_t* dig, int top, const uint8_t* secret_be, const uint32_t* gtab,
                                        bool dealer_bad) {
  fd_combine_eval(f, evals, n_padv, n_pad, m, e, d, dig, top);
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

struct FdPlan {
  bool use;          // finite differences pay off for this shape
  uint32_t m, h;     // parts per dealer and coefficients per part (m * h >= t)
  int32_t lo, hi;    // seed points lo..hi (hi - lo + 1 == h, lo <= 1 <= hi)
  uint32_t steps;    // extension steps: n_r - hi
  uint64_t cost_fd, cost_horner;  // field products per dealer, both ways (excluding G*s)
};

// cheapest window of h consecutive integers containing 1 for polynomials of h coefficients
inline uint64_t fd_best_window(uint32_t h, uint32_t n_r, int32_t* lo_out) {
  uint64_t* pre = new uint64_t[h + 1];  // pre[a] = sum_{x=1..a} cost(x)
  pre[0] = 0;
  for (uint32_t a = 1; a <= h; a++) pre[a] = pre[a - 1] + fd_horner_cost(h, a);
  uint64_t best = ~0ull;
  for (int32_t lo = 2 - (int32_t)h; lo <= 1; lo++) {
    int32_t hi = lo + (int32_t)h - 1;
    uint64_t c = pre[hi] + (lo < 0 ? pre[-lo] : 0);
    c += (uint64_t)h * (h - 1) / 2 * 12 + (uint64_t)(n_r - (uint32_t)hi) * (h - 1) * 12;
    if (c < best) {
      best = c;
      *lo_out = lo;
    }
  }
  delete[] pre;
  return best;
}

// ids must already be known to be a permutation of 1..n_r.  m_force != 0 fixes the number of parts.
inline FdPlan fd_make_plan(uint32_t t, uint32_t n_r, uint32_t m_force = 0) {
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
    uint64_t c = (uint64_t)m * fd_best_window(h, n_r, &lo);
    if (m > 1) c += (uint64_t)n_r * (255 * 8 + (uint64_t)(m - 1) * 85 * 12 + 12);  // joint NAF double-and-add per share
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

}  // namespace dkgv
