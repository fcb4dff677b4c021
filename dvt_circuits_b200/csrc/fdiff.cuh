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

// A <- f_d(x) for a small signed x (Horner with the signed-digit chain of |x|; [-x]P = -([x]P))
DKGV_HD void fd_seed_eval(const OpFile& f, const VVView& v, uint32_t t, uint32_t d, int32_t x) {
  if (t == 0) {
    vm_set_point(f, AX, g1_identity());
    return;
  }
  if (x == 0) {
    vm_load_coeff(f, AX, v, 0, d);
    return;
  }
  bool negx = x < 0;
  SmallChain chain = make_small_chain(negx ? (uint32_t)(-(int64_t)x) : (uint32_t)x);
  vm_load_coeff(f, AX, v, t - 1, d);
#pragma unroll 1
  for (int k = (int)t - 2; k >= 0; k--) {
    vm_g1_mul_chain(f, chain);
    if (negx) vm_neg(f, AY, AY);
    vm_load_coeff(f, BX, v, (uint32_t)k, d);
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

// compare the evaluation (an evals entry) with G * s: the tail of verify_seed_exchange_commitment
// (crates/dkg/src/verification.rs:92-99,138-146), same status contract as vm_share_check
DKGV_HD uint8_t fd_compare_item(const OpFile& f, const uint32_t* ent, uint32_t n_pad, const uint8_t* secret_be,
                                const uint32_t* gtab, bool dealer_bad) {
  uint32_t s[8];
  bool in_range = fr_raw_from_be32(s, secret_be);
  vm_fixed_base_mul(f, gtab, s);
  fd_load(f, AX, ent, n_pad);
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
  int32_t lo, hi;    // seed points lo..hi (hi - lo + 1 == t, lo <= 1 <= hi)
  uint32_t steps;    // extension steps: n_r - hi
  uint64_t cost_fd, cost_horner;  // field products per dealer, both ways (excluding G*s)
};

// ids must already be known to be a permutation of 1..n_r
inline FdPlan fd_make_plan(uint32_t t, uint32_t n_r) {
  FdPlan p{};
  p.use = false;
  for (uint32_t j = 1; j <= n_r; j++) p.cost_horner += fd_horner_cost(t, j);
  if (t < 2 || n_r <= t) return p;
  // window of t consecutive integers containing 1 with the cheapest Horner total + extension length
  uint64_t* pre = new uint64_t[t + 1];  // pre[a] = sum_{x=1..a} cost(x)
  pre[0] = 0;
  for (uint32_t a = 1; a <= t; a++) pre[a] = pre[a - 1] + fd_horner_cost(t, a);
  uint64_t best = ~0ull;
  for (int32_t lo = 2 - (int32_t)t; lo <= 1; lo++) {
    int32_t hi = lo + (int32_t)t - 1;
    uint64_t c = pre[hi] + (lo < 0 ? pre[-lo] : 0);
    c += (uint64_t)t * (t - 1) / 2 * 12 + (uint64_t)(n_r - (uint32_t)hi) * (t - 1) * 12;
    if (c < best) {
      best = c;
      p.lo = lo;
      p.hi = hi;
    }
  }
  delete[] pre;
  p.steps = n_r - (uint32_t)p.hi;
  p.cost_fd = best;
  p.use = best * 10 < p.cost_horner * 9;
  return p;
}

}  // namespace dkgv
