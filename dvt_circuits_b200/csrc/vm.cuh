// "Operand file" execution model for the hot kernels.
//
// Why: a G1 formula with every 381-bit Montgomery product inlined is ~4k SASS instructions; the
// Horner loop body (double, add, mixed add) came to 331 KB of code and ncu showed the warps stalled
// on instruction fetch 65 % of the time (profiles/r1_share_verify_v0_inlined.md).  Here every field
// routine exists ONCE (noinline), operands live in a per-thread file in shared memory, and a point
// formula is a short sequence of calls - the whole kernel is a few thousand instructions and stays
// resident in the instruction cache.  Register use drops to ~70, so occupancy is bounded by the
// operand file (13 slots x 48 B per thread) instead.
//
// Shared-memory layout: slot s of thread t is three 16-byte chunks at
//     file[(s*3 + c) * NT + t]            (NT = threads per block)
// so a warp's LDS.128/STS.128 touch 32 consecutive 16-byte words: conflict-free.
//
// The same code runs on the host (tests/hostemu) with the file in ordinary memory.
#pragma once
#include "feldman.cuh"

#if defined(__CUDACC__)
#define DKGV_NI static __host__ __device__ __noinline__
#else
#define DKGV_NI static
#endif

namespace dkgv {

struct alignas(16) U4 {
  uint32_t x, y, z, w;
};

struct OpFile {
  U4* base;         // this thread's chunk 0 of slot 0
  uint32_t stride;  // U4 elements between consecutive chunks (= threads per block)
};

DKGV_HD Fp of_load(const OpFile& f, int s) {
  Fp r;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    U4 v = f.base[(size_t)(s * 3 + c) * f.stride];
    r.l[4 * c] = v.x;
    r.l[4 * c + 1] = v.y;
    r.l[4 * c + 2] = v.z;
    r.l[4 * c + 3] = v.w;
  }
  return r;
}
DKGV_HD void of_store(const OpFile& f, int s, const Fp& a) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    U4 v;
    v.x = a.l[4 * c];
    v.y = a.l[4 * c + 1];
    v.z = a.l[4 * c + 2];
    v.w = a.l[4 * c + 3];
    f.base[(size_t)(s * 3 + c) * f.stride] = v;
  }
}

// ---- field routines on slots (one copy each) ---------------------------------------------
DKGV_NI void vm_mul(OpFile f, int d, int a, int b) { of_store(f, d, mul(of_load(f, a), of_load(f, b))); }
DKGV_NI void vm_add(OpFile f, int d, int a, int b) { of_store(f, d, add(of_load(f, a), of_load(f, b))); }
DKGV_NI void vm_sub(OpFile f, int d, int a, int b) { of_store(f, d, sub(of_load(f, a), of_load(f, b))); }
DKGV_NI void vm_mul12(OpFile f, int d, int a) { of_store(f, d, fp_mul12(of_load(f, a))); }
// d = a*b + c*e with one Montgomery reduction (field.cuh mul2add: 444 instead of 600 wide products).
// The instruction cache holds ~3.4k instructions of a hot loop (profiles/r1_fd_kernels.md: a 4.6k-instruction
// recombination kernel fell to a 77 % hit rate and lost 18 %), so the two product routines exist exactly once and
// everything else (sums of operands, negation, constants) goes through the small routines and a spare slot.
DKGV_NI void vm_mul2add(OpFile f, int d, int a, int b, int c, int e) {
  of_store(f, d, mul2add(of_load(f, a), of_load(f, b), of_load(f, c), of_load(f, e)));
}
DKGV_NI void vm_copy3(OpFile f, int d, int a) {
#pragma unroll
  for (int i = 0; i < 9; i++) f.base[(size_t)(d * 3 + i) * f.stride] = f.base[(size_t)(a * 3 + i) * f.stride];
}
DKGV_NI bool vm_eq(OpFile f, int a, int b) { return eq(of_load(f, a), of_load(f, b)); }
DKGV_NI void vm_neg(OpFile f, int d, int a) { of_store(f, d, neg(of_load(f, a))); }

// ---- slot map --------------------------------------------------------------------------------
enum : int { AX = 0, AY, AZ, BX, BY, BZ, T0, T1, T2, T3, T4, T5, T6, VM_SLOTS };

DKGV_HD void vm_set_point(const OpFile& f, int s, const G1Proj& p) {
  of_store(f, s, p.x);
  of_store(f, s + 1, p.y);
  of_store(f, s + 2, p.z);
}
DKGV_HD G1Proj vm_get_point(const OpFile& f, int s) {
  G1Proj p;
  p.x = of_load(f, s);
  p.y = of_load(f, s + 1);
  p.z = of_load(f, s + 2);
  return p;
}

// A <- A + B, both projective (RCB Alg. 7, complete; B may be the identity (0:1:0)); the three output lines are
// sums of two products each: 6 products + 3 fused pairs = 3132 instead of 3600 wide multiply-accumulates
DKGV_NI void vm_g1_add(OpFile f) {
  vm_mul(f, T0, AX, BX);
  vm_mul(f, T1, AY, BY);
  vm_mul(f, T2, AZ, BZ);
  vm_add(f, T3, AX, AY);
  vm_add(f, T4, BX, BY);
  vm_mul(f, T3, T3, T4);
  vm_add(f, T4, T0, T1);
  vm_sub(f, T3, T3, T4);
  vm_add(f, T4, AY, AZ);
  vm_add(f, T5, BY, BZ);
  vm_mul(f, T4, T4, T5);
  vm_add(f, T5, T1, T2);
  vm_sub(f, T4, T4, T5);
  vm_add(f, T5, AX, AZ);
  vm_add(f, T6, BX, BZ);
  vm_mul(f, T5, T5, T6);
  vm_add(f, T6, T0, T2);
  vm_sub(f, T5, T5, T6);  // "Y3" of the paper; A is dead from here on
  vm_add(f, AX, T0, T0);
  vm_add(f, T0, AX, T0);
  vm_mul12(f, T2, T2);
  vm_add(f, AZ, T1, T2);
  vm_sub(f, T1, T1, T2);
  vm_mul12(f, T5, T5);
  vm_neg(f, T2, T5);
  vm_mul2add(f, AX, T3, T1, T4, T2);  // X3 = t3 t1 - t4 y3
  vm_mul2add(f, AY, T1, AZ, T5, T0);  // Y3 = t1 z3 + y3 t0
  vm_mul2add(f, AZ, AZ, T4, T0, T3);  // Z3 = z3 t4 + t0 t3
}

// A <- 2A (RCB Alg. 9)
DKGV_NI void vm_g1_dbl(OpFile f) {
  vm_mul(f, T0, AY, AY);
  vm_add(f, T3, T0, T0);
  vm_add(f, T3, T3, T3);
  vm_add(f, T3, T3, T3);  // Z3' = 8 Y^2
  vm_mul(f, T1, AY, AZ);
  vm_mul(f, T2, AZ, AZ);
  vm_mul12(f, T2, T2);    // t2 = b3 Z^2
  vm_add(f, T5, T0, T2);  // Y3'
  vm_mul(f, T4, AX, AY);  // X Y
  vm_mul(f, AZ, T1, T3);  // Z3
  vm_add(f, T6, T2, T2);
  vm_add(f, T6, T6, T2);
  vm_sub(f, T0, T0, T6);               // t0 - 3 t2
  vm_mul2add(f, AY, T2, T3, T0, T5);   // Y3 = t2 Z3' + (t0 - 3 t2) Y3'
  vm_mul(f, T4, T0, T4);
  vm_add(f, AX, T4, T4);
}

// P <- P + Q with P projective at slots (p, p+1, p+2) and Q = (x, y) affine, never the identity,
// at slots (T5, T6) (RCB Alg. 8).  P may alias nothing in T0..T6.
DKGV_NI void vm_g1_madd(OpFile f, int p) {
  const int X1 = p, Y1 = p + 1, Z1 = p + 2, QX = T5, QY = T6;
  vm_mul(f, T0, X1, QX);
  vm_mul(f, T1, Y1, QY);
  vm_add(f, T3, QX, QY);
  vm_add(f, T4, X1, Y1);
  vm_mul(f, T3, T3, T4);
  vm_add(f, T4, T0, T1);
  vm_sub(f, T3, T3, T4);
  vm_mul(f, T4, QY, Z1);
  vm_add(f, T4, T4, Y1);
  vm_mul(f, T2, QX, Z1);
  vm_add(f, T2, T2, X1);  // "Y3" pre; Q, X1, Y1 dead
  vm_mul12(f, T5, Z1);    // t2
  vm_add(f, X1, T0, T0);
  vm_add(f, T0, X1, T0);
  vm_add(f, Z1, T1, T5);
  vm_sub(f, T1, T1, T5);
  vm_mul12(f, T2, T2);
  vm_neg(f, T5, T2);
  vm_mul2add(f, X1, T3, T1, T4, T5);  // X3 = t3 t1 - t4 y3
  vm_mul2add(f, Y1, T1, Z1, T2, T0);  // Y3 = t1 z3 + y3 t0
  vm_mul2add(f, Z1, Z1, T4, T0, T3);  // Z3 = z3 t4 + t0 t3
}

// Signed-digit chain for the small public scalar (recipient id): [k]P = sum d_i 2^i P with
// d_i in {-1, 0, +1}.  The non-adjacent form has ~1/3 non-zero digits instead of ~1/2, so
// fewer point additions (12 M each) at the price of at most one more doubling (8 M); whichever of
// plain binary and NAF is cheaper for this k is used.  Exact arithmetic either way - only the
// addition chain changes.  Warp-uniform: every lane of a warp shares k.
struct SmallChain {
  unsigned long long pos, neg;  // digit masks
  int top;                      // index of the leading (+1) digit; -1 for k == 0
};
DKGV_HD int chain_cost(unsigned long long pos, unsigned long long neg, int top) {
  int nz = 0;
  for (int i = 0; i <= top; i++) nz += (int)(((pos | neg) >> i) & 1);
  return 8 * top + 12 * (nz - 1);
}
DKGV_HD SmallChain make_small_chain(uint32_t k) {
  SmallChain c;
  c.pos = k;
  c.neg = 0;
  c.top = -1;
  if (k == 0) return c;
  int tb = 31;
  while (!((k >> tb) & 1)) tb--;
  c.top = tb;
  unsigned long long k3 = 3ull * k, kk = k;
  unsigned long long np = (k3 & ~kk) >> 1, nn = (~k3 & kk) >> 1;
  int tn = 33;
  while (!((np >> tn) & 1)) tn--;
  if (chain_cost(np, nn, tn) < chain_cost(c.pos, 0, tb)) {
    c.pos = np;
    c.neg = nn;
    c.top = tn;
  }
  return c;
}

// A <- [k]A along the chain; uses B as the base copy (BY is negated in place when the digit sign flips)
DKGV_HD void vm_g1_mul_chain(const OpFile& f, const SmallChain& c) {
  if (c.top < 0) {
    vm_set_point(f, AX, g1_identity());
    return;
  }
  if (((c.pos | c.neg) & ~(1ull << c.top)) != 0) vm_copy3(f, BX, AX);
  bool b_neg = false;
#pragma unroll 1
  for (int b = c.top - 1; b >= 0; b--) {
    vm_g1_dbl(f);
    bool p = (c.pos >> b) & 1, n = (c.neg >> b) & 1;
    if (p || n) {
      if (n != b_neg) {
        vm_neg(f, BY, BY);
        b_neg = n;
      }
      vm_g1_add(f);
    }
  }
}
DKGV_HD void vm_g1_mul_small(const OpFile& f, uint32_t k) { vm_g1_mul_chain(f, make_small_chain(k)); }

// coefficient k of dealer d -> projective point at slots (s, s+1, s+2); identity -> (0 : 1 : 0)
DKGV_HD void vm_load_coeff(const OpFile& f, int s, const VVView& v, uint32_t k, uint32_t d) {
  G1Aff a = vv_load(v, k, d);
  vm_set_point(f, s, g1_from_affine(a));
}

// Horner in the exponent (dkg_math.rs:160-174): A <- sum_k C_k id^k
DKGV_HD void vm_feldman_eval(const OpFile& f, const VVView& v, uint32_t t, uint32_t d, uint32_t id) {
  if (t == 0) {
    vm_set_point(f, AX, g1_identity());
    return;
  }
  vm_load_coeff(f, AX, v, t - 1, d);
  SmallChain chain = make_small_chain(id);
#pragma unroll 1
  for (int k = (int)t - 2; k >= 0; k--) {
    vm_g1_mul_chain(f, chain);
    vm_load_coeff(f, BX, v, (uint32_t)k, d);
    vm_g1_add(f);
  }
}

// B <- G * s  (s raw little-endian limbs; the caller range-checks it: feldman.cuh gtab_scalar)
DKGV_HD void vm_fixed_base_mul(const OpFile& f, GTab gtab, const uint32_t* s_raw) {
  uint32_t u[8];
  gtab_scalar(u, s_raw);
  G1Aff first;
  first.inf = 0;
  gtab_lookup(gtab, u, 0, &first.x, &first.y);
  vm_set_point(f, BX, g1_from_affine(first));
#pragma unroll 1
  for (uint32_t w = 1; w < gtab.windows; w++) {
    Fp x, y;
    gtab_lookup(gtab, u, w, &x, &y);
    of_store(f, T5, x);
    of_store(f, T6, y);
    vm_g1_madd(f, BX);
  }
}

// A == B as projective points (verification.rs:140 compares the compressed encodings)
DKGV_HD bool vm_g1_eq_ab(const OpFile& f) {
  bool ia = is_zero(of_load(f, AZ)), ib = is_zero(of_load(f, BZ));
  vm_mul(f, T0, AX, BZ);
  vm_mul(f, T1, BX, AZ);
  vm_mul(f, T2, AY, BZ);
  vm_mul(f, T3, BY, AZ);
  bool e = vm_eq(f, T0, T1) && vm_eq(f, T2, T3);
  return (ia || ib) ? (ia && ib) : e;
}

// one share (same contract as share_check in feldman.cuh)
DKGV_HD uint8_t vm_share_check(const OpFile& f, const VVView& vv, uint32_t t, uint32_t d, uint32_t id, const uint8_t* secret_be,
                               GTab gtab, bool dealer_bad) {
  vm_feldman_eval(f, vv, t, d, id);
  uint32_t s[8];
  bool in_range = fr_raw_from_be32(s, secret_be);
  vm_fixed_base_mul(f, gtab, s);
  uint8_t st = vm_g1_eq_ab(f) ? DKGV_OK : DKGV_SLASHABLE_SHARE_MISMATCH;
  if (dealer_bad) st = DKGV_PANIC_BAD_G1;
  if (!in_range) st = DKGV_SLASHABLE_SECRET_RANGE;
  return st;
}

}  // namespace dkgv
