// Share-matrix verification through finite differences (fdiff.cuh): the row-of-recipients form of
// verify_seed_exchange_commitment (crates/dkg/src/verification.rs:129-146) when the recipient ids
// are the consecutive ranks 1..n (verification.rs:50-66).  Each dealer polynomial is split into m
// parts of h coefficients ("virtual dealers", column part * n_pad + d of planes m * n_pad wide).
// Four phases on one stream:
//   k_fd_seed     h Horner evaluations of h-1 steps per virtual dealer (LPT order)
//   k_fd_init     h-1 rounds of pairwise differences           (1 point addition per item)
//   k_fd_ext      steps + h - 2 wavefront ticks                (1 point addition per item)
//   k_fd_digits   signed digits of the public recombination scalars x^(h i) mod r, one thread per id
//   k_fd_combine  sum_i [y^i] f_i(x) by joint double-and-add, G * s, compare: one thread per share
// before them the consistency shortcut (k_fd_tables / k_fd_share_limbs / k_fd_difftab / k_fd_coefpoint / k_fd_coefsign /
// k_fd_need / k_fd_fill_ok, see below; k_fd_polycheck / k_fd_interp / k_fd_coefcheck are the formulations they replaced, kept
// behind DKGV_FD_DIFFTAB=0 / DKGV_FD_BYTES=0 for A/B tests): a dealer group whose shares are provably all valid never enters
// the evaluation, and its commitments are never decompressed.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ctx.hpp"
#include "fdiff.cuh"

using namespace dkgv;

int dkgv_session_redecode_checked(dkgv_ctx* ctx, cudaStream_t s);  // dkgv.cu

#ifndef DKGV_FD_BYTES_DEFAULT
#define DKGV_FD_BYTES_DEFAULT 1  // B200 parity run green (tests/test_gpu_share.py test_shortcut_formulations_agree); 0: lazy decode + k_fd_coefcheck
#endif
constexpr bool FD_BYTES_DEFAULT = DKGV_FD_BYTES_DEFAULT != 0;
static bool g_fd_bytes = FD_BYTES_DEFAULT;         // condition (3) against the compressed commitments, decode deferred (k_fd_coefpoint + k_fd_coefsign);
                                       // DKGV_FD_BYTES=0: lazy decode + k_fd_coefcheck (A/B tests)
static bool g_fd_difftab = true;      // fused difference table (k_fd_difftab); DKGV_FD_DIFFTAB=0: k_fd_polycheck + k_fd_interp (A/B tests)
static uint32_t g_fd_ipb_force = 0;  // items per block of the difference / extension launches, DKGV_FD_IPB (experiments)
constexpr int FD_NT = 32;  // one warp per block: 32 consecutive dealers, one entry (cf. SVM_NT in dkgv.cu)
constexpr size_t FD_SMEM = (size_t)VM_SLOTS * 3 * FD_NT * sizeof(U4);

__global__ void __launch_bounds__(FD_NT)
k_fd_seed(VVView vv, const int32_t* __restrict__ seed_x, int32_t lo, uint32_t* __restrict__ evals, uint32_t n_d, uint32_t t,
          uint32_t h, uint32_t part0, uint32_t n_padv, uint32_t d0, uint32_t n_cols, const uint8_t* __restrict__ need_group) {
  extern __shared__ U4 opfile[];
  if (need_group && !need_group[blockIdx.x]) return;  // only the dealer groups that need the evaluation
  uint32_t d = blockIdx.x * 32 + threadIdx.x;  // column of this dealer chunk; dealer d0 + d
  int32_t x = seed_x[blockIdx.y];              // most expensive points first
  uint32_t part = part0 + blockIdx.z;
  uint32_t dd = d0 + d < n_d ? d0 + d : n_d - 1;
  OpFile f{opfile + threadIdx.x, FD_NT};
  fd_seed_eval(f, vv, t, dd, x, part * h, h);
  fd_store(f, AX, fd_entry(evals, n_padv, (size_t)(x - lo), part * n_cols + d), n_padv);
}

__global__ void __launch_bounds__(FD_NT)
k_fd_init(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t* __restrict__ da, uint32_t* __restrict__ db,
          uint32_t n_pad, uint32_t t, uint32_t r, uint32_t col0, uint32_t ipb, const uint8_t* __restrict__ need_group, uint32_t groups) {
  extern __shared__ U4 opfile[];
  if (need_group && !need_group[(col0 / 32 + blockIdx.x) % groups]) return;
  uint32_t d = col0 + blockIdx.x * 32 + threadIdx.x;
  OpFile f{opfile + threadIdx.x, FD_NT};
#pragma unroll 1
  for (uint32_t i = blockIdx.y * ipb; i < (blockIdx.y + 1) * ipb && i + r < t; i++) fd_init_item(f, src, dst, da, db, n_pad, t, r, i, d);
}

__global__ void __launch_bounds__(FD_NT)
k_fd_ext(const uint32_t* __restrict__ old, uint32_t* __restrict__ cur, uint32_t* __restrict__ evals, uint32_t n_pad, uint32_t t,
         uint32_t tick, uint32_t k_lo, uint32_t k_hi, size_t e_hi, uint32_t col0, uint32_t ipb, const uint8_t* __restrict__ need_group,
         uint32_t groups) {
  extern __shared__ U4 opfile[];
  if (need_group && !need_group[(col0 / 32 + blockIdx.x) % groups]) return;  // continuation only for dealer groups that need it
  uint32_t d = col0 + blockIdx.x * 32 + threadIdx.x;
  OpFile f{opfile + threadIdx.x, FD_NT};
#pragma unroll 1
  for (uint32_t k = k_lo + blockIdx.y * ipb; k < k_lo + (blockIdx.y + 1) * ipb && k <= k_hi; k++)
    fd_ext_item(f, old, cur, evals, n_pad, t, tick, k, e_hi, d);
}

__global__ void __launch_bounds__(128)
k_fd_digits(uint32_t n_r, uint32_t h, uint32_t m, int8_t* __restrict__ dig, int32_t* __restrict__ top) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (x > n_r) return;
  top[x - 1] = fd_comb_digits(x, h, m, dig + (size_t)(x - 1) * fd_dig_bytes(m));
}

// recipients: columns j0 .. j0 + gridDim.y - 1, or (cols != nullptr) cols[j0 ..] = columns in ascending-id order;
// tab holds one table plane per (tab_r0 + blockIdx.y, point, slot)
__global__ void __launch_bounds__(FD_NT)
k_fd_combine(const uint32_t* __restrict__ evals, int32_t lo, uint32_t m, const int8_t* __restrict__ dig, const int32_t* __restrict__ top,
             const uint32_t* __restrict__ ids, const uint8_t* __restrict__ shares, const uint32_t* __restrict__ gtab,
             const uint8_t* __restrict__ dealer_bad, uint8_t* __restrict__ status, uint32_t* __restrict__ tab, uint32_t n_pad,
             uint32_t n_d, uint32_t n_r, uint32_t j0, const uint32_t* __restrict__ cols, uint32_t tab_r0, uint32_t d0,
             const uint8_t* __restrict__ need_group) {
  extern __shared__ U4 opfile[];
  if (need_group && !need_group[blockIdx.x]) return;
  uint32_t d = blockIdx.x * 32 + threadIdx.x;  // column of this dealer chunk (n_pad columns); dealer d0 + d
  uint32_t j = cols ? cols[j0 + blockIdx.y] : j0 + blockIdx.y;
  bool active = d0 + d < n_d;
  uint32_t dd = active ? d : n_d - 1 - d0;  // column whose data an inactive lane borrows
  uint32_t gd = d0 + dd;
  OpFile f{opfile + threadIdx.x, FD_NT};
  uint32_t x = ids[j];
  size_t e = (size_t)((int64_t)x - lo);
  uint32_t* my_tab = tab + (size_t)(tab_r0 + blockIdx.y) * (m - 1) * FD_TAB_SLOTS * 36 * n_pad;  // column d, not dd: private to this thread
  uint8_t st = fd_combine_compare_item(f, evals, n_pad * m, n_pad, m, e, dd, dig + (size_t)(x - 1) * fd_dig_bytes(m),
                                       m > 1 ? top[x - 1] : -1, my_tab + (d - dd), shares + ((size_t)gd * n_r + j) * 32, gtab,
                                       dealer_bad[gd] != 0);
  if (active) status[(size_t)gd * n_r + j] = st;
}

// ---- consistency shortcut -------------------------------------------------------------------------------
// All n shares of a dealer are valid  <=>  (1) every share is < r, (2) the shares s(1..n) lie on a polynomial p of
// degree <= t-1 over Fr, and (3) G * p_k == C_k for every coefficient k.  (The commitments C_k = a_k G define
// A(x) = sum a_k x^k with f(x) = A(x) G; (3) says A = p, so s(x) = p(x) = A(x) for every id; the converse is
// trivial.)  (1) and (2) are scalar arithmetic - (2): the t-th forward differences sum_j (-1)^j C(t,j) s(x+j) vanish
// for x = 1..n-t - p is the Newton interpolant of s(1..t) converted to monomial coefficients, and (3) costs t
// fixed-base multiplications per dealer instead of n evaluations in the exponent.  A group of 32 dealers in which
// some dealer fails a condition (or has an undecodable commitment) goes through the full evaluation, which yields
// the exact per-share verdicts.  Exact and deterministic - no random linear combination.
// Default formulation: (2) and the interpolation in one difference table per dealer (k_fd_difftab), (3) against the
// compressed commitments (k_fd_coefpoint + k_fd_coefsign) with the decode deferred until a group needs the evaluation.
constexpr uint32_t FD_SHORTCUT_MAX_T = 1024;  // k_fd_interp: one thread per coefficient; fr_submul_small: factors j < 2^10 (dkgv.cu's lazy_subgroup uses the same bound)

// c[j] = (-1)^j C(t, j) mod r (j = 0..t), inv[j] = 1/j mod r (j = 1..t) and ifact[j] = 1/j! mod r, Montgomery form; one thread per j
__global__ void __launch_bounds__(128) k_fd_tables(uint32_t t, uint32_t* __restrict__ c, uint32_t* __restrict__ inv, uint32_t* __restrict__ ifact) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > t) return;
  Fr num = one<FrParams>(), den = one<FrParams>(), jm = zero<FrParams>();
  for (uint32_t i = 1; i <= j; i++) {
    Fr a = zero<FrParams>(), b = zero<FrParams>();
    a.l[0] = t - i + 1;
    b.l[0] = i;
    num = mul(num, to_mont(a));
    den = mul(den, to_mont(b));
  }
  jm.l[0] = j ? j : 1;
  jm = to_mont(jm);
  Fr iden = one<FrParams>(), ij = one<FrParams>();  // den^(r-2), j^(r-2)
  for (int l = 7; l >= 0; l--) {
    uint32_t w = consts::R_MINUS_2(l);
    for (int b = 31; b >= 0; b--) {
      iden = mul(iden, iden);
      ij = mul(ij, ij);
      if ((w >> b) & 1) {
        iden = mul(iden, den);
        ij = mul(ij, jm);
      }
    }
  }
  Fr v = mul(num, iden);
  if (j & 1) v = neg(v);
  for (int l = 0; l < 8; l++) {
    c[(size_t)j * 8 + l] = v.l[l];
    inv[(size_t)j * 8 + l] = ij.l[l];
    ifact[(size_t)j * 8 + l] = iden.l[l];  // 1/j!
  }
}

// shares of one dealer chunk as little-endian limbs in ascending-id order: sl[d - d0][x - 1][8]; poly_ok[d] = 0 when a share is >= r
__global__ void __launch_bounds__(128)
k_fd_share_limbs(const uint8_t* __restrict__ shares, const uint32_t* __restrict__ cols, uint32_t* __restrict__ sl, uint8_t* __restrict__ poly_ok,
                 uint32_t d0, uint32_t n_cols, uint32_t n_d, uint32_t n_r) {
  uint32_t xi = blockIdx.x * blockDim.x + threadIdx.x, dl = blockIdx.y;
  if (xi >= n_r || d0 + dl >= n_d) return;
  uint32_t l[8];
  bool ok = fr_raw_from_be32(l, shares + ((size_t)(d0 + dl) * n_r + cols[xi]) * 32);
  if (!ok) poly_ok[d0 + dl] = 0;
  uint32_t* o = sl + ((size_t)dl * n_r + xi) * 8;
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = l[i];
}

// condition (2): one block per dealer, one thread per window x = w + 1; poly_ok[d] = 0 when a t-th difference is non-zero
__global__ void __launch_bounds__(256)
k_fd_polycheck(const uint32_t* __restrict__ sl, const uint32_t* __restrict__ c, uint8_t* __restrict__ poly_ok, uint32_t d0, uint32_t n_d,
               uint32_t n_r, uint32_t t) {
  uint32_t dl = blockIdx.x;
  if (d0 + dl >= n_d) return;
  const uint32_t* row = sl + (size_t)dl * n_r * 8;
  bool bad = false;
  for (uint32_t w = threadIdx.x; w + t < n_r; w += blockDim.x) {
    Fr acc = zero<FrParams>();
#pragma unroll 1
    for (uint32_t j = 0; j <= t; j++) {
      Fr sv, cv;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sv.l[i] = row[(size_t)(w + j) * 8 + i];
        cv.l[i] = c[(size_t)j * 8 + i];
      }
      acc = add(acc, mul(sv, cv));  // canonical share x Montgomery coefficient = canonical product
    }
    bad |= !is_zero(acc);
  }
  if (bad) poly_ok[d0 + dl] = 0;
}

// p = the polynomial of degree <= t-1 through (x, s(x)), x = 1..t, as canonical monomial coefficients coef[dl][k][8]:
// forward differences D_j = Delta^j s(1), then Horner in the Newton basis, P <- P (x - j) / j + D_{j-1}, on coefficient
// vectors: new_c[i] = c[i-1] / j - c[i].  One block per dealer, thread i owns coefficient i; 3 x t Fr values in shared memory.
__global__ void __launch_bounds__(1024)
k_fd_interp(const uint32_t* __restrict__ sl, const uint32_t* __restrict__ inv, uint32_t* __restrict__ coef, const uint8_t* __restrict__ poly_ok,
            uint32_t d0, uint32_t n_d, uint32_t n_r, uint32_t t) {
  extern __shared__ uint32_t fr_sm[];  // [3][t][8]
  uint32_t dl = blockIdx.x, i = threadIdx.x;
  if (d0 + dl >= n_d || !poly_ok[d0 + dl]) return;  // whole block; a dealer that already failed (1) or (2) needs no interpolation
  Fr* A = (Fr*)fr_sm;
  Fr* B = A + t;
  Fr* D = B + t;
  if (i < t) {
    Fr v;
#pragma unroll
    for (int l = 0; l < 8; l++) v.l[l] = sl[((size_t)dl * n_r + i) * 8 + l];
    A[i] = to_mont(v);
  }
  __syncthreads();
  Fr* cur = A;
  Fr* nxt = B;
  for (uint32_t r = 1; r < t; r++) {
    if (i < t) nxt[i] = i >= r ? sub(cur[i], cur[i - 1]) : cur[i];
    __syncthreads();
    Fr* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  if (i < t) D[i] = cur[i];
  __syncthreads();
  if (i < t) cur[i] = i == 0 ? D[t - 1] : zero<FrParams>();
  __syncthreads();
  for (uint32_t j = t - 1; j >= 1; j--) {
    if (i < t) {
      Fr ij;
#pragma unroll
      for (int l = 0; l < 8; l++) ij.l[l] = inv[(size_t)j * 8 + l];
      Fr v = i >= 1 ? sub(mul(cur[i - 1], ij), cur[i]) : neg(cur[0]);
      if (i == 0) v = add(v, D[j - 1]);
      nxt[i] = v;
    }
    __syncthreads();
    Fr* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  if (i < t) {
    Fr v = from_mont(cur[i]);
#pragma unroll
    for (int l = 0; l < 8; l++) coef[((size_t)dl * t + i) * 8 + l] = v.l[l];
  }
}

// Conditions (2) and the interpolation in ONE difference table per dealer (replaces k_fd_polycheck + k_fd_interp for
// n_r <= 2048), canonical residues, no products except by small integers (fdiff.cuh, "difference table"):
//   phase 1  t rounds e[k] <- e[k] - e[k-1] (k >= r) over all n_r shares: e[k] = Delta^k s(1) for k < t, and the entries
//            k >= t are the t-th differences Delta^t s(k-t+1) - all zero <=> condition (2);
//   phase 2  E_k = e[k] / k!, then P <- P (x - j) + E_{j-1} for j = t-1 .. 1 on monomial coefficients: c[k] <- c[k-1] - j c[k].
// One block per dealer, thread i owns entries 2i and 2i+1 in registers and publishes only e[2i+1] per round (double-buffered,
// one barrier per round); entries that are already final (phase 1: k < r) or still zero (phase 2: k > t - j) are skipped, so
// whole warps drop out.  Work per dealer: t n - t^2/2 subtractions + t^2/2 small products, against (n - t)(t + 1) + t^2/2 full
// Montgomery products before.
__global__ void __launch_bounds__(1024)
k_fd_difftab(const uint32_t* __restrict__ sl, const uint32_t* __restrict__ ifact, uint32_t* __restrict__ coef, uint8_t* __restrict__ poly_ok,
             uint32_t d0, uint32_t n_d, uint32_t n_r, uint32_t t) {
  extern __shared__ uint32_t fr_sm[];  // pub[2][blockDim.x], E[t]
  const uint32_t dl = blockIdx.x, i = threadIdx.x, nt = blockDim.x;
  if (d0 + dl >= n_d || !poly_ok[d0 + dl]) return;  // whole block; a share >= r already failed condition (1)
  Fr* pub = (Fr*)fr_sm;
  Fr* E = pub + 2 * (size_t)nt;
  const uint32_t k0 = 2 * i, k1 = 2 * i + 1;
  DtPair p;
  p.a = zero<FrParams>();
  p.b = zero<FrParams>();
  const uint32_t* row = sl + (size_t)dl * n_r * 8;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    if (k0 < n_r) p.a.l[l] = row[(size_t)k0 * 8 + l];
    if (k1 < n_r) p.b.l[l] = row[(size_t)k1 * 8 + l];
  }
#pragma unroll 1
  for (uint32_t r = 1; r <= t; r++) {
    Fr* pr = pub + (size_t)(r & 1) * nt;
    if (dt1_publishes(i, r)) pr[i] = p.b;
    __syncthreads();
    if (dt1_active(i, r)) dt1_step(p, i, r, pr);
  }
  bool bad = (k0 >= t && k0 < n_r && !is_zero(p.a)) || (k1 >= t && k1 < n_r && !is_zero(p.b));
  if (__syncthreads_or(bad)) {  // some t-th difference is not zero: the shares are not on a polynomial of degree < t
    if (i == 0) poly_ok[d0 + dl] = 0;
    return;
  }
  Fr f;
  if (k0 < t) {
#pragma unroll
    for (int l = 0; l < 8; l++) f.l[l] = ifact[(size_t)k0 * 8 + l];
    E[k0] = mul(p.a, f);  // canonical x Montgomery = canonical
  }
  if (k1 < t) {
#pragma unroll
    for (int l = 0; l < 8; l++) f.l[l] = ifact[(size_t)k1 * 8 + l];
    E[k1] = mul(p.b, f);
  }
  __syncthreads();
  p.a = i == 0 ? E[t - 1] : zero<FrParams>();
  p.b = zero<FrParams>();
#pragma unroll 1
  for (uint32_t j = t - 1; j >= 1; j--) {
    Fr* pr = pub + (size_t)(j & 1) * nt;
    const bool act = dt2_active(i, j, t);
    if (act) pr[i] = p.b;
    __syncthreads();
    if (act) dt2_step(p, i, j, pr, E);
  }
  uint32_t* o = coef + (size_t)dl * t * 8;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    if (k0 < t) o[(size_t)k0 * 8 + l] = p.a.l[l];
    if (k1 < t) o[(size_t)k1 * 8 + l] = p.b.l[l];
  }
}

// condition (3): G * p_k against the decoded commitment C_k; warp = 32 dealers x one k (the layout of the seeds)
__global__ void __launch_bounds__(FD_NT)
k_fd_coefcheck(VVView vv, const uint32_t* __restrict__ coef, const uint32_t* __restrict__ gtab, uint8_t* __restrict__ poly_ok, uint32_t d0,
               uint32_t n_d, uint32_t t) {
  extern __shared__ U4 opfile[];
  uint32_t dl = blockIdx.x * 32 + threadIdx.x, k = blockIdx.y;
  bool active = d0 + dl < n_d;
  uint32_t dc = active ? dl : n_d - 1 - d0;
#if defined(__CUDA_ARCH__)
  if (__ballot_sync(0xffffffffu, active && poly_ok[d0 + dl]) == 0) return;  // nobody in this group can still pass
#endif
  OpFile f{opfile + threadIdx.x, FD_NT};
  uint32_t sc[8];
#pragma unroll
  for (int l = 0; l < 8; l++) sc[l] = coef[((size_t)dc * t + k) * 8 + l];
  vm_fixed_base_mul(f, gtab, sc);  // B = G * p_k
  G1Aff c = vv_load(vv, k, d0 + dc);
  bool same;
  if (c.inf) {
    same = is_zero(of_load(f, BZ));
  } else {
    of_store(f, T0, c.x);
    of_store(f, T1, c.y);
    vm_mul(f, T2, T0, BZ);
    vm_mul(f, T3, T1, BZ);
    same = !is_zero(of_load(f, BZ)) && vm_eq(f, T2, BX) && vm_eq(f, T3, BY);
  }
  if (active && !same) poly_ok[d0 + dl] = 0;
}

// condition (3) WITHOUT decoding the commitments (fdiff.cuh, "against the COMPRESSED commitment"): x half.  Same launch shape
// as k_fd_coefcheck; writes Z (limbs 0..11) and Y (12..23) of G * p_k to the chunk-local planes yz[(k*24 + w) * n_cols + dl].
__global__ void __launch_bounds__(FD_NT)
k_fd_coefpoint(const uint8_t* __restrict__ vv, const uint32_t* __restrict__ coef, const uint32_t* __restrict__ gtab,
               uint8_t* __restrict__ poly_ok, uint32_t* __restrict__ yz, uint32_t d0, uint32_t n_d, uint32_t n_cols, uint32_t t) {
  extern __shared__ U4 opfile[];
  uint32_t dl = blockIdx.x * 32 + threadIdx.x, k = blockIdx.y;
  bool active = d0 + dl < n_d;
  uint32_t dc = active ? dl : n_d - 1 - d0;
#if defined(__CUDA_ARCH__)
  if (__ballot_sync(0xffffffffu, active && poly_ok[d0 + dl]) == 0) return;  // nobody in this group can still pass
#endif
  OpFile f{opfile + threadIdx.x, FD_NT};
  uint32_t sc[8];
#pragma unroll
  for (int l = 0; l < 8; l++) sc[l] = coef[((size_t)dc * t + k) * 8 + l];
  Fp y, z;
  bool same = fd_coef_point(f, gtab, sc, vv + ((size_t)(d0 + dc) * t + k) * 48, &y, &z);
  uint32_t* o = yz + (size_t)k * 24 * n_cols + dl;
#pragma unroll
  for (int w = 0; w < 12; w++) {
    o[(size_t)w * n_cols] = z.l[w];
    o[(size_t)(12 + w) * n_cols] = y.l[w];
  }
  if (active && !same) poly_ok[d0 + dl] = 0;
}

// sign half: thread = (dealer, batch of FD_SIGN_K consecutive coefficients), lanes = consecutive dealers (coalesced planes).
// A dealer still marked ok here had every one of its points written by k_fd_coefpoint (poly_ok only ever drops).
__global__ void __launch_bounds__(128)
k_fd_coefsign(const uint8_t* __restrict__ vv, const uint32_t* __restrict__ yz, uint8_t* __restrict__ poly_ok, uint32_t d0, uint32_t n_d,
              uint32_t n_cols, uint32_t t) {
  uint32_t dl = blockIdx.x * 32 + threadIdx.x, k0 = (blockIdx.y * blockDim.y + threadIdx.y) * FD_SIGN_K;
  if (d0 + dl >= n_d || k0 >= t || !poly_ok[d0 + dl]) return;
  int cnt = (int)(t - k0 < (uint32_t)FD_SIGN_K ? t - k0 : (uint32_t)FD_SIGN_K);
  Fp z[FD_SIGN_K], y[FD_SIGN_K];
  uint8_t fs[FD_SIGN_K];
#pragma unroll 1
  for (int i = 0; i < cnt; i++) {
    const uint32_t* e = yz + (size_t)(k0 + i) * 24 * n_cols + dl;
#pragma unroll
    for (int w = 0; w < 12; w++) {
      z[i].l[w] = e[(size_t)w * n_cols];
      y[i].l[w] = e[(size_t)(12 + w) * n_cols];
    }
    fs[i] = (vv[((size_t)(d0 + dl) * t + k0 + i) * 48] >> 5) & 1;
  }
  if (!fd_coef_signs<FD_SIGN_K>(z, y, fs, cnt)) poly_ok[d0 + dl] = 0;
}

// need_group[g] = 1 when some dealer of the 32-dealer group g (of this chunk) fails a condition or has an undecodable commitment
__global__ void __launch_bounds__(128)
k_fd_need(const uint8_t* __restrict__ poly_ok, const uint8_t* __restrict__ dealer_bad, uint32_t d0, uint32_t n_cols, uint32_t n_d,
          uint8_t* __restrict__ need_group, uint32_t* __restrict__ any_need) {
  uint32_t dl = blockIdx.x * blockDim.x + threadIdx.x;
  if (dl >= n_cols || d0 + dl >= n_d) return;
  if (dealer_bad[d0 + dl] != 0 || poly_ok[d0 + dl] == 0) {
    need_group[dl / 32] = 1;
    atomicOr(any_need, 1u);
  }
}

// verdict OK for every share of the dealer groups that met the three conditions
__global__ void __launch_bounds__(128)
k_fd_fill_ok(uint8_t* __restrict__ status, const uint8_t* __restrict__ need_group, uint32_t d0, uint32_t n_cols, uint32_t n_d, uint32_t n_r) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, dl = blockIdx.y;
  if (j >= n_r || d0 + dl >= n_d || need_group[dl / 32]) return;
  status[(size_t)(d0 + dl) * n_r + j] = DKGV_OK;
}

// evaluate_polynomial output instead of the share comparison: out[dealer][column j] = compress(f_d(ids[j]))
__global__ void __launch_bounds__(FD_NT)
k_fd_combine_out(const uint32_t* __restrict__ evals, int32_t lo, uint32_t m, const int8_t* __restrict__ dig, const int32_t* __restrict__ top,
                 const uint32_t* __restrict__ ids, uint8_t* __restrict__ out48, uint32_t* __restrict__ tab, uint32_t n_pad, uint32_t n_d,
                 uint32_t n_r, uint32_t j0, const uint32_t* __restrict__ cols, uint32_t tab_r0, uint32_t d0) {
  extern __shared__ U4 opfile[];
  uint32_t d = blockIdx.x * 32 + threadIdx.x;
  uint32_t j = cols ? cols[j0 + blockIdx.y] : j0 + blockIdx.y;
  bool active = d0 + d < n_d;
  uint32_t dd = active ? d : n_d - 1 - d0;
  OpFile f{opfile + threadIdx.x, FD_NT};
  uint32_t x = ids[j];
  size_t e = (size_t)((int64_t)x - lo);
  uint32_t* my_tab = tab + (size_t)(tab_r0 + blockIdx.y) * (m - 1) * FD_TAB_SLOTS * 36 * n_pad;
  fd_combine_eval(f, evals, n_pad * m, n_pad, m, e, dd, dig + (size_t)(x - 1) * fd_dig_bytes(m), m > 1 ? top[x - 1] : -1,
                  my_tab + (d - dd));
  uint8_t enc[48];
  g1_compress(g1_to_affine(vm_get_point(f, AX)), enc);
  if (active) {
    uint8_t* o = out48 + ((size_t)(d0 + d) * n_r + j) * 48;
#pragma unroll
    for (int i = 0; i < 48; i++) o[i] = enc[i];
  }
}

int dkgv_fd_setup(dkgv_ctx* ctx) {
  for (const void* k : {(const void*)k_fd_seed, (const void*)k_fd_init, (const void*)k_fd_ext, (const void*)k_fd_combine,
                        (const void*)k_fd_combine_out, (const void*)k_fd_coefcheck, (const void*)k_fd_coefpoint}) {
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FD_SMEM));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  if (const char* e = getenv("DKGV_FD_IPB")) g_fd_ipb_force = (uint32_t)atoi(e);
  CK(cudaFuncSetAttribute(k_fd_interp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * FD_SHORTCUT_MAX_T * 32)));
  CK(cudaFuncSetAttribute(k_fd_difftab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * 1024 + FD_SHORTCUT_MAX_T) * 32)));
  CK(cudaFuncSetAttribute(k_fd_difftab, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  size_t stack = 0;  // k_fd_coefsign keeps 3 x FD_SIGN_K field elements in local memory (1.2 KB frame); only ever raise the limit
  CK(cudaDeviceGetLimit(&stack, cudaLimitStackSize));
  if (stack < 2048) CK(cudaDeviceSetLimit(cudaLimitStackSize, 2048));
  const char* by = getenv("DKGV_FD_BYTES");
  g_fd_bytes = by ? atoi(by) != 0 : FD_BYTES_DEFAULT;
  const char* dt = getenv("DKGV_FD_DIFFTAB");
  g_fd_difftab = !dt || atoi(dt) != 0;
  for (int i = 0; i < 5; i++) CK(cudaEventCreate(&ctx->ev_fd[i]));
  CK(cudaEventCreateWithFlags(&ctx->fd_fork, cudaEventDisableTiming));
  for (uint32_t i = 0; i < FD_MAX_PARTS; i++) {
    CK(cudaStreamCreateWithFlags(&ctx->fd_streams[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->fd_join[i], cudaEventDisableTiming));
  }
  return 0;
}

// the share-matrix entry may skip the decode of the commitments until a dealer group needs the evaluation
bool dkgv_fd_defers_decode() { return g_fd_bytes; }

// ids (host copy) a permutation of 1..n_r ?
bool dkgv_fd_ids_consecutive(const uint32_t* h_ids, uint32_t n_r) {
  std::vector<uint8_t> seen(n_r, 0);
  for (uint32_t j = 0; j < n_r; j++) {
    uint32_t x = h_ids[j];
    if (x < 1 || x > n_r || seen[x - 1]) return false;
    seen[x - 1] = 1;
  }
  return true;
}

// Items (point additions) per block of the difference / extension launches.  One per block is the
// measured optimum on B200 (n=1024, t=683, N=1: extension 336 ms with 1, 344 ms with 2, 367 ms with 4 items):
// block scheduling is not what the one-addition blocks lose time on.  DKGV_FD_IPB overrides (experiments).
static inline uint32_t items_per_block(uint32_t, uint32_t) { return g_fd_ipb_force ? g_fd_ipb_force : 1; }

// dealers d0 .. d0 + n_pad - 1 (n_pad a multiple of 32: the column count of this chunk's planes)
static int fd_run(dkgv_ctx* ctx, const VVView& view, uint32_t d0, uint32_t n_pad, uint32_t n_d, uint32_t n_r, uint32_t t,
                  const FdPlan& plan, const uint32_t* d_ids, const uint32_t* h_ids, const uint8_t* d_shares, uint8_t* d_status,
                  uint8_t* d_out48, cudaStream_t s) {
  const uint32_t m = plan.m, h = plan.h;
  const uint32_t n_padv = n_pad * m;  // plane width: one column per virtual dealer
  const uint32_t groups = n_pad / 32, n_here = std::min(n_pad, n_d - d0);
  const size_t ent_words = (size_t)36 * n_padv, ent_bytes = ent_words * 4;
  const size_t n_evals = (size_t)((int64_t)n_r - plan.lo + 1);
  CK(ctx->fd_cols.reserve((size_t)n_r * 4));
  const uint32_t* cols = (const uint32_t*)ctx->fd_cols.p;
  ctx->fd_cols_host.resize(n_r);  // columns in ascending-id order
  for (uint32_t j = 0; j < n_r; j++) ctx->fd_cols_host[h_ids[j] - 1] = j;
  CK(cudaMemcpyAsync(ctx->fd_cols.p, ctx->fd_cols_host.data(), (size_t)n_r * 4, cudaMemcpyHostToDevice, s));
  CK(cudaEventRecord(ctx->ev_fd[0], s));
  CK(cudaEventRecord(ctx->ev_hot0, s));

  // ---- consistency shortcut: settle whole dealer groups by scalar arithmetic + t fixed-base multiplications
  const uint8_t* filter = nullptr;  // dealer groups that go through the evaluation (nullptr: all)
  ctx->fd_last_need = true;
  if (ctx->fd_polycheck && d_shares && !d_out48 && n_r > t && t <= FD_SHORTCUT_MAX_T) {
    CK(ctx->fd_sl.reserve((size_t)n_pad * n_r * 32));
    CK(ctx->fd_coef.reserve((size_t)n_pad * t * 32));
    CK(ctx->fd_flags.reserve((size_t)n_d + groups + 16));
    uint8_t* poly_ok = (uint8_t*)ctx->fd_flags.p;
    uint8_t* need_group = poly_ok + n_d;
    uint32_t* any_need = (uint32_t*)(((uintptr_t)(need_group + groups) + 7) & ~(uintptr_t)7);
    if (ctx->fd_binom_t != t) {
      CK(ctx->fd_binom.reserve((size_t)(t + 1) * 96));
      k_fd_tables<<<(t + 128) / 128, 128, 0, s>>>(t, (uint32_t*)ctx->fd_binom.p, (uint32_t*)ctx->fd_binom.p + (size_t)(t + 1) * 8,
                                                   (uint32_t*)ctx->fd_binom.p + (size_t)(t + 1) * 16);
      ctx->fd_binom_t = t;
      ctx->launches++;
    }
    const uint32_t* binom = (const uint32_t*)ctx->fd_binom.p;
    const uint32_t* invtab = binom + (size_t)(t + 1) * 8;
    CK(cudaMemsetAsync(poly_ok + d0, 1, n_here, s));
    CK(cudaMemsetAsync(need_group, 0, groups + 16, s));
    k_fd_share_limbs<<<dim3((n_r + 127) / 128, n_here), 128, 0, s>>>(d_shares, cols, (uint32_t*)ctx->fd_sl.p, poly_ok, d0, n_pad, n_d, n_r);
    const bool fused = g_fd_difftab && n_r <= 2048 && t >= 1;
    if (fused) {
      const uint32_t nt = (((n_r + 1) / 2 + 31) / 32) * 32;
      k_fd_difftab<<<n_here, nt, ((size_t)2 * nt + t) * 32, s>>>((const uint32_t*)ctx->fd_sl.p, invtab + (size_t)(t + 1) * 8,
                                                               (uint32_t*)ctx->fd_coef.p, poly_ok, d0, n_d, n_r, t);
    } else {
      k_fd_polycheck<<<n_here, 256, 0, s>>>((const uint32_t*)ctx->fd_sl.p, binom, poly_ok, d0, n_d, n_r, t);
      k_fd_interp<<<n_here, ((t + 31) / 32) * 32, (size_t)3 * t * 32, s>>>((const uint32_t*)ctx->fd_sl.p, invtab, (uint32_t*)ctx->fd_coef.p, poly_ok,
                                                                         d0, n_d, n_r, t);
    }
    CK(cudaEventRecord(ctx->ev_fd[1], s));  // shortcut phases: [limbs + difference table | x halves | sign halves | flags]
    const bool bytes = !ctx->vv_decoded;  // the decode was deferred: compare against the compressed commitments
    if (bytes) {
      CK(ctx->fd_yz.reserve((size_t)t * 24 * n_pad * 4));
      k_fd_coefpoint<<<dim3(groups, t), FD_NT, FD_SMEM, s>>>(ctx->vv_src, (const uint32_t*)ctx->fd_coef.p, ctx->gtab, poly_ok,
                                                             (uint32_t*)ctx->fd_yz.p, d0, n_d, n_pad, t);
      CK(cudaEventRecord(ctx->ev_fd[2], s));
      const uint32_t batches = (t + FD_SIGN_K - 1) / FD_SIGN_K;
      k_fd_coefsign<<<dim3(groups, (batches + 3) / 4), dim3(32, 4), 0, s>>>(ctx->vv_src, (const uint32_t*)ctx->fd_yz.p, poly_ok, d0, n_d,
                                                                         n_pad, t);
      ctx->launches++;
    } else {
      k_fd_coefcheck<<<dim3(groups, t), FD_NT, FD_SMEM, s>>>(view, (const uint32_t*)ctx->fd_coef.p, ctx->gtab, poly_ok, d0, n_d, t);
      CK(cudaEventRecord(ctx->ev_fd[2], s));
    }
    CK(cudaEventRecord(ctx->ev_fd[3], s));
    k_fd_need<<<(n_pad + 127) / 128, 128, 0, s>>>(poly_ok, (const uint8_t*)ctx->dealer_bad.p, d0, n_pad, n_d, need_group, any_need);
    k_fd_fill_ok<<<dim3((n_r + 127) / 128, n_here), 128, 0, s>>>(d_status, need_group, d0, n_pad, n_d, n_r);
    ctx->launches += fused ? 5 : 6;
    uint32_t h_any = 0;
    CK(cudaMemcpyAsync(&h_any, any_need, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (!h_any) {  // an honest chunk: every verdict is OK and already written
      ctx->fd_last_need = false;
      CK(cudaEventRecord(ctx->ev_hot1, s));
      CK(cudaEventRecord(ctx->ev_fd[4], s));
      ctx->hot_recorded = true;
      ctx->fd_recorded = true;
      return 0;
    }
    filter = need_group;
  }
  // the evaluation (and its PANIC_BAD_G1 verdicts) needs subgroup-checked commitments; only the share-matrix entry decodes
  // lazily (the evaluation-output callers bring their own fully decoded view)
  if (d_shares && !d_out48)
    if (int rc = dkgv_session_redecode_checked(ctx, s)) return rc;

  // ---- evaluation of f_d at every id (for the dealer groups of `filter`)
  CK(ctx->fd_evals.reserve(n_evals * ent_bytes));
  CK(ctx->fd_p0.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_p1.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_da.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_db.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_seedx.reserve((size_t)h * 4));
  CK(ctx->fd_dig.reserve((size_t)n_r * fd_dig_bytes(m)));
  CK(ctx->fd_top.reserve((size_t)n_r * 4));
  // per-share tables of the recombination: recipients are processed in chunks that fit the budget
  const size_t tab_per_recipient = (size_t)(m > 1 ? m - 1 : 0) * FD_TAB_SLOTS * 36 * n_pad * 4;
  const size_t tab_budget = (size_t)4 << 30;
  uint32_t chunk_r = n_r;
  if (tab_per_recipient && tab_per_recipient * chunk_r > tab_budget) chunk_r = (uint32_t)(tab_budget / tab_per_recipient);
  if (chunk_r == 0) chunk_r = 1;
  CK(ctx->fd_tab.reserve(tab_per_recipient ? tab_per_recipient * chunk_r : 16));
  uint32_t* evals = (uint32_t*)ctx->fd_evals.p;
  uint32_t* pp[2] = {(uint32_t*)ctx->fd_p0.p, (uint32_t*)ctx->fd_p1.p};
  uint32_t* dd[2] = {(uint32_t*)ctx->fd_da.p, (uint32_t*)ctx->fd_db.p};

  // seed points, most expensive first (blocks are dispatched in increasing blockIdx.y)
  std::vector<int32_t> order(h);
  for (uint32_t i = 0; i < h; i++) order[i] = plan.lo + (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    return fd_horner_cost(h, (uint32_t)(a < 0 ? -a : a)) > fd_horner_cost(h, (uint32_t)(b < 0 ? -b : b));
  });
  ctx->fd_seed_host.assign(order.begin(), order.end());  // must outlive the async copy
  CK(cudaMemcpyAsync(ctx->fd_seedx.p, ctx->fd_seed_host.data(), (size_t)h * 4, cudaMemcpyHostToDevice, s));

  const unsigned gx = groups, gxv = n_padv / 32;
  const size_t e_hi = (size_t)(plan.hi - plan.lo);  // == h - 1
  const uint32_t ticks = plan.steps + h - 2;
  const bool overlap = ctx->fd_overlap && m > 1;
  if (!overlap) {
    // one stream, phase after phase over all parts at once (also the mode that yields per-phase times)
    k_fd_seed<<<dim3(gx, h, m), FD_NT, FD_SMEM, s>>>(view, (const int32_t*)ctx->fd_seedx.p, plan.lo, evals, n_d, t, h, 0, n_padv, d0, n_pad,
                                                     filter);
    CK(cudaEventRecord(ctx->ev_hot1, s));
    CK(cudaEventRecord(ctx->ev_fd[1], s));
    ctx->launches++;
    // backward differences of every part at hi
    CK(cudaMemcpyAsync(dd[0], evals + e_hi * ent_words, ent_bytes, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(dd[1], evals + e_hi * ent_words, ent_bytes, cudaMemcpyDeviceToDevice, s));
    const uint32_t* src = evals;
    for (uint32_t r = 1; r < h; r++) {
      uint32_t* dst = pp[r & 1];
      uint32_t ipb = items_per_block(gxv, h - r);
      k_fd_init<<<dim3(gxv, (h - r + ipb - 1) / ipb), FD_NT, FD_SMEM, s>>>(src, dst, dd[0], dd[1], n_padv, h, r, 0, ipb, filter, groups);
      ctx->launches++;
      src = dst;
    }
    CK(cudaEventRecord(ctx->ev_fd[2], s));
    // wavefront extension hi+1 .. n_r
    for (uint32_t tick = 1; tick <= ticks; tick++) {
      int32_t k_lo, k_hi;
      fd_ext_band(h, plan.steps, tick, &k_lo, &k_hi);
      if (k_lo > k_hi) continue;
      uint32_t cnt = (uint32_t)(k_hi - k_lo + 1), ipb = items_per_block(gxv, cnt);
      k_fd_ext<<<dim3(gxv, (cnt + ipb - 1) / ipb), FD_NT, FD_SMEM, s>>>(dd[(tick & 1) ^ 1], dd[tick & 1], evals, n_padv, h, tick,
                                                                      (uint32_t)k_lo, (uint32_t)k_hi, e_hi, 0, ipb, filter, groups);
      ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev_fd[3], s));
  } else {
    // The parts are independent until the recombination: each runs its seed -> differences -> extension
    // chain on its own stream, so the tail of one part's kernel is filled by blocks of the others
    // (no grid-wide barrier per round / tick; matters when a rank holds few dealers).
    CK(cudaEventRecord(ctx->fd_fork, s));
    for (uint32_t p = 0; p < m; p++) {
      cudaStream_t sp = ctx->fd_streams[p];
      CK(cudaStreamWaitEvent(sp, ctx->fd_fork, 0));
      k_fd_seed<<<dim3(gx, h, 1), FD_NT, FD_SMEM, sp>>>(view, (const int32_t*)ctx->fd_seedx.p, plan.lo, evals, n_d, t, h, p, n_padv, d0,
                                                        n_pad, filter);
      ctx->launches++;
      const size_t w = (size_t)n_pad * 4, pitch = (size_t)n_padv * 4;
      const uint32_t* col = evals + e_hi * ent_words + (size_t)p * n_pad;
      CK(cudaMemcpy2DAsync(dd[0] + (size_t)p * n_pad, pitch, col, pitch, w, 36, cudaMemcpyDeviceToDevice, sp));
      CK(cudaMemcpy2DAsync(dd[1] + (size_t)p * n_pad, pitch, col, pitch, w, 36, cudaMemcpyDeviceToDevice, sp));
    }
    for (uint32_t r = 1; r < h; r++) {
      uint32_t ipb = items_per_block(gxv, h - r);
      for (uint32_t p = 0; p < m; p++) {
        k_fd_init<<<dim3(gx, (h - r + ipb - 1) / ipb), FD_NT, FD_SMEM, ctx->fd_streams[p]>>>(r == 1 ? evals : pp[(r - 1) & 1], pp[r & 1], dd[0],
                                                                                            dd[1], n_padv, h, r, p * n_pad, ipb, filter, groups);
        ctx->launches++;
      }
    }
    for (uint32_t tick = 1; tick <= ticks; tick++) {
      int32_t k_lo, k_hi;
      fd_ext_band(h, plan.steps, tick, &k_lo, &k_hi);
      if (k_lo > k_hi) continue;
      uint32_t cnt = (uint32_t)(k_hi - k_lo + 1), ipb = items_per_block(gxv, cnt);
      for (uint32_t p = 0; p < m; p++) {
        k_fd_ext<<<dim3(gx, (cnt + ipb - 1) / ipb), FD_NT, FD_SMEM, ctx->fd_streams[p]>>>(dd[(tick & 1) ^ 1], dd[tick & 1], evals, n_padv, h,
                                                                                         tick, (uint32_t)k_lo, (uint32_t)k_hi, e_hi,
                                                                                         p * n_pad, ipb, filter, groups);
        ctx->launches++;
      }
    }
    for (uint32_t p = 0; p < m; p++) {
      CK(cudaEventRecord(ctx->fd_join[p], ctx->fd_streams[p]));
      CK(cudaStreamWaitEvent(s, ctx->fd_join[p], 0));
    }
    CK(cudaEventRecord(ctx->ev_hot1, s));
    for (int i = 1; i <= 3; i++) CK(cudaEventRecord(ctx->ev_fd[i], s));  // phases overlap: only their sum is defined
  }
  ctx->hot_recorded = true;

  // ---- recombination + comparison (or evaluation output), ids in ascending order, in chunks that fit the table budget
  if (m > 1) {
    k_fd_digits<<<(n_r + 127) / 128, 128, 0, s>>>(n_r, h, m, (int8_t*)ctx->fd_dig.p, (int32_t*)ctx->fd_top.p);
    ctx->launches++;
  }
  for (uint32_t r0 = 0; r0 < n_r; r0 += chunk_r) {
    uint32_t nj = std::min(chunk_r, n_r - r0);
    if (d_out48)
      k_fd_combine_out<<<dim3(gx, nj), FD_NT, FD_SMEM, s>>>(evals, plan.lo, m, (const int8_t*)ctx->fd_dig.p, (const int32_t*)ctx->fd_top.p, d_ids,
                                                           d_out48, (uint32_t*)ctx->fd_tab.p, n_pad, n_d, n_r, r0, cols, 0, d0);
    else
      k_fd_combine<<<dim3(gx, nj), FD_NT, FD_SMEM, s>>>(evals, plan.lo, m, (const int8_t*)ctx->fd_dig.p, (const int32_t*)ctx->fd_top.p, d_ids,
                                                       d_shares, ctx->gtab, (const uint8_t*)ctx->dealer_bad.p, d_status, (uint32_t*)ctx->fd_tab.p,
                                                       n_pad, n_d, n_r, r0, cols, 0, d0, filter);
    ctx->launches++;
  }
  CK(cudaEventRecord(ctx->ev_fd[4], s));
  ctx->fd_recorded = true;
  CK(cudaGetLastError());
  return 0;
}

// Dealers are independent, so a ceremony whose planes would not fit the memory budget is processed in
// dealer chunks (multiples of 32 columns).  (1024, 683) on one GPU needs 1.5 GB of planes + 3.6 GB of tables.
static int fd_run_chunks(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint32_t* h_ids, const uint8_t* d_shares, uint8_t* d_status, uint8_t* d_out48,
                         cudaStream_t s) {
  const size_t n_evals = (size_t)((int64_t)n_r - plan.lo + 1);
  const size_t bytes_per_col = (n_evals + 4 * (size_t)plan.h) * 36 * plan.m * 4;
  size_t budget = (size_t)12 << 30;
  if (const char* e = getenv("DKGV_FD_PLANE_BUDGET_MB")) budget = (size_t)atoll(e) << 20;  // tests force several chunks
  uint32_t chunk = (uint32_t)std::min<size_t>(view.n_pad, std::max<size_t>(32, (budget / bytes_per_col) & ~(size_t)31));
  for (uint32_t d0 = 0; d0 < view.n_pad && d0 < n_d; d0 += chunk) {
    uint32_t cols = std::min(chunk, view.n_pad - d0);
    if (int rc = fd_run(ctx, view, d0, cols, n_d, n_r, t, plan, d_ids, h_ids, d_shares, d_status, d_out48, s)) return rc;
  }
  return 0;
}

int dkgv_share_matrix_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint32_t* h_ids, const uint8_t* d_shares, uint8_t* d_status, cudaStream_t s) {
  return fd_run_chunks(ctx, view, n_d, n_r, t, plan, d_ids, h_ids, d_shares, d_status, nullptr, s);
}

// evaluate_polynomial (dkg_math.rs:160-174) of every dealer at every id, ids a permutation of 1..n_r:
// d_out48[d][j] = compress(f_d(ids[j])).  Used for the final keys K_j of agg_coefficients (one "dealer": the
// column sums) and the batched dkgv_feldman_eval.
int dkgv_feldman_eval_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint32_t* h_ids, uint8_t* d_out48, cudaStream_t s) {
  return fd_run_chunks(ctx, view, n_d, n_r, t, plan, d_ids, h_ids, nullptr, nullptr, d_out48, s);
}
