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
// before them the consistency shortcut (k_fd_cols / k_fd_tables / k_fd_share_limbs / k_fd_difftab / k_fd_coefpoint /
// k_fd_coefsign / k_fd_need / k_fd_fill_ok, see below): a dealer group whose shares are provably all valid never enters
// the evaluation, and its commitments are never decompressed.  The shortcut is queued WITHOUT any host synchronisation
// (dkgv_share_submit); whether some dealer group still needs the evaluation is a device flag the caller reads together
// with its results (dkgv_share_finish).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ctx.hpp"
#include "fdiff.cuh"
#include "share_rs.cuh"

using namespace dkgv;


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
             const uint32_t* __restrict__ ids, const uint8_t* __restrict__ shares, GTab gtab,
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
constexpr uint32_t FD_SHORTCUT_MAX_T = 2047;  // lz_muladd_small: factors j < 2^11; t < n_r <= FD_SHORTCUT_MAX_N
constexpr uint32_t FD_SHORTCUT_MAX_N = 2048;  // k_fd_difftab: one block per dealer, two entries per thread

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

// shares of one dealer chunk as little-endian limbs in ascending-id order: sl[d - d0][x - 1][8].  A share >= r fails condition (1):
// poly_ok[d] = 0, the dealer is marked for the repair route (share_rs.cuh), which treats the share as a wrong value (0 in the table,
// oor[d - d0][x - 1] = 1 so that its verdict stays SECRET_RANGE whatever the decoder finds)
__global__ void __launch_bounds__(128)
k_fd_share_limbs(const uint8_t* __restrict__ shares, const uint32_t* __restrict__ cols, uint32_t* __restrict__ sl, uint8_t* __restrict__ poly_ok,
                 uint8_t* __restrict__ state, uint8_t* __restrict__ oor, uint32_t d0, uint32_t n_cols, uint32_t n_d, uint32_t n_r) {
  uint32_t xi = blockIdx.x * blockDim.x + threadIdx.x, dl = blockIdx.y;
  if (xi >= n_r || d0 + dl >= n_d) return;
  uint32_t l[8];
  uint32_t c = cols[xi];
  if (c >= n_r) c = 0;  // ids that are not a permutation: the speculative run reads in bounds, its results are discarded
  bool ok = fr_raw_from_be32(l, shares + ((size_t)(d0 + dl) * n_r + c) * 32);
  if (!ok) {
    poly_ok[d0 + dl] = 0;
    state[d0 + dl] = RS_REPAIR;
  }
  if (oor) oor[(size_t)dl * n_r + xi] = ok ? 0 : 1;
  uint32_t* o = sl + ((size_t)dl * n_r + xi) * 8;
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = ok ? l[i] : 0u;
}

// Condition (2) and the interpolation in ONE difference table per dealer (n_r <= 2048), lazy residues (fdiff.cuh, "lazy residues for the
// difference table": 9-limb values reduced every 30 / every dt2_period(t) rounds), no products except by small integers:
//   phase 1  t rounds e[k] <- e[k] - e[k-1] (k >= r) over all n_r shares: e[k] = Delta^k s(1) for k < t, and the entries
//            k >= t are the t-th differences Delta^t s(k-t+1) - all zero <=> condition (2);
//   phase 2  E_k = e[k] / k!, then P <- P (x - j) + E_{j-1} for j = t-1 .. 1 on monomial coefficients: c[k] <- c[k-1] - j c[k].
// One block per dealer, thread i owns entries 2i and 2i+1 in registers and publishes only e[2i+1] per round (double-buffered,
// one barrier per round); entries that are already final (phase 1: k < r) or still zero (phase 2: k > t - j) are skipped, so
// whole warps drop out.  Work per dealer: t n - t^2/2 nine-limb subtractions + t^2/2 nine-limb multiply-adds by a small integer.
__global__ void __launch_bounds__(1024)
k_fd_difftab(const uint32_t* __restrict__ sl, const uint32_t* __restrict__ ifact, uint32_t* __restrict__ coef, uint8_t* __restrict__ poly_ok,
             uint8_t* __restrict__ state, uint32_t d0, uint32_t n_d, uint32_t n_r, uint32_t t, const uint32_t* __restrict__ map,
             const uint32_t* __restrict__ n_map) {
  extern __shared__ uint32_t fr_sm[];  // pub[2][9 * blockDim.x] (lz_publish planes), E'[t]
  if (map && blockIdx.x >= *n_map) return;  // second pass of the repair route: one block per candidate
  const uint32_t dl = map ? map[blockIdx.x] : blockIdx.x, i = threadIdx.x, nt = blockDim.x;
  if (d0 + dl >= n_d || !poly_ok[d0 + dl]) return;  // whole block; a share >= r already failed condition (1)
  uint32_t* pub = fr_sm;
  Fr* E = (Fr*)(fr_sm + 18 * (size_t)nt);
  const uint32_t k0 = 2 * i, k1 = 2 * i + 1;
  DtPair p;
  p.a = lz_zero();
  p.b = lz_zero();
  const uint32_t* row = sl + (size_t)dl * n_r * 8;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    if (k0 < n_r) p.a.l[l] = row[(size_t)k0 * 8 + l];
    if (k1 < n_r) p.b.l[l] = row[(size_t)k1 * 8 + l];
  }
  uint32_t left = DT1_PERIOD;
#pragma unroll 1
  for (uint32_t r = 1; r <= t; r++) {
    uint32_t* pr = pub + (size_t)(r & 1) * 9 * nt;
    if (dt1_publishes(i, r)) lz_publish(pr, nt, i, p.b);
    __syncthreads();
    const bool red = --left == 0;
    if (red) left = DT1_PERIOD;
    if (dt1_active(i, r)) dt1_step(p, i, r, red, pr, nt);
  }
  dt1_finish(p);
  bool bad = (k0 >= t && k0 < n_r && !lz_is_zero(p.a)) || (k1 >= t && k1 < n_r && !lz_is_zero(p.b));
  if (__syncthreads_or(bad)) {  // some t-th difference is not zero: the shares are not on a polynomial of degree < t
    if (i == 0) {
      poly_ok[d0 + dl] = 0;
      if (state) state[d0 + dl] = RS_REPAIR;  // the shares are not on one polynomial: the repair route may still find it
    }
    return;
  }
  Fr f;
  if (k0 < t) {
#pragma unroll
    for (int l = 0; l < 8; l++) f.l[l] = ifact[(size_t)k0 * 8 + l];
    E[k0] = dt2_signed_e(mul(lz_low(p.a), f), t, k0);  // canonical x Montgomery = canonical
  }
  if (k1 < t) {
#pragma unroll
    for (int l = 0; l < 8; l++) f.l[l] = ifact[(size_t)k1 * 8 + l];
    E[k1] = dt2_signed_e(mul(lz_low(p.b), f), t, k1);
  }
  __syncthreads();
  p.a = i == 0 ? lz_from(E[t - 1]) : lz_zero();
  p.b = lz_zero();
  const uint32_t period = dt2_period(t);
  left = period;
#pragma unroll 1
  for (uint32_t j = t - 1; j >= 1; j--) {
    uint32_t* pr = pub + (size_t)(j & 1) * 9 * nt;
    const bool act = dt2_active(i, j, t);
    if (act) lz_publish(pr, nt, i, p.b);
    __syncthreads();
    const bool red = --left == 0;
    if (red) left = period;
    if (act) dt2_step(p, i, j, red, pr, nt, E);
  }
  dt2_finish(p, i, t);
  uint32_t* o = coef + (size_t)dl * t * 8;
#pragma unroll
  for (int l = 0; l < 8; l++) {
    if (k0 < t) o[(size_t)k0 * 8 + l] = p.a.l[l];
    if (k1 < t) o[(size_t)k1 * 8 + l] = p.b.l[l];
  }
}

// condition (3) WITHOUT decoding the commitments (fdiff.cuh, "against the COMPRESSED commitment"): x half.  Same launch shape
// (warp = 32 dealers x one k, the layout of the seeds); writes Z (limbs 0..11) and Y (12..23) of G * p_k to the chunk-local planes yz[(k*24 + w) * n_cols + dl].
__global__ void __launch_bounds__(FD_NT)
k_fd_coefpoint(const uint8_t* __restrict__ vv, const uint32_t* __restrict__ coef, GTab gtab,
               uint8_t* __restrict__ poly_ok, uint32_t* __restrict__ yz, uint32_t d0, uint32_t n_d, uint32_t n_cols, uint32_t t,
               const uint32_t* __restrict__ map, const uint32_t* __restrict__ n_map) {
  extern __shared__ U4 opfile[];
  // column c of the Y / Z planes; without a map column c is dealer d0 + c, with one (second pass of the repair route: the
  // candidates compacted into dense groups) it is dealer d0 + map[c], c < *n_map
  const uint32_t c = blockIdx.x * 32 + threadIdx.x, k = blockIdx.y;
  uint32_t dl = c;
  bool active = d0 + c < n_d;
  if (map) {
    const uint32_t nm = *n_map;
    if (blockIdx.x * 32 >= nm) return;
    active = c < nm;
    dl = map[active ? c : 0];
  }
  uint32_t dc = active ? dl : (map ? dl : n_d - 1 - d0);
#if defined(__CUDA_ARCH__)
  if (__ballot_sync(0xffffffffu, active && poly_ok[d0 + dl]) == 0) return;  // nobody in this group can still pass
#endif
  OpFile f{opfile + threadIdx.x, FD_NT};
  uint32_t sc[8];
#pragma unroll
  for (int l = 0; l < 8; l++) sc[l] = coef[((size_t)dc * t + k) * 8 + l];
  Fp y, z;
  bool same = fd_coef_point(f, gtab, sc, vv + ((size_t)(d0 + dc) * t + k) * 48, &y, &z);
  uint32_t* o = yz + (size_t)k * 24 * n_cols + c;
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
              uint32_t n_cols, uint32_t t, const uint32_t* __restrict__ map, const uint32_t* __restrict__ n_map) {
  const uint32_t c = blockIdx.x * 32 + threadIdx.x, k0 = (blockIdx.y * blockDim.y + threadIdx.y) * FD_SIGN_K;
  uint32_t dl = c;
  if (map) {
    if (c >= *n_map) return;
    dl = map[c];
  }
  if (d0 + dl >= n_d || k0 >= t || !poly_ok[d0 + dl]) return;
  int cnt = (int)(t - k0 < (uint32_t)FD_SIGN_K ? t - k0 : (uint32_t)FD_SIGN_K);
  Fp z[FD_SIGN_K], y[FD_SIGN_K];
  uint8_t fs[FD_SIGN_K];
#pragma unroll 1
  for (int i = 0; i < cnt; i++) {
    const uint32_t* e = yz + (size_t)(k0 + i) * 24 * n_cols + c;
#pragma unroll
    for (int w = 0; w < 12; w++) {
      z[i].l[w] = e[(size_t)w * n_cols];
      y[i].l[w] = e[(size_t)(12 + w) * n_cols];
    }
    fs[i] = (vv[((size_t)(d0 + dl) * t + k0 + i) * 48] >> 5) & 1;
  }
  if (!fd_coef_signs<FD_SIGN_K>(z, y, fs, cnt)) poly_ok[d0 + dl] = 0;
}

// cols[x - 1] = the column j with ids[j] == x (columns in ascending-id order); flags[0] = 1 when the ids are not a permutation of
// 1..n_r (then nothing the speculative shortcut wrote counts and dkgv_share_finish takes the Horner route).  cols starts at 0xffffffff.
__global__ void __launch_bounds__(128) k_fd_cols(const uint32_t* __restrict__ ids, uint32_t n_r, uint32_t* __restrict__ cols, uint32_t* __restrict__ flags) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_r) return;
  uint32_t x = ids[j];
  if (x < 1 || x > n_r) {
    atomicOr(&flags[0], 1u);
    return;
  }
  if (atomicExch(&cols[x - 1], j) != 0xffffffffu) atomicOr(&flags[0], 1u);
}

// start of a submitted job: cols = 0xffffffff, poly_ok = 1, need_group = 0, flags = {0, pending0}
__global__ void __launch_bounds__(256) k_fd_prep(uint32_t* __restrict__ cols, uint32_t n_r, uint8_t* __restrict__ poly_ok, uint32_t n_pad,
                                                 uint8_t* __restrict__ need_group, uint32_t* __restrict__ flags, uint32_t pending0,
                                                 uint8_t* __restrict__ state) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_r) cols[i] = 0xffffffffu;
  if (i < n_pad) {
    poly_ok[i] = 1;
    state[i] = RS_IDLE;
  }
  if (i < n_pad / 32) need_group[i] = 0;
  if (i == 0) {
    flags[0] = 0;
    flags[1] = pending0;
  }
}

// need_group[g] = 1 when some dealer of the 32-dealer group g fails a condition; flags[1] counts the failing dealers
__global__ void __launch_bounds__(128)
k_fd_need(const uint8_t* __restrict__ poly_ok, uint32_t n_d, uint8_t* __restrict__ need_group, uint32_t* __restrict__ flags) {
  uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_d) return;
  if (poly_ok[d] == 0) {
    need_group[d / 32] = 1;
    atomicAdd(&flags[1], 1u);
  }
}

// verdict OK for every share of the dealer groups that met the three conditions (block = 128 columns of one dealer)
__global__ void __launch_bounds__(128)
k_fd_fill_ok(uint8_t* __restrict__ status, const uint8_t* __restrict__ need_group, const uint8_t* __restrict__ state, uint32_t n_d, uint32_t n_r) {
  uint32_t j = blockIdx.y * blockDim.x + threadIdx.x, d = blockIdx.x;
  if (j >= n_r || d >= n_d || need_group[d / 32] || state[d] == RS_CANDIDATE) return;  // a repaired dealer keeps its own verdicts
  status[(size_t)d * n_r + j] = DKGV_OK;
}

// ---- per-dealer fallback: the dealers still unsettled after the shortcut and the repair route, compacted into dense groups ------------
// ordered list of the dealers with poly_ok == 0 (one warp; n_d <= 65535)
__global__ void __launch_bounds__(32) k_fd_need_list(const uint8_t* __restrict__ poly_ok, uint32_t n_d, uint32_t* __restrict__ list, uint32_t cap) {
  uint32_t count = 0;
  for (uint32_t base = 0; base < n_d; base += 32) {
    const uint32_t d = base + threadIdx.x;
    const bool f = d < n_d && poly_ok[d] == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, f);
    const uint32_t pos = count + __popc(m & ((1u << threadIdx.x) - 1));
    if (f && pos < cap) list[pos] = d;
    count += __popc(m);
  }
}
// verdict OK for the settled dealers inside the groups k_fd_fill_ok left alone (a repaired dealer keeps its own verdicts)
__global__ void __launch_bounds__(128)
k_fd_fill_ok_dealers(uint8_t* __restrict__ status, const uint8_t* __restrict__ need_group, const uint8_t* __restrict__ poly_ok,
                     const uint8_t* __restrict__ state, uint32_t n_d, uint32_t n_r) {
  uint32_t j = blockIdx.y * blockDim.x + threadIdx.x, d = blockIdx.x;
  if (j >= n_r || d >= n_d || !need_group[d / 32] || !poly_ok[d] || state[d] == RS_CANDIDATE) return;
  status[(size_t)d * n_r + j] = DKGV_OK;
}
// share rows (n_r x 32 B, 16-byte pieces) of the listed dealers into a dense table; status rows of the dense table back
__global__ void __launch_bounds__(128)
k_fd_gather_rows(const uint4* __restrict__ src, uint4* __restrict__ dst, const uint32_t* __restrict__ list, uint32_t row16) {
  const uint32_t c = blockIdx.x, i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i < row16) dst[(size_t)c * row16 + i] = src[(size_t)list[c] * row16 + i];
}
__global__ void __launch_bounds__(128)
k_fd_scatter_status(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const uint32_t* __restrict__ list, uint32_t n_r) {
  const uint32_t c = blockIdx.x, j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j < n_r) dst[(size_t)list[c] * n_r + j] = src[(size_t)c * n_r + j];
}

// The `count` dealers still unsettled (flags[1], read back by the caller) as a dense session: *list (device) their indices in ascending
// order, *shares_c / *status_c dense share and status tables; the settled dealers of their groups get their OK verdicts here.
int dkgv_fd_compact_unsettled(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t count, const uint8_t* d_shares, uint8_t* d_status, cudaStream_t s,
                              const uint32_t** list, const uint8_t** shares_c, uint8_t** status_c) {
  const uint32_t n_pad = (n_d + 31) & ~31u, groups = n_pad / 32;
  const uint8_t* poly_ok = (const uint8_t*)ctx->fd_flags.p;
  const uint8_t* need_group = poly_ok + n_pad;
  const uint8_t* state = need_group + groups;
  CK(ctx->fd_cmp_list.reserve((size_t)count * 4));
  CK(ctx->fd_cmp_sh.reserve((size_t)count * n_r * 32));
  CK(ctx->fd_cmp_st.reserve((size_t)count * n_r));
  const unsigned gy = (n_r + 127) / 128;
  k_fd_need_list<<<1, 32, 0, s>>>(poly_ok, n_d, (uint32_t*)ctx->fd_cmp_list.p, count);
  k_fd_fill_ok_dealers<<<dim3(n_d, gy), 128, 0, s>>>(d_status, need_group, poly_ok, state, n_d, n_r);
  k_fd_gather_rows<<<dim3(count, (n_r * 2 + 127) / 128), 128, 0, s>>>((const uint4*)d_shares, (uint4*)ctx->fd_cmp_sh.p, (const uint32_t*)ctx->fd_cmp_list.p,
                                                                   n_r * 2);
  ctx->launches += 3;
  CK(cudaGetLastError());
  *list = (const uint32_t*)ctx->fd_cmp_list.p;
  *shares_c = (const uint8_t*)ctx->fd_cmp_sh.p;
  *status_c = (uint8_t*)ctx->fd_cmp_st.p;
  return 0;
}
int dkgv_fd_scatter_status(dkgv_ctx* ctx, uint32_t n_r, uint32_t count, uint8_t* d_status, cudaStream_t s) {
  k_fd_scatter_status<<<dim3(count, (n_r + 127) / 128), 128, 0, s>>>((const uint8_t*)ctx->fd_cmp_st.p, d_status, (const uint32_t*)ctx->fd_cmp_list.p, n_r);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
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
                        (const void*)k_fd_combine_out, (const void*)k_fd_coefpoint}) {
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FD_SMEM));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  if (const char* e = getenv("DKGV_FD_IPB")) ctx->fd_ipb_force = (uint32_t)atoi(e);  // experiments; read once per ctx
  CK(cudaFuncSetAttribute(k_fd_difftab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 9 * 1024 * 4 + FD_SHORTCUT_MAX_T * 32)));
  CK(cudaFuncSetAttribute(k_fd_difftab, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  size_t stack = 0;  // k_fd_coefsign keeps 3 x FD_SIGN_K field elements in local memory (1.2 KB frame); only ever raise the limit
  CK(cudaDeviceGetLimit(&stack, cudaLimitStackSize));
  if (stack < 2048) CK(cudaDeviceSetLimit(cudaLimitStackSize, 2048));
  for (int i = 0; i < 5; i++) {
    CK(cudaEventCreate(&ctx->ev_fd[i]));
    CK(cudaEventCreate(&ctx->ev_sc[i]));
  }
  CK(cudaEventCreateWithFlags(&ctx->fd_fork, cudaEventDisableTiming));
  for (uint32_t i = 0; i < FD_MAX_PARTS; i++) {
    CK(cudaStreamCreateWithFlags(&ctx->fd_streams[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->fd_join[i], cudaEventDisableTiming));
  }
  return 0;
}

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

// the consistency shortcut can settle this shape (else every share goes through the evaluation)
bool dkgv_fd_shortcut_applies(const dkgv_ctx* ctx, uint32_t n_r, uint32_t t) {
  return ctx->fd_polycheck && t >= 1 && n_r > t && t <= FD_SHORTCUT_MAX_T && n_r <= FD_SHORTCUT_MAX_N;
}

// Queues, WITHOUT synchronising: cols from the device ids (flags[0] = ids are not a permutation of 1..n_r) and, when `shortcut`,
// the consistency shortcut over all dealers - verdict OK written for every dealer group that met the three conditions,
// need_group / flags[1] = the dealers that did not.  Without the shortcut flags[1] = 1: everything is still pending.
int dkgv_take_vv_wait(dkgv_ctx* ctx, cudaStream_t s);  // dkgv.cu

int dkgv_fd_submit(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint32_t* d_ids, const uint8_t* d_shares,
                   uint8_t* d_status, bool shortcut, uint32_t* d_flags, cudaStream_t s) {
  const uint32_t n_pad = (n_d + 31) & ~31u, groups = n_pad / 32;
  CK(ctx->fd_cols.reserve((size_t)n_r * 4));
  CK(ctx->fd_flags.reserve((size_t)n_pad * 3 + groups + (size_t)n_pad * 12 + 64));
  uint32_t* cols = (uint32_t*)ctx->fd_cols.p;
  uint8_t* poly_ok = (uint8_t*)ctx->fd_flags.p;
  uint8_t* need_group = poly_ok + n_pad;
  uint8_t* state = need_group + groups;  // [n_pad] dealer states of the repair route; behind it ok2 [n_pad], deg / cnt [n_pad] u32, counter
  if (!shortcut)
    if (int rc = dkgv_take_vv_wait(ctx, s)) return rc;
  CK(cudaEventRecord(ctx->ev_sc[0], s));
  CK(cudaEventRecord(ctx->ev_hot0, s));
  const uint32_t prep_n = std::max(n_r, n_pad);
  k_fd_prep<<<(prep_n + 255) / 256, 256, 0, s>>>(cols, n_r, poly_ok, n_pad, need_group, d_flags, shortcut ? 0u : 1u, state);
  k_fd_cols<<<(n_r + 127) / 128, 128, 0, s>>>(d_ids, n_r, cols, d_flags);
  ctx->launches += 2;
  if (shortcut) {
    if (ctx->fd_binom_t != t) {
      CK(ctx->fd_binom.reserve((size_t)(t + 1) * 96));
      k_fd_tables<<<(t + 128) / 128, 128, 0, s>>>(t, (uint32_t*)ctx->fd_binom.p, (uint32_t*)ctx->fd_binom.p + (size_t)(t + 1) * 8,
                                                   (uint32_t*)ctx->fd_binom.p + (size_t)(t + 1) * 16);
      ctx->fd_binom_t = t;
      ctx->launches++;
    }
    const uint32_t* ifact = (const uint32_t*)ctx->fd_binom.p + (size_t)(t + 1) * 16;
    // dealer chunks: share limbs + coefficients + Y/Z planes of a chunk stay within a 4 GB budget (123 MB at (1024, 683))
    const size_t per_dealer = (size_t)n_r * 32 + (size_t)t * 128;
    const uint32_t chunk = (uint32_t)std::min<size_t>(std::min<size_t>(n_pad, 32768), std::max<size_t>(32, (((size_t)4 << 30) / per_dealer) & ~(size_t)31));
    CK(ctx->fd_sl.reserve((size_t)chunk * n_r * 32));
    CK(ctx->fd_coef.reserve((size_t)chunk * t * 32));
    CK(ctx->fd_yz.reserve((size_t)t * 24 * chunk * 4));
    const uint32_t nt = (((n_r + 1) / 2 + 31) / 32) * 32;
    const uint32_t batches = (t + FD_SIGN_K - 1) / FD_SIGN_K;
    for (uint32_t d0 = 0; d0 < n_d; d0 += chunk) {
      const uint32_t n_cols = std::min(chunk, n_pad - d0), n_here = std::min(n_cols, n_d - d0), g_here = n_cols / 32;
      const bool first = d0 == 0, last = d0 + chunk >= n_d;
      k_fd_share_limbs<<<dim3((n_r + 127) / 128, n_here), 128, 0, s>>>(d_shares, cols, (uint32_t*)ctx->fd_sl.p, poly_ok, state, nullptr, d0, n_cols, n_d,
                                                                       n_r);
      k_fd_difftab<<<n_here, nt, (size_t)nt * 72 + (size_t)t * 32, s>>>((const uint32_t*)ctx->fd_sl.p, ifact, (uint32_t*)ctx->fd_coef.p, poly_ok, state, d0, n_d,
                                                               n_r, t, nullptr, nullptr);
      if (first) CK(cudaEventRecord(ctx->ev_sc[1], s));  // phases (of the first chunk): [limbs + difference table | x halves | sign halves | flags]
      if (int rc = dkgv_take_vv_wait(ctx, s)) return rc;
      k_fd_coefpoint<<<dim3(g_here, t), FD_NT, FD_SMEM, s>>>(d_vv, (const uint32_t*)ctx->fd_coef.p, ctx->gtab, poly_ok, (uint32_t*)ctx->fd_yz.p, d0,
                                                             n_d, n_cols, t, nullptr, nullptr);
      if (first) CK(cudaEventRecord(ctx->ev_sc[2], s));
      k_fd_coefsign<<<dim3(g_here, (batches + 3) / 4), dim3(32, 4), 0, s>>>(d_vv, (const uint32_t*)ctx->fd_yz.p, poly_ok, d0, n_d, n_cols, t, nullptr,
                                                                         nullptr);
      if (last) CK(cudaEventRecord(ctx->ev_sc[3], s));
      ctx->launches += 4;
    }
    k_fd_need<<<(n_d + 127) / 128, 128, 0, s>>>(poly_ok, n_d, need_group, d_flags);
    k_fd_fill_ok<<<dim3(n_d, (n_r + 127) / 128), 128, 0, s>>>(d_status, need_group, state, n_d, n_r);
    ctx->launches += 2;
  } else {
    for (int i = 1; i <= 3; i++) CK(cudaEventRecord(ctx->ev_sc[i], s));
  }
  CK(cudaEventRecord(ctx->ev_sc[4], s));
  CK(cudaEventRecord(ctx->ev_hot1, s));
  ctx->hot_recorded = true;
  ctx->sc_recorded = true;
  CK(cudaGetLastError());
  return 0;
}
const uint8_t* dkgv_fd_need_groups(const dkgv_ctx* ctx, uint32_t n_d) { return (const uint8_t*)ctx->fd_flags.p + ((n_d + 31) & ~31u); }

// the repair route applies: the shortcut's shapes, at least one correctable error, blocks that fit (k_rs_bm: tau + 2 threads)
bool dkgv_fd_repair_applies(const dkgv_ctx* ctx, uint32_t n_r, uint32_t t) {
  return ctx->fd_repair && dkgv_fd_shortcut_applies(ctx, n_r, t) && (n_r - t) / 2 >= 1 && (n_r - t) / 2 + 2 <= 1024 && t >= 2;
}

// Repair route (share_rs.cuh) for the dealers a submitted job marked RS_REPAIR, then the need flags again: asynchronous on s; d_flags[1]
// ends as the number of dealers STILL unsettled.  Called by share_finish after it has seen flags[1] != 0.
int dkgv_fd_repair(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint8_t* d_shares, uint8_t* d_status,
                   uint32_t* d_flags, cudaStream_t s) {
  const uint32_t n_pad = (n_d + 31) & ~31u, groups = n_pad / 32;
  const uint32_t nsyn = n_r - t, tau = nsyn / 2;
  uint32_t* cols = (uint32_t*)ctx->fd_cols.p;
  uint8_t* poly_ok = (uint8_t*)ctx->fd_flags.p;
  uint8_t* need_group = poly_ok + n_pad;
  uint8_t* state = need_group + groups;
  uint8_t* ok2 = state + n_pad;
  uint32_t* deg = (uint32_t*)(((uintptr_t)(ok2 + n_pad) + 15) & ~(uintptr_t)15);
  uint32_t* cnt = deg + n_pad;
  uint32_t* repaired = cnt + n_pad;
  uint32_t* n_cand = repaired + 1;
  uint32_t* cand = n_cand + 3;  // [n_pad] dealers (chunk-local) of the second pass, compacted
  if (ctx->rs_n != n_r || ctx->rs_t != t) {  // tables of the shape: dual weights u[n], syndrome matrix mt[nsyn][nsyn] (3.7 MB at (1024, 683)), scratch
    CK(ctx->rs_tab.reserve(((size_t)n_r + (size_t)nsyn * nsyn + (size_t)3 * nsyn) * 32));
    uint32_t* tab = (uint32_t*)ctx->rs_tab.p;
    k_rs_tables<<<(n_r + 127) / 128, 128, 0, s>>>(n_r, tab);
    k_rs_mtab<<<1, 1024, 0, s>>>(n_r, nsyn, tab + (size_t)n_r * 8, tab + ((size_t)n_r + (size_t)nsyn * nsyn) * 8);
    ctx->rs_n = n_r;
    ctx->rs_t = t;
    ctx->launches += 2;
  }
  const uint32_t* tab_u = (const uint32_t*)ctx->rs_tab.p;
  const uint32_t* tab_mt = tab_u + (size_t)n_r * 8;
  const uint32_t* ifact = (const uint32_t*)ctx->fd_binom.p + (size_t)(t + 1) * 16;
  const size_t per_dealer = (size_t)n_r * 35 + (size_t)t * 128 + (size_t)nsyn * 64 + (size_t)(tau + 1) * 96 + 256;
  const uint32_t chunk = (uint32_t)std::min<size_t>(std::min<size_t>(n_pad, 32768), std::max<size_t>(32, (((size_t)4 << 30) / per_dealer) & ~(size_t)31));
  CK(ctx->fd_sl.reserve((size_t)chunk * n_r * 32));
  CK(ctx->fd_coef.reserve((size_t)chunk * t * 32));
  CK(ctx->fd_yz.reserve((size_t)t * 24 * chunk * 4));
  const uint32_t bm_threads = std::max<uint32_t>(64, (tau + 2 + 31) & ~31u);
  CK(ctx->rs_work.reserve((size_t)chunk * ((size_t)n_r * 2 + 1 + (size_t)nsyn * 64 + (size_t)(tau + 1) * 32 + (size_t)bm_threads * 64 + 64) + 512));
  uint8_t* w = (uint8_t*)ctx->rs_work.p;
  uint32_t* syn = (uint32_t*)w;
  uint32_t* gdf = syn + (size_t)chunk * nsyn * 8;
  uint32_t* lam = gdf + (size_t)chunk * nsyn * 8;
  uint8_t* err = (uint8_t*)(lam + (size_t)chunk * (tau + 1) * 8);
  uint8_t* oor = err + (size_t)chunk * n_r;
  uint8_t* bmdone = oor + (size_t)chunk * n_r;
  uint32_t* park = (uint32_t*)(((uintptr_t)(bmdone + chunk) + 63) & ~(uintptr_t)63);  // [chunk][bm_threads * 16 + 16] words
  CK(cudaMemsetAsync(deg, 0, (size_t)n_pad * 8 + 16, s));  // deg, cnt, repaired
  CK(cudaMemsetAsync(ok2, 0, n_pad, s));
  static bool attr = false;
  if (!attr) {
    CK(cudaFuncSetAttribute(k_rs_syndromes, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 32));
    CK(cudaFuncSetAttribute(k_rs_gdiff, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 9 * 1024 * 4));
    CK(cudaFuncSetAttribute(k_rs_bm, cudaFuncAttributeMaxDynamicSharedMemorySize, (2048 + 1024 + 40) * 32));
    CK(cudaFuncSetAttribute(k_rs_forney, cudaFuncAttributeMaxDynamicSharedMemorySize, (4 * 1024 + 2) * 32 + 2048 * 4));
    attr = true;
  }
  const uint32_t nt = (((n_r + 1) / 2 + 31) / 32) * 32, batches = (t + FD_SIGN_K - 1) / FD_SIGN_K;
  const unsigned gy = (n_r + 127) / 128;
  for (uint32_t d0 = 0; d0 < n_d; d0 += chunk) {
    const uint32_t n_cols = std::min(chunk, n_pad - d0), n_here = std::min(n_cols, n_d - d0), g_here = n_cols / 32;
    // the share table of this chunk again (chunks of a large session share the buffers), now with the out-of-range marks
    k_fd_share_limbs<<<dim3(gy, n_here), 128, 0, s>>>(d_shares, cols, (uint32_t*)ctx->fd_sl.p, poly_ok, state, oor, d0, n_cols, n_d, n_r);
    k_rs_gdiff<<<n_here, nt, (size_t)nt * 72, s>>>((const uint32_t*)ctx->fd_sl.p, state, gdf, d0, n_r, t, nsyn);
    // two stages (share_rs.cuh k_rs_bm): few wrong shares per dealer finish on the first RS_STAGE1 syndromes
    CK(cudaMemsetAsync(bmdone, 0, n_here, s));
    const uint32_t stage1 = std::min<uint32_t>(RS_STAGE1, nsyn);
    k_rs_syndromes<<<dim3(n_here, (stage1 + 127) / 128), 128, (size_t)stage1 * 32, s>>>(gdf, state, bmdone, tab_mt, syn, d0, nsyn, 0, stage1);
    k_rs_bm<<<n_here, bm_threads, ((size_t)stage1 + bm_threads + bm_threads / 32 + 1) * 32, s>>>(syn, state, lam, deg, bmdone, park, d0, nsyn, 0, stage1, tau);
    if (stage1 < nsyn) {
      k_rs_syndromes<<<dim3(n_here, (nsyn - stage1 + 127) / 128), 128, (size_t)nsyn * 32, s>>>(gdf, state, bmdone, tab_mt, syn, d0, nsyn, stage1, nsyn);
      k_rs_bm<<<n_here, bm_threads, ((size_t)nsyn + bm_threads + bm_threads / 32 + 1) * 32, s>>>(syn, state, lam, deg, bmdone, park, d0, nsyn, stage1, nsyn,
                                                                                                 tau);
      ctx->launches += 2;
    }
    k_rs_chien<<<dim3(n_here, gy), 128, 0, s>>>(lam, deg, state, err, cnt, d0, n_r, tau);
    k_rs_forney<<<n_here, 256, ((size_t)4 * tau + 2) * 32 + (size_t)n_r * 4, s>>>((uint32_t*)ctx->fd_sl.p, err, cnt, deg, state, syn, lam, tab_u, d0, n_r,
                                                                                  nsyn, tau);
    CK(cudaMemsetAsync(n_cand, 0, 4, s));
    k_rs_stage<<<(n_here + 127) / 128, 128, 0, s>>>(state, ok2, cand, n_cand, d0, n_here);
    // second pass of the exact conditions on the corrected table: t-th differences + coefficients, compress(G * p_k) == C_k
    // (the candidates are compacted into dense groups: a handful of wrong dealers costs a handful of warps, not every group)
    k_fd_difftab<<<n_here, nt, (size_t)nt * 72 + (size_t)t * 32, s>>>((const uint32_t*)ctx->fd_sl.p, ifact, (uint32_t*)ctx->fd_coef.p, ok2, nullptr, d0, n_d, n_r, t,
                                                             cand, n_cand);
    k_fd_coefpoint<<<dim3(g_here, t), FD_NT, FD_SMEM, s>>>(d_vv, (const uint32_t*)ctx->fd_coef.p, ctx->gtab, ok2, (uint32_t*)ctx->fd_yz.p, d0, n_d, n_cols, t,
                                                           cand, n_cand);
    k_fd_coefsign<<<dim3(g_here, (batches + 3) / 4), dim3(32, 4), 0, s>>>(d_vv, (const uint32_t*)ctx->fd_yz.p, ok2, d0, n_d, n_cols, t, cand, n_cand);
    k_rs_verdicts<<<dim3(n_here, gy), 128, 0, s>>>(d_status, err, oor, ok2, poly_ok, state, cols, repaired, d0, n_r);
    ctx->launches += 11;
  }
  CK(cudaMemsetAsync(need_group, 0, groups, s));
  CK(cudaMemsetAsync(d_flags + 1, 0, 4, s));
  k_fd_need<<<(n_d + 127) / 128, 128, 0, s>>>(poly_ok, n_d, need_group, d_flags);
  k_fd_fill_ok<<<dim3(n_d, gy), 128, 0, s>>>(d_status, need_group, state, n_d, n_r);
  ctx->launches += 2;
  CK(cudaMemcpyAsync(&ctx->h_job_flags[2], repaired, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaGetLastError());
  return 0;
}

// Items (point additions) per block of the difference / extension launches.  One per block is the
// measured optimum on B200 (n=1024, t=683, N=1: extension 336 ms with 1, 344 ms with 2, 367 ms with 4 items):
// block scheduling is not what the one-addition blocks lose time on.  DKGV_FD_IPB overrides (experiments).
static inline uint32_t items_per_block(const dkgv_ctx* ctx) { return ctx->fd_ipb_force ? ctx->fd_ipb_force : 1; }

// Evaluation of dealers d0 .. d0 + n_pad - 1 (n_pad a multiple of 32: the column count of this chunk's planes) at every id;
// cols (device) = columns in ascending-id order; filter (per 32-dealer group of this chunk, or nullptr = all): the groups to evaluate
static int fd_run(dkgv_ctx* ctx, const VVView& view, uint32_t d0, uint32_t n_pad, uint32_t n_d, uint32_t n_r, uint32_t t,
                  const FdPlan& plan, const uint32_t* d_ids, const uint32_t* cols, const uint8_t* d_shares, uint8_t* d_status,
                  uint8_t* d_out48, const uint8_t* filter, cudaStream_t s) {
  const uint32_t m = plan.m, h = plan.h;
  const uint32_t n_padv = n_pad * m;  // plane width: one column per virtual dealer
  const uint32_t groups = n_pad / 32;
  const size_t ent_words = (size_t)36 * n_padv, ent_bytes = ent_words * 4;
  const size_t n_evals = (size_t)((int64_t)n_r - plan.lo + 1);
  CK(cudaEventRecord(ctx->ev_fd[0], s));
  CK(cudaEventRecord(ctx->ev_hot0, s));
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
  const uint32_t ipb = items_per_block(ctx);
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
      uint32_t cnt = (uint32_t)(k_hi - k_lo + 1);
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
      uint32_t cnt = (uint32_t)(k_hi - k_lo + 1);
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
// filter: need flags per 32-dealer group of the whole session (nullptr = every group)
static int fd_run_chunks(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint32_t* cols, const uint8_t* d_shares, uint8_t* d_status, uint8_t* d_out48,
                         const uint8_t* filter, cudaStream_t s) {
  const size_t n_evals = (size_t)((int64_t)n_r - plan.lo + 1);
  const size_t bytes_per_col = (n_evals + 4 * (size_t)plan.h) * 36 * plan.m * 4;
  size_t budget = (size_t)12 << 30;
  if (const char* e = getenv("DKGV_FD_PLANE_BUDGET_MB")) budget = (size_t)atoll(e) << 20;  // tests force several chunks
  uint32_t chunk = (uint32_t)std::min<size_t>(view.n_pad, std::max<size_t>(32, (budget / bytes_per_col) & ~(size_t)31));
  for (uint32_t d0 = 0; d0 < view.n_pad && d0 < n_d; d0 += chunk) {
    uint32_t ncols = std::min(chunk, view.n_pad - d0);
    if (int rc = fd_run(ctx, view, d0, ncols, n_d, n_r, t, plan, d_ids, cols, d_shares, d_status, d_out48, filter ? filter + d0 / 32 : nullptr, s))
      return rc;
  }
  return 0;
}

// share verdicts by evaluation for the dealer groups of `filter` (nullptr: all); cols were computed by dkgv_fd_submit
int dkgv_share_matrix_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint8_t* d_shares, uint8_t* d_status, const uint8_t* filter, cudaStream_t s) {
  return fd_run_chunks(ctx, view, n_d, n_r, t, plan, d_ids, (const uint32_t*)ctx->fd_cols.p, d_shares, d_status, nullptr, filter, s);
}

// evaluate_polynomial (dkg_math.rs:160-174) of every dealer at every id, ids a permutation of 1..n_r (h_ids: the caller's host
// copy): d_out48[d][j] = compress(f_d(ids[j])).  Used for the final keys K_j of agg_coefficients (one "dealer": the
// column sums) and the batched dkgv_feldman_eval.
int dkgv_feldman_eval_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint32_t* h_ids, uint8_t* d_out48, cudaStream_t s) {
  CK(ctx->fd_cols.reserve((size_t)n_r * 4));
  ctx->fd_cols_host.resize(n_r);  // columns in ascending-id order; must outlive the async copy
  for (uint32_t j = 0; j < n_r; j++) ctx->fd_cols_host[h_ids[j] - 1] = j;
  CK(cudaMemcpyAsync(ctx->fd_cols.p, ctx->fd_cols_host.data(), (size_t)n_r * 4, cudaMemcpyHostToDevice, s));
  return fd_run_chunks(ctx, view, n_d, n_r, t, plan, d_ids, (const uint32_t*)ctx->fd_cols.p, nullptr, nullptr, d_out48, nullptr, s);
}
