// Repair route of the consistency shortcut: a dealer whose n shares do NOT lie on one polynomial of degree < t is not sent to the
// evaluation in the exponent right away.  The shares of a dealer are a Reed-Solomon codeword over Fr (evaluations of a polynomial
// of degree < t at the n points 1..n) with errors; as long as at most tau = floor((n - t) / 2) of them are wrong the committed
// polynomial is recovered by scalar arithmetic alone:
//   k_rs_gdiff / k_rs_syndromes   S_j = sum_i u_i r_i i^j, j < n - t (u_i = 1 / prod_{k != i} (i - k): the dual code's weights), out of the
//                    difference table of the shares carried on to order n - 1 and a fixed triangular matrix of the shape
//   k_rs_bm          inversion-free Berlekamp-Massey on S -> error locator Lambda, degree L = number of wrong shares
//   k_rs_chien       the positions x with Lambda*(x) = sum_k Lambda_k x^(L-k) = 0
//   k_rs_forney      the error values at the located positions by Forney's formula; share - error replaces the wrong values in the table
// after which the SAME exact conditions as for an honest dealer are checked on the corrected table - the t-th differences of all n
// values vanish (k_fd_difftab, which also yields the monomial coefficients) and compress(G * p_k) == C_k for every k
// (k_fd_coefpoint / k_fd_coefsign).  Only then the verdicts are written: located position with p(x) != share -> SHARE_MISMATCH, every
// other share OK.  The decoder merely PROPOSES the corrected table; exactness rests on the two conditions, as for the honest path: p equals the
// committed polynomial, so a share is valid iff it equals p(x).  A dealer the decoder cannot repair (more than tau wrong shares,
// or a proposal that fails a condition) goes to the evaluation as before.  verify_seed_exchange_commitment
// (crates/dkg/src/verification.rs:68-149) per share, unchanged verdicts; no randomness anywhere.
#pragma once
#include "fdiff.cuh"

namespace dkgv {

// dealer states of a submitted job
enum : uint8_t {
  RS_IDLE = 0,       // settled (or failed) by the first pass: the repair kernels skip it
  RS_REPAIR = 1,     // conditions (1) / (2) failed: decode
  RS_CANDIDATE = 2,  // a corrected share table is in place: second pass of the conditions
  RS_FAILED = 3      // not repairable here: evaluation
};

DKGV_HD Fr fr_from_small(uint32_t x) {
  Fr r = zero<FrParams>();
  r.l[0] = x;
  return to_mont(r);
}
struct ExpRm2S {
  DKGV_HD uint32_t operator()(int i) const { return consts::R_MINUS_2(i); }
};
DKGV_HD Fr fr_inverse(const Fr& a) { return pow_const<FrParams>(a, ExpRm2S(), 8); }

DKGV_HD Fr fr_load(const uint32_t* p) {
  Fr v;
#pragma unroll
  for (int l = 0; l < 8; l++) v.l[l] = p[l];
  return v;
}
DKGV_HD void fr_store(uint32_t* p, const Fr& v) {
#pragma unroll
  for (int l = 0; l < 8; l++) p[l] = v.l[l];
}

}  // namespace dkgv

#if defined(__CUDACC__)
using namespace dkgv;

// table of a shape (n, t), Montgomery form: u[i] = 1 / prod_{k != i} (x_i - x_k) for the point x_i = i + 1 (the dual code's weights)
__global__ void __launch_bounds__(128) k_rs_tables(uint32_t n, uint32_t* __restrict__ u) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // prod_{k != x} (x - k) = (x-1)! * (-1)^(n-x) (n-x)!,  x = i + 1
  Fr a = one<FrParams>();
  for (uint32_t k = 2; k <= i; k++) a = mul(a, fr_from_small(k));
  for (uint32_t k = 2; k <= n - 1 - i; k++) a = mul(a, fr_from_small(k));
  if ((n - 1 - i) & 1) a = neg(a);
  fr_store(u + (size_t)i * 8, fr_inverse(a));
}

// The syndromes come out of the DIFFERENCE TABLE of the shares, not out of n (n - t) products per dealer.  With the weights above
//   S_j = sum_i u_i r_i x_i^j = Delta^(n-1)[x^j r(x)](1) / (n-1)!
// (u_i = (-1)^(n-i) C(n-1, i-1) / (n-1)!), and by the Leibniz rule for differences, x^j having degree j < n - t,
//   S_j = sum_{k <= j} M[j][k] G_k,   G_k = Delta^(n-1-k) r(1 + k),   M[j][k] = C(n-1, k) Delta^k[x^j](1) / (n-1)!.
// G_k is the LAST entry of the table after n - 1 - k rounds - orders t .. n-1, i.e. the table of condition (2) simply carried on
// (k_rs_gdiff: (n - t)^2 / 2 more subtractions) - and M is a fixed triangular matrix of the shape: (n - t)^2 / 2 products per
// dealer instead of n (n - t).
// k_rs_mtab: A(j, k) = Delta^k[x^j](1) by A(j+1, k) = (1 + k) A(j, k) + k A(j, k-1), A(0, 0) = 1, one block, thread(s) per k, a
// round per j; out mt[k][j] (k-major: a thread per j reads consecutive words), Montgomery form, entries k > j are never read.
__global__ void __launch_bounds__(1024) k_rs_mtab(uint32_t n, uint32_t nsyn, uint32_t* __restrict__ mt, uint32_t* __restrict__ scratch) {
  // scratch: A[2][nsyn] (global, double-buffered), cf[nsyn]
  Fr* A0 = (Fr*)scratch;
  Fr* A1 = A0 + nsyn;
  Fr* cf = A1 + nsyn;
  for (uint32_t k = threadIdx.x; k < nsyn; k += blockDim.x) {
    // C(n-1, k) / (n-1)! = 1 / (k! (n-1-k)!)
    Fr a = one<FrParams>();
    for (uint32_t m = 2; m <= k; m++) a = mul(a, fr_from_small(m));
    for (uint32_t m = 2; m <= n - 1 - k; m++) a = mul(a, fr_from_small(m));
    cf[k] = fr_inverse(a);
    A0[k] = k == 0 ? one<FrParams>() : zero<FrParams>();
  }
  __syncthreads();
  Fr* cur = A0;
  Fr* nxt = A1;
#pragma unroll 1
  for (uint32_t j = 0; j < nsyn; j++) {
    for (uint32_t k = threadIdx.x; k < nsyn; k += blockDim.x) {
      const Fr a = cur[k];
      fr_store(mt + ((size_t)k * nsyn + j) * 8, mul(a, cf[k]));
      Fr v = mul(a, fr_from_small(k + 1));
      if (k) v = add(v, mul(cur[k - 1], fr_from_small(k)));
      nxt[k] = v;
    }
    __syncthreads();
    Fr* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
}

// G_k for the dealers under repair: the difference table of k_fd_difftab (same lazy per-thread routines) carried on to order n - 1;
// the thread that owns the last entry stores it (canonical) after every round r >= t: g[dl][n - 1 - r].  One block per dealer.
__global__ void __launch_bounds__(1024)
k_rs_gdiff(const uint32_t* __restrict__ sl, const uint8_t* __restrict__ state, uint32_t* __restrict__ g, uint32_t d0, uint32_t n_r, uint32_t t,
           uint32_t nsyn) {
  extern __shared__ uint32_t rs_sm[];  // pub[2][9 * blockDim.x]
  const uint32_t dl = blockIdx.x, i = threadIdx.x, nt = blockDim.x;
  if (state[d0 + dl] != RS_REPAIR) return;
  const uint32_t k0 = 2 * i, k1 = 2 * i + 1, last = n_r - 1;
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
  for (uint32_t r = 1; r < n_r; r++) {
    uint32_t* pr = rs_sm + (size_t)(r & 1) * 9 * nt;
    if (dt1_publishes(i, r)) lz_publish(pr, nt, i, p.b);
    __syncthreads();
    const bool red = --left == 0;
    if (red) left = DT1_PERIOD;
    if (dt1_active(i, r)) dt1_step(p, i, r, red, pr, nt);
    if (r >= t && (k0 == last || k1 == last)) {
      Lz v = k0 == last ? p.a : p.b;
      lz_reduce(v, true);
      fr_store(g + ((size_t)dl * nsyn + (last - r)) * 8, lz_low(v));
    }
  }
}

// S_j = sum_{k <= j} M[j][k] G_k, row0 <= j < rows, for the dealers under repair the first stage of k_rs_bm has not finished: block = (dealer,
// 128 syndromes), G (to Montgomery form) staged in shared memory; the longest rows first (blockIdx.y counts down)
__global__ void __launch_bounds__(128)
k_rs_syndromes(const uint32_t* __restrict__ g, const uint8_t* __restrict__ state, const uint8_t* __restrict__ done, const uint32_t* __restrict__ mt,
               uint32_t* __restrict__ syn, uint32_t d0, uint32_t nsyn, uint32_t row0, uint32_t rows) {
  extern __shared__ uint32_t rs_sm[];  // G[rows][8]
  const uint32_t dl = blockIdx.x;
  if (state[d0 + dl] != RS_REPAIR || done[dl]) return;
  const uint32_t yb = gridDim.y - 1 - blockIdx.y, jmax = min(rows, row0 + (yb + 1) * blockDim.x);
  for (uint32_t k = threadIdx.x; k < jmax; k += blockDim.x) fr_store(rs_sm + (size_t)k * 8, to_mont(fr_load(g + ((size_t)dl * nsyn + k) * 8)));
  __syncthreads();
  const uint32_t j = row0 + yb * blockDim.x + threadIdx.x;
  if (j >= rows) return;
  Fr acc = zero<FrParams>();
#pragma unroll 1
  for (uint32_t k = 0; k <= j; k++) acc = add(acc, mul(fr_load(rs_sm + (size_t)k * 8), fr_load(mt + ((size_t)k * nsyn + j) * 8)));
  fr_store(syn + ((size_t)dl * nsyn + j) * 8, acc);
}

// block-wide sum of one Fr per thread (Montgomery or canonical alike); every thread gets the result.  red: blockDim.x / 32 + 1 entries.
__device__ __forceinline__ Fr rs_block_sum(Fr v, Fr* red) {
#pragma unroll 1
  for (int off = 16; off > 0; off >>= 1) {
    Fr o;
#pragma unroll
    for (int l = 0; l < 8; l++) o.l[l] = __shfl_down_sync(0xffffffffu, v.l[l], off);
    v = add(v, o);
  }
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();  // red may still be read from the previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    Fr s = lane < nw ? red[lane] : zero<FrParams>();
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) {
      Fr o;
#pragma unroll
      for (int l = 0; l < 8; l++) o.l[l] = __shfl_down_sync(0xffffffffu, s.l[l], off);
      s = add(s, o);
    }
    if (lane == 0) red[nw] = s;
  }
  __syncthreads();
  return red[nw];
}

constexpr uint32_t RS_BM_QUIET = 16;
constexpr uint32_t RS_STAGE1 = 96;  // syndromes of the first stage: finishes dealers with up to (96 - 16) / 2 = 40 wrong shares
// Inversion-free Berlekamp-Massey, one block per dealer under repair, thread k owns Lambda_k (k <= tau + 1).
//   d = sum_{k <= L} Lambda_k S_{r-k};  d != 0:  Lambda <- b Lambda - d z^m B  (and, when 2 L <= r: B <- old Lambda, L <- r + 1 - L, b <- d, m <- 1)
// out: lam[dl][0..tau] (Montgomery), deg[d] = L, or state -> RS_FAILED when L > tau (more wrong shares than the code corrects)
// Two stages: the first sees only the first `avail` < nsyn syndromes - a dealer with few wrong shares is finished there (done[dl] = 1)
// and costs neither the long rows of the syndrome product nor the long recurrence; an unfinished dealer parks its state in `park`
// (per dealer: blockDim.x x {c, B} then b, then L, m, quiet) and the second stage (first = avail of the first stage, avail == nsyn)
// resumes from it.
__global__ void __launch_bounds__(1024)
k_rs_bm(const uint32_t* __restrict__ syn, uint8_t* __restrict__ state, uint32_t* __restrict__ lam, uint32_t* __restrict__ deg, uint8_t* __restrict__ done,
        uint32_t* __restrict__ park, uint32_t d0, uint32_t nsyn, uint32_t first, uint32_t avail, uint32_t tau) {
  extern __shared__ uint32_t rs_sm[];  // S[avail], B[blockDim.x], red[blockDim.x / 32 + 1]
  const uint32_t dl = blockIdx.x, k = threadIdx.x;
  if (state[d0 + dl] != RS_REPAIR || done[dl]) return;
  Fr* S = (Fr*)rs_sm;
  Fr* B = S + avail;
  Fr* red = B + blockDim.x;
  uint32_t* pk = park + (size_t)dl * ((size_t)blockDim.x * 16 + 16);
  for (uint32_t j = k; j < avail; j += blockDim.x) S[j] = fr_load(syn + ((size_t)dl * nsyn + j) * 8);
  Fr c = k == 0 ? one<FrParams>() : zero<FrParams>();
  Fr b = one<FrParams>();
  uint32_t L = 0, m = 1, quiet = 0;  // quiet: consecutive rounds without a discrepancy
  if (first) {
    c = fr_load(pk + (size_t)k * 16);
    B[k] = fr_load(pk + (size_t)k * 16 + 8);
    b = fr_load(pk + (size_t)blockDim.x * 16);
    L = pk[(size_t)blockDim.x * 16 + 8], m = pk[(size_t)blockDim.x * 16 + 9], quiet = pk[(size_t)blockDim.x * 16 + 10];
  } else {
    B[k] = c;
  }
  __syncthreads();
  bool settled = false;
#pragma unroll 1
  for (uint32_t r = first; r < avail; r++) {
    // The locator is final once 2 L syndromes have gone in and the recurrence keeps predicting the next ones; RS_BM_QUIET predicted
    // syndromes are taken as enough.  Stopping early cannot cost exactness - the decoder only proposes, the second pass decides -
    // only, against shares crafted to fool the stop, the repair of that dealer (it goes to the evaluation).
    if (quiet >= RS_BM_QUIET && 2 * L <= r) {
      settled = true;
      break;
    }
    Fr term = (k <= L && k <= r) ? mul(c, S[r - k]) : zero<FrParams>();
    Fr d = rs_block_sum(term, red);
    if (is_zero(d)) {
      m++;
      quiet++;
      continue;
    }
    quiet = 0;
    Fr shifted = k >= m ? B[k - m] : zero<FrParams>();
    Fr nc = sub(mul(b, c), mul(d, shifted));
    const bool grow = 2 * L <= r;
    __syncthreads();  // every thread has read B[k - m]
    if (grow) {
      B[k] = c;
      L = r + 1 - L;
      b = d;
      m = 1;
    } else {
      m++;
    }
    c = nc;
    __syncthreads();
  }
  if (!settled && avail < nsyn) {  // first stage, not finished: the second stage resumes here with all the syndromes
    fr_store(pk + (size_t)k * 16, c);
    fr_store(pk + (size_t)k * 16 + 8, B[k]);
    if (k == 0) {
      fr_store(pk + (size_t)blockDim.x * 16, b);
      pk[(size_t)blockDim.x * 16 + 8] = L, pk[(size_t)blockDim.x * 16 + 9] = m, pk[(size_t)blockDim.x * 16 + 10] = quiet;
    }
    return;
  }
  if (k == 0) done[dl] = 1;
  if (L > tau) {
    if (k == 0) state[d0 + dl] = RS_FAILED;
    return;
  }
  if (k <= tau) fr_store(lam + ((size_t)dl * (tau + 1) + k) * 8, c);
  if (k == 0) deg[d0 + dl] = L;
}

// err[dl][x - 1] = 1 where Lambda*(x) = sum_k Lambda_k x^(L - k) vanishes; cnt[d] counts them
__global__ void __launch_bounds__(128)
k_rs_chien(const uint32_t* __restrict__ lam, const uint32_t* __restrict__ deg, const uint8_t* __restrict__ state, uint8_t* __restrict__ err,
           uint32_t* __restrict__ cnt, uint32_t d0, uint32_t n_r, uint32_t tau) {
  const uint32_t dl = blockIdx.x, xi = blockIdx.y * blockDim.x + threadIdx.x;
  if (state[d0 + dl] != RS_REPAIR || xi >= n_r) return;
  const uint32_t L = deg[d0 + dl];
  Fr x = fr_from_small(xi + 1), v = zero<FrParams>();
  const uint32_t* lp = lam + (size_t)dl * (tau + 1) * 8;
#pragma unroll 1
  for (uint32_t k = 0; k <= L; k++) v = add(mul(v, x), fr_load(lp + (size_t)k * 8));
  const bool root = is_zero(v);
  err[(size_t)dl * n_r + xi] = root ? 1 : 0;
  if (root) atomicAdd(&cnt[d0 + dl], 1u);
}

// Forney's formula: the VALUES of the located errors straight from the syndromes and the locator, no interpolation.  With
// S(z) = sum_j S_j z^j = sum_i Y_i / (1 - X_i z) (Y_i = u_i eps_i, X_i the wrong positions, eps_i = share - p(X_i)) and
// Lambda(z) = lambda_0 prod_i (1 - X_i z):  Omega(z) = S(z) Lambda(z) mod z^L,  Y_i = -X_i Omega(1 / X_i) / Lambda'(1 / X_i).  Both sides are
// evaluated through their reversed polynomials (the common factor X_i^-(L-1) cancels - no inverse of a position is needed):
//   Om*(x) = sum_{k < L} Omega_k x^(L-1-k),  Omega_k = sum_{m <= k} Lambda_m S_{k-m};      D*(x) = sum_{1 <= k <= L} k Lambda_k x^(L-k)
//   eps_i = -X_i Om*(X_i) / (D*(X_i) u_i),   corrected share = share - eps_i.
// One block per dealer under repair, one thread per coefficient / per error: O(L^2) products + one inversion per error, against the
// O(t^2) of an interpolation through t good shares.  The located count must equal the locator's degree.  eps_i == 0: the locator was
// wrong about this position (err <- 0; the second pass then decides about the dealer, exactly).  A vanishing D*(X_i) (a repeated
// root) -> RS_FAILED.
__global__ void __launch_bounds__(256)
k_rs_forney(uint32_t* __restrict__ sl, uint8_t* __restrict__ err, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ deg,
            uint8_t* __restrict__ state, const uint32_t* __restrict__ syn, const uint32_t* __restrict__ lam, const uint32_t* __restrict__ u,
            uint32_t d0, uint32_t n_r, uint32_t nsyn, uint32_t tau) {
  extern __shared__ uint32_t rs_sm[];  // LAM[tau + 1], SY[tau], OM[tau], DL[tau + 1], list[n_r]
  __shared__ uint32_t n_err, failed;
  const uint32_t dl = blockIdx.x, tid = threadIdx.x;
  if (state[d0 + dl] != RS_REPAIR) return;
  const uint32_t L = deg[d0 + dl];
  if (cnt[d0 + dl] != L || L == 0 || L > tau) {  // the locator does not split over 1..n: not a correctable error pattern
    if (tid == 0) state[d0 + dl] = RS_FAILED;
    return;
  }
  Fr* LAM = (Fr*)rs_sm;
  Fr* SY = LAM + (tau + 1);
  Fr* OM = SY + tau;
  Fr* DL = OM + tau;
  uint32_t* list = (uint32_t*)(DL + (tau + 1));
  if (tid == 0) n_err = 0, failed = 0;
  for (uint32_t k = tid; k <= L; k += blockDim.x) {
    Fr l = fr_load(lam + ((size_t)dl * (tau + 1) + k) * 8);
    LAM[k] = l;
    DL[k] = mul(l, fr_from_small(k));
  }
  for (uint32_t k = tid; k < L; k += blockDim.x) SY[k] = fr_load(syn + ((size_t)dl * nsyn + k) * 8);
  __syncthreads();
  for (uint32_t x = tid; x < n_r; x += blockDim.x)
    if (err[(size_t)dl * n_r + x]) list[atomicAdd(&n_err, 1u)] = x;
  for (uint32_t k = tid; k < L; k += blockDim.x) {
    Fr a = zero<FrParams>();
#pragma unroll 1
    for (uint32_t m = 0; m <= k; m++) a = add(a, mul(LAM[m], SY[k - m]));
    OM[k] = a;
  }
  __syncthreads();
#pragma unroll 1
  for (uint32_t e = tid; e < n_err; e += blockDim.x) {
    const uint32_t xi = list[e];
    const Fr x = fr_from_small(xi + 1);
    Fr om = zero<FrParams>(), dd = zero<FrParams>();
#pragma unroll 1
    for (uint32_t k = 0; k < L; k++) {
      om = add(mul(om, x), OM[k]);
      dd = add(mul(dd, x), DL[k + 1]);
    }
    const Fr den = mul(dd, fr_load(u + (size_t)xi * 8));
    if (is_zero(den)) {
      failed = 1;
      continue;
    }
    const Fr eps = from_mont(neg(mul(mul(x, om), fr_inverse(den))));
    uint32_t* sp = sl + ((size_t)dl * n_r + xi) * 8;
    if (is_zero(eps))
      err[(size_t)dl * n_r + xi] = 0;
    else
      fr_store(sp, sub(fr_load(sp), eps));  // canonical residues: the modular subtraction does not care about the form
  }
  __syncthreads();
  if (failed && tid == 0) state[d0 + dl] = RS_FAILED;
}

// after the correction: the dealers under repair become candidates of the second pass (ok2 = 1, listed in cand[0 .. *n_cand)),
// everyone else 0
__global__ void __launch_bounds__(128)
k_rs_stage(uint8_t* __restrict__ state, uint8_t* __restrict__ ok2, uint32_t* __restrict__ cand, uint32_t* __restrict__ n_cand, uint32_t d0,
           uint32_t n_here) {
  uint32_t dl = blockIdx.x * blockDim.x + threadIdx.x;
  if (dl >= n_here) return;
  const bool c = state[d0 + dl] == RS_REPAIR;
  ok2[d0 + dl] = c ? 1 : 0;
  if (c) {
    state[d0 + dl] = RS_CANDIDATE;
    cand[atomicAdd(n_cand, 1u)] = dl;
  }
}

// verdicts of the dealers the second pass confirmed (ok2 still 1): located -> SHARE_MISMATCH, out of range keeps SECRET_RANGE, else OK;
// the dealer counts as settled (poly_ok <- 1).  block = (dealer, 128 ids)
__global__ void __launch_bounds__(128)
k_rs_verdicts(uint8_t* __restrict__ status, const uint8_t* __restrict__ err, const uint8_t* __restrict__ oor, const uint8_t* __restrict__ ok2,
              uint8_t* __restrict__ poly_ok, uint8_t* __restrict__ state, const uint32_t* __restrict__ cols, uint32_t* __restrict__ repaired,
              uint32_t d0, uint32_t n_r) {
  const uint32_t dl = blockIdx.x, xi = blockIdx.y * blockDim.x + threadIdx.x;
  const uint32_t d = d0 + dl;
  if (state[d] != RS_CANDIDATE) return;
  if (!ok2[d]) {
    if (xi == 0) state[d] = RS_FAILED;  // (other threads of the dealer return on either value)
    return;
  }
  if (xi >= n_r) return;
  uint32_t c = cols[xi];
  if (c >= n_r) c = 0;
  const size_t e = (size_t)dl * n_r + xi;
  status[(size_t)d * n_r + c] = oor[e] ? DKGV_SLASHABLE_SECRET_RANGE : (err[e] ? DKGV_SLASHABLE_SHARE_MISMATCH : DKGV_OK);
  if (xi == 0) {
    poly_ok[d] = 1;
    atomicAdd(repaired, 1u);
  }
}
#endif
