// Repair route of the consistency shortcut: a dealer whose n shares do NOT lie on one polynomial of degree < t is not sent to the
// evaluation in the exponent right away.  The shares of a dealer are a Reed-Solomon codeword over Fr (evaluations of a polynomial
// of degree < t at the n points 1..n) with errors; as long as at most tau = floor((n - t) / 2) of them are wrong the committed
// polynomial is recovered by scalar arithmetic alone:
//   k_rs_syndromes   S_j = sum_i u_i r_i i^j, j < n - t          (u_i = 1 / prod_{k != i} (i - k): the dual code's weights)
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

// tables of a shape (n, t), Montgomery form: u[i] (i = 0..n-1 for the point x = i + 1), inv[d] = 1 / d (d = 1..n-1; inv[0] unused),
// pw[i][j] = (i + 1)^j for j < n - t
__global__ void __launch_bounds__(128)
k_rs_tables(uint32_t n, uint32_t nsyn, uint32_t* __restrict__ u, uint32_t* __restrict__ inv, uint32_t* __restrict__ pw) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // prod_{k != x} (x - k) = (x-1)! * (-1)^(n-x) (n-x)!,  x = i + 1
  Fr a = one<FrParams>();
  for (uint32_t k = 2; k <= i; k++) a = mul(a, fr_from_small(k));
  for (uint32_t k = 2; k <= n - 1 - i; k++) a = mul(a, fr_from_small(k));
  if ((n - 1 - i) & 1) a = neg(a);
  fr_store(u + (size_t)i * 8, fr_inverse(a));
  fr_store(inv + (size_t)i * 8, i ? fr_inverse(fr_from_small(i)) : zero<FrParams>());
  Fr x = fr_from_small(i + 1), p = one<FrParams>();
  for (uint32_t j = 0; j < nsyn; j++) {
    fr_store(pw + ((size_t)i * nsyn + j) * 8, p);
    p = mul(p, x);
  }
}

// S_j for the dealers under repair: block = (dealer, 128 syndromes), the weighted shares w_i = u_i r_i staged in shared memory
__global__ void __launch_bounds__(128)
k_rs_syndromes(const uint32_t* __restrict__ sl, const uint8_t* __restrict__ state, const uint32_t* __restrict__ u, const uint32_t* __restrict__ pw,
               uint32_t* __restrict__ syn, uint32_t d0, uint32_t n_r, uint32_t nsyn) {
  extern __shared__ uint32_t rs_sm[];  // w[n_r][8]
  const uint32_t dl = blockIdx.x;
  if (state[d0 + dl] != RS_REPAIR) return;
  for (uint32_t i = threadIdx.x; i < n_r; i += blockDim.x) {
    Fr r = to_mont(fr_load(sl + ((size_t)dl * n_r + i) * 8));
    fr_store(rs_sm + (size_t)i * 8, mul(r, fr_load(u + (size_t)i * 8)));
  }
  __syncthreads();
  const uint32_t j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j >= nsyn) return;
  Fr acc = zero<FrParams>();
#pragma unroll 1
  for (uint32_t i = 0; i < n_r; i++) acc = add(acc, mul(fr_load(rs_sm + (size_t)i * 8), fr_load(pw + ((size_t)i * nsyn + j) * 8)));
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

// Inversion-free Berlekamp-Massey, one block per dealer under repair, thread k owns Lambda_k (k <= tau + 1).
//   d = sum_{k <= L} Lambda_k S_{r-k};  d != 0:  Lambda <- b Lambda - d z^m B  (and, when 2 L <= r: B <- old Lambda, L <- r + 1 - L, b <- d, m <- 1)
// out: lam[dl][0..tau] (Montgomery), deg[d] = L, or state -> RS_FAILED when L > tau (more wrong shares than the code corrects)
__global__ void __launch_bounds__(1024)
k_rs_bm(const uint32_t* __restrict__ syn, uint8_t* __restrict__ state, uint32_t* __restrict__ lam, uint32_t* __restrict__ deg, uint32_t d0,
        uint32_t nsyn, uint32_t tau) {
  extern __shared__ uint32_t rs_sm[];  // S[nsyn], B[blockDim.x], red[blockDim.x / 32 + 1]
  const uint32_t dl = blockIdx.x, k = threadIdx.x;
  if (state[d0 + dl] != RS_REPAIR) return;
  Fr* S = (Fr*)rs_sm;
  Fr* B = S + nsyn;
  Fr* red = B + blockDim.x;
  for (uint32_t j = k; j < nsyn; j += blockDim.x) S[j] = fr_load(syn + ((size_t)dl * nsyn + j) * 8);
  Fr c = k == 0 ? one<FrParams>() : zero<FrParams>();
  B[k] = c;
  Fr b = one<FrParams>();
  uint32_t L = 0, m = 1;
  __syncthreads();
#pragma unroll 1
  for (uint32_t r = 0; r < nsyn; r++) {
    Fr term = (k <= L && k <= r) ? mul(c, S[r - k]) : zero<FrParams>();
    Fr d = rs_block_sum(term, red);
    if (is_zero(d)) {
      m++;
      continue;
    }
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
