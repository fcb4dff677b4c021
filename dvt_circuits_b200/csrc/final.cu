// Final / partial key aggregation on the GPU (C ABI part 3):
//   agg_coefficients        crates/dkg/src/dkg_math.rs:230-248  (column sums, then K_j = P(id_j))
//   lagrange_interpolation  crates/dkg/src/dkg_math.rs:178-227  (at 0, with its error exits)
//   evaluate_polynomial     crates/dkg/src/dkg_math.rs:160-174  (over a point vector, arbitrary ids)
// These run once per ceremony on n, t ~ 10^3 inputs: thread-per-output kernels with the inlined
// formulas of g1.cuh; the verification vectors come decoded in the limb-planar layout of feldman.cuh.
#include "ctx.hpp"
#include "fdiff.cuh"

using namespace dkgv;

// share_fd.cu
bool dkgv_fd_ids_consecutive(const uint32_t* h_ids, uint32_t n_r);
int dkgv_feldman_eval_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint32_t* h_ids, uint8_t* d_out48, cudaStream_t s);

namespace {
struct ExpRm2 {
  DKGV_HD uint32_t operator()(int i) const { return consts::R_MINUS_2(i); }
};
DKGV_HD Fr fr_inv(const Fr& a) { return pow_const<FrParams>(a, ExpRm2(), 8); }
DKGV_HD Fr fr_from_u32(uint32_t x) {
  Fr r = zero<FrParams>();
  r.l[0] = x;
  return to_mont(r);
}
__device__ __forceinline__ G1Proj warp_reduce_points(G1Proj p) {
#pragma unroll 1
  for (int off = 16; off > 0; off >>= 1) {
    G1Proj q;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      q.x.l[i] = __shfl_down_sync(0xffffffffu, p.x.l[i], off);
      q.y.l[i] = __shfl_down_sync(0xffffffffu, p.y.l[i], off);
      q.z.l[i] = __shfl_down_sync(0xffffffffu, p.z.l[i], off);
    }
    p = g1_add(p, q);
  }
  return p;
}
__device__ __forceinline__ void store_affine25(uint32_t* dst, const G1Aff& a) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    dst[i] = a.x.l[i];
    dst[12 + i] = a.y.l[i];
  }
  dst[24] = a.inf;
}
__device__ __forceinline__ G1Aff load_affine25(const uint32_t* src) {
  G1Aff a;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    a.x.l[i] = src[i];
    a.y.l[i] = src[12 + i];
  }
  a.inf = src[24];
  return a;
}
}  // namespace

// decode (same kernel as the share path, duplicated symbol-locally to keep the TUs independent)
__global__ void __launch_bounds__(128)
k_decompress_vv_f(const uint8_t* __restrict__ vv, uint32_t n_d, uint32_t t, uint32_t n_pad, uint32_t* __restrict__ limbs,
                  uint8_t* __restrict__ inf, uint8_t* __restrict__ dealer_bad) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)n_pad * t) return;
  uint32_t d = (uint32_t)(p % n_pad), k = (uint32_t)(p / n_pad);
  G1Aff a;
  a.x = zero<FpParams>();
  a.y = zero<FpParams>();
  a.inf = 1;
  if (d < n_d) {
    uint32_t st = g1_decompress(vv + ((size_t)d * t + k) * 48, &a, true);
    if (st != G1_DEC_OK) dealer_bad[d] = 1;
  }
  vv_store(limbs, inf, n_pad, k, d, a);
}

// C_k = sum_d vv[d][k]: one warp per coefficient index k; padding dealers are stored as identity.
// out: 25 words per coefficient (affine Montgomery x, y, inf) + optional 48-byte encoding
__global__ void __launch_bounds__(128)
k_column_sums(VVView vv, uint32_t t, uint32_t* __restrict__ coeffs25, uint8_t* __restrict__ enc_out) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= t) return;
  G1Proj acc = g1_identity();
#pragma unroll 1
  for (uint32_t d = lane; d < vv.n_pad; d += 32) acc = g1_add_mixed(acc, vv_load(vv, warp, d));
  acc = warp_reduce_points(acc);
  if (lane == 0) {
    G1Aff a = g1_to_affine(acc);
    store_affine25(coeffs25 + (size_t)warp * 25, a);
    if (enc_out) {
      uint8_t enc[48];
      g1_compress(a, enc);
      for (int i = 0; i < 48; i++) enc_out[(size_t)warp * 48 + i] = enc[i];
    }
  }
}

// the column sums as the verification vector of ONE dealer (column 0 of a 32-wide plane, the rest identity):
// input of the finite-difference evaluation of the final keys
__global__ void __launch_bounds__(128)
k_coeffs_to_planar(const uint32_t* __restrict__ coeffs25, uint32_t t, uint32_t* __restrict__ limbs, uint8_t* __restrict__ inf) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= t * 32) return;
  uint32_t k = p >> 5, d = p & 31;
  G1Aff a;
  a.x = zero<FpParams>();
  a.y = zero<FpParams>();
  a.inf = 1;
  if (d == 0) a = load_affine25(coeffs25 + (size_t)k * 25);
  vv_store(limbs, inf, 32, k, d, a);
}

// K_j = sum_k C_k id_j^k (Horner from the top, dkg_math.rs:160-174); one thread per id
__global__ void __launch_bounds__(64)
k_eval_points_at_ids(const uint32_t* __restrict__ coeffs25, uint32_t t, const uint32_t* __restrict__ ids, uint32_t n_ids,
                     uint8_t* __restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_ids) return;
  uint32_t id = ids[j];
  G1Proj acc = g1_identity();
  if (t > 0) {
    acc = g1_from_affine(load_affine25(coeffs25 + (size_t)(t - 1) * 25));
#pragma unroll 1
    for (int k = (int)t - 2; k >= 0; k--) {
      acc = g1_mul_small(acc, id);
      acc = g1_add_mixed(acc, load_affine25(coeffs25 + (size_t)k * 25));
    }
  }
  uint8_t enc[48];
  g1_compress(g1_to_affine(acc), enc);
  for (int i = 0; i < 48; i++) out[(size_t)j * 48 + i] = enc[i];
}

// Lagrange weights at 0:  l_i = a * (x_i * prod_{j != i} (x_j - x_i))^-1,  a = prod x_j
// flags[0] |= 1 when some id is zero (a == 0), flags[1] |= 1 on a duplicate id
__global__ void __launch_bounds__(128)
k_lagrange_weights(const uint32_t* __restrict__ ids, uint32_t k, uint32_t* __restrict__ weights /*[k][8] raw*/,
                   uint32_t* __restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  Fr xi = fr_from_u32(ids[i]);
  Fr a = one<FrParams>(), b = xi;
  bool dup = false;
#pragma unroll 1
  for (uint32_t j = 0; j < k; j++) {
    Fr xj = fr_from_u32(ids[j]);
    a = mul(a, xj);
    if (j != i) {
      Fr v = sub(xj, xi);
      dup |= is_zero(v);
      b = mul(b, v);
    }
  }
  if (is_zero(a)) atomicOr(&flags[0], 1u);
  if (dup) atomicOr(&flags[1], 1u);
  Fr w = from_mont(mul(a, fr_inv(b)));
#pragma unroll
  for (int l = 0; l < 8; l++) weights[(size_t)i * 8 + l] = w.l[l];
}

// partial[i] = [l_i] y_i (full-width variable-base multiplication); bad[0] |= 1 if y_i fails to decode
__global__ void __launch_bounds__(64)
k_lagrange_terms(const uint8_t* __restrict__ pts, const uint32_t* __restrict__ weights, uint32_t k, uint32_t* __restrict__ partial /*[k][36]*/,
                 uint32_t* __restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  G1Aff y;
  if (g1_decompress(pts + (size_t)i * 48, &y, true) != G1_DEC_OK) atomicOr(&flags[2], 1u);
  G1Proj base = g1_from_affine(y), acc = g1_identity();
  bool started = false;
#pragma unroll 1
  for (int l = 7; l >= 0; l--) {
    uint32_t w = weights[(size_t)i * 8 + l];
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      if (started) acc = g1_dbl(acc);
      if ((w >> b) & 1) {
        acc = started ? g1_add(acc, base) : base;
        started = true;
      }
    }
  }
  uint32_t* o = partial + (size_t)i * 36;
#pragma unroll
  for (int l = 0; l < 12; l++) {
    o[l] = acc.x.l[l];
    o[12 + l] = acc.y.l[l];
    o[24 + l] = acc.z.l[l];
  }
}

// BlsG1::mul_scalar (crates/dkg/src/dkg_math.rs:122-127) for m independent (point, scalar) pairs:
// out[i] = compress([s_i] P_i); status OK / PANIC_BAD_G1 (P_i undecodable) / PANIC_BAD_SCALAR (s_i >= r)
__global__ void __launch_bounds__(64)
k_g1_mul_batch(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ scalars, uint8_t* __restrict__ out, uint8_t* __restrict__ status,
               uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  G1Aff y;
  uint32_t s[8];
  bool pt_ok = g1_decompress(pts + (size_t)i * 48, &y, true) == G1_DEC_OK;
  bool sc_ok = fr_raw_from_be32(s, scalars + (size_t)i * 32);
  G1Proj base = g1_from_affine(y), acc = g1_identity();
  bool started = false;
#pragma unroll 1
  for (int l = 7; l >= 0; l--) {
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      if (started) acc = g1_dbl(acc);
      if ((s[l] >> b) & 1) {
        acc = started ? g1_add(acc, base) : base;
        started = true;
      }
    }
  }
  uint8_t enc[48];
  g1_compress(g1_to_affine(acc), enc);
  for (int k = 0; k < 48; k++) out[(size_t)i * 48 + k] = (pt_ok && sc_ok) ? enc[k] : 0;
  status[i] = !sc_ok ? DKGV_PANIC_BAD_SCALAR : (!pt_ok ? DKGV_PANIC_BAD_G1 : DKGV_OK);
}

// sum of k projective points (one warp) -> 48-byte encoding
__global__ void __launch_bounds__(32) k_sum_points(const uint32_t* __restrict__ partial, uint32_t k, uint8_t* __restrict__ out) {
  uint32_t lane = threadIdx.x;
  G1Proj acc = g1_identity();
#pragma unroll 1
  for (uint32_t i = lane; i < k; i += 32) {
    G1Proj p;
    const uint32_t* o = partial + (size_t)i * 36;
#pragma unroll
    for (int l = 0; l < 12; l++) {
      p.x.l[l] = o[l];
      p.y.l[l] = o[12 + l];
      p.z.l[l] = o[24 + l];
    }
    acc = g1_add(acc, p);
  }
  acc = warp_reduce_points(acc);
  if (lane == 0) {
    uint8_t enc[48];
    g1_compress(g1_to_affine(acc), enc);
    for (int i = 0; i < 48; i++) out[i] = enc[i];
  }
}

// projective column sums of a row block, for the sharded aggregation (comm.cu): part36[k] = sum over this block's dealers of vv[d][k]
// (X, Y, Z Montgomery limbs), and after the t points a 4-word tail whose first word is 1 when some commitment failed to decode
__global__ void __launch_bounds__(128)
k_column_partials(VVView vv, uint32_t t, const uint8_t* __restrict__ dealer_bad, uint32_t n_d, uint32_t* __restrict__ part36) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= t) return;
  G1Proj acc = g1_identity();
#pragma unroll 1
  for (uint32_t d = lane; d < vv.n_pad; d += 32) acc = g1_add_mixed(acc, vv_load(vv, warp, d));
  acc = warp_reduce_points(acc);
  if (lane == 0) {
    uint32_t* o = part36 + (size_t)warp * 36;
#pragma unroll
    for (int l = 0; l < 12; l++) {
      o[l] = acc.x.l[l];
      o[12 + l] = acc.y.l[l];
      o[24 + l] = acc.z.l[l];
    }
  }
  if (warp == 0) {
    uint32_t bad = 0;
    for (uint32_t d = lane; d < n_d; d += 32) bad |= dealer_bad[d];
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (lane == 0) {
      uint32_t* tail = part36 + (size_t)t * 36;
      tail[0] = bad ? 1u : 0u;
      tail[1] = tail[2] = tail[3] = 0;
    }
  }
}

// ============================================================================ C ABI
// decode + subgroup-check the rows vv [n][t][48] (host) into the session planes; asynchronous on s
static int decode_rows(dkgv_ctx* ctx, uint32_t n, uint32_t t, const uint8_t* vv, cudaStream_t s, VVView* view) {
  uint32_t n_pad = (n + 31) & ~31u, tt = t ? t : 1;
  size_t vvb = (size_t)n * t * 48;
  CK(ctx->in_a.reserve(vvb ? vvb : 1));
  CK(ctx->vv_limbs.reserve((size_t)tt * 24 * n_pad * 4));
  CK(ctx->vv_inf.reserve((size_t)tt * n_pad));
  CK(ctx->dealer_bad.reserve(n_pad));
  if (vvb) CK(cudaMemcpyAsync(ctx->in_a.p, vv, vvb, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(ctx->dealer_bad.p, 0, n_pad, s));
  if (t) {
    size_t total = (size_t)n_pad * t;
    k_decompress_vv_f<<<(unsigned)((total + 127) / 128), 128, 0, s>>>((const uint8_t*)ctx->in_a.p, n, t, n_pad, (uint32_t*)ctx->vv_limbs.p,
                                                                    (uint8_t*)ctx->vv_inf.p, (uint8_t*)ctx->dealer_bad.p);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  *view = VVView{(const uint32_t*)ctx->vv_limbs.p, (const uint8_t*)ctx->vv_inf.p, n_pad};
  return 0;
}

int dkgv_column_partials_internal(dkgv_ctx* ctx, uint32_t n_local, uint32_t t, const uint8_t* h_vv, uint32_t* d_partial36, cudaStream_t s) {
  VVView view;
  if (int rc = decode_rows(ctx, n_local, t, h_vv, s, &view)) return rc;
  k_column_partials<<<(t * 32 + 127) / 128, 128, 0, s>>>(view, t, (const uint8_t*)ctx->dealer_bad.p, n_local, d_partial36);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

// K_j = evaluate_polynomial(C, ids[j]) from the column sums in ctx->scratch_a (25-word affine records); keys_out: host
int dkgv_keys_from_coeffs_internal(dkgv_ctx* ctx, uint32_t t, const uint32_t* ids, uint32_t n_ids, uint8_t* keys_out, cudaStream_t s) {
  CK(ctx->in_b.reserve((size_t)n_ids * 4));
  CK(ctx->out_b.reserve((size_t)n_ids * 48));
  CK(cudaMemcpyAsync(ctx->in_b.p, ids, (size_t)n_ids * 4, cudaMemcpyHostToDevice, s));
  // ids 1..n_ids (the ranks of a ceremony): the n final keys come from t-ish Horner seeds + finite differences
  // (csrc/fdiff.cuh) instead of n Horner chains of t-1 steps; same points, same encodings
  FdPlan plan{};
  bool use_fd = false;
  if (ctx->share_path != DKGV_SHARE_PATH_HORNER && t >= 2 && n_ids >= 3 && n_ids <= 65535 && dkgv_fd_ids_consecutive(ids, n_ids)) {
    plan = fd_make_plan(t, n_ids, ctx->share_parts);
    use_fd = plan.cost_fd != ~0ull && (plan.use || ctx->share_path == DKGV_SHARE_PATH_FDIFF);
  }
  if (use_fd) {
    CK(ctx->scratch_b.reserve((size_t)t * 24 * 32 * 4));
    CK(ctx->scratch_c.reserve((size_t)t * 32));
    k_coeffs_to_planar<<<(t * 32 + 127) / 128, 128, 0, s>>>((const uint32_t*)ctx->scratch_a.p, t, (uint32_t*)ctx->scratch_b.p,
                                                            (uint8_t*)ctx->scratch_c.p);
    ctx->launches++;
    VVView one{(const uint32_t*)ctx->scratch_b.p, (const uint8_t*)ctx->scratch_c.p, 32};
    if (int rc = dkgv_feldman_eval_fd(ctx, one, 1, n_ids, t, plan, (const uint32_t*)ctx->in_b.p, ids, (uint8_t*)ctx->out_b.p, s)) return rc;
  } else {
    k_eval_points_at_ids<<<(n_ids + 63) / 64, 64, 0, s>>>((const uint32_t*)ctx->scratch_a.p, t, (const uint32_t*)ctx->in_b.p, n_ids,
                                                         (uint8_t*)ctx->out_b.p);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(keys_out, ctx->out_b.p, (size_t)n_ids * 48, cudaMemcpyDeviceToHost, s));
  return 0;
}

extern "C" int dkgv_agg_final_keys(dkgv_ctx* ctx, uint32_t n, uint32_t t, const uint8_t* vv, const uint32_t* ids, uint32_t n_ids,
                                   uint8_t* coeff_out, uint8_t* keys_out, uint8_t* status) {
  if (!ctx || !status) return -1;
  *status = DKGV_OK;
  if (n == 0) return dkgv_fail(ctx, "agg_coefficients needs at least one verification vector (the reference indexes vv[0])");
  if ((t && !vv) || (n_ids && (!ids || !keys_out))) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint32_t n_pad = (n + 31) & ~31u, tt = t ? t : 1;
  CK(ctx->scratch_a.reserve((size_t)tt * 25 * 4));
  CK(ctx->out_a.reserve((size_t)tt * 48));
  VVView view;
  if (int rc = decode_rows(ctx, n, t, vv, s, &view)) return rc;
  if (t) {
    k_column_sums<<<(t * 32 + 127) / 128, 128, 0, s>>>(view, t, (uint32_t*)ctx->scratch_a.p, (uint8_t*)ctx->out_a.p);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  if (n_ids)
    if (int rc = dkgv_keys_from_coeffs_internal(ctx, t, ids, n_ids, keys_out, s)) return rc;
  if (coeff_out && t) CK(cudaMemcpyAsync(coeff_out, ctx->out_a.p, (size_t)t * 48, cudaMemcpyDeviceToHost, s));
  std::string bad(n_pad, 0);
  CK(cudaMemcpyAsync(&bad[0], ctx->dealer_bad.p, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (uint32_t i = 0; i < n; i++)
    if (bad[i]) *status = DKGV_PANIC_BAD_G1;
  return 0;
}

extern "C" int dkgv_lagrange_at_zero(dkgv_ctx* ctx, uint32_t k, const uint8_t* pts, const uint32_t* ids, uint8_t* out, uint8_t* status) {
  if (!ctx || !status || !out) return -1;
  if (k == 0) {
    *status = DKGV_ERR_LEN;  // dkg_math.rs:183-188
    return 0;
  }
  if (!pts || !ids) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve((size_t)k * 48));
  CK(ctx->in_b.reserve((size_t)k * 4));
  CK(ctx->scratch_a.reserve((size_t)k * 32));
  CK(ctx->scratch_b.reserve((size_t)k * 36 * 4));
  CK(ctx->scratch_c.reserve(16));
  CK(ctx->out_a.reserve(48));
  CK(cudaMemcpyAsync(ctx->in_a.p, pts, (size_t)k * 48, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, ids, (size_t)k * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(ctx->scratch_c.p, 0, 16, s));
  uint32_t flags[4] = {0, 0, 0, 0};
  if (k == 1) {
    // the reference returns y_0 untouched (dkg_math.rs:189-191); still decode it to report a bad point
    uint8_t dst = 0;
    CK(cudaStreamSynchronize(s));
    int rc = dkgv_g1_decompress_check(ctx, 1, pts, &dst);
    if (rc) return rc;
    *status = dst ? DKGV_PANIC_BAD_G1 : DKGV_OK;
    for (int i = 0; i < 48; i++) out[i] = pts[i];
    return 0;
  }
  k_lagrange_weights<<<(k + 127) / 128, 128, 0, s>>>((const uint32_t*)ctx->in_b.p, k, (uint32_t*)ctx->scratch_a.p, (uint32_t*)ctx->scratch_c.p);
  k_lagrange_terms<<<(k + 63) / 64, 64, 0, s>>>((const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->scratch_a.p, k, (uint32_t*)ctx->scratch_b.p,
                                               (uint32_t*)ctx->scratch_c.p);
  k_sum_points<<<1, 32, 0, s>>>((const uint32_t*)ctx->scratch_b.p, k, (uint8_t*)ctx->out_a.p);
  ctx->launches += 3;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, 48, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(flags, ctx->scratch_c.p, 16, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  // exit order of the reference: points are decoded by the caller before lagrange_interpolation runs
  // (verification.rs:282-291,311-316), then zero id (dkg_math.rs:201-206), then duplicate id (:212-217)
  if (flags[2]) *status = DKGV_PANIC_BAD_G1;
  else if (flags[0]) *status = DKGV_ERR_ZERO_ID;
  else if (flags[1]) *status = DKGV_ERR_DUP_ID;
  else *status = DKGV_OK;
  return 0;
}

// evaluate_polynomial over an arbitrary point vector at several ids (dkg_math.rs:160-174);
// used by compute_pubkey_share (verification.rs:523-551, quirk Q1: Horner over the final keys)
extern "C" int dkgv_eval_points(dkgv_ctx* ctx, uint32_t t, const uint8_t* coeffs, const uint32_t* ids, uint32_t n_ids, uint8_t* out,
                                uint8_t* status) {
  if (!ctx || !status) return -1;
  *status = DKGV_OK;
  if (n_ids == 0) return 0;
  if ((t && !coeffs) || !ids || !out) return dkgv_fail(ctx, "null pointer argument");
  // a one-"dealer" session: decode + column sum over a single row is the identity map
  return dkgv_agg_final_keys(ctx, 1, t, coeffs, ids, n_ids, nullptr, out, status);
}

extern "C" int dkgv_g1_mul_batch(dkgv_ctx* ctx, uint32_t m, const uint8_t* pts, const uint8_t* scalars, uint8_t* out, uint8_t* status) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!pts || !scalars || !out || !status) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve((size_t)m * 48));
  CK(ctx->in_b.reserve((size_t)m * 32));
  CK(ctx->out_a.reserve((size_t)m * 48));
  CK(ctx->out_b.reserve(m));
  CK(cudaMemcpyAsync(ctx->in_a.p, pts, (size_t)m * 48, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, scalars, (size_t)m * 32, cudaMemcpyHostToDevice, s));
  k_g1_mul_batch<<<(m + 63) / 64, 64, 0, s>>>((const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, (uint8_t*)ctx->out_a.p,
                                             (uint8_t*)ctx->out_b.p, m);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, (size_t)m * 48, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(status, ctx->out_b.p, m, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}
