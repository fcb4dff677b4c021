// Flow-level batch entry points of the C ABI (part 4): many items of one proof type over one session in a handful of launches.
//   dkgv_bad_partial_key_verify_batch   prove_wrong_final_key_generation, crates/dkg/src/verification.rs:422-466 with
//                                       verify_expected_key :399-420 and compute_pubkey_share :523-551 (quirk Q1), as the guest
//                                       crates/bad_parial_key_prove/src/main.rs:16-51 runs it - for m items over ONE session.
// The hash / sort / identity-signature part of the flow (verification.rs:428-438) is per session or per item SHA-256 / ECDSA work
// and stays with the host (host/dkg_host.cpp); this file is everything that touches the curve.
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "g1.cuh"

using namespace dkgv;

// status of item i in the reference's order of checks (verification.rs:440-463):
//   key undecodable -> SLASHABLE_BAD_PK, signature undecodable -> SLASHABLE_BAD_SIG, pairing equality false -> SLASHABLE_SIG_INVALID,
//   then verify_expected_key: a session with an undecodable commitment panics there (PANIC_BAD_G1), expected != key -> KEY_MISMATCH
__global__ void __launch_bounds__(128)
k_bpk_status(const uint8_t* __restrict__ pair_st, const uint8_t* __restrict__ pk_st, const uint8_t* __restrict__ sig_st,
             const uint32_t* __restrict__ perp, const uint8_t* __restrict__ pk, const uint8_t* __restrict__ expected, uint8_t session_status,
             uint8_t* __restrict__ status, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  uint8_t st;
  if (pk_st[i] != G1_DEC_OK) {
    st = DKGV_SLASHABLE_BAD_PK;
  } else if (sig_st[i] != G1_DEC_OK) {
    st = DKGV_SLASHABLE_BAD_SIG;
  } else if (pair_st[i] != DKGV_OK) {
    st = DKGV_SLASHABLE_SIG_INVALID;
  } else if (session_status != DKGV_OK) {
    st = session_status;
  } else {
    const uint32_t* e = (const uint32_t*)(expected + (size_t)perp[i] * 48);
    const uint32_t* k = (const uint32_t*)(pk + (size_t)i * 48);
    uint32_t diff = 0;
#pragma unroll
    for (int w = 0; w < 12; w++) diff |= e[w] ^ k[w];
    st = diff ? DKGV_SLASHABLE_KEY_MISMATCH : DKGV_OK;
  }
  status[i] = st;
}

extern "C" int dkgv_bad_partial_key_verify_batch(dkgv_ctx* ctx, uint32_t n, uint32_t t, const uint8_t* vv, uint32_t m, const uint32_t* perp,
                                                 const uint8_t* pk, const uint8_t* sig, uint32_t n_msg, const uint8_t* msgs,
                                                 const uint32_t* msg_offsets, const uint32_t* msg_idx, uint8_t* status, uint8_t* expected_out,
                                                 uint8_t* session_status) {
  if (!ctx || !session_status) return -1;
  *session_status = DKGV_OK;
  if (n == 0) return dkgv_fail(ctx, "the session needs at least one generation (the reference indexes generations[0])");
  if ((t && !vv) || (m && (!perp || !pk || !sig || !status || !msg_offsets || n_msg == 0))) return dkgv_fail(ctx, "null pointer argument");
  for (uint32_t i = 0; i < m; i++)
    if (perp[i] >= n || (msg_idx && msg_idx[i] >= n_msg)) return dkgv_fail(ctx, "perpetrator or message index out of range");
  // ---- once per session: K_j = agg_coefficients(vv, 1..n) (dkg_math.rs:230-248), then the "expected key" of EVERY perpetrator
  // index: evaluate_polynomial over the K_j as if they were coefficients (quirk Q1, verification.rs:548-549)
  std::vector<uint32_t> ids(n);
  for (uint32_t j = 0; j < n; j++) ids[j] = j + 1;
  std::vector<uint8_t> keys((size_t)n * 48), expected((size_t)n * 48);
  uint8_t sst = DKGV_OK;
  if (int rc = dkgv_agg_final_keys(ctx, n, t, vv, ids.data(), n, nullptr, keys.data(), &sst)) return rc;
  if (sst == DKGV_OK)
    if (int rc = dkgv_eval_points(ctx, n, keys.data(), ids.data(), n, expected.data(), &sst)) return rc;
  *session_status = sst;
  if (expected_out && sst == DKGV_OK) memcpy(expected_out, expected.data(), expected.size());
  if (m == 0) return 0;
  // ---- the items: all messages hashed in one launch, all decodes + pairing checks in one batch
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  std::vector<uint8_t> hm((size_t)n_msg * 96);
  if (int rc = dkgv_hash_to_g2(ctx, n_msg, msgs, msg_offsets, hm.data())) return rc;
  CK(ctx->in_a.reserve((size_t)m * 48));
  CK(ctx->in_b.reserve((size_t)m * 96));
  CK(ctx->in_c.reserve((size_t)n_msg * 96));
  CK(ctx->out_a.reserve((size_t)m * 4));
  CK(ctx->out_b.reserve((size_t)m * 2));
  const size_t exp_bytes = ((size_t)n * 48 + 15) & ~(size_t)15;
  CK(ctx->scratch_d.reserve(exp_bytes + (size_t)m * 4));
  uint8_t* d_expected = (uint8_t*)ctx->scratch_d.p;
  uint32_t* d_perp = (uint32_t*)(d_expected + exp_bytes);
  CK(cudaMemcpyAsync(ctx->in_a.p, pk, (size_t)m * 48, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, sig, (size_t)m * 96, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_c.p, hm.data(), (size_t)n_msg * 96, cudaMemcpyHostToDevice, s));
  if (msg_idx) CK(cudaMemcpyAsync(ctx->out_a.p, msg_idx, (size_t)m * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_expected, expected.data(), (size_t)n * 48, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_perp, perp, (size_t)m * 4, cudaMemcpyHostToDevice, s));
  uint8_t* d_pair = (uint8_t*)ctx->out_b.p;
  uint8_t* d_status = d_pair + m;
  if (int rc = dkgv_bls_verify_batch_dev(ctx, m, (const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, n_msg, (const uint8_t*)ctx->in_c.p,
                                         msg_idx ? (const uint32_t*)ctx->out_a.p : nullptr, d_pair, s))
    return rc;
  const uint8_t* pk_st = (const uint8_t*)ctx->bls_st.p;  // decode statuses left by the pairing batch: keys, then signatures
  k_bpk_status<<<(m + 127) / 128, 128, 0, s>>>(d_pair, pk_st, pk_st + m, d_perp, (const uint8_t*)ctx->in_a.p, d_expected, sst, d_status, m);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(status, d_status, m, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}
