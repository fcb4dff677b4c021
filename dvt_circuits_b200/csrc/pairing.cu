// BLS partial-signature checks on the GPU (C ABI part 2): G2 decoding, hash-to-G2 and the batched
// pairing-equality kernel.  Replaces crates/dkg/src/crypto/bls_common.rs:11-40 and the signature
// loop of verify_generation_hashes (crates/dkg/src/verification.rs:237-248).
#include <cstdlib>
#include <cstring>

#include "ctx.hpp"
#include "h2c.cuh"
#include "pairing_vm.cuh"

using namespace dkgv;

// one thread per hashed message: decode (subgroup-checked) into an affine struct in global memory
__global__ void __launch_bounds__(32) k_g2_decode(const uint8_t* __restrict__ in, G2Aff* __restrict__ out, uint8_t* __restrict__ st,
                                                  uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  G2Aff a;
  uint32_t s = g2_decompress(in + (size_t)i * 96, &a, true);
  out[i] = a;
  st[i] = (uint8_t)s;
}

// one thread per hashed message: its 68 Miller-loop lines, shared by every check against that message
__global__ void __launch_bounds__(32) k_g2_prepare(const G2Aff* __restrict__ hm, const uint8_t* __restrict__ st, G2Line* __restrict__ lines,
                                                   uint32_t n_hm) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hm) return;
  G2Aff a = hm[i];
  if (st[i] != G1_DEC_OK || a.inf) return;  // never read: the check is skipped / the pair is the Gt identity
  g2_prepare(lines + (size_t)i * G2_PREP_LINES, &a);
}

__global__ void __launch_bounds__(32) k_g2_decompress_check(const uint8_t* __restrict__ in, uint8_t* __restrict__ st, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  G2Aff a;
  st[i] = (uint8_t)g2_decompress(in + (size_t)i * 96, &a, true);
}

// one thread per message; msgs are concatenated, offsets[i]..offsets[i+1]
__global__ void __launch_bounds__(32) k_hash_to_g2(const uint8_t* __restrict__ msgs, const uint32_t* __restrict__ offsets,
                                                   uint8_t* __restrict__ out, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  G2Aff h;
  hash_to_g2(&h, msgs + offsets[i], offsets[i + 1] - offsets[i]);
  uint8_t enc[96];
  g2_compress(&h, enc);
  for (int k = 0; k < 96; k++) out[(size_t)i * 96 + k] = enc[k];
}

// G1 decode (subgroup-checked) into an affine struct in global memory, one thread per key
__global__ void __launch_bounds__(64) k_g1_decode(const uint8_t* __restrict__ in, G1Aff* __restrict__ out, uint8_t* __restrict__ st,
                                                  uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  G1Aff a;
  uint32_t s = g1_decompress(in + (size_t)i * 48, &a, true);
  out[i] = a;
  st[i] = (uint8_t)s;
}

// one thread per check: status = OK | SLASHABLE_SIG_INVALID (pairing equality false)
//                               | PANIC_BAD_G1 (pk undecodable) | PANIC_BAD_G2 (signature undecodable)
// (the caller maps the decode failures to the reference's exit for its call site:
//  .expect -> panic in verify_generation_hashes, slashable in prove_wrong_final_key_generation).
// Keys and signatures arrive decoded (k_g1_decode / k_g2_decode): the decoders' code (square roots, subgroup
// checks, ~200 KB of SASS) stays out of this kernel, whose working set of instructions is the Miller loop and
// the final exponentiation only - with everything in one kernel ncu showed a 79 % instruction-cache hit rate and
// 22 % of the issue stalls on instruction fetch (profiles/r1_bls_verify.md).
#ifndef DKGV_BLS_MINB
#define DKGV_BLS_MINB 1
#endif
__global__ void __launch_bounds__(32, DKGV_BLS_MINB)
k_bls_verify(const G1Aff* __restrict__ pk, const uint8_t* __restrict__ pk_st, const G2Aff* __restrict__ sig,
             const uint8_t* __restrict__ sig_st, const G2Aff* __restrict__ hm, const G2Line* __restrict__ hm_lines,
             const uint32_t* __restrict__ hm_idx, uint8_t* __restrict__ status, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  // decode order of the reference: signature first, then key (verification.rs:238-241)
  uint8_t st;
  if (sig_st[i] != G1_DEC_OK) {
    st = DKGV_PANIC_BAD_G2;
  } else if (pk_st[i] != G1_DEC_OK) {
    st = DKGV_PANIC_BAD_G1;
  } else {
    G2Aff s = sig[i];
    G1Aff p = pk[i];
    uint32_t hi = hm_idx ? hm_idx[i] : 0;
    G2Aff h = hm[hi];
    st = bls_verify_prepared(&p, &s, &h, hm_lines ? hm_lines + (size_t)hi * G2_PREP_LINES : nullptr) ? DKGV_OK
                                                                                                       : DKGV_SLASHABLE_SIG_INVALID;
  }
  status[i] = st;
}

// The pairing VM (pairing_vm.cuh): PVM_R warps x 32 checks per block, every Fp2 of a check in shared memory.  status as
// k_bls_verify.  scratch: two Fp12 per check that the final exponentiation parks in global memory (f and t2 / t3 of
// tower.cuh final_exponentiation), chunk-planar: U4 index ((g * 6 + k) * 6 + chunk) * m_pad + check - coalesced.
constexpr size_t PVM_SMEM = (size_t)PVM_SLOTS * 6 * PVM_LANES * sizeof(U4);
__global__ void __launch_bounds__(PVM_LANES * PVM_R, 2)
k_pairing_vm(const G1Aff* __restrict__ pk, const uint8_t* __restrict__ pk_st, const G2Aff* __restrict__ sig, const uint8_t* __restrict__ sig_st,
             const G2Aff* __restrict__ hm, const G2Line* __restrict__ hm_lines, const uint32_t* __restrict__ hm_idx, uint8_t* __restrict__ status,
             uint32_t m, U4* __restrict__ scratch, uint32_t m_pad) {
  extern __shared__ U4 pvm_file[];
  const uint32_t lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * PVM_LANES + lane, ii = i < m ? i : m - 1;
  const uint32_t hi = hm_idx ? hm_idx[ii] : 0;
  const PvmCtx c{pvm_file + lane, (const uint32_t*)(hm_lines + (size_t)hi * G2_PREP_LINES), (const uint32_t*)&pk[ii], (const uint32_t*)&sig[ii], nullptr};
  if (role == 0) pvm_init_point(c);
  __syncthreads();
#pragma unroll 1
  for (uint32_t ci = 0; ci < PVM_N_CALLS; ci++) {
    const PvmCall k = pvm_call(ci);
    if (k.kind == 0) {
      uint32_t pc = pvm_seg_start[k.a][role];
#pragma unroll 1
      for (;;) {
        pc = pvm_exec(c, pc, k.b);
        if (pc & PVM_END_FLAG) break;
        __syncthreads();
      }
    } else {
#pragma unroll 1
      for (uint32_t s = role; s < 6; s += PVM_R) {
#pragma unroll
        for (uint32_t ch = 0; ch < 6; ch++) {
          if (k.kind == 1) {
            c.file[(size_t)((pvm_reg_slot(k.a) + s) * 6 + ch) * PVM_LANES] = c.file[(size_t)((pvm_reg_slot(k.b) + s) * 6 + ch) * PVM_LANES];
          } else if (k.kind == 2) {
            scratch[(size_t)((k.a * 6 + s) * 6 + ch) * m_pad + i] = c.file[(size_t)((pvm_reg_slot(k.b) + s) * 6 + ch) * PVM_LANES];
          } else {
            c.file[(size_t)((pvm_reg_slot(k.a) + s) * 6 + ch) * PVM_LANES] = scratch[(size_t)((k.b * 6 + s) * 6 + ch) * m_pad + i];
          }
        }
      }
      __syncthreads();
    }
  }
  if (role == 0 && i < m) {
    uint8_t st;
    if (sig_st[i] != G1_DEC_OK) {  // decode order of the reference: signature first, then key (verification.rs:238-241)
      st = DKGV_PANIC_BAD_G2;
    } else if (pk_st[i] != G1_DEC_OK) {
      st = DKGV_PANIC_BAD_G1;
    } else {
      st = pvm_status(pk[i].inf != 0, sig[i].inf != 0, hm[hi].inf != 0, pvm_result_is_one(c));
    }
    status[i] = st;
  }
}

// The Miller-loop lines of the hashed messages by the same VM (segments P_D / P_A): PVM_R warps per 32 messages instead of one
// thread per message - with ONE message (a finalization) the one-thread kernel was a 68-step serial chain of ~2 ms.
__global__ void __launch_bounds__(PVM_LANES * PVM_R, 2)
k_g2_prepare_vm(const G2Aff* __restrict__ hm, const uint8_t* __restrict__ st, G2Line* __restrict__ lines, uint32_t n_hm) {
  extern __shared__ U4 pvm_file[];
  const uint32_t lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * PVM_LANES + lane, ii = i < n_hm ? i : n_hm - 1;
  // lanes beyond n_hm run message n_hm - 1 again and store the same values to the same places; a message that did not decode or is
  // the identity gets lines nobody reads (the check is decided without them)
  (void)st;
  const PvmCtx c{pvm_file + lane, nullptr, nullptr, (const uint32_t*)&hm[ii], (uint32_t*)(lines + (size_t)ii * G2_PREP_LINES)};
  if (role == 0) pvm_init_point(c);
  __syncthreads();
#pragma unroll 1
  for (uint32_t ci = 0; ci < PVM_N_PREP_CALLS; ci++) {
    const PvmCall k = pvm_prep_call(ci);
    uint32_t pc = pvm_seg_start[k.a][role];
#pragma unroll 1
    for (;;) {
      pc = pvm_exec(c, pc, k.b);
      if (pc & PVM_END_FLAG) break;
      __syncthreads();
    }
  }
}

// signer-side helper for synthetic ceremonies (not a verification step): out[i] = [scalars[i]] * base
struct LimbArr {
  const uint32_t* w;
  DKGV_HD uint32_t operator()(int i) const { return w[i]; }
};
__global__ void __launch_bounds__(32) k_g2_mul_batch(const uint8_t* __restrict__ base96, const uint8_t* __restrict__ scalars,
                                                     uint8_t* __restrict__ out, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  G2Aff b;
  g2_decompress(base96, &b, false);
  uint32_t s[8];
  fr_raw_from_be32(s, scalars + (size_t)i * 32);
  G2Proj p = g2_from_affine(b), r;
  g2_mul_public(&r, &p, LimbArr{s}, 8);
  G2Aff a;
  g2_to_affine(&a, &r);
  uint8_t enc[96];
  g2_compress(&a, enc);
  for (int k = 0; k < 96; k++) out[(size_t)i * 96 + k] = enc[k];
}

// base_hash_d = SHA-256(gen_id(16) || n || k || len as u8 || vv[d][0..t))  (verification.rs:151-175);
// one thread per dealer, 48*t + 19 byte preimage streamed from HBM
__global__ void __launch_bounds__(64) k_initial_commitment_hashes(const uint8_t* __restrict__ vv, uint32_t n_d, uint32_t t,
                                                                  const uint8_t* __restrict__ hdr19, uint8_t* __restrict__ out) {
  uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_d) return;
  Sha256 s;
  sha_init(&s);
  uint8_t h[19];
  for (int i = 0; i < 19; i++) h[i] = hdr19[i];
  sha_update(&s, h, 19);
  const uint8_t* row = vv + (size_t)d * t * 48;
  uint8_t chunk[48];
#pragma unroll 1
  for (uint32_t k = 0; k < t; k++) {
    for (int i = 0; i < 48; i++) chunk[i] = row[(size_t)k * 48 + i];
    sha_update(&s, chunk, 48);
  }
  uint8_t dg[32];
  sha_finish(&s, dg);
  for (int i = 0; i < 32; i++) out[(size_t)d * 32 + i] = dg[i];
}

static int ensure_stack(dkgv_ctx* ctx) {
  if (!ctx->stack_set) {
    CK(cudaDeviceSetLimit(cudaLimitStackSize, 16 * 1024));
    ctx->stack_set = true;
  }
  return 0;
}

extern "C" int dkgv_g2_decompress_check(dkgv_ctx* ctx, uint32_t m, const uint8_t* in, uint8_t* decode_status) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!in || !decode_status) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  if (int rc = ensure_stack(ctx)) return rc;
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve((size_t)m * 96));
  CK(ctx->out_b.reserve(m));
  CK(cudaMemcpyAsync(ctx->in_a.p, in, (size_t)m * 96, cudaMemcpyHostToDevice, s));
  k_g2_decompress_check<<<(m + 31) / 32, 32, 0, s>>>((const uint8_t*)ctx->in_a.p, (uint8_t*)ctx->out_b.p, m);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(decode_status, ctx->out_b.p, m, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int dkgv_hash_to_g2(dkgv_ctx* ctx, uint32_t m, const uint8_t* msgs, const uint32_t* offsets, uint8_t* out) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!offsets || !out) return dkgv_fail(ctx, "null pointer argument");
  size_t total = offsets[m];
  if (total && !msgs) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  if (int rc = ensure_stack(ctx)) return rc;
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve(total ? total : 1));
  CK(ctx->in_b.reserve((size_t)(m + 1) * 4));
  CK(ctx->out_a.reserve((size_t)m * 96));
  if (total) CK(cudaMemcpyAsync(ctx->in_a.p, msgs, total, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, offsets, (size_t)(m + 1) * 4, cudaMemcpyHostToDevice, s));
  k_hash_to_g2<<<(m + 31) / 32, 32, 0, s>>>((const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, (uint8_t*)ctx->out_a.p, m);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, (size_t)m * 96, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int dkgv_bls_verify_batch_dev(dkgv_ctx* ctx, uint32_t m, const uint8_t* d_pk, const uint8_t* d_sig, uint32_t n_hm,
                                         const uint8_t* d_hm, const uint32_t* d_hm_idx, uint8_t* d_status, void* stream) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!d_pk || !d_sig || !d_hm || !d_status || n_hm == 0) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  if (int rc = ensure_stack(ctx)) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  CK(ctx->scratch_a.reserve((size_t)n_hm * sizeof(G2Aff)));
  CK(ctx->scratch_b.reserve(n_hm));
  CK(ctx->bls_pk.reserve((size_t)m * sizeof(G1Aff)));
  CK(ctx->bls_sig.reserve((size_t)m * sizeof(G2Aff)));
  CK(ctx->bls_st.reserve((size_t)m * 2));
  uint8_t* pk_st = (uint8_t*)ctx->bls_st.p;
  uint8_t* sig_st = pk_st + m;
  const bool want_vm = ctx->bls_path != DKGV_BLS_PATH_THREAD;
  const bool prep = ((size_t)n_hm * 4 <= m || want_vm) && (size_t)n_hm * G2_PREP_LINES * sizeof(G2Line) <= ((size_t)1 << 30);
  if (prep) CK(ctx->scratch_c.reserve((size_t)n_hm * G2_PREP_LINES * sizeof(G2Line)));
  if (!ctx->pvm_attr_set) {
    CK(cudaFuncSetAttribute(k_pairing_vm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PVM_SMEM));
    CK(cudaFuncSetAttribute(k_pairing_vm, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_g2_prepare_vm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PVM_SMEM));
    ctx->pvm_attr_set = true;
  }
  // Three independent chains before the pairing kernel - hashed messages: decode -> lines; signatures: decode; keys: decode - each a
  // long serial computation per item.  They run on three streams of the ctx and join before the pairing kernel: for a small batch
  // (a finalization: 1 024 checks) their latencies overlap instead of adding up.
  cudaStream_t sa = ctx->fd_streams[0], sb = ctx->fd_streams[1], sc = ctx->fd_streams[2];
  CK(cudaEventRecord(ctx->fd_fork, s));
  CK(cudaStreamWaitEvent(sa, ctx->fd_fork, 0));
  CK(cudaStreamWaitEvent(sb, ctx->fd_fork, 0));
  CK(cudaStreamWaitEvent(sc, ctx->fd_fork, 0));
  k_g2_decode<<<(n_hm + 31) / 32, 32, 0, sa>>>(d_hm, (G2Aff*)ctx->scratch_a.p, (uint8_t*)ctx->scratch_b.p, n_hm);
  ctx->launches++;
  // messages shared by several checks: their lines are computed once (19.6 KB per message)
  const G2Line* lines = nullptr;
  if (prep) {
    if (want_vm)
      k_g2_prepare_vm<<<(n_hm + PVM_LANES - 1) / PVM_LANES, PVM_LANES * PVM_R, PVM_SMEM, sa>>>((const G2Aff*)ctx->scratch_a.p, (const uint8_t*)ctx->scratch_b.p,
                                                                                             (G2Line*)ctx->scratch_c.p, n_hm);
    else
      k_g2_prepare<<<(n_hm + 31) / 32, 32, 0, sa>>>((const G2Aff*)ctx->scratch_a.p, (const uint8_t*)ctx->scratch_b.p, (G2Line*)ctx->scratch_c.p, n_hm);
    ctx->launches++;
    lines = (const G2Line*)ctx->scratch_c.p;
  }
  // decode keys and signatures in their own kernels
  k_g2_decode<<<(m + 31) / 32, 32, 0, sb>>>(d_sig, (G2Aff*)ctx->bls_sig.p, sig_st, m);
  k_g1_decode<<<(m + 63) / 64, 64, 0, sc>>>(d_pk, (G1Aff*)ctx->bls_pk.p, pk_st, m);
  ctx->launches += 2;
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->fd_join[0], sa));
  CK(cudaEventRecord(ctx->fd_join[1], sb));
  CK(cudaEventRecord(ctx->fd_join[2], sc));
  for (int i = 0; i < 3; i++) CK(cudaStreamWaitEvent(s, ctx->fd_join[i], 0));
  if (lines && ctx->bls_path != DKGV_BLS_PATH_THREAD) {
    // the VM needs the prepared lines of every hashed message (19.6 KB each); batches with more distinct messages than
    // that budget allows take the one-thread-per-check kernel below
    const uint32_t blocks = (m + PVM_LANES - 1) / PVM_LANES, m_pad = blocks * PVM_LANES;
    CK(ctx->bls_scratch.reserve((size_t)2 * 6 * 6 * m_pad * sizeof(U4)));
    CK(cudaEventRecord(ctx->ev_bls0, s));
    k_pairing_vm<<<blocks, PVM_LANES * PVM_R, PVM_SMEM, s>>>((const G1Aff*)ctx->bls_pk.p, pk_st, (const G2Aff*)ctx->bls_sig.p, sig_st,
                                                            (const G2Aff*)ctx->scratch_a.p, lines, d_hm_idx, d_status, m, (U4*)ctx->bls_scratch.p, m_pad);
    CK(cudaEventRecord(ctx->ev_bls1, s));
    ctx->bls_recorded = true;
    ctx->last_bls_path = DKGV_BLS_PATH_VM;
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
  }
  ctx->last_bls_path = DKGV_BLS_PATH_THREAD;
  CK(cudaEventRecord(ctx->ev_bls0, s));
  k_bls_verify<<<(m + 31) / 32, 32, 0, s>>>((const G1Aff*)ctx->bls_pk.p, pk_st, (const G2Aff*)ctx->bls_sig.p, sig_st,
                                            (const G2Aff*)ctx->scratch_a.p, lines, d_hm_idx, d_status, m);
  CK(cudaEventRecord(ctx->ev_bls1, s));
  ctx->bls_recorded = true;
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int dkgv_bls_verify_batch(dkgv_ctx* ctx, uint32_t m, const uint8_t* pk, const uint8_t* sig, uint32_t n_hm,
                                     const uint8_t* hm, const uint32_t* hm_idx, uint8_t* status) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!pk || !sig || !hm || !status || n_hm == 0) return dkgv_fail(ctx, "null pointer argument");
  if (hm_idx)
    for (uint32_t i = 0; i < m; i++)
      if (hm_idx[i] >= n_hm) return dkgv_fail(ctx, "hm_idx out of range");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve((size_t)m * 48));
  CK(ctx->in_b.reserve((size_t)m * 96));
  CK(ctx->in_c.reserve((size_t)n_hm * 96));
  CK(ctx->out_a.reserve(hm_idx ? (size_t)m * 4 : 4));
  CK(ctx->out_b.reserve(m));
  CK(cudaMemcpyAsync(ctx->in_a.p, pk, (size_t)m * 48, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, sig, (size_t)m * 96, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_c.p, hm, (size_t)n_hm * 96, cudaMemcpyHostToDevice, s));
  if (hm_idx) CK(cudaMemcpyAsync(ctx->out_a.p, hm_idx, (size_t)m * 4, cudaMemcpyHostToDevice, s));
  int rc = dkgv_bls_verify_batch_dev(ctx, m, (const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, n_hm, (const uint8_t*)ctx->in_c.p,
                                     hm_idx ? (const uint32_t*)ctx->out_a.p : nullptr, (uint8_t*)ctx->out_b.p, s);
  if (rc) return rc;
  CK(cudaMemcpyAsync(status, ctx->out_b.p, m, cudaMemcpyDeviceToHost, s));
  // an undecodable hashed message is a caller bug, not an item outcome
  uint8_t hst[16];
  uint32_t chk = n_hm < 16 ? n_hm : 16;
  CK(cudaMemcpyAsync(hst, ctx->scratch_b.p, chk, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (uint32_t i = 0; i < chk; i++)
    if (hst[i]) return dkgv_fail(ctx, "hashed message is not a valid G2 encoding");
  return 0;
}

extern "C" int dkgv_g2_mul_batch(dkgv_ctx* ctx, uint32_t m, const uint8_t* base96, const uint8_t* scalars, uint8_t* out) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!base96 || !scalars || !out) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  if (int rc = ensure_stack(ctx)) return rc;
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve(96));
  CK(ctx->in_b.reserve((size_t)m * 32));
  CK(ctx->out_a.reserve((size_t)m * 96));
  CK(cudaMemcpyAsync(ctx->in_a.p, base96, 96, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, scalars, (size_t)m * 32, cudaMemcpyHostToDevice, s));
  k_g2_mul_batch<<<(m + 31) / 32, 32, 0, s>>>((const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_b.p, (uint8_t*)ctx->out_a.p, m);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, (size_t)m * 96, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int dkgv_initial_commitment_hashes(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t t, const uint8_t* vv, const uint8_t* gen_id16,
                                              uint8_t n, uint8_t k, uint8_t* out) {
  if (!ctx) return -1;
  if (n_dealers == 0) return 0;
  if ((t && !vv) || !gen_id16 || !out) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  if (int rc = ensure_stack(ctx)) return rc;
  cudaStream_t s = ctx->stream;
  size_t vvb = (size_t)n_dealers * t * 48;
  uint8_t hdr[19];
  memcpy(hdr, gen_id16, 16);
  hdr[16] = n;
  hdr[17] = k;
  hdr[18] = (uint8_t)t;  // `len as u8` truncation of the reference
  CK(ctx->in_a.reserve(vvb ? vvb : 1));
  CK(ctx->in_b.reserve(32));
  CK(ctx->out_a.reserve((size_t)n_dealers * 32));
  if (vvb) CK(cudaMemcpyAsync(ctx->in_a.p, vv, vvb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, hdr, 19, cudaMemcpyHostToDevice, s));
  k_initial_commitment_hashes<<<(n_dealers + 63) / 64, 64, 0, s>>>((const uint8_t*)ctx->in_a.p, n_dealers, t, (const uint8_t*)ctx->in_b.p,
                                                                 (uint8_t*)ctx->out_a.p);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, (size_t)n_dealers * 32, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int dkgv_set_bls_path(dkgv_ctx* ctx, int mode) {
  if (!ctx) return -1;
  if (mode < DKGV_BLS_PATH_AUTO || mode > DKGV_BLS_PATH_THREAD) return dkgv_fail(ctx, "unknown pairing path");
  ctx->bls_path = mode;
  return 0;
}
extern "C" int dkgv_last_bls_path(const dkgv_ctx* ctx) { return ctx ? ctx->last_bls_path : -1; }
extern "C" int dkgv_last_bls_kernel_ms(dkgv_ctx* ctx, float* ms) {
  if (!ctx || !ms) return -1;
  if (!ctx->bls_recorded) return dkgv_fail(ctx, "no pairing batch launched yet");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->ev_bls1));
  CK(cudaEventElapsedTime(ms, ctx->ev_bls0, ctx->ev_bls1));
  return 0;
}
