// Multi-GPU inside the library (C ABI part 5): a dkgv_comm per ctx (one process per GPU, NCCL over NVLink / NVSwitch) and the
// sharded forms of the three paths SURVEY 8(e) shards:
//   share matrix      dealer row blocks, no exchange during compute; ONE all-gather of (verdict bitmask + the two job flags) per call
//   pairing checks    items sharded; one all-gather of the status bytes
//   agg_coefficients  each rank decodes and sums its dealers' rows (dkg_math.rs:234-241), one all-gather of the t projective partial
//                     sums (144 B each - point addition is not an NCCL reduce op), every rank adds the partials and evaluates the keys
// NCCL is bound at run time (dlopen of libnccl.so.2 - inside a torch process that is the copy torch already loaded), so the library
// itself links against nothing but the CUDA runtime and loads on machines without NCCL; a host in any language gets multi-GPU
// without bringing its own collectives: rank 0 asks for a unique id, hands its 128 bytes to the other ranks by whatever
// means it has (the bench uses torch.distributed's store), every rank calls dkgv_comm_init.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.hpp"
#include "g1.cuh"

using namespace dkgv;

namespace {
struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, ncclUniqueIdBytes, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
Nccl* nccl() {
  static Nccl n;
  if (n.h || !n.err.empty()) return &n;
  const char* names[] = {getenv("DKGV_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.h) break;
  }
  if (!n.h) {
    n.err = std::string("cannot load NCCL (libnccl.so.2; set DKGV_NCCL_LIB): ") + dlerror();
    return &n;
  }
  n.GetUniqueId = (int (*)(void*))dlsym(n.h, "ncclGetUniqueId");
  n.CommInitRank = (int (*)(void**, int, ncclUniqueIdBytes, int))dlsym(n.h, "ncclCommInitRank");
  n.CommDestroy = (int (*)(void*))dlsym(n.h, "ncclCommDestroy");
  n.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(n.h, "ncclAllGather");
  n.GetErrorString = (const char* (*)(int))dlsym(n.h, "ncclGetErrorString");
  if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllGather || !n.GetErrorString) {
    n.err = "libnccl lacks an expected symbol";
    n.h = nullptr;
  }
  return &n;
}
constexpr int NCCL_UINT8 = 1;  // ncclUint8 (stable across NCCL 2.x)
}  // namespace

#define NK(call)                                                                        \
  do {                                                                                  \
    int r_ = (call);                                                                    \
    if (r_ != 0) {                                                                      \
      ctx->err = std::string(#call) + ": " + nccl()->GetErrorString(r_);               \
      return -4;                                                                        \
    }                                                                                   \
  } while (0)

extern "C" int dkgv_comm_unique_id(uint8_t id_out[128]) {
  Nccl* n = nccl();
  if (!n->h || !id_out) return -4;
  ncclUniqueIdBytes id;
  if (n->GetUniqueId(&id) != 0) return -4;
  memcpy(id_out, id.internal, 128);
  return 0;
}

extern "C" int dkgv_comm_init(dkgv_ctx* ctx, const uint8_t id[128], int rank, int world) {
  if (!ctx) return -1;
  if (!id || world < 1 || rank < 0 || rank >= world) return dkgv_fail(ctx, "bad communicator arguments");
  if (ctx->comm) return dkgv_fail(ctx, "this ctx already has a communicator");
  Nccl* n = nccl();
  if (!n->h) return dkgv_fail(ctx, n->err.c_str());
  CK(cudaSetDevice(ctx->device));
  ncclUniqueIdBytes uid;
  memcpy(uid.internal, id, 128);
  NK(n->CommInitRank(&ctx->comm, world, uid, rank));
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return 0;
}

extern "C" int dkgv_comm_destroy(dkgv_ctx* ctx) {
  if (!ctx) return -1;
  if (ctx->comm) {
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    nccl()->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_world = 1;
  ctx->comm_rank = 0;
  return 0;
}
extern "C" int dkgv_comm_world(const dkgv_ctx* ctx) { return ctx ? ctx->comm_world : -1; }
extern "C" int dkgv_comm_rank(const dkgv_ctx* ctx) { return ctx ? ctx->comm_rank : -1; }

// all-gather of `bytes` per rank; world 1 (no communicator): a device copy
static int all_gather(dkgv_ctx* ctx, const void* d_send, void* d_recv, size_t bytes, cudaStream_t s) {
  if (!ctx->comm) {
    if (d_send != d_recv) CK(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  NK(nccl()->AllGather(d_send, d_recv, bytes, NCCL_UINT8, ctx->comm, s));
  ctx->collectives++;
  return 0;
}

// all-gather of `bytes` per rank over the ctx's communicator (rank r's block lands at d_recv + r * bytes on every rank; world 1: a
// copy); asynchronous on `stream`.  For callers that shard a batch entry point themselves (e.g. the bad-partial-key items).
extern "C" int dkgv_all_gather_dev(dkgv_ctx* ctx, const void* d_send, void* d_recv, size_t bytes, void* stream) {
  if (!ctx) return -1;
  if (!d_send || !d_recv || bytes == 0) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  return all_gather(ctx, d_send, (uint8_t*)d_recv, bytes, stream ? (cudaStream_t)stream : ctx->stream);
}

// bits[w] of the local verdicts + the two job flags appended: the payload one rank contributes to the gather
__global__ void __launch_bounds__(256)
k_pack_verdicts_flags(const uint8_t* __restrict__ status, uint32_t* __restrict__ out, size_t n, size_t words, const uint32_t* __restrict__ flags) {
  size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < words) {
    uint32_t v = 0;
#pragma unroll 4
    for (int b = 0; b < 32; b++) {
      size_t i = w * 32 + b;
      if (i < n && status[i] != DKGV_OK) v |= 1u << b;
    }
    out[w] = v;
  }
  if (w < 2) out[words + w] = flags[w];
}

// ---- share matrix, dealer row blocks -----------------------------------------------------------------------------------
// Every rank calls this with ITS dealers' rows (n_local dealers; all ranks the same n_local) and the full id list.
// d_gather [world][chunk] u32 with chunk = dkgv_share_gather_words(n_local, n_r): rank r's chunk holds the verdict bitmask of its
// rows (bit i % 32 of word i / 32 = share i of the row block is NOT ok) followed by its two job flags.  The honest path is the
// asynchronous submit + pack + ONE all-gather; the call then synchronises once and - only if some rank reported unsettled dealers or
// foreign ids - runs that rank's evaluation and a second gather (every rank sees every flag, so all of them agree on it).
extern "C" uint32_t dkgv_share_gather_words(uint32_t n_local, uint32_t n_r) {
  size_t words = ((size_t)n_local * n_r + 31) / 32 + 2;
  return (uint32_t)((words + 3) & ~(size_t)3);
}
int dkgv_share_submit_internal(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint32_t* d_ids,
                               const uint8_t* d_shares, uint8_t* d_status, uint32_t* d_flags, cudaStream_t s);  // dkgv.cu
int dkgv_share_finish_internal(dkgv_ctx* ctx, const uint32_t* h_flags, cudaStream_t s);

extern "C" int dkgv_share_matrix_verify_sharded_dev(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_r, uint32_t t, const uint8_t* d_vv_local,
                                                    const uint32_t* d_ids, const uint8_t* d_shares_local, uint8_t* d_status_local,
                                                    uint32_t* d_gather, void* stream) {
  if (!ctx) return -1;
  if (n_local == 0 || n_r == 0) return dkgv_fail(ctx, "empty row block");
  if (!d_ids || !d_shares_local || !d_status_local || !d_gather || (t && !d_vv_local)) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  const int world = ctx->comm_world, rank = ctx->comm_rank;
  const size_t n = (size_t)n_local * n_r, words = (n + 31) / 32;
  const uint32_t chunk = dkgv_share_gather_words(n_local, n_r);
  CK(ctx->comm_flags.reserve((size_t)world * 8));
  if (!ctx->h_comm_flags || ctx->h_comm_flags_cap < (size_t)world * 2) {
    if (ctx->h_comm_flags) cudaFreeHost(ctx->h_comm_flags);
    CK(cudaMallocHost(&ctx->h_comm_flags, (size_t)world * 8));
    ctx->h_comm_flags_cap = (size_t)world * 2;
  }
  if (int rc = dkgv_share_submit_internal(ctx, n_local, n_r, t, d_vv_local, d_ids, d_shares_local, d_status_local, nullptr, s)) return rc;
  uint32_t* mine = d_gather + (size_t)rank * chunk;
  for (int round = 0; round < 2; round++) {
    k_pack_verdicts_flags<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(d_status_local, mine, n, words, ctx->job.d_flags);
    ctx->launches++;
    CK(cudaGetLastError());
    if (int rc = all_gather(ctx, mine, d_gather, (size_t)chunk * 4, s)) return rc;
    if (round == 1) break;
    // every rank's two flag words, in one strided copy
    CK(cudaMemcpy2DAsync(ctx->h_comm_flags, 8, d_gather + words, (size_t)chunk * 4, 8, world, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    bool any = false;
    for (int r = 0; r < world; r++) any |= ctx->h_comm_flags[2 * r] != 0 || ctx->h_comm_flags[2 * r + 1] != 0;
    if (int rc = dkgv_share_finish_internal(ctx, ctx->h_comm_flags + 2 * rank, s)) return rc;
    if (!any) break;  // the honest ceremony: one gather, one synchronisation
    CK(cudaMemsetAsync(ctx->job.d_flags, 0, 8, s));  // the second payload reports this rank as settled
  }
  return 0;
}

// Pipelined form of the same call for a host that verifies many ceremonies: `enqueue` queues the shortcut, the pack, the ONE all-gather
// and the copy of every rank's two flag words to h_flags (pinned host memory of the caller, 2 * world words) and returns WITHOUT
// synchronising - any number of ceremonies can be in flight on a stream (the ctx's scratch is reused in stream order), and the call is
// CUDA-graph capturable.  After the caller has synchronised (stream, event) it calls `settle` with the same arguments: flags all zero
// (the honest ceremony, every rank settled by the shortcut) - nothing to do, the verdicts are final; otherwise the ceremony is run again
// through the synchronous entry point (repair route / evaluation, second gather).  Every rank sees every flag, so all ranks take the
// same branch.  *reran (optional) tells which.
extern "C" int dkgv_share_matrix_enqueue_sharded_dev(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_r, uint32_t t, const uint8_t* d_vv_local,
                                                     const uint32_t* d_ids, const uint8_t* d_shares_local, uint8_t* d_status_local,
                                                     uint32_t* d_gather, uint32_t* h_flags, void* stream) {
  if (!ctx) return -1;
  if (n_local == 0 || n_r == 0) return dkgv_fail(ctx, "empty row block");
  if (!d_ids || !d_shares_local || !d_status_local || !d_gather || !h_flags || (t && !d_vv_local)) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  const int world = ctx->comm_world, rank = ctx->comm_rank;
  const size_t n = (size_t)n_local * n_r, words = (n + 31) / 32;
  const uint32_t chunk = dkgv_share_gather_words(n_local, n_r);
  if (int rc = dkgv_share_submit_internal(ctx, n_local, n_r, t, d_vv_local, d_ids, d_shares_local, d_status_local, nullptr, s)) return rc;
  uint32_t* mine = d_gather + (size_t)rank * chunk;
  k_pack_verdicts_flags<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(d_status_local, mine, n, words, ctx->job.d_flags);
  ctx->launches++;
  CK(cudaGetLastError());
  if (int rc = all_gather(ctx, mine, d_gather, (size_t)chunk * 4, s)) return rc;
  CK(cudaMemcpy2DAsync(h_flags, 8, d_gather + words, (size_t)chunk * 4, 8, world, cudaMemcpyDeviceToHost, s));
  ctx->job.open = false;  // nothing is pending on the ctx: `settle` starts over from the caller's arguments if it has to
  return 0;
}

extern "C" int dkgv_share_matrix_settle_sharded_dev(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_r, uint32_t t, const uint8_t* d_vv_local,
                                                    const uint32_t* d_ids, const uint8_t* d_shares_local, uint8_t* d_status_local,
                                                    uint32_t* d_gather, const uint32_t* h_flags, void* stream, int* reran) {
  if (!ctx) return -1;
  if (!h_flags) return dkgv_fail(ctx, "null pointer argument");
  if (reran) *reran = 0;
  bool any = false;
  for (int r = 0; r < ctx->comm_world; r++) any |= h_flags[2 * r] != 0 || h_flags[2 * r + 1] != 0;
  if (!any) return 0;
  if (reran) *reran = 1;
  return dkgv_share_matrix_verify_sharded_dev(ctx, n_local, n_r, t, d_vv_local, d_ids, d_shares_local, d_status_local, d_gather, stream);
}

// The same pair with HOST buffers (pinned memory for the copies to be asynchronous): the rows, ids and shares go to the ctx's staging
// buffers, the status bytes of the own rows, the gathered chunks (gather may be NULL) and every rank's flag words come back - all queued
// on the ctx's own stream, nothing synchronised.  A host keeps one ctx per ceremony in flight on a GPU (the ctxs of a device share the
// fixed-base table): the copies of one ceremony then run under the kernels of another.  dkgv_sync, then settle.
int dkgv_take_vv_wait(dkgv_ctx* ctx, cudaStream_t s);  // dkgv.cu
static int stage_rows(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_r, uint32_t t, const uint8_t* vv, const uint32_t* ids, const uint8_t* shares,
                      bool vv_on_second_stream, cudaStream_t s) {
  const size_t vvb = (size_t)n_local * t * 48, idb = (size_t)n_r * 4, shb = (size_t)n_local * n_r * 32;
  CK(ctx->in_a.reserve(vvb ? vvb : 1));
  CK(ctx->in_b.reserve(idb));
  CK(ctx->in_c.reserve(shb));
  CK(ctx->out_a.reserve((size_t)n_local * n_r));
  CK(ctx->share_gather.reserve((size_t)ctx->comm_world * dkgv_share_gather_words(n_local, n_r) * 4));
  CK(cudaMemcpyAsync(ctx->in_b.p, ids, idb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_c.p, shares, shb, cudaMemcpyHostToDevice, s));
  if (!vvb) return 0;
  if (!vv_on_second_stream) {
    CK(cudaMemcpyAsync(ctx->in_a.p, vv, vvb, cudaMemcpyHostToDevice, s));
    return 0;
  }
  // as dkgv_share_matrix_verify: the verification vectors follow on a second stream, under the difference tables; the first kernel
  // that reads them waits for ev_vv.  ev_sh orders the copy behind everything already queued on s (the previous ceremony of this ctx).
  cudaStream_t s2 = ctx->fd_streams[0];
  CK(cudaEventRecord(ctx->ev_sh, s));
  CK(cudaStreamWaitEvent(s2, ctx->ev_sh, 0));
  CK(cudaMemcpyAsync(ctx->in_a.p, vv, vvb, cudaMemcpyHostToDevice, s2));
  CK(cudaEventRecord(ctx->ev_vv, s2));
  ctx->vv_wait = ctx->ev_vv;
  return 0;
}
static int unstage_results(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_r, uint8_t* status, uint32_t* gather, cudaStream_t s) {
  CK(cudaMemcpyAsync(status, ctx->out_a.p, (size_t)n_local * n_r, cudaMemcpyDeviceToHost, s));
  if (gather)
    CK(cudaMemcpyAsync(gather, ctx->share_gather.p, (size_t)ctx->comm_world * dkgv_share_gather_words(n_local, n_r) * 4, cudaMemcpyDeviceToHost, s));
  return 0;
}

extern "C" int dkgv_share_matrix_enqueue_sharded(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_r, uint32_t t, const uint8_t* vv_local,
                                                 const uint32_t* ids, const uint8_t* shares_local, uint8_t* status_local, uint32_t* gather,
                                                 uint32_t* h_flags) {
  if (!ctx) return -1;
  if (n_local == 0 || n_r == 0) return dkgv_fail(ctx, "empty row block");
  if (!ids || !shares_local || !status_local || !h_flags || (t && !vv_local)) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (int rc = stage_rows(ctx, n_local, n_r, t, vv_local, ids, shares_local, true, s)) return rc;
  const size_t n = (size_t)n_local * n_r, words = (n + 31) / 32;
  const uint32_t chunk = dkgv_share_gather_words(n_local, n_r);
  if (int rc = dkgv_share_submit_internal(ctx, n_local, n_r, t, (const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p,
                                          (const uint8_t*)ctx->in_c.p, (uint8_t*)ctx->out_a.p, nullptr, s))
    return rc;
  if (int rc = dkgv_take_vv_wait(ctx, s)) return rc;  // a path that never read them: the copy still ends before the next ceremony's
  uint32_t* all = (uint32_t*)ctx->share_gather.p;
  uint32_t* mine = all + (size_t)ctx->comm_rank * chunk;
  k_pack_verdicts_flags<<<(unsigned)((words + 255) / 256), 256, 0, s>>>((const uint8_t*)ctx->out_a.p, mine, n, words, ctx->job.d_flags);
  ctx->launches++;
  CK(cudaGetLastError());
  if (int rc = all_gather(ctx, mine, all, (size_t)chunk * 4, s)) return rc;
  if (int rc = unstage_results(ctx, n_local, n_r, status_local, gather, s)) return rc;
  CK(cudaMemcpy2DAsync(h_flags, 8, all + words, (size_t)chunk * 4, 8, ctx->comm_world, cudaMemcpyDeviceToHost, s));
  ctx->job.open = false;
  return 0;
}

extern "C" int dkgv_share_matrix_settle_sharded(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_r, uint32_t t, const uint8_t* vv_local,
                                                const uint32_t* ids, const uint8_t* shares_local, uint8_t* status_local, uint32_t* gather,
                                                const uint32_t* h_flags, int* reran) {
  if (!ctx) return -1;
  if (!h_flags) return dkgv_fail(ctx, "null pointer argument");
  if (reran) *reran = 0;
  bool any = false;
  for (int r = 0; r < ctx->comm_world; r++) any |= h_flags[2 * r] != 0 || h_flags[2 * r + 1] != 0;
  if (!any) return 0;
  if (reran) *reran = 1;
  if (n_local == 0 || n_r == 0) return dkgv_fail(ctx, "empty row block");
  if (!ids || !shares_local || !status_local || (t && !vv_local)) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (int rc = stage_rows(ctx, n_local, n_r, t, vv_local, ids, shares_local, false, s)) return rc;
  if (int rc = dkgv_share_matrix_verify_sharded_dev(ctx, n_local, n_r, t, (const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p,
                                                    (const uint8_t*)ctx->in_c.p, (uint8_t*)ctx->out_a.p, (uint32_t*)ctx->share_gather.p, s))
    return rc;
  if (int rc = unstage_results(ctx, n_local, n_r, status_local, gather, s)) return rc;
  CK(cudaStreamSynchronize(s));
  return 0;
}

// ---- pairing checks, items sharded --------------------------------------------------------------------------------------
// every rank: its m_local (pk, sig) pairs -> d_status_all [world][m_local]
extern "C" int dkgv_bls_verify_batch_sharded_dev(dkgv_ctx* ctx, uint32_t m_local, const uint8_t* d_pk, const uint8_t* d_sig, uint32_t n_hm,
                                                 const uint8_t* d_hm, const uint32_t* d_hm_idx, uint8_t* d_status_all, void* stream) {
  if (!ctx) return -1;
  if (m_local == 0) return dkgv_fail(ctx, "empty item block");
  if (!d_status_all) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  uint8_t* mine = d_status_all + (size_t)ctx->comm_rank * m_local;
  if (int rc = dkgv_bls_verify_batch_dev(ctx, m_local, d_pk, d_sig, n_hm, d_hm, d_hm_idx, mine, s)) return rc;
  return all_gather(ctx, mine, d_status_all, m_local, s);
}

// ---- agg_coefficients, dealers sharded ------------------------------------------------------------------------------------
// sum of `world` projective partial sums per coefficient: partial[r][k][36] -> affine 25-word records + 48-byte encodings
__global__ void __launch_bounds__(64)
k_add_partials(const uint32_t* __restrict__ partial, uint32_t world, uint32_t t, uint32_t* __restrict__ coeffs25, uint8_t* __restrict__ enc_out) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t) return;
  G1Proj acc = g1_identity();
  for (uint32_t r = 0; r < world; r++) {
    const uint32_t* o = partial + ((size_t)r * t + k) * 36;
    G1Proj p;
#pragma unroll
    for (int l = 0; l < 12; l++) {
      p.x.l[l] = o[l];
      p.y.l[l] = o[12 + l];
      p.z.l[l] = o[24 + l];
    }
    acc = g1_add(acc, p);
  }
  G1Aff a = g1_to_affine(acc);
  uint32_t* c = coeffs25 + (size_t)k * 25;
#pragma unroll
  for (int l = 0; l < 12; l++) {
    c[l] = a.x.l[l];
    c[12 + l] = a.y.l[l];
  }
  c[24] = a.inf;
  uint8_t enc[48];
  g1_compress(a, enc);
  for (int i = 0; i < 48; i++) enc_out[(size_t)k * 48 + i] = enc[i];
}

int dkgv_column_partials_internal(dkgv_ctx* ctx, uint32_t n_local, uint32_t t, const uint8_t* h_vv, uint32_t* d_partial36, cudaStream_t s);  // final.cu
int dkgv_keys_from_coeffs_internal(dkgv_ctx* ctx, uint32_t t, const uint32_t* ids, uint32_t n_ids, uint8_t* keys_out, cudaStream_t s);   // final.cu

// vv_local [n_local][t][48]: this rank's generations (any split of the n generations over the ranks; the sum is commutative);
// coeff_out [t][48], keys_out [n_ids][48] on every rank; *status OK / PANIC_BAD_G1 (some rank met an undecodable commitment)
extern "C" int dkgv_agg_final_keys_sharded(dkgv_ctx* ctx, uint32_t n_local, uint32_t t, const uint8_t* vv_local, const uint32_t* ids,
                                           uint32_t n_ids, uint8_t* coeff_out, uint8_t* keys_out, uint8_t* status) {
  if (!ctx || !status) return -1;
  *status = DKGV_OK;
  if (t == 0 || n_local == 0) return dkgv_fail(ctx, "empty row block");
  if (!vv_local || (n_ids && (!ids || !keys_out))) return dkgv_fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int world = ctx->comm_world, rank = ctx->comm_rank;
  const size_t rec = (size_t)t * 36 * 4 + 16;  // t projective points + a 16-byte tail carrying this rank's "undecodable" flag
  CK(ctx->comm_buf.reserve(rec * world));
  uint8_t* base = (uint8_t*)ctx->comm_buf.p;
  uint32_t* mine = (uint32_t*)(base + rec * rank);
  if (int rc = dkgv_column_partials_internal(ctx, n_local, t, vv_local, mine, s)) return rc;
  if (int rc = all_gather(ctx, mine, base, rec, s)) return rc;
  // the gathered records are not contiguous as points (tails in between): add them record by record
  CK(ctx->scratch_a.reserve((size_t)t * 25 * 4));
  CK(ctx->out_a.reserve((size_t)t * 48));
  CK(ctx->scratch_d.reserve((size_t)world * t * 36 * 4));
  for (int r = 0; r < world; r++)
    CK(cudaMemcpyAsync((uint8_t*)ctx->scratch_d.p + (size_t)r * t * 144, base + rec * r, (size_t)t * 144, cudaMemcpyDeviceToDevice, s));
  k_add_partials<<<(t + 63) / 64, 64, 0, s>>>((const uint32_t*)ctx->scratch_d.p, (uint32_t)world, t, (uint32_t*)ctx->scratch_a.p, (uint8_t*)ctx->out_a.p);
  ctx->launches++;
  CK(cudaGetLastError());
  std::vector<uint32_t> tails((size_t)world * 4);
  CK(cudaMemcpy2DAsync(tails.data(), 16, base + (size_t)t * 144, rec, 16, world, cudaMemcpyDeviceToHost, s));
  if (coeff_out) CK(cudaMemcpyAsync(coeff_out, ctx->out_a.p, (size_t)t * 48, cudaMemcpyDeviceToHost, s));
  if (n_ids)
    if (int rc = dkgv_keys_from_coeffs_internal(ctx, t, ids, n_ids, keys_out, s)) return rc;
  CK(cudaStreamSynchronize(s));
  for (int r = 0; r < world; r++)
    if (tails[(size_t)r * 4]) *status = DKGV_PANIC_BAD_G1;
  return 0;
}
