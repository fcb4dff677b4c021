// CUDA kernels (sm_100a) and the C ABI of include/dkgv.h.
// Hot path: Feldman share verification for a (dealer x recipient) matrix -
//   crates/dkg/src/verification.rs:129-146 + crates/dkg/src/dkg_math.rs:160-174.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <algorithm>
#include <string>
#include <vector>

#include "fdiff.cuh"

using namespace dkgv;

// share_fd.cu
int dkgv_fd_setup(dkgv_ctx* ctx);
bool dkgv_fd_ids_consecutive(const uint32_t* h_ids, uint32_t n_r);
bool dkgv_fd_shortcut_applies(const dkgv_ctx* ctx, uint32_t n_r, uint32_t t);
int dkgv_fd_submit(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint32_t* d_ids, const uint8_t* d_shares,
                   uint8_t* d_status, bool shortcut, uint32_t* d_flags, cudaStream_t s);
const uint8_t* dkgv_fd_need_groups(const dkgv_ctx* ctx, uint32_t n_d);
int dkgv_fd_compact_unsettled(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t count, const uint8_t* d_shares, uint8_t* d_status, cudaStream_t s,
                              const uint32_t** list, const uint8_t** shares_c, uint8_t** status_c);
int dkgv_fd_scatter_status(dkgv_ctx* ctx, uint32_t n_r, uint32_t count, uint8_t* d_status, cudaStream_t s);
bool dkgv_fd_repair_applies(const dkgv_ctx* ctx, uint32_t n_r, uint32_t t);
int dkgv_fd_repair(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint8_t* d_shares, uint8_t* d_status,
                   uint32_t* d_flags, cudaStream_t s);
int dkgv_share_matrix_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint8_t* d_shares, uint8_t* d_status, const uint8_t* filter, cudaStream_t s);
int dkgv_feldman_eval_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint32_t* h_ids, uint8_t* d_out48, cudaStream_t s);

// ============================================================================ kernels
// Fixed-base table of the generator (layout and the per-thread routines in feldman.cuh): the W window bases first, then one thread
// per run of GTAB_RUN entries.
__global__ void __launch_bounds__(32) k_gtab_bases(uint32_t bits, uint32_t windows, uint32_t* __restrict__ base) {
  if (threadIdx.x < windows) gtab_base(bits, threadIdx.x, base);
}
__global__ void __launch_bounds__(128) k_gtab_fill(uint32_t bits, uint32_t windows, const uint32_t* __restrict__ base, uint32_t* __restrict__ gtab) {
  const uint32_t run = blockIdx.x * blockDim.x + threadIdx.x, per_window = (1u << (bits - 1)) / GTAB_RUN;
  if (run >= windows * per_window) return;
  gtab_fill_run(bits, base, run / per_window, (run % per_window) * GTAB_RUN, gtab);
}
// debug: entries idx = first, first + stride, ... against the slow definition (gtab_entry); *bad counts the differences
__global__ void __launch_bounds__(128) k_gtab_check(GTab g, uint32_t first, uint32_t stride, uint32_t count, uint32_t* __restrict__ bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t idx = first + i * stride;
  if (idx >= gtab_entries(g.bits)) return;
  G1Aff a = gtab_entry(g.bits, idx);
  const uint32_t* e = g.p + (size_t)idx * 24;
  bool same = true;
#pragma unroll
  for (int k = 0; k < 12; k++) same = same && e[k] == a.x.l[k] && e[12 + k] == a.y.l[k];
  if (!same) atomicAdd(bad, 1u);
}

// Decode + subgroup-check the verification vectors into the limb-planar layout of feldman.cuh.
//   vv [n_d][t][48]  ->  limbs/inf;  dealer_bad[d] |= 1 when any coefficient fails to decode
//   (the reference panics on `.expect("Invalid pubkey")`, verification.rs:132-137).
__global__ void __launch_bounds__(128)
k_decompress_vv(const uint8_t* __restrict__ vv, uint32_t n_d, uint32_t t, uint32_t n_pad, uint32_t* __restrict__ limbs,
                uint8_t* __restrict__ inf, uint8_t* __restrict__ dealer_bad, uint8_t* __restrict__ point_status,
                const uint8_t* __restrict__ group_filter, const uint32_t* __restrict__ map) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)n_pad * t) return;
  uint32_t d = (uint32_t)(p % n_pad), k = (uint32_t)(p / n_pad);
  // a warp = one 32-dealer group and one coefficient: only the groups the evaluation will read (nullptr: all of them)
  if (group_filter && !group_filter[d / 32]) return;
  G1Aff a;
  a.x = zero<FpParams>();
  a.y = zero<FpParams>();
  a.inf = 1;
  if (d < n_d) {
    const uint32_t src = map ? map[d] : d;  // dense session of listed dealers (per-dealer fallback): column d holds dealer map[d]
    uint32_t st = g1_decompress(vv + ((size_t)src * t + k) * 48, &a, true);
    if (st != G1_DEC_OK) dealer_bad[d] = 1;
    if (point_status) point_status[(size_t)d * t + k] = (uint8_t)st;
  }
  vv_store(limbs, inf, n_pad, k, d, a);
}

// The hot kernel.  Thread = one share (dealer d, recipient column j); the 32 lanes of a warp hold
// 32 consecutive dealers and ONE recipient id, so the double-and-add over the id bits is
// warp-uniform (no divergence) and every coefficient load is a fully coalesced 128 B line per limb.
// Field operands live in the shared-memory operand file of vm.cuh (13 slots x 48 B per thread).
constexpr int SV_WARPS = 4;   // k_feldman_eval (inlined formulas, cold path)
#ifndef DKGV_SVM_NT
#define DKGV_SVM_NT 32  // measured on B200: 32 -> 306k, 64 -> 286k, 128 -> 276k shares/s (n_r=1024,t=683,n_d=256)
#endif
constexpr int SVM_NT = DKGV_SVM_NT;  // threads per block of the hot kernel: one recipient id per warp
constexpr size_t SVM_SMEM = (size_t)VM_SLOTS * 3 * SVM_NT * sizeof(U4);
typedef VVView HotView;
__global__ void __launch_bounds__(SVM_NT)
k_share_verify(HotView vv, const uint8_t* __restrict__ dealer_bad, const uint32_t* __restrict__ ids,
               const uint8_t* __restrict__ shares, GTab gtab, uint8_t* __restrict__ status,
               uint32_t n_d, uint32_t n_r, uint32_t t) {
  extern __shared__ U4 opfile[];
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t d = blockIdx.x * 32 + lane;
  uint32_t j = blockIdx.y * (SVM_NT / 32) + warp;
  if (j >= n_r) return;
  // blocks are dispatched in increasing blockIdx; recipient ids usually ascend (1..n) and the chain
  // for a larger id is longer, so walk the columns from the back: expensive units first, cheap ones
  // fill the tail of the last wave (longest-processing-time-first)
  j = n_r - 1 - j;
  bool active = d < n_d;
  uint32_t dd = active ? d : n_d - 1;
  OpFile f{opfile + threadIdx.x, SVM_NT};
  uint8_t st = vm_share_check(f, vv, t, dd, ids[j], shares + ((size_t)dd * n_r + j) * 32, gtab, dealer_bad[dd] != 0);
  if (active) status[(size_t)d * n_r + j] = st;
}

// Sparse item list over one session: a warp = up to 32 items that share ONE recipient (the host sorts
// the items by recipient), so the chain over the id bits stays warp-uniform; dealers differ per lane.
__global__ void __launch_bounds__(32)
k_share_items(VVView vv, const uint8_t* __restrict__ dealer_bad, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ s_dealer,
              const uint32_t* __restrict__ s_col, const uint32_t* __restrict__ s_orig, const uint32_t* __restrict__ warp_first,
              const uint8_t* __restrict__ secrets, GTab gtab, uint8_t* __restrict__ status, uint32_t t) {
  extern __shared__ U4 opfile[];
  uint32_t first = warp_first[blockIdx.x], cnt = warp_first[blockIdx.x + 1] - first;
  bool active = threadIdx.x < cnt;
  uint32_t it = first + (active ? threadIdx.x : cnt - 1);
  uint32_t d = s_dealer[it], id = ids[s_col[first]], orig = s_orig[it];
  OpFile f{opfile + threadIdx.x, 32};
  uint8_t st = vm_share_check(f, vv, t, d, id, secrets + (size_t)orig * 32, gtab, dealer_bad[d] != 0);
  if (active) status[orig] = st;
}

// bit i of bits = (status[i] != DKGV_OK), little-endian within each 32-bit word; n_words = ceil(n / 32)
__global__ void __launch_bounds__(256) k_pack_verdicts(const uint8_t* __restrict__ status, uint32_t* __restrict__ bits, size_t n) {
  size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 32 >= n) return;
  uint32_t v = 0;
#pragma unroll 4
  for (int b = 0; b < 32; b++) {
    size_t i = w * 32 + b;
    if (i < n && status[i] != DKGV_OK) v |= 1u << b;
  }
  bits[w] = v;
}

// evaluate_polynomial for every (dealer, id) with compressed output
__global__ void __launch_bounds__(SV_WARPS * 32)
k_feldman_eval(VVView vv, const uint32_t* __restrict__ ids, uint8_t* __restrict__ out, uint32_t n_d, uint32_t n_r, uint32_t t) {
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t d = blockIdx.x * 32 + lane;
  uint32_t j = blockIdx.y * SV_WARPS + warp;
  if (j >= n_r) return;
  bool active = d < n_d;
  uint32_t dd = active ? d : n_d - 1;
  G1Proj ev = feldman_eval(vv, t, dd, ids[j]);
  G1Aff a = g1_to_affine(ev);
  uint8_t enc[48];
  g1_compress(a, enc);
  if (active) {
    uint8_t* o = out + ((size_t)d * n_r + j) * 48;
#pragma unroll
    for (int i = 0; i < 48; i++) o[i] = enc[i];
  }
}

__global__ void __launch_bounds__(128)
k_fixed_base_mul(const uint8_t* __restrict__ scalars, GTab gtab, uint8_t* __restrict__ out,
                 uint8_t* __restrict__ status, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  uint32_t s[8];
  bool ok = fr_raw_from_be32(s, scalars + (size_t)i * 32);
  G1Aff a = g1_to_affine(fixed_base_mul(gtab, s));
  uint8_t enc[48];
  g1_compress(a, enc);
#pragma unroll
  for (int k = 0; k < 48; k++) out[(size_t)i * 48 + k] = ok ? enc[k] : 0;
  status[i] = ok ? DKGV_OK : DKGV_SLASHABLE_SECRET_RANGE;
}

__global__ void __launch_bounds__(128)
k_g1_decompress_check(const uint8_t* __restrict__ in, uint8_t* __restrict__ st, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  G1Aff a;
  st[i] = (uint8_t)g1_decompress(in + (size_t)i * 48, &a, true);
}

// Dealer-side helper used to build synthetic ceremonies (not part of verification):
// out[d][j] = sum_k coeffs[d][k] * ids[j]^k mod r, 32-byte big-endian in and out.
__global__ void __launch_bounds__(128)
k_fr_poly_eval(const uint8_t* __restrict__ coeffs, const uint32_t* __restrict__ ids, uint8_t* __restrict__ out, uint32_t n_d,
               uint32_t n_r, uint32_t t) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t d = blockIdx.y;
  if (j >= n_r || d >= n_d) return;
  Fr x = zero<FrParams>();
  x.l[0] = ids[j];
  x = to_mont(x);
  Fr acc = zero<FrParams>();
#pragma unroll 1
  for (int k = (int)t - 1; k >= 0; k--) {
    Fr c;
    fr_raw_from_be32(c.l, coeffs + ((size_t)d * t + k) * 32);
    acc = add(mul(acc, x), to_mont(c));
  }
  acc = from_mont(acc);
  uint8_t* o = out + ((size_t)d * n_r + j) * 32;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* q = o + 28 - 4 * i;
    q[0] = (uint8_t)(acc.l[i] >> 24);
    q[1] = (uint8_t)(acc.l[i] >> 16);
    q[2] = (uint8_t)(acc.l[i] >> 8);
    q[3] = (uint8_t)acc.l[i];
  }
}

// ============================================================================ host side
#include "ctx.hpp"
using dkgv_host::DevBuf;
namespace {
thread_local std::string g_create_error;
// The fixed-base table is read-only once built: every ctx of a process on the same device with the same window width shares ONE copy
// (reference-counted), so a host that keeps several ctxs per GPU - one per ceremony in flight, see dkgv_share_matrix_enqueue_sharded_dev -
// pays for the table (2.4 GB at 22 bits, 32 GB at 26) once.
struct SharedTab {
  int device;
  uint32_t bits;
  uint32_t* mem;
  int refs;
};
std::mutex g_tab_mu;
std::vector<SharedTab> g_tabs;
thread_local bool g_tab_locked = false;  // this thread is inside dkgv_ctx_create_ex and holds g_tab_mu (its failure paths call dkgv_ctx_destroy)
struct TabLock {
  std::unique_lock<std::mutex> l;
  TabLock() : l(g_tab_mu) { g_tab_locked = true; }
  ~TabLock() { g_tab_locked = false; }
};
}

static int fail(dkgv_ctx* ctx, const char* msg) { return dkgv_fail(ctx, msg); }

extern "C" int dkgv_ctx_create(int device, dkgv_ctx** out) { return dkgv_ctx_create_ex(device, 0, out); }

extern "C" int dkgv_ctx_create_ex(int device, uint32_t gtab_bits, dkgv_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return -2;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return -1;
  }
  dkgv_ctx* ctx = new (std::nothrow) dkgv_ctx();
  if (!ctx) return -3;
  ctx->device = device;
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    dkgv_ctx_destroy(ctx);
    return -2;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaEventCreate(&ctx->ev_hot0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_hot1)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev_dec0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_dec1)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev_bls0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_bls1)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_vv, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&ctx->ev_sh, cudaEventDisableTiming)) != cudaSuccess)
    return bail("cudaEventCreate", e);
  // fixed-base table: the window width asked for (0 = DKGV_GTAB_BITS from the environment, else the default), stepping down while the
  // device cannot hold it
  if (gtab_bits == 0) {
    const char* env = getenv("DKGV_GTAB_BITS");
    gtab_bits = env ? (uint32_t)atoi(env) : GTAB_BITS_DEFAULT;
  }
  if (gtab_bits < GTAB_BITS_MIN || gtab_bits > GTAB_BITS_MAX) {
    g_create_error = "gtab_bits out of range";
    dkgv_ctx_destroy(ctx);
    return -1;
  }
  TabLock tab_lock;  // held until the table is built: a second ctx created meanwhile waits and shares it
  bool tab_shared = false;
  for (;; gtab_bits -= 2) {
    for (SharedTab& st : g_tabs)
      if (st.device == device && st.bits == gtab_bits) {
        st.refs++;
        ctx->gtab_mem = st.mem;
        tab_shared = true;
        break;
      }
    if (tab_shared) break;
    e = cudaMalloc(&ctx->gtab_mem, gtab_words(gtab_bits) * 4);
    if (e == cudaSuccess) break;
    ctx->gtab_mem = nullptr;
    cudaGetLastError();
    if (gtab_bits < 16 + 2) return bail("cudaMalloc gtab", e);
  }
  if (!tab_shared) g_tabs.push_back(SharedTab{device, gtab_bits, ctx->gtab_mem, 1});
  ctx->gtab = GTab{ctx->gtab_mem, gtab_bits, gtab_windows(gtab_bits)};
  if ((e = cudaFuncSetAttribute(k_share_verify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SVM_SMEM)) != cudaSuccess)
    return bail("cudaFuncSetAttribute smem", e);
  if ((e = cudaFuncSetAttribute(k_share_verify, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess)
    return bail("cudaFuncSetAttribute carveout", e);
  if ((e = cudaFuncSetAttribute(k_share_items, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess)
    return bail("cudaFuncSetAttribute carveout", e);
  if ((e = cudaMalloc(&ctx->job_flags, 16)) != cudaSuccess) return bail("cudaMalloc flags", e);
  if ((e = cudaMallocHost(&ctx->h_job_flags, 16)) != cudaSuccess) return bail("cudaMallocHost flags", e);
  if (dkgv_fd_setup(ctx) != 0) {
    g_create_error = "finite-difference path setup: " + ctx->err;
    dkgv_ctx_destroy(ctx);
    return -2;
  }
  if (!tab_shared) {
    uint32_t* base = nullptr;
    if ((e = cudaMalloc(&base, 32 * 48 * 4)) != cudaSuccess) return bail("cudaMalloc gtab bases", e);
    const uint32_t runs = ctx->gtab.windows * ((1u << (gtab_bits - 1)) / GTAB_RUN);
    k_gtab_bases<<<1, 32, 0, ctx->stream>>>(gtab_bits, ctx->gtab.windows, base);
    k_gtab_fill<<<(runs + 127) / 128, 128, 0, ctx->stream>>>(gtab_bits, ctx->gtab.windows, base, ctx->gtab_mem);
    ctx->launches += 2;
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(base);
    if (e != cudaSuccess) return bail("k_gtab_fill", e);
  }
  *out = ctx;
  return 0;
}

extern "C" void dkgv_ctx_destroy(dkgv_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (cudaEvent_t ev : ctx->ev_sc)
    if (ev) cudaEventDestroy(ev);
  dkgv_comm_destroy(ctx);
  if (ctx->h_comm_flags) cudaFreeHost(ctx->h_comm_flags);
  if (ctx->job_flags) cudaFree(ctx->job_flags);
  if (ctx->h_job_flags) cudaFreeHost(ctx->h_job_flags);
  for (cudaEvent_t ev : ctx->ev_fd)
    if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < 16; i++) {
    if (ctx->fd_streams[i]) cudaStreamSynchronize(ctx->fd_streams[i]), cudaStreamDestroy(ctx->fd_streams[i]);
    if (ctx->fd_join[i]) cudaEventDestroy(ctx->fd_join[i]);
  }
  if (ctx->fd_fork) cudaEventDestroy(ctx->fd_fork);
  if (ctx->gtab_mem) {  // the last ctx that uses the shared table frees it
    std::unique_lock<std::mutex> tab_lock(g_tab_mu, std::defer_lock);
    if (!g_tab_locked) tab_lock.lock();
    for (size_t i = 0; i < g_tabs.size(); i++)
      if (g_tabs[i].mem == ctx->gtab_mem) {
        if (--g_tabs[i].refs == 0) {
          cudaFree(ctx->gtab_mem);
          g_tabs.erase(g_tabs.begin() + (long)i);
        }
        break;
      }
  }
  if (ctx->ev_hot0) cudaEventDestroy(ctx->ev_hot0);
  if (ctx->ev_hot1) cudaEventDestroy(ctx->ev_hot1);
  if (ctx->ev_dec0) cudaEventDestroy(ctx->ev_dec0);
  if (ctx->ev_dec1) cudaEventDestroy(ctx->ev_dec1);
  if (ctx->ev_bls0) cudaEventDestroy(ctx->ev_bls0);
  if (ctx->ev_bls1) cudaEventDestroy(ctx->ev_bls1);
  if (ctx->ev_vv) cudaEventDestroy(ctx->ev_vv);
  if (ctx->ev_sh) cudaEventDestroy(ctx->ev_sh);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;  // every DevBuf member frees its allocation (ctx.hpp)
}

extern "C" const char* dkgv_last_error(const dkgv_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" uint64_t dkgv_launch_count(const dkgv_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint32_t dkgv_gtab_bits(const dkgv_ctx* ctx) { return ctx ? ctx->gtab.bits : 0; }
extern "C" int dkgv_gtab_selfcheck(dkgv_ctx* ctx, uint32_t first, uint32_t stride, uint32_t count, uint32_t* n_bad) {
  if (!ctx || !n_bad) return -1;
  CK(cudaSetDevice(ctx->device));
  uint32_t* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 4);
  if (e != cudaSuccess) return fail(ctx, "cudaMalloc");
  cudaMemsetAsync(d, 0, 4, ctx->stream);
  if (count) k_gtab_check<<<(count + 127) / 128, 128, 0, ctx->stream>>>(ctx->gtab, first, stride, count, d);
  ctx->launches++;
  e = cudaMemcpyAsync(n_bad, d, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  return e == cudaSuccess ? 0 : fail(ctx, cudaGetErrorString(e));
}
extern "C" int dkgv_last_hot_kernel_ms(dkgv_ctx* ctx, float* ms) {
  if (!ctx || !ms) return -1;
  if (!ctx->hot_recorded) return fail(ctx, "no hot kernel launched yet");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->ev_hot1));
  CK(cudaEventElapsedTime(ms, ctx->ev_hot0, ctx->ev_hot1));
  return 0;
}
extern "C" int dkgv_last_decode_ms(dkgv_ctx* ctx, float* ms, int* subgroup_checked) {
  if (!ctx || !ms) return -1;
  if (!ctx->dec_recorded) return fail(ctx, "no verification-vector decode launched yet");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->ev_dec1));
  CK(cudaEventElapsedTime(ms, ctx->ev_dec0, ctx->ev_dec1));
  if (subgroup_checked) *subgroup_checked = 1;
  return 0;
}
extern "C" int dkgv_sync(dkgv_ctx* ctx) {
  if (!ctx) return -1;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// decode + subgroup-check vv into the ctx session buffers (asynchronous on s)
static int session_decode(dkgv_ctx* ctx, uint32_t n_d, uint32_t t, const uint8_t* d_vv, cudaStream_t s, VVView* view, uint32_t* n_pad_out,
                          const uint8_t* group_filter = nullptr, const uint32_t* map = nullptr) {
  uint32_t n_pad = (n_d + 31) & ~31u;
  uint32_t tt = t ? t : 1;
  CK(ctx->vv_limbs.reserve((size_t)tt * 24 * n_pad * 4));
  CK(ctx->vv_inf.reserve((size_t)tt * n_pad));
  CK(ctx->dealer_bad.reserve(n_pad));
  CK(cudaMemsetAsync(ctx->dealer_bad.p, 0, n_pad, s));
  if (t) {
    size_t total = (size_t)n_pad * t;
    if (ctx->ev_dec0) CK(cudaEventRecord(ctx->ev_dec0, s));
    k_decompress_vv<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(d_vv, n_d, t, n_pad, (uint32_t*)ctx->vv_limbs.p, (uint8_t*)ctx->vv_inf.p,
                                                                  (uint8_t*)ctx->dealer_bad.p, nullptr, group_filter, map);
    if (ctx->ev_dec1) CK(cudaEventRecord(ctx->ev_dec1, s));
    ctx->dec_recorded = true;
    ctx->launches++;
    CK(cudaGetLastError());
  }
  ctx->vv_decoded = true;
  view->limbs = (const uint32_t*)ctx->vv_limbs.p;
  view->inf = (const uint8_t*)ctx->vv_inf.p;
  view->n_pad = n_pad;
  *n_pad_out = n_pad;
  return 0;
}

extern "C" int dkgv_set_share_path(dkgv_ctx* ctx, int mode) {
  if (!ctx) return -1;
  if (mode < DKGV_SHARE_PATH_AUTO || mode > DKGV_SHARE_PATH_FDIFF) return fail(ctx, "unknown share path");
  ctx->share_path = mode;
  return 0;
}
extern "C" int dkgv_set_share_parts(dkgv_ctx* ctx, uint32_t parts) {
  if (!ctx) return -1;
  if (parts > FD_MAX_PARTS) return fail(ctx, "too many parts");
  ctx->share_parts = parts;
  return 0;
}
extern "C" int dkgv_set_share_overlap(dkgv_ctx* ctx, int on) {
  if (!ctx) return -1;
  ctx->fd_overlap = on != 0;
  return 0;
}
extern "C" int dkgv_set_share_shortcut(dkgv_ctx* ctx, int on) {
  if (!ctx) return -1;
  ctx->fd_polycheck = on != 0;
  return 0;
}
extern "C" int dkgv_set_share_repair(dkgv_ctx* ctx, int on) {
  if (!ctx) return -1;
  ctx->fd_repair = on != 0;
  return 0;
}
extern "C" int dkgv_last_share_repaired(const dkgv_ctx* ctx) { return ctx ? (int)ctx->last_repaired : -1; }
extern "C" int dkgv_last_share_decoded(const dkgv_ctx* ctx) { return ctx ? (ctx->vv_decoded ? 1 : 0) : -1; }
extern "C" int dkgv_last_share_continued(const dkgv_ctx* ctx) { return ctx ? (ctx->fd_last_need ? 1 : 0) : -1;
}
extern "C" int dkgv_share_fd_plan(uint32_t t, uint32_t n_r, uint32_t parts_force, uint32_t n_opt, uint32_t* parts, uint32_t* h, int32_t* lo,
                                  int32_t* hi, uint32_t* steps, uint64_t* modmul_fd, uint64_t* modmul_horner) {
  FdPlan p = fd_make_plan(t, n_r, parts_force, n_opt);
  if (parts) *parts = p.m;
  if (h) *h = p.h;
  if (lo) *lo = p.lo;
  if (hi) *hi = p.hi;
  if (steps) *steps = p.steps;
  if (modmul_fd) *modmul_fd = p.cost_fd;
  if (modmul_horner) *modmul_horner = p.cost_horner;
  return p.cost_fd == ~0ull ? -1 : (p.use ? 1 : 0);
}
extern "C" int dkgv_last_share_path(const dkgv_ctx* ctx) { return ctx ? ctx->last_share_path : -1; }
extern "C" int dkgv_last_share_phases_ms(dkgv_ctx* ctx, float* ms4) {
  if (!ctx || !ms4) return -1;
  const bool eval = ctx->fd_last_need;  // the last call continued into the evaluation: its phases; else the shortcut's
  if (eval ? !ctx->fd_recorded : !ctx->sc_recorded) return fail(ctx, "no finite-difference share verification launched yet");
  CK(cudaSetDevice(ctx->device));
  cudaEvent_t* ev = eval ? ctx->ev_fd : ctx->ev_sc;
  CK(cudaEventSynchronize(ev[4]));
  for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(ms4 + i, ev[i], ev[i + 1]));
  return 0;
}

// Horner per share over the whole matrix (arbitrary ids, or shapes where finite differences do not pay); asynchronous
static int share_matrix_horner(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint32_t* d_ids,
                               const uint8_t* d_shares, uint8_t* d_status, cudaStream_t s) {
  VVView view;
  uint32_t n_pad;
  if (int rc = session_decode(ctx, n_d, t, d_vv, s, &view, &n_pad)) return rc;
  ctx->last_share_path = DKGV_SHARE_PATH_HORNER;
  dim3 grid(n_pad / 32, (n_r + SVM_NT / 32 - 1) / (SVM_NT / 32));
  CK(cudaEventRecord(ctx->ev_hot0, s));
  k_share_verify<<<grid, SVM_NT, SVM_SMEM, s>>>(view, (const uint8_t*)ctx->dealer_bad.p, d_ids, d_shares, ctx->gtab, d_status, n_d, n_r, t);
  CK(cudaEventRecord(ctx->ev_hot1, s));
  ctx->hot_recorded = true;
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

// Queue the default path without synchronising.  Recipient ids that are the consecutive ranks 1..n_r (always the case for a
// ceremony, verification.rs:50-66,129) allow the consistency shortcut and, behind it, t Horner evaluations + finite differences
// per dealer.  Whether the ids are such a permutation is decided ON THE DEVICE (flags[0]) while the shortcut already runs
// speculatively; whether any dealer group still needs the evaluation is flags[1].  share_finish acts on the two words.
// the stream's next kernel reads the verification vectors: wait for their copy if one is in flight (dkgv_share_matrix_verify)
int dkgv_take_vv_wait(dkgv_ctx* ctx, cudaStream_t s) {
  if (ctx->vv_wait) {
    cudaEvent_t ev = ctx->vv_wait;
    ctx->vv_wait = nullptr;
    CK(cudaStreamWaitEvent(s, ev, 0));
  }
  return 0;
}

static int share_submit(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint32_t* d_ids,
                        const uint8_t* d_shares, uint8_t* d_status, uint32_t* d_flags, cudaStream_t s) {
  dkgv_ctx::ShareJob& job = ctx->job;
  job = dkgv_ctx::ShareJob();
  job.n_d = n_d, job.n_r = n_r, job.t = t;
  job.d_vv = d_vv, job.d_ids = d_ids, job.d_shares = d_shares, job.d_status = d_status;
  job.d_flags = d_flags ? d_flags : ctx->job_flags;
  job.parts = ctx->share_parts;
  job.open = true;
  ctx->fd_last_need = false;
  ctx->last_repaired = 0;
  if (ctx->share_path != DKGV_SHARE_PATH_HORNER && t >= 2 && n_r >= 3 && n_r <= 65535 && t <= 65535 * FD_MAX_PARTS) {
    // planned for the full evaluation (n_opt = 0): measured within 0.5 % of the shortcut-optimal split on an honest ceremony
    // (553 vs 551 ms at 1024 / 683) and 2 % better when the groups have to continue (734 vs 750 ms)
    FdPlan plan = fd_make_plan(t, n_r, ctx->share_parts, 0);
    job.fd = plan.cost_fd != ~0ull && (plan.use || ctx->share_path == DKGV_SHARE_PATH_FDIFF);
  }
  if (!job.fd) {  // nothing to speculate on: the whole Horner route is queued now
    if (int rc = dkgv_take_vv_wait(ctx, s)) return rc;
    CK(cudaMemsetAsync(job.d_flags, 0, 8, s));
    return share_matrix_horner(ctx, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, s);
  }
  job.shortcut = dkgv_fd_shortcut_applies(ctx, n_r, t);
  ctx->last_share_path = DKGV_SHARE_PATH_FDIFF;
  ctx->vv_decoded = false;  // compress(G * p_k) == C_k needs no decompression; the decode waits until a group needs the evaluation
  return dkgv_fd_submit(ctx, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, job.shortcut, job.d_flags, s);
}

// h_flags: the two flag words of the job as read back by the caller (after synchronising the stream), or nullptr: read them
// here (one stream synchronisation).  Queues whatever the flags ask for: nothing (an honest ceremony), the evaluation of the
// dealer groups the shortcut could not settle, or the Horner route when the ids were no permutation of 1..n_r.
static int share_finish(dkgv_ctx* ctx, const uint32_t* h_flags, cudaStream_t s) {
  dkgv_ctx::ShareJob job = ctx->job;
  if (!job.open) return dkgv_fail(ctx, "no share-matrix job submitted");
  ctx->job.open = false;
  if (!job.fd) return 0;
  if (!h_flags) {
    CK(cudaMemcpyAsync(ctx->h_job_flags, job.d_flags, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    h_flags = ctx->h_job_flags;
  }
  if (h_flags[0]) return share_matrix_horner(ctx, job.n_d, job.n_r, job.t, job.d_vv, job.d_ids, job.d_shares, job.d_status, s);
  if (!h_flags[1]) return 0;  // every verdict is OK and already written
  uint32_t unsettled = h_flags[1];  // with the shortcut: the dealers that failed a condition
  if (job.shortcut && dkgv_fd_repair_applies(ctx, job.n_r, job.t)) {
    // some dealers' shares are not on one polynomial: decode them as Reed-Solomon words with errors before anything is evaluated
    // in the exponent; what the decoder settles (exactly - share_rs.cuh) needs no evaluation.  One more read-back of the flags.
    if (int rc = dkgv_fd_repair(ctx, job.n_d, job.n_r, job.t, job.d_vv, job.d_shares, job.d_status, job.d_flags, s)) return rc;
    CK(cudaMemcpyAsync(ctx->h_job_flags, job.d_flags, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ctx->last_repaired = ctx->h_job_flags[2];
    if (!ctx->h_job_flags[1]) return 0;
    unsettled = ctx->h_job_flags[1];
  }
  ctx->fd_last_need = true;
  VVView view;
  uint32_t n_pad;
  FdPlan plan = fd_make_plan(job.t, job.n_r, job.parts, 0);
  if (job.shortcut && unsettled && (size_t)unsettled * 2 <= job.n_d && ((uintptr_t)job.d_shares & 15) == 0) {
    // per-dealer fallback: the few dealers left are evaluated as a dense session of their own (8 wrong dealers spread over 8 groups cost
    // one group, not eight); the settled dealers of their groups get their verdicts without evaluation
    const uint32_t* list;
    const uint8_t* sh_c;
    uint8_t* st_c;
    if (int rc = dkgv_fd_compact_unsettled(ctx, job.n_d, job.n_r, unsettled, job.d_shares, job.d_status, s, &list, &sh_c, &st_c)) return rc;
    if (int rc = session_decode(ctx, unsettled, job.t, job.d_vv, s, &view, &n_pad, nullptr, list)) return rc;
    if (int rc = dkgv_share_matrix_fd(ctx, view, unsettled, job.n_r, job.t, plan, job.d_ids, sh_c, st_c, nullptr, s)) return rc;
    return dkgv_fd_scatter_status(ctx, job.n_r, unsettled, job.d_status, s);
  }
  // only the commitments of the dealer groups that go on to the evaluation are decoded
  const uint8_t* need = job.shortcut ? dkgv_fd_need_groups(ctx, job.n_d) : nullptr;
  if (int rc = session_decode(ctx, job.n_d, job.t, job.d_vv, s, &view, &n_pad, need)) return rc;
  return dkgv_share_matrix_fd(ctx, view, job.n_d, job.n_r, job.t, plan, job.d_ids, job.d_shares, job.d_status, need, s);
}

int dkgv_share_submit_internal(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv, const uint32_t* d_ids,
                               const uint8_t* d_shares, uint8_t* d_status, uint32_t* d_flags, cudaStream_t s) {
  return share_submit(ctx, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, d_flags, s);
}
int dkgv_share_finish_internal(dkgv_ctx* ctx, const uint32_t* h_flags, cudaStream_t s) { return share_finish(ctx, h_flags, s); }

extern "C" int dkgv_share_matrix_submit_dev(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv,
                                            const uint32_t* d_ids, const uint8_t* d_shares, uint8_t* d_status, uint32_t* d_flags2,
                                            void* stream) {
  if (!ctx) return -1;
  if (!d_ids || !d_shares || !d_status || (t && !d_vv) || n_d == 0 || n_r == 0) return fail(ctx, "null pointer or empty matrix");
  CK(cudaSetDevice(ctx->device));
  return share_submit(ctx, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, d_flags2, stream ? (cudaStream_t)stream : ctx->stream);
}
extern "C" int dkgv_share_matrix_finish_dev(dkgv_ctx* ctx, const uint32_t* h_flags2, void* stream) {
  if (!ctx) return -1;
  CK(cudaSetDevice(ctx->device));
  return share_finish(ctx, h_flags2, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int dkgv_share_matrix_verify_dev(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* d_vv,
                                            const uint32_t* d_ids, const uint8_t* d_shares, uint8_t* d_status, void* stream) {
  if (!ctx) return -1;
  if (n_d == 0 || n_r == 0) return 0;
  if (!d_ids || !d_shares || !d_status || (t && !d_vv)) return fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  if (int rc = share_submit(ctx, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, nullptr, s)) return rc;
  return share_finish(ctx, nullptr, s);
}

extern "C" int dkgv_share_matrix_verify(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* vv,
                                        const uint32_t* ids, const uint8_t* shares, uint8_t* status) {
  if (!ctx) return -1;
  if (n_d == 0 || n_r == 0) return 0;
  if (!ids || !shares || !status || (t && !vv)) return fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t vvb = (size_t)n_d * t * 48, idb = (size_t)n_r * 4, shb = (size_t)n_d * n_r * 32, stb = (size_t)n_d * n_r;
  CK(ctx->in_a.reserve(vvb ? vvb : 1));
  CK(ctx->in_b.reserve(idb));
  CK(ctx->in_c.reserve(shb));
  CK(ctx->out_a.reserve(stb));
  // ids and shares first; the verification vectors follow on a second stream, so the share half of the default path (limbs, difference
  // tables) runs under their copy - the first kernel that reads them waits for ev_vv (dkgv_take_vv_wait)
  CK(cudaMemcpyAsync(ctx->in_b.p, ids, idb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_c.p, shares, shb, cudaMemcpyHostToDevice, s));
  if (vvb) {
    cudaStream_t s2 = ctx->fd_streams[0];
    CK(cudaEventRecord(ctx->ev_sh, s));
    CK(cudaStreamWaitEvent(s2, ctx->ev_sh, 0));
    CK(cudaMemcpyAsync(ctx->in_a.p, vv, vvb, cudaMemcpyHostToDevice, s2));
    CK(cudaEventRecord(ctx->ev_vv, s2));
    ctx->vv_wait = ctx->ev_vv;
  }
  if (int rc = share_submit(ctx, n_d, n_r, t, (const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, (const uint8_t*)ctx->in_c.p,
                            (uint8_t*)ctx->out_a.p, nullptr, s))
    return rc;
  if (int rc = dkgv_take_vv_wait(ctx, s)) return rc;  // (a path that never read them: the copy still ends before the call returns)
  // the verdicts and the two flag words come back together: ONE synchronisation when the shortcut settled the ceremony
  CK(cudaMemcpyAsync(status, ctx->out_a.p, stb, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(ctx->h_job_flags, ctx->job.d_flags, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const bool more = ctx->job.fd && (ctx->h_job_flags[0] || ctx->h_job_flags[1]);
  if (int rc = share_finish(ctx, ctx->h_job_flags, s)) return rc;
  if (more) {
    CK(cudaMemcpyAsync(status, ctx->out_a.p, stb, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return 0;
}

extern "C" int dkgv_pack_verdicts_dev(dkgv_ctx* ctx, uint64_t n, const uint8_t* d_status, uint32_t* d_bits, void* stream) {
  if (!ctx) return -1;
  if (n == 0) return 0;
  if (!d_status || !d_bits) return fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  size_t words = (size_t)((n + 31) / 32);
  k_pack_verdicts<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(d_status, d_bits, (size_t)n);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int dkgv_share_items_verify(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_r, uint32_t t, const uint8_t* vv, const uint32_t* ids,
                                       uint32_t m, const uint32_t* item_dealer, const uint32_t* item_recipient, const uint8_t* secrets,
                                       uint8_t* status) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (n_d == 0 || n_r == 0) return fail(ctx, "items given for an empty session");
  if (!ids || !item_dealer || !item_recipient || !secrets || !status || (t && !vv)) return fail(ctx, "null pointer argument");
  for (uint32_t i = 0; i < m; i++)
    if (item_dealer[i] >= n_d || item_recipient[i] >= n_r) return fail(ctx, "item index out of range");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // sort by recipient column (stable: dealers stay in caller order), cut into warps of one recipient each
  std::vector<uint32_t> perm(m);
  for (uint32_t i = 0; i < m; i++) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return item_recipient[a] < item_recipient[b]; });
  std::vector<uint32_t> sd(m), sc(m), wf;
  for (uint32_t i = 0; i < m; i++) {
    sd[i] = item_dealer[perm[i]];
    sc[i] = item_recipient[perm[i]];
    if (i == 0 || sc[i] != sc[i - 1] || i - wf.back() == 32) wf.push_back(i);
  }
  uint32_t n_warps = (uint32_t)wf.size();
  wf.push_back(m);
  size_t vvb = (size_t)n_d * t * 48;
  CK(ctx->in_a.reserve(vvb ? vvb : 1));
  CK(ctx->in_b.reserve((size_t)n_r * 4));
  CK(ctx->in_c.reserve((size_t)m * 32));
  CK(ctx->out_a.reserve(m));
  CK(ctx->scratch_a.reserve((size_t)m * 12 + (size_t)(n_warps + 1) * 4));
  uint32_t* d_sd = (uint32_t*)ctx->scratch_a.p;
  uint32_t *d_sc = d_sd + m, *d_so = d_sc + m, *d_wf = d_so + m;
  if (vvb) CK(cudaMemcpyAsync(ctx->in_a.p, vv, vvb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, ids, (size_t)n_r * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_c.p, secrets, (size_t)m * 32, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_sd, sd.data(), (size_t)m * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_sc, sc.data(), (size_t)m * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_so, perm.data(), (size_t)m * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_wf, wf.data(), (size_t)(n_warps + 1) * 4, cudaMemcpyHostToDevice, s));
  VVView view;
  uint32_t n_pad;
  int rc = session_decode(ctx, n_d, t, (const uint8_t*)ctx->in_a.p, s, &view, &n_pad);
  if (rc) return rc;
  k_share_items<<<n_warps, 32, (size_t)VM_SLOTS * 3 * 32 * sizeof(U4), s>>>(view, (const uint8_t*)ctx->dealer_bad.p,
                                                                           (const uint32_t*)ctx->in_b.p, d_sd, d_sc, d_so, d_wf,
                                                                           (const uint8_t*)ctx->in_c.p, ctx->gtab, (uint8_t*)ctx->out_a.p, t);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(status, ctx->out_a.p, m, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));  // also keeps the host vectors alive until the copies are done
  return 0;
}

extern "C" int dkgv_feldman_eval(dkgv_ctx* ctx, uint32_t n_d, uint32_t n_ids, uint32_t t, const uint8_t* vv, const uint32_t* ids,
                                 uint8_t* out, uint8_t* row_status) {
  if (!ctx) return -1;
  if (n_d == 0 || n_ids == 0) return 0;
  if (!ids || !out || !row_status || (t && !vv)) return fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t vvb = (size_t)n_d * t * 48, idb = (size_t)n_ids * 4, ob = (size_t)n_d * n_ids * 48;
  CK(ctx->in_a.reserve(vvb ? vvb : 1));
  CK(ctx->in_b.reserve(idb));
  CK(ctx->out_a.reserve(ob));
  if (vvb) CK(cudaMemcpyAsync(ctx->in_a.p, vv, vvb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, ids, idb, cudaMemcpyHostToDevice, s));
  VVView view;
  uint32_t n_pad;
  int rc = session_decode(ctx, n_d, t, (const uint8_t*)ctx->in_a.p, s, &view, &n_pad);
  if (rc) return rc;
  FdPlan plan{};
  bool use_fd = false;
  if (ctx->share_path != DKGV_SHARE_PATH_HORNER && t >= 2 && n_ids >= 3 && n_ids <= 65535 && dkgv_fd_ids_consecutive(ids, n_ids)) {
    plan = fd_make_plan(t, n_ids, ctx->share_parts);
    use_fd = plan.cost_fd != ~0ull && (plan.use || ctx->share_path == DKGV_SHARE_PATH_FDIFF);
  }
  if (use_fd) {
    if (int rc2 = dkgv_feldman_eval_fd(ctx, view, n_d, n_ids, t, plan, (const uint32_t*)ctx->in_b.p, ids, (uint8_t*)ctx->out_a.p, s)) return rc2;
  } else {
    dim3 grid(n_pad / 32, (n_ids + SV_WARPS - 1) / SV_WARPS);
    k_feldman_eval<<<grid, SV_WARPS * 32, 0, s>>>(view, (const uint32_t*)ctx->in_b.p, (uint8_t*)ctx->out_a.p, n_d, n_ids, t);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, ob, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(row_status, ctx->dealer_bad.p, n_d, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (uint32_t i = 0; i < n_d; i++) row_status[i] = row_status[i] ? DKGV_PANIC_BAD_G1 : DKGV_OK;
  return 0;
}

extern "C" int dkgv_g1_fixed_base_mul(dkgv_ctx* ctx, uint32_t m, const uint8_t* scalars, uint8_t* out, uint8_t* status) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!scalars || !out || !status) return fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve((size_t)m * 32));
  CK(ctx->out_a.reserve((size_t)m * 48));
  CK(ctx->out_b.reserve(m));
  CK(cudaMemcpyAsync(ctx->in_a.p, scalars, (size_t)m * 32, cudaMemcpyHostToDevice, s));
  k_fixed_base_mul<<<(m + 127) / 128, 128, 0, s>>>((const uint8_t*)ctx->in_a.p, ctx->gtab, (uint8_t*)ctx->out_a.p,
                                                  (uint8_t*)ctx->out_b.p, m);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, (size_t)m * 48, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(status, ctx->out_b.p, m, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int dkgv_g1_decompress_check(dkgv_ctx* ctx, uint32_t m, const uint8_t* in, uint8_t* decode_status) {
  if (!ctx) return -1;
  if (m == 0) return 0;
  if (!in || !decode_status) return fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  CK(ctx->in_a.reserve((size_t)m * 48));
  CK(ctx->out_b.reserve(m));
  CK(cudaMemcpyAsync(ctx->in_a.p, in, (size_t)m * 48, cudaMemcpyHostToDevice, s));
  k_g1_decompress_check<<<(m + 127) / 128, 128, 0, s>>>((const uint8_t*)ctx->in_a.p, (uint8_t*)ctx->out_b.p, m);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(decode_status, ctx->out_b.p, m, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int dkgv_fr_poly_eval(dkgv_ctx* ctx, uint32_t n_d, uint32_t t, const uint8_t* coeffs, uint32_t n_r, const uint32_t* ids,
                                 uint8_t* out) {
  if (!ctx) return -1;
  if (n_d == 0 || n_r == 0) return 0;
  if (!ids || !out || (t && !coeffs)) return fail(ctx, "null pointer argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  size_t cb = (size_t)n_d * t * 32, ob = (size_t)n_d * n_r * 32;
  CK(ctx->in_a.reserve(cb ? cb : 1));
  CK(ctx->in_b.reserve((size_t)n_r * 4));
  CK(ctx->out_a.reserve(ob));
  if (cb) CK(cudaMemcpyAsync(ctx->in_a.p, coeffs, cb, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->in_b.p, ids, (size_t)n_r * 4, cudaMemcpyHostToDevice, s));
  dim3 grid((n_r + 127) / 128, n_d);
  k_fr_poly_eval<<<grid, 128, 0, s>>>((const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, (uint8_t*)ctx->out_a.p, n_d, n_r, t);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->out_a.p, ob, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}
