// Share-matrix verification through finite differences (fdiff.cuh): the row-of-recipients form of
// verify_seed_exchange_commitment (crates/dkg/src/verification.rs:129-146) when the recipient ids
// are the consecutive ranks 1..n (verification.rs:50-66).  Four phases on one stream:
//   k_fd_seed     t Horner evaluations per dealer (the integer-pipe bound kernel; LPT order)
//   k_fd_init     t-1 rounds of pairwise differences           (1 point addition per item)
//   k_fd_ext      steps + t - 2 wavefront ticks                (1 point addition per item)
//   k_fd_compare  G * s against the evaluation, one thread per share
#include <algorithm>
#include <vector>

#include "ctx.hpp"
#include "fdiff.cuh"

using namespace dkgv;

constexpr int FD_NT = 32;  // one warp per block: 32 consecutive dealers, one entry (cf. SVM_NT in dkgv.cu)
constexpr size_t FD_SMEM = (size_t)VM_SLOTS * 3 * FD_NT * sizeof(U4);

__global__ void __launch_bounds__(FD_NT)
k_fd_seed(VVView vv, const int32_t* __restrict__ seed_x, int32_t lo, uint32_t* __restrict__ evals, uint32_t n_d, uint32_t t) {
  extern __shared__ U4 opfile[];
  uint32_t d = blockIdx.x * 32 + threadIdx.x;
  int32_t x = seed_x[blockIdx.y];  // most expensive points first
  uint32_t dd = d < n_d ? d : n_d - 1;
  OpFile f{opfile + threadIdx.x, FD_NT};
  fd_seed_eval(f, vv, t, dd, x);
  fd_store(f, AX, fd_entry(evals, vv.n_pad, (size_t)(x - lo), d), vv.n_pad);
}

__global__ void __launch_bounds__(FD_NT)
k_fd_init(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t* __restrict__ da, uint32_t* __restrict__ db,
          uint32_t n_pad, uint32_t t, uint32_t r) {
  extern __shared__ U4 opfile[];
  uint32_t d = blockIdx.x * 32 + threadIdx.x;
  OpFile f{opfile + threadIdx.x, FD_NT};
  fd_init_item(f, src, dst, da, db, n_pad, t, r, blockIdx.y, d);
}

__global__ void __launch_bounds__(FD_NT)
k_fd_ext(const uint32_t* __restrict__ old, uint32_t* __restrict__ cur, uint32_t* __restrict__ evals, uint32_t n_pad, uint32_t t,
         uint32_t tick, uint32_t k_lo, size_t e_hi) {
  extern __shared__ U4 opfile[];
  uint32_t d = blockIdx.x * 32 + threadIdx.x;
  OpFile f{opfile + threadIdx.x, FD_NT};
  fd_ext_item(f, old, cur, evals, n_pad, t, tick, k_lo + blockIdx.y, e_hi, d);
}

__global__ void __launch_bounds__(FD_NT)
k_fd_compare(const uint32_t* __restrict__ evals, int32_t lo, const uint32_t* __restrict__ ids, const uint8_t* __restrict__ shares,
             const uint32_t* __restrict__ gtab, const uint8_t* __restrict__ dealer_bad, uint8_t* __restrict__ status,
             uint32_t n_pad, uint32_t n_d, uint32_t n_r) {
  extern __shared__ U4 opfile[];
  uint32_t d = blockIdx.x * 32 + threadIdx.x;
  uint32_t j = blockIdx.y;
  bool active = d < n_d;
  uint32_t dd = active ? d : n_d - 1;
  OpFile f{opfile + threadIdx.x, FD_NT};
  size_t e = (size_t)((int64_t)ids[j] - lo);
  uint8_t st = fd_compare_item(f, fd_entry(evals, n_pad, e, dd), n_pad, shares + ((size_t)dd * n_r + j) * 32, gtab,
                               dealer_bad[dd] != 0);
  if (active) status[(size_t)d * n_r + j] = st;
}

int dkgv_fd_setup(dkgv_ctx* ctx) {
  for (const void* k : {(const void*)k_fd_seed, (const void*)k_fd_init, (const void*)k_fd_ext, (const void*)k_fd_compare}) {
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FD_SMEM));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  for (int i = 0; i < 5; i++) CK(cudaEventCreate(&ctx->ev_fd[i]));
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

int dkgv_share_matrix_fd(dkgv_ctx* ctx, const VVView& view, uint32_t n_d, uint32_t n_r, uint32_t t, const FdPlan& plan,
                         const uint32_t* d_ids, const uint8_t* d_shares, uint8_t* d_status, cudaStream_t s) {
  const uint32_t n_pad = view.n_pad;
  const size_t ent_words = (size_t)36 * n_pad, ent_bytes = ent_words * 4;
  const size_t n_evals = (size_t)((int64_t)n_r - plan.lo + 1);
  CK(ctx->fd_evals.reserve(n_evals * ent_bytes));
  CK(ctx->fd_p0.reserve((size_t)t * ent_bytes));
  CK(ctx->fd_p1.reserve((size_t)t * ent_bytes));
  CK(ctx->fd_da.reserve((size_t)t * ent_bytes));
  CK(ctx->fd_db.reserve((size_t)t * ent_bytes));
  CK(ctx->fd_seedx.reserve((size_t)t * 4));
  uint32_t* evals = (uint32_t*)ctx->fd_evals.p;
  uint32_t* pp[2] = {(uint32_t*)ctx->fd_p0.p, (uint32_t*)ctx->fd_p1.p};
  uint32_t* dd[2] = {(uint32_t*)ctx->fd_da.p, (uint32_t*)ctx->fd_db.p};

  // seed points, most expensive first (blocks are dispatched in increasing blockIdx.y)
  std::vector<int32_t> order(t);
  for (uint32_t i = 0; i < t; i++) order[i] = plan.lo + (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    return fd_horner_cost(t, (uint32_t)(a < 0 ? -a : a)) > fd_horner_cost(t, (uint32_t)(b < 0 ? -b : b));
  });
  ctx->fd_seed_host.assign(order.begin(), order.end());  // must outlive the async copy
  CK(cudaMemcpyAsync(ctx->fd_seedx.p, ctx->fd_seed_host.data(), (size_t)t * 4, cudaMemcpyHostToDevice, s));

  dim3 gx(n_pad / 32);
  CK(cudaEventRecord(ctx->ev_fd[0], s));
  CK(cudaEventRecord(ctx->ev_hot0, s));
  k_fd_seed<<<dim3(gx.x, t), FD_NT, FD_SMEM, s>>>(view, (const int32_t*)ctx->fd_seedx.p, plan.lo, evals, n_d, t);
  CK(cudaEventRecord(ctx->ev_hot1, s));
  CK(cudaEventRecord(ctx->ev_fd[1], s));
  ctx->hot_recorded = true;
  ctx->launches++;

  // backward differences of f at hi
  const size_t e_hi = (size_t)(plan.hi - plan.lo);  // == t - 1
  CK(cudaMemcpyAsync(dd[0], evals + e_hi * ent_words, ent_bytes, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(dd[1], evals + e_hi * ent_words, ent_bytes, cudaMemcpyDeviceToDevice, s));
  const uint32_t* src = evals;
  for (uint32_t r = 1; r < t; r++) {
    uint32_t* dst = pp[r & 1];
    k_fd_init<<<dim3(gx.x, t - r), FD_NT, FD_SMEM, s>>>(src, dst, dd[0], dd[1], n_pad, t, r);
    ctx->launches++;
    src = dst;
  }
  CK(cudaEventRecord(ctx->ev_fd[2], s));

  // wavefront extension hi+1 .. n_r
  const uint32_t ticks = plan.steps + t - 2;
  for (uint32_t tick = 1; tick <= ticks; tick++) {
    int32_t k_lo, k_hi;
    fd_ext_band(t, plan.steps, tick, &k_lo, &k_hi);
    if (k_lo > k_hi) continue;
    k_fd_ext<<<dim3(gx.x, (unsigned)(k_hi - k_lo + 1)), FD_NT, FD_SMEM, s>>>(dd[(tick & 1) ^ 1], dd[tick & 1], evals, n_pad, t, tick,
                                                                             (uint32_t)k_lo, e_hi);
    ctx->launches++;
  }
  CK(cudaEventRecord(ctx->ev_fd[3], s));

  k_fd_compare<<<dim3(gx.x, n_r), FD_NT, FD_SMEM, s>>>(evals, plan.lo, d_ids, d_shares, ctx->gtab, (const uint8_t*)ctx->dealer_bad.p,
                                                       d_status, n_pad, n_d, n_r);
  ctx->launches++;
  CK(cudaEventRecord(ctx->ev_fd[4], s));
  ctx->fd_recorded = true;
  CK(cudaGetLastError());
  return 0;
}
