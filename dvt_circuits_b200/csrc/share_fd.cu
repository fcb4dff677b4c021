// Share-matrix verification through finite differences (fdiff.cuh): the row-of-recipients form of
// verify_seed_exchange_commitment (crates/dkg/src/verification.rs:129-146) when the recipient ids
// are the consecutive ranks 1..n (verification.rs:50-66).  Each dealer polynomial is split into m
// parts of h coefficients ("virtual dealers", column part * n_pad + d of planes m * n_pad wide).
// Four phases on one stream:
//   k_fd_seed     h Horner evaluations of h-1 steps per virtual dealer (LPT order)
//   k_fd_init     h-1 rounds of pairwise differences           (1 point addition per item)
//   k_fd_ext      steps + h - 2 wavefront ticks                (1 point addition per item)
//   k_fd_digits   NAF digits of the public recombination scalars x^(h i) mod r, one thread per id
//   k_fd_combine  sum_i [y^i] f_i(x) by joint double-and-add, G * s, compare: one thread per share
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ctx.hpp"
#include "fdiff.cuh"

using namespace dkgv;

static uint32_t g_fd_ipb_force = 0;  // items per block of the difference / extension launches, DKGV_FD_IPB (experiments)
constexpr int FD_NT = 32;  // one warp per block: 32 consecutive dealers, one entry (cf. SVM_NT in dkgv.cu)
constexpr size_t FD_SMEM = (size_t)VM_SLOTS * 3 * FD_NT * sizeof(U4);

__global__ void __launch_bounds__(FD_NT)
k_fd_seed(VVView vv, const int32_t* __restrict__ seed_x, int32_t lo, uint32_t* __restrict__ evals, uint32_t n_d, uint32_t t,
          uint32_t h, uint32_t part0, uint32_t n_padv, uint32_t d0, uint32_t n_cols) {
  extern __shared__ U4 opfile[];
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
          uint32_t n_pad, uint32_t t, uint32_t r, uint32_t col0, uint32_t ipb) {
  extern __shared__ U4 opfile[];
  uint32_t d = col0 + blockIdx.x * 32 + threadIdx.x;
  OpFile f{opfile + threadIdx.x, FD_NT};
#pragma unroll 1
  for (uint32_t i = blockIdx.y * ipb; i < (blockIdx.y + 1) * ipb && i + r < t; i++) fd_init_item(f, src, dst, da, db, n_pad, t, r, i, d);
}

__global__ void __launch_bounds__(FD_NT)
k_fd_ext(const uint32_t* __restrict__ old, uint32_t* __restrict__ cur, uint32_t* __restrict__ evals, uint32_t n_pad, uint32_t t,
         uint32_t tick, uint32_t k_lo, uint32_t k_hi, size_t e_hi, uint32_t col0, uint32_t ipb) {
  extern __shared__ U4 opfile[];
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
             uint32_t n_d, uint32_t n_r, uint32_t j0, const uint32_t* __restrict__ cols, uint32_t tab_r0, uint32_t d0) {
  extern __shared__ U4 opfile[];
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

// evaluate_polynomial output instead of the share comparison: out[dealer][column j] = compress(f_d(ids[j]))
__global__ void __launch_bounds__(FD_NT)
k_fd_combine_out(const uint32_t* __restrict__ evals, int32_t lo, uint32_t m, const int8_t* __restrict__ dig, const int32_t* __restrict__ top,
                 const uint32_t* __restrict__ ids, uint8_t* __restrict__ out48, uint32_t* __restrict__ tab, uint32_t n_pad, uint32_t n_d,
                 uint32_t n_r, uint32_t j0, uint32_t tab_r0, uint32_t d0) {
  extern __shared__ U4 opfile[];
  uint32_t d = blockIdx.x * 32 + threadIdx.x;
  uint32_t j = j0 + blockIdx.y;
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
                        (const void*)k_fd_combine_out}) {
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FD_SMEM));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  if (const char* e = getenv("DKGV_FD_IPB")) g_fd_ipb_force = (uint32_t)atoi(e);
  for (int i = 0; i < 5; i++) CK(cudaEventCreate(&ctx->ev_fd[i]));
  CK(cudaEventCreateWithFlags(&ctx->fd_fork, cudaEventDisableTiming));
  // the extension is a long dependent chain of small launches: its streams get the highest priority so
  // that recombination blocks (comb stream, lowest priority) only fill the slots it leaves idle
  int prio_least = 0, prio_greatest = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  for (uint32_t i = 0; i < FD_MAX_PARTS; i++) {
    CK(cudaStreamCreateWithPriority(&ctx->fd_streams[i], cudaStreamNonBlocking, prio_greatest));
    CK(cudaEventCreateWithFlags(&ctx->fd_join[i], cudaEventDisableTiming));
  }
  CK(cudaStreamCreateWithPriority(&ctx->fd_comb_stream, cudaStreamNonBlocking, prio_least));
  CK(cudaEventCreateWithFlags(&ctx->fd_comb_done, cudaEventDisableTiming));
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

static inline const uint32_t* evals_c(dkgv_ctx* ctx) { return (const uint32_t*)ctx->fd_evals.p; }
// Items (point additions) per block of the difference / extension launches.  One per block is the
// measured optimum on B200 (n=1024, t=683, N=1: extension 336 ms with 1, 344 ms with 2, 367 ms with 4 items):
// block scheduling is not what the one-addition blocks lose time on.  DKGV_FD_IPB overrides (experiments).
static inline uint32_t items_per_block(uint32_t, uint32_t) { return g_fd_ipb_force ? g_fd_ipb_force : 1; }
constexpr uint32_t FD_COMB_CHUNKS = 8;  // recombination launches pipelined behind the extension

// dealers d0 .. d0 + n_pad - 1 (n_pad a multiple of 32: the column count of this chunk's planes)
static int fd_run(dkgv_ctx* ctx, const VVView& view, uint32_t d0, uint32_t n_pad, uint32_t n_d, uint32_t n_r, uint32_t t,
                  const FdPlan& plan, const uint32_t* d_ids, const uint32_t* h_ids, const uint8_t* d_shares, uint8_t* d_status,
                  uint8_t* d_out48, cudaStream_t s) {
  const uint32_t m = plan.m, h = plan.h;
  const uint32_t n_padv = n_pad * m;  // plane width: one column per virtual dealer
  const size_t ent_words = (size_t)36 * n_padv, ent_bytes = ent_words * 4;
  const size_t n_evals = (size_t)((int64_t)n_r - plan.lo + 1);
  CK(ctx->fd_evals.reserve(n_evals * ent_bytes));
  CK(ctx->fd_p0.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_p1.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_da.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_db.reserve((size_t)h * ent_bytes));
  CK(ctx->fd_seedx.reserve((size_t)h * 4));
  CK(ctx->fd_dig.reserve((size_t)n_r * fd_dig_bytes(m)));
  // per-share tables of the recombination: recipients are processed in chunks that fit the budget
  const size_t tab_per_recipient = (size_t)(m > 1 ? m - 1 : 0) * FD_TAB_SLOTS * 36 * n_pad * 4;
  const size_t tab_budget = (size_t)4 << 30;
  uint32_t chunk_r = n_r;
  if (tab_per_recipient && tab_per_recipient * chunk_r > tab_budget) chunk_r = (uint32_t)(tab_budget / tab_per_recipient);
  if (chunk_r == 0) chunk_r = 1;
  CK(ctx->fd_tab.reserve(tab_per_recipient ? tab_per_recipient * chunk_r : 16));
  // columns in ascending-id order: the extension produces f(x) in that order, so the recombination of a
  // range of ids can start as soon as the wavefront has passed it
  const bool pipelined = ctx->fd_overlap && ctx->fd_pipeline && m > 1 && chunk_r == n_r && !d_out48;
  if (pipelined) {
    ctx->fd_cols_host.resize(n_r);
    for (uint32_t j = 0; j < n_r; j++) ctx->fd_cols_host[h_ids[j] - 1] = j;
    CK(ctx->fd_cols.reserve((size_t)n_r * 4));
    CK(cudaMemcpyAsync(ctx->fd_cols.p, ctx->fd_cols_host.data(), (size_t)n_r * 4, cudaMemcpyHostToDevice, s));
    while (ctx->fd_chunk_ev.size() < (size_t)(FD_COMB_CHUNKS + 1) * FD_MAX_PARTS) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      ctx->fd_chunk_ev.push_back(ev);
    }
  }
  auto launch_combine = [&](cudaStream_t cs, uint32_t r0, uint32_t nj, const uint32_t* cols, uint32_t tab_r0) {
    if (d_out48) {  // evaluation output (cols == nullptr: this mode is never pipelined)
      k_fd_combine_out<<<dim3(n_pad / 32, nj), FD_NT, FD_SMEM, cs>>>(evals_c(ctx), plan.lo, m, (const int8_t*)ctx->fd_dig.p,
                                                                     (const int32_t*)ctx->fd_top.p, d_ids, d_out48,
                                                                     (uint32_t*)ctx->fd_tab.p, n_pad, n_d, n_r, r0, tab_r0, d0);
      ctx->launches++;
      return;
    }
    k_fd_combine<<<dim3(n_pad / 32, nj), FD_NT, FD_SMEM, cs>>>(evals_c(ctx), plan.lo, m, (const int8_t*)ctx->fd_dig.p,
                                                               (const int32_t*)ctx->fd_top.p, d_ids, d_shares, ctx->gtab,
                                                               (const uint8_t*)ctx->dealer_bad.p, d_status, (uint32_t*)ctx->fd_tab.p,
                                                               n_pad, n_d, n_r, r0, cols, tab_r0, d0);
    ctx->launches++;
  };
  // recombination of the ids (x0, x1] on the comb stream once every part stream has produced them
  uint32_t chunk_no = 0, next_chunk = 1;
  auto chunk_end = [&](uint32_t c) { return (uint32_t)(((uint64_t)plan.steps * c + FD_COMB_CHUNKS - 1) / FD_COMB_CHUNKS); };
  auto combine_after_parts = [&](uint32_t x0, uint32_t x1) -> int {
    for (uint32_t p = 0; p < m; p++) {
      cudaEvent_t ev = ctx->fd_chunk_ev[(size_t)chunk_no * FD_MAX_PARTS + p];
      CK(cudaEventRecord(ev, ctx->fd_streams[p]));
      CK(cudaStreamWaitEvent(ctx->fd_comb_stream, ev, 0));
    }
    chunk_no++;
    if (x1 > x0) launch_combine(ctx->fd_comb_stream, x0, x1 - x0, (const uint32_t*)ctx->fd_cols.p, x0);
    return 0;
  };
  CK(ctx->fd_top.reserve((size_t)n_r * 4));
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

  const unsigned gx = n_pad / 32, gxv = n_padv / 32;
  const size_t e_hi = (size_t)(plan.hi - plan.lo);  // == h - 1
  const uint32_t ticks = plan.steps + h - 2;
  const bool overlap = ctx->fd_overlap && m > 1;
  CK(cudaEventRecord(ctx->ev_fd[0], s));
  CK(cudaEventRecord(ctx->ev_hot0, s));
  if (!overlap) {
    // one stream, phase after phase over all parts at once (also the mode that yields per-phase times)
    k_fd_seed<<<dim3(gx, h, m), FD_NT, FD_SMEM, s>>>(view, (const int32_t*)ctx->fd_seedx.p, plan.lo, evals, n_d, t, h, 0, n_padv, d0, n_pad);
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
      k_fd_init<<<dim3(gxv, (h - r + ipb - 1) / ipb), FD_NT, FD_SMEM, s>>>(src, dst, dd[0], dd[1], n_padv, h, r, 0, ipb);
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
                                                                      (uint32_t)k_lo, (uint32_t)k_hi, e_hi, 0, ipb);
      ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev_fd[3], s));
  } else {
    // The parts are independent until the recombination: each runs its seed -> differences -> extension
    // chain on its own stream, so the tail of one part's kernel is filled by blocks of the others
    // (no grid-wide barrier per round / tick; matters when a rank holds few dealers).
    CK(cudaEventRecord(ctx->fd_fork, s));
    if (pipelined) {
      CK(cudaStreamWaitEvent(ctx->fd_comb_stream, ctx->fd_fork, 0));
      k_fd_digits<<<(n_r + 127) / 128, 128, 0, ctx->fd_comb_stream>>>(n_r, h, m, (int8_t*)ctx->fd_dig.p, (int32_t*)ctx->fd_top.p);
      ctx->launches++;
    }
    for (uint32_t p = 0; p < m; p++) {
      cudaStream_t sp = ctx->fd_streams[p];
      CK(cudaStreamWaitEvent(sp, ctx->fd_fork, 0));
      k_fd_seed<<<dim3(gx, h, 1), FD_NT, FD_SMEM, sp>>>(view, (const int32_t*)ctx->fd_seedx.p, plan.lo, evals, n_d, t, h, p, n_padv, d0,
                                                        n_pad);
      ctx->launches++;
      const size_t w = (size_t)n_pad * 4, pitch = (size_t)n_padv * 4;
      const uint32_t* col = evals + e_hi * ent_words + (size_t)p * n_pad;
      CK(cudaMemcpy2DAsync(dd[0] + (size_t)p * n_pad, pitch, col, pitch, w, 36, cudaMemcpyDeviceToDevice, sp));
      CK(cudaMemcpy2DAsync(dd[1] + (size_t)p * n_pad, pitch, col, pitch, w, 36, cudaMemcpyDeviceToDevice, sp));
    }
    if (pipelined)  // ids 1..hi are seed values
      if (int rc = combine_after_parts(0, (uint32_t)plan.hi)) return rc;
    for (uint32_t r = 1; r < h; r++) {
      uint32_t ipb = items_per_block(gxv, h - r);
      for (uint32_t p = 0; p < m; p++) {
        k_fd_init<<<dim3(gx, (h - r + ipb - 1) / ipb), FD_NT, FD_SMEM, ctx->fd_streams[p]>>>(r == 1 ? evals : pp[(r - 1) & 1], pp[r & 1], dd[0],
                                                                                            dd[1], n_padv, h, r, p * n_pad, ipb);
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
                                                                                         p * n_pad, ipb);
        ctx->launches++;
      }
      if (pipelined && tick >= h - 1) {  // step tick - (h - 2) of every part is queued: ids up to hi + that step exist after it
        uint32_t sdone = tick - (h - 2);
        while (next_chunk <= FD_COMB_CHUNKS && chunk_end(next_chunk) <= sdone) {
          uint32_t s_beg = chunk_end(next_chunk - 1), s_end = chunk_end(next_chunk);
          next_chunk++;
          if (s_end > s_beg)
            if (int rc = combine_after_parts((uint32_t)plan.hi + s_beg, (uint32_t)plan.hi + s_end)) return rc;
        }
      }
    }
    for (uint32_t p = 0; p < m; p++) {
      CK(cudaEventRecord(ctx->fd_join[p], ctx->fd_streams[p]));
      CK(cudaStreamWaitEvent(s, ctx->fd_join[p], 0));
    }
    if (pipelined) {
      CK(cudaEventRecord(ctx->fd_comb_done, ctx->fd_comb_stream));
      CK(cudaStreamWaitEvent(s, ctx->fd_comb_done, 0));
    }
    CK(cudaEventRecord(ctx->ev_hot1, s));
    for (int i = 1; i <= 3; i++) CK(cudaEventRecord(ctx->ev_fd[i], s));  // phases overlap: only their sum is defined
  }
  ctx->hot_recorded = true;

  if (!pipelined) {
    if (m > 1) {
      k_fd_digits<<<(n_r + 127) / 128, 128, 0, s>>>(n_r, h, m, (int8_t*)ctx->fd_dig.p, (int32_t*)ctx->fd_top.p);
      ctx->launches++;
    }
    for (uint32_t j0 = 0; j0 < n_r; j0 += chunk_r) launch_combine(s, j0, n_r - j0 < chunk_r ? n_r - j0 : chunk_r, nullptr, 0);
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
