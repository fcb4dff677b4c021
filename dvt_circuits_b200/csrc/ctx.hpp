// Shared host-side definitions of the C ABI implementation (ctx, error macro, device buffers).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/dkgv.h"
#include "gtab.hpp"

struct ncclUniqueIdBytes {
  char internal[128];
};
namespace dkgv_host {
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }  // dkgv_ctx_destroy selects the device before the ctx (and with it every buffer) is deleted
};
}  // namespace dkgv_host

struct dkgv_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint32_t* gtab_mem = nullptr;  // fixed-base table of the generator (feldman.cuh), built once per ctx
  dkgv::GTab gtab{nullptr, 0, 0};
  uint64_t launches = 0;
  std::string err;
  dkgv_host::DevBuf vv_limbs, vv_inf, dealer_bad;      // session scratch (decoded verification vectors)
  dkgv_host::DevBuf in_a, in_b, in_c, out_a, out_b;    // staging for the host-pointer entry points
  dkgv_host::DevBuf scratch_a, scratch_b, scratch_c, scratch_d;   // intermediates of the aggregation / pairing paths
  dkgv_host::DevBuf bls_pk, bls_sig, bls_st;           // decoded keys / signatures of a pairing batch
  dkgv_host::DevBuf bls_scratch;                       // Fp12 values the pairing VM parks in global memory
  cudaEvent_t ev_bls0 = nullptr, ev_bls1 = nullptr;    // bracket the pairing kernel
  // host-buffer share path: the verification vectors are copied on a second stream WHILE the difference tables run; the first kernel
  // that reads them waits for this event (dkgv_take_vv_wait), nullptr = nothing to wait for
  cudaEvent_t vv_wait = nullptr, ev_vv = nullptr, ev_sh = nullptr;
  bool bls_recorded = false, pvm_attr_set = false;
  int bls_path = 0, last_bls_path = 0;                 // enum dkgv_bls_path
  // communicator (comm.cu): NCCL, one process per GPU
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  uint64_t collectives = 0;
  dkgv_host::DevBuf comm_flags, comm_buf, share_gather;
  uint32_t* h_comm_flags = nullptr;
  size_t h_comm_flags_cap = 0;
  cudaEvent_t ev_hot0 = nullptr, ev_hot1 = nullptr;    // bracket the hot kernel (roofline timing)
  bool vv_decoded = true;            // the last share-matrix call decoded its commitments (false: settled against their encodings)
  cudaEvent_t ev_dec0 = nullptr, ev_dec1 = nullptr;  // bracket the last verification-vector decode of the share path
  bool dec_recorded = false;
  bool hot_recorded = false;
  bool stack_set = false;
  // finite-difference share path (share_fd.cu)
  dkgv_host::DevBuf fd_evals, fd_p0, fd_p1, fd_da, fd_db, fd_seedx, fd_dig, fd_top, fd_tab, fd_cols, fd_sl, fd_flags, fd_binom, fd_coef, fd_yz, fd_cmp_list, fd_cmp_sh, fd_cmp_st;
  std::vector<int32_t> fd_seed_host;
  cudaEvent_t ev_fd[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // phase boundaries of the evaluation
  cudaEvent_t ev_sc[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // ... of the consistency shortcut
  bool fd_recorded = false, sc_recorded = false;
  uint32_t fd_ipb_force = 0;  // items per block of the difference / extension launches (DKGV_FD_IPB, read once at create)
  // share-matrix job between dkgv_share_matrix_submit_dev and dkgv_share_matrix_finish_dev
  struct ShareJob {
    bool open = false, fd = false, shortcut = false;
    uint32_t n_d = 0, n_r = 0, t = 0;
    const uint8_t* d_vv = nullptr;
    const uint32_t* d_ids = nullptr;
    const uint8_t* d_shares = nullptr;
    uint8_t* d_status = nullptr;
    uint32_t* d_flags = nullptr;
    uint32_t parts = 0;
  } job;
  uint32_t* job_flags = nullptr;    // device: {ids are not a permutation of 1..n, dealers the shortcut could not settle}
  uint32_t* h_job_flags = nullptr;  // pinned host copy
  bool fd_overlap = true;  // one stream per part (default) or everything on the caller's stream
  bool fd_repair = true;     // repair route: Reed-Solomon decoding of the share sequence of an inconsistent dealer (share_rs.cuh)
  dkgv_host::DevBuf rs_tab, rs_work;
  uint32_t rs_n = 0, rs_t = 0;    // shape the cached tables belong to
  uint32_t last_repaired = 0;     // dealers the repair route settled in the last share-matrix call
  bool fd_polycheck = true;  // consistency shortcut: ids beyond t only for dealer groups that fail the scalar-side conditions
  uint32_t fd_binom_t = 0;   // t the cached binomial coefficients belong to
  bool fd_last_need = false; // the last finite-difference run had to continue beyond t for some dealer group
  cudaStream_t fd_streams[16] = {};
  cudaEvent_t fd_fork = nullptr, fd_join[16] = {};
  std::vector<uint32_t> fd_cols_host;
  int share_path = 0;       // DKGV_SHARE_PATH_* requested
  uint32_t share_parts = 0; // 0: planner's choice of parts per dealer; else forced
  int last_share_path = 0;  // path taken by the most recent share-matrix call
};

#define CK(call)                                                     \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); \
      return -2;                                                     \
    }                                                                \
  } while (0)

static inline int dkgv_fail(dkgv_ctx* ctx, const char* msg) {
  ctx->err = msg;
  return -1;
}
