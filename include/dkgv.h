/* dkgv.h - C ABI of the B200-native batch verifier for the DKG checks of metacraft-labs/dvt-circuits
 * `crates/dkg`.  This is the drop-in boundary: a Rust `-sys` crate binds
 * exactly these symbols; no CUDA or torch types appear in the signatures.
 *
 * Conventions
 *  - All byte encodings are the reference's wire formats (crates/dkg/src/types.rs:396-407):
 *    G1 = 48 B compressed big-endian, G2 = 96 B compressed (c1 || c0), scalars = 32 B big-endian,
 *    hashes = 32 B.  Arrays are dense row-major.
 *  - Every function returns 0 on success, <0 on argument / CUDA error (message via
 *    dkgv_last_error).  Per-item outcomes are written to caller-allocated `status` arrays using
 *    the dkgv_status codes below: one code per distinct exit of the reference
 *    (Ok / SlashableError / UnslashableError / io::Error / panic!).
 *  - Pointers named `h_*`/unprefixed are HOST pointers; functions ending in `_dev` take DEVICE
 *    pointers (current device of the ctx) and a stream handle.  They queue their work on that stream
 *    and do not synchronise it, with ONE documented exception: dkgv_share_matrix_verify_dev reads a
 *    two-word flag back (one stream synchronisation) to decide whether anything is left to evaluate;
 *    its split form dkgv_share_matrix_submit_dev / dkgv_share_matrix_finish_dev leaves that read-back
 *    to the caller and is fully asynchronous (CUDA-graph capturable once its buffers exist); the sharded
 *    synchronous calls (dkgv_share_matrix_verify_sharded_dev, dkgv_agg_final_keys_sharded) synchronise once
 *    after their all-gather, their pipelined forms (dkgv_share_matrix_enqueue_sharded[_dev] + _settle_) do not.
 *  - A ctx is single-owner (Send, not Sync), bound to one GPU.  Multi-GPU: one process per GPU, a
 *    dkgv_comm per ctx (NCCL inside the library: dkgv_comm_*, *_sharded entry points).  Several ctxs of one
 *    process may share a GPU (one per ceremony in flight); they share ONE fixed-base table per window width.
 *  - There is no CPU fallback: without a CUDA device dkgv_ctx_create fails.
 */
#ifndef DKGV_H
#define DKGV_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dkgv_ctx dkgv_ctx;

/* per-item outcome codes; comments cite the reference exit each one stands for */
enum dkgv_status {
  DKGV_OK = 0,
  DKGV_SLASHABLE_SECRET_RANGE = 1,   /* verification.rs:92-99   secret >= r                      */
  DKGV_SLASHABLE_COMMIT_HASH = 2,    /* verification.rs:101-114 [auth] seed-exchange hash        */
  DKGV_SLASHABLE_DST_NOT_FOUND = 3,  /* verification.rs:116-126 dst_base_hash not in base_hashes */
  DKGV_SLASHABLE_SHARE_MISMATCH = 4, /* verification.rs:140-146 G*s != sum C_k id^k              */
  DKGV_SLASHABLE_BAD_PK = 5,         /* verification.rs:440-446 partial_pubkey undecodable       */
  DKGV_SLASHABLE_BAD_SIG = 6,        /* verification.rs:448-454 message_signature undecodable    */
  DKGV_SLASHABLE_SIG_INVALID = 7,    /* verification.rs:456-461 pairing check false              */
  DKGV_SLASHABLE_KEY_MISMATCH = 8,   /* verification.rs:413-418 expected key != partial key (Q1) */
  DKGV_SLASHABLE_BAD_ENCRYPTED_MSG = 9, /* bad_encrypted_share_prove/main.rs:359-369 decrypted message does not parse */
  DKGV_UNSLASHABLE_COMMIT_SIG = 16,  /* verification.rs:76-89, 488-493 identity signature        */
  DKGV_UNSLASHABLE_COMMIT_HASH = 17, /* verification.rs:470-477                                  */
  DKGV_UNSLASHABLE_GEN_HASH = 18,    /* verification.rs:250-257, 388-393                         */
  DKGV_UNSLASHABLE_PERP_NOT_FOUND = 19, /* verification.rs:511-519                               */
  DKGV_UNSLASHABLE_SIG_INVALID = 20, /* verification.rs:243-248                                  */
  DKGV_ERR_LEN = 32,                 /* verification.rs:218-223, 270-275; dkg_math.rs:183-188    */
  DKGV_ERR_MSG_MISMATCH = 33,        /* verification.rs:224-231                                  */
  DKGV_ERR_AGG_MISMATCH_VV = 34,     /* verification.rs:301-309                                  */
  DKGV_ERR_AGG_MISMATCH_PK = 35,     /* verification.rs:320-328                                  */
  DKGV_ERR_ZERO_ID = 36,             /* dkg_math.rs:201-206                                      */
  DKGV_ERR_DUP_ID = 37,              /* dkg_math.rs:212-217                                      */
  DKGV_PANIC_BAD_G1 = 48,            /* .expect on G1 decode: verification.rs:136,241,288,314    */
  DKGV_PANIC_BAD_G2 = 49,            /* .expect on G2 decode: verification.rs:238                */
  DKGV_PANIC_BAD_SCALAR = 50,        /* dkg_math.rs:84                                           */
  DKGV_PANIC_INDEX = 51,             /* dkg_math.rs:235-239 ragged verification vectors          */
  DKGV_PANIC_PRECHECK = 52,          /* guest pre-checks, bad_share_exchange_prove/main.rs:24-43 */
  DKGV_PANIC_BAD_IDENTITY = 53       /* verification.rs:369-372, 478-481                         */
};

/* G1 / G2 decode results (dkgv_g1_decompress_check) */
enum dkgv_decode { DKGV_DEC_OK = 0, DKGV_DEC_BAD_FLAGS = 1, DKGV_DEC_X_RANGE = 2, DKGV_DEC_NOT_ON_CURVE = 3, DKGV_DEC_NOT_IN_SUBGROUP = 4 };

/* ---- context ------------------------------------------------------------------------------ */
/* Binds to CUDA device `device`, builds the fixed-base table for G1 generator multiplication.   */
int dkgv_ctx_create(int device, dkgv_ctx** out);
/* Same with the window width of the fixed-base table chosen by the caller: G * s (bls_keys.rs:98-114; t times per dealer on the default
 * share path) costs ceil(256 / gtab_bits) - 1 mixed additions against a table of ceil(256 / gtab_bits) * 2^(gtab_bits - 1) affine
 * points (96 B each) in device memory - 16: 50 MB / 15 additions, 22: 2.4 GB / 11, 26: 32 GB / 9.  gtab_bits = 0: the environment
 * variable DKGV_GTAB_BITS if set, else 22.  8 <= gtab_bits <= 26; when the device cannot hold the table the width steps down by 2
 * (not below 16) - dkgv_gtab_bits tells what the ctx ended up with. */
int dkgv_ctx_create_ex(int device, uint32_t gtab_bits, dkgv_ctx** out);
uint32_t dkgv_gtab_bits(const dkgv_ctx* ctx);
/* debug: compares `count` entries of the fixed-base table (index first, first + stride, ...) with their definition computed the slow
 * way on the device; *n_bad = how many differ.  Synchronous. */
int dkgv_gtab_selfcheck(dkgv_ctx* ctx, uint32_t first, uint32_t stride, uint32_t count, uint32_t* n_bad);
void dkgv_ctx_destroy(dkgv_ctx* ctx);
const char* dkgv_last_error(const dkgv_ctx* ctx); /* ctx may be NULL: last create error */
/* number of kernel launches issued through this ctx so far (bench accounting) */
uint64_t dkgv_launch_count(const dkgv_ctx* ctx);
int dkgv_sync(dkgv_ctx* ctx);
/* device time (CUDA events on the launching stream) of the most recent hot-kernel launch
 * (k_share_verify) issued through this ctx; blocks until that launch has finished */
int dkgv_last_hot_kernel_ms(dkgv_ctx* ctx, float* ms);
/* device time of the most recent verification-vector decode of the share path (k_decompress_vv) and whether it included the
 * subgroup checks.  A share-matrix call settled by the consistency shortcut decodes nothing (dkgv_last_share_decoded == 0):
 * the value then belongs to an earlier call; fails when this ctx has not decoded yet.                                 */
int dkgv_last_decode_ms(dkgv_ctx* ctx, float* ms, int* subgroup_checked);

/* ---- Feldman share verification (replaces the loop body of verify_seed_exchange_commitment,
 *      crates/dkg/src/verification.rs:129-146, for a whole (dealer x recipient) matrix) ------- */
/* vv      [n_dealers][t][48]   verification vectors (base_pubkeys of each dealer)
 * ids     [n_recipients]       recipient id = 1 + index of dst_base_hash in sorted base_hashes
 * shares  [n_dealers][n_recipients][32] big-endian secrets
 * status  [n_dealers][n_recipients]  out: OK / SLASHABLE_SECRET_RANGE / PANIC_BAD_G1 / SLASHABLE_SHARE_MISMATCH
 * t == 0 evaluates to the identity, t == 1 to C_0 (dkg_math.rs:161-166).                        */
int dkgv_share_matrix_verify(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t n_recipients, uint32_t t,
                             const uint8_t* vv, const uint32_t* ids, const uint8_t* shares, uint8_t* status);
/* same with device-resident buffers on `stream` (a cudaStream_t, may be NULL = the ctx stream): submit + finish below, i.e.
 * everything is queued asynchronously except ONE stream synchronisation that reads the two flag words.                   */
int dkgv_share_matrix_verify_dev(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t n_recipients, uint32_t t,
                                 const uint8_t* d_vv, const uint32_t* d_ids, const uint8_t* d_shares,
                                 uint8_t* d_status, void* stream);
/* The same call in two halves, for callers that overlap it with other work or capture it into a CUDA graph.
 * submit: queues the default path on `stream` and returns WITHOUT synchronising.  When the ids look like ceremony ranks the
 *   consistency shortcut (below) runs speculatively while a kernel checks that they really are a permutation of 1..n;
 *   d_status receives OK for every dealer group the shortcut settles.  d_flags2 (device, 2 x u32; NULL = a ctx-owned pair):
 *     [0] != 0  the ids are not a permutation of 1..n_recipients: nothing written so far counts;
 *     [1]       number of dealers the shortcut could not settle (their 32-dealer groups still hold no verdicts).
 * finish: h_flags2 = those two words as the caller read them back after synchronising (typically in the same copy as its
 *   results), or NULL to let the library read them (one synchronisation).  {0, 0}: returns at once - every verdict is OK
 *   and already in d_status.  Otherwise queues the evaluation of the pending dealer groups (or the Horner route over the
 *   whole matrix) on `stream`, asynchronously.  The buffers of submit must stay valid until finish's work has run.     */
int dkgv_share_matrix_submit_dev(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t n_recipients, uint32_t t, const uint8_t* d_vv,
                                 const uint32_t* d_ids, const uint8_t* d_shares, uint8_t* d_status, uint32_t* d_flags2, void* stream);
int dkgv_share_matrix_finish_dev(dkgv_ctx* ctx, const uint32_t* h_flags2, void* stream);

/* Sparse item list over one session (e.g. only the complained-about shares): item i is the pair
 * (item_dealer[i] < n_dealers, item_recipient[i] < n_recipients = column into ids) with secret secrets[i][32];
 * status[i] as in the matrix call.  Horner per item; the items are grouped by recipient internally.   */
int dkgv_share_items_verify(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t n_recipients, uint32_t t, const uint8_t* vv,
                            const uint32_t* ids, uint32_t m, const uint32_t* item_dealer, const uint32_t* item_recipient,
                            const uint8_t* secrets, uint8_t* status);
/* verdict bitmask of a status array on the device: bit (i % 32) of word i / 32 is set when status[i] != OK
 * (ceil(n / 32) words; what the ranks exchange instead of the status bytes).  Asynchronous on `stream`. */
int dkgv_pack_verdicts_dev(dkgv_ctx* ctx, uint64_t n, const uint8_t* d_status, uint32_t* d_bits, void* stream);

/* Evaluation strategy of the share-matrix entry points.  AUTO (default): when the recipient ids are a
 * permutation of 1..n_recipients (always so for a ceremony: verification.rs:50-66,129), split every
 * dealer polynomial into `parts` pieces of h coefficients, evaluate h points per piece by Horner, all
 * further points by finite differences, and recombine per share (exact group arithmetic, identical
 * verdicts); otherwise, or when that is not cheaper, Horner per share.                              */
enum dkgv_share_path { DKGV_SHARE_PATH_AUTO = 0, DKGV_SHARE_PATH_HORNER = 1, DKGV_SHARE_PATH_FDIFF = 2 };
int dkgv_set_share_path(dkgv_ctx* ctx, int mode); /* FDIFF: use it whenever applicable, even if not cheaper */
int dkgv_set_share_parts(dkgv_ctx* ctx, uint32_t parts); /* 0 = planner's choice (default), else 1..16 */
/* 1 (default): the parts of the finite-difference path run on one internal stream each and join before the
 * recombination; 0: everything on the caller's stream, phase after phase (gives per-phase device times).
 * (Starting the recombination of an id range while the extension is still running was measured slower
 * at N = 1 and N = 8 and removed.)                                                                     */
int dkgv_set_share_overlap(dkgv_ctx* ctx, int on);
/* Consistency shortcut of the finite-difference path (default on).  All n shares of a dealer are valid exactly when
 * (1) each is < r, (2) they lie on a polynomial of degree <= t-1 over Fr - pure scalar arithmetic: the t-th forward
 * differences of the share sequence vanish - and (3) G * p_k == C_k for each coefficient of the interpolated polynomial p.
 * So no share goes through the group arithmetic; a group of 32 dealers in which some dealer fails a condition continues with the full
 * evaluation, which yields the exact per-share verdicts.  Deterministic and exact (no random linear combination).
 * Applies for t <= 1024 and t < n_recipients <= 2048 (other shapes: every share through the evaluation).
 * dkgv_last_share_continued: 1 when the last call had to continue into the evaluation for some dealer group, 0 when the
 * shortcut settled everything.                                                                                   */
int dkgv_set_share_shortcut(dkgv_ctx* ctx, int on);
int dkgv_last_share_continued(const dkgv_ctx* ctx);
/* Repair route behind the shortcut (default on).  A dealer whose shares are NOT on one polynomial of degree < t is first decoded as
 * a Reed-Solomon word with errors over Fr (syndromes, Berlekamp-Massey, root search, Newton interpolation through the shares not
 * located as wrong - scalar arithmetic only): with at most floor((n - t) / 2) wrong shares this recovers the dealer's polynomial p,
 * which is then put through the SAME exact conditions as an honest dealer's (t-th differences of the corrected sequence,
 * compress(G * p_k) == C_k).  Only a confirmed p yields verdicts - share != p(id) -> SLASHABLE_SHARE_MISMATCH, >= r ->
 * SLASHABLE_SECRET_RANGE, else OK - so they are exact and no randomness is involved; anything the decoder cannot settle goes to the
 * evaluation as before.  dkgv_last_share_repaired: dealers the route settled in the last share-matrix call.                       */
int dkgv_set_share_repair(dkgv_ctx* ctx, int on);
int dkgv_last_share_repaired(const dkgv_ctx* ctx);
/* 1 when the commitments of the last share-matrix call were decoded (square roots, subgroup tests), 0 when the shortcut
 * settled the call against their compressed encodings: compress(G * p_k) == C_k needs no decompression, and an encoding
 * that is not a subgroup point can never agree, so the decode waits until some dealer group needs the evaluation.   */
int dkgv_last_share_decoded(const dkgv_ctx* ctx);
int dkgv_last_share_path(const dkgv_ctx* ctx);    /* HORNER or FDIFF: what the last share-matrix call ran */
/* the plan for ids 1..n_recipients (parts_force 0 = cheapest; n_opt = number of ids the cost model assumes are
 * evaluated in the group, 0 = all, t when the consistency shortcut is on): parts, h = ceil(t / parts), Horner seed points
 * lo..hi (hi - lo + 1 == h), extension steps, field products per dealer by this plan and by per-share Horner
 * (evaluation only, G*s excluded).  Returns 1 when AUTO would take this plan, 0 when Horner per share is
 * cheaper, -1 when no plan exists for the shape.  Pure host function.                                */
int dkgv_share_fd_plan(uint32_t t, uint32_t n_recipients, uint32_t parts_force, uint32_t n_opt, uint32_t* parts, uint32_t* h,
                       int32_t* lo, int32_t* hi, uint32_t* steps, uint64_t* modmul_fd, uint64_t* modmul_horner);
/* device times of the last finite-difference run: seed Horner, differences, extension, recombine + G*s compare;
 * with overlap on the first three run concurrently and only ms4[0] (their total) and ms4[3] are meaningful.
 * When the consistency shortcut settled the call (dkgv_last_share_continued == 0) the four intervals are its own:
 * share limbs + difference table, x halves of G*p_k == C_k, sign halves, flags + verdict fill.                */
int dkgv_last_share_phases_ms(dkgv_ctx* ctx, float* ms4);

/* ---- evaluate_polynomial (crates/dkg/src/dkg_math.rs:160-174), batched ---------------------- */
/* out[d][j] = compress( sum_k vv[d][k] * ids[j]^k ), 48 B each; status per dealer row
 * (OK / PANIC_BAD_G1).                                                                          */
int dkgv_feldman_eval(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t n_ids, uint32_t t, const uint8_t* vv,
                      const uint32_t* ids, uint8_t* out, uint8_t* row_status);

/* ---- G * s (BlsSecretKey::to_public_key, crates/dkg/src/crypto/bls_keys.rs:133-137) --------- */
/* out[i] = compress(G * scalars[i]); status OK / SLASHABLE_SECRET_RANGE (scalar >= r)           */
int dkgv_g1_fixed_base_mul(dkgv_ctx* ctx, uint32_t m, const uint8_t* scalars, uint8_t* out, uint8_t* status);

/* ---- P * s (BlsG1::mul_scalar, crates/dkg/src/dkg_math.rs:122-127; the ECDH step of
 *      crates/bad_encrypted_share_prove/src/main.rs:339-342), m independent pairs ----------------- */
/* out[i] = compress([scalars[i]] pts[i]); status OK / PANIC_BAD_G1 / PANIC_BAD_SCALAR (scalar >= r) */
int dkgv_g1_mul_batch(dkgv_ctx* ctx, uint32_t m, const uint8_t* pts, const uint8_t* scalars, uint8_t* out, uint8_t* status);

/* ---- G1 decoding with subgroup check (to_g1_affine, crypto/bls_common.rs:108-112) ----------- */
int dkgv_g1_decompress_check(dkgv_ctx* ctx, uint32_t m, const uint8_t* in, uint8_t* decode_status);

/* ---- final / partial key aggregation (crates/dkg/src/dkg_math.rs:178-248) -------------------- */
/* agg_coefficients: vv [n][t][48] (t = vv[0].len(); ragged inputs are the caller's PANIC_INDEX),
 * coeff_out [t][48] = column sums C_k (may be NULL), keys_out [n_ids][48] = evaluate_polynomial(C, ids[j]).
 * *status: OK / PANIC_BAD_G1 (some coefficient undecodable; outputs are then meaningless).       */
int dkgv_agg_final_keys(dkgv_ctx* ctx, uint32_t n, uint32_t t, const uint8_t* vv, const uint32_t* ids, uint32_t n_ids,
                        uint8_t* coeff_out, uint8_t* keys_out, uint8_t* status);
/* lagrange_interpolation at 0 over (ids[i], pts[i]); ids are u32 scalars (bls_id_from_u32).
 * *status: OK / ERR_LEN (k == 0) / ERR_ZERO_ID / ERR_DUP_ID / PANIC_BAD_G1; k == 1 returns pts[0]. */
int dkgv_lagrange_at_zero(dkgv_ctx* ctx, uint32_t k, const uint8_t* pts, const uint32_t* ids, uint8_t* out, uint8_t* status);
/* evaluate_polynomial over an arbitrary point vector: out[j] = sum_k coeffs[k] * ids[j]^k          */
int dkgv_eval_points(dkgv_ctx* ctx, uint32_t t, const uint8_t* coeffs, const uint32_t* ids, uint32_t n_ids, uint8_t* out,
                     uint8_t* status);

/* ---- BLS partial-signature checks (crates/dkg/src/crypto/bls_common.rs:11-40) ----------------- */
/* G2 decoding with subgroup check (BlsSignature::from_bytes, crypto/bls_keys.rs:165-176)          */
int dkgv_g2_decompress_check(dkgv_ctx* ctx, uint32_t m, const uint8_t* in, uint8_t* decode_status);
/* hash_message_to_g2 for m messages: msgs = concatenation, offsets [m+1]; out [m][96] compressed  */
int dkgv_hash_to_g2(dkgv_ctx* ctx, uint32_t m, const uint8_t* msgs, const uint32_t* offsets, uint8_t* out);
/* bls_verify_precomputed_hash for m (pk, sig) pairs: e(pk, H) == e(G1, sig), H = hm[hm_idx[i]]
 * (hm_idx NULL: all use hm[0]).  hm [n_hm][96] compressed hashed messages.
 * status[i]: OK / SLASHABLE_SIG_INVALID (equality false) / PANIC_BAD_G2 (sig undecodable; checked
 * first, as verification.rs:238-241) / PANIC_BAD_G1 (pk undecodable).  Callers map the decode
 * outcomes to their call site's exit (panic in verify_generation_hashes, slashable in
 * prove_wrong_final_key_generation, unslashable-invalid in verification.rs:243-248).              */
int dkgv_bls_verify_batch(dkgv_ctx* ctx, uint32_t m, const uint8_t* pk, const uint8_t* sig, uint32_t n_hm, const uint8_t* hm,
                          const uint32_t* hm_idx, uint8_t* status);
int dkgv_bls_verify_batch_dev(dkgv_ctx* ctx, uint32_t m, const uint8_t* d_pk, const uint8_t* d_sig, uint32_t n_hm,
                              const uint8_t* d_hm, const uint32_t* d_hm_idx, uint8_t* d_status, void* stream);

/* Execution of the pairing batch.  AUTO (default): the pairing VM - several warps cooperate on each group of 32 checks with
 * every Fp2 of a check in shared memory (csrc/pairing_vm.cuh) - whenever the prepared Miller-loop lines of all hashed messages fit
 * a 1 GB budget (19.6 KB per distinct message), else one thread per check.  THREAD forces the latter (parity tests: two independent
 * implementations of the same check).  dkgv_last_bls_kernel_ms: device time of the last pairing kernel alone (without decoding). */
enum dkgv_bls_path { DKGV_BLS_PATH_AUTO = 0, DKGV_BLS_PATH_VM = 1, DKGV_BLS_PATH_THREAD = 2 };
int dkgv_set_bls_path(dkgv_ctx* ctx, int mode);
int dkgv_last_bls_path(const dkgv_ctx* ctx);
int dkgv_last_bls_kernel_ms(dkgv_ctx* ctx, float* ms);

/* ---- prove_wrong_final_key_generation (crates/dkg/src/verification.rs:422-466) for m items over ONE session -------------
 * The curve part of the flow, batched: one hash-to-G2 launch over all messages, one decode + pairing batch over all items, one
 * aggregation per session and the "expected key" of every perpetrator index at once (verify_expected_key :399-420 ->
 * compute_pubkey_share :523-551, including its quirk: evaluate_polynomial over the n FINAL KEYS K_j at the perpetrator's id).
 * The caller has done the per-item hash / identity-signature checks and the sort by base_hash (:428-438).
 *   vv [n][t][48]  verification vectors of the n generations in base_hash-sorted order (ids 1..n; t = len of vv[0])
 *   perp [m]       index of the accused generation in that order (find_perpetrator_index :498-521)
 *   pk [m][48], sig [m][96], msgs/msg_offsets [n_msg + 1]/msg_idx [m] (NULL: message 0)   partial_pubkey, message_signature,
 *                  message_cleartext of bad_partial.data
 *   status [m]     SLASHABLE_BAD_PK / SLASHABLE_BAD_SIG / SLASHABLE_SIG_INVALID / SLASHABLE_KEY_MISMATCH / OK in the reference's
 *                  order of checks; PANIC_BAD_G1 for an item that reaches verify_expected_key of a session with an undecodable
 *                  commitment (the reference's .expect, verification.rs:534)
 *   expected_out [n][48] (may be NULL)  the expected key per perpetrator index - with pk[i] the (expected, got) pair of the
 *                  reference's message (verification.rs:414-417)
 *   *session_status  OK / PANIC_BAD_G1                                                                                  */
int dkgv_bad_partial_key_verify_batch(dkgv_ctx* ctx, uint32_t n, uint32_t t, const uint8_t* vv, uint32_t m, const uint32_t* perp,
                                      const uint8_t* pk, const uint8_t* sig, uint32_t n_msg, const uint8_t* msgs,
                                      const uint32_t* msg_offsets, const uint32_t* msg_idx, uint8_t* status, uint8_t* expected_out,
                                      uint8_t* session_status);

/* ---- multi-GPU: one process (and one ctx) per GPU, the collectives INSIDE the library (NCCL over NVLink / NVSwitch) -----------
 * Rank 0 asks for a 128-byte unique id and hands it to the other ranks (any transport); every rank then calls dkgv_comm_init on
 * its ctx.  NCCL is bound at run time (libnccl.so.2; DKGV_NCCL_LIB overrides), the library links against nothing but the CUDA
 * runtime.  A ctx without a communicator behaves as world = 1 in every *_sharded call.  dkgv_ctx_destroy releases the comm.   */
int dkgv_comm_unique_id(uint8_t id_out[128]);
int dkgv_comm_init(dkgv_ctx* ctx, const uint8_t id[128], int rank, int world);
int dkgv_comm_destroy(dkgv_ctx* ctx);
int dkgv_comm_world(const dkgv_ctx* ctx);
int dkgv_comm_rank(const dkgv_ctx* ctx);
/* rank r's `bytes` land at d_recv + r * bytes on every rank (world 1: a copy); asynchronous on `stream` */
int dkgv_all_gather_dev(dkgv_ctx* ctx, const void* d_send, void* d_recv, size_t bytes, void* stream);
/* Share matrix of one ceremony, sharded by dealer row blocks: every rank passes ITS n_local dealers' verification
 * vectors and shares (the same n_local on every rank) and the full id list; no exchange during compute.  d_gather [world][chunk]
 * u32, chunk = dkgv_share_gather_words(n_local, n_recipients): after the call, on EVERY rank, rank r's chunk holds the verdict
 * bitmask of its row block (bit i % 32 of word i / 32 set = share i of the block is NOT ok) and, behind the bitmask, its two job
 * flags.  Honest ceremony: asynchronous submit + pack + ONE all-gather + one synchronisation; otherwise the ranks that have
 * unsettled dealers evaluate them and a second all-gather follows (every rank sees every flag, so all agree on that).
 * d_status_local [n_local][n_recipients] keeps the status bytes (outcome classes) of the own rows.                          */
uint32_t dkgv_share_gather_words(uint32_t n_local, uint32_t n_recipients);
int dkgv_share_matrix_verify_sharded_dev(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_recipients, uint32_t t, const uint8_t* d_vv_local,
                                         const uint32_t* d_ids, const uint8_t* d_shares_local, uint8_t* d_status_local, uint32_t* d_gather,
                                         void* stream);
/* Pipelined form for a host with many ceremonies in flight.  enqueue: the shortcut, the pack, the ONE all-gather and the copy of every
 * rank's two flag words to h_flags (PINNED host memory, 2 * world words) are queued on `stream`; no synchronisation, CUDA-graph
 * capturable; the ctx's scratch is reused in stream order, so ceremonies may be queued back to back.  settle (after the caller has
 * synchronised): flags all zero - an honest ceremony, verdicts final, nothing is launched; otherwise the ceremony goes through
 * dkgv_share_matrix_verify_sharded_dev with the same arguments (all ranks see all flags and take the same branch); *reran = 0 / 1. */
int dkgv_share_matrix_enqueue_sharded_dev(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_recipients, uint32_t t, const uint8_t* d_vv_local,
                                          const uint32_t* d_ids, const uint8_t* d_shares_local, uint8_t* d_status_local, uint32_t* d_gather,
                                          uint32_t* h_flags, void* stream);
int dkgv_share_matrix_settle_sharded_dev(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_recipients, uint32_t t, const uint8_t* d_vv_local,
                                         const uint32_t* d_ids, const uint8_t* d_shares_local, uint8_t* d_status_local, uint32_t* d_gather,
                                         const uint32_t* h_flags, void* stream, int* reran);
/* The same pair with HOST buffers - the reference-facing form for a host with many ceremonies (verification.rs:68-149 per ceremony):
 * rows, ids and shares in (PINNED memory, or the copies are not asynchronous), status_local [n_local][n_recipients], gather
 * [world][chunk] (may be NULL) and h_flags (2 * world words) out; everything queued on the ctx's own stream, nothing synchronised.
 * One ctx per ceremony in flight on a GPU (ctxs of a device share the fixed-base table): the copies of one ceremony run under the
 * kernels of another.  dkgv_sync(ctx), then settle with the same arguments (synchronous when it has to run the ceremony again). */
int dkgv_share_matrix_enqueue_sharded(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_recipients, uint32_t t, const uint8_t* vv_local,
                                      const uint32_t* ids, const uint8_t* shares_local, uint8_t* status_local, uint32_t* gather,
                                      uint32_t* h_flags);
int dkgv_share_matrix_settle_sharded(dkgv_ctx* ctx, uint32_t n_local, uint32_t n_recipients, uint32_t t, const uint8_t* vv_local,
                                     const uint32_t* ids, const uint8_t* shares_local, uint8_t* status_local, uint32_t* gather,
                                     const uint32_t* h_flags, int* reran);
/* Pairing checks sharded by items: every rank its m_local pairs; d_status_all [world][m_local] on every rank.  Asynchronous. */
int dkgv_bls_verify_batch_sharded_dev(dkgv_ctx* ctx, uint32_t m_local, const uint8_t* d_pk, const uint8_t* d_sig, uint32_t n_hm,
                                      const uint8_t* d_hm, const uint32_t* d_hm_idx, uint8_t* d_status_all, void* stream);
/* agg_coefficients sharded by generations (dkg_math.rs:230-248): every rank decodes and sums ITS n_local rows, one all-gather of
 * the t projective partial sums (144 B each), every rank adds the partials and evaluates the keys.  Outputs on every rank.    */
int dkgv_agg_final_keys_sharded(dkgv_ctx* ctx, uint32_t n_local, uint32_t t, const uint8_t* vv_local, const uint32_t* ids, uint32_t n_ids,
                                uint8_t* coeff_out, uint8_t* keys_out, uint8_t* status);

/* ---- initial-commitment hashes (crates/dkg/src/verification.rs:151-175) on the GPU ------------- */
/* out[d] = SHA-256(gen_id(16) || n || k || (t as u8) || vv[d][0..t)), one per dealer, out [n_dealers][32] */
int dkgv_initial_commitment_hashes(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t t, const uint8_t* vv, const uint8_t* gen_id16,
                                   uint8_t n, uint8_t k, uint8_t* out);

/* ---- dealer-side helper for building synthetic ceremonies (not a verification step) --------- */
/* out[d][j] = sum_k coeffs[d][k] * ids[j]^k mod r ; coeffs [n_dealers][t][32] BE (< r), out BE   */
int dkgv_fr_poly_eval(dkgv_ctx* ctx, uint32_t n_dealers, uint32_t t, const uint8_t* coeffs, uint32_t n_ids,
                      const uint32_t* ids, uint8_t* out);
/* signer-side helper: out[i] = compress([scalars[i]] * base), base a compressed G2 point (e.g. H(m)) */
int dkgv_g2_mul_batch(dkgv_ctx* ctx, uint32_t m, const uint8_t* base96, const uint8_t* scalars, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* DKGV_H */
