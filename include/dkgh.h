/* dkgh.h - host-level entry points layered on dkgv.h: the reference's `dkg` crate flows and the
 * `dkg_prover_host execute` contract (src/main.rs:306-345, script/run.sh:78-82), implemented in
 * dvt_circuits_b200/host/dkg_host.cpp.  All curve / pairing arithmetic goes through dkgv_* (GPU). */
#ifndef DKGH_H
#define DKGH_H
#include <stddef.h>
#include <stdint.h>

#include "dkgv.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Runs one dkg_prover_host input (JSON text, crates/dkg/src/types.rs:27-203) through the guest logic
 *   type = "bad-share"        crates/bad_share_exchange_prove/src/main.rs:16-82
 *        | "finalization"     crates/finalization_prove/src/main.rs:7-33   (BLS identity setup, as the reference)
 *        | "bad-partial-key"  crates/bad_parial_key_prove/src/main.rs:16-51
 *        | "bad-encrypted-share" crates/bad_encrypted_share_prove/src/main.rs:281-405 (incl. quirk Q2)
 *        | "fn:verify_seed_exchange_commitment" | "fn:verify_generations" | "fn:prove_wrong_final_key_generation"
 *                             the crate functions of crates/dkg/src/lib.rs:6-9 ALONE, without a guest's pre-checks and outcome
 *                             mapping (inputs: SharedData / FinalizationData / BadPartialShareData); return value 0 = Ok(()), 1 = Err
 *   auth = cargo feature auth_commitment; bls_identity = BlsDkgWithBlsCommitment instead of secp256k1.
 * Returns the reference's process exit code (0 = misbehaviour proven / ceremony valid, 1 otherwise);
 * *status = dkgv_status reached, or 255 when the input is rejected while parsing (serde error).   */
int dkgh_execute(dkgv_ctx* ctx, const char* type, const char* json_text, int auth, int bls_identity, int* status, char* msg,
                 size_t msg_cap);

/* The same run, also returning what the reference's guest hands to the outside world:
 *  - the public values it commits with sp1_zkvm::io::commit, in commit order (bad_share_exchange_prove/src/main.rs:57-71: every
 *    verification hash, then the perpetrator's identity key; finalization_prove/src/main.rs:26-32: every generation's base_hash
 *    in input order, then the aggregate key; bad_parial_key_prove/src/main.rs:31-41; bad_encrypted_share_prove/src/main.rs:359-369).
 *    Only runs that exit 0 commit anything.  public_values (caller-allocated, public_cap bytes) receives n_public entries of
 *    [u32 little-endian length][bytes]; public_len is the size needed (nothing is written when it exceeds public_cap).
 *  - the two keys the reference prints in its message for a share mismatch (verification.rs:141-145: expected = the Feldman
 *    evaluation, got = G * s), an aggregate-key mismatch (:304-307, :323-326: expected = the claimed aggregate key, got = the
 *    computed one) and a partial-key mismatch (:414-417: expected = compute_pubkey_share, got = the accused key).            */
typedef struct dkgh_report {
  uint8_t* public_values;
  size_t public_cap, public_len;
  uint32_t n_public;
  int have_keys;
  uint8_t expected[48], got[48];
} dkgh_report;
int dkgh_execute_report(dkgv_ctx* ctx, const char* type, const char* json_text, int auth, int bls_identity, int* status, char* msg,
                        size_t msg_cap, dkgh_report* report);

/* compute_initial_commitment_hash (crates/dkg/src/verification.rs:151-175) */
void dkgh_initial_commitment_hash(const uint8_t* gen_id16, uint8_t n, uint8_t k, const uint8_t* base_pubkeys, uint32_t count,
                                  uint8_t* out32);

#ifdef __cplusplus
}
#endif
#endif /* DKGH_H */
