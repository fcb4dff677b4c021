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
 *   auth = cargo feature auth_commitment; bls_identity = BlsDkgWithBlsCommitment instead of secp256k1.
 * Returns the reference's process exit code (0 = misbehaviour proven / ceremony valid, 1 otherwise);
 * *status = dkgv_status reached, or 255 when the input is rejected while parsing (serde error).   */
int dkgh_execute(dkgv_ctx* ctx, const char* type, const char* json_text, int auth, int bls_identity, int* status, char* msg,
                 size_t msg_cap);

/* compute_initial_commitment_hash (crates/dkg/src/verification.rs:151-175) */
void dkgh_initial_commitment_hash(const uint8_t* gen_id16, uint8_t n, uint8_t k, const uint8_t* base_pubkeys, uint32_t count,
                                  uint8_t* out32);

#ifdef __cplusplus
}
#endif
#endif /* DKGH_H */
