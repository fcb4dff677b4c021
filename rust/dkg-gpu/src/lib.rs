//! `dkg-gpu`: the checks of `crates/dkg` executed on a B200 through `libdkgv.so`.
//!
//! The three flow functions and the two hash helpers have EXACTLY the signatures of `crates/dkg/src/lib.rs:6-12`, so a caller
//! switches by changing `use dkg::{...}` into `use dkg_gpu::{...}`; a single item is a batch of one.  They run on a lazily
//! created per-thread context (`Gpu`); the `*_on` variants take an explicit one, the `*_batch` functions are the entry points a
//! service uses (whole share matrices, many pairing checks, many bad-partial-key items in one call).
//!
//! Inputs cross the C ABI as the JSON the reference's own `dkg_prover_host` takes (the serde structs of `crates/dkg/src/types.rs`),
//! or as dense byte arrays in the reference's wire formats (G1 48 B, G2 96 B, scalars 32 B big-endian).  Every distinct exit of the
//! reference comes back as one `dkgv_status` code and is mapped to the same `Result` / panic here.
//!
//! NOTE: written against the reference's public API; this repository's image has no Rust toolchain, so the crate is
//! source-only here (see INTEGRATION.md for how it is wired into the reference's workspace).
use std::cell::RefCell;
use std::error::Error;
use std::ffi::{CStr, CString};
use std::os::raw::c_int;

use dkg::{
    BadPartialShareData, ByteConvertible, DkgSetup, DkgSetupTypes, GenerateSettings, Generation, InitialCommitment, RawBytes,
    SeedExchangeCommitment, VerificationErrors, VerificationHashes, AsByteArr, SHA256Raw,
};
use dkg_cuda_sys as sys;
use serde::Serialize;

/// One `dkgv_ctx`: bound to one GPU, single owner.
pub struct Gpu {
    ctx: *mut sys::dkgv_ctx,
}
// the C side keeps no thread affinity; a ctx must simply not be used from two threads at once
unsafe impl Send for Gpu {}

#[derive(Debug)]
pub struct GpuError(pub String);
impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "dkgv: {}", self.0)
    }
}
impl Error for GpuError {}

impl Gpu {
    /// Binds to CUDA device `device`.  There is no CPU fallback: without a GPU this fails.
    pub fn new(device: i32) -> Result<Gpu, GpuError> {
        let mut ctx: *mut sys::dkgv_ctx = std::ptr::null_mut();
        let rc = unsafe { sys::dkgv_ctx_create(device as c_int, &mut ctx) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::dkgv_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(GpuError(format!("dkgv_ctx_create({device}) failed ({rc}): {msg}")));
        }
        Ok(Gpu { ctx })
    }

    /// Same with the window width of the fixed-base table chosen by the caller (`dkgv_ctx_create_ex`): 16 = 50 MB of device memory and
    /// 15 mixed additions per `G * s`, 22 (the default of [`Gpu::new`]) = 2.4 GB / 11, 26 = 32 GB / 9.
    pub fn with_table(device: i32, gtab_bits: u32) -> Result<Gpu, GpuError> {
        let mut ctx: *mut sys::dkgv_ctx = std::ptr::null_mut();
        let rc = unsafe { sys::dkgv_ctx_create_ex(device as c_int, gtab_bits, &mut ctx) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::dkgv_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(GpuError(format!("dkgv_ctx_create_ex({device}, {gtab_bits}) failed ({rc}): {msg}")));
        }
        Ok(Gpu { ctx })
    }

    /// This rank's row block of dealers through the default share path, then ONE all-gather inside the library: `d_gather` receives the
    /// verdict bitmask of the whole ceremony (world x ceil(n_local * n_recipients / 32) words + the job flags).
    ///
    /// # Safety
    /// All pointers are DEVICE pointers of this `Gpu`'s device, sized as `include/dkgv.h` states; `stream` is a CUDA stream handle or null.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn share_matrix_verify_sharded_dev(&mut self, n_local: u32, n_recipients: u32, t: u32, d_vv_local: *const u8, d_ids: *const u32,
        d_shares_local: *const u8, d_status_local: *mut u8, d_gather: *mut u32, stream: *mut std::os::raw::c_void) -> Result<(), GpuError> {
        self.check(sys::dkgv_share_matrix_verify_sharded_dev(self.ctx, n_local, n_recipients, t, d_vv_local, d_ids, d_shares_local,
            d_status_local, d_gather, stream))
    }

    /// Pipelined form: queues the ceremony (shortcut, pack, all-gather, flag copy to `h_flags` = pinned host memory of 2 x world words)
    /// without synchronising; any number may be in flight on one stream.
    ///
    /// # Safety
    /// As `share_matrix_verify_sharded_dev`; `h_flags` must stay valid (and pinned) until the stream has been synchronised.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn share_matrix_enqueue_sharded_dev(&mut self, n_local: u32, n_recipients: u32, t: u32, d_vv_local: *const u8, d_ids: *const u32,
        d_shares_local: *const u8, d_status_local: *mut u8, d_gather: *mut u32, h_flags: *mut u32, stream: *mut std::os::raw::c_void) -> Result<(), GpuError> {
        self.check(sys::dkgv_share_matrix_enqueue_sharded_dev(self.ctx, n_local, n_recipients, t, d_vv_local, d_ids, d_shares_local,
            d_status_local, d_gather, h_flags, stream))
    }

    /// After the stream has been synchronised: `Ok(false)` - the flags of every rank were zero, the verdicts are final; `Ok(true)` - the
    /// ceremony had corrupted shares or foreign ids and was run again through the synchronous entry point (same on every rank).
    ///
    /// # Safety
    /// As `share_matrix_enqueue_sharded_dev`, same arguments.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn share_matrix_settle_sharded_dev(&mut self, n_local: u32, n_recipients: u32, t: u32, d_vv_local: *const u8, d_ids: *const u32,
        d_shares_local: *const u8, d_status_local: *mut u8, d_gather: *mut u32, h_flags: *const u32, stream: *mut std::os::raw::c_void) -> Result<bool, GpuError> {
        let mut reran: c_int = 0;
        self.check(sys::dkgv_share_matrix_settle_sharded_dev(self.ctx, n_local, n_recipients, t, d_vv_local, d_ids, d_shares_local,
            d_status_local, d_gather, h_flags, stream, &mut reran))?;
        Ok(reran != 0)
    }

    /// Host-buffer form of the pipelined pair: this rank's rows, the ids and the shares are copied in, the status bytes of the own rows,
    /// the gathered chunks (`gather` may be null) and every rank's flag words come back - all queued on the ctx's own stream, nothing
    /// synchronised.  Call `sync`, then `share_matrix_settle_sharded` with the same arguments.
    ///
    /// # Safety
    /// Every pointer is HOST memory sized as `include/dkgv.h` states, pinned (page-locked) for the copies to be asynchronous, and must
    /// stay valid until `sync` has returned.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn share_matrix_enqueue_sharded(&mut self, n_local: u32, n_recipients: u32, t: u32, vv_local: *const u8, ids: *const u32,
        shares_local: *const u8, status_local: *mut u8, gather: *mut u32, h_flags: *mut u32) -> Result<(), GpuError> {
        self.check(sys::dkgv_share_matrix_enqueue_sharded(self.ctx, n_local, n_recipients, t, vv_local, ids, shares_local, status_local, gather, h_flags))
    }

    /// `Ok(false)`: honest ceremony, the verdicts `enqueue` delivered are final; `Ok(true)`: it was run again (synchronously).
    ///
    /// # Safety
    /// As `share_matrix_enqueue_sharded`, same arguments, after `sync`.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn share_matrix_settle_sharded(&mut self, n_local: u32, n_recipients: u32, t: u32, vv_local: *const u8, ids: *const u32,
        shares_local: *const u8, status_local: *mut u8, gather: *mut u32, h_flags: *const u32) -> Result<bool, GpuError> {
        let mut reran: c_int = 0;
        self.check(sys::dkgv_share_matrix_settle_sharded(self.ctx, n_local, n_recipients, t, vv_local, ids, shares_local, status_local, gather,
            h_flags, &mut reran))?;
        Ok(reran != 0)
    }

    /// Waits for everything queued on the ctx's own stream.
    pub fn sync(&mut self) -> Result<(), GpuError> {
        self.check(unsafe { sys::dkgv_sync(self.ctx) })
    }

    fn check(&self, rc: c_int) -> Result<(), GpuError> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(sys::dkgv_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(GpuError(format!("call failed ({rc}): {msg}")))
    }

    /// Joins a multi-GPU job: one process (and one `Gpu`) per device, the collectives live inside the library.
    /// Rank 0 obtains `id` from [`comm_unique_id`] and hands the 128 bytes to the other ranks.
    pub fn comm_init(&mut self, id: &[u8; 128], rank: i32, world: i32) -> Result<(), GpuError> {
        self.check(unsafe { sys::dkgv_comm_init(self.ctx, id.as_ptr(), rank as c_int, world as c_int) })
    }

    /// Runs one input through a flow; returns (status, message, report).
    fn execute(&mut self, ty: &str, json: &str, bls_identity: bool) -> Result<Outcome, GpuError> {
        let ty = CString::new(ty).unwrap();
        let json = CString::new(json).map_err(|_| GpuError("input contains a NUL byte".into()))?;
        let mut status: c_int = 255;
        let mut msg = vec![0u8; 512];
        let mut public = vec![0u8; 1 << 20];
        let mut rep = sys::dkgh_report {
            public_values: public.as_mut_ptr(),
            public_cap: public.len(),
            public_len: 0,
            n_public: 0,
            have_keys: 0,
            expected: [0u8; 48],
            got: [0u8; 48],
        };
        let exit_code = unsafe {
            sys::dkgh_execute_report(
                self.ctx,
                ty.as_ptr(),
                json.as_ptr(),
                cfg!(feature = "auth_commitment") as c_int,
                bls_identity as c_int,
                &mut status,
                msg.as_mut_ptr() as *mut _,
                msg.len(),
                &mut rep,
            )
        };
        let text = CStr::from_bytes_until_nul(&msg).map(|c| c.to_string_lossy().into_owned()).unwrap_or_default();
        let mut values = Vec::new();
        if rep.public_len <= public.len() {
            let mut o = 0usize;
            for _ in 0..rep.n_public {
                let len = u32::from_le_bytes([public[o], public[o + 1], public[o + 2], public[o + 3]]) as usize;
                values.push(public[o + 4..o + 4 + len].to_vec());
                o += 4 + len;
            }
        }
        Ok(Outcome {
            exit_code,
            status: status as u8,
            message: text,
            public_values: values,
            keys: if rep.have_keys != 0 { Some((rep.expected, rep.got)) } else { None },
        })
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { sys::dkgv_ctx_destroy(self.ctx) }
    }
}

/// 128 bytes for [`Gpu::comm_init`] (rank 0 of a multi-GPU job).
pub fn comm_unique_id() -> Result<[u8; 128], GpuError> {
    let mut id = [0u8; 128];
    if unsafe { sys::dkgv_comm_unique_id(id.as_mut_ptr()) } != 0 {
        return Err(GpuError("NCCL is not loadable (libnccl.so.2; DKGV_NCCL_LIB overrides)".into()));
    }
    Ok(id)
}

/// What a run hands back: the reference's process exit code, the `dkgv_status` reached, the guest's committed public values
/// (in commit order) and the (expected, got) keys of the reference's error message where it prints them.
pub struct Outcome {
    pub exit_code: i32,
    pub status: u8,
    pub message: String,
    pub public_values: Vec<Vec<u8>>,
    pub keys: Option<([u8; 48], [u8; 48])>,
}

thread_local! {
    static DEFAULT_GPU: RefCell<Option<Gpu>> = RefCell::new(None);
}

/// Runs `f` on this thread's default context (device 0, created on first use).
pub fn with_default_gpu<R>(f: impl FnOnce(&mut Gpu) -> R) -> Result<R, GpuError> {
    DEFAULT_GPU.with(|cell| {
        let mut slot = cell.borrow_mut();
        if slot.is_none() {
            *slot = Some(Gpu::new(0)?);
        }
        Ok(f(slot.as_mut().unwrap()))
    })
}

fn identity_is_bls<Setup: DkgSetup + DkgSetupTypes<Setup>>() -> bool {
    // BlsDkgWithBlsCommitment: 48-byte identity keys; BlsDkgWithSecp256kCommitment: 33 bytes (crates/dkg/src/types.rs:9-25)
    std::mem::size_of::<RawBytes<Setup::CommitmentPubkey>>() == 48
}

/// `dkgv_status` -> the reference's outcome: Ok / SlashableError / UnslashableError / io::Error(InvalidData) / panic.
fn outcome_to_result(o: Outcome) -> Result<(), Box<dyn Error>> {
    let detail = match (&o.keys, o.status) {
        (Some((e, g)), sys::DKGV_SLASHABLE_SHARE_MISMATCH) => {
            format!("Bad secret field : Expected secret with public key: {}, got public key: {}\n", hex::encode(e), hex::encode(g))
        }
        (Some((e, g)), sys::DKGV_SLASHABLE_KEY_MISMATCH) => format!("Computed key {} does not match expected key {}", hex::encode(e), hex::encode(g)),
        (Some((e, g)), sys::DKGV_ERR_AGG_MISMATCH_VV) | (Some((e, g)), sys::DKGV_ERR_AGG_MISMATCH_PK) => {
            format!("Computed key {} does not match aggregate public key {}", hex::encode(g), hex::encode(e))
        }
        _ => format!("dkgv status {}{}{}", o.status, if o.message.is_empty() { "" } else { ": " }, o.message),
    };
    match o.status {
        0 => Ok(()),
        1..=15 => Err(Box::new(VerificationErrors::SlashableError(detail))),
        16..=31 => Err(Box::new(VerificationErrors::UnslashableError(detail))),
        32..=47 | 255 => Err(Box::new(std::io::Error::new(std::io::ErrorKind::InvalidData, detail))),
        // the reference `.expect(...)`s / `panic!`s on undecodable points and scalars (verification.rs:136,238,241,288,314; dkg_math.rs:84,235)
        _ => panic!("{detail}"),
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// crates/dkg/src/lib.rs:6-9 - same names, same signatures
// ---------------------------------------------------------------------------------------------------------------------------

/// `crates/dkg/src/verification.rs:68-149` on the GPU (Feldman evaluation + `G * s` + comparison; hashing, sorting and the
/// identity signature on the host side of the library, in the reference's order of checks).
pub fn verify_seed_exchange_commitment<Setup>(
    verification_hashes: &VerificationHashes,
    seed_exchange: &SeedExchangeCommitment<Setup>,
    initial_commitment: &InitialCommitment<Setup>,
) -> Result<(), Box<dyn Error>>
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
    SeedExchangeCommitment<Setup>: Serialize,
    InitialCommitment<Setup>: Serialize,
{
    with_default_gpu(|gpu| verify_seed_exchange_commitment_on::<Setup>(gpu, verification_hashes, seed_exchange, initial_commitment))?
}

pub fn verify_seed_exchange_commitment_on<Setup>(
    gpu: &mut Gpu,
    verification_hashes: &VerificationHashes,
    seed_exchange: &SeedExchangeCommitment<Setup>,
    initial_commitment: &InitialCommitment<Setup>,
) -> Result<(), Box<dyn Error>>
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
    SeedExchangeCommitment<Setup>: Serialize,
    InitialCommitment<Setup>: Serialize,
{
    let json = serde_json::json!({
        "base_hashes": verification_hashes,
        "initial_commitment": initial_commitment,
        "seeds_exchange_commitment": seed_exchange,
    })
    .to_string();
    outcome_to_result(gpu.execute("fn:verify_seed_exchange_commitment", &json, identity_is_bls::<Setup>())?)
}

/// `crates/dkg/src/verification.rs:262-331`: one hash-to-G2, all partial-signature checks in one pairing batch, aggregation of
/// the verification vectors, two Lagrange interpolations at 0 against the claimed aggregate key.
pub fn verify_generations<Setup>(
    generations: &[Generation<Setup>],
    settings: &GenerateSettings,
    agg_key: &Setup::DkgPubkey,
) -> Result<(), Box<dyn Error>>
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
    Generation<Setup>: Serialize,
{
    with_default_gpu(|gpu| verify_generations_on::<Setup>(gpu, generations, settings, agg_key))?
}

pub fn verify_generations_on<Setup>(
    gpu: &mut Gpu,
    generations: &[Generation<Setup>],
    settings: &GenerateSettings,
    agg_key: &Setup::DkgPubkey,
) -> Result<(), Box<dyn Error>>
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
    Generation<Setup>: Serialize,
{
    let json = serde_json::json!({
        "settings": settings,
        "generations": generations,
        "aggregate_pubkey": hex::encode(agg_key.to_bytes().as_arr()),
    })
    .to_string();
    // the target cryptography is BLS in both setups; finalization always uses BLS identity keys (finalization_prove/src/main.rs:8)
    outcome_to_result(gpu.execute("fn:verify_generations", &json, true)?)
}

/// `crates/dkg/src/verification.rs:422-466`, including `compute_pubkey_share`'s evaluation over the final keys.
pub fn prove_wrong_final_key_generation<Setup>(data: &BadPartialShareData<Setup>) -> Result<(), Box<dyn Error>>
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
    BadPartialShareData<Setup>: Serialize,
{
    with_default_gpu(|gpu| prove_wrong_final_key_generation_on::<Setup>(gpu, data))?
}

pub fn prove_wrong_final_key_generation_on<Setup>(gpu: &mut Gpu, data: &BadPartialShareData<Setup>) -> Result<(), Box<dyn Error>>
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
    BadPartialShareData<Setup>: Serialize,
{
    let json = serde_json::to_string(data)?;
    outcome_to_result(gpu.execute("fn:prove_wrong_final_key_generation", &json, identity_is_bls::<Setup>())?)
}

/// `crates/dkg/src/verification.rs:151-175`: SHA-256(gen_id || n || k || len as u8 || base_pubkeys).
pub fn compute_initial_commitment_hash<Setup>(commitment: &InitialCommitment<Setup>) -> SHA256Raw
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
{
    let flat: Vec<u8> = commitment.base_pubkeys.iter().flat_map(|p| p.as_arr().to_vec()).collect();
    let mut out = [0u8; 32];
    unsafe {
        sys::dkgh_initial_commitment_hash(
            commitment.settings.gen_id.as_ref().as_ptr(),
            commitment.settings.n,
            commitment.settings.k,
            flat.as_ptr(),
            commitment.base_pubkeys.len() as u32,
            out.as_mut_ptr(),
        )
    };
    SHA256Raw::try_from(&out[..]).ok().expect("32-byte digest")
}

/// `crates/dkg/src/verification.rs:177-183`
pub fn verify_initial_commitment_hash<Setup>(commitment: &InitialCommitment<Setup>) -> bool
where
    Setup: DkgSetup + DkgSetupTypes<Setup>,
{
    compute_initial_commitment_hash::<Setup>(commitment).as_ref() == commitment.hash.as_ref()
}

// ---------------------------------------------------------------------------------------------------------------------------
// the guests, natively: `dkg_prover_host execute --type ...` semantics (exit code, status, public values)
// ---------------------------------------------------------------------------------------------------------------------------

/// `type_`: "bad-share" | "finalization" | "bad-partial-key" | "bad-encrypted-share"; `json`: the `dkg_prover_host` input file.
pub fn execute(gpu: &mut Gpu, type_: &str, json: &str, bls_identity: bool) -> Result<Outcome, GpuError> {
    gpu.execute(type_, json, bls_identity)
}

// ---------------------------------------------------------------------------------------------------------------------------
// batch entry points (dense byte arrays in the reference's wire formats)
// ---------------------------------------------------------------------------------------------------------------------------

/// Whole (dealer x recipient) share matrix of one ceremony: `vv` [n_dealers][t][48], `ids` [n_recipients] (rank of the recipient's
/// commitment hash + 1), `shares` [n_dealers][n_recipients][32] -> status byte per share (0 = valid).
pub fn verify_share_matrix(gpu: &mut Gpu, n_dealers: u32, t: u32, vv: &[u8], ids: &[u32], shares: &[u8]) -> Result<Vec<u8>, GpuError> {
    let n_r = ids.len() as u32;
    assert_eq!(vv.len(), n_dealers as usize * t as usize * 48);
    assert_eq!(shares.len(), n_dealers as usize * n_r as usize * 32);
    let mut status = vec![0u8; n_dealers as usize * n_r as usize];
    gpu.check(unsafe { sys::dkgv_share_matrix_verify(gpu.ctx, n_dealers, n_r, t, vv.as_ptr(), ids.as_ptr(), shares.as_ptr(), status.as_mut_ptr()) })?;
    Ok(status)
}

/// `bls_verify_precomputed_hash` for many (key, signature) pairs against `hm` (compressed hashed messages, `hm_idx` picks one per
/// pair; `None`: all use the first) -> status per pair: 0 valid, 7 invalid, 48 / 49 undecodable key / signature.
pub fn bls_verify_batch(gpu: &mut Gpu, pk: &[u8], sig: &[u8], hm: &[u8], hm_idx: Option<&[u32]>) -> Result<Vec<u8>, GpuError> {
    let m = (pk.len() / 48) as u32;
    assert_eq!(sig.len(), m as usize * 96);
    let mut status = vec![0u8; m as usize];
    gpu.check(unsafe {
        sys::dkgv_bls_verify_batch(
            gpu.ctx,
            m,
            pk.as_ptr(),
            sig.as_ptr(),
            (hm.len() / 96) as u32,
            hm.as_ptr(),
            hm_idx.map_or(std::ptr::null(), |i| i.as_ptr()),
            status.as_mut_ptr(),
        )
    })?;
    Ok(status)
}

/// Many bad-partial-key items over one session (generations in base_hash-sorted order): status per item in the reference's order
/// of checks (5 bad key, 6 bad signature, 7 signature invalid, 8 key mismatch, 0 nothing to slash) + the expected key per
/// perpetrator index.
pub fn bad_partial_key_verify_batch(
    gpu: &mut Gpu,
    n: u32,
    t: u32,
    vv_sorted: &[u8],
    perpetrator: &[u32],
    pk: &[u8],
    sig: &[u8],
    message: &[u8],
) -> Result<(Vec<u8>, Vec<u8>, u8), GpuError> {
    let m = perpetrator.len() as u32;
    let offsets = [0u32, message.len() as u32];
    let mut status = vec![0u8; m as usize];
    let mut expected = vec![0u8; n as usize * 48];
    let mut session_status = 0u8;
    gpu.check(unsafe {
        sys::dkgv_bad_partial_key_verify_batch(
            gpu.ctx,
            n,
            t,
            vv_sorted.as_ptr(),
            m,
            perpetrator.as_ptr(),
            pk.as_ptr(),
            sig.as_ptr(),
            1,
            message.as_ptr(),
            offsets.as_ptr(),
            std::ptr::null(),
            status.as_mut_ptr(),
            expected.as_mut_ptr(),
            &mut session_status,
        )
    })?;
    Ok((status, expected, session_status))
}
