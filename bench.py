#!/usr/bin/env python
"""bench.py - verified shares/s on the synthetic DKG ceremony n=1024, t=683 and BLS pairing checks/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole share matrix.  With N ranks the ONE ceremony is sharded by dealer row blocks
(strong scaling, as BASELINE.json names it) through the LIBRARY's own multi-GPU entry point `dkgv_share_matrix_verify_sharded_dev`:
local rows, verdict bitmask packed, one NCCL all-gather inside the library (dkgv_comm_*), all inside the timed region.
torch.distributed is only the plumbing that hands the 128-byte communicator id round and takes the max over ranks of the timings.

`value`      device-timed, inputs resident in HBM, max over ranks: the library's DEFAULT path on the synthetic (honest) ceremony - the
             exact consistency shortcut: no share is evaluated in the exponent, no commitment decompressed.
`e2e`        the same metric with HOST (pinned) buffers: H2D of vv + shares + ids and D2H of the verdicts inside the timed region.
`roofline`   integer-pipe roofline of the dominant kernel of the timed steps (k_fd_coefpoint), all shortcut kernels and the
             evaluation kernels, against the IMAD.WIDE peak measured live by bench/imad_peak.
`corruption` the same matrix with 1 share / 1 dealer / 1 % / 10 % / 50 % (BASELINE config 5) of the shares corrupted: verdicts must
             flag exactly the corrupted shares; `full_evaluation`: shortcut switched off.
`pairing`    second BASELINE metric: 262 144 bls_verify_precomputed_hash checks (every 7th with a wrong signature), sharded over the
             ranks; device-timed value, kernel-only roofline (executed products from the pairing VM program), e2e from host buffers,
             cpu_baseline (oracle, two full pairings per check as the reference does).
`bad_partial_key`  BASELINE config 5, second half: 1 M bad-partial-key items over a (64, 43) session, half of them corrupted.
`config_a`   BASELINE config 2: n=64, t=43 full matrix on one GPU, with the CPU full-matrix baseline.
`finalization`  BASELINE config 4 at (n, t): sharded aggregation + Lagrange + the n partial-signature checks; CPU baseline on a
             bounded (64, 43) ceremony.
`cpu_baseline`  the CPU oracle in reference-faithful mode (per-op affine round trips, constant-time 255-step scalar multiplication - the
             reference's operation sequence) on a bounded sample of the same matrix, all host cores; `same_algorithm`: the GPU
             library's own shortcut run on the CPU (oracle shortcut leg), so algorithmic and hardware gain separate.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_PART, THRESH = 1024, 683
MAC_PER_MODMUL = 300          # 2*12^2 + 12 wide multiply-accumulates per 12-limb Montgomery product
PAPER_PEAK_MAC = 148 * 64 * 1.965e9
METRIC = "verified shares/sec (n=1024,t=683)"


def canonical_modmul_per_share(n, t):
    """Horner, left-to-right binary [j]y (double 8, add 12, mixed add 11), + 356 for
    G*s and the comparison; averaged over ids 1..n.  (n=1024, t=683) -> 84 314."""
    tot = sum(8 * (j.bit_length() - 1) + 12 * (bin(j).count("1") - 1) + 11 for j in range(1, n + 1))
    return (t - 1) * tot / n + 356


def canonical_horner_modmul(t, x):
    """SURVEY.md 8(d) W_eval for one evaluation at |x| (binary chain, mixed add); f(0) = C_0 is free"""
    x = abs(x)
    return 0 if x == 0 else (t - 1) * (8 * (x.bit_length() - 1) + 12 * (bin(x).count("1") - 1) + 11)


# executed product-equivalents (300 wide MACs each) of the vm.cuh formulas: the fused sum-of-two-products routine
# (mul2add, 444 MACs) replaces three pairs in the additions and one in the doubling
EXEC_ADD, EXEC_DBL, EXEC_MADD = 6 + 3 * 444 / 300, 6 + 444 / 300, 5 + 3 * 444 / 300
# fixed-base multiplication G * s: the canonical algorithm of SURVEY 8(d) is 8-bit windows (32 + 1 mixed additions); the library's table
# has signed odd B-bit windows (csrc/feldman.cuh; B per ctx, default 22): the first entry initialises the sum, ceil(256 / B) - 1 mixed additions
CANON_FIX_MADDS = 33


def executed_horner_modmul(t, x):
    """what fd_seed_eval / vm_feldman_eval execute at |x|: signed-digit chain (chosen with the canonical 8 / 12
    weights, vm.cuh make_small_chain) + a full addition per step, in executed product-equivalents"""
    x = abs(x)
    if x == 0:
        return 0

    def shape(pos, neg):
        m = pos | neg
        return m.bit_length() - 1, bin(m).count("1") - 1
    k3 = 3 * x
    dbl, adds = min((shape(x, 0), shape((k3 & ~x) >> 1, (~k3 & x) >> 1)), key=lambda s_: 8 * s_[0] + 12 * s_[1])
    return (t - 1) * (EXEC_DBL * dbl + EXEC_ADD * adds + EXEC_ADD)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_PART, help="participants (default: the BASELINE config)")
    ap.add_argument("--t", type=int, default=THRESH)
    ap.add_argument("--cpu-sample", type=int, default=0, help="recipient ids per host thread in the CPU-baseline sample (0 = 24)")
    ap.add_argument("--no-cpu", action="store_true", help="skip every CPU-baseline leg")
    ap.add_argument("--gtab-bits", type=int, default=26,
                    help="window width of the fixed-base table of the ctx (dkgv_ctx_create_ex): 26 = 32 GB of the 180 GB, 9 mixed additions per G * s; "
                         "the library's own default is 22 (2.4 GB, 11)")
    ap.add_argument("--quick", action="store_true", help="main leg only (value, e2e, roofline of the default path): for profiling runs")
    ap.add_argument("--no-peak", action="store_true", help="do not run bench/imad_peak (use the paper peak); for runs under ncu")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="single GPU: process only the row block one rank of a W-rank job would hold (no collective) - the per-kernel view "
                         "of the N = W step for ncu; the printed value is NOT a whole-ceremony number")
    ap.add_argument("--pipeline", type=int, default=1, choices=[0, 1],
                    help="1 (default): value = K ceremonies queued back to back through dkgv_share_matrix_enqueue_sharded_dev, one synchronisation at the "
                         "end; 0: value = K synchronous calls (one host synchronisation per ceremony), reported as `sync_call` otherwise")
    ap.add_argument("--lanes", type=int, default=0, help="ctxs (and streams) per GPU the pipelined ceremonies alternate between (0 = by row-block size)")
    ap.add_argument("--overlap", type=int, default=1, choices=[0, 1], help="dkgv_set_share_overlap mode of the evaluation steps")
    ap.add_argument("--parts", type=int, default=0, help="parts per dealer polynomial on the finite-difference path (0 = planner)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """SM clocks / throttle reasons of one GPU during the timed region (B200_PROFILING.md recipe: the fields of its nvidia-smi line).
    Default: NVML inside this process (the library nvidia-smi itself reads), one device handle, a thread polling every 20 ms - a
    separate `nvidia-smi -lms` process attaches to every GPU of the box and was seen to perturb the 1.2 ms steps of an 8-rank job
    (profiles/r2_n8_rest_of_step.md).  DKGV_BENCH_SAMPLER=smi: the nvidia-smi process; =off: no samples (diagnosis only)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stamps, self.proc, self.thr = gpu_index, [], [], None, None
        self.mode = os.environ.get("DKGV_BENCH_SAMPLER", "nvml")
        self.stop_flag = threading.Event()
        self.source = None

    def start(self):
        if self.mode == "off":
            return
        if self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
                self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
                self.source = "NVML in-process, 20 ms period"
                self.thr = threading.Thread(target=self._poll_nvml, daemon=True)
                self.thr.start()
                return
            except Exception:  # noqa: BLE001 - no NVML binding: the nvidia-smi process below
                pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 50"
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown, nv.nvmlClocksEventReasonSwThermalSlowdown,
                nv.nvmlClocksEventReasonSwPowerCap]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            mx = ""
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append([str(self.gpu), str(sm), str(mx), "", hex(rs)] + ["Active" if rs & b else "Not Active" for b in bits])
                self.stamps.append(time.perf_counter())
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])
            self.stamps.append(time.perf_counter())

    def stop(self, since=None):
        """summary of the samples taken after perf_counter() time `since` (None: all of them)"""
        if self.mode == "off":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler switched off (DKGV_BENCH_SAMPLER=off)"]}
        if self.proc is None and self.thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"]}
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        elif self.thr:
            self.thr.join(timeout=2)
        if since is not None:
            n_rows = min(len(self.rows), len(self.stamps))
            self.rows = [self.rows[i] for i in range(n_rows) if self.stamps[i] >= since]
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(self.NAMES, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def measured_int_peak():
    """IMAD.WIDE lane-ops/s measured live with bench/imad_peak; falls back to the paper figure."""
    exe = os.path.join(ROOT, "bench", "imad_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120, check=True).stdout
        j = json.loads(out)
        return {"imad_wide": j["imad_wide"]["lane_ops_per_s"], "imad_wide_x": j["imad_wide_x"]["lane_ops_per_s"],
                "fp_mul_per_s": j["fp_mul_32warps_per_sm"]["modmul_per_s"], "source": "measured live (bench/imad_peak)"}
    except Exception as e:  # noqa: BLE001
        return {"imad_wide": PAPER_PEAK_MAC, "imad_wide_x": None, "fp_mul_per_s": None,
                "source": f"fallback paper peak 148 SM x 64 lanes x 1.965 GHz ({type(e).__name__})"}


# ------------------------------------------------------------------------------------------ CPU legs (the only users of oracle/)
def cpu_share_session(n, t, rows, ids):
    """`rows` copies of one dealer of the synthetic ceremony at the recipient ids `ids` (built with the oracle itself)"""
    import numpy as np
    import oracle_lib as O
    from dvt_circuits_b200 import synthetic
    coeffs = synthetic.make_coefficients(1, t)
    vv1 = np.zeros((t, 48), dtype=np.uint8)
    cs = [int.from_bytes(coeffs[0, k].tobytes(), "big") for k in range(t)]
    for k in range(t):
        st, pk = O.g1_fixed_base(coeffs[0, k].tobytes(), O.FAST)
        vv1[k] = np.frombuffer(pk, dtype=np.uint8)
    sh1 = np.zeros((len(ids), 32), dtype=np.uint8)
    for j, i in enumerate(ids.tolist()):
        acc = 0
        for c in reversed(cs):
            acc = (acc * i + c) % synthetic.R_INT
        sh1[j] = np.frombuffer(acc.to_bytes(32, "big"), dtype=np.uint8)
    return np.broadcast_to(vv1, (rows, t, 48)).copy(), np.broadcast_to(sh1, (rows, len(ids), 32)).copy()


def cpu_baseline(n, t, cols, threads):
    """oracle in reference-faithful mode on a bounded sample of the same matrix (threads x cols shares: one dealer row per host
    thread, `cols` recipient ids spread over 1..n).  The sample rows reuse one dealer's polynomial - the faithful cost per share
    depends on t only - and the verification vector is decoded once per row, which under-counts the reference (it decodes all t
    points per share, verification.rs:132-137): the baseline is favoured, not the GPU."""
    import numpy as np
    import oracle_lib as O
    rows = max(1, threads)
    ids = np.unique(np.linspace(1, n, cols).astype(np.uint32))
    cols = len(ids)
    vv, shares = cpu_share_session(n, t, rows, ids)
    t0 = time.perf_counter()
    st = O.share_matrix(vv, ids, shares, O.FAITHFUL, threads=threads)
    dt = time.perf_counter() - t0
    assert not st.any(), "CPU oracle rejected a valid synthetic share"
    t1 = time.perf_counter()
    st = O.share_matrix(vv, ids, shares, O.FAST, threads=threads)
    dt_fast = time.perf_counter() - t1
    return {"value": rows * cols / dt, "unit": "shares/s", "cores": threads, "kind": "port",
            "sample": f"{rows}x{cols} shares of the n={n},t={t} matrix, oracle faithful mode (reference op sequence), {dt:.1f}s wall",
            "fast_mode_value": rows * cols / dt_fast, "n_shares": rows * cols}


def cpu_same_algorithm(n, t, threads):
    """the GPU library's own exact shortcut (range, t-th differences, compress(G p_k) == C_k) on the host cores: `threads` dealers
    of the ceremony with all n shares each, oracle shortcut leg.  Separates the algorithmic gain from the hardware gain."""
    import numpy as np
    import oracle_lib as O
    ids = np.arange(1, n + 1, dtype=np.uint32)
    rows = threads * 8
    vv, shares = cpu_share_session(n, t, rows, ids)
    O.share_matrix_shortcut(vv[:1], shares[:1], threads=1)  # builds the fixed-base table (untimed, as the GPU's)
    t0 = time.perf_counter()
    st, fb = O.share_matrix_shortcut(vv, shares, threads=threads)
    dt = time.perf_counter() - t0
    assert fb == 0 and not st.any()
    return {"value": rows * n / dt, "unit": "shares/s", "cores": threads, "kind": "port",
            "sample": f"{rows} dealers x {n} shares, the consistency shortcut of the GPU path run by the CPU oracle, {dt:.2f}s wall"}


def cpu_pairing(threads, per_thread=96):
    """oracle bls_verify_precomputed_hash exactly as the reference computes it (two full pairings + Gt equality, decoding included)"""
    import numpy as np
    import oracle_lib as O
    from oracle.pyref import bls12_381 as B
    hmp = B.hash_to_g2(b"Sign with new partial key")
    hm = B.g2_compress(hmp)
    pks, sigs = [], []
    for i in range(4):
        sk = 0x1234567 + i
        pks.append(list(B.g1_compress(B.g1_mul(B.G1, sk))))
        sigs.append(list(B.g2_compress(B.g2_mul(hmp, sk))))
    m = threads * per_thread
    pk = np.array([pks[i % 4] for i in range(m)], dtype=np.uint8)
    sg = np.array([sigs[i % 4] for i in range(m)], dtype=np.uint8)
    O.fp_mul_count(reset=True)
    O.bls_verify_hm(bytes(pk[0]), bytes(sg[0]), hm)
    modmul = O.fp_mul_count(reset=True)
    t0 = time.perf_counter()
    out = O.bls_verify_batch(pk, sg, hm, threads=threads)
    dt = time.perf_counter() - t0
    assert (out == 1).all()
    return {"value": m / dt, "unit": "checks/s", "cores": threads, "kind": "port",
            "sample": f"{m} checks (decode + 2 full pairings each, as bls_common.rs:26-35), {dt:.1f}s wall",
            "modmul_per_check_reference_sequence": modmul}


def cpu_config_a(threads):
    """BASELINE config 2 on the CPU: the FULL 64 x 64 matrix at t = 43, oracle faithful mode, all cores"""
    import numpy as np
    import oracle_lib as O
    n, t = 64, 43
    ids = np.arange(1, n + 1, dtype=np.uint32)
    vv, shares = cpu_share_session(n, t, n, ids)
    t0 = time.perf_counter()
    st = O.share_matrix(vv, ids, shares, O.FAITHFUL, threads=threads)
    dt = time.perf_counter() - t0
    assert not st.any()
    return {"value": n * n / dt, "unit": "shares/s", "cores": threads, "kind": "port",
            "sample": f"full {n}x{n} matrix, t={t}, oracle faithful mode, {dt:.1f}s wall"}


def cpu_finalization(fin, threads):
    """config 4 on the CPU at a bounded size (the (64, 43) ceremony `fin`): agg_coefficients + keys, two Lagrange interpolations and
    the n partial-signature checks, oracle faithful mode (the aggregation and Lagrange are single-threaded, as in the reference)"""
    import oracle_lib as O
    n = fin["ids"].shape[0]
    t0 = time.perf_counter()
    ast, co, keys = O.agg_coefficients(fin["vv"], fin["ids"], O.FAITHFUL)
    l1 = O.lagrange(keys, fin["ids"], O.FAITHFUL)
    l2 = O.lagrange(fin["partial_pubkeys"], fin["ids"], O.FAITHFUL)
    out = O.bls_verify_batch(fin["partial_pubkeys"], fin["signatures"], bytes(fin["hm"]), threads=threads)
    dt = time.perf_counter() - t0
    assert ast == 0 and l1 == l2 and l1[1] == bytes(co[0]) and (out == 1).all()
    return {"total_ms": dt * 1e3, "n": n, "t": int(fin["vv"].shape[1]), "cores": threads, "kind": "port",
            "sample": f"whole finalization of a ({n}, {fin['vv'].shape[1]}) ceremony, oracle faithful mode"}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric.  The Rust crate cannot be built
    in this image (no cargo/rustc, un-vendored git dependencies), so this times the C++ oracle in
    reference-faithful mode on all host cores, one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_lib as O
    O.use_native_build()
    threads = os.cpu_count() or 1
    cols = args.cpu_sample or 24
    vals, cb = [], None
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue  # one warm-up pass is enough for a CPU loop; keeps the run within minutes
        cb = cpu_baseline(args.n, args.t, cols, threads)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = statistics.mean(vals)
    line = {"metric": METRIC, "value": v, "unit": "shares/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cb["n_shares"] / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64 limbs (381-bit Montgomery Fp)", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"synthetic DKG n={args.n}, t={args.t}: share-matrix verification, bounded sample of {cb['n_shares']} shares per step",
                       "n": args.n, "t": args.t},
            "cpu_baseline": {"value": v, "unit": "shares/s", "cores": threads, "kind": "port", "sample": cb["sample"]},
            "e2e": {"value": v, "unit": "shares/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "C++ oracle restating the reference's operation sequence; the Rust reference itself cannot be built here"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import pipeline, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): whatever native libraries print there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    n, t = args.n, args.t
    split = args.emulate_world if (args.emulate_world and world == 1) else world
    if n % split:
        raise SystemExit("participants must divide by the number of ranks")
    rows = n // split

    v = dk.Verifier(local, gtab_bits=int(os.environ.get("DKGV_GTAB_BITS", args.gtab_bits)))
    gtab_bits = v.gtab_bits()
    EXEC_FIX_MADDS = (256 + gtab_bits - 1) // gtab_bits - 1
    if world > 1:  # the library owns the collectives; torch.distributed only carries the 128-byte id to the other ranks
        box = [dk.Verifier.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        v.comm_init(box[0], rank, world)
    v.set_share_parts(args.parts)
    v.set_share_overlap(args.overlap)
    sess = synthetic.make_session(v, rows, n, t, dealer_offset=rank * rows)  # set-up, untimed
    ts = torch.cuda.Stream(device=dev)
    stream = ts.cuda_stream
    chunk = v.share_gather_words(rows, n)
    words = (rows * n + 31) // 32

    def dmax(x):
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(ts):
        d_vv = torch.from_numpy(sess["vv"]).to(dev)
        d_ids = torch.from_numpy(sess["ids"].view(np.int32)).to(dev)
        d_sh = torch.from_numpy(sess["shares"]).to(dev)
        d_st = torch.empty((rows, n), dtype=torch.uint8, device=dev)
        d_gather = torch.zeros((world, chunk), dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # pinned host copies for the end-to-end leg
    h_vv = torch.from_numpy(sess["vv"]).pin_memory()
    h_ids = torch.from_numpy(sess["ids"].view(np.int32)).pin_memory()
    h_sh = torch.from_numpy(sess["shares"]).pin_memory()
    h_st = torch.empty((rows, n), dtype=torch.uint8).pin_memory()
    h_gather = torch.empty((world, chunk), dtype=torch.int32).pin_memory()

    def step_device(shares=None):
        v.share_matrix_verify_sharded_dev(rows, n, t, d_vv.data_ptr(), d_ids.data_ptr(), (d_sh if shares is None else shares).data_ptr(),
                                          d_st.data_ptr(), d_gather.data_ptr(), stream)

    def step_e2e():
        if world == 1:
            v._ck(v._lib.dkgv_share_matrix_verify(v._h, rows, n, t, h_vv.data_ptr(), h_ids.data_ptr(), h_sh.data_ptr(), h_st.data_ptr()))
        else:  # host rows in, gathered verdict bitmask of the WHOLE ceremony out
            d_vv.copy_(h_vv, non_blocking=True)
            d_ids.copy_(h_ids, non_blocking=True)
            d_sh.copy_(h_sh, non_blocking=True)
            step_device()
            h_gather.copy_(d_gather, non_blocking=True)
            h_st.copy_(d_st, non_blocking=True)
            ts.synchronize()

    def timed_steps(fn, count, do_flush=True, ranks_together=True):
        """per-step CUDA events around fn(); ranks_together: every rank calls this the same number of times (a barrier precedes each step) -
        False inside a leg only ONE rank runs (a barrier there would wait for ranks that never come)"""
        out = []
        for _ in range(count):
            if do_flush and os.environ.get("DKGV_BENCH_FLUSH", "1") != "0":  # (=0: diagnosis only, the number is then not a bench value)
                flush.fill_(1)  # L2 flush between timed iterations (not timed)
            if world > 1 and ranks_together:
                barrier()  # every step starts on all ranks together: its collective is not charged with host-side skew between the ranks
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            h0 = time.perf_counter()
            fn()
            host_ms.append((time.perf_counter() - h0) * 1e3)
            e1.record(ts)
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return out

    host_ms = []  # host wall clock inside every timed call (launches + the call's own synchronisation), same order as the device times

    def all_ranks(x):  # per-step list of this rank -> [rank][step] on every rank
        a = torch.tensor(x, dtype=torch.float64, device=dev)
        if world == 1:
            return [a.tolist()]
        box = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(box, a)
        return [b.tolist() for b in box]

    def bad_bits():
        g = d_gather[:, :words]
        return int(torch.count_nonzero(g).item())

    peak = None
    if rank == 0:
        peak = measured_int_peak() if not args.no_peak else {"imad_wide": PAPER_PEAK_MAC, "imad_wide_x": None, "fp_mul_per_s": None,
                                                            "source": "paper peak 148 SM x 64 lanes x 1.965 GHz (--no-peak)"}
    line = {}
    with torch.cuda.stream(ts):
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()  # before the warm-up: nvidia-smi needs a moment before its first sample; only samples from wall0 on count
        for _ in range(args.warmup):
            step_device()
        barrier()
        launches0 = v.launch_count
        barrier()
        wall0 = time.perf_counter()
        step_ms, short_ms = [], []
        del host_ms[:]
        for _ in range(args.steps):
            step_ms += timed_steps(step_device, 1)
            if v.last_share_path == v.PATH_FDIFF and not v.last_share_continued:
                short_ms.append(v.last_share_phases_ms())  # [share limbs + difference table, x halves of G*p_k == C_k, sign halves, flags]
        barrier()
        wall = time.perf_counter() - wall0
        launches = v.launch_count - launches0
        # The timed region of the default path is short (10 steps of ~13 ms on one GPU, less on eight) against nvidia-smi's 50 ms
        # sampling period: keep the same steps running, untimed, until the GPU has been under this load for ~0.6 s, so that the
        # clocks line rests on several samples.  The count comes from the max-over-ranks step time: identical on every rank.
        ms_per_step = dmax(sum(step_ms)) / args.steps
        per_rank = all_ranks(step_ms)
        per_rank_host = all_ranks(host_ms[:args.steps])
        step_spread = {"per_step_max_over_ranks_ms": [round(max(r[i] for r in per_rank), 4) for i in range(args.steps)],
                       "per_rank_mean_ms": [round(sum(r) / args.steps, 4) for r in per_rank],
                       "median_of_per_step_max_ms": round(statistics.median(max(r[i] for r in per_rank) for i in range(args.steps)), 4),
                       "per_rank_host_ms_in_call_mean": [round(sum(r) / args.steps, 4) for r in per_rank_host],
                       "note": "ms_per_step (the contract's number) = max over ranks of the mean; the per-step maxima show how much of it is single slow steps"}
        sync_call = {"ms_per_step": ms_per_step, "value": rows * n * world / (ms_per_step * 1e-3), "unit": "shares/s", "step_spread": step_spread,
                     "what": "K synchronous dkgv_share_matrix_verify_sharded_dev calls: every ceremony ends with its own host synchronisation "
                             "(flag read-back); L2 flushed (256 MB fill) before every call; per-call CUDA events, max over ranks of the mean"}
        pipe = None
        if args.pipeline:
            # ---- the headline: K ceremonies in flight.  A host with many ceremonies queues them back to back (enqueue), synchronises once and
            # settles them; nothing on the honest path waits for the host between two ceremonies.  Inputs: a ring of distinct device copies
            # larger than twice the L2, so no ceremony finds its verification vectors or shares in the cache; one flush before the region.
            n_lanes = args.lanes or pipeline.lanes_for(rows)
            lane_v, lane_s = [v], [ts]
            for _ in range(n_lanes - 1):
                lv = dk.Verifier(local, gtab_bits=gtab_bits)
                if world > 1:
                    box = [dk.Verifier.comm_unique_id() if rank == 0 else None]
                    dist.broadcast_object_list(box, src=0)
                    lv.comm_init(box[0], rank, world)
                lv.set_share_parts(args.parts)
                lv.set_share_overlap(args.overlap)
                lane_v.append(lv)
                lane_s.append(torch.cuda.Stream(device=dev))
            bytes_step = d_vv.numel() + d_sh.numel()
            ring = pipeline.ring_size(bytes_step)
            slots = [{"vv": d_vv.clone() if i else d_vv, "sh": d_sh.clone() if i else d_sh} for i in range(ring)]
            depth = min(args.steps, 256)
            outs = [{"st": torch.empty_like(d_st), "g": torch.zeros_like(d_gather), "hf": torch.zeros(2 * world, dtype=torch.int32).pin_memory()}
                    for _ in range(max(depth, args.warmup, n_lanes))]

            def job_args(k, lane):
                sl, o = slots[k % ring], outs[k % len(outs)]
                return (rows, n, t, sl["vv"].data_ptr(), d_ids.data_ptr(), sl["sh"].data_ptr(), o["st"].data_ptr(), o["g"].data_ptr(), o["hf"].data_ptr(),
                        lane_s[lane].cuda_stream)

            def pipe_pass(count):
                """count ceremonies, device time from before the first enqueue to after the last settle (ms) and the ceremonies settle ran again"""
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reran, first = 0, 0
                e0.record(ts)
                for s_ in lane_s[1:]:
                    s_.wait_event(e0)
                for k in range(count):
                    lane_v[k % n_lanes].share_matrix_enqueue_sharded_dev(*job_args(k, k % n_lanes))
                    if (k + 1 - first) == len(outs) or k == count - 1:  # a wave is in flight: join the lanes, ONE synchronisation, settle
                        for s_ in lane_s[1:]:
                            ev = torch.cuda.Event()
                            ev.record(s_)
                            ts.wait_event(ev)
                        ts.synchronize()
                        for kk in range(first, k + 1):
                            reran += lane_v[kk % n_lanes].share_matrix_settle_sharded_dev(*job_args(kk, kk % n_lanes))
                        first = k + 1
                e1.record(ts)
                e1.synchronize()
                return e0.elapsed_time(e1), reran

            pipe_pass(max(args.warmup, n_lanes))  # warm-up of every lane (its scratch is allocated on first use)
            flush.fill_(1)
            barrier()
            launches0 = sum(lv.launch_count for lv in lane_v)
            wall0p = time.perf_counter()
            pipe_ms, pipe_reran = pipe_pass(args.steps)
            wall_pipe = time.perf_counter() - wall0p
            launches_pipe = sum(lv.launch_count for lv in lane_v) - launches0
            barrier()
            pipe_bad = sum(int(torch.count_nonzero(o["g"][:, :words]).item()) for o in outs[:min(depth, args.steps)]) + pipe_reran
            repeats = []
            for _ in range(2):  # spread only: the number is the first pass
                flush.fill_(1)
                barrier()
                repeats.append(dmax(pipe_pass(args.steps)[0]) / args.steps)
            ms_per_step = dmax(pipe_ms) / args.steps
            pipe = {"lanes": n_lanes, "ring": ring, "ring_bytes": int(ring * bytes_step), "bad": pipe_bad, "repeat_ms_per_step": repeats,
                    "wall_s": wall_pipe, "launches": launches_pipe}
        extra_steps = max(0, min(2000, int(0.6 / max(ms_per_step * 1e-3, 1e-4)) - args.steps))
        for _ in range(extra_steps):
            step_device()
        barrier()
        clocks = sampler.stop(since=wall0) if rank == 0 else None
        if clocks is not None:
            clocks["untimed_steps_under_the_same_load"] = extra_steps
        bad = bad_bits()
        settled_by_shortcut = v.last_share_path == v.PATH_FDIFF and not v.last_share_continued
        shares_total = rows * n * world

        # end-to-end with host buffers: synchronous calls (one ceremony per call) ...
        step_e2e()
        barrier()
        e2e_t = []
        for _ in range(args.steps):
            barrier()
            t0 = time.perf_counter()
            step_e2e()
            torch.cuda.synchronize()
            e2e_t.append(time.perf_counter() - t0)
        e2e_sync_s_per_step = dmax(sum(e2e_t)) / args.steps
        e2e_s_per_step = e2e_sync_s_per_step
        bad_e2e = int(h_st.count_nonzero().item())
        e2e_pipe = None
        if pipe:
            # ... and the same K ceremonies in flight: dkgv_share_matrix_enqueue_sharded (pinned host rows in; status bytes, gathered bitmask and
            # flags back on the host) alternating over the lanes, dkgv_sync per lane, settle for each - host wall clock around all of it
            h_outs = [{"st": torch.empty((rows, n), dtype=torch.uint8).pin_memory(), "g": torch.empty((world, chunk), dtype=torch.int32).pin_memory(),
                       "hf": torch.zeros(2 * world, dtype=torch.int32).pin_memory()} for _ in range(len(outs))]

            def e2e_args(k):
                o = h_outs[k % len(h_outs)]
                return (rows, n, t, h_vv.data_ptr(), h_ids.data_ptr(), h_sh.data_ptr(), o["st"].data_ptr(), o["g"].data_ptr(), o["hf"].data_ptr())

            def e2e_pass(count):
                reran, first = 0, 0
                t0 = time.perf_counter()
                for k in range(count):
                    lane_v[k % n_lanes].share_matrix_enqueue_sharded(*e2e_args(k))
                    if (k + 1 - first) == len(h_outs) or k == count - 1:
                        for lv in lane_v:
                            lv.sync()
                        for kk in range(first, k + 1):
                            reran += lane_v[kk % n_lanes].share_matrix_settle_sharded(*e2e_args(kk))
                        first = k + 1
                return time.perf_counter() - t0, reran

            e2e_pass(max(args.warmup, n_lanes))
            barrier()
            e2e_wall, e2e_reran = e2e_pass(args.steps)
            barrier()
            e2e_s_per_step = dmax(e2e_wall) / args.steps
            bad_e2e += e2e_reran + sum(int(o["st"].count_nonzero().item()) + int(torch.count_nonzero(o["g"][:, :words]).item())
                                       for o in h_outs[:min(len(h_outs), args.steps)])
            e2e_pipe = True
            for lv in lane_v[1:]:
                lv.close()

        legs = {}
        if not args.quick:
            # ---- the evaluation route: shortcut off (every share through the group arithmetic), production mode then phase by phase
            aux = min(args.steps, 3)
            v.set_share_shortcut(False)
            step_device()  # untimed: the evaluation buffers are allocated on first use
            full_ms = timed_steps(step_device, aux)
            v.set_share_overlap(0)
            phase_ms, serial_ms = [], []
            for _ in range(aux):
                serial_ms += timed_steps(step_device, 1)
                phase_ms.append(v.last_share_phases_ms())
            v.set_share_overlap(args.overlap)
            v.set_share_shortcut(True)
            barrier()
            full_ms_step = dmax(sum(full_ms)) / aux
            legs["full_evaluation"] = {"metric": "verified shares/sec, every share evaluated in the group (consistency shortcut off)",
                                       "value": shares_total / (full_ms_step * 1e-3), "unit": "shares/s", "ms_per_step": full_ms_step}

            # ---- corrupted shares: the reference exists to PROVE misbehaviour; how the default path degrades with the corruption rate
            rng = np.random.Generator(np.random.PCG64([0xBAD, rank]))
            corr = {}
            for name, kind in (("one_share", 1), ("one_dealer", 2), ("one_dealer_in_every_group", 3), ("p_1pct", 0.01), ("p_10pct", 0.10),
                               ("p_50pct_config5", 0.50)):
                mask = np.zeros((rows, n), dtype=bool)
                if kind == 1:
                    if rank == 0:
                        mask[rows // 2, n // 3] = True
                elif kind == 2:
                    if rank == 0:
                        mask[rows // 2, :] = True
                elif kind == 3:  # every 32-dealer group holds one dealer that is wrong throughout: the per-dealer fallback evaluates rows / 32 dealers
                    mask[5::32, :] = True
                else:
                    mask = rng.random((rows, n)) < kind
                sh_bad = sess["shares"].copy()
                sh_bad[mask, 31] ^= 1
                d_bad = torch.from_numpy(sh_bad).to(dev)
                step_device(d_bad)
                ms = timed_steps(lambda: step_device(d_bad), 2)
                ok = bool(((d_st != 0) == torch.from_numpy(mask).to(dev)).all().item())
                ok_all = dmax(0.0 if ok else 1.0) == 0.0
                ms_step = dmax(sum(ms)) / len(ms)
                corr[name] = {"corrupted_shares_this_rank0": int(mask.sum()), "ms_per_step": ms_step,
                              "value": shares_total / (ms_step * 1e-3), "unit": "shares/s",
                              "verdicts_flag_exactly_the_corrupted_shares": ok_all, "took_the_evaluation": bool(v.last_share_continued)}
                del d_bad
            legs["corruption"] = corr
            legs["mixed_items"] = dict(corr["p_50pct_config5"], metric="verified shares/sec, 50 % of the shares corrupted (BASELINE config 5)")

            # ---- second BASELINE metric: BLS pairing checks/s (bls_verify_precomputed_hash, one common message), 262 144 checks sharded
            # over the ranks, every 7th with another signer's (valid, wrong) signature; status bytes all-gathered inside the library
            m_total = 262144
            m_loc = m_total // world
            fin = synthetic.make_finalization(v, 64, 8)
            reps = (m_loc + 63) // 64
            pk_np = np.tile(fin["partial_pubkeys"], (reps, 1))[:m_loc].copy()
            sg_np = np.tile(fin["signatures"], (reps, 1))[:m_loc].copy()
            wrong = np.arange(m_loc) % 7 == 3
            sg_np[wrong] = np.roll(sg_np, 1, axis=0)[wrong]
            d_pk, d_sg = torch.from_numpy(pk_np).to(dev), torch.from_numpy(sg_np).to(dev)
            d_hm = torch.from_numpy(fin["hm"].copy()).to(dev)
            d_pall = torch.empty((world, m_loc), dtype=torch.uint8, device=dev)

            def step_pairing():
                v.bls_verify_batch_sharded_dev(m_loc, d_pk.data_ptr(), d_sg.data_ptr(), 1, d_hm.data_ptr(), None, d_pall.data_ptr(), stream)

            step_pairing()
            barrier()
            pair_ms, pair_kernel_ms = [], []
            for _ in range(max(2, min(args.steps, 5))):
                pair_ms += timed_steps(step_pairing, 1, do_flush=False)
                pair_kernel_ms.append(v.last_bls_kernel_ms())
            pair_ms_step = dmax(sum(pair_ms)) / len(pair_ms)
            want = torch.from_numpy(np.where(wrong, 7, 0).astype(np.uint8)).to(dev)
            pair_ok = bool((d_pall == want.unsqueeze(0)).all().item())
            # e2e: host buffers through dkgv_bls_verify_batch (H2D of keys + signatures, D2H of the statuses inside)
            h_pk, h_sg = torch.from_numpy(pk_np).pin_memory(), torch.from_numpy(sg_np).pin_memory()
            h_hm, h_pst = torch.from_numpy(fin["hm"].copy()).pin_memory(), torch.empty(m_loc, dtype=torch.uint8).pin_memory()
            pair_e2e_t = []
            for _ in range(3):  # the first call is the warm-up of the staging buffers
                barrier()
                t0 = time.perf_counter()
                v._ck(v._lib.dkgv_bls_verify_batch(v._h, m_loc, h_pk.data_ptr(), h_sg.data_ptr(), 1, h_hm.data_ptr(), None, h_pst.data_ptr()))
                pair_e2e_t.append(time.perf_counter() - t0)
            pair_e2e_s = dmax(sum(pair_e2e_t[1:])) / 2
            st_h = h_pst.numpy()
            prog = json.load(open(os.path.join(ROOT, "dvt_circuits_b200", "csrc", "pairing_prog.json")))
            kms = statistics.mean(pair_kernel_ms)
            legs["pairing"] = {
                "metric": "BLS pairing checks/sec", "value": m_total / (pair_ms_step * 1e-3), "unit": "checks/s",
                "checks_per_step": m_total, "ms_per_step": pair_ms_step, "wrong_signatures_per_step": int(wrong.sum()) * world,
                "verdicts_flag_exactly_the_wrong_signatures": pair_ok and bool((st_h == np.where(wrong, 7, 0)).all()),
                "path": {1: "pairing VM (6 warps per 32 checks, operands in shared memory)", 2: "one thread per check"}[v.last_bls_path],
                "e2e": {"value": m_total / pair_e2e_s, "unit": "checks/s", "h2d_bytes_per_step": int(m_total * 144 + 96), "d2h_bytes_per_step": m_total,
                        "timing": "host wall clock around dkgv_bls_verify_batch (pinned host buffers), mean of 2 calls after one warm-up, max over ranks"},
                "note": "e(pk,H(m)) == e(G1,sig) as ONE product of two Miller loops + one final exponentiation per check, incl. G1/G2 decoding "
                        "with subgroup checks and the preparation of the hashed message's lines",
            }
            if rank == 0:
                macs = prog["wide_macs_per_check"]
                ach = m_loc * macs / (kms * 1e-3)
                legs["pairing"]["roofline"] = {
                    "bound": "int_pipe", "kernel": "k_pairing_vm", "kernel_ms": kms, "kernel_share_of_step": kms / statistics.mean(pair_ms),
                    "achieved": ach / 1e9, "peak": peak["imad_wide"] / 1e9, "unit": "G wide-MAC/s (32x32->64)", "frac": ach / peak["imad_wide"],
                    "frac_of_carry_chain_peak": (ach / peak["imad_wide_x"]) if peak["imad_wide_x"] else None,
                    "executed_wide_macs_per_check": macs, "executed_fp_mul_equivalents_per_check": prog["fp_mul_equivalents_per_check"],
                    "fp2_products_per_check": prog["fp2_mul_per_check"], "fp2_squarings_per_check": prog["fp2_sqr_per_check"],
                    "barriers_per_check": prog["barriers_per_check"], "units_per_launch": m_loc, "unit_is": "one pairing check (Miller loops + final exponentiation; decoding is in other kernels)",
                    "traffic": 327 * m_loc, "traffic_ref": "profiles/r2_pairing_vm.md: dram__bytes_read.sum + dram__bytes_write.sum of one k_pairing_vm launch = 327 B per check (no local memory)",
                    "algorithmic_bytes": 145 * m_loc,
                    "note": "executed products counted from the VM program (csrc/pairing_prog.json); an Fp2 product = 2 fused sums of two products (888 MACs), a squaring = 2 products (600)"}
            del d_pk, d_sg, d_pall

            # ---- hash_message_to_g2 on the GPU: 65 536 distinct 32-byte messages through dkgv_hash_to_g2 (host buffers)
            if rank == 0:
                msgs_np = np.random.Generator(np.random.PCG64(0x42)).integers(0, 256, size=(65536, 32), dtype=np.uint8)
                msgs = [bytes(r_) for r_ in msgs_np[:64]]
                assert bytes(v.hash_to_g2(msgs[:1])[0]) == bytes(v.hash_to_g2(msgs[:2])[0])
                blob = np.ascontiguousarray(msgs_np).reshape(-1)
                offs = (np.arange(65537, dtype=np.uint32) * 32).astype(np.uint32)
                outb = np.zeros((65536, 96), dtype=np.uint8)
                import ctypes as _ct
                t0 = time.perf_counter()
                v._ck(v._lib.dkgv_hash_to_g2(v._h, 65536, blob.ctypes.data_as(_ct.c_void_p), offs.ctypes.data_as(_ct.c_void_p),
                                             outb.ctypes.data_as(_ct.c_void_p)))
                h2c_s = time.perf_counter() - t0
                legs["hash_to_g2"] = {"metric": "hash_message_to_g2 (RFC 9380, G2, SHA-256 XMD, SSWU, 3-isogeny, h_eff) messages/sec", "value": 65536 / h2c_s,
                                      "unit": "messages/s", "messages": 65536, "s": h2c_s, "timing": "host wall clock around dkgv_hash_to_g2 (host buffers)",
                                      "note": "one thread per message; distinct from the pairing batch, where one message is hashed once"}

            # ---- BASELINE config 5, second half: 1 M bad-partial-key items over a (64, 43) session, half of them corrupted
            m_bp = (1 << 20) // world
            fin_a = synthetic.make_finalization(v, 64, 43)
            items = synthetic.make_bad_partial_items(v, fin_a, m_bp, p_bad=0.5, seed=synthetic.DEFAULT_SEED + rank)
            st_bp, exp_keys, sst = v.bad_partial_key_verify_batch(fin_a["vv"], items["perp"], items["pk"], items["sig"], [fin_a["message"]])
            barrier()
            t0 = time.perf_counter()
            st_bp, exp_keys, sst = v.bad_partial_key_verify_batch(fin_a["vv"], items["perp"], items["pk"], items["sig"], [fin_a["message"]])
            # verdict = 1 bit per item, gathered over NVLink by the library's communicator
            d_bst = torch.from_numpy(st_bp).to(dev)
            d_bbits = torch.zeros((world, (m_bp + 31) // 32), dtype=torch.int32, device=dev)
            v.pack_verdicts_dev(m_bp, d_bst.data_ptr(), d_bbits[rank].data_ptr(), stream)
            v.all_gather_dev(d_bbits[rank].data_ptr(), d_bbits.data_ptr(), d_bbits[rank].numel() * 4, stream)
            ts.synchronize()
            bp_s = dmax(time.perf_counter() - t0)
            bp_ok = dmax(0.0 if (sst == 0 and (st_bp == items["expected"]).all()) else 1.0) == 0.0
            legs["bad_partial_key"] = {
                "metric": "bad-partial-key items/sec (prove_wrong_final_key_generation, 1 M items over a (64, 43) session, 50 % corrupted)",
                "value": m_bp * world / bp_s, "unit": "items/s", "items_per_step": m_bp * world, "s_per_step": bp_s,
                "statuses_equal_the_expected_ones": bp_ok, "status_histogram_rank0": {int(k): int(c) for k, c in zip(*np.unique(st_bp, return_counts=True))},
                "slashable_bits_gathered": int(torch.count_nonzero(d_bbits).item() > 0),
                "timing": "host wall clock around dkgv_bad_partial_key_verify_batch (host buffers: H2D of keys, signatures, perpetrator "
                          "indices; aggregation + expected keys of the session; hash-to-G2; decode + pairing of every item; D2H of the statuses) "
                          "+ bitmask all-gather, max over ranks"}
            del d_bst, d_bbits

            # ---- BASELINE config 4: finalization of the whole ceremony.  Aggregation sharded by generations (each rank decodes and
            # sums its rows, one all-gather of the t partial sums), the n partial-signature checks sharded by items; the Lagrange
            # interpolations and the evaluation of the n final keys are latency-bound one-thread-per-term kernels and run replicated.
            ff = synthetic.make_finalization(v, n, t) if rank == 0 or world > 1 else None
            fr = slice(rank * (n // world), (rank + 1) * (n // world))
            best = None
            for _ in range(2):
                barrier()
                t0 = time.perf_counter()
                ast, co, keys = v.agg_final_keys_sharded(ff["vv"][fr], ff["ids"])
                t1 = time.perf_counter()
                l1 = v.lagrange_at_zero(keys, ff["ids"])
                l2 = v.lagrange_at_zero(ff["partial_pubkeys"], ff["ids"])
                t2 = time.perf_counter()
                st_f = v.bls_verify_batch(ff["partial_pubkeys"][fr], ff["signatures"][fr], ff["hm"])
                t3 = time.perf_counter()
                ok = bool(ast == 0 and (keys == ff["partial_pubkeys"]).all() and l1 == (0, bytes(co[0])) and l2 == l1 and not st_f.any())
                cur = {"metric": "finalization of one ceremony (host buffers, wall clock, max over ranks)", "n": n, "t": t,
                       "total_ms": dmax(t3 - t0) * 1e3, "agg_final_keys_ms": dmax(t1 - t0) * 1e3, "lagrange_x2_ms": dmax(t2 - t1) * 1e3,
                       "partial_signature_checks_ms": dmax(t3 - t2) * 1e3, "all_checks_hold": dmax(0.0 if ok else 1.0) == 0.0,
                       "sharding": f"aggregation: {n // world} generations per rank + all-gather of {t} x 144 B partial sums; signature checks: "
                                   f"{n // world} per rank; final keys + Lagrange replicated"}
                if best is None or cur["total_ms"] < best["total_ms"]:
                    best = cur
            legs["finalization"] = best

            # ---- BASELINE config 2 (rank 0, one GPU): n = 64, t = 43, the full 64 x 64 matrix
            if rank == 0:
                sa = synthetic.make_session(v, 64, 64, 43)
                da_vv, da_ids = torch.from_numpy(sa["vv"]).to(dev), torch.from_numpy(sa["ids"].view(np.int32)).to(dev)
                da_sh, da_st = torch.from_numpy(sa["shares"]).to(dev), torch.empty((64, 64), dtype=torch.uint8, device=dev)

                def step_a():
                    v.share_matrix_verify_dev(64, 64, 43, da_vv.data_ptr(), da_ids.data_ptr(), da_sh.data_ptr(), da_st.data_ptr(), stream)
                step_a()
                a_ms = statistics.mean(timed_steps(step_a, 5, do_flush=False, ranks_together=False))
                v.set_share_shortcut(False)
                step_a()
                a_full_ms = statistics.mean(timed_steps(step_a, 3, do_flush=False, ranks_together=False))
                v.set_share_shortcut(True)
                legs["config_a"] = {"metric": "verified shares/sec (n=64,t=43), full 64x64 matrix on one GPU", "value": 4096 / (a_ms * 1e-3),
                                    "unit": "shares/s", "ms_per_step": a_ms, "bad_verdicts": int(da_st.count_nonzero().item()),
                                    "every_share_evaluated": {"value": 4096 / (a_full_ms * 1e-3), "ms_per_step": a_full_ms},
                                    "note": "latency-bound: 4 096 shares are a fraction of one wave"}

    if rank == 0:
        value = shares_total / (ms_per_step * 1e-3)
        MODMUL_PER_SHARE = canonical_modmul_per_share(n, t)
        step_mean = statistics.mean(step_ms)

        def kernel_entry(name, units, unit_is, canon_unit, exec_unit, ms, ref_ms):
            a, e = units * canon_unit * MAC_PER_MODMUL / (ms * 1e-3), units * exec_unit * MAC_PER_MODMUL / (ms * 1e-3)
            return {"kernel": name, "achieved": e / 1e9, "frac": e / peak["imad_wide"],
                    "frac_of_carry_chain_peak": (e / peak["imad_wide_x"]) if peak["imad_wide_x"] else None,
                    "canonical_gmac_per_s": a / 1e9, "canonical_frac": a / peak["imad_wide"],
                    "kernel_ms": ms, "kernel_share_of_step": ms / ref_ms, "units_per_launch_total": units, "unit_is": unit_is,
                    "modmul_per_unit": exec_unit, "canonical_modmul_per_unit": canon_unit}

        roof = {"bound": "int_pipe", "peak": peak["imad_wide"] / 1e9, "unit": "G wide-MAC/s (32x32->64)", "peak_source": peak["source"],
                "peak_carry_chain": (peak["imad_wide_x"] or 0) / 1e9, "mac_per_modmul": MAC_PER_MODMUL, "traffic": None, "algorithmic_bytes": None}
        if settled_by_shortcut and len(short_ms) == args.steps:
            # no decode at all - compress(G * p_k) == C_k per coefficient.  x halves (k_fd_coefpoint): fixed-base multiplication (ceil(256 / B) - 1 mixed
            # additions over the signed B-bit-window table; canonical: 33 over byte windows) + to-Montgomery + x_C * Z; sign halves (k_fd_coefsign): batches of 8 with one inversion (binary extended Euclid:
            # ALU work + 2 products) - per point 3 products of the simultaneous inversion + Y / Z + the canonical form for the sign
            sp = [statistics.mean(p_[i] for p_ in short_ms) for i in range(4)]
            pt_canon, pt_exec = CANON_FIX_MADDS * 11 + 2, EXEC_FIX_MADDS * EXEC_MADD + 2
            sg = 2 / 8 + 3 + 2
            top = kernel_entry("k_fd_coefpoint", rows * t, "one coefficient: G * p_k by the fixed-base table, x_C * Z == X against the "
                               "compressed commitment", pt_canon, pt_exec, sp[1], step_mean)
            roof.update({k_: top[k_] for k_ in ("kernel", "achieved", "frac", "frac_of_carry_chain_peak", "canonical_gmac_per_s", "canonical_frac",
                                                "kernel_ms", "kernel_share_of_step", "unit_is", "modmul_per_unit", "canonical_modmul_per_unit")})
            roof["units_per_launch"] = top["units_per_launch_total"]
            roof["shortcut_kernels"] = [
                top,
                kernel_entry("k_fd_coefsign", rows * t, "one coefficient: sign of Y / Z (simultaneous inversion over 8, inversion by binary extended Euclid)",
                             sg, sg, sp[2], step_mean),
                {"kernel": "k_fd_prep + k_fd_cols + k_fd_share_limbs + k_fd_difftab", "kernel_ms": sp[0], "kernel_share_of_step": sp[0] / step_mean,
                 "unit_is": "one dealer: t rounds of subtractions over the n shares + basis conversion by small integers (Fr, no wide products to speak of)"},
                {"kernel": "k_fd_need + k_fd_fill_ok", "kernel_ms": sp[3], "kernel_share_of_step": sp[3] / step_mean},
                {"kernel": "pack + all-gather + flag read-back (the rest of the step)", "kernel_ms": step_mean - sum(sp),
                 "kernel_share_of_step": (step_mean - sum(sp)) / step_mean}]
            roof["algorithmic_bytes"] = rows * t * (48 + 32 + 96)  # commitment + coefficient in, Y and Z planes out
            # ncu --set full of k_fd_coefpoint at 699 392 coefficients, dram__bytes_read.sum + dram__bytes_write.sum per launch:
            #   26-bit windows (profiles/r2_default_path_26bit.md): 1 270.76 MB + 79.03 MB - the 32 GB table does not fit any cache, every one of
            #     the 9 table entries of a coefficient (96 B, random) comes from DRAM in 64-byte pieces: the price of 9 instead of 21 additions;
            #   13-bit windows (profiles/r2_default_path.md): 71.46 MB + 44.39 MB (the 15.7 MB table stays in L2)
            if gtab_bits == 26:
                per_coef, ref_file = (1270.76e6 + 79.032832e6) / 699392, "profiles/r2_default_path_26bit.md"
            elif gtab_bits <= 16:
                per_coef, ref_file = (71463936 + 44385280) / 699392, "profiles/r2_default_path.md (13-bit table, L2-resident)"
            else:
                per_coef, ref_file = None, None
            if per_coef:
                roof["traffic"] = int(round(per_coef * rows * t))
                roof["traffic_ref"] = (f"{ref_file}: dram__bytes_read.sum + dram__bytes_write.sum of one k_fd_coefpoint launch "
                                       f"({per_coef:.0f} B per coefficient, scaled to this launch's coefficients)")
            else:
                roof["traffic_ref"] = f"no ncu capture of k_fd_coefpoint with a {gtab_bits}-bit table (captured: 13 and 26 bits)"
        if "full_evaluation" in legs:
            plan = dk.share_fd_plan(t, n, args.parts)
            m_parts, h_part = plan["parts"], plan["h"]
            seeds = range(plan["lo"], plan["hi"] + 1)
            ph = [statistics.mean(p[i] for p in phase_ms) for i in range(4)]
            ref = statistics.mean(serial_ms)
            # 128 shared doublings (GLV), per point: table 3P/5P/7P (44) + 2 x 128/5 additions + 128/5 products by beta; + part 0; + G*s, compare
            comb_canon = (128 * 8 + (m_parts - 1) * (44 + 52 * 12 + 26) + 12 if m_parts > 1 else 0) + CANON_FIX_MADDS * 11 + 4
            comb_exec = ((128 * EXEC_DBL + (m_parts - 1) * (EXEC_DBL + 3 * EXEC_ADD + 52 * EXEC_ADD + 26) + EXEC_ADD if m_parts > 1 else 0)
                         + EXEC_FIX_MADDS * EXEC_MADD + 4)
            kernels = [
                kernel_entry("k_fd_seed", rows * m_parts * h_part, "one Horner evaluation of a part (dealer, part, seed point)",
                             sum(canonical_horner_modmul(h_part, x) for x in seeds) / h_part,
                             sum(executed_horner_modmul(h_part, x) for x in seeds) / h_part, ph[0], ref),
                kernel_entry("k_fd_init", rows * m_parts * h_part * (h_part - 1) // 2, "one point subtraction (all rounds)", 12, EXEC_ADD, ph[1], ref),
                kernel_entry("k_fd_ext", rows * m_parts * plan["steps"] * (h_part - 1), "one point addition (all ticks)", 12, EXEC_ADD, ph[2], ref),
                kernel_entry("k_fd_combine", rows * n, "one share: joint GLV / width-4 double-and-add over the parts, G*s, compare",
                             comb_canon, comb_exec, ph[3], ref),
            ]
            path_canon = rows * n * MODMUL_PER_SHARE * MAC_PER_MODMUL / (statistics.mean(full_ms) * 1e-3)
            roof["kernels"] = kernels
            roof["evaluation_top_kernel"] = max(kernels, key=lambda k_: k_["kernel_ms"])
            if "kernel" not in roof:
                roof.update({k_: roof["evaluation_top_kernel"][k_] for k_ in ("kernel", "achieved", "frac", "kernel_ms", "kernel_share_of_step", "unit_is")})
            roof["whole_path"] = {"canonical_gmac_per_s": path_canon / 1e9, "frac_of_peak": path_canon / peak["imad_wide"],
                                  "modmul_per_share_canonical": MODMUL_PER_SHARE,
                                  "note": "canonical per-share Horner work of all verified shares / full-evaluation step time; finite "
                                          "differences execute fewer products than that, so this exceeds the kernels' own utilisation"}
            roof["fdiff"] = {"parts_per_dealer": plan["parts"], "coefficients_per_part": plan["h"], "seed_points": [plan["lo"], plan["hi"]],
                             "extension_steps": plan["steps"],
                             "phase_ms": {"seed_horner": ph[0], "differences": ph[1], "extension": ph[2], "recombine_gs_compare": ph[3]},
                             "serial_step_ms": ref, "overlapped_step_ms": full_ms_step,
                             "modmul_per_dealer_fdiff": plan["modmul_fd"], "modmul_per_dealer_horner": plan["modmul_horner"]}
        roof["note"] = ("`kernel` ... `modmul_per_unit`: the dominant kernel of the timed (default-path) steps, timed live inside them, all of them under "
                        "`shortcut_kernels`; `kernels` / `evaluation_top_kernel`: the evaluation kernels (the steps behind `full_evaluation`, run "
                        "phase after phase).  `achieved` / `frac` = the wide MACs the kernel EXECUTES per launch (`modmul_per_unit` Montgomery products of 300 MACs, "
                        "fused sums of two products counted as 444) / its launch time, against the measured carry-free IMAD.WIDE.U32 peak; "
                        "`frac_of_carry_chain_peak`: against the measured rate of the carry-chained IMAD.WIDE.U32.X the product is built from; "
                        "`canonical_*`: SURVEY 8(d)'s textbook count for the same result (33 additions of 11 products per fixed-base multiplication, "
                        "12 per projective addition) - above 1 where the table / formulas do less work than the textbook, so not a utilisation.  "
                        "Kernel times: CUDA events of the library inside the synchronous calls of `sync_call` (same kernels, same inputs as the "
                        "pipelined headline steps, where two lanes may overlap them)")
        line = {
            "metric": METRIC, "value": value, "unit": "shares/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (381-bit Montgomery Fp, 255-bit Fr)", "data": "synthetic",
            "config": {"workload": f"synthetic DKG n={n}, t={t}: full {n}x{n} share-matrix verification, dealer row blocks over {world} GPU(s)",
                       "n": n, "t": t, "shares_per_step": shares_total, "l2": "flushed (256 MB fill) between timed iterations",
                       "share_path": ("consistency shortcut (range, t-th differences of the shares, compress(G*p_k) == C_k per coefficient) settled every "
                                      "dealer: no evaluation in the exponent, no commitment decoded" if settled_by_shortcut else "evaluation"),
                       "fixed_base_table": (f"{gtab_bits}-bit signed odd windows: {EXEC_FIX_MADDS} mixed additions per G * s, "
                                            f"{(EXEC_FIX_MADDS + 1) * (1 << (gtab_bits - 1)) * 96 / 1e6:.0f} MB per ctx"),
                       "parallelism": (f"row-block x{world}; inside dkgv_share_matrix_verify_sharded_dev: one NCCL all-gather of the verdict bitmask + job flags "
                                       f"({chunk * 4} B per rank), one host synchronisation") if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": shares_total / e2e_s_per_step, "unit": "shares/s",
                    "h2d_bytes_per_step": int(h_vv.numel() + h_sh.numel() + h_ids.numel() * 4) * world,
                    "d2h_bytes_per_step": int(h_st.numel() + (h_gather.numel() * 4 + 8 * world if (world > 1 or e2e_pipe) else 0)) * world,
                    "timing": ("host wall clock around K ceremonies through dkgv_share_matrix_enqueue_sharded (pinned host rows in; status bytes, the gathered "
                               "bitmask and the flags back on the host), alternating over the lanes, dkgv_sync, settle; max over ranks" if e2e_pipe else
                               "host wall clock, max over ranks: pinned host buffers in, verdicts (and the gathered bitmask) back on the host"),
                    "sync_call_value": shares_total / e2e_sync_s_per_step,
                    "sync_call": "one synchronous call per ceremony (N = 1: dkgv_share_matrix_verify), host wall clock per call"},
            "gpu_launches": int(launches), "collectives_in_library": world > 1,
            "roofline": roof,
            "parity": {"bad_verdict_bits_device": bad, "bad_verdicts_e2e": bad_e2e, "expected": 0},
            "wall_s_timed_region": wall_pipe if pipe else wall,
            "sync_call": sync_call,
        }
        if pipe:
            line["config"]["l2"] = (f"inputs larger than L2: the K ceremonies read their verification vectors and shares from a ring of {pipe['ring']} distinct "
                                    f"device copies ({pipe['ring_bytes'] / 1e6:.0f} MB > 2 x 126 MB L2), L2 flushed (256 MB fill) once before the timed region")
            line["config"]["pipeline"] = (f"K = {args.steps} ceremonies queued back to back through dkgv_share_matrix_enqueue_sharded_dev over {pipe['lanes']} "
                                          f"ctx / stream lane(s) per GPU, ONE host synchronisation, then dkgv_share_matrix_settle_sharded_dev for each; CUDA events "
                                          f"from before the first enqueue to after the last settle, max over ranks")
            line["gpu_launches"] = int(pipe["launches"])
            line["parity"]["bad_verdict_bits_pipelined"] = pipe["bad"]
            line["pipeline_repeats_ms_per_step"] = pipe["repeat_ms_per_step"]
        if args.emulate_world and world == 1:
            line["emulated"] = (f"ONE rank's row block of a {split}-rank job ({rows} dealers) on one GPU, no collective: value is NOT a whole-ceremony "
                                f"number; ms_per_step is the per-rank step time to expect at N = {split}")
        line.update(legs)
        if not args.no_cpu and not args.quick:
            import oracle_lib as O
            O.use_native_build()
            threads = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_baseline(n, t, args.cpu_sample or 24, threads)
            line["cpu_baseline"]["same_algorithm"] = cpu_same_algorithm(n, t, threads)
            line["pairing"]["cpu_baseline"] = cpu_pairing(threads)
            line["config_a"]["cpu_baseline"] = cpu_config_a(threads)
            line["finalization"]["cpu_baseline"] = cpu_finalization(fin_a, threads)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)  # anything printed during teardown stays out of stdout as well
    if world > 1:
        dist.barrier()
    v.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
