#!/usr/bin/env python
"""bench.py - verified shares/s on the synthetic DKG ceremony n=1024, t=683 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole share matrix through `dkgv_share_matrix_verify[_dev]` (decode of the
verification vectors, verdict for every (dealer, recipient) share).  With N ranks the ONE ceremony is sharded by dealer row
blocks (strong scaling, as BASELINE.json names it); the only exchange is an NCCL all-gather of the per-share verdict bitmask,
inside the timed region.

`value`     device-timed, inputs resident in HBM, max over ranks, the library's DEFAULT path on the synthetic (honest) ceremony:
            the consistency shortcut proves every dealer's shares valid by scalar arithmetic + t fixed-base
            multiplications compared with the COMPRESSED commitments, exactly (no randomness): no share is evaluated in the
            exponent and no commitment is decompressed.
`full_evaluation`  the same steps with the shortcut off - every share through the group arithmetic (finite differences);
`mixed_items`      the same matrix with half of the shares corrupted (BASELINE config 5): the shortcut fails for every dealer
            and the evaluation produces the per-share verdicts, which must flag exactly the corrupted shares.
`e2e`       `value`'s metric through the reference-facing C ABI with HOST (pinned) buffers:
            H2D of vv + shares + ids and D2H of the verdicts inside the timed region.
`roofline`  integer-pipe roofline of the evaluation kernels (`roofline.kernels`, from the full-evaluation steps run phase after
            phase): canonical 32x32->64 multiply-accumulates per second (84 314 modmul/share x 300 MAC)
            against the IMAD.WIDE peak measured live by bench/imad_peak on the same GPU.  The top-level fields describe the
            dominant kernel of the timed (default-path) steps - k_fd_coefpoint, G * p_k against the compressed commitment (the
            default path decodes no commitment) - and `roofline.shortcut_kernels` all kernels of such a step, timed live.
`cpu_baseline` the CPU oracle in reference-faithful mode (per-op affine round trips, constant-time
            255-step scalar multiplication - the reference's operation sequence) on a bounded sample
            of the same matrix, all host cores.  A restatement, not the Rust binary (no cargo here).
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


def canonical_modmul_per_share(n, t):
    """Horner, left-to-right binary [j]y (double 8, add 12, mixed add 11), + 356 for
    G*s and the comparison; averaged over ids 1..n.  (n=1024, t=683) -> 84 314."""
    tot = sum(8 * (j.bit_length() - 1) + 12 * (bin(j).count("1") - 1) + 11 for j in range(1, n + 1))
    return (t - 1) * tot / n + 356


def executed_modmul_per_share(n, t):
    """what k_share_verify really executes: signed-digit chain (vm.cuh make_small_chain), full add (12)
    for the coefficient, 33 mixed adds (11) + 4 for G*s and the comparison."""
    def cost(pos, neg):
        m = pos | neg
        return 8 * (m.bit_length() - 1) + 12 * (bin(m).count("1") - 1)
    return sum(executed_horner_modmul(t, j) for j in range(1, n + 1)) / n + 33 * EXEC_MADD + 4


def canonical_horner_modmul(t, x):
    """SURVEY.md 8(d) W_eval for one evaluation at |x| (binary chain, mixed add); f(0) = C_0 is free"""
    x = abs(x)
    return 0 if x == 0 else (t - 1) * (8 * (x.bit_length() - 1) + 12 * (bin(x).count("1") - 1) + 11)


# executed product-equivalents (300 wide MACs each) of the vm.cuh formulas: the fused sum-of-two-products routine
# (mul2add, 444 MACs) replaces three pairs in the additions and one in the doubling
EXEC_ADD, EXEC_DBL, EXEC_MADD = 6 + 3 * 444 / 300, 6 + 444 / 300, 5 + 3 * 444 / 300


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


PAPER_PEAK_MAC = 148 * 64 * 1.965e9
METRIC = "verified shares/sec (n=1024,t=683)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_PART, help="participants (default: the BASELINE config)")
    ap.add_argument("--t", type=int, default=THRESH)
    ap.add_argument("--cpu-sample", type=int, default=0, help="recipient ids per host thread in the CPU-baseline sample (0 = 24)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-shortcut", dest="shortcut", action="store_false",
                    help="evaluate every id in the group even when the scalar-side consistency conditions hold (dkgv_set_share_shortcut 0)")
    ap.add_argument("--no-finalization", action="store_true", help="skip the config-4 finalization leg")
    ap.add_argument("--no-peak", action="store_true", help="do not run bench/imad_peak (use the paper peak); for runs under ncu")
    ap.add_argument("--overlap", type=int, default=1, choices=[0, 1], help="dkgv_set_share_overlap mode of the evaluation steps")
    ap.add_argument("--parts", type=int, default=0, help="parts per dealer polynomial on the finite-difference path (0 = planner)")
    ap.add_argument("--share-path", default="auto", choices=["auto", "horner", "fdiff"],
                    help="evaluation strategy (enum dkgv_share_path); auto = finite differences for ids 1..n, n > t")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stamps, self.proc = gpu_index, [], [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])
            self.stamps.append(time.perf_counter())

    def stop(self, since=None):
        """summary of the samples taken after perf_counter() time `since` (None: all of them)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        if since is not None:
            n_rows = min(len(self.rows), len(self.stamps))
            self.rows = [self.rows[i] for i in range(n_rows) if self.stamps[i] >= since]
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def cpu_baseline(n, t, cols, threads):
    """oracle in reference-faithful mode on a bounded sample of the same matrix (threads x cols shares:
    one dealer row per host thread, `cols` recipient ids spread over 1..n).  The sample rows reuse one
    dealer's polynomial - the faithful cost per share depends on t only - and the verification vector is
    decoded once per row, which under-counts the reference (it decodes all t points per share,
    verification.rs:132-137): the baseline is favoured, not the GPU."""
    import numpy as np
    import oracle_lib as O
    from dvt_circuits_b200 import synthetic
    rows = max(1, threads)
    coeffs = synthetic.make_coefficients(1, t)
    ids = np.unique(np.linspace(1, n, cols).astype(np.uint32))
    cols = len(ids)
    vv1 = np.zeros((t, 48), dtype=np.uint8)
    cs = [int.from_bytes(coeffs[0, k].tobytes(), "big") for k in range(t)]
    for k in range(t):
        st, pk = O.g1_fixed_base(coeffs[0, k].tobytes(), O.FAST)
        vv1[k] = np.frombuffer(pk, dtype=np.uint8)
    sh1 = np.zeros((cols, 32), dtype=np.uint8)
    for j, i in enumerate(ids.tolist()):
        acc = 0
        for c in reversed(cs):
            acc = (acc * i + c) % synthetic.R_INT
        sh1[j] = np.frombuffer(acc.to_bytes(32, "big"), dtype=np.uint8)
    vv = np.broadcast_to(vv1, (rows, t, 48)).copy()
    shares = np.broadcast_to(sh1, (rows, cols, 32)).copy()
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


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric.  The Rust crate cannot be built
    in this image (no cargo/rustc, un-vendored git dependencies), so this times the C++ oracle in
    reference-faithful mode on all host cores, one bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
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
    from dvt_circuits_b200 import synthetic

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
    if n % world:
        raise SystemExit("participants must divide by the number of ranks")
    rows = n // world

    v = dk.Verifier(local)
    v.set_share_parts(args.parts)
    v.set_share_overlap(args.overlap)
    v.set_share_shortcut(args.shortcut)
    v.set_share_path({"auto": v.PATH_AUTO, "horner": v.PATH_HORNER, "fdiff": v.PATH_FDIFF}[args.share_path])
    sess = synthetic.make_session(v, rows, n, t, dealer_offset=rank * rows)  # set-up, untimed
    ts = torch.cuda.Stream(device=dev)
    stream = ts.cuda_stream
    with torch.cuda.stream(ts):
        d_vv = torch.from_numpy(sess["vv"]).to(dev)
        d_ids = torch.from_numpy(sess["ids"].view(np.int32)).to(dev)
        d_sh = torch.from_numpy(sess["shares"]).to(dev)
        d_st = torch.empty((rows, n), dtype=torch.uint8, device=dev)
        # ranks exchange the verdict BITMASK (1 bit per share), not the status bytes
        d_bits = torch.zeros(((rows * n + 31) // 32,), dtype=torch.int32, device=dev)
        d_all = torch.zeros((world * d_bits.numel(),), dtype=torch.int32, device=dev) if world > 1 else d_st
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # pinned host copies for the end-to-end leg
    h_vv = torch.from_numpy(sess["vv"]).pin_memory()
    h_ids = torch.from_numpy(sess["ids"].view(np.int32)).pin_memory()
    h_sh = torch.from_numpy(sess["shares"]).pin_memory()
    h_st = torch.empty((rows, n), dtype=torch.uint8).pin_memory()

    def step_device():
        v.share_matrix_verify_dev(rows, n, t, d_vv.data_ptr(), d_ids.data_ptr(), d_sh.data_ptr(), d_st.data_ptr(), stream)
        if world > 1:
            v.pack_verdicts_dev(rows * n, d_st.data_ptr(), d_bits.data_ptr(), stream)
            dist.all_gather_into_tensor(d_all, d_bits)

    def step_e2e():
        rc = v._lib.dkgv_share_matrix_verify(v._h, rows, n, t, h_vv.data_ptr(), h_ids.data_ptr(), h_sh.data_ptr(), h_st.data_ptr())
        v._ck(rc)
        if world > 1:
            d_st.copy_(h_st, non_blocking=True)
            v.pack_verdicts_dev(rows * n, d_st.data_ptr(), d_bits.data_ptr(), stream)
            dist.all_gather_into_tensor(d_all, d_bits)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak = None
    if rank == 0:
        peak = measured_int_peak() if not args.no_peak else {"imad_wide": PAPER_PEAK_MAC, "imad_wide_x": None, "fp_mul_per_s": None,
                                                            "source": "paper peak 148 SM x 64 lanes x 1.965 GHz (--no-peak)"}

    with torch.cuda.stream(ts):
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()  # before the warm-up: nvidia-smi needs a moment before its first sample; only samples from wall0 on count
        for _ in range(args.warmup):
            step_device()
        barrier()
        launches0 = v.launch_count
        step_ms, hot_ms, phase_ms, dec_ms, short_ms, decoded_steps = [], [], [], [], [], 0
        barrier()
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)  # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            step_device()
            e1.record(ts)
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            hot_ms.append(v.last_hot_kernel_ms())
            # the default path settles an honest ceremony against the COMPRESSED commitments: no decode inside such a step
            step_decoded = bool(getattr(v, "last_share_decoded", 1))
            decoded_steps += step_decoded
            if step_decoded:
                dec_ms.append(v.last_decode_ms())
            if v.last_share_path == v.PATH_FDIFF:
                try:  # shortcut-settled step: [share limbs + difference table, x halves of G*p_k == C_k, sign halves, flags]
                    short_ms.append(v.last_share_phases_ms())
                except Exception:  # noqa: BLE001  (timing detail only)
                    pass
        barrier()
        wall = time.perf_counter() - wall0
        launches = v.launch_count - launches0
        # The timed region of the default path is short (10 steps of ~15 ms on one GPU, less on eight) against nvidia-smi's 50 ms
        # sampling period: keep the same steps running, untimed, until the GPU has been under this load for ~0.6 s, so that the
        # clocks line rests on several samples.  The count comes from the max-over-ranks step time: identical on every rank.
        t_sum = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_sum, op=dist.ReduceOp.MAX)
        per_step_s = max(float(t_sum.item()) / args.steps * 1e-3, 1e-4)
        extra_steps = max(0, min(2000, int(0.6 / per_step_s) - args.steps))
        for _ in range(extra_steps):
            step_device()
        barrier()
        clocks = sampler.stop(since=wall0) if rank == 0 else None
        if clocks is not None:
            clocks["untimed_steps_under_the_same_load"] = extra_steps
        continued = bool(v.last_share_continued) if v.last_share_path == v.PATH_FDIFF else False
        full_ms, serial_ms = [], []
        if v.last_share_path == v.PATH_FDIFF:
            # The evaluation kernels: the same steps with the consistency shortcut off (every share through the group
            # arithmetic) - first as in production (parts on concurrent streams), then phase after phase on one stream,
            # where single kernels have a duration of their own (roofline.kernels)
            v.set_share_shortcut(False)
            aux_steps = min(args.steps, 3)  # the evaluation legs take ~0.7 s per step
            step_device()  # untimed: the evaluation buffers are allocated on first use
            torch.cuda.synchronize()
            for mode, acc in ((args.overlap, full_ms), (0, serial_ms)):
                v.set_share_overlap(mode)
                for _ in range(aux_steps):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(ts)
                    step_device()
                    e1.record(ts)
                    e1.synchronize()
                    acc.append(e0.elapsed_time(e1))
                    if mode == 0:
                        phase_ms.append(v.last_share_phases_ms())
            v.set_share_overlap(args.overlap)
            v.set_share_shortcut(args.shortcut)
            barrier()
            full_total = torch.tensor([sum(full_ms)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(full_total, op=dist.ReduceOp.MAX)
            full_ms_step = float(full_total.item()) / aux_steps

        total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        ms_per_step = float(total_ms.item()) / args.steps
        bad = int(d_all.count_nonzero().item())

        # end-to-end through the host-pointer C ABI
        step_e2e()
        barrier()
        e2e_t = []
        for _ in range(args.steps):
            barrier()
            t0 = time.perf_counter()
            step_e2e()
            torch.cuda.synchronize()
            e2e_t.append(time.perf_counter() - t0)
        e2e_total = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
        e2e_s_per_step = float(e2e_total.item()) / args.steps
        bad_e2e = int(h_st.count_nonzero().item())

        # BASELINE config 5: the same matrix with half of the shares corrupted (one flipped bit each): every dealer group
        # takes the full evaluation; verdicts must flag exactly the corrupted shares
        rng = np.random.Generator(np.random.PCG64([0xBAD, rank]))
        mask = rng.random((rows, n)) < 0.5
        sh_bad = sess["shares"].copy()
        sh_bad[mask, 31] ^= 1
        d_sh_bad = torch.from_numpy(sh_bad).to(dev)
        mixed_ms = []
        for i in range(1 + max(1, min(args.steps, 3) - 1)):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            v.share_matrix_verify_dev(rows, n, t, d_vv.data_ptr(), d_ids.data_ptr(), d_sh_bad.data_ptr(), d_st.data_ptr(), stream)
            e1.record(ts)
            e1.synchronize()
            if i:
                mixed_ms.append(e0.elapsed_time(e1))
        mixed_ok = bool(((d_st != 0) == torch.from_numpy(mask).to(dev)).all().item())
        mixed_total = torch.tensor([sum(mixed_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(mixed_total, op=dist.ReduceOp.MAX)
        mixed_ms_step = float(mixed_total.item()) / len(mixed_ms)
        del d_sh_bad

        # second BASELINE metric: BLS pairing checks/s (bls_verify_precomputed_hash, one common message),
        # 262144 checks (a batch that saturates the GPU) sharded over the ranks, inputs resident in HBM, verdict bytes all-gathered
        m_total = 262144
        m_loc = m_total // world
        fin = synthetic.make_finalization(v, 64, 8)
        reps = (m_loc + 63) // 64
        d_pk = torch.from_numpy(np.tile(fin["partial_pubkeys"], (reps, 1))[:m_loc].copy()).to(dev)
        d_sg = torch.from_numpy(np.tile(fin["signatures"], (reps, 1))[:m_loc].copy()).to(dev)
        d_hm = torch.from_numpy(fin["hm"].copy()).to(dev)
        d_ps = torch.empty((m_loc,), dtype=torch.uint8, device=dev)
        d_pall = torch.empty((m_total,), dtype=torch.uint8, device=dev) if world > 1 else d_ps

        def step_pairing():
            v._ck(v._lib.dkgv_bls_verify_batch_dev(v._h, m_loc, d_pk.data_ptr(), d_sg.data_ptr(), 1, d_hm.data_ptr(), None,
                                                   d_ps.data_ptr(), stream))
            if world > 1:
                dist.all_gather_into_tensor(d_pall, d_ps)

        step_pairing()
        barrier()
        pair_ms = []
        for _ in range(max(2, args.steps)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            step_pairing()
            e1.record(ts)
            e1.synchronize()
            pair_ms.append(e0.elapsed_time(e1))
        pair_total = torch.tensor([sum(pair_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pair_total, op=dist.ReduceOp.MAX)
        pair_ms_step = float(pair_total.item()) / len(pair_ms)
        pair_bad = int(d_pall.count_nonzero().item())

        # BASELINE config 4: finalization of the whole ceremony through the host-pointer C ABI (rank 0, one GPU):
        # agg_coefficients + the n final keys, two Lagrange interpolations at 0, n partial-signature checks
        fin_line = None
        if rank == 0 and not args.no_finalization:
            ff = synthetic.make_finalization(v, n, t)
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                ast, co, keys = v.agg_final_keys(ff["vv"], ff["ids"])
                t1 = time.perf_counter()
                l1 = v.lagrange_at_zero(keys, ff["ids"])
                l2 = v.lagrange_at_zero(ff["partial_pubkeys"], ff["ids"])
                t2 = time.perf_counter()
                st_f = v.bls_verify_batch(ff["partial_pubkeys"], ff["signatures"], ff["hm"])
                t3 = time.perf_counter()
                ok = bool(ast == 0 and (keys == ff["partial_pubkeys"]).all() and l1 == (0, bytes(co[0])) and l2 == l1 and not st_f.any())
                cur = {"metric": "finalization of one ceremony (host buffers, wall clock)", "n": n, "t": t, "total_ms": (t3 - t0) * 1e3,
                       "agg_final_keys_ms": (t1 - t0) * 1e3, "lagrange_x2_ms": (t2 - t1) * 1e3, "partial_signature_checks_ms": (t3 - t2) * 1e3,
                       "final_keys_per_s": n / (t1 - t0), "all_checks_hold": ok}
                if best is None or cur["total_ms"] < best["total_ms"]:
                    best = cur
            fin_line = best

    if rank == 0:
        shares = n * n
        value = shares / (ms_per_step * 1e-3)
        hot = statistics.mean(hot_ms)
        MODMUL_PER_SHARE = canonical_modmul_per_share(n, t)
        fdiff = v.last_share_path == v.PATH_FDIFF
        step_mean = statistics.mean(step_ms)
        # whole path: canonical per-share work (84 314 modmul) of every verified share per second of step time
        path_canon = rows * n * MODMUL_PER_SHARE * MAC_PER_MODMUL / ((statistics.mean(full_ms) if full_ms else step_mean) * 1e-3)

        def kernel_entry(name, units, unit_is, canon_unit, exec_unit, ms, ref_ms):
            a, e = units * canon_unit * MAC_PER_MODMUL / (ms * 1e-3), units * exec_unit * MAC_PER_MODMUL / (ms * 1e-3)
            return {"kernel": name, "achieved": a / 1e9, "frac": a / peak["imad_wide"],
                    "frac_of_carry_chain_peak": (a / peak["imad_wide_x"]) if peak["imad_wide_x"] else None,
                    "executed_gmac_per_s": e / 1e9,
                    "executed_frac_of_carry_chain_peak": (e / peak["imad_wide_x"]) if peak["imad_wide_x"] else None,
                    "kernel_ms": ms, "kernel_share_of_step": ms / ref_ms, "units_per_launch_total": units, "unit_is": unit_is,
                    "modmul_per_unit": canon_unit, "executed_modmul_per_unit": exec_unit}

        if fdiff:
            plan = dk.share_fd_plan(t, n, args.parts)
            m_parts, h_part = plan["parts"], plan["h"]
            seeds = range(plan["lo"], plan["hi"] + 1)
            ph = [statistics.mean(p[i] for p in phase_ms) for i in range(4)]
            ref = statistics.mean(serial_ms)
            # 128 shared doublings (GLV), per point: table 3P/5P/7P (44) + 2 x 128/5 additions + 128/5 products by beta; + part 0; + G*s, compare
            comb_canon = (128 * 8 + (m_parts - 1) * (44 + 52 * 12 + 26) + 12 if m_parts > 1 else 0) + 33 * 11 + 4
            comb_exec = ((128 * EXEC_DBL + (m_parts - 1) * (EXEC_DBL + 3 * EXEC_ADD + 52 * EXEC_ADD + 26) + EXEC_ADD if m_parts > 1 else 0)
                         + 33 * EXEC_MADD + 4)
            short = bool(args.shortcut and n > t and not continued)  # the timed steps were settled by the consistency shortcut
            ids_eval, ext_steps = n, plan["steps"]  # the kernel entries below come from the full-evaluation steps
            kernels = [
                kernel_entry("k_fd_seed", rows * m_parts * h_part, "one Horner evaluation of a part (dealer, part, seed point)",
                             sum(canonical_horner_modmul(h_part, x) for x in seeds) / h_part,
                             sum(executed_horner_modmul(h_part, x) for x in seeds) / h_part, ph[0], ref),
                kernel_entry("k_fd_init", rows * m_parts * h_part * (h_part - 1) // 2, "one point subtraction (all rounds)", 12, EXEC_ADD, ph[1], ref),
                kernel_entry("k_fd_ext", rows * m_parts * ext_steps * (h_part - 1), "one point addition (all ticks)", 12, EXEC_ADD, ph[2], ref),
                kernel_entry("k_fd_combine", rows * ids_eval, "one share: joint GLV / width-4 double-and-add over the parts, G*s, compare",
                             comb_canon, comb_exec, ph[3], ref),
            ]
            top = max(kernels, key=lambda k_: k_["kernel_ms"])
            algo_bytes = rows * t * 100 + rows * m_parts * h_part * 144
            short_kernels = []
            if short and decoded_steps == args.steps and dec_ms:
                # (DKGV_FD_BYTES=0) dominant kernel of the shortcut path: the lazy decode of the commitments - flags, x < p, square root, curve equation
                dms = statistics.mean(d_[0] for d_ in dec_ms)
                dec_canon = 379 + 228 + 4  # a^((p+1)/4) by square-and-multiply + x^3 + 4, y^2 check
                dec = kernel_entry("k_decompress_vv (no subgroup test)", rows * t, "one commitment: decompression without the subgroup test",
                                   dec_canon, dec_canon, dms, step_mean)
                dec["subgroup_checked"] = bool(dec_ms[-1][1])
            elif short and len(short_ms) == args.steps:
                # default: no decode at all - compress(G * p_k) == C_k per coefficient.  x halves (k_fd_coefpoint): fixed-base
                # multiplication (33 mixed additions) + to-Montgomery + x_C * Z; sign halves (k_fd_coefsign): batches of 8 with
                # one inversion (p - 2: 380 squarings + 227 products) - per point 607 / 8 + 3 products of the simultaneous
                # inversion + Y / Z + the canonical form for the sign
                sp = [statistics.mean(p_[i] for p_ in short_ms) for i in range(4)]
                pt_canon, pt_exec = 33 * 11 + 2, 33 * EXEC_MADD + 2
                sg = 607 / 8 + 3 + 2
                dec = kernel_entry("k_fd_coefpoint", rows * t, "one coefficient: G * p_k by the fixed-base table, x_C * Z == X against the "
                                   "compressed commitment", pt_canon, pt_exec, sp[1], step_mean)
                short_kernels = [dec,
                                 kernel_entry("k_fd_coefsign", rows * t, "one coefficient: sign of Y / Z (simultaneous inversion over 8)", sg, sg,
                                              sp[2], step_mean),
                                 {"kernel": "k_fd_share_limbs + k_fd_difftab", "kernel_ms": sp[0], "kernel_share_of_step": sp[0] / step_mean,
                                  "unit_is": "one dealer: t rounds of subtractions over the n shares + basis conversion by small integers (Fr, no wide products to speak of)"},
                                 {"kernel": "k_fd_need + k_fd_fill_ok + flag read-back", "kernel_ms": sp[3], "kernel_share_of_step": sp[3] / step_mean}]
            else:
                short = False
        else:
            plan = None
            kernels = [kernel_entry("k_share_verify", rows * n, "one share", MODMUL_PER_SHARE, executed_modmul_per_share(n, t), hot, step_mean)]
            top = kernels[0]
            algo_bytes = rows * n * (32 + 1) + rows * t * 100 + n * 4  # shares + verdicts + decoded vv + ids
        eval_top = top
        if fdiff and short:
            top = dec  # dominant kernel of the timed (default-path) steps, timed live inside them
        roof = {"bound": "int_pipe", "kernel": top["kernel"], "achieved": top["achieved"], "peak": peak["imad_wide"] / 1e9,
                "unit": "G wide-MAC/s (32x32->64)", "frac": top["frac"], "peak_source": peak["source"],
                "peak_carry_chain": (peak["imad_wide_x"] or 0) / 1e9,
                "frac_of_carry_chain_peak": top["frac_of_carry_chain_peak"],
                "executed_gmac_per_s": top["executed_gmac_per_s"],
                "executed_frac_of_carry_chain_peak": top["executed_frac_of_carry_chain_peak"],
                "kernel_ms": top["kernel_ms"], "kernel_share_of_step": top["kernel_share_of_step"],
                "units_per_launch": top["units_per_launch_total"], "unit_is": top["unit_is"],
                "modmul_per_unit": top["modmul_per_unit"], "mac_per_modmul": MAC_PER_MODMUL,
                "traffic": None, "traffic_ref": "profiles/ (ncu --set full captures: DRAM bytes per launch are negligible on this integer-bound path)",
                "algorithmic_bytes": None,
                "kernels": kernels,
                "whole_path": {"canonical_gmac_per_s": path_canon / 1e9, "frac_of_peak": path_canon / peak["imad_wide"],
                               "modmul_per_share_canonical": MODMUL_PER_SHARE,
                               "note": "canonical per-share Horner work of all verified shares / full-evaluation step time; finite "
                                       "differences execute fewer products than that, so this exceeds the kernels' own utilisation"},
                "hbm": {"algorithmic_bytes_per_launch": algo_bytes, "achieved_gbs": algo_bytes / (eval_top["kernel_ms"] * 1e-3) / 1e9,
                        "note": "integer-bound path: HBM use is a rounding error"}}
        if fdiff and short and short_kernels:
            roof["shortcut_kernels"] = short_kernels
            roof["algorithmic_bytes"] = rows * t * (48 + 32 + 96)  # commitment + coefficient in, Y and Z planes out
            # ncu --set full of k_fd_coefpoint at 699 392 coefficients (profiles/r1_default_path_v3.md): dram read 57 367 296 B +
            # write 32 190 208 B per launch; part of the Y / Z planes stays in L2 for k_fd_coefsign
            roof["traffic"] = int(round((57367296 + 32190208) / 699392 * rows * t))
            roof["traffic_ref"] = ("profiles/r1_default_path_v3.md: dram__bytes_read.sum + dram__bytes_write.sum of one k_fd_coefpoint launch "
                                   "(128.1 B per coefficient, scaled to this launch's coefficients)")
        elif fdiff and short:
            # ncu --set full of k_decompress_vv at 699 392 commitments (profiles/r1_default_path.md): dram read 34 069 760 B +
            # write 21 060 352 B per launch; algorithmic 48 B in + 100 B planar out per commitment (the planes mostly stay in L2)
            roof["traffic"] = int(round((34069760 + 21060352) / 699392 * rows * t))
            roof["algorithmic_bytes"] = rows * t * 148
            roof["traffic_ref"] = ("profiles/r1_default_path.md: dram__bytes_read.sum + dram__bytes_write.sum of one k_decompress_vv launch "
                                   "(78.8 B per commitment, scaled to this launch's commitments)")
        if fdiff:
            roof["note"] = (("the timed steps behind `value` were settled by the consistency shortcut: `kernel` ... `modmul_per_unit` describe their "
                             "dominant kernel (timed live inside them; all of them under `shortcut_kernels`); " if short else "")
                            + "`kernels` / `evaluation_top_kernel` describe the evaluation kernels (every share through the group arithmetic: "
                              "the steps behind `full_evaluation`, run phase after phase)")
            roof["evaluation_top_kernel"] = eval_top
            roof["fdiff"] = {"consistency_shortcut_settled_the_timed_steps": bool(short),
                             "parts_per_dealer": plan["parts"], "coefficients_per_part": plan["h"],
                             "seed_points": [plan["lo"], plan["hi"]], "extension_steps": plan["steps"],
                             "phase_ms": {"seed_horner": ph[0], "differences": ph[1], "extension": ph[2], "recombine_gs_compare": ph[3]},
                             "serial_step_ms": ref, "overlapped_step_ms": full_ms_step,
                             "note": "full evaluation (shortcut off): phase times from steps run phase-after-phase on one stream; "
                                     "`overlapped_step_ms` with the parts on concurrent streams",
                             "modmul_per_dealer_fdiff": plan["modmul_fd"], "modmul_per_dealer_horner": plan["modmul_horner"]}
        line = {
            "metric": METRIC, "value": value, "unit": "shares/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (381-bit Montgomery Fp, 255-bit Fr)", "data": "synthetic",
            "config": {"workload": f"synthetic DKG n={n}, t={t}: full {n}x{n} share-matrix verification, dealer row blocks over {world} GPU(s)",
                       "n": n, "t": t, "shares_per_step": shares, "l2": "flushed (256 MB fill) between timed iterations",
                       "share_path": (("consistency shortcut (range, t-th differences of the shares, G*p_k == C_k per coefficient) settled every dealer; "
                                       if (fdiff and short) else "") +
                                      (f"evaluation by finite differences: {plan['parts']} parts x {plan['h']} coefficients per dealer, "
                                       f"{plan['h']} Horner seeds per part, differences, recombination") if fdiff else "Horner per share"),
                       "parallelism": f"row-block x{world}, NCCL all-gather of the verdict bitmask ({n * n // 8} B)" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": shares / e2e_s_per_step, "unit": "shares/s",
                    "h2d_bytes_per_step": int(h_vv.numel() + h_sh.numel() + h_ids.numel() * 4) * world,
                    "d2h_bytes_per_step": int(h_st.numel()) * world, "timing": "host wall clock around dkgv_share_matrix_verify, max over ranks"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "pairing": {"metric": "BLS pairing checks/sec", "value": m_total / (pair_ms_step * 1e-3), "unit": "checks/s",
                        "checks_per_step": m_total, "ms_per_step": pair_ms_step, "bad_verdicts": pair_bad,
                        "note": "e(pk,H(m)) == e(G1,sig) as 2 Miller loops + 1 final exponentiation per check, incl. G1/G2 decoding with subgroup checks"},
            "full_evaluation": ({"metric": "verified shares/sec, every share evaluated in the group (consistency shortcut off)",
                                 "value": shares / (full_ms_step * 1e-3), "unit": "shares/s", "ms_per_step": full_ms_step} if fdiff else None),
            "mixed_items": {"metric": "verified shares/sec, 50 % of the shares corrupted (BASELINE config 5)", "value": shares / (mixed_ms_step * 1e-3),
                            "unit": "shares/s", "ms_per_step": mixed_ms_step, "verdicts_flag_exactly_the_corrupted_shares": mixed_ok,
                            "note": "every dealer group fails the consistency conditions and takes the full evaluation"},
            "finalization": fin_line,
            "parity": {"bad_verdicts_device": bad, "bad_verdicts_e2e": bad_e2e, "expected": 0},
            "wall_s_timed_region": wall,
        }
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_baseline(n, t, args.cpu_sample or 24, threads)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)  # anything printed during teardown stays out of stdout as well
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    v.close()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
