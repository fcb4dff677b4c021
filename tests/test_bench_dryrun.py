"""CPU dry run of bench.py's B200 arm at world_size 2 over gloo: the WHOLE control flow of the full (not --quick) run - warm-up,
synchronous steps, ceremonies in flight over the lanes, end-to-end legs, evaluation / corruption / pairing / hash-to-G2 /
bad-partial-key / finalization / config-A legs, line assembly, teardown - with stand-ins for the GPU: torch.cuda streams and events
are no-ops, "device" tensors live on the CPU, and the Verifier is a fake that computes nothing but performs a REAL gloo collective
wherever the library performs an NCCL one.  What this catches: a collective that only some ranks reach (the full run at N = 2 once hung
on a barrier inside the rank-0 config-A leg), exceptions in branches only N > 1 takes, and a line that is not the contract's JSON.
It checks no verdict (the fake has none to give) - parity is the business of the -m gpu tests."""
import json
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R_INT = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def _install_standins():
    import contextlib
    import torch
    import torch.distributed as dist
    import dvt_circuits_b200 as dk

    cpu = torch.device("cpu")
    real_init = dist.init_process_group

    class Stream:
        cuda_stream = 0

        def __init__(self, device=None):
            pass

        def synchronize(self):
            pass

        def wait_event(self, ev):
            pass

    class Event:
        def __init__(self, enable_timing=False):
            pass

        def record(self, stream=None):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return 1.0

    torch.cuda.Stream, torch.cuda.Event = Stream, Event
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.Tensor.pin_memory = lambda self: self
    real_device = torch.device
    torch.device = lambda *a, **k: cpu if (a and isinstance(a[0], str) and a[0].startswith("cuda")) else real_device(*a, **k)
    dist.init_process_group = lambda backend, **k: real_init("gloo")

    def collective():  # stands for one NCCL call of the library: blocks until every rank has made the same call
        if dist.is_initialized() and dist.get_world_size() > 1:
            box = [torch.zeros(1) for _ in range(dist.get_world_size())]
            dist.all_gather(box, torch.zeros(1))

    class FakeLib:
        def __getattr__(self, name):
            return lambda *a: 0

    class FakeVerifier:
        PATH_AUTO, PATH_HORNER, PATH_FDIFF = 0, 1, 2

        def __init__(self, device=0, gtab_bits=None):
            self._bits, self._lib, self._h, self.launch_count = gtab_bits or 22, FakeLib(), 1, 0
            self.last_share_path, self.last_share_continued, self.last_share_repaired, self.last_bls_path = self.PATH_FDIFF, 0, 0, 1
            self._world, self._shortcut = 1, True

        def gtab_bits(self):
            return self._bits

        @staticmethod
        def comm_unique_id():
            return bytes(128)

        def comm_init(self, uid, rank, world):
            self._world = world
            collective()

        def _ck(self, rc):
            assert rc == 0

        def set_share_parts(self, p):
            pass

        def set_share_overlap(self, o):
            pass

        def set_share_shortcut(self, on):
            self._shortcut = bool(on)

        def share_gather_words(self, n_local, n_r):
            return (((n_local * n_r + 31) // 32 + 2) + 3) & ~3

        def _step(self):
            self.launch_count += 9
            self.last_share_continued = 0 if self._shortcut else 1

        def share_matrix_verify_sharded_dev(self, *a):
            self._step()
            if self._world > 1:
                collective()

        def share_matrix_enqueue_sharded_dev(self, *a):
            self.share_matrix_verify_sharded_dev()

        share_matrix_enqueue_sharded = share_matrix_enqueue_sharded_dev

        def share_matrix_settle_sharded_dev(self, *a):
            return False

        share_matrix_settle_sharded = share_matrix_settle_sharded_dev

        def share_matrix_verify_dev(self, *a):
            self._step()

        def sync(self):
            pass

        def last_share_phases_ms(self):
            return [0.4, 0.4, 0.1, 0.01]

        def last_bls_kernel_ms(self):
            return 0.8

        def bls_verify_batch_sharded_dev(self, *a):
            if self._world > 1:
                collective()

        def bls_verify_batch(self, pk, sig, hm, hm_idx=None):
            return np.zeros((np.asarray(pk).reshape(-1, 48).shape[0],), dtype=np.uint8)

        def hash_to_g2(self, msgs):
            return [np.zeros(96, dtype=np.uint8) for _ in msgs]

        def g1_fixed_base_mul(self, scalars):
            m = np.asarray(scalars).reshape(-1, 32).shape[0]
            return np.zeros((m, 48), dtype=np.uint8), np.zeros((m,), dtype=np.uint8)

        def g2_mul_batch(self, base96, scalars):
            return np.zeros((np.asarray(scalars).reshape(-1, 32).shape[0], 96), dtype=np.uint8)

        def fr_poly_eval(self, coeffs, ids):
            n_d, t, _ = coeffs.shape
            out = np.zeros((n_d, len(ids), 32), dtype=np.uint8)
            for d in range(n_d):
                c = [int.from_bytes(coeffs[d, k].tobytes(), "big") for k in range(t)]
                for j, x in enumerate(ids):
                    acc = 0
                    for ck in reversed(c):
                        acc = (acc * int(x) + ck) % R_INT
                    out[d, j] = np.frombuffer(acc.to_bytes(32, "big"), dtype=np.uint8)
            return out

        def bad_partial_key_verify_batch(self, vv, perp, pk, sig, msgs, msg_idx=None):
            return np.zeros((len(perp),), dtype=np.uint8), np.zeros((vv.shape[0], 48), dtype=np.uint8), 0

        def pack_verdicts_dev(self, *a):
            pass

        def all_gather_dev(self, *a):
            if self._world > 1:
                collective()

        def agg_final_keys_sharded(self, vv_local, ids):
            if self._world > 1:
                collective()
            return 0, np.zeros((vv_local.shape[1], 48), dtype=np.uint8), np.zeros((len(ids), 48), dtype=np.uint8)

        def lagrange_at_zero(self, pts, ids):
            return 0, bytes(48)

        def close(self):
            pass

    dk.Verifier = FakeVerifier


def _worker(rank, world, port, outdir, argv):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      DKGV_BENCH_SAMPLER="off")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    _install_standins()
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_dry", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    bench.measured_int_peak = lambda: {"imad_wide": bench.PAPER_PEAK_MAC, "imad_wide_x": bench.PAPER_PEAK_MAC / 2, "fp_mul_per_s": None, "source": "dry run"}
    out = open(os.path.join(outdir, f"rank{rank}.out"), "w")
    sys.stdout.flush()
    os.dup2(out.fileno(), 1)
    sys.argv = ["bench.py"] + argv
    rc = bench.main()
    sys.stdout.flush()
    with open(os.path.join(outdir, f"rank{rank}.rc"), "w") as f:
        f.write(str(rc))


def _run(world, argv, timeout):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    outdir = tempfile.mkdtemp(prefix="bench_dry_")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, outdir, argv)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    hung = [p.pid for p in procs if p.is_alive()]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert not hung, f"bench.py did not finish on every rank within {timeout} s (a collective only some ranks reach?)"
    assert [p.exitcode for p in procs] == [0] * world
    for r in range(world):
        assert open(os.path.join(outdir, f"rank{r}.rc")).read() == "0"
    lines = [ln for ln in open(os.path.join(outdir, "rank0.out")).read().splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # exactly ONE JSON line on rank 0's stdout
    for r in range(1, world):
        assert open(os.path.join(outdir, f"rank{r}.out")).read().strip() == ""
    return json.loads(lines[0])


CONTRACT_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                 "config", "clocks", "e2e", "gpu_launches", "roofline"]


@pytest.mark.parametrize("world", [2, 1])
def test_full_bench_control_flow(world):
    line = _run(world, ["--gpus", str(world), "--n", "8", "--t", "3", "--steps", "2", "--warmup", "1", "--no-cpu"], timeout=240)
    for k in CONTRACT_KEYS:
        assert k in line, k
    assert line["n_gpus"] == world and line["steps"] == 2 and line["scaling"] == "strong" and line["higher_is_better"] is True
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["gpu_launches"] > 0 and "workload" in line["config"] and "pipeline" in line["config"]
    for leg in ("sync_call", "full_evaluation", "corruption", "pairing", "hash_to_g2", "bad_partial_key", "finalization", "config_a"):
        assert leg in line, leg
    assert set(line["corruption"]) == {"one_share", "one_dealer", "one_dealer_in_every_group", "p_1pct", "p_10pct", "p_50pct_config5"}
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in line["roofline"], k


def test_quick_bench_control_flow_two_ranks():
    line = _run(2, ["--gpus", "2", "--n", "8", "--t", "3", "--steps", "3", "--warmup", "1", "--quick", "--no-cpu", "--no-peak"], timeout=120)
    assert line["n_gpus"] == 2 and "sync_call" in line and "pairing" not in line


def test_more_steps_than_output_slots_and_a_single_step():
    """K = 300 ceremonies go through the pipelined legs in waves of 256 (the output slots); K = 1, W = 0 uses one lane only"""
    line = _run(2, ["--gpus", "2", "--n", "8", "--t", "3", "--steps", "300", "--warmup", "1", "--quick", "--no-cpu", "--no-peak"], timeout=200)
    assert line["steps"] == 300 and line["gpu_launches"] == 300 * 9
    line = _run(1, ["--n", "8", "--t", "3", "--steps", "1", "--warmup", "0", "--quick", "--no-cpu", "--no-peak"], timeout=100)
    assert line["steps"] == 1 and line["warmup"] == 0 and line["gpu_launches"] == 9


def test_reference_arm_under_two_ranks():
    """--impl reference launched like the B200 arm (torchrun, N = 2): rank 0 alone runs the oracle and prints the one line, rank 1
    exits 0 without work and without touching a process group"""
    line = _run(2, ["--impl", "reference", "--gpus", "2", "--n", "8", "--t", "3", "--steps", "1", "--warmup", "0", "--cpu-sample", "2"], timeout=200)
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["gpu_launches"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
