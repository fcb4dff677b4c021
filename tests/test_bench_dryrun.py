"""CPU-only dry run of bench.py's B200 arm: the control flow and the JSON assembly of `run_b200` with the CUDA pieces of torch
and the verifier replaced by stand-ins (no kernel runs, timings are made up).  It exists because the bench line is the
round's deliverable and most of its Python cannot otherwise be exercised without a GPU: a renamed key, an unbound variable or
a mismatch with the binding's API shows up here.  Numbers produced by this test mean nothing."""
import contextlib
import ctypes
import importlib.util
import io
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 15.0


class FakeStream:
    cuda_stream = 1

    def __init__(self, device=None):
        pass


class FakeLib:
    def dkgv_share_matrix_verify(self, h, rows, n, t, vv, ids, sh, st):
        ctypes.memset(st, 0, rows * n)
        return 0

    def dkgv_bls_verify_batch_dev(self, h, m, pk, sg, one, hm, x, st, stream):
        ctypes.memset(st, 0, m)
        return 0


def make_fake_verifier(real_cls, decode_free):
    class FakeVerifier:
        PATH_AUTO, PATH_HORNER, PATH_FDIFF = real_cls.PATH_AUTO, real_cls.PATH_HORNER, real_cls.PATH_FDIFF

        def __init__(self, device):
            self._lib, self._h = FakeLib(), None
            self.launch_count = 0
            self.shortcut = True
            self.last_share_path = self.PATH_FDIFF
            self.last_share_continued = 0
            self.last_share_decoded = 0

        def _ck(self, rc):
            assert rc == 0

        def set_share_parts(self, p):
            pass

        def set_share_overlap(self, m):
            pass

        def set_share_shortcut(self, on):
            self.shortcut = bool(on)

        def set_share_path(self, m):
            pass

        def share_matrix_verify_dev(self, rows, n, t, vv, ids, sh, st, stream):
            ctypes.memset(st, 0, rows * n)
            self.launch_count += 6
            self.last_share_continued = 0 if self.shortcut else 1
            self.last_share_decoded = 0 if (self.shortcut and decode_free) else 1

        def pack_verdicts_dev(self, n, st, bits, stream):
            pass

        def last_hot_kernel_ms(self):
            return 14.0

        def last_decode_ms(self):
            if decode_free and self.shortcut:
                raise RuntimeError("no verification-vector decode launched yet")
            return 15.0, not self.shortcut

        def last_share_phases_ms(self):
            return [3.5, 8.9, 2.2, 0.04]

        def close(self):
            pass

    if not hasattr(real_cls, "last_share_decoded"):
        pytest.fail("binding lacks last_share_decoded")
    for name in ("last_hot_kernel_ms", "last_decode_ms", "last_share_phases_ms", "share_matrix_verify_dev", "pack_verdicts_dev",
                 "set_share_parts", "set_share_overlap", "set_share_shortcut", "set_share_path", "launch_count", "last_share_continued"):
        assert hasattr(real_cls, name), name  # the stand-in only has what the real binding has
    return FakeVerifier


@pytest.mark.parametrize("decode_free,world", [(True, 1), (False, 1), (True, 2)])
def test_bench_b200_arm_dry_run(monkeypatch, capsys, decode_free, world):
    import torch
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import synthetic

    spec = importlib.util.spec_from_file_location("bench_dry", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    cpu = torch.device("cpu")
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch, "device", lambda s: cpu)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch, "empty", (lambda real: (lambda *a, **k: real(*((min(a[0], 1 << 16),) if a and isinstance(a[0], int) else a), **k)))(torch.empty))
    monkeypatch.setattr(dk, "Verifier", make_fake_verifier(dk.Verifier, decode_free))

    def fake_session(v, n_d, n_r, t, dealer_offset=0, **kw):
        return {"vv": np.zeros((n_d, t, 48), np.uint8), "ids": np.arange(1, n_r + 1, dtype=np.uint32), "shares": np.zeros((n_d, n_r, 32), np.uint8)}

    def fake_final(v, n, t, **kw):
        return {"partial_pubkeys": np.zeros((n, 48), np.uint8), "signatures": np.zeros((n, 96), np.uint8), "hm": np.zeros((96,), np.uint8)}

    monkeypatch.setattr(synthetic, "make_session", fake_session)
    monkeypatch.setattr(synthetic, "make_finalization", fake_final)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)  # no nvidia-smi here: stop() reports it unavailable
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    if world > 1:  # rank 0 of a 2-rank job with the collectives stubbed out: the N > 1 branches of the control flow
        import torch.distributed as dist
        monkeypatch.setenv("WORLD_SIZE", str(world))
        monkeypatch.setenv("RANK", "0")
        monkeypatch.setenv("LOCAL_RANK", "0")
        for name in ("init_process_group", "all_reduce", "all_gather_into_tensor", "barrier", "destroy_process_group"):
            monkeypatch.setattr(dist, name, lambda *a, **k: None)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--n", "64", "--t", "43", "--steps", "3", "--warmup", "1", "--no-cpu", "--no-peak",
                                      "--no-finalization"])
    # run_b200 parks fd 1 on stderr (native banners) and prints the one JSON line through sys.stdout: undo the fd games afterwards
    saved = os.dup(1)
    try:
        rc = bench.main()
        sys.stdout.flush()
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    out = capsys.readouterr().out
    assert rc in (0, None)
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "full_evaluation", "mixed_items", "pairing"):
        assert key in line, key
    assert line["steps"] == 3 and line["n_gpus"] == world and line["gpu_launches"] == 18
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "kernel_ms", "kernel_share_of_step"):
        assert key in roof, key
    if decode_free:
        assert roof["kernel"] == "k_fd_coefpoint" and [k["kernel"] for k in roof["shortcut_kernels"]][:2] == ["k_fd_coefpoint", "k_fd_coefsign"]
        assert roof["traffic"] == int(round((57367296 + 32190208) / 699392 * (64 // world) * 43))
    else:
        assert roof["kernel"].startswith("k_decompress_vv")
    assert line["clocks"]["untimed_steps_under_the_same_load"] == int(0.6 / 0.015) - 3
