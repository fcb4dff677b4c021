"""CPU-only checks of bench.py: the work-per-unit arithmetic of SURVEY.md 8(d), the reference arm's JSON line on a tiny ceremony and
the CPU legs that separate algorithmic from hardware gain (they are the only code of the bench that may touch oracle/)."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_canonical_work_per_share():
    b = load_bench()
    assert round(b.canonical_modmul_per_share(1024, 683)) == 84314
    assert round(b.canonical_modmul_per_share(64, 43)) == 3220       # config A
    assert b.canonical_horner_modmul(10, 0) == 0 and b.canonical_horner_modmul(10, 1) == 9 * 11
    assert b.executed_horner_modmul(10, 3) <= b.canonical_horner_modmul(10, 3) + 9 * 2


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "16", "--t", "5", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "4"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in j, key
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["e2e"]["h2d_bytes_per_step"] == 0


def test_cpu_legs_on_a_small_ceremony():
    b = load_bench()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    same = b.cpu_same_algorithm(24, 7, 2)
    assert same["value"] > 0 and same["cores"] == 2
    base = b.cpu_baseline(24, 7, 4, 2)
    assert base["value"] > 0 and base["fast_mode_value"] > base["value"]
    pr = b.cpu_pairing(2, per_thread=1)
    assert pr["value"] > 0 and pr["modmul_per_check_reference_sequence"] > 20000
