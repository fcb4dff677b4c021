"""-m gpu: the PTX field routines that host emulation cannot reach.  bench/dfma_mul_bench compares, on the GPU,
  * mul2add(a, b, c, d) (one interleaved Montgomery reduction for ab + cd, csrc/field.cuh) and its subtraction form, and
  * mul_dfma(a, b) (the FP64-pipe product of csrc/field_dfma.cuh, an experiment that is not wired into the kernels)
with the carry-chain product mul() on 19.4 M random and extreme operand sets each; mul() itself is pinned on the oracle by every
other GPU test (golden vectors, reference KATs)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_and_fp64_products_are_bit_exact():
    exe = os.path.join(ROOT, "bench", "dfma_mul_bench")
    if not os.path.exists(exe):
        pytest.skip("bench/dfma_mul_bench not built (python __graft_entry__.py)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, check=True).stdout
    j = json.loads(out)
    assert j["status"] == "no error"
    assert j["mul2add_checked"] > 10_000_000 and j["mul2add_mismatches"] == 0
    assert j["checked_pairs"] > 10_000_000 and j["mismatches"] == 0
