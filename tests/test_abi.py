"""CPU-only: the C-ABI library loads without a GPU, exports every symbol include/*.h declares, and the
product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

import dvt_circuits_b200 as dk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in ("dkgv.h", "dkgh.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(dkg[vh]_\w+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    lib = dk.load_library()
    decl = declared_functions()
    assert decl, "no declarations found"
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by libdkgv.so"
    assert decl == set(dk.DECLARED_SYMBOLS), decl ^ set(dk.DECLARED_SYMBOLS)


def test_status_codes_match_header():
    text = open(os.path.join(ROOT, "include", "dkgv.h")).read()
    for name, val in re.findall(r"DKGV_(\w+) = (\d+)", text):
        if name.startswith(("DEC_", "SHARE_PATH_", "BLS_PATH_")):
            continue
        assert dk.Status[name] == int(val)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dk.DkgvError):
        dk.Verifier(0)


def test_product_never_touches_the_oracle():
    """nothing under dvt_circuits_b200/ or bench's b200 arm may import / link oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "dvt_circuits_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "oracle/" not in text.replace("never the oracle/", "") and "liboracle" not in text and "oracle_lib" not in text, f


def test_rust_sys_crate_matches_the_headers():
    """rust/dkg-cuda-sys/src/lib.rs is generated from include/*.h (tools/gen_rust_sys.py): committed file == generator output, and
    every declared C function has its `pub fn`; the wrapper crate only calls functions the sys crate declares"""
    import subprocess
    import sys
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py")], capture_output=True, text=True, check=True).stdout
    committed = open(os.path.join(ROOT, "rust", "dkg-cuda-sys", "src", "lib.rs")).read()
    assert gen == committed, "run: python tools/gen_rust_sys.py > rust/dkg-cuda-sys/src/lib.rs"
    for name in declared_functions():
        assert f"pub fn {name}(" in committed, name
    wrapper = open(os.path.join(ROOT, "rust", "dkg-gpu", "src", "lib.rs")).read()
    for name in set(re.findall(r"sys::(dkg[vh]_\w+)\(", wrapper)):
        assert f"pub fn {name}(" in committed, name
    # the wrapper keeps the names of crates/dkg/src/lib.rs:6-12
    for fn in ("verify_seed_exchange_commitment", "verify_generations", "prove_wrong_final_key_generation",
               "compute_initial_commitment_hash", "verify_initial_commitment_hash"):
        assert re.search(rf"pub fn {fn}<Setup>\(", wrapper), fn
