"""CPU-only: the pairing VM (dvt_circuits_b200/csrc/pairing_vm.cuh + the generated pairing_prog.inc).
  * tools/gen_pairing_vm.py simulates the generated per-role instruction streams on Python integers against the Python
    restatement of the reference (oracle/pyref): a valid signature gives 1, an invalid one the exact Fp12 value
    e(pk, H)^3 e(-G, sig)^3; no slot is touched by two roles within a level;
  * the committed program is what the generator writes today;
  * the C++ interpreter itself (compiled for the host by tests/hostemu) runs the committed program for all roles of one check and
    must agree with the C++ oracle on the reference's KAT (crates/dkg/src/dkg_math.rs:259-278), its negatives, identity
    arguments and random keys - verdict AND pairing value."""
import ctypes
import os
import subprocess
import sys

import pytest

import oracle_lib as O
from test_oracle import KAT_BAD_SIG, KAT_MSG, KAT_PK, KAT_SIG, KAT_WRONG_PK

sys.path.insert(0, os.path.join(O.ROOT, "tools"))


@pytest.fixture(scope="module")
def L():
    path = os.path.join(O.ROOT, "tests", "hostemu", "libhostemu.so")
    src = os.path.join(O.ROOT, "tests", "hostemu", "hostemu.cpp")
    deps = [src] + [os.path.join(O.ROOT, "dvt_circuits_b200", "csrc", f) for f in ("pairing_vm.cuh", "pairing_prog.inc", "tower.cuh")]
    if not os.path.exists(path) or any(os.path.getmtime(path) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", path, src])
    return ctypes.CDLL(path)


def test_generated_program_simulates_against_pyref():
    import gen_pairing_vm as G
    segs, streams, n_slots = G.self_check(6, verbose=False)
    assert n_slots <= 37  # two blocks of 32 checks per SM: 2 x (slots x 3 KB + 1 KB) <= 227 KB
    # every role of a segment passes the same number of barriers
    for name, st in streams.items():
        assert len({sum(1 for op, _ in s if op == "BAR") for s in st}) == 1, name


def test_committed_program_is_current(tmp_path):
    import gen_pairing_vm as G
    out = tmp_path / "prog.inc"
    G.write_inc(str(out), 6)
    assert out.read_text() == open(os.path.join(O.ROOT, "dvt_circuits_b200", "csrc", "pairing_prog.inc")).read(), \
        "run: python tools/gen_pairing_vm.py --write"


def _pairing_product(pk, sig, hm):
    """e(pk, hm)^3 * e(-G, sig)^3 from the C++ oracle's pairing values, as the VM lays RA out (c0.c0 .. c1.c2, c0 then c1 of each)"""
    from oracle.pyref import bls12_381 as B
    neg_g = B.g1_compress((B.G1[0], (-B.G1[1]) % B.P))

    def f12(raw):
        c = [int.from_bytes(raw[48 * i:48 * i + 48], "big") for i in range(12)]
        return (((c[0], c[1]), (c[2], c[3]), (c[4], c[5])), ((c[6], c[7]), (c[8], c[9]), (c[10], c[11])))
    rc1, a = O.pairing_bytes(pk, hm)
    rc2, b = O.pairing_bytes(neg_g, sig)
    assert rc1 == 0 and rc2 == 0
    prod = B.f12_mul(f12(a), f12(b))
    return b"".join(x.to_bytes(48, "big") for half in prod for c in half for x in c)


def test_interpreter_on_host_against_oracle(L):
    hm = O.hash_to_g2(KAT_MSG)
    out = ctypes.create_string_buffer(576)
    assert L.he_pairing_vm(KAT_PK, KAT_SIG, hm, out) == 0          # dkg_math.rs:259-278
    assert out.raw == (1).to_bytes(48, "big") + bytes(576 - 48)
    for pk, sig, h, want in ((KAT_PK, KAT_BAD_SIG, hm, 7), (KAT_WRONG_PK, KAT_SIG, hm, 7), (KAT_PK, KAT_SIG, O.hash_to_g2(b"\x00"), 7)):
        assert L.he_pairing_vm(pk, sig, h, out) == want
        assert out.raw == _pairing_product(pk, sig, h)              # the Fp12 value itself, not only the verdict
    inf1, inf2 = bytes([0xC0]) + bytes(47), bytes([0xC0]) + bytes(95)
    assert L.he_pairing_vm(inf1, inf2, hm, None) == 0 and L.he_pairing_vm(inf1, KAT_SIG, hm, None) == 7
    assert L.he_pairing_vm(KAT_PK, inf2, hm, None) == 7
    assert L.he_pairing_vm(bytes(48), KAT_SIG, hm, None) == 48 and L.he_pairing_vm(KAT_PK, bytes(96), hm, None) == 49
    assert L.he_pairing_vm(bytes(48), bytes(96), hm, None) == 49    # signature first (verification.rs:238-241)


def test_interpreter_on_host_random_keys(L):
    from oracle.pyref import bls12_381 as B
    import random
    rnd = random.Random(5)
    msg = b"Sign with new partial key"
    hmp = B.hash_to_g2(msg)
    hm = B.g2_compress(hmp)
    assert hm == O.hash_to_g2(msg)
    for it in range(3):
        sk = rnd.randrange(1, B.R)
        pk, sig = B.g1_compress(B.g1_mul(B.G1, sk)), B.g2_compress(B.g2_mul(hmp, sk))
        assert L.he_pairing_vm(pk, sig, hm, None) == 0 == (1 - O.bls_verify_hm(pk, sig, hm))
        bad = B.g2_compress(B.g2_mul(hmp, sk ^ 1))
        assert L.he_pairing_vm(pk, bad, hm, None) == 7 and O.bls_verify_hm(pk, bad, hm) == 0
