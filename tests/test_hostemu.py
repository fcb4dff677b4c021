"""CPU-only: the per-thread logic of the CUDA kernels (dvt_circuits_b200/csrc/*.cuh compiled for the
host by tests/hostemu) against the oracles - field arithmetic, G1 formulas and codecs, the operand-file
(vm.cuh) formulation of the share check, the tower,
pairing and hash-to-G2.  The PTX carry-chain product itself is pinned by the -m gpu tests."""
import ctypes
import hashlib
import os
import random
import subprocess

import pytest

import oracle_lib as O
from oracle.pyref import bls12_381 as B
from test_oracle import EVAL_PKS, EVAL_TARGET, KAT_BAD_SIG, KAT_MSG, KAT_PK, KAT_SIG

H = bytes.fromhex


@pytest.fixture(scope="module")
def L():
    path = os.path.join(O.ROOT, "tests", "hostemu", "libhostemu.so")
    src = os.path.join(O.ROOT, "tests", "hostemu", "hostemu.cpp")
    csrc = os.path.join(O.ROOT, "dvt_circuits_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp", ".inc"))]
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", path, src])
    return ctypes.CDLL(path)


def buf(n):
    return ctypes.create_string_buffer(n)


def test_field_ops(L):
    rnd = random.Random(7)
    for it in range(100):
        a, b = rnd.randrange(B.P), rnd.randrange(B.P)
        if it == 0:
            a = 0
        if it == 1:
            a = b = B.P - 1
        m, ad, sb, iv = buf(48), buf(48), buf(48), buf(48)
        L.he_fp_ops(a.to_bytes(48, "big"), b.to_bytes(48, "big"), m, ad, sb, iv)
        assert int.from_bytes(m.raw, "big") == a * b % B.P
        assert int.from_bytes(ad.raw, "big") == (a + b) % B.P
        assert int.from_bytes(sb.raw, "big") == (a - b) % B.P
        assert int.from_bytes(iv.raw, "big") == pow(a, B.P - 2, B.P)
        x, y, r = rnd.randrange(B.R), rnd.randrange(B.R), buf(32)
        L.he_fr_mul(x.to_bytes(32, "big"), y.to_bytes(32, "big"), r)
        assert int.from_bytes(r.raw, "big") == x * y % B.R


def test_inversion_by_binary_gcd(L):
    """fp_inv_bgcd (binary extended Euclid with count-trailing-zeros strips) == a^(p-2) on random and extreme operands"""
    rnd = random.Random(21)
    cases = [0, 1, 2, 3, B.P - 1, B.P - 2, (B.P - 1) // 2, (B.P + 1) // 2, 1 << 32, 1 << 64, (1 << 380) + 1, 1 << 380, 3 << 200]
    cases += [pow(2, k, B.P) for k in (31, 32, 33, 95, 96, 383, 384)] + [rnd.randrange(B.P) for _ in range(400)]
    for a in cases:
        o = buf(48)
        L.he_fp_inv_bgcd(a.to_bytes(48, "big"), o)
        assert int.from_bytes(o.raw, "big") == pow(a, B.P - 2, B.P), hex(a)


def test_g1_formulas_and_codec(L):
    rnd = random.Random(3)
    for it in range(6):
        p = B.g1_mul(B.G1, rnd.randrange(B.R))
        q = B.g1_mul(B.G1, rnd.randrange(B.R))
        for x, y in [(p, q), (p, p), (p, B.E1.neg(p)), (None, q), (p, None), (None, None)]:
            k = rnd.choice([0, 1, 2, 3, 5, 1023, 1024, 0xFFFFFFFF, rnd.randrange(2 ** 32)])
            o = [buf(48) for _ in range(4)]
            assert L.he_g1_ops(B.g1_compress(x), B.g1_compress(y), k, *o) == 0
            assert o[0].raw == B.g1_compress(B.g1_add(x, y)) == o[1].raw
            assert o[2].raw == B.g1_compress(B.g1_add(x, x))
            assert o[3].raw == B.g1_compress(B.g1_mul(x, k))
    o = buf(48)
    for enc in (KAT_PK, bytes([0xC0]) + bytes(47), bytes(48), bytes([0xE0]) + bytes(47)):
        assert L.he_g1_decompress(enc, o) == O.g1_decompress(enc)[0]
    pb = bytearray(B.P.to_bytes(48, "big"))
    pb[0] |= 0x80
    assert L.he_g1_decompress(bytes(pb), o) == 2
    x = 1
    while True:  # on the curve, outside the subgroup (cf. report-1-bad-aggregate-pubkey)
        yy = B.fp_sqrt((x ** 3 + 4) % B.P)
        if yy is not None and not B.g1_in_subgroup((x, yy)):
            break
        x += 1
    assert L.he_g1_decompress(B.g1_compress((x, yy)), o) == 4


def test_share_check_all_formulations(L):
    """he_share_check runs the inlined and operand-file variants and returns >= 0x100 on any
    disagreement between them"""
    ev, pk = buf(48), buf(48)
    vv = b"".join(H(h) for h in EVAL_PKS)
    assert L.he_share_check(vv, 3, 1, (5).to_bytes(32, "big"), ev, pk) == 4
    assert ev.raw.hex() == EVAL_TARGET  # dkg_math.rs:281-298
    assert pk.raw == B.g1_compress(B.g1_mul(B.G1, 5))
    rnd = random.Random(11)
    for t in (0, 1, 2, 6):
        coef = [rnd.randrange(B.R) for _ in range(t)]
        vv = b"".join(B.g1_compress(B.g1_mul(B.G1, c)) for c in coef)
        for i in (0, 1, 2, 7, 1023, 1024, 0xFFFFFFFF):
            s = sum(c * pow(i, k, B.R) for k, c in enumerate(coef)) % B.R
            assert L.he_share_check(vv, t, i, s.to_bytes(32, "big"), ev, pk) == 0, (t, i)
            assert L.he_share_check(vv, t, i, ((s + 1) % B.R).to_bytes(32, "big"), ev, pk) == 4
        assert L.he_share_check(vv, t, 3, B.R.to_bytes(32, "big"), ev, pk) == 1
    # the signed-odd fixed-base table (feldman.cuh gtab_*): G * s on the digit extremes - even / odd scalars, all-zero and all-one
    # windows, carries into the next window, s = 0 (the sum meets P + (-P))
    edge = [0, 1, 2, 3, B.R - 1, B.R - 2, 0xFFFF, 0x10000, 0x10001, 0xFFFF0000FFFF, (1 << 240) + 1, (1 << 240), (1 << 254) - 1,
            0x8000800080008000800080008000800080008000800080008000800080008000 % B.R, int("55" * 32, 16) % B.R, int("aa" * 32, 16) % B.R]
    for s in edge + [rnd.randrange(B.R) for _ in range(24)]:
        L.he_share_check(b"", 0, 1, s.to_bytes(32, "big"), ev, pk)
        assert pk.raw == B.g1_compress(B.g1_mul(B.G1, s)), hex(s)
    # every window width a ctx accepts (dkgv_ctx_create_ex): the digit extraction crosses limb boundaries differently for each
    for bits in range(8, 27):
        for s in [0, 1, 2, B.R - 1, B.R - 2, (1 << 254) - 1, int("55" * 32, 16) % B.R, rnd.randrange(B.R), rnd.randrange(B.R)]:
            L.he_fixed_base_bits(bits, s.to_bytes(32, "big"), pk)
            assert pk.raw == B.g1_compress(B.g1_mul(B.G1, s)), (bits, hex(s))
    # the table builder (one small multiplication, a walk of mixed additions, one inversion per run) against the definition
    for bits, w, m in [(8, 0, 0), (8, 31, 127), (13, 5, 4000), (16, 15, 32767), (22, 11, (1 << 21) - 1), (22, 0, 0), (26, 9, (1 << 25) - 5)]:
        assert L.he_gtab_builder(bits, w, m) == 0, (bits, w, m)


def test_tower_pairing_h2c(L):
    for m in (b"", b"abc", bytes(range(200))):
        o = buf(32)
        L.he_sha256(m, ctypes.c_size_t(len(m)), o)
        assert o.raw == hashlib.sha256(m).digest()
    for m in (b"Sign with new partial key", KAT_MSG, b"", b"x" * 300):
        o = buf(96)
        L.he_hash_to_g2(m, ctypes.c_size_t(len(m)), o)
        assert o.raw == O.hash_to_g2(m)
    o = buf(96)
    assert L.he_g2_decompress(KAT_SIG, o) == 0 and o.raw == KAT_SIG
    assert L.he_g2_decompress(bytes(96), o) == 1
    hm = O.hash_to_g2(KAT_MSG)
    assert L.he_bls_verify_hm(KAT_PK, KAT_SIG, hm) == 1  # dkg_math.rs:259-278
    assert L.he_bls_verify_hm(KAT_PK, KAT_BAD_SIG, hm) == 0
    assert L.he_bls_verify_hm(KAT_PK, KAT_SIG, O.hash_to_g2(b"\x00")) == 0
    inf1, inf2 = bytes([0xC0]) + bytes(47), bytes([0xC0]) + bytes(95)
    assert L.he_bls_verify_hm(inf1, inf2, hm) == 1 and L.he_bls_verify_hm(inf1, KAT_SIG, hm) == 0
    a = buf(576)
    assert L.he_pairing_bytes(KAT_PK, KAT_SIG, a) == 0
    assert O.pairing_bytes(KAT_PK, KAT_SIG) == (0, a.raw)


def test_finite_difference_row(L):
    """fdiff.cuh: polynomial split into m parts, seeds by Horner on a window around 0, backward
    differences, wavefront extension, joint double-and-add recombination - every f(1..n) must equal the
    oracle's evaluate_polynomial (dkg_math.rs:160-174)"""
    rnd = random.Random(5)
    NONE = 0x7FFFFFFF
    cases = [(2, 5, 1, None), (3, 9, 1, None), (3, 9, 1, 1), (3, 9, 1, -1), (5, 12, 1, None), (5, 12, 1, -3), (5, 6, 1, 1), (8, 20, 0, None),
             (8, 20, 2, None), (9, 20, 2, 0), (7, 7, 2, None), (12, 9, 3, None), (12, 30, 4, None), (13, 16, 5, 1)]
    for t, n_r, m, lo in cases:
        coef = [rnd.randrange(B.R) for _ in range(t)]
        if t == 5:
            coef[2] = 0  # an identity coefficient
        vv = b"".join(B.g1_compress(B.g1_mul(B.G1, c)) for c in coef)
        plan = (ctypes.c_int32 * 6)()
        out = buf(48 * n_r)
        rc = L.he_fd_row(vv, t, n_r, m, ctypes.c_int32(NONE if lo is None else lo), plan, out)
        assert rc == 0, (t, n_r, m, lo, rc)
        _, pm, ph, plo, phi, steps = list(plan)
        assert (m == 0 or pm == m) and ph == -(-t // pm) and phi - plo + 1 == ph and plo <= 1 <= phi and steps == n_r - phi
        for j in range(n_r):
            want = B.g1_compress(B.g1_mul(B.G1, sum(c * pow(j + 1, k, B.R) for k, c in enumerate(coef)) % B.R))
            assert out.raw[48 * j:48 * j + 48] == want, (t, n_r, m, lo, j)
    plan = (ctypes.c_int32 * 6)()
    assert L.he_fd_row(bytes(48), 1, 5, 0, ctypes.c_int32(NONE), plan, buf(48 * 5)) == -1 and plan[0] == 0
    L.he_fd_row(bytes(48 * 683), 683, 1024, 0, ctypes.c_int32(NONE), plan, buf(48))  # plan only (vv undecodable -> -3)
    assert plan[0] == 1 and plan[1] > 1 and plan[3] < 0 < plan[4]


def test_recombination_digits(L):
    """fd_comb_digits: k = x^(h i) mod r splits as k1 + k2 z^2 (GLV), each half in width-4 signed digits"""
    Z2 = 0xD201000000010000 ** 2
    out = (ctypes.c_int8 * (132 * 2 * 5))()
    for x, h, m in [(1, 7, 2), (2, 114, 6), (1024, 114, 6), (777, 1, 3), (0xFFFFFFFF, 65535, 4), (3, (B.R - 1) // 2, 2)]:
        top = L.he_fd_digits(x, h & 0xFFFFFFFF, m, out)
        tops = []
        for i in range(1, m):
            ks = []
            for half in (0, 1):
                dg = [out[((i - 1) * 2 + half) * 132 + b] for b in range(132)]
                assert all(dv == 0 or (dv % 2 and abs(dv) <= 7) for dv in dg)
                nz = [b for b in range(132) if dg[b]]
                assert all(b2 - b1 >= 4 for b1, b2 in zip(nz, nz[1:]))
                tops.append(max(nz, default=-1))
                ks.append(sum(dv << b for b, dv in enumerate(dg)))
            assert 0 <= ks[0] < Z2 and 0 <= ks[1] < 2 ** 128
            assert ks[0] + ks[1] * Z2 == pow(x, (h & 0xFFFFFFFF) * i, B.R)
        assert top == max(tops)
    # z^2 acts as -phi on G1: [z^2](x, y) = (beta x, -y)
    p = B.g1_mul(B.G1, 12345)
    q = B.g1_mul(p, Z2 % B.R)
    assert q[1] == (-p[1]) % B.P and pow(q[0] * pow(p[0], -1, B.P) % B.P, 3, B.P) == 1


def test_difference_table_shortcut(L):
    """k_fd_difftab (share_fd.cu), the fused conditions-(2)-and-interpolation kernel of the consistency shortcut, run thread by
    thread on the host (tests/hostemu he_difftab: the kernel's own per-thread routines, predicates and double buffer):
    fr_submul_small against Python integers on random and extreme operands; the table returns the dealer's monomial
    coefficients for consistent shares and flags every kind of inconsistency the t-th differences catch."""
    from math import factorial
    R = B.R
    rnd = random.Random(23)
    o = buf(32)
    edge = [0, 1, 2, R - 1, R - 2, R >> 1, (1 << 224) - 1, 1 << 224, (1 << 234) - 1, 1 << 234]
    cases = [(p_, a_, j_) for p_ in edge for a_ in edge for j_ in (0, 1, 2, 511, 682, 1022, 1023)]
    cases += [(rnd.randrange(R), rnd.randrange(R), rnd.randrange(1024)) for _ in range(3000)]
    for p_, a_, j_ in cases:
        L.he_fr_submul_small(p_.to_bytes(32, "big"), a_.to_bytes(32, "big"), j_, o)
        assert int.from_bytes(o.raw, "big") == (p_ - j_ * a_) % R, (hex(p_), hex(a_), j_)

    # lz_reduce (fdiff.cuh): the lazy 9-limb values of the table back to canonical residues, on the extremes of both ranges
    import struct

    def reduce(v, signed):
        w = v % (1 << 288)
        L.he_lz_reduce(struct.pack("<9I", *[(w >> (32 * i)) & 0xffffffff for i in range(9)]), int(signed), o)
        return int.from_bytes(o.raw, "big")

    sig_edge = [0, 1, -1, R, -R, R - 1, 1 - R, (1 << 285), -(1 << 285), (1 << 285) - 1, 1 - (1 << 285), (1 << 256), -(1 << 256),
                3 * R, -3 * R, (R << 30), -(R << 30), (R << 30) - 1, (R << 30) + 1, ((1 << 285) // R) * R, -(((1 << 285) // R) * R)]
    for v in sig_edge + [rnd.randrange(-(1 << 285), (1 << 285) + 1) for _ in range(3000)] + [rnd.randrange(-R, R) for _ in range(500)]:
        assert reduce(v, True) == v % R, hex(v)
    uns_edge = [0, 1, R - 1, R, R + 1, 2 * R, 3 * R - 1, 1 << 256, (1 << 286), (1 << 286) - 1, ((1 << 286) // R) * R, ((1 << 286) // R) * R - 1,
                (1 << 224) - 1, 1 << 224, (1 << 255) - 1]
    for v in uns_edge + [rnd.randrange((1 << 286) + 1) for _ in range(3000)] + [rnd.randrange(1 << rnd.randrange(1, 287)) for _ in range(1000)]:
        assert reduce(v, False) == v % R, hex(v)

    def run(s, t):
        n = len(s)
        ifact = b"".join(pow(factorial(k), -1, R).to_bytes(32, "big") for k in range(t))
        coef = buf(32 * t)
        rc = L.he_difftab(b"".join(x.to_bytes(32, "big") for x in s), n, t, ifact, coef)
        return rc, [int.from_bytes(coef.raw[32 * k:32 * k + 32], "big") for k in range(t)]

    def horner(a, x):
        v = 0
        for c in reversed(a):
            v = (v * x + c) % R
        return v

    for t, n in [(1, 2), (1, 5), (2, 3), (2, 5), (3, 9), (7, 12), (8, 9), (20, 33), (33, 64), (43, 64), (64, 65), (97, 200), (130, 131),
                 (683, 1024), (1024, 1025), (1023, 2048), (1500, 2048), (2047, 2048)]:
        a = [rnd.randrange(R) for _ in range(t)]
        if t > 2:
            a[t - 1] = 0 if n % 2 else a[t - 1]  # a polynomial of lower degree is fine too
        s = [horner(a, x) for x in range(1, n + 1)]
        rc, c = run(s, t)
        assert rc == 0 and c == a, (t, n)
        # one flipped bit anywhere (inside or beyond the first t shares)
        for pos in {0, t - 1, t, n - 1, rnd.randrange(n)}:
            bad = list(s)
            bad[pos] ^= 1 << rnd.randrange(250)
            bad[pos] %= R
            assert run(bad, t)[0] == 1, (t, n, pos)
        # a deviation c * prod_{i<=t}(x - i): s(1..t) intact, caught by the differences beyond t
        dev = list(s)
        for x in range(1, n + 1):
            d = 5
            for i in range(1, t + 1):
                d = d * (x - i) % R
            dev[x - 1] = (dev[x - 1] + d) % R
        assert dev[:t] == s[:t] and run(dev, t)[0] == 1
        # consistent shares of a polynomial of degree exactly t (one too many)
        a2 = a + [rnd.randrange(1, R)]
        s2 = [horner(a2, x) for x in range(1, n + 1)]
        assert run(s2, t)[0] == 1


def test_compressed_commitment_check(L):
    """fd_coef_point + fd_coef_signs (k_fd_coefpoint / k_fd_coefsign): compress(G * p_k) == C_k decided without decompressing C_k
    (x_C * Z == X, batched 1/Z for the sign).  Against the Python oracle: the true encoding agrees; the other sign, another
    point, x + p (non-canonical), missing compression flag, infinity flag on a point, the identity encoding against a non-zero
    scalar and a non-identity encoding against the zero scalar do not; scalar 0 with the identity encoding agrees."""
    rnd = random.Random(31)
    G = B.G1
    scal, encs, want = [], [], []

    def add(k, e, w):
        scal.append(k.to_bytes(32, "big"))
        encs.append(bytes(e))
        want.append(w)

    ks = [1, 2, B.R - 1, rnd.randrange(B.R), rnd.randrange(B.R), rnd.randrange(1 << 64), rnd.randrange(B.R)]
    for k in ks:
        good = bytearray(B.g1_compress(B.g1_mul(G, k)))
        add(k, good, 1)
        e = bytearray(good)
        e[0] ^= 0x20  # the other square root
        add(k, e, 0)
        add(k, B.g1_compress(B.g1_mul(G, (k + 1) % B.R or 2)), 0)
        e = bytearray(good)
        e[0] &= 0x7F  # compression flag missing
        add(k, e, 0)
        e = bytearray(good)
        e[0] |= 0x40  # infinity flag on a point
        add(k, e, 0)
        x = int.from_bytes(bytes([good[0] & 0x1F]) + bytes(good[1:]), "big")
        if x + B.P < 1 << 381:  # same residue, non-canonical encoding
            e = bytearray((x + B.P).to_bytes(48, "big"))
            e[0] |= good[0] & 0xE0
            add(k, e, 0)
        add(k, bytes([0xC0]) + bytes(47), 0)  # identity encoding, non-zero scalar
    add(0, bytes([0xC0]) + bytes(47), 1)
    add(0, bytes([0xE0]) + bytes(47), 0)  # identity with the sign flag
    add(0, B.g1_compress(G), 0)
    add(0, bytes([0xC0]) + bytes(46) + b"\x01", 0)
    n = len(want)
    out = buf(n)
    L.he_coef_bytes_check(b"".join(scal), b"".join(encs), n, out)
    assert list(out.raw) == want, [i for i in range(n) if out.raw[i] != want[i]]
