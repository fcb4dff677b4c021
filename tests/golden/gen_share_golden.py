"""Generate tests/golden/share_small.json with the pure-Python restatement (oracle/pyref).
    python tests/golden/gen_share_golden.py
Sessions are tiny so the file stays small; every expected value is computed by pyref only."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.pyref import bls12_381 as B  # noqa: E402
from oracle.pyref import dkg as D  # noqa: E402


def prng(label, i):
    return int.from_bytes(hashlib.sha256(f"dkgv-golden|{label}|{i}".encode()).digest(), "big")


def non_subgroup_point():
    x = 1
    while True:
        y = B.fp_sqrt((x ** 3 + 4) % B.P)
        if y is not None and not B.g1_in_subgroup((x, y)):
            return (x, y)
        x += 1


def session(name, n_d, t, ids, mutate):
    coefs = [[prng(f"{name}-coef", d * 1000 + k) % B.R for k in range(t)] for d in range(n_d)]
    pts = [[B.g1_mul(B.G1, c) for c in row] for row in coefs]
    vv = [[B.g1_compress(p) for p in row] for row in pts]
    shares = [[sum(c * pow(i, k, B.R) for k, c in enumerate(row)) % B.R for i in ids] for row in coefs]
    shares_b = [[s.to_bytes(32, "big") for s in row] for row in shares]
    bad_dealers = set()
    mutate(vv, pts, shares_b, bad_dealers)
    status, evals, pks = [], [], []
    for d in range(n_d):
        srow, erow, prow = [], [], []
        for j, i in enumerate(ids):
            sk = B.fr_from_be(shares_b[d][j])
            if sk is None:
                srow.append(D.SLASHABLE_SECRET_RANGE)
                erow.append("")
                prow.append("")
                continue
            if d in bad_dealers:
                srow.append(D.PANIC_BAD_G1)
                erow.append("")
                prow.append("")
                continue
            ev = D.evaluate_polynomial([B.g1_decompress(c) for c in vv[d]], i)
            pk = B.g1_mul(B.G1, sk)
            erow.append(B.g1_compress(ev).hex())
            prow.append(B.g1_compress(pk).hex())
            srow.append(D.OK if B.g1_compress(ev) == B.g1_compress(pk) else D.SLASHABLE_SHARE_MISMATCH)
        status.append(srow)
        evals.append(erow)
        pks.append(prow)
    return {"name": name, "n_d": n_d, "t": t, "ids": ids, "vv": [[c.hex() for c in row] for row in vv],
            "shares": [[s.hex() for s in row] for row in shares_b], "status": status, "eval": evals, "pk": pks}


def mut_main(vv, pts, shares, bad):
    # flip one bit of a share (-bad-secret-key), swap in another dealer's share (-wrong-), share >= r,
    # identity coefficient, non-subgroup coefficient (-> the reference panics on decode)
    s = bytearray(shares[0][1]); s[31] ^= 1; shares[0][1] = bytes(s)
    shares[1][2] = shares[2][2]
    shares[1][3] = B.R.to_bytes(32, "big")
    shares[2][0] = bytes([0xFF]) * 32
    shares[3][4] = bytes(32)  # zero secret: pk = identity, mismatch unless the evaluation is the identity
    vv[3][1] = B.g1_compress(None)
    vv[4][2] = B.g1_compress(non_subgroup_point())
    bad.add(4)
    shares[4][1] = bytes([0xFF]) * 32  # range check precedes the decode panic


def mut_none(vv, pts, shares, bad):
    pass


def mut_identity_eval(vv, pts, shares, bad):
    # all-identity verification vector: evaluation is the identity and the zero secret verifies
    for k in range(len(vv[0])):
        vv[0][k] = B.g1_compress(None)
    for j in range(len(shares[0])):
        shares[0][j] = bytes(32)


def mut_ranks_split(vv, pts, shares, bad):
    s = bytearray(shares[0][5]); s[0] ^= 0x40; shares[0][5] = bytes(s)   # may leave the range or just mismatch
    shares[1][0] = shares[2][0]
    shares[2][11] = bytes(32)
    vv[1][4] = B.g1_compress(None)       # identity coefficient in the middle of a part
    vv[2][8] = vv[2][7]                  # repeated coefficient


out = {"sessions": [
    session("main", 5, 3, [1, 2, 3, 4, 5, 37], mut_main),
    session("t1", 2, 1, [1, 2, 9], mut_none),
    session("t0", 2, 0, [1, 2], mut_none),
    session("ragged-n", 33, 2, [1, 2, 1024, 0], mut_none),
    session("identity", 2, 3, [1, 6], mut_identity_eval),
    # recipient ids = the ranks 1..n in arbitrary column order (what a ceremony always has): the shape that
    # takes the finite-difference route of the share matrix (csrc/fdiff.cuh)
    session("ranks", 5, 3, [4, 1, 6, 2, 5, 3], mut_main),
    session("ranks-split", 3, 9, [7, 12, 1, 3, 10, 2, 9, 4, 11, 6, 8, 5], mut_ranks_split),
]}
# fixed-base G*s known answers incl. edge scalars
ks = [0, 1, 2, 255, 256, B.R - 1, prng("fb", 0) % B.R, prng("fb", 1) % B.R, 1 << 248, (1 << 255) % B.R]
out["fixed_base"] = [{"s": k.to_bytes(32, "big").hex(), "pk": B.g1_compress(B.g1_mul(B.G1, k)).hex()} for k in ks]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "share_small.json"), "w"), indent=0)
print("sessions", [(s["name"], s["status"]) for s in out["sessions"][:1]])
