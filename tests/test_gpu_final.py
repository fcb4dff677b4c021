"""-m gpu parity tests of the finalization path through the C ABI: key aggregation, Lagrange at 0,
hash-to-G2, G2 decoding and the batched BLS pairing checks - against the reference's own KATs
(crates/dkg/src/dkg_math.rs:259-375) and the CPU oracle on seeded random inputs."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_oracle import (EVAL_PKS, EVAL_TARGET, KAT_BAD_SIG, KAT_MSG, KAT_PK, KAT_SIG, KAT_WRONG_PK, LAG_PKS, LAG_TARGET, pts)

pytestmark = pytest.mark.gpu
H = bytes.fromhex
INF1, INF2 = bytes([0xC0]) + bytes(47), bytes([0xC0]) + bytes(95)


def test_hash_to_g2(verifier):
    msgs = [b"Sign with new partial key", KAT_MSG, b"", b"x" * 300, b"hello"]
    out = verifier.hash_to_g2(msgs)
    for m, o in zip(msgs, out):
        assert bytes(o) == O.hash_to_g2(m)


def test_g2_decompress(verifier):
    cases = [KAT_SIG, KAT_BAD_SIG, INF2, bytes(96), bytes([0xE0]) + bytes(95), O.hash_to_g2(b"abc")]
    st = verifier.g2_decompress_check(np.array([list(c) for c in cases], dtype=np.uint8))
    assert st.tolist() == [O.g2_decompress(c)[0] for c in cases]
    assert st.tolist()[:3] == [0, 0, 0]


def test_kat_bls_verify(verifier):
    # dkg_math.rs:259-278 incl. the three negatives, + identity semantics
    hm = [O.hash_to_g2(KAT_MSG), O.hash_to_g2(b"\x00")]
    pk = [KAT_PK, KAT_PK, KAT_WRONG_PK, KAT_PK, INF1, INF1, KAT_PK, bytes(48), KAT_PK]
    sig = [KAT_SIG, KAT_SIG, KAT_SIG, KAT_BAD_SIG, INF2, KAT_SIG, INF2, KAT_SIG, bytes(96)]
    idx = [0, 1, 0, 0, 0, 0, 0, 0, 0]
    st = verifier.bls_verify_batch(np.array([list(x) for x in pk], dtype=np.uint8), np.array([list(x) for x in sig], dtype=np.uint8),
                                   np.array([list(x) for x in hm], dtype=np.uint8), idx)
    assert st.tolist() == [0, 7, 7, 7, 0, 7, 7, 48, 49]


def test_bls_batch_random(verifier):
    rng = np.random.default_rng(5)
    m = 40
    msg = b"Sign with new partial key"
    hm = O.hash_to_g2(msg)
    sks = [int.from_bytes(rng.bytes(31), "big") + 1 for _ in range(m)]
    pks = [O.g1_fixed_base(s.to_bytes(32, "big"))[1] for s in sks]
    # sig_i = sk_i * H(m): computed with the pure-Python restatement (G2 scalar multiplication)
    from oracle.pyref import bls12_381 as B
    hpt = B.g2_decompress(hm)
    sigs = [B.g2_compress(B.g2_mul(hpt, s)) for s in sks]
    for i in range(0, m, 5):  # corrupt every 5th: wrong key for that signature
        pks[i] = pks[(i + 1) % m]
    st = verifier.bls_verify_batch(np.array([list(x) for x in pks], dtype=np.uint8), np.array([list(x) for x in sigs], dtype=np.uint8),
                                   np.array([list(hm)], dtype=np.uint8))
    exp = O.bls_verify_batch(np.array([list(x) for x in pks], dtype=np.uint8), np.array([list(x) for x in sigs], dtype=np.uint8), hm, threads=4)
    assert st.tolist() == [0 if e == 1 else 7 for e in exp.tolist()]
    assert st.tolist().count(7) == m // 5


def test_kat_lagrange(verifier):
    assert verifier.lagrange_at_zero(pts(LAG_PKS), [1, 2, 3, 4, 5]) == (0, H(LAG_TARGET))
    assert verifier.lagrange_at_zero(pts(LAG_PKS[4:] + LAG_PKS[:4]), [5, 1, 2, 3, 4]) == (0, H(LAG_TARGET))
    st, out = verifier.lagrange_at_zero(pts([LAG_PKS[1], LAG_PKS[0]] + LAG_PKS[2:]), [1, 2, 3, 4, 5])
    assert st == 0 and out != H(LAG_TARGET)
    assert verifier.lagrange_at_zero(pts(LAG_PKS), [1, 2, 0, 4, 5])[0] == 36
    assert verifier.lagrange_at_zero(pts(LAG_PKS), [1, 2, 2, 4, 5])[0] == 37
    assert verifier.lagrange_at_zero(pts(LAG_PKS[:1]), [9]) == (0, H(LAG_PKS[0]))
    assert verifier.lagrange_at_zero(np.zeros((0, 48), dtype=np.uint8), [])[0] == 32
    bad = pts(LAG_PKS).copy()
    bad[2] = 0
    assert verifier.lagrange_at_zero(bad, [1, 2, 3, 4, 5])[0] == 48


def test_kat_eval_points(verifier):
    st, out = verifier.eval_points(pts(EVAL_PKS), [1, 0, 7])
    assert st == 0 and bytes(out[0]).hex() == EVAL_TARGET
    assert bytes(out[1]) == H(EVAL_PKS[0])  # id 0 -> C_0
    assert bytes(out[2]) == O.evaluate_polynomial(b"".join(H(h) for h in EVAL_PKS), 3, 7)[1]


@pytest.mark.parametrize("n,t", [(3, 2), (7, 5), (40, 9), (33, 1)])
def test_agg_final_keys_vs_oracle(verifier, n, t):
    from dvt_circuits_b200 import synthetic
    s = synthetic.make_session(verifier, n, n, t, seed=99 + n)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    st, co, keys = verifier.agg_final_keys(s["vv"], ids)
    ost, oco, okeys = O.agg_coefficients(s["vv"], ids, O.FAST)
    assert st == ost == 0
    assert (co == oco).all() and (keys == okeys).all()
    # size-independent property: Lagrange at 0 over the final keys returns the aggregate C_0
    lst, agg = verifier.lagrange_at_zero(keys, ids)
    if n >= t:
        assert lst == 0 and agg == bytes(co[0])
    assert O.lagrange(keys, ids, O.FAST) == (lst, agg)


def test_agg_bad_point(verifier):
    from dvt_circuits_b200 import synthetic
    s = synthetic.make_session(verifier, 4, 4, 3)
    vv = s["vv"].copy()
    vv[2, 1] = 0
    st, _, _ = verifier.agg_final_keys(vv, [1, 2, 3, 4])
    assert st == 48


@pytest.mark.parametrize("n,t", [(8, 5), (48, 17)])
def test_synthetic_finalization_properties(verifier, n, t):
    """size-independent properties of a whole synthetic ceremony (config 4): every partial signature
    verifies, the partial public keys equal the final keys K_j, and both Lagrange interpolations give
    the aggregate key C_0 - then the same through the JSON flow (n <= 255 fits the reference format)."""
    import json
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import synthetic
    f = synthetic.make_finalization(verifier, n, t)
    st = verifier.bls_verify_batch(f["partial_pubkeys"], f["signatures"], f["hm"])
    assert not st.any()
    ast, co, keys = verifier.agg_final_keys(f["vv"], f["ids"])
    assert ast == 0 and (keys == f["partial_pubkeys"]).all()
    assert verifier.lagrange_at_zero(keys, f["ids"]) == (0, bytes(co[0]))
    # a corrupted signature set: swap two signatures -> exactly those two checks fail
    sig = f["signatures"].copy()
    sig[[0, 1]] = sig[[1, 0]]
    st = verifier.bls_verify_batch(f["partial_pubkeys"], sig, f["hm"])
    assert st.tolist() == [7, 7] + [0] * (n - 2)
    # JSON flow
    gen_id = bytes(range(16))
    gens = []
    for i in range(n):
        gens.append({"base_pubkeys": [bytes(p).hex() for p in f["vv"][i]],
                     "base_hash": dk.initial_commitment_hash(gen_id, n, t, f["vv"][i]).hex(),
                     "partial_pubkey": "", "message_cleartext": f["message"].decode(), "message_signature": ""})
    # recipient ids follow the sorted base hashes: generation at sorted position p owns id p+1
    order = sorted(range(n), key=lambda i: bytes.fromhex(gens[i]["base_hash"]))
    for pos, i in enumerate(order):
        gens[i]["partial_pubkey"] = bytes(f["partial_pubkeys"][pos]).hex()
        gens[i]["message_signature"] = bytes(f["signatures"][pos]).hex()
    # the aggregate key does not depend on which id a dealer holds
    data = {"settings": {"n": n, "k": t, "gen_id": gen_id.hex()}, "generations": gens, "aggregate_pubkey": bytes(co[0]).hex()}
    assert verifier.execute("finalization", json.dumps(data))[:2] == (0, 0)
    data["aggregate_pubkey"] = bytes(co[1]).hex()
    assert verifier.execute("finalization", json.dumps(data))[:2] == (1, 34)


def test_initial_commitment_hashes(verifier):
    import hashlib
    import dvt_circuits_b200 as dk
    from dvt_circuits_b200 import synthetic
    gen_id = bytes(range(16))
    for n, t in ((5, 3), (70, 43), (3, 300)):
        vv = synthetic.make_session(verifier, n, 1, t)["vv"]
        out = verifier.initial_commitment_hashes(vv, gen_id, n, t)
        for d in range(n):
            exp = hashlib.sha256(gen_id + bytes([n & 0xFF, t & 0xFF, t & 0xFF]) + vv[d].tobytes()).digest()
            assert bytes(out[d]) == exp
            assert dk.initial_commitment_hash(gen_id, n, t, vv[d]) == exp


def test_q1_expected_key_on_gpu(verifier):
    """quirk Q1 on the GPU, pinned on SURVEY App. C3: a valid-signature item built from the reference's finalization/report-1.json
    reaches verify_expected_key (verification.rs:399-420, 523-551) through the dkg_prover_host front end: status 8, exit 0, the
    (expected, got) keys of the reference's message and the guest's committed public values."""
    import json
    import q1_fixture as Q
    st, out = verifier.eval_points(np.array([list(bytes.fromhex(h)) for h in Q.FINAL_KEYS], dtype=np.uint8), [1, 2, 3])
    assert st == 0 and [bytes(o).hex() for o in out] == Q.Q1_EXPECTED
    for perp in range(3):
        item = Q.q1_item(perp)
        code, status, msg, public, keys = verifier.execute_report("bad-partial-key", json.dumps(item), auth=False)
        assert (code, status) == (0, 8), msg
        assert keys is not None and keys[0].hex() == Q.Q1_EXPECTED[perp] and keys[1].hex() == Q.FINAL_KEYS[perp]
        # bad_parial_key_prove/src/main.rs:31-41: every generation's base_hash in INPUT order, then the perpetrator's identity key
        assert [p.hex() for p in public] == [g["base_hash"] for g in item["generations"]] + [item["bad_partial"]["commitment"]["pubkey"]]
    # the same items through the batch entry: one call, three perpetrators, + an item accusing the expected key itself
    gens = Q.sorted_generations()
    vv = np.array([[list(bytes.fromhex(p)) for p in g["base_pubkeys"]] for g in gens], dtype=np.uint8)
    pk = np.array([list(bytes.fromhex(g["partial_pubkey"])) for g in gens], dtype=np.uint8)
    sig = np.array([list(bytes.fromhex(g["message_signature"])) for g in gens], dtype=np.uint8)
    stt, exp, sst = verifier.bad_partial_key_verify_batch(vv, [0, 1, 2], pk, sig, [gens[0]["message_cleartext"].encode()])
    assert sst == 0 and stt.tolist() == [8, 8, 8] and [bytes(e).hex() for e in exp] == Q.Q1_EXPECTED
    # accused key == expected key but the signature is for another key: the pairing check fires first (7)
    stt, _, _ = verifier.bad_partial_key_verify_batch(vv, [0], exp[:1], sig[:1], [gens[0]["message_cleartext"].encode()])
    assert stt.tolist() == [7]


def test_guest_public_values(verifier):
    """the public values each guest commits, in commit order (bad_share_exchange_prove/src/main.rs:57-71,
    finalization_prove/src/main.rs:26-32), and the (expected, got) pair of a share mismatch (verification.rs:141-145)"""
    import json
    vec = os.path.join(O.ROOT, "tests", "golden", "reference_vectors", "no_auth")
    j = json.load(open(os.path.join(vec, "share", "seeds-commitment-from-2-to-1-bad-secret-key.json")))["scenario"]
    code, status, msg, public, keys = verifier.execute_report("bad-share", json.dumps(j), auth=False)
    assert (code, status) == (0, 4), msg
    assert [p.hex() for p in public] == j["base_hashes"] + [j["seeds_exchange_commitment"]["commitment"]["pubkey"]]
    assert keys[0].hex() == "b29c8acec16b2193a36e635e65727bbbad73bbbbc933295863b92fe1dee26fe3581d0ecb22cdfc09a714ad1ffa0e8ccb"  # C1
    assert keys[1].hex() == "8d13ea70941e0ac58adeecc4bd4105d213de0113369e69c7d1b9d0222e31b163015ed08bc95f458dba3599504ce3d401"  # C2
    j = json.load(open(os.path.join(vec, "finalization", "report-1.json")))["scenario"]
    code, status, msg, public, keys = verifier.execute_report("finalization", json.dumps(j))
    assert (code, status) == (0, 0) and keys is None
    assert [p.hex() for p in public] == [g["base_hash"] for g in j["generations"]] + [j["aggregate_pubkey"]]
    # a valid share proves nothing: exit 1, nothing committed
    j = json.load(open(os.path.join(vec, "share", "seeds-commitment-from-2-to-1.json")))["scenario"]
    code, status, msg, public, keys = verifier.execute_report("bad-share", json.dumps(j), auth=False)
    assert (code, status, public, keys) == (1, 0, [], None)


@pytest.mark.parametrize("n,t,m", [(8, 5, 96), (3, 2, 40)])
def test_bad_partial_key_batch_against_oracle(verifier, n, t, m):
    """dkgv_bad_partial_key_verify_batch on a synthetic session with every kind of item (valid under quirk Q1, flipped key /
    signature bits, wrong-but-valid key / signature, the honest partial key) against the C++ oracle composed per item in the
    reference's order of checks (verification.rs:440-463)."""
    from dvt_circuits_b200 import synthetic
    fin = synthetic.make_finalization(verifier, n, t)
    it = synthetic.make_bad_partial_items(verifier, fin, m, p_bad=0.7)
    st, exp, sst = verifier.bad_partial_key_verify_batch(fin["vv"], it["perp"], it["pk"], it["sig"], [fin["message"]])
    assert sst == 0
    assert (st == it["expected"]).all(), np.argwhere(st != it["expected"])[:10]
    assert set(st.tolist()) == {0, 5, 6, 7, 8}
    ast, co, keys = O.agg_coefficients(fin["vv"], fin["ids"])
    assert ast == 0
    for p in range(n):
        est, q1 = O.evaluate_polynomial(keys.tobytes(), n, p + 1)
        assert est == 0 and q1 == bytes(exp[p]) == bytes(it["q1_keys"][p])
    for i in range(m):
        pk, sg = bytes(it["pk"][i]), bytes(it["sig"][i])
        if O.g1_decompress(pk)[0]:
            want = 5
        elif O.g2_decompress(sg)[0]:
            want = 6
        elif O.bls_verify(pk, sg, fin["message"]) != 1:
            want = 7
        else:
            want = 0 if pk == bytes(exp[it["perp"][i]]) else 8
        assert st[i] == want, (i, int(it["kind"][i]), st[i], want)
    # an undecodable commitment in the session: items that reach verify_expected_key panic (48), earlier exits are unchanged
    vv2 = fin["vv"].copy()
    vv2[1, 0, 0] &= 0x7F
    st2, _, sst2 = verifier.bad_partial_key_verify_batch(vv2, it["perp"], it["pk"], it["sig"], [fin["message"]])
    assert sst2 == 48
    assert (st2 == np.where((st == 0) | (st == 8), 48, st)).all()


@pytest.mark.parametrize("m", [1, 31, 33, 200])
def test_pairing_vm_against_thread_kernel_and_oracle(verifier, m):
    """the pairing VM (several warps per 32 checks, csrc/pairing_vm.cuh) and the one-thread-per-check kernel are two independent
    implementations of bls_verify_precomputed_hash: same statuses on valid, wrong, undecodable and identity arguments, with
    several hashed messages - and both equal the C++ oracle's."""
    from dvt_circuits_b200 import synthetic
    fin = synthetic.make_finalization(verifier, 8, 3)
    msgs = [b"Sign with new partial key", b"another message", b""]
    hms = verifier.hash_to_g2(msgs)
    sk = fin["partial_secrets"]
    sigs = [verifier.g2_mul_batch(hms[k].tobytes(), sk) for k in range(3)]
    rng = np.random.Generator(np.random.PCG64(m))
    who = rng.integers(0, 8, size=m)
    idx = rng.integers(0, 3, size=m).astype(np.uint32)
    pk = fin["partial_pubkeys"][who].copy()
    sg = np.stack([sigs[idx[i]][who[i]] for i in range(m)])
    kind = rng.integers(0, 8, size=m)
    for i in range(m):
        if kind[i] == 0:
            sg[i] = sigs[idx[i]][(who[i] + 1) % 8]      # wrong signer
        elif kind[i] == 1:
            sg[i] = sigs[(idx[i] + 1) % 3][who[i]]      # wrong message
        elif kind[i] == 2:
            pk[i, 17] ^= 4                              # undecodable key
        elif kind[i] == 3:
            sg[i, 60] ^= 1                              # undecodable signature
        elif kind[i] == 4:
            pk[i], sg[i] = np.frombuffer(INF1, dtype=np.uint8), np.frombuffer(INF2, dtype=np.uint8)  # identity = identity
        elif kind[i] == 5:
            pk[i] = np.frombuffer(INF1, dtype=np.uint8)  # e(O, H) = 1 != e(G, sig)
    try:
        verifier.set_bls_path(verifier.BLS_VM)
        st_vm = verifier.bls_verify_batch(pk, sg, hms, idx)
        assert verifier.last_bls_path == verifier.BLS_VM
        verifier.set_bls_path(verifier.BLS_THREAD)
        st_th = verifier.bls_verify_batch(pk, sg, hms, idx)
        assert verifier.last_bls_path == verifier.BLS_THREAD
    finally:
        verifier.set_bls_path(verifier.BLS_AUTO)
    assert (st_vm == st_th).all(), np.argwhere(st_vm != st_th)[:10]
    for i in range(m):
        r = O.bls_verify_hm(bytes(pk[i]), bytes(sg[i]), bytes(hms[idx[i]]))
        want = {1: 0, 0: 7, -48: 48, -49: 49}[r]
        if kind[i] == 3 and r == -48:
            want = 49
        assert st_vm[i] == want, (i, int(kind[i]), int(st_vm[i]), r)


def test_bad_partial_key_batch_at_scale(verifier):
    """BASELINE config 5, second half, at a size the oracle can follow: 8 192 items over a (64, 43) session, half of them corrupted.
    Every status equals the one the construction predicts; every item that is NOT valid and a 1/64 sample of the valid ones is
    re-decided by the C++ oracle (decoding, pairing equality on all host threads, expected key of quirk Q1)."""
    from dvt_circuits_b200 import synthetic
    fin = synthetic.make_finalization(verifier, 64, 43)
    m = 8192
    it = synthetic.make_bad_partial_items(verifier, fin, m, p_bad=0.5)
    st, exp, sst = verifier.bad_partial_key_verify_batch(fin["vv"], it["perp"], it["pk"], it["sig"], [fin["message"]])
    assert sst == 0 and (st == it["expected"]).all(), np.argwhere(st != it["expected"])[:10]
    hist = {int(k): int(c) for k, c in zip(*np.unique(st, return_counts=True))}
    assert set(hist) == {0, 5, 6, 7, 8} and 0.4 * m < hist[0] < 0.6 * m
    ast, co, keys = O.agg_coefficients(fin["vv"], fin["ids"])
    q1 = np.array([list(O.evaluate_polynomial(keys.tobytes(), 64, p + 1)[1]) for p in range(64)], dtype=np.uint8)
    assert (q1 == exp).all()
    check = np.nonzero((st != 0) | (np.arange(m) % 64 == 0))[0]
    want = np.zeros(len(check), dtype=np.uint8)
    need_pairing = []
    for n_, i in enumerate(check):
        if O.g1_decompress(bytes(it["pk"][i]))[0]:
            want[n_] = 5
        elif O.g2_decompress(bytes(it["sig"][i]))[0]:
            want[n_] = 6
        else:
            need_pairing.append(n_)
    idx = check[need_pairing]
    out = O.bls_verify_batch(it["pk"][idx], it["sig"][idx], bytes(fin["hm"]), threads=os.cpu_count() or 1)
    for n_, i, ok in zip(need_pairing, idx, out):
        want[n_] = (0 if bytes(it["pk"][i]) == bytes(exp[it["perp"][i]]) else 8) if ok == 1 else 7
    assert (st[check] == want).all(), np.argwhere(st[check] != want)[:10]
