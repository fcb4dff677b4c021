"""Deterministic synthetic DKG ceremonies: dealer polynomials, Feldman verification
vectors vv[i][k] = G * a_{i,k}, shares s_{i,j} = f_i(id_j) mod r, and seeded corruptions drawn from
the reference's own mutation taxonomy (test_vectors/*: -bad- = one bit flipped, -wrong- = another
valid value, out-of-range secret).  Set-up work (G*a, f_i(id)) runs on the GPU through the C ABI and
is excluded from every timed region."""
import numpy as np

R_INT = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
DEFAULT_SEED = 0xD1C62026


def make_coefficients(n_dealers, t, seed=DEFAULT_SEED, dealer_offset=0):
    """[n_dealers, t, 32] big-endian scalars a_{i,k} = SHA-256(seed || "coef" || i || k) mod r (seeded PRNG = SHA-256 in
    counter mode over seed || label || index; the whole range of Fr).  Dealer i's row depends only on (seed, i), so every rank of a
    multi-GPU job generates exactly its own rows."""
    import hashlib
    out = np.zeros((n_dealers, t, 32), dtype=np.uint8)
    base = hashlib.sha256(int(seed).to_bytes(8, "little") + b"coef")
    for d in range(n_dealers):
        hd = base.copy()
        hd.update(int(dealer_offset + d).to_bytes(4, "little"))
        row = bytearray(t * 32)
        for k in range(t):
            h = hd.copy()
            h.update(k.to_bytes(4, "little"))
            row[k * 32:(k + 1) * 32] = (int.from_bytes(h.digest(), "big") % R_INT).to_bytes(32, "big")
        out[d] = np.frombuffer(bytes(row), dtype=np.uint8).reshape(t, 32)
    return out


def make_session(verifier, n_dealers, n_recipients, t, seed=DEFAULT_SEED, dealer_offset=0, ids=None):
    """-> dict(vv [n_d,t,48], ids [n_r] u32, shares [n_d,n_r,32], coeffs)"""
    coeffs = make_coefficients(n_dealers, t, seed, dealer_offset)
    if ids is None:
        ids = np.arange(1, n_recipients + 1, dtype=np.uint32)
    vv, st = verifier.g1_fixed_base_mul(coeffs.reshape(-1, 32))
    assert not st.any()
    shares = verifier.fr_poly_eval(coeffs, ids)
    return {"vv": vv.reshape(n_dealers, t, 48), "ids": np.asarray(ids, dtype=np.uint32), "shares": shares, "coeffs": coeffs}


def corrupt_shares(shares, p_bad, seed=DEFAULT_SEED):
    """Bernoulli(p_bad) per share; returns (mutated copy, expected status array).
    kinds: 0 flip one bit (-> SHARE_MISMATCH), 1 replace by the share of the next dealer, same
    recipient (-> SHARE_MISMATCH), 2 secret >= r (-> SECRET_RANGE)."""
    n_d, n_r, _ = shares.shape
    rng = np.random.Generator(np.random.PCG64([seed, 0xBAD]))
    bad = rng.random((n_d, n_r)) < p_bad
    kind = rng.integers(0, 3, size=(n_d, n_r))
    bit = rng.integers(0, 248, size=(n_d, n_r))
    out = shares.copy()
    expected = np.zeros((n_d, n_r), dtype=np.uint8)
    di, ji = np.nonzero(bad)
    for d, j in zip(di.tolist(), ji.tolist()):
        k = kind[d, j]
        if k == 0 or n_d == 1 and k == 1:
            b = int(bit[d, j])
            out[d, j, 31 - b // 8] ^= 1 << (b % 8)
            # a flipped bit can push the value past r (probability ~1e-4): then the range check fires first
            expected[d, j] = 4 if int.from_bytes(out[d, j].tobytes(), "big") < R_INT else 1
        elif k == 1:
            out[d, j] = shares[(d + 1) % n_d, j]
            expected[d, j] = 4 if (shares[(d + 1) % n_d, j] != shares[d, j]).any() else 0
        else:
            out[d, j] = 0xFF
            expected[d, j] = 1
    return out, expected


def make_finalization(verifier, n, t, seed=DEFAULT_SEED, message=b"Sign with new partial key"):
    """Config 4: partial secret S_j = sum_i s_{i,j}, partial_pubkey_j = G*S_j,
    sig_j = S_j * H(m).  -> dict(vv, ids, partial_pubkeys [n,48], signatures [n,96], hm [96], message)"""
    s = make_session(verifier, n, n, t, seed)
    sh = s["shares"]
    sums = []
    for j in range(n):
        acc = 0
        for i in range(n):
            acc += int.from_bytes(sh[i, j].tobytes(), "big")
        sums.append((acc % R_INT).to_bytes(32, "big"))
    sk = np.frombuffer(b"".join(sums), dtype=np.uint8).reshape(n, 32)
    pks, st = verifier.g1_fixed_base_mul(sk)
    assert not st.any()
    hm = verifier.hash_to_g2([message])[0]
    sigs = verifier.g2_mul_batch(hm.tobytes(), sk)
    s.update({"partial_pubkeys": pks, "signatures": sigs, "hm": hm, "message": message, "partial_secrets": sk})
    return s


def make_bad_partial_items(verifier, fin, m, p_bad=0.5, seed=DEFAULT_SEED):
    """BASELINE config 5, second half: m bad-partial-key items over the finalization session `fin` (make_finalization), cycling
    the perpetrator index, a Bernoulli(p_bad) half of them corrupted with the reference's own mutation taxonomy
    (test_vectors/*/wrong_final_key_generation: -bad- = one flipped bit, -wrong- = another valid value).
    A VALID item is one prove_wrong_final_key_generation cannot slash: the accused key is the "expected key" of
    compute_pubkey_share (quirk Q1: Horner over the final keys K_j at the perpetrator's id, verification.rs:548-549) and the
    signature verifies under it - built here from the partial secrets.  kinds of corruption -> expected status:
      0 one bit of the key flipped            -> SLASHABLE_BAD_PK (5)      1 one bit of the signature flipped -> SLASHABLE_BAD_SIG (6)
      2 another perpetrator's signature       -> SLASHABLE_SIG_INVALID (7) 3 the honest partial key K_p + its honest signature
      4 another perpetrator's expected key    -> SLASHABLE_SIG_INVALID (7)   -> SLASHABLE_KEY_MISMATCH (8)
    -> dict(perp [m] u32, pk [m,48], sig [m,96], expected [m] u8, q1_keys [n,48])"""
    n = fin["ids"].shape[0]
    S = [int.from_bytes(fin["partial_secrets"][j].tobytes(), "big") for j in range(n)]
    q = []
    for p in range(n):
        acc = 0
        for s_j in reversed(S):  # evaluate_polynomial(K, id = p + 1) in the exponent
            acc = (acc * (p + 1) + s_j) % R_INT
        q.append(acc.to_bytes(32, "big"))
    qsk = np.frombuffer(b"".join(q), dtype=np.uint8).reshape(n, 32)
    qpk, st = verifier.g1_fixed_base_mul(qsk)
    assert not st.any()
    qsig = verifier.g2_mul_batch(fin["hm"].tobytes(), qsk)
    rng = np.random.Generator(np.random.PCG64([seed, 0xB9C]))
    perp = (np.arange(m, dtype=np.uint32) % n).astype(np.uint32)
    bad = rng.random(m) < p_bad
    kind = rng.integers(0, 5, size=m)
    other = ((perp + 1 + rng.integers(0, max(n - 1, 1), size=m)) % n).astype(np.uint32) if n > 1 else perp
    pk, sig = qpk[perp].copy(), qsig[perp].copy()
    expected = np.zeros((m,), dtype=np.uint8)
    k0 = np.nonzero(bad & (kind == 0))[0]
    pk[k0, 1 + rng.integers(0, 47, size=k0.size)] ^= (1 << rng.integers(0, 8, size=k0.size)).astype(np.uint8)
    expected[k0] = 5
    k1 = np.nonzero(bad & (kind == 1))[0]
    sig[k1, 1 + rng.integers(0, 95, size=k1.size)] ^= (1 << rng.integers(0, 8, size=k1.size)).astype(np.uint8)
    expected[k1] = 6
    k2 = np.nonzero(bad & (kind == 2))[0]
    sig[k2] = qsig[other[k2]]
    expected[k2] = 7 if n > 1 else 0
    k3 = np.nonzero(bad & (kind == 3))[0]
    pk[k3], sig[k3] = fin["partial_pubkeys"][perp[k3]], fin["signatures"][perp[k3]]
    expected[k3] = 8
    k4 = np.nonzero(bad & (kind == 4))[0]
    pk[k4] = qpk[other[k4]]
    expected[k4] = 7 if n > 1 else 0
    return {"perp": perp, "pk": pk, "sig": sig, "expected": expected, "q1_keys": qpk, "kind": np.where(bad, kind, -1)}
