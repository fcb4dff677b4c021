"""Pure-Python restatement of the crates/dkg verification flows and the guests' outcome mapping
(TEST INFRASTRUCTURE ONLY).  Follows, line by line:
  crates/dkg/src/verification.rs:29-551, crates/dkg/src/dkg_math.rs:160-248,
  crates/bad_share_exchange_prove/src/main.rs:16-82, crates/finalization_prove/src/main.rs:7-33,
  crates/bad_parial_key_prove/src/main.rs:16-51, crates/dkg/src/crypto/secp256k1_keys.rs:14-64.
Every reference exit is mapped to one status code of include/dkgv.h (same numbers).
"""
import hashlib

from . import bls12_381 as B

# status codes == include/dkgv.h
OK = 0
SLASHABLE_SECRET_RANGE = 1
SLASHABLE_COMMIT_HASH = 2
SLASHABLE_DST_NOT_FOUND = 3
SLASHABLE_SHARE_MISMATCH = 4
SLASHABLE_BAD_PK = 5
SLASHABLE_BAD_SIG = 6
SLASHABLE_SIG_INVALID = 7
SLASHABLE_KEY_MISMATCH = 8
UNSLASHABLE_COMMIT_SIG = 16
UNSLASHABLE_COMMIT_HASH = 17
UNSLASHABLE_GEN_HASH = 18
UNSLASHABLE_PERP_NOT_FOUND = 19
UNSLASHABLE_SIG_INVALID = 20
ERR_LEN = 32
ERR_MSG_MISMATCH = 33
ERR_AGG_MISMATCH_VV = 34
ERR_AGG_MISMATCH_PK = 35
ERR_ZERO_ID = 36
ERR_DUP_ID = 37
PANIC_BAD_G1 = 48
PANIC_BAD_G2 = 49
PANIC_BAD_SCALAR = 50
PANIC_INDEX = 51
PANIC_PRECHECK = 52
PANIC_BAD_IDENTITY = 53

STATUS_NAMES = {v: k for k, v in list(globals().items()) if isinstance(v, int) and k.isupper()}


def is_slashable(s):
    return 1 <= s < 16


class Panic(Exception):
    def __init__(self, code):
        self.code = code


# ----------------------------------------------------------------------------- secp256k1 ECDSA (identity crypto of BlsDkgWithSecp256kCommitment)
SP = 2 ** 256 - 2 ** 32 - 977
SN = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
SG = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
      0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)


def _s_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % SP == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], SP - 2, SP) % SP
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], SP - 2, SP) % SP
    x = (lam * lam - a[0] - b[0]) % SP
    return (x, (lam * (a[0] - x) - a[1]) % SP)


def _s_mul(pt, k):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = _s_add(acc, acc)
        if bit == "1":
            acc = _s_add(acc, pt)
    return acc


def secp_pubkey_parse(b):
    """secp256k1::PublicKey::from_slice on 33 bytes; None if invalid."""
    if len(b) != 33 or b[0] not in (2, 3):
        return None
    x = int.from_bytes(b[1:], "big")
    if x >= SP:
        return None
    y2 = (x * x * x + 7) % SP
    y = pow(y2, (SP + 1) // 4, SP)
    if y * y % SP != y2:
        return None
    if (y & 1) != (b[0] & 1):
        y = SP - y
    return (x, y)


def secp_sig_parse(b):
    """ecdsa::Signature::from_compact: fails when r or s overflow the group order."""
    if len(b) != 64:
        return None
    r, s = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big")
    if r >= SN or s >= SN:
        return None
    return (r, s)


def secp_verify(pk, digest, sig):
    """secp256k1_keys.rs:51-64; libsecp256k1 rejects high-S and zero r/s; non-32-byte msg -> false."""
    if len(digest) != 32:
        return False
    r, s = sig
    if r == 0 or s == 0 or s > SN // 2:
        return False
    z = int.from_bytes(digest, "big")
    w = pow(s, SN - 2, SN)
    pt = _s_add(_s_mul(SG, z * w % SN), _s_mul(pk, r * w % SN))
    return pt is not None and pt[0] % SN == r


# ----------------------------------------------------------------------------- setups (types.rs:9-25)
class Setup:
    """identity = 'secp256k1' (BlsDkgWithSecp256kCommitment) or 'bls' (BlsDkgWithBlsCommitment)."""

    def __init__(self, identity="secp256k1", auth=False):
        self.identity, self.auth = identity, auth

    def verify_identity_sig(self, pk_bytes, msg, sig_bytes, pk_safe):
        """verification.rs:364-374 / :478-493.  pk decode failure panics in both call sites."""
        if self.identity == "secp256k1":
            pk = secp_pubkey_parse(pk_bytes)
            if pk is None:
                raise Panic(PANIC_BAD_IDENTITY)
            sig = secp_sig_parse(sig_bytes)
            if sig is None:
                raise Panic(PANIC_BAD_IDENTITY)
            return secp_verify(pk, msg, sig)
        try:
            pk = B.g1_decompress(pk_bytes)
        except ValueError:
            raise Panic(PANIC_BAD_G1)
        try:
            sig = B.g2_decompress(sig_bytes)
        except ValueError:
            raise Panic(PANIC_BAD_G2)
        return B.bls_verify(pk, sig, msg)


def hx(s):
    return bytes.fromhex(s)


# ----------------------------------------------------------------------------- dkg_math.rs
def g1_from_bytes_expect(b):
    """dkg_math.rs:24-31 -> bls_common.rs:108-112 (.expect => panic)."""
    try:
        return B.g1_decompress(b)
    except ValueError:
        raise Panic(PANIC_BAD_G1)


def evaluate_polynomial(cfs, x):
    """dkg_math.rs:160-174."""
    if len(cfs) == 0:
        return None
    if len(cfs) == 1:
        return cfs[0]
    y = cfs[-1]
    for i in range(2, len(cfs) + 1):
        y = B.g1_mul(y, x)
        y = B.g1_add(y, cfs[len(cfs) - i])
    return y


class DkgError(Exception):
    def __init__(self, code):
        self.code = code


def lagrange_interpolation(ys, xs):
    """dkg_math.rs:178-227; xs are scalars mod r."""
    k = len(xs)
    if k == 0 or k != len(ys):
        raise DkgError(ERR_LEN)
    if k == 1:
        return ys[0]
    a = 1
    for x in xs:
        a = a * x % B.R
    if a == 0:
        raise DkgError(ERR_ZERO_ID)
    r = None
    for i in range(k):
        b = xs[i]
        for j in range(k):
            if j != i:
                v = (xs[j] - xs[i]) % B.R
                if v == 0:
                    raise DkgError(ERR_DUP_ID)
                b = b * v % B.R
        li0 = a * pow(b, B.R - 2, B.R) % B.R
        r = B.g1_add(r, B.g1_mul(ys[i], li0))
    return r


def agg_coefficients(vvs, ids):
    """dkg_math.rs:230-248: returns (final keys K_j, column sums C_k)."""
    cfs = []
    for i in range(len(vvs[0])):
        s = None
        for v in vvs:
            if i >= len(v):
                raise Panic(PANIC_INDEX)
            s = B.g1_add(s, v[i])
        cfs.append(s)
    return [evaluate_polynomial(cfs, x) for x in ids], cfs


# ----------------------------------------------------------------------------- verification.rs
def compute_initial_commitment_hash(settings, base_pubkeys_hex):
    """verification.rs:151-175."""
    h = hashlib.sha256()
    h.update(hx(settings["gen_id"]))
    h.update(bytes([settings["n"] & 0xFF, settings["k"] & 0xFF, len(base_pubkeys_hex) & 0xFF]))
    for pk in base_pubkeys_hex:
        h.update(hx(pk))
    return h.digest()


def get_index_in_commitments(hashes_hex, dst_hex):
    """verification.rs:50-66."""
    srt = sorted(hx(h) for h in hashes_hex)
    d = hx(dst_hex)
    for i, h in enumerate(srt):
        if h == d:
            return i
    return None


def verify_seed_exchange_commitment(setup, hashes, seed_exchange, initial_commitment, detail=None):
    """verification.rs:68-149.  Returns a status code; raises Panic for .expect sites."""
    commitment = seed_exchange["commitment"]
    if setup.auth:
        if not setup.verify_identity_sig(hx(commitment["pubkey"]), hx(commitment["hash"]),
                                         hx(commitment["signature"]), True):
            return UNSLASHABLE_COMMIT_SIG
    ss = seed_exchange["ssecret"]
    sk = B.fr_from_be(hx(ss["shared_secret"]))
    if sk is None:
        return SLASHABLE_SECRET_RANGE
    if setup.auth:
        h = hashlib.sha256(hx(seed_exchange["initial_commitment_hash"]) + sk.to_bytes(32, "big")
                           + hx(ss["dst_base_hash"])).digest()
        if h != hx(commitment["hash"]):
            return SLASHABLE_COMMIT_HASH
    idx = get_index_in_commitments(hashes, ss["dst_base_hash"])
    if idx is None:
        return SLASHABLE_DST_NOT_FOUND
    cid = idx + 1
    cfs = [g1_from_bytes_expect(hx(pk)) for pk in initial_commitment["base_pubkeys"]]
    ev = evaluate_polynomial(cfs, cid)
    got = B.g1_mul(B.G1, sk)
    if detail is not None:
        detail["expected"] = B.g1_compress(ev).hex()
        detail["got"] = B.g1_compress(got).hex()
        detail["id"] = cid
    if B.g1_compress(got) != B.g1_compress(ev):
        return SLASHABLE_SHARE_MISMATCH
    return OK


def guest_bad_share(setup, data, detail=None):
    """crates/bad_share_exchange_prove/src/main.rs:16-82 -> (status, exit_code)."""
    try:
        ic = data["initial_commitment"]
        st = ic["settings"]
        if len(data["base_hashes"]) != st["n"]:
            raise Panic(PANIC_PRECHECK)
        if st["n"] < st["k"]:
            raise Panic(PANIC_PRECHECK)
        if ic["hash"].lower() not in [h.lower() for h in data["base_hashes"]]:
            raise Panic(PANIC_PRECHECK)
        if compute_initial_commitment_hash(st, ic["base_pubkeys"]) != hx(ic["hash"]):
            raise Panic(PANIC_PRECHECK)
        s = verify_seed_exchange_commitment(setup, data["base_hashes"], data["seeds_exchange_commitment"], ic,
                                            detail)
    except Panic as e:
        return e.code, 1
    return s, (0 if is_slashable(s) else 1)


def verify_generation_hashes(setup, generations, settings):
    """verification.rs:211-260."""
    if not generations:
        return ERR_LEN
    for g in generations[1:]:
        if g["message_cleartext"] != generations[0]["message_cleartext"]:
            return ERR_MSG_MISMATCH
    hm = B.hash_to_g2(generations[0]["message_cleartext"].encode())
    for g in generations:
        try:
            sig = B.g2_decompress(hx(g["message_signature"]))
        except ValueError:
            raise Panic(PANIC_BAD_G2)
        try:
            key = B.g1_decompress(hx(g["partial_pubkey"]))
        except ValueError:
            raise Panic(PANIC_BAD_G1)
        if not B.bls_verify_precomputed_hash(key, sig, hm):
            return UNSLASHABLE_SIG_INVALID
        if compute_initial_commitment_hash(settings, g["base_pubkeys"]) != hx(g["base_hash"]):
            return UNSLASHABLE_GEN_HASH
    return OK


def verify_generations(setup, generations, settings, agg_key_bytes, detail=None):
    """verification.rs:262-331."""
    if len(generations) != settings["n"]:
        return ERR_LEN
    s = verify_generation_hashes(setup, generations, settings)
    if s != OK:
        return s
    srt = sorted(generations, key=lambda g: hx(g["base_hash"]))  # python sort is stable
    vvs = [[g1_from_bytes_expect(hx(p)) for p in g["base_pubkeys"]] for g in srt]
    ids = list(range(1, len(srt) + 1))
    try:
        keys, cfs = agg_coefficients(vvs, ids)
        computed = lagrange_interpolation(keys, ids)
    except DkgError as e:
        return e.code
    if detail is not None:
        detail["final_keys"] = [B.g1_compress(k).hex() for k in keys]
        detail["coefficients"] = [B.g1_compress(c).hex() for c in cfs]
        detail["computed"] = B.g1_compress(computed).hex()
    if agg_key_bytes != B.g1_compress(computed):
        return ERR_AGG_MISMATCH_VV
    pks = [g1_from_bytes_expect(hx(g["partial_pubkey"])) for g in srt]
    try:
        computed = lagrange_interpolation(pks, ids)
    except DkgError as e:
        return e.code
    if agg_key_bytes != B.g1_compress(computed):
        return ERR_AGG_MISMATCH_PK
    return OK


def guest_finalization(setup, data, detail=None):
    """crates/finalization_prove/src/main.rs:7-33 -> (status, exit_code); exit 0 = ceremony valid."""
    try:
        try:
            agg = B.g1_decompress(hx(data["aggregate_pubkey"]))
        except ValueError:
            raise Panic(PANIC_BAD_G1)
        s = verify_generations(setup, data["generations"], data["settings"], B.g1_compress(agg), detail)
    except Panic as e:
        return e.code, 1
    return s, (0 if s == OK else 1)


def compute_partial_share_hash(settings, bp):
    """verification.rs:333-362."""
    d = bp["data"]
    h = hashlib.sha256()
    h.update(hx(settings["gen_id"]))
    h.update(bytes([settings["n"] & 0xFF, settings["k"] & 0xFF, len(d["base_pubkeys"]) & 0xFF]))
    for pk in d["base_pubkeys"]:
        h.update(hx(pk))
    h.update(hx(d["base_hash"]))
    h.update(hx(d["partial_pubkey"]))
    msg = d["message_cleartext"].encode()
    h.update(bytes([len(msg) & 0xFF]))
    h.update(msg)
    h.update(hx(d["message_signature"]))
    return h.digest()


def prove_wrong_final_key_generation(setup, data, detail=None):
    """verification.rs:422-466 (incl. quirk Q1 in compute_pubkey_share :523-551)."""
    bp = data["bad_partial"]
    if setup.auth:
        c = bp["commitment"]
        if compute_partial_share_hash(data["settings"], bp) != hx(c["hash"]):
            return UNSLASHABLE_COMMIT_HASH
        if not setup.verify_identity_sig(hx(c["pubkey"]), hx(c["hash"]), hx(c["signature"]), False):
            return UNSLASHABLE_COMMIT_SIG
    for g in data["generations"]:
        if compute_initial_commitment_hash(data["settings"], g["base_pubkeys"]) != hx(g["base_hash"]):
            return UNSLASHABLE_GEN_HASH
    srt = sorted(data["generations"], key=lambda g: hx(g["base_hash"]))
    perp = None
    for i, g in enumerate(srt):
        if hx(g["base_hash"]) == hx(bp["data"]["base_hash"]):
            perp = i
    if perp is None:
        return UNSLASHABLE_PERP_NOT_FOUND
    try:
        key = B.g1_decompress(hx(bp["data"]["partial_pubkey"]))
    except ValueError:
        return SLASHABLE_BAD_PK
    try:
        sig = B.g2_decompress(hx(bp["data"]["message_signature"]))
    except ValueError:
        return SLASHABLE_BAD_SIG
    if not B.bls_verify(key, sig, bp["data"]["message_cleartext"].encode()):
        return SLASHABLE_SIG_INVALID
    # verify_expected_key -> compute_pubkey_share
    vvs = [[g1_from_bytes_expect(hx(p)) for p in g["base_pubkeys"]] for g in srt]
    ids = list(range(1, len(srt) + 1))
    keys, _ = agg_coefficients(vvs, ids)
    expected = evaluate_polynomial(keys, perp + 1)  # Q1: Horner over the K_j
    if detail is not None:
        detail["expected"] = B.g1_compress(expected).hex()
    if expected != key:
        return SLASHABLE_KEY_MISMATCH
    return OK


def guest_bad_partial_key(setup, data, detail=None):
    """crates/bad_parial_key_prove/src/main.rs:16-51 -> (status, exit_code)."""
    try:
        s = prove_wrong_final_key_generation(setup, data, detail)
    except Panic as e:
        return e.code, 1
    return s, (0 if is_slashable(s) else 1)


# ----------------------------------------------------------------------------- bad-encrypted-share guest
SLASHABLE_BAD_ENCRYPTED_MSG = 9
STATUS_NAMES[SLASHABLE_BAD_ENCRYPTED_MSG] = "SLASHABLE_BAD_ENCRYPTED_MSG"


def chacha20_xor(key, nonce, data, counter=0):
    """RFC 8439 ChaCha20 (32-byte key, 12-byte nonce, 32-bit block counter)."""
    def rotl(v, n):
        return ((v << n) & 0xFFFFFFFF) | (v >> (32 - n))

    def qr(s, a, b, c, d):
        s[a] = (s[a] + s[b]) & 0xFFFFFFFF; s[d] = rotl(s[d] ^ s[a], 16)
        s[c] = (s[c] + s[d]) & 0xFFFFFFFF; s[b] = rotl(s[b] ^ s[c], 12)
        s[a] = (s[a] + s[b]) & 0xFFFFFFFF; s[d] = rotl(s[d] ^ s[a], 8)
        s[c] = (s[c] + s[d]) & 0xFFFFFFFF; s[b] = rotl(s[b] ^ s[c], 7)

    out = bytearray()
    k = [int.from_bytes(key[4 * i:4 * i + 4], "little") for i in range(8)]
    nn = [int.from_bytes(nonce[4 * i:4 * i + 4], "little") for i in range(3)]
    for blk in range((len(data) + 63) // 64):
        init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + k + [(counter + blk) & 0xFFFFFFFF] + nn
        s = list(init)
        for _ in range(10):
            qr(s, 0, 4, 8, 12); qr(s, 1, 5, 9, 13); qr(s, 2, 6, 10, 14); qr(s, 3, 7, 11, 15)
            qr(s, 0, 5, 10, 15); qr(s, 1, 6, 11, 12); qr(s, 2, 7, 8, 13); qr(s, 3, 4, 9, 14)
        ks = b"".join(((s[i] + init[i]) & 0xFFFFFFFF).to_bytes(4, "little") for i in range(16))
        chunk = data[64 * blk:64 * blk + 64]
        out += bytes(x ^ y for x, y in zip(chunk, ks))
    return bytes(out)


def guest_bad_encrypted_share(setup, data, detail=None):
    """crates/bad_encrypted_share_prove/src/main.rs:281-405 -> (status, exit_code).
    Quirk Q2 kept: once the message parses, every path ends in the final panic (exit 1)."""
    try:
        st = data["settings"]
        sender_hash = compute_initial_commitment_hash(st, data["sender_base_pubkeys"])
        hashes = [hx(h) for h in data["base_hashes"]]
        if sender_hash not in hashes:
            raise Panic(PANIC_PRECHECK)
        receiver_hash = compute_initial_commitment_hash(st, data["receiver_base_pubkeys"])
        if receiver_hash not in hashes:
            raise Panic(PANIC_PRECHECK)
        sk = B.fr_from_be(hx(data["receiver_encr_seckey"]))
        if sk is None:
            raise Panic(PANIC_BAD_SCALAR)
        rpk = B.g1_compress(B.g1_mul(B.G1, sk))
        if not data["receiver_base_pubkeys"]:
            raise Panic(PANIC_INDEX)
        if rpk != max(hx(p) for p in data["receiver_base_pubkeys"]):
            raise Panic(PANIC_PRECHECK)
        if not data["sender_base_pubkeys"]:
            raise Panic(PANIC_INDEX)
        if hx(data["sender_encr_pubkey"]) != max(hx(p) for p in data["sender_base_pubkeys"]):
            raise Panic(PANIC_PRECHECK)
        if len(hashes) != st["n"] or st["n"] < st["k"]:
            raise Panic(PANIC_PRECHECK)
        their = g1_from_bytes_expect(hx(data["sender_encr_pubkey"]))
        shared = B.g1_compress(B.g1_mul(their, sk))
        digest = hashlib.sha256(shared).digest()
        try:
            enc = bytes.fromhex(data["encrypted_data"])
        except ValueError:
            raise Panic(PANIC_PRECHECK)
        msg = chacha20_xor(digest, digest[:12], enc)
        if detail is not None:
            detail["ecdh"] = shared.hex()
            detail["plaintext_prefix"] = msg[:17].hex()
        idlen = 33 if setup.identity == "secp256k1" else 48
        siglen = 64 if setup.identity == "secp256k1" else 96
        want = 16 + 1 + 32 + (32 + idlen + siglen if setup.auth else idlen)
        if len(msg) < want:
            return SLASHABLE_BAD_ENCRYPTED_MSG, 0          # ReadError -> commit + return
        if len(msg) > want:
            raise Panic(PANIC_PRECHECK)                    # stream.finalize() assert
        gen_id, mtype, secret = msg[:16], msg[16], msg[17:49]
        if gen_id != hx(st["gen_id"]) or mtype != 3:
            return SLASHABLE_BAD_ENCRYPTED_MSG, 0
        if setup.auth:
            commitment = {"hash": msg[49:81].hex(), "pubkey": msg[81:81 + idlen].hex(), "signature": msg[81 + idlen:].hex()}
        else:
            commitment = {"pubkey": msg[49:].hex()}
        seed = {"initial_commitment_hash": sender_hash.hex(),
                "ssecret": {"shared_secret": secret.hex(), "dst_base_hash": receiver_hash.hex()}, "commitment": commitment}
        ic = {"hash": sender_hash.hex(), "settings": st, "base_pubkeys": data["sender_base_pubkeys"]}
        s = verify_seed_exchange_commitment(setup, data["base_hashes"], seed, ic, detail)
        return s, 1                                        # Q2: falls through to the final panic
    except Panic as e:
        return e.code, 1
