"""ctypes access to the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_L = None
FAITHFUL, FAST = 0, 1


def use_native_build():
    """bench.py's CPU-baseline legs: build the oracle with -march=native ON this machine (oracle/Makefile `native`) and use it from
    now on; silently keeps the portable build when that fails.  Call before the first oracle function."""
    global _L
    import subprocess
    try:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _L = None
        return _load(os.path.join(ROOT, "oracle", "liboracle_native.so"))
    except Exception:  # noqa: BLE001
        return lib()


def _load(path):
    global _L
    _L = ctypes.CDLL(path)
    _L.orc_fp_mul_count.restype = ctypes.c_uint64
    _L.orc_init()
    return _L


def lib():
    global _L
    if _L is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
        _L = ctypes.CDLL(path)
        _L.orc_fp_mul_count.restype = ctypes.c_uint64
        _L.orc_init()
    return _L


def _b(n):
    return ctypes.create_string_buffer(n)


def g1_decompress(b48):
    o = _b(48)
    st = lib().orc_g1_decompress(bytes(b48), o)
    return st, o.raw


def g2_decompress(b96):
    o = _b(96)
    st = lib().orc_g2_decompress(bytes(b96), o)
    return st, o.raw


def g1_fixed_base(s32, mode=FAST):
    o = _b(48)
    st = lib().orc_g1_fixed_base(bytes(s32), mode, o)
    return st, o.raw


def evaluate_polynomial(vv_bytes, t, id_, mode=FAST):
    o = _b(48)
    st = lib().orc_evaluate_polynomial(bytes(vv_bytes), t, id_, mode, o)
    return st, o.raw


def share_verify(vv_bytes, t, id_, secret32, mode=FAST):
    return lib().orc_share_verify(bytes(vv_bytes), t, id_, bytes(secret32), mode)


def share_matrix(vv, ids, shares, mode=FAST, threads=1, sample_stride=1):
    vv = np.ascontiguousarray(vv, dtype=np.uint8)
    n_d, t = vv.shape[0], vv.shape[1]
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    shares = np.ascontiguousarray(shares, dtype=np.uint8)
    st = np.empty((n_d, ids.shape[0]), dtype=np.uint8)
    lib().orc_share_matrix(n_d, ids.shape[0], t, vv.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p),
                           shares.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p), mode, threads, sample_stride)
    return st


def agg_coefficients(vv, ids, mode=FAST):
    vv = np.ascontiguousarray(vv, dtype=np.uint8)
    n, t = vv.shape[0], vv.shape[1]
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    co = np.empty((t, 48), dtype=np.uint8)
    keys = np.empty((ids.shape[0], 48), dtype=np.uint8)
    st = lib().orc_agg_coefficients(n, t, vv.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p), ids.shape[0], mode,
                                    co.ctypes.data_as(ctypes.c_void_p), keys.ctypes.data_as(ctypes.c_void_p))
    return st, co, keys


def lagrange(pts, ids, mode=FAST):
    pts = np.ascontiguousarray(pts, dtype=np.uint8)
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    o = _b(48)
    st = lib().orc_lagrange(ids.shape[0], pts.ctypes.data_as(ctypes.c_void_p), ids.ctypes.data_as(ctypes.c_void_p), mode, o)
    return st, o.raw


def hash_to_g2(msg):
    o = _b(96)
    lib().orc_hash_to_g2(bytes(msg), ctypes.c_size_t(len(msg)), o)
    return o.raw


def bls_verify(pk48, sig96, msg):
    return lib().orc_bls_verify(bytes(pk48), bytes(sig96), bytes(msg), ctypes.c_size_t(len(msg)))


def bls_verify_hm(pk48, sig96, hm96):
    return lib().orc_bls_verify_hm(bytes(pk48), bytes(sig96), bytes(hm96))


def bls_verify_batch(pk, sig, hm96, threads=1):
    pk = np.ascontiguousarray(pk, dtype=np.uint8)
    sig = np.ascontiguousarray(sig, dtype=np.uint8)
    out = np.empty((pk.shape[0],), dtype=np.int8)
    lib().orc_bls_verify_batch(pk.shape[0], pk.ctypes.data_as(ctypes.c_void_p), sig.ctypes.data_as(ctypes.c_void_p), bytes(hm96),
                               out.ctypes.data_as(ctypes.c_void_p), threads)
    return out


def sha256(msg):
    o = _b(32)
    lib().orc_sha256(bytes(msg), ctypes.c_size_t(len(msg)), o)
    return o.raw


def pairing_bytes(p48, q96):
    o = _b(576)
    rc = lib().orc_pairing_bytes(bytes(p48), bytes(q96), o)
    return rc, o.raw


def fp_mul_count(reset=False):
    return int(lib().orc_fp_mul_count(1 if reset else 0))


def share_matrix_shortcut(vv, shares, threads=1):
    """the GPU library's exact consistency shortcut on the CPU (ids 1..n_r in order) -> (status [n_d, n_r], dealers that fell back)"""
    vv = np.ascontiguousarray(vv, dtype=np.uint8)
    shares = np.ascontiguousarray(shares, dtype=np.uint8)
    n_d, t = vv.shape[0], vv.shape[1]
    n_r = shares.shape[1]
    st = np.empty((n_d, n_r), dtype=np.uint8)
    fb = ctypes.c_uint32(0)
    lib().orc_share_matrix_shortcut(n_d, n_r, t, vv.ctypes.data_as(ctypes.c_void_p), shares.ctypes.data_as(ctypes.c_void_p),
                                    st.ctypes.data_as(ctypes.c_void_p), threads, ctypes.byref(fb))
    return st, int(fb.value)
