"""ctypes binding of include/dkgv.h.  numpy arrays for host buffers, raw device pointers (e.g.
torch.Tensor.data_ptr()) for the *_dev entry points."""
import ctypes
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class DkgvError(RuntimeError):
    pass


class Status(enum.IntEnum):
    """dkgv_status of include/dkgv.h (one code per reference exit)."""
    OK = 0
    SLASHABLE_SECRET_RANGE = 1
    SLASHABLE_COMMIT_HASH = 2
    SLASHABLE_DST_NOT_FOUND = 3
    SLASHABLE_SHARE_MISMATCH = 4
    SLASHABLE_BAD_PK = 5
    SLASHABLE_BAD_SIG = 6
    SLASHABLE_SIG_INVALID = 7
    SLASHABLE_KEY_MISMATCH = 8
    SLASHABLE_BAD_ENCRYPTED_MSG = 9
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


class Report(ctypes.Structure):
    """dkgh_report of include/dkgh.h"""
    _fields_ = [("public_values", ctypes.c_void_p), ("public_cap", ctypes.c_size_t), ("public_len", ctypes.c_size_t),
                ("n_public", ctypes.c_uint32), ("have_keys", ctypes.c_int), ("expected", ctypes.c_uint8 * 48), ("got", ctypes.c_uint8 * 48)]


c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
_vp = ctypes.c_void_p
_u32 = ctypes.c_uint32

# symbol -> (restype, argtypes); must list every function include/dkgv.h declares
DECLARED_SYMBOLS = {
    "dkgv_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "dkgv_ctx_create_ex": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(_vp)]),
    "dkgv_gtab_bits": (ctypes.c_uint32, [_vp]),
    "dkgv_gtab_selfcheck": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]),
    "dkgv_ctx_destroy": (None, [_vp]),
    "dkgv_last_error": (ctypes.c_char_p, [_vp]),
    "dkgv_launch_count": (ctypes.c_uint64, [_vp]),
    "dkgv_sync": (ctypes.c_int, [_vp]),
    "dkgv_last_hot_kernel_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "dkgv_last_decode_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
    "dkgv_share_matrix_verify": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp]),
    "dkgv_share_matrix_verify_dev": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "dkgv_share_matrix_submit_dev": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dkgv_share_matrix_finish_dev": (ctypes.c_int, [_vp, _vp, _vp]),
    "dkgv_share_items_verify": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "dkgv_pack_verdicts_dev": (ctypes.c_int, [_vp, ctypes.c_uint64, _vp, _vp, _vp]),
    "dkgv_set_share_path": (ctypes.c_int, [_vp, ctypes.c_int]),
    "dkgv_last_share_path": (ctypes.c_int, [_vp]),
    "dkgv_set_share_parts": (ctypes.c_int, [_vp, _u32]),
    "dkgv_set_share_overlap": (ctypes.c_int, [_vp, ctypes.c_int]),
    "dkgv_set_share_shortcut": (ctypes.c_int, [_vp, ctypes.c_int]),
    "dkgv_last_share_continued": (ctypes.c_int, [_vp]),
    "dkgv_set_share_repair": (ctypes.c_int, [_vp, ctypes.c_int]),
    "dkgv_last_share_repaired": (ctypes.c_int, [_vp]),
    "dkgv_last_share_decoded": (ctypes.c_int, [_vp]),
    "dkgv_share_fd_plan": (ctypes.c_int, [_u32, _u32, _u32, _u32, ctypes.POINTER(_u32), ctypes.POINTER(_u32), ctypes.POINTER(ctypes.c_int32),
                                          ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_u32), ctypes.POINTER(ctypes.c_uint64),
                                          ctypes.POINTER(ctypes.c_uint64)]),
    "dkgv_last_share_phases_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "dkgv_feldman_eval": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp]),
    "dkgv_g1_fixed_base_mul": (ctypes.c_int, [_vp, _u32, _vp, _vp, _vp]),
    "dkgv_g1_decompress_check": (ctypes.c_int, [_vp, _u32, _vp, _vp]),
    "dkgv_g1_mul_batch": (ctypes.c_int, [_vp, _u32, _vp, _vp, _vp, _vp]),
    "dkgv_fr_poly_eval": (ctypes.c_int, [_vp, _u32, _u32, _vp, _u32, _vp, _vp]),
    "dkgv_agg_final_keys": (ctypes.c_int, [_vp, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp]),
    "dkgv_lagrange_at_zero": (ctypes.c_int, [_vp, _u32, _vp, _vp, _vp, _vp]),
    "dkgv_eval_points": (ctypes.c_int, [_vp, _u32, _vp, _vp, _u32, _vp, _vp]),
    "dkgv_g2_decompress_check": (ctypes.c_int, [_vp, _u32, _vp, _vp]),
    "dkgv_hash_to_g2": (ctypes.c_int, [_vp, _u32, _vp, _vp, _vp]),
    "dkgv_bls_verify_batch": (ctypes.c_int, [_vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp]),
    "dkgv_bls_verify_batch_dev": (ctypes.c_int, [_vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "dkgv_comm_unique_id": (ctypes.c_int, [_vp]),
    "dkgv_comm_init": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int]),
    "dkgv_comm_destroy": (ctypes.c_int, [_vp]),
    "dkgv_comm_world": (ctypes.c_int, [_vp]),
    "dkgv_comm_rank": (ctypes.c_int, [_vp]),
    "dkgv_all_gather_dev": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_size_t, _vp]),
    "dkgv_share_gather_words": (_u32, [_u32, _u32]),
    "dkgv_share_matrix_verify_sharded_dev": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dkgv_share_matrix_enqueue_sharded": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dkgv_share_matrix_settle_sharded": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "dkgv_share_matrix_enqueue_sharded_dev": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dkgv_share_matrix_settle_sharded_dev": (ctypes.c_int, [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "dkgv_bls_verify_batch_sharded_dev": (ctypes.c_int, [_vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "dkgv_agg_final_keys_sharded": (ctypes.c_int, [_vp, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp]),
    "dkgv_set_bls_path": (ctypes.c_int, [_vp, ctypes.c_int]),
    "dkgv_last_bls_path": (ctypes.c_int, [_vp]),
    "dkgv_last_bls_kernel_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "dkgv_bad_partial_key_verify_batch": (ctypes.c_int, [_vp, _u32, _u32, _vp, _u32, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dkgv_g2_mul_batch": (ctypes.c_int, [_vp, _u32, _vp, _vp, _vp]),
    "dkgv_initial_commitment_hashes": (ctypes.c_int, [_vp, _u32, _u32, _vp, _vp, ctypes.c_uint8, ctypes.c_uint8, _vp]),
    # include/dkgh.h
    "dkgh_execute": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                    ctypes.c_char_p, ctypes.c_size_t]),
    "dkgh_execute_report": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                           ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(Report)]),
    "dkgh_initial_commitment_hash": (None, [_vp, ctypes.c_uint8, ctypes.c_uint8, _vp, _u32, _vp]),
}


def lib_path():
    # DKGV_LIB lets experiments point at an alternative build of the SAME CUDA library
    return os.environ.get("DKGV_LIB") or os.path.join(_HERE, "libdkgv.so")


_LIB = None


def load_library():
    """dlopen libdkgv.so and set prototypes.  Raises DkgvError when the CUDA extension is missing -
    there is deliberately no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise DkgvError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    for name, (res, args) in DECLARED_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _host(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None and a.size else None


class Verifier:
    """One dkgv_ctx bound to one GPU."""

    def __init__(self, device=0, gtab_bits=0):
        """gtab_bits: window width of the fixed-base table (dkgv_ctx_create_ex; 0 = DKGV_GTAB_BITS or the library's default)"""
        self._lib = load_library()
        h = _vp()
        rc = self._lib.dkgv_ctx_create_ex(int(device), int(gtab_bits), ctypes.byref(h))
        if rc != 0:
            raise DkgvError(f"dkgv_ctx_create failed ({rc}): {self._lib.dkgv_last_error(None).decode()}")
        self._h = h
        self.device = device

    def gtab_bits(self):
        return int(self._lib.dkgv_gtab_bits(self._h))

    def gtab_selfcheck(self, first, stride, count):
        bad = ctypes.c_uint32(0)
        self._ck(self._lib.dkgv_gtab_selfcheck(self._h, int(first), int(stride), int(count), ctypes.byref(bad)))
        return int(bad.value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dkgv_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise DkgvError(f"dkgv call failed ({rc}): {self._lib.dkgv_last_error(self._h).decode()}")

    @property
    def launch_count(self):
        return int(self._lib.dkgv_launch_count(self._h))

    def last_hot_kernel_ms(self):
        ms = ctypes.c_float()
        self._ck(self._lib.dkgv_last_hot_kernel_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def last_decode_ms(self):
        """(device ms of the last verification-vector decode, whether it included the subgroup checks)"""
        ms, chk = ctypes.c_float(), ctypes.c_int()
        self._ck(self._lib.dkgv_last_decode_ms(self._h, ctypes.byref(ms), ctypes.byref(chk)))
        return float(ms.value), bool(chk.value)

    def sync(self):
        self._ck(self._lib.dkgv_sync(self._h))

    # share-matrix evaluation strategy (enum dkgv_share_path)
    PATH_AUTO, PATH_HORNER, PATH_FDIFF = 0, 1, 2

    def set_share_path(self, mode):
        self._ck(self._lib.dkgv_set_share_path(self._h, int(mode)))

    def set_share_overlap(self, on):
        """1 / True (default): one internal stream per part; 0 / False: one stream, phase after phase"""
        self._ck(self._lib.dkgv_set_share_overlap(self._h, int(on)))

    def set_share_shortcut(self, on):
        """True (default): ids beyond t only for dealer groups that fail the scalar-side consistency conditions"""
        self._ck(self._lib.dkgv_set_share_shortcut(self._h, int(bool(on))))

    def set_share_repair(self, on):
        """True (default): inconsistent dealers are first decoded as Reed-Solomon words (scalar arithmetic), evaluation only for the rest"""
        self._ck(self._lib.dkgv_set_share_repair(self._h, int(bool(on))))

    @property
    def last_share_repaired(self):
        return int(self._lib.dkgv_last_share_repaired(self._h))

    @property
    def last_share_continued(self):
        return int(self._lib.dkgv_last_share_continued(self._h))

    @property
    def last_share_decoded(self):
        """1: the last share-matrix call decoded the commitments; 0: settled against their compressed encodings"""
        return int(self._lib.dkgv_last_share_decoded(self._h))

    def set_share_parts(self, parts):
        """parts per dealer polynomial on the finite-difference path (0 = planner's choice)"""
        self._ck(self._lib.dkgv_set_share_parts(self._h, int(parts)))

    @property
    def last_share_path(self):
        return int(self._lib.dkgv_last_share_path(self._h))

    def last_share_phases_ms(self):
        """(seed Horner, differences, extension, G*s compare) device ms of the last finite-difference run"""
        ms = (ctypes.c_float * 4)()
        self._ck(self._lib.dkgv_last_share_phases_ms(self._h, ms))
        return [float(x) for x in ms]

    # ---- share verification ------------------------------------------------------------------
    def share_matrix_verify(self, vv, ids, shares):
        """vv [n_d, t, 48] u8, ids [n_r] u32, shares [n_d, n_r, 32] u8 -> status [n_d, n_r] u8"""
        vv = np.ascontiguousarray(vv, dtype=np.uint8)
        n_d, t = vv.shape[0], vv.shape[1]
        ids = _host(ids, np.uint32)
        n_r = ids.shape[0]
        shares = _host(shares, np.uint8, (n_d, n_r, 32))
        status = np.empty((n_d, n_r), dtype=np.uint8)
        self._ck(self._lib.dkgv_share_matrix_verify(self._h, n_d, n_r, t, _p(vv), _p(ids), _p(shares), _p(status)))
        return status

    def share_matrix_verify_dev(self, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, stream=None):
        """device pointers (ints); asynchronous"""
        self._ck(self._lib.dkgv_share_matrix_verify_dev(self._h, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, stream))

    def share_matrix_submit_dev(self, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, d_flags2=None, stream=None):
        """queues the default path WITHOUT synchronising; d_flags2 (device, 2 x u32; None = ctx-owned) receives
        [ids are no permutation of 1..n, dealers the consistency shortcut could not settle]"""
        self._ck(self._lib.dkgv_share_matrix_submit_dev(self._h, n_d, n_r, t, d_vv, d_ids, d_shares, d_status, d_flags2, stream))

    def share_matrix_finish_dev(self, h_flags2=None, stream=None):
        """h_flags2: the two flag words read back by the caller (numpy u32[2]) or None (read here: one synchronisation);
        queues the evaluation of the unsettled dealer groups / the Horner route when the flags ask for it"""
        fl = _host(h_flags2, np.uint32, (2,)) if h_flags2 is not None else None
        self._ck(self._lib.dkgv_share_matrix_finish_dev(self._h, _p(fl) if fl is not None else None, stream))

    def share_items_verify(self, vv, ids, item_dealer, item_recipient, secrets):
        """sparse (dealer, recipient-column) items of one session -> status [m] u8"""
        vv = np.ascontiguousarray(vv, dtype=np.uint8)
        n_d, t = vv.shape[0], vv.shape[1]
        ids = _host(ids, np.uint32)
        it_d, it_r = _host(item_dealer, np.uint32), _host(item_recipient, np.uint32)
        m = it_d.shape[0]
        secrets = _host(secrets, np.uint8, (m, 32))
        status = np.empty((m,), dtype=np.uint8)
        self._ck(self._lib.dkgv_share_items_verify(self._h, n_d, ids.shape[0], t, _p(vv), _p(ids), m, _p(it_d), _p(it_r), _p(secrets), _p(status)))
        return status

    def pack_verdicts_dev(self, n, d_status, d_bits, stream=None):
        """device pointers; bit i of d_bits (u32 words) = status[i] != OK; asynchronous"""
        self._ck(self._lib.dkgv_pack_verdicts_dev(self._h, int(n), d_status, d_bits, stream))

    def feldman_eval(self, vv, ids):
        vv = np.ascontiguousarray(vv, dtype=np.uint8)
        n_d, t = vv.shape[0], vv.shape[1]
        ids = _host(ids, np.uint32)
        out = np.empty((n_d, ids.shape[0], 48), dtype=np.uint8)
        row = np.empty((n_d,), dtype=np.uint8)
        self._ck(self._lib.dkgv_feldman_eval(self._h, n_d, ids.shape[0], t, _p(vv), _p(ids), _p(out), _p(row)))
        return out, row

    def g1_fixed_base_mul(self, scalars):
        scalars = _host(scalars, np.uint8)
        m = scalars.shape[0]
        out = np.empty((m, 48), dtype=np.uint8)
        st = np.empty((m,), dtype=np.uint8)
        self._ck(self._lib.dkgv_g1_fixed_base_mul(self._h, m, _p(scalars), _p(out), _p(st)))
        return out, st

    def g1_mul_batch(self, pts, scalars):
        pts = np.ascontiguousarray(pts, dtype=np.uint8).reshape(-1, 48)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros((pts.shape[0], 48), dtype=np.uint8)
        st = np.zeros((pts.shape[0],), dtype=np.uint8)
        self._ck(self._lib.dkgv_g1_mul_batch(self._h, pts.shape[0], _p(pts), _p(scalars), _p(out), _p(st)))
        return out, st

    def g1_decompress_check(self, pts):
        pts = _host(pts, np.uint8)
        m = pts.shape[0]
        st = np.empty((m,), dtype=np.uint8)
        self._ck(self._lib.dkgv_g1_decompress_check(self._h, m, _p(pts), _p(st)))
        return st

    def fr_poly_eval(self, coeffs, ids):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint8)
        n_d, t = coeffs.shape[0], coeffs.shape[1]
        ids = _host(ids, np.uint32)
        out = np.empty((n_d, ids.shape[0], 32), dtype=np.uint8)
        self._ck(self._lib.dkgv_fr_poly_eval(self._h, n_d, t, _p(coeffs), ids.shape[0], _p(ids), _p(out)))
        return out

    # ---- aggregation / Lagrange -----------------------------------------------------------------
    def agg_final_keys(self, vv, ids, want_coeffs=True):
        """vv [n, t, 48], ids [m] -> (status, coeffs [t,48], keys [m,48])"""
        vv = np.ascontiguousarray(vv, dtype=np.uint8)
        n, t = vv.shape[0], vv.shape[1]
        ids = _host(ids, np.uint32)
        co = np.zeros((t, 48), dtype=np.uint8)
        keys = np.zeros((ids.shape[0], 48), dtype=np.uint8)
        st = ctypes.c_uint8(0)
        self._ck(self._lib.dkgv_agg_final_keys(self._h, n, t, _p(vv), _p(ids), ids.shape[0], _p(co) if want_coeffs else None, _p(keys),
                                               ctypes.cast(ctypes.byref(st), _vp)))
        return int(st.value), co, keys

    def lagrange_at_zero(self, pts, ids):
        pts = np.ascontiguousarray(pts, dtype=np.uint8).reshape(-1, 48)
        ids = _host(ids, np.uint32)
        out = np.zeros((48,), dtype=np.uint8)
        st = ctypes.c_uint8(0)
        k = ids.shape[0]
        if pts.shape[0] != k:
            return int(Status.ERR_LEN), bytes(48)
        self._ck(self._lib.dkgv_lagrange_at_zero(self._h, k, _p(pts), _p(ids), _p(out), ctypes.cast(ctypes.byref(st), _vp)))
        return int(st.value), out.tobytes()

    def eval_points(self, coeffs, ids):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint8).reshape(-1, 48)
        ids = _host(ids, np.uint32)
        out = np.zeros((ids.shape[0], 48), dtype=np.uint8)
        st = ctypes.c_uint8(0)
        self._ck(self._lib.dkgv_eval_points(self._h, coeffs.shape[0], _p(coeffs), _p(ids), ids.shape[0], _p(out),
                                            ctypes.cast(ctypes.byref(st), _vp)))
        return int(st.value), out

    # ---- BLS checks --------------------------------------------------------------------------------
    def g2_decompress_check(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.uint8).reshape(-1, 96)
        st = np.empty((pts.shape[0],), dtype=np.uint8)
        self._ck(self._lib.dkgv_g2_decompress_check(self._h, pts.shape[0], _p(pts), _p(st)))
        return st

    def hash_to_g2(self, msgs):
        """list of bytes -> [m, 96] compressed G2 points"""
        offs = np.zeros((len(msgs) + 1,), dtype=np.uint32)
        for i, m in enumerate(msgs):
            offs[i + 1] = offs[i] + len(m)
        blob = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8).copy()
        out = np.zeros((len(msgs), 96), dtype=np.uint8)
        self._ck(self._lib.dkgv_hash_to_g2(self._h, len(msgs), _p(blob), _p(offs), _p(out)))
        return out

    def bls_verify_batch(self, pk, sig, hm, hm_idx=None):
        pk = np.ascontiguousarray(pk, dtype=np.uint8).reshape(-1, 48)
        sig = np.ascontiguousarray(sig, dtype=np.uint8).reshape(-1, 96)
        hm = np.ascontiguousarray(hm, dtype=np.uint8).reshape(-1, 96)
        idx = _host(hm_idx, np.uint32) if hm_idx is not None else None
        st = np.empty((pk.shape[0],), dtype=np.uint8)
        self._ck(self._lib.dkgv_bls_verify_batch(self._h, pk.shape[0], _p(pk), _p(sig), hm.shape[0], _p(hm),
                                                 _p(idx) if idx is not None else None, _p(st)))
        return st

    def bad_partial_key_verify_batch(self, vv, perp, pk, sig, msgs, msg_idx=None):
        """vv [n,t,48] (base_hash-sorted generations), perp [m] u32, pk [m,48], sig [m,96], msgs = list of bytes, msg_idx [m] or None
        -> (status [m] u8, expected keys [n,48], session status)"""
        vv = np.ascontiguousarray(vv, dtype=np.uint8)
        n, t = vv.shape[0], vv.shape[1]
        perp = _host(perp, np.uint32)
        m = perp.shape[0]
        pk = _host(pk, np.uint8, (m, 48))
        sig = _host(sig, np.uint8, (m, 96))
        offs = np.zeros((len(msgs) + 1,), dtype=np.uint32)
        for i, mm in enumerate(msgs):
            offs[i + 1] = offs[i] + len(mm)
        blob = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8).copy()
        idx = _host(msg_idx, np.uint32, (m,)) if msg_idx is not None else None
        st = np.empty((m,), dtype=np.uint8)
        exp = np.zeros((n, 48), dtype=np.uint8)
        sst = ctypes.c_uint8(0)
        self._ck(self._lib.dkgv_bad_partial_key_verify_batch(self._h, n, t, _p(vv), m, _p(perp), _p(pk), _p(sig), len(msgs), _p(blob), _p(offs),
                                                             _p(idx) if idx is not None else None, _p(st), _p(exp),
                                                             ctypes.cast(ctypes.byref(sst), _vp)))
        return st, exp, int(sst.value)

    # ---- multi-GPU (include/dkgv.h dkgv_comm_*): one Verifier per process / GPU
    @staticmethod
    def comm_unique_id():
        """rank 0: 128 bytes to hand to every rank"""
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        if lib.dkgv_comm_unique_id(ctypes.cast(buf, _vp)) != 0:
            raise DkgvError("dkgv_comm_unique_id failed (NCCL not loadable)")
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._ck(self._lib.dkgv_comm_init(self._h, ctypes.cast(buf, _vp), int(rank), int(world)))

    def comm_destroy(self):
        self._ck(self._lib.dkgv_comm_destroy(self._h))

    @property
    def comm_world(self):
        return int(self._lib.dkgv_comm_world(self._h))

    @property
    def comm_rank(self):
        return int(self._lib.dkgv_comm_rank(self._h))

    def all_gather_dev(self, d_send, d_recv, nbytes, stream=None):
        self._ck(self._lib.dkgv_all_gather_dev(self._h, d_send, d_recv, int(nbytes), stream))

    def share_gather_words(self, n_local, n_r):
        return int(self._lib.dkgv_share_gather_words(n_local, n_r))

    def share_matrix_verify_sharded_dev(self, n_local, n_r, t, d_vv, d_ids, d_shares, d_status, d_gather, stream=None):
        """this rank's dealer row block; d_gather [world, share_gather_words(n_local, n_r)] u32 on the device"""
        self._ck(self._lib.dkgv_share_matrix_verify_sharded_dev(self._h, n_local, n_r, t, d_vv, d_ids, d_shares, d_status, d_gather, stream))

    def share_matrix_enqueue_sharded_dev(self, n_local, n_r, t, d_vv, d_ids, d_shares, d_status, d_gather, h_flags, stream=None):
        """pipelined form: queued without synchronising; h_flags = pinned host memory, 2 * world u32 (settle reads it after a sync)"""
        self._ck(self._lib.dkgv_share_matrix_enqueue_sharded_dev(self._h, n_local, n_r, t, d_vv, d_ids, d_shares, d_status, d_gather, h_flags, stream))

    def share_matrix_settle_sharded_dev(self, n_local, n_r, t, d_vv, d_ids, d_shares, d_status, d_gather, h_flags, stream=None):
        """after the stream has been synchronised: True when the ceremony had to be run again (corrupted shares / foreign ids)"""
        reran = ctypes.c_int(0)
        self._ck(self._lib.dkgv_share_matrix_settle_sharded_dev(self._h, n_local, n_r, t, d_vv, d_ids, d_shares, d_status, d_gather, h_flags, stream,
                                                               ctypes.byref(reran)))
        return bool(reran.value)

    def share_matrix_enqueue_sharded(self, n_local, n_r, t, h_vv, h_ids, h_shares, h_status, h_gather, h_flags):
        """host-buffer form (addresses of PINNED host memory; h_gather may be None): queued on the ctx's own stream, no synchronisation"""
        self._ck(self._lib.dkgv_share_matrix_enqueue_sharded(self._h, n_local, n_r, t, h_vv, h_ids, h_shares, h_status, h_gather, h_flags))

    def share_matrix_settle_sharded(self, n_local, n_r, t, h_vv, h_ids, h_shares, h_status, h_gather, h_flags):
        """after sync(): True when the ceremony had to be run again"""
        reran = ctypes.c_int(0)
        self._ck(self._lib.dkgv_share_matrix_settle_sharded(self._h, n_local, n_r, t, h_vv, h_ids, h_shares, h_status, h_gather, h_flags, ctypes.byref(reran)))
        return bool(reran.value)

    def bls_verify_batch_sharded_dev(self, m_local, d_pk, d_sig, n_hm, d_hm, d_hm_idx, d_status_all, stream=None):
        self._ck(self._lib.dkgv_bls_verify_batch_sharded_dev(self._h, m_local, d_pk, d_sig, n_hm, d_hm, d_hm_idx, d_status_all, stream))

    def agg_final_keys_sharded(self, vv_local, ids):
        """this rank's generations [n_local, t, 48] -> (status, coeffs [t,48], keys [m,48]) on every rank"""
        vv = np.ascontiguousarray(vv_local, dtype=np.uint8)
        n, t = vv.shape[0], vv.shape[1]
        ids = _host(ids, np.uint32)
        co = np.zeros((t, 48), dtype=np.uint8)
        keys = np.zeros((ids.shape[0], 48), dtype=np.uint8)
        st = ctypes.c_uint8(0)
        self._ck(self._lib.dkgv_agg_final_keys_sharded(self._h, n, t, _p(vv), _p(ids), ids.shape[0], _p(co), _p(keys),
                                                       ctypes.cast(ctypes.byref(st), _vp)))
        return int(st.value), co, keys

    BLS_AUTO, BLS_VM, BLS_THREAD = 0, 1, 2

    def set_bls_path(self, mode):
        """enum dkgv_bls_path: AUTO / VM (pairing VM, several warps per 32 checks) / THREAD (one thread per check)"""
        self._ck(self._lib.dkgv_set_bls_path(self._h, int(mode)))

    @property
    def last_bls_path(self):
        return int(self._lib.dkgv_last_bls_path(self._h))

    def last_bls_kernel_ms(self):
        ms = ctypes.c_float()
        self._ck(self._lib.dkgv_last_bls_kernel_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def initial_commitment_hashes(self, vv, gen_id, n, k):
        """base hashes of all dealers at once: vv [n_d, t, 48] -> [n_d, 32]"""
        vv = np.ascontiguousarray(vv, dtype=np.uint8)
        gid = np.frombuffer(bytes(gen_id), dtype=np.uint8)
        out = np.zeros((vv.shape[0], 32), dtype=np.uint8)
        self._ck(self._lib.dkgv_initial_commitment_hashes(self._h, vv.shape[0], vv.shape[1], _p(vv), _p(gid), n & 0xFF, k & 0xFF, _p(out)))
        return out

    def g2_mul_batch(self, base96, scalars):
        scalars = np.ascontiguousarray(scalars, dtype=np.uint8).reshape(-1, 32)
        base = np.frombuffer(bytes(base96), dtype=np.uint8)
        out = np.zeros((scalars.shape[0], 96), dtype=np.uint8)
        self._ck(self._lib.dkgv_g2_mul_batch(self._h, scalars.shape[0], _p(base), _p(scalars), _p(out)))
        return out

    # ---- dkg_prover_host `execute` contract (include/dkgh.h) ------------------------------------------
    def execute(self, type_, json_text, auth=False, bls_identity=False):
        """-> (exit_code, status, message)"""
        st = ctypes.c_int(0)
        msg = ctypes.create_string_buffer(512)
        if isinstance(json_text, str):
            json_text = json_text.encode()
        code = self._lib.dkgh_execute(self._h, type_.encode(), json_text, int(auth), int(bls_identity), ctypes.byref(st), msg, 512)
        return int(code), int(st.value), msg.value.decode(errors="replace")


    def execute_report(self, type_, json_text, auth=False, bls_identity=False):
        """-> (exit_code, status, message, public values [bytes, in commit order], (expected48, got48) or None)"""
        st = ctypes.c_int(0)
        msg = ctypes.create_string_buffer(512)
        pub = ctypes.create_string_buffer(1 << 20)
        rep = Report()
        rep.public_values = ctypes.cast(pub, ctypes.c_void_p)
        rep.public_cap = len(pub)
        if isinstance(json_text, str):
            json_text = json_text.encode()
        code = self._lib.dkgh_execute_report(self._h, type_.encode(), json_text, int(auth), int(bls_identity), ctypes.byref(st), msg, 512,
                                             ctypes.byref(rep))
        vals, raw, o = [], pub.raw, 0
        for _ in range(rep.n_public):
            ln = int.from_bytes(raw[o:o + 4], "little")
            vals.append(raw[o + 4:o + 4 + ln])
            o += 4 + ln
        keys = (bytes(rep.expected), bytes(rep.got)) if rep.have_keys else None
        return int(code), int(st.value), msg.value.decode(errors="replace"), vals, keys


def verdict_bits_to_matrix(words, n_dealers, n_recipients):
    """Unpack the bitmask of dkgv_pack_verdicts_dev (u32 words, bit i%32 of word i//32 = share i is NOT ok,
    shares in row-major (dealer, recipient) order) into a bool matrix [n_dealers, n_recipients]; the
    bad-participant list of the share proof type is `np.nonzero(m.any(axis=1))[0]`."""
    w = np.ascontiguousarray(words).view(np.uint8)
    bits = np.unpackbits(w, bitorder="little")[:n_dealers * n_recipients]
    return bits.reshape(n_dealers, n_recipients).astype(bool)


def share_fd_plan(t, n_recipients, parts=0, n_opt=0):
    """dkgv_share_fd_plan -> dict(use, parts, h, lo, hi, steps, modmul_fd, modmul_horner) (per dealer, evaluation only);
    n_opt = ids assumed evaluated in the group (0 = all; t = what the share matrix uses with the consistency shortcut on)"""
    lib = load_library()
    m, h, lo, hi, steps = _u32(), _u32(), ctypes.c_int32(), ctypes.c_int32(), _u32()
    cf, ch = ctypes.c_uint64(), ctypes.c_uint64()
    use = lib.dkgv_share_fd_plan(t, n_recipients, parts, n_opt, ctypes.byref(m), ctypes.byref(h), ctypes.byref(lo), ctypes.byref(hi),
                                 ctypes.byref(steps), ctypes.byref(cf), ctypes.byref(ch))
    return {"use": use == 1, "exists": use >= 0, "parts": m.value, "h": h.value, "lo": lo.value, "hi": hi.value, "steps": steps.value,
            "modmul_fd": cf.value, "modmul_horner": ch.value}


def initial_commitment_hash(gen_id, n, k, base_pubkeys):
    """compute_initial_commitment_hash (verification.rs:151-175); base_pubkeys [count, 48] u8"""
    lib = load_library()
    pk = np.ascontiguousarray(base_pubkeys, dtype=np.uint8).reshape(-1, 48)
    gid = np.frombuffer(bytes(gen_id), dtype=np.uint8)
    out = np.zeros((32,), dtype=np.uint8)
    lib.dkgh_initial_commitment_hash(_p(gid), n & 0xFF, k & 0xFF, _p(pk) if pk.size else None, pk.shape[0], _p(out))
    return out.tobytes()
