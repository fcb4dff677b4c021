"""dvt_circuits_b200 - B200-native batch verifier for the DKG checks of metacraft-labs/dvt-circuits
`crates/dkg` (Feldman share verification, key aggregation, BLS partial-signature checks).

This package is a thin ctypes binding over the C ABI in include/dkgv.h (libdkgv.so, hand-written
CUDA for sm_100a).  There is NO CPU fallback: importing works without a GPU (so symbol/ABI tests can
run), but creating a `Verifier` raises when the library or a CUDA device is missing.
"""
from .binding import (DkgvError, Verifier, Status, lib_path, load_library, DECLARED_SYMBOLS, initial_commitment_hash, share_fd_plan, verdict_bits_to_matrix)  # noqa: F401
