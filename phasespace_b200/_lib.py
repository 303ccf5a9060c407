"""ctypes binding of libphasespace_b200.so (include/phasespace_b200.h).

There is no CPU fallback: if the library is missing the import fails loudly, and every compute call
needs a CUDA device.
"""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PS_B200_LIB") or os.path.join(_HERE, "libphasespace_b200.so")

PS_MASS_FIXED, PS_MASS_GAUSS, PS_MASS_BW, PS_MASS_RELBW, PS_MASS_USER = range(5)
PS_GEN_NORMALIZE, PS_GEN_EXACT = 1, 2
PS_E_INVALID, PS_E_CUDA, PS_E_COMPILE, PS_E_FORBIDDEN = -1, -2, -3, -4
TILE = 256  # events per tile of the unweighting mask kernel
ABI_VERSION = 4  # PS_ABI_VERSION of include/phasespace_b200.h this binding was written against


class NodeDesc(ctypes.Structure):
    _fields_ = [("parent", ctypes.c_int32), ("kind", ctypes.c_int32), ("mass", ctypes.c_double),
                ("width", ctypes.c_double)]


class PhasespaceLibraryError(ImportError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise PhasespaceLibraryError(
            f"{LIB_PATH} not found: build it with `python -m phasespace_b200.build` "
            "(phasespace_b200 has no CPU fallback)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, i64, u32, u64, dbl = c.c_void_p, c.c_int32, c.c_int64, c.c_uint32, c.c_uint64, c.c_double
    sigs = {
        "ps_tree_create": (c.c_int, [c.POINTER(NodeDesc), i32, c.POINTER(vp)]),
        "ps_tree_destroy": (None, [vp]),
        "ps_tree_n_particles": (i32, [vp]),
        "ps_tree_slot": (i32, [vp, i32]),
        "ps_tree_n_draws": (i32, [vp]),
        "ps_tree_n_user": (i32, [vp]),
        "ps_tree_user_column": (i32, [vp, i32]),
        "ps_tree_min_mass": (dbl, [vp, i32]),
        "ps_tree_w_max": (dbl, [vp]),
        "ps_tree_signature": (c.c_char_p, [vp]),
        "ps_tree_relbw_nodes": (i32, [vp]),
        "ps_tree_node_table_size": (i32, [vp]),
        "ps_tree_node_table": (i32, [vp, vp, i32, vp, vp]),
        "ps_tree_is_aot": (i32, [vp, i32, i32, u32]),
        "ps_tree_prepare": (c.c_int, [vp, i32, i32, u32, i32]),
        "ps_generate": (c.c_int, [vp, u64, u64, i64, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, u32, vp, dbl]),
        "ps_generate_indexed": (c.c_int, [vp, u64, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, u32, vp]),
        "ps_generate_host": (c.c_int, [vp, u64, u64, i64, vp, i64, vp, vp, vp, i64, u32]),
        "ps_unweight_mask": (c.c_int, [vp, u64, u64, i64, dbl, vp, vp, vp, u32, vp]),
        "ps_unweight_scan": (c.c_int, [vp, i64, vp, vp, vp]),
        "ps_unweight_index": (c.c_int, [vp, vp, u64, i64, vp, vp]),
        "ps_histogram": (c.c_int, [vp, u64, u64, i64, vp, i64, i32, dbl, dbl, dbl, dbl, dbl, vp, vp, u32, vp]),
        "ps_fonll_boost": (c.c_int, [vp, vp, i32, vp, vp, i32, dbl, dbl, u64, u64, i64, vp, vp]),
        "ps_sample_mass": (c.c_int, [i32, dbl, dbl, vp, vp, i64, u64, u64, vp, vp]),
        "ps_lorentz_boost": (c.c_int, [vp, vp, i64, i32, i64, vp, vp]),
        "ps_mass": (c.c_int, [vp, i64, vp, vp]),
        "ps_last_error": (c.c_char_p, []),
        "ps_abi_version": (i32, []),
        "ps_launch_count": (i64, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    found = lib.ps_abi_version()
    if found != ABI_VERSION:  # a stale build would be called with the wrong argument lists
        raise PhasespaceLibraryError(
            f"{LIB_PATH} has ABI version {found}, this package needs {ABI_VERSION}: rebuild it with "
            "`python -m phasespace_b200.build --force`."
        )
    return lib, tuple(sigs)


lib, EXPORTS = _load()


class ForbiddenDecayError(ValueError):
    """Kinematically forbidden decay (the reference raises tf.errors.InvalidArgumentError
    "Forbidden decay", phasespace.py:357-363)."""


def last_error() -> str:
    return lib.ps_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc == 0:
        return
    msg = last_error()
    if rc == PS_E_INVALID:
        raise ValueError(msg)
    if rc == PS_E_FORBIDDEN:
        raise ForbiddenDecayError(msg)
    raise RuntimeError(f"phasespace_b200 (code {rc}): {msg}")
