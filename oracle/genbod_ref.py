"""ctypes loader for oracle/libgenbod_ref.so (scalar C GENBOD; see genbod_ref.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libgenbod_ref.so")
    src = os.path.join(_HERE, "genbod_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libgenbod_ref.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.genbod_generate.restype = ctypes.c_int
        lib.genbod_generate.argtypes = [
            ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int64, ctypes.c_uint64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        lib.genbod_wtmax.restype = ctypes.c_double
        lib.genbod_wtmax.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        _LIB = lib
    return _LIB


def generate(m_top, masses, n_events, seed=1, n_threads=0, store=True):
    """Returns ``(weights (N,), particles (n,N,4), threads_used)``; weights are prod(pd)/wtmax."""
    n = len(masses)
    m = (ctypes.c_double * n)(*[float(x) for x in masses])
    weights = np.empty(n_events, dtype=np.float64)
    parts = np.empty((n, n_events, 4), dtype=np.float64) if store else None
    used = _lib().genbod_generate(
        float(m_top), n, m, int(n_events), int(seed), weights.ctypes.data,
        parts.ctypes.data if store else None, int(n_threads),
    )
    if used < 0:
        raise ValueError("Forbidden decay" if used == -2 else "bad particle count")
    return weights, parts, used


def wtmax(m_top, masses):
    n = len(masses)
    return _lib().genbod_wtmax(float(m_top), n, (ctypes.c_double * n)(*[float(x) for x in masses]))
