"""TEST INFRASTRUCTURE: ctypes binding of oracle/emd_netsimplex.c (network simplex restatement of POT's `ot.emd2`,
FilteringMergingModule.py:160-166).  Only tests/ and bench.py's CPU-baseline leg import this module."""
import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libmarsoracle.so")
_lib = None
_lock = threading.Lock()  # bench.py calls from a thread pool: one thread builds / loads


def available() -> bool:
    return os.path.exists(_PATH) or os.path.exists(os.path.join(_HERE, "Makefile"))


def _load():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_PATH):  # built by __graft_entry__.build(); a checkout that skipped it builds on first use
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        lib = ctypes.CDLL(_PATH)
        lib.mars_oracle_emd_netsimplex.restype = ctypes.c_int
        lib.mars_oracle_emd_netsimplex.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                                   ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]
        _lib = lib
        return _lib


def emd_network_simplex_c(cost, details: bool = False):
    """Optimal transport cost of `cost` [T, M] with uniform marginals 1/T and 1/M (an empty marginal is defined as 0).
    details=True also returns the optimum on the integer-scaled costs and the number of pivots."""
    c = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
    if c.ndim != 2:
        raise ValueError("cost must be a [T, M] matrix")
    t, m = c.shape
    if t == 0 or m == 0:
        return (0.0, 0.0, 0) if details else 0.0
    obj, obj_int, piv = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    rc = _load().mars_oracle_emd_netsimplex(c.ctypes.data, t, m, ctypes.byref(obj), ctypes.byref(obj_int), ctypes.byref(piv))
    if rc != 0:
        raise RuntimeError(f"mars_oracle_emd_netsimplex failed with code {rc}")
    return (obj.value, obj_int.value, piv.value) if details else obj.value
