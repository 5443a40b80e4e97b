"""ctypes loader for oracle/psmc_oracle.c (test infrastructure; see that file's header)."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpsmc_oracle.so")
    src = os.path.join(_HERE, "psmc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libpsmc_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_loglik_batch.restype = ctypes.c_int
        _LIB.oracle_loglik_batch.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        _LIB.oracle_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def loglik_batch(data, row_index, params, grad=False, want_alpha=False, n_threads=0):
    """data int8 [N, L]; row_index [n_pairs]; params float64 [n_pairs, 7, M].
    Returns ll [n_pairs] (+ dlog [n_pairs, 7, M] if grad, + final alpha [n_pairs, M] if want_alpha)."""
    data = np.ascontiguousarray(data, dtype=np.int8)
    row_index = np.ascontiguousarray(row_index, dtype=np.int64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    n_pairs, rows, m = params.shape
    assert rows == 7 and row_index.shape == (n_pairs,)
    assert row_index.min() >= 0 and row_index.max() < data.shape[0]
    ll = np.empty(n_pairs)
    dlog = np.empty((n_pairs, 7, m)) if grad else None
    alpha = np.empty((n_pairs, m)) if (want_alpha and not grad) else None
    rc = lib().oracle_loglik_batch(
        m, data.ctypes.data, data.shape[1], data.shape[1], n_pairs, row_index.ctypes.data, params.ctypes.data,
        ll.ctypes.data, dlog.ctypes.data if grad else None, alpha.ctypes.data if alpha is not None else None,
        int(n_threads),
    )
    if rc != 0:
        raise RuntimeError("oracle_loglik_batch failed")
    out = (ll,)
    if grad:
        out += (dlog,)
    if alpha is not None:
        out += (alpha,)
    return out if len(out) > 1 else ll
