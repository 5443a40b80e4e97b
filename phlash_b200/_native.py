"""ctypes binding of include/phlash_b200.h.  Loading fails loudly when the library has not been
built: there is no Python/CPU implementation to fall back to."""

from __future__ import annotations

import ctypes
import os

from phlash_b200 import build as _build

PHB_OK = 0
PHB_E_INVALID, PHB_E_CUDA, PHB_E_NOMEM, PHB_E_DATA = -1, -2, -3, -4

_lib = None

# name -> (restype, argtypes); also the list of symbols the header declares (tests check all load)
_vp, _i, _i64, _dp = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_double)
SIGNATURES = {
    "phb_abi_version": (_i, []),
    "phb_last_error": (ctypes.c_char_p, []),
    "phb_device_count": (_i, []),
    "phb_create": (_i, [_i, _vp, _i64, _i64, _i, _i, ctypes.POINTER(_vp)]),
    "phb_create_chunks": (_i, [_i, _vp, _i64, _i64, _i64, _i, _i, ctypes.POINTER(_vp)]),
    "phb_reserve": (_i, [_vp, _i64, _i64, _i64, _i]),
    "phb_allocation_count": (_i64, [_vp]),
    "phb_sample_minibatch_device": (_i, [_vp, ctypes.c_uint64, _i64, _vp, _vp]),
    "phb_set_iteration": (_i, [_vp, ctypes.c_uint64, _vp]),
    "phb_minibatch_indices": (None, [ctypes.c_uint64, ctypes.c_uint64, _i64, _i64, _vp]),
    "phb_measure_fp32_peak": (_i, [_i, _dp, _dp]),
    "phb_create_from_contig": (_i, [_i, _vp, _i64, _i64, _i64, _i64, _i, _i, ctypes.POINTER(_vp)]),
    "phb_download_data": (_i, [_vp, _vp]),
    "phb_destroy": (None, [_vp]),
    "phb_M": (_i, [_vp]),
    "phb_double_precision": (_i, [_vp]),
    "phb_num_rows": (_i64, [_vp]),
    "phb_row_length": (_i64, [_vp]),
    "phb_device": (_i, [_vp]),
    "phb_set_threads_per_pair": (_i, [_vp, _i]),
    "phb_set_store_all": (_i, [_vp, _i]),
    "phb_set_parallel_in_time": (_i, [_vp, _i]),
    "phb_set_precision_escalation": (_i, [_vp, _i]),
    "phb_num_escalated_rows": (_i64, [_vp]),
    "phb_loglik_host": (_i, [_vp, _vp, _vp, _i64, _i64, _i, _vp, _vp]),
    "phb_loglik_shared_host": (_i, [_vp, _vp, _vp, _i, _vp, _i64, _i64, _i, _vp, _vp]),
    "phb_loglik_device": (_i, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i, _vp, _vp, _vp]),
    "phb_loglik_warmup_device": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i, _vp, _vp, _vp]),
    "phb_loglik_warmup_host": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i, _vp, _vp]),
    "phb_params_from_particles": (_i, [_vp, _vp, _i64, _vp, _i, ctypes.c_double, _vp, _vp]),
    "phb_params_vjp": (_i, [_vp, _vp, _i64, _vp, _i, ctypes.c_double, _vp, _vp, _vp]),
    "phb_hmm_term_device": (_i, [_vp, _vp, _i64, _vp, _i, ctypes.c_double, _vp, _i64, _i64, ctypes.c_double, _vp, _vp, _vp]),
    "phb_hmm_term_host": (_i, [_vp, _vp, _i64, _vp, _i, ctypes.c_double, _vp, _i64, _i64, ctypes.c_double, _vp, _vp]),
    "phb_hmm_term_sums_device": (_i, [_vp, _vp, _i64, _vp, _i, ctypes.c_double, _vp, _i64, _i64, _i, _vp, _vp]),
    "phb_hmm_term_finish_device": (_i, [_vp, _vp, _i64, _vp, _i, ctypes.c_double, _vp, ctypes.c_double, _vp, _vp, _vp]),
    "phb_hmm_term_sharded_plan": (_i, [_vp, _i64, _i64, _i64, _i, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "phb_hmm_term_sharded_begin": (_i, [_vp, _vp, _i64, _vp, _i, ctypes.c_double, _vp, _i64, _i64, _i, _i, _vp, _vp]),
    "phb_hmm_term_sharded_end": (_i, [_vp, _vp, _i64, _i64, _i64, _i, _i, _vp, _vp, _vp]),
    "phb_sum_over_chunks_device": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "phb_stream": (_vp, [_vp]),
    "phb_sync": (_i, [_vp]),
    "phb_device_data": (_vp, [_vp, ctypes.POINTER(_i64)]),
    "phb_last_kernel_ms": (ctypes.c_float, [_vp]),
    "phb_last_kernel_name": (ctypes.c_char_p, [_vp]),
    "phb_launch_count": (_i64, [_vp]),
}


def library_path() -> str:
    # PHB_LIBRARY: an alternative build of the same ABI (kernel experiments)
    return os.environ.get("PHB_LIBRARY") or _build.LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} has not been built (run `python -m phlash_b200.build`); "
                "phlash_b200 has no CPU fallback"
            )
        handle = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().phb_last_error().decode()
