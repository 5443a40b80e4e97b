"""The C-ABI shared library: it loads, exports every symbol include/phlash_b200.h declares, and
its argument validation works without a GPU (no compute calls here)."""

import ctypes
import os
import re

import numpy as np
import pytest

from phlash_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()  # no-op when the in-tree .so is newer than its sources
    return _native.lib()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "phlash_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phb_[a-z_A-Z0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("phb_create", "phb_destroy", "phb_loglik_host", "phb_loglik_shared_host", "phb_loglik_device",
                 "phb_sync", "phb_last_error"):
        assert must in names


def test_every_declared_symbol_is_exported_and_bound(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature in phlash_b200/_native.py"
    assert sorted(_native.SIGNATURES) == declared_functions()


def test_no_torch_or_python_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "phlash_b200.h")).read()
    assert 'extern "C"' in text
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code and "at::" not in code and "PyObject" not in code and "std::" not in code


def test_abi_version(lib):
    assert lib.phb_abi_version() == 3


def test_create_rejects_bad_arguments_before_touching_cuda(lib):
    data = np.zeros((2, 32), dtype=np.int8)
    h = ctypes.c_void_p()
    rc = lib.phb_create(7, data.ctypes.data, 2, 32, 0, 0, ctypes.byref(h))
    assert rc == _native.PHB_E_INVALID and "M=7" in _native.last_error()
    rc = lib.phb_create(16, None, 2, 32, 0, 0, ctypes.byref(h))
    assert rc == _native.PHB_E_INVALID
    rc = lib.phb_create_chunks(16, data.ctypes.data, 2, 32, 32, 0, 0, ctypes.byref(h))
    assert rc == _native.PHB_E_INVALID and "overlap" in _native.last_error()
    assert not h.value
    # (the checks of the matrix itself - values >= -1, every row observed - run on the device:
    # tests/test_gpu_step_plumbing.py)


def test_minibatch_generator_on_the_host(lib):
    """phb_minibatch_indices is a pure function of (seed, iteration): with replacement, in range, and
    different across iterations / seeds (reference: jax.random.choice(subkey, N, (S,)), mcmc.py:277)."""
    from phlash_b200.gpu import minibatch_indices

    a = minibatch_indices(7, 3, 595, 5)
    assert a.shape == (5,) and a.dtype == np.int64 and (a >= 0).all() and (a < 595).all()
    np.testing.assert_array_equal(a, minibatch_indices(7, 3, 595, 5))
    assert not np.array_equal(a, minibatch_indices(7, 4, 595, 5))
    assert not np.array_equal(a, minibatch_indices(8, 3, 595, 5))
    # a prefix property: the first draws do not depend on S
    np.testing.assert_array_equal(minibatch_indices(7, 3, 595, 64)[:5], a)
    # uniform over the rows, with replacement
    big = np.concatenate([minibatch_indices(1, it, 10, 1000) for it in range(20)])
    counts = np.bincount(big, minlength=10)
    assert counts.min() > 1800 and counts.max() < 2200


def test_null_handle_is_an_error_not_a_crash(lib):
    assert lib.phb_sync(None) == _native.PHB_E_INVALID
    assert lib.phb_M(None) == 0
    lib.phb_destroy(None)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under phlash_b200/ may reference it."""
    pkg = os.path.join(ROOT, "phlash_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("# no oracle", ""), f"{fn} mentions the oracle"
