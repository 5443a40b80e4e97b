"""Parity of the CUDA path (through the C ABI / PSMCKernel) against the fp64 oracle.
Tolerances: BASELINE.json north_star - loglik 1e-5 relative, gradients 1e-4 relative in fp32.
Mirrors the reference's tests/test_gpu.py and tests/test_model.py."""

import numpy as np
import pytest

from conftest import fixture_data
from oracle import c_oracle, psmc_oracle as orc

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-5
GRAD_RTOL = 1e-4


def grad_close(got, want, rtol, what=""):
    """Relative to each entry, with an absolute floor tied to the largest entry of the same row
    (entries many orders of magnitude below the row's scale carry no information in fp32)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max(axis=-1, keepdims=True)
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + rtol * 1e-3 * scale
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} entries off; worst rel {np.max(err / np.maximum(np.abs(want), 1e-300)):.3e}"


def make_kernel(M, data, double_precision=False, T=0):
    from phlash_b200.gpu import PSMCKernel

    k = PSMCKernel(M=M, data=data, double_precision=double_precision, num_gpus=1)
    if T:
        k.set_threads_per_pair(T)
    return k


def oracle_eval(data, inds, pa):
    B, S = pa.shape[:2]
    rows = np.tile(np.asarray(inds), B)
    ll, dlog = c_oracle.loglik_batch(data, rows, pa.reshape(B * S, 7, -1).astype(np.float64), grad=True)
    return ll.reshape(B, S), dlog.reshape(B, S, 7, -1)


@pytest.mark.parametrize("T", [0, 2, 4])
@pytest.mark.parametrize("missing", [False, True])
def test_reference_fixture_fp32(golden, seed, missing, T):
    """tests/test_gpu.py:44-64 of the reference: CUDA vs the HMM definition, with and without
    missing data; ll and every gradient leaf."""
    from phlash_b200.params import PSMCParams

    data, miss = fixture_data(seed)
    data = miss if missing else data
    pp = golden["dm16_pp"]
    kern = make_kernel(16, data, T=T)
    inds = np.arange(len(data))
    ll, dll = kern(PSMCParams.from_block(pp), inds, grad=True)
    ll_nograd = kern(PSMCParams.from_block(pp), inds, grad=False)
    ref_ll, ref_dlog = oracle_eval(data, inds, np.broadcast_to(pp, (1, len(inds), 7, 16)))
    np.testing.assert_allclose(ll, ref_ll[0], rtol=LL_RTOL)
    np.testing.assert_allclose(ll_nograd, ll, rtol=1e-12)  # test_eq_grad_nograd
    grad_close(dll.to_block(), ref_dlog[0], GRAD_RTOL, f"T={T}")
    # the golden ll of the reference's own psmc_ll for rows 0..2
    off = 3 if missing else 0
    np.testing.assert_allclose(ll[:3], golden[f"hmm_ll_s{seed}"][off : off + 3], rtol=LL_RTOL)


@pytest.mark.parametrize("T", [0, 4])
def test_reference_fixture_fp64(golden, seed, T):
    """double_precision=True against the oracle at the reference's own tolerance
    (tests/test_gpu.py:59-64: atol 1e-8, rtol 1e-5) - and much tighter."""
    from phlash_b200.params import PSMCParams

    _, data = fixture_data(seed)
    pp = golden["part_pp"][2]
    kern = make_kernel(16, data, double_precision=True, T=T)
    inds = np.array([4, 0, 9, 4])
    ll, dll = kern(PSMCParams.from_block(pp), inds, grad=True)
    ref_ll, ref_dlog = oracle_eval(data, inds, np.broadcast_to(pp, (1, 4, 7, 16)))
    np.testing.assert_allclose(ll, ref_ll[0], rtol=1e-12)
    np.testing.assert_allclose(dll.to_block(), ref_dlog[0], rtol=1e-8, atol=1e-12)
