"""The reference's OWN CUDA kernels (KERNEL_SRC of src/phlash/gpu.py compiled by
oracle/build_ref.py into oracle/_ref/*.cubin) run on the same GPU:
  * in double precision they are a second oracle - they pin oracle/psmc_oracle.* against the
    reference's own implementation (not just against its Python twin);
  * in single precision they are what our fp32 kernel replaces."""

import numpy as np
import pytest

from conftest import fixture_data
from oracle import c_oracle, psmc_oracle as orc, ref_cuda

pytestmark = pytest.mark.gpu


def oracle_eval(data, inds, pa):
    B, S = pa.shape[:2]
    ll, dlog = c_oracle.loglik_batch(data, np.tile(inds, B), pa.reshape(B * S, 7, -1).astype(np.float64), grad=True)
    return ll.reshape(B, S), dlog.reshape(B, S, 7, -1)


needs_ref = pytest.mark.skipif(not ref_cuda.available(16, True), reason="oracle/_ref cubins not built")


@needs_ref
@pytest.mark.parametrize("missing", [False, True])
def test_oracle_matches_reference_cuda_fp64(golden, seed, missing):
    data, miss = fixture_data(seed)
    data = miss if missing else data
    pps = np.concatenate([golden["dm16_pp"][None], golden["part_pp"][1:4]])
    inds = np.array([0, 7, 3])
    pa = np.broadcast_to(pps[:, None], (4, 3, 7, 16)).copy()
    ref = ref_cuda.ReferenceKernel(16, data, double_precision=True)
    ll_ref, dlog_ref = ref(pa, inds, grad=True)
    ll_o, dlog_o = oracle_eval(data, inds, pa)
    np.testing.assert_allclose(ll_o, ll_ref, rtol=1e-11)
    np.testing.assert_allclose(dlog_o, dlog_ref, rtol=1e-7, atol=1e-12)
    # the reference's forward-only kernel shares the parameters of s == 0 inside a block
    # (gpu.py:548-551); with identical parameters across s it must agree with its grad kernel
    np.testing.assert_allclose(ref(pa, inds, grad=False), ll_ref, rtol=1e-12)
    ref.close()


@needs_ref
def test_ours_matches_reference_cuda(golden):
    """fp64 vs fp64 at the reference's own test tolerance (tests/test_gpu.py:59-64), and our fp32
    kernel against the reference's fp32 kernel at the north-star tolerance."""
    from phlash_b200.gpu import PSMCKernel
    from phlash_b200.params import PSMCParams

    _, data = fixture_data(1)
    pps = golden["part_pp"][:5]
    inds = np.arange(10)
    pa = np.broadcast_to(pps[:, None], (5, 10, 7, 16)).copy()
    ref64 = ref_cuda.ReferenceKernel(16, data, double_precision=True)
    ll_ref, dlog_ref = ref64(pa, inds, grad=True)
    ours64 = PSMCKernel(16, data, double_precision=True, num_gpus=1)
    ll, dll = ours64(PSMCParams.from_block(pa), inds, grad=True)
    np.testing.assert_allclose(ll, ll_ref, atol=1e-8, rtol=1e-11)
    np.testing.assert_allclose(dll.to_block(), dlog_ref, atol=1e-8, rtol=1e-7)
    # fp32: both kernels get the same fp32-rounded parameters; the truth is the reference's fp64
    # kernel evaluated at those rounded values
    pa32 = pa.astype(np.float32)
    ll_t, dlog_t = ref64(pa32.astype(np.float64), inds, grad=True)
    ref32 = ref_cuda.ReferenceKernel(16, data, double_precision=False)
    ll_r32, dlog_r32 = ref32(pa32, inds, grad=True)
    ours32 = PSMCKernel(16, data, double_precision=False, num_gpus=1)
    ll32, dll32 = ours32(PSMCParams.from_block(pa32), inds, grad=True)
    np.testing.assert_allclose(ll32, ll_t, rtol=1e-5)
    # our fp32 result is at least as close to the fp64 truth as the reference's fp32 kernel
    err_ours = np.abs(ll32 - ll_t).max()
    err_ref = np.abs(ll_r32 - ll_t).max()
    assert err_ours <= max(2 * err_ref, 1e-5 * np.abs(ll_t).max())
    scale = np.abs(dlog_t).max(axis=-1, keepdims=True)
    assert np.all(np.abs(dll32.to_block() - dlog_t) <= 1e-4 * np.abs(dlog_t) + 1e-7 * scale)
    g_err_ours = (np.abs(dll32.to_block() - dlog_t) / (np.abs(dlog_t) + 1e-3 * scale)).max()
    g_err_ref = (np.abs(dlog_r32 - dlog_t) / (np.abs(dlog_t) + 1e-3 * scale)).max()
    assert g_err_ours <= max(2 * g_err_ref, 1e-4)
    ref64.close()
    ref32.close()
