"""Fused warm-up (model.py:50-57): log p(chunk | pi_s) with pi_s the filtered distribution after the
`overlap` warm-up bins started from the stationary pi, and its gradient THROUGH the warm-up."""

import numpy as np
import pytest

from oracle import c_oracle, psmc_oracle as orc

PATTERN16 = "14*1+1*2"


def oracle_warmup(chunks, inds, pps, overlap, kernel_dtype=np.float64, with_parts=False):
    """Oracle of the fused evaluation by the identity LL(warm-up + chunk) - LL(warm-up)."""
    B, S = len(pps), len(inds)
    pa = np.broadcast_to(np.asarray(pps).astype(kernel_dtype).astype(np.float64)[:, None], (B, S, 7, pps.shape[-1]))
    pa = pa.reshape(B * S, 7, -1)
    rows = np.tile(inds, B)
    ll_f, g_f = c_oracle.loglik_batch(chunks, rows, pa, grad=True)
    ll_w, g_w = c_oracle.loglik_batch(np.ascontiguousarray(chunks[:, :overlap]), rows, pa, grad=True)
    if with_parts:
        return (ll_f - ll_w).reshape(B, S), (g_f - g_w).reshape(B, S, 7, -1), (np.abs(g_f) + np.abs(g_w)).reshape(B, S, 7, -1)
    return (ll_f - ll_w).reshape(B, S), (g_f - g_w).reshape(B, S, 7, -1)


def test_identity_against_reference_log_density(golden):
    """CPU: the identity reproduces the HMM term of the reference's own log_density (golden F),
    and its gradient agrees with finite differences of that composition."""
    chunks, inds = golden["model_chunks"], golden["model_inds"]
    pps = []
    for i in range(3):
        t, c, rho = orc.particle_to_dm(golden["part_x"][i], PATTERN16, 1e-2)
        pps.append(orc.params_from_dm(t, c, 1e-2, rho))
    pps = np.stack(pps)
    ll, dlog = oracle_warmup(chunks, inds, pps, 50)
    np.testing.assert_allclose(ll.sum(1), golden["model_l2"], rtol=1e-12)
    warm, body = chunks[:, :50], chunks[:, 50:]
    base = pps[1]
    total = dlog[1].sum(0)
    h = 1e-5  # central differences in log space (larger steps push d = 1 - 5e-5 past 1)
    for g, k in ((0, 3), (1, 7), (2, 5), (3, 9), (4, 11), (5, 12), (6, 2), (6, 10)):
        up, dn = base.copy(), base.copy()
        up[g, k] *= np.exp(h)
        dn[g, k] *= np.exp(-h)
        fd = (orc.hmm_term(up, warm[inds], body[inds]) - orc.hmm_term(dn, warm[inds], body[inds])) / (2 * h)
        np.testing.assert_allclose(total[g, k], fd, rtol=3e-5, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("dbl", [True, False])
def test_fused_warmup_on_gpu(golden, dbl):
    from phlash_b200.gpu import _PSMCKernelBase

    chunks, inds = golden["model_chunks"], golden["model_inds"]
    pps = golden["part_pp"][:5]
    kern = _PSMCKernelBase(16, chunks, double_precision=dbl)
    ll, dlog = kern.evaluate_warmup(pps, inds, 50, True)
    ref_ll, ref_dlog = oracle_warmup(chunks, inds, pps, 50, np.float64 if dbl else np.float32)
    if dbl:
        np.testing.assert_allclose(ll, ref_ll, rtol=1e-11)
        np.testing.assert_allclose(dlog, ref_dlog, rtol=1e-7, atol=1e-11)
        # the reference's own number for particles 0..2 of the golden set
        pps3 = []
        for i in range(3):
            t, c, rho = orc.particle_to_dm(golden["part_x"][i], PATTERN16, 1e-2)
            pps3.append(orc.params_from_dm(t, c, 1e-2, rho))
        ll3 = kern.evaluate_warmup(np.stack(pps3), inds, 50, False)
        np.testing.assert_allclose(ll3.sum(1), golden["model_l2"], rtol=1e-11)
    else:
        np.testing.assert_allclose(ll, ref_ll, rtol=1e-5)
        _, _, parts = oracle_warmup(chunks, inds, pps, 50, np.float32, with_parts=True)
        assert np.all(np.abs(dlog - ref_dlog) <= 1e-4 * np.abs(ref_dlog) + 1e-6 * parts.max(-1, keepdims=True))
    np.testing.assert_allclose(kern.evaluate_warmup(pps, inds, 50, False), ll, rtol=1e-6)
    # overlap = 0 degenerates to the plain evaluation with the stationary pi
    ll0, dlog0 = kern.evaluate_warmup(pps, inds, 0, True)
    pa = np.broadcast_to(pps[:, None], (5, len(inds), 7, 16)).copy()
    ll1, dlog1 = kern.evaluate(pa, inds, True)
    np.testing.assert_array_equal(ll0, ll1)
    np.testing.assert_array_equal(dlog0, dlog1)


@pytest.mark.gpu
def test_fused_warmup_benchmark_geometry():
    """overlap 500 + 50 000-bin chunks; per-particle sums against the oracle."""
    import torch

    from benchdata import synth
    from phlash_b200.data import _chunk_het_matrix
    from phlash_b200.distributed import pack_per_particle, unpack_per_particle
    from phlash_b200.gpu import _PSMCKernelBase

    het = synth.het_matrix(1, 260_000, seed=5)
    chunks = _chunk_het_matrix(het, 500, 50_000)[:5]
    pps = synth.particles(16, 4).astype(np.float32).astype(np.float64)
    inds = np.array([0, 3, 1, 4, 4, 2])
    kern = _PSMCKernelBase(16, chunks)
    dev = torch.device("cuda:0")
    ll, dlog = kern.evaluate_warmup_device(
        torch.tensor(pps, dtype=torch.float32, device=dev), torch.tensor(inds, device=dev), 500, True)
    kern.sync()
    ll_b, dlog_b = unpack_per_particle(pack_per_particle(ll, dlog), 16)
    ref_ll, ref_dlog, parts = oracle_warmup(chunks, inds, pps, 500, with_parts=True)
    np.testing.assert_allclose(ll_b.cpu().numpy(), ref_ll.sum(1), rtol=1e-6)
    want = ref_dlog.sum(1)
    # the result is a difference of two gradients (pi row: of two nearly equal ones, the memory of
    # the start decays over 500 warm-up bins), so fp32 error is judged against the un-cancelled size
    size = parts.sum(1)
    assert np.all(np.abs(dlog_b.cpu().numpy() - want) <= 1e-4 * np.abs(want) + 2e-6 * size.max(-1, keepdims=True))
