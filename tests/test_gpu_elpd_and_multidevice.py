"""(i) The ELPD path of the reference (mcmc.py:213-238): forward-only log-likelihood of whole,
un-chunked test contigs for every particle, through `log_density` with a single missing warm-up
bin.  (ii) In-process multi-device evaluation, PSMCKernel(num_gpus=2) (gpu.py:386-438)."""

import numpy as np
import pytest

from oracle import c_oracle, psmc_oracle as orc

pytestmark = pytest.mark.gpu


def test_elpd_path_long_unchunked_contig():
    from benchdata import synth
    from phlash_b200.gpu import PSMCKernel
    from phlash_b200.params import PSMCParams

    het = synth.het_matrix(2, 300_001, seed=9)  # two test diploids, odd length
    pps = synth.particles(16, 6).astype(np.float32).astype(np.float64)
    kern = PSMCKernel(M=16, data=het, double_precision=False, num_gpus=1)
    inds = np.arange(2)
    pa = np.broadcast_to(pps[:, None], (6, 2, 7, 16)).copy()
    ll = kern(PSMCParams.from_block(pa), inds, grad=False)
    ref = c_oracle.loglik_batch(het, np.tile(inds, 6), pa.reshape(12, 7, 16)).reshape(6, 2)
    np.testing.assert_allclose(ll, ref, rtol=1e-5)
    # the reference's elpd() prepends ONE missing warm-up bin (mcmc.py:229-233): pi -> pi A
    full = np.concatenate([np.full((2, 1), -1, dtype=np.int8), het], axis=1)
    kern_w = PSMCKernel(M=16, data=full, double_precision=False, num_gpus=1)
    ll_w = kern_w.gpu_kernels[0].evaluate_warmup(pps, inds, 1, False)
    want = np.array([[orc.hmm_term(p, full[i : i + 1, :1], full[i : i + 1, 1:]) for i in range(2)] for p in pps[:2]])
    np.testing.assert_allclose(ll_w[:2], want, rtol=1e-5)
    # ... and the same through the one-call entry from the particles themselves
    import torch

    from phlash_b200 import model

    pps_x, xs, pattern = orc.synth_particles(16, 4, seed=5)
    x = torch.tensor(xs, device="cuda:0")
    tk = model.elpd_kernel(16, het)
    np.testing.assert_array_equal(tk.download_data(), full)
    got = float(model.elpd_hmm_term(tk, x, pattern, 1e-2))
    # (theta of the synthetic particles is 1e-2, orc.synth_particles)
    want_e = np.mean([sum(orc.hmm_term(p.astype(np.float32).astype(np.float64), full[i : i + 1, :1], full[i : i + 1, 1:])
                          for i in range(2)) for p in pps_x])
    np.testing.assert_allclose(got, want_e, rtol=1e-5)


def test_two_devices_in_one_process(golden):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import fixture_data
    from phlash_b200.gpu import PSMCKernel
    from phlash_b200.params import PSMCParams

    data, _ = fixture_data(0)
    pps = golden["part_pp"][:4]
    inds = np.array([9, 0, 3, 3, 7])
    pa = np.broadcast_to(pps[:, None], (4, 5, 7, 16)).copy()
    one = PSMCKernel(16, data, num_gpus=1)(PSMCParams.from_block(pa), inds, grad=True)
    two = PSMCKernel(16, data, num_gpus=2)
    assert len(two.gpu_kernels) == 2
    ll2, dll2 = two(PSMCParams.from_block(pa), inds, grad=True)
    np.testing.assert_array_equal(ll2, one[0])
    np.testing.assert_array_equal(dll2.to_block(), one[1].to_block())
    # a single chunk (fewer chunks than devices) must not dead-lock (reference FIXME, gpu.py:404)
    ll1 = two(PSMCParams.from_block(pa[:, :1]), inds[:1], grad=False)
    np.testing.assert_array_equal(ll1, one[0][:, :1])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_chunking_on_the_device_is_bit_exact(golden, tag):
    """phb_create_from_contig cuts the windows of _chunk_het_matrix (data.py:37-61) on the device;
    compared with the reference function's own output (golden G)."""
    from phlash_b200.gpu import _PSMCKernelBase

    ov, cs = (int(v) for v in golden[f"chunk_{tag}_geom"])
    het = golden[f"chunk_{tag}_in"]
    want = golden[f"chunk_{tag}_out"]
    if not np.all(want[:, ov:].max(axis=1) > -1):  # the reference checks the data part (mcmc.py:203, gpu.py:111-113)
        with pytest.raises(AssertionError):
            _PSMCKernelBase.from_contig(16, het, ov, cs)
        return
    kern = _PSMCKernelBase.from_contig(16, het, ov, cs)
    np.testing.assert_array_equal(kern.download_data(), want)


def test_contig_to_likelihood_in_one_object(golden):
    """binned contig -> device chunking -> fused warm-up evaluation, against the host-chunked path."""
    from phlash_b200.gpu import _PSMCKernelBase

    het, inds = golden["model_het"], golden["model_inds"]
    pps = golden["part_pp"][:3]
    k_dev = _PSMCKernelBase.from_contig(16, het, 50, 500, double_precision=True)
    k_host = _PSMCKernelBase(16, golden["model_chunks"], double_precision=True)
    np.testing.assert_array_equal(k_dev.download_data(), golden["model_chunks"])
    a = k_dev.evaluate_warmup(pps, inds, 50, True)
    b = k_host.evaluate_warmup(pps, inds, 50, True)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
