"""Everything around the kernels that an SVGD iteration touches (SURVEY.md section 8f-3 and the constructor):
device-side validation of the observation matrix, scratch reservation (no allocation in the step), minibatch
sampling on the device, CUDA-graph capture of a whole likelihood step."""

import ctypes

import numpy as np
import pytest

from oracle import psmc_oracle as orc

pytestmark = pytest.mark.gpu


def _chunks(n_bins=260_000, seed=2):
    het = orc.synth_het_matrix(1, n_bins, seed=seed)
    return orc.chunk_het_matrix(het, 500, 50_000)


def test_constructor_checks_run_on_the_device():
    """gpu.py:106-113: values >= -1, clipping to <= 1, every row holds an observation - here evaluated by
    fixup_rows_kernel after the upload."""
    from phlash_b200 import _native
    from phlash_b200.gpu import _PSMCKernelBase

    lib = _native.lib()
    rng = np.random.default_rng(0)
    data = (rng.random((7, 1003)) < 0.1).astype(np.int8)
    data[3, 5] = 4            # clipped to 1
    data[6, 1002] = 3
    kern = _PSMCKernelBase(16, data)
    np.testing.assert_array_equal(kern.download_data(), np.minimum(data, 1))
    h = ctypes.c_void_p()
    bad = data.copy()
    bad[5, 1001] = -2
    bad[6, 17] = -3
    rc = lib.phb_create(16, bad.ctypes.data, 7, 1003, 0, 0, ctypes.byref(h))
    assert rc == _native.PHB_E_DATA and "data[5, 1001] < -1" in _native.last_error(), _native.last_error()
    allmiss = data.copy()
    allmiss[4] = -1
    allmiss[2] = -1
    rc = lib.phb_create(16, allmiss.ctypes.data, 7, 1003, 0, 0, ctypes.byref(h))
    assert rc == _native.PHB_E_DATA and "all missing values (row 2)" in _native.last_error(), _native.last_error()
    assert not h.value
    with pytest.raises(AssertionError):
        _PSMCKernelBase(16, allmiss)
    # full chunks: a row whose DATA part is all missing is rejected even if its warm-up bins are observed
    # (the reference splits first, mcmc.py:203), and accepted the other way round
    full = data.copy()
    full[1, 100:] = -1
    assert (full[1, :100] > -1).any()
    with pytest.raises(AssertionError):
        _PSMCKernelBase(16, full, overlap=100)
    _PSMCKernelBase(16, full, overlap=0)
    ok = data.copy()
    ok[1, :100] = -1
    _PSMCKernelBase(16, ok, overlap=100)


def test_upload_of_a_multi_slab_matrix_is_bit_exact():
    """more than one 128 MB staging slab, row length not a multiple of 16 (pitch > L)"""
    from phlash_b200.gpu import _PSMCKernelBase

    rng = np.random.default_rng(1)
    data = rng.integers(-1, 3, size=(6000, 50_500), dtype=np.int8)  # ~300 MB, values -1..2
    kern = _PSMCKernelBase(16, data)
    np.testing.assert_array_equal(kern.download_data(), np.minimum(data, 1))


def test_sampling_on_the_device_matches_the_host_generator():
    import torch

    from phlash_b200.gpu import _PSMCKernelBase, minibatch_indices

    chunks = _chunks()
    kern = _PSMCKernelBase(16, chunks, overlap=500)
    n = chunks.shape[0]
    kern.set_iteration(11)
    a = kern.sample_minibatch(seed=5, S=7)
    b = kern.sample_minibatch(seed=5, S=7)  # the counter advanced
    torch.cuda.synchronize()
    np.testing.assert_array_equal(a.cpu().numpy(), minibatch_indices(5, 11, n, 7))
    np.testing.assert_array_equal(b.cpu().numpy(), minibatch_indices(5, 12, n, 7))


@pytest.mark.parametrize("S", [1, 5])
def test_whole_step_is_allocation_free_and_graph_capturable(S):
    """phb_reserve, then: sample a minibatch on the device + the whole HMM term (particles -> parameters ->
    fused warm-up loglik + gradient -> VJP) recorded ONCE into a CUDA graph; every replay draws the next
    minibatch and reproduces the eager evaluation of the same indices bit for bit, without any device
    allocation (mcmc.py:275-279 is one host round trip per iteration in the reference)."""
    import torch

    from phlash_b200.gpu import _PSMCKernelBase, minibatch_indices

    chunks = _chunks()
    n = chunks.shape[0]
    _, xs, pattern = orc.synth_particles(16, 64, seed=4)
    dev = torch.device("cuda:0")
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    kern = _PSMCKernelBase(16, chunks, overlap=500)
    kern.reserve(B=64, S_max=8, overlap=500)
    allocs = kern.allocation_count
    inds = torch.empty(S, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream()
    # warm-up on the capture stream (cuda graphs want the kernels loaded)
    with torch.cuda.stream(side):
        kern.set_iteration(0)
        kern.sample_minibatch(3, S, out=inds)
        kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=n / S)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    kern.set_iteration(100)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=side):
        kern.sample_minibatch(3, S, out=inds)
        value, grad = kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=n / S)
    for it in (100, 101, 102):
        graph.replay()
        torch.cuda.synchronize()
        want_inds = minibatch_indices(3, it, n, S)
        np.testing.assert_array_equal(inds.cpu().numpy(), want_inds)
        v_e, g_e = kern.hmm_term(x, pattern, 1e-2, torch.tensor(want_inds, device=dev), 500, weight=n / S)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(value.cpu().numpy(), v_e.cpu().numpy())
        np.testing.assert_array_equal(grad.cpu().numpy(), g_e.cpu().numpy())
    assert kern.allocation_count == allocs, "the step allocated device memory after phb_reserve"
    kern.sync()


def test_fp32_peak_measurement():
    from phlash_b200.gpu import measure_fp32_peak

    ind, acc = measure_fp32_peak(0)
    assert 40.0 < ind < 80.0 and 25.0 < acc <= ind * 1.02, (ind, acc)


def test_fit_loop_follows_the_reference_schedule():
    """phlash_b200.mcmc.fit_loop: S rule, device-side sampling (reproducible from the seed), N / S weight, ELPD
    every 10th iteration with the early stop (mcmc.py:116-140, 275-304); one CUDA-graph replay per iteration
    gives exactly the particles of the eager loop."""
    import os
    import sys

    import torch

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import svgd_torch

    from phlash_b200 import mcmc, model
    from phlash_b200.gpu import _PSMCKernelBase, minibatch_indices

    chunks = _chunks(310_000, seed=7)[:6]  # 6 full chunks
    _, xs, pattern = orc.synth_particles(16, 48, seed=2)
    test_het = orc.synth_het_matrix(1, 20_000, seed=9)
    assert model.default_minibatch_size(len(chunks), 1000) == 1 and model.default_minibatch_size(6000, 1000) == 5
    out = []
    for use_graph in (True, False):
        x0 = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
        opt = svgd_torch.SvgdAmsgrad(x0)
        res = mcmc.fit_loop(chunks, x0, pattern, 1e-2, opt, niter=24, overlap=500, minibatch_size=2, test_het=test_het,
                            log_prior_grad=svgd_torch.log_prior_grad, seed=11, use_graph=use_graph)
        assert res.iterations == 24 and res.minibatch_size == 2 and res.n_chunks == 6 and not res.stopped_early
        assert len(res.elpd_trace) == 3  # iterations 0, 10, 20 (mcmc.py:287)
        assert res.graph_replays == (24 - 3 if use_graph else 0)
        out.append(res.particles.cpu().numpy())
    np.testing.assert_array_equal(out[0], out[1])
    assert np.abs(out[0] - xs).max() > 1e-3  # the particles moved
    # the first iteration, by hand: the sampler's indices, the weighted HMM term, the prior, one update
    x = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
    kern = _PSMCKernelBase(16, chunks, overlap=500)
    inds = torch.tensor(minibatch_indices(11, 0, 6, 2), device="cuda:0")
    _, g = kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=6 / 2)
    opt = svgd_torch.SvgdAmsgrad(x)
    x1 = x.clone()
    opt(x1, g + svgd_torch.log_prior_grad(x))
    x0 = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
    res = mcmc.fit_loop(chunks, x0, pattern, 1e-2, svgd_torch.SvgdAmsgrad(x0), niter=1, overlap=500, minibatch_size=2,
                        log_prior_grad=svgd_torch.log_prior_grad, seed=11, use_graph=False)
    np.testing.assert_array_equal(res.particles.cpu().numpy(), x1.cpu().numpy())
    # early stop: an update that makes the particles worse every time trips the ELPD cut-off
    x0 = torch.tensor(xs, dtype=torch.float64, device="cuda:0")

    def worsen(x, score):
        x[:, 2:-1] += 0.5  # population sizes drift away

    res = mcmc.fit_loop(chunks, x0, pattern, 1e-2, worsen, niter=200, overlap=500, minibatch_size=1, test_het=test_het,
                        elpd_cutoff=15, seed=1, use_graph=False)
    assert res.stopped_early and res.iterations < 60
