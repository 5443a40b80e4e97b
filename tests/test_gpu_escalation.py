"""Precision escalation of single-precision kernel objects (include/phlash_b200.h): rows holding a
long run of identical observations are marked at construction and their pairs are evaluated with
double arithmetic by a second launch; the minibatch is split on the device."""

import numpy as np
import pytest

from oracle import c_oracle, psmc_oracle as orc
from test_gpu_parity import GRAD_RTOL, LL_RTOL, grad_close, oracle_eval

pytestmark = pytest.mark.gpu

WINDOW = 1024


def expected_flags(data):
    """the rule of flag_long_runs_kernel: some ALIGNED window of 1024 sites holds a single value"""
    n, L = data.shape
    nw = L // WINDOW
    if nw == 0:
        return np.zeros(n, dtype=bool)
    w = np.clip(data[:, : nw * WINDOW], -1, 1).reshape(n, nw, WINDOW)
    return (w == w[:, :, :1]).all(axis=2).any(axis=1)


def planted_rows(L, seed=0):
    """rows with planted constant runs of several lengths, values and alignments"""
    rng = np.random.default_rng(seed)
    plants = [
        None,                      # ordinary row
        (0, 1024, -1),             # exactly one aligned window of missing data
        (1, 2046, 0),              # 2046 zeros starting at 1: covers no aligned window
        (1, 2047, 0),              # 2047 zeros starting at 1: covers [1024, 2048)
        (3000, 2500, 1),           # run of ones
        (L - 1500, 1500, -1),      # missing tail (window fully inside only if aligned)
        (L - 4096, 4096, -1),      # long missing tail
        None,
        (5000, 1023, -1),          # one short of a window
    ]
    rows = []
    for plant in plants:
        row = (rng.random(L) < 0.07).astype(np.int8)
        # make sure no accidental constant window exists
        row[:: 512] = 1
        row[1:: 512] = 0
        if plant is not None:
            start, length, value = plant
            row[start : start + length] = value
        rows.append(row)
    return np.stack(rows)


@pytest.mark.parametrize("L", [1000, 4096, 9000, 10_240])
def test_row_marking_follows_the_window_rule(L):
    from phlash_b200.gpu import _PSMCKernelBase

    data = planted_rows(max(L, 8192))[:, :L].copy()
    data[:, 0] = np.where((data > -1).any(axis=1), data[:, 0], 0)
    kern = _PSMCKernelBase(16, data)
    want = expected_flags(data)
    assert kern.num_escalated_rows == int(want.sum())
    # a double-precision object has nothing to escalate
    assert _PSMCKernelBase(16, data, double_precision=True).num_escalated_rows == 0


@pytest.mark.parametrize("M,T", [(16, 0), (16, 2), (16, 4), (32, 0), (8, 0)])
def test_mixed_minibatch_matches_oracle(M, T):
    """marked and ordinary rows in one minibatch, repeated indices, odd batch shape"""
    from phlash_b200.gpu import _PSMCKernelBase

    L = 9000
    data = planted_rows(L, seed=M)
    want = expected_flags(data)
    assert 2 <= want.sum() < len(data)
    pps, _, _ = orc.synth_particles(M, 7, seed=1)
    inds = np.array([3, 0, 6, 6, 2, 8, 1, 4, 3, 7, 5])
    pa = np.broadcast_to(pps[:, None], (7, len(inds), 7, M)).copy()
    kern = _PSMCKernelBase(M, data)
    if T:
        kern.set_threads_per_pair(T)
    assert kern.num_escalated_rows == int(want.sum())
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    for store_all in (0, 1):
        kern.set_store_all(store_all)
        ll, dlog = kern.evaluate(pa, inds, True)
        np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
        grad_close(dlog, ref_dlog, GRAD_RTOL, f"M={M} T={T} store_all={store_all}")
    # forward-only path and per-pair parameter blocks are unaffected by the split
    ll_f = kern.evaluate(pa, inds, False)
    np.testing.assert_allclose(ll_f, ref_ll, rtol=LL_RTOL)
    # only marked rows / only ordinary rows (one of the two lists is empty)
    for sel in (np.flatnonzero(want), np.flatnonzero(~want)):
        ll, dlog = kern.evaluate(pa[:, : len(sel)], sel, True)
        r_ll, r_dlog = oracle_eval(data, sel, pa[:, : len(sel)])
        np.testing.assert_allclose(ll, r_ll, rtol=LL_RTOL)
        grad_close(dlog, r_dlog, GRAD_RTOL, "single list")


def test_escalated_rows_are_scored_in_double():
    """on a marked row the result agrees with the double-precision object far beyond fp32 accuracy"""
    from phlash_b200.gpu import _PSMCKernelBase

    L = 30_000
    rng = np.random.default_rng(5)
    data = (rng.random((3, L)) < 0.07).astype(np.int8)
    data[1, 8000:28_000] = -1
    pps, _, _ = orc.synth_particles(16, 4, seed=2)
    pa = np.broadcast_to(pps[:, None], (4, 3, 7, 16)).astype(np.float32).astype(np.float64)
    k32 = _PSMCKernelBase(16, data)
    k64 = _PSMCKernelBase(16, data, double_precision=True)
    assert k32.num_escalated_rows == 1
    inds = np.arange(3)
    ll32, d32 = k32.evaluate(pa, inds, True)
    ll64, d64 = k64.evaluate(pa, inds, True)
    np.testing.assert_allclose(ll32[:, 1], ll64[:, 1], rtol=1e-12)
    np.testing.assert_allclose(d32[:, 1], d64[:, 1].astype(np.float32), rtol=2e-7, atol=1e-30)
    assert np.abs(ll32[:, 0] / ll64[:, 0] - 1).max() > 1e-12  # ordinary rows stay on the fp32 kernel
    k32.set_precision_escalation(False)
    ll_off, _ = k32.evaluate(pa, inds, True)
    assert np.abs(ll_off[:, 1] / ll64[:, 1] - 1).max() > 1e-12


def test_warmup_evaluation_with_marked_rows():
    """fused warm-up (two launches, the second subtracts) on a minibatch with marked rows"""
    from phlash_b200.gpu import _PSMCKernelBase

    ov, L = 100, 6000
    data = planted_rows(ov + L, seed=3)
    want = expected_flags(data)
    assert want.any() and not want.all()
    pps, _, _ = orc.synth_particles(16, 5, seed=4)
    pps = pps.astype(np.float32).astype(np.float64)
    inds = np.array([6, 0, 4, 6, 1])
    kern = _PSMCKernelBase(16, data)
    ll, dlog = kern.evaluate_warmup(pps, inds, ov, True)
    for b in range(5):
        for j, r in enumerate(inds):
            full_ll, full_g = c_oracle.loglik_batch(data, np.array([r]), pps[b : b + 1], grad=True)
            warm_ll, warm_g = c_oracle.loglik_batch(data[:, :ov].copy(), np.array([r]), pps[b : b + 1], grad=True)
            np.testing.assert_allclose(ll[b, j], full_ll[0] - warm_ll[0], rtol=LL_RTOL)
            want_g = (full_g - warm_g).reshape(7, 16)
            scale = np.abs(full_g).reshape(7, 16).max(axis=-1, keepdims=True)
            assert (np.abs(dlog[b, j] - want_g) <= GRAD_RTOL * (np.abs(want_g) + 1e-3 * scale) + 2e-7 * scale).all()


def test_bad_index_is_still_reported():
    """device entry: the split sends out-of-range rows to the ordinary kernel, which flags them"""
    import torch
    from phlash_b200.gpu import _PSMCKernelBase

    data = planted_rows(5000)
    kern = _PSMCKernelBase(16, data)
    assert kern.num_escalated_rows > 0
    pps, _, _ = orc.synth_particles(16, 2, seed=0)
    dev = torch.device("cuda:0")
    p6 = torch.tensor(pps[:, :6], dtype=torch.float32, device=dev).contiguous()
    pi = torch.tensor(pps[:, 6], dtype=torch.float32, device=dev).contiguous()
    flagged = int(np.flatnonzero(expected_flags(data))[0])
    inds = torch.tensor([0, len(data), flagged, -1], device=dev)
    ll, _ = kern.evaluate_device(p6, pi, inds, True)
    with pytest.raises(AssertionError):
        kern.sync()
    ll = ll.cpu().numpy()
    assert np.isnan(ll[:, [1, 3]]).all() and np.isfinite(ll[:, [0, 2]]).all()


def test_small_minibatch_keeps_parallel_in_time_with_marked_rows_in_the_data():
    """The split is decided per call: real data always holds marked rows (the padded last chunk of every
    contig), and a minibatch of ONE or FIVE chunks must still take the parallel-in-time gradient paths -
    with marked chunks in the minibatch scored in double by the second launch."""
    from phlash_b200.gpu import _PSMCKernelBase

    het = orc.synth_het_matrix(1, 420_000, seed=5)
    data = orc.chunk_het_matrix(het, 500, 50_000)[:, 500:].copy()  # 9 chunks, the last one padded
    pps, _, _ = orc.synth_particles(16, 40, seed=3)
    kern = _PSMCKernelBase(16, data)
    assert kern.num_escalated_rows == 1
    for inds in (np.array([2]), np.array([8]), np.array([1, 8, 3, 8, 0])):
        pa = np.broadcast_to(pps[:, None], (len(pps), len(inds), 7, 16)).copy()
        ll, dlog = kern.evaluate(pa, inds, True)
        assert "segments" in kern.last_kernel_name, kern.last_kernel_name
        ref_ll, ref_dlog = oracle_eval(data, inds, pa)
        np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
        grad_close(dlog, ref_dlog, GRAD_RTOL, f"inds={inds}")
    # fused warm-up form (second launch subtracts): marked pairs must not be subtracted twice
    full = orc.chunk_het_matrix(het, 500, 50_000)
    kw = _PSMCKernelBase(16, full)
    inds = np.array([8, 4])
    p7 = pps[:12].astype(np.float32).astype(np.float64)
    ll, dlog = kw.evaluate_warmup(p7, inds, 500, True)
    assert "segments" in kw.last_kernel_name or "storeall" in kw.last_kernel_name
    from test_warmup import oracle_warmup

    want_ll, want_dlog, parts = oracle_warmup(full, inds, p7, 500, np.float32, with_parts=True)
    np.testing.assert_allclose(ll, want_ll, rtol=LL_RTOL)
    # (the difference of two evaluations: tolerance relative to the terms that were subtracted)
    scale = np.abs(parts).max(-1, keepdims=True)
    assert np.all(np.abs(dlog - want_dlog) <= GRAD_RTOL * parts + GRAD_RTOL * 1e-3 * scale)
