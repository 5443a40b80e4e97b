"""Parallel-in-time forward evaluation (segment transfer operators chained in float64; the path of
the reference's ELPD, mcmc.py:213-238) against the fp64 oracle and the sequential kernel."""

import numpy as np
import pytest

from oracle import c_oracle, psmc_oracle as orc
from test_gpu_parity import LL_RTOL

pytestmark = pytest.mark.gpu


def rows_with_missing(n, L, seed):
    rng = np.random.default_rng(seed)
    data = (rng.random((n, L)) < 0.07).astype(np.int8)
    data[rng.random((n, L)) < 0.02] = -1
    data[:, 0] = np.where(data[:, 0] < 0, 0, data[:, 0])
    data[0, 10:90] = -1  # a run of missing observations across a block boundary
    data[-1, L - 37 :] = -1  # missing tail
    return data


@pytest.mark.parametrize("M", [4, 8, 16])
@pytest.mark.parametrize("L", [200, 1003, 20_000])
def test_transfer_operators_match_oracle_and_sequential_kernel(M, L):
    from phlash_b200.gpu import _PSMCKernelBase

    data = rows_with_missing(3, L, seed=M + L)
    pps, _, _ = orc.synth_particles(M, 5, seed=3)
    pps = pps.astype(np.float32).astype(np.float64)
    inds = np.array([2, 0, 1, 2])
    pa = np.broadcast_to(pps[:, None], (5, len(inds), 7, M)).copy()
    want = c_oracle.loglik_batch(data, np.tile(inds, 5), pa.reshape(-1, 7, M)).reshape(5, len(inds))
    kern = _PSMCKernelBase(M, data)
    kern.set_parallel_in_time(1)
    ll_pit = kern.evaluate(pa, inds, False)
    assert "transfer" in kern.last_kernel_name
    kern.set_parallel_in_time(0)
    ll_seq = kern.evaluate(pa, inds, False)
    assert "transfer" not in kern.last_kernel_name
    np.testing.assert_allclose(ll_pit, want, rtol=LL_RTOL)
    np.testing.assert_allclose(ll_pit, ll_seq, rtol=2e-6)


def test_dispatch_rules():
    from phlash_b200.gpu import _PSMCKernelBase

    data = rows_with_missing(2, 40_000, seed=1)
    pps, _, _ = orc.synth_particles(16, 4, seed=0)
    kern = _PSMCKernelBase(16, data)
    pa = np.broadcast_to(pps[:, None], (4, 2, 7, 16)).copy()
    kern.evaluate(pa, np.arange(2), False)
    assert "transfer" in kern.last_kernel_name            # 8 pairs x 40 000 sites
    kern.evaluate(pa, np.arange(2), True)
    assert "segments" in kern.last_kernel_name            # gradient path: also parallel in time
    kern.set_threads_per_pair(2)
    kern.evaluate(pa, np.arange(2), False)
    assert "transfer" not in kern.last_kernel_name        # a forced lane layout wins
    kern.set_threads_per_pair(0)
    big = np.broadcast_to(pps[:1, None], (1, 3000, 7, 16)).copy()
    kern.evaluate(big, np.zeros(3000, dtype=np.int64), False)
    assert "transfer" not in kern.last_kernel_name        # 3000 pairs: the sequential kernel is as fast
    short = _PSMCKernelBase(16, data[:, :5000].copy())
    short.evaluate(pa, np.arange(2), False)
    assert "transfer" not in short.last_kernel_name       # segments would be shorter than 4096 sites
    dbl = _PSMCKernelBase(16, data, double_precision=True)
    dbl.evaluate(pa, np.arange(2), False)
    assert "transfer" not in dbl.last_kernel_name


def test_elpd_shape_through_the_one_call_entry():
    """B particles x one un-chunked contig behind one missing bin (mcmc.py:229-233): the evaluation of
    all bins goes through the transfer operators, the subtraction of the warm-up bin through the
    sequential kernel; compared with the same call with the parallel-in-time path switched off."""
    import torch

    from benchdata import synth
    from phlash_b200 import model

    het = synth.het_matrix(1, 400_003, seed=4)
    _, xs, pattern = orc.synth_particles(16, 8, seed=2)
    x = torch.tensor(xs, device="cuda:0")
    tk = model.elpd_kernel(16, het)
    inds = torch.arange(1, device="cuda:0")
    v_pit, _ = tk.hmm_term(x, pattern, 1e-2, inds, 1, 1.0, grad=False)
    tk.set_parallel_in_time(0)
    v_seq, _ = tk.hmm_term(x, pattern, 1e-2, inds, 1, 1.0, grad=False)
    torch.testing.assert_close(v_pit, v_seq, rtol=2e-6, atol=0)
    pps = orc.synth_particles(16, 8, seed=2)[0].astype(np.float32).astype(np.float64)
    full = tk.download_data()
    want = np.array([orc.hmm_term(p, full[:, :1], full[:, 1:]) for p in pps[:3]])
    np.testing.assert_allclose(v_pit.cpu().numpy()[:3], want, rtol=LL_RTOL)


def test_bad_index_is_reported():
    import torch

    from phlash_b200.gpu import _PSMCKernelBase

    data = rows_with_missing(2, 40_000, seed=1)
    pps, _, _ = orc.synth_particles(16, 2, seed=0)
    kern = _PSMCKernelBase(16, data)
    dev = torch.device("cuda:0")
    p6 = torch.tensor(pps[:, :6], dtype=torch.float32, device=dev).contiguous()
    pi = torch.tensor(pps[:, 6], dtype=torch.float32, device=dev).contiguous()
    ll, _ = kern.evaluate_device(p6, pi, torch.tensor([0, 2, 1], device=dev), False)
    assert "transfer" in kern.last_kernel_name
    with pytest.raises(AssertionError):
        kern.sync()
    ll = ll.cpu().numpy()
    assert np.isnan(ll[:, 1]).all() and np.isfinite(ll[:, [0, 2]]).all()


@pytest.mark.parametrize("M", [4, 8, 16])
@pytest.mark.parametrize("L", [200, 3001, 20_000])
def test_gradient_through_segments_matches_oracle(M, L):
    """Parallel-in-time gradient: boundary vectors from the chained operators, store-all passes per
    segment, partial gradients added up - against the fp64 oracle and the sequential kernels."""
    from test_gpu_parity import GRAD_RTOL, grad_close, oracle_eval

    from phlash_b200.gpu import _PSMCKernelBase

    data = rows_with_missing(3, L, seed=2 * M + L)
    pps, _, _ = orc.synth_particles(M, 5, seed=7)
    inds = np.array([2, 0, 1, 2])
    pa = np.broadcast_to(pps[:, None], (5, len(inds), 7, M)).copy()
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    kern = _PSMCKernelBase(M, data)
    kern.set_parallel_in_time(1)
    ll, dlog = kern.evaluate(pa, inds, True)
    assert "segments" in kern.last_kernel_name
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dlog, ref_dlog, GRAD_RTOL, f"parallel in time, M={M}, L={L}")
    kern.set_parallel_in_time(0)
    ll_s, dlog_s = kern.evaluate(pa, inds, True)
    assert "segments" not in kern.last_kernel_name
    np.testing.assert_allclose(ll, ll_s, rtol=2e-6)
    grad_close(dlog, dlog_s, 2e-5, "vs sequential kernel")
    # structural zeros stay exactly zero
    assert np.all(dlog[:, :, 0, -1] == 0) and np.all(dlog[:, :, 2, -1] == 0) and np.all(dlog[:, :, 3, 0] == 0)


def test_gradient_dispatch_and_warmup_composition(golden):
    """automatic choice at the reference's default shape for one genome (S = 1: few pairs, long chunks),
    and the fused warm-up evaluation (second launch subtracts) on top of it"""
    from test_gpu_parity import GRAD_RTOL

    from phlash_b200.gpu import _PSMCKernelBase

    ov, L = 100, 12_000
    data = rows_with_missing(4, ov + L, seed=11)
    pps, _, _ = orc.synth_particles(16, 6, seed=4)
    pps = pps.astype(np.float32).astype(np.float64)
    kern = _PSMCKernelBase(16, data)
    inds = np.array([3])
    ll, dlog = kern.evaluate_warmup(pps, inds, ov, True)
    assert "warm-up term" in kern.last_kernel_name          # scored as one more segment, no second launch
    pa = np.broadcast_to(pps[:, None], (6, 1, 7, 16)).copy()
    kern.evaluate(pa, inds, True)
    assert "segments" in kern.last_kernel_name              # 6 pairs x 12 100 sites
    kern.evaluate(np.broadcast_to(pps[:1, None], (1, 4000, 7, 16)).copy(), np.zeros(4000, dtype=np.int64), True)
    assert "transfer_rows" not in kern.last_kernel_name     # 4000 pairs: not worth M x the forward work
    kern.set_parallel_in_time(0)
    ll_s, dlog_s = kern.evaluate_warmup(pps, inds, ov, True)
    np.testing.assert_allclose(ll, ll_s, rtol=1e-6)
    # the result is a DIFFERENCE of two gradients (all bins - warm-up bins): fp32 round-off is relative to
    # the un-cancelled size
    _, full = kern.evaluate(pa, inds, True)
    scale = np.abs(dlog_s).max(axis=-1, keepdims=True)
    scale_full = np.abs(full).max(axis=-1, keepdims=True)
    assert (np.abs(dlog - dlog_s) <= GRAD_RTOL * (np.abs(dlog_s) + 1e-3 * scale) + 2e-6 * scale_full).all()


@pytest.mark.parametrize("B,S,path", [(6, 1, "transfer_rows"), (300, 3, "boundary_sweep")])
def test_fused_warmup_term_matches_oracle(B, S, path, monkeypatch):
    """LL(whole row) - LL(first `ov` sites) and its gradient when the warm-up term rides along with the segment
    passes (operator path and two-sweep path), against the oracle's difference and against the two-launch form."""
    from test_gpu_parity import GRAD_RTOL, oracle_eval

    from phlash_b200.gpu import _PSMCKernelBase

    ov, L = 132, 8_000  # (ov is not a multiple of the checkpoint spacing)
    data = rows_with_missing(4, ov + L, seed=31 + S)
    pps, _, _ = orc.synth_particles(16, B, seed=5)
    pps = pps.astype(np.float32).astype(np.float64)
    inds = np.arange(S) + 1
    pa = np.broadcast_to(pps[:, None], (B, S, 7, 16)).copy()
    full_ll, full_dlog = oracle_eval(data, inds, pa)
    warm_ll, warm_dlog = oracle_eval(np.ascontiguousarray(data[:, :ov]), inds, pa)
    kern = _PSMCKernelBase(16, data)
    ll, dlog = kern.evaluate_warmup(pps, inds, ov, True)
    assert path in kern.last_kernel_name and "warm-up term" in kern.last_kernel_name, kern.last_kernel_name
    np.testing.assert_allclose(ll, full_ll - warm_ll, rtol=2e-6)
    # a DIFFERENCE of two gradients: fp32 round-off is relative to the un-cancelled size
    want = full_dlog - warm_dlog
    scale = np.abs(want).max(axis=-1, keepdims=True)
    scale_full = np.abs(full_dlog).max(axis=-1, keepdims=True)
    assert (np.abs(dlog - want) <= GRAD_RTOL * (np.abs(want) + 1e-3 * scale) + 2e-6 * scale_full).all()
    monkeypatch.setenv("PHB_FUSE_WARMUP", "0")
    two = _PSMCKernelBase(16, data)
    ll2, dlog2 = two.evaluate_warmup(pps, inds, ov, True)
    assert "warm-up term" not in two.last_kernel_name
    np.testing.assert_allclose(ll, ll2, rtol=1e-6)
    assert (np.abs(dlog - dlog2) <= 2e-5 * (np.abs(dlog2) + 1e-3 * scale) + 2e-6 * scale_full).all()


def test_per_pair_parameter_blocks_and_subtracting_launch():
    """(i) parameter rows that differ between the chunks of a particle (pa [B, S, 7, M] not shared);
    (ii) a warm-up long enough for the SECOND launch of the fused evaluation (which subtracts) to be
    parallel in time as well."""
    from test_gpu_parity import GRAD_RTOL, grad_close, oracle_eval

    from phlash_b200.gpu import _PSMCKernelBase

    L = 9000
    data = rows_with_missing(3, L, seed=21)
    pps, _, _ = orc.synth_particles(16, 6, seed=9)
    inds = np.array([1, 2])
    pa = np.stack([pps[:3], pps[3:6]], axis=1)  # [3, 2, 7, 16]: chunk s of particle b has its own block
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    kern = _PSMCKernelBase(16, data)
    kern.set_parallel_in_time(1)
    ll, dlog = kern.evaluate(pa, inds, True)
    assert "segments" in kern.last_kernel_name
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dlog, ref_dlog, GRAD_RTOL, "per-pair parameter blocks")
    ll_f = kern.evaluate(pa, inds, False)
    assert "transfer" in kern.last_kernel_name
    np.testing.assert_allclose(ll_f, ref_ll, rtol=LL_RTOL)
    # (ii)
    ov = 4000
    p7 = pps[:4].astype(np.float32).astype(np.float64)
    got_ll, got_dlog = kern.evaluate_warmup(p7, inds, ov, True)
    assert "segments" in kern.last_kernel_name           # the launch over the first `ov` bins
    kern.set_parallel_in_time(0)
    seq_ll, seq_dlog = kern.evaluate_warmup(p7, inds, ov, True)
    np.testing.assert_allclose(got_ll, seq_ll, rtol=1e-6)
    scale = np.abs(seq_dlog).max(axis=-1, keepdims=True)
    assert (np.abs(got_dlog - seq_dlog) <= GRAD_RTOL * (np.abs(seq_dlog) + 1e-3 * scale) + 2e-6 * scale).all()


@pytest.mark.parametrize("M", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("L", [150, 3001, 20_000])
def test_two_sweep_gradient_matches_oracle(M, L):
    """Boundary vectors from the forward / adjoint-only sweeps (mode 2), segment passes, summed partial
    gradients - against the fp64 oracle and the sequential kernels; any M."""
    from test_gpu_parity import GRAD_RTOL, grad_close, oracle_eval

    from phlash_b200.gpu import _PSMCKernelBase

    data = rows_with_missing(3, L, seed=3 * M + L)
    pps, _, _ = orc.synth_particles(M, 5, seed=8)
    inds = np.array([2, 0, 1, 2])
    pa = np.broadcast_to(pps[:, None], (5, len(inds), 7, M)).copy()
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    kern = _PSMCKernelBase(M, data)
    kern.set_parallel_in_time(2)
    ll, dlog = kern.evaluate(pa, inds, True)
    assert "boundary_sweep" in kern.last_kernel_name
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dlog, ref_dlog, GRAD_RTOL, f"two sweeps, M={M}, L={L}")
    kern.set_parallel_in_time(0)
    ll_s, dlog_s = kern.evaluate(pa, inds, True)
    np.testing.assert_allclose(ll, ll_s, rtol=2e-6)
    grad_close(dlog, dlog_s, 2e-5, "vs sequential kernel")
    assert np.all(dlog[:, :, 0, -1] == 0) and np.all(dlog[:, :, 2, -1] == 0) and np.all(dlog[:, :, 3, 0] == 0)


@pytest.mark.parametrize("M,TF,TB,LL", [(8, 2, 2, 1), (8, 2, 2, 0), (16, 4, 4, 1), (16, 4, 2, 1), (16, 2, 2, 1),
                                        (16, 4, 4, 0), (16, 1, 1, 0), (32, 8, 8, 1), (32, 8, 4, 1), (32, 4, 4, 1), (32, 4, 2, 1),
                                        (32, 2, 2, 1), (32, 8, 8, 0)])
def test_every_sweep_lane_layout(M, TF, TB, LL, monkeypatch):
    """The sweeps choose their lane layout (lanes per pair, separately for the forward and the adjoint
    direction, generic or low-latency site functions) by the number of pairs; every layout is forced here
    through the experiment knobs, which are read when the kernel object is created."""
    from test_gpu_parity import GRAD_RTOL, grad_close, oracle_eval

    from phlash_b200.gpu import _PSMCKernelBase

    L = 2_531  # ragged last block, ragged last segment
    data = rows_with_missing(3, L, seed=M + TF + 7 * TB)
    pps, _, _ = orc.synth_particles(M, 37, seed=9)  # pairs do not fill the last warp of either direction
    inds = np.array([1, 2, 0])
    pa = np.broadcast_to(pps[:, None], (37, len(inds), 7, M)).copy()
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    monkeypatch.setenv("PHB_SWEEP_TF", str(TF))
    monkeypatch.setenv("PHB_SWEEP_TB", str(TB))
    monkeypatch.setenv("PHB_SWEEP_LL", str(LL))
    kern = _PSMCKernelBase(M, data)
    kern.set_parallel_in_time(2)
    ll, dlog = kern.evaluate(pa, inds, True)
    want = f"boundary_sweep_kernel<float,TF={TF},TB={TB}" + (",LL>" if LL else ">")
    assert want in kern.last_kernel_name, kern.last_kernel_name
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dlog, ref_dlog, GRAD_RTOL, f"sweep layout TF={TF} TB={TB} LL={LL}, M={M}")


def test_sweep_layout_gives_every_warp_a_scheduler():
    """2 500 pairs (the reference's S = 5 with 500 particles) at M = 16: four lanes per pair in both
    directions would need 158 four-warp CTAs for 148 SMs; the picker narrows the adjoint sweep instead."""
    from phlash_b200.gpu import _PSMCKernelBase

    data = rows_with_missing(4, 12_000, seed=5)
    pps, _, _ = orc.synth_particles(16, 1, seed=4)
    kern = _PSMCKernelBase(16, data)
    for n_chunks, expect in ((1500, "TF=4,TB=4,LL"), (2500, "TF=4,TB=2,LL"), (4000, "TF=2,TB=2,LL"), (6000, "TF=1,TB=1>")):
        pa = np.broadcast_to(pps[:1, None], (1, n_chunks, 7, 16)).copy()
        kern.evaluate(pa, np.zeros(n_chunks, dtype=np.int64), True)
        assert expect in kern.last_kernel_name, (n_chunks, kern.last_kernel_name)


def test_two_sweep_is_chosen_between_operators_and_store_all():
    from phlash_b200.gpu import _PSMCKernelBase

    data = rows_with_missing(4, 12_000, seed=5)
    pps, _, _ = orc.synth_particles(16, 1, seed=4)
    kern = _PSMCKernelBase(16, data)
    for n_chunks, expect in ((1, "transfer_rows"), (2500, "boundary_sweep"), (40_000, "psmc_loglik_kernel")):
        pa = np.broadcast_to(pps[:1, None], (1, n_chunks, 7, 16)).copy()
        kern.evaluate(pa, np.zeros(n_chunks, dtype=np.int64), True)
        assert expect in kern.last_kernel_name, (n_chunks, kern.last_kernel_name)


@pytest.mark.parametrize("S,world", [(1, 2), (1, 8), (3, 4), (5, 8)])
def test_time_axis_sharding_equals_the_single_process_term(S, world):
    """phb_hmm_term_sharded_*: the segments of the parallel-in-time gradient sharded over `world` processes.
    Emulated on ONE GPU: every "process" fills its slot of the gather buffer (what the all-gather would
    deliver), then every one of them finishes its share; the partial per-particle sums add up to the sums of
    the single-process call (different segmentation: agreement to fp32 round-off, not bit for bit).  The
    data hold a marked row (padded last chunk), which process 0 scores in double."""
    import torch

    from phlash_b200.gpu import _PSMCKernelBase

    het = orc.synth_het_matrix(1, 420_000, seed=11)
    chunks = orc.chunk_het_matrix(het, 500, 50_000)
    _, xs, pattern = orc.synth_particles(16, 40, seed=6)
    dev = torch.device("cuda:0")
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    kern = _PSMCKernelBase(16, chunks, overlap=500)
    assert kern.num_escalated_rows == 1
    inds = torch.tensor([8, 2, 5, 0, 7][:S], device=dev)  # chunk 8 is the marked one
    want = kern.hmm_term_sums(x, pattern, 1e-2, inds, 500, True)
    n_seg, slot = kern.sharded_plan(40, S, 500, world)
    assert n_seg >= world and slot > 0
    gather = torch.zeros(world * slot, dtype=torch.uint8, device=dev)
    for rank in range(world):  # every "process" leaves the product of its segment operators in its slot
        kern.sharded_begin(x, pattern, 1e-2, inds, 500, rank, world, gather)
    total = torch.zeros_like(want)
    for rank in range(world):
        # (one kernel object plays all processes here: its own segment operators, which a real process still
        # holds from its call to _begin, are recomputed first)
        kern.sharded_begin(x, pattern, 1e-2, inds, 500, rank, world, gather)
        total += kern.sharded_end(inds, 40, 500, rank, world, gather)
        assert "time-sharded" in kern.last_kernel_name or rank == 0
    kern.sync()
    want, total = want.cpu().numpy(), total.cpu().numpy()
    np.testing.assert_allclose(total[:, 0], want[:, 0], rtol=1e-6)
    g_want, g_got = want[:, 1:].reshape(40, 7, 16), total[:, 1:].reshape(40, 7, 16)
    scale = np.abs(g_want).max(-1, keepdims=True)
    # (the sums are differences LL(all) - LL(warm-up): pi row entries are small remainders of O(1) terms)
    assert np.all(np.abs(g_got - g_want) <= 2e-4 * np.abs(g_want) + 2e-6 * np.maximum(scale, 1.0))
