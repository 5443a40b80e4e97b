"""The N > 1 host logic (shard -> per-particle sums -> one all-reduce) on CPU with gloo,
world_size 2.  The per-shard "kernel results" come from the oracle; what is tested is the
sharding / packing / collective plumbing that bench.py and ShardedPSMCKernel use on the GPUs."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle, psmc_oracle as orc
from phlash_b200 import distributed as pd


def test_shard_bounds_cover_everything_exactly_once():
    for n in (0, 1, 5, 8, 595, 1000):
        for world in (1, 2, 3, 8):
            spans = [pd.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    ll = torch.tensor(rng.normal(size=(3, 5)))
    dlog = torch.tensor(rng.normal(size=(3, 5, 7, 4)).astype(np.float32))
    packed = pd.pack_per_particle(ll, dlog)
    ll_sum, dlog_sum = pd.unpack_per_particle(packed, 4)
    np.testing.assert_allclose(ll_sum.numpy(), ll.numpy().sum(1))
    np.testing.assert_allclose(dlog_sum.numpy(), dlog.numpy().astype(np.float64).sum(1), rtol=1e-12)
    empty = pd.pack_per_particle(ll[:, :0], dlog[:, :0])
    assert empty.shape == (3, 29) and float(empty.abs().sum()) == 0.0


class _OracleKernel:
    """Stands in for gpu._PSMCKernelBase.evaluate_device in the CPU test."""

    def __init__(self, data):
        self.data = data

    def evaluate_device(self, params6, pi, inds, grad):
        B, S = params6.shape[0], inds.shape[0]
        pa = np.concatenate([params6.numpy(), pi.numpy()[:, None]], axis=1)  # [B, 7, M]
        pa = np.broadcast_to(pa[:, None], (B, S, 7, pa.shape[-1])).reshape(B * S, 7, -1)
        ll, dlog = c_oracle.loglik_batch(self.data, np.tile(inds.numpy(), B), pa, grad=True)
        return torch.tensor(ll.reshape(B, S)), torch.tensor(dlog.reshape(B, S, 7, -1))


def _worker(rank, world, port, data, pps, inds, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sharded = pd.ShardedPSMCKernel(_OracleKernel(data))
        assert (sharded.rank, sharded.world) == (rank, world)
        ll, dlog = sharded.loglik_grad_sum(torch.tensor(pps[:, :6]), torch.tensor(pps[:, 6]), torch.tensor(inds))
        out[rank] = (ll.numpy().copy(), dlog.numpy().copy())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_chunks", [5, 1])  # 1 chunk: rank 1 gets an empty shard
def test_two_ranks_reproduce_the_single_rank_sum(n_chunks):
    rng = np.random.default_rng(1)
    data = (rng.uniform(size=(6, 300)) < 0.08).astype(np.int8)
    pps, _, _ = orc.synth_particles(8, 3, seed=2)
    inds = rng.integers(0, 6, size=n_chunks).astype(np.int64)
    manager = mp.Manager()
    out = manager.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, data, pps, inds, out), nprocs=2, join=True)
    whole_ll, whole_dlog = _OracleKernel(data).evaluate_device(
        torch.tensor(pps[:, :6]), torch.tensor(pps[:, 6]), torch.tensor(inds), True)
    for rank in (0, 1):
        ll, dlog = out[rank]
        np.testing.assert_allclose(ll, whole_ll.numpy().sum(1), rtol=1e-12)
        np.testing.assert_allclose(dlog, whole_dlog.numpy().sum(1), rtol=1e-10, atol=1e-14)


class _FakeElpdKernel:
    """Stands in for the ELPD kernel object: hmm_term(x, ...) -> one value per particle (a fixed function of x)."""

    _N = 3

    def hmm_term(self, x, pattern, theta, inds, overlap, weight, grad=False):
        assert not grad and int(inds.shape[0]) == self._N
        return (x * x).sum(dim=1) * weight - theta, None


def _elpd_worker(rank, world, port, xs, out):
    from phlash_b200 import model

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = model.elpd_hmm_term(_FakeElpdKernel(), torch.tensor(xs), "14*1+1*2", 0.25, rank=rank, world=world)
        out[rank] = float(e)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_particles", [7, 1])  # 1 particle: rank 1 scores nothing
def test_elpd_with_particles_sharded_over_two_ranks(n_particles):
    """model.elpd_hmm_term shards the PARTICLES over the processes (the ELPD is a mean over particles, mcmc.py:221-236):
    every rank ends up with the single-process value."""
    from phlash_b200 import model

    xs = np.random.default_rng(3).normal(size=(n_particles, 5))
    want = float(model.elpd_hmm_term(_FakeElpdKernel(), torch.tensor(xs), "14*1+1*2", 0.25))
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_elpd_worker, args=(2, _free_port(), xs, out), nprocs=2, join=True)
    for rank in (0, 1):
        np.testing.assert_allclose(out[rank], want, rtol=1e-13)
