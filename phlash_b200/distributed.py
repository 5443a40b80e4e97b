"""Multi-GPU evaluation, one process per GPU (torch.distributed; NCCL over NVLink on the box).

The reference replicates the data on every device and splits the minibatch indices across
Python threads (src/phlash/gpu.py:386-438), then concatenates on the host.  Here every rank owns a
kernel object on its own device, scores a contiguous shard of the minibatch and contributes to
ONE all-reduce of the per-particle sums [B, 1 + 7 M] per step: with the
6 parameter rows and pi shared by all chunks of a particle, both the log-likelihood and the
gradient are additive over chunks (reference: model.py:57 sums the minibatch).
"""

from __future__ import annotations

from typing import Tuple


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of ``n_items`` for ``rank`` (first ranks get the extra
    item; an empty shard is legal - the reference dead-locks on it, gpu.py:404)."""
    assert 0 <= rank < world
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_per_particle(ll, dlog, out=None):
    """ll [B, S] float64 and dlog [B, S, 7, M] -> [B, 1 + 7 M] float64 sums over the chunk axis
    (torch tensors on any device)."""
    import torch

    B = ll.shape[0]
    width = 1 + dlog.shape[2] * dlog.shape[3]
    if out is None:
        out = torch.empty((B, width), dtype=torch.float64, device=ll.device)
    if ll.shape[1] == 0:
        out.zero_()
        return out
    out[:, 0] = ll.sum(1)
    out[:, 1:] = dlog.sum(1, dtype=torch.float64).reshape(B, width - 1)
    return out


def unpack_per_particle(packed, m: int):
    """[B, 1 + 7 M] -> (ll [B], dlog [B, 7, M])"""
    return packed[:, 0], packed[:, 1:].reshape(packed.shape[0], 7, m)


def all_reduce_sum(packed):
    """The single collective of a step; a no-op without an initialised process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(packed)
    return packed


class ShardedPSMCKernel:
    """Per-particle log-likelihood and gradient of a minibatch, sharded over the ranks of the
    default process group.  ``kern`` is this rank's ``gpu._PSMCKernelBase`` (data replicated)."""

    def __init__(self, kern, rank: int = None, world: int = None):
        import torch.distributed as dist

        self.kern = kern
        live = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank() if live else 0)
        self.world = world if world is not None else (dist.get_world_size() if live else 1)

    def loglik_grad_sum(self, params6, pi, inds):
        """params6 [B, 6, M], pi [B, M] SHARED by the chunks of a particle, inds [S] (global minibatch,
        identical on every rank), all device tensors.  Returns (ll [B], dlog [B, 7, M]) summed over ALL S
        chunks, on every rank, in a fresh buffer per call (B (1 + 7 M) doubles) that the caller owns.
        With a shared pi the terms are additive over chunks; for the reference's per-chunk warm-up pi
        (model.py:52-55) use model.hmm_term_value_and_grad, whose fused warm-up keeps them additive."""
        import torch

        lo, hi = shard_bounds(int(inds.shape[0]), self.rank, self.world)
        B, M = int(params6.shape[0]), int(params6.shape[2])
        packed = torch.empty((B, 1 + 7 * M), dtype=torch.float64, device=params6.device)
        if hi > lo:
            ll, dlog = self.kern.evaluate_device(params6, pi, inds[lo:hi].contiguous(), True)
            pack_per_particle(ll, dlog, out=packed)
        else:
            packed.zero_()
        all_reduce_sum(packed)
        return unpack_per_particle(packed, M)
