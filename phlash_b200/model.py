"""The HMM term of ``phlash.model.log_density`` (reference: src/phlash/model.py:50-57, weighted as
in :71-72 and src/phlash/mcmc.py:240-247) evaluated entirely on the device for all particles at
once, with its gradient w.r.t. the particles:

    x [B, P] --params_from_particles--> theta [B, 7, M]
             --fused warm-up loglik+grad over the minibatch--> l2 [B], d l2 / d log(theta) [B, 7, M]
             --params_vjp--> d l2 / d x [B, P]

The reference does the first and last arrow in XLA (float64), the warm-up in a 500-step lax.scan
differentiated by JAX, and only the middle in CUDA, with a host round trip per evaluation.  The
prior and AFS terms of log_density (model.py:11-21, 58-70) are O(dim) per particle and stay with
the caller.
"""

from __future__ import annotations

from phlash_b200.distributed import all_reduce_sum, shard_bounds


_GATHER = {}


def _gather_buffer(kern, n_bytes: int, device):
    """the all-gather buffer of the time-sharded path, kept per kernel object (stable address: CUDA graphs)"""
    import torch

    buf = _GATHER.get(id(kern))
    if buf is None or buf.numel() < n_bytes or buf.device != device:
        buf = torch.empty(n_bytes, dtype=torch.uint8, device=device)
        _GATHER[id(kern)] = buf
    return buf[:n_bytes]


def sample_minibatch(rng, n_chunks: int, minibatch_size: int):
    """inds ~ choice(N, (S,)) WITH replacement (reference: mcmc.py:277)."""
    return rng.integers(0, n_chunks, size=minibatch_size)


def default_minibatch_size(n_chunks: int, niter: int) -> int:
    """reference: mcmc.py:119-121"""
    return max(1, min(5, int(n_chunks / niter)))


def hmm_term_value_and_grad(kern, x, pattern: str, theta: float, inds, overlap: int, weight: float = 1.0,
                            rank: int = 0, world: int = 1, grad: bool = True):
    """kern: ``gpu._PSMCKernelBase`` built on FULL chunks [N, overlap + L]; x: torch float64 CUDA
    tensor [B, P]; inds: torch int64 CUDA tensor [S] (the whole minibatch, identical on all ranks).
    Returns (weight * l2 [B], weight * d l2 / d x [B, P]) as float64 tensors.  One library call on one
    GPU (phb_hmm_term_device); with ``world > 1`` every rank scores its shard of the minibatch, one
    all-reduce joins the per-particle sums [B, 1 + 7 M], and every rank finishes on the total."""
    if world == 1:
        return kern.hmm_term(x, pattern, theta, inds, overlap, weight, grad)
    S = int(inds.shape[0])
    if grad and S < world:
        # fewer chunks than GPUs (the reference's default S <= 5): shard the SEGMENTS of the parallel-in-time
        # gradient instead of the chunks - one all-gather of the segment operators, then the usual all-reduce
        n_seg, slot = kern.sharded_plan(int(x.shape[0]), S, overlap, world)
        if n_seg > 0:
            import torch
            import torch.distributed as dist

            gather = _gather_buffer(kern, world * slot, x.device)
            kern.sharded_begin(x, pattern, theta, inds, overlap, rank, world, gather)
            dist.all_gather_into_tensor(gather, gather[rank * slot:(rank + 1) * slot])
            sums = kern.sharded_end(inds, int(x.shape[0]), overlap, rank, world, gather)
            all_reduce_sum(sums)
            return kern.hmm_term_finish(x, pattern, theta, sums, weight, grad)
    lo, hi = shard_bounds(S, rank, world)
    sums = kern.hmm_term_sums(x, pattern, theta, inds[lo:hi].contiguous(), overlap, grad)
    all_reduce_sum(sums)
    return kern.hmm_term_finish(x, pattern, theta, sums, weight, grad)


def elpd_kernel(M: int, test_het, double_precision: bool = False, device: int = 0):
    """Kernel object for the reference's ELPD evaluation (mcmc.py:213-236): the un-chunked test contigs
    [N_test, L] with ONE missing warm-up bin in front of every row (`warmup = full([N_test, 1], -1)`)."""
    import numpy as np

    from phlash_b200.gpu import _PSMCKernelBase

    het = np.asarray(test_het)
    full = np.concatenate([np.full((het.shape[0], 1), -1, dtype=np.int8), np.clip(het, -1, 1).astype(np.int8)], axis=1)
    return _PSMCKernelBase(M, full, double_precision=double_precision, device=device)


def elpd_hmm_term(test_kern, x, pattern: str, theta: float, rank: int = 0, world: int = 1):
    """HMM part of the expected log-predictive density: mean over particles of log_density with
    c = (0, 1, 1) over all test contigs (mcmc.py:221-236; the AFS part stays with the caller).
    Forward only: no gradient is formed.

    With one process per GPU (``world`` > 1, process group initialised by the caller) the PARTICLES are sharded:
    the evaluation is parallel in time through segment transfer operators, i.e. M times the arithmetic of a plain
    forward pass whatever the number of pairs (88 ms for 500 particles x 2.5 M bins on one GPU, as much as thirty
    S = 1 steps), so every process scores its block of the particles and one all-reduce of a scalar joins them."""
    import torch

    inds = torch.arange(test_kern._N, dtype=torch.int64, device=x.device)
    B = int(x.shape[0])
    if world <= 1:
        value, _ = test_kern.hmm_term(x, pattern, theta, inds, 1, 1.0, grad=False)
        return value.mean()
    import torch.distributed as dist

    lo, hi = shard_bounds(B, rank, world)
    total = torch.zeros((), dtype=torch.float64, device=x.device)
    if hi > lo:
        value, _ = test_kern.hmm_term(x[lo:hi].contiguous(), pattern, theta, inds, 1, 1.0, grad=False)
        total = value.sum()
    dist.all_reduce(total)
    return total / B


def downsample_chunks(chunks, minibatch_size: int, niter: int, rng):
    """Keep at most 5 * S * niter chunk rows, chosen without replacement (reference: mcmc.py:124-139:
    "in expectation, we will sample at most S * niter rows of the data").  `rng` is a
    numpy.random.Generator (the reference seeds one from its jax key, so the choice itself is not
    reproducible across the two code bases - only the rule is)."""
    limit = 5 * minibatch_size * niter
    if len(chunks) > limit:
        chunks = rng.choice(chunks, size=(limit,), replace=False)
    return chunks


def minibatch_weight(n_chunks: int, minibatch_size: int) -> float:
    """N / S: the factor that makes the minibatch term an unbiased estimate of the sum over all chunks
    (reference: mcmc.py:240-247, c = [1, N / S, 1])."""
    return n_chunks / minibatch_size
