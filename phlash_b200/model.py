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

from phlash_b200.distributed import all_reduce_sum, pack_per_particle, shard_bounds, unpack_per_particle


def sample_minibatch(rng, n_chunks: int, minibatch_size: int):
    """inds ~ choice(N, (S,)) WITH replacement (reference: mcmc.py:277)."""
    return rng.integers(0, n_chunks, size=minibatch_size)


def default_minibatch_size(n_chunks: int, niter: int) -> int:
    """reference: mcmc.py:119-121"""
    return max(1, min(5, int(n_chunks / niter)))


def hmm_term_value_and_grad(kern, x, pattern: str, theta: float, inds, overlap: int, weight: float = 1.0,
                            rank: int = 0, world: int = 1):
    """kern: ``gpu._PSMCKernelBase`` built on FULL chunks [N, overlap + L]; x: torch float64 CUDA
    tensor [B, P]; inds: torch int64 CUDA tensor [S] (the whole minibatch, identical on all ranks).
    Returns (weight * l2 [B], weight * d l2 / d x [B, P]) as float64 tensors; with ``world > 1``
    every rank scores its shard of the minibatch and one all-reduce joins the per-particle sums."""
    import torch

    B = int(x.shape[0])
    M = kern._M
    params7 = kern.params_from_particles(x, pattern, theta)
    lo, hi = shard_bounds(int(inds.shape[0]), rank, world)
    packed = torch.zeros((B, 1 + 7 * M), dtype=torch.float64, device=x.device)
    if hi > lo:
        ll, dlog = kern.evaluate_warmup_device(params7, inds[lo:hi].contiguous(), overlap, True)
        pack_per_particle(ll, dlog, out=packed)
    if world > 1:
        all_reduce_sum(packed)
    l2, dlog_b = unpack_per_particle(packed, M)
    cot = dlog_b.to(params7.dtype).contiguous()
    grad_x = kern.params_vjp(x, pattern, theta, cot)
    return weight * l2, weight * grad_x
