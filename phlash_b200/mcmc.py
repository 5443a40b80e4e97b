"""The per-iteration likelihood dispatch of ``phlash.mcmc.fit`` (reference: src/phlash/mcmc.py:116-140,
201-247, 275-304) on top of the device-resident kernels of this package.

What the reference does every iteration on the HOST - draw ``inds`` with jax.random.choice, gather
``warmup_chunks[inds]`` in NumPy, hand both to the jitted SVGD step, which calls back into Python for the
CUDA kernel (mcmc.py:275-279, gpu.py:441-465) - is here one sampling kernel plus one library call, both on
the device, recorded ONCE into a CUDA graph together with the particle update and replayed per iteration.

The particle update itself (blackjax.svgd + optax.amsgrad, mcmc.py:178-199) is third-party code that stays
with the caller: ``fit_loop`` takes it as ``update`` - any callable
``update(x, score) -> None`` that moves the particle matrix ``x`` [B, P] in place given
``score`` = d log_density / d x.  tools/svgd_torch.py holds a plain-torch stand-in (RBF-kernel SVGD +
AMSGrad) used by the benchmark; the prior term of log_density (model.py:11-21) is passed as
``log_prior_grad``; the AFS term (model.py:58-70) is omitted when no AFS is given, as in the reference.
"""

from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from phlash_b200 import model


@dataclass
class FitResult:
    particles: "object"                 # torch float64 [B, P] on the device
    iterations: int                     # iterations actually run (early stop: mcmc.py:298-303)
    seconds: float                      # wall-clock of the loop (device synchronised on both sides)
    minibatch_size: int
    n_chunks: int
    elpd_trace: List[float] = field(default_factory=list)   # the EMA the reference shows in its progress bar
    stopped_early: bool = False
    graph_replays: int = 0
    timed_iterations: int = 0           # iterations inside `seconds` (those after the warm-up)

    @property
    def iters_per_s(self) -> float:
        return self.timed_iterations / self.seconds if self.seconds > 0 else float("nan")


def fit_loop(chunks: np.ndarray, x0, pattern: str, theta: float, update: Callable, *, niter: int = 1000,
             overlap: int = 500, minibatch_size: Optional[int] = None, M: int = 16, test_het: Optional[np.ndarray] = None,
             max_samples: Optional[int] = None, elpd_cutoff: int = 100, elpd_every: int = 10,
             log_prior_grad: Optional[Callable] = None, seed: int = 0, device: int = 0, rank: int = 0, world: int = 1,
             use_graph: bool = True, downsample_rng: Optional[np.random.Generator] = None, warmup_iters: int = 0,
             kern=None, test_kern=None) -> FitResult:
    """Run the reference's schedule.

    chunks: int8 [N, overlap + L] as ``init_mcmc_data`` returns them (every rank passes the same matrix: the
    data are replicated like in the reference, gpu.py:346-351); x0: float64 [B, P] initial particles (NumPy
    or torch); ``update(x, score)`` moves them.  With ``world`` > 1 (one process per GPU, process group
    initialised by the caller) every rank scores its shard of the minibatch and ONE all-reduce of the
    per-particle sums [B, 1 + 7 M] joins them; all ranks then make the same update.
    ``warmup_iters`` iterations are run (and counted in ``iterations``) before the clock starts."""
    import torch

    from phlash_b200.gpu import _PSMCKernelBase

    dev = torch.device("cuda", device)
    # ---- mcmc.py:116-139: minibatch size, down-sampling of the chunk matrix
    S = minibatch_size or model.default_minibatch_size(len(chunks), niter)
    if downsample_rng is None:
        downsample_rng = np.random.default_rng(seed)
    # ---- mcmc.py:201-209: the kernel object; the warm-up columns stay resident with the data (fused warm-up)
    if kern is None:
        chunks = model.downsample_chunks(chunks, S, niter, downsample_rng)
        kern = _PSMCKernelBase(M, np.ascontiguousarray(chunks), double_precision=False, device=device, overlap=overlap)
    N = kern._N  # (a kernel object handed in already holds the rows to sample from)
    x = torch.as_tensor(np.asarray(x0) if not torch.is_tensor(x0) else x0, dtype=torch.float64, device=dev).contiguous().clone()
    B = int(x.shape[0])
    # ---- mcmc.py:211-238: the ELPD on held-out contigs (forward only, un-chunked)
    if test_het is not None and test_kern is None:
        test = np.asarray(test_het)
        if max_samples is not None:
            test = test[:max_samples]  # mcmc.py:208-210
        test_kern = model.elpd_kernel(M, test, device=device)
    weight = model.minibatch_weight(N, S)  # mcmc.py:240-247: c = [1, N / S, 1]
    kern.reserve(B, S, overlap)
    inds = torch.empty(S, dtype=torch.int64, device=dev)
    kern.set_iteration(0)

    def one_iteration():
        # mcmc.py:277-279
        kern.sample_minibatch(seed, S, out=inds)
        # one library call on one GPU; with several processes the chunks - or, for fewer chunks than processes,
        # the segments of the parallel-in-time gradient - are sharded and joined by one all-reduce
        _, g = model.hmm_term_value_and_grad(kern, x, pattern, theta, inds, overlap, weight, rank=rank, world=world)
        if log_prior_grad is not None:
            g = g + log_prior_grad(x)
        update(x, g)

    graph = None
    side = torch.cuda.Stream(device=dev)
    eager_before_capture = 3  # the first iterations run eagerly on the capture stream (kernels loaded, NCCL warmed up)
    result = FitResult(particles=x, iterations=0, seconds=0.0, minibatch_size=S, n_chunks=N)
    ema = None
    best = None  # (iteration, ema)
    t0 = None
    for i in range(niter):
        if i == warmup_iters:
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
        if use_graph and graph is None and i >= eager_before_capture:
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                one_iteration()
            # (recording does not execute: the sampler's iteration counter is untouched)
        if graph is not None:
            graph.replay()
            result.graph_replays += 1
        else:
            with torch.cuda.stream(side):
                one_iteration()
        result.iterations = i + 1
        if test_kern is not None and i % elpd_every == 0:
            # mcmc.py:287-304
            side.synchronize()
            torch.cuda.synchronize(dev)
            assert bool(torch.isfinite(x).all()), "non-finite particle (mcmc.py:281-285)"
            e = float(model.elpd_hmm_term(test_kern, x, pattern, theta, rank=rank, world=world))  # particles sharded
            ema = e if ema is None else 0.9 * ema + 0.1 * e
            result.elpd_trace.append(ema)
            if best is None or ema > best[1]:
                best = (i, ema)
            if i - best[0] > elpd_cutoff:
                result.stopped_early = True
                break
    torch.cuda.synchronize(dev)
    if t0 is not None:
        result.seconds = time.perf_counter() - t0
        result.timed_iterations = result.iterations - warmup_iters
    assert bool(torch.isfinite(x).all()), "non-finite particle (mcmc.py:281-285)"
    return result
