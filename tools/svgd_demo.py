"""SVGD iterations per second with the device-side likelihood of this repository, on the reference's
schedule (phlash_b200.mcmc.fit_loop: minibatch rule, down-sampling, device-side sampling, N / S weight, ELPD
every 10th iteration with the early stop) and a plain-torch optimiser stand-in (tools/svgd_torch.py - NOT
blackjax / optax, parity unpinned).

    python tools/svgd_demo.py [iterations] [minibatch S]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import svgd_torch  # noqa: E402
from benchdata import synth  # noqa: E402
from phlash_b200 import mcmc  # noqa: E402
from phlash_b200.data import _chunk_het_matrix  # noqa: E402

PATTERNS = {16: "14*1+1*2", 32: "30*1+1*2", 64: "62*1+1*2"}
THETA, OVERLAP = 1e-2, 500


def run(n_iter=50, S=5, n_bins=3_000_000, device=0, seed=0, M=16, particles=500, rank=0, world=1, use_graph=True,
        chunks=None, test_bins=0, kern=None):
    """n_iter timed iterations (after 5 untimed ones) of the whole per-iteration path."""
    dev = torch.device("cuda", device)
    if chunks is None:
        chunks = _chunk_het_matrix(synth.het_matrix(1, n_bins, seed), OVERLAP, 50_000)
    xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:particles]
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    opt = svgd_torch.SvgdAmsgrad(x)
    test_het = synth.het_matrix(1, test_bins, seed=101) if test_bins else None
    warm = 5
    res = mcmc.fit_loop(chunks, x, PATTERNS[M], THETA, opt, niter=warm + n_iter, overlap=OVERLAP, minibatch_size=S, M=M,
                        test_het=test_het, log_prior_grad=svgd_torch.log_prior_grad, seed=seed, device=device, rank=rank,
                        world=world, use_graph=use_graph, warmup_iters=warm, kern=kern)
    timed = res.timed_iterations
    return {"svgd_iters_per_s": timed / res.seconds, "ms_per_iter": 1e3 * res.seconds / timed, "particles": int(x.shape[0]),
            "M": M, "minibatch_chunks": S, "bins_per_chunk": int(chunks.shape[1]), "n_chunks": int(res.n_chunks),
            "log_density_evaluations_per_iteration": 1, "cuda_graph": bool(use_graph), "world": world,
            "elpd_every_10th": bool(test_bins), "elpd_trace_first_last": (res.elpd_trace[:1] + res.elpd_trace[-1:]),
            "note": "reference schedule (mcmc.py:116-140, 275-304) with a plain-torch RBF SVGD + AMSGrad stand-in, not "
                    "blackjax/optax (parity unpinned); sampling, warm-up, likelihood, parameter construction and their "
                    "gradients run in this repository's CUDA kernels, one CUDA-graph replay per iteration"}


if __name__ == "__main__":
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    print(json.dumps(run(n_iter, S)))
