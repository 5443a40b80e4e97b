"""SVGD iterations per second with the device-side likelihood of this repository.

NOT the reference's optimiser stack: phlash uses blackjax.svgd + optax.amsgrad
(src/phlash/mcmc.py:178-199), neither of which is installable here, and the reference pins no
numbers for them - so this harness is a plain-torch RBF-kernel SVGD with
AMSGrad, used ONLY to exercise the whole per-iteration path the way mcmc.fit does
(mcmc.py:275-279): sample a minibatch with replacement, evaluate
c[1] * l2 (model.py:57, weight N / S, mcmc.py:244) and the N(0, 1) prior on log(rho / theta)
(model.py:11-21) for all particles, update the particles.  The AFS term is omitted (no AFS data).

    python tools/svgd_demo.py [iterations] [minibatch S]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402
from phlash_b200 import model  # noqa: E402
from phlash_b200.data import _chunk_het_matrix  # noqa: E402
from phlash_b200.gpu import _PSMCKernelBase  # noqa: E402

PATTERN, THETA, OVERLAP = "14*1+1*2", 1e-2, 500


def log_prior_and_grad(x):
    """-0.5 * log(rho/theta)^2 with rho/theta = 0.1 + 9.9 sigmoid(x[:, -1]) (params.py:110-113)."""
    z = x[:, -1].detach().requires_grad_(True)
    lp = -0.5 * torch.log(0.1 + 9.9 * torch.sigmoid(z)) ** 2
    (g,) = torch.autograd.grad(lp.sum(), z)
    grad = torch.zeros_like(x)
    grad[:, -1] = g
    return lp.detach(), grad


def svgd_direction(x, score):
    """phi_i = mean_j [ k(x_j, x_i) score_j + grad_{x_j} k(x_j, x_i) ], RBF kernel, median bandwidth."""
    d2 = torch.cdist(x, x) ** 2
    h = torch.median(d2) / np.log(x.shape[0] + 1.0) + 1e-12
    k = torch.exp(-d2 / h)
    repulse = (2.0 / h) * (k.sum(1, keepdim=True) * x - k @ x)
    return (k @ score + repulse) / x.shape[0]


def run(n_iter=50, S=5, n_bins=3_000_000, device=0, seed=0):
    dev = torch.device("cuda", device)
    het = synth.het_matrix(1, n_bins, seed)
    chunks = _chunk_het_matrix(het, OVERLAP, 50_000)
    kern = _PSMCKernelBase(16, chunks, device=device)
    xs = np.load(os.path.join(ROOT, "benchdata", "particles_M16.npz"))["xs"]
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    n_chunks = chunks.shape[0]
    m = torch.zeros_like(x)
    v = torch.zeros_like(x)
    vmax = torch.zeros_like(x)
    gen = torch.Generator(device=dev).manual_seed(seed)
    lr, b1, b2, eps = 0.1, 0.9, 0.999, 1e-8
    history = []

    def one_iteration(it):
        nonlocal x, m, v, vmax
        inds = torch.randint(0, n_chunks, (S,), generator=gen, device=dev)
        l2, g_l2 = model.hmm_term_value_and_grad(kern, x, PATTERN, THETA, inds, OVERLAP, weight=n_chunks / S)
        lp, g_lp = log_prior_and_grad(x)
        phi = svgd_direction(x, g_l2 + g_lp)
        m = b1 * m + (1 - b1) * phi
        v = b2 * v + (1 - b2) * phi * phi
        vmax = torch.maximum(vmax, v)
        step = lr * (m / (1 - b1 ** (it + 1))) / (torch.sqrt(vmax / (1 - b2 ** (it + 1))) + eps)
        x = (x + step).contiguous()  # ascent on the log density
        return float((l2 + lp).mean())

    for it in range(3):
        one_iteration(it)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(3, 3 + n_iter):
        history.append(one_iteration(it))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert torch.isfinite(x).all()
    return {"svgd_iters_per_s": n_iter / dt, "ms_per_iter": 1e3 * dt / n_iter, "particles": int(x.shape[0]),
            "minibatch_chunks": S, "bins_per_chunk": int(chunks.shape[1]), "n_chunks": int(n_chunks),
            "log_density_evaluations_per_iteration": 1,
            "mean_log_density_first": history[0], "mean_log_density_last": history[-1],
            "note": "plain-torch RBF SVGD + AMSGrad harness, not blackjax/optax; likelihood, warm-up, parameter "
                    "construction and their gradients run in this repository's CUDA kernels"}


if __name__ == "__main__":
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    print(json.dumps(run(n_iter, S)))
