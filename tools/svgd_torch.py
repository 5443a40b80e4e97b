"""Plain-torch stand-in for the reference's optimiser stack, used ONLY to drive the likelihood path the way
mcmc.fit does when measuring SVGD iterations per second.

NOT blackjax.svgd + optax.amsgrad (src/phlash/mcmc.py:178-199): neither is installable in this image and the
reference pins no numbers for them ("SVGD iters/sec: parity unpinned").  What is here:
RBF-kernel SVGD with the median bandwidth heuristic, AMSGrad on the SVGD direction, and the N(0, 1) prior on
log(rho / theta) of model.py:11-21.  Every operation is a fixed-shape torch op on device tensors (the step
counter included), so that a whole iteration can be recorded into a CUDA graph by phlash_b200.mcmc.fit_loop.
"""
import numpy as np
import torch


def log_prior_grad(x):
    """d/dx of -0.5 * log(rho/theta)^2 with rho/theta = 0.1 + 9.9 sigmoid(x[:, -1]) (params.py:110-113,
    model.py:11-21), in closed form (no autograd: capturable)."""
    z = x[:, -1]
    s = torch.sigmoid(z)
    r = 0.1 + 9.9 * s
    g = torch.zeros_like(x)
    g[:, -1] = -torch.log(r) / r * 9.9 * s * (1.0 - s)
    return g


def svgd_direction(x, score):
    """phi_i = mean_j [ k(x_j, x_i) score_j + grad_{x_j} k(x_j, x_i) ], RBF kernel, median bandwidth."""
    d2 = torch.cdist(x, x) ** 2
    h = torch.median(d2) / np.log(x.shape[0] + 1.0) + 1e-12
    k = torch.exp(-d2 / h)
    repulse = (2.0 / h) * (k.sum(1, keepdim=True) * x - k @ x)
    return (k @ score + repulse) / x.shape[0]


class SvgdAmsgrad:
    """update(x, score): one SVGD + AMSGrad ascent step on the particle matrix, in place."""

    def __init__(self, x, lr=0.1, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.m = torch.zeros_like(x)
        self.v = torch.zeros_like(x)
        self.vmax = torch.zeros_like(x)
        self.t = torch.zeros((), dtype=torch.float64, device=x.device)  # step counter ON THE DEVICE

    def __call__(self, x, score):
        phi = svgd_direction(x, score)
        self.t += 1.0
        self.m.mul_(self.b1).add_(phi, alpha=1 - self.b1)
        self.v.mul_(self.b2).addcmul_(phi, phi, value=1 - self.b2)
        torch.maximum(self.vmax, self.v, out=self.vmax)
        mhat = self.m / (1 - self.b1 ** self.t)
        vhat = self.vmax / (1 - self.b2 ** self.t)
        x.add_(self.lr * mhat / (torch.sqrt(vhat) + self.eps))  # ascent on the log density
