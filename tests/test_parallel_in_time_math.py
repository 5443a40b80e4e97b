"""The parallel-in-time decomposition used by the CUDA kernels (segment transfer operators chained in
float64; boundary vectors from the chain or from two sequential sweeps; independent gradient passes per
segment), restated in NumPy on top of the oracle's primitives and checked against the oracle's own
sequential log-likelihood and gradient.  Exactness of the decomposition is a property of the algebra,
not of the GPU: CPU only."""

import numpy as np
import pytest

from conftest import fixture_data
from oracle import psmc_oracle as orc

ROW_B, ROW_D, ROW_U, ROW_V, ROW_E0, ROW_E1, ROW_PI = range(7)


def propagate(pp, x, data):
    """x through all sites of `data` (transition, emission, rescale): (normalised vector, log of the scale)"""
    emis = orc._emission_rows(pp)
    log_scale = 0.0
    for ob in data:
        x = orc.matvec_smc(x, pp) * emis[int(ob)]
        s = x.sum()
        x = x / s
        log_scale += np.log(s)
    return x, log_scale


def transfer_operator(pp, data):
    """rows[i] = e_i propagated through the segment, normalised; logs[i] = log of what was divided out"""
    m = pp.shape[1]
    rows, logs = np.empty((m, m)), np.empty(m)
    for i in range(m):
        rows[i], logs[i] = propagate(pp, np.eye(m)[i], data)
    return rows, logs


def adjoint_step(pp, beta, ob):
    """beta <- A (emis(ob) .* beta)"""
    b, d, u, v = pp[ROW_B], pp[ROW_D], pp[ROW_U], pp[ROW_V]
    w = orc._emission_rows(pp)[int(ob)] * beta
    below = np.cumsum(b * w) - b * w
    tail = np.cumsum((v * w)[::-1])[::-1] - v * w
    return below + d * w + u * tail


def segment_gradient(pp, data, alpha_in, beta_out):
    """the oracle's adjoint recursion (psmc_ll_grad) over ONE segment, started from alpha_in and closed with
    beta_out rescaled to beta . alpha == 1; the pi row is alpha_in .* beta at the segment start"""
    pp_seg = pp.copy()
    pp_seg[ROW_PI] = alpha_in
    m, length = pp.shape[1], len(data)
    b, d, u, v = pp[ROW_B], pp[ROW_D], pp[ROW_U], pp[ROW_V]
    emis = orc._emission_rows(pp)
    alphas = np.empty((length + 1, m))
    alphas[0] = alpha_in
    for s in range(length):
        a = orc.matvec_smc(alphas[s], pp) * emis[int(data[s])]
        alphas[s + 1] = a / a.sum()
    grad = np.zeros((7, m))
    beta = beta_out / np.dot(beta_out, alphas[length])
    for s in range(length - 1, -1, -1):
        ob, x = int(data[s]), alphas[s]
        if ob >= 0:
            grad[ROW_E0 + ob] += alphas[s + 1] * beta
        w = emis[ob] * beta
        tail = np.cumsum((v * w)[::-1])[::-1] - v * w
        beta_new = adjoint_step(pp, beta, ob)
        scale = 1.0 / np.dot(x, beta_new)
        above = np.cumsum(x[::-1])[::-1] - x
        weighted = np.cumsum(u * x) - u * x
        grad[ROW_B] += b * above * w * scale
        grad[ROW_D] += d * x * w * scale
        grad[ROW_V] += v * weighted * w * scale
        grad[ROW_U] += u * x * tail * scale
        beta = beta_new * scale
    grad[ROW_PI] = alpha_in * beta
    return grad


def boundaries_from_operators(pp, segments):
    ops = [transfer_operator(pp, seg) for seg in segments]
    alpha = pp[ROW_PI] / pp[ROW_PI].sum()
    ll = np.log(pp[ROW_PI].sum())
    alphas = [alpha]
    for rows, logs in ops:
        nxt = (alpha * np.exp(logs - logs.max())) @ rows
        ll += logs.max() + np.log(nxt.sum())
        alpha = nxt / nxt.sum()
        alphas.append(alpha)
    beta = np.ones(pp.shape[1])
    betas = [beta]
    for rows, logs in reversed(ops):
        beta = np.exp(logs - logs.max()) * (rows @ beta)
        beta = beta / beta.max()
        betas.append(beta)
    return ll, alphas, betas[::-1]


def boundaries_from_sweeps(pp, segments):
    alpha = pp[ROW_PI] / pp[ROW_PI].sum()
    alphas = [alpha]
    for seg in segments:
        alpha, _ = propagate(pp, alpha, seg)
        alphas.append(alpha)
    beta = np.ones(pp.shape[1])
    betas = [beta]
    for seg in reversed(segments):
        for ob in seg[::-1]:
            beta = adjoint_step(pp, beta, ob)
            beta = beta / beta.sum()
        betas.append(beta)
    return alphas, betas[::-1]


@pytest.mark.parametrize("missing", [False, True])
@pytest.mark.parametrize("n_seg", [1, 3, 7])
def test_decomposition_is_exact(golden, missing, n_seg):
    data, miss = fixture_data(0)
    row = (miss if missing else data)[2][:700]
    pp = golden["dm16_pp"].astype(np.float64)
    want_ll, want_grad = orc.psmc_ll_grad(pp, row)
    cuts = np.linspace(0, len(row), n_seg + 1).astype(int)
    segments = [row[a:b] for a, b in zip(cuts[:-1], cuts[1:])]

    ll, alphas, betas = boundaries_from_operators(pp, segments)
    np.testing.assert_allclose(ll, want_ll, rtol=1e-12)
    grad = np.zeros_like(want_grad)
    for g, seg in enumerate(segments):
        part = segment_gradient(pp, seg, alphas[g], betas[g + 1])
        grad[:ROW_PI] += part[:ROW_PI]
        if g == 0:
            grad[ROW_PI] = part[ROW_PI]
    np.testing.assert_allclose(grad, want_grad, rtol=1e-9, atol=1e-12)

    # the two sequential sweeps give the same boundary vectors up to scale
    a2, b2 = boundaries_from_sweeps(pp, segments)
    for g in range(n_seg + 1):
        np.testing.assert_allclose(a2[g], alphas[g], rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(b2[g] / b2[g].max(), betas[g] / betas[g].max(), rtol=1e-10, atol=1e-300)
