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


# ---- round 2: the forward sweep in "z-form" with delayed rescaling, checkpoints from the sweep, and the warm-up
# ---- term as one more segment (boundary_sweep_kernel, KernelArgs::ext_ck / warm_len)

def z_form_sweep(pp, data, delay_sites=2, block=4, every=8):
    """The forward sweep as boundary_sweep_kernel's low-latency path runs it: it carries the PREDICTED vector
    z(t) = alpha(t - 1) A, forms alpha(t) = emis(ob_t) .* z(t) as a by-product, and applies the rescaling factor of
    a block of `block` sites `delay_sites` sites into the NEXT block (exact: the recursion is linear).
    Returns (log-likelihood, [alpha after every `every`-th site, any scale], alpha after the last site)."""
    emis = orc._emission_rows(pp)
    pi = pp[ROW_PI]
    ll = np.log(pi.sum())
    z = orc.matvec_smc(pi / pi.sum(), pp)  # a step with the emission row of ones
    pending = 1.0
    checkpoints = []
    w = None
    for t, ob in enumerate(data):
        if t % block == delay_sites:
            z = z * pending  # the factor of the previous block
            pending = 1.0
        w = emis[int(ob)] * z  # alpha(t), scaled like z
        z = orc.matvec_smc(w, pp)
        if (t + 1) % block == 0:
            tot = z.sum()
            pending = 1.0 / tot
            ll += np.log(tot)
        if (t + 1) % every == 0:
            checkpoints.append(w.copy())
    z = z * pending
    # sum z(L) = sum alpha(L - 1): the rows of A sum to one
    return ll + np.log(z.sum()), checkpoints, w


def test_z_form_sweep_with_delayed_rescaling_is_the_forward_recursion(golden):
    data, _ = fixture_data(0)
    row = data[1][:403]  # ragged last block
    pp = golden["dm16_pp"].astype(np.float64)
    want_ll, _ = orc.psmc_ll_grad(pp, row)
    ll, cks, last = z_form_sweep(pp, row)
    np.testing.assert_allclose(ll, want_ll, rtol=1e-12)
    # the by-product is the forward vector, up to scale: checkpoints after sites 7, 15, ...
    alpha = pp[ROW_PI] / pp[ROW_PI].sum()
    for m, ck in enumerate(cks):
        alpha, _ = propagate(pp, alpha, row[8 * m : 8 * m + 8])
        np.testing.assert_allclose(ck / ck.sum(), alpha, rtol=1e-11, atol=1e-300)
    full, _ = propagate(pp, pp[ROW_PI] / pp[ROW_PI].sum(), row)
    np.testing.assert_allclose(last / last.sum(), full, rtol=1e-11, atol=1e-300)


def segment_gradient_from_checkpoints(pp, data, alpha_in, alpha_end, checkpoints, beta_out, every=8):
    """segment_gradient() without its own forward pass over the segment: the forward vectors come from
    checkpoints[m] = vector after site `every` (m + 1) - 1 of the segment (any scale) and are re-run for `every`
    sites at a time, as the segment passes do with the vectors the forward sweep left (KernelArgs::ext_ck)."""
    m_states, length = pp.shape[1], len(data)
    emis = orc._emission_rows(pp)
    b, d, u, v = pp[ROW_B], pp[ROW_D], pp[ROW_U], pp[ROW_V]
    grad = np.zeros((7, m_states))
    beta = beta_out / np.dot(beta_out, alpha_end)
    ob_last = int(data[-1])
    if ob_last >= 0:
        grad[ROW_E0 + ob_last] += alpha_end * beta
    n_groups = (length + every - 1) // every
    for grp in range(n_groups - 1, -1, -1):
        lo, hi = grp * every, min(length, (grp + 1) * every)
        x = alpha_in if grp == 0 else checkpoints[grp - 1]
        xs = [x / x.sum()]
        for s in range(lo, hi):
            a = orc.matvec_smc(xs[-1], pp) * emis[int(data[s])]
            xs.append(a / a.sum())
        beta = beta / np.dot(beta, xs[-1])  # re-impose beta . alpha == 1 against the re-run vector
        for s in range(hi - 1, lo - 1, -1):
            ob, xin = int(data[s]), xs[s - lo]
            w = emis[ob] * beta
            tail = np.cumsum((v * w)[::-1])[::-1] - v * w
            beta_new = adjoint_step(pp, beta, ob)
            scale = 1.0 / np.dot(xin, beta_new)
            above = np.cumsum(xin[::-1])[::-1] - xin
            weighted = np.cumsum(u * xin) - u * xin
            grad[ROW_B] += b * above * w * scale
            grad[ROW_D] += d * xin * w * scale
            grad[ROW_V] += v * weighted * w * scale
            grad[ROW_U] += u * xin * tail * scale
            beta = beta_new * scale
            if s > 0:
                ob_prev = int(data[s - 1])
                if ob_prev >= 0:
                    grad[ROW_E0 + ob_prev] += xin * beta
    grad[ROW_PI] = (alpha_in / alpha_in.sum()) * beta
    return grad


@pytest.mark.parametrize("n_seg", [2, 5])
def test_segment_passes_from_sweep_checkpoints_and_fused_warmup(golden, n_seg):
    """(i) the segment passes fed by the sweep's checkpoints give the oracle's gradient; (ii) the warm-up term
    LL(first ov sites) scored as one more segment - started from pi, closed with a vector of ones - and
    subtracted gives the gradient of LL(whole row) - LL(first ov sites)."""
    data, miss = fixture_data(0)
    row = miss[0][:640]
    pp = golden["dm16_pp"].astype(np.float64)
    want_ll, want_grad = orc.psmc_ll_grad(pp, row)
    seg_len = 640 // n_seg // 16 * 16 + 16  # a multiple of 16 like the kernel's; the last segment is shorter
    cuts = list(range(0, 640, seg_len)) + [640]
    segments = [row[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    ll, cks, last = z_form_sweep(pp, row)
    np.testing.assert_allclose(ll, want_ll, rtol=1e-12)
    _, betas = boundaries_from_sweeps(pp, segments)
    pi_n = pp[ROW_PI] / pp[ROW_PI].sum()
    grad = np.zeros_like(want_grad)
    for g, seg in enumerate(segments):
        first = cuts[g] // 8  # record m of the sweep = vector after site 8 (m + 1) - 1
        alpha_in = pi_n if g == 0 else cks[first - 1]
        alpha_end = last if cuts[g + 1] == 640 else cks[cuts[g + 1] // 8 - 1]
        part = segment_gradient_from_checkpoints(pp, seg, alpha_in, alpha_end, cks[first:], betas[g + 1])
        grad[:ROW_PI] += part[:ROW_PI]
        if g == 0:
            grad[ROW_PI] = part[ROW_PI]
    np.testing.assert_allclose(grad, want_grad, rtol=1e-9, atol=1e-12)

    ov = 132  # not a multiple of the checkpoint spacing: the warm-up segment runs its own forward pass
    warm_ll, warm_grad = orc.psmc_ll_grad(pp, row[:ov])
    warm_part = segment_gradient(pp, row[:ov], pi_n, np.ones(pp.shape[1]))
    np.testing.assert_allclose(warm_part, warm_grad, rtol=1e-9, atol=1e-12)
    fused = grad - warm_part
    np.testing.assert_allclose(fused, want_grad - warm_grad, rtol=1e-9, atol=1e-11)
    assert want_ll - warm_ll < 0  # (the term the whole-term entries return: log p(chunk | warm-up))
