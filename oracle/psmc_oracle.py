"""CPU oracle (float64, NumPy) for the PSMC coalescent-HMM hot path of jthlab/phlash.

TEST INFRASTRUCTURE ONLY.  Nothing under ``phlash_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker / reported
baseline, never as the product path.

It is a *restatement* of the reference's algorithm (the reference needs JAX,
which is absent from this image), written from the maths rather than translated
line by line.  Every function names the reference lines it follows.

Pinning status ("Oracle"):
  * pinned against the reference's own known-answer / property tests
    (tests/test_oracle_reference_pins.py) and against golden vectors produced by
    executing the reference's *unmodified Python sources* through a NumPy-backed
    shim of the jax API (tests/golden/make_golden.py; jax itself is not
    installable here), and - on the GPU box - against the reference's own CUDA
    kernel compiled in double precision (oracle/build_ref.py -> oracle/_ref/).
  * the XLA/JAX execution of the reference cannot run anywhere in this project,
    so for that leg specifically: parity unpinned.
"""

from __future__ import annotations

import numpy as np

# order of the rows of the [7, M] parameter block (reference: src/phlash/params.py:16-23,
# src/phlash/gpu.py:483-491)
ROW_B, ROW_D, ROW_U, ROW_V, ROW_E0, ROW_E1, ROW_PI = range(7)
P = 7


# --------------------------------------------------------------------------------------
# data layer
# --------------------------------------------------------------------------------------
def chunk_het_matrix(het_matrix: np.ndarray, overlap: int, chunk_size: int) -> np.ndarray:
    """Chunk geometry of ``_chunk_het_matrix`` (reference: src/phlash/data.py:37-61).

    Every row of ``het_matrix`` is clipped to [-1, 1], right-padded with -1 up to a
    multiple of ``W = chunk_size + overlap`` and cut into ``ceil(L / W)`` windows of
    ``W`` bins that *start* every ``chunk_size`` bins (so consecutive windows share
    ``overlap`` bins, and the tail of a long row is never covered).
    """
    data = np.clip(np.asarray(het_matrix), -1, 1).astype(np.int8)
    assert data.ndim == 2
    n, length = data.shape
    width = chunk_size + overlap
    n_chunks = -(-length // width)
    padded = np.full((n, n_chunks * width), -1, dtype=np.int8)
    padded[:, :length] = data
    out = np.empty((n, n_chunks, width), dtype=np.int8)
    for k in range(n_chunks):
        out[:, k, :] = padded[:, k * chunk_size : k * chunk_size + width]
    return out.reshape(n * n_chunks, width)


# --------------------------------------------------------------------------------------
# size history -> stationary distribution and expected coalescence times
# --------------------------------------------------------------------------------------
def expm1inv(x):
    """1 / expm1(x), stable for large x (reference: src/phlash/size_history.py:17-22)."""
    x = np.asarray(x, dtype=np.float64)
    big = x > 10.0
    xs = np.where(big, 1.0, x)
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        return np.where(big, -np.exp(-x) / np.expm1(-x), 1.0 / np.expm1(xs))


def surv(t, c):
    """Survival function of the coalescence time at t[1:], then 0
    (reference: src/phlash/size_history.py:123-128)."""
    t = np.asarray(t, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    hazard = np.cumsum(c[:-1] * np.diff(t))
    return np.append(np.exp(-hazard), 0.0)


def stationary_pi(t, c):
    """P(coalescence in interval i) (reference: src/phlash/size_history.py:131-138)."""
    mass = -np.diff(surv(t, c))
    return np.concatenate([[1.0 - mass.sum()], mass])


def ect(t, c):
    """Expected coalescence time inside each interval
    (reference: src/phlash/size_history.py:170-193)."""
    t = np.asarray(t, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    ci = c[:-1]
    t0, t1 = t[:-1], t[1:]
    dt = t1 - t0
    is_zero = np.isclose(ci, 0.0)
    is_huge = np.isinf(ci) | (ci > 100.0)
    cs = np.where(is_zero | is_huge, 1.0, ci)
    inner = 1.0 / cs + t0 - dt * expm1inv(cs * dt)
    e = np.where(is_zero, 0.5 * (t0 + t1), np.where(is_huge, t0, inner))
    e = np.append(e, t[-1] + 1.0 / c[-1])
    return np.maximum(e, 1e-20)


# --------------------------------------------------------------------------------------
# transition matrix of the SMC' coalescent HMM
# --------------------------------------------------------------------------------------
def expQ(r, c, n):
    """Closed-form exp(Q) of the 3-state chain (unrecombined / floating / recombined)
    (reference: src/phlash/transition.py:9-34)."""
    r = np.float64(r)
    c = np.float64(c)
    disc = np.sqrt((c * n) ** 2 - 2.0 * c * (n - 2) * r + r * r) / 2.0
    mean = (r + c * n) / 2.0
    half = (r - c * n) / 2.0
    cosh_term = (np.exp(disc - mean) + np.exp(-(disc + mean))) / 2.0
    if disc < 1e-6:
        sinhc_term = np.exp(-mean) * (1.0 + 1.0 / 6.0)  # NB: u_safe == 1 on this branch
        # The reference evaluates exp(-v) * (1 + u_safe**2 / 6) with u_safe = 1 when u is
        # small (src/phlash/transition.py:18-22), i.e. the series is applied to the
        # *substituted* value.  Reproduced as written.
    else:
        sinhc_term = (np.exp(disc - mean) - np.exp(-(disc + mean))) / 2.0 / disc
    p11 = cosh_term - half * sinhc_term
    p12 = r * sinhc_term
    p21 = c * sinhc_term
    p22 = cosh_term + half * sinhc_term
    return np.array(
        [[p11, p12, 1.0 - p11 - p12], [p21, p22, 1.0 - p21 - p22], [0.0, 0.0, 1.0]]
    )


def _transition_pieces(t, c, rho, n=2):
    """Shared by transition_matrix / transition_bands (reference: src/phlash/transition.py:37-83)."""
    t = np.asarray(t, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    m = len(t)
    e = ect(t, c)
    c_adj = c * (n - 1)
    dt = np.diff(t)
    # time grid t0, e0, t1, e1, ..., t_{M-1}, e_{M-1}; 2M-1 steps, rate c_k on both halves of k
    grid = np.stack([t, e], axis=1).reshape(-1)
    step = np.diff(grid)
    rate = np.repeat(c, 2)[:-1]
    cum = [np.eye(3)]
    for k in range(2 * m - 1):
        if np.isclose(step[k], 0.0):
            pk = np.eye(3)
        else:
            pk = expQ(2.0 * step[k] * rho, step[k] * rate[k], n)
        cum.append(cum[-1] @ pk)
    absorbing = np.array([[0.0, 0.0, 1.0]] * 3)
    cum.append(cum[-1] @ absorbing)
    cum = np.array(cum)  # 2M+1 cumulative products
    at_t = cum[0::2]  # M+1: state at t_0..t_{M-1}, then "infinity"
    at_e = cum[1::2]  # M:   state at ect_0..ect_{M-1}
    lower_col = np.diff(at_t[:, 0, 2])  # value of every entry below the diagonal in column j
    rest = np.append(t[1:] - e[:-1], np.nan)  # time left in interval k after ect_k
    with np.errstate(invalid="ignore"):
        back = np.append(-np.expm1(-rest[:-1] * c_adj[:-1]), 1.0)
        stay = np.append(np.exp(-rest[:-1] * c_adj[:-1]), 0.0)
    diag = at_e[:, 0, 0] + at_e[:, 0, 1] * back + at_e[:, 0, 2] - at_t[:-1, 0, 2]
    p_float = at_e[:, 0, 1] * stay
    p_pass = np.append(np.exp(-dt * c_adj[:-1]), 0.0)
    p_coal = np.append(-np.expm1(-dt * c_adj[:-1]), 1.0)
    lo, hi = 1e-8, 1.0 - 1e-8
    return lower_col, diag, np.clip(p_float, lo, hi), np.clip(p_pass, lo, hi), np.clip(p_coal, lo, hi)


def transition_matrix(t, c, rho, n=2):
    """Dense M x M transition matrix (reference: src/phlash/transition.py:37-85)."""
    lower_col, diag, p_float, p_pass, p_coal = _transition_pieces(t, c, rho, n)
    m = len(diag)
    a = np.zeros((m, m))
    for i in range(m):
        for j in range(m):
            if i > j:
                a[i, j] = lower_col[j]
            elif i == j:
                a[i, j] = diag[i]
            else:
                a[i, j] = p_float[i] * np.prod(p_pass[i + 1 : j]) * p_coal[j]
    return a


# --------------------------------------------------------------------------------------
# HMM parameter block
# --------------------------------------------------------------------------------------
def params_from_dm(t, c, theta, rho, require_m16=False):
    """[7, M] block (b, d, u, v, emis0, emis1, pi) (reference: src/phlash/params.py:32-55).

    The reference hard-asserts M == 16 (params.py:35); the restatement works for any M
    (``require_m16`` re-enables the assertion for drop-in tests).
    """
    t = np.asarray(t, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    m = len(t)
    if require_m16:
        assert m == 16, "require M=16"
    lo, hi = 1e-20, 1.0 - 1e-20
    mut = theta * ect(t, c)
    emis0 = np.clip(np.exp(-mut), lo, hi)
    emis1 = np.clip(-np.expm1(-mut), lo, hi)
    pi = np.clip(stationary_pi(t, c), lo, hi)
    a = np.clip(transition_matrix(t, c, rho), lo, hi)
    sub = np.diagonal(a, -1)
    sup = np.diagonal(a, 1)
    ratio = a[0, 1:] / a[0, 1]
    out = np.zeros((P, m))
    out[ROW_B, :-1] = sub
    out[ROW_D] = np.diagonal(a)
    out[ROW_U, :-1] = sup / ratio
    out[ROW_V, 1:] = ratio
    out[ROW_E0] = emis0
    out[ROW_E1] = emis1
    out[ROW_PI] = pi
    return out


def parse_pattern(pattern: str):
    """PSMC pattern string -> epoch widths (reference: src/phlash/util.py:8-37)."""
    widths = []
    for tok in pattern.split("+"):
        if "*" in tok:
            k, w = (int(x) for x in tok.split("*"))
        else:
            k, w = 1, int(tok)
        widths += [w] * k
    if not widths or any(w <= 0 for w in widths):
        raise ValueError("could not parse pattern")
    return widths


def softplus(x):
    x = np.asarray(x, dtype=np.float64)
    return np.logaddexp(0.0, x)


def softplus_inv(y):
    """reference: src/phlash/util.py:49-51"""
    y = np.asarray(y, dtype=np.float64)
    return y + np.log1p(-np.exp(-y))


def particle_from_linear(pattern, t1, tM, c, theta, rho):
    """Unconstrained particle vector [t_tr(2), c_tr(len(pattern)), rho_over_theta_tr]
    (reference: src/phlash/params.py:68-92; flattening order = dataclass field order
    t_tr, c_tr, rho_over_theta_tr, params.py:58-66)."""
    widths = parse_pattern(pattern)
    c = np.asarray(c, dtype=np.float64)
    assert len(widths) == len(c)
    ratio = (rho / theta - 0.1) / 9.9
    return np.concatenate(
        [[np.log(t1), np.log(tM - t1)], softplus_inv(c), [np.log(ratio) - np.log1p(-ratio)]]
    )


def particle_to_dm(x, pattern, theta):
    """Particle -> (t[M], c[M], rho) (reference: src/phlash/params.py:94-131)."""
    widths = parse_pattern(pattern)
    m = sum(widths)
    x = np.asarray(x, dtype=np.float64)
    assert x.shape == (2 + len(widths) + 1,)
    t1, dtm = np.exp(x[0]), np.exp(x[1])
    t = np.concatenate([[0.0], np.geomspace(t1, t1 + dtm, m - 1)])
    c = np.repeat(softplus(x[2 : 2 + len(widths)]), widths)
    rho = theta * (0.1 + 9.9 / (1.0 + np.exp(-x[-1])))
    return t, c, rho


def default_dm(m, theta, rho=None, t_max=15.0):
    """DemographicModel.default("<M>*1", ...) (reference: src/phlash/size_history.py:303-326)."""
    t = np.concatenate([[0.0], np.geomspace(1e-3, t_max, m - 1)])
    return t, np.ones(m), (theta if rho is None else rho)


# --------------------------------------------------------------------------------------
# the HMM: structured transition, scaled forward recursion, gradient
# --------------------------------------------------------------------------------------
def matvec_smc(x, pp):
    """x @ A in O(M) from the (b, d, u, v) factorisation (reference: src/phlash/hmm.py:52-65)."""
    x = np.asarray(x, dtype=np.float64)
    above = np.cumsum(x[::-1])[::-1] - x  # sum_{i>j} x[i]
    weighted = np.cumsum(pp[ROW_U] * x) - pp[ROW_U] * x  # sum_{i<j} u[i] x[i]
    return pp[ROW_B] * above + pp[ROW_D] * x + pp[ROW_V] * weighted


def dense_from_pp(pp):
    """The dense A that matvec_smc applies (test helper)."""
    m = pp.shape[1]
    a = np.zeros((m, m))
    for i in range(m):
        for j in range(m):
            a[i, j] = pp[ROW_B, j] if i > j else pp[ROW_D, j] if i == j else pp[ROW_U, i] * pp[ROW_V, j]
    return a


def _emission_rows(pp):
    # row -1 (missing) is all ones (reference: src/phlash/hmm.py:70-71)
    return np.stack([pp[ROW_E0], pp[ROW_E1], np.ones_like(pp[ROW_E0])])


def psmc_ll(pp, data):
    """Scaled forward algorithm: returns (final filtered distribution, log-likelihood).
    Transition first, then emission, then rescale (reference: src/phlash/hmm.py:68-82)."""
    pp = np.asarray(pp, dtype=np.float64)
    emis = _emission_rows(pp)
    alpha = pp[ROW_PI].copy()
    ll = 0.0
    for ob in np.asarray(data):
        alpha = matvec_smc(alpha, pp) * emis[int(ob)]
        norm = alpha.sum()
        alpha /= norm
        ll += np.log(norm)
    return alpha, ll


def psmc_ll_grad(pp, data):
    """Log-likelihood and its gradient with the layout the reference kernel returns after
    the host-side roll: ``dlog[g, m] = d ll / d log(theta[g, m])`` for rows
    (b, d, u, v, emis0, emis1) and ``pi[m] * d ll / d pi[m]`` for the pi row, zero where
    the parameter is zero (reference: src/phlash/gpu.py:575-692 for the quantity,
    :303-313 for the layout).  Computed here by the adjoint (backward) recursion."""
    pp = np.asarray(pp, dtype=np.float64)
    data = np.asarray(data)
    m = pp.shape[1]
    length = len(data)
    b, d, u, v = pp[ROW_B], pp[ROW_D], pp[ROW_U], pp[ROW_V]
    emis = _emission_rows(pp)
    alphas = np.empty((length + 1, m))
    alphas[0] = pp[ROW_PI]
    ll = 0.0
    for s in range(length):
        a = matvec_smc(alphas[s], pp) * emis[int(data[s])]
        norm = a.sum()
        alphas[s + 1] = a / norm
        ll += np.log(norm)
    grad = np.zeros((P, m))
    beta = np.ones(m)  # alphas[length] . beta == 1 throughout
    for s in range(length - 1, -1, -1):
        ob = int(data[s])
        x = alphas[s]
        if ob >= 0:
            grad[ROW_E0 + ob] += alphas[s + 1] * beta
        w = emis[ob] * beta
        below = np.cumsum(b * w) - b * w  # sum_{j<i} b[j] w[j]
        tail = np.cumsum((v * w)[::-1])[::-1] - v * w  # sum_{j>i} v[j] w[j]
        beta_new = below + d * w + u * tail
        scale = 1.0 / np.dot(x, beta_new)
        above = np.cumsum(x[::-1])[::-1] - x
        weighted = np.cumsum(u * x) - u * x
        ws = w * scale
        grad[ROW_B] += b * above * ws
        grad[ROW_D] += d * x * ws
        grad[ROW_V] += v * weighted * ws
        grad[ROW_U] += u * x * tail * scale
        beta = beta_new * scale
    grad[ROW_PI] = pp[ROW_PI] * beta
    return ll, grad


def psmc_ll_grad_forward_mode(pp, data):
    """The reference kernel's own algorithm - forward-mode sensitivities kept in log-parameter
    space for every row except pi, which is multiplied by pi at the end
    (src/phlash/gpu.py:600-691) - evaluated densely in float64.  O(7 M^3) per site, test use
    only: an independent check of the adjoint recursion in psmc_ll_grad."""
    pp = np.asarray(pp, dtype=np.float64)
    m = pp.shape[1]
    b, d, u, v = pp[ROW_B], pp[ROW_D], pp[ROW_U], pp[ROW_V]
    a_mat = dense_from_pp(pp)
    emis = _emission_rows(pp)
    h = pp[ROW_PI].copy()
    sens = np.zeros((P, m, m))  # sens[g, k, :] = d h / d (log) theta[g, k], scaled like h
    sens[ROW_PI] = np.eye(m)
    ll = 0.0
    for ob in np.asarray(data):
        ob = int(ob)
        new = sens @ a_mat
        for k in range(m):
            new[ROW_B, k, k] += b[k] * h[k + 1 :].sum()
            new[ROW_D, k, k] += d[k] * h[k]
            new[ROW_U, k, k + 1 :] += u[k] * h[k] * v[k + 1 :]
            new[ROW_V, k, k] += v[k] * np.dot(u[:k], h[:k])
        h = h @ a_mat
        if ob >= 0:
            new[ROW_E0 + ob] += np.diag(h)
        h = h * emis[ob]
        norm = h.sum()
        h /= norm
        sens = new * emis[ob] / norm
        ll += np.log(norm)
    grad = sens.sum(axis=2)
    grad[ROW_PI] *= pp[ROW_PI]
    return ll, grad


def warmup_pis(pp, warmup_rows):
    """Filtered distributions after the warm-up bins, starting from the stationary pi
    (reference: src/phlash/model.py:52-54)."""
    return np.stack([psmc_ll(pp, row)[0] for row in warmup_rows])


def hmm_term(pp, warmup_rows, data_rows):
    """Sum over the minibatch of the chunk log-likelihoods, each started from its own
    warm-up distribution (reference: src/phlash/model.py:52-57)."""
    total = 0.0
    for wrow, drow in zip(warmup_rows, data_rows):
        q = np.array(pp, dtype=np.float64)
        q[ROW_PI] = psmc_ll(pp, wrow)[0]
        total += psmc_ll(q, drow)[1]
    return total


# --------------------------------------------------------------------------------------
# synthetic "msprime-shaped" inputs - shared by tests and bench
# --------------------------------------------------------------------------------------
def synth_het_matrix(n_dip, n_bins, seed, het_lo=0.02, het_hi=0.12, mean_run=2000, miss_frac=0.01, miss_run=50):
    rng = np.random.default_rng(seed)
    out = np.empty((n_dip, n_bins), dtype=np.int8)
    for i in range(n_dip):
        n_runs = int(n_bins / mean_run * 2) + 16
        lens = rng.geometric(1.0 / mean_run, size=n_runs)
        while lens.sum() < n_bins:
            lens = np.concatenate([lens, rng.geometric(1.0 / mean_run, size=n_runs)])
        state0 = rng.integers(0, 2)
        levels = np.where((np.arange(len(lens)) + state0) % 2 == 0, het_lo, het_hi)
        rate = np.repeat(levels, lens)[:n_bins]
        row = (rng.random(n_bins) < rate).astype(np.int8)
        n_miss = int(miss_frac * n_bins / miss_run)
        starts = rng.integers(0, max(1, n_bins - miss_run), size=n_miss)
        for s in starts:
            row[s : s + miss_run] = -1
        out[i] = row
    return out


def synth_particles(m, n_particles, seed, theta=1e-2, rho=1e-2, sigma=1.0):
    """B particles around DemographicModel.default, jittered in unconstrained space
    (law of src/phlash/mcmc.py:186-195).  Returns ([B, 7, M] params, particle matrix, pattern)."""
    pattern = f"{m - 2}*1+1*2"
    t, c, _ = default_dm(m, theta, rho)
    n_epochs = len(parse_pattern(pattern))
    x0 = particle_from_linear(pattern, t[1], t[-1], np.ones(n_epochs), theta, rho)
    rng = np.random.default_rng(1000 + seed)
    xs = x0[None] + sigma * rng.standard_normal((n_particles, len(x0)))
    pps = []
    for x in xs:
        t_x, c_x, rho_x = particle_to_dm(x, pattern, theta)
        pps.append(params_from_dm(t_x, c_x, theta, rho_x))
    return np.stack(pps), xs, pattern
