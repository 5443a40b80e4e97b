"""NumPy restatement of the SCALED site recursion the sm_100a throughput kernel runs
(phlash_b200/csrc/psmc_scaled.cuh), written on the oracle's conventions so that the algebra can be
checked on the CPU against oracle.psmc_oracle.psmc_ll_grad (tests/test_scaled_recursion_math.py).
Test infrastructure only.

The reference recursion (src/phlash/hmm.py:68-82) is  x_t = e(ob_t) .* (x_{t-1} A)  with
A = strictLower(1 b^T) + diag(d) + strictUpper(u v^T)  (hmm.py:52-65).  The kernel carries the
PRE-emission vector z_t = x_{t-1} A divided by v:  zh = z / s,  s = (1, v_1, ..., v_{M-1}), and treats
"emission of site t, then transition" as one step whose coefficients for ob = 0 are

    beta_k = b_k / s_k     sigma_i = e0_i s_i     delta_k = d_k e0_k     pi_i = u_i e0_i s_i
    zh'_k = beta_k S_k + delta_k y_k + P_k,   S_k = sum_{i>k} sigma_i y_i,   P_k = sum_{i<k} pi_i y_i

(4 fused multiply-adds per state, no separate multiply), where y = zh for ob = 0 and
y = (e(ob) / e0) .* zh otherwise.  The normaliser  sum_k x_k = sum_k sigma_k y_k  falls out of the S chain.
"""

import numpy as np

ROW_B, ROW_D, ROW_U, ROW_V, ROW_E0, ROW_E1, ROW_PI = range(7)


def coefficients(pp, dtype=np.float64):
    pp = np.asarray(pp, dtype=dtype)
    b, d, u, v, e0, e1, pi = pp
    s = v.copy()
    s[0] = 1
    co = dict(
        s=s, beta=b / s, sigma=e0 * s, delta=d * e0, pi_c=u * e0 * s,
        ratio=np.stack([np.ones_like(e0), e1 / e0, 1 / e0]),  # ob = 0, 1, -1 (missing) -> index ob
        zh0=pi / s,
    )
    return co


def _step(co, y):
    """zh' and the two chains for input y (already multiplied by the emission ratio)."""
    sy = co["sigma"] * y
    py = co["pi_c"] * y
    S = np.cumsum(sy[::-1])[::-1] - sy  # sum_{i>k}
    P = np.cumsum(py) - py              # sum_{i<k}
    return co["beta"] * S + co["delta"] * y + P, S, P, S[0] + sy[0]


def loglik_grad(pp, data, dtype=np.float64, rescale_every=4, telescope_v=False):
    """ll and dlog (layout of oracle.psmc_ll_grad) by the scaled recursion, all arithmetic in `dtype`
    except the accumulation of log-scales / window sums, which the kernel does in double too."""
    data = np.asarray(data)
    L = len(data)
    co = coefficients(pp, dtype)
    m = len(co["s"])
    one = dtype(1)
    # ---- forward: step j consumes the emission of obs[j - 1] (j = 0: none) and one transition
    zh = np.empty((L + 1, m), dtype=dtype)
    fac = np.ones(L + 1, dtype=dtype)  # factor applied to zh[j + 1] after step j
    zh[0] = co["zh0"]
    log_scale = 0.0
    for j in range(L):
        ob = -1 if j == 0 else int(data[j - 1])
        y = zh[j] * co["ratio"][ob]
        out, _, _, tot = _step(co, y)
        if j % rescale_every == rescale_every - 1:
            f = one / tot
            fac[j + 1] = f
            out = out * f
            log_scale -= np.log(np.float64(f))
        zh[j + 1] = out
    y = zh[L] * co["ratio"][int(data[L - 1])]
    total = np.sum(co["sigma"] * y, dtype=dtype)
    ll = log_scale + np.log(np.float64(total))
    # ---- adjoint
    A = {k: np.zeros(m, dtype=np.float64) for k in ("delta", "sigma", "pi", "beta", "P")}
    gam = np.zeros((3, m), dtype=np.float64)  # posterior sums of ob = 0 (unused), 1, missing
    w = co["sigma"] * co["ratio"][int(data[L - 1])] / total  # adjoint of zh[L]; w . zh[L] == 1
    gamma_last = (zh[L] * w).astype(np.float64)
    gam_tot_extra = gamma_last.copy()
    if int(data[L - 1]) != 0:
        gam[int(data[L - 1])] += gamma_last
    for j in range(L - 1, -1, -1):
        ob = -1 if j == 0 else int(data[j - 1])
        wp = w * fac[j + 1]  # adjoint of the un-rescaled output of step j
        y = zh[j] * co["ratio"][ob]
        _, S, P, _ = _step(co, y)
        bw = co["beta"] * wp
        Bp = np.cumsum(bw) - bw              # sum_{k<i} beta_k w'_k
        Q = np.cumsum(wp[::-1])[::-1] - wp   # sum_{k>i} w'_k
        wy = co["sigma"] * Bp + co["delta"] * wp + co["pi_c"] * Q
        A["delta"] += y * wp
        A["sigma"] += y * Bp
        A["pi"] += y * Q
        A["beta"] += S * wp
        A["P"] += P * wp
        if ob != 0:
            gam[ob] += y * wy
        w = wy * co["ratio"][ob]
    gamma_first = (zh[0] * w).astype(np.float64)
    # ---- to the reference's layout
    g_b = co["beta"] * A["beta"]
    g_d = co["delta"] * A["delta"]
    g_u = co["pi_c"] * A["pi"]
    g_sigma = co["sigma"] * A["sigma"]
    if telescope_v:
        g_v = g_sigma + g_u - g_b + gamma_last - gamma_first
    else:
        g_v = A["P"].copy()
    g_v[0] = 0.0
    gam_tot = g_d + g_sigma + g_u + gam_tot_extra
    dlog = np.zeros((7, m))
    dlog[ROW_B], dlog[ROW_D], dlog[ROW_U], dlog[ROW_V] = g_b, g_d, g_u, g_v
    dlog[ROW_E1] = gam[1]
    dlog[ROW_E0] = gam_tot - gam[1] - gam[-1]
    dlog[ROW_PI] = gamma_first
    return ll, dlog
