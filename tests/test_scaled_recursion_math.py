"""CPU check of an alternative formulation that was analysed in round 2 and NOT adopted ("What bounds the
throughput kernel"): the site recursion on the diagonally rescaled pre-emission vector (tests/scaled_math.py) with
4 FMAs per state in the forward step.  The algebra is exact - it reproduces the oracle's log-likelihood and gradient to
round-off - but (i) its pre-multiplied coefficients fl(d * e) would bias the stay weights in fp32, and (ii) obtaining the
v row of the gradient by telescoping from the other accumulators loses ~7 digits.  Kept as a record of why."""

import numpy as np

import scaled_math as sm
from oracle import psmc_oracle as orc


def test_scaled_recursion_reproduces_the_oracle_and_telescoping_loses_digits():
    pps, _, _ = orc.synth_particles(16, 3, seed=0)
    het = orc.synth_het_matrix(1, 1500, seed=1)[0]
    worst_direct, worst_telescoped = 0.0, 0.0
    for pp in pps:
        ll0, g0 = orc.psmc_ll_grad(pp, het)
        for telescope in (False, True):
            ll, g = sm.loglik_grad(pp, het, telescope_v=telescope)
            np.testing.assert_allclose(ll, ll0, rtol=1e-13)
            err = np.max(np.abs(g - g0) / (np.abs(g0) + 1e-12 * np.abs(g0).max(-1, keepdims=True)))
            if telescope:
                worst_telescoped = max(worst_telescoped, err)
            else:
                worst_direct = max(worst_direct, err)
    assert worst_direct < 1e-11
    assert 1e-12 < worst_telescoped < 1e-6  # exact in exact arithmetic, ~1e-9 in float64: useless in float32


def test_a_pre_rounded_stay_weight_is_amplified_in_the_gradient():
    """Rounding d to float32 (a 3e-8 relative perturbation, the size of the error of a pre-multiplied coefficient
    fl(d * e)) moves the gradient by two orders of magnitude more than the perturbation itself - a SYSTEMATIC
    error that per-site roundings of separate multiplications do not have - but stays well inside the 1e-4 bar."""
    pps, _, _ = orc.synth_particles(16, 2, seed=3)
    het = orc.synth_het_matrix(1, 20_000, seed=2)[0]
    for pp in pps:
        _, g = orc.psmc_ll_grad(pp, het)
        pp32 = pp.copy()
        pp32[1] = pp[1].astype(np.float32).astype(np.float64)
        _, g32 = orc.psmc_ll_grad(pp32, het)
        pert = np.max(np.abs(pp32[1] - pp[1]) / pp[1])
        rel = np.max(np.abs(g32[1] - g[1]) / np.abs(g[1]))
        assert 30 * pert < rel < 1e-4, (pert, rel)
