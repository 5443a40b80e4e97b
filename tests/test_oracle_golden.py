"""The oracle (oracle/psmc_oracle.py, oracle/psmc_oracle.c) against golden vectors produced by the
reference's own unmodified sources (tests/golden/make_golden.py).  CPU only."""

import numpy as np
import pytest

from oracle import c_oracle, psmc_oracle as orc

PATTERN16 = "14*1+1*2"


def rel(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def test_size_history_default_model(golden):
    t, c = golden["dm16_t"], golden["dm16_c"]
    assert rel(orc.ect(t, c), golden["dm16_ect"]) < 1e-14
    assert rel(orc.stationary_pi(t, c), golden["dm16_pi"]) < 1e-14
    np.testing.assert_allclose(orc.surv(t, c), golden["dm16_surv"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("n", [2, 5])
def test_transition_matrix_default_model(golden, n):
    t, c = golden["dm16_t"], golden["dm16_c"]
    key = "dm16_A" if n == 2 else "dm16_A_n5"
    assert rel(orc.transition_matrix(t, c, 1e-2, n), golden[key]) < 1e-13


@pytest.mark.parametrize("m", [32, 64])
def test_transition_matrix_larger_m(golden, m):
    t, c, rho = orc.default_dm(m, 1e-2, 2e-2)
    assert rel(orc.transition_matrix(t, c, rho), golden[f"dm{m}_A"]) < 1e-13
    assert rel(orc.ect(t, c), golden[f"dm{m}_ect"]) < 1e-14
    assert rel(orc.stationary_pi(t, c), golden[f"dm{m}_pi"]) < 1e-14


def test_extreme_rates(golden):
    """c ~ 0, c > 100 and ordinary intervals: the guards of SizeHistory.ect (size_history.py:170-193)."""
    t = np.concatenate([[0.0], np.geomspace(1e-3, 15.0, 15)])
    c = golden["odd_c"]
    assert rel(orc.ect(t, c), golden["odd_ect"]) < 1e-14
    assert rel(orc.stationary_pi(t, c), golden["odd_pi"], 1e-30) < 1e-12
    assert rel(orc.transition_matrix(t, c, 5e-3), golden["odd_A"], 1e-30) < 1e-12
    assert rel(orc.params_from_dm(t, c, 2e-2, 5e-3), golden["odd_pp"], 1e-30) < 1e-9


def test_expq_grid(golden):
    got = np.stack([orc.expQ(r, c, int(n)) for r, c, n in golden["expq_args"]])
    np.testing.assert_allclose(got, golden["expq_vals"], rtol=1e-14, atol=1e-300)


def test_params_from_dm(golden):
    got = orc.params_from_dm(golden["dm16_t"], golden["dm16_c"], 1e-2, 1e-2, require_m16=True)
    assert rel(got, golden["dm16_pp"], 1e-30) < 1e-12
    # the structural zeros of the reference (params.py:48-51)
    assert got[orc.ROW_B, -1] == 0 and got[orc.ROW_U, -1] == 0 and got[orc.ROW_V, 0] == 0


def test_particles_to_params(golden):
    """MCMCParams.from_linear / to_dm / PSMCParams.from_dm (params.py:68-131, 32-55)."""
    x0 = orc.particle_from_linear(PATTERN16, 1e-4, 15.0, np.ones(15), 1e-2, 1e-2)
    np.testing.assert_allclose(x0, golden["part_x"][0], rtol=1e-14, atol=1e-15)
    for i, x in enumerate(golden["part_x"]):
        t, c, rho = orc.particle_to_dm(x, PATTERN16, 1e-2)
        np.testing.assert_allclose(t, golden["part_t"][i], rtol=1e-14)
        np.testing.assert_allclose(c, golden["part_c"][i], rtol=1e-14)
        np.testing.assert_allclose(rho, golden["part_rho"][i], rtol=1e-14)
        assert rel(orc.params_from_dm(t, c, 1e-2, rho), golden["part_pp"][i], 1e-30) < 1e-8


def test_matvec(golden):
    np.testing.assert_allclose(orc.matvec_smc(golden["matvec_v"], golden["dm16_pp"]), golden["matvec_out"], rtol=1e-14)


def test_forward_recursion(golden, seed):
    """psmc_ll of the reference on its own test fixtures, with and without missing data."""
    k = 0
    for params in (golden["dm16_pp"], golden["part_pp"][2]):
        for mat in (golden[f"data_s{seed}"], golden[f"missing_s{seed}"]):
            ll_c, alpha_c = c_oracle.loglik_batch(mat, np.arange(3), np.stack([params] * 3), want_alpha=True)
            for row in range(3):
                alpha, ll = orc.psmc_ll(params, mat[row])
                np.testing.assert_allclose(ll, golden[f"hmm_ll_s{seed}"][k], rtol=1e-13)
                np.testing.assert_allclose(alpha, golden[f"hmm_alpha_s{seed}"][k], rtol=1e-10, atol=1e-18)
                np.testing.assert_allclose(ll_c[row], golden[f"hmm_ll_s{seed}"][k], rtol=1e-13)
                np.testing.assert_allclose(alpha_c[row], golden[f"hmm_alpha_s{seed}"][k], rtol=1e-10, atol=1e-18)
                k += 1


def test_gradient_vs_reference_finite_differences(golden):
    """The adjoint gradient against central differences of the REFERENCE's psmc_ll in
    log-parameter space (make_golden.py, section E)."""
    ll, grad = orc.psmc_ll_grad(golden["fd_pp"], golden["fd_row"])
    np.testing.assert_allclose(ll, golden["fd_ll"], rtol=1e-13)
    np.testing.assert_allclose(grad, golden["fd_grad"], rtol=2e-4, atol=1e-7)
    zero = golden["fd_pp"] == 0
    assert np.all(grad[zero] == 0)


def test_gradient_adjoint_vs_forward_mode(golden, seed):
    """Two independent derivations: the adjoint recursion and the reference kernel's own
    forward-mode algorithm (gpu.py:600-691) restated densely."""
    pp = golden["part_pp"][1 + seed]
    row = golden[f"missing_s{seed}"][5][:400]
    ll1, g1 = orc.psmc_ll_grad(pp, row)
    ll2, g2 = orc.psmc_ll_grad_forward_mode(pp, row)
    np.testing.assert_allclose(ll1, ll2, rtol=1e-13)
    np.testing.assert_allclose(g1, g2, rtol=1e-9, atol=1e-14)
    ll3, g3 = c_oracle.loglik_batch(row[None], np.zeros(1, dtype=np.int64), pp[None], grad=True)
    np.testing.assert_allclose(ll3[0], ll1, rtol=1e-13)
    np.testing.assert_allclose(g3[0], g1, rtol=1e-10, atol=1e-16)


def test_gradient_vs_torch_autograd(golden):
    """Third derivation: reverse-mode autodiff of the same recursion (closest analogue of
    jax.value_and_grad(psmc_ll), tests/test_gpu.py:59-64)."""
    import torch

    pp = torch.tensor(golden["part_pp"][3], dtype=torch.float64)
    row = golden["missing_s1"][2][:200]
    logp = torch.where(pp > 0, pp.log(), torch.zeros_like(pp)).requires_grad_(True)
    par = torch.where(pp > 0, logp.exp(), torch.zeros_like(pp))
    b, d, u, v, e0, e1, pi = par
    emis = torch.stack([e0, e1, torch.ones_like(e0)])
    alpha, ll = pi, torch.zeros((), dtype=torch.float64)
    for ob in row:
        above = torch.flip(torch.cumsum(torch.flip(alpha, [0]), 0), [0]) - alpha
        weighted = torch.cumsum(u * alpha, 0) - u * alpha
        alpha = (b * above + d * alpha + v * weighted) * emis[int(ob)]
        norm = alpha.sum()
        alpha = alpha / norm
        ll = ll + norm.log()
    ll.backward()
    ll_o, g_o = orc.psmc_ll_grad(golden["part_pp"][3], row)
    np.testing.assert_allclose(ll.item(), ll_o, rtol=1e-13)
    np.testing.assert_allclose(logp.grad.numpy(), g_o, rtol=1e-9, atol=1e-14)


def test_hmm_term_of_log_density(golden):
    """Warm-up from the stationary pi, then the chunk (model.py:52-57) against the reference's
    log_density run with its own PureJaxPSMCKernel."""
    chunks = golden["model_chunks"]
    warm, body = chunks[:, :50], chunks[:, 50:]
    inds = golden["model_inds"]
    for i in range(3):
        t, c, rho = orc.particle_to_dm(golden["part_x"][i], PATTERN16, 1e-2)
        pp = orc.params_from_dm(t, c, 1e-2, rho)
        np.testing.assert_allclose(orc.hmm_term(pp, warm[inds], body[inds]), golden["model_l2"][i], rtol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_chunk_geometry(golden, tag):
    ov, cs = (int(v) for v in golden[f"chunk_{tag}_geom"])
    got = orc.chunk_het_matrix(golden[f"chunk_{tag}_in"], ov, cs)
    assert got.dtype == np.int8
    np.testing.assert_array_equal(got, golden[f"chunk_{tag}_out"])
