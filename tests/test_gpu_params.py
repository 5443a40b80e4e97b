"""Parameter construction on the device (particle -> b, d, u, v, emis0, emis1, pi) and its VJP,
against the reference's own outputs (golden B) and the fp64 oracle."""

import numpy as np
import pytest

from oracle import psmc_oracle as orc

pytestmark = pytest.mark.gpu
PATTERN16 = "14*1+1*2"


def make(M, dbl=True):
    from phlash_b200.gpu import _PSMCKernelBase

    return _PSMCKernelBase(M, np.zeros((1, 16), dtype=np.int8), double_precision=dbl)


def test_matches_reference_from_dm(golden):
    import torch

    kern = make(16)
    x = torch.tensor(golden["part_x"], dtype=torch.float64, device="cuda:0")
    pp = kern.params_from_particles(x, PATTERN16, 1e-2).cpu().numpy()
    want = golden["part_pp"]
    # entries ~1e-10 are differences of O(1) cumulative products: a few ulps of 1 absolute
    np.testing.assert_allclose(pp, want, rtol=1e-8, atol=2e-15)
    assert np.all(pp[:, 0, -1] == 0) and np.all(pp[:, 2, -1] == 0) and np.all(pp[:, 3, 0] == 0)
    # float32 kernels get the same values rounded
    pp32 = make(16, dbl=False).params_from_particles(x, PATTERN16, 1e-2).cpu().numpy()
    np.testing.assert_allclose(pp32, want.astype(np.float32), rtol=1e-6, atol=2e-15)


@pytest.mark.parametrize("M", [4, 8, 32, 64])
def test_other_state_counts(M):
    import torch

    pps, xs, pattern = orc.synth_particles(M, 12, seed=M)
    got = make(M).params_from_particles(torch.tensor(xs, device="cuda:0"), pattern, 1e-2).cpu().numpy()
    np.testing.assert_allclose(got, pps, rtol=1e-8, atol=2e-15)


def test_vjp_against_finite_differences(golden):
    import torch

    rng = np.random.default_rng(0)
    kern = make(16)
    xs = golden["part_x"][1:5]
    cot = rng.normal(size=(4, 7, 16))

    def theta_of(x):
        t, cc, rho = orc.particle_to_dm(x, PATTERN16, 1e-2)
        return orc.params_from_dm(t, cc, 1e-2, rho)

    # objective sum(cot * theta): its log-space cotangent is cot * theta.  (A random cotangent on
    # log(theta) itself would weight entries like b[0] ~ 1e-10 +- 1e-16, whose logarithm is noisy at
    # the size of the finite-difference step.)
    want = np.zeros((4, 18))
    h = 1e-6
    for b in range(4):
        for p in range(18):
            up, dn = xs[b].copy(), xs[b].copy()
            up[p] += h
            dn[p] -= h
            want[b, p] = np.sum(cot[b] * (theta_of(up) - theta_of(dn))) / (2 * h)
    cot_log = np.stack([cot[b] * theta_of(xs[b]) for b in range(4)])
    got = kern.params_vjp(torch.tensor(xs, device="cuda:0"), PATTERN16, 1e-2,
                          torch.tensor(cot_log, device="cuda:0")).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-6 * np.abs(want).max())


def test_whole_chain_particle_to_gradient(golden):
    """particle -> params (device) -> fused warm-up loglik+grad (device) -> VJP (device), against
    finite differences of the oracle's composition of the same reference functions."""
    import torch

    from phlash_b200.distributed import pack_per_particle, unpack_per_particle
    from phlash_b200.gpu import _PSMCKernelBase

    chunks, inds = golden["model_chunks"], golden["model_inds"]
    kern = _PSMCKernelBase(16, chunks, double_precision=True)
    xs = golden["part_x"][:2]
    dev = torch.device("cuda:0")
    x_d = torch.tensor(xs, device=dev)
    pp = kern.params_from_particles(x_d, PATTERN16, 1e-2)
    ll, dlog = kern.evaluate_warmup_device(pp, torch.tensor(inds, device=dev), 50, True)
    ll_b, dlog_b = unpack_per_particle(pack_per_particle(ll, dlog), 16)
    grad_x = kern.params_vjp(x_d, PATTERN16, 1e-2, dlog_b.contiguous()).cpu().numpy()
    np.testing.assert_allclose(ll_b.cpu().numpy(), golden["model_l2"][:2], rtol=1e-10)

    warm, body = chunks[:, :50], chunks[:, 50:]

    def l2(x):
        t, c, rho = orc.particle_to_dm(x, PATTERN16, 1e-2)
        return orc.hmm_term(orc.params_from_dm(t, c, 1e-2, rho), warm[inds], body[inds])

    h = 1e-6
    for p in (0, 1, 2, 9, 16, 17):
        up, dn = xs[1].copy(), xs[1].copy()
        up[p] += h
        dn[p] -= h
        fd = (l2(up) - l2(dn)) / (2 * h)
        np.testing.assert_allclose(grad_x[1, p], fd, rtol=2e-5, atol=1e-6)


def test_model_hmm_term(golden):
    """phlash_b200.model.hmm_term_value_and_grad against the reference's log_density (golden F)
    with the N / S weight of mcmc.py:244."""
    import torch

    from phlash_b200 import model
    from phlash_b200.gpu import _PSMCKernelBase

    chunks, inds = golden["model_chunks"], golden["model_inds"]
    kern = _PSMCKernelBase(16, chunks, double_precision=True)
    dev = torch.device("cuda:0")
    x = torch.tensor(golden["part_x"][:3], device=dev)
    w = chunks.shape[0] / len(inds)
    val, grad = model.hmm_term_value_and_grad(kern, x, PATTERN16, 1e-2, torch.tensor(inds, device=dev), 50, weight=w)
    np.testing.assert_allclose(val.cpu().numpy(), w * golden["model_l2"], rtol=1e-10)
    assert grad.shape == (3, 18) and torch.isfinite(grad).all()
    assert model.default_minibatch_size(595, 1000) == 1 and model.default_minibatch_size(59500, 1000) == 5
