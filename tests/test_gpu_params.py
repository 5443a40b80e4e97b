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


@pytest.mark.parametrize("dbl", [True, False])
def test_whole_term_entry_equals_the_composition(golden, dbl):
    """phb_hmm_term_device = params_from_particles -> fused warm-up loglik+grad -> sum over the
    minibatch -> params_vjp; also the split form around the all-reduce (sums / finish)."""
    import torch

    from phlash_b200.gpu import _PSMCKernelBase

    chunks, inds_np = golden["model_chunks"], golden["model_inds"]
    kern = _PSMCKernelBase(16, chunks, double_precision=dbl)
    dev = torch.device("cuda:0")
    x = torch.tensor(golden["part_x"][:5], device=dev)
    inds = torch.tensor(inds_np, device=dev)
    w = 3.5
    val, grad = kern.hmm_term(x, PATTERN16, 1e-2, inds, 50, weight=w)
    # the composition, step by step
    params7 = kern.params_from_particles(x, PATTERN16, 1e-2)
    ll, dlog = kern.evaluate_warmup_device(params7, inds, 50, True)
    cot = dlog.double().sum(dim=1)
    want_val = w * ll.sum(dim=1)
    want_grad = w * kern.params_vjp(x, PATTERN16, 1e-2, cot.to(params7.dtype).contiguous())
    torch.testing.assert_close(val, want_val, rtol=1e-13, atol=0)
    # (the library keeps the cotangent in double; the step-by-step path rounds it to the kernel's type)
    torch.testing.assert_close(grad, want_grad, rtol=1e-12 if dbl else 2e-6, atol=1e-9 if dbl else 1e-4)
    # forward only
    val_f, none = kern.hmm_term(x, PATTERN16, 1e-2, inds, 50, weight=w, grad=False)
    assert none is None
    torch.testing.assert_close(val_f, val, rtol=1e-6 if not dbl else 1e-13, atol=0)
    # split around the collective: two "ranks", one of them possibly without work
    for cut in (0, 2, len(inds_np)):
        sums = kern.hmm_term_sums(x, PATTERN16, 1e-2, inds[:cut].contiguous(), 50)
        sums = sums + kern.hmm_term_sums(x, PATTERN16, 1e-2, inds[cut:].contiguous(), 50)
        v2, g2 = kern.hmm_term_finish(x, PATTERN16, 1e-2, sums, weight=w)
        torch.testing.assert_close(v2, val, rtol=1e-13, atol=0)
        torch.testing.assert_close(g2, grad, rtol=1e-11, atol=1e-9)
    empty = kern.hmm_term_sums(x, PATTERN16, 1e-2, inds[:0].contiguous(), 50)
    assert torch.count_nonzero(empty) == 0


def test_whole_term_gradient_by_finite_differences(golden):
    import torch

    from phlash_b200.gpu import _PSMCKernelBase

    chunks, inds_np = golden["model_chunks"], golden["model_inds"]
    kern = _PSMCKernelBase(16, chunks, double_precision=True)
    dev = torch.device("cuda:0")
    x = torch.tensor(golden["part_x"][:2], device=dev)
    inds = torch.tensor(inds_np, device=dev)
    val, grad = kern.hmm_term(x, PATTERN16, 1e-2, inds, 50, weight=2.0)
    h = 1e-5
    for p in (0, 1, 5, 16, 17):
        xp, xm = x.clone(), x.clone()
        xp[:, p] += h
        xm[:, p] -= h
        fd = (kern.hmm_term(xp, PATTERN16, 1e-2, inds, 50, 2.0, grad=False)[0]
              - kern.hmm_term(xm, PATTERN16, 1e-2, inds, 50, 2.0, grad=False)[0]) / (2 * h)
        torch.testing.assert_close(grad[:, p], fd, rtol=2e-5, atol=1e-5)


def test_whole_term_host_entry(golden):
    """phb_hmm_term_host (NumPy buffers, blocking) equals the device entry; input checks raise."""
    import torch

    from phlash_b200.gpu import _PSMCKernelBase

    chunks, inds_np = golden["model_chunks"], golden["model_inds"]
    kern = _PSMCKernelBase(16, chunks)
    x_np = golden["part_x"][:5]
    val_h, grad_h = kern.hmm_term_host(x_np, PATTERN16, 1e-2, inds_np, 50, weight=2.5)
    dev = torch.device("cuda:0")
    val_d, grad_d = kern.hmm_term(torch.tensor(x_np, device=dev), PATTERN16, 1e-2, torch.tensor(inds_np, device=dev), 50, weight=2.5)
    np.testing.assert_array_equal(val_h, val_d.cpu().numpy())
    np.testing.assert_array_equal(grad_h, grad_d.cpu().numpy())
    val_f, none = kern.hmm_term_host(x_np, PATTERN16, 1e-2, inds_np, 50, weight=2.5, grad=False)
    assert none is None
    np.testing.assert_allclose(val_f, val_h, rtol=1e-6)
    with pytest.raises(AssertionError):
        kern.hmm_term_host(x_np, PATTERN16, 1e-2, np.array([len(chunks)]), 50)
    bad = x_np.copy()
    bad[0, 0] = np.nan
    with pytest.raises(AssertionError):
        kern.hmm_term_host(bad, PATTERN16, 1e-2, inds_np, 50)
