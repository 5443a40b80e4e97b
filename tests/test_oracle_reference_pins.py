"""The reference's own known-answer and property tests for the hot path, re-expressed on the
oracle without jax.  Each test names the reference test it mirrors."""

import numpy as np
import scipy.linalg
import scipy.stats

from oracle import psmc_oracle as orc


def _Q(r, c, n):
    # tests/test_transition.py:11-18
    return np.array([[-r, r, 0.0], [1.0 * c, -(n * c), (n - 1) * c], [0.0, 0.0, -0.0]])


def test_expq(seed):
    """tests/test_transition.py:21-28"""
    rng = np.random.default_rng(seed)
    for sigma in 1e-2, 1, 10, 100:
        r, c = sigma**2 * rng.chisquare(1, (2,))
        for n in [2, 10, 20, 50, 100]:
            np.testing.assert_allclose(scipy.linalg.expm(_Q(r, c, n)), orc.expQ(r, c, n), rtol=1e-4)


def test_transition():
    """tests/test_transition.py:31-35"""
    t, c, rho = orc.default_dm(16, 1e-2, 1e-2)
    for n in 2, 5, 10, 50:
        m = orc.transition_matrix(t, c, rho, n)
        assert np.all(m >= 0.0)
        np.testing.assert_allclose(m.sum(1), 1.0)


def test_matvec(seed):
    """tests/test_hmm.py:10-19"""
    rng = np.random.default_rng(seed)
    t, c, rho = orc.default_dm(16, 1e-2, 1e-2)
    a = orc.transition_matrix(t, c, rho)
    v = rng.uniform(size=16)
    v /= v.sum()
    pp = orc.params_from_dm(t, c, 1e-2, rho)
    np.testing.assert_allclose(v @ a, orc.matvec_smc(v, pp))
    np.testing.assert_allclose(orc.dense_from_pp(pp), a, rtol=1e-9, atol=1e-20)


def test_pi():
    """tests/test_size_history.py:30-40"""
    s = orc.surv(np.array([0.0, 1.0, 2.0, 3.0]), np.ones(4))
    np.testing.assert_allclose(s[0], np.exp(-1))
    q = scipy.stats.expon.ppf([0.1, 0.2, 0.3])
    np.testing.assert_allclose(orc.surv(np.concatenate([[0.0], q]), np.ones(4)), [0.9, 0.8, 0.7, 0.0])
    q = scipy.stats.expon.ppf([0.25, 0.5, 0.75])
    np.testing.assert_allclose(orc.stationary_pi(np.concatenate([[0.0], q]), np.ones(4)), 0.25)


def test_expm1inv(seed):
    """tests/test_size_history.py:125-127"""
    y = np.random.default_rng(seed).normal(size=100) * 10
    np.testing.assert_allclose(1.0 / np.expm1(y), orc.expm1inv(y))


def test_chunk(seed):
    """tests/test_data.py:18-28"""
    rng = np.random.default_rng(seed)
    h = rng.integers(0, 2, size=(1, 10_000))
    overlap, chunk_size = 123, 4_567
    ch = orc.chunk_het_matrix(h, overlap=overlap, chunk_size=chunk_size)
    assert ch.shape == (3, overlap + chunk_size)
    b = 0
    for ch_i in ch:
        q = min(chunk_size + overlap, len(h[0, b:]))
        assert np.all(ch_i[:q] == h[0, b : b + q])
        assert np.all(ch_i[q:] == -1)
        b += chunk_size


def test_chunk_geometry_at_benchmark_scale():
    """30 M bins, chunk 50 000, overlap 500 -> 595 chunks and the last
    249 500 bins never covered (checked on the index arithmetic, not on 30 MB of data)."""
    length, cs, ov = 30_000_000, 50_000, 500
    width = cs + ov
    n_chunks = -(-length // width)
    assert n_chunks == 595
    assert length - (n_chunks * cs + ov) == 249_500
    small = orc.chunk_het_matrix(np.zeros((1, 1_000_000), dtype=np.int8), 500, 10_000)
    assert small.shape == (96, 10_500)


def test_missing_data_is_neutral_for_the_likelihood():
    """A run of missing observations multiplies by A only; rows of A sum to one, so the
    log-likelihood of an all-missing sequence is zero (hmm.py:70-71 semantics)."""
    t, c, rho = orc.default_dm(16, 1e-2, 1e-2)
    pp = orc.params_from_dm(t, c, 1e-2, rho)
    _, ll = orc.psmc_ll(pp, np.full(300, -1, dtype=np.int8))
    assert abs(ll) < 1e-10
