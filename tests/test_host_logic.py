"""Host-side logic of the PSMCKernel mirror that needs no GPU."""

import numpy as np
import pytest
import torch

from phlash_b200.gpu import CudaError, _normalise_call
from phlash_b200.params import PSMCParams


def block(*lead, m=16, seed=0):
    return np.random.default_rng(seed).uniform(0.1, 0.9, size=(*lead, 7, m))


def test_psmcparams_roundtrip():
    blk = block(3, 2)
    pp = PSMCParams.from_block(blk)
    assert pp.M == 16 and pp.b.shape == (3, 2, 16)
    np.testing.assert_array_equal(pp.to_block(), blk)
    assert PSMCParams._fields == ("b", "d", "u", "v", "emis0", "emis1", "pi")  # params.py:16-23


@pytest.mark.parametrize(
    "pshape,index,expect",
    [
        ((), 3, (1, 1, True, True)),            # pp [M], scalar index     (gpu.py:192-195, 203-207)
        ((), [3, 1, 3], (1, 3, True, False)),   # pp [M], index [S]: broadcast over S
        ((3,), [0, 1, 2], (1, 3, True, False)),  # pp [S, M]                (gpu.py:208-211)
        ((4, 3), [0, 1, 2], (4, 3, False, False)),  # pp [B, S, M]
    ],
)
def test_argument_shapes_follow_the_reference(pshape, index, expect):
    pp = PSMCParams.from_block(block(*pshape))
    pa, inds, added_b, added_s = _normalise_call(pp, index, 16)
    b, s, eb, es = expect
    assert pa.shape == (b, s, 7, 16) and inds.shape == (s,)
    assert (added_b, added_s) == (eb, es)


def test_shape_mismatch_is_an_assertion_error():
    pp = PSMCParams.from_block(block(2))
    with pytest.raises(AssertionError):
        _normalise_call(pp, [0, 1, 2], 16)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_fallback_without_a_gpu():
    """kernel.get_kernel must raise, not fall back (the reference falls back to pure JAX,
    kernel.py:14-24; the north star forbids that)."""
    from phlash_b200.kernel import get_kernel

    data = np.zeros((2, 64), dtype=np.int8)
    with pytest.raises((CudaError, RuntimeError)):
        get_kernel(M=16, data=data, double_precision=False)


def test_fit_dispatch_rules():
    """mcmc.py:116-139, 240-247: minibatch size, down-sampling bound, minibatch weight."""
    import numpy as np

    from phlash_b200 import model

    assert model.default_minibatch_size(595, 1000) == 1       # BASELINE config 2: one genome
    assert model.default_minibatch_size(5950, 1000) == 5      # config 3
    assert model.default_minibatch_size(59_500, 1000) == 5    # config 4, before down-sampling
    assert model.default_minibatch_size(10, 1000) == 1
    rng = np.random.default_rng(0)
    chunks = np.arange(59_500 * 3, dtype=np.int32).reshape(59_500, 3)
    kept = model.downsample_chunks(chunks, 5, 1000, rng)       # 5 * S * niter = 25 000 rows
    assert kept.shape == (25_000, 3)
    assert len(np.unique(kept[:, 0])) == 25_000                # without replacement, whole rows
    assert np.all(kept[:, 1] == kept[:, 0] + 1)
    same = model.downsample_chunks(chunks[:20_000], 5, 1000, rng)
    assert same.shape == (20_000, 3)
    assert model.minibatch_weight(595, 1) == 595.0 and model.minibatch_weight(25_000, 5) == 5000.0
    inds = model.sample_minibatch(np.random.default_rng(1), 595, 5)
    assert inds.shape == (5,) and inds.min() >= 0 and inds.max() < 595
