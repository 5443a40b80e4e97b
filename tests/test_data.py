"""Product-side chunking (phlash_b200/data.py) against the reference function's own output and the
reference's test (tests/test_data.py:18-28).  CPU only."""

import numpy as np
import pytest

from phlash_b200 import data as pdata


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_chunk_matches_reference_output(golden, tag):
    ov, cs = (int(v) for v in golden[f"chunk_{tag}_geom"])
    got = pdata._chunk_het_matrix(golden[f"chunk_{tag}_in"], ov, cs)
    assert got.dtype == np.int8 and got.flags.c_contiguous
    np.testing.assert_array_equal(got, golden[f"chunk_{tag}_out"])


def test_chunk(seed):
    rng = np.random.default_rng(seed)
    h = rng.integers(0, 2, size=(1, 10_000))
    overlap, chunk_size = 123, 4_567
    ch = pdata._chunk_het_matrix(h, overlap=overlap, chunk_size=chunk_size)
    assert ch.shape == (3, overlap + chunk_size)
    b = 0
    for ch_i in ch:
        q = min(chunk_size + overlap, len(h[0, b:]))
        assert np.all(ch_i[:q] == h[0, b : b + q])
        b += chunk_size


def test_benchmark_geometry_and_split():
    het = np.zeros((2, 1_000_000), dtype=np.int8)
    chunks = pdata.init_mcmc_data([het[:1], het[1:]], overlap=500, chunk_size=10_000)
    assert chunks.shape == (2 * 96, 10_500)  # 96 chunks, 39 500 bins dropped
    warm, body = pdata.split_warmup(chunks, 500)
    assert warm.shape == (192, 500) and body.shape == (192, 10_000) and body.flags.c_contiguous
    assert pdata.default_chunk_size([10_000_000], 100) == 20_000  # data.py:520-521


def test_values_are_clipped_and_padding_is_missing():
    h = np.array([[5, -1, 0, 2, 1, 0, 0]])
    ch = pdata._chunk_het_matrix(h, overlap=2, chunk_size=3)
    assert ch.shape == (2, 5)
    np.testing.assert_array_equal(ch[0], [1, -1, 0, 1, 1])
    np.testing.assert_array_equal(ch[1], [1, 1, 0, 0, -1])
