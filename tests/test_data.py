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


def _write_psmcfa(path, records, width=60, gz=False):
    import gzip

    opener = gzip.open if gz else open
    with opener(path, "wt") as fh:
        for name, seq in records:
            fh.write(f">{name} some description\n")
            for i in range(0, len(seq), width):
                fh.write(seq[i : i + width] + "\n")
            fh.write("\n")


@pytest.mark.parametrize("gz", [False, True])
def test_read_psmcfa_follows_the_reference_decoding(tmp_path, gz):
    """RawContig.from_psmcfa_iter (data.py:139-149): K -> 1, N -> -1, anything else -> 0 (case
    sensitive), one row per record, records in file order."""
    from phlash_b200.data import psmc_inputs, read_psmcfa

    rng = np.random.default_rng(0)
    records = [(f"chr{i}", "".join(rng.choice(list("TTTTTTKNnk"), size=n))) for i, n in enumerate([1, 59, 60, 61, 1000])]
    path = str(tmp_path / ("x.psmcfa.gz" if gz else "x.psmcfa"))
    _write_psmcfa(path, records, gz=gz)
    got = list(read_psmcfa(path))
    assert [n for n, _ in got] == [n for n, _ in records]
    for (_, het), (_, seq) in zip(got, records):
        arr = np.array(list(seq), dtype="c")            # the reference's own three lines
        want = (arr == b"K").astype(np.int8)
        want[arr == b"N"] = -1
        assert het.dtype == np.int8 and het.shape == (1, len(seq))
        np.testing.assert_array_equal(het[0], want)
    contigs, test = psmc_inputs([path, path], hold_out=True)        # psmc.py:22-28
    assert len(contigs) == 2 * len(records) - 1
    np.testing.assert_array_equal(test, got[0][1])
    contigs, test = psmc_inputs([path], hold_out=False)
    assert test is None and len(contigs) == len(records)


def test_read_psmcfa_rejects_headerless_input(tmp_path):
    from phlash_b200.data import read_psmcfa

    path = tmp_path / "bad.psmcfa"
    path.write_text("TTTKTT\n")
    with pytest.raises(ValueError):
        list(read_psmcfa(str(path)))


def test_raw_contig_surface_and_init_from_contigs(tmp_path):
    """RawContig (data.py:114-170) and init_mcmc_data (data.py:506-558) on .psmcfa input: N, L, size,
    get_data's window check, default chunk size ~1/5th of the shortest contig, stacked chunks."""
    from phlash_b200.data import RawContig, _chunk_het_matrix, init_mcmc_data_from_contigs

    rng = np.random.default_rng(1)
    records = [(f"chr{i}", "".join(rng.choice(list("TTTTTTTKN"), size=n))) for i, n in enumerate([30_000, 12_345])]
    path = str(tmp_path / "two.psmcfa")
    _write_psmcfa(path, records)
    contigs = list(RawContig.from_psmcfa_iter(path, window_size=100))
    assert [c.N for c in contigs] == [2, 2]
    assert [c.L for c in contigs] == [3_000_000, 1_234_500]
    assert contigs[0].size == 6_000_000
    with pytest.raises(ValueError):
        contigs[0].get_data(50)
    afs, chunks = init_mcmc_data_from_contigs(contigs, window_size=100, overlap=500)
    cs = int(0.2 * 1_234_500 / 100)                       # data.py:520-521
    want = np.concatenate([_chunk_het_matrix(c.het_matrix, 500, cs) for c in contigs], 0)
    np.testing.assert_array_equal(chunks, want)
    assert chunks.shape[1] == cs + 500
    np.testing.assert_array_equal(afs, np.full(1, 2.0))
    empty = RawContig(het_matrix=None, afs=None, window_size=100)
    assert empty.N is None and empty.L is None and empty.size is None
    with pytest.raises(ValueError):
        init_mcmc_data_from_contigs([empty], 100, 500)
