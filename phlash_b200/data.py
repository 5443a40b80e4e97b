"""Host-side data layer on the hot path: cutting binned contigs into overlapping chunks
(reference: src/phlash/data.py:22-24, 37-61, 102-112, 506-558).  The reference does this once,
in NumPy on the host; so does this mirror.  Of the file readers only the .psmcfa one is mirrored
(data.py:122-149, psmc.py:8-29: plain text, no third-party parser needed); VCF and tree-sequence
input need pysam / tskit and stay with the reference."""

from __future__ import annotations

import gzip
from dataclasses import dataclass
from typing import Iterator, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np


class ChunkedContig(NamedTuple):
    chunks: np.ndarray  # int8 [N, overlap + chunk_size]
    afs: Optional[np.ndarray]


def _chunk_het_matrix(het_matrix: np.ndarray, overlap: int, chunk_size: int) -> np.ndarray:
    """Windows of ``overlap + chunk_size`` bins starting every ``chunk_size`` bins, rows padded
    with -1 (reference: data.py:37-61; same quirks: ``ceil(L / W)`` windows per row, so the tail of
    a long row is not covered, and window 0's first ``overlap`` bins only ever serve as warm-up)."""
    data = np.ascontiguousarray(np.clip(het_matrix, -1, 1).astype(np.int8))
    assert data.ndim == 2
    n, length = data.shape
    width = chunk_size + overlap
    n_chunks = -(-length // width)
    reach = (n_chunks - 1) * chunk_size + width  # last bin (exclusive) any window touches
    padded = np.full((n, max(reach, length)), -1, dtype=np.int8)
    padded[:, :length] = data
    windows = np.lib.stride_tricks.sliding_window_view(padded, width, axis=1)[:, :: chunk_size][:, :n_chunks]
    return np.ascontiguousarray(windows).reshape(n * n_chunks, width)


def default_chunk_size(lengths_in_bp: Sequence[int], window_size: int) -> int:
    """~1/5th of the shortest contig, in bins (reference: data.py:520-521)."""
    return int(min(0.2 * length / window_size for length in lengths_in_bp if length))


def init_mcmc_data(het_matrices: Sequence[np.ndarray], overlap: int, chunk_size: int) -> np.ndarray:
    """Chunk every contig with the same geometry and stack the chunks
    (reference: data.py:506-558 without the process pool and the AFS bookkeeping)."""
    chunks = [_chunk_het_matrix(h, overlap, chunk_size) for h in het_matrices]
    assert len({c.shape[-1] for c in chunks}) == 1
    return np.concatenate(chunks, 0)


def split_warmup(chunks: np.ndarray, overlap: int):
    """warmup_chunks, data_chunks = np.split(chunks, [overlap], axis=1) (reference: mcmc.py:203)."""
    return chunks[:, :overlap], np.ascontiguousarray(chunks[:, overlap:])


_PSMCFA_CODE = np.zeros(256, dtype=np.int8)  # every letter is "homozygous" ...
_PSMCFA_CODE[ord("K")] = 1                   # ... except K = at least one heterozygote in the window
_PSMCFA_CODE[ord("N")] = -1                  # and N = missing (reference: data.py:146-147, case sensitive)


def read_psmcfa(path: str) -> Iterator[Tuple[str, np.ndarray]]:
    """(name, het_matrix int8 [1, L]) for every record of a PSMC FASTA file, one entry per window of
    the `fq2psmcfa -s` size (reference: RawContig.from_psmcfa_iter, data.py:122-149, which reads the
    records with pysam.FastxFile; this is a plain FASTA reader, gzip transparently)."""
    opener = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    name, parts = None, []

    def record():
        seq = np.frombuffer(b"".join(parts), dtype=np.uint8)
        return name, _PSMCFA_CODE[seq][None, :]

    with opener(path, "rb") as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(b">"):
                if name is not None:
                    yield record()
                name, parts = line[1:].split()[0].decode() if len(line) > 1 else "", []
            else:
                if name is None:
                    raise ValueError(f"{path}: sequence data before the first '>' header")
                parts.append(line)
    if name is not None:
        yield record()


def psmc_inputs(psmcfa_files: Sequence[str], hold_out: bool = True) -> Tuple[List[np.ndarray], Optional[np.ndarray]]:
    """het matrices of all contigs of the files, in order, and the held-out test contig: the first one,
    when there is more than one (reference: psmc.psmc, psmc.py:22-28)."""
    contigs = [het for f in psmcfa_files for _, het in read_psmcfa(f)]
    test = contigs.pop(0) if hold_out and len(contigs) > 1 else None
    return contigs, test


@dataclass(frozen=True)
class RawContig:
    """A contig with a pre-computed het matrix and AFS: the reference's `RawContig` (data.py:114-170)
    with the part of the `Contig` surface the hot path uses (`N`, `L`, `size`, `get_data`,
    `to_chunked`, data.py:63-112).  The other Contig kinds of the reference (VCF, tree sequence) need
    pysam / tskit and convert to this one with `Contig.to_raw`."""

    het_matrix: Optional[np.ndarray]  # int8 [n_diploids, L / window_size]
    afs: Optional[np.ndarray]
    window_size: int

    @classmethod
    def from_psmcfa_iter(cls, psmcfa_path: str, window_size: int) -> Iterator["RawContig"]:
        """data.py:122-149 (afs = ones(1) for a single diploid)."""
        for _, het in read_psmcfa(psmcfa_path):
            yield cls(het_matrix=het, afs=np.ones(1), window_size=window_size)

    @property
    def N(self) -> Optional[int]:
        """number of ploids: twice the number of rows (data.py:150-156)"""
        return None if self.het_matrix is None else 2 * self.het_matrix.shape[0]

    @property
    def L(self) -> Optional[int]:
        """sequence length in base pairs (data.py:158-162)"""
        return None if self.het_matrix is None else self.het_matrix.shape[1] * self.window_size

    @property
    def size(self) -> Optional[int]:
        return None if self.L is None or self.N is None else self.L * self.N

    def get_data(self, window_size: int) -> dict:
        if window_size != self.window_size:
            raise ValueError(
                f"This contig was created with a window size of {self.window_size} but you requested {window_size}"
            )
        return {"het_matrix": self.het_matrix, "afs": self.afs, "window_size": self.window_size}

    def to_chunked(self, overlap: int, chunk_size: int, window_size: int = 100) -> ChunkedContig:
        d = self.get_data(window_size)
        ch = None if d["het_matrix"] is None else _chunk_het_matrix(d["het_matrix"], overlap, chunk_size)
        return ChunkedContig(chunks=ch, afs=d["afs"])


def init_mcmc_data_from_contigs(data: Sequence[RawContig], window_size: int, overlap: int, chunk_size: Optional[int] = None):
    """(summed AFS, stacked chunks) of a list of contigs - the reference's init_mcmc_data (data.py:506-558)
    without the process pool: default chunk size ~1/5th of the shortest contig, every contig chunked with
    the same geometry, AFS summed, chunks concatenated."""
    if all(ds.L is None for ds in data):
        raise ValueError("None of the contigs have a length")
    if chunk_size is None:
        chunk_size = default_chunk_size([ds.L for ds in data if ds.L], window_size)
    parts = [ds.to_chunked(overlap=overlap, chunk_size=chunk_size, window_size=window_size) for ds in data]
    afss = [p.afs for p in parts if p.afs is not None]
    chunks = [p.chunks for p in parts if p.chunks is not None]
    assert all(a.ndim == 1 for a in afss) and len({a.shape for a in afss}) <= 1
    assert len({ch.shape[-1] for ch in chunks}) == 1 and all(ch.ndim == 2 for ch in chunks)
    return (np.sum(afss, 0) if afss else None), np.concatenate(chunks, 0)
