"""Parameter container of the HMM, mirroring ``phlash.params.PSMCParams``
(reference: src/phlash/params.py:16-30): seven length-M vectors in the order
b, d, u, v, emis0, emis1, pi.  Leaves may carry leading batch axes."""

from __future__ import annotations

from typing import NamedTuple

import numpy as np


class PSMCParams(NamedTuple):
    b: np.ndarray
    d: np.ndarray
    u: np.ndarray
    v: np.ndarray
    emis0: np.ndarray
    emis1: np.ndarray
    pi: np.ndarray

    @property
    def M(self) -> int:
        "The number of discretization intervals"
        M = np.shape(self.d)[-1]
        assert all(np.shape(a)[-1] == M for a in self)
        return M

    @classmethod
    def from_block(cls, block) -> "PSMCParams":
        """Split a [..., 7, M] array into its seven rows."""
        block = np.asarray(block)
        assert block.shape[-2] == 7
        return cls(*(block[..., i, :] for i in range(7)))

    def to_block(self, dtype=None) -> np.ndarray:
        """Stack to [..., 7, M] (what the reference does with np.stack(pp, -2), gpu.py:189)."""
        return np.stack([np.asarray(a, dtype=dtype) for a in self], axis=-2)


def parse_pattern(pattern: str):
    """PSMC-style pattern string -> list of epoch widths (reference: src/phlash/util.py:8-37)."""
    try:
        epochs = []
        for tok in pattern.split("+"):
            if "*" in tok:
                k, width = map(int, tok.split("*"))
            else:
                k, width = 1, int(tok)
            epochs += [width] * k
    except Exception:
        raise ValueError("could not parse pattern")
    if len(epochs) == 0:
        raise ValueError("pattern must contain at least one epoch")
    if any(e <= 0 for e in epochs):
        raise ValueError("epochs must be positive")
    return epochs
