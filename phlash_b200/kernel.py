"""The plugin seam (reference: src/phlash/kernel.py:7-24).  The reference falls back to a pure-JAX
kernel when the CUDA module cannot be loaded; this implementation deliberately does not - errors
propagate."""

from __future__ import annotations


def get_kernel(M: int, data, double_precision: bool, num_gpus: int = None):
    from phlash_b200.gpu import PSMCKernel

    return PSMCKernel(M=M, data=data, double_precision=double_precision, num_gpus=num_gpus)
