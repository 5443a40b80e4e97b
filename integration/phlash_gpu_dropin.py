"""Drop-in replacement for ``phlash/gpu.py`` (jthlab/phlash): same public names and JAX behaviour
(``PSMCKernel.loglik`` is differentiable and vmappable), with the B200 kernels of phlash_b200
underneath instead of the NVRTC-compiled reference kernels.

Install by copying this file over ``src/phlash/gpu.py`` (or by pointing ``phlash.kernel.get_kernel``
at it).  Needs jax (the reference pins jax<0.6, pyproject.toml:11); jax is
NOT installable in this repository's build image, so this module is exercised only by inspection
here - the kernel calls it makes are the ones tests/test_gpu_parity.py covers through the same
``PSMCKernel.__call__``.

Glue mirrored from the reference: custom_vjp over (log_params, index) with the kernel as a
non-differentiable argument, forward ALWAYS evaluates value and gradient, backward is g * dlog
(src/phlash/gpu.py:441-472); the callback is jax.pure_callback(..., vectorized=True).
"""

from __future__ import annotations

from functools import partial, singledispatchmethod

import jax
import jax.numpy as jnp
import numpy as np
from jax import custom_vjp

import phlash.size_history
from phlash.params import PSMCParams
from phlash_b200.gpu import CudaError, PSMCKernel as _B200Kernel  # noqa: F401  (CudaError re-exported)


class PSMCKernel:
    """Same constructor as the reference (src/phlash/gpu.py:338): M, data, double_precision, num_gpus."""

    def __init__(self, M, data, double_precision=False, num_gpus: int = None):
        self._impl = _B200Kernel(M=M, data=np.asarray(data), double_precision=double_precision, num_gpus=num_gpus)
        self.M = M
        self.double_precision = double_precision

    @property
    def float_type(self):
        return self._impl.float_type

    @singledispatchmethod
    def loglik(self, pp: PSMCParams, index: int):
        log_params = jax.tree.map(jnp.log, pp)
        return _psmc_ll(log_params, index=index, kern=self)

    @loglik.register
    def _(self, dm: phlash.size_history.DemographicModel, index):
        return self.loglik(PSMCParams.from_dm(dm), index)

    def __call__(self, pp: PSMCParams, index, grad: bool):
        """Host callback: NumPy in, NumPy out (src/phlash/gpu.py:386-423)."""
        from phlash_b200.params import PSMCParams as HostParams

        out = self._impl(HostParams(*(np.asarray(a) for a in pp)), np.asarray(index), grad)
        if not grad:
            return out
        ll, dll = out
        return ll, PSMCParams(*dll)


@partial(custom_vjp, nondiff_argnums=(2,))
def _psmc_ll(log_params: PSMCParams, index, kern) -> float:
    return _psmc_ll_helper(log_params, index=index, kern=kern, grad=False)


def _psmc_ll_fwd(log_params, index, kern):
    return _psmc_ll_helper(log_params, index=index, kern=kern, grad=True)


def _psmc_ll_helper(log_params: PSMCParams, index, kern, grad):
    params = jax.tree.map(jnp.exp, log_params)
    shape = jax.ShapeDtypeStruct(shape=(), dtype=jnp.float64)
    if grad:
        shape = (shape, PSMCParams(*[jax.ShapeDtypeStruct(shape=(params.M,), dtype=kern.float_type) for _ in params]))
    return jax.pure_callback(kern, shape, pp=params, index=index, grad=grad, vectorized=True)


def _psmc_ll_bwd(kern, df, g):
    return jax.tree.map(lambda a: g * a, df), None


_psmc_ll.defvjp(_psmc_ll_fwd, _psmc_ll_bwd)
