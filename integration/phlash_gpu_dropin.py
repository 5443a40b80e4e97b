"""Drop-in replacement for ``phlash/gpu.py`` (jthlab/phlash): same public names and JAX behaviour
(``PSMCKernel.loglik`` is differentiable and vmappable), with the B200 kernels of phlash_b200
underneath instead of the NVRTC-compiled reference kernels.

Install by copying this file over ``src/phlash/gpu.py`` (or by pointing ``phlash.kernel.get_kernel``
at it).  Needs jax (the reference pins jax<0.6, pyproject.toml:11); jax is
NOT installable in this repository's build image, so this module is exercised only by inspection
here - the kernel calls it makes are the ones tests/test_gpu_parity.py covers through the same
``PSMCKernel.__call__``.

The JAX contract it honours is the reference's (src/phlash/gpu.py:441-472), see
``_differentiable_loglik``; identity hashing keeps the object usable as a static jit argument
(src/phlash/mcmc.py:199).
"""

from __future__ import annotations

import jax
import jax.numpy as jnp
import numpy as np

import phlash.size_history
from phlash.params import PSMCParams
from phlash_b200.gpu import CudaError, PSMCKernel as _B200Kernel  # noqa: F401  (CudaError re-exported)
from phlash_b200.params import PSMCParams as _HostParams


def _differentiable_loglik(kern: "PSMCKernel"):
    """log-likelihood as a function of (log-parameters, index) whose reverse rule is the gradient the
    kernel returns.  Contract of the reference (src/phlash/gpu.py:441-472): the primal call asks the
    kernel for the value only, the forward rule of the VJP always asks for value + gradient and keeps
    the gradient as residual, the backward rule scales it by the cotangent; `index` gets no gradient."""

    def host_call(log_params, index, want_grad):
        params = jax.tree.map(jnp.exp, log_params)
        value_t = jax.ShapeDtypeStruct((), jnp.float64)
        if not want_grad:
            return jax.pure_callback(kern, value_t, pp=params, index=index, grad=False, vectorized=True)
        leaf_t = jax.ShapeDtypeStruct((kern.M,), kern.float_type)
        out_t = (value_t, PSMCParams(*([leaf_t] * len(PSMCParams._fields))))
        return jax.pure_callback(kern, out_t, pp=params, index=index, grad=True, vectorized=True)

    @jax.custom_vjp
    def loglik(log_params, index):
        return host_call(log_params, index, False)

    def forward(log_params, index):
        value, dlog = host_call(log_params, index, True)
        return value, dlog

    def backward(dlog, cotangent):
        return jax.tree.map(lambda leaf: cotangent * leaf, dlog), None

    loglik.defvjp(forward, backward)
    return loglik


class PSMCKernel:
    """Same constructor and attributes as the reference class (src/phlash/gpu.py:328-357)."""

    def __init__(self, M, data, double_precision=False, num_gpus: int = None):
        self._impl = _B200Kernel(M=M, data=np.asarray(data), double_precision=double_precision, num_gpus=num_gpus)
        self.M = M
        self.double_precision = double_precision
        self._loglik = _differentiable_loglik(self)

    @property
    def float_type(self):
        return self._impl.float_type

    def loglik(self, pp, index):
        """Differentiable, vmappable log-likelihood of data[index]; also accepts a DemographicModel
        like the reference's convenience overload (src/phlash/gpu.py:359-367)."""
        if isinstance(pp, phlash.size_history.DemographicModel):
            pp = PSMCParams.from_dm(pp)
        return self._loglik(jax.tree.map(jnp.log, pp), index)

    def __call__(self, pp, index, grad: bool):
        """The host callback: NumPy in, NumPy out."""
        out = self._impl(_HostParams(*(np.asarray(leaf) for leaf in pp)), np.asarray(index), grad)
        if not grad:
            return out
        value, dlog = out
        return value, PSMCParams(*dlog)


def make_hmm_term(full_chunks, M: int, pattern: str, theta: float, overlap: int):
    """The WHOLE HMM term of ``log_density`` (src/phlash/model.py:50-57) for all particles at once, as a
    JAX function with a custom VJP, on top of ONE library call per evaluation (``phb_hmm_term_host``:
    particles -> parameters -> fused warm-up likelihood + gradient -> VJP).

        hmm_term(x, inds, weight) -> [B]       x [B, P] float64: flattened particles (params.py:58-66)

    The host round trip is B * (2 P + 1) doubles (tens of kilobytes) instead of the reference's
    [B, S, 7, M] blocks; ``full_chunks`` is the chunk matrix BEFORE the warm-up split of mcmc.py:203.
    Use it in place of lines 50-57 of model.log_density with the vmap over particles lifted out."""
    from phlash_b200.gpu import _PSMCKernelBase

    kern = _PSMCKernelBase(M, np.asarray(full_chunks))

    def host_call(x, inds, weight, want_grad):
        value, grad = kern.hmm_term_host(np.asarray(x), pattern, theta, np.asarray(inds), overlap, float(weight), want_grad)
        return (value, grad) if want_grad else value

    def call(x, inds, weight, want_grad):
        value_t = jax.ShapeDtypeStruct(x.shape[:1], jnp.float64)
        if not want_grad:
            return jax.pure_callback(lambda *a: host_call(*a, False), value_t, x, inds, weight)
        return jax.pure_callback(lambda *a: host_call(*a, True), (value_t, jax.ShapeDtypeStruct(x.shape, jnp.float64)), x, inds, weight)

    @jax.custom_vjp
    def hmm_term(x, inds, weight):
        return call(x, inds, weight, False)

    def forward(x, inds, weight):
        value, grad = call(x, inds, weight, True)
        return value, grad

    def backward(grad, cotangent):
        return cotangent[:, None] * grad, None, None

    hmm_term.defvjp(forward, backward)
    return hmm_term
