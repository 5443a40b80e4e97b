"""The opt-in thread-per-pair kernels on the rescaled recursion (psmc_sform.cuh, PHB_SFORM=1; not the default: fewer
instructions but measured slower, see the header of that file) against the fp64 oracle.  The knob is read when a
kernel object is created, so the check runs in a child process."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = r"""
import numpy as np
from oracle import c_oracle, psmc_oracle as orc
from phlash_b200.gpu import _PSMCKernelBase

het = orc.synth_het_matrix(1, 130_000, seed=4)
data = orc.chunk_het_matrix(het, 500, 20_000)[:5, 500:].copy()
pps, _, _ = orc.synth_particles(16, 40, seed=2)
pa = np.broadcast_to(pps[:, None], (40, 5, 7, 16)).astype(np.float32)
inds = np.array([4, 0, 2, 2, 1])
kern = _PSMCKernelBase(16, data)
kern.set_threads_per_pair(1)
ll, dlog = kern.evaluate(pa, inds, True)
assert "psmc_sform_kernel<float,MT=16,T=1,K=8,grad" in kern.last_kernel_name, kern.last_kernel_name
ref_ll, ref = c_oracle.loglik_batch(data, np.tile(inds, 40), pa.reshape(-1, 7, 16).astype(np.float64), grad=True)
ref_ll, ref = ref_ll.reshape(40, 5), ref.reshape(40, 5, 7, 16)
np.testing.assert_allclose(ll, ref_ll, rtol=1e-5)
scale = np.abs(ref).max(-1, keepdims=True)
assert np.all(np.abs(dlog - ref) <= 1e-4 * np.abs(ref) + 1e-7 * scale), np.max(np.abs(dlog - ref) / (np.abs(ref) + 1e-3 * scale))
assert np.all(dlog[:, :, 0, -1] == 0) and np.all(dlog[:, :, 2, -1] == 0) and np.all(dlog[:, :, 3, 0] == 0)
ll_f = kern.evaluate(pa, inds, False)
assert "psmc_sform_kernel" in kern.last_kernel_name and "fwd" in kern.last_kernel_name
np.testing.assert_allclose(ll_f, ref_ll, rtol=1e-5)
# segment mode (the gradient passes of the parallel-in-time paths)
kern2 = _PSMCKernelBase(16, data)
ll2, dlog2 = kern2.evaluate(pa[:, :1], inds[:1], True)
assert "segments" in kern2.last_kernel_name, kern2.last_kernel_name
np.testing.assert_allclose(ll2, ref_ll[:, :1], rtol=1e-5)
assert np.all(np.abs(dlog2 - ref[:, :1]) <= 1e-4 * np.abs(ref[:, :1]) + 1e-7 * scale[:, :1])
# outside the kernel's domain (emis0 == 0): flagged, reported by sync
bad = pa.copy()
bad[0, :, 4, 3] = 0.0
try:
    kern.evaluate(bad, inds, True)
    raise SystemExit("domain violation not reported")
except AssertionError as e:
    assert "domain" in str(e)
print("SFORM-OK")
"""


def test_rescaled_kernels_match_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {**os.environ, "PHB_SFORM": "1", "PYTHONPATH": root}
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert "SFORM-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
