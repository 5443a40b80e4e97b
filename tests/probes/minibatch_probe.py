"""Small-minibatch latency (the reference's default S <= 5, mcmc.py:119-121): ours vs the
reference's own CUDA kernel on the same GPU and inputs.  B = 500 particles, L = 50 000 bins."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from benchdata import synth
from oracle import ref_cuda
from phlash_b200.data import _chunk_het_matrix, split_warmup
from phlash_b200.gpu import _PSMCKernelBase

het = synth.het_matrix(1, 3_000_000, 0)
_, data = split_warmup(_chunk_het_matrix(het, 500, 50_000), 500)
pps = synth.particles(16, 500)
ours = _PSMCKernelBase(16, data)
ref = ref_cuda.ReferenceKernel(16, data, double_precision=False)
for S in (1, 5, 32):
    inds = (np.arange(S) * 7) % data.shape[0]
    pa = np.broadcast_to(pps[:, None], (500, S, 7, 16)).astype(np.float32)
    for _ in range(3):
        ll, dlog = ours.evaluate(pa, inds, True)
        ms_ours = ours.last_kernel_ms
    ref(pa[:4], inds, grad=True)
    ll_r, dlog_r = ref(pa, inds, grad=True)
    ms_ref = ref.last_ms
    st = 500 * S * data.shape[1]
    print(json.dumps({"B": 500, "S": S, "L": int(data.shape[1]), "ours_kernel_ms": round(ms_ours, 3),
                      "reference_kernel_ms": round(ms_ref, 3), "speedup": round(ms_ref / ms_ours, 1),
                      "ours_st_per_s": st / ms_ours * 1e3, "reference_st_per_s": st / ms_ref * 1e3,
                      "ll_max_rel_diff": float(np.max(np.abs(ll - ll_r) / np.abs(ll_r)))}), flush=True)
