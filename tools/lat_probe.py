"""Whole likelihood step (phb_hmm_term_device) at small minibatches, CUDA-event timing:
    [PHB_SWEEP_T=2] python tools/lat_probe.py [M] [S,S,...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402
from phlash_b200.data import _chunk_het_matrix  # noqa: E402
from phlash_b200.gpu import _PSMCKernelBase  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
SS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 3, 5, 8]
pattern = {16: "14*1+1*2", 32: "30*1+1*2", 64: "62*1+1*2"}[M]
chunks = _chunk_het_matrix(synth.het_matrix(1, 3_000_000, 0), 500, 50_000)[:50]
kern = _PSMCKernelBase(M, chunks, overlap=500)
xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:500]
x = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
for S in SS:
    inds = torch.arange(S, device="cuda:0") * (50 // S)
    for _ in range(3):
        v, g = kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        v, g = kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=1.0)
    e1.record()
    e1.synchronize()
    print(json.dumps({"M": M, "S": S, "ms": round(e0.elapsed_time(e1) / 10, 3), "sweep_T": os.environ.get("PHB_SWEEP_T", "default"),
                      "value0": float(v[0]), "gsum": float(g.abs().sum()), "kernel": kern.last_kernel_name}), flush=True)
