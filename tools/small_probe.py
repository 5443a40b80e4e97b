"""Gradient-kernel time at small minibatches (B = 500, L = 50 000): python tools/small_probe.py [M] [S,S,...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth
from phlash_b200.data import _chunk_het_matrix, split_warmup
from phlash_b200.gpu import _PSMCKernelBase

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
SS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 5]
het = synth.het_matrix(1, 3_000_000, 0)
_, data = split_warmup(_chunk_het_matrix(het, 500, 50_000), 500)
pps = synth.particles(M, 500)
kern = _PSMCKernelBase(M, data)
for S in SS:
    inds = (np.arange(S) * 7) % data.shape[0]
    pa = np.broadcast_to(pps[:, None], (500, S, 7, M)).astype(np.float32)
    times = []
    for _ in range(9):
        ll, dlog = kern.evaluate(pa, inds, True)
        times.append(kern.last_kernel_ms)
    ms = float(np.median(times[2:]))
    print(json.dumps({"M": M, "S": S, "ms": round(ms, 3), "min_ms": round(min(times), 3), "max_ms": round(max(times[2:]), 3), "kernel": kern.last_kernel_name, "ll0": float(ll[0, 0]),
                      "lib": os.environ.get("PHB_LIBRARY", "default")}), flush=True)
