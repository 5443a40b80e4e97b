"""Quick device-resident throughput probe of the loglik(+grad) kernel for several lane layouts.
Usage: python tools/perf_probe.py [M] [B] [n_chunks] [L] [T,T,...] [dbl]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth
from phlash_b200.gpu import _PSMCKernelBase

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
NCH = int(sys.argv[3]) if len(sys.argv) > 3 else 595
L = int(sys.argv[4]) if len(sys.argv) > 4 else 50_000
TS = [int(t) for t in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
DBL = len(sys.argv) > 6 and sys.argv[6] == "dbl"

rng = np.random.default_rng(0)
het = synth.het_matrix(1, 200_000, 0)
base = het[0]
reps = -(-(NCH * L) // base.size)
data = np.tile(base, reps)[: NCH * L].reshape(NCH, L).copy()
data[:, 0] = np.where(data[:, 0] < 0, 0, data[:, 0])
pps = synth.particles(M, B)  # committed fixtures: M = 16, 32, 64
kern = _PSMCKernelBase(M, data, double_precision=DBL)
dt = torch.float64 if DBL else torch.float32
dev = torch.device("cuda:0")
p6 = torch.tensor(pps[:, :6], dtype=dt, device=dev).contiguous()
pi = torch.tensor(pps[:, 6], dtype=dt, device=dev).contiguous()
inds = torch.arange(NCH, dtype=torch.int64, device=dev)
for T in TS:
    kern.set_threads_per_pair(T)
    for grad in (True, False):
        for it in range(3):
            ll, dlog = kern.evaluate_device(p6, pi, inds, grad)
            kern.sync()
            ms = kern.last_kernel_ms
        print(json.dumps({"M": M, "B": B, "S": NCH, "L": L, "T": T, "grad": grad, "dbl": DBL, "ms": round(ms, 3),
                          "st_per_s": B * NCH * L / (ms * 1e-3), "ll_mean": float(ll.mean())}), flush=True)
