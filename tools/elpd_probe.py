"""Forward-only evaluation of few, long pairs (the reference's ELPD shape, mcmc.py:213-238: B particles
x N_test un-chunked contigs), for the lane layouts the dispatcher can choose from.
Usage: python tools/elpd_probe.py [M] [B] [L] [N_test,N_test,...] [T,T,...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchdata import synth
from phlash_b200.gpu import _PSMCKernelBase

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 500
L = int(sys.argv[3]) if len(sys.argv) > 3 else 2_500_000
NS = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
TS = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]

het = synth.het_matrix(max(NS), L, seed=11)
pps = synth.particles(M, B)
dev = torch.device("cuda:0")
p6 = torch.tensor(pps[:, :6], dtype=torch.float32, device=dev).contiguous()
pi = torch.tensor(pps[:, 6], dtype=torch.float32, device=dev).contiguous()
kern = _PSMCKernelBase(M, het)
for n in NS:
    inds = torch.arange(n, dtype=torch.int64, device=dev)
    for T in TS:
        kern.set_threads_per_pair(T)
        for _ in range(3):
            ll, _ = kern.evaluate_device(p6, pi, inds, False)
            kern.sync()
        ms = kern.last_kernel_ms
        print(json.dumps({"M": M, "B": B, "N_test": n, "L": L, "T": T, "ms": round(ms, 3), "ns_per_site": round(ms * 1e6 / L, 1),
                          "st_per_s": B * n * L / (ms * 1e-3), "kernel": kern.last_kernel_name}), flush=True)
