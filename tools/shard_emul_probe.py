"""What ONE process of a time-sharded step launches, timed on one GPU (no collectives): begin + end of rank 0 (which
also holds the log-likelihood and the warm-up term) for an emulated `world`, per-kernel times through CUPTI.
    [PHB_PIT_SEGMENTS=...] python tools/shard_emul_probe.py [world] [S,S,...]
The gather buffer holds rank 0's product operator only; the other slots are filled by running begin() for every rank
once (as the emulation test does), so that end() chains real operators."""
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402
from phlash_b200.data import _chunk_het_matrix  # noqa: E402
from phlash_b200.gpu import _PSMCKernelBase  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 5]
chunks = _chunk_het_matrix(synth.het_matrix(1, 3_000_000, 0), 500, 50_000)[:50]
kern = _PSMCKernelBase(16, chunks, overlap=500)
xs = np.load(os.path.join(ROOT, "benchdata", "particles_M16.npz"))["xs"][:500]
x = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
for S in SS:
    inds = torch.arange(S, device="cuda:0") * (50 // S)
    n_seg, slot = kern.sharded_plan(500, S, 500, world)
    if n_seg == 0:
        print(json.dumps({"world": world, "S": S, "plan": "does not apply"}))
        continue
    gather = torch.zeros(world * slot, dtype=torch.uint8, device="cuda:0")
    for r in range(world):
        kern.sharded_begin(x, "14*1+1*2", 1e-2, inds, 500, r, world, gather)

    def rank0():
        kern.sharded_begin(x, "14*1+1*2", 1e-2, inds, 500, 0, world, gather)
        return kern.sharded_end(inds, 500, 500, 0, world, gather)

    for _ in range(3):
        rank0()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sums = rank0()
    e1.record()
    e1.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            rank0()
        torch.cuda.synchronize()
    per = {}
    for ev in prof.key_averages():
        name = ev.key.split("(")[0].replace("void ", "").replace("phb::", "")
        per[name[:48]] = round(ev.device_time_total / 3 / 1000.0, 3)
    print(json.dumps({"world": world, "S": S, "segments": n_seg, "env": os.environ.get("PHB_PIT_SEGMENTS", "auto"),
                      "rank0_ms": round(e0.elapsed_time(e1) / 10, 3), "kernels_ms": per, "sum0": float(sums[0, 0])}), flush=True)
