"""Per-kernel durations of one likelihood step at small minibatches for every lane layout of the boundary sweeps
(CUPTI through torch.profiler; the layout knobs are read when a kernel object is created):
    python tools/sweep_probe.py"""
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

CASES = [  # M, S, TF, TB, LL   (TF = 0: the library's own choice)
    (16, 3, 0, 0, 1), (16, 3, 4, 4, 0),
    (16, 5, 0, 0, 1), (16, 5, 4, 4, 1), (16, 5, 4, 4, 0), (16, 5, 2, 2, 1), (16, 5, 1, 1, 0),
    (16, 8, 0, 0, 1), (16, 8, 4, 4, 0), (16, 8, 4, 2, 1), (16, 8, 1, 1, 0),
    (16, 12, 0, 0, 1), (16, 12, 4, 4, 0), (16, 12, 1, 1, 0),
    (32, 5, 0, 0, 1), (32, 5, 8, 8, 1), (32, 5, 8, 4, 1), (32, 5, 4, 4, 1), (32, 5, 4, 2, 1), (32, 5, 2, 2, 1),
    (32, 2, 0, 0, 1), (32, 2, 8, 8, 1),
]
if len(sys.argv) > 1:
    CASES = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1:]]
chunk_cache = {}
for M, S, TF, TB, LL in CASES:
    pattern = {16: "14*1+1*2", 32: "30*1+1*2", 64: "62*1+1*2"}[M]
    if "chunks" not in chunk_cache:
        chunk_cache["chunks"] = _chunk_het_matrix(synth.het_matrix(1, 3_000_000, 0), 500, 50_000)[:50]
    for key, val in (("PHB_SWEEP_TF", TF), ("PHB_SWEEP_TB", TB)):
        if val:
            os.environ[key] = str(val)
        else:
            os.environ.pop(key, None)
    os.environ["PHB_SWEEP_LL"] = str(LL)
    kern = _PSMCKernelBase(M, chunk_cache["chunks"], overlap=500)
    xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:500]
    x = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
    inds = torch.arange(S, device="cuda:0") * (50 // S)
    for _ in range(2):
        v, g = kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        v, g = kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=1.0)
    e1.record()
    e1.synchronize()
    step_ms = e0.elapsed_time(e1) / 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=1.0)
        torch.cuda.synchronize()
    per = {}
    for ev in prof.key_averages():
        name = ev.key
        for tag in ("boundary_sweep", "psmc_loglik_kernel", "storeall", "transfer_rows", "chain_boundaries", "params_warp_kernel<false>", "params_warp_kernel<true>", "params_warp"):
            if tag in name:
                per[tag] = round(per.get(tag, 0.0) + ev.device_time_total / 3 / 1000.0, 3)
    print(json.dumps({"M": M, "S": S, "forced": [TF, TB, LL], "step_ms": round(step_ms, 3), "kernels_ms": per, "value0": float(v[0]),
                      "gsum": float(g.abs().sum())}), flush=True)
    del kern
