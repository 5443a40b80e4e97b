"""A/B of the uniform-register throughput kernel (psmc_uniform.cuh) against the register-parameter kernel at
the same shape, with the two results compared:  python tools/uniform_probe.py [B] [S]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 3 and sys.argv[3] == "child":
    import torch

    from benchdata import synth
    from phlash_b200.data import _chunk_het_matrix
    from phlash_b200.gpu import _PSMCKernelBase

    B, S = int(sys.argv[1]), int(sys.argv[2])
    chunks = _chunk_het_matrix(synth.het_matrix(1, 30_000_000, 0), 500, 50_000)
    data = np.ascontiguousarray(chunks[:S, 500:])
    pps = synth.particles(16, B)
    kern = _PSMCKernelBase(16, data)
    dev = torch.device("cuda:0")
    p6 = torch.tensor(pps[:, :6], dtype=torch.float32, device=dev).contiguous()
    pi = torch.tensor(pps[:, 6], dtype=torch.float32, device=dev).contiguous()
    inds = torch.arange(S, dtype=torch.int64, device=dev)
    out = {}
    for grad in (True, False):
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ll, dlog = kern.evaluate_device(p6, pi, inds, grad)
            e1.record()
            e1.synchronize()
        kern.sync()
        ms = e0.elapsed_time(e1)
        out["grad" if grad else "fwd"] = {"ms": round(ms, 3), "st_per_s": B * S * 50_000 / (ms * 1e-3), "kernel": kern.last_kernel_name}
        if grad:
            np.save(f"/tmp/uni_ll_{os.environ.get('PHB_UNIFORM', '1')}.npy", ll.cpu().numpy())
            np.save(f"/tmp/uni_dlog_{os.environ.get('PHB_UNIFORM', '1')}.npy", dlog.cpu().numpy())
    print(json.dumps({"B": B, "S": S, "uniform": os.environ.get("PHB_UNIFORM", "1"), **out}), flush=True)
else:
    B = sys.argv[1] if len(sys.argv) > 1 else "124"
    S = sys.argv[2] if len(sys.argv) > 2 else "595"
    for mode in ("0", "1"):
        subprocess.check_call([sys.executable, __file__, B, S, "child"], env={**os.environ, "PHB_UNIFORM": mode})
    a, b = np.load("/tmp/uni_ll_0.npy"), np.load("/tmp/uni_ll_1.npy")
    ga, gb = np.load("/tmp/uni_dlog_0.npy").astype(np.float64), np.load("/tmp/uni_dlog_1.npy").astype(np.float64)
    scale = np.abs(ga).max(-1, keepdims=True)
    print(json.dumps({"ll_max_rel_diff": float(np.max(np.abs(a - b) / np.abs(a))),
                      "grad_max_diff_rel_to_row_max": float(np.max(np.abs(ga - gb) / scale)),
                      "grad_max_rel_diff_large_entries": float(np.max(np.where(np.abs(ga) > 1e-3 * scale, np.abs(ga - gb) / np.maximum(np.abs(ga), 1e-300), 0)))}))
