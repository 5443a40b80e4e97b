"""Whole likelihood step at small minibatches on several GPUs (one process per GPU, NCCL):
    [PHB_PIT_SEGMENTS=192] torchrun --nproc-per-node N tools/dist_probe.py [S,S,...]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402
from phlash_b200 import model  # noqa: E402
from phlash_b200.data import _chunk_het_matrix  # noqa: E402
from phlash_b200.gpu import _PSMCKernelBase  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
SS = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 5]
chunks = _chunk_het_matrix(synth.het_matrix(1, 3_000_000, 0), 500, 50_000)[:50]
kern = _PSMCKernelBase(16, chunks, device=local, overlap=500)
xs = np.load(os.path.join(ROOT, "benchdata", "particles_M16.npz"))["xs"][:500]
x = torch.tensor(xs, dtype=torch.float64, device=dev)
for S in SS:
    inds = torch.arange(S, device=dev) * (50 // S)
    for _ in range(3):
        v, g = model.hmm_term_value_and_grad(kern, x, "14*1+1*2", 1e-2, inds, 500, rank=rank, world=world)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        v, g = model.hmm_term_value_and_grad(kern, x, "14*1+1*2", 1e-2, inds, 500, rank=rank, world=world)
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "S": S, "ms": round(float(t), 3), "segments": os.environ.get("PHB_PIT_SEGMENTS", "auto"),
                          "plan": kern.sharded_plan(500, S, 500, world), "value0": float(v[0]), "gsum": float(g.abs().sum())}), flush=True)
dist.destroy_process_group()
