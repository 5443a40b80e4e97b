"""One whole likelihood step (phb_hmm_term_device) at a small minibatch, for `ncu --metrics gpu__time_duration.sum`:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/step_probe.py 1
prints nothing but the last kernel name; the launch list is the result."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchdata import synth  # noqa: E402
from phlash_b200.data import _chunk_het_matrix  # noqa: E402
from phlash_b200.gpu import _PSMCKernelBase  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pattern = {16: "14*1+1*2", 32: "30*1+1*2", 64: "62*1+1*2"}[M]
chunks = _chunk_het_matrix(synth.het_matrix(1, 3_000_000, 0), 500, 50_000)[:50]
kern = _PSMCKernelBase(M, chunks, overlap=500)
xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:500]
x = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
inds = torch.arange(S, device="cuda:0") * (50 // S)
for _ in range(reps):
    kern.hmm_term(x, pattern, 1e-2, inds, 500, weight=1.0)
torch.cuda.synchronize()
print(kern.last_kernel_name)
