"""Generate benchdata/particles_M{16,32,64}.npz: 500 jittered particles around
DemographicModel.default (law of src/phlash/mcmc.py:186-195) mapped to [7, M] HMM parameter
blocks with the CPU oracle's restatement of PSMCParams.from_dm.  Run once; outputs committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import psmc_oracle as orc  # noqa: E402

for m, n in ((16, 500), (32, 256), (64, 128)):
    pps, xs, pattern = orc.synth_particles(m, n, seed=0)
    assert np.isfinite(pps).all()
    path = os.path.join(ROOT, "benchdata", f"particles_M{m}.npz")
    np.savez_compressed(path, pps=pps, xs=xs, pattern=np.array(pattern))
    print(path, pps.shape, os.path.getsize(path))
