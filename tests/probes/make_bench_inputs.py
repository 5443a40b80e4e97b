"""Generate benchdata/particles_M{16,32,64}.npz: jittered particles around
DemographicModel.default (law of src/phlash/mcmc.py:186-195) mapped to [7, M] HMM parameter
blocks with the CPU oracle's restatement of PSMCParams.from_dm.  Run once; outputs committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import psmc_oracle as orc  # noqa: E402

# (M, particles whose unconstrained coordinates xs are stored, how many of them also get a stored
# [7, M] block): the blocks of the others are built on the device by phb_params_from_particles
for m, n, n_blocks in ((16, 500, 500), (32, 500, 256), (64, 1000, 128)):
    pps, xs, pattern = orc.synth_particles(m, n, seed=0)
    assert np.isfinite(pps).all()
    path = os.path.join(ROOT, "benchdata", f"particles_M{m}.npz")
    np.savez_compressed(path, pps=pps[:n_blocks], xs=xs, pattern=np.array(pattern))
    print(path, pps.shape, os.path.getsize(path))
