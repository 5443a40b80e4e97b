"""Generate tests/golden/*.npz by executing the reference's own, unmodified Python sources
(/root/reference/src/phlash/{hmm,params,transition,size_history,model,data}.py) with the jax API
backed by NumPy (jax_numpy_shim.py; jax is not installable in this image).

Run in the builder container only (``python tests/golden/make_golden.py``): /root/reference does
not exist on the GPU box, which is why the outputs are committed.  Nothing here is imported by
the product or by the test-suite.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jax_numpy_shim  # noqa: E402

jax_numpy_shim.install()

from phlash.data import _chunk_het_matrix  # noqa: E402
from phlash.hmm import PureJaxPSMCKernel, matvec_smc, psmc_ll  # noqa: E402
from phlash.model import log_density  # noqa: E402
from phlash.params import MCMCParams, PSMCParams  # noqa: E402
from phlash.size_history import DemographicModel, SizeHistory  # noqa: E402
from phlash.transition import _expQ, transition_matrix  # noqa: E402


def pp_block(pp: PSMCParams) -> np.ndarray:
    return np.stack([np.asarray(a, dtype=np.float64) for a in pp])


def fixture_data(seed):
    """tests/conftest.py:14-21 and tests/test_gpu.py:16-20 of the reference."""
    rng = np.random.default_rng(seed)
    data = (rng.uniform(size=(10, 1000)) < 0.05).astype(np.int8)
    missing = data.copy()
    inds = rng.integers(0, missing.size, size=int(0.01 * missing.size))
    missing.flat[inds] = -1
    return data, missing.clip(-1, 1)


def jittered_particles(n, seed):
    init = MCMCParams.from_linear(
        pattern="14*1+1*2", t1=1e-4, tM=15.0, c=np.ones(15), theta=1e-2, rho=1e-2
    )
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        out.append(
            MCMCParams(
                pattern=init.pattern,
                t_tr=init.t_tr + rng.standard_normal(2),
                c_tr=init.c_tr + rng.standard_normal(15),
                rho_over_theta_tr=init.rho_over_theta_tr + rng.standard_normal(),
                theta=init.theta,
                alpha=0.0,
                beta=0.0,
            )
        )
    return init, out


def main():
    out = {}
    # ---- A: the reference's test model (tests/conftest.py:24-31)
    dm = DemographicModel.default(pattern="16*1", theta=1e-2, rho=1e-2)
    pp = PSMCParams.from_dm(dm)
    out["dm16_t"] = np.asarray(dm.eta.t)
    out["dm16_c"] = np.asarray(dm.eta.c)
    out["dm16_ect"] = dm.eta.ect()
    out["dm16_pi"] = dm.eta.pi
    out["dm16_surv"] = dm.eta.surv()
    out["dm16_A"] = transition_matrix(dm)
    out["dm16_A_n5"] = transition_matrix(dm, 5)
    out["dm16_pp"] = pp_block(pp)

    # ---- B: jittered particles through to_dm / from_dm
    init, parts = jittered_particles(6, 7)
    xs, ts, cs, rhos, pps = [], [], [], [], []
    for mcp in [init] + parts:
        d = mcp.to_dm()
        xs.append(np.concatenate([np.ravel(mcp.t_tr), np.ravel(mcp.c_tr), [mcp.rho_over_theta_tr]]))
        ts.append(np.asarray(d.eta.t))
        cs.append(np.asarray(d.eta.c, dtype=np.float64))
        rhos.append(float(d.rho))
        pps.append(pp_block(PSMCParams.from_dm(d)))
    out["part_x"] = np.stack(xs)
    out["part_t"] = np.stack(ts)
    out["part_c"] = np.stack(cs)
    out["part_rho"] = np.array(rhos)
    out["part_pp"] = np.stack(pps)

    # ---- C: transition matrices at M = 32, 64 (from_dm itself asserts M == 16)
    for m in (32, 64):
        d = DemographicModel.default(pattern=f"{m}*1", theta=1e-2, rho=2e-2)
        out[f"dm{m}_A"] = transition_matrix(d)
        out[f"dm{m}_ect"] = d.eta.ect()
        out[f"dm{m}_pi"] = d.eta.pi
    # a non-constant size history with extreme rates (exercises the ect guards)
    t = np.concatenate([[0.0], np.geomspace(1e-3, 15.0, 15)])
    c = np.array([1.0, 1e-9, 150.0, 0.3, 2.0, 1.0, 0.01, 5.0, 1.0, 1.0, 40.0, 1.0, 0.5, 1.0, 3.0, 1.0])
    d = DemographicModel(eta=SizeHistory(t=t, c=c), theta=2e-2, rho=5e-3)
    out["odd_c"] = c
    out["odd_ect"] = d.eta.ect()
    out["odd_pi"] = d.eta.pi
    out["odd_A"] = transition_matrix(d)
    out["odd_pp"] = pp_block(PSMCParams.from_dm(d))

    # ---- expQ grid
    grid = [(r, cc, n) for r in (1e-9, 1e-3, 0.5, 30.0) for cc in (1e-9, 1e-2, 1.0, 50.0) for n in (2, 10)]
    out["expq_args"] = np.array(grid, dtype=np.float64)
    out["expq_vals"] = np.stack([np.asarray(_expQ(r, cc, int(n)), dtype=np.float64) for r, cc, n in grid])

    # ---- D: forward recursion on the reference's test fixtures
    pp_alt = PSMCParams(*out["part_pp"][2])
    for seed in (0, 1, 2):
        data, missing = fixture_data(seed)
        out[f"data_s{seed}"] = data
        out[f"missing_s{seed}"] = missing
        lls, alphas = [], []
        for which, params in (("dm", pp), ("alt", pp_alt)):
            for name, mat in (("data", data), ("missing", missing)):
                for row in (0, 1, 2):
                    a, ll = psmc_ll(params, mat[row])
                    lls.append(ll)
                    alphas.append(a)
        out[f"hmm_ll_s{seed}"] = np.array(lls)  # order: (dm|alt) x (data|missing) x row
        out[f"hmm_alpha_s{seed}"] = np.stack(alphas)
    v = np.random.default_rng(0).uniform(size=16)
    v /= v.sum()
    out["matvec_v"] = v
    out["matvec_out"] = matvec_smc(v, pp)

    # ---- E: central finite differences of the reference psmc_ll in log-parameter space
    data, missing = fixture_data(0)
    row = missing[1][:300]
    base = pp_block(pp_alt)
    h = 1e-5
    fd = np.zeros_like(base)
    for g in range(7):
        for k in range(16):
            if base[g, k] == 0.0:
                continue
            up, dn = base.copy(), base.copy()
            up[g, k] *= np.exp(h)
            dn[g, k] *= np.exp(-h)
            fd[g, k] = (psmc_ll(PSMCParams(*up), row)[1] - psmc_ll(PSMCParams(*dn), row)[1]) / (2 * h)
    out["fd_row"] = row
    out["fd_pp"] = base
    out["fd_grad"] = fd
    out["fd_ll"] = psmc_ll(PSMCParams(*base), row)[1]

    # ---- F: the HMM term of log_density (warm-up from the stationary pi, then the chunk)
    rng = np.random.default_rng(11)
    het = (rng.uniform(size=(2, 3000)) < 0.06).astype(np.int8)
    het[0, 100:130] = -1
    chunks = _chunk_het_matrix(het, overlap=50, chunk_size=500)
    warm, body = np.split(chunks, [50], axis=1)
    kern = PureJaxPSMCKernel(M=16, data=np.ascontiguousarray(body))
    inds = np.array([3, 0, 3, 7])
    out["model_het"] = het
    out["model_inds"] = inds
    out["model_chunks"] = chunks
    vals = []
    for mcp in [init] + parts[:2]:
        vals.append(
            log_density(mcp, c=np.array([0.0, 1.0, 0.0]), inds=inds, warmup=warm[inds], kern=kern, afs=None)
        )
    out["model_l2"] = np.array(vals, dtype=np.float64)

    # ---- G: chunk geometry (tests/test_data.py:18-28 and neighbours)
    rng = np.random.default_rng(5)
    for tag, shape, ov, cs in (("a", (1, 10_000), 123, 4_567), ("b", (3, 1_000), 10, 90), ("c", (2, 70), 20, 100)):
        h_in = rng.integers(-1, 4, size=shape)
        out[f"chunk_{tag}_in"] = h_in.astype(np.int8)
        out[f"chunk_{tag}_geom"] = np.array([ov, cs])
        out[f"chunk_{tag}_out"] = _chunk_het_matrix(h_in, overlap=ov, chunk_size=cs)

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
