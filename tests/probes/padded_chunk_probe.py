"""fp32 error on a chunk that ends in a long run of missing observations (the padded last chunk
of a contig): ours vs the reference's own fp32 kernel, both against fp64 at the same inputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import c_oracle, psmc_oracle as orc, ref_cuda
from phlash_b200.gpu import _PSMCKernelBase

L = 50_000
het = orc.synth_het_matrix(1, 8 * L + 20_000, seed=0)
data = orc.chunk_het_matrix(het, 500, L)[:, 500:].copy()
pps, _, _ = orc.synth_particles(16, 6, seed=0)
pa = pps[:, None].astype(np.float32).astype(np.float64)        # [6, 1, 7, 16]
for row in (8, 3):
    inds = np.array([row])
    ll_o, g_o = c_oracle.loglik_batch(data, np.tile(inds, 6), pa.reshape(6, 7, 16), grad=True)
    g_o = g_o.reshape(6, 1, 7, 16)
    scale = np.abs(g_o).max(-1, keepdims=True)
    def err(g):
        return (np.abs(g - g_o) / (np.abs(g_o) + 1e-3 * scale)).max(axis=(0, 1, 3))
    ours = _PSMCKernelBase(16, data)
    ll, g = ours.evaluate(pa, inds, True)
    ref = ref_cuda.ReferenceKernel(16, data, double_precision=False)
    ll_r, g_r = ref(pa, inds, grad=True)
    print(f"row {row} (missing tail: {(data[row] < 0).sum()} sites)")
    print("  ll rel err  ours %.2e  reference-fp32 %.2e" % (np.abs(ll.ravel() - ll_o).max() / np.abs(ll_o).max(), np.abs(ll_r.ravel() - ll_o).max() / np.abs(ll_o).max()))
    print("  grad err per row (b,d,u,v,e0,e1,pi)  ours", np.array2string(err(g), precision=1))
    print("  grad err per row            reference-fp32", np.array2string(err(g_r), precision=1))
    print("  non-finite entries: ours", int((~np.isfinite(g)).sum()), "reference-fp32", int((~np.isfinite(g_r)).sum()),
          "| sample d-row: oracle", g_o[3, 0, 1, :3], "ours", g[3, 0, 1, :3], "ref", g_r[3, 0, 1, :3], "ref ms", ref.last_ms)
