"""Where does the fp32 gradient error at full chunk length come from?  (diagnostic)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import c_oracle, psmc_oracle as orc
from phlash_b200.gpu import _PSMCKernelBase

L = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
het = orc.synth_het_matrix(1, 8 * L + 20_000, seed=0)
chunks = orc.chunk_het_matrix(het, 500, L)
data = chunks[:, 500:].copy()
pps, _, _ = orc.synth_particles(16, 6, seed=0)
inds = np.arange(data.shape[0])
pa = np.broadcast_to(pps[:, None], (6, len(inds), 7, 16)).astype(np.float32).astype(np.float64).copy()
B, S = pa.shape[:2]
ll_o, g_o = c_oracle.loglik_batch(data, np.tile(inds, B), pa.reshape(B * S, 7, 16), grad=True)
g_o = g_o.reshape(B, S, 7, 16)
names = ["b", "d", "u", "v", "e0", "e1", "pi"]
for T in (2, 4):
    for dbl in (False,):
        k = _PSMCKernelBase(16, data, double_precision=dbl)
        k.set_threads_per_pair(T)
        ll, g = k.evaluate(pa, inds, True)
        print(f"T={T} ll rel err max {np.max(np.abs(ll - ll_o.reshape(B, S)) / np.abs(ll_o.reshape(B, S))):.2e}")
        rel = np.abs(g - g_o) / np.maximum(np.abs(g_o), 1e-300)
        scale = np.abs(g_o).max(-1, keepdims=True)
        signif = np.abs(g_o) > 1e-3 * scale
        for r, nm in enumerate(names):
            rr = np.where(signif[:, :, r], rel[:, :, r], 0)
            worst = np.unravel_index(np.argmax(rr), rr.shape)
            wv = (worst[0], worst[1], r, worst[2])
            print(f"  row {nm}: got {g[wv]:.6e} want {g_o[wv]:.6e} max rel (significant entries) {rr.max():.2e} at b,s,m={worst} value {g_o[worst[0], worst[1], r, worst[2]]:.3e}; "
                  f"median rel {np.median(rel[:, :, r][signif[:, :, r]]):.2e}; per-state max {np.array2string(rr.max((0, 1)), precision=1)}")
