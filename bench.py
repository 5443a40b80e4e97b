"""Benchmark of the PSMC HMM log-likelihood + gradient hot path (BASELINE.json metric:
HMM site-transitions/sec, loglik+grad, and SVGD iters/sec at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3-strong|c4|c5]

Configurations (BASELINE.json `configs`; chunk geometry of the reference: 50 000-bin chunks + 500 overlap):

  c2 (default)  configs[1]: per GPU one diploid of 30 M bins -> 595 chunks, M = 16, 500 particles; one STEP =
                loglik + gradient of all 500 x 595 pairs = 1.4875e10 site-transitions per GPU (weak scaling).
  c3-strong     configs[2]: 10 diploids -> 5 950 chunks IN TOTAL, sharded over the N ranks (strong scaling);
                one STEP = all 500 x 5 950 pairs + the all-reduce of the per-particle sums.
  c4            configs[3]: M = 32, 500 particles, 25 000 resident chunks per GPU (what the reference keeps of
                59 500 after its down-sampling rule, mcmc.py:126-139); one STEP = fused warm-up loglik + gradient
                of a 2 048-chunk minibatch per GPU; SVGD iterations at the reference's own minibatch (S = 5).
  c5            configs[4]: M = 64, 1 000 particles, 10^6 resident chunks (50.5 GB of observations per GPU,
                constructor time reported); STEP = fused warm-up loglik + gradient of a 512-chunk minibatch.

Prints ONE JSON line (see the task contract); details in DESIGN.md section "Measurement".
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BINS = 30_000_000
CHUNK = 50_000
OVERLAP = 500
BYTES_PER_ST = 1.0            # int8 observation, not shared across particles (conservative)
CPU_SAMPLE = (16, 64)         # particles x chunks scored by the CPU baseline per step
PATTERNS = {16: "14*1+1*2", 32: "30*1+1*2", 64: "62*1+1*2"}
THETA = 1e-2

CONFIGS = {
    "c2": dict(M=16, B=500, scaling="weak", fused=False,
               workload="1 diploid whole-genome-scale synthetic (30M bins -> 595 chunks x 50000 bins + 500 overlap), M=16, "
                        "500 particles, per GPU"),
    "c3-strong": dict(M=16, B=500, scaling="strong", fused=False, diploids=10,
                      workload="10 diploids x 30M bins synthetic -> 5950 chunks x 50000 bins (+500 overlap) IN TOTAL, sharded "
                               "across the GPUs, M=16, 500 particles"),
    "c4": dict(M=32, B=500, scaling="weak", fused=True, rows=25_000, minibatch=2048, base_diploids=8,
               workload="100 diploids x 30M bins -> 59500 chunks, 25000 kept by the reference's down-sampling rule and resident "
                        "per GPU, M=32, 500 particles; step = fused warm-up loglik+grad of a 2048-chunk minibatch per GPU"),
    "c5": dict(M=64, B=1000, scaling="weak", fused=True, rows=1_000_000, minibatch=512, base_diploids=4,
               workload="stress: M=64, 1000 particles, 10^6 resident chunks x 50500 bins (50.5 GB int8 per GPU); step = fused "
                        "warm-up loglik+grad of a 512-chunk minibatch per GPU"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def diploid_chunks(seed):
    """full chunks [595, 50 500] of one synthetic diploid (reference geometry, data.py:37-61)"""
    from benchdata import synth
    from phlash_b200.data import _chunk_het_matrix

    return _chunk_het_matrix(synth.het_matrix(1, N_BINS, seed), OVERLAP, CHUNK)


def build_inputs(seed=0, with_full_chunks=False):
    """config c2: (data part [595, 50 000], parameter blocks [500, 7, 16] (, full chunks))"""
    from benchdata import synth
    from phlash_b200.data import split_warmup

    chunks = diploid_chunks(seed)
    _, data = split_warmup(chunks, OVERLAP)
    if with_full_chunks:
        return data, synth.particles(16, 500), chunks
    return data, synth.particles(16, 500)


def tiled_rows(base, n_rows):
    """[n_rows, W] made of copies of `base` (synthetic data: the kernel's speed does not depend on the content,
    and 50 GB of distinct synthetic bins would take minutes of host time to draw)"""
    out = np.empty((n_rows, base.shape[1]), dtype=np.int8)
    for r0 in range(0, n_rows, len(base)):
        n = min(len(base), n_rows - r0)
        out[r0:r0 + n] = base[:n]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows), "reasons": reasons}


def host_threads():
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would otherwise silently
    # make the CPU arm a single-thread run)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline(data, pps, steps=1, warmup=0):
    """The fp64 C/OpenMP port of the reference recursion (oracle/psmc_oracle.c) on all host cores,
    on a bounded sample of the same workload."""
    from oracle import c_oracle

    b, s = CPU_SAMPLE
    n_threads = host_threads()
    rows = np.tile(np.arange(s) * (data.shape[0] // s), b)
    params = np.repeat(pps[:b], s, axis=0)
    n_st = b * s * data.shape[1]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        c_oracle.loglik_batch(data, rows, params, grad=True, n_threads=n_threads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    m = pps.shape[-1]
    return {"value": n_st / dt, "unit": "site-transitions/s", "cores": n_threads, "kind": "port",
            "sample": f"{b} particles x {s} chunks x {data.shape[1]} bins, M={m}, loglik+grad, fp64 C/OpenMP restatement "
                      f"of hmm.py:52-82 (the reference's JAX CPU path needs jax, absent here)",
            "seconds_per_step": dt}


def reference_gpu(data, pps, M):
    """The reference's own CUDA kernel (fp32, NVRTC-compiled KERNEL_SRC, its own launch geometry)
    on the same GPU and inputs, on a bounded sample; reported for context (north-star target is
    10x this path)."""
    from oracle import ref_cuda

    if not ref_cuda.available(M, False):
        return {"unavailable": f"oracle/_ref cubin for M={M} not built (the reference's kernel does not compile beyond M=32)"}
    b, s = 500, 8
    inds = np.arange(s) * (data.shape[0] // s)
    pa = np.broadcast_to(pps[:b, None], (b, s, 7, M)).astype(np.float32)
    ref = ref_cuda.ReferenceKernel(M, data, double_precision=False)
    ref(pa[:8], inds, grad=True)  # warm-up
    ms = []
    for _ in range(2):
        ref(pa, inds, grad=True)
        ms.append(ref.last_ms)
    ref.close()
    return {"value": b * s * data.shape[1] / (min(ms) * 1e-3), "unit": "site-transitions/s",
            "sample": f"{b} particles x {s} chunks x {data.shape[1]} bins", "kernel_ms": min(ms),
            "what": "reference loglik_grad (src/phlash/gpu.py:575-692), fp32, grid (B,S) x block (7,M)"}


def likelihood_step(kern, M, B, n_rows, rank, world, local_rank, sizes):
    """One whole likelihood evaluation of an SVGD iteration on the device: particles -> parameters
    -> fused warm-up loglik+grad over the minibatch -> VJP to the particles (what model.log_density's
    HMM term and its reverse pass do in the reference: model.py:50-57, params.py:32-55), the minibatch
    sharded over the ranks + one all-reduce when world > 1.  Timed at the reference's default minibatch
    (mcmc.py:119-121: S = min(5, N / niter), i.e. S = 1 for one genome, S = 5 from ~5 000 chunks on) and at
    larger ones; eagerly and as ONE CUDA-graph replay (sampling on the device included)."""
    import torch

    from phlash_b200 import model

    dev = torch.device("cuda", local_rank)
    xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:B]
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    pattern = PATTERNS[M]
    out = {}
    for name, S in sizes:
        inds = torch.arange(S, dtype=torch.int64, device=dev) * (n_rows // S)
        kern.reserve(B, S, OVERLAP)

        def call():
            return model.hmm_term_value_and_grad(kern, x, pattern, THETA, inds, OVERLAP, rank=rank, world=world)

        for _ in range(2):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if S <= 8 else 2
        e0.record()
        for _ in range(reps):
            val, grad = call()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        assert torch.isfinite(val).all() and torch.isfinite(grad).all()
        rec = {"particles": B, "minibatch_chunks": S, "bins_per_chunk": OVERLAP + CHUNK, "ms": ms, "evaluations_per_s": 1e3 / ms}
        if S <= 8:
            # the same step recorded into a CUDA graph (phb_reserve makes it allocation free), new minibatch per replay
            try:
                side = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(side):
                    kern.sample_minibatch(1, S, out=inds)
                    call()
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    kern.sample_minibatch(1, S, out=inds)
                    call()
                for _ in range(3):
                    graph.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(20):
                    graph.replay()
                e1.record()
                e1.synchronize()
                rec["ms_cuda_graph_replay"] = e0.elapsed_time(e1) / 20
                del graph
            except Exception as e:  # an extra, never allowed to take the benchmark down
                rec["ms_cuda_graph_replay"] = None
                rec["cuda_graph_error"] = repr(e)[:160]
        out[name] = rec
    return out


def e2e_small(kern, M, B, n_rows):
    """The blocking whole-term entry a jax.pure_callback binds (phb_hmm_term_host) with PAGEABLE NumPy buffers,
    wall clock over >= 50 calls, at the reference's default minibatch sizes."""
    xs = np.ascontiguousarray(np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:B])
    out = {}
    for name, S in (("S1", 1), ("S5", 5)):
        inds = (np.arange(S) * (n_rows // S)).astype(np.int64)
        for _ in range(3):
            kern.hmm_term_host(xs, PATTERNS[M], THETA, inds, OVERLAP, weight=n_rows / S)
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            value, grad = kern.hmm_term_host(xs, PATTERNS[M], THETA, inds, OVERLAP, weight=n_rows / S)
        dt = (time.perf_counter() - t0) / n
        assert np.isfinite(value).all() and np.isfinite(grad).all()
        out[name] = {"ms_per_call": dt * 1e3, "site_transitions_per_s": B * S * CHUNK / dt, "calls": n,
                     "h2d_bytes_per_call": int(xs.nbytes + inds.nbytes), "d2h_bytes_per_call": int(value.nbytes + grad.nbytes),
                     "api": "phb_hmm_term_host, pageable NumPy buffers, blocking"}
    return out


def elpd_step(local_rank, M=16, B=500, rank=0, world=1):
    """The reference's ELPD evaluation (mcmc.py:213-238): forward-only HMM term of all particles over one
    un-chunked held-out contig (2.5 M bins = a 250 Mb chromosome at 100 bp), through the one-call entry;
    with the parallel-in-time path (automatic) and with the sequential kernel.  With several processes the
    particles are sharded (one all-reduce of a scalar) and the time is the maximum over the ranks."""
    import torch

    from benchdata import synth
    from phlash_b200 import model

    dev = torch.device("cuda", local_rank)
    n_bins = 2_500_000
    tk = model.elpd_kernel(M, synth.het_matrix(1, n_bins, seed=101), device=local_rank)
    xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:B]
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    out = {"particles": B, "test_contigs": 1, "bins": n_bins, "world": world,
           "path": "transfer_rows_kernel + chain_transfer_kernel (parallel in time), then the 1-bin warm-up term"
                   + ("; particles sharded over the processes" if world > 1 else "")}
    for name, mode in (("ms", -1), ("ms_sequential_kernel", 0)):
        if mode == 0 and world > 1:
            continue
        tk.set_parallel_in_time(mode)
        for _ in range(2):
            e = model.elpd_hmm_term(tk, x, PATTERNS[M], THETA, rank=rank, world=world)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            e = model.elpd_hmm_term(tk, x, PATTERNS[M], THETA, rank=rank, world=world)
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t)
        out["elpd"] = float(e)
        assert bool(torch.isfinite(e))
    return out


def parity_block(host_rows, fused, inds_np, pps32, ll, dlog, n_check=16, seed=0):
    """After timing: (1) oracle spot check of n_check sampled pairs of the LAST timed step's outputs, at the
    tolerances of tests/test_gpu_parity.py; (2) oracle-free invariants over ALL pairs of that step (every
    posterior sums to one): sum_m(dlog_e0 + dlog_e1) = #observed bins, sum_m(dlog_b + dlog_d + dlog_v) = #bins,
    sum_m dlog_pi = 1 (0 for the fused warm-up difference); structural zeros exactly 0.
    host_rows(i) -> the int8 row the kernel scored for chunk position i (full chunk when fused)."""
    from oracle import c_oracle

    B, S = ll.shape
    M = dlog.shape[-1]
    rng = np.random.default_rng(seed)
    pick_b = rng.integers(0, B, n_check)
    pick_s = rng.integers(0, S, n_check)
    rows = np.stack([host_rows(int(inds_np[s])) for s in pick_s])
    pa = pps32[pick_b].astype(np.float64)
    idx = np.arange(n_check)
    ll_ref, g_ref = c_oracle.loglik_batch(rows, idx, pa, grad=True)
    size = np.abs(g_ref)
    if fused:
        ll_w, g_w = c_oracle.loglik_batch(np.ascontiguousarray(rows[:, :OVERLAP]), idx, pa, grad=True)
        ll_ref, g_ref, size = ll_ref - ll_w, g_ref - g_w, np.abs(g_ref) + np.abs(g_w)
    got_ll = ll[pick_b, pick_s]
    got_g = dlog[pick_b, pick_s].astype(np.float64)
    ll_err = float(np.max(np.abs(got_ll - ll_ref) / np.abs(ll_ref)))
    tol = 1e-4 * size + 1e-7 * size.max(-1, keepdims=True)
    g_ok = bool(np.all(np.abs(got_g - g_ref) <= tol))
    g_err = float(np.max(np.abs(got_g - g_ref) / np.maximum(size, 1e-3 * size.max(-1, keepdims=True))))
    # invariants over all pairs
    first = OVERLAP if fused else 0
    n_obs = np.array([(host_rows(int(i))[first:] >= 0).sum() for i in inds_np], dtype=np.float64)
    n_bins = float(len(host_rows(int(inds_np[0]))) - first)
    emis = dlog[:, :, 4].sum(-1, dtype=np.float64) + dlog[:, :, 5].sum(-1, dtype=np.float64)
    trans = sum(dlog[:, :, r].sum(-1, dtype=np.float64) for r in (0, 1, 3))
    pim = dlog[:, :, 6].sum(-1, dtype=np.float64)
    inv = {
        "emission_mass_rel_err": float(np.max(np.abs(emis - n_obs[None]) / n_obs[None])),
        "transition_mass_rel_err": float(np.max(np.abs(trans - n_bins) / n_bins)),
        "pi_mass_abs_err": float(np.max(np.abs(pim - (0.0 if fused else 1.0)))),
        "structural_zeros_exact": bool(np.all(dlog[:, :, 0, -1] == 0) and np.all(dlog[:, :, 2, -1] == 0) and np.all(dlog[:, :, 3, 0] == 0)),
        "all_finite": bool(np.isfinite(ll).all() and np.isfinite(dlog).all()),
    }
    ok = (ll_err <= 1e-5 and g_ok and inv["emission_mass_rel_err"] <= 5e-5 and inv["transition_mass_rel_err"] <= 5e-5
          and inv["pi_mass_abs_err"] <= 5e-4 and inv["structural_zeros_exact"] and inv["all_finite"])
    return {"ok": bool(ok), "oracle_pairs": int(n_check), "pairs_checked_by_invariants": int(B * S),
            "ll_max_rel_err": ll_err, "ll_tol": 1e-5, "grad_max_rel_err": g_err, "grad_tol": 1e-4,
            "grad_within_tol": g_ok, "invariants": inv,
            "oracle": "oracle/psmc_oracle.c (fp64) at the fp32-rounded parameters, outside the timed region"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    M = cfg["M"]
    from benchdata import synth
    from phlash_b200.data import split_warmup

    chunks = diploid_chunks(0)
    _, data = split_warmup(chunks, OVERLAP)
    pps = synth.particles(M, CPU_SAMPLE[0])
    res = cpu_baseline(data, pps, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "HMM site-transitions/sec (loglik+grad)", "value": res["value"],
        "unit": "site-transitions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": args.config, "M": M, "particles": cfg["B"], "chunk_bins": CHUNK,
                   "overlap": OVERLAP, "step": "bounded sample: " + res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "site-transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pcie_rate_gbs(dev):
    """pinned host -> device copy rate measured now (the yard-stick for the constructor's upload)"""
    import torch

    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d.copy_(h, non_blocking=True)
    e1.record()
    e1.synchronize()
    return n / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=int, default=0, help="c4 / c5: resident chunk rows per GPU (default: the configuration's)")
    ap.add_argument("--skip-baselines", action="store_true", help="omit the cpu_baseline / reference_gpu / extra legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: phlash_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from benchdata import synth
    from phlash_b200.distributed import all_reduce_sum, shard_bounds
    from phlash_b200.gpu import _PSMCKernelBase, measure_fp32_peak

    cfg = CONFIGS[args.config]
    M, B, fused = cfg["M"], cfg["B"], cfg["fused"]
    flop_per_st = 36 * M          # loglik + grad
    pps = synth.particles(M, B) if M == 16 else None
    create_info = None

    # ---- resident observations
    if args.config == "c2":
        # every rank scores its own diploid (weak scaling; the observation matrix is resident per GPU)
        chunks_full = diploid_chunks(rank)
        data = np.ascontiguousarray(chunks_full[:, OVERLAP:])
        kern = _PSMCKernelBase(M, data, device=local_rank)
        kern_full = None  # built later for the whole-likelihood legs
        S = data.shape[0]
        global_S = S * world
        host_row = lambda i: data[i]  # noqa: E731
    elif args.config == "c3-strong":
        per = 595
        total = per * cfg["diploids"]
        lo, hi = shard_bounds(total, rank, world)
        parts = [diploid_chunks(d) for d in range(lo // per, (hi - 1) // per + 1)]
        chunks_full = np.concatenate(parts)[lo - (lo // per) * per:][: hi - lo]
        data = np.ascontiguousarray(chunks_full[:, OVERLAP:])
        kern = _PSMCKernelBase(M, data, device=local_rank)
        kern_full = None
        S = data.shape[0]
        global_S = total
        host_row = lambda i: data[i]  # noqa: E731
    else:
        n_rows = args.rows or cfg["rows"]
        base = np.concatenate([diploid_chunks(100 * rank + d) for d in range(cfg["base_diploids"])])
        t0 = time.perf_counter()
        chunks_full = tiled_rows(base, n_rows)
        t_host = time.perf_counter() - t0
        rate = pcie_rate_gbs(dev)
        t0 = time.perf_counter()
        kern = _PSMCKernelBase(M, chunks_full, device=local_rank, overlap=OVERLAP)
        t_create = time.perf_counter() - t0
        create_info = {"resident_bytes": int(chunks_full.nbytes), "rows": int(n_rows), "create_seconds": t_create,
                       "pinned_h2d_gbs_measured": rate, "pcie_copy_seconds_at_that_rate": chunks_full.nbytes / (rate * 1e9),
                       "create_over_pcie_copy": t_create / (chunks_full.nbytes / (rate * 1e9)),
                       "host_tiling_seconds": t_host,
                       "what": "phb_create_chunks: pageable host matrix -> 2 pinned slabs filled by up to 16 host threads -> H2D, "
                               "validated / clipped / padded on the device per slab, rows with long constant runs marked"}
        kern_full = kern
        S = cfg["minibatch"]
        global_S = S * world
        host_row = lambda i: chunks_full[i]  # noqa: E731
    n_chunks, length = (kern._N, CHUNK)
    st_per_step = B * S * length

    # ---- device-resident inputs for `value`
    if pps is None:
        xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:B]
        p7 = kern.params_from_particles(torch.tensor(xs, dtype=torch.float64, device=dev), PATTERNS[M], THETA)
        pps = p7.double().cpu().numpy()
    else:
        p7 = torch.tensor(pps, dtype=torch.float32, device=dev).contiguous()
    p6 = p7[:, :6].contiguous()
    pi = p7[:, 6].contiguous()
    if fused:
        inds_np = np.sort(np.random.default_rng(rank).choice(n_chunks, size=S, replace=False)).astype(np.int64)
    else:
        inds_np = np.arange(S, dtype=np.int64)
    inds_d = torch.tensor(inds_np, device=dev)
    ll_d = torch.empty((B, S), dtype=torch.float64, device=dev)
    dlog_d = torch.empty((B, S, 7, M), dtype=torch.float32, device=dev)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    reduced = torch.empty((B, 1 + 7 * M), dtype=torch.float64, device=dev)
    sum_stream = torch.cuda.current_stream(dev).cuda_stream

    def step_device():
        if fused:
            kern.evaluate_warmup_device(p7, inds_d, OVERLAP, True, ll=ll_d, dlog=dlog_d)
        else:
            kern.evaluate_device(p6, pi, inds_d, True, ll=ll_d, dlog=dlog_d)
        if world > 1:
            # per-particle sums over this rank's chunks (the library's own summing kernel), then ONE all-reduce
            kern.sum_over_chunks(ll_d, dlog_d, out=reduced)
            all_reduce_sum(reduced)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = kern.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    torch.cuda.synchronize()
    for e0, e1 in ev:
        l2_flush.fill_(1)  # evict L2 between timed steps (untimed)
        e0.record()
        step_device()
        e1.record()
        e1.synchronize()
        kernel_ms.append(kern.last_kernel_ms)
    kernel_name = kern.last_kernel_name
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = kern.launch_count - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = B * global_S * length / (ms_per_step * 1e-3)

    # ---- parity of the timed step's outputs (oracle: outside the timed region, rank 0)
    parity = None
    if rank == 0:
        try:
            parity = parity_block(host_row, fused, inds_np, pps.astype(np.float32), ll_d.cpu().numpy(), dlog_d.cpu().numpy())
        except Exception as e:
            parity = {"ok": False, "error": repr(e)[:200]}

    # ---- end to end through the reference-facing call with HOST buffers (pageable NumPy, what jax.pure_callback hands over)
    inds_h = inds_np.copy()
    ll_np = np.empty((B, S), dtype=np.float64)
    dlog_np = np.empty((B, S, 7, M), dtype=np.float32)
    if fused:
        p7_np = pps.astype(np.float32)

        def e2e_call():
            return kern.evaluate_warmup(p7_np, inds_h, OVERLAP, True)

        h2d = p7_np.nbytes + inds_h.nbytes
        api = "phb_loglik_warmup_host, params7 [B,7,M] fp32 + inds, pageable NumPy"
    else:
        pa_np = np.ascontiguousarray(np.broadcast_to(pps[:, None], (B, S, 7, M)).astype(np.float32))

        def e2e_call():
            return kern.evaluate(pa_np, inds_h, True, ll_out=ll_np, dlog_out=dlog_np)

        h2d = pa_np.nbytes + inds_h.nbytes
        api = "phb_loglik_host via PSMCKernel host entry, pa [B,S,7,M] fp32, pageable NumPy in and out"
    e2e_steps = max(1, min(args.steps, 3))
    e2e_call()  # warm-up (grows scratch)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ll_e, dlog_e = e2e_call()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * global_S * length / float(t.item())
    assert np.isfinite(ll_e).all() and np.allclose(ll_e, ll_d.cpu().numpy(), rtol=1e-9)
    d2h = ll_e.nbytes + dlog_e.nbytes
    del ll_e, dlog_e

    # ---- the whole likelihood step / SVGD iteration legs (every N: the minibatch is sharded over the ranks)
    lik_step = svgd = small = None
    if not args.skip_baselines:
        if kern_full is None:
            kern_full = _PSMCKernelBase(M, chunks_full, device=local_rank, overlap=OVERLAP)
        sizes = [("S1", 1), ("S5", 5)]
        sizes.append(("SN", kern_full._N) if not fused else ("S64", 64))
        lik_step = likelihood_step(kern_full, M, B, kern_full._N, rank, world, local_rank, sizes)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import svgd_demo

            svgd = {}
            for name, S_it, n_it in (("S1", 1, 60), ("S5", 5, 40)):
                svgd[name] = svgd_demo.run(n_iter=n_it, S=S_it, device=local_rank, M=M, particles=B, rank=rank, world=world,
                                           chunks=chunks_full, kern=kern_full)
        except Exception as e:  # an extra, never allowed to take the benchmark down
            svgd = {"unavailable": repr(e)[:300]}
        if world == 1:
            try:
                small = e2e_small(kern_full, M, B, kern_full._N)
            except Exception as e:
                small = {"unavailable": repr(e)[:200]}
    elpd = None
    if not args.skip_baselines and M == 16:
        try:
            elpd = elpd_step(local_rank, rank=rank, world=world)
        except Exception as e:  # an extra, never allowed to take the benchmark down
            elpd = {"unavailable": repr(e)[:200]}
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        k_ms = float(np.mean(kernel_ms))
        ffma_peak, ffma_acc = measure_fp32_peak(local_rank)
        # the fused evaluation is two launches (all bins, then the warm-up bins); kernel_ms is the last one's, so
        # the roofline of the fused configurations is quoted on the whole step
        roof_ms = k_ms if not fused else ms_per_step
        hbm_achieved = st_per_step * BYTES_PER_ST / (roof_ms * 1e-3) / 1e9
        fp32_achieved = st_per_step * flop_per_st / (roof_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        if os.path.exists(tpath) and args.config == "c2":
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        line = {
            "metric": "HMM site-transitions/sec (loglik+grad)", "value": value, "unit": "site-transitions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "M": M, "particles": B, "chunks_per_step_per_gpu": S,
                       "resident_chunks_per_gpu": n_chunks, "chunk_bins": length, "overlap": OVERLAP, "pairs_per_gpu": B * S,
                       "site_transitions_per_step_per_gpu": st_per_step,
                       "rows_scored_in_double": kern.num_escalated_rows,
                       "l2": "256 MiB L2 flush between timed steps",
                       "step": ("fused warm-up loglik+grad (phb_loglik_warmup_device): LL(all 50500 bins) - LL(500 warm-up bins); "
                                "site-transitions count the 50000 data bins only") if fused else
                               "loglik+grad of the data part of every chunk (phb_loglik_device)",
                       "parallelism": (f"dp{world}: chunks sharded, per-particle sums (sum_over_chunks_kernel) + 1 all-reduce of "
                                       f"[B,1+7M] per step") if world > 1 else "single GPU"},
            "roofline": {"bound": "fp32_fma", "achieved": fp32_achieved, "peak": ffma_peak, "unit": "TFLOP/s",
                         "frac": fp32_achieved / ffma_peak, "traffic": traffic,
                         "peak_source": "independent-FFMA chains measured in this run (phb_measure_fp32_peak); "
                                        "MEASURED_PEAKS.json has no fp32 entry",
                         "flop_per_site_transition": flop_per_st, "kernel": kernel_name, "kernel_ms": roof_ms,
                         "peak_accumulate_pattern": ffma_acc,
                         "note": "algorithmic flops (36 M per site-transition, SURVEY 8d) / CUDA-event time of the kernel; the "
                                 "FP32 FMA pipe is the binding roof, HBM is secondary (see roofline_secondary)"},
            "roofline_secondary": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": hbm_achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                                   "note": "1 algorithmic byte per site-transition (int8 observation)"},
            "e2e": {"value": e2e_value, "unit": "site-transitions/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": api, "steps": e2e_steps},
            "parity": parity,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if create_info is not None:
            line["constructor"] = create_info
        if lik_step is not None:
            line["likelihood_step"] = lik_step
        if svgd is not None:
            line["svgd_harness"] = svgd
        if small is not None:
            line["e2e_small"] = small
        if elpd is not None:
            line["elpd_step"] = elpd
        if not args.skip_baselines:
            cpu_rows = np.ascontiguousarray(chunks_full[:595, OVERLAP:])
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(cpu_rows, pps).items() if k != "seconds_per_step"}
            try:
                line["reference_gpu"] = reference_gpu(cpu_rows, pps, M)
            except Exception as e:  # the comparator must never take the benchmark down
                line["reference_gpu"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
