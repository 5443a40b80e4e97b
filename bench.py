"""Benchmark of the PSMC HMM log-likelihood + gradient hot path (BASELINE.json metric:
HMM site-transitions/sec, loglik+grad).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every N (BASELINE.json configs[1], weak scaling = configs[2]'s sharding): per GPU one
diploid, 30 M bins chunked with the reference geometry (chunk 50 000 + overlap 500 -> 595 chunks),
M = 16, 500 SVGD particles; one STEP = loglik + gradient of all 500 x 595 (particle, chunk) pairs =
1.4875e10 site-transitions per GPU.  Prints ONE JSON line (see the task contract); details in
DESIGN.md section "Measurement".
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

M = 16
N_PARTICLES = 500
N_BINS = 30_000_000
CHUNK = 50_000
OVERLAP = 500
FLOP_PER_ST = 36 * M          # loglik + grad
BYTES_PER_ST = 1.0            # int8 observation, not shared across particles (conservative)
FFMA_PEAK_TFLOPS = 71.76      # measured on this pool's B200, profiles/r01_microbench_b200.json
CPU_SAMPLE = (16, 64)         # particles x chunks scored by the CPU baseline per step


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def build_inputs(seed=0, with_full_chunks=False):
    from benchdata import synth
    from phlash_b200.data import _chunk_het_matrix, split_warmup

    het = synth.het_matrix(1, N_BINS, seed)
    chunks = _chunk_het_matrix(het, OVERLAP, CHUNK)
    _, data = split_warmup(chunks, OVERLAP)
    # the reference rejects rows without a single observation (gpu.py:111-113)
    assert np.all(data.max(axis=1) > -1)
    if with_full_chunks:
        return data, synth.particles(M, N_PARTICLES), chunks
    return data, synth.particles(M, N_PARTICLES)


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


def cpu_baseline(data, pps, steps=1, warmup=0):
    """The fp64 C/OpenMP port of the reference recursion (oracle/psmc_oracle.c) on all host cores,
    on a bounded sample of the same workload."""
    from oracle import c_oracle

    b, s = CPU_SAMPLE
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would
    # otherwise silently make this a single-thread run)
    n_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
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
    return {"value": n_st / dt, "unit": "site-transitions/s", "cores": n_threads, "kind": "port",
            "sample": f"{b} particles x {s} chunks x {data.shape[1]} bins, loglik+grad, fp64 C/OpenMP restatement "
                      f"of hmm.py:52-82 (the reference's JAX CPU path needs jax, absent here)",
            "seconds_per_step": dt}


def reference_gpu(data, pps):
    """The reference's own CUDA kernel (fp32, NVRTC-compiled KERNEL_SRC, its own launch geometry)
    on the same GPU and inputs, on a bounded sample; reported for context (north-star target is
    10x this path)."""
    from oracle import ref_cuda

    if not ref_cuda.available(M, False):
        return {"unavailable": "oracle/_ref cubins not built"}
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


def likelihood_step(local_rank, chunks_full, rank, world):
    """One whole likelihood evaluation of an SVGD iteration on the device: particles -> parameters
    -> fused warm-up loglik+grad over the minibatch -> VJP to the particles (what model.log_density's
    HMM term and its reverse pass do in the reference: model.py:50-57, params.py:32-55).  Timed at
    the reference's default minibatch (mcmc.py:119-121: S = min(5, N / niter), i.e. S = 1 for one genome
    = this workload, S = 5 from ~5 000 chunks on) and at S = N.  The SVGD update
    itself (blackjax) is not part of this repository."""
    import torch

    from phlash_b200 import model
    from phlash_b200.gpu import _PSMCKernelBase

    dev = torch.device("cuda", local_rank)
    kern = _PSMCKernelBase(M, chunks_full, double_precision=False, device=local_rank)
    xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:N_PARTICLES]
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    out = {}
    for name, S in (("S1", 1), ("S5", 5), ("SN", chunks_full.shape[0])):
        inds = torch.arange(S, dtype=torch.int64, device=dev) * (chunks_full.shape[0] // S)
        for _ in range(2):
            model.hmm_term_value_and_grad(kern, x, "14*1+1*2", 1e-2, inds, OVERLAP, rank=rank, world=world)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if S <= 5 else 2
        e0.record()
        for _ in range(reps):
            val, grad = model.hmm_term_value_and_grad(kern, x, "14*1+1*2", 1e-2, inds, OVERLAP, rank=rank, world=world)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        assert torch.isfinite(val).all() and torch.isfinite(grad).all()
        out[name] = {"particles": N_PARTICLES, "minibatch_chunks": S, "bins_per_chunk": int(chunks_full.shape[1]),
                     "ms": ms, "evaluations_per_s": 1e3 / ms}
    return out


def elpd_step(local_rank):
    """The reference's ELPD evaluation (mcmc.py:213-238): forward-only HMM term of all particles over one
    un-chunked held-out contig (2.5 M bins = a 250 Mb chromosome at 100 bp), through the one-call entry;
    with the parallel-in-time path (automatic) and with the sequential kernel."""
    import torch

    from benchdata import synth
    from phlash_b200 import model

    dev = torch.device("cuda", local_rank)
    n_bins = 2_500_000
    tk = model.elpd_kernel(M, synth.het_matrix(1, n_bins, seed=101), device=local_rank)
    xs = np.load(os.path.join(ROOT, "benchdata", f"particles_M{M}.npz"))["xs"][:N_PARTICLES]
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    out = {"particles": N_PARTICLES, "test_contigs": 1, "bins": n_bins,
           "path": "transfer_rows_kernel + chain_transfer_kernel (parallel in time), then the 1-bin warm-up term"}
    for name, mode in (("ms", -1), ("ms_sequential_kernel", 0)):
        tk.set_parallel_in_time(mode)
        for _ in range(2):
            e = model.elpd_hmm_term(tk, x, "14*1+1*2", 1e-2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            e = model.elpd_hmm_term(tk, x, "14*1+1*2", 1e-2)
        e1.record()
        e1.synchronize()
        out[name] = e0.elapsed_time(e1) / 3
        assert bool(torch.isfinite(e))
    return out


def run_reference_arm(args, rank):
    if rank != 0:
        return
    data, pps = build_inputs()
    res = cpu_baseline(data, pps, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "HMM site-transitions/sec (loglik+grad)", "value": res["value"],
        "unit": "site-transitions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(), "M": M, "particles": N_PARTICLES, "chunk_bins": CHUNK,
                   "overlap": OVERLAP, "step": "bounded sample: " + res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "site-transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name():
    return ("1 diploid whole-genome-scale synthetic (30M bins -> 595 chunks x 50000 bins + 500 overlap), M=16, "
            "500 particles, per GPU")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-baselines", action="store_true", help="omit the cpu_baseline / reference_gpu legs")
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
    from phlash_b200.distributed import all_reduce_sum, pack_per_particle
    from phlash_b200.gpu import _PSMCKernelBase

    # every rank scores its own diploid (weak scaling; the observation matrix is resident per GPU)
    data, pps, chunks_full = build_inputs(seed=rank, with_full_chunks=True)
    kern = _PSMCKernelBase(M, data, double_precision=False, device=local_rank)
    n_chunks, length = data.shape
    B, S = N_PARTICLES, n_chunks
    st_per_step = B * S * length

    # ---- device-resident inputs for `value`
    p6 = torch.tensor(pps[:, :6], dtype=torch.float32, device=dev).contiguous()
    pi = torch.tensor(pps[:, 6], dtype=torch.float32, device=dev).contiguous()
    inds_d = torch.arange(S, dtype=torch.int64, device=dev)
    ll_d = torch.empty((B, S), dtype=torch.float64, device=dev)
    dlog_d = torch.empty((B, S, 7, M), dtype=torch.float32, device=dev)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    reduced = torch.empty((B, 1 + 7 * M), dtype=torch.float64, device=dev)

    def step_device():
        kern.evaluate_device(p6, pi, inds_d, True, ll=ll_d, dlog=dlog_d)
        if world > 1:
            # per-particle sums over this rank's chunks, then ONE all-reduce per step
            pack_per_particle(ll_d, dlog_d, out=reduced)
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
    value = world * st_per_step / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing call with HOST buffers (pinned)
    pa_h = torch.empty((B, S, 7, M), dtype=torch.float32).pin_memory()
    pa_h.copy_(torch.tensor(np.broadcast_to(pps[:, None], (B, S, 7, M)), dtype=torch.float32))
    ll_h = torch.empty((B, S), dtype=torch.float64).pin_memory()
    dlog_h = torch.empty((B, S, 7, M), dtype=torch.float32).pin_memory()
    inds_h = np.arange(S, dtype=np.int64)
    pa_np, ll_np, dlog_np = pa_h.numpy(), ll_h.numpy(), dlog_h.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    kern.evaluate(pa_np, inds_h, True, ll_out=ll_np, dlog_out=dlog_np)  # warm-up (grows scratch)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        kern.evaluate(pa_np, inds_h, True, ll_out=ll_np, dlog_out=dlog_np)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * st_per_step / float(t.item())
    assert np.isfinite(ll_np).all() and np.allclose(ll_np, ll_d.cpu().numpy(), rtol=1e-9)

    lik_step = None
    if not args.skip_baselines:
        lik_step = likelihood_step(local_rank, chunks_full, rank, world)
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        k_ms = float(np.mean(kernel_ms))
        hbm_achieved = st_per_step * BYTES_PER_ST / (k_ms * 1e-3) / 1e9
        fp32_achieved = st_per_step * FLOP_PER_ST / (k_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        line = {
            "metric": "HMM site-transitions/sec (loglik+grad)", "value": value, "unit": "site-transitions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(), "M": M, "particles": B, "chunks": S, "chunk_bins": length,
                       "overlap": OVERLAP, "pairs_per_gpu": B * S, "site_transitions_per_step_per_gpu": st_per_step,
                       "rows_scored_in_double": kern.num_escalated_rows,
                       "l2": "256 MiB L2 flush between timed steps",
                       "parallelism": f"dp{world}: chunks sharded, 1 all-reduce of [B,1+7M] per step" if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": hbm_achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                         "kernel": kernel_name, "kernel_ms": k_ms,
                         "note": "1 algorithmic byte per site-transition; the binding pipe is FP32 FMA, see fp32"},
            "fp32": {"achieved_tflops": fp32_achieved, "peak_tflops": FFMA_PEAK_TFLOPS,
                     "frac": fp32_achieved / FFMA_PEAK_TFLOPS, "flop_per_site_transition": FLOP_PER_ST,
                     "peak_source": "independent-FFMA microbenchmark on this pool's B200 (profiles/r01_microbench_b200.json)"},
            "e2e": {"value": e2e_value, "unit": "site-transitions/s",
                    "h2d_bytes_per_step": int(pa_np.nbytes + inds_h.nbytes), "d2h_bytes_per_step": int(ll_np.nbytes + dlog_np.nbytes),
                    "api": "phb_loglik_host via PSMCKernel host entry, pa [B,S,7,M] fp32 pinned", "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if lik_step is not None:
            line["likelihood_step"] = lik_step
        if not args.skip_baselines and world == 1:
            try:
                line["elpd_step"] = elpd_step(local_rank)
            except Exception as e:  # an extra, never allowed to take the benchmark down
                line["elpd_step"] = {"unavailable": repr(e)[:200]}
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import svgd_demo

                # S = 1 is the reference's default minibatch for this workload (595 chunks, mcmc.py:119-121)
                line["svgd_harness"] = {"S1": svgd_demo.run(n_iter=60, S=1, device=local_rank),
                                        "S5": svgd_demo.run(n_iter=40, S=5, device=local_rank)}
            except Exception as e:  # an extra, never allowed to take the benchmark down
                line["svgd_harness"] = {"unavailable": repr(e)[:200]}
        if not args.skip_baselines:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(data, pps).items() if k != "seconds_per_step"}
            try:
                line["reference_gpu"] = reference_gpu(data, pps)
            except Exception as e:  # the comparator must never take the benchmark down
                line["reference_gpu"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
