"""One process per GPU over NCCL (needs >= 2 GPUs; skipped on a single-GPU lease - the log of a 2-GPU run is
kept under profiles/): the sharded likelihood step equals the single-process one, both for a minibatch with
at least one chunk per process (chunk axis sharded) and for fewer chunks than processes (time axis sharded)."""

import os

import numpy as np
import pytest

from oracle import psmc_oracle as orc

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, chunks, xs, pattern, cases, out_dir):
    import torch
    import torch.distributed as dist

    from phlash_b200 import model
    from phlash_b200.gpu import _PSMCKernelBase

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    kern = _PSMCKernelBase(16, chunks, device=rank, overlap=500)
    x = torch.tensor(xs, dtype=torch.float64, device=dev)
    for name, inds in cases:
        inds_d = torch.tensor(inds, device=dev)
        value, grad = model.hmm_term_value_and_grad(kern, x, pattern, 1e-2, inds_d, 500, weight=2.5, rank=rank, world=world)
        torch.cuda.synchronize()
        if rank == world - 1:  # (process 0's last launch is the warm-up term it subtracts)
            np.savez(os.path.join(out_dir, f"{name}.npz"), value=value.cpu().numpy(), grad=grad.cpu().numpy(),
                     kernel=np.array(kern.last_kernel_name))
    # the ELPD (forward only, un-chunked test contig) with the particles sharded over the processes
    test_kern = model.elpd_kernel(16, chunks[:2, 500:30_500], device=rank)
    e = model.elpd_hmm_term(test_kern, x, pattern, 1e-2, rank=rank, world=world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "elpd.npz"), elpd=float(e))
    dist.destroy_process_group()


def test_sharded_step_equals_single_process(tmp_path):
    import torch
    import torch.multiprocessing as mp

    from phlash_b200.gpu import _PSMCKernelBase

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    het = orc.synth_het_matrix(1, 420_000, seed=21)
    chunks = orc.chunk_het_matrix(het, 500, 50_000)
    _, xs, pattern = orc.synth_particles(16, 48, seed=8)
    cases = [("time_S1", [3]), ("chunks_S5", [8, 1, 4, 4, 0]), ("chunks_S9", list(range(9)))]
    if world > 2:
        cases.append(("time_S3", [8, 2, 6]))
    mp.spawn(_worker, args=(world, 29561, chunks, xs, pattern, cases, str(tmp_path)), nprocs=world, join=True)
    kern = _PSMCKernelBase(16, chunks, overlap=500)
    x = torch.tensor(xs, dtype=torch.float64, device="cuda:0")
    for name, inds in cases:
        got = np.load(tmp_path / f"{name}.npz")
        value, grad = kern.hmm_term(x, pattern, 1e-2, torch.tensor(inds, device="cuda:0"), 500, weight=2.5)
        want_v, want_g = value.cpu().numpy(), grad.cpu().numpy()
        if name.startswith("time"):
            assert "time-sharded" in str(got["kernel"]), str(got["kernel"])
        np.testing.assert_allclose(got["value"], want_v, rtol=1e-6)
        scale = np.abs(want_g).max(-1, keepdims=True)
        assert np.all(np.abs(got["grad"] - want_g) <= 2e-4 * np.abs(want_g) + 1e-5 * scale), name
    from phlash_b200 import model

    want = float(model.elpd_hmm_term(model.elpd_kernel(16, chunks[:2, 500:30_500]), x, pattern, 1e-2))
    np.testing.assert_allclose(float(np.load(tmp_path / "elpd.npz")["elpd"]), want, rtol=1e-9)
