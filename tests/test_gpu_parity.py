"""Parity of the CUDA path (through the C ABI / PSMCKernel) against the fp64 oracle.
Tolerances: BASELINE.json north_star - loglik 1e-5 relative, gradients 1e-4 relative in fp32.
Mirrors the reference's tests/test_gpu.py and tests/test_model.py."""

import numpy as np
import pytest

from conftest import fixture_data
from oracle import c_oracle, psmc_oracle as orc

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-5
GRAD_RTOL = 1e-4


def grad_close(got, want, rtol, what=""):
    """Relative to each entry, with an absolute floor tied to the largest entry of the same row
    (entries many orders of magnitude below the row's scale carry no information in fp32)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max(axis=-1, keepdims=True)
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + rtol * 1e-3 * scale
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} entries off; worst rel {np.max(err / np.maximum(np.abs(want), 1e-300)):.3e}"


def make_kernel(M, data, double_precision=False, T=0):
    from phlash_b200.gpu import PSMCKernel

    k = PSMCKernel(M=M, data=data, double_precision=double_precision, num_gpus=1)
    if T:
        k.set_threads_per_pair(T)
    return k


def oracle_eval(data, inds, pa, kernel_dtype=np.float32):
    """fp64 oracle on "the same inputs": the kernel's interface is FLOAT parameters (the reference
    casts with pa.astype(float32), gpu.py:226), so the oracle is evaluated at the parameter values
    the kernel actually receives.  (Rounding d = 1 - 5e-5 to fp32 alone moves d ll / d log d by
    6e-4 - a property of the fp32 interface, not of any kernel.)"""
    B, S = pa.shape[:2]
    rows = np.tile(np.asarray(inds), B)
    pa = np.asarray(pa).astype(kernel_dtype).astype(np.float64)
    ll, dlog = c_oracle.loglik_batch(data, rows, pa.reshape(B * S, 7, -1), grad=True)
    return ll.reshape(B, S), dlog.reshape(B, S, 7, -1)


@pytest.mark.parametrize("T", [0, 2, 4])
@pytest.mark.parametrize("missing", [False, True])
def test_reference_fixture_fp32(golden, seed, missing, T):
    """tests/test_gpu.py:44-64 of the reference: CUDA vs the HMM definition, with and without
    missing data; ll and every gradient leaf."""
    from phlash_b200.params import PSMCParams

    data, miss = fixture_data(seed)
    data = miss if missing else data
    pp = golden["dm16_pp"]
    kern = make_kernel(16, data, T=T)
    inds = np.arange(len(data))
    ll, dll = kern(PSMCParams.from_block(pp), inds, grad=True)
    ll_nograd = kern(PSMCParams.from_block(pp), inds, grad=False)
    ref_ll, ref_dlog = oracle_eval(data, inds, np.broadcast_to(pp, (1, len(inds), 7, 16)))
    np.testing.assert_allclose(ll, ref_ll[0], rtol=LL_RTOL)
    np.testing.assert_allclose(ll_nograd, ll, rtol=1e-12)  # test_eq_grad_nograd
    grad_close(dll.to_block(), ref_dlog[0], GRAD_RTOL, f"T={T}")
    # the golden ll of the reference's own psmc_ll for rows 0..2
    off = 3 if missing else 0
    np.testing.assert_allclose(ll[:3], golden[f"hmm_ll_s{seed}"][off : off + 3], rtol=LL_RTOL)


@pytest.mark.parametrize("T", [0, 4])
def test_reference_fixture_fp64(golden, seed, T):
    """double_precision=True against the oracle at the reference's own tolerance
    (tests/test_gpu.py:59-64: atol 1e-8, rtol 1e-5) - and much tighter."""
    from phlash_b200.params import PSMCParams

    _, data = fixture_data(seed)
    pp = golden["part_pp"][2]
    kern = make_kernel(16, data, double_precision=True, T=T)
    inds = np.array([4, 0, 9, 4])
    ll, dll = kern(PSMCParams.from_block(pp), inds, grad=True)
    ref_ll, ref_dlog = oracle_eval(data, inds, np.broadcast_to(pp, (1, 4, 7, 16)), np.float64)
    np.testing.assert_allclose(ll, ref_ll[0], rtol=1e-12)
    np.testing.assert_allclose(dll.to_block(), ref_dlog[0], rtol=1e-8, atol=1e-12)


# --------------------------------------------------------------------------------------------
# wider shapes
# --------------------------------------------------------------------------------------------
def random_data(rng, n, length, het=0.07, miss=0.02):
    data = (rng.uniform(size=(n, length)) < het).astype(np.int8)
    data[rng.uniform(size=(n, length)) < miss] = -1
    data[:, 0] = np.maximum(data[:, 0], 0)  # no all-missing rows
    return data


@pytest.mark.parametrize("M,T", [(4, 0), (8, 0), (8, 2), (16, 1), (32, 0), (32, 2), (32, 8), (64, 0), (64, 4), (64, 16)])
def test_other_state_counts_fp32(M, T):
    """M = 32 / 64 are BASELINE.json configs 4 and 5; the reference cannot run them end to end
    (params.py:35) so the oracle is the only comparator."""
    rng = np.random.default_rng(M)
    data = random_data(rng, 6, 777)
    pps, _, _ = orc.synth_particles(M, 5, seed=M)
    kern = make_kernel(M, data, T=T)
    inds = np.array([5, 0, 3])
    pa = np.broadcast_to(pps[:, None], (5, 3, 7, M)).copy()
    from phlash_b200.params import PSMCParams

    ll, dll = kern(PSMCParams.from_block(pa), inds, grad=True)
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dll.to_block(), ref_dlog, GRAD_RTOL, f"M={M} T={T}")
    np.testing.assert_allclose(kern(PSMCParams.from_block(pa), inds, grad=False), ll, rtol=1e-6)


@pytest.mark.parametrize("M", [16, 32])
def test_other_state_counts_fp64(M):
    rng = np.random.default_rng(100 + M)
    data = random_data(rng, 4, 501)
    pps, _, _ = orc.synth_particles(M, 3, seed=M + 1)
    kern = make_kernel(M, data, double_precision=True)
    inds = np.array([1, 1, 3, 0])
    pa = np.broadcast_to(pps[:, None], (3, 4, 7, M)).copy()
    from phlash_b200.params import PSMCParams

    ll, dll = kern(PSMCParams.from_block(pa), inds, grad=True)
    ref_ll, ref_dlog = oracle_eval(data, inds, pa, np.float64)
    np.testing.assert_allclose(ll, ref_ll, rtol=1e-12)
    np.testing.assert_allclose(dll.to_block(), ref_dlog, rtol=1e-8, atol=1e-13)


def test_every_pair_has_its_own_parameters():
    """pa [B, S, 7, M] with all B*S blocks different (the general contract of gpu.py:182-213), and
    per-pair pi the way model.py:55 builds it."""
    rng = np.random.default_rng(3)
    data = random_data(rng, 5, 400)
    pps, _, _ = orc.synth_particles(16, 12, seed=9)
    pa = pps.reshape(3, 4, 7, 16).copy()
    inds = np.array([4, 2, 2, 0])
    kern = make_kernel(16, data)
    from phlash_b200.params import PSMCParams

    ll, dll = kern(PSMCParams.from_block(pa), inds, grad=True)
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dll.to_block(), ref_dlog, GRAD_RTOL, "per-pair")
    # shared rows + per-pair pi through the dedicated entry point
    pis = rng.dirichlet(np.ones(16), size=(3, 4))
    pa2 = np.broadcast_to(pps[:3, None], (3, 4, 7, 16)).copy()
    pa2[:, :, 6] = pis
    base = kern.gpu_kernels[0]
    ll_a, dlog_a = base.evaluate(pa2, inds, True)
    ll_b, dlog_b = base.evaluate_shared(pps[:3, :6], pis, inds, True)
    np.testing.assert_array_equal(ll_a, ll_b)
    np.testing.assert_array_equal(dlog_a, dlog_b)
    ref_ll, ref_dlog = oracle_eval(data, inds, pa2)
    np.testing.assert_allclose(ll_b, ref_ll, rtol=LL_RTOL)
    grad_close(dlog_b, ref_dlog, GRAD_RTOL, "shared+pi")


@pytest.mark.parametrize("length", [1, 2, 3, 4, 5, 15, 16, 17, 31, 33, 1003])
def test_ragged_lengths(golden, length):
    """Row lengths that are not multiples of the 4-site rescaling block, the 16-site checkpoint
    segment or the 16-byte row pitch."""
    rng = np.random.default_rng(length)
    data = random_data(rng, 3, length, het=0.2, miss=0.1)
    pp = golden["part_pp"][4]
    from phlash_b200.params import PSMCParams

    for T in (0, 4):
        kern = make_kernel(16, data, T=T)
        ll, dll = kern(PSMCParams.from_block(pp), np.arange(3), grad=True)
        ref_ll, ref_dlog = oracle_eval(data, np.arange(3), np.broadcast_to(pp, (1, 3, 7, 16)))
        np.testing.assert_allclose(ll, ref_ll[0], rtol=LL_RTOL, atol=1e-6)
        grad_close(dll.to_block(), ref_dlog[0], GRAD_RTOL, f"L={length}")


def test_argument_shapes_and_scalar_index(golden):
    """pp leaves [M] with a scalar index, [S, M] and [B, S, M] (gpu.py:186-213, 319-325)."""
    from phlash_b200.params import PSMCParams

    data, _ = fixture_data(0)
    pp = golden["dm16_pp"]
    kern = make_kernel(16, data)
    ll0 = kern.loglik(PSMCParams.from_block(pp), 3)
    assert np.ndim(ll0) == 0
    np.testing.assert_allclose(ll0, golden["hmm_ll_s0"][0] if False else orc.psmc_ll(pp, data[3])[1], rtol=LL_RTOL)
    ll, dll = kern(PSMCParams.from_block(pp), 3, grad=True)
    assert np.ndim(ll) == 0 and dll.b.shape == (16,)
    ll_s, dll_s = kern(PSMCParams.from_block(np.stack([pp, pp])), np.array([3, 4]), grad=True)
    assert ll_s.shape == (2,) and dll_s.pi.shape == (2, 16)
    np.testing.assert_allclose(ll_s[0], ll, rtol=1e-12)


def test_input_validation(golden):
    """The reference's assertions (gpu.py:106-113, 197-199, 214)."""
    from phlash_b200.gpu import PSMCKernel
    from phlash_b200.params import PSMCParams

    data, _ = fixture_data(1)
    pp = golden["dm16_pp"]
    kern = make_kernel(16, data)
    with pytest.raises(AssertionError):
        kern(PSMCParams.from_block(pp), np.array([0, 10]), grad=True)  # 10 == N
    with pytest.raises(AssertionError):
        kern(PSMCParams.from_block(pp), np.array([-1]), grad=False)
    bad = pp.copy()
    bad[1, 3] = np.nan
    with pytest.raises(AssertionError):
        kern(PSMCParams.from_block(bad), 0, grad=True)
    allmiss = data.copy()
    allmiss[2] = -1
    with pytest.raises(AssertionError):
        PSMCKernel(M=16, data=allmiss)
    with pytest.raises(AssertionError):
        PSMCKernel(M=16, data=data.astype(np.int32))
    clipped = data.copy()
    clipped[0, :5] = 3  # values > 1 are clipped to 1 (gpu.py:108-110)
    k2 = make_kernel(16, clipped)
    np.testing.assert_allclose(
        k2.loglik(PSMCParams.from_block(pp), 0), orc.psmc_ll(pp, clipped[0].clip(-1, 1))[1], rtol=LL_RTOL
    )


def test_scratch_grows_between_calls(golden):
    """The reference sizes its device buffers by the first call (gpu.py:222-237); ours must grow."""
    from phlash_b200.params import PSMCParams

    data, _ = fixture_data(2)
    pp = golden["dm16_pp"]
    kern = make_kernel(16, data)
    small = kern(PSMCParams.from_block(pp), np.array([1]), grad=True)
    pa = np.broadcast_to(pp, (9, 10, 7, 16)).copy()
    ll, dll = kern(PSMCParams.from_block(pa), np.arange(10), grad=True)
    np.testing.assert_allclose(ll[4, 1], small[0][0], rtol=1e-12)
    np.testing.assert_allclose(dll.d[8, 1], small[1].d[0], rtol=1e-6)


def test_device_buffer_entry_matches_host_entry(golden):
    import torch

    data, _ = fixture_data(0)
    pps, _, _ = orc.synth_particles(16, 7, seed=4)
    kern = make_kernel(16, data)
    base = kern.gpu_kernels[0]
    inds = np.array([0, 5, 9, 5, 2])
    pa = np.broadcast_to(pps[:, None], (7, 5, 7, 16)).copy()
    ll_h, dlog_h = base.evaluate(pa, inds, True)
    dev = torch.device("cuda:0")
    p6 = torch.tensor(pps[:, :6], dtype=torch.float32, device=dev).contiguous()
    pi = torch.tensor(pps[:, 6], dtype=torch.float32, device=dev).contiguous()
    ll_d, dlog_d = base.evaluate_device(p6, pi, torch.tensor(inds, device=dev), True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ll_d.cpu().numpy(), ll_h)
    np.testing.assert_array_equal(dlog_d.cpu().numpy(), dlog_h)


# --------------------------------------------------------------------------------------------
# full chunk length (BASELINE.json config 2: 50 000 + bins per chunk)
# --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def long_case():
    het = orc.synth_het_matrix(1, 420_000, seed=0)
    chunks = orc.chunk_het_matrix(het, 500, 50_000)  # 8 chunks x 50 500
    pps, _, _ = orc.synth_particles(16, 24, seed=0)
    return chunks[:, 500:].copy(), pps


def test_full_length_chunks_against_oracle(long_case):
    """Every chunk of a 420 000-bin row at the benchmark geometry (50 000-bin chunks).  Chunks
    0..7 are ordinary; chunk 8 is the padded last chunk of the contig: 19 500 observed bins
    followed by 30 500 missing ones (data.py:37-61 pads with -1)."""
    data, pps = long_case
    kern = make_kernel(16, data)
    inds = np.arange(data.shape[0])
    pa = np.broadcast_to(pps[:6, None], (6, len(inds), 7, 16)).copy()
    ll, dlog = kern.gpu_kernels[0].evaluate(pa, inds, True)
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    assert (data[8] < 0).sum() > 30_000 and (data[:8] < 0).mean() < 0.05
    grad_close(dlog[:, :8], ref_dlog[:, :8], GRAD_RTOL, "L=50000")
    # Through >= 1e4 consecutive missing bins an fp32 adjoint vector sits at a floating-point fixed
    # point and the transition rows of the gradient lose up to ~4e-4 relative (DESIGN.md,
    # "Accuracy").  Such rows are marked at construction and scored with double arithmetic:
    assert kern.gpu_kernels[0].num_escalated_rows == 1
    grad_close(dlog[:, 8:], ref_dlog[:, 8:], GRAD_RTOL, "padded last chunk")
    # ... and this is what the plain fp32 path gives on that row (emission / pi rows unaffected)
    kern.gpu_kernels[0].set_precision_escalation(False)
    ll32, dlog32 = kern.gpu_kernels[0].evaluate(pa[:, 8:], inds[8:], True)
    np.testing.assert_allclose(ll32, ref_ll[:, 8:], rtol=LL_RTOL)
    grad_close(dlog32, ref_dlog[:, 8:], 1e-3, "padded last chunk, fp32 only")
    grad_close(dlog32[:, :, 4:], ref_dlog[:, 8:, 4:], GRAD_RTOL, "padded last chunk, fp32 only, emission and pi rows")
    k64 = make_kernel(16, data, double_precision=True)
    ll64, dlog64 = k64.gpu_kernels[0].evaluate(pa[:2, 8:], inds[8:], True)
    ref_ll64, ref_dlog64 = oracle_eval(data, inds[8:], pa[:2, 8:], np.float64)
    np.testing.assert_allclose(ll64, ref_ll64, rtol=1e-10)
    np.testing.assert_allclose(dlog64, ref_dlog64, rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("M,T", [(16, 1), (32, 2)])
def test_headline_throughput_kernel_full_length(long_case, M, T):
    """THE kernel bench.py times - the thread-per-pair throughput kernel (MT = 16; non-segment mode,
    checkpoints + recompute, 1024-site fp32 windows flushed into fp64 slots) - at the benchmark's chunk
    length, against the fp64 oracle.  threads_per_pair forces the throughput path (no parallel-in-time /
    store-all dispatch), so all 6 250 checkpoint segments and 48 window flushes of a 50 000-bin chunk run.
    8 particles x all 9 chunks including the padded one (scored by the double build), and the same
    layout with two lanes per pair at M = 32.  Mirrors tests/test_gpu.py:59-64 of the reference."""
    data, pps16 = long_case
    pps = pps16 if M == 16 else orc.synth_particles(M, 8, seed=M)[0]
    kern = make_kernel(M, data, T=T).gpu_kernels[0]
    inds = np.arange(data.shape[0])
    pa = np.broadcast_to(pps[:8, None], (8, len(inds), 7, M)).copy()
    ll, dlog = kern.evaluate(pa, inds, True)
    assert f"psmc_loglik_kernel<float,MT=16,T={T},K=8,grad" in kern.last_kernel_name, kern.last_kernel_name
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dlog, ref_dlog, GRAD_RTOL, f"throughput kernel M={M} T={T} L=50000")
    # forward-only build of the same layout
    ll_fwd = kern.evaluate(pa, inds, False)
    assert f"T={T}" in kern.last_kernel_name and "fwd" in kern.last_kernel_name
    np.testing.assert_allclose(ll_fwd, ref_ll, rtol=LL_RTOL)


def test_full_length_invariants(long_case):
    """Properties that hold for any parameters, checked on every pair without the oracle:
    the posterior of each site sums to one, so  sum_m (dlog_e0 + dlog_e1)[m] = #non-missing sites,
    sum_m (dlog_b + dlog_d + dlog_v)[m] = L, and sum_m dlog_pi[m] = 1."""
    data, pps = long_case
    kern = make_kernel(16, data)
    inds = np.arange(data.shape[0])
    pa = np.broadcast_to(pps[:, None], (len(pps), len(inds), 7, 16)).copy()
    ll, dlog = kern.gpu_kernels[0].evaluate(pa, inds, True)
    assert np.isfinite(ll).all() and np.isfinite(dlog).all() and (ll < 0).all()
    n_obs = (data >= 0).sum(axis=1)
    emis_mass = dlog[:, :, 4].sum(-1, dtype=np.float64) + dlog[:, :, 5].sum(-1, dtype=np.float64)
    np.testing.assert_allclose(emis_mass, np.broadcast_to(n_obs, emis_mass.shape), rtol=2e-5)
    trans_mass = sum(dlog[:, :, r].sum(-1, dtype=np.float64) for r in (0, 1, 3))
    np.testing.assert_allclose(trans_mass, data.shape[1], rtol=2e-5)
    np.testing.assert_allclose(dlog[:, :, 6].sum(-1, dtype=np.float64), 1.0, rtol=2e-5)
    # structural zeros stay exactly zero
    assert np.all(dlog[:, :, 0, -1] == 0) and np.all(dlog[:, :, 2, -1] == 0) and np.all(dlog[:, :, 3, 0] == 0)
    # forward-only path gives the same likelihood
    ll2 = kern.gpu_kernels[0].evaluate(pa, inds, False)
    np.testing.assert_allclose(ll2, ll, rtol=1e-6)


def test_forward_only_uniform_register_kernel():
    """Forward-only evaluation of a large minibatch at M = 16 with shared parameter rows runs with the parameters
    in uniform registers (psmc_uniform.cuh): more particles than constant-bank slots (two batches), a chunk count
    that is not a multiple of 32, missing data; against the oracle and against the gradient kernel's ll."""
    rng = np.random.default_rng(77)
    data = random_data(rng, 70, 1801, het=0.08, miss=0.03)
    pps, _, _ = orc.synth_particles(16, 131, seed=5)
    inds = np.arange(70)[::-1].copy()
    pa = np.broadcast_to(pps[:, None], (131, 70, 7, 16)).copy()
    kern = make_kernel(16, data).gpu_kernels[0]
    ll = kern.evaluate(pa, inds, False)
    assert "psmc_uniform_forward_kernel" in kern.last_kernel_name, kern.last_kernel_name
    ref_ll, _ = oracle_eval(data, inds, pa)
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    ll_g, _ = kern.evaluate(pa, inds, True)
    np.testing.assert_allclose(ll, ll_g, rtol=1e-6)
    # per-pair parameter rows cannot be warp-uniform: the register-parameter kernel scores them
    pa2 = pa.copy()
    pa2[3, 5, 1] *= 0.999
    ll2 = kern.evaluate(pa2, inds, False)
    assert "psmc_uniform" not in kern.last_kernel_name
    np.testing.assert_allclose(np.delete(ll2.ravel(), 3 * 70 + 5), np.delete(ll.ravel(), 3 * 70 + 5), rtol=1e-6)


def test_empty_batches(golden):
    """B = 0 or S = 0: nothing to do, correctly shaped empty results (no launch)."""
    data, _ = fixture_data(0)
    kern = make_kernel(16, data).gpu_kernels[0]
    n0 = kern.launch_count
    ll, dlog = kern.evaluate(np.zeros((3, 0, 7, 16)), np.zeros(0, dtype=np.int64), True)
    assert ll.shape == (3, 0) and dlog.shape == (3, 0, 7, 16)
    ll = kern.evaluate(np.zeros((0, 2, 7, 16)), np.array([1, 2]), False)
    assert ll.shape == (0, 2)
    assert kern.launch_count == n0


def test_device_entry_reports_bad_index_at_sync(golden):
    """The device-buffer entry cannot check indices on the host: the kernel flags them, the
    affected pairs get NaN and phb_sync() raises (include/phlash_b200.h)."""
    import torch

    data, _ = fixture_data(0)
    pps = golden["part_pp"][:2]
    kern = make_kernel(16, data).gpu_kernels[0]
    dev = torch.device("cuda:0")
    p6 = torch.tensor(pps[:, :6], dtype=torch.float32, device=dev).contiguous()
    pi = torch.tensor(pps[:, 6], dtype=torch.float32, device=dev).contiguous()
    inds = torch.tensor([1, 10, 3], device=dev)  # 10 == N is out of range
    ll, _ = kern.evaluate_device(p6, pi, inds, True)
    with pytest.raises(AssertionError):
        kern.sync()
    ll = ll.cpu().numpy()
    assert np.isnan(ll[:, 1]).all() and np.isfinite(ll[:, [0, 2]]).all()
    kern.sync()  # the flag is cleared once reported


@pytest.mark.parametrize("M", [4, 8, 16, 32, 64])
def test_store_all_kernel(M):
    """The small-minibatch gradient kernel (keeps every forward vector instead of recomputing)
    against the oracle and against the checkpointing kernel."""
    rng = np.random.default_rng(200 + M)
    data = random_data(rng, 5, 1003, het=0.1, miss=0.05)
    pps, _, _ = orc.synth_particles(M, 6, seed=M + 3)
    pa = np.broadcast_to(pps[:, None], (6, 3, 7, M)).copy()
    inds = np.array([4, 0, 2])
    kern = make_kernel(M, data).gpu_kernels[0]
    kern.set_store_all(1)
    ll, dlog = kern.evaluate(pa, inds, True)
    assert "storeall" in kern.last_kernel_name
    ref_ll, ref_dlog = oracle_eval(data, inds, pa)
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL)
    grad_close(dlog, ref_dlog, GRAD_RTOL, f"store-all M={M}")
    kern.set_store_all(0)
    ll2, dlog2 = kern.evaluate(pa, inds, True)
    assert "storeall" not in kern.last_kernel_name
    np.testing.assert_allclose(ll2, ll, rtol=1e-6)
    grad_close(dlog2, dlog, 1e-4, "store-all vs checkpointing")


@pytest.mark.parametrize("length", [1, 3, 4, 5, 17, 64, 130])
def test_store_all_ragged_lengths(golden, length):
    rng = np.random.default_rng(length + 77)
    data = random_data(rng, 3, length, het=0.2, miss=0.1)
    pp = golden["part_pp"][4]
    pa = np.broadcast_to(pp, (2, 3, 7, 16)).copy()
    kern = make_kernel(16, data).gpu_kernels[0]
    kern.set_store_all(1)
    ll, dlog = kern.evaluate(pa, np.arange(3), True)
    ref_ll, ref_dlog = oracle_eval(data, np.arange(3), pa)
    np.testing.assert_allclose(ll, ref_ll, rtol=LL_RTOL, atol=1e-6)
    grad_close(dlog, ref_dlog, GRAD_RTOL, f"store-all L={length}")


def test_store_all_with_fused_warmup(golden):
    """The two-launch warm-up evaluation goes through the same dispatcher (second launch subtracts)."""
    chunks, inds = golden["model_chunks"], golden["model_inds"]
    pps = golden["part_pp"][:4].astype(np.float32).astype(np.float64)
    from phlash_b200.gpu import _PSMCKernelBase

    k1 = _PSMCKernelBase(16, chunks)
    k1.set_store_all(1)
    a = k1.evaluate_warmup(pps, inds, 50, True)
    k0 = _PSMCKernelBase(16, chunks)
    k0.set_store_all(0)
    b = k0.evaluate_warmup(pps, inds, 50, True)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-6)
    scale = np.abs(b[1]).max(-1, keepdims=True)
    assert np.all(np.abs(a[1] - b[1]) <= 1e-4 * np.abs(b[1]) + 1e-5 * scale)


def test_matrix_larger_than_2_gib_uses_64_bit_offsets():
    """BASELINE configs 4 / 5 hold 1.3 - 50 GB of observations per GPU: row offsets must be 64-bit.
    The last rows of a 2.3 GB matrix (46 000 x 50 000) score exactly like the same rows in a small one."""
    from phlash_b200.gpu import _PSMCKernelBase

    L, reps = 50_000, 11_500
    base = orc.synth_het_matrix(4, L, seed=3).astype(np.int8)
    base[:, 0] = np.where(base[:, 0] < 0, 0, base[:, 0])
    big = np.tile(base, (reps, 1))  # row i == base[i % 4]
    assert big.nbytes > 2**31
    pps, _, _ = orc.synth_particles(16, 3, seed=1)
    small = _PSMCKernelBase(16, base)
    large = _PSMCKernelBase(16, big)
    inds_big = np.array([len(big) - 1, len(big) - 2, 0, len(big) // 2 + 1])
    inds_small = inds_big % 4
    pa = np.broadcast_to(pps[:, None], (3, 4, 7, 16)).copy()
    for grad in (True, False):
        got = large.evaluate(pa, inds_big, grad)
        want = small.evaluate(pa, inds_small, grad)
        if grad:
            np.testing.assert_array_equal(got[0], want[0])
            np.testing.assert_array_equal(got[1], want[1])
        else:
            np.testing.assert_array_equal(got, want)
    assert large.num_escalated_rows == 0
