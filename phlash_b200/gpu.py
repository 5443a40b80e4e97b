"""Host-side mirror of ``phlash.gpu`` (reference: src/phlash/gpu.py:101-438): the same
``PSMCKernel`` constructor / ``__call__`` / ``loglik`` / ``float_type`` surface, backed by the
sm_100a kernels behind the C ABI in ``include/phlash_b200.h``.

Differences from the reference, all deliberate:
  * no NVRTC: the kernels are compiled ahead of time for sm_100a;
  * no fallback: a missing library, a missing GPU or a CUDA failure raises;
  * scratch buffers grow with the call (the reference sizes them by the first call, gpu.py:222-237);
  * ``num_gpus > 1`` splits the *chunk* axis S across devices and joins on that axis (the
    reference joins on the particle axis, gpu.py:425-429, which is only right for B == 1);
  * the forward-only path gives every pair its own parameter block (the reference's ``loglik``
    kernel silently uses the parameters of s == 0 for the whole block, gpu.py:548-551).
"""

from __future__ import annotations

import ctypes
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from phlash_b200 import _native
from phlash_b200.params import PSMCParams


class CudaError(RuntimeError):
    """Raised for CUDA runtime failures (reference: gpu.py:23-32)."""


def _check(rc: int) -> None:
    if rc == _native.PHB_OK:
        return
    msg = _native.last_error()
    if rc == _native.PHB_E_NOMEM:
        raise MemoryError(msg)
    if rc == _native.PHB_E_CUDA:
        raise CudaError(msg)
    # PHB_E_INVALID / PHB_E_DATA: the reference signals these with bare asserts
    raise AssertionError(msg)


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


class _PSMCKernelBase:
    "PSMC kernel running on a single GPU (reference: gpu.py:101-325)"

    def __init__(self, M: int, data: np.ndarray, double_precision: bool = False, device: int = 0, overlap: int = 0):
        """``overlap`` > 0: ``data`` are FULL chunks [N, overlap + L] (for the fused warm-up entries); the
        "every row has an observation" check then covers the data part [overlap:], which is what the
        reference's constructor sees after mcmc.py:203 has split the warm-up columns off."""
        data = np.asarray(data)
        assert data.ndim == 2
        assert data.dtype == np.int8
        # The reference's checks (min >= -1, clip to <= 1, every row observed; gpu.py:106-113) run on the
        # DEVICE inside phb_create_chunks, after the copy: the host never walks a matrix that can be tens
        # of GB (BASELINE config 5), and no clipped host copy is made.
        data = np.ascontiguousarray(data)
        self.double_precision = bool(double_precision)
        self._N, self._L = data.shape
        self._M = int(M)
        self._lib = _native.lib()
        handle = ctypes.c_void_p()
        _check(
            self._lib.phb_create_chunks(
                self._M, _ptr(data), self._N, self._L, int(overlap), int(self.double_precision), int(device),
                ctypes.byref(handle)
            )
        )
        self._handle = handle
        self.device = int(device)

    @classmethod
    def from_contig(cls, M: int, het_matrix: np.ndarray, overlap: int, chunk_size: int,
                    double_precision: bool = False, device: int = 0) -> "_PSMCKernelBase":
        """Kernel object on the chunks of one binned contig, cut ON THE DEVICE with the geometry of
        _chunk_het_matrix (data.py:37-61).  The object holds full chunks [N, overlap + chunk_size]:
        use evaluate_warmup(..., overlap=overlap)."""
        het = np.asarray(het_matrix)
        assert het.ndim == 2
        if het.dtype != np.int8:  # counts of any integer type: bring them to int8 without wrapping
            assert het.min() >= -1
            het = het.clip(-1, 1).astype(np.int8)
        het = np.ascontiguousarray(het)  # (int8 input is validated and clipped on the device)
        self = cls.__new__(cls)
        self.double_precision = bool(double_precision)
        self._M = int(M)
        self._lib = _native.lib()
        handle = ctypes.c_void_p()
        _check(
            self._lib.phb_create_from_contig(
                self._M, _ptr(het), het.shape[0], het.shape[1], int(overlap), int(chunk_size),
                int(self.double_precision), int(device), ctypes.byref(handle)
            )
        )
        self._handle = handle
        self.device = int(device)
        self._N = int(self._lib.phb_num_rows(handle))
        self._L = int(self._lib.phb_row_length(handle))
        return self

    def download_data(self) -> np.ndarray:
        """The resident observation matrix [N, L] copied back to the host."""
        out = np.empty((self._N, self._L), dtype=np.int8)
        _check(self._lib.phb_download_data(self._handle, _ptr(out)))
        return out

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            self._lib.phb_destroy(handle)
            self._handle = None

    @property
    def float_type(self):
        return np.float64 if self.double_precision else np.float32

    def set_threads_per_pair(self, t: int) -> None:
        _check(self._lib.phb_set_threads_per_pair(self._handle, int(t)))

    def set_store_all(self, mode: int) -> None:
        """-1 auto, 0 never, 1 always: the store-all gradient kernel for small minibatches."""
        _check(self._lib.phb_set_store_all(self._handle, int(mode)))

    def set_parallel_in_time(self, mode: int) -> None:
        """-1 auto, 0 never, 1 whenever possible: forward-only evaluation of few, long pairs through
        segment transfer operators (include/phlash_b200.h)."""
        _check(self._lib.phb_set_parallel_in_time(self._handle, int(mode)))

    def set_precision_escalation(self, enabled: bool) -> None:
        """Rows with a long run of identical observations are scored with double arithmetic (default
        on for single-precision objects; see include/phlash_b200.h)."""
        _check(self._lib.phb_set_precision_escalation(self._handle, int(bool(enabled))))

    @property
    def num_escalated_rows(self) -> int:
        return int(self._lib.phb_num_escalated_rows(self._handle))

    @property
    def last_kernel_ms(self) -> float:
        return float(self._lib.phb_last_kernel_ms(self._handle))

    @property
    def last_kernel_name(self) -> str:
        return self._lib.phb_last_kernel_name(self._handle).decode()

    @property
    def launch_count(self) -> int:
        return int(self._lib.phb_launch_count(self._handle))

    def evaluate(self, pa: np.ndarray, inds: np.ndarray, grad: bool, ll_out=None, dlog_out=None):
        """pa [B, S, 7, M] (any float dtype), inds [S] -> ll [B, S] (, dlog [B, S, 7, M]).
        ``ll_out`` / ``dlog_out``: optional preallocated (e.g. pinned) result arrays."""
        M = self._M
        B, S = pa.shape[:2]
        assert pa.shape == (B, S, 7, M)
        assert inds.shape == (S,)
        assert np.all(0 <= inds) & np.all(inds < self._N), f"0 <= {inds.min()=} < {inds.max()=} < N"
        # (the finiteness check of gpu.py:214 is done on the device by phb_loglik_host)
        pa = np.ascontiguousarray(pa, dtype=self.float_type)
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        ll = np.zeros([B, S], dtype=np.float64) if ll_out is None else ll_out
        dlog = None
        if grad:
            dlog = np.zeros([B, S, 7, M], dtype=self.float_type) if dlog_out is None else dlog_out
            assert dlog.shape == (B, S, 7, M) and dlog.dtype == self.float_type and dlog.flags.c_contiguous
        assert ll.shape == (B, S) and ll.dtype == np.float64 and ll.flags.c_contiguous
        _check(
            self._lib.phb_loglik_host(
                self._handle, _ptr(pa), _ptr(inds), B, S, int(grad), _ptr(ll), _ptr(dlog) if grad else None
            )
        )
        return (ll, dlog) if grad else ll

    def evaluate_shared(self, params6: np.ndarray, pi: np.ndarray, inds: np.ndarray, grad: bool):
        """params6 [B, 6, M] shared by the S chunks of a particle, pi [B, S, M] or [B, M]."""
        M = self._M
        B = params6.shape[0]
        S = inds.shape[0]
        assert params6.shape == (B, 6, M)
        assert pi.shape in ((B, S, M), (B, M))
        assert np.all(0 <= inds) & np.all(inds < self._N)
        assert np.isfinite(params6).all() and np.isfinite(pi).all(), "not all parameters finite"
        params6 = np.ascontiguousarray(params6, dtype=self.float_type)
        pi_c = np.ascontiguousarray(pi, dtype=self.float_type)
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        ll = np.zeros([B, S], dtype=np.float64)
        dlog = np.zeros([B, S, 7, M], dtype=self.float_type) if grad else None
        _check(
            self._lib.phb_loglik_shared_host(
                self._handle, _ptr(params6), _ptr(pi_c), int(pi.ndim == 3), _ptr(inds), B, S, int(grad),
                _ptr(ll), _ptr(dlog) if grad else None,
            )
        )
        return (ll, dlog) if grad else ll

    def evaluate_warmup(self, params7: np.ndarray, inds: np.ndarray, overlap: int, grad: bool):
        """Fused warm-up evaluation (model.py:50-57) on a kernel built from FULL chunks
        [N, overlap + L]: params7 [B, 7, M] (pi row = the particle's stationary pi).  Returns
        ll [B, S] = log p(chunk | warm-up) and, if grad, dlog [B, S, 7, M] w.r.t. the particle's
        own rows; both additive over the chunk axis."""
        M = self._M
        B, S = params7.shape[0], inds.shape[0]
        assert params7.shape == (B, 7, M)
        assert 0 <= overlap < self._L
        assert np.all(0 <= inds) & np.all(inds < self._N)
        assert np.isfinite(params7).all(), "not all parameters finite"
        params7 = np.ascontiguousarray(params7, dtype=self.float_type)
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        ll = np.zeros([B, S], dtype=np.float64)
        dlog = np.zeros([B, S, 7, M], dtype=self.float_type) if grad else None
        _check(
            self._lib.phb_loglik_warmup_host(
                self._handle, _ptr(params7), _ptr(inds), B, S, int(overlap), int(grad), _ptr(ll),
                _ptr(dlog) if grad else None,
            )
        )
        return (ll, dlog) if grad else ll

    def evaluate_warmup_device(self, params7, inds, overlap: int, grad: bool, ll=None, dlog=None, stream=None):
        """Device-buffer form of evaluate_warmup (torch CUDA tensors; asynchronous)."""
        import torch

        M = self._M
        B, S = int(params7.shape[0]), int(inds.shape[0])
        tdtype = torch.float64 if self.double_precision else torch.float32
        assert params7.is_cuda and inds.is_cuda and params7.device.index == self.device
        assert params7.shape == (B, 7, M) and params7.dtype == tdtype and params7.is_contiguous()
        assert inds.dtype == torch.int64 and inds.is_contiguous()
        if ll is None:
            ll = torch.empty((B, S), dtype=torch.float64, device=params7.device)
        if grad and dlog is None:
            dlog = torch.empty((B, S, 7, M), dtype=tdtype, device=params7.device)
        if stream is None:
            stream = torch.cuda.current_stream(params7.device).cuda_stream
        _check(
            self._lib.phb_loglik_warmup_device(
                self._handle, params7.data_ptr(), inds.data_ptr(), B, S, int(overlap), int(grad), ll.data_ptr(),
                dlog.data_ptr() if grad else None, ctypes.c_void_p(stream),
            )
        )
        return ll, (dlog if grad else None)

    # ---- parameter construction on the device (params.py:32-55, 94-131; transition.py; size_history.py)
    def params_from_particles(self, x, pattern: str, theta: float, stream=None):
        """x: torch float64 CUDA tensor [B, P] of flattened particles (t_tr[2], c_tr, rho_over_theta_tr).
        Returns the [B, 7, M] parameter blocks (kernel float type) as a torch tensor on the device."""
        import torch

        from phlash_b200.params import parse_pattern

        widths = np.ascontiguousarray(parse_pattern(pattern), dtype=np.int32)
        B, P = int(x.shape[0]), int(x.shape[1])
        assert P == 2 + len(widths) + 1, "particle length does not match the pattern"
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.device.index == self.device
        tdtype = torch.float64 if self.double_precision else torch.float32
        out = torch.empty((B, 7, self._M), dtype=tdtype, device=x.device)
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(
            self._lib.phb_params_from_particles(
                self._handle, x.data_ptr(), B, _ptr(widths), len(widths), float(theta), out.data_ptr(),
                ctypes.c_void_p(stream),
            )
        )
        return out

    def _term_args(self, x, pattern: str):
        import torch

        from phlash_b200.params import parse_pattern

        widths = np.ascontiguousarray(parse_pattern(pattern), dtype=np.int32)
        B, P = int(x.shape[0]), int(x.shape[1])
        assert P == 2 + len(widths) + 1, "particle length does not match the pattern"
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.device.index == self.device
        return widths, B, P

    def hmm_term_sums(self, x, pattern: str, theta: float, inds, overlap: int, grad: bool = True, stream=None):
        """Per-particle sums over (this rank's part of) the minibatch: [B, 1 + 7 M] float64 =
        (sum_s ll, sum_s d ll / d log theta), warm-up fused (include/phlash_b200.h, phb_hmm_term_sums_device).
        x: torch float64 CUDA [B, P]; inds: torch int64 CUDA [S] (may be empty)."""
        import torch

        widths, B, _ = self._term_args(x, pattern)
        S = int(inds.shape[0])
        assert inds.is_cuda and inds.dtype == torch.int64 and inds.is_contiguous()
        sums = torch.empty((B, 1 + 7 * self._M), dtype=torch.float64, device=x.device)
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(
            self._lib.phb_hmm_term_sums_device(
                self._handle, x.data_ptr(), B, _ptr(widths), len(widths), float(theta), inds.data_ptr() if S else None,
                S, int(overlap), int(grad), sums.data_ptr(), ctypes.c_void_p(stream),
            )
        )
        return sums

    def hmm_term_finish(self, x, pattern: str, theta: float, sums, weight: float = 1.0, grad: bool = True, stream=None):
        """(weight * l2 [B], weight * d l2 / d x [B, P] or None) from the (all-reduced) sums."""
        import torch

        widths, B, P = self._term_args(x, pattern)
        assert sums.shape == (B, 1 + 7 * self._M) and sums.dtype == torch.float64 and sums.is_contiguous() and sums.is_cuda
        value = torch.empty((B,), dtype=torch.float64, device=x.device)
        grad_x = torch.empty((B, P), dtype=torch.float64, device=x.device) if grad else None
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(
            self._lib.phb_hmm_term_finish_device(
                self._handle, x.data_ptr(), B, _ptr(widths), len(widths), float(theta), sums.data_ptr(), float(weight),
                value.data_ptr(), grad_x.data_ptr() if grad else None, ctypes.c_void_p(stream),
            )
        )
        return value, grad_x

    def hmm_term(self, x, pattern: str, theta: float, inds, overlap: int, weight: float = 1.0, grad: bool = True, stream=None):
        """The whole HMM term of log_density and its gradient w.r.t. the particles in ONE library call
        (phb_hmm_term_device): (weight * l2 [B], weight * d l2 / d x [B, P] or None)."""
        import torch

        widths, B, P = self._term_args(x, pattern)
        S = int(inds.shape[0])
        assert inds.is_cuda and inds.dtype == torch.int64 and inds.is_contiguous()
        value = torch.empty((B,), dtype=torch.float64, device=x.device)
        grad_x = torch.empty((B, P), dtype=torch.float64, device=x.device) if grad else None
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(
            self._lib.phb_hmm_term_device(
                self._handle, x.data_ptr(), B, _ptr(widths), len(widths), float(theta), inds.data_ptr() if S else None,
                S, int(overlap), float(weight), value.data_ptr(), grad_x.data_ptr() if grad else None,
                ctypes.c_void_p(stream),
            )
        )
        return value, grad_x

    def hmm_term_host(self, x: np.ndarray, pattern: str, theta: float, inds: np.ndarray, overlap: int,
                      weight: float = 1.0, grad: bool = True):
        """`hmm_term` with NumPy buffers, blocking (phb_hmm_term_host): x float64 [B, P], inds int [S] ->
        (weight * l2 [B], weight * d l2 / d x [B, P] or None).  No torch involved: the entry a
        `jax.pure_callback` binds."""
        from phlash_b200.params import parse_pattern

        widths = np.ascontiguousarray(parse_pattern(pattern), dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        B, P = x.shape
        assert P == 2 + len(widths) + 1, "particle length does not match the pattern"
        assert inds.ndim == 1 and (len(inds) == 0 or (inds.min() >= 0 and inds.max() < self._N))
        assert np.isfinite(x).all(), "not all particle coordinates finite"
        value = np.empty(B, dtype=np.float64)
        grad_x = np.empty((B, P), dtype=np.float64) if grad else None
        _check(
            self._lib.phb_hmm_term_host(
                self._handle, _ptr(x), B, _ptr(widths), len(widths), float(theta), _ptr(inds) if len(inds) else None,
                len(inds), int(overlap), float(weight), _ptr(value), _ptr(grad_x) if grad else None,
            )
        )
        return value, grad_x

    def params_vjp(self, x, pattern: str, theta: float, cotangent, stream=None):
        """cotangent [B, 7, M] = d l / d log(theta) (kernel float type).  Returns d l / d x [B, P]."""
        import torch

        from phlash_b200.params import parse_pattern

        widths = np.ascontiguousarray(parse_pattern(pattern), dtype=np.int32)
        B, P = int(x.shape[0]), int(x.shape[1])
        tdtype = torch.float64 if self.double_precision else torch.float32
        assert P == 2 + len(widths) + 1
        assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
        assert cotangent.shape == (B, 7, self._M) and cotangent.dtype == tdtype and cotangent.is_contiguous()
        out = torch.empty((B, P), dtype=torch.float64, device=x.device)
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(
            self._lib.phb_params_vjp(
                self._handle, x.data_ptr(), B, _ptr(widths), len(widths), float(theta), cotangent.data_ptr(),
                out.data_ptr(), ctypes.c_void_p(stream),
            )
        )
        return out

    # ---- device-resident entry (torch tensors are only used as device buffers here)
    def evaluate_device(self, params6, pi, inds, grad: bool, ll=None, dlog=None, stream=None):
        """Asynchronous evaluation on device buffers (torch CUDA tensors on this kernel's device).
        params6: [B, 6, M] (shared across chunks) or [B, S, 6, M]; pi: [B, M] or [B, S, M];
        inds: int64 [S].  Returns (ll [B, S] float64, dlog [B, S, 7, M] or None)."""
        import torch

        M = self._M
        S = int(inds.shape[0])
        B = int(params6.shape[0])
        tdtype = torch.float64 if self.double_precision else torch.float32
        assert params6.is_cuda and pi.is_cuda and inds.is_cuda and params6.device.index == self.device
        assert params6.dtype == tdtype and pi.dtype == tdtype and inds.dtype == torch.int64
        assert params6.is_contiguous() and pi.is_contiguous() and inds.is_contiguous()
        if params6.dim() == 3:
            assert params6.shape == (B, 6, M)
            ps_b, ps_s = 6 * M, 0
        else:
            assert params6.shape == (B, S, 6, M)
            ps_b, ps_s = S * 6 * M, 6 * M
        if pi.dim() == 2:
            assert pi.shape == (B, M)
            pis_b, pis_s = M, 0
        else:
            assert pi.shape == (B, S, M)
            pis_b, pis_s = S * M, M
        if ll is None:
            ll = torch.empty((B, S), dtype=torch.float64, device=params6.device)
        if grad and dlog is None:
            dlog = torch.empty((B, S, 7, M), dtype=tdtype, device=params6.device)
        if stream is None:
            stream = torch.cuda.current_stream(params6.device).cuda_stream
        _check(
            self._lib.phb_loglik_device(
                self._handle, params6.data_ptr(), ps_b, ps_s, pi.data_ptr(), pis_b, pis_s, inds.data_ptr(),
                B, S, int(grad), ll.data_ptr(), dlog.data_ptr() if grad else None, ctypes.c_void_p(stream),
            )
        )
        return ll, (dlog if grad else None)

    # ---- time-axis sharding of a small minibatch over several processes (include/phlash_b200.h)
    def sharded_plan(self, B: int, S: int, overlap: int, world: int):
        """(segments, bytes per process of the all-gather buffer); segments == 0: does not apply"""
        n_seg, slot = ctypes.c_int64(), ctypes.c_int64()
        _check(self._lib.phb_hmm_term_sharded_plan(self._handle, int(B), int(S), int(overlap), int(world),
                                                   ctypes.byref(n_seg), ctypes.byref(slot)))
        return n_seg.value, slot.value

    def sharded_begin(self, x, pattern: str, theta: float, inds, overlap: int, rank: int, world: int, gather, stream=None):
        import torch

        widths, B, _ = self._term_args(x, pattern)
        assert inds.is_cuda and inds.dtype == torch.int64 and inds.is_contiguous()
        assert gather.is_cuda and gather.dtype == torch.uint8 and gather.is_contiguous()
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        _check(
            self._lib.phb_hmm_term_sharded_begin(
                self._handle, x.data_ptr(), B, _ptr(widths), len(widths), float(theta), inds.data_ptr(), int(inds.shape[0]),
                int(overlap), int(rank), int(world), gather.data_ptr(), ctypes.c_void_p(stream),
            )
        )

    def sharded_end(self, inds, B: int, overlap: int, rank: int, world: int, gather, stream=None):
        """this process's partial per-particle sums [B, 1 + 7 M] (to be all-reduced)"""
        import torch

        sums = torch.empty((int(B), 1 + 7 * self._M), dtype=torch.float64, device=inds.device)
        if stream is None:
            stream = torch.cuda.current_stream(inds.device).cuda_stream
        _check(
            self._lib.phb_hmm_term_sharded_end(
                self._handle, inds.data_ptr(), int(B), int(inds.shape[0]), int(overlap), int(rank), int(world),
                gather.data_ptr(), sums.data_ptr(), ctypes.c_void_p(stream),
            )
        )
        return sums

    def sum_over_chunks(self, ll, dlog, out=None, stream=None):
        """ll [B, S] float64, dlog [B, S, 7, M] (torch CUDA) -> per-particle sums [B, 1 + 7 M] float64: what one
        process per GPU contributes to the all-reduce of a step (phb_sum_over_chunks_device)."""
        import torch

        B, S = int(ll.shape[0]), int(ll.shape[1])
        assert ll.is_cuda and ll.dtype == torch.float64 and ll.is_contiguous()
        assert dlog is None or (dlog.is_contiguous() and dlog.shape == (B, S, 7, self._M))
        if out is None:
            out = torch.empty((B, 1 + 7 * self._M), dtype=torch.float64, device=ll.device)
        assert out.shape == (B, 1 + 7 * self._M) and out.dtype == torch.float64 and out.is_contiguous()
        if stream is None:
            stream = torch.cuda.current_stream(ll.device).cuda_stream
        _check(
            self._lib.phb_sum_over_chunks_device(
                self._handle, ll.data_ptr(), dlog.data_ptr() if dlog is not None else None, B, S, out.data_ptr(),
                ctypes.c_void_p(stream),
            )
        )
        return out

    def sync(self) -> None:
        _check(self._lib.phb_sync(self._handle))

    # ---- allocation-free, capturable steps; minibatch sampling on the device (mcmc.py:277-278)
    def reserve(self, B: int, S_max: int, overlap: int = 0, grad: bool = True) -> None:
        """Size every scratch buffer a call with B particles and up to S_max chunks can need, so that later
        calls neither allocate nor free (phb_reserve): a whole step can then be recorded into a CUDA graph."""
        _check(self._lib.phb_reserve(self._handle, int(B), int(S_max), int(overlap), int(grad)))

    @property
    def allocation_count(self) -> int:
        return int(self._lib.phb_allocation_count(self._handle))

    def set_iteration(self, iteration: int, stream=None) -> None:
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self._lib.phb_set_iteration(self._handle, int(iteration), ctypes.c_void_p(stream)))

    def sample_minibatch(self, seed: int, S: int, out=None, stream=None):
        """inds ~ choice(N, (S,)) with replacement, drawn ON THE DEVICE for the object's current iteration
        counter (which advances): torch int64 CUDA tensor [S]."""
        import torch

        if out is None:
            out = torch.empty((int(S),), dtype=torch.int64, device=torch.device("cuda", self.device))
        assert out.is_cuda and out.dtype == torch.int64 and out.is_contiguous() and out.shape == (int(S),)
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self._lib.phb_sample_minibatch_device(self._handle, int(seed), int(S), out.data_ptr(), ctypes.c_void_p(stream)))
        return out


def minibatch_indices(seed: int, iteration: int, n_chunks: int, S: int) -> np.ndarray:
    """The generator of ``sample_minibatch`` on the host (pure function of (seed, iteration))."""
    out = np.empty(int(S), dtype=np.int64)
    _native.lib().phb_minibatch_indices(int(seed), int(iteration), int(n_chunks), int(S), _ptr(out))
    return out


def measure_fp32_peak(device: int = 0):
    """(independent-FFMA TFLOP/s, accumulate-pattern TFLOP/s) measured now on ``device``."""
    a, b = ctypes.c_double(), ctypes.c_double()
    _check(_native.lib().phb_measure_fp32_peak(int(device), ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def _normalise_call(pp: PSMCParams, index, M: int):
    """Bring (pp, index) to a parameter block [B, S, 7, M] and indices [S].

    Accepted ranks, as in the reference (gpu.py:186-213): leaves [M] with a scalar index; leaves [M],
    [S, M] or [B, S, M] with an index vector [S] ([M] is broadcast over the S chunks).  Returns the
    block, the index vector and which of the two leading axes were added (so that the caller can
    strip them from the results, gpu.py:319-325)."""
    block = np.stack([np.asarray(leaf) for leaf in pp], axis=-2)
    assert block.shape[-2:] == (7, M), f"expected parameter leaves ending in M={M}"
    index = np.asarray(index)
    assert index.ndim in (0, 1)
    inds = index.reshape(-1)
    n_chunks = inds.shape[0]
    lead = block.ndim - 2  # number of batch axes the caller supplied
    assert lead in (0, 1, 2)
    if index.ndim == 0:
        assert lead == 0, "a scalar index takes unbatched parameters"
        return block[None, None], inds, True, True
    if lead == 0:
        block = np.broadcast_to(block, (1, n_chunks, 7, M))
    elif lead == 1:
        assert block.shape[0] == n_chunks
        block = block[None]
    else:
        assert block.shape[1] == n_chunks
    return block, inds, lead < 2, False


class PSMCKernel:
    """Evaluate the PSMC HMM log-likelihood (and gradient) of rows of ``data`` on B200 GPUs.

    Args (same as the reference, gpu.py:328-351):
        - M: discretization level (4, 8, 16, 32 or 64; the reference supports 16).
        - data: int8 data matrix [N, L], -1 = missing.
        - double_precision: if True, use float64 on the GPU.
        - num_gpus: number of devices to spread the chunk axis over (default: all visible).
    """

    def __init__(self, M, data, double_precision=False, num_gpus: int = None):
        if num_gpus is not None:
            assert num_gpus > 0
        self.double_precision = bool(double_precision)
        self.M = int(M)
        n = _native.lib().phb_device_count()
        if n <= 0:
            raise CudaError(_native.last_error() or "no CUDA device is visible")
        if num_gpus is not None:
            n = min(num_gpus, n)
        self.devices = list(range(n))
        self.gpu_kernels = [_PSMCKernelBase(M, data, double_precision, device=d) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=n) if n > 1 else None

    @property
    def float_type(self):
        return np.float64 if self.double_precision else np.float32

    def set_threads_per_pair(self, t: int) -> None:
        for k in self.gpu_kernels:
            k.set_threads_per_pair(t)

    def loglik(self, pp: PSMCParams, index):
        """Log-likelihood of data[index] (reference: gpu.py:359-362; there it is the
        JAX-differentiable scalar, here the value - use ``__call__(..., grad=True)`` for the
        gradient that the reference's custom_vjp consumes)."""
        return self(pp, index, grad=False)

    def __call__(self, pp: PSMCParams, index, grad: bool):
        pa, inds, added_B, added_S = _normalise_call(pp, index, self.M)
        D = len(self.gpu_kernels)
        if D == 1 or inds.shape[0] < D:
            res = self.gpu_kernels[0].evaluate(pa, inds, grad)
        else:
            parts = np.array_split(np.arange(inds.shape[0]), D)
            futs = [
                self._pool.submit(kern.evaluate, pa[:, sel], inds[sel], grad)
                for kern, sel in zip(self.gpu_kernels, parts)
            ]
            outs = [f.result() for f in futs]
            if grad:
                res = (np.concatenate([o[0] for o in outs], 1), np.concatenate([o[1] for o in outs], 1))
            else:
                res = np.concatenate(outs, 1)
        if grad:
            ll, dlog = res
            ret = (ll, PSMCParams.from_block(dlog))
        else:
            ret = res

        def strip(a):
            if added_B:
                a = a[0]
                return a[0] if added_S else a
            return a

        if grad:
            return strip(ret[0]), PSMCParams(*(strip(a) for a in ret[1]))
        return strip(ret)
