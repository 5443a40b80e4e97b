"""Run the reference's own CUDA kernels (cubins built by oracle/build_ref.py) on the current GPU
with the reference's launch geometry (gpu.py:261-279) and host post-processing (gpu.py:303-313).
Test / benchmark infrastructure: the fp64 build at M=16 is the reference's own implementation and
serves as a second oracle; the fp32 builds are the "reference GPU path" timed beside ours."""

from __future__ import annotations

import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available(m: int, double_precision: bool) -> bool:
    tag = f"M{m}_{'f64' if double_precision else 'f32'}"
    return os.path.exists(os.path.join(REF_DIR, f"ref_{tag}.cubin"))


def manifest() -> dict:
    path = os.path.join(REF_DIR, "manifest.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def _ok(res, what):
    from cuda.bindings import driver as cu

    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"{what}: {err}")
    return res[1] if len(res) == 2 else res[1:]


class ReferenceKernel:
    """data int8 [N, L]; call(pa [B, S, 7, M], inds [S], grad) like _PSMCKernelBase.__call__."""

    def __init__(self, m: int, data: np.ndarray, double_precision: bool = False):
        from cuda.bindings import driver as cu

        self.cu = cu
        self.m = m
        self.dbl = double_precision
        self.ft = np.float64 if double_precision else np.float32
        tag = f"M{m}_{'f64' if double_precision else 'f32'}"
        cubin = open(os.path.join(REF_DIR, f"ref_{tag}.cubin"), "rb").read()
        _ok(cu.cuInit(0), "cuInit")
        dev = _ok(cu.cuDeviceGet(0), "cuDeviceGet")
        self.ctx = _ok(cu.cuDevicePrimaryCtxRetain(dev), "ctx")
        _ok(cu.cuCtxSetCurrent(self.ctx), "setctx")
        self.mod = _ok(cu.cuModuleLoadData(cubin), "cuModuleLoadData")
        self.f_grad = _ok(cu.cuModuleGetFunction(self.mod, b"loglik_grad"), "getfn")
        self.f_ll = _ok(cu.cuModuleGetFunction(self.mod, b"loglik"), "getfn")
        data = np.ascontiguousarray(data.clip(-1, 1), dtype=np.int8)
        self.n, self.length = data.shape
        self.d_data = _ok(cu.cuMemAlloc(data.nbytes), "alloc data")
        _ok(cu.cuMemcpyHtoD(self.d_data, data.ctypes.data, data.nbytes), "h2d data")
        self.stream = _ok(cu.cuStreamCreate(0), "stream")
        self.ev0 = _ok(cu.cuEventCreate(0), "event")
        self.ev1 = _ok(cu.cuEventCreate(0), "event")
        self.last_ms = None

    def __call__(self, pa: np.ndarray, inds: np.ndarray, grad: bool = True):
        cu = self.cu
        _ok(cu.cuCtxSetCurrent(self.ctx), "setctx")
        b, s = pa.shape[:2]
        assert pa.shape == (b, s, 7, self.m)
        pa = np.ascontiguousarray(pa, dtype=self.ft)
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        ll = np.zeros((b, s), dtype=np.float64)
        dlog = np.zeros((b, s, 7, self.m), dtype=self.ft)
        d_inds = _ok(cu.cuMemAlloc(inds.nbytes), "alloc")
        d_pa = _ok(cu.cuMemAlloc(pa.nbytes), "alloc")
        d_ll = _ok(cu.cuMemAlloc(ll.nbytes), "alloc")
        d_dlog = _ok(cu.cuMemAlloc(dlog.nbytes), "alloc")
        try:
            # NB: always pass explicit addresses - cuda.bindings coerces a 1-element integer array
            # to an *address* (int(array)) instead of using its buffer, which segfaults for S == 1
            for d, h in ((d_inds, inds), (d_pa, pa), (d_ll, ll), (d_dlog, dlog)):
                _ok(cu.cuMemcpyHtoD(d, h.ctypes.data, h.nbytes), "h2d")
            vals = [self.d_data, np.int64(self.length), np.int64(self.n), d_inds, d_pa, d_ll]
            types = [None, ctypes.c_int64, ctypes.c_int64, None, None, None]
            if grad:
                vals.append(d_dlog)
                types.append(None)
                fn, grid, block = self.f_grad, (b, s, 1), (7, self.m, 1)
            else:
                fn, grid, block = self.f_ll, (b, 1, 1), (s, 1, 1)
            _ok(cu.cuEventRecord(self.ev0, self.stream), "rec")
            _ok(cu.cuLaunchKernel(fn, *grid, *block, 0, self.stream, (tuple(vals), tuple(types)), 0), "launch")
            _ok(cu.cuEventRecord(self.ev1, self.stream), "rec")
            _ok(cu.cuStreamSynchronize(self.stream), "sync")
            self.last_ms = _ok(cu.cuEventElapsedTime(self.ev0, self.ev1), "elapsed")
            _ok(cu.cuMemcpyDtoH(ll.ctypes.data, d_ll, ll.nbytes), "d2h")
            if grad:
                _ok(cu.cuMemcpyDtoH(dlog.ctypes.data, d_dlog, dlog.nbytes), "d2h")
        finally:
            for d in (d_inds, d_pa, d_ll, d_dlog):
                cu.cuMemFree(d)
        if not grad:
            return ll
        dlog[..., 3, :] = np.roll(dlog[..., 3, :], 1, axis=-1)  # gpu.py:303-313
        return ll, dlog

    def close(self):
        cu = self.cu
        cu.cuMemFree(self.d_data)
        cu.cuModuleUnload(self.mod)
        cu.cuStreamDestroy(self.stream)
