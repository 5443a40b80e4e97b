"""Compile the reference's own CUDA kernels (the KERNEL_SRC string of
/root/reference/src/phlash/gpu.py:475-693) into cubins under oracle/_ref/ (git-ignored; they
travel to the GPU box with the gpurun snapshot).  Test / benchmark infrastructure only.

The source is read where it lies with ``ast`` (importing the module would need jax), prefixed with
the same two lines the reference prepends (gpu.py:131-138) and compiled by NVRTC for sm_100 the way
gpu.py:49-70 does.  No reference source is copied into this repository.

M=16 builds in float and double, M=32 in float only; M=32 double and M=64 exceed the 48 KiB static
shared-memory limit of the reference kernel and cannot be built (recorded in manifest.json)."""

from __future__ import annotations

import ast
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_GPU_PY = "/root/reference/src/phlash/gpu.py"


def kernel_source() -> str:
    tree = ast.parse(open(REF_GPU_PY).read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(getattr(t, "id", None) == "KERNEL_SRC" for t in node.targets):
            return ast.literal_eval(node.value)
    raise RuntimeError("KERNEL_SRC not found in the reference")


def compile_cubin(src: str, arch: str = "sm_100") -> bytes:
    from cuda.bindings import nvrtc

    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"kern.cu", 0, [], [])
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
    (err,) = nvrtc.nvrtcCompileProgram(prog, 1, [f"--gpu-architecture={arch}".encode()])
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        _, n = nvrtc.nvrtcGetProgramLogSize(prog)
        log = bytearray(n)
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError(log.decode(errors="replace"))
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    data = bytearray(n)
    nvrtc.nvrtcGetCUBIN(prog, data)
    return bytes(data)


def build(verbose: bool = True) -> dict:
    if not os.path.exists(REF_GPU_PY):
        raise FileNotFoundError(REF_GPU_PY)
    os.makedirs(OUT, exist_ok=True)
    body = kernel_source()
    manifest = {}
    for m in (16, 32, 64):
        for dbl in (False, True):
            tag = f"M{m}_{'f64' if dbl else 'f32'}"
            src = "\n".join([f"#define M {m}", "typedef double FLOAT;" if dbl else "typedef float FLOAT;", body])
            try:
                cubin = compile_cubin(src)
                with open(os.path.join(OUT, f"ref_{tag}.cubin"), "wb") as fh:
                    fh.write(cubin)
                manifest[tag] = {"ok": True, "bytes": len(cubin)}
            except RuntimeError as e:
                first = [ln for ln in str(e).splitlines() if "error" in ln.lower()][:1]
                manifest[tag] = {"ok": False, "why": first[0] if first else str(e)[:200]}
            if verbose:
                print(tag, manifest[tag], file=sys.stderr)
    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    return manifest


if __name__ == "__main__":
    build()
