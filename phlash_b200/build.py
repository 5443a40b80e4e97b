"""Builds csrc/ into phlash_b200/_lib/libphlash_b200.so with nvcc for sm_100a (in-tree, so the
.so travels with a gpurun snapshot; *.so is git-ignored)."""

from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_lib", "libphlash_b200.so")
SOURCES = [os.path.join(_PKG, "csrc", "phlash_b200.cu")]
import glob  # noqa: E402

# every source / header the library is built from: editing any of them rebuilds
DEPS = sorted(set(SOURCES + glob.glob(os.path.join(_PKG, "csrc", "*.cu*")) + glob.glob(os.path.join(_PKG, "csrc", "*.h"))
                  + glob.glob(os.path.join(os.path.dirname(_PKG), "include", "*.h"))))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, *SOURCES, "-o", LIB_PATH]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
