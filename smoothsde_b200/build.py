"""In-tree build of the CUDA library (nvcc cross-compiles sm_100a without a GPU).

    python -m smoothsde_b200.build

produces smoothsde_b200/lib/libsmoothsde_b200.so, the C-ABI shared library declared in
include/smoothsde_b200.h.  The .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsmoothsde_b200" + os.environ.get("SSDE_LIB_SUFFIX", "") + ".so")   # suffix: diagnostics builds

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]
# SSDE_FAST_BUILD=1: optimise the kernels of a translation unit in parallel (1.5 min instead of 4).
# Development only -- the split changes register allocation of the two hot kernels and measured
# 4 % slower on the B200 (12.9 vs 12.45 ms / 1.024e8 rows), so release builds stay single-unit.
FAST_FLAGS = ["--split-compile", "0", "--threads", "0"]
LINK_FLAGS = ["-lcusolver", "-lpthread"]          # dense Cholesky of the random-effect Hessian block (ssde_laplace.cu)


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "smoothsde_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("SSDE_NVCC_EXTRA", "").split()      # e.g. -DSSDE_STATS for the diagnostics build
    if os.environ.get("SSDE_FAST_BUILD"):
        extra = FAST_FLAGS + extra
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources() + LINK_FLAGS
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
