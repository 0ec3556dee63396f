"""Builds libb200rt.so (the C-ABI CUDA library) in-tree for sm_100a.

    python -m pgr_raytracing_project_b200.build [--force]

nvcc cross-compiles without a GPU.  Flags that are part of the arithmetic contract
(DESIGN.md): -fmad=false (no implicit FMA contraction; kernels spell their FMAs out), IEEE
division / sqrt (nvcc defaults), never --use_fast_math.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200rt.so")
SOURCES = ["rt_api.cu", "rt_kernels.cu", "rt_wavefront.cu", "rt_lbvh.cu", "rt_refit.cu", "rt_display.cu", "rt_bvh.cpp"]
HEADERS = ["rt_device.cuh", "rt_kernels.h", "rt_kernel_common.cuh", "rt_bvh.h", "rt_lbvh.h", "rt_refit.h", "rt_display.h", os.path.join("..", "..", "include", "b200rt.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def flags() -> list:
    return [
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-std=c++17", "-lineinfo", "-fmad=false", *os.environ.get("B200RT_NVCC_EXTRA", "").split(),
        "-Xcompiler", "-fPIC,-fopenmp,-O2,-fno-fast-math,-ffp-contract=off",
        "-Xptxas", "-v",
        "-shared",
    ]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + flags() + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-lgomp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
