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
SOURCES = ["rt_api.cu", "rt_kernels.cu", "rt_wavefront.cu", "rt_tiny.cu", "rt_lbvh.cu", "rt_refit.cu", "rt_display.cu", "rt_bvh.cpp"]
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
    ]


def _deps() -> list:
    return [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    """Every translation unit to an object file (in parallel; an object is reused while it is newer than its source,
    every header and this script), then one link.  The objects live in build/ (git-ignored)."""
    out = os.environ.get("B200RT_LIB_OUT")                       # A/B builds (tools/): another library file, its own objects
    if out:
        return _build_to(os.path.abspath(out), os.path.join(HERE, "build", "alt_" + os.path.basename(out)), True, verbose)
    if not force and not needs_build():
        return LIB
    return _build_to(LIB, os.path.join(HERE, "build"), force, verbose)


def _build_to(LIB: str, objdir: str, force: bool, verbose: bool) -> str:
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    extra = os.environ.get("B200RT_NVCC_EXTRA", "")
    stamp = os.path.join(objdir, "flags.txt")
    if not os.path.exists(stamp) or open(stamp).read() != extra:
        force = True
        with open(stamp, "w") as f:
            f.write(extra)

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objlog = obj + ".log"                                    # ptxas -v output of this unit, kept next to a reused object
        if not force and os.path.exists(obj) and os.path.exists(objlog) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t):
            return obj, open(objlog).read(), 0
        cmd = [nvcc_path()] + flags() + ["-c", path, "-o", obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        text = " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        with open(objlog, "w") as f:
            f.write(text)
        return obj, text, proc.returncode

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = "".join(r[1] for r in results)
    bad = [r for r in results if r[2] != 0]
    if not bad:
        cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + [r[0] for r in results] + ["-o", LIB, "-lgomp"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        if proc.returncode != 0:
            bad = [(LIB, log, proc.returncode)]
    with open(os.path.join(HERE, "build.log") if LIB == globals()["LIB"] else LIB + ".log", "w") as f:     # A/B builds keep their own log
        f.write(log)
    if bad:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
