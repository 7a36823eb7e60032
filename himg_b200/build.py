"""Build recipe for the product library: hand-written sm_100a CUDA + thin C++ host code.

    himg_b200/_lib/libhimgcu.so   the C ABI of include/himg_cuda.h (kernels + launch sequences)

nvcc cross-compiles for sm_100a without a GPU; the built .so stays in-tree (git-ignored, but it
travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libhimgcu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _deps():
    return _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))) + [
        os.path.join(os.path.dirname(HERE), "include", "himg_cuda.h")
    ]


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    deps = _deps()
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [NVCC] + NVCC_FLAGS + list(extra) + ["-o", LIB] + _sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else []))
