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


HOST = os.path.join(HERE, "host")
HOST_LIB = os.path.join(LIBDIR, "libhimg.so")
DRIVERS = ["chimg", "dhimg", "benchmark"]


def build_host(force: bool = False) -> dict:
    """C++ host layer: libhimg.so (himg::Encoder / himg::Decoder on top of libhimgcu.so) and the
    chimg / dhimg / benchmark drivers (target names as in the reference's CMake files)."""
    build(force)
    srcs = [os.path.join(HOST, "encoder.cpp"), os.path.join(HOST, "decoder.cpp")]
    deps = srcs + sorted(glob.glob(os.path.join(HOST, "*.h"))) + [LIB]
    out = {"libhimg": HOST_LIB}
    common = ["g++", "-std=c++11", "-O2", "-Wall", "-Wextra", "-fPIC"]
    rpath = "-Wl,-rpath,$ORIGIN"
    if force or not os.path.exists(HOST_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOST_LIB) for d in deps):
        proc = subprocess.run(common + ["-shared", "-o", HOST_LIB] + srcs + ["-L" + LIBDIR, "-lhimgcu", rpath],
                              capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("host library build failed")
    for d in DRIVERS:
        exe = os.path.join(LIBDIR, d)
        src = os.path.join(HOST, d + ".cpp")
        out[d] = exe
        if force or not os.path.exists(exe) or any(os.path.getmtime(x) > os.path.getmtime(exe) for x in [src, HOST_LIB] + deps):
            proc = subprocess.run(common + ["-o", exe, src, "-L" + LIBDIR, "-lhimg", "-lhimgcu", rpath],
                                  capture_output=True, text=True)
            if proc.returncode != 0:
                sys.stderr.write(proc.stdout + proc.stderr)
                raise RuntimeError(f"driver build failed: {d}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else []))
    print(build_host(force="--force" in sys.argv))
