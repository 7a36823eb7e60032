"""Build recipe for the ORACLE (test infrastructure only).

Two artefacts:

* ``oracle/_build/libhimg_oracle.so`` -- our own plain-C restatement (``himg_oracle.c``).
* ``oracle/_ref/libhimg_ref.so``      -- the UNMODIFIED reference ``src/lib`` compiled from the
  sources where they lie under ``/root/reference`` plus the small extern "C" shim
  ``ref_shim.cpp``.  Only built when ``/root/reference`` exists (i.e. in the build container);
  the GPU box uses the prebuilt file that travels with the snapshot.  No reference source is
  ever copied into the repository.

The reference's own CMake build is not used (its top level hard-requires FreeImage,
src/CMakeLists.txt:22-25); ``src/lib`` needs nothing but libstdc++/pthread.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB_DIR = "/root/reference/src/lib"
ORACLE_SO = os.path.join(HERE, "_build", "libhimg_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhimg_ref.so")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("oracle build failed: " + " ".join(cmd))


def build_port(force: bool = False) -> str:
    """Compile the plain-C restatement."""
    srcs = [os.path.join(HERE, "himg_oracle.c"), os.path.join(HERE, "himg_oracle.h")]
    if not force and _newer(ORACLE_SO, srcs):
        return ORACLE_SO
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    _run(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", ORACLE_SO, srcs[0]])
    return ORACLE_SO


def build_ref(force: bool = False) -> str | None:
    """Compile the reference's src/lib (in place) + shim into oracle/_ref/libhimg_ref.so."""
    if not os.path.isdir(REF_LIB_DIR):
        return REF_SO if os.path.exists(REF_SO) else None
    ref_srcs = sorted(glob.glob(os.path.join(REF_LIB_DIR, "*.cpp")))
    srcs = ref_srcs + [os.path.join(HERE, "ref_shim.cpp")]
    if not force and _newer(REF_SO, srcs):
        return REF_SO
    os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
    # -O2: the reference's CMake sets no build type; an explicit level is stated.
    _run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-pthread", "-I", REF_LIB_DIR, "-o", REF_SO] + srcs)
    return REF_SO


def build_all(force: bool = False) -> dict:
    return {"port": build_port(force), "ref": build_ref(force)}


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
