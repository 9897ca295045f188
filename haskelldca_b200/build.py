"""Build libdca_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdca_b200.so")
SOURCES = ["dca_b200.cu"]
def deps():
    """every source the library is built from: all of csrc/ plus the C-ABI header"""
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
        os.path.join(os.path.dirname(HERE), "include", "dca_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",  # parity: the only fused multiply-adds are the explicit __fmaf_rn / __fma_rn
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdca_b200.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str | None = None) -> str:
    """`extra_flags` / `out` build a development variant (e.g. -DDCA_MIN_CTAS=4) next to the product library."""
    out = out or LIB
    if not force and out == LIB and not is_stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra_flags, "-o", out, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    env.pop("CC", None)  # the image's CC wrapper is not a valid nvcc host compiler
    env.pop("CXX", None)
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return out


HOST_DIR = os.path.join(os.path.dirname(HERE), "host")
EXE = os.path.join(OUT_DIR, "dca-exe")


def build_host(force: bool = False) -> str:
    """g++ host/main.cpp (the C++ host layer + CLI over the C ABI) -> _build/dca-exe, linked against libdca_b200.so."""
    srcs = [os.path.join(HOST_DIR, "main.cpp"), os.path.join(HOST_DIR, "dca_host.hpp"), LIB]
    if not force and os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(x) for x in srcs):
        return EXE
    gxx = shutil.which("g++") or "/usr/bin/g++"
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-Wall", "-I", os.path.join(os.path.dirname(HERE), "include"),
                           "-o", EXE, os.path.join(HOST_DIR, "main.cpp"), "-L", OUT_DIR, "-ldca_b200",
                           "-Wl,-rpath,$ORIGIN"], env=env)
    return EXE


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_host(force=True))
