"""Builds libnngp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python <package>/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
# the library is built under <repo>/lib (in-tree, short path: the package directory's name is 150 characters long, which made
# the driver's loaded-library hook drop the line); the package directory keeps a symlink to it
SO = os.path.join(os.path.dirname(PKG), "lib", "libnngp_b200.so")
SO_LINK = os.path.join(PKG, "libnngp_b200.so")
SOURCES = ["nngp_b200.cu", "host_graph.cpp", "host_shard.cpp", "r_stream.cpp"]
HEADERS = ["kernels.cuh", "device_math.cuh", "nngp_internal.h", os.path.join("..", "..", "include", "nngp_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
         "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-O2", "-Xlinker", "-lgomp"]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, f) for f in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libnngp_b200.so")
    if not os.path.islink(SO_LINK) or os.path.realpath(SO_LINK) != os.path.realpath(SO):
        if os.path.lexists(SO_LINK):
            os.remove(SO_LINK)
        os.symlink(os.path.relpath(SO, PKG), SO_LINK)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
