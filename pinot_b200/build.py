"""In-tree builds: libpinot_b200.so (CUDA, sm_100a) and the CPU-only segment-writer helper."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libpinot_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "pb_engine.cu"), os.path.join(HERE, "csrc", "host", "pb_host.cpp")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "pb_device.cuh"), os.path.join(HERE, "csrc", "pb_internal.h"),
                  os.path.join(ROOT, "include", "pinot_b200.h"), os.path.join(ROOT, "include", "pinot_b200_host.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def build_native(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a without a GPU."""
    stale = force or not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(d) for d in DEPS)
    if stale:
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        subprocess.check_call(cmd)
    return LIB


def build_all(force: bool = False) -> None:
    from .segment_writer import build_segwriter
    build_segwriter(force)
    build_native(force)
