"""In-tree builds: libpinot_b200.so (CUDA, sm_100a) and the CPU-only segment-writer helper.

The library is compiled per translation unit (objects under pinot_b200/build/, rebuilt only when stale, in parallel) and
linked with nvcc: pb_engine.cu (runtime + generic kernels), pb_filter_spec.cu (the plan-time specialised instantiations of
the filter kernel: one small kernel per bit width and predicate kind) and host/pb_host.cpp (the planning layer)."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# PB_SCRATCH_BUILD=<dir>: compile and link into <dir> instead of the tree (a compile check that leaves the in-tree library,
# which a GPU run in flight may be about to snapshot, untouched)
_SCRATCH = os.environ.get("PB_SCRATCH_BUILD")
LIB = os.path.join(_SCRATCH or HERE, "libpinot_b200.so")
OBJ_DIR = os.path.join(_SCRATCH, "obj") if _SCRATCH else os.path.join(HERE, "build")
SOURCES = [os.path.join(HERE, "csrc", "pb_engine.cu"), os.path.join(HERE, "csrc", "pb_filter_spec.cu"),
           os.path.join(HERE, "csrc", "host", "pb_host.cpp")]
HEADERS = [os.path.join(HERE, "csrc", "pb_device.cuh"), os.path.join(HERE, "csrc", "pb_internal.h"),
           os.path.join(ROOT, "include", "pinot_b200.h"), os.path.join(ROOT, "include", "pinot_b200_host.h")]
DEPS = SOURCES + HEADERS

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(src) + ".o")


def build_native(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a without a GPU."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in HEADERS)

    def stale(src):
        o = _obj(src)
        return force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(src), newest_header)

    todo = [s for s in SOURCES if stale(s)]

    def compile_one(src):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(src), src]
        subprocess.check_call(cmd)

    if todo:
        with ThreadPoolExecutor(max_workers=len(todo)) as ex:
            list(ex.map(compile_one, todo))
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(_obj(s)) for s in SOURCES):
        subprocess.check_call([nvcc_path()] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + [_obj(s) for s in SOURCES] + ["-ldl"])
    return LIB


def build_all(force: bool = False) -> None:
    from .segment_writer import build_segwriter
    build_segwriter(force)
    build_native(force)
