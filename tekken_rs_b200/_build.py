"""Build the CUDA library in-tree: tekken_rs_b200/libtekken_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# TEKKEN_B200_LIB: load (and build into) another file, e.g. an experimental build with TEKKEN_B200_NVCC_FLAGS="-DPT_MINB=5"
LIB = os.environ.get("TEKKEN_B200_LIB") or os.path.join(HERE, "libtekken_b200.so")
SOURCES = ["tk_kernels.cu", "tk_decode.cu", "tk_api.cu", "tk_host.cpp"]
HEADERS = ["tk_common.h", "tk_host.h", "tk_pretok.h", "tk_pretok_cfg.h", "tk_small.cuh", "tk_device.cuh", "tk_kernels.h", "unicode_ranges.inc", "unicode_subclasses.inc",
           os.path.join("..", "..", "include", "tekken_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the library cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd[1:1] = os.environ.get("TEKKEN_B200_NVCC_FLAGS", "").split()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))


DEBUG_LIB = os.path.join(HERE, "libtekken_b200_dbg.so")


def build_debug(force: bool = False) -> str:
    """The bounds-checked variant (-DTK_DEBUG_BOUNDS, see tk_kernels.cu): libtekken_b200_dbg.so next to the library.
    tests/test_gpu_debug_bounds.py runs the parity corpus through it (compute-sanitizer is closed on the GPU pool)."""
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    if not force and os.path.exists(DEBUG_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(DEBUG_LIB) for d in deps if os.path.exists(d)):
        return DEBUG_LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DTK_DEBUG_BOUNDS",
           "-Xcompiler", "-fPIC", "-shared", "-o", DEBUG_LIB + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    os.replace(DEBUG_LIB + ".tmp", DEBUG_LIB)
    return DEBUG_LIB
