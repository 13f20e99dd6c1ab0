"""Builds libruf_b200.so in-tree with nvcc for sm_100a (and nothing else).

    python -m realtime_urdf_filter_b200.build [--force]

The kernels rely on `-fmad=false` (explicit fmaf only) for bit-exact parity with the raster
specification; do not add -use_fast_math.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
HOST = os.path.join(PKG_DIR, "host")
# RUF_LIB_PATH / RUF_EXTRA_NVCC: build or load an experimental variant next to the default library
LIB_PATH = os.environ.get("RUF_LIB_PATH") or os.path.join(PKG_DIR, "libruf_b200.so")

CUDA_SOURCES = ["ruf_kernels.cu", "ruf_api.cu"]
HOST_SOURCES = ["ruf_host.cpp", "ruf_meshlet.cpp"]
FACADE_SOURCES = ["urdf_model.cpp", "urdf_filter.cpp", "facade_c.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (need CUDA 12.9 with sm_100a support)")


def sources() -> list[str]:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES + HOST_SOURCES]
    srcs += [os.path.join(HOST, s) for s in FACADE_SOURCES if os.path.exists(os.path.join(HOST, s))]
    return srcs


def _deps() -> list[str]:
    deps = sources()
    for d in (CSRC, HOST, os.path.join(PKG_DIR, "..", "include")):
        if os.path.isdir(d):
            deps += [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".h", ".cuh", ".hpp"))]
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    # Several ranks of one torchrun job may arrive here together (a snapshot whose sources look newer than the
    # library): one of them builds under an exclusive lock, into a temporary file that replaces the library
    # atomically; the others wait and then find it up to date.
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB_PATH
        tmp = f"{LIB_PATH}.tmp{os.getpid()}"
        extra = os.environ.get("RUF_EXTRA_NVCC", "").split()
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", os.path.join(PKG_DIR, "..", "include"), "-o", tmp, *sources()]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB_PATH)
        if verbose:
            print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
