"""The C-ABI library loads and exports every symbol include/ruf_b200.h declares; without a GPU the
compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

import conftest
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    out = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            out += re.findall(r"RUF_API\s+[\w\s\*]+?\b(ruf\w+)\s*\(", src)
    return sorted(set(out))


def test_header_symbols_exported():
    syms = declared_symbols()
    assert len(syms) >= 25 and "ruf_filter" in syms and "ruf_filter_batch_device" in syms
    lib = ctypes.CDLL(ruf.lib_path())
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    # and the Python binding covers every declared symbol
    missing = [s for s in syms if s not in _lib.SIGNATURES]
    assert not missing, missing


def test_only_ruf_symbols_are_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", ruf.lib_path()], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert names and all(n.startswith("ruf_") for n in names), [n for n in names if not n.startswith("ruf_")][:5]


def test_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", ruf.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    src_dir = os.path.join(ROOT, "realtime_urdf_filter_b200")
    for dp, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle_py" not in text and "ruf_oracle" not in text, f


@pytest.mark.skipif(conftest.HAS_GPU, reason="only meaningful on a machine without a GPU")
def test_no_gpu_means_loud_failure():
    with pytest.raises(ruf.RufError) as e:
        ruf.Context(640, 480)
    assert e.value.code == ruf.RUF_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_argument_validation_without_gpu():
    lib = ruf.load()
    h = ctypes.c_void_p()
    assert lib.ruf_create(ctypes.byref(h), 0, 0, 480, 0.1, 8.0) == ruf.RUF_ERR_INVALID
    assert lib.ruf_create(ctypes.byref(h), 0, 640, 480, 8.0, 0.1) == ruf.RUF_ERR_INVALID
    assert b"z_near" in lib.ruf_last_error(None)
    assert lib.ruf_destroy(None) == 0
