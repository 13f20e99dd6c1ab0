"""Committed golden vectors of the whole path (tests/golden/make_path_golden.py): frozen oracle outputs.

CPU half: the oracle, rebuilt from today's sources, still reproduces them bit for bit.
GPU half: the CUDA path (through the C ABI) reproduces the committed bytes without the oracle in the loop.
The reference itself has no golden vectors for this path, so these pin regressions, not GL parity.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import helpers
import oracle_py as orc
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = np.load(os.path.join(GOLD, "path_small.npz"))
with open(os.path.join(GOLD, "path_hashes.json")) as _f:
    HASHES = json.load(_f)["cases"]
SMALL_FRAMES = [(k, enc) for k in (0, 5) for enc in ("u16", "f32")]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def small(k, enc, key):
    return SMALL[f"f{k}_{enc}_{key}"]


# ------------------------------------------------------------------ CPU: oracle vs committed vectors
@pytest.mark.parametrize("k,enc", SMALL_FRAMES)
def test_oracle_reproduces_small_golden(k, enc):
    n_parts = int(SMALL["n_parts"])
    mvp = orc.compose_mvp(small(k, enc, "proj"), small(k, enc, "view"), small(k, enc, "pm"), n_parts)
    assert same_bits(mvp, small(k, enc, "mvp"))
    out, mask, zbuf = orc.filter_frame(small(k, enc, "depth"), SMALL["tri"], SMALL["tri_part"], mvp,
                                       SMALL["z_near"][()], SMALL["z_far"][()], SMALL["max_diff"][()],
                                       SMALL["replace_value"][()], want_mask=True, nthreads=2, want_zbuf=True)
    assert same_bits(zbuf, small(k, enc, "zbuf"))
    assert same_bits(mask, small(k, enc, "mask"))
    assert same_bits(out, small(k, enc, "out"))


def test_small_golden_exercises_every_outcome():
    for k, enc in SMALL_FRAMES:
        mask, z, d = small(k, enc, "mask"), small(k, enc, "zbuf"), small(k, enc, "depth")
        assert set(np.unique(mask)) == {0, 255}                       # F2: 0 / 255 only
        robot = z < z.max()
        assert 0 < np.count_nonzero(robot) < z.size                   # model and background pixels
        assert np.count_nonzero(robot & (mask == 0)) > 0              # kept-in-front (occluders / invalid) on the model
        assert np.count_nonzero(robot & (mask == 255)) > 0            # filtered on the model surface
        invalid = (d == 0) if enc == "u16" else np.isnan(d)
        assert np.count_nonzero(invalid) > 0                          # invalid sensor pixels present
        if enc == "f32":
            assert not np.any(mask[invalid] == 255)                   # NaN > x is false: never filtered (frag:23)
            assert np.all(np.isnan(small(k, enc, "out")[invalid]))    # and passes through as NaN
    assert sorted(np.unique(SMALL["tri_part"])) == list(range(int(SMALL["n_parts"])))


@pytest.mark.parametrize("case", HASHES, ids=lambda c: f"{c['scene']}-{c['frame']}")
def test_oracle_and_generators_reproduce_hashes(case):
    sc = helpers.scene(case["scene"])
    assert (sc.width, sc.height, sc.n_tris, sc.n_parts) == (case["width"], case["height"], case["n_tris"], case["n_parts"])
    assert sha(sc.tri) == case["tri"] and sha(sc.tri_part) == case["tri_part"]
    for enc in ("u16", "f32"):
        fr = helpers.make_frame(sc, case["frame"], enc)
        want = case[enc]
        assert sha(helpers.oracle_mvp(sc, fr["view"], fr["pm"])) == want["mvp"]
        assert sha(fr["depth"]) == want["depth_in"]
        out, mask, zbuf = helpers.oracle_filter(sc, fr)
        assert sha(zbuf) == want["zbuf"] and sha(mask) == want["mask"] and sha(out) == want["depth_out"]
        assert int(np.count_nonzero(mask)) == want["masked_px"]


# ------------------------------------------------------------------ GPU: CUDA path vs committed vectors
@pytest.mark.gpu
@pytest.mark.parametrize("k,enc", SMALL_FRAMES)
def test_cuda_reproduces_small_golden(k, enc):
    with ruf.Context(int(SMALL["width"]), int(SMALL["height"])) as ctx:
        ctx.set_model(SMALL["tri"], SMALL["tri_part"], int(SMALL["n_parts"]))
        got_d, got_m = ctx.filter(small(k, enc, "depth"), small(k, enc, "proj"), small(k, enc, "view"),
                                  small(k, enc, "pm"), float(SMALL["max_diff"]), float(SMALL["replace_value"]))
    assert same_bits(got_m, small(k, enc, "mask"))
    assert same_bits(got_d, small(k, enc, "out"))


@pytest.mark.gpu
@pytest.mark.parametrize("case", HASHES, ids=lambda c: f"{c['scene']}-{c['frame']}")
def test_cuda_reproduces_hashes(case):
    sc = helpers.scene(case["scene"])
    proj, _, _ = sc.proj()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        for enc in ("u16", "f32"):
            fr = helpers.make_frame(sc, case["frame"], enc)
            assert sha(fr["depth"]) == case[enc]["depth_in"]
            got_d, got_m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
            assert sha(got_m) == case[enc]["mask"]
            assert sha(got_d) == case[enc]["depth_out"]
