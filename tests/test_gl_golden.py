"""Parity pinned against the reference itself: tests/golden/gl_llvmpipe.npz holds what the reference's own shader files
(unmodified) produce in a real GL driver (Mesa 18.1.9 llvmpipe) for the GL call sequence of RealtimeURDFFilter::render --
filtered depth (attachment 1) and mask (attachment 3), see tests/golden/make_gl_golden.py.  The CPU oracle must reproduce them
bit for bit here; tests/test_gpu_gl_golden.py holds the CUDA path to the same vectors."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "gl_ref"))
import gl_case  # noqa: E402
import oracle_py as orc  # noqa: E402
from realtime_urdf_filter_b200 import synth  # noqa: E402


def load_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "gl_llvmpipe.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, [(c, z[f"depth_{i}"].view(np.float32), z[f"mask_{i}"]) for i, c in enumerate(meta["cases"])]


META, CASES = load_golden()
_Z = np.load(os.path.join(ROOT, "tests", "golden", "gl_llvmpipe.npz"))
FUZZ = [(f, _Z[f"fuzz_depth_{j}"].view(np.float32), _Z[f"fuzz_mask_{j}"]) for j, f in enumerate(META["fuzz"])]


def test_the_vectors_come_from_a_real_driver_and_the_unmodified_shaders():
    assert "llvmpipe" in META["gl"] and "Mesa" in META["gl"] and "unmodified" in META["shaders"]
    assert len(CASES) >= 6 and {c["scene"] for c, _, _ in CASES} >= {"small:example", "small:pr2_small", "small:walls"}
    for c, d, m in CASES:
        assert d.shape == (c["height"], c["width"]) and m.shape == d.shape and set(np.unique(m)) <= {0, 255}
        assert 0 < np.count_nonzero(m) < m.size                 # both outcomes of the comparison in every image


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_reproduces_the_gl_driver_bit_for_bit(i):
    c, gl_depth, gl_mask = CASES[i]
    sc = gl_case._scene(c["scene"])
    depth = gl_case._frame_depth(sc, c["frame"])
    proj, _, _ = sc.proj()
    view, pm = sc.frame(c["frame"])
    mvp = orc.compose_mvp(proj, view, pm, sc.n_parts)
    want_d, want_m, _ = orc.filter_frame(depth, sc.tri, sc.tri_part, mvp, np.float32(synth.Z_NEAR), np.float32(synth.Z_FAR),
                                         np.float32(sc.max_diff), np.float32(sc.replace_value), want_mask=True, nthreads=4)
    assert np.array_equal(want_m, gl_mask), f"{np.count_nonzero(want_m != gl_mask)} mask pixels differ from the GL driver"
    assert np.array_equal(want_d.view(np.uint32), gl_depth.view(np.uint32))


@pytest.mark.parametrize("j", range(len(FUZZ)))
def test_oracle_against_the_gl_driver_on_hostile_soups(j):
    """Slivers, triangles through the near plane, non-finite and far-away vertices, mirrored parts (tests/helpers.py::fuzz_case):
    the filtered depth is identical wherever the mask agrees; the mask differs from the driver's on the recorded handful of
    pixels (GL multiplies its matrix stack and clips in float) -- never more than 5 in 10000."""
    import helpers
    f, gl_depth, gl_mask = FUZZ[j]
    fc = helpers.fuzz_case(f["seed"], special=f.get("special", False))
    want_d, want_m = helpers.fuzz_oracle(fc)
    dm = want_m != gl_mask
    assert int(dm.sum()) == f["mask_pixels_differing_from_oracle"] <= max(4, 5e-4 * dm.size)
    assert np.array_equal(want_d.view(np.uint32)[~dm], gl_depth.view(np.uint32)[~dm])
