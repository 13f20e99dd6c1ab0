"""GPU: the CUDA path against what the reference's own shaders produce in a real GL driver (tests/golden/gl_llvmpipe.npz,
Mesa 18.1.9 llvmpipe; see tests/test_gl_golden.py) -- the float image the reference's filter() works on, and the 16UC1
image filter_callback publishes (convertTo(CV_16U, 1000) of the float result, src/urdf_filter.cpp:309-312)."""
import numpy as np
import pytest

import oracle_py as orc
import realtime_urdf_filter_b200 as ruf
import helpers
from test_gl_golden import CASES, FUZZ, gl_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cluster", ["0", "1"])
@pytest.mark.parametrize("i", range(len(CASES)))
def test_cuda_path_reproduces_the_gl_driver_bit_for_bit(i, cluster, monkeypatch):
    monkeypatch.setenv("RUF_CLUSTER", cluster)
    monkeypatch.setenv("RUF_FINE_MESHLETS", cluster)
    c, gl_depth, gl_mask = CASES[i]
    sc = gl_case._scene(c["scene"])
    depth = gl_case._frame_depth(sc, c["frame"])                 # float metres, as filter_callback hands them to filter()
    proj, _, _ = sc.proj()
    view, pm = sc.frame(c["frame"])
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        got_d, got_m = ctx.filter(depth, proj, view, pm, sc.max_diff, sc.replace_value)
        assert np.array_equal(got_m, gl_mask), f"{np.count_nonzero(got_m != gl_mask)} mask pixels differ from the GL driver"
        assert np.array_equal(got_d.view(np.uint32), gl_depth.view(np.uint32))
        # the 16UC1 wire format: same frame in millimetres in, convertTo(CV_16U, 1000) of the GL result out
        u16 = orc.f32_to_u16(depth.reshape(-1) ).reshape(depth.shape)
        assert np.array_equal(orc.u16_to_f32(u16.reshape(-1)).reshape(depth.shape), depth)
        got_u, got_m2 = ctx.filter(u16, proj, view, pm, sc.max_diff, sc.replace_value)
        assert np.array_equal(got_m2, gl_mask)
        assert np.array_equal(got_u, orc.f32_to_u16(gl_depth.reshape(-1)).reshape(depth.shape))


@pytest.mark.parametrize("j", range(len(FUZZ)))
def test_cuda_path_against_the_gl_driver_on_hostile_soups(j):
    f, gl_depth, gl_mask = FUZZ[j]
    fc = helpers.fuzz_case(f["seed"], special=f.get("special", False))
    with ruf.Context(fc["W"], fc["H"]) as ctx:
        ctx.set_model(fc["tri"], fc["part"], fc["n_parts"])
        got_d, got_m = ctx.filter(fc["depth"], fc["proj"], fc["view"], fc["pm"], fc["max_diff"], fc["replace_value"])
    dm = got_m != gl_mask
    assert int(dm.sum()) == f["mask_pixels_differing_from_oracle"]          # the very pixels the oracle differs on
    assert np.array_equal(got_d.view(np.uint32)[~dm], gl_depth.view(np.uint32)[~dm])
