"""Property tests of the raster specification: fill rule, watertightness, invariances."""
import numpy as np
from hypothesis import given, settings, strategies as st

import helpers
import oracle_py as orc
from realtime_urdf_filter_b200 import synth

W, H = 96, 64
P = synth.kinect_P(W, H)
PROJ, TX, TY = orc.projection_matrix(P, W, H)
VIEW = orc.view_matrix((0, 0, 0, 1), (0, 0, 0), (0, 0, 0, 1), (0, 0, 0), TX, TY)
MVP = orc.compose_mvp(PROJ, VIEW, np.eye(4).reshape(-1), 1)


def cover(tris):
    tris = np.asarray(tris, np.float32).reshape(-1, 9)
    z = orc.render(tris, np.zeros(len(tris), np.uint32), MVP, W, H, np.float32(0))   # bg disabled
    return z < 1.0, z


def win_to_eye(xw, yw, z):
    """inverse of x_w = fx x/z + cx ; y_w = fy y/z + (H - cy)"""
    return [(xw - P[2]) / P[0] * z, (yw - (H - P[6])) / P[5] * z, z]


pt = st.tuples(st.floats(2, W - 2), st.floats(2, H - 2))


@settings(max_examples=60, deadline=None)
@given(pt, pt, pt, pt, st.floats(0.5, 6.0))
def test_shared_edge_is_watertight_and_exclusive(a, b, c, d, z):
    """Two triangles sharing edge (a,b): no pixel is covered twice and the union equals the
    coverage of ... themselves rendered together (top-left rule, S6)."""
    va, vb, vc, vd = (win_to_eye(*p, z) for p in (a, b, c, d))
    # put c and d on opposite sides of ab, otherwise the triangles overlap legitimately
    side = lambda p: (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
    if side(c) * side(d) >= 0:
        return
    c1, _ = cover([va + vb + vc])
    c2, _ = cover([vb + va + vd])
    both, _ = cover([va + vb + vc, vb + va + vd])
    assert not np.any(c1 & c2)
    assert np.array_equal(c1 | c2, both)


@settings(max_examples=40, deadline=None)
@given(pt, pt, pt, st.floats(0.5, 6.0))
def test_winding_does_not_matter(a, b, c, z):
    va, vb, vc = (win_to_eye(*p, z) for p in (a, b, c))
    c1, z1 = cover([va + vb + vc])
    c2, z2 = cover([va + vc + vb])            # no culling in the reference (no glCullFace call)
    assert np.array_equal(c1, c2)
    assert np.allclose(z1, z2, atol=2e-7)


def test_pixel_centre_sampling_exact_square():
    # axis-aligned square from (10,10) to (20,20) in window coords: covers pixel centres 10.5..19.5
    z = 2.0
    q = [win_to_eye(10, 10, z), win_to_eye(20, 10, z), win_to_eye(20, 20, z), win_to_eye(10, 20, z)]
    c, _ = cover([q[0] + q[1] + q[2], q[0] + q[2] + q[3]])
    ys, xs = np.nonzero(c)
    assert xs.min() == 10 and xs.max() == 19 and ys.min() == 10 and ys.max() == 19 and c.sum() == 100


def test_far_clip_per_pixel_and_behind_camera():
    s = 5.0
    beyond = [[-s, -s, 8.5, s, -s, 8.5, s, s, 8.5]]
    assert not cover(beyond)[0].any()
    behind = [[-s, -s, -1.0, s, -s, -1.0, s, s, -1.0]]
    assert not cover(behind)[0].any()
    sloped = [[-s, -s, 7.0, s, -s, 7.0, 0.0, s, 9.5]]       # crosses the far plane: partially visible
    c, z = cover(sloped)
    assert c.any() and synth.linear_depth(z[c]).max() <= 8.0 + 1e-3


def test_degenerate_and_nan_triangles_ignored():
    tris = [[0, 0, 2, 0, 0, 2, 0, 0, 2], [0, 0, 2, 1, 1, 2, 2, 2, 2], [np.nan, 0, 2, 1, 0, 2, 0, 1, 2],
            [np.inf, 0, 2, 1, 0, 2, 0, 1, 2]]
    assert not cover(tris)[0].any()
