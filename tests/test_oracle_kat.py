"""Analytic known-answer tests of the oracle."""
import numpy as np
import pytest

import helpers
import oracle_py as orc
from realtime_urdf_filter_b200 import synth

ZN, ZF = np.float32(0.1), np.float32(8.0)
IDQ, ZERO = (0, 0, 0, 1), (0, 0, 0)


def cam_setup(W=640, H=480, P=None):
    P = synth.kinect_P(W, H) if P is None else P
    proj, tx, ty = orc.projection_matrix(P, W, H)
    view = orc.view_matrix(IDQ, ZERO, IDQ, ZERO, tx, ty)       # camera frame == fixed frame
    return proj, view


def render_tris(tri, W=640, H=480, P=None, model=None):
    proj, view = cam_setup(W, H, P)
    model = np.eye(4).T.reshape(-1) if model is None else model
    mvp = orc.compose_mvp(proj, view, model, 1)
    tri = np.asarray(tri, np.float32).reshape(-1, 9)
    return orc.render(tri, np.zeros(len(tri), np.uint32), mvp, W, H, helpers.BG_Z)


def test_projection_matrix_literal():
    P = [585.26, 0, 317.387, 0, 0, 585.028, 239.264, 0, 0, 0, 1, 0]      # src/urdf_filter.cpp:464-466
    g, tx, ty = orc.projection_matrix(P, 640, 480)
    assert g[0] == -2.0 * 585.26 / 640 and g[5] == 2.0 * 585.028 / 480
    assert g[8] == 2.0 * (0.5 - 317.387 / 640) and g[9] == 2.0 * (239.264 / 480 - 0.5)
    assert g[10] == -(8 + 0.1) / (8 - 0.1) and g[14] == -2.0 * 8 * 0.1 / (8 - 0.1) and g[11] == -1
    assert tx == 0 and ty == 0 and np.count_nonzero(g) == 7
    _, tx, ty = orc.projection_matrix([525, 0, 319.5, -39.375, 0, 525, 239.5, 5.25, 0, 0, 1, 0], 640, 480)
    assert tx == 0.075 and ty == -0.01                                   # -P[3]/fx, -P[7]/fy


def test_lookat_is_diag_m1_1_m1():
    assert np.array_equal(np.abs(orc.lookat()), np.eye(4).reshape(-1))
    assert orc.lookat()[[0, 5, 10, 15]].tolist() == [-1.0, 1.0, -1.0, 1.0]


def test_optical_axis_lands_at_cx_H_minus_cy():
    """F5: x_w = fx*x/z + cx, y_w = fy*y/z + (H - cy).  A square of 0.75 px half-width centred on
    the optical axis spans (318.75, 320.25) x (239.75, 241.25) in window coordinates and therefore
    covers exactly the pixel whose centre is (319.5, 240.5): column 319, row 240."""
    z, r = 2.0, 0.75 * 2.0 / 525.0
    sq = [[-r, -r, z, r, -r, z, r, r, z], [-r, -r, z, r, r, z, -r, r, z]]
    zb = render_tris(sq)
    hit = np.argwhere(zb < orc.render(np.zeros((0, 9), np.float32), np.zeros(0, np.uint32),
                                      orc.compose_mvp(*cam_setup(), np.zeros(16), 0), 640, 480, helpers.BG_Z))
    rows, cols = sorted(set(hit[:, 0])), sorted(set(hit[:, 1]))
    assert cols == [319] and rows == [240]
    # asymmetric check with an off-centre principal point: cy = 200 -> rows around H - cy = 280
    P = synth.kinect_P(640, 480); P[6] = 200.0
    zb2 = render_tris(sq, P=P)
    rows2 = np.argwhere(zb2 < zb2.max())[:, 0]
    assert abs(rows2.mean() - (480 - 200.0 - 0.5)) <= 1.0


@pytest.mark.parametrize("z", [0.11, 0.5, 1.0, 2.5, 5.0, 7.9])
def test_fronto_parallel_plane_depth(z):
    """A plane at eye depth z: to_linear_depth(window z) == z within 25 um (float32 bound)."""
    s = 3.0 * z
    q = [[-s, -s, z, s, -s, z, s, s, z], [-s, -s, z, s, s, z, -s, s, z]]
    zb = render_tris(q)
    virt = synth.linear_depth(zb)
    assert np.all(np.abs(virt - z) < 2.5e-5 * max(1.0, z)), float(np.abs(virt - z).max())


def test_background_quad_is_7_92_everywhere():
    """F3: with no model the virtual depth is 0.99 * far on every pixel."""
    zb = render_tris(np.zeros((0, 9)))
    assert len(np.unique(zb)) == 1 and zb[0, 0] < 1.0
    assert abs(orc.to_linear_depth(float(zb[0, 0])) - 7.92) < 1e-4


def test_background_filters_far_readings_only():
    proj, view = cam_setup()
    mvp = orc.compose_mvp(proj, view, np.zeros(16), 0)
    depth = np.full((480, 640), 3000, np.uint16)
    depth[0, :10] = 7900          # > 7.92 - 0.05
    depth[1, :10] = 7860          # < 7.87: kept
    depth[2, :10] = 0             # invalid: kept (0 > 7.87 is false)
    out, mask, _ = orc.filter_frame(depth, np.zeros((0, 9), np.float32), np.zeros(0, np.uint32), mvp, ZN, ZF,
                                    np.float32(0.05), np.float32(5.0))
    assert set(np.unique(mask)) == {0, 255}                      # F2: wire values 0 / 255
    assert (mask[0, :10] == 255).all() and (out[0, :10] == 5000).all()
    assert (mask[1, :10] == 0).all() and (out[1, :10] == 7860).all()
    assert (mask[2, :10] == 0).all() and (out[2, :10] == 0).all()
    assert mask[3:].max() == 0 and (out[3:] == 3000).all()


def test_shader_is_one_sided():
    """F1: everything at or behind (virtual - max_diff) is filtered, in front is kept."""
    z = 2.0
    s = 10.0
    q = np.float32([[-s, -s, z, s, -s, z, s, s, z], [-s, -s, z, s, s, z, -s, s, z]])
    proj, view = cam_setup()
    mvp = orc.compose_mvp(proj, view, np.eye(4).reshape(-1), 1)
    depth = np.zeros((480, 640), np.float32)
    depth[:, 0:100] = 1.90        # 10 cm in front: kept
    depth[:, 100:200] = 1.96      # within threshold of the surface: filtered
    depth[:, 200:300] = 2.00      # on the surface: filtered
    depth[:, 300:400] = 6.00      # far behind: filtered too (shadow)
    depth[:, 400:500] = np.nan    # invalid: kept as NaN
    depth[:, 500:] = 1.9499       # just in front of virt - 0.05
    out, mask, _ = orc.filter_frame(depth, q, np.zeros(2, np.uint32), mvp, ZN, ZF, np.float32(0.05), np.float32(5.0))
    assert mask[:, 0:100].max() == 0 and np.all(out[:, 0:100] == np.float32(1.90))
    for a in (100, 200, 300):
        assert mask[:, a:a + 100].min() == 255 and np.all(out[:, a:a + 100] == 5.0)
    assert mask[:, 400:500].max() == 0 and np.isnan(out[:, 400:500]).all()
    assert mask[:, 500:].max() == 0


def test_no_mask_requested_and_replace_default_zero():
    proj, view = cam_setup()
    mvp = orc.compose_mvp(proj, view, np.zeros(16), 0)
    depth = np.full((480, 640), 7950, np.uint16)
    out, mask, _ = orc.filter_frame(depth, np.zeros((0, 9), np.float32), np.zeros(0, np.uint32), mvp, ZN, ZF,
                                    np.float32(0.05), np.float32(0.0), want_mask=False)
    assert mask is None and (out == 0).all()       # filter_replace_value defaults to 0 (:110-111)


def test_example_urdf_double_box():
    """F4: RenderableBox draws the (dx,dy,dz) box AND a (dx^2, dx*dy, dx*dz) box.  With only the
    first, wall1 (4 x 0.5 x 2 at (0,5,0)) leaves most of the image at the background; the second
    (16 x 2 x 8 m) swallows the camera, so almost every pixel sees a surface closer than 7.92 m."""
    sc = helpers.scene("example")
    assert sc.n_parts == 4 and sc.n_tris == 48
    zb = helpers.oracle_zbuf(sc, 0)
    virt = synth.linear_depth(zb)
    assert (virt < 7.9).mean() > 0.95
    # without the glutSolidCube parts only the two 4 m walls at ~5 m remain
    keep = np.isin(sc.tri_part, [0, 2])
    view, pm = sc.frame(0)
    z1 = orc.render(sc.tri[keep], sc.tri_part[keep], helpers.oracle_mvp(sc, view, pm), 640, 480, helpers.BG_Z)
    v1 = synth.linear_depth(z1)
    assert 0.05 < (v1 < 7.9).mean() < 0.6
    centre = v1[240, 320]
    assert 4.4 < centre < 5.0           # the wall corner nearest the camera: 5 - 0.25*sqrt(2) .. 5


def test_near_clip_partial_triangle():
    """A triangle crossing the near plane is clipped, not dropped: its visible part starts at 0.1 m."""
    tri = [[-0.5, 0.02, -1.0, 0.5, 0.02, -1.0, 0.0, 0.02, 3.0]]     # from behind the camera to 3 m
    tri2 = [[-0.5, -0.3, -1.0, 0.5, -0.3, -1.0, 0.0, 0.5, 3.0]]
    zb = render_tris(tri2)
    virt = synth.linear_depth(zb)
    hit = virt < 7.9
    assert hit.sum() > 1000
    assert virt[hit].min() >= 0.1 - 1e-4 and virt[hit].min() < 0.2


def test_render_is_order_and_thread_independent():
    sc = helpers.scene("pr2_small")
    view, pm = sc.frame(5)
    mvp = helpers.oracle_mvp(sc, view, pm)
    z1 = orc.render(sc.tri, sc.tri_part, mvp, 640, 480, helpers.BG_Z, nthreads=1)
    z8 = orc.render(sc.tri, sc.tri_part, mvp, 640, 480, helpers.BG_Z, nthreads=8)
    perm = np.random.default_rng(0).permutation(sc.n_tris)
    zp = orc.render(sc.tri[perm], sc.tri_part[perm], mvp, 640, 480, helpers.BG_Z, nthreads=3)
    assert np.array_equal(z1, z8) and np.array_equal(z1, zp)


# ---- the raster kernel's integer form of the shader compare for 16UC1 frames (csrc/ruf_kernels.cu: u16_threshold) ----
def _u16_threshold_np(thr):
    """numpy float32 transcription of u16_threshold(): largest raw with float32(raw) * 0.001f <= thr."""
    thr = np.asarray(thr, np.float32)
    k = np.float32(0.001)
    with np.errstate(invalid="ignore", over="ignore"):
        c = np.minimum(np.floor(thr * np.float32(1000.0)), np.float32(65535.0))
        c = np.where(np.isfinite(c), c, 0).astype(np.int64)
        c = np.clip(c, 0, 65535)
        down = (c.astype(np.float32) * k) > thr
        up = ~down & (c < 65535) & ~(((c + 1).astype(np.float32) * k) > thr)
        c = c - down + up
        c = np.where(thr < np.float32(0.0), -1, c)
        c = np.where(~(thr < np.float32(65.536)), 65535, c)
    return c


def test_u16_threshold_matches_float_compare_on_every_boundary():
    """`float(raw) * 0.001f > thr` (urdf_filter.frag:23 after convertTo(CV_32F, 0.001), src/urdf_filter.cpp:288) must
    equal `raw > R(thr)` for every raw: check thr = the sensor value of every raw and its float neighbours."""
    raw = np.arange(65536, dtype=np.int64)
    g = raw.astype(np.float32) * np.float32(0.001)
    assert np.all(np.diff(g) > 0)                       # strictly increasing: the integer form exists
    cands = [g, np.nextafter(g, np.float32(-np.inf)), np.nextafter(g, np.float32(np.inf)),
             ((g[:-1].astype(np.float64) + g[1:]) / 2).astype(np.float32)]
    rng = np.random.default_rng(3)
    cands.append(rng.uniform(-1, 70, 400000).astype(np.float32))
    cands.append(np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 65.535, 65.536, 1e-30, -1e-30, 7.87, 3e38], np.float32))
    for thr in cands:
        R = _u16_threshold_np(thr)
        # reference: count of raw values that are NOT filtered = first index with g > thr
        with np.errstate(invalid="ignore"):
            want = np.where(np.isnan(thr), 65535, np.searchsorted(g, thr, side="right") - 1)
        assert np.array_equal(R, want)
