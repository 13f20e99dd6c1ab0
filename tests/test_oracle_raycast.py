"""Independent float64 ray-caster vs the oracle's rasteriser (geometry parity of the primitives).

For every pixel centre a ray is cast through the *inverse* of the reference's projection
(x_w = fx x/z + cx, y_w = fy y/z + (H - cy)) and intersected analytically / by Moeller-Trumbore
with the same triangle soup; interior pixels must agree to float32 depth precision, silhouette
pixels (where the nearest hit is within a pixel of an edge) are excluded."""
import numpy as np
import pytest

import helpers
import oracle_py as orc
from realtime_urdf_filter_b200 import synth

W, H = 160, 120


def raycast(tri_world, P):
    """tri_world: (T,3,3) float64 in the camera's optical frame.  Returns eye depth z per pixel (inf = miss)
    and a flag telling whether a second-nearest/edge ambiguity exists within 1.5 px."""
    fx, fy, cx, cy = P[0], P[5], P[2], P[6]
    jj, ii = np.mgrid[0:H, 0:W]
    xw, yw = ii + 0.5, jj + 0.5
    d = np.stack([(xw - cx) / fx, (yw - (H - cy)) / fy, np.ones_like(xw)], -1).reshape(-1, 3)   # ray dirs, z = 1
    best = np.full(d.shape[0], np.inf)
    for t in tri_world:
        e1, e2 = t[1] - t[0], t[2] - t[0]
        pv = np.cross(d, e2)
        det = pv @ e1
        ok = np.abs(det) > 1e-14
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = -t[0]                                # origin at 0
        u = (pv @ tv) * inv
        qv = np.cross(tv, e1)
        v = (d @ qv) * inv
        tt = (e2 @ qv) * inv
        hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (tt > 0.1) & (tt < 8.0)
        best = np.where(hit & (tt < best), tt, best)
    return best.reshape(H, W)


def scene_tris(kind):
    if kind == "sphere":
        tri, model = orc.sphere_triangles(0.4), orc.link_model((0, 0, 0, 1), (0.1, -0.05, 1.5))
    elif kind == "cylinder":
        tri = orc.cylinder_triangles(0.25, 0.9)
        q = (np.sin(0.6) * 0.6, np.sin(0.6) * 0.8, 0.0, np.cos(0.6))
        model = orc.link_model(q, (-0.1, 0.05, 1.8), suffix=synth.translate_suffix(0, 0, -0.45))
    else:
        tri = orc.box_triangles(0.8, 0.5, 0.3)
        q = (0.2, 0.3, 0.1, 0.927)
        model = orc.link_model(q, (0.0, 0.0, 1.2))
    return tri, model


@pytest.mark.parametrize("kind", ["sphere", "cylinder", "box"])
def test_primitive_depth_matches_raycaster(kind):
    P = synth.kinect_P(W, H)
    proj, tx, ty = orc.projection_matrix(P, W, H)
    view = orc.view_matrix((0, 0, 0, 1), (0, 0, 0), (0, 0, 0, 1), (0, 0, 0), tx, ty)
    tri, model = scene_tris(kind)
    mvp = orc.compose_mvp(proj, view, model, 1)
    zb = orc.render(tri, np.zeros(len(tri), np.uint32), mvp, W, H, helpers.BG_Z)
    virt = synth.linear_depth(zb).astype(np.float64)
    M = model.reshape(4, 4).T
    tw = (tri.reshape(-1, 3, 3).astype(np.float64) @ M[:3, :3].T) + M[:3, 3]
    ref = raycast(tw, P)
    hit_r, hit_o = np.isfinite(ref), virt < 7.9
    # interior = a 3x3 neighbourhood agrees about being hit
    def erode(m):
        e = m.copy()
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                e &= np.roll(np.roll(m, dy, 0), dx, 1)
        return e
    interior = erode(hit_r) & erode(hit_o)
    assert interior.sum() > 200
    assert hit_r.sum() > 0 and abs(int(hit_r.sum()) - int(hit_o.sum())) <= 0.02 * hit_r.sum() + 8
    # same surface, same depth: float32 window-z quantisation is < 25 um * z, silhouettes of inner edges aside
    err = np.abs(virt - ref)[interior]
    assert np.quantile(err, 0.99) < 1e-4, float(np.quantile(err, 0.99))
    # coverage agreement away from the silhouette
    assert np.array_equal(hit_r[erode(hit_r) | erode(~hit_r)], hit_o[erode(hit_r) | erode(~hit_r)])


def test_tessellation_counts():
    assert len(orc.sphere_triangles(1.0)) == 180          # glutSolidSphere(r,10,10)
    assert len(orc.cylinder_triangles(1.0, 1.0)) == 220   # glutSolidCylinder(r,l,10,10)
    assert len(orc.box_triangles(1, 2, 3)) == 12 and len(orc.cube_triangles(1)) == 12
    s = orc.sphere_triangles(0.5).reshape(-1, 3)
    assert np.allclose(np.linalg.norm(s, axis=1), 0.5, atol=1e-6)
    c = orc.cylinder_triangles(0.5, 2.0).reshape(-1, 3)
    assert c[:, 2].min() == 0.0 and c[:, 2].max() == 2.0
    b = orc.box_triangles(4, 0.5, 2).reshape(-1, 3)
    assert np.array_equal(np.abs(b).max(0), [2.0, 0.25, 1.0])
    k = orc.cube_triangles(4).reshape(-1, 3)
    assert np.array_equal(np.abs(k).max(0), [2.0, 2.0, 2.0])
