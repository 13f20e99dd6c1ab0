"""Seeded random triangle soups through the CUDA path against the oracle, bit for bit (depth, mask AND the raw z-buffer):
what no hand-made scene covers -- slivers, huge and sub-pixel triangles, degenerate and non-finite vertices, triangles
through the near plane and far outside the guard band, parts whose matrices mirror or shear, image sizes that are not
multiples of anything, both encodings, all four kinds of instantiations of the raster kernel (RUF_MULTIPASS x
RUF_CLUSTER: the automatic choice would give these small launches the cluster-split variant only)."""
import numpy as np
import pytest

import oracle_py as orc
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

pytestmark = pytest.mark.gpu


def _soup(rng, n_parts):
    tris, parts = [], []

    def add(t, p):
        tris.append(np.asarray(t, np.float32).reshape(-1, 9))
        parts.append(np.full(len(tris[-1]), p, np.uint32))
    for p in range(n_parts):
        kind = rng.integers(0, 6)
        c = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.0, 1.0), rng.uniform(0.05, 6.0)])
        n = int(rng.integers(20, 400))
        if kind == 0:      # cloud of small triangles
            v = c + rng.normal(0, 0.25, (n, 1, 3)) + rng.normal(0, 0.02, (n, 3, 3))
        elif kind == 1:    # long thin slivers
            a = c + rng.normal(0, 0.3, (n, 3))
            d = rng.normal(0, 1.0, (n, 3))
            v = np.stack([a, a + d, a + d * 1.001 + rng.normal(0, 0.002, (n, 3))], 1)
        elif kind == 2:    # a few window-sized triangles, some through the near plane / behind the camera
            n = int(rng.integers(2, 12))
            v = rng.uniform(-6, 6, (n, 3, 3)) + np.array([0, 0, rng.uniform(-1.0, 3.0)])
        elif kind == 3:    # sub-pixel triangles far away
            v = np.array([0, 0, 6.5]) + rng.normal(0, 1.5, (n, 1, 3)) * [1, 1, 0.1] + rng.normal(0, 0.004, (n, 3, 3))
        elif kind == 4:    # a closed blob (depth cull / front-back runs)
            v = synth.blob_mesh(rng, (0.3, 0.25, 0.35), 14, 10).reshape(-1, 3, 3) + c
        else:              # degenerate and hostile input
            v = c + rng.normal(0, 0.2, (n, 3, 3))
            v[::5, 1] = v[::5, 0]                                   # zero area
            v[1::7, 2] = (v[1::7, 0] + v[1::7, 1]) / 2              # collinear
            v[2::11, 0, 0] = np.nan
            v[3::13, 1, 2] = np.inf
            v[4::17] *= 1e6                                         # far outside the guard band
        add(v, p)
    return np.concatenate(tris), np.concatenate(parts)


def _part_models(rng, n_parts):
    pm = np.zeros((n_parts, 16))
    for p in range(n_parts):
        M = np.eye(4)
        A = np.eye(3) + rng.normal(0, 0.15, (3, 3))
        if rng.random() < 0.3:
            A[:, 0] *= -1                                            # mirrored part: winding flips
        M[:3, :3] = A
        M[:3, 3] = rng.normal(0, 0.2, 3)
        pm[p] = M.T.reshape(-1)                                      # column-major
    return pm


@pytest.mark.parametrize("cluster", ["0", "1"])
@pytest.mark.parametrize("mode", ["0", "1"])
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_soups_match_the_oracle(monkeypatch, mode, seed, cluster):
    monkeypatch.setenv("RUF_MULTIPASS", mode)
    monkeypatch.setenv("RUF_CLUSTER", cluster)
    monkeypatch.setenv("RUF_FINE_MESHLETS", cluster)
    rng = np.random.default_rng(1000 + seed)
    W, H = [(640, 480), (200, 151), (333, 77), (96, 200), (1280, 96), (64, 64)][seed - 1]
    n_parts = int(rng.integers(3, 14))
    tri, part = _soup(rng, n_parts)
    P = synth.kinect_P(W, H, fx=float(rng.uniform(0.5, 1.6)) * 525.0 * W / 640.0)
    proj = orc.projection_matrix(P, W, H)[0]
    ex = synth.example_scene()
    Tinv = np.linalg.inv(synth.make_T(ex.cam_R, (0.0, 0.0, 0.0)))
    # the soup is modelled in camera coordinates (x right, y down, z forward): bring it into the world the view undoes
    view = ruf.view_matrix((0, 0, 0, 1), (0, 0, 0), synth.quat_from_matrix(Tinv[:3, :3]), Tinv[:3, 3], 0.0, 0.0)
    world_from_cam = synth.make_T(ex.cam_R, (0.0, 0.0, 0.0)).T.reshape(-1)
    with ruf.Context(W, H) as ctx:
        ctx.set_model(tri, part, n_parts)
        for frame in range(2):
            pm = _part_models(rng, n_parts)
            pm = np.stack([(world_from_cam.reshape(4, 4).T @ m.reshape(4, 4).T).T.reshape(-1) for m in pm])
            mvp = orc.compose_mvp(proj, view, pm, n_parts)
            for enc in ("u16", "f32"):
                if enc == "u16":
                    depth = rng.integers(0, 9000, (H, W)).astype(np.uint16)
                else:
                    depth = rng.uniform(0.0, 9.0, (H, W)).astype(np.float32)
                    depth[rng.random((H, W)) < 0.05] = np.nan
                want_d, want_m, want_z = orc.filter_frame(depth, tri, part, mvp, np.float32(0.1), np.float32(8.0),
                                                          np.float32(0.05), np.float32(5.0), want_mask=True,
                                                          want_zbuf=True, nthreads=8)
                got_d, got_m = ctx.filter(depth, proj, view, pm, 0.05, 5.0)
                assert np.array_equal(got_m, want_m), (seed, frame, enc, int(np.count_nonzero(got_m != want_m)))
                assert np.array_equal(got_d.view(np.uint8), want_d.view(np.uint8)), (seed, frame, enc)
            covered = float((want_z < np.float32(0.98)).mean())
            assert covered > 0.005, covered                           # the soup is actually in view
