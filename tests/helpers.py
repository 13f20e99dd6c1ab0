"""Shared helpers of the parity tests: run a synthetic scene through the CPU oracle."""
from __future__ import annotations

import functools

import numpy as np

import oracle_py as orc
from realtime_urdf_filter_b200 import synth

BG_Z = np.float32(8.0 * 0.99)


@functools.lru_cache(maxsize=8)
def scene(name: str, **kw):
    if name == "example":
        return synth.example_scene(**kw)
    if name == "pr2":
        return synth.pr2_like_scene(**kw)
    if name == "pr2_small":
        return synth.pr2_like_scene(n_tris=6000, name="pr2_like_small", **kw)
    if name == "walls":
        return synth.walls_scene(**kw)
    if name == "multi":
        return synth.multi_robot_scene(**kw)
    raise KeyError(name)


def oracle_mvp(sc, view, pm):
    proj, _, _ = sc.proj()
    return orc.compose_mvp(proj, view, pm, sc.n_parts)


def oracle_zbuf(sc, k, nthreads=4):
    view, pm = sc.frame(k)
    return orc.render(sc.tri, sc.tri_part, oracle_mvp(sc, view, pm), sc.width, sc.height, BG_Z, nthreads=nthreads)


def make_frame(sc, k, enc="u16", nthreads=4):
    """-> dict(view, pm, depth, zbuf) with a synthetic sensor frame derived from the oracle's
    own virtual depth so that every shader outcome occurs."""
    view, pm = sc.frame(k)
    z = orc.render(sc.tri, sc.tri_part, oracle_mvp(sc, view, pm), sc.width, sc.height, BG_Z, nthreads=nthreads)
    depth = synth.synth_depth(synth.linear_depth(z), k, enc)
    return dict(view=view, pm=pm, depth=depth, zbuf=z)


def oracle_filter(sc, fr, want_mask=True, nthreads=4, max_diff=None, replace_value=None):
    md = sc.max_diff if max_diff is None else max_diff
    rv = sc.replace_value if replace_value is None else replace_value
    out, mask, zbuf = orc.filter_frame(fr["depth"], sc.tri, sc.tri_part, oracle_mvp(sc, fr["view"], fr["pm"]),
                                       np.float32(synth.Z_NEAR), np.float32(synth.Z_FAR), np.float32(md),
                                       np.float32(rv), want_mask=want_mask, nthreads=nthreads, want_zbuf=True)
    return out, mask, zbuf


def fuzz_case(seed, special=False):
    """One hostile random soup of tests/test_gpu_fuzz.py with its matrices and a random float depth image, sized so that the
    reference's GL read-back (default GL_PACK_ALIGNMENT) is happy: W % 4 == 0.  Deterministic in `seed`.
    special: 15 % of the sensor pixels are NaN, +-inf, +-0.0, negative, 3e38 or denormal (what a REP-117 float depth image
    may hold, and what mix() in the reference's shader treats in its own way)."""
    import realtime_urdf_filter_b200 as ruf
    import test_gpu_fuzz as fz
    rng = np.random.default_rng(1000 + seed)
    W, H = [(640, 480), (200, 152), (336, 76), (96, 200), (1280, 96), (64, 64)][(seed - 1) % 6]
    n_parts = int(rng.integers(3, 14))
    tri, part = fz._soup(rng, n_parts)
    P = synth.kinect_P(W, H, fx=float(rng.uniform(0.5, 1.6)) * 525.0 * W / 640.0)
    proj = orc.projection_matrix(P, W, H)[0]
    ex = synth.example_scene()
    Tinv = np.linalg.inv(synth.make_T(ex.cam_R, (0.0, 0.0, 0.0)))
    view = ruf.view_matrix((0, 0, 0, 1), (0, 0, 0), synth.quat_from_matrix(Tinv[:3, :3]), Tinv[:3, 3], 0.0, 0.0)
    world_from_cam = synth.make_T(ex.cam_R, (0.0, 0.0, 0.0)).T.reshape(-1)
    pm = np.stack([(world_from_cam.reshape(4, 4).T @ m.reshape(4, 4).T).T.reshape(-1) for m in fz._part_models(rng, n_parts)])
    depth = rng.uniform(0.0, 9.0, (H, W)).astype(np.float32)
    if special:
        sel = np.random.default_rng(seed).random(depth.shape)
        for lo, hi, v in ((0.00, 0.03, np.nan), (0.03, 0.05, np.inf), (0.05, 0.07, -np.inf), (0.07, 0.09, 0.0), (0.09, 0.11, -1.5),
                          (0.11, 0.12, 3.0e38), (0.12, 0.13, 1e-42), (0.13, 0.14, -1e-42), (0.14, 0.15, -0.0)):
            depth[(sel >= lo) & (sel < hi)] = np.float32(v)
    return dict(W=W, H=H, n_parts=n_parts, tri=tri, part=part, proj=proj, view=view, pm=pm, depth=depth,
                z_near=0.1, z_far=8.0, max_diff=0.05, replace_value=5.0)


def fuzz_oracle(fc, nthreads=8):
    mvp = orc.compose_mvp(fc["proj"], fc["view"], fc["pm"], fc["n_parts"])
    d, m, _ = orc.filter_frame(fc["depth"], fc["tri"], fc["part"], mvp, np.float32(fc["z_near"]), np.float32(fc["z_far"]),
                               np.float32(fc["max_diff"]), np.float32(fc["replace_value"]), want_mask=True, nthreads=nthreads)
    return d, m
