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
