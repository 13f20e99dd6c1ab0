"""Model ingest (CPU): the meshlets ruf_set_model builds expand back to the input soup bit for bit,
respect their limits, and weld shared vertices (the vertex stage then runs once per distinct vertex)."""
import numpy as np
import pytest

import helpers
from realtime_urdf_filter_b200 import _lib


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", ["example", "pr2_small", "pr2"])
def test_roundtrip_is_the_identity(name):
    sc = helpers.scene(name)
    xyz, part, cnt = _lib.meshlet_roundtrip(sc.tri, sc.tri_part, sc.n_parts)
    T = sc.n_tris
    assert cnt["tris"] == T + 2
    assert np.array_equal(_bits(xyz[:T]), _bits(sc.tri.reshape(-1, 9)))
    assert np.array_equal(part[:T], sc.tri_part)
    # background quad, src/urdf_filter.cpp:591-596
    z = np.float32(8.0 * 0.99)
    assert np.array_equal(part[T:], [sc.n_parts, sc.n_parts])
    assert np.array_equal(xyz[T], np.array([-100, -100, z, 100, -100, z, 100, 100, z], np.float32))
    assert np.array_equal(xyz[T + 1], np.array([-100, -100, z, 100, 100, z, -100, 100, z], np.float32))


def test_closed_meshes_are_welded():
    sc = helpers.scene("pr2")
    _, _, cnt = _lib.meshlet_roundtrip(sc.tri, sc.tri_part, sc.n_parts)
    # a closed lat/long mesh has V ~ T/2; cutting it into meshlets duplicates the seams
    assert cnt["verts"] < 0.8 * sc.n_tris, cnt
    assert cnt["meshlets"] < sc.n_tris / 256, cnt


@pytest.mark.parametrize("limits", [(3, 1, 1), (4, 2, 1), (16, 7, 2), (256, 512, 32), (1024, 1023, 32)])
def test_limits_and_adversarial_order(limits):
    rng = np.random.default_rng(5)
    # shared vertices from a small pool, parts in random order, NaN / inf / -0.0 vertices, exact duplicates
    pool = rng.standard_normal((40, 3)).astype(np.float32)
    pool[0] = [np.nan, 0, 0]
    pool[1] = [np.inf, 1, 2]
    pool[2] = [-0.0, 0.0, 1.0]
    pool[3] = [0.0, 0.0, 1.0]
    idx = rng.integers(0, 40, size=(700, 3))
    tri = pool[idx].reshape(-1, 9)
    tri[10] = tri[9]
    part = rng.integers(0, 90, size=700).astype(np.uint32)
    xyz, p2, cnt = _lib.meshlet_roundtrip(tri, part, 90, max_verts=limits[0], max_tris=limits[1], max_parts=limits[2])
    assert np.array_equal(_bits(xyz[:700]), _bits(tri))
    assert np.array_equal(p2[:700], part)
    assert cnt["meshlets"] >= 700 / limits[1]


def test_empty_model_is_just_the_background_quad():
    xyz, part, cnt = _lib.meshlet_roundtrip(np.zeros((0, 9), np.float32), np.zeros(0, np.uint32), 0)
    assert cnt == dict(meshlets=1, verts=4, tris=2) and np.array_equal(part, [0, 0])


@pytest.mark.parametrize("name", ["pr2_small", "example"])
def test_both_cuts_of_the_model_hold_every_triangle_once(name):
    """ruf_set_model keeps two cuts of the model in one set of arrays (throughput cut, then the fine cut that launches of
    one frame use, its offsets shifted behind the first): each expands to the input soup bit for bit + the background quad."""
    sc = helpers.scene(name)
    xyz, part, cnt = _lib.meshlet_sets_roundtrip(sc.tri, sc.tri_part, sc.n_parts)
    T = len(sc.tri)
    for cut in range(2):
        assert np.array_equal(_bits(xyz[cut, :T]), _bits(np.asarray(sc.tri, np.float32).reshape(-1, 9)))
        assert np.array_equal(part[cut, :T], sc.tri_part)
        assert np.all(part[cut, T:] == sc.n_parts)                  # the background quad, last in both cuts
    assert cnt["tris"] == 2 * (T + 2) and cnt["fine_meshlets"] >= cnt["meshlets"] and cnt["fine_meshlets"] >= (T + 2) / 256
