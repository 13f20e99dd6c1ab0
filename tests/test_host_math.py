"""Product host-side matrices and tessellations (libruf_b200.so, CPU code) against the oracle,
bit for bit.  These are the doubles the reference computes on the host before glMultMatrixd."""
import numpy as np

import oracle_py as orc
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth


def _rq(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def test_projection_lookat_bit_exact():
    rng = np.random.default_rng(1)
    for _ in range(50):
        P = np.zeros(12)
        P[0], P[5], P[2], P[6] = rng.uniform(200, 1200, 2).tolist() + rng.uniform(100, 900, 2).tolist()
        P[3], P[7] = rng.uniform(-50, 50, 2)
        W, H = int(rng.integers(64, 2000)), int(rng.integers(64, 1200))
        a, atx, aty = ruf.projection_matrix(P, W, H)
        b, btx, bty = orc.projection_matrix(P, W, H)
        assert np.array_equal(a, b) and atx == btx and aty == bty
    assert np.array_equal(ruf.lookat().view(np.uint64), orc.lookat().view(np.uint64))     # incl. signed zeros


def test_view_and_part_model_bit_exact():
    rng = np.random.default_rng(2)
    for _ in range(200):
        oq, cq, lq = _rq(rng), _rq(rng), _rq(rng)
        fq = rng.normal(size=4) * rng.uniform(0.1, 3)      # un-normalised URDF origin quaternion
        ot, ct, lt, ft = (rng.normal(size=3) for _ in range(4))
        tx, ty = rng.normal(size=2) * 0.05
        assert np.array_equal(ruf.view_matrix(oq, ot, cq, ct, tx, ty), orc.view_matrix(oq, ot, cq, ct, tx, ty))
        sfx = synth.scale_suffix(*rng.uniform(0.001, 3, 3)) if rng.random() < 0.5 else None
        assert np.array_equal(ruf.part_model(lq, lt, fq, ft, sfx), orc.link_model(lq, lt, fq, ft, sfx))


def test_view_matrix_semantics():
    # identity everything -> view == LookAt ; camera offset is applied inverted
    v = ruf.view_matrix((0, 0, 0, 1), (0, 0, 0), (0, 0, 0, 1), (0, 0, 0))
    assert np.array_equal(v, ruf.lookat())
    v = ruf.view_matrix((0, 0, 0, 1), (1, 2, 3), (0, 0, 0, 1), (0, 0, 0)).reshape(4, 4).T
    assert np.allclose(v[:3, 3], [1.0, -2.0, 3.0])      # LookAt * translate(-offset) = diag(-1,1,-1) * (-1,-2,-3)
    # tx shifts the camera origin along its own x axis (src/urdf_filter.cpp:607-608)
    v = ruf.view_matrix((0, 0, 0, 1), (0, 0, 0), (0, 0, 0, 1), (0, 0, 0), 0.075, 0.0).reshape(4, 4).T
    assert np.allclose(v[:3, 3], [-0.075, 0.0, 0.0])


def test_primitives_bit_exact():
    for dims in [(4, 0.5, 2), (0.1, 0.2, 0.3), (1, 1, 1)]:
        assert np.array_equal(ruf.box_triangles(*dims), orc.box_triangles(*dims))
        assert np.array_equal(ruf.cube_triangles(dims[0]), orc.cube_triangles(dims[0]))
    for r in (0.05, 0.3, 1.7):
        assert np.array_equal(ruf.sphere_triangles(r), orc.sphere_triangles(r))
        assert np.array_equal(ruf.cylinder_triangles(r, 2 * r + 0.1), orc.cylinder_triangles(r, 2 * r + 0.1))
    assert ruf.sphere_triangles(1.0, 6, 4).shape[0] == 2 * 6 + 2 * 6 * 2
    assert ruf.cylinder_triangles(1.0, 1.0, 5, 3).shape[0] == 2 * 5 + 2 * 5 * 3


def test_synth_scenes_shape():
    sc = synth.pr2_like_scene()
    assert 85 <= sc.n_parts <= 92 and abs(sc.n_tris - 90000) <= 900     # ~88 parts, 90k +- 1 %
    assert sc.tri_part.max() == sc.n_parts - 1
    view, pm = sc.frame(0)
    assert view.shape == (16,) and pm.shape == (sc.n_parts, 16)
    ex = synth.example_scene()
    assert ex.n_tris == 48 and ex.n_parts == 4


def test_binding_keeps_converted_arguments_alive():
    """lists, float32 and transposed (non-contiguous) matrices are copied by the binding; the copies
    must outlive the C call (numpy's small-block cache hands a freed temporary's address to the next one)."""
    import realtime_urdf_filter_b200 as ruf
    rng = np.random.default_rng(11)
    for _ in range(20):
        q1, q2 = rng.normal(size=4), rng.normal(size=4)
        t1, t2 = rng.normal(size=3), rng.normal(size=3)
        want = ruf.view_matrix(q1, t1, q2, t2, 0.1, -0.2)
        got = ruf.view_matrix(list(q1), t1.astype(np.float64)[::-1][::-1].tolist(), tuple(q2), list(t2), 0.1, -0.2)
        assert np.array_equal(want, got)
        # float32 / strided inputs take the copying branch for every argument at once
        q1f, q2f = q1.astype(np.float32), q2.astype(np.float32)
        t1s, t2s = np.stack([t1, t1], 1)[:, 0], np.stack([t2, t2], 1)[:, 0]
        want32 = ruf.view_matrix(q1f.astype(np.float64), t1, q2f.astype(np.float64), t2)
        assert np.array_equal(want32, ruf.view_matrix(q1f, t1s, q2f, t2s))
        want_pm = ruf.part_model(q1, t1, q2, t2)
        assert np.array_equal(want_pm, ruf.part_model(list(q1), list(t1), list(q2), list(t2)))


def test_scenes_do_not_depend_on_the_math_provider():
    """bench.py's reference arm builds its scene with the oracle's host math (so that the CPU arm never loads the
    product library): same triangles, same matrices, bit for bit."""
    import bench
    from realtime_urdf_filter_b200 import synth
    a = synth.pr2_like_scene(n_tris=5000, name="prov_a")
    va, pa = a.frame(7)
    try:
        synth.use_math(bench.OracleMath(orc))
        b = synth.pr2_like_scene(n_tris=5000, name="prov_b")
        vb, pb = b.frame(7)
        ex = synth.example_scene()
        pj = b.proj()
    finally:
        synth.use_math(synth._lib)
    assert np.array_equal(a.tri.view(np.uint32), b.tri.view(np.uint32)) and np.array_equal(a.tri_part, b.tri_part)
    assert np.array_equal(va.view(np.uint64), vb.view(np.uint64)) and np.array_equal(pa.view(np.uint64), pb.view(np.uint64))
    assert np.array_equal(ex.tri.view(np.uint32), synth.example_scene().tri.view(np.uint32))
    assert np.array_equal(np.asarray(pj[0]).view(np.uint64), np.asarray(a.proj()[0]).view(np.uint64))
