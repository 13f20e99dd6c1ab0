"""GPU parity on the other BASELINE.json configurations and on the edge cases of the path:
odd image sizes, no mask, empty model, clipping-heavy views, capacity overflow + retry, models whose
setup CTAs span many parts, more than 512 setup CTAs (multi-round gather), determinism."""
import numpy as np
import pytest

import helpers
import oracle_py as orc
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

pytestmark = pytest.mark.gpu


def run_and_compare(sc, k=0, enc="u16", want_mask=True, ctx=None, max_diff=None, replace_value=None):
    proj, _, _ = sc.proj()
    fr = helpers.make_frame(sc, k, enc, nthreads=8)
    md = sc.max_diff if max_diff is None else max_diff
    rv = sc.replace_value if replace_value is None else replace_value
    want_d, want_m, _ = helpers.oracle_filter(sc, fr, want_mask=want_mask, nthreads=8, max_diff=md, replace_value=rv)
    own = ctx is None
    if own:
        ctx = ruf.Context(sc.width, sc.height)
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    try:
        got_d, got_m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], md, rv, want_mask=want_mask)
        st = ctx.stats()
    finally:
        if own:
            ctx.close()
    assert np.array_equal(got_d.view(np.uint8), want_d.view(np.uint8)), \
        f"depth differs at {np.count_nonzero(got_d != want_d)} px"
    if want_mask:
        assert np.array_equal(got_m, want_m), f"mask differs at {np.count_nonzero(got_m != want_m)} px"
    else:
        assert got_m is None
    return st, want_m


def test_c3_walls_1280x960():
    st, m = run_and_compare(helpers.scene("walls"), k=3)
    assert 0.05 < (m == 255).mean() < 0.95


def test_c5_four_robots_1920x1080_500k_triangles():
    sc = helpers.scene("multi")
    assert sc.n_tris > 490000 and sc.n_parts > 300
    for k in (0, 9):
        st, m = run_and_compare(sc, k=k)
        assert st["visible_tris"] > 10000


@pytest.mark.parametrize("size", [(100, 75), (648, 488), (64, 32), (1, 1), (37, 300), (72, 40)])
def test_odd_image_sizes(size):
    W, H = size
    sc = synth.pr2_like_scene(W, H, n_tris=4000, name=f"odd{W}x{H}")
    run_and_compare(sc, k=2, enc="u16")
    run_and_compare(sc, k=2, enc="f32")


def test_no_mask_and_default_replace_value():
    sc = helpers.scene("pr2_small")
    run_and_compare(sc, k=1, want_mask=False, replace_value=0.0)
    run_and_compare(sc, k=1, enc="f32", want_mask=False, max_diff=0.2, replace_value=-1.0)


def test_empty_model_only_background():
    sc = helpers.scene("example")
    empty = synth.Scene("empty", 640, 480, sc.P, [], [], np.zeros((0, 9), np.float32), np.zeros(0, np.uint32),
                        -1, (0, 0, 0), sc.cam_R)
    depth = np.random.default_rng(0).integers(0, 9000, (480, 640)).astype(np.uint16)
    proj, _, _ = empty.proj()
    view, pm = empty.frame(0)
    mvp = orc.compose_mvp(proj, view, pm, 0)
    want_d, want_m, _ = orc.filter_frame(depth, empty.tri, empty.tri_part, mvp, np.float32(0.1), np.float32(8.0),
                                         np.float32(0.05), np.float32(5.0))
    with ruf.Context(640, 480) as ctx:
        ctx.set_model(empty.tri, empty.tri_part, 0)
        got_d, got_m = ctx.filter(depth, proj, view, pm, 0.05, 5.0)
    assert np.array_equal(got_d, want_d) and np.array_equal(got_m, want_m)
    # float(7870) * 0.001f = 7.87 > to_linear(z_bg) - 0.05 = 7.86996..: 7870 mm itself is already filtered
    assert (got_m == 255).sum() == (depth >= 7870).sum()


def test_filter_before_set_model_is_an_error():
    with ruf.Context(64, 48) as ctx:
        with pytest.raises(ruf.RufError) as e:
            ctx.filter(np.zeros((48, 64), np.uint16), np.zeros(16), np.zeros(16), np.zeros(0), 0.05, 5.0)
        assert e.value.code == ruf.RUF_ERR_NO_MODEL


def _camera_in_the_arm_scene():
    """The camera sits inside the left upper arm: ~700 triangles cross the near plane (clipper output ->
    big list), many are huge on screen."""
    sc = synth.pr2_like_scene(640, 480, n_tris=60000, name="pr2_clip")
    sc.cam_link = [l.name for l in sc.links].index("r0/l_upper_arm")
    sc.cam_xyz = (0.15, 0.0, 0.0)
    return sc


def test_clipping_heavy_view_and_big_list_growth():
    sc = _camera_in_the_arm_scene()
    st, m = run_and_compare(sc, k=4)
    assert st["big_tris"] > 20      # bg quad (2) + what the clipper emitted inside the viewport
    # force the big list and the reference buffer to overflow: the host call grows them and retries
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.reserve(1, big_capacity=2, bin_capacity=64)
        run_and_compare(sc, k=4, ctx=ctx)
        run_and_compare(sc, k=5, ctx=ctx, enc="f32")


def test_device_call_reports_overflow_then_succeeds_after_growth():
    import torch
    sc = helpers.scene("pr2_small")
    proj, _, _ = sc.proj()
    fr = helpers.make_frame(sc, 0, "u16")
    want_d, want_m, _ = helpers.oracle_filter(sc, fr)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in, d_proj, d_view, d_pm = t(fr["depth"].view(np.int16)), t(proj), t(fr["view"]), t(fr["pm"])
    d_out = torch.empty_like(d_in)
    d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.reserve(1, big_capacity=1024, bin_capacity=32)
        args = (1, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
        ctx.filter_batch_device(*args)
        with pytest.raises(ruf.RufError) as e:
            ctx.sync()
        assert e.value.code == ruf.RUF_ERR_OVERFLOW
        for _ in range(12):                      # capacity doubles per failed attempt
            ctx.filter_batch_device(*args)
            try:
                ctx.sync()
                break
            except ruf.RufError as err:
                assert err.code == ruf.RUF_ERR_OVERFLOW
        else:
            pytest.fail("never recovered from overflow")
        assert np.array_equal(d_out.cpu().numpy().view(np.uint16), want_d)
        assert np.array_equal(d_mask.cpu().numpy(), want_m)


def test_one_triangle_per_part_disables_part_culling_but_stays_exact():
    rng = np.random.default_rng(8)
    n = 3000
    sc0 = helpers.scene("example")
    c = rng.uniform([-1.5, -1.0, 0.5], [1.5, 1.0, 4.0], (n, 1, 3))
    tri = (c + rng.normal(0, 0.05, (n, 3, 3))).reshape(n, 9).astype(np.float32)
    links = [synth.Link("l", -1)]
    parts = [synth.Part(0, (0, 0, 0, 1), tuple(rng.normal(0, 0.01, 3))) for _ in range(n)]
    sc = synth.Scene("soup", 640, 480, sc0.P, links, parts, tri, np.arange(n, dtype=np.uint32), -1, (0, 0, 0),
                     np.eye(3))
    run_and_compare(sc, k=0)


def test_more_than_512_setup_ctas_multi_round_gather():
    """> 512 * 1024 triangles: the raster kernel gathers its segment list in several rounds."""
    rng = np.random.default_rng(9)
    n = 600000
    sc0 = synth.example_scene(320, 240)
    c = rng.uniform([-1.2, -0.9, 0.6], [1.2, 0.9, 3.0], (n, 1, 3))
    tri = (c + rng.normal(0, 0.01, (n, 3, 3))).reshape(n, 9).astype(np.float32)
    nparts = 6
    part = np.sort(rng.integers(0, nparts, n)).astype(np.uint32)
    links = [synth.Link("l", -1)]
    parts = [synth.Part(0, (0, 0, 0, 1), (0.01 * i, 0, 0)) for i in range(nparts)]
    sc = synth.Scene("soup600k", 320, 240, sc0.P, links, parts, tri, part, -1, (0, 0, 0), np.eye(3))
    st, _ = run_and_compare(sc, k=0)
    assert st["binned_refs"] > 100000


def test_batched_results_are_deterministic_and_frame_independent():
    import torch
    sc = helpers.scene("pr2_small")
    proj, _, _ = sc.proj()
    ks = list(range(6))
    frames = [helpers.make_frame(sc, k, "f32") for k in ks]
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in = t(np.stack([f["depth"] for f in frames]))
    d_view, d_pm, d_proj = t(np.stack([f["view"] for f in frames])), t(np.stack([f["pm"] for f in frames])), t(proj)
    outs = []
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        for rep in range(3):
            d_out = torch.full_like(d_in, -7.0)
            d_mask = torch.full(d_in.shape, 3, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            ctx.filter_batch_device(len(ks), d_in.data_ptr(), ruf.ENC_F32_M, d_proj.data_ptr(), d_view.data_ptr(),
                                    d_pm.data_ptr(), sc.max_diff, sc.replace_value, d_out.data_ptr(),
                                    d_mask.data_ptr(), 0)
            ctx.sync()
            outs.append((d_out.cpu().numpy(), d_mask.cpu().numpy()))
    for o, m in outs[1:]:
        assert np.array_equal(o.view(np.uint32), outs[0][0].view(np.uint32)) and np.array_equal(m, outs[0][1])
    for i, fr in enumerate(frames):           # each frame of the batch equals its single-frame oracle result
        want_d, want_m, _ = helpers.oracle_filter(sc, fr)
        assert np.array_equal(outs[0][0][i].view(np.uint32), want_d.view(np.uint32))
        assert np.array_equal(outs[0][1][i], want_m)


def test_all_u16_values_pass_through_unfiltered_pixels():
    """The fused 16UC1 path writes the input bits for unfiltered pixels (round-trip identity)."""
    sc = helpers.scene("example")
    empty_tri, empty_part = np.zeros((0, 9), np.float32), np.zeros(0, np.uint32)
    depth = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    P = synth.kinect_P(256, 256)
    proj, tx, ty = ruf.projection_matrix(P, 256, 256)
    view = ruf.view_matrix((0, 0, 0, 1), (0, 0, 0), (0, 0, 0, 1), (0, 0, 0), tx, ty)
    mvp = orc.compose_mvp(proj, view, np.zeros(16), 0)
    want_d, want_m, _ = orc.filter_frame(depth, empty_tri, empty_part, mvp, np.float32(0.1), np.float32(8.0),
                                         np.float32(0.05), np.float32(5.0))
    with ruf.Context(256, 256) as ctx:
        ctx.set_model(empty_tri, empty_part, 0)
        got_d, got_m = ctx.filter(depth, proj, view, np.zeros(0), 0.05, 5.0)
    assert np.array_equal(got_d, want_d) and np.array_equal(got_m, want_m)
    keep = want_m == 0
    assert np.array_equal(got_d[keep], depth[keep]) and keep.sum() > 7000


def test_concurrent_contexts_are_independent():
    """C4 in miniature: several camera streams, one context (and stream) each, driven from separate host
    threads at the same time (ctypes releases the GIL): no process-global state, results stay bit-exact."""
    import threading
    scs = [helpers.scene("pr2_small"), helpers.scene("example"), synth.pr2_like_scene(320, 240, n_tris=3000, name="t3"),
           helpers.scene("pr2_small")]
    jobs = []
    for i, sc in enumerate(scs):
        fr = helpers.make_frame(sc, i + 1, "u16" if i % 2 == 0 else "f32")
        want_d, want_m, _ = helpers.oracle_filter(sc, fr)
        jobs.append((sc, fr, want_d, want_m))
    errors = []

    def worker(sc, fr, want_d, want_m):
        try:
            proj, _, _ = sc.proj()
            with ruf.Context(sc.width, sc.height) as ctx:
                ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
                for _ in range(25):
                    d, m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
                    if not (np.array_equal(d.view(np.uint8), want_d.view(np.uint8)) and np.array_equal(m, want_m)):
                        errors.append(f"{sc.name}: mismatch")
                        return
        except Exception as e:          # noqa: BLE001
            errors.append(f"{sc.name}: {e!r}")

    threads = [threading.Thread(target=worker, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def _facing_scene():
    """Meshes that defeat every assumption of the raster kernel's drawing-order hint (front faces first, the
    rest depth-culled): a blob wound clockwise, a mirrored blob (negative scale), an open half blob seen from
    its inside, two interpenetrating blobs, a doubled (coplanar) blob and one lying across the near plane.  The
    hint may only cost time -- every pixel must still match the oracle."""
    rng = np.random.default_rng(11)
    links = [synth.Link("world", -1, (0, 0, 0), (0, 0, 0))]
    parts, tris, pidx = [], [], []
    blob = lambda r, n: synth.blob_for_budget(rng, r, n)
    m = blob((0.25, 0.2, 0.3), 3000)
    synth.add_mesh(parts, tris, pidx, 0, m[:, [0, 1, 2, 6, 7, 8, 3, 4, 5]], off_t=(-0.7, 0.3, 1.6))      # clockwise
    synth.add_mesh(parts, tris, pidx, 0, blob((0.25, 0.2, 0.3), 3000), scale=(-1.0, 1.0, 1.0), off_t=(0.7, 0.3, 1.6))
    half = blob((0.4, 0.4, 0.4), 4000)
    synth.add_mesh(parts, tris, pidx, 0, half[half[:, [2, 5, 8]].max(1) <= 0.0], off_t=(0.0, -0.4, 1.2))  # open, inside visible
    synth.add_mesh(parts, tris, pidx, 0, blob((0.3, 0.3, 0.3), 3000), off_t=(-0.15, 0.35, 2.4))
    synth.add_mesh(parts, tris, pidx, 0, blob((0.3, 0.3, 0.3), 3000), off_t=(0.15, 0.35, 2.5))            # interpenetrating
    dbl = blob((0.2, 0.2, 0.2), 2000)
    synth.add_mesh(parts, tris, pidx, 0, np.concatenate([dbl, dbl]), off_t=(0.0, 0.9, 2.0))               # coplanar twice
    synth.add_mesh(parts, tris, pidx, 0, blob((0.3, 0.3, 0.5), 3000), off_t=(0.5, -0.5, 0.3))             # across the near plane
    tri, tp = synth._finish(tris, pidx)
    return synth.Scene("facing", 640, 480, synth.kinect_P(640, 480), links, parts, tri, tp, -1, (0, 0, 0), np.eye(3),
                       label="adversarial winding / mirroring / open meshes")


def test_depth_cull_hint_is_result_neutral_on_adversarial_meshes():
    sc = _facing_scene()
    for enc in ("u16", "f32"):
        st, m = run_and_compare(sc, k=0, enc=enc)
        assert st["visible_tris"] > 5000 and 0.02 < (m == 255).mean() < 0.9


def test_sliced_two_stream_launch_matches_single_launch(monkeypatch):
    """Batches of >= 512 frames are cut into slices that alternate between two streams: same bits as one launch."""
    import torch
    sc = synth.pr2_like_scene(160, 120, n_tris=6000, name="pr2_160x120")
    proj, _, _ = sc.proj()
    n = 530
    views, pms = sc.frames([k % 40 for k in range(n)])
    rng = np.random.default_rng(3)
    depth = rng.integers(0, 4000, (n, sc.height, sc.width)).astype(np.uint16)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in, d_proj, d_view, d_pm = t(depth.view(np.int16)), t(proj), t(views), t(pms)
    outs = []
    for slice_frames in ("256", "0"):
        monkeypatch.setenv("RUF_SLICE_FRAMES", slice_frames)
        d_out = torch.zeros_like(d_in)
        d_mask = torch.zeros(d_in.shape, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        with ruf.Context(sc.width, sc.height) as ctx:
            ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
            ctx.filter_batch_device(n, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                                    sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
            ctx.sync()
            launches = ctx.stats()["kernel_launches"]
        outs.append((d_out.cpu().numpy().view(np.uint16), d_mask.cpu().numpy(), launches))
    assert outs[1][2] in (3, 4) and outs[0][2] == 3 * outs[1][2]     # 3 slices x 4 kernels vs one launch sequence (3 with RUF_CLUSTER=1)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    # and a few frames against the oracle
    for k in (0, 255, 256, 529):
        want_d, want_m, _ = orc.filter_frame(depth[k], sc.tri, sc.tri_part, helpers.oracle_mvp(sc, views[k], pms[k]),
                                             np.float32(synth.Z_NEAR), np.float32(synth.Z_FAR), np.float32(sc.max_diff),
                                             np.float32(sc.replace_value))
        assert np.array_equal(outs[0][0][k], want_d) and np.array_equal(outs[0][1][k], want_m)


@pytest.mark.parametrize("fpc", ["1", "4", "3"])
def test_setup_frame_runs_match_the_oracle_frame_by_frame(monkeypatch, fpc):
    """A setup CTA keeps its meshlet in registers over a run of frames (matrices staged two frames ahead, one
    barrier per frame): force runs of 1 / 4 / 3 frames over a 10-frame batch with moving joints (parts enter and
    leave the view volume between frames) and compare every frame with the oracle."""
    import torch
    monkeypatch.setenv("RUF_SETUP_FRAMES_FORCE", fpc)
    sc = helpers.scene("pr2")
    proj, _, _ = sc.proj()
    ks = [0, 3, 7, 11, 17, 19, 23, 31, 40, 47]
    frames = [helpers.make_frame(sc, k, "u16", nthreads=8) for k in ks]
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in = t(np.stack([f["depth"] for f in frames]).view(np.int16))
    d_out = torch.empty_like(d_in)
    d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
    d_z = torch.empty(d_in.shape, dtype=torch.float32, device=dev)
    d_proj, d_view, d_pm = t(proj), t(np.stack([f["view"] for f in frames])), t(np.stack([f["pm"] for f in frames]))
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.filter_batch_device(len(ks), d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                                sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), d_z.data_ptr())
        ctx.sync()
    out, mask, z = d_out.cpu().numpy().view(np.uint16), d_mask.cpu().numpy(), d_z.cpu().numpy()
    for i, fr in enumerate(frames):
        want_d, want_m, want_z = helpers.oracle_filter(sc, fr, nthreads=8)
        assert np.array_equal(z[i].view(np.uint32), want_z.view(np.uint32)), f"frame {i}: z-buffer differs"
        assert np.array_equal(out[i], want_d) and np.array_equal(mask[i], want_m), f"frame {i}"


def test_million_frames_stay_identical():
    """Soak: ~1 M frames through the device path, output of the last launch == output of the first.  (The raster
    kernel's record ring once mistook a not-yet-issued chunk for a landed one -- a parity wait two phases ahead --
    about once per half a million frames, only when the depth-culled pass made batches cheap.)"""
    import torch
    sc = helpers.scene("pr2")
    proj, _, _ = sc.proj()
    n = 256
    views, pms = sc.frames([k % 48 for k in range(n)])
    frames = [helpers.make_frame(sc, k, "u16", nthreads=8)["depth"] for k in range(8)]
    depth = np.stack([frames[k % 8] for k in range(n)])
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in, d_proj, d_view, d_pm = t(depth.view(np.int16)), t(proj), t(views), t(pms)
    d_out = torch.zeros_like(d_in)
    d_mask = torch.zeros(d_in.shape, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.reserve(n)
        args = (n, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
        ctx.filter_batch_device(*args)
        ctx.sync()
        first_d, first_m = d_out.clone(), d_mask.clone()
        for it in range(4000):
            ctx.filter_batch_device(*args)
            if it % 500 == 499:
                ctx.sync()
                assert torch.equal(d_out, first_d) and torch.equal(d_mask, first_m), f"differs after {it + 1} launches"
        ctx.sync()
    assert torch.equal(d_out, first_d) and torch.equal(d_mask, first_m)


@pytest.mark.parametrize("n", [5, 12])
def test_small_launch_soak_cluster_split_variant(n, monkeypatch):
    """Soak of what small launches get (cluster-split raster variant: four CTAs per tile, DSMEM merge, cluster barriers;
    fine meshlet cut): 20000 launches of n frames -- several waves of clusters per launch --, every 1000th compared with
    the first, the first with the oracle."""
    import torch
    monkeypatch.setenv("RUF_CLUSTER", "1")
    monkeypatch.setenv("RUF_FINE_MESHLETS", "1")
    sc = helpers.scene("pr2")
    proj, _, _ = sc.proj()
    frs = [helpers.make_frame(sc, k, "u16", nthreads=8) for k in range(n)]
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in = t(np.stack([f["depth"] for f in frs]).view(np.int16))
    d_proj, d_view, d_pm = t(proj), t(np.stack([f["view"] for f in frs])), t(np.stack([f["pm"] for f in frs]))
    d_out = torch.zeros_like(d_in)
    d_mask = torch.zeros(d_in.shape, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        args = (n, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
        ctx.filter_batch_device(*args)
        ctx.sync()
        first_d, first_m = d_out.clone(), d_mask.clone()
        for i in (0, n - 1):
            want_d, want_m, _ = helpers.oracle_filter(sc, frs[i], nthreads=8)
            assert np.array_equal(first_d[i].cpu().numpy().view(np.uint16), want_d) and np.array_equal(first_m[i].cpu().numpy(), want_m)
        for it in range(20000):
            ctx.filter_batch_device(*args)
            if it % 1000 == 999:
                ctx.sync()
                assert torch.equal(d_out, first_d) and torch.equal(d_mask, first_m), f"differs after {it + 1} launches"
        ctx.sync()
    assert torch.equal(d_out, first_d) and torch.equal(d_mask, first_m)


@pytest.mark.parametrize("enc", ["u16", "f32"])
def test_sensor_values_straddling_the_threshold(enc):
    """Every pixel's sensor depth sits within a few units of its own `virtual - threshold`: the compare of
    urdf_filter.frag:23 (and the raster kernel's integer form of it for 16UC1 runs that share one virtual depth)
    must flip at exactly the oracle's value, for robot, background and never-drawn pixels alike."""
    for sc, k in ((helpers.scene("pr2_small"), 4), (helpers.scene("example"), 0)):
        proj, _, _ = sc.proj()
        view, pm = sc.frame(k)
        z = orc.render(sc.tri, sc.tri_part, helpers.oracle_mvp(sc, view, pm), sc.width, sc.height, helpers.BG_Z, nthreads=8)
        virt = synth.linear_depth(z).astype(np.float32)
        rng = np.random.default_rng(77)
        for md in (0.05, 0.0, -0.013, 2.5):
            thr = virt - np.float32(md)
            if enc == "u16":
                depth = np.clip(np.floor(thr.astype(np.float64) * 1000.0) + rng.integers(-2, 4, thr.shape), 0, 65535)
                depth = depth.astype(np.uint16)
                depth[::7, ::5] = rng.integers(0, 65536, depth[::7, ::5].shape).astype(np.uint16)
            else:
                steps = rng.integers(-2, 3, thr.shape)
                depth = thr.copy()
                for s in range(2):
                    depth = np.where(steps > s, np.nextafter(depth, np.float32(np.inf)), depth)
                    depth = np.where(steps < -s, np.nextafter(depth, np.float32(-np.inf)), depth)
                depth[::9, ::4] = np.nan
            fr = dict(view=view, pm=pm, depth=depth)
            want_d, want_m, _ = helpers.oracle_filter(sc, fr, nthreads=8, max_diff=md)
            with ruf.Context(sc.width, sc.height) as ctx:
                ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
                got_d, got_m = ctx.filter(depth, proj, view, pm, md, sc.replace_value)
            assert np.array_equal(got_m, want_m), f"mask differs at {np.count_nonzero(got_m != want_m)} px (md={md})"
            assert np.array_equal(got_d.view(np.uint8), want_d.view(np.uint8))
            assert 0.02 < (want_m == 255).mean() < 0.98


@pytest.mark.parametrize("enc", ["u16", "f32"])
def test_bit_packed_mask_equals_the_byte_mask(enc):
    """RUF_MASK_BITS (opt-in, 1 bit per pixel on the wire): unpacked it is the 0 / 255 image of the default format, on
    flat tiles, busy tiles and through the batched host call; widths that are not a multiple of 8 are refused."""
    sc = helpers.scene("pr2_small")
    proj, _, _ = sc.proj()
    frs = [helpers.make_frame(sc, k, enc) for k in (2, 9, 30)]
    depth = np.stack([f["depth"] for f in frs])
    views, pms = np.stack([f["view"] for f in frs]), np.stack([f["pm"] for f in frs])
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        want_d, want_m = ctx.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value)
        ctx.set_mask_format(ruf.MASK_BITS)
        got_d, bits = ctx.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value)
        assert bits.shape == (3, sc.height, sc.width // 8)
        one_d, one_bits = ctx.filter(depth[1], proj, views[1], pms[1], sc.max_diff, sc.replace_value)
        ctx.set_mask_format(ruf.MASK_BYTES)
        again_d, again_m = ctx.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value)
    unpacked = np.unpackbits(bits, axis=-1, bitorder="little") * np.uint8(255)
    assert np.array_equal(unpacked, want_m) and np.array_equal(got_d.view(np.uint8), want_d.view(np.uint8))
    assert np.array_equal(np.unpackbits(one_bits, axis=-1, bitorder="little") * np.uint8(255), want_m[1])
    assert np.array_equal(again_m, want_m)
    for i, f in enumerate(frs):
        _, om, _ = helpers.oracle_filter(sc, f)
        assert np.array_equal(unpacked[i], om)
    with ruf.Context(100, 75) as ctx:
        with pytest.raises(ruf.RufError) as e:
            ctx.set_mask_format(ruf.MASK_BITS)
        assert e.value.code == ruf.RUF_ERR_INVALID


def _pinned(shape, dtype):
    import ctypes
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = ruf.host_alloc(n)
    buf = (ctypes.c_uint8 * n).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), ptr


@pytest.mark.parametrize("direct", ["7", "3", "0", "5"])
def test_single_frame_graph_path_with_pinned_buffers(direct, monkeypatch):
    """ruf_filter with pinned host buffers runs as ONE captured CUDA graph.  RUF_DIRECT = 7 (default): three kernel nodes, the
    kernels read / write the caller's mapped buffers, the matrices and the status words themselves; 3: matrices and status
    by copy nodes; 0: every buffer by copy nodes (uploads, memset, kernels, read-backs); 5: a mix.  Same results as the
    staged pipeline and the oracle: first call (capture), same buffers again, other buffers (copy nodes / the raster
    kernel's arguments retargeted), other shader scalars / encoding / no mask (re-capture)."""
    monkeypatch.setenv("RUF_DIRECT", direct)
    sc = helpers.scene("pr2_small")
    proj, _, _ = sc.proj()
    lib = ruf.load()
    bufs = []
    try:
        sets = []
        for _ in range(2):
            d_in, p1 = _pinned((sc.height, sc.width), np.uint16)
            d_out, p2 = _pinned((sc.height, sc.width), np.uint16)
            m_out, p3 = _pinned((sc.height, sc.width), np.uint8)
            bufs += [p1, p2, p3]
            sets.append((d_in, d_out, m_out))
        f_in, p4 = _pinned((sc.height, sc.width), np.float32)
        f_out, p5 = _pinned((sc.height, sc.width), np.float32)
        bufs += [p4, p5]
        with ruf.Context(sc.width, sc.height) as ctx:
            ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)

            def call(d_in, d_out, m_out, fr, enc, md):
                v = np.ascontiguousarray(fr["view"], np.float64)
                pm = np.ascontiguousarray(fr["pm"], np.float64)
                pr = np.ascontiguousarray(proj, np.float64)
                rc = lib.ruf_filter(ctx._h, d_in.ctypes.data, enc, pr.ctypes.data, v.ctypes.data, pm.ctypes.data, md,
                                    sc.replace_value, d_out.ctypes.data, m_out.ctypes.data if m_out is not None else None)
                assert rc == 0, lib.ruf_last_error(ctx._h)
                return ctx.stats()

            for i, k in enumerate((0, 5, 5, 12, 20)):
                d_in, d_out, m_out = sets[i % 2] if i != 2 else sets[1]
                fr = helpers.make_frame(sc, k, "u16")
                d_in[:] = fr["depth"]
                d_out[:] = 0xFFFF
                m_out[:] = 7
                st = call(d_in, d_out, m_out, fr, ruf.ENC_U16_MM, sc.max_diff)
                assert st["kernel_launches"] == 3 and st["d2h_bytes"] == sc.width * sc.height * 3
                want_d, want_m, _ = helpers.oracle_filter(sc, fr)
                assert np.array_equal(d_out, want_d) and np.array_equal(m_out, want_m)
            # other threshold -> re-captured graph; then no mask; then 32FC1
            fr = helpers.make_frame(sc, 3, "u16")
            d_in, d_out, m_out = sets[0]
            d_in[:] = fr["depth"]
            call(d_in, d_out, m_out, fr, ruf.ENC_U16_MM, 0.2)
            want_d, want_m, _ = helpers.oracle_filter(sc, fr, max_diff=0.2)
            assert np.array_equal(d_out, want_d) and np.array_equal(m_out, want_m)
            m_out[:] = 9
            call(d_in, d_out, None, fr, ruf.ENC_U16_MM, 0.2)
            assert np.array_equal(d_out, want_d) and np.all(m_out == 9)
            fr = helpers.make_frame(sc, 3, "f32")
            f_in[:] = fr["depth"]
            call(f_in, f_out, m_out, fr, ruf.ENC_F32_M, sc.max_diff)
            want_d, want_m, _ = helpers.oracle_filter(sc, fr)
            assert np.array_equal(f_out.view(np.uint32), want_d.view(np.uint32)) and np.array_equal(m_out, want_m)
            # pageable buffers still take the staged pipeline on the same context
            got_d, got_m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
            assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32)) and np.array_equal(got_m, want_m)
    finally:
        for p in bufs:
            ruf.host_free(p)


def test_single_frame_graph_on_a_callers_stream():
    """ruf_set_stream + ruf_filter with pinned buffers: the graph is captured on the context's own (idle) stream and
    launched on the caller's, behind whatever the caller queued there (here: a device batch that uses the same workspace)."""
    import torch
    sc = helpers.scene("pr2_small")
    proj, _, _ = sc.proj()
    lib = ruf.load()
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    frs = [helpers.make_frame(sc, k, "u16") for k in (2, 7, 13)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in = t(np.stack([f["depth"] for f in frs]).view(np.int16)); d_out = torch.empty_like(d_in)
    d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
    d_proj, d_view, d_pm = t(proj), t(np.stack([f["view"] for f in frs])), t(np.stack([f["pm"] for f in frs]))
    h_in = torch.empty((sc.height, sc.width), dtype=torch.int16).pin_memory()
    h_out, h_mask = torch.empty_like(h_in).pin_memory(), torch.empty((sc.height, sc.width), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx, torch.cuda.stream(stream):
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.set_stream(stream.cuda_stream)
        for rep in range(3):
            ctx.filter_batch_device(3, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                                    sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)      # asynchronous
            fr = frs[rep]
            h_in.numpy()[:] = fr["depth"].view(np.int16)
            v, pm, pr = [np.ascontiguousarray(a, np.float64) for a in (fr["view"], fr["pm"], proj)]
            rc = lib.ruf_filter(ctx._h, h_in.data_ptr(), ruf.ENC_U16_MM, pr.ctypes.data, v.ctypes.data, pm.ctypes.data, sc.max_diff,
                                sc.replace_value, h_out.data_ptr(), h_mask.data_ptr())
            assert rc == 0, lib.ruf_last_error(ctx._h)
            want_d, want_m, _ = helpers.oracle_filter(sc, fr)
            assert np.array_equal(h_out.numpy().view(np.uint16), want_d) and np.array_equal(h_mask.numpy(), want_m)
        ctx.sync()
    for i, fr in enumerate(frs):
        want_d, want_m, _ = helpers.oracle_filter(sc, fr)
        assert np.array_equal(d_out[i].cpu().numpy().view(np.uint16), want_d) and np.array_equal(d_mask[i].cpu().numpy(), want_m)


@pytest.mark.parametrize("env", [{}, {"RUF_HOST_STAGING": "0"}, {"RUF_NO_GRAPH": "1"}])
@pytest.mark.parametrize("enc", ["u16", "f32"])
def test_pageable_single_frame_paths(env, enc, monkeypatch):
    """ruf_filter with pageable buffers (numpy arrays): by default through the context's page-locked staging around the
    single-frame graph, with RUF_HOST_STAGING=0 or RUF_NO_GRAPH=1 through the staged pipeline -- the same bits."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sc = helpers.scene("pr2_small")
    lib = ruf.load()
    a = np.zeros(16, np.uint8)
    assert lib.ruf_host_is_pinned(a.ctypes.data) == 0 and lib.ruf_host_is_pinned(None) == 0
    p = ruf.host_alloc(64)
    assert lib.ruf_host_is_pinned(p) == 1 and lib.ruf_host_is_pinned(p + 32) == 1
    ruf.host_free(p)
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        for k in (1, 6, 6):
            run_and_compare(sc, k=k, enc=enc, ctx=ctx)
        run_and_compare(sc, k=2, enc=enc, ctx=ctx, want_mask=False)


@pytest.mark.parametrize("bg_cache", ["1", "0"])
def test_single_frame_graph_background_seed_follows_the_projection_matrix(bg_cache, monkeypatch):
    """The single-frame graph seeds the big list with the background quad's records, set up once per projection matrix
    (ruf_api.cu fill_bg_seed).  A caller that changes camera_info between frames -- other focal lengths, other principal
    point, back again -- gets the records of the matrix it passed: oracle-exact every time, with the cache and without."""
    monkeypatch.setenv("RUF_BG_CACHE", bg_cache)
    import dataclasses
    sc = helpers.scene("pr2_small")
    P2 = np.array(sc.P, dtype=np.float64, copy=True)
    P2[0] *= 0.7; P2[5] *= 0.85; P2[2] += 31.0; P2[6] -= 17.0
    sc2 = dataclasses.replace(sc, P=P2)
    assert not np.array_equal(sc.proj()[0], sc2.proj()[0])
    lib = ruf.load()
    bufs = []
    try:
        d_in, p1 = _pinned((sc.height, sc.width), np.uint16)
        d_out, p2 = _pinned((sc.height, sc.width), np.uint16)
        m_out, p3 = _pinned((sc.height, sc.width), np.uint8)
        bufs += [p1, p2, p3]
        with ruf.Context(sc.width, sc.height) as ctx:
            ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
            for s, k in ((sc, 0), (sc, 3), (sc2, 4), (sc2, 9), (sc, 6)):
                fr = helpers.make_frame(s, k, "u16")
                d_in[:] = fr["depth"]; d_out[:] = 0xFFFF; m_out[:] = 7
                pr = np.ascontiguousarray(s.proj()[0], np.float64)
                v, pm = np.ascontiguousarray(fr["view"], np.float64), np.ascontiguousarray(fr["pm"], np.float64)
                rc = lib.ruf_filter(ctx._h, d_in.ctypes.data, ruf.ENC_U16_MM, pr.ctypes.data, v.ctypes.data, pm.ctypes.data,
                                    sc.max_diff, sc.replace_value, d_out.ctypes.data, m_out.ctypes.data)
                assert rc == 0, lib.ruf_last_error(ctx._h)
                want_d, want_m, _ = helpers.oracle_filter(s, fr)
                assert np.array_equal(d_out, want_d) and np.array_equal(m_out, want_m), (k, bg_cache)
    finally:
        for p in bufs:
            ruf.host_free(p)


@pytest.mark.parametrize("mode", ["0", "1", "auto"])
def test_both_raster_kernel_variants_are_bit_exact(monkeypatch, mode):
    """ruf_raster_filter_kernel<ENC, MP>: records above kMaxUnits units either parked for the cooperative footprint walk
    (MP = 0) or dealt out over several unit passes (MP = 1); the host picks per launch from the previous launches'
    statistics (RUF_MULTIPASS forces one).  Same bits either way, on a model whose records are mostly wide (6000
    triangles at 640x480) and on the 90k one (none are), across the switch-over of the automatic mode."""
    if mode != "auto":
        monkeypatch.setenv("RUF_MULTIPASS", mode)
    else:
        monkeypatch.delenv("RUF_MULTIPASS", raising=False)
    for name in ("pr2_small", "pr2"):
        sc = helpers.scene(name)
        proj, _, _ = sc.proj()
        frs = [helpers.make_frame(sc, k, "u16") for k in (1, 6, 14)]
        depth = np.stack([f["depth"] for f in frs])
        views, pms = np.stack([f["view"] for f in frs]), np.stack([f["pm"] for f in frs])
        with ruf.Context(sc.width, sc.height) as ctx:
            ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
            for rep in range(3):       # automatic mode: the first call runs the default variant, later ones what the statistics say
                got_d, got_m = ctx.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value)
                for i, f in enumerate(frs):
                    want_d, want_m, _ = helpers.oracle_filter(sc, f)
                    assert np.array_equal(got_d[i], want_d) and np.array_equal(got_m[i], want_m), (name, rep, i)


def test_device_batch_larger_than_the_workspace_budget_runs_as_sub_batches():
    """The record lists are sized per (frame, tile), so the workspace grows with the batch; it is capped at a byte budget
    (RUF_WORKSPACE_GB) and a device batch that does not fit runs as equal sub-batches through the same workspace."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent('''
        import os, sys
        import numpy as np, torch
        sys.path[:0] = [%r, %r, %r]
        import helpers, realtime_urdf_filter_b200 as ruf
        sc = helpers.scene("pr2_small")
        proj, _, _ = sc.proj()
        ks = list(range(11))
        frs = [helpers.make_frame(sc, k, "u16") for k in ks]
        dev = torch.device("cuda:0")
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d_in = t(np.stack([f["depth"] for f in frs]).view(np.int16)); d_out = torch.zeros_like(d_in)
        d_mask = torch.zeros(d_in.shape, dtype=torch.uint8, device=dev)
        d_proj, d_view, d_pm = t(proj), t(np.stack([f["view"] for f in frs])), t(np.stack([f["pm"] for f in frs]))
        torch.cuda.synchronize()
        with ruf.Context(sc.width, sc.height) as ctx:
            ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
            ctx.filter_batch_device(len(ks), d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                                    sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
            ctx.sync()
            st = ctx.stats()
        print("LAUNCHES", st["kernel_launches"])
        for i, fr in enumerate(frs):
            want_d, want_m, _ = helpers.oracle_filter(sc, fr)
            assert np.array_equal(d_out[i].cpu().numpy().view(np.uint16), want_d), i
            assert np.array_equal(d_mask[i].cpu().numpy(), want_m), i
        print("OK")
    ''') % (helpers.__file__.rsplit("/", 2)[0], helpers.__file__.rsplit("/", 2)[0] + "/oracle", helpers.__file__.rsplit("/", 1)[0])
    # pr2_small at 640x480: 80 tiles x 2048 records x 32 B = 5.2 MB per frame -> a 0.02 GB budget holds 3 frames
    env = dict(__import__("os").environ, RUF_WORKSPACE_GB="0.02")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr
    launches = int(res.stdout.split("LAUNCHES")[1].split()[0])
    assert launches >= 3 * 3, launches          # 11 frames in sub-batches of at most 3-4 frames: at least 3 launch sequences


def test_copy_ceiling_call_moves_the_same_bytes():
    sc = helpers.scene("pr2_small")
    proj, _, _ = sc.proj()
    frs = [helpers.make_frame(sc, k, "u16") for k in (0, 1, 2, 3, 4)]
    depth = np.stack([f["depth"] for f in frs])
    views, pms = np.stack([f["view"] for f in frs]), np.stack([f["pm"] for f in frs])
    out, mask = np.empty_like(depth), np.empty(depth.shape, np.uint8)
    lib = ruf.load()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value)
        want = ctx.stats()
        pr = np.ascontiguousarray(proj, np.float64)
        rc = lib.ruf_host_copy_ceiling(ctx._h, 5, depth.ctypes.data, ruf.ENC_U16_MM, pr.ctypes.data, views.ctypes.data,
                                       pms.ctypes.data, out.ctypes.data, mask.ctypes.data)
        assert rc == 0
        got = ctx.stats()
    assert got["h2d_bytes"] == want["h2d_bytes"] and got["d2h_bytes"] == want["d2h_bytes"] and got["kernel_launches"] == 0
