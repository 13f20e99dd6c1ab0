"""Forward kinematics: the oracle's restatement against an independent numpy FK (CPU), and the device
kernels against the oracle bit for bit (GPU)."""
import math

import numpy as np
import pytest

import helpers
import oracle_py as orc
import realtime_urdf_filter_b200 as ruf


def test_oracle_sincos_accuracy():
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.linspace(-10, 10, 4001), rng.uniform(-3000, 3000, 4000), [0.0, math.pi / 4, -math.pi / 2]])
    err = max(max(abs(orc.sincos(x)[0] - math.sin(x)), abs(orc.sincos(x)[1] - math.cos(x))) for x in xs)
    assert err < 4e-16
    assert orc.sincos(0.0) == (0.0, 1.0)


@pytest.mark.parametrize("name", ["pr2_small", "example", "multi"])
def test_oracle_fk_matches_numpy_fk(name):
    sc = helpers.scene(name)
    kin = sc.kinematics()
    _, tx, ty = sc.proj()
    kin["tx"], kin["ty"] = tx, ty
    for k in (0, 7):
        links, pm, view = orc.fk(kin, sc.joint_q(k))
        Ts = sc.link_poses(k)
        for i, T in enumerate(Ts):
            assert np.allclose(links[i].reshape(4, 4).T, T, atol=1e-12)
        want_view, want_pm = sc.frame(k)             # quaternion route of the TF-fed entry point
        assert np.allclose(pm, want_pm, atol=1e-11) and np.allclose(view, want_view, atol=1e-11)


@pytest.mark.gpu
def test_device_fk_bit_exact_and_filter_through_fk():
    import torch
    sc = helpers.scene("pr2_small")
    kin = sc.kinematics()
    proj, tx, ty = sc.proj()
    kin["tx"], kin["ty"] = tx, ty
    ks = [0, 3, 11, 29]
    q = np.stack([sc.joint_q(k) for k in ks])
    want = [orc.fk(kin, q[i]) for i in range(len(ks))]
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_q = t(q)
    d_pm = torch.empty((len(ks), sc.n_parts, 16), dtype=torch.float64, device=dev)
    d_view = torch.empty((len(ks), 16), dtype=torch.float64, device=dev)
    frames = []
    for i, k in enumerate(ks):       # sensor frames derived from the oracle-FK poses
        fr = dict(view=want[i][2], pm=want[i][1])
        z = orc.render(sc.tri, sc.tri_part, helpers.oracle_mvp(sc, fr["view"], fr["pm"]), sc.width, sc.height, helpers.BG_Z, nthreads=4)
        from realtime_urdf_filter_b200 import synth
        fr["depth"] = synth.synth_depth(synth.linear_depth(z), k, "u16")
        frames.append(fr)
    d_in = t(np.stack([f["depth"] for f in frames]).view(np.int16))
    d_out = torch.empty_like(d_in)
    d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
    d_proj = t(proj)
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.set_kinematics(**{k_: kin[k_] for k_ in ("parent", "joint_type", "origin", "axis", "part_link", "part_local",
                                                    "cam_link", "cam_mount", "view_pre")})
        ctx.fk_batch_device(len(ks), d_q.data_ptr(), tx, ty, d_pm.data_ptr(), d_view.data_ptr())
        ctx.sync()
        pm, view = d_pm.cpu().numpy(), d_view.cpu().numpy()
        for i in range(len(ks)):
            assert np.array_equal(pm[i].view(np.uint64), want[i][1].view(np.uint64))
            assert np.array_equal(view[i].view(np.uint64), want[i][2].view(np.uint64))
        ctx.filter_batch_device_fk(len(ks), d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_q.data_ptr(), tx, ty,
                                   sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
        ctx.sync()
        assert ctx.stats()["kernel_launches"] in (5, 6)      # 2 FK kernels + 3 (small launch: cluster-split raster variant) or 4
    for i, fr in enumerate(frames):
        want_d, want_m, _ = helpers.oracle_filter(sc, fr)
        assert np.array_equal(d_out[i].cpu().numpy().view(np.uint16), want_d)
        assert np.array_equal(d_mask[i].cpu().numpy(), want_m)


@pytest.mark.gpu
def test_kinematics_argument_checks():
    sc = helpers.scene("pr2_small")
    kin = sc.kinematics()
    with ruf.Context(sc.width, sc.height) as ctx:
        with pytest.raises((ruf.RufError, ValueError)):          # model first
            ctx.set_kinematics(**{k_: kin[k_] for k_ in ("parent", "joint_type", "origin", "axis", "part_link",
                                                        "part_local", "cam_link", "cam_mount", "view_pre")})
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        bad = dict(kin)
        bad["parent"] = kin["parent"].copy()
        bad["parent"][3] = 5                        # child before parent
        with pytest.raises(ruf.RufError) as e:
            ctx.set_kinematics(**{k_: bad[k_] for k_ in ("parent", "joint_type", "origin", "axis", "part_link",
                                                        "part_local", "cam_link", "cam_mount", "view_pre")})
        assert e.value.code == ruf.RUF_ERR_INVALID


@pytest.mark.gpu
def test_fk_buffers_follow_a_larger_model():
    """the library's internal FK buffers are frames x links / frames x parts; a second model with more
    links and parts (same frame count) must get larger buffers, not write past the old ones."""
    import torch
    dev = torch.device("cuda:0")
    KEYS = ("parent", "joint_type", "origin", "axis", "part_link", "part_local", "cam_link", "cam_mount", "view_pre")
    small, large = helpers.scene("example"), helpers.scene("pr2_small")
    assert large.n_parts > small.n_parts
    ks = [1, 5, 8]
    # a context per resolution is avoided on purpose: both scenes are 640x480, ONE context sees both models
    assert (small.width, small.height) == (large.width, large.height)
    with ruf.Context(large.width, large.height) as ctx:
        for sc in (small, large):
            kin = sc.kinematics()
            proj, tx, ty = sc.proj()
            kin["tx"], kin["ty"] = tx, ty
            q = np.stack([sc.joint_q(k) for k in ks])
            want = [orc.fk(kin, q[i]) for i in range(len(ks))]
            frames = []
            from realtime_urdf_filter_b200 import synth
            for i, k in enumerate(ks):
                fr = dict(view=want[i][2], pm=want[i][1])
                z = orc.render(sc.tri, sc.tri_part, helpers.oracle_mvp(sc, fr["view"], fr["pm"]), sc.width, sc.height,
                               helpers.BG_Z, nthreads=4)
                fr["depth"] = synth.synth_depth(synth.linear_depth(z), k, "u16")
                frames.append(fr)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            d_q, d_proj = t(q), t(proj)
            d_in = t(np.stack([f["depth"] for f in frames]).view(np.int16))
            d_out = torch.empty_like(d_in)
            d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
            ctx.set_kinematics(**{k_: kin[k_] for k_ in KEYS})
            ctx.filter_batch_device_fk(len(ks), d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_q.data_ptr(), tx, ty,
                                       sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
            ctx.sync()
            for i, fr in enumerate(frames):
                want_d, want_m, _ = helpers.oracle_filter(sc, fr)
                assert np.array_equal(d_out[i].cpu().numpy().view(np.uint16), want_d)
                assert np.array_equal(d_mask[i].cpu().numpy(), want_m)
