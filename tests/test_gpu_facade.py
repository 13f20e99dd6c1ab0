"""GPU: the C++ RealtimeURDFFilter facade driven like the ROS node (params, TF, Image+CameraInfo)
against the oracle -- reads like a test of the reference's own class."""
import numpy as np
import pytest

import helpers
import oracle_py as orc
from realtime_urdf_filter_b200 import facade, synth

pytestmark = pytest.mark.gpu

PARAMS = {"fixed_frame": "/world", "camera_frame": "/camera_rgb_optical_frame",
          "depth_distance_threshold": 0.05, "filter_replace_value": 5.0, "show_gui": False,
          "robot_description": synth.example_urdf_xml()}
MODELS = [{"model": "robot_description", "tf_prefix": "/EXAMPLE", "geometry_type": "visual", "scale": 1.0}]


def make_node(sc, **extra):
    n = facade.FilterNode(dict(PARAMS, **extra), MODELS, camera_offset=((0, 0, 0), (0, 0, 0, 1)))
    Ts = sc.link_poses(0)
    n.set_tf("/world", (0, 0, 0, 1), (0, 0, 0))
    for ln, T in zip(sc.links, Ts):
        n.set_tf("/EXAMPLE/" + ln.name, synth.quat_from_matrix(T[:3, :3]), T[:3, 3])
    Tc = synth.make_T(sc.cam_R, sc.cam_xyz)
    n.set_tf("/camera_rgb_optical_frame", synth.quat_from_matrix(Tc[:3, :3]), Tc[:3, 3])
    return n


@pytest.mark.parametrize("pinned_staging", [True, False])
@pytest.mark.parametrize("enc", ["u16", "f32"])
def test_filter_callback_matches_oracle(enc, pinned_staging):
    """pinned_staging (default): the message's pageable data goes through the facade's page-locked staging buffers, so that
    ruf_filter runs its single-frame graph; off: the pageable pointers go straight to ruf_filter (staged pipeline)."""
    sc = helpers.scene("example")
    fr = helpers.make_frame(sc, 0, enc)
    want_d, want_m, _ = helpers.oracle_filter(sc, fr)
    with make_node(sc, pinned_staging=pinned_staging) as n:
        n.callback(fr["depth"], sc.P, stamp=1.0)
        assert n.published() == (1, 1)
        c = n.counts()
        assert c["renderers"] == 1 and c["renderables"] == 2 and c["parts"] == 4 and c["triangles"] == 48
        assert c["tf_lookups"] == 1 + 2          # camera + one per renderable (src/urdf_renderer.cpp:173-190)
        got_d, e = n.last_image(0, fr["depth"].shape, fr["depth"].dtype)
        got_m, em = n.last_image(1, fr["depth"].shape, np.uint8)
        assert e == ("16UC1" if enc == "u16" else "32FC1") and em == "mono8"      # :316, :324
        assert np.array_equal(got_m, want_m)
        assert np.array_equal(got_d.view(np.uint8), want_d.view(np.uint8))


def test_tracker_style_filter_and_masked_depth():
    """src/urdf_filtered_tracker.cpp:161-167,239-241: filter(buffer, glTf, W, H) + getMaskedDepth()."""
    sc = helpers.scene("example")
    fr = helpers.make_frame(sc, 0, "f32")
    want_d, _, _ = helpers.oracle_filter(sc, fr)
    proj, _, _ = sc.proj()
    with make_node(sc) as n:
        got = n.filter(fr["depth"], proj)
        assert np.array_equal(got.view(np.uint32), want_d.view(np.uint32))
        assert n.get("width") == 640 and n.get("height") == 480


def test_need_mask_follows_subscribers_and_depth_only_if_subscribed():
    sc = helpers.scene("example")
    fr = helpers.make_frame(sc, 0, "u16")
    with make_node(sc) as n:
        n.subscribers(depth=0, mask=0)
        n.callback(fr["depth"], sc.P)
        assert n.published() == (0, 0)
        n.subscribers(depth=1, mask=0)
        n.callback(fr["depth"], sc.P)
        assert n.published() == (1, 0)


def test_camera_tf_failure_skips_render_and_link_tf_failure_reuses_previous():
    sc = helpers.scene("example")
    fr = helpers.make_frame(sc, 0, "u16")
    with make_node(sc) as n:
        n.erase_tf("/camera_rgb_optical_frame")
        n.callback(fr["depth"], sc.P)
        assert "does not exist" in n.log() and n.published() == (0, 0)
    with make_node(sc) as n:
        # wall2's TF is missing: the reference silently reuses `t` of the previous loop iteration,
        # i.e. wall2 is drawn with wall1's pose (src/urdf_renderer.cpp:175-188)
        n.erase_tf("/EXAMPLE/wall2")
        n.callback(fr["depth"], sc.P)
        got_m, _ = n.last_image(1, fr["depth"].shape, np.uint8)
        view, pm = sc.frame(0)
        pm2 = pm.copy()
        pm2[2], pm2[3] = pm[0], pm[1]
        fr2 = dict(fr, pm=pm2)
        _, want_m, _ = helpers.oracle_filter(sc, fr2)
        assert np.array_equal(got_m, want_m)


def test_no_models_raises_like_initGL():
    with facade.FilterNode(dict(PARAMS), []) as n:
        n.set_tf("/world", (0, 0, 0, 1), (0, 0, 0))
        with pytest.raises(RuntimeError, match="Could not load any models"):
            n.callback(np.zeros((480, 640), np.uint16), synth.kinect_P(640, 480))


def test_image_size_change_reinitialises():
    sc = helpers.scene("example")
    fr = helpers.make_frame(sc, 0, "u16")
    with make_node(sc) as n:
        n.callback(fr["depth"], sc.P)
        sc2 = synth.example_scene(320, 240)
        fr2 = helpers.make_frame(sc2, 0, "u16")
        want_d, want_m, _ = helpers.oracle_filter(sc2, fr2)
        n.callback(fr2["depth"], sc2.P)
        assert "image size has changed" in n.log()
        got_d, _ = n.last_image(0, (240, 320), np.uint16)
        assert np.array_equal(got_d, want_d)


def test_packed_mask_readback_publishes_the_same_mono8_mask():
    sc = helpers.scene("example")
    fr = helpers.make_frame(sc, 0, "u16")
    _, want_m, _ = helpers.oracle_filter(sc, fr)
    n = facade.FilterNode(dict(PARAMS, packed_mask_readback=True), MODELS, camera_offset=((0, 0, 0), (0, 0, 0, 1)))
    with n:
        Ts = sc.link_poses(0)
        n.set_tf("/world", (0, 0, 0, 1), (0, 0, 0))
        for ln, T in zip(sc.links, Ts):
            n.set_tf("/EXAMPLE/" + ln.name, synth.quat_from_matrix(T[:3, :3]), T[:3, 3])
        Tc = synth.make_T(sc.cam_R, sc.cam_xyz)
        n.set_tf("/camera_rgb_optical_frame", synth.quat_from_matrix(Tc[:3, :3]), Tc[:3, 3])
        n.callback(fr["depth"], sc.P, stamp=1.0)
        got_m, em = n.last_image(1, fr["depth"].shape, np.uint8)
    assert em == "mono8" and np.array_equal(got_m, want_m)


def test_missing_mesh_is_counted_and_strict_meshes_raises():
    xml = ('<robot name="m"><link name="l"><visual><geometry><mesh filename="package://pkg/nope.stl"/></geometry></visual>'
           '</link><link name="b"><visual><geometry><box size="1 1 1"/></geometry></visual></link>'
           '<joint name="j" type="fixed"><parent link="l"/><child link="b"/></joint></robot>')
    params = dict(PARAMS, robot_description=xml)
    depth = np.full((480, 640), 1500, np.uint16)
    with facade.FilterNode(params, MODELS, camera_offset=((0, 0, 0), (0, 0, 0, 1))) as n:
        for f in ("/world", "/EXAMPLE/l", "/EXAMPLE/b", "/camera_rgb_optical_frame"):
            n.set_tf(f, (0, 0, 0, 1), (0, 0, 0))
        n.callback(depth, synth.kinect_P(640, 480))
        assert n.counts()["mesh_errors"] == 1 and "Could not load resource" in n.log()
    with facade.FilterNode(dict(params, strict_meshes=True), MODELS, camera_offset=((0, 0, 0), (0, 0, 0, 1))) as n:
        n.set_tf("/world", (0, 0, 0, 1), (0, 0, 0))
        with pytest.raises(RuntimeError, match="Could not load 1 mesh"):
            n.callback(depth, synth.kinect_P(640, 480))
