#!/usr/bin/env python
"""Latency of the C++ RealtimeURDFFilter facade driven like the ROS node (pageable message buffers in and out, TF table,
rosparams): filter_callback with a 16UC1 image and filter() + getMaskedDepth() with a float buffer, with and without the
facade's page-locked staging (`pinned_staging`).  PR2-like synthetic model is not expressible as a URDF here, so this uses
example.urdf (48 triangles): the copies and the call overhead are what is being compared."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import helpers
from realtime_urdf_filter_b200 import facade, synth

PARAMS = {"fixed_frame": "/world", "camera_frame": "/camera_rgb_optical_frame", "depth_distance_threshold": 0.05,
          "filter_replace_value": 5.0, "show_gui": False, "robot_description": synth.example_urdf_xml()}
MODELS = [{"model": "robot_description", "tf_prefix": "/EXAMPLE", "geometry_type": "visual", "scale": 1.0}]
sc = helpers.scene("example")
for staging in (True, False):
    n = facade.FilterNode(dict(PARAMS, pinned_staging=staging), MODELS, camera_offset=((0, 0, 0), (0, 0, 0, 1)))
    with n:
        n.set_tf("/world", (0, 0, 0, 1), (0, 0, 0))
        for ln, T in zip(sc.links, sc.link_poses(0)):
            n.set_tf("/EXAMPLE/" + ln.name, synth.quat_from_matrix(T[:3, :3]), T[:3, 3])
        Tc = synth.make_T(sc.cam_R, sc.cam_xyz)
        n.set_tf("/camera_rgb_optical_frame", synth.quat_from_matrix(Tc[:3, :3]), Tc[:3, 3])
        fr16 = helpers.make_frame(sc, 0, "u16"); fr32 = helpers.make_frame(sc, 0, "f32")
        proj, _, _ = sc.proj()
        for name, call in (("filter_callback 16UC1", lambda: n.callback(fr16["depth"], sc.P, stamp=1.0)),
                           ("filter() 32FC1 buffer", lambda: n.filter(fr32["depth"], proj))):
            ts = []
            for k in range(230):
                t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
            ts = np.array(ts[30:]) * 1e6
            print("pinned_staging=%s  %-22s median %.1f us  p99 %.1f us" % (staging, name, np.median(ts), np.percentile(ts, 99)), flush=True)
