"""Section \"hostile random soups\" of profiles/r02_gl_crosscheck.txt: tests/helpers.py-style soups through the reference's shaders on
llvmpipe (oracle/_ref/gl_crosscheck_glx, `make -C oracle/gl_ref glx`) vs the CPU oracle.  TEST INFRASTRUCTURE ONLY."""
import os, sys, subprocess, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/oracle'); sys.path.insert(0, ROOT+'/oracle/gl_ref'); sys.path.insert(0, ROOT+'/tests')
import gl_case, oracle_py as orc
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth
import test_gpu_fuzz as fz
mesa = subprocess.run(["make","-s","-C",ROOT+"/oracle/gl_ref","mesa_dir"],capture_output=True,text=True).stdout.strip()
env = dict(os.environ, LD_LIBRARY_PATH=ROOT+"/oracle/_ref/fakex:"+mesa)
for seed in range(1, 13):
    rng = np.random.default_rng(1000 + seed)
    W, H = [(640, 480), (200, 152), (336, 76), (96, 200), (1280, 96), (64, 64)][(seed - 1) % 6]     # W % 4 == 0 (GL pack alignment in the reference)
    n_parts = int(rng.integers(3, 14))
    tri, part = fz._soup(rng, n_parts)
    P = synth.kinect_P(W, H, fx=float(rng.uniform(0.5, 1.6)) * 525.0 * W / 640.0)
    proj = orc.projection_matrix(P, W, H)[0]
    ex = synth.example_scene()
    Tinv = np.linalg.inv(synth.make_T(ex.cam_R, (0.0, 0.0, 0.0)))
    view = ruf.view_matrix((0, 0, 0, 1), (0, 0, 0), synth.quat_from_matrix(Tinv[:3, :3]), Tinv[:3, 3], 0.0, 0.0)
    world_from_cam = synth.make_T(ex.cam_R, (0.0, 0.0, 0.0)).T.reshape(-1)
    pm = fz._part_models(rng, n_parts)
    pm = np.stack([(world_from_cam.reshape(4, 4).T @ m.reshape(4, 4).T).T.reshape(-1) for m in pm])
    mvp = orc.compose_mvp(proj, view, pm, n_parts)
    depth = rng.uniform(0.0, 9.0, (H, W)).astype(np.float32)
    want_d, want_m, want_z = orc.filter_frame(depth, tri, part, mvp, np.float32(0.1), np.float32(8.0), np.float32(0.05), np.float32(5.0), want_mask=True, want_zbuf=True, nthreads=8)
    la = np.asarray(orc.lookat()).reshape(4, 4).T
    cam = (np.linalg.inv(la) @ np.asarray(view).reshape(4, 4).T).T.reshape(-1)
    with tempfile.TemporaryDirectory() as td:
        gl_case.write_case_raw(td+"/c.bin", W, H, proj, np.eye(4).reshape(-1), cam, pm, tri, part, depth, 0.1, 8.0, 0.05, 5.0)
        r = subprocess.run([ROOT+"/oracle/_ref/gl_crosscheck_glx", td+"/c.bin", "-", td+"/o.bin"], capture_output=True, text=True, env=env)
        if r.returncode: print(seed, "GL failed", r.stderr[-300:]); continue
        d, m = gl_case.read_dump(td+"/o.bin")
    dm = m != want_m
    dd = (d.view(np.uint32) != want_d.view(np.uint32)) & ~dm
    print("seed %2d %4dx%-4d tris %5d  mask diff %5d (%.3f %%)  depth diff (mask equal) %5d  covered %.2f" % (seed, W, H, len(tri), dm.sum(), 100.0*dm.mean(), dd.sum(), float((want_z < np.float32(0.98)).mean())))
