#!/usr/bin/env python
"""Golden vectors from the REFERENCE ITSELF: the reference's shader files (include/shaders/urdf_filter.vert|.frag, loaded
unmodified from /root/reference) driven through the GL call sequence of RealtimeURDFFilter::render in a real GL driver --
Mesa 18.1.9 llvmpipe, the libGL NVIDIA ships inside Nsight Compute, on top of oracle/gl_ref/fakex11 (no X server) -- by
oracle/gl_ref/gl_crosscheck.cpp.  The reference checkout does not travel to the GPU box; these vectors do.

    make -C oracle/gl_ref glx && python tests/golden/make_gl_golden.py        -> tests/golden/gl_llvmpipe.npz

Per case (scene at 160 x 120, frame k): attachment 1 (filtered depth, float bits) and attachment 3 (mask) as the driver
returned them; the input depth image is regenerated from (scene, k) by oracle/gl_ref/gl_case.py::_frame_depth.
Plus hostile random soups (tests/helpers.py::fuzz_case, regenerated from their seed): there the driver and the oracle differ
on a few mask pixels per image (float matrix stack, float clipper) -- their number is recorded with the vectors."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GLREF = os.path.join(ROOT, "oracle", "gl_ref")
sys.path.insert(0, GLREF); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gl_case  # noqa: E402
import helpers  # noqa: E402
import oracle_py as orc  # noqa: E402

FUZZ_SEEDS = [2, 4, 6, 9, 12]

CASES = [("small:example", 0), ("small:example", 7), ("small:pr2_small", 0), ("small:pr2_small", 7), ("small:pr2_small", 19),
         ("small:walls", 0), ("small:walls", 7), ("kinds", 0), ("kinds", 1), ("kinds", 9)]
SHADERS = os.environ.get("RUF_REFERENCE_SHADERS", "/root/reference/include/shaders")


def run_gl(sc, k, depth, raw=None):
    mesa = subprocess.run(["make", "-s", "-C", GLREF, "mesa_dir"], capture_output=True, text=True).stdout.strip()
    exe = os.path.join(ROOT, "oracle", "_ref", "gl_crosscheck_glx")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "oracle", "_ref", "fakex") + ":" + mesa)
    with tempfile.TemporaryDirectory() as td:
        case, dump = os.path.join(td, "case.bin"), os.path.join(td, "dump.bin")
        if raw is None:
            gl_case.write_case(case, sc, k, depth)
        else:
            la = np.asarray(orc.lookat()).reshape(4, 4).T
            cam = (np.linalg.inv(la) @ np.asarray(raw["view"]).reshape(4, 4).T).T.reshape(-1)          # MODELVIEW = LookAt * cam
            gl_case.write_case_raw(case, raw["W"], raw["H"], raw["proj"], np.eye(4).reshape(-1), cam, raw["pm"], raw["tri"], raw["part"],
                                   raw["depth"], raw["z_near"], raw["z_far"], raw["max_diff"], raw["replace_value"])
        r = subprocess.run([exe, case, SHADERS, dump], capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        d, m = gl_case.read_dump(dump)
        return d.copy(), m.copy(), r.stderr.strip().splitlines()[0]


if __name__ == "__main__":
    out, meta = {}, {"cases": [], "shaders": "include/shaders/urdf_filter.vert|.frag of the reference, unmodified"}
    for i, (name, k) in enumerate(CASES):
        sc = gl_case._scene(name)
        depth = gl_case._frame_depth(sc, k)
        d, m, gl_info = run_gl(sc, k, depth)
        rep = gl_case.compare(sc, k, depth, d, m)
        out[f"depth_{i}"], out[f"mask_{i}"] = d.view(np.uint32), m
        meta["cases"].append(dict(scene=name, frame=k, width=sc.width, height=sc.height, oracle_vs_gl=rep))
        meta["gl"] = gl_info
        print(name, k, rep)
    meta["fuzz"] = []
    for j, (seed, special) in enumerate([(sd, False) for sd in FUZZ_SEEDS] + [(6, True), (9, True)]):
        fc = helpers.fuzz_case(seed, special=special)
        d, m, _ = run_gl(None, 0, None, raw=fc)
        want_d, want_m = helpers.fuzz_oracle(fc)
        dm = m != want_m
        assert np.array_equal(d.view(np.uint32)[~dm], want_d.view(np.uint32)[~dm])
        out[f"fuzz_depth_{j}"], out[f"fuzz_mask_{j}"] = d.view(np.uint32), m
        meta["fuzz"].append(dict(seed=seed, special=special, width=fc["W"], height=fc["H"], triangles=int(len(fc["tri"])), mask_pixels_differing_from_oracle=int(dm.sum())))
        print("fuzz", seed, meta["fuzz"][-1])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gl_llvmpipe.npz"), meta=np.frombuffer(json.dumps(meta).encode(), np.uint8), **out)
    print(meta["gl"])
