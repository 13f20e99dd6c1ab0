"""Generates tests/golden/path_small.npz and tests/golden/path_hashes.json.

    python tests/golden/make_path_golden.py

The reference (blodow/realtime_urdf_filter) has no tests, golden vectors or fixtures for its render + filter
path and cannot be built or run in this image (OpenGL/GLEW/freeglut/ROS/Assimp are absent),
so these vectors are NOT reference outputs: they freeze the outputs of the CPU oracle (`oracle/ruf_oracle.c`)
after it was pinned by the known-answer, ray-caster and property tests.  Their job is regression: an edit
of the oracle, of the synthetic scene generators or of the CUDA path that changes a single bit of any frame shows
up against a committed file, not only against an oracle rebuilt from the same edit.

* path_small.npz  -- self-contained raw inputs AND outputs of a 160x120 scene (the two example.urdf walls with
  their doubled boxes, a sphere and a cylinder; 2 frames x 2 encodings): triangles, part indices, the float32
  MVP table, the sensor images, z-buffer, filtered depth and mask.  No generator is needed to replay it.
* path_hashes.json -- SHA-256 of inputs and outputs of frames of the full-size configurations (C1 example.urdf
  640x480, the 6,000-triangle and the 89,780-triangle PR2-like model, C3 walls 1280x960), which are too
  large to commit raw.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_py as orc  # noqa: E402
from realtime_urdf_filter_b200 import synth  # noqa: E402

HASH_CASES = [("example", 0), ("example", 11), ("pr2_small", 3), ("pr2", 0), ("pr2", 17), ("walls", 2)]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_scene():
    """example.urdf's two walls plus a swinging sphere, a sliding cylinder and a scaled blob mesh, 160x120:
    every renderable kind of the reference (box + doubled cube, sphere, cylinder, mesh) in one small frame."""
    base = synth.example_scene(width=160, height=120)
    links = list(base.links) + [
        synth.Link("ball", -1, (-0.4, 2.0, 0.2), (0, 0, 0), (0, 0, 1), "revolute", 0.3, 0.5, 0.7, 0.1),
        synth.Link("rod", 2, (0.9, 0.0, -0.3), (0.4, 0.2, 0.0), (1, 0, 0), "prismatic", 0.0, 0.2, 1.1, 0.5),
        synth.Link("blob", -1, (0.1, 1.2, -0.25), (0.1, 0.3, 0.9)),
    ]
    parts, tris, pidx = [], [], []
    synth.add_box(parts, tris, pidx, 0, (4, 0.5, 2))
    synth.add_box(parts, tris, pidx, 1, (4, 0.5, 2))
    synth.add_sphere(parts, tris, pidx, 2, 0.35, off_t=(0.05, 0.0, 0.0))
    synth.add_cylinder(parts, tris, pidx, 3, 0.12, 0.8)
    synth.add_mesh(parts, tris, pidx, 4, synth.blob_mesh(np.random.default_rng(5), (0.2, 0.15, 0.1), 12, 8),
                   scale=(1.5, 1.0, 0.75))
    tri, tp = synth._finish(tris, pidx)
    return synth.Scene("golden_small", 160, 120, synth.kinect_P(160, 120), links, parts, tri, tp, -1, (0, 0, 0),
                       base.cam_R, label="walls + sphere + cylinder + mesh, 160x120")


def run_frame(sc, k, enc, nthreads=4):
    proj, _, _ = sc.proj()
    view, pm = sc.frame(k)
    mvp = orc.compose_mvp(proj, view, pm, sc.n_parts)
    z = orc.render(sc.tri, sc.tri_part, mvp, sc.width, sc.height, np.float32(synth.Z_FAR * 0.99), nthreads=nthreads)
    depth = synth.synth_depth(synth.linear_depth(z), k, enc)
    out, mask, zbuf = orc.filter_frame(depth, sc.tri, sc.tri_part, mvp, np.float32(synth.Z_NEAR),
                                       np.float32(synth.Z_FAR), np.float32(sc.max_diff),
                                       np.float32(sc.replace_value), want_mask=True, nthreads=nthreads,
                                       want_zbuf=True)
    return dict(proj=proj, view=view, pm=pm, mvp=mvp, depth=depth, out=out, mask=mask, zbuf=zbuf)


def hash_case(name, k):
    import helpers
    sc = helpers.scene(name)
    rec = {"scene": name, "frame": k, "width": sc.width, "height": sc.height, "n_tris": int(sc.n_tris),
           "n_parts": int(sc.n_parts), "tri": sha(sc.tri), "tri_part": sha(sc.tri_part)}
    for enc in ("u16", "f32"):
        r = run_frame(sc, k, enc)
        rec[enc] = {"mvp": sha(r["mvp"]), "depth_in": sha(r["depth"]), "zbuf": sha(r["zbuf"]),
                    "depth_out": sha(r["out"]), "mask": sha(r["mask"]),
                    "masked_px": int(np.count_nonzero(r["mask"]))}
    return rec


def main():
    sc = small_scene()
    arrays = {"tri": sc.tri, "tri_part": sc.tri_part, "n_parts": np.int32(sc.n_parts),
              "width": np.int32(sc.width), "height": np.int32(sc.height),
              "max_diff": np.float32(sc.max_diff), "replace_value": np.float32(sc.replace_value),
              "z_near": np.float32(synth.Z_NEAR), "z_far": np.float32(synth.Z_FAR)}
    for k in (0, 5):
        for enc in ("u16", "f32"):
            r = run_frame(sc, k, enc)
            for key, val in r.items():
                arrays[f"f{k}_{enc}_{key}"] = val
    np.savez_compressed(os.path.join(HERE, "path_small.npz"), **arrays)
    hashes = {"_made_by": "tests/golden/make_path_golden.py (CPU oracle outputs; not reference outputs)",
              "cases": [hash_case(n, k) for n, k in HASH_CASES]}
    with open(os.path.join(HERE, "path_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    print("wrote", os.path.getsize(os.path.join(HERE, "path_small.npz")), "bytes npz;", len(hashes["cases"]), "hash cases")


if __name__ == "__main__":
    main()
