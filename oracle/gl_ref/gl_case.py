"""Case files for the GL cross-check harness (oracle/gl_ref/gl_crosscheck.cpp) and the comparison of its dump with
the CPU oracle.  TEST INFRASTRUCTURE ONLY (same rules as oracle_py.py).

    python oracle/gl_ref/gl_case.py make  <scene: example|pr2_small|pr2|walls|small:<scene>> <frame> <case.bin>
    python oracle/gl_ref/gl_case.py compare <case.bin> <dump.bin>      # prints the differing-pixel report as JSON

The case keeps the GL matrix-stack operands SEPARATE (projection, inverse(camera_offset), camera transform, per part
link_to_fixed * link_offset and the glTranslatef / glScalef suffix), because the point of the harness is that GL
multiplies them in float inside the driver, which the oracle restates as one double product rounded once.
"""
from __future__ import annotations

import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def _quat_matrix(q):
    x, y, z, w = [float(v) for v in q]
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
                     [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


def _gl(R, t):
    """tf::Transform::getOpenGLMatrix: column-major 4x4."""
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = R, t
    return M.T.reshape(-1).copy()


def camera_operands(sc, k):
    """inverse(camera_offset) and the camera transform shifted by camera_tx/ty (src/urdf_filter.cpp:602-614)."""
    from realtime_urdf_filter_b200 import synth
    _, tx, ty = sc.proj()
    Ts = sc.link_poses(k)
    Tc = (np.eye(4) if sc.cam_link < 0 else Ts[sc.cam_link]) @ synth.make_T(sc.cam_R, sc.cam_xyz)
    Tinv = np.linalg.inv(Tc)                               # lookupTransform(cam_frame, fixed_frame)
    R, t = Tinv[:3, :3], Tinv[:3, 3].copy()
    t = t + R @ np.array([1.0, 0, 0]) * tx + R @ np.array([0, 1.0, 0]) * ty
    Ro = _quat_matrix(sc.offset_q)
    to = np.asarray(sc.offset_t, float)
    return _gl(Ro.T, -Ro.T @ to), _gl(R, t)


def write_case(path, sc, k, depth_f32, max_diff=None, replace_value=None):
    from realtime_urdf_filter_b200 import synth
    import oracle_py as orc
    proj, _, _ = sc.proj()
    off_inv, cam = camera_operands(sc, k)
    Ts = sc.link_poses(k)
    order = np.argsort(sc.tri_part, kind="stable")          # the harness draws part by part
    with open(path, "wb") as f:
        f.write(b"RUFGLC01")
        f.write(struct.pack("<4i", sc.width, sc.height, sc.n_parts, sc.n_tris))
        f.write(np.asarray(proj, np.float64).tobytes() + off_inv.tobytes() + cam.tobytes())
        f.write(struct.pack("<4f", synth.Z_NEAR, synth.Z_FAR, sc.max_diff if max_diff is None else max_diff,
                            sc.replace_value if replace_value is None else replace_value))
        for p in sc.parts:
            T = Ts[p.link]
            model = orc.link_model(synth.quat_from_matrix(T[:3, :3]), T[:3, 3], p.off_q, p.off_t, None)
            kind, s = 0, (0.0, 0.0, 0.0)
            if p.suffix is not None:
                S = np.asarray(p.suffix, float).reshape(4, 4).T     # column-major -> math layout
                if np.any(S[:3, 3] != 0):
                    kind, s = 1, tuple(S[:3, 3])                    # glTranslatef (cylinder, src/renderable.cpp:95)
                else:
                    kind, s = 2, (S[0, 0], S[1, 1], S[2, 2])        # glScalef (doubled cube :128, mesh :427)
            f.write(np.asarray(model, np.float64).tobytes() + struct.pack("<i3f", kind, *s))
        f.write(np.ascontiguousarray(sc.tri[order], np.float32).tobytes())
        f.write(np.ascontiguousarray(sc.tri_part[order], np.uint32).tobytes())
        f.write(np.ascontiguousarray(depth_f32, np.float32).tobytes())


def write_case_raw(path, W, H, proj, off_inv, cam, part_models, tri, tri_part, depth_f32, z_near, z_far, max_diff, replace_value):
    """The same file from raw operands (random soups: no Scene behind them); part_models: (P, 16) column-major doubles."""
    tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
    tri_part = np.ascontiguousarray(tri_part, np.uint32)
    order = np.argsort(tri_part, kind="stable")
    with open(path, "wb") as f:
        f.write(b"RUFGLC01")
        f.write(struct.pack("<4i", W, H, len(part_models), len(tri)))
        f.write(np.asarray(proj, np.float64).tobytes() + np.asarray(off_inv, np.float64).tobytes() + np.asarray(cam, np.float64).tobytes())
        f.write(struct.pack("<4f", z_near, z_far, max_diff, replace_value))
        for m in part_models:
            f.write(np.asarray(m, np.float64).reshape(-1).tobytes() + struct.pack("<i3f", 0, 0.0, 0.0, 0.0))
        f.write(tri[order].tobytes())
        f.write(tri_part[order].tobytes())
        f.write(np.ascontiguousarray(depth_f32, np.float32).tobytes())


def read_case_header(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"RUFGLC01"
        W, H, P, T = struct.unpack("<4i", f.read(16))
    want = 8 + 16 + 3 * 128 + 16 + P * (128 + 16) + T * 40 + W * H * 4
    return dict(width=W, height=H, parts=P, tris=T, bytes=os.path.getsize(path), expected_bytes=want)


def read_dump(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"RUFGLO01"
        W, H = struct.unpack("<2i", f.read(8))
        d = np.frombuffer(f.read(W * H * 4), np.float32).reshape(H, W)
        m = np.frombuffer(f.read(W * H), np.uint8).reshape(H, W)
    return d, m


def compare(sc, k, depth_f32, gl_depth, gl_mask, max_diff=None, replace_value=None):
    """Differing pixels between a real GL driver and the CPU oracle; `silhouette` = the oracle's virtual depth jumps by
    more than 1 cm inside the pixel's 3x3 neighbourhood (where fill-rule / sub-pixel-precision differences may show)."""
    import oracle_py as orc
    from realtime_urdf_filter_b200 import synth
    proj, _, _ = sc.proj()
    view, pm = sc.frame(k)
    mvp = orc.compose_mvp(proj, view, pm, sc.n_parts)
    md = np.float32(sc.max_diff if max_diff is None else max_diff)
    rv = np.float32(sc.replace_value if replace_value is None else replace_value)
    want_d, want_m, z = orc.filter_frame(np.ascontiguousarray(depth_f32, np.float32), sc.tri, sc.tri_part, mvp,
                                         np.float32(synth.Z_NEAR), np.float32(synth.Z_FAR), md, rv, want_mask=True,
                                         want_zbuf=True, nthreads=4)
    virt = synth.linear_depth(z)
    pad = np.pad(virt, 1, mode="edge")
    nb = np.stack([pad[1 + dy:1 + dy + virt.shape[0], 1 + dx:1 + dx + virt.shape[1]] for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
    silhouette = (nb.max(0) - nb.min(0)) > 0.01
    dm = gl_mask != want_m
    same = ~dm
    dd = np.zeros_like(dm)
    dd[same] = gl_depth[same].view(np.uint32) != want_d[same].view(np.uint32)
    return dict(pixels=int(dm.size), mask_diff=int(dm.sum()), mask_diff_on_silhouette=int((dm & silhouette).sum()),
                mask_diff_elsewhere=int((dm & ~silhouette).sum()), depth_diff_where_mask_agrees=int(dd.sum()),
                silhouette_pixels=int(silhouette.sum()))


def _scene(name):
    from realtime_urdf_filter_b200 import synth
    if name == "kinds":                                 # 160 x 120, every renderable kind: box + doubled cube (glScalef), sphere,
        import importlib.util                           # cylinder (glTranslatef), scaled mesh -- tests/golden/make_path_golden.py
        spec = importlib.util.spec_from_file_location("make_path_golden", os.path.join(ROOT, "tests", "golden", "make_path_golden.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.small_scene()
    if name.startswith("small:"):                      # small:<scene>: the same scene at 160 x 120 (golden fixtures)
        base = name.split(":", 1)[1]
        return {"example": lambda: synth.example_scene(160, 120), "walls": lambda: synth.walls_scene(160, 120),
                "pr2_small": lambda: synth.pr2_like_scene(160, 120, n_tris=6000, name="pr2_like_small_160")}[base]()
    return {"example": synth.example_scene, "walls": synth.walls_scene, "pr2": synth.pr2_like_scene, "multi": synth.multi_robot_scene,
            "pr2_small": lambda: synth.pr2_like_scene(n_tris=6000, name="pr2_like_small")}[name]()


def _frame_depth(sc, k):
    import oracle_py as orc
    from realtime_urdf_filter_b200 import synth
    proj, _, _ = sc.proj()
    view, pm = sc.frame(k)
    z = orc.render(sc.tri, sc.tri_part, orc.compose_mvp(proj, view, pm, sc.n_parts), sc.width, sc.height,
                   np.float32(8.0 * 0.99), nthreads=4)
    u16 = synth.synth_depth(synth.linear_depth(z), k, "u16")
    return orc.u16_to_f32(u16.reshape(-1)).reshape(u16.shape)       # what filter_callback hands to filter() (:287-288)


if __name__ == "__main__":
    if sys.argv[1] == "make":
        sc = _scene(sys.argv[2])
        k = int(sys.argv[3])
        write_case(sys.argv[4], sc, k, _frame_depth(sc, k))
        print(json.dumps(read_case_header(sys.argv[4])))
    else:
        hdr_scene, k = sys.argv[4], int(sys.argv[5])
        sc = _scene(hdr_scene)
        d, m = read_dump(sys.argv[3])
        print(json.dumps(compare(sc, k, _frame_depth(sc, k), d, m)))
