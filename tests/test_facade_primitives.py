"""Rows a9 / a11 of SURVEY.md 8(a): a URDF whose links carry a <sphere>, a <cylinder> (the -len/2 translate suffix of
src/renderable.cpp:92-98), a <box> with a rotated visual origin and an STL <mesh scale=...> (glScalef suffix,
src/renderable.cpp:424-452) goes through the C++ RealtimeURDFFilter facade -- URDF text, TF lookups, Image + CameraInfo
-- to the kernels and is compared with the oracle bit for bit."""
import math
import struct

import numpy as np
import pytest

import helpers
import oracle_py as orc
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import facade, synth

W, H = 640, 480
SCALE = 0.9          # the model's `scale` rosparam (src/urdf_filter.cpp:166): multiplies every dimension


def _blob_stl(path):
    tris = synth.blob_mesh(np.random.default_rng(21), (0.25, 0.18, 0.3), 12, 9)
    with open(path, "wb") as f:
        f.write(b"blob".ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for t in tris:
            f.write(struct.pack("<12fH", 0, 0, 1, *[float(v) for v in t], 0))
    return np.asarray(tris, np.float32)


def _rpy_quat(r, p, y):
    """urdf::Rotation::setFromRPY, the formula host/urdf_model.cpp restates."""
    phi, the, psi = r / 2.0, p / 2.0, y / 2.0
    return (math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi),
            math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi),
            math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi),
            math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi))


# link name -> (geometry xml, visual origin xyz, rpy); names sort in this order (urdf::Model::getLinks is a std::map)
LINKS = {
    "a_box": ('<box size="0.5 0.4 0.3"/>', (0.1, 0.0, 0.05), (0.0, 0.0, 0.3)),
    "b_ball": ('<sphere radius="0.25"/>', (0.0, -0.05, 0.0), (0.0, 0.0, 0.0)),
    "c_rod": ('<cylinder radius="0.1" length="0.8"/>', (0.0, 0.0, 0.1), (0.5, 0.2, 0.0)),
    "d_blob": ('<mesh filename="package://pkg/meshes/blob.stl" scale="0.5 1.5 1.0"/>', (0.02, 0.03, -0.04), (0.1, -0.2, 0.7)),
}
# where the links stand in the fixed frame (camera at the origin looking along +y of the world, like example.urdf)
POSES = {
    "a_box": ((0.0, 0.0, math.sin(0.2), math.cos(0.2)), (-0.9, 3.0, 0.3)),
    "b_ball": ((0.0, 0.0, 0.0, 1.0), (0.3, 1.6, -0.2)),
    "c_rod": ((math.sin(0.35), 0.0, 0.0, math.cos(0.35)), (0.9, 2.2, 0.4)),
    "d_blob": ((0.0, math.sin(-0.4), 0.0, math.cos(-0.4)), (-0.2, 1.1, 0.35)),
}


def _xml():
    links = "".join(f'<link name="{n}"><visual><origin xyz="{o[0]} {o[1]} {o[2]}" rpy="{r[0]} {r[1]} {r[2]}"/>'
                    f'<geometry>{g}</geometry></visual></link>' for n, (g, o, r) in LINKS.items())
    names = list(LINKS)
    joints = "".join(f'<joint name="j{i}" type="fixed"><parent link="{names[0]}"/><child link="{n}"/></joint>'
                     for i, n in enumerate(names[1:]))
    return f'<robot name="prims">{links}{joints}</robot>'


def _expected_parts(identity_tf):
    """(off_q, off_t, suffix, link) of every part, in the facade's order: box -> VBO box + the doubled cube,
    sphere, cylinder with its translate suffix, mesh with its scale suffix; float members like the reference's."""
    f32 = lambda v: float(np.float32(v))
    out = []
    for name, (geom, o, r) in LINKS.items():
        q = _rpy_quat(*r)
        if "box" in geom:
            dx, dy, dz = f32(SCALE * 0.5), f32(SCALE * 0.4), f32(SCALE * 0.3)
            out += [(q, o, None, name), (q, o, synth.scale_suffix(dx, dy, dz), name)]
        elif "sphere" in geom:
            out.append((q, o, None, name))
        elif "cylinder" in geom:
            out.append((q, o, synth.translate_suffix(0.0, 0.0, float(-np.float32(SCALE * 0.8) / 2)), name))
        else:
            out.append((q, o, synth.scale_suffix(f32(SCALE * 0.5), f32(SCALE * 1.5), f32(SCALE * 1.0)), name))
    pm = np.zeros((len(out), 16))
    for i, (q, o, sfx, name) in enumerate(out):
        lq, lt = ((0, 0, 0, 1), (0, 0, 0)) if identity_tf else POSES[name]
        pm[i] = ruf.part_model(lq, lt, q, o, sfx)
    return pm


def _parse(tmp_path):
    (tmp_path / "pkg" / "meshes").mkdir(parents=True, exist_ok=True)
    blob = _blob_stl(tmp_path / "pkg" / "meshes" / "blob.stl")
    tri, part, pm_id = facade.parse_urdf(_xml(), "visual", scale=SCALE, resource_root=str(tmp_path))
    return blob, tri, part, pm_id


def test_primitive_and_mesh_parts_as_the_reference_builds_them(tmp_path):
    """CPU: triangles, part order and part matrices (identity TF) of the URDF reader against the generators and the
    matrix chain of src/renderable.cpp:59-68,80-98,107-131,424-452."""
    blob, tri, part, pm_id = _parse(tmp_path)
    assert pm_id.shape[0] == 5 and part.max() == 4
    f32 = lambda v: float(np.float32(v))
    assert np.array_equal(tri[part == 0], ruf.box_triangles(f32(SCALE * 0.5), f32(SCALE * 0.4), f32(SCALE * 0.3)))
    assert np.array_equal(tri[part == 1], ruf.cube_triangles(f32(SCALE * 0.5)))
    assert np.array_equal(tri[part == 2], ruf.sphere_triangles(f32(SCALE * 0.25), 10, 10)) and (part == 2).sum() == 180
    assert np.array_equal(tri[part == 3], ruf.cylinder_triangles(f32(SCALE * 0.1), f32(SCALE * 0.8), 10, 10)) and (part == 3).sum() == 220
    assert np.array_equal(tri[part == 4], blob)                          # as stored in the file: the scale is a matrix suffix
    assert np.array_equal(pm_id.view(np.uint64), _expected_parts(True).view(np.uint64))
    # the cylinder spans z in [0, len] in its own frame; the suffix centres it (src/renderable.cpp:95)
    z = tri[part == 3].reshape(-1, 3)[:, 2]
    assert z.min() == 0.0 and np.isclose(z.max(), SCALE * 0.8)


@pytest.mark.gpu
@pytest.mark.parametrize("enc", ["u16", "f32"])
def test_sphere_cylinder_box_and_scaled_mesh_links_through_the_facade(tmp_path, enc):
    blob, tri, part, pm_id = _parse(tmp_path)
    pm = _expected_parts(False)
    P = synth.kinect_P(W, H)
    proj = orc.projection_matrix(P, W, H)[0]
    cam_R = synth.example_scene().cam_R                      # optical axis along world +y
    Tc = synth.make_T(cam_R, (0.0, 0.0, 0.0))
    Tinv = np.linalg.inv(Tc)
    view = ruf.view_matrix((0, 0, 0, 1), (0, 0, 0), synth.quat_from_matrix(Tinv[:3, :3]), Tinv[:3, 3], 0.0, 0.0)
    mvp = orc.compose_mvp(proj, view, pm, pm.shape[0])
    z = orc.render(tri, part, mvp, W, H, helpers.BG_Z, nthreads=8)
    depth = synth.synth_depth(synth.linear_depth(z), 5, enc)
    want_d, want_m, _ = orc.filter_frame(depth, tri, part, mvp, np.float32(0.1), np.float32(8.0), np.float32(0.05),
                                         np.float32(5.0), want_mask=True, nthreads=8)
    covered = synth.linear_depth(z) < 7.0
    assert covered.mean() > 0.05                              # the four links are in view
    params = {"fixed_frame": "/world", "camera_frame": "/cam", "depth_distance_threshold": 0.05,
              "filter_replace_value": 5.0, "show_gui": False, "robot_description": _xml()}
    models = [{"model": "robot_description", "tf_prefix": "/R", "geometry_type": "visual", "scale": SCALE}]
    with facade.FilterNode(params, models, camera_offset=((0, 0, 0), (0, 0, 0, 1))) as n:
        n.add_resource_root(str(tmp_path))
        n.set_tf("/world", (0, 0, 0, 1), (0, 0, 0))
        for name, (q, t) in POSES.items():
            n.set_tf("/R/" + name, q, t)
        n.set_tf("/cam", synth.quat_from_matrix(Tc[:3, :3]), Tc[:3, 3])
        n.callback(depth, P, stamp=2.0)
        c = n.counts()
        assert c["renderables"] == 4 and c["parts"] == 5 and c["triangles"] == len(tri)
        assert "Could not load" not in n.log()
        got_d, _ = n.last_image(0, depth.shape, depth.dtype)
        got_m, _ = n.last_image(1, depth.shape, np.uint8)
    assert np.array_equal(got_m, want_m), f"mask differs at {np.count_nonzero(got_m != want_m)} px"
    assert np.array_equal(got_d.view(np.uint8), want_d.view(np.uint8))
    # every renderable kind contributes pixels of its own: drop one part from the oracle's model and the mask changes
    for p in range(5):
        keep = part != p
        _, m2, _ = orc.filter_frame(depth, tri[keep], part[keep], mvp, np.float32(0.1), np.float32(8.0), np.float32(0.05),
                                    np.float32(5.0), want_mask=True, nthreads=8)
        # (part 1 is the box's doubled cube: glScalef(dx,dy,dz); glutSolidCube(dx) = dx^2 x dx*dy x dx*dz, inside the box for dx < 1)
        assert p == 1 or not np.array_equal(m2, want_m), f"part {p} is invisible in this test scene"
