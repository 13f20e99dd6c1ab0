"""Host facade pieces that need no GPU: URDF subset parser, renderable construction (box F4 doubles,
cylinder/sphere tessellation + suffix matrices, scale, ignore list, geometry_type), STL loader."""
import os
import struct

import numpy as np
import pytest

import oracle_py as orc
from realtime_urdf_filter_b200 import facade, synth


def test_example_urdf_parts_match_scene():
    tri, part, pm = facade.parse_urdf(synth.example_urdf_xml(), "visual")
    sc = synth.example_scene()
    assert np.array_equal(tri, sc.tri) and np.array_equal(part, sc.tri_part)
    assert pm.shape == (4, 16)
    # identity TF: part 0 = identity, part 1 = glScalef(4, 0.5, 2)
    assert np.array_equal(pm[0], np.eye(4).reshape(-1))
    assert np.array_equal(pm[1], synth.scale_suffix(4, 0.5, 2))


def test_geometry_type_and_ignore_and_scale():
    xml = synth.example_urdf_xml()
    t_col, _, _ = facade.parse_urdf(xml, "collision")
    assert len(t_col) == 48
    t_ign, _, pm = facade.parse_urdf(xml, "", ignore=["wall2"])
    assert len(t_ign) == 24 and pm.shape[0] == 2
    t2, _, pm2 = facade.parse_urdf(xml, "visual", scale=0.5)
    assert np.array_equal(t2[:12], orc.box_triangles(2, 0.25, 1))
    assert np.array_equal(t2[12:24], orc.cube_triangles(2))
    assert np.array_equal(pm2[1], synth.scale_suffix(2, 0.25, 1))
    t_bad, _, _ = facade.parse_urdf(xml, "nonsense")        # ROS_FATAL + nothing rendered
    assert len(t_bad) == 0


def test_primitives_origin_and_suffix():
    xml = """<?xml version="1.0"?>
    <!-- comment -->
    <robot name="p">
      <link name="a"><visual><origin xyz="0.1 0.2 0.3" rpy="0.1 -0.2 0.3"/><geometry><cylinder radius="0.2" length="1.0"/></geometry></visual>
                     <visual><geometry><sphere radius="0.25"/></geometry></visual></link>
      <link name="b"/>
      <joint name="ab" type="fixed"><parent link="a"/><child link="b"/></joint>
    </robot>"""
    tri, part, pm = facade.parse_urdf(xml, "visual")
    assert len(tri) == 220 + 180 and pm.shape == (2, 16)
    assert np.array_equal(tri[:220], orc.cylinder_triangles(0.2, 1.0))
    assert np.array_equal(tri[220:], orc.sphere_triangles(0.25))
    # cylinder: link_offset * glTranslatef(0,0,-length/2)
    r, p, y = 0.1, -0.2, 0.3
    R = synth.rpy_matrix(r, p, y)
    M = pm[0].reshape(4, 4).T
    assert np.allclose(M[:3, :3], R, atol=1e-12)
    assert np.allclose(M[:3, 3], np.array([0.1, 0.2, 0.3]) + R @ np.array([0, 0, -0.5]), atol=1e-12)
    # bit-exact against the oracle's link_model with the same quaternion
    q = synth.quat_from_matrix(R)
    want = orc.link_model((0, 0, 0, 1), (0, 0, 0), q, (0.1, 0.2, 0.3), synth.translate_suffix(0, 0, -0.5))
    assert np.allclose(pm[0], want, atol=1e-15)


def test_malformed_urdf_is_rejected():
    for bad in ["", "<robot", "<notrobot/>", "<robot><link></link></robot>",
                '<robot><link name="a"><visual><geometry><cone/></geometry></visual></link></robot>']:
        with pytest.raises(ValueError):
            facade.parse_urdf(bad)


def test_urdfdom_validation_rules():
    """What urdf::Model::initString (urdfdom parseURDF / initTree / initRoot) rejects, the subset parser rejects
    with the same reason; src/urdf_renderer.cpp:69-75 then logs and builds no renderables."""
    link = lambda n: f'<link name="{n}"/>'
    joint = lambda n, p, c, t="fixed", extra="": (f'<joint name="{n}" type="{t}"><parent link="{p}"/>'
                                                  f'<child link="{c}"/>{extra}</joint>')
    cases = {
        "No name given for the robot": f'<robot>{link("a")}</robot>',
        "No link elements found": '<robot name="r"/>',
        "link 'a' is not unique": f'<robot name="r">{link("a")}{link("a")}</robot>',
        "joint 'j' is not unique": f'<robot name="r">{link("a")}{link("b")}{link("c")}{joint("j", "a", "b")}{joint("j", "a", "c")}</robot>',
        "has no known type [hinge]": f'<robot name="r">{link("a")}{link("b")}{joint("j", "a", "b", "hinge")}</robot>',
        "does not specify limits": f'<robot name="r">{link("a")}{link("b")}{joint("j", "a", "b", "revolute")}</robot>',
        "child link [zz] of joint [j] not found": f'<robot name="r">{link("a")}{joint("j", "a", "zz")}</robot>',
        "parent link [zz] of joint [j] not found": f'<robot name="r">{link("a")}{joint("j", "zz", "a")}</robot>',
        "Two root links found: [a] and [c]": f'<robot name="r">{link("a")}{link("b")}{link("c")}{joint("j", "a", "b")}</robot>',
        "No root link found": f'<robot name="r">{link("a")}{link("b")}{joint("j", "a", "b")}{joint("k", "b", "a")}</robot>',
        "missing a parent and/or child": f'<robot name="r">{link("a")}{link("b")}<joint name="j" type="fixed"><parent link="a"/></joint></robot>',
    }
    for reason, xml in cases.items():
        with pytest.raises(ValueError, match=reason.replace("[", r"\[").replace("]", r"\]")):
            facade.parse_urdf(xml)
    # accepted: continuous needs no limit, revolute with one, entities in attribute values
    ok = (f'<robot name="r &amp; d">{link("a")}{link("b")}{link("c")}{joint("j", "a", "b", "continuous")}'
          f'{joint("k", "b", "c", "revolute", "<limit lower=\"0\" upper=\"1\" effort=\"1\" velocity=\"1\"/>")}</robot>')
    tri, _, pm = facade.parse_urdf(ok)
    assert len(tri) == 0 and pm.shape[0] == 0


def test_renderables_follow_getLinks_order_not_document_order():
    """urdf::ModelInterface::getLinks iterates a std::map keyed by link name (src/urdf_renderer.cpp:87-95): the
    renderables come out sorted by link name whatever the document order is."""
    def vis(size):
        return f'<visual><geometry><sphere radius="{size}"/></geometry></visual>'
    xml = (f'<robot name="o"><link name="zeta">{vis(0.3)}</link><link name="Alpha">{vis(0.1)}</link>'
           f'<link name="beta">{vis(0.2)}</link>'
           '<joint name="j1" type="fixed"><parent link="zeta"/><child link="Alpha"/></joint>'
           '<joint name="j2" type="fixed"><parent link="zeta"/><child link="beta"/></joint></robot>')
    tri, part, pm = facade.parse_urdf(xml, "visual")
    assert pm.shape[0] == 3 and len(tri) == 3 * 180
    radii = [float(np.abs(tri[part == k]).max()) for k in range(3)]
    assert np.allclose(radii, [0.1, 0.2, 0.3])          # "Alpha" < "beta" < "zeta" (byte-wise: upper case first)


def _write_binary_stl(path, tris, header=b"binary"):
    with open(path, "wb") as f:
        f.write(header.ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for t in tris:
            f.write(struct.pack("<12fH", 0, 0, 1, *t, 0))


def test_stl_loader_binary_ascii_and_solid_prefixed_binary(tmp_path):
    tris = np.random.default_rng(3).normal(size=(7, 9)).astype(np.float32)
    os.makedirs(tmp_path / "pkg" / "meshes")
    _write_binary_stl(tmp_path / "pkg" / "meshes" / "a.stl", tris)
    _write_binary_stl(tmp_path / "pkg" / "meshes" / "solid.stl", tris, header=b"solid but binary")   # README.md:121-141
    with open(tmp_path / "pkg" / "meshes" / "ascii.stl", "w") as f:
        f.write("solid x\n")
        for t in tris:
            f.write("facet normal 0 0 1\n outer loop\n")
            for k in range(3):
                f.write("  vertex %r %r %r\n" % tuple(float(v) for v in t[3 * k:3 * k + 3]))
            f.write(" endloop\nendfacet\n")
        f.write("endsolid x\n")
    for name in ("a.stl", "solid.stl", "ascii.stl"):
        xml = f'''<robot name="m"><link name="l"><visual><geometry>
                  <mesh filename="package://pkg/meshes/{name}" scale="0.001 0.002 0.003"/></geometry></visual></link></robot>'''
        tri, part, pm = facade.parse_urdf(xml, "", resource_root=str(tmp_path))
        assert np.array_equal(tri, tris), name
        assert np.array_equal(pm[0], synth.scale_suffix(np.float32(0.001), np.float32(0.002), np.float32(0.003)))
    # a missing mesh logs an error and yields an empty renderable, it does not abort the model
    xml = '<robot name="m"><link name="l"><visual><geometry><mesh filename="package://pkg/nope.stl"/></geometry></visual></link></robot>'
    tri, _, pm = facade.parse_urdf(xml, "", resource_root=str(tmp_path))
    assert len(tri) == 0 and pm.shape[0] == 1


def test_constructor_reads_params_and_logs_fatal_on_missing():
    with facade.FilterNode({"fixed_frame": "/world", "camera_frame": "/cam", "depth_distance_threshold": 0.05,
                            "filter_replace_value": 5.0}, []) as n:
        assert n.get("depth_distance_threshold") == 0.05 and n.get("filter_replace_value") == 5.0
        assert n.get("far_plane") == 8.0 and n.get("near_plane") == 0.1
        assert "FATAL" not in n.log()
        g = n.projection(640, 480, [525, 0, 319.5, -39.375, 0, 525, 239.5, 0, 0, 0, 1, 0])
        assert np.array_equal(g, orc.projection_matrix([525, 0, 319.5, -39.375, 0, 525, 239.5, 0, 0, 0, 1, 0], 640, 480)[0])
        assert n.get("camera_tx") == 0.075
    with facade.FilterNode({}, []) as n:          # missing required params: FATAL logs, construction continues
        log = n.log()
        assert log.count("FATAL") == 3 and n.get("filter_replace_value") == 0.0


def test_tracker_caller_conversions():
    """The second caller of filter() (src/urdf_filtered_tracker.cpp:201-249): mirrored mm -> m in double, the
    hard-coded intrinsics held as float, and the truncating, un-mirrored m -> mm of the result."""
    rng = np.random.default_rng(5)
    mm = rng.integers(0, 65536, (7, 12)).astype(np.uint16)
    mm[0, :4] = [0, 1, 65535, 739]
    buf = facade.tracker_depth_to_buffer(mm)
    want = (mm[:, ::-1].astype(np.float64) * 0.001).astype(np.float32)
    assert np.array_equal(buf.view(np.uint32), want.view(np.uint32))
    # not the same thing as filter_callback's float product for every value (cv::Mat::convertTo, :288)
    all_mm = np.arange(65536, dtype=np.uint16).reshape(1, -1)
    d = facade.tracker_depth_to_buffer(all_mm)[0, ::-1]
    f = orc.u16_to_f32(all_mm.reshape(-1))
    assert 0 < np.count_nonzero(d != f) < 65536
    # projection: the KAT set of SURVEY.md 8(d), computed from float-held intrinsics
    g = facade.tracker_projection(640, 480)
    P = [float(np.float32(585.260)), 0, float(np.float32(317.387)), 0, 0, float(np.float32(585.028)),
         float(np.float32(239.264)), 0, 0, 0, 1, 0]
    assert np.array_equal(g, orc.projection_matrix(P, 640, 480)[0])
    # output: truncation, no mirroring.  Through the tracker's own input conversion an untouched pixel loses
    # 1 mm for 739 of the 65536 values; filter_callback's float product would round-trip all
    back = facade.tracker_masked_depth_to_mm(d.reshape(1, -1))[0]
    assert np.array_equal(back, (d * np.float32(1000)).astype(np.uint16))
    assert np.count_nonzero(back != np.arange(65536)) == 739
    assert np.array_equal(facade.tracker_masked_depth_to_mm(f.reshape(1, -1))[0], np.arange(65536))
    edge = np.array([[np.nan, -1.0, 1e9, 65.535, 5.0]], np.float32)
    assert facade.tracker_masked_depth_to_mm(edge).tolist() == [[0, 0, 65535, 65535, 5000]]


def test_urdf_parser_survives_mutated_and_hostile_input():
    """The hand-written XML subset reader must reject, never crash: seeded mutations of the example URDF,
    truncations, and nesting far beyond what a robot description has."""
    import random
    base = synth.example_urdf_xml()
    rnd = random.Random(11)
    chars = '<>/="\' \n&;!-?abz019.'
    outcomes = {"ok": 0, "rejected": 0}
    for _ in range(400):
        s = list(base if rnd.random() < 0.7 else base[:rnd.randrange(len(base))])
        for _ in range(rnd.randrange(1, 6)):
            i = rnd.randrange(len(s)) if s else 0
            op = rnd.randrange(3)
            if op == 0 and s:
                del s[i:i + rnd.randrange(1, 20)]
            elif op == 1:
                s[i:i] = [rnd.choice(chars) for _ in range(rnd.randrange(1, 5))]
            elif s:
                s[i] = rnd.choice(chars)
        try:
            facade.parse_urdf("".join(s))
            outcomes["ok"] += 1
        except ValueError:
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 100 and outcomes["ok"] > 0
    deep = "<robot name='d'>" + "<a>" * 100000 + "</a>" * 100000 + "</robot>"
    with pytest.raises(ValueError, match="nested deeper"):
        facade.parse_urdf(deep)


def test_binary_stl_with_trailing_bytes_and_visible_mesh_failures(tmp_path):
    """exporters append bytes after the declared facets (accepted, ignored); a mesh that cannot be loaded is
    recorded per renderer (`mesh_errors`) and, with the `strict_meshes` parameter, stops the filter like a model-less one."""
    tris = np.random.default_rng(4).normal(size=(5, 9)).astype(np.float32)
    os.makedirs(tmp_path / "pkg" / "meshes")
    _write_binary_stl(tmp_path / "pkg" / "meshes" / "padded.stl", tris)
    with open(tmp_path / "pkg" / "meshes" / "padded.stl", "ab") as f:
        f.write(b"\x00" * 37 + b"COLOR=\n")
    xml = '<robot name="m"><link name="l"><visual><geometry><mesh filename="package://pkg/meshes/padded.stl"/></geometry></visual></link></robot>'
    tri, _, _ = facade.parse_urdf(xml, "", resource_root=str(tmp_path))
    assert np.array_equal(tri, tris)
    # an ASCII file that is long enough to hold the "declared" facet count of its own text is still read as ASCII
    with open(tmp_path / "pkg" / "meshes" / "ascii_long.stl", "w") as f:
        f.write("solid " + "x" * 200 + "\n")
        for t in tris:
            f.write("facet normal 0 0 1\n outer loop\n")
            for k in range(3):
                f.write("  vertex %r %r %r\n" % tuple(float(v) for v in t[3 * k:3 * k + 3]))
            f.write(" endloop\nendfacet\n")
        f.write("endsolid x\n")
    tri, _, _ = facade.parse_urdf(xml.replace("padded", "ascii_long"), "", resource_root=str(tmp_path))
    assert np.array_equal(tri, tris)
