"""Synthetic scenes for the parity tests and the benchmark.

Everything here is *input generation*: seeded procedural geometry, joint sweeps, camera
intrinsics and depth frames.  The matrices handed to the filter are built with the library's own
host-side functions (ruf_projection_matrix / ruf_view_matrix / ruf_part_model), i.e. exactly what
a ROS host would compute from CameraInfo and TF.

No PR2 assets exist on disk (and there is no network), so "PR2" means a PR2-*like* articulated
model: same kinematic layout (base, 4 casters x 2 wheels, torso lift, pan/tilt head, two 7-DoF arms
with grippers), ~88 drawn parts and ~90k triangles.  Results obtained with it are labelled
"PR2-like synthetic".
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib

# Host-math provider: the matrices and primitive tessellations come from libruf_b200.so's host functions by default.
# bench.py's reference arm swaps in the CPU oracle's bit-identical twins (tests/test_host_math.py) with use_math(), so
# that the CPU arm never loads the product library.
_math = _lib


def use_math(provider):
    """provider: object with projection_matrix, view_matrix, part_model, box/cube/sphere/cylinder_triangles."""
    global _math
    _math = provider

Z_NEAR, Z_FAR = 0.1, 8.0


# ------------------------------------------------------------------------------------------------
# small rigid-body helpers (float64, numpy)
# ------------------------------------------------------------------------------------------------
def rpy_matrix(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def axis_angle_matrix(axis, a):
    x, y, z = np.asarray(axis, float) / np.linalg.norm(axis)
    c, s, C = math.cos(a), math.sin(a), 1 - math.cos(a)
    return np.array([[c + x * x * C, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, c + z * z * C]])


def quat_from_matrix(R):
    """(x, y, z, w) of a rotation matrix."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s]
    q = np.array(q)
    return q / np.linalg.norm(q)


def make_T(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def scale_suffix(sx, sy, sz):
    """glScalef(sx, sy, sz) as a column-major double[16]."""
    m = np.zeros(16)
    m[0], m[5], m[10], m[15] = sx, sy, sz, 1.0
    return m


def translate_suffix(x, y, z):
    """glTranslatef(x, y, z) as a column-major double[16] (arguments are GLfloat)."""
    m = np.zeros(16)
    m[0] = m[5] = m[10] = m[15] = 1.0
    m[12], m[13], m[14] = np.float32(x), np.float32(y), np.float32(z)
    return m


# ------------------------------------------------------------------------------------------------
# scene description
# ------------------------------------------------------------------------------------------------
@dataclass
class Link:
    name: str
    parent: int                 # -1 = fixed frame
    xyz: tuple = (0.0, 0.0, 0.0)
    rpy: tuple = (0.0, 0.0, 0.0)
    axis: tuple = (0.0, 0.0, 1.0)
    jtype: str = "fixed"        # fixed | revolute | prismatic
    q0: float = 0.0
    amp: float = 0.0
    freq: float = 0.0
    phase: float = 0.0


@dataclass
class Part:
    """One draw call of the reference (one model matrix)."""
    link: int
    off_q: tuple = (0.0, 0.0, 0.0, 1.0)
    off_t: tuple = (0.0, 0.0, 0.0)
    suffix: np.ndarray | None = None


@dataclass
class Scene:
    name: str
    width: int
    height: int
    P: np.ndarray                       # CameraInfo.P (12,)
    links: list
    parts: list
    tri: np.ndarray                     # (T, 9) float32
    tri_part: np.ndarray                # (T,) uint32
    cam_link: int = -1                  # link the camera rides on (-1: fixed frame)
    cam_xyz: tuple = (0.0, 0.0, 0.0)
    cam_R: np.ndarray = field(default_factory=lambda: np.eye(3))   # optical frame in the link frame
    offset_q: tuple = (0.0, 0.0, 0.0, 1.0)
    offset_t: tuple = (0.0, 0.0, 0.0)
    max_diff: float = 0.05
    replace_value: float = 5.0
    fps: float = 30.0
    label: str = ""

    @property
    def n_parts(self):
        return len(self.parts)

    @property
    def n_tris(self):
        return int(self.tri.shape[0])

    def proj(self):
        return _math.projection_matrix(self.P, self.width, self.height, Z_NEAR, Z_FAR)

    def link_poses(self, k: int):
        """Forward kinematics at frame k -> list of 4x4 link_to_fixed."""
        t = k / self.fps
        Ts = []
        for ln in self.links:
            Tp = np.eye(4) if ln.parent < 0 else Ts[ln.parent]
            T = Tp @ make_T(rpy_matrix(*ln.rpy), ln.xyz)
            q = ln.q0 + ln.amp * math.sin(2 * math.pi * ln.freq * t + ln.phase)
            if ln.jtype == "revolute":
                T = T @ make_T(axis_angle_matrix(ln.axis, q), (0, 0, 0))
            elif ln.jtype == "prismatic":
                T = T @ make_T(np.eye(3), np.asarray(ln.axis, float) * q)
            Ts.append(T)
        return Ts

    def frame(self, k: int):
        """-> (view[16], part_models[n_parts, 16]) for frame k, built like the reference's host."""
        Ts = self.link_poses(k)
        _, tx, ty = self.proj()
        Tc = (np.eye(4) if self.cam_link < 0 else Ts[self.cam_link]) @ make_T(self.cam_R, self.cam_xyz)
        Tinv = np.linalg.inv(Tc)              # lookupTransform(cam_frame, fixed_frame)
        view = _math.view_matrix(self.offset_q, self.offset_t, quat_from_matrix(Tinv[:3, :3]), Tinv[:3, 3], tx, ty)
        pm = np.zeros((self.n_parts, 16))
        cache = {}
        for i, p in enumerate(self.parts):
            if p.link not in cache:
                T = Ts[p.link]
                cache[p.link] = (quat_from_matrix(T[:3, :3]), T[:3, 3].copy())
            q, t = cache[p.link]
            pm[i] = _math.part_model(q, t, p.off_q, p.off_t, p.suffix)
        return view, pm

    # ---- device-side forward kinematics inputs (ruf_set_kinematics / ruf_fk_batch_device) ----
    def joint_q(self, k: int) -> np.ndarray:
        """Joint positions at frame k (the same sweep link_poses() uses)."""
        t = k / self.fps
        return np.array([ln.q0 + ln.amp * math.sin(2 * math.pi * ln.freq * t + ln.phase) for ln in self.links])

    def kinematics(self) -> dict:
        jt = {"fixed": 0, "revolute": 1, "prismatic": 2}
        n = len(self.links)
        origin = np.zeros((n, 16))
        axis = np.zeros((n, 3))
        for i, ln in enumerate(self.links):
            origin[i] = make_T(rpy_matrix(*ln.rpy), ln.xyz).T.reshape(-1)      # column-major
            axis[i] = np.asarray(ln.axis, float) / np.linalg.norm(ln.axis)
        part_local = np.zeros((self.n_parts, 16))
        for i, p in enumerate(self.parts):
            part_local[i] = _math.part_model((0, 0, 0, 1), (0, 0, 0), p.off_q, p.off_t, p.suffix)
        return dict(parent=np.array([ln.parent for ln in self.links], np.int32),
                    joint_type=np.array([jt[ln.jtype] for ln in self.links], np.int32),
                    origin=origin, axis=axis, part_link=np.array([p.link for p in self.parts], np.int32),
                    part_local=part_local, cam_link=self.cam_link,
                    cam_mount=make_T(self.cam_R, self.cam_xyz).T.reshape(-1),
                    view_pre=_math.view_matrix(self.offset_q, self.offset_t, (0, 0, 0, 1), (0, 0, 0), 0.0, 0.0))

    def frames(self, ks):
        views = np.zeros((len(ks), 16))
        pms = np.zeros((len(ks), self.n_parts, 16))
        for i, k in enumerate(ks):
            views[i], pms[i] = self.frame(k)
        return views, pms


def kinect_P(width, height, fx=None):
    """Kinect/openni default intrinsics scaled to the image size."""
    f = 525.0 * width / 640.0 if fx is None else fx
    return np.array([f, 0, (width - 1) / 2.0, 0, 0, f, (height - 1) / 2.0, 0, 0, 0, 1, 0], float)


# camera optical frame (z forward, x right, y down) expressed in a body frame (x forward, z up)
OPTICAL_IN_BODY = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


# ------------------------------------------------------------------------------------------------
# primitive renderables, exactly as the reference draws them
# ------------------------------------------------------------------------------------------------
def _add(tris, parts_idx, t, idx):
    tris.append(np.asarray(t, np.float32).reshape(-1, 9))
    parts_idx.append(np.full(len(tris[-1]), idx, np.uint32))


def add_box(parts, tris, pidx, link, dims, off_q=(0, 0, 0, 1), off_t=(0, 0, 0)):
    """RenderableBox::render (src/renderable.cpp:107-131): the VBO box, then
    glScalef(dx,dy,dz); glutSolidCube(dx)  -- two draw calls, two model matrices (F4)."""
    dx, dy, dz = (float(np.float32(v)) for v in dims)
    parts.append(Part(link, off_q, off_t, None))
    _add(tris, pidx, _math.box_triangles(dx, dy, dz), len(parts) - 1)
    parts.append(Part(link, off_q, off_t, scale_suffix(np.float32(dx), np.float32(dy), np.float32(dz))))
    _add(tris, pidx, _math.cube_triangles(dx), len(parts) - 1)


def add_sphere(parts, tris, pidx, link, radius, off_q=(0, 0, 0, 1), off_t=(0, 0, 0)):
    parts.append(Part(link, off_q, off_t, None))
    _add(tris, pidx, _math.sphere_triangles(float(np.float32(radius)), 10, 10), len(parts) - 1)


def add_cylinder(parts, tris, pidx, link, radius, length, off_q=(0, 0, 0, 1), off_t=(0, 0, 0)):
    r, l = float(np.float32(radius)), float(np.float32(length))
    parts.append(Part(link, off_q, off_t, translate_suffix(0, 0, -np.float32(l) / np.float32(2))))
    _add(tris, pidx, _math.cylinder_triangles(r, l, 10, 10), len(parts) - 1)


def add_mesh(parts, tris, pidx, link, tri, scale=(1.0, 1.0, 1.0), off_q=(0, 0, 0, 1), off_t=(0, 0, 0)):
    """RenderableMesh::render (src/renderable.cpp:424-452): glScalef then the indexed triangles."""
    s = [np.float32(v) for v in scale]
    parts.append(Part(link, off_q, off_t, scale_suffix(*s)))
    _add(tris, pidx, tri, len(parts) - 1)


def blob_mesh(rng, radii, slices, stacks, bump=0.04):
    """Closed lat/long mesh of a bumpy ellipsoid: 2*slices*(stacks-1) triangles."""
    a, b = rng.uniform(0, 2 * np.pi, 2)
    th = np.linspace(0, np.pi, stacks + 1)
    ph = np.linspace(0, 2 * np.pi, slices + 1)[:-1]
    TH, PH = np.meshgrid(th, ph, indexing="ij")
    r = 1.0 + bump * np.sin(3 * TH + a) * np.sin(2 * PH + b)
    X = radii[0] * r * np.sin(TH) * np.cos(PH)
    Y = radii[1] * r * np.sin(TH) * np.sin(PH)
    Z = radii[2] * r * np.cos(TH)
    V = np.stack([X, Y, Z], -1)                      # (stacks+1, slices, 3)
    j1 = (np.arange(slices) + 1) % slices
    p00, p01 = V[:-1], V[:-1][:, j1]                 # (stacks, slices, 3)
    p10, p11 = V[1:], V[1:][:, j1]
    upper = np.concatenate([p00, p10, p01], -1)[1:]  # rows i > 0        -> (p00, p10, p01)
    lower = np.concatenate([p01, p10, p11], -1)[:-1] # rows i < stacks-1 -> (p01, p10, p11)
    # same triangle order as the original per-(i, j) loop: for each (i, j): upper (if i>0) then lower
    out = np.zeros((stacks, slices, 2, 9))
    has = np.zeros((stacks, slices, 2), bool)
    out[1:, :, 0] = upper; has[1:, :, 0] = True
    out[:-1, :, 1] = lower; has[:-1, :, 1] = True
    return out[has].astype(np.float32)


def blob_for_budget(rng, radii, budget, bump=0.04):
    stacks = max(3, int(round(math.sqrt(budget / 4.0))))
    slices = max(3, int(round(budget / (2.0 * (stacks - 1)))))
    return blob_mesh(rng, radii, slices, stacks, bump)


def _finish(tris, pidx):
    return np.concatenate(tris, 0).astype(np.float32), np.concatenate(pidx).astype(np.uint32)


# ------------------------------------------------------------------------------------------------
# C1: urdf/example.urdf.xml
# ------------------------------------------------------------------------------------------------
def example_scene(width=640, height=480):
    """Two 4 x 0.5 x 2 box walls at (0,5,0), yaw +-pi/4 (urdf/example.urdf.xml:3-37); camera at the
    world origin looking along world +y (x_c = x_w, y_c = -z_w, z_c = y_w)."""
    links = [Link("wall1", -1, (0, 5, 0), (0, 0, 0.785398163)), Link("wall2", -1, (0, 5, 0), (0, 0, -0.785398163))]
    parts, tris, pidx = [], [], []
    add_box(parts, tris, pidx, 0, (4, 0.5, 2))
    add_box(parts, tris, pidx, 1, (4, 0.5, 2))
    tri, tp = _finish(tris, pidx)
    cam_R = np.array([[1.0, 0, 0], [0, 0, 1.0], [0, -1.0, 0]])   # columns: optical x, y, z in world
    return Scene("example_urdf", width, height, kinect_P(width, height), links, parts, tri, tp, -1, (0, 0, 0), cam_R,
                 label="urdf/example.urdf.xml, static pose")


def example_urdf_xml():
    """A URDF text equivalent to the reference's urdf/example.urdf.xml (two 4 x 0.5 x 2 box walls hanging
    off `world` by fixed joints at (0,5,0), yaw +-pi/4), generated here rather than copied; it keeps the
    stray '>' after </visual> that the original has, which a tolerant parser must skip."""
    def wall(n):
        return (f'  <link name="wall{n}">\n    <visual>\n      <geometry><box size="4 0.5 2" /></geometry>\n'
                f'    </visual>>\n    <collision>\n      <geometry><box size="4 0.5 2" /></geometry>\n'
                f'    </collision>>\n  </link>\n')
    def joint(n, yaw):
        return (f'  <joint name="wall{n}_joint" type="fixed">\n    <origin xyz="0 5 0" rpy="0 0 {yaw}"/>\n'
                f'    <parent link="world"/>\n    <child link="wall{n}"/>\n  </joint>\n')
    return ('<robot name="example">\n  <link name="world"/>\n' + wall(1) + wall(2) + joint(1, "0.785398163")
            + joint(2, "-0.785398163") + '</robot>\n')


# ------------------------------------------------------------------------------------------------
# C2: PR2-like articulated robot
# ------------------------------------------------------------------------------------------------
def pr2_like_scene(width=640, height=480, n_tris=90000, seed=7, name="pr2_like", walls=False, copies=1,
                   spacing=1.2):
    rng = np.random.default_rng(seed)
    links, specs = [], []      # specs: (link index, radii, weight, off_t, scale)

    def L(name, parent, xyz=(0, 0, 0), rpy=(0, 0, 0), axis=(0, 0, 1), jtype="fixed", q0=0.0, amp=0.0):
        links.append(Link(name, parent, xyz, rpy, axis, jtype, q0, amp,
                          freq=float(rng.uniform(0.1, 0.5)), phase=float(rng.uniform(0, 2 * np.pi))))
        return len(links) - 1

    def G(link, radii, weight, off_t=(0, 0, 0), scale=1.0):
        specs.append((link, radii, weight, off_t, scale))

    def accessories(link, n, extent, size=0.015):
        # small visuals riding on a link (LEDs, cameras, motor housings, cable covers ...)
        for _ in range(n):
            off = tuple(float(rng.uniform(-e, e)) for e in extent)
            G(link, tuple(float(rng.uniform(0.6, 1.4) * size) for _ in range(3)), 0.25, off)

    cam_link = -1
    for c in range(copies):
        y0 = (c - (copies - 1) / 2.0) * spacing
        base = L(f"r{c}/base_link", -1, (0.0, y0, 0.051))
        G(base, (0.33, 0.33, 0.13), 10, (0, 0, 0.15))
        G(base, (0.05, 0.05, 0.03), 1, (0.275, 0, 0.25))                         # base laser
        accessories(base, 6, (0.3, 0.3, 0.05))
        for ci, (cx, cy) in enumerate([(0.2246, 0.2246), (0.2246, -0.2246), (-0.2246, 0.2246), (-0.2246, -0.2246)]):
            cas = L(f"r{c}/caster{ci}", base, (cx, cy, 0.0282), jtype="revolute", q0=0.0, amp=0.6)
            G(cas, (0.09, 0.06, 0.05), 1.5, (0, 0, 0.03))
            for wi, wy in enumerate([0.049, -0.049]):
                wh = L(f"r{c}/caster{ci}_wheel{wi}", cas, (0, wy, 0), axis=(0, 1, 0), jtype="revolute", amp=3.0)
                G(wh, (0.078, 0.02, 0.078), 1.2)
        torso = L(f"r{c}/torso_lift", base, (-0.05, 0, 0.74), jtype="prismatic", q0=0.15, amp=0.1)
        G(torso, (0.16, 0.2, 0.42), 8, (0, 0, 0.0))
        G(torso, (0.06, 0.08, 0.05), 1, (0.1, 0, 0.25))                          # imu / sensor bumps
        accessories(torso, 4, (0.15, 0.18, 0.3))
        lmount = L(f"r{c}/laser_tilt_mount", torso, (0.098, 0, 0.227), axis=(0, 1, 0), jtype="revolute", q0=0.3,
                   amp=0.5)
        G(lmount, (0.05, 0.05, 0.04), 1.5)
        pan = L(f"r{c}/head_pan", torso, (-0.017, 0, 0.381), jtype="revolute", q0=0.0, amp=0.15)
        G(pan, (0.09, 0.12, 0.05), 2, (0, 0, 0.03))
        tilt = L(f"r{c}/head_tilt", pan, (0.068, 0, 0), axis=(0, 1, 0), jtype="revolute", q0=0.55, amp=0.06)
        G(tilt, (0.08, 0.15, 0.07), 4, (0.02, 0, 0.08))
        for si, sy in enumerate([-0.09, -0.045, 0.0, 0.045, 0.09]):                 # stereo / prosilica / projector
            s = L(f"r{c}/head_sensor{si}", tilt, (0.09, sy, 0.1))
            G(s, (0.02, 0.018, 0.018), 0.5)
        accessories(tilt, 6, (0.07, 0.14, 0.05))
        if c == 0:
            cam_link = tilt
        for side, sy in (("l", 0.188), ("r", -0.188)):
            sgn = 1.0 if side == "l" else -1.0
            sp = L(f"r{c}/{side}_shoulder_pan", torso, (0, sy, 0), jtype="revolute", q0=0.12 * sgn, amp=0.15)
            G(sp, (0.12, 0.1, 0.2), 5, (0.02, 0, -0.1))
            sl = L(f"r{c}/{side}_shoulder_lift", sp, (0.1, 0, 0), axis=(0, 1, 0), jtype="revolute", q0=0.25, amp=0.2)
            G(sl, (0.09, 0.08, 0.08), 3)
            ur = L(f"r{c}/{side}_upper_arm_roll", sl, (0, 0, 0), axis=(1, 0, 0), jtype="revolute", q0=0.3 * sgn, amp=0.3)
            G(ur, (0.06, 0.06, 0.06), 1.5, (0.08, 0, 0))
            ua = L(f"r{c}/{side}_upper_arm", ur)
            G(ua, (0.17, 0.07, 0.07), 6, (0.21, 0, 0), scale=0.001)             # mm mesh + 0.001 scale
            ef = L(f"r{c}/{side}_elbow_flex", ua, (0.4, 0, 0), axis=(0, 1, 0), jtype="revolute", q0=-1.35, amp=0.35)
            G(ef, (0.07, 0.06, 0.06), 2.5)
            fr = L(f"r{c}/{side}_forearm_roll", ef, (0, 0, 0), axis=(1, 0, 0), jtype="revolute", q0=0.0, amp=0.8)
            G(fr, (0.05, 0.05, 0.05), 1.2, (0.05, 0, 0))
            fa = L(f"r{c}/{side}_forearm", fr)
            G(fa, (0.14, 0.055, 0.055), 6, (0.18, 0, 0), scale=0.001)
            G(fa, (0.02, 0.015, 0.015), 0.4, (0.135, 0, 0.045))                  # forearm camera
            wf = L(f"r{c}/{side}_wrist_flex", fa, (0.321, 0, 0), axis=(0, 1, 0), jtype="revolute", q0=-0.4, amp=0.4)
            G(wf, (0.04, 0.04, 0.04), 1.5)
            wr = L(f"r{c}/{side}_wrist_roll", wf, (0, 0, 0), axis=(1, 0, 0), jtype="revolute", q0=0.0, amp=1.0)
            G(wr, (0.03, 0.035, 0.035), 1.0, (0.03, 0, 0))
            palm = L(f"r{c}/{side}_gripper_palm", wr, (0.0, 0, 0))
            G(palm, (0.05, 0.05, 0.025), 3, (0.075, 0, 0))
            G(palm, (0.012, 0.012, 0.01), 0.3, (0.06, 0, 0.03))                  # accelerometer
            accessories(ua, 3, (0.15, 0.06, 0.06))
            accessories(fa, 3, (0.12, 0.05, 0.05))
            accessories(palm, 2, (0.04, 0.04, 0.02), 0.008)
            for fi, fy in (("l", 0.01), ("r", -0.01)):
                fs = 1.0 if fi == "l" else -1.0
                fg = L(f"r{c}/{side}_gripper_{fi}_finger", palm, (0.07691, fy, 0), axis=(0, 0, fs), jtype="revolute",
                       q0=0.25, amp=0.2)
                G(fg, (0.045, 0.012, 0.012), 1.2, (0.045, fs * 0.01, 0))
                ft = L(f"r{c}/{side}_gripper_{fi}_finger_tip", fg, (0.09137, fs * 0.00495, 0), axis=(0, 0, -fs),
                       jtype="revolute", q0=0.25, amp=0.2)
                G(ft, (0.02, 0.008, 0.011), 0.8, (0.02, fs * 0.005, 0))

    parts, tris, pidx = [], [], []
    wall_tris = 0
    if walls:
        # C3: two static wall meshes (Automatica-style cell walls): box-shaped RenderableMesh parts, 2.6 m ahead
        w1 = len(links); links.append(Link("wall1", -1, (2.6, 0.9, 1.0), (0, 0, math.pi / 2 + 0.5)))
        w2 = len(links); links.append(Link("wall2", -1, (2.6, -0.9, 1.0), (0, 0, math.pi / 2 - 0.5)))
        slab = _math.box_triangles(3.0, 0.2, 2.5)
        add_mesh(parts, tris, pidx, w1, slab)
        add_mesh(parts, tris, pidx, w2, slab)
        wall_tris = 24
    wsum = sum(s[2] for s in specs)
    for (link, radii, weight, off_t, scale) in specs:
        budget = max(16, n_tris * weight / wsum)
        m = blob_for_budget(rng, radii, budget)
        if scale != 1.0:
            m = (m / np.float32(scale)).astype(np.float32)
        yaw = float(rng.uniform(-0.2, 0.2))
        off_q = (0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2))
        add_mesh(parts, tris, pidx, link, m, (scale, scale, scale), off_q, off_t)
    tri, tp = _finish(tris, pidx)
    label = f"PR2-like synthetic ({len(parts)} parts, {tri.shape[0]} triangles)"
    # head camera: optical frame on the head-tilt link, looking along the link's +x
    sc = Scene(name, width, height, kinect_P(width, height), links, parts, tri, tp, cam_link, (0.07, 0.03, 0.11),
               OPTICAL_IN_BODY.copy(), label=label)
    sc.wall_tris = wall_tris
    return sc


def walls_scene(width=1280, height=960):
    """C3: PR2-like model plus two static wall boxes, 1280x960."""
    return pr2_like_scene(width, height, 90000, seed=7, name="pr2_like_walls", walls=True)


def multi_robot_scene(width=1920, height=1080, n_tris=500000):
    """C5: four articulated PR2-like URDFs (~125k triangles each) with animated joint sweep; the
    camera rides on robot 0's head but is pulled back so that the neighbours are in view."""
    sc = pr2_like_scene(width, height, n_tris, seed=11, name="pr2_like_x4", copies=4)
    # external camera 3.2 m behind the row of robots, 1.3 m up, pitched 0.15 rad down, looking along +x
    c, s_ = math.cos(0.15), math.sin(0.15)
    sc.cam_link = -1
    sc.cam_xyz = (-3.2, 0.0, 1.3)
    sc.cam_R = np.array([[0.0, -s_, c], [-1.0, 0.0, 0.0], [0.0, -c, -s_]])   # columns: right, down, forward
    return sc


# ------------------------------------------------------------------------------------------------
# depth frames
# ------------------------------------------------------------------------------------------------
def linear_depth(zbuf):
    """to_linear_depth of the fragment shader in float32 numpy (input synthesis only)."""
    zn, zf = np.float32(Z_NEAR), np.float32(Z_FAR)
    k1 = (zn * zf) / (zn - zf)
    k2 = zf / (zf - zn)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (k1 / (zbuf.astype(np.float32) - k2)).astype(np.float32)


def synth_depth(virt_m, frame, enc="u16", bg_m=7.92):
    """Sensor frame for a given virtual depth image (metres; >= bg_m - eps where nothing was hit):
    model surface / 3 m back wall + N(0, 3 mm) noise, 10 % invalid, 5 % occluders 0.3 m in front,
    2 % beyond 7.9 m.  enc: "u16" (mm, 0 invalid) or "f32" (m, NaN invalid)."""
    rng = np.random.default_rng(1234 + frame)
    H, W = virt_m.shape
    hit = virt_m < (bg_m - 0.01)
    d = np.where(hit, virt_m, 3.0).astype(np.float64)
    d += rng.normal(0.0, 0.003, d.shape)
    # occluder blobs: 5 % of the area as discs
    n_blobs = max(1, int(0.05 * H * W / (math.pi * 12 * 12)))
    rad = 12
    oy, ox = np.mgrid[-rad:rad + 1, -rad:rad + 1]
    disc = (oy * oy + ox * ox) <= rad * rad
    for _ in range(n_blobs):
        cy, cx = int(rng.integers(0, H)), int(rng.integers(0, W))
        y0, y1, x0, x1 = max(cy - rad, 0), min(cy + rad + 1, H), max(cx - rad, 0), min(cx + rad + 1, W)
        m = disc[y0 - cy + rad:y1 - cy + rad, x0 - cx + rad:x1 - cx + rad]
        sub = d[y0:y1, x0:x1]
        sub[m] = np.maximum(0.2, sub[m] - 0.3)
    far = rng.random(d.shape) < 0.02
    d[far] = rng.uniform(7.9, 9.5, int(far.sum()))
    invalid = rng.random(d.shape) < 0.10
    if enc == "u16":
        u = np.clip(np.rint(d * 1000.0), 0, 65535).astype(np.uint16)
        u[invalid] = 0
        return u
    f = d.astype(np.float32)
    f[invalid] = np.nan
    return f
