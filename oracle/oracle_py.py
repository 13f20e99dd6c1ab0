"""ctypes loader for the CPU oracle (oracle/ruf_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from the realtime_urdf_filter_b200
package.  PARITY STATUS of the oracle itself: pinned against the reference's own shaders run on Mesa
llvmpipe (oracle/gl_ref, tests/golden/gl_llvmpipe.npz; see the header of ruf_oracle.c and DESIGN.md section 2).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
ENC_F32_M, ENC_U16_MM = 0, 1

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_libs: dict[str, C.CDLL] = {}


def build(native: bool = False) -> str:
    target = "native" if native else "all"
    name = "libruf_oracle_native.so" if native else "libruf_oracle.so"
    path = os.path.join(BUILD, name)
    src = os.path.join(HERE, "ruf_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        res = subprocess.run(["make", "-C", HERE, target], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return path


def load(native: bool = False) -> C.CDLL:
    key = "native" if native else "portable"
    if key in _libs:
        return _libs[key]
    lib = C.CDLL(build(native))
    lib.orc_projection_matrix.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp]
    lib.orc_lookat.argtypes = [_dp]
    lib.orc_transform_to_gl.argtypes = [_dp, _dp, _dp]
    lib.orc_view_matrix.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    lib.orc_link_model.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp]
    lib.orc_compose_mvp.argtypes = [_dp, _dp, _dp, C.c_int, _fp]
    lib.orc_box_triangles.argtypes = [C.c_float, C.c_float, C.c_float, _fp]
    lib.orc_box_triangles.restype = C.c_int
    lib.orc_cube_triangles.argtypes = [C.c_float, _fp]
    lib.orc_cube_triangles.restype = C.c_int
    lib.orc_sphere_triangles.argtypes = [C.c_float, C.c_int, C.c_int, _fp]
    lib.orc_sphere_triangles.restype = C.c_int
    lib.orc_cylinder_triangles.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, _fp]
    lib.orc_cylinder_triangles.restype = C.c_int
    lib.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_float, C.c_void_p, C.c_int]
    lib.orc_render.restype = C.c_int
    lib.orc_u16_to_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.orc_f32_to_u16.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.orc_to_linear_depth.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.orc_to_linear_depth.restype = C.c_float
    lib.orc_filter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                               C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_filter_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_filter_frame.restype = C.c_int
    lib.orc_max_threads.restype = C.c_int
    lib.orc_sincos.argtypes = [C.c_double, _dp, _dp]
    lib.orc_fk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp]
    lib.orc_fk_outputs.argtypes = [_dp, C.c_int, C.c_void_p, _dp, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp, _dp]
    _libs[key] = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, n):
    a = np.ascontiguousarray(a, np.float64).reshape(-1)
    assert a.size == n, (a.size, n)
    return a


def projection_matrix(P, width, height, z_near=0.1, z_far=8.0):
    out = np.zeros(16)
    tx, ty = C.c_double(), C.c_double()
    load().orc_projection_matrix(_d(_f64(P, 12)), width, height, z_near, z_far, _d(out), C.byref(tx), C.byref(ty))
    return out, tx.value, ty.value


def lookat():
    out = np.zeros(16)
    load().orc_lookat(_d(out))
    return out


def view_matrix(offset_q, offset_t, cam_q, cam_t, tx=0.0, ty=0.0):
    out = np.zeros(16)
    load().orc_view_matrix(_d(_f64(offset_q, 4)), _d(_f64(offset_t, 3)), _d(_f64(cam_q, 4)), _d(_f64(cam_t, 3)),
                           tx, ty, _d(out))
    return out


def link_model(link_q, link_t, off_q=(0, 0, 0, 1), off_t=(0, 0, 0), suffix=None):
    out = np.zeros(16)
    s = None if suffix is None else _f64(suffix, 16)
    load().orc_link_model(_d(_f64(link_q, 4)), _d(_f64(link_t, 3)), _d(_f64(off_q, 4)), _d(_f64(off_t, 3)),
                          None if s is None else _d(s), _d(out))
    return out


def compose_mvp(proj, view, link_models, L):
    lm = _f64(link_models, 16 * L) if L else np.zeros(1)
    out = np.zeros((L + 1, 16), np.float32)
    load().orc_compose_mvp(_d(_f64(proj, 16)), _d(_f64(view, 16)), _d(lm), L, out.ctypes.data_as(_fp))
    return out


def box_triangles(dx, dy, dz):
    out = np.zeros((12, 9), np.float32)
    n = load().orc_box_triangles(dx, dy, dz, out.ctypes.data_as(_fp))
    return out[:n]


def cube_triangles(size):
    out = np.zeros((12, 9), np.float32)
    n = load().orc_cube_triangles(size, out.ctypes.data_as(_fp))
    return out[:n]


def sphere_triangles(radius, slices=10, stacks=10):
    out = np.zeros((2 * slices * stacks, 9), np.float32)
    n = load().orc_sphere_triangles(radius, slices, stacks, out.ctypes.data_as(_fp))
    return out[:n].copy()


def cylinder_triangles(radius, height, slices=10, stacks=10):
    out = np.zeros((2 * slices * (stacks + 1), 9), np.float32)
    n = load().orc_cylinder_triangles(radius, height, slices, stacks, out.ctypes.data_as(_fp))
    return out[:n].copy()


def render(tri, tri_link, mvp, W, H, bg_z, nthreads=1, native=False):
    tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
    tl = np.ascontiguousarray(tri_link, np.uint32)
    mvp = np.ascontiguousarray(mvp, np.float32)
    L = mvp.reshape(-1, 16).shape[0] - 1
    z = np.empty((H, W), np.float32)
    rc = load(native).orc_render(tri.ctypes.data, tl.ctypes.data, tri.shape[0], mvp.ctypes.data, L, W, H,
                                 bg_z, z.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_render rc={rc}")
    return z


def filter_frame(depth, tri, tri_link, mvp, z_near, z_far, max_diff, replace_value, want_mask=True,
                 nthreads=1, native=False, want_zbuf=False):
    """Whole frame through the oracle -> (depth_out, mask, zbuf)."""
    H, W = depth.shape
    enc = ENC_U16_MM if depth.dtype == np.uint16 else ENC_F32_M
    d = np.ascontiguousarray(depth)
    tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
    tl = np.ascontiguousarray(tri_link, np.uint32)
    mvp = np.ascontiguousarray(mvp, np.float32)
    L = mvp.reshape(-1, 16).shape[0] - 1
    out = np.empty_like(d)
    mask = np.empty((H, W), np.uint8) if want_mask else None
    zbuf = np.empty((H, W), np.float32) if want_zbuf else None
    rc = load(native).orc_filter_frame(d.ctypes.data, enc, W, H, tri.ctypes.data, tl.ctypes.data, tri.shape[0],
                                       mvp.ctypes.data, L, z_near, z_far, max_diff, replace_value,
                                       out.ctypes.data, mask.ctypes.data if want_mask else None,
                                       zbuf.ctypes.data if want_zbuf else None, nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_filter_frame rc={rc}")
    return out, mask, zbuf


def to_linear_depth(d, z_near=0.1, z_far=8.0):
    return float(load().orc_to_linear_depth(d, z_near, z_far))


def u16_to_f32(a):
    a = np.ascontiguousarray(a, np.uint16)
    out = np.empty(a.shape, np.float32)
    load().orc_u16_to_f32(a.ctypes.data, out.ctypes.data, a.size)
    return out


def f32_to_u16(a):
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty(a.shape, np.uint16)
    load().orc_f32_to_u16(a.ctypes.data, out.ctypes.data, a.size)
    return out


def max_threads() -> int:
    return int(load().orc_max_threads())


def sincos(x):
    s, c = C.c_double(), C.c_double()
    load().orc_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def fk(kin: dict, q):
    """Oracle forward kinematics -> (links[n,16], part_models[P,16], view[16]) for joint positions q."""
    n = len(kin["parent"])
    parent = np.ascontiguousarray(kin["parent"], np.int32)
    jt = np.ascontiguousarray(kin["joint_type"], np.int32)
    origin = _f64(kin["origin"], 16 * n) if n else np.zeros(1)
    axis = _f64(kin["axis"], 3 * n) if n else np.zeros(1)
    qq = _f64(q, n) if n else np.zeros(1)
    links = np.zeros((max(n, 1), 16))
    load().orc_fk(n, parent.ctypes.data, jt.ctypes.data, _d(origin), _d(axis), _d(qq), _d(links))
    P = len(kin["part_link"])
    pl = np.ascontiguousarray(kin["part_link"], np.int32)
    plocal = _f64(kin["part_local"], 16 * P) if P else np.zeros(1)
    pm = np.zeros((max(P, 1), 16))
    view = np.zeros(16)
    load().orc_fk_outputs(_d(links), P, pl.ctypes.data, _d(plocal), int(kin["cam_link"]), _d(_f64(kin["cam_mount"], 16)),
                          _d(_f64(kin["view_pre"], 16)), float(kin.get("tx", 0.0)), float(kin.get("ty", 0.0)), _d(pm), _d(view))
    return links[:n], pm[:P], view
