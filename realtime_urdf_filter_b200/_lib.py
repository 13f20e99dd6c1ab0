"""ctypes binding of libruf_b200.so (the C ABI declared in include/ruf_b200.h).

This is the binding a Python host would use; the ROS/C++ host links the same library directly
(INTEGRATION.md).  There is no fallback: if the shared library is missing the import of the
compute entry points raises, and every compute call fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

RUF_OK = 0
RUF_ERR_INVALID = -1
RUF_ERR_CUDA = -2
RUF_ERR_NO_MODEL = -3
RUF_ERR_OVERFLOW = -4
RUF_ERR_NOMEM = -5

ENC_F32_M = 0
ENC_U16_MM = 1
MASK_BYTES = 0
MASK_BITS = 1

_c_double_p = C.POINTER(C.c_double)
_c_float_p = C.POINTER(C.c_float)


class RufStats(C.Structure):
    _fields_ = [
        ("frames", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("visible_tris", C.c_int64),
        ("binned_refs", C.c_int64),
        ("big_tris", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
    ]


# name -> (restype, argtypes); mirrors include/ruf_b200.h one to one
SIGNATURES = {
    "ruf_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]),
    "ruf_destroy": (C.c_int, [C.c_void_p]),
    "ruf_last_error": (C.c_char_p, [C.c_void_p]),
    "ruf_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ruf_sync": (C.c_int, [C.c_void_p]),
    "ruf_set_model": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "ruf_set_model_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "ruf_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64]),
    "ruf_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "ruf_filter_batch_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "ruf_filter_batch_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "ruf_set_mask_format": (C.c_int, [C.c_void_p, C.c_int]),
    "ruf_host_copy_ceiling": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "ruf_group_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]),
    "ruf_group_destroy": (C.c_int, [C.c_void_p]),
    "ruf_group_size": (C.c_int, [C.c_void_p]),
    "ruf_group_context": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ruf_group_last_error": (C.c_char_p, [C.c_void_p]),
    "ruf_group_set_model": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "ruf_group_broadcast_bytes": (C.c_int64, [C.c_void_p]),
    "ruf_group_filter_batch_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int]),
    "ruf_set_kinematics": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ruf_fk_batch_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "ruf_filter_batch_device_fk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_double, C.c_double, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    "ruf_meshlet_roundtrip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ruf_meshlet_sets_roundtrip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ruf_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ruf_host_free": (C.c_int, [C.c_void_p]),
    "ruf_host_is_pinned": (C.c_int, [C.c_void_p]),
    "ruf_get_stats": (C.c_int, [C.c_void_p, C.POINTER(RufStats)]),
    "ruf_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ruf_get_stage_times": (C.c_int, [C.c_void_p, _c_double_p, C.POINTER(C.c_int64), C.c_int]),
    "ruf_projection_matrix": (None, [_c_double_p, C.c_int, C.c_int, C.c_double, C.c_double, _c_double_p,
                                     _c_double_p, _c_double_p]),
    "ruf_lookat": (None, [_c_double_p]),
    "ruf_view_matrix": (None, [_c_double_p, _c_double_p, _c_double_p, _c_double_p, C.c_double, C.c_double,
                               _c_double_p]),
    "ruf_part_model": (None, [_c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p]),
    "ruf_box_triangles": (C.c_int, [C.c_float, C.c_float, C.c_float, _c_float_p]),
    "ruf_cube_triangles": (C.c_int, [C.c_float, _c_float_p]),
    "ruf_sphere_triangles": (C.c_int, [C.c_float, C.c_int, C.c_int, _c_float_p]),
    "ruf_cylinder_triangles": (C.c_int, [C.c_float, C.c_float, C.c_int, C.c_int, _c_float_p]),
    "ruf_sphere_triangle_count": (C.c_int, [C.c_int, C.c_int]),
    "ruf_cylinder_triangle_count": (C.c_int, [C.c_int, C.c_int]),
    "ruf_version": (C.c_char_p, []),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libruf_b200.so (building it in-tree first when nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and _build.needs_build():
        try:
            _build.build_library()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -m realtime_urdf_filter_b200.build` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RufError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ruf error {code}: {msg}")
        self.code = code


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_double_p)


def _as_f64(a, n=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} doubles, got {a.size}")
    return a


# ---------------------------------------------------------------------------------------------
# host math / geometry wrappers (no GPU needed)
# ---------------------------------------------------------------------------------------------
def projection_matrix(P, width, height, z_near=0.1, z_far=8.0):
    """getProjectionMatrix (src/urdf_filter.cpp:459-501) -> (glTf[16], camera_tx, camera_ty)."""
    lib = load()
    P = _as_f64(P, 12)
    out = np.zeros(16)
    tx, ty = C.c_double(0), C.c_double(0)
    lib.ruf_projection_matrix(_dp(P), int(width), int(height), float(z_near), float(z_far), _dp(out),
                              C.byref(tx), C.byref(ty))
    return out, tx.value, ty.value


def lookat():
    out = np.zeros(16)
    load().ruf_lookat(_dp(out))
    return out


def view_matrix(offset_q, offset_t, cam_q, cam_t, camera_tx=0.0, camera_ty=0.0):
    out = np.zeros(16)
    # every converted array is bound to a local: a temporary would be freed before the C call reads it
    oq, ot, cq, ct = _as_f64(offset_q, 4), _as_f64(offset_t, 3), _as_f64(cam_q, 4), _as_f64(cam_t, 3)
    load().ruf_view_matrix(_dp(oq), _dp(ot), _dp(cq), _dp(ct), float(camera_tx), float(camera_ty), _dp(out))
    return out


def part_model(link_q, link_t, off_q=(0, 0, 0, 1), off_t=(0, 0, 0), suffix=None):
    out = np.zeros(16)
    sfx = None if suffix is None else _as_f64(suffix, 16)
    lq, lt, oq, ot = _as_f64(link_q, 4), _as_f64(link_t, 3), _as_f64(off_q, 4), _as_f64(off_t, 3)
    load().ruf_part_model(_dp(lq), _dp(lt), _dp(oq), _dp(ot), None if sfx is None else _dp(sfx), _dp(out))
    return out


def _fp(a):
    return a.ctypes.data_as(_c_float_p)


def box_triangles(dx, dy, dz):
    out = np.zeros((12, 9), np.float32)
    n = load().ruf_box_triangles(dx, dy, dz, _fp(out))
    return out[:n]


def cube_triangles(size):
    out = np.zeros((12, 9), np.float32)
    n = load().ruf_cube_triangles(size, _fp(out))
    return out[:n]


def sphere_triangles(radius, slices=10, stacks=10):
    lib = load()
    out = np.zeros((lib.ruf_sphere_triangle_count(slices, stacks), 9), np.float32)
    n = lib.ruf_sphere_triangles(radius, slices, stacks, _fp(out))
    return out[:n]


def cylinder_triangles(radius, height, slices=10, stacks=10):
    lib = load()
    out = np.zeros((lib.ruf_cylinder_triangle_count(slices, stacks), 9), np.float32)
    n = lib.ruf_cylinder_triangles(radius, height, slices, stacks, _fp(out))
    return out[:n]


# ---------------------------------------------------------------------------------------------
# context wrapper
# ---------------------------------------------------------------------------------------------
def meshlet_roundtrip(tri_xyz, tri_part, n_parts, z_far=8.0, max_verts=256, max_tris=512, max_parts=32):
    """Model-ingest diagnostics (CPU): soup -> meshlets -> soup.  -> (xyz (T+2, 9), part (T+2,), counts dict)."""
    lib = load()
    xyz = np.ascontiguousarray(tri_xyz, dtype=np.float32).reshape(-1, 9)
    part = np.ascontiguousarray(tri_part, dtype=np.uint32)
    n = xyz.shape[0]
    out_xyz = np.empty((n + 2, 9), np.float32)
    out_part = np.empty(n + 2, np.uint32)
    counts = np.zeros(3, np.int64)
    rc = lib.ruf_meshlet_roundtrip(xyz.ctypes.data, part.ctypes.data, n, int(n_parts), float(z_far), max_verts, max_tris,
                                   max_parts, out_xyz.ctypes.data, out_part.ctypes.data, counts.ctypes.data)
    if rc != RUF_OK:
        raise RufError(rc, "ruf_meshlet_roundtrip failed")
    return out_xyz, out_part, dict(meshlets=int(counts[0]), verts=int(counts[1]), tris=int(counts[2]))


def meshlet_sets_roundtrip(tri_xyz, tri_part, n_parts, z_far=8.0, max_verts=512, max_tris=1023, fine_tris=256, max_parts=32):
    """Both cuts of the model as ruf_set_model uploads them (throughput cut, then fine cut), each expanded to a soup again.
    -> (xyz (2, T+2, 9), part (2, T+2), counts dict)."""
    lib = load()
    xyz = np.ascontiguousarray(tri_xyz, dtype=np.float32).reshape(-1, 9)
    part = np.ascontiguousarray(tri_part, dtype=np.uint32)
    n = xyz.shape[0]
    out_xyz = np.empty((2, n + 2, 9), np.float32)
    out_part = np.empty((2, n + 2), np.uint32)
    counts = np.zeros(4, np.int64)
    rc = lib.ruf_meshlet_sets_roundtrip(xyz.ctypes.data, part.ctypes.data, n, int(n_parts), float(z_far), max_verts, max_tris,
                                        fine_tris, max_parts, out_xyz.ctypes.data, out_part.ctypes.data, counts.ctypes.data)
    if rc != RUF_OK:
        raise RufError(rc, "ruf_meshlet_sets_roundtrip failed")
    return out_xyz, out_part, dict(meshlets=int(counts[0]), fine_meshlets=int(counts[1]), verts=int(counts[2]), tris=int(counts[3]))


class Context:
    """Thin RAII wrapper over ruf_context."""

    def __init__(self, width: int, height: int, device: int = 0, z_near: float = 0.1, z_far: float = 8.0):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.ruf_create(C.byref(self._h), device, width, height, z_near, z_far)
        if rc != RUF_OK:
            raise RufError(rc, self._lib.ruf_last_error(None).decode())
        self.width, self.height, self.device = width, height, device
        self.n_parts = 0
        self.n_tris = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ruf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != RUF_OK:
            raise RufError(rc, self._lib.ruf_last_error(self._h).decode())

    def set_stream(self, cuda_stream: int | None):
        self._check(self._lib.ruf_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self._check(self._lib.ruf_sync(self._h))

    def set_mask_format(self, fmt: int):
        """MASK_BYTES (0 / 255 per pixel, default) or MASK_BITS (1 bit per pixel, little bit order; width % 8 == 0)."""
        self._check(self._lib.ruf_set_mask_format(self._h, int(fmt)))
        self.mask_format = int(fmt)

    def _mask_shape(self, n=None):
        w = self.width // 8 if getattr(self, "mask_format", MASK_BYTES) == MASK_BITS else self.width
        return (self.height, w) if n is None else (n, self.height, w)

    def set_model(self, tri_xyz: np.ndarray, tri_part: np.ndarray, n_parts: int):
        tri = np.ascontiguousarray(tri_xyz, np.float32).reshape(-1, 9)
        part = np.ascontiguousarray(tri_part, np.uint32).reshape(-1)
        if part.size != tri.shape[0]:
            raise ValueError("tri_part length mismatch")
        self._check(self._lib.ruf_set_model(self._h, tri.ctypes.data, part.ctypes.data, tri.shape[0], int(n_parts)))
        self.n_parts, self.n_tris = int(n_parts), tri.shape[0]

    def set_model_device(self, d_tri_ptr: int, d_part_ptr: int, n_tris: int, n_parts: int):
        self._check(self._lib.ruf_set_model_device(self._h, C.c_void_p(d_tri_ptr), C.c_void_p(d_part_ptr),
                                                   int(n_tris), int(n_parts)))
        self.n_parts, self.n_tris = int(n_parts), int(n_tris)

    def reserve(self, max_batch: int, big_capacity: int = 0, bin_capacity: int = 0):
        """bin_capacity: records per (frame, tile) list (0 = automatic)."""
        self._check(self._lib.ruf_reserve(self._h, max_batch, big_capacity, bin_capacity))

    def _enc_dtype(self, enc):
        return np.uint16 if enc == ENC_U16_MM else np.float32

    def filter(self, depth, proj, view, part_models, max_diff, replace_value, want_mask=True):
        """One frame with host numpy buffers -> (depth_out, mask or None)."""
        enc = ENC_U16_MM if depth.dtype == np.uint16 else ENC_F32_M
        d = np.ascontiguousarray(depth, self._enc_dtype(enc)).reshape(self.height, self.width)
        out = np.empty_like(d)
        mask = np.empty(self._mask_shape(), np.uint8) if want_mask else None
        pm = _as_f64(part_models, 16 * self.n_parts)
        pr, vw = _as_f64(proj, 16), _as_f64(view, 16)     # locals: the copies must outlive the call
        self._check(self._lib.ruf_filter(self._h, d.ctypes.data, enc, pr.ctypes.data,
                                         vw.ctypes.data, pm.ctypes.data if pm.size else None,
                                         max_diff, replace_value, out.ctypes.data,
                                         mask.ctypes.data if want_mask else None))
        return out, mask

    def _frames(self, a, dtype, n, what):
        """A caller-supplied (n, H, W) buffer: right dtype, C-contiguous, right shape -- never converted silently
        (an output must be written in place)."""
        if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous or \
                a.shape != (n, self.height, self.width):
            raise ValueError(f"{what} must be a C-contiguous {np.dtype(dtype).name} array of shape "
                             f"({n}, {self.height}, {self.width})")
        return a

    def filter_batch_host(self, depth, proj, views, part_models, max_diff, replace_value, out=None, mask=None,
                          want_mask=True):
        """n frames with host buffers (numpy arrays or raw pointers via .ctypes.data)."""
        depth = np.asarray(depth)
        enc = ENC_U16_MM if depth.dtype == np.uint16 else ENC_F32_M
        dt = self._enc_dtype(enc)
        if depth.ndim != 3 or depth.shape[1:] != (self.height, self.width):
            raise ValueError(f"depth must have shape (n, {self.height}, {self.width}), got {depth.shape}")
        depth = np.ascontiguousarray(depth, dt)        # float64 / strided input is converted like Context.filter does
        n = depth.shape[0]
        out = np.empty_like(depth) if out is None else self._frames(out, dt, n, "out")
        if mask is None and want_mask:
            mask = np.empty(self._mask_shape(n), np.uint8)
        elif mask is not None and (mask.dtype != np.uint8 or not mask.flags.c_contiguous or mask.shape != self._mask_shape(n)):
            raise ValueError(f"mask must be a C-contiguous uint8 array of shape {self._mask_shape(n)}")
        v = _as_f64(views, 16 * n)
        pm = _as_f64(part_models, 16 * n * self.n_parts)
        pr = _as_f64(proj, 16)
        self._check(self._lib.ruf_filter_batch_host(self._h, n, depth.ctypes.data, enc,
                                                    pr.ctypes.data, v.ctypes.data,
                                                    pm.ctypes.data if pm.size else None, max_diff, replace_value,
                                                    out.ctypes.data, mask.ctypes.data if mask is not None else None))
        return out, mask

    def filter_batch_device(self, n_frames, d_depth_in, enc, d_proj, d_view, d_part_model, max_diff,
                            replace_value, d_depth_out, d_mask_out=0, d_zbuf_out=0):
        """Raw device pointers (ints); asynchronous on the context's stream."""
        self._check(self._lib.ruf_filter_batch_device(
            self._h, n_frames, C.c_void_p(d_depth_in), enc, C.c_void_p(d_proj), C.c_void_p(d_view),
            C.c_void_p(d_part_model or 0), max_diff, replace_value, C.c_void_p(d_depth_out),
            C.c_void_p(d_mask_out or 0), C.c_void_p(d_zbuf_out or 0)))

    def set_kinematics(self, parent, joint_type, origin, axis, part_link, part_local, cam_link, cam_mount, view_pre):
        """Kinematic tree for the device-side forward kinematics (ruf_set_kinematics)."""
        parent = np.ascontiguousarray(parent, np.int32)
        jt = np.ascontiguousarray(joint_type, np.int32)
        n = parent.size
        origin = _as_f64(origin, 16 * n)
        axis = _as_f64(axis, 3 * n)
        pl = np.ascontiguousarray(part_link, np.int32)
        plocal = _as_f64(part_local, 16 * self.n_parts)
        mount, pre = _as_f64(cam_mount, 16), _as_f64(view_pre, 16)
        self._check(self._lib.ruf_set_kinematics(self._h, n, parent.ctypes.data, jt.ctypes.data, origin.ctypes.data,
                                                 axis.ctypes.data, pl.ctypes.data, plocal.ctypes.data, int(cam_link),
                                                 mount.ctypes.data, pre.ctypes.data))
        self.n_links = n

    def fk_batch_device(self, n_frames, d_joint_q, tx, ty, d_part_model_out, d_view_out):
        self._check(self._lib.ruf_fk_batch_device(self._h, n_frames, C.c_void_p(d_joint_q), tx, ty,
                                                  C.c_void_p(d_part_model_out), C.c_void_p(d_view_out)))

    def filter_batch_device_fk(self, n_frames, d_depth_in, enc, d_proj, d_joint_q, tx, ty, max_diff, replace_value,
                               d_depth_out, d_mask_out=0, d_zbuf_out=0):
        self._check(self._lib.ruf_filter_batch_device_fk(
            self._h, n_frames, C.c_void_p(d_depth_in), enc, C.c_void_p(d_proj), C.c_void_p(d_joint_q), tx, ty,
            max_diff, replace_value, C.c_void_p(d_depth_out), C.c_void_p(d_mask_out or 0), C.c_void_p(d_zbuf_out or 0)))

    STAGES = ("pose", "setup_bin", "raster_filter")

    def set_profiling(self, enable: bool):
        self._check(self._lib.ruf_set_profiling(self._h, int(enable)))

    def stage_times(self, reset: bool = True):
        """-> ({stage: accumulated ms}, launch sequences covered)."""
        ms = np.zeros(len(self.STAGES))
        calls = C.c_int64(0)
        self._check(self._lib.ruf_get_stage_times(self._h, _dp(ms), C.byref(calls), int(reset)))
        return dict(zip(self.STAGES, ms.tolist())), calls.value

    def stats(self) -> dict:
        s = RufStats()
        self._check(self._lib.ruf_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in RufStats._fields_}


class Group:
    """ruf_group: N devices in one process (one context + one host thread per device, NCCL broadcast of the model)."""

    def __init__(self, width: int, height: int, devices=None, n_devices: int | None = None, z_near=0.1, z_far=8.0):
        self._lib = load()
        self._h = C.c_void_p()
        devs = None if devices is None else np.ascontiguousarray(devices, np.int32)
        n = int(n_devices if devs is None else devs.size)
        rc = self._lib.ruf_group_create(C.byref(self._h), n, None if devs is None else devs.ctypes.data, width, height,
                                        z_near, z_far)
        if rc != RUF_OK:
            raise RufError(rc, self._lib.ruf_group_last_error(None).decode())
        self.width, self.height, self.size, self.n_parts = width, height, n, 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ruf_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != RUF_OK:
            raise RufError(rc, self._lib.ruf_group_last_error(self._h).decode())

    def set_model(self, tri_xyz, tri_part, n_parts: int):
        tri = np.ascontiguousarray(tri_xyz, np.float32).reshape(-1, 9)
        part = np.ascontiguousarray(tri_part, np.uint32).reshape(-1)
        self._check(self._lib.ruf_group_set_model(self._h, tri.ctypes.data, part.ctypes.data, tri.shape[0], int(n_parts)))
        self.n_parts = int(n_parts)
        return int(self._lib.ruf_group_broadcast_bytes(self._h))

    def filter_batch_host(self, depth, proj, views, part_models, max_diff, replace_value, out=None, mask=None,
                          frames_per_chunk=0):
        depth = np.asarray(depth)
        enc = ENC_U16_MM if depth.dtype == np.uint16 else ENC_F32_M
        if depth.ndim != 3 or depth.shape[1:] != (self.height, self.width) or not depth.flags.c_contiguous:
            raise ValueError("depth must be a C-contiguous (n, H, W) array")
        n = depth.shape[0]
        out = np.empty_like(depth) if out is None else out
        mask = np.empty(depth.shape, np.uint8) if mask is None else mask
        v, pm, pr = _as_f64(views, 16 * n), _as_f64(part_models, 16 * n * self.n_parts), _as_f64(proj, 16)
        self._check(self._lib.ruf_group_filter_batch_host(self._h, n, depth.ctypes.data, enc, pr.ctypes.data, v.ctypes.data,
                                                          pm.ctypes.data if pm.size else None, max_diff, replace_value,
                                                          out.ctypes.data, mask.ctypes.data, int(frames_per_chunk)))
        return out, mask


def host_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    rc = load().ruf_host_alloc(C.byref(p), nbytes)
    if rc != RUF_OK:
        raise RufError(rc, load().ruf_last_error(None).decode())
    return p.value


def host_free(ptr: int):
    load().ruf_host_free(C.c_void_p(ptr))
