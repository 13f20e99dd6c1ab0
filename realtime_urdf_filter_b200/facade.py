"""Python driver for the C++ `RealtimeURDFFilter` facade (host/urdf_filter.h) through the flat hooks
of host/facade_c.cpp.  It plays the role of the ROS graph in tests: parameter server, TF broadcaster,
image publisher and subscriber of the outputs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _sig(lib):
    vp, cp, dp = C.c_void_p, C.c_char_p, C.POINTER(C.c_double)
    lib.ruf_facade_new.restype = vp
    lib.ruf_facade_delete.argtypes = [vp]
    lib.ruf_facade_param_str.argtypes = [vp, cp, cp]
    lib.ruf_facade_param_double.argtypes = [vp, cp, C.c_double]
    lib.ruf_facade_param_bool.argtypes = [vp, cp, C.c_int]
    lib.ruf_facade_camera_offset.argtypes = [vp, dp, dp]
    lib.ruf_facade_add_model.argtypes = [vp, cp, cp, cp, C.c_double, cp]
    lib.ruf_facade_construct.argtypes = [vp]
    lib.ruf_facade_add_resource_root.argtypes = [vp, cp]
    lib.ruf_facade_set_tf.argtypes = [vp, cp, dp, dp]
    lib.ruf_facade_erase_tf.argtypes = [vp, cp]
    lib.ruf_facade_subscribers.argtypes = [vp, C.c_int, C.c_int]
    lib.ruf_facade_callback.argtypes = [vp, cp, C.c_int, C.c_int, C.c_int, vp, dp, C.c_double]
    lib.ruf_facade_filter.argtypes = [vp, vp, dp, C.c_int, C.c_int, vp]
    lib.ruf_facade_projection.argtypes = [vp, C.c_int, C.c_int, dp, dp]
    lib.ruf_facade_published.argtypes = [vp, C.c_int]
    lib.ruf_facade_published.restype = C.c_long
    lib.ruf_facade_last_image.argtypes = [vp, C.c_int, vp, C.c_long, cp]
    lib.ruf_facade_last_image.restype = C.c_long
    lib.ruf_facade_counts.argtypes = [vp, C.c_int]
    lib.ruf_facade_counts.restype = C.c_long
    lib.ruf_facade_log.argtypes = [vp]
    lib.ruf_facade_log.restype = cp
    lib.ruf_facade_clear_log.argtypes = [vp]
    lib.ruf_facade_error.argtypes = [vp]
    lib.ruf_facade_error.restype = cp
    lib.ruf_facade_get_double.argtypes = [vp, cp]
    lib.ruf_facade_get_double.restype = C.c_double
    lib.ruf_facade_parse_urdf.argtypes = [cp, cp, C.c_double, cp, cp, vp, vp, C.c_long, C.POINTER(C.c_long), vp]
    lib.ruf_facade_parse_urdf.restype = C.c_long
    lib.ruf_facade_tracker_depth_to_buffer.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.ruf_facade_tracker_projection.argtypes = [C.c_int, C.c_int, vp]
    lib.ruf_facade_tracker_masked_depth_to_mm.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.ruf_facade_last_parse_error.restype = cp
    return lib


_ready = None


def lib():
    global _ready
    if _ready is None:
        _ready = _sig(_lib.load())
    return _ready


def _d(a, n):
    a = np.ascontiguousarray(a, np.float64).reshape(-1)
    assert a.size == n
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def tracker_depth_to_buffer(depth_mm: np.ndarray) -> np.ndarray:
    """OpenNITrackerLoopback::runOnce input conversion (mirror in x, mm -> m), urdf_filtered_tracker.cpp:201-207."""
    d = np.ascontiguousarray(depth_mm, np.uint16)
    out = np.empty(d.shape, np.float32)
    lib().ruf_facade_tracker_depth_to_buffer(C.c_void_p(d.ctypes.data), d.shape[1], d.shape[0], C.c_void_p(out.ctypes.data))
    return out


def tracker_projection(xres: int, yres: int) -> np.ndarray:
    g = np.zeros(16)
    lib().ruf_facade_tracker_projection(xres, yres, C.c_void_p(g.ctypes.data))
    return g


def tracker_masked_depth_to_mm(masked: np.ndarray) -> np.ndarray:
    """...:243-249: truncating m -> mm of getMaskedDepth(), not mirrored back."""
    m = np.ascontiguousarray(masked, np.float32)
    out = np.empty(m.shape, np.uint16)
    lib().ruf_facade_tracker_masked_depth_to_mm(C.c_void_p(m.ctypes.data), m.shape[1], m.shape[0], C.c_void_p(out.ctypes.data))
    return out


def parse_urdf(xml: str, geometry_type: str = "", scale: float = 1.0, ignore=(), resource_root: str = ""):
    """URDF text -> (tri[T,9], tri_part[T], part_models[P,16] with identity TF).  CPU only."""
    L = lib()
    n_parts = C.c_long(0)
    n = L.ruf_facade_parse_urdf(xml.encode(), geometry_type.encode(), scale, ",".join(ignore).encode(),
                                resource_root.encode(), None, None, 0, C.byref(n_parts), None)
    if n < 0:
        raise ValueError("URDF failed to parse: " + L.ruf_facade_last_parse_error().decode())
    tri = np.zeros((n, 9), np.float32)
    part = np.zeros(n, np.uint32)
    pm = np.zeros((n_parts.value, 16))
    L.ruf_facade_parse_urdf(xml.encode(), geometry_type.encode(), scale, ",".join(ignore).encode(),
                            resource_root.encode(), tri.ctypes.data, part.ctypes.data, n, C.byref(n_parts),
                            pm.ctypes.data)
    return tri, part, pm


class FilterNode:
    """A RealtimeURDFFilter instance plus the shimmed ROS graph around it."""

    def __init__(self, params: dict, models: list[dict], camera_offset=None):
        self.L = lib()
        self.h = self.L.ruf_facade_new()
        for k, v in params.items():
            if isinstance(v, bool):
                self.L.ruf_facade_param_bool(self.h, k.encode(), int(v))
            elif isinstance(v, (int, float)):
                self.L.ruf_facade_param_double(self.h, k.encode(), float(v))
            else:
                self.L.ruf_facade_param_str(self.h, k.encode(), str(v).encode())
        if camera_offset is not None:
            t, tp = _d(camera_offset[0], 3)
            q, qp = _d(camera_offset[1], 4)
            self.L.ruf_facade_camera_offset(self.h, tp, qp)
        for m in models:
            self.L.ruf_facade_add_model(self.h, m["model"].encode(), m.get("tf_prefix", "").encode(),
                                        m.get("geometry_type", "").encode(), float(m.get("scale", 1.0)),
                                        ",".join(m.get("ignore", [])).encode())
        self.L.ruf_facade_construct(self.h)

    def close(self):
        if self.h:
            self.L.ruf_facade_delete(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def add_resource_root(self, d):
        self.L.ruf_facade_add_resource_root(self.h, d.encode())

    def set_tf(self, frame, q, t):
        _, qp = _d(q, 4)
        _, tp = _d(t, 3)
        self.L.ruf_facade_set_tf(self.h, frame.encode(), qp, tp)

    def erase_tf(self, frame):
        self.L.ruf_facade_erase_tf(self.h, frame.encode())

    def subscribers(self, depth=1, mask=1):
        self.L.ruf_facade_subscribers(self.h, depth, mask)

    def callback(self, depth: np.ndarray, P, stamp=0.0, encoding=None):
        d = np.ascontiguousarray(depth)
        enc = encoding or ("16UC1" if d.dtype == np.uint16 else "32FC1")
        H, W = d.shape
        _, pp = _d(P, 12)
        rc = self.L.ruf_facade_callback(self.h, enc.encode(), W, H, W * d.itemsize, d.ctypes.data, pp, stamp)
        if rc != 0:
            raise RuntimeError(self.L.ruf_facade_error(self.h).decode())

    def filter(self, depth_f32: np.ndarray, glTf):
        d = np.ascontiguousarray(depth_f32, np.float32)
        H, W = d.shape
        out = np.zeros_like(d)
        _, gp = _d(glTf, 16)
        rc = self.L.ruf_facade_filter(self.h, d.ctypes.data, gp, W, H, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(self.L.ruf_facade_error(self.h).decode())
        return out

    def projection(self, W, H, P):
        out = np.zeros(16)
        _, pp = _d(P, 12)
        self.L.ruf_facade_projection(self.h, W, H, pp, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def published(self):
        return self.L.ruf_facade_published(self.h, 0), self.L.ruf_facade_published(self.h, 1)

    def last_image(self, which, shape, dtype):
        out = np.zeros(shape, dtype)
        enc = C.create_string_buffer(16)
        n = self.L.ruf_facade_last_image(self.h, which, out.ctypes.data, out.nbytes, enc)
        return (out if n == out.nbytes else None), enc.value.decode()

    def counts(self):
        names = ["renderers", "parts", "triangles", "tf_lookups", "frames", "renderables", "mesh_errors"]
        return {n: self.L.ruf_facade_counts(self.h, i) for i, n in enumerate(names)}

    def log(self, clear=False):
        s = self.L.ruf_facade_log(self.h).decode()
        if clear:
            self.L.ruf_facade_clear_log(self.h)
        return s

    def get(self, name):
        return self.L.ruf_facade_get_double(self.h, name.encode())
