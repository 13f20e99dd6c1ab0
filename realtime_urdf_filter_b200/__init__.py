"""B200-native (sm_100a) URDF depth-image self-filter: the per-frame hot path of
blodow/realtime_urdf_filter as hand-written CUDA behind a C ABI (include/ruf_b200.h).

Python here is only the ctypes binding (`Context`), the build helper and synthetic-scene
generation for tests/bench; the product is libruf_b200.so.  There is no CPU fallback.
"""
from ._lib import (Context, Group, RufError, ENC_F32_M, ENC_U16_MM, MASK_BYTES, MASK_BITS, RUF_OK, RUF_ERR_CUDA, RUF_ERR_INVALID,
                   RUF_ERR_NO_MODEL, RUF_ERR_OVERFLOW, load, lib_path, projection_matrix, lookat, view_matrix,
                   part_model, box_triangles, cube_triangles, sphere_triangles, cylinder_triangles,
                   host_alloc, host_free)

__all__ = ["Context", "Group", "RufError", "ENC_F32_M", "ENC_U16_MM", "MASK_BYTES", "MASK_BITS", "load", "lib_path", "projection_matrix",
           "lookat", "view_matrix", "part_model", "box_triangles", "cube_triangles", "sphere_triangles",
           "cylinder_triangles", "host_alloc", "host_free"]
