"""ctypes view of include/rt3.h: the structs that cross the C ABI and their numpy dtypes.

Layouts are asserted against the C side by tests (sizeof checks through the loaded library).
"""
import ctypes as C

import numpy as np


class RenderSettings(C.Structure):
    """rt3_render_settings — replaces RenderSettings (reference src/shader/shader_data.h:71-114)."""

    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("samples_per_launch", C.c_uint32), ("subframe_index", C.c_uint32),
        ("eye", C.c_float * 3), ("U", C.c_float * 3), ("V", C.c_float * 3), ("W", C.c_float * 3),
        ("max_depth", C.c_int32), ("mode", C.c_int32),
        ("miss_color", C.c_float * 3),
        ("accum_mode", C.c_int32),
    ]


class Stats(C.Structure):
    """rt3_stats."""

    _fields_ = [
        ("rays_primary", C.c_uint64), ("rays_bounce", C.c_uint64), ("rays_shadow", C.c_uint64),
        ("samples", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("ms_generate", C.c_float), ("ms_extend", C.c_float), ("ms_shade", C.c_float),
        ("ms_connect", C.c_float), ("ms_resolve", C.c_float), ("ms_total", C.c_float),
        ("max_stack_depth", C.c_uint32), ("error_flags", C.c_uint32), ("flattened_instances", C.c_uint32), ("traversal_passes", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# rt3_ray (48 B) / rt3_hit (32 B)
RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmin", "<f4"), ("d", "<f4", 3), ("tmax", "<f4"), ("time", "<f4"), ("pad", "<f4", 3)])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<i4"), ("inst", "<i4"), ("pad", "<i4", 3)])
LOCAL_GEOMETRY_DTYPE = np.dtype([("P", "<f4", 3), ("N", "<f4", 3), ("Ng", "<f4", 3), ("UV", "<f4", 2), ("dndu", "<f4", 3), ("dndv", "<f4", 3),
                                 ("dpdu", "<f4", 3), ("dpdv", "<f4", 3), ("color", "<f4", 4)])   # rt3_local_geometry, cuda/LocalGeometry.h:40-58
assert RAY_DTYPE.itemsize == 48 and HIT_DTYPE.itemsize == 32 and LOCAL_GEOMETRY_DTYPE.itemsize == 108
LIGHT_BYTES = 68  # rendertoy3o::Light (reference src/light.h:13-22)


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def bptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))
