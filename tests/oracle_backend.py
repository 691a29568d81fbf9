"""ctypes binding of the CPU oracle (oracle/_ref/librt3o.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from rendertoy3c_b200._abi import HIT_DTYPE, LIGHT_BYTES, RAY_DTYPE, RenderSettings, Stats, bptr, fptr, iptr

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(_ROOT, "oracle", "_ref", "librt3o.so")


def build_oracle():
    """(Re)build librt3o.so from oracle/*.cpp when missing or stale (g++ only, a few seconds)."""
    srcs = [os.path.join(_ROOT, "oracle", f) for f in ("rt3o.cpp", "rt3o.h", "rt3o_math.hpp", "rt3o_prims.hpp", "rt3o_curve.hpp")]
    srcs.append(os.path.join(_ROOT, "include", "rt3.h"))
    if os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs):
        return _LIB
    subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle"), "_ref/librt3o.so"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        L = _lib
        L.rt3o_scene_create.restype = C.c_void_p
        L.rt3o_last_error.restype = C.c_char_p
        L.rt3o_kat_rnd.restype = C.c_float
        L.rt3o_kat_tea4.restype = C.c_uint32
        L.rt3o_kat_tea4.argtypes = [C.c_uint32, C.c_uint32]
        for name in ("rt3o_scene_destroy", "rt3o_mesh_create", "rt3o_mesh_set_colors", "rt3o_spheres_create", "rt3o_curves_create", "rt3o_texture_create",
                     "rt3o_accel_append_instance", "rt3o_accel_append_animated_instance", "rt3o_accel_build",
                     "rt3o_scene_set_hitgroup", "rt3o_scene_set_texture_transform", "rt3o_scene_set_lights", "rt3o_trace", "rt3o_get_local_geometry", "rt3o_launch_subframe",
                     "rt3o_download_accum", "rt3o_download_frame", "rt3o_get_stats", "rt3o_reset_stats", "rt3o_kat_fetch_texture", "rt3o_kat_sample_texture"):
            getattr(L, name).argtypes = None
    return _lib


class OracleError(RuntimeError):
    pass


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleScene:
    """Same operator surface as rendertoy3c_b200.api.Context, backed by the scalar CPU oracle."""

    def __init__(self, nthreads=0):
        self.L = lib()
        self.s = C.c_void_p(self.L.rt3o_scene_create())
        self.nthreads = nthreads
        self.width = self.height = 0

    def close(self):
        if self.s:
            self.L.rt3o_scene_destroy(self.s)
            self.s = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc < 0:
            raise OracleError(self.L.rt3o_last_error().decode())
        return rc

    def mesh_create(self, verts, idx, normals, uvs):
        v = _f32(verts)
        n = _f32(normals) if normals is not None else None
        t = _f32(uvs) if uvs is not None else None
        i = np.ascontiguousarray(idx, dtype=np.int32)
        keys, nv = (v.shape[0], v.shape[1]) if v.ndim == 3 else (1, len(v))
        return self._chk(self.L.rt3o_mesh_create(self.s, fptr(v), C.c_int(keys), C.c_int(nv), iptr(i), C.c_int(len(i)), fptr(n) if n is not None else None,
                                                 fptr(t) if t is not None else None))

    def mesh_set_colors(self, blas, rgba):
        c = _f32(rgba)
        self._chk(self.L.rt3o_mesh_set_colors(self.s, C.c_int(blas), fptr(c)))

    def spheres_create(self, cr):
        c = _f32(cr)
        return self._chk(self.L.rt3o_spheres_create(self.s, fptr(c), C.c_int(len(c))))

    def curves_create(self, degree, cp, seg):
        c = _f32(cp)
        s = np.ascontiguousarray(seg, dtype=np.int32)
        return self._chk(self.L.rt3o_curves_create(self.s, C.c_int(degree), fptr(c), C.c_int(len(c)), iptr(s), C.c_int(len(s))))

    def texture_create(self, rgba, address=0, filt=0):
        r = np.ascontiguousarray(rgba, dtype=np.uint8)
        return self._chk(self.L.rt3o_texture_create(self.s, bptr(r), C.c_int(r.shape[1]), C.c_int(r.shape[0]), C.c_int(address), C.c_int(filt)))

    def fetch_texture(self, tex, u, v):
        out = np.zeros(3, dtype=np.float32)
        self._chk(self.L.rt3o_kat_fetch_texture(self.s, C.c_int(tex), C.c_float(u), C.c_float(v), fptr(out)))
        return out

    def sample_texture(self, iid, u, v):
        out = np.zeros(3, dtype=np.float32)
        self._chk(self.L.rt3o_kat_sample_texture(self.s, C.c_int(iid), C.c_float(u), C.c_float(v), fptr(out)))
        return out

    def append_instance(self, blas, xform):
        x = _f32(xform)
        return self._chk(self.L.rt3o_accel_append_instance(self.s, C.c_int(blas), fptr(x)))

    def append_animated_instance(self, blas, keys, t_begin, t_end, static_xform):
        k, x = _f32(keys), _f32(static_xform)
        return self._chk(self.L.rt3o_accel_append_animated_instance(self.s, C.c_int(blas), fptr(k), C.c_int(len(k)), C.c_float(t_begin), C.c_float(t_end), fptr(x)))

    def set_option(self, key, value):
        """the build options that change results: "flatten" (process-wide in the oracle, read by accel_build); the kernels'
        tuning switches ("split", "tlas_sah", ...) leave results alone and mean nothing here"""
        if key == "flatten" or (key == "merge_identity" and not value):   # without the merged BLAS nothing is flattened
            self.L.rt3o_set_flatten(C.c_int(int(value)))

    def accel_build(self):
        self._chk(self.L.rt3o_accel_build(self.s))

    def set_hitgroup(self, iid, emission, diffuse, tex):
        e, d = _f32(emission), _f32(diffuse)
        self._chk(self.L.rt3o_scene_set_hitgroup(self.s, C.c_int(iid), fptr(e), fptr(d), C.c_int(tex)))

    def set_texture_transform(self, iid, scale, rotation, offset):
        s, r, o = _f32(scale), _f32(rotation), _f32(offset)
        self._chk(self.L.rt3o_scene_set_texture_transform(self.s, C.c_int(iid), fptr(s), fptr(r), fptr(o)))

    def light_make(self, e, v0, v1, v2):
        buf = C.create_string_buffer(LIGHT_BYTES)
        a, b, c, d = _f32(e), _f32(v0), _f32(v1), _f32(v2)
        self.L.rt3o_kat_light_make(fptr(a), fptr(b), fptr(c), fptr(d), buf)
        return buf.raw

    def set_lights(self, blob, n):
        self._chk(self.L.rt3o_scene_set_lights(self.s, C.c_char_p(blob), C.c_int(n)))

    def camera_uvw(self, eye, lookat, up, fovy, aspect):
        out = np.zeros(9, dtype=np.float32)
        a, b, c = _f32(eye), _f32(lookat), _f32(up)
        self.L.rt3o_kat_camera_uvw(fptr(a), fptr(b), fptr(c), C.c_float(fovy), C.c_float(aspect), fptr(out))
        return out[0:3].copy(), out[3:6].copy(), out[6:9].copy()

    def trace(self, rays, any_hit=False, accel=1):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self._chk(self.L.rt3o_trace(self.s, rays.ctypes.data_as(C.c_void_p), C.c_int(len(rays)), C.c_int(1 if any_hit else 0),
                                    hits.ctypes.data_as(C.c_void_p), C.c_int(accel), C.c_int(self.nthreads)))
        return hits

    def get_local_geometry(self, rays, hits):
        from rendertoy3c_b200._abi import LOCAL_GEOMETRY_DTYPE
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        out = np.zeros(len(rays), dtype=LOCAL_GEOMETRY_DTYPE)
        self._chk(self.L.rt3o_get_local_geometry(self.s, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p), C.c_int(len(rays)),
                                                 out.ctypes.data_as(C.c_void_p)))
        return out

    def launch_subframe(self, settings: RenderSettings):
        self.width, self.height = settings.width, settings.height
        self._chk(self.L.rt3o_launch_subframe(self.s, C.byref(settings), C.c_int(self.nthreads)))

    def download_accum(self):
        out = np.zeros((self.height, self.width, 4), dtype=np.float32)
        self._chk(self.L.rt3o_download_accum(self.s, fptr(out)))
        return out

    def download_frame(self):
        out = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        self._chk(self.L.rt3o_download_frame(self.s, bptr(out)))
        return out

    def stats(self):
        st = Stats()
        self._chk(self.L.rt3o_get_stats(self.s, C.byref(st)))
        return st.as_dict()

    def reset_stats(self):
        self.L.rt3o_reset_stats(self.s)

    def sync(self):
        pass
