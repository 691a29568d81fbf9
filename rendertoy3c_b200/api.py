"""Host-side mirror of rendertoy3o's device-scene operators over the librt3.so C ABI.

Class / method names follow the reference's operators so that a caller of the reference reads
the same code here:
    Context                       <- OptixContext (src/cuda/optix_context.h:231-271) + CUDAScene (src/cuda/cuda_scene.h:124-183)
    Context.mesh_create           <- CUDAMesh ctor (src/cuda/cuda_mesh.h:33-155)
    Context.texture_create        <- CUDATexture<uchar4> ctor (src/cuda/cuda_texture.h:46-75)
    Context.append_instance / append_animated_instance / accel_build
                                  <- CUDAAccel (src/cuda/cuda_accel.h:38-150)
    Context.set_hitgroup          <- CUDAScene::create_sbt (src/cuda/cuda_scene.h:54-88)
    Context.set_lights            <- buildLightSampler (src/wavefront.cpp:257-275)
    Context.launch_subframe       <- launchSubframe (src/wavefront.cpp:203-222)
Errors raise Rt3Error (reference: rendertoy3o::Exception, src/util/exception.h:28-60).
There is no CPU fallback: loading fails loudly if librt3.so is missing, and context creation fails
without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

from ._abi import HIT_DTYPE, LIGHT_BYTES, RAY_DTYPE, RenderSettings, Stats, bptr, fptr, iptr

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librt3.so")

RT3_SYMBOLS = [
    "rt3_context_create", "rt3_context_destroy", "rt3_sync", "rt3_last_error", "rt3_get_stream", "rt3_get_stats", "rt3_reset_stats", "rt3_get_debug_counters",
    "rt3_set_option", "rt3_mesh_create", "rt3_mesh_set_colors", "rt3_spheres_create", "rt3_curves_create", "rt3_texture_create",
    "rt3_accel_append_instance", "rt3_accel_append_animated_instance", "rt3_accel_build", "rt3_scene_set_hitgroup",
    "rt3_scene_set_lights", "rt3_light_make", "rt3_camera_uvw", "rt3_launch_subframe", "rt3_trace", "rt3_trace_device", "rt3_get_local_geometry", "rt3_scene_set_texture_transform",
    "rt3_download_accum", "rt3_download_frame", "rt3_download_frame_async", "rt3_accum_device_ptr", "rt3_clear_accum", "rt3_finalize_accum",
    "rt3_allreduce_accum",
]


class Rt3Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rt3 error %d: %s" % (code, msg))
        self.code = code


_libs = {}


def load_library(path=None):
    """dlopen librt3.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    p = path or os.environ.get("RT3_LIB") or LIB_PATH  # RT3_LIB: tuning builds of the same sources (tools/)
    if p not in _libs:
        if not os.path.exists(p):
            raise Rt3Error(-3, "librt3.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback" % p)
        L = C.CDLL(p)
        L.rt3_last_error.restype = C.c_char_p
        L.rt3_context_destroy.restype = None
        _libs[p] = L
    return _libs[p]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def camera_uvw(eye, lookat, up, fovy, aspect, L=None):
    """sutil::Camera::UVWFrame (sutil/Camera.cpp:34-45) through the library's host helper."""
    L = L or load_library()
    e, l, u = _f32(eye), _f32(lookat), _f32(up)
    U, V, W = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    rc = L.rt3_camera_uvw(fptr(e), fptr(l), fptr(u), C.c_float(fovy), C.c_float(aspect), fptr(U), fptr(V), fptr(W))
    if rc != 0:
        raise Rt3Error(rc, L.rt3_last_error().decode())
    return U, V, W


def make_settings(desc, uvw, subframe_index=0, samples_per_launch=8, accum_mode=0, max_depth=None, width=None, height=None, mode=0, miss=0.01):
    """RenderSettings for a SceneDesc (src/shader/shader_data.h:100-112 + handleCameraUpdate src/wavefront.cpp:167-176)."""
    rs = RenderSettings()
    rs.width = width or desc.width
    rs.height = height or desc.height
    rs.samples_per_launch = samples_per_launch
    rs.subframe_index = subframe_index
    U, V, W = uvw
    for i in range(3):
        rs.eye[i] = float(np.float32(desc.camera.eye[i]))
        rs.U[i], rs.V[i], rs.W[i] = float(U[i]), float(V[i]), float(W[i])
        rs.miss_color[i] = miss
    rs.max_depth = desc.max_depth if max_depth is None else max_depth
    rs.mode = mode  # 0 = REFERENCE_FAITHFUL, 1 = CORRECTED, 2 = CORRECTED + power light sampler
    rs.accum_mode = accum_mode
    return rs


class Context:
    """One GPU's device scene + renderer.  Move-only in spirit, like the reference's RAII owners."""

    def __init__(self, device=0, lib_path=None):
        self.L = load_library(lib_path)
        self.ctx = C.c_void_p()
        self._chk(self.L.rt3_context_create(C.c_int(device), C.byref(self.ctx)))
        self.width = self.height = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.L.rt3_context_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise Rt3Error(rc, self.L.rt3_last_error().decode())

    # ---- geometry
    def mesh_create(self, verts, idx, normals, uvs):
        v = _f32(verts)  # verts [nv,3] or [keys,nv,3] (vertex-key motion)
        n = _f32(normals) if normals is not None else None   # None: the SDK's fallbacks (N = geometric normal / UV = barycentrics)
        t = _f32(uvs) if uvs is not None else None
        i = np.ascontiguousarray(idx, dtype=np.int32)
        h = C.c_uint64()
        keys, nv = (v.shape[0], v.shape[1]) if v.ndim == 3 else (1, len(v))
        self._chk(self.L.rt3_mesh_create(self.ctx, fptr(v), C.c_int(keys), C.c_int(nv), iptr(i), C.c_int(len(i)), fptr(n) if n is not None else None,
                                         fptr(t) if t is not None else None, C.byref(h)))
        return h.value

    def mesh_set_colors(self, blas, rgba):
        c = _f32(rgba)
        self._chk(self.L.rt3_mesh_set_colors(self.ctx, C.c_uint64(blas), fptr(c)))

    def spheres_create(self, cr):
        c = _f32(cr)
        h = C.c_uint64()
        self._chk(self.L.rt3_spheres_create(self.ctx, fptr(c), C.c_int(len(c)), C.byref(h)))
        return h.value

    def curves_create(self, degree, cp, seg):
        c = _f32(cp)
        s = np.ascontiguousarray(seg, dtype=np.int32)
        h = C.c_uint64()
        self._chk(self.L.rt3_curves_create(self.ctx, C.c_int(degree), fptr(c), C.c_int(len(c)), iptr(s), C.c_int(len(s)), C.byref(h)))
        return h.value

    def texture_create(self, rgba, address=0, filt=0):
        r = np.ascontiguousarray(rgba, dtype=np.uint8)
        tid = C.c_int()
        self._chk(self.L.rt3_texture_create(self.ctx, bptr(r), C.c_int(r.shape[1]), C.c_int(r.shape[0]), C.c_int(address), C.c_int(filt), C.byref(tid)))
        return tid.value

    # ---- instances
    def append_instance(self, blas, xform):
        x = _f32(xform)
        iid = C.c_int()
        self._chk(self.L.rt3_accel_append_instance(self.ctx, C.c_uint64(blas), fptr(x), C.byref(iid)))
        return iid.value

    def append_animated_instance(self, blas, keys, t_begin, t_end, static_xform):
        k, x = _f32(keys), _f32(static_xform)
        iid = C.c_int()
        self._chk(self.L.rt3_accel_append_animated_instance(self.ctx, C.c_uint64(blas), fptr(k), C.c_int(len(k)), C.c_float(t_begin),
                                                            C.c_float(t_end), fptr(x), C.byref(iid)))
        return iid.value

    def accel_build(self):
        self._chk(self.L.rt3_accel_build(self.ctx))

    # ---- shading records
    def set_hitgroup(self, iid, emission, diffuse, tex):
        e, d = _f32(emission), _f32(diffuse)
        self._chk(self.L.rt3_scene_set_hitgroup(self.ctx, C.c_int(iid), fptr(e), fptr(d), C.c_int(tex)))

    def light_make(self, e, v0, v1, v2):
        buf = C.create_string_buffer(LIGHT_BYTES)
        a, b, c, d = _f32(e), _f32(v0), _f32(v1), _f32(v2)
        self._chk(self.L.rt3_light_make(fptr(a), fptr(b), fptr(c), fptr(d), buf))
        return buf.raw

    def set_lights(self, blob, n):
        self._chk(self.L.rt3_scene_set_lights(self.ctx, C.c_char_p(blob), C.c_int(n)))

    def camera_uvw(self, eye, lookat, up, fovy, aspect):
        return camera_uvw(eye, lookat, up, fovy, aspect, L=self.L)

    def clear_accum(self):
        self._chk(self.L.rt3_clear_accum(self.ctx))

    # ---- hot path
    def launch_subframe(self, settings: RenderSettings):
        self.width, self.height = settings.width, settings.height
        self._chk(self.L.rt3_launch_subframe(self.ctx, C.byref(settings)))

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self._chk(self.L.rt3_trace(self.ctx, rays.ctypes.data_as(C.c_void_p), C.c_int(len(rays)), C.c_int(1 if any_hit else 0),
                                   hits.ctypes.data_as(C.c_void_p)))
        return hits

    def set_texture_transform(self, iid, scale, rotation, offset):
        """texcoord transform of the SDK's sampleTexture (cuda/LocalShading.h:37-54); rotation = (sin, cos)"""
        s, r, o = (np.ascontiguousarray(x, dtype=np.float32) for x in (scale, rotation, offset))
        self._chk(self.L.rt3_scene_set_texture_transform(self.ctx, C.c_int(iid), fptr(s), fptr(r), fptr(o)))

    def get_local_geometry(self, rays, hits):
        """getLocalGeometry (cuda/LocalGeometry.h:61-175) for rays and the hit records trace() returned for them"""
        from ._abi import LOCAL_GEOMETRY_DTYPE
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        out = np.zeros(len(rays), dtype=LOCAL_GEOMETRY_DTYPE)
        self._chk(self.L.rt3_get_local_geometry(self.ctx, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p), C.c_int(len(rays)),
                                                out.ctypes.data_as(C.c_void_p)))
        return out

    def trace_device(self, d_rays_ptr, n, any_hit, d_hits_ptr):
        self._chk(self.L.rt3_trace_device(self.ctx, C.c_void_p(d_rays_ptr), C.c_int(n), C.c_int(1 if any_hit else 0), C.c_void_p(d_hits_ptr)))

    # ---- results
    def download_accum(self):
        out = np.zeros((self.height, self.width, 4), dtype=np.float32)
        self._chk(self.L.rt3_download_accum(self.ctx, fptr(out)))
        return out

    def download_frame(self):
        out = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        self._chk(self.L.rt3_download_frame(self.ctx, bptr(out)))
        return out

    def accum_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._chk(self.L.rt3_accum_device_ptr(self.ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def finalize_accum(self, total_subframes):
        self._chk(self.L.rt3_finalize_accum(self.ctx, C.c_uint32(total_subframes)))

    def sync(self):
        self._chk(self.L.rt3_sync(self.ctx))

    def stream(self):
        p = C.c_void_p()
        self._chk(self.L.rt3_get_stream(self.ctx, C.byref(p)))
        return p.value or 0

    def download_frame_into(self, host_ptr):
        self._chk(self.L.rt3_download_frame(self.ctx, C.c_void_p(host_ptr)))

    def download_frame_async_into(self, pinned_host_ptr):
        """rt3_download_frame_async: returns at once; sync() completes the copy"""
        self._chk(self.L.rt3_download_frame_async(self.ctx, C.c_void_p(pinned_host_ptr)))

    def stats(self):
        st = Stats()
        self._chk(self.L.rt3_get_stats(self.ctx, C.byref(st)))
        return st.as_dict()

    def reset_stats(self):
        self._chk(self.L.rt3_reset_stats(self.ctx))

    def debug_counters(self):
        out = (C.c_uint32 * 16)()
        self._chk(self.L.rt3_get_debug_counters(self.ctx, out))
        return list(out)

    def set_option(self, key, value):
        self._chk(self.L.rt3_set_option(self.ctx, key.encode(), C.c_int(int(value))))


def camera_rays(desc, uvw, width, height, rng=None, n=None):
    """Pinhole rays through pixel centres of a (width x height) film for rt3_trace tests/benches
    (same direction formula as raygen.cu:32-39 with jitter 0.5), optionally a random subset of n."""
    U, V, W = uvw
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    if n is not None:
        sel = rng.choice(width * height, size=n, replace=False)
        xs, ys = xs.reshape(-1)[sel], ys.reshape(-1)[sel]
    dx = (2.0 * ((xs.reshape(-1) + 0.5) / np.float32(width)) - 1.0).astype(np.float32)
    dy = (2.0 * ((ys.reshape(-1) + 0.5) / np.float32(height)) - 1.0).astype(np.float32)
    d = dx[:, None] * U[None, :] + dy[:, None] * V[None, :] + W[None, :]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros(len(d), dtype=RAY_DTYPE)
    rays["o"] = np.asarray(desc.camera.eye, dtype=np.float32)[None, :]
    rays["d"] = d
    rays["tmin"] = 0.01
    rays["tmax"] = 1e16
    rays["time"] = 0.0
    return rays
