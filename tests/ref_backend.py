"""ctypes binding of oracle/_ref/librt3ref.so: the REFERENCE'S OWN device programs (src/shader/raygen.cu,
closehit_radiance.cu, miss.cu, test.cu) compiled where they lie with a functional host OptiX stand-in
(oracle/ref_shim/optix.h).  TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the build
container): it produces the reference-held images committed under tests/golden/ref_images.npz.

RefShaderScene has the operator surface scenes.replay() drives; traversal and texel fetches go to an OracleScene it
owns (the reference has no source for either), everything else is the reference's code."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_backend import OracleScene, lib as oracle_lib
from rendertoy3c_b200._abi import RenderSettings, bptr, fptr, iptr
from rendertoy3c_b200.scenes import IDENTITY

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(_ROOT, "oracle", "_ref", "librt3ref.so")


def available():
    return os.path.isdir("/root/reference/src/shader")


_lib = None


def lib():
    global _lib
    if _lib is None:
        oracle_lib()   # builds librt3o.so first
        subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle"), "ref_shaders"], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
        _lib.rt3ref_create.restype = C.c_void_p
        _lib.rt3ref_create.argtypes = [C.c_void_p]
        _lib.rt3ref_destroy.argtypes = [C.c_void_p]
    return _lib


class RefShaderScene:
    """Reference scope only: triangle meshes under identity instances (all src/ ever creates, cuda_scene.h:141-146)."""

    def __init__(self, nthreads=0):
        self.o = OracleScene(nthreads)
        self.L = lib()
        self.r = C.c_void_p(self.L.rt3ref_create(self.o.s))
        self.nthreads = nthreads
        self.meshes, self.inst, self.em, self.df, self.tex = [], [], [], [], []
        self.lights = (b"", 0)
        self.accum = self.frame = None

    def close(self):
        if self.r:
            self.L.rt3ref_destroy(self.r)
            self.r = None
        self.o.close()

    def mesh_create(self, verts, idx, normals, uvs):
        v = np.ascontiguousarray(verts, dtype=np.float32)
        assert v.ndim == 2, "vertex keys are outside what this harness pins"
        self.meshes.append((v, np.ascontiguousarray(idx, dtype=np.int32), np.ascontiguousarray(normals, dtype=np.float32), np.ascontiguousarray(uvs, dtype=np.float32)))
        return self.o.mesh_create(verts, idx, normals, uvs)

    def texture_create(self, rgba, address=0, filt=0):
        return self.o.texture_create(rgba, address, filt)

    def append_instance(self, blas, xform):
        assert np.array_equal(np.asarray(xform, dtype=np.float32), IDENTITY), "the reference's programs assume identity instances (Q13)"
        iid = self.o.append_instance(blas, xform)
        assert iid == len(self.inst)
        self.inst.append(blas)
        v, i, n, t = self.meshes[blas]
        self.L.rt3ref_set_mesh(self.r, C.c_int(iid), fptr(v), C.c_int(len(v)), iptr(i), C.c_int(len(i)), fptr(n), fptr(t))
        return iid

    def set_hitgroup(self, iid, emission, diffuse, tex):
        assert iid == len(self.em)
        self.em.append(np.asarray(emission, dtype=np.float32)); self.df.append(np.asarray(diffuse, dtype=np.float32)); self.tex.append(int(tex))
        self.o.set_hitgroup(iid, emission, diffuse, tex)

    def light_make(self, e, v0, v1, v2):
        return self.o.light_make(e, v0, v1, v2)

    def set_lights(self, blob, n):
        self.lights = (blob, n)
        self.o.set_lights(blob, n)

    def accel_build(self):
        self.o.accel_build()

    def camera_uvw(self, *a):
        return self.o.camera_uvw(*a)

    def launch_subframe(self, rs: RenderSettings):
        if self.accum is None or self.accum.shape[:2] != (rs.height, rs.width):
            self.accum = np.zeros((rs.height, rs.width, 4), dtype=np.float32)
            self.frame = np.zeros((rs.height, rs.width, 4), dtype=np.uint8)
        em = np.ascontiguousarray(np.stack(self.em), dtype=np.float32)
        df = np.ascontiguousarray(np.stack(self.df), dtype=np.float32)
        tx = np.ascontiguousarray(self.tex, dtype=np.int32)
        self.L.rt3ref_finish(self.r, fptr(em), fptr(df), iptr(tx), C.c_char_p(self.lights[0]), C.c_int(self.lights[1]))
        self.L.rt3ref_launch(self.r, C.byref(rs), fptr(self.accum), bptr(self.frame), C.c_int(self.nthreads))

    def download_accum(self):
        return self.accum.copy()

    def download_frame(self):
        return self.frame.copy()
