"""Writes tests/golden/ref_kat_sdk.npz: known answers of the OptiX-SDK device code the reference ships under cuda/ (the stage
surface of SURVEY rows A9-A11), evaluated WHERE IT LIES through oracle/_ref/librt3ref.so (oracle/ref_shim/ref_sdk.cpp):
    cuda/sphere.cu:37-97 __intersection__sphere, cuda/LocalGeometry.h:59-178 getLocalGeometry,
    cuda/LocalShading.h:37-54 sampleTexture, cuda/curve.h:38-443 interpolators / surfaceNormal / curveTangent.
Inputs are seeded random; inputs and outputs are both stored.  Run by hand in the build container:
    python tests/golden/make_ref_kat_sdk.py
tests/test_reference_pins_sdk.py compares the oracle's restatements with these answers."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rendertoy3c_b200._abi import fptr  # noqa: E402


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def sphere_inputs(n=600, seed=41):
    r = np.random.RandomState(seed)
    cr = f32(np.concatenate([r.uniform(-5, 5, (n, 3)), r.uniform(0.05, 2.0, (n, 1))], axis=1))
    o = f32(r.uniform(-8, 8, (n, 3)))
    far = r.rand(n) < 0.3                                   # far origins: |root1| > 10 r takes the refinement branch
    o[far] = f32(o[far] * 60.0)
    inside = r.rand(n) < 0.15                               # origin inside the sphere: the second root is the answer
    o[inside] = cr[inside, :3] + f32(r.uniform(-0.3, 0.3, (inside.sum(), 3))) * cr[inside, 3:4]
    target = cr[:, :3] + f32(r.uniform(-1.2, 1.2, (n, 3))) * cr[:, 3:4]   # aims near the sphere: hits, grazes and misses
    d = f32(target - o)
    d *= f32(r.uniform(0.2, 3.0, (n, 1)))                  # unnormalised directions
    tmin = f32(np.where(r.rand(n) < 0.2, r.uniform(0.0, 1.5, n), 1e-3))
    tmax = f32(np.where(r.rand(n) < 0.2, r.uniform(0.5, 2.0, n), 1e16))
    return o, d, tmin, tmax, cr


def mesh_inputs(seed=43, nv=60, nt=40):
    r = np.random.RandomState(seed)
    P = f32(r.uniform(-2, 2, (nv, 3)))
    N = f32(r.randn(nv, 3)); N /= np.linalg.norm(N, axis=1, keepdims=True)
    UV = f32(r.uniform(0, 4, (nv, 2)))
    COL = f32(r.uniform(0, 1, (nv, 4)))
    idx = np.ascontiguousarray(np.stack([r.choice(nv, 3, replace=False) for _ in range(nt)]), dtype=np.int32)
    n = 300
    prim = r.randint(0, nt, n).astype(np.int32)
    b = r.dirichlet([1, 1, 1], n)
    bu, bv = f32(b[:, 1]), f32(b[:, 2])
    bu[:10] = 0.0; bv[10:20] = 0.0; bu[20:25] = 1.0; bv[20:25] = 0.0   # edges and a vertex
    # instance transforms: identity, and a rigid rotation + anisotropic scale + translation
    a = 0.7
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ np.array([[1, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]])
    m = rot @ np.diag([1.5, 0.75, 2.0])
    xf = f32(np.concatenate([m, np.array([[3.0], [-1.0], [0.5]])], axis=1).reshape(12))
    ident = f32([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0])
    return P, N, UV, COL, idx, prim, bu, bv, [ident, xf]


def texture_inputs(seed=47, n=400):
    r = np.random.RandomState(seed)
    tex = r.randint(0, 256, (64, 48, 4)).astype(np.uint8)
    ang = r.uniform(0, 2 * np.pi, n)
    scale = f32(r.uniform(0.25, 3.0, (n, 2)))
    rot = f32(np.stack([np.sin(ang), np.cos(ang)], axis=1))
    rot[:40] = f32([0.0, 1.0])
    off = f32(r.uniform(-2, 2, (n, 2)))
    uv = f32(r.uniform(-3, 5, (n, 2)))
    return tex, scale, rot, off, uv


def curve_inputs(seed=53, n=120):
    r = np.random.RandomState(seed)
    out = {}
    for basis, ncp in ((0, 2), (1, 3), (2, 4), (3, 4), (4, 4)):
        cp = f32(r.uniform(-1, 1, (n, ncp, 4)))
        cp[..., :3] += np.arange(ncp, dtype=np.float32)[None, :, None] * f32([0.8, 0.1, -0.2])   # a strand that moves on
        cp[..., 3] = f32(r.uniform(0.02, 0.2, (n, ncp)))
        u = f32(r.uniform(0.0, 1.0, n))
        u[:6] = 0.0; u[6:12] = 1.0
        out[basis] = (cp, u, f32(r.randn(n, 3)))
    return out


def main():
    from oracle_backend import OracleScene
    from ref_backend import available, lib
    if not available():
        sys.exit("/root/reference is not present")
    L = lib()
    out = {}
    # ---- sphere
    o, d, tmin, tmax, cr = sphere_inputs()
    res = np.zeros((len(o), 6), np.float32)
    for i in range(len(o)):
        L.rt3ref_sphere(fptr(o[i]), fptr(d[i]), C.c_float(tmin[i]), C.c_float(tmax[i]), fptr(cr[i]), fptr(res[i]))
    out["sphere"] = res
    print("sphere: %d of %d report an intersection" % (int(res[:, 0].sum()), len(o)))
    # ---- LocalGeometry (full attribute set, the two fallbacks, vertex colours)
    P, N, UV, COL, idx, prim, bu, bv, xforms = mesh_inputs()
    o_scene = OracleScene()
    for k, xf in enumerate(xforms):
        inv = np.zeros(12, np.float32)
        o_scene.L.rt3o_kat_invert_affine(fptr(xf), fptr(inv))
        for tag, (nn, uu, cc) in (("full", (N, UV, None)), ("nonormals", (None, UV, None)), ("nouvs", (N, None, None)), ("colors", (N, UV, COL))):
            lg = np.zeros((len(prim), 27), np.float32)
            for i in range(len(prim)):
                L.rt3ref_local_geometry(fptr(P), idx.ctypes.data_as(C.c_void_p), fptr(nn) if nn is not None else None, fptr(uu) if uu is not None else None,
                                        fptr(cc) if cc is not None else None, C.c_uint(int(prim[i])), C.c_float(bu[i]), C.c_float(bv[i]), fptr(xf), fptr(inv), fptr(lg[i]))
            out["lg_%s_%d" % (tag, k)] = lg
    # ---- sampleTexture
    tex, scale, rot, off, uv = texture_inputs()
    tid = o_scene.texture_create(tex, 0, 0)
    rgb = np.zeros((len(uv), 3), np.float32)
    uvt = np.zeros((len(uv), 2), np.float32)
    for i in range(len(uv)):
        L.rt3ref_sample_texture(o_scene.s, C.c_int(tid), fptr(scale[i]), fptr(rot[i]), fptr(off[i]), C.c_float(uv[i, 0]), C.c_float(uv[i, 1]), fptr(rgb[i]), fptr(uvt[i]))
    out["tex_rgb"], out["tex_uv"] = rgb, uvt
    # ---- curves
    for basis, (cp, u, dirs) in curve_inputs().items():
        ev = np.zeros((len(u), 16), np.float32)
        nrm = np.zeros((len(u), 3), np.float32)
        ps_in = np.zeros((len(u), 3), np.float32)
        ps_out = np.zeros((len(u), 3), np.float32)
        for i in range(len(u)):
            L.rt3ref_curve(C.c_int(basis), fptr(cp[i]), C.c_float(u[i]), None, fptr(ev[i]), None)
            # a point near the offset surface: centre + radius * (1 + 1e-3) along a direction orthogonal to the tangent
            t = ev[i, 12:15].astype(np.float64)
            w = dirs[i].astype(np.float64); w -= w.dot(t) * t; w /= np.linalg.norm(w)
            ps = f32(ev[i, 0:3].astype(np.float64) + w * float(ev[i, 3]) * 1.001)
            ps_in[i] = ps
            L.rt3ref_curve(C.c_int(basis), fptr(cp[i]), C.c_float(u[i]), fptr(ps), fptr(ev[i]), fptr(nrm[i]))
            ps_out[i] = ps
        out["curve%d_eval" % basis], out["curve%d_normal" % basis], out["curve%d_ps_in" % basis], out["curve%d_ps_out" % basis] = ev, nrm, ps_in, ps_out
    np.savez_compressed(os.path.join(HERE, "ref_kat_sdk.npz"), **out)
    print("wrote tests/golden/ref_kat_sdk.npz (%.1f KB)" % (os.path.getsize(os.path.join(HERE, "ref_kat_sdk.npz")) / 1024.0))


if __name__ == "__main__":
    main()
