"""Pins of the SDK stage surface to REFERENCE CODE: tests/golden/ref_kat_sdk.npz holds answers of
cuda/sphere.cu, cuda/LocalGeometry.h, cuda/LocalShading.h and cuda/curve.h evaluated where they lie through the host
OptiX stand-in (generator tests/golden/make_ref_kat_sdk.py, which also defines the seeded inputs).  The oracle's
restatements must reproduce them: bit for bit where the arithmetic is the reference's alone, to a stated tolerance where
an OptiX-internal step sits in between (object<->world transforms) or where the oracle derives a quantity another way
(the sphere normal from the hit point instead of from the root)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_ref_kat_sdk as gen  # noqa: E402
from oracle_backend import OracleScene  # noqa: E402
from rendertoy3c_b200._abi import HIT_DTYPE, RAY_DTYPE, fptr  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_kat_sdk.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_sphere_intersection_program():
    """cuda/sphere.cu:37-97: which root is reported, and its t, bit for bit; the normal attribute within 1e-5"""
    o, d, tmin, tmax, cr = gen.sphere_inputs()
    want = GOLD["sphere"]
    s = OracleScene()
    n_hit = 0
    for i in range(len(o)):
        t = C.c_float(0)
        hit = s.L.rt3o_kat_hit_sphere(fptr(o[i]), fptr(d[i]), fptr(cr[i]), C.c_float(tmin[i]), C.c_float(tmax[i]), C.byref(t))
        assert bool(hit) == bool(want[i, 0]), i
        if hit:
            n_hit += 1
            assert bits(np.float32(t.value)) == bits(want[i, 1]), (i, t.value, want[i, 1])
            p = o[i].astype(np.float64) + np.float64(t.value) * d[i].astype(np.float64)
            nrm = (p - cr[i, :3]) / cr[i, 3]
            scale = max(1.0, np.linalg.norm(o[i] - cr[i, :3]) / cr[i, 3])          # a far origin loses that many digits in o + t d
            assert np.abs(nrm - want[i, 2:5]).max() <= 2e-6 * scale, (i, nrm, want[i, 2:5])
            assert bits(want[i, 5]) == bits(cr[i, 3])
    assert 200 < n_hit < 500


def _lg_scene(with_normals, with_uvs, with_colors):
    P, N, UV, COL, idx, prim, bu, bv, xforms = gen.mesh_inputs()
    s = OracleScene()
    b = s.mesh_create(P, idx, N if with_normals else None, UV if with_uvs else None)
    if with_colors:
        s.mesh_set_colors(b, COL)
    for xf in xforms:
        s.append_instance(b, xf)
    s.accel_build()
    return s, prim, bu, bv, len(xforms)


def _lg_records(backend, prim, bu, bv, inst):
    hits = np.zeros(len(prim), dtype=HIT_DTYPE)
    hits["prim"], hits["u"], hits["v"], hits["inst"], hits["t"] = prim, bu, bv, inst, 1.0
    rays = np.zeros(len(prim), dtype=RAY_DTYPE)
    rays["d"] = (0, 0, 1)
    lg = backend.get_local_geometry(rays, hits)
    return np.stack([np.concatenate([np.atleast_1d(r[k]).ravel() for k in ("P", "N", "Ng", "UV", "dndu", "dndv", "dpdu", "dpdv", "color")]) for r in lg]).astype(np.float32)


@pytest.mark.parametrize("tag,flags", [("full", (1, 1, 0)), ("nonormals", (0, 1, 0)), ("nouvs", (1, 0, 0)), ("colors", (1, 1, 1))])
def test_get_local_geometry(tag, flags):
    """cuda/LocalGeometry.h:59-160 incl. vertex colours (:99-110) and the fallbacks without normals (:120-124) / texcoords
    (:150-158).  Identity instance: every field bit for bit.  Transformed instance: P, N, Ng pass through OptiX's
    object->world helpers (not reference code) -> 1e-6 relative; UV, the object-space derivatives and colour stay exact."""
    s, prim, bu, bv, ninst = _lg_scene(*flags)
    for k in range(ninst):
        got = _lg_records(s, prim, bu, bv, k)
        want = GOLD["lg_%s_%d" % (tag, k)]
        exact = slice(0, 27) if k == 0 else slice(9, 27)
        if tag == "nonormals" and k > 0:
            exact = slice(9, 11)       # the derivatives of N are differences of the (transformed) geometric normal: tolerance below
        same = (bits(got[:, exact]) == bits(want[:, exact])) | (np.isnan(got[:, exact]) & np.isnan(want[:, exact]))
        assert same.all(), (tag, k, np.argwhere(~same)[:5])
        if k > 0:
            fin = np.isfinite(want)
            assert np.allclose(got[fin], want[fin], rtol=2e-6, atol=2e-6), (tag, k)


def test_sample_texture_transform():
    """cuda/LocalShading.h:37-54: UV * scale, rotated by (sin, cos), + offset, then one texel"""
    tex, scale, rot, off, uv = gen.texture_inputs()
    s = OracleScene()
    tid = s.texture_create(tex, 0, 0)
    b = s.mesh_create(np.eye(3, dtype=np.float32), np.array([[0, 1, 2]], np.int32), np.eye(3, dtype=np.float32), np.zeros((3, 2), np.float32))
    iid = s.append_instance(b, gen.f32([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]))
    s.set_hitgroup(iid, (0, 0, 0), (1, 1, 1), tid)
    same = 0
    for i in range(len(uv)):
        s.set_texture_transform(iid, scale[i], rot[i], off[i])
        got = s.sample_texture(iid, float(uv[i, 0]), float(uv[i, 1]))
        same += int(np.array_equal(bits(got), bits(GOLD["tex_rgb"][i])))
    assert same == len(uv), "%d of %d transformed fetches hit another texel than the SDK's" % (len(uv) - same, len(uv))


BASIS = {0: 1, 1: 2, 2: 3, 3: 4, 4: 5}   # generator's basis index -> the ABI's curve type (1 linear, 2 / 3 B-spline, 4 Catmull-Rom, 5 Bezier)


@pytest.mark.parametrize("basis", sorted(BASIS))
def test_curve_interpolators(basis):
    """cuda/curve.h:38-309: position4, velocity4, acceleration4 and curveTangent of every basis, bit for bit"""
    s = OracleScene()
    cp, u, _ = gen.curve_inputs()[basis]
    want = GOLD["curve%d_eval" % basis]
    for i in range(len(u)):
        out = np.zeros(16, np.float32)
        s.L.rt3o_kat_curve_eval(C.c_int(BASIS[basis]), fptr(cp[i]), C.c_float(u[i]), fptr(out))
        assert np.array_equal(bits(out), bits(want[i])), (basis, i, out, want[i])


@pytest.mark.parametrize("basis", sorted(BASIS))
def test_curve_surface_normals(basis):
    """cuda/curve.h:311-425 surfaceNormal<>: conic normal with round end caps for linear segments, the bona fide normal of
    the TRUE curve with flat end caps for quadratic / cubic ones.  The oracle's LocalGeometry of a curve hit at parameter u
    with hit point ps has exactly this normal, whatever sub-segment of its tessellation the hit was found on."""
    cp, u, _ = gen.curve_inputs()[basis]
    ps = GOLD["curve%d_ps_in" % basis]
    want = GOLD["curve%d_normal" % basis]
    for i in range(len(u)):
        s = OracleScene()
        b = s.curves_create(BASIS[basis], cp[i], np.array([0], np.int32))
        iid = s.append_instance(b, gen.f32([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]))
        s.accel_build()
        hits = np.zeros(1, dtype=HIT_DTYPE)
        hits["prim"], hits["u"], hits["inst"], hits["t"] = 0, u[i], iid, 0.0
        rays = np.zeros(1, dtype=RAY_DTYPE)
        rays["o"], rays["d"] = ps[i], (0, 0, 1)                    # t = 0: the hit point is exactly ps
        lg = s.get_local_geometry(rays, hits)[0]
        if basis == 0 or u[i] in (0.0, 1.0):
            assert np.array_equal(bits(lg["N"]), bits(want[i])), (basis, i, u[i], lg["N"], want[i])
        else:
            # the ABI carries u through the sub-segment translation (k + u_sub) / K, which can move it by an ulp
            assert np.abs(lg["N"] - want[i]).max() <= 2e-5, (basis, i, u[i], lg["N"], want[i])


def test_goldens_are_what_the_sdk_code_returns(tmp_path, monkeypatch):
    from ref_backend import available
    if not available():
        pytest.skip("/root/reference not present")
    monkeypatch.setattr(gen, "HERE", str(tmp_path))
    gen.main()
    fresh = np.load(str(tmp_path / "ref_kat_sdk.npz"))
    assert sorted(fresh.files) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(bits(fresh[k]), bits(GOLD[k])), "committed golden is stale: " + k
