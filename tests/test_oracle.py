"""The oracle against (a) known answers produced by the REFERENCE'S OWN headers compiled as host code
(tests/golden/ref_kat.json, generator oracle/ref_kat/gen_ref_kat.cpp), (b) its own brute force, and
(c) analytic cases.  The reference ships no tests or goldens, so (a) is everything reference
code can pin; hits and images are pinned by (b)/(c) and the committed oracle goldens."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_backend as ob
from parity_common import SMALL, random_rays
from rendertoy3c_b200 import scenes
from rendertoy3c_b200._abi import RAY_DTYPE, fptr
from rendertoy3c_b200.api import camera_rays, make_settings

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "ref_kat.json")))
L = ob.lib()


def f3(x):
    return np.asarray(x, dtype=np.float32)


def close(a, b, rel=1e-6, abs_=1e-7):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= abs_ + rel * np.abs(b))


def test_rng_kat_exact():
    for c in KAT["tea4_rnd"]:
        seed = L.rt3o_kat_tea4(c["v0"], c["v1"])
        assert seed == c["seed"]
        s = C.c_uint32(seed)
        r = [L.rt3o_kat_rnd(C.byref(s)) for _ in range(3)]
        assert [np.float32(x) for x in r] == [np.float32(x) for x in c["rnd"]]
        assert s.value == c["state"]


def test_cosine_sample_kat():
    # the oracle replaces libm sinf/cosf by an explicit polynomial (bit-reproducible on the GPU):
    # agreement with the reference's libm-based values is to ~1e-7 absolute, not bitwise
    for c in KAT["cosine"]:
        out = np.zeros(4, np.float32)
        L.rt3o_kat_cosine_sample(C.c_float(c["u"][0]), C.c_float(c["u"][1]), fptr(out))
        assert close(out[:3], c["w"], rel=2e-6, abs_=3e-7), (out, c)
        assert close(out[3], c["pdf"], rel=2e-6, abs_=3e-7)


def test_sincos_accuracy():
    u = np.linspace(0, 1, 20001, dtype=np.float32)[:-1]
    out = np.zeros(2, np.float32)
    worst = 0.0
    for x in u[::7]:
        L.rt3o_kat_sincos_2pi(C.c_float(x), fptr(out))
        worst = max(worst, abs(out[0] - np.sin(2 * np.pi * np.float64(x))), abs(out[1] - np.cos(2 * np.pi * np.float64(x))))
    assert worst < 4e-7


def test_onb_kat_exact():
    for c in KAT["onb"]:
        out = np.zeros(9, np.float32)
        n, w = f3(c["n"]), f3([0.3, 0.4, 0.8660254])
        L.rt3o_kat_onb(fptr(n), fptr(w), fptr(out))
        assert np.array_equal(out[0:3], f3(c["T"])) and np.array_equal(out[3:6], f3(c["B"])) and np.array_equal(out[6:9], f3(c["p"]))


def test_light_kat_exact():
    k = KAT["light"]
    buf = C.create_string_buffer(68)
    e, v0, v1, v2 = f3([17, 12, 4]), f3([343, 548.7, 227]), f3([343, 548.7, 332]), f3([213, 548.7, 332])
    L.rt3o_kat_light_make(fptr(e), fptr(v0), fptr(v1), fptr(v2), buf)
    raw = np.frombuffer(buf.raw, dtype=np.float32)
    assert k["sizeof"] == 68 and raw[16] == np.float32(k["area"]) and np.array_equal(raw[13:16], f3(k["normal"]))
    for s in k["samples"]:
        seed = C.c_uint32(s["seed"])
        out = np.zeros(7, np.float32)
        P = f3(s["P"])
        L.rt3o_kat_light_sample(buf, fptr(P), C.byref(seed), fptr(out))
        assert np.array_equal(out[0:3], f3(s["pos"])) and np.array_equal(out[3:6], f3(s["em"])) and out[6] == np.float32(s["pdf"])
        assert seed.value == s["state"]


def test_make_color_kat_exact():
    for c in KAT["make_color"]:
        out = np.zeros(4, np.uint8)
        v = f3(c["c"])
        L.rt3o_kat_make_color(fptr(v), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert list(out) == c["rgba"]


def test_camera_kat_exact():
    for c in KAT["camera"]:
        out = np.zeros(9, np.float32)
        e, l, u = f3(c["eye"]), f3(c["lookat"]), f3(c["up"])
        L.rt3o_kat_camera_uvw(fptr(e), fptr(l), fptr(u), C.c_float(c["fovy"]), C.c_float(c["aspect"]), fptr(out))
        assert np.array_equal(out, f3(c["U"] + c["V"] + c["W"]))


def test_power_heuristic_kat():
    for a, b, r in KAT["power_heuristic"]:
        a, b = np.float32(a), np.float32(b)
        assert np.float32(a * a / (a * a + b * b)) == np.float32(r)


# ---------------------------------------------------------------- analytic primitive cases
def tri(o, d, v, tmin=0.0, tmax=1e30):
    out = np.zeros(3, np.float32)
    o, d, v = f3(o), f3(d), f3(v).reshape(-1)
    hit = L.rt3o_kat_hit_triangle(fptr(o), fptr(d), fptr(v), C.c_float(tmin), C.c_float(tmax), fptr(out))
    return hit, out


def test_triangle_analytic():
    V = [[0, 0, 5], [4, 0, 5], [0, 4, 5]]
    hit, o = tri([1, 1, 0], [0, 0, 1], V)
    assert hit and o[0] == 5 and o[1] == 0.25 and o[2] == 0.25           # u -> v1, v -> v2 (OptiX convention)
    hit, o = tri([1, 1, 0], [0, 0, 2], V)
    assert hit and o[0] == 2.5                                            # t is in units of |d|
    assert tri([1, 1, 10], [0, 0, 1], V)[0] == 0                          # behind
    assert tri([1, 1, 10], [0, 0, -1], V)[0] == 1                         # no back-face culling
    assert tri([3, 3, 0], [0, 0, 1], V)[0] == 0                           # outside
    assert tri([1, 1, 0], [0, 0, 1], V, tmin=5.0)[0] == 0                 # open interval
    assert tri([1, 1, 0], [0, 0, 1], V, tmax=5.0)[0] == 0
    # watertight: a ray through the shared edge of two triangles hits at least one of them
    A = [[0, 0, 5], [4, 0, 5], [0, 4, 5]]
    B = [[4, 0, 5], [4, 4, 5], [0, 4, 5]]
    rng = np.random.RandomState(3)
    for _ in range(2000):
        s = rng.rand()
        p = np.array([4 * s, 4 * (1 - s), 5.0])
        o_ = rng.randn(3) * 3 + np.array([2, 2, -4.0])
        d_ = p - o_
        assert tri(o_, d_, A)[0] or tri(o_, d_, B)[0]


def test_sphere_analytic():
    def sph(o, d, cr, tmin=0.0, tmax=1e30):
        t = C.c_float()
        o, d, cr = f3(o), f3(d), f3(cr)
        return L.rt3o_kat_hit_sphere(fptr(o), fptr(d), fptr(cr), C.c_float(tmin), C.c_float(tmax), C.byref(t)), t.value
    assert sph([0, 0, -5], [0, 0, 1], [0, 0, 0, 1]) == (1, 4.0)
    assert sph([0, 0, 0], [0, 0, 1], [0, 0, 0, 1]) == (1, 1.0)              # inside: far root
    assert sph([0, 0, -5], [0, 0, 2], [0, 0, 0, 1]) == (1, 2.0)             # unnormalised direction
    assert sph([0, 2, -5], [0, 0, 1], [0, 0, 0, 1])[0] == 0
    hit, t = sph([0, 0, -1000], [0, 0, 1], [0, 0, 0, 0.5])                   # refinement path (|root| > 10 r)
    assert hit and abs(t - 999.5) < 1e-3


def curve(o, d, a, b, tmin=0.0, tmax=1e30):
    out = np.zeros(2, np.float32)
    o, d, a, b = f3(o), f3(d), f3(a), f3(b)
    return L.rt3o_kat_hit_curve(fptr(o), fptr(d), fptr(a), fptr(b), C.c_float(tmin), C.c_float(tmax), fptr(out)), out


def test_curve_analytic_and_union_of_spheres():
    hit, o = curve([0.5, 0, -5], [0, 0, 1], [0, 0, 0, 0.2], [1, 0, 0, 0.2])
    assert hit and abs(o[0] - 4.8) < 1e-5 and abs(o[1] - 0.5) < 1e-5         # cylinder body
    hit, o = curve([-0.1, 0, -5], [0, 0, 1], [0, 0, 0, 0.2], [1, 0, 0, 0.2])
    assert hit and o[1] == 0.0                                               # end cap a
    assert curve([0.5, 0, 0], [0, 0, 1], [0, 0, 0, 0.2], [1, 0, 0, 0.2])[0] == 0  # origin inside: entry hits only
    # against a dense union of spheres (the definition of the primitive)
    rng = np.random.RandomState(5)
    a = np.array([0.1, -0.2, 0.3, 0.25]); b = np.array([0.9, 0.4, -0.1, 0.08])
    s = np.linspace(0, 1, 4001)
    cen = a[None, :3] + s[:, None] * (b[:3] - a[:3])[None, :]
    rad = a[3] + s * (b[3] - a[3])
    n_hit = 0
    for _ in range(400):
        o_ = rng.randn(3); o_ = 3 * o_ / np.linalg.norm(o_)
        d_ = (a[:3] + rng.rand() * (b[:3] - a[:3]) + 0.3 * rng.randn(3)) - o_
        d_ /= np.linalg.norm(d_)
        oc = o_[None, :] - cen
        bq = oc @ d_
        disc = bq * bq - (np.sum(oc * oc, axis=1) - rad * rad)
        ok = disc > 0
        tref = np.min(-bq[ok] - np.sqrt(disc[ok])) if ok.any() else None
        hit, o = curve(o_, d_, a, b)
        if tref is None or not hit:
            if tref is not None and (ok.sum() > 8):
                assert hit, "missed a clear hit"
            continue
        n_hit += 1
        assert abs(o[0] - tref) < 2e-3, (o, tref)
    assert n_hit > 50


def test_invert_affine_identity_is_exact():
    m = scenes.IDENTITY.copy()
    out = np.zeros(12, np.float32)
    L.rt3o_kat_invert_affine(fptr(m), fptr(out))
    assert np.array_equal(np.abs(out), m)


# ---------------------------------------------------------------- BVH2 path == brute force, images, goldens
@pytest.mark.parametrize("name", sorted(SMALL))
def test_oracle_bvh_equals_bruteforce(name):
    desc = SMALL[name]()
    o = ob.OracleScene()
    scenes.replay(desc, o)
    uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
    rays = np.concatenate([camera_rays(desc, uvw, 64, 64), random_rays(desc, 3000, 23)])
    for any_hit in (False, True):
        a, b = o.trace(rays, any_hit=any_hit, accel=0), o.trace(rays, any_hit=any_hit, accel=1)
        if any_hit:
            assert np.array_equal(a["prim"] >= 0, b["prim"] >= 0)
        else:
            assert a.tobytes() == b.tobytes()
    assert (a["prim"] >= 0).mean() > 0.05


def test_edge_cases_empty_and_ragged():
    desc = SMALL["cornell"]()
    o = ob.OracleScene()
    scenes.replay(desc, o)
    assert len(o.trace(np.zeros(0, RAY_DTYPE))) == 0
    r = np.zeros(3, RAY_DTYPE)
    r["o"] = [278, 273, -800]; r["d"] = [[0, 0, 1], [0, 0, -1], [0, 0, 1]]
    r["tmin"] = [0.01, 0.01, 0.01]; r["tmax"] = [1e16, 1e16, 5.0]      # hit / pointing away / interval too short
    h = o.trace(r, accel=0)
    assert h["prim"][0] >= 0 and h["prim"][1] == -1 and h["prim"][2] == -1 and h["inst"][1] == -1


def test_chain_sum_association_is_rounding_only():
    """The reference keeps one running sum over the samples of a launch (raygen.cu:27,58-59); the oracle's
    default (per-sample sums, the wavefront association) differs by fp32 rounding only."""
    desc = SMALL["cornell"]()
    imgs = []
    for chain in (0, 1):
        L.rt3o_set_chain_sum(chain)
        o = ob.OracleScene()
        scenes.replay(desc, o)
        uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, 1.0)
        o.launch_subframe(make_settings(desc, uvw, 0))
        imgs.append(o.download_accum()[..., :3].astype(np.float64))
    L.rt3o_set_chain_sum(0)
    rel = np.abs(imgs[0] - imgs[1]) / np.maximum(np.abs(imgs[1]), 1e-3)
    assert rel.max() < 5e-6


def test_oracle_golden_fixture():
    """tests/golden/oracle_golden.json (made by tests/golden/make_golden.py) pins the oracle itself."""
    import make_golden
    gold = json.load(open(os.path.join(GOLD, "oracle_golden.json")))
    now = make_golden.compute(ob)
    assert now == gold


def test_vertex_key_motion_analytic():
    """N2: a unit triangle sliding +2 in x over the shutter (2 keys); a fixed ray at x = 1.5 only hits it
    while the interpolated triangle covers that x, and 3 keys interpolate piecewise."""
    from rendertoy3c_b200.scenes import IDENTITY
    tri0 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    keys = np.stack([tri0, tri0 + np.array([2, 0, 0], np.float32)])
    o = ob.OracleScene()
    h = o.mesh_create(keys, np.array([[0, 1, 2]], np.int32), np.tile([[0, 0, 1]], (3, 1)), np.zeros((3, 2)))
    o.append_instance(h, IDENTITY)
    o.accel_build()
    r = np.zeros(5, RAY_DTYPE)
    r["o"] = [1.6, 0.2, -1]; r["d"] = [0, 0, 1]; r["tmin"] = 0; r["tmax"] = 10
    r["time"] = [0.0, 0.3, 0.5, 0.75, 1.0]          # triangle spans x in [2t, 2t+1-y]: covers 1.6 (y=0.2) for t in (0.4, 0.8)
    hits = o.trace(r, accel=0)
    assert list(hits["prim"] >= 0) == [False, False, True, True, False]
    assert np.allclose(hits["t"][2:4], 1.0)
    assert np.allclose(hits["u"][2], 1.6 - 1.0) and np.allclose(hits["u"][3], 1.6 - 1.5)   # u = weight of v1 = x offset in the moved triangle
    assert o.trace(r, accel=1).tobytes() == hits.tobytes()


def test_corrected_mode_matches_analytic_direct_lighting():
    """N4: NEE + BSDF-sampled emitter hits with power-heuristic MIS reproduce the closed-form radiance of a
    diffuse floor under a rectangular Lambertian emitter; the faithful estimator (quirks Q2-Q4) does not."""
    import corrected_cases as cc
    desc = cc.furnace_scene()
    want = cc.analytic_radiance()
    o = ob.OracleScene()
    scenes.replay(desc, o)
    got = cc.render_mean(o, desc, subframes=24, mode=1, max_depth=2)     # 24*24 px * 192 spp
    assert abs(got - want) / want < 0.02, (got, want)
    o2 = ob.OracleScene()
    scenes.replay(desc, o2)
    faithful = cc.render_mean(o2, desc, subframes=24, mode=0, max_depth=2)
    assert abs(faithful - want) / want > 0.1                               # documents that mode 0 is not physically correct (F6)


def test_power_light_sampler_matches_analytic_two_lights():
    """N4 / reference README "power light sampler": mode 2 chooses lights in proportion to luminance x area; with a
    bright and a dim emitter it converges to the same closed form as the uniform sampler (mode 1), from different samples"""
    import corrected_cases as cc
    desc = cc.two_light_scene()
    want = cc.analytic_two_lights()
    got = {}
    for mode in (1, 2):
        o = ob.OracleScene()
        scenes.replay(desc, o)
        got[mode] = cc.render_mean(o, desc, subframes=24, mode=mode, max_depth=2)
        assert abs(got[mode] - want) / want < 0.02, (mode, got[mode], want)
    assert got[1] != got[2]


def test_local_geometry_record_analytic():
    """A9 (cuda/LocalGeometry.h:61-175): world P from interpolated vertices, unit N / Ng in world space, UV, and the
    object-space derivatives of a triangle whose uv map is the identity on its edges"""
    from rendertoy3c_b200._abi import RAY_DTYPE
    from rendertoy3c_b200.scenes import IDENTITY
    o = ob.OracleScene()
    verts = np.array([[0, 0, 0], [2, 0, 0], [0, 3, 0]], np.float32)
    normals = np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1]], np.float32)
    uvs = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    m = o.mesh_create(verts, np.array([[0, 1, 2]], np.int32), normals, uvs)
    xf = IDENTITY.copy()                                            # rotate 90 degrees about x (z -> -y), then translate by (5, 6, 7)
    xf[:] = [1, 0, 0, 5, 0, 0, -1, 6, 0, 1, 0, 7]
    i0 = o.append_instance(m, IDENTITY)
    i1 = o.append_instance(m, xf)
    for i in (i0, i1):
        o.set_hitgroup(i, (0, 0, 0), (0.5, 0.5, 0.5), -1)
    o.accel_build()
    rays = np.zeros(3, RAY_DTYPE)
    rays["o"] = [[0.5, 0.75, 4.0], [5.5, 10.0, 7.75], [50, 50, 50]]
    rays["d"] = [[0, 0, -1], [0, -1, 0], [0, 0, 1]]
    rays["tmax"] = 1e16
    hits = o.trace(rays, accel=0)
    assert list(hits["inst"]) == [i0, i1, -1] and list(hits["prim"][:2]) == [0, 0]
    lg = o.get_local_geometry(rays, hits)
    assert np.allclose(lg["P"][0], [0.5, 0.75, 0.0], atol=1e-6) and np.allclose(lg["P"][1], [5.5, 6.0, 7.75], atol=1e-5)
    assert np.allclose(lg["P"][:2], rays["o"][:2] + hits["t"][:2, None] * rays["d"][:2], atol=1e-5)
    assert np.allclose(lg["Ng"][0], [0, 0, 1]) and np.allclose(lg["N"][0], [0, 0, 1])
    assert np.allclose(lg["Ng"][1], [0, -1, 0], atol=1e-6) and np.allclose(lg["N"][1], [0, -1, 0], atol=1e-6)   # normals follow the rotation
    assert np.allclose(lg["UV"][0], [0.25, 0.25], atol=1e-6) and np.allclose(lg["UV"][1], [0.25, 0.25], atol=1e-5)
    for k in range(2):                                              # derivatives stay in object space, like the SDK's
        assert np.allclose(lg["dpdu"][k], [2, 0, 0]) and np.allclose(lg["dpdv"][k], [0, 3, 0])
        assert np.allclose(lg["dndu"][k], 0) and np.allclose(lg["dndv"][k], 0)
        assert np.allclose(lg["color"][k], 1)
    assert not lg[2].tobytes().strip(b"\0")                         # miss: all-zero record


def test_texture_address_and_filter_modes():
    """tex2D restatement: point / bilinear x wrap / clamp / mirror / border on a 4x2 texture with known texels"""
    o = ob.OracleScene()
    img = np.zeros((2, 4, 4), np.uint8)
    img[0, :, 0] = [0, 51, 102, 153]                                 # row 0: red ramp 0, .2, .4, .6
    img[1, :, 0] = [255, 204, 153, 102]                              # row 1: 1, .8, .6, .4
    img[..., 3] = 255
    ids = {(a, f): o.texture_create(img, a, f) for a in range(4) for f in range(2)}
    r = lambda a, f, u, v: float(o.fetch_texture(ids[(a, f)], u, v)[0])
    for a in range(4):                                               # inside the image every address mode agrees
        assert abs(r(a, 0, 0.375, 0.25) - 0.2) < 1e-6                # point: texel (1, 0)
        assert abs(r(a, 1, 0.375, 0.25) - 0.2) < 1e-6                # bilinear at a texel centre = that texel
        assert abs(r(a, 1, 0.5, 0.25) - 0.3) < 1e-6                  # half way between texels 1 and 2 of row 0
        assert abs(r(a, 1, 0.375, 0.5) - 0.5) < 1e-6                 # half way between rows: (0.2 + 0.8) / 2
    # outside: u = 1.125 is texel centre 4 (one past the edge), u = -0.125 is texel centre -1
    assert abs(r(0, 0, 1.125, 0.25) - 0.0) < 1e-6 and abs(r(0, 0, -0.125, 0.25) - 0.6) < 1e-6     # wrap
    assert abs(r(1, 0, 1.125, 0.25) - 0.6) < 1e-6 and abs(r(1, 0, -0.125, 0.25) - 0.0) < 1e-6     # clamp
    assert abs(r(2, 0, 1.125, 0.25) - 0.6) < 1e-6 and abs(r(2, 0, 1.375, 0.25) - 0.4) < 1e-6      # mirror: 4 -> 3, 5 -> 2
    assert abs(r(2, 0, -0.125, 0.25) - 0.0) < 1e-6 and abs(r(2, 0, -0.375, 0.25) - 0.2) < 1e-6    # mirror: -1 -> 0, -2 -> 1
    assert r(3, 0, 1.125, 0.25) == 0.0 and r(3, 0, -0.125, 0.25) == 0.0                           # border
    assert abs(r(3, 1, 1.0, 0.25) - 0.3) < 1e-6                      # bilinear on the edge: half texel 3 (0.6), half border (0)
    assert abs(r(0, 1, 1.0, 0.25) - 0.3) < 1e-6                      # wrap: half texel 3 (0.6), half texel 0 (0.0)
    assert abs(r(1, 1, 1.0, 0.25) - 0.6) < 1e-6                      # clamp: both neighbours are texel 3
    assert abs(r(0, 1, 0.375 + 1.0 / 1024, 0.25) - (0.2 + 0.2 / 256)) < 1e-6                      # weights have 8 fractional bits: 1/256 steps


def test_sample_texture_texcoord_transform():
    """cuda/LocalShading.h:37-54: UV' = (dot(UV*scale, (cos, sin)), dot(UV*scale, (-sin, cos))) + offset, then tex2D"""
    from rendertoy3c_b200.scenes import IDENTITY
    o = ob.OracleScene()
    img = np.zeros((4, 4, 4), np.uint8)
    img[..., 0] = np.arange(16).reshape(4, 4) * 16                  # red = 16 * (4 * row + column)
    t = o.texture_create(img, 0, 0)
    m = o.mesh_create(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.int32),
                      np.array([[0, 0, 1]] * 3, np.float32), np.array([[0, 0], [1, 0], [0, 1]], np.float32))
    plain, moved = o.append_instance(m, IDENTITY), o.append_instance(m, IDENTITY)
    for i in (plain, moved):
        o.set_hitgroup(i, (0, 0, 0), (1, 1, 1), t)
    o.set_texture_transform(moved, (2.0, 0.5), (1.0, 0.0), (0.25, 0.75))   # 90 degrees: (u, v) -> (v', -u') with u' = 2u, v' = v/2
    texel = lambda u, v: float(o.fetch_texture(t, u, v)[0])
    assert float(o.sample_texture(plain, 0.3, 0.6)[0]) == texel(0.3, 0.6)
    for u, v in ((0.1, 0.2), (0.3, 0.6), (0.45, 0.9)):
        su, sv = np.float32(u) * np.float32(2.0), np.float32(v) * np.float32(0.5)
        want = texel(float(sv) + 0.25, float(-su) + 0.75)
        assert float(o.sample_texture(moved, u, v)[0]) == want


def test_bspline_curves_analytic():
    """degree 2 / 3 curves (SDK Quadratic / CubicInterpolator from uniform B-spline control points, cuda/curve.h:98-230): with collinear,
    equally spaced control points of one radius the segment is a straight tube between P(0) and P(1); hits report (segment, u)"""
    from rendertoy3c_b200._abi import RAY_DTYPE
    from rendertoy3c_b200.scenes import IDENTITY
    o = ob.OracleScene()
    cp = np.array([[x, 0, 0, 0.1] for x in range(6)], np.float32)
    quad = o.curves_create(2, cp, np.array([0, 1, 2, 3], np.int32))      # segment s spans x in [s + 0.5, s + 1.5]
    cubic = o.curves_create(3, cp, np.array([0, 1, 2], np.int32))         # segment s spans x in [s + 1, s + 2]
    shift = IDENTITY.copy(); shift[7] = 10.0                                # the cubic strand 10 higher
    iq, ic = o.append_instance(quad, IDENTITY), o.append_instance(cubic, shift)
    for i in (iq, ic):
        o.set_hitgroup(i, (0, 0, 0), (0.5, 0.5, 0.5), -1)
    o.accel_build()
    rays = np.zeros(4, RAY_DTYPE)
    rays["o"] = [[1.0, 5.0, 0.0], [2.75, 5.0, 0.0], [1.25, 15.0, 0.0], [0.2, 5.0, 0.0]]
    rays["d"] = [[0, -1, 0]] * 4
    rays["tmax"] = 1e16
    h = o.trace(rays, accel=0)
    assert list(h["inst"]) == [iq, iq, ic, -1]                            # x = 0.2 is before the first quadratic segment starts (0.5 - r)
    assert list(h["prim"][:3]) == [0, 2, 0]
    assert np.allclose(h["t"][:3], 4.9, atol=1e-5)
    assert np.allclose(h["u"][:3], [0.5, 0.25, 0.25], atol=1e-5)
    assert o.trace(rays, accel=1).tobytes() == h.tobytes()
    lg = o.get_local_geometry(rays, h)
    assert np.allclose(lg["N"][:3], [0, 1, 0], atol=1e-5) and np.allclose(lg["UV"][:3, 0], [0.5, 0.25, 0.25], atol=1e-5)
    assert np.allclose(lg["P"][:3, 1], [0.1, 0.1, 10.1], atol=1e-5)


def test_spline_tessellation_tolerance():
    """the oracle's adaptive split of spline segments keeps hits within the stated tolerance of the true curve"""
    from parity_common import check_spline_tessellation
    check_spline_tessellation(ob.OracleScene())


def test_flattened_instances_bvh_equals_bruteforce_far_from_the_origin():
    """flattened instances are culled with the object-space BVH but tested in world space; the culling slack must cover the
    rounding of the ray transform where it is worst — small, rotated, sheared copies far from the origin, and big ones"""
    rng = np.random.RandomState(31)
    blob = scenes.grid_mesh(24, 24, scenes._blob_pos)
    inst, centres, radii = [], [], []
    for k in range(10):
        c = (rng.rand(3) - 0.5) * [2000.0, 600.0, 2000.0]
        sc = [0.01, 0.05, 1.0, 40.0, 0.3][k % 5]
        m = scenes.rigid(rng, 180.0, c, sc).reshape(3, 4)
        if k % 2:   # anisotropic scale + shear
            m[:, :3] = m[:, :3] @ np.array([[1.0, 0.4, 0.0], [0.0, 2.5, 0.0], [0.3, 0.0, 0.6]], np.float32)
        inst.append(scenes.Instance(0, xform=m.reshape(12).astype(np.float32)))
        centres.append(c)
        radii.append(sc)
    desc = scenes.SceneDesc("far_instances", [blob], inst, [], scenes.Camera(eye=(0, 0, 3000), lookat=(0, 0, 0), fovy=45.0), 8, 8, 1, 1)
    o = ob.OracleScene()
    scenes.replay(desc, o)
    n = 20000
    rays = np.zeros(n, dtype=RAY_DTYPE)
    which = rng.randint(0, len(inst), n)
    c = np.asarray(centres)[which]
    r = np.asarray(radii)[which][:, None]
    target = c + (rng.rand(n, 3) - 0.5) * 1.6 * r
    origin = np.where(rng.rand(n, 1) < 0.5, c + rng.randn(n, 3) * 6.0 * r, (rng.rand(n, 3) - 0.5) * 3000.0)
    rays["o"] = origin.astype(np.float32)
    rays["d"] = (target - origin).astype(np.float32)
    rays["tmin"], rays["tmax"] = 1e-4, 1e16
    a, b = o.trace(rays, accel=0), o.trace(rays, accel=1)
    assert (a["prim"] >= 0).mean() > 0.3
    assert a.tobytes() == b.tobytes()
    assert np.array_equal(o.trace(rays, any_hit=True, accel=0)["prim"] >= 0, o.trace(rays, any_hit=True, accel=1)["prim"] >= 0)
