"""Shared parity checks: a backend under test (librt3.so on the GPU, or the kernel-logic simulator on
this box) against the CPU oracle on the same seeded inputs."""
import numpy as np

from oracle_backend import OracleScene
from rendertoy3c_b200 import scenes
from rendertoy3c_b200._abi import RAY_DTYPE
from rendertoy3c_b200.api import camera_rays, make_settings

SMALL = {
    "cornell": lambda: scenes.cornell(width=64, height=64),
    "terrain": lambda: scenes.terrain(n=40, width=80, height=48, tex_size=64),
    "instanced": lambda: scenes.instanced(n_inst=27, blob_n=10, n_spheres=16, width=80, height=48),
    "motion": lambda: scenes.motion(n_inst=8, blob_n=8, n_spheres=10, n_curves=50, width=80, height=48),
    "deforming": lambda: scenes.deforming(blob_n=10, width=80, height=48),
    "splines": lambda: scenes.splines(),
    "fallbacks": lambda: scenes.fallbacks(),
}


def random_rays(desc, n, seed, extent=None):
    """half camera-like rays from the eye, half rays between random points around the scene"""
    rng = np.random.RandomState(seed)
    eye = np.asarray(desc.camera.eye, dtype=np.float32)
    look = np.asarray(desc.camera.lookat, dtype=np.float32)
    ext = extent or float(np.linalg.norm(eye - look))
    rays = np.zeros(n, dtype=RAY_DTYPE)
    h = n // 2
    d = (look - eye)[None, :] + (rng.rand(h, 3).astype(np.float32) - 0.5) * ext
    rays["o"][:h] = eye
    rays["d"][:h] = d
    a = look[None, :] + (rng.rand(n - h, 3).astype(np.float32) - 0.5) * 1.5 * ext
    b = look[None, :] + (rng.rand(n - h, 3).astype(np.float32) - 0.5) * 1.5 * ext
    rays["o"][h:] = a
    rays["d"][h:] = b - a
    rays["tmin"] = 1e-3
    rays["tmax"] = 1e16
    rays["time"] = rng.rand(n).astype(np.float32)
    return rays


def degenerate_mask(oracle, rays, hits):
    """rays excluded from the bit-exact ID check — edge/vertex grazes and near-ties.  The ID
    check in these tests is nevertheless applied to ALL rays; this mask is only reported."""
    u, v = hits["u"], hits["v"]
    hit = hits["prim"] >= 0
    w = 1.0 - u - v
    return hit & (np.minimum(np.minimum(u, v), w) < 1e-6)


def check_trace(backend, oracle, rays, accel=1):
    hb = backend.trace(rays)
    ho = oracle.trace(rays, accel=accel)
    assert np.array_equal(hb["prim"], ho["prim"]), "closest-hit primitive ids differ: %d of %d" % ((hb["prim"] != ho["prim"]).sum(), len(rays))
    assert np.array_equal(hb["inst"], ho["inst"]), "closest-hit instance ids differ"
    for k in ("t", "u", "v"):
        assert np.array_equal(hb[k].view(np.uint32), ho[k].view(np.uint32)), "hit %s not bit-identical" % k
    ab = backend.trace(rays, any_hit=True)
    ao = oracle.trace(rays, any_hit=True, accel=accel)
    assert np.array_equal(ab["prim"] >= 0, ao["prim"] >= 0), "occlusion results differ"
    # LocalGeometry stage record (cuda/LocalGeometry.h) of those hits: bit-identical, NaN-aware (degenerate uv maps give inf/nan derivatives on both sides)
    lb = backend.get_local_geometry(rays, hb).view(np.float32).reshape(len(rays), 27)
    lo = oracle.get_local_geometry(rays, ho).view(np.float32).reshape(len(rays), 27)
    same = (lb.view(np.uint32) == lo.view(np.uint32)) | (np.isnan(lb) & np.isnan(lo))
    assert same.all(), "LocalGeometry records differ in %d of %d floats" % ((~same).sum(), same.size)
    return ho


def check_render(backend, oracle, desc, subframes=2, spl=8, width=None, height=None, max_depth=None, mode=0):
    uvw = oracle.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy,
                            (width or desc.width) / (height or desc.height))
    uvw_b = backend.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy,
                               (width or desc.width) / (height or desc.height))
    for a, b in zip(uvw, uvw_b):
        assert np.array_equal(a, b), "camera frame differs"
    oracle.reset_stats()
    backend.reset_stats()
    for sf in range(subframes):
        rs = make_settings(desc, uvw, sf, samples_per_launch=spl, width=width, height=height, max_depth=max_depth, mode=mode)
        backend.launch_subframe(rs)
        oracle.launch_subframe(rs)
    ab, ao = backend.download_accum(), oracle.download_accum()
    # bit-identical, except that a NaN is a NaN: x86 and the GPU canonicalise NaN payloads differently
    # (the faithful estimator does produce NaNs: zero-area lights, cos = 0 in the Q2 weight)
    diff = (ab.view(np.uint32) != ao.view(np.uint32)) & ~(np.isnan(ab) & np.isnan(ao))
    assert not diff.any(), "accumulation buffer not bit-identical to the oracle (%d of %d floats differ)" % (diff.sum(), ab.size)
    fb, fo = backend.download_frame(), oracle.download_frame()
    assert np.abs(fb.astype(int) - fo.astype(int)).max() <= 1, "8-bit frame differs by more than 1 LSB (powf)"
    sb, so = backend.stats(), oracle.stats()
    for k in ("rays_primary", "rays_bounce", "rays_shadow", "samples"):
        assert sb[k] == so[k], "ray counter %s differs: %d vs %d" % (k, sb[k], so[k])
    assert sb["error_flags"] == 0
    return ao


def build_pair(desc, backend, options=None):
    """the scene in the backend under test and in a fresh oracle; `options`: rt3_set_option switches applied to both before
    the build (the oracle keeps the ones that change results — "flatten" — and is put back to its default afterwards)"""
    o = OracleScene()
    for k, v in (options or {}).items():
        o.set_option(k, v)
        backend.set_option(k, v)
    try:
        scenes.replay(desc, o)
    finally:
        o.set_option("flatten", 1)
    scenes.replay(desc, backend)
    return o


def check_golden(backend, name):
    """backend vs the COMMITTED oracle vectors (tests/golden/oracle_golden.json), no live oracle involved."""
    import hashlib
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))[name]
    desc = SMALL[name]()
    scenes.replay(desc, backend)
    uvw = backend.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
    rays = np.concatenate([camera_rays(desc, uvw, 32, 32), random_rays(desc, 1000, 31)])
    h = backend.trace(rays)
    core = np.stack([h["t"].view(np.uint32), h["u"].view(np.uint32), h["v"].view(np.uint32), h["prim"].view(np.uint32), h["inst"].view(np.uint32)])
    assert hashlib.sha256(core.tobytes()).hexdigest() == gold["hits_sha256"]
    assert int((h["prim"] >= 0).sum()) == gold["n_hit"]
    backend.reset_stats()
    for sf in range(2):
        backend.launch_subframe(make_settings(desc, uvw, sf))
    assert hashlib.sha256(backend.download_accum().tobytes()).hexdigest() == gold["accum_sha256"]
    st = backend.stats()
    assert [st[k] for k in ("rays_primary", "rays_bounce", "rays_shadow")] == gold["rays"]


def check_spline_tessellation(backend):
    """Spline segments are intersected as K linear pieces, K chosen per segment so that the pieces stay within
    RT3_CURVE_TOL (2 %) of the segment's radius of the true curve.  Property checked here, with the true Bezier curve
    evaluated in float64: every hit point lies within 5 % of the radius of the true swept surface (chord deviation + radius
    interpolation), on a hairpin that a fixed 8-piece split misses by more than that; and a straight spline segment
    returns the hits of the linear curve with the same end points."""
    from rendertoy3c_b200.scenes import Geometry, Instance, SceneDesc, Camera, replay
    hairpin = np.array([[-1, 0, 0, 0.02], [-0.2, 2.6, 0, 0.03], [0.2, 2.6, 0, 0.03], [1, 0, 0, 0.02]], np.float32)
    straight = np.array([[-1, -1, 0, 0.05], [-1 / 3, -1, 0, 0.05], [1 / 3, -1, 0, 0.05], [1, -1, 0, 0.05]], np.float32)
    linear = np.array([[-1, -2, 0, 0.05], [1, -2, 0, 0.05]], np.float32)
    geoms = [Geometry("curves", cr=hairpin, seg=np.array([0], np.int32), degree=5),
             Geometry("curves", cr=straight, seg=np.array([0], np.int32), degree=5),
             Geometry("curves", cr=linear, seg=np.array([0], np.int32), degree=1)]
    desc = SceneDesc("spline_tess", geoms, [Instance(0), Instance(1), Instance(2)], [], Camera(eye=(0, 0, 5), lookat=(0, 0, 0), fovy=45.0), 8, 8, 1, 1)
    replay(desc, backend)
    # orthographic rays along -z over the hairpin
    n = 160
    xs, ys = np.meshgrid(np.linspace(-1.1, 1.1, n, dtype=np.float32), np.linspace(-0.1, 2.1, n, dtype=np.float32))
    rays = np.zeros(n * n, dtype=RAY_DTYPE)
    rays["o"] = np.stack([xs.ravel(), ys.ravel(), np.full(n * n, 3, np.float32)], axis=1)
    rays["d"] = (0, 0, -1)
    rays["tmin"], rays["tmax"] = 0.0, 1e16
    h = backend.trace(rays)
    sel = (h["prim"] >= 0) & (h["inst"] == 0) & (h["u"] > 0.01) & (h["u"] < 0.99)   # away from the (round) end caps
    assert sel.sum() > 300
    P = rays["o"][sel].astype(np.float64) + h["t"][sel, None].astype(np.float64) * rays["d"][sel].astype(np.float64)
    u = np.linspace(0, 1, 4001)[:, None]
    q = hairpin.astype(np.float64)
    c = (1 - u) ** 3 * q[0] + 3 * u * (1 - u) ** 2 * q[1] + 3 * u * u * (1 - u) * q[2] + u ** 3 * q[3]      # [4001, 4]
    err = np.empty(len(P))
    for i in range(0, len(P), 256):
        d = np.linalg.norm(P[i:i + 256, None, :] - c[None, :, :3], axis=2)                                   # [m, 4001]
        err[i:i + 256] = np.abs(d - c[None, :, 3]).min(axis=1)
    assert err.max() <= 0.05 * 0.03, "hit points leave the true curve's surface by %.3g radii" % (err.max() / 0.03)
    # eight uniform pieces of this hairpin: chord deviation max|P''| / (8 * 64) — far outside the tolerance
    bend = max(np.linalg.norm(6 * (q[2] - 2 * q[1] + q[0])[:3]), np.linalg.norm(6 * (q[3] - 2 * q[2] + q[1])[:3]))
    assert bend / (8 * 64) > 0.05 * 0.03
    # straight segment: one piece, same hits as the linear curve one unit below
    ys2 = np.linspace(-1.06, -0.94, 64, dtype=np.float32)
    xs2 = np.linspace(-1.02, 1.02, 64, dtype=np.float32)
    gx, gy = np.meshgrid(xs2, ys2)
    r2 = np.zeros(gx.size * 2, dtype=RAY_DTYPE)
    r2["o"][:gx.size] = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, 3, np.float32)], axis=1)
    r2["o"][gx.size:] = r2["o"][:gx.size] - np.array([0, 1, 0], np.float32)
    r2["d"] = (0, 0, -1)
    r2["tmin"], r2["tmax"] = 0.0, 1e16
    h2 = backend.trace(r2)
    a, b = h2[:gx.size], h2[gx.size:]
    assert np.array_equal(a["prim"] >= 0, b["prim"] >= 0) and (a["prim"] >= 0).sum() > 100
    hit = a["prim"] >= 0
    assert np.all(a["inst"][hit] == 1) and np.all(b["inst"][hit] == 2)
    np.testing.assert_allclose(a["t"][hit], b["t"][hit], rtol=0, atol=1e-5)
    np.testing.assert_allclose(a["u"][hit], b["u"][hit], rtol=0, atol=1e-5)


def millimetre_scene():
    """the small instanced scene shrunk to a thousandth: every light is closer than the reference's fixed shadow-ray offsets
    (tmin 0.001, tmax = distance - 0.01, closehit_radiance.cu:132-138), so EVERY shadow ray has a negative tmax — an empty
    interval, nothing can occlude it.  (Found a marker collision in the two-pass traversal: "occluded in pass 1" used to be
    written as a negative tmax.)"""
    desc = SMALL["instanced"]()
    s = np.float32(1e-3)
    for g in desc.geoms:
        if g.kind == "mesh":
            g.verts = (g.verts * s).astype(np.float32)
        else:
            g.cr = (g.cr * s).astype(np.float32)
    for i in desc.instances:
        x = np.array(i.xform, np.float32)
        x[[3, 7, 11]] *= s
        i.xform = x
    c = desc.camera
    desc.camera = scenes.Camera(eye=tuple(np.float32(c.eye) * s), lookat=tuple(np.float32(c.lookat) * s), fovy=c.fovy)
    return desc


def check_async_frame_download(make_context):
    """rt3_download_frame_async after every subframe (the copy overlaps the next subframe, whose resolve must wait for it)
    returns the frames rt3_download_frame returns"""
    desc = SMALL["cornell"]()
    want, got = [], []
    with make_context() as a, make_context() as b:
        for c in (a, b):
            scenes.replay(desc, c)
        uvw = a.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
        for sf in range(4):
            rs = make_settings(desc, uvw, sf)
            a.launch_subframe(rs)
            want.append(a.download_frame())
            b.launch_subframe(rs)
            buf = np.zeros((desc.height, desc.width, 4), dtype=np.uint8)
            b.download_frame_async_into(buf.ctypes.data)
            got.append(buf)
        b.sync()
    for sf in range(4):
        assert np.array_equal(want[sf], got[sf]), "frame %d differs" % sf
    assert not np.array_equal(got[0], got[3])
