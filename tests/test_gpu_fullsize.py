"""BASELINE.json configs at FULL geometric size on the B200 against the oracle: hit parity on large
random ray batches (oracle BVH2 + a brute-force subset) and a bit-exact subframe at reduced
resolution (the oracle is a scalar CPU tracer; full 1080p x 64 spp is minutes of CPU time), plus
size-independent properties at the full 1080p film: determinism, N-way sample partition == single
run, ray-count conservation."""
import numpy as np
import pytest

from parity_common import build_pair, check_render, check_trace, random_rays
from rendertoy3c_b200 import scenes
from rendertoy3c_b200.api import Context, camera_rays, make_settings

pytestmark = pytest.mark.gpu


def _fullsize(desc, n_rays, brute, width, height, options=None, expect=None):
    with Context(0) as g:
        o = build_pair(desc, g, options)
        for k, v in (expect or {}).items():
            assert g.stats()[k] == v, (k, g.stats()[k], v)
        uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, width / height)
        rays = np.concatenate([camera_rays(desc, uvw, 256, 144), random_rays(desc, n_rays, 41)])
        ho = check_trace(g, o, rays, accel=1)
        assert (ho["prim"] >= 0).mean() > 0.1
        check_trace(g, o, rays[:: max(1, len(rays) // brute)], accel=0)
        check_render(g, o, desc, subframes=1, width=width, height=height)
        assert g.stats()["error_flags"] == 0


def test_c2_terrain_1m_triangles():
    _fullsize(scenes.terrain(), 150000, 192, 480, 270)           # 1,002,530 triangles, textured, single level


def test_c3_thousand_instances_plus_spheres():
    # 1000 x 100,352-triangle instances + 1000 spheres: the instances are flattened into one 100 M-triangle world BLAS (pass 1),
    # the spheres stay behind a TLAS (pass 2)
    _fullsize(scenes.instanced(), 100000, 24, 320, 180, expect={"flattened_instances": 1000, "traversal_passes": 2})


def test_c3_thousand_instances_kept_as_instances():
    _fullsize(scenes.instanced(), 100000, 24, 320, 180, options={"flatten": 0}, expect={"flattened_instances": 0})   # one BLAS, 1000 TLAS leaves


def test_c4_motion_blur_tris_spheres_curves():
    _fullsize(scenes.motion(), 100000, 48, 320, 180)             # 64 two-key instances, 256 moving spheres, 10k curve segments


def test_c2_full_1080p_subframe_bit_identical():
    """the headline workload as it is timed — 1,002,530 triangles, 1920x1080, 8 samples per pixel, depth 8: 16.6 M paths, ~41 M
    rays — against the oracle on all host cores: every float of the accumulation buffer, every ray counter"""
    desc = scenes.terrain()
    with Context(0) as g:
        o = build_pair(desc, g)
        check_render(g, o, desc, subframes=1, width=1920, height=1080)


def test_c2_full_film_properties():
    """1920x1080, 8 spl: same input twice -> identical bits; subframes {0,1} rendered as one context or as
    two sample partitions (SUM mode) agree; device ray counters add up."""
    desc = scenes.terrain()
    with Context(0) as a, Context(0) as b:
        for c in (a, b):
            scenes.replay(desc, c)
        uvw = a.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, 1920 / 1080)
        a.clear_accum()
        for sf in (0, 1):
            a.launch_subframe(make_settings(desc, uvw, sf, accum_mode=1))
        img_a = a.download_accum()
        sa = a.stats()
        # partitioned: context b renders subframe 1 then 0 into separate sums
        b.clear_accum()
        b.launch_subframe(make_settings(desc, uvw, 1, accum_mode=1))
        p1 = b.download_accum()
        b.clear_accum()
        b.launch_subframe(make_settings(desc, uvw, 0, accum_mode=1))
        p0 = b.download_accum()
        sb = b.stats()
        assert np.array_equal((p0 + p1).view(np.uint32), img_a.view(np.uint32))      # fp32 add of two terms commutes
        for k in ("rays_primary", "rays_bounce", "rays_shadow", "samples"):
            assert sa[k] == sb[k]
        assert sa["rays_primary"] == 2 * 1920 * 1080 * 8
        assert np.isfinite(img_a[..., :3]).mean() > 0.9999


def test_c5_4k_film_one_subframe():
    """BASELINE configs[4] film size: 3840x2160 x 8 spl = 66.4 M paths in flight (20 GB of queue pools).
    Rendered twice: identical bits; counters consistent; no stack overflow."""
    desc = scenes.terrain()
    with Context(0) as g:
        scenes.replay(desc, g)
        uvw = g.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, 3840 / 2160)
        imgs = []
        for _ in range(2):
            g.reset_stats()
            g.clear_accum()
            g.launch_subframe(make_settings(desc, uvw, 5, width=3840, height=2160, accum_mode=1))
            imgs.append(g.download_accum())
            st = g.stats()
            assert st["rays_primary"] == 3840 * 2160 * 8 and st["error_flags"] == 0
            assert st["rays_bounce"] > 0 and st["rays_shadow"] > 0
        same = (imgs[0].view(np.uint32) == imgs[1].view(np.uint32)) | (np.isnan(imgs[0]) & np.isnan(imgs[1]))
        assert same.all()
        # 4K subframe 5 restricted to even pixels is NOT the 1080p image (different pixel seeds), but the mean radiance must agree
        g.launch_subframe(make_settings(desc, uvw, 0, width=1920, height=1080))
        lo = g.download_accum()
        m4, m2 = np.nanmean(imgs[0][..., :3]), np.nanmean(lo[..., :3])
        assert abs(m4 - m2) / m2 < 0.05
