"""Parity tests proper: librt3.so on the B200 (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar: hit ids, t, u, v and the float accumulation buffer BIT-IDENTICAL; the 8-bit
sRGB frame within 1 LSB (device powf)."""
import numpy as np
import pytest

from parity_common import SMALL, build_pair, check_render, check_trace, random_rays
from rendertoy3c_b200 import scenes
from rendertoy3c_b200.api import Context, camera_rays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(SMALL))
def test_small_scenes_trace_and_render(product_lib, name):
    desc = SMALL[name]()
    with Context(0) as g:
        o = build_pair(desc, g)
        uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
        rays = np.concatenate([camera_rays(desc, uvw, 48, 48), random_rays(desc, 1500, 7)])
        check_trace(g, o, rays, accel=0)  # brute-force oracle
        check_render(g, o, desc, subframes=2)


def test_cornell_c1_full(product_lib):
    """BASELINE.json configs[0]: Cornell 512x512, 16 spp (2 subframes x 8), depth 4 — full size."""
    desc = scenes.cornell()
    with Context(0) as g:
        o = build_pair(desc, g)
        check_render(g, o, desc, subframes=2)


def test_terrain_medium_trace(product_lib):
    desc = scenes.terrain(n=256, width=320, height=180, tex_size=256)  # 131k triangles
    with Context(0) as g:
        o = build_pair(desc, g)
        uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, 16 / 9)
        rays = np.concatenate([camera_rays(desc, uvw, 320, 180), random_rays(desc, 100000, 11)])
        check_trace(g, o, rays, accel=1)           # oracle BVH2 (validated against brute force in test_oracle.py)
        check_trace(g, o, rays[::97], accel=0)      # brute-force subset
        check_render(g, o, desc, subframes=1, width=160, height=90)


def test_instanced_motion_medium(product_lib):
    desc = scenes.motion(n_inst=27, blob_n=32, n_spheres=64, n_curves=500, width=160, height=90)
    with Context(0) as g:
        o = build_pair(desc, g)
        rays = random_rays(desc, 50000, 13)
        check_trace(g, o, rays, accel=1)
        check_render(g, o, desc, subframes=1)
    desc = scenes.instanced(n_inst=125, blob_n=32, n_spheres=100, width=160, height=90)
    with Context(0) as g:
        o = build_pair(desc, g)
        rays = random_rays(desc, 50000, 17)
        check_trace(g, o, rays, accel=1)
        check_render(g, o, desc, subframes=1)


def test_unbounded_depth(product_lib):
    desc = SMALL["cornell"]()
    with Context(0) as g:
        o = build_pair(desc, g)
        check_render(g, o, desc, subframes=1, width=48, height=48, max_depth=0)


@pytest.mark.parametrize("name", sorted(SMALL))
def test_committed_golden_vectors(product_lib, name):
    """kernels vs tests/golden/oracle_golden.json — committed vectors, no live oracle on the GPU box"""
    from parity_common import check_golden
    with Context(0) as g:
        check_golden(g, name)


def test_empty_and_ragged_batches(product_lib):
    from rendertoy3c_b200._abi import RAY_DTYPE
    desc = SMALL["cornell"]()
    with Context(0) as g:
        o = build_pair(desc, g)
        assert len(g.trace(np.zeros(0, RAY_DTYPE))) == 0
        for n in (1, 31, 33, 1000):                     # not multiples of the warp size
            rays = random_rays(desc, n, 100 + n)
            check_trace(g, o, rays, accel=0)


@pytest.mark.parametrize("offset", [0.0, 4096.0])
def test_adversarial_geometry_bit_exact(product_lib, offset):
    """ties, degenerate triangles, flat boxes, far-from-origin scene, axis-parallel and in-plane rays"""
    import adversarial
    desc, off = adversarial.make_scene(offset)
    with Context(0) as g:
        o = build_pair(desc, g)
        ho = check_trace(g, o, adversarial.make_rays(off, n=20000), accel=0)
        assert (ho["prim"] >= 0).mean() > 0.1
        check_render(g, o, desc, subframes=2)


def test_stacked_layers_queue_overflow_and_ties(product_lib):
    """single-level kernel: dozens of coincident / near-coincident triangles pending per ray and round"""
    import adversarial
    desc = adversarial.make_stacked_scene()
    with Context(0) as g:
        o = build_pair(desc, g)
        ho = check_trace(g, o, adversarial.make_stacked_rays(), accel=0)
        assert (ho["prim"] >= 0).mean() > 0.5
        check_render(g, o, desc, subframes=2)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", sorted(SMALL))
def test_corrected_mode_matches_oracle(product_lib, name, mode):
    """mode 1 (unbiased Lambert + NEE + MIS) and mode 2 (the same with the power light sampler):
    bit-identical to the oracle's corrected integrator"""
    desc = SMALL[name]()
    with Context(0) as g:
        o = build_pair(desc, g)
        check_render(g, o, desc, subframes=2, mode=mode)


def test_power_light_sampler_analytic_two_lights(product_lib):
    import corrected_cases as cc
    desc = cc.two_light_scene(width=64, height=64)
    want = cc.analytic_two_lights()
    with Context(0) as g:
        scenes.replay(desc, g)
        got = cc.render_mean(g, desc, subframes=32, mode=2, max_depth=2)
    assert abs(got - want) / want < 0.02, (got, want)


def test_corrected_mode_analytic_direct_lighting(product_lib):
    import corrected_cases as cc
    desc = cc.furnace_scene(width=64, height=64)
    with Context(0) as g:
        scenes.replay(desc, g)
        got = cc.render_mean(g, desc, subframes=32, mode=1, max_depth=2)   # 64*64 px * 256 spp
    want = cc.analytic_radiance()
    assert abs(got - want) / want < 0.02, (got, want)


@pytest.mark.parametrize("address,filt", [(2, 0), (3, 0), (0, 1), (1, 1), (2, 1), (3, 1)])
def test_texture_modes_match_oracle(product_lib, address, filt):
    """CUDATexture's other address modes (mirror, border) and the hardware-bilinear filter (the reference's FilterMode::Point = 1)"""
    desc = SMALL["terrain"]()
    for t in desc.textures:
        t.address, t.filter = address, filt
    with Context(0) as g:
        o = build_pair(desc, g)
        check_render(g, o, desc, subframes=2)


def test_texcoord_transform_matches_oracle(product_lib):
    """sampleTexture's scale / rotation / offset (cuda/LocalShading.h:37-54) on the textured terrain"""
    desc = SMALL["terrain"]()
    for inst in desc.instances:
        if inst.tex >= 0:
            inst.tex_xform = ((1.5, 0.75), (float(np.sin(0.6)), float(np.cos(0.6))), (0.125, -0.3))
    with Context(0) as g:
        o = build_pair(desc, g)
        check_render(g, o, desc, subframes=1)


@pytest.mark.parametrize("mode", [1, 2])
def test_corrected_mode_with_emitters_nee_cannot_reach(product_lib, mode):
    """every instance of the motion scene emits — spheres, curves, animated and transformed meshes.  Their BSDF-sampled
    hits keep the full weight (the light list only holds object-space triangles of identity meshes, Q15); none of them
    may be looked up as a static triangle mesh (null vertex arrays on spheres / curves)"""
    desc = SMALL["motion"]()
    for k, inst in enumerate(desc.instances):
        inst.emission = (0.5 + 0.25 * (k % 3), 0.4, 0.3)
    with Context(0) as e:
        o = build_pair(desc, e)
        check_render(e, o, desc, subframes=2, mode=mode)


def test_spline_tessellation_tolerance(product_lib):
    """adaptive split of spline segments: hits stay within the stated tolerance of the true curve (parity_common)"""
    from parity_common import check_spline_tessellation
    with Context(0) as e:
        check_spline_tessellation(e)


@pytest.mark.parametrize("options", [{"split": 2}, {"split": 0}, {"flatten": 0}, {"flatten": 0, "split": 2}, {"merge_identity": 0, "flatten": 0}],
                         ids=lambda o: ",".join("%s=%d" % kv for kv in o.items()))
@pytest.mark.parametrize("name", ["instanced", "motion", "deforming", "splines"])
def test_flatten_and_split_variants_match_oracle(product_lib, name, options):
    """static instances flattened into the merged world BLAS or entered through the TLAS, traversed in one launch or two
    (single-level kernel, then the rest seeded with its result): bit-identical to the oracle either way"""
    desc = SMALL[name]()
    with Context(0) as e:
        o = build_pair(desc, e, options)
        uvw = e.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
        rays = np.concatenate([camera_rays(desc, uvw, 40, 24), random_rays(desc, 1500, seed=5)])
        check_trace(e, o, rays)
        check_render(e, o, desc, subframes=1)


@pytest.mark.parametrize("split", [0, 2])
def test_shadow_rays_with_negative_tmax(product_lib, split):
    from parity_common import millimetre_scene
    desc = millimetre_scene()
    with Context(0) as e:
        o = build_pair(desc, e, {"split": split})
        check_render(e, o, desc, subframes=2)
        assert e.stats()["rays_shadow"] > 100


def test_async_frame_download(product_lib):
    from parity_common import check_async_frame_download
    check_async_frame_download(lambda: Context(0))


def test_camera_ray_packets_with_axis_parallel_view(product_lib):
    """camera rays are traversed in packets of eight whose direction interval is tested against the boxes; looking straight down
    an axis puts zero inside that interval on two axes for the packets around the image centre (and on one axis along the
    centre row and column) — the hits must still be the per-ray ones, bit for bit"""
    desc = scenes.terrain(n=120, width=96, height=64, tex_size=64)
    for eye, look in (((0.0, 0.5, 16.0), (0.0, 0.5, 0.0)), ((0.0, 14.0, 0.0), (0.0, 0.0, 1e-3)), ((16.0, 2.0, 0.0), (0.0, 2.0, 0.0))):
        desc.camera = scenes.Camera(eye=eye, lookat=look, fovy=40.0)
        for packets in (1, 0):
            with Context(0) as e:
                o = build_pair(desc, e, {"packets": packets})
                check_render(e, o, desc, subframes=2)


@pytest.mark.parametrize("name", ["cornell", "terrain", "instanced"])
def test_many_subframes_in_flight(product_lib, name):
    """consecutive subframes overlap on the GPU (two pool slots, alternating stream pairs; only the resolves stay in order):
    eight of them launched back to back, the running-mean film compared with the oracle's bit for bit — with and without
    the pipeline, and with a stats / frame read in the middle"""
    desc = SMALL[name]()
    for pipeline in (1, 0):
        with Context(0) as e:
            o = build_pair(desc, e, {"pipeline": pipeline})
            check_render(e, o, desc, subframes=8)
            uvw = e.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
            from rendertoy3c_b200.api import make_settings
            frames = []
            for sf in range(5):
                e.launch_subframe(make_settings(desc, uvw, sf))
                if sf == 2:
                    frames.append(e.download_frame())
                    assert e.stats()["error_flags"] == 0
            frames.append(e.download_frame())
            assert not np.array_equal(frames[0], frames[1])
