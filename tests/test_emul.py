"""CPU-side coverage of the kernel logic: the kernel bodies of rendertoy3c_b200/csrc run through the
kernel-logic simulator (conftest.emul_lib) against the oracle.  These are NOT the parity tests proper
(those are tests/test_gpu_parity.py, on the B200); they keep the host logic, the BVH builder and the
wavefront schedule covered on a box without a GPU."""
import numpy as np
import pytest

from parity_common import SMALL, build_pair, check_render, check_trace, random_rays
from rendertoy3c_b200 import scenes
from rendertoy3c_b200.api import Context, camera_rays, make_settings


@pytest.mark.parametrize("name", sorted(SMALL))
def test_trace_and_render_match_oracle(emul_lib, name):
    desc = SMALL[name]()
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
        rays = np.concatenate([camera_rays(desc, uvw, 48, 48), random_rays(desc, 1500, 7)])
        check_trace(e, o, rays, accel=0)
        check_render(e, o, desc, subframes=2)


def test_unbounded_depth_matches_oracle(emul_lib):
    desc = SMALL["cornell"]()
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        check_render(e, o, desc, subframes=1, width=32, height=32, max_depth=0)


def test_api_errors(emul_lib):
    from rendertoy3c_b200.api import Rt3Error
    with Context(0, lib_path=emul_lib) as e:
        with pytest.raises(Rt3Error):
            e.accel_build()  # no instances
        with pytest.raises(Rt3Error):
            e.append_instance(5, np.zeros(12, np.float32))  # bad handle
        with pytest.raises(Rt3Error):
            e.mesh_create(np.zeros((3, 3)), np.array([[0, 1, 7]]), np.zeros((3, 3)), np.zeros((3, 2)))  # index out of range
        with pytest.raises(Rt3Error):
            e.texture_create(np.zeros((4, 4, 4), np.uint8), 0, 2)  # filter 0 (point) or 1 (bilinear)
        with pytest.raises(Rt3Error):
            e.texture_create(np.zeros((4, 4, 4), np.uint8), 4, 0)  # address modes 0..3


@pytest.mark.parametrize("name", sorted(SMALL))
def test_committed_golden_vectors(emul_lib, name):
    from parity_common import check_golden
    with Context(0, lib_path=emul_lib) as e:
        check_golden(e, name)


@pytest.mark.parametrize("offset", [0.0, 4096.0])
def test_adversarial_geometry_bit_exact(emul_lib, offset):
    import adversarial
    desc, off = adversarial.make_scene(offset)
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        ho = check_trace(e, o, adversarial.make_rays(off), accel=0)
        assert (ho["prim"] >= 0).mean() > 0.1
        check_render(e, o, desc, subframes=1)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", ["cornell", "terrain", "motion"])
def test_corrected_mode_matches_oracle(emul_lib, name, mode):
    desc = SMALL[name]()
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        check_render(e, o, desc, subframes=2, mode=mode)


def test_stacked_layers_ties(emul_lib):
    """the simulator's per-lane loop on the stacked-layer scene (the GPU suite runs it through the deferred queue)"""
    import adversarial
    desc = adversarial.make_stacked_scene(layers=16)
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        ho = check_trace(e, o, adversarial.make_stacked_rays(n=1500), accel=0)
        assert (ho["prim"] >= 0).mean() > 0.5


@pytest.mark.parametrize("address,filt", [(2, 0), (3, 0), (0, 1), (1, 1), (2, 1), (3, 1)])
def test_texture_modes_match_oracle(emul_lib, address, filt):
    """every CUDATexture address / filter mode: the textured terrain (uv in [0,8]: wraps, mirrors, runs into the border) renders bit-identically"""
    desc = SMALL["terrain"]()
    for t in desc.textures:
        t.address, t.filter = address, filt
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        check_render(e, o, desc, subframes=1)


def test_texcoord_transform_matches_oracle(emul_lib):
    """sampleTexture's scale / rotation / offset (cuda/LocalShading.h:37-54) on the textured terrain"""
    desc = SMALL["terrain"]()
    for inst in desc.instances:
        if inst.tex >= 0:
            inst.tex_xform = ((1.5, 0.75), (float(np.sin(0.6)), float(np.cos(0.6))), (0.125, -0.3))
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        check_render(e, o, desc, subframes=1)


@pytest.mark.parametrize("mode", [1, 2])
def test_corrected_mode_with_emitters_nee_cannot_reach(emul_lib, mode):
    """every instance of the motion scene emits — spheres, curves, animated and transformed meshes.  Their BSDF-sampled
    hits keep the full weight (the light list only holds object-space triangles of identity meshes, Q15); none of them
    may be looked up as a static triangle mesh (null vertex arrays on spheres / curves)"""
    desc = SMALL["motion"]()
    for k, inst in enumerate(desc.instances):
        inst.emission = (0.5 + 0.25 * (k % 3), 0.4, 0.3)
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        check_render(e, o, desc, subframes=2, mode=mode)


def test_stack_overflow_is_reported(tmp_path):
    """a traversal stack that overflows drops a subtree: rt3_trace and the download calls must say so instead of returning
    an incomplete result (simulator built with a 2-entry stack; the product's holds 64)"""
    import os
    import subprocess
    from rendertoy3c_b200.api import Rt3Error
    from rendertoy3c_b200 import scenes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / "librt3_emul_smallstack.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared", "-DRT3_EMULATE", "-DRT3_STACK_SIZE=2",
                    "-x", "c++", os.path.join(root, "rendertoy3c_b200", "csrc", "rt3_lib.cu"), "-I/usr/local/cuda/include", "-o", lib], check=True)
    desc = SMALL["instanced"]()
    with Context(0, lib_path=lib) as e:
        scenes.replay(desc, e)
        uvw = e.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, 1.0)
        rays = camera_rays(desc, uvw, 32, 32)
        with pytest.raises(Rt3Error, match="stack overflowed"):
            e.trace(rays)
        assert e.stats()["error_flags"] & 1
        e.reset_stats()
        assert e.stats()["error_flags"] == 0


def test_spline_tessellation_tolerance(emul_lib):
    """adaptive split of spline segments: hits stay within the stated tolerance of the true curve (parity_common)"""
    from parity_common import check_spline_tessellation
    with Context(0, lib_path=emul_lib) as e:
        check_spline_tessellation(e)


@pytest.mark.parametrize("options", [{"split": 2}, {"split": 0}, {"flatten": 0}, {"flatten": 0, "split": 2}, {"merge_identity": 0, "flatten": 0}],
                         ids=lambda o: ",".join("%s=%d" % kv for kv in o.items()))
@pytest.mark.parametrize("name", ["instanced", "motion", "deforming", "splines"])
def test_flatten_and_split_variants_match_oracle(emul_lib, name, options):
    """how static instances reach the kernels — flattened into the merged world BLAS or entered through the TLAS, in one
    launch or two (single-level kernel, then the rest seeded with its result) — changes nothing the oracle can see;
    `flatten` itself changes the hit arithmetic (world-space vertices) and is mirrored by the oracle"""
    desc = SMALL[name]()
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e, options)
        uvw = e.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
        rays = np.concatenate([camera_rays(desc, uvw, 40, 24), random_rays(desc, 1500, seed=5)])
        check_trace(e, o, rays)
        check_render(e, o, desc, subframes=1)


def test_collapsed_instance_is_not_flattened(emul_lib):
    """an instance whose transform is singular (a mesh squashed into a plane) or not finite keeps its object-space semantics in
    the kernels and in the oracle alike; its neighbours are flattened as usual"""
    desc = SMALL["instanced"]()
    m = np.array(desc.instances[0].xform, np.float32).reshape(3, 4)
    m[:, 2] = 0.0                      # rank 2
    desc.instances[0].xform = m.reshape(12)
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e)
        assert e.stats()["flattened_instances"] == sum(1 for i in desc.instances[1:] if desc.geoms[i.geom].kind == "mesh" and i.keys is None and not np.array_equal(np.asarray(i.xform, np.float32), scenes.IDENTITY))
        uvw = e.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
        check_trace(e, o, np.concatenate([camera_rays(desc, uvw, 40, 24), random_rays(desc, 1500, seed=6)]))


def test_unbounded_depth_running_out_of_slots_is_reported(emul_lib):
    """max_depth <= 0 means "until the path dies" (the reference has no bound); the library has 1022 depth slots.  Inside a
    closed white box no path dies: the launch ends them at the last slot and says so in error_flags bit 1"""
    s = 1.0
    q = [[[-s, -s, -s], [s, -s, -s], [s, -s, s], [-s, -s, s]], [[-s, s, -s], [-s, s, s], [s, s, s], [s, s, -s]],
         [[-s, -s, -s], [-s, -s, s], [-s, s, s], [-s, s, -s]], [[s, -s, -s], [s, s, -s], [s, s, s], [s, -s, s]],
         [[-s, -s, -s], [-s, s, -s], [s, s, -s], [s, -s, -s]], [[-s, -s, s], [s, -s, s], [s, s, s], [-s, s, s]]]
    box = scenes._quad_mesh(q)
    lamp = scenes._quad_mesh([[[-0.2, 0.99, -0.2], [0.2, 0.99, -0.2], [0.2, 0.99, 0.2], [-0.2, 0.99, 0.2]]])
    desc = scenes.SceneDesc("closed_box", [box, lamp], [scenes.Instance(0, diffuse=(1.0, 1.0, 1.0)), scenes.Instance(1, diffuse=(1.0, 1.0, 1.0), emission=(1.0, 1.0, 1.0))],
                            [], scenes.Camera(eye=(0.0, 0.0, 0.5), lookat=(0.0, 0.0, -1.0), fovy=60.0), 8, 8, 1, 0)
    with Context(0, lib_path=emul_lib) as e:
        scenes.replay(desc, e)
        uvw = e.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, 1.0)
        e.launch_subframe(make_settings(desc, uvw, 0, samples_per_launch=1, max_depth=0))
        e.sync()
        st = e.stats()
        assert st["error_flags"] & 2 and not st["error_flags"] & 1
        assert st["rays_bounce"] > 1000
        e.download_accum()          # informative only: the image is still handed out
        e.reset_stats()
        assert e.stats()["error_flags"] == 0


@pytest.mark.parametrize("split", [0, 2])
def test_shadow_rays_with_negative_tmax(emul_lib, split):
    from parity_common import millimetre_scene
    desc = millimetre_scene()
    with Context(0, lib_path=emul_lib) as e:
        o = build_pair(desc, e, {"split": split})
        check_render(e, o, desc, subframes=2)
        assert e.stats()["rays_shadow"] > 100


def test_async_frame_download(emul_lib):
    from parity_common import check_async_frame_download
    check_async_frame_download(lambda: Context(0, lib_path=emul_lib))
