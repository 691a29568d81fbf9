"""CORRECTED mode checks shared by the CPU and GPU suites: an analytic direct-lighting
case (point-to-rectangle configuration factor) and bit-exact parity with the oracle."""
import numpy as np

from rendertoy3c_b200 import scenes
from rendertoy3c_b200.api import make_settings
from rendertoy3c_b200.scenes import Camera, Instance, SceneDesc, _quad_mesh


def furnace_scene(width=24, height=24):
    """diffuse floor (albedo 0.5) at y=0, 2x2 emissive quad (Le=10) centred 1 above the origin, narrow camera
    looking at the origin from the side: outgoing floor radiance = albedo * Le * F, F = 4 * corner factor."""
    floor = _quad_mesh([[[-50, 0, -50], [-50, 0, 50], [50, 0, 50], [50, 0, -50]]])
    light = _quad_mesh([[[-1, 1, -1], [1, 1, -1], [1, 1, 1], [-1, 1, 1]]])
    inst = [Instance(0, diffuse=(0.5, 0.5, 0.5)), Instance(1, diffuse=(0.0, 0.0, 0.0), emission=(10.0, 10.0, 10.0))]
    cam = Camera(eye=(4.0, 0.6, 0.0), lookat=(0.0, 0.0, 0.0), fovy=1.5)
    return SceneDesc("furnace", [floor, light], inst, [], cam, width, height, 8, 2)


def analytic_radiance(albedo=0.5, le=10.0, a=1.0, b=1.0, h=1.0):
    X, Y = a / h, b / h
    corner = (X / np.sqrt(1 + X * X) * np.arctan(Y / np.sqrt(1 + X * X)) + Y / np.sqrt(1 + Y * Y) * np.arctan(X / np.sqrt(1 + Y * Y))) / (2 * np.pi)
    return albedo * le * 4 * corner


def _corner(X, Y):
    """configuration factor from a surface point to a rectangle X x Y (in units of its height) with one corner on the point's normal"""
    return (X / np.sqrt(1 + X * X) * np.arctan(Y / np.sqrt(1 + X * X)) + Y / np.sqrt(1 + Y * Y) * np.arctan(X / np.sqrt(1 + Y * Y))) / (2 * np.pi)


def two_light_scene(width=24, height=24):
    """furnace_scene plus a second, dim emitter (Le=1) of the same size next to the first (x in [1,3]): the power
    sampler (mode 2) picks it 1 time in 11, the uniform sampler every other time; both must converge to the same value"""
    d = furnace_scene(width, height)
    dim = _quad_mesh([[[1, 1, -1], [3, 1, -1], [3, 1, 1], [1, 1, 1]]])
    d.geoms.append(dim)
    d.camera.fovy = 0.5  # the dim emitter is off-centre: its factor varies linearly across the footprint, keep the footprint small
    d.instances.append(Instance(2, diffuse=(0.0, 0.0, 0.0), emission=(1.0, 1.0, 1.0)))
    return d


def analytic_two_lights(albedo=0.5, le_a=10.0, le_b=1.0):
    return albedo * (le_a * 4 * _corner(1.0, 1.0) + le_b * 2 * (_corner(3.0, 1.0) - _corner(1.0, 1.0)))


def render_mean(backend, desc, subframes, mode, max_depth):
    uvw = backend.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
    backend.clear_accum() if hasattr(backend, "clear_accum") else None
    for sf in range(subframes):
        backend.launch_subframe(make_settings(desc, uvw, sf, mode=mode, max_depth=max_depth, miss=0.0))
    return float(backend.download_accum()[..., 0].mean())
