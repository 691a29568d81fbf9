"""The C++ host (rendertoy3c_b200/host: RAII mirrors of the reference's operators, own .obj/.mtl/PPM
loader, headless wavefront main) renders the same frame as the Python host for the same scene.
CPU: linked against the kernel-logic simulator; GPU: the product binary against librt3.so."""
import os
import subprocess

import numpy as np
import pytest

from rendertoy3c_b200 import scenes
from rendertoy3c_b200.api import Context, make_settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rendertoy3c_b200", "host")


def build_host(libdir, libname, out):
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", out, os.path.join(HOST, "wavefront.cpp"), "-L" + libdir, "-l:" + libname,
                    "-Wl,-rpath," + libdir], check=True)
    return out


def read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = (int(x) for x in f.readline().split())
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(h, w, 3)


def run_case(tmp_path, exe, lib_path, desc, spp=16, tex_format="ppm"):
    obj = str(tmp_path / "scene.obj")
    scenes.write_obj(desc, obj, tex_format=tex_format)
    out = str(tmp_path / "out.ppm")
    c = desc.camera
    cmd = [exe, "--scene", obj, "--width", str(desc.width), "--height", str(desc.height), "--spp", str(spp), "--max-depth", str(desc.max_depth),
           "--fovy", repr(c.fovy), "--out", out, "--eye", *map(repr, c.eye), "--lookat", *map(repr, c.lookat), "--up", *map(repr, c.up)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    img = read_ppm(out)
    with Context(0, lib_path=lib_path) as g:
        scenes.replay(desc, g)
        uvw = g.camera_uvw(c.eye, c.lookat, c.up, c.fovy, desc.width / desc.height)
        for sf in range(spp // 8):
            g.launch_subframe(make_settings(desc, uvw, sf))
        ref = g.download_frame()[::-1, :, :3]
    assert np.array_equal(img, ref), "C++ host frame differs from the Python host frame"


def test_cpp_host_cornell_and_textured_terrain_simulator(tmp_path, emul_lib):
    exe = build_host(os.path.dirname(emul_lib), os.path.basename(emul_lib), str(tmp_path / "wavefront_emul"))
    run_case(tmp_path, exe, emul_lib, scenes.cornell(width=48, height=48))
    run_case(tmp_path, exe, emul_lib, scenes.terrain(n=12, width=48, height=32, tex_size=16))
    run_case(tmp_path, exe, emul_lib, scenes.terrain(n=12, width=48, height=32, tex_size=32), tex_format="png")   # map_Kd .png through the own PNG decoder


@pytest.mark.gpu
def test_cpp_host_on_gpu(tmp_path, product_lib):
    exe = build_host(os.path.dirname(product_lib), os.path.basename(product_lib), str(tmp_path / "wavefront"))
    run_case(tmp_path, exe, None, scenes.cornell(width=128, height=128))
    run_case(tmp_path, exe, None, scenes.terrain(n=32, width=128, height=72, tex_size=64))
