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
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", out, os.path.join(HOST, "wavefront.cpp"), "-L" + libdir, "-l:" + libname,
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


def _read_png_rgba(path):
    import struct
    import zlib
    d = open(path, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(d):
        n, tag = struct.unpack(">I4s", d[pos:pos + 8])
        body = d[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", d[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF, "chunk CRC"
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, il = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, il) == (8, 6, 0)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)   # zlib verifies the Adler-32
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 4)


def _read_exr_rgba(path):
    import struct
    d = open(path, "rb").read()
    assert d[:8] == bytes([0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0])
    pos, attrs = 8, {}
    while d[pos] != 0:
        e = d.index(b"\0", pos); name = d[pos:e].decode(); pos = e + 1
        e = d.index(b"\0", pos); typ = d[pos:e].decode(); pos = e + 1
        n = struct.unpack("<i", d[pos:pos + 4])[0]; pos += 4
        attrs[name] = (typ, d[pos:pos + n]); pos += n
    pos += 1
    assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names = [c.split(b"\0")[0].decode() for c in [attrs["channels"][1][i * 18:(i + 1) * 18] for i in range(4)]]
    assert names == ["A", "B", "G", "R"]
    offs = struct.unpack("<%dQ" % h, d[pos:pos + 8 * h])
    img = np.empty((h, w, 4), np.float32)
    for y in range(h):
        yy, nb = struct.unpack("<ii", d[offs[y]:offs[y] + 8])
        assert yy == y and nb == 16 * w
        line = np.frombuffer(d[offs[y] + 8:offs[y] + 8 + nb], dtype="<f4").reshape(4, w)
        img[y] = line[[3, 2, 1, 0]].T
    return img


def test_cpp_host_output_formats(tmp_path, emul_lib):
    """N3 (image output): .png and .exr written by host/image_writer.hpp parse with the standard library and hold the
    frame / the float accumulation buffer exactly; --tonemap aces applies the reference viewer's display curve"""
    exe = build_host(os.path.dirname(emul_lib), os.path.basename(emul_lib), str(tmp_path / "wavefront_emul"))
    desc = scenes.cornell(width=40, height=28)
    obj = str(tmp_path / "scene.obj")
    scenes.write_obj(desc, obj)
    c = desc.camera
    base = [exe, "--scene", obj, "--width", str(desc.width), "--height", str(desc.height), "--spp", "16", "--max-depth", str(desc.max_depth),
            "--fovy", repr(c.fovy), "--eye", *map(repr, c.eye), "--lookat", *map(repr, c.lookat), "--up", *map(repr, c.up)]
    with Context(0, lib_path=emul_lib) as g:
        scenes.replay(desc, g)
        uvw = g.camera_uvw(c.eye, c.lookat, c.up, c.fovy, desc.width / desc.height)
        for sf in range(2):
            g.launch_subframe(make_settings(desc, uvw, sf))
        frame, accum = g.download_frame()[::-1], g.download_accum()[::-1]
    for ext in ("png", "exr", "ppm"):
        out = str(tmp_path / ("o." + ext))
        r = subprocess.run(base + ["--out", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        if ext == "png":
            assert np.array_equal(_read_png_rgba(out), frame)
        elif ext == "exr":
            assert np.array_equal(_read_exr_rgba(out).view(np.uint32), np.ascontiguousarray(accum).view(np.uint32))
        else:
            assert np.array_equal(read_ppm(out), frame[..., :3])
    out = str(tmp_path / "aces.ppm")
    r = subprocess.run(base + ["--out", out, "--tonemap", "aces"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    x = frame[..., :3].astype(np.float32) / np.float32(255)
    want = np.clip((x * (2.51 * x + 0.03)) / (x * (2.43 * x + 0.59) + 0.14), 0, 1)
    assert np.abs(read_ppm(out).astype(np.int32) - np.rint(want * 255).astype(np.int32)).max() <= 1
    r = subprocess.run(base + ["--out", str(tmp_path / "o.gif")], capture_output=True, text=True)
    assert r.returncode != 0 and "unsupported extension" in r.stderr


def test_cpp_host_vertex_keyframes_from_obj_files(tmp_path, emul_lib):
    """N1 + N2: several .obj files of one topology are key-frames (src/mesh.cpp:39-110): `--scene k0.obj --key k1.obj
    --key k2.obj` renders the same frame as the Python host fed with the [keys, nv, 3] vertex arrays"""
    from rendertoy3c_b200.scenes import Camera, Geometry, Instance, SceneDesc, _blob_pos, _quad_mesh, grid_mesh
    exe = build_host(os.path.dirname(emul_lib), os.path.basename(emul_lib), str(tmp_path / "wavefront_emul"))
    blob = grid_mesh(10, 10, _blob_pos)
    nk = 3
    vk = np.stack([blob.verts * np.float32(1.0 + 0.25 * k) + np.array([0.15 * k, 0.1 * k * k, 0.0], np.float32) for k in range(nk)]).astype(np.float32)
    geoms = [Geometry("mesh", verts=np.ascontiguousarray(vk[0]), idx=blob.idx, normals=blob.normals, uvs=blob.uvs, vert_keys=vk)]
    for quad in ([[-2, 2.5, -2], [2, 2.5, -2], [2, 2.5, 2], [-2, 2.5, 2]], [[-4, -0.8, -4], [-4, -0.8, 4], [4, -0.8, 4], [4, -0.8, -4]]):
        q = _quad_mesh([quad])
        q.vert_keys = np.stack([q.verts] * nk)                       # every mesh of a key-framed scene carries all keys, like the reference's
        geoms.append(q)
    inst = [Instance(0, diffuse=(0.8, 0.4, 0.3)), Instance(1, diffuse=(0.8, 0.8, 0.8), emission=(8.0, 8.0, 7.0)), Instance(2, diffuse=(0.6, 0.6, 0.6))]
    desc = SceneDesc("keyframes", geoms, inst, [], Camera(eye=(0.7, 0.8, 3.6), lookat=(0.4, 0.1, 0.0), fovy=45.0), 48, 32, 16, 4)
    files = []
    for k in range(nk):
        files.append(str(tmp_path / ("key%d.obj" % k)))
        scenes.write_obj(desc, files[-1], key=k)
    out = str(tmp_path / "out.ppm")
    c = desc.camera
    cmd = [exe, "--scene", files[0], "--key", files[1], "--key", files[2], "--width", str(desc.width), "--height", str(desc.height), "--spp", "16",
           "--max-depth", str(desc.max_depth), "--fovy", repr(c.fovy), "--out", out, "--eye", *map(repr, c.eye), "--lookat", *map(repr, c.lookat), "--up", *map(repr, c.up)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with Context(0, lib_path=emul_lib) as g:
        scenes.replay(desc, g)
        uvw = g.camera_uvw(c.eye, c.lookat, c.up, c.fovy, desc.width / desc.height)
        for sf in range(2):
            g.launch_subframe(make_settings(desc, uvw, sf))
        ref = g.download_frame()[::-1, :, :3]
        static = None
    assert np.array_equal(read_ppm(out), ref), "key-framed C++ host frame differs from the Python host frame"
    with Context(0, lib_path=emul_lib) as g:                         # and the motion is really there: key 0 alone renders something else
        for gm in desc.geoms:
            gm.vert_keys = None
        scenes.replay(desc, g)
        for sf in range(2):
            g.launch_subframe(make_settings(desc, uvw, sf))
        static = g.download_frame()[::-1, :, :3]
    assert not np.array_equal(static, ref)
    # a key-frame file of another topology is refused
    bad = str(tmp_path / "bad.obj")
    open(bad, "w").write("v 0 0 0\nvn 0 1 0\nvt 0 0\n")
    r = subprocess.run([exe, "--scene", files[0], "--key", bad, "--out", out], capture_output=True, text=True)
    assert r.returncode != 0 and "fewer v / vn / vt" in r.stderr


@pytest.mark.gpu
def test_cpp_host_two_gpus_one_process(tmp_path, product_lib):
    """two contexts in one process (every entry point selects its context's device), sample-partitioned subframes and the
    NCCL sum of rt3_allreduce_accum: same image as one GPU.  Needs a 2-GPU box (skipped otherwise)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = build_host(os.path.dirname(product_lib), os.path.basename(product_lib), str(tmp_path / "wavefront"))
    desc = scenes.cornell(width=128, height=128)
    obj = str(tmp_path / "scene.obj")
    scenes.write_obj(desc, obj)
    c = desc.camera
    imgs = []
    for n in (1, 2):
        out = str(tmp_path / ("o%d.ppm" % n))
        r = subprocess.run([exe, "--scene", obj, "--gpus", str(n), "--width", "128", "--height", "128", "--spp", "64", "--max-depth", "4", "--fovy", repr(c.fovy),
                            "--out", out, "--eye", *map(repr, c.eye), "--lookat", *map(repr, c.lookat), "--up", *map(repr, c.up)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        imgs.append(read_ppm(out).astype(np.int32))
    assert np.abs(imgs[0] - imgs[1]).max() <= 1   # same samples, another summation order
