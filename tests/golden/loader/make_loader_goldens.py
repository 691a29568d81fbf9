"""Writes the loader-parity fixtures (tests/golden/loader/cases/<case>/*.obj, *.mtl, textures) and, when the
REFERENCE'S OWN loader has been compiled here (oracle/Makefile target `ref_loader` -> oracle/_ref/dump_ref_loader, i.e.
/root/reference/src/mesh.cpp + its vendored tinyobj / stb_image), the golden dumps <case>/golden.rt3l it produces for
them.  tests/test_loader_parity.py then requires this repo's loadOBJ to reproduce every golden byte for byte.

Needs Pillow (for the JPEG / BMP / TGA / PNM encoders) and is run by hand in the build container:
    make -C oracle ref_loader && python tests/golden/loader/make_loader_goldens.py
Deterministic: fixed seeds, no timestamps.  Fixtures and goldens are committed; /root/reference is not needed to run
the tests (it is used, when present, to re-derive the goldens and to fuzz)."""
import math
import os
import shutil
import struct
import subprocess
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CASES = os.path.join(HERE, "cases")
REF_DUMPER = os.path.join(ROOT, "oracle", "_ref", "dump_ref_loader")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def picture(w, h, seed=0):
    """a smooth + noisy RGB test picture (uint8 [h, w, 3])"""
    rng = np.random.RandomState(100 + seed)
    y, x = np.mgrid[0:h, 0:w]
    a = np.stack([128 + 100 * np.sin(x / 5.0 + y / 9.0 + seed), 128 + 100 * np.cos(x / 3.0 - y / 7.0), (x * 7 + y * 13 + 31 * seed) % 256], -1)
    a = a + rng.randint(-20, 20, a.shape)
    return np.clip(a, 0, 255).astype(np.uint8)


def tri_scene(d, textures, mtl_extra=""):
    """one triangle per texture file, one material each: scene.obj + scene.mtl in directory d"""
    with open(os.path.join(d, "scene.mtl"), "w") as m, open(os.path.join(d, "scene.obj"), "w") as o:
        o.write("mtllib scene.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nvt 1 0\nvt 0 1\n")
        for i, fn in enumerate(textures):
            m.write("newmtl m%d\nKd 1 1 1\nmap_Kd %s\n" % (i, fn))
            o.write("o s%d\nusemtl m%d\nf 1/1/1 2/2/1 3/3/1\n" % (i, i))
        m.write(mtl_extra)


def case_dir(name):
    d = os.path.join(CASES, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    return d


def write(path, text, eol="\n"):
    with open(path, "w", newline="") as f:
        f.write(text.replace("\n", eol))


# ------------------------------------------------------------------------------------------------ geometry cases
def case_polygons():
    d = case_dir("polygons")
    write(os.path.join(d, "scene.mtl"), "newmtl a\nKd 0.8 0.1 0.1\nnewmtl b\nKd 0.1 0.1 0.8\nKe 1 2 3\n")
    L = ["mtllib scene.mtl", "vn 0 0 1", "vn 0 1 0", "vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1", "usemtl a"]
    nv = 0

    def poly(points, reverse=False):
        nonlocal nv
        for p in points:
            L.append("v %.9g %.9g %.9g" % tuple(p))
        ids = list(range(nv + 1, nv + 1 + len(points)))
        nv += len(points)
        if reverse:
            ids.reverse()
        L.append("f " + " ".join("%d/%d/1" % (i, 1 + k % 4) for k, i in enumerate(ids)))

    rng = np.random.RandomState(5)
    poly([(0, 0, 0), (1, 0, 0), (0, 1, 0)])                                           # triangle
    poly([(0, 0, 0), (2, 0, 0), (2, 1, 0), (0, 1, 0)])                                # planar quad, diagonals equal
    poly([(0, 0, 0), (3, 0, 0.2), (3.5, 1, -0.3), (0, 1.5, 0.4)])                     # non-planar quad, 0-2 shorter
    poly([(0, 0, 0), (1, 0, 0.2), (4, 3, -0.3), (0, 1, 0.4)])                         # non-planar quad, 1-3 shorter
    for n in (5, 6, 7, 9, 12):                                                        # convex n-gons in the three axis planes
        for ax in range(3):
            pts = []
            for k in range(n):
                a = 2 * math.pi * k / n
                p = [math.cos(a) * (1 + 0.1 * ax), math.sin(a)]
                p.insert(ax, 0.25 * ax)
                pts.append(p)
            poly(pts, reverse=(n + ax) % 2 == 1)
    L.append("usemtl b")
    for n in (5, 6, 8, 11):                                                           # concave (star-like) n-gons, both windings, random planes
        for rep in range(3):
            pts = []
            u = rng.randn(3); u /= np.linalg.norm(u)
            v = np.cross(u, rng.randn(3)); v /= np.linalg.norm(v)
            for k in range(n):
                a = 2 * math.pi * k / n
                r = rng.uniform(0.25, 1.0)
                pts.append(tuple(r * math.cos(a) * u + r * math.sin(a) * v + 0.02 * rng.randn(3) * (rep == 2)))
            poly(pts, reverse=rep == 1)
    poly([(0, 0, 0), (1, 0, 0), (2, 0, 0), (2, 1, 0), (1, 1, 0), (0, 1, 0)])          # collinear edges (degenerate first corners)
    poly([(0, 0, 0), (1, 1, 0), (2, 2, 0), (3, 3, 0), (4, 4, 0)])                     # a fully degenerate pentagon
    poly([(0, 0, 0), (2, 0, 0), (2, 2, 0), (1, 0.2, 0), (0, 2, 0)])                   # an arrow: a reflex vertex inside the first candidate ears
    L.append("f 1/1/1 2/2/1")                                                         # a 2-gon: dropped
    write(os.path.join(d, "scene.obj"), "\n".join(L) + "\n")
    return d, ["scene.obj"]


def case_groups():
    d = case_dir("groups")
    write(os.path.join(d, "scene.mtl"),
          "# materials\nnewmtl red\nKd 0.8 0.1 0.1\nPr 0.3\naniso 0.2\nNi 1.5\nTf 0.5 0.4 0.3\n\nnewmtl blue\n\tKd 0.1 0.1 0.8\nKe 1 2 3\nKt 0.25 0.5 0.75\n"
          "newmtl red\nKd 0 1 0\nnewmtl green\nKd 0 0.5 0\n")   # the second `red` never wins
    L = ["# comment", "mtllib scene.mtl", ""]
    rng = np.random.RandomState(11)
    for i in range(24):
        L.append("v %.9g %.9g %.9g" % tuple(rng.uniform(-2, 2, 3)))
    for i in range(6):
        L.append("vn %.9g %.9g %.9g" % tuple(rng.uniform(-1, 1, 3)))
    for i in range(8):
        L.append("vt %.9g %.9g" % tuple(rng.uniform(0, 4, 2)))
    L += ["usemtl red", "f 1/1/1 2/2/2 3/3/3", "f 2/2/2 3/3/3 4/4/4",        # shape 0 (no name): red, blue, red again, green
          "usemtl blue", "f 4/1/1 5/2/1 6/3/1 7/4/1", "usemtl red", "f 3/3/3 2/2/2 8/8/6",
          "usemtl green", "f -1/-1/-1 -2/-2/-2 -3/-3/-3",
          "o second", "f 9/1/1 10/1/1 11/1/1",                                 # material carries over into the next object (green)
          "usemtl blue", "f 9/1/1 10/1/1 12/1/1", "f 9/2/1 10/1/1 12/1/1",       # same position, other vt: a new vertex
          "g grp one", "usemtl blue", "f 13/1/1 14/1/1 15/1/1",
          "g", "g empty", "usemtl red", "g third", "f 16/5/2 17/6/2 18/7/2 19/8/2 20/1/2",
          "o", "usemtl nosuch"]                                                 # unknown material with no faces after it: harmless
    write(os.path.join(d, "scene.obj"), "\n".join(L))                           # no newline at the end of the file
    return d, ["scene.obj"]


def case_syntax():
    d = case_dir("syntax")
    write(os.path.join(d, "scene.mtl"), "newmtl m \nKd 1 .5 5e-1\nKe +1.5e+0 -.25 1E1 \n", eol="\r\n")
    L = ["mtllib   scene.mtl", "v 0.1 0.2 0.3", "v 1e0 -1E-1 +2.5e+0", "v .5 -.5 5.", "v 0.12345678901234567 123456789.123456789 -0.000001234567",
         "v 1 2", "v 7 8 9 1.0", "v 3 4 5 0.1 0.2 0.3", "v nan inf -inf", "v 1e 2e+ 3.e1", "v 0.30000001192092896 16777217 1e-45",
         "v  \t 4.0\t5.0   6.0  ", "vn 0 0 1", "vn 1e-3 .7 -.7", "vt 0.25", "vt 1 2 3", "vt .5 .75",
         "usemtl m", "f 1/1/1 2/2/2 3/3/1", "f\t4/1/1\t5/2/2\t6/3/1 ", "f 7//1 8//2 9//1", "f 9/1 10/2 11/3", "f 1 2 3", "f 1/1/1 2/2 3",
         "f 1/0/1 2/0/0 3/1/1", "f 4/3/2 5/3/2 6/3/2 7/3/2", "s 1", "s off", "l 1 2 3", "p 1", "vp 0.5", "usemtlx", "t foo 0/0/0"]
    write(os.path.join(d, "scene.obj"), "\n".join(L) + "\n", eol="\r\n")
    return d, ["scene.obj"]


def case_keyframes():
    d = case_dir("keyframes")
    write(os.path.join(d, "k0.mtl"), "newmtl a\nKd 0.5 0.5 0.5\nnewmtl b\nKd 0.9 0.2 0.2\n")
    rng = np.random.RandomState(21)
    V = rng.uniform(-1, 1, (3, 10, 3))
    N = rng.uniform(-1, 1, (3, 4, 3))
    T = rng.uniform(0, 1, (3, 5, 2))
    faces = "usemtl a\nf 1/1/1 2/2/2 3/3/3\nf 3/3/3 2/2/2 4/4/4 5/5/1\nusemtl b\nf 6/1/1 7/2/1 8/3/1\no two\nf 8/1/2 9/2/2 10/3/2 1/4/2 2/5/2\n"
    for k in range(3):
        s = "mtllib k0.mtl\n" if k == 0 else "mtllib none.mtl\n"
        s += "".join("v %.9g %.9g %.9g\n" % tuple(v) for v in V[k]) + "".join("vn %.9g %.9g %.9g\n" % tuple(v) for v in N[k])
        s += "".join("vt %.9g %.9g\n" % tuple(v) for v in T[k])
        s += faces if k < 2 else "# key 2 carries more vertices than the faces use and no faces of its own\nv 9 9 9\n"
        write(os.path.join(d, "k%d.obj" % k), s)
    return d, ["k0.obj", "k1.obj", "k2.obj"]


def case_scenes():
    """the repo's own exporter: Cornell box + a textured terrain (PPM and PNG textures)"""
    from rendertoy3c_b200 import scenes
    d = case_dir("scenes")
    scenes.write_obj(scenes.cornell(width=16, height=16), os.path.join(d, "cornell.obj"))
    scenes.write_obj(scenes.terrain(n=6, width=16, height=16, tex_size=16), os.path.join(d, "terrain.obj"), tex_format="png")
    return d, None   # two independent scenes: see run()


# ------------------------------------------------------------------------------------------------ .mtl / texture-list cases
def case_mtl_textures():
    from PIL import Image
    d = case_dir("mtl_textures")
    os.makedirs(os.path.join(d, "sub dir"))
    Image.fromarray(picture(9, 7, 1)).save(os.path.join(d, "a.png"))
    Image.fromarray(picture(5, 6, 2)).save(os.path.join(d, "sub dir", "b c.png"))
    Image.fromarray(picture(4, 4, 3)).save(os.path.join(d, "e.png"))
    write(os.path.join(d, "scene.mtl"),
          "map_Kd a.png\n"                                   # statements before any newmtl belong to an unnamed material that is dropped
          "newmtl first\nmap_Kd a.png\n"                     # no Kd seen yet in this file: Kd becomes 0.6
          "newmtl second\nKd 0.25 0.5 0.75\nmap_Kd -s 2 2 2 -o 0.5 0.5 0 -clamp on sub dir\\b c.png\nmap_Ke e.png\nmap_Pr a.png\nnorm -bm 0.5 a.png\n"
          "newmtl third\nmap_Kd a.png\n"                     # has_kd survives `newmtl`: Kd stays 0
          "newmtl fourth\nKd 1 1 1\nmap_Kd missing.png\nmap_Ke a.png\n"
          "newmtl fifth\nKd 1 0 1\nmap_Kd -blendu off\n")     # options only: no file name
    L = ["mtllib nothere.mtl scene.mtl", "v 0 0 0", "v 1 0 0", "v 0 1 0", "v 1 1 0", "vn 0 0 1", "vt 0 0", "vt 1 1"]
    for i, m in enumerate(("first", "second", "third", "fourth", "fifth", "first")):
        L += ["o s%d" % i, "usemtl %s" % m, "f 1/1/1 2/2/1 3/1/1", "f 2/2/1 4/2/1 3/1/1"]
    write(os.path.join(d, "scene.obj"), "\n".join(L) + "\n")
    return d, ["scene.obj"]


def _png_variants(d):
    from test_image_loader import write_png, write_png_adam7
    rng = np.random.RandomState(31)
    names = []
    w, h = 11, 6
    for ctype, chan in ((0, 1), (2, 3), (4, 2), (6, 4)):
        for depth in (8, 16):
            px = rng.randint(0, 256, (h, w, chan * depth // 8)).astype(np.uint8)
            fn = "c%d_d%d.png" % (ctype, depth)
            write_png(os.path.join(d, fn), [px[y].tobytes() for y in range(h)], ctype, depth, [0, 1, 2, 3, 4], split=2)
            names.append(fn)
    for depth in (1, 2, 4):   # packed grey and palette rows
        nbytes = (w * depth + 7) // 8
        rows = [bytes(rng.randint(0, 256, nbytes).astype(np.uint8)) for _ in range(h)]
        write_png(os.path.join(d, "grey_d%d.png" % depth), rows, 0, depth, [0, 2])
        plte = rng.randint(0, 256, 3 * (1 << depth)).astype(np.uint8)
        write_png(os.path.join(d, "pal_d%d.png" % depth), rows, 3, depth, [0], plte=plte, trns=plte[:(1 << depth) - 1][::-1])
        names += ["grey_d%d.png" % depth, "pal_d%d.png" % depth]
    px = rng.randint(0, 200, (h, w, 1)).astype(np.uint8)
    plte = rng.randint(0, 256, 3 * 200).astype(np.uint8)
    write_png(os.path.join(d, "pal_d8.png"), [px[y].tobytes() for y in range(h)], 3, 8, [4], plte=plte)
    names.append("pal_d8.png")
    # colour-key transparency on grey / RGB (tRNS holds ONE colour, 16 bits per channel)
    g = rng.randint(0, 4, (h, w, 1)).astype(np.uint8) * 80
    write_png(os.path.join(d, "grey_key.png"), [g[y].tobytes() for y in range(h)], 0, 8, [1], trns=struct.pack(">H", 80))
    c = rng.randint(0, 2, (h, w, 3)).astype(np.uint8) * 255
    write_png(os.path.join(d, "rgb_key.png"), [c[y].tobytes() for y in range(h)], 2, 8, [2], trns=struct.pack(">HHH", 255, 0, 255))
    g16 = rng.randint(0, 3, (h, w)).astype(np.uint16) * 0x1234
    write_png(os.path.join(d, "grey16_key.png"), [g16[y].astype(">u2").tobytes() for y in range(h)], 0, 16, [0], trns=struct.pack(">H", 0x1234))
    names += ["grey_key.png", "rgb_key.png", "grey16_key.png"]
    for (hh, ww, chan, ctype) in ((9, 13, 3, 2), (1, 1, 4, 6), (5, 3, 1, 0), (8, 8, 2, 4)):
        px = rng.randint(0, 256, (hh, ww, chan)).astype(np.uint8)
        fn = "adam7_%dx%d_c%d.png" % (ww, hh, ctype)
        write_png_adam7(os.path.join(d, fn), px, ctype, [0, 1, 2, 3, 4])
        names.append(fn)
    return names


def case_textures_png():
    d = case_dir("textures_png")
    tri_scene(d, _png_variants(d))
    return d, ["scene.obj"]


def case_textures_other():
    from PIL import Image
    d = case_dir("textures_other")
    names = []
    rgb = picture(13, 9, 4)
    rgba = np.dstack([rgb, picture(13, 9, 5)[..., 0]])
    Image.fromarray(rgb).save(os.path.join(d, "rgb24.bmp")); names.append("rgb24.bmp")
    Image.fromarray(rgba).save(os.path.join(d, "rgba32.bmp")); names.append("rgba32.bmp")
    Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=16).save(os.path.join(d, "pal8.bmp")); names.append("pal8.bmp")
    Image.fromarray(rgb).convert("L").save(os.path.join(d, "grey8.bmp")); names.append("grey8.bmp")
    Image.fromarray(rgb).convert("1").save(os.path.join(d, "mono1.bmp")); names.append("mono1.bmp")
    # top-down 32-bit BMP written by hand (negative height)
    h, w = rgba.shape[:2]
    body = b"".join(rgba[y][:, [2, 1, 0, 3]].tobytes() for y in range(h))
    open(os.path.join(d, "topdown32.bmp"), "wb").write(b"BM" + struct.pack("<IHHI", 54 + len(body), 0, 0, 54) +
                                                         struct.pack("<IiiHHIIiiII", 40, w, -h, 1, 32, 0, len(body), 2835, 2835, 0, 0) + body)
    names.append("topdown32.bmp")
    Image.fromarray(rgb).save(os.path.join(d, "rgb24.tga")); names.append("rgb24.tga")
    Image.fromarray(rgba).save(os.path.join(d, "rgba32_rle.tga"), compression="tga_rle"); names.append("rgba32_rle.tga")
    Image.fromarray(rgb).convert("L").save(os.path.join(d, "grey8.tga")); names.append("grey8.tga")
    Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=32).save(os.path.join(d, "pal8.tga")); names.append("pal8.tga")
    flat = rgb.copy(); flat[:, 3:9] = flat[:, 3:4]
    Image.fromarray(flat).save(os.path.join(d, "rgb24_rle_top.tga"), compression="tga_rle", orientation=1); names.append("rgb24_rle_top.tga")
    open(os.path.join(d, "rgb.ppm"), "wb").write(b"P6\n# comment\n%d %d\n255\n" % (w, h) + rgb.tobytes()); names.append("rgb.ppm")
    open(os.path.join(d, "grey.pgm"), "wb").write(b"P5 %d %d 255\n" % (w, h) + rgb[..., 1].tobytes()); names.append("grey.pgm")
    open(os.path.join(d, "rgb_max100.ppm"), "wb").write(b"P6 %d %d 100\n" % (w, h) + (rgb // 3).tobytes()); names.append("rgb_max100.ppm")
    open(os.path.join(d, "rgb16.ppm"), "wb").write(b"P6 %d %d 65535\n" % (w, h) + (rgb.astype(">u2") * 256 + (255 - rgb)).astype(">u2").tobytes()); names.append("rgb16.ppm")
    open(os.path.join(d, "garbage.png"), "wb").write(b"\x89PNG\r\n\x1a\nnot really"); names.append("garbage.png")
    open(os.path.join(d, "empty.jpg"), "wb").write(b""); names.append("empty.jpg")
    tri_scene(d, names)
    return d, ["scene.obj"]


def case_textures_jpeg():
    from PIL import Image
    from test_image_loader import write_baseline_jpeg
    d = case_dir("textures_jpeg")
    names = []

    def save(fn, im, **kw):
        im.save(os.path.join(d, fn), "JPEG", **kw)
        names.append(fn)

    for i, (w, h) in enumerate(((37, 29), (16, 16), (1, 1), (17, 1), (2, 33), (8, 8))):
        im = Image.fromarray(picture(w, h, i))
        save("s444_%dx%d.jpg" % (w, h), im, subsampling=0, quality=85)
        save("s422_%dx%d.jpg" % (w, h), im, subsampling=1, quality=70)
        save("s420_%dx%d.jpg" % (w, h), im, subsampling=2, quality=90)
        save("grey_%dx%d.jpg" % (w, h), im.convert("L"), quality=80)
    im = Image.fromarray(picture(41, 27, 9))
    save("s420_q30.jpg", im, subsampling=2, quality=30)
    save("s420_opt.jpg", im, subsampling=2, quality=75, optimize=True)
    save("s420_rst3.jpg", im, subsampling=2, quality=75, restart_marker_blocks=3)
    save("s444_rstrow.jpg", im, subsampling=0, quality=75, restart_marker_rows=1)
    save("prog420.jpg", im, subsampling=2, quality=80, progressive=True)
    save("prog444.jpg", im, subsampling=0, quality=95, progressive=True)
    save("prog_grey.jpg", im.convert("L"), quality=80, progressive=True)
    save("cmyk.jpg", im.convert("CMYK"), quality=80)
    save("rgb_keep.jpg", im, quality=90, keep_rgb=True)
    # the repo's own test encoder: sampling factors PIL does not offer (4:4:0, 4:1:1, 4:1:0), flat Huffman tables, restart interval
    px = picture(29, 21, 12)
    for tag, sampling, kw in (("440", (1, 2), {}), ("411", (4, 1), {}), ("410", (4, 2), dict(restart=2)), ("420flat", (2, 2), dict(optimal=False))):
        fn = "own_%s.jpg" % tag
        write_baseline_jpeg(os.path.join(d, fn), px, sampling=sampling, **kw)
        names.append(fn)
    # a file cut inside its entropy-coded data (no restart markers): stb feeds zero bits from there on and still decodes
    # every block.  (With restart markers the rest of the planes stays uninitialised malloc memory in stb: not a fixture.)
    whole = open(os.path.join(d, "s420_q30.jpg"), "rb").read()
    sos = whole.index(b"\xff\xda")
    open(os.path.join(d, "cut.jpg"), "wb").write(whole[:sos + 14 + (len(whole) - sos - 14) // 2]); names.append("cut.jpg")
    tri_scene(d, names)
    return d, ["scene.obj"]


# ------------------------------------------------------------------------------------------------ GIF / PSD / PIC / HDR, written by hand
def gif_lzw(indices, min_bits):
    """GIF flavour of LZW (clear code first, variable width, LSB first) cut into sub-blocks"""
    clear, eoi = 1 << min_bits, (1 << min_bits) + 1
    out, acc, nbits = bytearray(), 0, 0

    def emit(code, width):
        nonlocal acc, nbits
        acc |= code << nbits
        nbits += width
        while nbits >= 8:
            out.append(acc & 255)
            acc >>= 8
            nbits -= 8

    table = {(i,): i for i in range(clear)}
    width, nxt = min_bits + 1, eoi + 1
    emit(clear, width)
    cur = ()
    for px in indices:
        if cur + (px,) in table:
            cur = cur + (px,)
            continue
        emit(table[cur], width)
        if nxt < 4096:
            table[cur + (px,)] = nxt
            nxt += 1
            if nxt - 1 == (1 << width) and width < 12:
                width += 1
        else:
            emit(clear, width)
            table = {(i,): i for i in range(clear)}
            width, nxt = min_bits + 1, eoi + 1
        cur = (px,)
    if cur:
        emit(table[cur], width)
    emit(eoi, width)
    if nbits:
        out.append(acc & 255)
    blocks = b"".join(bytes([len(out[i:i + 255])]) + bytes(out[i:i + 255]) for i in range(0, len(out), 255))
    return bytes([min_bits]) + blocks + b"\0"


def gif_file(W, H, palette, rect, indices, bg=0, transparent=None, interlace=False, local_palette=None, version=b"GIF89a", min_bits=None):
    """one-image GIF: global `palette` ([n,3], n a power of two), image `rect` (x, y, w, h) of palette `indices` (row-major)"""
    def table_bits(n):
        return int(math.log2(n)) - 1
    x, y, w, h = rect
    out = bytearray(version + struct.pack("<HHBBB", W, H, (0x80 | table_bits(len(palette))) if palette is not None else 0, bg, 0))
    if palette is not None:
        out += np.asarray(palette, np.uint8).tobytes()
    out += b"\x21\xfe\x05hello\x00"                                     # a comment extension
    if transparent is not None:
        out += b"\x21\xf9\x04" + struct.pack("<BHB", 1, 0, transparent) + b"\0"
    rows = np.asarray(indices, np.uint8).reshape(h, w)
    if interlace:
        order = list(range(0, h, 8)) + list(range(4, h, 8)) + list(range(2, h, 4)) + list(range(1, h, 2))
        rows = rows[order]
    out += b"\x2c" + struct.pack("<HHHHB", x, y, w, h, (0x40 if interlace else 0) | ((0x80 | table_bits(len(local_palette))) if local_palette is not None else 0))
    if local_palette is not None:
        out += np.asarray(local_palette, np.uint8).tobytes()
    ncol = len(local_palette) if local_palette is not None else len(palette)
    out += gif_lzw(rows.ravel().tolist(), min_bits or max(2, int(math.log2(ncol))))
    return bytes(out + b"\x3b")


def packbits(row):
    """PackBits of one row of bytes (runs of >= 3 equal bytes, literal stretches otherwise)"""
    out, i, n = bytearray(), 0, len(row)
    while i < n:
        j = i
        while j + 1 < n and row[j + 1] == row[i] and j - i < 127:
            j += 1
        if j - i >= 2:
            out += bytes([257 - (j - i + 1), row[i]])
            i = j + 1
            continue
        k = i
        while k < n and k - i < 128 and not (k + 2 < n and row[k] == row[k + 1] == row[k + 2]):
            k += 1
        out += bytes([k - i - 1]) + bytes(row[i:k])
        i = k
    return bytes(out)


def psd_file(planes, depth=8, rle=False):
    """flattened RGB PSD: planes [channels, h, w] of uint8 (depth 8) or uint16 (depth 16)"""
    c, h, w = planes.shape
    out = bytearray(b"8BPS" + struct.pack(">H6xHIIHH", 1, c, h, w, depth, 3))
    out += struct.pack(">I", 0) + struct.pack(">I", 6) + b"resrcs" + struct.pack(">I", 0)      # mode data, image resources, layers
    out += struct.pack(">H", 1 if rle else 0)
    if rle:
        rows = [packbits(planes[ch, y].astype(">u2" if depth == 16 else np.uint8).tobytes()) for ch in range(c) for y in range(h)]
        out += b"".join(struct.pack(">H", len(r)) for r in rows) + b"".join(rows)
    else:
        out += planes.astype(">u2" if depth == 16 else np.uint8).tobytes()
    return bytes(out)


def pic_file(w, h, packets, rows_coder):
    """Softimage PIC: packets = [(type, channel mask)], rows_coder(y, packet index) -> bytes of that packet's row"""
    out = bytearray(struct.pack(">If", 0x5380F634, 3.71) + b"rt3 loader fixture".ljust(80, b"\0") + b"PICT" + struct.pack(">HHfHH", w, h, 1.0, 3, 0))
    for i, (typ, chan) in enumerate(packets):
        out += bytes([1 if i + 1 < len(packets) else 0, 8, typ, chan])
    for y in range(h):
        for i in range(len(packets)):
            out += rows_coder(y, i)
    return bytes(out)


def rgbe(rgb):
    """float RGB [h, w, 3] -> Radiance RGBE bytes [h, w, 4]"""
    m = rgb.max(axis=2)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0)
    scale = np.where(m > 1e-32, 256.0 / np.exp2(e), 0)
    out = np.zeros(rgb.shape[:2] + (4,), np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    return out


def hdr_file(px, rle=True, signature=b"#?RADIANCE"):
    h, w = px.shape[:2]
    out = bytearray(signature + b"\n# made by hand\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %d +X %d\n" % (h, w))
    for y in range(h):
        if not rle:
            out += px[y].tobytes()
            continue
        out += bytes([2, 2, w >> 8, w & 255])
        for k in range(4):
            row, i = px[y, :, k].tolist(), 0
            while i < w:
                j = i
                while j + 1 < w and row[j + 1] == row[i] and j - i < 126:
                    j += 1
                if j - i >= 2:
                    out += bytes([128 + (j - i + 1), row[i]])
                    i = j + 1
                    continue
                k2 = i
                while k2 < w and k2 - i < 128 and not (k2 + 2 < w and row[k2] == row[k2 + 1] == row[k2 + 2]):
                    k2 += 1
                out += bytes([k2 - i]) + bytes(row[i:k2])
                i = k2
    return bytes(out)


def case_textures_more():
    d = case_dir("textures_more")
    names = []

    def put(fn, data):
        open(os.path.join(d, fn), "wb").write(data)
        names.append(fn)

    rng = np.random.RandomState(77)
    # ---- GIF
    pal16 = rng.randint(0, 256, (16, 3))
    w, h = 23, 17
    smooth = ((np.add.outer(np.arange(h), np.arange(w)) // 3) % 16).astype(np.uint8)
    noisy = rng.randint(0, 16, (h, w)).astype(np.uint8)
    put("plain.gif", gif_file(w, h, pal16, (0, 0, w, h), smooth))
    put("noisy87.gif", gif_file(w, h, pal16, (0, 0, w, h), noisy, version=b"GIF87a"))
    put("interlaced.gif", gif_file(w, h, pal16, (0, 0, w, h), noisy, interlace=True))
    put("transparent.gif", gif_file(w, h, pal16, (0, 0, w, h), smooth, transparent=5))
    put("subrect_bg.gif", gif_file(w, h, pal16, (4, 3, 11, 9), noisy[:9, :11], bg=7))            # uncovered pixels: background entry, red / blue exchanged
    put("subrect_bg0.gif", gif_file(w, h, pal16, (4, 3, 11, 9), noisy[:9, :11], bg=0))           # background index 0: uncovered pixels stay 0,0,0,0
    put("local_palette.gif", gif_file(w, h, pal16, (0, 0, w, h), noisy % 8, local_palette=rng.randint(0, 256, (8, 3)), transparent=2))
    put("only_local.gif", gif_file(w, h, None, (0, 0, w, h), noisy % 4, local_palette=rng.randint(0, 256, (4, 3))))
    pal256 = rng.randint(0, 256, (256, 3))
    big = rng.randint(0, 256, (90, 120)).astype(np.uint8)
    put("big256.gif", gif_file(120, 90, pal256, (0, 0, 120, 90), big))                           # fills the code table: clear codes in mid-stream
    put("beyond_table.gif", gif_file(w, h, pal16[:4], (0, 0, w, h), noisy % 8, min_bits=3))          # indices 4..7 have no palette entry: not drawn
    whole = gif_file(w, h, pal16, (0, 0, w, h), noisy, bg=3)
    put("cut.gif", whole[:len(whole) - 60])                                                      # data ends early: what was drawn stays, the rest is background
    put("no_image.gif", gif_file(w, h, pal16, (0, 0, w, h), smooth)[:13 + 48] + b"\x3b")
    try:
        from PIL import Image
        Image.fromarray(picture(31, 19, 3)).convert("P", palette=Image.ADAPTIVE, colors=64).save(os.path.join(d, "pil.gif"))
        names.append("pil.gif")
    except ImportError:
        pass
    # ---- PSD
    rgb = picture(19, 11, 6).transpose(2, 0, 1)
    alpha = picture(19, 11, 7)[..., 0][None]
    alpha[0, :2] = 255
    alpha[0, 2:4] = 0
    put("rgb8.psd", psd_file(rgb))
    put("rgba8.psd", psd_file(np.concatenate([rgb, alpha])))                                     # white matte removal
    put("rgba8_rle.psd", psd_file(np.concatenate([rgb // 32 * 32, alpha // 64 * 64]), rle=True))
    put("rgb16.psd", psd_file(rgb.astype(np.uint16) * 257 - 100, depth=16))
    put("rgba16.psd", psd_file(np.concatenate([rgb, alpha]).astype(np.uint16) * 256 + 17, depth=16))
    put("grey_as_rgb.psd", psd_file(rgb[:1]))                                                    # one channel in RGB mode: green and blue 0
    put("five.psd", psd_file(np.concatenate([rgb, alpha, alpha // 2]), rle=True))                # a fifth channel is ignored
    put("cmyk.psd", psd_file(rgb).replace(struct.pack(">HH", 8, 3), struct.pack(">HH", 8, 4), 1))  # refused
    # ---- PIC
    img = np.dstack([picture(21, 13, 8) // 16 * 16, picture(21, 13, 9)[..., :1]])
    put("raw_rgb.pic", pic_file(21, 13, [(0, 0xE0)], lambda y, i: img[y, :, :3].tobytes()))

    def mixed(vals):                                                                             # vals [n, k]: runs of equal pixels, literals otherwise
        out, i, n = bytearray(), 0, len(vals)
        while i < n:
            j = i
            while j + 1 < n and (vals[j + 1] == vals[i]).all() and j - i < 127:
                j += 1
            if j > i:
                out += bytes([127 + (j - i + 1)]) + vals[i].tobytes()
                i = j + 1
            else:
                k = i
                while k < n and k - i < 128 and not (k + 1 < n and (vals[k] == vals[k + 1]).all()):
                    k += 1
                out += bytes([k - i - 1]) + vals[i:k].tobytes()
                i = k
        return bytes(out)

    def pure(vals):
        out, i, n = bytearray(), 0, len(vals)
        while i < n:
            j = i
            while j + 1 < n and (vals[j + 1] == vals[i]).all() and j - i < 254:
                j += 1
            out += bytes([j - i + 1]) + vals[i].tobytes()
            i = j + 1
        return bytes(out)

    put("mixed_rgb_pure_a.pic", pic_file(21, 13, [(2, 0xE0), (1, 0x10)], lambda y, i: mixed(img[y, :, :3]) if i == 0 else pure(img[y, :, 3:] // 64 * 64)))
    put("long_run.pic", pic_file(300, 2, [(2, 0xE0)], lambda y, i: bytes([128]) + struct.pack(">H", 300) + bytes([10 * y + 1, 20, 30])))
    put("green_only.pic", pic_file(21, 13, [(0, 0x40)], lambda y, i: img[y, :, 1].tobytes()))     # red / blue / alpha stay 255
    put("overrun.pic", pic_file(21, 2, [(1, 0xE0)], lambda y, i: bytes([200, 1, 2, 3])))          # pure runs are cut at the row end
    # ---- Radiance HDR
    x = np.linspace(0, 1, 40)[None, :, None]
    yv = np.linspace(0, 1, 12)[:, None, None]
    lin = (np.array([0.02, 0.5, 3.0])[None, None, :] * (0.2 + x) * (0.3 + yv) * 2.0)
    lin[3:5, 10:30] = 0.0
    lin[6, :, :] = 1.0
    px = rgbe(lin)
    put("rle.hdr", hdr_file(px))
    put("flat.hdr", hdr_file(px, rle=False))
    put("narrow.hdr", hdr_file(px[:, :5]))                                                       # width < 8: always flat, whatever the bytes look like
    put("rgbe_sig.hdr", hdr_file(px, signature=b"#?RGBE"))
    put("xyze.hdr", hdr_file(px).replace(b"32-bit_rle_rgbe", b"32-bit_rle_xyze"))                 # refused
    tri_scene(d, names)
    return d, ["scene.obj"]


def run(case, paths, dumper, out):
    d = os.path.join(CASES, case)
    r = subprocess.run([dumper, out] + [os.path.join(d, p) for p in paths], capture_output=True, text=True)
    return r.returncode, r.stderr


def jobs():
    """(case directory name, golden file name, list of .obj paths) for every dump"""
    out = []
    for name in sorted(os.listdir(CASES)):
        d = os.path.join(CASES, name)
        if not os.path.isdir(d):
            continue
        if name == "keyframes":
            out.append((name, "golden.rt3l", ["k0.obj", "k1.obj", "k2.obj"]))
        else:
            for f in sorted(os.listdir(d)):
                if f.endswith(".obj"):
                    out.append((name, "golden.rt3l" if f == "scene.obj" else f[:-4] + ".golden.rt3l", [f]))
    return out


if __name__ == "__main__":
    for fn in (case_polygons, case_groups, case_syntax, case_keyframes, case_scenes, case_mtl_textures, case_textures_png, case_textures_other, case_textures_jpeg, case_textures_more):
        fn()
    if not os.path.exists(REF_DUMPER):
        sys.exit("fixtures written; oracle/_ref/dump_ref_loader is missing (make -C oracle ref_loader): goldens NOT refreshed")
    for case, golden, paths in jobs():
        rc, err = run(case, paths, REF_DUMPER, os.path.join(CASES, case, golden))
        print("%-16s %-24s reference loader rc=%d" % (case, golden, rc))
        assert rc == 0, err
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(CASES) for f in fs)
    print("fixtures + goldens: %.1f KB" % (total / 1024.0))
