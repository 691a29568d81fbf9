"""N1 (texture ingest): rendertoy3c_b200/host/image_loader.hpp decodes PNG (own inflate: stored, fixed and dynamic
Huffman blocks; all five scanline filters; grey, grey+alpha, RGB, RGBA, palette + tRNS, 16-bit, sub-byte depths),
BMP, TGA and PNM files to the RGBA8, bottom-row-first layout the reference's loadOBJ produces (src/mesh.cpp:137-159).
The files are written here with the standard library only (struct + zlib)."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from test_host_cpp import build_host


@pytest.fixture(scope="module")
def exe(tmp_path_factory, emul_lib):
    d = tmp_path_factory.mktemp("imgexe")
    return build_host(os.path.dirname(emul_lib), os.path.basename(emul_lib), str(d / "wavefront_emul"))


def decode(exe, path, tmp_path):
    out = str(tmp_path / "decoded.bin")
    r = subprocess.run([exe, "--decode-image", path, out], capture_output=True, text=True)
    if r.returncode != 0:
        return None, r.stderr
    raw = open(out, "rb").read()
    w, h = struct.unpack("<ii", raw[:8])
    return np.frombuffer(raw[8:], dtype=np.uint8).reshape(h, w, 4)[::-1], ""   # back to top row first


def png_chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def write_png(path, rows, ctype, depth, filters, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, plte=None, trns=None, split=1, interlace=0):
    """rows: list of bytes objects (packed scanlines); filters: per-row filter type"""
    bpp = max(1, {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype] * depth // 8)
    out, prev = bytearray(), bytes(len(rows[0]))
    for y, row in enumerate(rows):
        ft = filters[y % len(filters)]
        enc = bytearray()
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = [0, a, b, (a + b) >> 1, paeth(a, b, c)][ft]
            enc.append((v - pred) & 255)
        out += bytes([ft]) + enc
        prev = row
    co = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
    z = co.compress(bytes(out)) + co.flush()
    w = len(rows[0]) * 8 // ({0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype] * depth)
    f = b"\x89PNG\r\n\x1a\n" + png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, len(rows), depth, ctype, 0, 0, interlace))
    if plte is not None:
        f += png_chunk(b"PLTE", bytes(plte))
    if trns is not None:
        f += png_chunk(b"tRNS", bytes(trns))
    step = (len(z) + split - 1) // split
    for k in range(0, len(z), step):
        f += png_chunk(b"IDAT", z[k:k + step])
    f += png_chunk(b"IEND", b"")
    open(path, "wb").write(f)


def test_png_variants(exe, tmp_path):
    rng = np.random.RandomState(5)
    w, h = 37, 23
    smooth = (np.add.outer(np.arange(h) * 5, np.arange(w) * 3)[..., None] + np.array([0, 40, 90, 200])) % 256   # compressible
    noise = rng.randint(0, 256, size=(h, w, 4))
    for name, img in (("smooth", smooth), ("noise", noise)):
        img = img.astype(np.uint8)
        cases = [
            ("rgba", 6, img, [0, 1, 2, 3, 4], 6, zlib.Z_DEFAULT_STRATEGY, 1),
            ("rgb_fixed", 2, img[..., :3], [4, 3], 6, zlib.Z_FIXED, 3),
            ("rgb_stored", 2, img[..., :3], [1], 0, zlib.Z_DEFAULT_STRATEGY, 1),
            ("grey", 0, img[..., :1], [2, 0, 4], 9, zlib.Z_DEFAULT_STRATEGY, 2),
            ("greya", 4, img[..., [0, 3]], [3], 9, zlib.Z_DEFAULT_STRATEGY, 1),
        ]
        for tag, ctype, px, filt, level, strat, split in cases:
            p = str(tmp_path / ("%s_%s.png" % (name, tag)))
            write_png(p, [px[y].tobytes() for y in range(h)], ctype, 8, filt, level, strat, split=split)
            got, err = decode(exe, p, tmp_path)
            assert got is not None, err
            want = np.empty((h, w, 4), np.uint8)
            if ctype == 6:
                want[:] = px
            elif ctype == 2:
                want[..., :3] = px; want[..., 3] = 255
            elif ctype == 0:
                want[..., :3] = px; want[..., 3] = 255
            else:
                want[..., :3] = px[..., :1]; want[..., 3] = px[..., 1]
            assert np.array_equal(got, want), (name, tag)


def test_png_palette_16bit_and_subbyte(exe, tmp_path):
    rng = np.random.RandomState(6)
    w, h = 19, 11
    pal = rng.randint(0, 256, size=(16, 3)).astype(np.uint8)
    alpha = rng.randint(0, 256, size=7).astype(np.uint8)            # tRNS shorter than the palette: the rest is opaque
    idx = rng.randint(0, 16, size=(h, w)).astype(np.uint8)
    p = str(tmp_path / "pal8.png")
    write_png(p, [idx[y].tobytes() for y in range(h)], 3, 8, [0, 4], plte=pal.tobytes(), trns=alpha.tobytes())
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    want = np.concatenate([pal[idx], np.where(idx < 7, alpha[np.minimum(idx, 6)], 255)[..., None]], axis=-1).astype(np.uint8)
    assert np.array_equal(got, want)
    # 4-bit palette: two indices per byte, high nibble first, rows padded to a byte
    rows = []
    for y in range(h):
        r = list(idx[y]) + [0] * (w % 2)
        rows.append(bytes((r[i] << 4) | r[i + 1] for i in range(0, len(r), 2)))
    p = str(tmp_path / "pal4.png")
    bpp_w = rows[0]
    # write_png derives the width from the row length; patch IHDR by writing with the true width through a wrapper
    write_png(p, rows, 3, 4, [0], plte=pal.tobytes())
    data = bytearray(open(p, "rb").read())
    data[16:20] = struct.pack(">I", w)                              # true width (odd) in IHDR ...
    data[29:33] = struct.pack(">I", zlib.crc32(bytes(data[12:29])) & 0xFFFFFFFF)   # ... and its CRC (not checked by the loader, kept valid anyway)
    open(p, "wb").write(data)
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got[..., :3], pal[idx]) and (got[..., 3] == 255).all()
    # 16-bit RGB: the high byte is kept
    px16 = rng.randint(0, 65536, size=(h, w, 3)).astype(">u2")
    p = str(tmp_path / "rgb16.png")
    write_png(p, [px16[y].tobytes() for y in range(h)], 2, 16, [1, 2])
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got[..., :3], (px16 >> 8).astype(np.uint8))
    # 1-bit grey scales to 0 / 255
    bits = rng.randint(0, 2, size=(h, 24)).astype(np.uint8)
    p = str(tmp_path / "grey1.png")
    write_png(p, [np.packbits(bits[y]).tobytes() for y in range(h)], 0, 1, [0])
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got[..., 0], bits * 255)


def test_bmp_tga_pnm(exe, tmp_path):
    rng = np.random.RandomState(7)
    w, h = 13, 9
    img = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    # BMP 24 bit, bottom-up, rows padded to 4 bytes
    stride = (w * 3 + 3) & ~3
    body = b"".join(img[y, :, [2, 1, 0]].T.tobytes() + bytes(stride - w * 3) for y in range(h - 1, -1, -1))
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(body), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(body), 2835, 2835, 0, 0)
    p = str(tmp_path / "a.bmp"); open(p, "wb").write(hdr + body)
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got[..., :3], img[..., :3]) and (got[..., 3] == 255).all()
    # BMP 32 bit, top-down (negative height)
    body = b"".join(img[y][:, [2, 1, 0, 3]].tobytes() for y in range(h))
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(body), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, -h, 1, 32, 0, len(body), 2835, 2835, 0, 0)
    p = str(tmp_path / "b.bmp"); open(p, "wb").write(hdr + body)
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got, img)
    # TGA 24 bit raw, bottom-left origin
    body = b"".join(img[y][:, [2, 1, 0]].tobytes() for y in range(h - 1, -1, -1))
    p = str(tmp_path / "a.tga"); open(p, "wb").write(struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, w, h, 24, 0) + body)
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got[..., :3], img[..., :3])
    # TGA 32 bit RLE, top-left origin, with an image-id field: alternate run and raw packets
    flat = img.copy()
    flat[:, 4:9] = flat[:, 4:5]                                     # a run of 5 equal pixels per row
    px = flat[:, :, [2, 1, 0, 3]].reshape(-1, 4)
    body, i = bytearray(), 0
    while i < len(px):
        run = 1
        while i + run < len(px) and run < 128 and (px[i + run] == px[i]).all():
            run += 1
        if run > 1:
            body += bytes([0x80 | (run - 1)]) + px[i].tobytes(); i += run
        else:
            n = min(3, len(px) - i)
            body += bytes([n - 1]) + px[i:i + n].tobytes(); i += n
    p = str(tmp_path / "b.tga"); open(p, "wb").write(struct.pack("<BBBHHBHHHHBB", 5, 0, 10, 0, 0, 0, 0, 0, w, h, 32, 0x28) + b"hello" + bytes(body))
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got, flat)
    # PGM (P5) with a comment, PPM (P6)
    p = str(tmp_path / "a.pgm"); open(p, "wb").write(b"P5\n# made here\n%d %d\n255\n" % (w, h) + img[..., 0].tobytes())
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got[..., 1], img[..., 0])
    p = str(tmp_path / "a.ppm"); open(p, "wb").write(b"P6 %d %d 255\n" % (w, h) + img[..., :3].tobytes())
    got, err = decode(exe, p, tmp_path)
    assert got is not None, err
    assert np.array_equal(got[..., :3], img[..., :3])


def test_rejects_what_it_cannot_decode(exe, tmp_path):
    p = str(tmp_path / "x.jpg"); open(p, "wb").write(b"\xff\xd8\xff\xe0" + bytes(64))
    got, err = decode(exe, p, tmp_path)
    assert got is None and "JPEG" in err
    rows = [bytes(12)] * 4
    p = str(tmp_path / "i.png"); write_png(p, rows, 2, 8, [0], interlace=1)
    got, err = decode(exe, p, tmp_path)
    assert got is None and "interlaced" in err
    p = str(tmp_path / "t.png"); write_png(p, rows, 2, 8, [0])
    data = open(p, "rb").read()
    open(p, "wb").write(data[:len(data) // 2])
    got, err = decode(exe, p, tmp_path)
    assert got is None
