"""N1 (texture ingest): rendertoy3c_b200/host/image_loader.hpp decodes PNG (own inflate: stored, fixed and dynamic
Huffman blocks; all five scanline filters; grey, grey+alpha, RGB, RGBA, palette + tRNS, 16-bit, sub-byte depths, Adam7 interlacing),
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


def write_png_adam7(path, px, ctype, filters, level=6):
    """8-bit Adam7-interlaced PNG of px [h, w, chan]: seven sub-images, each with its own filtered scanlines"""
    h, w, chan = px.shape
    out = bytearray()
    n = 0
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = px[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        prev = bytes(sub.shape[1] * chan)
        for y in range(sub.shape[0]):
            row = sub[y].tobytes()
            ft = filters[n % len(filters)]; n += 1
            enc = bytearray()
            for i, v in enumerate(row):
                a = row[i - chan] if i >= chan else 0
                b = prev[i]
                c = prev[i - chan] if i >= chan else 0
                enc.append((v - [0, a, b, (a + b) >> 1, paeth(a, b, c)][ft]) & 255)
            out += bytes([ft]) + enc
            prev = row
    f = b"\x89PNG\r\n\x1a\n" + png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 1))
    f += png_chunk(b"IDAT", zlib.compress(bytes(out), level)) + png_chunk(b"IEND", b"")
    open(path, "wb").write(f)


def test_png_adam7_interlaced(exe, tmp_path):
    rng = np.random.RandomState(9)
    for (h, w) in ((23, 37), (1, 1), (3, 2), (8, 8), (9, 17)):          # small sizes leave some passes empty
        img = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
        for tag, ctype, px in (("rgba", 6, img), ("rgb", 2, img[..., :3]), ("grey", 0, img[..., :1])):
            p = str(tmp_path / ("adam7_%s_%dx%d.png" % (tag, w, h)))
            write_png_adam7(p, np.ascontiguousarray(px), ctype, [0, 1, 2, 3, 4])
            got, err = decode(exe, p, tmp_path)
            assert got is not None, (tag, w, h, err)
            want = np.full((h, w, 4), 255, np.uint8)
            want[..., :px.shape[2] if ctype != 0 else 3] = px if ctype != 0 else np.repeat(px, 3, axis=2)
            assert np.array_equal(got, want), (tag, w, h)


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


ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def _dct_matrix():
    m = np.zeros((8, 8))
    for u in range(8):
        for x in range(8):
            m[u, x] = (np.sqrt(0.5) if u == 0 else 1.0) * np.cos((2 * x + 1) * u * np.pi / 16) / 2
    return m


def write_baseline_jpeg(path, rgb, sampling=(1, 1), restart=0, grey=False, q=3, optimal=True):
    """A small baseline JPEG encoder for the tests (Huffman tables built from the symbols that occur: optimal lengths
    with the reserved all-ones code, or equal-length codes).  Returns the image a straightforward decoder must produce (float IDCT, replicated chroma, JFIF colour)."""
    h, w = rgb.shape[:2]
    f = rgb.astype(np.float64)
    if grey:
        planes, samp = [np.rint(0.299 * f[..., 0] + 0.587 * f[..., 1] + 0.114 * f[..., 2])], [(1, 1)]
    else:
        y = 0.299 * f[..., 0] + 0.587 * f[..., 1] + 0.114 * f[..., 2]
        cb = 128 - 0.168736 * f[..., 0] - 0.331264 * f[..., 1] + 0.5 * f[..., 2]
        cr = 128 + 0.5 * f[..., 0] - 0.418688 * f[..., 1] - 0.081312 * f[..., 2]
        planes, samp = [np.rint(y), np.rint(cb), np.rint(cr)], [sampling, (1, 1), (1, 1)]
    hmax, vmax = max(s[0] for s in samp), max(s[1] for s in samp)
    mcux, mcuy = -(-w // (8 * hmax)), -(-h // (8 * vmax))
    qt = [np.clip(np.add.outer(np.arange(8), np.arange(8)) * q + 2, 1, 255).astype(np.int64), np.full((8, 8), 4 * q + 3, np.int64)]
    D = _dct_matrix()
    coefs, recon = [], []
    for ci, (pl, (sh, sv)) in enumerate(zip(planes, samp)):
        fx, fy = hmax // sh, vmax // sv
        ph, pw = mcuy * sv * 8, mcux * sh * 8
        full = np.pad(pl, ((0, mcuy * vmax * 8 - h), (0, mcux * hmax * 8 - w)), mode="edge")
        sub = full.reshape(ph, fy, pw, fx).mean(axis=(1, 3))        # box-filtered chroma
        qq = qt[0 if ci == 0 else 1]
        blocks = sub.reshape(ph // 8, 8, pw // 8, 8).transpose(0, 2, 1, 3) - 128.0
        co = np.rint(np.einsum("ux,abxy,vy->abuv", D, blocks, D) / qq).astype(np.int64)   # co[by, bx, v(row freq), u]
        coefs.append(co)
        rec = np.einsum("ux,abuv,vy->abxy", D, (co * qq).astype(np.float64), D) + 128.0
        recon.append(np.clip(np.rint(rec), 0, 255).transpose(0, 2, 1, 3).reshape(ph, pw))
    # symbol stream
    def cat(v):
        return 0 if v == 0 else int(abs(int(v))).bit_length()

    def bits_of(v, s):
        return (v if v >= 0 else v + (1 << s) - 1, s)
    stream = []                                                     # items: ("dc"|"ac", table, symbol) or ("raw", value, nbits) or ("rst", n)
    pred = [0] * len(planes)
    n_mcu, rst = 0, 0
    for my in range(mcuy):
        for mx in range(mcux):
            if restart and n_mcu and n_mcu % restart == 0:
                stream.append(("rst", rst)); rst = (rst + 1) & 7; pred = [0] * len(planes)
            for ci, (sh, sv) in enumerate(samp):
                tb = 0 if ci == 0 else 1
                for by in range(sv):
                    for bx in range(sh):
                        zzv = coefs[ci][my * sv + by, mx * sh + bx].reshape(64)[ZIGZAG]
                        d = int(zzv[0]) - pred[ci]; pred[ci] = int(zzv[0])
                        s = cat(d)
                        stream.append(("dc", tb, s))
                        if s:
                            stream.append(("raw",) + bits_of(d, s))
                        run = 0
                        last = max([k for k in range(1, 64) if zzv[k] != 0], default=0)
                        for k in range(1, last + 1):
                            v = int(zzv[k])
                            if v == 0:
                                run += 1
                                continue
                            while run > 15:
                                stream.append(("ac", tb, 0xF0)); run -= 16
                            s = cat(v)
                            stream.append(("ac", tb, (run << 4) | s))
                            stream.append(("raw",) + bits_of(v, s))
                            run = 0
                        if last < 63:
                            stream.append(("ac", tb, 0x00))
            n_mcu += 1
    tables = {}
    for kind in ("dc", "ac"):
        for tb in range(2 if not grey else 1):
            occ = [it[2] for it in stream if it[0] == kind and it[1] == tb] or [0]
            syms = sorted(set(occ))
            if not optimal:
                L = max(1, int(len(syms)).bit_length())             # 2**L > len(syms): the all-ones code stays unused
                tables[(kind, tb)] = ({sy: (i, L) for i, sy in enumerate(syms)}, [0] * (L - 1) + [len(syms)] + [0] * (16 - L), syms)
                continue
            # Huffman code lengths with a reserved (never used) symbol that takes the all-ones code, T.81 K.2
            import heapq
            heap = [(occ.count(sy), i, [sy]) for i, sy in enumerate(syms)] + [(0.5, len(syms), [None])]
            length = {sy: 0 for sy in syms}; length[None] = 0
            heapq.heapify(heap)
            tick = len(heap)
            while len(heap) > 1:
                a, b = heapq.heappop(heap), heapq.heappop(heap)
                for sy in a[2] + b[2]:
                    length[sy] += 1
                heapq.heappush(heap, (a[0] + b[0], tick, a[2] + b[2])); tick += 1
            assert max(length.values()) <= 16
            order = sorted(syms, key=lambda sy: (length[sy], sy)) + [None]   # the reserved symbol last among the longest
            assert length[None] == max(length.values())
            codes, code, prev = {}, 0, 0
            for sy in order:
                code <<= length[sy] - prev; prev = length[sy]
                codes[sy] = (code, length[sy]); code += 1
            assert codes[None][0] == (1 << length[None]) - 1
            counts = [sum(1 for sy in syms if length[sy] == l) for l in range(1, 17)]
            tables[(kind, tb)] = ({sy: codes[sy] for sy in syms}, counts, order[:-1])
    out, acc, nacc = bytearray(), 0, 0

    def put(v, n):
        nonlocal acc, nacc
        acc = (acc << n) | v; nacc += n
        while nacc >= 8:
            b = (acc >> (nacc - 8)) & 255
            out.append(b)
            if b == 0xFF:
                out.append(0)
            nacc -= 8
        acc &= (1 << nacc) - 1
    for it in stream:
        if it[0] == "raw":
            put(it[1], it[2])
        elif it[0] == "rst":
            if nacc:
                put((1 << (8 - nacc)) - 1, 8 - nacc)
            out += bytes([0xFF, 0xD0 + it[1]])
        else:
            put(*tables[(it[0], it[1])][0][it[2]])
    if nacc:
        put((1 << (8 - nacc)) - 1, 8 - nacc)
    seg = lambda m, body: bytes([0xFF, m]) + struct.pack(">H", len(body) + 2) + body
    jf = b"\xff\xd8" + seg(0xE0, b"JFIF\0\1\1\0\0\1\0\1\0\0")
    for i in range(1 if grey else 2):
        jf += seg(0xDB, bytes([i]) + bytes(int(qt[i].reshape(64)[ZIGZAG[k]]) for k in range(64)))
    jf += seg(0xC0, struct.pack(">BHHB", 8, h, w, len(planes)) + b"".join(bytes([ci + 1, (sh << 4) | sv, 0 if ci == 0 else 1]) for ci, (sh, sv) in enumerate(samp)))
    for (kind, tb), (_, counts, syms) in tables.items():
        jf += seg(0xC4, bytes([(16 if kind == "ac" else 0) | tb]) + bytes(counts) + bytes(syms))
    if restart:
        jf += seg(0xDD, struct.pack(">H", restart))
    jf += seg(0xDA, bytes([len(planes)]) + b"".join(bytes([ci + 1, 0x00 if ci == 0 else 0x11]) for ci in range(len(planes))) + bytes([0, 63, 0]))
    open(path, "wb").write(jf + bytes(out) + b"\xff\xd9")
    # what a decoder must deliver
    up = []
    for ci, (sh, sv) in enumerate(samp):
        up.append(np.repeat(np.repeat(recon[ci], vmax // sv, axis=0), hmax // sh, axis=1)[:h, :w])
    if grey:
        return np.stack([up[0]] * 3, axis=-1)
    yy, cb, cr = up[0], up[1] - 128.0, up[2] - 128.0
    return np.clip(np.rint(np.stack([yy + 1.402 * cr, yy - 0.344136 * cb - 0.714136 * cr, yy + 1.772 * cb], axis=-1)), 0, 255)


def test_baseline_jpeg(exe, tmp_path):
    rng = np.random.RandomState(8)
    h, w = 37, 53                                                   # not multiples of the MCU size
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 7.0), 128 + 100 * np.cos(yy / 5.0), (xx * 3 + yy * 5) % 256], axis=-1) + rng.randint(-12, 13, size=(h, w, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    for tag, kw in (("444", {}), ("420", {"sampling": (2, 2)}), ("422_rst", {"sampling": (2, 1), "restart": 3}), ("grey_rst", {"grey": True, "restart": 2}), ("fine", {"q": 1}), ("flatcodes", {"optimal": False, "sampling": (1, 2)})):
        p = str(tmp_path / (tag + ".jpg"))
        want = write_baseline_jpeg(p, img, **kw)
        got, err = decode(exe, p, tmp_path)
        assert got is not None, (tag, err)
        # The decoder's exact bytes are pinned against the reference's stb_image in tests/test_loader_parity.py (integer
        # IDCT, filtered chroma upsampling, fixed-point YCbCr).  `want` is a float model with replicated chroma: it only
        # bounds the decode from the outside here - close where nothing is subsampled, the same picture elsewhere.
        diff = np.abs(got[..., :3].astype(np.int32) - want.astype(np.int32))
        if "sampling" not in kw:
            assert diff.max() <= 4 and diff.mean() < 0.6, (tag, diff.max(), diff.mean())
        else:
            assert diff.mean() < 6, (tag, diff.mean())
        assert (got[..., 3] == 255).all()
        orig = img.astype(np.float64)
        if kw.get("grey"):
            orig = np.repeat((0.299 * orig[..., 0] + 0.587 * orig[..., 1] + 0.114 * orig[..., 2])[..., None], 3, axis=-1)
        assert np.abs(got[..., :3].astype(np.float64) - orig).mean() < (12 if tag != "fine" else 6)   # and it is the picture, not noise


def test_rejects_what_it_cannot_decode(exe, tmp_path):
    p = str(tmp_path / "x.jpg"); open(p, "wb").write(b"\xff\xd8\xff\xc9" + struct.pack(">H", 17) + bytes(15) + b"\xff\xd9")   # SOF9: arithmetic coding
    got, err = decode(exe, p, tmp_path)
    assert got is None and "arithmetic" in err
    p = str(tmp_path / "y.jpg"); open(p, "wb").write(b"\xff\xd8\xff\xe0" + bytes(64))
    got, err = decode(exe, p, tmp_path)
    assert got is None
    rows = [bytes(12)] * 4
    p = str(tmp_path / "i.png"); write_png(p, rows, 2, 8, [0], interlace=2)
    got, err = decode(exe, p, tmp_path)
    assert got is None and "interlace" in err
    # malformed headers must be refused before anything is indexed or allocated from them (bit depths 0 / 3 / 5 / 6 / 7,
    # 16-bit palette, sub-byte RGB, absurd sizes)
    for ctype, depth in ((0, 5), (0, 0), (3, 3), (0, 6), (3, 7), (3, 16), (2, 4), (6, 2), (4, 1)):
        p = str(tmp_path / ("bad_%d_%d.png" % (ctype, depth)))
        write_png(p, rows, ctype, 8, [0])
        data = bytearray(open(p, "rb").read())
        data[24] = depth; data[25] = ctype                             # IHDR depth / colour type (the CRC is not checked by the loader)
        open(p, "wb").write(bytes(data))
        got, err = decode(exe, p, tmp_path)
        assert got is None and "bit depth" in err, (ctype, depth, err)
    p = str(tmp_path / "huge.png"); write_png(p, rows, 2, 8, [0])
    data = bytearray(open(p, "rb").read())
    data[16:24] = struct.pack(">II", 60000, 60000)
    open(p, "wb").write(bytes(data))
    got, err = decode(exe, p, tmp_path)
    assert got is None and "large" in err
    p = str(tmp_path / "t.png"); write_png(p, rows, 2, 8, [0])
    data = open(p, "rb").read()
    open(p, "wb").write(data[:len(data) // 2])
    got, err = decode(exe, p, tmp_path)
    assert got is None
