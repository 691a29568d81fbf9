"""Parser of the canonical loadOBJ dump ("RT3L", oracle/ref_loader/loader_dump.hpp) + a readable diff (test infrastructure)."""
import struct

import numpy as np


def parse(path):
    d = open(path, "rb").read()
    assert d[:4] == b"RT3L", "not a loader dump"
    pos = 4

    def u(n=1):
        nonlocal pos
        v = struct.unpack_from("<%dI" % n, d, pos)
        pos += 4 * n
        return v

    nm, nt = u(2)
    meshes = []
    for _ in range(nm):
        keys, nv, ntri = u(3)
        per_key = []
        for _k in range(keys):
            arrs = []
            for c in (3, 3, 2):
                n, = u()
                arrs.append(np.frombuffer(d, np.float32, c * n, pos).reshape(n, c))
                pos += 4 * c * n
            per_key.append(arrs)
        idx = np.frombuffer(d, np.int32, 3 * ntri, pos).reshape(ntri, 3)
        pos += 12 * ntri
        mat_f = np.frombuffer(d, np.float32, 10, pos)
        pos += 40
        mat_i = np.frombuffer(d, np.int32, 4, pos)
        pos += 16
        meshes.append(dict(num_keys=keys, nv=nv, nt=ntri, keys=per_key, idx=idx, mat_f=mat_f, mat_i=mat_i))
    textures = []
    for _ in range(nt):
        w, h = u(2)
        textures.append(np.frombuffer(d, np.uint8, 4 * w * h, pos).reshape(h, w, 4))
        pos += 4 * w * h
    assert pos == len(d), "trailing bytes in loader dump"
    return meshes, textures


def diff(path_a, path_b):
    """'' when both dumps hold the same result, else a one-line description of the first difference."""
    (ma, ta), (mb, tb) = parse(path_a), parse(path_b)
    if len(ma) != len(mb):
        return "mesh count %d vs %d" % (len(ma), len(mb))
    for i, (a, b) in enumerate(zip(ma, mb)):
        for k in ("num_keys", "nv", "nt"):
            if a[k] != b[k]:
                return "mesh %d: %s %d vs %d" % (i, k, a[k], b[k])
        if not np.array_equal(a["idx"], b["idx"]):
            return "mesh %d: indices differ: %s vs %s" % (i, a["idx"].tolist()[:8], b["idx"].tolist()[:8])
        for k, (ka, kb) in enumerate(zip(a["keys"], b["keys"])):
            for name, x, y in zip(("vertices", "normals", "texcoords"), ka, kb):
                if x.shape != y.shape or not np.array_equal(x.view(np.uint32), y.view(np.uint32)):
                    return "mesh %d key %d: %s differ (%s vs %s)" % (i, k, name, x.shape, y.shape)
        if not np.array_equal(a["mat_f"].view(np.uint32), b["mat_f"].view(np.uint32)):
            return "mesh %d: material floats %s vs %s" % (i, a["mat_f"], b["mat_f"])
        if not np.array_equal(a["mat_i"], b["mat_i"]):
            return "mesh %d: texture ids %s vs %s" % (i, a["mat_i"], b["mat_i"])
    if len(ta) != len(tb):
        return "texture count %d vs %d" % (len(ta), len(tb))
    for i, (x, y) in enumerate(zip(ta, tb)):
        if x.shape != y.shape:
            return "texture %d: shape %s vs %s" % (i, x.shape, y.shape)
        if not np.array_equal(x, y):
            dd = np.abs(x.astype(int) - y.astype(int))
            return "texture %d: %.2f %% of bytes differ, max %d LSB" % (i, 100.0 * (dd > 0).mean(), dd.max())
    return ""


if __name__ == "__main__":
    import sys
    r = diff(sys.argv[1], sys.argv[2])
    print(r or "identical")
    if "-v" in sys.argv:
        for p in sys.argv[1:3]:
            ms, tx = parse(p)
            print(p, len(ms), "meshes", len(tx), "textures")
            for m in ms:
                print("  nv %d nt %d idx %s mat %s %s" % (m["nv"], m["nt"], m["idx"].tolist(), m["mat_f"], m["mat_i"]))
