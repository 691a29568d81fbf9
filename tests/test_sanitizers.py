"""AddressSanitizer + UndefinedBehaviorSanitizer over (a) the kernel bodies, through the kernel-logic simulator, and (b) the
scene loader on damaged input.  The GPU pool offers no compute-sanitizer; the kernels' memory safety otherwise rests on
bit-exact parity and the argument checks at the ABI."""
import os
import random
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAN = ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-O1", "-g"]


def _runtime(name):
    p = subprocess.run(["g++", "-print-file-name=" + name], capture_output=True, text=True).stdout.strip()
    return p if os.path.isabs(p) and os.path.exists(p) else None


def _clean(stderr):
    return "runtime error" not in stderr and "AddressSanitizer" not in stderr


def test_kernel_bodies_under_asan_ubsan(tmp_path):
    asan, ubsan = _runtime("libasan.so"), _runtime("libubsan.so")
    if not asan or not ubsan:
        pytest.skip("sanitizer runtimes not installed")
    lib = str(tmp_path / "librt3_emul_asan.so")
    subprocess.run(["g++", "-std=c++17", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared", "-DRT3_EMULATE"] + SAN +
                   ["-x", "c++", os.path.join(ROOT, "rendertoy3c_b200", "csrc", "rt3_lib.cu"), "-I/usr/local/cuda/include", "-o", lib], check=True)
    env = dict(os.environ, LD_PRELOAD=asan + ":" + ubsan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "sanitize_emul.py"), lib], capture_output=True, text=True, env=env, timeout=1200)
    assert r.returncode == 0 and "SANITIZED RUN COMPLETE" in r.stdout and _clean(r.stderr), r.stderr[-3000:]


def test_loader_on_damaged_files_under_asan_ubsan(tmp_path):
    """400 mutated texture files (every format) and 300 mutated .obj / .mtl files: refused or loaded, never a bad access or
    undefined arithmetic"""
    if not _runtime("libasan.so"):
        pytest.skip("sanitizer runtimes not installed")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "loader"))
    import make_loader_goldens as gen
    exe = str(tmp_path / "dump_own_loader_asan")
    subprocess.run(["g++", "-std=c++17"] + SAN + ["-o", exe, os.path.join(ROOT, "tests", "tools", "dump_own_loader.cpp")], check=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    r = random.Random(99)
    pool = []
    for case in ("textures_more", "textures_other", "textures_png", "textures_jpeg"):
        d = os.path.join(gen.CASES, case)
        pool += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.rsplit(".", 1)[-1] in ("gif", "psd", "pic", "hdr", "bmp", "tga", "ppm", "pgm", "png", "jpg")]

    def mutate(b, text):
        for _ in range(r.randrange(1, 6)):
            if not b:
                b += b"v 0 0 0\n"
            mode, i = r.randrange(4), r.randrange(len(b))
            if mode == 0:
                b[i] = r.choice(b"0123456789-+./ eE\n\t#fvntmlgous") if text else r.randrange(256)
            elif mode == 1:
                if text:
                    del b[i:i + r.randrange(1, 20)]
                elif len(b) > 1:
                    del b[r.randrange(1, len(b)):]
            elif mode == 2:
                if text:
                    b[i:i] = r.choice([b"-1", b"99999999", b"/", b"//", b" ", b"\n", b"1e400", b"nan", b"f ", b"usemtl x\n", b"-0"])
                else:
                    b[i] ^= 1 << r.randrange(8)
            else:
                b[i:i + 4] = bytes([255, 255, 255, 127])[:max(0, min(4, len(b) - i))]
        return b

    work = tmp_path / "w"
    for it in range(400):
        shutil.rmtree(work, ignore_errors=True)
        work.mkdir()
        src = r.choice(pool)
        data = bytearray(open(src, "rb").read())
        if not data:
            continue
        open(work / os.path.basename(src), "wb").write(mutate(data, False))
        gen.tri_scene(str(work), [os.path.basename(src)])
        o = subprocess.run([exe, str(tmp_path / "o.bin"), str(work / "scene.obj")], capture_output=True, env=env, timeout=120)
        assert o.returncode >= 0 and _clean(o.stderr.decode("latin1")), (src, o.stderr[-2000:])
    for it in range(300):
        case = r.choice(["polygons", "groups", "syntax", "mtl_textures", "scenes"])
        shutil.rmtree(work, ignore_errors=True)
        shutil.copytree(os.path.join(gen.CASES, case), work)
        texts = sorted(f for f in os.listdir(work) if f.endswith(".obj") or f.endswith(".mtl"))
        objs = [f for f in texts if f.endswith(".obj")]
        f = r.choice(texts)
        open(work / f, "wb").write(mutate(bytearray(open(work / f, "rb").read()), True))
        o = subprocess.run([exe, str(tmp_path / "o.bin"), str(work / r.choice(objs))], capture_output=True, env=env, timeout=120)
        assert o.returncode >= 0 and _clean(o.stderr.decode("latin1")), (case, f, o.stderr[-2000:])
