import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


EMUL_LIB = os.path.join(ROOT, "tests", "_emul", "librt3_emul.so")


@pytest.fixture(scope="session")
def emul_lib():
    """Kernel-logic simulator: the SAME kernel bodies (rendertoy3c_b200/csrc/*.cuh) compiled by g++ with
    -DRT3_EMULATE, each launch a host loop over thread ids.  Test infrastructure for this GPU-less box;
    never part of librt3.so."""
    csrc = os.path.join(ROOT, "rendertoy3c_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", "rt3.h")]
    if not os.path.exists(EMUL_LIB) or any(os.path.getmtime(s) > os.path.getmtime(EMUL_LIB) for s in srcs):
        os.makedirs(os.path.dirname(EMUL_LIB), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared", "-DRT3_EMULATE",
                        "-x", "c++", os.path.join(csrc, "rt3_lib.cu"), "-I/usr/local/cuda/include", "-o", EMUL_LIB], check=True)
    return EMUL_LIB


@pytest.fixture(scope="session")
def product_lib():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    return g.LIB
