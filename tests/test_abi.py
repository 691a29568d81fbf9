"""The C-ABI library loads and exports every symbol include/rt3.h declares; without a GPU the product
fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

from rendertoy3c_b200 import _abi
from rendertoy3c_b200.api import RT3_SYMBOLS, Context, Rt3Error, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rt3.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt3_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(RT3_SYMBOLS)


def test_library_exports_every_declared_symbol(product_lib):
    L = load_library(product_lib)
    for s in declared_symbols():
        assert getattr(L, s) is not None


def test_struct_sizes_match_header():
    assert C.sizeof(_abi.RenderSettings) == 4 * 4 + 12 * 4 + 2 * 4 + 3 * 4 + 4
    assert _abi.RAY_DTYPE.itemsize == 48 and _abi.HIT_DTYPE.itemsize == 32
    assert C.sizeof(_abi.Stats) == 5 * 8 + 6 * 4 + 4 * 4


def test_product_has_no_oracle_or_emulation_inside(product_lib):
    """librt3.so must not contain the oracle's or the simulator's entry points."""
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", product_lib], capture_output=True, text=True).stdout
    assert "rt3o_" not in syms
    pkg = os.path.join(ROOT, "rendertoy3c_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            txt = open(os.path.join(pkg, f)).read()
            assert "oracle_backend" not in txt and "librt3o" not in txt and "librt3_emul" not in txt


def test_no_gpu_means_loud_failure(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Rt3Error) as e:
        Context(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
