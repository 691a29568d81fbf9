"""Pins of the hot-path bodies to REFERENCE CODE.

tests/golden/ref_images.npz holds subframes rendered by the reference's own device programs — src/shader/raygen.cu,
closehit_radiance.cu, miss.cu, test.cu with src/light.h, cuda/random.h, cuda/helpers.h, src/util/sampling.h — compiled
where they lie and run on the host through the functional OptiX stand-in oracle/ref_shim (generator
tests/golden/make_ref_images.py).  Only traversal and the texel fetch are not reference code there (the reference has no
source for them): they come from the oracle's brute-force intersector.

What is asserted:
  1. with its two documented deviations switched off — D1 (explicit sincos polynomial instead of libm cosf / sinf) and D3
     (per-sample instead of running radiance sum) — the oracle reproduces the reference-held float accumulation buffers of
     every subframe BIT FOR BIT, and the 8-bit frames;
  2. D3 alone moves a pixel by at most 1e-6 relative (measured 2.7e-7);
  3. D1 is not a per-pixel bound on every scene: the reference offsets its shadow rays by a fixed tmin = 0.001 from a hit
     point that is itself rounded (closehit_radiance.cu:80,132-138; its own "TODO: smarter offset", raygen.cu:54), so a
     1-ulp change of a bounce direction flips self-shadowing at grazing light angles.  On the open terrain the bound is
     per pixel (max 1e-3 relative, 99 % of the pixels within 1e-4); in the closed Cornell box it is statistical (relative
     MSE <= 2e-3 at 24 spp, image mean within 5e-4), which is what "image within a stated relMSE of the reference" means;
  4. on a GPU, the kernels meet the same bounds against the same reference-held images;
  5. where /root/reference exists, the goldens are re-rendered from it and must equal the committed file."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_ref_images import pin_scenes, render  # noqa: E402
from oracle_backend import OracleScene, lib as oracle_lib  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_images.npz"))
NAMES = sorted(pin_scenes())


def oracle_render(name, chain_sum, libm_sincos):
    L = oracle_lib()
    L.rt3o_set_chain_sum(chain_sum)
    L.rt3o_set_libm_sincos(libm_sincos)
    try:
        desc, n = pin_scenes()[name]
        return render(OracleScene(), desc, n)
    finally:
        L.rt3o_set_chain_sum(0)
        L.rt3o_set_libm_sincos(0)


def deviation(accum, name):
    a = accum[-1][..., :3].astype(np.float64)
    b = GOLD[name + "_accum"][-1][..., :3].astype(np.float64)
    rel = (np.abs(a - b) / np.maximum(np.abs(b), 1e-3)).max(axis=2)
    return dict(max_rel=float(rel.max()), frac_1e4=float((rel > 1e-4).mean()), rel_mse=float(np.mean((a - b) ** 2 / (b * b + 1e-4))),
                mean_ratio=float(a.mean() / b.mean()))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_is_bit_identical_to_the_reference_programs(name):
    accum, frame = oracle_render(name, chain_sum=1, libm_sincos=1)
    want = GOLD[name + "_accum"]
    assert accum.shape == want.shape
    same = (accum.view(np.uint32) == want.view(np.uint32)).all(axis=-1)
    assert same.all(), "%s: %d of %d pixel-subframes differ from the reference's programs (libm cosf / sinf of this machine must be the one the goldens were made with)" % (name, (~same).sum(), same.size)
    assert np.array_equal(frame, GOLD[name + "_frame"])


@pytest.mark.parametrize("name", NAMES)
def test_d3_sum_order_alone(name):
    accum, frame = oracle_render(name, chain_sum=0, libm_sincos=1)
    d = deviation(accum, name)
    assert d["max_rel"] <= 1e-6, d
    assert np.array_equal(frame, GOLD[name + "_frame"])


def check_default_mode(accum, frame, name):
    d = deviation(accum, name)
    if name == "terrain":
        assert d["max_rel"] <= 1e-3 and d["frac_1e4"] <= 0.01 and d["rel_mse"] <= 1e-9, d
        assert np.abs(frame.astype(int) - GOLD[name + "_frame"].astype(int)).max() <= 1
    else:
        assert d["rel_mse"] <= 2e-3 and abs(d["mean_ratio"] - 1.0) <= 5e-4, d
    return d


@pytest.mark.parametrize("name", NAMES)
def test_default_mode_within_stated_bounds(name):
    accum, frame = oracle_render(name, chain_sum=0, libm_sincos=0)
    check_default_mode(accum, frame, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_kernels_against_reference_held_images(name):
    from rendertoy3c_b200.api import Context
    desc, n = pin_scenes()[name]
    with Context(0) as g:
        accum, frame = render(g, desc, n)
        assert g.stats()["error_flags"] == 0
    check_default_mode(accum, frame, name)
    # and the kernels are the oracle's default mode exactly
    o_accum, o_frame = oracle_render(name, chain_sum=0, libm_sincos=0)
    assert np.array_equal(accum.view(np.uint32), o_accum.view(np.uint32)) and np.array_equal(frame, o_frame)


@pytest.mark.parametrize("name", NAMES)
def test_goldens_are_what_the_reference_programs_render(name):
    from ref_backend import RefShaderScene, available
    if not available():
        pytest.skip("/root/reference not present")
    desc, n = pin_scenes()[name]
    r = RefShaderScene()
    accum, frame = render(r, desc, n)
    r.close()
    assert np.array_equal(accum.view(np.uint32), GOLD[name + "_accum"].view(np.uint32)), "committed golden is stale"
    assert np.array_equal(frame, GOLD[name + "_frame"])
