"""Writes tests/golden/ref_images.npz: float accumulation buffers and 8-bit frames rendered by the REFERENCE'S OWN device
programs (src/shader/raygen.cu, closehit_radiance.cu, miss.cu, test.cu, src/light.h, cuda/random.h, cuda/helpers.h)
compiled where they lie and run on the host through oracle/ref_shim (make -C oracle ref_shaders).  Traversal and texel
fetches come from the oracle (the reference has no source for them).  Run by hand in the build container:
    python tests/golden/make_ref_images.py
tests/test_reference_pins.py compares the oracle (and, on the GPU, the kernels) with these images."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rendertoy3c_b200 import scenes  # noqa: E402
from rendertoy3c_b200.api import make_settings  # noqa: E402


def pin_scenes():
    """name -> (SceneDesc, subframes): the reference's scope — triangle meshes under identity instances, unbounded depth"""
    return {
        "cornell": (scenes.cornell(width=48, height=48), 3),                               # C1 geometry: Ke, Kd, uniform-light NEE, long paths in a closed box
        "terrain": (scenes.terrain(n=12, width=48, height=32, tex_size=16), 2),            # C2 geometry: map_Kd texture with wrap, open sky (miss program)
    }


def render(backend, desc, subframes):
    scenes.replay(desc, backend)
    uvw = backend.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
    per_subframe = []
    for sf in range(subframes):
        rs = make_settings(desc, uvw, sf)
        rs.max_depth = 0      # unbounded, like the reference's for(;;) (raygen.cu:48)
        backend.launch_subframe(rs)
        per_subframe.append(backend.download_accum())
    return np.stack(per_subframe), backend.download_frame()


if __name__ == "__main__":
    from ref_backend import RefShaderScene, available
    if not available():
        sys.exit("/root/reference is not present: the reference's programs cannot be compiled here")
    out = {}
    for name, (desc, n) in pin_scenes().items():
        r = RefShaderScene()
        accum, frame = render(r, desc, n)
        out[name + "_accum"] = accum      # [subframe][h][w][4] float32, the running mean after each launch
        out[name + "_frame"] = frame      # [h][w][4] u8 after the last launch
        print("%-8s %s  mean %.6f" % (name, accum.shape, float(accum[-1, ..., :3].mean())))
        r.close()
    np.savez_compressed(os.path.join(HERE, "ref_images.npz"), **out)
    print("wrote tests/golden/ref_images.npz (%.1f KB)" % (os.path.getsize(os.path.join(HERE, "ref_images.npz")) / 1024.0))
