"""Small all-features workload for compute-sanitizer: Cornell (single-level), instanced (TLAS + merged +
spheres), motion (keys + curves): build, rt3_trace closest/any, one 64x36 subframe each."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from parity_common import SMALL, random_rays  # noqa: E402
from rendertoy3c_b200 import scenes  # noqa: E402
from rendertoy3c_b200.api import Context, make_settings  # noqa: E402

for name in ("cornell", "instanced", "motion", "terrain"):
    d = SMALL[name]()
    with Context(0) as g:
        scenes.replay(d, g)
        uvw = g.camera_uvw(d.camera.eye, d.camera.lookat, d.camera.up, d.camera.fovy, 64 / 36)
        rays = random_rays(d, 2000, 5)
        h = g.trace(rays)
        a = g.trace(rays, any_hit=True)
        g.launch_subframe(make_settings(d, uvw, 0, width=64, height=36))
        img = g.download_accum()
        print(name, int((h["prim"] >= 0).sum()), int((a["prim"] >= 0).sum()), float(np.nanmean(img[..., :3])), g.stats()["error_flags"])
