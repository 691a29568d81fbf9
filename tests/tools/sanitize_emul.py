"""Every small test scene through the kernel-logic simulator built with AddressSanitizer + UndefinedBehaviorSanitizer
(tests/test_sanitizers.py builds it and runs this file with the sanitizer runtimes preloaded): builds with the tuning
switches both ways, closest / any-hit queries, LocalGeometry, all three estimator modes.  compute-sanitizer is not available
on the GPU pool; this is the memory-safety check of the kernel bodies that is."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_common import SMALL, random_rays  # noqa: E402
from rendertoy3c_b200 import scenes  # noqa: E402
from rendertoy3c_b200.api import Context, camera_rays, make_settings  # noqa: E402

lib = sys.argv[1]
for name in sorted(SMALL):
    for opts in ({}, {"split": 2}, {"flatten": 0}, {"ploc": 1}, {"tlas_sah": 0, "sah_collapse": 0, "bsphere_cull": 0}):
        d = SMALL[name]()
        with Context(0, lib_path=lib) as g:
            for k, v in opts.items():
                g.set_option(k, v)
            scenes.replay(d, g)
            uvw = g.camera_uvw(d.camera.eye, d.camera.lookat, d.camera.up, d.camera.fovy, 32 / 18)
            rays = np.concatenate([random_rays(d, 1500, 5), camera_rays(d, uvw, 32, 18)])
            h = g.trace(rays)
            g.trace(rays, any_hit=True)
            g.get_local_geometry(rays, h)
            for mode in (0, 1, 2):
                g.launch_subframe(make_settings(d, uvw, 0, width=32, height=18, mode=mode))
            g.download_accum()
            g.download_frame()
            assert g.stats()["error_flags"] == 0
print("SANITIZED RUN COMPLETE")
