"""Regenerates tests/golden/oracle_golden.json: SHA-256 of oracle hits and accumulation buffers on the
small seeded scenes.  Pins the oracle against silent change; the GPU tests compare against the same file
so the B200 box checks the kernels against committed vectors as well as against the live oracle.
Run from the repo root: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def compute(ob):
    from parity_common import SMALL, random_rays
    from rendertoy3c_b200 import scenes
    from rendertoy3c_b200.api import camera_rays, make_settings
    out = {}
    for name in sorted(SMALL):
        desc = SMALL[name]()
        o = ob.OracleScene()
        scenes.replay(desc, o)
        uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
        rays = np.concatenate([camera_rays(desc, uvw, 32, 32), random_rays(desc, 1000, 31)])
        h = o.trace(rays, accel=0)
        core = np.stack([h["t"].view(np.uint32), h["u"].view(np.uint32), h["v"].view(np.uint32), h["prim"].view(np.uint32), h["inst"].view(np.uint32)])
        for sf in range(2):
            o.launch_subframe(make_settings(desc, uvw, sf))
        out[name] = {
            "hits_sha256": hashlib.sha256(core.tobytes()).hexdigest(),
            "n_hit": int((h["prim"] >= 0).sum()),
            "accum_sha256": hashlib.sha256(o.download_accum().tobytes()).hexdigest(),
            "rays": [int(o.stats()[k]) for k in ("rays_primary", "rays_bounce", "rays_shadow")],
        }
    return out


if __name__ == "__main__":
    import oracle_backend as ob
    json.dump(compute(ob), open(os.path.join(HERE, "oracle_golden.json"), "w"), indent=1, sort_keys=True)
    print(open(os.path.join(HERE, "oracle_golden.json")).read())
