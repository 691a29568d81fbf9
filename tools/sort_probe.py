"""How much would binning bounce rays by direction octant (or finer) buy?  Same rays, different order."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rendertoy3c_b200 import scenes  # noqa: E402
from rendertoy3c_b200.api import Context, camera_rays  # noqa: E402

w, h = 1920, 1080
d = scenes.terrain(n=708, width=w, height=h, tex_size=64)
g = Context(0)
scenes.replay(d, g)
uvw = g.camera_uvw(d.camera.eye, d.camera.lookat, d.camera.up, d.camera.fovy, w / h)
prim = camera_rays(d, uvw, w, h)
prim = np.tile(prim, 4)  # 4 spl in pixel-major order would interleave; here: 4 copies back to back (same coherence per copy)
hp = g.trace(prim)
hit = hp["prim"] >= 0
rng = np.random.RandomState(1)
inc = prim[hit].copy()
inc["o"] = prim["o"][hit] + prim["d"][hit] * hp["t"][hit][:, None]
dirs = rng.randn(len(inc), 3).astype(np.float32)
dirs[:, 1] = np.abs(dirs[:, 1]) * 0.7
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
inc["d"] = dirs
inc["tmin"] = 0.01


def run(rays, any_hit, reps=3):
    dr = torch.from_numpy(rays.view(np.float32).reshape(-1, 12)).cuda()
    dh = torch.empty((len(rays), 8), dtype=torch.float32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    st = torch.cuda.ExternalStream(g.stream())
    best = 1e9
    for _ in range(reps):
        g.sync()
        ev[0].record(st)
        g.trace_device(dr.data_ptr(), len(rays), any_hit, dh.data_ptr())
        ev[1].record(st)
        g.sync()
        best = min(best, ev[0].elapsed_time(ev[1]))
    return len(rays) / best / 1e3


octant = (inc["d"][:, 0] >= 0).astype(np.int32) | ((inc["d"][:, 1] >= 0).astype(np.int32) << 1) | ((inc["d"][:, 2] >= 0).astype(np.int32) << 2)
orders = {"as produced (pixel order)": np.arange(len(inc)), "binned by direction octant (stable)": np.argsort(octant, kind="stable")}
# finer: octant, then 16 sub-bins of the dominant direction angle
az = ((np.arctan2(inc["d"][:, 2], inc["d"][:, 0]) + np.pi) / (2 * np.pi) * 16).astype(np.int32).clip(0, 15)
orders["binned by 16 azimuth sectors (stable)"] = np.argsort(az, kind="stable")
# block-local binning: octant sort inside blocks of 4096 consecutive rays (what a shade CTA could do in shared memory)
blk = np.arange(len(inc)) // 4096
orders["octant bins inside blocks of 4096"] = np.lexsort((octant, blk))
blk = np.arange(len(inc)) // 256
orders["octant bins inside blocks of 256"] = np.lexsort((octant, blk))
for name, o in orders.items():
    r = inc[o]
    print("%-42s closest %7.1f  any-hit %7.1f Mrays/s" % (name, run(r, False), run(r, True)))
