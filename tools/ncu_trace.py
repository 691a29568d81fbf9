"""Small fixed workload for ncu: terrain BLAS (n x n quads), then three rt3_trace launches —
(1) 1080p primary rays, closest hit; (2) incoherent rays, closest hit; (3) incoherent rays, any hit.
Usage: python tools/ncu_trace.py [n] ; under ncu use -k regex:k_traverse"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rendertoy3c_b200 import scenes  # noqa: E402
from rendertoy3c_b200.api import Context, camera_rays  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 708
w, h = 1920, 1080
scene = os.environ.get("SCENE", "terrain")
d = scenes.terrain(n=n, width=w, height=h, tex_size=64) if scene == "terrain" else (scenes.instanced(width=w, height=h) if scene == "instanced" else scenes.motion(width=w, height=h))
g = Context(0)
for kv in os.environ.get("OPTS", "").split(","):
    if kv:
        g.set_option(kv.split("=")[0], int(kv.split("=")[1]))
scenes.replay(d, g)
uvw = g.camera_uvw(d.camera.eye, d.camera.lookat, d.camera.up, d.camera.fovy, w / h)
prim = camera_rays(d, uvw, w, h)
# bounce-like incoherent rays: origins = primary hit points, cosine-ish random directions in the upper hemisphere
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
inc["time"] = rng.rand(len(inc)).astype(np.float32)
prim["time"] = rng.rand(len(prim)).astype(np.float32)


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
    dc = g.debug_counters()
    if dc[5]:
        print("   per ray: %.1f wide nodes, %.1f prim tests, %.1f rounds" % (dc[2] / dc[5], dc[3] / dc[5], dc[4] / dc[5]))
        if dc[6]:
            print("   per warp round: no-triangle %.2f, <=1 per lane %.2f, lanes with triangles %.1f, triangles %.1f, busy lanes %.1f" %
                  (dc[7] / dc[6], dc[8] / dc[6], dc[9] / dc[6], dc[10] / dc[6], dc[11] / dc[6]))
    return len(rays) / best / 1e3


reps = int(os.environ.get("REPS", "3"))
repl = int(os.environ.get("REPL", "1"))
if repl > 1:
    perm = np.random.RandomState(9).permutation(len(inc) * repl) % len(inc)
    inc = inc[perm]  # more incoherent work per launch (shuffled copies)
    prim = np.tile(prim, repl)
print("primary closest  %.1f Mrays/s (%d rays)" % (run(prim, False, reps), len(prim)))
print("bounce  closest  %.1f Mrays/s (%d rays)" % (run(inc, False, reps), len(inc)))
print("bounce  any-hit  %.1f Mrays/s" % run(inc, True, reps))
print("stats", g.stats()["max_stack_depth"], g.stats()["error_flags"])
