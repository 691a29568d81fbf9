"""Quick on-GPU performance probe (not the bench): BLAS build time, rt3_trace throughput for coherent
and incoherent rays, and per-stage times of one 1080p subframe.  Usage: python tools/perf_probe.py [n] [w h]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rendertoy3c_b200 import scenes  # noqa: E402
from rendertoy3c_b200.api import Context, camera_rays, make_settings  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 708
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
scene = sys.argv[4] if len(sys.argv) > 4 else "terrain"
out = {}
t0 = time.time()
if scene == "terrain":
    d = scenes.terrain(n=n, width=w, height=h)
elif scene == "instanced":
    d = scenes.instanced(width=w, height=h)
else:
    d = scenes.motion(width=w, height=h)
out["scene_gen_s"] = time.time() - t0
g = Context(0)
for kv in os.environ.get("OPTS", "").split(","):   # e.g. OPTS=ploc=0,tlas_sah=0
    if kv:
        g.set_option(kv.split("=")[0], int(kv.split("=")[1]))
t0 = time.time()
scenes.replay(d, g)
g.sync()
out["upload_and_build_s"] = time.time() - t0
out["prims"] = d.total_instanced_prims()
uvw = g.camera_uvw(d.camera.eye, d.camera.lookat, d.camera.up, d.camera.fovy, w / h)


def time_trace(rays, any_hit, reps=5):
    dr = torch.from_numpy(rays.view(np.float32).reshape(-1, 12)).cuda()
    dh = torch.empty((len(rays), 8), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        g.sync()
        t = time.time()
        g.trace_device(dr.data_ptr(), len(rays), any_hit, dh.data_ptr())
        g.sync()
        best = min(best, time.time() - t)
    hits = dh.cpu().numpy()
    return len(rays) / best / 1e6, float((hits[:, 3].view(np.int32) >= 0).mean())


prim = camera_rays(d, uvw, w, h)
out["primary_Mrays_s"], out["primary_hit_frac"] = time_trace(prim, False)
# incoherent: random origins above the scene, random directions
rng = np.random.RandomState(1)
inc = prim.copy()
hit_o = np.asarray(d.camera.lookat, np.float32)[None, :] + (rng.rand(len(inc), 3).astype(np.float32) - 0.5) * np.array([16, 4, 16], np.float32) + np.array([0, 4, 0], np.float32)
dirs = rng.randn(len(inc), 3).astype(np.float32)
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
inc["o"] = hit_o
inc["d"] = dirs
out["incoherent_Mrays_s"], out["incoherent_hit_frac"] = time_trace(inc, False)
out["incoherent_anyhit_Mrays_s"], _ = time_trace(inc, True)

g.set_option("timing", 1)
for sf in range(2):
    g.launch_subframe(make_settings(d, uvw, sf))
    g.sync()
st = g.stats()
out["subframe_stats"] = st
g.set_option("timing", 0)
g.reset_stats()
g.sync()
t = time.time()
nsf = 4
for sf in range(nsf):
    g.launch_subframe(make_settings(d, uvw, sf))
g.sync()
dt = time.time() - t
st = g.stats()
rays = st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"]
out["render_Mrays_s"] = rays / dt / 1e6
out["render_Msamples_s"] = st["samples"] / dt / 1e6
out["render_ms_per_subframe"] = dt / nsf * 1e3
out["rays_per_sample"] = rays / st["samples"]
if os.environ.get("RT3_LIB", "").endswith("_stats.so"):   # the -DRT3_STATS twin: device counters of one 960x540 subframe
    g.debug_counters()
    uvw2 = g.camera_uvw(d.camera.eye, d.camera.lookat, d.camera.up, d.camera.fovy, 960 / 540)
    g.launch_subframe(make_settings(d, uvw2, 0, width=960, height=540))
    g.sync()
    c = g.debug_counters()
    out["counters_per_ray"] = {"wide_nodes": c[2] / max(1, c[5]), "primitive_tests": c[3] / max(1, c[5]), "rounds": c[4] / max(1, c[5]), "instance_entries": c[13] / max(1, c[5]), "rays": c[5]}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/perf_probe_%s.json" % scene, "w"), indent=1)
