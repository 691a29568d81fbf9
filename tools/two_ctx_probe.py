import sys, os, time, json
sys.path.insert(0,'/root/repo')
import torch
from rendertoy3c_b200 import scenes
from rendertoy3c_b200.api import Context, make_settings
d = scenes.terrain()
ctxs=[Context(0) for _ in range(int(sys.argv[1]))]
for g in ctxs: scenes.replay(d,g)
uvw = ctxs[0].camera_uvw(d.camera.eye, d.camera.lookat, d.camera.up, d.camera.fovy, d.width/d.height)
def run(nsub):
    for g in ctxs: g.reset_stats()
    torch.cuda.synchronize(); t=time.perf_counter()
    for sf in range(nsub):
        ctxs[sf % len(ctxs)].launch_subframe(make_settings(d, uvw, sf, samples_per_launch=8, accum_mode=1, max_depth=8))
    for g in ctxs: g.sync()
    dt=time.perf_counter()-t
    rays=sum(g.stats()[k] for g in ctxs for k in ("rays_primary","rays_bounce","rays_shadow"))
    return rays/dt/1e6, dt/nsub*1e3
run(6)
print(len(ctxs), "context(s):", run(40))
