"""Probe: what does C3 cost when its static triangle instances are flattened into one world-space mesh?
(1000 x 100k triangles = 100M triangles, ~6 GB of BVH + triangle records).  Usage: python tools/flatten_probe.py [n_inst] [keep_spheres]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rendertoy3c_b200 import scenes  # noqa: E402
from rendertoy3c_b200.api import Context, make_settings  # noqa: E402
from bench import time_subframes  # noqa: E402

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
keep_spheres = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = scenes.instanced(n_inst=n_inst)
out = {"n_inst": n_inst, "keep_spheres": keep_spheres}


def run(desc, tag):
    t0 = time.time()
    g = Context(0)
    scenes.replay(desc, g)
    g.sync()
    out[tag + "_build_s"] = time.time() - t0
    out[tag + "_mem_GB"] = (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9
    uvw = g.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width / desc.height)
    stream = torch.cuda.ExternalStream(g.stream(), device=torch.device("cuda", 0))
    ms, st = time_subframes(g, lambda i: make_settings(desc, uvw, i, samples_per_launch=8), 2, 3, stream, torch)
    rays = st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"]
    out[tag + "_Mrays_s"] = rays / (ms * 1e-3) / 1e6
    out[tag + "_ms_per_step"] = ms / 3
    out[tag + "_rays"] = {k: st["rays_" + k] // 3 for k in ("primary", "bounce", "shadow")}
    out[tag + "_error_flags"] = st["error_flags"]
    g.set_option("timing", 1)
    g.reset_stats()
    g.launch_subframe(make_settings(desc, uvw, 9, samples_per_launch=8))
    g.sync()
    s = g.stats()
    out[tag + "_stage_ms"] = {k: s[k] for k in ("ms_generate", "ms_extend", "ms_shade", "ms_connect", "ms_resolve", "ms_total")}
    g.close()
    print(json.dumps(out), flush=True)


if os.environ.get("BASE", "1") == "1":
    run(d, "instanced")

# flatten: every static instance of a plain triangle mesh -> one mesh, one identity instance
t0 = time.time()
V, N, U, I = [], [], [], []
base = 0
rest = []
for inst in d.instances:
    g = d.geoms[inst.geom]
    if g.kind != "mesh" or inst.keys is not None or g.vert_keys is not None or np.any(np.asarray(inst.emission) > 0):
        rest.append(inst)
        continue
    m = np.asarray(inst.xform, np.float32).reshape(3, 4)
    V.append(g.verts @ m[:, :3].T + m[:, 3])
    N.append(g.normals @ np.linalg.inv(m[:, :3].astype(np.float64)).astype(np.float32))   # (M^-1)^T n as a row-vector product
    U.append(g.uvs)
    I.append(g.idx + base)
    base += len(g.verts)
flat = scenes.Geometry("mesh", verts=np.concatenate(V).astype(np.float32), idx=np.concatenate(I).astype(np.int32),
                       normals=np.concatenate(N).astype(np.float32), uvs=np.concatenate(U).astype(np.float32))
geoms = list(d.geoms) + [flat]
insts = [scenes.Instance(len(geoms) - 1, diffuse=(0.55, 0.55, 0.55))]
for inst in rest:
    if d.geoms[inst.geom].kind == "spheres" and not keep_spheres:
        continue
    insts.append(inst)
fd = scenes.SceneDesc("C3_flattened", geoms, insts, d.textures, d.camera, d.width, d.height, d.spp, d.max_depth)
out["flatten_host_s"] = time.time() - t0
out["flat_triangles"] = int(len(flat.idx))
run(fd, "flat")
