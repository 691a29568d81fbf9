import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import adversarial
from parity_common import build_pair
from rendertoy3c_b200.api import Context, make_settings
desc, off = adversarial.make_scene(0.0)
with Context(0) as g:
    o = build_pair(desc, g)
    uvw = o.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, desc.width/desc.height)
    for sf in range(1):
        rs = make_settings(desc, uvw, sf)
        g.launch_subframe(rs); o.launch_subframe(rs)
    a, b = g.download_accum(), o.download_accum()
    d = (a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))
    ys, xs, cs = np.nonzero(d)
    print('ndiff', d.sum(), 'nan gpu', np.isnan(a).sum(), 'nan cpu', np.isnan(b).sum(), 'inf gpu', np.isinf(a).sum(), 'inf cpu', np.isinf(b).sum())
    for y,x,c in list(zip(ys,xs,cs))[:10]: print(y,x,c,a[y,x,c],b[y,x,c])
    print(g.stats(), o.stats())
with Context(0) as g:
    o = build_pair(desc, g)
    for sf in range(2):
        rs = make_settings(desc, uvw, sf)
        g.launch_subframe(rs); o.launch_subframe(rs)
    a, b = g.download_accum(), o.download_accum()
    fa, fb = g.download_frame().astype(int), o.download_frame().astype(int)
    bad = np.abs(fa - fb).max(axis=2) > 1
    ys, xs = np.nonzero(bad)
    print('frame bad', bad.sum())
    for y, x in list(zip(ys, xs))[:8]: print(y, x, a[y, x], b[y, x], fa[y, x], fb[y, x])
