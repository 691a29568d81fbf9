"""N>1 host logic on CPU: world_size 2, gloo, kernel-logic simulator.  Rank r renders subframes r, r+N, ...
into a SUM accumulation buffer; one all-reduce of the buffer + rt3_finalize_accum must reproduce the
single-rank render of the same subframe set (up to fp32 summation order)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from rendertoy3c_b200 import scenes
from rendertoy3c_b200.api import Context, make_settings

SUBFRAMES = 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _render(lib, rank, world, accum_mode):
    desc = scenes.cornell(width=40, height=40)
    g = Context(0, lib_path=lib)
    scenes.replay(desc, g)
    uvw = g.camera_uvw(desc.camera.eye, desc.camera.lookat, desc.camera.up, desc.camera.fovy, 1.0)
    g.clear_accum()
    for sf in range(rank, SUBFRAMES, world):
        g.launch_subframe(make_settings(desc, uvw, sf, accum_mode=accum_mode))
    g.sync()
    return g


def _worker(rank, world, port, lib, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = _render(lib, rank, world, 1)
    ptr, n = g.accum_device_ptr()  # simulator: device memory is host memory
    buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,))
    t = torch.from_numpy(buf)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    g.finalize_accum(SUBFRAMES)
    g.sync()
    if rank == 0:
        np.save(out_path, np.concatenate([g.download_accum().reshape(-1), g.download_frame().reshape(-1).astype(np.float32)]))
    st = g.stats()
    cnt = torch.tensor([st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"], st["samples"]], dtype=torch.float64)
    dist.all_reduce(cnt)
    if rank == 0:
        np.save(out_path + ".cnt.npy", cnt.numpy())
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank(emul_lib, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(2, _free_port(), emul_lib, out), nprocs=2, join=True)
    multi = np.load(out)
    g = _render(emul_lib, 0, 1, 1)
    st = g.stats()
    g.finalize_accum(SUBFRAMES)
    single = np.concatenate([g.download_accum().reshape(-1), g.download_frame().reshape(-1).astype(np.float32)])
    na = 40 * 40 * 4
    np.testing.assert_allclose(multi[:na], single[:na], rtol=2e-6, atol=1e-7)
    assert np.abs(multi[na:] - single[na:]).max() <= 1          # 8-bit frame
    cnt = np.load(out + ".cnt.npy")
    assert int(cnt[0]) == st["rays_primary"] + st["rays_bounce"] + st["rays_shadow"] and int(cnt[1]) == st["samples"]
    # and the SUM/finalize path agrees with the reference-style running mean (mode 0) of the same subframes
    g0 = _render(emul_lib, 0, 1, 0)
    np.testing.assert_allclose(g0.download_accum().reshape(-1), single[:na], rtol=3e-6, atol=1e-7)
