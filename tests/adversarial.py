"""Adversarial geometry / rays for the bit-exact hit parity claim: duplicated and coplanar triangles
(exact-t ties -> lowest (instance, primitive) must win), zero-area and sliver triangles, axis-aligned
flat meshes, a scene far from the origin (large |coordinates| vs small features), axis-parallel rays,
rays starting exactly on surfaces and at vertices, and the same mesh instanced twice at one place."""
import numpy as np

from rendertoy3c_b200 import scenes
from rendertoy3c_b200._abi import RAY_DTYPE
from rendertoy3c_b200.scenes import IDENTITY, Geometry, Instance


def make_scene(offset=0.0, seed=3):
    rng = np.random.RandomState(seed)
    off = np.array([offset, -offset * 0.5, offset * 0.25], np.float32)
    # a 12x12 flat grid in the plane y=0 (axis-aligned, zero-thickness boxes), every quad duplicated once
    g = scenes.grid_mesh(12, 12, lambda U, V: np.stack([U * 4 - 2, 0 * U, V * 4 - 2], axis=-1))
    verts = g.verts + off
    idx = np.concatenate([g.idx, g.idx[::3]])                      # duplicated triangles: exact ties
    # slivers and zero-area triangles
    extra_v = np.array([[0, 1, 0], [1e-6, 1, 0], [0, 1, 2], [1, 1.5, 1], [1, 1.5, 1], [1, 1.5, 1], [-1, 0.5, -1], [1, 0.5, -1], [0, 0.5, -1 + 1e-5]], np.float32) + off
    base = len(verts)
    verts = np.concatenate([verts, extra_v]).astype(np.float32)
    idx = np.concatenate([idx, np.array([[base, base + 1, base + 2], [base + 3, base + 4, base + 5], [base + 6, base + 7, base + 8]], np.int32)]).astype(np.int32)
    normals = np.tile(np.array([[0, 1, 0]], np.float32), (len(verts), 1))
    uvs = np.zeros((len(verts), 2), np.float32)
    flat = Geometry("mesh", verts=verts, idx=idx, normals=normals, uvs=uvs)
    # a small random soup, instanced twice at the SAME place (identity + an exactly equal non-merged copy via keys)
    sv = (rng.rand(60, 3).astype(np.float32) * 2 - 1) * np.float32(0.8) + np.array([0, 1.2, 0], np.float32) + off
    si = rng.randint(0, 60, size=(40, 3)).astype(np.int32)
    soup = Geometry("mesh", verts=sv, idx=si, normals=np.tile(np.array([[0, 0, 1]], np.float32), (60, 1)), uvs=np.zeros((60, 2), np.float32))
    keys = np.stack([IDENTITY, IDENTITY]).astype(np.float32)       # animated instance with identical keys: same place, separate BLAS path
    inst = [Instance(0, emission=(1.0, 1.0, 1.0)), Instance(1), Instance(1, keys=keys), Instance(0, xform=IDENTITY.copy())]
    cam = scenes.Camera(eye=tuple(np.array([0.0, 3.0, 5.0]) + off), lookat=tuple(np.array([0.0, 0.5, 0.0]) + off))
    return scenes.SceneDesc("adversarial", [flat, soup], inst, [], cam, 48, 32, 8, 4), off


def make_rays(off, n=4000, seed=9):
    rng = np.random.RandomState(seed)
    rays = np.zeros(n, RAY_DTYPE)
    o = (rng.rand(n, 3).astype(np.float32) * 6 - 3) + np.array([0, 1, 0], np.float32)
    d = rng.randn(n, 3).astype(np.float32)
    k = n // 8
    d[:k] = np.eye(3, dtype=np.float32)[rng.randint(0, 3, k)] * rng.choice([-1, 1], k)[:, None].astype(np.float32)   # axis-parallel
    o[k:2 * k, 1] = 0.0                                                      # origins exactly in the flat plane
    d[2 * k:3 * k, 1] = 0.0                                                  # rays parallel to the flat plane ...
    o[2 * k:3 * k, 1] = 0.0                                                  # ... and inside it
    gx = np.round(o[3 * k:4 * k, 0] * 3) / 3
    gz = np.round(o[3 * k:4 * k, 2] * 3) / 3
    tgt = np.stack([gx, np.zeros_like(gx), gz], axis=1).astype(np.float32)   # aim exactly at grid vertices / edges
    d[3 * k:4 * k] = tgt - o[3 * k:4 * k]
    rays["o"] = o + off
    rays["d"] = d
    rays["tmin"] = 0.0
    rays["tmin"][k:2 * k] = 1e-3
    rays["tmax"] = 1e16
    rays["time"] = rng.rand(n).astype(np.float32)
    return rays


def make_stacked_scene(layers=48, n=6):
    """ONE static identity mesh (-> the single-level kernel and its deferred triangle queue): an n x n flat grid,
    `layers` copies, half of them exactly coincident (exact-t ties between many primitives: the lowest id must win)
    and half a hair apart.  A ray down the stack has dozens of triangles pending per round, far more than the
    per-warp queue holds, so the partial-queue / forced-pass path and the order-free fold are exercised."""
    g = scenes.grid_mesh(n, n, lambda U, V: np.stack([U * 2 - 1, 0 * U, V * 2 - 1], axis=-1))
    verts, idx = [], []
    for k in range(layers):
        v = g.verts.copy()
        if k % 2:
            v[:, 1] += np.float32(1e-4 * k)
        idx.append(g.idx + len(g.verts) * k)
        verts.append(v)
    verts = np.concatenate(verts).astype(np.float32)
    idx = np.concatenate(idx).astype(np.int32)
    normals = np.tile(np.array([[0, 1, 0]], np.float32), (len(verts), 1))
    uvs = np.zeros((len(verts), 2), np.float32)
    mesh = Geometry("mesh", verts=verts, idx=idx, normals=normals, uvs=uvs)
    light = scenes.grid_mesh(1, 1, lambda U, V: np.stack([U - 0.5, 0 * U + 2.0, V - 0.5], axis=-1))
    inst = [Instance(0, diffuse=(0.7, 0.6, 0.5)), Instance(1, emission=(5.0, 5.0, 5.0))]
    cam = scenes.Camera(eye=(0.3, 2.5, 1.5), lookat=(0.0, 0.0, 0.0))
    return scenes.SceneDesc("stacked", [mesh, light], inst, [], cam, 40, 30, 8, 3)


def make_stacked_rays(n=6000, seed=5):
    rng = np.random.RandomState(seed)
    rays = np.zeros(n, RAY_DTYPE)
    o = rng.rand(n, 3).astype(np.float32) * np.array([2.4, 0, 2.4], np.float32) - np.array([1.2, 0, 1.2], np.float32)
    o[:, 1] = rng.choice([-1.0, 1.5], n).astype(np.float32)          # from below and from above
    d = rng.randn(n, 3).astype(np.float32) * np.float32(0.3)
    d[:, 1] = -np.sign(o[:, 1])
    d[: n // 4, 0] = 0.0
    d[: n // 4, 2] = 0.0                                             # straight through every layer
    rays["o"] = o
    rays["d"] = d
    rays["tmin"] = 0.0
    rays["tmax"] = 1e16
    return rays
