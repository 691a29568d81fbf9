"""Synthetic scene descriptions for the five BASELINE.json configs.

A SceneDesc is plain numpy data (host "scene ingest", the step before the hot path) that can be
replayed into any backend exposing the rt3 operator surface — the product Context
(rendertoy3c_b200.api) or, in tests only, the CPU oracle.  Nothing here computes rays.

Reference pointers: instance/SBT mapping src/cuda/cuda_scene.h:141-147 (one identity instance per
mesh, hit-group i <-> instance i); light list src/wavefront.cpp:257-275 (every triangle of every
mesh with |Ke| >= 1e-5, object-space key-0 vertices, Q15).
"""
from dataclasses import dataclass, field

import numpy as np

IDENTITY = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)


@dataclass
class Geometry:
    kind: str  # "mesh" | "spheres" | "curves"
    verts: np.ndarray = None      # mesh [nv,3] f32
    idx: np.ndarray = None        # mesh [nt,3] i32
    normals: np.ndarray = None    # mesh [nv,3]
    uvs: np.ndarray = None        # mesh [nv,2]
    cr: np.ndarray = None         # spheres [n,4] / curves control points [ncp,4]
    seg: np.ndarray = None        # curves [nseg] i32 first control point
    degree: int = 1               # curves: 1 linear, 2 / 3 quadratic / cubic uniform B-spline, 4 Catmull-Rom, 5 cubic Bezier (rt3.h RT3_CURVE_*)
    vert_keys: np.ndarray = None  # mesh, optional [keys,nv,3]: vertex-key (deformation) motion; verts = key 0
    colors: np.ndarray = None     # mesh, optional [nv,4] vertex colours (cuda/LocalGeometry.h:99-110); normals / uvs may be None too (SDK fallbacks)

    @property
    def nprims(self):
        return len(self.idx) if self.kind == "mesh" else (len(self.cr) if self.kind == "spheres" else len(self.seg))


@dataclass
class Instance:
    geom: int
    xform: np.ndarray = field(default_factory=lambda: IDENTITY.copy())
    keys: np.ndarray = None        # [nkeys,12] or None
    t_begin: float = 0.0
    t_end: float = 1.0
    emission: tuple = (0.0, 0.0, 0.0)
    diffuse: tuple = (0.8, 0.8, 0.8)
    tex: int = -1
    tex_xform: tuple = None        # optional (scale[2], rotation (sin, cos), offset[2]) of the SDK's sampleTexture (cuda/LocalShading.h:37-54)


@dataclass
class Texture:
    rgba: np.ndarray  # [h,w,4] u8
    address: int = 0
    filter: int = 0


@dataclass
class Camera:
    eye: tuple
    lookat: tuple
    up: tuple = (0.0, 1.0, 0.0)
    fovy: float = 45.0


@dataclass
class SceneDesc:
    name: str
    geoms: list
    instances: list
    textures: list
    camera: Camera
    width: int = 512
    height: int = 512
    spp: int = 16
    max_depth: int = 4

    def total_instanced_prims(self):
        return sum(self.geoms[i.geom].nprims for i in self.instances)


def replay(desc, be):
    """Replay a SceneDesc into a backend (same call sequence as CUDAScene's ctor + buildLightSampler,
    reference src/cuda/cuda_scene.h:124-159, src/wavefront.cpp:257-275)."""
    handles = []
    for g in desc.geoms:
        if g.kind == "mesh":
            handles.append(be.mesh_create(g.verts if g.vert_keys is None else g.vert_keys, g.idx, g.normals, g.uvs))
            if g.colors is not None:
                be.mesh_set_colors(handles[-1], g.colors)
        elif g.kind == "spheres":
            handles.append(be.spheres_create(g.cr))
        else:
            handles.append(be.curves_create(g.degree, g.cr, g.seg))
    tex_ids = [be.texture_create(t.rgba, t.address, t.filter) for t in desc.textures]
    lights = []
    for inst in desc.instances:
        if inst.keys is not None:
            iid = be.append_animated_instance(handles[inst.geom], inst.keys, inst.t_begin, inst.t_end, inst.xform)
        else:
            iid = be.append_instance(handles[inst.geom], inst.xform)
        be.set_hitgroup(iid, inst.emission, inst.diffuse, tex_ids[inst.tex] if inst.tex >= 0 else -1)
        if inst.tex_xform is not None:
            be.set_texture_transform(iid, *inst.tex_xform)
        g = desc.geoms[inst.geom]
        e = np.asarray(inst.emission, dtype=np.float32)
        if g.kind == "mesh" and float(np.sqrt(np.float32(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]))) >= 1e-5:
            for tri in g.idx:
                lights.append(be.light_make(e, g.verts[tri[0]], g.verts[tri[1]], g.verts[tri[2]]))
    be.accel_build()
    if lights:
        be.set_lights(b"".join(lights), len(lights))
    return handles


# --------------------------------------------------------------------------------------------- helpers
def _quad_mesh(quads):
    """quads: [nq,4,3] -> flat-shaded mesh (4 verts per quad, 2 tris: (0,1,2),(0,2,3))."""
    quads = np.asarray(quads, dtype=np.float32)
    nq = len(quads)
    verts = quads.reshape(-1, 3).copy()
    n = np.cross(quads[:, 1] - quads[:, 0], quads[:, 2] - quads[:, 0]).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
    normals = np.repeat(n, 4, axis=0).astype(np.float32)
    uvs = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32), (nq, 1))
    base = (4 * np.arange(nq, dtype=np.int32))[:, None]
    idx = np.concatenate([base + np.array([0, 1, 2], dtype=np.int32), base + np.array([0, 2, 3], dtype=np.int32)], axis=1).reshape(-1, 3)
    return Geometry("mesh", verts=verts, idx=idx.astype(np.int32), normals=normals, uvs=uvs)


def value_noise_texture(size, seed, cells=32):
    """RGBA8 procedural value-noise texture (tileable), albedo range ~[0.2, 0.9]."""
    rng = np.random.RandomState(seed)
    lattice = rng.rand(cells, cells, 3).astype(np.float32)
    t = (np.arange(size, dtype=np.float32) + 0.5) * (cells / size)
    i0 = np.floor(t).astype(np.int64) % cells
    i1 = (i0 + 1) % cells
    f = t - np.floor(t)
    f = f * f * (3 - 2 * f)
    a = lattice[i0][:, i0] * (1 - f)[None, :, None] + lattice[i0][:, i1] * f[None, :, None]
    b = lattice[i1][:, i0] * (1 - f)[None, :, None] + lattice[i1][:, i1] * f[None, :, None]
    v = a * (1 - f)[:, None, None] + b * f[:, None, None]
    checker = ((np.arange(size)[:, None] // (size // 16) + np.arange(size)[None, :] // (size // 16)) % 2).astype(np.float32)
    v = 0.2 + 0.7 * (0.75 * v + 0.25 * checker[:, :, None] * v)
    rgba = np.empty((size, size, 4), dtype=np.uint8)
    rgba[..., :3] = np.clip(v * 255.0 + 0.5, 0, 255).astype(np.uint8)
    rgba[..., 3] = 255
    return rgba


def rigid(rng, max_angle_deg=180.0, translate=(0, 0, 0), scale=1.0):
    """random rotation (axis-angle) * scale + translation as a row-major 3x4."""
    axis = rng.randn(3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(max_angle_deg) * (2 * rng.rand() - 1)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    m = np.zeros((3, 4))
    m[:, :3] = R * scale
    m[:, 3] = translate
    return m.astype(np.float32).reshape(12)


def compose(a, b):
    """row-major 3x4 product a∘b (apply b first)."""
    A = np.vstack([a.reshape(3, 4).astype(np.float64), [0, 0, 0, 1]])
    B = np.vstack([b.reshape(3, 4).astype(np.float64), [0, 0, 0, 1]])
    return (A @ B)[:3].astype(np.float32).reshape(12)


def grid_mesh(nx, ny, pos_fn, uv_scale=1.0):
    """(nx x ny) quads; pos_fn(u,v)->[...,3]; smooth normals from accumulated face normals."""
    u = np.linspace(0.0, 1.0, nx + 1, dtype=np.float64)
    v = np.linspace(0.0, 1.0, ny + 1, dtype=np.float64)
    U, V = np.meshgrid(u, v, indexing="xy")
    P = pos_fn(U, V).astype(np.float32)  # [ny+1,nx+1,3]
    verts = P.reshape(-1, 3)
    uvs = (np.stack([U, V], axis=-1) * uv_scale).astype(np.float32).reshape(-1, 2)
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    idx = np.empty((2 * nx * ny, 3), dtype=np.int32)
    idx[0::2] = np.stack([v00, v10, v11], axis=1)
    idx[1::2] = np.stack([v00, v11, v01], axis=1)
    fn = np.cross(verts[idx[:, 1]] - verts[idx[:, 0]], verts[idx[:, 2]] - verts[idx[:, 0]]).astype(np.float64)
    nrm = np.zeros((len(verts), 3), dtype=np.float64)
    for k in range(3):
        np.add.at(nrm, idx[:, k], fn)
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    normals = (nrm / ln).astype(np.float32)
    return Geometry("mesh", verts=np.ascontiguousarray(verts), idx=idx, normals=normals, uvs=uvs)


# --------------------------------------------------------------------------------------------- C1
def cornell(width=512, height=512, spp=16, max_depth=4):
    """C1: classic Cornell box, 32 triangles in 6 meshes (white walls, red, green, light, 2 blocks)."""
    white_q = [
        [[552.8, 0, 0], [0, 0, 0], [0, 0, 559.2], [549.6, 0, 559.2]],              # floor
        [[556, 548.8, 0], [556, 548.8, 559.2], [0, 548.8, 559.2], [0, 548.8, 0]],  # ceiling
        [[549.6, 0, 559.2], [0, 0, 559.2], [0, 548.8, 559.2], [556, 548.8, 559.2]] # back
    ]
    green_q = [[[0, 0, 559.2], [0, 0, 0], [0, 548.8, 0], [0, 548.8, 559.2]]]
    red_q = [[[552.8, 0, 0], [549.6, 0, 559.2], [556, 548.8, 559.2], [556, 548.8, 0]]]
    light_q = [[[343, 548.7, 227], [343, 548.7, 332], [213, 548.7, 332], [213, 548.7, 227]]]
    short_q = [
        [[130, 165, 65], [82, 165, 225], [240, 165, 272], [290, 165, 114]],
        [[290, 0, 114], [290, 165, 114], [240, 165, 272], [240, 0, 272]],
        [[130, 0, 65], [130, 165, 65], [290, 165, 114], [290, 0, 114]],
        [[82, 0, 225], [82, 165, 225], [130, 165, 65], [130, 0, 65]],
        [[240, 0, 272], [240, 165, 272], [82, 165, 225], [82, 0, 225]],
    ]
    tall_q = [
        [[423, 330, 247], [265, 330, 296], [314, 330, 456], [472, 330, 406]],
        [[423, 0, 247], [423, 330, 247], [472, 330, 406], [472, 0, 406]],
        [[472, 0, 406], [472, 330, 406], [314, 330, 456], [314, 0, 456]],
        [[314, 0, 456], [314, 330, 456], [265, 330, 296], [265, 0, 296]],
        [[265, 0, 296], [265, 330, 296], [423, 330, 247], [423, 0, 247]],
    ]
    geoms = [_quad_mesh(q) for q in (white_q, red_q, green_q, light_q, short_q, tall_q)]
    W, R, G = (0.73, 0.73, 0.73), (0.65, 0.05, 0.05), (0.12, 0.45, 0.15)
    inst = [
        Instance(0, diffuse=W), Instance(1, diffuse=R), Instance(2, diffuse=G),
        Instance(3, diffuse=(0.78, 0.78, 0.78), emission=(17.0, 12.0, 4.0)),
        Instance(4, diffuse=W), Instance(5, diffuse=W),
    ]
    cam = Camera(eye=(278.0, 273.0, -800.0), lookat=(278.0, 273.0, 279.6), fovy=39.3)
    return SceneDesc("C1_cornell", geoms, inst, [], cam, width, height, spp, max_depth)


# --------------------------------------------------------------------------------------------- C2
def _terrain_pos(size):
    def fn(U, V):
        x = (U - 0.5) * size
        z = (V - 0.5) * size
        y = (1.6 * np.sin(0.9 * x) * np.cos(0.7 * z) + 0.8 * np.sin(2.3 * x + 1.0) * np.sin(1.9 * z)
             + 0.25 * np.sin(7.1 * x) * np.cos(6.3 * z + 0.5))
        for (cx, cz, r, hgt) in ((-3.0, 2.0, 2.0, 3.0), (4.0, -1.5, 1.5, 2.5), (0.5, -5.0, 2.5, 2.0), (-5.5, -4.0, 1.2, 2.2)):
            d2 = (x - cx) ** 2 + (z - cz) ** 2
            y = y + hgt * np.exp(-d2 / (r * r))
        return np.stack([x, y, z], axis=-1)
    return fn


def terrain(n=708, width=1920, height=1080, spp=64, max_depth=8, tex_size=2048):
    """C2: displaced tessellated grid ("terrain + blobs"), n x n quads (n=708 -> 1,002,528 tris), smooth
    normals, uvs in [0,8]^2 (wrap), one value-noise RGBA8 texture, one emissive quad above (2 lights)."""
    size = 20.0
    g = grid_mesh(n, n, _terrain_pos(size), uv_scale=8.0)
    lq = [[[-3.0, 9.0, -3.0], [3.0, 9.0, -3.0], [3.0, 9.0, 3.0], [-3.0, 9.0, 3.0]]]
    geoms = [g, _quad_mesh(lq)]
    inst = [Instance(0, tex=0), Instance(1, diffuse=(0.8, 0.8, 0.8), emission=(30.0, 28.0, 24.0))]
    cam = Camera(eye=(0.0, 9.0, 16.0), lookat=(0.0, 0.5, 0.0), fovy=45.0)
    return SceneDesc("C2_terrain_%dtris" % len(g.idx), geoms, inst, [Texture(value_noise_texture(tex_size, 1))], cam, width, height, spp, max_depth)


# --------------------------------------------------------------------------------------------- C3 / C4
def _blob_pos(U, V):
    th = V * np.pi
    ph = U * 2 * np.pi
    r = 0.45 * (1.0 + 0.15 * np.sin(5 * ph) * np.sin(4 * th) + 0.08 * np.cos(9 * th + 2 * ph))
    return np.stack([r * np.sin(th) * np.cos(ph), r * np.cos(th), r * np.sin(th) * np.sin(ph)], axis=-1)


def instanced(n_inst=1000, blob_n=224, n_spheres=1000, width=1920, height=1080, spp=64, max_depth=8):
    """C3: two-level AS: n_inst rigid instances of one blob mesh (224x224 quads -> 100,352 tris) on a jittered
    lattice + analytic spheres + one emissive quad."""
    rng = np.random.RandomState(2)
    blob = grid_mesh(blob_n, blob_n, _blob_pos)
    side = int(round(n_inst ** (1.0 / 3.0)))
    while side ** 3 < n_inst:
        side += 1
    geoms = [blob]
    inst = []
    extent = 1.2 * side
    for k in range(n_inst):
        ix, iy, iz = k % side, (k // side) % side, k // (side * side)
        c = (np.array([ix, iy, iz]) + 0.5 + 0.3 * (rng.rand(3) - 0.5)) * 1.2 - extent / 2
        col = tuple(float(x) for x in (0.25 + 0.6 * rng.rand(3)))
        inst.append(Instance(0, xform=rigid(rng, 180.0, c, 0.8 + 0.4 * rng.rand()), diffuse=col))
    rng3 = np.random.RandomState(3)
    if n_spheres > 0:
        cr = np.empty((n_spheres, 4), dtype=np.float32)
        cr[:, :3] = (rng3.rand(n_spheres, 3) - 0.5) * extent
        cr[:, 3] = 0.2 + 0.4 * rng3.rand(n_spheres)
        cr[:, 3] *= 0.5
        geoms.append(Geometry("spheres", cr=cr))
        inst.append(Instance(len(geoms) - 1, diffuse=(0.7, 0.7, 0.75)))
    h = extent / 2 + 2.0
    w = extent / 2
    geoms.append(_quad_mesh([[[-w, h, -w], [w, h, -w], [w, h, w], [-w, h, w]]]))
    inst.append(Instance(len(geoms) - 1, diffuse=(0.8, 0.8, 0.8), emission=(6.0, 6.0, 5.5)))
    geoms.append(_quad_mesh([[[-2 * w, -h, -2 * w], [-2 * w, -h, 2 * w], [2 * w, -h, 2 * w], [2 * w, -h, -2 * w]]]))
    inst.append(Instance(len(geoms) - 1, diffuse=(0.5, 0.5, 0.5)))
    cam = Camera(eye=(0.0, 0.3 * extent, 1.35 * extent), lookat=(0.0, -0.05 * extent, 0.0), fovy=45.0)
    return SceneDesc("C3_instanced_%dx%d" % (n_inst, len(blob.idx)), geoms, inst, [], cam, width, height, spp, max_depth)


def motion(n_inst=64, blob_n=224, n_spheres=256, n_curves=10000, width=1920, height=1080, spp=128, max_depth=8):
    """C4: per-instance 2-key matrix motion (key1 = key0 ∘ small random rigid), triangles + spheres + linear curves."""
    rng = np.random.RandomState(4)
    blob = grid_mesh(blob_n, blob_n, _blob_pos)
    geoms = [blob]
    inst = []
    side = 4
    extent = 1.4 * side
    for k in range(n_inst):
        ix, iy, iz = k % side, (k // side) % side, k // (side * side)
        c = (np.array([ix, iy, iz]) + 0.5 + 0.3 * (rng.rand(3) - 0.5)) * 1.4 - extent / 2
        key0 = rigid(rng, 180.0, (0, 0, 0), 0.9 + 0.3 * rng.rand())
        delta = rigid(rng, 10.0, 0.5 * (rng.rand(3) - 0.5), 1.0)
        key1 = compose(delta, key0)
        stat = IDENTITY.copy()
        stat[[3, 7, 11]] = c
        col = tuple(float(x) for x in (0.25 + 0.6 * rng.rand(3)))
        inst.append(Instance(0, xform=stat, keys=np.stack([key0, key1]).astype(np.float32), diffuse=col))
    rng3 = np.random.RandomState(3)
    if n_spheres > 0:
        cr = np.empty((n_spheres, 4), dtype=np.float32)
        cr[:, :3] = (rng3.rand(n_spheres, 3) - 0.5) * extent
        cr[:, 3] = 0.1 + 0.2 * rng3.rand(n_spheres)
        geoms.append(Geometry("spheres", cr=cr))
        k0 = IDENTITY.copy()
        k1 = IDENTITY.copy()
        k1[7] = 0.3
        inst.append(Instance(len(geoms) - 1, keys=np.stack([k0, k1]).astype(np.float32), diffuse=(0.7, 0.7, 0.75)))
    if n_curves > 0:
        rng5 = np.random.RandomState(5)
        seg_per = 5
        n_strands = max(1, n_curves // seg_per)
        cps, segs = [], []
        for s in range(n_strands):
            root = np.array([(rng5.rand() - 0.5) * extent, -extent / 2 - 1.0, (rng5.rand() - 0.5) * extent])
            d = np.array([0.0, 1.0, 0.0]) + 0.5 * (rng5.rand(3) - 0.5)
            p = root.copy()
            r0 = 0.01 + 0.02 * rng5.rand()
            base = len(cps)
            for j in range(seg_per + 1):
                cps.append([p[0], p[1], p[2], r0 * (1.0 - 0.6 * j / seg_per)])
                d = d + 0.25 * (rng5.rand(3) - 0.5)
                p = p + 0.12 * d / np.linalg.norm(d)
            segs.extend(range(base, base + seg_per))
        geoms.append(Geometry("curves", cr=np.asarray(cps, dtype=np.float32), seg=np.asarray(segs, dtype=np.int32)))
        inst.append(Instance(len(geoms) - 1, diffuse=(0.6, 0.45, 0.25)))
    h = extent / 2 + 2.0
    w = extent / 2
    geoms.append(_quad_mesh([[[-w, h, -w], [w, h, -w], [w, h, w], [-w, h, w]]]))
    inst.append(Instance(len(geoms) - 1, diffuse=(0.8, 0.8, 0.8), emission=(8.0, 8.0, 7.5)))
    geoms.append(_quad_mesh([[[-2 * w, -h + 1.0, -2 * w], [-2 * w, -h + 1.0, 2 * w], [2 * w, -h + 1.0, 2 * w], [2 * w, -h + 1.0, -2 * w]]]))
    inst.append(Instance(len(geoms) - 1, diffuse=(0.5, 0.5, 0.5)))
    cam = Camera(eye=(0.0, 0.2 * extent, 1.6 * extent), lookat=(0.0, -0.1 * extent, 0.0), fovy=45.0)
    return SceneDesc("C4_motion", geoms, inst, [], cam, width, height, spp, max_depth)


def deforming(blob_n=24, width=160, height=90, spp=16, max_depth=6, keys=3):
    """vertex-key motion: a blob whose vertices move through `keys` key-frames over the
    shutter, next to a static copy (transformed instance), above an emissive-lit floor."""
    blob = grid_mesh(blob_n, blob_n, _blob_pos)
    vk = np.stack([blob.verts * np.float32(1.0 + 0.25 * k) + np.array([0.15 * k, 0.1 * k * k, 0.0], np.float32) for k in range(keys)]).astype(np.float32)
    moving = Geometry("mesh", verts=np.ascontiguousarray(vk[0]), idx=blob.idx, normals=blob.normals, uvs=blob.uvs, vert_keys=vk)
    geoms = [moving, blob]
    stat = IDENTITY.copy()
    stat[[3, 7, 11]] = (1.4, 0.0, 0.0)
    inst = [Instance(0, diffuse=(0.8, 0.4, 0.3)), Instance(1, xform=stat, diffuse=(0.3, 0.5, 0.8))]
    geoms.append(_quad_mesh([[[-2, 2.5, -2], [2, 2.5, -2], [2, 2.5, 2], [-2, 2.5, 2]]]))
    inst.append(Instance(2, diffuse=(0.8, 0.8, 0.8), emission=(8.0, 8.0, 7.0)))
    geoms.append(_quad_mesh([[[-4, -0.8, -4], [-4, -0.8, 4], [4, -0.8, 4], [4, -0.8, -4]]]))
    inst.append(Instance(3, diffuse=(0.6, 0.6, 0.6)))
    cam = Camera(eye=(0.7, 0.8, 3.6), lookat=(0.7, 0.1, 0.0), fovy=45.0)
    return SceneDesc("N2_deforming", geoms, inst, [], cam, width, height, spp, max_depth)


def splines(n_strands=40, width=80, height=48, spp=16, max_depth=4, seed=11):
    """spline curve strands of every type rt3_curves_create offers besides linear — quadratic and cubic B-spline, Catmull-Rom,
    cubic Bezier (random walks of control points, varying radius) — over a lit floor"""
    rng = np.random.RandomState(seed)
    geoms, inst = [], []
    for degree in (2, 3, 4, 5):
        cps, segs = [], []
        ncp = 3 if degree == 2 else 4           # control points per segment
        for _ in range(n_strands if degree < 4 else n_strands // 2):
            n = rng.randint(ncp + 1, ncp + 6)
            p = np.cumsum(rng.randn(n, 3).astype(np.float32) * np.float32(0.35), axis=0) + (rng.rand(3).astype(np.float32) * 3 - 1.5) * np.array([1, 0.3, 1], np.float32)
            p[:, 1] = np.abs(p[:, 1]) + 0.2
            r = (0.03 + 0.05 * rng.rand(n)).astype(np.float32)
            base = sum(len(c) for c in cps)
            cps.append(np.concatenate([p, r[:, None]], axis=1))
            segs.extend(range(base, base + n - ncp + 1))
        geoms.append(Geometry("curves", cr=np.concatenate(cps).astype(np.float32), seg=np.array(segs, np.int32), degree=degree))
        inst.append(Instance(len(geoms) - 1, diffuse=[(0.7, 0.5, 0.3), (0.3, 0.6, 0.7), (0.5, 0.7, 0.3), (0.7, 0.3, 0.6)][degree - 2]))
    geoms.append(_quad_mesh([[[-3, 3.0, -3], [3, 3.0, -3], [3, 3.0, 3], [-3, 3.0, 3]]]))
    inst.append(Instance(len(geoms) - 1, diffuse=(0.8, 0.8, 0.8), emission=(7.0, 7.0, 6.5)))
    geoms.append(_quad_mesh([[[-5, 0, -5], [-5, 0, 5], [5, 0, 5], [5, 0, -5]]]))
    inst.append(Instance(len(geoms) - 1, diffuse=(0.6, 0.6, 0.6)))
    cam = Camera(eye=(0.0, 2.2, 5.0), lookat=(0.0, 0.6, 0.0), fovy=45.0)
    return SceneDesc("splines", geoms, inst, [], cam, width, height, spp, max_depth)


def fallbacks(blob_n=10, width=80, height=48, spp=16, max_depth=5):
    """the deforming scene with the SDK's optional vertex attributes exercised (cuda/LocalGeometry.h:99-124,150-158): the moving
    blob has no normals and carries vertex colours, its static (transformed) copy has no texcoords, the floor has neither and
    is textured, so its texture is addressed by the barycentrics"""
    d = deforming(blob_n=blob_n, width=width, height=height, spp=spp, max_depth=max_depth)
    rng = np.random.RandomState(17)
    d.geoms[0].normals = None
    d.geoms[0].colors = rng.rand(len(d.geoms[0].verts), 4).astype(np.float32)
    d.geoms[1].uvs = None
    d.geoms[1].colors = rng.rand(len(d.geoms[1].verts), 4).astype(np.float32)
    d.geoms[3].normals = None
    d.geoms[3].uvs = None
    d.textures.append(Texture(value_noise_texture(32, 5, cells=8)))
    d.instances[3].tex = 0
    d.name = "fallbacks"
    return d


def by_name(name, **kw):
    return {"cornell": cornell, "terrain": terrain, "instanced": instanced, "motion": motion, "deforming": deforming}[name](**kw)


def _png_bytes(rgb_top_down):
    """8-bit RGB PNG (sub filter, zlib) of an [h,w,3] array, standard library only"""
    import struct
    import zlib
    a = np.ascontiguousarray(rgb_top_down, dtype=np.uint8)
    h, w = a.shape[:2]
    d = a.astype(np.int16)
    d[:, 1:] -= a[:, :-1].astype(np.int16)                        # filter type 1 (sub)
    raw = np.concatenate([np.ones((h, 1), np.uint8), (d & 255).astype(np.uint8).reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")


def write_obj(desc, path, tex_format="ppm", key=0):
    """Export the mesh instances of a SceneDesc (identity transforms only) as .obj + .mtl (+ binary PPM or
    PNG textures), one `o` shape and one material per instance, for the C++ host (host/wavefront.cpp) and the
    loader round-trip tests.  Rows of textures are written top-down (the loader flips them back)."""
    import os
    base = os.path.splitext(path)[0]
    with open(path, "w") as f, open(base + ".mtl", "w") as m:
        f.write("mtllib %s.mtl\n" % os.path.basename(base))
        vo = 1
        for i, inst in enumerate(desc.instances):
            g = desc.geoms[inst.geom]
            assert g.kind == "mesh" and inst.keys is None and np.array_equal(inst.xform, IDENTITY)
            m.write("newmtl m%d\nKd %.9g %.9g %.9g\nKe %.9g %.9g %.9g\n" % ((i,) + tuple(np.float32(x) for x in inst.diffuse) + tuple(np.float32(x) for x in inst.emission)))
            if inst.tex >= 0:
                tname = "%s_tex%d.%s" % (os.path.basename(base), inst.tex, tex_format)
                m.write("map_Kd %s\n" % tname)
                rgba = desc.textures[inst.tex].rgba
                with open(os.path.join(os.path.dirname(path), tname), "wb") as t:
                    if tex_format == "png":
                        t.write(_png_bytes(rgba[::-1, :, :3]))
                    else:
                        t.write(b"P6\n%d %d\n255\n" % (rgba.shape[1], rgba.shape[0]))
                        t.write(np.ascontiguousarray(rgba[::-1, :, :3]).tobytes())
            f.write("o shape%d\nusemtl m%d\n" % (i, i))
            for v in (g.verts if g.vert_keys is None else g.vert_keys[key]):  # key-frame files: same topology, other positions
                f.write("v %.9g %.9g %.9g\n" % tuple(v))
            for n in g.normals:
                f.write("vn %.9g %.9g %.9g\n" % tuple(n))
            for t in g.uvs:
                f.write("vt %.9g %.9g\n" % tuple(t))
            for tri in g.idx:
                f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % tuple(int(x) + vo for x in np.repeat(tri, 3)))
            vo += len(g.verts)
