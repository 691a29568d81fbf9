// ORACLE — TEST INFRASTRUCTURE ONLY (see rt3o_math.hpp header).
//
// Scalar CPU restatement of rendertoy3o's path (Appendix A of SURVEY.md, draw for draw):
//   raygen + path loop + RR + accumulate : src/shader/raygen.cu:14-87
//   ray time draws                       : src/shader/shader_common.h:64,125
//   closest-hit shading + NEE            : src/shader/closehit_radiance.cu:60-160
//   miss                                 : src/shader/miss.cu:29-32 + src/shader/test.cu:5
//   instance / SBT mapping               : src/cuda/cuda_accel.h:75-85, src/cuda/cuda_scene.h:60-82
//   texture semantics (point, wrap)      : src/cuda/cuda_texture.h:52-74 (Q9)
// Intersection is brute force over every instance x primitive (accel=0) or a binned-SAH
// BVH2 with padded boxes (accel=1, validated against accel=0 in tests/); tie-break on equal
// t: lowest instance id, then lowest primitive id.
// PARITY NOTE: the reference ships no tests/goldens and cannot be built here (OptiX), so
// only the KAT-level functions are pinned by reference code (tests/golden/ref_kat.json);
// hits and images are "parity unpinned" by the reference and pinned by this oracle.
#include "rt3o.h"
#include "rt3o_math.hpp"
#include "rt3o_prims.hpp"
#include "rt3o_curve.hpp"

#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace rt3o;

namespace {

thread_local std::string g_err;
bool g_chain_sum = false;  // see render_pixel
bool g_flatten = true;     // see rt3o_accel_build

struct Box {
    f3 lo{3e38f, 3e38f, 3e38f}, hi{-3e38f, -3e38f, -3e38f};
    void grow(f3 p) {
        lo = {fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z)};
        hi = {fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z)};
    }
    void grow(const Box& b) { grow(b.lo); grow(b.hi); }
    float area() const {
        f3 e = hi - lo;
        return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x);
    }
};

// ------------------------------------------------------------------ BVH2 (oracle-only accelerator)
struct Bvh2 {
    struct Node { Box box; int left, right, first, count, axis; };  // leaf when count > 0
    std::vector<Node> nodes;
    std::vector<int> order;

    void build(const std::vector<Box>& boxes) {
        nodes.clear();
        int n = (int)boxes.size();
        order.resize(n);
        for (int i = 0; i < n; i++) order[i] = i;
        if (n == 0) return;
        std::vector<f3> cent(n);
        for (int i = 0; i < n; i++) cent[i] = (boxes[i].lo + boxes[i].hi) * 0.5f;
        nodes.reserve(2 * n);
        nodes.push_back({});
        struct Item { int node, first, count; };
        std::vector<Item> todo{{0, 0, n}};
        while (!todo.empty()) {
            Item it = todo.back();
            todo.pop_back();
            Box b, cb;
            for (int i = it.first; i < it.first + it.count; i++) { b.grow(boxes[order[i]]); cb.grow(cent[order[i]]); }
            // pad: culling must stay conservative w.r.t. the brute-force primitive tests
            f3 pad = {1e-5f * (fabsf(b.lo.x) + fabsf(b.hi.x)) + 1e-7f, 1e-5f * (fabsf(b.lo.y) + fabsf(b.hi.y)) + 1e-7f,
                      1e-5f * (fabsf(b.lo.z) + fabsf(b.hi.z)) + 1e-7f};
            Node nd;
            nd.box.lo = b.lo - pad;
            nd.box.hi = b.hi + pad;
            nd.left = nd.right = -1;
            nd.first = it.first;
            nd.count = it.count;
            nd.axis = 0;
            if (it.count > 4) {
                f3 ext = cb.hi - cb.lo;
                int axis = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
                float lo = get(cb.lo, axis), e = get(ext, axis);
                int mid = -1;
                if (e > 0) {
                    const int NB = 16;
                    Box bb[NB];
                    int bc[NB] = {0};
                    float k = NB * (1.0f - 1e-6f) / e;
                    for (int i = it.first; i < it.first + it.count; i++) {
                        int bi = (int)((get(cent[order[i]], axis) - lo) * k);
                        bi = std::min(std::max(bi, 0), NB - 1);
                        bb[bi].grow(boxes[order[i]]);
                        bc[bi]++;
                    }
                    float la[NB], ra[NB];
                    int lc[NB], rc[NB];
                    Box acc;
                    int c = 0;
                    for (int i = 0; i < NB; i++) { if (bc[i]) acc.grow(bb[i]); c += bc[i]; la[i] = c ? acc.area() : 0; lc[i] = c; }
                    acc = Box();
                    c = 0;
                    for (int i = NB - 1; i >= 0; i--) { if (bc[i]) acc.grow(bb[i]); c += bc[i]; ra[i] = c ? acc.area() : 0; rc[i] = c; }
                    float best = 3e38f;
                    int bs = -1;
                    for (int i = 0; i < NB - 1; i++) {
                        if (lc[i] == 0 || rc[i + 1] == 0) continue;
                        float cost = la[i] * lc[i] + ra[i + 1] * rc[i + 1];
                        if (cost < best) { best = cost; bs = i; }
                    }
                    if (bs >= 0) {
                        auto pivot = std::partition(order.begin() + it.first, order.begin() + it.first + it.count, [&](int p) {
                            int bi = (int)((get(cent[p], axis) - lo) * k);
                            bi = std::min(std::max(bi, 0), NB - 1);
                            return bi <= bs;
                        });
                        mid = (int)(pivot - order.begin());
                    }
                }
                if (mid <= it.first || mid >= it.first + it.count) mid = it.first + it.count / 2;  // median fallback
                nd.count = 0;
                nd.axis = axis;
                nd.left = (int)nodes.size();
                nd.right = nd.left + 1;
                nodes.push_back({});
                nodes.push_back({});
                todo.push_back({nd.left, it.first, mid - it.first});
                todo.push_back({nd.right, mid, it.first + it.count - mid});
            }
            nodes[it.node] = nd;
        }
    }

    // calls leaf(prim) for every primitive whose padded box the ray may touch within [tmin, *tfar].  `slack` (3 absolute
    // distances) and `widen` loosen the test further for callers whose primitive test does not run on this ray (flattened
    // instances: the boxes are object space, the triangles are tested in world space)
    template <class F>
    bool traverse(f3 o, f3 d, float tmin, const float* tfar, F&& leaf, const float* slack = nullptr, float widen = 4e-7f) const {
        if (nodes.empty()) return false;
        f3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node& nd = nodes[stack[--sp]];
            float t0 = tmin, t1 = *tfar;
            bool ok = true;
            for (int a = 0; a < 3 && ok; a++) {
                float oa = get(o, a), da = get(d, a), lo = get(nd.box.lo, a), hi = get(nd.box.hi, a);
                if (slack) { lo -= slack[a]; hi += slack[a]; }
                if (da == 0.0f) { ok = (oa >= lo && oa <= hi); continue; }
                float ia = get(inv, a);
                float ta = (lo - oa) * ia, tb = (hi - oa) * ia;
                if (ta > tb) std::swap(ta, tb);
                // widen by a few ulps (robust slab test)
                ta -= fabsf(ta) * widen;
                tb += fabsf(tb) * widen;
                t0 = fmaxf(t0, ta);
                t1 = fminf(t1, tb);
                ok = t0 <= t1;
            }
            if (!ok) continue;
            if (nd.count > 0) {
                for (int i = nd.first; i < nd.first + nd.count; i++)
                    if (leaf(order[i])) return true;
            } else {
                if (get(d, nd.axis) > 0.0f) { stack[sp++] = nd.right; stack[sp++] = nd.left; }
                else { stack[sp++] = nd.left; stack[sp++] = nd.right; }
            }
        }
        return false;
    }
};

enum { PRIM_TRI = 0, PRIM_SPHERE = 1, PRIM_CURVE = 2 };

struct Blas {
    int type = PRIM_TRI;
    std::vector<f3> verts, normals;    // verts: [vkeys][nv] (vertex-key motion, cuda_mesh.h:85-88: keys spread over time [0,1])
    int vkeys = 1, nv = 0;
    std::vector<int> sub_first;        // spline curves: first linear sub-segment of every USER segment (+ end); empty otherwise
    std::vector<int> sub_seg;          // spline curves: user segment of every sub-segment (hits translated at the API boundary)
    bool spline() const { return !sub_first.empty(); }
    std::vector<CurvePoly> poly;       // spline curves: the true curve of every USER segment (normals are taken from it, cuda/curve.h:311-379)
    std::vector<f2> uvs;               // normals / uvs may be empty: the SDK's fallbacks (cuda/LocalGeometry.h:120-124,150-158)
    std::vector<float> colors;         // optional vertex colours, 4 per vertex (LocalGeometry.h:99-110)
    std::vector<int32_t> idx;          // tris: 3 per prim
    std::vector<float> cr;             // spheres: 4 per prim; curves: control points 4 per cp
    std::vector<int32_t> seg;          // curves: first cp per segment
    int nprims = 0;
    Bvh2 bvh;
    Box bounds;

    Box prim_box(int p) const {
        Box b;
        if (type == PRIM_TRI) {
            for (int k = 0; k < vkeys; k++)  // vertices move on straight segments between keys: the key boxes bound the motion
                for (int c = 0; c < 3; c++) b.grow(verts[(size_t)k * nv + idx[3 * p + c]]);
        } else if (type == PRIM_SPHERE) {
            f3 c = {cr[4 * p], cr[4 * p + 1], cr[4 * p + 2]};
            float r = cr[4 * p + 3];
            b.grow(c - mk3(r, r, r)); b.grow(c + mk3(r, r, r));
        } else {
            int a = seg[p];
            for (int k = 0; k < 2; k++) {
                f3 c = {cr[4 * (a + k)], cr[4 * (a + k) + 1], cr[4 * (a + k) + 2]};
                float r = cr[4 * (a + k) + 3];
                b.grow(c - mk3(r, r, r)); b.grow(c + mk3(r, r, r));
            }
        }
        return b;
    }
};

struct Instance {
    int blas;
    Affine stat, stat_inv;
    int nkeys = 0;
    bool identity = false;
    bool flat = false;      // flattened (see rt3o_accel_build): its triangles are tested in WORLD space, vertices = stat * v
    float flat_w[3] = {0, 0, 0};  // magnitude bound of the instance's world-space vertices per axis (padding of the object-space culling)
    std::vector<float> keys;
    float t0 = 0, t1 = 1;
    f3 emission{0, 0, 0}, diffuse{0.8f, 0.8f, 0.8f};
    int tex = -1;
    bool has_xf = false;  // texcoord transform of the SDK's sampleTexture (cuda/LocalShading.h:37-54)
    f2 tex_scale{1, 1}, tex_rot{0, 1}, tex_off{0, 0};
};

struct Texture { int w, h, addr, filt; std::vector<uint8_t> px; };

struct Hit { float t = 0, u = 0, v = 0; int prim = -1, inst = -1; };

static inline bool better(float t, int inst, int prim, const Hit& h) {
    if (h.prim < 0) return true;
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (inst != h.inst) return inst < h.inst;
    return prim < h.prim;
}

}  // namespace

struct rt3o_scene {
    std::vector<std::unique_ptr<Blas>> blas;
    std::vector<Instance> inst;
    std::vector<Texture> tex;
    std::vector<Light> lights;
    std::vector<float> light_cdf;  // running sum of luminance(emission) * area (mode 2, power light sampler)
    Bvh2 tlas;
    bool built = false;
    uint32_t w = 0, h = 0;
    std::vector<float> accum;
    std::vector<uint8_t> frame;
    std::atomic<uint64_t> n_primary{0}, n_bounce{0}, n_shadow{0}, n_samples{0};

    void to_object(const Instance& in, float time, f3 o, f3 d, f3& oo, f3& od) const {
        // an instance whose transform is bit-for-bit the identity (and has no motion keys) — every
        // instance the reference creates, cuda_scene.h:141-146 — sees the world-space ray unchanged
        if (in.identity) { oo = o; od = d; return; }
        oo = xform_point(in.stat_inv, o);
        od = xform_vector(in.stat_inv, d);
        if (in.nkeys > 0) {
            Affine m = lerp_keys(in.keys.data(), in.nkeys, in.t0, in.t1, time);
            Affine mi = invert_affine(m);
            oo = xform_point(mi, oo);
            od = xform_vector(mi, od);
        }
    }

    // test one primitive of one BLAS in object space
    static inline bool test_prim(const Blas& b, int p, f3 oo, f3 od, const RayShear& sh, float tmin, float tmax, float time,
                                 float& t, float& u, float& v) {
        if (b.type == PRIM_TRI) {
            if (b.vkeys == 1)
                return hit_triangle(oo, sh, b.verts[b.idx[3 * p]], b.verts[b.idx[3 * p + 1]], b.verts[b.idx[3 * p + 2]], tmin, tmax, t, u, v);
            // vertex-key motion: per-vertex linear interpolation of the bracketing keys at the ray time
            // (OptiX motion GAS semantics, timeBegin 0 / timeEnd 1, clamped; cuda_mesh.h:82-88)
            const float tc = fminf(fmaxf(time, 0.0f), 1.0f);
            const float f = tc * (float)(b.vkeys - 1);
            int ki = (int)floorf(f);
            if (ki > b.vkeys - 2) ki = b.vkeys - 2;
            const float a = f - (float)ki, w = 1.0f - a;
            f3 q[3];
            for (int c = 0; c < 3; c++) {
                const f3 v0 = b.verts[(size_t)ki * b.nv + b.idx[3 * p + c]], v1 = b.verts[(size_t)(ki + 1) * b.nv + b.idx[3 * p + c]];
                q[c] = {w * v0.x + a * v1.x, w * v0.y + a * v1.y, w * v0.z + a * v1.z};
            }
            return hit_triangle(oo, sh, q[0], q[1], q[2], tmin, tmax, t, u, v);
        } else if (b.type == PRIM_SPHERE) {
            u = v = 0;
            return hit_sphere(oo, od, {b.cr[4 * p], b.cr[4 * p + 1], b.cr[4 * p + 2]}, b.cr[4 * p + 3], tmin, tmax, t);
        } else {
            int a = b.seg[p];
            v = 0;
            return hit_curve_linear(oo, od, {b.cr[4 * a], b.cr[4 * a + 1], b.cr[4 * a + 2]}, b.cr[4 * a + 3],
                                    {b.cr[4 * a + 4], b.cr[4 * a + 5], b.cr[4 * a + 6]}, b.cr[4 * a + 7], tmin, tmax, t, u);
        }
    }

    // closest hit (any_hit=false) or occlusion (any_hit=true)
    Hit trace(f3 o, f3 d, float tmin, float tmax, float time, bool any_hit, int accel) const {
        Hit best;
        float tfar = tmax;
        auto visit_instance = [&](int ii) -> bool {
            const Instance& in = inst[ii];
            const Blas& b = *blas[in.blas];
            f3 oo, od;
            to_object(in, time, o, d, oo, od);
            RayShear sh = make_shear(in.flat ? d : od);
            auto visit_prim = [&](int p) -> bool {
                float t, u, v;
                if (in.flat) {  // the world-space ray against the world-space triangle, exactly what the flattened BLAS of the product holds
                    if (!hit_triangle(o, sh, xform_point(in.stat, b.verts[b.idx[3 * p]]), xform_point(in.stat, b.verts[b.idx[3 * p + 1]]),
                                      xform_point(in.stat, b.verts[b.idx[3 * p + 2]]), tmin, tmax, t, u, v)) return false;
                } else if (!test_prim(b, p, oo, od, sh, tmin, tmax, time, t, u, v)) return false;
                if (better(t, ii, p, best)) {
                    best.t = t; best.u = u; best.v = v; best.prim = p; best.inst = ii;
                    tfar = t;
                }
                return any_hit;
            };
            if (accel == 0) {
                for (int p = 0; p < b.nprims; p++)
                    if (visit_prim(p)) return true;
                return false;
            }
            if (in.flat) {
                // The BVH is the object-space one, walked by the object-space ray, but the hit decision is taken in world space:
                // the culling must allow for the rounding of the ray transform and of the transformed vertices.  Both are a
                // few 1e-7 of sum_j |minv_ij| (|world coordinate_j|) + |minv_i3| in object space; 1e-4 of it is taken.
                float slack[3];
                for (int i = 0; i < 3; i++) {
                    const float* m = in.stat_inv.m + 4 * i;
                    slack[i] = 1e-4f * (fabsf(m[0]) * (in.flat_w[0] + fabsf(o.x)) + fabsf(m[1]) * (in.flat_w[1] + fabsf(o.y)) + fabsf(m[2]) * (in.flat_w[2] + fabsf(o.z)) + fabsf(m[3]));
                }
                return b.bvh.traverse(oo, od, tmin, &tfar, visit_prim, slack, 1e-4f);
            }
            return b.bvh.traverse(oo, od, tmin, &tfar, visit_prim);
        };
        if (accel == 0) {
            for (int ii = 0; ii < (int)inst.size(); ii++)
                if (visit_instance(ii)) break;
        } else {
            tlas.traverse(o, d, tmin, &tfar, visit_instance);
        }
        return best;
    }

    Box instance_world_box(const Instance& in) const {
        const Blas& b = *blas[in.blas];
        Box wb;
        int nk = in.nkeys > 0 ? in.nkeys : 1;
        for (int k = 0; k < nk; k++) {
            for (int c = 0; c < 8; c++) {
                f3 p = {(c & 1) ? b.bounds.hi.x : b.bounds.lo.x, (c & 2) ? b.bounds.hi.y : b.bounds.lo.y,
                        (c & 4) ? b.bounds.hi.z : b.bounds.lo.z};
                if (in.nkeys > 0) {
                    Affine m;
                    std::memcpy(m.m, in.keys.data() + 12 * k, sizeof(m.m));
                    p = xform_point(m, p);
                }
                wb.grow(xform_point(in.stat, p));
            }
        }
        return wb;
    }

    // ------------------------------------------------------------------ shading helpers
    // Texel index of an integer texel coordinate under an address mode (CUDA programming guide, texture fetching):
    // wrap = modulo, clamp = nearest edge texel, mirror = reflected every N texels, border = -1 (reads as 0).
    static int resolve_texel(int i, int n, int mode) {
        if (mode == RT3_ADDRESS_WRAP) { int m = i % n; return m < 0 ? m + n : m; }
        if (mode == RT3_ADDRESS_CLAMP) return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
        if (mode == RT3_ADDRESS_MIRROR) { int m = i % (2 * n); if (m < 0) m += 2 * n; return m < n ? m : 2 * n - 1 - m; }
        return (i >= 0 && i < n) ? i : -1;
    }
    f3 fetch_texture(int id, float u, float v) const {
        // normalised coords, RGBA8 -> [0,1], no sRGB decode (Q10).  filter 0 = point sampling (what the reference's
        // `FilterMode::Linear = 0` really selects, Q9); filter 1 = the hardware's bilinear filter (what its
        // `FilterMode::Point = 1` selects): texel centres at i + 0.5, weights rounded to 8 fractional bits.
        const Texture& tx = tex[id];
        auto texel = [&](int x, int y) -> f3 {
            if (x < 0 || y < 0) return {0.0f, 0.0f, 0.0f};  // border
            const uint8_t* p = &tx.px[4 * ((size_t)y * tx.w + x)];
            return {(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f};
        };
        if (tx.filt == 0 && (tx.addr == RT3_ADDRESS_WRAP || tx.addr == RT3_ADDRESS_CLAMP)) {
            auto addr = [&](float c, int n) -> int {
                if (tx.addr == RT3_ADDRESS_WRAP) {
                    float f = c - floorf(c);
                    int i = (int)(f * (float)n);
                    return i > n - 1 ? n - 1 : i;
                }
                float f = fminf(fmaxf(c, 0.0f), 1.0f);
                int i = (int)(f * (float)n);
                return i > n - 1 ? n - 1 : i;
            };
            return texel(addr(u, tx.w), addr(v, tx.h));
        }
        const float fx = u * (float)tx.w, fy = v * (float)tx.h;
        if (tx.filt == 0) return texel(resolve_texel((int)floorf(fx), tx.w, tx.addr), resolve_texel((int)floorf(fy), tx.h, tx.addr));
        const float bx = fx - 0.5f, by = fy - 0.5f;
        const float ix = floorf(bx), iy = floorf(by);
        const float al = floorf((bx - ix) * 256.0f + 0.5f) / 256.0f, be = floorf((by - iy) * 256.0f + 0.5f) / 256.0f;
        const int x0 = resolve_texel((int)ix, tx.w, tx.addr), x1 = resolve_texel((int)ix + 1, tx.w, tx.addr);
        const int y0 = resolve_texel((int)iy, tx.h, tx.addr), y1 = resolve_texel((int)iy + 1, tx.h, tx.addr);
        const f3 t00 = texel(x0, y0), t10 = texel(x1, y0), t01 = texel(x0, y1), t11 = texel(x1, y1);
        const float w00 = (1.0f - al) * (1.0f - be), w10 = al * (1.0f - be), w01 = (1.0f - al) * be, w11 = al * be;
        return {((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x, ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y,
                ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z};
    }

    // sampleTexture (cuda/LocalShading.h:37-54): the instance's texcoord transform, then tex2D
    f3 sample_texture(const Instance& in, f2 uv) const {
        if (!in.has_xf) return fetch_texture(in.tex, uv.x, uv.y);
        const float sx = uv.x * in.tex_scale.x, sy = uv.y * in.tex_scale.y;
        const float tu = (sx * in.tex_rot.y + sy * in.tex_rot.x) + in.tex_off.x;
        const float tv = (sx * (-in.tex_rot.x) + sy * in.tex_rot.y) + in.tex_off.y;
        return fetch_texture(in.tex, tu, tv);
    }

    // "LocalGeometry" of the new shade stage: object-space N/uv per closehit_radiance.cu:66-74,
    // N moved to world space by the inverse-transpose (cuda/LocalGeometry.h:110,119) — exact
    // identity for the reference's identity instances.
    // the three object-space vertices of a triangle at a ray time (vertex keys spread evenly over [0, 1], cuda_mesh.h:82-88)
    static void tri_verts(const Blas& b, int prim, float time, f3& P0, f3& P1, f3& P2) {
        const int i0 = b.idx[3 * prim], i1 = b.idx[3 * prim + 1], i2 = b.idx[3 * prim + 2];
        if (b.vkeys <= 1) { P0 = b.verts[i0]; P1 = b.verts[i1]; P2 = b.verts[i2]; return; }
        const float tc = fminf(fmaxf(time, 0.0f), 1.0f);
        const float f = tc * (float)(b.vkeys - 1);
        int ki = (int)floorf(f);
        if (ki > b.vkeys - 2) ki = b.vkeys - 2;
        const float al = f - (float)ki, w = 1.0f - al;
        const f3* k0 = &b.verts[(size_t)ki * b.nv];
        const f3* k1 = k0 + b.nv;
        P0 = {w * k0[i0].x + al * k1[i0].x, w * k0[i0].y + al * k1[i0].y, w * k0[i0].z + al * k1[i0].z};
        P1 = {w * k0[i1].x + al * k1[i1].x, w * k0[i1].y + al * k1[i1].y, w * k0[i1].z + al * k1[i1].z};
        P2 = {w * k0[i2].x + al * k1[i2].x, w * k0[i2].y + al * k1[i2].y, w * k0[i2].z + al * k1[i2].z};
    }

    void local_geometry(const Hit& h, f3 o, f3 d, float time, f3& N, f2& uv) const {
        const Instance& in = inst[h.inst];
        const Blas& b = *blas[in.blas];
        f3 n_obj;
        if (b.type == PRIM_TRI) {
            const int i0 = b.idx[3 * h.prim], i1 = b.idx[3 * h.prim + 1], i2 = b.idx[3 * h.prim + 2];
            const float w0 = 1.0f - h.u - h.v;
            if (!b.normals.empty()) n_obj = w0 * b.normals[i0] + h.u * b.normals[i1] + h.v * b.normals[i2];
            else {  // no vertex normals: the geometric normal (cuda/LocalGeometry.h:120-124)
                f3 P0, P1, P2;
                tri_verts(b, h.prim, time, P0, P1, P2);
                n_obj = cross(P1 - P0, P2 - P0);
            }
            if (!b.uvs.empty()) {
                uv.x = w0 * b.uvs[i0].x + h.u * b.uvs[i1].x + h.v * b.uvs[i2].x;
                uv.y = w0 * b.uvs[i0].y + h.u * b.uvs[i1].y + h.v * b.uvs[i2].y;
            } else uv = {h.u, h.v};  // no texcoords: the barycentrics (LocalGeometry.h:150-152)
        } else {
            f3 oo, od;
            to_object(in, time, o, d, oo, od);
            f3 ps = oo + h.t * od;
            if (b.type == PRIM_SPHERE) {
                f3 c = {b.cr[4 * h.prim], b.cr[4 * h.prim + 1], b.cr[4 * h.prim + 2]};
                n_obj = (ps - c) / b.cr[4 * h.prim + 3];
                uv = {0, 0};
            } else if (b.spline()) {  // spline curve: the SDK's bona fide normal of the TRUE curve (cuda/curve.h:311-379) at the hit's curve parameter
                const int sg = b.sub_seg[(size_t)h.prim], first = b.sub_first[(size_t)sg], K = b.sub_first[(size_t)sg + 1] - first;
                const float uu = ((float)(h.prim - first) + h.u) / (float)K;
                n_obj = curve_surface_normal_raw(b.poly[(size_t)sg], uu, ps);
                uv = {uu, 0};
            } else {  // cuda/curve.h:382-425 surfaceNormal<LinearInterpolator>
                int a = b.seg[h.prim];
                f3 p0 = {b.cr[4 * a], b.cr[4 * a + 1], b.cr[4 * a + 2]};
                f3 p1 = {b.cr[4 * a + 4], b.cr[4 * a + 5], b.cr[4 * a + 6]};
                float r0 = b.cr[4 * a + 3], r1 = b.cr[4 * a + 7];
                if (h.u == 0.0f) n_obj = ps - p0;
                else if (h.u >= 1.0f) n_obj = ps - ((p1 - p0) + p0);  // the SDK rebuilds the end point from its pre-transformed coefficients (curve.h:389-394)
                else {
                    f3 dd3 = p1 - p0;
                    float dr = r1 - r0;
                    f3 p = p0 + h.u * dd3;
                    float r = r0 + h.u * dr;
                    float dd = dot(dd3, dd3);
                    f3 o1 = ps - p;
                    o1 = o1 - (dot(o1, dd3) / dd) * dd3;
                    o1 = o1 * (r / length(o1));
                    n_obj = dd * o1 - (dr * r) * dd3;
                }
                uv = {h.u, 0};
            }
        }
        // object -> world normal: (M^-1)^T n, M = static o motion(t)
        f3 n = n_obj;
        if (in.nkeys > 0) {
            Affine m = lerp_keys(in.keys.data(), in.nkeys, in.t0, in.t1, time);
            n = xform_normal_by_inverse(invert_affine(m), n);
        }
        n = xform_normal_by_inverse(in.stat_inv, n);
        N = normalize(n);
    }


    // The SDK's complete stage record (cuda/LocalGeometry.h:40-175, one texcoord set): world P from the interpolated
    // object-space vertices, Ng, N, UV and the object-space derivatives exactly as getLocalGeometry forms them
    // (LocalGeometry.h:126-160).  Spheres / curves (empty in the SDK): P = o + t d, N = Ng = surface normal, zero derivatives.
    void local_geometry_full(const Hit& h, f3 o, f3 d, float time, float out[27]) const {
        for (int k = 0; k < 27; k++) out[k] = 0.0f;
        const Instance& in = inst[h.inst];
        const Blas& b = *blas[in.blas];
        f3 P, N, Ng, dndu{0, 0, 0}, dndv{0, 0, 0}, dpdu{0, 0, 0}, dpdv{0, 0, 0};
        f2 uv{0, 0};
        float color[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if (b.type == PRIM_TRI) {
            const bool moving = in.nkeys > 0;
            Affine m{}, mi{};
            if (moving) { m = lerp_keys(in.keys.data(), in.nkeys, in.t0, in.t1, time); mi = invert_affine(m); }
            const int i0 = b.idx[3 * h.prim], i1 = b.idx[3 * h.prim + 1], i2 = b.idx[3 * h.prim + 2];
            f3 P0, P1, P2;
            tri_verts(b, h.prim, time, P0, P1, P2);
            const float w0 = 1.0f - h.u - h.v;
            f3 p = w0 * P0 + h.u * P1 + h.v * P2;
            if (moving) p = xform_point(m, p);
            P = xform_point(in.stat, p);
            if (!b.colors.empty())   // LocalGeometry.h:99-106
                for (int k = 0; k < 4; k++) color[k] = w0 * b.colors[4 * i0 + k] + h.u * b.colors[4 * i1 + k] + h.v * b.colors[4 * i2 + k];
            f3 ng = cross(P1 - P0, P2 - P0);
            if (moving) ng = xform_normal_by_inverse(mi, ng);
            Ng = normalize(xform_normal_by_inverse(in.stat_inv, ng));
            f3 N0, N1, N2;
            if (!b.normals.empty()) {
                N0 = b.normals[i0]; N1 = b.normals[i1]; N2 = b.normals[i2];
                f3 n = w0 * N0 + h.u * N1 + h.v * N2;
                if (moving) n = xform_normal_by_inverse(mi, n);
                N = normalize(xform_normal_by_inverse(in.stat_inv, n));
            } else N = N0 = N1 = N2 = Ng;   // LocalGeometry.h:120-124: the (world-space, unit) geometric normal stands in for all three
            const f3 dp1 = P0 - P2, dp2 = P1 - P2, dn1 = N0 - N2, dn2 = N1 - N2;
            if (!b.uvs.empty()) {
                const f2 U0 = b.uvs[i0], U1 = b.uvs[i1], U2 = b.uvs[i2];
                uv.x = w0 * U0.x + h.u * U1.x + h.v * U2.x;
                uv.y = w0 * U0.y + h.u * U1.y + h.v * U2.y;
                const float du1 = U0.x - U2.x, du2 = U1.x - U2.x, dv1 = U0.y - U2.y, dv2 = U1.y - U2.y;
                const float det = du1 * dv2 - dv1 * du2;
                const float invdet = 1.0f / det;
                dpdu = (dv2 * dp1 - dv1 * dp2) * invdet;
                dpdv = ((-du2) * dp1 + du1 * dp2) * invdet;
                dndu = (dv2 * dn1 - dv1 * dn2) * invdet;
                dndv = ((-du2) * dn1 + du1 * dn2) * invdet;
            } else {   // LocalGeometry.h:150-158
                uv = {h.u, h.v};
                dpdu = -dp1;
                dpdv = -dp1 + dp2;
                dndu = -dn1;
                dndv = -dn1 + dn2;
            }
        } else {
            local_geometry(h, o, d, time, N, uv);
            Ng = N;
            P = o + h.t * d;
        }
        const float v[27] = {P.x, P.y, P.z, N.x, N.y, N.z, Ng.x, Ng.y, Ng.z, uv.x, uv.y, dndu.x, dndu.y, dndu.z, dndv.x, dndv.y, dndv.z,
                             dpdu.x, dpdu.y, dpdu.z, dpdv.x, dpdv.y, dpdv.z, color[0], color[1], color[2], color[3]};
        for (int k = 0; k < 27; k++) out[k] = v[k];
    }

    // ---------------------------------------------------------------------------------- CORRECTED mode (mode = 1)
    // the same stages with the estimator errors Q2-Q5, Q7, Q8 fixed — unbiased
    // Lambertian path tracing with uniform-light NEE and two-strategy MIS (power heuristic):
    //   * per-sample independent stream: seed = tea<4>(pixel, subframe * spl + k)            (Q8)
    //   * one ray time per path, drawn after the jitter                                        (Q7)
    //   * throughput *= albedo (f * cos / pdf of the cosine-sampled Lambert lobe)              (Q2)
    //   * NEE: Le * (albedo/pi) * cos_s * w / pdf_light, pdf_light = dist^2 / (N * area * cos_l),
    //     weighted by the throughput BEFORE this bounce's BSDF factor                          (Q3, Q4)
    //   * BSDF-sampled emitter hits count at every depth with the complementary MIS weight     (Q3)
    //   * Russian roulette with p = min(lum(throughput), 1)                                    (Q5)
    // Emissive meshes must be identity instances (the light list holds object-space vertices, Q15).
    // mode 2 (the reference README's unchecked "power light sampler"): light k is chosen with
    // probability power_k / total, power = luminance(emission) * area; cdf = sequential fp32 running sum
    uint32_t pick_light_by_power(float xi01, float& p_sel) const {
        const uint32_t nl = (uint32_t)lights.size();
        const float total = light_cdf[nl - 1];
        const float xi = xi01 * total;
        uint32_t lo = 0, hi = nl - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (xi < light_cdf[mid]) hi = mid; else lo = mid + 1;
        }
        p_sel = light_power(lights[lo].emission, lights[lo].area) / total;
        return lo;
    }

    f3 render_pixel_corrected(const rt3_render_settings& rs, uint32_t x, uint32_t y, int accel, uint64_t cnt[3]) const {
        const bool by_power = rs.mode == 2 && light_cdf.back() > 0.0f;
        const uint32_t w = rs.width, h = rs.height;
        const f3 eye = {rs.eye[0], rs.eye[1], rs.eye[2]}, U = {rs.U[0], rs.U[1], rs.U[2]}, V = {rs.V[0], rs.V[1], rs.V[2]},
                 W = {rs.W[0], rs.W[1], rs.W[2]};
        const f3 miss = {rs.miss_color[0], rs.miss_color[1], rs.miss_color[2]};
        const int max_depth = rs.max_depth > 0 ? rs.max_depth : (1 << 30);
        const uint32_t nl = (uint32_t)lights.size();
        const float inv_pi = (float)(1.0 / 3.14159265358979323846);
        f3 result = {0, 0, 0};
        for (uint32_t k = 0; k < rs.samples_per_launch; k++) {
            uint32_t seed = tea4(y * w + x, rs.subframe_index * rs.samples_per_launch + k);
            const float jx = rnd(seed);
            const float jy = rnd(seed);
            const float time = rnd(seed);
            const float dx = 2.0f * (((float)x + jx) / (float)w) - 1.0f;
            const float dy = 2.0f * (((float)y + jy) / (float)h) - 1.0f;
            f3 dir = normalize(dx * U + dy * V + W);
            f3 org = eye;
            f3 beta = {1, 1, 1};
            f3 L = {0, 0, 0};
            float pdf_prev = 0.0f;
            int depth = 0;
            for (;;) {
                cnt[depth == 0 ? 0 : 1]++;
                Hit hit = trace(org, dir, 0.01f, 1e16f, time, false, accel);
                if (hit.prim < 0) { L = L + beta * miss; break; }
                const Instance& in = inst[hit.inst];
                f3 N;
                f2 uv;
                local_geometry(hit, org, dir, time, N, uv);
                const f3 Ns = faceforward(N, -dir, N);
                const f3 P = org + hit.t * dir;
                if (in.emission.x != 0.0f || in.emission.y != 0.0f || in.emission.z != 0.0f) {
                    float wgt = 1.0f;
                    // BSDF-sampled emitter hit: weight against the NEE strategy where NEE can produce this point at all: the light
                    // list holds object-space key-0 triangles (Q15), so only static meshes under identity instances qualify
                    const Blas& b = *blas[in.blas];
                    if (depth > 0 && b.type == PRIM_TRI && b.vkeys == 1 && in.identity) {
                        const f3 v0 = b.verts[b.idx[3 * hit.prim]], v1 = b.verts[b.idx[3 * hit.prim + 1]], v2 = b.verts[b.idx[3 * hit.prim + 2]];
                        const f3 nrm = cross(v1 - v0, v2 - v0);
                        const float area = 0.5f * length(nrm);
                        const float cos_l = fabsf(dot(normalize(nrm), dir));
                        const float dist2 = hit.t * hit.t * dot(dir, dir);
                        const float pdf_light = by_power ? (dist2 / (area * cos_l)) * (light_power(in.emission, area) / light_cdf.back())
                                                         : dist2 / ((float)nl * area * cos_l);
                        wgt = power_heuristic(pdf_prev, pdf_light);
                    }
                    L = L + beta * in.emission * wgt;
                }
                const f3 albedo = in.tex >= 0 ? sample_texture(in, uv) : in.diffuse;
                // next event estimation
                float p_sel = 0.0f;
                const float xi_l = rnd(seed);
                const Light& lt = lights[by_power ? pick_light_by_power(xi_l, p_sel) : (uint32_t)(int)(xi_l * (float)nl)];
                const float u = rnd(seed);
                const float v = rnd(seed);
                const float su0 = sqrtf(u);
                const float b0 = 1.0f - su0, b1 = v * su0;
                const f3 lpos = b0 * lt.v0 + b1 * lt.v1 + (1.0f - b0 - b1) * lt.v2;
                const f3 dl = lpos - P;
                const float dist2 = dot(dl, dl);
                if (dist2 > 1e-10f) {
                    const float dist = sqrtf(dist2);
                    const f3 Ld = dl / dist;
                    const float cos_s = dot(Ns, Ld);
                    const float cos_l = fabsf(dot(Ld, lt.normal));
                    if (cos_s > 0.0f && cos_l > 0.0f && lt.area > 0.0f) {
                        const float pdf_light = by_power ? (dist2 / (lt.area * cos_l)) * p_sel : dist2 / ((float)nl * lt.area * cos_l);
                        const float pdf_bsdf = cos_s * inv_pi;
                        const float wgt = power_heuristic(pdf_light, pdf_bsdf);
                        const f3 contrib = beta * albedo * lt.emission * (inv_pi * cos_s * wgt / pdf_light);
                        cnt[2]++;
                        Hit sh = trace(P, Ld, 0.001f, dist - 0.01f, time, true, accel);
                        if (sh.prim < 0) L = L + contrib;
                    }
                }
                // BSDF sample
                const float u1 = rnd(seed);
                const float u2 = rnd(seed);
                const f3 w_in = sample_cosine_hemisphere(u1, u2);
                if (!(w_in.z > 0.0f)) break;
                pdf_prev = w_in.z * inv_pi;
                Onb onb(Ns);
                dir = onb.inverse_transform(w_in);
                org = P;
                beta = beta * albedo;
                const float p = fminf(beta.x * 0.30f + beta.y * 0.59f + beta.z * 0.11f, 1.0f);
                if (rnd(seed) > p) break;
                beta = beta / p;
                ++depth;
                if (depth >= max_depth) break;
            }
            result = result + L;
        }
        return result / (float)rs.samples_per_launch;
    }

    // one pixel, one launch: returns result/spl (raygen.cu:28-76)
    f3 render_pixel(const rt3_render_settings& rs, uint32_t x, uint32_t y, int accel, uint64_t cnt[3]) const {
        const uint32_t w = rs.width, h = rs.height;
        const f3 eye = {rs.eye[0], rs.eye[1], rs.eye[2]}, U = {rs.U[0], rs.U[1], rs.U[2]}, V = {rs.V[0], rs.V[1], rs.V[2]},
                 W = {rs.W[0], rs.W[1], rs.W[2]};
        const f3 miss = {rs.miss_color[0], rs.miss_color[1], rs.miss_color[2]};
        uint32_t seed = tea4(y * w + x, rs.subframe_index);
        f3 result = {0, 0, 0};
        const int max_depth = rs.max_depth > 0 ? rs.max_depth : (1 << 30);
        const uint32_t nl = (uint32_t)lights.size();
        int i = (int)rs.samples_per_launch;
        do {
            const float jx = rnd(seed);
            const float jy = rnd(seed);
            const float dx = 2.0f * (((float)x + jx) / (float)w) - 1.0f;
            const float dy = 2.0f * (((float)y + jy) / (float)h) - 1.0f;
            f3 dir = normalize(dx * U + dy * V + W);
            f3 org = eye;
            f3 att = {1, 1, 1};
            uint32_t pseed = seed;
            int depth = 0;
            f3 last_att = att;
            // Association of the radiance sum: the reference keeps ONE running `result` across the
            // samples of a launch (raygen.cu:27,58-59).  The wavefront kernels sum each path on its
            // own and then the samples of a pixel in order; the oracle follows that association by
            // default (fp32 reassociation only, quantified in tests/test_oracle.py) and the
            // reference's single chain when g_chain_sum is set.
            f3 sres = {0, 0, 0};
            f3& acc = g_chain_sum ? result : sres;
            for (;;) {
                // traceRadiance (shader_common.h:50-106)
                const float time = rnd(pseed);
                cnt[depth == 0 ? 0 : 1]++;
                Hit hit = trace(org, dir, 0.01f, 1e16f, time, false, accel);
                f3 emitted, radiance, norg = org, ndir = dir;
                bool done;
                if (hit.prim < 0) {  // miss.cu:29-32
                    radiance = miss;
                    emitted = {0, 0, 0};
                    done = true;
                } else {  // closehit_radiance.cu:60-160
                    const Instance& in = inst[hit.inst];
                    f3 N;
                    f2 uv;
                    local_geometry(hit, org, dir, time, N, uv);
                    const f3 Ns = faceforward(N, -dir, N);
                    const f3 P = org + hit.t * dir;
                    emitted = depth == 0 ? in.emission : mk3(0, 0, 0);
                    uint32_t s = pseed;
                    (void)rnd(s);
                    (void)rnd(s);  // z1, z2 discarded (Q6)
                    const float u1 = rnd(s);
                    const float u2 = rnd(s);
                    f3 w_in = sample_cosine_hemisphere(u1, u2);
                    const float pdf_prev = (float)((double)w_in.z / 3.14159265358979323846);
                    Onb onb(Ns);
                    ndir = onb.inverse_transform(w_in);
                    norg = P;
                    const float bsdf = (float)(1.0 / 3.14159265358979323846);
                    const f3 albedo = in.tex >= 0 ? sample_texture(in, uv) : in.diffuse;
                    att = att * albedo;
                    att = att * (bsdf / pdf_prev);
                    // NEE
                    const Light& lt = lights[(int)(rnd(s) * (float)nl)];
                    f3 lpos, lem;
                    float pdf_light;
                    light_sample(lt, P, s, lpos, lem, pdf_light);
                    pdf_light = pdf_light / (float)nl;
                    pseed = s;
                    const float Ldist = length(lpos - P);
                    const f3 L = normalize(lpos - P);
                    const float nDl = dot(Ns, L);
                    f3 weight = {0, 0, 0};
                    if (nDl > 0.0f) {
                        const float tshadow = rnd(s);  // local copy only (Q7)
                        cnt[2]++;
                        Hit sh = trace(P, L, 0.001f, Ldist - 0.01f, tshadow, true, accel);
                        if (sh.prim < 0) {
                            const float pdf_scat = (float)((double)fabsf(dot(L, Ns)) / 3.14159265358979323846);
                            weight = albedo * (power_heuristic(pdf_light, pdf_scat) * bsdf);
                        }
                    }
                    radiance = lem * weight;
                    done = false;
                }
                acc = acc + emitted;
                acc = acc + radiance * last_att;
                last_att = att;
                const float p = att.x * 0.30f + att.y * 0.59f + att.z * 0.11f;
                if (done || rnd(pseed) > p) break;
                att = att / p;
                org = norg;
                dir = ndir;
                ++depth;
                if (depth >= max_depth) break;  // extension
            }
            if (!g_chain_sum) result = result + sres;
        } while (--i);
        return result / (float)rs.samples_per_launch;
    }
};

template <class F>
static void parallel_for(int n, int nthreads, int grain, F&& f) {
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next{0};
    auto worker = [&]() {
        for (;;) {
            int b = next.fetch_add(grain);
            if (b >= n) break;
            int e = std::min(n, b + grain);
            for (int i = b; i < e; i++) f(i);
        }
    };
    if (nthreads == 1) { worker(); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

// ====================================================================== C API
#define RT3O_TRY try {
#define RT3O_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

rt3o_scene* rt3o_scene_create(void) { return new rt3o_scene(); }
void rt3o_scene_destroy(rt3o_scene* s) { delete s; }
const char* rt3o_last_error(void) { return g_err.c_str(); }

static int finish_blas(rt3o_scene* s, std::unique_ptr<Blas> b) {
    std::vector<Box> boxes(b->nprims);
    for (int p = 0; p < b->nprims; p++) { boxes[p] = b->prim_box(p); b->bounds.grow(boxes[p]); }
    b->bvh.build(boxes);
    s->blas.push_back(std::move(b));
    return (int)s->blas.size() - 1;
}

int rt3o_mesh_create(rt3o_scene* s, const float* verts, int num_keys, int nv, const int32_t* idx, int nt,
                     const float* normals, const float* uvs) {
    RT3O_TRY
    if (!s || !verts || !idx || nv <= 0 || nt <= 0 || num_keys < 1) { g_err = "mesh_create: bad argument"; return -1; }
    for (int i = 0; i < 3 * nt; i++)
        if (idx[i] < 0 || idx[i] >= nv) { g_err = "mesh_create: index out of range"; return -1; }
    auto b = std::make_unique<Blas>();
    b->type = PRIM_TRI;
    b->nprims = nt;
    b->vkeys = num_keys; b->nv = nv;
    b->verts.resize((size_t)nv * num_keys);
    std::memcpy(b->verts.data(), verts, sizeof(f3) * (size_t)nv * num_keys);  // [key][vertex]; normals / uvs: key 0 (create_sbt binds the buffer start)
    if (normals) { b->normals.resize(nv); std::memcpy(b->normals.data(), normals, sizeof(f3) * nv); }
    if (uvs) { b->uvs.resize(nv); std::memcpy(b->uvs.data(), uvs, sizeof(f2) * nv); }
    b->idx.assign(idx, idx + 3 * nt);
    return finish_blas(s, std::move(b));
    RT3O_CATCH(-1)
}
int rt3o_mesh_set_colors(rt3o_scene* s, int blas, const float* rgba) {
    RT3O_TRY
    if (!s || !rgba || blas < 0 || blas >= (int)s->blas.size() || s->blas[blas]->type != PRIM_TRI) { g_err = "mesh_set_colors: bad argument"; return -1; }
    s->blas[blas]->colors.assign(rgba, rgba + 4 * (size_t)s->blas[blas]->nv);
    return 0;
    RT3O_CATCH(-1)
}
int rt3o_spheres_create(rt3o_scene* s, const float* cr, int n) {
    RT3O_TRY
    if (!s || !cr || n <= 0) { g_err = "spheres_create: bad argument"; return -1; }
    auto b = std::make_unique<Blas>();
    b->type = PRIM_SPHERE;
    b->nprims = n;
    b->cr.assign(cr, cr + 4 * n);
    return finish_blas(s, std::move(b));
    RT3O_CATCH(-1)
}
// Spline curves (quadratic / cubic B-spline, Catmull-Rom, Bezier: the SDK's interpolators, cuda/curve.h:98-243): every user
// segment is intersected as K round linear sub-segments between the points P(k / K) of its TRUE polynomial; normals come
// from the true curve (local_geometry).  K is chosen per segment from the polynomial itself: a chord over 1 / K of the
// parameter range deviates from the curve by at most max|P''| / (8 K^2), and P'' is linear in u, so its maximum is at
// u = 0 or 1.  K is the smallest count that keeps that bound — for the axis and for the radius — under RT3_CURVE_TOL
// times the segment's largest radius (taken at u = 0, 1/2, 1), clamped to [1, RT3_CURVE_MAX_SUBDIV].  Inside the library the sub-segments are
// ordinary linear-curve primitives; hit records are translated at the API boundary through sub_first / sub_seg:
// prim = user segment, u = (k + u_sub) / K.
#ifndef RT3_CURVE_TOL
#define RT3_CURVE_TOL 0.02
#endif
#ifndef RT3_CURVE_MAX_SUBDIV
#define RT3_CURVE_MAX_SUBDIV 64
#endif
static int curve_pieces(const CurvePoly& p) {
    double m_axis = 0.0, m_rad = 0.0;
    for (int end = 0; end < 2; end++) {  // P''(u) = 6 c0 u + 2 c1
        const double ax = 2.0 * (double)p.c[1].x + (end ? 6.0 * (double)p.c[0].x : 0.0);
        const double ay = 2.0 * (double)p.c[1].y + (end ? 6.0 * (double)p.c[0].y : 0.0);
        const double az = 2.0 * (double)p.c[1].z + (end ? 6.0 * (double)p.c[0].z : 0.0);
        const double aw = 2.0 * (double)p.c[1].w + (end ? 6.0 * (double)p.c[0].w : 0.0);
        m_axis = std::max(m_axis, std::sqrt(ax * ax + ay * ay + az * az));
        m_rad = std::max(m_rad, std::fabs(aw));
    }
    const double r0 = (double)p.c[3].w;
    const double r1 = (double)p.c[0].w + (double)p.c[1].w + (double)p.c[2].w + (double)p.c[3].w;
    const double rh = 0.125 * (double)p.c[0].w + 0.25 * (double)p.c[1].w + 0.5 * (double)p.c[2].w + (double)p.c[3].w;
    const double rref = std::max(r0, std::max(r1, rh));
    if (!(rref > 0.0)) return 1;   // nothing to see
    const double need = std::max(m_axis, m_rad) / (8.0 * RT3_CURVE_TOL * rref);   // K^2 >= need
    if (!(need <= (double)RT3_CURVE_MAX_SUBDIV * RT3_CURVE_MAX_SUBDIV)) return RT3_CURVE_MAX_SUBDIV;
    const int K = (int)std::ceil(std::sqrt(need));
    return K < 1 ? 1 : (K > RT3_CURVE_MAX_SUBDIV ? RT3_CURVE_MAX_SUBDIV : K);
}
static void tessellate_curves(int basis, const float* cp, const int32_t* seg, int nseg, Blas& b, std::vector<float>& out_cp, std::vector<int32_t>& out_seg) {
    b.poly.resize((size_t)nseg);
    b.sub_first.assign((size_t)nseg + 1, 0);
    for (int s = 0; s < nseg; s++) {
        b.poly[(size_t)s] = curve_poly(basis, cp + 4 * (size_t)seg[s]);
        b.sub_first[(size_t)s + 1] = b.sub_first[(size_t)s] + curve_pieces(b.poly[(size_t)s]);
    }
    const size_t nsub = (size_t)b.sub_first[(size_t)nseg];
    out_cp.resize(4 * (nsub + (size_t)nseg));
    out_seg.resize(nsub);
    b.sub_seg.resize(nsub);
    for (int s = 0; s < nseg; s++) {
        const int first = b.sub_first[(size_t)s], K = b.sub_first[(size_t)s + 1] - first;
        for (int k = 0; k <= K; k++) {
            const f4 v = curve_position(b.poly[(size_t)s], (float)k / (float)K);
            float* o = &out_cp[4 * ((size_t)first + (size_t)s + (size_t)k)];   // K + 1 points per segment
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            if (k < K) { out_seg[(size_t)first + (size_t)k] = first + s + k; b.sub_seg[(size_t)first + (size_t)k] = s; }
        }
    }
}
static inline void curve_hit_to_internal(const Blas& b, int32_t& prim, float& u) {
    const int first = b.sub_first[(size_t)prim], K = b.sub_first[(size_t)prim + 1] - first;
    const float f = u * (float)K;
    int k = (int)f;
    k = k > K - 1 ? K - 1 : (k < 0 ? 0 : k);
    u = f - (float)k;
    prim = first + k;
}

static inline void curve_hit_to_user(const Blas& b, int32_t& prim, float& u) {
    const int s = b.sub_seg[(size_t)prim], first = b.sub_first[(size_t)s], K = b.sub_first[(size_t)s + 1] - first;
    u = ((float)(prim - first) + u) / (float)K;
    prim = s;
}

int rt3o_curves_create(rt3o_scene* s, int degree, const float* cp, int ncp, const int32_t* seg, int nseg) {
    RT3O_TRY
    if (!s || !cp || !seg || ncp < 2 || nseg <= 0) { g_err = "curves_create: bad argument"; return -1; }
    if (degree < CURVE_LINEAR || degree > CURVE_BEZIER) { g_err = "curves_create: basis must be 1 (linear), 2 / 3 (quadratic / cubic B-spline), 4 (Catmull-Rom) or 5 (Bezier)"; return -5; }
    for (int i = 0; i < nseg; i++)
        if (seg[i] < 0 || seg[i] + curve_control_points(degree) > ncp) { g_err = "curves_create: segment out of range"; return -1; }
    auto b = std::make_unique<Blas>();
    b->type = PRIM_CURVE;
    std::vector<float> tcp;
    std::vector<int32_t> tseg;
    if (degree > 1) {
        tessellate_curves(degree, cp, seg, nseg, *b, tcp, tseg);
        cp = tcp.data(); ncp = (int)(tcp.size() / 4); seg = tseg.data(); nseg = (int)tseg.size();
    }
    b->nprims = nseg;
    b->cr.assign(cp, cp + 4 * ncp);
    b->seg.assign(seg, seg + nseg);
    return finish_blas(s, std::move(b));
    RT3O_CATCH(-1)
}
int rt3o_texture_create(rt3o_scene* s, const uint8_t* rgba8, int w, int h, int address_mode, int filter_mode) {
    RT3O_TRY
    if (!s || !rgba8 || w <= 0 || h <= 0) { g_err = "texture_create: bad argument"; return -1; }
    if (filter_mode != 0 && filter_mode != 1) { g_err = "texture_create: filter_mode must be 0 (point) or 1 (bilinear)"; return -5; }
    if (address_mode < RT3_ADDRESS_WRAP || address_mode > RT3_ADDRESS_BORDER) { g_err = "texture_create: address mode unsupported"; return -5; }
    Texture t;
    t.w = w; t.h = h; t.addr = address_mode; t.filt = filter_mode;
    t.px.assign(rgba8, rgba8 + (size_t)4 * w * h);
    s->tex.push_back(std::move(t));
    return (int)s->tex.size() - 1;
    RT3O_CATCH(-1)
}
static int add_instance(rt3o_scene* s, int blas, const float* stat, const float* keys, int nkeys, float t0, float t1) {
    if (!s || blas < 0 || blas >= (int)s->blas.size() || !stat) { g_err = "append_instance: bad argument"; return -1; }
    Instance in;
    in.blas = blas;
    std::memcpy(in.stat.m, stat, sizeof(float) * 12);
    in.stat_inv = invert_affine(in.stat);
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    in.identity = !keys && std::memcmp(stat, ident, sizeof(ident)) == 0;
    if (keys) {
        if (nkeys < 2 || !(t1 > t0)) { g_err = "append_animated_instance: need >=2 keys and t_end > t_begin"; return -1; }
        in.nkeys = nkeys;
        in.keys.assign(keys, keys + 12 * nkeys);
        in.t0 = t0; in.t1 = t1;
    }
    s->inst.push_back(std::move(in));
    s->built = false;
    return (int)s->inst.size() - 1;
}
int rt3o_accel_append_instance(rt3o_scene* s, int blas, const float xform[12]) {
    RT3O_TRY return add_instance(s, blas, xform, nullptr, 0, 0, 1); RT3O_CATCH(-1)
}
int rt3o_accel_append_animated_instance(rt3o_scene* s, int blas, const float* keys, int nkeys, float t_begin, float t_end,
                                        const float static_xform[12]) {
    RT3O_TRY
    if (!keys) { g_err = "append_animated_instance: keys null"; return -1; }
    return add_instance(s, blas, static_xform, keys, nkeys, t_begin, t_end);
    RT3O_CATCH(-1)
}
// finite and not (numerically) singular: the same decision, in the same double arithmetic, as the product's rt3_accel_build
static bool flattenable(const float* m) {
    for (int k = 0; k < 12; k++) if (!std::isfinite(m[k])) return false;
    const double det = (double)m[0] * ((double)m[5] * (double)m[10] - (double)m[6] * (double)m[9])
                     - (double)m[1] * ((double)m[4] * (double)m[10] - (double)m[6] * (double)m[8])
                     + (double)m[2] * ((double)m[4] * (double)m[9] - (double)m[5] * (double)m[8]);
    return std::fabs(det) >= 1e-30;
}

int rt3o_accel_build(rt3o_scene* s) {
    RT3O_TRY
    if (!s || s->inst.empty()) { g_err = "accel_build: no instances"; return -4; }
    // Flattening (rt3_accel_build of the product, "flatten"): static, transformed instances of plain triangle meshes are
    // intersected in world space — vertices transformed once, x' = ((m0 x + m1 y) + m2 z) + m3 — instead of transforming every
    // ray.  The hit records differ from the object-space ones in the last bits, so the oracle applies the same rule: all such
    // instances if, together with the identity ones, they stay below 2^27 triangles; none otherwise.
    {
        uint64_t sum = 0;
        for (Instance& in : s->inst) {
            const Blas& b = *s->blas[in.blas];
            in.flat = false;
            if (in.nkeys == 0 && b.type == PRIM_TRI && b.vkeys == 1 && (in.identity || flattenable(in.stat.m))) sum += (uint64_t)b.nprims;
        }
        if (g_flatten && sum < (1ull << 27))
            for (Instance& in : s->inst) {
                const Blas& b = *s->blas[in.blas];
                if (in.nkeys != 0 || in.identity || b.type != PRIM_TRI || b.vkeys != 1 || !flattenable(in.stat.m)) continue;
                in.flat = true;
                const float bx = std::max(fabsf(b.bounds.lo.x), fabsf(b.bounds.hi.x)), by = std::max(fabsf(b.bounds.lo.y), fabsf(b.bounds.hi.y)),
                            bz = std::max(fabsf(b.bounds.lo.z), fabsf(b.bounds.hi.z));
                for (int i = 0; i < 3; i++) {
                    const float* m = in.stat.m + 4 * i;
                    in.flat_w[i] = fabsf(m[0]) * bx + fabsf(m[1]) * by + fabsf(m[2]) * bz + fabsf(m[3]);
                }
            }
    }
    std::vector<Box> boxes(s->inst.size());
    for (size_t i = 0; i < s->inst.size(); i++) boxes[i] = s->instance_world_box(s->inst[i]);
    s->tlas.build(boxes);
    s->built = true;
    return 0;
    RT3O_CATCH(-1)
}
int rt3o_scene_set_hitgroup(rt3o_scene* s, int id, const float e[3], const float d[3], int tex) {
    if (!s || id < 0 || id >= (int)s->inst.size() || tex >= (int)s->tex.size()) { g_err = "set_hitgroup: bad argument"; return -1; }
    s->inst[id].emission = {e[0], e[1], e[2]};
    s->inst[id].diffuse = {d[0], d[1], d[2]};
    s->inst[id].tex = tex;
    return 0;
}
int rt3o_scene_set_texture_transform(rt3o_scene* s, int id, const float scale[2], const float rotation[2], const float offset[2]) {
    if (!s || id < 0 || id >= (int)s->inst.size() || !scale || !rotation || !offset) { g_err = "set_texture_transform: bad argument"; return -1; }
    Instance& in = s->inst[id];
    in.tex_scale = {scale[0], scale[1]}; in.tex_rot = {rotation[0], rotation[1]}; in.tex_off = {offset[0], offset[1]};
    in.has_xf = true;
    return 0;
}
int rt3o_scene_set_lights(rt3o_scene* s, const void* lights68, int n) {
    if (!s || !lights68 || n <= 0) { g_err = "set_lights: need at least one light (Q17)"; return -1; }
    s->lights.resize(n);
    std::memcpy(s->lights.data(), lights68, sizeof(Light) * n);
    s->light_cdf.resize(n);
    float run = 0.0f;
    for (int k = 0; k < n; k++) {
        run = run + light_power(s->lights[k].emission, s->lights[k].area);
        s->light_cdf[k] = run;
    }
    return 0;
}

int rt3o_trace(rt3o_scene* s, const rt3_ray* rays, int n, int any_hit, rt3_hit* hits, int accel, int nthreads) {
    RT3O_TRY
    if (!s || !s->built) { g_err = "trace: accel not built"; return -4; }
    if (n < 0 || (n > 0 && (!rays || !hits))) { g_err = "trace: bad argument"; return -1; }
    parallel_for(n, nthreads, accel == 0 ? 1 : 256, [&](int i) {
        const rt3_ray& r = rays[i];
        Hit h = s->trace({r.o[0], r.o[1], r.o[2]}, {r.d[0], r.d[1], r.d[2]}, r.tmin, r.tmax, r.time, any_hit != 0, accel);
        rt3_hit& o = hits[i];
        std::memset(&o, 0, sizeof(o));
        o.t = h.t; o.u = h.u; o.v = h.v; o.prim = h.prim; o.inst = h.inst;
        if (h.prim >= 0) {
            const Blas& b = *s->blas[s->inst[h.inst].blas];
            if (b.spline()) curve_hit_to_user(b, o.prim, o.u);
        }
    });
    return 0;
    RT3O_CATCH(-1)
}

int rt3o_get_local_geometry(rt3o_scene* s, const rt3_ray* rays, const rt3_hit* hits, int n, rt3_local_geometry* out) {
    RT3O_TRY
    if (!s || !s->built) { g_err = "get_local_geometry: accel not built"; return -4; }
    if (n < 0 || (n > 0 && (!rays || !hits || !out))) { g_err = "get_local_geometry: bad argument"; return -1; }
    static_assert(sizeof(rt3_local_geometry) == 27 * sizeof(float), "record = 27 floats");
    for (int i = 0; i < n; i++) {
        float* o = reinterpret_cast<float*>(out + i);
        if (hits[i].prim < 0) { for (int k = 0; k < 27; k++) o[k] = 0.0f; continue; }
        Hit h; h.t = hits[i].t; h.u = hits[i].u; h.v = hits[i].v; h.prim = hits[i].prim; h.inst = hits[i].inst;
        const Blas& b = *s->blas[s->inst[h.inst].blas];
        if (b.spline()) {
            if (h.prim >= (int)b.poly.size()) { g_err = "get_local_geometry: hit record does not belong to this scene"; return -1; }
            curve_hit_to_internal(b, h.prim, h.u);
        }
        const rt3_ray& r = rays[i];
        s->local_geometry_full(h, {r.o[0], r.o[1], r.o[2]}, {r.d[0], r.d[1], r.d[2]}, r.time, o);
        if (b.spline()) o[9] = hits[i].u;  // UV.x is the u along the user's segment, not along the sub-segment
    }
    return 0;
    RT3O_CATCH(-1)
}

int rt3o_launch_subframe(rt3o_scene* s, const rt3_render_settings* rs, int nthreads) {
    RT3O_TRY
    if (!s || !rs || !s->built) { g_err = "launch_subframe: accel not built"; return -4; }
    if (s->lights.empty()) { g_err = "launch_subframe: no lights (Q17)"; return -4; }
    if (rs->width == 0 || rs->height == 0 || rs->samples_per_launch == 0) { g_err = "launch_subframe: bad settings"; return -1; }
    if (s->w != rs->width || s->h != rs->height) {
        s->w = rs->width; s->h = rs->height;
        s->accum.assign((size_t)4 * s->w * s->h, 0.0f);
        s->frame.assign((size_t)4 * s->w * s->h, 0);
    }
    const int W = (int)rs->width, H = (int)rs->height;
    const int tx = (W + 15) / 16, ty = (H + 15) / 16;
    std::atomic<uint64_t> c0{0}, c1{0}, c2{0};
    parallel_for(tx * ty, nthreads, 1, [&](int tile) {
        uint64_t cnt[3] = {0, 0, 0};
        int bx = (tile % tx) * 16, by = (tile / tx) * 16;
        for (int y = by; y < std::min(by + 16, H); y++)
            for (int x = bx; x < std::min(bx + 16, W); x++) {
                f3 c = rs->mode != 0 ? s->render_pixel_corrected(*rs, (uint32_t)x, (uint32_t)y, 1, cnt) : s->render_pixel(*rs, (uint32_t)x, (uint32_t)y, 1, cnt);
                size_t pi = (size_t)y * W + x;
                float* a = &s->accum[4 * pi];
                if (rs->accum_mode == 0) {  // raygen.cu:75-86
                    if (rs->subframe_index > 0) {
                        const float k = 1.0f / (float)(rs->subframe_index + 1);
                        f3 prev = {a[0], a[1], a[2]};
                        c = prev + k * (c - prev);  // lerp: a + t*(b-a)
                    }
                    a[0] = c.x; a[1] = c.y; a[2] = c.z; a[3] = 1.0f;
                    make_color(c, &s->frame[4 * pi]);
                } else {
                    a[0] += c.x; a[1] += c.y; a[2] += c.z; a[3] += 1.0f;
                }
            }
        c0 += cnt[0]; c1 += cnt[1]; c2 += cnt[2];
    });
    s->n_primary += c0; s->n_bounce += c1; s->n_shadow += c2;
    s->n_samples += (uint64_t)W * H * rs->samples_per_launch;
    return 0;
    RT3O_CATCH(-1)
}
int rt3o_download_accum(rt3o_scene* s, float* rgba) {
    if (!s || !rgba || s->accum.empty()) { g_err = "download_accum: nothing rendered"; return -4; }
    std::memcpy(rgba, s->accum.data(), s->accum.size() * sizeof(float));
    return 0;
}
int rt3o_download_frame(rt3o_scene* s, uint8_t* rgba8) {
    if (!s || !rgba8 || s->frame.empty()) { g_err = "download_frame: nothing rendered"; return -4; }
    std::memcpy(rgba8, s->frame.data(), s->frame.size());
    return 0;
}
int rt3o_get_stats(rt3o_scene* s, rt3_stats* st) {
    if (!s || !st) return -1;
    std::memset(st, 0, sizeof(*st));
    st->rays_primary = s->n_primary; st->rays_bounce = s->n_bounce; st->rays_shadow = s->n_shadow; st->samples = s->n_samples;
    return 0;
}
int rt3o_reset_stats(rt3o_scene* s) {
    if (!s) return -1;
    s->n_primary = 0; s->n_bounce = 0; s->n_shadow = 0; s->n_samples = 0;
    return 0;
}

void rt3o_set_chain_sum(int on) { g_chain_sum = on != 0; }
void rt3o_set_flatten(int on) { g_flatten = on != 0; }
void rt3o_set_libm_sincos(int on) { g_libm_sincos = on != 0; }

// ---------------------------------------------------------------- KAT hooks
uint32_t rt3o_kat_tea4(uint32_t a, uint32_t b) { return tea4(a, b); }
float rt3o_kat_rnd(uint32_t* seed) { return rnd(*seed); }
void rt3o_kat_cosine_sample(float u1, float u2, float o[4]) {
    f3 p = sample_cosine_hemisphere(u1, u2);
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = (float)((double)p.z / 3.14159265358979323846);
}
void rt3o_kat_onb(const float n[3], const float w[3], float o[9]) {
    Onb b({n[0], n[1], n[2]});
    f3 p = b.inverse_transform({w[0], w[1], w[2]});
    o[0] = b.t.x; o[1] = b.t.y; o[2] = b.t.z; o[3] = b.b.x; o[4] = b.b.y; o[5] = b.b.z; o[6] = p.x; o[7] = p.y; o[8] = p.z;
}
int rt3o_kat_fetch_texture(rt3o_scene* s, int t, float u, float v, float out[3]) {
    if (!s || t < 0 || (size_t)t >= s->tex.size()) return -1;
    const f3 c = s->fetch_texture(t, u, v);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
    return 0;
}
int rt3o_kat_sample_texture(rt3o_scene* s, int id, float u, float v, float out[3]) {
    if (!s || id < 0 || (size_t)id >= s->inst.size() || s->inst[id].tex < 0) return -1;
    const f3 c = s->sample_texture(s->inst[id], {u, v});
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
    return 0;
}
void rt3o_kat_light_make(const float e[3], const float v0[3], const float v1[3], const float v2[3], void* out) {
    Light l = light_make({e[0], e[1], e[2]}, {v0[0], v0[1], v0[2]}, {v1[0], v1[1], v1[2]}, {v2[0], v2[1], v2[2]});
    std::memcpy(out, &l, sizeof(l));
}
void rt3o_kat_light_sample(const void* light68, const float P[3], uint32_t* seed, float o[7]) {
    Light l;
    std::memcpy(&l, light68, sizeof(l));
    f3 pos, em;
    float pdf;
    light_sample(l, {P[0], P[1], P[2]}, *seed, pos, em, pdf);
    o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = em.x; o[4] = em.y; o[5] = em.z; o[6] = pdf;
}
void rt3o_kat_make_color(const float c[3], uint8_t out[4]) { make_color({c[0], c[1], c[2]}, out); }
void rt3o_kat_camera_uvw(const float e[3], const float l[3], const float u[3], float fovy, float aspect, float o[9]) {
    f3 U, V, W;
    camera_uvw({e[0], e[1], e[2]}, {l[0], l[1], l[2]}, {u[0], u[1], u[2]}, fovy, aspect, U, V, W);
    o[0] = U.x; o[1] = U.y; o[2] = U.z; o[3] = V.x; o[4] = V.y; o[5] = V.z; o[6] = W.x; o[7] = W.y; o[8] = W.z;
}
void rt3o_kat_curve_eval(int basis, const float* cp, float u, float out[16]) {
    const CurvePoly p = curve_poly(basis, cp);
    const f4 a = curve_position(p, u), v = curve_velocity(p, u), c = curve_acceleration(p, u);
    const f3 t = curve_tangent(p, u);
    const float o[16] = {a.x, a.y, a.z, a.w, v.x, v.y, v.z, v.w, c.x, c.y, c.z, c.w, t.x, t.y, t.z, 0.0f};
    for (int k = 0; k < 16; k++) out[k] = o[k];
}
void rt3o_kat_sincos_2pi(float u, float o[2]) { sincos_2pi(u, o[0], o[1]); }
void rt3o_kat_invert_affine(const float m[12], float out[12]) {
    Affine a;
    std::memcpy(a.m, m, sizeof(a.m));
    Affine r = invert_affine(a);
    std::memcpy(out, r.m, sizeof(r.m));
}
int rt3o_kat_hit_triangle(const float o[3], const float d[3], const float v[9], float tmin, float tmax, float out[3]) {
    RayShear s = make_shear({d[0], d[1], d[2]});
    return hit_triangle({o[0], o[1], o[2]}, s, {v[0], v[1], v[2]}, {v[3], v[4], v[5]}, {v[6], v[7], v[8]}, tmin, tmax, out[0], out[1], out[2]) ? 1 : 0;
}
int rt3o_kat_hit_sphere(const float o[3], const float d[3], const float cr[4], float tmin, float tmax, float* t) {
    return hit_sphere({o[0], o[1], o[2]}, {d[0], d[1], d[2]}, {cr[0], cr[1], cr[2]}, cr[3], tmin, tmax, *t) ? 1 : 0;
}
int rt3o_kat_hit_curve(const float o[3], const float d[3], const float a[4], const float b[4], float tmin, float tmax, float out[2]) {
    return hit_curve_linear({o[0], o[1], o[2]}, {d[0], d[1], d[2]}, {a[0], a[1], a[2]}, a[3], {b[0], b[1], b[2]}, b[3], tmin, tmax, out[0], out[1]) ? 1 : 0;
}

}  // extern "C"
