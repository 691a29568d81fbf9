// rt3_traverse.cuh — two-level software traversal (replaces optixTraverse on RT cores; reference
// call sites src/shader/shader_common.h:74-88 closest hit, :119-133 occlusion; no reference source).
//
// Per-thread state machine over the compressed BVH8 (rt3_bvh.cuh):
//   step() is one round of a while-while loop: pop -> intersect one wide node -> test all pending
//   primitives / enter one instance.  The persistent kernels in rt3_wavefront.cuh drive 32 of these
//   per warp and refill finished lanes from the ray queue (dynamic fetch).
// Primitive tests (same operation order as the CPU oracle, oracle/rt3o_prims.hpp):
//   triangle: watertight (Woop/Benthin/Wald 2013), barycentrics u->v1, v->v2;
//   sphere  : cuda/sphere.cu:44-96;  curve: round linear segment, entry hits only.
// Closest hit keeps the lexicographically smallest (t, instance, primitive) so the result does
// not depend on traversal order; box culling is padded so it never removes such a candidate.
#pragma once
#include "rt3_bvh.cuh"
#ifndef RT3_EMULATE
#include <cuda_fp16.h>
#endif

namespace rt3 {

enum { PRIM_TRI = 0, PRIM_SPHERE = 1, PRIM_CURVE = 2, PRIM_TRI_MOTION = 3 };  // 3: triangles with vertex keys (deformation blur)
#ifndef RT3_STACK_SIZE
#define RT3_STACK_SIZE 64
#endif
#ifndef RT3_COOP_CAP
#define RT3_COOP_CAP 32   // (owner, triangle) pairs a warp shares per round
#endif
#ifndef RT3_TRAV_THREADS
#define RT3_TRAV_THREADS 128
#endif
#ifndef RT3_SMEM_STACK
#define RT3_SMEM_STACK 0   // >0: keep that many bottom stack entries per lane in shared memory (measured slower: ptxas spills)
#endif
#ifndef RT3_DEFER
#define RT3_DEFER 1       // single-level kernel: triangles wait in a per-warp queue until a pass is worth it (step_warp_deferred)
#endif
#ifndef RT3_QCAP
#define RT3_QCAP 64       // capacity of that queue
#endif
#ifndef RT3_DEFER_ITEMS
#define RT3_DEFER_ITEMS 32    // run the triangle pass once this many pairs are queued (a full-width pass) ...
#endif
#ifndef RT3_DEFER_BLOCKED
#define RT3_DEFER_BLOCKED 4   // ... or more than this many lanes have nothing else left to do
#endif
#ifndef RT3_DEFER_PARTIAL
#define RT3_DEFER_PARTIAL 1   // 1: a pass takes 32 pairs at most and leaves the rest queued (no near-empty second pass: +3-5 %)
#endif
#ifndef RT3_BSPHERE_TNEAR
#define RT3_BSPHERE_TNEAR 0   // 1: instance bounding spheres also cull by entry distance against the closest hit so far (measured: entries per ray 3.38 -> 3.31 on C3, not worth a sqrt)
#endif
#ifndef RT3_TRI_BATCH
#define RT3_TRI_BATCH 16    // general kernel: run the triangle phase only when the warp holds at least this many pending pairs (0 / 1 = every round)
#endif
#ifndef RT3_ENTRY_BATCH
#define RT3_ENTRY_BATCH 8   // general kernel: run the entry / sphere / curve phase only when at least this many lanes want it (0 / 1 = every round)
#endif
#ifndef RT3_COOP_MIN
#define RT3_COOP_MIN 0    // general kernel: redistribute a round's triangles only when the warp holds at least this many (0 = always when a lane has two)
#endif
#ifndef RT3_COOP
#define RT3_COOP 1        // 1: warp-cooperative triangle phase (step_warp), 0: per-lane loop (step)
#endif

struct BlasDev {             // per geometry, device-resident table entry
    const Node8* nodes;
    const float4* prims;     // 3 x float4 per primitive, node-contiguous order (see k_pack_*)
    uint32_t type, nprims;
    // shading attributes, ORIGINAL primitive order (reference HitGroupData arrays, shader_data.h:125-136)
    const int32_t* idx;      // mesh [nt][3]
    const float* normals;    // mesh [nv][3]
    const float* uvs;        // mesh [nv][2]
    const float4* cr;        // spheres [n] / curve control points [ncp]
    const int32_t* seg;      // curves [nseg]
    uint32_t vkeys;          // PRIM_TRI_MOTION: vertex keys per triangle record (record = vkeys x 3 float4)
    const float* verts;      // mesh [vkeys][nv][3] (corrected mode: area of BSDF-sampled emitter hits; rt3_get_local_geometry)
    uint32_t nv;             // vertices per key
    const uint32_t* sub;     // spline curves: (user segment, k | K << 16) of every linear sub-segment; null otherwise
    const float* colors;     // mesh, optional [nv][4] vertex colours (cuda/LocalGeometry.h:99-110); normals / uvs may be null too (SDK fallbacks)
    const float4* poly;      // spline curves: power-basis coefficients c0 u^3 + c1 u^2 + c2 u + c3 of every USER segment (xyz + radius)
    uint32_t curve_cubic;    // spline curves: 1 = cubic interpolator
};

struct InstanceDev {         // traversal record (64 B)
    float inv_static[12];    // world -> object of the static instance transform
    uint32_t blas;
    uint32_t nkeys : 16;     // 0 = no motion
    uint32_t identity : 16;  // 1: static transform is exactly the identity and there is no motion; 2: merged world BLAS
    uint32_t key_offset;     // first float of this instance's keys in TravScene::keys
    float t0;                // motion begin; end in t1 (kept in the shading record to stay at 64 B)
};
static_assert(sizeof(InstanceDev) == 64, "instance record is 64 bytes");

struct HitGroupDev {         // shading record (reference HitGroupData + motion end time)
    float emission[3];
    float diffuse[3];
    int32_t tex;
    float t1;
    // MaterialData::Texture of the SDK's sampleTexture (cuda/LocalShading.h:37-54): UV * scale, rotated by (sin, cos), + offset;
    // has_xf = 0 (the default, and all the reference's src/ path ever does) fetches at the plain UV
    float tex_scale[2], tex_rot[2], tex_off[2];
    uint32_t has_xf, pad_;
};
static_assert(sizeof(HitGroupDev) == 64, "shading record is 64 bytes");

struct TravScene {
    const Node8* tlas_nodes;
    const uint32_t* tlas_order;     // TLAS leaf slot -> instance id
    const InstanceDev* instances;
    const HitGroupDev* hitgroups;
    const BlasDev* blas;
    const float* keys;
    const float* inst_fwd;          // [instance][12] object -> world of the static instance transform (rt3_get_local_geometry)
    // [instance][2] world-space bounding sphere of the instanced geometry: {centre at the first motion key, radius} {centre(last key) -
    // centre(first key), 1 / (t1 - t0) or 0 when static}; radius < 0 = none (identity instances, more than two keys).  A TLAS leaf
    // whose sphere the ray misses is not entered: the leaf box is the world AABB of a rotated object, and for compact objects about half
    // of the rays through such a box miss the object's bounding sphere, while an entry costs a transform, three divisions and a BLAS root
    const float4* inst_bsphere;
    uint32_t* error_flags;          // bit0 stack overflow
    uint32_t* max_stack;
    // single-level fast path: all identity, static triangle-mesh instances (every instance the
    // reference creates, cuda_scene.h:141-146) are merged into ONE world-space BLAS whose primitive
    // ids index merged_map -> (instance, primitive).  If nothing else is in the scene the TLAS is
    // skipped altogether (root_is_blas) and tlas_nodes points at the merged BLAS.
    uint32_t magic_h2;              // 0x64646464 and
    float bias_h;                   // -1024.0f: see widen_bytes / byte_to_float (kept opaque to the compiler on purpose)
    const uint2* merged_map;
    const float4* root_prims;
    uint32_t root_is_blas;
};
#define RT3_MERGED_INST 0x7fffffff

struct HitRec { float t, u, v; int prim, inst; };

// object-space ray of an instance at a ray time (shared by traversal and the shade stage)
RT3_HD void instance_ray(const TravScene& sc, const InstanceDev* in, float t1, float time, float3 wo, float3 wd, float3& oo, float3& od) {
    if (in->identity) { oo = wo; od = wd; return; }
    Affine si;
#pragma unroll
    for (int j = 0; j < 12; j++) si.m[j] = in->inv_static[j];
    oo = xform_point(si, wo);
    od = xform_vector(si, wd);
    if (in->nkeys > 0) {
        const Affine m = lerp_keys(sc.keys + in->key_offset, (int)in->nkeys, in->t0, t1, time);
        const Affine mi = invert_affine(m);
        oo = xform_point(mi, oo);
        od = xform_vector(mi, od);
    }
}

// ------------------------------------------------------------------------------------ primitive tests
struct Shear { int kx, ky, kz; float Sx, Sy, Sz; };

// Shear constants of the watertight test.  The reciprocal of the dominant component is the ONLY division: Sx, Sy are
// products with it (the test stays watertight — every triangle of a ray sees the same constants), and it is the same
// clamped reciprocal the slab test uses for that axis, so a ray set-up costs three divisions instead of six.
#define RT3_DIR_EPS 8.271806e-25f  // 2^-80: keeps 1/d finite for axis-parallel rays
RT3_HD float clamp_dir(float d) { return fabsf(d) > RT3_DIR_EPS ? d : copysignf(RT3_DIR_EPS, d); }
RT3_HD Shear make_shear(float3 d, float3 idir) {  // idir = 1 / clamp_dir(d) per component
    Shear s;
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    s.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    s.kx = s.kz + 1; if (s.kx == 3) s.kx = 0;
    s.ky = s.kx + 1; if (s.ky == 3) s.ky = 0;
    const float dz = comp(d, s.kz);
    if (dz < 0.0f) { const int t = s.kx; s.kx = s.ky; s.ky = t; }
    s.Sz = comp(idir, s.kz);
    s.Sx = comp(d, s.kx) * s.Sz;
    s.Sy = comp(d, s.ky) * s.Sz;
    return s;
}

RT3_HD bool test_triangle(float3 o, const Shear& s, float3 v0, float3 v1, float3 v2, float& t, float& u, float& v) {
    const float3 A = sub(v0, o), B = sub(v1, o), C = sub(v2, o);
    const float Akz = comp(A, s.kz), Bkz = comp(B, s.kz), Ckz = comp(C, s.kz);
    const float Ax = comp(A, s.kx) - s.Sx * Akz;
    const float Ay = comp(A, s.ky) - s.Sy * Akz;
    const float Bx = comp(B, s.kx) - s.Sx * Bkz;
    const float By = comp(B, s.ky) - s.Sy * Bkz;
    const float Cx = comp(C, s.kx) - s.Sx * Ckz;
    const float Cy = comp(C, s.ky) - s.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {  // edge-on: redo the edge functions in double
        const double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        const double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        const double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = s.Sz * Akz, Bz = s.Sz * Bkz, Cz = s.Sz * Ckz;
    const float T = U * Az + V * Bz + W * Cz;
    const float rcp = 1.0f / det;
    t = T * rcp;
    u = V * rcp;
    v = W * rcp;
    return true;
}

// returns up to two candidate roots in order (cuda/sphere.cu:44-96); caller applies the interval
RT3_HD int test_sphere(float3 o, float3 d, float3 center, float radius, float& ta, float& tb) {
    const float3 O = sub(o, center);
    const float l = 1.0f / length(d);
    const float3 D = mul(d, l);
    float b = dot(O, D);
    float c = dot(O, O) - radius * radius;
    float disc = b * b - c;
    if (!(disc > 0.0f)) return 0;
    float sdisc = sqrtf(disc);
    const float root1 = (-b - sdisc);
    float root11 = 0.0f;
    const bool do_refine = fabsf(root1) > (10.0f * radius);
    if (do_refine) {
        const float3 O1 = add(O, mul(D, root1));
        b = dot(O1, D);
        c = dot(O1, O1) - radius * radius;
        disc = b * b - c;
        if (disc > 0.0f) {
            sdisc = sqrtf(disc);
            root11 = (-b - sdisc);
        }
    }
    ta = (root1 + root11) * l;
    const float root2 = (-b + sdisc) + (do_refine ? root1 : 0.0f);
    tb = root2 * l;
    return 2;
}

RT3_HD bool test_curve_linear(float3 o, float3 d, float3 pa, float ra, float3 pb, float rb, float& t, float& u) {
    const float l = 1.0f / length(d);
    const float3 D = mul(d, l);
    const float t0 = dot(sub(pa, o), D);
    const float3 ro = add(o, mul(D, t0));
    const float3 ba = sub(pb, pa);
    const float3 oa = sub(ro, pa);
    const float3 ob = sub(ro, pb);
    const float rr = ra - rb;
    const float m0 = dot(ba, ba);
    const float m1 = dot(ba, oa);
    const float m2 = dot(ba, D);
    const float m3 = dot(D, oa);
    const float m5 = dot(oa, oa);
    const float m6 = dot(ob, D);
    const float m7 = dot(ob, ob);
    const float d2 = m0 - rr * rr;
    bool found = false;
    float tn = 0.0f, un = 0.0f;
    if (d2 > 0.0f) {
        const float k2 = d2 - m2 * m2;
        const float k1 = d2 * m3 - m1 * m2 + m2 * rr * ra;
        const float k0 = d2 * m5 - m1 * m1 + m1 * rr * ra * 2.0f - m0 * ra * ra;
        const float h = k1 * k1 - k0 * k2;
        if (h > 0.0f && k2 != 0.0f) {
            const float tb = (-sqrtf(h) - k1) / k2;
            const float y = m1 - ra * rr + tb * m2;
            if (y > 0.0f && y < d2) { found = true; tn = tb; un = y / d2; }
        }
    }
    if (!found) {
        const float h1 = m3 * m3 - m5 + ra * ra;
        const float h2 = m6 * m6 - m7 + rb * rb;
        float best = 3.0e38f;
        if (h1 > 0.0f) { best = -m3 - sqrtf(h1); un = 0.0f; found = true; }
        if (h2 > 0.0f) {
            const float tc = -m6 - sqrtf(h2);
            if (tc < best) { best = tc; un = 1.0f; }
            found = true;
        }
        tn = best;
    }
    if (!found) return false;
    t = (t0 + tn) * l;
    u = un;
    return true;
}

// hit word of child j: its child bits (1 for an internal child, the unary primitive count for a leaf) at its bit index
RT3_HD uint32_t child_word(uint32_t child_bits4, uint32_t bit_index4, int j) {
#ifdef RT3_EMULATE
    return ((child_bits4 >> (8 * j)) & 0xffu) << ((bit_index4 >> (8 * j)) & 31u);
#else
    return __byte_perm(child_bits4, 0u, 0x4440u + (uint32_t)j) << ((bit_index4 >> (8 * j)) & 31u);  // PRMT, SHF.R, SHF.L.W
#endif
}

// The four quantised bytes of a packed word as floats, exact.  Device: bytes (0,1) and (2,3) are widened to
// fp16 pairs {0x64 b} = 1024 + b with ONE PRMT per pair, and the mixed-precision add of sm_100
// (add.f32.f16 -> FHADD, half operand selected in place) removes the 1024 and widens to fp32: 24 PRMT +
// 48 FHADD per wide node instead of 48 PRMT + 48 FADD.  `magic` (0x64646464) and `bias` (-1024.0f) are
// kernel parameters so that ptxas cannot treat them as constants: PRMT encodes ONE immediate, and with
// both operands constant ptxas keeps the magic as the immediate and parks the selectors in uniform
// registers, which costs a UR->R move in front of every PRMT (205 of 1584 SASS instructions, r01k).
struct Bytes4 { uint32_t h01, h23; };
RT3_HD Bytes4 widen_bytes(uint32_t w, uint32_t magic) {
    Bytes4 r;
#ifdef RT3_EMULATE
    (void)magic;
    r.h01 = w; r.h23 = w;
#else
    asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(r.h01) : "r"(w), "r"(magic));
    asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(r.h23) : "r"(w), "r"(magic));
#endif
    return r;
}
RT3_HD float byte_to_float(const Bytes4& b, int j, float bias) {
#ifdef RT3_EMULATE
    (void)bias;
    return (float)((b.h01 >> (8 * j)) & 0xffu);
#else
    float f;
    const uint32_t h2 = j < 2 ? b.h01 : b.h23;
    if (j & 1) asm("{\n .reg .b16 lo, hi;\n mov.b32 {lo, hi}, %1;\n add.f32.f16 %0, hi, %2;\n}" : "=f"(f) : "r"(h2), "f"(bias));
    else asm("{\n .reg .b16 lo, hi;\n mov.b32 {lo, hi}, %1;\n add.f32.f16 %0, lo, %2;\n}" : "=f"(f) : "r"(h2), "f"(bias));
    return f;
#endif
}

// (a.x * m + c, a.y * m + c), each one IEEE fma like fmaf(): ONE packed FFMA2 on sm_100 (the scalars are broadcast operands)
RT3_HD float2 fma2(float ax, float ay, float m, float c) {
#ifdef RT3_EMULATE
    return make_float2(fmaf(ax, m, c), fmaf(ay, m, c));
#else
    float2 r;
    asm("{\n .reg .b64 a, m, c, d;\n mov.b64 a, {%2, %3};\n mov.b64 m, {%4, %4};\n mov.b64 c, {%5, %5};\n fma.rn.f32x2 d, a, m, c;\n mov.b64 {%0, %1}, d;\n}"
        : "=f"(r.x), "=f"(r.y) : "f"(ax), "f"(ay), "f"(m), "f"(c));
    return r;
#endif
}

// two fp16 values packed in a word -> float2 (x = low half); exact.  Device: 2 x HADD2.F32
RT3_HD float2 half2_to_float2(uint32_t w) {
#ifdef RT3_EMULATE
    float r[2];
    for (int i = 0; i < 2; i++) {
        const uint32_t h = (w >> (16 * i)) & 0xffffu;
        const uint32_t e = (h >> 10) & 31u, m = h & 0x3ffu;
        r[i] = e == 0 ? ldexpf((float)m, -24) : ldexpf((float)(m | 0x400u), (int)e - 25);
        if (h & 0x8000u) r[i] = -r[i];
    }
    return make_float2(r[0], r[1]);
#else
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
#endif
}

// ------------------------------------------------------------------------------------ traversal state machine
#ifndef RT3_PREFETCH
#define RT3_PREFETCH 0
#endif
RT3_HD void prefetch_l1(const void* p) {
#if !defined(RT3_EMULATE) && RT3_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Register diet: only what every node / triangle test needs lives in registers (origin, 1/d,
// shear constants, interval, hit, cursors).  The rest — the current-space direction, the ray
// time and the saved world-space ray — sits in a "frame" behind the traversal stack in local
// memory (L1-resident) and is touched only by sphere / curve tests and instance entry / exit.
enum { FR_D_XY = 0, FR_DZ_TIME, FR_WO_XY, FR_WOZ_WDX, FR_WD_YZ, FR_WIDIR_XY, FR_WIDIRZ_INV, FR_WS_XY, FR_WSZ, FR_COUNT };

#if !defined(RT3_EMULATE) && RT3_SMEM_STACK > 0
static __shared__ uint2 rt3_s_stack[RT3_SMEM_STACK][RT3_TRAV_THREADS];  // [entry][thread of the CTA]
#endif

// SINGLE = the scene is the merged world BLAS only (sc.root_is_blas): no TLAS, no instance entry /
// exit, static triangles only.  The kernel instantiated with SINGLE = true drops every other path
// (their code costs the hot loop registers even when it never runs).
template <bool ANY_HIT, bool SINGLE>
struct Trav {
    float3 o;          // current-space origin
    float3 idir;       // 1 / current-space direction (clamped)
    float tmin, tbest;
    uint32_t inv;      // bits 0-2: d_k >= 0 per axis; bits 8-9 kx, 10-11 ky, 12-13 kz (watertight shear axes)
    float Sx, Sy, Sz;  // watertight shear constants
    float hu, hv;      // closest hit so far
    int hprim, hinst;
    const Node8* nodes;
    const float4* prims;
    uint32_t ptype;
    int cur_inst;      // -1 while in the TLAS
    uint2 ng, tg;
    int sp;
    uint32_t pend;     // triangles of this lane waiting in the warp's queue (step_warp_deferred)
#ifdef RT3_STATS
    uint32_t c_nodes, c_prims, c_rounds, c_entries;  // diagnostic build only (-DRT3_STATS): wide nodes, primitive tests, rounds, instance entries
    uint32_t* dbg;
#endif
    // the traversal stack (+ frame) is a separate per-thread array owned by the kernel: as a member its
    // dynamically indexed stores forced the WHOLE state into local memory (every push was followed by a
    // reload of origin, 1/d, ... and every node step stored the cursors back)
    uint2* stack;
    // ncu: the per-thread stack in local memory (1280 threads/SM) competes with nodes and triangles for
    // L1 and spilled to DRAM (99 B written per ray for a 20 B hit record); the hot bottom of the stack
    // therefore sits in shared memory, column per thread (bank-conflict free), the rest stays local.
#if !defined(RT3_EMULATE) && RT3_SMEM_STACK > 0
    RT3_HD void st_put(int i, uint2 e) { if (i < RT3_SMEM_STACK) rt3_s_stack[i][threadIdx.x] = e; else stack[i - RT3_SMEM_STACK] = e; }
    RT3_HD uint2 st_get(int i) const { return i < RT3_SMEM_STACK ? rt3_s_stack[i][threadIdx.x] : stack[i - RT3_SMEM_STACK]; }
#else
    RT3_HD void st_put(int i, uint2 e) { stack[i] = e; }
    RT3_HD uint2 st_get(int i) const { return stack[i]; }
#endif

    RT3_HD void fr_set(int k, float a, float b) { stack[RT3_STACK_SIZE + k] = make_uint2(rt3_f2u(a), rt3_f2u(b)); }
    RT3_HD float2 fr_get(int k) const { const uint2 v = stack[RT3_STACK_SIZE + k]; return make_float2(rt3_u2f(v.x), rt3_u2f(v.y)); }
    RT3_HD float3 cur_d() const { const float2 a = fr_get(FR_D_XY), b = fr_get(FR_DZ_TIME); return v3(a.x, a.y, b.x); }
    RT3_HD float ray_time() const { return fr_get(FR_DZ_TIME).y; }

    RT3_HD void set_space(float3 oo, float3 dd, float time, bool keep_frame = true) {
        o = oo;
        if (keep_frame) {
            fr_set(FR_D_XY, dd.x, dd.y);
            fr_set(FR_DZ_TIME, dd.z, time);
        }
        const float dx = clamp_dir(dd.x), dy = clamp_dir(dd.y), dz = clamp_dir(dd.z);
        idir = v3(1.0f / dx, 1.0f / dy, 1.0f / dz);
        const Shear s = make_shear(dd, idir);
        Sx = s.Sx; Sy = s.Sy; Sz = s.Sz;
        inv = (dx >= 0.0f ? 1u : 0u) | (dy >= 0.0f ? 2u : 0u) | (dz >= 0.0f ? 4u : 0u) | ((uint32_t)s.kx << 8) | ((uint32_t)s.ky << 10) |
              ((uint32_t)s.kz << 12);
    }
    RT3_HD Shear shear() const {
        Shear s;
        s.kx = (int)((inv >> 8) & 3u); s.ky = (int)((inv >> 10) & 3u); s.kz = (int)((inv >> 12) & 3u);
        s.Sx = Sx; s.Sy = Sy; s.Sz = Sz;
        return s;
    }

    RT3_HD void init(const TravScene& sc, float3 ro, float3 rd, float rtmin, float rtmax, float rtime) {
        tmin = rtmin; tbest = rtmax;
        pend = 0u;
        hu = hv = 0.0f; hprim = -1;
        if (!SINGLE) hinst = -1;
        if (!SINGLE) { nodes = sc.tlas_nodes; prims = sc.root_prims; ptype = PRIM_TRI; cur_inst = sc.root_is_blas ? RT3_MERGED_INST : -1; }
        ng = make_uint2(0u, 0x80000000u);
        tg = make_uint2(0u, 0u);
        sp = 0;
#ifdef RT3_STATS
        dbg = sc.error_flags;
#endif
#ifdef RT3_STATS
        c_nodes = c_prims = c_rounds = c_entries = 0;
#endif
        // single-level scenes (merged world BLAS only) never read the frame: skip its local-memory stores
        set_space(ro, rd, rtime, !SINGLE);
        if (!SINGLE) {  // world-space copy for leaving transformed instances
            fr_set(FR_WO_XY, ro.x, ro.y);
            fr_set(FR_WOZ_WDX, ro.z, rd.x);
            fr_set(FR_WD_YZ, rd.y, rd.z);
            fr_set(FR_WIDIR_XY, idir.x, idir.y);
            fr_set(FR_WIDIRZ_INV, idir.z, rt3_u2f(inv));
            fr_set(FR_WS_XY, Sx, Sy);
            fr_set(FR_WSZ, Sz, 0.0f);
        }
    }
    RT3_HD void restore_world() {
        const float2 a = fr_get(FR_WO_XY), b = fr_get(FR_WOZ_WDX), c = fr_get(FR_WD_YZ), e = fr_get(FR_WIDIR_XY), f = fr_get(FR_WIDIRZ_INV),
                     g = fr_get(FR_WS_XY), h = fr_get(FR_WSZ);
        o = v3(a.x, a.y, b.x);
        const float time = ray_time();
        fr_set(FR_D_XY, b.y, c.x);
        fr_set(FR_DZ_TIME, c.y, time);
        idir = v3(e.x, e.y, f.x);
        inv = rt3_f2u(f.y);
        Sx = g.x; Sy = g.y; Sz = h.x;
    }

    RT3_HD void push(const TravScene& sc, uint2 e) {
        if (sp < RT3_STACK_SIZE) st_put(sp++, e);
        else rt3_atomic_or(sc.error_flags, 1u);   // the subtree is lost: rt3_trace / rt3_download_* report it
#if defined(RT3_STATS) && !defined(RT3_EMULATE)
        if ((uint32_t)sp > *sc.max_stack) rt3_atomic_max(sc.max_stack, (uint32_t)sp);   // high-water mark, diagnostic builds
#endif
    }

    RT3_HD bool in_range(float t) const { return t > tmin && (hprim < 0 ? t < tbest : t <= tbest); }

    // returns true if the candidate was accepted
    RT3_HD bool accept(const TravScene& sc, float t, float u, float v, int prim) {
        if (!in_range(t)) return false;
        if (!ANY_HIT && hprim >= 0 && t == tbest) {  // exact tie: lowest (instance, primitive) wins
            int ci = SINGLE ? RT3_MERGED_INST : cur_inst, cp = prim, bi = SINGLE ? RT3_MERGED_INST : hinst, bp = hprim;
            if (!SINGLE && ci == RT3_MERGED_INST) { const uint2 m = sc.merged_map[cp]; ci = (int)m.x; cp = (int)m.y; }
            if (!SINGLE && bi == RT3_MERGED_INST) { const uint2 m = sc.merged_map[bp]; bi = (int)m.x; bp = (int)m.y; }  // SINGLE: merged ids already order like (instance, primitive)
            if (!(ci < bi || (ci == bi && cp < bp))) return false;
        }
        tbest = t; hu = u; hv = v; hprim = prim; hinst = SINGLE ? RT3_MERGED_INST : cur_inst;
        return true;
    }

    RT3_HD void node_step(const TravScene& sc) {
#ifdef RT3_STATS
        c_nodes++;
#endif
        const uint32_t hits = ng.y;
        const int bit = 31 - rt3_clz(hits);
        ng.y &= ~(1u << bit);
        if (ng.y & 0xff000000u) push(sc, ng);
        const uint32_t oct = inv & 7u;
        const uint32_t slot = ((uint32_t)bit - 24u) ^ oct;
        const uint32_t rel = (uint32_t)rt3_popc(hits & 0xffu & ((1u << slot) - 1u));
        const uint4* np = reinterpret_cast<const uint4*>((SINGLE ? sc.tlas_nodes : nodes) + (ng.x + rel));
        const uint4 n0 = rt3_ldg(np + 0), n1 = rt3_ldg(np + 1);

        const float adjx = rt3_u2f((n0.w & 0xffu) << 23) * idir.x;
        const float adjy = rt3_u2f(((n0.w >> 8) & 0xffu) << 23) * idir.y;
        const float adjz = rt3_u2f(((n0.w >> 16) & 0xffu) << 23) * idir.z;
        const float orgx = (rt3_u2f(n0.x) - o.x) * idir.x;
        const float orgy = (rt3_u2f(n0.y) - o.y) * idir.y;
        const float orgz = (rt3_u2f(n0.z) - o.z) * idir.z;
        // conservative padding of the slab distances (a few ulps of the largest operand)
        const float keps = 4.76837158e-7f;  // 2^-21
        const float padx = keps * (fabsf(orgx) + (float)RT3_QMAX * fabsf(adjx));
        const float pady = keps * (fabsf(orgy) + (float)RT3_QMAX * fabsf(adjy));
        const float padz = keps * (fabsf(orgz) + (float)RT3_QMAX * fabsf(adjz));
        const float nox = orgx - padx, fox = orgx + padx;
        const float noy = orgy - pady, foy = orgy + pady;
        const float noz = orgz - padz, foz = orgz + padz;

        uint32_t hitmask = 0;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            // meta bytes of 4 children at once (after Ylitie et al.): internal children land on bit
            // 24 + (slot ^ oct) so that the nearest child is the highest set bit; leaf children put
            // their unary primitive count at their primitive offset; empty slots contribute 0 bits
            const uint32_t meta4 = half ? n1.w : n1.z;
            const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t oct_inner4 = (is_inner4 >> 4) * oct;  // the ray octant in the bytes of internal children
            uint32_t bit_index4 = (meta4 ^ oct_inner4) & 0x1f1f1f1fu;
            uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#ifndef RT3_EMULATE
            // pin the 4-wide decode: ptxas otherwise sinks it into the predicated per-child code and
            // redoes it for every child (87 instead of ~50 instructions for the eight hit words)
            asm volatile("" : "+r"(bit_index4), "+r"(child_bits4));
#endif
#if RT3_NODE_FP16
            const uint4 qx = rt3_ldg(np + 2 + 3 * half), qy = rt3_ldg(np + 3 + 3 * half), qz = rt3_ldg(np + 4 + 3 * half);
            // {lo01, lo23, hi01, hi23} per axis; near plane = lo when d >= 0
            const uint32_t nx[2] = {(oct & 1u) ? qx.x : qx.z, (oct & 1u) ? qx.y : qx.w}, fx[2] = {(oct & 1u) ? qx.z : qx.x, (oct & 1u) ? qx.w : qx.y};
            const uint32_t ny[2] = {(oct & 2u) ? qy.x : qy.z, (oct & 2u) ? qy.y : qy.w}, fy[2] = {(oct & 2u) ? qy.z : qy.x, (oct & 2u) ? qy.w : qy.y};
            const uint32_t nz[2] = {(oct & 4u) ? qz.x : qz.z, (oct & 4u) ? qz.y : qz.w}, fz[2] = {(oct & 4u) ? qz.z : qz.x, (oct & 4u) ? qz.w : qz.y};
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {
                const float2 anx = half2_to_float2(nx[pr]), any_ = half2_to_float2(ny[pr]), anz = half2_to_float2(nz[pr]);
                const float2 afx = half2_to_float2(fx[pr]), afy = half2_to_float2(fy[pr]), afz = half2_to_float2(fz[pr]);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = 2 * pr + e;
                    const float tnx = fmaf(e ? anx.y : anx.x, adjx, nox);
                    const float tny = fmaf(e ? any_.y : any_.x, adjy, noy);
                    const float tnz = fmaf(e ? anz.y : anz.x, adjz, noz);
                    const float tfx = fmaf(e ? afx.y : afx.x, adjx, fox);
                    const float tfy = fmaf(e ? afy.y : afy.x, adjy, foy);
                    const float tfz = fmaf(e ? afz.y : afz.x, adjz, foz);
                    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
                    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tbest));
                    if (tn <= tf) hitmask |= ((child_bits4 >> (8 * j)) & 0xffu) << ((bit_index4 >> (8 * j)) & 0xffu);
                }
            }
#else
            const uint4 n2 = rt3_ldg(np + 2), n3 = rt3_ldg(np + 3), n4 = rt3_ldg(np + 4);
            const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
            const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
            const uint32_t nx = (oct & 1u) ? lox : hix, fx = (oct & 1u) ? hix : lox;
            const uint32_t ny = (oct & 2u) ? loy : hiy, fy = (oct & 2u) ? hiy : loy;
            const uint32_t nz = (oct & 4u) ? loz : hiz, fz = (oct & 4u) ? hiz : loz;
            const Bytes4 bnx = widen_bytes(nx, sc.magic_h2), bny = widen_bytes(ny, sc.magic_h2), bnz = widen_bytes(nz, sc.magic_h2);
            const Bytes4 bfx = widen_bytes(fx, sc.magic_h2), bfy = widen_bytes(fy, sc.magic_h2), bfz = widen_bytes(fz, sc.magic_h2);
#pragma unroll
            for (int pr = 0; pr < 2; pr++) {  // two children per packed fma
                const int j = 2 * pr;
                const float2 tnx = fma2(byte_to_float(bnx, j, sc.bias_h), byte_to_float(bnx, j + 1, sc.bias_h), adjx, nox);
                const float2 tny = fma2(byte_to_float(bny, j, sc.bias_h), byte_to_float(bny, j + 1, sc.bias_h), adjy, noy);
                const float2 tnz = fma2(byte_to_float(bnz, j, sc.bias_h), byte_to_float(bnz, j + 1, sc.bias_h), adjz, noz);
                const float2 tfx = fma2(byte_to_float(bfx, j, sc.bias_h), byte_to_float(bfx, j + 1, sc.bias_h), adjx, fox);
                const float2 tfy = fma2(byte_to_float(bfy, j, sc.bias_h), byte_to_float(bfy, j + 1, sc.bias_h), adjy, foy);
                const float2 tfz = fma2(byte_to_float(bfz, j, sc.bias_h), byte_to_float(bfz, j + 1, sc.bias_h), adjz, foz);
                const float tn0 = fmaxf(fmaxf(tnx.x, tny.x), fmaxf(tnz.x, tmin)), tf0 = fminf(fminf(tfx.x, tfy.x), fminf(tfz.x, tbest));
                const float tn1 = fmaxf(fmaxf(tnx.y, tny.y), fmaxf(tnz.y, tmin)), tf1 = fminf(fminf(tfx.y, tfy.y), fminf(tfz.y, tbest));
                hitmask |= (tn0 <= tf0) ? child_word(child_bits4, bit_index4, j) : 0u;
                hitmask |= (tn1 <= tf1) ? child_word(child_bits4, bit_index4, j + 1) : 0u;
            }
#endif
        }
        ng = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
        tg = make_uint2(n1.y, hitmask & 0x00ffffffu);
        // warm L1 for what this lane touches next: its first pending primitive and its nearest child
        if (!SINGLE && tg.y != 0u && cur_inst >= 0) prefetch_l1(prims + 3u * (tg.x + (uint32_t)(31 - rt3_clz(tg.y & (0u - tg.y)))));
        if (ng.y & 0xff000000u) {
            const uint32_t nslot = ((uint32_t)(31 - rt3_clz(ng.y)) - 24u) ^ oct;
            prefetch_l1((SINGLE ? sc.tlas_nodes : nodes) + (ng.x + (uint32_t)rt3_popc(ng.y & 0xffu & ((1u << nslot) - 1u))));
        }
    }

#ifndef RT3_EMULATE
    // ---------------------------------------------------------------------------------- primary-ray packets (k_extend_packets)
    // Eight consecutive camera rays — the samples of one pixel in the pixel-major path order — leave the same origin in almost
    // the same direction and walk the same nodes.  Here they walk them ONCE: the eight lanes of a group keep identical
    // traversal state (node groups, stack) and share a node step, lane j testing child j of the wide node against the
    // PACKET — the interval [ilo, ihi] of the group's 1/d per axis — instead of every lane testing all eight children for
    // its own ray.  The packet test is a superset of every ray's own padded slab test (so is the packet's far bound, the
    // largest tbest of the group), the triangles of every leaf reached are tested per lane with the ray's own watertight
    // test, and the closest hit is an order-free choice: the result is the one the per-ray traversal finds, bit for bit.
    // unbounded: bit k set = the group's directions straddle zero on axis k; ilo = 1 / (most negative d), ihi = 1 / (most positive d)
    // there, and the axis only bounds the entry distance.
    __device__ __forceinline__ void node_step_packet(const TravScene& sc, uint32_t gmask, uint32_t gl, uint32_t oct, float3 ilo, float3 ihi, float3 iabs,
                                                     uint32_t unbounded, float ptmin, float ptbest) {
        const uint32_t hits = ng.y;
        const int bit = 31 - __clz((int)hits);
        ng.y &= ~(1u << bit);
        if (ng.y & 0xff000000u) push(sc, ng);
        const uint32_t slot = ((uint32_t)bit - 24u) ^ oct;
        const uint32_t rel = (uint32_t)__popc(hits & 0xffu & ((1u << slot) - 1u));
        const uint4* np = reinterpret_cast<const uint4*>(sc.tlas_nodes + (ng.x + rel));
        const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        const uint32_t m = __byte_perm(n1.z, n1.w, gl) & 0xffu;   // meta byte of child slot gl
        uint32_t word = 0u;
        if (m != 0u) {
            // plane distances from the origin, formed like the per-ray test forms them — (p - o) first, then the quantised offset — so
            // that their rounding is relative to the distance itself; the slack covers the per-ray test's own padding (4.8e-7 of
            // |(p - o) / d| + 255 |scale / d|) and the rounding on both sides, a few times over
            const float ox = __uint_as_float(n0.x) - o.x, oy = __uint_as_float(n0.y) - o.y, oz = __uint_as_float(n0.z) - o.z;
            const float sx = __uint_as_float((n0.w & 0xffu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23), sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
            const float lox = fmaf((float)(__byte_perm(n2.x, n2.y, gl) & 0xffu), sx, ox), loy = fmaf((float)(__byte_perm(n2.z, n2.w, gl) & 0xffu), sy, oy),
                        loz = fmaf((float)(__byte_perm(n3.x, n3.y, gl) & 0xffu), sz, oz);
            const float hix = fmaf((float)(__byte_perm(n3.z, n3.w, gl) & 0xffu), sx, ox), hiy = fmaf((float)(__byte_perm(n4.x, n4.y, gl) & 0xffu), sy, oy),
                        hiz = fmaf((float)(__byte_perm(n4.z, n4.w, gl) & 0xffu), sz, oz);
            float tn = ptmin, tf = ptbest;
            {
                const float e = 3e-6f * (fabsf(ox) + 512.0f * sx) * iabs.x;
                if (!(unbounded & 1u)) {
                    const float pn = (oct & 1u) ? lox : hix, pf = (oct & 1u) ? hix : lox;
                    tn = fmaxf(tn, fminf(pn * ilo.x, pn * ihi.x) - e);
                    tf = fminf(tf, fmaxf(pf * ilo.x, pf * ihi.x) + e);
                } else {   // directions of both signs: the beam [o + t dlo, o + t dhi] reaches a box beside the origin only from some t on
                    const float k = 3e-6f * (fabsf(ox) + 512.0f * sx);
                    if (lox > 0.0f) tn = fmaxf(tn, lox * ihi.x - k * ihi.x);
                    if (hix < 0.0f) tn = fmaxf(tn, hix * ilo.x + k * ilo.x);
                }
            }
            {
                const float e = 3e-6f * (fabsf(oy) + 512.0f * sy) * iabs.y;
                if (!(unbounded & 2u)) {
                    const float pn = (oct & 2u) ? loy : hiy, pf = (oct & 2u) ? hiy : loy;
                    tn = fmaxf(tn, fminf(pn * ilo.y, pn * ihi.y) - e);
                    tf = fminf(tf, fmaxf(pf * ilo.y, pf * ihi.y) + e);
                } else {   // directions of both signs: the beam [o + t dlo, o + t dhi] reaches a box beside the origin only from some t on
                    const float k = 3e-6f * (fabsf(oy) + 512.0f * sy);
                    if (loy > 0.0f) tn = fmaxf(tn, loy * ihi.y - k * ihi.y);
                    if (hiy < 0.0f) tn = fmaxf(tn, hiy * ilo.y + k * ilo.y);
                }
            }
            {
                const float e = 3e-6f * (fabsf(oz) + 512.0f * sz) * iabs.z;
                if (!(unbounded & 4u)) {
                    const float pn = (oct & 4u) ? loz : hiz, pf = (oct & 4u) ? hiz : loz;
                    tn = fmaxf(tn, fminf(pn * ilo.z, pn * ihi.z) - e);
                    tf = fminf(tf, fmaxf(pf * ilo.z, pf * ihi.z) + e);
                } else {   // directions of both signs: the beam [o + t dlo, o + t dhi] reaches a box beside the origin only from some t on
                    const float k = 3e-6f * (fabsf(oz) + 512.0f * sz);
                    if (loz > 0.0f) tn = fmaxf(tn, loz * ihi.z - k * ihi.z);
                    if (hiz < 0.0f) tn = fmaxf(tn, hiz * ilo.z + k * ilo.z);
                }
            }
            if (tn - 1e-5f * fabsf(tn) <= tf + 1e-5f * fabsf(tf)) {
                const uint32_t inner = (m & (m << 1)) & 0x10u;   // 001sssss with sssss >= 24: bits 4 and 3 set
                word = ((m >> 5) & 7u) << ((m ^ (inner ? oct : 0u)) & 0x1fu);
            }
        }
        word |= __shfl_xor_sync(gmask, word, 1);
        word |= __shfl_xor_sync(gmask, word, 2);
        word |= __shfl_xor_sync(gmask, word, 4);
        ng = make_uint2(n1.x, (word & 0xff000000u) | (n0.w >> 24));
        tg = make_uint2(n1.y, word & 0x00ffffffu);
#ifdef RT3_STATS
        c_nodes++;
#endif
    }
#endif

    // returns true when an any-hit ray is finished
    RT3_HD bool prim_step(const TravScene& sc) {
        const int bit = 31 - rt3_clz(tg.y & (0u - tg.y));  // lowest set bit
        tg.y &= tg.y - 1u;
        const uint32_t pi = tg.x + (uint32_t)bit;
        if (!SINGLE && cur_inst < 0) {  // TLAS leaf: enter the instance
            const int inst = (int)rt3_ldg(sc.tlas_order + pi);
            {
                const float4 b0 = rt3_ldg(sc.inst_bsphere + 2 * (size_t)inst);
                if (b0.w >= 0.0f) {
                    const float4 b1 = rt3_ldg(sc.inst_bsphere + 2 * (size_t)inst + 1);
                    float3 c = v3(b0);
                    if (b1.w != 0.0f) {  // two motion keys: the centre moves on a straight line (the transform is linear in the key weight)
                        const float a = fminf(fmaxf((ray_time() - sc.instances[inst].t0) * b1.w, 0.0f), 1.0f);
                        c = add(c, mul(v3(b1), a));
                    }
                    const float3 O = sub(o, c), D = cur_d();   // in the TLAS the current space is the world
                    const float OO = dot(O, O), RR = b0.w * b0.w, bq = dot(O, D), cq = OO - RR;
                    if (cq > 1e-5f * (OO + RR)) {                // origin safely outside
                        const float aq = dot(D, D), disc = bq * bq - aq * cq;
                        if (bq >= 0.0f || disc < -1e-5f * (bq * bq + aq * cq)) return false;   // pointing away, or passing by
#if RT3_BSPHERE_TNEAR
                        // entering the sphere beyond the closest hit so far (the leaf BOX was nearer, its corner is not the object)
                        if (disc > 0.0f && (-bq - sqrtf(disc)) * (1.0f - 1e-5f) > tbest * aq) return false;
#endif
                    }
                }
            }
#ifdef RT3_STATS
            c_entries++;
#endif
            if (sp + 3 > RT3_STACK_SIZE) {  // no room for what has to come back (node group, leaf group, exit sentinel): skip the instance, report it
                rt3_atomic_or(sc.error_flags, 1u);
                return false;
            }
            if (ng.y & 0xff000000u) push(sc, ng);
            if (tg.y) push(sc, tg);
            const InstanceDev* in = sc.instances + inst;
            if (in->identity) {  // exact identity, no motion: the object-space ray IS the world-space ray
                push(sc, make_uint2(0xfffffffeu, 0u));
            } else {
                push(sc, make_uint2(0xffffffffu, 0u));  // sentinel: restore the world-space ray on the way out
                const float2 a = fr_get(FR_WO_XY), b = fr_get(FR_WOZ_WDX), c = fr_get(FR_WD_YZ);
                const float time = ray_time();
                float3 oo, od;
                instance_ray(sc, in, sc.hitgroups[inst].t1, time, v3(a.x, a.y, b.x), v3(b.y, c.x, c.y), oo, od);
                set_space(oo, od, time);
            }
            const BlasDev* bl = sc.blas + in->blas;
            nodes = bl->nodes; prims = bl->prims; ptype = bl->type == PRIM_TRI_MOTION ? (PRIM_TRI_MOTION | (bl->vkeys << 8)) : bl->type;
            cur_inst = in->identity == 2u ? RT3_MERGED_INST : inst;  // 2 = the merged world BLAS pseudo-instance
            ng = make_uint2(0u, 0x80000000u);
            tg = make_uint2(0u, 0u);
            return false;
        }
#ifdef RT3_STATS
        c_prims++;
#endif
        const float4* pr = (SINGLE ? sc.root_prims : prims) + 3u * ((!SINGLE && (ptype & 0xffu) == PRIM_TRI_MOTION) ? (ptype >> 8) * pi : pi);
        const float4 a = rt3_ldg(pr), b = rt3_ldg(pr + 1);
        bool got = false;
        if (SINGLE || ptype == PRIM_TRI || (ptype & 0xffu) == PRIM_TRI_MOTION) {
            float3 q0, q1, q2;
            int id;
            if (SINGLE || ptype == PRIM_TRI) {
                const float4 c = rt3_ldg(pr + 2);
                q0 = v3(a); q1 = v3(b); q2 = v3(c);
                id = (int)rt3_f2u(a.w);
            } else {
                // vertex-key motion (the reference's num_keys GAS, cuda_mesh.h:82-88): lerp the bracketing keys at the ray time
                const uint32_t vk = ptype >> 8;
                const float tc = fminf(fmaxf(ray_time(), 0.0f), 1.0f);
                const float f = tc * (float)(vk - 1u);
                int ki = (int)floorf(f);
                if (ki > (int)vk - 2) ki = (int)vk - 2;
                const float al = f - (float)ki, w = 1.0f - al;
                const float4* k0 = pr + 3 * ki;
                const float4 p0 = rt3_ldg(k0), p1 = rt3_ldg(k0 + 1), p2 = rt3_ldg(k0 + 2), r0 = rt3_ldg(k0 + 3), r1 = rt3_ldg(k0 + 4), r2 = rt3_ldg(k0 + 5);
                q0 = v3(w * p0.x + al * r0.x, w * p0.y + al * r0.y, w * p0.z + al * r0.z);
                q1 = v3(w * p1.x + al * r1.x, w * p1.y + al * r1.y, w * p1.z + al * r1.z);
                q2 = v3(w * p2.x + al * r2.x, w * p2.y + al * r2.y, w * p2.z + al * r2.z);
                id = (int)rt3_f2u(a.w);
            }
            float t, u, v;
            if (test_triangle(o, shear(), q0, q1, q2, t, u, v)) got = accept(sc, t, u, v, id);
        } else if (ptype == PRIM_SPHERE) {
            float ta, tb;
            if (test_sphere(o, cur_d(), v3(a), a.w, ta, tb)) {
                const int prim = (int)rt3_f2u(b.x);
                // first root if it lies in the interval, else the second (cuda/sphere.cu:76-94)
                if (in_range(ta)) got = accept(sc, ta, 0.0f, 0.0f, prim);
                else got = accept(sc, tb, 0.0f, 0.0f, prim);
            }
        } else {
            const float4 c = rt3_ldg(pr + 2);
            float t, u;
            if (test_curve_linear(o, cur_d(), v3(a), a.w, v3(b), b.w, t, u)) got = accept(sc, t, u, 0.0f, (int)rt3_f2u(c.x));
        }
        return ANY_HIT && got;
    }

    // One round of the while-while loop; returns false when the ray is finished.  Every lane of a
    // warp runs the three phases in lock step (pop -> one wide node -> all pending primitives), so
    // lanes that are in the same phase execute it together instead of serialising per action.
    RT3_HD bool step(const TravScene& sc) {
#ifdef RT3_STATS
        c_rounds++;
#endif
        while (tg.y == 0u && !(ng.y & 0xff000000u)) {
            if (sp == 0) return false;
            const uint2 e = st_get(--sp);
            if (!SINGLE && e.y == 0u) {  // sentinel: leave the instance
                if (e.x == 0xffffffffu) restore_world();  // the instance had its own ray space
                nodes = sc.tlas_nodes;
                cur_inst = -1;
            } else if (e.y & 0xff000000u) {
                ng = e;
            } else {
                tg = e;
            }
        }
        if (tg.y == 0u) node_step(sc);
        while (tg.y != 0u) {
            if (prim_step(sc)) return false;
        }
        return true;
    }

#ifndef RT3_EMULATE
    // ---------------------------------------------------------------------------------- warp-cooperative round
    // Same round as step(), but executed by ALL 32 lanes of the warp (finished lanes pass active =
    // false) so that the triangle phase can be shared: ncu showed that after a wide-node step only
    // ~2.7 lanes have triangles pending, some of them several, and the per-lane loop ran 3.5 passes
    // of a ~270-instruction test at <10 % lane utilisation — more issue slots than the node phase.
    // Here every lane posts its pending (owner lane, triangle) pairs to a shared-memory work list,
    // the pairs are dealt out one per lane, each lane tests its pair against the OWNER's ray
    // (fetched with shuffles), and the owners then fold the results of their own pairs through the
    // usual accept() rule — so the outcome is identical to the per-lane loop, pass for pass.
    // Spheres, curves and TLAS leaves (instance entry) keep the per-lane path.
    __device__ __forceinline__ bool step_warp(const TravScene& sc, bool active, uint32_t* s_items, float4* s_res) {
        const uint32_t lane = threadIdx.x & 31u;
        if (active) {
#ifdef RT3_STATS
            c_rounds++;
#endif
            while (tg.y == 0u && !(ng.y & 0xff000000u)) {
                if (sp == 0) { active = false; break; }
                const uint2 e = st_get(--sp);
                if (!SINGLE && e.y == 0u) {  // sentinel: leave the instance
                    if (e.x == 0xffffffffu) restore_world();
                    nodes = sc.tlas_nodes;
                    cur_inst = -1;
                } else if (e.y & 0xff000000u) {
                    ng = e;
                } else {
                    tg = e;
                }
            }
            if (active && tg.y == 0u) node_step(sc);
        }
        // ---- cooperative triangle phase (warp-uniform control flow from here)
        bool tri_lane = active && tg.y != 0u && (SINGLE || (cur_inst >= 0 && ptype == PRIM_TRI));
        const bool other_lane = !SINGLE && active && tg.y != 0u && !tri_lane;   // a TLAS leaf to enter, spheres, curves
#if RT3_TRI_BATCH > 1 || RT3_ENTRY_BATCH > 1
        // Minority phases run at a handful of lanes when every round serves whoever happens to need them.  A lane with
        // pending primitives / a pending entry simply WAITS (it holds its place, does no node work) until enough lanes of
        // the warp want the same phase — or until no lane could use the round for node work anyway.
        const bool can_node = active && tg.y == 0u;   // will pop / step a node next round
        const bool nobody_ready = __ballot_sync(0xffffffffu, can_node) == 0u;
#endif
#if RT3_TRI_BATCH > 1
        if (!SINGLE) {
            const uint32_t pending = __reduce_add_sync(0xffffffffu, tri_lane ? (uint32_t)__popc(tg.y) : 0u);
            if (pending < RT3_TRI_BATCH && !nobody_ready) tri_lane = false;   // keep tg: the pairs wait in the lanes
        }
#endif
        const uint32_t cnt = tri_lane ? (uint32_t)__popc(tg.y) : 0u;
        const uint32_t maxc = __reduce_max_sync(0xffffffffu, cnt);
#ifdef RT3_STATS
        {   // warp-round histogram: [6] rounds, [7] rounds without triangles, [8] rounds with <= 1 per lane, [9] lanes with triangles, [10] triangles, [11] busy lanes
            const uint32_t tl = __popc(__ballot_sync(0xffffffffu, tri_lane)), tt = __reduce_add_sync(0xffffffffu, cnt), bl = __popc(__ballot_sync(0xffffffffu, active));
            if (lane == 0) {
                uint32_t* d = const_cast<uint32_t*>(sc.error_flags);
                atomicAdd(d + 6, 1u); atomicAdd(d + 7, maxc == 0u ? 1u : 0u); atomicAdd(d + 8, maxc == 1u ? 1u : 0u);
                atomicAdd(d + 9, tl); atomicAdd(d + 10, tt); atomicAdd(d + 11, bl);
            }
        }
#endif
#if RT3_COOP_MIN > 0
        const uint32_t warp_total = __reduce_add_sync(0xffffffffu, cnt);
        if (maxc > 1u && warp_total < RT3_COOP_MIN) {  // too few pairs to be worth the redistribution: each lane loops over its own
            if (tri_lane) {
                while (tg.y != 0u) {
                    if (prim_step(sc)) { active = false; break; }
                }
            }
        } else
#endif
        if (maxc == 1u) {  // one triangle per lane at most: nothing to redistribute, test in place
            if (tri_lane) {
#ifdef RT3_STATS
                c_prims++;
#endif
                const float4* pr = (SINGLE ? sc.root_prims : prims) + 3u * (tg.x + (uint32_t)(__ffs((int)tg.y) - 1));
                tg.y = 0u;
                const float4 a = __ldg(pr), b = __ldg(pr + 1), c = __ldg(pr + 2);
                float t, u, v;
                if (test_triangle(o, shear(), v3(a), v3(b), v3(c), t, u, v) && accept(sc, t, u, v, (int)__float_as_uint(a.w)) && ANY_HIT) active = false;
            }
        } else if (maxc > 1u) {
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total != 0u) {
            const uint32_t excl = incl - cnt;
            uint32_t posted = 0;
            if (tri_lane) {  // post (owner, triangle) pairs; what does not fit stays in tg for the next round
                uint32_t m = tg.y, pos = excl;
                while (m != 0u && pos < RT3_COOP_CAP) {
                    const uint32_t b = (uint32_t)(__ffs((int)m) - 1);
                    m &= m - 1u;
                    s_items[pos++] = (lane << 27) | (tg.x + b);
                    posted++;
                }
                tg.y = m;
            }
            __syncwarp();
            const uint32_t n_items = total < RT3_COOP_CAP ? total : RT3_COOP_CAP;
            const uint32_t kpack = inv >> 8;
            const uint64_t pbase = SINGLE ? (uint64_t)sc.root_prims : (uint64_t)prims;  // finished lanes help too: never use their stale members
            for (uint32_t base = 0; base < n_items; base += 32u) {
                const uint32_t g = base + lane;
                const bool have = g < n_items;
                const uint32_t item = have ? s_items[g] : 0u;
                const int owner = (int)(item >> 27);
                const float ox = __shfl_sync(0xffffffffu, o.x, owner), oy = __shfl_sync(0xffffffffu, o.y, owner), oz = __shfl_sync(0xffffffffu, o.z, owner);
                Shear s;
                s.Sx = __shfl_sync(0xffffffffu, Sx, owner); s.Sy = __shfl_sync(0xffffffffu, Sy, owner); s.Sz = __shfl_sync(0xffffffffu, Sz, owner);
                const uint32_t kp = __shfl_sync(0xffffffffu, kpack, owner);
                const float otmin = __shfl_sync(0xffffffffu, tmin, owner), otbest = __shfl_sync(0xffffffffu, tbest, owner);
                const uint64_t op = SINGLE ? pbase : __shfl_sync(0xffffffffu, pbase, owner);
                if (have) {
                    s.kx = (int)(kp & 3u); s.ky = (int)((kp >> 2) & 3u); s.kz = (int)((kp >> 4) & 3u);
                    const float4* pr = reinterpret_cast<const float4*>(op) + 3u * (item & 0x07ffffffu);
                    const float4 a = __ldg(pr), b = __ldg(pr + 1), c = __ldg(pr + 2);
                    float t, u, v;
                    const bool hit = test_triangle(v3(ox, oy, oz), s, v3(a), v3(b), v3(c), t, u, v) && t > otmin && t <= otbest;
                    s_res[g] = make_float4(hit ? t : __int_as_float(0x7fc00000), u, v, a.w);  // NaN marks a miss
                }
            }
            __syncwarp();
            if (tri_lane) {  // owners fold their own results, in posting order
                for (uint32_t k = 0; k < posted; k++) {
#ifdef RT3_STATS
                    c_prims++;
#endif
                    const float4 r = s_res[excl + k];
                    if (r.x == r.x) {
                        if (accept(sc, r.x, r.y, r.z, (int)__float_as_uint(r.w)) && ANY_HIT) { active = false; break; }
                    }
                }
            }
            __syncwarp();
        }
        }
        // ---- everything else that is pending: instance entry, spheres, curves (per lane)
#if RT3_ENTRY_BATCH > 1
        const bool go_other = (uint32_t)__popc(__ballot_sync(0xffffffffu, other_lane)) >= RT3_ENTRY_BATCH || nobody_ready;
#else
        const bool go_other = true;
#endif
        if (other_lane && go_other) {
            while (tg.y != 0u) {
                if (prim_step(sc)) { active = false; break; }
            }
        }
        return active;
    }
#endif

#ifndef RT3_EMULATE
    // ---------------------------------------------------------------------------------- deferred triangle queue
    // Single-level scenes.  Counters (RT3_STATS build, C2 bounce rays): a round leaves triangles on only 4
    // of 25 busy lanes, 10.7 in total, in 9 rounds out of 10 — so a triangle pass per round runs its ~300
    // instructions at a third of the lanes and costs as many issue slots as the wide-node step.  Here the
    // (owner lane, triangle) pairs stay in a per-warp shared-memory queue ACROSS rounds while their
    // owners go on with node work, and one full-width pass tests 32 of them when RT3_DEFER_ITEMS have gathered
    // (or fewer when more than RT3_DEFER_BLOCKED lanes have nothing else left, or a lane could not queue
    // everything); pairs beyond 32 stay queued, per-owner counts in shared memory tell when a lane is drained.
    // The fold is order-free: a candidate is valid under the owner's accept() rule, the owner's new hit
    // is the valid candidate with the smallest (t, primitive id) — which is what accept() yields for any
    // order of the same candidates — found with shared-memory atomicMin in two steps (t, then id among
    // equal t).  A late tbest only means a few more nodes are visited; the result is the same.
    __device__ __forceinline__ bool step_warp_deferred(const TravScene& sc, bool active, uint32_t* s_items, float4* s_res, uint32_t* s_best) {
        static_assert(SINGLE, "deferred queue: single-level kernel only");
        const uint32_t lane = threadIdx.x & 31u;
        if (active && tg.y == 0u) {  // (triangles that did not fit the queue last round are queued first)
#ifdef RT3_STATS
            c_rounds++;
#endif
            while (!(ng.y & 0xff000000u) && sp != 0) ng = st_get(--sp);  // the stack holds node groups only here
            if (ng.y & 0xff000000u) node_step(sc);
        }
        // ---- queue this round's triangles: the few lanes that have some reserve their slots with one shared-
        // memory atomicAdd each (the order of the pairs does not matter to the fold)
        if (active && tg.y != 0u) {
            uint32_t m = tg.y, pos = atomicAdd(&s_items[RT3_QCAP], (uint32_t)__popc(m));
            while (m != 0u && pos < RT3_QCAP) {
                const uint32_t b = (uint32_t)(__ffs((int)m) - 1);
                m &= m - 1u;
                s_items[pos++] = (lane << 27) | (tg.x + b);
                pend++;
            }
            tg.y = m;  // what did not fit waits in the lane (and forces the pass below)
#if RT3_DEFER_PARTIAL
            s_best[64u + lane] = pend;  // pairs of this lane in the queue (the pass counts them down)
#endif
#ifdef RT3_STATS
            if (m != 0u) atomicAdd(const_cast<uint32_t*>(sc.error_flags) + 12, 1u);  // queue-full events
#endif
        }
        __syncwarp();
        const uint32_t qn = s_items[RT3_QCAP] < RT3_QCAP ? s_items[RT3_QCAP] : RT3_QCAP;  // warp-uniform from here
        const bool ready = active && tg.y == 0u && ((ng.y & 0xff000000u) != 0u || sp != 0);
        const uint32_t act_mask = __ballot_sync(0xffffffffu, active), ready_mask = __ballot_sync(0xffffffffu, ready);
        const uint32_t waiting = (uint32_t)(__popc(act_mask) - __popc(ready_mask));  // lanes that wait for the pass (or could not queue)
#ifdef RT3_STATS
        if (lane == 0) {
            uint32_t* d = const_cast<uint32_t*>(sc.error_flags);
            atomicAdd(d + 6, 1u); atomicAdd(d + 11, (uint32_t)__popc(act_mask));
        }
#endif
        if (qn != 0u && (qn >= RT3_DEFER_ITEMS || waiting > RT3_DEFER_BLOCKED || ready_mask == 0u)) {
#ifdef RT3_STATS
            if (lane == 0) { uint32_t* d = const_cast<uint32_t*>(sc.error_flags); atomicAdd(d + 7, 1u); atomicAdd(d + 9, qn); }
#endif
            __syncwarp();
            const uint32_t kpack = inv >> 8;
#if RT3_DEFER_PARTIAL
            const uint32_t qpass = qn < 32u ? qn : 32u;  // ONE full-width pass; the rest stays queued
#else
            const uint32_t qpass = qn;
#endif
            for (uint32_t base = 0; base < qpass; base += 32u) {
                s_best[lane] = 0xffffffffu;       // smallest ordered t per owner
                s_best[32u + lane] = 0xffffffffu; // smallest primitive id among those
                __syncwarp();
                const uint32_t g = base + lane;
                const bool have = g < qn;
                const uint32_t item = have ? s_items[g] : 0u;
                const int owner = (int)(item >> 27);
                const float ox = __shfl_sync(0xffffffffu, o.x, owner), oy = __shfl_sync(0xffffffffu, o.y, owner), oz = __shfl_sync(0xffffffffu, o.z, owner);
                Shear s;
                s.Sx = __shfl_sync(0xffffffffu, Sx, owner); s.Sy = __shfl_sync(0xffffffffu, Sy, owner); s.Sz = __shfl_sync(0xffffffffu, Sz, owner);
                const uint32_t kp = __shfl_sync(0xffffffffu, kpack, owner);
                const float otmin = __shfl_sync(0xffffffffu, tmin, owner), otbest = __shfl_sync(0xffffffffu, tbest, owner);
                const int ohprim = __shfl_sync(0xffffffffu, hprim, owner);
                bool valid = false;
                float t = 0.0f, u = 0.0f, v = 0.0f;
                uint32_t id = 0u, key = 0u;
                if (have) {
#ifdef RT3_STATS
                    c_prims++;
#endif
                    s.kx = (int)(kp & 3u); s.ky = (int)((kp >> 2) & 3u); s.kz = (int)((kp >> 4) & 3u);
                    const float4* pr = sc.root_prims + 3u * (item & 0x07ffffffu);
                    const float4 a = __ldg(pr), b = __ldg(pr + 1), c = __ldg(pr + 2);
                    id = __float_as_uint(a.w);
#if RT3_DEFER_PARTIAL
                    atomicSub(&s_best[64 + owner], 1u);
#endif
                    // the owner's accept() rule
                    valid = test_triangle(v3(ox, oy, oz), s, v3(a), v3(b), v3(c), t, u, v) && t > otmin &&
                            (ohprim < 0 ? t < otbest : (t < otbest || (t == otbest && (int)id < ohprim)));
                    if (valid) {
                        const uint32_t tb = __float_as_uint(t);
                        key = tb ^ ((uint32_t)((int)tb >> 31) | 0x80000000u);  // unsigned order == float order
                        atomicMin(&s_best[owner], key);
                    }
                }
                __syncwarp();
                const bool first = valid && s_best[owner] == key;
                if (first) atomicMin(&s_best[32 + owner], id);
                __syncwarp();
                if (first && s_best[32 + owner] == id) s_res[owner] = make_float4(t, u, v, __uint_as_float(id));
                __syncwarp();
                if (s_best[lane] != 0xffffffffu) {
                    const float4 r = s_res[lane];
                    tbest = r.x; hu = r.y; hv = r.z; hprim = (int)__float_as_uint(r.w);
#if RT3_DEFER_PARTIAL
                    if (ANY_HIT) { tg.y = 0u; ng.y = 0u; sp = 0; }  // done; goes inactive below once none of its pairs is left in the queue
#else
                    if (ANY_HIT) { active = false; tg.y = 0u; }
#endif
                }
                __syncwarp();
            }
#if RT3_DEFER_PARTIAL
            const uint32_t rest = qn - qpass;  // <= 32: move it to the front
            const uint32_t moved = lane < rest ? s_items[32u + lane] : 0u;
            pend = s_best[64u + lane];
            __syncwarp();
            if (lane < rest) s_items[lane] = moved;
            if (lane == 0) s_items[RT3_QCAP] = rest;
#else
            if (lane == 0) s_items[RT3_QCAP] = 0u;
            pend = 0u;
#endif
            __syncwarp();
        }
        if (active && pend == 0u && tg.y == 0u && !(ng.y & 0xff000000u) && sp == 0) active = false;  // nothing left anywhere
        return active;
    }
#endif

    RT3_HD HitRec result(const TravScene& sc) const {
#ifdef RT3_STATS
        rt3_atomic_add(const_cast<uint32_t*>(dbg) + 2, c_nodes);
        rt3_atomic_add(const_cast<uint32_t*>(dbg) + 3, c_prims);
        rt3_atomic_add(const_cast<uint32_t*>(dbg) + 4, c_rounds);
        rt3_atomic_add(const_cast<uint32_t*>(dbg) + 13, c_entries);
        rt3_atomic_add(const_cast<uint32_t*>(dbg) + 5, 1u);
#endif
        HitRec h;
        h.t = hprim >= 0 ? tbest : 0.0f;
        h.u = hu; h.v = hv; h.prim = hprim; h.inst = hprim < 0 ? -1 : (SINGLE ? RT3_MERGED_INST : hinst);
        if (hprim >= 0 && h.inst == RT3_MERGED_INST) { const uint2 m = sc.merged_map[hprim]; h.inst = (int)m.x; h.prim = (int)m.y; }
        return h;
    }
};

}  // namespace rt3
