// rt3_wavefront.cuh — the wavefront stages of the path tracer (sm_100a).
//
// The reference runs the whole path loop inside one OptiX raygen megakernel
// (src/shader/raygen.cu:14-87, one thread per pixel, traversal on RT cores).  Here the same
// computation is staged over SoA queues in HBM:
//   generate (raygen.cu:16-46)  ->  [ extend (optixTraverse, shader_common.h:74-88)
//                                     shade  (closehit_radiance.cu:60-160 + miss.cu:22-35 + RR raygen.cu:58-71)
//                                     connect(traceOcclusion, shader_common.h:109-134) ] x depth
//   -> resolve (raygen.cu:75-86)
// Queue records are 16-byte float4 planes indexed by queue slot (coalesced 128-bit accesses);
// surviving paths are compacted into the next queue with one atomic per warp (ballot + popc).
// A path's radiance is accumulated in bounce order into result[path]; resolve sums the samples
// of a pixel in sample order, so the image is bit-identical to the CPU oracle.
#pragma once
#include "rt3_traverse.cuh"

namespace rt3 {

struct TexDev { const uchar4* px; int32_t w, h, addr, filt; };

struct RayPlanes {           // stride in float4 units: 1 for SoA planes, 3 for the AoS rt3_ray
    const float4* r0;        // {o.xyz, tmin}
    const float4* r1;        // {d.xyz, tmax}
    const float4* r2;        // {time, path id (bits), -, -}
    uint32_t stride;
};

struct TraverseArgs {
    TravScene scene;
    RayPlanes rays;
    const uint32_t* count_ptr;   // number of rays (device-resident queue counter) or null
    uint32_t count;              // used when count_ptr == null
    uint32_t* fetch;             // persistent-kernel work counter (zeroed by the host)
    // outputs
    float4* hit0;                // extend: {t,u,v,prim}; trace: rt3_hit AoS (2 float4 per hit)
    int32_t* hit_inst;           // extend
    const float4* contrib;       // connect: {contribution.xyz, -}
    float4* result;              // connect: per-path radiance
    unsigned long long* stat;    // ray counter to bump by count
    uint32_t faithful;           // connect: reproduce the reference's (Le*0)*last_att NaN propagation on occluded rays (mode 0)
    // split scenes (a large merged world BLAS beside other instances) are traversed in two launches over the same rays:
    // pass 1 = the single-level kernel on the merged BLAS, pass 2 = the general kernel on the TLAS of the remaining
    // instances, seeded with pass 1's result (closest hit: t, ids for the tie rule; occlusion: nothing left to do).
    // pass 1 of a connect launch leaves its verdict in the shadow ray itself (tmax = -infinity: occluded — a value no shadow ray
    // has by itself: tmax = distance - 0.01 may well be negative, or NaN, but not that); 0 = the only pass
    uint32_t pass;
};

enum { TRAV_EXTEND = 0, TRAV_CONNECT = 1, TRAV_TRACE_CLOSEST = 2, TRAV_TRACE_ANY = 3 };

template <int MODE, bool SINGLE>
RT3_HD bool trav_begin(const TraverseArgs& a, uint32_t i, Trav<(MODE == TRAV_CONNECT || MODE == TRAV_TRACE_ANY), SINGLE>& tr) {
    const float4 r0 = rt3_ldcs(&a.rays.r0[(size_t)i * a.rays.stride]);
    const float4 r1 = rt3_ldcs(&a.rays.r1[(size_t)i * a.rays.stride]);
    const float4 r2 = rt3_ldcs(&a.rays.r2[(size_t)i * a.rays.stride]);
    tr.init(a.scene, v3(r0), v3(r1), r0.w, r1.w, r2.x);
    if (!SINGLE && a.pass == 2u) {  // what pass 1 found in the merged BLAS
        if (MODE == TRAV_CONNECT) {
            if (r1.w == rt3_u2f(0xff800000u)) { tr.hprim = 0; return false; }
        } else {
            const float4 h0 = MODE == TRAV_EXTEND ? a.hit0[i] : a.hit0[2 * (size_t)i];
            const int hp = (int)rt3_f2u(h0.w);
            if (hp >= 0) {
                tr.tbest = h0.x; tr.hu = h0.y; tr.hv = h0.z; tr.hprim = hp;
                tr.hinst = MODE == TRAV_EXTEND ? a.hit_inst[i] : (int)rt3_f2u(a.hit0[2 * (size_t)i + 1].x);
                if (MODE == TRAV_TRACE_ANY) return false;
            }
        }
    }
    return true;
}

template <int MODE, bool SINGLE>
RT3_HD void trav_end(const TraverseArgs& a, uint32_t i, const Trav<(MODE == TRAV_CONNECT || MODE == TRAV_TRACE_ANY), SINGLE>& tr) {
    const HitRec h = tr.result(a.scene);
    if (MODE == TRAV_EXTEND) {
        rt3_stcs(&a.hit0[i], make_float4(h.t, h.u, h.v, rt3_u2f((uint32_t)h.prim)));
        rt3_stcs(&a.hit_inst[i], h.inst);
    } else if (MODE == TRAV_CONNECT) {
        if (SINGLE && a.pass == 1u) {  // the epilogue belongs to pass 2; an occluded ray is marked for it
            if (h.prim >= 0) reinterpret_cast<float*>(const_cast<float4*>(&a.rays.r1[(size_t)i * a.rays.stride]))[3] = rt3_u2f(0xff800000u);
            return;
        }
        const uint32_t path = rt3_f2u(a.rays.r2[(size_t)i * a.rays.stride].y);
        const float4 c = rt3_ldcs(&a.contrib[i]);
        if (h.prim < 0) {  // unoccluded: result += radiance * last_attenuation (raygen.cu:59)
            float4 r = a.result[path];
            r.x = r.x + c.x; r.y = r.y + c.y; r.z = r.z + c.z;
            a.result[path] = r;
        } else if (a.faithful) {  // occluded: the reference adds (Le*0)*last_att, which is NaN iff the product is not finite
            const float zx = c.x * 0.0f, zy = c.y * 0.0f, zz = c.z * 0.0f;
            if (zx != 0.0f || zy != 0.0f || zz != 0.0f) {
                float4 r = a.result[path];
                r.x = r.x + zx; r.y = r.y + zy; r.z = r.z + zz;
                a.result[path] = r;
            }
        }
    } else {
        rt3_stcs(&a.hit0[2 * (size_t)i], make_float4(h.t, h.u, h.v, rt3_u2f((uint32_t)h.prim)));
        rt3_stcs(&a.hit0[2 * (size_t)i + 1], make_float4(rt3_u2f((uint32_t)h.inst), 0.0f, 0.0f, 0.0f));
    }
}

#ifdef RT3_EMULATE
template <int MODE, bool SINGLE>
static void k_traverse(TraverseArgs a) {
    const uint32_t n = a.count_ptr ? *a.count_ptr : a.count;
    if (a.stat) *a.stat += n;
    for (uint32_t i = 0; i < n; i++) {
        Trav<(MODE == TRAV_CONNECT || MODE == TRAV_TRACE_ANY), SINGLE> tr;
        uint2 stack_mem[RT3_STACK_SIZE + FR_COUNT];
        tr.stack = stack_mem;
        uint32_t hw = 0;
        if (trav_begin<MODE, SINGLE>(a, i, tr))
            while (tr.step(a.scene)) { if ((uint32_t)tr.sp > hw) hw = (uint32_t)tr.sp; }
        if (hw > *a.scene.max_stack) *a.scene.max_stack = hw;
        trav_end<MODE, SINGLE>(a, i, tr);
    }
}
#else
#ifndef RT3_TRAV_MIN_BLOCKS
#define RT3_TRAV_MIN_BLOCKS 7          // general kernel (TLAS, instances, all primitive types): 72 regs.  Measured (r02f, C3 / C4 Mrays/s): 5 CTAs 630 / 836, 6 670 / 876, 7 712 / 931, 8 (64 regs, 110 B spilled) 679 / 878, 9 -7 %
#endif
#ifndef RT3_TRAV_MIN_BLOCKS_SINGLE
#define RT3_TRAV_MIN_BLOCKS_SINGLE 8   // single-level kernel (merged world BLAS only): 64 regs, no spills (9 CTAs = 56 regs spills 44 B: -5 %)
#endif
#ifndef RT3_REFILL_THRESHOLD
#define RT3_REFILL_THRESHOLD 26
#endif
#ifndef RT3_REFILL_THRESHOLD_GENERAL
#define RT3_REFILL_THRESHOLD_GENERAL 22   // two-level kernel: rounds are longer (entry, per-lane primitives), refilling earlier pays.  r02g, C3 / C4 Mrays/s: 16 713 / 961, 22 736 / 950, 26 727 / 920, 29 701 / 876
#endif
// Persistent threads with dynamic fetch: a warp keeps traversing until fewer than
// RT3_REFILL_THRESHOLD lanes are busy, then refills the idle lanes from the queue with a single
// atomicAdd per warp (warp-aggregated fetch).
template <int MODE, bool SINGLE>
__global__ void __launch_bounds__(RT3_TRAV_THREADS, SINGLE ? RT3_TRAV_MIN_BLOCKS_SINGLE : RT3_TRAV_MIN_BLOCKS) k_traverse(TraverseArgs a) {
    const uint32_t n = a.count_ptr ? *a.count_ptr : a.count;
    if (a.stat && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.stat, (unsigned long long)n);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
#if RT3_COOP
    constexpr bool DEFER = SINGLE && RT3_DEFER;
    __shared__ uint32_t s_items[RT3_TRAV_THREADS / 32][DEFER ? RT3_QCAP + 1 : RT3_COOP_CAP];  // deferred queue: [RT3_QCAP] = its fill count
    __shared__ float4 s_res[RT3_TRAV_THREADS / 32][32];
    __shared__ uint32_t s_best[DEFER ? RT3_TRAV_THREADS / 32 : 1][96];  // [0,32) t, [32,64) id, [64,96) queued pairs per owner
    if (DEFER) s_best[DEFER ? threadIdx.x >> 5 : 0][64u + (threadIdx.x & 31u)] = 0u;
    if (DEFER && (threadIdx.x & 31u) == 0u) s_items[threadIdx.x >> 5][DEFER ? RT3_QCAP : 0] = 0u;
    __syncwarp();
#endif
    Trav<(MODE == TRAV_CONNECT || MODE == TRAV_TRACE_ANY), SINGLE> tr;
    uint2 stack_mem[RT3_STACK_SIZE + FR_COUNT];
    tr.stack = stack_mem;
    bool active = false;
    bool exhausted = false;
    uint32_t my = 0;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !exhausted) {
            const uint32_t nidle = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.fetch, nidle);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!active) {
                const uint32_t cand = base + __popc(idle & lt_mask);
                if (cand < n) {
                    my = cand;
                    active = trav_begin<MODE, SINGLE>(a, my, tr);
                    if (!active) trav_end<MODE, SINGLE>(a, my, tr);   // settled by pass 1
                }
            }
            if (base + nidle >= n) exhausted = true;
        }
        uint32_t busy = __ballot_sync(0xffffffffu, active);
        if (busy == 0u) break;
        const uint32_t threshold = exhausted ? 1u : (SINGLE ? RT3_REFILL_THRESHOLD : RT3_REFILL_THRESHOLD_GENERAL);
        while (__popc(busy) >= threshold) {
#if RT3_COOP
            const bool was = active;
            if constexpr (DEFER) active = tr.step_warp_deferred(a.scene, active, s_items[threadIdx.x >> 5], s_res[threadIdx.x >> 5], s_best[threadIdx.x >> 5]);
            else active = tr.step_warp(a.scene, active, s_items[threadIdx.x >> 5], s_res[threadIdx.x >> 5]);
            if (was && !active) trav_end<MODE, SINGLE>(a, my, tr);
#else
            if (active) {
                if (!tr.step(a.scene)) {
                    trav_end<MODE, SINGLE>(a, my, tr);
                    active = false;
                }
            }
#endif
            busy = __ballot_sync(0xffffffffu, active);
        }
    }
}
#endif

#ifndef RT3_EMULATE
// Extension rays of depth 0 in a single-level scene (or pass 1 of a split one): camera rays, eight consecutive ones per packet
// (Trav::node_step_packet).  One thread per ray, 128 rays per CTA, no persistence: the packets of a launch cost about the same.
__global__ void __launch_bounds__(128, 8) k_extend_packets(TraverseArgs a) {
    const uint32_t n = a.count;
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    if (a.stat && i == 0u) atomicAdd(a.stat, (unsigned long long)n);
    const uint32_t lane = threadIdx.x & 31u, gl = lane & 7u, gmask = 0xffu << (lane & 24u);
    const uint32_t ii = i < n ? i : n - 1u;   // a ragged last packet is filled with copies of the last ray (they are not written)
    Trav<false, true> tr;
    uint2 stack_mem[RT3_STACK_SIZE + FR_COUNT];
    tr.stack = stack_mem;
    trav_begin<TRAV_EXTEND, true>(a, ii, tr);
    const float4 r1 = a.rays.r1[(size_t)ii * a.rays.stride];
    // the packet: shared origin (all camera rays leave the eye; a group whose origins differ falls back to boxes of its own), 1/d interval
    float dlx = r1.x, dhx = r1.x, dly = r1.y, dhy = r1.y, dlz = r1.z, dhz = r1.z;
    bool same_o = true;
#pragma unroll
    for (int sft = 1; sft < 8; sft <<= 1) {
        dlx = fminf(dlx, __shfl_xor_sync(0xffffffffu, dlx, sft)); dhx = fmaxf(dhx, __shfl_xor_sync(0xffffffffu, dhx, sft));
        dly = fminf(dly, __shfl_xor_sync(0xffffffffu, dly, sft)); dhy = fmaxf(dhy, __shfl_xor_sync(0xffffffffu, dhy, sft));
        dlz = fminf(dlz, __shfl_xor_sync(0xffffffffu, dlz, sft)); dhz = fmaxf(dhz, __shfl_xor_sync(0xffffffffu, dhz, sft));
        same_o = same_o && __shfl_xor_sync(0xffffffffu, tr.o.x, sft) == tr.o.x && __shfl_xor_sync(0xffffffffu, tr.o.y, sft) == tr.o.y &&
                 __shfl_xor_sync(0xffffffffu, tr.o.z, sft) == tr.o.z;
    }
    same_o = __all_sync(0xffffffffu, same_o) != 0;   // (checked per warp: one launch kind)
    // per axis: directions of one sign -> the interval of 1/d; of both signs (or zero) -> 1 / the extreme of either sign
    uint32_t unbounded = 0u;
    float3 ilo, ihi, iabs;
    if (dlx > 0.0f || dhx < 0.0f) { ilo.x = 1.0f / clamp_dir(dhx); ihi.x = 1.0f / clamp_dir(dlx); } else { unbounded |= 1u; ilo.x = 1.0f / fminf(dlx, -RT3_DIR_EPS); ihi.x = 1.0f / fmaxf(dhx, RT3_DIR_EPS); }
    if (dly > 0.0f || dhy < 0.0f) { ilo.y = 1.0f / clamp_dir(dhy); ihi.y = 1.0f / clamp_dir(dly); } else { unbounded |= 2u; ilo.y = 1.0f / fminf(dly, -RT3_DIR_EPS); ihi.y = 1.0f / fmaxf(dhy, RT3_DIR_EPS); }
    if (dlz > 0.0f || dhz < 0.0f) { ilo.z = 1.0f / clamp_dir(dhz); ihi.z = 1.0f / clamp_dir(dlz); } else { unbounded |= 4u; ilo.z = 1.0f / fminf(dlz, -RT3_DIR_EPS); ihi.z = 1.0f / fmaxf(dhz, RT3_DIR_EPS); }
    iabs = make_float3(fmaxf(fabsf(ilo.x), fabsf(ihi.x)), fmaxf(fabsf(ilo.y), fabsf(ihi.y)), fmaxf(fabsf(ilo.z), fabsf(ihi.z)));
    const uint32_t oct = (dlx >= 0.0f ? 1u : 0u) | (dly >= 0.0f ? 2u : 0u) | (dlz >= 0.0f ? 4u : 0u);   // the group's front-to-back order
    float ptmin = tr.tmin;   // the packet's near bound: the smallest tmin of the group (camera rays all have the same)
#pragma unroll
    for (int sft = 1; sft < 8; sft <<= 1) ptmin = fminf(ptmin, __shfl_xor_sync(0xffffffffu, ptmin, sft));
    if (!same_o) {   // not a camera wave after all: every lane on its own (kept for safety; launch_subframe only sends depth 0 here)
        while (tr.step(a.scene)) {}
    } else {
        for (;;) {
            while (tr.tg.y != 0u) tr.prim_step(a.scene);
            while (!(tr.ng.y & 0xff000000u) && tr.sp != 0) tr.ng = tr.st_get(--tr.sp);   // the stack holds node groups only
            if (!(tr.ng.y & 0xff000000u)) break;
            float ptbest = tr.tbest;
            ptbest = fmaxf(ptbest, __shfl_xor_sync(gmask, ptbest, 1));
            ptbest = fmaxf(ptbest, __shfl_xor_sync(gmask, ptbest, 2));
            ptbest = fmaxf(ptbest, __shfl_xor_sync(gmask, ptbest, 4));
            tr.node_step_packet(a.scene, gmask, gl, oct, ilo, ihi, iabs, unbounded, ptmin, ptbest);
        }
    }
    if (i < n) trav_end<TRAV_EXTEND, true>(a, i, tr);
}
#endif

// ------------------------------------------------------------------------------------ geometry packing / boxes
RT3_GLOBAL(k_tri_boxes, const float* verts, const int32_t* idx, float4* lo, float4* hi) {
    const uint32_t p = RT3_THREAD_ID();
    if (p >= rt3_n_) return;
    const float3 a = ld3(verts + 3 * (size_t)idx[3 * (size_t)p]), b = ld3(verts + 3 * (size_t)idx[3 * (size_t)p + 1]),
                 c = ld3(verts + 3 * (size_t)idx[3 * (size_t)p + 2]);
    lo[p] = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.0f);
    hi[p] = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.0f);
}
// vertex-key meshes: verts [vkeys][nv][3]; a vertex moves on straight segments between keys, so the
// union of the key boxes bounds the triangle at every time
RT3_GLOBAL(k_tri_boxes_motion, const float* verts, const int32_t* idx, uint32_t vkeys, uint32_t nv, float4* lo, float4* hi) {
    const uint32_t p = RT3_THREAD_ID();
    if (p >= rt3_n_) return;
    float3 mn = v3(3e38f, 3e38f, 3e38f), mx = v3(-3e38f, -3e38f, -3e38f);
    for (uint32_t k = 0; k < vkeys; k++)
        for (int c = 0; c < 3; c++) {
            const float3 a = ld3(verts + 3 * ((size_t)k * nv + (size_t)idx[3 * (size_t)p + c]));
            mn = v3(fminf(mn.x, a.x), fminf(mn.y, a.y), fminf(mn.z, a.z));
            mx = v3(fmaxf(mx.x, a.x), fmaxf(mx.y, a.y), fmaxf(mx.z, a.z));
        }
    lo[p] = make_float4(mn.x, mn.y, mn.z, 0.0f);
    hi[p] = make_float4(mx.x, mx.y, mx.z, 0.0f);
}
RT3_GLOBAL(k_pack_tris_motion, const float* verts, const int32_t* idx, uint32_t vkeys, uint32_t nv, const uint32_t* order, float4* out) {
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const uint32_t p = order[j];
    for (uint32_t k = 0; k < vkeys; k++)
        for (int c = 0; c < 3; c++) {
            const float3 a = ld3(verts + 3 * ((size_t)k * nv + (size_t)idx[3 * (size_t)p + c]));
            out[3 * ((size_t)vkeys * j + k) + c] = make_float4(a.x, a.y, a.z, (k == 0 && c == 0) ? rt3_u2f(p) : 0.0f);
        }
}
RT3_GLOBAL(k_sphere_boxes, const float4* cr, float4* lo, float4* hi) {
    const uint32_t p = RT3_THREAD_ID();
    if (p >= rt3_n_) return;
    const float4 s = cr[p];
    const float r = fabsf(s.w);
    lo[p] = make_float4(s.x - r, s.y - r, s.z - r, 0.0f);
    hi[p] = make_float4(s.x + r, s.y + r, s.z + r, 0.0f);
}
RT3_GLOBAL(k_curve_boxes, const float4* cp, const int32_t* seg, float4* lo, float4* hi) {
    const uint32_t p = RT3_THREAD_ID();
    if (p >= rt3_n_) return;
    const float4 a = cp[seg[p]], b = cp[seg[p] + 1];
    const float ra = fabsf(a.w), rb = fabsf(b.w);
    lo[p] = make_float4(fminf(a.x - ra, b.x - rb), fminf(a.y - ra, b.y - rb), fminf(a.z - ra, b.z - rb), 0.0f);
    hi[p] = make_float4(fmaxf(a.x + ra, b.x + rb), fmaxf(a.y + ra, b.y + rb), fmaxf(a.z + ra, b.z + rb), 0.0f);
}
// primitive records in node-contiguous order: 3 x float4 each, original id kept for the hit record
RT3_GLOBAL(k_pack_tris, const float* verts, const int32_t* idx, const uint32_t* order, float4* out) {
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const uint32_t p = order[j];
    const float3 a = ld3(verts + 3 * (size_t)idx[3 * (size_t)p]), b = ld3(verts + 3 * (size_t)idx[3 * (size_t)p + 1]),
                 c = ld3(verts + 3 * (size_t)idx[3 * (size_t)p + 2]);
    out[3 * (size_t)j] = make_float4(a.x, a.y, a.z, rt3_u2f(p));
    out[3 * (size_t)j + 1] = make_float4(b.x, b.y, b.z, 0.0f);
    out[3 * (size_t)j + 2] = make_float4(c.x, c.y, c.z, 0.0f);
}
RT3_GLOBAL(k_pack_spheres, const float4* cr, const uint32_t* order, float4* out) {
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const uint32_t p = order[j];
    out[3 * (size_t)j] = cr[p];
    out[3 * (size_t)j + 1] = make_float4(rt3_u2f(p), 0.0f, 0.0f, 0.0f);
    out[3 * (size_t)j + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}
RT3_GLOBAL(k_pack_curves, const float4* cp, const int32_t* seg, const uint32_t* order, float4* out) {
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const uint32_t p = order[j];
    out[3 * (size_t)j] = cp[seg[p]];
    out[3 * (size_t)j + 1] = cp[seg[p] + 1];
    out[3 * (size_t)j + 2] = make_float4(rt3_u2f(p), 0.0f, 0.0f, 0.0f);
}
// merged world BLAS: records of the static triangle-mesh instances that are traversed without an instance transform —
// identity instances as they are, the others ("flattened") with their vertices brought to world space once, here:
// x' = ((m0 x + m1 y) + m2 z) + m3 per row, unfused (the oracle rebuilds the same vertices).  Primitive id = merged index.
struct MergedRange { uint32_t first, inst; const float* verts; const int32_t* idx; uint32_t has_xf; float xf[12]; };
RT3_HD float3 merged_vertex(const MergedRange& r, uint32_t p, int corner) {
    const float3 v = ld3(r.verts + 3 * (size_t)r.idx[3 * (size_t)p + corner]);
    if (!r.has_xf) return v;
    const float* m = r.xf;
    return v3(((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3], ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7], ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11]);
}
RT3_HD uint32_t merged_range_of(const MergedRange* ranges, uint32_t nranges, uint32_t g) {
    uint32_t lo = 0, hi = nranges;  // last range with first <= g
    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (ranges[mid].first <= g) lo = mid; else hi = mid; }
    return lo;
}
RT3_GLOBAL(k_merged_boxes, const MergedRange* ranges, uint32_t nranges, float4* lo, float4* hi) {
    const uint32_t g = RT3_THREAD_ID();
    if (g >= rt3_n_) return;
    const MergedRange& r = ranges[merged_range_of(ranges, nranges, g)];
    const float3 a = merged_vertex(r, g - r.first, 0), b = merged_vertex(r, g - r.first, 1), c = merged_vertex(r, g - r.first, 2);
    lo[g] = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.0f);
    hi[g] = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.0f);
}
RT3_GLOBAL(k_pack_merged, const MergedRange* ranges, uint32_t nranges, const uint32_t* order, float4* out, uint2* map) {
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const uint32_t g = order[j];
    const MergedRange& r = ranges[merged_range_of(ranges, nranges, g)];
    const uint32_t p = g - r.first;
    const float3 a = merged_vertex(r, p, 0), b = merged_vertex(r, p, 1), c = merged_vertex(r, p, 2);
    out[3 * (size_t)j] = make_float4(a.x, a.y, a.z, rt3_u2f(g));
    out[3 * (size_t)j + 1] = make_float4(b.x, b.y, b.z, 0.0f);
    out[3 * (size_t)j + 2] = make_float4(c.x, c.y, c.z, 0.0f);
    map[g] = make_uint2(r.inst, p);
}
// decoded (conservative) box of child slot s of a wide node
RT3_HD void node_child_box(const Node8& n, int s, float3& lo, float3& hi) {
    const float sx = rt3_u2f((uint32_t)n.ex << 23), sy = rt3_u2f((uint32_t)n.ey << 23), sz = rt3_u2f((uint32_t)n.ez << 23);
#if RT3_NODE_FP16
    const float2 xl = half2_to_float2(n.q[s >> 2][0][0][s & 3]), xh = half2_to_float2(n.q[s >> 2][0][1][s & 3]);
    const float2 yl = half2_to_float2(n.q[s >> 2][1][0][s & 3]), yh = half2_to_float2(n.q[s >> 2][1][1][s & 3]);
    const float2 zl = half2_to_float2(n.q[s >> 2][2][0][s & 3]), zh = half2_to_float2(n.q[s >> 2][2][1][s & 3]);
    lo = v3(n.px + xl.x * sx, n.py + yl.x * sy, n.pz + zl.x * sz);
    hi = v3(n.px + xh.x * sx, n.py + yh.x * sy, n.pz + zh.x * sz);
#else
    lo = v3(n.px + (float)n.qlo[0][s] * sx, n.py + (float)n.qlo[1][s] * sy, n.pz + (float)n.qlo[2][s] * sz);
    hi = v3(n.px + (float)n.qhi[0][s] * sx, n.py + (float)n.qhi[1][s] * sy, n.pz + (float)n.qhi[2][s] * sz);
#endif
    // one ulp-scale step outward: p + q*scale is rounded here, the traversal evaluates it exactly in t-space
    const float3 e = v3(1e-6f * (fabsf(lo.x) + fabsf(hi.x)), 1e-6f * (fabsf(lo.y) + fabsf(hi.y)), 1e-6f * (fabsf(lo.z) + fabsf(hi.z)));
    lo = sub(lo, e);
    hi = add(hi, e);
}

// world-space box of every instance: union over motion keys of the transformed BLAS boxes.  Instead of
// the 8 corners of the BLAS root box, the (up to 64) grandchild boxes of the BLAS root are transformed:
// for a rotated, roughly round object that is close to its true world-space extent, whereas the box of a
// rotated box is up to sqrt(3) larger per side — and every false TLAS hit costs a ray transform, three
// reciprocals, the shear constants and a wide-node step.
struct BlasBounds { float lo[3], hi[3]; };
struct InstBoxAcc {
    float3 mn, mx;
    Affine st;
    const float* keys;
    int nk;
    bool moving;
    RT3_HD void grow(float3 blo, float3 bhi) {
        for (int k = 0; k < nk; k++) {
            Affine km;
            if (moving) for (int j = 0; j < 12; j++) km.m[j] = keys[12 * k + j];
            for (int c = 0; c < 8; c++) {
                float3 p = v3((c & 1) ? bhi.x : blo.x, (c & 2) ? bhi.y : blo.y, (c & 4) ? bhi.z : blo.z);
                if (moving) p = xform_point(km, p);
                p = xform_point(st, p);
                mn = v3(fminf(mn.x, p.x), fminf(mn.y, p.y), fminf(mn.z, p.z));
                mx = v3(fmaxf(mx.x, p.x), fmaxf(mx.y, p.y), fmaxf(mx.z, p.z));
            }
        }
    }
};
RT3_GLOBAL(k_instance_boxes, const InstanceDev* inst, const float* inst_static, const BlasBounds* bb, const BlasDev* blas, const float* keys, int refine, float4* lo, float4* hi) {
    const uint32_t i = RT3_THREAD_ID();
    if (i >= rt3_n_) return;
    const InstanceDev in = inst[i];
    const BlasBounds b = bb[in.blas];
    InstBoxAcc acc;
    for (int j = 0; j < 12; j++) acc.st.m[j] = inst_static[12 * (size_t)i + j];
    acc.mn = v3(3e38f, 3e38f, 3e38f);
    acc.mx = v3(-3e38f, -3e38f, -3e38f);
    acc.moving = in.nkeys > 0;
    acc.nk = in.nkeys > 0 ? (int)in.nkeys : 1;
    acc.keys = keys + in.key_offset;
    const Node8* nodes = blas[in.blas].nodes;
    if (in.identity || nodes == nullptr || !refine) {
        acc.grow(v3(b.lo[0], b.lo[1], b.lo[2]), v3(b.hi[0], b.hi[1], b.hi[2]));
    } else {
        const Node8 root = nodes[0];
        for (int s = 0; s < 8; s++) {
            if (root.meta[s] == 0) continue;
            float3 clo, chi;
            node_child_box(root, s, clo, chi);
            if ((root.meta[s] & 0x18) == 0x18) {  // internal child: descend one more level
                const Node8 ch = nodes[root.child_base + (uint32_t)rt3_popc((uint32_t)root.imask & ((1u << s) - 1u))];
                for (int s2 = 0; s2 < 8; s2++) {
                    if (ch.meta[s2] == 0) continue;
                    float3 glo, ghi;
                    node_child_box(ch, s2, glo, ghi);
                    // a grandchild box may stick out of its (tighter, exact) parent box by quantisation: clip to it
                    glo = v3(fmaxf(glo.x, clo.x), fmaxf(glo.y, clo.y), fmaxf(glo.z, clo.z));
                    ghi = v3(fminf(ghi.x, chi.x), fminf(ghi.y, chi.y), fminf(ghi.z, chi.z));
                    acc.grow(glo, ghi);
                }
            } else {
                acc.grow(clo, chi);
            }
        }
    }
    const float3 mn = acc.mn, mx = acc.mx;
    // pad by a few ulps: the per-ray inverse transform is not exactly the inverse of these corners
    const float3 pad = v3(1e-5f * (fabsf(mn.x) + fabsf(mx.x)) + 1e-7f, 1e-5f * (fabsf(mn.y) + fabsf(mx.y)) + 1e-7f,
                          1e-5f * (fabsf(mn.z) + fabsf(mx.z)) + 1e-7f);
    lo[i] = make_float4(mn.x - pad.x, mn.y - pad.y, mn.z - pad.z, 0.0f);
    hi[i] = make_float4(mx.x + pad.x, mx.y + pad.y, mx.z + pad.z, 0.0f);
}
RT3_GLOBAL(k_invert_static, const float* xf, InstanceDev* inst) {
    const uint32_t i = RT3_THREAD_ID();
    if (i >= rt3_n_) return;
    Affine a;
    for (int j = 0; j < 12; j++) a.m[j] = xf[12 * (size_t)i + j];
    const Affine r = invert_affine(a);
    for (int j = 0; j < 12; j++) inst[i].inv_static[j] = r.m[j];
}

// ------------------------------------------------------------------------------------ frame parameters
struct FrameParams {
    uint32_t width, height, spl, subframe;
    float eye[3], U[3], V[3], W[3];
    float miss[3];
    int32_t max_depth, accum_mode, mode;  // mode 0 = REFERENCE_FAITHFUL, 1 = CORRECTED (unbiased), 2 = CORRECTED + power light sampler
    const Light* lights;
    uint32_t nlights;
    const float* light_cdf;               // running sum of light_power() over the lights (mode 2)
    const TexDev* tex;
    uint32_t path_base;                   // first path of this chain (a subframe may be issued as two independent halves)
};

struct Queues {
    float4* ray0; float4* ray1; float4* ray2;     // current rays
    float4* st0; float4* st1;                      // {att, seed}, {last_att, depth}
    float4* nray0; float4* nray1; float4* nray2;   // next rays
    float4* nst0; float4* nst1;
    float4* hit0; int32_t* hit_inst;
    float4* sh0; float4* sh1; float4* sh2; float4* sh3;  // shadow rays + contribution
    float4* result;                                // per path
    uint32_t* n_cur;                               // device counters
    uint32_t* n_next;
    uint32_t* n_shadow;
};

// ------------------------------------------------------------------------------------ generate (raygen.cu:16-46)
RT3_GLOBAL(k_generate, FrameParams f, Queues q) {
    const uint32_t slot = RT3_THREAD_ID();  // queue slot of this chain; path id = path_base + slot
    if (slot >= rt3_n_) return;
    const uint32_t p = f.path_base + slot;
    const uint32_t npix = f.width * f.height;
    // path id = pixel * spl + sample: the samples of a pixel sit in adjacent lanes, so a warp traces 32 / spl pixels'
    // worth of near-identical primary rays (and first bounces from neighbouring points): +4 % on C2 against the
    // sample-major order; tiling the pixels on top of it (2x2 ... 8x8) measured 0 %
    const uint32_t pix = p / f.spl, k = p % f.spl;
    (void)npix;
    const uint32_t x = pix % f.width, y = pix / f.width;
    uint32_t seed;
    if (f.mode == 0) {
        seed = tea4(pix, f.subframe);
        for (uint32_t i = 0; i < 2u * k; i++) seed = 1664525u * seed + 1013904223u;  // samples of a launch share one LCG stream (Q8)
    } else {
        seed = tea4(pix, f.subframe * f.spl + k);  // corrected: an independent stream per sample
    }
    const float jx = rnd(seed);
    const float jy = rnd(seed);
    const float path_time = f.mode == 0 ? 0.0f : rnd(seed);  // corrected: one ray time per path
    const float dx = 2.0f * (((float)x + jx) / (float)f.width) - 1.0f;
    const float dy = 2.0f * (((float)y + jy) / (float)f.height) - 1.0f;
    const float3 dir = normalize(add(add(mul(ld3(f.U), dx), mul(ld3(f.V), dy)), ld3(f.W)));
    uint32_t pseed = seed;
    const float time = f.mode == 0 ? rnd(pseed) : path_time;  // faithful: traceRadiance draws the ray time first (shader_common.h:64)
    rt3_stcs(&q.ray0[slot], make_float4(f.eye[0], f.eye[1], f.eye[2], 0.01f));
    rt3_stcs(&q.ray1[slot], make_float4(dir.x, dir.y, dir.z, 1e16f));
    rt3_stcs(&q.ray2[slot], make_float4(time, rt3_u2f(p), 0.0f, 0.0f));
    rt3_stcs(&q.st0[slot], make_float4(1.0f, 1.0f, 1.0f, rt3_u2f(pseed)));
    rt3_stcs(&q.st1[slot], make_float4(1.0f, 1.0f, 1.0f, rt3_u2f(0u)));
    rt3_stcs(&q.result[p], make_float4(0.0f, 0.0f, 0.0f, 0.0f));  // result is indexed by path id (all chains share it)
    if (slot == 0) *q.n_cur = rt3_n_;
}

// ------------------------------------------------------------------------------------ LocalGeometry / LocalShading
struct LocalGeometry {   // subset of cuda/LocalGeometry.h:40-58 that the Lambert closure consumes
    float3 P, N;
    float2 UV;
};

// Texel index of an integer texel coordinate under an address mode (CUDA programming guide, texture fetching):
// wrap = modulo, clamp = nearest edge texel, mirror = reflected every N texels, border = -1 (reads as 0).
RT3_HD int resolve_texel(int i, int n, int mode) {
    if (mode == 0) { int m = i % n; return m < 0 ? m + n : m; }
    if (mode == 1) return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    if (mode == 2) { int m = i % (2 * n); if (m < 0) m += 2 * n; return m < n ? m : 2 * n - 1 - m; }
    return (i >= 0 && i < n) ? i : -1;
}
RT3_HD float3 load_texel(const TexDev& tx, int x, int y) {
    if (x < 0 || y < 0) return v3(0.0f, 0.0f, 0.0f);  // border
    const uchar4 t = rt3_ldg(tx.px + (size_t)y * tx.w + x);
    return v3((float)t.x / 255.0f, (float)t.y / 255.0f, (float)t.z / 255.0f);
}
// tex2D on normalised coordinates, RGBA8 -> [0,1], no sRGB decode (cuda_texture.h:52-74, Q10).  filter 0 = point sampling
// (what the reference's `FilterMode::Linear = 0` really selects, Q9); filter 1 = the hardware's bilinear filter (what
// its `FilterMode::Point = 1` selects): texel centres at i + 0.5, weights rounded to 8 fractional bits.
RT3_HD float3 fetch_texture(const TexDev& tx, float u, float v) {
    if (tx.filt == 0 && tx.addr <= 1) {
        int x, y;
        if (tx.addr == 0) {
            const float fu = u - floorf(u), fv = v - floorf(v);
            x = (int)(fu * (float)tx.w);
            y = (int)(fv * (float)tx.h);
        } else {
            x = (int)(fminf(fmaxf(u, 0.0f), 1.0f) * (float)tx.w);
            y = (int)(fminf(fmaxf(v, 0.0f), 1.0f) * (float)tx.h);
        }
        x = x > tx.w - 1 ? tx.w - 1 : x;
        y = y > tx.h - 1 ? tx.h - 1 : y;
        return load_texel(tx, x, y);
    }
    const float fx = u * (float)tx.w, fy = v * (float)tx.h;
    if (tx.filt == 0) return load_texel(tx, resolve_texel((int)floorf(fx), tx.w, tx.addr), resolve_texel((int)floorf(fy), tx.h, tx.addr));
    const float bx = fx - 0.5f, by = fy - 0.5f;
    const float ix = floorf(bx), iy = floorf(by);
    const float al = floorf((bx - ix) * 256.0f + 0.5f) / 256.0f, be = floorf((by - iy) * 256.0f + 0.5f) / 256.0f;
    const int x0 = resolve_texel((int)ix, tx.w, tx.addr), x1 = resolve_texel((int)ix + 1, tx.w, tx.addr);
    const int y0 = resolve_texel((int)iy, tx.h, tx.addr), y1 = resolve_texel((int)iy + 1, tx.h, tx.addr);
    const float3 t00 = load_texel(tx, x0, y0), t10 = load_texel(tx, x1, y0), t01 = load_texel(tx, x0, y1), t11 = load_texel(tx, x1, y1);
    const float w00 = (1.0f - al) * (1.0f - be), w10 = al * (1.0f - be), w01 = (1.0f - al) * be, w11 = al * be;
    return v3(((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x, ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y,
              ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z);
}

// sampleTexture (cuda/LocalShading.h:37-54): the hit group's texcoord transform, then tex2D
RT3_HD float3 sample_texture(const TexDev& tx, const HitGroupDev& hg, float2 uv) {
    if (hg.has_xf == 0u) return fetch_texture(tx, uv.x, uv.y);
    const float sx = uv.x * hg.tex_scale[0], sy = uv.y * hg.tex_scale[1];
    const float tu = (sx * hg.tex_rot[1] + sy * hg.tex_rot[0]) + hg.tex_off[0];       // dot(UV, (rot.y,  rot.x)) + offset.x
    const float tv = (sx * (-hg.tex_rot[0]) + sy * hg.tex_rot[1]) + hg.tex_off[1];    // dot(UV, (-rot.x, rot.y)) + offset.y
    return fetch_texture(tx, tu, tv);
}

// Surface normal (not normalised) of a quadratic / cubic round curve segment given by its power-basis coefficients
// c[0] u^3 + c[1] u^2 + c[2] u + c[3] (xyz + radius; c[0] = 0 for quadratics), at parameter u, for a point ps near the offset
// surface — cuda/curve.h:311-379 with type = 2 (the bona fide normal): flat end caps at u = 0 / 1 (-+ velocity; the cubic
// interpolator evaluates its velocity a hair inside, :281-288), else ps is projected onto the plane through the curve point
// orthogonal to the tangent, dropped onto the surface, and the normal corrected for the radius derivative and the curvature.
RT3_HD float3 spline_surface_normal(const float4* c, bool cubic, float u, float3 ps) {
    const float4 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    if (u == 0.0f || u == 1.0f) {
        const float ue = cubic ? (u == 0.0f ? 0.000001f : 0.999999f) : u;
        const float3 vel = v3((3.0f * c0.x * ue + 2.0f * c1.x) * ue + c2.x, (3.0f * c0.y * ue + 2.0f * c1.y) * ue + c2.y, (3.0f * c0.z * ue + 2.0f * c1.z) * ue + c2.z);
        return u == 0.0f ? neg(vel) : vel;
    }
    const float3 p = v3(((c0.x * u + c1.x) * u + c2.x) * u + c3.x, ((c0.y * u + c1.y) * u + c2.y) * u + c3.y, ((c0.z * u + c1.z) * u + c2.z) * u + c3.z);
    const float r = ((c0.w * u + c1.w) * u + c2.w) * u + c3.w;
    const float3 d = v3((3.0f * c0.x * u + 2.0f * c1.x) * u + c2.x, (3.0f * c0.y * u + 2.0f * c1.y) * u + c2.y, (3.0f * c0.z * u + 2.0f * c1.z) * u + c2.z);
    const float dr = (3.0f * c0.w * u + 2.0f * c1.w) * u + c2.w;
    const float3 acc = v3(6.0f * c0.x * u + 2.0f * c1.x, 6.0f * c0.y * u + 2.0f * c1.y, 6.0f * c0.z * u + 2.0f * c1.z);
    float dd = dot(d, d);
    float3 o1 = sub(ps, p);
    o1 = sub(o1, mul(d, dot(o1, d) / dd));
    o1 = mul(o1, r / length(o1));
    dd -= dot(acc, o1);
    return sub(mul(o1, dd), mul(d, dr * r));
}

// the three object-space vertices of a triangle at a ray time (vertex keys spread evenly over [0, 1], cuda_mesh.h:82-88), as in the traversal
RT3_HD void triangle_vertices(const BlasDev* b, int i0, int i1, int i2, float time, float3& P0, float3& P1, float3& P2) {
    if (b->vkeys <= 1u) {
        P0 = ld3(b->verts + 3 * (size_t)i0); P1 = ld3(b->verts + 3 * (size_t)i1); P2 = ld3(b->verts + 3 * (size_t)i2);
        return;
    }
    const float tc = fminf(fmaxf(time, 0.0f), 1.0f);
    const float f = tc * (float)(b->vkeys - 1u);
    int ki = (int)floorf(f);
    if (ki > (int)b->vkeys - 2) ki = (int)b->vkeys - 2;
    const float al = f - (float)ki, w = 1.0f - al;
    const float* k0 = b->verts + 3 * (size_t)ki * b->nv;
    const float* k1 = k0 + 3 * (size_t)b->nv;
    const float3 a0 = ld3(k0 + 3 * (size_t)i0), a1 = ld3(k0 + 3 * (size_t)i1), a2 = ld3(k0 + 3 * (size_t)i2);
    const float3 c0 = ld3(k1 + 3 * (size_t)i0), c1 = ld3(k1 + 3 * (size_t)i1), c2 = ld3(k1 + 3 * (size_t)i2);
    P0 = v3(w * a0.x + al * c0.x, w * a0.y + al * c0.y, w * a0.z + al * c0.z);
    P1 = v3(w * a1.x + al * c1.x, w * a1.y + al * c1.y, w * a1.z + al * c1.z);
    P2 = v3(w * a2.x + al * c2.x, w * a2.y + al * c2.y, w * a2.z + al * c2.z);
}

RT3_HD LocalGeometry local_geometry(const TravScene& sc, const HitRec& h, float3 o, float3 d, float time) {
    LocalGeometry lg;
    const InstanceDev* in = sc.instances + h.inst;
    const BlasDev* b = sc.blas + in->blas;
    const float t1 = sc.hitgroups[h.inst].t1;
    float3 n_obj;
    if (b->type == PRIM_TRI || b->type == PRIM_TRI_MOTION) {  // closehit_radiance.cu:66-73 (key-0 normals / uvs: create_sbt binds the buffer start)
        const int i0 = b->idx[3 * (size_t)h.prim], i1 = b->idx[3 * (size_t)h.prim + 1], i2 = b->idx[3 * (size_t)h.prim + 2];
        const float w0 = 1.0f - h.u - h.v;
        if (b->normals) n_obj = add(add(mul(ld3(b->normals + 3 * (size_t)i0), w0), mul(ld3(b->normals + 3 * (size_t)i1), h.u)), mul(ld3(b->normals + 3 * (size_t)i2), h.v));
        else {  // the SDK's fallback (cuda/LocalGeometry.h:120-124): the geometric normal
            float3 P0, P1, P2;
            triangle_vertices(b, i0, i1, i2, time, P0, P1, P2);
            n_obj = cross(sub(P1, P0), sub(P2, P0));
        }
        if (b->uvs) {
            lg.UV.x = w0 * b->uvs[2 * (size_t)i0] + h.u * b->uvs[2 * (size_t)i1] + h.v * b->uvs[2 * (size_t)i2];
            lg.UV.y = w0 * b->uvs[2 * (size_t)i0 + 1] + h.u * b->uvs[2 * (size_t)i1 + 1] + h.v * b->uvs[2 * (size_t)i2 + 1];
        } else lg.UV = make_float2(h.u, h.v);  // LocalGeometry.h:150-152
    } else {
        float3 oo, od;
        instance_ray(sc, in, t1, time, o, d, oo, od);
        float3 ps = add(oo, mul(od, h.t));
        if (b->type == PRIM_SPHERE) {  // cuda/sphere.cu:79: normal = (O + t*D) / radius
            const float4 s = b->cr[h.prim];
            n_obj = divs(sub(ps, v3(s)), s.w);
            lg.UV = make_float2(0.0f, 0.0f);
        } else if (b->sub) {  // spline curves: the SDK's surfaceNormal<> of the TRUE curve at the hit's parameter (cuda/curve.h:311-379)
            const uint32_t sg = b->sub[2 * (size_t)h.prim], kK = b->sub[2 * (size_t)h.prim + 1];
            const float uu = ((float)(kK & 0xffffu) + h.u) / (float)(kK >> 16);   // the hit was found on linear sub-segment k of K of its segment
            n_obj = spline_surface_normal(b->poly + 4 * (size_t)sg, b->curve_cubic != 0u, uu, ps);
            lg.UV = make_float2(uu, 0.0f);
        } else {  // cuda/curve.h:382-425 surfaceNormal<LinearInterpolator>
            const float4 c0 = b->cr[b->seg[h.prim]], c1 = b->cr[b->seg[h.prim] + 1];
            if (h.u == 0.0f) n_obj = sub(ps, v3(c0));
            else if (h.u >= 1.0f) n_obj = sub(ps, add(sub(v3(c1), v3(c0)), v3(c0)));  // end point rebuilt from the interpolator's coefficients, as the SDK does (curve.h:389-394)
            else {
                const float3 dd3 = sub(v3(c1), v3(c0));
                const float dr = c1.w - c0.w;
                const float3 p = add(v3(c0), mul(dd3, h.u));
                const float r = c0.w + h.u * dr;
                const float dd = dot(dd3, dd3);
                float3 o1 = sub(ps, p);
                o1 = sub(o1, mul(dd3, dot(o1, dd3) / dd));
                o1 = mul(o1, r / length(o1));
                n_obj = sub(mul(o1, dd), mul(dd3, dr * r));
            }
            lg.UV = make_float2(h.u, 0.0f);
        }
    }
    float3 n = n_obj;
    if (in->nkeys > 0) {
        const Affine m = lerp_keys(sc.keys + in->key_offset, (int)in->nkeys, in->t0, t1, time);
        n = xform_normal_by_inverse(invert_affine(m), n);
    }
    Affine si;
#pragma unroll
    for (int j = 0; j < 12; j++) si.m[j] = in->inv_static[j];
    n = xform_normal_by_inverse(si, n);
    lg.N = normalize(n);
    lg.P = add(o, mul(d, h.t));
    return lg;
}

// The complete stage record of the SDK (cuda/LocalGeometry.h:40-175, one texcoord set), for rt3_get_local_geometry:
// world-space P from the interpolated object-space vertices, geometric normal Ng, shading normal N, UV and the
// object-space derivatives dpdu, dpdv, dndu, dndv exactly as the SDK forms them (not transformed, LocalGeometry.h:126-160).
// Spheres and curves (left empty by the SDK, LocalGeometry.h:164-167) get P = o + t d, N = Ng = their D7 normal,
// UV as in the shade stage and zero derivatives.
struct LocalGeometryFull { float3 P, N, Ng; float2 UV; float3 dndu, dndv, dpdu, dpdv; float4 color; };

RT3_HD LocalGeometryFull local_geometry_full(const TravScene& sc, const HitRec& h, float3 o, float3 d, float time) {
    LocalGeometryFull lg;
    const float3 z = v3(0.0f, 0.0f, 0.0f);
    lg.P = z; lg.N = z; lg.Ng = z; lg.UV = make_float2(0.0f, 0.0f); lg.dndu = z; lg.dndv = z; lg.dpdu = z; lg.dpdv = z;
    lg.color = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    const InstanceDev* in = sc.instances + h.inst;
    const BlasDev* b = sc.blas + in->blas;
    const float t1 = sc.hitgroups[h.inst].t1;
    Affine si, fs, m, mi;
#pragma unroll
    for (int j = 0; j < 12; j++) { si.m[j] = in->inv_static[j]; fs.m[j] = sc.inst_fwd[12 * (size_t)h.inst + j]; }
    const bool moving = in->nkeys > 0;
    if (moving) { m = lerp_keys(sc.keys + in->key_offset, (int)in->nkeys, in->t0, t1, time); mi = invert_affine(m); }
    if (b->type == PRIM_TRI || b->type == PRIM_TRI_MOTION) {
        const int i0 = b->idx[3 * (size_t)h.prim], i1 = b->idx[3 * (size_t)h.prim + 1], i2 = b->idx[3 * (size_t)h.prim + 2];
        float3 P0, P1, P2;
        triangle_vertices(b, i0, i1, i2, time, P0, P1, P2);
        const float w0 = 1.0f - h.u - h.v;
        float3 p = add(add(mul(P0, w0), mul(P1, h.u)), mul(P2, h.v));
        if (moving) p = xform_point(m, p);
        lg.P = xform_point(fs, p);
        if (b->colors) {  // interpolated vertex colours, LocalGeometry.h:99-106
            const float* c0 = b->colors + 4 * (size_t)i0, *c1 = b->colors + 4 * (size_t)i1, *c2 = b->colors + 4 * (size_t)i2;
            lg.color = make_float4(w0 * c0[0] + h.u * c1[0] + h.v * c2[0], w0 * c0[1] + h.u * c1[1] + h.v * c2[1], w0 * c0[2] + h.u * c1[2] + h.v * c2[2],
                                   w0 * c0[3] + h.u * c1[3] + h.v * c2[3]);
        }
        float3 ng = cross(sub(P1, P0), sub(P2, P0));
        if (moving) ng = xform_normal_by_inverse(mi, ng);
        lg.Ng = normalize(xform_normal_by_inverse(si, ng));
        float3 N0, N1, N2;
        if (b->normals) {
            N0 = ld3(b->normals + 3 * (size_t)i0); N1 = ld3(b->normals + 3 * (size_t)i1); N2 = ld3(b->normals + 3 * (size_t)i2);
            float3 n = add(add(mul(N0, w0), mul(N1, h.u)), mul(N2, h.v));
            if (moving) n = xform_normal_by_inverse(mi, n);
            lg.N = normalize(xform_normal_by_inverse(si, n));
        } else {  // no vertex normals: the unit world-space geometric normal stands in for all three (LocalGeometry.h:120-124)
            lg.N = lg.Ng; N0 = lg.Ng; N1 = lg.Ng; N2 = lg.Ng;
        }
        const float3 dp1 = sub(P0, P2), dp2 = sub(P1, P2), dn1 = sub(N0, N2), dn2 = sub(N1, N2);
        if (b->uvs) {
            const float u0x = b->uvs[2 * (size_t)i0], u0y = b->uvs[2 * (size_t)i0 + 1], u1x = b->uvs[2 * (size_t)i1], u1y = b->uvs[2 * (size_t)i1 + 1],
                        u2x = b->uvs[2 * (size_t)i2], u2y = b->uvs[2 * (size_t)i2 + 1];
            lg.UV.x = w0 * u0x + h.u * u1x + h.v * u2x;
            lg.UV.y = w0 * u0y + h.u * u1y + h.v * u2y;
            const float du1 = u0x - u2x, du2 = u1x - u2x, dv1 = u0y - u2y, dv2 = u1y - u2y;
            const float det = du1 * dv2 - dv1 * du2;
            const float invdet = 1.0f / det;
            lg.dpdu = mul(sub(mul(dp1, dv2), mul(dp2, dv1)), invdet);
            lg.dpdv = mul(add(mul(dp1, -du2), mul(dp2, du1)), invdet);
            lg.dndu = mul(sub(mul(dn1, dv2), mul(dn2, dv1)), invdet);
            lg.dndv = mul(add(mul(dn1, -du2), mul(dn2, du1)), invdet);
        } else {  // no texcoords: the barycentrics and edge differences (LocalGeometry.h:150-158)
            lg.UV = make_float2(h.u, h.v);
            lg.dpdu = neg(dp1);
            lg.dpdv = add(neg(dp1), dp2);
            lg.dndu = neg(dn1);
            lg.dndv = add(neg(dn1), dn2);
        }
    } else {
        const LocalGeometry g = local_geometry(sc, h, o, d, time);
        lg.P = g.P; lg.N = g.N; lg.Ng = g.N; lg.UV = g.UV;
    }
    return lg;
}

// rays [n] x {o,tmin | d,tmax | time,...}, hits [n] x {t,u,v,prim | inst,...} -> 27 floats per record (rt3_local_geometry); misses: zeros
RT3_GLOBAL(k_local_geometry, TravScene sc, const float4* rays, const float4* hits, float* out) {
    const uint32_t i = RT3_THREAD_ID();
    if (i >= rt3_n_) return;
    float* r = out + 27 * (size_t)i;
    const float4 h0 = hits[2 * (size_t)i], h1 = hits[2 * (size_t)i + 1];
    HitRec h;
    h.t = h0.x; h.u = h0.y; h.v = h0.z; h.prim = (int)rt3_f2u(h0.w); h.inst = (int)rt3_f2u(h1.x);
    if (h.prim < 0) { for (int k = 0; k < 27; k++) r[k] = 0.0f; return; }
    const float4 r0 = rays[3 * (size_t)i], r1 = rays[3 * (size_t)i + 1], r2 = rays[3 * (size_t)i + 2];
    const LocalGeometryFull g = local_geometry_full(sc, h, v3(r0), v3(r1), r2.x);
    r[0] = g.P.x; r[1] = g.P.y; r[2] = g.P.z; r[3] = g.N.x; r[4] = g.N.y; r[5] = g.N.z; r[6] = g.Ng.x; r[7] = g.Ng.y; r[8] = g.Ng.z;
    r[9] = g.UV.x; r[10] = g.UV.y;
    r[11] = g.dndu.x; r[12] = g.dndu.y; r[13] = g.dndu.z; r[14] = g.dndv.x; r[15] = g.dndv.y; r[16] = g.dndv.z;
    r[17] = g.dpdu.x; r[18] = g.dpdu.y; r[19] = g.dpdu.z; r[20] = g.dpdv.x; r[21] = g.dpdv.y; r[22] = g.dpdv.z;
    r[23] = g.color.x; r[24] = g.color.y; r[25] = g.color.z; r[26] = g.color.w;
}

// rt3_trace output for spline curves: the traversal reports linear sub-segments; the caller sees
// prim = segment, u = (k + u_sub) / K for sub-segment k of the segment's K
RT3_GLOBAL(k_curve_hits_to_user, TravScene sc, float4* hits) {
    const uint32_t i = RT3_THREAD_ID();
    if (i >= rt3_n_) return;
    float4 h0 = hits[2 * (size_t)i];
    const int prim = (int)rt3_f2u(h0.w);
    if (prim < 0) return;
    const int inst = (int)rt3_f2u(hits[2 * (size_t)i + 1].x);
    const uint32_t* sub = sc.blas[sc.instances[inst].blas].sub;
    if (!sub) return;
    const uint32_t kK = sub[2 * (size_t)prim + 1];
    h0.y = ((float)(kK & 0xffffu) + h0.y) / (float)(kK >> 16);
    h0.w = rt3_u2f(sub[2 * (size_t)prim]);
    hits[2 * (size_t)i] = h0;
}

// one atomic per warp: ballot + popc prefix (device); sequential counter (emulation)
RT3_HD uint32_t warp_append(uint32_t* counter, bool pred) {
#ifdef RT3_EMULATE
    return pred ? rt3_atomic_add(counter, 1u) : 0xffffffffu;
#else
    const uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs((int)m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + (uint32_t)__popc(m & ((1u << lane) - 1u)) : 0xffffffffu;
#endif
}

// ------------------------------------------------------------------------------------ shade
// closest-hit / miss programs + the tail of the path loop, for queue slot i (valid == false: the
// lane only takes part in the warp-wide compaction votes)
RT3_HD void shade_slot(const FrameParams& f, const TravScene& sc, const Queues& q, uint32_t i, bool valid) {
    bool push_ray = false, push_shadow = false;
    float3 P = v3(0, 0, 0), ndir = v3(0, 0, 0), L = v3(0, 0, 0), att = v3(0, 0, 0), new_last = v3(0, 0, 0), contrib = v3(0, 0, 0);
    float shadow_tmax = 0.0f, tshadow = 0.0f, next_time = 0.0f;
    uint32_t pseed = 0, path = 0, depth = 0;
    if (valid) {
        const float4 r0 = rt3_ldcs(&q.ray0[i]), r1 = rt3_ldcs(&q.ray1[i]), r2 = rt3_ldcs(&q.ray2[i]);
        const float4 h0 = rt3_ldcs(&q.hit0[i]);
        const float4 s0 = rt3_ldcs(&q.st0[i]), s1 = rt3_ldcs(&q.st1[i]);
        const float3 org = v3(r0), dir = v3(r1);
        const float time = r2.x;
        path = rt3_f2u(r2.y);
        HitRec h;
        h.t = h0.x; h.u = h0.y; h.v = h0.z; h.prim = (int)rt3_f2u(h0.w); h.inst = rt3_ldcs(&q.hit_inst[i]);
        att = v3(s0);
        pseed = rt3_f2u(s0.w);
        const float3 last_att = v3(s1);
        depth = rt3_f2u(s1.w);
        if (h.prim < 0) {  // __miss__radiance (miss.cu:29-32): radiance = callable(0.01), done
            const float3 c = mul(ld3(f.miss), last_att);
            float4 r = q.result[path];
            r.x = r.x + c.x; r.y = r.y + c.y; r.z = r.z + c.z;
            q.result[path] = r;
        } else {           // __closesthit__radiance
            const HitGroupDev hg = sc.hitgroups[h.inst];
            const LocalGeometry lg = local_geometry(sc, h, org, dir, time);
            const float3 Ns = faceforward(lg.N, neg(dir), lg.N);
            P = lg.P;
            if (depth == 0u) {  // emitted only on the camera ray (closehit_radiance.cu:78-83)
                float4 r = q.result[path];
                r.x = r.x + hg.emission[0]; r.y = r.y + hg.emission[1]; r.z = r.z + hg.emission[2];
                q.result[path] = r;
            }
            uint32_t s = pseed;
            (void)rnd(s);
            (void)rnd(s);  // z1, z2 discarded (Q6)
            const float u1 = rnd(s);
            const float u2 = rnd(s);
            const float3 w_in = sample_cosine_hemisphere(u1, u2);
            const float pdf_prev = (float)((double)w_in.z / 3.14159265358979323846);
            ndir = onb_inverse_transform(Ns, w_in);
            const float bsdf = (float)(1.0 / 3.14159265358979323846);
            const float3 albedo = hg.tex >= 0 ? sample_texture(f.tex[hg.tex], hg, lg.UV) : ld3(hg.diffuse);
            att = mul(att, albedo);
            att = mul(att, bsdf / pdf_prev);
            // next event estimation: uniform light pick (closehit_radiance.cu:10-15,117-156)
            const Light* lt = f.lights + (int)(rnd(s) * (float)f.nlights);
            float3 lpos, lem;
            float pdf_light;
            light_sample(lt, P, s, lpos, lem, pdf_light);
            pdf_light = pdf_light / (float)f.nlights;
            pseed = s;
            const float3 dl = sub(lpos, P);
            const float Ldist = length(dl);
            L = normalize(dl);
            const float nDl = dot(Ns, L);
            if (nDl > 0.0f) {
                tshadow = rnd(s);  // local copy: same value as the RR draw below (Q7)
                const float pdf_scat = (float)((double)fabsf(dot(L, Ns)) / 3.14159265358979323846);
                const float3 weight = mul(albedo, power_heuristic(pdf_light, pdf_scat) * bsdf);
                contrib = mul(mul(lem, weight), last_att);
                shadow_tmax = Ldist - 0.01f;
                push_shadow = true;
            } else {  // weight = 0: the reference still adds (Le*0)*last_att — NaN iff not finite
                const float3 z = mul(mul(lem, 0.0f), last_att);
                if (z.x != 0.0f || z.y != 0.0f || z.z != 0.0f) {
                    float4 r = q.result[path];
                    r.x = r.x + z.x; r.y = r.y + z.y; r.z = r.z + z.z;
                    q.result[path] = r;
                }
            }
            new_last = att;
            // Russian roulette (raygen.cu:62-66) + depth bound (extension)
            const float p = att.x * 0.30f + att.y * 0.59f + att.z * 0.11f;
            if (!(rnd(pseed) > p)) {
                att = divs(att, p);
                depth += 1u;
                if (f.max_depth <= 0 || (int)depth < f.max_depth) {
                    next_time = rnd(pseed);
                    push_ray = true;
                }
            }
        }
    }
    const uint32_t sslot = warp_append(q.n_shadow, push_shadow);
    if (push_shadow) {
        rt3_stcs(&q.sh0[sslot], make_float4(P.x, P.y, P.z, 0.001f));
        rt3_stcs(&q.sh1[sslot], make_float4(L.x, L.y, L.z, shadow_tmax));
        rt3_stcs(&q.sh2[sslot], make_float4(tshadow, rt3_u2f(path), 0.0f, 0.0f));
        rt3_stcs(&q.sh3[sslot], make_float4(contrib.x, contrib.y, contrib.z, 0.0f));
    }
    const uint32_t rslot = warp_append(q.n_next, push_ray);
    if (push_ray) {
        rt3_stcs(&q.nray0[rslot], make_float4(P.x, P.y, P.z, 0.01f));
        rt3_stcs(&q.nray1[rslot], make_float4(ndir.x, ndir.y, ndir.z, 1e16f));
        rt3_stcs(&q.nray2[rslot], make_float4(next_time, rt3_u2f(path), 0.0f, 0.0f));
        rt3_stcs(&q.nst0[rslot], make_float4(att.x, att.y, att.z, rt3_u2f(pseed)));
        rt3_stcs(&q.nst1[rslot], make_float4(new_last.x, new_last.y, new_last.z, rt3_u2f(depth)));
    }
}

// ------------------------------------------------------------------------------------ shade, CORRECTED mode
// Same stage, unbiased estimator (draw order identical to oracle/rt3o.cpp
// render_pixel_corrected): throughput *= albedo, NEE with solid-angle light pdf and power-heuristic
// MIS against the cosine lobe, BSDF-sampled emitter hits weighted by the complementary heuristic,
// Russian roulette with p = min(luminance, 1), one ray time per path.
// selection weight of the power light sampler (mode 2): luminance (the Russian-roulette weights of raygen.cu:66) x area
RT3_HD float light_power(float3 e, float area) { return (e.x * 0.30f + e.y * 0.59f + e.z * 0.11f) * area; }

RT3_HD void shade_slot_corrected(const FrameParams& f, const TravScene& sc, const Queues& q, uint32_t i, bool valid) {
    bool push_ray = false, push_shadow = false;
    float3 P = v3(0, 0, 0), ndir = v3(0, 0, 0), Ld = v3(0, 0, 0), beta = v3(0, 0, 0), contrib = v3(0, 0, 0);
    float shadow_tmax = 0.0f, time = 0.0f, pdf_prev = 0.0f;
    uint32_t seed = 0, path = 0, depth = 0;
    if (valid) {
        const float4 r0 = rt3_ldcs(&q.ray0[i]), r1 = rt3_ldcs(&q.ray1[i]), r2 = rt3_ldcs(&q.ray2[i]);
        const float4 h0 = rt3_ldcs(&q.hit0[i]);
        const float4 s0 = rt3_ldcs(&q.st0[i]), s1 = rt3_ldcs(&q.st1[i]);
        const float3 org = v3(r0), dir = v3(r1);
        time = r2.x;
        path = rt3_f2u(r2.y);
        pdf_prev = r2.z;
        HitRec h;
        h.t = h0.x; h.u = h0.y; h.v = h0.z; h.prim = (int)rt3_f2u(h0.w); h.inst = rt3_ldcs(&q.hit_inst[i]);
        beta = v3(s0);
        seed = rt3_f2u(s0.w);
        depth = rt3_f2u(s1.w);
        const float inv_pi = (float)(1.0 / 3.14159265358979323846);
        const float light_total = f.mode == 2 ? f.light_cdf[f.nlights - 1u] : 0.0f;
        const bool by_power = light_total > 0.0f;
        if (h.prim < 0) {
            const float3 c = mul(beta, ld3(f.miss));
            float4 r = q.result[path];
            r.x = r.x + c.x; r.y = r.y + c.y; r.z = r.z + c.z;
            q.result[path] = r;
        } else {
            const HitGroupDev hg = sc.hitgroups[h.inst];
            const LocalGeometry lg = local_geometry(sc, h, org, dir, time);
            const float3 Ns = faceforward(lg.N, neg(dir), lg.N);
            P = lg.P;
            if (hg.emission[0] != 0.0f || hg.emission[1] != 0.0f || hg.emission[2] != 0.0f) {
                float wgt = 1.0f;
                // BSDF-sampled emitter hit: weight against the NEE strategy — where NEE can produce this point at all.  The
                // light list holds the OBJECT-space key-0 triangles of emissive meshes (buildLightSampler, src/wavefront.cpp:
                // 257-275, Q15), so only static triangle meshes under an identity instance qualify; an emissive sphere, curve,
                // deforming mesh or transformed instance is reached by BSDF sampling alone and keeps the full weight.
                const BlasDev* b = sc.blas + sc.instances[h.inst].blas;
                if (depth > 0u && b->type == PRIM_TRI && sc.instances[h.inst].identity == 1u) {
                    const float3 v0 = ld3(b->verts + 3 * (size_t)b->idx[3 * (size_t)h.prim]), v1 = ld3(b->verts + 3 * (size_t)b->idx[3 * (size_t)h.prim + 1]),
                                 v2 = ld3(b->verts + 3 * (size_t)b->idx[3 * (size_t)h.prim + 2]);
                    const float3 nrm = cross(sub(v1, v0), sub(v2, v0));
                    const float area = 0.5f * length(nrm);
                    const float cos_l = fabsf(dot(normalize(nrm), dir));
                    const float dist2 = h.t * h.t * dot(dir, dir);
                    const float pdf_light = by_power ? (dist2 / (area * cos_l)) * (light_power(ld3(hg.emission), area) / light_total)
                                                     : dist2 / ((float)f.nlights * area * cos_l);
                    wgt = power_heuristic(pdf_prev, pdf_light);
                }
                const float3 c = mul(mul(beta, ld3(hg.emission)), wgt);
                float4 r = q.result[path];
                r.x = r.x + c.x; r.y = r.y + c.y; r.z = r.z + c.z;
                q.result[path] = r;
            }
            const float3 albedo = hg.tex >= 0 ? sample_texture(f.tex[hg.tex], hg, lg.UV) : ld3(hg.diffuse);
            // next event estimation
            const float xi_l = rnd(seed);
            float p_sel = 0.0f;
            uint32_t lk;
            if (by_power) {  // smallest k with xi < cdf[k]
                const float xi = xi_l * light_total;
                uint32_t lo = 0, hi = f.nlights - 1u;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (xi < f.light_cdf[mid]) hi = mid; else lo = mid + 1u;
                }
                lk = lo;
                p_sel = light_power(ld3(f.lights[lk].emission), f.lights[lk].area) / light_total;
            } else {
                lk = (uint32_t)(int)(xi_l * (float)f.nlights);
            }
            const Light* lt = f.lights + lk;
            const float u = rnd(seed);
            const float v = rnd(seed);
            const float su0 = sqrtf(u);
            const float b0 = 1.0f - su0, b1 = v * su0;
            const float3 lpos = add(add(mul(ld3(lt->v0), b0), mul(ld3(lt->v1), b1)), mul(ld3(lt->v2), 1.0f - b0 - b1));
            const float3 dl = sub(lpos, P);
            const float dist2 = dot(dl, dl);
            if (dist2 > 1e-10f) {
                const float dist = sqrtf(dist2);
                Ld = divs(dl, dist);
                const float cos_s = dot(Ns, Ld);
                const float cos_l = fabsf(dot(Ld, ld3(lt->normal)));
                if (cos_s > 0.0f && cos_l > 0.0f && lt->area > 0.0f) {
                    const float pdf_light = by_power ? (dist2 / (lt->area * cos_l)) * p_sel : dist2 / ((float)f.nlights * lt->area * cos_l);
                    const float pdf_bsdf = cos_s * inv_pi;
                    const float wgt = power_heuristic(pdf_light, pdf_bsdf);
                    contrib = mul(mul(mul(beta, albedo), ld3(lt->emission)), inv_pi * cos_s * wgt / pdf_light);
                    shadow_tmax = dist - 0.01f;
                    push_shadow = true;
                }
            }
            // BSDF sample + Russian roulette + depth bound
            const float u1 = rnd(seed);
            const float u2 = rnd(seed);
            const float3 w_in = sample_cosine_hemisphere(u1, u2);
            if (w_in.z > 0.0f) {
                pdf_prev = w_in.z * inv_pi;
                ndir = onb_inverse_transform(Ns, w_in);
                beta = mul(beta, albedo);
                const float p = fminf(beta.x * 0.30f + beta.y * 0.59f + beta.z * 0.11f, 1.0f);
                if (!(rnd(seed) > p)) {
                    beta = divs(beta, p);
                    depth += 1u;
                    if (f.max_depth <= 0 || (int)depth < f.max_depth) push_ray = true;
                }
            }
        }
    }
    const uint32_t sslot = warp_append(q.n_shadow, push_shadow);
    if (push_shadow) {
        rt3_stcs(&q.sh0[sslot], make_float4(P.x, P.y, P.z, 0.001f));
        rt3_stcs(&q.sh1[sslot], make_float4(Ld.x, Ld.y, Ld.z, shadow_tmax));
        rt3_stcs(&q.sh2[sslot], make_float4(time, rt3_u2f(path), 0.0f, 0.0f));
        rt3_stcs(&q.sh3[sslot], make_float4(contrib.x, contrib.y, contrib.z, 0.0f));
    }
    const uint32_t rslot = warp_append(q.n_next, push_ray);
    if (push_ray) {
        rt3_stcs(&q.nray0[rslot], make_float4(P.x, P.y, P.z, 0.01f));
        rt3_stcs(&q.nray1[rslot], make_float4(ndir.x, ndir.y, ndir.z, 1e16f));
        rt3_stcs(&q.nray2[rslot], make_float4(time, rt3_u2f(path), pdf_prev, 0.0f));
        rt3_stcs(&q.nst0[rslot], make_float4(beta.x, beta.y, beta.z, rt3_u2f(seed)));
        rt3_stcs(&q.nst1[rslot], make_float4(0.0f, 0.0f, 0.0f, rt3_u2f(depth)));
    }
}

#ifndef RT3_SHADE_MIN_BLOCKS
#define RT3_SHADE_MIN_BLOCKS 4   // grid = SMs x this, all CTAs resident (one wave): 64 regs
#endif
#ifdef RT3_EMULATE
static void k_shade(FrameParams f, TravScene sc, Queues q) {
    const uint32_t n = *q.n_cur;
    for (uint32_t i = 0; i < n; i++) { if (f.mode == 0) shade_slot(f, sc, q, i, true); else shade_slot_corrected(f, sc, q, i, true); }
}
#else
__global__ void __launch_bounds__(256, RT3_SHADE_MIN_BLOCKS) k_shade(FrameParams f, TravScene sc, Queues q) {
    const uint32_t n = *q.n_cur;
    const uint32_t n32 = (n + 31u) & ~31u;
    if (f.mode == 0) { for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += gridDim.x * blockDim.x) shade_slot(f, sc, q, i, i < n); }
    else { for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += gridDim.x * blockDim.x) shade_slot_corrected(f, sc, q, i, i < n); }
}
#endif

// ------------------------------------------------------------------------------------ resolve (raygen.cu:75-86)
RT3_GLOBAL(k_resolve, FrameParams f, const float4* result, float4* accum, uchar4* frame) {
    const uint32_t pix = RT3_THREAD_ID();
    if (pix >= rt3_n_) return;
    const uint32_t npix = f.width * f.height;
    float3 res = v3(0.0f, 0.0f, 0.0f);
    (void)npix;
    for (uint32_t k = 0; k < f.spl; k++) res = add(res, v3(result[(size_t)pix * f.spl + k]));
    float3 c = divs(res, (float)f.spl);
    if (f.accum_mode == 0) {
        if (f.subframe > 0u) {
            const float a = 1.0f / (float)(f.subframe + 1u);
            const float3 prev = v3(accum[pix]);
            c = add(prev, mul(sub(c, prev), a));  // lerp(prev, c, a) = prev + a*(c-prev), vec_math.h:515-518
        }
        accum[pix] = make_float4(c.x, c.y, c.z, 1.0f);
        frame[pix] = make_color(c);
    } else {
        const float4 a = accum[pix];  // SUM mode: cleared by rt3_clear_accum / film (re)allocation
        accum[pix] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + 1.0f);
    }
}
// after the multi-GPU SUM reduce: accum = sum / total_subframes, refresh the 8-bit frame
RT3_GLOBAL(k_finalize, float4* accum, uchar4* frame, float inv_n) {
    const uint32_t pix = RT3_THREAD_ID();
    if (pix >= rt3_n_) return;
    const float4 a = accum[pix];
    const float3 c = v3(a.x * inv_n, a.y * inv_n, a.z * inv_n);
    accum[pix] = make_float4(c.x, c.y, c.z, 1.0f);
    frame[pix] = make_color(c);
}

}  // namespace rt3
