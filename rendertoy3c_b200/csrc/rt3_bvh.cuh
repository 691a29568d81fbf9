// rt3_bvh.cuh — GPU builder for the compressed 8-wide BVH used by both BLAS and TLAS.
//
// Replaces optixAccelBuild / optixAccelCompact (reference call sites src/cuda/cuda_mesh.h:92-146,
// src/cuda/cuda_accel.h:118-146; no reference source exists for the builder itself).
//
// Pipeline (all on the GPU, one stream):
//   1. centroid bounds           (atomic min/max on order-preserving uint keys)
//   2. 63-bit Morton codes       (21 bits / axis)
//   3. sort (key,prim)           (hand-written bitonic sort, k_bitonic_stage)
//   4. Karras 2012 radix tree    (one thread per internal node)
//   5. bottom-up refit           (one thread per leaf, atomic arrival counters)
//   6. top-down collapse BVH2 -> BVH8 by largest-surface-area opening (level-synchronous work
//      queue), octant-ordered child slots, leaf children of <= 3 primitives, primitives
//      re-ordered into node-contiguous blocks
//   7. 80-byte compressed nodes: per-node origin + per-axis power-of-two scale, 8-bit child
//      boxes rounded outward (verified in double), after Ylitie/Karras/Laine 2017.
#pragma once
#include "rt3_common.cuh"
#include "rt3_rt.h"

#include <algorithm>
#include <vector>

namespace rt3 {

// ------------------------------------------------------------------------------------ node layout
// RT3_NODE_FP16 = 1 (default): 128-byte, cache-line-aligned node; child boxes are 11-bit integers
//   (0..2047, in units of the per-axis power-of-two scale) stored as fp16, so the decode is ONE
//   full-rate HADD2.F32 per plane on the FMA pipe (the 8-bit layout needs PRMT + FADD, and the ALU
//   pipe is the busier one in the slab test), the grid is 8x finer, and a node never straddles lines.
// RT3_NODE_FP16 = 0: the 80-byte layout of Ylitie/Karras/Laine 2017 (8-bit boxes).
#ifndef RT3_NODE_FP16
#define RT3_NODE_FP16 0
#endif
#if RT3_NODE_FP16
#define RT3_QMAX 2047
struct alignas(128) Node8 {
    float px, py, pz;        // quantisation origin = node box lo
    uint8_t ex, ey, ez;      // biased exponents: scale_k = 2^(e_k - 127)
    uint8_t imask;           // bit s set: slot s holds an internal child
    uint32_t child_base;     // index of the first internal child (children stored contiguously in slot order)
    uint32_t prim_base;      // index of the first primitive of this node's leaf children
    uint8_t meta[8];         // 0 = empty | internal: 001sssss (sssss = 24+slot) | leaf: (unary count)<<5 | offset
    uint16_t q[2][3][2][4];  // [half: slots 0-3 / 4-7][axis][lo,hi][slot in half] fp16 bit patterns
};
static_assert(sizeof(Node8) == 128, "wide node must be one 128-byte line");
RT3_HD uint16_t int_to_half_bits(int n) {  // exact for 0 <= n < 2048
    if (n <= 0) return 0;
    const int e = 31 - rt3_clz((uint32_t)n);
    return (uint16_t)(((uint32_t)(e + 15) << 10) | (((uint32_t)n << (10 - e)) & 0x3ffu));
}
#else
#define RT3_QMAX 255
struct alignas(16) Node8 {
    float px, py, pz;        // quantisation origin = node box lo
    uint8_t ex, ey, ez;      // biased exponents: scale_k = 2^(e_k - 127)
    uint8_t imask;           // bit s set: slot s holds an internal child
    uint32_t child_base;     // index of the first internal child (children stored contiguously in slot order)
    uint32_t prim_base;      // index of the first primitive of this node's leaf children
    uint8_t meta[8];         // 0 = empty | internal: 001sssss (sssss = 24+slot) | leaf: (unary count)<<5 | offset
    uint8_t qlo[3][8];       // [axis][slot]
    uint8_t qhi[3][8];
};
static_assert(sizeof(Node8) == 80, "compressed wide node must be 80 bytes");
#endif

struct Bvh8 {            // one acceleration structure (BLAS or TLAS) in device memory
    Node8* nodes = nullptr;
    uint32_t* prim_order = nullptr;  // node-contiguous order -> original primitive index
    uint32_t num_nodes = 0, num_prims = 0;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};  // root bounds (host copy)
};

#ifndef RT3_ADAPTIVE_MORTON
#define RT3_ADAPTIVE_MORTON 1
#endif

// ------------------------------------------------------------------------------------ helpers
RT3_HD uint32_t float_to_ordered(float f) { const uint32_t u = rt3_f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
RT3_HD float ordered_to_float(uint32_t u) { return rt3_u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
RT3_HD uint64_t expand21(uint32_t v) {  // spread 21 bits to every third bit
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

struct BuildArrays {  // device pointers of one build
    const float4* plo;  // primitive boxes
    const float4* phi;
    uint32_t n;
    uint32_t* bounds;   // 6 ordered uints: centroid lo xyz, hi xyz
    uint64_t* keys;
    uint32_t* vals;     // after sort: sorted position -> primitive
    // BVH2: ids [0,n-1) internal, (n-1)+j leaf j
    float4* nlo;
    float4* nhi;
    int* left;
    int* right;
    int* parent;
    int* first;   // internal: first sorted position of the range
    int* last;
    uint32_t* flags;
    // BVH8 output
    Node8* nodes;
    uint32_t* prim_order;
    uint32_t* counters;  // [0] nodes allocated, [1] prims allocated, [2] out-queue size
    int2* q_in;
    int2* q_out;
    int leaf_max;        // primitives per leaf child, 1 .. RT3_LEAF_MAX
    // SAH-optimal collapse (k_bvh_dp): per BVH2 node the cost of representing its subtree with at most i = 1..7 children of a
    // wide node, and the split that achieves it; null = greedy collapse
    float* dp_cost;      // [2n-1][7]
    uint8_t* dp_split;   // [2n-1][8]: [j-1] = slots given to the left child when the node is spread over j slots (0: j-1 slots are as good)
    uint32_t* dp_flags;  // arrival counters of the bottom-up pass
};

// ------------------------------------------------------------------------------------ kernels
RT3_GLOBAL(k_bvh_bounds, BuildArrays b) {
    const uint32_t i = RT3_THREAD_ID();
    if (i >= rt3_n_) return;
    const float4 lo = b.plo[i], hi = b.phi[i];
    const float cx = (lo.x + hi.x) * 0.5f, cy = (lo.y + hi.y) * 0.5f, cz = (lo.z + hi.z) * 0.5f;
    rt3_atomic_min(&b.bounds[0], float_to_ordered(cx));
    rt3_atomic_min(&b.bounds[1], float_to_ordered(cy));
    rt3_atomic_min(&b.bounds[2], float_to_ordered(cz));
    rt3_atomic_max(&b.bounds[3], float_to_ordered(cx));
    rt3_atomic_max(&b.bounds[4], float_to_ordered(cy));
    rt3_atomic_max(&b.bounds[5], float_to_ordered(cz));
}

RT3_GLOBAL(k_bvh_morton, BuildArrays b) {
    const uint32_t i = RT3_THREAD_ID();
    if (i >= rt3_n_) return;
    const float4 lo = b.plo[i], hi = b.phi[i];
    const float c[3] = {(lo.x + hi.x) * 0.5f, (lo.y + hi.y) * 0.5f, (lo.z + hi.z) * 0.5f};
    uint64_t key = 0;
    uint32_t q[3];
    float ext[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float l = ordered_to_float(b.bounds[k]), h = ordered_to_float(b.bounds[3 + k]);
        ext[k] = h - l;
        float t = ext[k] > 0.0f ? (c[k] - l) / ext[k] : 0.0f;
        t = fminf(fmaxf(t * 2097152.0f, 0.0f), 2097151.0f);
        q[k] = (uint32_t)t;
    }
#if RT3_ADAPTIVE_MORTON
    // Extent-adaptive bit order (after Vinkler et al. 2017, "Extended Morton codes"): instead of the
    // fixed xyz interleave, each key bit halves the axis whose CELL is currently longest, so a flat
    // scene (terrain: 20 x 8 x 20) is not split in y as often as in x and z.  The axis sequence
    // depends only on the scene bounds, so a common key prefix is still one spatial cell.
    int used[3] = {0, 0, 0};
    for (int bit = 0; bit < 63; bit++) {
        int a = -1;
        float best = -1.0f;
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (used[k] < 21 && ext[k] > best) { best = ext[k]; a = k; }
        key = (key << 1) | ((q[a] >> (20 - used[a])) & 1u);
        used[a]++;
        ext[a] *= 0.5f;
    }
#else
#pragma unroll
    for (int k = 0; k < 3; k++) key |= expand21(q[k]) << k;
#endif
    b.keys[i] = key;
    b.vals[i] = i;
}

// common-prefix length of sorted keys i and j (Karras 2012), ties broken by position
RT3_HD int bvh_delta(const uint64_t* RT3_RESTRICT keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], c = keys[j];
    if (a == c) return 64 + rt3_clz((uint32_t)i ^ (uint32_t)j);
    return rt3_clzll(a ^ c);
}

RT3_GLOBAL(k_bvh_hierarchy, BuildArrays b) {
    const int i = (int)RT3_THREAD_ID();
    const int n = (int)b.n;
    if (i >= n - 1) return;
    const uint64_t* keys = b.keys;
    const int d = (bvh_delta(keys, n, i, i + 1) - bvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = bvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (bvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (bvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = bvh_delta(keys, n, i, j);
    int s = 0;
    int div = 2;
    for (;;) {
        const int t = (l + div - 1) / div;
        if (bvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
        div *= 2;
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    b.left[i] = lc;
    b.right[i] = rc;
    b.first[i] = lo;
    b.last[i] = hi;
    b.parent[lc] = i;
    b.parent[rc] = i;
    if (i == 0) b.parent[0] = -1;
}

RT3_GLOBAL(k_bvh_refit, BuildArrays b) {
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const int n = (int)b.n;
    const uint32_t prim = b.vals[j];
    int id = n - 1 + (int)j;
    b.nlo[id] = b.plo[prim];
    b.nhi[id] = b.phi[prim];
    if (n == 1) return;
    int cur = b.parent[id];
    while (cur >= 0) {
        rt3_threadfence();
        if (rt3_atomic_add(&b.flags[cur], 1u) == 0u) return;  // first arrival: the sibling finishes this node
        rt3_threadfence();
        const int lc = b.left[cur], rc = b.right[cur];
#ifdef RT3_EMULATE
        const float4 a0 = b.nlo[lc], a1 = b.nhi[lc], c0 = b.nlo[rc], c1 = b.nhi[rc];
#else
        const float4 a0 = __ldcg(&b.nlo[lc]), a1 = __ldcg(&b.nhi[lc]), c0 = __ldcg(&b.nlo[rc]), c1 = __ldcg(&b.nhi[rc]);
#endif
        b.nlo[cur] = make_float4(fminf(a0.x, c0.x), fminf(a0.y, c0.y), fminf(a0.z, c0.z), 0.0f);
        b.nhi[cur] = make_float4(fmaxf(a1.x, c1.x), fmaxf(a1.y, c1.y), fmaxf(a1.z, c1.z), 0.0f);
        cur = b.parent[cur];
    }
}

RT3_HD int bvh2_count(const BuildArrays& b, int id) { return id >= (int)b.n - 1 ? 1 : (b.last[id] - b.first[id] + 1); }
// the (at most RT3_LEAF_MAX) sorted positions under a small subtree, left to right.  A walk instead of [first, last]: the
// leaves of a PLOC subtree are not a contiguous range of the Morton order
RT3_HD int bvh2_leaves(const BuildArrays& b, int id, int out[4]) {
    const int n = (int)b.n;
    int stack[8], sp = 0, cnt = 0;
    stack[sp++] = id;
    while (sp > 0 && cnt < 4) {
        const int cur = stack[--sp];
        if (cur >= n - 1) { out[cnt++] = cur - (n - 1); continue; }
        if (sp < 7) { stack[sp++] = b.right[cur]; stack[sp++] = b.left[cur]; }
    }
    return cnt;
}
RT3_HD float bvh2_area(const BuildArrays& b, int id) {
    const float4 lo = b.nlo[id], hi = b.nhi[id];
    const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
    return ex * ey + ey * ez + ez * ex;
}

// ------------------------------------------------------------------------------------ PLOC (SAH-quality binary tree)
// Parallel locally-ordered clustering (Meister & Bittner 2018): the clusters — at first the primitives in Morton order —
// each look RT3_PLOC_RADIUS neighbours to either side for the partner that gives the smallest merged surface area; mutual
// nearest neighbours merge into a new node; the array is compacted and the round repeats until one cluster is left.
// Agglomerative by surface area, so the tree approaches a full SAH build where the Morton-order tree (k_bvh_hierarchy)
// splits at key-bit boundaries whatever they cut through.  Everything order dependent goes through prefix sums (no
// atomics): the tree is the same run to run and in the kernel-logic simulator.  Internal node ids descend from n-2 so that
// the last merge is node 0, the root, as the collapse expects; leaf of sorted position j is id n-1+j as before.
#ifndef RT3_PLOC_RADIUS
#define RT3_PLOC_RADIUS 16
#endif
struct PlocArrays {
    const int* in;      // clusters of this round (node ids)
    int* out;           // clusters of the next round
    int* nn;            // nearest neighbour (index into `in`)
    uint32_t* merge;    // 1: cluster i merges with nn[i] and i < nn[i]   -> after the scan: exclusive prefix
    uint32_t* keep;     // 1: cluster i survives (as itself or as the merged node) -> after the scan: exclusive prefix
    uint32_t* sums;     // [2][chunks] chunk totals, then their exclusive prefix
    uint32_t count;     // clusters in `in`
    int next_id;        // the first merge of this round gets next_id - 1
};
RT3_HD float ploc_union_area(const float4 a0, const float4 a1, const float4 c0, const float4 c1) {
    const float ex = fmaxf(a1.x, c1.x) - fminf(a0.x, c0.x), ey = fmaxf(a1.y, c1.y) - fminf(a0.y, c0.y), ez = fmaxf(a1.z, c1.z) - fminf(a0.z, c0.z);
    return ex * ey + ey * ez + ez * ex;
}
RT3_GLOBAL(k_ploc_leaves, BuildArrays b, int* clusters) {   // leaf boxes + the first cluster list
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const int id = (int)b.n - 1 + (int)j;
    const uint32_t prim = b.vals[j];
    b.nlo[id] = b.plo[prim];
    b.nhi[id] = b.phi[prim];
    clusters[j] = id;
}
RT3_GLOBAL(k_ploc_nearest, BuildArrays b, PlocArrays p) {
    const int i = (int)RT3_THREAD_ID();
    if (i >= (int)p.count) return;
    const float4 lo = b.nlo[p.in[i]], hi = b.nhi[p.in[i]];
    const int from = i - RT3_PLOC_RADIUS < 0 ? 0 : i - RT3_PLOC_RADIUS, to = i + RT3_PLOC_RADIUS > (int)p.count - 1 ? (int)p.count - 1 : i + RT3_PLOC_RADIUS;
    // Ties are the rule on tessellated grids (every neighbour of a regular mesh gives the same merged area), and a one-sided
    // tie-break ("the lower index") would chain i -> i-1 -> i-2 ... with ONE mutual pair per round.  The tie-break is therefore
    // symmetric: among equal areas the partner with the smallest i XOR j, i.e. the one sharing the longest index prefix —
    // whatever i prefers about j, j prefers about i, so tied neighbours pair up like the siblings of a binary heap.
    float best = 3.4e38f;
    uint32_t best_x = 0xffffffffu;
    int bj = -1;
    for (int j = from; j <= to; j++) {
        if (j == i) continue;
        const float a = ploc_union_area(lo, hi, b.nlo[p.in[j]], b.nhi[p.in[j]]);
        const uint32_t x = (uint32_t)i ^ (uint32_t)j;
        if (a < best || (a == best && x < best_x)) { best = a; best_x = x; bj = j; }
    }
    p.nn[i] = bj;
}
RT3_GLOBAL(k_ploc_flags, PlocArrays p) {
    const int i = (int)RT3_THREAD_ID();
    if (i >= (int)p.count) return;
    const int j = p.nn[i];
    const bool mutual = j >= 0 && p.nn[j] == i;
    p.merge[i] = (mutual && i < j) ? 1u : 0u;
    p.keep[i] = (mutual && i > j) ? 0u : 1u;
}
// exclusive prefix sums of merge[] and keep[] in three launches of independent threads (chunks of 256)
RT3_GLOBAL(k_ploc_scan_chunks, PlocArrays p, uint32_t chunks) {
    const uint32_t c = RT3_THREAD_ID();
    if (c >= rt3_n_) return;
    uint32_t m = 0, k = 0;
    const uint32_t end = (c + 1) * 256u < p.count ? (c + 1) * 256u : p.count;
    for (uint32_t i = c * 256u; i < end; i++) { m += p.merge[i]; k += p.keep[i]; }
    p.sums[c] = m;
    p.sums[chunks + c] = k;
}
RT3_GLOBAL(k_ploc_scan_sums, PlocArrays p, uint32_t chunks) {   // one thread; <= n / 256 entries
    if (RT3_THREAD_ID() != 0) return;
    uint32_t m = 0, k = 0;
    for (uint32_t c = 0; c < chunks; c++) {
        const uint32_t a = p.sums[c], d = p.sums[chunks + c];
        p.sums[c] = m; p.sums[chunks + c] = k;
        m += a; k += d;
    }
    p.sums[2 * chunks] = m;       // merges of this round
    p.sums[2 * chunks + 1] = k;   // clusters of the next round
}
RT3_GLOBAL(k_ploc_merge, BuildArrays b, PlocArrays p, uint32_t chunks) {   // one thread per chunk: finishes the scan and applies it
    const uint32_t c = RT3_THREAD_ID();
    if (c >= rt3_n_) return;
    uint32_t m = p.sums[c], k = p.sums[chunks + c];
    const uint32_t end = (c + 1) * 256u < p.count ? (c + 1) * 256u : p.count;
    for (uint32_t i = c * 256u; i < end; i++) {
        const uint32_t is_merge = p.merge[i], is_keep = p.keep[i];
        if (is_merge) {
            const int id = p.next_id - 1 - (int)m, l = p.in[i], r = p.in[p.nn[i]];
            const float4 a0 = b.nlo[l], a1 = b.nhi[l], c0 = b.nlo[r], c1 = b.nhi[r];
            b.nlo[id] = make_float4(fminf(a0.x, c0.x), fminf(a0.y, c0.y), fminf(a0.z, c0.z), 0.0f);
            b.nhi[id] = make_float4(fmaxf(a1.x, c1.x), fmaxf(a1.y, c1.y), fmaxf(a1.z, c1.z), 0.0f);
            b.left[id] = l; b.right[id] = r;
            b.parent[l] = id; b.parent[r] = id;
            const int nl = l >= (int)b.n - 1 ? 1 : b.last[l] + 1, nr = r >= (int)b.n - 1 ? 1 : b.last[r] + 1;
            b.first[id] = 0; b.last[id] = nl + nr - 1;   // only the COUNT of a PLOC subtree is meaningful (bvh2_count)
            p.out[k] = id;
        } else if (is_keep) p.out[k] = p.in[i];
        m += is_merge; k += is_keep;
    }
}

// smallest biased exponent E with p + RT3_QMAX * 2^(E-127) >= hi (checked in double)
RT3_HD uint32_t quant_exponent(float p, float hi) {
    const float ext = hi - p;
    uint32_t E = 1;
    if (ext > 0.0f) {
        const uint32_t bits = rt3_f2u(ext / (float)RT3_QMAX);
        E = (bits >> 23) & 0xffu;
        if (bits & 0x7fffffu) E += 1;
        if (E < 1) E = 1;
    }
    while (E < 254 && (double)p + (double)RT3_QMAX * (double)rt3_u2f(E << 23) < (double)hi) E++;
    return E;
}

// ------------------------------------------------------------------------------------ SAH-optimal collapse (dynamic program)
// Which descendants of a binary node become the (at most) 8 children of its wide node decides how many wide nodes a ray
// visits.  After Ylitie, Karras & Laine 2017 (section 3.1): bottom-up over the binary tree, C(n, i) = the least SAH cost of
// representing the subtree of n by at most i child slots —
//     C(n, 1) = min( leaf: A(n) * P(n) * c_prim  if P(n) <= leaf_max,   wide node: D(n, 8) + A(n) * c_node )
//     C(n, i) = min( D(n, i), C(n, i - 1) ),     D(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
// — and the collapse then follows the recorded splits instead of greedily opening the largest child.
#define RT3_DP_CNODE 1.0f
#ifndef RT3_DP_CPRIM
#define RT3_DP_CPRIM 0.6f   // cost of a primitive test relative to a wide-node test.  The paper uses 0.3; measured here (r02k, Mrays/s C2 / C3 / C4): greedy 2581 / 822 / 953, DP 0.3: 2641 / 825 / 979, DP 0.6: 2651 / 836 / 1005
#endif
RT3_HD void bvh_dp_node(const BuildArrays& b, int id) {
    const int n = (int)b.n;
    float* c = b.dp_cost + 7 * (size_t)id;
    uint8_t* sp = b.dp_split + 8 * (size_t)id;
    const float area = bvh2_area(b, id);
    if (id >= n - 1) {  // a primitive
        for (int i = 0; i < 7; i++) c[i] = area * RT3_DP_CPRIM;
        for (int i = 0; i < 8; i++) sp[i] = 0;
        return;
    }
    const float* cl = b.dp_cost + 7 * (size_t)b.left[id];
    const float* cr = b.dp_cost + 7 * (size_t)b.right[id];
    float dist[9];
    uint8_t kbest[9];
    for (int j = 2; j <= 8; j++) {
        float best = 3.4e38f;
        int bk = 1;
        for (int k = 1; k < j; k++) {
            if (k > 7 || j - k > 7) continue;
            const float v = cl[k - 1] + cr[j - k - 1];
            if (v < best) { best = v; bk = k; }
        }
        dist[j] = best;
        kbest[j] = (uint8_t)bk;
    }
    const int cnt = bvh2_count(b, id);
    const float c_leaf = cnt <= b.leaf_max ? area * (float)cnt * RT3_DP_CPRIM : 3.4e38f;
    const float c_node = dist[8] + area * RT3_DP_CNODE;
    c[0] = c_leaf < c_node ? c_leaf : c_node;
    sp[0] = 0;
    for (int i = 2; i <= 7; i++) {
        if (dist[i] < c[i - 2]) { c[i - 1] = dist[i]; sp[i - 1] = kbest[i]; }
        else { c[i - 1] = c[i - 2]; sp[i - 1] = 0; }
    }
    sp[7] = kbest[8];
}
RT3_GLOBAL(k_bvh_dp, BuildArrays b) {   // one thread per primitive, walking up; the second arrival at a node computes it
    const uint32_t j = RT3_THREAD_ID();
    if (j >= rt3_n_) return;
    const int n = (int)b.n;
    int id = n - 1 + (int)j;
    bvh_dp_node(b, id);
    if (n == 1) return;
    int cur = b.parent[id];
    while (cur >= 0) {
        rt3_threadfence();
        if (rt3_atomic_add(&b.dp_flags[cur], 1u) == 0u) return;
        rt3_threadfence();
        bvh_dp_node(b, cur);
        cur = b.parent[cur];
    }
}
// the children of the wide node rooted at binary node `root`, following the recorded splits
RT3_HD int bvh_dp_children(const BuildArrays& b, int root, int ch[8]) {
    const int n = (int)b.n;
    int nch = 0;
    int st_node[16], st_slots[16], sp = 0;
    {   // the root of a wide node is opened by definition: its 8 slots go to its two binary children
        const int k = b.dp_split[8 * (size_t)root + 7];
        st_node[sp] = b.right[root]; st_slots[sp++] = 8 - k;
        st_node[sp] = b.left[root]; st_slots[sp++] = k;
    }
    while (sp > 0) {
        const int m = st_node[--sp];
        int i = st_slots[sp];
        if (m >= n - 1 || i <= 1) { ch[nch++] = m; continue; }
        if (i > 7) i = 7;
        int k = b.dp_split[8 * (size_t)m + (size_t)(i - 1)];
        while (k == 0 && i > 1) { --i; k = i > 1 ? b.dp_split[8 * (size_t)m + (size_t)(i - 1)] : 0; }   // fewer slots are as good
        if (i <= 1) { ch[nch++] = m; continue; }
        st_node[sp] = b.right[m]; st_slots[sp++] = i - k;
        st_node[sp] = b.left[m]; st_slots[sp++] = k;
    }
    return nch;
}

#ifndef RT3_LEAF_MAX
#define RT3_LEAF_MAX 3  // primitives per leaf child (the meta byte holds a unary count of up to 3)
#endif
RT3_GLOBAL(k_bvh_collapse, BuildArrays b) {
    const uint32_t item = RT3_THREAD_ID();
    if (item >= rt3_n_) return;
    const int n = (int)b.n;
    const int root = b.q_in[item].x;
    const uint32_t widx = (uint32_t)b.q_in[item].y;

    int ch[8];
    int nch = 1;
    ch[0] = root;
    if (b.dp_split != nullptr && root < n - 1) nch = bvh_dp_children(b, root, ch);
    else
    // phase 1: open the largest-area child holding more than 3 primitives; phase 2: use free slots
    // to split the remaining multi-primitive leaves (tighter boxes at no extra nodes)
    for (int phase = 0; phase < 2; phase++) {
        const int min_count = phase == 0 ? b.leaf_max + 1 : 2;
        while (nch < 8) {
            int best = -1;
            float best_a = -1.0f;
            for (int c = 0; c < nch; c++) {
                if (ch[c] >= n - 1) continue;  // BVH2 leaf
                if (bvh2_count(b, ch[c]) < min_count) continue;
                const float a = bvh2_area(b, ch[c]);
                if (a > best_a) { best_a = a; best = c; }
            }
            if (best < 0) break;
            const int o = ch[best];
            ch[best] = b.left[o];
            ch[nch++] = b.right[o];
        }
    }

    const float4 rlo = b.nlo[root], rhi = b.nhi[root];
    const float cen[3] = {(rlo.x + rhi.x) * 0.5f, (rlo.y + rhi.y) * 0.5f, (rlo.z + rhi.z) * 0.5f};
    // octant-ordered slot assignment (greedy): slot bit k set <=> child lies on the + side of axis k
    float cd[8][3];
    for (int c = 0; c < nch; c++) {
        const float4 lo = b.nlo[ch[c]], hi = b.nhi[ch[c]];
        cd[c][0] = (lo.x + hi.x) * 0.5f - cen[0];
        cd[c][1] = (lo.y + hi.y) * 0.5f - cen[1];
        cd[c][2] = (lo.z + hi.z) * 0.5f - cen[2];
    }
    int slot_child[8];
    for (int s = 0; s < 8; s++) slot_child[s] = -1;
    uint32_t child_done = 0;
    for (int round = 0; round < nch; round++) {
        float best = -3.4e38f;
        int bc = -1, bs = -1;
        for (int c = 0; c < nch; c++) {
            if (child_done & (1u << c)) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_child[s] >= 0) continue;
                const float cost = ((s & 1) ? cd[c][0] : -cd[c][0]) + ((s & 2) ? cd[c][1] : -cd[c][1]) + ((s & 4) ? cd[c][2] : -cd[c][2]);
                if (cost > best) { best = cost; bc = c; bs = s; }
            }
        }
        slot_child[bs] = bc;
        child_done |= 1u << bc;
    }

    uint32_t imask = 0, nprims = 0;
    for (int s = 0; s < 8; s++) {
        if (slot_child[s] < 0) continue;
        const int cnt = bvh2_count(b, ch[slot_child[s]]);
        if (cnt > b.leaf_max) imask |= 1u << s;
        else nprims += (uint32_t)cnt;
    }
    const uint32_t nint = (uint32_t)rt3_popc(imask);
    const uint32_t child_base = nint ? rt3_atomic_add(&b.counters[0], nint) : 0u;
    const uint32_t prim_base = nprims ? rt3_atomic_add(&b.counters[1], nprims) : 0u;

    Node8 nd;
    nd.px = rlo.x; nd.py = rlo.y; nd.pz = rlo.z;
    const uint32_t E[3] = {quant_exponent(rlo.x, rhi.x), quant_exponent(rlo.y, rhi.y), quant_exponent(rlo.z, rhi.z)};
    nd.ex = (uint8_t)E[0]; nd.ey = (uint8_t)E[1]; nd.ez = (uint8_t)E[2];
    nd.imask = (uint8_t)imask;
    nd.child_base = child_base;
    nd.prim_base = prim_base;
    const float p[3] = {rlo.x, rlo.y, rlo.z};
    uint32_t offset = 0;
    for (int s = 0; s < 8; s++) {
        const int c = slot_child[s];
        if (c < 0) {
            nd.meta[s] = 0;
            for (int k = 0; k < 3; k++) {  // inverted box; the empty meta byte is what makes it unhittable
#if RT3_NODE_FP16
                nd.q[s >> 2][k][0][s & 3] = int_to_half_bits(RT3_QMAX);
                nd.q[s >> 2][k][1][s & 3] = 0;
#else
                nd.qlo[k][s] = 255; nd.qhi[k][s] = 0;
#endif
            }
            continue;
        }
        const int id = ch[c];
        const int cnt = bvh2_count(b, id);
        if (cnt > b.leaf_max) {
            nd.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            const uint32_t cw = child_base + (uint32_t)rt3_popc(imask & ((1u << s) - 1u));
            const uint32_t q = rt3_atomic_add(&b.counters[2], 1u);
            b.q_out[q] = make_int2(id, (int)cw);
        } else {
            const uint32_t unary = cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u);
            nd.meta[s] = (uint8_t)((unary << 5) | offset);
            int pos[4];
            bvh2_leaves(b, id, pos);
            for (int k = 0; k < cnt; k++) b.prim_order[prim_base + offset + (uint32_t)k] = b.vals[pos[k]];
            offset += (uint32_t)cnt;
        }
        const float4 clo4 = b.nlo[id], chi4 = b.nhi[id];
        const float clo[3] = {clo4.x, clo4.y, clo4.z}, chi[3] = {chi4.x, chi4.y, chi4.z};
        for (int k = 0; k < 3; k++) {
            const float scale = rt3_u2f(E[k] << 23);
            int ql = (int)floorf((clo[k] - p[k]) / scale);
            ql = ql < 0 ? 0 : (ql > RT3_QMAX ? RT3_QMAX : ql);
            while (ql > 0 && (double)p[k] + (double)ql * (double)scale > (double)clo[k]) ql--;
            int qh = (int)ceilf((chi[k] - p[k]) / scale);
            qh = qh < 0 ? 0 : (qh > RT3_QMAX ? RT3_QMAX : qh);
            while (qh < RT3_QMAX && (double)p[k] + (double)qh * (double)scale < (double)chi[k]) qh++;
#if RT3_NODE_FP16
            nd.q[s >> 2][k][0][s & 3] = int_to_half_bits(ql);
            nd.q[s >> 2][k][1][s & 3] = int_to_half_bits(qh);
#else
            nd.qlo[k][s] = (uint8_t)ql;
            nd.qhi[k][s] = (uint8_t)qh;
#endif
        }
    }
    b.nodes[widx] = nd;
}

// ------------------------------------------------------------------------------------ host driver
// Hand-written bitonic sort of (Morton key, primitive) pairs, ascending by (key, primitive) so that
// equal keys keep primitive order (= a stable sort by key; the simulator's std::stable_sort agrees).
// One compare-exchange stage per launch on a power-of-two padded copy: O(n log^2 n), a few
// milliseconds for 1 M primitives — a one-off inside the build, kept simple on purpose.
RT3_GLOBAL(k_bitonic_stage, uint64_t* keys, uint32_t* vals, uint32_t j, uint32_t k) {
    const uint32_t i = RT3_THREAD_ID();
    if (i >= rt3_n_) return;
    const uint32_t l = i ^ j;
    if (l <= i) return;
    const uint64_t ka = keys[i], kb = keys[l];
    const uint32_t va = vals[i], vb = vals[l];
    const bool a_gt_b = ka > kb || (ka == kb && va > vb);
    const bool ascending = (i & k) == 0u;
    if (a_gt_b == ascending) {
        keys[i] = kb; keys[l] = ka;
        vals[i] = vb; vals[l] = va;
    }
}

inline void sort_pairs(uint64_t* keys, uint32_t* vals, uint32_t n, Stream st) {
#ifdef RT3_EMULATE
    std::vector<std::pair<uint64_t, uint32_t>> v(n);
    for (uint32_t i = 0; i < n; i++) v[i] = {keys[i], vals[i]};
    std::stable_sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (uint32_t i = 0; i < n; i++) { keys[i] = v[i].first; vals[i] = v[i].second; }
#else
    if (n < 2) return;
    uint32_t m = 1;
    while (m < n) m <<= 1;
    DevBuf<uint64_t> k2(m);
    DevBuf<uint32_t> v2(m);
    dev_memset(k2.p, 0xff, sizeof(uint64_t) * m, st);  // padding sorts to the end
    dev_memset(v2.p, 0xff, sizeof(uint32_t) * m, st);
    d2d(k2.p, keys, sizeof(uint64_t) * n, st);
    d2d(v2.p, vals, sizeof(uint32_t) * n, st);
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) RT3_LAUNCH_1D(k_bitonic_stage, m, st, k2.p, v2.p, j, k);
    d2d(keys, k2.p, sizeof(uint64_t) * n, st);
    d2d(vals, v2.p, sizeof(uint32_t) * n, st);
    stream_sync(st);
#endif
}

// ------------------------------------------------------------------------------------ host SAH builder (small inputs)
// The TLAS holds a few thousand instance boxes of very different sizes that overlap their neighbours: the case where a
// Morton-order tree is at its worst (an instance visit costs a ray a transform, three divisions and a BLAS root).  For such
// small inputs the binary tree is built on the host with the full-sweep surface-area heuristic (every split position along
// every axis, primitives sorted by centroid), written in the SAME arrays the GPU hierarchy kernels produce — internal
// nodes 0 .. n-2 (root 0), leaf of sorted position j at id n-1+j, [first, last] = the node's range of sorted positions —
// and collapsed to the wide layout by the same kernel.  O(n log^2 n); ~1 ms for 1000 boxes.
struct HostBvh2 {
    std::vector<int> left, right, first, last, parent;   // parent: [2n - 1], -1 at the root
    std::vector<float4> nlo, nhi;   // [2n - 1]
    std::vector<uint32_t> vals;     // sorted position -> primitive
};
inline void build_bvh2_sah_host(const std::vector<float4>& plo, const std::vector<float4>& phi, HostBvh2& t) {
    const int n = (int)plo.size();
    t.left.assign((size_t)n - 1, 0); t.right.assign((size_t)n - 1, 0); t.first.assign((size_t)n - 1, 0); t.last.assign((size_t)n - 1, 0);
    t.nlo.assign((size_t)2 * n, make_float4(0, 0, 0, 0)); t.nhi.assign((size_t)2 * n, make_float4(0, 0, 0, 0));
    t.vals.resize((size_t)n);
    t.parent.assign((size_t)2 * n, -1);
    for (int i = 0; i < n; i++) t.vals[(size_t)i] = (uint32_t)i;
    struct Box { float lo[3], hi[3]; };
    auto grow = [&](Box& b, uint32_t p) {
        const float l[3] = {plo[p].x, plo[p].y, plo[p].z}, h[3] = {phi[p].x, phi[p].y, phi[p].z};
        for (int k = 0; k < 3; k++) { b.lo[k] = l[k] < b.lo[k] ? l[k] : b.lo[k]; b.hi[k] = h[k] > b.hi[k] ? h[k] : b.hi[k]; }
    };
    auto area = [](const Box& b) { const float e[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]}; return e[0] * e[1] + e[1] * e[2] + e[2] * e[0]; };
    const Box empty = {{3.4e38f, 3.4e38f, 3.4e38f}, {-3.4e38f, -3.4e38f, -3.4e38f}};
    int next_internal = 0;
    struct Job { int a, b, id; };
    std::vector<Job> todo;
    std::vector<float> right_area;
    std::vector<uint32_t> best_order;
    todo.push_back({0, n - 1, next_internal++});
    while (!todo.empty()) {
        const Job j = todo.back();
        todo.pop_back();
        const int cnt = j.b - j.a + 1;
        float best_cost = 3.4e38f;
        int best_split = -1;
        for (int axis = 0; axis < 3; axis++) {
            std::stable_sort(t.vals.begin() + j.a, t.vals.begin() + j.b + 1, [&](uint32_t p, uint32_t q) {
                const float cp = axis == 0 ? plo[p].x + phi[p].x : (axis == 1 ? plo[p].y + phi[p].y : plo[p].z + phi[p].z);
                const float cq = axis == 0 ? plo[q].x + phi[q].x : (axis == 1 ? plo[q].y + phi[q].y : plo[q].z + phi[q].z);
                return cp < cq;
            });
            right_area.assign((size_t)cnt, 0.0f);
            Box rb = empty;
            for (int i = cnt - 1; i > 0; i--) { grow(rb, t.vals[(size_t)(j.a + i)]); right_area[(size_t)i] = area(rb); }
            Box lb = empty;
            for (int i = 0; i < cnt - 1; i++) {   // split after sorted element i
                grow(lb, t.vals[(size_t)(j.a + i)]);
                const float cost = area(lb) * (float)(i + 1) + right_area[(size_t)(i + 1)] * (float)(cnt - 1 - i);
                if (cost < best_cost) { best_cost = cost; best_split = i; best_order.assign(t.vals.begin() + j.a, t.vals.begin() + j.b + 1); }
            }
        }
        std::copy(best_order.begin(), best_order.end(), t.vals.begin() + j.a);
        const int m = j.a + best_split;   // left = [a, m], right = [m + 1, b]
        t.first[(size_t)j.id] = j.a; t.last[(size_t)j.id] = j.b;
        const int lc = m == j.a ? n - 1 + j.a : next_internal++;
        const int rc = m + 1 == j.b ? n - 1 + j.b : next_internal++;
        t.left[(size_t)j.id] = lc; t.right[(size_t)j.id] = rc;
        t.parent[(size_t)lc] = j.id; t.parent[(size_t)rc] = j.id;
        if (lc < n - 1) todo.push_back({j.a, m, lc});
        if (rc < n - 1) todo.push_back({m + 1, j.b, rc});
    }
    // bounds: leaves from their primitive, internal nodes from their range (ids of children are larger than the parent's)
    for (int p = 0; p < n; p++) { t.nlo[(size_t)(n - 1 + p)] = plo[t.vals[(size_t)p]]; t.nhi[(size_t)(n - 1 + p)] = phi[t.vals[(size_t)p]]; }
    for (int id = n - 2; id >= 0; id--) {
        const float4 a0 = t.nlo[(size_t)t.left[(size_t)id]], a1 = t.nhi[(size_t)t.left[(size_t)id]], c0 = t.nlo[(size_t)t.right[(size_t)id]], c1 = t.nhi[(size_t)t.right[(size_t)id]];
        t.nlo[(size_t)id] = make_float4(fminf(a0.x, c0.x), fminf(a0.y, c0.y), fminf(a0.z, c0.z), 0.0f);
        t.nhi[(size_t)id] = make_float4(fmaxf(a1.x, c1.x), fmaxf(a1.y, c1.y), fmaxf(a1.z, c1.z), 0.0f);
    }
}

// Builds a BVH8 over n primitive boxes (device arrays).  Synchronises the stream (one-off build).
// sah_host: the binary tree under the collapse comes from build_bvh2_sah_host (for small n: instance lists) instead of the LBVH.
inline void build_bvh8(const float4* d_plo, const float4* d_phi, uint32_t n, Stream st, DevBuf<Node8>& out_nodes,
                       DevBuf<uint32_t>& out_order, Bvh8& out, bool sah_host = false, bool ploc = false, int leaf_max = RT3_LEAF_MAX, bool sah_collapse = false) {
    RT3_REQUIRE(n > 0, -1, "build_bvh8: no primitives");
    // the traversal kernels tag queued triangles as (lane << 27 | index): 2^27 primitives per acceleration structure
    RT3_REQUIRE(n < (1u << 27), -1, "build_bvh8: more than 134,217,727 primitives in one acceleration structure");
    const uint32_t nn = 2 * n;
    DevBuf<uint32_t> bounds(6), flags(n), counters(4);
    DevBuf<uint64_t> keys(n);
    DevBuf<uint32_t> vals(n), order(n);
    DevBuf<float4> nlo(nn), nhi(nn);
    DevBuf<int> left(n), right(n), parent(nn), first(n), last(n);
    DevBuf<Node8> nodes(n);  // worst case; compacted below
    DevBuf<int2> qa(n), qb(n);

    BuildArrays b;
    b.plo = d_plo; b.phi = d_phi; b.n = n;
    b.bounds = bounds.p; b.keys = keys.p; b.vals = vals.p;
    b.nlo = nlo.p; b.nhi = nhi.p; b.left = left.p; b.right = right.p; b.parent = parent.p;
    b.first = first.p; b.last = last.p; b.flags = flags.p;
    b.nodes = nodes.p; b.prim_order = order.p; b.counters = counters.p; b.q_in = qa.p; b.q_out = qb.p;
    b.leaf_max = leaf_max < 1 ? 1 : (leaf_max > RT3_LEAF_MAX ? RT3_LEAF_MAX : leaf_max);

    const uint32_t init_bounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    h2d(bounds.p, init_bounds, sizeof(init_bounds), st);
    dev_memset(flags.p, 0, sizeof(uint32_t) * n, st);
    dev_memset(parent.p, 0xff, sizeof(int) * nn, st);
    if (sah_host && n > 2) {
        // small inputs (the TLAS): a full-sweep SAH binary tree built on the host, in the arrays the collapse reads
        std::vector<float4> hlo(n), hhi(n);
        d2h(hlo.data(), d_plo, sizeof(float4) * n, st);
        d2h(hhi.data(), d_phi, sizeof(float4) * n, st);
        stream_sync(st);
        HostBvh2 t;
        build_bvh2_sah_host(hlo, hhi, t);
        h2d(vals.p, t.vals.data(), sizeof(uint32_t) * n, st);
        h2d(nlo.p, t.nlo.data(), sizeof(float4) * nn, st);
        h2d(nhi.p, t.nhi.data(), sizeof(float4) * nn, st);
        h2d(left.p, t.left.data(), sizeof(int) * (n - 1), st);
        h2d(right.p, t.right.data(), sizeof(int) * (n - 1), st);
        h2d(first.p, t.first.data(), sizeof(int) * (n - 1), st);
        h2d(last.p, t.last.data(), sizeof(int) * (n - 1), st);
        h2d(parent.p, t.parent.data(), sizeof(int) * (nn - 1), st);
        stream_sync(st);
    } else {
        RT3_LAUNCH_1D(k_bvh_bounds, n, st, b);
        RT3_LAUNCH_1D(k_bvh_morton, n, st, b);
        sort_pairs(keys.p, vals.p, n, st);
        if (ploc && n > 2) {
            const uint32_t max_chunks = (n + 255) / 256;
            DevBuf<int> ca(n), cb(n), nn(n);
            DevBuf<uint32_t> fm(n), fk(n), sums(2 * (size_t)max_chunks + 2);
            RT3_LAUNCH_1D(k_ploc_leaves, n, st, b, ca.p);
            PlocArrays p;
            p.in = ca.p; p.out = cb.p; p.nn = nn.p; p.merge = fm.p; p.keep = fk.p; p.sums = sums.p;
            p.count = n; p.next_id = (int)n - 1;
            int rounds = 0;
            while (p.count > 1) {
                const uint32_t chunks = (p.count + 255) / 256;
                RT3_LAUNCH_1D(k_ploc_nearest, p.count, st, b, p);
                RT3_LAUNCH_1D(k_ploc_flags, p.count, st, p);
                RT3_LAUNCH_1D(k_ploc_scan_chunks, chunks, st, p, chunks);
                RT3_LAUNCH_1D(k_ploc_scan_sums, 1, st, p, chunks);
                RT3_LAUNCH_1D(k_ploc_merge, chunks, st, b, p, chunks);
                uint32_t tot[2];
                d2h(tot, sums.p + 2 * (size_t)chunks, sizeof(tot), st);
                stream_sync(st);
                RT3_REQUIRE(tot[0] > 0 && tot[1] == p.count - tot[0], -2, "build_bvh8: PLOC round made no progress");
                p.next_id -= (int)tot[0];
                p.count = tot[1];
                const int* t = p.in; p.in = p.out; p.out = const_cast<int*>(t);
                RT3_REQUIRE(++rounds < 4096, -2, "build_bvh8: PLOC did not converge");
            }
            RT3_REQUIRE(p.next_id == 0, -2, "build_bvh8: PLOC node count mismatch");
            if (getenv("RT3_BUILD_VERBOSE")) fprintf(stderr, "rt3 build: PLOC %u primitives, %d rounds\n", n, rounds);
        } else {
            if (n > 1) RT3_LAUNCH_1D(k_bvh_hierarchy, n - 1, st, b);
            RT3_LAUNCH_1D(k_bvh_refit, n, st, b);
        }
    }

    DevBuf<float> dp_cost;
    DevBuf<uint8_t> dp_split;
    DevBuf<uint32_t> dp_flags;
    b.dp_cost = nullptr; b.dp_split = nullptr; b.dp_flags = nullptr;
    if (sah_collapse && n > 8) {
        dp_cost.alloc(7 * (size_t)nn); dp_split.alloc(8 * (size_t)nn); dp_flags.alloc(n);
        dev_memset(dp_flags.p, 0, sizeof(uint32_t) * n, st);
        b.dp_cost = dp_cost.p; b.dp_split = dp_split.p; b.dp_flags = dp_flags.p;
        RT3_LAUNCH_1D(k_bvh_dp, n, st, b);
    }
    // collapse, level by level
    const uint32_t init_counters[4] = {1u, 0u, 0u, 0u};  // node 0 = root
    h2d(counters.p, init_counters, sizeof(init_counters), st);
    const int2 root_item = make_int2(n == 1 ? 0 : 0, 0);  // BVH2 id 0 is the root (leaf 0 when n == 1)
    h2d(qa.p, &root_item, sizeof(root_item), st);
    uint32_t qn = 1;
    int levels = 0;
    while (qn > 0) {
        RT3_LAUNCH_1D(k_bvh_collapse, qn, st, b);
        uint32_t c[4];
        d2h(c, counters.p, sizeof(c), st);
        stream_sync(st);
        qn = c[2];
        const uint32_t zero = 0;
        h2d(&counters.p[2], &zero, sizeof(zero), st);
        int2* t = b.q_in; b.q_in = b.q_out; b.q_out = t;
        RT3_REQUIRE(++levels < 256, -2, "build_bvh8: collapse did not converge");
    }
    uint32_t c[4];
    d2h(c, counters.p, sizeof(c), st);
    float4 rb[2];
    d2h(&rb[0], &nlo.p[0], sizeof(float4), st);
    d2h(&rb[1], &nhi.p[0], sizeof(float4), st);
    stream_sync(st);
    RT3_REQUIRE(c[1] == n, -2, "build_bvh8: primitive count mismatch after collapse");
    out_nodes.alloc(c[0]);
    d2d(out_nodes.p, nodes.p, sizeof(Node8) * c[0], st);  // compaction (reference: optixAccelCompact, cuda_mesh.h:146)
    out_order = std::move(order);
    stream_sync(st);
    out.nodes = out_nodes.p;
    out.prim_order = out_order.p;
    out.num_nodes = c[0];
    out.num_prims = n;
    out.lo[0] = rb[0].x; out.lo[1] = rb[0].y; out.lo[2] = rb[0].z;
    out.hi[0] = rb[1].x; out.hi[1] = rb[1].y; out.hi[2] = rb[1].z;
}

}  // namespace rt3
