// rt3_lib.cu — the C ABI of include/rt3.h over the sm_100a kernels (single translation unit).
//
// Host-side orchestration only: device-scene assembly (geometry upload + BLAS build, instance
// table + TLAS build, hit-group / light / texture tables) and the per-subframe stage schedule
//   generate -> { extend -> shade -> connect } x depth -> resolve
// issued on one CUDA stream with no host synchronisation inside the loop (queue sizes stay in
// device memory; persistent kernels read them there).
// Compiled with: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
#include "../../include/rt3.h"
#include "rt3_wavefront.cuh"

#include <memory>
#include <vector>
#ifndef RT3_EMULATE
#include <dlfcn.h>
#endif

namespace rt3 {
thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launch_count{0};   // all contexts, any host thread
}  // namespace rt3

using namespace rt3;

namespace {

struct Geometry {
    uint32_t type = PRIM_TRI, nprims = 0, vkeys = 1, nv = 0;
    std::vector<uint32_t> sub_first;   // spline curves: first linear sub-segment of every USER segment (+ end); hit records are translated at the API boundary
    DevBuf<uint32_t> sub;              // spline curves: (user segment, k | K << 16) per sub-segment
    bool spline() const { return !sub_first.empty(); }
    uint32_t user_prims() const { return spline() ? (uint32_t)sub_first.size() - 1u : nprims; }
    DevBuf<float> verts, normals, uvs, colors;   // normals / uvs / colors stay empty when the caller has none (SDK fallbacks)
    DevBuf<int32_t> idx, seg;
    DevBuf<float4> cr, poly;   // poly: spline curves, 4 power-basis coefficients (float4: xyz + radius) per USER segment
    uint32_t curve_cubic = 0;  // spline curves: 1 = cubic (its velocity nudges u = 0 / 1, curve.h:281-288)
    DevBuf<Node8> nodes;
    DevBuf<uint32_t> order;
    DevBuf<float4> prims;
    Bvh8 bvh;
    bool has_blas = false;  // built lazily: meshes that are only used by merged (identity) instances never need one
    float bsphere[4] = {0, 0, 0, -1};  // object-space bounding sphere (centre of the bounding box, radius to the farthest point), from the host arrays at creation
};

struct InstanceHost {
    uint32_t blas = 0;
    float xform[12];
    std::vector<float> keys;
    uint32_t nkeys = 0;
    float t0 = 0, t1 = 1;
    HitGroupDev hg;
};

struct TextureObj { DevBuf<uchar4> px; int w = 0, h = 0, addr = 0, filt = 0; };

constexpr uint32_t MAX_DEPTH_SLOTS = 1024;
#ifndef RT3_TRAV_BLOCKS_SINGLE_SHARED
// persistent CTAs per SM of the single-level kernel while other subframes are in flight: a smaller grid leaves room for their
// kernels (8 / 7 / 6 / 5 / 4 / 3: 2977 / 2983 / 2995 / 3010 / 3011 / 2903 Mrays/s on C2); alone on the GPU the kernel keeps its 8
#define RT3_TRAV_BLOCKS_SINGLE_SHARED 5
#endif
#ifndef RT3_PIPE_SLOTS
#define RT3_PIPE_SLOTS 3   // subframes in flight (pool sets); measured on C2: 1 -> 2863, 2 -> 2953, 3 -> see DESIGN
#endif

}  // namespace

struct rt3_context {
    int device = 0;
    Stream stream = 0;
    int num_sms = 148;
    std::vector<std::unique_ptr<Geometry>> geoms;
    std::vector<InstanceHost> inst;
    std::vector<std::unique_ptr<TextureObj>> textures;
    bool built = false;
    // device scene tables
    DevBuf<BlasDev> d_blas;
    DevBuf<InstanceDev> d_inst;
    DevBuf<HitGroupDev> d_hg;
    DevBuf<float> d_keys, d_static;
    DevBuf<TexDev> d_tex;
    DevBuf<Light> d_lights;
    DevBuf<float> d_light_cdf;  // running sum of the power-sampler weights (mode 2)
    uint32_t nlights = 0;
    DevBuf<Node8> tlas_nodes;
    DevBuf<uint32_t> tlas_order;
    Bvh8 tlas;
    // merged world BLAS of the identity static mesh instances (single-level fast path)
    DevBuf<Node8> m_nodes;
    DevBuf<uint32_t> m_order;
    DevBuf<float4> m_prims;
    DevBuf<uint2> m_map;
    DevBuf<uint8_t> m_slab;
    Node8* m_nodes_p = nullptr;
    float4* m_prims_p = nullptr;
    int opt_l2_persist = 0;  // measured: -5 % on C2 (the carve-out starves the streaming queue traffic), off by default
    Bvh8 m_bvh;
    bool has_merged = false, single_level = false;
    bool has_subdiv_curves = false;  // some instance refers to a spline curve geometry (set in rt3_accel_build)
    int opt_merge = 1;
    int opt_packets = 1;   // camera rays (depth 0) traverse the merged BLAS in packets of eight (k_extend_packets): 1 = when the BLAS fits the L2, 2 = always
    size_t l2_bytes = 0;
    int opt_flatten = 1;   // static, transformed triangle-mesh instances join the merged world BLAS (vertices transformed once, at build)
    int opt_split = 1;     // merged BLAS beside other instances: 0 = one TLAS over both, 2 = always two passes (single-level kernel, then the rest), 1 = two passes when the merged BLAS is large
    bool split = false;
    uint32_t host_flags = 0;   // error_flags bits raised on the host (bit 1: an unbounded-depth launch ran out of depth slots)
    DevBuf<float4> q_rays, q_hits;   // staging of rt3_trace / rt3_get_local_geometry (grown on demand, kept between calls)
    DevBuf<float> q_out;
    uint32_t flattened = 0;  // instances flattened by the last build
    int opt_tlas_sah = 1;
    int opt_tlas_leaf = 1;      // instances per TLAS leaf child (r02j, C3 / C4 Mrays/s: 3 -> 762 / 957, 2 -> 785 / 927, 1 -> 815 / 953: an own box per instance culls more entries than the extra TLAS nodes cost)
    int opt_sah_collapse = 1;   // binary -> wide collapse by the SAH dynamic program (k_bvh_dp); 0 = greedy largest-area opening
    int opt_bsphere_cull = 1;   // skip TLAS leaves whose instance bounding sphere the ray misses
    int opt_ploc = 0;       // 1: BLAS binary tree by parallel locally-ordered clustering instead of the Morton-order tree (measured neutral on the tessellated BASELINE meshes)   // TLAS binary tree from the host full-sweep SAH builder (small inputs) instead of the LBVH
    int opt_tlas_refine = 1;  // instance boxes from the BLAS root's grandchild boxes instead of its root box
    DevBuf<float4> d_bsphere;  // [instance][2] world-space bounding spheres (TravScene::inst_bsphere)
    DevBuf<uint32_t> d_flags;  // [0] error flags, [1] max stack
    // film
    uint32_t width = 0, height = 0;
    DevBuf<float4> accum;
    DevBuf<uchar4> frame;
    // wavefront pools
    // RT3_PIPE_SLOTS sets ("slots") of them: consecutive subframes rotate through the slots, each with its own pair of streams, so that
    // the head of subframe k + 1 (generate, the camera-ray packets) fills the GPU while the last bounces of subframe k — a dozen
    // launches over the few paths still alive — drain ("pipeline").  Only the resolves, which update the film in subframe
    // order, stay on the context's stream.  The slots are allocated together, at the first pipelined launch.
    struct Pools {
        size_t paths = 0;
        DevBuf<float4> ray[2][3], st[2][2], hit0, sh[4], result;
        DevBuf<int32_t> hit_inst;
        DevBuf<uint32_t> counters;       // 2 chains x 6 x MAX_DEPTH_SLOTS
        Event done;                      // recorded on the context's stream behind the resolve of the subframe that used the slot
    } pool[RT3_PIPE_SLOTS];
    Stream pipe_stream[RT3_PIPE_SLOTS][2] = {};   // [slot][main, aux]
    Event ev_pipe_join[RT3_PIPE_SLOTS];
    int pipe_slot = 0;
    bool shared_gpu = false;   // this subframe is launched while earlier ones are still running: see RT3_TRAV_BLOCKS_SINGLE_SHARED
    int opt_pipeline = 1;
    DevBuf<unsigned long long> d_stats;  // primary, bounce, shadow
    DevBuf<uint32_t> trace_fetch;        // [2]: one work counter per pass
    // options / stats
    int opt_timing = 0;
    // connect(d) (shadow rays) and extend(d+1) (next bounce) are independent: connect runs on a second stream so that
    // the next extend fills the tail of its persistent CTAs; shade(d+1) waits for it (shadow queue + radiance RMW)
    int opt_overlap = 1;
    Stream stream2 = 0, stream3 = 0, stream4 = 0;   // aux of chain 0; main + aux of chain 1 ("overlap" >= 2)
    Stream stream_copy = 0;                         // rt3_download_frame_async: the copy engine works beside the next subframe
    Event ev_frame_ready, ev_frame_copied;
    bool frame_copy_pending = false;
    Event ev_shade[RT3_PIPE_SLOTS][2], ev_connect[RT3_PIPE_SLOTS][2], ev_fork, ev_join;   // [chain, or slot when pipelined][depth parity]
    int opt_ctas_per_sm = 0;
    uint64_t samples = 0;
    float ms[6] = {0, 0, 0, 0, 0, 0};
    bool hitgroups_dirty = true;
#ifndef RT3_EMULATE
    // single-process multi-GPU reduce (rt3_allreduce_accum): this context's communicator, kept between calls, and the
    // device list + rank it was created for; destroyed with the context
    void* nccl_comm = nullptr;
    std::vector<int> nccl_devs;
    int nccl_rank = -1;
#endif

    // merged_only: the scene as pass 1 of a split traversal sees it — the merged world BLAS and nothing else
    TravScene trav_scene(bool merged_only = false) {
        const bool single_level = this->single_level || merged_only;
        TravScene s;
        s.tlas_nodes = single_level ? m_nodes_p : tlas_nodes.p;
        s.tlas_order = tlas_order.p;
        s.merged_map = m_map.p;
        s.magic_h2 = 0x64646464u;
        s.bias_h = -1024.0f;
        s.root_prims = single_level ? m_prims_p : nullptr;
        s.root_is_blas = single_level ? 1u : 0u;
        s.instances = d_inst.p;
        s.hitgroups = d_hg.p;
        s.blas = d_blas.p;
        s.keys = d_keys.p;
        s.inst_fwd = d_static.p;
        s.inst_bsphere = d_bsphere.p;
        s.error_flags = d_flags.p;
        s.max_stack = d_flags.p + 1;
        return s;
    }
    int trav_grid(bool single) const {
#ifdef RT3_EMULATE
        (void)single;
        return 1;
#else
        const int per_sm = opt_ctas_per_sm > 0 ? opt_ctas_per_sm : (single ? (shared_gpu ? RT3_TRAV_BLOCKS_SINGLE_SHARED : RT3_TRAV_MIN_BLOCKS_SINGLE) : RT3_TRAV_MIN_BLOCKS);
        return num_sms * per_sm;
#endif
    }
};

namespace {

template <int MODE, bool SINGLE>
void launch_traverse_kernel(rt3_context* c, const TraverseArgs& a, Stream st) {
#ifdef RT3_EMULATE
    (void)c; (void)st;
    k_traverse<MODE, SINGLE>(a);
#else
    k_traverse<MODE, SINGLE><<<c->trav_grid(SINGLE), RT3_TRAV_THREADS, 0, st>>>(a);
    RT3_CUDA(cudaGetLastError());
#endif
    count_launch();
}
// camera rays (depth 0 of a subframe: `packet_rays` of them, known on the host) through the merged BLAS in packets of eight
void launch_extend_packets(rt3_context* c, TraverseArgs a, uint32_t packet_rays, Stream st) {
#ifdef RT3_EMULATE
    (void)packet_rays; (void)st;
    k_traverse<TRAV_EXTEND, true>(a);   // the simulator has no warps: the per-ray traversal returns the same hits
#else
    a.count_ptr = nullptr;
    a.count = packet_rays;
    k_extend_packets<<<(packet_rays + 127u) / 128u, 128, 0, st>>>(a);
    RT3_CUDA(cudaGetLastError());
#endif
    count_launch();
}
// fetch2: the work counter of the second launch of a split traversal (zeroed like a.fetch); packet_rays > 0: the rays are the
// camera rays of a subframe, that many
template <int MODE>
void launch_traverse(rt3_context* c, TraverseArgs a, uint32_t* fetch2, Stream st, uint32_t packet_rays = 0) {
    a.pass = 0u;
    // packets put FOUR independent traversals into a warp where the per-ray kernel has thirty-two: fine while most nodes come from
    // L2, latency-bound once they come from DRAM.  Measured (C2's terrain at growing size, Mrays/s without -> with packets): 1.0 M
    // triangles / 64 MB 2656 -> 2861, 1.5 M 2561 -> 2742, 2 M / 128 MB 2506 -> 2664, 3.9 M / 250 MB 2370 -> 2472, 8 M / 512 MB
    // 2235 -> 2265, 15.7 M / 1 GB 2101 -> 2047; C3 flattened (100 M / 5.9 GB, many silhouettes) 1326 -> 287.  Used up to four times
    // the L2 ("packets" = 2: always)
    const bool packets = MODE == TRAV_EXTEND && packet_rays > 0 && c->opt_packets && (c->opt_packets >= 2 || c->m_slab.bytes() <= 4 * c->l2_bytes);
    if (c->single_level) {   // merged world BLAS only: the lean instantiation
        if (packets) launch_extend_packets(c, a, packet_rays, st);
        else launch_traverse_kernel<MODE, true>(c, a, st);
    } else if (c->split) {   // the bulk of the triangles with the lean kernel, then everything else, seeded with what that found
        TraverseArgs p1 = a;
        p1.scene = c->trav_scene(true);
        p1.pass = 1u;
        if (packets) launch_extend_packets(c, p1, packet_rays, st);
        else launch_traverse_kernel<MODE, true>(c, p1, st);
        a.pass = 2u;
        a.fetch = fetch2;
        a.stat = nullptr;
        launch_traverse_kernel<MODE, false>(c, a, st);
    } else launch_traverse_kernel<MODE, false>(c, a, st);
}

void upload_hitgroups(rt3_context* c) {
    if (!c->hitgroups_dirty || c->inst.empty()) return;
    std::vector<HitGroupDev> hg(c->inst.size() + 1);  // + the merged pseudo-instance (never shaded)
    for (size_t i = 0; i < c->inst.size(); i++) hg[i] = c->inst[i].hg;
    hg[c->inst.size()] = HitGroupDev{{0, 0, 0}, {0, 0, 0}, -1, 1.0f, {1, 1}, {0, 1}, {0, 0}, 0u, 0u};
    stream_sync(c->stream);   // subframes in flight on the pipeline's streams still read the old records
    c->d_hg.ensure(hg.size());
    h2d(c->d_hg.p, hg.data(), sizeof(HitGroupDev) * hg.size(), c->stream);
    stream_sync(c->stream);
    c->hitgroups_dirty = false;
}

// object-space bounding sphere of n points p[i] (stride floats apart) each with its own radius r[i] (or 0)
static void bounding_sphere(const float* p, size_t stride, const float* r, size_t rstride, size_t n, float out[4]) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            const double rr = r ? (double)r[i * rstride] : 0.0, v = p[i * stride + k];
            lo[k] = v - rr < lo[k] ? v - rr : lo[k];
            hi[k] = v + rr > hi[k] ? v + rr : hi[k];
        }
    const double c[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
    double rad = 0.0;
    for (size_t i = 0; i < n; i++) {
        const double dx = p[i * stride] - c[0], dy = p[i * stride + 1] - c[1], dz = p[i * stride + 2] - c[2];
        const double d = std::sqrt(dx * dx + dy * dy + dz * dz) + (r ? (double)r[i * rstride] : 0.0);
        rad = d > rad ? d : rad;
    }
    out[0] = (float)c[0]; out[1] = (float)c[1]; out[2] = (float)c[2];
    out[3] = (float)(rad * (1.0 + 1e-5) + 1e-30);
}
// largest singular value of the 3x3 part of a row-major 3x4 matrix (power iteration on A^T A), rounded up
static double spectral_norm(const float* m) {
    double a[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}}, ata[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) ata[i][j] = a[0][i] * a[0][j] + a[1][i] * a[1][j] + a[2][i] * a[2][j];
    double v[3] = {0.577, 0.577, 0.577}, lam = 0.0;
    for (int it = 0; it < 64; it++) {
        double w[3];
        for (int i = 0; i < 3; i++) w[i] = ata[i][0] * v[0] + ata[i][1] * v[1] + ata[i][2] * v[2];
        lam = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (lam < 1e-300) return 0.0;
        for (int i = 0; i < 3; i++) v[i] = w[i] / lam;
    }
    // power iteration converges from below: bound it from above by the Frobenius norm where it has not converged
    const double fro = std::sqrt(ata[0][0] + ata[1][1] + ata[2][2]);
    const double s = std::sqrt(lam) * 1.001;
    return s < fro ? s : fro;
}

uint64_t finish_geometry(rt3_context* c, std::unique_ptr<Geometry> g) {
    c->geoms.push_back(std::move(g));
    c->built = false;
    return (uint64_t)c->geoms.size();  // handle = index + 1
}

// BLAS of one geometry (reference: the GAS build inside the CUDAMesh ctor, cuda_mesh.h:92-146), on first need
void ensure_blas(rt3_context* c, Geometry* g) {
    if (g->has_blas) return;
    DevBuf<float4> lo(g->nprims), hi(g->nprims);
    if (g->type == PRIM_TRI) RT3_LAUNCH_1D(k_tri_boxes, g->nprims, c->stream, (const float*)g->verts.p, (const int32_t*)g->idx.p, lo.p, hi.p);
    else if (g->type == PRIM_TRI_MOTION) RT3_LAUNCH_1D(k_tri_boxes_motion, g->nprims, c->stream, (const float*)g->verts.p, (const int32_t*)g->idx.p, g->vkeys, g->nv, lo.p, hi.p);
    else if (g->type == PRIM_SPHERE) RT3_LAUNCH_1D(k_sphere_boxes, g->nprims, c->stream, (const float4*)g->cr.p, lo.p, hi.p);
    else RT3_LAUNCH_1D(k_curve_boxes, g->nprims, c->stream, (const float4*)g->cr.p, (const int32_t*)g->seg.p, lo.p, hi.p);
    build_bvh8(lo.p, hi.p, g->nprims, c->stream, g->nodes, g->order, g->bvh, /*sah_host=*/false, /*ploc=*/c->opt_ploc != 0, RT3_LEAF_MAX, /*sah_collapse=*/c->opt_sah_collapse != 0);
    g->prims.alloc(3 * (size_t)g->nprims * g->vkeys);
    if (g->type == PRIM_TRI_MOTION) RT3_LAUNCH_1D(k_pack_tris_motion, g->nprims, c->stream, (const float*)g->verts.p, (const int32_t*)g->idx.p, g->vkeys, g->nv, (const uint32_t*)g->order.p, g->prims.p);
    else if (g->type == PRIM_TRI) RT3_LAUNCH_1D(k_pack_tris, g->nprims, c->stream, (const float*)g->verts.p, (const int32_t*)g->idx.p, (const uint32_t*)g->order.p, g->prims.p);
    else if (g->type == PRIM_SPHERE) RT3_LAUNCH_1D(k_pack_spheres, g->nprims, c->stream, (const float4*)g->cr.p, (const uint32_t*)g->order.p, g->prims.p);
    else RT3_LAUNCH_1D(k_pack_curves, g->nprims, c->stream, (const float4*)g->cr.p, (const int32_t*)g->seg.p, (const uint32_t*)g->order.p, g->prims.p);
    stream_sync(c->stream);
    g->has_blas = true;
}

// Keep the merged BLAS (nodes + primitive records, re-read by every ray) resident in L2 against the
// streaming queue traffic: persisting access-policy window on the context's stream.
void set_l2_window(rt3_context* c) {
#ifndef RT3_EMULATE
    if (!c->opt_l2_persist || !c->m_slab.p) return;
    cudaDeviceProp prop;
    RT3_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    const size_t want = c->m_slab.bytes();
    const size_t carve = want < (size_t)prop.persistingL2CacheMaxSize ? want : (size_t)prop.persistingL2CacheMaxSize;
    RT3_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = c->m_slab.p;
    v.accessPolicyWindow.num_bytes = want < (size_t)prop.accessPolicyMaxWindowSize ? want : (size_t)prop.accessPolicyMaxWindowSize;
    v.accessPolicyWindow.hitRatio = want <= carve ? 1.0f : (float)carve / (float)want;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    RT3_CUDA(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &v));
#else
    (void)c;
#endif
}

void ensure_pools(rt3_context* c, int slot, size_t paths) {
    rt3_context::Pools& p = c->pool[slot];
    if (!p.counters.p) {
        p.counters.alloc(12 * MAX_DEPTH_SLOTS);  // two chains
        dev_memset(p.counters.p, 0, p.counters.bytes(), c->stream);
    }
    if (paths <= p.paths) return;
    stream_sync(c->stream);   // every subframe in flight ends in a resolve on this stream
    for (int b = 0; b < 2; b++) {
        for (int k = 0; k < 3; k++) p.ray[b][k].alloc(paths);
        for (int k = 0; k < 2; k++) p.st[b][k].alloc(paths);
    }
    p.hit0.alloc(paths);
    p.hit_inst.alloc(paths);
    for (int k = 0; k < 4; k++) p.sh[k].alloc(paths);
    p.result.alloc(paths);
    p.paths = paths;
}

void ensure_film(rt3_context* c, uint32_t w, uint32_t h) {
    if (c->width == w && c->height == h && c->accum.p) return;
    if (c->frame_copy_pending) { stream_sync(c->stream_copy); c->frame_copy_pending = false; }   // the old frame is still being read
    c->width = w;
    c->height = h;
    c->accum.alloc((size_t)w * h);  // handleResize reallocates the accumulation buffer (src/wavefront.cpp:178-191)
    c->frame.alloc((size_t)w * h);
    dev_memset(c->accum.p, 0, c->accum.bytes(), c->stream);
    dev_memset(c->frame.p, 0, c->frame.bytes(), c->stream);
}

}  // namespace

// several contexts (one per GPU) may live in one process: every entry point makes its context's device current
static inline void use_device(rt3_context* c) {
#ifndef RT3_EMULATE
    if (c) cudaSetDevice(c->device);
#else
    (void)c;
#endif
}
#define RT3_API_BEGIN try {
#define RT3_API_END                                       \
    }                                                     \
    catch (const rt3::Error& e) {                         \
        rt3::g_last_error = e.what();                     \
        return e.code;                                    \
    }                                                     \
    catch (const std::exception& e) {                     \
        rt3::g_last_error = e.what();                     \
        return RT3_ERR_INVALID;                           \
    }                                                     \
    return RT3_OK;


// Device-side error word (bit 0: a traversal stack overflowed and dropped a subtree).  Read wherever the call synchronises
// anyway, so that a wrong image or hit list is reported instead of returned silently.
static void require_no_device_error(rt3_context* c, const char* where) {
    uint32_t fl = 0;
    d2h(&fl, c->d_flags.p, sizeof(fl), c->stream);
    stream_sync(c->stream);
    if (fl & 1u) throw Error(RT3_ERR_STATE, std::string(where) + ": a traversal stack overflowed (more than " + std::to_string(RT3_STACK_SIZE) +
                                                " pending entries): results are incomplete; rt3_reset_stats clears the flag");
}

#ifndef RT3_EMULATE
// NCCL is resolved at run time so that librt3.so has no link-time dependency on it
struct NcclApi {
    bool ok = false;
    int (*initAll)(void**, int, const int*) = nullptr;
    int (*allReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*groupStart)(void) = nullptr;
    int (*groupEnd)(void) = nullptr;
    int (*commDestroy)(void*) = nullptr;
};
static const NcclApi& nccl_api() {
    static const NcclApi api = [] {
        NcclApi a;
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return a;
        a.initAll = (int (*)(void**, int, const int*))dlsym(lib, "ncclCommInitAll");
        a.allReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
        a.groupStart = (int (*)(void))dlsym(lib, "ncclGroupStart");
        a.groupEnd = (int (*)(void))dlsym(lib, "ncclGroupEnd");
        a.commDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
        a.ok = a.initAll && a.allReduce && a.groupStart && a.groupEnd && a.commDestroy;
        return a;
    }();
    return api;
}
static void drop_nccl_comm(rt3_context* c) {
    if (c->nccl_comm) { cudaSetDevice(c->device); nccl_api().commDestroy(c->nccl_comm); }
    c->nccl_comm = nullptr; c->nccl_devs.clear(); c->nccl_rank = -1;
}
#endif

extern "C" {

const char* rt3_last_error(void) { return rt3::g_last_error.c_str(); }

int rt3_context_create(int device, rt3_context_t* out) {
    RT3_API_BEGIN
    RT3_REQUIRE(out, RT3_ERR_INVALID, "context_create: null output");
    *out = nullptr;
    auto c = std::make_unique<rt3_context>();
    c->device = device;
#ifndef RT3_EMULATE
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        throw Error(RT3_ERR_NO_DEVICE, std::string("context_create: no usable CUDA device (") + cudaGetErrorString(e) +
                                           "); librt3 has no CPU fallback");
    RT3_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RT3_CUDA(cudaGetDeviceProperties(&prop, device));
    RT3_REQUIRE(prop.major >= 10, RT3_ERR_NO_DEVICE, "context_create: kernels are built for sm_100a only");
    c->num_sms = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    RT3_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RT3_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    RT3_CUDA(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
    RT3_CUDA(cudaStreamCreateWithFlags(&c->stream4, cudaStreamNonBlocking));
    RT3_CUDA(cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking));
    for (int sl = 0; sl < RT3_PIPE_SLOTS; sl++)
        for (int k = 0; k < 2; k++) RT3_CUDA(cudaStreamCreateWithFlags(&c->pipe_stream[sl][k], cudaStreamNonBlocking));
    if (const char* e = getenv("RT3_PIPELINE")) c->opt_pipeline = atoi(e);
    if (const char* e = getenv("RT3_CTAS")) c->opt_ctas_per_sm = atoi(e);
    if (const char* e = getenv("RT3_PACKETS")) c->opt_packets = atoi(e);
    if (const char* e = getenv("RT3_OVERLAP")) c->opt_overlap = atoi(e);  // A/B switch for measurements; rt3_set_option("overlap", v) is the API
#endif
    c->d_flags.alloc(16);  // [0] error flags, [1] max stack, [2..15] diagnostic counters (RT3_STATS builds)
    c->d_stats.alloc(4);
    c->trace_fetch.alloc(2);
    dev_memset(c->d_flags.p, 0, c->d_flags.bytes(), c->stream);
    dev_memset(c->d_stats.p, 0, c->d_stats.bytes(), c->stream);
    stream_sync(c->stream);
    *out = c.release();
    RT3_API_END
}

void rt3_context_destroy(rt3_context_t c) {
    if (!c) return;
#ifndef RT3_EMULATE
    cudaSetDevice(c->device);
    drop_nccl_comm(c);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream2); cudaStreamSynchronize(c->stream3); cudaStreamSynchronize(c->stream4); cudaStreamSynchronize(c->stream_copy);
    cudaStream_t s = c->stream, s2 = c->stream2, s3 = c->stream3, s4 = c->stream4, s5 = c->stream_copy;
    cudaStream_t ps[2 * RT3_PIPE_SLOTS];
    for (int sl = 0; sl < RT3_PIPE_SLOTS; sl++) { ps[2 * sl] = c->pipe_stream[sl][0]; ps[2 * sl + 1] = c->pipe_stream[sl][1]; }
    for (cudaStream_t q : ps) cudaStreamSynchronize(q);
    delete c;
    cudaStreamDestroy(s); cudaStreamDestroy(s2); cudaStreamDestroy(s3); cudaStreamDestroy(s4); cudaStreamDestroy(s5);
    for (cudaStream_t q : ps) cudaStreamDestroy(q);
#else
    delete c;
#endif
}

int rt3_sync(rt3_context_t c) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c, RT3_ERR_INVALID, "sync: null context");
    stream_sync(c->stream);
    if (c->frame_copy_pending) {   // an asynchronous frame download: finished, and as trustworthy as a synchronous one
        stream_sync(c->stream_copy);
        c->frame_copy_pending = false;
        require_no_device_error(c, "download_frame_async");
    }
    RT3_API_END
}

int rt3_get_stream(rt3_context_t c, void** stream) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && stream, RT3_ERR_INVALID, "get_stream: null argument");
#ifdef RT3_EMULATE
    *stream = nullptr;
#else
    *stream = (void*)c->stream;
#endif
    RT3_API_END
}

int rt3_set_option(rt3_context_t c, const char* key, int value) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && key, RT3_ERR_INVALID, "set_option: null argument");
    const std::string k(key);
    if (k == "timing") c->opt_timing = value;
    else if (k == "overlap") c->opt_overlap = value;
    else if (k == "persist_ctas_per_sm") c->opt_ctas_per_sm = value;
    else if (k == "merge_identity") { c->opt_merge = value; c->built = false; }
    else if (k == "flatten") { c->opt_flatten = value; c->built = false; }
    else if (k == "packets") c->opt_packets = value;
    else if (k == "pipeline") c->opt_pipeline = value;
    else if (k == "split") { c->opt_split = value; c->built = false; }
    else if (k == "tlas_sah") { c->opt_tlas_sah = value; c->built = false; }
    else if (k == "bsphere_cull") { c->opt_bsphere_cull = value; c->built = false; }
    else if (k == "tlas_leaf") { c->opt_tlas_leaf = value; c->built = false; }
    else if (k == "sah_collapse") { c->opt_sah_collapse = value; c->built = false; for (auto& g : c->geoms) g->has_blas = false; }
    else if (k == "ploc") { c->opt_ploc = value; c->built = false; for (auto& g : c->geoms) g->has_blas = false; }
    else if (k == "l2_persist") { c->opt_l2_persist = value; c->built = false; }
    else if (k == "tlas_refine") { c->opt_tlas_refine = value; c->built = false; }
    else if (k == "sort_rays" || k == "sort_materials") { RT3_REQUIRE(value == 0, RT3_ERR_UNSUPPORTED, "set_option: sorting stages are not built yet"); }
    else throw Error(RT3_ERR_INVALID, "set_option: unknown key " + k);
    RT3_API_END
}

// ------------------------------------------------------------------------------------ geometry
int rt3_mesh_create(rt3_context_t c, const float* verts, int num_keys, int nv, const int32_t* idx, int nt, const float* normals,
                    const float* uvs, rt3_handle_t* blas) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && verts && idx && blas, RT3_ERR_INVALID, "mesh_create: null argument");
    RT3_REQUIRE(nv > 0 && nt > 0 && num_keys >= 1, RT3_ERR_INVALID, "mesh_create: empty mesh");
    for (size_t i = 0; i < 3 * (size_t)nt; i++) RT3_REQUIRE(idx[i] >= 0 && idx[i] < nv, RT3_ERR_INVALID, "mesh_create: index out of range");
    auto g = std::make_unique<Geometry>();
    g->type = num_keys > 1 ? PRIM_TRI_MOTION : PRIM_TRI;  // motionOptions.numKeys = mesh.num_keys, cuda_mesh.h:85
    g->vkeys = (uint32_t)num_keys;
    g->nv = (uint32_t)nv;
    g->nprims = (uint32_t)nt;
    g->verts.alloc(3 * (size_t)nv * (size_t)num_keys);
    g->idx.alloc(3 * (size_t)nt);
    bounding_sphere(verts, 3, nullptr, 0, (size_t)nv * (size_t)num_keys, g->bsphere);
    h2d(g->verts.p, verts, g->verts.bytes(), c->stream);  // [key][vertex][3]
    if (normals) { g->normals.alloc(3 * (size_t)nv); h2d(g->normals.p, normals, g->normals.bytes(), c->stream); }
    if (uvs) { g->uvs.alloc(2 * (size_t)nv); h2d(g->uvs.p, uvs, g->uvs.bytes(), c->stream); }
    h2d(g->idx.p, idx, g->idx.bytes(), c->stream);
    stream_sync(c->stream);  // host arrays are borrowed for the duration of the call only
    *blas = finish_geometry(c, std::move(g));
    RT3_API_END
}

int rt3_mesh_set_colors(rt3_context_t c, rt3_handle_t blas, const float* rgba) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && rgba && blas >= 1 && blas <= c->geoms.size(), RT3_ERR_INVALID, "mesh_set_colors: bad argument");
    Geometry& g = *c->geoms[(size_t)blas - 1];
    RT3_REQUIRE(g.type == PRIM_TRI || g.type == PRIM_TRI_MOTION, RT3_ERR_INVALID, "mesh_set_colors: not a triangle mesh");
    g.colors.alloc(4 * (size_t)g.nv);
    h2d(g.colors.p, rgba, g.colors.bytes(), c->stream);
    stream_sync(c->stream);
    c->built = false;   // the device-side geometry table is rewritten by the next rt3_accel_build
    RT3_API_END
}

int rt3_spheres_create(rt3_context_t c, const float* cr, int n, rt3_handle_t* blas) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && cr && blas && n > 0, RT3_ERR_INVALID, "spheres_create: bad argument");
    auto g = std::make_unique<Geometry>();
    g->type = PRIM_SPHERE;
    g->nprims = (uint32_t)n;
    g->cr.alloc(n);
    bounding_sphere(cr, 4, cr + 3, 4, (size_t)n, g->bsphere);
    h2d(g->cr.p, cr, g->cr.bytes(), c->stream);
    stream_sync(c->stream);
    *blas = finish_geometry(c, std::move(g));
    RT3_API_END
}


// Spline curves (quadratic / cubic B-spline, Catmull-Rom, Bezier through the SDK's interpolators, cuda/curve.h:98-243):
// every user segment is intersected as K round linear sub-segments between the points P(k / K) of its true polynomial, and
// shaded with the SDK's surfaceNormal<> of that polynomial (k_shade / k_local_geometry).  K adapts to the segment: a chord
// over 1 / K of the parameter range leaves the curve by at most max|P''| / (8 K^2), and P'' is linear in u, so the maximum
// sits at u = 0 or 1.  K is the smallest count that keeps this bound — for the axis and for the radius — below RT3_CURVE_TOL
// times the segment's largest radius (taken at u = 0, 1/2, 1), clamped to [1, RT3_CURVE_MAX_SUBDIV].  Inside the library the
// sub-segments are ordinary linear-curve primitives; hit records are translated at the API boundary (Geometry::sub_first on
// the host, BlasDev::sub on the device): prim = user segment, u = (k + u_sub) / K.
#ifndef RT3_CURVE_TOL
#define RT3_CURVE_TOL 0.02
#endif
#ifndef RT3_CURVE_MAX_SUBDIV
#define RT3_CURVE_MAX_SUBDIV 64
#endif
// Spline segments -> power-basis coefficients, table driven.  A row lists the terms (control point, factor) of one coefficient
// in the order the SDK's interpolators add them (cuda/curve.h:43-47 linear, :102-110 quadratic B-spline, :176-186 cubic
// B-spline, :209-219 Catmull-Rom, :233-243 Bezier); `div` is the common denominator, applied as a multiplication by its
// rounded reciprocal like vec_math's float4 / float.  Coefficients are stored cubic-style, highest power first, with the
// leading ones zero for lower degrees: the cubic Horner forms then evaluate to exactly the lower-degree ones.
struct BasisRow { int n; int cp[4]; float f[4]; };
struct BasisTable { int ncp; float div; BasisRow row[4]; };
static const BasisTable& basis_table(int basis) {
    static const BasisTable T[5] = {
        /* linear      */ {2, 0.0f, {{0, {0}, {0}}, {0, {0}, {0}}, {2, {1, 0}, {1.0f, -1.0f}}, {1, {0}, {1.0f}}}},
        /* quad bspl   */ {3, 2.0f, {{0, {0}, {0}}, {3, {0, 1, 2}, {1.0f, -2.0f, 1.0f}}, {2, {0, 1}, {-2.0f, 2.0f}}, {2, {0, 1}, {1.0f, 1.0f}}}},
        /* cubic bspl  */ {4, 6.0f, {{4, {0, 1, 2, 3}, {-1.0f, 3.0f, -3.0f, 1.0f}}, {3, {0, 1, 2}, {3.0f, -6.0f, 3.0f}}, {2, {0, 2}, {-3.0f, 3.0f}}, {3, {0, 1, 2}, {1.0f, 4.0f, 1.0f}}}},
        /* catmull-rom */ {4, 2.0f, {{4, {0, 1, 2, 3}, {-1.0f, 3.0f, -3.0f, 1.0f}}, {4, {0, 1, 2, 3}, {2.0f, -5.0f, 4.0f, -1.0f}}, {2, {0, 2}, {-1.0f, 1.0f}}, {1, {1}, {2.0f}}}},
        /* bezier      */ {4, 0.0f, {{4, {0, 1, 2, 3}, {-1.0f, 3.0f, -3.0f, 1.0f}}, {3, {0, 1, 2}, {3.0f, -6.0f, 3.0f}}, {2, {0, 1}, {-3.0f, 3.0f}}, {1, {0}, {1.0f}}}},
    };
    return T[basis - 1];
}
// coef[4][4]: coefficient j (u^(3-j)) x component (x, y, z, radius) of the segment whose first control point is q
static void curve_coefficients(int basis, const float* q, float coef[4][4]) {
    const BasisTable& t = basis_table(basis);
    const float inv = t.div != 0.0f ? 1.0f / t.div : 1.0f;
    for (int j = 0; j < 4; j++)
        for (int c = 0; c < 4; c++) {
            const BasisRow& r = t.row[j];
            float v = 0.0f;
            for (int k = 0; k < r.n; k++) {
                const float term = q[4 * r.cp[k] + c] * r.f[k];
                v = k == 0 ? term : v + term;
            }
            coef[j][c] = (r.n && t.div != 0.0f) ? v * inv : v;
        }
}
static int curve_pieces(const float coef[4][4]) {
    double bend = 0.0;   // max |P''| over the segment, axis and radius
    for (int end = 0; end < 2; end++) {
        double acc[4];
        for (int c = 0; c < 4; c++) acc[c] = 2.0 * (double)coef[1][c] + (end ? 6.0 * (double)coef[0][c] : 0.0);
        bend = std::max(bend, std::max(std::sqrt(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]), std::fabs(acc[3])));
    }
    const double r0 = (double)coef[3][3];
    const double r1 = (double)coef[0][3] + (double)coef[1][3] + (double)coef[2][3] + (double)coef[3][3];
    const double rh = 0.125 * (double)coef[0][3] + 0.25 * (double)coef[1][3] + 0.5 * (double)coef[2][3] + (double)coef[3][3];
    const double rref = std::max(r0, std::max(r1, rh));
    if (!(rref > 0.0)) return 1;
    const double need = bend / (8.0 * RT3_CURVE_TOL * rref);   // K^2 >= need
    if (!(need <= (double)RT3_CURVE_MAX_SUBDIV * RT3_CURVE_MAX_SUBDIV)) return RT3_CURVE_MAX_SUBDIV;
    const int K = (int)std::ceil(std::sqrt(need));
    return K < 1 ? 1 : (K > RT3_CURVE_MAX_SUBDIV ? RT3_CURVE_MAX_SUBDIV : K);
}
// out_sub: (user segment, k | K << 16) per sub-segment; sub_first: first sub-segment of every user segment (+ end)
static void tessellate_curves(int basis, const float* cp, const int32_t* seg, int nseg, std::vector<float>& out_cp, std::vector<int32_t>& out_seg,
                              std::vector<float>& out_coef, std::vector<uint32_t>& out_sub, std::vector<uint32_t>& sub_first) {
    out_coef.resize((size_t)16 * nseg);
    sub_first.assign((size_t)nseg + 1, 0u);
    for (int s = 0; s < nseg; s++) {
        float coef[4][4];
        curve_coefficients(basis, cp + 4 * (size_t)seg[s], coef);
        memcpy(&out_coef[16 * (size_t)s], coef, sizeof(coef));
        sub_first[(size_t)s + 1] = sub_first[(size_t)s] + (uint32_t)curve_pieces(coef);
        RT3_REQUIRE(sub_first[(size_t)s + 1] < (1u << 27), RT3_ERR_INVALID, "curves_create: too many segments");
    }
    const size_t nsub = sub_first[(size_t)nseg];
    out_cp.resize(4 * (nsub + (size_t)nseg));
    out_seg.resize(nsub);
    out_sub.resize(2 * nsub);
    for (int s = 0; s < nseg; s++) {
        const float* coef = &out_coef[16 * (size_t)s];
        const uint32_t first = sub_first[(size_t)s], K = sub_first[(size_t)s + 1] - first;
        for (uint32_t k = 0; k <= K; k++) {
            const float u = (float)k / (float)K;
            for (int c = 0; c < 4; c++) out_cp[4 * ((size_t)first + (size_t)s + k) + (size_t)c] = ((coef[c] * u + coef[4 + c]) * u + coef[8 + c]) * u + coef[12 + c];   // Horner
            if (k < K) {
                out_seg[(size_t)first + k] = (int32_t)(first + (uint32_t)s + k);   // K + 1 points per segment
                out_sub[2 * ((size_t)first + k)] = (uint32_t)s;
                out_sub[2 * ((size_t)first + k) + 1] = k | (K << 16);
            }
        }
    }
}
static inline void curve_hit_to_internal(const std::vector<uint32_t>& sub_first, int32_t& prim, float& u) {
    const uint32_t first = sub_first[(size_t)prim];
    const int K = (int)(sub_first[(size_t)prim + 1] - first);
    const float f = u * (float)K;
    int k = (int)f;
    k = k > K - 1 ? K - 1 : (k < 0 ? 0 : k);
    u = f - (float)k;
    prim = (int32_t)first + k;
}

int rt3_curves_create(rt3_context_t c, int degree, const float* cp, int ncp, const int32_t* seg, int nseg, rt3_handle_t* blas) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && cp && seg && blas && ncp >= 2 && nseg > 0, RT3_ERR_INVALID, "curves_create: bad argument");
    RT3_REQUIRE(degree >= RT3_CURVE_LINEAR && degree <= RT3_CURVE_BEZIER, RT3_ERR_UNSUPPORTED,
                "curves_create: curve type must be 1 (linear), 2 / 3 (quadratic / cubic uniform B-spline), 4 (Catmull-Rom) or 5 (cubic Bezier)");
    for (int i = 0; i < nseg; i++) RT3_REQUIRE(seg[i] >= 0 && seg[i] + basis_table(degree).ncp <= ncp, RT3_ERR_INVALID, "curves_create: segment out of range");
    auto g = std::make_unique<Geometry>();
    g->type = PRIM_CURVE;
    std::vector<float> tcp, coef;
    std::vector<int32_t> tseg;
    std::vector<uint32_t> sub;
    if (degree > RT3_CURVE_LINEAR) {
        tessellate_curves(degree, cp, seg, nseg, tcp, tseg, coef, sub, g->sub_first);
        cp = tcp.data(); ncp = (int)(tcp.size() / 4); seg = tseg.data(); nseg = (int)tseg.size();
        g->sub.alloc(sub.size());
        h2d(g->sub.p, sub.data(), g->sub.bytes(), c->stream);
        g->curve_cubic = degree >= RT3_CURVE_CUBIC_BSPLINE ? 1u : 0u;
        g->poly.alloc(coef.size() / 4);   // the true curve of every user segment: normals are taken from it (cuda/curve.h:311-379)
        h2d(g->poly.p, coef.data(), g->poly.bytes(), c->stream);
    }
    g->nprims = (uint32_t)nseg;
    bounding_sphere(cp, 4, cp + 3, 4, (size_t)ncp, g->bsphere);   // (tessellated) control points with their radii: every round segment lies in the hull of its end spheres
    g->cr.alloc(ncp);
    g->seg.alloc(nseg);
    h2d(g->cr.p, cp, g->cr.bytes(), c->stream);
    h2d(g->seg.p, seg, g->seg.bytes(), c->stream);
    stream_sync(c->stream);
    *blas = finish_geometry(c, std::move(g));
    RT3_API_END
}

int rt3_texture_create(rt3_context_t c, const uint8_t* rgba8, int w, int h, int address_mode, int filter_mode, int* tex_id) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && rgba8 && tex_id && w > 0 && h > 0, RT3_ERR_INVALID, "texture_create: bad argument");
    RT3_REQUIRE(filter_mode == 0 || filter_mode == 1, RT3_ERR_UNSUPPORTED, "texture_create: filter_mode must be 0 (point) or 1 (bilinear)");
    RT3_REQUIRE(address_mode >= RT3_ADDRESS_WRAP && address_mode <= RT3_ADDRESS_BORDER, RT3_ERR_UNSUPPORTED, "texture_create: address mode unsupported");
    auto t = std::make_unique<TextureObj>();
    t->w = w; t->h = h; t->addr = address_mode; t->filt = filter_mode;
    t->px.alloc((size_t)w * h);
    h2d(t->px.p, rgba8, t->px.bytes(), c->stream);
    stream_sync(c->stream);
    c->textures.push_back(std::move(t));
    *tex_id = (int)c->textures.size() - 1;
    std::vector<TexDev> td(c->textures.size());
    for (size_t i = 0; i < td.size(); i++) td[i] = TexDev{c->textures[i]->px.p, c->textures[i]->w, c->textures[i]->h, c->textures[i]->addr, c->textures[i]->filt};
    c->d_tex.ensure(td.size());
    h2d(c->d_tex.p, td.data(), sizeof(TexDev) * td.size(), c->stream);
    stream_sync(c->stream);
    RT3_API_END
}

// ------------------------------------------------------------------------------------ instances
static int append_instance_impl(rt3_context_t c, rt3_handle_t blas, const float* xform, const float* keys, int nkeys, float t0, float t1, int* iid) {
    RT3_REQUIRE(c && xform && iid, RT3_ERR_INVALID, "append_instance: null argument");
    RT3_REQUIRE(blas >= 1 && blas <= c->geoms.size(), RT3_ERR_INVALID, "append_instance: invalid BLAS handle");
    InstanceHost in;
    in.blas = (uint32_t)(blas - 1);
    memcpy(in.xform, xform, sizeof(in.xform));
    if (keys) {
        RT3_REQUIRE(nkeys >= 2 && t1 > t0, RT3_ERR_INVALID, "append_animated_instance: need >= 2 keys and t_end > t_begin");
        RT3_REQUIRE(nkeys < 65536, RT3_ERR_INVALID, "append_animated_instance: at most 65535 keys (16-bit field of the instance record)");
        in.nkeys = (uint32_t)nkeys;
        in.keys.assign(keys, keys + 12 * (size_t)nkeys);
        in.t0 = t0; in.t1 = t1;
    }
    in.hg = HitGroupDev{{0, 0, 0}, {0.8f, 0.8f, 0.8f}, -1, in.t1, {1, 1}, {0, 1}, {0, 0}, 0u, 0u};
    c->inst.push_back(std::move(in));
    c->built = false;
    c->hitgroups_dirty = true;
    *iid = (int)c->inst.size() - 1;  // instanceId = sbtOffset = index (cuda_accel.h:78-79)
    return RT3_OK;
}

int rt3_accel_append_instance(rt3_context_t c, rt3_handle_t blas, const float xform[12], int* instance_id) {
    RT3_API_BEGIN
    use_device(c);
    append_instance_impl(c, blas, xform, nullptr, 0, 0.0f, 1.0f, instance_id);
    RT3_API_END
}

int rt3_accel_append_animated_instance(rt3_context_t c, rt3_handle_t blas, const float* keys, int nkeys, float t_begin, float t_end,
                                       const float static_xform[12], int* instance_id) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(keys, RT3_ERR_INVALID, "append_animated_instance: null keys");
    append_instance_impl(c, blas, static_xform, keys, nkeys, t_begin, t_end, instance_id);
    RT3_API_END
}

// a transform whose triangles may be intersected in world space: finite and not (numerically) singular — a collapsed instance
// keeps its object-space semantics, where the inverse is what it is.  Evaluated in double from the float entries, in this
// order, so that the oracle (rt3o_accel_build) takes the same decision
static bool flattenable(const float* m) {
    for (int k = 0; k < 12; k++) if (!std::isfinite(m[k])) return false;
    const double det = (double)m[0] * ((double)m[5] * (double)m[10] - (double)m[6] * (double)m[9])
                     - (double)m[1] * ((double)m[4] * (double)m[10] - (double)m[6] * (double)m[8])
                     + (double)m[2] * ((double)m[4] * (double)m[9] - (double)m[5] * (double)m[8]);
    return std::fabs(det) >= 1e-30;
}

int rt3_accel_build(rt3_context_t c) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c, RT3_ERR_INVALID, "accel_build: null context");
    RT3_REQUIRE(!c->inst.empty(), RT3_ERR_STATE, "accel_build: no instances");
    const uint32_t ni = (uint32_t)c->inst.size();
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    // ---- which instances go into the merged world BLAS, which stay in the TLAS
    // Identity instances (all the reference creates) always qualify.  Transformed static instances of plain triangle meshes
    // are flattened into it as well ("flatten", default on): instancing saves memory the B200 has (48 B per triangle + nodes,
    // ~100 B in all), and costs a ray transform, three divisions and a second BVH root per instance visit.  The merged BLAS is
    // one acceleration structure and stays below the 2^27-primitive limit of build_bvh8: if flattening would exceed it, only
    // the identity instances are merged; if those alone exceed it, nothing is.
    std::vector<uint32_t> merged, tl;
    auto select = [&](bool flatten) {
        merged.clear(); tl.clear();
        uint64_t sum = 0;
        for (uint32_t i = 0; i < ni; i++) {
            const InstanceHost& in = c->inst[i];
            const bool identity = in.nkeys == 0 && memcmp(in.xform, ident, sizeof(ident)) == 0;
            if (c->opt_merge && in.nkeys == 0 && (identity || (flatten && flattenable(in.xform))) && c->geoms[in.blas]->type == PRIM_TRI) { merged.push_back(i); sum += c->geoms[in.blas]->nprims; }
            else tl.push_back(i);
        }
        return sum < (1ull << 27);
    };
    if (!(c->opt_flatten && select(true)) && !select(false)) {
        merged.clear();
        tl.clear();
        for (uint32_t i = 0; i < ni; i++) tl.push_back(i);
    }
    c->flattened = 0;
    for (uint32_t i : merged) c->flattened += memcmp(c->inst[i].xform, ident, sizeof(ident)) != 0 ? 1u : 0u;
    c->has_merged = !merged.empty();
    c->single_level = c->has_merged && tl.empty();
    {   // two passes pay when the merged BLAS carries real work; a floor quad beside animated instances does not
        uint64_t sum = 0;
        for (uint32_t i : merged) sum += c->geoms[c->inst[i].blas]->nprims;
        c->split = c->has_merged && !tl.empty() && (c->opt_split >= 2 || (c->opt_split == 1 && sum >= 1024));
    }
    for (uint32_t i : tl) ensure_blas(c, c->geoms[c->inst[i].blas].get());
    BlasBounds merged_bounds{};
    if (c->has_merged) try {
        std::vector<MergedRange> ranges;
        uint32_t total = 0;
        for (uint32_t i : merged) {
            const Geometry& g = *c->geoms[c->inst[i].blas];
            MergedRange r{total, i, g.verts.p, g.idx.p, memcmp(c->inst[i].xform, ident, sizeof(ident)) != 0 ? 1u : 0u, {}};
            memcpy(r.xf, c->inst[i].xform, sizeof(r.xf));
            ranges.push_back(r);
            total += g.nprims;
        }
        DevBuf<MergedRange> d_ranges(ranges.size());
        h2d(d_ranges.p, ranges.data(), sizeof(MergedRange) * ranges.size(), c->stream);
        {
            DevBuf<float4> lo(total), hi(total);
            RT3_LAUNCH_1D(k_merged_boxes, total, c->stream, (const MergedRange*)d_ranges.p, (uint32_t)ranges.size(), lo.p, hi.p);
            build_bvh8(lo.p, hi.p, total, c->stream, c->m_nodes, c->m_order, c->m_bvh, /*sah_host=*/false, /*ploc=*/c->opt_ploc != 0, RT3_LEAF_MAX, /*sah_collapse=*/c->opt_sah_collapse != 0);
        }
        c->m_prims.alloc(3 * (size_t)total);
        c->m_map.alloc(total);
        RT3_LAUNCH_1D(k_pack_merged, total, c->stream, (const MergedRange*)d_ranges.p, (uint32_t)ranges.size(), (const uint32_t*)c->m_order.p, c->m_prims.p, c->m_map.p);
        stream_sync(c->stream);
        // nodes + primitive records in ONE slab so that a single L2 access-policy window can pin them
        {
            const size_t nb = (c->m_nodes.bytes() + 255) & ~(size_t)255, pb = c->m_prims.bytes();
            c->m_slab.alloc(nb + pb);
            d2d(c->m_slab.p, c->m_nodes.p, c->m_nodes.bytes(), c->stream);
            d2d(c->m_slab.p + nb, c->m_prims.p, pb, c->stream);
            stream_sync(c->stream);
            c->m_nodes.release();
            c->m_prims.release();
            c->m_nodes_p = reinterpret_cast<Node8*>(c->m_slab.p);
            c->m_prims_p = reinterpret_cast<float4*>(c->m_slab.p + nb);
            set_l2_window(c);
        }
        for (int k = 0; k < 3; k++) { merged_bounds.lo[k] = c->m_bvh.lo[k]; merged_bounds.hi[k] = c->m_bvh.hi[k]; }
    } catch (const Error& e) {
        if (c->flattened == 0) throw;
        throw Error(e.code, std::string(e.what()) + " — while building the merged world BLAS with " + std::to_string(c->flattened) +
                                " flattened instance(s) (about 350 B of device memory per instanced triangle at the peak of the build); "
                                "rt3_set_option(ctx, \"flatten\", 0) keeps transformed instances as instances");
    } else {
        c->m_map.alloc(1);
    }
    // ---- BLAS table: one entry per geometry (+ one for the merged BLAS)
    const size_t ng = c->geoms.size();
    std::vector<BlasDev> bt(ng + 1);
    std::vector<BlasBounds> bb(ng + 1);
    for (size_t i = 0; i < ng; i++) {
        const Geometry& g = *c->geoms[i];
        bt[i] = BlasDev{g.nodes.p, g.prims.p, g.type, g.nprims, g.idx.p, g.normals.p, g.uvs.p, g.cr.p, g.seg.p, g.vkeys, g.verts.p, g.nv, g.sub.p, g.colors.p, g.poly.p, g.curve_cubic};
        for (int k = 0; k < 3; k++) { bb[i].lo[k] = g.bvh.lo[k]; bb[i].hi[k] = g.bvh.hi[k]; }
    }
    bt[ng] = BlasDev{c->m_nodes_p, c->m_prims_p, PRIM_TRI, c->m_bvh.num_prims, nullptr, nullptr, nullptr, nullptr, nullptr, 1u, nullptr};
    bb[ng] = merged_bounds;
    c->d_blas.alloc(bt.size());
    h2d(c->d_blas.p, bt.data(), sizeof(BlasDev) * bt.size(), c->stream);
    DevBuf<BlasBounds> d_bb(bb.size());
    h2d(d_bb.p, bb.data(), sizeof(BlasBounds) * bb.size(), c->stream);
    // ---- instance table: the real instances (shading reads them by hit id) + one pseudo-instance for the merged BLAS
    std::vector<InstanceDev> it(ni + 1);
    std::vector<float> keys, stat(12 * (size_t)(ni + 1));
    for (uint32_t i = 0; i < ni; i++) {
        const InstanceHost& in = c->inst[i];
        memset(&it[i], 0, sizeof(InstanceDev));
        it[i].blas = in.blas;
        it[i].nkeys = in.nkeys;
        it[i].identity = (in.nkeys == 0 && memcmp(in.xform, ident, sizeof(ident)) == 0) ? 1u : 0u;
        it[i].key_offset = (uint32_t)keys.size();
        it[i].t0 = in.t0;
        keys.insert(keys.end(), in.keys.begin(), in.keys.end());
        memcpy(&stat[12 * (size_t)i], in.xform, sizeof(in.xform));
    }
    memset(&it[ni], 0, sizeof(InstanceDev));
    it[ni].blas = (uint32_t)ng;
    it[ni].identity = 2u;
    memcpy(&stat[12 * (size_t)ni], ident, sizeof(ident));
    c->d_inst.alloc(ni + 1);
    c->d_static.alloc(stat.size());
    c->d_keys.alloc(keys.size() ? keys.size() : 12);
    h2d(c->d_inst.p, it.data(), sizeof(InstanceDev) * (ni + 1), c->stream);
    h2d(c->d_static.p, stat.data(), sizeof(float) * stat.size(), c->stream);
    h2d(c->d_keys.p, keys.data(), sizeof(float) * keys.size(), c->stream);
    RT3_LAUNCH_1D(k_invert_static, ni + 1, c->stream, (const float*)c->d_static.p, c->d_inst.p);
    {   // world-space bounding spheres of the instances (TravScene::inst_bsphere)
        std::vector<float> bs(8 * (size_t)(ni + 1), 0.0f);
        auto xform = [](const float* m, const double p[3], double out[3]) { for (int k = 0; k < 3; k++) out[k] = (double)m[4 * k] * p[0] + (double)m[4 * k + 1] * p[1] + (double)m[4 * k + 2] * p[2] + (double)m[4 * k + 3]; };
        for (uint32_t i = 0; i <= ni; i++) {
            float* o = &bs[8 * (size_t)i];
            o[3] = -1.0f;
            if (i == ni) continue;   // the merged pseudo-instance
            const InstanceHost& in = c->inst[i];
            const Geometry& g = *c->geoms[in.blas];
            const bool identity = in.nkeys == 0 && memcmp(in.xform, ident, sizeof(ident)) == 0;
            if (identity || in.nkeys > 2 || !c->opt_bsphere_cull || !(g.bsphere[3] >= 0.0f)) continue;
            const double c0[3] = {g.bsphere[0], g.bsphere[1], g.bsphere[2]};
            double w0[3], w1[3], t[3];
            double scale = spectral_norm(in.xform);
            if (in.nkeys == 2) {
                xform(in.keys.data(), c0, t); xform(in.xform, t, w0);
                xform(in.keys.data() + 12, c0, t); xform(in.xform, t, w1);
                const double k0 = spectral_norm(in.keys.data()), k1 = spectral_norm(in.keys.data() + 12);
                scale *= k0 > k1 ? k0 : k1;
            } else {
                xform(in.xform, c0, w0);
                for (int k = 0; k < 3; k++) w1[k] = w0[k];
            }
            const double cn = std::sqrt(w0[0] * w0[0] + w0[1] * w0[1] + w0[2] * w0[2]) + std::sqrt(w1[0] * w1[0] + w1[1] * w1[1] + w1[2] * w1[2]);
            o[0] = (float)w0[0]; o[1] = (float)w0[1]; o[2] = (float)w0[2];
            o[3] = (float)((double)g.bsphere[3] * scale * (1.0 + 1e-4) + 1e-6 * cn + 1e-30);   // padded: float centres, float ray arithmetic
            o[4] = (float)(w1[0] - w0[0]); o[5] = (float)(w1[1] - w0[1]); o[6] = (float)(w1[2] - w0[2]);
            o[7] = in.nkeys == 2 ? 1.0f / (in.t1 - in.t0) : 0.0f;
        }
        c->d_bsphere.alloc(2 * (size_t)(ni + 1));
        h2d(c->d_bsphere.p, bs.data(), sizeof(float) * bs.size(), c->stream);
        stream_sync(c->stream);
    }
    // ---- TLAS over the remaining instances (+ the merged pseudo-instance), unless the scene is single-level
    if (!c->single_level) {
        DevBuf<float4> lo(ni + 1), hi(ni + 1);
        RT3_LAUNCH_1D(k_instance_boxes, ni + 1, c->stream, (const InstanceDev*)c->d_inst.p, (const float*)c->d_static.p, (const BlasBounds*)d_bb.p,
                      (const BlasDev*)c->d_blas.p, (const float*)c->d_keys.p, c->opt_tlas_refine, lo.p, hi.p);
        std::vector<uint32_t> sel(tl);
        if (c->has_merged && !c->split) sel.push_back(ni);
        // gather the selected boxes, build, then translate the TLAS leaf order back to instance ids
        const uint32_t ns = (uint32_t)sel.size();
        DevBuf<float4> slo(ns), shi(ns);
        for (uint32_t k = 0; k < ns; k++) {
            d2d(slo.p + k, lo.p + sel[k], sizeof(float4), c->stream);
            d2d(shi.p + k, hi.p + sel[k], sizeof(float4), c->stream);
        }
        build_bvh8(slo.p, shi.p, ns, c->stream, c->tlas_nodes, c->tlas_order, c->tlas, /*sah_host=*/c->opt_tlas_sah && ns <= (1u << 16), /*ploc=*/false, /*leaf_max=*/c->opt_tlas_leaf, /*sah_collapse=*/c->opt_sah_collapse != 0);
        std::vector<uint32_t> order(ns);
        d2h(order.data(), c->tlas_order.p, sizeof(uint32_t) * ns, c->stream);
        stream_sync(c->stream);
        for (uint32_t k = 0; k < ns; k++) order[k] = sel[order[k]];
        h2d(c->tlas_order.p, order.data(), sizeof(uint32_t) * ns, c->stream);
        stream_sync(c->stream);
    }
    c->hitgroups_dirty = true;
    upload_hitgroups(c);
    c->has_subdiv_curves = false;
    for (const InstanceHost& in : c->inst) c->has_subdiv_curves = c->has_subdiv_curves || c->geoms[in.blas]->spline();
    c->built = true;
    RT3_API_END
}

// ------------------------------------------------------------------------------------ shading records
int rt3_scene_set_hitgroup(rt3_context_t c, int id, const float e[3], const float d[3], int tex) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && e && d, RT3_ERR_INVALID, "set_hitgroup: null argument");
    RT3_REQUIRE(id >= 0 && id < (int)c->inst.size(), RT3_ERR_INVALID, "set_hitgroup: instance id out of range");
    RT3_REQUIRE(tex >= -1 && tex < (int)c->textures.size(), RT3_ERR_INVALID, "set_hitgroup: texture id out of range");
    HitGroupDev& hg = c->inst[id].hg;
    for (int k = 0; k < 3; k++) { hg.emission[k] = e[k]; hg.diffuse[k] = d[k]; }
    hg.tex = tex;
    c->hitgroups_dirty = true;
    RT3_API_END
}

// MaterialData::Texture::texcoord_scale / _rotation (sin, cos) / _offset of the SDK's sampleTexture (cuda/LocalShading.h:37-54)
int rt3_scene_set_texture_transform(rt3_context_t c, int id, const float scale[2], const float rotation[2], const float offset[2]) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && scale && rotation && offset, RT3_ERR_INVALID, "set_texture_transform: null argument");
    RT3_REQUIRE(id >= 0 && id < (int)c->inst.size(), RT3_ERR_INVALID, "set_texture_transform: instance id out of range");
    HitGroupDev& hg = c->inst[id].hg;
    for (int k = 0; k < 2; k++) { hg.tex_scale[k] = scale[k]; hg.tex_rot[k] = rotation[k]; hg.tex_off[k] = offset[k]; }
    hg.has_xf = 1u;
    c->hitgroups_dirty = true;
    RT3_API_END
}

int rt3_scene_set_lights(rt3_context_t c, const void* lights68, int n) {
    RT3_API_BEGIN
    use_device(c);
    static_assert(sizeof(Light) == 68, "rendertoy3o::Light is 68 bytes");
    RT3_REQUIRE(c && lights68 && n > 0, RT3_ERR_INVALID, "set_lights: at least one light is required (Q17)");
    c->d_lights.alloc(n);
    h2d(c->d_lights.p, lights68, sizeof(Light) * (size_t)n, c->stream);
    // running sum of the power-sampler weights (mode 2); sequential fp32, same recipe as the kernels' light_power()
    std::vector<float> cdf((size_t)n);
    const Light* L = static_cast<const Light*>(lights68);
    float run = 0.0f;
    for (int k = 0; k < n; k++) {
        Light l;
        memcpy(&l, reinterpret_cast<const char*>(L) + sizeof(Light) * (size_t)k, sizeof(l));
        run = run + (l.emission[0] * 0.30f + l.emission[1] * 0.59f + l.emission[2] * 0.11f) * l.area;
        cdf[(size_t)k] = run;
    }
    c->d_light_cdf.alloc(n);
    h2d(c->d_light_cdf.p, cdf.data(), sizeof(float) * (size_t)n, c->stream);
    stream_sync(c->stream);
    c->nlights = (uint32_t)n;
    RT3_API_END
}

// host helpers: pure fp32 arithmetic, no device work (nvcc passes -ffp-contract=off to the host compiler)
int rt3_light_make(const float e[3], const float v0[3], const float v1[3], const float v2[3], void* out) {  // Light ctor, src/light.h:24-30
    RT3_API_BEGIN
    RT3_REQUIRE(e && v0 && v1 && v2 && out, RT3_ERR_INVALID, "light_make: null argument");
    Light l;
    l.type = 0;
    const float a[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, b[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
    float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    l.area = 0.5f * len;
    const float inv = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int k = 0; k < 3; k++) { l.emission[k] = e[k]; l.v0[k] = v0[k]; l.v1[k] = v1[k]; l.v2[k] = v2[k]; l.normal[k] = n[k] * inv; }
    memcpy(out, &l, sizeof(l));
    RT3_API_END
}

int rt3_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy, float aspect, float U[3], float V[3], float W[3]) {  // sutil/Camera.cpp:34-45
    RT3_API_BEGIN
    RT3_REQUIRE(eye && lookat && up && U && V && W, RT3_ERR_INVALID, "camera_uvw: null argument");
    for (int k = 0; k < 3; k++) W[k] = lookat[k] - eye[k];
    const float wlen = sqrtf(W[0] * W[0] + W[1] * W[1] + W[2] * W[2]);
    float u[3] = {W[1] * up[2] - W[2] * up[1], W[2] * up[0] - W[0] * up[2], W[0] * up[1] - W[1] * up[0]};
    float inv = 1.0f / sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int k = 0; k < 3; k++) u[k] = u[k] * inv;
    float v[3] = {u[1] * W[2] - u[2] * W[1], u[2] * W[0] - u[0] * W[2], u[0] * W[1] - u[1] * W[0]};
    inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int k = 0; k < 3; k++) v[k] = v[k] * inv;
    const float vlen = wlen * tanf(0.5f * fovy * 3.14159265358979323846f / 180.0f);
    const float ulen = vlen * aspect;
    for (int k = 0; k < 3; k++) { V[k] = v[k] * vlen; U[k] = u[k] * ulen; }
    RT3_API_END
}

// ------------------------------------------------------------------------------------ the hot path
int rt3_launch_subframe(rt3_context_t c, const rt3_render_settings* rs) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && rs, RT3_ERR_INVALID, "launch_subframe: null argument");
    RT3_REQUIRE(c->built, RT3_ERR_STATE, "launch_subframe: rt3_accel_build has not been called");
    RT3_REQUIRE(c->nlights > 0, RT3_ERR_STATE, "launch_subframe: no lights set (Q17)");
    RT3_REQUIRE(rs->width > 0 && rs->height > 0 && rs->samples_per_launch > 0, RT3_ERR_INVALID, "launch_subframe: bad film settings");
    RT3_REQUIRE(rs->mode >= 0 && rs->mode <= 2, RT3_ERR_UNSUPPORTED, "launch_subframe: mode must be 0 (REFERENCE_FAITHFUL), 1 (CORRECTED) or 2 (CORRECTED + power light sampler)");
    const uint64_t paths64 = (uint64_t)rs->width * rs->height * rs->samples_per_launch;
    RT3_REQUIRE(paths64 < 0xfffffff0ull, RT3_ERR_INVALID, "launch_subframe: too many paths per launch");
    const uint32_t P = (uint32_t)paths64;
    upload_hitgroups(c);
    ensure_film(c, rs->width, rs->height);

    FrameParams f;
    f.width = rs->width; f.height = rs->height; f.spl = rs->samples_per_launch; f.subframe = rs->subframe_index;
    for (int k = 0; k < 3; k++) { f.eye[k] = rs->eye[k]; f.U[k] = rs->U[k]; f.V[k] = rs->V[k]; f.W[k] = rs->W[k]; f.miss[k] = rs->miss_color[k]; }
    f.max_depth = rs->max_depth; f.accum_mode = rs->accum_mode; f.mode = rs->mode;
    f.lights = c->d_lights.p; f.nlights = c->nlights; f.light_cdf = c->d_light_cdf.p; f.tex = c->d_tex.p;
    f.path_base = 0;
    const TravScene sc = c->trav_scene();

    const uint32_t M = MAX_DEPTH_SLOTS;
    const bool timing = c->opt_timing != 0;
    const bool unbounded = rs->max_depth <= 0;
    const uint32_t depth_limit = unbounded ? M - 2 : (uint32_t)rs->max_depth;
    RT3_REQUIRE(depth_limit <= M - 2, RT3_ERR_INVALID, "launch_subframe: max_depth too large");
#ifdef RT3_EMULATE
    const bool overlap = false;
#else
    const bool overlap = c->opt_overlap != 0 && !timing && !unbounded;
#endif
    // Schedule.  A subframe is one dependency chain generate -> {extend -> shade -> connect} x depth -> resolve in which
    // only connect(d) and extend(d+1) are independent; every kernel is a persistent grid whose tail (the longest rays)
    // leaves SMs idle, and at 1080p the tails cost ~10 % (the same kernels run 10 % faster per ray at 4K).  With
    // "overlap" (a) connect(d) runs on an auxiliary stream beside extend(d+1), and (b) the paths are split into two
    // halves issued as two independent chains on their own stream pairs, so that there is always another chain's
    // kernel to fill a tail.  Paths are independent and `result` is indexed by path id: the image is unchanged.
    const int nchains = (overlap && c->opt_overlap >= 2 && P >= (1u << 20)) ? 2 : 1;
    // consecutive subframes alternate between two slots (pools + streams) unless something in this launch needs the host
    const bool pipelined = overlap && nchains == 1 && c->opt_pipeline != 0;
    const int slot = pipelined ? (c->pipe_slot = (c->pipe_slot + 1) % RT3_PIPE_SLOTS) : 0;
    if (pipelined) for (int sl = 0; sl < RT3_PIPE_SLOTS; sl++) ensure_pools(c, sl, P);   // all at once: no allocation in the middle of a sequence
    else ensure_pools(c, slot, P);
    rt3_context::Pools& pool = c->pool[slot];
    const Stream s_main = pipelined ? c->pipe_stream[slot][0] : c->stream, s_aux = pipelined ? c->pipe_stream[slot][1] : c->stream2;
    if (pipelined && event_recorded(pool.done)) stream_wait(s_main, pool.done);   // the slot's previous subframe has been resolved
    c->shared_gpu = false;
#ifndef RT3_EMULATE
    if (pipelined)   // is an earlier subframe still running?  (a caller that synchronises after every subframe never sees one)
        for (int sl = 0; sl < RT3_PIPE_SLOTS; sl++)
            if (sl != slot && event_recorded(c->pool[sl].done) && cudaEventQuery(c->pool[sl].done.e) == cudaErrorNotReady) c->shared_gpu = true;
#endif
    struct Chain {
        Stream main = 0, aux = 0;
        uint32_t base = 0, count = 0;
        uint32_t* cnt = nullptr;   // [0,M): rays per depth  [M,2M): shadow rays per depth  [2M,3M): extend fetch  [3M,4M): connect fetch  [4M,6M): the same for pass 2 of a split traversal
        int cur = 0, connect_pending = -1;
        Event* ev_shade = nullptr; Event* ev_connect = nullptr;
        Queues q;
    } ch[2];
    dev_memset(pool.counters.p, 0, pool.counters.bytes(), s_main);
    for (int k = 0; k < nchains; k++) {
        Chain& h = ch[k];
        h.main = k == 0 ? s_main : c->stream3;
        h.aux = k == 0 ? s_aux : c->stream4;
        const uint32_t half = (P / 2u) & ~31u;
        h.base = k == 0 ? 0u : half;
        h.count = nchains == 1 ? P : (k == 0 ? half : P - half);
        h.cnt = pool.counters.p + (size_t)k * 6 * M;
        h.ev_shade = c->ev_shade[pipelined ? slot : k]; h.ev_connect = c->ev_connect[pipelined ? slot : k];
    }
    auto bind = [&](Chain& h, uint32_t depth) {
        Queues& q = h.q;
        const size_t o = h.base;
        q.ray0 = pool.ray[h.cur][0].p + o; q.ray1 = pool.ray[h.cur][1].p + o; q.ray2 = pool.ray[h.cur][2].p + o;
        q.st0 = pool.st[h.cur][0].p + o; q.st1 = pool.st[h.cur][1].p + o;
        q.nray0 = pool.ray[h.cur ^ 1][0].p + o; q.nray1 = pool.ray[h.cur ^ 1][1].p + o; q.nray2 = pool.ray[h.cur ^ 1][2].p + o;
        q.nst0 = pool.st[h.cur ^ 1][0].p + o; q.nst1 = pool.st[h.cur ^ 1][1].p + o;
        q.hit0 = pool.hit0.p + o; q.hit_inst = pool.hit_inst.p + o;
        q.sh0 = pool.sh[0].p + o; q.sh1 = pool.sh[1].p + o; q.sh2 = pool.sh[2].p + o; q.sh3 = pool.sh[3].p + o;
        q.result = pool.result.p;
        q.n_cur = h.cnt + depth; q.n_next = h.cnt + depth + 1; q.n_shadow = h.cnt + M + depth;
    };

    Event ev[8];
    float acc_ms[4] = {0, 0, 0, 0};
    if (timing) event_record(ev[0], c->stream);
    if (nchains > 1) {  // the second chain starts behind everything issued so far on the context's stream (counter reset, the previous subframe)
        event_record(c->ev_fork, c->stream);
        stream_wait(c->stream3, c->ev_fork);
    }
    for (int k = 0; k < nchains; k++) {
        bind(ch[k], 0);
        FrameParams fk = f;
        fk.path_base = ch[k].base;
        RT3_LAUNCH_1D(k_generate, ch[k].count, ch[k].main, fk, ch[k].q);
    }
    if (timing) event_record(ev[1], c->stream);

    for (uint32_t depth = 0; depth < depth_limit; depth++) {
        Event e0, e1, e2, e3;
        if (timing) event_record(e0, c->stream);
        for (int k = 0; k < nchains; k++) {
            Chain& h = ch[k];
            bind(h, depth);
            TraverseArgs a;
            a.scene = sc;
            a.rays = RayPlanes{h.q.ray0, h.q.ray1, h.q.ray2, 1u};
            a.count_ptr = h.q.n_cur; a.count = 0; a.fetch = h.cnt + 2 * M + depth;
            a.hit0 = h.q.hit0; a.hit_inst = h.q.hit_inst; a.contrib = nullptr; a.result = nullptr;
            a.stat = c->d_stats.p + (depth == 0 ? 0 : 1);
            a.faithful = 0;
            launch_traverse<TRAV_EXTEND>(c, a, h.cnt + 4 * M + depth, h.main, depth == 0 ? h.count : 0u);
        }
        if (timing) event_record(e1, c->stream);
        for (int k = 0; k < nchains; k++) {
            Chain& h = ch[k];
            // shade reuses the shadow queue and adds to the same radiance slots as the previous connect of this chain
            if (overlap && h.connect_pending >= 0) { stream_wait(h.main, h.ev_connect[h.connect_pending]); h.connect_pending = -1; }
#ifdef RT3_EMULATE
            k_shade(f, sc, h.q);
#else
            k_shade<<<c->num_sms * RT3_SHADE_MIN_BLOCKS, 256, 0, h.main>>>(f, sc, h.q);
            RT3_CUDA(cudaGetLastError());
#endif
            count_launch();
        }
        if (timing) event_record(e2, c->stream);
        for (int k = 0; k < nchains; k++) {
            Chain& h = ch[k];
            if (overlap) {  // shadow rays of this bounce on the auxiliary stream, behind shade(depth)
                event_record(h.ev_shade[depth & 1u], h.main);
                stream_wait(h.aux, h.ev_shade[depth & 1u]);
            }
            TraverseArgs s;
            s.scene = sc;
            s.rays = RayPlanes{h.q.sh0, h.q.sh1, h.q.sh2, 1u};
            s.count_ptr = h.q.n_shadow; s.count = 0; s.fetch = h.cnt + 3 * M + depth;
            s.hit0 = nullptr; s.hit_inst = nullptr; s.contrib = h.q.sh3; s.result = h.q.result;
            s.stat = c->d_stats.p + 2;
            s.faithful = rs->mode == 0 ? 1u : 0u;
            launch_traverse<TRAV_CONNECT>(c, s, h.cnt + 5 * M + depth, overlap ? h.aux : h.main);
            if (overlap) { event_record(h.ev_connect[depth & 1u], h.aux); h.connect_pending = (int)(depth & 1u); }
            h.cur ^= 1;
        }
        if (timing) {
            event_record(e3, c->stream);
            stream_sync(c->stream);
            acc_ms[0] += event_ms(e0, e1); acc_ms[1] += event_ms(e1, e2); acc_ms[2] += event_ms(e2, e3);
        }
        if (unbounded) {  // reference semantics: paths end only by miss / Russian roulette (raygen.cu:48-72)
            uint32_t nnext = 0;
            d2h(&nnext, ch[0].q.n_next, sizeof(nnext), c->stream);
            stream_sync(c->stream);
            if (nnext == 0) break;
            if (depth + 1 == depth_limit) c->host_flags |= 2u;   // paths still alive where the depth slots end: reported, not silent
        }
    }
    for (int k = 0; k < nchains; k++) {  // join: resolve (context stream) waits for every chain's last kernels
        Chain& h = ch[k];
        if (overlap && h.connect_pending >= 0) stream_wait(c->stream, h.ev_connect[h.connect_pending]);
        if (k > 0) { event_record(c->ev_join, h.main); stream_wait(c->stream, c->ev_join); }
    }
    if (pipelined) { event_record(c->ev_pipe_join[slot], s_main); stream_wait(c->stream, c->ev_pipe_join[slot]); }
    if (timing) event_record(ev[2], c->stream);
    if (c->frame_copy_pending) stream_wait(c->stream, c->ev_frame_copied);   // resolve overwrites the frame a pending download reads
    RT3_LAUNCH_1D(k_resolve, rs->width * rs->height, c->stream, f, (const float4*)pool.result.p, c->accum.p, c->frame.p);
    event_record(pool.done, c->stream);   // (whatever stream this subframe ran on: a later pipelined one may take the slot)
    if (timing) {
        event_record(ev[3], c->stream);
        stream_sync(c->stream);
        c->ms[0] = event_ms(ev[0], ev[1]); c->ms[1] = acc_ms[0]; c->ms[2] = acc_ms[1]; c->ms[3] = acc_ms[2];
        c->ms[4] = event_ms(ev[2], ev[3]); c->ms[5] = event_ms(ev[0], ev[3]);
    }
    c->samples += paths64;
    RT3_API_END
}

int rt3_trace_device(rt3_context_t c, const void* d_rays, int n, int any_hit, void* d_hits) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && c->built, RT3_ERR_STATE, "trace: rt3_accel_build has not been called");
    RT3_REQUIRE(n >= 0 && (n == 0 || (d_rays && d_hits)), RT3_ERR_INVALID, "trace: bad argument");
    if (n == 0) return RT3_OK;
    upload_hitgroups(c);
    dev_memset(c->trace_fetch.p, 0, 2 * sizeof(uint32_t), c->stream);
    TraverseArgs a;
    a.scene = c->trav_scene();
    const float4* r = (const float4*)d_rays;
    a.rays = RayPlanes{r, r + 1, r + 2, 3u};
    a.count_ptr = nullptr; a.count = (uint32_t)n; a.fetch = c->trace_fetch.p;
    a.hit0 = (float4*)d_hits; a.hit_inst = nullptr; a.contrib = nullptr; a.result = nullptr; a.stat = nullptr; a.faithful = 0;
    if (any_hit) launch_traverse<TRAV_TRACE_ANY>(c, a, c->trace_fetch.p + 1, c->stream);
    else launch_traverse<TRAV_TRACE_CLOSEST>(c, a, c->trace_fetch.p + 1, c->stream);
    // spline curves: internal sub-segment hits -> (segment, u along the segment)
    if (c->has_subdiv_curves) RT3_LAUNCH_1D(k_curve_hits_to_user, (uint32_t)n, c->stream, a.scene, (float4*)d_hits);
    RT3_API_END
}

int rt3_trace(rt3_context_t c, const rt3_ray* rays, int n, int any_hit, rt3_hit* hits) {
    RT3_API_BEGIN
    use_device(c);
    static_assert(sizeof(rt3_ray) == 48 && sizeof(rt3_hit) == 32, "ABI record sizes");
    RT3_REQUIRE(c && c->built, RT3_ERR_STATE, "trace: rt3_accel_build has not been called");
    RT3_REQUIRE(n >= 0 && (n == 0 || (rays && hits)), RT3_ERR_INVALID, "trace: bad argument");
    if (n == 0) return RT3_OK;
    c->q_rays.ensure(3 * (size_t)n);
    c->q_hits.ensure(2 * (size_t)n);
    DevBuf<float4>&d_rays = c->q_rays, &d_hits = c->q_hits;
    h2d(d_rays.p, rays, sizeof(rt3_ray) * (size_t)n, c->stream);
    const int rc = rt3_trace_device(c, d_rays.p, n, any_hit, d_hits.p);
    if (rc != RT3_OK) return rc;
    d2h(hits, d_hits.p, sizeof(rt3_hit) * (size_t)n, c->stream);
    require_no_device_error(c, "trace");
    RT3_API_END
}

// getLocalGeometry (cuda/LocalGeometry.h:61-175) for a batch of rays and the hit records rt3_trace returned for them
int rt3_get_local_geometry(rt3_context_t c, const rt3_ray* rays, const rt3_hit* hits, int n, rt3_local_geometry* out) {
    RT3_API_BEGIN
    use_device(c);
    static_assert(sizeof(rt3_local_geometry) == 108, "ABI record size");
    RT3_REQUIRE(c && c->built, RT3_ERR_STATE, "get_local_geometry: rt3_accel_build has not been called");
    RT3_REQUIRE(n >= 0 && (n == 0 || (rays && hits && out)), RT3_ERR_INVALID, "get_local_geometry: bad argument");
    for (int i = 0; i < n; i++)
        RT3_REQUIRE(hits[i].prim < 0 || (hits[i].inst >= 0 && (size_t)hits[i].inst < c->inst.size() &&
                                        (uint32_t)hits[i].prim < c->geoms[c->inst[(size_t)hits[i].inst].blas]->user_prims()),
                    RT3_ERR_INVALID, "get_local_geometry: hit record does not belong to this scene");
    if (n == 0) return RT3_OK;
    upload_hitgroups(c);
    std::vector<rt3_hit> internal;
    const rt3_hit* user_hits = hits;
    if (c->has_subdiv_curves) {  // (segment, u) of a spline curve -> its linear sub-segment
        internal.assign(hits, hits + n);
        for (int i = 0; i < n; i++)
            if (internal[(size_t)i].prim >= 0) {
                const Geometry& g = *c->geoms[c->inst[(size_t)internal[(size_t)i].inst].blas];
                if (g.spline()) curve_hit_to_internal(g.sub_first, internal[(size_t)i].prim, internal[(size_t)i].u);
            }
        hits = internal.data();
    }
    c->q_rays.ensure(3 * (size_t)n);
    c->q_hits.ensure(2 * (size_t)n);
    c->q_out.ensure(27 * (size_t)n);
    DevBuf<float4>&d_rays = c->q_rays, &d_hits = c->q_hits;
    DevBuf<float>& d_out = c->q_out;
    h2d(d_rays.p, rays, sizeof(rt3_ray) * (size_t)n, c->stream);
    h2d(d_hits.p, hits, sizeof(rt3_hit) * (size_t)n, c->stream);
    RT3_LAUNCH_1D(k_local_geometry, (uint32_t)n, c->stream, c->trav_scene(), (const float4*)d_rays.p, (const float4*)d_hits.p, d_out.p);
    d2h(out, d_out.p, sizeof(rt3_local_geometry) * (size_t)n, c->stream);
    stream_sync(c->stream);
    if (c->has_subdiv_curves)  // UV.x is the u along the user's segment, not along the sub-segment
        for (int i = 0; i < n; i++)
            if (user_hits[i].prim >= 0 && c->geoms[c->inst[(size_t)user_hits[i].inst].blas]->spline()) out[i].UV[0] = user_hits[i].u;
    RT3_API_END
}

// ------------------------------------------------------------------------------------ results
int rt3_download_accum(rt3_context_t c, float* rgba) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && rgba && c->accum.p, RT3_ERR_STATE, "download_accum: nothing rendered");
    d2h(rgba, c->accum.p, c->accum.bytes(), c->stream);
    require_no_device_error(c, "download_accum");
    RT3_API_END
}
int rt3_download_frame(rt3_context_t c, uint8_t* rgba8) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && rgba8 && c->frame.p, RT3_ERR_STATE, "download_frame: nothing rendered");
    d2h(rgba8, c->frame.p, c->frame.bytes(), c->stream);
    require_no_device_error(c, "download_frame");
    RT3_API_END
}
// The frame of the subframes launched so far, copied on the context's copy stream: returns at once, the copy runs beside
// whatever is launched next (the next subframe's resolve waits for it), rt3_sync completes it and reports device errors
int rt3_download_frame_async(rt3_context_t c, uint8_t* rgba8_pinned) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && rgba8_pinned && c->frame.p, RT3_ERR_STATE, "download_frame_async: nothing rendered");
    event_record(c->ev_frame_ready, c->stream);
    stream_wait(c->stream_copy, c->ev_frame_ready);
    d2h(rgba8_pinned, c->frame.p, c->frame.bytes(), c->stream_copy);
    event_record(c->ev_frame_copied, c->stream_copy);
    c->frame_copy_pending = true;
    RT3_API_END
}
int rt3_accum_device_ptr(rt3_context_t c, void** p, uint64_t* n) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && p && n && c->accum.p, RT3_ERR_STATE, "accum_device_ptr: nothing rendered");
    stream_sync(c->stream);
    *p = c->accum.p;
    *n = 4ull * c->width * c->height;
    RT3_API_END
}
int rt3_clear_accum(rt3_context_t c) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c, RT3_ERR_INVALID, "clear_accum: null context");
    if (c->accum.p) dev_memset(c->accum.p, 0, c->accum.bytes(), c->stream);
    RT3_API_END
}
int rt3_finalize_accum(rt3_context_t c, uint32_t total_subframes) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && c->accum.p && total_subframes > 0, RT3_ERR_STATE, "finalize_accum: nothing rendered");
    if (c->frame_copy_pending) stream_wait(c->stream, c->ev_frame_copied);
    RT3_LAUNCH_1D(k_finalize, c->width * c->height, c->stream, c->accum.p, c->frame.p, 1.0f / (float)total_subframes);
    RT3_API_END
}

int rt3_get_stats(rt3_context_t c, rt3_stats* st) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && st, RT3_ERR_INVALID, "get_stats: null argument");
    unsigned long long s[4] = {0, 0, 0, 0};
    uint32_t fl[2] = {0, 0};
    d2h(s, c->d_stats.p, sizeof(s), c->stream);
    d2h(fl, c->d_flags.p, sizeof(fl), c->stream);
    stream_sync(c->stream);
    memset(st, 0, sizeof(*st));
    st->rays_primary = s[0]; st->rays_bounce = s[1]; st->rays_shadow = s[2];
    st->samples = c->samples;
    st->kernel_launches = g_launch_count;
    st->ms_generate = c->ms[0]; st->ms_extend = c->ms[1]; st->ms_shade = c->ms[2]; st->ms_connect = c->ms[3]; st->ms_resolve = c->ms[4]; st->ms_total = c->ms[5];
    st->error_flags = fl[0] | c->host_flags; st->max_stack_depth = fl[1];
    st->flattened_instances = c->flattened; st->traversal_passes = c->built && c->split ? 2u : 1u;
    RT3_API_END
}
int rt3_get_debug_counters(rt3_context_t c, uint32_t out[16]) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c && out, RT3_ERR_INVALID, "get_debug_counters: null argument");
    d2h(out, c->d_flags.p, 16 * sizeof(uint32_t), c->stream);
    stream_sync(c->stream);
    dev_memset(c->d_flags.p + 2, 0, 14 * sizeof(uint32_t), c->stream);
    RT3_API_END
}
int rt3_reset_stats(rt3_context_t c) {
    RT3_API_BEGIN
    use_device(c);
    RT3_REQUIRE(c, RT3_ERR_INVALID, "reset_stats: null context");
    stream_sync(c->stream);   // subframes in flight still count
    dev_memset(c->d_stats.p, 0, c->d_stats.bytes(), c->stream);
    dev_memset(c->d_flags.p, 0, 2 * sizeof(uint32_t), c->stream);   // error word + stack high-water mark
    c->host_flags = 0;
    c->samples = 0;
    stream_sync(c->stream);
    RT3_API_END
}

// ------------------------------------------------------------------------------------ multi-GPU (single process, one context per GPU)
int rt3_allreduce_accum(rt3_context_t* ctxs, int n, uint32_t total_subframes) {
    RT3_API_BEGIN
    RT3_REQUIRE(ctxs && n >= 1, RT3_ERR_INVALID, "allreduce_accum: bad argument");
#ifdef RT3_EMULATE
    throw Error(RT3_ERR_UNSUPPORTED, "allreduce_accum: not available in the kernel-logic simulator");
#else
    if (n > 1) {
        const NcclApi& nccl = nccl_api();
        RT3_REQUIRE(nccl.ok, RT3_ERR_NCCL, "allreduce_accum: libnccl.so.2 (ncclCommInitAll / ncclAllReduce / ncclGroupStart / ncclGroupEnd / ncclCommDestroy) not found");
        std::vector<int> devs(n);
        for (int i = 0; i < n; i++) { RT3_REQUIRE(ctxs[i] && ctxs[i]->accum.p, RT3_ERR_STATE, "allreduce_accum: context has no film"); devs[i] = ctxs[i]->device; }
        for (int i = 0; i < n; i++)
            RT3_REQUIRE(ctxs[i]->width == ctxs[0]->width && ctxs[i]->height == ctxs[0]->height, RT3_ERR_STATE, "allreduce_accum: films differ in size");
        // communicators are created once per (device list, rank) and live in the contexts
        bool cached = true;
        for (int i = 0; i < n; i++) cached = cached && ctxs[i]->nccl_comm && ctxs[i]->nccl_rank == i && ctxs[i]->nccl_devs == devs;
        if (!cached) {
            for (int i = 0; i < n; i++) drop_nccl_comm(ctxs[i]);
            std::vector<void*> comms(n, nullptr);
            RT3_REQUIRE(nccl.initAll(comms.data(), n, devs.data()) == 0, RT3_ERR_NCCL, "ncclCommInitAll failed");
            for (int i = 0; i < n; i++) { ctxs[i]->nccl_comm = comms[i]; ctxs[i]->nccl_devs = devs; ctxs[i]->nccl_rank = i; }
        }
        RT3_REQUIRE(nccl.groupStart() == 0, RT3_ERR_NCCL, "ncclGroupStart failed");
        int rc_all = 0;
        for (int i = 0; i < n && rc_all == 0; i++) {
            cudaSetDevice(ctxs[i]->device);
            const size_t count = 4ull * ctxs[i]->width * ctxs[i]->height;
            rc_all = nccl.allReduce(ctxs[i]->accum.p, ctxs[i]->accum.p, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, ctxs[i]->nccl_comm, ctxs[i]->stream);
        }
        const int rc_end = nccl.groupEnd();   // always closed, also after a failed call inside the group
        RT3_REQUIRE(rc_all == 0 && rc_end == 0, RT3_ERR_NCCL, "ncclAllReduce failed");
        for (int i = 0; i < n; i++) { RT3_CUDA(cudaSetDevice(ctxs[i]->device)); stream_sync(ctxs[i]->stream); }
    }
    for (int i = 0; i < n; i++) {
        RT3_CUDA(cudaSetDevice(ctxs[i]->device));
        const int rc = rt3_finalize_accum(ctxs[i], total_subframes);
        if (rc != RT3_OK) return rc;
        stream_sync(ctxs[i]->stream);
    }
#endif
    RT3_API_END
}

}  // extern "C"
