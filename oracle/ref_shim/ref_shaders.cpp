// ref_shaders.cpp — TEST INFRASTRUCTURE.  Host harness around the REFERENCE'S OWN device programs
// (src/shader/raygen.cu, closehit_radiance.cu, miss.cu, test.cu — compiled where they lie under /root/reference with
// the functional OptiX stand-in oracle/ref_shim/optix.h and linked into oracle/_ref/librt3ref.so by oracle/Makefile
// target `ref_shaders`).  One call renders a subframe the way optixLaunch(w, h, 1) would: every launch index runs
// __raygen__rg, which runs the reference's path loop, closest-hit / miss programs, Light::Sample, RNG and accumulate.
//
// What is NOT reference code here, because the reference has none: ray traversal (optixTraverse -> the
// oracle's brute-force closest-hit / occlusion query, rt3o_trace with accel = 0) and the texture unit (tex2D -> the
// oracle's texel fetch).  Everything downstream of a hit record is the reference's own arithmetic, so the images this
// library writes are the pins for oracle rows A2, A7, A8, A12-A16 (tests/test_reference_pins.py).
#include <src/shader/shader_common.h>
#include <src/light.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../rt3o.h"

extern "C" {
extern rendertoy3o::RenderSettings params;   // the weak object defined by every program file
void __raygen__rg();
}

namespace rt3shim {
Hooks hooks;
thread_local Lane lane;
}

namespace {
struct Mesh { std::vector<float3> vertices, normals; std::vector<int3> indices; std::vector<float2> texcoords; };
struct RefScene {
    rt3o_scene* oracle = nullptr;
    std::vector<Mesh> meshes;                       // one per instance (hit group i <-> instance i, cuda_scene.h:60-82)
    std::vector<rendertoy3o::HitGroupData> sbt;
    std::vector<rendertoy3o::Light> lights;
    rendertoy3o::MissData miss{};
};

rt3shim::Hit trace_hook(void* user, float3 o, float3 d, float tmin, float tmax, float time, int any_hit) {
    RefScene* s = static_cast<RefScene*>(user);
    rt3_ray r{};
    r.o[0] = o.x; r.o[1] = o.y; r.o[2] = o.z; r.tmin = tmin;
    r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z; r.tmax = tmax;
    r.time = time;
    rt3_hit h{};
    rt3o_trace(s->oracle, &r, 1, any_hit, &h, /*accel: brute force*/ 0, /*threads*/ 1);
    rt3shim::Hit out;
    out.hit = h.prim >= 0;
    out.t = h.t; out.u = h.u; out.v = h.v; out.prim = h.prim; out.inst = h.inst;
    return out;
}
float4 tex_hook(void* user, unsigned long long tex, float u, float v) {
    RefScene* s = static_cast<RefScene*>(user);
    float rgb[3] = {0, 0, 0};
    rt3o_kat_fetch_texture(s->oracle, (int)tex, u, v, rgb);
    return make_float4(rgb[0], rgb[1], rgb[2], 1.0f);
}
void* sbt_hook(void* user, int inst) { return &static_cast<RefScene*>(user)->sbt[(size_t)inst]; }
}  // namespace

extern "C" {

void* rt3ref_create(rt3o_scene* oracle_scene) {
    RefScene* s = new RefScene;
    s->oracle = oracle_scene;
    return s;
}
void rt3ref_destroy(void* p) { delete static_cast<RefScene*>(p); }

// the attribute arrays behind HitGroupData of instance `inst` (src/shader/shader_data.h:125-136)
int rt3ref_set_mesh(void* p, int inst, const float* vertices, int nv, const int* indices, int nt, const float* normals, const float* texcoords) {
    RefScene* s = static_cast<RefScene*>(p);
    if (inst < 0) return -1;
    if ((size_t)inst >= s->meshes.size()) s->meshes.resize((size_t)inst + 1);
    Mesh& m = s->meshes[(size_t)inst];
    m.vertices.resize((size_t)nv); m.normals.resize((size_t)nv); m.texcoords.resize((size_t)nv); m.indices.resize((size_t)nt);
    for (int i = 0; i < nv; ++i) {
        m.vertices[(size_t)i] = make_float3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
        m.normals[(size_t)i] = make_float3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        m.texcoords[(size_t)i] = make_float2(texcoords[2 * i], texcoords[2 * i + 1]);
    }
    for (int i = 0; i < nt; ++i) m.indices[(size_t)i] = make_int3(indices[3 * i], indices[3 * i + 1], indices[3 * i + 2]);
    return 0;
}
// the SBT (create_sbt, src/cuda/cuda_scene.h:60-82: hit group i <-> instance i) and the light array (buildLightSampler,
// src/wavefront.cpp:257-275: 68-byte rendertoy3o::Light records)
int rt3ref_finish(void* p, const float* emission /*[n][3]*/, const float* diffuse /*[n][3]*/, const int* tex_ids /*[n]*/, const void* lights68, int nlights) {
    RefScene* s = static_cast<RefScene*>(p);
    const size_t n = s->meshes.size();
    s->sbt.assign(n, rendertoy3o::HitGroupData{});
    for (size_t i = 0; i < n; ++i) {
        rendertoy3o::HitGroupData& h = s->sbt[i];
        h.emission_color = make_float3(emission[3 * i], emission[3 * i + 1], emission[3 * i + 2]);
        h.diffuse_color = make_float3(diffuse[3 * i], diffuse[3 * i + 1], diffuse[3 * i + 2]);
        h.vertices = s->meshes[i].vertices.data();
        h.indices = s->meshes[i].indices.data();
        h.normals = s->meshes[i].normals.data();
        h.texcoords = s->meshes[i].texcoords.data();
        h.hasTexture = tex_ids[i] >= 0;
        h.texture = (cudaTextureObject_t)(tex_ids[i] >= 0 ? tex_ids[i] : 0);
    }
    static_assert(sizeof(rendertoy3o::Light) == 68, "Light is the 68-byte record of src/light.h:13-22");
    const rendertoy3o::Light* L = static_cast<const rendertoy3o::Light*>(lights68);
    s->lights.assign(L, L + nlights);
    return 0;
}

// One subframe = optixLaunch(pipeline, stream, d_params, sizeof, &sbt, w, h, 1) (src/wavefront.cpp:203-222).
// accum [h][w][4] float in/out (read when subframe_index > 0), frame [h][w][4] u8 out.
int rt3ref_launch(void* p, const rt3_render_settings* rs, float* accum, uint8_t* frame, int nthreads) {
    RefScene* s = static_cast<RefScene*>(p);
    rt3shim::hooks.user = s;
    rt3shim::hooks.trace = trace_hook;
    rt3shim::hooks.tex2d = tex_hook;
    rt3shim::hooks.sbt_hit = sbt_hook;
    rt3shim::hooks.sbt_miss = &s->miss;
    params.film_settings.subframe_index = rs->subframe_index;
    params.film_settings.accum_buffer = reinterpret_cast<float4*>(accum);
    params.film_settings.frame_buffer = reinterpret_cast<uchar4*>(frame);
    params.film_settings.width = rs->width;
    params.film_settings.height = rs->height;
    params.film_settings.samples_per_launch = rs->samples_per_launch;
    params.camera_settings.eye = make_float3(rs->eye[0], rs->eye[1], rs->eye[2]);
    params.camera_settings.U = make_float3(rs->U[0], rs->U[1], rs->U[2]);
    params.camera_settings.V = make_float3(rs->V[0], rs->V[1], rs->V[2]);
    params.camera_settings.W = make_float3(rs->W[0], rs->W[1], rs->W[2]);
    params.light_settings.light_count = (unsigned int)s->lights.size();
    params.light_settings.lights = s->lights.data();
    params.handle = 1;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    std::atomic<unsigned int> next_row{0};
    auto work = [&]() {
        for (;;) {
            const unsigned int y = next_row.fetch_add(1);
            if (y >= rs->height) break;
            for (unsigned int x = 0; x < rs->width; ++x) {
                rt3shim::lane = rt3shim::Lane();
                rt3shim::lane.launch_index = make_uint3(x, y, 0);
                rt3shim::lane.launch_dims = make_uint3(rs->width, rs->height, 1);
                __raygen__rg();
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return 0;
}

}  // extern "C"
