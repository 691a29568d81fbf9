// rt3_host.hpp — C++ host-side mirror of rendertoy3o's device-scene operators over the rt3 C ABI.
//
// Class and method names follow the reference so that src/wavefront.cpp reads the same after the
// switch:
//   rt3host::Exception        <- rendertoy3o::Exception            (src/util/exception.h:28-60)
//   rt3host::Context          <- rendertoy3o::OptixContext         (src/cuda/optix_context.h:16-272)
//   rt3host::CUDAMesh         <- rendertoy3o::CUDAMesh             (src/cuda/cuda_mesh.h:8-185)
//   rt3host::CUDATexture      <- rendertoy3o::CUDATexture<uchar4>  (src/cuda/cuda_texture.h:13-89)
//   rt3host::CUDAAccel        <- rendertoy3o::CUDAAccel            (src/cuda/cuda_accel.h:15-162)
//   rt3host::CUDAScene        <- rendertoy3o::CUDAScene            (src/cuda/cuda_scene.h:12-184)
//   rt3host::RenderSettings   <- rendertoy3o::RenderSettings       (src/shader/shader_data.h:71-114)
//   rt3host::launchSubframe   <- launchSubframe                    (src/wavefront.cpp:203-222)
// Move-only RAII, exceptions on error (the C ABI itself never throws).
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt3.h"

namespace rt3host {

struct Exception : std::runtime_error {
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
inline void check(int rc, const char* call, const char* file, int line) {
    if (rc != RT3_OK) throw Exception(std::string("rt3 call '") + call + "' failed: " + rt3_last_error() + " (" + file + ":" + std::to_string(line) + ")");
}
#define RT3HOST_CHECK(call) ::rt3host::check((call), #call, __FILE__, __LINE__)

struct float3_ { float x, y, z; };

// host-side containers filled by the loader (reference: src/mesh.h:12-28, src/material.h:15-38)
struct Material {
    float3_ m_diffuse{0.8f, 0.8f, 0.8f};
    float3_ m_emissive{0, 0, 0};
    int m_diffuseTextureID = -1;
    // carried like the reference does (src/mesh.cpp:186-197) and, like there, not read by any closure (src/material.h:33-36)
    int m_emissiveTextureID = -1;
    float m_roughness = 0.0f;
    int m_roughnessTextureID = -1;
    float m_anisotropy = 0.0f;
    float m_ior = 1.0f;
    float m_transmittance = 0.0f;
    int m_normalTextureID = -1;
};
struct Mesh {
    unsigned int num_keys = 1;
    std::vector<std::vector<float>> vertices;   // [key][3*nv]
    std::vector<std::vector<float>> normals;    // [key][3*nv]
    std::vector<std::vector<float>> texcoords;  // [key][2*nv]
    std::vector<int32_t> indices;               // 3*nt
    Material material;
};
struct Texture {
    int width = 0, height = 0;
    std::vector<uint8_t> pixel;  // RGBA8, row 0 = image bottom (mesh.cpp:151-159 flips at load)
};

class Context {
    rt3_context_t _ctx{nullptr};
public:
    explicit Context(int device = 0) { RT3HOST_CHECK(rt3_context_create(device, &_ctx)); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ~Context() { rt3_context_destroy(_ctx); }
    rt3_context_t ctx() const { return _ctx; }
};

class CUDAMesh {
    rt3_handle_t _gas_handle{0};
public:
    CUDAMesh(const CUDAMesh&) = delete;
    CUDAMesh(CUDAMesh&& o) noexcept : _gas_handle(o._gas_handle) { o._gas_handle = 0; }
    CUDAMesh(rt3_context_t ctx, const Mesh& mesh) {
        // the reference uploads whatever the loader produced and its closest-hit program reads normals[] / texcoords[] for every
        // vertex (closehit_radiance.cu:71-75): a mesh whose .obj left some corner without vn / vt is out-of-bounds there (Q11)
        for (unsigned k = 0; k < mesh.num_keys; ++k)
            if (mesh.normals[k].size() != mesh.vertices[k].size() || mesh.texcoords[k].size() / 2 != mesh.vertices[k].size() / 3)
                throw Exception("CUDAMesh: mesh has vertices without a normal or a texcoord (both are required, Q11)");
        std::vector<float> keys;
        for (unsigned k = 0; k < mesh.num_keys; ++k) keys.insert(keys.end(), mesh.vertices[k].begin(), mesh.vertices[k].end());
        RT3HOST_CHECK(rt3_mesh_create(ctx, keys.data(), (int)mesh.num_keys, (int)(mesh.vertices[0].size() / 3), mesh.indices.data(),
                                      (int)(mesh.indices.size() / 3), mesh.normals[0].data(), mesh.texcoords[0].data(), &_gas_handle));
    }
    rt3_handle_t gas_handle() const { return _gas_handle; }
};

class CUDATexture {
    int _id{-1};
public:
    enum struct AddressMode : int32_t { Wrap = 0, Clamp = 1, Mirror = 2, Border = 3 };
    enum struct FilterMode : int32_t { Linear = 0, Point = 1 };  // values as in the reference (Q9: 0 ends up as point sampling)
    CUDATexture(rt3_context_t ctx, size_t width, size_t height, const void* data, AddressMode a, FilterMode f) {
        RT3HOST_CHECK(rt3_texture_create(ctx, static_cast<const uint8_t*>(data), (int)width, (int)height, (int)a, (int)f, &_id));
    }
    int texture_object() const { return _id; }
};

class CUDAAccel {
    rt3_context_t _ctx;
    size_t _n{0};
public:
    explicit CUDAAccel(rt3_context_t ctx) : _ctx(ctx) {}
    int append_instance(const CUDAMesh& mesh, const float transformation[12]) { return append_instance(mesh.gas_handle(), transformation); }
    int append_instance(rt3_handle_t handle, const float transformation[12]) {
        int id = -1;
        RT3HOST_CHECK(rt3_accel_append_instance(_ctx, handle, transformation, &id));
        ++_n;
        return id;
    }
    // motion_options {numKeys = motion_matrix.size(), timeBegin, timeEnd} as OptixMotionOptions in the reference
    int append_animated_instance(const CUDAMesh& mesh, const std::vector<std::array<float, 12>>& motion_matrix, float time_begin, float time_end,
                                 const float static_transformation[12]) {
        int id = -1;
        RT3HOST_CHECK(rt3_accel_append_animated_instance(_ctx, mesh.gas_handle(), motion_matrix.data()->data(), (int)motion_matrix.size(), time_begin,
                                                         time_end, static_transformation, &id));
        ++_n;
        return id;
    }
    void build() { RT3HOST_CHECK(rt3_accel_build(_ctx)); }
    size_t instance_size() const { return _n; }
};

struct RenderSettings : rt3_render_settings {
    RenderSettings(int w, int h, unsigned int spl) {
        std::memset(static_cast<rt3_render_settings*>(this), 0, sizeof(rt3_render_settings));
        width = (uint32_t)w; height = (uint32_t)h; samples_per_launch = spl; subframe_index = 0;
        max_depth = 0; mode = 0; accum_mode = 0;
        miss_color[0] = miss_color[1] = miss_color[2] = 0.01f;
    }
};

class CUDAScene {
    rt3_context_t _ctx;
    std::vector<CUDAMesh> _cuda_meshes;
    std::vector<CUDATexture> _cuda_textures;
    CUDAAccel _accel;
public:
    CUDAScene(const CUDAScene&) = delete;
    // same order of work as the reference ctor (cuda_scene.h:124-159) + buildLightSampler (wavefront.cpp:257-275)
    CUDAScene(Context& context, const std::vector<Mesh>& meshes, const std::vector<Texture>& textures) : _ctx(context.ctx()), _accel(context.ctx()) {
        for (const auto& mesh : meshes) _cuda_meshes.emplace_back(_ctx, mesh);
        for (const auto& t : textures)
            _cuda_textures.emplace_back(_ctx, (size_t)t.width, (size_t)t.height, t.pixel.data(), CUDATexture::AddressMode::Wrap, CUDATexture::FilterMode::Linear);
        const float transformation[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        std::vector<uint8_t> lights;
        int nlights = 0;
        for (size_t i = 0; i < meshes.size(); ++i) {
            const int id = _accel.append_instance(_cuda_meshes[i], transformation);
            const Material& m = meshes[i].material;
            const int tex = m.m_diffuseTextureID >= 0 ? _cuda_textures[(size_t)m.m_diffuseTextureID].texture_object() : -1;
            RT3HOST_CHECK(rt3_scene_set_hitgroup(_ctx, id, &m.m_emissive.x, &m.m_diffuse.x, tex));
            const float len = std::sqrt(m.m_emissive.x * m.m_emissive.x + m.m_emissive.y * m.m_emissive.y + m.m_emissive.z * m.m_emissive.z);
            if (len < 1e-5f) continue;
            const std::vector<float>& v = meshes[i].vertices[0];
            for (size_t t = 0; t + 2 < meshes[i].indices.size(); t += 3) {
                lights.resize(lights.size() + 68);
                RT3HOST_CHECK(rt3_light_make(&m.m_emissive.x, &v[3 * (size_t)meshes[i].indices[t]], &v[3 * (size_t)meshes[i].indices[t + 1]],
                                             &v[3 * (size_t)meshes[i].indices[t + 2]], lights.data() + lights.size() - 68));
                ++nlights;
            }
        }
        _accel.build();
        if (nlights > 0) RT3HOST_CHECK(rt3_scene_set_lights(_ctx, lights.data(), nlights));
    }
    const CUDAAccel& accel() const { return _accel; }
};

inline void launchSubframe(Context& ctx, const RenderSettings& params) {
    RT3HOST_CHECK(rt3_launch_subframe(ctx.ctx(), &params));
    RT3HOST_CHECK(rt3_sync(ctx.ctx()));  // CUDA_SYNC_CHECK, wavefront.cpp:221
}

}  // namespace rt3host
