// dump_ref_loader.cpp — TEST INFRASTRUCTURE.  Runs the REFERENCE'S OWN loadOBJ (src/mesh.cpp:37-210, compiled
// where it lies under /root/reference together with its vendored tinyobj + stb_image; nothing is copied) and
// writes the Mesh / Texture lists it returns in the canonical "RT3L" dump format (see loader_dump.hpp).
// Built by oracle/Makefile target `ref_loader` into oracle/_ref/ (git-ignored).
//   usage: dump_ref_loader out.bin key0.obj [key1.obj ...]
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <src/mesh.h>
#include "loader_dump.hpp"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s out.bin key0.obj [key1.obj ...]\n", argv[0]); return 2; }
    std::vector<std::string> paths(argv + 2, argv + argc);
    auto [meshes, textures] = rendertoy3o::loadOBJ(paths);
    rt3dump::Writer w(argv[1]);
    w.header((uint32_t)meshes.size(), (uint32_t)textures.size());
    for (const auto& m : meshes) {
        const uint32_t nv = (uint32_t)m.vertices[0].size();
        w.mesh_begin(m.num_keys, nv, (uint32_t)m.indices.size());
        for (unsigned k = 0; k < m.num_keys; ++k) {
            w.count((uint32_t)m.vertices[k].size());  w.floats(&m.vertices[k][0].x, 3 * m.vertices[k].size());
            w.count((uint32_t)m.normals[k].size());   w.floats(m.normals[k].empty() ? nullptr : &m.normals[k][0].x, 3 * m.normals[k].size());
            w.count((uint32_t)m.texcoords[k].size()); w.floats(m.texcoords[k].empty() ? nullptr : &m.texcoords[k][0].x, 2 * m.texcoords[k].size());
        }
        w.ints(m.indices.empty() ? nullptr : &m.indices[0].x, 3 * m.indices.size());
        const auto& a = m.material;
        const float mat_f[10] = {a.m_diffuse.x, a.m_diffuse.y, a.m_diffuse.z, a.m_emissive.x, a.m_emissive.y, a.m_emissive.z,
                                 a.m_roughness, a.m_anisotropy, a.m_ior, a.m_transmittance};
        const int mat_i[4] = {a.m_diffuseTextureID, a.m_emissiveTextureID, a.m_roughnessTextureID, a.m_normalTextureID};
        w.floats(mat_f, 10);
        w.ints(mat_i, 4);
    }
    for (const auto& t : textures) {
        w.count((uint32_t)t.resolution.x);
        w.count((uint32_t)t.resolution.y);
        w.bytes(t.pixel.data(), 4 * t.pixel.size());
    }
    return 0;
}
