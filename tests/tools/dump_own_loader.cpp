// dump_own_loader.cpp — TEST INFRASTRUCTURE.  Runs this repo's loadOBJ (rendertoy3c_b200/host/obj_loader.hpp) and
// writes its result in the canonical "RT3L" dump (oracle/ref_loader/loader_dump.hpp), to be compared byte for byte
// with the dump of the reference's own loader (oracle/_ref/dump_ref_loader, tests/golden/loader/*.rt3l).
//   usage: dump_own_loader out.bin key0.obj [key1.obj ...]       exit 4 = the loader threw (message on stderr)
#include <cstdio>
#include <string>
#include <vector>
#include "../../rendertoy3c_b200/host/obj_loader.hpp"
#include "../../oracle/ref_loader/loader_dump.hpp"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s out.bin key0.obj [key1.obj ...]\n", argv[0]); return 2; }
    std::vector<std::string> paths(argv + 2, argv + argc);
    std::vector<rt3host::Mesh> meshes;
    std::vector<rt3host::Texture> textures;
    try {
        rt3host::loadOBJ(paths, meshes, textures);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 4;
    }
    rt3dump::Writer w(argv[1]);
    w.header((uint32_t)meshes.size(), (uint32_t)textures.size());
    for (const auto& m : meshes) {
        w.mesh_begin(m.num_keys, (uint32_t)(m.vertices[0].size() / 3), (uint32_t)(m.indices.size() / 3));
        for (unsigned k = 0; k < m.num_keys; ++k) {
            w.count((uint32_t)(m.vertices[k].size() / 3));  w.floats(m.vertices[k].data(), m.vertices[k].size());
            w.count((uint32_t)(m.normals[k].size() / 3));   w.floats(m.normals[k].data(), m.normals[k].size());
            w.count((uint32_t)(m.texcoords[k].size() / 2)); w.floats(m.texcoords[k].data(), m.texcoords[k].size());
        }
        w.ints(m.indices.data(), m.indices.size());
        const auto& a = m.material;
        const float mat_f[10] = {a.m_diffuse.x, a.m_diffuse.y, a.m_diffuse.z, a.m_emissive.x, a.m_emissive.y, a.m_emissive.z,
                                 a.m_roughness, a.m_anisotropy, a.m_ior, a.m_transmittance};
        const int mat_i[4] = {a.m_diffuseTextureID, a.m_emissiveTextureID, a.m_roughnessTextureID, a.m_normalTextureID};
        w.floats(mat_f, 10);
        w.ints(mat_i, 4);
    }
    for (const auto& t : textures) {
        w.count((uint32_t)t.width);
        w.count((uint32_t)t.height);
        w.bytes(t.pixel.data(), t.pixel.size());
    }
    return 0;
}
