// loader_dump.hpp — TEST INFRASTRUCTURE.  The canonical binary dump ("RT3L") of a loadOBJ result, shared by the
// reference-side dumper (oracle/ref_loader/dump_ref_loader.cpp, linked with the reference's src/mesh.cpp) and the
// dumper of this repo's own loader (tests/tools/dump_own_loader.cpp).  Two loaders agree iff their dumps are
// byte-identical.  Little-endian u32 / i32 / f32:
//   "RT3L" n_meshes n_textures
//   per mesh:    num_keys nv nt; per key: n,verts[3n] n,normals[3n] n,texcoords[2n]; indices[3nt];
//                material floats[10] = Kd Ke roughness anisotropy ior transmittance; ints[4] = diffuse / emissive / roughness / normal texture id
//   per texture: width height rgba8[4wh]  (row 0 = image bottom, src/mesh.cpp:151-159)
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace rt3dump {
struct Writer {
    std::FILE* f;
    explicit Writer(const char* path) : f(std::fopen(path, "wb")) { if (!f) { std::perror(path); std::exit(3); } }
    ~Writer() { std::fclose(f); }
    void count(uint32_t v) { std::fwrite(&v, 4, 1, f); }
    void floats(const float* p, size_t n) { if (n) std::fwrite(p, 4, n, f); }
    void ints(const int* p, size_t n) { if (n) std::fwrite(p, 4, n, f); }
    void bytes(const void* p, size_t n) { if (n) std::fwrite(p, 1, n, f); }
    void header(uint32_t nm, uint32_t nt) { std::fwrite("RT3L", 1, 4, f); count(nm); count(nt); }
    void mesh_begin(uint32_t keys, uint32_t nv, uint32_t nt) { count(keys); count(nv); count(nt); }
};
}  // namespace rt3dump
