// obj_loader.hpp — minimal Wavefront .obj/.mtl ingest producing the same Mesh list, primitive order
// and vertex order as the reference's loadOBJ (src/mesh.cpp:37-210), written from scratch (the
// reference vendors tinyobj + stb_image; neither is used here).
//   * one Mesh per (shape x ascending material id), shapes split at every `o` / `g` statement;
//   * faces in file order; vertices de-duplicated per mesh on the (v, vn, vt) triple, numbered by
//     first use (mesh.cpp:78-110);
//   * quads split along the shorter diagonal (tinyobj rule), larger polygons as a fan;
//   * material fields Kd, Ke, map_Kd (+ map_Ke, Pr, map_Pr, aniso, Ni, Tf, norm: carried, unused by shading, as in the
//     reference); textures through image_loader.hpp (PNG, BMP, TGA, PPM/PGM -> RGBA8,
//     rows flipped so that v = 0 is the image bottom, mesh.cpp:151-159), de-duplicated by file name;
//   * normals and texcoords are required (the reference reads them unconditionally, Q11).
#pragma once
#include <cstdio>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <tuple>

#include "image_loader.hpp"
#include "rt3_host.hpp"

namespace rt3host {

namespace detail {
struct ObjIndex { int v, vt, vn; bool operator<(const ObjIndex& o) const { return std::tie(v, vn, vt) < std::tie(o.v, o.vn, o.vt); } };
struct ObjShape { std::vector<ObjIndex> idx; std::vector<int> mat; };  // 3 indices per triangle
struct ObjMaterial {
    std::string name;
    float3_ Kd{0.8f, 0.8f, 0.8f}, Ke{0, 0, 0};
    float Pr = 0.0f, aniso = 0.0f, Ni = 1.0f, Tf = 0.0f;  // PBR extension: roughness, anisotropy; ior; transmittance (first component)
    std::string map_Kd, map_Ke, map_Pr, norm;
};

inline int fix_index(int i, int n) { return i > 0 ? i - 1 : n + i; }

}  // namespace detail

// paths[0] is the scene; paths[1..] are key-frames of the SAME topology: only their v / vn / vt arrays are read and
// indexed with the first file's faces (src/mesh.cpp:39-110) -> Mesh::num_keys = paths.size(), vertex-key motion blur
inline void loadOBJ(const std::vector<std::string>& paths, std::vector<Mesh>& meshes, std::vector<Texture>& textures) {
    using namespace detail;
    if (paths.empty()) throw Exception("loadOBJ: no file given");
    const std::string& path = paths[0];
    std::ifstream in(path);
    if (!in) throw Exception("loadOBJ: cannot open " + path);
    const std::string dir = path.substr(0, path.rfind('/') + 1);
    const size_t nkeys = paths.size();
    std::vector<std::vector<float>> KV(nkeys), KVN(nkeys), KVT(nkeys);
    for (size_t k = 1; k < nkeys; ++k) {
        std::ifstream kf(paths[k]);
        if (!kf) throw Exception("loadOBJ: cannot open " + paths[k]);
        std::string l;
        while (std::getline(kf, l)) {
            std::istringstream ss(l);
            std::string t;
            ss >> t;
            if (t == "v") { float x, y, z; ss >> x >> y >> z; KV[k].insert(KV[k].end(), {x, y, z}); }
            else if (t == "vn") { float x, y, z; ss >> x >> y >> z; KVN[k].insert(KVN[k].end(), {x, y, z}); }
            else if (t == "vt") { float x, y = 0; ss >> x >> y; KVT[k].insert(KVT[k].end(), {x, y}); }
        }
    }
    std::vector<float>&V = KV[0], &VN = KVN[0], &VT = KVT[0];
    std::vector<ObjShape> shapes(1);
    std::vector<ObjMaterial> mats;
    std::map<std::string, int> mat_id;
    int cur_mat = -1;
    std::string line;
    auto load_mtl = [&](const std::string& file) {
        std::ifstream m(dir + file);
        std::string l;
        while (std::getline(m, l)) {
            std::istringstream ss(l);
            std::string k;
            ss >> k;
            if (k == "newmtl") { ObjMaterial mm; ss >> mm.name; mat_id[mm.name] = (int)mats.size(); mats.push_back(mm); }
            else if (mats.empty()) continue;
            else if (k == "Kd") ss >> mats.back().Kd.x >> mats.back().Kd.y >> mats.back().Kd.z;
            else if (k == "Ke") ss >> mats.back().Ke.x >> mats.back().Ke.y >> mats.back().Ke.z;
            else if (k == "map_Kd") ss >> mats.back().map_Kd;
            else if (k == "map_Ke") ss >> mats.back().map_Ke;
            else if (k == "Pr") ss >> mats.back().Pr;
            else if (k == "map_Pr") ss >> mats.back().map_Pr;
            else if (k == "aniso") ss >> mats.back().aniso;
            else if (k == "Ni") ss >> mats.back().Ni;
            else if (k == "Tf") ss >> mats.back().Tf;
            else if (k == "norm") ss >> mats.back().norm;
        }
    };
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string k;
        ss >> k;
        if (k == "v") { float x, y, z; ss >> x >> y >> z; V.insert(V.end(), {x, y, z}); }
        else if (k == "vn") { float x, y, z; ss >> x >> y >> z; VN.insert(VN.end(), {x, y, z}); }
        else if (k == "vt") { float x, y = 0; ss >> x >> y; VT.insert(VT.end(), {x, y}); }
        else if (k == "o" || k == "g") { if (!shapes.back().idx.empty()) shapes.emplace_back(); }
        else if (k == "mtllib") { std::string f; ss >> f; load_mtl(f); }
        else if (k == "usemtl") { std::string n; ss >> n; auto it = mat_id.find(n); cur_mat = it == mat_id.end() ? -1 : it->second; }
        else if (k == "f") {
            std::vector<ObjIndex> poly;
            std::string tok;
            while (ss >> tok) {
                ObjIndex ix{0, -1, -1};
                int a = 0, b = 0, c = 0;
                if (std::sscanf(tok.c_str(), "%d/%d/%d", &a, &b, &c) == 3) { ix.v = fix_index(a, (int)V.size() / 3); ix.vt = fix_index(b, (int)VT.size() / 2); ix.vn = fix_index(c, (int)VN.size() / 3); }
                else if (std::sscanf(tok.c_str(), "%d//%d", &a, &c) == 2) { ix.v = fix_index(a, (int)V.size() / 3); ix.vn = fix_index(c, (int)VN.size() / 3); }
                else if (std::sscanf(tok.c_str(), "%d/%d", &a, &b) == 2) { ix.v = fix_index(a, (int)V.size() / 3); ix.vt = fix_index(b, (int)VT.size() / 2); }
                else if (std::sscanf(tok.c_str(), "%d", &a) == 1) { ix.v = fix_index(a, (int)V.size() / 3); }
                poly.push_back(ix);
            }
            auto emit = [&](int a, int b, int c) { shapes.back().idx.insert(shapes.back().idx.end(), {poly[(size_t)a], poly[(size_t)b], poly[(size_t)c]}); shapes.back().mat.push_back(cur_mat); };
            if (poly.size() == 3) emit(0, 1, 2);
            else if (poly.size() == 4) {
                auto d2 = [&](int a, int b) { float s = 0; for (int q = 0; q < 3; ++q) { const float e = V[3 * (size_t)poly[(size_t)a].v + q] - V[3 * (size_t)poly[(size_t)b].v + q]; s += e * e; } return s; };
                if (d2(0, 2) < d2(1, 3)) { emit(0, 1, 2); emit(0, 2, 3); } else { emit(0, 1, 3); emit(1, 2, 3); }
            } else for (size_t q = 1; q + 1 < poly.size(); ++q) emit(0, (int)q, (int)q + 1);
        }
    }
    std::map<std::string, int> known_tex;
    for (const ObjShape& sh : shapes) {
        std::set<int> ids(sh.mat.begin(), sh.mat.end());
        for (int mid : ids) {
            if (mid < 0) throw Exception("loadOBJ: face without material (the reference dereferences materials[-1], Q11)");
            Mesh mesh;
            mesh.num_keys = (unsigned)nkeys;
            mesh.vertices.resize(nkeys); mesh.normals.resize(nkeys); mesh.texcoords.resize(nkeys);
            std::map<ObjIndex, int> known;
            for (size_t f = 0; f < sh.mat.size(); ++f) {
                if (sh.mat[f] != mid) continue;
                for (int c = 0; c < 3; ++c) {
                    const ObjIndex ix = sh.idx[3 * f + (size_t)c];
                    if (ix.vn < 0 || ix.vt < 0) throw Exception("loadOBJ: vertex without normal or texcoord (required, Q11)");
                    auto it = known.find(ix);
                    int id;
                    if (it != known.end()) id = it->second;
                    else {
                        id = (int)mesh.vertices[0].size() / 3;
                        known[ix] = id;
                        for (size_t k = 0; k < nkeys; ++k) {
                            if (3 * (size_t)ix.v + 2 >= KV[k].size() || 3 * (size_t)ix.vn + 2 >= KVN[k].size() || 2 * (size_t)ix.vt + 1 >= KVT[k].size())
                                throw Exception("loadOBJ: key-frame " + paths[k] + " has fewer v / vn / vt entries than the faces of " + path + " use");
                            for (int q = 0; q < 3; ++q) mesh.vertices[k].push_back(KV[k][3 * (size_t)ix.v + q]);
                            for (int q = 0; q < 3; ++q) mesh.normals[k].push_back(KVN[k][3 * (size_t)ix.vn + q]);
                            for (int q = 0; q < 2; ++q) mesh.texcoords[k].push_back(KVT[k][2 * (size_t)ix.vt + q]);
                        }
                    }
                    mesh.indices.push_back(id);
                }
            }
            const ObjMaterial& m = mats[(size_t)mid];
            mesh.material.m_diffuse = m.Kd;
            mesh.material.m_emissive = m.Ke;
            auto texture_id = [&](const std::string& name) -> int {  // addTextureAndGetTextureId, mesh.cpp:112-168
                if (name.empty()) return -1;
                auto it = known_tex.find(name);
                if (it != known_tex.end()) return it->second;
                int id = -1;
                Texture t;
                std::string fn = name;
                for (char& ch : fn) if (ch == '\\') ch = '/';
                std::string why;
                if (load_image(dir + fn, t, &why)) { id = (int)textures.size(); textures.push_back(std::move(t)); }
                else std::fprintf(stderr, "Error loading texture %s (%s).\n", fn.c_str(), why.c_str());
                known_tex[name] = id;
                return id;
            };
            mesh.material.m_diffuseTextureID = texture_id(m.map_Kd);    // same order as the reference: diffuse, emissive, roughness, normal
            mesh.material.m_emissiveTextureID = texture_id(m.map_Ke);
            mesh.material.m_roughness = m.Pr;
            mesh.material.m_roughnessTextureID = texture_id(m.map_Pr);
            mesh.material.m_anisotropy = m.aniso;
            mesh.material.m_ior = m.Ni;
            mesh.material.m_transmittance = m.Tf;
            mesh.material.m_normalTextureID = texture_id(m.norm);
            if (!mesh.vertices[0].empty()) meshes.push_back(std::move(mesh));
        }
    }
}

inline void loadOBJ(const std::string& path, std::vector<Mesh>& meshes, std::vector<Texture>& textures) {
    loadOBJ(std::vector<std::string>{path}, meshes, textures);
}

}  // namespace rt3host
