// obj_loader.hpp — Wavefront .obj/.mtl ingest that returns the SAME Mesh / Texture lists as the reference's loadOBJ
// (src/mesh.cpp:37-210), written from scratch: the reference hands the parsing to its vendored tinyobjloader 2.0
// (support/tinyobj/tiny_obj_loader.h) and the textures to stb_image; neither is used here, their published behaviour
// is restated.  Pinned byte for byte against the reference's own loader compiled where it lies
// (oracle/Makefile target `ref_loader`, tests/test_loader_parity.py, tests/golden/loader/).
//
// What "the same" takes (tiny_obj_loader.h line numbers):
//   * lines end at \n, \r\n or \r (:764-797); numbers go through tinyobj's own decimal parser (:890-1022: digits
//     accumulated in a double, fraction digits weighted by a 10^-k table, exponent applied as ldexp(m * 5^e, e), then
//     narrowed to float) — not strtod, whose last bit can differ; an unparsable number is 0;
//   * face corners `v`, `v/vt`, `v//vn`, `v/vt/vn` with atoi semantics, 1-based or relative (negative) indices
//     (:808-845, :1160-1213); a zero or unresolvable vertex index aborts the load (the reference exit(1)s,
//     src/mesh.cpp:46-51; here an Exception);
//   * faces collect in a group that is triangulated and flushed — with the vertex array as it stands at that point —
//     when `usemtl` CHANGES the material, at `g`, at `o` and at the end of the file (:2836-2862, :2908-2976,
//     :3088-3098); `g` / `o` also close the shape if it holds any triangle;
//   * triangles as written; quads along the shorter diagonal (:1483-1592); larger polygons by tinyobj's built-in ear
//     clipping (:1714-1938): projection axes from the first non-degenerate corner, ears tested in order from a moving
//     guess vertex, convexity by the sign of cross x (signed area term of the first edge), other vertices tested with
//     the even-odd point-in-triangle rule, at most n idle iterations, the last three vertices as the final triangle;
//   * .mtl (:2042-2424): a material is committed at the next `newmtl` (only if named) and at the end of the file
//     (always); the FIRST material of a name wins (map::insert); defaults Kd = Ke = 0, Pr = 0, aniso = 0, Ni = 1,
//     Tf = 0; `map_Kd` without a preceding `Kd` ANYWHERE earlier in the file sets Kd = 0.6 (has_kd is never reset);
//     texture statements skip their -options and take the rest of the line as the file name (:1246-1327);
//   * loadOBJ itself (src/mesh.cpp:62-206): one Mesh per (shape x ascending material id); vertices de-duplicated per
//     mesh on (v, vn, vt), numbered by first use; a corner without vn / vt leaves a gap that the NEXT corner that has
//     one fills (mesh.cpp:97-106), so normals / texcoords may end up shorter than vertices — CUDAMesh refuses such a
//     mesh at upload, the reference reads out of bounds there (Q11); textures de-duplicated PER MESH by the name in
//     the .mtl (a file used by two meshes is loaded twice, mesh.cpp:72,119-122); loaded in the order diffuse,
//     emissive, roughness, normal; the directory prefix is everything up to the last '/' of the .obj path plus "/"
//     (so an .obj named without any '/' looks for "/<texture>", mesh.cpp:124-135,170); failures give id -1;
//   * key-frame files paths[1..]: only their v / vn / vt arrays are used, indexed with the first file's faces
//     (mesh.cpp:39-58,88-107) -> Mesh::num_keys = paths.size().
// Where the reference has undefined behaviour (faces without a material -> materials[-1]; indices beyond the arrays)
// this loader throws.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <tuple>

#include "image_loader.hpp"
#include "rt3_host.hpp"

namespace rt3host {

namespace detail {
struct ObjIndex { int v = -1, vt = -1, vn = -1; bool operator<(const ObjIndex& o) const { return std::tie(v, vn, vt) < std::tie(o.v, o.vn, o.vt); } };
struct ObjShape { std::vector<ObjIndex> idx; std::vector<int> mat; };  // 3 corners and one material id per triangle
struct ObjMaterial {
    std::string name;
    float Kd[3] = {0, 0, 0}, Ke[3] = {0, 0, 0}, Tf[3] = {0, 0, 0};
    float Pr = 0.0f, aniso = 0.0f, Ni = 1.0f;
    std::string map_Kd, map_Ke, map_Pr, norm;
};
struct ObjArrays { std::vector<float> v, vn, vt; };

inline bool obj_space(char c) { return c == ' ' || c == '\t'; }
inline bool obj_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }
inline bool obj_digit(char c) { return (unsigned)(c - '0') < 10u; }

// the whole file as lines; \n, \r\n and a lone \r all end a line
inline bool obj_read_lines(const std::string& path, std::vector<std::string>& lines) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string s = ss.str();
    std::string cur;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (c == '\n') { lines.push_back(cur); cur.clear(); }
        else if (c == '\r') { if (i + 1 < s.size() && s[i + 1] == '\n') ++i; lines.push_back(cur); cur.clear(); }
        else cur += c;
    }
    if (!cur.empty()) lines.push_back(cur);
    return true;
}

// tinyobj's decimal parser (tiny_obj_loader.h:890-1022), greedy over [s, e); false = no number (caller keeps its default)
inline bool obj_parse_double(const char* s, const char* e, double& out) {
    if (s >= e) return false;
    static const double frac_weight[8] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
    double mant = 0.0;
    int expo = 0;
    bool neg = false, exp_neg = false, lead_dot = false;
    const char* c = s;
    if (*c == '+' || *c == '-') {
        neg = *c == '-';
        ++c;
        if (c != e && *c == '.') lead_dot = true;
    } else if (*c == '.') lead_dot = true;
    else if (!obj_digit(*c)) return false;
    if (!lead_dot) {
        int nd = 0;
        for (; c != e && obj_digit(*c); ++c, ++nd) mant = mant * 10 + (double)(*c - '0');
        if (nd == 0) return false;
    }
    bool more = c != e;
    if (more && *c == '.') {
        ++c;
        for (int k = 1; c != e && obj_digit(*c); ++c, ++k) mant += (double)(*c - '0') * (k < 8 ? frac_weight[k] : std::pow(10.0, -k));
        more = c != e;
    } else if (more && !(*c == 'e' || *c == 'E')) more = false;
    if (more && (*c == 'e' || *c == 'E')) {
        ++c;
        if (c != e && (*c == '+' || *c == '-')) { exp_neg = *c == '-'; ++c; }
        else if (c == e || !obj_digit(*c)) return false;   // a bare 'e'
        int nd = 0;
        for (; c != e && obj_digit(*c); ++c, ++nd) {
            if (expo > 2147483647 / 10) return false;
            expo = expo * 10 + (*c - '0');
        }
        if (nd == 0) return false;
        if (exp_neg) expo = -expo;
    }
    out = (neg ? -1 : 1) * (expo ? std::ldexp(mant * std::pow(5.0, expo), expo) : mant);
    return true;
}
inline float obj_real(const char*& tok, double dflt = 0.0) {
    tok += std::strspn(tok, " \t");
    const char* end = tok + std::strcspn(tok, " \t\r");
    double v = dflt;
    obj_parse_double(tok, end, v);
    tok = end;
    return (float)v;
}
inline std::string obj_word(const char*& tok) {
    tok += std::strspn(tok, " \t");
    const size_t n = std::strcspn(tok, " \t\r");
    std::string s(tok, tok + n);
    tok += n;
    return s;
}
// 1-based / relative index -> 0-based; a literal 0 gives -1 and is an error only where `zero_ok` is false
inline bool obj_fix_index(int idx, int n, bool zero_ok, int& out) {
    if (idx > 0) { out = idx - 1; return true; }
    if (idx == 0) { out = -1; return zero_ok; }
    out = n + idx;
    return out >= 0;
}
inline bool obj_corner(const char*& tok, int nv, int nvn, int nvt, ObjIndex& out) {
    ObjIndex ix;
    if (!obj_fix_index(std::atoi(tok), nv, false, ix.v)) return false;
    tok += std::strcspn(tok, "/ \t\r");
    if (*tok != '/') { out = ix; return true; }
    ++tok;
    if (*tok == '/') {  // v//vn
        ++tok;
        if (!obj_fix_index(std::atoi(tok), nvn, true, ix.vn)) return false;
        tok += std::strcspn(tok, "/ \t\r");
        out = ix;
        return true;
    }
    if (!obj_fix_index(std::atoi(tok), nvt, true, ix.vt)) return false;
    tok += std::strcspn(tok, "/ \t\r");
    if (*tok != '/') { out = ix; return true; }
    ++tok;
    if (!obj_fix_index(std::atoi(tok), nvn, true, ix.vn)) return false;
    tok += std::strcspn(tok, "/ \t\r");
    out = ix;
    return true;
}

// even-odd rule for a triangle (W. R. Franklin's pnpoly, as tinyobj uses it)
inline bool obj_in_triangle(const float x[3], const float y[3], float tx, float ty) {
    bool in = false;
    for (int i = 0, j = 2; i < 3; j = i++)
        if (((y[i] > ty) != (y[j] > ty)) && (tx < (x[j] - x[i]) * (ty - y[i]) / (y[j] - y[i]) + x[i])) in = !in;
    return in;
}

// triangulate one face into `sh` (tiny_obj_loader.h:1476-1965); V = the vertex array read so far
inline void obj_emit_face(const std::vector<ObjIndex>& face, const std::vector<float>& V, int mat, ObjShape& sh) {
    const size_t n = face.size();
    if (n < 3) return;
    auto emit = [&](const ObjIndex& a, const ObjIndex& b, const ObjIndex& c) { sh.idx.insert(sh.idx.end(), {a, b, c}); sh.mat.push_back(mat); };
    auto inside = [&](const ObjIndex& i) { return 3 * (size_t)i.v + 2 < V.size(); };
    if (n == 3) { emit(face[0], face[1], face[2]); return; }
    if (n == 4) {
        if (!inside(face[0]) || !inside(face[1]) || !inside(face[2]) || !inside(face[3])) return;
        const float* p0 = &V[3 * (size_t)face[0].v], *p1 = &V[3 * (size_t)face[1].v], *p2 = &V[3 * (size_t)face[2].v], *p3 = &V[3 * (size_t)face[3].v];
        const float ax = p2[0] - p0[0], ay = p2[1] - p0[1], az = p2[2] - p0[2];
        const float bx = p3[0] - p1[0], by = p3[1] - p1[1], bz = p3[2] - p1[2];
        const float d02 = ax * ax + ay * ay + az * az, d13 = bx * bx + by * by + bz * bz;
        if (d02 < d13) { emit(face[0], face[1], face[2]); emit(face[0], face[2], face[3]); }
        else { emit(face[0], face[1], face[3]); emit(face[1], face[2], face[3]); }
        return;
    }
    // the two projection axes: drop the dominant axis of the first corner whose edges are not parallel
    size_t ax0 = 1, ax1 = 2;
    for (size_t k = 0; k < n; ++k) {
        const ObjIndex &i0 = face[k % n], &i1 = face[(k + 1) % n], &i2 = face[(k + 2) % n];
        if (!inside(i0) || !inside(i1) || !inside(i2)) continue;
        const float* a = &V[3 * (size_t)i0.v], *b = &V[3 * (size_t)i1.v], *c = &V[3 * (size_t)i2.v];
        const float e0x = b[0] - a[0], e0y = b[1] - a[1], e0z = b[2] - a[2];
        const float e1x = c[0] - b[0], e1y = c[1] - b[1], e1z = c[2] - b[2];
        const float cx = std::fabs(e0y * e1z - e0z * e1y), cy = std::fabs(e0z * e1x - e0x * e1z), cz = std::fabs(e0x * e1y - e0y * e1x);
        const float eps = 1.1920929e-07f;  // FLT_EPSILON
        if (cx > eps || cy > eps || cz > eps) {
            if (!(cx > cy && cx > cz)) {
                ax0 = 0;
                if (cz > cx && cz > cy) ax1 = 1;
            }
            break;
        }
    }
    std::vector<ObjIndex> rest = face;
    size_t guess = 0, idle_left = n, prev_size = n;
    while (rest.size() > 3 && idle_left > 0) {
        const size_t m = rest.size();
        if (guess >= m) guess -= m;
        if (prev_size != m) { prev_size = m; idle_left = m; } else --idle_left;
        ObjIndex tri[3];
        float px[3], py[3];
        for (size_t k = 0; k < 3; ++k) {
            tri[k] = rest[(guess + k) % m];
            const size_t vi = (size_t)tri[k].v;
            const bool ok = vi * 3 + ax0 < V.size() && vi * 3 + ax1 < V.size();
            px[k] = ok ? V[vi * 3 + ax0] : 0.0f;
            py[k] = ok ? V[vi * 3 + ax1] : 0.0f;
        }
        const float e0x = px[1] - px[0], e0y = py[1] - py[0], e1x = px[2] - px[1], e1y = py[2] - py[1];
        const float turn = e0x * e1y - e0y * e1x;
        const float area = (px[0] * py[1] - py[0] * px[1]) * 0.5f;
        if (turn * area < 0.0f) { ++guess; continue; }   // reflex corner
        bool covered = false;
        for (size_t o = 3; o < m && !covered; ++o) {
            const size_t vi = (size_t)rest[(guess + o) % m].v;
            if (vi * 3 + ax0 >= V.size() || vi * 3 + ax1 >= V.size()) continue;
            covered = obj_in_triangle(px, py, V[vi * 3 + ax0], V[vi * 3 + ax1]);
        }
        if (covered) { ++guess; continue; }
        emit(tri[0], tri[1], tri[2]);
        rest.erase(rest.begin() + (std::ptrdiff_t)((guess + 1) % m));   // the ear's tip leaves the polygon
    }
    if (rest.size() == 3) emit(rest[0], rest[1], rest[2]);
}

// a texture statement: options first (each with its fixed number of arguments), then the file name = rest of the line
inline std::string obj_texture_name(const char* tok) {
    std::string name;
    auto opt = [&](const char* key, size_t len) { return std::strncmp(tok, key, len) == 0 && obj_space(tok[len]); };
    while (!obj_eol(*tok)) {
        tok += std::strspn(tok, " \t");
        if (opt("-blendu", 7) || opt("-blendv", 7)) { tok += 8; obj_word(tok); }
        else if (opt("-clamp", 6)) { tok += 7; obj_word(tok); }
        else if (opt("-boost", 6)) { tok += 7; obj_real(tok); }
        else if (opt("-bm", 3)) { tok += 4; obj_real(tok); }
        else if (opt("-o", 2) || opt("-s", 2) || opt("-t", 2)) { tok += 3; obj_real(tok); obj_real(tok); obj_real(tok); }
        else if (opt("-type", 5)) { tok += 5; obj_word(tok); }
        else if (opt("-texres", 7)) { tok += 7; obj_word(tok); }
        else if (opt("-imfchan", 8)) { tok += 9; obj_word(tok); }
        else if (opt("-mm", 3)) { tok += 4; obj_real(tok); obj_real(tok); }
        else if (opt("-colorspace", 11)) { tok += 12; obj_word(tok); }
        else { name = tok; tok += name.size(); }
    }
    return name;
}

inline void obj_load_mtl(const std::string& file, std::vector<ObjMaterial>& mats, std::map<std::string, int>& ids) {
    std::vector<std::string> lines;
    obj_read_lines(file, lines);
    ObjMaterial cur;
    bool has_kd = false;   // never reset at `newmtl`, like the reference's local (tiny_obj_loader.h:2057,2108-2109)
    auto commit = [&]() { ids.insert({cur.name, (int)mats.size()}); mats.push_back(cur); };
    auto key = [](const char* t, const char* k) { const size_t n = std::strlen(k); return std::strncmp(t, k, n) == 0 && obj_space(t[n]); };
    for (std::string l : lines) {
        const size_t last = l.find_last_not_of(" \t");
        l = last == std::string::npos ? std::string() : l.substr(0, last + 1);
        const char* t = l.c_str();
        t += std::strspn(t, " \t");
        if (*t == '\0' || *t == '#') continue;
        if (key(t, "newmtl")) {
            if (!cur.name.empty()) commit();
            cur = ObjMaterial();
            t += 7;
            cur.name = obj_word(t);
        }
        else if (key(t, "Kd")) { t += 2; for (float& c : cur.Kd) c = obj_real(t); has_kd = true; }
        else if (key(t, "Ke")) { t += 2; for (float& c : cur.Ke) c = obj_real(t); }
        else if (key(t, "Kt") || key(t, "Tf")) { t += 2; for (float& c : cur.Tf) c = obj_real(t); }
        else if (key(t, "Ni")) { t += 2; cur.Ni = obj_real(t); }
        else if (key(t, "Pr")) { t += 2; cur.Pr = obj_real(t); }
        else if (key(t, "aniso")) { t += 6; cur.aniso = obj_real(t); }
        else if (key(t, "map_Kd")) { cur.map_Kd = obj_texture_name(t + 7); if (!has_kd) cur.Kd[0] = cur.Kd[1] = cur.Kd[2] = 0.6f; }
        else if (key(t, "map_Ke")) cur.map_Ke = obj_texture_name(t + 7);
        else if (key(t, "map_Pr")) cur.map_Pr = obj_texture_name(t + 7);
        else if (key(t, "norm")) cur.norm = obj_texture_name(t + 5);
    }
    commit();
}

// v / vn / vt arrays, shapes and materials of one file (shapes / materials only wanted for the first key)
inline void obj_parse_file(const std::string& path, ObjArrays& A, std::vector<ObjShape>* shapes, std::vector<ObjMaterial>* mats) {
    std::vector<std::string> lines;
    if (!obj_read_lines(path, lines)) throw Exception("loadOBJ: cannot open " + path);
    std::string mtl_dir;
    const size_t sep = path.find_last_of("/\\");
    if (sep != std::string::npos && sep > 0) mtl_dir = path.substr(0, sep) + "/";
    std::map<std::string, int> mat_id;
    std::set<std::string> mtl_loaded;
    std::vector<ObjMaterial> local_mats;
    std::vector<ObjMaterial>& M = mats ? *mats : local_mats;
    std::vector<std::vector<ObjIndex>> group;   // faces since the last flush
    ObjShape shape;
    int material = -1;
    auto flush = [&]() { for (const auto& f : group) obj_emit_face(f, A.v, material, shape); };
    auto close_shape = [&]() { if (!shape.idx.empty() && shapes) shapes->push_back(shape); shape = ObjShape(); };
    size_t line_no = 0;
    for (const std::string& l : lines) {
        ++line_no;
        const char* t = l.c_str();
        t += std::strspn(t, " \t");
        if (*t == '\0' || *t == '#') continue;
        if (t[0] == 'v' && obj_space(t[1])) { t += 2; for (int q = 0; q < 3; ++q) A.v.push_back(obj_real(t)); }
        else if (t[0] == 'v' && t[1] == 'n' && obj_space(t[2])) { t += 3; for (int q = 0; q < 3; ++q) A.vn.push_back(obj_real(t)); }
        else if (t[0] == 'v' && t[1] == 't' && obj_space(t[2])) { t += 3; for (int q = 0; q < 2; ++q) A.vt.push_back(obj_real(t)); }
        else if (t[0] == 'f' && obj_space(t[1])) {
            t += 2;
            t += std::strspn(t, " \t");
            std::vector<ObjIndex> face;
            while (!obj_eol(*t)) {
                ObjIndex ix;
                if (!obj_corner(t, (int)A.v.size() / 3, (int)A.vn.size() / 3, (int)A.vt.size() / 2, ix))
                    throw Exception("loadOBJ: " + path + " line " + std::to_string(line_no) + ": bad face index (zero, or a relative index before the start)");
                face.push_back(ix);
                t += std::strspn(t, " \t\r");
            }
            group.push_back(std::move(face));
        }
        else if (std::strncmp(t, "usemtl", 6) == 0) {
            t += 6;
            const std::string name = obj_word(t);
            const auto it = mat_id.find(name);
            const int id = it == mat_id.end() ? -1 : it->second;
            if (id != material) { flush(); group.clear(); material = id; }
        }
        else if (std::strncmp(t, "mtllib", 6) == 0 && obj_space(t[6])) {
            // file names separated by blanks, '\' escapes; the first one that opens is read
            std::vector<std::string> names;
            std::string cur;
            bool esc = false;
            for (const char* c = t + 7; *c; ++c) {
                if (esc) { esc = false; cur += *c; }
                else if (*c == '\\') esc = true;
                else if (*c == ' ') { if (!cur.empty()) names.push_back(cur); cur.clear(); }
                else cur += *c;
            }
            names.push_back(cur);
            for (const std::string& nm : names) {
                if (mtl_loaded.count(nm)) continue;
                const std::string file = mtl_dir + nm;
                if (!std::ifstream(file)) continue;
                obj_load_mtl(file, M, mat_id);
                mtl_loaded.insert(nm);
                break;
            }
        }
        else if ((t[0] == 'g' || t[0] == 'o') && obj_space(t[1])) { flush(); group.clear(); close_shape(); }
    }
    flush();
    close_shape();
}

}  // namespace detail

// paths[0] is the scene; paths[1..] are key-frames of the SAME topology: only their v / vn / vt arrays are read and
// indexed with the first file's faces (src/mesh.cpp:39-110) -> Mesh::num_keys = paths.size(), vertex-key motion blur
inline void loadOBJ(const std::vector<std::string>& paths, std::vector<Mesh>& meshes, std::vector<Texture>& textures) {
    using namespace detail;
    if (paths.empty()) throw Exception("loadOBJ: no file given");
    const std::string& path = paths[0];
    const size_t nkeys = paths.size();
    std::vector<ObjArrays> K(nkeys);
    std::vector<ObjShape> shapes;
    std::vector<ObjMaterial> mats;
    obj_parse_file(path, K[0], &shapes, &mats);
    for (size_t k = 1; k < nkeys; ++k) obj_parse_file(paths[k], K[k], nullptr, nullptr);
    const std::string tex_dir = path.substr(0, path.rfind('/') + 1) + "/";   // src/mesh.cpp:170,135
    for (const ObjShape& sh : shapes) {
        std::set<int> ids(sh.mat.begin(), sh.mat.end());
        for (int mid : ids) {
            if (mid < 0 || (size_t)mid >= mats.size()) throw Exception("loadOBJ: face without material (the reference dereferences materials[-1], Q11)");
            Mesh mesh;
            mesh.num_keys = (unsigned)nkeys;
            mesh.vertices.resize(nkeys); mesh.normals.resize(nkeys); mesh.texcoords.resize(nkeys);
            std::map<ObjIndex, int> known;
            std::map<std::string, int> known_tex;   // per mesh, like the reference
            auto vertex_id = [&](const ObjIndex& ix) -> int {   // addVertexAndGetIndexInMesh, mesh.cpp:78-110
                const auto it = known.find(ix);
                if (it != known.end()) return it->second;
                const int id = (int)mesh.vertices[0].size() / 3;
                known[ix] = id;
                for (size_t k = 0; k < nkeys; ++k) {
                    if (3 * (size_t)ix.v + 2 >= K[k].v.size() || (ix.vn >= 0 && 3 * (size_t)ix.vn + 2 >= K[k].vn.size()) || (ix.vt >= 0 && 2 * (size_t)ix.vt + 1 >= K[k].vt.size()))
                        throw Exception("loadOBJ: " + paths[k] + " has fewer v / vn / vt entries than the faces of " + path + " use");
                    for (int q = 0; q < 3; ++q) mesh.vertices[k].push_back(K[k].v[3 * (size_t)ix.v + q]);
                    if (ix.vn >= 0) while (mesh.normals[k].size() < mesh.vertices[k].size()) for (int q = 0; q < 3; ++q) mesh.normals[k].push_back(K[k].vn[3 * (size_t)ix.vn + q]);
                    if (ix.vt >= 0) while (mesh.texcoords[k].size() / 2 < mesh.vertices[k].size() / 3) for (int q = 0; q < 2; ++q) mesh.texcoords[k].push_back(K[k].vt[2 * (size_t)ix.vt + q]);
                }
                return id;
            };
            auto texture_id = [&](const std::string& name) -> int {  // addTextureAndGetTextureId, mesh.cpp:112-168
                if (name.empty()) return -1;
                const auto it = known_tex.find(name);
                if (it != known_tex.end()) return it->second;
                int id = -1;
                Texture t;
                std::string fn = name;
                for (char& ch : fn) if (ch == '\\') ch = '/';
                std::string why;
                if (load_image(tex_dir + fn, t, &why)) { id = (int)textures.size(); textures.push_back(std::move(t)); }
                else std::fprintf(stderr, "Error loading texture %s (%s).\n", fn.c_str(), why.c_str());
                known_tex[name] = id;
                return id;
            };
            const ObjMaterial& m = mats[(size_t)mid];
            bool first = true;
            for (size_t f = 0; f < sh.mat.size(); ++f) {
                if (sh.mat[f] != mid) continue;
                for (int c = 0; c < 3; ++c) mesh.indices.push_back(vertex_id(sh.idx[3 * f + (size_t)c]));
                if (!first) continue;   // the reference re-assigns the same material per face; once is the same
                first = false;
                mesh.material.m_diffuse = {m.Kd[0], m.Kd[1], m.Kd[2]};
                mesh.material.m_diffuseTextureID = texture_id(m.map_Kd);
                mesh.material.m_emissive = {m.Ke[0], m.Ke[1], m.Ke[2]};
                mesh.material.m_emissiveTextureID = texture_id(m.map_Ke);
                mesh.material.m_roughness = m.Pr;
                mesh.material.m_roughnessTextureID = texture_id(m.map_Pr);
                mesh.material.m_anisotropy = m.aniso;
                mesh.material.m_ior = m.Ni;
                mesh.material.m_transmittance = m.Tf[0];
                mesh.material.m_normalTextureID = texture_id(m.norm);
            }
            if (!mesh.vertices[0].empty()) meshes.push_back(std::move(mesh));
        }
    }
}

inline void loadOBJ(const std::string& path, std::vector<Mesh>& meshes, std::vector<Texture>& textures) {
    loadOBJ(std::vector<std::string>{path}, meshes, textures);
}

}  // namespace rt3host
