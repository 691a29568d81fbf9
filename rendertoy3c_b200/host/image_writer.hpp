// image_writer.hpp — saveImage for the headless host, written from scratch (the reference's
// sutil::saveImage, sutil/sutil.cpp:542-700, picks the format by extension and goes through stb_image_write /
// tinyexr; neither is used here):
//   * .ppm — binary P6 from the 8-bit sRGB frame buffer
//   * .png — 8-bit RGBA from the frame buffer (stored deflate blocks: valid PNG, no compression)
//   * .exr — 32-bit float RGBA scanline file, uncompressed, from the float accumulation buffer
// Buffers arrive with row 0 = image bottom (raygen.cu launch index y, Q19); files are written top row first,
// i.e. flipped, as the reference does (sutil.cpp:552, stbi_flip_vertically_on_write).
// tonemap_aces() is the display curve of the reference's GL viewer (src/gui/display.cpp:119-127, Narkowicz 2015),
// applied there to the frame texture as fetched, i.e. to the sRGB-encoded values; same here.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "rt3_host.hpp"

namespace rt3host {
namespace detail {

inline uint32_t crc32_of(const uint8_t* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    return ~crc;
}
inline void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); }
inline void png_chunk(std::vector<uint8_t>& file, const char tag[4], const std::vector<uint8_t>& data) {
    put_be32(file, (uint32_t)data.size());
    const size_t at = file.size();
    file.insert(file.end(), tag, tag + 4);
    file.insert(file.end(), data.begin(), data.end());
    put_be32(file, crc32_of(&file[at], file.size() - at));
}
template <class T> inline void put_le(std::vector<uint8_t>& v, T x) { uint8_t b[sizeof(T)]; std::memcpy(b, &x, sizeof(T)); v.insert(v.end(), b, b + sizeof(T)); }
inline void put_str(std::vector<uint8_t>& v, const char* s) { v.insert(v.end(), s, s + std::strlen(s) + 1); }
inline bool write_file(const std::string& path, const std::vector<uint8_t>& data) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(data.data(), 1, data.size(), f) == data.size();
    return std::fclose(f) == 0 && ok;
}

}  // namespace detail

// display curve of the reference viewer, per channel on [0,1] values
inline float tonemap_aces(float x) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return (x * (a * x + b)) / (x * (c * x + d) + e);
}
inline void tonemap_frame_aces(std::vector<uint8_t>& rgba8) {
    for (size_t i = 0; i < rgba8.size(); ++i) {
        if ((i & 3) == 3) continue;
        float v = tonemap_aces((float)rgba8[i] / 255.0f);
        v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        rgba8[i] = (uint8_t)std::lround(v * 255.0f);
    }
}

inline void save_ppm(const std::string& path, int w, int h, const uint8_t* frame) {
    std::vector<uint8_t> f;
    char hdr[64];
    const int n = std::snprintf(hdr, sizeof(hdr), "P6\n%d %d\n255\n", w, h);
    f.insert(f.end(), hdr, hdr + n);
    for (int y = h - 1; y >= 0; --y)
        for (int x = 0; x < w; ++x) { const uint8_t* p = &frame[4 * ((size_t)y * w + x)]; f.insert(f.end(), p, p + 3); }
    if (!detail::write_file(path, f)) throw Exception("cannot write " + path);
}

inline void save_png(const std::string& path, int w, int h, const uint8_t* frame) {
    using namespace detail;
    // filter byte 0 + RGBA per scanline, top row first
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)4 * w + 1) * h);
    for (int y = h - 1; y >= 0; --y) {
        raw.push_back(0);
        raw.insert(raw.end(), &frame[(size_t)4 * w * y], &frame[(size_t)4 * w * (y + 1)]);
    }
    // zlib stream of stored blocks (<= 65535 bytes each) + Adler-32
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t s1 = 1, s2 = 0;
    size_t pos = 0;
    do {
        const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        z.push_back(pos + n >= raw.size() ? 1 : 0);  // BFINAL on the last block, BTYPE = 00 (stored)
        z.push_back((uint8_t)(n & 255)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 255)); z.push_back((uint8_t)((~n >> 8) & 255));
        z.insert(z.end(), raw.begin() + (std::ptrdiff_t)pos, raw.begin() + (std::ptrdiff_t)(pos + n));
        for (size_t i = pos; i < pos + n; ++i) { s1 = (s1 + raw[i]) % 65521u; s2 = (s2 + s1) % 65521u; }
        pos += n;
    } while (pos < raw.size());
    put_be32(z, (s2 << 16) | s1);
    std::vector<uint8_t> f = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a}, ihdr;
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(f, "IHDR", ihdr);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
    if (!write_file(path, f)) throw Exception("cannot write " + path);
}

inline void save_exr(const std::string& path, int w, int h, const float* accum) {
    using namespace detail;
    std::vector<uint8_t> f = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    auto attr = [&](const char* name, const char* type, const std::vector<uint8_t>& val) {
        put_str(f, name); put_str(f, type); put_le<int32_t>(f, (int32_t)val.size());
        f.insert(f.end(), val.begin(), val.end());
    };
    std::vector<uint8_t> v;
    for (const char* ch : {"A", "B", "G", "R"}) {  // channel list, alphabetical; FLOAT = 2
        put_str(v, ch); put_le<int32_t>(v, 2); v.push_back(0); v.push_back(0); v.push_back(0); v.push_back(0);
        put_le<int32_t>(v, 1); put_le<int32_t>(v, 1);
    }
    v.push_back(0);
    attr("channels", "chlist", v);
    attr("compression", "compression", {0});
    v.clear(); put_le<int32_t>(v, 0); put_le<int32_t>(v, 0); put_le<int32_t>(v, w - 1); put_le<int32_t>(v, h - 1);
    attr("dataWindow", "box2i", v);
    attr("displayWindow", "box2i", v);
    attr("lineOrder", "lineOrder", {0});
    v.clear(); put_le<float>(v, 1.0f);
    attr("pixelAspectRatio", "float", v);
    v.clear(); put_le<float>(v, 0.0f); put_le<float>(v, 0.0f);
    attr("screenWindowCenter", "v2f", v);
    v.clear(); put_le<float>(v, 1.0f);
    attr("screenWindowWidth", "float", v);
    f.push_back(0);
    const size_t table = f.size(), line_bytes = (size_t)16 * w;
    for (int y = 0; y < h; ++y) put_le<uint64_t>(f, (uint64_t)(table + (size_t)8 * h + (size_t)y * (8 + line_bytes)));
    static const int chan_of[4] = {3, 2, 1, 0};  // A, B, G, R from RGBA
    for (int y = 0; y < h; ++y) {
        put_le<int32_t>(f, y); put_le<int32_t>(f, (int32_t)line_bytes);
        const float* row = accum + (size_t)4 * w * (size_t)(h - 1 - y);  // file row 0 = image top
        for (int c = 0; c < 4; ++c)
            for (int x = 0; x < w; ++x) put_le<float>(f, row[4 * x + chan_of[c]]);
    }
    if (!write_file(path, f)) throw Exception("cannot write " + path);
}

// by extension, like sutil::saveImage; `frame` = uchar4 frame buffer, `accum` = float4 accumulation buffer (needed for .exr)
inline void saveImage(const std::string& path, int w, int h, const uint8_t* frame, const float* accum) {
    if (path.size() < 5) throw Exception("saveImage: failed to determine filename extension");
    std::string ext = path.substr(path.size() - 3);
    for (char& c : ext) c = (char)std::tolower((unsigned char)c);
    if (ext == "ppm") save_ppm(path, w, h, frame);
    else if (ext == "png") save_png(path, w, h, frame);
    else if (ext == "exr") { if (!accum) throw Exception("saveImage: .exr needs the float accumulation buffer"); save_exr(path, w, h, accum); }
    else throw Exception("saveImage: unsupported extension '" + ext + "' (ppm, png, exr)");
}

}  // namespace rt3host
