// image_loader.hpp — texture file ingest for the .obj/.mtl loader, written from scratch (the reference
// calls stbi_load(..., STBI_rgb_alpha), src/mesh.cpp:137; stb_image is not used here).
// Decodes to RGBA8, rows flipped so that v = 0 is the image bottom (mesh.cpp:151-159):
//   * PNG  — 1-16 bit, grey / grey+alpha / RGB / RGBA / palette (+ tRNS), plain or Adam7-interlaced; own inflate
//   * BMP  — uncompressed: 1 / 4 / 8-bit palettised, 16 / 32-bit with masks, 24-bit; core, info and V4 / V5 headers
//   * TGA  — true-colour 15 / 16 / 24 / 32 bit, grey, grey + alpha, colour-mapped; raw or RLE; either vertical origin
//   * PPM / PGM — binary P6 / P5, 8 or 16 bits per sample
//   * JPEG — baseline, extended-sequential and progressive (Huffman, 8 bit); grey, YCbCr, RGB, CMYK / YCCK; integer
//            sampling ratios; restart intervals; decoded with stb_image's arithmetic so that the bytes are the reference's
//   * GIF (first image), Photoshop PSD (flattened RGB composite), Softimage PIC, Radiance HDR (tone-mapped to 8 bits as stbi_load does)
// Arithmetic-coded / lossless JPEG are not supported (load_image returns false and says why), as in stb_image.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "rt3_host.hpp"

namespace rt3host {
namespace detail {

inline bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    if (n < 0) return false;
    f.seekg(0);
    out.resize((size_t)n);
    f.read(reinterpret_cast<char*>(out.data()), n);
    return (bool)f;
}

// rows top-down RGBA in `top`, stored bottom row first in the texture
inline void store_flipped(const std::vector<uint8_t>& top, int w, int h, Texture& t) {
    t.width = w; t.height = h;
    t.pixel.resize((size_t)4 * w * h);
    if (t.pixel.empty()) return;
    for (int y = 0; y < h; ++y) std::memcpy(&t.pixel[(size_t)4 * w * y], &top[(size_t)4 * w * (h - 1 - y)], (size_t)4 * w);
}

// ------------------------------------------------------------------------------------ inflate (RFC 1951)
struct BitSrc {
    const uint8_t* p; size_t n, pos = 0; uint32_t acc = 0; int cnt = 0; bool bad = false;
    BitSrc(const uint8_t* d, size_t len) : p(d), n(len) {}
    uint32_t bits(int k) {  // k <= 16, LSB first
        while (cnt < k) {
            if (pos >= n) { bad = true; return 0; }
            acc |= (uint32_t)p[pos++] << cnt;
            cnt += 8;
        }
        const uint32_t v = acc & ((1u << k) - 1u);
        acc >>= k; cnt -= k;
        return v;
    }
    void align() { acc = 0; cnt = 0; }
};
struct HuffTable {  // canonical code: symbols sorted by (length, value)
    uint16_t count[16] = {0};
    std::vector<uint16_t> sym;
    void build(const uint8_t* len, int n) {
        for (int i = 0; i < 16; ++i) count[i] = 0;
        for (int i = 0; i < n; ++i) count[len[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int i = 1; i < 15; ++i) offs[i + 1] = (uint16_t)(offs[i] + count[i]);
        sym.assign((size_t)n, 0);
        for (int i = 0; i < n; ++i) if (len[i]) sym[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(BitSrc& b) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; ++l) {
            code |= (int)b.bits(1);
            if (b.bad) return -1;
            const int c = count[l];
            if (code - c < first) return sym[(size_t)(index + (code - first))];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};
inline bool inflate_raw(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    BitSrc b(src, n);
    for (;;) {
        const uint32_t final = b.bits(1), type = b.bits(2);
        if (b.bad) return false;
        if (type == 0) {
            b.align();
            if (b.pos + 4 > n) return false;
            const uint32_t len = b.p[b.pos] | (b.p[b.pos + 1] << 8), nlen = b.p[b.pos + 2] | (b.p[b.pos + 3] << 8);
            b.pos += 4;
            if ((len ^ 0xffffu) != nlen || b.pos + len > n) return false;
            out.insert(out.end(), b.p + b.pos, b.p + b.pos + len);
            b.pos += len;
        } else if (type == 1 || type == 2) {
            HuffTable lit, dist;
            uint8_t lens[320];
            if (type == 1) {
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                lit.build(lens, 288);
                for (int i = 0; i < 30; ++i) lens[i] = 5;
                dist.build(lens, 30);
            } else {
                const int nlit = (int)b.bits(5) + 257, ndist = (int)b.bits(5) + 1, ncode = (int)b.bits(4) + 4;
                if (b.bad || nlit > 286 || ndist > 30) return false;
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)b.bits(3);
                HuffTable clt;
                clt.build(cl, 19);
                int i = 0;
                while (i < nlit + ndist) {
                    const int s = clt.decode(b);
                    if (s < 0) return false;
                    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                    int rep; uint8_t val = 0;
                    if (s == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + (int)b.bits(2); }
                    else if (s == 17) rep = 3 + (int)b.bits(3);
                    else rep = 11 + (int)b.bits(7);
                    if (b.bad || i + rep > nlit + ndist) return false;
                    while (rep--) lens[i++] = val;
                }
                lit.build(lens, nlit);
                dist.build(lens + nlit, ndist);
            }
            for (;;) {
                const int s = lit.decode(b);
                if (s < 0) return false;
                if (s < 256) { out.push_back((uint8_t)s); continue; }
                if (s == 256) break;
                if (s > 285) return false;
                const int len = len_base[s - 257] + (int)b.bits(len_extra[s - 257]);
                const int ds = dist.decode(b);
                if (ds < 0 || ds > 29) return false;
                const size_t d = (size_t)dist_base[ds] + b.bits(dist_extra[ds]);
                if (b.bad || d > out.size()) return false;
                const size_t from = out.size() - d;
                for (int k = 0; k < len; ++k) out.push_back(out[from + (size_t)k]);  // may overlap: byte by byte
            }
        } else {
            return false;
        }
        if (final) return true;
    }
}

// ------------------------------------------------------------------------------------ PNG
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool load_png(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (f.size() < 8 || std::memcmp(f.data(), sig, 8) != 0) return false;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t pos = 8;
    while (pos + 12 <= f.size()) {
        const uint32_t len = be32(&f[pos]);
        const uint8_t* type = &f[pos + 4];
        const uint8_t* data = &f[pos + 8];
        if (pos + 12 + (size_t)len > f.size()) { why = "truncated PNG chunk"; return false; }
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) { w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; }
        else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || w > 65535u || h > 65535u) { why = "bad PNG header"; return false; }
    if (interlace > 1) { why = "unknown PNG interlace method"; return false; }
    const int chan = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depth_ok = depth == 8 || (depth == 16 && ctype != 3) || ((depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3));
    if (!chan || !depth_ok) { why = "unsupported PNG colour type / bit depth"; return false; }
    if ((uint64_t)w * (uint64_t)h > (1ull << 28)) { why = "PNG too large"; return false; }
    if (idat.size() < 6) { why = "PNG without image data"; return false; }
    std::vector<uint8_t> raw;   // grows with the data actually present: the header alone never sizes an allocation
    if (!inflate_raw(idat.data() + 2, idat.size() - 2, raw)) { why = "corrupt PNG data stream"; return false; }  // 2-byte zlib header; Adler-32 not checked
    const size_t bpp = (size_t)((chan * depth + 7) / 8);
    // colour-key transparency of grey / RGB images (tRNS holds ONE colour, 16 bits per channel): compared at the full
    // 16 bits for 16-bit images, else on the low byte scaled like the samples (stb_image.h:5155-5166,5218-5226)
    const bool keyed = (ctype == 0 || ctype == 2) && !trns.empty();
    uint32_t key[3] = {0, 0, 0};
    if (keyed) {
        if (trns.size() != (size_t)2 * chan) { why = "bad PNG tRNS chunk"; return false; }
        for (int c = 0; c < chan; ++c) {
            const uint32_t v = ((uint32_t)trns[2 * (size_t)c] << 8) | trns[2 * (size_t)c + 1];
            key[c] = depth == 16 ? v : (uint8_t)((v & 255u) * (depth == 8 ? 1u : 255u / ((1u << depth) - 1u)));
        }
    }
    std::vector<uint8_t> top((size_t)4 * w * h), prev;
    // one pass = a sub-image with its own scanlines: the whole image, or the seven Adam7 passes {x0, y0, dx, dy}
    static const int adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const int whole[1][4] = {{0, 0, 1, 1}};
    const int (*passes)[4] = interlace ? adam7 : whole;
    size_t pos_raw = 0;
    for (int pass = 0; pass < (interlace ? 7 : 1); ++pass) {
        const uint32_t x0 = (uint32_t)passes[pass][0], y0 = (uint32_t)passes[pass][1], dx = (uint32_t)passes[pass][2], dy = (uint32_t)passes[pass][3];
        if (x0 >= w || y0 >= h) continue;
        const uint32_t pw = (w - x0 + dx - 1) / dx, ph = (h - y0 + dy - 1) / dy;
        const size_t stride = ((size_t)pw * chan * depth + 7) / 8;
        if (raw.size() < pos_raw + (stride + 1) * ph) { why = "short PNG data stream"; return false; }
        prev.assign(stride, 0);
        for (uint32_t j = 0; j < ph; ++j) {
            uint8_t* row = &raw[pos_raw + (stride + 1) * j + 1];
            const int ft = raw[pos_raw + (stride + 1) * j];
            for (size_t i = 0; i < stride; ++i) {  // un-filter in place (PNG spec 9.2)
                const int a = i >= bpp ? row[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int pred = 0;
                if (ft == 1) pred = a;
                else if (ft == 2) pred = b;
                else if (ft == 3) pred = (a + b) >> 1;
                else if (ft == 4) { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
                else if (ft != 0) { why = "bad PNG filter type"; return false; }
                row[i] = (uint8_t)(row[i] + pred);
            }
            std::memcpy(prev.data(), row, stride);
            const uint32_t y = y0 + j * dy;
            for (uint32_t x = 0; x < pw; ++x) {
                uint8_t* d = &top[4 * ((size_t)y * w + (x0 + x * dx))];
                auto sample = [&](int c) -> int {  // channel c of pixel x of this pass as 8 bit (16 bit: high byte)
                    if (depth == 8) return row[(size_t)x * chan + c];
                    if (depth == 16) return row[((size_t)x * chan + c) * 2];
                    const int per = 8 / depth, v = (row[x / per] >> ((per - 1 - (int)(x % per)) * depth)) & ((1 << depth) - 1);
                    return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
                };
                auto full = [&](int c) -> uint32_t {  // the value the colour key is compared with
                    if (depth == 16) return ((uint32_t)row[((size_t)x * chan + c) * 2] << 8) | row[((size_t)x * chan + c) * 2 + 1];
                    return (uint32_t)sample(c);
                };
                if (ctype == 0) { const int g = sample(0); d[0] = d[1] = d[2] = (uint8_t)g; d[3] = (keyed && full(0) == key[0]) ? 0 : 255; }
                else if (ctype == 2) {
                    d[0] = (uint8_t)sample(0); d[1] = (uint8_t)sample(1); d[2] = (uint8_t)sample(2);
                    d[3] = (keyed && full(0) == key[0] && full(1) == key[1] && full(2) == key[2]) ? 0 : 255;
                }
                else if (ctype == 4) { const int g = sample(0); d[0] = d[1] = d[2] = (uint8_t)g; d[3] = (uint8_t)sample(1); }
                else if (ctype == 6) { d[0] = (uint8_t)sample(0); d[1] = (uint8_t)sample(1); d[2] = (uint8_t)sample(2); d[3] = (uint8_t)sample(3); }
                else {
                    const size_t k = (size_t)sample(0);
                    if (3 * k + 2 >= plte.size()) { why = "PNG palette index out of range"; return false; }
                    d[0] = plte[3 * k]; d[1] = plte[3 * k + 1]; d[2] = plte[3 * k + 2];
                    d[3] = k < trns.size() ? trns[k] : 255;
                }
            }
        }
        pos_raw += (stride + 1) * ph;
    }
    store_flipped(top, (int)w, (int)h, t);
    return true;
}

// ------------------------------------------------------------------------------------ BMP
// Uncompressed Windows / OS2 bitmaps the way stb_image reads them (support/stb/stb_image.h:5445-5720): core (12), info
// (40 / 56) and V4 / V5 (108 / 124) headers; 1 / 4 / 8-bit palettised, 16-bit (5-5-5 or BI_BITFIELDS), 24-bit, 32-bit
// (BGRA or BI_BITFIELDS); masked channels widened to 8 bits by bit replication; an alpha channel that is zero everywhere
// is taken as opaque; positive height = bottom-up.  RLE-compressed bitmaps are refused (stb refuses them too).
inline bool load_bmp(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    if (f.size() < 2 || f[0] != 'B' || f[1] != 'M') return false;
    size_t pos = 2;
    auto g8 = [&]() -> uint32_t { return pos < f.size() ? f[pos++] : 0u; };
    auto g16 = [&]() -> uint32_t { const uint32_t a = g8(); return a | (g8() << 8); };
    auto g32 = [&]() -> uint32_t { const uint32_t a = g16(); return a | (g16() << 16); };
    auto skip = [&](int k) { if (k < 0) pos = f.size(); else pos = pos + (size_t)k > f.size() ? f.size() : pos + (size_t)k; };
    g32(); g16(); g16();
    const int offset = (int)g32(), hsz = (int)g32();
    uint32_t mr = 0, mg = 0, mb = 0, ma = 0, all_a = 255;
    int extra_read = 14, W, Hs;
    if (offset < 0) { why = "bad BMP"; return false; }
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) { why = "unknown BMP header"; return false; }
    if (hsz == 12) { W = (int)g16(); Hs = (int)g16(); } else { W = (int)g32(); Hs = (int)g32(); }
    if (g16() != 1) { why = "bad BMP"; return false; }
    const int bpp = (int)g16();
    auto default_masks = [&](int compress) {
        if (compress != 0) return;
        if (bpp == 16) { mr = 31u << 10; mg = 31u << 5; mb = 31u; }
        else if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; all_a = 0; }
        else mr = mg = mb = ma = 0;
    };
    if (hsz != 12) {
        const int compress = (int)g32();
        if (compress == 1 || compress == 2) { why = "RLE-compressed BMP is not supported"; return false; }
        if (compress >= 4 || compress < 0) { why = "BMP with embedded JPEG / PNG is not supported"; return false; }
        if (compress == 3 && bpp != 16 && bpp != 32) { why = "bad BMP"; return false; }
        g32(); g32(); g32(); g32(); g32();
        if (hsz == 40 || hsz == 56) {
            if (hsz == 56) { g32(); g32(); g32(); g32(); }
            if (bpp == 16 || bpp == 32) {
                if (compress == 0) default_masks(0);
                else {
                    mr = g32(); mg = g32(); mb = g32();
                    extra_read += 12;
                    if (mr == mg && mg == mb) { why = "bad BMP"; return false; }
                }
            }
        } else {
            mr = g32(); mg = g32(); mb = g32(); ma = g32();
            if (compress != 3) default_masks(compress);
            for (int i = 0; i < 13; ++i) g32();
            if (hsz == 124) { g32(); g32(); g32(); g32(); }
        }
    }
    const bool bottom_up = Hs > 0;
    const int H = Hs < 0 ? -Hs : Hs;
    if (W <= 0 || H <= 0 || W > (1 << 24) || H > (1 << 24) || (uint64_t)W * (uint64_t)H > (1ull << 28)) { why = "bad BMP size"; return false; }
    int psize = 0;
    if (hsz == 12) { if (bpp < 24) psize = (offset - extra_read - 24) / 3; }
    else if (bpp < 16) psize = (offset - extra_read - hsz) >> 2;
    if (psize == 0) {
        const int so_far = (int)pos;
        if (so_far <= 0 || so_far > 1024 || offset < so_far || offset - so_far > 1024) { why = "corrupt BMP (data offset)"; return false; }
        skip(offset - so_far);
    }
    std::vector<uint8_t> rows((size_t)4 * W * H);   // in file order
    size_t z = 0;
    if (bpp < 16) {
        if (psize == 0 || psize > 256) { why = "corrupt BMP (palette)"; return false; }
        uint8_t pal[256][3];
        std::memset(pal, 0, sizeof(pal));
        for (int i = 0; i < psize; ++i) { pal[i][2] = (uint8_t)g8(); pal[i][1] = (uint8_t)g8(); pal[i][0] = (uint8_t)g8(); if (hsz != 12) g8(); }
        skip(offset - extra_read - hsz - psize * (hsz == 12 ? 3 : 4));
        int width;
        if (bpp == 1) width = (W + 7) >> 3; else if (bpp == 4) width = (W + 1) >> 1; else if (bpp == 8) width = W; else { why = "corrupt BMP (bits per pixel)"; return false; }
        const int pad = (-width) & 3;
        for (int j = 0; j < H; ++j) {
            uint32_t v = 0;
            for (int i = 0; i < W; ++i) {
                uint32_t c;
                if (bpp == 1) { if ((i & 7) == 0) v = g8(); c = (v >> (7 - (i & 7))) & 1u; }
                else if (bpp == 4) { if ((i & 1) == 0) { v = g8(); c = v >> 4; } else c = v & 15u; }
                else c = g8();
                rows[z++] = pal[c][0]; rows[z++] = pal[c][1]; rows[z++] = pal[c][2]; rows[z++] = 255;
            }
            skip(pad);
        }
    } else {
        skip(offset - extra_read - hsz);
        const int pad = bpp == 24 ? (-(3 * W)) & 3 : (bpp == 16 ? (-(2 * W)) & 3 : 0);
        const int easy = bpp == 24 ? 1 : ((bpp == 32 && mb == 0xffu && mg == 0xff00u && mr == 0x00ff0000u && ma == 0xff000000u) ? 2 : 0);
        if (bpp != 16 && bpp != 24 && bpp != 32) { why = "corrupt BMP (bits per pixel)"; return false; }
        auto high_bit = [](uint32_t m) { int n = -1; while (m) { ++n; m >>= 1; } return n; };
        auto bit_count = [](uint32_t m) { int n = 0; while (m) { n += (int)(m & 1u); m >>= 1; } return n; };
        // the masked field, top-aligned to 8 bits and filled up by repeating its bit pattern
        auto widen = [](uint32_t v, int shift, int bits) -> uint32_t {
            static const uint32_t mul[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
            static const uint32_t shr[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
            if (shift < 0) v <<= -shift; else v >>= shift;
            v >>= (8 - bits);
            return (v * mul[bits]) >> shr[bits];
        };
        int rs = 0, gs = 0, bs = 0, as = 0, rc = 0, gc = 0, bc = 0, ac = 0;
        if (!easy) {
            if (!mr || !mg || !mb) { why = "corrupt BMP (masks)"; return false; }
            rs = high_bit(mr) - 7; rc = bit_count(mr); gs = high_bit(mg) - 7; gc = bit_count(mg);
            bs = high_bit(mb) - 7; bc = bit_count(mb); as = high_bit(ma) - 7; ac = bit_count(ma);
            if (rc > 8 || gc > 8 || bc > 8 || ac > 8) { why = "corrupt BMP (masks)"; return false; }
        }
        for (int j = 0; j < H; ++j) {
            for (int i = 0; i < W; ++i) {
                uint32_t a;
                if (easy) {
                    rows[z + 2] = (uint8_t)g8(); rows[z + 1] = (uint8_t)g8(); rows[z] = (uint8_t)g8();
                    a = easy == 2 ? g8() : 255u;
                } else {
                    const uint32_t v = bpp == 16 ? g16() : g32();
                    rows[z] = (uint8_t)widen(v & mr, rs, rc); rows[z + 1] = (uint8_t)widen(v & mg, gs, gc); rows[z + 2] = (uint8_t)widen(v & mb, bs, bc);
                    a = ma ? widen(v & ma, as, ac) : 255u;
                }
                all_a |= a;
                rows[z + 3] = (uint8_t)a;
                z += 4;
            }
            skip(pad);
        }
    }
    if (all_a == 0) for (size_t i = 3; i < rows.size(); i += 4) rows[i] = 255;
    t.width = W; t.height = H;
    t.pixel.resize(rows.size());   // the texture wants the bottom row first, which is the file order of a bottom-up bitmap
    for (int y = 0; y < H; ++y) std::memcpy(&t.pixel[(size_t)4 * W * y], &rows[(size_t)4 * W * (size_t)(bottom_up ? y : H - 1 - y)], (size_t)4 * W);
    return true;
}

// ------------------------------------------------------------------------------------ TGA
// Truevision files as stb_image reads them (support/stb/stb_image.h:5733-6060): true-colour 15 / 16 / 24 / 32 bit, grey 8
// bit and grey + alpha 16 bit, colour-mapped with 8 / 16-bit indices into a 15 / 16 / 24 / 32-bit palette, raw or RLE;
// 5-bit channels scale as c * 255 / 31; rows bottom-up unless descriptor bit 5 is set; the right-to-left bit is ignored
// (stb ignores it).  TGA has no magic number: tga_plausible() is the header check that stands in for one.
inline bool tga_plausible(const std::vector<uint8_t>& f) {
    if (f.size() < 18) return false;
    const int cmap = f[1], type = f[2], bpp = f[16];
    if (cmap > 1) return false;
    if (cmap == 1) {
        if (type != 1 && type != 9) return false;
        const int pb = f[7];
        if (pb != 8 && pb != 15 && pb != 16 && pb != 24 && pb != 32) return false;
    } else if (type != 2 && type != 3 && type != 10 && type != 11) return false;
    if ((f[12] | (f[13] << 8)) < 1 || (f[14] | (f[15] << 8)) < 1) return false;
    if (cmap == 1 && bpp != 8 && bpp != 16) return false;
    return bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32;
}
inline bool load_tga(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    if (!tga_plausible(f)) return false;
    size_t pos = 0;
    auto g8 = [&]() -> int { return pos < f.size() ? f[pos++] : 0; };
    auto g16 = [&]() -> int { const int a = g8(); return a | (g8() << 8); };
    const int idlen = g8(), indexed = g8();
    int type = g8();
    const int pal_start = g16(), pal_len = g16(), pal_bits = g8();
    g16(); g16();
    const int w = g16(), h = g16(), bpp = g8(), desc = g8();
    const bool rle = type >= 8;
    if (rle) type -= 8;
    const bool bottom_up = ((desc >> 5) & 1) == 0;
    // components of a pixel as stored / looked up: 1 grey, 2 grey + alpha, 3 or 4 colour; 15 / 16-bit colour expands to 3
    auto comps = [](int bits, bool grey, bool& rgb16) { rgb16 = false; if (bits == 8) return 1; if (bits == 16 && grey) return 2; if (bits == 15 || bits == 16) { rgb16 = true; return 3; } return (bits == 24 || bits == 32) ? bits / 8 : 0; };
    bool rgb16 = false;
    const int nc = indexed ? comps(pal_bits, false, rgb16) : comps(bpp, type == 3, rgb16);
    if (!nc) { why = "unsupported TGA pixel format"; return false; }
    if ((uint64_t)w * (uint64_t)h > (1ull << 28)) { why = "TGA too large"; return false; }
    pos += (size_t)idlen;
    auto read16 = [&](uint8_t* o) { const int px = g16(); o[0] = (uint8_t)((((px >> 10) & 31) * 255) / 31); o[1] = (uint8_t)((((px >> 5) & 31) * 255) / 31); o[2] = (uint8_t)(((px & 31) * 255) / 31); };
    std::vector<uint8_t> pal;
    if (indexed) {
        if (pal_len == 0) { why = "corrupt TGA (empty palette)"; return false; }
        pos += (size_t)pal_start;
        pal.assign((size_t)pal_len * nc, 0);
        if (rgb16) for (int i = 0; i < pal_len; ++i) read16(&pal[(size_t)i * 3]);
        else {
            if (pos + pal.size() > f.size()) { why = "truncated TGA"; return false; }
            std::memcpy(pal.data(), &f[pos], pal.size());
            pos += pal.size();
        }
    }
    std::vector<uint8_t> data((size_t)w * h * nc);   // pixels in file order
    uint8_t px[4] = {0, 0, 0, 0};
    int run = 0;
    bool repeating = false;
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        bool fetch = true;
        if (rle) {
            if (run == 0) { const int c = g8(); run = 1 + (c & 127); repeating = (c >> 7) != 0; }
            else if (repeating) fetch = false;
        }
        if (fetch) {
            if (indexed) {
                int k = bpp == 8 ? g8() : g16();
                if (k >= pal_len) k = 0;
                std::memcpy(px, &pal[(size_t)k * nc], (size_t)nc);
            } else if (rgb16) read16(px);
            else for (int j = 0; j < nc; ++j) px[j] = (uint8_t)g8();
        }
        std::memcpy(&data[i * nc], px, (size_t)nc);
        --run;
    }
    t.width = w; t.height = h;
    t.pixel.resize((size_t)4 * w * h);
    for (int y = 0; y < h; ++y)   // texture row 0 = image bottom = the first row of a bottom-up file
        for (int x = 0; x < w; ++x) {
            const uint8_t* s = &data[((size_t)(bottom_up ? y : h - 1 - y) * w + (size_t)x) * nc];
            uint8_t* d = &t.pixel[4 * ((size_t)y * w + x)];
            if (nc == 1) { d[0] = d[1] = d[2] = s[0]; d[3] = 255; }
            else if (nc == 2) { d[0] = d[1] = d[2] = s[0]; d[3] = s[1]; }
            else if (rgb16) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = 255; }
            else { d[0] = s[2]; d[1] = s[1]; d[2] = s[0]; d[3] = nc == 4 ? s[3] : 255; }
        }
    return true;
}

// ------------------------------------------------------------------------------------ PPM / PGM
// Binary P5 / P6 as stb_image reads them (support/stb/stb_image.h:7503-7621): samples are taken as they are, whatever
// maxval says (no rescaling); with maxval > 255 a sample is two bytes and the texture gets the SECOND one — stb reads the
// big-endian pairs as host-order 16-bit words and keeps their upper half, which on the little-endian hosts the reference
// runs on is the low-order byte of the sample.
inline bool load_pnm(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    if (f.size() < 3 || f[0] != 'P' || (f[1] != '6' && f[1] != '5')) return false;
    const int chan = f[1] == '6' ? 3 : 1;
    size_t pos = 2;
    auto eof = [&]() { return pos >= f.size(); };
    auto g8 = [&]() -> char { return (char)(pos < f.size() ? f[pos++] : 0); };
    char c = g8();
    auto is_space = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\v' || ch == '\f' || ch == '\r'; };
    auto skip_space = [&]() {
        for (;;) {
            while (!eof() && is_space(c)) c = g8();
            if (eof() || c != '#') break;
            while (!eof() && c != '\n' && c != '\r') c = g8();
        }
    };
    auto integer = [&]() -> long {
        long v = 0;
        while (!eof() && c >= '0' && c <= '9') { v = v * 10 + (c - '0'); c = g8(); if (v > 214748364L) return -1; }
        return v;
    };
    skip_space();
    const long w = integer();
    skip_space();
    const long h = integer();
    skip_space();
    const long maxv = integer();
    if (w <= 0 || h <= 0 || maxv < 0 || maxv > 65535) { why = "bad PNM header"; return false; }
    const size_t bytes = maxv > 255 ? 2 : 1;
    if ((uint64_t)w * (uint64_t)h > (1ull << 28)) { why = "PNM too large"; return false; }
    if (pos + (size_t)w * h * chan * bytes > f.size()) { why = "truncated PNM"; return false; }
    std::vector<uint8_t> top((size_t)4 * w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const uint8_t* s = &f[pos + i * chan * bytes + (bytes - 1)];
        uint8_t* d = &top[4 * i];
        d[0] = s[0]; d[1] = s[chan == 3 ? bytes : 0]; d[2] = s[chan == 3 ? 2 * bytes : 0]; d[3] = 255;
    }
    store_flipped(top, (int)w, (int)h, t);
    return true;
}

// ------------------------------------------------------------------------------------ GIF / PSD / PIC / Radiance HDR
// The remaining formats stbi_load accepts (support/stb/stb_image.h v2.29), restated with its lenient reading rules so that
// the same files — damaged ones included — give the same RGBA bytes: a read past the end of the file yields zero bytes.
struct ByteCursor {
    const uint8_t* p; size_t n, pos = 0;
    explicit ByteCursor(const std::vector<uint8_t>& f) : p(f.data()), n(f.size()) {}
    bool eof() const { return pos >= n; }
    int u8() { return pos < n ? p[pos++] : 0; }
    int u16le() { const int a = u8(); return a | (u8() << 8); }
    int u16be() { const int a = u8(); return (a << 8) | u8(); }
    uint32_t u32be() { const uint32_t a = (uint32_t)u16be(); return (a << 16) | (uint32_t)u16be(); }
    void skip(long k) { pos = (k < 0 || (size_t)k > n - (pos < n ? pos : n)) ? n : pos + (size_t)k; }
};
// stb lets these four formats declare a width or height of zero (a texture without texels, refused later at upload)
inline bool image_size_ok(long w, long h) { return w >= 0 && h >= 0 && w <= (1 << 24) && h <= (1 << 24) && (uint64_t)w * (uint64_t)h <= (1ull << 28); }

// GIF87a / GIF89a, first image only (stb_image.h:6575-6947).  stb's choices, all visible in the texture: pixels of the
// transparent index are not drawn (they stay 0,0,0,0); codes beyond the colour table are transparent; when the background
// index is non-zero the pixels the first image does not cover get the background entry WITH RED AND BLUE EXCHANGED (stb
// keeps its palettes as B,G,R,A and copies that entry unconverted, :6895-6903); interlaced rows in the 8/8/4/2 order; a
// raster must start with a clear code; a data stream that ends early keeps what was drawn.
inline bool load_gif(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    ByteCursor s(f);
    if (s.u8() != 'G' || s.u8() != 'I' || s.u8() != 'F' || s.u8() != '8') return false;
    const int version = s.u8();
    if ((version != '7' && version != '9') || s.u8() != 'a') return false;
    const int W = s.u16le(), H = s.u16le(), flags = s.u8(), bgindex = s.u8();
    s.u8();   // aspect ratio
    if (!image_size_ok(W, H)) { why = "bad GIF size"; return false; }
    struct Rgba { uint8_t r, g, b, a; };
    std::vector<Rgba> global(256, Rgba{0, 0, 0, 0}), local(256, Rgba{0, 0, 0, 0});
    auto read_table = [&](std::vector<Rgba>& pal, int entries, int transparent) {
        for (int i = 0; i < entries; ++i) {
            pal[(size_t)i].r = (uint8_t)s.u8(); pal[(size_t)i].g = (uint8_t)s.u8(); pal[(size_t)i].b = (uint8_t)s.u8();
            pal[(size_t)i].a = transparent == i ? 0 : 255;
        }
    };
    if (flags & 0x80) read_table(global, 2 << (flags & 7), -1);
    std::vector<uint8_t> top((size_t)4 * W * H, 0), drawn((size_t)W * H, 0);
    int eflags = 0, transparent = -1;
    for (;;) {
        const int tag = s.u8();
        if (tag == 0x21) {   // extension; only the graphic control block matters (transparent index)
            const int ext = s.u8();
            if (ext == 0xF9) {
                const int len = s.u8();
                if (len != 4) { s.skip(len); continue; }   // stb goes back to the tag loop here, without walking the sub-blocks
                eflags = s.u8();
                s.u16le();   // delay
                if (transparent >= 0) global[(size_t)transparent].a = 255;
                if (eflags & 1) { transparent = s.u8(); global[(size_t)transparent].a = 0; }
                else { s.skip(1); transparent = -1; }
            }
            for (int len; (len = s.u8()) != 0;) s.skip(len);
            continue;
        }
        if (tag != 0x2C) { why = tag == 0x3B ? "GIF without an image" : "corrupt GIF"; return false; }
        break;
    }
    const int x0 = s.u16le(), y0 = s.u16le(), w = s.u16le(), h = s.u16le();
    if (x0 + w > W || y0 + h > H) { why = "bad GIF image descriptor"; return false; }
    const int lflags = s.u8();
    // byte offsets into `top`, the way stb walks them
    const long line = 4L * W, start_x = 4L * x0, start_y = (long)y0 * line, max_x = start_x + 4L * w, max_y = start_y + (long)h * line;
    long cur_x = start_x, cur_y = w == 0 ? max_y : start_y, step = (lflags & 0x40) ? 8 * line : line;
    int pass = (lflags & 0x40) ? 3 : 0;
    const std::vector<Rgba>* table;
    if (lflags & 0x80) { read_table(local, 2 << (lflags & 7), (eflags & 1) ? transparent : -1); table = &local; }
    else if (flags & 0x80) table = &global;
    else { why = "GIF without a colour table"; return false; }
    auto put = [&](int index) {
        if (cur_y >= max_y) return;
        const size_t at = (size_t)(cur_x + cur_y);
        drawn[at / 4] = 1;
        const Rgba c = (*table)[(size_t)index];
        if (c.a > 128) { top[at] = c.r; top[at + 1] = c.g; top[at + 2] = c.b; top[at + 3] = c.a; }
        cur_x += 4;
        if (cur_x >= max_x) {
            cur_x = start_x;
            cur_y += step;
            while (cur_y >= max_y && pass > 0) { step = (1L << pass) * line; cur_y = start_y + (step >> 1); --pass; }
        }
    };
    // LZW (variable code width, LSB first, data in sub-blocks)
    const int min_bits = s.u8();
    if (min_bits > 12) { why = "corrupt GIF"; return false; }
    struct Code { int16_t prefix; uint8_t first, suffix; };
    std::vector<Code> codes(8192);
    const int clear = 1 << min_bits;
    for (int i = 0; i < clear; ++i) codes[(size_t)i] = Code{-1, (uint8_t)i, (uint8_t)i};
    int width = min_bits + 1, mask = (1 << width) - 1, avail = clear + 2, old = -1, have = 0, block = 0;
    int32_t acc = 0;
    bool seen_clear = false;
    std::vector<uint8_t> chain;
    for (;;) {
        if (have < width) {
            if (block == 0) { block = s.u8(); if (block == 0) break; }   // also where a truncated file ends
            --block;
            acc |= (int32_t)((uint32_t)s.u8() << have);
            have += 8;
            continue;
        }
        const int code = acc & mask;
        acc >>= width; have -= width;
        if (code == clear) { width = min_bits + 1; mask = (1 << width) - 1; avail = clear + 2; old = -1; seen_clear = true; continue; }
        if (code == clear + 1) break;   // end of information (the trailing sub-blocks are of no interest here)
        if (code > avail) { why = "corrupt GIF (illegal code)"; return false; }
        if (!seen_clear) { why = "corrupt GIF (no clear code)"; return false; }
        if (old >= 0) {
            if (avail + 1 > 8192) { why = "corrupt GIF (too many codes)"; return false; }
            Code& c = codes[(size_t)avail++];
            c.prefix = (int16_t)old;
            c.first = codes[(size_t)old].first;
            c.suffix = code == avail ? c.first : codes[(size_t)code].first;
        } else if (code == avail) { why = "corrupt GIF (illegal code)"; return false; }
        chain.clear();
        for (int k = code; k >= 0; k = codes[(size_t)k].prefix) chain.push_back(codes[(size_t)k].suffix);
        for (size_t k = chain.size(); k-- > 0;) put(chain[k]);
        if ((avail & mask) == 0 && avail <= 0x0FFF) { ++width; mask = (1 << width) - 1; }
        old = code;
    }
    if (bgindex > 0) {
        const Rgba bg = global[(size_t)bgindex];
        for (size_t i = 0; i < drawn.size(); ++i)
            if (!drawn[i]) { top[4 * i] = bg.b; top[4 * i + 1] = bg.g; top[4 * i + 2] = bg.r; top[4 * i + 3] = 255; }
    }
    store_flipped(top, W, H, t);
    return true;
}

// Photoshop PSD, the flattened composite only (stb_image.h:6078-6326): version 1, RGB colour mode, 8 or 16 bits per
// channel (the high byte is kept), raw or PackBits rows, channel planes R, G, B, A in that order (missing ones: 0, alpha
// 255, extra ones ignored).  With four or more channels stb "removes the white matte": for 0 < a < 255 every colour byte
// becomes c / a' + 255 (1 - 1 / a'), a' = a / 255, in float, truncated and reduced modulo 256.
inline bool load_psd(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    ByteCursor s(f);
    if (s.u32be() != 0x38425053u) return false;
    if (s.u16be() != 1) { why = "unsupported PSD version"; return false; }
    s.skip(6);
    const int channels = s.u16be();
    if (channels > 16) { why = "unsupported number of PSD channels"; return false; }
    const uint32_t H = s.u32be(), W = s.u32be();
    if (H > (1u << 24) || W > (1u << 24) || !image_size_ok((long)W, (long)H)) { why = "bad PSD size"; return false; }
    const int depth = s.u16be();
    if (depth != 8 && depth != 16) { why = "PSD bit depth is not 8 or 16"; return false; }
    if (s.u16be() != 3) { why = "PSD is not in RGB colour mode"; return false; }
    for (int k = 0; k < 3; ++k) s.skip((long)(int32_t)s.u32be());   // mode data, image resources, layer and mask information
    const int compression = s.u16be();
    if (compression > 1) { why = "unknown PSD compression"; return false; }
    const size_t count = (size_t)W * H;
    std::vector<uint8_t> top(4 * count);
    if (count == 0) { store_flipped(top, (int)W, (int)H, t); return true; }
    if (compression) s.skip((long)H * channels * 2);   // byte counts of the packed rows
    for (int ch = 0; ch < 4; ++ch) {
        uint8_t* d = top.data() + ch;
        if (ch >= channels) { for (size_t i = 0; i < count; ++i) d[4 * i] = ch == 3 ? 255 : 0; continue; }
        if (!compression) {
            for (size_t i = 0; i < count; ++i) d[4 * i] = (uint8_t)(depth == 16 ? s.u16be() >> 8 : s.u8());
            continue;
        }
        // PackBits over the whole plane, bytes whatever the depth says (stb decodes packed 16-bit files this way too)
        for (size_t done = 0; done < count;) {
            int len = s.u8();
            if (len == 128) continue;
            const bool run = len > 128;
            len = run ? 257 - len : len + 1;
            if ((size_t)len > count - done) { why = "bad PSD RLE data"; return false; }
            const int v = run ? s.u8() : 0;
            for (int k = 0; k < len; ++k) d[4 * (done + (size_t)k)] = (uint8_t)(run ? v : s.u8());
            done += (size_t)len;
        }
    }
    if (channels >= 4)
        for (size_t i = 0; i < count; ++i) {
            uint8_t* px = &top[4 * i];
            if (px[3] == 0 || px[3] == 255) continue;
            const float a = px[3] / 255.0f, ra = 1.0f / a, inv_a = 255.0f * (1 - ra);
            for (int c = 0; c < 3; ++c) px[c] = (uint8_t)(int32_t)(px[c] * ra + inv_a);
        }
    store_flipped(top, (int)W, (int)H, t);
    return true;
}

// Softimage PIC (stb_image.h:6333-6536): up to ten 8-bit channel packets, each raw, run-length or mixed run-length coded,
// applied row by row to a picture preset to 255.
inline bool pic_plausible(const std::vector<uint8_t>& f) {
    return f.size() >= 92 && f[0] == 0x53 && f[1] == 0x80 && f[2] == 0xF6 && f[3] == 0x34 && std::memcmp(&f[88], "PICT", 4) == 0;
}
inline bool load_pic(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    ByteCursor s(f);
    s.skip(92);
    const int W = s.u16be(), H = s.u16be();
    if (s.eof() || !image_size_ok(W, H)) { why = "bad PIC header"; return false; }
    s.skip(8);   // ratio, fields, pad
    struct Packet { int type, channels; };
    std::vector<Packet> packets;
    for (int chained = 1; chained;) {
        if (packets.size() == 10) { why = "PIC: too many packets"; return false; }
        chained = s.u8();
        const int size = s.u8();
        Packet pk;
        pk.type = s.u8(); pk.channels = s.u8();
        packets.push_back(pk);
        if (s.eof()) { why = "PIC file too short"; return false; }
        if (size != 8) { why = "PIC packet is not 8 bits per channel"; return false; }
    }
    std::vector<uint8_t> top((size_t)4 * W * H, 255);
    if (top.empty()) { store_flipped(top, W, H, t); return true; }   // no texels: the rows read nothing
    bool short_file = false;
    auto read_value = [&](int channels, uint8_t* d) {   // 0x80 red, 0x40 green, 0x20 blue, 0x10 alpha
        for (int i = 0; i < 4; ++i)
            if (channels & (0x80 >> i)) {
                if (s.eof()) { short_file = true; return; }
                d[i] = (uint8_t)s.u8();
            }
    };
    auto copy_value = [](int channels, uint8_t* d, const uint8_t* v) {
        for (int i = 0; i < 4; ++i) if (channels & (0x80 >> i)) d[i] = v[i];
    };
    for (int y = 0; y < H; ++y)
        for (const Packet& pk : packets) {
            uint8_t* d = &top[(size_t)4 * W * y];
            if (pk.type == 0) {
                for (int x = 0; x < W && !short_file; ++x, d += 4) read_value(pk.channels, d);
            } else if (pk.type == 1) {   // runs only; a run that overshoots the row is cut
                for (int left = W; left > 0 && !short_file;) {
                    int count = s.u8();
                    if (s.eof()) { short_file = true; break; }
                    if (count > left) count = (uint8_t)left;
                    uint8_t v[4] = {0, 0, 0, 0};
                    read_value(pk.channels, v);
                    if (short_file) break;
                    for (int i = 0; i < count; ++i, d += 4) copy_value(pk.channels, d, v);
                    left -= count;
                }
            } else if (pk.type == 2) {   // runs and literal stretches
                for (int left = W; left > 0 && !short_file;) {
                    int count = s.u8();
                    if (s.eof()) { short_file = true; break; }
                    if (count >= 128) {
                        count = count == 128 ? s.u16be() : count - 127;
                        if (count > left) { why = "PIC scanline overrun"; return false; }
                        uint8_t v[4] = {0, 0, 0, 0};
                        read_value(pk.channels, v);
                        if (short_file) break;
                        for (int i = 0; i < count; ++i, d += 4) copy_value(pk.channels, d, v);
                    } else {
                        ++count;
                        if (count > left) { why = "PIC scanline overrun"; return false; }
                        for (int i = 0; i < count && !short_file; ++i, d += 4) read_value(pk.channels, d);
                    }
                    left -= count;
                }
            } else { why = "PIC packet has a bad compression type"; return false; }
            if (short_file) { why = "PIC file too short"; return false; }
        }
    store_flipped(top, W, H, t);
    return true;
}

// Radiance RGBE (stb_image.h:7086-7272) brought to 8 bits the way stbi_load does for an HDR file (:1883-1907): linear value
// v = mantissa * 2^(exponent - 136), byte = (int)((float)pow(v, 1 / 2.2f) * 255 + 0.5f) clamped to 0..255, alpha 255.
// Header: "#?RADIANCE" or "#?RGBE", a FORMAT=32-bit_rle_rgbe line, a blank line, "-Y h +X w".  Rows are new-style run-length
// coded for widths 8..32767 (flat RGBE otherwise, or when the first row does not start with the 2 2 marker).
inline bool hdr_plausible(const std::vector<uint8_t>& f) {
    return (f.size() >= 11 && std::memcmp(f.data(), "#?RADIANCE\n", 11) == 0) || (f.size() >= 7 && std::memcmp(f.data(), "#?RGBE\n", 7) == 0);
}
inline bool load_hdr(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    ByteCursor s(f);
    auto line = [&]() {   // up to the next '\n'; at most 1022 characters are kept (a character that ends the file is dropped, as in stb)
        std::string out;
        char c = (char)s.u8();
        while (!s.eof() && c != '\n') {
            out.push_back(c);
            if (out.size() == 1023) { while (!s.eof() && s.u8() != '\n') {} break; }
            c = (char)s.u8();
        }
        return out;
    };
    line();   // the signature, checked by hdr_plausible
    bool rle_rgbe = false;
    for (;;) {
        const std::string tok = line();
        if (tok.empty()) break;
        if (tok == "FORMAT=32-bit_rle_rgbe") rle_rgbe = true;
    }
    if (!rle_rgbe) { why = "unsupported HDR format"; return false; }
    const std::string dims = line();
    if (dims.compare(0, 3, "-Y ") != 0) { why = "unsupported HDR data layout"; return false; }
    char* end = nullptr;
    const long H = std::strtol(dims.c_str() + 3, &end, 10);
    while (*end == ' ') ++end;
    if (std::strncmp(end, "+X ", 3) != 0) { why = "unsupported HDR data layout"; return false; }
    const long W = std::strtol(end + 3, nullptr, 10);
    if (!image_size_ok(W, H)) { why = "bad HDR size"; return false; }
    std::vector<uint8_t> rgbe((size_t)4 * W * H);
    // flat pixels are read four bytes at a time into one buffer; where the file ends, the bytes it no longer supplies keep the
    // previous pixel's values (stb's read leaves its buffer untouched there), so a cut file repeats its last pixel
    uint8_t quad[4] = {0, 0, 0, 0};
    auto flat_from = [&](size_t first) {
        for (size_t i = first; i < (size_t)W * H; ++i) {
            for (int k = 0; k < 4 && !s.eof(); ++k) quad[k] = (uint8_t)s.u8();
            std::memcpy(&rgbe[4 * i], quad, 4);
        }
    };
    if (W < 8 || W >= 32768) flat_from(0);
    else
        for (long y = 0; y < H; ++y) {
            const int c1 = s.u8(), c2 = s.u8(), hi = s.u8();
            if (c1 != 2 || c2 != 2 || (hi & 0x80)) {   // not run-length coded: these four bytes are the first pixel of a flat file
                // (on a later row too: stb then starts the picture over, flat, from this point of the file)
                quad[0] = (uint8_t)c1; quad[1] = (uint8_t)c2; quad[2] = (uint8_t)hi; quad[3] = (uint8_t)s.u8();
                std::memcpy(&rgbe[0], quad, 4);
                flat_from(1);
                break;
            }
            if (((hi << 8) | s.u8()) != W) { why = "corrupt HDR (scanline length)"; return false; }
            for (int k = 0; k < 4; ++k)
                for (long i = 0; i < W;) {
                    int count = s.u8();
                    const bool run = count > 128;
                    if (run) count -= 128;
                    const int v = run ? s.u8() : 0;
                    if (count == 0 || count > W - i) { why = "corrupt HDR (bad RLE data)"; return false; }
                    for (int z = 0; z < count; ++z, ++i) rgbe[4 * ((size_t)y * W + (size_t)i) + (size_t)k] = (uint8_t)(run ? v : s.u8());
                }
        }
    std::vector<uint8_t> top((size_t)4 * W * H);
    const double gamma = (double)(1.0f / 2.2f);
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        const uint8_t* in = &rgbe[4 * i];
        const float scale = in[3] ? (float)std::ldexp(1.0f, in[3] - 136) : 0.0f;
        for (int c = 0; c < 3; ++c) {
            const float linear = in[3] ? in[c] * scale : 0.0f;
            float z = (float)std::pow((double)(linear * 1.0f), gamma) * 255 + 0.5f;
            z = z < 0 ? 0 : (z > 255 ? 255 : z);
            top[4 * i + (size_t)c] = (uint8_t)(int)z;
        }
        top[4 * i + 3] = 255;
    }
    store_flipped(top, (int)W, (int)H, t);
    return true;
}

// ------------------------------------------------------------------------------------ JPEG (ITU T.81), stb_image's arithmetic
// The reference decodes map_Kd files with stbi_load (src/mesh.cpp:137); a JPEG decoder is not specified to the bit
// by T.81, so "the same texture bytes" means stb_image's choices (support/stb/stb_image.h), restated here and pinned
// against the reference's own loader (tests/test_loader_parity.py):
//   * coefficients dequantised into 16-bit integers while decoding (:2209-2262), progressive scans refined in place
//     and dequantised at the end (:2264-2412, :3072-3096);
//   * the 8x8 inverse DCT in integers: the "islow" factorisation with 12-bit constants, a column pass that keeps
//     2 extra bits (+512 >> 10), a row pass that removes 17 bits with the +128 level shift folded in (:2431-2517);
//   * chroma planes brought to full resolution row by row with jfif-centred filters: 3:1 taps vertically / horizontally
//     for factor 2, the 9:3:3:1 tent for 2x2, replication for every other ratio (:3453-3657, :3897-3945);
//   * YCbCr -> RGB in 20-bit fixed point with the 16 fractional bits of the Cb term of green masked off
//     (:3659-3688); RGB-tagged ('R','G','B' ids, or Adobe transform 0 without JFIF) and CMYK / YCCK files as stb
//     treats them (:3947-3981);
//   * an entropy segment that ends early yields zero bits, a missing restart marker ends the scan quietly (:2075-2092,
//     :2962-2969) — lenient like stb, so that the same damaged files load to the same bytes.
// 8-bit baseline, extended-sequential and progressive Huffman files with 1, 3 or 4 components; arithmetic coding and
// lossless / hierarchical frames are refused, as stb refuses them.
struct JpegHuff {
    uint8_t size[257], values[256];
    uint16_t code[256];
    uint32_t maxcode[18];   // per length: (largest code + 1) aligned to 16 bits
    int delta[17];          // per length: symbol index of the first code minus that code
    int16_t fast_ac[512];   // AC tables: (value << 8 | run << 4 | total length) for short codes with short magnitudes, else 0
    bool build(const int count[16]) {
        int k = 0;
        for (int l = 0; l < 16; ++l)
            for (int j = 0; j < count[l]; ++j) { size[k++] = (uint8_t)(l + 1); if (k >= 257) return false; }
        size[k] = 0;
        uint32_t c = 0;
        k = 0;
        for (int l = 1; l <= 16; ++l) {
            delta[l] = k - (int)c;
            if (size[k] == l) {
                while (size[k] == l) code[k++] = (uint16_t)(c++);
                if (c - 1 >= (1u << l)) return false;
            }
            maxcode[l] = c << (16 - l);
            c <<= 1;
        }
        maxcode[17] = 0xffffffffu;
        return true;
    }
    // symbol index + length of the code at the top of a 9-bit window, or -1 when the code is longer
    int peek9(uint32_t window9, int& len) const {
        const uint32_t top16 = window9 << 7;
        for (int l = 1; l <= 9; ++l)
            if (top16 < maxcode[l]) { len = l; const int sym = (int)(top16 >> (16 - l)) + delta[l]; return (sym >= 0 && sym < 256 && size[sym] == l) ? sym : -1; }
        return -1;
    }
    void build_fast_ac() {
        for (uint32_t i = 0; i < 512; ++i) {
            fast_ac[i] = 0;
            int len = 0;
            const int sym = peek9(i, len);
            if (sym < 0) continue;
            const int rs = values[sym], run = rs >> 4, mag = rs & 15;
            if (mag && len + mag <= 9) {
                int k = (int)(((i << len) & 511u) >> (9 - mag));
                if (k < (1 << (mag - 1))) k -= (1 << mag) - 1;
                if (k >= -128 && k <= 127) fast_ac[i] = (int16_t)(k * 256 + run * 16 + len + mag);
            }
        }
    }
};

struct JpegComp {
    int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dc_pred = 0;
    int x = 0, y = 0, w2 = 0, h2 = 0, coeff_w = 0;
    std::vector<uint8_t> data;     // w2 x h2 samples
    std::vector<int16_t> coeff;    // progressive: 64 per block
};

struct JpegDec {
    const uint8_t* p; size_t n, pos = 0;
    uint32_t buf = 0; int bits = 0; int marker = 0xff; bool nomore = false;   // entropy reader; marker 0xff = none
    JpegHuff hdc[4], hac[4];
    uint16_t dequant[4][64];
    JpegComp comp[4];
    int ncomp = 0, W = 0, H = 0, hmax = 1, vmax = 1, mcu_x = 0, mcu_y = 0;
    bool progressive = false, jfif = false;
    int adobe_transform = -1, rgb_ids = 0;
    int scan_n = 0, order[4] = {0, 0, 0, 0}, spec_start = 0, spec_end = 0, succ_high = 0, succ_low = 0, eob_run = 0;
    int restart_interval = 0, todo = 0;
    std::string why;

    int get8() { return pos < n ? p[pos++] : 0; }
    int get16() { const int a = get8(); return (a << 8) | get8(); }
    bool fail(const char* m) { if (why.empty()) why = m; return false; }

    void grow() {
        do {
            const uint32_t b = nomore ? 0u : (uint32_t)get8();
            if (b == 0xff) {
                int c = get8();
                while (c == 0xff) c = get8();
                if (c != 0) { marker = c; nomore = true; return; }
            }
            buf |= b << (24 - bits);
            bits += 8;
        } while (bits <= 24);
    }
    int huff(const JpegHuff& h) {
        if (bits < 16) grow();
        const uint32_t top = buf >> 16;
        int k = 1;
        while (k <= 16 && top >= h.maxcode[k]) ++k;
        if (k == 17 || k > bits) return -1;
        const int sym = (int)(buf >> (32 - k)) + h.delta[k];
        if (sym < 0 || sym >= 256) return -1;
        bits -= k;
        buf <<= k;
        return h.values[sym];
    }
    int take(int k) {   // k unsigned bits, 0 once the stream has run dry
        if (bits < k) grow();
        if (bits < k || k == 0) return 0;
        const uint32_t v = buf >> (32 - k);
        buf <<= k;
        bits -= k;
        return (int)v;
    }
    int extend(int k) {  // T.81 F.2.2.1 RECEIVE + EXTEND
        if (bits < k) grow();
        if (bits < k) return 0;
        const int v = take(k);
        return v < (1 << (k - 1)) ? v - (1 << k) + 1 : v;
    }
    int get_marker() {
        if (marker != 0xff) { const int m = marker; marker = 0xff; return m; }
        int x = get8();
        if (x != 0xff) return 0xff;
        while (x == 0xff) x = get8();
        return x;
    }
    void reset() {
        bits = 0; buf = 0; nomore = false; marker = 0xff; eob_run = 0;
        for (JpegComp& c : comp) c.dc_pred = 0;
        todo = restart_interval ? restart_interval : 0x7fffffff;
    }

    bool table_marker(int m) {   // DRI, DQT, DHT, APPn, COM
        static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                       35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        if (m == 0xff) return fail("corrupt JPEG (marker expected)");
        if (m == 0xdd) { if (get16() != 4) return fail("corrupt JPEG (DRI)"); restart_interval = get16(); return true; }
        if (m == 0xdb) {
            int L = get16() - 2;
            while (L > 0) {
                const int q = get8(), prec = q >> 4, t = q & 15;
                if (prec > 1 || t > 3) return fail("corrupt JPEG (DQT)");
                for (int i = 0; i < 64; ++i) dequant[t][zz[i]] = (uint16_t)(prec ? get16() : get8());
                L -= prec ? 129 : 65;
            }
            return L == 0 ? true : fail("corrupt JPEG (DQT)");
        }
        if (m == 0xc4) {
            int L = get16() - 2;
            while (L > 0) {
                const int q = get8(), tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) return fail("corrupt JPEG (DHT)");
                int count[16], total = 0;
                for (int& c : count) { c = get8(); total += c; }
                if (total > 256) return fail("corrupt JPEG (DHT)");
                JpegHuff& h = tc ? hac[th] : hdc[th];
                if (!h.build(count)) return fail("corrupt JPEG (DHT)");
                for (int i = 0; i < total; ++i) h.values[i] = (uint8_t)get8();
                if (tc) h.build_fast_ac();
                L -= 17 + total;
            }
            return L == 0 ? true : fail("corrupt JPEG (DHT)");
        }
        if ((m >= 0xe0 && m <= 0xef) || m == 0xfe) {
            int L = get16();
            if (L < 2) return fail("corrupt JPEG (APP / COM)");
            L -= 2;
            if (m == 0xe0 && L >= 5) {
                static const char tag[5] = {'J', 'F', 'I', 'F', 0};
                bool ok = true;
                for (char t : tag) if (get8() != (uint8_t)t) ok = false;
                L -= 5;
                if (ok) jfif = true;
            } else if (m == 0xee && L >= 12) {
                static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
                bool ok = true;
                for (char t : tag) if (get8() != (uint8_t)t) ok = false;
                L -= 6;
                if (ok) { get8(); get16(); get16(); adobe_transform = get8(); L -= 6; }
            }
            pos = L < 0 ? n : (pos + (size_t)L > n ? n : pos + (size_t)L);
            return true;
        }
        return fail("unsupported JPEG (lossless, hierarchical or arithmetic-coded frame, or an unknown marker)");
    }

    bool frame_header() {
        const int Lf = get16();
        if (Lf < 11) return fail("corrupt JPEG (SOF)");
        if (get8() != 8) return fail("JPEG sample precision other than 8 bit");
        H = get16(); W = get16();
        if (H == 0) return fail("JPEG with delayed height is not supported");
        if (W == 0) return fail("corrupt JPEG (width 0)");
        ncomp = get8();
        if (ncomp != 1 && ncomp != 3 && ncomp != 4) return fail("corrupt JPEG (component count)");
        if (Lf != 8 + 3 * ncomp) return fail("corrupt JPEG (SOF)");
        rgb_ids = 0;
        for (int i = 0; i < ncomp; ++i) {
            JpegComp& c = comp[i];
            c.id = get8();
            if (ncomp == 3 && c.id == "RGB"[i]) ++rgb_ids;
            const int q = get8();
            c.h = q >> 4; c.v = q & 15; c.tq = get8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("corrupt JPEG (sampling factors)");
        }
        if ((uint64_t)W * (uint64_t)H * (uint64_t)ncomp > 0x7fffffffull) return fail("JPEG too large");
        hmax = vmax = 1;
        for (int i = 0; i < ncomp; ++i) { hmax = comp[i].h > hmax ? comp[i].h : hmax; vmax = comp[i].v > vmax ? comp[i].v : vmax; }
        for (int i = 0; i < ncomp; ++i) if (hmax % comp[i].h || vmax % comp[i].v) return fail("JPEG with fractional sampling ratios is not supported");
        mcu_x = (W + 8 * hmax - 1) / (8 * hmax);
        mcu_y = (H + 8 * vmax - 1) / (8 * vmax);
        for (int i = 0; i < ncomp; ++i) {
            JpegComp& c = comp[i];
            c.x = (W * c.h + hmax - 1) / hmax;
            c.y = (H * c.v + vmax - 1) / vmax;
            c.w2 = mcu_x * c.h * 8;
            c.h2 = mcu_y * c.v * 8;
            c.data.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) { c.coeff_w = c.w2 / 8; c.coeff.assign((size_t)c.w2 * c.h2, 0); }
        }
        return true;
    }

    bool scan_header() {
        const int Ls = get16();
        scan_n = get8();
        if (scan_n < 1 || scan_n > 4 || scan_n > ncomp) return fail("corrupt JPEG (SOS)");
        if (Ls != 6 + 2 * scan_n) return fail("corrupt JPEG (SOS)");
        for (int i = 0; i < scan_n; ++i) {
            const int id = get8(), q = get8();
            int which = 0;
            while (which < ncomp && comp[which].id != id) ++which;
            if (which == ncomp) return fail("corrupt JPEG (SOS component)");
            comp[which].hd = q >> 4; comp[which].ha = q & 15;
            if (comp[which].hd > 3 || comp[which].ha > 3) return fail("corrupt JPEG (SOS table)");
            order[i] = which;
        }
        spec_start = get8(); spec_end = get8();
        const int aa = get8();
        succ_high = aa >> 4; succ_low = aa & 15;
        if (progressive) {
            if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) return fail("corrupt JPEG (SOS)");
        } else {
            if (spec_start != 0 || succ_high != 0 || succ_low != 0) return fail("corrupt JPEG (SOS)");
            spec_end = 63;
        }
        return true;
    }

    static const uint8_t* zigzag() {   // 15 more entries so that a corrupt run cannot index past the block
        static const uint8_t zz[79] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                       35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                       63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};
        return zz;
    }

    bool block_sequential(int16_t d[64], JpegComp& c) {
        const uint8_t* zz = zigzag();
        const JpegHuff& ac = hac[c.ha];
        const uint16_t* dq = dequant[c.tq];
        if (bits < 16) grow();
        const int t = huff(hdc[c.hd]);
        if (t < 0 || t > 15) return fail("corrupt JPEG data");
        std::memset(d, 0, 64 * sizeof(int16_t));
        const int diff = t ? extend(t) : 0;
        c.dc_pred += diff;
        d[0] = (int16_t)(c.dc_pred * dq[0]);
        int k = 1;
        do {
            if (bits < 16) grow();
            const int r = ac.fast_ac[buf >> 23];
            if (r) {
                k += (r >> 4) & 15;
                const int s = r & 15;
                if (s > bits) return fail("corrupt JPEG data");
                buf <<= s; bits -= s;
                const int z = zz[k++];
                d[z] = (int16_t)((r >> 8) * dq[z]);
            } else {
                const int rs = huff(ac);
                if (rs < 0) return fail("corrupt JPEG data");
                const int s = rs & 15, run = rs >> 4;
                if (s == 0) { if (rs != 0xf0) break; k += 16; }
                else { k += run; const int z = zz[k++]; d[z] = (int16_t)(extend(s) * dq[z]); }
            }
        } while (k < 64);
        return true;
    }
    bool block_prog_dc(int16_t d[64], JpegComp& c) {
        if (spec_end != 0) return fail("corrupt JPEG (DC scan with AC band)");
        if (bits < 16) grow();
        if (succ_high == 0) {
            std::memset(d, 0, 64 * sizeof(int16_t));
            const int t = huff(hdc[c.hd]);
            if (t < 0 || t > 15) return fail("corrupt JPEG data");
            const int diff = t ? extend(t) : 0;
            c.dc_pred += diff;
            d[0] = (int16_t)(c.dc_pred * (1 << succ_low));
        } else if (take(1)) d[0] = (int16_t)(d[0] + (int16_t)(1 << succ_low));
        return true;
    }
    bool block_prog_ac(int16_t d[64], JpegComp& c) {
        const uint8_t* zz = zigzag();
        const JpegHuff& ac = hac[c.ha];
        if (spec_start == 0) return fail("corrupt JPEG (AC scan with DC)");
        if (succ_high == 0) {
            if (eob_run) { --eob_run; return true; }
            int k = spec_start;
            do {
                if (bits < 16) grow();
                const int r = ac.fast_ac[buf >> 23];
                if (r) {
                    k += (r >> 4) & 15;
                    const int s = r & 15;
                    if (s > bits) return fail("corrupt JPEG data");
                    buf <<= s; bits -= s;
                    d[zz[k++]] = (int16_t)((r >> 8) * (1 << succ_low));
                } else {
                    const int rs = huff(ac);
                    if (rs < 0) return fail("corrupt JPEG data");
                    const int s = rs & 15, run = rs >> 4;
                    if (s == 0) {
                        if (run < 15) { eob_run = 1 << run; if (run) eob_run += take(run); --eob_run; break; }
                        k += 16;
                    } else { k += run; d[zz[k++]] = (int16_t)(extend(s) * (1 << succ_low)); }
                }
            } while (k <= spec_end);
            return true;
        }
        const int16_t bit = (int16_t)(1 << succ_low);
        auto refine = [&](int16_t& v) { if (take(1) && (v & bit) == 0) v = (int16_t)(v > 0 ? v + bit : v - bit); };
        if (eob_run) {
            --eob_run;
            for (int k = spec_start; k <= spec_end; ++k) { int16_t& v = d[zz[k]]; if (v != 0) refine(v); }
            return true;
        }
        int k = spec_start;
        do {
            const int rs = huff(ac);
            if (rs < 0) return fail("corrupt JPEG data");
            int s = rs & 15, run = rs >> 4;
            if (s == 0) {
                if (run < 15) { eob_run = (1 << run) - 1; if (run) eob_run += take(run); run = 64; }
            } else {
                if (s != 1) return fail("corrupt JPEG data");
                s = take(1) ? bit : -bit;
            }
            while (k <= spec_end) {   // skip `run` zero coefficients, refining the non-zero ones on the way
                int16_t& v = d[zz[k++]];
                if (v != 0) refine(v);
                else { if (run == 0) { v = (int16_t)s; break; } --run; }
            }
        } while (k <= spec_end);
        return true;
    }

    // 32-bit arithmetic that wraps, as stb's int arithmetic does on the two's-complement machines it runs on: damaged files
    // reach coefficients whose products leave the int range (valid ones never do), and the bytes should still be stb's
    static int wadd(int a, int b) { return (int)((uint32_t)a + (uint32_t)b); }
    static int wsub(int a, int b) { return (int)((uint32_t)a - (uint32_t)b); }
    static int wmul(int a, int b) { return (int)((uint32_t)a * (uint32_t)b); }
    // 1-D pass of the integer inverse DCT on s[0..7]; returns the even part in x[] and the odd part in t[]
    static void idct_1d(const int s[8], int x[4], int t[4]) {
        int p2 = s[2], p3 = s[6];
        int p1 = wmul(wadd(p2, p3), 2217);
        const int t2 = wadd(p1, wmul(p3, -7567)), t3 = wadd(p1, wmul(p2, 3135));
        p2 = s[0]; p3 = s[4];
        const int t0 = wmul(wadd(p2, p3), 4096), t1 = wmul(wsub(p2, p3), 4096);
        x[0] = wadd(t0, t3); x[3] = wsub(t0, t3); x[1] = wadd(t1, t2); x[2] = wsub(t1, t2);
        int o0 = s[7], o1 = s[5], o2 = s[3], o3 = s[1];
        p3 = wadd(o0, o2);
        int p4 = wadd(o1, o3);
        p1 = wadd(o0, o3); p2 = wadd(o1, o2);
        const int p5 = wmul(wadd(p3, p4), 4816);
        o0 = wmul(o0, 1223); o1 = wmul(o1, 8410); o2 = wmul(o2, 12586); o3 = wmul(o3, 6149);
        p1 = wadd(p5, wmul(p1, -3685)); p2 = wadd(p5, wmul(p2, -10497));
        p3 = wmul(p3, -8034); p4 = wmul(p4, -1597);
        t[3] = wadd(wadd(o3, p1), p4); t[2] = wadd(wadd(o2, p2), p3); t[1] = wadd(wadd(o1, p2), p4); t[0] = wadd(wadd(o0, p1), p3);
    }
    static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
    static void idct_block(uint8_t* out, int stride, const int16_t d[64]) {
        int val[64];
        for (int i = 0; i < 8; ++i) {   // columns
            const int16_t* c = d + i;
            int* v = val + i;
            if (!(c[8] | c[16] | c[24] | c[32] | c[40] | c[48] | c[56])) {
                const int dc = c[0] * 4;
                for (int r = 0; r < 8; ++r) v[8 * r] = dc;
                continue;
            }
            const int s[8] = {c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]};
            int x[4], t[4];
            idct_1d(s, x, t);
            for (int q = 0; q < 4; ++q) x[q] = wadd(x[q], 512);
            v[0] = wadd(x[0], t[3]) >> 10; v[56] = wsub(x[0], t[3]) >> 10;
            v[8] = wadd(x[1], t[2]) >> 10; v[48] = wsub(x[1], t[2]) >> 10;
            v[16] = wadd(x[2], t[1]) >> 10; v[40] = wsub(x[2], t[1]) >> 10;
            v[24] = wadd(x[3], t[0]) >> 10; v[32] = wsub(x[3], t[0]) >> 10;
        }
        for (int i = 0; i < 8; ++i) {   // rows
            int x[4], t[4];
            idct_1d(val + 8 * i, x, t);
            for (int q = 0; q < 4; ++q) x[q] = wadd(x[q], 65536 + (128 << 17));
            uint8_t* o = out + (size_t)stride * i;
            o[0] = clamp8(wadd(x[0], t[3]) >> 17); o[7] = clamp8(wsub(x[0], t[3]) >> 17);
            o[1] = clamp8(wadd(x[1], t[2]) >> 17); o[6] = clamp8(wsub(x[1], t[2]) >> 17);
            o[2] = clamp8(wadd(x[2], t[1]) >> 17); o[5] = clamp8(wsub(x[2], t[1]) >> 17);
            o[3] = clamp8(wadd(x[3], t[0]) >> 17); o[4] = clamp8(wsub(x[3], t[0]) >> 17);
        }
    }

    // true = go on with the scan; false = the scan ends here (without being an error)
    bool mcu_done() {
        if (--todo > 0) return true;
        if (bits < 24) grow();
        if (!(marker >= 0xd0 && marker <= 0xd7)) return false;
        reset();
        return true;
    }
    bool entropy_scan() {
        reset();
        int16_t blk[64];
        if (scan_n == 1) {   // non-interleaved: the component's own blocks in raster order
            JpegComp& c = comp[order[0]];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int j = 0; j < bh; ++j)
                for (int i = 0; i < bw; ++i) {
                    if (!progressive) {
                        if (!block_sequential(blk, c)) return false;
                        idct_block(&c.data[(size_t)c.w2 * j * 8 + (size_t)i * 8], c.w2, blk);
                    } else {
                        int16_t* d = &c.coeff[64 * ((size_t)i + (size_t)j * c.coeff_w)];
                        if (!(spec_start == 0 ? block_prog_dc(d, c) : block_prog_ac(d, c))) return false;
                    }
                    if (!mcu_done()) return true;
                }
            return true;
        }
        for (int j = 0; j < mcu_y; ++j)
            for (int i = 0; i < mcu_x; ++i) {
                for (int k = 0; k < scan_n; ++k) {
                    JpegComp& c = comp[order[k]];
                    for (int y = 0; y < c.v; ++y)
                        for (int x = 0; x < c.h; ++x) {
                            const int bx = i * c.h + x, by = j * c.v + y;
                            if (!progressive) {
                                if (!block_sequential(blk, c)) return false;
                                idct_block(&c.data[(size_t)c.w2 * by * 8 + (size_t)bx * 8], c.w2, blk);
                            } else if (!block_prog_dc(&c.coeff[64 * ((size_t)bx + (size_t)by * c.coeff_w)], c)) return false;
                        }
                }
                if (!mcu_done()) return true;
            }
        return true;
    }
    void finish_progressive() {
        for (int ci = 0; ci < ncomp; ++ci) {
            JpegComp& c = comp[ci];
            const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int j = 0; j < bh; ++j)
                for (int i = 0; i < bw; ++i) {
                    int16_t* d = &c.coeff[64 * ((size_t)i + (size_t)j * c.coeff_w)];
                    for (int q = 0; q < 64; ++q) d[q] = (int16_t)(d[q] * dequant[c.tq][q]);
                    idct_block(&c.data[(size_t)c.w2 * j * 8 + (size_t)i * 8], c.w2, d);
                }
        }
    }

    bool decode_planes() {
        if (get_marker() != 0xd8) return fail("not a JPEG");
        int m = get_marker();
        while (!(m == 0xc0 || m == 0xc1 || m == 0xc2)) {
            if (!table_marker(m)) return false;
            m = get_marker();
            while (m == 0xff) {   // padding between segments
                if (pos >= n) return fail("JPEG without a frame header");
                m = get_marker();
            }
        }
        progressive = m == 0xc2;
        if (!frame_header()) return false;
        m = get_marker();
        while (m != 0xd9) {
            if (m == 0xda) {
                if (!scan_header() || !entropy_scan()) return false;
                if (marker == 0xff) {   // skip whatever follows the scan up to the next thing that looks like a marker
                    while (pos < n) {
                        int x = get8();
                        while (x == 0xff) {
                            if (pos >= n) break;
                            x = get8();
                            if (x != 0x00 && x != 0xff) { marker = x; break; }
                        }
                        if (marker != 0xff) break;
                    }
                }
                m = get_marker();
                if (m >= 0xd0 && m <= 0xd7) m = get_marker();
            } else if (m == 0xdc) {
                const int Ld = get16(), NL = get16();
                if (Ld != 4 || NL != H) return fail("corrupt JPEG (DNL)");
                m = get_marker();
            } else {
                if (!table_marker(m)) { why.clear(); break; }   // whatever was decoded so far is the image
                m = get_marker();
            }
        }
        if (progressive) finish_progressive();
        return true;
    }
};

inline uint8_t jpeg_mul255(uint8_t a, uint8_t b) { const unsigned t = (unsigned)a * b + 128u; return (uint8_t)((t + (t >> 8)) >> 8); }

inline bool load_jpeg(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    if (f.size() < 4 || f[0] != 0xff || f[1] != 0xd8) return false;
    JpegDec z{f.data(), f.size()};
    std::memset(z.dequant, 0, sizeof(z.dequant));
    for (JpegHuff& h : z.hdc) { std::memset(&h, 0, sizeof(h)); }
    for (JpegHuff& h : z.hac) { std::memset(&h, 0, sizeof(h)); }
    if (!z.decode_planes()) { why = z.why.empty() ? "corrupt JPEG" : z.why; return false; }
    const int W = z.W, H = z.H, nc = z.ncomp;
    const bool is_rgb = nc == 3 && (z.rgb_ids == 3 || (z.adobe_transform == 0 && !z.jfif));
    // per component: the two source rows an output row is blended from, advanced as in a streaming decoder
    struct Up { int hs, vs, ystep, w_lores, ypos; const uint8_t* line0; const uint8_t* line1; std::vector<uint8_t> buf; };
    Up up[4];
    for (int k = 0; k < nc; ++k) {
        const JpegComp& c = z.comp[k];
        up[k].hs = z.hmax / c.h; up[k].vs = z.vmax / c.v;
        up[k].ystep = up[k].vs >> 1;
        up[k].w_lores = (W + up[k].hs - 1) / up[k].hs;
        up[k].ypos = 0;
        up[k].line0 = up[k].line1 = c.data.data();
        up[k].buf.assign((size_t)W + 3 + 8, 0);
    }
    auto div4 = [](int v) { return (uint8_t)(v >> 2); };
    auto div16 = [](int v) { return (uint8_t)(v >> 4); };
    std::vector<uint8_t> top((size_t)4 * W * H);
    const uint8_t* row[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int j = 0; j < H; ++j) {
        for (int k = 0; k < nc; ++k) {
            Up& r = up[k];
            const bool bot = r.ystep >= (r.vs >> 1);
            const uint8_t* nr = bot ? r.line1 : r.line0;   // the nearer source row
            const uint8_t* fr = bot ? r.line0 : r.line1;
            uint8_t* o = r.buf.data();
            const int w = r.w_lores;
            if (r.hs == 1 && r.vs == 1) row[k] = nr;
            else if (r.hs == 1 && r.vs == 2) { for (int i = 0; i < w; ++i) o[i] = div4(3 * nr[i] + fr[i] + 2); row[k] = o; }
            else if (r.hs == 2 && r.vs == 1) {
                if (w == 1) o[0] = o[1] = nr[0];
                else {
                    o[0] = nr[0];
                    o[1] = div4(nr[0] * 3 + nr[1] + 2);
                    int i = 1;
                    for (; i < w - 1; ++i) { const int c3 = 3 * nr[i] + 2; o[2 * i] = div4(c3 + nr[i - 1]); o[2 * i + 1] = div4(c3 + nr[i + 1]); }
                    o[2 * i] = div4(nr[w - 2] * 3 + nr[w - 1] + 2);
                    o[2 * i + 1] = nr[w - 1];
                }
                row[k] = o;
            } else if (r.hs == 2 && r.vs == 2) {
                if (w == 1) o[0] = o[1] = div4(3 * nr[0] + fr[0] + 2);
                else {
                    int t1 = 3 * nr[0] + fr[0];
                    o[0] = div4(t1 + 2);
                    for (int i = 1; i < w; ++i) {
                        const int t0 = t1;
                        t1 = 3 * nr[i] + fr[i];
                        o[2 * i - 1] = div16(3 * t0 + t1 + 8);
                        o[2 * i] = div16(3 * t1 + t0 + 8);
                    }
                    o[2 * w - 1] = div4(t1 + 2);
                }
                row[k] = o;
            } else {
                if (r.buf.size() < (size_t)w * r.hs) r.buf.resize((size_t)w * r.hs), o = r.buf.data();
                for (int i = 0; i < w; ++i) for (int q = 0; q < r.hs; ++q) o[i * r.hs + q] = nr[i];
                row[k] = o;
            }
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < z.comp[k].y) r.line1 += z.comp[k].w2;
            }
        }
        uint8_t* out = &top[(size_t)4 * W * j];
        auto ycc = [&](uint8_t* d, int Y, int cb, int cr) {
            const int yf = (Y << 20) + (1 << 19);
            cr -= 128; cb -= 128;
            int r = yf + cr * 1470208;
            int g = yf + cr * -748800 + (int)((uint32_t)(cb * -360960) & 0xffff0000u);
            int b = yf + cb * 1858048;
            r >>= 20; g >>= 20; b >>= 20;
            d[0] = JpegDec::clamp8(r); d[1] = JpegDec::clamp8(g); d[2] = JpegDec::clamp8(b); d[3] = 255;
        };
        for (int i = 0; i < W; ++i) {
            uint8_t* d = out + 4 * i;
            if (nc == 1) { d[0] = d[1] = d[2] = row[0][i]; d[3] = 255; }
            else if (nc == 3) {
                if (is_rgb) { d[0] = row[0][i]; d[1] = row[1][i]; d[2] = row[2][i]; d[3] = 255; }
                else ycc(d, row[0][i], row[1][i], row[2][i]);
            } else if (z.adobe_transform == 0) {   // CMYK
                const uint8_t k = row[3][i];
                d[0] = jpeg_mul255(row[0][i], k); d[1] = jpeg_mul255(row[1][i], k); d[2] = jpeg_mul255(row[2][i], k); d[3] = 255;
            } else if (z.adobe_transform == 2) {   // YCCK
                ycc(d, row[0][i], row[1][i], row[2][i]);
                const uint8_t k = row[3][i];
                d[0] = jpeg_mul255((uint8_t)(255 - d[0]), k); d[1] = jpeg_mul255((uint8_t)(255 - d[1]), k); d[2] = jpeg_mul255((uint8_t)(255 - d[2]), k);
            } else ycc(d, row[0][i], row[1][i], row[2][i]);   // four components without an Adobe tag: the fourth is ignored
        }
    }
    store_flipped(top, W, H, t);
    return true;
}

}  // namespace detail

// Decodes an image file by content, never by extension, trying the formats in stbi_load's order (stb_image.h:1131-1187):
// PNG, BMP, GIF, PSD, PIC, JPEG, PNM, Radiance HDR, and TGA — which has no magic number — last.  false + `why` if it cannot.
inline bool load_image(const std::string& path, Texture& t, std::string* why_out = nullptr) {
    std::vector<uint8_t> f;
    std::string why;
    bool ok = false;
    if (!detail::read_file(path, f)) why = "cannot read file";
    else if (f.size() >= 8 && f[0] == 0x89 && f[1] == 'P') ok = detail::load_png(f, t, why);
    else if (f.size() >= 2 && f[0] == 'B' && f[1] == 'M') ok = detail::load_bmp(f, t, why);
    else if (f.size() >= 6 && std::memcmp(f.data(), "GIF8", 4) == 0 && (f[4] == '7' || f[4] == '9') && f[5] == 'a') ok = detail::load_gif(f, t, why);
    else if (f.size() >= 4 && std::memcmp(f.data(), "8BPS", 4) == 0) ok = detail::load_psd(f, t, why);
    else if (detail::pic_plausible(f)) ok = detail::load_pic(f, t, why);
    else if (f.size() >= 3 && f[0] == 0xff && f[1] == 0xd8) ok = detail::load_jpeg(f, t, why);
    else if (f.size() >= 2 && f[0] == 'P' && (f[1] == '5' || f[1] == '6')) ok = detail::load_pnm(f, t, why);
    else if (detail::hdr_plausible(f)) ok = detail::load_hdr(f, t, why);
    else if (detail::tga_plausible(f)) ok = detail::load_tga(f, t, why);
    else why = "unknown image format";
    if (!ok && why.empty()) why = "not a valid image of its kind";
    if (why_out) *why_out = why;
    return ok;
}

}  // namespace rt3host
