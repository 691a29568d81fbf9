// image_loader.hpp — texture file ingest for the .obj/.mtl loader, written from scratch (the reference
// calls stbi_load(..., STBI_rgb_alpha), src/mesh.cpp:137; stb_image is not used here).
// Decodes to RGBA8, rows flipped so that v = 0 is the image bottom (mesh.cpp:151-159):
//   * PNG  — 1-16 bit, grey / grey+alpha / RGB / RGBA / palette (+ tRNS), plain or Adam7-interlaced; own inflate
//   * BMP  — uncompressed 24 / 32 bit, bottom-up or top-down
//   * TGA  — true-colour 24 / 32 bit and 8-bit grey, raw or RLE, either origin
//   * PPM / PGM — binary P6 / P5, maxval 255
//   * JPEG — baseline sequential (Huffman, 8 bit), grey or YCbCr with any sampling factors, restart intervals
// Progressive / arithmetic / CMYK JPEG are not supported (load_image returns false and says why).
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
    if (!chan || !(depth == 8 || depth == 16 || (depth < 8 && (ctype == 0 || ctype == 3))) || (ctype == 3 && depth == 16)) { why = "unsupported PNG colour type / bit depth"; return false; }
    if (idat.size() < 6) { why = "PNG without image data"; return false; }
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)w * chan * depth / 8 + 2) * h);
    if (!inflate_raw(idat.data() + 2, idat.size() - 2, raw)) { why = "corrupt PNG data stream"; return false; }  // 2-byte zlib header; Adler-32 not checked
    const size_t bpp = (size_t)((chan * depth + 7) / 8);
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
                if (ctype == 0) { const int g = sample(0); d[0] = d[1] = d[2] = (uint8_t)g; d[3] = 255; }
                else if (ctype == 2) { d[0] = (uint8_t)sample(0); d[1] = (uint8_t)sample(1); d[2] = (uint8_t)sample(2); d[3] = 255; }
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
inline bool load_bmp(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    if (f.size() < 54 || f[0] != 'B' || f[1] != 'M') return false;
    auto le32 = [&](size_t o) { return (uint32_t)f[o] | ((uint32_t)f[o + 1] << 8) | ((uint32_t)f[o + 2] << 16) | ((uint32_t)f[o + 3] << 24); };
    const uint32_t off = le32(10), hdr = le32(14);
    const int32_t w = (int32_t)le32(18), hs = (int32_t)le32(22);
    const int bpp = f[28] | (f[29] << 8);
    const uint32_t comp = le32(30);
    if (hdr < 40 || w <= 0 || hs == 0 || (bpp != 24 && bpp != 32) || (comp != 0 && !(comp == 3 && bpp == 32))) { why = "unsupported BMP (need uncompressed 24 / 32 bit)"; return false; }
    const int h = hs < 0 ? -hs : hs;
    const size_t stride = (((size_t)w * bpp / 8) + 3) & ~(size_t)3;
    if ((size_t)off + stride * h > f.size()) { why = "truncated BMP"; return false; }
    std::vector<uint8_t> top((size_t)4 * w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* row = &f[off + stride * (size_t)(hs < 0 ? y : h - 1 - y)];
        for (int x = 0; x < w; ++x) {
            const uint8_t* s = row + (size_t)x * bpp / 8;
            uint8_t* d = &top[4 * ((size_t)y * w + x)];
            d[0] = s[2]; d[1] = s[1]; d[2] = s[0]; d[3] = bpp == 32 ? s[3] : 255;
        }
    }
    store_flipped(top, w, h, t);
    return true;
}

// ------------------------------------------------------------------------------------ TGA
inline bool load_tga(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    if (f.size() < 18) return false;
    const int idlen = f[0], cmap = f[1], type = f[2], w = f[12] | (f[13] << 8), h = f[14] | (f[15] << 8), bpp = f[16], desc = f[17];
    const bool rle = type == 10 || type == 11, grey = type == 3 || type == 11;
    if (cmap != 0 || !(type == 2 || type == 3 || rle) || w <= 0 || h <= 0 || !((grey && bpp == 8) || (!grey && (bpp == 24 || bpp == 32)))) { why = "unsupported TGA (need true-colour 24 / 32 bit or 8-bit grey)"; return false; }
    const size_t px = (size_t)bpp / 8;
    std::vector<uint8_t> data((size_t)w * h * px);
    size_t pos = 18 + (size_t)idlen, o = 0;
    if (!rle) {
        if (pos + data.size() > f.size()) { why = "truncated TGA"; return false; }
        std::memcpy(data.data(), &f[pos], data.size());
    } else {
        while (o < data.size()) {
            if (pos >= f.size()) { why = "truncated TGA"; return false; }
            const int hd = f[pos++], cnt = (hd & 127) + 1;
            if (hd & 128) {
                if (pos + px > f.size()) { why = "truncated TGA"; return false; }
                for (int k = 0; k < cnt && o < data.size(); ++k, o += px) std::memcpy(&data[o], &f[pos], px);
                pos += px;
            } else {
                const size_t nb = (size_t)cnt * px;
                if (pos + nb > f.size() || o + nb > data.size()) { why = "truncated TGA"; return false; }
                std::memcpy(&data[o], &f[pos], nb);
                pos += nb; o += nb;
            }
        }
    }
    const bool top_origin = (desc & 0x20) != 0, right_origin = (desc & 0x10) != 0;
    std::vector<uint8_t> top((size_t)4 * w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* s = &data[((size_t)(top_origin ? y : h - 1 - y) * w + (size_t)(right_origin ? w - 1 - x : x)) * px];
            uint8_t* d = &top[4 * ((size_t)y * w + x)];
            if (grey) { d[0] = d[1] = d[2] = s[0]; d[3] = 255; }
            else { d[0] = s[2]; d[1] = s[1]; d[2] = s[0]; d[3] = bpp == 32 ? s[3] : 255; }
        }
    store_flipped(top, w, h, t);
    return true;
}

// ------------------------------------------------------------------------------------ PPM / PGM
inline bool load_pnm(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    if (f.size() < 7 || f[0] != 'P' || (f[1] != '6' && f[1] != '5')) return false;
    const int chan = f[1] == '6' ? 3 : 1;
    size_t pos = 2;
    auto next_int = [&]() {
        for (;;) {
            while (pos < f.size() && (f[pos] == ' ' || f[pos] == '\t' || f[pos] == '\r' || f[pos] == '\n')) ++pos;
            if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') ++pos; continue; }
            break;
        }
        int v = -1;
        while (pos < f.size() && f[pos] >= '0' && f[pos] <= '9') { v = (v < 0 ? 0 : v * 10) + (f[pos] - '0'); ++pos; }
        return v;
    };
    const int w = next_int(), h = next_int(), maxv = next_int();
    ++pos;  // the single whitespace byte after maxval
    if (w <= 0 || h <= 0 || maxv != 255) { why = "unsupported PNM (need binary P5 / P6 with maxval 255)"; return false; }
    if (pos + (size_t)w * h * chan > f.size()) { why = "truncated PNM"; return false; }
    std::vector<uint8_t> top((size_t)4 * w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const uint8_t* s = &f[pos + i * chan];
        uint8_t* d = &top[4 * i];
        d[0] = s[0]; d[1] = s[chan == 3 ? 1 : 0]; d[2] = s[chan == 3 ? 2 : 0]; d[3] = 255;
    }
    store_flipped(top, w, h, t);
    return true;
}

// ------------------------------------------------------------------------------------ JPEG (baseline sequential, ITU T.81)
// 8-bit, Huffman, one interleaved scan, 1 or 3 components (grey / YCbCr, any sampling factors up to 4, chroma
// replicated on upsampling), restart intervals.  Progressive, arithmetic-coded and 4-component files are refused.
struct JpegHuff {
    uint8_t vals[256];
    int mincode[17], maxcode[17], valptr[17];  // per code length 1..16 (T.81 F.2.2.3); maxcode = -1: no code of that length
    bool set = false;
    void build(const uint8_t counts[16], const uint8_t* v, int nv) {
        std::memcpy(vals, v, (size_t)nv);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += counts[l - 1];
            k += counts[l - 1];
            maxcode[l] = counts[l - 1] ? code - 1 : -1;
            code <<= 1;
        }
        set = true;
    }
};
struct JpegBits {
    const uint8_t* p; size_t n, pos; uint32_t acc = 0; int cnt = 0; bool bad = false;
    JpegBits(const uint8_t* d, size_t len, size_t at) : p(d), n(len), pos(at) {}
    int bit() {
        if (cnt == 0) {
            if (pos >= n) { bad = true; return 0; }
            uint8_t b = p[pos++];
            if (b == 0xff) {
                if (pos < n && p[pos] == 0x00) ++pos;  // stuffed zero
                else { bad = true; --pos; return 0; }  // a marker inside the data: the scan ended early
            }
            acc = b; cnt = 8;
        }
        --cnt;
        return (int)((acc >> cnt) & 1u);
    }
    int receive(int k) { int v = 0; while (k--) v = (v << 1) | bit(); return v; }
    int decode(const JpegHuff& h) {
        int code = 0;
        for (int l = 1; l <= 16; ++l) {
            code = (code << 1) | bit();
            if (bad) return -1;
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        bad = true;
        return -1;
    }
    bool restart(int expect) {  // byte-align, then RSTn
        cnt = 0;
        if (pos + 2 > n || p[pos] != 0xff || p[pos + 1] != (uint8_t)(0xd0 + (expect & 7))) return false;
        pos += 2;
        return true;
    }
};

inline bool load_jpeg(const std::vector<uint8_t>& f, Texture& t, std::string& why) {
    static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    if (f.size() < 4 || f[0] != 0xff || f[1] != 0xd8) return false;
    struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0, pw = 0, ph = 0; std::vector<uint8_t> plane; };
    uint16_t quant[4][64] = {};
    JpegHuff dc[4], ac[4];
    std::vector<Comp> comps;
    int W = 0, H = 0, ri = 0;
    size_t pos = 2;
    auto be16 = [&](size_t o) { return (int)((f[o] << 8) | f[o + 1]); };
    for (;;) {
        while (pos < f.size() && f[pos] != 0xff) ++pos;
        while (pos < f.size() && f[pos] == 0xff) ++pos;
        if (pos >= f.size()) { why = "JPEG without image data"; return false; }
        const int m = f[pos++];
        if (m == 0xd9) { why = "JPEG without image data"; return false; }
        if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
        if (pos + 2 > f.size()) { why = "truncated JPEG"; return false; }
        const int len = be16(pos);
        if (len < 2 || pos + (size_t)len > f.size()) { why = "truncated JPEG"; return false; }
        const size_t seg = pos + 2, end = pos + (size_t)len;
        if (m == 0xc2 || m == 0xc6 || m == 0xca || m == 0xce) { why = "progressive JPEG is not supported (save it as baseline, or as PNG)"; return false; }
        if (m == 0xc3 || m == 0xc5 || m == 0xc7 || (m >= 0xc9 && m <= 0xcf && m != 0xcc)) { why = "lossless / arithmetic-coded JPEG is not supported"; return false; }
        if (m == 0xc0 || m == 0xc1) {
            if (len < 8 || f[seg] != 8) { why = "JPEG sample precision other than 8 bit"; return false; }
            H = be16(seg + 1); W = be16(seg + 3);
            const int nf = f[seg + 5];
            if (W <= 0 || H <= 0 || !(nf == 1 || nf == 3) || len < 8 + 3 * nf) { why = nf == 4 ? "4-component (CMYK) JPEG is not supported" : "bad JPEG frame header"; return false; }
            comps.assign((size_t)nf, Comp());
            for (int i = 0; i < nf; ++i) {
                Comp& c = comps[(size_t)i];
                c.id = f[seg + 6 + 3 * (size_t)i]; c.h = f[seg + 7 + 3 * (size_t)i] >> 4; c.v = f[seg + 7 + 3 * (size_t)i] & 15; c.tq = f[seg + 8 + 3 * (size_t)i] & 3;
                if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4) { why = "bad JPEG sampling factors"; return false; }
            }
        } else if (m == 0xdb) {
            size_t q = seg;
            while (q < end) {
                const int pq = f[q] >> 4, tq = f[q] & 15;
                ++q;
                if (tq > 3 || q + (size_t)(pq ? 128 : 64) > end) { why = "bad JPEG quantisation table"; return false; }
                for (int i = 0; i < 64; ++i) { quant[tq][zz[i]] = (uint16_t)(pq ? be16(q + 2 * (size_t)i) : f[q + (size_t)i]); }
                q += pq ? 128 : 64;
            }
        } else if (m == 0xc4) {
            size_t q = seg;
            while (q < end) {
                const int tc = f[q] >> 4, th = f[q] & 15;
                if (tc > 1 || th > 3 || q + 17 > end) { why = "bad JPEG Huffman table"; return false; }
                int nv = 0;
                for (int i = 0; i < 16; ++i) nv += f[q + 1 + (size_t)i];
                if (nv > 256 || q + 17 + (size_t)nv > end) { why = "bad JPEG Huffman table"; return false; }
                (tc ? ac[th] : dc[th]).build(&f[q + 1], &f[q + 17], nv);
                q += 17 + (size_t)nv;
            }
        } else if (m == 0xdd) {
            ri = be16(seg);
        } else if (m == 0xda) {
            const int ns = f[seg];
            if (comps.empty() || ns != (int)comps.size() || len < 6 + 2 * ns) { why = "JPEG with several scans is not supported"; return false; }
            for (int i = 0; i < ns; ++i) {
                const int cs = f[seg + 1 + 2 * (size_t)i], tdta = f[seg + 2 + 2 * (size_t)i];
                bool found = false;
                for (Comp& c : comps) if (c.id == cs) { c.td = tdta >> 4; c.ta = tdta & 15; found = true; }
                if (!found || (tdta >> 4) > 3 || (tdta & 15) > 3) { why = "bad JPEG scan header"; return false; }
            }
            pos = end;
            break;
        }
        pos = end;
    }
    int hmax = 1, vmax = 1;
    for (const Comp& c : comps) { hmax = c.h > hmax ? c.h : hmax; vmax = c.v > vmax ? c.v : vmax; if (!dc[c.td].set || !ac[c.ta].set) { why = "JPEG scan refers to a missing Huffman table"; return false; } }
    const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
    for (Comp& c : comps) { c.pw = mcux * c.h * 8; c.ph = mcuy * c.v * 8; c.plane.assign((size_t)c.pw * c.ph, 0); }
    float cosm[8][8];
    for (int x = 0; x < 8; ++x)
        for (int u = 0; u < 8; ++u) cosm[x][u] = (u == 0 ? 0.70710678118654752f : 1.0f) * std::cos((float)((2 * x + 1) * u) * 3.14159265358979323846f / 16.0f);
    JpegBits b(f.data(), f.size(), pos);
    int until_restart = ri, next_rst = 0;
    for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
            if (ri && until_restart == 0) {
                if (!b.restart(next_rst)) { why = "JPEG restart marker missing"; return false; }
                next_rst = (next_rst + 1) & 7; until_restart = ri;
                for (Comp& c : comps) c.pred = 0;
            }
            for (Comp& c : comps)
                for (int by = 0; by < c.v; ++by)
                    for (int bx = 0; bx < c.h; ++bx) {
                        float co[64] = {0};
                        const int s = b.decode(dc[c.td]);
                        if (s < 0 || s > 11) { why = "corrupt JPEG data"; return false; }
                        int diff = s ? b.receive(s) : 0;
                        if (s && diff < (1 << (s - 1))) diff -= (1 << s) - 1;  // EXTEND
                        c.pred += diff;
                        co[0] = (float)(c.pred * (int)quant[c.tq][0]);
                        for (int k = 1; k < 64;) {
                            const int rs = b.decode(ac[c.ta]);
                            if (rs < 0) { why = "corrupt JPEG data"; return false; }
                            const int r = rs >> 4, sz = rs & 15;
                            if (sz == 0) { if (r == 15) { k += 16; continue; } break; }  // ZRL / EOB
                            k += r;
                            if (k > 63) { why = "corrupt JPEG data"; return false; }
                            int v = b.receive(sz);
                            if (v < (1 << (sz - 1))) v -= (1 << sz) - 1;
                            co[zz[k]] = (float)(v * (int)quant[c.tq][zz[k]]);
                            ++k;
                        }
                        if (b.bad) { why = "truncated JPEG data"; return false; }
                        float tmp[64];
                        for (int y = 0; y < 8; ++y)       // rows: tmp[y][x] = sum_u cos[x][u] co[y][u]
                            for (int x = 0; x < 8; ++x) { float a = 0; for (int u = 0; u < 8; ++u) a += cosm[x][u] * co[8 * y + u]; tmp[8 * y + x] = a; }
                        const int ox = (mx * c.h + bx) * 8, oy = (my * c.v + by) * 8;
                        for (int x = 0; x < 8; ++x)       // columns
                            for (int y = 0; y < 8; ++y) {
                                float a = 0;
                                for (int v = 0; v < 8; ++v) a += cosm[y][v] * tmp[8 * v + x];
                                const long q = std::lround(a * 0.25f + 128.0f);
                                c.plane[(size_t)(oy + y) * c.pw + (size_t)(ox + x)] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
                            }
                    }
            if (ri) --until_restart;
        }
    std::vector<uint8_t> top((size_t)4 * W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t* d = &top[4 * ((size_t)y * W + x)];
            auto at = [&](const Comp& c) { return (float)c.plane[(size_t)(y * c.v / vmax) * c.pw + (size_t)(x * c.h / hmax)]; };
            if (comps.size() == 1) { d[0] = d[1] = d[2] = (uint8_t)at(comps[0]); }
            else {
                const float Y = at(comps[0]), cb = at(comps[1]) - 128.0f, cr = at(comps[2]) - 128.0f;
                const float rgb[3] = {Y + 1.402f * cr, Y - 0.344136f * cb - 0.714136f * cr, Y + 1.772f * cb};
                for (int k = 0; k < 3; ++k) { const long q = std::lround(rgb[k]); d[k] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q)); }
            }
            d[3] = 255;
        }
    store_flipped(top, W, H, t);
    return true;
}

}  // namespace detail

// Decodes an image file by content (PNG, BMP, PNM) or extension (TGA has no magic).  false + `why` if it cannot.
inline bool load_image(const std::string& path, Texture& t, std::string* why_out = nullptr) {
    std::vector<uint8_t> f;
    std::string why;
    bool ok = false;
    if (!detail::read_file(path, f)) why = "cannot read file";
    else if (f.size() >= 8 && f[0] == 0x89 && f[1] == 'P') ok = detail::load_png(f, t, why);
    else if (f.size() >= 2 && f[0] == 'B' && f[1] == 'M') ok = detail::load_bmp(f, t, why);
    else if (f.size() >= 2 && f[0] == 'P' && (f[1] == '5' || f[1] == '6')) ok = detail::load_pnm(f, t, why);
    else if (f.size() >= 3 && f[0] == 0xff && f[1] == 0xd8) ok = detail::load_jpeg(f, t, why);
    else {
        const size_t dot = path.rfind('.');
        std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
        for (char& c : ext) c = (char)std::tolower((unsigned char)c);
        if (ext == "tga") ok = detail::load_tga(f, t, why);
        else why = "unknown image format";
    }
    if (!ok && why.empty()) why = "not a valid image of its kind";
    if (why_out) *why_out = why;
    return ok;
}

}  // namespace rt3host
