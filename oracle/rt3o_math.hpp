// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.
//
// Scalar C++17 restatement of the arithmetic on rendertoy3o's hot path.
// Compile with -ffp-contract=off: every float op below is individually rounded
// (IEEE fp32); the CUDA kernels are compiled with -fmad=false and follow the same
// operation order, so results are meant to be bit-identical, not merely close.
//
// Reference pointers (all relative to /root/reference):
//   RNG           cuda/random.h:31-72
//   vector ops    sutil/vec_math.h:470-585  (normalize = v*(1/sqrt(dot)); a/s = a*(1/s))
//   cosine sample src/util/sampling.h:27-37, src/util/math.h:20-23
//   ONB           src/shader/shader_common.h:15-48
//   Light         src/light.h:13-61
//   sRGB          cuda/helpers.h:35-66
//   camera        sutil/Camera.cpp:34-45
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace rt3o {

struct f3 { float x, y, z; };
struct f2 { float x, y; };

static inline f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
static inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
static inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline f3 operator*(float s, f3 a) { return {a.x * s, a.y * s, a.z * s}; }
// sutil/vec_math.h:500-504: a / s is a * (1/s)
static inline f3 operator/(f3 a, float s) { float inv = 1.0f / s; return a * inv; }
static inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float length(f3 v) { return sqrtf(dot(v, v)); }
// sutil/vec_math.h:560-564
static inline f3 normalize(f3 v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
// sutil/vec_math.h:582-585
static inline f3 faceforward(f3 n, f3 i, f3 nref) { return n * copysignf(1.0f, dot(i, nref)); }
// clamp(NaN, 0, 1) = 0: nvcc compiles the reference's fmaxf(0, fminf(x, 1)) (cuda/helpers.h:50,59) to a
// saturating move, which maps NaN to +0, so NaN pixels of the reference come out black, not white
static inline float clampf(float x, float a, float b) { return x != x ? a : fmaxf(a, fminf(x, b)); }
static inline float get(f3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// ---------------------------------------------------------------- RNG (cuda/random.h)
static inline uint32_t tea4(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (int n = 0; n < 4; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
static inline uint32_t lcg(uint32_t& prev) {
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00FFFFFFu;
}
static inline float rnd(uint32_t& prev) { return (float)lcg(prev) / (float)0x01000000; }

// ---------------------------------------------------------------- sin/cos of 2*pi*u
// The reference calls cosf/sinf(2*pi*u2) under --use_fast_math (CMakeLists.txt:267), which
// is not reproducible across devices.  Both the oracle and the CUDA kernels instead use
// this explicit polynomial (quadrant reduction + cephes-style minimax on |x| <= pi/4),
// so the sampled directions are bit-identical on CPU and GPU.  Max abs error ~1.2e-7.
static inline void sincos_2pi(float u, float& s, float& c) {
    float a = u * 4.0f;
    float qf = floorf(a + 0.5f);
    float f = a - qf;                       // [-0.5, 0.5]
    float x = f * 1.57079632679489661923f;  // [-pi/4, pi/4]
    float x2 = x * x;
    float sp = ((-1.9515295891e-4f * x2 + 8.3321608736e-3f) * x2 - 1.6666654611e-1f) * x2 * x + x;
    float cp = ((2.443315711809948e-5f * x2 - 1.388731625493765e-3f) * x2 + 4.166664568298827e-2f) * x2 * x2
               - 0.5f * x2 + 1.0f;
    int q = ((int)qf) & 3;
    switch (q) {
        case 0: s = sp;  c = cp;  break;
        case 1: s = cp;  c = -sp; break;
        case 2: s = -sp; c = -cp; break;
        default: s = -cp; c = sp; break;
    }
}

// src/util/sampling.h:27-37 (with the sincos substitution above = deviation D1).  g_libm_sincos switches D1 OFF: phi =
// 2 pi u2 and the C library's cosf / sinf, i.e. the reference's own text compiled for the host — the mode in which the
// oracle is compared BIT FOR BIT with the reference's device programs run on the host shim (tests/test_reference_pins.py).
inline bool g_libm_sincos = false;
static inline f3 sample_cosine_hemisphere(float u1, float u2) {
    const float r = sqrtf(u1);
    float s, c;
    if (g_libm_sincos) {
        const float phi = 2.0f * 3.14159265358979323846f * u2;
        c = cosf(phi);
        s = sinf(phi);
    } else sincos_2pi(u2, s, c);
    f3 p;
    p.x = r * c;
    p.y = r * s;
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    return p;
}

// src/shader/shader_common.h:15-48
struct Onb {
    f3 t, b, n;
    explicit Onb(f3 normal) {
        n = normal;
        if (fabsf(n.x) > fabsf(n.z)) { b.x = -n.y; b.y = n.x; b.z = 0; }
        else { b.x = 0; b.y = -n.z; b.z = n.y; }
        b = normalize(b);
        t = cross(b, n);
    }
    f3 inverse_transform(f3 p) const { return p.x * t + p.y * b + p.z * n; }
};

// src/shader/shader_common.h:136-145
static inline float power_heuristic(float p1, float p2) {
    const float a = p1 * p1, b = p2 * p2;
    return a / (a + b);
}

// ---------------------------------------------------------------- Light (src/light.h), 68-byte AoS
struct Light {
    int32_t type;
    f3 emission, v0, v1, v2, normal;
    float area;
};
static_assert(sizeof(Light) == 68, "rendertoy Light layout is 68 bytes");

// selection weight of the power light sampler (mode 2): luminance with the Russian-roulette weights of raygen.cu:66, times area
static inline float light_power(f3 emission, float area) { return (emission.x * 0.30f + emission.y * 0.59f + emission.z * 0.11f) * area; }

static inline Light light_make(f3 emission, f3 v0, f3 v1, f3 v2) {  // light.h:24-30
    Light l;
    l.type = 0; l.emission = emission; l.v0 = v0; l.v1 = v1; l.v2 = v2;
    l.normal = cross(v1 - v0, v2 - v0);
    l.area = 0.5f * length(l.normal);
    l.normal = normalize(l.normal);
    return l;
}
static inline void light_sample(const Light& l, f3 P, uint32_t& seed, f3& pos, f3& emission, float& pdf) {  // light.h:32-60
    const float u = rnd(seed);
    const float v = rnd(seed);
    float su0 = sqrtf(u);
    float b0 = 1.0f - su0;
    float b1 = v * su0;
    pos = b0 * l.v0 + b1 * l.v1 + (1.0f - b0 - b1) * l.v2;
    f3 dv = pos - P;
    float dist2 = dot(dv, dv);
    if (dist2 < 1e-5f) { emission = {0, 0, 0}; pdf = 1.0f; return; }
    f3 nd = normalize(dv);
    float omega = fabsf(dot(nd, l.normal)) * l.area / dist2;
    if (omega < 1e-5f) { emission = {0, 0, 0}; pdf = 1.0f; return; }
    emission = l.emission * omega;
    pdf = 1.0f / omega;
}

// ---------------------------------------------------------------- sRGB quantise (cuda/helpers.h:35-66)
static inline float to_srgb1(float c) {
    float invGamma = 1.0f / 2.4f;
    float powed = powf(c, invGamma);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}
static inline uint8_t quantize_u8(float x) {
    x = clampf(x, 0.0f, 1.0f);
    uint32_t v = (uint32_t)(x * 256.0f);
    return (uint8_t)(v < 255u ? v : 255u);
}
static inline void make_color(f3 c, uint8_t out[4]) {
    out[0] = quantize_u8(to_srgb1(clampf(c.x, 0.0f, 1.0f)));
    out[1] = quantize_u8(to_srgb1(clampf(c.y, 0.0f, 1.0f)));
    out[2] = quantize_u8(to_srgb1(clampf(c.z, 0.0f, 1.0f)));
    out[3] = 255;
}

// ---------------------------------------------------------------- camera (sutil/Camera.cpp:34-45)
static inline void camera_uvw(f3 eye, f3 lookat, f3 up, float fovY, float aspect, f3& U, f3& V, f3& W) {
    W = lookat - eye;
    float wlen = length(W);
    U = normalize(cross(W, up));
    V = normalize(cross(U, W));
    float vlen = wlen * tanf(0.5f * fovY * 3.14159265358979323846f / 180.0f);
    V = V * vlen;
    float ulen = vlen * aspect;
    U = U * ulen;
}

// ---------------------------------------------------------------- affine 3x4 (row-major, OptixInstance::transform)
struct Affine { float m[12]; };

static inline f3 xform_point(const Affine& a, f3 p) {
    return {a.m[0] * p.x + a.m[1] * p.y + a.m[2] * p.z + a.m[3],
            a.m[4] * p.x + a.m[5] * p.y + a.m[6] * p.z + a.m[7],
            a.m[8] * p.x + a.m[9] * p.y + a.m[10] * p.z + a.m[11]};
}
static inline f3 xform_vector(const Affine& a, f3 v) {
    return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z,
            a.m[4] * v.x + a.m[5] * v.y + a.m[6] * v.z,
            a.m[8] * v.x + a.m[9] * v.y + a.m[10] * v.z};
}
// normal by inverse-transpose: pass the INVERSE matrix (cuda/LocalGeometry.h:110,119 semantics)
static inline f3 xform_normal_by_inverse(const Affine& inv, f3 n) {
    return {inv.m[0] * n.x + inv.m[4] * n.y + inv.m[8] * n.z,
            inv.m[1] * n.x + inv.m[5] * n.y + inv.m[9] * n.z,
            inv.m[2] * n.x + inv.m[6] * n.y + inv.m[10] * n.z};
}
// Inverse by cofactors; the operation order here is part of the spec shared with the kernels.
static inline Affine invert_affine(const Affine& a) {
    const float* m = a.m;
    const float c00 = m[5] * m[10] - m[6] * m[9];
    const float c01 = m[6] * m[8] - m[4] * m[10];
    const float c02 = m[4] * m[9] - m[5] * m[8];
    const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const float id = 1.0f / det;
    Affine r;
    r.m[0] = c00 * id;
    r.m[1] = (m[2] * m[9] - m[1] * m[10]) * id;
    r.m[2] = (m[1] * m[6] - m[2] * m[5]) * id;
    r.m[4] = c01 * id;
    r.m[5] = (m[0] * m[10] - m[2] * m[8]) * id;
    r.m[6] = (m[2] * m[4] - m[0] * m[6]) * id;
    r.m[8] = c02 * id;
    r.m[9] = (m[1] * m[8] - m[0] * m[9]) * id;
    r.m[10] = (m[0] * m[5] - m[1] * m[4]) * id;
    r.m[3] = -(r.m[0] * m[3] + r.m[1] * m[7] + r.m[2] * m[11]);
    r.m[7] = -(r.m[4] * m[3] + r.m[5] * m[7] + r.m[6] * m[11]);
    r.m[11] = -(r.m[8] * m[3] + r.m[9] * m[7] + r.m[10] * m[11]);
    return r;
}
// OptixMatrixMotionTransform semantics (src/cuda/cuda_accel.h:38-73): element-wise lerp of the
// bracketing keys, time clamped to [t0,t1] (OPTIX_MOTION_FLAG_NONE).
static inline Affine lerp_keys(const float* keys, int nkeys, float t0, float t1, float time) {
    Affine r;
    if (nkeys <= 1) { std::memcpy(r.m, keys, sizeof(r.m)); return r; }
    float tc = fminf(fmaxf(time, t0), t1);
    float f = (tc - t0) / (t1 - t0) * (float)(nkeys - 1);
    int i = (int)floorf(f);
    if (i > nkeys - 2) i = nkeys - 2;
    if (i < 0) i = 0;
    float a = f - (float)i;
    float b = 1.0f - a;
    const float* k0 = keys + 12 * i;
    const float* k1 = k0 + 12;
    for (int j = 0; j < 12; j++) r.m[j] = b * k0[j] + a * k1[j];
    return r;
}

}  // namespace rt3o
