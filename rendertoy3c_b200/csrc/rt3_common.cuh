// rt3_common.cuh — device-side arithmetic of the hot path (sm_100a).
//
// Every function is a per-thread body usable from a __global__ kernel.  The translation units
// are compiled with `-fmad=false` so that each fp32 operation is individually rounded unless an
// explicit fmaf() is written; that makes the results independent of compiler contraction and
// lets parity with the CPU oracle be bit-exact.
//
// RT3_EMULATE: tests/emul builds these same bodies with g++ as a kernel-logic simulator for the
// GPU-less CI box (test infrastructure; never compiled into librt3.so).
//
// Reference pointers (relative to the rendertoy3C tree):
//   RNG cuda/random.h:31-72 | vec ops sutil/vec_math.h:470-585 | cosine sample src/util/sampling.h:27-37
//   Onb src/shader/shader_common.h:15-48 | Light src/light.h:13-61 | sRGB cuda/helpers.h:35-66
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <vector_types.h>
#include <vector_functions.h>

#ifdef RT3_EMULATE
#define RT3_HD inline
#define RT3_RESTRICT
static inline uint32_t rt3_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float rt3_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int rt3_clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int rt3_clzll(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static inline int rt3_popc(uint32_t x) { return __builtin_popcount(x); }
template <class T> static inline T rt3_ldg(const T* p) { return *p; }
static inline uint32_t rt3_atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t rt3_atomic_max(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
static inline uint32_t rt3_atomic_min(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline uint32_t rt3_atomic_or(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline void rt3_atomic_add64(unsigned long long* p, unsigned long long v) { *p += v; }
static inline void rt3_threadfence() {}
#else
#define RT3_HD __device__ __forceinline__
#define RT3_RESTRICT __restrict__
RT3_HD uint32_t rt3_f2u(float f) { return __float_as_uint(f); }
RT3_HD float rt3_u2f(uint32_t u) { return __uint_as_float(u); }
RT3_HD int rt3_clz(uint32_t x) { return __clz((int)x); }
RT3_HD int rt3_clzll(uint64_t x) { return __clzll((long long)x); }
RT3_HD int rt3_popc(uint32_t x) { return __popc(x); }
template <class T> RT3_HD T rt3_ldg(const T* p) { return __ldg(p); }
RT3_HD uint32_t rt3_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
RT3_HD uint32_t rt3_atomic_max(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
RT3_HD uint32_t rt3_atomic_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
RT3_HD uint32_t rt3_atomic_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
RT3_HD void rt3_atomic_add64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
RT3_HD void rt3_threadfence() { __threadfence(); }
#endif

// Queue records are touched once per stage: stream them through L2 (evict-first) so that the BVH
// nodes and primitive records, which every ray re-reads, keep the L2 to themselves.
#ifndef RT3_STREAM_HINTS
#define RT3_STREAM_HINTS 1
#endif
#if defined(RT3_EMULATE) || !RT3_STREAM_HINTS
template <class T> RT3_HD T rt3_ldcs(const T* p) { return *p; }
template <class T> RT3_HD void rt3_stcs(T* p, T v) { *p = v; }
#else
template <class T> RT3_HD T rt3_ldcs(const T* p) { return __ldcs(p); }
template <class T> RT3_HD void rt3_stcs(T* p, T v) { __stcs(p, v); }
#endif

namespace rt3 {

// ------------------------------------------------------------------------------------ float3 ops
RT3_HD float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
RT3_HD float3 v3(float4 a) { return make_float3(a.x, a.y, a.z); }
RT3_HD float3 add(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT3_HD float3 sub(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT3_HD float3 neg(float3 a) { return v3(-a.x, -a.y, -a.z); }
RT3_HD float3 mul(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT3_HD float3 mul(float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
RT3_HD float3 divs(float3 a, float s) { const float inv = 1.0f / s; return mul(a, inv); }  // vec_math.h:500-504
RT3_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT3_HD float3 cross(float3 a, float3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RT3_HD float length(float3 a) { return sqrtf(dot(a, a)); }
RT3_HD float3 normalize(float3 a) { const float inv = 1.0f / sqrtf(dot(a, a)); return mul(a, inv); }  // vec_math.h:560-564
RT3_HD float3 faceforward(float3 n, float3 i, float3 nref) { return mul(n, copysignf(1.0f, dot(i, nref))); }  // vec_math.h:582-585
RT3_HD float comp(float3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

// ------------------------------------------------------------------------------------ RNG
RT3_HD uint32_t tea4(uint32_t val0, uint32_t val1) {  // cuda/random.h:31-46, N = 4
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 4; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
RT3_HD float rnd(uint32_t& s) {  // cuda/random.h:49-67
    s = 1664525u * s + 1013904223u;
    return (float)(s & 0x00FFFFFFu) / (float)0x01000000;
}

// ------------------------------------------------------------------------------------ sampling
// sin/cos(2*pi*u): explicit quadrant reduction + minimax polynomials instead of the reference's
// fast-math sinf/cosf (CMakeLists.txt:267), so CPU and GPU agree bit for bit.
RT3_HD void sincos_2pi(float u, float& s, float& c) {
    const float a = u * 4.0f;
    const float qf = floorf(a + 0.5f);
    const float f = a - qf;
    const float x = f * 1.57079632679489661923f;
    const float x2 = x * x;
    const float sp = ((-1.9515295891e-4f * x2 + 8.3321608736e-3f) * x2 - 1.6666654611e-1f) * x2 * x + x;
    const float cp = ((2.443315711809948e-5f * x2 - 1.388731625493765e-3f) * x2 + 4.166664568298827e-2f) * x2 * x2 - 0.5f * x2 + 1.0f;
    const int q = ((int)qf) & 3;
    s = (q == 0) ? sp : (q == 1) ? cp : (q == 2) ? -sp : -cp;
    c = (q == 0) ? cp : (q == 1) ? -sp : (q == 2) ? -cp : sp;
}
RT3_HD float3 sample_cosine_hemisphere(float u1, float u2) {  // src/util/sampling.h:27-37
    const float r = sqrtf(u1);
    float s, c;
    sincos_2pi(u2, s, c);
    float3 p;
    p.x = r * c;
    p.y = r * s;
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    return p;
}
// Onb (src/shader/shader_common.h:15-48): returns p.x*T + p.y*B + p.z*N
RT3_HD float3 onb_inverse_transform(float3 n, float3 p) {
    float3 b;
    if (fabsf(n.x) > fabsf(n.z)) { b.x = -n.y; b.y = n.x; b.z = 0.0f; }
    else { b.x = 0.0f; b.y = -n.z; b.z = n.y; }
    b = normalize(b);
    const float3 t = cross(b, n);
    return add(add(mul(t, p.x), mul(b, p.y)), mul(n, p.z));
}
RT3_HD float power_heuristic(float p1, float p2) {  // shader_common.h:136-145
    const float a = p1 * p1, b = p2 * p2;
    return a / (a + b);
}

// ------------------------------------------------------------------------------------ Light (68-byte AoS, src/light.h:13-22)
struct Light {
    int32_t type;
    float emission[3], v0[3], v1[3], v2[3], normal[3];
    float area;
};
RT3_HD float3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
// Light::Sample, src/light.h:32-60
RT3_HD void light_sample(const Light* RT3_RESTRICT l, float3 P, uint32_t& seed, float3& pos, float3& emission, float& pdf) {
    const float u = rnd(seed);
    const float v = rnd(seed);
    const float su0 = sqrtf(u);
    const float b0 = 1.0f - su0;
    const float b1 = v * su0;
    pos = add(add(mul(ld3(l->v0), b0), mul(ld3(l->v1), b1)), mul(ld3(l->v2), 1.0f - b0 - b1));
    const float3 dv = sub(pos, P);
    const float dist2 = dot(dv, dv);
    if (dist2 < 1e-5f) { emission = v3(0, 0, 0); pdf = 1.0f; return; }
    const float3 nd = normalize(dv);
    const float omega = fabsf(dot(nd, ld3(l->normal))) * l->area / dist2;
    if (omega < 1e-5f) { emission = v3(0, 0, 0); pdf = 1.0f; return; }
    emission = mul(ld3(l->emission), omega);
    pdf = 1.0f / omega;
}

// ------------------------------------------------------------------------------------ sRGB quantise (cuda/helpers.h:35-66)
RT3_HD float to_srgb1(float c) {
    const float invGamma = 1.0f / 2.4f;
    const float powed = powf(c, invGamma);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}
RT3_HD float clamp01(float x) { return x != x ? 0.0f : fmaxf(0.0f, fminf(x, 1.0f)); }  // NaN -> 0, as the saturating move nvcc emits for the reference
RT3_HD uint32_t quantize_u8(float x) {
    x = clamp01(x);
    const uint32_t v = (uint32_t)(x * 256.0f);
    return v < 255u ? v : 255u;
}
RT3_HD uchar4 make_color(float3 c) {
    uchar4 o;
    o.x = (unsigned char)quantize_u8(to_srgb1(clamp01(c.x)));
    o.y = (unsigned char)quantize_u8(to_srgb1(clamp01(c.y)));
    o.z = (unsigned char)quantize_u8(to_srgb1(clamp01(c.z)));
    o.w = 255;
    return o;
}

// ------------------------------------------------------------------------------------ affine 3x4 row-major
struct Affine { float m[12]; };

RT3_HD float3 xform_point(const Affine& a, float3 p) {
    return v3(a.m[0] * p.x + a.m[1] * p.y + a.m[2] * p.z + a.m[3], a.m[4] * p.x + a.m[5] * p.y + a.m[6] * p.z + a.m[7],
              a.m[8] * p.x + a.m[9] * p.y + a.m[10] * p.z + a.m[11]);
}
RT3_HD float3 xform_vector(const Affine& a, float3 v) {
    return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[4] * v.x + a.m[5] * v.y + a.m[6] * v.z,
              a.m[8] * v.x + a.m[9] * v.y + a.m[10] * v.z);
}
RT3_HD float3 xform_normal_by_inverse(const Affine& inv, float3 n) {  // (M^-1)^T n, cuda/LocalGeometry.h:110,119
    return v3(inv.m[0] * n.x + inv.m[4] * n.y + inv.m[8] * n.z, inv.m[1] * n.x + inv.m[5] * n.y + inv.m[9] * n.z,
              inv.m[2] * n.x + inv.m[6] * n.y + inv.m[10] * n.z);
}
RT3_HD Affine invert_affine(const Affine& a) {
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
// OptixMatrixMotionTransform (src/cuda/cuda_accel.h:38-73): element-wise lerp of bracketing keys, clamped time
RT3_HD Affine lerp_keys(const float* RT3_RESTRICT keys, int nkeys, float t0, float t1, float time) {
    Affine r;
    const float tc = fminf(fmaxf(time, t0), t1);
    const float f = (tc - t0) / (t1 - t0) * (float)(nkeys - 1);
    int i = (int)floorf(f);
    if (i > nkeys - 2) i = nkeys - 2;
    if (i < 0) i = 0;
    const float a = f - (float)i;
    const float b = 1.0f - a;
    const float* k0 = keys + 12 * i;
    const float* k1 = k0 + 12;
#pragma unroll
    for (int j = 0; j < 12; j++) r.m[j] = b * k0[j] + a * k1[j];
    return r;
}

}  // namespace rt3
