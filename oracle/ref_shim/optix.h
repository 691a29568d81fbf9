/* optix.h — TEST INFRASTRUCTURE.  A FUNCTIONAL host stand-in for the OptiX 8 device API, just wide enough that the
 * reference's own device programs
 *     src/shader/raygen.cu  closehit_radiance.cu  miss.cu  test.cu        (+ the SDK's cuda/sphere.cu, curve.h,
 *     LocalGeometry.h, LocalShading.h for the known-answer generator)
 * compile WHERE THEY LIE under /root/reference as ordinary C++ (g++ -x c++) and run on the CPU, one "launch index"
 * at a time.  Nothing of OptiX is here: traversal is delegated to whoever installed rt3shim::Hooks (the harness in
 * ref_shaders.cpp points it at the oracle's brute-force intersector; there is no reference source for traversal), and
 * tex2D to the same.  What this pins is the reference's SHADER arithmetic: seeds, draws, closures, NEE, path loop,
 * accumulation — see oracle/ref_shim/ref_shaders.cpp and tests/test_reference_pins.py.
 */
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <math.h>
#include <cmath>
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <vector_types.h>
#include <vector_functions.h>

/* the CUDA headers turn these into attributes g++ ignores; the programs are plain functions here.  __constant__
 * becomes `weak` so that the `params` every program file defines (raygen.cu:3-6, closehit_radiance.cu:5-8) is ONE
 * object after linking, like the single launch-parameter block of an OptiX pipeline. */
#undef __host__
#undef __device__
#undef __global__
#undef __forceinline__
#undef __inline__
#undef __constant__
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __inline__ inline
#define __constant__ __attribute__((weak))
#ifndef __align__
#define __align__(n) alignas(n)
#endif

/* Argument evaluation order.  Two expressions of the reference draw two random numbers as the ARGUMENTS of one call:
 *     make_float2(rnd(seed), rnd(seed))      src/shader/raygen.cu:30  (pixel jitter), cuda/random.h:71 (rnd2)
 * C++ leaves the order unspecified.  nvcc's device compiler evaluates left to right — `nvcc -arch=sm_100a -ptx` of rnd2
 * stores {first draw, second draw} to (.x, .y) (excerpt in DESIGN.md §2) — while g++ evaluates right to left, which would
 * swap every jitter and every cosine-sample pair.  A braced initialiser list is evaluated left to right by rule
 * ([dcl.init.list]/4), so make_float2 is routed through one here; nothing else of the reference's text is touched. */
#include <sutil/vec_math.h>
namespace rt3shim {
struct F2 {
    float2 v;
    template <class A, class B> F2(A x, B y) { v.x = (float)x; v.y = (float)y; }
    template <class A> F2(A a) : v(::make_float2(a)) {}
};
}
#define make_float2(...) (rt3shim::F2{__VA_ARGS__}.v)

typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
typedef int OptixResult;
enum OptixPayloadTypeID { OPTIX_PAYLOAD_TYPE_DEFAULT = 0, OPTIX_PAYLOAD_TYPE_ID_0 = 1 };
enum { OPTIX_SUCCESS = 0 };
enum { OPTIX_RAY_FLAG_NONE = 0, OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1, OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 4 };
enum { OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_READ = 1, OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_READ_WRITE = 3, OPTIX_PAYLOAD_SEMANTICS_CH_WRITE = 8,
       OPTIX_PAYLOAD_SEMANTICS_CH_READ_WRITE = 12, OPTIX_PAYLOAD_SEMANTICS_MS_WRITE = 32 };
#define OPTIX_SBT_RECORD_ALIGNMENT 16
#define OPTIX_SBT_RECORD_HEADER_SIZE 32
inline const char* optixGetErrorName(OptixResult) { return "shim"; }

inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }

namespace rt3shim {
struct Hit { bool hit = false; float t = 0, u = 0, v = 0; int prim = -1, inst = -1; };
struct Hooks {   /* installed by the harness */
    void* user = nullptr;
    /* closest hit (any_hit = 0) or occlusion (any_hit = 1) in (tmin, tmax) at a ray time */
    Hit (*trace)(void* user, float3 o, float3 d, float tmin, float tmax, float time, int any_hit) = nullptr;
    float4 (*tex2d)(void* user, unsigned long long tex, float u, float v) = nullptr;
    void* (*sbt_hit)(void* user, int inst) = nullptr;   /* HitGroupData of an instance (sbtOffset = instance index) */
    void* sbt_miss = nullptr;
};
struct Lane {     /* the state of one OptiX thread */
    uint3 launch_index = {0, 0, 0};
    uint3 launch_dims = {1, 1, 1};
    unsigned int payload[32] = {0};
    Hit hit;
    float3 ray_o = {0, 0, 0}, ray_d = {0, 0, 1};
    float ray_tmin = 0, ray_tmax = 0, ray_time = 0;
    /* custom-primitive programs (cuda/sphere.cu): what optixReportIntersection received */
    bool reported = false; float rep_t = 0; unsigned int rep_kind = 0, rep_attr[8] = {0}; int rep_nattr = 0;
    void* sbt_override = nullptr; unsigned int prim_index = 0;
    /* current instance transform, rows of the 3x4 object->world and world->object matrices (identity unless a harness sets them) */
    float obj2world[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, world2obj[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
};
extern Hooks hooks;
extern thread_local Lane lane;
template <class... P> inline void store_payload(P... p) { unsigned int v[] = {p...}; for (size_t i = 0; i < sizeof...(P); ++i) lane.payload[i] = v[i]; }
inline void load_payload(int) {}
template <class P0, class... P> inline void load_payload(int i, P0& p0, P&... p) { p0 = lane.payload[i]; load_payload(i + 1, p...); }
}  // namespace rt3shim

/* the two device programs optixInvoke can run, and the callable */
extern "C" void __closesthit__radiance();
extern "C" void __miss__radiance();
extern "C" float3 __direct_callable__test();

inline uint3 optixGetLaunchIndex() { return rt3shim::lane.launch_index; }
inline uint3 optixGetLaunchDimensions() { return rt3shim::lane.launch_dims; }

/* optixTraverse with payload (src/shader/shader_common.h:74-88): finds the hit object; the payload registers are
 * untouched until optixInvoke runs a program on them */
template <class... P>
inline void optixTraverse(OptixPayloadTypeID, OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float time, OptixVisibilityMask,
                          unsigned int flags, unsigned int, unsigned int, unsigned int, P&...) {
    rt3shim::Lane& L = rt3shim::lane;
    L.ray_o = o; L.ray_d = d; L.ray_tmin = tmin; L.ray_time = time;
    L.hit = rt3shim::hooks.trace(rt3shim::hooks.user, o, d, tmin, tmax, time, (flags & OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT) ? 1 : 0);
    L.ray_tmax = L.hit.hit ? L.hit.t : tmax;
}
/* without payload (shader_common.h:119-132) */
inline void optixTraverse(OptixTraversableHandle h, float3 o, float3 d, float tmin, float tmax, float time, OptixVisibilityMask m, unsigned int flags,
                          unsigned int a, unsigned int b, unsigned int c) {
    optixTraverse(OPTIX_PAYLOAD_TYPE_DEFAULT, h, o, d, tmin, tmax, time, m, flags, a, b, c);
}
inline void optixReorder() {}
inline bool optixHitObjectIsHit() { return rt3shim::lane.hit.hit; }
template <class... P>
inline void optixInvoke(OptixPayloadTypeID, P&... p) {
    rt3shim::store_payload(p...);
    if (rt3shim::lane.hit.hit) __closesthit__radiance(); else __miss__radiance();
    rt3shim::load_payload(0, p...);
}
inline void optixSetPayloadTypes(unsigned int) {}

#define RT3SHIM_PAYLOAD(n) \
    inline unsigned int optixGetPayload_##n() { return rt3shim::lane.payload[n]; } \
    inline void optixSetPayload_##n(unsigned int v) { rt3shim::lane.payload[n] = v; }
RT3SHIM_PAYLOAD(0) RT3SHIM_PAYLOAD(1) RT3SHIM_PAYLOAD(2) RT3SHIM_PAYLOAD(3) RT3SHIM_PAYLOAD(4) RT3SHIM_PAYLOAD(5) RT3SHIM_PAYLOAD(6) RT3SHIM_PAYLOAD(7)
RT3SHIM_PAYLOAD(8) RT3SHIM_PAYLOAD(9) RT3SHIM_PAYLOAD(10) RT3SHIM_PAYLOAD(11) RT3SHIM_PAYLOAD(12) RT3SHIM_PAYLOAD(13) RT3SHIM_PAYLOAD(14) RT3SHIM_PAYLOAD(15)
RT3SHIM_PAYLOAD(16) RT3SHIM_PAYLOAD(17) RT3SHIM_PAYLOAD(18) RT3SHIM_PAYLOAD(19) RT3SHIM_PAYLOAD(20) RT3SHIM_PAYLOAD(21) RT3SHIM_PAYLOAD(22) RT3SHIM_PAYLOAD(23)
#undef RT3SHIM_PAYLOAD

inline CUdeviceptr optixGetSbtDataPointer() {
    const rt3shim::Lane& L = rt3shim::lane;
    if (L.sbt_override) return (CUdeviceptr)L.sbt_override;
    return (CUdeviceptr)(L.hit.hit ? rt3shim::hooks.sbt_hit(rt3shim::hooks.user, L.hit.inst) : rt3shim::hooks.sbt_miss);
}
inline unsigned int optixGetPrimitiveIndex() { return rt3shim::lane.hit.hit ? (unsigned int)rt3shim::lane.hit.prim : rt3shim::lane.prim_index; }
inline unsigned int optixGetInstanceId() { return (unsigned int)rt3shim::lane.hit.inst; }
inline float3 optixGetWorldRayOrigin() { return rt3shim::lane.ray_o; }
inline float3 optixGetWorldRayDirection() { return rt3shim::lane.ray_d; }
inline float3 optixGetObjectRayOrigin() { return rt3shim::lane.ray_o; }      /* identity instances only (all the reference creates) */
inline float3 optixGetObjectRayDirection() { return rt3shim::lane.ray_d; }
inline float optixGetRayTmin() { return rt3shim::lane.ray_tmin; }
inline float optixGetRayTmax() { return rt3shim::lane.ray_tmax; }
inline float optixGetRayTime() { return rt3shim::lane.ray_time; }
inline float2 optixGetTriangleBarycentrics() { return make_float2(rt3shim::lane.hit.u, rt3shim::lane.hit.v); }
template <class R, class... A> inline R optixDirectCall(unsigned int, A...) { return __direct_callable__test(); }

/* object <-> world helpers used by cuda/LocalGeometry.h:95,110,119 — the arithmetic of the OptiX SDK's
 * optix_device_impl_transformations.h: a point goes through the rows of object->world, a normal through the COLUMNS of
 * world->object (inverse transpose), each a left-to-right sum */
inline float3 optixTransformPointFromObjectToWorldSpace(float3 p) {
    const float* m = rt3shim::lane.obj2world;
    return make_float3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
inline float3 optixTransformNormalFromObjectToWorldSpace(float3 n) {
    const float* w = rt3shim::lane.world2obj;
    return make_float3(w[0] * n.x + w[4] * n.y + w[8] * n.z, w[1] * n.x + w[5] * n.y + w[9] * n.z, w[2] * n.x + w[6] * n.y + w[10] * n.z);
}

/* custom-primitive intersection programs (cuda/sphere.cu:37-97) */
template <class... A>
inline bool optixReportIntersection(float t, unsigned int kind, A... attrs) {
    rt3shim::Lane& L = rt3shim::lane;
    if (!(t > L.ray_tmin && t < L.ray_tmax)) return false;   /* OptiX accepts a report only inside the current interval */
    const unsigned int v[] = {attrs..., 0u};
    L.reported = true; L.rep_t = t; L.rep_kind = kind; L.rep_nattr = (int)sizeof...(A);
    for (int i = 0; i < L.rep_nattr; ++i) L.rep_attr[i] = v[i];
    L.ray_tmax = t;
    return true;
}

/* tex2D<float4>(texture object, u, v) (closehit_radiance.cu:105,146) */
template <class T> inline T tex2D(cudaTextureObject_t tex, float u, float v);
template <> inline float4 tex2D<float4>(cudaTextureObject_t tex, float u, float v) { return rt3shim::hooks.tex2d(rt3shim::hooks.user, tex, u, v); }
