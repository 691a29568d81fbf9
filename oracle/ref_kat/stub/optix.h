/* Minimal stand-in for the OptiX SDK header, ONLY so that the reference's own OptiX-free
 * arithmetic headers (cuda/random.h, cuda/helpers.h, src/light.h, src/util/sampling.h,
 * src/shader/shader_common.h) can be compiled as host code by gen_ref_kat.cpp.
 * Nothing here computes anything; OptiX itself is absent from this image. */
#pragma once
#include <cstddef>
typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
typedef int OptixResult;
typedef int OptixPayloadTypeID;
enum { OPTIX_SUCCESS = 0, OPTIX_PAYLOAD_TYPE_ID_0 = 1, OPTIX_RAY_FLAG_NONE = 0, OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 4,
       OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1, OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_READ_WRITE = 3,
       OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_READ = 1, OPTIX_PAYLOAD_SEMANTICS_CH_READ_WRITE = 12,
       OPTIX_PAYLOAD_SEMANTICS_CH_WRITE = 8, OPTIX_PAYLOAD_SEMANTICS_MS_WRITE = 32 };
#define OPTIX_SBT_RECORD_ALIGNMENT 16
#define OPTIX_SBT_RECORD_HEADER_SIZE 32
#ifndef __align__
#define __align__(n) alignas(n)
#endif
inline const char* optixGetErrorName(OptixResult) { return "stub"; }
template <class... A> inline void optixTraverse(A&&...) {}
template <class... A> inline void optixReorder(A&&...) {}
template <class... A> inline void optixInvoke(A&&...) {}
inline bool optixHitObjectIsHit() { return false; }
inline unsigned int __float_as_uint(float f) { unsigned int u; __builtin_memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned int u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
