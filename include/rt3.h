/* rt3.h — C ABI of librt3.so: the B200-native (sm_100a) wavefront path tracer that replaces
 * the OptiX engine underneath rendertoy3o's device-scene operators.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * rendertoy3C tree).  Conventions:
 *   - every call returns 0 on success, a negative rt3_status otherwise; rt3_last_error()
 *     returns the thread-local message (reference: exceptions from src/util/exception.h:11-96).
 *   - host pointers are borrowed for the duration of the call only; device memory never
 *     crosses the ABI except through the explicit download calls / rt3_accum_device_ptr.
 *   - one context per GPU, externally synchronised (one host thread per context).
 *   - there is NO CPU fallback: without a CUDA device rt3_context_create fails with
 *     RT3_ERR_NO_DEVICE.
 */
#ifndef RT3_H
#define RT3_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RT3_OK = 0,
    RT3_ERR_INVALID = -1,   /* bad argument */
    RT3_ERR_CUDA = -2,      /* CUDA runtime error (message has file:line) */
    RT3_ERR_NO_DEVICE = -3, /* no usable CUDA device: there is no CPU fallback */
    RT3_ERR_STATE = -4,     /* call out of order (e.g. launch before accel build) */
    RT3_ERR_UNSUPPORTED = -5,
    RT3_ERR_NCCL = -6
} rt3_status;

typedef struct rt3_context* rt3_context_t; /* replaces OptixContext + CUDAScene (src/cuda/optix_context.h:231-271, src/cuda/cuda_scene.h:124-183) */
typedef uint64_t rt3_handle_t;             /* replaces OptixTraversableHandle of a GAS (src/cuda/cuda_mesh.h:165) */

/* Batch ray / hit records of the rt3_trace test hook (no reference equivalent: optixTraverse
 * is device-only, src/shader/shader_common.h:74-88).  16-byte aligned on purpose. */
typedef struct { float o[3]; float tmin; float d[3]; float tmax; float time; float pad[3]; } rt3_ray;   /* 48 B */
typedef struct { float t, u, v; int32_t prim; int32_t inst; int32_t pad[3]; } rt3_hit;                  /* 32 B; prim = inst = -1 on miss */

/* Launch parameters: replaces RenderSettings (src/shader/shader_data.h:71-114).  The accum /
 * frame buffers and the light array live inside the context instead of being raw pointers. */
typedef struct {
    uint32_t width, height;        /* film_settings.width/height */
    uint32_t samples_per_launch;   /* film_settings.samples_per_launch (reference default 8, src/wavefront.cpp:55) */
    uint32_t subframe_index;       /* film_settings.subframe_index: seeds tea<4>(pixel, subframe) (src/shader/raygen.cu:25) */
    float eye[3], U[3], V[3], W[3];/* camera_settings (sutil/Camera.cpp:34-45) */
    int32_t max_depth;             /* extension: max extension rays per path; <=0 = unbounded like the reference (src/shader/raygen.cu:48) */
    int32_t mode;                  /* 0 = REFERENCE_FAITHFUL estimator (quirks Q2-Q8 reproduced); 1 = CORRECTED: unbiased Lambert + NEE + MIS; 2 = CORRECTED with lights chosen in proportion to luminance(emission) x area (the reference README's unchecked "power light sampler") */
    float miss_color[3];           /* __direct_callable__test returns (0.01,0.01,0.01) (src/shader/test.cu:5) */
    int32_t accum_mode;            /* 0 = running mean exactly as src/shader/raygen.cu:79-85; 1 = per-pixel SUM of subframe means (for the multi-GPU reduce) */
} rt3_render_settings;

typedef struct {
    uint64_t rays_primary, rays_bounce, rays_shadow; /* rays actually traversed, device-counted */
    uint64_t samples;                                /* pixel-samples completed */
    uint64_t kernel_launches;                        /* number of librt3 kernels launched so far */
    float ms_generate, ms_extend, ms_shade, ms_connect, ms_resolve; /* CUDA-event time of the LAST subframe, per stage (0 if timing disabled) */
    float ms_total;
    uint32_t max_stack_depth;                        /* traversal stack high-water mark; recorded by diagnostic (-DRT3_STATS) builds only, 0 otherwise */
    uint32_t error_flags;                            /* bit0: a traversal stack overflowed and dropped a subtree: rt3_trace and rt3_download_* fail with RT3_ERR_STATE while it is set.  bit1: a launch with max_depth <= 0 (unbounded, like the reference) reached the 1022 bounces the library has slots for with paths still alive; they were ended there (informative).  rt3_reset_stats clears both */
    uint32_t flattened_instances;                    /* transformed static mesh instances the last rt3_accel_build merged into the world-space BLAS ("flatten") */
    uint32_t traversal_passes;                       /* launches per ray batch: 1, or 2 when the merged BLAS and the remaining instances are traversed one after the other ("split") */
} rt3_stats;

/* texture enums: identical values to CUDATexture<T>::AddressMode / FilterMode
 * (src/cuda/cuda_texture.h:16-28).  NOTE the reference casts FilterMode::Linear (=0) to
 * cudaTextureFilterMode where 0 is POINT; filter_mode 0 therefore means point sampling. */
enum { RT3_ADDRESS_WRAP = 0, RT3_ADDRESS_CLAMP = 1, RT3_ADDRESS_MIRROR = 2, RT3_ADDRESS_BORDER = 3 };
enum { RT3_FILTER_REFERENCE_LINEAR_IS_POINT = 0, RT3_FILTER_REFERENCE_POINT_IS_BILINEAR = 1 };  /* the reference's FilterMode values, by what the hardware does with them */

/* ---- context -------------------------------------------------------------------------- */
int rt3_context_create(int device, rt3_context_t* out);   /* OptixContext() src/cuda/optix_context.h:231-243 */
void rt3_context_destroy(rt3_context_t ctx);              /* ~CUDAScene src/cuda/cuda_scene.h:161-169 */
int rt3_sync(rt3_context_t ctx);                          /* CUDA_SYNC_CHECK src/wavefront.cpp:221 */
const char* rt3_last_error(void);
int rt3_get_stream(rt3_context_t ctx, void** cuda_stream); /* the cudaStream_t every launch of this context goes to (reference: state.stream, src/wavefront.cpp:302) */
int rt3_get_stats(rt3_context_t ctx, rt3_stats* out);
int rt3_reset_stats(rt3_context_t ctx);
int rt3_get_debug_counters(rt3_context_t ctx, uint32_t out[16]); /* diagnostic builds (-DRT3_STATS): [2] wide nodes visited, [3] primitives tested, [4] rounds, [5] rays, [6..12] warp-round histogram, [12] = queue-full events, [13] = instance entries (rt3_traverse.cuh); reads and clears */
int rt3_set_option(rt3_context_t ctx, const char* key, int value); /* tuning switches: "timing", "overlap" (0 = serial schedule, 1 = shadow rays of a bounce on a second stream beside the next extension (default), 2 = additionally two half-frame chains), "persist_ctas_per_sm", "merge_identity" (1 = single-level fast path for identity instances; call before rt3_accel_build), "flatten" (1, default: static transformed instances of plain triangle meshes are merged into the same world-space BLAS, vertices transformed at build — ~100 B of device memory per instanced triangle instead of a ray transform per visit; hit records then carry world-space arithmetic; needs merge_identity; not applied when the merged BLAS would exceed 2^27 triangles), "pipeline" (1, default: consecutive rt3_launch_subframe calls overlap on the GPU — three sets of queue pools and stream pairs take turns, so the head of the next subframe fills the GPU while the last bounces of the previous ones drain; only the resolves, which update the film in subframe order, stay on the context's stream; 0 = one subframe at a time; results identical, device memory for the pools x 3), "packets" (1, default: the camera rays of a subframe — depth 0 — traverse the merged BLAS in packets of eight consecutive rays, the eight lanes sharing one traversal and one wide-node test per step, when that BLAS is at most four times the L2 cache; 2 = whatever its size; 0 = never; results identical), "split" (merged BLAS beside other instances: 1, default = two launches per batch when the merged BLAS holds >= 1024 triangles — the single-level kernel, then the general kernel on the rest seeded with its result; 0 = never, 2 = always; results identical) */

/* ---- geometry (BLAS) ------------------------------------------------------------------- */
/* CUDAMesh(ctx, mesh) src/cuda/cuda_mesh.h:33-155: uploads vertex/index/normal/uv arrays and
 * builds the BLAS (there: optixAccelBuild + compaction; here: GPU LBVH -> compressed BVH8).
 * verts [num_keys][nv][3]: num_keys > 1 = vertex-key (deformation) motion blur, keys spread evenly over ray
 * time [0,1] like the reference's motionOptions (cuda_mesh.h:82-88), per-vertex linear interpolation; idx [nt][3],
 * normals [nv][3], uvs [nv][2].  The reference's own closest-hit program needs both (Q11); either may be NULL for the SDK's
 * fallbacks (cuda/LocalGeometry.h:120-124,150-158): without normals N = the geometric normal, without texcoords UV = the
 * barycentrics, in the shade stage and in rt3_get_local_geometry. */
int rt3_mesh_create(rt3_context_t ctx, const float* verts, int num_keys, int nv, const int32_t* idx, int nt,
                    const float* normals, const float* uvs, rt3_handle_t* blas);
/* optional vertex colours rgba [nv][4] of a triangle mesh (GeometryData::TriangleMesh::colors, cuda/LocalGeometry.h:99-110):
 * interpolated into rt3_local_geometry::color; call before rt3_accel_build */
int rt3_mesh_set_colors(rt3_context_t ctx, rt3_handle_t blas, const float* rgba);
/* analytic spheres, center_radius [n][4] (cuda/GeometryData.h:83-87, test per cuda/sphere.cu:37-97) */
int rt3_spheres_create(rt3_context_t ctx, const float* center_radius, int n, rt3_handle_t* blas);
/* round curves.  cp_radius [ncp][4]; `degree` selects the curve type (the round OptixPrimitiveTypes the SDK's cuda/curve.h
 * evaluates): segment i uses control points seg_first_cp[i] .. + 1 (linear), + 2 (quadratic B-spline), + 3 (the cubic types).
 * Linear segments are intersected directly (cuda/curve.h:38-80, cuda/GeometryData.h:127-133).  The other types are converted
 * with the SDK's Quadratic / CubicInterpolator::initializeFrom{BSpline, Catrom, Bezier} (cuda/curve.h:98-243); rays are
 * intersected with K round linear pieces per segment — K adapts per segment (1..64) so that the pieces stay within 2 % of
 * the segment's radius of the true curve — and the shading normal is the SDK's surfaceNormal<> of the TRUE
 * curve at the hit's parameter (cuda/curve.h:311-379; flat end caps at u = 0 / 1).  Hits report the segment and u in [0,1]. */
enum { RT3_CURVE_LINEAR = 1, RT3_CURVE_QUADRATIC_BSPLINE = 2, RT3_CURVE_CUBIC_BSPLINE = 3, RT3_CURVE_CATMULLROM = 4, RT3_CURVE_BEZIER = 5 };
int rt3_curves_create(rt3_context_t ctx, int degree, const float* cp_radius, int ncp, const int32_t* seg_first_cp,
                      int nseg, rt3_handle_t* blas);
/* CUDATexture<uchar4>(w,h,data,address,filter) src/cuda/cuda_texture.h:46-75 */
int rt3_texture_create(rt3_context_t ctx, const uint8_t* rgba8, int w, int h, int address_mode, int filter_mode, int* tex_id);

/* ---- instances (TLAS) ------------------------------------------------------------------ */
/* CUDAAccel::append_instance src/cuda/cuda_accel.h:75-90 (instanceId = sbtOffset = index) */
int rt3_accel_append_instance(rt3_context_t ctx, rt3_handle_t blas, const float xform[12], int* instance_id);
/* CUDAAccel::append_animated_instance src/cuda/cuda_accel.h:38-73: keys [nkeys][12] row-major 3x4,
 * OptixMotionOptions{numKeys,timeBegin,timeEnd}, plus the static instance transform */
int rt3_accel_append_animated_instance(rt3_context_t ctx, rt3_handle_t blas, const float* keys, int nkeys,
                                       float t_begin, float t_end, const float static_xform[12], int* instance_id);
int rt3_accel_build(rt3_context_t ctx);                   /* CUDAAccel::build src/cuda/cuda_accel.h:92-150 */

/* ---- shading records / lights ----------------------------------------------------------- */
/* one HitGroupRecord per instance: CUDAScene::create_sbt src/cuda/cuda_scene.h:54-88, HitGroupData src/shader/shader_data.h:125-136 */
int rt3_scene_set_hitgroup(rt3_context_t ctx, int instance_id, const float emission[3], const float diffuse[3], int tex_id);
/* texcoord transform of the SDK's sampleTexture (cuda/LocalShading.h:37-54): UV' = R(UV * scale) + offset with
 * rotation = (sin, cos); optional, per instance; without it the texture is fetched at the interpolated UV as in
 * the reference's closest-hit program (closehit_radiance.cu:105) */
int rt3_scene_set_texture_transform(rt3_context_t ctx, int instance_id, const float scale[2], const float rotation[2], const float offset[2]);
/* buildLightSampler src/wavefront.cpp:257-275: array of 68-byte rendertoy3o::Light (src/light.h:13-22) */
int rt3_scene_set_lights(rt3_context_t ctx, const void* lights68, int n);
/* Light ctor src/light.h:24-30 (host helper, pure arithmetic) */
int rt3_light_make(const float emission[3], const float v0[3], const float v1[3], const float v2[3], void* light68_out);
/* sutil::Camera::UVWFrame sutil/Camera.cpp:34-45 (host helper, pure arithmetic) */
int rt3_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy_deg, float aspect,
                   float U[3], float V[3], float W[3]);

/* ---- the hot path ---------------------------------------------------------------------- */
/* launchSubframe src/wavefront.cpp:203-222 = update_cuda_params_async + optixLaunch(w,h,1):
 * renders samples_per_launch paths per pixel through generate/extend/shade/connect/resolve
 * and folds them into the context's float4 accumulation buffer.  Asynchronous on the
 * context's stream. */
int rt3_launch_subframe(rt3_context_t ctx, const rt3_render_settings* settings);
/* batch closest-hit (any_hit=0) / occlusion (any_hit=1; hits[i].prim>=0 means occluded) query
 * through the same traversal kernels, host buffers in and out.  Test hook + e2e bench leg. */
int rt3_trace(rt3_context_t ctx, const rt3_ray* rays, int n, int any_hit, rt3_hit* hits);
/* same, with rays/hits already resident in device memory of this context's GPU */
int rt3_trace_device(rt3_context_t ctx, const void* d_rays, int n, int any_hit, void* d_hits);

/* ---- results ---------------------------------------------------------------------------- */
/* The SDK's stage record (cuda/LocalGeometry.h:40-58, one texcoord set) for hits returned by rt3_trace: world-space P
 * (interpolated vertices, object -> world at the ray time), shading normal N and geometric normal Ng (world space,
 * unit length), UV, the object-space derivatives dndu/dndv/dpdu/dpdv as getLocalGeometry forms them
 * (LocalGeometry.h:126-160) and color = the interpolated vertex colours (1 without them).  Spheres and curves (empty in the SDK): P = o + t d, N = Ng = surface
 * normal, UV = (0,0) / (u,0), zero derivatives.  Misses give an all-zero record. */
typedef struct rt3_local_geometry {
    float P[3], N[3], Ng[3];
    float UV[2], dndu[3], dndv[3], dpdu[3], dpdv[3];
    float color[4];
} rt3_local_geometry;
int rt3_get_local_geometry(rt3_context_t ctx, const rt3_ray* rays, const rt3_hit* hits, int n, rt3_local_geometry* out);

int rt3_download_accum(rt3_context_t ctx, float* rgba);   /* float4 accum_buffer [h][w][4]; row 0 = image bottom (Q19) */
int rt3_download_frame(rt3_context_t ctx, uint8_t* rgba8);/* uchar4 frame_buffer, make_color cuda/helpers.h:57-66 */
/* the same copy, asynchronous (the reference displays from a mapped GL buffer and never waits for a host copy, src/gui/display.cpp):
 * returns at once; the frame of everything launched so far is copied to `rgba8_pinned` (page-locked host memory, valid until
 * rt3_sync) on a copy stream beside the next subframe; rt3_sync completes it and fails like rt3_download_frame on a device error */
int rt3_download_frame_async(rt3_context_t ctx, uint8_t* rgba8_pinned);
int rt3_accum_device_ptr(rt3_context_t ctx, void** d_ptr, uint64_t* n_floats); /* for an external collective (torch.distributed / NCCL) */
int rt3_clear_accum(rt3_context_t ctx);                  /* restart accumulation (reference: subframe_index = 0 on camera change, src/wavefront.cpp:193-201) */
/* after an external SUM reduce in accum_mode 1: accum = sum / total_subframes, refresh the u8 frame */
int rt3_finalize_accum(rt3_context_t ctx, uint32_t total_subframes);
/* single-process multi-GPU: NCCL sum of the accumulation buffers of n contexts (one per GPU), then
 * finalize on each.  No reference equivalent (the reference is single-GPU). */
int rt3_allreduce_accum(rt3_context_t* ctxs, int n, uint32_t total_subframes);

#ifdef __cplusplus
}
#endif
#endif
