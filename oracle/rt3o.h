/* ORACLE — TEST INFRASTRUCTURE ONLY.  C API of the scalar CPU oracle (librt3o.so).
 * Mirrors include/rt3.h entry for entry so the same Python scene description can be
 * replayed into either library; the product (librt3.so) never links or loads this.
 * Struct layouts (rt3_ray, rt3_hit, rt3_render_settings, rt3_stats) are shared with
 * include/rt3.h on purpose: they are the data contract under test.
 */
#ifndef RT3O_H
#define RT3O_H
#include <stdint.h>
#include "../include/rt3.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt3o_scene rt3o_scene;

rt3o_scene* rt3o_scene_create(void);
void rt3o_scene_destroy(rt3o_scene*);
int rt3o_mesh_create(rt3o_scene*, const float* verts, int num_keys, int nv, const int32_t* idx, int nt,
                     const float* normals, const float* uvs);                 /* returns blas id or <0; normals / uvs may be NULL (SDK fallbacks) */
int rt3o_mesh_set_colors(rt3o_scene*, int blas, const float* rgba);          /* [nv][4] vertex colours, cuda/LocalGeometry.h:99-110 */
int rt3o_spheres_create(rt3o_scene*, const float* center_radius, int n);
int rt3o_curves_create(rt3o_scene*, int degree, const float* cp_radius, int ncp, const int32_t* seg_first_cp, int nseg);
int rt3o_texture_create(rt3o_scene*, const uint8_t* rgba8, int w, int h, int address_mode, int filter_mode);
int rt3o_accel_append_instance(rt3o_scene*, int blas, const float xform[12]); /* returns instance id */
int rt3o_accel_append_animated_instance(rt3o_scene*, int blas, const float* keys, int nkeys, float t_begin,
                                        float t_end, const float static_xform[12]);
int rt3o_accel_build(rt3o_scene*);
int rt3o_scene_set_hitgroup(rt3o_scene*, int instance_id, const float emission[3], const float diffuse[3], int tex_id);
int rt3o_scene_set_texture_transform(rt3o_scene*, int instance_id, const float scale[2], const float rotation[2], const float offset[2]);  /* cuda/LocalShading.h:37-54 */
int rt3o_scene_set_lights(rt3o_scene*, const void* lights68, int n);
/* accel: 0 = brute force over every instance x primitive, 1 = BVH2 (validated against 0) */
int rt3o_trace(rt3o_scene*, const rt3_ray* rays, int n, int any_hit, rt3_hit* hits, int accel, int nthreads);
int rt3o_get_local_geometry(rt3o_scene*, const rt3_ray* rays, const rt3_hit* hits, int n, rt3_local_geometry* out);  /* cuda/LocalGeometry.h:61-175 */
int rt3o_launch_subframe(rt3o_scene*, const rt3_render_settings*, int nthreads);
int rt3o_download_accum(rt3o_scene*, float* rgba);
int rt3o_download_frame(rt3o_scene*, uint8_t* rgba8);
int rt3o_get_stats(rt3o_scene*, rt3_stats*);
int rt3o_reset_stats(rt3o_scene*);
const char* rt3o_last_error(void);
/* 1 = sum radiance in the reference's single running chain across the samples of a launch (raygen.cu:27,58-59);
 * 0 (default) = per-sample sums added in sample order, the association the wavefront kernels use */
void rt3o_set_chain_sum(int on);
/* 1 (default): static transformed triangle-mesh instances are intersected in world space, like the product's flattened
 * merged BLAS ("flatten" option of rt3_set_option); 0: every non-identity instance transforms the ray.  Takes effect at the
 * next rt3o_accel_build. */
void rt3o_set_flatten(int on);
/* 1 = deviation D1 off: the cosine sample uses cosf / sinf(2 pi u) of the C library like the reference's text compiled
 * for the host (src/util/sampling.h:27-37); 0 (default) = the explicit polynomial the kernels share */
void rt3o_set_libm_sincos(int on);

/* known-answer hooks (each follows the reference line cited in rt3o_math.hpp) */
uint32_t rt3o_kat_tea4(uint32_t v0, uint32_t v1);
float rt3o_kat_rnd(uint32_t* seed);
void rt3o_kat_cosine_sample(float u1, float u2, float out_xyz_pdf[4]);
void rt3o_kat_onb(const float n[3], const float w[3], float out_t_b_p[9]);
void rt3o_kat_light_make(const float e[3], const float v0[3], const float v1[3], const float v2[3], void* light68);
void rt3o_kat_light_sample(const void* light68, const float P[3], uint32_t* seed, float out_pos_em_pdf[7]);
void rt3o_kat_make_color(const float c[3], uint8_t out[4]);
void rt3o_kat_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy, float aspect, float out_uvw[9]);
void rt3o_kat_sincos_2pi(float u, float out_sc[2]);
void rt3o_kat_curve_eval(int basis, const float* cp, float u, float out[16]);  /* cuda/curve.h: position4, velocity4, acceleration4, curveTangent of one segment */
void rt3o_kat_invert_affine(const float m[12], float out[12]);
int rt3o_kat_fetch_texture(rt3o_scene*, int tex, float u, float v, float out_rgb[3]);
int rt3o_kat_sample_texture(rt3o_scene*, int instance_id, float u, float v, float out_rgb[3]);  /* with the instance's texcoord transform */  /* the shade stage's tex2D restatement */
int rt3o_kat_hit_triangle(const float o[3], const float d[3], const float v[9], float tmin, float tmax, float out_tuv[3]);
int rt3o_kat_hit_sphere(const float o[3], const float d[3], const float cr[4], float tmin, float tmax, float* t);
int rt3o_kat_hit_curve(const float o[3], const float d[3], const float a[4], const float b[4], float tmin, float tmax, float out_tu[2]);

#ifdef __cplusplus
}
#endif
#endif
