// ref_sdk.cpp — TEST INFRASTRUCTURE.  Entry points of oracle/_ref/librt3ref.so that evaluate the OptiX-SDK device code
// the reference ships under cuda/ — the stage surface north_star names — where it lies:
//     cuda/sphere.cu:37-97            __intersection__sphere (compiled as its own object)
//     cuda/LocalGeometry.h:59-178     getLocalGeometry
//     cuda/LocalShading.h:37-54       sampleTexture
//     cuda/curve.h:38-443             Linear / Quadratic / CubicInterpolator, surfaceNormal<>, curveTangent
// through the functional OptiX stand-in (oracle/ref_shim/optix.h).  tests/test_reference_pins_sdk.py compares the oracle's
// restatements with them on random inputs and keeps the goldens tests/golden/ref_kat_sdk.npz.
#include <cuda/LocalGeometry.h>
#include <cuda/LocalShading.h>
#include <cuda/curve.h>
#include <cuda/whitted.h>

#include "../rt3o.h"

extern "C" void __intersection__sphere();

template <class C> static void curve_out(const C& bc, float u, float* ps, float out[16], float out_n[3]) {
    const float4 p = bc.position4(u), v = bc.velocity4(u), a = bc.acceleration4(u);
    const float3 t = curveTangent(bc, u);
    const float o[16] = {p.x, p.y, p.z, p.w, v.x, v.y, v.z, v.w, a.x, a.y, a.z, a.w, t.x, t.y, t.z, 0.0f};
    for (int k = 0; k < 16; ++k) out[k] = o[k];
    if (ps) {
        float3 q = make_float3(ps[0], ps[1], ps[2]);
        const float3 n = surfaceNormal(bc, u, q);
        out_n[0] = n.x; out_n[1] = n.y; out_n[2] = n.z;
        ps[0] = q.x; ps[1] = q.y; ps[2] = q.z;
    }
}

extern "C" {

// out = {reported (0/1), t, nx, ny, nz, radius attribute}
void rt3ref_sphere(const float o[3], const float d[3], float tmin, float tmax, const float center_radius[4], float out[6]) {
    whitted::HitGroupData hg{};
    GeometryData::Sphere s;
    s.center = make_float3(center_radius[0], center_radius[1], center_radius[2]);
    s.radius = center_radius[3];
    hg.geometry_data.setSphere(s);
    rt3shim::Lane& L = rt3shim::lane;
    L = rt3shim::Lane();
    L.sbt_override = &hg;
    L.ray_o = make_float3(o[0], o[1], o[2]); L.ray_d = make_float3(d[0], d[1], d[2]);
    L.ray_tmin = tmin; L.ray_tmax = tmax;
    __intersection__sphere();
    out[0] = L.reported ? 1.0f : 0.0f;
    out[1] = L.rep_t;
    for (int k = 0; k < 4; ++k) out[2 + k] = __uint_as_float(L.rep_attr[k]);
    L.sbt_override = nullptr;
}

// getLocalGeometry of one triangle hit.  indices: uint32 triples (or null: vertex soup); normals / uvs / colors may be null
// (the SDK's fallbacks, LocalGeometry.h:99-124,150-158).  out[27] = P N Ng UV dndu dndv dpdu dpdv color.
void rt3ref_local_geometry(const float* positions, const unsigned int* indices, const float* normals, const float* uvs, const float* colors,
                           unsigned int prim, float bary_u, float bary_v, const float obj2world[12], const float world2obj[12], float out[27]) {
    GeometryData g;
    GeometryData::TriangleMesh m{};
    m.positions.data = (CUdeviceptr)positions; m.positions.byte_stride = 0; m.positions.elmt_byte_size = 12; m.positions.count = 1u << 30;
    if (indices) { m.indices.data = (CUdeviceptr)indices; m.indices.elmt_byte_size = 4; m.indices.byte_stride = 0; m.indices.count = 1u << 30; }
    if (normals) { m.normals.data = (CUdeviceptr)normals; m.normals.elmt_byte_size = 12; m.normals.count = 1u << 30; }
    if (uvs) { m.texcoords[0].data = (CUdeviceptr)uvs; m.texcoords[0].elmt_byte_size = 8; m.texcoords[0].count = 1u << 30; }
    if (colors) { m.colors.data = (CUdeviceptr)colors; m.colors.elmt_byte_size = 16; m.colors.count = 1u << 30; }
    g.setTriangleMesh(m);
    rt3shim::Lane& L = rt3shim::lane;
    L = rt3shim::Lane();
    L.hit.hit = true; L.hit.prim = (int)prim; L.hit.u = bary_u; L.hit.v = bary_v;
    for (int k = 0; k < 12; ++k) { L.obj2world[k] = obj2world[k]; L.world2obj[k] = world2obj[k]; }
    const LocalGeometry lg = getLocalGeometry(g);
    const LocalGeometry::Texcoord& tc = lg.texcoord[0];
    const float v[27] = {lg.P.x, lg.P.y, lg.P.z, lg.N.x, lg.N.y, lg.N.z, lg.Ng.x, lg.Ng.y, lg.Ng.z, tc.UV.x, tc.UV.y, tc.dndu.x, tc.dndu.y, tc.dndu.z,
                         tc.dndv.x, tc.dndv.y, tc.dndv.z, tc.dpdu.x, tc.dpdu.y, tc.dpdu.z, tc.dpdv.x, tc.dpdv.y, tc.dpdv.z, lg.color.x, lg.color.y, lg.color.z, lg.color.w};
    for (int k = 0; k < 27; ++k) out[k] = v[k];
}

// sampleTexture<float4>: the texel the SDK fetches for UV under a texcoord transform; the fetch itself (tex2D) goes to the
// oracle's texture `tex` of `scene`.  out_uv = the transformed coordinates handed to tex2D.
static thread_local float g_last_uv[2];
static thread_local rt3o_scene* g_tex_scene;
static float4 sdk_tex_hook(void*, unsigned long long tex, float u, float v) {
    g_last_uv[0] = u; g_last_uv[1] = v;
    float rgb[3] = {0, 0, 0};
    if (g_tex_scene) rt3o_kat_fetch_texture(g_tex_scene, (int)tex - 1, u, v, rgb);
    return make_float4(rgb[0], rgb[1], rgb[2], 1.0f);
}
void rt3ref_sample_texture(rt3o_scene* scene, int tex, const float scale[2], const float rotation[2], const float offset[2], float u, float v,
                           float out_rgb[3], float out_uv[2]) {
    MaterialData::Texture t{};
    t.tex = (cudaTextureObject_t)(tex + 1);   // 0 means "no texture" to sampleTexture
    t.texcoord = 0;
    t.texcoord_scale = make_float2(scale[0], scale[1]);
    t.texcoord_rotation = make_float2(rotation[0], rotation[1]);
    t.texcoord_offset = make_float2(offset[0], offset[1]);
    LocalGeometry lg{};
    lg.texcoord[0].UV = make_float2(u, v);
    g_tex_scene = scene;
    auto saved = rt3shim::hooks.tex2d;
    rt3shim::hooks.tex2d = sdk_tex_hook;
    const float4 c = sampleTexture<float4>(t, lg);
    rt3shim::hooks.tex2d = saved;
    out_rgb[0] = c.x; out_rgb[1] = c.y; out_rgb[2] = c.z;
    out_uv[0] = g_last_uv[0]; out_uv[1] = g_last_uv[1];
}

// curve.h.  basis: 0 linear (2 control points), 1 quadratic B-spline (3), 2 cubic B-spline, 3 Catmull-Rom, 4 Bezier (4 each).
// out[16] = position4(u), velocity4(u), acceleration4(u), curveTangent(u) + 0;  if ps != null: out_n = surfaceNormal(bc, u, ps)
// (the bona fide normal, type 2; type 1 for linear) and ps is updated in place like the SDK does.
int rt3ref_curve(int basis, const float* cp /*[n][4]*/, float u, float* ps /*[3] in/out or null*/, float out[16], float out_n[3]) {
    const float4* q = reinterpret_cast<const float4*>(cp);
    if (basis == 0) { LinearInterpolator bc; bc.initialize(q); curve_out(bc, u, ps, out, out_n); }
    else if (basis == 1) { QuadraticInterpolator bc; bc.initializeFromBSpline(q); curve_out(bc, u, ps, out, out_n); }
    else if (basis == 2) { CubicInterpolator bc; bc.initializeFromBSpline(q); curve_out(bc, u, ps, out, out_n); }
    else if (basis == 3) { CubicInterpolator bc; bc.initializeFromCatrom(q); curve_out(bc, u, ps, out, out_n); }
    else if (basis == 4) { CubicInterpolator bc; bc.initializeFromBezier(q); curve_out(bc, u, ps, out, out_n); }
    else return -1;
    return 0;
}

}  // extern "C"
