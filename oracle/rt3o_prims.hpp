// ORACLE — TEST INFRASTRUCTURE ONLY (see rt3o_math.hpp header).
//
// Ray/primitive tests.  The reference has NO source for these (they live inside closed-source
// OptiX, src/shader/shader_common.h:74-88,119-133), so this file DEFINES them ("parity
// unpinned" by the reference; pinned by brute force + analytic cases in tests/):
//   * triangle : watertight test after Woop, Benthin, Wald, "Watertight Ray/Triangle
//                Intersection", JCGT 2(1) 2013 — published algorithm restated;
//                barycentrics in OptiX convention (u -> v1, v -> v2).
//   * sphere   : restates cuda/sphere.cu:44-96 (normalised-direction quadratic with root refinement).
//   * curve    : round linear segment (convex hull of two spheres); entry hits only, like
//                OptiX curves; u = curve parameter (cuda/curve.h:382-425 consumes it).
// Hit interval is the open interval (tmin, tmax).
#pragma once
#include "rt3o_math.hpp"

namespace rt3o {

struct RayShear {  // per-ray (per-object-space-ray) precomputation for the watertight test
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

static inline RayShear make_shear(f3 d) {
    RayShear s;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    s.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    s.kx = s.kz + 1; if (s.kx == 3) s.kx = 0;
    s.ky = s.kx + 1; if (s.ky == 3) s.ky = 0;
    if (get(d, s.kz) < 0.0f) { int t = s.kx; s.kx = s.ky; s.ky = t; }
    // one division: the reciprocal of the dominant component (clamped away from zero like the slab test's 1/d), Sx and Sy
    // are products with it.  Every triangle of a ray sees the same constants, which is all watertightness needs.
    const float eps = 8.271806e-25f;  // 2^-80
    float dz = get(d, s.kz);
    if (!(fabsf(dz) > eps)) dz = copysignf(eps, dz);
    s.Sz = 1.0f / dz;
    s.Sx = get(d, s.kx) * s.Sz;
    s.Sy = get(d, s.ky) * s.Sz;
    return s;
}

// returns true and (t,u,v) if the ray hits in (tmin,tmax)
static inline bool hit_triangle(f3 o, const RayShear& s, f3 v0, f3 v1, f3 v2, float tmin, float tmax,
                                float& t_out, float& u_out, float& v_out) {
    const f3 A = v0 - o, B = v1 - o, C = v2 - o;
    const float Akz = get(A, s.kz), Bkz = get(B, s.kz), Ckz = get(C, s.kz);
    const float Ax = get(A, s.kx) - s.Sx * Akz;
    const float Ay = get(A, s.ky) - s.Sy * Akz;
    const float Bx = get(B, s.kx) - s.Sx * Bkz;
    const float By = get(B, s.ky) - s.Sy * Bkz;
    const float Cx = get(C, s.kx) - s.Sx * Ckz;
    const float Cy = get(C, s.ky) - s.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = s.Sz * Akz, Bz = s.Sz * Bkz, Cz = s.Sz * Ckz;
    const float T = U * Az + V * Bz + W * Cz;
    const float rcp = 1.0f / det;
    const float t = T * rcp;
    if (!(t > tmin && t < tmax)) return false;
    t_out = t;
    u_out = V * rcp;  // weight of v1
    v_out = W * rcp;  // weight of v2
    return true;
}

// cuda/sphere.cu:44-96.  cr = (center, radius).
static inline bool hit_sphere(f3 o, f3 d, f3 center, float radius, float tmin, float tmax, float& t_out) {
    const f3 O = o - center;
    const float l = 1.0f / length(d);
    const f3 D = d * l;
    float b = dot(O, D);
    float c = dot(O, O) - radius * radius;
    float disc = b * b - c;
    if (disc > 0.0f) {
        float sdisc = sqrtf(disc);
        float root1 = (-b - sdisc);
        float root11 = 0.0f;
        const bool do_refine = fabsf(root1) > (10.0f * radius);
        if (do_refine) {
            f3 O1 = O + root1 * D;
            b = dot(O1, D);
            c = dot(O1, O1) - radius * radius;
            disc = b * b - c;
            if (disc > 0.0f) {
                sdisc = sqrtf(disc);
                root11 = (-b - sdisc);
            }
        }
        float t = (root1 + root11) * l;
        if (t > tmin && t < tmax) { t_out = t; return true; }
        float root2 = (-b + sdisc) + (do_refine ? root1 : 0.0f);
        t = root2 * l;
        if (t > tmin && t < tmax) { t_out = t; return true; }
    }
    return false;
}

// Round linear curve segment: convex hull of spheres (pa,ra) and (pb,rb).  Entry hit only.
// The ray origin is first advanced to the point nearest pa (t0) to keep the quartic-in-
// coordinates coefficients well conditioned; all arithmetic is on the normalised direction
// and the returned t is rescaled by l = 1/|d| (as cuda/sphere.cu does).
static inline bool hit_curve_linear(f3 o, f3 d, f3 pa, float ra, f3 pb, float rb, float tmin, float tmax,
                                    float& t_out, float& u_out) {
    const float l = 1.0f / length(d);
    const f3 D = d * l;
    const float t0 = dot(pa - o, D);
    const f3 ro = o + t0 * D;
    const f3 ba = pb - pa;
    const f3 oa = ro - pa;
    const f3 ob = ro - pb;
    const float rr = ra - rb;
    const float m0 = dot(ba, ba);
    const float m1 = dot(ba, oa);
    const float m2 = dot(ba, D);
    const float m3 = dot(D, oa);
    const float m5 = dot(oa, oa);
    const float m6 = dot(ob, D);
    const float m7 = dot(ob, ob);
    const float d2 = m0 - rr * rr;
    bool found = false;
    float tn = 0.0f, un = 0.0f;
    if (d2 > 0.0f) {  // otherwise one end sphere contains the other: caps only
        const float k2 = d2 - m2 * m2;
        const float k1 = d2 * m3 - m1 * m2 + m2 * rr * ra;
        const float k0 = d2 * m5 - m1 * m1 + m1 * rr * ra * 2.0f - m0 * ra * ra;
        const float h = k1 * k1 - k0 * k2;
        if (h > 0.0f && k2 != 0.0f) {
            const float tb = (-sqrtf(h) - k1) / k2;
            const float y = m1 - ra * rr + tb * m2;
            if (y > 0.0f && y < d2) { found = true; tn = tb; un = y / d2; }
        }
    }
    if (!found) {
        const float h1 = m3 * m3 - m5 + ra * ra;
        const float h2 = m6 * m6 - m7 + rb * rb;
        float best = 3.0e38f;
        if (h1 > 0.0f) { best = -m3 - sqrtf(h1); un = 0.0f; found = true; }
        if (h2 > 0.0f) {
            float tc = -m6 - sqrtf(h2);
            if (tc < best) { best = tc; un = 1.0f; }
            found = true;
        }
        tn = best;
    }
    if (!found) return false;
    const float t = (t0 + tn) * l;
    if (!(t > tmin && t < tmax)) return false;
    t_out = t;
    u_out = un;
    return true;
}

}  // namespace rt3o
