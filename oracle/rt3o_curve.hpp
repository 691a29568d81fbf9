// ORACLE — TEST INFRASTRUCTURE ONLY (see rt3o_math.hpp header).
//
// Polynomial curve segments of the SDK's cuda/curve.h, restated: every basis is converted once to power-basis
// coefficients c[0] u^3 + c[1] u^2 + c[2] u + c[3] (lower degrees leave the leading coefficients zero, which makes the
// cubic Horner forms below evaluate to exactly the quadratic / linear ones), with the operation order of
//     LinearInterpolator::initialize             curve.h:43-47
//     QuadraticInterpolator::initializeFromBSpline  :102-110
//     CubicInterpolator::initializeFromBSpline / Catrom / Bezier   :176-186, :209-219, :233-243
// (sums left to right; float4 / float = multiplication by the rounded reciprocal, sutil/vec_math.h:735-739), the
// evaluators position4 / velocity4 / acceleration4 (:50-79, :124-156, :264-309; the cubic velocity nudges u = 0 / 1 to
// 1e-6 / 0.999999) and surfaceNormal<> (:311-379 bona fide normal with flat end caps for degree 2 / 3, :380-425 conic
// normal with round end caps for degree 1).  Pinned against the SDK header compiled where it lies:
// tests/test_reference_pins_sdk.py.
#pragma once
#include "rt3o_math.hpp"

namespace rt3o {

enum { CURVE_LINEAR = 1, CURVE_QUADRATIC_BSPLINE = 2, CURVE_CUBIC_BSPLINE = 3, CURVE_CATMULLROM = 4, CURVE_BEZIER = 5 };
static inline int curve_control_points(int basis) { return basis == CURVE_LINEAR ? 2 : (basis == CURVE_QUADRATIC_BSPLINE ? 3 : 4); }

struct f4 { float x, y, z, w; };
static inline f4 operator+(f4 a, f4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
static inline f4 operator-(f4 a, f4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
static inline f4 operator*(f4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
static inline f4 operator*(float s, f4 a) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
static inline f4 div_by(f4 a, float s) { const float inv = 1.0f / s; return a * inv; }
static inline f3 xyz(f4 a) { return {a.x, a.y, a.z}; }

struct CurvePoly {
    int basis = CURVE_LINEAR;
    f4 c[4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    bool cubic() const { return basis >= CURVE_CUBIC_BSPLINE; }
};

static inline CurvePoly curve_poly(int basis, const float* cp) {
    const f4* q = reinterpret_cast<const f4*>(cp);
    CurvePoly p;
    p.basis = basis;
    switch (basis) {
        case CURVE_LINEAR:
            p.c[3] = q[0];
            p.c[2] = q[1] - q[0];
            break;
        case CURVE_QUADRATIC_BSPLINE:
            p.c[1] = div_by(q[0] - 2.0f * q[1] + q[2], 2.0f);
            p.c[2] = div_by(-2.0f * q[0] + 2.0f * q[1], 2.0f);
            p.c[3] = div_by(q[0] + q[1], 2.0f);
            break;
        case CURVE_CUBIC_BSPLINE:
            p.c[0] = div_by(q[0] * -1.0f + q[1] * 3.0f + q[2] * -3.0f + q[3], 6.0f);
            p.c[1] = div_by(q[0] * 3.0f + q[1] * -6.0f + q[2] * 3.0f, 6.0f);
            p.c[2] = div_by(q[0] * -3.0f + q[2] * 3.0f, 6.0f);
            p.c[3] = div_by(q[0] * 1.0f + q[1] * 4.0f + q[2] * 1.0f, 6.0f);
            break;
        case CURVE_CATMULLROM:
            p.c[0] = div_by(-1.0f * q[0] + 3.0f * q[1] + -3.0f * q[2] + 1.0f * q[3], 2.0f);
            p.c[1] = div_by(2.0f * q[0] + -5.0f * q[1] + 4.0f * q[2] + -1.0f * q[3], 2.0f);
            p.c[2] = div_by(-1.0f * q[0] + 1.0f * q[2], 2.0f);
            p.c[3] = div_by(2.0f * q[1], 2.0f);
            break;
        default:  // CURVE_BEZIER
            p.c[0] = q[0] * -1.0f + q[1] * 3.0f + q[2] * -3.0f + q[3];
            p.c[1] = q[0] * 3.0f + q[1] * -6.0f + q[2] * 3.0f;
            p.c[2] = q[0] * -3.0f + q[1] * 3.0f;
            p.c[3] = q[0];
            break;
    }
    return p;
}

static inline f4 curve_position(const CurvePoly& p, float u) { return ((p.c[0] * u + p.c[1]) * u + p.c[2]) * u + p.c[3]; }
static inline f4 curve_velocity(const CurvePoly& p, float u) {
    if (p.cubic()) {  // "adjust u to avoid problems with triple knots"
        if (u == 0.0f) u = 0.000001f;
        if (u == 1.0f) u = 0.999999f;
    }
    return (3.0f * p.c[0] * u + 2.0f * p.c[1]) * u + p.c[2];
}
static inline f4 curve_acceleration(const CurvePoly& p, float u) { return 6.0f * p.c[0] * u + 2.0f * p.c[1]; }
static inline f3 curve_tangent(const CurvePoly& p, float u) { return normalize(xyz(curve_velocity(p, u))); }

// surface normal (NOT yet normalised: the SDK normalises once, at the end) at curve parameter u for a point ps near the offset
// surface; ps is moved onto the surface like the SDK does
static inline f3 curve_surface_normal_raw(const CurvePoly& bc, float u, f3& ps) {
    f3 n;
    const bool linear = bc.basis == CURVE_LINEAR;
    if (u == 0.0f) n = linear ? ps - xyz(bc.c[3]) : -xyz(curve_velocity(bc, 0.0f));
    else if (linear ? u >= 1.0f : u == 1.0f) n = linear ? ps - (xyz(bc.c[2]) + xyz(bc.c[3])) : xyz(curve_velocity(bc, 1.0f));
    else {
        const f4 p4 = curve_position(bc, u);
        const f3 p = xyz(p4);
        const float r = p4.w;
        const f4 d4 = curve_velocity(bc, u);
        const f3 d = xyz(d4);
        const float dr = d4.w;
        float dd = dot(d, d);
        f3 o1 = ps - p;
        o1 = o1 - (dot(o1, d) / dd) * d;
        o1 = o1 * (r / length(o1));
        ps = p + o1;
        if (!linear) dd -= dot(xyz(curve_acceleration(bc, u)), o1);
        n = dd * o1 - (dr * r) * d;
    }
    return n;
}
static inline f3 curve_surface_normal(const CurvePoly& bc, float u, f3& ps) { return normalize(curve_surface_normal_raw(bc, u, ps)); }

}  // namespace rt3o
