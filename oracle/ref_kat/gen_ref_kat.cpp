// Generates tests/golden/ref_kat.json by compiling the REFERENCE'S OWN headers (where they lie
// under /root/reference, never copied) as host code and evaluating them on fixed inputs.
// Built by oracle/Makefile target `ref_kat` into oracle/_ref/ (git-ignored).  IEEE fp32, no fast-math.
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <src/shader/shader_common.h>  // Onb, powerHeuristic (+ random.h, helpers.h, vec_math.h)
#include <src/light.h>
#include <src/util/sampling.h>
#include <sutil/Camera.h>

static void p3(const char* k, float3 v, const char* end = ",") { printf("\"%s\": [%.9g, %.9g, %.9g]%s\n", k, v.x, v.y, v.z, end); }

int main() {
    printf("{\n");
    // ---- RNG: cuda/random.h:31-72
    printf("\"tea4_rnd\": [\n");
    unsigned int cases[][2] = {{0, 0}, {1, 0}, {0, 1}, {262143, 0}, {131328, 3}, {2073599, 63}, {8294399, 127}, {12345, 6789}};
    int nc = sizeof(cases) / sizeof(cases[0]);
    for (int i = 0; i < nc; i++) {
        unsigned int s = tea<4>(cases[i][0], cases[i][1]);
        unsigned int s0 = s;
        float a = rnd(s), b = rnd(s), c = rnd(s);
        printf("  {\"v0\": %u, \"v1\": %u, \"seed\": %u, \"rnd\": [%.9g, %.9g, %.9g], \"state\": %u}%s\n", cases[i][0], cases[i][1], s0, a, b, c, s, i + 1 < nc ? "," : "");
    }
    printf("],\n");
    // ---- cosine hemisphere: src/util/sampling.h:27-37
    printf("\"cosine\": [\n");
    float cu[][2] = {{0.25f, 0.5f}, {0.294449925f, 0.695515215f}, {0.0f, 0.0f}, {0.999998987f, 0.75f}, {0.5f, 0.125f}, {0.9f, 0.9f}, {0.01f, 0.3333f}};
    nc = sizeof(cu) / sizeof(cu[0]);
    for (int i = 0; i < nc; i++) {
        float3 w = rendertoy3o::SampleCosineHemisphere(make_float2(cu[i][0], cu[i][1]));
        float pdf = w.z / M_PI;
        printf("  {\"u\": [%.9g, %.9g], \"w\": [%.9g, %.9g, %.9g], \"pdf\": %.9g}%s\n", cu[i][0], cu[i][1], w.x, w.y, w.z, pdf, i + 1 < nc ? "," : "");
    }
    printf("],\n");
    // ---- Onb: src/shader/shader_common.h:15-48
    printf("\"onb\": [\n");
    float ns[][3] = {{0, 1, 0}, {0, 0, 1}, {1, 0, 0}, {0.57735026f, 0.57735026f, 0.57735026f}, {-0.6f, 0.0f, 0.8f}, {0.0f, -1.0f, 0.0f}};
    nc = sizeof(ns) / sizeof(ns[0]);
    for (int i = 0; i < nc; i++) {
        float3 n = make_float3(ns[i][0], ns[i][1], ns[i][2]);
        rendertoy3o::Onb onb(n);
        float3 w = make_float3(0.3f, 0.4f, 0.8660254f);
        onb.inverse_transform(w);
        printf("  {\"n\": [%.9g, %.9g, %.9g], \"T\": [%.9g, %.9g, %.9g], \"B\": [%.9g, %.9g, %.9g], \"p\": [%.9g, %.9g, %.9g]}%s\n", n.x, n.y, n.z,
               onb.m_tangent.x, onb.m_tangent.y, onb.m_tangent.z, onb.m_binormal.x, onb.m_binormal.y, onb.m_binormal.z, w.x, w.y, w.z, i + 1 < nc ? "," : "");
    }
    printf("],\n");
    printf("\"power_heuristic\": [[%.9g, %.9g, %.9g], [%.9g, %.9g, %.9g]],\n", 0.5f, 0.25f, rendertoy3o::powerHeuristic(0.5f, 0.25f), 22.26721f, 0.1591549f,
           rendertoy3o::powerHeuristic(22.26721f, 0.1591549f));
    // ---- Light: src/light.h:13-61
    {
        rendertoy3o::Light l(make_float3(17, 12, 4), make_float3(343, 548.7f, 227), make_float3(343, 548.7f, 332), make_float3(213, 548.7f, 332));
        printf("\"light\": {\"sizeof\": %zu, \"area\": %.9g, ", sizeof(l), l.m_area);
        p3("normal", l.m_normal);
        printf("\"samples\": [\n");
        float Ps[][3] = {{278, 0, 279.5f}, {100, 200, 300}, {278, 548.0f, 279.5f}, {343, 548.7f, 227}};
        unsigned int seeds[] = {0x5df5f2bfu, 0xc09848f2u, 12345u, 777u};
        for (int i = 0; i < 4; i++) {
            unsigned int s = seeds[i];
            float3 pos, em;
            float pdf;
            l.Sample(make_float3(Ps[i][0], Ps[i][1], Ps[i][2]), s, pos, em, pdf);
            printf("  {\"P\": [%.9g, %.9g, %.9g], \"seed\": %u, \"pos\": [%.9g, %.9g, %.9g], \"em\": [%.9g, %.9g, %.9g], \"pdf\": %.9g, \"state\": %u}%s\n", Ps[i][0],
                   Ps[i][1], Ps[i][2], seeds[i], pos.x, pos.y, pos.z, em.x, em.y, em.z, pdf, s, i < 3 ? "," : "");
        }
        printf("]},\n");
    }
    // ---- make_color: cuda/helpers.h:35-66
    printf("\"make_color\": [\n");
    float cs[][3] = {{0, 0.0031308f, 1}, {0.5f, 0.18f, 0.002f}, {2, -1, 0.999f}, {0.25f, 0.75f, 0.01f}, {0.0031307f, 0.0031309f, 0.9999f}};
    nc = sizeof(cs) / sizeof(cs[0]);
    for (int i = 0; i < nc; i++) {
        uchar4 c = make_color(make_float3(cs[i][0], cs[i][1], cs[i][2]));
        printf("  {\"c\": [%.9g, %.9g, %.9g], \"rgba\": [%u, %u, %u, %u]}%s\n", cs[i][0], cs[i][1], cs[i][2], c.x, c.y, c.z, c.w, i + 1 < nc ? "," : "");
    }
    printf("],\n");
    // ---- camera: sutil/Camera.cpp:34-45
    printf("\"camera\": [\n");
    float asp[] = {1.0f, 16.0f / 9.0f};
    for (int i = 0; i < 2; i++) {
        sutil::Camera cam(make_float3(5, 5, 5), make_float3(0, 1, 0), make_float3(0, 1, 0), 45.0f, asp[i]);
        float3 U, V, W;
        cam.UVWFrame(U, V, W);
        printf("  {\"eye\": [5,5,5], \"lookat\": [0,1,0], \"up\": [0,1,0], \"fovy\": 45, \"aspect\": %.9g, \"U\": [%.9g, %.9g, %.9g], \"V\": [%.9g, %.9g, %.9g], \"W\": [%.9g, %.9g, %.9g]}%s\n",
               asp[i], U.x, U.y, U.z, V.x, V.y, V.z, W.x, W.y, W.z, i < 1 ? "," : "");
    }
    printf("]\n}\n");
    return 0;
}
