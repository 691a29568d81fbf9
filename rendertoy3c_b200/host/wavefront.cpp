// wavefront.cpp — headless host driver with the shape of the reference's src/wavefront.cpp:
// loadOBJ -> Context -> CUDAScene (meshes, textures, instances, hit groups, light sampler) ->
// RenderSettings + camera -> render loop of launchSubframe -> image to disk (the reference
// displays through GL instead; saveImage in sutil/sutil.cpp:542-700 flips rows the same way).
//
//   wavefront --scene scene.obj [--key frame1.obj ...] [--width 768 --height 768 --spp 64 --spl 8 --max-depth 0]
//             [--eye x y z --lookat x y z --up x y z --fovy 45] [--gpus N] [--mode 0|1|2]
//             [--out out.ppm|.png|.exr] [--tonemap none|aces]
// Multi-GPU: scene replicated, GPU g renders subframes g, g+N, ...; one NCCL sum of the float4
// accumulation buffers (rt3_allreduce_accum).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>

#include "image_writer.hpp"
#include "obj_loader.hpp"

using namespace rt3host;

int main(int argc, char** argv) {
    std::string scene, out = "out.ppm", tonemap = "none";
    std::vector<std::string> key_files;  // further .obj files of the same topology = vertex key-frames (src/mesh.cpp:39)
    int width = 768, height = 768, spp = 64, spl = 8, max_depth = 0, gpus = 1, mode = 0;  // reference defaults (wavefront.cpp:55,300)
    float eye[3] = {5, 5, 5}, lookat[3] = {0, 1, 0}, up[3] = {0, 1, 0}, fovy = 45.0f;  // initCameraState, wavefront.cpp:238-243
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&]() -> const char* {   // the argument of option `a`
            if (i + 1 >= argc) { std::fprintf(stderr, "option %s needs a value\n", a.c_str()); std::exit(2); }
            return argv[++i];
        };
        auto f3 = [&](float* v) { for (int k = 0; k < 3; ++k) v[k] = (float)std::atof(value()); };
        if (a == "--scene") scene = value();
        else if (a == "--key") key_files.push_back(value());
        else if (a == "--out") out = value();
        else if (a == "--width") width = std::atoi(value());
        else if (a == "--height") height = std::atoi(value());
        else if (a == "--spp") spp = std::atoi(value());
        else if (a == "--spl") spl = std::atoi(value());
        else if (a == "--max-depth") max_depth = std::atoi(value());
        else if (a == "--gpus") gpus = std::atoi(value());
        else if (a == "--decode-image" && i + 2 < argc) {  // utility: decode a texture file as loadOBJ would, dump w, h + RGBA8 (bottom row first)
            Texture t;
            std::string why;
            if (!load_image(argv[i + 1], t, &why)) { std::fprintf(stderr, "%s: %s\n", argv[i + 1], why.c_str()); return 1; }
            FILE* f = std::fopen(argv[i + 2], "wb");
            if (!f) return 1;
            const int32_t wh[2] = {t.width, t.height};
            std::fwrite(wh, sizeof(wh), 1, f);
            std::fwrite(t.pixel.data(), 1, t.pixel.size(), f);
            std::fclose(f);
            return 0;
        }
        else if (a == "--mode") mode = std::atoi(value());
        else if (a == "--tonemap") tonemap = value();  // none | aces (the reference viewer's display curve; 8-bit outputs only)  // 0 reference-faithful, 1 corrected, 2 corrected + power light sampler
        else if (a == "--fovy") fovy = (float)std::atof(value());
        else if (a == "--eye") f3(eye);
        else if (a == "--lookat") f3(lookat);
        else if (a == "--up") f3(up);
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (scene.empty()) { std::fprintf(stderr, "usage: wavefront --scene file.obj [options]\n"); return 2; }
    if (width <= 0 || height <= 0 || spp <= 0 || spl <= 0 || gpus <= 0) { std::fprintf(stderr, "wavefront: width, height, spp, spl and gpus must be positive\n"); return 2; }
    const int subframes = (spp + spl - 1) / spl;
    if (gpus > subframes) gpus = subframes;   // a GPU without a subframe would have no film to reduce
    try {
        std::vector<Mesh> meshes;
        std::vector<Texture> textures;
        std::vector<std::string> paths = {scene};
        paths.insert(paths.end(), key_files.begin(), key_files.end());
        loadOBJ(paths, meshes, textures);
        std::vector<std::unique_ptr<Context>> ctx;
        std::vector<std::unique_ptr<CUDAScene>> scenes;
        for (int g = 0; g < gpus; ++g) {
            ctx.emplace_back(new Context(g));
            scenes.emplace_back(new CUDAScene(*ctx.back(), meshes, textures));
        }
        RenderSettings params(width, height, (unsigned)spl);
        params.max_depth = max_depth;
        params.mode = mode;
        params.accum_mode = gpus > 1 ? 1 : 0;
        for (int k = 0; k < 3; ++k) params.eye[k] = eye[k];
        RT3HOST_CHECK(rt3_camera_uvw(eye, lookat, up, fovy, (float)width / (float)height, params.U, params.V, params.W));  // handleCameraUpdate
        const auto t0 = std::chrono::steady_clock::now();
        // GPU g renders subframes g, g + N, ... from its own host thread (one thread per context): a launch with unbounded
        // depth waits for its queue counters every bounce, which must not hold up the other devices
        std::vector<std::string> failures((size_t)gpus);
        auto drive = [&](int g) {
            try {
                RenderSettings mine = params;
                for (int sf = g; sf < subframes; sf += gpus) {
                    mine.subframe_index = (uint32_t)sf;
                    RT3HOST_CHECK(rt3_launch_subframe(ctx[(size_t)g]->ctx(), &mine));
                }
                RT3HOST_CHECK(rt3_sync(ctx[(size_t)g]->ctx()));
            } catch (const std::exception& e) { failures[(size_t)g] = e.what(); }
        };
        std::vector<std::thread> workers;
        for (int g = 1; g < gpus; ++g) workers.emplace_back(drive, g);
        drive(0);
        for (auto& w : workers) w.join();
        for (const std::string& f : failures) if (!f.empty()) throw Exception(f);
        if (gpus > 1) {
            std::vector<rt3_context_t> raw;
            for (auto& c : ctx) raw.push_back(c->ctx());
            RT3HOST_CHECK(rt3_allreduce_accum(raw.data(), gpus, (uint32_t)subframes));
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        unsigned long long rays = 0, samples = 0;
        for (auto& c : ctx) { rt3_stats st; RT3HOST_CHECK(rt3_get_stats(c->ctx(), &st)); rays += st.rays_primary + st.rays_bounce + st.rays_shadow; samples += st.samples; }
        std::vector<uint8_t> frame((size_t)4 * width * height);
        RT3HOST_CHECK(rt3_download_frame(ctx[0]->ctx(), frame.data()));
        if (tonemap == "aces") tonemap_frame_aces(frame);
        else if (tonemap != "none") throw Exception("unknown --tonemap " + tonemap);
        std::vector<float> accum;
        if (out.size() >= 4 && (out.substr(out.size() - 4) == ".exr" || out.substr(out.size() - 4) == ".EXR")) {
            accum.resize((size_t)4 * width * height);
            RT3HOST_CHECK(rt3_download_accum(ctx[0]->ctx(), accum.data()));
        }
        saveImage(out, width, height, frame.data(), accum.empty() ? nullptr : accum.data());  // .ppm / .png / .exr by extension
        std::printf("{\"meshes\": %zu, \"gpus\": %d, \"subframes\": %d, \"seconds\": %.4f, \"Mrays_s\": %.1f, \"Msamples_s\": %.1f, \"out\": \"%s\"}\n", meshes.size(), gpus,
                    subframes, secs, rays / secs / 1e6, samples / secs / 1e6, out.c_str());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "wavefront: %s\n", e.what());
        return 1;
    }
    return 0;
}
