// main.cpp -- headless stand-in for the reference's GLFW application (PT_sv5_/main.cpp:297-481): the same call
// sequence against the same classes -- loadOBJ, SampleRenderer ctor, resize, setCamera, setProbe, then per frame
// set the gaze launchParams.frame.c, render(), downloadPixels() -- with a PPM written instead of a window.
//
//   fovpt_demo scene.obj [width height frames gaze_x gaze_y out.ppm]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "SimplePathtracer.h"

int main(int argc, char** argv)
{
    try {
        if (argc < 2) { std::cerr << "usage: " << argv[0] << " scene.obj [width height frames gaze_x gaze_y out.ppm]\n"; return 2; }
        const int w = argc > 2 ? std::atoi(argv[2]) : 960, h = argc > 3 ? std::atoi(argv[3]) : 540;
        const int frames = argc > 4 ? std::atoi(argv[4]) : 8;
        const unsigned gx = argc > 5 ? (unsigned)std::atoi(argv[5]) : (unsigned)w / 2, gy = argc > 6 ? (unsigned)std::atoi(argv[6]) : (unsigned)h / 2;
        const char* out = argc > 7 ? argv[7] : "fovpt_demo.ppm";

        Model* model = loadOBJ(argv[1]);
        float3 lo = make_float3(1e30f, 1e30f, 1e30f), hi = make_float3(-1e30f, -1e30f, -1e30f);
        for (auto m : model->meshes) for (auto& v : m->vertex) {
            lo = make_float3(fminf(lo.x, v.x), fminf(lo.y, v.y), fminf(lo.z, v.z)); hi = make_float3(fmaxf(hi.x, v.x), fmaxf(hi.y, v.y), fmaxf(hi.z, v.z));
        }
        const float3 centre = make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
        const float extent = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
        sutil::Camera camera(make_float3(centre.x, centre.y, centre.z + 1.4f * extent), centre, make_float3(0.f, 1.f, 0.f), 45.0f, w / float(h));

        // constant ambient probe, the reference's loadColor(2.5) path (main.cpp:175-187,229)
        ProbeData probe; probe.width = 64; probe.height = 32;
        std::vector<Color> texels((size_t)probe.width * probe.height, make_float4(2.5f, 2.5f, 2.5f, 1.0f));
        probe.data = texels.data();
        probe.BuildCDF();

        fovpt_config cfg; fovpt_default_config(&cfg);
        cfg.tmin = 1e-4f * extent;                       // the reference's 0.01 is tuned to Sponza's scale
        cfg.accumulate = 1; cfg.seed_all_passes_with_subframe = 1;
        cfg.exposure = 1.0f;                             // the reference's 2^4 (deviceProgram.cu:586) saturates a 2.5 ambient probe to white
        SampleRenderer sample(model, 0, &cfg);
        sample.resize(make_int2(w, h));
        sample.setCamera(camera);
        sample.setProbe(probe);

        std::vector<uint32_t> pixels((size_t)w * h);
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; ++f) {
            sample.launchParams.frame.c = make_uint2(gx, gy);      // gaze, as main.cpp:366-367 sets it every frame
            sample.render();
        }
        sample.downloadPixels(pixels.data());
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const fovpt_stats st = sample.stats();
        std::printf("%d frames %dx%d: %.2f ms/frame, %.1f Msamples/s, %.2f Grays/s, BVH %llu nodes built in %.2f ms\n", frames, w, h, ms / frames,
                    st.samples / ms / 1e3, (st.radiance_rays + st.shadow_rays) / ms / 1e6, (unsigned long long)st.bvh_nodes, st.bvh_build_ms);
        FILE* fp = std::fopen(out, "wb");
        if (fp) {
            std::fprintf(fp, "P6\n%d %d\n255\n", w, h);
            for (int y = h - 1; y >= 0; --y)                       // row 0 is the bottom of the image
                for (int x = 0; x < w; ++x) { const uint32_t p = pixels[(size_t)y * w + x]; std::fputc(p & 255, fp); std::fputc((p >> 8) & 255, fp); std::fputc((p >> 16) & 255, fp); }
            std::fclose(fp);
        }
        delete model;
        return 0;
    } catch (const std::runtime_error& e) {                        // the reference's convention (main.cpp:487-490)
        std::cerr << "FATAL ERROR: " << e.what() << std::endl;
        return 1;
    }
}
