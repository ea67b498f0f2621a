// oracle/common.h -- TEST INFRASTRUCTURE: plumbing shared by the two CPU checkers
// (scene container, software texture fetch, row-parallel launch loop, three-pass schedule).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../include/fovpt_exact.h"
#include "bvh.h"
#include "orc_api.h"

namespace ocommon {

struct HostTexture { int w = 0, h = 0; std::vector<uint32_t> px; };

struct HostMesh {
    std::vector<float>    vtx;      // 3 per vertex
    std::vector<uint32_t> idx;      // 3 per triangle
    std::vector<float>    tc;       // 2 per vertex or empty
    unsigned char         material[104];
    int                   texture_id = -1;
};

// RGBA8 texture, normalized coordinates, wrap addressing, bilinear filter, normalized float read, no sRGB
// decode, no mip maps: the sampler state of SimplePathtracer.cpp:775-789, evaluated with the texture unit's
// fixed-point arithmetic as pinned against a real cudaTextureObject_t (include/fovpt_exact.h).
inline void tex_bilinear(const HostTexture& t, float x, float y, float out[4])
{
    fx_tex_bilinear_rgba8(t.px.data(), t.w, t.h, x, y, out);
}

inline int resolve_threads(int req)
{
    if (req > 0) return req;
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

// Runs fn(y) for y in [0,rows) on `threads` host threads, rows interleaved.  The result of a
// launch must not depend on the schedule; the only shared writes are the splat stores, which
// are disjoint between launch indices except for the clamped out-of-frame ones,
// so for those configurations callers pass threads = 1.
inline void parallel_rows(unsigned rows, int threads, const std::function<void(unsigned)>& fn)
{
    if (threads <= 1 || rows <= 1) { for (unsigned y = 0; y < rows; ++y) fn(y); return; }
    std::vector<std::thread> pool;
    std::atomic<unsigned> next{0};
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() { for (;;) { unsigned y = next.fetch_add(1); if (y >= rows) break; fn(y); } });
    for (auto& th : pool) th.join();
}

// One optixLaunch-equivalent of the three-pass frame (SimplePathtracer.cpp:137-209).
struct PassParams {
    uint32_t factor; int32_t fill; float r_inner, r_outer; uint32_t spp;
    uint32_t off_x, off_y; uint32_t redraw; bool zero_subframe; unsigned lw, lh;
};

// The reference's launch dimensions are float expressions truncated to unsigned
// (launchParams.frame.r_outer passed as a dimension, :184,206).
inline std::vector<PassParams> three_pass_schedule(const orc_fov_cfg& c, int w, int h, uint32_t cx, uint32_t cy)
{
    std::vector<PassParams> v;
    if (!c.foveated) {
        v.push_back({1u, 1, 0.f, 1000000000.f, (uint32_t)c.spp_uniform, 0u, 0u, 0u, true, (unsigned)w, (unsigned)h});
        return v;
    }
    const int ri = c.inner_radius, ro = c.outer_radius;
    v.push_back({4u, 4, (float)ro, 1000000000.f, (uint32_t)c.spp_periphery, 0u, 0u, 0u, false,
                 (unsigned)(w / 4), (unsigned)(h / 4)});
    float r2 = (float)(ro + 2);
    v.push_back({2u, 2, (float)ri, r2, (uint32_t)c.spp_mid, cx - (uint32_t)(ro + 2), cy - (uint32_t)(ro + 2), 1u, true,
                 (unsigned)r2, (unsigned)r2});
    float r3 = (float)(ri + 1);
    v.push_back({1u, 1, 0.f, r3, (uint32_t)c.spp_fovea, cx - (uint32_t)(ri + 1), cy - (uint32_t)(ri + 1), 1u, true,
                 (unsigned)(r3 * 2), (unsigned)(r3 * 2)});
    return v;
}

}  // namespace ocommon
