// wavefront.cuh -- device-side data model of the wavefront path tracer (kernels: wavefront.cu).
//
// One camera path ("sample", one iteration of reference deviceProgram.cu:448-539) owns slot `pid`
// of the per-path accumulators; its ray state travels through dense per-bounce buffers that connect the stages:
//
//   k_owner_map -> k_raygen -> [ k_trace -> k_shade ] x (max_depth+1) -> k_trace -> k_resolve_li -> k_resolve_px
// (k_trace(i) extends the radiance rays of iteration i and resolves the shadow rays emitted by k_shade(i-1))
//
// pid = pass.pid_base + li * pass.spp + k, with li the row-major launch index of the pass
// (optixGetLaunchIndex) and k the sample number, so the spp samples of one launch index are
// adjacent and can be summed in the reference's order by k_resolve_li.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "bvh.cuh"
#include "fovpt_layout.h"
#include "shade.cuh"

namespace fovpt {

constexpr uint32_t kNoHit = 0xffffffffu;
constexpr uint32_t kNoOwner = 0xffffffffu;
constexpr int kMaxPasses = 3;

// path flag bits (stored as int bits in alpha.w)
constexpr uint32_t kFlagSecondary = 1u << 8;   // RAY_STATE_FLAGS_SECONDARY_RAY
constexpr uint32_t kFlagAlphaOne = 1u << 9;    // prd.alpha was set to 1 by a non-catcher hit
constexpr uint32_t kDepthMask = 0xffu;

// Per-path accumulators, indexed by path slot pid (what k_resolve_li sums per launch index).
struct PathState {
    float4* direct;             // directLight.xyz
    float4* indirect;           // indirectLight.xyz
    float4* alpha;              // alpha.xyz
    // denoiser features (fovpt_config.features; null when off): what prd.normal / prd.albedo held after the FIRST trace
    // of the path (deviceProgram.cu:509-512, :653-665); feat_n.w counts how many times the raygen loop adds them
    float4* feat_n;             // normal.xyz, multiplicity
    float4* feat_a;             // albedo.xyz
};

// The wavefront itself is DENSE: the state of the rays of one bounce lives in queue order, so every kernel reads and
// writes it fully coalesced and without a pid indirection; k_shade appends the continuing paths to the `next` buffer.
struct RayBuf {                 // by queue slot
    float4* o;                  // origin.xyz
    float4* d;                  // direction.xyz
    float4* thr;                // pathThroughput.xyz, rayEta
    uint2*  rng;                // Random state (maths.h:170-227)
    uint2*  meta;               // (pid, depth | flags)
};
struct ShadowBuf {              // by shadow-queue slot: the NEE rays emitted by one shading pass
    float4*   o;                // origin.xyz
    float4*   d;                // direction.xyz
    float4*   pend;             // throughput * NEE value, bits(mode): bit0 catcher (add to alpha when occluded), bit1 primary
    uint32_t* pid;
};

struct SceneDev {
    DeviceBvh           bvh;
    const uint32_t*     tri_mesh;     // global triangle id -> mesh id
    const uint32_t*     tri_vidx;     // 3 global vertex indices per global triangle
    const float2*       texcoords;    // per global vertex (zero where the mesh has none)
    const MaterialBlob* materials;    // per mesh
    const int2*         mesh_tex;     // per mesh: (texture id or -1, has texcoords)
    const int4*         tex_info;     // per texture: (texel offset, w, h, 0)
    const uint32_t*     texels;
    int                 any_catcher;
    uint32_t            n_meshes, n_textures, n_vertices;     // table sizes (bounds-checked build)
};

struct PassDev {
    unsigned lw, lh, factor, fill, spp, redraw, subframe;
    unsigned sub_step;                // subframe of layer s of a batch = subframe + s * sub_step (0: the pass reseeds with the same subframe)
    int      off_x, off_y;            // offset as a signed number (uint2 offset wraps)
    unsigned uoff_x, uoff_y;          // the same offset, raw
    float    r_inner, r_outer;
    unsigned pid_base, li_base;
};

struct FrameDev {
    int       w, h;
    unsigned  cx, cy;
    float     eye[3], U[3], V[3], W[3];
    ProbeBlob probe;
    ProbeGuide guide;           // null unless the probe's CDFs are the ones fovpt_probe_upload built guides for
    float4*   accum;
    uint32_t* frame;
    float4   *feat_color, *feat_normal, *feat_albedo;      // LaunchParams.frame.{color,normal,albedo}_buffer when fovpt_config.features, else null
    int       npass;
    PassDev   pass[kMaxPasses];
    // a batch of nsub accumulated subframes traced as ONE wavefront (fovpt_render_frames): per-launch-index buffers hold
    // nsub layers of li_total entries, k_resolve_px replays the per-frame accumulation over the layers in order
    int       nsub;
    unsigned  li_total;
    unsigned  path_cap;         // slots of the per-path / per-ray buffers (bounds-checked build)
    // config
    int   max_depth;
    float tmin, tmax, exposure;
    int   accumulate, skip_wasted, filter;
    // partition
    const uint8_t* tile_owner;
    int   tile_size, tiles_x, tiles_y, rank;
};

// Work of one raygen / resolve launch: up to kMaxChunks row ranges of passes, each over a range of batch layers.  Item i of
// chunk c (first[c] <= i < first[c+1]) decomposes as ((layer - s0) * rows * lw + (ly - row0) * lw + lx) * spp + k, and the
// path slot is pid_base + (i - first[c]), so the spp samples of one (layer, launch index) are adjacent.
constexpr int kMaxChunks = 8;
struct ChunkDev { int pass; unsigned row0, rows, s0, ns, pid_base, first; };
struct ChunkTable { int n; unsigned total; ChunkDev c[kMaxChunks]; };

struct Queues {
    RayBuf    cur, next;
    ShadowBuf shadow;
    float4*   hit;              // by slot of `cur`: t, u, v, bits(sorted triangle slot | kNoHit)
    uint32_t* counts;   // [0] cur [1] next [2],[5] shadow (alternating per iteration) [3] fetch cursor of k_trace [6] CTAs of k_shade done  (8 words)
};

struct DevStats {
    unsigned long long samples, radiance_rays, shadow_rays;
    unsigned long long dbg[16];     // -DFOVPT_TRACE_STATS builds only: warp-step counters of k_trace (closest 0..7, any 8..15)
};

}  // namespace fovpt
