// wavefront_launch.h -- host-callable wrappers around the kernels of wavefront.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "wavefront.cuh"

namespace fovpt {

struct LaunchCfg {
    cudaStream_t stream;
    int grid_trace;   // persistent grid of the traversal kernels  (SMs x resident CTAs)
    int grid_shade;   // persistent grid of the shading kernel
    int grid_light;   // grid of the light-weight per-pixel / per-launch-index kernels
};

void launch_owner_map(const LaunchCfg&, const FrameDev&, uint32_t* owner_map, uint8_t* needed);
void launch_raygen(const LaunchCfg&, const FrameDev&, const ChunkTable&, const uint8_t* needed, const PathState&, const Queues&,
                   float4* li_backplate, DevStats*);
// shadow_slot: index into Queues::counts (2 or 5, alternating per iteration) of the shadow-ray counter in use
void launch_trace(const LaunchCfg&, const SceneDev&, const PathState&, const Queues&, int shadow_slot, float tmin, float tmax, DevStats*);
void launch_shade(const LaunchCfg&, const FrameDev&, const SceneDev&, const PathState&, const Queues&, int shadow_slot, bool fast);
// primary-hit normal / albedo into PathState::feat_n / feat_a; between the first launch_trace and the first launch_shade
void launch_features(const LaunchCfg&, const SceneDev&, const PathState&, const Queues&, bool fast);
// peer-frame flags (api.cu fovpt_peer_frame_*): raise *flag to seq / wait until flags[0..count) (except `skip`) reached seq
void launch_peer_signal(const LaunchCfg&, uint32_t* flag, uint32_t seq);
void launch_peer_wait(const LaunchCfg&, const uint32_t* flags, int count, int skip, uint32_t seq, uint32_t* err, long long timeout_clocks);
// li_normal / li_albedo: per-(layer, launch index) denoiser features, or null (fovpt_config.features off)
void launch_resolve_li(const LaunchCfg&, const FrameDev&, const ChunkTable&, const uint8_t* needed, const PathState&,
                       const float4* li_backplate, float4* li_color, float4* li_normal, float4* li_albedo);
// accum_mirror: second destination of the accumulation value (the owner rank's shared buffer of a multi-GPU job) or null
void launch_resolve_px(const LaunchCfg&, const FrameDev&, const uint32_t* owner_map, const float4* li_color, const uint8_t* needed,
                       float4* accum_mirror, const float4* li_normal, const float4* li_albedo, bool fast);
void launch_probe_guides(const LaunchCfg&, const ProbeBlob&, uint16_t* gy, uint16_t* gx);
void launch_primary_hits(const LaunchCfg&, const FrameDev&, const SceneDev&, int32_t* tri, float* tuv, float* dir);
void launch_trace_query(const LaunchCfg&, const SceneDev&, const float* rays, size_t n, int mode, int32_t* tri, float* tuv);
void launch_texture_probe(const LaunchCfg&, cudaTextureObject_t tex, const uint32_t* texels, int w, int h, const float2* uv, size_t n, int mode,
                          float4* sw, float4* hw);
void query_occupancy(int* extend_blocks_per_sm, int* shade_blocks_per_sm);
unsigned read_check_failure();      // FOVPT_CHECKED builds: site of the first failed bounds check (bvh.cuh), else 0

}  // namespace fovpt
