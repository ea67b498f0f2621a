// MultiGpu.h -- the host side of a multi-GPU job for C++ users of SampleRenderer: which rank renders which screen tiles, and
// the shared frame the ranks' resolve kernels store into.  One process per GPU (the reference has no multi-GPU code,
// SURVEY.md 8e; main.cpp:328-330 only ever uses device 0).  Mirrors fovpathtracing_optixcodelatest_b200/partition.py
// function for function -- same cost model, same tie-breaks, byte-identical owner tables (tests/test_host_cpp.py).
//
//   fovpt_multigpu::partition     tiles of `tile` pixels assigned by foveal sample density (longest-processing-time greedy):
//                                 a pixel costs spp / factor^2 samples of the ring that finally owns it
//                                 (SimplePathtracer.cpp:137-209: fovea 32, mid 16/4, periphery 8/16)
//   fovpt_multigpu::interleaved   gaze-independent static assignment for a gaze that moves every frame
//   fovpt_multigpu::SharedFrame   fovpt_peer_frame_* (include/fovpt.h) as an object: rank 0 creates the frame and publishes a
//                                 128-byte handle, the others open it; present() is the per-frame exchange step
#pragma once
#include <cstdint>
#include <vector>

#include "SimplePathtracer.h"

namespace fovpt_multigpu {

struct TilePartition {
    int tile = 0, tiles_x = 0, tiles_y = 0;
    std::vector<uint8_t> owner;        // tiles_y x tiles_x, row-major: rank that renders the tile
    std::vector<double> loads;         // samples per rank under the cost model
};

// samples attributable to each pixel (row-major height x width)
std::vector<float> pixel_cost(int width, int height, uint2 gaze, int inner_radius, int outer_radius,
                              const int spp[3], bool foveated, int spp_uniform);
TilePartition partition(int width, int height, uint2 gaze, int inner_radius, int outer_radius, int n_ranks, int tile = 32);
TilePartition partition(int width, int height, uint2 gaze, const fovpt_config& cfg, int n_ranks, int tile = 32);
TilePartition interleaved(int width, int height, int n_ranks, int tile = 16);
// per-rank sample load of any owner table under the same cost model
std::vector<double> loads_of(const TilePartition& p, int width, int height, uint2 gaze, const fovpt_config& cfg, int n_ranks);

// this renderer traces only the launch indices that own pixels of `rank`'s tiles (fovpt_set_tile_owner)
void apply(SampleRenderer& r, const TilePartition& p, int rank);

class SharedFrame {
public:
    // rank 0 of `world`: allocates the shared RGBA8 + float4 frame for r's current frame size; handle() goes to the other ranks
    static SharedFrame create(SampleRenderer& r, int world);
    // ranks 1..world-1: map the owner's frame (CUDA IPC; NVLink peer stores)
    static SharedFrame open(SampleRenderer& r, const fovpt_peer_handle& handle, int rank);
    SharedFrame(SharedFrame&& o) noexcept;
    SharedFrame(const SharedFrame&) = delete;
    ~SharedFrame();

    const fovpt_peer_handle& handle() const { return m_handle; }
    bool owner() const { return m_owner; }
    // call after render()/render_async of every frame.  Contributor: raises its arrival flag behind the pixel stores.
    // Owner: makes its stream wait for every rank's arrival; the frame is complete for whatever is enqueued next.
    void present();
    // owner only, once the frame has been consumed (downloaded / displayed): contributors may overwrite it
    void release();

private:
    SharedFrame() = default;
    SampleRenderer* m_r = nullptr;
    fovpt_peer_handle m_handle{};
    bool m_owner = false;
};

}  // namespace fovpt_multigpu
