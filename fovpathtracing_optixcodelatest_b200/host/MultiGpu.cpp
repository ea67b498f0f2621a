// MultiGpu.cpp -- see MultiGpu.h
#include "MultiGpu.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>
#include <string>

namespace fovpt_multigpu {

std::vector<float> pixel_cost(int width, int height, uint2 gaze, int inner_radius, int outer_radius,
                              const int spp[3], bool foveated, int spp_uniform)
{
    std::vector<float> cost((size_t)width * height);
    if (!foveated) { std::fill(cost.begin(), cost.end(), (float)spp_uniform); return cost; }
    const float periphery = (float)(spp[0] / 16.0), mid = (float)(spp[1] / 4.0), fovea = (float)spp[2];
    const float gx = (float)gaze.x, gy = (float)gaze.y;
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const float dx = (float)x - gx, dy = (float)y - gy;
            const float r = std::sqrt(dx * dx + dy * dy);          // float32 throughout, like the ring test it models
            float c = periphery;
            if (r <= (float)(outer_radius + 2)) c = mid;
            if (r <= (float)(inner_radius + 1)) c = fovea;
            cost[(size_t)y * width + x] = c;
        }
    return cost;
}

namespace {

std::vector<double> tile_costs(const std::vector<float>& cost, int width, int height, int tile, int& tx, int& ty)
{
    tx = (width + tile - 1) / tile; ty = (height + tile - 1) / tile;
    std::vector<double> t((size_t)tx * ty, 0.0);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) t[(size_t)(y / tile) * tx + x / tile] += (double)cost[(size_t)y * width + x];
    return t;
}

// heaviest tile first onto the least loaded rank; ties: lower tile index, lower rank
std::vector<uint8_t> assign_tiles(const std::vector<double>& costs, int n_ranks)
{
    std::vector<uint8_t> owner(costs.size(), 0);
    if (n_ranks <= 1) return owner;
    std::vector<size_t> order(costs.size());
    std::iota(order.begin(), order.end(), (size_t)0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return costs[a] > costs[b]; });
    std::vector<double> load((size_t)n_ranks, 0.0);
    for (size_t t : order) {
        const size_t r = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        owner[t] = (uint8_t)r;
        load[r] += costs[t];
    }
    return owner;
}

std::vector<double> loads(const std::vector<double>& costs, const std::vector<uint8_t>& owner, int n_ranks)
{
    std::vector<double> l((size_t)std::max(1, n_ranks), 0.0);
    for (size_t i = 0; i < costs.size(); ++i) l[owner[i]] += costs[i];
    return l;
}

void check(SampleRenderer& r, int rc)
{
    if (rc != 0) throw std::runtime_error(std::string("libfovpt: ") + fovpt_last_error(r.context()));
}

}  // namespace

TilePartition partition(int width, int height, uint2 gaze, const fovpt_config& cfg, int n_ranks, int tile)
{
    if (n_ranks < 1 || n_ranks > 256) throw std::runtime_error("partition: 1..256 ranks (the owner table holds one byte per tile)");
    if (tile < 1 || width < 1 || height < 1) throw std::runtime_error("partition: bad frame or tile size");
    const int spp[3] = {cfg.spp_periphery, cfg.spp_mid, cfg.spp_fovea};
    TilePartition p; p.tile = tile;
    const std::vector<double> costs = tile_costs(pixel_cost(width, height, gaze, cfg.inner_radius, cfg.outer_radius, spp, cfg.foveated != 0, cfg.spp_uniform),
                                                 width, height, tile, p.tiles_x, p.tiles_y);
    p.owner = assign_tiles(costs, n_ranks);
    p.loads = loads(costs, p.owner, n_ranks);
    return p;
}

TilePartition partition(int width, int height, uint2 gaze, int inner_radius, int outer_radius, int n_ranks, int tile)
{
    fovpt_config cfg; fovpt_default_config(&cfg);
    cfg.inner_radius = inner_radius; cfg.outer_radius = outer_radius;
    return partition(width, height, gaze, cfg, n_ranks, tile);
}

TilePartition interleaved(int width, int height, int n_ranks, int tile)
{
    if (n_ranks < 1 || n_ranks > 256 || tile < 1) throw std::runtime_error("interleaved: bad rank count or tile size");
    TilePartition p; p.tile = tile;
    p.tiles_x = (width + tile - 1) / tile; p.tiles_y = (height + tile - 1) / tile;
    p.owner.resize((size_t)p.tiles_x * p.tiles_y);
    for (int y = 0; y < p.tiles_y; ++y)
        for (int x = 0; x < p.tiles_x; ++x) p.owner[(size_t)y * p.tiles_x + x] = (uint8_t)((x + 3 * y) % n_ranks);
    return p;
}

std::vector<double> loads_of(const TilePartition& p, int width, int height, uint2 gaze, const fovpt_config& cfg, int n_ranks)
{
    const int spp[3] = {cfg.spp_periphery, cfg.spp_mid, cfg.spp_fovea};
    int tx = 0, ty = 0;
    const std::vector<double> costs = tile_costs(pixel_cost(width, height, gaze, cfg.inner_radius, cfg.outer_radius, spp, cfg.foveated != 0, cfg.spp_uniform),
                                                 width, height, p.tile, tx, ty);
    if (tx != p.tiles_x || ty != p.tiles_y) throw std::runtime_error("loads_of: the partition was made for another frame size");
    return loads(costs, p.owner, n_ranks);
}

void apply(SampleRenderer& r, const TilePartition& p, int rank)
{
    check(r, fovpt_set_tile_owner(r.context(), p.tile, p.tiles_x, p.tiles_y, p.owner.empty() ? nullptr : p.owner.data(), rank));
}

SharedFrame SharedFrame::create(SampleRenderer& r, int world)
{
    SharedFrame f; f.m_r = &r; f.m_owner = true;
    uint32_t* ptr = nullptr;
    const size_t npix = (size_t)r.launchParams.frame.size.x * r.launchParams.frame.size.y;
    check(r, fovpt_peer_frame_create(r.context(), npix, world, 0, &f.m_handle, &ptr));
    r.launchParams.frame.frame_buffer = (uchar4*)ptr;
    return f;
}

SharedFrame SharedFrame::open(SampleRenderer& r, const fovpt_peer_handle& handle, int rank)
{
    SharedFrame f; f.m_r = &r; f.m_owner = false; f.m_handle = handle;
    uint32_t* ptr = nullptr;
    check(r, fovpt_peer_frame_open(r.context(), &handle, rank, &ptr));
    r.launchParams.frame.frame_buffer = (uchar4*)ptr;
    return f;
}

SharedFrame::SharedFrame(SharedFrame&& o) noexcept : m_r(o.m_r), m_handle(o.m_handle), m_owner(o.m_owner) { o.m_r = nullptr; }

SharedFrame::~SharedFrame()
{
    if (m_r) fovpt_peer_frame_close(m_r->context());
}

void SharedFrame::present() { check(*m_r, m_owner ? fovpt_peer_frame_wait(m_r->context()) : fovpt_peer_frame_signal(m_r->context())); }
void SharedFrame::release() { if (m_owner) check(*m_r, fovpt_peer_frame_release(m_r->context())); }

}  // namespace fovpt_multigpu
