"""Screen-tile partition of one frame across GPUs, weighted by foveal sample density.

Pure host logic (numpy), unit-tested on CPU.  Every pixel belongs to exactly one tile and every
tile to exactly one rank; a rank renders the launch indices that own pixels of its tiles
(csrc/wavefront.cu k_owner_map) and nothing else, so the data path needs no exchange until the
final gather of the frame.  The reference has no multi-GPU code; the cost model is
the three-pass schedule of SimplePathtracer.cpp:137-209: a pixel costs spp / factor^2 samples of
the ring that finally owns it (fovea 32, mid 16/4, periphery 8/16 at the reference settings).
"""
import numpy as np


def pixel_cost(width, height, gaze, inner_radius, outer_radius, spp=(8, 16, 32), foveated=True, spp_uniform=4):
    """samples attributable to each pixel (float32 [h,w])"""
    if not foveated:
        return np.full((height, width), float(spp_uniform), np.float32)
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    r = np.sqrt((x - np.float32(gaze[0])) ** 2 + (y - np.float32(gaze[1])) ** 2)
    cost = np.full((height, width), spp[0] / 16.0, np.float32)
    cost[r <= outer_radius + 2] = spp[1] / 4.0
    cost[r <= inner_radius + 1] = float(spp[2])
    return cost


def tile_costs(cost, tile):
    h, w = cost.shape
    ty, tx = (h + tile - 1) // tile, (w + tile - 1) // tile
    pad = np.zeros((ty * tile, tx * tile), np.float64)
    pad[:h, :w] = cost
    return pad.reshape(ty, tile, tx, tile).sum(axis=(1, 3))


def assign_tiles(costs, n_ranks):
    """Longest-processing-time greedy: heaviest tile first onto the least loaded rank.
    Deterministic (ties: lower tile index, lower rank).  Returns uint8 [ty,tx] of owners."""
    ty, tx = costs.shape
    owner = np.zeros((ty, tx), np.uint8)
    if n_ranks <= 1:
        return owner
    flat = costs.reshape(-1)
    order = np.lexsort((np.arange(flat.size), -flat))
    load = np.zeros(n_ranks, np.float64)
    for t in order:
        r = int(np.argmin(load))
        owner.reshape(-1)[t] = r
        load[r] += flat[t]
    return owner


def partition(width, height, gaze, inner_radius, outer_radius, n_ranks, tile=32, spp=(8, 16, 32), foveated=True, spp_uniform=4):
    cost = pixel_cost(width, height, gaze, inner_radius, outer_radius, spp, foveated, spp_uniform)
    costs = tile_costs(cost, tile)
    owner = assign_tiles(costs, n_ranks)
    loads = np.array([costs[owner == r].sum() for r in range(max(1, n_ranks))])
    return owner, loads


def interleaved(width, height, n_ranks, tile=16):
    """Gaze-independent partition for a moving gaze point (BASELINE configs[3]): tile (tx,ty) goes to rank (tx + 3 ty) mod n.
    Every ring then covers many tiles of every rank wherever the gaze is, so the owner table never has to be rebuilt or
    re-uploaded between frames (the LPT assignment above costs ~150 ms of host time per gaze point at 2160x2160; a frame
    takes ~2 ms).  With 16-pixel tiles the per-rank sample load stays within 2 % of the mean for 2/4/8 ranks along the
    C4 gaze path (tests/test_host_logic.py); 16 is a multiple of the periphery splat (4), so only mid-ring 2x2 blocks can
    straddle tile edges."""
    ty, tx = (height + tile - 1) // tile, (width + tile - 1) // tile
    yy, xx = np.mgrid[0:ty, 0:tx]
    return ((xx + 3 * yy) % max(1, n_ranks)).astype(np.uint8)


def loads_of(owner, width, height, gaze, inner_radius, outer_radius, n_ranks, tile, spp=(8, 16, 32), foveated=True, spp_uniform=4):
    """per-rank sample load of a given owner table under the cost model of pixel_cost"""
    costs = tile_costs(pixel_cost(width, height, gaze, inner_radius, outer_radius, spp, foveated, spp_uniform), tile)
    return np.array([costs[owner == r].sum() for r in range(max(1, n_ranks))])


def frame_sample_count(width, height, gaze, inner_radius, outer_radius, spp=(8, 16, 32)):
    """camera paths one foveated frame starts: the closed form of SURVEY 8(d), evaluated in float32 exactly as the ring
    test of deviceProgram.cu:432-440 (periphery on the /4 grid, mid ring on a /2 grid around the gaze, fovea at full
    resolution; unsigned wrap of the pixel offset when the gaze is closer than a radius to the frame edge)"""
    f = np.float32
    c, ri, ro = gaze, inner_radius, outer_radius

    def count(lw, lh, factor, off, rin, rout):
        y, x = np.mgrid[0:lh, 0:lw]
        px = (x * factor + off[0]).astype(np.int64) & 0xffffffff
        py = (y * factor + off[1]).astype(np.int64) & 0xffffffff
        dx = px.astype(f) - f(c[0]); dy = py.astype(f) - f(c[1])
        r = np.sqrt(dx * dx + dy * dy, dtype=f)
        return int((~((r < f(rin)) | (r > f(rout)))).sum())
    n1 = count(width // 4, height // 4, 4, (0, 0), ro, 1e9)
    n2 = count(ro + 2, ro + 2, 2, (c[0] - (ro + 2), c[1] - (ro + 2)), ri, ro + 2)
    n3 = count(2 * (ri + 1), 2 * (ri + 1), 1, (c[0] - (ri + 1), c[1] - (ri + 1)), 0, ri + 1)
    return spp[0] * n1 + spp[1] * n2 + spp[2] * n3
