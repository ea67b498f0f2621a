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
