// bvh_build.cu -- on-device BVH construction; replaces optixAccelBuild/optixAccelCompact
// (reference PT_sv5_/SimplePathtracer.cpp:602-746).  Pipeline, all kernels on the context stream:
//
//   k_rebase_indices  (scene commit) mesh-local indices -> indices into the vertex soup, mesh id per triangle
//   k_tri_setup     per triangle: fx_make_tri record, AABB, scene bounds (atomic min/max)
//   k_morton        63-bit Morton code of the AABB centre (21 bits per axis)
//   radix sort      hand-written LSD sort, 8 x 8-bit passes over (u64 key, u32 triangle id):
//                   k_radix_hist -> scan -> k_radix_scatter (stable, match_any ranking); scans of more than
//                   128 k entries run tiled over the whole GPU (k_scan_tile_sums / k_scan_tiles)
//   PLOC            parallel locally-ordered clustering (Meister & Bittner 2018) over the Morton-ordered
//                   triangles: every round each cluster finds, within +-kPlocRadius positions, the
//                   neighbour whose merged box has the smallest surface area (k_ploc_nearest, boxes
//                   staged in shared memory); mutual nearest neighbours merge into a new binary node
//                   (k_ploc_merge); survivors are compacted in order (k_scan_u32 + k_ploc_compact).
//                   Bottom-up, surface-area driven: trees close to a binned-SAH build at a few dozen
//                   rounds of trivially parallel kernels (the first path used a Karras LBVH: +25 % nodes
//                   and +24 % triangle tests per ray on the atrium).
//   top_sah         (host, top_sah.h) once at most 1 M clusters are left -- at once for scenes of up to 1 M triangles -- the
//                   tree ABOVE them is built top-down with the binned surface-area heuristic; scenes under 4096 triangles
//                   stay pure PLOC.  Measured: k_trace -11 % (C3) / -12.5 % (C5) against the pure PLOC tree.
//   k_collapse4     breadth-first collapse of the binary tree into the 4-wide, 8-bit-quantised nodes of bvh.cuh:
//                   a wide node takes a binary node's two children and twice replaces the child
//                   of largest surface area by its own two children; subtrees with <= leaf_max (default 2, at
//                   most 4) triangles become leaves and receive consecutive slots of the final triangle order.  One launch
//                   per level, work queue + node allocation through device counters.
//   k_gather_tris   48-byte triangle records into leaf order
//
// Boxes are padded by pad = max|coordinate| / 65536 so that the fma-form slab test of bvh.cuh can
// never reject a box whose triangle the exact test (fovpt_exact.h) accepts.  Hit results are
// independent of the tree (tie-break in fx_hit_better), so tree quality is purely a speed matter.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <exception>
#include <thread>
#include <vector>

#include "bvh.cuh"
#include "bvh_build.h"
#include "top_sah.h"

namespace fovpt {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ int float_as_ordered(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_as_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_init_bounds(int* b)
{
    if (threadIdx.x < 3) b[threadIdx.x] = 0x7fffffff;          // mins (ordered-int space)
    else if (threadIdx.x < 6) b[threadIdx.x] = (int)0x80000000; // maxs
}

// vertices: global float3 soup ; indices: global (already rebased) uint3
__global__ void k_tri_setup(const float* __restrict__ vtx, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ tri_mesh, uint32_t ntri,
                            float4* __restrict__ tri_rec, float4* __restrict__ box_lo, float4* __restrict__ box_hi,
                            int* __restrict__ bounds)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < ntri) {
        const uint32_t ia = idx[3 * (size_t)i], ib = idx[3 * (size_t)i + 1], ic = idx[3 * (size_t)i + 2];
        fx_vec3 A = {vtx[3 * (size_t)ia], vtx[3 * (size_t)ia + 1], vtx[3 * (size_t)ia + 2]};
        fx_vec3 B = {vtx[3 * (size_t)ib], vtx[3 * (size_t)ib + 1], vtx[3 * (size_t)ib + 2]};
        fx_vec3 C = {vtx[3 * (size_t)ic], vtx[3 * (size_t)ic + 1], vtx[3 * (size_t)ic + 2]};
        fx_tri t = fx_make_tri(A, B, C);
        tri_rec[3 * (size_t)i + 0] = make_float4(t.v0.x, t.v0.y, t.v0.z, __int_as_float((int)i));
        tri_rec[3 * (size_t)i + 1] = make_float4(t.v1.x, t.v1.y, t.v1.z, __int_as_float((int)tri_mesh[i]));   // mesh id rides along: saves the shading kernel a dependent load
        tri_rec[3 * (size_t)i + 2] = make_float4(t.v2.x, t.v2.y, t.v2.z, 0.f);
        lo[0] = fminf(A.x, fminf(B.x, C.x)); hi[0] = fmaxf(A.x, fmaxf(B.x, C.x));
        lo[1] = fminf(A.y, fminf(B.y, C.y)); hi[1] = fmaxf(A.y, fmaxf(B.y, C.y));
        lo[2] = fminf(A.z, fminf(B.z, C.z)); hi[2] = fmaxf(A.z, fmaxf(B.z, C.z));
        box_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        box_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    // warp reduce then one atomic per warp per component
    for (int a = 0; a < 3; ++a) {
        float l = lo[a], h = hi[a];
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(&bounds[a], float_as_ordered(l));
            atomicMax(&bounds[3 + a], float_as_ordered(h));
        }
    }
}

__device__ __forceinline__ uint64_t expand21(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, uint32_t ntri,
                         const int* __restrict__ bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntri) return;
    float lo[3], ext[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = ordered_as_float(bounds[a]);
        ext[a] = ordered_as_float(bounds[3 + a]) - lo[a];
    }
    const float4 bl = box_lo[i], bh = box_hi[i];
    const float c[3] = {0.5f * (bl.x + bh.x), 0.5f * (bl.y + bh.y), 0.5f * (bl.z + bh.z)};
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
        float f = ext[a] > 0.f ? (c[a] - lo[a]) / ext[a] : 0.f;
        f = fminf(fmaxf(f * 2097152.0f, 0.f), 2097151.0f);
        q[a] = (uint32_t)f;
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[i] = i;
}

// ------------------------------------------------------------------ radix sort (8-bit digits)
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                           uint32_t* __restrict__ hist, uint32_t nblocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];     // digit-major
}

// single-block exclusive scan of n u32 values (in place); n up to a few million is fine.
__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ total_out)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? data[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sums[lane];
            uint32_t s = w;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (uint32_t)o) s += y; }
            warp_sums[lane] = s - w;    // exclusive
        }
        __syncthreads();
        const uint32_t c = carry;
        const uint32_t incl = x + warp_sums[wid] + c;
        if (i < n) data[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// ---- three-step exclusive scan for large arrays: per-tile sums -> single-block scan of the sums -> per-tile scan + offset.
// (k_scan_u32 alone walks the whole array with one CTA: fine for the 300 k-entry histograms, 10 ms per call at 10 M.)
constexpr int SC_TILE = 2048;          // entries per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) k_scan_tile_sums(const uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ sums)
{
    __shared__ uint32_t ws[8];
    const uint32_t base = blockIdx.x * SC_TILE;
    uint32_t v = 0;
    for (int r = 0; r < 8; ++r) { const uint32_t i = base + r * 256 + threadIdx.x; if (i < n) v += data[i]; }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += ws[w]; sums[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(256) k_scan_tiles(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ tile_offsets)
{
    __shared__ uint32_t ws[8];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = tile_offsets[blockIdx.x];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * SC_TILE;
    for (int r = 0; r < 8; ++r) {
        const uint32_t i = base + r * 256 + threadIdx.x;
        const uint32_t v = i < n ? data[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
        if (lane == 31) ws[wid] = x;
        __syncthreads();
        uint32_t before = carry;
        for (uint32_t w = 0; w < wid; ++w) before += ws[w];
        if (i < n) data[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                              uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                              uint32_t n, int shift, const uint32_t* __restrict__ offsets, uint32_t nblocks)
{
    __shared__ uint32_t running[256];               // keys of each digit already placed by earlier rounds
    __shared__ uint32_t warp_cnt[RS_THREADS / 32][256];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    running[threadIdx.x] = offsets[threadIdx.x * nblocks + blockIdx.x];
    const uint32_t base = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; ++r) {
        for (int w = 0; w < RS_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        uint64_t key = 0; uint32_t val = 0; uint32_t digit = 0;
        if (valid) { key = keys_in[i]; val = vals_in[i]; digit = (uint32_t)(key >> shift) & 255u; }
        // invalid lanes get a digit no valid lane shares in the match (use bit 8)
        const uint32_t mkey = valid ? digit : 256u + lane;
        const uint32_t peers = __match_any_sync(0xffffffffu, mkey);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_cnt[wid][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t before = 0;
            for (uint32_t w = 0; w < wid; ++w) before += warp_cnt[w][digit];
            const uint32_t pos = running[digit] + before + rank_in_warp;
            keys_out[pos] = key; vals_out[pos] = val;
        }
        __syncthreads();
        uint32_t tot = 0;
        for (int w = 0; w < RS_THREADS / 32; ++w) tot += warp_cnt[w][threadIdx.x];
        running[threadIdx.x] += tot;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ PLOC
// node ids: >= 0 inner binary node, < 0 leaf = ~(global triangle id)
#ifndef FOVPT_PLOC_RADIUS
#define FOVPT_PLOC_RADIUS 32
#endif
constexpr int kPlocRadius = FOVPT_PLOC_RADIUS;
constexpr int kPlocThreads = 256;

__device__ __forceinline__ float union_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi)
{
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x);
    const float dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y);
    const float dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_ploc_init(uint32_t n, const uint32_t* __restrict__ sorted_ids, const float4* __restrict__ box_lo,
                            const float4* __restrict__ box_hi, int* __restrict__ cid, float4* __restrict__ clo, float4* __restrict__ chi)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = sorted_ids[i];
    cid[i] = ~(int)t; clo[i] = box_lo[t]; chi[i] = box_hi[t];
}

// nn[i] = position of the cluster within [i-R, i+R] whose union with cluster i has the smallest area, boxes of the
// block's window staged in shared memory.  Ties are broken SYMMETRICALLY: first the partner i ^ 1, then the nearer
// position -- in a run of clusters with identical boxes (duplicated or coincident triangles) every cluster then finds a
// mutual partner and the run halves per round; "smallest position wins" made only the pair (0, 1) mutual, one merge per
// round and a left-deep chain the collapse cannot rebalance.  forced != 0: nn[i] = i ^ 1 regardless of area -- a
// balanced tree over the Morton order, the fallback when the area-driven tree comes out too deep for the traversal stack.
__global__ void __launch_bounds__(kPlocThreads) k_ploc_nearest(uint32_t n, const float4* __restrict__ clo, const float4* __restrict__ chi,
                                                               uint32_t* __restrict__ nn, int forced)
{
    if (forced) {
        const uint32_t i = blockIdx.x * kPlocThreads + threadIdx.x;
        if (i < n) nn[i] = (i ^ 1u) < n ? (i ^ 1u) : i;
        return;
    }
    __shared__ float4 slo[kPlocThreads + 2 * kPlocRadius], shi[kPlocThreads + 2 * kPlocRadius];
    const int base = (int)(blockIdx.x * kPlocThreads) - kPlocRadius;
    for (int k = threadIdx.x; k < kPlocThreads + 2 * kPlocRadius; k += kPlocThreads) {
        const int j = base + k;
        if (j >= 0 && j < (int)n) { slo[k] = clo[j]; shi[k] = chi[j]; }
    }
    __syncthreads();
    const int i = (int)(blockIdx.x * kPlocThreads + threadIdx.x);
    if (i >= (int)n) return;
    const float4 lo = slo[threadIdx.x + kPlocRadius], hi = shi[threadIdx.x + kPlocRadius];
    float best = 3.0e38f; int bestj = i, bestp = 0x7fffffff;
    for (int d = -kPlocRadius; d <= kPlocRadius; ++d) {
        const int j = i + d;
        if (d == 0 || j < 0 || j >= (int)n) continue;
        const float a = union_area(lo, hi, slo[threadIdx.x + kPlocRadius + d], shi[threadIdx.x + kPlocRadius + d]);
        const int pref = j == (i ^ 1) ? 0 : 2 * (d < 0 ? -d : d) + (d > 0 ? 1 : 0);      // symmetric between i and i ^ 1
        if (a < best || (a == best && pref < bestp)) { best = a; bestj = j; bestp = pref; }
    }
    nn[i] = (uint32_t)bestj;
}

// mutual nearest neighbours (i < j) become binary node `id`; cluster i carries it on, cluster j dies
__global__ void k_ploc_merge(uint32_t n, const uint32_t* __restrict__ nn, int* __restrict__ cid, float4* __restrict__ clo, float4* __restrict__ chi,
                             uint32_t* __restrict__ node_counter, int* __restrict__ left, int* __restrict__ right,
                             float4* __restrict__ nlo, float4* __restrict__ nhi, uint32_t* __restrict__ cnt, uint32_t* __restrict__ alive)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = nn[i];
    uint32_t keep = 1u;
    if (j != i && nn[j] == i) {
        if (i < j) {
            const uint32_t id = atomicAdd(node_counter, 1u);
            const int a = cid[i], b = cid[j];
            const float4 alo = clo[i], ahi = chi[i], blo = clo[j], bhi = chi[j];
            const float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
            const float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
            left[id] = a; right[id] = b; nlo[id] = lo; nhi[id] = hi;
            cnt[id] = (a < 0 ? 1u : cnt[a]) + (b < 0 ? 1u : cnt[b]);
            cid[i] = (int)id; clo[i] = lo; chi[i] = hi;
        } else keep = 0u;
    }
    alive[i] = keep;
}

__global__ void k_ploc_compact(uint32_t n, const uint32_t* __restrict__ alive, const uint32_t* __restrict__ pos,
                               const int* __restrict__ cid, const float4* __restrict__ clo, const float4* __restrict__ chi,
                               int* __restrict__ cid_out, float4* __restrict__ clo_out, float4* __restrict__ chi_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !alive[i]) return;
    const uint32_t p = pos[i];
    cid_out[p] = cid[i]; clo_out[p] = clo[i]; chi_out[p] = chi[i];
}

// ------------------------------------------------------------------ top of the tree: binned SAH over the last clusters
// PLOC's +-radius window over the Morton order is a good neighbourhood while clusters are small; further up Morton neighbours
// are far apart in space and greedy pairwise merging decides the most visited nodes of the tree.  So the bottom-up phase stops
// once at most kTopSahMax clusters are left (scenes of up to that many triangles skip it altogether), and the tree ABOVE them is
// built top-down on the host with the surface-area heuristic: 32 centroid bins per axis, cost = area x triangles below, subtrees
// of more than 16 k clusters on threads of their own.  Measured against the pure PLOC tree (same kernels, k_trace time per step):
// C3 -11 %, C5 (10 M triangles, the top 1 M clusters) -12.5 %; the host pass takes ~50 ms per 64 k clusters on one thread.
// Scenes below kTopSahMinTris keep the PLOC tree: for a few dozen triangles pairing by merged area (the two halves of a quad
// first) beats splitting by centroid (Cornell box: +7 % with SAH).
#ifndef FOVPT_TOP_SAH
#define FOVPT_TOP_SAH 1048576
#endif
constexpr uint32_t kTopSahMax = FOVPT_TOP_SAH, kTopSahMinTris = 4096;

__global__ void k_cluster_counts(uint32_t n, const int* __restrict__ cid, const uint32_t* __restrict__ cnt, uint32_t* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cid[i] < 0 ? 1u : cnt[cid[i]];
}

using topsah::TopCluster; using topsah::TopNode; using topsah::top_sah;

// ------------------------------------------------------------------ 4-wide collapse
struct CollapseItem { int bvh2; int node4; uint32_t first; uint32_t pad_; };   // first = slot of the subtree's first triangle in the final order

__device__ __forceinline__ float box_area(const float4 lo, const float4 hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Encodes one 4-wide node in the 64-byte quantised layout of bvh.cuh.  Child boxes are first padded
// by `pad` (absorbs the rounding of the fma slab test, see the header of this file), then snapped
// OUTWARD onto the node's 8-bit grid  plane(q) = p + q * 2^(e-127),  q in 0..255 : the check against
// the exact (double) plane position makes the outward rounding unconditional.  Unused slots get the
// inverted byte box lo = 255, hi = 0, which the sign-selecting slab test can never enter.
__device__ __forceinline__ void write_node4(uint4* __restrict__ o, const float4 lo[4], const float4 hi[4], const int enc[4], int nc, float pad)
{
    float p[3] = {3.0e38f, 3.0e38f, 3.0e38f}, top[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int j = 0; j < nc; ++j) {
        p[0] = fminf(p[0], lo[j].x - pad); p[1] = fminf(p[1], lo[j].y - pad); p[2] = fminf(p[2], lo[j].z - pad);
        top[0] = fmaxf(top[0], hi[j].x + pad); top[1] = fmaxf(top[1], hi[j].y + pad); top[2] = fmaxf(top[2], hi[j].z + pad);
    }
    uint32_t ebits[3]; double scale[3];
    for (int a = 0; a < 3; ++a) {
        const double ext = (double)top[a] - (double)p[a];
        int e = 0;
        (void)frexp(fmax(ext, 1e-30) / 255.0, &e);             // ext/255 = m * 2^e, m in [0.5,1)  ->  2^e >= ext/255
        e = max(-120, min(e, 120));
        ebits[a] = (uint32_t)(e + 127);
        scale[a] = ldexp(1.0, e);
    }
    uint32_t qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
    for (int j = 0; j < 4; ++j) {
        uint32_t l3[3] = {255u, 255u, 255u}, h3[3] = {0u, 0u, 0u};
        if (j < nc) {
            const float cl[3] = {lo[j].x - pad, lo[j].y - pad, lo[j].z - pad}, ch[3] = {hi[j].x + pad, hi[j].y + pad, hi[j].z + pad};
            for (int a = 0; a < 3; ++a) {
                int ql = (int)floor(((double)cl[a] - (double)p[a]) / scale[a]);
                int qh = (int)ceil(((double)ch[a] - (double)p[a]) / scale[a]);
                ql = max(0, min(ql, 255)); qh = max(0, min(qh, 255));
                while (ql > 0 && (double)p[a] + ql * scale[a] > (double)cl[a]) --ql;
                while (qh < 255 && (double)p[a] + qh * scale[a] < (double)ch[a]) ++qh;
                l3[a] = (uint32_t)ql; h3[a] = (uint32_t)qh;
            }
        }
        for (int a = 0; a < 3; ++a) { qlo[a] |= l3[a] << (8 * j); qhi[a] |= h3[a] << (8 * j); }
    }
    o[0] = make_uint4(__float_as_uint(p[0]), __float_as_uint(p[1]), __float_as_uint(p[2]), ebits[0] | (ebits[1] << 8) | (ebits[2] << 16));
    o[1] = make_uint4((uint32_t)enc[0], (uint32_t)enc[1], (uint32_t)enc[2], (uint32_t)enc[3]);
    o[2] = make_uint4(qlo[0], qhi[0], qlo[1], qhi[1]);
    o[3] = make_uint4(qlo[2], qhi[2], 0u, 0u);
}

// writes the (<= 4) triangle ids below binary node / leaf `c` to sorted_ids[first ...], depth-first left to right
__device__ __forceinline__ void emit_leaf_ids(int c, uint32_t first, const int* __restrict__ left, const int* __restrict__ right,
                                              uint32_t* __restrict__ sorted_ids)
{
    int stack[8]; int sp = 0;
    stack[sp++] = c;
    while (sp > 0) {
        const int x = stack[--sp];
        if (x < 0) sorted_ids[first++] = (uint32_t)~x;
        else { stack[sp++] = right[x]; stack[sp++] = left[x]; }
    }
}

// one level: items_in[0 .. counts[level]) -> nodes4 + items_out[0 .. counts[level+1])
__global__ void k_collapse4(const CollapseItem* __restrict__ items_in, CollapseItem* __restrict__ items_out,
                            uint32_t* __restrict__ counts, int level, uint32_t* __restrict__ node_counter,
                            const int* __restrict__ left, const int* __restrict__ right, const uint32_t* __restrict__ cnt,
                            const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                            const float4* __restrict__ nlo, const float4* __restrict__ nhi,
                            int leaf_max, float pad, uint4* __restrict__ nodes4, uint32_t* __restrict__ sorted_ids)
{
    const uint32_t n = counts[level];
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x) {
        const CollapseItem item = items_in[it];
        int c[4]; int nc = 2;
        c[0] = left[item.bvh2]; c[1] = right[item.bvh2]; c[2] = c[3] = 0;
        for (int rep = 0; rep < 2; ++rep) {
            int pick = -1; float best = -1.f;
            for (int j = 0; j < nc; ++j) {
                if (c[j] < 0 || (int)cnt[c[j]] <= leaf_max) continue;             // a leaf as a whole
                const float a = box_area(nlo[c[j]], nhi[c[j]]);
                if (a > best) { best = a; pick = j; }
            }
            if (pick < 0) break;
            const int x = c[pick];
            c[pick] = left[x]; c[nc++] = right[x];
        }
        int n_inner = 0;
        bool inner[4];
        for (int j = 0; j < 4; ++j) {
            inner[j] = j < nc && c[j] >= 0 && (int)cnt[c[j]] > leaf_max;
            n_inner += inner[j] ? 1 : 0;
        }
        uint32_t base = 0, qbase = 0;
        if (n_inner) { base = atomicAdd(node_counter, (uint32_t)n_inner); qbase = atomicAdd(&counts[level + 1], (uint32_t)n_inner); }
        float4 lo[4], hi[4]; int enc[4];
        int k = 0;
        uint32_t first = item.first;
        for (int j = 0; j < 4; ++j) {
            if (j >= nc) {                                                         // unused slot (write_node4 gives it an inverted box)
                lo[j] = make_float4(0.f, 0.f, 0.f, 0.f); hi[j] = lo[j];
                enc[j] = 0x7fffffff;
                continue;
            }
            const uint32_t ntri = c[j] < 0 ? 1u : cnt[c[j]];
            if (c[j] < 0) { const uint32_t id = (uint32_t)~c[j]; lo[j] = box_lo[id]; hi[j] = box_hi[id]; }
            else { lo[j] = nlo[c[j]]; hi[j] = nhi[c[j]]; }
            if (inner[j]) {
                enc[j] = (int)(base + k);
                items_out[qbase + k] = CollapseItem{c[j], (int)(base + k), first, 0u};
                ++k;
            } else {
                enc[j] = ~(int)((first << 2) | (ntri - 1u));
                emit_leaf_ids(c[j], first, left, right, sorted_ids);
            }
            first += ntri;
        }
        write_node4(nodes4 + 4 * (size_t)item.node4, lo, hi, enc, nc, pad);
    }
}

// scenes with <= leaf_max triangles: a root whose slot 0 is the only leaf, boxed by the scene bounds
__global__ void k_emit_single4(int n, const int* __restrict__ bounds, float pad, uint4* __restrict__ nodes4)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float4 lo[4], hi[4]; int enc[4];
    for (int j = 0; j < 4; ++j) { lo[j] = make_float4(0.f, 0.f, 0.f, 0.f); hi[j] = lo[j]; enc[j] = 0x7fffffff; }
    lo[0] = make_float4(ordered_as_float(bounds[0]), ordered_as_float(bounds[1]), ordered_as_float(bounds[2]), 0.f);
    hi[0] = make_float4(ordered_as_float(bounds[3]), ordered_as_float(bounds[4]), ordered_as_float(bounds[5]), 0.f);
    enc[0] = ~(int)((0u << 2) | (uint32_t)(n - 1));
    write_node4(nodes4, lo, hi, enc, 1, pad);
}

__global__ void k_seed_collapse(CollapseItem* items, uint32_t* counts, uint32_t* node_counter, const int* __restrict__ root_cluster)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { items[0] = CollapseItem{root_cluster[0], 0, 0u, 0u}; counts[0] = 1; *node_counter = 1; }
}

__global__ void k_gather_tris(uint32_t n, const uint32_t* __restrict__ sorted_ids, const float4* __restrict__ tri_rec,
                              float4* __restrict__ tri_sorted)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = sorted_ids[i];
    tri_sorted[3 * (size_t)i + 0] = tri_rec[3 * (size_t)id + 0];
    tri_sorted[3 * (size_t)i + 1] = tri_rec[3 * (size_t)id + 1];
    tri_sorted[3 * (size_t)i + 2] = tri_rec[3 * (size_t)id + 2];
}

// ------------------------------------------------------------------ scene assembly
__global__ void k_rebase_indices(const uint32_t* __restrict__ idx_local, const uint2* __restrict__ mesh_table, uint32_t n_meshes, uint32_t ntri,
                                 uint32_t* __restrict__ idx_global, uint32_t* __restrict__ tri_mesh)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntri; t += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n_meshes;                   // last mesh whose first triangle is <= t (empty meshes share a start: take the last)
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (mesh_table[mid].x <= t) lo = mid; else hi = mid; }
        const uint32_t vb = mesh_table[lo].y;
        idx_global[3 * (size_t)t] = idx_local[3 * (size_t)t] + vb;
        idx_global[3 * (size_t)t + 1] = idx_local[3 * (size_t)t + 1] + vb;
        idx_global[3 * (size_t)t + 2] = idx_local[3 * (size_t)t + 2] + vb;
        tri_mesh[t] = lo;
    }
}
void launch_rebase_indices(cudaStream_t st, const uint32_t* idx_local, const uint2* mesh_table, uint32_t n_meshes, uint32_t ntri,
                           uint32_t* idx_global, uint32_t* tri_mesh)
{
    k_rebase_indices<<<148 * 8, 256, 0, st>>>(idx_local, mesh_table, n_meshes, ntri, idx_global, tri_mesh);
}

// ------------------------------------------------------------------ host orchestration
// Every temporary of one build; freed when the build returns, on the error paths too.  Scratch arrays are carved out of ONE
// slab (a 10 M-triangle build used to spend 60 of its 70 ms in ~40 cudaMalloc / cudaFree calls, 12 ms in kernels); what the
// caller keeps (own = true) gets an allocation of its own.
struct Temps {
    std::vector<void*> ptrs;
    unsigned char* slab = nullptr; size_t slab_size = 0, slab_used = 0;
    ~Temps() { for (void* p : ptrs) cudaFree(p); cudaFree(slab); }
    void reserve(size_t bytes) { if (cudaMalloc((void**)&slab, bytes) == cudaSuccess) slab_size = bytes; else { slab = nullptr; cudaGetLastError(); } }
    template <typename T> cudaError_t alloc(T** p, size_t n, bool own = false)
    {
        const size_t bytes = (n * sizeof(T) + 32 + 255) & ~(size_t)255;
        if (!own && slab && slab_used + bytes <= slab_size) { *p = (T*)(slab + slab_used); slab_used += bytes; return cudaSuccess; }
        cudaError_t e = cudaMalloc((void**)p, n * sizeof(T) + 32);
        if (e == cudaSuccess) ptrs.push_back((void*)*p); else *p = nullptr;
        return e;
    }
    void release(const void* p) { for (auto& q : ptrs) if (q == p) { q = ptrs.back(); ptrs.pop_back(); return; } }   // the caller keeps it
};

// exclusive scan of data[0..n) in place; *total_out (device, may be null) receives the sum.  tile_sums: scratch of >= n / SC_TILE + 2 words
static void scan_u32(uint32_t* data, uint32_t n, uint32_t* total_out, uint32_t* tile_sums, cudaStream_t st, uint64_t* launches)
{
    if (n <= 64u * SC_TILE) { k_scan_u32<<<1, 1024, 0, st>>>(data, n, total_out); *launches += 1; return; }
    const uint32_t tiles = (n + SC_TILE - 1) / SC_TILE;
    k_scan_tile_sums<<<tiles, 256, 0, st>>>(data, n, tile_sums);
    k_scan_u32<<<1, 1024, 0, st>>>(tile_sums, tiles, total_out);
    k_scan_tiles<<<tiles, 256, 0, st>>>(data, n, tile_sums);
    *launches += 3;
}

cudaError_t build_bvh(const float* d_vtx, const uint32_t* d_idx, const uint32_t* d_tri_mesh, uint32_t ntri, int leaf_max, int wide, cudaStream_t st,
                      BuiltBvh* out, uint64_t* launches)
{
    const size_t nstride = 4;                          // uint4 per node
    out->wide = wide;
    out->nodes = nullptr; out->tris = nullptr; out->tri_rec = nullptr; out->sorted_ids = nullptr;
    out->n_nodes = 0; out->n_tris = ntri; out->max_depth = 0; out->balanced_fallback = 0;
    if (ntri == 0) return cudaSuccess;
    // FOVPT_BUILD_TRACE=1: host wall time per phase on stderr (phases end with a stream synchronise only when tracing)
    const bool trace = getenv("FOVPT_BUILD_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_last = now();
    auto phase = [&](const char* name) {
        if (!trace) return;
        cudaStreamSynchronize(st);
        const double t = now();
        fprintf(stderr, "[fovpt build] %-28s %9.3f ms\n", name, t - t_last);
        t_last = now();
    };
    Temps tmp;
    const int T = 256;
    const uint32_t gb = (ntri + T - 1) / T;
    float4 *tri_rec = nullptr, *box_lo = nullptr, *box_hi = nullptr, *nlo = nullptr, *nhi = nullptr;
    float4 *clo[2] = {nullptr, nullptr}, *chi[2] = {nullptr, nullptr};
    int* cid[2] = {nullptr, nullptr};
    int* bounds = nullptr;
    uint64_t *keys[2] = {nullptr, nullptr};
    uint32_t *vals[2] = {nullptr, nullptr};
    uint32_t *hist = nullptr, *scal = nullptr, *nn = nullptr, *alive = nullptr, *pos = nullptr, *cnt = nullptr, *sorted = nullptr;
    int *left = nullptr, *right = nullptr;
    CollapseItem* items[2] = {nullptr, nullptr};
    uint32_t* lvl_counts = nullptr;
    const uint32_t nblocks = (ntri + RS_TILE - 1) / RS_TILE;
    const uint32_t ninner = ntri > 1 ? ntri - 1 : 1;
    // scratch per triangle: boxes 32, keys 16, ids 8, clusters 72, nn/alive/pos 12, binary nodes 44, wide nodes 64, work items 32
    tmp.reserve((size_t)ntri * 288 + 256 * (size_t)nblocks * 4 + ((size_t)ntri / SC_TILE + 64) * 8 + (1u << 20));
    CK(tmp.alloc(&tri_rec, 3 * (size_t)ntri, true)); CK(tmp.alloc(&box_lo, ntri)); CK(tmp.alloc(&box_hi, ntri));
    CK(tmp.alloc(&bounds, 8)); CK(tmp.alloc(&keys[0], ntri)); CK(tmp.alloc(&keys[1], ntri));
    CK(tmp.alloc(&vals[0], ntri)); CK(tmp.alloc(&vals[1], ntri)); CK(tmp.alloc(&hist, 256 * (size_t)nblocks));
    CK(tmp.alloc(&scal, 8)); CK(tmp.alloc(&sorted, ntri, true));
    uint32_t* tile_sums = nullptr;
    CK(tmp.alloc(&tile_sums, std::max<size_t>(ntri, 256 * (size_t)nblocks) / SC_TILE + 2));
    CK(cudaMemsetAsync(scal, 0, 8 * sizeof(uint32_t), st));
    phase("allocate (sort stage)");

    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_tri_setup<<<gb, T, 0, st>>>(d_vtx, d_idx, d_tri_mesh, ntri, tri_rec, box_lo, box_hi, bounds);
    k_morton<<<gb, T, 0, st>>>(box_lo, box_hi, ntri, bounds, keys[0], vals[0]);
    *launches += 3;
    phase("triangle setup + morton");
    int cur = 0;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = pass * 8;
        k_radix_hist<<<nblocks, RS_THREADS, 0, st>>>(keys[cur], ntri, shift, hist, nblocks);
        scan_u32(hist, 256 * nblocks, nullptr, tile_sums, st, launches);
        k_radix_scatter<<<nblocks, RS_THREADS, 0, st>>>(keys[cur], vals[cur], keys[1 - cur], vals[1 - cur], ntri, shift, hist, nblocks);
        cur = 1 - cur;
        *launches += 2;
    }
    phase("radix sort");
    // scene pad needs the bounds on the host
    int hb[6];
    CK(cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float S = 0.f;
    for (int a = 0; a < 6; ++a) {
        int i = hb[a]; i = i >= 0 ? i : i ^ 0x7fffffff;
        float f; memcpy(&f, &i, 4);
        out->bounds[a] = f;
        S = fmaxf(S, fabsf(f));
    }
    const float pad = S * (1.0f / 65536.0f);
    out->pad = pad;

    float4* tri_sorted = nullptr; uint4* nodes = nullptr;
    CK(tmp.alloc(&tri_sorted, 3 * (size_t)ntri, true));
    if ((int)ntri <= leaf_max) {
        CK(tmp.alloc(&nodes, nstride, true));
        k_emit_single4<<<1, 32, 0, st>>>((int)ntri, bounds, pad, nodes);
        CK(cudaMemcpyAsync(sorted, vals[cur], ntri * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        *launches += 1;
        out->n_nodes = 1; out->max_depth = 1;
    } else {
        for (int k = 0; k < 2; ++k) { CK(tmp.alloc(&cid[k], ntri)); CK(tmp.alloc(&clo[k], ntri)); CK(tmp.alloc(&chi[k], ntri)); }
        CK(tmp.alloc(&nn, ntri)); CK(tmp.alloc(&alive, ntri)); CK(tmp.alloc(&pos, ntri));
        CK(tmp.alloc(&left, ninner)); CK(tmp.alloc(&right, ninner)); CK(tmp.alloc(&nlo, ninner)); CK(tmp.alloc(&nhi, ninner)); CK(tmp.alloc(&cnt, ninner));
        const uint32_t cap = ninner;                     // a wide node consumes at least one binary node
        const uint32_t max_levels = kMaxWideDepth + 1;
        CK(tmp.alloc(&nodes, nstride * (size_t)cap));
        CK(tmp.alloc(&items[0], cap)); CK(tmp.alloc(&items[1], cap));
        CK(tmp.alloc(&lvl_counts, (size_t)max_levels + 2));
        phase("allocate (tree stage)");
        // Attempt 0: the area-driven PLOC tree.  Attempt 1 (only if that tree is deeper than the traversal stack allows --
        // adversarial input such as long runs of nested boxes): forced pairing of Morton neighbours, a balanced tree of
        // depth ceil(log2 n) <= 29, which always fits.  OptiX accepts any input; so does this builder.
        for (int attempt = 0; attempt < 2; ++attempt) {
            // ---- PLOC: bottom-up binary tree over the Morton-ordered clusters
            CK(cudaMemsetAsync(scal, 0, 8 * sizeof(uint32_t), st));
            k_ploc_init<<<gb, T, 0, st>>>(ntri, vals[cur], box_lo, box_hi, cid[0], clo[0], chi[0]);
            *launches += 1;
            uint32_t n = ntri; int cb = 0; uint32_t rounds = 0;
            const uint32_t stop_at = (!attempt && kTopSahMax > 1 && ntri >= kTopSahMinTris) ? kTopSahMax : 1u;
            while (n > stop_at) {
                const uint32_t g = (n + kPlocThreads - 1) / kPlocThreads;
                k_ploc_nearest<<<g, kPlocThreads, 0, st>>>(n, clo[cb], chi[cb], nn, attempt);
                k_ploc_merge<<<g, kPlocThreads, 0, st>>>(n, nn, cid[cb], clo[cb], chi[cb], scal + 3, left, right, nlo, nhi, cnt, alive);
                CK(cudaMemcpyAsync(pos, alive, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
                scan_u32(pos, n, scal + 4, tile_sums, st, launches);
                k_ploc_compact<<<g, kPlocThreads, 0, st>>>(n, alive, pos, cid[cb], clo[cb], chi[cb], cid[1 - cb], clo[1 - cb], chi[1 - cb]);
                uint32_t n_next = 0;
                CK(cudaMemcpyAsync(&n_next, scal + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                *launches += 3;
                if (n_next == 0 || n_next >= n) return cudaErrorUnknown;      // every round merges at least the globally closest pair
                n = n_next; cb = 1 - cb; ++rounds;
            }
            out->ploc_rounds = rounds;
            phase("PLOC rounds");
            if (n > 1) try {                                 // the tree above the remaining clusters: top-down SAH on the host
                std::vector<int> h_cid(n); std::vector<float4> h_lo(n), h_hi(n); std::vector<uint32_t> h_cnt(n);
                uint32_t next_id = 0;
                k_cluster_counts<<<(n + 255) / 256, 256, 0, st>>>(n, cid[cb], cnt, pos);
                *launches += 1;
                CK(cudaMemcpyAsync(h_cid.data(), cid[cb], n * sizeof(int), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(h_lo.data(), clo[cb], n * sizeof(float4), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(h_hi.data(), chi[cb], n * sizeof(float4), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(h_cnt.data(), pos, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(&next_id, scal + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                std::vector<TopCluster> tc(n);
                for (uint32_t i = 0; i < n; ++i)
                    tc[i] = TopCluster{{h_lo[i].x, h_lo[i].y, h_lo[i].z}, h_cid[i], {h_hi[i].x, h_hi[i].y, h_hi[i].z}, h_cnt[i]};
                std::vector<TopNode> tn(n - 1);              // n - 1 new binary nodes, ids next_id ..
                const int root = top_sah(tc.data(), 0, n, (int)next_id, tn.data(), 0, 0);
                const size_t m = tn.size();
                std::vector<int> hl(m), hr(m); std::vector<float4> nl(m), nh(m); std::vector<uint32_t> nc(m);
                for (size_t i = 0; i < m; ++i) {
                    hl[i] = tn[i].left; hr[i] = tn[i].right; nc[i] = tn[i].cnt;
                    nl[i] = make_float4(tn[i].lo[0], tn[i].lo[1], tn[i].lo[2], 0.f); nh[i] = make_float4(tn[i].hi[0], tn[i].hi[1], tn[i].hi[2], 0.f);
                }
                CK(cudaMemcpyAsync(left + next_id, hl.data(), m * sizeof(int), cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(right + next_id, hr.data(), m * sizeof(int), cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(nlo + next_id, nl.data(), m * sizeof(float4), cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(nhi + next_id, nh.data(), m * sizeof(float4), cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(cnt + next_id, nc.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(cid[cb], &root, sizeof(int), cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));               // the host vectors go out of scope
                phase("top-level SAH (host)");
            } catch (const std::exception&) { return cudaErrorMemoryAllocation; }       // host memory for the cluster list
            // ---- 4-wide collapse, one launch per level of the wide tree, until a level comes up empty
            CK(cudaMemsetAsync(lvl_counts, 0, ((size_t)max_levels + 2) * sizeof(uint32_t), st));
            k_seed_collapse<<<1, 32, 0, st>>>(items[0], lvl_counts, scal + 2, cid[cb]);
            const int cgrid = (int)std::min<uint32_t>((cap + 127) / 128, 148u * 8u);
            for (uint32_t lvl = 0; lvl < max_levels; ++lvl) {
                k_collapse4<<<cgrid, 128, 0, st>>>(items[lvl & 1], items[(lvl + 1) & 1], lvl_counts, (int)lvl, scal + 2, left, right, cnt,
                                                  box_lo, box_hi, nlo, nhi, leaf_max, pad, nodes, sorted);
                *launches += 1;
            }
            std::vector<uint32_t> hc((size_t)max_levels + 2);
            uint32_t n4 = 0;
            CK(cudaMemcpyAsync(hc.data(), lvl_counts, hc.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&n4, scal + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            uint32_t depth4 = 0;
            while (depth4 < hc.size() && hc[depth4] > 0) ++depth4;
            out->n_nodes = n4; out->max_depth = depth4;      // depth4 > kMaxWideDepth: the tree was cut off at the last level
            phase("collapse");
            if (depth4 <= kMaxWideDepth) break;
            out->balanced_fallback = 1;
        }
        {                                                    // the node array the caller keeps, sized to fit
            uint4* exact = nullptr;
            CK(tmp.alloc(&exact, nstride * (size_t)out->n_nodes, true));
            CK(cudaMemcpyAsync(exact, nodes, nstride * (size_t)out->n_nodes * sizeof(uint4), cudaMemcpyDeviceToDevice, st));
            CK(cudaStreamSynchronize(st));
            nodes = exact;
        }
    }
    phase("shrink to fit");
    k_gather_tris<<<gb, T, 0, st>>>(ntri, sorted, tri_rec, tri_sorted);
    *launches += 1;
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    phase("gather triangles");
    out->nodes = nodes; out->tris = tri_sorted; out->tri_rec = tri_rec; out->sorted_ids = sorted;
    tmp.release(nodes); tmp.release(tri_sorted); tmp.release(tri_rec); tmp.release(sorted);
    return cudaSuccess;
}

void free_bvh(BuiltBvh* b)
{
    cudaFree((void*)b->nodes); cudaFree((void*)b->tris); cudaFree((void*)b->tri_rec); cudaFree((void*)b->sorted_ids);
    b->nodes = nullptr; b->tris = nullptr; b->tri_rec = nullptr; b->sorted_ids = nullptr; b->n_nodes = 0; b->n_tris = 0;
}

}  // namespace fovpt
