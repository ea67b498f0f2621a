// bvh_build.cu -- on-device BVH construction; replaces optixAccelBuild/optixAccelCompact
// (reference PT_sv5_/SimplePathtracer.cpp:602-746).  Pipeline, all kernels on the context stream:
//
//   k_tri_setup     per triangle: fx_make_tri record, AABB, scene bounds (atomic min/max)
//   k_morton        63-bit Morton code of the AABB centre (21 bits per axis)
//   radix sort      hand-written LSD sort, 8 x 8-bit passes over (u64 key, u32 triangle id):
//                   k_radix_hist -> k_scan_u32 -> k_radix_scatter (stable, match_any ranking)
//   k_hierarchy     Karras 2012: one thread per inner node finds its key range and split
//   k_refit         bottom-up AABBs, second arriver proceeds (atomic counter per node)
//   k_collapse4     breadth-first collapse of the binary tree into the 4-wide, 8-bit-quantised nodes of bvh.cuh:
//                   a wide node takes a binary node's two children and twice replaces the child
//                   of largest surface area by its own two children; subtrees with <= 4 triangles
//                   become leaves (their triangles are contiguous in Morton order).  One launch
//                   per level, work queue + node allocation through device counters.
//
// Boxes are padded by pad = max|coordinate| / 65536 so that the fma-form slab test of bvh.cuh can
// never reject a box whose triangle the exact test (fovpt_exact.h) accepts.  Hit results are
// independent of the tree (tie-break in fx_hit_better), so tree quality is purely a speed matter.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <vector>

#include "bvh.cuh"
#include "bvh_build.h"

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
__global__ void k_tri_setup(const float* __restrict__ vtx, const uint32_t* __restrict__ idx, uint32_t ntri,
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
        tri_rec[3 * (size_t)i + 1] = make_float4(t.e1.x, t.e1.y, t.e1.z, 0.f);
        tri_rec[3 * (size_t)i + 2] = make_float4(t.e2.x, t.e2.y, t.e2.z, 0.f);
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

// ------------------------------------------------------------------ Karras hierarchy
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

// child encoding during the build: >= 0 inner node, < 0 : ~leaf_position
__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                            int* __restrict__ parent_inner, int* __restrict__ parent_leaf,
                            int* __restrict__ rng_first, int* __restrict__ rng_last)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int lc = (first == gamma) ? ~gamma : gamma;
    const int rc = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc; right[i] = rc;
    rng_first[i] = first; rng_last[i] = last;
    if (lc >= 0) parent_inner[lc] = i; else parent_leaf[~lc] = i;
    if (rc >= 0) parent_inner[rc] = i; else parent_leaf[~rc] = i;
    if (i == 0) parent_inner[0] = -1;
}

__global__ void k_refit(int n, const uint32_t* __restrict__ sorted_ids, const float4* __restrict__ box_lo,
                        const float4* __restrict__ box_hi, const int* __restrict__ left, const int* __restrict__ right,
                        const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf,
                        float4* __restrict__ nlo, float4* __restrict__ nhi, uint32_t* __restrict__ visits,
                        uint32_t* __restrict__ max_depth)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent_leaf[i];
    {   // depth of this leaf (every leaf walks to the root once; bounds the traversal stack)
        uint32_t depth = 1;
        for (int p = node; p >= 0; p = parent_inner[p]) ++depth;
        atomicMax(max_depth, depth);
    }
    while (node >= 0) {
        const uint32_t prev = atomicAdd(&visits[node], 1u);
        if (prev == 0) return;              // first arriver stops; the second sees both children done
        __threadfence();
        float4 lo[2], hi[2];
        const int ch[2] = {left[node], right[node]};
        for (int s = 0; s < 2; ++s) {
            if (ch[s] < 0) { const uint32_t id = sorted_ids[~ch[s]]; lo[s] = box_lo[id]; hi[s] = box_hi[id]; }
            else { lo[s] = __ldcg(&nlo[ch[s]]); hi[s] = __ldcg(&nhi[ch[s]]); }
        }
        nlo[node] = make_float4(fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.f);
        nhi[node] = make_float4(fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.f);
        __threadfence();
        node = parent_inner[node];
    }
}

// ------------------------------------------------------------------ 4-wide collapse
struct CollapseItem { int bvh2; int node4; };

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

// one level: items_in[0 .. counts[level]) -> nodes4 + items_out[0 .. counts[level+1])
__global__ void k_collapse4(const CollapseItem* __restrict__ items_in, CollapseItem* __restrict__ items_out,
                            uint32_t* __restrict__ counts, int level, uint32_t* __restrict__ node_counter,
                            const int* __restrict__ left, const int* __restrict__ right,
                            const int* __restrict__ rng_first, const int* __restrict__ rng_last,
                            const uint32_t* __restrict__ sorted_ids, const float4* __restrict__ box_lo,
                            const float4* __restrict__ box_hi, const float4* __restrict__ nlo, const float4* __restrict__ nhi,
                            int leaf_max, float pad, uint4* __restrict__ nodes4)
{
    const uint32_t n = counts[level];
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x) {
        const CollapseItem item = items_in[it];
        int c[4]; int nc = 2;
        c[0] = left[item.bvh2]; c[1] = right[item.bvh2]; c[2] = c[3] = 0;
        for (int rep = 0; rep < 2; ++rep) {
            int pick = -1; float best = -1.f;
            for (int j = 0; j < nc; ++j) {
                if (c[j] < 0) continue;                                            // single triangle
                if (rng_last[c[j]] - rng_first[c[j]] + 1 <= leaf_max) continue;    // becomes a leaf as a whole
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
            inner[j] = j < nc && c[j] >= 0 && (rng_last[c[j]] - rng_first[c[j]] + 1 > leaf_max);
            n_inner += inner[j] ? 1 : 0;
        }
        uint32_t base = 0, qbase = 0;
        if (n_inner) { base = atomicAdd(node_counter, (uint32_t)n_inner); qbase = atomicAdd(&counts[level + 1], (uint32_t)n_inner); }
        float4 lo[4], hi[4]; int enc[4];
        int k = 0;
        for (int j = 0; j < 4; ++j) {
            if (j >= nc) {                                                         // unused slot (write_node4 gives it an inverted box)
                lo[j] = make_float4(0.f, 0.f, 0.f, 0.f); hi[j] = lo[j];
                enc[j] = 0x7fffffff;
            } else if (c[j] < 0) {
                const int pos = ~c[j];
                const uint32_t id = sorted_ids[pos];
                lo[j] = box_lo[id]; hi[j] = box_hi[id];
                enc[j] = ~(int)(((uint32_t)pos << 2) | 0u);
            } else {
                lo[j] = nlo[c[j]]; hi[j] = nhi[c[j]];
                if (inner[j]) {
                    enc[j] = (int)(base + k);
                    items_out[qbase + k] = CollapseItem{c[j], (int)(base + k)};
                    ++k;
                } else {
                    const int first = rng_first[c[j]], cnt = rng_last[c[j]] - first + 1;
                    enc[j] = ~(int)(((uint32_t)first << 2) | (uint32_t)(cnt - 1));
                }
            }
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

__global__ void k_seed_collapse(CollapseItem* items, uint32_t* counts, uint32_t* node_counter)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { items[0] = CollapseItem{0, 0}; counts[0] = 1; *node_counter = 1; }
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

// ------------------------------------------------------------------ host orchestration
template <typename T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T) + 32); }

cudaError_t build_bvh(const float* d_vtx, const uint32_t* d_idx, uint32_t ntri, int leaf_max, cudaStream_t st,
                      BuiltBvh* out, uint64_t* launches)
{
    out->nodes = nullptr; out->tris = nullptr; out->tri_rec = nullptr; out->sorted_ids = nullptr;
    out->n_nodes = 0; out->n_tris = ntri; out->max_depth = 0;
    if (ntri == 0) return cudaSuccess;
    const int T = 256;
    const uint32_t gb = (ntri + T - 1) / T;
    float4 *tri_rec = nullptr, *box_lo = nullptr, *box_hi = nullptr, *nlo = nullptr, *nhi = nullptr;
    int* bounds = nullptr;
    uint64_t *keys[2] = {nullptr, nullptr};
    uint32_t *vals[2] = {nullptr, nullptr};
    uint32_t *hist = nullptr, *visits = nullptr, *scal = nullptr;
    int *left = nullptr, *right = nullptr, *par_in = nullptr, *par_leaf = nullptr, *rf = nullptr, *rl = nullptr;
    const uint32_t nblocks = (ntri + RS_TILE - 1) / RS_TILE;
    const uint32_t ninner = ntri > 1 ? ntri - 1 : 1;
    CK(dalloc(&tri_rec, 3 * (size_t)ntri)); CK(dalloc(&box_lo, ntri)); CK(dalloc(&box_hi, ntri));
    CK(dalloc(&bounds, 8)); CK(dalloc(&keys[0], ntri)); CK(dalloc(&keys[1], ntri));
    CK(dalloc(&vals[0], ntri)); CK(dalloc(&vals[1], ntri)); CK(dalloc(&hist, 256 * (size_t)nblocks));
    CK(dalloc(&scal, 8));
    CK(cudaMemsetAsync(scal, 0, 8 * sizeof(uint32_t), st));

    k_init_bounds<<<1, 32, 0, st>>>(bounds);
    k_tri_setup<<<gb, T, 0, st>>>(d_vtx, d_idx, ntri, tri_rec, box_lo, box_hi, bounds);
    k_morton<<<gb, T, 0, st>>>(box_lo, box_hi, ntri, bounds, keys[0], vals[0]);
    *launches += 3;
    int cur = 0;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = pass * 8;
        k_radix_hist<<<nblocks, RS_THREADS, 0, st>>>(keys[cur], ntri, shift, hist, nblocks);
        k_scan_u32<<<1, 1024, 0, st>>>(hist, 256 * nblocks, nullptr);
        k_radix_scatter<<<nblocks, RS_THREADS, 0, st>>>(keys[cur], vals[cur], keys[1 - cur], vals[1 - cur], ntri, shift, hist, nblocks);
        cur = 1 - cur;
        *launches += 3;
    }
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
    CK(dalloc(&tri_sorted, 3 * (size_t)ntri));
    k_gather_tris<<<gb, T, 0, st>>>(ntri, vals[cur], tri_rec, tri_sorted);
    *launches += 1;

    CollapseItem* items[2] = {nullptr, nullptr};
    uint32_t* lvl_counts = nullptr;
    if ((int)ntri <= leaf_max) {
        CK(dalloc(&nodes, 4));
        k_emit_single4<<<1, 32, 0, st>>>((int)ntri, bounds, pad, nodes);
        *launches += 1;
        out->n_nodes = 1; out->max_depth = 1;
    } else {
        CK(dalloc(&left, ninner)); CK(dalloc(&right, ninner)); CK(dalloc(&par_in, ninner)); CK(dalloc(&par_leaf, ntri));
        CK(dalloc(&rf, ninner)); CK(dalloc(&rl, ninner)); CK(dalloc(&nlo, ninner)); CK(dalloc(&nhi, ninner));
        CK(dalloc(&visits, ninner));
        CK(cudaMemsetAsync(visits, 0, ninner * sizeof(uint32_t), st));
        const uint32_t gi = (ninner + T - 1) / T;
        k_hierarchy<<<gi, T, 0, st>>>(keys[cur], (int)ntri, left, right, par_in, par_leaf, rf, rl);
        k_refit<<<gb, T, 0, st>>>((int)ntri, vals[cur], box_lo, box_hi, left, right, par_in, par_leaf, nlo, nhi, visits, scal + 1);
        uint32_t depth2 = 0;
        CK(cudaMemcpyAsync(&depth2, scal + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        *launches += 2;
        // 4-wide collapse, one launch per level of the wide tree (at most the binary depth)
        const uint32_t cap = ninner;                     // a wide node consumes at least one binary node
        CK(dalloc(&nodes, 4 * (size_t)cap));
        CK(dalloc(&items[0], cap)); CK(dalloc(&items[1], cap));
        CK(dalloc(&lvl_counts, (size_t)depth2 + 2));
        CK(cudaMemsetAsync(lvl_counts, 0, ((size_t)depth2 + 2) * sizeof(uint32_t), st));
        k_seed_collapse<<<1, 32, 0, st>>>(items[0], lvl_counts, scal + 2);
        const int cgrid = (int)std::min<uint32_t>((cap + 127) / 128, 148u * 8u);
        for (uint32_t lvl = 0; lvl < depth2; ++lvl) {
            k_collapse4<<<cgrid, 128, 0, st>>>(items[lvl & 1], items[(lvl + 1) & 1], lvl_counts, (int)lvl, scal + 2, left, right, rf, rl,
                                              vals[cur], box_lo, box_hi, nlo, nhi, leaf_max, pad, nodes);
            *launches += 1;
        }
        std::vector<uint32_t> hc((size_t)depth2 + 2);
        uint32_t n4 = 0;
        CK(cudaMemcpyAsync(hc.data(), lvl_counts, hc.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&n4, scal + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        uint32_t depth4 = 0;
        while (depth4 < hc.size() && hc[depth4] > 0) ++depth4;
        out->n_nodes = n4; out->max_depth = depth4;
        if (n4 < cap) {                                  // shrink to fit
            uint4* exact = nullptr;
            CK(dalloc(&exact, 4 * (size_t)n4));
            CK(cudaMemcpyAsync(exact, nodes, 4 * (size_t)n4 * sizeof(uint4), cudaMemcpyDeviceToDevice, st));
            CK(cudaStreamSynchronize(st));
            cudaFree(nodes); nodes = exact;
        }
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    out->nodes = nodes; out->tris = tri_sorted; out->tri_rec = tri_rec; out->sorted_ids = vals[cur];
    cudaFree(box_lo); cudaFree(box_hi); cudaFree(bounds); cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(vals[1 - cur]);
    cudaFree(hist); cudaFree(scal); cudaFree(left); cudaFree(right); cudaFree(par_in); cudaFree(par_leaf);
    cudaFree(rf); cudaFree(rl); cudaFree(nlo); cudaFree(nhi); cudaFree(visits); cudaFree(items[0]); cudaFree(items[1]); cudaFree(lvl_counts);
    return cudaSuccess;
}

void free_bvh(BuiltBvh* b)
{
    cudaFree((void*)b->nodes); cudaFree((void*)b->tris); cudaFree((void*)b->tri_rec); cudaFree((void*)b->sorted_ids);
    b->nodes = nullptr; b->tris = nullptr; b->tri_rec = nullptr; b->sorted_ids = nullptr; b->n_nodes = 0; b->n_tris = 0;
}

}  // namespace fovpt
