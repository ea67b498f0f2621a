// wavefront.cu -- the sm_100a kernels that replace the reference's OptiX device programs
// (PT_sv5_/deviceProgram.cu): __raygen__renderFrame (:392-617) is split into k_raygen +
// k_resolve_li + k_resolve_px, optixTrace(RADIANCE) (:196-222) is k_extend, __closesthit__radiance
// / __miss__radiance (:619-732, :253-282) are k_shade, optixTrace(OCCLUSION) + its three programs
// (:224-300) are k_shadow.  All kernels are persistent grid-stride loops over device-resident
// queues, so a whole frame is enqueued without a host round trip.

#include "shade.cuh"
#include "wavefront.cuh"
#include "wavefront_launch.h"

namespace fovpt {

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ bool li_active(const FrameDev& F, const PassDev& P, unsigned lx, unsigned ly, uint32_t* px, uint32_t* py)
{
    fx_launch_to_pixel(lx, ly, P.factor, P.factor, P.uoff_x, P.uoff_y, px, py);       // deviceProgram.cu:433
    const float range = fx_gaze_range(*px, *py, F.cx, F.cy);                          // :435
    return fx_in_ring(range, P.r_inner, P.r_outer) != 0;                              // :437
}

__device__ __forceinline__ long long floor_div(long long a, long long b) { long long q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__device__ __forceinline__ long long ceil_div(long long a, long long b) { return -floor_div(-a, b); }

// launch coordinates (one axis) whose splat reaches destination coordinate X of an n-wide frame:
// px0 = l*factor + off, writes px0 .. px0+fill-1, clamped to [0, n-1] (deviceProgram.cu:546-556)
__device__ __forceinline__ void splat_candidates(int X, int n, const PassDev& P, int off, unsigned ldim, int* lo, int* hi)
{
    long long h = floor_div((long long)X - off, (long long)P.factor);
    long long l = ceil_div((long long)X - ((long long)P.fill - 1) - off, (long long)P.factor);
    if (X == n - 1) h = (long long)ldim - 1;     // everything beyond the frame clamps onto the border
    if (l < 0) l = 0;
    if (h > (long long)ldim - 1) h = (long long)ldim - 1;
    *lo = (int)l; *hi = (int)h;
}

// The launch index of pass P whose write to pixel (X,Y) survives: the row-major LAST active
// writer (sequential semantics of one launch; only border pixels ever have more than one).
__device__ __forceinline__ uint32_t pass_owner(const FrameDev& F, const PassDev& P, int X, int Y)
{
    int lox, hix, loy, hiy;
    splat_candidates(X, F.w, P, P.off_x, P.lw, &lox, &hix);
    splat_candidates(Y, F.h, P, P.off_y, P.lh, &loy, &hiy);
    for (int ly = hiy; ly >= loy; --ly)
        for (int lx = hix; lx >= lox; --lx) {
            uint32_t px, py;
            if (li_active(F, P, (unsigned)lx, (unsigned)ly, &px, &py)) return (uint32_t)ly * P.lw + (uint32_t)lx;
        }
    return kNoOwner;
}

__device__ __forceinline__ bool tile_is_mine(const FrameDev& F, int X, int Y)
{
    if (!F.tile_owner) return true;
    const int tx = X / F.tile_size, ty = Y / F.tile_size;
    FOVPT_CHECK(tx >= 0 && tx < F.tiles_x && ty >= 0 && ty < F.tiles_y, 6);
    return F.tile_owner[ty * F.tiles_x + tx] == (uint8_t)F.rank;
}

// ------------------------------------------------------------------ k_owner_map
// Each frame pixel is owned by exactly one (pass, launch index): the last pass that writes it
// (passes overdraw in order, SimplePathtracer.cpp:137-209) and within it the row-major last writer.
__global__ void k_owner_map(FrameDev F, uint32_t* __restrict__ owner_map, uint8_t* __restrict__ needed)
{
    const size_t npix = (size_t)F.w * F.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % F.w), Y = (int)(i / F.w);
        uint32_t own = kNoOwner;
        if (tile_is_mine(F, X, Y)) {
            for (int p = F.npass - 1; p >= 0; --p) {
                const uint32_t li = pass_owner(F, F.pass[p], X, Y);
                if (li != kNoOwner) { own = ((uint32_t)p << 28) | li; needed[F.pass[p].li_base + li] = 1; break; }
            }
        }
        owner_map[i] = own;
    }
}

// ------------------------------------------------------------------ camera
struct CamRay { float3 dir; uint32_t px, py; };

__device__ __forceinline__ float3 camera_dir(const FrameDev& F, uint32_t px, uint32_t py, float jx, float jy)
{
    const fx_vec3 U = {F.U[0], F.U[1], F.U[2]}, V = {F.V[0], F.V[1], F.V[2]}, W = {F.W[0], F.W[1], F.W[2]};
    const fx_vec3 d = fx_camera_dir(px, py, jx, jy, F.w, F.h, U, V, W);                 // :483-491
    return make_float3(d.x, d.y, d.z);
}

// ------------------------------------------------------------------ k_raygen
// which chunk of the table holds work item i (tables have <= kMaxChunks entries; first[] ascends)
__device__ __forceinline__ int chunk_of(const ChunkTable& T, uint32_t i)
{
    int c = 0;
#pragma unroll
    for (int j = 1; j < kMaxChunks; ++j) c += (j < T.n && i >= T.c[j].first) ? 1 : 0;
    return c;
}

// ONE launch for all passes (and all layers of a batch): one thread per (layer, launch index, sample).
__global__ void k_raygen(FrameDev F, ChunkTable T, const uint8_t* __restrict__ needed, PathState ps, Queues q,
                         float4* __restrict__ li_backplate, DevStats* __restrict__ stats)
{
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint32_t s_cnt[32];
    __shared__ uint32_t s_base;
    const uint32_t total = T.total;
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long started = 0;
    // queue slots are claimed with ONE atomicAdd per CTA and iteration (a warp-level claim put ~70 k same-address atomics
    // per frame on counts[0]); the trip count is CTA-uniform
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        unsigned k = 0, lx = 0, ly = 0, layer = 0; uint32_t li = 0, pid = 0; int pass = 0;
        bool active = i < total;                                   // (run_frame keeps total far enough below 2^32 for i not to wrap)
        if (active) {
            const ChunkDev ch = T.c[chunk_of(T, i)];
            pass = ch.pass;
            const PassDev& P = F.pass[pass];
            const uint32_t local = i - ch.first;
            const uint32_t per_layer = ch.rows * P.lw * P.spp;
            layer = ch.s0 + local / per_layer;
            const uint32_t r = local % per_layer;
            k = r % P.spp;
            const uint32_t lloc = r / P.spp;
            lx = lloc % P.lw; ly = ch.row0 + lloc / P.lw;
            li = ly * P.lw + lx;
            pid = ch.pid_base + local;
            active = needed[P.li_base + li] != 0;
        }
        const unsigned m = __ballot_sync(FULL, active);
        if (lane == 0) s_cnt[wib] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (unsigned w = 0; w < nw; ++w) tot += s_cnt[w];
            s_base = tot ? atomicAdd(&q.counts[0], tot) : 0u;
        }
        __syncthreads();
        if (active) {
            const PassDev& P = F.pass[pass];
            uint32_t slot = s_base + __popc(m & ((1u << lane) - 1u));
            for (unsigned w = 0; w < wib; ++w) slot += s_cnt[w];
            FOVPT_CHECK(slot < F.path_cap && pid < F.path_cap && P.li_base + li < F.li_total, 7);
            uint32_t px, py;
            fx_launch_to_pixel(lx, ly, P.factor, P.factor, P.uoff_x, P.uoff_y, &px, &py);
            uint32_t seed = fx_tea4(ly * (unsigned)F.w + lx, P.subframe + layer * P.sub_step);   // :411
            seed = fx_seed_after(seed, k);
            const fx_rng rng = fx_rng_init(seed);                                           // :464 (before the jitter draws)
            const float jx = fx_rnd(&seed);                                                 // :479, x first
            const float jy = fx_rnd(&seed);
            const float3 dir = camera_dir(F, px, py, jx, jy);
            q.cur.o[slot] = make_float4(F.eye[0], F.eye[1], F.eye[2], 0.f);
            q.cur.d[slot] = make_float4(dir.x, dir.y, dir.z, 0.f);
            q.cur.thr[slot] = make_float4(1.f, 1.f, 1.f, 1.f);                               // throughput 1, rayEta 1 (:467-468)
            q.cur.rng[slot] = make_uint2(rng.s1, rng.s2);
            q.cur.meta[slot] = make_uint2(pid, 0u);
            ps.direct[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
            ps.indirect[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
            ps.alpha[pid] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ps.feat_n) { ps.feat_n[pid] = make_float4(0.f, 0.f, 0.f, 1.f); ps.feat_a[pid] = make_float4(0.f, 0.f, 0.f, 0.f); }   // a miss adds zeros (:271-272)
            if (k == P.spp - 1) {                                                           // backplate of the LAST sample (:495)
                const float4 bp = probe_eval(F.probe, probe_dir_to_uv(dir));
                li_backplate[(size_t)layer * F.li_total + P.li_base + li] = bp;
            }
            ++started;
        }
        __syncthreads();                                   // s_cnt / s_base are rewritten by the next iteration
    }
    // warp-level count of started samples
    for (int o = 16; o > 0; o >>= 1) started += __shfl_xor_sync(FULL, started, o);
    if (lane == 0 && started) atomicAdd(&stats->samples, started);
}

// the tail of SampleLights / SampleShadow once the occlusion bit is known (deviceProgram.cu:323-341, :367-384)
__device__ __forceinline__ void shadow_resolve(const PathState& ps, const ShadowBuf& sb, uint32_t k, bool occluded)
{
    const float4 pe = sb.pend[k];
    const uint32_t pid = sb.pid[k];
    const uint32_t mode = (uint32_t)__float_as_int(pe.w);
    if ((mode & 1u) == 0u) {
        if (!occluded) {
            float4* dst = (mode & 2u) ? ps.direct : ps.indirect;
            float4 a = dst[pid];
            a.x += pe.x; a.y += pe.y; a.z += pe.z;
            dst[pid] = a;
        }
    } else if (occluded) {                                 // shadow catcher
        float4 a = ps.alpha[pid];
        a.x += pe.x; a.y += pe.y; a.z += pe.z;
        ps.alpha[pid] = a;
    }
}

// ------------------------------------------------------------------ k_trace  (optixTrace)
// Persistent warps with dynamic ray fetch: every warp pulls rays from the stage queues through one
// warp-aggregated atomicAdd and REPLACES terminated lanes while the others keep their traversal state
// in registers, so a few long rays do not hold 31 idle lanes hostage.  One launch serves BOTH ray
// kinds of a wavefront iteration: the shadow rays q.shadow[0..counts[shadow_slot]) of the previous
// shading pass (any front-facing hit, reference deviceProgram.cu:224-248, then the NEE accumulate of
// SampleLights :323-341), handed out first, then the radiance rays q.cur[0..counts[0]) (closest hit,
// :196-222) -- they are independent, and together they keep the machine filled through the late, thin bounces.
#ifndef FOVPT_FETCH_MIN
#define FOVPT_FETCH_MIN 8
#endif
#ifndef FOVPT_TRACE_BLOCKS
#define FOVPT_TRACE_BLOCKS 9
#endif
#ifndef FOVPT_NODE_REPEAT
#define FOVPT_NODE_REPEAT 2   // node steps per vote: lanes still at an inner node after a node step take the next one without a new vote
#endif
#ifndef FOVPT_NODE_ADAPT
#define FOVPT_NODE_ADAPT 0    // > 0: repeat only while this many eighths of the first step's lanes are still at inner nodes
#endif
#ifndef FOVPT_LEAF_NUM
#define FOVPT_LEAF_NUM 2      // run the leaf phase when LEAF_NUM * lanes_at_leaf > LEAF_DEN * lanes_at_node (leaves are cheap:
#define FOVPT_LEAF_DEN 1      // <= 2 triangles; measured best at 2:1 on C3)
#endif
constexpr int kFetchMin = FOVPT_FETCH_MIN;          // refill as soon as this many lanes are idle

// The persistent loop of k_trace for one warp.  A lane is idle exactly when its traversal is done(); `keep` is the number
// of busy lanes at or below which the warp stops stepping and refills first (0 once the queues are drained: then it runs
// until every lane is done).
template <bool THIN>
__device__ __forceinline__ void trace_loop(const SceneDev& S, const PathState& ps, const Queues& q, int shadow_slot, float tmin, float tmax,
                                           const RayStack& stack, const RayRows& rows, uint32_t n_any, uint32_t n, uint32_t cap_in, DevStats* stats)
{
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t cap = THIN ? cap_in : 32u;
    const unsigned lanes = THIN ? ((1u << cap) - 1u) : FULL;               // lanes that take rays in this launch
    const int fetch_min = THIN ? min(kFetchMin, (int)((cap + 3u) / 4u)) : kFetchMin;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t* fetch = q.counts + 3;
    FOVPT_TRAV(kMixed) T;
    uint32_t pid = 0;
    int keep = (int)cap - fetch_min;
    bool exhausted = false;
    T.reset();
#ifdef FOVPT_TRACE_STATS
    unsigned long long st_node_exec = 0, st_node_lanes = 0, st_leaf_exec = 0, st_leaf_lanes = 0, st_leaf_tris = 0, st_leaf_nmax = 0, st_fetch = 0;
#endif
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, T.done()) & lanes;
        if (idle == lanes || (!exhausted && __popc(idle) >= fetch_min)) {
            if (!exhausted) {
                const int want = __popc(idle);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(fetch, (uint32_t)want);
                base = __shfl_sync(FULL, base, 0);
                const uint32_t mine = base + __popc(idle & ((1u << lane) - 1u));
                if (((idle >> lane) & 1u) && mine < n) {
                    // shadow rays are handed out FIRST: while they last every warp is all-shadow and steps without sorting
                    // children, and the mixed phase that follows is short because shadow rays are short (measured on C3:
                    // k_trace -4.7 % against radiance rays first)
                    const bool is_any = mine < n_any;
                    pid = is_any ? mine : mine - n_any;                // slot in the shadow / radiance buffer
                    const float4 o = is_any ? q.shadow.o[pid] : q.cur.o[pid], d = is_any ? q.shadow.d[pid] : q.cur.d[pid];
                    T.begin(make_float3(o.x, o.y, o.z), make_float3(d.x, d.y, d.z), tmin, tmax, rows);
                    T.any = is_any;
                }
                if (base + (uint32_t)want >= n) { exhausted = true; keep = 0; }
            }
            if (__ballot_sync(FULL, !T.done()) == 0u) break;
        }
        for (;;) {
            // each iteration runs the step kind most lanes are waiting for; the others keep their state
            const bool node = T.at_node(), leaf = T.at_leaf();
            const unsigned mn = __ballot_sync(FULL, node), ml = __ballot_sync(FULL, leaf);
            if (__popc(mn | ml) <= keep) break;                          // enough idle lanes (refill first), or warp drained
            bool stepped;
            if (FOVPT_LEAF_NUM * __popc(ml) <= FOVPT_LEAF_DEN * __popc(mn)) {
#ifdef FOVPT_TRACE_STATS
                st_node_exec++; st_node_lanes += __popc(mn);
#endif
                // when every lane that takes the step is a shadow ray the children need no distance order (warp-uniform flag)
                const bool all_any = __ballot_sync(FULL, node && !T.any) == 0u;
                if (node) {
                    T.node_step(S.bvh, stack, all_any);
#if FOVPT_NODE_REPEAT > 1
                    // consecutive node steps only ever lose lanes to the leaf phase, so the vote between them decides nothing but
                    // when the leaf phase starts: take the next step(s) right away and save the loop bookkeeping (~19 instructions)
#if FOVPT_NODE_ADAPT
                    // keep stepping while at least ADAPT/8 of the lanes that took the first step are still at an inner node
                    const int first = __popc(mn);
#pragma unroll 1
                    for (int rep = 1; rep < FOVPT_NODE_REPEAT; ++rep) {
                        const bool again = T.at_node();
                        if (8 * __popc(__ballot_sync(__activemask(), again)) < FOVPT_NODE_ADAPT * first) break;
                        if (again) T.node_step(S.bvh, stack, all_any);
                    }
#else
#pragma unroll 1
                    for (int rep = 1; rep < FOVPT_NODE_REPEAT; ++rep) if (T.at_node()) T.node_step(S.bvh, stack, all_any);
#endif
#endif
                }
                stepped = node;
            } else {
                const uint32_t nmax = __reduce_max_sync(FULL, leaf ? T.leaf_work() : 0u);
#ifdef FOVPT_TRACE_STATS
                st_leaf_exec++; st_leaf_lanes += __popc(ml); st_leaf_nmax += nmax; st_leaf_tris += __reduce_add_sync(FULL, leaf ? T.leaf_work() : 0u);
#endif
                if (leaf) T.leaf_step(S.bvh, stack, rows, nmax);
                stepped = leaf;
            }
            if (stepped && T.done()) {
                if (!T.any) q.hit[pid] = T.found() ? make_float4(T.best_t, T.best_u, T.best_v, __int_as_float((int)T.best_k))
                                                  : make_float4(0.f, 0.f, 0.f, __int_as_float((int)kNoHit));
                else shadow_resolve(ps, q.shadow, pid, T.found());
            }
        }
#ifdef FOVPT_TRACE_STATS
        st_fetch++;
#endif
    }
#ifdef FOVPT_TRACE_STATS
    if (lane == 0) {
        unsigned long long* d = stats->dbg;
        atomicAdd(d + 0, st_node_exec); atomicAdd(d + 1, st_node_lanes); atomicAdd(d + 2, st_leaf_exec); atomicAdd(d + 3, st_leaf_lanes);
        atomicAdd(d + 4, st_leaf_tris); atomicAdd(d + 5, st_leaf_nmax); atomicAdd(d + 6, st_fetch);
        atomicAdd(d + 7, (unsigned long long)(blockIdx.x == 0 && threadIdx.x == 0 ? n : 0));
    }
#else
    (void)stats;
#endif
}

__global__ void __launch_bounds__(128, FOVPT_TRACE_BLOCKS) k_trace(SceneDev S, PathState ps, Queues q, int shadow_slot, float tmin, float tmax,
                                                                    DevStats* __restrict__ stats)
{
    const uint32_t n_closest = q.counts[0], n_any = q.counts[shadow_slot];
    const uint32_t n = n_closest + n_any;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&stats->radiance_rays, (unsigned long long)n_closest);
        atomicAdd(&stats->shadow_rays, (unsigned long long)n_any);
    }
    if (n == 0) return;
    if (S.bvh.n_tris == 0) {                                   // empty scene: every ray misses
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            if (i < n_closest) q.hit[i] = make_float4(0.f, 0.f, 0.f, __int_as_float((int)kNoHit));
            else shadow_resolve(ps, q.shadow, i - n_closest, false);
        }
        return;
    }
    // thin launch (late bounces, small per-GPU shares): spread the rays over ALL resident warps, `cap` lanes each.
    // A warp's time is its longest ray plus the serialisation of divergent lanes, so fewer rays per warp finish sooner
    // when there are SMs to spare (measured: 240 us -> 60 us floor per launch at 1/4 frame per GPU).
    const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t cap = min(32u, max(1u, (n + total_warps - 1u) / total_warps));
    int2 stack_local[kStackSize];
#if FOVPT_SMEM_STACK > 0
    __shared__ int2 stack_shared[FOVPT_SMEM_STACK * 128];
    const RayStack stack{stack_local, stack_shared + threadIdx.x, FOVPT_SMEM_STACK, 128};
#else
    const RayStack stack = local_stack(stack_local);
#endif
    __shared__ float rows_shared[9 * 128];
    const RayRows rows{rows_shared + threadIdx.x, 128};
    // the full-width case gets its own copy of the loop with cap / lanes / fetch_min folded into constants
    if (cap >= 32u) trace_loop<false>(S, ps, q, shadow_slot, tmin, tmax, stack, rows, n_any, n, 32u, stats);
    else trace_loop<true>(S, ps, q, shadow_slot, tmin, tmax, stack, rows, n_any, n, cap, stats);
}

// ------------------------------------------------------------------ k_shade  (closest-hit / miss program)
#ifndef FOVPT_SHADE_BLOCKS
#define FOVPT_SHADE_BLOCKS 7
#endif
// What one shaded path hands back: the continuation ray (if push_next) and the NEE shadow ray (if push_shadow); the
// caller appends them to the dense buffers of the next bounce once it knows their slots.
struct ShadeOut {
    bool   push_next = false, push_shadow = false;
    float4 o, d, thr;            // continuation: origin (also the shadow ray's origin), direction, throughput + rayEta
    uint2  rng, meta;
    float4 s_d, s_pend;          // shadow ray: direction, pending NEE value + mode bits
};

// closest-hit program of ONE path (reference __closesthit__radiance, deviceProgram.cu:619-732); `i` = its slot in q.cur
template <bool FAST>
__device__ __forceinline__ void shade_path(const FrameDev& F, const SceneDev& S, const PathState& ps, const Queues& q, uint32_t i, ShadeOut& out)
{
    // The kernel is bound by chains of dependent loads, not by arithmetic (removing the BSDF and probe code changes its time by
    // 10 %): everything indexed by the path id is requested together with the hit record, the material in one batch.
    const float4 h4 = q.hit[i];
    const float4 ro4 = q.cur.o[i], rd4 = q.cur.d[i], th4 = q.cur.thr[i];
    const uint2 r2 = q.cur.rng[i], meta = q.cur.meta[i];
    const uint32_t pid = meta.x;
    const uint32_t slot = (uint32_t)__float_as_int(h4.w);
    if (slot == kNoHit) return;                               // __miss__radiance: DONE, adds nothing (:253-282)
    const float3 ro = f3(ro4.x, ro4.y, ro4.z), rd = f3(rd4.x, rd4.y, rd4.z);
    FOVPT_CHECK(slot < S.bvh.n_tris && pid < F.path_cap, 5);
    const float4* tp = S.bvh.tris + 3 * (size_t)slot;
    const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
    const uint32_t gid = (uint32_t)__float_as_int(ta.w);
    FOVPT_CHECK(gid < S.bvh.n_tris && (uint32_t)__float_as_int(tb.w) < S.n_meshes, 5);
    // geometric normal, flipped toward the viewer (:632-634)
    const float3 v0 = f3(ta.x, ta.y, ta.z);
    const float3 n0 = unit3<FAST>(cross3(f3(tb.x, tb.y, tb.z) - v0, f3(tc.x, tc.y, tc.z) - v0));
    const float3 wo = -rd;
    const float3 N = n0 * copysignf(1.0f, dot3(wo, n0));
    const float3 P = ro + h4.x * rd;                         // :636-638
    const uint32_t mesh = (uint32_t)__float_as_int(tb.w);      // rides in the triangle record (bvh_build.cu k_tri_setup)
    const MaterialBlob* M = S.materials + mesh;
    // 104-byte Material as ten 8-byte loads issued back to back (the blob is 8-byte aligned, stride 104)
    const float2* M2 = reinterpret_cast<const float2*>(M);
    const float2 m0 = __ldg(M2 + 0), m1 = __ldg(M2 + 1), m2 = __ldg(M2 + 2), m4 = __ldg(M2 + 4), m5 = __ldg(M2 + 5), m6 = __ldg(M2 + 6),
                 m7 = __ldg(M2 + 7), m9 = __ldg(M2 + 9), m10 = __ldg(M2 + 10), m12 = __ldg(M2 + 12);
    const int2 mt = __ldg(&S.mesh_tex[mesh]);
    const int mflags = __float_as_int(m12.y);                // flags @100
    uint32_t st = meta.y;
    const uint32_t depth = st & kDepthMask;
    const bool secondary = (st & kFlagSecondary) != 0;
    const bool catcher = (mflags & kMaterialShadowCatcher) != 0;
    if (catcher && secondary) {                              // pass-through, --depth then ++depth (:646-651, :529)
        // a pass-through at depth 1 drops prd.depth to 0 with prd.normal / prd.albedo still holding the first hit, which the
        // raygen loop then adds once more (:509-512)
        if (ps.feat_n && depth == 1u) ps.feat_n[pid].w += 1.f;
        out.o = make_float4(P.x, P.y, P.z, 0.f); out.d = rd4; out.thr = th4; out.rng = r2; out.meta = meta;
        out.push_next = true;
        return;
    }
    BsdfParams bp;
    bp.color = f3(m1.y, m2.x, m2.y);                          // color @12,16,20
    bp.metallic = m5.x; bp.subsurface = m5.y; bp.specular = m6.x;      // @40,44,48
    bp.roughness = m6.y; bp.specularTint = m7.x; bp.clearcoat = m9.x;  // @52,56,72
    bp.clearcoatGloss = m9.y; bp.transmission = m10.x;        // @76,80
    float3 albedo = bp.color;
    if (mt.x >= 0 && mt.y) {                                 // :655-665
        const uint32_t* vi = S.tri_vidx + 3 * (size_t)gid;
        FOVPT_CHECK((uint32_t)mt.x < S.n_textures && vi[0] < S.n_vertices && vi[1] < S.n_vertices && vi[2] < S.n_vertices, 8);
        const float2 t0 = __ldg(&S.texcoords[vi[0]]), t1 = __ldg(&S.texcoords[vi[1]]), t2 = __ldg(&S.texcoords[vi[2]]);
        const float u = h4.y, v = h4.z, w0 = 1.f - u - v;
        const int4 ti = __ldg(&S.tex_info[mt.x]);
        albedo = texture_bilinear(S.texels + ti.x, ti.y, ti.z, w0 * t0.x + u * t1.x + v * t2.x, w0 * t0.y + u * t1.y + v * t2.y);
    }
    float3 T = f3(th4.x, th4.y, th4.z);
    float ray_eta = th4.w;
    float out_eta = 1.0f;                                    // :673-683
    if (ray_eta == 1.0f) {
        const float eta = m4.y;                               // eta @36
        // (resolving Material::ior() once per material at scene commit instead was measured: k_shade 6.63 -> 6.82 ms per C3 step)
        out_eta = eta == 0.f ? 2.f / (1.f - sqrtf(0.08f * bp.specular)) - 1.f : eta;
    }
    fx_rng rng; rng.s1 = r2.x; rng.s2 = r2.y;

    // Both random decisions of the hit first, in the reference's draw order: the probe sample of the next-event
    // estimate (:303-344 / :347-387), then the continuation direction (:700-707).  The BSDF is evaluated afterwards for both
    // directions at once, so its view-dependent terms (Fresnel at the view angle, Smith shadowing, tinted specular colour) are
    // computed once per hit instead of once per pdf / eval call.
    float3 wi, sky; float sky_pdf;
    probe_sample<FAST>(F.probe, F.guide, rng, wi, sky, sky_pdf);
    BsdfView view;
    view.ndv = dot3(N, wo); view.a = fmaxf(0.001f, bp.roughness);
    view.F = fresnel_dielectric(view.ndv, ray_eta, out_eta);
    float3 tu, tv, L = f3(0.f);
    tangent_frame<FAST>(N, tu, tv);
    float pdf = bsdf_sample_dir<FAST>(bp, view, ray_eta, out_eta, tu, tv, N, wo, rng, L);
    bsdf_view_terms<FAST>(bp, albedo, view);
    BsdfDir dl;
    bsdf_dir_terms<FAST>(view, N, wo, L, dl);
    if (pdf < 0.f) pdf = bsdf_pdf<FAST>(bp, view, dl);
    float3 pending = f3(0.f);
    {
        BsdfDir di;
        bsdf_dir_terms<FAST>(view, N, wo, wi, di);
        const float pdfb = bsdf_pdf<FAST>(bp, view, di);
        if (pdfb > 0.f) {
            const float wgt = 0.5f * sky_pdf / (0.5f * pdfb + 0.5f * sky_pdf);
            if (wgt > 0.f) {
                const float3 f = bsdf_eval<FAST>(bp, view, di, albedo, ray_eta, out_eta);
                pending = T * ((wgt * sky * f * fabsf(dot3(wi, N))) * Mth<FAST>::rcp(sky_pdf));
            }
        }
    }
    if (!catcher && !(st & kFlagAlphaOne)) { ps.alpha[pid] = make_float4(1.f, 1.f, 1.f, 0.f); st |= kFlagAlphaOne; }   // prd.alpha = 1 (:689), written once per path
    float3 E = f3(0.f);
    if (!secondary) E = f3(m0.x, m0.y, m1.x);                // emission @0,4,8 (:696-698)

    const bool done = !(pdf > 0.f);                          // :708-711
    const bool commit = !done && (int)depth < F.max_depth;   // radiance of this segment counts (:515-527)
    if (!done) {
        const float3 f = bsdf_eval<FAST>(bp, view, dl, albedo, ray_eta, out_eta);
        if (dot3(L, N) <= 0.f) ray_eta = out_eta;            // :717-721
        T = T * ((f * fabsf(dot3(N, L))) * Mth<FAST>::rcp(pdf));   // :724
    }
    const bool has_pending = pending.x != 0.f || pending.y != 0.f || pending.z != 0.f;
    const bool shadow = catcher ? has_pending : (commit && has_pending);
    // !secondary implies depth == 0 (depth only advances together with the secondary flag), so the
    // emission term lands in directLight exactly once; the NEE term is added by k_shadow.
    if (commit && depth == 0) ps.direct[pid] = make_float4(E.x, E.y, E.z, 0.f);
    out.o = make_float4(P.x, P.y, P.z, 0.f);
    if (shadow) {
        out.s_d = make_float4(wi.x, wi.y, wi.z, 0.f);
        const uint32_t mode = (catcher ? 1u : 0u) | (depth == 0 ? 2u : 0u);
        out.s_pend = make_float4(pending.x, pending.y, pending.z, __int_as_float((int)mode));
        out.push_shadow = true;
    }
    bool cont = commit;
    if (commit) {
        st = (st & ~kDepthMask) | ((depth + 1) & kDepthMask) | kFlagSecondary;
        // the (max_depth+1)-th segment only ever sets alpha; skip it when alpha is already 1
        if (F.skip_wasted && (int)(depth + 1) >= F.max_depth && (st & kFlagAlphaOne) && !S.any_catcher) cont = false;
    }
    out.meta = make_uint2(pid, st);                          // the shadow ray needs the pid even when the path ends here
    if (cont) {
        out.d = make_float4(L.x, L.y, L.z, 0.f);
        out.thr = make_float4(T.x, T.y, T.z, ray_eta);
        out.rng = make_uint2(rng.s1, rng.s2);
        out.push_next = true;
    }
}

// Denoiser features of the primary hits (fovpt_config.features): what __closesthit__radiance leaves in prd.normal /
// prd.albedo after the FIRST trace of a path (deviceProgram.cu:632-634, :653-665).  Runs between the first k_trace and the
// first k_shade, one thread per primary ray; a kernel of its own so that k_shade carries no extra registers for a
// feature that is off by default.  Misses keep the zeros k_raygen wrote (:271-272).
template <bool FAST>
__global__ void k_features(SceneDev S, PathState ps, Queues q)
{
    const uint32_t n = q.counts[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 h4 = q.hit[i];
        const uint32_t slot = (uint32_t)__float_as_int(h4.w);
        if (slot == kNoHit) continue;
        const uint32_t pid = q.cur.meta[i].x;
        const float4 rd4 = q.cur.d[i];
        const float4* tp = S.bvh.tris + 3 * (size_t)slot;
        const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
        const float3 v0 = f3(ta.x, ta.y, ta.z);
        const float3 n0 = unit3<FAST>(cross3(f3(tb.x, tb.y, tb.z) - v0, f3(tc.x, tc.y, tc.z) - v0));
        const float3 N = n0 * copysignf(1.0f, dot3(f3(-rd4.x, -rd4.y, -rd4.z), n0));
        const uint32_t gid = (uint32_t)__float_as_int(ta.w), mesh = (uint32_t)__float_as_int(tb.w);
        const float* M = reinterpret_cast<const float*>(S.materials + mesh);
        float3 albedo = f3(__ldg(M + 3), __ldg(M + 4), __ldg(M + 5));                 // color @12,16,20
        const int2 mt = __ldg(&S.mesh_tex[mesh]);
        if (mt.x >= 0 && mt.y) {
            const uint32_t* vi = S.tri_vidx + 3 * (size_t)gid;
            const float2 t0 = __ldg(&S.texcoords[vi[0]]), t1 = __ldg(&S.texcoords[vi[1]]), t2 = __ldg(&S.texcoords[vi[2]]);
            const float u = h4.y, v = h4.z, w0 = 1.f - u - v;
            const int4 ti = __ldg(&S.tex_info[mt.x]);
            albedo = texture_bilinear(S.texels + ti.x, ti.y, ti.z, w0 * t0.x + u * t1.x + v * t2.x, w0 * t0.y + u * t1.y + v * t2.y);
        }
        ps.feat_n[pid] = make_float4(N.x, N.y, N.z, 1.f);
        ps.feat_a[pid] = make_float4(albedo.x, albedo.y, albedo.z, 0.f);
    }
}

// One thread per queued path.  Queue appends are aggregated per CTA (ballot + shared-memory prefix, ONE atomicAdd per CTA
// and queue per iteration): with 7 M shaded paths per frame, warp-level aggregation still put ~440 k same-address
// atomics on the two queue counters, and waiting for them showed up as the top stall of the kernel.
template <bool FAST>
__global__ void __launch_bounds__(128, FOVPT_SHADE_BLOCKS) k_shade(FrameDev F, SceneDev S, PathState ps, Queues q, int shadow_slot)
{
    constexpr unsigned FULL = 0xffffffffu;
    // two copies of the append bookkeeping, used alternately: iteration i + 1 writes the other copy, and nobody reaches iteration
    // i + 2 before every thread has passed both barriers of iteration i + 1, i.e. finished reading copy i -- two barriers per
    // iteration instead of three
    __shared__ uint32_t s_cnt2[2][2][4];
    __shared__ uint32_t s_base2[2][2];
    int par = 0;
    const uint32_t n = q.counts[0];
    __syncthreads();      // every warp of the CTA has read counts[0] before thread 0 can take the end-of-iteration ticket below
                          // (a CTA with zero loop iterations would otherwise let late warps see the rotated counters)
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {      // CTA-uniform trip count
        const uint32_t i = base + threadIdx.x;
        ShadeOut out;
        if (i < n) shade_path<FAST>(F, S, ps, q, i, out);
        uint32_t (*s_cnt)[4] = s_cnt2[par];
        uint32_t* s_base = s_base2[par];
        par ^= 1;
        const unsigned mn = __ballot_sync(FULL, out.push_next), ms = __ballot_sync(FULL, out.push_shadow);
        if (lane == 0) { s_cnt[0][wib] = __popc(mn); s_cnt[1][wib] = __popc(ms); }
        __syncthreads();
        if (threadIdx.x < 2) {
            const uint32_t tot = s_cnt[threadIdx.x][0] + s_cnt[threadIdx.x][1] + s_cnt[threadIdx.x][2] + s_cnt[threadIdx.x][3];
            s_base[threadIdx.x] = tot ? atomicAdd(&q.counts[threadIdx.x == 0 ? 1 : shadow_slot], tot) : 0u;
        }
        __syncthreads();
        uint32_t on = s_base[0], os = s_base[1];
        for (unsigned w = 0; w < wib; ++w) { on += s_cnt[0][w]; os += s_cnt[1][w]; }
        const unsigned lt = (1u << lane) - 1u;
        if (out.push_next) {
            const uint32_t j = on + __popc(mn & lt);
            FOVPT_CHECK(j < F.path_cap, 9);
            q.next.o[j] = out.o; q.next.d[j] = out.d; q.next.thr[j] = out.thr; q.next.rng[j] = out.rng; q.next.meta[j] = out.meta;
        }
        if (out.push_shadow) {
            const uint32_t k = os + __popc(ms & lt);
            FOVPT_CHECK(k < F.path_cap, 9);
            q.shadow.o[k] = out.o; q.shadow.d[k] = out.s_d; q.shadow.pend[k] = out.s_pend; q.shadow.pid[k] = out.meta.x;
        }
    }
    // End of the wavefront iteration, done by the last CTA to finish instead of a launch of its own: next becomes cur,
    // the fetch cursor rewinds, and the shadow counter that this iteration's k_trace consumed is cleared for the shading
    // pass after next.  (Every CTA read counts[0] before it took its ticket, so nobody sees the new value.)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&q.counts[6], 1u) == gridDim.x - 1u) {
            __threadfence();
            volatile uint32_t* cnt = q.counts;
            cnt[0] = cnt[1]; cnt[1] = 0u; cnt[3] = 0u; cnt[shadow_slot == 2 ? 5 : 2] = 0u; cnt[6] = 0u;
        }
    }
}

// ------------------------------------------------------------------ k_resolve_li
// per (layer, launch index): sum its spp samples in order, apply the backplate/alpha composite (:536-560).  One launch for
// all passes and layers; work item i of chunk c = (layer - s0) * rows * lw + (ly - row0) * lw + lx.
__global__ void k_resolve_li(FrameDev F, ChunkTable T, const uint8_t* __restrict__ needed, PathState ps,
                             const float4* __restrict__ li_backplate, float4* __restrict__ li_color,
                             float4* __restrict__ li_normal, float4* __restrict__ li_albedo)
{
    const uint32_t total = T.total;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const ChunkDev ch = T.c[chunk_of(T, i)];
        const PassDev& P = F.pass[ch.pass];
        const uint32_t local = i - ch.first;
        const uint32_t per_layer = ch.rows * P.lw;
        const uint32_t layer = ch.s0 + local / per_layer;
        const uint32_t li = ch.row0 * P.lw + local % per_layer;
        if (!needed[P.li_base + li]) continue;
        float3 result = f3(0.f), alpha = f3(0.f);
        const uint32_t pid0 = ch.pid_base + local * P.spp;
        for (unsigned k = 0; k < P.spp; ++k) {
            const float4 d = ps.direct[pid0 + k], in = ps.indirect[pid0 + k], a = ps.alpha[pid0 + k];
            result = result + (f3(d.x, d.y, d.z) + f3(in.x, in.y, in.z));
            alpha = alpha + f3(a.x, a.y, a.z);
        }
        const float sppf = (float)P.spp;
        const float inv = 1.0f / sppf;
        alpha = alpha * inv;
        const size_t at = (size_t)layer * F.li_total + P.li_base + li;
        const float4 bp = li_backplate[at];
        const float3 color = (f3(bp.x, bp.y, bp.z) * sppf) * (f3(1.f) - alpha) + result;
        const float3 acc = color * inv;
        li_color[at] = make_float4(acc.x, acc.y, acc.z, 1.0f);
        if (li_normal) {                                     // :509-512, :541-542: sums in sample order, then / spp
            float3 nsum = f3(0.f), asum = f3(0.f);
            for (unsigned k = 0; k < P.spp; ++k) {
                const float4 n4 = ps.feat_n[pid0 + k], a4 = ps.feat_a[pid0 + k];
                for (int j = (int)n4.w; j > 0; --j) { nsum = nsum + f3(n4.x, n4.y, n4.z); asum = asum + f3(a4.x, a4.y, a4.z); }
            }
            nsum = nsum * inv; asum = asum * inv;
            li_normal[at] = make_float4(nsum.x, nsum.y, nsum.z, 1.0f);
            li_albedo[at] = make_float4(asum.x, asum.y, asum.z, 1.0f);
        }
    }
}

// ------------------------------------------------------------------ k_resolve_px
// per owned pixel: fetch its owner's colour, optional accumulation lerp, exposure + Reinhard + sRGB8 (:558-600)
// F.filter == 1: instead of the reference's box splat, a pixel of a sub-sampled ring (fill > 1) blends its own launch
// index with the three neighbours of the same pass on the side the pixel lies on, weights = distance from the block
// centre in block units (bilinear reconstruction; the reference's Filter.h is an empty class).  Neighbours that were
// not rendered (outside the ring / not owned by this rank) fall back to the pixel's own sample.
__device__ __forceinline__ float3 filtered_color(const FrameDev& F, const PassDev& P, uint32_t li, int X, int Y,
                                                 const float4* __restrict__ li_color, const uint8_t* __restrict__ needed)
{
    const float4 c = li_color[P.li_base + li];
    const float3 own = f3(c.x, c.y, c.z);
    if (!F.filter || P.fill <= 1u) return own;
    const int lx = (int)(li % P.lw), ly = (int)(li / P.lw);
    const float half = 0.5f * (float)(P.fill - 1u);
    const float dx = ((float)X - ((float)(lx * (int)P.factor + P.off_x) + half)) / (float)P.factor;
    const float dy = ((float)Y - ((float)(ly * (int)P.factor + P.off_y) + half)) / (float)P.factor;
    const int sx = dx < 0.f ? -1 : 1, sy = dy < 0.f ? -1 : 1;
    const float wx = fminf(fabsf(dx), 1.f), wy = fminf(fabsf(dy), 1.f);
    auto fetch = [&](int ax, int ay) {
        if (ax < 0 || ay < 0 || ax >= (int)P.lw || ay >= (int)P.lh) return own;
        const uint32_t j = (uint32_t)ay * P.lw + (uint32_t)ax;
        if (!needed[P.li_base + j]) return own;
        const float4 v = li_color[P.li_base + j];
        return f3(v.x, v.y, v.z);
    };
    const float3 c10 = fetch(lx + sx, ly), c01 = fetch(lx, ly + sy), c11 = fetch(lx + sx, ly + sy);
    return mix3(mix3(own, c10, wx), mix3(c01, c11, wx), wy);
}

// A batch (F.nsub > 1) replays, per pixel and in order, what nsub successive frames would have done to the accumulation
// buffer: the result is bit-identical to rendering them one by one; only the final state is tonemapped.
template <bool FAST>
__global__ void k_resolve_px(FrameDev F, const uint32_t* __restrict__ owner_map, const float4* __restrict__ li_color,
                             const uint8_t* __restrict__ needed, float4* __restrict__ accum_mirror,
                             const float4* __restrict__ li_normal, const float4* __restrict__ li_albedo)
{
    const size_t npix = (size_t)F.w * F.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t own = owner_map[i];
        if (own == kNoOwner) continue;
        FOVPT_CHECK((int)(own >> 28) < F.npass && (own & 0x0fffffffu) < F.pass[own >> 28].lw * F.pass[own >> 28].lh, 10);
        const PassDev& P = F.pass[own >> 28];
        const bool lerps = F.accumulate && !P.redraw;
        float3 acc = f3(0.f);
        bool have = false;                                       // acc holds the running value (else it is still in F.accum)
        for (int s = 0; s < F.nsub; ++s) {
            float3 cur = filtered_color(F, P, own & 0x0fffffffu, (int)(i % F.w), (int)(i / F.w), li_color + (size_t)s * F.li_total, needed);
            const unsigned sub = P.subframe + (unsigned)s * P.sub_step;
            if (lerps && sub > 0) {                              // PT_sv4_vmv2/deviceProgram.cu:545-553
                cur = f3(fmaxf(0.f, fminf(cur.x, 10.f)), fmaxf(0.f, fminf(cur.y, 10.f)), fmaxf(0.f, fminf(cur.z, 10.f)));
                const float a = 1.0f / (float)(sub + 1);
                if (!have) { const float4 prev = F.accum[i]; acc = f3(prev.x, prev.y, prev.z); }
                cur = mix3(acc, cur, a);
            }
            acc = cur; have = true;
        }
        const float4 out = make_float4(acc.x, acc.y, acc.z, 1.0f);
        F.accum[i] = out;
        if (accum_mirror) accum_mirror[i] = out;                 // multi-GPU: the owner rank's shared accumulation buffer (peer memory)
        F.frame[i] = tonemap_rgba8<FAST>(acc, F.exposure);
        if (F.feat_color) F.feat_color[i] = out;                 // :612-614: what the last frame of the batch leaves behind
        if (li_normal) {
            const size_t at = (size_t)(F.nsub - 1) * F.li_total + P.li_base + (own & 0x0fffffffu);
            if (F.feat_normal) F.feat_normal[i] = li_normal[at];
            if (F.feat_albedo) F.feat_albedo[i] = li_albedo[at];
        }
    }
}

// ------------------------------------------------------------------ k_probe_guides
// gy[k] = lower_bound(cdfY, k / kGuide); gx[row][k] = lower_bound(cdfX[row], k / kGuide), k = 0..kGuide
__global__ void k_probe_guides(ProbeBlob p, uint16_t* __restrict__ gy, uint16_t* __restrict__ gx)
{
    const int per = kGuide + 1;
    const size_t total = (size_t)(p.height + 1) * per;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / per), k = (int)(i % per);
        const float v = (float)k / (float)kGuide;
        if (r == p.height) gy[k] = (uint16_t)cdf_lower_bound(p.cdfValuesY, 0, p.height, v);
        else gx[(size_t)r * per + k] = (uint16_t)(cdf_lower_bound(p.cdfValuesX, r * p.width, (r + 1) * p.width, v) - r * p.width);
    }
}

// ------------------------------------------------------------------ debug / query kernels
__global__ void k_primary_hits(FrameDev F, SceneDev S, int32_t* __restrict__ tri, float* __restrict__ tuv, float* __restrict__ dir)
{
    const PassDev P = F.pass[0];
    const size_t total = (size_t)P.lw * P.lh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned lx = (unsigned)(i % P.lw), ly = (unsigned)(i / P.lw);
        uint32_t px, py;
        if (!li_active(F, P, lx, ly, &px, &py)) { tri[i] = -2; continue; }
        uint32_t seed = fx_tea4(ly * (unsigned)F.w + lx, P.subframe);
        const float jx = fx_rnd(&seed), jy = fx_rnd(&seed);
        const float3 d = camera_dir(F, px, py, jx, jy);
        HitRec h;
        const bool ok = trace_closest(S.bvh, make_float3(F.eye[0], F.eye[1], F.eye[2]), d, F.tmin, F.tmax, h);
        tri[i] = ok ? (int32_t)h.gid : -1;
        tuv[3 * i] = ok ? h.t : 0.f; tuv[3 * i + 1] = ok ? h.u : 0.f; tuv[3 * i + 2] = ok ? h.v : 0.f;
        dir[3 * i] = d.x; dir[3 * i + 1] = d.y; dir[3 * i + 2] = d.z;
    }
}

__global__ void k_trace_query(SceneDev S, const float* __restrict__ rays, size_t n, int mode, int32_t* __restrict__ tri, float* __restrict__ tuv)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float* r = rays + 8 * i;
        const float3 o = make_float3(r[0], r[1], r[2]), d = make_float3(r[4], r[5], r[6]);
        if (mode == 0) {
            HitRec h;
            const bool ok = trace_closest(S.bvh, o, d, fmaxf(r[3], 0.f), r[7], h);
            tri[i] = ok ? (int32_t)h.gid : -1;
            tuv[3 * i] = ok ? h.t : 0.f; tuv[3 * i + 1] = ok ? h.u : 0.f; tuv[3 * i + 2] = ok ? h.v : 0.f;
        } else if (mode == 1) {
            tri[i] = trace_any_front(S.bvh, o, d, fmaxf(r[3], 0.f), r[7]) ? 1 : 0;
        } else {                                       // modes 2 / 3: closest / any with visit counters in tuv
            float nodes = 0.f, leaves = 0.f, tris = 0.f;
            int2 stack_local[kStackSize];
            float rows_local[9];
            const RayStack stack = local_stack(stack_local);
            const RayRows rows{rows_local, 1};
            if (S.bvh.n_tris == 0) { tri[i] = mode == 2 ? -1 : 0; }
            else if (mode == 2) {
                FOVPT_TRAV(kClosest) T; T.begin(o, d, fmaxf(r[3], 0.f), r[7], rows);
                while (!T.done()) {
                    if (T.at_node()) { nodes += 1.f; T.node_step(S.bvh, stack); }
                    else { leaves += 1.f; tris += (float)T.leaf_work(); T.leaf_step(S.bvh, stack, rows, T.leaf_work()); }
                }
                tri[i] = T.found() ? (int32_t)T.best_gid : -1;
            } else {
                FOVPT_TRAV(kAny) T; T.begin(o, d, fmaxf(r[3], 0.f), r[7], rows);
                while (!T.done()) {
                    if (T.at_node()) { nodes += 1.f; T.node_step(S.bvh, stack); }
                    else { leaves += 1.f; tris += (float)T.leaf_work(); T.leaf_step(S.bvh, stack, rows, T.leaf_work()); }
                }
                tri[i] = T.found() ? 1 : 0;
            }
            tuv[3 * i] = nodes; tuv[3 * i + 1] = leaves; tuv[3 * i + 2] = tris;
        }
    }
}

// ------------------------------------------------------------------ texture-unit cross-check (test hook)
// hw: the texture unit with the reference's sampler state (SimplePathtracer.cpp:775-789); sw: candidate software definitions
//   mode 0 float weights | 1 frac rounded to 8 bits | 2 coordinate rounded to 1/256 texel (rni) | 3 frac truncated | 4 coordinate floored
__device__ __forceinline__ void tex_axis(float x, int n, int mode, int* i0, int* i1, float* a)
{
    const float xs = x * (float)n;
    long long i; float al;
    if (mode == 2 || mode == 4) {
        const float f = xs * 256.0f;
        const long long fx = (mode == 2 ? (long long)rintf(f) : (long long)floorf(f)) - 128;
        i = fx >> 8; al = (float)(fx & 255) * (1.0f / 256.0f);
    } else {
        const float xb = xs - 0.5f, fl = floorf(xb);
        i = (long long)fl; al = xb - fl;
        if (mode == 1) al = floorf(al * 256.0f + 0.5f) * (1.0f / 256.0f);
        if (mode == 3) al = floorf(al * 256.0f) * (1.0f / 256.0f);
    }
    long long m = i % n; if (m < 0) m += n;
    *i0 = (int)m; *i1 = (int)(m + 1 == n ? 0 : m + 1); *a = al;
}
__global__ void k_texture_probe(cudaTextureObject_t tex, const uint32_t* __restrict__ texels, int w, int h, const float2* __restrict__ uv, size_t n,
                                int mode, float4* __restrict__ sw, float4* __restrict__ hw)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 t = uv[i];
        hw[i] = tex2D<float4>(tex, t.x, t.y);
        if (mode < 0) { float o4[4]; fx_tex_bilinear_rgba8(texels, w, h, t.x, t.y, o4); sw[i] = make_float4(o4[0], o4[1], o4[2], o4[3]); continue; }   // the product's fetch (shade.cuh texture_bilinear)
        int x0, x1, y0, y1; float a, b;
        tex_axis(t.x, w, mode, &x0, &x1, &a); tex_axis(t.y, h, mode, &y0, &y1, &b);
        const uint32_t p00 = texels[(size_t)y0 * w + x0], p10 = texels[(size_t)y0 * w + x1], p01 = texels[(size_t)y1 * w + x0], p11 = texels[(size_t)y1 * w + x1];
        const float w00 = (1.f - a) * (1.f - b), w10 = a * (1.f - b), w01 = (1.f - a) * b, w11 = a * b;
        float o[4];
        for (int c = 0; c < 4; ++c) {
            const float k = 1.0f / 255.0f;
            o[c] = w00 * ((float)((p00 >> (8 * c)) & 255u) * k) + w10 * ((float)((p10 >> (8 * c)) & 255u) * k) +
                   w01 * ((float)((p01 >> (8 * c)) & 255u) * k) + w11 * ((float)((p11 >> (8 * c)) & 255u) * k);
        }
        sw[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
void launch_texture_probe(const LaunchCfg& L, cudaTextureObject_t tex, const uint32_t* texels, int w, int h, const float2* uv, size_t n, int mode,
                          float4* sw, float4* hw)
{
    k_texture_probe<<<L.grid_light, 256, 0, L.stream>>>(tex, texels, w, h, uv, n, mode, sw, hw);
}

// ------------------------------------------------------------------ peer-frame flags (multi-GPU gather over NVLink peer memory)
// The shared frame lives in the owner's memory; contributors store their tiles into it from k_resolve_px.  Ordering between
// the processes is carried by 32-bit sequence numbers in the same allocation, written and polled at system scope: a
// kernel boundary on the writer's stream precedes the signal, so every pixel store of the frame is performed before the
// flag becomes visible.  Sequence numbers compare modulo 2^32.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__global__ void k_peer_signal(uint32_t* flag, uint32_t seq)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
    }
}
// thread i waits until flags[i] has reached seq (skip = the index that is not waited for, -1: none); a peer that does not
// arrive within `timeout` clocks raises *err instead of hanging the stream
__global__ void k_peer_wait(const uint32_t* flags, int count, int skip, uint32_t seq, uint32_t* err, long long timeout)
{
    const int i = (int)threadIdx.x;
    if (i >= count || i == skip) return;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + i) - seq) < 0) {
        if (clock64() - t0 > timeout) { atomicExch(err, 1u + (uint32_t)i); return; }
        __nanosleep(256);
    }
}

// ------------------------------------------------------------------ launch wrappers (host)
void launch_owner_map(const LaunchCfg& L, const FrameDev& F, uint32_t* owner_map, uint8_t* needed)
{
    k_owner_map<<<L.grid_light, 256, 0, L.stream>>>(F, owner_map, needed);
}
void launch_raygen(const LaunchCfg& L, const FrameDev& F, const ChunkTable& T, const uint8_t* needed, const PathState& ps, const Queues& q,
                   float4* li_backplate, DevStats* stats)
{
    k_raygen<<<L.grid_light, 256, 0, L.stream>>>(F, T, needed, ps, q, li_backplate, stats);
}
void launch_trace(const LaunchCfg& L, const SceneDev& S, const PathState& ps, const Queues& q, int shadow_slot, float tmin, float tmax, DevStats* stats)
{
    k_trace<<<L.grid_trace, 128, 0, L.stream>>>(S, ps, q, shadow_slot, tmin, tmax, stats);
}
void launch_shade(const LaunchCfg& L, const FrameDev& F, const SceneDev& S, const PathState& ps, const Queues& q, int shadow_slot, bool fast)
{
    if (fast) k_shade<true><<<L.grid_shade, 128, 0, L.stream>>>(F, S, ps, q, shadow_slot);
    else k_shade<false><<<L.grid_shade, 128, 0, L.stream>>>(F, S, ps, q, shadow_slot);
}
void launch_features(const LaunchCfg& L, const SceneDev& S, const PathState& ps, const Queues& q, bool fast)
{
    if (fast) k_features<true><<<L.grid_light, 256, 0, L.stream>>>(S, ps, q);
    else k_features<false><<<L.grid_light, 256, 0, L.stream>>>(S, ps, q);
}
void launch_peer_signal(const LaunchCfg& L, uint32_t* flag, uint32_t seq) { k_peer_signal<<<1, 32, 0, L.stream>>>(flag, seq); }
void launch_peer_wait(const LaunchCfg& L, const uint32_t* flags, int count, int skip, uint32_t seq, uint32_t* err, long long timeout_clocks)
{
    k_peer_wait<<<1, 64, 0, L.stream>>>(flags, count, skip, seq, err, timeout_clocks);
}
void launch_resolve_li(const LaunchCfg& L, const FrameDev& F, const ChunkTable& T, const uint8_t* needed, const PathState& ps,
                       const float4* li_backplate, float4* li_color, float4* li_normal, float4* li_albedo)
{
    k_resolve_li<<<L.grid_light, 256, 0, L.stream>>>(F, T, needed, ps, li_backplate, li_color, li_normal, li_albedo);
}
void launch_resolve_px(const LaunchCfg& L, const FrameDev& F, const uint32_t* owner_map, const float4* li_color, const uint8_t* needed,
                       float4* accum_mirror, const float4* li_normal, const float4* li_albedo, bool fast)
{
    if (fast) k_resolve_px<true><<<L.grid_light, 256, 0, L.stream>>>(F, owner_map, li_color, needed, accum_mirror, li_normal, li_albedo);
    else k_resolve_px<false><<<L.grid_light, 256, 0, L.stream>>>(F, owner_map, li_color, needed, accum_mirror, li_normal, li_albedo);
}
void launch_probe_guides(const LaunchCfg& L, const ProbeBlob& p, uint16_t* gy, uint16_t* gx)
{
    k_probe_guides<<<L.grid_light, 256, 0, L.stream>>>(p, gy, gx);
}
void launch_primary_hits(const LaunchCfg& L, const FrameDev& F, const SceneDev& S, int32_t* tri, float* tuv, float* dir)
{
    k_primary_hits<<<L.grid_trace, 128, 0, L.stream>>>(F, S, tri, tuv, dir);
}
void launch_trace_query(const LaunchCfg& L, const SceneDev& S, const float* rays, size_t n, int mode, int32_t* tri, float* tuv)
{
    k_trace_query<<<L.grid_trace, 128, 0, L.stream>>>(S, rays, n, mode, tri, tuv);
}

// bounds-checked build: the site number of the first failed check since the last call (0 = none); always 0 in the product build
unsigned read_check_failure()
{
#if FOVPT_CHECKED
    unsigned v = 0, zero = 0;
    cudaMemcpyFromSymbol(&v, g_fovpt_check_site, sizeof(v));
    if (v) cudaMemcpyToSymbol(g_fovpt_check_site, &zero, sizeof(zero));
    return v;
#else
    return 0u;
#endif
}

void query_occupancy(int* extend_blocks_per_sm, int* shade_blocks_per_sm)
{
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(extend_blocks_per_sm, k_trace, 128, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(shade_blocks_per_sm, k_shade<false>, 128, 0);
}

}  // namespace fovpt
