// bvh.cuh -- device BVH layout and the traversal state machine that replaces optixTrace
// (reference call sites: PT_sv5_/deviceProgram.cu:196-222 radiance / :224-248 occlusion).
//
// Layout in HBM (all 32-byte aligned, read through the read-only path with 256-bit loads):
//   nodes : 64 B per 4-wide inner node = 4 x uint4; child boxes are quantised to 8 bits per plane on
//           a per-node grid  plane(q) = p + q * 2^(e-127)  (snapped outward, so always conservative):
//             n0 = (bits(p.x), bits(p.y), bits(p.z), e.x | e.y << 8 | e.z << 16)
//             n1 = child[0..3]
//             n2 = (qlo.x[0..3], qhi.x[0..3], qlo.y[0..3], qhi.y[0..3])      one byte per child slot
//             n3 = (qlo.z[0..3], qhi.z[0..3], 0, 0)
//           child >= 0 : index of an inner node;  child < 0 : leaf,
//           ~child = (first_sorted_tri << 2) | (count - 1), count in 1..4;
//           an unused slot has the inverted byte box lo = 255, hi = 0 and can never be entered.
//           Two sectors per node instead of four: the traversal kernels are bound by L1 sector
//           throughput (profiles/), so bytes per visited node are what buys speed.
//   tris  : 48 B per triangle in leaf (Morton) order = 3 x float4
//             (v0.xyz, bits(global id)), (v1.xyz, bits(mesh id)), (v2.xyz, 0)      [the vertices, unmodified: the
//             watertight test of fovpt_exact.h needs bit-identical end points on both sides of a shared edge]
// Algorithmic bytes per ray are counted on the oracle's canonical binary tree:
// 64 per node visit + 48 per triangle test + 48 (ray in, hit out); see DESIGN.md section 5.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fovpt_exact.h"

// Bounds-checked build (make -C csrc checked -> libfovpt_checked.so; compute-sanitizer is not available on the B200 pool, so the
// memory-safety evidence is this: every index the kernels derive from device data is tested before use, a violation latches its
// site number in a device word that fovpt_sync reports, and the GPU test suite runs against that build through FOVPT_LIB).
// In the product build the macro expands to nothing.
#ifndef FOVPT_CHECKED
#define FOVPT_CHECKED 0
#endif
#if FOVPT_CHECKED
static __device__ unsigned int g_fovpt_check_site = 0;
#define FOVPT_CHECK(cond, site) do { if (!(cond)) atomicCAS(&g_fovpt_check_site, 0u, (unsigned)(site)); } while (0)
#else
#define FOVPT_CHECK(cond, site) do { } while (0)
#endif

namespace fovpt {

constexpr int kBvhWide = 4;   // branching factor of the device BVH (an 8-wide compressed layout was built and measured in round 1:
                              // 16 % slower on C3 -- octant-ordered visiting without pop-time culling costs more node visits
                              // than the wider nodes save, DESIGN.md section 6 -- and removed in round 2)

struct DeviceBvh {
    const uint4*  nodes = nullptr;
    const float4* tris = nullptr;
    uint32_t      n_nodes = 0;
    uint32_t      n_tris = 0;
};

constexpr int kStackSize = 96;          // entries; scene commit refuses trees deeper than kStackSize / 3
constexpr int kDone = 0x7fffffff;       // `cur` of a finished ray (never a valid node index)

struct HitRec { float t, u, v; uint32_t k; uint32_t gid; };   // k = sorted triangle slot, gid = global id

// Per-ray traversal stack: the youngest n_shared entries live in shared memory (entry-major, one column per thread:
// bank = thread, so lanes at different depths never conflict), deeper ones spill to the thread's local array.  Stack
// traffic is ~45 % of k_trace's L1 requests when it all goes through local memory (profiles/), and shared-memory
// accesses bypass the L1 tag stage.
#ifndef FOVPT_SMEM_STACK
#define FOVPT_SMEM_STACK 8
#endif
#ifndef FOVPT_BRANCHFREE_PUSH
#define FOVPT_BRANCHFREE_PUSH 0     // measured: -1 % alone, but +4 % once node steps repeat (FOVPT_NODE_REPEAT): the predicated stores lengthen the step
#endif
#ifndef FOVPT_RCP_APPROX
#define FOVPT_RCP_APPROX 1
#endif
struct RayStack {
    int2* local;          // kStackSize entries
    int2* shared;         // this thread's column: entry e at shared[e * stride]
    int   n_shared, stride;
    __device__ __forceinline__ void put(int i, int2 v) const
    {
        FOVPT_CHECK(i >= 0 && i - n_shared < kStackSize, 3);
        if (i < n_shared) shared[i * stride] = v; else local[i - n_shared] = v;
    }
    __device__ __forceinline__ int2 get(int i) const { return i < n_shared ? shared[i * stride] : local[i - n_shared]; }
    // store under a predicate, meant to compile to predicated STS / STL without a branch (i may be out of range when !p)
    __device__ __forceinline__ void put_if(bool p, int i, int2 v) const
    {
        const bool sh = i < n_shared;
        if (p && sh) shared[i * stride] = v;
        if (p && !sh) local[i - n_shared] = v;
    }
};
__device__ __forceinline__ RayStack local_stack(int2* local) { return RayStack{local, nullptr, 0, 0}; }

// Per-ray constants of the watertight triangle test (fx_rayspace of fovpt_exact.h): the nine row coefficients are written
// once per ray and read once per leaf step; k_trace keeps them in shared memory (coefficient-major, one column per thread:
// bank = thread), which costs the persistent warps no registers.  The three offsets stay in registers (they replace the
// ray origin, which the slab test does not need: it uses o * idir).
struct RayRows {
    float* p; int stride;
    __device__ __forceinline__ void put(const fx_rayspace& r) const
    {
        p[0] = r.ex.x; p[stride] = r.ex.y; p[2 * stride] = r.ex.z;
        p[3 * stride] = r.ey.x; p[4 * stride] = r.ey.y; p[5 * stride] = r.ey.z;
        p[6 * stride] = r.ez.x; p[7 * stride] = r.ez.y; p[8 * stride] = r.ez.z;
    }
    __device__ __forceinline__ fx_vec3 row(int k) const { fx_vec3 r; r.x = p[3 * k * stride]; r.y = p[(3 * k + 1) * stride]; r.z = p[(3 * k + 2) * stride]; return r; }
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

struct u8x { uint4 a, b; };
// one 256-bit read-only load (LDG.E.256.CONSTANT); p must be 32-byte aligned
__device__ __forceinline__ u8x ldg256(const uint4* p)
{
    u8x r;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t leaf_count(int child) { return (((uint32_t)~child) & 3u) + 1u; }
__device__ __forceinline__ uint32_t leaf_first(int child) { return ((uint32_t)~child) >> 2; }

__device__ __forceinline__ float safe_rcp(float d)
{
    // direction component 0 -> a huge finite reciprocal keeps 0*inf NaNs out of the slab test
    const float kTiny = 1e-30f;
    float ad = fabsf(d) < kTiny ? copysignf(kTiny, d) : d;
#if FOVPT_RCP_APPROX
    // MUFU.RCP alone (1 ulp) instead of the correctly rounded reciprocal (MUFU + Newton step + range fix-up, ~14 instructions,
    // three per ray).  idir only feeds the SLAB test, whose boxes are padded by 2^-16 of the scene extent (bvh_build.cu): an
    // error of 2^-23 relative moves an entry point by < 2^-20 of the extent, so the test stays conservative and hits -- decided
    // by the exact triangle test, which does not use idir -- are unchanged.
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ad));
    return r;
#else
    return __frcp_rn(ad);          // correctly rounded, like 1.0f / ad, without the division's slow path
#endif
}

// slab test of child slot J of a quantised node: near / far byte planes were already selected by the
// sign of the ray direction (so an inverted box yields an empty interval), a* = 2^e * idir and
// b* = (p - o) * idir fold the de-quantisation into one fma per plane.  Boxes are padded at build time
// (bvh_build.cu), which is what makes the fma form conservative with respect to fx_ray_tri.
#define FOVPT_BYTE(w, J) ((float)(((w) >> (8 * (J))) & 0xffu))
#define FOVPT_SLAB(J, tn, tf)                                                                                          \
    do {                                                                                                               \
        const float nx_ = fmaf(FOVPT_BYTE(nxw, J), ax, bx), ny_ = fmaf(FOVPT_BYTE(nyw, J), ay, by), nz_ = fmaf(FOVPT_BYTE(nzw, J), az, bz); \
        const float fx_ = fmaf(FOVPT_BYTE(fxw, J), ax, bx), fy_ = fmaf(FOVPT_BYTE(fyw, J), ay, by), fz_ = fmaf(FOVPT_BYTE(fzw, J), az, bz); \
        tn = fmaxf(fmaxf(nx_, ny_), fmaxf(nz_, tmin));                                                                 \
        tf = fminf(fminf(fx_, fy_), fminf(fz_, tlimit));                                                               \
    } while (0)

__device__ __forceinline__ void cswap(int& a, int& b) { const int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

// One ray's traversal as a resumable state machine.  A ray is always in one of three states:
//   at_node()  : `cur` is an inner node  -> node_step() tests its 4 child boxes (branch-free)
//   at_leaf()  : `cur` is a leaf         -> leaf_step() tests its 1..4 triangles
//   done()     : finished
// Kernels run 32 rays per warp, execute per iteration the step kind that the majority of lanes is
// waiting for, and replace finished rays from the stage queue (wavefront.cu k_trace).
// (Tried and measured slower on C3: parking a reached leaf and descending speculatively until a second
// leaf blocks the lane -- fuller leaf steps, but 9 % more node visits)
//   closest : min t in (tmin, tmax), ties -> smallest global triangle id (fx_hit_better), so the
//             result does not depend on tree shape or visiting order.
//   any     : true iff some FRONT-facing triangle intersects (tmin, tmax)
//             (OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES + payload-setting any-hit, reference
//             deviceProgram.cu:242,284-300); stops at the first one.
//   MODE = kClosest / kAny fix the kind at compile time; MODE = kMixed keeps it per ray in `any`, so
//   radiance rays of bounce k+1 and shadow rays of bounce k share one launch (and one warp).
// The stack lives in the caller's local array: closest-hit entries are (child, bits(t_near)) so that
// entries that fell behind the current best hit are dropped when popped, without touching the node.
constexpr int kClosest = 0, kAny = 1, kMixed = 2;

template <int MODE>
struct Traversal {
    float3 idir, ood;
    float rs_ox, rs_oy, rs_oz;     // offsets of the ray-space map (fovpt_exact.h); its rows live behind RayRows.  Origin and
                                   // direction themselves are not kept: the slab test needs idir and o * idir only.
    float tmin, tmax;
    float best_t, best_u, best_v;
    uint32_t best_gid, best_k;
    int cur, sp;
    bool any = (MODE == kAny);

    __device__ __forceinline__ bool is_any() const { return MODE == kMixed ? any : MODE == kAny; }

    __device__ __forceinline__ void begin(float3 o_, float3 d_, float tmin_, float tmax_, const RayRows& rows)
    {
        // tmax finite: unused child slots sit at +inf.  The caller guarantees tmin >= 0 (fovpt_set_config enforces it, the
        // query kernels clamp): every entry distance is then >= 0 and its bit pattern orders like the float (node_step keys).
        tmin = tmin_; tmax = fminf(tmax_, 3.0e38f);
        idir = make_float3(safe_rcp(d_.x), safe_rcp(d_.y), safe_rcp(d_.z));
        ood = make_float3(o_.x * idir.x, o_.y * idir.y, o_.z * idir.z);
        const fx_vec3 O = {o_.x, o_.y, o_.z}, D = {d_.x, d_.y, d_.z};
        const fx_rayspace rs = fx_rayspace_init(O, D);
        rows.put(rs);
        rs_ox = rs.ox; rs_oy = rs.oy; rs_oz = rs.oz;
        best_t = __int_as_float(0x7f800000); best_u = best_v = 0.f; best_gid = 0xffffffffu; best_k = 0;
        cur = 0; sp = 0;
    }
    __device__ __forceinline__ void reset() { cur = kDone; sp = 0; }
    __device__ __forceinline__ uint32_t leaf_work() const { return leaf_count(cur); }
    __device__ __forceinline__ bool found() const { return best_gid != 0xffffffffu; }
    __device__ __forceinline__ bool done() const { return cur == kDone; }
    __device__ __forceinline__ bool at_node() const { return cur >= 0 && cur != kDone; }
    __device__ __forceinline__ bool at_leaf() const { return cur < 0; }

    __device__ __forceinline__ void pop(const RayStack& stack)
    {
        if (MODE == kAny) { cur = sp > 0 ? stack.get(--sp).x : kDone; return; }
        // drop entries whose near distance is already beyond the best hit (strictly: equal t can still win on id)
        // the top entry is taken without a branch (9 pops in 10 end here); only a culled top entry enters the loop
        int2 e = make_int2(kDone, 0x7fffffff);
        if (sp > 0) e = stack.get(sp - 1);
        sp = max(sp - 1, 0);
        const bool ok = __int_as_float(e.y) <= best_t;        // bits 0x7fffffff are a NaN: never passes
        cur = ok ? e.x : kDone;
        if (!ok) {
            while (sp > 0) {
                const int2 f = stack.get(--sp);
                if (__int_as_float(f.y) <= best_t) { cur = f.x; break; }
            }
        }
    }

    // `unordered` (warp-uniform): every lane that takes this step is an any-hit ray, for which the visiting order does not
    // matter -- skip the distance keys and the sorting network
    __device__ __forceinline__ void node_step(const DeviceBvh& bvh, const RayStack& stack, bool unordered = false)
    {
        FOVPT_CHECK((uint32_t)cur < bvh.n_nodes, 1);
        const uint4* np = bvh.nodes + 4 * (size_t)cur;
        const u8x A = ldg256(np), B = ldg256(np + 2);           // A = (n0, n1), B = (n2, n3)
        const float tlimit = is_any() ? tmax : fminf(tmax, best_t);
        const float ax = __uint_as_float((A.a.w & 0xffu) << 23) * idir.x;
        const float ay = __uint_as_float(((A.a.w >> 8) & 0xffu) << 23) * idir.y;
        const float az = __uint_as_float(((A.a.w >> 16) & 0xffu) << 23) * idir.z;
        const float bx = fmaf(__uint_as_float(A.a.x), idir.x, -ood.x);
        const float by = fmaf(__uint_as_float(A.a.y), idir.y, -ood.y);
        const float bz = fmaf(__uint_as_float(A.a.z), idir.z, -ood.z);
        const bool px = idir.x >= 0.f, py = idir.y >= 0.f, pz = idir.z >= 0.f;
        const uint32_t nxw = px ? B.a.x : B.a.y, fxw = px ? B.a.y : B.a.x;
        const uint32_t nyw = py ? B.a.z : B.a.w, fyw = py ? B.a.w : B.a.z;
        const uint32_t nzw = pz ? B.b.x : B.b.y, fzw = pz ? B.b.y : B.b.x;
        float tn0, tn1, tn2, tn3, tf0, tf1, tf2, tf3;
        FOVPT_SLAB(0, tn0, tf0); FOVPT_SLAB(1, tn1, tf1); FOVPT_SLAB(2, tn2, tf2); FOVPT_SLAB(3, tn3, tf3);
        const int c0 = (int)A.b.x, c1 = (int)A.b.y, c2 = (int)A.b.z, c3 = (int)A.b.w;
        if (MODE == kAny || unordered) {
            // order is irrelevant for occlusion: enter the first hit child, push the others
            const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1, h2 = tn2 <= tf2, h3 = tn3 <= tf3;
#if FOVPT_BRANCHFREE_PUSH
            // enter the lowest hit slot, defer every other hit child (any order): predicated stores, no branches
            const bool p1 = h1 && h0, p2 = h2 && (h0 || h1), p3 = h3 && (h0 || h1 || h2);
            stack.put_if(p1, sp, make_int2(c1, 0));
            stack.put_if(p2, sp + (p1 ? 1 : 0), make_int2(c2, 0));
            stack.put_if(p3, sp + (p1 ? 1 : 0) + (p2 ? 1 : 0), make_int2(c3, 0));
            sp += (p1 ? 1 : 0) + (p2 ? 1 : 0) + (p3 ? 1 : 0);
            const int next = h0 ? c0 : h1 ? c1 : h2 ? c2 : h3 ? c3 : kDone;
#else
            int next = kDone;
            if (h3) { next = c3; }
            if (h2) { if (next != kDone) stack.put(sp++, make_int2(next, 0)); next = c2; }
            if (h1) { if (next != kDone) stack.put(sp++, make_int2(next, 0)); next = c1; }
            if (h0) { if (next != kDone) stack.put(sp++, make_int2(next, 0)); next = c0; }
#endif
            if (next != kDone) cur = next; else pop(stack);
        } else {
            // sort the (up to 4) hit children by entry distance with a 5-exchange network on integer keys:
            // key = bits(t_near) with the slot number in the two low mantissa bits (t_near >= tmin >= 0, so
            // the bit pattern orders like the float), misses = INT_MAX.
            const int kMiss = 0x7fffffff;
            int k0 = tn0 <= tf0 ? ((__float_as_int(tn0) & ~3) | 0) : kMiss;
            int k1 = tn1 <= tf1 ? ((__float_as_int(tn1) & ~3) | 1) : kMiss;
            int k2 = tn2 <= tf2 ? ((__float_as_int(tn2) & ~3) | 2) : kMiss;
            int k3 = tn3 <= tf3 ? ((__float_as_int(tn3) & ~3) | 3) : kMiss;
            cswap(k0, k1); cswap(k2, k3); cswap(k0, k2); cswap(k1, k3); cswap(k1, k2);
#define FOVPT_CHILD(k) (((k) & 2) ? (((k) & 1) ? c3 : c2) : (((k) & 1) ? c1 : c0))
            // far to near, so that the nearest of the deferred children is on top; clearing the key's low
            // bits only makes the stored distance a hair smaller, which keeps the pop-time cull conservative
#if FOVPT_BRANCHFREE_PUSH
            // misses sort to the end, so the hits are a prefix k0..k(n-1) and the deferred children k1..k(n-1) land, far to
            // near, in the slots below the new stack pointer: k1 at top-1, k2 at top-2, k3 at top-3 -- three predicated stores
            // instead of three divergent branches (each ran for 1-6 of ~24 lanes in nearly every step)
            const bool h1 = k1 != kMiss, h2 = k2 != kMiss, h3 = k3 != kMiss;
            const int top = sp + (h1 ? 1 : 0) + (h2 ? 1 : 0) + (h3 ? 1 : 0);
            stack.put_if(h3, top - 3, make_int2(FOVPT_CHILD(k3), k3 & ~3));
            stack.put_if(h2, top - 2, make_int2(FOVPT_CHILD(k2), k2 & ~3));
            stack.put_if(h1, top - 1, make_int2(FOVPT_CHILD(k1), k1 & ~3));
            sp = top;
#else
            if (k3 != kMiss) stack.put(sp++, make_int2(FOVPT_CHILD(k3), k3 & ~3));
            if (k2 != kMiss) stack.put(sp++, make_int2(FOVPT_CHILD(k2), k2 & ~3));
            if (k1 != kMiss) stack.put(sp++, make_int2(FOVPT_CHILD(k1), k1 & ~3));
#endif
            if (k0 != kMiss) cur = FOVPT_CHILD(k0); else pop(stack);
#undef FOVPT_CHILD
        }
    }

    // `nmax` = largest triangle count among the lanes of this warp that execute the step together
    // (keeps the loop trip count warp-uniform; each lane predicates on its own count)
    __device__ __forceinline__ void leaf_step(const DeviceBvh& bvh, const RayStack& stack, const RayRows& rows, uint32_t nmax)
    {
        const uint32_t first = leaf_first(cur), count = leaf_count(cur);
        FOVPT_CHECK(first + count <= bvh.n_tris, 2);
        bool occluded = false;
        for (uint32_t i = 0; i < nmax; ++i) {
            if (i < count) {
                const float4* tp = bvh.tris + 3 * (size_t)(first + i);
                const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
                fx_tri tr;
                tr.v0.x = a.x; tr.v0.y = a.y; tr.v0.z = a.z;
                tr.v1.x = b.x; tr.v1.y = b.y; tr.v1.z = b.z;
                tr.v2.x = c.x; tr.v2.y = c.y; tr.v2.z = c.z;
                // fx_ray_tri_wt of fovpt_exact.h, row by row: each row of the ray-space map is fetched when its coordinate is
                // computed, so only three coefficients are live at a time (all nine at once spilled at 56 registers)
                float t, u, v;
                fx_vec3 rw = rows.row(0);
                const float Ax = fx_rayspace_coord(tr.v0, rw, rs_ox), Bx = fx_rayspace_coord(tr.v1, rw, rs_ox), Cx = fx_rayspace_coord(tr.v2, rw, rs_ox);
                rw = rows.row(1);
                const float Ay = fx_rayspace_coord(tr.v0, rw, rs_oy), By = fx_rayspace_coord(tr.v1, rw, rs_oy), Cy = fx_rayspace_coord(tr.v2, rw, rs_oy);
                rw = rows.row(2);
                const float Az = fx_rayspace_coord(tr.v0, rw, rs_oz), Bz = fx_rayspace_coord(tr.v1, rw, rs_oz), Cz = fx_rayspace_coord(tr.v2, rw, rs_oz);
                const int side = fx_ray_tri_2d(Ax, Ay, Az, Bx, By, Bz, Cx, Cy, Cz, FX_ADD(FX_ADD(rw.x, rw.y), rw.z), &t, &u, &v);
                const bool in_range = t > tmin && t < tmax;
                if (MODE != kClosest) occluded = occluded || (side == 1 && in_range);
                if (MODE != kAny) {
                    const uint32_t gid = (uint32_t)__float_as_int(a.w);
                    if (!is_any() && side != 0 && in_range && fx_hit_better(t, gid, best_t, best_gid)) {
                        best_t = t; best_gid = gid; best_k = first + i; best_u = u; best_v = v;
                    }
                }
            }
        }
        if (MODE != kClosest && is_any() && occluded) { best_gid = 0; cur = kDone; return; }
        pop(stack);
    }

    // scalar driver (query kernels): one ray per thread, no warp cooperation
    __device__ __forceinline__ void run(const DeviceBvh& bvh, const RayStack& stack, const RayRows& rows)
    {
        while (!done()) {
            if (at_node()) node_step(bvh, stack);
            else leaf_step(bvh, stack, rows, leaf_count(cur));
        }
    }
};

#define FOVPT_TRAV(MODE) Traversal<MODE>

__device__ __forceinline__ bool trace_closest(const DeviceBvh& bvh, float3 o, float3 d, float tmin, float tmax, HitRec& out)
{
    if (bvh.n_tris == 0) return false;
    FOVPT_TRAV(kClosest) T;
    int2 stack[kStackSize];
    float rows_local[9];
    const RayRows rows{rows_local, 1};
    T.begin(o, d, tmin, tmax, rows);
    T.run(bvh, local_stack(stack), rows);
    if (!T.found()) return false;
    out.t = T.best_t; out.u = T.best_u; out.v = T.best_v; out.k = T.best_k; out.gid = T.best_gid;
    return true;
}

__device__ __forceinline__ bool trace_any_front(const DeviceBvh& bvh, float3 o, float3 d, float tmin, float tmax)
{
    if (bvh.n_tris == 0) return false;
    FOVPT_TRAV(kAny) T;
    int2 stack[kStackSize];
    float rows_local[9];
    const RayRows rows{rows_local, 1};
    T.begin(o, d, tmin, tmax, rows);
    T.run(bvh, local_stack(stack), rows);
    return T.found();
}

}  // namespace fovpt
