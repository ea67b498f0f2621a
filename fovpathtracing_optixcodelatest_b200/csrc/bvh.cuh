// bvh.cuh -- device BVH layout and the two traversal routines that replace optixTrace
// (reference call sites: PT_sv5_/deviceProgram.cu:196-222 radiance / :224-248 occlusion).
//
// Layout in HBM (all 32-byte aligned, read through the read-only path):
//   nodes : 64 B per inner node = 4 x float4
//             n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//             n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//             n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//             n3 = (child0, child1, -, -) as int bits; child >= 0: inner node index,
//                  child < 0: leaf, ~child = (first_sorted_tri << 2) | (count - 1), count in 1..4
//   tris  : 48 B per triangle in leaf (Morton) order = 3 x float4
//             (v0.xyz, bits(global id)), (e1.xyz, 0), (e2.xyz, 0)      [fx_make_tri of fovpt_exact.h]
// Algorithmic bytes per ray: 64 per node visit + 48 per triangle test + 48 (ray in, hit out).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fovpt_exact.h"

namespace fovpt {

struct DeviceBvh {
    const float4* nodes = nullptr;
    const float4* tris = nullptr;
    uint32_t      n_nodes = 0;
    uint32_t      n_tris = 0;
};

constexpr int kStackSize = 96;

struct HitRec { float t, u, v; uint32_t k; uint32_t gid; };   // k = sorted triangle slot, gid = global id

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// slab test of one child box against the ray (origin folded into ood = o * idir).  Boxes are
// padded at build time (bvh_build.cu), which is what makes the fma form conservative.
__device__ __forceinline__ void slab2(const float4 n0, const float4 n1, const float4 n2, const float3 idir, const float3 ood,
                                      float tmin, float tmax, float& tn0, float& tf0, float& tn1, float& tf1)
{
    float c0lox = fmaf(n0.x, idir.x, -ood.x), c0hix = fmaf(n0.y, idir.x, -ood.x);
    float c0loy = fmaf(n0.z, idir.y, -ood.y), c0hiy = fmaf(n0.w, idir.y, -ood.y);
    float c0loz = fmaf(n2.x, idir.z, -ood.z), c0hiz = fmaf(n2.y, idir.z, -ood.z);
    float c1lox = fmaf(n1.x, idir.x, -ood.x), c1hix = fmaf(n1.y, idir.x, -ood.x);
    float c1loy = fmaf(n1.z, idir.y, -ood.y), c1hiy = fmaf(n1.w, idir.y, -ood.y);
    float c1loz = fmaf(n2.z, idir.z, -ood.z), c1hiz = fmaf(n2.w, idir.z, -ood.z);
    tn0 = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tmin));
    tf0 = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tmax));
    tn1 = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tmin));
    tf1 = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tmax));
}

__device__ __forceinline__ float safe_rcp(float d)
{
    // direction component 0 -> a huge finite reciprocal keeps 0*inf NaNs out of the slab test
    const float kTiny = 1e-30f;
    float ad = fabsf(d) < kTiny ? copysignf(kTiny, d) : d;
    return 1.0f / ad;
}

// closest hit: min t in (tmin, tmax), ties -> smallest global triangle id (fx_hit_better).
__device__ __forceinline__ bool trace_closest(const DeviceBvh& bvh, float3 o, float3 d, float tmin, float tmax, HitRec& out)
{
    float best_t = __int_as_float(0x7f800000);   // +inf
    uint32_t best_gid = 0xffffffffu, best_k = 0;
    float best_u = 0.f, best_v = 0.f;
    if (bvh.n_tris == 0) return false;
    const float3 idir = make_float3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    const float3 ood = make_float3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    const fx_vec3 O = {o.x, o.y, o.z}, D = {d.x, d.y, d.z};
    int stack[kStackSize];
    int sp = 0;
    int cur = 0;
    for (;;) {
        while (cur >= 0) {
            const float4* np = bvh.nodes + 4 * (size_t)cur;
            const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3);
            float tn0, tf0, tn1, tf1;
            slab2(n0, n1, n2, idir, ood, tmin, fminf(tmax, best_t), tn0, tf0, tn1, tf1);
            const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = tn1 < tn0;
                cur = swap ? c1 : c0;
                if (sp < kStackSize) stack[sp++] = swap ? c0 : c1;
            } else if (h0) cur = c0;
            else if (h1) cur = c1;
            else {
                if (sp == 0) goto done;
                cur = stack[--sp];
            }
        }
        {   // leaf
            const uint32_t enc = (uint32_t)~cur;
            const uint32_t first = enc >> 2, count = (enc & 3u) + 1u;
            for (uint32_t i = 0; i < count; ++i) {
                const float4* tp = bvh.tris + 3 * (size_t)(first + i);
                const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
                fx_tri tr;
                tr.v0.x = a.x; tr.v0.y = a.y; tr.v0.z = a.z;
                tr.e1.x = b.x; tr.e1.y = b.y; tr.e1.z = b.z;
                tr.e2.x = c.x; tr.e2.y = c.y; tr.e2.z = c.z;
                float t, u, v;
                if (fx_ray_tri(O, D, tr, &t, &u, &v) == 0) continue;
                if (!(t > tmin && t < tmax)) continue;
                const uint32_t gid = (uint32_t)__float_as_int(a.w);
                if (fx_hit_better(t, gid, best_t, best_gid)) { best_t = t; best_gid = gid; best_k = first + i; best_u = u; best_v = v; }
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
done:
    if (best_gid == 0xffffffffu) return false;
    out.t = best_t; out.u = best_u; out.v = best_v; out.k = best_k; out.gid = best_gid;
    return true;
}

// any-hit with back-face culling: true iff some FRONT-facing triangle intersects (tmin, tmax).
// (OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES + payload-setting any-hit, deviceProgram.cu:242,284-300.)
__device__ __forceinline__ bool trace_any_front(const DeviceBvh& bvh, float3 o, float3 d, float tmin, float tmax)
{
    if (bvh.n_tris == 0) return false;
    const float3 idir = make_float3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    const float3 ood = make_float3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    const fx_vec3 O = {o.x, o.y, o.z}, D = {d.x, d.y, d.z};
    int stack[kStackSize];
    int sp = 0;
    int cur = 0;
    for (;;) {
        while (cur >= 0) {
            const float4* np = bvh.nodes + 4 * (size_t)cur;
            const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3);
            float tn0, tf0, tn1, tf1;
            slab2(n0, n1, n2, idir, ood, tmin, tmax, tn0, tf0, tn1, tf1);
            const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                cur = c0;
                if (sp < kStackSize) stack[sp++] = c1;
            } else if (h0) cur = c0;
            else if (h1) cur = c1;
            else {
                if (sp == 0) return false;
                cur = stack[--sp];
            }
        }
        const uint32_t enc = (uint32_t)~cur;
        const uint32_t first = enc >> 2, count = (enc & 3u) + 1u;
        for (uint32_t i = 0; i < count; ++i) {
            const float4* tp = bvh.tris + 3 * (size_t)(first + i);
            const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
            fx_tri tr;
            tr.v0.x = a.x; tr.v0.y = a.y; tr.v0.z = a.z;
            tr.e1.x = b.x; tr.e1.y = b.y; tr.e1.z = b.z;
            tr.e2.x = c.x; tr.e2.y = c.y; tr.e2.z = c.z;
            float t, u, v;
            if (fx_ray_tri(O, D, tr, &t, &u, &v) != 1) continue;
            if (t > tmin && t < tmax) return true;
        }
        if (sp == 0) return false;
        cur = stack[--sp];
    }
}

}  // namespace fovpt
