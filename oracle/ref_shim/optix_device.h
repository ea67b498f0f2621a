// CPU emulation of the slice of the OptiX 8 *device* API that
// /root/reference/PT_sv5_/deviceProgram.cu uses  --  TEST INFRASTRUCTURE ONLY.
//
// Purpose: let g++ compile the reference's deviceProgram.cu UNMODIFIED so that its
// __raygen__/__closesthit__/__miss__ programs can be executed on host cores and used
// (a) to pin the CPU restatement in oracle/ and (b) as the "reference" CPU baseline.
// What OptiX itself did (BVH build, traversal, triangle test, SBT dispatch) has no
// source in the reference; here it is provided by oracle/bvh.h through the four
// fovref:: hooks declared below ("parity unpinned" for that part).
#pragma once
#include <optix.h>
#include <vector_types.h>
#include <vector_functions.h>

enum {
    OPTIX_RAY_FLAG_NONE                        = 0,
    OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT      = 1 << 2,
    OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES  = 1 << 4
};
typedef unsigned int OptixVisibilityMask;

namespace fovref {
struct Hit { int mesh; int prim; float t, u, v; };
struct ThreadState {
    uint3        launch_index;
    unsigned int p0, p1;          // payload registers of the trace in flight
    float3       ray_o, ray_d;
    float        ray_t;           // optixGetRayTmax() inside closest-hit == hit distance
    int          prim;
    const void*  sbt;
    float2       bary;
    unsigned int traces_this_index;   // trace counter since the launch loop reset it
};
extern thread_local ThreadState ts;
// provided by oracle/ref_driver.cpp
bool        trace_closest(const float3& o, const float3& d, float tmin, float tmax, Hit* hit);
bool        trace_any_front(const float3& o, const float3& d, float tmin, float tmax);
const void* sbt_record(int mesh);
float4      tex_fetch(unsigned long long tex, float x, float y);
}  // namespace fovref

extern "C" void __closesthit__radiance();
extern "C" void __miss__radiance();
extern "C" void __anyhit__occlusion();
extern "C" void __closesthit__occlusion();
extern "C" void __miss__occlusion();

inline uint3        optixGetLaunchIndex()           { return fovref::ts.launch_index; }
inline unsigned int optixGetPayload_0()             { return fovref::ts.p0; }
inline unsigned int optixGetPayload_1()             { return fovref::ts.p1; }
inline void         optixSetPayload_0(unsigned v)   { fovref::ts.p0 = v; }
inline float3       optixGetWorldRayDirection()     { return fovref::ts.ray_d; }
inline float3       optixGetWorldRayOrigin()        { return fovref::ts.ray_o; }
inline float        optixGetRayTmax()               { return fovref::ts.ray_t; }
inline unsigned int optixGetPrimitiveIndex()        { return (unsigned)fovref::ts.prim; }
inline unsigned long long optixGetSbtDataPointer()  { return (unsigned long long)fovref::ts.sbt; }
inline float2       optixGetTriangleBarycentrics()  { return fovref::ts.bary; }

// radiance trace: two payload registers carrying the PRD pointer (deviceProgram.cu:209)
inline void optixTrace(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float,
                       OptixVisibilityMask, unsigned int /*flags*/, unsigned int, unsigned int, unsigned int,
                       unsigned int& p0, unsigned int& p1)
{
    fovref::ThreadState& s = fovref::ts;
    fovref::ThreadState saved = s;
    s.p0 = p0; s.p1 = p1; s.ray_o = o; s.ray_d = d;
    fovref::Hit h;
    if (fovref::trace_closest(o, d, tmin, tmax, &h)) {
        s.ray_t = h.t; s.prim = h.prim; s.sbt = fovref::sbt_record(h.mesh);
        s.bary = make_float2(h.u, h.v);
        __closesthit__radiance();
    } else {
        s.ray_t = tmax;
        __miss__radiance();
    }
    p0 = s.p0; p1 = s.p1;
    unsigned int n = s.traces_this_index;
    s = saved;
    s.traces_this_index = n;
}

// occlusion trace: one payload register (deviceProgram.cu:234).  With
// CULL_BACK_FACING and a non-terminating any-hit that sets the payload, the result
// is "some front-facing triangle intersects (tmin,tmax)"; closest-hit is a no-op and
// miss (payload <- 0) runs only when nothing was hit (deviceProgram.cu:284-300).
inline void optixTrace(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float,
                       OptixVisibilityMask, unsigned int /*flags*/, unsigned int, unsigned int, unsigned int,
                       unsigned int& p0)
{
    fovref::ThreadState& s = fovref::ts;
    fovref::ThreadState saved = s;
    s.p0 = p0; s.ray_o = o; s.ray_d = d;
    if (fovref::trace_any_front(o, d, tmin, tmax)) { __anyhit__occlusion(); __closesthit__occlusion(); }
    else                                            { __miss__occlusion(); }
    p0 = s.p0;
    unsigned int n = s.traces_this_index;
    s = saved;
    s.traces_this_index = n;
}

// tex2D<float4> on a cudaTextureObject_t (deviceProgram.cu:664); the CUDA header
// only declares it under __CUDACC__.
template <typename T> inline T tex2D(cudaTextureObject_t tex, float x, float y);
template <> inline float4 tex2D<float4>(cudaTextureObject_t tex, float x, float y)
{
    return fovref::tex_fetch((unsigned long long)tex, x, y);
}
