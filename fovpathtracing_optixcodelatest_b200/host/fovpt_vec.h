// fovpt_vec.h -- the CUDA vector types the reference's host headers use (float3, int2, uint3, ...;
// reference Model.h:4 includes <vector_types.h>) plus the handful of float3 operations the host side
// needs.  Host-only code: the CUDA headers are used for the type definitions alone.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector_functions.h>
#include <vector_types.h>

namespace fovpt_host {
inline float3 sub(const float3& a, const float3& b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 mul(const float3& a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(const float3& a) { return sqrtf(dot(a, a)); }
// sutil/vec_math.h:553-561 normalizes with a reciprocal square root: v * (1 / sqrt(dot))
inline float3 normalize(const float3& a) { const float inv = 1.0f / sqrtf(dot(a, a)); return mul(a, inv); }
}  // namespace fovpt_host
