// LaunchParams.h -- the per-launch contract between host and device: field for field the reference's
// struct LaunchParams (PT_sv5_/LaunchParams.h:49-91, 248 bytes) and device Probe (Probe.cuh:6-21, 64 bytes).
// libfovpt receives this struct as a byte blob (include/fovpt.h); the asserts pin the layout.
#pragma once
#include <cstddef>

#include "Probe.h"

struct Probe {
    int    width;
    int    height;
    Color* data;              // device pointer
    float3 offset;
    float* pdfValuesX;        // device pointers
    float* cdfValuesX;
    float* pdfValuesY;
    float* cdfValuesY;
};
static_assert(sizeof(Probe) == 64, "Probe layout");

typedef unsigned long long OptixTraversableHandle;    // what the reference stores at offset 160; here: an opaque libfovpt handle

struct LaunchParams {
    struct {
        float4* accum_buffer;
        uchar4* frame_buffer;
        float4* color_buffer;
        float4* normal_buffer;
        float4* albedo_buffer;
        int2    size;
        unsigned int subframe_index;
        uint3   factor;          // foveation: launch index -> pixel scale
        int     fillSize;        // splat block edge
        uint2   c;               // gaze point (centre of foveation), pixels
        float   r_inner, r_outer;
        uint2   offset;
        unsigned int redraw;
    } frame;
    struct {
        float3 eye, U, V, W;
    } camera;
    unsigned int samples_per_launch;
    OptixTraversableHandle traversable;
    Probe probe;
    int2  viewportSize;
    float white;
};
static_assert(sizeof(LaunchParams) == 248, "LaunchParams must keep the reference's 248-byte layout");
static_assert(offsetof(LaunchParams, frame.size) == 40 && offsetof(LaunchParams, frame.subframe_index) == 48 &&
              offsetof(LaunchParams, frame.factor) == 52 && offsetof(LaunchParams, frame.fillSize) == 64 &&
              offsetof(LaunchParams, frame.c) == 72 && offsetof(LaunchParams, frame.r_inner) == 80 &&
              offsetof(LaunchParams, frame.offset) == 88 && offsetof(LaunchParams, frame.redraw) == 96 &&
              offsetof(LaunchParams, camera) == 104 && offsetof(LaunchParams, samples_per_launch) == 152 &&
              offsetof(LaunchParams, traversable) == 160 && offsetof(LaunchParams, probe) == 168 &&
              offsetof(LaunchParams, viewportSize) == 232 && offsetof(LaunchParams, white) == 240, "LaunchParams offsets");
