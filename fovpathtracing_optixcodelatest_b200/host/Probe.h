// Probe.h -- host light probe with luminance CDFs; same struct and BuildCDF arithmetic as the
// reference's ProbeData (PT_sv5_/Probe.h:5-100), so the tables are bit-identical to the ones the
// reference would upload.  The device-side Probe blob (64 B, Probe.cuh:6-21) is in LaunchParams.h.
#pragma once
#include "fovpt_vec.h"

typedef float4 Color;                               // maths.h:161

inline float Luminance(const Color& c) { return c.x * 0.3f + c.y * 0.6f + c.z * 0.1f; }   // maths.h:163-166

struct ProbeData {
    int    width = 0;
    int    height = 0;
    Color* data = nullptr;
    float3 offset = make_float3(0.f, 0.f, 0.f);   // world space offset to warp the skybox
    bool   valid = false;
    float* pdfValuesX = nullptr;
    float* cdfValuesX = nullptr;
    float* pdfValuesY = nullptr;
    float* cdfValuesY = nullptr;

    // float32 running sums in the reference's order (Probe.h:29-77)
    inline void BuildCDF()
    {
        pdfValuesX = new float[(size_t)width * height]; pdfValuesY = new float[height];
        cdfValuesX = new float[(size_t)width * height]; cdfValuesY = new float[height];
        float totalY = 0.0f;
        for (int j = 0; j < height; ++j) {
            float totalX = 0.0f;
            float* px = pdfValuesX + (size_t)j * width; float* cx = cdfValuesX + (size_t)j * width;
            for (int i = 0; i < width; ++i) { const float wgt = Luminance(data[(size_t)j * width + i]); totalX += wgt; px[i] = wgt; cx[i] = totalX; }
            const float inv = 1.0f / totalX;
            for (int i = 0; i < width; ++i) { px[i] *= inv; cx[i] *= inv; }
            totalY += totalX;
            pdfValuesY[j] = totalX; cdfValuesY[j] = totalY;
        }
        for (int j = 0; j < height; ++j) { cdfValuesY[j] /= totalY; pdfValuesY[j] /= totalY; }
        valid = true;
    }
};
