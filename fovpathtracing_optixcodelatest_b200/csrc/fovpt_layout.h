// fovpt_layout.h -- byte layouts of the blobs that cross the C ABI, as the device code sees them.
// They reproduce the reference's structs field for field (PT_sv5_/LaunchParams.h:49-91,
// Material.h:11-70, Probe.cuh:6-21); sizes and offsets are the ones g++ reports for the unmodified
// reference headers and are asserted below so a drift fails the build.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace fovpt {

struct MaterialBlob {            // 104 B
    float emission[3], color[3], absorption[3];
    float eta, metallic, subsurface, specular, roughness, specularTint, anisotropic, sheen, sheenTint,
          clearcoat, clearcoatGloss, transmission, bump;
    float bumpTile[3];
    int   flags;
};
static_assert(sizeof(MaterialBlob) == 104, "Material must stay 104 bytes");
static_assert(offsetof(MaterialBlob, color) == 12 && offsetof(MaterialBlob, metallic) == 40 && offsetof(MaterialBlob, specular) == 48 &&
              offsetof(MaterialBlob, roughness) == 52 && offsetof(MaterialBlob, specularTint) == 56 && offsetof(MaterialBlob, clearcoat) == 72 &&
              offsetof(MaterialBlob, clearcoatGloss) == 76, "k_shade reads the material as float2 words at these offsets");
static_assert(offsetof(MaterialBlob, eta) == 36 && offsetof(MaterialBlob, transmission) == 80 &&
              offsetof(MaterialBlob, flags) == 100, "Material field offsets");
constexpr int kMaterialShadowCatcher = 1;     // Material.h:9

struct ProbeBlob {               // 64 B
    int          width, height;
    const float4* data;
    float        offset[3];
    uint32_t     pad_;
    const float* pdfValuesX;
    const float* cdfValuesX;
    const float* pdfValuesY;
    const float* cdfValuesY;
};
static_assert(sizeof(ProbeBlob) == 64 && offsetof(ProbeBlob, pdfValuesX) == 32, "Probe layout");

struct LaunchParamsBlob {        // 248 B
    struct Frame {
        float4*   accum_buffer;
        uint32_t* frame_buffer;          // uchar4 in the reference
        float4*   color_buffer;
        float4*   normal_buffer;
        float4*   albedo_buffer;
        int       size_x, size_y;
        unsigned  subframe_index;
        unsigned  factor_x, factor_y, factor_z;
        int       fillSize;
        unsigned  pad0_;
        unsigned  c_x, c_y;
        float     r_inner, r_outer;
        unsigned  offset_x, offset_y;
        unsigned  redraw;
        unsigned  pad1_;
    } frame;
    struct Camera { float eye[3], U[3], V[3], W[3]; } camera;
    unsigned           samples_per_launch;
    unsigned           pad0_;
    unsigned long long traversable;
    ProbeBlob          probe;
    int                viewport_x, viewport_y;
    float              white;
    unsigned           pad1_;
};
static_assert(sizeof(LaunchParamsBlob) == 248, "LaunchParams must stay 248 bytes");
static_assert(offsetof(LaunchParamsBlob, frame.size_x) == 40 && offsetof(LaunchParamsBlob, frame.subframe_index) == 48 &&
              offsetof(LaunchParamsBlob, frame.factor_x) == 52 && offsetof(LaunchParamsBlob, frame.fillSize) == 64 &&
              offsetof(LaunchParamsBlob, frame.c_x) == 72 && offsetof(LaunchParamsBlob, frame.r_inner) == 80 &&
              offsetof(LaunchParamsBlob, frame.offset_x) == 88 && offsetof(LaunchParamsBlob, frame.redraw) == 96,
              "LaunchParams.frame offsets");
static_assert(offsetof(LaunchParamsBlob, camera) == 104 && offsetof(LaunchParamsBlob, samples_per_launch) == 152 &&
              offsetof(LaunchParamsBlob, traversable) == 160 && offsetof(LaunchParamsBlob, probe) == 168 &&
              offsetof(LaunchParamsBlob, viewport_x) == 232 && offsetof(LaunchParamsBlob, white) == 240,
              "LaunchParams offsets");

}  // namespace fovpt
