// Material.h -- Disney parameter block of one mesh.  Same fields, order, defaults and size (104 B) as
// the reference's struct Material (PT_sv5_/Material.h:11-70): it crosses the C ABI as a byte blob.
#pragma once
#include <cmath>
#include <cstddef>

#include "fovpt_vec.h"

static const int MATERIAL_FLAG_NONE = 0;
static const int MATERIAL_FLAG_SHADOW_CATCHER = 1 << 0;      // Material.h:8-9

struct Material {
    // defaults of the reference constructor (Material.h:13-38)
    float3 emission = make_float3(1.0f, 1.0f, 1.0f);
    float3 color = make_float3(1.0f, 0.0f, 0.0f);
    float3 absorption = make_float3(1.0f, 1.0f, 1.0f);
    float eta = 1.4f;              // 0 -> index of refraction inferred from `specular`
    float metallic = 0.5f;
    float subsurface = 0.0f;
    float specular = 1.0f;
    float roughness = 1.0f;
    float specularTint = 1.0f;
    float anisotropic = 0.0f;
    float sheen = 0.0f;
    float sheenTint = 0.0f;
    float clearcoat = 0.0f;
    float clearcoatGloss = 1.0f;
    float transmission = 0.4f;
    float bump = 0.0f;
    float3 bumpTile = make_float3(1.0f, 1.0f, 1.0f);
    int flags = 0;

    float GetIndexOfRefraction() const      // Material.h:40-46
    {
        return eta == 0.0f ? 2.0f / (1.0f - std::sqrt(0.08f * specular)) - 1.0f : eta;
    }
};
static_assert(sizeof(Material) == 104, "Material must keep the reference's 104-byte layout");
static_assert(offsetof(Material, eta) == 36 && offsetof(Material, transmission) == 80 && offsetof(Material, flags) == 100, "Material offsets");
