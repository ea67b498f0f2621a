// shade.cuh -- device-side material, light-probe and output arithmetic of the wavefront path tracer.
// Behavioural source (what must come out, not how it is written):
//   Disney-style BSDF          reference PT_sv5_/Disney.cuh:36-98 (helpers), :152-193 (pdf),
//                              :197-315 (sample), :318-427 (eval)
//   light probe                reference PT_sv5_/Probe.cuh:38-67 (dir<->uv, eval), :119-169 (CDF sample)
//   samplers / basis           reference PT_sv5_/maths.h:94-108, :229-276
//   tonemap + sRGB + quantise  reference PT_sv5_/deviceProgram.cu:126-131,586,597; cuda/helpers.h:35-62
// RNG draws are consumed in exactly the reference's order because seed-matched
// comparison against the CPU oracle depends on it.  Float results agree with the oracle to rounding
// (device libm vs glibc), which is what the statistical parity tests measure.
//
// template<bool FAST>: FAST=false uses IEEE-accurate device math; FAST=true uses the SFU
// approximations the reference's --use_fast_math build would have used (CMakeLists.txt:181).
#pragma once
#include <cuda_runtime.h>

#include "../../include/fovpt_exact.h"
#include "fovpt_layout.h"

namespace fovpt {

// ---------------------------------------------------------------- float3 algebra
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float s) { return make_float3(s, s, s); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float3 mix3(float3 a, float3 b, float t) { return a + t * (b - a); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a + t * (b - a); }
__device__ __forceinline__ float sq(float a) { return a * a; }
__device__ __forceinline__ float clamp01(float x) { return fmaxf(0.f, fminf(x, 1.f)); }

#define FOVPT_PI 3.141592653589793f
#define FOVPT_2PI (3.141592653589793f * 2.0f)
#define FOVPT_INV_PI (1.0f / 3.141592653589793f)
#define FOVPT_INV_2PI (1.0f / (3.141592653589793f * 2.0f))

template <bool FAST> struct Mth;
template <> struct Mth<false> {
    static __device__ __forceinline__ float sin(float x) { return sinf(x); }
    static __device__ __forceinline__ float cos(float x) { return cosf(x); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float log(float x) { return logf(x); }
    static __device__ __forceinline__ float pow(float x, float y) { return powf(x, y); }
    static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
    static __device__ __forceinline__ float rsqrt(float x) { return 1.0f / sqrtf(x); }
};
template <> struct Mth<true> {
    static __device__ __forceinline__ float sin(float x) { return __sinf(x); }
    static __device__ __forceinline__ float cos(float x) { return __cosf(x); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { __sincosf(x, s, c); }
    static __device__ __forceinline__ float log(float x) { return __logf(x); }
    static __device__ __forceinline__ float pow(float x, float y) { return __powf(x, y); }
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ float rsqrt(float x) { return rsqrtf(x); }
};

template <bool FAST> __device__ __forceinline__ float3 unit3(float3 v) { return v * Mth<FAST>::rsqrt(dot3(v, v)); }

// ---------------------------------------------------------------- sampling helpers
// tangent frame of a unit vector (maths.h:94-108)
template <bool FAST> __device__ __forceinline__ void tangent_frame(float3 n, float3& t, float3& b)
{
    if (fabsf(n.x) > fabsf(n.y)) { float il = Mth<FAST>::rsqrt(n.x * n.x + n.z * n.z); t = f3(-n.z * il, 0.f, n.x * il); }
    else { float il = Mth<FAST>::rsqrt(n.y * n.y + n.z * n.z); t = f3(0.f, n.z * il, -n.y * il); }
    b = cross3(n, t);
}

// ---------------------------------------------------------------- BSDF
struct BsdfParams {        // the Disney parameters the BSDF actually reads (of Material's 104 bytes)
    float3 color;
    float metallic, subsurface, specular, roughness, specularTint, clearcoat, clearcoatGloss, transmission;
};

__device__ __forceinline__ BsdfParams bsdf_params(const MaterialBlob& m)
{
    BsdfParams p;
    p.color = f3(m.color[0], m.color[1], m.color[2]);
    p.metallic = m.metallic; p.subsurface = m.subsurface; p.specular = m.specular; p.roughness = m.roughness;
    p.specularTint = m.specularTint; p.clearcoat = m.clearcoat; p.clearcoatGloss = m.clearcoatGloss;
    p.transmission = m.transmission;
    return p;
}

__device__ __forceinline__ float schlick_weight(float u) { float m = clamp01(1.f - u); float m2 = m * m; return m2 * m2 * m; }
__device__ __forceinline__ float gtr2(float ndh, float a) { float a2 = a * a; float t = 1.f + (a2 - 1.f) * ndh * ndh; return a2 / (FOVPT_PI * t * t); }
template <bool FAST> __device__ __forceinline__ float gtr1(float ndh, float a)
{
    if (a >= 1.f) return FOVPT_INV_PI;
    float a2 = a * a;
    float t = 1.f + (a2 - 1.f) * ndh * ndh;
    return (a2 - 1.f) / (FOVPT_PI * Mth<FAST>::log(a2) * t);
}
__device__ __forceinline__ float smith_ggx(float ndv, float ag) { float a = ag * ag, b = ndv * ndv; return 1.f / (ndv + sqrtf(a + b - a * b)); }

// exact dielectric Fresnel (Disney.cuh:81-98)
__device__ __forceinline__ float fresnel_dielectric(float vdn, float eta_i, float eta_t)
{
    float s2 = sq(eta_i / eta_t) * (1.f - vdn * vdn);
    if (s2 > 1.f) return 1.f;
    float ldn = sqrtf(1.f - s2);
    float eta = eta_t / eta_i;
    float r1 = (vdn - eta * ldn) / (vdn + eta * ldn);
    float r2 = (ldn - eta * vdn) / (ldn + eta * vdn);
    return 0.5f * (sq(r1) + sq(r2));
}

// Terms of the BSDF that depend on the material, the normal and the VIEW direction only.  One hit evaluates the BSDF for two
// light directions (the probe sample of the next-event estimate and the sampled continuation) and its pdf for both; the
// reference recomputes these terms in every call (Disney.cuh:152-193, :318-427), here they are computed once per hit.
// Every term keeps the reference's expression, so the values are the ones the separate calls produced.
struct BsdfView {
    float ndv, a;          // dot(n, v); max(0.001, roughness)
    float F;               // fresnel_dielectric(ndv, eta_i, eta_o)
    float Gv, FV;          // smith_ggx(ndv, a); schlick_weight(ndv)
    float3 spec0;
};

template <bool FAST>
__device__ __forceinline__ void bsdf_view_terms(const BsdfParams& m, float3 albedo, BsdfView& w)
{
    const float lum = 0.3f * albedo.x + 0.6f * albedo.y + 0.1f * albedo.z;
    const float3 tint = lum > 0.f ? albedo * Mth<FAST>::rcp(lum) : f3(1.f);
    w.spec0 = mix3((m.specular * 0.08f) * mix3(f3(1.f), tint, m.specularTint), albedo, m.metallic);
    w.Gv = smith_ggx(w.ndv, w.a);
    w.FV = schlick_weight(w.ndv);
}

// ... and the terms that depend on one light direction, shared by its pdf and its value: the half vector's cosines and the
// GGX distribution at it (BSDFPdf and BSDFEval both normalise l + v and both evaluate GTR2 there, Disney.cuh:168-178, :330-352)
struct BsdfDir { float ndl, ndh, ldh, D; };

template <bool FAST>
__device__ __forceinline__ void bsdf_dir_terms(const BsdfView& w, float3 n, float3 v, float3 l, BsdfDir& d)
{
    d.ndl = dot3(n, l);
    float3 h = l + v;
    const float hl = dot3(h, h);
    h = hl > 0.f ? h * Mth<FAST>::rsqrt(hl) : f3(0.f);
    d.ndh = dot3(n, h); d.ldh = dot3(l, h);
    d.D = d.ndl > 0.f ? gtr2(d.ndh, w.a) : 0.f;           // only the upper hemisphere branches read it
}

template <bool FAST>
__device__ __forceinline__ float bsdf_pdf(const BsdfParams& m, const BsdfView& w, const BsdfDir& d)
{
    const float ldn = d.ndl;
    if (ldn <= 0.f) return mixf(FOVPT_INV_2PI * m.subsurface * 0.5f, 0.f, m.transmission);
    const float cth = fabsf(d.ndh);
    const float pdf_half = d.D * cth;
    const float pdf_spec = 0.25f * pdf_half / fmaxf(1.e-6f, d.ldh);
    const float pdf_diff = fabsf(ldn) * FOVPT_INV_PI * (1.f - m.subsurface);
    return mixf(mixf(pdf_diff, pdf_spec, 0.5f), pdf_spec * w.F, m.transmission);
}

template <bool FAST>
__device__ __forceinline__ float3 bsdf_eval(const BsdfParams& m, const BsdfView& w, const BsdfDir& d, float3 albedo, float eta_i, float eta_o)
{
    const float ndl = d.ndl, ndv = w.ndv;
    const float ndh = d.ndh, ldh = d.ldh;
    const float3 spec0 = w.spec0;
    float3 bsdf = f3(0.f), brdf = f3(0.f);
    const float a = w.a;
    if (m.transmission > 0.f) {
        if (ndl <= 0.f) {
            bsdf = f3(m.transmission * (1.f - w.F) / fabsf(ndl) * (1.f - m.metallic));
        } else {
            const float Ds = d.D;
            const float FH = fresnel_dielectric(ldh, eta_i, eta_o);
            const float3 Fs = mix3(spec0, f3(1.f), FH);
            const float Gs = w.Gv * smith_ggx(ndl, a);
            bsdf = Gs * Fs * Ds;
        }
    }
    if (m.transmission < 1.f) {
        if (ndl <= 0.f) {
            if (m.subsurface > 0.f) {
                // uses the material colour, not the textured albedo (Disney.cuh:374)
                const float3 s = f3(sqrtf(m.color.x), sqrtf(m.color.y), sqrtf(m.color.z));
                const float FL = schlick_weight(fabsf(ndl));
                const float Fd = (1.f - 0.5f * FL) * (1.f - 0.5f * w.FV);
                brdf = FOVPT_INV_PI * s * m.subsurface * Fd * (1.f - m.metallic);
            }
        } else {
            const float Ds = d.D;
            const float FH = schlick_weight(ldh);
            const float3 Fs = mix3(spec0, f3(1.f), FH);
            const float Gs = w.Gv * smith_ggx(ndl, a);
            const float FL = schlick_weight(ndl);
            const float Fd90 = 0.5f + 2.f * ldh * ldh * m.roughness;
            const float Fd = mixf(1.f, Fd90, FL) * mixf(1.f, Fd90, w.FV);
            // clearcoat lobe (Disney.cuh:411-417): the reference evaluates it for every material and multiplies by m.clearcoat;
            // with clearcoat == 0 (the constructor default, Material.h:31) the product is an exact zero, so the log, the two
            // Smith terms and their divisions are skipped without changing a bit of the result
            float cc = 0.f;
            if (m.clearcoat != 0.f) {
                const float Dr = gtr1<FAST>(ndh, mixf(.1f, .001f, m.clearcoatGloss));
                const float Fc = mixf(.04f, 1.f, FH);
                const float Gr = smith_ggx(ndl, .25f) * smith_ggx(ndv, .25f);
                cc = m.clearcoat * Gr * Fc * Dr;
            }
            brdf = FOVPT_INV_PI * Fd * albedo * (1.f - m.metallic) * (1.f - m.subsurface) + Gs * Fs * Ds + f3(cc);
        }
    }
    return mix3(brdf, bsdf, m.transmission);
}

// GGX half-vector sample shared by the two specular branches (Disney.cuh:211-219 and :287-301)
template <bool FAST>
__device__ __forceinline__ float3 ggx_reflect(float a, float3 tu, float3 tv, float3 n, float3 view, float r1, float r2)
{
    const float cth = sqrtf((1.f - r2) / (1.f + (sq(a) - 1.f) * r2));
    const float sth = sqrtf(fmaxf(0.f, 1.f - sq(cth)));
    float sp, cp;
    Mth<FAST>::sincos(r1 * FOVPT_2PI, &sp, &cp);
    float3 h = tu * (sth * cp) + tv * (sth * sp) + n * cth;
    if (dot3(h, view) <= 0.f) h = -h;
    return 2.f * dot3(view, h) * h - view;
}

// Draws the continuation direction (Disney.cuh:197-315 up to its closing BSDFPdf call).  Returns < 0 when the pdf of `light` is
// still to be evaluated with bsdf_pdf (every branch but smooth refraction), else the final pdf (<= 0: path ends).
// Consumes 2, 4, 5 or 7 uniform draws depending on the branch.  w.ndv, w.a and w.F must be set.
template <bool FAST>
__device__ __forceinline__ float bsdf_sample_dir(const BsdfParams& m, const BsdfView& w, float eta_i, float eta_o, float3 tu, float3 tv, float3 n,
                                                 float3 view, fx_rng& rng, float3& light)
{
    if (fx_rng_f(&rng) < m.transmission) {
        if (fx_rng_f(&rng) < w.F) {
            const float r1 = fx_rng_f(&rng), r2 = fx_rng_f(&rng);
            light = ggx_reflect<FAST>(w.a, tu, tv, n, view, r1, r2);
        } else {
            // smooth refraction (Disney.cuh:36-49,231-247)
            const float eta = eta_i / eta_o;
            const float ci = w.ndv;
            const float s2t = eta * eta * fmaxf(0.f, 1.f - ci * ci);
            if (s2t >= 1.f) return 0.f;
            const float ct = sqrtf(1.f - s2t);
            light = eta * (-view) + (eta * ci - ct) * n;
            return (1.f - w.F) * m.transmission;
        }
    } else {
        const float r1 = fx_rng_f(&rng), r2 = fx_rng_f(&rng);
        if (fx_rng_f(&rng) < 0.5f) {
            if (fx_rng_f(&rng) < m.subsurface) {
                const float z = fx_rng_f(&rng);
                const float wz = sqrtf(1.f - z * z);
                float s, c;
                Mth<FAST>::sincos(FOVPT_2PI * fx_rng_f(&rng), &s, &c);
                light = tu * (c * wz) + tv * (s * wz) - n * z;
            } else {
                const float r = sqrtf(r1);
                float s, c;
                Mth<FAST>::sincos(FOVPT_2PI * r2, &s, &c);
                const float dx = r * c, dy = r * s;
                const float dz = sqrtf(fmaxf(0.f, 1.f - dx * dx - dy * dy));
                light = tu * dx + tv * dy + n * dz;
            }
        } else {
            light = ggx_reflect<FAST>(w.a, tu, tv, n, view, r1, r2);
        }
    }
    return -1.f;
}

// ---------------------------------------------------------------- probe
__device__ __forceinline__ float2 probe_dir_to_uv(float3 d)
{
    const float theta = acosf(fmaxf(-1.f, fminf(d.y, 1.f)));
    const float phi = (d.x == 0.f && d.z == 0.f) ? 0.f : atan2f(d.z, d.x);
    return make_float2((FOVPT_PI + phi) * FOVPT_INV_PI * 0.5f, theta * FOVPT_INV_PI);
}

__device__ __forceinline__ float4 probe_eval(const ProbeBlob& p, float2 uv)
{
    const int px = max(0, min((int)(uv.x * p.width), p.width - 1));
    const int py = max(0, min((int)(uv.y * p.height), p.height - 1));
    return __ldg(&p.data[py * p.width + px]);
}

__device__ __forceinline__ int cdf_lower_bound(const float* __restrict__ a, int lo, int hi, float v)
{
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (__ldg(&a[mid]) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Guide tables for the CDF inversion (built once per probe by k_probe_guides): gy[k] / gx[row][k] =
// lower_bound(cdf, k / kGuide) for k = 0..kGuide.  A sample u with floor(u * kGuide) = k has its lower
// bound inside [g[k], g[k+1]] (lower_bound is monotone and k / kGuide <= u < (k+1) / kGuide holds exactly,
// kGuide being a power of two), so searching only that range returns the SAME index as the reference's
// full-row binary search (Probe.cuh:119-136,147-153) with 1-3 dependent loads instead of 10-11 per axis.
#ifndef FOVPT_GUIDE
#define FOVPT_GUIDE 512
#endif
constexpr int kGuide = FOVPT_GUIDE;
struct ProbeGuide { const uint16_t* gy = nullptr; const uint16_t* gx = nullptr; };

// one importance sample of the probe: row by cdfY then column by that row's cdfX (Probe.cuh:138-169)
template <bool FAST>
__device__ __forceinline__ void probe_sample(const ProbeBlob& p, const ProbeGuide& g, fx_rng& rng, float3& dir, float3& color, float& pdf)
{
    const float r1 = fx_rng_f(&rng), r2 = fx_rng_f(&rng);
    int row, col;
    if (g.gy) {
        const int ky = min((int)(r1 * (float)kGuide), kGuide - 1);
        row = cdf_lower_bound(p.cdfValuesY, (int)__ldg(&g.gy[ky]), (int)__ldg(&g.gy[ky + 1]), r1);
        const int kx = min((int)(r2 * (float)kGuide), kGuide - 1);
        const uint16_t* gr = g.gx + (size_t)min(row, p.height - 1) * (kGuide + 1);
        const int base = row * p.width;
        col = cdf_lower_bound(p.cdfValuesX, base + (int)__ldg(&gr[kx]), base + (int)__ldg(&gr[kx + 1]), r2) - base;
    } else {
        row = cdf_lower_bound(p.cdfValuesY, 0, p.height, r1);
        col = cdf_lower_bound(p.cdfValuesX, row * p.width, (row + 1) * p.width, r2) - row * p.width;
    }
    const float4 c = __ldg(&p.data[row * p.width + col]);
    color = f3(c.x, c.y, c.z);
    pdf = __ldg(&p.pdfValuesX[row * p.width + col]) * __ldg(&p.pdfValuesY[row]);
    const float u = col / float(p.width), v = row / float(p.height);
    float st, ct;
    Mth<FAST>::sincos(v * FOVPT_PI, &st, &ct);
    if (st == 0.f) pdf = 0.f;
    else pdf *= (float)(p.width * p.height) / (2.f * FOVPT_PI * FOVPT_PI * st);
    float sp, cp;
    Mth<FAST>::sincos(u * 2.f * FOVPT_PI, &sp, &cp);
    dir = f3(-st * cp, ct, -st * sp);
}

// ---------------------------------------------------------------- output
template <bool FAST> __device__ __forceinline__ float srgb_oetf(float c)
{
    c = clamp01(c);
    return c < 0.0031308f ? 12.92f * c : 1.055f * Mth<FAST>::pow(c, 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ uint32_t quant8(float x) { x = clamp01(x); return min((uint32_t)(x * 256.0f), 255u); }

// exposure, Reinhard with white = 1, sRGB, 8-bit pack (deviceProgram.cu:586,597,126-131; helpers.h:35-62)
template <bool FAST> __device__ __forceinline__ uint32_t tonemap_rgba8(float3 accum, float exposure)
{
    const float3 e = accum * exposure;
    const float lum = 0.2126f * e.x + 0.7152f * e.y + 0.0722f * e.z;
    const float3 m = e * (1.0f / (1.0f + lum));
    return quant8(srgb_oetf<FAST>(m.x)) | (quant8(srgb_oetf<FAST>(m.y)) << 8) | (quant8(srgb_oetf<FAST>(m.z)) << 16) | (255u << 24);
}

// RGBA8 texture through the reference's sampler (linear, wrap, normalised coordinates, normalised float;
// SimplePathtracer.cpp:775-789): the texture unit's own fixed-point arithmetic, fx_tex_bilinear_rgba8 of fovpt_exact.h
__device__ __forceinline__ float3 texture_bilinear(const uint32_t* __restrict__ texels, int w, int h, float x, float y)
{
    float o[4];
    fx_tex_bilinear_rgba8(texels, w, h, x, y, o);
    return f3(o[0], o[1], o[2]);
}

}  // namespace fovpt
