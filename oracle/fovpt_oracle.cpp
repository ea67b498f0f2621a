// oracle/fovpt_oracle.cpp -- THE CPU ORACLE ("port").  TEST INFRASTRUCTURE ONLY: nothing in the
// product path may include, link, import or execute this file (see orc_api.h).
//
// A self-contained scalar C++ restatement of the reference's foveated path-tracing hot path.
// Every function cites the reference lines it follows (paths relative to /root/reference).
// It is pinned in two ways (tests/test_oracle_*.py):
//   * against oracle/_ref/libfovref.so -- the reference's own deviceProgram.cu, Disney.cuh,
//     Probe.cuh, maths.h, sample.h, random.h, helpers.h, Probe.h, Camera.cpp compiled unmodified
//     for the host -- function by function and on whole frames, BIT-EXACT (same compiler flags);
//   * against tests/golden/*.npz generated from that library (script: tests/golden/make_golden.py),
//     so the pin also holds where /root/reference does not exist (the GPU box).
// What is NOT pinned by any reference source: BVH traversal and the ray/triangle test, which the
// reference delegates to closed-source OptiX ("parity unpinned" for that part; the specification is
// include/fovpt_exact.h + oracle/bvh.h).
//
// Extensions beyond sv5's literals (all off by default, used by BASELINE configs C1-C3/C5):
// runtime max_depth, exposure, the vmv2 accumulation lerp, subframe seeding of every pass.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "common.h"

namespace {

// ---------------------------------------------------------------- small vector algebra
// same operator semantics as sutil/vec_math.h:380-470 (notably  v / s  ==  v * (1.0f / s)).
struct f3 { float x, y, z; };
struct f2 { float x, y; };
struct f4 { float x, y, z, w; };
inline f3 mk3(float x, float y, float z) { return {x, y, z}; }
inline f3 mk3(float s) { return {s, s, s}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator-(float a, f3 b) { return {a - b.x, a - b.y, a - b.z}; }
inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator*(float s, f3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator/(f3 a, float s) { float inv = 1.0f / s; return a * inv; }
inline f3 operator/(f3 a, f3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline void operator+=(f3& a, f3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline void operator*=(f3& a, f3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
inline void operator*=(f3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
inline void operator/=(f3& a, float s) { float inv = 1.0f / s; a *= inv; }
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline f3 cross(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline f3 normalize(f3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
inline float lerpf(float a, float b, float t) { return a + t * (b - a); }
inline f3 lerp3(f3 a, f3 b, float t) { return a + t * (b - a); }
inline f3 faceforward(f3 n, f3 i, f3 nref) { return n * copysignf(1.0f, dot(i, nref)); }   // vec_math.h:580-583
inline float sqr(float a) { return a * a; }

const float kPi = 3.141592653589793f;            // maths.h:28-31
const float k2Pi = 3.141592653589793f * 2.0f;
const float kInvPi = 1.0f / kPi;
const float kInv2Pi = 1.0f / k2Pi;

// ---------------------------------------------------------------- launch-contract structs
struct Material {                                  // Material.h:11-70, 104 bytes
    f3 emission, color, absorption;
    float eta, metallic, subsurface, specular, roughness, specularTint, anisotropic, sheen, sheenTint,
          clearcoat, clearcoatGloss, transmission, bump;
    f3 bumpTile;
    int flags;
    float ior() const { return eta == 0.0f ? 2.0f / (1.0f - sqrtf(0.08f * specular)) - 1.0f : eta; }   // :40-46
};
static_assert(sizeof(Material) == 104, "Material layout");
const int kShadowCatcher = 1;                      // Material.h:9

struct Probe {                                     // Probe.cuh:6-21, 64 bytes
    int width, height; f4* data; f3 offset; float* pdfX; float* cdfX; float* pdfY; float* cdfY;
};
static_assert(sizeof(Probe) == 64, "Probe layout");

struct LaunchParams {                              // LaunchParams.h:49-91, 248 bytes
    struct {
        f4* accum_buffer; uint32_t* frame_buffer; f4* color_buffer; f4* normal_buffer; f4* albedo_buffer;
        int size_x, size_y; unsigned subframe_index; unsigned factor_x, factor_y, factor_z; int fillSize;
        alignas(8) unsigned c_x; unsigned c_y; float r_inner, r_outer; alignas(8) unsigned offset_x; unsigned offset_y;
        unsigned redraw;
    } frame;
    struct { f3 eye, U, V, W; } camera;
    unsigned samples_per_launch;
    unsigned long long traversable;
    Probe probe;
    int viewport_x, viewport_y;
    float white;
};
static_assert(sizeof(LaunchParams) == 248, "LaunchParams layout");
static_assert(offsetof(LaunchParams, camera) == 104 && offsetof(LaunchParams, samples_per_launch) == 152, "layout");
static_assert(offsetof(LaunchParams, traversable) == 160 && offsetof(LaunchParams, probe) == 168, "layout");

// ---------------------------------------------------------------- RNG (maths.h:170-227, sample.h:254-260)
struct Random {
    fx_rng s;
    explicit Random(uint32_t seed = 0) : s(fx_rng_init(seed)) {}
    float Randf() { return fx_rng_f(&s); }
    float Randf01() { float t = Randf(); return (1.0f - t) * 0.0f + t * 1.0f; }   // Randf(min,max), maths.h:213-217
};
inline void Sample2D(Random& r, float& u1, float& u2) { u1 = r.Randf01(); u2 = r.Randf01(); }

// ---------------------------------------------------------------- samplers (maths.h:94-108,229-276)
inline void BasisFromVector(f3 w, f3* u, f3* v)
{
    if (fabsf(w.x) > fabsf(w.y)) {
        float invLen = (float)(1.0 / (double)sqrtf(w.x * w.x + w.z * w.z));
        *u = mk3(-w.z * invLen, 0.0f, w.x * invLen);
    } else {
        float invLen = (float)(1.0 / (double)sqrtf(w.y * w.y + w.z * w.z));
        *u = mk3(0.0f, w.z * invLen, -w.y * invLen);
    }
    *v = cross(w, *u);
}
inline f3 SafeNormalize(f3 a)
{
    float m = dot(a, a);
    if (m > 0.0) return a * (float)(1.0 / (double)sqrtf(m));
    return mk3(0.0f);
}
inline f3 UniformSampleHemisphere(Random& rand)
{
    float z = rand.Randf01();
    float w = sqrtf(1.0f - z * z);
    float phi = k2Pi * rand.Randf01();
    return mk3(cosf(phi) * w, sinf(phi) * w, z);
}
inline f3 CosineSampleHemisphere(float u1, float u2)
{
    float r = sqrtf(u1), theta = k2Pi * u2;
    float sx = r * cosf(theta), sy = r * sinf(theta);
    float z = sqrtf(fmaxf(0.0f, 1.0f - sx * sx - sy * sy));      // maxf(a,b) = a > b ? a : b ; same for finite input
    return mk3(sx, sy, z);
}

// ---------------------------------------------------------------- Disney BSDF (Disney.cuh)
enum BSDFType { eReflected, eTransmitted, eSpecular };

inline bool Refract(f3 wi, f3 n, float eta, f3& wt)               // Disney.cuh:36-49
{
    float cosThetaI = dot(n, wi);
    float sin2ThetaI = fmaxf(0.0f, 1.0f - cosThetaI * cosThetaI);
    float sin2ThetaT = eta * eta * sin2ThetaI;
    if (sin2ThetaT >= 1) return false;
    float cosThetaT = sqrtf(1.0f - sin2ThetaT);
    wt = eta * -wi + (eta * cosThetaI - cosThetaT) * n;
    return true;
}
inline float SchlickFresnel(float u)                               // :51-56
{
    float m = clampf(1 - u, 0.0f, 1.0f);
    float m2 = m * m;
    return m2 * m2 * m;
}
inline float GTR1(float NDotH, float a)                            // :58-64
{
    if (a >= 1) return kInvPi;
    float a2 = a * a;
    float t = 1 + (a2 - 1) * NDotH * NDotH;
    return (a2 - 1) / (kPi * logf(a2) * t);
}
inline float GTR2(float NDotH, float a)                            // :66-71
{
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return a2 / (kPi * t * t);
}
inline float SmithGGX(float NDotv, float alphaG)                   // :73-78
{
    float a = alphaG * alphaG, b = NDotv * NDotv;
    return 1 / (NDotv + sqrtf(a + b - a * b));
}
inline float Fr(float VDotN, float etaI, float etaT)               // :81-98
{
    float SinThetaT2 = sqr(etaI / etaT) * (1.0f - VDotN * VDotN);
    if (SinThetaT2 > 1.0f) return 1.0f;
    float LDotN = sqrtf(1.0f - SinThetaT2);
    float eta = etaT / etaI;
    float r1 = (VDotN - eta * LDotN) / (VDotN + eta * LDotN);
    float r2 = (LDotN - eta * VDotN) / (LDotN + eta * VDotN);
    return 0.5f * (sqr(r1) + sqr(r2));
}

inline float BSDFPdf(const Material& mat, float etaI, float etaO, f3 n, f3 V, f3 L)      // :152-193
{
    if (dot(L, n) <= 0.0f) {
        float bsdfPdf = 0.0f;
        float brdfPdf = kInv2Pi * mat.subsurface * 0.5f;
        return lerpf(brdfPdf, bsdfPdf, mat.transmission);
    }
    float F = Fr(dot(n, V), etaI, etaO);
    const float a = fmaxf(0.001f, mat.roughness);
    const f3 half = SafeNormalize(L + V);
    const float cosThetaHalf = fabsf(dot(half, n));
    const float pdfHalf = GTR2(cosThetaHalf, a) * cosThetaHalf;
    float pdfSpec = 0.25f * pdfHalf / fmaxf(1.e-6f, dot(L, half));
    float pdfDiff = fabsf(dot(L, n)) * kInvPi * (1.0f - mat.subsurface);
    float bsdfPdf = pdfSpec * F;
    float brdfPdf = lerpf(pdfDiff, pdfSpec, 0.5f);
    return lerpf(brdfPdf, bsdfPdf, mat.transmission);
}

inline f3 ggx_half(const Material& mat, f3 U, f3 V, f3 N, float r1, float r2)            // :211-219 == :287-301
{
    const float a = fmaxf(0.001f, mat.roughness);
    const float phiHalf = r1 * k2Pi;
    const float cosThetaHalf = sqrtf((1.0f - r2) / (1.0f + (sqr(a) - 1.0f) * r2));
    const float sinThetaHalf = sqrtf(fmaxf(0.0f, 1.0f - sqr(cosThetaHalf)));
    const float sinPhiHalf = sinf(phiHalf);
    const float cosPhiHalf = cosf(phiHalf);
    return U * (sinThetaHalf * cosPhiHalf) + V * (sinThetaHalf * sinPhiHalf) + N * cosThetaHalf;
}

inline void BSDFSample(const Material& mat, float etaI, float etaO, f3 U, f3 V, f3 N, f3 view,
                       f3& light, float& pdf, BSDFType& type, Random& rand)                // :197-315
{
    if (rand.Randf() < mat.transmission) {
        float F = Fr(dot(N, view), etaI, etaO);
        if (rand.Randf() < F) {
            float r1, r2;
            Sample2D(rand, r1, r2);
            f3 half = ggx_half(mat, U, V, N, r1, r2);
            if (dot(half, view) <= 0.0f) half *= -1.0f;
            type = eReflected;
            light = 2.0f * dot(view, half) * half - view;
        } else {
            float eta = etaI / etaO;
            if (Refract(view, N, eta, light)) { type = eSpecular; pdf = (1.0f - F) * mat.transmission; return; }
            pdf = 0.0f;
            return;
        }
    } else {
        float r1, r2;
        Sample2D(rand, r1, r2);
        if (rand.Randf() < 0.5f) {
            if (rand.Randf() < mat.subsurface) {
                const f3 d = UniformSampleHemisphere(rand);
                light = U * d.x + V * d.y - N * d.z;
                type = eTransmitted;
            } else {
                const f3 d = CosineSampleHemisphere(r1, r2);
                light = U * d.x + V * d.y + N * d.z;
                type = eReflected;
            }
        } else {
            f3 half = ggx_half(mat, U, V, N, r1, r2);
            if (dot(half, view) <= 0.0f) half *= -1.0f;
            light = 2.0f * dot(view, half) * half - view;
            type = eReflected;
        }
    }
    pdf = BSDFPdf(mat, etaI, etaO, N, view, light);
}

inline f3 BSDFEval(const Material& mat, f3 albedo, float etaI, float etaO, f3 N, f3 V, f3 L)   // :318-427
{
    float NDotL = dot(N, L), NDotV = dot(N, V);
    f3 H = normalize(L + V);
    float NDotH = dot(N, H), LDotH = dot(L, H);
    f3 Cdlin = albedo;
    float Cdlum = (float)(.3 * Cdlin.x + .6 * Cdlin.y + .1 * Cdlin.z);        // double literals in the reference
    f3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : mk3(1.0f);
    f3 Cspec0 = lerp3((float)(mat.specular * .08) * lerp3(mk3(1.0f), Ctint, mat.specularTint), Cdlin, mat.metallic);
    f3 bsdf = mk3(0.0f), brdf = mk3(0.0f);
    if (mat.transmission > 0.0f) {
        if (NDotL <= 0) {
            float F = Fr(NDotV, etaI, etaO);
            bsdf = mk3(mat.transmission * (1.0f - F) / fabsf(NDotL) * (1.0f - mat.metallic));
        } else {
            float a = fmaxf(0.001f, mat.roughness);
            float Ds = GTR2(NDotH, a);
            float FH = Fr(LDotH, etaI, etaO);
            f3 Fs = lerp3(Cspec0, mk3(1.0f), FH);
            float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
            bsdf = Gs * Fs * Ds;
        }
    }
    if (mat.transmission < 1.0f) {
        if (NDotL <= 0) {
            if (mat.subsurface > 0.0f) {
                f3 s = mk3(sqrtf(mat.color.x), sqrtf(mat.color.y), sqrtf(mat.color.z));
                float FL = SchlickFresnel(fabsf(NDotL)), FV = SchlickFresnel(NDotV);
                float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
                brdf = kInvPi * s * mat.subsurface * Fd * (1.0f - mat.metallic);
            }
        } else {
            float a = fmaxf(0.001f, mat.roughness);
            float Ds = GTR2(NDotH, a);
            float FH = SchlickFresnel(LDotH);
            f3 Fs = lerp3(Cspec0, mk3(1.f), FH);
            float Gs = SmithGGX(NDotV, a) * SmithGGX(NDotL, a);
            float FL = SchlickFresnel(NDotL), FV = SchlickFresnel(NDotV);
            float Fd90 = (float)(0.5 + 2.0f * LDotH * LDotH * mat.roughness);
            float Fd = lerpf(1.0f, Fd90, FL) * lerpf(1.0f, Fd90, FV);
            float Dr = GTR1(NDotH, lerpf(.1f, .001f, mat.clearcoatGloss));
            float Fc = lerpf(.04f, 1.0f, FH);
            float Gr = SmithGGX(NDotL, .25f) * SmithGGX(NDotV, .25f);
            brdf = kInvPi * Fd * Cdlin * (1.0f - mat.metallic) * (1.0f - mat.subsurface) + Gs * Fs * Ds + mk3(mat.clearcoat * Gr * Fc * Dr);
        }
    }
    return lerp3(brdf, bsdf, mat.transmission);
}

// ---------------------------------------------------------------- probe (Probe.cuh, Probe.h)
inline f2 ProbeDirToUV(f3 dir)                                    // Probe.cuh:38-46
{
    float theta = acosf(clampf(dir.y, -1.0f, 1.0f));
    float phi = (dir.x == 0.0f && dir.z == 0.0f) ? 0.0f : atan2f(dir.z, dir.x);
    float u = (kPi + phi) * kInvPi * 0.5f;
    float v = theta * kInvPi;
    return {u, v};
}
inline f3 ProbeUVToDir(f2 uv)                                     // :48-58
{
    float theta = uv.y * kPi;
    float phi = uv.x * 2.0f * kPi;
    return mk3(-sinf(theta) * cosf(phi), cosf(theta), -sinf(theta) * sinf(phi));
}
inline int clampi(int v, int a, int b) { return std::max(a, std::min(v, b)); }
inline f4 ProbeEval(const Probe& p, f2 uv)                        // :61-67
{
    int px = clampi((int)(uv.x * p.width), 0, p.width - 1);
    int py = clampi((int)(uv.y * p.height), 0, p.height - 1);
    return p.data[py * p.width + px];
}
inline int LowerBound(const float* a, int lower, int upper, float value)     // :119-136
{
    while (lower < upper) {
        int mid = lower + (upper - lower) / 2;
        if (a[mid] < value) lower = mid + 1; else upper = mid;
    }
    return lower;
}
inline void ProbeSample(const Probe& p, f3& dir, f3& color, float& pdf, Random& rand)   // :138-169
{
    float r1, r2;
    Sample2D(rand, r1, r2);
    int row = LowerBound(p.cdfY, 0, p.height, r1);
    int col = LowerBound(p.cdfX, row * p.width, (row + 1) * p.width, r2) - row * p.width;
    f4 c = p.data[row * p.width + col];
    color = mk3(c.x, c.y, c.z);
    pdf = p.pdfX[row * p.width + col] * p.pdfY[row];
    float u = col / float(p.width), v = row / float(p.height);
    float sinTheta = sinf(v * kPi);
    if (sinTheta == 0.0f) pdf = 0.0f;
    else pdf *= p.width * p.height / (2.0f * kPi * kPi * sinTheta);
    dir = ProbeUVToDir({u, v});
}

struct HostProbe {                                                 // Probe.h:5-77 ProbeData
    int w = 0, h = 0; f3 offset{0, 0, 0};
    std::vector<f4> data; std::vector<float> pdfX, cdfX, pdfY, cdfY;
    void BuildCDF()
    {
        pdfX.assign((size_t)w * h, 0.f); cdfX.assign((size_t)w * h, 0.f); pdfY.assign(h, 0.f); cdfY.assign(h, 0.f);
        float totalWeightY = 0.0f;
        for (int j = 0; j < h; ++j) {
            float totalWeightX = 0.0f;
            for (int i = 0; i < w; ++i) {
                const f4& c = data[(size_t)j * w + i];
                float weight = c.x * 0.3f + c.y * 0.6f + c.z * 0.1f;     // Luminance, maths.h:163-166
                totalWeightX += weight;
                pdfX[(size_t)j * w + i] = weight;
                cdfX[(size_t)j * w + i] = totalWeightX;
            }
            float inv = 1.0f / totalWeightX;
            for (int i = 0; i < w; ++i) { pdfX[(size_t)j * w + i] *= inv; cdfX[(size_t)j * w + i] *= inv; }
            totalWeightY += totalWeightX;
            pdfY[j] = totalWeightX;
            cdfY[j] = totalWeightY;
        }
        for (int j = 0; j < h; ++j) { cdfY[j] /= totalWeightY; pdfY[j] /= totalWeightY; }
    }
    Probe view()
    {
        Probe p; p.width = w; p.height = h; p.data = data.data(); p.offset = offset;
        p.pdfX = pdfX.data(); p.cdfX = cdfX.data(); p.pdfY = pdfY.data(); p.cdfY = cdfY.data();
        return p;
    }
};

// ---------------------------------------------------------------- output (deviceProgram.cu:126-131, cuda/helpers.h:35-62)
inline float srgb1(float c, float powed) { return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f; }
inline uint32_t make_color(f3 c)
{
    f3 cc = mk3(clampf(c.x, 0.0f, 1.0f), clampf(c.y, 0.0f, 1.0f), clampf(c.z, 0.0f, 1.0f));
    float invGamma = 1.0f / 2.4f;
    f3 powed = mk3(powf(cc.x, invGamma), powf(cc.y, invGamma), powf(cc.z, invGamma));
    f3 s = mk3(srgb1(cc.x, powed.x), srgb1(cc.y, powed.y), srgb1(cc.z, powed.z));
    auto q = [](float x) { x = clampf(x, 0.0f, 1.0f); return (uint32_t)std::min((unsigned int)(x * 256.0f), 255u); };
    return q(s.x) | (q(s.y) << 8) | (q(s.z) << 16) | (255u << 24);
}
inline uint32_t tonemap(f3 accum, float exposure)
{
    f3 e = accum * exposure;
    const float luminance = 0.2126f * e.x + 0.7152f * e.y + 0.0722f * e.z;
    f3 m = (e * 1.0f) / (1.0f + luminance / 1.0f);
    return make_color(m);
}

}  // namespace

// ---------------------------------------------------------------- context
struct orc_ctx {
    std::vector<ocommon::HostMesh> meshes;
    std::vector<ocommon::HostTexture> textures;
    obvh::Scene scene;
    obvh::Counters ctr;
    std::atomic<uint64_t> samples{0};
    HostProbe probe;
    bool has_probe = false;
    std::string err;
    std::vector<float> prev_accum;          // snapshot of the accumulation buffer at frame start (extension mode)
    bool prev_accum_active = false;
    // orc_opts.filter: what the passes of one frame leave behind for the reconstruction pass of orc_render
    std::vector<f4> li_color;               // per launch index of the pass in flight: colour before accumulation
    std::vector<std::atomic<uint32_t>> owner;   // per pixel: (pass << 28 | launch index) of the surviving write, kNoOwner = none
    int pass_in_flight = -1;
};

namespace {

struct PRD {                                                       // RadiancePRD, deviceProgram.cu:60-89
    f3 radiance, alpha, origin, direction;
    float bsdfPdf; f3 pathThroughput; float rayEta;
    int depth; bool done, secondary;
    Random rand;
    f3 normal, albedo;                                               // denoiser features of the last shaded hit (:69-70)
};

struct Frame {                                                     // one launch in flight
    orc_ctx* c; const LaunchParams* lp; const orc_opts* o; Probe probe;
    // extension mode only (every ring accumulates, orc_opts.seed_all_passes_with_subframe): the accumulation
    // buffer as the PREVIOUS frame left it.  Each pixel is then lerped exactly once per frame, against last
    // frame's value, by whichever pass writes it last -- a running mean of per-frame estimates.  Reading the
    // live buffer instead (the literal vmv2 statement, used when this is null) would lerp the thin rings
    // where two passes overlap twice in one frame.  See DESIGN.md "accumulation".
    const f4* prev_accum = nullptr;
};

// SampleLights / SampleShadow, deviceProgram.cu:303-387 (kProbeSamples = kBsdfSamples = 1)
f3 sample_lights(const Frame& F, const Material& mat, f3 albedo, float etaI, float etaO, f3 P, f3 N, f3 wo,
                 Random& rand, bool want_occluded)
{
    f3 sum = mk3(0.0f);
    f3 skyColor, wi; float skyPdf;
    ProbeSample(F.probe, wi, skyColor, skyPdf, rand);
    fx_vec3 O = {P.x, P.y, P.z}, D = {wi.x, wi.y, wi.z};
    bool occluded = F.c->scene.any_front(O, D, F.o->tmin, F.o->tmax, &F.c->ctr);
    if (occluded == want_occluded) {
        float bsdfPdf = BSDFPdf(mat, etaI, etaO, N, wo, wi);
        f3 f = BSDFEval(mat, albedo, etaI, etaO, N, wo, wi);
        if (bsdfPdf > 0.0f) {
            int Nn = 2;
            float cbsdf = 1.f / Nn, csky = float(1.f) / Nn;
            float weight = csky * skyPdf / (cbsdf * bsdfPdf + csky * skyPdf);
            if (weight > 0.0f) {
                f3 val = weight * skyColor * f * fabsf(dot(wi, N)) / skyPdf * (1.0f / 1.f);
                sum += val;
            }
        }
    }
    return sum;
}

// __closesthit__radiance, deviceProgram.cu:619-732
void closest_hit(const Frame& F, PRD& prd, const obvh::Hit& hit, f3 ray_org, f3 ray_dir)
{
    const ocommon::HostMesh& mesh = F.c->meshes[(size_t)hit.mesh];
    Material mat; memcpy(&mat, mesh.material, sizeof(Material));
    const uint32_t* ix = &mesh.idx[3 * (size_t)hit.prim];
    auto vert = [&](uint32_t i) { return mk3(mesh.vtx[3 * i], mesh.vtx[3 * i + 1], mesh.vtx[3 * i + 2]); };
    const f3 v0 = vert(ix[0]), v1 = vert(ix[1]), v2 = vert(ix[2]);
    const f3 N_0 = normalize(cross(v1 - v0, v2 - v0));
    f3 N = faceforward(N_0, -ray_dir, N_0);
    const f3 P = ray_org + hit.t * ray_dir;

    if ((mat.flags & kShadowCatcher) != 0 && prd.secondary) {       // :646-651
        prd.origin = P; prd.direction = ray_dir; --prd.depth; return;
    }
    prd.normal = N;                                                  // :653
    f3 albedo = mat.color;
    if (mesh.texture_id >= 0 && mesh.texture_id < (int)F.c->textures.size() && !mesh.tc.empty()) {   // :655-665
        const float u = hit.u, v = hit.v;
        auto tcv = [&](uint32_t i) { return f2{mesh.tc[2 * i], mesh.tc[2 * i + 1]}; };
        f2 t0 = tcv(ix[0]), t1 = tcv(ix[1]), t2 = tcv(ix[2]);
        float w0 = 1.f - u - v;
        f2 tc = {w0 * t0.x + u * t1.x + v * t2.x, w0 * t0.y + u * t1.y + v * t2.y};
        float px[4];
        ocommon::tex_bilinear(F.c->textures[(size_t)mesh.texture_id], tc.x, tc.y, px);
        albedo = mk3(px[0], px[1], px[2]);
    }
    prd.albedo = albedo;                                             // :654, :665
    float outEta;
    if (prd.rayEta == 1.0f) outEta = mat.ior(); else outEta = 1.0f;  // :673-683 (absorption is tracked but never applied)

    if ((mat.flags & kShadowCatcher) == 0) {                         // :686-694
        f3 lightSample = sample_lights(F, mat, albedo, prd.rayEta, outEta, P, N, -ray_dir, prd.rand, false);
        prd.radiance += prd.pathThroughput * lightSample;
        prd.alpha = mk3(1.0f);
    } else {
        f3 shadowSample = sample_lights(F, mat, albedo, prd.rayEta, outEta, P, N, -ray_dir, prd.rand, true);
        prd.alpha += prd.pathThroughput * shadowSample;
    }
    if (!prd.secondary) prd.radiance += mat.emission;                // :696-698

    f3 u, v;
    BasisFromVector(N, &u, &v);
    f3 bsdfDir = mk3(0.f); BSDFType type = eReflected;
    BSDFSample(mat, prd.rayEta, outEta, u, v, N, -ray_dir, bsdfDir, prd.bsdfPdf, type, prd.rand);
    if (prd.bsdfPdf <= 0.0f) { prd.done = true; return; }            // :708-711
    f3 f = BSDFEval(mat, albedo, prd.rayEta, outEta, N, -ray_dir, bsdfDir);
    if (dot(bsdfDir, N) <= 0.0f) prd.rayEta = outEta;                // :717-721
    prd.pathThroughput *= f * fabsf(dot(N, bsdfDir)) / prd.bsdfPdf;  // :724
    prd.direction = bsdfDir; prd.origin = P; prd.secondary = true;   // :727-731
}

// __raygen__renderFrame for one launch index, deviceProgram.cu:392-617
bool raygen(const Frame& F, unsigned lx, unsigned ly, unsigned lw)
{
    const LaunchParams& p = *F.lp;
    const orc_opts& o = *F.o;
    const int w = p.frame.size_x, h = p.frame.size_y;
    const unsigned subframe_index = p.frame.subframe_index;
    const int spp = (int)p.samples_per_launch;
    uint32_t seed = fx_tea4(ly * (unsigned)w + lx, subframe_index);   // :411
    uint32_t px, py;
    fx_launch_to_pixel(lx, ly, p.frame.factor_x, p.frame.factor_y, p.frame.offset_x, p.frame.offset_y, &px, &py);   // :433
    float range = fx_gaze_range(px, py, p.frame.c_x, p.frame.c_y);    // :435
    const size_t li = (size_t)ly * lw + lx;
    if (!fx_in_ring(range, p.frame.r_inner, p.frame.r_outer)) {       // :437-440
        if (o.primary_tri) o.primary_tri[li] = -2;
        return false;
    }
    fx_vec3 U = {p.camera.U.x, p.camera.U.y, p.camera.U.z}, V = {p.camera.V.x, p.camera.V.y, p.camera.V.z},
            Wv = {p.camera.W.x, p.camera.W.y, p.camera.W.z};
    f3 result = mk3(0.f), alpha = mk3(0.f), backplate = mk3(0.f);
    f3 normal = mk3(0.f), albedo = mk3(0.f);                          // :443-444
    int i = spp;
    bool first = true;
    do {
        f3 direct = mk3(0.f), indirect = mk3(0.f);
        PRD prd;
        prd.radiance = mk3(0.f); prd.alpha = mk3(0.f);
        prd.rand = Random(seed);                                       // :464, BEFORE the jitter advances seed
        prd.rayEta = 1.0f; prd.pathThroughput = mk3(1.f); prd.bsdfPdf = 1.0f;
        prd.done = false; prd.secondary = false; prd.depth = 0;
        prd.normal = mk3(0.f); prd.albedo = mk3(0.f);                  // :471-472
        float jx, jy;                                                  // :479 (argument evaluation order)
        if (o.jitter_yx) { jy = fx_rnd(&seed); jx = fx_rnd(&seed); } else { jx = fx_rnd(&seed); jy = fx_rnd(&seed); }
        fx_vec3 d = fx_camera_dir(px, py, jx, jy, w, h, U, V, Wv);     // :483-491
        f3 ray_direction = mk3(d.x, d.y, d.z), ray_origin = p.camera.eye;
        f4 bp = ProbeEval(F.probe, ProbeDirToUV(ray_direction));       // :495
        backplate = mk3(bp.x, bp.y, bp.z);
        for (;;) {                                                     // :497-534
            prd.radiance = mk3(0.f);
            obvh::Hit hit;
            fx_vec3 O = {ray_origin.x, ray_origin.y, ray_origin.z}, D = {ray_direction.x, ray_direction.y, ray_direction.z};
            bool ok = F.c->scene.closest(O, D, o.tmin, o.tmax, &hit, &F.c->ctr);
            if (first) {
                first = false;
                if (o.primary_tri) o.primary_tri[li] = ok ? (int32_t)hit.tri : -1;
                if (o.primary_tuv) { float* q = o.primary_tuv + 3 * li; q[0] = ok ? hit.t : 0.f; q[1] = ok ? hit.u : 0.f; q[2] = ok ? hit.v : 0.f; }
                if (o.primary_dir) { float* q = o.primary_dir + 3 * li; q[0] = d.x; q[1] = d.y; q[2] = d.z; }
            }
            if (ok) closest_hit(F, prd, hit, ray_origin, ray_direction);
            else { prd.done = true; prd.normal = mk3(0.f); prd.albedo = mk3(0.f); }      // __miss__radiance :253-282
            // :509-512.  A secondary ray that passes through a shadow catcher decrements depth (:650) and leaves prd.normal
            // untouched, so a pass-through at depth 1 adds the first hit's features once more: kept as written.
            if (prd.depth == 0) { normal += prd.normal; albedo += prd.albedo; }
            if (prd.done || prd.depth >= o.max_depth) break;           // :515
            if (prd.depth == 0) direct += prd.radiance; else indirect += prd.radiance;
            ++prd.depth;
            ray_origin = prd.origin; ray_direction = prd.direction;
        }
        result += direct + indirect;                                   // :536-537
        alpha += prd.alpha;
    } while (--i);
    normal /= (float)spp;                                              // :541-543
    albedo /= (float)spp;
    alpha /= (float)spp;
    if (!F.c->li_color.empty()) {                                      // orc_opts.filter: keep the launch index's colour, unaccumulated
        const f3 c0 = ((backplate * (float)spp) * (1.0f - alpha) + result) / (float)spp;
        F.c->li_color[li] = {c0.x, c0.y, c0.z, 1.0f};
    }

    for (int a = 0; a < p.frame.fillSize; ++a) {                       // :546-616
        for (int b = 0; b < p.frame.fillSize; ++b) {
            unsigned ix = lx * p.frame.factor_x + (unsigned)a + p.frame.offset_x;
            unsigned iy = ly * p.frame.factor_y + (unsigned)b + p.frame.offset_y;
            ix = std::max(0u, std::min(ix, (unsigned)(w - 1)));        // clamp(uint2, 0, (w-1,h-1))
            iy = std::max(0u, std::min(iy, (unsigned)(h - 1)));
            const unsigned image_index = iy * (unsigned)w + ix;
            f3 color = (backplate * (float)spp) * (1.0f - alpha) + result;
            f3 accum_color = color / (float)spp;
            if (o.accumulate && subframe_index > 0 && !p.frame.redraw) {     // vmv2 :545-553 (commented out in sv5 :565-581)
                accum_color = mk3(clampf(accum_color.x, 0.f, 10.f), clampf(accum_color.y, 0.f, 10.f), clampf(accum_color.z, 0.f, 10.f));
                const float a_w = 1.0f / (float)(subframe_index + 1);
                const f4 prev4 = F.prev_accum ? F.prev_accum[image_index] : p.frame.accum_buffer[image_index];
                accum_color = lerp3(mk3(prev4.x, prev4.y, prev4.z), accum_color, a_w);
            }
            p.frame.accum_buffer[image_index] = {accum_color.x, accum_color.y, accum_color.z, 1.0f};
            p.frame.frame_buffer[image_index] = tonemap(accum_color, o.exposure);
            if (o.features) {                                          // :612-614 (commented out in sv5)
                if (p.frame.normal_buffer) p.frame.normal_buffer[image_index] = {normal.x, normal.y, normal.z, 1.0f};
                if (p.frame.color_buffer) p.frame.color_buffer[image_index] = {accum_color.x, accum_color.y, accum_color.z, 1.0f};
                if (p.frame.albedo_buffer) p.frame.albedo_buffer[image_index] = {albedo.x, albedo.y, albedo.z, 1.0f};
            }
            if (!F.c->owner.empty()) {                                 // surviving write = last pass, row-major last launch index
                const uint32_t me = ((uint32_t)F.c->pass_in_flight << 28) | (uint32_t)li;
                std::atomic<uint32_t>& w8 = F.c->owner[image_index];
                uint32_t cur = w8.load(std::memory_order_relaxed);
                while ((cur == 0xffffffffu || cur < me) && !w8.compare_exchange_weak(cur, me, std::memory_order_relaxed)) {}
            }
        }
    }
    return true;
}

}  // namespace

// ---------------------------------------------------------------- C interface
extern "C" {

const char* orc_kind(void) { return "port"; }
orc_ctx* orc_create(void) { return new orc_ctx(); }
void orc_destroy(orc_ctx* c) { delete c; }
const char* orc_last_error(orc_ctx* c) { return c->err.c_str(); }

int orc_add_texture(orc_ctx* c, const uint32_t* rgba8, int w, int h)
{
    ocommon::HostTexture t; t.w = w; t.h = h; t.px.assign(rgba8, rgba8 + (size_t)w * h);
    c->textures.push_back(std::move(t));
    return (int)c->textures.size() - 1;
}
int orc_add_mesh(orc_ctx* c, const float* vtx, int nvtx, const uint32_t* idx, int ntri, const float* tc,
                 const void* material104, int texture_id)
{
    for (int i = 0; i < 3 * ntri; ++i) if (idx[i] >= (uint32_t)nvtx) { c->err = "index out of range"; return -1; }
    ocommon::HostMesh m;
    m.vtx.assign(vtx, vtx + 3 * (size_t)nvtx); m.idx.assign(idx, idx + 3 * (size_t)ntri);
    if (tc) m.tc.assign(tc, tc + 2 * (size_t)nvtx);
    memcpy(m.material, material104, 104); m.texture_id = texture_id;
    c->meshes.push_back(std::move(m));
    return (int)c->meshes.size() - 1;
}
int orc_load_obj(orc_ctx* c, const char*)
{
    c->err = "the port has no OBJ reader: loader parity is product (host/Model.cpp) vs oracle/_ref and tests/golden";
    return -1;
}
int orc_add_box(orc_ctx* c, const void* material104, const float pos[3], const float e[3])   // addBox, Model.cpp:219-291
{
    float X[2] = {-e[0] + pos[0], e[0] + pos[0]}, Y[2] = {-e[1] + pos[1], e[1] + pos[1]}, Z[2] = {-e[2] + pos[2], e[2] + pos[2]};
    float A[3] = {X[0], Y[0], Z[1]}, B[3] = {X[1], Y[0], Z[1]}, Cc[3] = {X[1], Y[1], Z[1]}, D[3] = {X[0], Y[1], Z[1]};
    float E[3] = {X[0], Y[0], Z[0]}, Fp[3] = {X[1], Y[0], Z[0]}, G[3] = {X[1], Y[1], Z[0]}, H[3] = {X[0], Y[1], Z[0]};
    const float* order[36] = {A, B, Cc, A, Cc, D, E, H, G, E, G, Fp, E, A, D, E, D, H, B, Fp, G, B, G, Cc, D, Cc, G, D, G, H, E, A, B, E, B, Fp};
    ocommon::HostMesh m;
    for (int i = 0; i < 36; ++i) { m.vtx.insert(m.vtx.end(), order[i], order[i] + 3); m.tc.push_back(0.f); m.tc.push_back(0.f); m.idx.push_back((uint32_t)i); }
    memcpy(m.material, material104, 104);
    m.texture_id = 1;    // TriangleMesh::diffuseTextureID default, Model.h:19 (addBox never sets it)
    c->meshes.push_back(std::move(m));
    return (int)c->meshes.size() - 1;
}
int orc_mesh_count(orc_ctx* c) { return (int)c->meshes.size(); }
int orc_mesh_info(orc_ctx* c, int i, int* nvtx, int* ntri, int* ntc, int* tex)
{
    if (i < 0 || i >= (int)c->meshes.size()) return -1;
    auto& m = c->meshes[(size_t)i];
    *nvtx = (int)m.vtx.size() / 3; *ntri = (int)m.idx.size() / 3; *ntc = (int)m.tc.size() / 2; *tex = m.texture_id;
    return 0;
}
int orc_mesh_get(orc_ctx* c, int i, float* vtx, uint32_t* idx, float* tc, void* material104)
{
    if (i < 0 || i >= (int)c->meshes.size()) return -1;
    auto& m = c->meshes[(size_t)i];
    if (vtx) memcpy(vtx, m.vtx.data(), m.vtx.size() * 4);
    if (idx) memcpy(idx, m.idx.data(), m.idx.size() * 4);
    if (tc && !m.tc.empty()) memcpy(tc, m.tc.data(), m.tc.size() * 4);
    if (material104) memcpy(material104, m.material, 104);
    return 0;
}
int orc_texture_count(orc_ctx* c) { return (int)c->textures.size(); }
int orc_texture_info(orc_ctx* c, int i, int* w, int* h)
{
    if (i < 0 || i >= (int)c->textures.size()) return -1;
    *w = c->textures[(size_t)i].w; *h = c->textures[(size_t)i].h; return 0;
}
int orc_texture_get(orc_ctx* c, int i, uint32_t* rgba8)
{
    if (i < 0 || i >= (int)c->textures.size()) return -1;
    memcpy(rgba8, c->textures[(size_t)i].px.data(), c->textures[(size_t)i].px.size() * 4); return 0;
}
int orc_commit(orc_ctx* c)
{
    c->scene.clear();
    for (auto& m : c->meshes) c->scene.add_mesh(m.vtx.data(), m.idx.data(), m.idx.size() / 3);
    c->scene.commit();
    return 0;
}

int orc_set_probe(orc_ctx* c, int w, int h, const float* rgba, const float offset[3])
{
    c->probe.w = w; c->probe.h = h; c->probe.offset = mk3(offset[0], offset[1], offset[2]);
    c->probe.data.resize((size_t)w * h);
    memcpy(c->probe.data.data(), rgba, (size_t)w * h * 16);
    c->probe.BuildCDF();
    c->has_probe = true;
    return 0;
}
int orc_get_probe(orc_ctx* c, float* pdfX, float* cdfX, float* pdfY, float* cdfY)
{
    if (!c->has_probe) return -1;
    size_t n = (size_t)c->probe.w * c->probe.h;
    memcpy(pdfX, c->probe.pdfX.data(), n * 4); memcpy(cdfX, c->probe.cdfX.data(), n * 4);
    memcpy(pdfY, c->probe.pdfY.data(), (size_t)c->probe.h * 4); memcpy(cdfY, c->probe.cdfY.data(), (size_t)c->probe.h * 4);
    return 0;
}

int orc_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy, float aspect, float o[9])
{   // sutil/Camera.cpp:34-46
    f3 W = mk3(lookat[0], lookat[1], lookat[2]) - mk3(eye[0], eye[1], eye[2]);
    float wlen = sqrtf(dot(W, W));
    f3 U = normalize(cross(W, mk3(up[0], up[1], up[2])));
    f3 V = normalize(cross(U, W));
    float vlen = wlen * tanf(0.5f * fovy * 3.14159265358979323846f / 180.0f);
    V *= vlen;
    float ulen = vlen * aspect;
    U *= ulen;
    o[0] = U.x; o[1] = U.y; o[2] = U.z; o[3] = V.x; o[4] = V.y; o[5] = V.z; o[6] = W.x; o[7] = W.y; o[8] = W.z;
    return 0;
}

int orc_launch(orc_ctx* c, void* lp248, unsigned lw, unsigned lh, const orc_opts* o)
{
    if (!c->has_probe) { c->err = "no probe"; return -1; }
    LaunchParams* lp = (LaunchParams*)lp248;
    lp->traversable = 1;
    Frame F{c, lp, o, c->probe.view()};
    F.prev_accum = c->prev_accum_active ? reinterpret_cast<const f4*>(c->prev_accum.data()) : nullptr;
    lp->probe = F.probe;
    std::atomic<uint64_t> nsamp{0};
    ocommon::parallel_rows(lh, ocommon::resolve_threads(o->threads), [&](unsigned y) {
        uint64_t local = 0;
        for (unsigned x = 0; x < lw; ++x) if (raygen(F, x, y, lw)) local += lp->samples_per_launch;
        nsamp += local;
    });
    c->samples += nsamp;
    return 0;
}

int orc_render(orc_ctx* c, void* lp248, const orc_fov_cfg* cfg, const orc_opts* o)    // SimplePathtracer.cpp:77-214
{
    LaunchParams* lp = (LaunchParams*)lp248;
    if (lp->frame.size_x == 0) return 0;
    auto passes = ocommon::three_pass_schedule(*cfg, lp->frame.size_x, lp->frame.size_y, lp->frame.c_x, lp->frame.c_y);
    unsigned saved = lp->frame.subframe_index;
    c->prev_accum_active = false;
    if (o->seed_all_passes_with_subframe && o->accumulate && saved > 0) {
        const size_t nfl = (size_t)lp->frame.size_x * lp->frame.size_y * 4;
        c->prev_accum.assign((const float*)lp->frame.accum_buffer, (const float*)lp->frame.accum_buffer + nfl);
        c->prev_accum_active = true;
    }
    orc_opts local = *o;
    local.primary_tri = nullptr; local.primary_tuv = nullptr; local.primary_dir = nullptr;
    const bool filter = o->filter != 0;
    const int w = lp->frame.size_x, h = lp->frame.size_y;
    std::vector<std::vector<f4>> li_colors;
    std::vector<f4> before;                                            // the accumulation buffer as the previous frame left it
    if (filter) {
        before.assign(lp->frame.accum_buffer, lp->frame.accum_buffer + (size_t)w * h);
        c->owner = std::vector<std::atomic<uint32_t>>((size_t)w * h);
        for (auto& a : c->owner) a.store(0xffffffffu, std::memory_order_relaxed);
    }
    std::vector<unsigned> pass_sub;
    for (auto& p : passes) {
        if (filter) { c->pass_in_flight = (int)li_colors.size(); c->li_color.assign((size_t)p.lw * p.lh, f4{0.f, 0.f, 0.f, 0.f}); }
        lp->frame.factor_x = lp->frame.factor_y = p.factor; lp->frame.factor_z = 1;
        lp->frame.fillSize = p.fill;
        lp->frame.r_inner = p.r_inner; lp->frame.r_outer = p.r_outer;
        lp->samples_per_launch = p.spp;
        lp->frame.offset_x = p.off_x; lp->frame.offset_y = p.off_y;
        lp->frame.redraw = o->seed_all_passes_with_subframe ? 0u : p.redraw;
        if (p.zero_subframe && !o->seed_all_passes_with_subframe) lp->frame.subframe_index = 0;
        else lp->frame.subframe_index = saved;
        pass_sub.push_back(lp->frame.subframe_index);
        if (orc_launch(c, lp, p.lw, p.lh, &local)) return -1;
        if (filter) li_colors.push_back(std::move(c->li_color));
    }
    if (filter) {
        // Reconstruction pass ("reconstruction filter"): every pixel takes the colour of the launch index
        // whose write survived; in a sub-sampled ring (fillSize > 1) it is blended with the three neighbouring launch indices of
        // the SAME pass on the side of the block the pixel lies on, weights = distance from the block centre in block units.
        // A neighbour takes part only if it is inside the launch and itself owns at least one pixel of the frame.
        c->li_color.clear(); c->pass_in_flight = -1;
        std::vector<std::vector<uint8_t>> owns(passes.size());
        for (size_t p = 0; p < passes.size(); ++p) owns[p].assign((size_t)passes[p].lw * passes[p].lh, 0);
        for (size_t i = 0; i < (size_t)w * h; ++i) { const uint32_t ow = c->owner[i].load(); if (ow != 0xffffffffu) owns[ow >> 28][ow & 0x0fffffffu] = 1; }
        for (int Y = 0; Y < h; ++Y) for (int X = 0; X < w; ++X) {
            const size_t i = (size_t)Y * w + X;
            const uint32_t ow = c->owner[i].load();
            if (ow == 0xffffffffu) { lp->frame.accum_buffer[i] = before[i]; continue; }      // nobody wrote it: untouched
            const size_t pi = ow >> 28; const uint32_t li = ow & 0x0fffffffu;
            const auto& P = passes[pi];
            const std::vector<f4>& col = li_colors[pi];
            const f3 own = mk3(col[li].x, col[li].y, col[li].z);
            f3 cur = own;
            if (P.fill > 1) {
                const int lx = (int)(li % P.lw), ly = (int)(li / P.lw);
                const float half = 0.5f * (float)(P.fill - 1);
                const float dx = ((float)X - ((float)(lx * (int)P.factor + (int)P.off_x) + half)) / (float)P.factor;
                const float dy = ((float)Y - ((float)(ly * (int)P.factor + (int)P.off_y) + half)) / (float)P.factor;
                const int sx = dx < 0.f ? -1 : 1, sy = dy < 0.f ? -1 : 1;
                const float wx = fminf(fabsf(dx), 1.f), wy = fminf(fabsf(dy), 1.f);
                auto fetch = [&](int ax, int ay) {
                    if (ax < 0 || ay < 0 || ax >= (int)P.lw || ay >= (int)P.lh) return own;
                    const size_t j = (size_t)ay * P.lw + (size_t)ax;
                    if (!owns[pi][j]) return own;
                    return mk3(col[j].x, col[j].y, col[j].z);
                };
                const f3 c10 = fetch(lx + sx, ly), c01 = fetch(lx, ly + sy), c11 = fetch(lx + sx, ly + sy);
                cur = lerp3(lerp3(own, c10, wx), lerp3(c01, c11, wx), wy);
            }
            const bool redraw = o->seed_all_passes_with_subframe ? false : P.redraw != 0;
            if (o->accumulate && pass_sub[pi] > 0 && !redraw) {
                cur = mk3(clampf(cur.x, 0.f, 10.f), clampf(cur.y, 0.f, 10.f), clampf(cur.z, 0.f, 10.f));
                cur = lerp3(mk3(before[i].x, before[i].y, before[i].z), cur, 1.0f / (float)(pass_sub[pi] + 1));
            }
            lp->frame.accum_buffer[i] = {cur.x, cur.y, cur.z, 1.0f};
            lp->frame.frame_buffer[i] = tonemap(cur, o->exposure);
            if (o->features && lp->frame.color_buffer) lp->frame.color_buffer[i] = {cur.x, cur.y, cur.z, 1.0f};
        }
        c->owner.clear();
    }
    lp->frame.subframe_index = saved + 1;
    c->prev_accum_active = false;
    return 0;
}

void orc_get_counters(orc_ctx* c, uint64_t out[8])
{
    out[0] = c->ctr.closest_rays; out[1] = c->ctr.any_rays; out[2] = c->ctr.closest_nodes; out[3] = c->ctr.closest_tris;
    out[4] = c->ctr.any_nodes; out[5] = c->ctr.any_tris; out[6] = c->samples; out[7] = 0;
}
void orc_reset_counters(orc_ctx* c) { c->ctr.reset(); c->samples = 0; }

int orc_trace(orc_ctx* c, const float* rays, int n, int mode, int32_t* tri, float* tuv)
{
    ocommon::parallel_rows((unsigned)((n + 1023) / 1024), ocommon::resolve_threads(0), [&](unsigned blk) {
        int b = (int)blk * 1024, e = std::min(n, b + 1024);
        for (int i = b; i < e; ++i) {
            const float* r = rays + 8 * (size_t)i;
            fx_vec3 O = {r[0], r[1], r[2]}, D = {r[4], r[5], r[6]};
            if (mode == 0 || mode == 2) {
                obvh::Hit h;
                bool ok = mode == 0 ? c->scene.closest(O, D, r[3], r[7], &h) : c->scene.closest_brute(O, D, r[3], r[7], &h);
                tri[i] = ok ? (int32_t)h.tri : -1;
                if (tuv) { tuv[3 * i] = ok ? h.t : 0.f; tuv[3 * i + 1] = ok ? h.u : 0.f; tuv[3 * i + 2] = ok ? h.v : 0.f; }
            } else {
                bool occ = mode == 1 ? c->scene.any_front(O, D, r[3], r[7]) : c->scene.any_front_brute(O, D, r[3], r[7]);
                tri[i] = occ ? 1 : 0;
            }
        }
    });
    return 0;
}

uint32_t orc_tea4(uint32_t a, uint32_t b) { return fx_tea4(a, b); }
void orc_rnd_seq(uint32_t seed, int n, float* out, uint32_t* seed_out)
{
    for (int i = 0; i < n; ++i) out[i] = fx_rnd(&seed);
    *seed_out = seed;
}
void orc_random_seq(uint32_t seed, int n, uint32_t* u, float* f)
{
    fx_rng a = fx_rng_init(seed), b = fx_rng_init(seed);
    for (int i = 0; i < n; ++i) { if (u) u[i] = fx_rng_u32(&a); if (f) f[i] = fx_rng_f(&b); }
}
void orc_bsdf_sample(const void* material104, const float albedo[3], float etaI, float etaO,
                     const float Nn[3], const float Vv[3], uint32_t seed, float out[10])
{
    Material mat; memcpy(&mat, material104, 104);
    f3 N = mk3(Nn[0], Nn[1], Nn[2]), V = mk3(Vv[0], Vv[1], Vv[2]), u, v, L = mk3(0.f);
    BasisFromVector(N, &u, &v);
    Random rand(seed), probe(seed);
    float pdf = 0.f; BSDFType type = eReflected;
    BSDFSample(mat, etaI, etaO, u, v, N, V, L, pdf, type, rand);
    int draws = 0;
    for (; draws < 16; ++draws) { if (probe.s.s1 == rand.s.s1 && probe.s.s2 == rand.s.s2) break; fx_rng_u32(&probe.s); }
    out[0] = L.x; out[1] = L.y; out[2] = L.z; out[3] = pdf; out[4] = (float)type;
    if (pdf > 0.f) {
        f3 f = BSDFEval(mat, mk3(albedo[0], albedo[1], albedo[2]), etaI, etaO, N, V, L);
        out[5] = f.x; out[6] = f.y; out[7] = f.z; out[8] = BSDFPdf(mat, etaI, etaO, N, V, L);
    } else { out[5] = out[6] = out[7] = out[8] = 0.f; }
    out[9] = (float)draws;
}
void orc_bsdf_eval(const void* material104, const float albedo[3], float etaI, float etaO,
                   const float Nn[3], const float Vv[3], const float Ll[3], float out[4])
{
    Material mat; memcpy(&mat, material104, 104);
    f3 N = mk3(Nn[0], Nn[1], Nn[2]), V = mk3(Vv[0], Vv[1], Vv[2]), L = mk3(Ll[0], Ll[1], Ll[2]);
    f3 f = BSDFEval(mat, mk3(albedo[0], albedo[1], albedo[2]), etaI, etaO, N, V, L);
    out[0] = f.x; out[1] = f.y; out[2] = f.z; out[3] = BSDFPdf(mat, etaI, etaO, N, V, L);
}
int orc_probe_sample(orc_ctx* c, uint32_t seed, float out[7])
{
    if (!c->has_probe) return -1;
    Probe p = c->probe.view();
    Random rand(seed); f3 dir, col; float pdf;
    ProbeSample(p, dir, col, pdf, rand);
    out[0] = dir.x; out[1] = dir.y; out[2] = dir.z; out[3] = col.x; out[4] = col.y; out[5] = col.z; out[6] = pdf;
    return 0;
}
int orc_probe_eval(orc_ctx* c, const float d[3], float out[6])
{
    if (!c->has_probe) return -1;
    Probe p = c->probe.view();
    f2 uv = ProbeDirToUV(mk3(d[0], d[1], d[2]));
    f4 v = ProbeEval(p, uv);
    out[0] = uv.x; out[1] = uv.y; out[2] = v.x; out[3] = v.y; out[4] = v.z; out[5] = v.w;
    return 0;
}
uint32_t orc_tonemap(const float accum[3], float exposure) { return tonemap(mk3(accum[0], accum[1], accum[2]), exposure); }

}  // extern "C"
