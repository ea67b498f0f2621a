// Camera.h -- the pinhole camera the renderer is driven with; same class and UVW construction as
// the reference's sutil::Camera (sutil/Camera.h:37-79, sutil/Camera.cpp:34-46).
#pragma once
#include "fovpt_vec.h"

namespace sutil {

class Camera {
public:
    Camera() : m_eye(make_float3(1.f, 1.f, 1.f)), m_lookat(make_float3(0.f, 0.f, 0.f)), m_up(make_float3(0.f, 1.f, 0.f)), m_fovY(35.0f), m_aspectRatio(1.0f) {}
    Camera(const float3& eye, const float3& lookat, const float3& up, float fovY, float aspectRatio)
        : m_eye(eye), m_lookat(lookat), m_up(up), m_fovY(fovY), m_aspectRatio(aspectRatio) {}

    const float3& eye() const { return m_eye; }
    void setEye(const float3& v) { m_eye = v; }
    const float3& lookat() const { return m_lookat; }
    void setLookat(const float3& v) { m_lookat = v; }
    const float3& up() const { return m_up; }
    void setUp(const float3& v) { m_up = v; }
    const float& fovY() const { return m_fovY; }
    void setFovY(const float& v) { m_fovY = v; }
    const float& aspectRatio() const { return m_aspectRatio; }
    void setAspectRatio(const float& v) { m_aspectRatio = v; }

    // U, V, W: orthogonal, NOT orthonormal; |W| is the focal length, |V| = |W| tan(fovY / 2), |U| = |V| aspect
    void UVWFrame(float3& U, float3& V, float3& W) const
    {
        using namespace fovpt_host;
        W = sub(m_lookat, m_eye);
        const float wlen = length(W);
        U = normalize(cross(W, m_up));
        V = normalize(cross(U, W));
        const float vlen = wlen * tanf(0.5f * m_fovY * 3.14159265358979323846f / 180.0f);
        V = mul(V, vlen);
        U = mul(U, vlen * m_aspectRatio);
    }

private:
    float3 m_eye, m_lookat, m_up;
    float  m_fovY, m_aspectRatio;
};

}  // namespace sutil
