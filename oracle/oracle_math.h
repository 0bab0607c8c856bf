// oracle_math.h -- TEST INFRASTRUCTURE (CPU oracle), never linked into the product.
//
// Scalar float3 arithmetic restating the helpers the reference's device programs call
// from the NVIDIA OptiX SDK 4.1.1 header optixu/optixu_math_namespace.h (NOT present
// under /root/reference; pinned only by reflectcuts.vcxproj:92,139 and README.md:24;
// restated from its published definitions).  Call sites in the
// reference: rtmaterial.cuh:58-60,81,106,122,136-137; lighttracing.cu:115-116,236,292;
// triangleintersect.cu:27.
//
// Parity status: PINNED BY THE REFERENCE'S OWN DEVICE CODE for the CUDA programs -- lighttracing.cu, pathtracing.cu,
// triangleintersect.cu, rtmaterial.cuh, rtmath.cuh, rtlightsource.cuh are compiled UNMODIFIED from /root/reference
// (oracle/ref_device*.cu + oracle/ref_shim/, `make -C oracle refdevice` -> oracle/_ref/libref_device*.so), run on a B200,
// and this oracle must equal what they return: flags / counts / RNG consumption exactly, floats to 1e-5
// (tests/ref_cases.py; live: tests/test_gpu_reference.py; on CPU from the stored outputs: tests/test_ref_goldens.py).
// Still UNPINNED (the reference holds no runnable code or vectors for them): the GLSL stages (photon splat, G-buffer,
// final composite), the host-side schedule and the LVC splatColor loop (lvclighttracing.cu:348-387 uses an MSVC-only cast).
// Also pinned: XORWOW against cuRAND itself (device curand_init), the mt19937 standard vector, detmath against libm.
//
// Compile with -ffp-contract=off: every expression below must round exactly as written.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../evplp_b200/csrc/detmath.h"  // the shared deterministic libm (sinf/cosf/powf/asinf)

namespace orc {

using evplp::det_asinf;
using evplp::det_cosf;
using evplp::det_powf;
using evplp::det_sinf;

static const float M_PIf_ = 3.14159265358979323846f;
static const float M_Inv_PIf = 0.318309886183790671537767526745028724068919291480912897495f;  // rtmath.cuh:12

struct F3 {
    float x, y, z;
};
struct F2 {
    float x, y;
};
inline F3 mk3(float x, float y, float z) { return F3{x, y, z}; }
inline F3 mk3(float s) { return F3{s, s, s}; }
inline F3 operator+(const F3& a, const F3& b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline F3 operator-(const F3& a, const F3& b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 operator-(const F3& a) { return F3{-a.x, -a.y, -a.z}; }
inline F3 operator*(const F3& a, const F3& b) { return F3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline F3 operator*(const F3& a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
inline F3 operator*(float s, const F3& a) { return F3{s * a.x, s * a.y, s * a.z}; }
// optixu: float3 / float multiplies by the reciprocal
inline F3 operator/(const F3& a, float s) {
    float inv = 1.0f / s;
    return a * inv;
}
inline void operator+=(F3& a, const F3& b) { a = a + b; }
inline void operator*=(F3& a, const F3& b) { a = a * b; }
inline void operator/=(F3& a, float s) {
    float inv = 1.0f / s;
    a = a * inv;
}
inline float dot(const F3& a, const F3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 cross(const F3& a, const F3& b) {
    return F3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline F3 normalize(const F3& v) {
    float invLen = 1.0f / sqrtf(dot(v, v));
    return v * invLen;
}
inline F3 reflect(const F3& i, const F3& n) { return i - 2.0f * n * dot(n, i); }
inline F3 faceforward(const F3& n, const F3& i, const F3& nref) { return n * copysignf(1.0f, dot(i, nref)); }
inline F3 fminf3(const F3& a, const F3& b) { return F3{fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline F3 fmaxf3(const F3& a, const F3& b) { return F3{fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }

// optix::Onb
struct Onb {
    F3 m_tangent, m_binormal, m_normal;
    explicit Onb(const F3& normal) {
        m_normal = normal;
        if (fabsf(m_normal.x) > fabsf(m_normal.z)) {
            m_binormal.x = -m_normal.y;
            m_binormal.y = m_normal.x;
            m_binormal.z = 0;
        } else {
            m_binormal.x = 0;
            m_binormal.y = -m_normal.z;
            m_binormal.z = m_normal.y;
        }
        m_binormal = normalize(m_binormal);
        m_tangent = cross(m_binormal, m_normal);
    }
    void inverse_transform(F3& p) const { p = p.x * m_tangent + p.y * m_binormal + p.z * m_normal; }
};

// optix::cosine_sample_hemisphere
inline void cosine_sample_hemisphere(float u1, float u2, F3& p) {
    const float r = sqrtf(u1);
    const float phi = 2.0f * M_PIf_ * u2;
    p.x = r * det_cosf(phi);
    p.y = r * det_sinf(phi);
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}

}  // namespace orc
