// optix_shim.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the parts of the NVIDIA OptiX 4.1.1
// SDK device API that the reference's OptiX programs use, so that the UNMODIFIED reference sources
//   /root/reference/reflectcuts/realtimetechniques/{lighttracing.cu, pathtracing.cu, triangleintersect.cu,
//   rtmaterial.cuh, rtmath.cuh, rtlightsource.cuh, all.cuh, rtcomphoton/rtphotonrecord.h}
// compile with nvcc 12.9 for sm_100a (oracle/Makefile target `refdevice` -> oracle/_ref/libref_device.so).
// Nothing here is part of the product; only tests/ and scripts/make_ref_goldens.py load the result.
//
// The OptiX SDK itself is not in /root/reference (include path reflectcuts.vcxproj:92,139), so the helper
// semantics below are restated from the SDK's published definitions.  How OptiX's
// execution model is mapped:
//   * every OptiX launch index runs as ONE CUDA thread in its OWN block, so the per-thread semantic
//     variables (rtLaunchIndex, rtCurrentRay, rtPayload, attributes, ...) and the per-instance variables
//     (material textures, lightIntensity, mesh buffers) can be namespace-scope __shared__ objects: the
//     reference declares them with rtDeclareVariable / rtBuffer / rtTextureSampler and reads them as plain
//     names, which is exactly what a __shared__ object gives a one-thread block;
//   * rtTrace = brute-force loop over all triangles, calling the reference's own meshFineIntersect
//     (triangleintersect.cu:17-41) per primitive in ascending global primitive order, then the reference's
//     own rtMaterialClosestHit / rtMaterialAnyHit (ref_device.cu).  Strict "t < current tmax" acceptance
//     (rtPotentialIntersection) makes the smallest primitive id win equal-t ties, the definition the oracle
//     and the product use;
//   * tex2D on a G-buffer sampler reads the addressed texel exactly (the reference samples texel centres,
//     lighttracing.cu:350-351); tex2D on a material sampler is bilinear / repeat with full-float weights
//     (the oracle's stated deviation from the texture unit's 8-bit weights).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>

typedef unsigned int uint;
typedef int rtObject;
enum RTtransformkind { RT_OBJECT_TO_WORLD = 0, RT_WORLD_TO_OBJECT = 1 };

#ifndef M_PIf
#define M_PIf 3.14159265358979323846f
#endif
#define RT_DEFAULT_MAX 1.e27f

#define RT_PROGRAM __device__
#define rtDeclareVariable(type, name, semantic, annotation) __shared__ type name
#define rtBuffer __shared__ ::refshim::Buffer
#define rtTextureSampler __shared__ ::refshim::TexSampler
#define rtPrintf(...) ((void)0)
#define rtPrintExceptionDetails() ((void)0)

namespace optix {
using ::float2; using ::float3; using ::float4; using ::int2; using ::int3; using ::uint2; using ::uint3; using ::uchar4;

// ---- float2 ----
__device__ __forceinline__ float2 make_float2(float s) { return ::make_float2(s, s); }
__device__ __forceinline__ float2 make_float2(const uint2& v) { return ::make_float2((float)v.x, (float)v.y); }
__device__ __forceinline__ float2 operator+(const float2& a, const float2& b) { return ::make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(const float2& a, const float2& b) { return ::make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 operator*(const float2& a, float s) { return ::make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 operator*(float s, const float2& a) { return ::make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 operator/(const float2& a, const float2& b) { return ::make_float2(a.x / b.x, a.y / b.y); }

// ---- float3 ----
__device__ __forceinline__ float3 make_float3(float s) { return ::make_float3(s, s, s); }
__device__ __forceinline__ float3 make_float3(const float4& v) { return ::make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator-(const float3& a) { return ::make_float3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator+(const float3& a, const float3& b) { return ::make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(const float3& a, const float3& b) { return ::make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(const float3& a, const float3& b) { return ::make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(const float3& a, float s) { return ::make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, const float3& a) { return ::make_float3(a.x * s, a.y * s, a.z * s); }
// the SDK divides a vector by a scalar as a multiplication by the reciprocal
__device__ __forceinline__ float3 operator/(const float3& a, float s) { const float inv = 1.0f / s; return a * inv; }
__device__ __forceinline__ float3 operator/(const float3& a, const float3& b) { return ::make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ void operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ void operator-=(float3& a, const float3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
__device__ __forceinline__ void operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
__device__ __forceinline__ void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
__device__ __forceinline__ void operator/=(float3& a, float s) { const float inv = 1.0f / s; a.x *= inv; a.y *= inv; a.z *= inv; }

// ---- float4 ----
__device__ __forceinline__ float4 make_float4(const float3& v) { return ::make_float4(v.x, v.y, v.z, 0.0f); }
__device__ __forceinline__ float4 make_float4(float s) { return ::make_float4(s, s, s, s); }
__device__ __forceinline__ float4 operator+(const float4& a, const float4& b) { return ::make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ void operator+=(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ float4 operator*(const float4& a, float s) { return ::make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 operator*(float s, const float4& a) { return ::make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// ---- scalar / vector helpers (optixu_math_namespace.h) ----
using ::fminf; using ::fmaxf;   // optix::fminf(float, float) is CUDA's own under nvcc
__device__ __forceinline__ float3 fminf(const float3& a, const float3& b) { return ::make_float3(::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z)); }
__device__ __forceinline__ float3 fmaxf(const float3& a, const float3& b) { return ::make_float3(::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z)); }
__device__ __forceinline__ float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(const float3& a, const float3& b) {
    return ::make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length(const float3& v) { return sqrtf(dot(v, v)); }
__device__ __forceinline__ float3 normalize(const float3& v) { const float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
__device__ __forceinline__ float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
__device__ __forceinline__ float3 faceforward(const float3& n, const float3& i, const float3& nref) { return n * copysignf(1.0f, dot(i, nref)); }

__device__ __forceinline__ void cosine_sample_hemisphere(const float u1, const float u2, float3& p) {
    const float r = sqrtf(u1);
    const float phi = 2.0f * M_PIf * u2;
    p.x = r * cosf(phi);
    p.y = r * sinf(phi);
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}

struct Onb {
    __device__ __forceinline__ Onb(const float3& normal) {
        m_normal = normal;
        if (fabs(m_normal.x) > fabs(m_normal.z)) {
            m_binormal.x = -m_normal.y; m_binormal.y = m_normal.x; m_binormal.z = 0;
        } else {
            m_binormal.x = 0; m_binormal.y = -m_normal.z; m_binormal.z = m_normal.y;
        }
        m_binormal = normalize(m_binormal);
        m_tangent = cross(m_binormal, m_normal);
    }
    __device__ __forceinline__ void inverse_transform(float3& p) const { p = p.x * m_tangent + p.y * m_binormal + p.z * m_normal; }
    float3 m_tangent, m_binormal, m_normal;
};

struct Ray {
    __device__ __forceinline__ Ray() {}
    __device__ __forceinline__ Ray(float3 o, float3 d, unsigned int type, float tmin_, float tmax_ = RT_DEFAULT_MAX)
        : origin(o), direction(d), ray_type(type), tmin(tmin_), tmax(tmax_) {}
    float3 origin, direction;
    unsigned int ray_type;
    float tmin, tmax;
};

__device__ __forceinline__ bool intersect_triangle_branchless(const Ray& ray, const float3& p0, const float3& p1, const float3& p2,
                                                              float3& n, float& t, float& beta, float& gamma) {
    const float3 e0 = p1 - p0;
    const float3 e1 = p0 - p2;
    n = cross(e1, e0);
    const float3 e2 = (1.0f / dot(n, ray.direction)) * (p0 - ray.origin);
    const float3 i = cross(ray.direction, e2);
    beta = dot(i, e1);
    gamma = dot(i, e0);
    t = dot(n, e2);
    return ((t < ray.tmax) & (t > ray.tmin) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1));
}
// (meshIntersect, the early-exit variant, is never bound by the reference host code: rtcomphoton.h:435-439)
__device__ __forceinline__ bool intersect_triangle_earlyexit(const Ray& ray, const float3& p0, const float3& p1, const float3& p2,
                                                             float3& n, float& t, float& beta, float& gamma) {
    return intersect_triangle_branchless(ray, p0, p1, p2, n, t, beta, gamma);
}

struct Aabb {
    float3 m_min, m_max;
    __device__ __forceinline__ void invalidate() { m_min = make_float3(1e37f); m_max = make_float3(-1e37f); }
};
}  // namespace optix

namespace refshim {

template <typename T, int DIM = 1>
struct Buffer {
    T* data;
    size_t count;   // DIM == 1: elements; DIM == 2: width (elements per row)
    __device__ __forceinline__ size_t size() const { return count; }
    __device__ __forceinline__ T& operator[](size_t i) { return data[i]; }
    __device__ __forceinline__ const T& operator[](size_t i) const { return data[i]; }
    __device__ __forceinline__ T& operator[](const uint2& i) { return data[(size_t)i.y * count + i.x]; }
};

// kind 0: exact texel at the addressed position (G-buffer planes); kind 1: bilinear, repeat, full-float weights
template <typename T, int DIM>
struct TexSampler {
    const float4* texels;
    int w, h, kind;
};

template <typename T, int DIM>
__device__ inline float4 tex2D(const TexSampler<T, DIM>& t, float u, float v) {
    if (t.kind == 0) {
        int x = (int)floorf(u * (float)t.w), y = (int)floorf(v * (float)t.h);
        x = x < 0 ? 0 : (x >= t.w ? t.w - 1 : x);
        y = y < 0 ? 0 : (y >= t.h ? t.h - 1 : y);
        return t.texels[(size_t)y * t.w + x];
    }
    if (t.w == 1 && t.h == 1) return t.texels[0];
    const float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    int i0 = (int)fx % t.w; if (i0 < 0) i0 += t.w;
    int j0 = (int)fy % t.h; if (j0 < 0) j0 += t.h;
    int i1 = i0 + 1; if (i1 == t.w) i1 = 0;
    int j1 = j0 + 1; if (j1 == t.h) j1 = 0;
    const float4 t00 = t.texels[j0 * t.w + i0], t10 = t.texels[j0 * t.w + i1];
    const float4 t01 = t.texels[j1 * t.w + i0], t11 = t.texels[j1 * t.w + i1];
    float4 r;
    float lo, hi;
    lo = t00.x + a * (t10.x - t00.x); hi = t01.x + a * (t11.x - t01.x); r.x = lo + b * (hi - lo);
    lo = t00.y + a * (t10.y - t00.y); hi = t01.y + a * (t11.y - t01.y); r.y = lo + b * (hi - lo);
    lo = t00.z + a * (t10.z - t00.z); hi = t01.z + a * (t11.z - t01.z); r.z = lo + b * (hi - lo);
    lo = t00.w + a * (t10.w - t00.w); hi = t01.w + a * (t11.w - t01.w); r.w = lo + b * (hi - lo);
    return r;
}

// ---- the traversal state of the one ray this thread is tracing (one thread per block) ----
struct TraceState {
    float tmin, tmax;      // current acceptance interval (tmax shrinks as hits are committed)
    float tCandidate;      // t passed to rtPotentialIntersection
    int hit;               // a hit was committed
    int terminate;         // rtTerminateRay() was called by the any-hit program
    int anyHitRay;         // ray type 1: run the any-hit program on every reported intersection
    int curMesh, curPrim;  // primitive being intersected
    int hitMesh, hitPrim;
    float3 hitNormal;      // committed attributes
    float2 hitTexcoord;
};
__shared__ TraceState g_trace;

}  // namespace refshim

using refshim::tex2D;

__device__ __forceinline__ float3 rtTransformNormal(RTtransformkind, const float3& n) { return n; }  // no transforms in the scene graph
__device__ __forceinline__ bool rtPotentialIntersection(float t) {
    if (t > refshim::g_trace.tmin && t < refshim::g_trace.tmax) { refshim::g_trace.tCandidate = t; return true; }
    return false;
}
__device__ void rtReportIntersection(unsigned int material);  // ref_device.cu (needs the programs' attribute variables)
__device__ __forceinline__ void rtTerminateRay() { refshim::g_trace.terminate = 1; }

__device__ void refshim_trace(const optix::Ray& ray, void* prd);  // ref_device.cu
template <class PRD>
__device__ __forceinline__ void rtTrace(rtObject, const optix::Ray& ray, PRD& prd) { refshim_trace(ray, &prd); }
