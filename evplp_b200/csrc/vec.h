// vec.h -- float3 helpers with a FIXED evaluation order (no contraction).
// The forms restate the OptiX SDK 4.1.1 helpers the reference calls
// (optixu_math_namespace.h): normalize = v * (1/sqrt(dot)),
// reflect(i,n) = i - 2*n*dot(n,i), faceforward(n,i,nref) = n*copysign(1,dot(i,nref)).
#pragma once
#include "detmath.h"

namespace evplp {

struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};

EVPLP_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
EVPLP_HD V3 v3s(float s) { return v3(s, s, s); }
EVPLP_HD V3 v3p(const float* p) { return v3(p[0], p[1], p[2]); }
EVPLP_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
EVPLP_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
EVPLP_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
EVPLP_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
EVPLP_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
EVPLP_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
// OptiX float3/float: multiplies by the reciprocal (optixu_math_namespace.h operator/).
EVPLP_HD V3 operator/(V3 a, float s) { float inv = det_div(1.0f, s); return v3(a.x * inv, a.y * inv, a.z * inv); }
EVPLP_HD V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
EVPLP_HD V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
EVPLP_HD V3& operator/=(V3& a, float s) { a = a / s; return a; }

EVPLP_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
EVPLP_HD V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
EVPLP_HD V3 normalize(V3 v) { float inv = det_div(1.0f, det_sqrtf(dot(v, v))); return v * inv; }
EVPLP_HD V3 reflect(V3 i, V3 n) { return i - 2.0f * n * dot(n, i); }
EVPLP_HD V3 faceforward(V3 n, V3 i, V3 nref) { return n * copysignf(1.0f, dot(i, nref)); }
EVPLP_HD V3 vmin(V3 a, V3 b) { return v3(det_min(a.x, b.x), det_min(a.y, b.y), det_min(a.z, b.z)); }
EVPLP_HD V3 vmax(V3 a, V3 b) { return v3(det_max(a.x, b.x), det_max(a.y, b.y), det_max(a.z, b.z)); }
EVPLP_HD float max_color(V3 c) { return det_max(det_max(c.x, c.y), c.z); }  // rtmaterial.cuh:25-28

// OptiX Onb (optixu_math_namespace.h): branch on |n.x| > |n.z|.
struct Onb {
    V3 tangent, binormal, normal;
};
EVPLP_HD Onb make_onb(V3 n) {
    Onb o;
    o.normal = n;
    if (fabsf(n.x) > fabsf(n.z)) {
        o.binormal = v3(-n.y, n.x, 0.0f);
    } else {
        o.binormal = v3(0.0f, -n.z, n.y);
    }
    o.binormal = normalize(o.binormal);
    o.tangent = cross(o.binormal, o.normal);
    return o;
}
EVPLP_HD V3 onb_inverse_transform(const Onb& o, V3 p) {
    return p.x * o.tangent + p.y * o.binormal + p.z * o.normal;
}

}  // namespace evplp
