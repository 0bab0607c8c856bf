// shading.h -- BRDF / light / texture arithmetic of the EVPLP hot path for the sm_100a
// kernels (host+device inline so that a host build can unit-test it).
//
// Every function names the reference program it replaces:
//   rtmaterial.cuh:25-155        Lambert/Phong Sample, EvalF, PdfW, PdfA, MaxColor
//   rtlightsource.cuh:24-80      LightSample (CDF lower-bound + SquareToBarycentric rtmath.cuh:23-28)
//   lighttracing.cu:254-346      MIS heuristics and vplSplat's shading tail
//   photonsplatinstanced.frag    GLSL BRDF variants (thresholds 1e-5, two-sided Lambert)
// All arithmetic is written with the fixed evaluation order of vec.h / detmath.h and the
// translation units are compiled with -fmad=false, so results are bit-identical to the
// scalar CPU oracle used by the tests.
#pragma once
#include "vec.h"
#include "xorwow.h"

namespace evplp {

// ------------------------------------------------------------------ textures ------------
struct DevTexture {
    int w, h;
    int offset;  // first texel (float4 index) inside the texture pool
    int pad;
};

struct DevMaterial {
    DevTexture lambert, phong, exponent;
    float lightIntensity[4];
};

struct Texel4 {
    float x, y, z, w;
};

// Bilinear, repeat-wrapped, full-float weights (RT_WRAP_REPEAT + RT_FILTER_LINEAR,
// rtcommon.h:225-244; GL_LINEAR/GL_REPEAT 203-208).  1x1 textures (constant colours,
// rtcommon.h:80-90) short-circuit.
template <typename F4>
EVPLP_HD Texel4 tex_fetch(const DevTexture& t, const F4* __restrict__ pool, float u, float v) {
    Texel4 r;
    const F4* base = pool + t.offset;
    if (t.w == 1 && t.h == 1) {
        F4 c = base[0];
        r.x = c.x; r.y = c.y; r.z = c.z; r.w = c.w;
        return r;
    }
    float x = u * (float)t.w - 0.5f;
    float y = v * (float)t.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx % t.w; if (i0 < 0) i0 += t.w;
    int j0 = (int)fy % t.h; if (j0 < 0) j0 += t.h;
    int i1 = i0 + 1; if (i1 == t.w) i1 = 0;
    int j1 = j0 + 1; if (j1 == t.h) j1 = 0;
    F4 t00 = base[j0 * t.w + i0];
    F4 t10 = base[j0 * t.w + i1];
    F4 t01 = base[j1 * t.w + i0];
    F4 t11 = base[j1 * t.w + i1];
    float lo, hi;
    lo = t00.x + a * (t10.x - t00.x); hi = t01.x + a * (t11.x - t01.x); r.x = lo + b * (hi - lo);
    lo = t00.y + a * (t10.y - t00.y); hi = t01.y + a * (t11.y - t01.y); r.y = lo + b * (hi - lo);
    lo = t00.z + a * (t10.z - t00.z); hi = t01.z + a * (t11.z - t01.z); r.z = lo + b * (hi - lo);
    lo = t00.w + a * (t10.w - t00.w); hi = t01.w + a * (t11.w - t01.w); r.w = lo + b * (hi - lo);
    return r;
}

// ------------------------------------------------------------------ BRDF library --------
EVPLP_HD float lambert_pdf_w(V3 n1, V3 v12) {  // rtmaterial.cuh:40-44 (no 1/pi: kept)
    return det_max(dot(n1, normalize(v12)), 0.f);
}

EVPLP_HD float lambert_pdf_a(V3 n1, V3 n2, V3 v12) {  // :46-54
    const float c1 = det_max(dot(n1, v12), 0.f);
    const float c2 = det_max(-dot(n2, v12), 0.f);
    const float d2 = dot(v12, v12);
    return det_div(c1 * c2, d2 * d2) * kInvPi;
}

EVPLP_HD void cosine_sample_hemisphere(float u1, float u2, V3& p) {  // optixu_math_namespace.h
    const float r = det_sqrtf(u1);
    const float phi = 2.0f * kPi * u2;
    float s, c;
    det_sincosf(phi, &s, &c);
    p.x = r * c;
    p.y = r * s;
    p.z = det_sqrtf(det_max(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}

// :56-67; the two uniforms are drawn left to right (definition)
EVPLP_HD V3 lambert_sample(V3* out, float* pdfW, V3 normal, V3 lambertReflectance, Xorwow& rng) {
    float u1 = xorwow_uniform(rng);
    float u2 = xorwow_uniform(rng);
    cosine_sample_hemisphere(u1, u2, *out);
    Onb onb = make_onb(normal);
    *out = onb_inverse_transform(onb, *out);
    *pdfW = det_max(dot(*out, normal), 0.f) * kInvPi;
    return lambertReflectance;
}

EVPLP_HD float phong_pdf_w(V3 n1, V3 v12, V3 in, V3 phongReflectance, float phongExponent) {  // :79-86
    V3 wi12 = normalize(v12);
    V3 reflectVec = normalize(reflect(-in, n1));
    float cosReflect = det_max(dot(wi12, reflectVec), 0.f);
    if (cosReflect <= 0.000001f || phongReflectance.x <= 0.000001f) return 0.0f;
    return (phongExponent + 1.0f) * 0.5f * kInvPi * det_powf(cosReflect, phongExponent);
}

EVPLP_HD float phong_pdf_a(V3 n1, V3 n2, V3 v12, V3 in, V3 phongReflectance, float phongExponent) {  // :88-103
    V3 wi12 = normalize(v12);
    V3 reflectVec = normalize(reflect(-in, n1));
    float cosReflect = det_max(dot(wi12, reflectVec), 0.f);
    if (cosReflect <= 0.000001f || phongReflectance.x <= 0.000001f) return 0.0f;
    float pdfW = (phongExponent + 1.0f) * 0.5f * kInvPi * det_powf(cosReflect, phongExponent);
    float cos2 = det_max(-dot(n2, wi12), 0.0f);
    float dist2 = dot(v12, v12);
    return det_div(pdfW * cos2, dist2);
}

EVPLP_HD float phong_eval_f(V3 out, V3 in, V3 normal, float phongExponent) {  // :113-119
    V3 reflectVec = reflect(-in, normal);
    float dotWrWo = det_max(dot(out, reflectVec), 0.0f);
    if (dotWrWo <= 0.000001f) return 0.0f;
    return (phongExponent + 2.0f) * det_powf(dotWrWo, phongExponent) * (kInvPi) * 0.5f;
}

EVPLP_HD V3 phong_sample(V3* out, float* pdfW, V3 in, V3 normal, V3 phongReflectance, float phongExponent,
                         Xorwow& rng) {  // :121-155
    V3 reflectVec = reflect(-in, normal);
    float sampleX = xorwow_uniform(rng);
    float sampleY = xorwow_uniform(rng);
    float cosTheta = det_powf(sampleX, det_div(1.f, phongExponent + 1.f));
    float sinTheta = det_sqrtf(1.0f - cosTheta * cosTheta);
    float phi = 2.f * kPi * sampleY;
    float sinPhi, cosPhi;
    det_sincosf(phi, &sinPhi, &cosPhi);
    *out = v3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
    Onb onb = make_onb(reflectVec);
    *out = onb_inverse_transform(onb, *out);
    float unsafeCosNormal = dot(*out, normal);
    float cosNormal = det_max(unsafeCosNormal, 0.f);
    float cosReflect = det_max(dot(*out, reflectVec), 0.f);
    if (unsafeCosNormal > 0.0f) {
        *pdfW = (phongExponent + 1.0f) * 0.5f * det_powf(cosReflect, phongExponent) * kInvPi;
    } else {
        *pdfW = 0.0f;
    }
    return det_div(phongExponent + 2.0f, phongExponent + 1.0f) * cosNormal * phongReflectance;
}

// rtmaterial.cuh:104-111 (the CUDA PhongEval: threshold 1e-6 and the Ks.x test; the GLSL one differs, see below)
EVPLP_HD V3 phong_eval(V3 out, V3 in, V3 normal, V3 ks, float e) {
    V3 reflectVec = reflect(-in, normal);
    float dotWrWo = det_max(dot(out, reflectVec), 0.0f);
    if (dotWrWo <= 0.000001f || ks.x <= 0.000001f) return v3s(0.0f);
    return ks * (e + 2.0f) * det_powf(dotWrWo, e) * (kInvPi) * 0.5f;
}
// rtmaterial.cuh:30-38
EVPLP_HD float geometry_term(V3 n1, V3 n2, V3 v12) {
    const float c1 = det_max(dot(n1, v12), 0.f);
    const float c2 = det_max(-dot(n2, v12), 0.f);
    const float d2 = dot(v12, v12);
    return det_div(c1 * c2, d2 * d2);
}
// pathtracing.cu:85-95
EVPLP_HD float pdf_w2a(V3 n2, V3 v12) {
    V3 nv12 = normalize(v12);
    return det_div(det_max(-dot(n2, nv12), 0.f), dot(v12, v12));
}
// pathtracing.cu:49-52 (quirk kept: the probability is never below 0.98 and may exceed 1)
EVPLP_HD float pt_russian_prob(V3 throughput) {
    return det_max(det_max(throughput.x, 0.98f), det_max(throughput.y, throughput.z));
}

// rtmath.cuh:23-28
EVPLP_HD void square_to_barycentric(float* beta, float* gamma, float x, float y) {
    const float sqrtX = det_sqrtf(x);
    *beta = sqrtX * (1.0f - y);
    *gamma = sqrtX * y;
}

// lighttracing.cu:254-272
EVPLP_HD float balance_heuristic(float a, float b) { return det_div(a, a + b); }
EVPLP_HD float max_heuristic(float a, float b) { return a > b ? 1.f : 0.f; }
EVPLP_HD float power_heuristic2(float a, float b) { return balance_heuristic(a * a, b * b); }

// lighttracing.cu:93-96
EVPLP_HD float russian_prob(V3 throughput) {
    return det_min(det_max(throughput.x, det_max(throughput.y, throughput.z)), 0.98f);
}

// A light-path vertex unpacked for shading (RtPhotonRecord fields, rtphotonrecord.h:17-25).
struct Vertex {
    V3 pos, normal, flux, fluxDir, kd, ks;
    float exponent, pSel;
};

// A G-buffer texel (deferred.frag outputs).
struct Surface {
    V3 pos, normal, kd, ks;
    float exponent;
};

// Terms of the shading tail that depend on the VPL alone.  The gather kernel evaluates them once per VPL when it
// stages a batch (the same expressions on the same inputs give the same bits wherever they are evaluated).
struct VplPre {
    V3 refl;    // reflect(-fluxDir, normal)               (PhongEval, rtmaterial.cuh:113-119)
    V3 reflN;   // normalize(refl)                         (PhongPdfA, rtmaterial.cuh:88-103)
    V3 kdPi;    // kInvPi * kd
};
EVPLP_HD VplPre vpl_precompute(const Vertex& vp) {
    VplPre p;
    p.refl = reflect(-vp.fluxDir, vp.normal);
    p.reflN = normalize(p.refl);
    p.kdPi = kInvPi * vp.kd;
    return p;
}
// phong_eval_f / phong_pdf_a with the reflected direction supplied
EVPLP_HD float phong_eval_f_refl(V3 out, V3 reflectVec, float phongExponent) {
    float dotWrWo = det_max(dot(out, reflectVec), 0.0f);
    if (dotWrWo <= 0.000001f) return 0.0f;
    return (phongExponent + 2.0f) * det_powf(dotWrWo, phongExponent) * (kInvPi) * 0.5f;
}
EVPLP_HD float phong_pdf_a_refl(V3 n2, V3 v12, V3 reflectVecN, V3 phongReflectance, float phongExponent) {
    V3 wi12 = normalize(v12);
    float cosReflect = det_max(dot(wi12, reflectVecN), 0.f);
    if (cosReflect <= 0.000001f || phongReflectance.x <= 0.000001f) return 0.0f;
    float pdfW = (phongExponent + 1.0f) * 0.5f * kInvPi * det_powf(cosReflect, phongExponent);
    float cos2 = det_max(-dot(n2, wi12), 0.0f);
    float dist2 = dot(v12, v12);
    return det_div(pdfW * cos2, dist2);
}

// Shading tail of vplSplat (lighttracing.cu:296-345) after the cosine test and the shadow
// ray; c1c2 = unnormCos1 * unnormCos2, v12 = vpl.pos - x.
EVPLP_HD V3 vpl_shade_pre(const Surface& sf, V3 wi10, const Vertex& vp, const VplPre& pre, V3 v12, float c1c2, unsigned misMode,
                          float pdfMc, float clampingValue) {
    float dist2 = dot(v12, v12);
    float dist = det_sqrtf(dist2);
    V3 wi12 = v12 / dist;
    V3 brdf2 = pre.kdPi + phong_eval_f_refl(-wi12, pre.refl, vp.exponent) * vp.ks;
    V3 brdf1 = kInvPi * sf.kd + phong_eval_f(wi10, wi12, sf.normal, sf.exponent) * sf.ks;
    float g21 = det_div(c1c2, dist2 * dist2);
    if (misMode == 0) {
        return vp.flux * brdf1 * brdf2 * g21;
    } else if (misMode <= 3) {
        float pdfDe = lambert_pdf_a(vp.normal, sf.normal, -v12) * vp.pSel;
        pdfDe += phong_pdf_a_refl(sf.normal, -v12, pre.reflN, vp.ks, vp.exponent) * (1.0f - vp.pSel);
        float weight = misMode == 1 ? balance_heuristic(pdfMc, pdfDe)
                     : misMode == 2 ? max_heuristic(pdfMc, pdfDe) : power_heuristic2(pdfMc, pdfDe);
        return weight * vp.flux * brdf1 * brdf2 * g21;
    } else if (misMode == 4) {
        return vp.flux * det_min(g21, clampingValue) * brdf1 * brdf2;
    } else {
        V3 gb = g21 * brdf1 * brdf2;
        return vp.flux * vmin(gb, v3s(clampingValue));
    }
}
EVPLP_HD V3 vpl_shade(const Surface& sf, V3 wi10, const Vertex& vp, V3 v12, float c1c2, unsigned misMode,
                      float pdfMc, float clampingValue) {
    return vpl_shade_pre(sf, wi10, vp, vpl_precompute(vp), v12, c1c2, misMode, pdfMc, clampingValue);
}

// ------------------------------------------------------------------ GLSL variants -------
// photonsplatinstanced.frag:36-79
EVPLP_HD V3 glsl_lambert_eval(V3 w10, V3 w12, V3 normal, V3 kd) {
    if (dot(w10, normal) <= 0.0f || dot(w12, normal) <= 0.0f) return v3s(0.0f);
    return kInvPi * kd;
}
EVPLP_HD V3 glsl_phong_eval(V3 outVec, V3 inVec, V3 normal, V3 ks, float e) {
    V3 reflectVec = reflect(-inVec, normal);
    float dotWrWo = dot(outVec, reflectVec);
    if (dotWrWo <= 0.00001f) return v3s(0.0f);
    return ks * (e + 2.0f) * det_powf(dotWrWo, e) * kInvPi * 0.5f;
}
EVPLP_HD float glsl_lambert_pdf_w(V3 n1, V3 v12) { return det_max(dot(n1, normalize(v12)), 0.f) * kInvPi; }
EVPLP_HD float glsl_phong_pdf_w(V3 n1, V3 wi12, V3 inVec, V3 ks, float e) {
    V3 reflectVec = reflect(-inVec, n1);
    float dotWrWo = det_max(dot(wi12, reflectVec), 0.f);
    if (dotWrWo <= 0.00001f || ks.x <= 0.00001f) return 0.0f;
    return (e + 1.0f) * 0.5f * kInvPi * det_powf(dotWrWo, e);
}

struct SplatUniforms {
    V3 cameraPosition;
    float radius;
    float pdfMc;
    float clampingValue;
    unsigned misMode;
    unsigned numLightPaths;
};

// The fragment shader (photonsplatinstanced.frag:146-240) split into its per-photon part
// (everything that does not depend on the shaded texel) and its per-fragment part, with the
// arithmetic of each expression unchanged.
struct SplatPhoton {
    V3 pos, w12, flux;   // photon position, direction to its predecessor, flux
    float weight;        // modes 1-3: MIS weight h(mixPdfA, pdfMc)
    float dist2;         // |p_{k-1} - p_k|^2 (modes 4, 5)
    V3 prevN;            // predecessor normal (modes 4, 5)
    V3 brdf2;            // BRDF at the predecessor (mode 5)
    int live;            // mixPdfW > 0
};

EVPLP_HD SplatPhoton splat_prepare(const SplatUniforms& U, const Vertex& ph, const Vertex& prev) {
    SplatPhoton sp;
    sp.pos = ph.pos; sp.flux = ph.flux;
    V3 v12 = prev.pos - ph.pos;
    V3 w12 = normalize(v12);
    sp.w12 = w12;
    float mixPdfW = glsl_lambert_pdf_w(prev.normal, -w12) * prev.pSel;
    mixPdfW += glsl_phong_pdf_w(prev.normal, -w12, prev.fluxDir, prev.ks, prev.exponent) * (1.0f - prev.pSel);
    sp.live = (mixPdfW > 0.0f) ? 1 : 0;
    sp.weight = 1.0f; sp.dist2 = dot(v12, v12); sp.prevN = prev.normal; sp.brdf2 = v3s(0.0f);
    if (U.misMode >= 1 && U.misMode <= 3) {
        float mixPdfA = det_div(mixPdfW * det_max(dot(ph.normal, w12), 0.0f), dot(v12, v12));
        sp.weight = U.misMode == 1 ? balance_heuristic(mixPdfA, U.pdfMc)
                  : U.misMode == 2 ? max_heuristic(mixPdfA, U.pdfMc) : power_heuristic2(mixPdfA, U.pdfMc);
    } else if (U.misMode == 5) {
        sp.brdf2 = glsl_lambert_eval(-w12, prev.fluxDir, prev.normal, prev.kd) +
                   glsl_phong_eval(-w12, prev.fluxDir, prev.normal, prev.ks, prev.exponent);
    }
    return sp;
}

// w10 = normalize(cameraPosition - sf.pos); invR2 = 1 / r^2 (rtcomphoton.h:819); invN = 1 / numLightPaths (:820).
// The radius test is done by the caller.  Returns false when the fragment is discarded.
EVPLP_HD bool splat_shade(const SplatUniforms& U, float invR2, float invN, const Surface& sf, V3 w10, const SplatPhoton& sp, V3* color) {
    V3 brdf1 = glsl_lambert_eval(w10, sp.w12, sf.normal, sf.kd) + glsl_phong_eval(w10, sp.w12, sf.normal, sf.ks, sf.exponent);
    if (!sp.live) {
        *color = v3s(0.0f);
        return true;
    }
    if (U.misMode == 0) {
        *color = brdf1 * (kInvPi * invR2) * sp.flux * invN;
    } else if (U.misMode <= 3) {
        *color = brdf1 * (kInvPi * invR2) * sp.flux * invN * sp.weight;
    } else {
        float cosCos = det_max(dot(sf.normal, sp.w12), 0.0f) * det_max(-dot(sp.prevN, sp.w12), 0.0f);
        if (cosCos <= 0.0f) return false;
        float geometryTerm = det_div(cosCos, sp.dist2);
        if (U.misMode == 4) {
            *color = brdf1 * (kInvPi * invR2) * sp.flux * invN * det_max(geometryTerm - U.clampingValue, 0.0f) / geometryTerm;
        } else {
            V3 num = vmax((brdf1 * sp.brdf2 * geometryTerm) - v3s(U.clampingValue), v3s(0.0f));
            V3 den = geometryTerm * sp.brdf2;
            V3 pre = (kInvPi * invR2) * sp.flux * invN;
            *color = v3(det_div(pre.x * num.x, den.x), det_div(pre.y * num.y, den.y), det_div(pre.z * num.z, den.z));
        }
    }
    return true;
}

EVPLP_HD bool splat_fragment(const SplatUniforms& U, const Surface& sf, const Vertex& ph, const Vertex& prev, V3* color) {
    const SplatPhoton sp = splat_prepare(U, ph, prev);
    const float invR2 = det_div(1.0f, U.radius * U.radius);
    const float invN = det_div(1.0f, (float)U.numLightPaths);
    return splat_shade(U, invR2, invN, sf, normalize(U.cameraPosition - sf.pos), sp, color);
}

// Q31.32 fixed point of the accumulation layers; non-finite / huge values are dropped.
EVPLP_HD long long to_fixed(float c) {
    if (!(fabsf(c) < 1.0e9f)) return 0;
#if defined(__CUDA_ARCH__)
    return __float2ll_rn(c * 4294967296.0f);
#else
    return (long long)llrintf(c * 4294967296.0f);
#endif
}

}  // namespace evplp
