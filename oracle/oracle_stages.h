// oracle_stages.h -- TEST INFRASTRUCTURE (CPU oracle): scalar restatement, statement by
// statement, of the reference's device programs for the EVPLP hot path.
//   BRDF library ......... realtimetechniques/rtmaterial.cuh:25-155
//   LightSample .......... realtimetechniques/rtlightsource.cuh:24-80, rtmath.cuh:23-28
//   tracePhotons ......... realtimetechniques/lighttracing.cu:192-250
//   rtMaterialClosestHit . realtimetechniques/lighttracing.cu:113-182
//   vplSplat/splatColor .. realtimetechniques/lighttracing.cu:254-379
//   VSL .................. realtimetechniques/lighttracing.cu:382-722
//   LVC splatColor ....... realtimetechniques/lvclighttracing.cu:348-387
//   photon splat ......... shaders/photonsplatinstanced.{vert,geom,frag}
//   G-buffer ............. shaders/deferred.{geom,frag} (ray-cast definition)
// Parity status: pinned by the reference's own device code for the CUDA programs, see oracle_math.h.
#pragma once
#include "oracle_scene.h"

namespace orc {

// ------------------------------- rtmaterial.cuh -------------------------------------
inline float MaxColor(const F3& color) { return fmaxf(fmaxf(color.x, color.y), color.z); }

inline float LambertPdfW(const F3& n1, const F3& v12) {  // :40-44 (no 1/pi: quirk kept)
    const float cos1Unnorm = fmaxf(dot(n1, normalize(v12)), 0.f);
    return cos1Unnorm;
}

inline float LambertPdfA(const F3& n1, const F3& n2, const F3& v12) {  // :46-54
    const float cos1Unnorm = fmaxf(dot(n1, v12), 0.f);
    const float cos2Unnorm = fmaxf(-dot(n2, v12), 0.f);
    const float d2 = dot(v12, v12);
    return cos1Unnorm * cos2Unnorm / (d2 * d2) * M_Inv_PIf;
}

// Argument evaluation order of the two curand_uniform() calls at :58 is unspecified in C++;
// the oracle DEFINES left-to-right.
inline F3 LambertSample(F3* out, float* pdfW, const F3& in, const F3& normal, const F3& lambertReflectance,
                        CurandState* rngState) {  // :56-67
    (void)in;
    float u1 = curand_uniform(rngState);
    float u2 = curand_uniform(rngState);
    cosine_sample_hemisphere(u1, u2, *out);
    Onb onb(normal);
    onb.inverse_transform(*out);
    *pdfW = fmaxf(dot(*out, normal), 0.f) * M_Inv_PIf;
    return lambertReflectance;
}

inline float LambertEvalF(const F3&, const F3&, const F3&) { return M_Inv_PIf; }  // :74-77

inline float PhongPdfW(const F3& n1, const F3& v12, const F3& in, const F3& phongReflectance,
                       const float phongExponent) {  // :79-86
    F3 wi12 = normalize(v12);
    F3 reflectVec = normalize(reflect(-in, n1));
    float cosReflect = fmaxf(dot(wi12, reflectVec), 0.f);
    if (cosReflect <= 0.000001f || phongReflectance.x <= 0.000001f) { return 0.0f; }
    return (phongExponent + 1.0f) * 0.5f * M_Inv_PIf * det_powf(cosReflect, phongExponent);
}

inline float PhongPdfA(const F3& n1, const F3& n2, const F3& v12, const F3& in, const F3& phongReflectance,
                       const float phongExponent) {  // :88-103
    (void)n2;
    F3 wi12 = normalize(v12);
    F3 reflectVec = normalize(reflect(-in, n1));
    float cosReflect = fmaxf(dot(wi12, reflectVec), 0.f);
    if (cosReflect <= 0.000001f || phongReflectance.x <= 0.000001f) { return 0.0f; }
    float pdfW = (phongExponent + 1.0f) * 0.5f * M_Inv_PIf * det_powf(cosReflect, phongExponent);
    float cos2 = fmaxf(-dot(n2, wi12), 0.0f);
    float dist2 = dot(v12, v12);
    return pdfW * cos2 / dist2;
}

inline float PhongEvalF(const F3& out, const F3& in, const F3& normal, const float phongExponent) {  // :113-119
    F3 reflectVec = reflect(-in, normal);
    float dotWrWo = fmaxf(dot(out, reflectVec), 0.0f);
    if (dotWrWo <= 0.000001f) { return 0.0f; }
    return (phongExponent + 2.0f) * det_powf(dotWrWo, phongExponent) * (M_Inv_PIf) * 0.5f;
}

inline F3 PhongSample(F3* out, float* pdfW, const F3& in, const F3& normal, const F3& phongReflectance,
                      const float phongExponent, CurandState* rngState) {  // :121-155
    F3 reflectVec = reflect(-in, normal);
    float sampleX = curand_uniform(rngState);
    float sampleY = curand_uniform(rngState);
    float cosTheta = det_powf(sampleX, 1.f / (phongExponent + 1.f));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float phi = 2.f * M_PIf_ * sampleY;
    float cosPhi = det_cosf(phi);
    float sinPhi = det_sinf(phi);
    *out = mk3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
    Onb onb(reflectVec);
    onb.inverse_transform(*out);
    float unsafeCosNormal = dot(*out, normal);
    float cosNormal = fmaxf(unsafeCosNormal, 0.f);
    float cosReflect = fmaxf(dot(*out, reflectVec), 0.f);
    if (unsafeCosNormal > 0.0f) {
        *pdfW = (phongExponent + 1.0f) * 0.5f * det_powf(cosReflect, phongExponent) * M_Inv_PIf;
    } else {
        *pdfW = 0.0f;
    }
    F3 result = (phongExponent + 2.0f) / (phongExponent + 1.0f) * cosNormal * phongReflectance;
    return result;
}

// ------------------------------- rtlightsource.cuh ----------------------------------
inline void SquareToBarycentric(float* beta, float* gamma, const float x, const float y) {  // rtmath.cuh:23-28
    const float sqrtX = sqrtf(x);
    *beta = (sqrtX * (1.0f - y));
    *gamma = (sqrtX * y);
}

inline F3 LightSample(const Scene& s, F3* position, F3* normal, float* pdf, CurandState* state) {  // :24-80
    float randNum = curand_uniform(state);
    unsigned int count = (unsigned int)s.lightCdf.size();
    unsigned int step = 0;
    unsigned int first = 0;
    while (count > 0) {
        unsigned int it = first;
        step = count / 2;
        it += step;
        if (s.lightCdf[it] < randNum) {
            first = ++it;
            count -= step + 1;
        } else {
            count = step;
        }
    }
    unsigned int indicesIndex = first;
    const Tri& tri = s.tris[s.lightFirst + indicesIndex];
    float beta, gamma;
    // argument evaluation order at :62 unspecified; oracle defines left-to-right
    float bx = curand_uniform(state);
    float by = curand_uniform(state);
    SquareToBarycentric(&beta, &gamma, bx, by);
    const F3& pos1 = tri.p0;
    const F3& pos2 = tri.p1;
    const F3& pos3 = tri.p2;
    *position = pos1 * beta + pos2 * gamma + pos3 * (1.0f - gamma - beta);
    *normal = normalize(cross(pos2 - pos1, pos3 - pos1));
    *pdf = 1.f / s.lightArea;
    const float invPdf = s.lightArea;
    return mk3(s.lightIntensity[0], s.lightIntensity[1], s.lightIntensity[2]) * invPdf;
}

// ------------------------------- lighttracing.cu: light tracing ---------------------
inline float russianProb(const F3& throughput) {  // :93-96
    return fminf(fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)), 0.98f);
}

inline void setv(float* d, const F3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
inline F3 getv(const float* d) { return mk3(d[0], d[1], d[2]); }

struct PerRayData_radiance {
    bool done;
    CurandState* rngState;
    F3 nextPosition, nextDirection, flux;
    unsigned int photonIndex;
    int flag;
};

// rtMaterialClosestHit (:113-182); `photons` is the record window of this trace call.
inline void rtMaterialClosestHit(const Scene& s, EvplpRecord* photons, PerRayData_radiance& prdRadiance,
                                 const F3& rayOrigin, const F3& rayDirection, const Hit& hit) {
    const Tri& tri = s.tris[hit.prim];
    const Material& mat = s.mats[tri.mat];
    // attributes written by meshFineIntersect (triangleintersect.cu:31-36)
    F3 geometryNormal = normalize(hit.n);
    F2 texcoord;
    {
        float w0 = 1.0f - hit.beta - hit.gamma;
        texcoord.x = tri.t1.x * hit.beta + tri.t2.x * hit.gamma + tri.t0.x * w0;
        texcoord.y = tri.t1.y * hit.beta + tri.t2.y * hit.gamma + tri.t0.y * w0;
    }
    float tHit = hit.t;

    F3 worldGeometryNormal = normalize(geometryNormal);  // rtTransformNormal is the identity
    F3 ffNormal = faceforward(worldGeometryNormal, -rayDirection, worldGeometryNormal);

    F3 nextPosition = rayOrigin + tHit * rayDirection;
    F3 nextNormal = ffNormal;

    if (dot(geometryNormal, rayDirection) > 0.f || mat.lightIntensity[0] > 0.01f) {
        prdRadiance.done = true;
        return;
    }

    float tl[4], tp[4], te[4];
    tex2D(mat.lambert, texcoord.x, texcoord.y, tl);
    tex2D(mat.phong, texcoord.x, texcoord.y, tp);
    tex2D(mat.exponent, texcoord.x, texcoord.y, te);
    F3 lambertReflectance = mk3(tl[0], tl[1], tl[2]);
    F3 phongReflectance = mk3(tp[0], tp[1], tp[2]);
    float phongExponent = te[0];

    const unsigned int index = prdRadiance.photonIndex;
    F3 direction;
    float pdfW;

    float maxLambert = MaxColor(lambertReflectance);
    float maxPhong = MaxColor(phongReflectance);
    if (maxLambert + maxPhong <= 0.000001f) {
        prdRadiance.done = true;
        return;
    }

    setv(photons[index].fluxDir, -rayDirection);
    setv(photons[index].position, nextPosition);
    setv(photons[index].normal, nextNormal);
    setv(photons[index].flux, prdRadiance.flux);
    setv(photons[index].lambertReflectance, lambertReflectance);
    setv(photons[index].phongReflectance, phongReflectance);
    photons[index].phongExponent = phongExponent;
    photons[index].flags = (uint32_t)prdRadiance.flag;

    float pSelectLambert = maxLambert / (maxPhong + maxLambert);
    float chooseMaterial = fminf(curand_uniform(prdRadiance.rngState), 0.999999f);
    photons[index].pSelectLambert = pSelectLambert;

    float russian = russianProb(prdRadiance.flux);
    prdRadiance.flux /= russian;
    prdRadiance.done = (curand_uniform(prdRadiance.rngState) >= russian);
    if (prdRadiance.done) { return; }

    if (chooseMaterial < pSelectLambert) {
        prdRadiance.flux *= LambertSample(&direction, &pdfW, -rayDirection, nextNormal, lambertReflectance,
                                          prdRadiance.rngState) / pSelectLambert;
        photons[index].flags = (uint32_t)(prdRadiance.flag | EVPLP_FLAG_LAMBERT_ONLY);
    } else {
        prdRadiance.flux *= PhongSample(&direction, &pdfW, -rayDirection, geometryNormal, phongReflectance,
                                        phongExponent, prdRadiance.rngState) / (1.0f - pSelectLambert);
        photons[index].flags = (uint32_t)(prdRadiance.flag | EVPLP_FLAG_PHONG_ONLY);
    }
    prdRadiance.nextPosition = nextPosition;
    prdRadiance.nextDirection = direction;
}

// tracePhotons (:192-250) for launch index `launchId`; records at photons[pmIndex..].
inline void tracePhotons(const Scene& s, EvplpRecord* photons, unsigned int launchId, unsigned int slot,
                         unsigned int numPhotonsPerLightPath, unsigned int rngSeed) {
    unsigned int pmIndex = slot * numPhotonsPerLightPath;
    for (unsigned int i = 0; i < numPhotonsPerLightPath; i++) photons[i + pmIndex].flags = 0;

    CurandState localState;
    curand_init(launchId, rngSeed, 0, &localState);

    F3 position, normal;
    float pdf;
    F3 flux = LightSample(s, &position, &normal, &pdf, &localState);

    F3 direction;
    float phongPdf;
    F3 att = PhongSample(&direction, &phongPdf, normal, normal, mk3(1.0f), s.lightIntensity[3], &localState);

    EvplpRecord& photon = photons[pmIndex];
    setv(photon.position, position);
    setv(photon.normal, normal);
    setv(photon.flux, flux);
    photon.flags = EVPLP_FLAG_USABLE_VPL;
    photon.pSelectLambert = 0.0f;
    setv(photon.lambertReflectance, mk3(0.0f));
    setv(photon.phongReflectance, mk3(1.0f));
    photon.phongExponent = s.lightIntensity[3];
    setv(photon.fluxDir, normal);

    PerRayData_radiance prd;
    prd.rngState = &localState;
    prd.flux = flux * att;
    prd.done = false;
    prd.nextPosition = position;
    prd.nextDirection = direction;

    for (unsigned int i = 1; i < numPhotonsPerLightPath; i++) {
        const F3 rayOrigin = prd.nextPosition, rayDirection = prd.nextDirection;
        prd.photonIndex = pmIndex + i;
        if (i != numPhotonsPerLightPath - 1) {
            prd.flag = EVPLP_FLAG_USABLE_VPL | EVPLP_FLAG_USABLE_PHOTON;
        } else {
            prd.flag = EVPLP_FLAG_USABLE_PHOTON;
        }
        // Ray ray(origin, direction, 0, 0.0001f) with tmax = RT_DEFAULT_MAX = 1e27f
        Hit hit = trace_closest(s, rayOrigin, rayDirection, 0.0001f, 1e27f);
        if (hit.prim < 0) { break; }  // miss program: none -> prd.done stays false, but nothing else happens;
                                      // the next rtTrace would re-trace the same ray: same miss. Equivalent to break.
        rtMaterialClosestHit(s, photons, prd, rayOrigin, rayDirection, hit);
        if (prd.done) { break; }
    }
}

// ------------------------------- G-buffer -------------------------------------------
struct GPixel {
    F3 position; float w;
    F3 normal;
    F3 lambert;
    F3 phong; float exponent;
    int prim;
};

inline F3 primary_dir(const EvplpParams& P, int W, int H, int x, int y) {
    float cx = ((float)x + 0.5f) / (float)W * 2.0f - 1.0f;
    float cy = ((float)y + 0.5f) / (float)H * 2.0f - 1.0f;
    float nx = (cx - P.jitter[0]) * P.tanHalfFovX;
    float ny = (cy - P.jitter[1]) * P.tanHalfFovY;
    return getv(P.camForward) + getv(P.camRight) * nx + getv(P.camUp) * ny;
}

inline GPixel gbuffer_pixel(const Scene& s, const EvplpParams& P, int W, int H, int x, int y) {
    GPixel g;
    g.position = mk3(0.f); g.w = 1.0f; g.normal = mk3(0.f); g.lambert = mk3(0.f); g.phong = mk3(0.f);
    g.exponent = 0.f; g.prim = -1;
    F3 org = getv(P.cameraPosition);
    F3 dir = primary_dir(P, W, H, x, y);
    Hit hit = trace_closest(s, org, dir, P.nearDist, P.farDist);
    if (hit.prim < 0) return g;
    const Tri& tri = s.tris[hit.prim];
    const Material& mat = s.mats[tri.mat];
    float w0 = 1.0f - hit.beta - hit.gamma;
    g.position = tri.p0 * w0 + tri.p1 * hit.beta + tri.p2 * hit.gamma;   // interpolated world position (deferred.geom:23)
    g.normal = normalize(cross(tri.p1 - tri.p0, tri.p2 - tri.p0));       // deferred.geom:16-18, not face-forwarded
    float u = tri.t0.x * w0 + tri.t1.x * hit.beta + tri.t2.x * hit.gamma;
    float v = tri.t0.y * w0 + tri.t1.y * hit.beta + tri.t2.y * hit.gamma;
    float tl[4], tp[4], te[4];
    tex2D(mat.lambert, u, v, tl);
    tex2D(mat.phong, u, v, tp);
    tex2D(mat.exponent, u, v, te);
    g.lambert = mk3(tl[0], tl[1], tl[2]);
    g.phong = mk3(tp[0], tp[1], tp[2]);
    g.exponent = te[0];
    g.prim = hit.prim;
    return g;
}

// ------------------------------- lighttracing.cu: VPL gather ------------------------
inline float BalanceHeuristic(const float pdfA, const float pdfB) { return pdfA / (pdfA + pdfB); }
inline float MaxHeuristic(const float pdfA, const float pdfB) { return pdfA > pdfB ? 1.f : 0.f; }
inline float PowerHeuristic2(const float pdfA, const float pdfB) {
    float pdfA2 = pdfA * pdfA;
    float pdfB2 = pdfB * pdfB;
    return BalanceHeuristic(pdfA2, pdfB2);
}

struct GatherCounters {
    uint64_t pairs = 0, shadowRays = 0;
};

inline F3 vplSplat(const Scene& s, const EvplpParams& P, const F3& wi10, const F3& firstPosition,
                   const F3& firstNormal, const F3& firstLambertReflectance, const F3& firstPhongReflectance,
                   const float firstPhongExponent, const EvplpRecord& rec, GatherCounters* cnt) {  // :275-346
    const F3 recPos = getv(rec.position), recN = getv(rec.normal), recFlux = getv(rec.flux);
    const F3 recKd = getv(rec.lambertReflectance), recKs = getv(rec.phongReflectance);
    F3 v12 = recPos - firstPosition;

    float unnormCos1 = fmaxf(dot(firstNormal, v12), 0.0f);
    float unnormCos2 = fmaxf(-dot(recN, v12), 0.0f);
    float unnormCos1Cos2 = unnormCos1 * unnormCos2;
    if (cnt) cnt->pairs++;
    if (unnormCos1Cos2 <= 0.000f) { return mk3(0.0f); }

    if (cnt) cnt->shadowRays++;
    // Ray ray(photonRecord.mPosition, -v12, 1, 0.0001, 1 - 0.0001): double literals narrowed to float
    if (trace_any(s, recPos, -v12, (float)0.0001, (float)(1 - 0.0001))) { return mk3(0.0f); }

    float dist2 = dot(v12, v12);
    float dist = sqrtf(dist2);

    F3 wi12 = v12 / dist;
    F3 incomingDir = getv(rec.fluxDir);

    F3 brdf2 = LambertEvalF(-wi12, incomingDir, recN) * recKd +
               PhongEvalF(-wi12, incomingDir, recN, rec.phongExponent) * recKs;
    F3 brdf1 = LambertEvalF(wi10, wi12, firstNormal) * firstLambertReflectance +
               PhongEvalF(wi10, wi12, firstNormal, firstPhongExponent) * firstPhongReflectance;

    float g21 = unnormCos1Cos2 / (dist2 * dist2);

    const unsigned misMode = P.misMode;
    if (misMode == 0) {
        return recFlux * brdf1 * brdf2 * g21;
    } else if (misMode == 1 || misMode == 2 || misMode == 3) {
        float pdfDe = LambertPdfA(recN, firstNormal, -v12) * rec.pSelectLambert;
        pdfDe += PhongPdfA(recN, firstNormal, -v12, incomingDir, recKs, rec.phongExponent) * (1.0f - rec.pSelectLambert);
        float weight = misMode == 1 ? BalanceHeuristic(P.pdfMc, pdfDe)
                     : misMode == 2 ? MaxHeuristic(P.pdfMc, pdfDe) : PowerHeuristic2(P.pdfMc, pdfDe);
        return weight * recFlux * brdf1 * brdf2 * g21;
    } else if (misMode == 4) {
        return recFlux * fminf(g21, P.clampingValue) * brdf1 * brdf2;
    } else {
        F3 gb = g21 * brdf1 * brdf2;
        return recFlux * fminf3(gb, mk3(P.clampingValue));
    }
}

// splatColor (:348-379): returns result / numVplLightPaths for one pixel (the caller accumulates).
inline F3 splatColor(const Scene& s, const EvplpParams& P, const GPixel& g, const EvplpRecord* photons,
                     GatherCounters* cnt) {
    if (g.w == 0.0f) return mk3(0.f);
    F3 wi01 = normalize(getv(P.cameraPosition) - g.position);
    F3 result = mk3(0.0f);
    unsigned numPhotons = P.numPhotonsPerLightPath * P.numVplLightPaths;
    for (unsigned i = 0; i < numPhotons; i++) {
        if ((photons[i].flags & EVPLP_FLAG_USABLE_VPL) != 0) {
            result += vplSplat(s, P, wi01, g.position, g.normal, g.lambert, g.phong, g.exponent, photons[i], cnt);
        }
    }
    return result / (float)P.numVplLightPaths;
}

// LVC splatColor (lvclighttracing.cu:348-387)
inline F3 splatColorLvc(const Scene& s, const EvplpParams& P, const GPixel& g, const EvplpRecord* photons,
                        unsigned launchId, GatherCounters* cnt) {
    if (g.w == 0.0f) return mk3(0.f);
    F3 wi01 = normalize(getv(P.cameraPosition) - g.position);
    F3 result = mk3(0.0f);
    CurandState localState;
    curand_init(launchId, P.rngSeed, 0, &localState);
    unsigned int lightPathOffset = (unsigned int)(fminf(curand_uniform(&localState), 0.999999f) * P.numLightPaths);
    for (unsigned i = 0; i < P.numVplLightPaths; i++) {
        unsigned int lightPathId = (i + lightPathOffset) % P.numLightPaths;
        unsigned int lightVertexOffset = lightPathId * P.numPhotonsPerLightPath;
        for (unsigned j = 0; j < P.numPhotonsPerLightPath; j++) {
            if ((photons[lightVertexOffset + j].flags & EVPLP_FLAG_USABLE_VPL) != 0) {
                result += vplSplat(s, P, wi01, g.position, g.normal, g.lambert, g.phong, g.exponent,
                                   photons[lightVertexOffset + j], cnt);
            }
        }
    }
    return result / (float)P.numVplLightPaths;
}

// ------------------------------- lighttracing.cu: VSL -------------------------------
inline F3 SquareToSolidAngle(const float sampleX, const float sampleY, const float halfAngleMax) {  // :382-390
    const float phi = 2.0f * M_PIf_ * sampleX;
    const float z = 1.0f - sampleY * (1.0f - det_cosf(halfAngleMax));
    const float l = sqrtf(1.0f - z * z);
    const float cosphi = det_cosf(phi);
    const float sinphi = det_sinf(phi);
    return mk3(cosphi * l, sinphi * l, z);
}

struct VslRec {
    F3 pos, normal, flux, fluxDir, kd, ks;
    float exponent;
};

inline F3 sampleCone(const EvplpParams& P, float* misWeight, const F3& wi01, const F3& normal, const F3& lambertRefl,
                     const F3& phongRefl, const float phongExp, const VslRec& rec, const float halfCone,
                     const float solidAngle, const float invSolidAngle, const F3& nd12, CurandState* rngState) {  // :395-446
    float maxLambert = MaxColor(lambertRefl);
    float maxPhong = MaxColor(phongRefl);
    if (maxLambert + maxPhong <= 0.000001f) { return mk3(0.0f); }
    float pSelectLambert = maxLambert / (maxPhong + maxLambert);
    float chooseMaterial = fminf(curand_uniform(rngState), 0.999999f);
    (void)chooseMaterial;
    // two curand_uniform() as call arguments at :417: oracle defines left-to-right
    float sx = curand_uniform(rngState);
    float sy = curand_uniform(rngState);
    F3 wi12 = normalize(SquareToSolidAngle(sx, sy, halfCone));
    Onb onb(nd12);
    onb.inverse_transform(wi12);
    wi12 = normalize(wi12);

    const float cos1cos2 = fmaxf(dot(normal, wi12), 0.0f) * fmaxf(-dot(rec.normal, wi12), 0.0f);
    if (cos1cos2 <= 0.000000001f) { return mk3(0.0f); }

    F3 incomingDir = rec.fluxDir;
    F3 brdf2 = LambertEvalF(-wi12, incomingDir, rec.normal) * rec.kd +
               PhongEvalF(-wi12, incomingDir, rec.normal, rec.exponent) * rec.ks;
    F3 brdf1 = LambertEvalF(wi01, wi12, normal) * lambertRefl + PhongEvalF(wi01, wi12, normal, phongExp) * phongRefl;
    float pdfCone = invSolidAngle;
    float pdfBrdf1 = LambertPdfW(normal, wi12) * pSelectLambert +
                     PhongPdfW(normal, wi12, wi01, phongRefl, phongExp) * (1.0f - pSelectLambert);
    float pdfBrdf2 = LambertPdfW(rec.normal, -wi12) * pSelectLambert +
                     PhongPdfW(rec.normal, -wi12, rec.fluxDir, rec.ks, rec.exponent);
    *misWeight = pdfCone / (pdfBrdf1 + pdfBrdf2 + pdfCone);
    return rec.flux * P.vslInvPiRadius2 * cos1cos2 * brdf1 * brdf2 * solidAngle;
}

inline F3 sampleBrdf1(const EvplpParams& P, float* misWeight, const F3& wi01, const F3& normal, const F3& lambertRefl,
                      const F3& phongRefl, const float phongExp, const VslRec& rec, const float cosHalfCone,
                      const float invSolidAngle, const F3& nd12, CurandState* rngState) {  // :448-521
    F3 wi12;
    F3 brdf1;
    float pdfW;
    {
        float maxLambert = MaxColor(lambertRefl);
        float maxPhong = MaxColor(phongRefl);
        if (maxLambert + maxPhong <= 0.000001f) { return mk3(0.0f); }
        float pSelectLambert = maxLambert / (maxPhong + maxLambert);
        float chooseMaterial = fminf(curand_uniform(rngState), 0.999999f);
        if (chooseMaterial < pSelectLambert) {
            brdf1 = LambertSample(&wi12, &pdfW, wi01, normal, lambertRefl, rngState) / pSelectLambert;
        } else {
            brdf1 = PhongSample(&wi12, &pdfW, wi01, normal, phongRefl, phongExp, rngState) / (1.0f - pSelectLambert);
        }
    }
    if (dot(wi12, nd12) <= cosHalfCone) { return mk3(0.0f); }
    const float cos1 = fmaxf(dot(normal, wi12), 0.0f);
    if (cos1 <= 0.000000001f) { return mk3(0.0f); }
    const float cos2 = fmaxf(-dot(rec.normal, wi12), 0.0f);
    F3 incomingDir = rec.fluxDir;
    F3 brdf2 = LambertEvalF(-wi12, incomingDir, rec.normal) * rec.kd +
               PhongEvalF(-wi12, incomingDir, rec.normal, rec.exponent) * rec.ks;
    {
        float maxLambert = MaxColor(lambertRefl);
        float maxPhong = MaxColor(phongRefl);
        if (maxLambert + maxPhong <= 0.000001f) { return mk3(0.0f); }
        float pSelectLambert = maxLambert / (maxPhong + maxLambert);
        float chooseMaterial = fminf(curand_uniform(rngState), 0.999999f);
        (void)chooseMaterial;
        float pdfCone = invSolidAngle;
        float pdfBrdf1 = LambertPdfW(normal, wi12) * pSelectLambert +
                         PhongPdfW(normal, wi12, wi01, phongRefl, phongExp) * (1.0f - pSelectLambert);
        float pdfBrdf2 = LambertPdfW(rec.normal, -wi12) * pSelectLambert +
                         PhongPdfW(rec.normal, -wi12, rec.fluxDir, rec.ks, rec.exponent);
        *misWeight = pdfBrdf1 / (pdfBrdf1 + pdfBrdf2 + pdfCone);
    }
    return rec.flux * P.vslInvPiRadius2 * cos2 * brdf1 * brdf2;
}

inline F3 sampleBrdf2(const EvplpParams& P, float* misWeight, const F3& wi10, const F3& normal, const F3& lambertRefl,
                      const F3& phongRefl, const float phongExp, const VslRec& rec, const float cosHalfCone,
                      const float invSolidAngle, const F3& nd12, CurandState* rngState) {  // :523-594
    F3 wi21;
    F3 brdf2;
    const F3& incomingDir = rec.fluxDir;
    float pdfW;
    {
        float maxLambert = MaxColor(rec.kd);
        float maxPhong = MaxColor(rec.ks);
        if (maxLambert + maxPhong <= 0.000001f) { return mk3(0.0f); }
        float pSelectLambert = maxLambert / (maxPhong + maxLambert);
        float chooseMaterial = fminf(curand_uniform(rngState), 0.999999f);
        if (chooseMaterial < pSelectLambert) {
            brdf2 = LambertSample(&wi21, &pdfW, incomingDir, rec.normal, rec.kd, rngState) / pSelectLambert;
        } else {
            brdf2 = PhongSample(&wi21, &pdfW, incomingDir, rec.normal, rec.ks, rec.exponent, rngState) / (1.0f - pSelectLambert);
        }
    }
    if (-dot(wi21, nd12) <= cosHalfCone) { return mk3(0.0f); }
    F3 brdf1 = LambertEvalF(wi10, -wi21, normal) * lambertRefl + PhongEvalF(wi10, -wi21, normal, phongExp) * phongRefl;
    const float cos2 = fmaxf(dot(rec.normal, wi21), 0.0f);
    if (cos2 <= 0.00000001f) { return mk3(0.0f); }
    const float cos1 = fmaxf(-dot(normal, wi21), 0.0f);
    {
        float maxLambert = MaxColor(lambertRefl);
        float maxPhong = MaxColor(phongRefl);
        if (maxLambert + maxPhong <= 0.000001f) { return mk3(0.0f); }
        float pSelectLambert = maxLambert / (maxPhong + maxLambert);
        float chooseMaterial = fminf(curand_uniform(rngState), 0.999999f);
        (void)chooseMaterial;
        float pdfCone = invSolidAngle;
        float pdfBrdf1 = LambertPdfW(normal, -wi21) * pSelectLambert +
                         PhongPdfW(normal, -wi21, wi10, phongRefl, phongExp) * (1.0f - pSelectLambert);
        float pdfBrdf2 = LambertPdfW(rec.normal, wi21) * pSelectLambert +
                         PhongPdfW(rec.normal, wi21, rec.fluxDir, rec.ks, rec.exponent);
        *misWeight = pdfBrdf2 / (pdfBrdf1 + pdfBrdf2 + pdfCone);
    }
    return rec.flux * P.vslInvPiRadius2 * cos1 * brdf1 * brdf2;
}

inline F3 vslSplat(const Scene& s, const EvplpParams& P, const F3& wi10, const F3& firstPosition, const F3& firstNormal,
                   const F3& firstLambertReflectance, const F3& firstPhongReflectance, const float firstPhongExponent,
                   const EvplpRecord& photonRecord, CurandState* localState, GatherCounters* cnt) {  // :596-686
    VslRec rec;
    rec.pos = getv(photonRecord.position); rec.normal = getv(photonRecord.normal); rec.flux = getv(photonRecord.flux);
    rec.fluxDir = getv(photonRecord.fluxDir); rec.kd = getv(photonRecord.lambertReflectance);
    rec.ks = getv(photonRecord.phongReflectance); rec.exponent = photonRecord.phongExponent;

    F3 v12 = rec.pos - firstPosition;
    float dist2 = dot(v12, v12);
    float dist = sqrtf(dist2);
    if (cnt) { cnt->pairs++; cnt->shadowRays++; }
    if (trace_any(s, rec.pos, -v12, (float)0.0001, (float)(1 - 0.0001))) { return mk3(0.0f); }

    F3 nv12 = v12 / dist;
    const float cos1cos2 = fmaxf(dot(firstNormal, nv12), 0.0f) * fmaxf(-dot(rec.normal, nv12), 0.0f);
    if (cos1cos2 <= 0.000000001f) { return mk3(0.0f); }

    const float rdratio = P.vslRadius / dist;
    const float halfCone = (rdratio >= 1.0) ? M_PIf_ / 2.0f : det_asinf(rdratio);
    const float cosHalfCone = det_cosf(halfCone);
    const float solidAngle = M_PIf_ * 2.0f * (1.0f - cosHalfCone);
    const float invSolidAngle = 1.0f / solidAngle;

    F3 result = mk3(0.0f);
    int numSamples = (int)(halfCone / M_PIf_ * 2.0f * 100.0f) + 1;
    for (int i = 0; i < numSamples; i++) {
        float coneWeight = 0.0f;
        F3 coneResult = sampleCone(P, &coneWeight, wi10, firstNormal, firstLambertReflectance, firstPhongReflectance,
                                   firstPhongExponent, rec, halfCone, solidAngle, invSolidAngle, nv12, localState);
        float brdf1Weight = 0.0f;
        F3 brdf1Result = sampleBrdf1(P, &brdf1Weight, wi10, firstNormal, firstLambertReflectance, firstPhongReflectance,
                                     firstPhongExponent, rec, cosHalfCone, invSolidAngle, nv12, localState);
        float brdf2Weight = 0.0f;
        F3 brdf2Result = sampleBrdf2(P, &brdf2Weight, wi10, firstNormal, firstLambertReflectance, firstPhongReflectance,
                                     firstPhongExponent, rec, cosHalfCone, invSolidAngle, nv12, localState);
        result += coneWeight * coneResult;
        result += brdf1Weight * brdf1Result;
        result += brdf2Weight * brdf2Result;
    }
    return result / (float)numSamples;
}

// splatSplotch (:689-722)
inline F3 splatSplotch(const Scene& s, const EvplpParams& P, const GPixel& g, const EvplpRecord* photons,
                       unsigned launchId, GatherCounters* cnt) {
    F3 wi10 = normalize(getv(P.cameraPosition) - g.position);
    F3 result = mk3(0.0f);
    unsigned numPhotons = P.numPhotonsPerLightPath * P.numVplLightPaths;
    CurandState localState;
    curand_init(launchId, P.rngSeed, 0, &localState);
    for (unsigned i = 0; i < numPhotons; i++) {
        if ((photons[i].flags & EVPLP_FLAG_USABLE_VPL) != 0) {
            result += vslSplat(s, P, wi10, g.position, g.normal, g.lambert, g.phong, g.exponent, photons[i],
                               &localState, cnt);
        }
    }
    return result / (float)P.numVplLightPaths;
}

// ------------------------------- pathtracing.cu (RtPt2) ----------------------------
// The reference's own ground-truth technique: pathtracing.cu:49-52, 85-95, 112-218, 232-377.
namespace pt {
inline float russianProb(const F3& throughput) {  // :49-52 (sic: never below 0.98)
    return fmaxf(fmaxf(throughput.x, 0.98f), fmaxf(throughput.y, throughput.z));
}
inline float MisWeight(const float pdf1, const float pdf2) { return pdf1 / (pdf1 + pdf2); }  // :85-88
inline float pdfW2A(const F3& n2, const F3& v12) {                                           // :90-94
    F3 nv12 = normalize(v12);
    return fmaxf(-dot(n2, nv12), 0.f) / dot(v12, v12);
}
inline float GeometryTerm(const F3& n1, const F3& n2, const F3 v12) {  // rtmaterial.cuh:30-38
    const float cos1Unnorm = fmaxf(dot(n1, v12), 0.f);
    const float cos2Unnorm = fmaxf(-dot(n2, v12), 0.f);
    const float d2 = dot(v12, v12);
    return cos1Unnorm * cos2Unnorm / (d2 * d2);
}
inline F3 LambertEval(const F3&, const F3&, const F3&, const F3& lambertReflectance) { return lambertReflectance * M_Inv_PIf; }  // :69-72
inline F3 PhongEval(const F3& out, const F3& in, const F3& normal, const F3& phongReflectance, const float phongExponent) {  // :104-111
    F3 reflectVec = reflect(-in, normal);
    float dotWrWo = fmaxf(dot(out, reflectVec), 0.0f);
    if (dotWrWo <= 0.000001f || phongReflectance.x <= 0.000001f) { return mk3(0.0f); }
    return phongReflectance * (phongExponent + 2.0f) * det_powf(dotWrWo, phongExponent) * (M_Inv_PIf) * 0.5f;
}

struct PerRayData_radiance {
    bool done;
    CurandState* rngState;
    float brdfPdfW;
    F3 result, position, direction, attenuation;
};

// rtMaterialClosestHit (:112-218)
inline void rtMaterialClosestHit(const Scene& s, PerRayData_radiance& prdRadiance, const F3& rayOrigin, const F3& rayDirection,
                                 const Hit& hit, uint64_t* rays) {
    const Tri& tri = s.tris[hit.prim];
    const Material& mat = s.mats[tri.mat];
    F3 geometryNormal = normalize(hit.n);
    F2 texcoord;
    {
        float w0 = 1.0f - hit.beta - hit.gamma;
        texcoord.x = tri.t1.x * hit.beta + tri.t2.x * hit.gamma + tri.t0.x * w0;
        texcoord.y = tri.t1.y * hit.beta + tri.t2.y * hit.gamma + tri.t0.y * w0;
    }
    F3 worldGeometryNormal = normalize(geometryNormal);
    F3 ffNormal = faceforward(worldGeometryNormal, -rayDirection, worldGeometryNormal);
    F3 nextPosition = rayOrigin + hit.t * rayDirection;

    if (dot(geometryNormal, rayDirection) > 0.f) {
        prdRadiance.result = mk3(0.0f);
        prdRadiance.done = true;
        return;
    }
    if (mat.lightIntensity[0] > 0.01f) {
        float brdfPdfA = (prdRadiance.brdfPdfW * pdfW2A(ffNormal, nextPosition - prdRadiance.position));
        float lightPdfA = 1.f / s.lightArea;
        float weight = MisWeight(brdfPdfA, lightPdfA);
        prdRadiance.result = weight * prdRadiance.attenuation *
                             PhongEvalF(geometryNormal, normalize(prdRadiance.position - nextPosition), geometryNormal, mat.lightIntensity[3]) *
                             mk3(mat.lightIntensity[0], mat.lightIntensity[1], mat.lightIntensity[2]);
        prdRadiance.done = true;
        return;
    }
    if (prdRadiance.done) { return; }

    float lightPdf;
    F3 lightPosition, lightNormal;
    F3 lightValue = LightSample(s, &lightPosition, &lightNormal, &lightPdf, prdRadiance.rngState);
    F3 toLight = lightPosition - nextPosition;
    F3 toLightNorm = normalize(toLight);
    (*rays)++;
    bool shadowHit = trace_any(s, lightPosition, -toLight, (float)0.00001, (float)0.99999);

    float tl[4], tp[4], te[4];
    tex2D(mat.lambert, texcoord.x, texcoord.y, tl);
    tex2D(mat.phong, texcoord.x, texcoord.y, tp);
    tex2D(mat.exponent, texcoord.x, texcoord.y, te);
    F3 lambertReflectance = mk3(tl[0], tl[1], tl[2]);
    F3 phongReflectance = mk3(tp[0], tp[1], tp[2]);
    float phongExponent = te[0];

    float maxLambert = MaxColor(lambertReflectance);
    float maxPhong = MaxColor(phongReflectance);
    prdRadiance.done = (maxLambert + maxPhong <= 0.000001f);
    if (prdRadiance.done) { return; }

    float pSelectLambert = maxLambert / (maxPhong + maxLambert);
    float chooseMaterial = fminf(curand_uniform(prdRadiance.rngState), 0.999999f);
    const float lightExp = s.lightIntensity[3];
    if (chooseMaterial < pSelectLambert) {
        if (!shadowHit) {
            float brdfPdf = LambertPdfA(ffNormal, lightNormal, toLight);
            float weight = MisWeight(lightPdf, brdfPdf);
            prdRadiance.result = weight * lightValue *
                                 LambertEval(toLightNorm, normalize(prdRadiance.position - nextPosition), ffNormal, lambertReflectance) *
                                 GeometryTerm(ffNormal, lightNormal, toLight) * prdRadiance.attenuation / pSelectLambert *
                                 PhongEvalF(lightNormal, -toLightNorm, lightNormal, lightExp);
        }
        prdRadiance.attenuation *= LambertSample(&prdRadiance.direction, &prdRadiance.brdfPdfW,
                                                 normalize(prdRadiance.position - nextPosition), geometryNormal, lambertReflectance,
                                                 prdRadiance.rngState) / pSelectLambert;
    } else {
        if (!shadowHit) {
            float brdfPdf = PhongPdfA(ffNormal, lightNormal, toLight, normalize(prdRadiance.position - nextPosition), phongReflectance, phongExponent);
            float weight = MisWeight(lightPdf, brdfPdf);
            prdRadiance.result = weight * lightValue *
                                 PhongEval(toLightNorm, normalize(prdRadiance.position - nextPosition), ffNormal, phongReflectance, phongExponent) *
                                 GeometryTerm(ffNormal, lightNormal, toLight) * prdRadiance.attenuation / (1.0f - pSelectLambert) *
                                 PhongEvalF(lightNormal, -toLightNorm, lightNormal, lightExp);
        }
        prdRadiance.attenuation *= PhongSample(&prdRadiance.direction, &prdRadiance.brdfPdfW,
                                               normalize(prdRadiance.position - nextPosition), geometryNormal, phongReflectance,
                                               phongExponent, prdRadiance.rngState) / (1.0f - pSelectLambert);
    }
    float russian = pt::russianProb(prdRadiance.attenuation);
    prdRadiance.done = (curand_uniform(prdRadiance.rngState) >= russian);
    if (prdRadiance.done) { return; }
    prdRadiance.position = nextPosition;
    prdRadiance.attenuation /= russian;
}

// pathTraceSimple (:232-321) with numSamples = 1
inline F3 pathTraceSimple(const Scene& s, const F3& cameraPos, const F3& firstPosition, const F3& firstNormal,
                          const F3& firstLambertReflectance, const F3& firstPhongReflectance, const float firstPhongExponent,
                          unsigned maxBounces, CurandState* rngState, uint64_t* rays) {
    F3 cameraVec = normalize(firstPosition - cameraPos);
    F3 result = mk3(0.0f);
    PerRayData_radiance prd;
    prd.rngState = rngState;
    F3 position = firstPosition;
    F3 normal = firstNormal;
    prd.position = firstPosition;
    prd.attenuation = mk3(1.0f);
    prd.brdfPdfW = 0.f;
    prd.direction = mk3(0.f);
    const float lightExp = s.lightIntensity[3];
    {
        float lightPdf;
        F3 lightPosition, lightNormal;
        F3 lightValue = LightSample(s, &lightPosition, &lightNormal, &lightPdf, rngState);
        F3 toLight = lightPosition - position;
        F3 toLightNorm = normalize(toLight);
        (*rays)++;
        bool shadowHit = trace_any(s, lightPosition, -toLight, 0.0001f, 1.0f - 0.0001f);
        float maxLambert = MaxColor(firstLambertReflectance);
        float maxPhong = MaxColor(firstPhongReflectance);
        float pSelectLambert = maxLambert / (maxPhong + maxLambert);
        if (maxLambert + maxPhong <= 0.000001f) { return mk3(0.0f); }
        float chooseMaterial = fminf(curand_uniform(prd.rngState), 0.999999f);
        if (chooseMaterial < pSelectLambert) {
            if (!shadowHit) {
                float brdfPdf = LambertPdfA(normal, lightNormal, toLight);
                float weight = MisWeight(lightPdf, brdfPdf);
                result += weight * lightValue * LambertEval(-cameraVec, toLightNorm, normal, firstLambertReflectance) *
                          GeometryTerm(normal, lightNormal, toLight) / pSelectLambert *
                          PhongEvalF(lightNormal, -toLightNorm, lightNormal, lightExp);
            }
            prd.attenuation *= LambertSample(&prd.direction, &prd.brdfPdfW, -cameraVec, normal, firstLambertReflectance, prd.rngState) / pSelectLambert;
        } else {
            if (!shadowHit) {
                float brdfPdf = PhongPdfA(normal, lightNormal, toLight, -cameraVec, firstPhongReflectance, firstPhongExponent);
                float weight = MisWeight(lightPdf, brdfPdf);
                result += weight * lightValue * PhongEval(-cameraVec, toLightNorm, normal, firstPhongReflectance, firstPhongExponent) *
                          GeometryTerm(normal, lightNormal, toLight) / (1.0f - pSelectLambert) *
                          PhongEvalF(lightNormal, -toLightNorm, lightNormal, lightExp);
            }
            prd.attenuation *= PhongSample(&prd.direction, &prd.brdfPdfW, -cameraVec, normal, firstPhongReflectance, firstPhongExponent,
                                           prd.rngState) / (1.0f - pSelectLambert);
        }
    }
    for (size_t i = 0; i < maxBounces; i++) {
        prd.done = (i == maxBounces - 1);
        prd.result = mk3(0.f);
        const F3 rayOrigin = prd.position, rayDirection = prd.direction;
        (*rays)++;
        Hit hit = trace_closest(s, rayOrigin, rayDirection, (float)0.00001, 1e27f);
        if (hit.prim < 0) { break; }  // no miss program: the same miss repeats until the last bounce, nothing is added
        rtMaterialClosestHit(s, prd, rayOrigin, rayDirection, hit, rays);
        result += prd.result;
        if (prd.done) { break; }
    }
    return result;
}

// splatColor (:323-354): result of one pixel (the caller accumulates)
inline F3 splatColor(const Scene& s, const EvplpParams& P, const GPixel& g, unsigned launchId, unsigned maxBounces, uint64_t* rays) {
    if (g.w == 0.0f) return mk3(0.f);
    CurandState localState;
    curand_init(launchId, P.rngSeed, 0, &localState);
    return pathTraceSimple(s, getv(P.cameraPosition), g.position, g.normal, g.lambert, g.phong, g.exponent, maxBounces, &localState, rays);
}
}  // namespace pt

// ------------------------------- photonsplatinstanced.frag --------------------------
namespace glsl {
inline F3 LambertEval(const F3& w10, const F3& w12, const F3& normal, const F3& lambertReflectance) {  // .frag:36-44
    if (dot(w10, normal) <= 0.0f || dot(w12, normal) <= 0.0f) { return mk3(0.0f); }
    return M_Inv_PIf * lambertReflectance;
}
inline F3 PhongEval(const F3& outVec, const F3& inVec, const F3& normal, const F3& phongReflectance,
                    const float phongExponent) {  // .frag:46-52
    F3 reflectVec = reflect(-inVec, normal);
    float dotWrWo = dot(outVec, reflectVec);
    if (dotWrWo <= 0.00001f) { return mk3(0.0f); }
    return phongReflectance * (phongExponent + 2.0f) * det_powf(dotWrWo, phongExponent) * M_Inv_PIf * 0.5f;
}
inline float LambertPdfW(const F3& normal1, const F3& v12) {  // .frag:59-63
    float cos1Unnorm = fmaxf(dot(normal1, normalize(v12)), 0.f);
    return cos1Unnorm * M_Inv_PIf;
}
inline float PhongPdfW(const F3& normal1, const F3& wi12, const F3& inVec, const F3& phongReflectance,
                       const float phongExponent) {  // .frag:73-79
    F3 reflectVec = reflect(-inVec, normal1);
    float dotWrWo = fmaxf(dot(wi12, reflectVec), 0.f);
    if (dotWrWo <= 0.00001f || phongReflectance.x <= 0.00001f) { return 0.0f; }
    return (phongExponent + 1.0f) * 0.5f * M_Inv_PIf * det_powf(dotWrWo, phongExponent);
}
}  // namespace glsl

// Fragment of photon `k` (global record index inside `photons`) on G-buffer pixel g.
// Returns false when the fragment is discarded.  (photonsplatinstanced.frag:146-240)
inline bool splatFragment(const EvplpParams& P, const GPixel& g, const EvplpRecord* photons, uint64_t k, F3* color) {
    const EvplpRecord& ph = photons[k];
    F3 shadingPosition = g.position;
    float photonRadius2 = P.radius * P.radius;
    F3 dvec = getv(ph.position) - shadingPosition;
    if (dot(dvec, dvec) > photonRadius2) { return false; }

    F3 shadingNormal = g.normal;
    F3 shadingDiffuseColor = g.lambert;
    F3 shadingPhongReflectance = g.phong;
    float shadingPhongExponent = g.exponent;

    const EvplpRecord& prev = photons[k - 1];
    F3 v12 = getv(prev.position) - getv(ph.position);
    F3 w12 = normalize(v12);
    F3 n1 = getv(ph.normal);

    F3 brdf1 = mk3(0.0f);
    F3 w10 = normalize(getv(P.cameraPosition) - shadingPosition);
    F3 prevFluxDir = getv(prev.fluxDir);
    F3 prevN = getv(prev.normal);

    brdf1 = glsl::LambertEval(w10, w12, shadingNormal, shadingDiffuseColor) +
            glsl::PhongEval(w10, w12, shadingNormal, shadingPhongReflectance, shadingPhongExponent);
    F3 brdf2 = glsl::LambertEval(-w12, prevFluxDir, prevN, getv(prev.lambertReflectance)) +
               glsl::PhongEval(-w12, prevFluxDir, prevN, getv(prev.phongReflectance), prev.phongExponent);

    float mixPdfW = glsl::LambertPdfW(prevN, -w12) * prev.pSelectLambert;
    mixPdfW += glsl::PhongPdfW(prevN, -w12, prevFluxDir, getv(prev.phongReflectance), prev.phongExponent) *
               (1.0f - prev.pSelectLambert);
    float mixPdfA = mixPdfW * fmaxf(dot(n1, w12), 0.0f) / dot(v12, v12);

    const float uInvPhotonRadius2 = 1.0f / (P.radius * P.radius);                 // rtcomphoton.h:819
    const float uInvNumLightPaths = 1.0f / (float)P.numLightPaths;                // rtcomphoton.h:820
    const F3 flux = getv(ph.flux);
    const float InvPi = M_Inv_PIf;

    if (mixPdfW > 0.0f) {
        const unsigned uMisMode = P.misMode;
        if (uMisMode == 0) {
            *color = brdf1 * (InvPi * uInvPhotonRadius2) * flux * uInvNumLightPaths;
        } else if (uMisMode == 1) {
            float weight = BalanceHeuristic(mixPdfA, P.pdfMc);
            *color = brdf1 * (InvPi * uInvPhotonRadius2) * flux * uInvNumLightPaths * weight;
        } else if (uMisMode == 2) {
            float weight = MaxHeuristic(mixPdfA, P.pdfMc);
            *color = brdf1 * (InvPi * uInvPhotonRadius2) * flux * uInvNumLightPaths * weight;
        } else if (uMisMode == 3) {
            float weight = PowerHeuristic2(mixPdfA, P.pdfMc);
            *color = brdf1 * (InvPi * uInvPhotonRadius2) * flux * uInvNumLightPaths * weight;
        } else if (uMisMode == 4) {
            float distance2 = dot(v12, v12);
            float cosCos = fmaxf(dot(shadingNormal, w12), 0.0f) * fmaxf(-dot(prevN, w12), 0.0f);
            if (cosCos <= 0.0f) { return false; }
            float geometryTerm = cosCos / distance2;
            *color = brdf1 * (InvPi * uInvPhotonRadius2) * flux * uInvNumLightPaths *
                     fmaxf(geometryTerm - P.clampingValue, 0.0f) / geometryTerm;
        } else {
            float distance2 = dot(v12, v12);
            float cosCos = fmaxf(dot(shadingNormal, w12), 0.0f) * fmaxf(-dot(prevN, w12), 0.0f);
            if (cosCos <= 0.0f) { return false; }
            float geometryTerm = cosCos / distance2;
            F3 num = fmaxf3((brdf1 * brdf2 * geometryTerm) - mk3(P.clampingValue), mk3(0.0f));
            F3 den = geometryTerm * brdf2;
            F3 pre = (InvPi * uInvPhotonRadius2) * flux * uInvNumLightPaths;
            *color = mk3(pre.x * num.x / den.x, pre.y * num.y / den.y, pre.z * num.z / den.z);
        }
    } else {
        *color = mk3(0.0f);
    }
    return true;
}

// Fixed-point (Q31.32) conversion used by the accumulation layers.  Non-finite values are
// dropped (GL would poison the pixel forever; documented deviation).
inline int64_t to_fixed(float c) {
    if (!(fabsf(c) < 1.0e9f)) return 0;
    return (int64_t)llrintf(c * 4294967296.0f);
}

}  // namespace orc
