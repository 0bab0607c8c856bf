// oracle_capi.cpp -- TEST INFRASTRUCTURE (CPU oracle): C entry points for ctypes.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load liboracle.so.  The product (libevplp_b200.so, evplp_b200/) never does.
// Parity status: pinned by the reference's own device code for the CUDA programs, see oracle_math.h.
#include <omp.h>
#include <stdio.h>
#include <string.h>
#include <cmath>
#include <numeric>
#include "oracle_stages.h"

using namespace orc;

extern "C" {

// ------------------------------------------------------------------ scene ------------
void* orc_scene_create(const EvplpMeshDesc* meshes, int32_t numMeshes, const EvplpMaterialDesc* materials,
                       int32_t numMaterials, int32_t lightMeshIndex, const float lightIntensityPrecomputed[4],
                       const float lightIntensityDisplay[4], int32_t bruteForce) {
    Scene* s = new Scene();
    s->mats.resize(numMaterials);
    for (int m = 0; m < numMaterials; m++) {
        const EvplpMaterialDesc& d = materials[m];
        Material& mt = s->mats[m];
        mt.lambert.w = d.lambertW; mt.lambert.h = d.lambertH;
        mt.lambert.data.assign(d.lambertReflectance, d.lambertReflectance + (size_t)d.lambertW * d.lambertH * 4);
        mt.phong.w = d.phongW; mt.phong.h = d.phongH;
        mt.phong.data.assign(d.phongReflectance, d.phongReflectance + (size_t)d.phongW * d.phongH * 4);
        mt.exponent.w = d.exponentW; mt.exponent.h = d.exponentH;
        mt.exponent.data.assign(d.phongExponent, d.phongExponent + (size_t)d.exponentW * d.exponentH * 4);
        memcpy(mt.lightIntensity, d.lightIntensity, 16);
    }
    for (int mi = 0; mi < numMeshes; mi++) {
        const EvplpMeshDesc& d = meshes[mi];
        if (mi == lightMeshIndex) { s->lightFirst = (int)s->tris.size(); s->lightCount = d.numTriangles; }
        for (int t = 0; t < d.numTriangles; t++) {
            Tri tr;
            int i0 = d.indices[t * 3], i1 = d.indices[t * 3 + 1], i2 = d.indices[t * 3 + 2];
            tr.p0 = mk3(d.vertices[i0 * 3], d.vertices[i0 * 3 + 1], d.vertices[i0 * 3 + 2]);
            tr.p1 = mk3(d.vertices[i1 * 3], d.vertices[i1 * 3 + 1], d.vertices[i1 * 3 + 2]);
            tr.p2 = mk3(d.vertices[i2 * 3], d.vertices[i2 * 3 + 1], d.vertices[i2 * 3 + 2]);
            if (d.texcoords) {
                tr.t0 = F2{d.texcoords[i0 * 2], d.texcoords[i0 * 2 + 1]};
                tr.t1 = F2{d.texcoords[i1 * 2], d.texcoords[i1 * 2 + 1]};
                tr.t2 = F2{d.texcoords[i2 * 2], d.texcoords[i2 * 2 + 1]};
            } else {
                tr.t0 = tr.t1 = tr.t2 = F2{0.f, 0.f};
            }
            tr.mat = d.matIndex;
            s->tris.push_back(tr);
        }
    }
    memcpy(s->lightIntensity, lightIntensityPrecomputed, 16);
    memcpy(s->lightDisplay, lightIntensityDisplay, 16);
    // RtAreaLight::createOptixCdf (rtcommon.h:501-531) with Triangle::ComputeArea (trianglemesh.cpp:13-19)
    {
        float sumArea = 0.f;
        s->lightCdf.resize(s->lightCount);
        for (int i = 0; i < s->lightCount; i++) {
            const Tri& t = s->tris[s->lightFirst + i];
            F3 ab = t.p1 - t.p0, ac = t.p2 - t.p0;
            F3 c = cross(ab, ac);
            float area = sqrtf(dot(c, c)) / 2.0f;
            sumArea += area;
            s->lightCdf[i] = sumArea;
        }
        for (int i = 0; i < s->lightCount; i++) s->lightCdf[i] /= sumArea;
        s->lightArea = sumArea;
    }
    s->bruteForce = bruteForce != 0;
    if (!s->bruteForce) build_bvh(*s);
    return s;
}

void orc_scene_destroy(void* h) { delete (Scene*)h; }
int32_t orc_num_prims(void* h) { return (int32_t)((Scene*)h)->tris.size(); }
float orc_light_area(void* h) { return ((Scene*)h)->lightArea; }
void orc_light_cdf(void* h, float* out) {
    Scene* s = (Scene*)h;
    memcpy(out, s->lightCdf.data(), s->lightCdf.size() * 4);
}
int32_t orc_num_threads(void) { return omp_get_max_threads(); }
// torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; callers that time the oracle set the count explicitly
void orc_set_threads(int32_t n) { if (n > 0) omp_set_num_threads(n); }

// RtScene::totalArea (rtcommon.h:759-768) and findBoundingSphereRadius (rtcommon.h:805-814);
// meshStart[numMeshes+1] gives the primitive range of each mesh (sequential f32 sums per mesh).
float orc_total_area(void* h, const int32_t* meshStart, int32_t numMeshes) {
    Scene* s = (Scene*)h;
    float sumArea = 0.f;
    for (int m = 0; m < numMeshes; m++) {
        float meshArea = 0.f;
        for (int i = meshStart[m]; i < meshStart[m + 1]; i++) {
            const Tri& t = s->tris[i];
            F3 c = cross(t.p1 - t.p0, t.p2 - t.p0);
            meshArea += sqrtf(dot(c, c)) / 2.0f;
        }
        sumArea += meshArea;
    }
    return sumArea;
}

// ------------------------------------------------------------------ taps -------------
void orc_uniforms(uint32_t seed, uint32_t subsequence, uint32_t n, float* out) {
    CurandState st;
    curand_init(seed, subsequence, 0, &st);
    for (uint32_t i = 0; i < n; i++) out[i] = curand_uniform(&st);
}
void orc_raw_u32(uint32_t seed, uint32_t subsequence, uint32_t n, uint32_t* out) {
    CurandState st;
    curand_init(seed, subsequence, 0, &st);
    for (uint32_t i = 0; i < n; i++) out[i] = curand(&st);
}

void orc_math(int op, const float* x, const float* y, uint32_t n, float* out) {
    for (uint32_t i = 0; i < n; i++) {
        switch (op) {
            case 0: out[i] = det_sinf(x[i]); break;
            case 1: out[i] = det_cosf(x[i]); break;
            case 2: out[i] = det_powf(x[i], y[i]); break;
            case 3: out[i] = det_asinf(x[i]); break;
            case 4: out[i] = sqrtf(x[i]); break;
            default: out[i] = 0.f;
        }
    }
}

// std::mt19937 restated (host sampler: common/rng.h:14,31-40 + sampler/independent.h:37-40).
// nextFloat = uniform_real_distribution<float>(0,1)(mt19937): MSVC's mapping is unpinned;
// defined here as float(u32) * 2^-32 with the "== 1" guard.
struct Mt19937 {
    uint32_t mt[624]; int idx;
    explicit Mt19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; i++) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
};
void orc_mt19937(uint32_t seed, uint32_t n, uint32_t* out) {
    Mt19937 g(seed);
    for (uint32_t i = 0; i < n; i++) out[i] = g.next();
}
// jitter stream: out[2*i], out[2*i+1] = nextVec2() of iteration i (x drawn first: defined, §A.1)
void orc_jitter_stream(uint32_t rngOffset, uint32_t numIterations, float* out) {
    Mt19937 g(rngOffset);
    for (uint32_t i = 0; i < 2 * numIterations; i++) {
        float u = (float)g.next() * 2.3283064365386963e-10f;
        if (u >= 1.0f) u = nextafterf(1.0f, 0.0f);
        out[i] = u;
    }
}

// Progressive schedule (rtcomphoton.h:1033-1063, evaluated after numIterations++).
// state = {photonRadius, clampingValue, pdfMc, vslRadius, vslInvPiRadius2}
void orc_progressive_update(int32_t numIterations, float alpha, float clampingStart, uint32_t numVplLightPaths,
                            uint32_t numLightPaths, int32_t forceVsl, float* state) {
    float ratio = ((float)numIterations + alpha) / (float)(numIterations + 1);
    state[0] *= std::sqrt(ratio);
    state[1] = (float)((double)clampingStart * std::pow((double)numIterations, (double)alpha));
    state[2] = static_cast<float>(numVplLightPaths) / static_cast<float>(numLightPaths) * 0.318309886183790671537767526745028724068919291480912897495f /
               (state[0] * state[0]);
    if (forceVsl) {
        state[3] *= std::sqrt(ratio);
        if (state[3] <= 0.008f) state[3] = std::max(state[3], 0.008f);
        state[4] = 0.318309886183790671537767526745028724068919291480912897495f / (state[3] * state[3]);
    }
}

// ------------------------------------------------------------------ stages -----------
void orc_light_trace(void* h, const EvplpParams* P, uint32_t rngSeed, uint32_t firstPath, uint32_t numPaths,
                     EvplpRecord* out) {
    Scene* s = (Scene*)h;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)numPaths; i++) {
        tracePhotons(*s, out, firstPath + (uint32_t)i, (uint32_t)i, P->numPhotonsPerLightPath, rngSeed);
    }
}

static void unpack_gpixel(const float* planes, const int32_t* prims, int64_t n, int64_t i, GPixel& g) {
    const float* p = planes + i * 4;
    g.position = mk3(p[0], p[1], p[2]); g.w = p[3];
    p = planes + (n + i) * 4; g.normal = mk3(p[0], p[1], p[2]);
    p = planes + (2 * n + i) * 4; g.lambert = mk3(p[0], p[1], p[2]);
    p = planes + (3 * n + i) * 4; g.phong = mk3(p[0], p[1], p[2]); g.exponent = p[3];
    g.prim = prims ? prims[i] : 0;
}

void orc_gbuffer(void* h, const EvplpParams* P, int32_t W, int32_t H, float* planes, int32_t* primIds) {
    Scene* s = (Scene*)h;
    const int64_t n = (int64_t)W * H;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
        int x = (int)(i % W), y = (int)(i / W);
        GPixel g = gbuffer_pixel(*s, *P, W, H, x, y);
        float* p = planes + i * 4; p[0] = g.position.x; p[1] = g.position.y; p[2] = g.position.z; p[3] = g.w;
        p = planes + (n + i) * 4; p[0] = g.normal.x; p[1] = g.normal.y; p[2] = g.normal.z; p[3] = 0.f;
        p = planes + (2 * n + i) * 4; p[0] = g.lambert.x; p[1] = g.lambert.y; p[2] = g.lambert.z; p[3] = 0.f;
        p = planes + (3 * n + i) * 4; p[0] = g.phong.x; p[1] = g.phong.y; p[2] = g.phong.z; p[3] = g.exponent;
        primIds[i] = g.prim;
    }
}

// One gather pass over `tile` (NULL = all). outRGB[W*H*3] receives the per-iteration value
// result / numVplLightPaths (0 outside the tile); counters[2] += {pairs, shadowRays}.
void orc_vpl_gather(void* h, const EvplpParams* P, int32_t W, int32_t H, const float* planes, const int32_t* primIds,
                    const EvplpRecord* records, int32_t mode, const EvplpTile* tile, float* outRGB, uint64_t* counters) {
    Scene* s = (Scene*)h;
    const int64_t n = (int64_t)W * H;
    EvplpTile t = tile ? *tile : EvplpTile{0, 0, W, H};
    uint64_t pairs = 0, rays = 0;
    memset(outRGB, 0, sizeof(float) * 3 * n);
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : pairs, rays)
    for (int64_t i = 0; i < n; i++) {
        int x = (int)(i % W), y = (int)(i / W);
        if (x < t.x0 || x >= t.x1 || y < t.y0 || y >= t.y1) continue;
        GPixel g; unpack_gpixel(planes, primIds, n, i, g);
        GatherCounters c;
        F3 r;
        if (mode == EVPLP_GATHER_VPL) r = splatColor(*s, *P, g, records, &c);
        else if (mode == EVPLP_GATHER_VSL) r = splatSplotch(*s, *P, g, records, (unsigned)i, &c);
        else r = splatColorLvc(*s, *P, g, records, (unsigned)i, &c);
        outRGB[i * 3] = r.x; outRGB[i * 3 + 1] = r.y; outRGB[i * 3 + 2] = r.z;
        pairs += c.pairs; rays += c.shadowRays;
    }
    if (counters) { counters[0] += pairs; counters[1] += rays; }
}

// RtPt2 path tracer (pathtracing.cu): outRGB[W*H*3] = one path per pixel of the tile; counters[0] += rays traced
void orc_path_trace(void* h, const EvplpParams* P, int32_t W, int32_t H, const float* planes, const int32_t* primIds, uint32_t maxBounces,
                    const EvplpTile* tile, float* outRGB, uint64_t* counters) {
    Scene* s = (Scene*)h;
    const int64_t n = (int64_t)W * H;
    EvplpTile t = tile ? *tile : EvplpTile{0, 0, W, H};
    uint64_t rays = 0;
    memset(outRGB, 0, sizeof(float) * 3 * n);
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : rays)
    for (int64_t i = 0; i < n; i++) {
        int x = (int)(i % W), y = (int)(i / W);
        if (x < t.x0 || x >= t.x1 || y < t.y0 || y >= t.y1) continue;
        GPixel g; unpack_gpixel(planes, primIds, n, i, g);
        uint64_t r = 0;
        F3 c = pt::splatColor(*s, *P, g, (unsigned)i, maxBounces, &r);
        outRGB[i * 3] = c.x; outRGB[i * 3 + 1] = c.y; outRGB[i * 3 + 2] = c.z;
        rays += r;
    }
    if (counters) counters[0] += rays;
}

// accum[W*H*3] (int64, Q31.32) += fixed(value)
void orc_accumulate_fixed(const float* rgb, int64_t count, int64_t* accum) {
    for (int64_t i = 0; i < count; i++) accum[i] += to_fixed(rgb[i]);
}

// Conservative pixel rectangle of the sphere (center p, radius r) for the camera of P.
static void splat_rect(const EvplpParams* P, int32_t W, int32_t H, const float* p, float r, int* x0, int* y0, int* x1, int* y1) {
    double c[3] = {(double)p[0] - P->cameraPosition[0], (double)p[1] - P->cameraPosition[1], (double)p[2] - P->cameraPosition[2]};
    double z = c[0] * P->camForward[0] + c[1] * P->camForward[1] + c[2] * P->camForward[2];
    double xv = c[0] * P->camRight[0] + c[1] * P->camRight[1] + c[2] * P->camRight[2];
    double yv = c[0] * P->camUp[0] + c[1] * P->camUp[1] + c[2] * P->camUp[2];
    double rr = (double)r * 1.001 + 1e-6;
    // G-buffer points lie at depth >= nearDist along the forward axis (primary rays start at tmin = nearDist)
    const double zNear = (double)P->nearDist * 0.999;
    *x0 = 0; *y0 = 0; *x1 = 0; *y1 = 0;
    if (z + rr < zNear) return;  // sphere entirely behind the near plane: no texel can be inside it
    double zn = std::max(z - rr, zNear), zf = z + rr;
    auto range = [&](double v, double tanHalf, double jit, int N, int* lo, int* hi) {
        double a = (v - rr), b = (v + rr);
        double lo_ndc = std::min(a / zn, a / zf) / tanHalf + jit;
        double hi_ndc = std::max(b / zn, b / zf) / tanHalf + jit;
        double plo = (lo_ndc + 1.0) * 0.5 * N - 0.5, phi = (hi_ndc + 1.0) * 0.5 * N - 0.5;
        plo = std::floor(plo) - 1.0; phi = std::ceil(phi) + 2.0;
        *lo = (int)std::max(0.0, std::min((double)N, plo));
        *hi = (int)std::max(0.0, std::min((double)N, phi));
    };
    range(xv, P->tanHalfFovX, P->jitter[0], W, x0, x1);
    range(yv, P->tanHalfFovY, P->jitter[1], H, y0, y1);
}

// Photon splat of records [firstRecord, firstRecord+numRecords) of `records` (index 0 = first record
// of the window) into accum (Q31.32).  counters[2] += {usable photons, fragments}.
// Coverage definition: photon k contributes once to pixel x iff IsUsablePhoton(k) and
// |p_k - pos(x)|^2 <= r^2 using x's own G-buffer texel; the screen rectangle below is only a conservative
// cull.  bruteForce != 0 visits every pixel for every photon (validates the cull).
void orc_photon_splat(const EvplpParams* P, int32_t W, int32_t H, const float* planes, const int32_t* primIds,
                      const EvplpRecord* records, uint64_t firstRecord, uint64_t numRecords, const EvplpTile* tile,
                      int64_t* accum, uint64_t* counters, int32_t bruteForce) {
    const int64_t n = (int64_t)W * H;
    EvplpTile t = tile ? *tile : EvplpTile{0, 0, W, H};
    uint64_t frags = 0, usable = 0;
    const float r2 = P->radius * P->radius;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : frags, usable)
    for (int64_t kk = 0; kk < (int64_t)numRecords; kk++) {
        const uint64_t k = firstRecord + (uint64_t)kk;
        const EvplpRecord& ph = records[k];
        if (!(ph.flags & EVPLP_FLAG_USABLE_PHOTON)) continue;
        usable++;
        int x0 = 0, y0 = 0, x1 = W, y1 = H;
        if (!bruteForce) splat_rect(P, W, H, ph.position, P->radius, &x0, &y0, &x1, &y1);
        x0 = std::max(x0, t.x0); y0 = std::max(y0, t.y0); x1 = std::min(x1, t.x1); y1 = std::min(y1, t.y1);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                int64_t i = (int64_t)y * W + x;
                if (primIds[i] < 0) continue;
                const float* gp = planes + i * 4;
                F3 d = getv(ph.position) - mk3(gp[0], gp[1], gp[2]);
                if (dot(d, d) > r2) continue;
                GPixel g; unpack_gpixel(planes, primIds, n, i, g);
                F3 color;
                if (!splatFragment(*P, g, records, k, &color)) continue;
                frags++;
                int64_t q[3] = {to_fixed(color.x), to_fixed(color.y), to_fixed(color.z)};
                for (int c = 0; c < 3; c++) {
                    if (q[c] == 0) continue;
#pragma omp atomic
                    accum[i * 3 + c] += q[c];
                }
            }
    }
    if (counters) { counters[0] += usable; counters[1] += frags; }
}

// light layer: count[i] += 1 where the nearest surface is the light mesh (rtcomphoton.h:839-855)
// runLightProgram (rtcomphoton.h:839-855, 985-995): the light mesh through the UN-jittered camera ("we don't jitter light
// source"), nearest surface wins; the mask is written (1 / 0), not accumulated.
void orc_light_pass(void* h, const EvplpParams* params, int32_t W, int32_t H, uint32_t* light) {
    Scene* s = (Scene*)h;
    EvplpParams P = *params;
    P.jitter[0] = 0.f; P.jitter[1] = 0.f;
#pragma omp parallel for schedule(dynamic, 8)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const F3 dir = primary_dir(P, W, H, x, y);
            const Hit hit = trace_closest(*s, getv(P.cameraPosition), dir, P.nearDist, P.farDist);
            light[(size_t)y * W + x] = (hit.prim >= s->lightFirst && hit.prim < s->lightFirst + s->lightCount) ? 1u : 0u;
        }
}

// final.frag:19-35 on the fixed-point layers; rows bottom-up (glReadPixels order).
void orc_resolve(void* h, int32_t W, int32_t H, const int64_t* vpl, const int64_t* photon, const uint32_t* light,
                 float vplScale, float photonScale, float lightScale, int32_t doGamma, float* outRGB) {
    Scene* s = (Scene*)h;
    for (int64_t i = 0; i < (int64_t)W * H; i++) {
        float res[3];
        float lightX = (light[i] ? s->lightDisplay[0] : 0.f) * lightScale;
        float stepv = (0.0f >= lightX) ? 1.0f : 0.0f;  // step(edge = lightColor.x, x = 0)
        for (int c = 0; c < 3; c++) {
            float vplColor = (float)((double)vpl[i * 3 + c] * (1.0 / 4294967296.0)) * vplScale;
            float pmColor = (float)((double)photon[i * 3 + c] * (1.0 / 4294967296.0)) * photonScale;
            float lightColor = (light[i] ? s->lightDisplay[c] : 0.f) * lightScale;
            float sum = stepv * (vplColor + pmColor) + lightColor;
            res[c] = doGamma ? powf(sum, 1.0f / 2.2f) : sum;
        }
        outRGB[i * 3] = res[0]; outRGB[i * 3 + 1] = res[1]; outRGB[i * 3 + 2] = res[2];
    }
}

// BRDF library taps, item by item the same calls as oracle/ref_device.cu:k_brdf makes into the reference's own
// rtmaterial.cuh / rtmath.cuh / lighttracing.cu functions: in = 16 floats per item, out = 8 floats per item.
void orc_brdf(int32_t op, const float* in16, uint32_t n, float* out8) {
    for (uint32_t i = 0; i < n; i++) {
        const float* q = in16 + (size_t)i * 16;
        float* o = out8 + (size_t)i * 8;
        const F3 a = mk3(q[0], q[1], q[2]), b = mk3(q[3], q[4], q[5]), c = mk3(q[6], q[7], q[8]), refl = mk3(q[9], q[10], q[11]);
        const float e = q[12];
        for (int k = 0; k < 8; k++) o[k] = 0.f;
        CurandState st;
        curand_init((unsigned)q[13], (unsigned)q[14], 0, &st);
        F3 dir = mk3(0.f, 0.f, 0.f);
        float pdf = 0.f;
        switch (op) {
            case 0: { F3 r = LambertSample(&dir, &pdf, a, b, refl, &st); setv(o, dir); o[3] = pdf; setv(o + 4, r); o[7] = curand_uniform(&st); break; }
            case 1: { F3 r = PhongSample(&dir, &pdf, a, b, refl, e, &st); setv(o, dir); o[3] = pdf; setv(o + 4, r); o[7] = curand_uniform(&st); break; }
            case 2: o[0] = LambertPdfA(a, b, c); o[1] = LambertPdfW(a, c); o[2] = pt::GeometryTerm(a, b, c); break;
            case 3: o[0] = PhongPdfA(a, b, c, refl, mk3(q[15], q[15], q[15]), e); o[1] = PhongPdfW(a, c, refl, mk3(q[15], q[15], q[15]), e); break;
            case 4: { o[0] = PhongEvalF(a, b, c, e); F3 r = pt::PhongEval(a, b, c, refl, e); setv(o + 1, r); break; }
            case 5: { float be, ga; SquareToBarycentric(&be, &ga, q[0], q[1]); o[0] = be; o[1] = ga; F3 sa = SquareToSolidAngle(q[0], q[1], q[2]); setv(o + 2, sa); o[5] = russianProb(b); break; }
        }
    }
}

void orc_trace_rays(void* h, const float* rays, uint64_t numRays, int32_t anyHit, int32_t* outPrim, float* outT) {
    Scene* s = (Scene*)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)numRays; i++) {
        const float* r = rays + i * 8;
        F3 o = mk3(r[0], r[1], r[2]), d = mk3(r[3], r[4], r[5]);
        if (anyHit) {
            outPrim[i] = trace_any(*s, o, d, r[6], r[7]) ? 1 : 0;
            if (outT) outT[i] = 0.f;
        } else {
            Hit hit = trace_closest(*s, o, d, r[6], r[7]);
            outPrim[i] = hit.prim;
            if (outT) outT[i] = hit.prim >= 0 ? hit.t : 0.f;
        }
    }
}

// ------------------------------------------------------------------ LBVH twin --------
// CPU twin of the product's Morton-code LBVH (Karras 2012), for bit-exact comparison of
// Morton codes, sorted order and binary topology (BASELINE.json north_star).
static inline uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

void orc_lbvh(void* h, uint64_t* mortonCodes, uint32_t* sortedPrimIds, int32_t* left, int32_t* right, int32_t* parent,
              float* nodeBounds, float* sceneMinMax) {
    Scene* s = (Scene*)h;
    const int n = (int)s->tris.size();
    std::vector<float> lo(n * 3), hi(n * 3);
    float smin[3] = {INFINITY, INFINITY, INFINITY}, smax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++) {
        const Tri& t = s->tris[i];
        const float p[3][3] = {{t.p0.x, t.p0.y, t.p0.z}, {t.p1.x, t.p1.y, t.p1.z}, {t.p2.x, t.p2.y, t.p2.z}};
        for (int a = 0; a < 3; a++) {
            lo[i * 3 + a] = fminf(fminf(p[0][a], p[1][a]), p[2][a]);
            hi[i * 3 + a] = fmaxf(fmaxf(p[0][a], p[1][a]), p[2][a]);
            smin[a] = fminf(smin[a], lo[i * 3 + a]);
            smax[a] = fmaxf(smax[a], hi[i * 3 + a]);
        }
    }
    if (sceneMinMax) { memcpy(sceneMinMax, smin, 12); memcpy(sceneMinMax + 3, smax, 12); }
    std::vector<uint64_t> code(n);
    for (int i = 0; i < n; i++) {
        uint32_t g[3];
        for (int a = 0; a < 3; a++) {
            float c = (lo[i * 3 + a] + hi[i * 3 + a]) * 0.5f;
            float ext = smax[a] - smin[a];
            float q = ext > 0.0f ? (c - smin[a]) / ext : 0.0f;
            g[a] = (uint32_t)fminf(fmaxf(q * 2097152.0f, 0.0f), 2097151.0f);
        }
        code[i] = (expand21(g[0]) << 2) | (expand21(g[1]) << 1) | expand21(g[2]);
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
    std::vector<uint64_t> sc(n);
    for (int i = 0; i < n; i++) sc[i] = code[order[i]];
    if (mortonCodes) memcpy(mortonCodes, sc.data(), sizeof(uint64_t) * n);
    if (sortedPrimIds) memcpy(sortedPrimIds, order.data(), sizeof(uint32_t) * n);
    if (n < 2) return;
    auto delta = [&](int i, int j) -> int {
        if (j < 0 || j >= n) return -1;
        uint64_t a = sc[i], b = sc[j];
        if (a == b) return 64 + __builtin_clz((uint32_t)i ^ (uint32_t)j);
        return __builtin_clzll(a ^ b);
    };
    std::vector<int> L(n - 1), R(n - 1), par(n - 1, -1), leafPar(n, -1), first(n - 1), last(n - 1);
    for (int i = 0; i < n - 1; i++) {
        int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
        int dmin = delta(i, i - d);
        int lmax = 2;
        while (delta(i, i + lmax * d) > dmin) lmax *= 2;
        int l = 0;
        for (int t = lmax / 2; t >= 1; t /= 2)
            if (delta(i, i + (l + t) * d) > dmin) l += t;
        int j = i + l * d;
        int dnode = delta(i, j);
        int sp = 0, t = l;
        do {
            t = (t + 1) / 2;
            if (delta(i, i + (sp + t) * d) > dnode) sp += t;
        } while (t > 1);
        int gamma = i + sp * d + std::min(d, 0);
        int lo_ = std::min(i, j), hi_ = std::max(i, j);
        first[i] = lo_; last[i] = hi_;
        L[i] = (lo_ == gamma) ? ~gamma : gamma;
        R[i] = (hi_ == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
    }
    for (int i = 0; i < n - 1; i++) {
        if (L[i] >= 0) par[L[i]] = i; else leafPar[~L[i]] = i;
        if (R[i] >= 0) par[R[i]] = i; else leafPar[~R[i]] = i;
    }
    if (left) memcpy(left, L.data(), 4 * (n - 1));
    if (right) memcpy(right, R.data(), 4 * (n - 1));
    if (parent) memcpy(parent, par.data(), 4 * (n - 1));
    if (nodeBounds) {
        // bounds of node i = union of the primitive boxes in its range (min/max are exact and
        // order-independent, so this equals any bottom-up refit)
        for (int i = 0; i < n - 1; i++) {
            float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
            for (int k = first[i]; k <= last[i]; k++) {
                int p = order[k];
                for (int a = 0; a < 3; a++) { b[a] = fminf(b[a], lo[p * 3 + a]); b[3 + a] = fmaxf(b[3 + a], hi[p * 3 + a]); }
            }
            memcpy(nodeBounds + (size_t)i * 6, b, 24);
        }
    }
}

}  // extern "C"
