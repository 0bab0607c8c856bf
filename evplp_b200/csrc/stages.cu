// stages.cu -- the per-iteration kernels of the EVPLP hot path (sm_100a).
//
//   gbuffer_kernel       replaces runDeferredProgram + deferred.{vert,geom,frag}   (rtcomphoton.h:710-754)
//   light_trace_kernel   replaces tracePhotons + rtMaterialClosestHit              (lighttracing.cu:113-250)
//   gather_vpl_kernel    replaces splatColor + vplSplat + rtMaterialAnyHit         (lighttracing.cu:184-188,275-379)
//   gather_vsl_kernel    replaces splatSplotch + vslSplat + sample*                (lighttracing.cu:382-722)
//   gather_lvc_kernel    replaces the LVC splatColor                               (lvclighttracing.cu:348-387)
//   splat_kernel         replaces runPhotonSplat + photonsplatinstanced.*          (rtcomphoton.h:789-837)
//   light_pass_kernel    replaces runLightProgram + light.frag                     (rtcomphoton.h:839-855)
//   resolve_kernel       replaces runFinalProgram + final.frag                     (rtcomphoton.h:756-787)
//
// Compiled with -fmad=false: all shading arithmetic rounds exactly as written (shading.h).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <curand_kernel.h>
#include "context.h"

namespace evplp {

// ------------------------------------------------------------------ helpers -------------
__device__ __forceinline__ Vertex load_vertex(const float4* r) {
    Vertex v;
    float4 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], f = r[5];
    v.pos = v3(a.x, a.y, a.z);
    v.normal = v3(b.x, b.y, b.z); v.pSel = b.w;
    v.flux = v3(c.x, c.y, c.z);
    v.fluxDir = v3(d.x, d.y, d.z);
    v.kd = v3(e.x, e.y, e.z);
    v.ks = v3(f.x, f.y, f.z); v.exponent = f.w;
    return v;
}

__device__ __forceinline__ void store_record(EvplpRecord* rec, V3 pos, uint32_t flags, V3 n, float pSel, V3 flux, V3 fluxDir,
                                             V3 kd, V3 ks, float e) {
    float4* r = reinterpret_cast<float4*>(rec);
    r[0] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(flags));
    r[1] = make_float4(n.x, n.y, n.z, pSel);
    r[2] = make_float4(flux.x, flux.y, flux.z, 0.f);
    r[3] = make_float4(fluxDir.x, fluxDir.y, fluxDir.z, 0.f);
    r[4] = make_float4(kd.x, kd.y, kd.z, 0.f);
    r[5] = make_float4(ks.x, ks.y, ks.z, e);
}

__device__ __forceinline__ void zero_record(EvplpRecord* rec) {
    float4* r = reinterpret_cast<float4*>(rec);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 6; k++) r[k] = z;
}

__device__ __forceinline__ Surface load_surface(const float4* __restrict__ gbuf, size_t n, size_t i, float* w) {
    Surface s;
    float4 a = gbuf[i], b = gbuf[n + i], c = gbuf[2 * n + i], d = gbuf[3 * n + i];
    s.pos = v3(a.x, a.y, a.z); *w = a.w;
    s.normal = v3(b.x, b.y, b.z);
    s.kd = v3(c.x, c.y, c.z);
    s.ks = v3(d.x, d.y, d.z); s.exponent = d.w;
    return s;
}

struct CamParams {
    V3 pos, fwd, right, up;
    float tanX, tanY, jx, jy, nearD, farD;
};

static CamParams cam_of(const EvplpParams& P) {
    CamParams c;
    c.pos = v3p(P.cameraPosition); c.fwd = v3p(P.camForward); c.right = v3p(P.camRight); c.up = v3p(P.camUp);
    c.tanX = P.tanHalfFovX; c.tanY = P.tanHalfFovY; c.jx = P.jitter[0]; c.jy = P.jitter[1];
    c.nearD = P.nearDist; c.farD = P.farDist;
    return c;
}

// ------------------------------------------------------------------ G-buffer ------------
__global__ void __launch_bounds__(128) gbuffer_kernel(DevScene sc, CamParams cam, int W, int H, float4* __restrict__ gbuf,
                                                      int32_t* __restrict__ gprim, DevStats* stats) {
    // 8x4 pixel tiles per warp keep primary rays coherent
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= W || y >= H) return;
    const size_t n = (size_t)W * H, i = (size_t)y * W + x;
    float cx = det_div((float)x + 0.5f, (float)W) * 2.0f - 1.0f;
    float cy = det_div((float)y + 0.5f, (float)H) * 2.0f - 1.0f;
    float nx = (cx - cam.jx) * cam.tanX;
    float ny = (cy - cam.jy) * cam.tanY;
    V3 dir = cam.fwd + cam.right * nx + cam.up * ny;
    int ovf = 0;
    RayHit hit = trace_closest<false>(sc, cam.pos, dir, cam.nearD, cam.farD, &ovf);
    if (ovf) stats->stackOverflow = 1;
    float4 o0 = make_float4(0.f, 0.f, 0.f, 1.f), o1 = make_float4(0.f, 0.f, 0.f, 0.f), o2 = o1, o3 = o1;
    if (hit.prim >= 0) {
        const float4 a = sc.triVerts[3 * (size_t)hit.prim], b = sc.triVerts[3 * (size_t)hit.prim + 1],
                     c = sc.triVerts[3 * (size_t)hit.prim + 2];
        const V3 p0 = ld3(a), p1 = ld3(b), p2 = ld3(c);
        const float w0 = 1.0f - hit.beta - hit.gamma;
        V3 pos = p0 * w0 + p1 * hit.beta + p2 * hit.gamma;        // deferred.geom:23 (interpolated world position)
        V3 nrm = normalize(cross(p1 - p0, p2 - p0));              // deferred.geom:16-18 (flat, not face-forwarded)
        const float2 t0 = sc.triUV[3 * (size_t)hit.prim], t1 = sc.triUV[3 * (size_t)hit.prim + 1],
                     t2 = sc.triUV[3 * (size_t)hit.prim + 2];
        float u = t0.x * w0 + t1.x * hit.beta + t2.x * hit.gamma;
        float v = t0.y * w0 + t1.y * hit.beta + t2.y * hit.gamma;
        const DevMaterial& m = sc.mats[hit.mat];
        Texel4 kd = tex_fetch(m.lambert, sc.texPool, u, v);
        Texel4 ks = tex_fetch(m.phong, sc.texPool, u, v);
        Texel4 ex = tex_fetch(m.exponent, sc.texPool, u, v);
        o0 = make_float4(pos.x, pos.y, pos.z, 1.f);
        o1 = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
        o2 = make_float4(kd.x, kd.y, kd.z, 0.f);
        o3 = make_float4(ks.x, ks.y, ks.z, ex.x);
    }
    gbuf[i] = o0; gbuf[n + i] = o1; gbuf[2 * n + i] = o2; gbuf[3 * n + i] = o3;
    gprim[i] = hit.prim;
}

// LightSample -- rtlightsource.cuh:24-80: triangle by lower-bound search in the area CDF, uniform point on it;
// returns areaLightIntensity.rgb * area (pdf = 1 / area).  Three uniforms, drawn in order.
__device__ __forceinline__ V3 light_sample(const DevScene& sc, V3* position, V3* normal, Xorwow& rng) {
    float randNum = xorwow_uniform(rng);
    unsigned count = (unsigned)sc.lightCount, first = 0;
    while (count > 0) {
        unsigned step = count / 2;
        unsigned it = first + step;
        if (sc.lightCdf[it] < randNum) { first = it + 1; count -= step + 1; } else { count = step; }
    }
    const size_t prim = (size_t)sc.lightFirst + first;
    const V3 pos1 = ld3(sc.triVerts[3 * prim]), pos2 = ld3(sc.triVerts[3 * prim + 1]), pos3 = ld3(sc.triVerts[3 * prim + 2]);
    float bx = xorwow_uniform(rng);
    float by = xorwow_uniform(rng);
    float beta, gamma;
    square_to_barycentric(&beta, &gamma, bx, by);
    *position = pos1 * beta + pos2 * gamma + pos3 * (1.0f - gamma - beta);
    *normal = normalize(cross(pos2 - pos1, pos3 - pos1));
    return v3(sc.lightIntensity[0], sc.lightIntensity[1], sc.lightIntensity[2]) * sc.lightArea;
}

// ------------------------------------------------------------------ light tracing -------
__global__ void __launch_bounds__(128, 8) light_trace_kernel(DevScene sc, const uint32_t* __restrict__ skip,
                                                          EvplpRecord* __restrict__ records, uint32_t firstPath,
                                                          uint32_t numPaths, uint32_t B1, DevStats* stats) {
    __shared__ __align__(16) uint32_t sm[kSkipTableWords];   // `skip` = the launch's skip matrix as 4-bit tables (xorwow.h)
    for (int k = threadIdx.x; k < kSkipTableWords / 4; k += blockDim.x)
        reinterpret_cast<uint4*>(sm)[k] = __ldg(reinterpret_cast<const uint4*>(skip) + k);
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numPaths) return;
    EvplpRecord* rec = records + (size_t)i * B1;

    // curand_init(launchId, rngSeed, 0) -- lighttracing.cu:203
    Xorwow rng = xorwow_seed_skip(firstPath + i, sm);

    V3 position, normal;
    V3 flux = light_sample(sc, &position, &normal, rng);
    V3 direction;
    float pdfW;
    V3 att = phong_sample(&direction, &pdfW, normal, normal, v3s(1.0f), sc.lightIntensity[3], rng);
    store_record(rec, position, EVPLP_FLAG_USABLE_VPL, normal, 0.0f, flux, normal, v3s(0.0f), v3s(1.0f), sc.lightIntensity[3]);

    V3 pflux = flux * att;
    V3 nextPosition = position, nextDirection = direction;
    uint32_t written = 1;
    unsigned long long rays = 0;
    int ovf = 0;
    for (uint32_t b = 1; b < B1; b++) {
        const V3 rayOrigin = nextPosition, rayDirection = nextDirection;
        const uint32_t flag = (b != B1 - 1) ? (EVPLP_FLAG_USABLE_VPL | EVPLP_FLAG_USABLE_PHOTON) : EVPLP_FLAG_USABLE_PHOTON;
        rays++;
        RayHit hit = trace_closest<true>(sc, rayOrigin, rayDirection, 0.0001f, 1e27f, &ovf);
        if (hit.prim < 0) break;  // no miss program: the path ends
        // rtMaterialClosestHit -- lighttracing.cu:113-182
        const DevMaterial& mat = sc.mats[hit.mat];
        V3 geometryNormal = normalize(hit.n);
        V3 worldGeometryNormal = normalize(geometryNormal);
        V3 ffNormal = faceforward(worldGeometryNormal, -rayDirection, worldGeometryNormal);
        V3 hitPosition = rayOrigin + hit.t * rayDirection;
        if (dot(geometryNormal, rayDirection) > 0.f || mat.lightIntensity[0] > 0.01f) break;
        const float2 t0 = sc.triUV[3 * (size_t)hit.prim], t1 = sc.triUV[3 * (size_t)hit.prim + 1],
                     t2 = sc.triUV[3 * (size_t)hit.prim + 2];
        const float w0 = 1.0f - hit.beta - hit.gamma;
        const float u = t1.x * hit.beta + t2.x * hit.gamma + t0.x * w0;   // triangleintersect.cu:36
        const float v = t1.y * hit.beta + t2.y * hit.gamma + t0.y * w0;
        Texel4 tl = tex_fetch(mat.lambert, sc.texPool, u, v);
        Texel4 tp = tex_fetch(mat.phong, sc.texPool, u, v);
        Texel4 te = tex_fetch(mat.exponent, sc.texPool, u, v);
        V3 kd = v3(tl.x, tl.y, tl.z), ks = v3(tp.x, tp.y, tp.z);
        float phongExponent = te.x;
        float maxLambert = max_color(kd), maxPhong = max_color(ks);
        if (maxLambert + maxPhong <= 0.000001f) break;

        float pSelectLambert = det_div(maxLambert, maxPhong + maxLambert);
        float chooseMaterial = det_min(xorwow_uniform(rng), 0.999999f);
        const V3 recFlux = pflux;
        float russian = russian_prob(pflux);
        pflux /= russian;
        bool done = (xorwow_uniform(rng) >= russian);
        uint32_t recFlags = flag;
        if (!done) {
            if (chooseMaterial < pSelectLambert) {
                pflux *= lambert_sample(&direction, &pdfW, ffNormal, kd, rng) / pSelectLambert;
                recFlags = flag | EVPLP_FLAG_LAMBERT_ONLY;
            } else {
                pflux *= phong_sample(&direction, &pdfW, -rayDirection, geometryNormal, ks, phongExponent, rng) /
                         (1.0f - pSelectLambert);
                recFlags = flag | EVPLP_FLAG_PHONG_ONLY;
            }
            nextPosition = hitPosition;
            nextDirection = direction;
        }
        store_record(rec + b, hitPosition, recFlags, ffNormal, pSelectLambert, recFlux, -rayDirection, kd, ks, phongExponent);
        written = b + 1;
        if (done) break;
    }
    // Slots after the end of the path: the reference clears only mFlags (lighttracing.cu:197-200);
    // the whole record is zeroed here so that the buffer is a deterministic function of the inputs.
    // NOTE: a path can leave a hole (bounce b rejected => loop ends), never a gap followed by data.
    for (uint32_t b = written; b < B1; b++) zero_record(rec + b);
    if (ovf) stats->stackOverflow = 1;
    if (rays) atomicAdd(&stats->closestRays, rays);  // uniform address: ptxas aggregates per warp
}

// ------------------------------------------------------------------ VPL gather ----------
template <int MINB, bool SHAFT>
__global__ void __launch_bounds__(GATHER_WARPS * 32, MINB)
gather_vpl_kernel(DevScene sc, GatherParams gp, const float4* __restrict__ gbuf, const EvplpRecord* __restrict__ records,
                  const uint32_t* __restrict__ vplList, const uint32_t* __restrict__ vplCount, long long* __restrict__ acc,
                  DevStats* stats, uint32_t* __restrict__ tileCounter, const uint32_t* __restrict__ tileOrder,
                  uint32_t* __restrict__ tileCost) {
    // Every warp stages its own VPL batches (no block-wide barrier: warps whose pixels are
    // culled by the cosine test run ahead instead of waiting for the slowest warp of the block).
    // per VPL: the 6 float4 of its record + 3 float4 of shading terms that depend on the VPL alone (VplPre)
    constexpr int VS = 9;
    __shared__ float4 batchAll[GATHER_WARPS][GATHER_BATCH * VS];
    __shared__ uint32_t stacks[GATHER_WARPS][BVH_STACK];
    __shared__ uint32_t cands[SHAFT ? GATHER_WARPS : 1][SHAFT_CAND];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* batch = batchAll[warp];
    const uint32_t stackBase = opaque((uint32_t)__cvta_generic_to_shared(stacks[warp]));
    const uint32_t candBase = opaque((uint32_t)__cvta_generic_to_shared(cands[SHAFT ? warp : 0]));
    const uint32_t vTotal = gp.tilePartition ? gp.ownedTiles * gp.numChunks : gp.vgx * gp.vgy * gp.vgz * GATHER_WARPS;
    unsigned rays = 0;
    unsigned shaftCnt[3] = {0u, 0u, 0u}, shaftSteps = 0u;
    int ovf = 0;
    for (;;) {
    // virtual warp v = ((bz * vgy + by) * vgx + bx) * GATHER_WARPS + vw of the block grid
    uint32_t v;
    if (gp.persistent) {
        if (lane == 0) v = atomicAdd(tileCounter, 1u);
        v = __shfl_sync(0xffffffffu, v, 0);
        if (v >= vTotal) break;
        if (tileOrder) v = __ldg(tileOrder + v);  // most expensive tiles of the previous launch first
    } else {
        v = ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * GATHER_WARPS + warp;
    }
    const long long itemStart = clock64();
    int x, y;
    uint32_t bz;
    bool inTile = true;
    if (gp.tilePartition) {   // tile t = tOffset + k * tStride of the handle's share, VPL range bz
        bz = v / gp.ownedTiles;
        const uint32_t t = gp.tOffset + (v % gp.ownedTiles) * gp.tStride;
        const int ty = (int)(t / (uint32_t)gp.pitchX), tx = (int)(t % (uint32_t)gp.pitchX);
        x = gp.x0 + tx * 8 + (lane & 7);
        y = gp.y0 + ty * 4 + (lane >> 3);
        inTile = tx < gp.tilesX;
    } else {
        const uint32_t vw = v % GATHER_WARPS, vb = v / GATHER_WARPS;
        const uint32_t bx = vb % gp.vgx, by = (vb / gp.vgx) % gp.vgy;
        bz = vb / (gp.vgx * gp.vgy);
        x = gp.x0 + bx * 16 + (vw & 1) * 8 + (lane & 7);
        y = gp.y0 + (by * gp.bandStride + gp.bandOffset) * 16 + (vw >> 1) * 4 + (lane >> 3);
    }
    const bool inside = inTile && x < gp.x1 && y < gp.y1;
    const size_t n = (size_t)gp.W * gp.H;
    const size_t i = inside ? (size_t)y * gp.W + x : 0;
    float gw;
    Surface sf = load_surface(gbuf, n, i, &gw);
    const bool valid = inside && gw != 0.0f;
    const V3 wi01 = normalize(gp.cameraPosition - sf.pos);
    // bounds of the warp's surface points: the far end of every (VPL -> tile) shaft
    V3 tileLo = valid ? sf.pos : v3s(INFINITY), tileHi = valid ? sf.pos : v3s(-INFINITY);
    if (SHAFT) {
        for (int o = 16; o > 0; o >>= 1) {
            tileLo = vmin(tileLo, v3(__shfl_xor_sync(0xffffffffu, tileLo.x, o), __shfl_xor_sync(0xffffffffu, tileLo.y, o), __shfl_xor_sync(0xffffffffu, tileLo.z, o)));
            tileHi = vmax(tileHi, v3(__shfl_xor_sync(0xffffffffu, tileHi.x, o), __shfl_xor_sync(0xffffffffu, tileHi.y, o), __shfl_xor_sync(0xffffffffu, tileHi.z, o)));
        }
    }

    const uint32_t total = *vplCount;
    const uint32_t per = (total + gp.numChunks - 1) / gp.numChunks;
    const uint32_t begin = min(total, bz * per);
    const uint32_t end = __any_sync(0xffffffffu, valid) ? min(total, begin + per) : begin;

    V3 result = v3s(0.0f);
    int streak = 0, skipLeft = 0;
    for (uint32_t base = begin; base < end; base += GATHER_BATCH) {
        const uint32_t nb = min((uint32_t)GATHER_BATCH, end - base);
        __syncwarp();
        for (uint32_t k = lane; k < nb * 6; k += 32) {
            const uint32_t r = vplList[base + k / 6];
            batch[(k / 6) * VS + k % 6] = __ldg(reinterpret_cast<const float4*>(records + r) + k % 6);
        }
        __syncwarp();
        if (lane < nb) {
            const VplPre pre = vpl_precompute(load_vertex(&batch[lane * VS]));
            batch[lane * VS + 6] = make_float4(pre.refl.x, pre.refl.y, pre.refl.z, 0.f);
            batch[lane * VS + 7] = make_float4(pre.reflN.x, pre.reflN.y, pre.reflN.z, 0.f);
            batch[lane * VS + 8] = make_float4(pre.kdPi.x, pre.kdPi.y, pre.kdPi.z, 0.f);
        }
        __syncwarp();
        for (uint32_t j = 0; j < nb; j++) {
            const float4 a = batch[j * VS], b = batch[j * VS + 1];
            const V3 vpos = v3(a.x, a.y, a.z), vn = v3(b.x, b.y, b.z);
            const V3 v12 = vpos - sf.pos;
            const float c1 = det_max(dot(sf.normal, v12), 0.0f);
            const float c2 = det_max(-dot(vn, v12), 0.0f);
            const float c1c2 = c1 * c2;
            const bool active = valid && !(c1c2 <= 0.000f);
            rays += active ? 1u : 0u;
            // Ray(vpl.pos, -v12, shadow, 0.0001, 1 - 0.0001) -- lighttracing.cu:292
            bool occluded;
            if (SHAFT) {
                // Overflowing shafts come in runs (a tile in clutter or across a depth edge overflows towards most VPLs):
                // after shaftStreak overflows in a row the next shaftSkip steps go to the packet traversal directly
                // instead of paying for a descent that is expected to be thrown away.
                const bool any = __any_sync(0xffffffffu, active);
                const Shaft sh = make_shaft(vpos, tileLo, tileHi);
                const unsigned before = shaftCnt[0];
                occluded = trace_any_warp_shaft(sc, active, vpos, -v12, (float)0.0001, (float)(1 - 0.0001), sh, stackBase, candBase, stacks[warp], gp.shaftCandMax, &ovf, shaftCnt,
                                                skipLeft > 0);
                if (any) {
                    shaftSteps++;
                    if (skipLeft > 0) skipLeft--;
                    else if (shaftCnt[0] != before) { if (++streak >= gp.shaftStreak) { skipLeft = gp.shaftSkip; streak = 0; } }
                    else streak = 0;
                }
            } else {
                occluded = trace_any_warp(sc, active, vpos, -v12, (float)0.0001, (float)(1 - 0.0001), stacks[warp], &ovf);
            }
            if (active && !occluded) {
                const Vertex vp = load_vertex(&batch[j * VS]);
                const float4 p0 = batch[j * VS + 6], p1 = batch[j * VS + 7], p2 = batch[j * VS + 8];
                VplPre pre;
                pre.refl = v3(p0.x, p0.y, p0.z); pre.reflN = v3(p1.x, p1.y, p1.z); pre.kdPi = v3(p2.x, p2.y, p2.z);
                result += vpl_shade_pre(sf, wi01, vp, pre, v12, c1c2, gp.misMode, gp.pdfMc, gp.clampingValue);
            }
        }
    }
    if (inside) {
        const V3 out = result * gp.invNumVpl;  // result / (float)numVplLightPaths (reciprocal multiply, lighttracing.cu:378)
        const long long q[3] = {to_fixed(out.x), to_fixed(out.y), to_fixed(out.z)};
        if (gp.numChunks == 1) {
            for (int c = 0; c < 3; c++) acc[i * 3 + c] = gp.doAccumulate ? acc[i * 3 + c] + q[c] : q[c];
        } else {
            // (the tile was cleared beforehand when doAccumulate == 0)
            for (int c = 0; c < 3; c++)
                if (q[c]) atomicAdd(reinterpret_cast<unsigned long long*>(acc + i * 3 + c), (unsigned long long)q[c]);
        }
    }
    if (!gp.persistent) break;
    if (tileCost && lane == 0) {
        const long long dt = (clock64() - itemStart) >> 8;
        tileCost[v] = dt > 0xffffffffll ? 0xffffffffu : (uint32_t)dt;
    }
    }  // next tile
    if (ovf) stats->stackOverflow = 1;
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (lane == 0 && rays) atomicAdd(&stats->shadowRays, (unsigned long long)rays);
    if (SHAFT && lane == 0 && shaftSteps) {
        atomicAdd(&stats->shaftSteps, (unsigned long long)shaftSteps); atomicAdd(&stats->shaftFallbacks, (unsigned long long)shaftCnt[0]);
        atomicAdd(&stats->shaftNodeVisits, (unsigned long long)shaftCnt[1]); atomicAdd(&stats->shaftCandLeaves, (unsigned long long)shaftCnt[2]);
    }
}

__global__ void clear_tile_kernel(long long* acc, int W, int x0, int y0, int x1, int y1) {
    int x = x0 + blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y;
    if (x >= x1 || y >= y1) return;
    size_t i = (size_t)y * W + x;
    acc[i * 3] = 0; acc[i * 3 + 1] = 0; acc[i * 3 + 2] = 0;
}

// ------------------------------------------------------------------ VSL gather ----------
__device__ __forceinline__ V3 square_to_solid_angle(float sampleX, float sampleY, float halfAngleMax) {  // :382-390
    const float phi = 2.0f * kPi * sampleX;
    const float z = 1.0f - sampleY * (1.0f - det_cosf(halfAngleMax));
    const float l = det_sqrtf(1.0f - z * z);
    float s, c;
    det_sincosf(phi, &s, &c);
    return v3(c * l, s * l, z);
}

struct VslConsts {
    float vslInvPiRadius2;
};

__device__ __forceinline__ float pdf_brdf1(const Surface& sf, V3 wi01, V3 w, float pSel) {
    return lambert_pdf_w(sf.normal, w) * pSel + phong_pdf_w(sf.normal, w, wi01, sf.ks, sf.exponent) * (1.0f - pSel);
}
// quirk kept: no (1 - pSel) on the Phong term and the PIXEL's pSel (lighttracing.cu:440-441)
__device__ __forceinline__ float pdf_brdf2(const Vertex& vp, V3 w, float pSel) {
    return lambert_pdf_w(vp.normal, w) * pSel + phong_pdf_w(vp.normal, w, vp.fluxDir, vp.ks, vp.exponent);
}

__device__ V3 vsl_sample_cone(float K, float* misWeight, const Surface& sf, V3 wi01, const Vertex& vp, float halfCone,
                              float solidAngle, float invSolidAngle, V3 nd12, Xorwow& rng) {  // :395-446
    float maxLambert = max_color(sf.kd), maxPhong = max_color(sf.ks);
    if (maxLambert + maxPhong <= 0.000001f) return v3s(0.0f);
    float pSel = det_div(maxLambert, maxPhong + maxLambert);
    (void)xorwow_uniform(rng);  // chooseMaterial (drawn, unused)
    float sx = xorwow_uniform(rng);
    float sy = xorwow_uniform(rng);
    V3 wi12 = normalize(square_to_solid_angle(sx, sy, halfCone));
    Onb onb = make_onb(nd12);
    wi12 = onb_inverse_transform(onb, wi12);
    wi12 = normalize(wi12);
    const float cos1cos2 = det_max(dot(sf.normal, wi12), 0.0f) * det_max(-dot(vp.normal, wi12), 0.0f);
    if (cos1cos2 <= 0.000000001f) return v3s(0.0f);
    V3 brdf2 = kInvPi * vp.kd + phong_eval_f(-wi12, vp.fluxDir, vp.normal, vp.exponent) * vp.ks;
    V3 brdf1 = kInvPi * sf.kd + phong_eval_f(wi01, wi12, sf.normal, sf.exponent) * sf.ks;
    float pdfCone = invSolidAngle;
    float p1 = pdf_brdf1(sf, wi01, wi12, pSel);
    float p2 = pdf_brdf2(vp, -wi12, pSel);
    *misWeight = det_div(pdfCone, p1 + p2 + pdfCone);
    return vp.flux * K * cos1cos2 * brdf1 * brdf2 * solidAngle;
}

__device__ V3 vsl_sample_brdf1(float K, float* misWeight, const Surface& sf, V3 wi01, const Vertex& vp, float cosHalfCone,
                               float invSolidAngle, V3 nd12, Xorwow& rng) {  // :448-521
    float maxLambert = max_color(sf.kd), maxPhong = max_color(sf.ks);
    if (maxLambert + maxPhong <= 0.000001f) return v3s(0.0f);
    float pSel = det_div(maxLambert, maxPhong + maxLambert);
    float chooseMaterial = det_min(xorwow_uniform(rng), 0.999999f);
    V3 wi12, brdf1;
    float pdfW;
    if (chooseMaterial < pSel) {
        brdf1 = lambert_sample(&wi12, &pdfW, sf.normal, sf.kd, rng) / pSel;
    } else {
        brdf1 = phong_sample(&wi12, &pdfW, wi01, sf.normal, sf.ks, sf.exponent, rng) / (1.0f - pSel);
    }
    if (dot(wi12, nd12) <= cosHalfCone) return v3s(0.0f);
    const float cos1 = det_max(dot(sf.normal, wi12), 0.0f);
    if (cos1 <= 0.000000001f) return v3s(0.0f);
    const float cos2 = det_max(-dot(vp.normal, wi12), 0.0f);
    V3 brdf2 = kInvPi * vp.kd + phong_eval_f(-wi12, vp.fluxDir, vp.normal, vp.exponent) * vp.ks;
    (void)xorwow_uniform(rng);  // second chooseMaterial (drawn, unused)
    float pdfCone = invSolidAngle;
    float p1 = pdf_brdf1(sf, wi01, wi12, pSel);
    float p2 = pdf_brdf2(vp, -wi12, pSel);
    *misWeight = det_div(p1, p1 + p2 + pdfCone);
    return vp.flux * K * cos2 * brdf1 * brdf2;
}

__device__ V3 vsl_sample_brdf2(float K, float* misWeight, const Surface& sf, V3 wi10, const Vertex& vp, float cosHalfCone,
                               float invSolidAngle, V3 nd12, Xorwow& rng) {  // :523-594
    V3 wi21, brdf2;
    float pdfW;
    {
        float maxLambert = max_color(vp.kd), maxPhong = max_color(vp.ks);
        if (maxLambert + maxPhong <= 0.000001f) return v3s(0.0f);
        float pSelV = det_div(maxLambert, maxPhong + maxLambert);
        float chooseMaterial = det_min(xorwow_uniform(rng), 0.999999f);
        if (chooseMaterial < pSelV) {
            brdf2 = lambert_sample(&wi21, &pdfW, vp.normal, vp.kd, rng) / pSelV;
        } else {
            brdf2 = phong_sample(&wi21, &pdfW, vp.fluxDir, vp.normal, vp.ks, vp.exponent, rng) / (1.0f - pSelV);
        }
    }
    if (-dot(wi21, nd12) <= cosHalfCone) return v3s(0.0f);
    V3 brdf1 = kInvPi * sf.kd + phong_eval_f(wi10, -wi21, sf.normal, sf.exponent) * sf.ks;
    const float cos2 = det_max(dot(vp.normal, wi21), 0.0f);
    if (cos2 <= 0.00000001f) return v3s(0.0f);
    const float cos1 = det_max(-dot(sf.normal, wi21), 0.0f);
    float maxLambert = max_color(sf.kd), maxPhong = max_color(sf.ks);
    if (maxLambert + maxPhong <= 0.000001f) return v3s(0.0f);
    float pSel = det_div(maxLambert, maxPhong + maxLambert);
    (void)xorwow_uniform(rng);  // chooseMaterial (drawn, unused)
    float pdfCone = invSolidAngle;
    float p1 = pdf_brdf1(sf, wi10, -wi21, pSel);
    float p2 = pdf_brdf2(vp, wi21, pSel);
    *misWeight = det_div(p2, p1 + p2 + pdfCone);
    return vp.flux * K * cos1 * brdf1 * brdf2;
}

// vslSplat after the shadow ray (lighttracing.cu:615-686)
__device__ V3 vsl_shade(const GatherParams& gp, const Surface& sf, V3 wi10, const Vertex& vp, V3 v12, Xorwow& rng) {
    float dist2 = dot(v12, v12);
    float dist = det_sqrtf(dist2);
    V3 nv12 = v12 / dist;
    const float cos1cos2 = det_max(dot(sf.normal, nv12), 0.0f) * det_max(-dot(vp.normal, nv12), 0.0f);
    if (cos1cos2 <= 0.000000001f) return v3s(0.0f);
    const float rdratio = det_div(gp.vslRadius, dist);
    const float halfCone = (rdratio >= 1.0f) ? det_div(kPi, 2.0f) : det_asinf(rdratio);
    const float cosHalfCone = det_cosf(halfCone);
    const float solidAngle = kPi * 2.0f * (1.0f - cosHalfCone);
    const float invSolidAngle = det_div(1.0f, solidAngle);
    V3 result = v3s(0.0f);
    const int numSamples = (int)(det_div(halfCone, kPi) * 2.0f * 100.0f) + 1;
    const float K = gp.vslInvPiRadius2;
    for (int s = 0; s < numSamples; s++) {
        float wc = 0.0f, w1 = 0.0f, w2 = 0.0f;
        V3 rc = vsl_sample_cone(K, &wc, sf, wi10, vp, halfCone, solidAngle, invSolidAngle, nv12, rng);
        V3 r1 = vsl_sample_brdf1(K, &w1, sf, wi10, vp, cosHalfCone, invSolidAngle, nv12, rng);
        V3 r2 = vsl_sample_brdf2(K, &w2, sf, wi10, vp, cosHalfCone, invSolidAngle, nv12, rng);
        result += wc * rc;
        result += w1 * r1;
        result += w2 * r2;
    }
    return result / (float)numSamples;
}

// splatSplotch (lighttracing.cu:689-722): one thread per pixel, the per-pixel XORWOW stream
// runs through ALL VSLs in order, so the VSL list cannot be split.
#ifndef EVPLP_VSL_MINB
#define EVPLP_VSL_MINB 2   // blocks / SM the register allocation aims at (tuning: -DEVPLP_VSL_MINB=3 / 4)
#endif
__global__ void __launch_bounds__(GATHER_WARPS * 32, EVPLP_VSL_MINB)
gather_vsl_kernel(DevScene sc, GatherParams gp, const uint32_t* __restrict__ skip, const float4* __restrict__ gbuf,
                  const EvplpRecord* __restrict__ records, const uint32_t* __restrict__ vplList,
                  const uint32_t* __restrict__ vplCount, long long* __restrict__ acc, DevStats* stats) {
    __shared__ uint32_t sm[kSkipMatrixWords];
    __shared__ uint32_t stacks[GATHER_WARPS][BVH_STACK];
    __shared__ uint32_t cands[GATHER_WARPS][SHAFT_CAND];
    for (int k = threadIdx.x; k < kSkipMatrixWords; k += blockDim.x) sm[k] = skip[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t stackBase = opaque((uint32_t)__cvta_generic_to_shared(stacks[warp]));
    const uint32_t candBase = opaque((uint32_t)__cvta_generic_to_shared(cands[warp]));
    const int x = gp.x0 + blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = gp.y0 + (blockIdx.y * gp.bandStride + gp.bandOffset) * 16 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = x < gp.x1 && y < gp.y1;
    const size_t n = (size_t)gp.W * gp.H;
    const size_t i = inside ? (size_t)y * gp.W + x : 0;
    float gw;
    Surface sf = load_surface(gbuf, n, i, &gw);
    const V3 wi10 = normalize(gp.cameraPosition - sf.pos);
    // far end of every (VSL -> tile) shaft: bounds of the warp's texel positions (background texels included: the
    // reference traces their shadow rays too, lighttracing.cu:694-695)
    V3 tileLo = inside ? sf.pos : v3s(INFINITY), tileHi = inside ? sf.pos : v3s(-INFINITY);
    for (int o = 16; o > 0; o >>= 1) {
        tileLo = vmin(tileLo, v3(__shfl_xor_sync(0xffffffffu, tileLo.x, o), __shfl_xor_sync(0xffffffffu, tileLo.y, o), __shfl_xor_sync(0xffffffffu, tileLo.z, o)));
        tileHi = vmax(tileHi, v3(__shfl_xor_sync(0xffffffffu, tileHi.x, o), __shfl_xor_sync(0xffffffffu, tileHi.y, o), __shfl_xor_sync(0xffffffffu, tileHi.z, o)));
    }
    unsigned shaftCnt[3] = {0u, 0u, 0u};
    Xorwow rng = xorwow_seed((uint32_t)i);  // curand_init(launchIndex.y * W + launchIndex.x, rngSeed, 0) -- :711
    xorwow_apply_matrix(rng, sm);
    const uint32_t total = *vplCount;
    V3 result = v3s(0.0f);
    unsigned rays = 0;
    int ovf = 0;
    for (uint32_t j = 0; j < total; j++) {
        const float4* r = reinterpret_cast<const float4*>(records + vplList[j]);
        const Vertex vp = load_vertex(r);
        const V3 v12 = vp.pos - sf.pos;
        rays += inside ? 1u : 0u;
        // shadow ray FIRST, before the cosine test (lighttracing.cu:609-614)
        bool occluded;
        if (gp.shaftMode) {
            const Shaft sh = make_shaft(vp.pos, tileLo, tileHi);
            occluded = trace_any_warp_shaft(sc, inside, vp.pos, -v12, (float)0.0001, (float)(1 - 0.0001), sh, stackBase, candBase,
                                            stacks[warp], gp.shaftCandMax, &ovf, shaftCnt);
        } else {
            occluded = trace_any_warp(sc, inside, vp.pos, -v12, (float)0.0001, (float)(1 - 0.0001), stacks[warp], &ovf);
        }
        if (inside && !occluded) result += vsl_shade(gp, sf, wi10, vp, v12, rng);
    }
    if (ovf) stats->stackOverflow = 1;
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if (lane == 0 && rays) atomicAdd(&stats->shadowRays, (unsigned long long)rays);
    if (!inside) return;
    const V3 out = result * gp.invNumVpl;
    const long long q[3] = {to_fixed(out.x), to_fixed(out.y), to_fixed(out.z)};
    for (int c = 0; c < 3; c++) acc[i * 3 + c] = gp.doAccumulate ? acc[i * 3 + c] + q[c] : q[c];
}

// LVC splatColor (lvclighttracing.cu:348-387): every pixel gathers from its own window of
// light paths, so records are read straight from L2/HBM and rays are traced per lane.
__global__ void __launch_bounds__(GATHER_WARPS * 32)
gather_lvc_kernel(DevScene sc, GatherParams gp, const uint32_t* __restrict__ skip, const float4* __restrict__ gbuf,
                  const EvplpRecord* __restrict__ records, long long* __restrict__ acc, DevStats* stats) {
    __shared__ uint32_t sm[kSkipMatrixWords];
    for (int k = threadIdx.x; k < kSkipMatrixWords; k += blockDim.x) sm[k] = skip[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = gp.x0 + blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = gp.y0 + (blockIdx.y * gp.bandStride + gp.bandOffset) * 16 + (warp >> 1) * 4 + (lane >> 3);
    if (!(x < gp.x1 && y < gp.y1)) return;
    const size_t n = (size_t)gp.W * gp.H;
    const size_t i = (size_t)y * gp.W + x;
    float gw;
    Surface sf = load_surface(gbuf, n, i, &gw);
    unsigned long long rays = 0, pairs = 0;
    V3 result = v3s(0.0f);
    int ovf = 0;
    if (gw != 0.0f) {
        const V3 wi01 = normalize(gp.cameraPosition - sf.pos);
        Xorwow rng = xorwow_seed((uint32_t)i);
        xorwow_apply_matrix(rng, sm);
        const unsigned lightPathOffset = (unsigned)(det_min(xorwow_uniform(rng), 0.999999f) * (float)gp.numLightPaths);
        for (unsigned k = 0; k < gp.numVplLightPaths; k++) {
            const unsigned lightPathId = (k + lightPathOffset) % gp.numLightPaths;
            const size_t off = (size_t)lightPathId * gp.B1;
            for (unsigned j = 0; j < gp.B1; j++) {
                const float4* r = reinterpret_cast<const float4*>(records + off + j);
                const float4 a = r[0];
                if ((__float_as_uint(a.w) & EVPLP_FLAG_USABLE_VPL) == 0) continue;
                pairs++;
                const float4 b = r[1];
                const V3 vpos = v3(a.x, a.y, a.z), vn = v3(b.x, b.y, b.z);
                const V3 v12 = vpos - sf.pos;
                const float c1c2 = det_max(dot(sf.normal, v12), 0.0f) * det_max(-dot(vn, v12), 0.0f);
                if (c1c2 <= 0.000f) continue;
                rays++;
                if (trace_any(sc, vpos, -v12, (float)0.0001, (float)(1 - 0.0001), &ovf)) continue;
                const Vertex vp = load_vertex(r);
                result += vpl_shade(sf, wi01, vp, v12, c1c2, gp.misMode, gp.pdfMc, gp.clampingValue);
            }
        }
    }
    if (ovf) stats->stackOverflow = 1;
    if (rays) atomicAdd(&stats->shadowRays, rays);
    if (pairs) atomicAdd(&stats->gatherPairs, pairs);
    const V3 out = result * gp.invNumVpl;
    const long long q[3] = {to_fixed(out.x), to_fixed(out.y), to_fixed(out.z)};
    for (int c = 0; c < 3; c++) acc[i * 3 + c] = gp.doAccumulate ? acc[i * 3 + c] + q[c] : q[c];
}

// ------------------------------------------------------------------ path tracing (RtPt2) --
// pathtracing.cu:112-377: unidirectional path tracer from the G-buffer first hit with next-event estimation and
// balance-heuristic MIS -- the reference's own ground-truth generator.  One thread per pixel,
// per-pixel XORWOW stream curand_init(pixel, rngSeed, 0); per-lane traversal (paths are incoherent).
__global__ void __launch_bounds__(128) path_trace_kernel(DevScene sc, GatherParams gp, const uint32_t* __restrict__ skip,
                                                         const float4* __restrict__ gbuf, unsigned maxBounces,
                                                         long long* __restrict__ acc, DevStats* stats) {
    __shared__ uint32_t sm[kSkipMatrixWords];
    for (int k = threadIdx.x; k < kSkipMatrixWords; k += blockDim.x) sm[k] = skip[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = gp.x0 + blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = gp.y0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (!(x < gp.x1 && y < gp.y1)) return;
    const size_t n = (size_t)gp.W * gp.H;
    const size_t i = (size_t)y * gp.W + x;
    float gw;
    const Surface sf = load_surface(gbuf, n, i, &gw);
    V3 result = v3s(0.0f);
    int ovf = 0;
    unsigned long long rays = 0;
    const float lightExp = sc.lightIntensity[3];
    const float lightPdf = det_div(1.f, sc.lightArea);
    if (gw != 0.0f) {
        Xorwow rng = xorwow_seed((uint32_t)i);
        xorwow_apply_matrix(rng, sm);
        // ---- pathTraceSimple, first bounce (:232-305)
        const V3 cameraVec = normalize(sf.pos - gp.cameraPosition);
        V3 position = sf.pos, direction = v3s(0.f), attenuation = v3s(1.0f);
        float brdfPdfW = 0.f;
        bool alive = true;
        {
            V3 lightPosition, lightNormal;
            V3 lightValue = light_sample(sc, &lightPosition, &lightNormal, rng);
            V3 toLight = lightPosition - position;
            V3 toLightNorm = normalize(toLight);
            rays++;
            const bool hit = trace_any(sc, lightPosition, -toLight, 0.0001f, 1.0f - 0.0001f, &ovf);
            float maxLambert = max_color(sf.kd), maxPhong = max_color(sf.ks);
            float pSel = det_div(maxLambert, maxPhong + maxLambert);
            if (maxLambert + maxPhong <= 0.000001f) {
                alive = false;
            } else {
                float chooseMaterial = det_min(xorwow_uniform(rng), 0.999999f);
                if (chooseMaterial < pSel) {
                    if (!hit) {
                        float brdfPdf = lambert_pdf_a(sf.normal, lightNormal, toLight);
                        float weight = balance_heuristic(lightPdf, brdfPdf);
                        result += weight * lightValue * (sf.kd * kInvPi) * geometry_term(sf.normal, lightNormal, toLight) / pSel *
                                  phong_eval_f(lightNormal, -toLightNorm, lightNormal, lightExp);
                    }
                    attenuation *= lambert_sample(&direction, &brdfPdfW, sf.normal, sf.kd, rng) / pSel;
                } else {
                    if (!hit) {
                        float brdfPdf = phong_pdf_a(sf.normal, lightNormal, toLight, -cameraVec, sf.ks, sf.exponent);
                        float weight = balance_heuristic(lightPdf, brdfPdf);
                        result += weight * lightValue * phong_eval(-cameraVec, toLightNorm, sf.normal, sf.ks, sf.exponent) *
                                  geometry_term(sf.normal, lightNormal, toLight) / (1.0f - pSel) *
                                  phong_eval_f(lightNormal, -toLightNorm, lightNormal, lightExp);
                    }
                    attenuation *= phong_sample(&direction, &brdfPdfW, -cameraVec, sf.normal, sf.ks, sf.exponent, rng) / (1.0f - pSel);
                }
            }
        }
        // ---- bounces: rtTrace + rtMaterialClosestHit (:112-218, 307-318)
        for (unsigned b = 0; alive && b < maxBounces; b++) {
            const bool last = (b == maxBounces - 1);
            rays++;
            const RayHit h = trace_closest<true>(sc, position, direction, 0.00001f, 1e27f, &ovf);
            if (h.prim < 0) break;  // no miss program: nothing more is added
            const DevMaterial& mat = sc.mats[h.mat];
            const V3 geometryNormal = normalize(h.n);
            const V3 worldGeometryNormal = normalize(geometryNormal);
            const V3 ffNormal = faceforward(worldGeometryNormal, -direction, worldGeometryNormal);
            const V3 nextPosition = position + h.t * direction;
            if (dot(geometryNormal, direction) > 0.f) break;  // result = 0, done
            if (mat.lightIntensity[0] > 0.01f) {             // hit the light: MIS against the light sampling of the previous vertex
                float brdfPdfA = brdfPdfW * pdf_w2a(ffNormal, nextPosition - position);
                float weight = balance_heuristic(brdfPdfA, lightPdf);
                result += weight * attenuation * phong_eval_f(geometryNormal, normalize(position - nextPosition), geometryNormal, mat.lightIntensity[3]) *
                          v3(mat.lightIntensity[0], mat.lightIntensity[1], mat.lightIntensity[2]);
                break;
            }
            if (last) break;  // last bounce: no next-event estimation
            V3 lightPosition, lightNormal;
            V3 lightValue = light_sample(sc, &lightPosition, &lightNormal, rng);
            V3 toLight = lightPosition - nextPosition;
            V3 toLightNorm = normalize(toLight);
            rays++;
            const bool hit = trace_any(sc, lightPosition, -toLight, 0.00001f, 0.99999f, &ovf);
            const float2 t0 = sc.triUV[3 * (size_t)h.prim], t1 = sc.triUV[3 * (size_t)h.prim + 1], t2 = sc.triUV[3 * (size_t)h.prim + 2];
            const float w0 = 1.0f - h.beta - h.gamma;
            const float u = t1.x * h.beta + t2.x * h.gamma + t0.x * w0;
            const float v = t1.y * h.beta + t2.y * h.gamma + t0.y * w0;
            Texel4 tl = tex_fetch(mat.lambert, sc.texPool, u, v);
            Texel4 tp = tex_fetch(mat.phong, sc.texPool, u, v);
            Texel4 te = tex_fetch(mat.exponent, sc.texPool, u, v);
            const V3 kd = v3(tl.x, tl.y, tl.z), ks = v3(tp.x, tp.y, tp.z);
            const float phongExponent = te.x;
            float maxLambert = max_color(kd), maxPhong = max_color(ks);
            if (maxLambert + maxPhong <= 0.000001f) break;
            float pSel = det_div(maxLambert, maxPhong + maxLambert);
            float chooseMaterial = det_min(xorwow_uniform(rng), 0.999999f);
            const V3 toPrev = normalize(position - nextPosition);
            if (chooseMaterial < pSel) {
                if (!hit) {
                    float brdfPdf = lambert_pdf_a(ffNormal, lightNormal, toLight);
                    float weight = balance_heuristic(lightPdf, brdfPdf);
                    result += weight * lightValue * (kd * kInvPi) * geometry_term(ffNormal, lightNormal, toLight) * attenuation / pSel *
                              phong_eval_f(lightNormal, -toLightNorm, lightNormal, lightExp);
                }
                attenuation *= lambert_sample(&direction, &brdfPdfW, geometryNormal, kd, rng) / pSel;
            } else {
                if (!hit) {
                    float brdfPdf = phong_pdf_a(ffNormal, lightNormal, toLight, toPrev, ks, phongExponent);
                    float weight = balance_heuristic(lightPdf, brdfPdf);
                    result += weight * lightValue * phong_eval(toLightNorm, toPrev, ffNormal, ks, phongExponent) *
                              geometry_term(ffNormal, lightNormal, toLight) * attenuation / (1.0f - pSel) *
                              phong_eval_f(lightNormal, -toLightNorm, lightNormal, lightExp);
                }
                attenuation *= phong_sample(&direction, &brdfPdfW, toPrev, geometryNormal, ks, phongExponent, rng) / (1.0f - pSel);
            }
            float russian = pt_russian_prob(attenuation);
            if (xorwow_uniform(rng) >= russian) break;
            position = nextPosition;
            attenuation /= russian;
        }
    }
    if (ovf) stats->stackOverflow = 1;
    if (rays) atomicAdd(&stats->closestRays, rays);
    const long long q[3] = {to_fixed(result.x), to_fixed(result.y), to_fixed(result.z)};
    for (int c = 0; c < 3; c++) acc[i * 3 + c] = gp.doAccumulate ? acc[i * 3 + c] + q[c] : q[c];
}

// ------------------------------------------------------------------ photon splat --------
struct SplatParams {
    SplatUniforms U;
    V3 camFwd, camRight, camUp;
    float tanX, tanY, jx, jy, nearD;
    int x0, y0, x1, y1, W, H;
};

// Conservative pixel rectangle of the sphere (p, r): only a cull; the exact test is the
// per-texel |p - pos(x)|^2 <= r^2.
__device__ __forceinline__ void splat_rect(const SplatParams& sp, V3 p, int* rx0, int* ry0, int* rx1, int* ry1) {
    const double c0 = (double)p.x - sp.U.cameraPosition.x, c1 = (double)p.y - sp.U.cameraPosition.y,
                 c2 = (double)p.z - sp.U.cameraPosition.z;
    const double z = c0 * sp.camFwd.x + c1 * sp.camFwd.y + c2 * sp.camFwd.z;
    const double xv = c0 * sp.camRight.x + c1 * sp.camRight.y + c2 * sp.camRight.z;
    const double yv = c0 * sp.camUp.x + c1 * sp.camUp.y + c2 * sp.camUp.z;
    const double rr = (double)sp.U.radius * 1.001 + 1e-6;
    // G-buffer points lie at depth z >= nearDist along the forward axis (primary rays start at tmin = nearDist)
    const double zNear = (double)sp.nearD * 0.999;
    *rx0 = 0; *ry0 = 0; *rx1 = 0; *ry1 = 0;
    if (z + rr < zNear) return;  // sphere entirely behind the near plane: no texel can be inside it
    const double zn = fmax(z - rr, zNear), zf = z + rr;
    {
        const double a = xv - rr, b = xv + rr;
        const double lo = fmin(a / zn, a / zf) / sp.tanX + sp.jx, hi = fmax(b / zn, b / zf) / sp.tanX + sp.jx;
        const double plo = floor((lo + 1.0) * 0.5 * sp.W - 0.5) - 1.0, phi = ceil((hi + 1.0) * 0.5 * sp.W - 0.5) + 2.0;
        *rx0 = (int)fmax(0.0, fmin((double)sp.W, plo));
        *rx1 = (int)fmax(0.0, fmin((double)sp.W, phi));
    }
    {
        const double a = yv - rr, b = yv + rr;
        const double lo = fmin(a / zn, a / zf) / sp.tanY + sp.jy, hi = fmax(b / zn, b / zf) / sp.tanY + sp.jy;
        const double plo = floor((lo + 1.0) * 0.5 * sp.H - 0.5) - 1.0, phi = ceil((hi + 1.0) * 0.5 * sp.H - 0.5) + 2.0;
        *ry0 = (int)fmax(0.0, fmin((double)sp.H, plo));
        *ry1 = (int)fmax(0.0, fmin((double)sp.H, phi));
    }
}

// A group of G lanes (G = 1, 8 or 32) per usable photon: the group sweeps the photon's conservative
// screen rectangle, tests |p - pos(x)|^2 <= r^2 per texel and scatter-adds Q31.32 fixed-point
// contributions (order-independent => deterministic).  G is picked per launch from the expected
// footprint: tiny footprints (a few pixels) waste 31 of 32 lanes with a warp per photon, large
// ones (hundreds of pixels) want the whole warp.
template <int G>
__global__ void __launch_bounds__(256) splat_kernel(SplatParams sp, const float4* __restrict__ gbuf,
                                                    const int32_t* __restrict__ gprim,
                                                    const EvplpRecord* __restrict__ records,
                                                    const uint32_t* __restrict__ photonList,
                                                    const uint32_t* __restrict__ photonCount, long long* __restrict__ acc,
                                                    DevStats* stats) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % G;
    const uint32_t groupId = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint32_t numGroups = (gridDim.x * blockDim.x) / G;
    const uint32_t total = *photonCount;
    const size_t n = (size_t)sp.W * sp.H;
    const float r2 = sp.U.radius * sp.U.radius;
    unsigned frags = 0;
    for (uint32_t w = groupId; w < total; w += numGroups) {
        const uint32_t k = photonList[w];
        const float4* r = reinterpret_cast<const float4*>(records + k);
        const float4 r0 = __ldg(r);
        const V3 ppos = v3(r0.x, r0.y, r0.z);
        int rx0, ry0, rx1, ry1;
        splat_rect(sp, ppos, &rx0, &ry0, &rx1, &ry1);
        rx0 = max(rx0, sp.x0); ry0 = max(ry0, sp.y0); rx1 = min(rx1, sp.x1); ry1 = min(ry1, sp.y1);
        const int rw = rx1 - rx0, rh = ry1 - ry0;
        if (rw <= 0 || rh <= 0) continue;
        const int area = rw * rh;
        bool loaded = false;
        Vertex ph, prev;
        for (int t = sub; t < area; t += G) {
            const int py = t / rw, px = t - py * rw;
            const size_t i = (size_t)(ry0 + py) * sp.W + (rx0 + px);
            const float4 gp0 = __ldg(gbuf + i);
            const V3 d = ppos - v3(gp0.x, gp0.y, gp0.z);
            if (dot(d, d) > r2) continue;
            if (__ldg(gprim + i) < 0) continue;  // background texel (position 0): no surface
            if (!loaded) {
                ph = load_vertex(r);
                prev = load_vertex(r - 6);  // record k-1: same path (a photon is never a path's first record)
                loaded = true;
            }
            float gw;
            const Surface sf = load_surface(gbuf, n, i, &gw);
            V3 color;
            if (!splat_fragment(sp.U, sf, ph, prev, &color)) continue;
            frags++;
            const long long q[3] = {to_fixed(color.x), to_fixed(color.y), to_fixed(color.z)};
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (q[c]) atomicAdd(reinterpret_cast<unsigned long long*>(acc + i * 3 + c), (unsigned long long)q[c]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) frags += __shfl_xor_sync(0xffffffffu, frags, o);
    if (lane == 0 && frags) atomicAdd(&stats->splatFragments, (unsigned long long)frags);
}

// ---- tiled photon splat ------------------------------------------------------------------------
// Photons are binned to 16x16-pixel screen tiles; one block per (tile, photon chunk) keeps its 256
// G-buffer texels in registers, streams the tile's photons through shared memory, tests
// |p - pos(x)|^2 <= r^2 per (texel, photon), shades the hits and accumulates Q31.32 in registers:
// ONE accumulator update per pixel per launch, no per-fragment atomics, G-buffer read once.
// The per-photon half of the fragment shader (splat_prepare) runs once per photon in the binning pass.
constexpr int SPLAT_TILE = 16;
constexpr int SPLAT_BATCH = 64;      // photons staged per shared-memory batch
constexpr int SPLAT_CHUNK = 2048;    // photons of one tile handled by one block
constexpr int SPLAT_PREP_F4 = 5;     // float4s per prepared photon

struct TileGrid {
    int tx0, ty0, nx, ny;            // first tile (in tile units) and tile counts covering the launch rectangle
};

__device__ __forceinline__ void photon_tile_range(const SplatParams& sp, const TileGrid& tg, V3 pos, int* a0, int* b0, int* a1, int* b1) {
    int rx0, ry0, rx1, ry1;
    splat_rect(sp, pos, &rx0, &ry0, &rx1, &ry1);
    rx0 = max(rx0, sp.x0); ry0 = max(ry0, sp.y0); rx1 = min(rx1, sp.x1); ry1 = min(ry1, sp.y1);
    if (rx1 <= rx0 || ry1 <= ry0) { *a0 = *b0 = 0; *a1 = *b1 = -1; return; }
    *a0 = rx0 / SPLAT_TILE - tg.tx0; *a1 = (rx1 - 1) / SPLAT_TILE - tg.tx0;
    *b0 = ry0 / SPLAT_TILE - tg.ty0; *b1 = (ry1 - 1) / SPLAT_TILE - tg.ty0;
}

__global__ void splat_prepare_kernel(SplatParams sp, TileGrid tg, const EvplpRecord* __restrict__ records,
                                     const uint32_t* __restrict__ photonList, const uint32_t* __restrict__ photonCount,
                                     float4* __restrict__ prep, uint32_t* __restrict__ tileCount) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= *photonCount) return;
    const float4* r = reinterpret_cast<const float4*>(records + photonList[w]);
    const Vertex ph = load_vertex(r), prev = load_vertex(r - 6);  // record k-1: same path
    const SplatPhoton q = splat_prepare(sp.U, ph, prev);
    float4* o = prep + (size_t)w * SPLAT_PREP_F4;
    o[0] = make_float4(q.pos.x, q.pos.y, q.pos.z, __int_as_float(q.live));
    o[1] = make_float4(q.w12.x, q.w12.y, q.w12.z, q.weight);
    o[2] = make_float4(q.flux.x, q.flux.y, q.flux.z, q.dist2);
    o[3] = make_float4(q.prevN.x, q.prevN.y, q.prevN.z, 0.f);
    o[4] = make_float4(q.brdf2.x, q.brdf2.y, q.brdf2.z, 0.f);
    int a0, b0, a1, b1;
    photon_tile_range(sp, tg, q.pos, &a0, &b0, &a1, &b1);
    for (int b = b0; b <= b1; b++)
        for (int a = a0; a <= a1; a++) atomicAdd(&tileCount[b * tg.nx + a], 1u);
}

__global__ void splat_fill_kernel(SplatParams sp, TileGrid tg, const float4* __restrict__ prep, const uint32_t* __restrict__ photonCount,
                                  const uint32_t* __restrict__ tileOffset, uint32_t* __restrict__ tileCursor,
                                  uint32_t* __restrict__ tileList) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= *photonCount) return;
    const float4 p0 = prep[(size_t)w * SPLAT_PREP_F4];
    int a0, b0, a1, b1;
    photon_tile_range(sp, tg, v3(p0.x, p0.y, p0.z), &a0, &b0, &a1, &b1);
    for (int b = b0; b <= b1; b++)
        for (int a = a0; a <= a1; a++) {
            const int t = b * tg.nx + a;
            tileList[tileOffset[t] + atomicAdd(&tileCursor[t], 1u)] = w;
        }
}

#ifndef EVPLP_SPLAT_MINB
#define EVPLP_SPLAT_MINB 3   // 80 registers, 24 warps / SM (tuning: -DEVPLP_SPLAT_MINB=2 -> 101 registers)
#endif
__global__ void __launch_bounds__(256, EVPLP_SPLAT_MINB) splat_tile_kernel(SplatParams sp, TileGrid tg, const float4* __restrict__ gbuf,
                                                         const int32_t* __restrict__ gprim, const float4* __restrict__ prep,
                                                         const uint32_t* __restrict__ tileOffset, const uint32_t* __restrict__ tileList,
                                                         long long* __restrict__ acc, int useAtomics, DevStats* stats) {
    __shared__ float4 batch[SPLAT_BATCH * SPLAT_PREP_F4];
    const int tile = blockIdx.y * tg.nx + blockIdx.x;
    const uint32_t first = tileOffset[tile], count = tileOffset[tile + 1] - first;
    const uint32_t begin = blockIdx.z * SPLAT_CHUNK;
    if (begin >= count) return;
    const uint32_t end = min(count, begin + SPLAT_CHUNK);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = (tg.tx0 + blockIdx.x) * SPLAT_TILE + (warp & 1) * 8 + (lane & 7);
    const int y = (tg.ty0 + blockIdx.y) * SPLAT_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = x >= sp.x0 && x < sp.x1 && y >= sp.y0 && y < sp.y1;
    const size_t n = (size_t)sp.W * sp.H;
    const size_t i = inside ? (size_t)y * sp.W + x : 0;
    float gw;
    const Surface sf = load_surface(gbuf, n, i, &gw);
    const bool valid = inside && gprim[i] >= 0;
    const V3 w10 = normalize(sp.U.cameraPosition - sf.pos);
    const float r2 = sp.U.radius * sp.U.radius;
    const float r2Cull = r2 * 1.001f;
    const float invR2 = det_div(1.0f, sp.U.radius * sp.U.radius);
    const float invN = det_div(1.0f, (float)sp.U.numLightPaths);
    // box of the surface points of this warp's sub-block (empty when no texel of it is valid: then nothing survives the cull)
    V3 blo = valid ? sf.pos : v3s(INFINITY), bhi = valid ? sf.pos : v3s(-INFINITY);
    for (int o = 16; o > 0; o >>= 1) {
        blo = v3(fminf(blo.x, __shfl_xor_sync(0xffffffffu, blo.x, o)), fminf(blo.y, __shfl_xor_sync(0xffffffffu, blo.y, o)), fminf(blo.z, __shfl_xor_sync(0xffffffffu, blo.z, o)));
        bhi = v3(fmaxf(bhi.x, __shfl_xor_sync(0xffffffffu, bhi.x, o)), fmaxf(bhi.y, __shfl_xor_sync(0xffffffffu, bhi.y, o)), fmaxf(bhi.z, __shfl_xor_sync(0xffffffffu, bhi.z, o)));
    }
    long long a0 = 0, a1 = 0, a2 = 0;
    unsigned frags = 0;
    // The photons of a batch are gathered through the tile list (two dependent loads); the NEXT batch is fetched into registers
    // while the current one is processed, so that latency is off the critical path (ncu: 28 % of the kernel's stall samples sat
    // on this gather and 2.1 stalled warps per issue on its barrier).  256 threads x 2 float4 cover the 64 x 5 float4 of a batch.
    static_assert(SPLAT_BATCH * SPLAT_PREP_F4 <= 2 * 256, "two registers per thread stage one batch");
    float4 pre0 = make_float4(0.f, 0.f, 0.f, 0.f), pre1 = pre0;
    auto fetch = [&](uint32_t base) {
        const uint32_t nbn = base < end ? min((uint32_t)SPLAT_BATCH, end - base) : 0u;
        const uint32_t k0 = threadIdx.x, k1 = threadIdx.x + 256u;
        if (k0 < nbn * SPLAT_PREP_F4) pre0 = __ldg(prep + (size_t)__ldg(tileList + first + base + k0 / SPLAT_PREP_F4) * SPLAT_PREP_F4 + k0 % SPLAT_PREP_F4);
        if (k1 < nbn * SPLAT_PREP_F4) pre1 = __ldg(prep + (size_t)__ldg(tileList + first + base + k1 / SPLAT_PREP_F4) * SPLAT_PREP_F4 + k1 % SPLAT_PREP_F4);
    };
    fetch(begin);
    for (uint32_t base = begin; base < end; base += SPLAT_BATCH) {
        const uint32_t nb = min((uint32_t)SPLAT_BATCH, end - base);
        __syncthreads();   // every warp is done with the previous batch
        if (threadIdx.x < nb * SPLAT_PREP_F4) batch[threadIdx.x] = pre0;
        if (threadIdx.x + 256u < nb * SPLAT_PREP_F4) batch[threadIdx.x + 256u] = pre1;
        __syncthreads();
        fetch(base + SPLAT_BATCH);
        // phase 0 + 1, per half of the batch: lane l asks whether photon l of the half can touch this warp's 8x4-texel
        // sub-block at all (sphere vs. the box of the sub-block's surface points, with a margin far above the rounding of the
        // exact test, so the cull never changes a decision; a footprint covers a few of the tile's eight sub-blocks), then
        // every texel runs the exact radius test against the survivors only (warp-uniform loop) -> 64-bit hit mask
        unsigned hitsLo = 0u, hitsHi = 0u;
#pragma unroll
        for (int h = 0; h < SPLAT_BATCH / 32; h++) {
            const uint32_t jl = (uint32_t)(lane + 32 * h);
            bool keep = false;
            if (jl < nb) {
                const float4 p0 = batch[jl * SPLAT_PREP_F4];
                const float dx = fmaxf(fmaxf(blo.x - p0.x, p0.x - bhi.x), 0.0f), dy = fmaxf(fmaxf(blo.y - p0.y, p0.y - bhi.y), 0.0f),
                            dz = fmaxf(fmaxf(blo.z - p0.z, p0.z - bhi.z), 0.0f);
                keep = !(dx * dx + dy * dy + dz * dz > r2Cull);
            }
            unsigned hm = 0u;
            for (unsigned cm = __ballot_sync(0xffffffffu, keep); cm; cm &= cm - 1u) {
                const int b = __ffs((int)cm) - 1;
                const float4 p0 = batch[(b + 32 * h) * SPLAT_PREP_F4];
                const V3 d = v3(p0.x, p0.y, p0.z) - sf.pos;
                if (!(dot(d, d) > r2)) hm |= 1u << b;
            }
            if (h == 0) hitsLo = hm; else hitsHi = hm;
        }
        unsigned long long hits = valid ? (((unsigned long long)hitsHi << 32) | hitsLo) : 0ull;
        // phase 2: every lane shades ITS OWN hits (ascending j), so a warp iterates max-hits-per-lane times instead of
        // once per photon that any of its lanes touches
        while (hits) {
            const int j = __ffsll((long long)hits) - 1;
            hits &= hits - 1ull;
            const float4 p0 = batch[j * SPLAT_PREP_F4], p1 = batch[j * SPLAT_PREP_F4 + 1], p2 = batch[j * SPLAT_PREP_F4 + 2],
                         p3 = batch[j * SPLAT_PREP_F4 + 3], p4 = batch[j * SPLAT_PREP_F4 + 4];
            SplatPhoton q;
            q.pos = v3(p0.x, p0.y, p0.z); q.live = __float_as_int(p0.w);
            q.w12 = v3(p1.x, p1.y, p1.z); q.weight = p1.w;
            q.flux = v3(p2.x, p2.y, p2.z); q.dist2 = p2.w;
            q.prevN = v3(p3.x, p3.y, p3.z); q.brdf2 = v3(p4.x, p4.y, p4.z);
            V3 color;
            if (!splat_shade(sp.U, invR2, invN, sf, w10, q, &color)) continue;
            frags++;
            a0 += to_fixed(color.x); a1 += to_fixed(color.y); a2 += to_fixed(color.z);
        }
    }
    for (int o = 16; o > 0; o >>= 1) frags += __shfl_xor_sync(0xffffffffu, frags, o);
    if (lane == 0 && frags) atomicAdd(&stats->splatFragments, (unsigned long long)frags);
    if (!valid) return;
    if (useAtomics) {
        if (a0) atomicAdd(reinterpret_cast<unsigned long long*>(acc + i * 3), (unsigned long long)a0);
        if (a1) atomicAdd(reinterpret_cast<unsigned long long*>(acc + i * 3 + 1), (unsigned long long)a1);
        if (a2) atomicAdd(reinterpret_cast<unsigned long long*>(acc + i * 3 + 2), (unsigned long long)a2);
    } else {
        acc[i * 3] += a0; acc[i * 3 + 1] += a1; acc[i * 3 + 2] += a2;
    }
}

// out[0] = largest per-tile count; total = sum of the counts in 64 bits (the 32-bit scan wraps beyond 2^32 entries: the caller
// compares THIS total with the capacity before it trusts the offsets)
__global__ void tile_summary_kernel(const uint32_t* __restrict__ tileCount, int numTiles, uint32_t* __restrict__ out /* [0] max */,
                                    unsigned long long* __restrict__ total) {
    uint32_t m = 0;
    unsigned long long sum = 0ull;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < numTiles; t += gridDim.x * blockDim.x) { m = max(m, tileCount[t]); sum += tileCount[t]; }
    for (int o = 16; o > 0; o >>= 1) { m = max(m, __shfl_xor_sync(0xffffffffu, m, o)); sum += __shfl_xor_sync(0xffffffffu, sum, o); }
    if ((threadIdx.x & 31) == 0 && m) { atomicMax(out, m); atomicAdd(total, sum); }
}

// ------------------------------------------------------------------ light / resolve -----
// runLightProgram (rtcomphoton.h:839-855, 985-995): the reference draws the light mesh with the UN-jittered view-projection
// ("we don't jitter light source") against the frame's depth buffer.  Here: un-jittered primary rays over the screen rectangle
// of the light mesh; a pixel is lit iff its closest hit is a light triangle.  The mask is the same in every iteration, so it is
// written, not accumulated (a silhouette pixel that saw the light in one jittered G-buffer no longer stays masked for good).
__global__ void __launch_bounds__(128) light_pass_kernel(DevScene sc, CamParams cam, int W, int H, int rx0, int ry0, int rx1, int ry1,
                                                         uint32_t* __restrict__ light, DevStats* stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = rx0 + blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = ry0 + blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= rx1 || y >= ry1) return;
    const float cx = det_div((float)x + 0.5f, (float)W) * 2.0f - 1.0f;
    const float cy = det_div((float)y + 0.5f, (float)H) * 2.0f - 1.0f;
    const V3 dir = cam.fwd + cam.right * (cx * cam.tanX) + cam.up * (cy * cam.tanY);
    int ovf = 0;
    const RayHit hit = trace_closest<false>(sc, cam.pos, dir, cam.nearD, cam.farD, &ovf);
    if (ovf) stats->stackOverflow = 1;
    light[(size_t)y * W + x] = (hit.prim >= sc.lightFirst && hit.prim < sc.lightFirst + sc.lightCount) ? 1u : 0u;
}

__global__ void add_count_kernel(long long* count, long long n) { count[0] += n; }

struct ResolveParams {
    float vplScale, photonScale, lightScale;
    int gamma;
    float lightDisplay[3];
};

__global__ void resolve_kernel(ResolveParams rp, const long long* __restrict__ vpl, const long long* __restrict__ photon,
                               const uint32_t* __restrict__ light, size_t n, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool lit = light[i] != 0;
    const float lightX = (lit ? rp.lightDisplay[0] : 0.f) * rp.lightScale;
    const float stepv = (0.0f >= lightX) ? 1.0f : 0.0f;  // step(edge = lightColor.x, x = 0)  final.frag:26
    for (int c = 0; c < 3; c++) {
        float vplColor = (float)((double)vpl[i * 3 + c] * (1.0 / 4294967296.0)) * rp.vplScale;
        float pmColor = (float)((double)photon[i * 3 + c] * (1.0 / 4294967296.0)) * rp.photonScale;
        float lightColor = (lit ? rp.lightDisplay[c] : 0.f) * rp.lightScale;
        float sum = stepv * (vplColor + pmColor) + lightColor;
        out[i * 3 + c] = rp.gamma ? det_powf(sum, det_div(1.0f, 2.2f)) : sum;
    }
}

// ------------------------------------------------------------------ list compaction -----
struct FlagPred {
    const EvplpRecord* records;
    uint32_t mask;
    __device__ bool operator()(uint32_t k) const { return (records[k].flags & mask) != 0; }
};

static cudaError_t compact_records(EvplpContext* c, uint64_t first, uint64_t count, uint32_t mask, DevBuf<uint32_t>& list,
                                   uint32_t* devCount) {
    if (count > 0x7fffffffull || first + count > 0xffffffffull) return cudaErrorInvalidValue;   // 32-bit record indices, int num_items of cub
    cudaError_t e = list.reserve(count ? count : 1);
    if (e != cudaSuccess) return e;
    thrust::counting_iterator<uint32_t> it((uint32_t)first);
    FlagPred pred{c->records.p, mask};
    size_t tempBytes = 0;
    e = cub::DeviceSelect::If(nullptr, tempBytes, it, list.p, devCount, (int)count, pred, c->stream);
    if (e != cudaSuccess) return e;
    e = c->sortTemp.reserve(tempBytes);
    if (e != cudaSuccess) return e;
    e = cub::DeviceSelect::If(c->sortTemp.p, tempBytes, it, list.p, devCount, (int)count, pred, c->stream);
    c->launches += 2;
    return e;
}

__global__ void iota_kernel(uint32_t* out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// ------------------------------------------------------------------ launchers -----------
cudaError_t launch_gbuffer(EvplpContext* c) {
    dim3 grid((c->W + 15) / 16, (c->H + 7) / 8);
    c->stageBegin(ST_GBUFFER);
    gbuffer_kernel<<<grid, 128, 0, c->stream>>>(c->scene(), cam_of(c->params), c->W, c->H, c->gbuf.p, c->gprim.p, c->devStats.p);
    c->stageEnd(ST_GBUFFER);
    c->launches++;
    c->stats.closestRays += (uint64_t)c->W * c->H;
    return cudaGetLastError();
}

cudaError_t launch_light_trace(EvplpContext* c, uint32_t rngSeed, uint32_t firstPath, uint32_t numPaths) {
    (void)rngSeed;
    if (numPaths == 0) return cudaSuccess;
    const uint32_t B1 = c->params.numPhotonsPerLightPath;
    c->stageBegin(ST_LIGHT_TRACE);
    light_trace_kernel<<<(numPaths + 127) / 128, 128, 0, c->stream>>>(c->scene(), c->skipTable.p, c->records.p, firstPath,
                                                                      numPaths, B1, c->devStats.p);
    c->stageEnd(ST_LIGHT_TRACE);
    c->launches++;
    return cudaGetLastError();
}

static GatherParams gather_params(EvplpContext* c, EvplpTile t) {
    const EvplpParams& P = c->params;
    GatherParams g;
    g.cameraPosition = v3p(P.cameraPosition);
    g.misMode = P.misMode; g.pdfMc = P.pdfMc; g.clampingValue = P.clampingValue;
    g.invNumVpl = 1.0f / (float)P.numVplLightPaths;
    g.doAccumulate = P.doAccumulate;
    g.x0 = t.x0; g.y0 = t.y0; g.x1 = t.x1; g.y1 = t.y1; g.W = c->W; g.H = c->H;
    g.numChunks = 1;
    g.vslRadius = P.vslRadius; g.vslInvPiRadius2 = P.vslInvPiRadius2;
    g.numLightPaths = P.numLightPaths; g.numVplLightPaths = P.numVplLightPaths; g.B1 = P.numPhotonsPerLightPath;
    g.shaftMode = c->opt.gatherMode == 2 ? 1 : 0;  // VSL gather: sampling-bound, the shaft brings nothing there (measured); opt-in with gather_mode = 2
    g.shaftCandMax = c->opt.shaftCandMax < 1 ? 1 : (c->opt.shaftCandMax > SHAFT_CAND ? SHAFT_CAND : c->opt.shaftCandMax);
    g.bandStride = c->opt.bandStride > 0 ? c->opt.bandStride : 1;
    g.bandOffset = c->opt.bandStride > 0 ? c->opt.bandOffset : 0;
    return g;
}

cudaError_t launch_gather(EvplpContext* c, EvplpTile t, int mode) {
    const EvplpParams& P = c->params;
    GatherParams g = gather_params(c, t);
    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    if (tw <= 0 || th <= 0) return cudaSuccess;
    const int bands = (th + 15) / 16;
    if (g.bandOffset >= bands) return cudaSuccess;  // this rank owns no band of the tile
    const int ownBands = (bands - g.bandOffset + g.bandStride - 1) / g.bandStride;
    uint64_t ownRows = 0;
    for (int b = g.bandOffset; b < bands; b += g.bandStride) ownRows += (uint64_t)((b + 1) * 16 <= th ? 16 : th - b * 16);
    dim3 grid((tw + 15) / 16, ownBands, 1);
    cudaError_t e;
    if (mode == EVPLP_GATHER_LVC) {
        c->stageBegin(ST_GATHER);
        gather_lvc_kernel<<<grid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->skipMatrix.p, c->gbuf.p, c->records.p,
                                                                      c->accVpl.p, c->devStats.p);
        c->stageEnd(ST_GATHER);
        c->launches++;
        return cudaGetLastError();
    }
    // usable VPLs of the record prefix, in order
    const uint64_t prefix = (uint64_t)P.numPhotonsPerLightPath * P.numVplLightPaths;
    uint32_t* devCount = c->counters.p + 3;
    e = compact_records(c, 0, prefix, EVPLP_FLAG_USABLE_VPL, c->vplList, devCount);
    if (e != cudaSuccess) return e;
    uint32_t count = 0;
    e = cudaMemcpyAsync(&count, devCount, 4, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return e;
    // The cluster gather pays off when the VPLs are dense enough for 16 Morton neighbours to form a small box (measured:
    // +24 % at 47.8 k VPLs, -30 % at 1.5 k); gather_algo 1 (default) picks it from 16384 usable VPLs on, 2 forces it, 0 never
    // uses it.  gather_chunks = 1 always asks for the bit-exact record order.
    if (mode == EVPLP_GATHER_VPL && c->opt.gatherChunks != 1 &&
        (c->opt.gatherAlgo == 2 || (c->opt.gatherAlgo == 1 && count >= 16384u)))
        return launch_gather_cluster(c, t, g, count);
    const bool tilePartition = mode == EVPLP_GATHER_VPL && c->opt.gatherPersistent;
    const TileShare share = tile_share(c, t);
    if (tilePartition && share.ownedTiles == 0) return cudaSuccess;
    c->stats.gatherPairs += (uint64_t)count * (tilePartition ? share.pixels : (uint64_t)tw * ownRows);
    if (mode == EVPLP_GATHER_VSL) {
        c->stageBegin(ST_GATHER);
        gather_vsl_kernel<<<grid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->skipMatrix.p, c->gbuf.p, c->records.p,
                                                                      c->vplList.p, devCount, c->accVpl.p, c->devStats.p);
        c->stageEnd(ST_GATHER);
        c->launches++;
        return cudaGetLastError();
    }
    // split the VPL list over gridDim.z when the tile alone cannot fill the GPU
    unsigned chunks = 1;
    if (c->opt.gatherChunks > 0) {
        chunks = (unsigned)c->opt.gatherChunks;
    } else {
        // enough work items per resident warp that the tail stays short (small image shares: multi-GPU partition of one frame)
        const unsigned blocks = tilePartition ? (share.ownedTiles + GATHER_WARPS - 1) / GATHER_WARPS : grid.x * grid.y;
        const unsigned want = tilePartition ? 148u * 4u * 24u : 148u * 3u * 6u;
        if (blocks < want) chunks = (want + blocks - 1) / blocks;
        const unsigned maxChunks = (count + 4 * GATHER_BATCH - 1) / (4 * GATHER_BATCH);
        if (chunks > maxChunks) chunks = maxChunks ? maxChunks : 1;
    }
    g.numChunks = chunks;
    grid.z = chunks;
    if (chunks > 1 && !P.doAccumulate) {
        dim3 cg((tw + 127) / 128, th);
        clear_tile_kernel<<<cg, 128, 0, c->stream>>>(c->accVpl.p, c->W, t.x0, t.y0, t.x1, t.y1);
        c->launches++;
    }
    g.vgx = grid.x; g.vgy = grid.y; g.vgz = grid.z;
    g.persistent = c->opt.gatherPersistent;
    g.tilePartition = tilePartition ? 1 : 0;
    g.tilesX = share.tilesX; g.pitchX = share.pitchX; g.ownedTiles = share.ownedTiles; g.tStride = share.stride; g.tOffset = share.offset;
    g.shaftStreak = c->opt.shaftStreak > 0 ? c->opt.shaftStreak : 1;
    g.shaftSkip = c->opt.shaftSkip;
    uint32_t* tileCounter = c->counters.p + 2;  // (slots 0-2 belong to the BVH build, which is over by now)
    dim3 lgrid = grid;
    const uint32_t* tileOrder = nullptr;
    uint32_t* tileCost = nullptr;
    if (g.persistent) {
        e = cudaMemsetAsync(tileCounter, 0, sizeof(uint32_t), c->stream);
        if (e != cudaSuccess) return e;
        if (c->opt.gatherLpt) {
            // longest-processing-time-first: order the tiles by the cycles they took in the previous launch of this grid
            const uint32_t vTotal = tilePartition ? share.ownedTiles * chunks : grid.x * grid.y * grid.z * GATHER_WARPS;
            const uint64_t sig[4] = {((uint64_t)grid.x << 40) | ((uint64_t)grid.y << 20) | grid.z,
                                     ((uint64_t)(uint32_t)t.x0 << 32) | (uint32_t)t.y0, ((uint64_t)(uint32_t)t.x1 << 32) | (uint32_t)t.y1,
                                     ((uint64_t)(uint32_t)g.bandStride << 32) | (uint32_t)g.bandOffset};
            const bool same = c->gatherCostValid && memcmp(sig, c->gatherSig, sizeof(sig)) == 0;
            if (!same) {
                if ((e = c->gatherCost.reserve(vTotal)) != cudaSuccess) return e;
                if ((e = c->gatherCostSorted.reserve(vTotal)) != cudaSuccess) return e;
                if ((e = c->gatherIota.reserve(vTotal)) != cudaSuccess) return e;
                if ((e = c->gatherOrder.reserve(vTotal)) != cudaSuccess) return e;
                iota_kernel<<<(vTotal + 255) / 256, 256, 0, c->stream>>>(c->gatherIota.p, vTotal);
                c->launches++;
                memcpy(c->gatherSig, sig, sizeof(sig));
                c->gatherCostValid = true;
            } else {
                size_t tempBytes = 0;
                e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tempBytes, c->gatherCost.p, c->gatherCostSorted.p, c->gatherIota.p,
                                                              c->gatherOrder.p, (int)vTotal, 0, 32, c->stream);
                if (e != cudaSuccess) return e;
                if ((e = c->sortTemp.reserve(tempBytes)) != cudaSuccess) return e;
                e = cub::DeviceRadixSort::SortPairsDescending(c->sortTemp.p, tempBytes, c->gatherCost.p, c->gatherCostSorted.p, c->gatherIota.p,
                                                              c->gatherOrder.p, (int)vTotal, 0, 32, c->stream);
                if (e != cudaSuccess) return e;
                c->launches += 2;
                tileOrder = c->gatherOrder.p;
            }
            tileCost = c->gatherCost.p;
        }
        const unsigned resident = 148u * 5u;  // at most 5 blocks of 256 threads fit an SM at any register count used here
        const unsigned blocks = tilePartition ? (share.ownedTiles * chunks + GATHER_WARPS - 1) / GATHER_WARPS : grid.x * grid.y * grid.z;
        lgrid = dim3(blocks < resident ? blocks : resident, 1, 1);
    }
    c->stageBegin(ST_GATHER);
    if (c->opt.gatherMode >= 1) {
        const int mb = c->opt.gatherMinBlocks ? c->opt.gatherMinBlocks : 4;  // measured: 64 registers / 4 blocks per SM wins by 5 % here
        if (mb == 2)
            gather_vpl_kernel<2, true><<<lgrid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->gbuf.p, c->records.p, c->vplList.p, devCount, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost);
        else if (mb == 5)
            gather_vpl_kernel<5, true><<<lgrid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->gbuf.p, c->records.p, c->vplList.p, devCount, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost);
        else if (mb == 4)
            gather_vpl_kernel<4, true><<<lgrid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->gbuf.p, c->records.p, c->vplList.p, devCount, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost);
        else
            gather_vpl_kernel<3, true><<<lgrid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->gbuf.p, c->records.p, c->vplList.p, devCount, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost);
    } else {
        switch (c->opt.gatherMinBlocks) {
            case 2: gather_vpl_kernel<2, false><<<lgrid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->gbuf.p, c->records.p, c->vplList.p, devCount, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost); break;
            case 4: gather_vpl_kernel<4, false><<<lgrid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->gbuf.p, c->records.p, c->vplList.p, devCount, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost); break;
            default: gather_vpl_kernel<3, false><<<lgrid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), g, c->gbuf.p, c->records.p, c->vplList.p, devCount, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost); break;
        }
    }
    c->stageEnd(ST_GATHER);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_path_trace(EvplpContext* c, EvplpTile t, uint32_t maxBounces) {
    GatherParams g = gather_params(c, t);
    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    if (tw <= 0 || th <= 0) return cudaSuccess;
    dim3 grid((tw + 15) / 16, (th + 7) / 8);
    c->stageBegin(ST_GATHER);
    path_trace_kernel<<<grid, 128, 0, c->stream>>>(c->scene(), g, c->skipMatrix.p, c->gbuf.p, maxBounces, c->accVpl.p, c->devStats.p);
    c->stageEnd(ST_GATHER);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_splat(EvplpContext* c, uint64_t firstRecord, uint64_t numRecords, EvplpTile t) {
    const EvplpParams& P = c->params;
    if (numRecords == 0) return cudaSuccess;
    uint32_t* devCount = c->counters.p + 3;
    cudaError_t e = compact_records(c, firstRecord, numRecords, EVPLP_FLAG_USABLE_PHOTON, c->photonList, devCount);
    if (e != cudaSuccess) return e;
    uint32_t count = 0;
    e = cudaMemcpyAsync(&count, devCount, 4, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return e;
    c->stats.splatPhotons += count;
    if (count == 0 || !(P.radius > 0.0f)) return cudaSuccess;  // radiusPercentage 0 => degenerate spheres, no fragments
    SplatParams sp;
    sp.U.cameraPosition = v3p(P.cameraPosition);
    sp.U.radius = P.radius; sp.U.pdfMc = P.pdfMc; sp.U.clampingValue = P.clampingValue;
    sp.U.misMode = P.misMode; sp.U.numLightPaths = P.numLightPaths;
    sp.camFwd = v3p(P.camForward); sp.camRight = v3p(P.camRight); sp.camUp = v3p(P.camUp);
    sp.tanX = P.tanHalfFovX; sp.tanY = P.tanHalfFovY; sp.jx = P.jitter[0]; sp.jy = P.jitter[1]; sp.nearD = P.nearDist;
    sp.x0 = t.x0; sp.y0 = t.y0; sp.x1 = t.x1; sp.y1 = t.y1; sp.W = c->W; sp.H = c->H;
    if (c->opt.splatMode != 1) {
        // ---- tiled path: bin -> scan -> fill -> per-tile accumulation
        TileGrid tg;
        tg.tx0 = t.x0 / SPLAT_TILE; tg.ty0 = t.y0 / SPLAT_TILE;
        tg.nx = (t.x1 - 1) / SPLAT_TILE - tg.tx0 + 1; tg.ny = (t.y1 - 1) / SPLAT_TILE - tg.ty0 + 1;
        const int numTiles = tg.nx * tg.ny;
        e = c->splatPrep.reserve((size_t)count * SPLAT_PREP_F4); if (e != cudaSuccess) return e;
        e = c->tileCount.reserve((size_t)numTiles + 1); if (e != cudaSuccess) return e;
        e = c->tileOffset.reserve((size_t)numTiles + 1); if (e != cudaSuccess) return e;
        e = c->tileCursor.reserve((size_t)numTiles + 6); if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(c->tileCount.p, 0, sizeof(uint32_t) * (numTiles + 1), c->stream); if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(c->tileCursor.p, 0, sizeof(uint32_t) * (numTiles + 6), c->stream); if (e != cudaSuccess) return e;
        c->stageBegin(ST_SPLAT);
        const unsigned pb = (count + 255) / 256;
        splat_prepare_kernel<<<pb, 256, 0, c->stream>>>(sp, tg, c->records.p, c->photonList.p, devCount, c->splatPrep.p, c->tileCount.p);
        size_t tempBytes = 0;
        e = cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, c->tileCount.p, c->tileOffset.p, numTiles + 1, c->stream); if (e != cudaSuccess) return e;
        e = c->sortTemp.reserve(tempBytes); if (e != cudaSuccess) return e;
        e = cub::DeviceScan::ExclusiveSum(c->sortTemp.p, tempBytes, c->tileCount.p, c->tileOffset.p, numTiles + 1, c->stream); if (e != cudaSuccess) return e;
        uint32_t* summary = c->tileCursor.p + numTiles + 1;  // max photons per tile
        unsigned long long* total64 = reinterpret_cast<unsigned long long*>(c->tileCursor.p + (((size_t)numTiles + 3) & ~(size_t)1));  // 8-byte aligned slot
        tile_summary_kernel<<<32, 256, 0, c->stream>>>(c->tileCount.p, numTiles, summary, total64);
        c->launches += 4;
        unsigned long long totalEntries = 0;   // 64-bit: the 32-bit offsets are only trusted when this fits the capacity below
        uint32_t maxPerTile = 0;
        e = cudaMemcpyAsync(&totalEntries, total64, 8, cudaMemcpyDeviceToHost, c->stream); if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(&maxPerTile, summary, 4, cudaMemcpyDeviceToHost, c->stream); if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) return e;
        if (totalEntries == 0) { c->stageEnd(ST_SPLAT); return cudaSuccess; }
        if ((uint64_t)totalEntries <= (uint64_t)c->opt.splatMaxEntries && totalEntries <= 0xffffffffull) {
            e = c->tileList.reserve((size_t)totalEntries); if (e != cudaSuccess) return e;
            splat_fill_kernel<<<pb, 256, 0, c->stream>>>(sp, tg, c->splatPrep.p, devCount, c->tileOffset.p, c->tileCursor.p, c->tileList.p);
            const unsigned chunks = (maxPerTile + SPLAT_CHUNK - 1) / SPLAT_CHUNK;
            dim3 grid(tg.nx, tg.ny, chunks);
            splat_tile_kernel<<<grid, 256, 0, c->stream>>>(sp, tg, c->gbuf.p, c->gprim.p, c->splatPrep.p, c->tileOffset.p, c->tileList.p,
                                                           c->accPhoton.p, chunks > 1 ? 1 : 0, c->devStats.p);
            c->stageEnd(ST_SPLAT);
            c->launches += 2;
            return cudaGetLastError();
        }
        // footprints so large that the tile lists would not fit: fall through to the scatter kernel
    }
    const int G = c->opt.splatGroup > 0 ? c->opt.splatGroup : 32;
    const uint64_t threadsWanted = (uint64_t)count * (uint64_t)G;
    uint64_t blocks = (threadsWanted + 255) / 256;
    const uint64_t maxBlocks = 148ull * 8ull * 8ull;
    if (blocks > maxBlocks) blocks = maxBlocks;
    c->stageBegin(ST_SPLAT);
    if (G == 32) splat_kernel<32><<<(unsigned)blocks, 256, 0, c->stream>>>(sp, c->gbuf.p, c->gprim.p, c->records.p, c->photonList.p, devCount, c->accPhoton.p, c->devStats.p);
    else if (G == 8) splat_kernel<8><<<(unsigned)blocks, 256, 0, c->stream>>>(sp, c->gbuf.p, c->gprim.p, c->records.p, c->photonList.p, devCount, c->accPhoton.p, c->devStats.p);
    else splat_kernel<1><<<(unsigned)blocks, 256, 0, c->stream>>>(sp, c->gbuf.p, c->gprim.p, c->records.p, c->photonList.p, devCount, c->accPhoton.p, c->devStats.p);
    c->stageEnd(ST_SPLAT);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_light_pass(EvplpContext* c) {
    CamParams cam = cam_of(c->params);
    cam.jx = 0.f; cam.jy = 0.f;
    // conservative screen rectangle of the light mesh's bounding box (whole frame when a corner is not in front of the camera)
    int rx0 = 0, ry0 = 0, rx1 = c->W, ry1 = c->H;
    {
        double lo[2] = {1e30, 1e30}, hi[2] = {-1e30, -1e30};
        bool ok = true;
        for (int k = 0; k < 8 && ok; k++) {
            const double q[3] = {(k & 1) ? c->lightBoxMax[0] : c->lightBoxMin[0], (k & 2) ? c->lightBoxMax[1] : c->lightBoxMin[1],
                                 (k & 4) ? c->lightBoxMax[2] : c->lightBoxMin[2]};
            const double d[3] = {q[0] - cam.pos.x, q[1] - cam.pos.y, q[2] - cam.pos.z};
            const double z = d[0] * cam.fwd.x + d[1] * cam.fwd.y + d[2] * cam.fwd.z;
            if (!(z > 1e-3)) { ok = false; break; }
            const double nx = (d[0] * cam.right.x + d[1] * cam.right.y + d[2] * cam.right.z) / (z * cam.tanX);
            const double ny = (d[0] * cam.up.x + d[1] * cam.up.y + d[2] * cam.up.z) / (z * cam.tanY);
            const double pxl = (nx + 1.0) * 0.5 * c->W - 0.5, pyl = (ny + 1.0) * 0.5 * c->H - 0.5;
            lo[0] = fmin(lo[0], pxl); hi[0] = fmax(hi[0], pxl); lo[1] = fmin(lo[1], pyl); hi[1] = fmax(hi[1], pyl);
        }
        if (ok) {
            rx0 = (int)fmax(0.0, fmin((double)c->W, floor(lo[0]) - 1.0)); rx1 = (int)fmax(0.0, fmin((double)c->W, ceil(hi[0]) + 2.0));
            ry0 = (int)fmax(0.0, fmin((double)c->H, floor(lo[1]) - 1.0)); ry1 = (int)fmax(0.0, fmin((double)c->H, ceil(hi[1]) + 2.0));
        }
    }
    if (rx1 <= rx0 || ry1 <= ry0) return cudaSuccess;   // the light is off screen: the layer stays as cleared
    dim3 grid((rx1 - rx0 + 15) / 16, (ry1 - ry0 + 7) / 8);
    light_pass_kernel<<<grid, 128, 0, c->stream>>>(c->scene(), cam, c->W, c->H, rx0, ry0, rx1, ry1, c->accLight.p, c->devStats.p);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_add_count(EvplpContext* c, long long n) {
    add_count_kernel<<<1, 1, 0, c->stream>>>(c->accCount.p, n);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_resolve(EvplpContext* c, float vplScale, float photonScale, float lightScale, int gamma) {
    const size_t n = (size_t)c->W * c->H;
    ResolveParams rp;
    rp.vplScale = vplScale; rp.photonScale = photonScale; rp.lightScale = lightScale; rp.gamma = gamma;
    for (int k = 0; k < 3; k++) rp.lightDisplay[k] = c->lightDisplay[k];
    c->stageBegin(ST_RESOLVE);
    resolve_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(rp, c->accVpl.p, c->accPhoton.p, c->accLight.p, n, c->resolveOut.p);
    c->stageEnd(ST_RESOLVE);
    c->launches++;
    return cudaGetLastError();
}

// emitted VPL / photon counts of the current record window (evplp_stats)
__global__ void count_flags_kernel(const EvplpRecord* __restrict__ records, uint64_t n, unsigned long long* out) {
    unsigned v = 0, p = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t f = records[i].flags;
        v += (f & EVPLP_FLAG_USABLE_VPL) ? 1u : 0u;
        p += (f & EVPLP_FLAG_USABLE_PHOTON) ? 1u : 0u;
    }
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); p += __shfl_xor_sync(0xffffffffu, p, o); }
    if ((threadIdx.x & 31) == 0) { if (v) atomicAdd(&out[0], (unsigned long long)v); if (p) atomicAdd(&out[1], (unsigned long long)p); }
}

cudaError_t launch_count_flags(EvplpContext* c, unsigned long long counts[2]) {
    cudaError_t e = c->scratch64.reserve(2);   // owned by the handle: no allocation per call
    if (e != cudaSuccess) return e;
    unsigned long long* d = c->scratch64.p;
    cudaMemsetAsync(d, 0, 16, c->stream);
    count_flags_kernel<<<148 * 8, 256, 0, c->stream>>>(c->records.p, c->numRecords, d);
    c->launches++;
    e = cudaMemcpyAsync(counts, d, 16, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e;
}

// ------------------------------------------------------------------ debug taps ----------
__global__ void trace_rays_kernel(DevScene sc, const float* __restrict__ rays, uint64_t n, int anyHit, int32_t* outPrim,
                                  float* outT, DevStats* stats) {
    __shared__ uint32_t stacks[4][BVH_STACK];
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const float* r = rays + (live ? i : 0) * 8;
    V3 o = v3(r[0], r[1], r[2]), d = v3(r[3], r[4], r[5]);
    int ovf = 0;
    if (anyHit == 2) {  // warp-cooperative any-hit (the gather's traversal)
        bool occ = trace_any_warp(sc, live, o, d, r[6], r[7], stacks[threadIdx.x >> 5], &ovf);
        if (live) { outPrim[i] = occ ? 1 : 0; if (outT) outT[i] = 0.f; }
    } else if (live) {
        if (anyHit == 1) {
            outPrim[i] = trace_any(sc, o, d, r[6], r[7], &ovf) ? 1 : 0;
            if (outT) outT[i] = 0.f;
        } else {
            RayHit h = (anyHit == 3) ? trace_closest<true>(sc, o, d, r[6], r[7], &ovf) : trace_closest<false>(sc, o, d, r[6], r[7], &ovf);
            outPrim[i] = h.prim;
            if (outT) outT[i] = h.prim >= 0 ? h.t : 0.f;
        }
    }
    if (ovf) stats->stackOverflow = 1;
}

cudaError_t launch_trace_rays(EvplpContext* c, const float* devRays, uint64_t n, int anyHit, int32_t* devPrim, float* devT) {
    if (n == 0) return cudaSuccess;
    trace_rays_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->scene(), devRays, n, anyHit, devPrim, devT, c->devStats.p);
    c->launches++;
    return cudaGetLastError();
}

__global__ void debug_uniforms_kernel(const uint32_t* __restrict__ skip, uint32_t seed, uint32_t n, float* out) {
    __shared__ uint32_t sm[kSkipMatrixWords];
    for (int k = threadIdx.x; k < kSkipMatrixWords; k += blockDim.x) sm[k] = skip[k];
    __syncthreads();
    if (threadIdx.x != 0) return;
    Xorwow rng = xorwow_seed(seed);
    xorwow_apply_matrix(rng, sm);
    for (uint32_t i = 0; i < n; i++) out[i] = xorwow_uniform(rng);
}

cudaError_t launch_debug_uniforms(EvplpContext* c, uint32_t seed, uint32_t n, float* devOut) {
    debug_uniforms_kernel<<<1, 128, 0, c->stream>>>(c->skipMatrix.p, seed, n, devOut);
    c->launches++;
    return cudaGetLastError();
}

// The real cuRAND device API, exactly as the reference calls it (lighttracing.cu:202-203):
// pins the oracle's and the product's XORWOW restatements against cuRAND itself.
__global__ void debug_curand_kernel(uint32_t seed, uint32_t subsequence, uint32_t n, float* out) {
    curandState st;
    curand_init(seed, subsequence, 0, &st);
    for (uint32_t i = 0; i < n; i++) out[i] = curand_uniform(&st);
}

cudaError_t launch_debug_curand(EvplpContext* c, uint32_t seed, uint32_t subsequence, uint32_t n, float* devOut) {
    debug_curand_kernel<<<1, 1, 0, c->stream>>>(seed, subsequence, n, devOut);
    c->launches++;
    return cudaGetLastError();
}

__global__ void debug_math_kernel(int op, const float* x, const float* y, uint32_t n, float* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (op) {
        case 0: out[i] = det_sinf(x[i]); break;
        case 1: out[i] = det_cosf(x[i]); break;
        case 2: out[i] = det_powf(x[i], y[i]); break;
        case 3: out[i] = det_asinf(x[i]); break;
        case 4: out[i] = det_sqrtf(x[i]); break;
        default: out[i] = 0.f;
    }
}

cudaError_t launch_debug_math(EvplpContext* c, int op, const float* x, const float* y, uint32_t n, float* out) {
    if (n == 0) return cudaSuccess;
    if (op == 5) return launch_debug_fast_pow(c, x, y, n, out);   // the cluster gather's pow approximation
    debug_math_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(op, x, y, n, out);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace evplp
