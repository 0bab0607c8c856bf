// ref_glue.inl -- TEST INFRASTRUCTURE.  Included AFTER the reference's OptiX program sources (which were included
// unmodified from /root/reference): binds the names those programs declare (ray, tHit, texcoord, geometryNormal,
// the three material samplers, lightIntensity, prdRadiance, prdShadow, rtMaterialClosestHit, rtMaterialAnyHit) and
// the geometry programs of triangleintersect.cu (namespace tri) into a brute-force rtTrace.  See ref_shim/optix_shim.h.
#pragma once
#include "../include/evplp.h"

namespace refshim {

struct DMesh { const float3* verts; const float2* uvs; const int3* idx; int numVerts, numTris, mat, firstPrim; };
struct DMat { const float4* tex[3]; int w[3], h[3]; float4 lightIntensity; };
struct DScene {
    const DMesh* meshes; int numMeshes;
    const DMat* mats; int numMats;
    const float* lightCdf; int lightTris;
    const float3* lightVerts; const uint3* lightIdx;
    float lightArea; float4 lightIntensity;
};

__device__ __forceinline__ void bind_material(const DMat& m) {
    lambertReflectanceTexture.texels = m.tex[0]; lambertReflectanceTexture.w = m.w[0]; lambertReflectanceTexture.h = m.h[0]; lambertReflectanceTexture.kind = 1;
    phongReflectanceTexture.texels = m.tex[1]; phongReflectanceTexture.w = m.w[1]; phongReflectanceTexture.h = m.h[1]; phongReflectanceTexture.kind = 1;
    phongExponentTexture.texels = m.tex[2]; phongExponentTexture.w = m.w[2]; phongExponentTexture.h = m.h[2]; phongExponentTexture.kind = 1;
    lightIntensity = m.lightIntensity;
}

__shared__ const DScene* g_scenePtr;

}  // namespace refshim

// rtReportIntersection: commit the candidate; on shadow rays (type 1) run the material's any-hit program
// (rtcomphoton.h:451: setAnyHitProgram(1, rtMaterialAnyHit)), which terminates the ray.
__device__ void rtReportIntersection(unsigned int) {
    using namespace refshim;
    g_trace.tmax = g_trace.tCandidate;
    g_trace.hit = 1;
    g_trace.hitMesh = g_trace.curMesh; g_trace.hitPrim = g_trace.curPrim;
    g_trace.hitNormal = tri::geometryNormal;
    g_trace.hitTexcoord = tri::texcoord;
    if (g_trace.anyHitRay) rtMaterialAnyHit();
}

__device__ void refshim_trace(const optix::Ray& r, void* prd) {
    using namespace refshim;
    const DScene& sc = *g_scenePtr;
    g_trace.tmin = r.tmin; g_trace.tmax = r.tmax; g_trace.hit = 0; g_trace.terminate = 0;
    g_trace.anyHitRay = (r.ray_type == 1) ? 1 : 0;
    if (g_trace.anyHitRay) prdShadow = *reinterpret_cast<PerRayData_shadow*>(prd);
    for (int m = 0; m < sc.numMeshes && !g_trace.terminate; m++) {
        const DMesh& mesh = sc.meshes[m];
        tri::vertexBuffer.data = const_cast<float3*>(mesh.verts); tri::vertexBuffer.count = mesh.numVerts;
        tri::indexBuffer.data = const_cast<int3*>(mesh.idx); tri::indexBuffer.count = mesh.numTris;
        tri::texcoordBuffer.data = const_cast<float2*>(mesh.uvs); tri::texcoordBuffer.count = mesh.numVerts;
        g_trace.curMesh = m;
        for (int p = 0; p < mesh.numTris && !g_trace.terminate; p++) {
            g_trace.curPrim = mesh.firstPrim + p;
            tri::ray = r;
            tri::ray.tmax = g_trace.tmax;  // rtCurrentRay's interval shrinks as closer hits are committed
            tri::meshFineIntersect(p);
        }
    }
    if (g_trace.anyHitRay) { *reinterpret_cast<PerRayData_shadow*>(prd) = prdShadow; return; }
    if (!g_trace.hit) return;  // no miss program is bound (rtcomphoton.h): the payload is left untouched
    bind_material(sc.mats[sc.meshes[g_trace.hitMesh].mat]);
    ray = r;
    tHit = g_trace.tmax;
    geometryNormal = g_trace.hitNormal;
    texcoord = g_trace.hitTexcoord;
    prdRadiance = *reinterpret_cast<PerRayData_radiance*>(prd);
    rtMaterialClosestHit();
    *reinterpret_cast<PerRayData_radiance*>(prd) = prdRadiance;
}

// ---- host side: scene upload shared by the harness files ---------------------------------------------------------
#include <vector>
#include <string>
#include <stdio.h>

namespace refshim {

struct HostScene {
    std::vector<void*> allocs;
    DScene h;        // host copy of the descriptor (device pointers inside)
    DScene* d = nullptr;
    int numPrims = 0;
    std::string err;
    template <class T> T* up(const T* src, size_t n) {
        T* p = nullptr;
        if (n == 0) n = 1;
        if (cudaMalloc((void**)&p, n * sizeof(T)) != cudaSuccess) { err = "cudaMalloc failed"; return nullptr; }
        allocs.push_back(p);
        if (src) cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice);
        return p;
    }
    ~HostScene() { for (void* p : allocs) cudaFree(p); }
};

// Same scene description as evplp_upload_scene (include/evplp.h).  The light CDF is built exactly like
// RtAreaLight::createOptixCdf (rtcommon.h:501-531): sequential float sums of Triangle::ComputeArea, normalised.
static HostScene* make_scene(const EvplpMeshDesc* meshes, int numMeshes, const EvplpMaterialDesc* mats, int numMats, int lightMesh,
                             const float lightPre[4], const float* lightCdf, float lightArea) {
    HostScene* s = new HostScene();
    std::vector<DMesh> dm(numMeshes);
    int first = 0;
    for (int m = 0; m < numMeshes; m++) {
        const EvplpMeshDesc& e = meshes[m];
        std::vector<float2> uv(e.numVertices, make_float2(0.f, 0.f));
        if (e.texcoords) for (int v = 0; v < e.numVertices; v++) uv[v] = make_float2(e.texcoords[2 * v], e.texcoords[2 * v + 1]);
        dm[m].verts = s->up(reinterpret_cast<const float3*>(e.vertices), e.numVertices);
        dm[m].uvs = s->up(uv.data(), e.numVertices);
        dm[m].idx = s->up(reinterpret_cast<const int3*>(e.indices), e.numTriangles);
        dm[m].numVerts = e.numVertices; dm[m].numTris = e.numTriangles; dm[m].mat = e.matIndex; dm[m].firstPrim = first;
        first += e.numTriangles;
    }
    s->numPrims = first;
    std::vector<DMat> mt(numMats);
    for (int k = 0; k < numMats; k++) {
        const EvplpMaterialDesc& e = mats[k];
        const float* src[3] = {e.lambertReflectance, e.phongReflectance, e.phongExponent};
        const int w[3] = {e.lambertW, e.phongW, e.exponentW}, h[3] = {e.lambertH, e.phongH, e.exponentH};
        for (int t = 0; t < 3; t++) {
            mt[k].tex[t] = s->up(reinterpret_cast<const float4*>(src[t]), (size_t)w[t] * h[t]);
            mt[k].w[t] = w[t]; mt[k].h[t] = h[t];
        }
        mt[k].lightIntensity = make_float4(e.lightIntensity[0], e.lightIntensity[1], e.lightIntensity[2], e.lightIntensity[3]);
    }
    s->h.meshes = s->up(dm.data(), numMeshes); s->h.numMeshes = numMeshes;
    s->h.mats = s->up(mt.data(), numMats); s->h.numMats = numMats;
    const EvplpMeshDesc& L = meshes[lightMesh];
    s->h.lightTris = L.numTriangles;
    s->h.lightCdf = s->up(lightCdf, L.numTriangles);
    s->h.lightVerts = dm[lightMesh].verts;
    s->h.lightIdx = reinterpret_cast<const uint3*>(dm[lightMesh].idx);
    s->h.lightArea = lightArea;
    s->h.lightIntensity = make_float4(lightPre[0], lightPre[1], lightPre[2], lightPre[3]);
    s->d = s->up(&s->h, 1);
    return s;
}

__device__ __forceinline__ void bind_scene(const DScene* sc) {
    g_scenePtr = sc;
    areaLightCdf.data = const_cast<float*>(sc->lightCdf); areaLightCdf.count = sc->lightTris;
    areaLightVertices.data = const_cast<float3*>(sc->lightVerts); areaLightVertices.count = 0;
    areaLightIndices.data = const_cast<uint3*>(sc->lightIdx); areaLightIndices.count = sc->lightTris;
    areaLightArea = sc->lightArea;
    areaLightIntensity = sc->lightIntensity;
    topObject = 0;
}

template <class S>
__device__ __forceinline__ void bind_gbuffer_plane(S& smp, const float4* plane, int W, int H) {
    smp.texels = plane; smp.w = W; smp.h = H; smp.kind = 0;
}

}  // namespace refshim
