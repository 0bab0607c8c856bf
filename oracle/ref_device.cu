// ref_device.cu -- TEST INFRASTRUCTURE: the reference's OWN OptiX device programs, compiled unmodified from
// /root/reference (never copied into this repository) against the stand-in headers of ref_shim/, plus a C ABI
// that launches them.  `make -C oracle refdevice` -> oracle/_ref/libref_device.so (git-ignored; ships to the GPU box).
// Used ONLY to pin the oracle (tests/test_gpu_reference.py, scripts/make_ref_goldens.py): the oracle's record
// flags / RNG consumption / counts must equal the reference's, its floats must agree within 1e-5 relative
// (the reference build contracts FMAs and calls libdevice powf / sinf / cosf; the oracle rounds every operation).
#include "ref_shim/optix_shim.h"
namespace tri {
#include "triangleintersect.cu"   // reference: realtimetechniques/triangleintersect.cu (meshFineIntersect, meshBound)
}
#include "lighttracing.cu"        // reference: realtimetechniques/lighttracing.cu (+ rtmath / rtlightsource / rtmaterial .cuh)
#include "ref_glue.inl"

using namespace refshim;

struct RefGatherArgs {
    float cameraPosition[3];
    uint32_t misMode;
    float pdfMc, clampingValue;
    uint32_t numVplLightPaths, numPhotonsPerLightPath, doAccumulate, rngSeed;
    float vslRadius, vslInvPiRadius2;
    int32_t W, H, x0, y0, x1, y1;
};

__global__ void k_trace_photons(const DScene* sc, RtPhotonRecord* recs, uint32_t seed, uint32_t firstPath, uint32_t numPaths, uint32_t B1) {
    bind_scene(sc);
    launchIndex = make_uint2(firstPath + blockIdx.x, 0u);
    launchDimension = make_uint2(0x7fffffffu, 1u);      // launchId = launchIndex.y * launchDimension.x + launchIndex.x = the path id
    rngSeed = seed; numPhotonsPerLightPath = B1;
    photons.data = recs - (size_t)firstPath * B1;          // record slot (path - firstPath) * B1 + bounce, as evplp_light_trace
    photons.count = (size_t)(firstPath + numPaths) * B1;
    tracePhotons();
}

__device__ __forceinline__ void bind_gather(const DScene* sc, const RefGatherArgs& a, const float4* planes, RtPhotonRecord* recs, float4* out) {
    bind_scene(sc);
    const size_t n = (size_t)a.W * a.H;
    bind_gbuffer_plane(deferredPositionTexture, planes, a.W, a.H);
    bind_gbuffer_plane(deferredNormalTexture, planes + n, a.W, a.H);
    bind_gbuffer_plane(deferredDiffuseTexture, planes + 2 * n, a.W, a.H);
    bind_gbuffer_plane(deferredPhongReflectanceTexture, planes + 3 * n, a.W, a.H);
    bind_gbuffer_plane(deferredPhongExpostureTexture, planes + 3 * n, a.W, a.H);
    launchIndex = make_uint2(a.x0 + blockIdx.x, a.y0 + blockIdx.y);
    launchDimension = make_uint2((unsigned)a.W, (unsigned)a.H);
    cameraPosition = make_float3(a.cameraPosition[0], a.cameraPosition[1], a.cameraPosition[2]);
    doAccumulate = a.doAccumulate; pdfMc = a.pdfMc; misMode = a.misMode; clampingValue = a.clampingValue;
    vslRadius = a.vslRadius; vslInvPiRadius2 = a.vslInvPiRadius2; rngSeed = a.rngSeed;
    numVplLightPaths = a.numVplLightPaths; numPhotonsPerLightPath = a.numPhotonsPerLightPath;
    photons.data = recs; photons.count = (size_t)a.numVplLightPaths * a.numPhotonsPerLightPath;
    outputBuffer.data = out; outputBuffer.count = (size_t)a.W;
}

__global__ void k_splat_color(const DScene* sc, RefGatherArgs a, const float4* planes, RtPhotonRecord* recs, float4* out) {
    bind_gather(sc, a, planes, recs, out);
    splatColor();
}
__global__ void k_splat_splotch(const DScene* sc, RefGatherArgs a, const float4* planes, RtPhotonRecord* recs, float4* out) {
    bind_gather(sc, a, planes, recs, out);
    splatSplotch();
}

// closest / any hit of arbitrary rays through the reference's meshFineIntersect (ties: smallest primitive id)
__global__ void k_trace_rays(const DScene* sc, const float* rays, int anyHit, int32_t* outPrim, float* outT) {
    bind_scene(sc);
    const float* r = rays + (size_t)blockIdx.x * 8;
    optix::Ray q(make_float3(r[0], r[1], r[2]), make_float3(r[3], r[4], r[5]), anyHit ? 1u : 0u, r[6], r[7]);
    // bypass the material programs: walk the triangles exactly like refshim_trace does
    g_trace.tmin = q.tmin; g_trace.tmax = q.tmax; g_trace.hit = 0; g_trace.terminate = 0; g_trace.anyHitRay = anyHit ? 1 : 0;
    prdShadow.hit = false;
    for (int m = 0; m < sc->numMeshes && !g_trace.terminate; m++) {
        const DMesh& mesh = sc->meshes[m];
        tri::vertexBuffer.data = const_cast<float3*>(mesh.verts); tri::indexBuffer.data = const_cast<int3*>(mesh.idx);
        tri::texcoordBuffer.data = const_cast<float2*>(mesh.uvs);
        g_trace.curMesh = m;
        for (int p = 0; p < mesh.numTris && !g_trace.terminate; p++) {
            g_trace.curPrim = mesh.firstPrim + p;
            tri::ray = q; tri::ray.tmax = g_trace.tmax;
            tri::meshFineIntersect(p);
        }
    }
    if (anyHit) { outPrim[blockIdx.x] = g_trace.hit ? 1 : 0; outT[blockIdx.x] = 0.f; }
    else { outPrim[blockIdx.x] = g_trace.hit ? g_trace.hitPrim : -1; outT[blockIdx.x] = g_trace.hit ? g_trace.tmax : 0.f; }
}

// BRDF library taps (rtmaterial.cuh:25-155): in = 16 floats per item, out = 8 floats per item
//   in: [0..2] a  [3..5] b  [6..8] c  [9..11] refl  [12] exponent  [13] seed  [14] subsequence
__global__ void k_brdf(int op, const float* in, float* out) {
    const float* q = in + (size_t)blockIdx.x * 16;
    float* o = out + (size_t)blockIdx.x * 8;
    const float3 a = make_float3(q[0], q[1], q[2]), b = make_float3(q[3], q[4], q[5]), c = make_float3(q[6], q[7], q[8]);
    const float3 refl = make_float3(q[9], q[10], q[11]);
    const float e = q[12];
    for (int k = 0; k < 8; k++) o[k] = 0.f;
    curandState st;
    curand_init((unsigned)q[13], (unsigned)q[14], 0, &st);
    float3 dir; float pdf;
    switch (op) {
        case 0: { float3 r = LambertSample(&dir, &pdf, a, b, refl, &st); o[0] = dir.x; o[1] = dir.y; o[2] = dir.z; o[3] = pdf; o[4] = r.x; o[5] = r.y; o[6] = r.z; o[7] = curand_uniform(&st); break; }
        case 1: { float3 r = PhongSample(&dir, &pdf, a, b, refl, e, &st); o[0] = dir.x; o[1] = dir.y; o[2] = dir.z; o[3] = pdf; o[4] = r.x; o[5] = r.y; o[6] = r.z; o[7] = curand_uniform(&st); break; }
        case 2: o[0] = LambertPdfA(a, b, c); o[1] = LambertPdfW(a, c); o[2] = GeometryTerm(a, b, c); break;
        case 3: o[0] = PhongPdfA(a, b, c, refl, make_float3(q[15]), e); o[1] = PhongPdfW(a, c, refl, make_float3(q[15]), e); break;
        case 4: { o[0] = PhongEvalF(a, b, c, e); float3 r = PhongEval(a, b, c, refl, e); o[1] = r.x; o[2] = r.y; o[3] = r.z; break; }
        case 5: { float be, ga; SquareToBarycentric(&be, &ga, q[0], q[1]); o[0] = be; o[1] = ga; float3 s = SquareToSolidAngle(q[0], q[1], q[2]); o[2] = s.x; o[3] = s.y; o[4] = s.z; o[5] = russianProb(b); break; }
    }
}

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }
const char* ref_sources(void) { return "reflectcuts/realtimetechniques/{lighttracing.cu,triangleintersect.cu,rtmaterial.cuh,rtmath.cuh,rtlightsource.cuh,all.cuh,rtcomphoton/rtphotonrecord.h} (unmodified, nvcc sm_100a)"; }

void* ref_scene_create(const EvplpMeshDesc* meshes, int numMeshes, const EvplpMaterialDesc* mats, int numMats, int lightMesh,
                       const float lightPre[4], const float* lightCdf, float lightArea) {
    return make_scene(meshes, numMeshes, mats, numMats, lightMesh, lightPre, lightCdf, lightArea);
}
void ref_scene_destroy(void* s) { delete static_cast<HostScene*>(s); }

int ref_trace_photons(void* sv, uint32_t rngSeed, uint32_t firstPath, uint32_t numPaths, uint32_t B1, EvplpRecord* out) {
    static_assert(sizeof(RtPhotonRecord) == sizeof(EvplpRecord), "record layout");
    HostScene* s = static_cast<HostScene*>(sv);
    RtPhotonRecord* d = nullptr;
    const size_t n = (size_t)numPaths * B1;
    CK(cudaMalloc((void**)&d, n * sizeof(RtPhotonRecord)));
    CK(cudaMemset(d, 0, n * sizeof(RtPhotonRecord)));   // the reference clears only mFlags (lighttracing.cu:197-200)
    k_trace_photons<<<numPaths, 1>>>(s->d, d, rngSeed, firstPath, numPaths, B1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, d, n * sizeof(RtPhotonRecord), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

// planes: 4 x W*H float4 (position+w, normal, Kd, Ks+exponent); records: the VPL prefix; out: W*H float4, read when doAccumulate
int ref_gather(void* sv, int vsl, const RefGatherArgs* a, const float* planes, const EvplpRecord* records, float* out) {
    HostScene* s = static_cast<HostScene*>(sv);
    const size_t n = (size_t)a->W * a->H, nr = (size_t)a->numVplLightPaths * a->numPhotonsPerLightPath;
    float4 *dp = nullptr, *dout = nullptr;
    RtPhotonRecord* dr = nullptr;
    CK(cudaMalloc((void**)&dp, 4 * n * sizeof(float4)));
    CK(cudaMalloc((void**)&dout, n * sizeof(float4)));
    CK(cudaMalloc((void**)&dr, (nr ? nr : 1) * sizeof(RtPhotonRecord)));
    CK(cudaMemcpy(dp, planes, 4 * n * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dout, out, n * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, records, nr * sizeof(RtPhotonRecord), cudaMemcpyHostToDevice));
    dim3 grid(a->x1 - a->x0, a->y1 - a->y0);
    if (vsl) k_splat_splotch<<<grid, 1>>>(s->d, *a, dp, dr, dout);
    else k_splat_color<<<grid, 1>>>(s->d, *a, dp, dr, dout);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dout, n * sizeof(float4), cudaMemcpyDeviceToHost));
    cudaFree(dp); cudaFree(dout); cudaFree(dr);
    return 0;
}

int ref_trace_rays(void* sv, const float* rays, uint32_t n, int anyHit, int32_t* outPrim, float* outT) {
    HostScene* s = static_cast<HostScene*>(sv);
    float *dr = nullptr, *dt = nullptr; int32_t* dprim = nullptr;
    CK(cudaMalloc((void**)&dr, (size_t)n * 8 * sizeof(float)));
    CK(cudaMalloc((void**)&dt, (size_t)n * sizeof(float)));
    CK(cudaMalloc((void**)&dprim, (size_t)n * sizeof(int32_t)));
    CK(cudaMemcpy(dr, rays, (size_t)n * 8 * sizeof(float), cudaMemcpyHostToDevice));
    k_trace_rays<<<n, 1>>>(s->d, dr, anyHit, dprim, dt);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(outPrim, dprim, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(outT, dt, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(dr); cudaFree(dt); cudaFree(dprim);
    return 0;
}

int ref_brdf(int op, const float* in16, uint32_t n, float* out8) {
    float *di = nullptr, *dout = nullptr;
    CK(cudaMalloc((void**)&di, (size_t)n * 16 * sizeof(float)));
    CK(cudaMalloc((void**)&dout, (size_t)n * 8 * sizeof(float)));
    CK(cudaMemcpy(di, in16, (size_t)n * 16 * sizeof(float), cudaMemcpyHostToDevice));
    k_brdf<<<n, 1>>>(op, di, dout);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out8, dout, (size_t)n * 8 * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(di); cudaFree(dout);
    return 0;
}

}  // extern "C"
