// ref_device_pt.cu -- TEST INFRASTRUCTURE: the reference's RtPt2 device programs (realtimetechniques/pathtracing.cu:
// splatColor, pathTraceSimple, rtMaterialClosestHit, rtMaterialAnyHit), compiled unmodified from /root/reference against
// the stand-in OptiX headers of ref_shim/ -- a separate translation unit because pathtracing.cu and lighttracing.cu define
// the same program names.  `make -C oracle refdevice` -> oracle/_ref/libref_device_pt.so.  Pins the oracle's namespace pt.
#include "ref_shim/optix_shim.h"
namespace tri {
#include "triangleintersect.cu"   // reference: realtimetechniques/triangleintersect.cu
}
#include "pathtracing.cu"         // reference: realtimetechniques/pathtracing.cu
#include "ref_glue.inl"

using namespace refshim;

struct RefPtArgs {
    float cameraPosition[3];
    uint32_t doAccumulate, maxBounces, rngSeed;
    int32_t W, H, x0, y0, x1, y1;
};

__global__ void k_path_trace(const DScene* sc, RefPtArgs a, const float4* planes, float4* out) {
    bind_scene(sc);
    const size_t n = (size_t)a.W * a.H;
    bind_gbuffer_plane(deferredPositionTexture, planes, a.W, a.H);
    bind_gbuffer_plane(deferredNormalTexture, planes + n, a.W, a.H);
    bind_gbuffer_plane(deferredDiffuseTexture, planes + 2 * n, a.W, a.H);
    bind_gbuffer_plane(deferredPhongReflectanceTexture, planes + 3 * n, a.W, a.H);
    launchIndex = make_uint2(a.x0 + blockIdx.x, a.y0 + blockIdx.y);
    launchDimension = make_uint2((unsigned)a.W, (unsigned)a.H);
    cameraPosition = make_float3(a.cameraPosition[0], a.cameraPosition[1], a.cameraPosition[2]);
    doAccumulate = a.doAccumulate; maxBounces = a.maxBounces; rngSeed = a.rngSeed;
    outputBuffer.data = out; outputBuffer.count = (size_t)a.W;
    splatColor();
}

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" {

const char* refpt_last_error(void) { return g_err.c_str(); }
const char* refpt_sources(void) { return "reflectcuts/realtimetechniques/{pathtracing.cu,triangleintersect.cu,rtmaterial.cuh,rtmath.cuh,rtlightsource.cuh} (unmodified, nvcc sm_100a)"; }

void* refpt_scene_create(const EvplpMeshDesc* meshes, int numMeshes, const EvplpMaterialDesc* mats, int numMats, int lightMesh,
                         const float lightPre[4], const float* lightCdf, float lightArea) {
    return make_scene(meshes, numMeshes, mats, numMats, lightMesh, lightPre, lightCdf, lightArea);
}
void refpt_scene_destroy(void* s) { delete static_cast<HostScene*>(s); }

int refpt_path_trace(void* sv, const RefPtArgs* a, const float* planes, float* out) {
    HostScene* s = static_cast<HostScene*>(sv);
    const size_t n = (size_t)a->W * a->H;
    float4 *dp = nullptr, *dout = nullptr;
    CK(cudaMalloc((void**)&dp, 4 * n * sizeof(float4)));
    CK(cudaMalloc((void**)&dout, n * sizeof(float4)));
    CK(cudaMemcpy(dp, planes, 4 * n * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dout, out, n * sizeof(float4), cudaMemcpyHostToDevice));
    k_path_trace<<<dim3(a->x1 - a->x0, a->y1 - a->y0), 1>>>(s->d, *a, dp, dout);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dout, n * sizeof(float4), cudaMemcpyDeviceToHost));
    cudaFree(dp); cudaFree(dout);
    return 0;
}

}  // extern "C"
