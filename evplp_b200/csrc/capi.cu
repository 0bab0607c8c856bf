// capi.cu -- the C ABI of include/evplp.h: context management, uploads, stage dispatch,
// parity taps.  Every compute entry point launches sm_100a kernels; there is no CPU path.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>
#include "context.h"

using namespace evplp;

// defaults a new handle starts from (evplp_set_option with a NULL handle); guarded: handles may be created from several threads
static evplp::Options g_defaultOptions;
static std::mutex g_defaultsMutex;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(EVPLP_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)
#define NEED(cond, msg) do { if (!(cond)) return fail(EVPLP_ERR_INVALID, msg); } while (0)

DevScene EvplpContext::scene() const {
    DevScene s;
    s.triLeaf = triLeaf.p; s.triVerts = triVerts.p; s.triUV = triUV.p; s.mats = mats.p; s.texPool = texPool.p;
    s.lightCdf = lightCdf.p; s.nodes = nodes.p; s.cnodes = cnodes.p; s.numPrims = numPrims; s.numNodes = numNodes;
    s.shaftNodes = shaftNodes.p; s.numShaftNodes = numShaftNodes;
    s.lightFirst = lightFirst; s.lightCount = lightCount; s.lightArea = lightArea;
    for (int k = 0; k < 4; k++) { s.lightIntensity[k] = lightIntensity[k]; s.lightDisplay[k] = lightDisplay[k]; }
    return s;
}

extern "C" {

const char* evplp_last_error(void) { return g_err.c_str(); }
const char* evplp_version(void) { return "evplp-b200 0.1 (sm_100a)"; }

int evplp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // extern "C"

static int check_overflow(EvplpContext* c) {
    DevStats ds;
    CU(cudaMemcpyAsync(&ds, c->devStats.p, sizeof(ds), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (ds.stackOverflow) return fail(EVPLP_ERR_CUDA, "BVH traversal stack overflow (tree too deep for BVH_STACK)");
    return EVPLP_OK;
}

extern "C" {

int evplp_create(int device, int width, int height, evplp_handle* out) {
    NEED(out != nullptr, "evplp_create: out == NULL");
    NEED(width > 0 && height > 0 && (uint64_t)width * height < (1ull << 31), "evplp_create: bad resolution");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(EVPLP_ERR_NO_DEVICE, "evplp_create: no CUDA device (this library has no CPU fallback)");
    }
    NEED(device >= 0 && device < n, "evplp_create: device index out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(EVPLP_ERR_NO_DEVICE, "evplp_create: device is not sm_100 class (Blackwell B200 required)");
    CU(cudaSetDevice(device));
    EvplpContext* c = new EvplpContext();
    { std::lock_guard<std::mutex> lock(g_defaultsMutex); c->opt = g_defaultOptions; }
    c->device = device; c->W = width; c->H = height;
    // every failure below destroys the partially built context (stream, events, buffers) before returning
    struct Guard { EvplpContext* c; ~Guard() { if (c) evplp_destroy(c); } } guard{c};
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int s = 0; s < ST_COUNT; s++) {
        CU(cudaEventCreate(&c->stageA[s]));
        CU(cudaEventCreate(&c->stageB[s]));
    }
    for (int s = 0; s < 4; s++) CU(cudaEventCreate(&c->userEv[s]));
    const size_t n_px = (size_t)width * height;
    CU(c->gbuf.reserve(4 * n_px));
    CU(c->gprim.reserve(n_px));
    CU(c->accVpl.reserve(3 * n_px));
    CU(c->accPhoton.reserve(3 * n_px));
    CU(c->accLight.reserve(n_px));
    CU(c->accCount.reserve(2));
    CU(cudaMemsetAsync(c->accCount.p, 0, 2 * sizeof(long long), c->stream));
    CU(c->resolveOut.reserve(3 * n_px));
    CU(c->devStats.reserve(1));
    CU(c->skipMatrix.reserve(kSkipMatrixWords));
    CU(c->skipTable.reserve(kSkipTableWords));
    CU(c->counters.reserve(4));
    CU(cudaMemsetAsync(c->devStats.p, 0, sizeof(DevStats), c->stream));
    CU(cudaMemsetAsync(c->accVpl.p, 0, sizeof(long long) * 3 * n_px, c->stream));
    CU(cudaMemsetAsync(c->accPhoton.p, 0, sizeof(long long) * 3 * n_px, c->stream));
    CU(cudaMemsetAsync(c->accLight.p, 0, sizeof(uint32_t) * n_px, c->stream));
    CU(cudaMemsetAsync(c->gprim.p, 0xff, sizeof(int32_t) * n_px, c->stream));
    CU(cudaMemsetAsync(c->gbuf.p, 0, sizeof(float4) * 4 * n_px, c->stream));
    CU(cudaMallocHost((void**)&c->resolvePinned, sizeof(float) * 3 * n_px));
    memset(&c->stats, 0, sizeof(c->stats));
    memset(&c->params, 0, sizeof(c->params));
    CU(cudaStreamSynchronize(c->stream));
    guard.c = nullptr;
    *out = c;
    return EVPLP_OK;
}

int evplp_destroy(evplp_handle c) {
    NEED(c != nullptr, "evplp_destroy: NULL handle");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->triVerts.release(); c->triLeaf.release(); c->texPool.release(); c->triUV.release(); c->mats.release();
    c->lightCdf.release(); c->primLo.release(); c->primHi.release(); c->codes.release(); c->codesSorted.release();
    c->primIds.release(); c->primIdsSorted.release(); c->left.release(); c->right.release(); c->parent.release();
    c->leafParent.release(); c->rangeFirst.release(); c->rangeLast.release(); c->nodeBounds.release();
    c->refitFlags.release(); c->nodes.release(); c->cnodes.release(); c->shaftNodes.release(); c->sceneBoundsEnc.release(); c->sortTemp.release();
    c->gatherCost.release(); c->gatherCostSorted.release(); c->gatherIota.release(); c->gatherOrder.release();
    c->queueA.release(); c->queueB.release(); c->counters.release(); c->skipMatrix.release(); c->skipTable.release(); c->scratch64.release(); c->records.release();
    c->vplList.release(); c->vplKeys.release(); c->vplKeysSorted.release(); c->vplVals.release(); c->vplOrder.release(); c->vplPrepared.release(); c->clusterBox.release(); c->clusterSlots.release(); c->clusterList.release(); c->photonList.release(); c->splatPrep.release(); c->tileCount.release(); c->tileOffset.release(); c->tileCursor.release(); c->tileList.release(); c->gbuf.release(); c->gprim.release(); c->accVpl.release();
    c->accPhoton.release(); c->accLight.release(); c->accCount.release(); c->resolveOut.release(); c->devStats.release();
    if (c->resolvePinned) cudaFreeHost(c->resolvePinned);
    for (int s = 0; s < ST_COUNT; s++) { cudaEventDestroy(c->stageA[s]); cudaEventDestroy(c->stageB[s]); }
    for (int s = 0; s < 4; s++) cudaEventDestroy(c->userEv[s]);
    cudaStreamDestroy(c->stream);
    delete c;
    return EVPLP_OK;
}

int evplp_upload_scene(evplp_handle c, const EvplpMeshDesc* meshes, int32_t numMeshes, const EvplpMaterialDesc* materials,
                       int32_t numMaterials, int32_t lightMeshIndex, const float lightIntensityPrecomputed[4],
                       const float lightIntensityDisplay[4]) {
    NEED(c != nullptr, "evplp_upload_scene: NULL handle");
    NEED(meshes && numMeshes > 0 && materials && numMaterials > 0, "evplp_upload_scene: empty scene");
    NEED(lightMeshIndex >= 0 && lightMeshIndex < numMeshes, "evplp_upload_scene: lightMeshIndex out of range");
    NEED(lightIntensityPrecomputed && lightIntensityDisplay, "evplp_upload_scene: NULL light intensity");
    CU(cudaSetDevice(c->device));
    size_t numPrims = 0;
    for (int m = 0; m < numMeshes; m++) {
        NEED(meshes[m].vertices && meshes[m].indices && meshes[m].numTriangles >= 0 && meshes[m].numVertices >= 0,
             "evplp_upload_scene: bad mesh");
        NEED(meshes[m].matIndex >= 0 && meshes[m].matIndex < numMaterials, "evplp_upload_scene: matIndex out of range");
        numPrims += (size_t)meshes[m].numTriangles;
    }
    NEED(numPrims < (1u << 27), "evplp_upload_scene: too many triangles");
    NEED(meshes[lightMeshIndex].numTriangles > 0, "evplp_upload_scene: the light mesh has no triangles");
    std::vector<float4> verts(3 * numPrims);
    std::vector<float2> uvs(3 * numPrims);
    size_t p = 0;
    int lightFirst = 0, lightCount = 0;
    for (int m = 0; m < numMeshes; m++) {
        const EvplpMeshDesc& d = meshes[m];
        if (m == lightMeshIndex) { lightFirst = (int)p; lightCount = d.numTriangles; }
        float matBits;
        int32_t mi = d.matIndex;
        memcpy(&matBits, &mi, 4);
        for (int t = 0; t < d.numTriangles; t++, p++) {
            for (int k = 0; k < 3; k++) {
                const int32_t vi = d.indices[3 * t + k];
                NEED(vi >= 0 && vi < d.numVertices, "evplp_upload_scene: vertex index out of range");
                verts[3 * p + k] = make_float4(d.vertices[3 * vi], d.vertices[3 * vi + 1], d.vertices[3 * vi + 2], matBits);
                uvs[3 * p + k] = d.texcoords ? make_float2(d.texcoords[2 * vi], d.texcoords[2 * vi + 1]) : make_float2(0.f, 0.f);
            }
        }
    }
    // materials: pool all textures into one float4 array
    std::vector<DevMaterial> dm(numMaterials);
    std::vector<float4> pool;
    auto add_tex = [&](const float* data, int w, int h, DevTexture& t) -> bool {
        if (!data || w <= 0 || h <= 0) return false;
        t.w = w; t.h = h; t.offset = (int)pool.size(); t.pad = 0;
        const size_t n = (size_t)w * h;
        pool.resize(pool.size() + n);
        memcpy(&pool[t.offset], data, n * sizeof(float4));
        return true;
    };
    for (int m = 0; m < numMaterials; m++) {
        const EvplpMaterialDesc& d = materials[m];
        NEED(add_tex(d.lambertReflectance, d.lambertW, d.lambertH, dm[m].lambert), "evplp_upload_scene: bad lambert texture");
        NEED(add_tex(d.phongReflectance, d.phongW, d.phongH, dm[m].phong), "evplp_upload_scene: bad phong texture");
        NEED(add_tex(d.phongExponent, d.exponentW, d.exponentH, dm[m].exponent), "evplp_upload_scene: bad exponent texture");
        memcpy(dm[m].lightIntensity, d.lightIntensity, 16);
    }
    // RtAreaLight::createOptixCdf (rtcommon.h:501-531), sequential f32 like the reference host code
    std::vector<float> cdf(lightCount);
    float sumArea = 0.f;
    for (int i = 0; i < lightCount; i++) {
        const float4 a = verts[3 * (size_t)(lightFirst + i)], b = verts[3 * (size_t)(lightFirst + i) + 1],
                     cc = verts[3 * (size_t)(lightFirst + i) + 2];
        V3 ab = v3(b.x, b.y, b.z) - v3(a.x, a.y, a.z), ac = v3(cc.x, cc.y, cc.z) - v3(a.x, a.y, a.z);
        V3 cr = cross(ab, ac);
        float area = sqrtf(dot(cr, cr)) / 2.0f;  // Triangle::ComputeArea, trianglemesh.cpp:13-19
        sumArea += area;
        cdf[i] = sumArea;
    }
    for (int i = 0; i < lightCount; i++) cdf[i] /= sumArea;

    CU(c->triVerts.reserve(verts.size())); CU(c->triUV.reserve(uvs.size()));
    CU(c->mats.reserve(dm.size())); CU(c->texPool.reserve(pool.size())); CU(c->lightCdf.reserve(cdf.size()));
    CU(cudaMemcpyAsync(c->triVerts.p, verts.data(), verts.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->triUV.p, uvs.data(), uvs.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->mats.p, dm.data(), dm.size() * sizeof(DevMaterial), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->texPool.p, pool.data(), pool.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->lightCdf.p, cdf.data(), cdf.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->numPrims = (int)numPrims; c->numMats = numMaterials;
    c->lightFirst = lightFirst; c->lightCount = lightCount; c->lightArea = sumArea;
    for (int k = 0; k < 3; k++) { c->lightBoxMin[k] = 3.0e38f; c->lightBoxMax[k] = -3.0e38f; }
    for (size_t v = 3 * (size_t)lightFirst; v < 3 * (size_t)(lightFirst + lightCount); v++) {
        const float q[3] = {verts[v].x, verts[v].y, verts[v].z};
        for (int k = 0; k < 3; k++) { c->lightBoxMin[k] = fminf(c->lightBoxMin[k], q[k]); c->lightBoxMax[k] = fmaxf(c->lightBoxMax[k], q[k]); }
    }
    memcpy(c->lightIntensity, lightIntensityPrecomputed, 16);
    memcpy(c->lightDisplay, lightIntensityDisplay, 16);
    c->sceneLoaded = true;
    c->bvhBuilt = false;
    c->numNodes = 0;
    return EVPLP_OK;
}

int evplp_build_bvh(evplp_handle c) {
    NEED(c != nullptr, "evplp_build_bvh: NULL handle");
    NEED(c->sceneLoaded, "evplp_build_bvh: no scene uploaded");
    CU(cudaSetDevice(c->device));
    std::string err;
    cudaError_t e = build_bvh_device(c, &err);
    c->stageEnd(ST_BVH);
    if (e != cudaSuccess) return fail(EVPLP_ERR_CUDA, "evplp_build_bvh: " + err);
    CU(cudaStreamSynchronize(c->stream));
    return EVPLP_OK;
}

int evplp_set_params(evplp_handle c, const EvplpParams* params) {
    NEED(c != nullptr && params != nullptr, "evplp_set_params: NULL argument");
    NEED(params->numPhotonsPerLightPath >= 1 && params->numPhotonsPerLightPath <= 64, "evplp_set_params: numPhotonsPerLightPath out of range");
    NEED(params->misMode <= 5, "evplp_set_params: misMode out of range");
    NEED(params->numVplLightPaths <= params->numLightPaths || params->numLightPaths == 0, "evplp_set_params: numVplLightPaths > numLightPaths");
    c->params = *params;
    c->paramsSet = true;
    return EVPLP_OK;
}

static int ensure_skip_matrix(EvplpContext* c, uint32_t subsequence) {
    if (c->skipMatrixValid && c->skipMatrixSeed == subsequence) return EVPLP_OK;
    static thread_local uint32_t host[kSkipMatrixWords];
    xorwow_compose_skip(subsequence, host);
    static thread_local uint32_t hostTab[kSkipTableWords];
    xorwow_build_tables(host, hostTab);
    CU(cudaMemcpyAsync(c->skipMatrix.p, host, sizeof(host), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->skipTable.p, hostTab, sizeof(hostTab), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // `host` is reused by the next call
    c->skipMatrixSeed = subsequence;
    c->skipMatrixValid = true;
    return EVPLP_OK;
}

int evplp_clear_accum(evplp_handle c) {
    NEED(c != nullptr, "evplp_clear_accum: NULL handle");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->W * c->H;
    CU(cudaMemsetAsync(c->accVpl.p, 0, sizeof(long long) * 3 * n, c->stream));
    CU(cudaMemsetAsync(c->accPhoton.p, 0, sizeof(long long) * 3 * n, c->stream));
    CU(cudaMemsetAsync(c->accLight.p, 0, sizeof(uint32_t) * n, c->stream));
    CU(cudaMemsetAsync(c->accCount.p, 0, 2 * sizeof(long long), c->stream));
    return EVPLP_OK;
}

int evplp_add_iterations(evplp_handle c, int64_t n) {
    NEED(c != nullptr, "evplp_add_iterations: NULL handle");
    CU(cudaSetDevice(c->device));
    CU(launch_add_count(c, (long long)n));
    return EVPLP_OK;
}

int evplp_iterations(evplp_handle c, int64_t* n) {
    NEED(c != nullptr && n != nullptr, "evplp_iterations: NULL argument");
    CU(cudaSetDevice(c->device));
    long long v = 0;
    CU(cudaMemcpyAsync(&v, c->accCount.p, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *n = (int64_t)v;
    return EVPLP_OK;
}

int evplp_gbuffer(evplp_handle c) {
    NEED(c != nullptr, "evplp_gbuffer: NULL handle");
    NEED(c->bvhBuilt && c->paramsSet, "evplp_gbuffer: needs evplp_build_bvh and evplp_set_params first");
    CU(cudaSetDevice(c->device));
    CU(launch_gbuffer(c));
    c->gbufValid = true;
    return EVPLP_OK;
}

int evplp_light_trace(evplp_handle c, uint32_t rngSeed, uint32_t firstPath, uint32_t numPaths) {
    NEED(c != nullptr, "evplp_light_trace: NULL handle");
    NEED(c->bvhBuilt && c->paramsSet, "evplp_light_trace: needs evplp_build_bvh and evplp_set_params first");
    CU(cudaSetDevice(c->device));
    const uint64_t nrec = (uint64_t)numPaths * c->params.numPhotonsPerLightPath;
    NEED(nrec < (1ull << 32), "evplp_light_trace: more than 2^32 records in one call; trace in chunks");
    CU(c->records.reserve(nrec));
    int rc = ensure_skip_matrix(c, rngSeed);
    if (rc) return rc;
    CU(launch_light_trace(c, rngSeed, firstPath, numPaths));
    c->numRecords = nrec;
    c->recordsFirstPath = firstPath;
    return EVPLP_OK;
}

static int tile_of(EvplpContext* c, const EvplpTile* tile, EvplpTile* out) {
    if (!tile) { out->x0 = 0; out->y0 = 0; out->x1 = c->W; out->y1 = c->H; return EVPLP_OK; }
    NEED(tile->x0 >= 0 && tile->y0 >= 0 && tile->x1 <= c->W && tile->y1 <= c->H && tile->x0 <= tile->x1 && tile->y0 <= tile->y1,
         "tile outside the image");
    *out = *tile;
    return EVPLP_OK;
}

int evplp_vpl_gather(evplp_handle c, const EvplpTile* tile, int gatherMode) {
    NEED(c != nullptr, "evplp_vpl_gather: NULL handle");
    NEED(c->bvhBuilt && c->paramsSet, "evplp_vpl_gather: needs evplp_build_bvh and evplp_set_params first");
    NEED(gatherMode >= 0 && gatherMode <= 2, "evplp_vpl_gather: bad gather mode");
    NEED(c->params.numVplLightPaths > 0, "evplp_vpl_gather: numVplLightPaths == 0 (the reference disables the gather, rtcomphoton.h:200-203)");
    const uint64_t need = gatherMode == EVPLP_GATHER_LVC ? (uint64_t)c->params.numLightPaths * c->params.numPhotonsPerLightPath
                                                         : (uint64_t)c->params.numVplLightPaths * c->params.numPhotonsPerLightPath;
    NEED(c->numRecords >= need, "evplp_vpl_gather: record buffer holds fewer records than the gather reads");
    CU(cudaSetDevice(c->device));
    EvplpTile t;
    int rc = tile_of(c, tile, &t);
    if (rc) return rc;
    if (gatherMode != EVPLP_GATHER_VPL) { rc = ensure_skip_matrix(c, c->params.rngSeed); if (rc) return rc; }
    CU(launch_gather(c, t, gatherMode));
    return EVPLP_OK;
}

int evplp_path_trace(evplp_handle c, const EvplpTile* tile, uint32_t maxBounces) {
    NEED(c != nullptr, "evplp_path_trace: NULL handle");
    NEED(c->bvhBuilt && c->paramsSet && c->gbufValid, "evplp_path_trace: needs evplp_build_bvh, evplp_set_params and a G-buffer first");
    CU(cudaSetDevice(c->device));
    EvplpTile t;
    int rc = tile_of(c, tile, &t);
    if (rc) return rc;
    rc = ensure_skip_matrix(c, c->params.rngSeed);
    if (rc) return rc;
    CU(launch_path_trace(c, t, maxBounces));
    return EVPLP_OK;
}

int evplp_photon_splat(evplp_handle c, uint64_t firstRecord, uint64_t numRecords, const EvplpTile* tile) {
    NEED(c != nullptr, "evplp_photon_splat: NULL handle");
    NEED(c->paramsSet && c->gbufValid, "evplp_photon_splat: needs evplp_set_params and a G-buffer first");
    NEED(firstRecord + numRecords <= c->numRecords, "evplp_photon_splat: record window outside the traced records");
    CU(cudaSetDevice(c->device));
    EvplpTile t;
    int rc = tile_of(c, tile, &t);
    if (rc) return rc;
    CU(launch_splat(c, firstRecord, numRecords, t));
    return EVPLP_OK;
}

int evplp_light_pass(evplp_handle c) {
    NEED(c != nullptr, "evplp_light_pass: NULL handle");
    NEED(c->bvhBuilt && c->paramsSet, "evplp_light_pass: needs evplp_build_bvh and evplp_set_params first");
    CU(cudaSetDevice(c->device));
    CU(launch_light_pass(c));
    return EVPLP_OK;
}

// NCCL is resolved at run time: the symbols of an NCCL already loaded into the process
// (e.g. the one bundled with torch) win, else libnccl.so.2 is opened.
typedef int (*ncclAllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static ncclAllReduce_t find_nccl() {
    static ncclAllReduce_t fn = nullptr;
    if (fn) return fn;
    fn = (ncclAllReduce_t)dlsym(RTLD_DEFAULT, "ncclAllReduce");
    if (!fn) {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (lib) fn = (ncclAllReduce_t)dlsym(lib, "ncclAllReduce");
    }
    return fn;
}

int evplp_reduce(evplp_handle c, void* ncclComm) {
    NEED(c != nullptr && ncclComm != nullptr, "evplp_reduce: NULL argument");
    CU(cudaSetDevice(c->device));
    ncclAllReduce_t allReduce = find_nccl();
    if (!allReduce) return fail(EVPLP_ERR_NCCL, "evplp_reduce: ncclAllReduce not found (no NCCL in the process and libnccl.so.2 not loadable)");
    const size_t n = (size_t)c->W * c->H;
    const int ncclInt64 = 4, ncclUint32 = 3, ncclSum = 0;  // nccl.h: ncclInt64 = 4, ncclUint32 = 3, ncclSum = 0
    int r = allReduce(c->accVpl.p, c->accVpl.p, 3 * n, ncclInt64, ncclSum, ncclComm, c->stream);
    if (r == 0) r = allReduce(c->accPhoton.p, c->accPhoton.p, 3 * n, ncclInt64, ncclSum, ncclComm, c->stream);
    if (r == 0) r = allReduce(c->accLight.p, c->accLight.p, n, ncclUint32, ncclSum, ncclComm, c->stream);
    if (r == 0) r = allReduce(c->accCount.p, c->accCount.p, 1, ncclInt64, ncclSum, ncclComm, c->stream);
    if (r != 0) return fail(EVPLP_ERR_NCCL, "evplp_reduce: ncclAllReduce failed with code " + std::to_string(r));
    return EVPLP_OK;
}

int evplp_accum_layer(evplp_handle c, int layer, void** devPtr, uint64_t* numElems) {
    NEED(c != nullptr && devPtr && numElems, "evplp_accum_layer: NULL argument");
    const uint64_t n = (uint64_t)c->W * c->H;
    switch (layer) {
        case 0: *devPtr = c->accVpl.p; *numElems = 3 * n; break;
        case 1: *devPtr = c->accPhoton.p; *numElems = 3 * n; break;
        case 2: *devPtr = c->accLight.p; *numElems = n; break;
        case 3: *devPtr = c->accCount.p; *numElems = 1; break;
        default: return fail(EVPLP_ERR_INVALID, "evplp_accum_layer: layer must be 0, 1, 2 or 3");
    }
    return EVPLP_OK;
}

int evplp_resolve(evplp_handle c, float vplScale, float photonScale, float lightScale, int doGammaCorrection, float* hostRGB) {
    NEED(c != nullptr && hostRGB != nullptr, "evplp_resolve: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(launch_resolve(c, vplScale, photonScale, lightScale, doGammaCorrection));
    const size_t bytes = sizeof(float) * 3 * (size_t)c->W * c->H;
    CU(cudaMemcpyAsync(c->resolvePinned, c->resolveOut.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(hostRGB, c->resolvePinned, bytes);
    return check_overflow(c);
}

// ---------------------------------------------------------------- parity taps ----------
int evplp_download_records(evplp_handle c, uint64_t firstRecord, uint64_t numRecords, EvplpRecord* out) {
    NEED(c != nullptr && out != nullptr, "evplp_download_records: NULL argument");
    NEED(firstRecord + numRecords <= c->numRecords, "evplp_download_records: window outside the record buffer");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(out, c->records.p + firstRecord, numRecords * sizeof(EvplpRecord), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return check_overflow(c);
}

int evplp_upload_records(evplp_handle c, uint32_t firstPath, const EvplpRecord* records, uint64_t numRecords) {
    NEED(c != nullptr && records != nullptr, "evplp_upload_records: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(c->records.reserve(numRecords));
    CU(cudaMemcpyAsync(c->records.p, records, numRecords * sizeof(EvplpRecord), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->numRecords = numRecords;
    c->recordsFirstPath = firstPath;
    return EVPLP_OK;
}

int evplp_download_gbuffer(evplp_handle c, float* planes, int32_t* primIds) {
    NEED(c != nullptr, "evplp_download_gbuffer: NULL handle");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->W * c->H;
    if (planes) CU(cudaMemcpyAsync(planes, c->gbuf.p, sizeof(float4) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
    if (primIds) CU(cudaMemcpyAsync(primIds, c->gprim.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return check_overflow(c);
}

int evplp_upload_gbuffer(evplp_handle c, const float* planes, const int32_t* primIds) {
    NEED(c != nullptr && planes && primIds, "evplp_upload_gbuffer: NULL argument");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->W * c->H;
    CU(cudaMemcpyAsync(c->gbuf.p, planes, sizeof(float4) * 4 * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->gprim.p, primIds, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->gbufValid = true;
    return EVPLP_OK;
}

int evplp_bvh_info(evplp_handle c, EvplpBvhInfo* info) {
    NEED(c != nullptr && info != nullptr, "evplp_bvh_info: NULL argument");
    NEED(c->bvhBuilt, "evplp_bvh_info: BVH not built");
    info->numPrims = (uint32_t)c->numPrims;
    info->numInternal = c->numPrims > 1 ? (uint32_t)c->numPrims - 1 : 0;
    info->wideNodeBytes = (uint32_t)sizeof(WideNode);
    for (int k = 0; k < 3; k++) { info->sceneMin[k] = c->sceneMin[k]; info->sceneMax[k] = c->sceneMax[k]; }
    return EVPLP_OK;
}

int evplp_download_bvh(evplp_handle c, uint64_t* mortonCodes, uint32_t* sortedPrimIds, int32_t* left, int32_t* right,
                       int32_t* parent, float* nodeBounds) {
    NEED(c != nullptr, "evplp_download_bvh: NULL handle");
    NEED(c->bvhBuilt, "evplp_download_bvh: BVH not built");
    CU(cudaSetDevice(c->device));
    const size_t n = c->numPrims, ni = n > 1 ? n - 1 : 0;
    if (mortonCodes && n) CU(cudaMemcpyAsync(mortonCodes, c->codesSorted.p, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    if (sortedPrimIds && n) CU(cudaMemcpyAsync(sortedPrimIds, c->primIdsSorted.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    const bool haveTopo = (int)n > c->opt.bvhLeafMax;  // tiny scenes skip the radix tree
    if (ni && !haveTopo && (left || right || parent || nodeBounds))
        return fail(EVPLP_ERR_INVALID, "evplp_download_bvh: scenes with <= BVH_LEAF_MAX triangles have no radix tree");
    if (left && ni) CU(cudaMemcpyAsync(left, c->left.p, 4 * ni, cudaMemcpyDeviceToHost, c->stream));
    if (right && ni) CU(cudaMemcpyAsync(right, c->right.p, 4 * ni, cudaMemcpyDeviceToHost, c->stream));
    if (parent && ni) CU(cudaMemcpyAsync(parent, c->parent.p, 4 * ni, cudaMemcpyDeviceToHost, c->stream));
    if (nodeBounds && ni) CU(cudaMemcpyAsync(nodeBounds, c->nodeBounds.p, 24 * ni, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return EVPLP_OK;
}

int evplp_trace_rays(evplp_handle c, const float* rays, uint64_t numRays, int anyHit, int32_t* outPrim, float* outT) {
    NEED(c != nullptr && rays && outPrim, "evplp_trace_rays: NULL argument");
    NEED(c->bvhBuilt, "evplp_trace_rays: BVH not built");
    NEED(anyHit >= 0 && anyHit <= 3, "evplp_trace_rays: anyHit must be 0 (closest), 1 (any), 2 (warp-cooperative any) or 3 (closest through the quantised nodes)");
    CU(cudaSetDevice(c->device));
    if (numRays == 0) return EVPLP_OK;
    float* dRays = nullptr; int32_t* dPrim = nullptr; float* dT = nullptr;
    CU(cudaMalloc((void**)&dRays, numRays * 32));
    CU(cudaMalloc((void**)&dPrim, numRays * 4));
    CU(cudaMalloc((void**)&dT, numRays * 4));
    CU(cudaMemcpyAsync(dRays, rays, numRays * 32, cudaMemcpyHostToDevice, c->stream));
    cudaError_t e = launch_trace_rays(c, dRays, numRays, anyHit, dPrim, dT);
    if (e == cudaSuccess) e = cudaMemcpyAsync(outPrim, dPrim, numRays * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && outT) e = cudaMemcpyAsync(outT, dT, numRays * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dRays); cudaFree(dPrim); cudaFree(dT);
    CU(e);
    return check_overflow(c);
}

int evplp_download_accum(evplp_handle c, int64_t* vpl, int64_t* photon, uint32_t* light) {
    NEED(c != nullptr, "evplp_download_accum: NULL handle");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->W * c->H;
    if (vpl) CU(cudaMemcpyAsync(vpl, c->accVpl.p, 24 * n, cudaMemcpyDeviceToHost, c->stream));
    if (photon) CU(cudaMemcpyAsync(photon, c->accPhoton.p, 24 * n, cudaMemcpyDeviceToHost, c->stream));
    if (light) CU(cudaMemcpyAsync(light, c->accLight.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return check_overflow(c);
}

int evplp_debug_uniforms(evplp_handle c, uint32_t seed, uint32_t subsequence, uint32_t n, float* out) {
    NEED(c != nullptr && out != nullptr, "evplp_debug_uniforms: NULL argument");
    CU(cudaSetDevice(c->device));
    int rc = ensure_skip_matrix(c, subsequence);
    if (rc) return rc;
    float* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)(n ? n : 1) * 4));
    cudaError_t e = launch_debug_uniforms(c, seed, n, d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    CU(e);
    return EVPLP_OK;
}

int evplp_debug_curand(evplp_handle c, uint32_t seed, uint32_t subsequence, uint32_t n, float* out) {
    NEED(c != nullptr && out != nullptr, "evplp_debug_curand: NULL argument");
    CU(cudaSetDevice(c->device));
    float* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)(n ? n : 1) * 4));
    cudaError_t e = launch_debug_curand(c, seed, subsequence, n, d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    CU(e);
    return EVPLP_OK;
}

int evplp_debug_math(evplp_handle c, int op, const float* x, const float* y, uint32_t n, float* out) {
    NEED(c != nullptr && x && out, "evplp_debug_math: NULL argument");
    CU(cudaSetDevice(c->device));
    float *dx = nullptr, *dy = nullptr, *dout = nullptr;
    const size_t bytes = (size_t)(n ? n : 1) * 4;
    CU(cudaMalloc((void**)&dx, bytes)); CU(cudaMalloc((void**)&dy, bytes)); CU(cudaMalloc((void**)&dout, bytes));
    CU(cudaMemcpyAsync(dx, x, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    if (y) CU(cudaMemcpyAsync(dy, y, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    else CU(cudaMemsetAsync(dy, 0, bytes, c->stream));
    cudaError_t e = launch_debug_math(c, op, dx, dy, n, dout);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    CU(e);
    return EVPLP_OK;
}

int evplp_stats(evplp_handle c, EvplpStats* stats) {
    NEED(c != nullptr && stats != nullptr, "evplp_stats: NULL argument");
    CU(cudaSetDevice(c->device));
    DevStats ds;
    CU(cudaMemcpyAsync(&ds, c->devStats.p, sizeof(ds), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *stats = c->stats;
    stats->shadowRays += ds.shadowRays;
    stats->splatFragments += ds.splatFragments;
    stats->closestRays += ds.closestRays;
    stats->gatherPairs += ds.gatherPairs;
    // emitted counts of the current record window
    stats->emittedVpls = 0; stats->emittedPhotons = 0;
    if (c->numRecords) {
        unsigned long long counts[2] = {0, 0};
        CU(launch_count_flags(c, counts));
        stats->emittedVpls = counts[0];
        stats->emittedPhotons = counts[1];
    }
    if (ds.stackOverflow) return fail(EVPLP_ERR_CUDA, "BVH traversal stack overflow");
    return EVPLP_OK;
}

int evplp_reset_stats(evplp_handle c) {
    NEED(c != nullptr, "evplp_reset_stats: NULL handle");
    CU(cudaSetDevice(c->device));
    memset(&c->stats, 0, sizeof(c->stats));
    CU(cudaMemsetAsync(c->devStats.p, 0, sizeof(DevStats), c->stream));
    return EVPLP_OK;
}

int evplp_last_stage_ms(evplp_handle c, int stage, float* ms) {
    NEED(c != nullptr && ms != nullptr, "evplp_last_stage_ms: NULL argument");
    NEED(stage >= 0 && stage < ST_COUNT, "evplp_last_stage_ms: stage out of range");
    NEED(c->stageValid[stage], "evplp_last_stage_ms: the stage has not run yet");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->stageB[stage]));
    CU(cudaEventElapsedTime(ms, c->stageA[stage], c->stageB[stage]));
    return EVPLP_OK;
}

int evplp_synchronize(evplp_handle c) {
    NEED(c != nullptr, "evplp_synchronize: NULL handle");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return check_overflow(c);
}

int evplp_launch_count(evplp_handle c, uint64_t* count) {
    NEED(c != nullptr && count != nullptr, "evplp_launch_count: NULL argument");
    *count = c->launches;
    return EVPLP_OK;
}

int evplp_debug_counters(evplp_handle c, uint64_t out[8]) {
    NEED(c != nullptr && out != nullptr, "evplp_debug_counters: NULL argument");
    CU(cudaSetDevice(c->device));
    DevStats ds;
    CU(cudaMemcpyAsync(&ds, c->devStats.p, sizeof(ds), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out[0] = ds.shaftSteps; out[1] = ds.shaftFallbacks; out[2] = ds.shaftNodeVisits; out[3] = ds.shaftCandLeaves;
    out[4] = (uint64_t)c->numNodes; out[5] = (uint64_t)c->numShaftNodes; out[6] = ds.clusterDescents; out[7] = ds.clusterSplits;
    return EVPLP_OK;
}

int evplp_debug_cluster_hist(evplp_handle c, uint64_t out[16]) {
    NEED(c != nullptr && out != nullptr, "evplp_debug_cluster_hist: NULL argument");
    CU(cudaSetDevice(c->device));
    DevStats ds;
    CU(cudaMemcpyAsync(&ds, c->devStats.p, sizeof(ds), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 16; k++) out[k] = ds.clusterHist[k];
    return EVPLP_OK;
}

int evplp_event_record(evplp_handle c, int slot) {
    NEED(c != nullptr && slot >= 0 && slot < 4, "evplp_event_record: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->userEv[slot], c->stream));
    return EVPLP_OK;
}

int evplp_event_elapsed_ms(evplp_handle c, int slotA, int slotB, float* ms) {
    NEED(c != nullptr && ms && slotA >= 0 && slotA < 4 && slotB >= 0 && slotB < 4, "evplp_event_elapsed_ms: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->userEv[slotB]));
    CU(cudaEventElapsedTime(ms, c->userEv[slotA], c->userEv[slotB]));
    return EVPLP_OK;
}

int evplp_set_option(evplp_handle c, const char* name, int value) {
    NEED(name != nullptr, "evplp_set_option: NULL name");
    // c == NULL: the defaults a handle created AFTERWARDS starts from; otherwise only this handle
    std::unique_lock<std::mutex> lock(g_defaultsMutex, std::defer_lock);
    if (!c) lock.lock();
    evplp::Options& o = c ? c->opt : g_defaultOptions;
    struct Knob { const char* name; int* slot; int lo, hi; };
    int splatMaxEntries = 0;
    const Knob knobs[] = {
        {"gather_chunks", &o.gatherChunks, 0, 4096},
        {"gather_band_stride", &o.bandStride, 0, 65536}, {"gather_band_offset", &o.bandOffset, 0, 65535},
        {"gather_min_blocks", &o.gatherMinBlocks, 0, 5},
        {"gather_mode", &o.gatherMode, 0, 2}, {"gather_algo", &o.gatherAlgo, 0, 2}, {"gather_cluster_size", &o.clusterSize, 1, 32},
        {"shaft_max_candidates", &o.shaftCandMax, 1, evplp::SHAFT_CAND}, {"shaft_streak", &o.shaftStreak, 1, 1 << 20},
        {"shaft_skip", &o.shaftSkip, 0, 1 << 20},
        {"gather_shared_batches", &o.sharedBatches, 1, 1 << 20}, {"gather_vpl_batches", &o.vplBatches, 1, 1 << 20},
        {"gather_cluster_skip_max", &o.clusterSkipMax, 0, 1 << 20}, {"gather_cluster_extent_permille", &o.clusterExtentPermille, 0, 1000},
        {"gather_lpt", &o.gatherLpt, 0, 1}, {"gather_persistent", &o.gatherPersistent, 0, 1},
        {"splat_group", &o.splatGroup, 0, 32}, {"splat_mode", &o.splatMode, 0, 1}, {"splat_max_entries", &splatMaxEntries, 0, 0x7fffffff},
        {"bvh_leaf_max", &o.bvhLeafMax, 1, 8}, {"shaft_leaf_max", &o.shaftLeafMax, 1, 8},
    };
    for (const Knob& k : knobs) {
        if (strcmp(name, k.name) != 0) continue;
        if (value < k.lo || value > k.hi)
            return fail(EVPLP_ERR_INVALID, std::string("evplp_set_option: ") + name + " must be in [" + std::to_string(k.lo) + ", " + std::to_string(k.hi) + "]");
        if (k.slot == &o.splatGroup && !(value == 0 || value == 1 || value == 8 || value == 32))
            return fail(EVPLP_ERR_INVALID, "evplp_set_option: splat_group must be 0, 1, 8 or 32");
        *k.slot = value;
        if (k.slot == &splatMaxEntries) o.splatMaxEntries = value;
        if (c && (k.slot == &o.bandStride || k.slot == &o.bandOffset)) c->gatherCostValid = false;
        return EVPLP_OK;
    }
    return fail(EVPLP_ERR_INVALID, std::string("evplp_set_option: unknown option ") + name);
}

}  // extern "C"
