// context.h -- the opaque EvplpContext behind evplp_handle (one per GPU) and the
// internal launch interface between the C ABI (capi.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/evplp.h"
#include "device_scene.h"

namespace evplp {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;  // capacity in elements
    cudaError_t reserve(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

enum Stage { ST_BVH = 0, ST_GBUFFER, ST_LIGHT_TRACE, ST_GATHER, ST_SPLAT, ST_RESOLVE, ST_COUNT };

struct DevStats {
    unsigned long long shadowRays, splatPhotons, splatFragments, closestRays, gatherPairs;
    unsigned long long shaftSteps, shaftFallbacks, shaftNodeVisits, shaftCandLeaves;  // tuning counters of the shaft gather
    unsigned long long clusterDescents, clusterSplits;                               // cluster gather: descents, overflow splits
    unsigned long long clusterHist[16];   // cluster gather: candidates per descent (0, <=8, <=32, <=64, <=96, overflow by span), live clusters / VPLs
    int stackOverflow;
    int pad;
};

// Tuning knobs (evplp_set_option).  Every handle owns a copy; the process-wide defaults are only what a handle starts from.
struct Options {
    int gatherChunks = 0;        // 0 = automatic; 1 = one thread sums a pixel's VPLs in record order (bit-exact test mode)
    int bandStride = 0, bandOffset = 0;   // image partition of the gather (multi-GPU): this handle owns tiles t = offset (mod stride)
    int gatherMinBlocks = 0;     // 0 = per-kernel default
    int gatherMode = 1;          // 1 = shaft traversal of the 32-wide hierarchy, 0 = per-ray packet traversal, 2 = shaft in VSL too
    int gatherAlgo = 1;          // 1 = VPL-cluster gather (tolerance mode, default), 0 = per-VPL exact-order gather
    int clusterSize = 16;        // VPLs per cluster of the cluster gather
    int clusterExtentPermille = 70;  // cluster gather: a run of clusterSize VPLs whose box edge exceeds this many 1/1000 of the scene's longest edge is cut into halves / quarters (0 = never)
    int sharedBatches = 3, vplBatches = 3, clusterSkipMax = 64;   // cluster gather: candidate batches a shared / a per-VPL descent may stream; longest run of clusters that skip the shared attempt after a fat shaft
    int shaftCandMax = 128, shaftStreak = 3, shaftSkip = 256;
    int gatherLpt = 1, gatherPersistent = 1;
    int splatGroup = 0, splatMode = 0;
    long long splatMaxEntries = 256ll * 1024 * 1024;
    int bvhLeafMax = BVH_LEAF_MAX, shaftLeafMax = 2;
};

}  // namespace evplp

struct EvplpContext {
    evplp::Options opt;
    int device = 0;
    int W = 0, H = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t stageA[evplp::ST_COUNT] = {}, stageB[evplp::ST_COUNT] = {};  // device time of each stage's main kernel(s)
    bool stageValid[evplp::ST_COUNT] = {};
    cudaEvent_t userEv[4] = {};                                              // evplp_event_record slots
    void stageBegin(int s) { cudaEventRecord(stageA[s], stream); }
    void stageEnd(int s) { cudaEventRecord(stageB[s], stream); stageValid[s] = true; }
    uint64_t launches = 0;

    // scene
    evplp::DevBuf<float4> triVerts, triLeaf, texPool;
    evplp::DevBuf<float2> triUV;
    evplp::DevBuf<evplp::DevMaterial> mats;
    evplp::DevBuf<float> lightCdf;
    int numPrims = 0, numMats = 0;
    int lightFirst = 0, lightCount = 0;
    float lightArea = 0.f;
    float lightIntensity[4] = {0, 0, 0, 0}, lightDisplay[4] = {0, 0, 0, 0};
    bool sceneLoaded = false;

    // BVH
    evplp::DevBuf<float> primLo, primHi;          // 3 floats per primitive, original order
    evplp::DevBuf<uint64_t> codes, codesSorted;
    evplp::DevBuf<uint32_t> primIds, primIdsSorted;
    evplp::DevBuf<int32_t> left, right, parent, leafParent, rangeFirst, rangeLast;
    evplp::DevBuf<float> nodeBounds;              // 6 floats per binary internal node
    evplp::DevBuf<uint32_t> refitFlags;
    evplp::DevBuf<evplp::WideNode> nodes;
    evplp::DevBuf<evplp::CNode> cnodes;           // quantised twin of `nodes` for the closest-hit traversal
    evplp::DevBuf<evplp::ShaftNode> shaftNodes;
    int numShaftNodes = 0;
    evplp::DevBuf<uint32_t> sceneBoundsEnc;       // 6 ordered-uint encodings
    evplp::DevBuf<uint8_t> sortTemp;
    evplp::DevBuf<uint32_t> queueA, queueB, counters;
    float sceneMin[3] = {0, 0, 0}, sceneMax[3] = {0, 0, 0};
    float boxPad = 0.f;
    int numNodes = 0;
    bool bvhBuilt = false;

    // per-iteration state
    EvplpParams params;
    bool paramsSet = false;
    evplp::DevBuf<uint32_t> skipMatrix;           // 800 words: composed XORWOW skip matrix
    evplp::DevBuf<uint32_t> skipTable;            // kSkipTableWords: the same matrix as 4-bit look-up tables (light tracing)
    uint32_t skipMatrixSeed = 0xffffffffu;
    bool skipMatrixValid = false;

    evplp::DevBuf<EvplpRecord> records;
    uint64_t numRecords = 0;                      // valid records of the last trace / upload
    uint32_t recordsFirstPath = 0;
    evplp::DevBuf<uint32_t> vplList;              // indices of usable VPL records (gather prefix)
    // cluster gather: Morton keys / record indices of the usable VPLs, their sorted order, the prepared VPLs (6 float4 each)
    // and the cluster boxes (2 float4 each)
    evplp::DevBuf<uint32_t> vplKeys, vplKeysSorted, vplVals, vplOrder;
    evplp::DevBuf<float4> vplPrepared, clusterBox;
    evplp::DevBuf<uint32_t> clusterSlots, clusterList;   // cluster layout of the cluster gather: packed (first << 5 | count)
    evplp::DevBuf<uint32_t> photonList;           // indices of usable photon records
    evplp::DevBuf<float4> splatPrep;              // per-photon constants of the fragment shader (5 float4 each)
    evplp::DevBuf<uint32_t> tileCount, tileOffset, tileCursor, tileList;  // photon bins of the tiled splat

    evplp::DevBuf<float4> gbuf;                   // 4 planes of W*H
    evplp::DevBuf<int32_t> gprim;
    bool gbufValid = false;

    evplp::DevBuf<long long> accVpl, accPhoton;   // Q31.32, W*H*3
    evplp::DevBuf<uint32_t> accLight;             // W*H
    evplp::DevBuf<unsigned long long> scratch64;  // small device scratch of the diagnostic taps (evplp_stats: emitted counts)
    evplp::DevBuf<long long> accCount;            // [0] iterations accumulated into the layers (reduced with them)
    float lightBoxMin[3] = {0, 0, 0}, lightBoxMax[3] = {0, 0, 0};   // world bounds of the light mesh (light pass rectangle)
    evplp::DevBuf<float> resolveOut;              // W*H*3
    float* resolvePinned = nullptr;

    // VPL gather work order: cycles every 8x4-pixel tile took in the previous launch of the same grid (a progressive run
    // renders the same view again and again), sorted so that the next launch draws the expensive tiles first
    evplp::DevBuf<uint32_t> gatherCost, gatherCostSorted, gatherIota, gatherOrder;
    uint64_t gatherSig[4] = {0, 0, 0, 0};          // launch grid + tile + band signature the costs belong to
    bool gatherCostValid = false;

    evplp::DevBuf<evplp::DevStats> devStats;
    EvplpStats stats;

    evplp::DevScene scene() const;
};

namespace evplp {

constexpr int GATHER_BATCH = 16;    // VPL records staged per warp and shared-memory batch
constexpr int GATHER_WARPS = 8;

struct GatherParams {
    V3 cameraPosition;
    unsigned misMode;
    float pdfMc, clampingValue;
    float invNumVpl;  // 1 / (float)numVplLightPaths
    unsigned doAccumulate;
    int x0, y0, x1, y1;  // tile
    int W, H;
    unsigned numChunks;  // VPL list split over gridDim.z
    float vslRadius, vslInvPiRadius2;
    unsigned numLightPaths, numVplLightPaths, B1;
    int shaftMode;               // 1 = shaft traversal (gather_mode option)
    int shaftCandMax;            // shaft gather: candidate leaves beyond which a (warp, VPL) step falls back to the packet traversal
    int bandStride, bandOffset;  // this launch owns the 16-row bands b = bandOffset (mod bandStride) of the tile (multi-GPU interleave)
    // VPL gather: the launch's (16x16-pixel block, VPL chunk) grid.  With persistent != 0 the kernel is launched with one
    // resident wave of blocks and every WARP draws the next 8x4-pixel tile of that grid from a global counter, so warps
    // whose tile is cheap (culled by the cosine test, sky) go on to new work instead of idling until their block ends.
    unsigned vgx, vgy, vgz;
    int persistent;
    int shaftStreak, shaftSkip;  // shaft gather: overflows in a row before, and number of, steps sent straight to the packet traversal
    // persistent VPL gather: this handle's share of the image = the 8x4-pixel tiles t = tOffset + k * tStride, k < ownedTiles, of a
    // row-major numbering with row pitch pitchX (one phantom tile per row when tilesX is a multiple of the stride, so that
    // consecutive rows do not give a handle the same columns); work item v = chunk * ownedTiles + k
    int tilePartition, tilesX, pitchX;
    uint32_t ownedTiles, tStride, tOffset;
};

struct TileShare { int tilesX, pitchX; uint32_t ownedTiles, stride, offset; uint64_t pixels; };
// the tiles of rectangle t that belong to this handle (gather_band_stride / gather_band_offset) and their pixel count
inline TileShare tile_share(const EvplpContext* c, EvplpTile t) {
    TileShare s;
    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    s.stride = c->opt.bandStride > 0 ? (uint32_t)c->opt.bandStride : 1u;
    s.offset = c->opt.bandStride > 0 ? (uint32_t)c->opt.bandOffset : 0u;
    s.tilesX = (tw + 7) / 8;
    const int tilesY = (th + 3) / 4;
    s.pitchX = (s.stride > 1 && s.tilesX % (int)s.stride == 0) ? s.tilesX + 1 : s.tilesX;
    const uint32_t total = (uint32_t)s.pitchX * (uint32_t)tilesY;
    s.ownedTiles = s.offset < total ? (total - s.offset + s.stride - 1) / s.stride : 0u;
    s.pixels = 0;
    for (uint32_t k = 0; k < s.ownedTiles; k++) {
        const uint32_t tt = s.offset + k * s.stride;
        const int ty = (int)(tt / (uint32_t)s.pitchX), tx = (int)(tt % (uint32_t)s.pitchX);
        if (tx >= s.tilesX) continue;
        const int w = tw - tx * 8 < 8 ? tw - tx * 8 : 8, h = th - ty * 4 < 4 ? th - ty * 4 : 4;
        s.pixels += (uint64_t)w * (uint64_t)h;
    }
    return s;
}

// gather_fast.cu: the VPL-cluster gather (tolerance mode); `count` = usable VPLs in c->vplList (already compacted)
cudaError_t launch_gather_cluster(EvplpContext* c, EvplpTile tile, GatherParams g, uint32_t count);

// bvh.cu
cudaError_t build_bvh_device(EvplpContext* c, std::string* err);
// stages.cu
cudaError_t launch_gbuffer(EvplpContext* c);
cudaError_t launch_light_trace(EvplpContext* c, uint32_t rngSeed, uint32_t firstPath, uint32_t numPaths);
cudaError_t launch_gather(EvplpContext* c, EvplpTile tile, int mode);
cudaError_t launch_path_trace(EvplpContext* c, EvplpTile tile, uint32_t maxBounces);
cudaError_t launch_splat(EvplpContext* c, uint64_t firstRecord, uint64_t numRecords, EvplpTile tile);
cudaError_t launch_light_pass(EvplpContext* c);
cudaError_t launch_add_count(EvplpContext* c, long long n);
cudaError_t launch_resolve(EvplpContext* c, float vplScale, float photonScale, float lightScale, int gamma);
cudaError_t launch_trace_rays(EvplpContext* c, const float* devRays, uint64_t n, int anyHit, int32_t* devPrim, float* devT);
cudaError_t launch_debug_uniforms(EvplpContext* c, uint32_t seed, uint32_t n, float* devOut);
cudaError_t launch_debug_curand(EvplpContext* c, uint32_t seed, uint32_t subsequence, uint32_t n, float* devOut);
cudaError_t launch_count_flags(EvplpContext* c, unsigned long long counts[2]);
cudaError_t launch_debug_math(EvplpContext* c, int op, const float* x, const float* y, uint32_t n, float* out);
cudaError_t launch_debug_fast_pow(EvplpContext* c, const float* x, const float* y, uint32_t n, float* out);  // gather_fast.cu
// host_xorwow.cpp
void xorwow_compose_skip(uint32_t subsequence, uint32_t* out800);

}  // namespace evplp
