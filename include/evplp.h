/* evplp.h -- C ABI of the B200-native EVPLP hot path (libevplp_b200.so).
 *
 * This is the drop-in boundary for the reference's technique class
 *   class RtTechnique { virtual void render(shared_ptr<RtScene>&, const glm::vec2&, const nlohmann::json&) = 0; }
 *   (reference: reflectcuts/realtimetechniques/rttechnique.h:6-10)
 * and specifically for the per-iteration stages of RtComPhoton / RtLvcComPhoton
 *   (reference: reflectcuts/realtimetechniques/rtcomphoton/rtcomphoton.h:646-1068,
 *               reflectcuts/realtimetechniques/rtcomphoton/rtlvccomphoton.h).
 * Each entry point below names the reference member function / OptiX or GL program it
 * replaces.  The C++ host classes in evplp_b200/host/ (same names as the reference:
 * RtScene, RtComPhoton, ...) are written purely against this header.
 *
 * Conventions: plain C, opaque handle, int error codes (0 = EVPLP_OK) plus
 * evplp_last_error(); no exceptions cross the ABI; every host array is caller-owned
 * and copied during the call; one host thread per handle; one handle per GPU.
 * There is NO CPU fallback: every compute entry point runs sm_100a kernels or fails.
 */
#ifndef EVPLP_H_
#define EVPLP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVPLP_OK 0
#define EVPLP_ERR_INVALID 1   /* bad argument / call order                       */
#define EVPLP_ERR_CUDA 2      /* CUDA runtime error (text in evplp_last_error)   */
#define EVPLP_ERR_NO_DEVICE 3 /* no usable sm_100 GPU                            */
#define EVPLP_ERR_NCCL 4

/* PhotonRecordFlag -- reference: rtcomphoton/rtphotonrecord.h:9-15 */
#define EVPLP_FLAG_USABLE_VPL 1u
#define EVPLP_FLAG_USABLE_PHOTON 2u
#define EVPLP_FLAG_LAMBERT_ONLY 4u
#define EVPLP_FLAG_PHONG_ONLY 8u

/* EMis -- reference: rtcomphoton.h:1199-1206 (EMisModeStrMap order) */
#define EVPLP_MIS_ONE 0
#define EVPLP_MIS_BALANCE 1
#define EVPLP_MIS_MAX 2
#define EVPLP_MIS_POWER2 3
#define EVPLP_MIS_GEOMETRY_CLAMP 4
#define EVPLP_MIS_GEOMETRY_BRDF_CLAMP 5

/* Gather flavours: splatColor (lighttracing.cu:348-379), splatSplotch (lighttracing.cu:689-722),
 * LVC splatColor (lvclighttracing.cu:348-387). */
#define EVPLP_GATHER_VPL 0
#define EVPLP_GATHER_VSL 1
#define EVPLP_GATHER_LVC 2

/* One light-path vertex: byte-compatible with RtPhotonRecord (rtphotonrecord.h:17-25)
 * and its GLSL std430 mirror (photonsplatinstanced.vert:7-15). 96 bytes. */
typedef struct EvplpRecord {
    float position[3];
    uint32_t flags;
    float normal[3];
    float pSelectLambert;
    float flux[3];
    float padding1;
    float fluxDir[3];
    float padding2;
    float lambertReflectance[3];
    float padding3;
    float phongReflectance[3];
    float phongExponent;
} EvplpRecord;

/* One RtMesh (rtcommon.h:464-467): SoA float arrays + int32 triangle indices. */
typedef struct EvplpMeshDesc {
    const float* vertices;    /* numVertices * 3 */
    const float* texcoords;   /* numVertices * 2, may be NULL => (0,0) like rtcommon.h:700-704 */
    const int32_t* indices;   /* numTriangles * 3, local to this mesh */
    int32_t numVertices;
    int32_t numTriangles;
    int32_t matIndex;         /* RtMesh::mMatIndex */
} EvplpMeshDesc;

/* One RtMaterial (rtcommon.h: three RGBA32F RtTextures + lightIntensity).  Texture rows
 * are stored bottom-up exactly as RtTexture::mData (stb flip-on-load, rtcommon.h:32). */
typedef struct EvplpMaterialDesc {
    const float* lambertReflectance; /* w*h*4 floats */
    int32_t lambertW, lambertH;
    const float* phongReflectance;
    int32_t phongW, phongH;
    const float* phongExponent;      /* .x is used */
    int32_t exponentW, exponentH;
    float lightIntensity[4];         /* RtMaterial::mLightIntensity (pi-scaled for the light) */
} EvplpMaterialDesc;

/* The per-iteration "uniforms" the reference pushes into the OptiX context and the GL
 * programs (rtcomphoton.h:895-930, 873, 799-823) plus the camera of this iteration. */
typedef struct EvplpParams {
    float cameraPosition[3];
    /* lookAtRH basis of the camera (glm gtc/matrix_transform.inl:521-546): forward f,
     * right s, up u; a pixel with un-jittered NDC (nx, ny) looks along
     * f + s * (nx * tanHalfFovX) + u * (ny * tanHalfFovY). */
    float camForward[3];
    float camRight[3];
    float camUp[3];
    float tanHalfFovX;
    float tanHalfFovY;
    float jitter[2];      /* NDC translation of this iteration (rtcomphoton.h:949-951) */
    float nearDist;       /* 0.1  (rtcommon.h:586) */
    float farDist;        /* 100  (rtcommon.h:586) */
    uint32_t numLightPaths;
    uint32_t numVplLightPaths;
    uint32_t numPhotonsPerLightPath; /* numMaxBounces + 1 */
    float radius;          /* mPhotonRadius */
    float pdfMc;           /* mPrecomptedPdfMc */
    uint32_t misMode;      /* EVPLP_MIS_* */
    float clampingValue;
    uint32_t doAccumulate; /* 1 = accumulate, 0 = cleareveryframe */
    float vslRadius;
    float vslInvPiRadius2;
    uint32_t rngSeed;      /* numIterations + rngOffset; used by VSL / LVC gathers */
} EvplpParams;

typedef struct EvplpTile {
    int32_t x0, y0, x1, y1; /* half-open pixel rectangle [x0,x1) x [y0,y1) */
} EvplpTile;

typedef struct EvplpStats {
    uint64_t emittedVpls;     /* records with IsUsableVpl in the traced window      */
    uint64_t emittedPhotons;  /* records with IsUsablePhoton in the traced window   */
    uint64_t gatherPairs;     /* (pixel, usable VPL) pairs submitted, cumulative    */
    uint64_t shadowRays;      /* pairs that passed the cosine test, cumulative      */
    uint64_t splatPhotons;    /* usable photons submitted to the splat, cumulative  */
    uint64_t splatFragments;  /* (photon, pixel) pairs inside the radius, cumulative*/
    uint64_t closestRays;     /* closest-hit rays traced (light trace + gbuffer)    */
} EvplpStats;

/* BVH download (parity tap): binary LBVH arrays + the wide nodes built from them. */
typedef struct EvplpBvhInfo {
    uint32_t numPrims;        /* triangles in the BVH                                */
    uint32_t numInternal;     /* numPrims - 1 (0 when numPrims <= 1)                 */
    uint32_t wideNodeBytes;   /* sizeof one wide node                                */
    float sceneMin[3], sceneMax[3];
} EvplpBvhInfo;

typedef struct EvplpContext* evplp_handle;

const char* evplp_last_error(void);
const char* evplp_version(void);
int evplp_device_count(void);

/* replaces RtComPhoton::setup() context creation (rtcomphoton.h:646-708, 419-431) */
int evplp_create(int device, int width, int height, evplp_handle* out);
/* replaces RtComPhoton::destroy() (rtcomphoton.h:1135-1138) */
int evplp_destroy(evplp_handle h);

/* replaces RtMesh::createOptixMeshBuffer/createOptixGeometry, RtMaterial::createOptixTextures,
 * RtAreaLight::createOptixCdf and the areaLight* context variables
 * (rtcommon.h:330-429, 501-531; rtcomphoton.h:678-703).
 * lightIntensityPrecomputed = intensity.rgb * pi, .w = emission Phong exponent (rtcommon.h:780-790);
 * lightIntensityDisplay = the un-scaled intensity drawn by light.frag (rtcomphoton.h:845). */
int evplp_upload_scene(evplp_handle h, const EvplpMeshDesc* meshes, int32_t numMeshes,
                       const EvplpMaterialDesc* materials, int32_t numMaterials,
                       int32_t lightMeshIndex, const float lightIntensityPrecomputed[4],
                       const float lightIntensityDisplay[4]);

/* replaces the OptiX "Trbvh" acceleration build over all GeometryInstances
 * (rtcomphoton.h:705-707; bounds program triangleintersect.cu:62-82). */
int evplp_build_bvh(evplp_handle h);

/* replaces the rtContext[...]->set* / glUniform pushes (rtcomphoton.h:895-930, 1043-1061) */
int evplp_set_params(evplp_handle h, const EvplpParams* params);

/* clears the accumulation layers (glClear of the photon / light FBOs, and the VPL
 * buffer's first write with doAccumulate = 0; rtcomphoton.h:889-890, 978-981) */
int evplp_clear_accum(evplp_handle h);

/* replaces runDeferredProgram (rtcomphoton.h:710-754; shaders/deferred.*) */
int evplp_gbuffer(evplp_handle h);

/* replaces runOptixLightTracingProgram -> tracePhotons (rtcomphoton.h:869-881;
 * lighttracing.cu:192-250).  Traces light paths [firstPath, firstPath + numPaths); the
 * record of (path p, bounce b) lands in record slot (p - firstPath) * (B+1) + b. */
int evplp_light_trace(evplp_handle h, uint32_t rngSeed, uint32_t firstPath, uint32_t numPaths);

/* replaces runOptixVplProgram -> splatColor / splatSplotch / LVC splatColor
 * (rtcomphoton.h:857-867; lighttracing.cu:348-379, 689-722; lvclighttracing.cu:348-387).
 * tile == NULL means the whole image. */
int evplp_vpl_gather(evplp_handle h, const EvplpTile* tile, int gatherMode);

/* RtPt2, replaces runOptixPtProgram -> splatColor / pathTraceSimple / rtMaterialClosestHit
 * (rtpt/rtpt2.h:561-570; pathtracing.cu:112-377): one path per pixel from the G-buffer first hit with next-event
 * estimation + MIS, per-pixel stream curand_init(pixel, params.rngSeed, 0); the result is added to (doAccumulate) or
 * replaces the VPL accumulation layer (the technique's outputBuffer).  The reference's own ground-truth generator. */
int evplp_path_trace(evplp_handle h, const EvplpTile* tile, uint32_t maxBounces);

/* replaces runPhotonSplat (rtcomphoton.h:789-837; shaders/photonsplatinstanced.*).
 * firstRecord/numRecords index the record window written by the last evplp_light_trace. */
int evplp_photon_splat(evplp_handle h, uint64_t firstRecord, uint64_t numRecords, const EvplpTile* tile);

/* replaces runLightProgram (rtcomphoton.h:839-855, 985-995; shaders/light.*): the light mesh seen through the UN-jittered
 * camera ("we don't jitter light source"); the light layer is overwritten with 1 where the closest surface is the light. */
int evplp_light_pass(evplp_handle h);

/* The number of iterations accumulated into the layers (the reference's numIterations, the 1 / N of runFinalProgram,
 * rtcomphoton.h:1001, 1122): kept next to the layers on the device so that evplp_reduce sums it with them -- ranks that
 * stop at different iteration counts (time limit) still normalise by what was actually rendered.  evplp_clear_accum zeroes it. */
int evplp_add_iterations(evplp_handle h, int64_t n);
int evplp_iterations(evplp_handle h, int64_t* n);

/* NEW (the reference is single-GPU): sum the accumulation layers over all ranks of
 * an NCCL communicator (ncclComm_t passed as void*).  Exact: the layers are int64. */
int evplp_reduce(evplp_handle h, void* ncclComm);
/* Device pointers + element counts of the accumulation layers, so a host that already
 * owns a communicator (e.g. torch.distributed) can all-reduce them itself.
 * layer: 0 = VPL int64[W*H*3], 1 = photon int64[W*H*3], 2 = light uint32[W*H], 3 = iteration count int64[1]. */
int evplp_accum_layer(evplp_handle h, int layer, void** devPtr, uint64_t* numInt64);

/* replaces runFinalProgram + dumpImage (rtcomphoton.h:756-787, 225-249; shaders/final.frag):
 * out = step(light.x*lightScale, 0) * (vpl*vplScale + photon*photonScale) + light*lightScale,
 * optional gamma 1/2.2; rows bottom-up like glReadPixels; hostRGB = W*H*3 floats. */
int evplp_resolve(evplp_handle h, float vplScale, float photonScale, float lightScale,
                  int doGammaCorrection, float* hostRGB);

/* ---- debug / parity taps (no reference counterpart; used by tests and the oracle checks) ---- */
int evplp_download_records(evplp_handle h, uint64_t firstRecord, uint64_t numRecords, EvplpRecord* out);
int evplp_upload_records(evplp_handle h, uint32_t firstPath, const EvplpRecord* records, uint64_t numRecords);
/* planes: position.xyzw, normal.xyz_, diffuse.xyz_, phong.xyz+exponent: 4 * W*H*4 floats; primIds W*H */
int evplp_download_gbuffer(evplp_handle h, float* planes, int32_t* primIds);
int evplp_upload_gbuffer(evplp_handle h, const float* planes, const int32_t* primIds);
int evplp_bvh_info(evplp_handle h, EvplpBvhInfo* info);
/* mortonCodes[numPrims], sortedPrimIds[numPrims], left/right/parent[numInternal]
 * (child >= 0: internal node, child < 0: ~leafIndex), nodeBounds[numInternal*6] (min xyz, max xyz).
 * Any pointer may be NULL. */
int evplp_download_bvh(evplp_handle h, uint64_t* mortonCodes, uint32_t* sortedPrimIds,
                       int32_t* left, int32_t* right, int32_t* parent, float* nodeBounds);
/* Trace numRays arbitrary rays (origin xyz, dir xyz, tmin, tmax = 8 floats each) through the
 * product traversal kernels: closest (anyHit=0) writes primId (-1 = miss) and t; any (1) writes 0/1; 2 = the gather's
 * warp-cooperative any hit; 3 = closest hit through the quantised nodes light tracing uses (same hits as 0). */
int evplp_trace_rays(evplp_handle h, const float* rays, uint64_t numRays, int anyHit,
                     int32_t* outPrim, float* outT);
/* Raw accumulation layers: int64[W*H*3] VPL, int64[W*H*3] photon, uint32[W*H] light count. */
int evplp_download_accum(evplp_handle h, int64_t* vpl, int64_t* photon, uint32_t* light);
/* First n cuRAND-compatible uniforms of stream (seed, subsequence) produced on the device. */
int evplp_debug_uniforms(evplp_handle h, uint32_t seed, uint32_t subsequence, uint32_t n, float* out);
/* The same stream from the REAL cuRAND device API, called exactly as the reference does
 * (curand_init(seed, subsequence, 0, &state); curand_uniform(&state) -- lighttracing.cu:202-203).
 * Pins the XORWOW restatements (product and oracle) against cuRAND itself. */
int evplp_debug_curand(evplp_handle h, uint32_t seed, uint32_t subsequence, uint32_t n, float* out);
/* Host-only self-check: the 4-bit-table form of the per-launch XORWOW skip (what light tracing uses) against the 160x160
 * matrix form, over numSeeds launch indices; returns the number of differing states (0 expected).  No device needed. */
int evplp_debug_xorwow_tables(uint32_t subsequence, uint32_t firstSeed, uint32_t numSeeds);
/* detmath on device: op 0 sin, 1 cos, 2 pow(x,y), 3 asin, 4 sqrt; x,y,out are n floats */
int evplp_debug_math(evplp_handle h, int op, const float* x, const float* y, uint32_t n, float* out);
int evplp_stats(evplp_handle h, EvplpStats* stats);
int evplp_reset_stats(evplp_handle h);
/* Device time (ms, CUDA events on the handle's stream, bracketing the stage's main kernel) of the
 * most recent call of each stage: 0 bvh (whole build), 1 gbuffer, 2 light_trace, 3 gather, 4 splat, 5 resolve. */
int evplp_last_stage_ms(evplp_handle h, int stage, float* ms);
int evplp_synchronize(evplp_handle h);
/* Number of kernel launches issued by this handle since creation (bench.py gpu_launches). */
int evplp_launch_count(evplp_handle h, uint64_t* count);
/* Tuning counters since the last evplp_reset_stats: [0] (warp, VPL) steps of the shaft gather, [1] steps that fell back to the
 * per-ray packet traversal, [2] 32-wide nodes visited, [3] candidate leaves tested, [4] 4-wide nodes, [5] 32-wide nodes. */
int evplp_debug_counters(evplp_handle h, uint64_t out[8]);
/* Cluster gather (gather_algo 1) since the last evplp_reset_stats, filled only by a tuning build (-DEVPLP_GATHER_HIST):
 * [0..6] descents that collected 0 / <= 8 / <= 32 / <= 64 / <= 128 / <= 512 / more candidate leaves, [8] (cluster, tile) pairs,
 * [9] of which some VPL lights the tile, [10] such VPLs.  evplp_debug_counters [6], [7] = descents, candidate batches. */
int evplp_debug_cluster_hist(evplp_handle h, uint64_t out[16]);
/* CUDA events on the handle's stream (4 slots): device-side timing of any span of calls. */
int evplp_event_record(evplp_handle h, int slot);
int evplp_event_elapsed_ms(evplp_handle h, int slotA, int slotB, float* ms);
/* Tuning knobs (no reference counterpart).  "gather_chunks": number of slices the VPL list is
 * split into across thread blocks (0 = automatic; 1 = every pixel sums its VPLs in record
 * order in one thread, which makes the gather bit-identical to the scalar oracle; N > 1 = N contiguous ranges of
 * ceil(total / N) usable VPLs, each summed in record order and added in Q31.32 with integer atomics: deterministic, and
 * bit-identical to the oracle evaluated range by range).
 * "gather_band_stride" / "gather_band_offset": the gather only renders the 8x4-pixel tiles t = offset (mod stride) of its
 * rectangle -- the interleaved image partition of a single heavy frame over N GPUs (stride = N, offset = rank; VSL / LVC gathers:
 * 16-row bands).
 * "gather_mode": 1 (default) = the warp descends a 32-wide hierarchy with one conservative shaft-vs-box test per child
 * and runs the exact per-ray triangle tests on the collected candidate leaves; 0 = per-ray packet traversal of the 4-wide
 * hierarchy; 2 = shaft traversal in the VSL gather too (sampling-bound: no gain measured).  "shaft_max_candidates" (candidate leaves per step before falling back to mode 0, <= 128 = default),
 * "shaft_streak" / "shaft_skip" (after 3 overflowing steps in a row a warp sends its next 256 steps straight to the packet traversal: overflow is a property of the tile),
 * "gather_persistent" (default 1: one resident wave of blocks whose warps draw 8x4-pixel tiles from a global counter; 0 = one
 * tile per warp of a full grid), "gather_lpt" (default 1: the tiles are drawn in descending order of the cycles they took in the
 * previous launch of the same grid -- longest processing time first; results do not depend on the order),
 * "bvh_leaf_max" / "shaft_leaf_max" (before evplp_build_bvh), "gather_min_blocks", "splat_mode" (0 tiled, 1 scatter),
 * "splat_group", "splat_max_entries": kernel variants.
 * "gather_algo": the VPL-cluster gather (gather_fast.cu: the usable VPLs in Morton order, clusters of "gather_cluster_size"
 * <= 16, one double-shaft descent per (cluster, 8x4-pixel tile) where the shaft is thin enough, shading tail with FMA
 * contraction: radiance within the 1e-4 tolerance, NOT bit-identical to the oracle) is used 1 (default) = from 16384 usable
 * VPLs on (it needs dense VPLs), 2 = always, 0 = never (the per-VPL exact-order kernel).  gather_chunks = 1 always selects
 * the exact-order kernel (the bit-exact test mode).  Cluster-gather knobs: "gather_cluster_extent_permille" (default 70: a run
 * of gather_cluster_size VPLs whose box edge exceeds that many 1/1000 of the scene's longest edge is cut into halves / quarters;
 * 0 = never), "gather_shared_batches" / "gather_vpl_batches" (default 3 / 3: batches of 96 candidate leaves a cluster's shared
 * descent / a single VPL's descent may stream before falling back to per-VPL descents / the packet traversal),
 * "gather_cluster_skip_max" (default 64: longest run of clusters that skip the shared attempt after a fat shaft).
 * Limits: evplp_photon_splat / evplp_vpl_gather compact at most 2^31 - 1 records per call (stream larger frames in chunks,
 * as RtComPhoton::runStreamed does).
 * Every handle owns its options: a non-NULL handle sets that handle only; a NULL handle sets the defaults that handles
 * created AFTERWARDS start from.  Values are range-checked (EVPLP_ERR_INVALID). */
int evplp_set_option(evplp_handle h, const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* EVPLP_H_ */
