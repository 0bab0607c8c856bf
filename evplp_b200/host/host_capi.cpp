// host_capi.cpp -- C entry points over the C++ host classes, for ctypes callers (tests, bench.py).
// No compute happens here: scene generation / loading is host data preparation and every
// render call goes through libevplp_b200.so.
// Compiled twice: into libevplp_host.so (everything; links libevplp_b200.so) and, with -DEVPLP_SCENE_ONLY, into
// libevplp_scene.so (scene generation / loading / descriptors only; links NOTHING of the product, so that the CPU
// reference arm of bench.py can build its inputs without mapping the product library).
#include <cstring>
#ifndef EVPLP_SCENE_ONLY
#include "rtcomphoton.h"
#include "rtpt2.h"
#else
#include "rtcommon.h"
#include "realtime.h"
#endif
#include "scenegen.h"

using namespace evplp_host;

static thread_local std::string g_hostErr;

struct HostScene {
    shared_ptr<RtScene> scene;
    std::vector<EvplpMeshDesc> md;
    std::vector<EvplpMaterialDesc> mt;
};

#ifndef EVPLP_SCENE_ONLY
struct HostTechnique {
    std::unique_ptr<RtComPhoton> tech;
    shared_ptr<RtScene> scene;
};
#endif

#define GUARD(...) try { __VA_ARGS__ } catch (const std::exception& e) { g_hostErr = e.what(); return -1; }

extern "C" {

const char* evplp_host_last_error(void) { return g_hostErr.c_str(); }

// Write the procedural stand-in of a named reference scene (OBJ + MTL + PPM + technique JSONs).
int evplp_host_export_scene(const char* name, const char* outDir, uint32_t seed, int detail, int resX, int resY) {
    GUARD(GenScene g = GenerateNamed(name, seed, detail); ExportScene(g, outDir, resX, resY); return 0;)
}

// Build the same scene in memory (identical triangles / materials to loading the exported files).
void* evplp_host_generate_scene(const char* name, uint32_t seed, int detail, float aspect) {
    try {
        auto* hs = new HostScene();
        hs->scene = ToRtScene(GenerateNamed(name, seed, detail), aspect);
        hs->scene->descriptors(hs->md, hs->mt);
        return hs;
    } catch (const std::exception& e) { g_hostErr = e.what(); return nullptr; }
}

// LoadScene(json) (main.cpp:42-85)
void* evplp_host_load_scene(const char* jsonPath) {
    try {
        Json json = Json::parse_file(jsonPath);
        auto* hs = new HostScene();
        hs->scene = LoadScene(json, jsonPath);
        if (!hs->scene) throw std::runtime_error("no scene in json");
        hs->scene->descriptors(hs->md, hs->mt);
        return hs;
    } catch (const std::exception& e) { g_hostErr = e.what(); return nullptr; }
}

void evplp_host_scene_destroy(void* s) { delete (HostScene*)s; }

// Descriptor views (valid while the scene lives): the same structs evplp_upload_scene takes.
int evplp_host_scene_descriptors(void* s, const EvplpMeshDesc** meshes, int32_t* numMeshes, const EvplpMaterialDesc** materials,
                                 int32_t* numMaterials, int32_t* lightMeshIndex, float lightPre[4], float lightDisplay[4]) {
    HostScene* hs = (HostScene*)s;
    *meshes = hs->md.data(); *numMeshes = (int32_t)hs->md.size();
    *materials = hs->mt.data(); *numMaterials = (int32_t)hs->mt.size();
    *lightMeshIndex = hs->scene->lightMeshIndex();
    const Vec4& p = hs->scene->mArealight->mPrecomputedLightIntensity;
    const Vec4& d = hs->scene->mArealight->mLightIntensity;
    lightPre[0] = p.x; lightPre[1] = p.y; lightPre[2] = p.z; lightPre[3] = p.w;
    lightDisplay[0] = d.x; lightDisplay[1] = d.y; lightDisplay[2] = d.z; lightDisplay[3] = d.w;
    return 0;
}

// scalars: [0] bounding-sphere radius, [1] total area, [2] number of triangles; camera: origin, f, s, u, tanX, tanY (14 floats)
int evplp_host_scene_info(void* s, float scalars[3], float camera[14]) {
    GUARD(
        HostScene* hs = (HostScene*)s;
        scalars[0] = hs->scene->findBoundingSphereRadius();
        scalars[1] = hs->scene->totalArea();
        scalars[2] = (float)hs->scene->numTriangles();
        Vec3 f, sv, u; float tx, ty;
        hs->scene->mCamera->basis(&f, &sv, &u, &tx, &ty);
        Vec3 o = hs->scene->mCamera->getOrigin();
        const float c[14] = {o.x, o.y, o.z, f.x, f.y, f.z, sv.x, sv.y, sv.z, u.x, u.y, u.z, tx, ty};
        memcpy(camera, c, sizeof(c));
        return 0;
    )
}

#ifndef EVPLP_SCENE_ONLY
// RtComPhoton / RtLvcComPhoton over a scene; `techniqueJson` is the text of the "photonfam" object.
// partitionMode: 0 = iterations round-robin over the ranks, 1 = image bands + light-path ranges (RtComPhoton::EPartition)
void* evplp_host_technique_create(void* s, const char* techniqueJson, int resX, int resY, int device, int lvc, int rank, int worldSize,
                                  int partitionMode) {
    try {
        HostScene* hs = (HostScene*)s;
        auto* ht = new HostTechnique();
        ht->scene = hs->scene;
        ht->tech.reset(lvc ? new RtLvcComPhoton(device) : new RtComPhoton(device));
        ht->tech->setPartition(rank, worldSize, partitionMode ? RtComPhoton::PartitionImage : RtComPhoton::PartitionIterations);
        ht->tech->setWriteOutputs(false);
        Vec2 res; res.x = (float)resX; res.y = (float)resY;
        ht->tech->parse(ht->scene, res, Json::parse(techniqueJson));
        ht->tech->setup();
        return ht;
    } catch (const std::exception& e) { g_hostErr = e.what(); return nullptr; }
}

// The multi-rank host logic without a device: parse the technique, then plan numIterations passes of the loop as rank `rank` of
// `worldSize` would (RtComPhoton::planNext / advanceSchedule -- the same code iterate() runs).  out: 16 doubles per iteration =
// iteration, render, jitter x, jitter y, rngSeed, photonRadius, clampingValue, pdfMc, vslRadius, vslInvPiRadius2,
// splatFirstPath, splatNumPaths, tileStride, tileOffset, drawLight, countIteration.
int evplp_host_technique_plan(void* s, const char* techniqueJson, int resX, int resY, int lvc, int rank, int worldSize, int partitionMode,
                              int numIterations, double* out) {
    GUARD(
        HostScene* hs = (HostScene*)s;
        std::unique_ptr<RtComPhoton> t(lvc ? new RtLvcComPhoton(0) : new RtComPhoton(0));
        t->setPartition(rank, worldSize, partitionMode ? RtComPhoton::PartitionImage : RtComPhoton::PartitionIterations);
        Vec2 res; res.x = (float)resX; res.y = (float)resY;
        t->parse(hs->scene, res, Json::parse(techniqueJson));
        t->setupHostOnly();
        for (int k = 0; k < numIterations; k++) {
            const RtComPhoton::IterationPlan p = t->planNext();
            double* o = out + 16 * (size_t)k;
            o[0] = p.iteration; o[1] = p.render; o[2] = p.jitter.x; o[3] = p.jitter.y; o[4] = p.rngSeed; o[5] = p.photonRadius;
            o[6] = p.clampingValue; o[7] = p.pdfMc; o[8] = p.vslRadius; o[9] = p.vslInvPiRadius2; o[10] = (double)p.splatFirstPath;
            o[11] = (double)p.splatNumPaths; o[12] = p.tileStride; o[13] = p.tileOffset; o[14] = p.drawLight; o[15] = p.countIteration;
            t->advanceSchedule();
        }
        return 0;
    )
}

void evplp_host_technique_set_max_paths_per_trace(void* t, uint64_t n) { ((HostTechnique*)t)->tech->setMaxPathsPerTrace(n); }

void* evplp_host_technique_handle(void* t) { return ((HostTechnique*)t)->tech->handle(); }

// one pass of the per-iteration loop body; returns 1 to continue, 0 when the loop ends, -1 on error
int evplp_host_technique_iterate(void* t) {
    GUARD(return ((HostTechnique*)t)->tech->iterate() ? 1 : 0;)
}

// state: photonRadius, clampingValue, pdfMc, vslRadius, vslInvPiRadius2, numIterations
int evplp_host_technique_state(void* t, float state[6]) {
    RtComPhoton* c = ((HostTechnique*)t)->tech.get();
    state[0] = c->mPhotonRadius; state[1] = c->mClampingValue; state[2] = c->mPrecomptedPdfMc; state[3] = c->mVslRadius;
    state[4] = c->mVslInvPiRadius2; state[5] = (float)c->numIterations();
    return 0;
}

// resolve like the display path (runFinalProgram(param, param, 1, gamma)); rows bottom-up
int evplp_host_technique_final(void* t, float vplScale, float photonScale, float lightScale, int gamma, float* hostRGB) {
    GUARD(
        FloatImage img = ((HostTechnique*)t)->tech->runFinalProgram(vplScale, photonScale, lightScale, gamma != 0);
        memcpy(hostRGB, img.data(), sizeof(float) * 3 * img.width() * img.height());
        return 0;
    )
}

void evplp_host_technique_destroy(void* t) { delete (HostTechnique*)t; }

// RtPt2 stepped one iteration at a time; `ptJson` is the text of the "pt" object.
struct HostPt { std::unique_ptr<RtPt2> tech; shared_ptr<RtScene> scene; };
void* evplp_host_pt_create(void* s, const char* ptJson, int resX, int resY, int device, int rank, int worldSize) {
    try {
        HostScene* hs = (HostScene*)s;
        auto* hp = new HostPt();
        hp->scene = hs->scene;
        hp->tech.reset(new RtPt2(device));
        hp->tech->setPartition(rank, worldSize);
        hp->tech->setWriteOutputs(false);
        Vec2 res; res.x = (float)resX; res.y = (float)resY;
        hp->tech->parse(hp->scene, res, Json::parse(ptJson));
        hp->tech->setup();
        return hp;
    } catch (const std::exception& e) { g_hostErr = e.what(); return nullptr; }
}
int evplp_host_pt_iterate(void* t) { GUARD(return ((HostPt*)t)->tech->iterate() ? 1 : 0;) }
int evplp_host_pt_final(void* t, float ptScale, float lightScale, int gamma, float* hostRGB) {
    GUARD(
        FloatImage img = ((HostPt*)t)->tech->runFinalProgram(ptScale, lightScale, gamma != 0);
        memcpy(hostRGB, img.data(), sizeof(float) * 3 * img.width() * img.height());
        return 0;
    )
}
void* evplp_host_pt_handle(void* t) { return ((HostPt*)t)->tech->handle(); }
void evplp_host_pt_destroy(void* t) { delete (HostPt*)t; }

// The whole blocking call of the reference: RtTechnique::render(scene, resolution, json) + outputs.
int evplp_host_render_json(const char* jsonPath, int device) {
    GUARD(
        Json json = Json::parse_file(jsonPath);
        shared_ptr<RtScene> scene = LoadScene(json, jsonPath);
        if (!scene) throw std::runtime_error("no scene in json");
        Vec2 res; res.x = json["resX"].as_float(); res.y = json["resY"].as_float();
        if (!json["pt"].is_null()) { RtPt2 t(device); t.render(scene, res, json["pt"]); }
        if (!json["photonfam"].is_null()) { RtComPhoton t(device); t.render(scene, res, json["photonfam"]); }
        if (!json["lvcphotonfam"].is_null()) { RtLvcComPhoton t(device); t.render(scene, res, json["lvcphotonfam"]); }
        return 0;
    )
}

// host-only taps (no device): the progressive schedule and the jitter stream of RtComPhoton
void evplp_host_progressive_update(int32_t numIterations, float alpha, float clampingStart, uint32_t numVplLightPaths,
                                   uint32_t numLightPaths, int32_t forceVsl, float* state /* radius, clamp, pdfMc, vslRadius, vslInvPiR2 */) {
    RtComPhoton::ProgressiveUpdate(numIterations, alpha, clampingStart, numVplLightPaths, numLightPaths, forceVsl != 0, &state[0],
                                   &state[1], &state[2], &state[3], &state[4]);
}
void evplp_host_jitter_stream(uint32_t rngOffset, uint32_t numIterations, float* out) {
    IndependentSampler s(rngOffset);
    for (uint32_t i = 0; i < numIterations; i++) { Vec2 v = s.nextVec2(); out[2 * i] = v.x; out[2 * i + 1] = v.y; }
}
int evplp_host_save_pfm(const char* path, const float* rgbTopDown, int w, int h) {
    GUARD(FloatImage f((size_t)w, (size_t)h); memcpy(f.data(), rgbTopDown, sizeof(float) * 3 * (size_t)w * h); FloatImage::Save(f, path); return 0;)
}

float evplp_host_pfm_relmse(const char* a, const char* b) {
    try { return FloatImage::ComputeRelMse(FloatImage::LoadPFM(a), FloatImage::LoadPFM(b)); }
    catch (const std::exception& e) { g_hostErr = e.what(); return -1.f; }
}
// masked metrics (scene/conference/README.md): `maskPng` is read through the PNG decoder, first channel / 255 = weight; PFM files
// are stored bottom-up and loaded top-down, the mask is top-down like the PNG.  relative = 0: MSE, 1: relMSE.  -1 on error.
float evplp_host_pfm_masked_error(const char* a, const char* b, const char* maskPng, int relative) {
    try {
        const FloatImage ia = FloatImage::LoadPFM(a), ib = FloatImage::LoadPFM(b);
        const png::Image m = png::DecodeFile(maskPng, 1);
        std::vector<float> w(m.pixels.size());
        for (size_t i = 0; i < w.size(); i++) w[i] = (float)m.pixels[i] / 255.0f;
        return relative ? FloatImage::ComputeRelMse(ia, ib, w) : FloatImage::ComputeMse(ia, ib, w);
    } catch (const std::exception& e) { g_hostErr = e.what(); return -1.f; }
}


// Host-only configuration check: reads a scene JSON file exactly as main.cpp / LoadScene / the techniques' parse() do, but
// against an already built scene (the OBJ files a reference JSON names may be absent), and reports what was understood.
// out[0..2] = sections present (pt, photonfam, lvcphotonfam); out[3..4] = resX, resY; out[5..13] = camera origin, look-at,
// up; out[14] = fovy (radians); per family section f (base 16 for photonfam, 40 for lvcphotonfam): numLightPaths,
// numVplLightPaths, numMaxBounces, radiusPercentage, misMode, DoProgressive, AlphaProgressive, forceVsl, vplSplat on,
// photonSplat on, frameMode, rngOffset, numMaxIteration, timeLimitMs, clampingValue, vslRadiusPercentage, useJitter;
// pt (base 64): rngOffset, numMaxIteration, timeLimitMs, numMaxBounces, numSamplePerPixel, frameMode, useJitter.
int evplp_host_config_check(void* s, const char* jsonPath, double out[80]) {
    GUARD(
        HostScene* hs = (HostScene*)s;
        std::ifstream ifs(jsonPath);
        if (!ifs.is_open()) throw std::runtime_error(std::string("cannot open ") + jsonPath);
        std::stringstream buf; buf << ifs.rdbuf();
        const Json json = Json::parse(buf.str());
        for (int i = 0; i < 80; i++) out[i] = 0.0;
        Vec2 res; res.x = json.at("resX").as_float(); res.y = json.at("resY").as_float();
        out[3] = res.x; out[4] = res.y;
        const Json& cj = json.contains("camera") ? json["camera"] : json.at("stablecamera");
        RtStableCamera cam(cj, res.x / res.y);
        out[5] = cam.mOrigin.x; out[6] = cam.mOrigin.y; out[7] = cam.mOrigin.z;
        out[8] = cam.mLookAt.x; out[9] = cam.mLookAt.y; out[10] = cam.mLookAt.z;
        out[11] = cam.mUp.x; out[12] = cam.mUp.y; out[13] = cam.mUp.z; out[14] = cam.mFovy;
        const char* fam[2] = {"photonfam", "lvcphotonfam"};
        for (int f = 0; f < 2; f++) {
            if (!json.contains(fam[f]) || json[fam[f]].is_null()) continue;
            out[1 + f] = 1.0;
            std::unique_ptr<RtComPhoton> t(f ? new RtLvcComPhoton(0) : new RtComPhoton(0));
            t->parse(hs->scene, res, json[fam[f]]);
            double* o = out + 16 + 24 * f;
            o[0] = t->mNumLightPaths; o[1] = t->mNumVplLightPaths; o[2] = t->mNumMaxBounce; o[3] = t->mRadiusPercentage;
            o[4] = (double)t->mMisMode; o[5] = t->mDoProgressive; o[6] = t->mAlphaProgressive; o[7] = t->mForceVsl;
            o[8] = t->mDoVplSplat; o[9] = t->mDoPhotonSplat; o[10] = (double)t->mFrameMode; o[11] = t->mRngOffset;
            o[12] = t->mNumMaxIteration; o[13] = t->mTimelimitMs; o[14] = t->mClampingValue; o[15] = t->mVslRadiusPercentage;
            o[16] = t->mJitter;
        }
        if (json.contains("pt") && !json["pt"].is_null()) {
            out[0] = 1.0;
            RtPt2 t(0);
            t.parse(hs->scene, res, json["pt"]);
            double* o = out + 64;
            o[0] = t.mRngOffset; o[1] = t.mNumMaxIteration; o[2] = t.mTimelimitMs; o[3] = t.mNumMaxBounce; o[4] = t.mNumSamplePerPixel;
            o[5] = (double)t.mFrameMode; o[6] = t.mJitter;
        }
        return 0;)
}

#endif  // !EVPLP_SCENE_ONLY

// texture decoding taps: a JPEG byte stream -> top-down RGB8 (what stbi_load(path, .., 3) returns without the
// flip), and a texture file -> the RGBA32F texels RtTexture hands to evplp_upload_scene
int evplp_host_jpeg_info(const uint8_t* data, uint64_t n, int32_t* width, int32_t* height, int32_t* fileChannels) {
    GUARD(jpeg::Image img; jpeg::Decode(data, (size_t)n, &img); *width = img.width; *height = img.height;
          *fileChannels = img.fileChannels; return 0;)
}
int evplp_host_jpeg_decode(const uint8_t* data, uint64_t n, uint8_t* rgbOut, uint64_t rgbCapacity) {
    GUARD(jpeg::Image img; jpeg::Decode(data, (size_t)n, &img);
          if (img.rgb.size() > rgbCapacity) throw std::runtime_error("evplp_host_jpeg_decode: output buffer too small");
          memcpy(rgbOut, img.rgb.data(), img.rgb.size()); return 0;)
}
// RealTime::loop contract probe (common/realtime.h:100-141): beforeSwap returns true `beforeTrue` times, afterSwap returns true
// `afterTrue` times, the sink reports "window closed" once `closeAfter` frames were presented (< 0: never).
// out = {loop passes, frames presented, afterSwap calls}
void evplp_host_realtime_probe(int beforeTrue, int afterTrue, int closeAfter, uint64_t out[3]) {
    struct Probe : RealTime::Sink {
        int closeAfter; uint64_t presented = 0;
        void present(uint64_t) override { presented++; }
        bool shouldClose() override { return closeAfter >= 0 && presented >= (uint64_t)closeAfter; }
    } probe;
    probe.closeAfter = closeAfter;
    RealTime rt(&probe);
    int b = 0, a = 0;
    uint64_t afterCalls = 0;
    rt.loop([&](std::string*) { return b++ < beforeTrue; }, [&](std::string*) { afterCalls++; return a++ < afterTrue; });
    out[0] = rt.loopPasses(); out[1] = rt.framesPresented(); out[2] = afterCalls;
}

// a PNG byte stream -> top-down samples with `wantChannels` channels (0 = the file's), like stbi_load_from_memory(.., want)
int evplp_host_png_info(const uint8_t* data, uint64_t n, int32_t* width, int32_t* height, int32_t* fileChannels) {
    GUARD(png::Image img = png::Decode(data, (size_t)n, 0); *width = img.width; *height = img.height;
          *fileChannels = img.fileChannels; return 0;)
}
int evplp_host_png_decode(const uint8_t* data, uint64_t n, int wantChannels, uint8_t* out, uint64_t capacity) {
    GUARD(png::Image img = png::Decode(data, (size_t)n, wantChannels);
          if (img.pixels.size() > capacity) throw std::runtime_error("evplp_host_png_decode: output buffer too small");
          memcpy(out, img.pixels.data(), img.pixels.size()); return 0;)
}
int evplp_host_texture_load(const char* path, float gamma, int32_t* width, int32_t* height, float* rgbaOut, uint64_t capacityFloats) {
    GUARD(RtTexture t(std::string(path), gamma); *width = t.mWidth; *height = t.mHeight;
          if (rgbaOut) {
              if (t.mData.size() > capacityFloats) throw std::runtime_error("evplp_host_texture_load: output buffer too small");
              memcpy(rgbaOut, t.mData.data(), t.mData.size() * sizeof(float));
          }
          return 0;)
}

}  // extern "C"
