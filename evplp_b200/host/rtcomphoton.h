// rtcomphoton.h -- RtComPhoton / RtLvcComPhoton: the EVPLP technique, headless, over the C ABI.
//
// Mirrors reflectcuts/realtimetechniques/rtcomphoton/rtcomphoton.h (and rtlvccomphoton.h):
// same JSON keys (:114-218), same per-iteration order and stage names (:936-1068), same
// progressive schedule (:1033-1063), same three PFM outputs + stat JSON (:1107-1132).  The
// OptiX programs / GL passes behind each run*() member are the sm_100a kernels of
// libevplp_b200.so.  New here (the reference is single-GPU): setPartition(rank, worldSize)
// renders iterations k = rank (mod worldSize) and evplp_reduce() sums the accumulation layers.
#pragma once
#include <chrono>
#include <cmath>
#include <functional>
#include <iomanip>
#include <random>
#include "floatimage.h"
#include "realtime.h"
#include "rttechnique.h"

namespace evplp_host {

// IndependentSampler over std::mt19937 (sampler/independent.h:37-40, common/rng.h:14-40).
// nextFloat() = uniform_real_distribution<float>(0,1): defined here as float(u32) * 2^-32 with the
// "== 1" guard; nextVec2() draws x first (both unpinned in the reference).
class IndependentSampler {
public:
    explicit IndependentSampler(uint32_t seed) : mGen(seed) {}
    float nextFloat() {
        float u = (float)mGen() * 2.3283064365386963e-10f;
        if (u >= 1.0f) u = std::nextafter(1.0f, 0.0f);
        return u;
    }
    Vec2 nextVec2() { Vec2 v; v.x = nextFloat(); v.y = nextFloat(); return v; }
private:
    std::mt19937 mGen;
};

class StopWatch {  // common/stopwatch.h:6-30
public:
    StopWatch() { reset(); }
    void reset() { mStart = std::chrono::steady_clock::now(); }
    float timeMilliSec() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - mStart).count(); }
private:
    std::chrono::steady_clock::time_point mStart;
};

class RtComPhoton : public RtTechnique {
public:
    enum EFrame { ClearEveryFrame = 0, Accumulate = 1 };
    enum EMis { One = 0, Balance, Max, Power2, GeometryClamp, GeometryBrdfClamp };

    explicit RtComPhoton(int device = 0, int gatherMode = EVPLP_GATHER_VPL) : mDevice(device), mBaseGatherMode(gatherMode) {}
    ~RtComPhoton() override { destroy(); }

    // Multi-GPU partition (new; the reference is single-GPU).  PartitionIterations: rank g renders the iterations
    // k = g (mod N) end to end (progressive / accumulate runs).  PartitionImage: every rank renders every iteration but
    // only its interleaved 16-row bands of the gather and its contiguous range of light paths of the splat (one heavy frame).
    enum EPartition { PartitionIterations = 0, PartitionImage = 1 };
    void setPartition(int rank, int worldSize, EPartition mode = PartitionIterations) { mRank = rank; mWorldSize = worldSize; mPartition = mode; }
    void setNcclComm(void* comm) { mNcclComm = comm; }
    void setWriteOutputs(bool w) { mWriteOutputs = w; }

    static EMis misFromString(const std::string& s) {
        static const std::map<std::string, EMis> m = {{"one", One}, {"balance", Balance}, {"max", Max}, {"power2", Power2},
                                                      {"geometryClamp", GeometryClamp}, {"geometryBrdfClamp", GeometryBrdfClamp}};
        auto it = m.find(s);
        if (it == m.end()) throw std::runtime_error("unknown misMode " + s);
        return it->second;
    }

    // rtcomphoton.h:107-223
    void parse(shared_ptr<RtScene>& scene, const Vec2& resolution, const Json& json) {
        mScene = scene;
        mResolution = resolution;
        mInvResolution.x = 1.0f / resolution.x; mInvResolution.y = 1.0f / resolution.y;
        mNumLightPaths = (uint32_t)json.at("numLightPaths").as_int();
        mNumVplLightPaths = (uint32_t)json.at("numVplLightPaths").as_int();
        mNumMaxBounce = (uint32_t)json.at("numMaxBounces").as_int();
        mNumPhotonsPerLightPath = mNumMaxBounce + 1;
        mRadiusPercentage = json.at("radiusPercentage").as_float();
        mPhotonRadius = mScene->findBoundingSphereRadius() * mRadiusPercentage;
        mPrecomptedPdfMc = static_cast<float>(mNumVplLightPaths) / static_cast<float>(mNumLightPaths) * Math::InvPi / (mPhotonRadius * mPhotonRadius);
        mDoWriteEveryFrame = json.contains("writeEveryFrame") ? json["writeEveryFrame"].as_bool() : false;
        mNumMaxIteration = json.at("numMaxIteration").as_int();
        mTimelimitMs = json.at("timeLimitMs").as_float();
        const std::string frameMode = json.at("frameMode").as_string();
        if (frameMode == "accumulate") mFrameMode = Accumulate;
        else if (frameMode == "cleareveryframe") mFrameMode = ClearEveryFrame;
        else throw std::runtime_error("unknown frameMode " + frameMode);
        mMisMode = json.contains("misMode") ? misFromString(json["misMode"].as_string()) : Balance;
        if (json.contains("clampingStart")) throw std::runtime_error("clampingStart option is not use anymore; remove it from your JSON file");
        if (json.contains("targetRenderingTime")) mTargetRenderingTime = json["targetRenderingTime"].as_float();
        if (!json.contains("clampingCoeff")) {
            float totalArea = scene->totalArea();
            std::cout << "Total area computation: " << totalArea << "\n";
            mClampingValue = 1.f / totalArea;
            mClampingStart = 1.f / totalArea;
        } else {
            float clampCoeff = json["clampingCoeff"].as_float();
            mClampingValue = clampCoeff;
            mClampingStart = clampCoeff;
        }
        mRngOffset = (uint32_t)json.at("rngOffset").as_int();
        mDumpCombineFilename = json.at("combinedFilename").as_string();
        mDumpWeightedPhotonFilename = json.at("weightedPhotonFilename").as_string();
        mDumpWeightedVplFilename = json.at("weightedVplFilename").as_string();
        mStatFilename = json.at("statFilename").as_string();
        mJitter = json.at("useJitter").as_bool();
        mUseStat = json.at("useStat").as_bool();
        if (json.contains("DoProgressive")) mDoProgressive = json["DoProgressive"].as_bool();
        if (json.contains("AlphaProgressive")) mAlphaProgressive = json["AlphaProgressive"].as_float();
        if (json.contains("run")) {
            const Json& r = json["run"];
            if (r.contains("deferredShading")) mDoDeferredShading = r["deferredShading"].as_bool();
            if (r.contains("lightTracing")) mDoLightTracing = r["lightTracing"].as_bool();
            if (r.contains("vplSplat")) mDoVplSplat = r["vplSplat"].as_bool();
            if (r.contains("photonSplat")) mDoPhotonSplat = r["photonSplat"].as_bool();
            if (r.contains("lightRender")) mDoLightRender = r["lightRender"].as_bool();
            if (r.contains("finalize")) mDoFinalize = r["finalize"].as_bool();
        }
        if (mNumVplLightPaths == 0) {
            std::cout << "WARN: 0 VPL light paths. Disable mDoVplSplat\n";
            mDoVplSplat = false;
        }
        if (json.contains("forceVsl")) {
            mForceVsl = json["forceVsl"].as_bool();
            if (mForceVsl) {
                mVslRadiusPercentage = json.at("vslRadiusPercentage").as_float();
                mVslRadius = mScene->findBoundingSphereRadius() * mVslRadiusPercentage;
                if (mVslRadius <= 0.008) {
                    mVslRadius = std::max(mVslRadius, 0.008f);
                    std::cout << "warning : vslRadius is too small. clamped vslRadius" << std::endl;
                }
                mVslInvPiRadius2 = Math::InvPi / (mVslRadius * mVslRadius);
            }
        }
    }

    void render(shared_ptr<RtScene>& scene, const Vec2& resolution, const Json& json) override {
        parse(scene, resolution, json);
        setup();
        run();
        destroy();
    }

    // rtcomphoton.h:646-708: context, scene upload, acceleration structure
    void setup() {
        check(evplp_create(mDevice, (int)mResolution.x, (int)mResolution.y, &mHandle), "evplp_create");
        check(mScene->upload(mHandle), "evplp_upload_scene");
        check(evplp_build_bvh(mHandle), "evplp_build_bvh");
        mScene->mCamera->basis(&mCamF, &mCamS, &mCamU, &mTanHalfX, &mTanHalfY);
        mMainSampler.reset(new IndependentSampler(mRngOffset));
        mNumIterations = 0;
        check(evplp_clear_accum(mHandle), "evplp_clear_accum");
        const bool imageMode = mPartition == PartitionImage && mWorldSize > 1;
        // cleareveryframe shows ONE frame; dealing iterations round-robin would leave a different "last frame" on every rank
        // and the reduce would add them up.  A single frame is split over the ranks by image tiles / light-path ranges instead.
        if (mFrameMode == ClearEveryFrame && mWorldSize > 1 && !imageMode)
            throw std::runtime_error("frameMode cleareveryframe with several ranks needs the image partition (PartitionImage)");
        check(evplp_set_option(mHandle, "gather_band_stride", imageMode ? mWorldSize : 0), "evplp_set_option");
        check(evplp_set_option(mHandle, "gather_band_offset", imageMode ? mRank : 0), "evplp_set_option");
        mMasterWatch.reset();
        mPrevTiming = 0.f;
    }

    // ---- the per-iteration stages (reference member names) ----
    void runDeferredProgram() { check(evplp_gbuffer(mHandle), "evplp_gbuffer"); }                     // :710-754
    void runOptixLightTracingProgram(uint32_t rngSeed) {                                                // :869-881
        check(evplp_light_trace(mHandle, rngSeed, 0, mNumLightPaths), "evplp_light_trace");
    }
    void runOptixVplProgram() {                                                                         // :857-867
        const int mode = mForceVsl ? EVPLP_GATHER_VSL : mBaseGatherMode;
        check(evplp_vpl_gather(mHandle, nullptr, mode), "evplp_vpl_gather");
    }
    void runPhotonSplat() {                                                                             // :789-837
        uint64_t firstPath = 0, numPaths = mNumLightPaths;
        if (mPartition == PartitionImage && mWorldSize > 1) {  // this rank splats its contiguous range of light paths
            firstPath = (uint64_t)mNumLightPaths * (uint64_t)mRank / (uint64_t)mWorldSize;
            numPaths = (uint64_t)mNumLightPaths * (uint64_t)(mRank + 1) / (uint64_t)mWorldSize - firstPath;
        }
        check(evplp_photon_splat(mHandle, firstPath * mNumPhotonsPerLightPath, numPaths * mNumPhotonsPerLightPath, nullptr),
              "evplp_photon_splat");
    }
    void runLightProgram() { check(evplp_light_pass(mHandle), "evplp_light_pass"); }                    // :839-855

    // Photon counts whose record buffer (96 B x (B+1) per path) would not fit in HBM (BASELINE config 5: up to 2^28 paths =
    // 103 GB) are streamed: the VPL prefix is traced and gathered first, then the light paths are traced and splatted in
    // chunks of mMaxPathsPerTrace.  Results are identical to the one-pass order (every path keeps its own RNG stream,
    // lighttracing.cu:202-203, and the fixed-point splat is order-independent).
    void setMaxPathsPerTrace(uint64_t n) { mMaxPathsPerTrace = n; }
    void runStreamed(uint32_t rngSeed) {
        if (!mDoLightTracing) return;
        if (mBaseGatherMode == EVPLP_GATHER_LVC && mDoVplSplat)
            throw std::runtime_error("the LVC gather reads every light path and cannot be streamed; lower numLightPaths");
        if (mDoVplSplat) {
            if ((uint64_t)mNumVplLightPaths > mMaxPathsPerTrace) throw std::runtime_error("numVplLightPaths exceeds the streaming chunk");
            check(evplp_light_trace(mHandle, rngSeed, 0, mNumVplLightPaths), "evplp_light_trace");
            runOptixVplProgram();
        }
        if (!mDoPhotonSplat) return;
        uint64_t p0 = 0, p1 = mNumLightPaths;
        if (mPartition == PartitionImage && mWorldSize > 1) {
            p0 = (uint64_t)mNumLightPaths * (uint64_t)mRank / (uint64_t)mWorldSize;
            p1 = (uint64_t)mNumLightPaths * (uint64_t)(mRank + 1) / (uint64_t)mWorldSize;
        }
        for (uint64_t c = p0; c < p1; c += mMaxPathsPerTrace) {
            const uint64_t n = std::min<uint64_t>(mMaxPathsPerTrace, p1 - c);
            check(evplp_light_trace(mHandle, rngSeed, (uint32_t)c, (uint32_t)n), "evplp_light_trace");
            check(evplp_photon_splat(mHandle, 0, n * mNumPhotonsPerLightPath, nullptr), "evplp_photon_splat");
        }
    }
    FloatImage runFinalProgram(float vplScale, float photonScale, float lightScale, bool gamma) {        // :756-787 + dumpImage :225-249
        FloatImage img((size_t)mResolution.x, (size_t)mResolution.y);
        check(evplp_resolve(mHandle, vplScale, photonScale, lightScale, gamma ? 1 : 0, img.data()), "evplp_resolve");
        return img;  // rows bottom-up, like glReadPixels
    }

    void pushParams(const Vec2& jitter, uint32_t rngSeed) {  // the rtContext[...]->set* block, :895-930, 873
        EvplpParams P;
        memset(&P, 0, sizeof(P));
        const Vec3 o = mScene->mCamera->getOrigin();
        P.cameraPosition[0] = o.x; P.cameraPosition[1] = o.y; P.cameraPosition[2] = o.z;
        P.camForward[0] = mCamF.x; P.camForward[1] = mCamF.y; P.camForward[2] = mCamF.z;
        P.camRight[0] = mCamS.x; P.camRight[1] = mCamS.y; P.camRight[2] = mCamS.z;
        P.camUp[0] = mCamU.x; P.camUp[1] = mCamU.y; P.camUp[2] = mCamU.z;
        P.tanHalfFovX = mTanHalfX; P.tanHalfFovY = mTanHalfY;
        P.jitter[0] = jitter.x; P.jitter[1] = jitter.y;
        P.nearDist = 0.1f; P.farDist = 100.0f;
        P.numLightPaths = mNumLightPaths; P.numVplLightPaths = mNumVplLightPaths; P.numPhotonsPerLightPath = mNumPhotonsPerLightPath;
        P.radius = mPhotonRadius; P.pdfMc = mPrecomptedPdfMc; P.misMode = (uint32_t)mMisMode; P.clampingValue = mClampingValue;
        P.doAccumulate = mFrameMode == ClearEveryFrame ? 0u : 1u;
        P.vslRadius = mVslRadius; P.vslInvPiRadius2 = mVslInvPiRadius2;
        P.rngSeed = rngSeed;
        check(evplp_set_params(mHandle, &P), "evplp_set_params");
    }

    // The progressive schedule (rtcomphoton.h:1033-1063), evaluated after numIterations++.  Types as in the
    // reference: float ratio (int + float), float sqrt, pow(int, float) in double, product rounded to float.
    static void ProgressiveUpdate(int numIterations, float alphaProgressive, float clampingStart, uint32_t numVplLightPaths,
                                  uint32_t numLightPaths, bool forceVsl, float* photonRadius, float* clampingValue,
                                  float* pdfMc, float* vslRadius, float* vslInvPiRadius2) {
        float ratio = (numIterations + alphaProgressive) / (numIterations + 1);
        *photonRadius *= std::sqrt(ratio);
        *clampingValue = (float)(clampingStart * std::pow((double)numIterations, (double)alphaProgressive));
        *pdfMc = static_cast<float>(numVplLightPaths) / static_cast<float>(numLightPaths) * Math::InvPi / (*photonRadius * *photonRadius);
        if (forceVsl) {
            *vslRadius *= std::sqrt(ratio);
            if (*vslRadius <= 0.008f) *vslRadius = std::max(*vslRadius, 0.008f);
            *vslInvPiRadius2 = Math::InvPi / (*vslRadius * *vslRadius);
        }
    }

    // What ONE pass of the loop body (rtcomphoton.h:936-1068) does on THIS rank: everything the host decides, nothing of the
    // device.  planNext() draws the iteration's jitter and reads the schedule; execute() issues the stages through the C ABI;
    // advanceSchedule() is the numIterations++ / progressive-update tail.  (Split so that the multi-rank host logic --
    // which rank renders what, with which uniforms -- can be checked without a GPU: evplp_host_technique_plan.)
    struct IterationPlan {
        int iteration = 0;                 // k
        bool render = false;               // this rank renders (its share of) iteration k
        Vec2 jitter;
        uint32_t rngSeed = 0;              // k + rngOffset
        float photonRadius = 0, clampingValue = 0, pdfMc = 0, vslRadius = 0, vslInvPiRadius2 = 0;
        uint64_t splatFirstPath = 0, splatNumPaths = 0;   // light paths whose photons this rank splats
        uint32_t tileStride = 1, tileOffset = 0;          // 8x4-pixel tiles t = offset (mod stride) this rank gathers
        bool drawLight = false, countIteration = false;
    };

    bool imageMode() const { return mPartition == PartitionImage && mWorldSize > 1; }

    IterationPlan planNext() {
        IterationPlan p;
        p.iteration = mNumIterations;
        if (mJitter) {
            Vec2 xi = mMainSampler->nextVec2();
            p.jitter.x = (2.0f * xi.x - 1.0f) * mInvResolution.x;
            p.jitter.y = (2.0f * xi.y - 1.0f) * mInvResolution.y;
        }
        p.render = imageMode() || (mNumIterations % mWorldSize) == mRank;  // iteration partition: k = rank (mod N)
        p.rngSeed = (uint32_t)mNumIterations + mRngOffset;
        p.photonRadius = mPhotonRadius; p.clampingValue = mClampingValue; p.pdfMc = mPrecomptedPdfMc;
        p.vslRadius = mVslRadius; p.vslInvPiRadius2 = mVslInvPiRadius2;
        p.splatFirstPath = 0; p.splatNumPaths = mNumLightPaths;
        if (imageMode()) {  // this rank splats its contiguous range of light paths and gathers every Nth tile
            p.splatFirstPath = (uint64_t)mNumLightPaths * (uint64_t)mRank / (uint64_t)mWorldSize;
            p.splatNumPaths = (uint64_t)mNumLightPaths * (uint64_t)(mRank + 1) / (uint64_t)mWorldSize - p.splatFirstPath;
            p.tileStride = (uint32_t)mWorldSize; p.tileOffset = (uint32_t)mRank;
        }
        p.drawLight = mDoLightRender && (!imageMode() || mRank == 0);
        p.countIteration = !imageMode() || mRank == 0;   // image partition: every rank holds a share of it, rank 0 counts it
        return p;
    }

    void execute(const IterationPlan& p) {
        pushParams(p.jitter, p.rngSeed);
        if (mFrameMode == ClearEveryFrame) check(evplp_clear_accum(mHandle), "evplp_clear_accum");
        if (mDoDeferredShading) runDeferredProgram();
        if ((uint64_t)mNumLightPaths <= mMaxPathsPerTrace) {
            if (mDoLightTracing) runOptixLightTracingProgram(p.rngSeed);
            if (mDoVplSplat) runOptixVplProgram();
            if (mDoPhotonSplat) runPhotonSplat();
        } else {
            runStreamed(p.rngSeed);
        }
        if (p.drawLight) runLightProgram();
        if (p.countIteration) check(evplp_add_iterations(mHandle, 1), "evplp_add_iterations");
    }

    void advanceSchedule() {
        mNumIterations++;
        if (mDoProgressive) {
            ProgressiveUpdate(mNumIterations, mAlphaProgressive, mClampingStart, mNumVplLightPaths, mNumLightPaths, mForceVsl,
                              &mPhotonRadius, &mClampingValue, &mPrecomptedPdfMc, &mVslRadius, &mVslInvPiRadius2);
        }
    }

    // planning without a device: what setup() prepares of the host state
    void setupHostOnly() {
        mMainSampler.reset(new IndependentSampler(mRngOffset));
        mNumIterations = 0;
        if (mFrameMode == ClearEveryFrame && mWorldSize > 1 && !imageMode())
            throw std::runtime_error("frameMode cleareveryframe with several ranks needs the image partition (PartitionImage)");
    }

    // One pass of the loop body (rtcomphoton.h:936-1068).  Returns false when the loop must stop.
    bool iterate() {
        if (mNumIterations == mNumMaxIteration) return false;
        const IterationPlan p = planNext();
        if (p.render) execute(p);
        const bool report = (mNumIterations + 1) % 20 == 0 && mRank == 0;
        const float radiusBefore = mPhotonRadius, clampBefore = mClampingValue;
        advanceSchedule();
        if (report) {
            float currentTiming = mMasterWatch.timeMilliSec();
            const float prevTimingForAdvice = mPrevTiming;
            std::cout << "numIter: " << mNumIterations << " | raduis: " << radiusBefore << " | clamping: " << clampBefore
                      << " | timing: " << currentTiming - mPrevTiming << "\n";
            mPrevTiming = currentTiming;
            // targetRenderingTime (rtcomphoton.h:1017-1030): advice printed with the 20-iteration report, never applied.
            // frameTime = the time of those 20 iterations / 20 (rtcomphoton.h:1010-1012).
            if (mTargetRenderingTime != -1.f) {
                const float frameTime = (currentTiming - prevTimingForAdvice) / 20.0f;
                const float factor = mTargetRenderingTime / frameTime;
                if (factor != 1.f) {
                    std::cout << "change number of samples: " << factor << " | currFrame time: " << frameTime << "\n";
                    if (mNumVplLightPaths != 0) {
                        const int newNbVPL = (int)(mNumVplLightPaths * factor);
                        std::cout << "Nb light paths: " << newNbVPL * (mNumLightPaths / mNumVplLightPaths) << "\n";
                        std::cout << "Nb VPL paths: " << newNbVPL << "\n";
                    } else {
                        std::cout << "Nb light paths: " << mNumLightPaths * factor << "\n";
                    }
                }
            }
        }
        // rtcomphoton.h:1065 compares unconditionally (a non-positive limit ends the run after one iteration).  With several ranks
        // each rank watches its own clock and may leave the loop at a different iteration; finish() normalises by the number
        // of iterations that were actually accumulated (reduced with the layers), so the image stays correctly exposed.
        if (mMasterWatch.timeMilliSec() >= mTimelimitMs) return false;
        return true;
    }

    // rtcomphoton.h:883-1133 (headless: RealTime::loop without the window)
    void run() {
        RealTime rt(mSink);
        rt.loop([&](std::string*) { return iterate(); },          // before swap: one iteration, false = stop
                [&](std::string* titleExtend) {                    // after swap (:1070-1104): progress title, writeEveryFrame
                    if (mNumMaxIteration > 0) {
                        const float iterRatio = (float)mNumIterations / (float)mNumMaxIteration;
                        const float time = (float)mMasterWatch.timeMilliSec() / 1000.0f;
                        const float etc = time / iterRatio - time;
                        *titleExtend = std::to_string(iterRatio * 100.0f) + "%, ETC : " + std::to_string(etc) + "s";
                    }
                    if (mDoWriteEveryFrame && mWorldSize == 1) writeFrame();
                    return true;
                });
        mLastTitle = rt.windowTitle();
        mFramesPresented = rt.framesPresented();
        finish();
    }
    // optional viewer / probe behind the headless loop (RealTime::Sink: present, shouldClose = ESC / window closed, title)
    void setSink(RealTime::Sink* sink) { mSink = sink; }

    void reduce() {
        if (mWorldSize > 1 && mNcclComm) check(evplp_reduce(mHandle, mNcclComm), "evplp_reduce");
    }

    // the output block (:1107-1132)
    void finish() {
        reduce();
        check(evplp_synchronize(mHandle), "evplp_synchronize");
        const float time = mMasterWatch.timeMilliSec();
        mElapsedMs = time;
        if (mRank != 0 || !mWriteOutputs) return;
        if (mUseStat) {
            Json result = Json::make_object();
            result.set("time", (double)time);
            result.set("numIterations", (double)mNumIterations);
            std::ofstream of(mStatFilename);
            of << result.dump(4);
        }
        const float param = (mFrameMode == ClearEveryFrame) ? 1.0f : (1.0f / (float)accumulatedIterations());
        FloatImage lightSourceImage = FloatImage::FlipY(runFinalProgram(0.0f, 0.0f, 1.0f, false));
        FloatImage photonImage = FloatImage::FlipY(runFinalProgram(0.0f, 1.0f, 0.0f, false));
        photonImage *= param;
        FloatImage vplImage = FloatImage::FlipY(runFinalProgram(1.0f, 0.0f, 0.0f, false));
        vplImage *= param;
        mCombined = lightSourceImage + vplImage + photonImage;
        FloatImage::Save(mCombined, mDumpCombineFilename);
        FloatImage::Save(lightSourceImage + vplImage, mDumpWeightedVplFilename);
        FloatImage::Save(photonImage, mDumpWeightedPhotonFilename);
    }

    // iterations accumulated into the layers (after reduce(): over all ranks); equals numIterations on a single rank
    int64_t accumulatedIterations() {
        int64_t n = 0;
        check(evplp_iterations(mHandle, &n), "evplp_iterations");
        return n > 0 ? n : 1;
    }

    void writeFrame() {  // writeEveryFrame (:1079-1102)
        const float param = (mFrameMode == ClearEveryFrame) ? 1.0f : (1.0f / (float)(mNumIterations));
        FloatImage light = runFinalProgram(0.0f, 0.0f, 1.0f, false);
        FloatImage photon = runFinalProgram(0.0f, 1.0f, 0.0f, false);
        photon *= param;
        FloatImage vpl = runFinalProgram(1.0f, 0.0f, 0.0f, false);
        vpl *= param;
        FloatImage result = light + photon + vpl;
        size_t i = mDumpWeightedPhotonFilename.find_last_of('.');
        std::string dotExtension = mDumpWeightedPhotonFilename.substr(i);
        FloatImage::Save(FloatImage::FlipY(result), mDumpWeightedPhotonFilename.substr(0, i) + "_" + std::to_string(mNumIterations) + dotExtension);
    }

    void destroy() {  // :1135-1138
        if (mHandle) { evplp_destroy(mHandle); mHandle = nullptr; }
    }

    evplp_handle handle() const { return mHandle; }
    int numIterations() const { return mNumIterations; }
    float elapsedMs() const { return mElapsedMs; }
    const FloatImage& combined() const { return mCombined; }

    // public like the reference's members so that callers (tests, bench) can override them
    uint32_t mNumLightPaths = 0, mNumVplLightPaths = 0, mNumMaxBounce = 0, mNumPhotonsPerLightPath = 0, mRngOffset = 0;
    float mRadiusPercentage = 0, mPhotonRadius = 0, mPrecomptedPdfMc = 0, mClampingValue = 0, mClampingStart = 0;
    float mTimelimitMs = 0, mTargetRenderingTime = -1, mAlphaProgressive = 0.7f;
    float mVslRadiusPercentage = 0, mVslRadius = 0, mVslInvPiRadius2 = 0;
    int mNumMaxIteration = -1;
    EFrame mFrameMode = Accumulate;
    EMis mMisMode = Balance;
    bool mJitter = true, mUseStat = false, mDoProgressive = false, mForceVsl = false, mDoWriteEveryFrame = false;
    bool mDoDeferredShading = true, mDoLightTracing = true, mDoVplSplat = true, mDoPhotonSplat = true, mDoLightRender = true, mDoFinalize = true;
    std::string mDumpCombineFilename, mDumpWeightedPhotonFilename, mDumpWeightedVplFilename, mStatFilename;

protected:
    void check(int rc, const char* what) {
        if (rc != EVPLP_OK) throw std::runtime_error(std::string(what) + ": " + evplp_last_error());
    }
    int mDevice = 0, mBaseGatherMode = EVPLP_GATHER_VPL, mRank = 0, mWorldSize = 1;
    EPartition mPartition = PartitionIterations;
    uint64_t mMaxPathsPerTrace = 32ull << 20;  // 32 Mi paths = 12.9 GB of records per chunk (180 GB of HBM per GPU)
    void* mNcclComm = nullptr;
    bool mWriteOutputs = true;
    evplp_handle mHandle = nullptr;
    shared_ptr<RtScene> mScene;
    Vec2 mResolution, mInvResolution;
    Vec3 mCamF, mCamS, mCamU;
    float mTanHalfX = 0, mTanHalfY = 0;
    std::unique_ptr<IndependentSampler> mMainSampler;
    int mNumIterations = 0;
    StopWatch mMasterWatch;
    RealTime::Sink* mSink = nullptr;
    std::string mLastTitle;
    uint64_t mFramesPresented = 0;
    float mPrevTiming = 0, mElapsedMs = 0;
    FloatImage mCombined;
};

// RtLvcComPhoton (rtlvccomphoton.h + lvclighttracing.cu:348-387): the same technique whose gather
// picks, per pixel, a random window of numVplLightPaths paths out of all numLightPaths.
class RtLvcComPhoton : public RtComPhoton {
public:
    explicit RtLvcComPhoton(int device = 0) : RtComPhoton(device, EVPLP_GATHER_LVC) {}
};

}  // namespace evplp_host
