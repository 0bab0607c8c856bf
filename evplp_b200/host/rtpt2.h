// rtpt2.h -- RtPt2: the reference's unidirectional path tracer (rtpt/rtpt2.h + pathtracing.cu), headless, over the C ABI.
// It is the reference's own ground-truth generator (scene/*/*_pt.json); here it doubles as an independent convergence
// check of the EVPLP path.  Same JSON keys (rtpt2.h:91-111), same loop (:608-665) and output (:704-721):
// result = light + pt / numIterations, flipped and saved as PFM.  Partition over GPUs as in RtComPhoton (iterations).
#pragma once
#include "rtcomphoton.h"

namespace evplp_host {

class RtPt2 : public RtTechnique {
public:
    enum EFrame { ClearEveryFrame = 0, Accumulate = 1 };
    explicit RtPt2(int device = 0) : mDevice(device) {}
    ~RtPt2() override { destroy(); }
    void setPartition(int rank, int worldSize) { mRank = rank; mWorldSize = worldSize; }
    void setWriteOutputs(bool w) { mWriteOutputs = w; }

    void parse(shared_ptr<RtScene>& scene, const Vec2& resolution, const Json& json) {  // rtpt2.h:86-112
        mScene = scene;
        mResolution = resolution;
        mInvResolution.x = 1.0f / resolution.x; mInvResolution.y = 1.0f / resolution.y;
        mRngOffset = (uint32_t)json.at("rngOffset").as_int();
        mNumMaxIteration = json.at("numMaxIteration").as_int();
        mTimelimitMs = json.at("timeLimitMs").as_float();
        const std::string frameMode = json.at("frameMode").as_string();
        if (frameMode == "accumulate") mFrameMode = Accumulate;
        else if (frameMode == "cleareveryframe") mFrameMode = ClearEveryFrame;
        else throw std::runtime_error("unknown frameMode " + frameMode);
        mOutputFilename = json.at("outputFilename").as_string();
        mStatFilename = json.at("statFilename").as_string();
        mJitter = json.at("useJitter").as_bool();
        mUseStat = json.at("useStat").as_bool();
        mNumSamplePerPixel = json.at("numSamplePerPixel").as_int();  // read but unused by the reference (pathtracing.cu:246 fixes 1)
        mNumMaxBounce = (uint32_t)json.at("numMaxBounces").as_int();
        mDoWriteEveryFrame = json.contains("writeEveryFrame") ? json["writeEveryFrame"].as_bool() : false;
    }

    void render(shared_ptr<RtScene>& scene, const Vec2& resolution, const Json& json) override {
        parse(scene, resolution, json);
        setup();
        RealTime rt(mSink);   // rtpt2.h:608-692: the same loop contract; after swap = time limit (checked in iterate()) + writeEveryFrame
        rt.loop([&](std::string*) { return iterate(); },
                [&](std::string*) {
                    if (mDoWriteEveryFrame && mWorldSize == 1) writeFrame();
                    return true;
                });
        finish();
        destroy();
    }

    void setup() {
        // iterations are dealt round-robin to the ranks and summed: a frame mode that keeps only the last frame has no meaning there
        if (mFrameMode == ClearEveryFrame && mWorldSize > 1)
            throw std::runtime_error("frameMode cleareveryframe cannot be split over several ranks by iterations");
        check(evplp_create(mDevice, (int)mResolution.x, (int)mResolution.y, &mHandle), "evplp_create");
        check(mScene->upload(mHandle), "evplp_upload_scene");
        check(evplp_build_bvh(mHandle), "evplp_build_bvh");
        mScene->mCamera->basis(&mCamF, &mCamS, &mCamU, &mTanHalfX, &mTanHalfY);
        mMainSampler.reset(new IndependentSampler(mRngOffset));
        mNumIterations = 0;
        check(evplp_clear_accum(mHandle), "evplp_clear_accum");
        mMasterWatch.reset();
    }

    bool iterate() {  // rtpt2.h:608-665
        if (mNumIterations == mNumMaxIteration) return false;
        Vec2 jitter;
        if (mJitter) {
            Vec2 xi = mMainSampler->nextVec2();
            jitter.x = (2.0f * xi.x - 1.0f) * mInvResolution.x;
            jitter.y = (2.0f * xi.y - 1.0f) * mInvResolution.y;
        }
        if ((mNumIterations % mWorldSize) == mRank) {
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
            P.numPhotonsPerLightPath = mNumMaxBounce + 1;
            P.doAccumulate = mFrameMode == ClearEveryFrame ? 0u : 1u;
            P.rngSeed = (uint32_t)mNumIterations + mRngOffset;
            check(evplp_set_params(mHandle, &P), "evplp_set_params");
            if (mFrameMode == ClearEveryFrame) check(evplp_clear_accum(mHandle), "evplp_clear_accum");
            check(evplp_gbuffer(mHandle), "evplp_gbuffer");                                // runDeferredProgram
            check(evplp_path_trace(mHandle, nullptr, mNumMaxBounce), "evplp_path_trace");  // runOptixPtProgram
            check(evplp_light_pass(mHandle), "evplp_light_pass");                          // runLightProgram
            check(evplp_add_iterations(mHandle, 1), "evplp_add_iterations");               // what finish() normalises by
        }
        mNumIterations++;
        if (mMasterWatch.timeMilliSec() >= mTimelimitMs) return false;   // unconditional, like rtpt2.h
        return true;
    }

    FloatImage runFinalProgram(float ptScaling, float lightScaling, bool gamma) {  // rtpt2.h:507-559
        FloatImage img((size_t)mResolution.x, (size_t)mResolution.y);
        check(evplp_resolve(mHandle, ptScaling, 0.0f, lightScaling, gamma ? 1 : 0, img.data()), "evplp_resolve");
        return img;
    }

    void finish() {  // rtpt2.h:690-721
        check(evplp_synchronize(mHandle), "evplp_synchronize");
        mElapsedMs = mMasterWatch.timeMilliSec();
        if (mRank != 0 || !mWriteOutputs) return;
        if (mUseStat) {
            Json result = Json::make_object();
            result.set("time", (double)mElapsedMs);
            result.set("numIterations", (double)mNumIterations);
            std::ofstream of(mStatFilename);
            of << result.dump(4);
        }
        FloatImage result;
        if (mFrameMode == ClearEveryFrame) {
            result = runFinalProgram(1.0f, 1.0f, false);
        } else {
            FloatImage lightImage = runFinalProgram(0.0f, 1.0f, false);
            FloatImage ptImage = runFinalProgram(1.0f, 0.0f, false);
            int64_t accumulated = 0;   // iterations in the layer (over all ranks once reduced); = mNumIterations on one rank
            check(evplp_iterations(mHandle, &accumulated), "evplp_iterations");
            ptImage *= 1.0f / (float)(accumulated > 0 ? accumulated : 1);  // the reference divides (FloatImage::operator/=); same up to 1 ulp
            result = lightImage + ptImage;
        }
        FloatImage::Save(FloatImage::FlipY(result), mOutputFilename);
    }

    void writeFrame() {  // writeEveryFrame (rtpt2.h:669-689): <output>_<numIterations>.<ext>
        check(evplp_synchronize(mHandle), "evplp_synchronize");
        FloatImage result;
        if (mFrameMode == ClearEveryFrame) {
            result = runFinalProgram(1.0f, 1.0f, false);
        } else {
            FloatImage lightImage = runFinalProgram(0.0f, 1.0f, false);
            FloatImage ptImage = runFinalProgram(1.0f, 0.0f, false);
            ptImage *= 1.0f / (float)mNumIterations;
            result = lightImage + ptImage;
        }
        const size_t i = mOutputFilename.find_last_of('.');
        const std::string ext = i == std::string::npos ? std::string() : mOutputFilename.substr(i);
        FloatImage::Save(FloatImage::FlipY(result), mOutputFilename.substr(0, i) + "_" + std::to_string(mNumIterations) + ext);
    }

    void destroy() {
        if (mHandle) { evplp_destroy(mHandle); mHandle = nullptr; }
    }

    evplp_handle handle() const { return mHandle; }
    int numIterations() const { return mNumIterations; }
    float elapsedMs() const { return mElapsedMs; }

    uint32_t mRngOffset = 0, mNumMaxBounce = 0;
    int mNumMaxIteration = 0, mNumSamplePerPixel = 0;
    float mTimelimitMs = 0;
    EFrame mFrameMode = Accumulate;
    bool mJitter = true, mUseStat = false, mDoWriteEveryFrame = false;
    RealTime::Sink* mSink = nullptr;
    void setSink(RealTime::Sink* sink) { mSink = sink; }
    std::string mOutputFilename, mStatFilename;

private:
    void check(int rc, const char* what) {
        if (rc != EVPLP_OK) throw std::runtime_error(std::string(what) + ": " + evplp_last_error());
    }
    int mDevice = 0, mRank = 0, mWorldSize = 1;
    bool mWriteOutputs = true;
    evplp_handle mHandle = nullptr;
    shared_ptr<RtScene> mScene;
    Vec2 mResolution, mInvResolution;
    Vec3 mCamF, mCamS, mCamU;
    float mTanHalfX = 0, mTanHalfY = 0;
    std::unique_ptr<IndependentSampler> mMainSampler;
    int mNumIterations = 0;
    StopWatch mMasterWatch;
    float mElapsedMs = 0;
};

}  // namespace evplp_host
