// realtime.h -- the headless sink: RealTime::loop's contract (reflectcuts/common/realtime.h:100-141) without GLFW / OpenGL.
//   do { title = ""; result = beforeSwap(&title); if (result) present();          // glfwSwapBuffers + glfwPollEvents
//        if (afterSwap) result = result && afterSwap(&title);                     // not called once beforeSwap said stop
//        frameCount++; once >= 1000 ms have passed: window title = "<fps>fps, <ms>ms, <title>", counters reset
//   } while (result && !escape && !windowShouldClose);
// "present" hands the frame to an optional Sink (a viewer, a test probe); escape / close are the Sink's shouldClose().
#pragma once
#include <chrono>
#include <functional>
#include <string>

namespace evplp_host {

class RealTime {
public:
    struct Sink {
        virtual ~Sink() {}
        virtual void present(uint64_t frameIndex) {}     // after a frame whose beforeSwap returned true
        virtual bool shouldClose() { return false; }     // ESC / window closed
        virtual void title(const std::string& t) {}      // setWindowTitle
    };

    explicit RealTime(Sink* sink = nullptr) : mSink(sink) {}

    void loop(const std::function<bool(std::string* titleExtend)>& beforeSwap,
              const std::function<bool(std::string* titleExtend)>& afterSwap = nullptr) {
        using clock = std::chrono::steady_clock;
        auto t0 = clock::now();
        bool result = false;
        double frameCount = 0;
        std::string titleExtend;
        do {
            titleExtend = "";
            result = beforeSwap(&titleExtend);
            if (result) {
                if (mSink) mSink->present(mFramesPresented);
                mFramesPresented++;
            }
            if (afterSwap != nullptr) result = result && afterSwap(&titleExtend);
            frameCount++;
            mLoopPasses++;
            const long long timeSpent = std::chrono::duration_cast<std::chrono::milliseconds>(clock::now() - t0).count();
            if (timeSpent >= 1000) {
                const double dTimeSpent = (double)timeSpent;
                const double fps = frameCount / dTimeSpent * 1000, spf = dTimeSpent / frameCount;
                setWindowTitle(std::to_string(fps) + "fps, " + std::to_string(spf) + "ms, " + titleExtend);
                frameCount = 0;
                t0 = clock::now();
            }
        } while (result && !(mSink && mSink->shouldClose()));
    }

    void setWindowTitle(const std::string& title) {
        mTitle = title;
        if (mSink) mSink->title(title);
    }
    const std::string& windowTitle() const { return mTitle; }
    uint64_t framesPresented() const { return mFramesPresented; }
    uint64_t loopPasses() const { return mLoopPasses; }

private:
    Sink* mSink;
    std::string mTitle;
    uint64_t mFramesPresented = 0, mLoopPasses = 0;
};

}  // namespace evplp_host
