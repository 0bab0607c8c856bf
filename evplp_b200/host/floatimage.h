// floatimage.h -- RGB float image, the headless sink of the technique.
// Replaces the parts of reflectcuts/common/floatimage/floatimage.{h,cpp} the path uses:
// SavePFM (:178-199, byte-compatible: "PF\n<w> <h>\n-1\n", rows written bottom-up),
// FlipY (:114-128), ComputeMse / ComputeRelMse (:64-112), operator+ / *=.
#pragma once
#include <cassert>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace evplp_host {

class FloatImage {
public:
    FloatImage() {}
    FloatImage(size_t width, size_t height) : mWidth(width), mHeight(height), mData(width * height * 3, 0.f) {}
    size_t width() const { return mWidth; }
    size_t height() const { return mHeight; }
    float* data() { return mData.data(); }
    const float* data() const { return mData.data(); }

    FloatImage& operator*=(float s) { for (float& v : mData) v *= s; return *this; }
    friend FloatImage operator+(const FloatImage& a, const FloatImage& b) {
        assert(a.mWidth == b.mWidth && a.mHeight == b.mHeight);
        FloatImage r(a.mWidth, a.mHeight);
        for (size_t i = 0; i < r.mData.size(); i++) r.mData[i] = a.mData[i] + b.mData[i];
        return r;
    }

    static FloatImage FlipY(const FloatImage& f) {
        FloatImage r(f.mWidth, f.mHeight);
        for (size_t row = 0; row < f.mHeight; row++)
            for (size_t k = 0; k < f.mWidth * 3; k++) r.mData[row * f.mWidth * 3 + k] = f.mData[(f.mHeight - 1 - row) * f.mWidth * 3 + k];
        return r;
    }

    static void SavePFM(const FloatImage& f, const std::string& filepath) {
        std::ofstream os(filepath, std::ios::binary);
        if (!os.is_open()) throw std::runtime_error("FloatImage::SavePFM: cannot write " + filepath);
        os << "PF" << std::endl;
        os << f.mWidth << " " << f.mHeight << std::endl;
        os << "-1" << std::endl;
        for (size_t i = 0; i < f.mHeight; i++)
            os.write((const char*)(&f.mData[f.mWidth * (f.mHeight - i - 1) * 3]), (std::streamsize)(sizeof(float) * f.mWidth * 3));
    }
    // FloatImage::Save dispatches on the extension (floatimage.cpp:260-273); only .pfm is built here.
    static void Save(const FloatImage& f, const std::string& filepath) {
        size_t i = filepath.find_last_of('.');
        if (i == std::string::npos || filepath.substr(i) != ".pfm") throw std::runtime_error("FloatImage::Save: only .pfm is supported: " + filepath);
        SavePFM(f, filepath);
    }
    static FloatImage LoadPFM(const std::string& filepath) {
        std::ifstream is(filepath, std::ios::binary);
        if (!is.is_open()) throw std::runtime_error("FloatImage::LoadPFM: cannot open " + filepath);
        std::string magic; size_t w, h; float scale;
        is >> magic >> w >> h >> scale;
        is.get();
        if (magic != "PF" || scale >= 0) throw std::runtime_error("FloatImage::LoadPFM: unsupported file " + filepath);
        FloatImage f(w, h);
        for (size_t i = 0; i < h; i++) is.read((char*)&f.mData[w * (h - i - 1) * 3], (std::streamsize)(sizeof(float) * w * 3));
        return f;
    }

    static float ComputeMse(const FloatImage& a, const FloatImage& ref) {
        assert(a.mWidth == ref.mWidth && a.mHeight == ref.mHeight);
        float result = 0;
        const float numPixels = (float)(a.mWidth * a.mHeight);
        for (size_t p = 0; p < a.mWidth * a.mHeight; p++) {
            float d0 = a.mData[3 * p] - ref.mData[3 * p], d1 = a.mData[3 * p + 1] - ref.mData[3 * p + 1], d2 = a.mData[3 * p + 2] - ref.mData[3 * p + 2];
            result += d0 * d0 + d1 * d1 + d2 * d2;
        }
        return result / numPixels;
    }
    static float ComputeRelMse(const FloatImage& a, const FloatImage& ref) {
        assert(a.mWidth == ref.mWidth && a.mHeight == ref.mHeight);
        float result = 0;
        const float numPixels = (float)(a.mWidth * a.mHeight);
        for (size_t p = 0; p < a.mWidth * a.mHeight; p++) {
            float r0 = ref.mData[3 * p], r1 = ref.mData[3 * p + 1], r2 = ref.mData[3 * p + 2];
            float d0 = a.mData[3 * p] - r0, d1 = a.mData[3 * p + 1] - r1, d2 = a.mData[3 * p + 2] - r2;
            float numerator = d0 * d0 + d1 * d1 + d2 * d2;
            float denominator = r0 * r0 + r1 * r1 + r2 * r2 + 0.001f;
            result += numerator / denominator;
        }
        return result / numPixels;
    }

    // Masked error metrics.  The reference ships scene/conference/conference_mask.png with the note that the error metric of
    // the conference scene (the only one with a visible light source, which is not anti-aliased) must be computed under it
    // (scene/conference/README.md:1-2); the metric code itself (floatimage.cpp:64-112) takes no mask, so the weighting is
    // ours: weight = mask value / 255 (white = counted, black = the light), result = sum(w * err) / sum(w).  `mask` holds one
    // weight per pixel in the images' row order.
    static float ComputeMse(const FloatImage& a, const FloatImage& ref, const std::vector<float>& mask) { return masked(a, ref, mask, false); }
    static float ComputeRelMse(const FloatImage& a, const FloatImage& ref, const std::vector<float>& mask) { return masked(a, ref, mask, true); }

private:
    static float masked(const FloatImage& a, const FloatImage& ref, const std::vector<float>& mask, bool relative) {
        if (a.mWidth != ref.mWidth || a.mHeight != ref.mHeight || mask.size() != a.mWidth * a.mHeight)
            throw std::runtime_error("FloatImage: image and mask sizes differ");
        float result = 0, weight = 0;
        for (size_t p = 0; p < a.mWidth * a.mHeight; p++) {
            float r0 = ref.mData[3 * p], r1 = ref.mData[3 * p + 1], r2 = ref.mData[3 * p + 2];
            float d0 = a.mData[3 * p] - r0, d1 = a.mData[3 * p + 1] - r1, d2 = a.mData[3 * p + 2] - r2;
            float e = d0 * d0 + d1 * d1 + d2 * d2;
            if (relative) e = e / (r0 * r0 + r1 * r1 + r2 * r2 + 0.001f);
            result += mask[p] * e; weight += mask[p];
        }
        return weight > 0 ? result / weight : 0.f;
    }
    size_t mWidth = 0, mHeight = 0;
    std::vector<float> mData;
};

}  // namespace evplp_host
