// rtcommon.h -- host scene model with the reference's names and semantics
// (reflectcuts/realtimetechniques/rtcommon.h): RtTexture, RtMaterial, RtMesh, RtAreaLight,
// RtStableCamera, RtScene::{addObject, addAreaLight, setCamera, totalArea,
// findBoundingSphereRadius}.  The OptiX / OpenGL upload members of the reference are
// replaced by RtScene::upload(evplp_handle), which feeds the C ABI.
//
// Asset loading: the reference imports OBJ/MTL through Assimp 3.3 and decodes textures with
// stb_image; here a small OBJ/MTL reader restates the parts of that import the path
// depends on (rtcommon.h:644-757): fan triangulation, one mesh per (object, material),
// identical (position, texcoord) vertices joined, material 0 = Assimp's "DefaultMaterial",
// Kd / Ks / Ns with the reference's "shininess / 4" correction of Assimp's 4x scaling
// (i.e. the exponent is the MTL Ns), map_Kd / map_Ks / map_Ns textures flipped vertically,
// value/255 with gamma 1, alpha 0.  Texture files: JPEG (jpegdecode.h, bit-exact with the
// reference's stb_image v2.16) or binary PPM/PGM (stb_image reads those too).
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/evplp.h"
#include "json.h"
#include "jpegdecode.h"
#include "pngdecode.h"

namespace evplp_host {

using std::make_shared;
using std::shared_ptr;

struct Vec2 { float x = 0, y = 0; };
struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : i == 1 ? y : z; }
};
struct Vec4 { float x = 0, y = 0, z = 0, w = 0; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return Vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }  // glm::cross
inline float length(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return a * inv; }  // glm::normalize = v * inversesqrt(dot)

namespace Math {
const float Pi = 3.14159265358979323846264338327950288f;      // math/math.h:14
const float InvPi = 0.318309886183790671537767526745028724f;  // math/math.h:15
}

inline Vec3 ToVec3(const Json& j) { return Vec3(j[0].as_float(), j[1].as_float(), j[2].as_float()); }
inline Vec4 ToVec4(const Json& j) { Vec4 v; v.x = j[0].as_float(); v.y = j[1].as_float(); v.z = j[2].as_float(); v.w = j[3].as_float(); return v; }

// Triangle::ComputeArea (shapes/trianglemesh.cpp:13-19)
inline float ComputeArea(Vec3 a, Vec3 b, Vec3 c) { return length(cross(b - a, c - a)) / 2.0f; }

struct RtTexture {
    int mWidth = 1, mHeight = 1;
    std::vector<float> mData;  // RGBA32F, row 0 = bottom (stbi flip-on-load, rtcommon.h:32)

    RtTexture() {}
    // RtTexture(r, g, b, gamma) (rtcommon.h:80-90)
    RtTexture(float r, float g, float b, float gamma) {
        mData = {std::pow(r, gamma), std::pow(g, gamma), std::pow(b, gamma), 0.f};
    }
    // RtTexture(filepath, gamma) (rtcommon.h:139-194): 3 channels forced, pow(byte/255, gamma), alpha 0
    RtTexture(const std::string& filepath, float gamma) {
        std::ifstream f(filepath, std::ios::binary);
        if (!f.is_open()) throw std::runtime_error("RtTexture: cannot open " + filepath);
        std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        std::vector<unsigned char> rgb;  // top-down, 3 channels (stbi_load(.., 3))
        if (jpeg::IsJpeg(bytes.data(), bytes.size())) {
            jpeg::Image img;
            try {
                jpeg::Decode(bytes.data(), bytes.size(), &img);
            } catch (const std::exception& e) {
                throw std::runtime_error("RtTexture: " + filepath + ": " + e.what());
            }
            mWidth = img.width; mHeight = img.height;
            rgb.swap(img.rgb);
        } else if (png::IsPng(bytes.data(), bytes.size())) {
            // stbi_load(.., 3): grey replicated, alpha dropped.  (The reference's own 4-channel branch then indexes this
            // 3-channel buffer with stride 4, rtcommon.h:178-189 -- a bug no bundled scene reaches; here every file is read
            // as the 3 channels stb returned.)
            png::Image img;
            try {
                img = png::Decode(bytes.data(), bytes.size(), 3);
            } catch (const std::exception& e) {
                throw std::runtime_error("RtTexture: " + filepath + ": " + e.what());
            }
            mWidth = img.width; mHeight = img.height;
            rgb.swap(img.pixels);
        } else {
            decodePnm(bytes, filepath, &rgb);
        }
        mData.resize((size_t)mWidth * mHeight * 4);
        for (int y = 0; y < mHeight; y++)
            for (int x = 0; x < mWidth; x++) {
                const unsigned char* src = &rgb[((size_t)(mHeight - 1 - y) * mWidth + x) * 3];  // flip vertically
                float* dst = &mData[((size_t)y * mWidth + x) * 4];
                for (int c = 0; c < 3; c++) dst[c] = std::pow((float)src[c] / 255.0f, gamma);
                dst[3] = 0.f;
            }
    }

private:
    // binary PPM / PGM (what the scene generator writes)
    void decodePnm(const std::vector<unsigned char>& bytes, const std::string& filepath, std::vector<unsigned char>* rgb) {
        size_t pos = 0;
        auto next_token = [&]() {
            for (;;) {
                while (pos < bytes.size() && isspace(bytes[pos])) pos++;
                if (pos < bytes.size() && bytes[pos] == '#') { while (pos < bytes.size() && bytes[pos] != '\n') pos++; continue; }
                break;
            }
            std::string tok;
            while (pos < bytes.size() && !isspace(bytes[pos])) tok.push_back((char)bytes[pos++]);
            if (tok.empty()) throw std::runtime_error("RtTexture: truncated header in " + filepath);
            return tok;
        };
        const std::string magic = bytes.size() >= 2 ? next_token() : std::string();
        if (magic != "P6" && magic != "P5")
            throw std::runtime_error("RtTexture: " + filepath + " is not a JPEG, PNG or binary PPM/PGM file");
        mWidth = atoi(next_token().c_str()); mHeight = atoi(next_token().c_str());
        const int maxv = atoi(next_token().c_str());
        pos++;  // the single whitespace after maxval
        if (maxv != 255 || mWidth <= 0 || mHeight <= 0) throw std::runtime_error("RtTexture: unsupported PPM in " + filepath);
        const int ch = magic == "P6" ? 3 : 1;
        const size_t n = (size_t)mWidth * mHeight;
        if (bytes.size() - pos < n * ch) throw std::runtime_error("RtTexture: truncated data in " + filepath);
        rgb->resize(n * 3);
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) (*rgb)[i * 3 + c] = bytes[pos + i * ch + (ch == 3 ? c : 0)];
    }
};

struct RtMaterial {
    shared_ptr<RtTexture> mLambertReflectance, mPhongReflectance, mPhongExponent;
    Vec4 mLightIntensity;
};

struct RtMesh {
    int32_t mNumVertices = 0, mNumTriangles = 0, mMatIndex = 0;
    std::vector<float> mVertices, mTexCoords;
    std::vector<int32_t> mTriIndices;

    Vec3 vertex(int32_t i) const { return Vec3(mVertices[3 * i], mVertices[3 * i + 1], mVertices[3 * i + 2]); }
    // rtcommon.h:439-456
    float recomputeArea() const {
        float sumArea = 0.f;
        for (int32_t i = 0; i < mNumTriangles; i++)
            sumArea += ComputeArea(vertex(mTriIndices[3 * i]), vertex(mTriIndices[3 * i + 1]), vertex(mTriIndices[3 * i + 2]));
        return sumArea;
    }
};

struct RtAreaLight {
    shared_ptr<RtMesh> mMesh;
    Vec4 mLightIntensity, mPrecomputedLightIntensity;
};

struct RtCameraBase {
    virtual ~RtCameraBase() {}
    virtual Vec3 getOrigin() const = 0;
};

// rtcommon.h:546-598.  computeVpMatrix() = perspectiveRH(fovy, aspect, 0.1, 100) * lookAtRH(origin, lookAt, up)
// (glm gtc/matrix_transform.inl:258-276, 521-546); the kernels take the same camera as the
// lookAt basis plus the half-angle tangents, which is what that matrix encodes.
struct RtStableCamera : RtCameraBase {
    RtStableCamera(const Json& json, float aspectRatio) {
        float fovy = 0;
        if (json.contains("fovy")) {
            fovy = json["fovy"].as_float() * 0.01745329251994329576923690768489f;  // glm::radians
        } else if (json.contains("fovx")) {
            float fovxDegree = json["fovx"].as_float();
            fovy = 2.0f * std::atan2(std::tan(fovxDegree * 0.01745329251994329576923690768489f * 0.5f), aspectRatio);
        } else {
            throw std::runtime_error("camera: forgot fov");
        }
        mOrigin = ToVec3(json.at("origin"));
        mLookAt = ToVec3(json.at("direction"));  // a look-at POINT (rtcommon.h:567,588)
        mUp = ToVec3(json.at("up"));
        mFovy = fovy;
        mAspectRatio = aspectRatio;
    }
    Vec3 getOrigin() const override { return mOrigin; }
    void basis(Vec3* f, Vec3* s, Vec3* u, float* tanHalfX, float* tanHalfY) const {
        *f = normalize(mLookAt - mOrigin);
        *s = normalize(cross(*f, mUp));
        *u = cross(*s, *f);
        *tanHalfY = std::tan(mFovy / 2.0f);
        *tanHalfX = mAspectRatio * *tanHalfY;
    }
    Vec3 mOrigin, mLookAt, mUp;
    float mFovy = 0, mAspectRatio = 1;
};

struct RtScene {
    std::vector<shared_ptr<RtMesh>> mMeshes;
    std::vector<shared_ptr<RtMaterial>> mMaterials;
    shared_ptr<RtAreaLight> mArealight;
    shared_ptr<RtStableCamera> mCamera;
    bool isAlreadyHaveLightSource = false;

    static std::string dir_of(const std::string& filepath) {
        size_t p = filepath.find_last_of("/\\");
        return p == std::string::npos ? std::string() : filepath.substr(0, p + 1);
    }

    struct MtlEntry {
        std::string name;
        float kd[3] = {0.6f, 0.6f, 0.6f}, ks[3] = {0.f, 0.f, 0.f}, ns = 0.f;
        std::string mapKd, mapKs, mapNs;
    };

    static void load_mtl(const std::string& path, std::vector<MtlEntry>& out) {
        std::ifstream f(path);
        if (!f.is_open()) { std::cerr << "warning: cannot open material library " << path << "\n"; return; }
        std::string line;
        while (std::getline(f, line)) {
            std::istringstream ss(line);
            std::string key;
            if (!(ss >> key) || key[0] == '#') continue;
            if (key == "newmtl") { MtlEntry e; ss >> e.name; out.push_back(e); continue; }
            if (out.empty()) continue;
            MtlEntry& e = out.back();
            if (key == "Kd") ss >> e.kd[0] >> e.kd[1] >> e.kd[2];
            else if (key == "Ks") ss >> e.ks[0] >> e.ks[1] >> e.ks[2];
            else if (key == "Ns") ss >> e.ns;
            else if (key == "map_Kd") e.mapKd = parse_map(ss);
            else if (key == "map_Ks") e.mapKs = parse_map(ss);
            else if (key == "map_Ns") e.mapNs = parse_map(ss);
        }
    }

    // "map_Kd [-option args ...] file name": Assimp's ObjFileMtlImporter::getTexture skips the texture options and takes the
    // rest of the line as the file name (it may contain spaces); files written on Windows use backslashes.
    static std::string parse_map(std::istringstream& ss) {
        std::vector<std::string> tok;
        std::string t;
        while (ss >> t) tok.push_back(t);
        auto numeric = [](const std::string& v) { char* e; strtod(v.c_str(), &e); return e != v.c_str() && *e == 0; };
        size_t i = 0;
        while (i < tok.size() && tok[i].size() > 1 && tok[i][0] == '-' && !numeric(tok[i])) {
            const std::string& o = tok[i++];
            if (o == "-o" || o == "-s" || o == "-t") { int n = 0; while (n < 3 && i < tok.size() && numeric(tok[i])) { i++; n++; } }
            else if (o == "-mm") i += 2;
            else i += 1;  // -blendu -blendv -cc -clamp -bm -boost -texres -imfchan -type: one argument
        }
        std::string name;
        for (; i < tok.size(); i++) name += (name.empty() ? "" : " ") + tok[i];
        for (char& c : name) if (c == '\\') c = '/';
        return name;
    }

    static shared_ptr<RtTexture> load_texture(const std::string& filedir, const std::string& map, const float* rgb, bool shininess) {
        static std::map<std::string, shared_ptr<RtTexture>> gTexturesMap;  // rtcommon.h:33
        if (!map.empty()) {
            const std::string p = filedir + map;
            auto it = gTexturesMap.find(p);
            if (it != gTexturesMap.end()) return it->second;
            auto t = make_shared<RtTexture>(p, 1.0f);
            gTexturesMap[p] = t;
            return t;
        }
        // Assimp reports 4 x Ns as AI_MATKEY_SHININESS and the reference divides by 4 (rtcommon.h:53-64): net = Ns
        if (shininess) return make_shared<RtTexture>(rgb[0], rgb[0], rgb[0], 1.0f);
        return make_shared<RtTexture>(rgb[0], rgb[1], rgb[2], 1.0f);
    }

    // rtcommon.h:644-757
    void addObject(const std::string& filepath, const Vec4& lightIntensity = Vec4(), bool overrideMaterial = false,
                   shared_ptr<RtMaterial> defaultMat = shared_ptr<RtMaterial>()) {
        std::ifstream f(filepath);
        if (!f.is_open()) throw std::runtime_error("Impossible to load the scene: " + filepath);
        const std::string filedir = dir_of(filepath);
        std::vector<float> pos, uv;
        std::vector<MtlEntry> mtl;
        struct Builder {
            shared_ptr<RtMesh> mesh;
            std::map<std::pair<int, int>, int32_t> remap;
        };
        std::vector<Builder> builders;
        std::map<std::pair<std::string, int>, size_t> builderOf;  // (object name, material) -> builder
        std::string objectName;
        int curMat = 0;  // 0 = DefaultMaterial
        std::string line;
        while (std::getline(f, line)) {
            const size_t lead = line.find_first_not_of(" \t");
            if (lead == std::string::npos) continue;
            if (lead) line.erase(0, lead);
            if (line.size() < 2) continue;
            if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
                float x, y, z;
                if (sscanf(line.c_str() + 2, "%f %f %f", &x, &y, &z) == 3) { pos.push_back(x); pos.push_back(y); pos.push_back(z); }
            } else if (line[0] == 'v' && line[1] == 't') {
                float u = 0, v = 0;
                sscanf(line.c_str() + 3, "%f %f", &u, &v);
                uv.push_back(u); uv.push_back(v);
            } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
                auto key = std::make_pair(objectName, curMat);
                auto it = builderOf.find(key);
                if (it == builderOf.end()) {
                    Builder b;
                    b.mesh = make_shared<RtMesh>();
                    b.mesh->mMatIndex = curMat;
                    builders.push_back(b);
                    it = builderOf.insert({key, builders.size() - 1}).first;
                }
                Builder& b = builders[it->second];
                int32_t face[64];
                int nf = 0;
                const char* p = line.c_str() + 2;
                while (*p && nf < 64) {
                    while (*p == ' ' || *p == '\t') p++;
                    if (!*p || *p == '\r' || *p == '\n') break;
                    char* end;
                    long vi = strtol(p, &end, 10), ti = 0;
                    bool hasT = false;
                    if (end == p) break;
                    p = end;
                    if (*p == '/') {
                        p++;
                        if (*p != '/') { ti = strtol(p, &end, 10); hasT = end != p; p = end; }
                        if (*p == '/') { p++; strtol(p, &end, 10); p = end; }
                    }
                    const int nv = (int)(pos.size() / 3), nt = (int)(uv.size() / 2);
                    int v0 = vi > 0 ? (int)vi - 1 : nv + (int)vi;
                    int t0 = hasT ? (ti > 0 ? (int)ti - 1 : nt + (int)ti) : -1;
                    if (v0 < 0 || v0 >= nv) throw std::runtime_error("obj: vertex index out of range in " + filepath);
                    auto rk = std::make_pair(v0, t0);
                    auto ri = b.remap.find(rk);
                    int32_t idx;
                    if (ri == b.remap.end()) {
                        idx = b.mesh->mNumVertices++;
                        b.remap[rk] = idx;
                        b.mesh->mVertices.push_back(pos[3 * v0]); b.mesh->mVertices.push_back(pos[3 * v0 + 1]); b.mesh->mVertices.push_back(pos[3 * v0 + 2]);
                        if (t0 >= 0 && t0 < nt) { b.mesh->mTexCoords.push_back(uv[2 * t0]); b.mesh->mTexCoords.push_back(uv[2 * t0 + 1]); }
                        else { b.mesh->mTexCoords.push_back(0.f); b.mesh->mTexCoords.push_back(0.f); }  // rtcommon.h:700-704
                    } else {
                        idx = ri->second;
                    }
                    face[nf++] = idx;
                }
                for (int k = 1; k + 1 < nf; k++) {  // aiProcess_Triangulate: fan
                    b.mesh->mTriIndices.push_back(face[0]); b.mesh->mTriIndices.push_back(face[k]); b.mesh->mTriIndices.push_back(face[k + 1]);
                    b.mesh->mNumTriangles++;
                }
            } else if (line.compare(0, 6, "usemtl") == 0) {
                std::istringstream ss(line.substr(6));
                std::string name;
                ss >> name;
                curMat = 0;
                for (size_t m = 0; m < mtl.size(); m++) if (mtl[m].name == name) curMat = (int)m + 1;
            } else if (line.compare(0, 6, "mtllib") == 0) {
                // the rest of the line is the file name (it may contain spaces)
                std::string name = line.substr(6);
                const size_t a = name.find_first_not_of(" \t"), b = name.find_last_not_of(" \t\r\n");
                name = a == std::string::npos ? std::string() : name.substr(a, b - a + 1);
                for (char& ch : name) if (ch == '\\') ch = '/';
                load_mtl(filedir + name, mtl);
            } else if ((line[0] == 'o' || line[0] == 'g') && line[1] == ' ') {
                std::istringstream ss(line.substr(2));
                ss >> objectName;
            }
        }
        const size_t matOffset = mMaterials.size();
        for (Builder& b : builders) {
            if (b.mesh->mNumTriangles == 0) continue;
            b.mesh->mMatIndex = overrideMaterial ? (int32_t)matOffset : (int32_t)(matOffset + b.mesh->mMatIndex);
            mMeshes.push_back(b.mesh);
        }
        if (overrideMaterial) {
            mMaterials.push_back(defaultMat);
        } else {
            // material 0: Assimp's DefaultMaterial (diffuse 0.6, no specular)
            MtlEntry def;
            std::vector<MtlEntry> all;
            all.push_back(def);
            all.insert(all.end(), mtl.begin(), mtl.end());
            for (const MtlEntry& e : all) {
                auto mat = make_shared<RtMaterial>();
                mat->mLambertReflectance = load_texture(filedir, e.mapKd, e.kd, false);
                mat->mPhongReflectance = load_texture(filedir, e.mapKs, e.ks, false);
                const float ns[3] = {e.ns, e.ns, e.ns};
                mat->mPhongExponent = load_texture(filedir, e.mapNs, ns, true);
                mat->mLightIntensity = lightIntensity;
                mMaterials.push_back(mat);
            }
        }
    }

    // rtcommon.h:759-768
    float totalArea() const {
        float sumArea = 0.f;
        for (auto& m : mMeshes) sumArea += m->recomputeArea();
        return sumArea;
    }

    // rtcommon.h:772-798
    void addAreaLight(const std::string& filepath, const Vec4& lightIntensity) {
        if (isAlreadyHaveLightSource) throw std::runtime_error("only one area light is supported");
        const size_t beforeSize = mMeshes.size();
        isAlreadyHaveLightSource = true;
        Vec4 pre = lightIntensity;
        pre.x = lightIntensity.x * Math::Pi; pre.y = lightIntensity.y * Math::Pi; pre.z = lightIntensity.z * Math::Pi;
        auto mat = make_shared<RtMaterial>();
        mat->mLambertReflectance = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
        mat->mPhongExponent = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
        mat->mPhongReflectance = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
        mat->mLightIntensity = pre;
        addObject(filepath, pre, true, mat);
        if (mMeshes.size() - beforeSize != 1) throw std::runtime_error("the area light OBJ must contain exactly one mesh (rtcommon.h:795)");
        mArealight = make_shared<RtAreaLight>();
        mArealight->mMesh = mMeshes.back();
        mArealight->mLightIntensity = lightIntensity;
        mArealight->mPrecomputedLightIntensity = pre;
    }

    void setCamera(shared_ptr<RtStableCamera> camera) { mCamera = camera; }

    // rtcommon.h:805-814 with Aabb::Union / DiagonalLength2 (math/aabb.h:27-41)
    float findBoundingSphereRadius() const {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (auto& m : mMeshes)
            for (int32_t i = 0; i < m->mNumVertices; i++)
                for (int k = 0; k < 3; k++) {
                    lo[k] = std::min(lo[k], m->mVertices[3 * i + k]);
                    hi[k] = std::max(hi[k], m->mVertices[3 * i + k]);
                }
        Vec3 diag(std::max(hi[0] - lo[0], 0.0f), std::max(hi[1] - lo[1], 0.0f), std::max(hi[2] - lo[2], 0.0f));
        float diameter = std::sqrt(dot(diag, diag));
        return diameter / 2.0f;
    }

    size_t numTriangles() const {
        size_t n = 0;
        for (auto& m : mMeshes) n += (size_t)m->mNumTriangles;
        return n;
    }

    int lightMeshIndex() const {
        for (size_t i = 0; i < mMeshes.size(); i++) if (mMeshes[i] == mArealight->mMesh) return (int)i;
        return -1;
    }

    // descriptors for evplp_upload_scene (host arrays stay owned by the scene, like rtcommon.h:464-467)
    void descriptors(std::vector<EvplpMeshDesc>& md, std::vector<EvplpMaterialDesc>& mt) const {
        md.resize(mMeshes.size());
        for (size_t i = 0; i < mMeshes.size(); i++) {
            const RtMesh& m = *mMeshes[i];
            md[i].vertices = m.mVertices.data(); md[i].texcoords = m.mTexCoords.data(); md[i].indices = m.mTriIndices.data();
            md[i].numVertices = m.mNumVertices; md[i].numTriangles = m.mNumTriangles; md[i].matIndex = m.mMatIndex;
        }
        mt.resize(mMaterials.size());
        for (size_t i = 0; i < mMaterials.size(); i++) {
            const RtMaterial& m = *mMaterials[i];
            mt[i].lambertReflectance = m.mLambertReflectance->mData.data(); mt[i].lambertW = m.mLambertReflectance->mWidth; mt[i].lambertH = m.mLambertReflectance->mHeight;
            mt[i].phongReflectance = m.mPhongReflectance->mData.data(); mt[i].phongW = m.mPhongReflectance->mWidth; mt[i].phongH = m.mPhongReflectance->mHeight;
            mt[i].phongExponent = m.mPhongExponent->mData.data(); mt[i].exponentW = m.mPhongExponent->mWidth; mt[i].exponentH = m.mPhongExponent->mHeight;
            mt[i].lightIntensity[0] = m.mLightIntensity.x; mt[i].lightIntensity[1] = m.mLightIntensity.y;
            mt[i].lightIntensity[2] = m.mLightIntensity.z; mt[i].lightIntensity[3] = m.mLightIntensity.w;
        }
    }

    int upload(evplp_handle h) const {
        std::vector<EvplpMeshDesc> md;
        std::vector<EvplpMaterialDesc> mt;
        descriptors(md, mt);
        const float pre[4] = {mArealight->mPrecomputedLightIntensity.x, mArealight->mPrecomputedLightIntensity.y,
                              mArealight->mPrecomputedLightIntensity.z, mArealight->mPrecomputedLightIntensity.w};
        const float disp[4] = {mArealight->mLightIntensity.x, mArealight->mLightIntensity.y, mArealight->mLightIntensity.z,
                               mArealight->mLightIntensity.w};
        return evplp_upload_scene(h, md.data(), (int32_t)md.size(), mt.data(), (int32_t)mt.size(), lightMeshIndex(), pre, disp);
    }
};

// LoadScene (main.cpp:42-85)
inline shared_ptr<RtScene> LoadScene(const Json& json, const std::string& jsonFilename) {
    auto rtScene = make_shared<RtScene>();
    if (json["scene"].is_null()) return nullptr;
    const std::string dir = RtScene::dir_of(jsonFilename);
    auto resolve = [&](const std::string& p) { return (!p.empty() && p[0] == '/') ? p : dir + p; };
    for (size_t i = 0; i < json["scene"].size(); i++) rtScene->addObject(resolve(json["scene"][i].as_string()));
    rtScene->addAreaLight(resolve(json["arealight"]["obj"].as_string()), ToVec4(json["arealight"]["intensity"]));
    const float aspect = json["resX"].as_float() / json["resY"].as_float();
    if (json.contains("camera")) rtScene->setCamera(make_shared<RtStableCamera>(json["camera"], aspect));
    else if (json.contains("stablecamera")) rtScene->setCamera(make_shared<RtStableCamera>(json["stablecamera"], aspect));
    return rtScene;
}

}  // namespace evplp_host
