// scenegen.h -- procedural stand-ins for the reference's bundled scenes.
//
// Every geometry asset of the reference (scene/*/*.obj, *.mtl) is a git-LFS pointer, not data
// (SURVEY.md F1), so the named scenes are re-created procedurally: deterministic from an
// integer seed, same file names and formats (OBJ + MTL + JSON with the reference's keys,
// textures as binary PPM which stb_image also reads), same cameras and light intensities as
// scene/{conference,livingroom,buddha}/*.json, Z up.  GenerateScene() builds the RtScene in
// memory; ExportScene() writes the OBJ/MTL/PPM/JSON files that LoadScene() (and the
// reference's own loader) read back.
#pragma once
#include <cstdio>
#include <functional>
#include <sys/stat.h>
#include "rtcommon.h"

namespace evplp_host {

struct GenMaterial {
    std::string name;
    float kd[3], ks[3], ns;
    std::string mapKd;  // file name of a generated PPM ("" = constant)
};

struct GenScene {
    std::string name;
    std::vector<GenMaterial> materials;                 // MTL entries (scene material index = 1 + k; 0 = DefaultMaterial)
    std::vector<std::pair<std::string, int>> meshMat;   // per mesh: object name, MTL entry
    std::vector<shared_ptr<RtMesh>> meshes;
    shared_ptr<RtMesh> light;
    float lightIntensity[4];
    float camOrigin[3], camLookAt[3], camUp[3], fovx;
    std::map<std::string, shared_ptr<RtTexture>> textures;  // generated PPM textures by file name
};

struct Lcg {  // tiny deterministic generator for layout jitter
    uint32_t s;
    explicit Lcg(uint32_t seed) : s(seed * 747796405u + 2891336453u) {}
    float next() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); }
    float range(float a, float b) { return a + (b - a) * next(); }
};

class MeshBuilder {
public:
    shared_ptr<RtMesh> mesh = make_shared<RtMesh>();
    int32_t addVertex(Vec3 p, float u, float v) {
        mesh->mVertices.push_back(p.x); mesh->mVertices.push_back(p.y); mesh->mVertices.push_back(p.z);
        mesh->mTexCoords.push_back(u); mesh->mTexCoords.push_back(v);
        return mesh->mNumVertices++;
    }
    void addTri(int32_t a, int32_t b, int32_t c) {
        mesh->mTriIndices.push_back(a); mesh->mTriIndices.push_back(b); mesh->mTriIndices.push_back(c);
        mesh->mNumTriangles++;
    }
    // (nu x nv) quads over p(s,t), s,t in [0,1]; front face = cross(dp/ds, dp/dt)
    void addSurface(int nu, int nv, const std::function<Vec3(float, float)>& p, float uvScaleU = 1.f, float uvScaleV = 1.f) {
        const int32_t base = mesh->mNumVertices;
        for (int j = 0; j <= nv; j++)
            for (int i = 0; i <= nu; i++) {
                float s = (float)i / (float)nu, t = (float)j / (float)nv;
                addVertex(p(s, t), s * uvScaleU, t * uvScaleV);
            }
        for (int j = 0; j < nv; j++)
            for (int i = 0; i < nu; i++) {
                int32_t a = base + j * (nu + 1) + i, b = a + 1, c = a + nu + 2, d = a + nu + 1;
                addTri(a, b, c);
                addTri(a, c, d);
            }
    }
    void addGrid(Vec3 o, Vec3 du, Vec3 dv, int nu, int nv, float su = 1.f, float sv = 1.f) {
        addSurface(nu, nv, [=](float s, float t) { return o + du * s + dv * t; }, su, sv);
    }
    // axis-aligned box, outward faces, each face tessellated n x n
    void addBox(Vec3 lo, Vec3 hi, int n = 1) {
        Vec3 d = hi - lo;
        addGrid(Vec3(lo.x, lo.y, hi.z), Vec3(d.x, 0, 0), Vec3(0, d.y, 0), n, n);   // +z
        addGrid(Vec3(lo.x, hi.y, lo.z), Vec3(d.x, 0, 0), Vec3(0, -d.y, 0), n, n);  // -z
        addGrid(Vec3(lo.x, lo.y, lo.z), Vec3(d.x, 0, 0), Vec3(0, 0, d.z), n, n);   // -y
        addGrid(Vec3(hi.x, hi.y, lo.z), Vec3(-d.x, 0, 0), Vec3(0, 0, d.z), n, n);  // +y
        addGrid(Vec3(hi.x, lo.y, lo.z), Vec3(0, d.y, 0), Vec3(0, 0, d.z), n, n);   // +x
        addGrid(Vec3(lo.x, hi.y, lo.z), Vec3(0, -d.y, 0), Vec3(0, 0, d.z), n, n);  // -x
    }
    // closed room: inward faces
    void addRoom(Vec3 lo, Vec3 hi, int n, int which /* bit mask: 1 floor 2 ceiling 4 walls */) {
        Vec3 d = hi - lo;
        if (which & 1) addGrid(Vec3(lo.x, lo.y, lo.z), Vec3(d.x, 0, 0), Vec3(0, d.y, 0), n, n, d.x * 0.25f, d.y * 0.25f);
        if (which & 2) addGrid(Vec3(lo.x, hi.y, hi.z), Vec3(d.x, 0, 0), Vec3(0, -d.y, 0), n, n);
        if (which & 4) {
            addGrid(Vec3(lo.x, hi.y, lo.z), Vec3(d.x, 0, 0), Vec3(0, 0, d.z), n, n);    // back  (normal -y)
            addGrid(Vec3(hi.x, lo.y, lo.z), Vec3(-d.x, 0, 0), Vec3(0, 0, d.z), n, n);   // front (normal +y)
            addGrid(Vec3(lo.x, lo.y, lo.z), Vec3(0, d.y, 0), Vec3(0, 0, d.z), n, n);    // left  (normal +x)
            addGrid(Vec3(hi.x, hi.y, lo.z), Vec3(0, -d.y, 0), Vec3(0, 0, d.z), n, n);   // right (normal -x)
        }
    }
};

inline shared_ptr<RtTexture> MakeProceduralTexture(int w, int h, int kind, uint32_t seed) {
    auto t = make_shared<RtTexture>();
    t->mWidth = w; t->mHeight = h;
    t->mData.resize((size_t)w * h * 4);
    Lcg rng(seed);
    std::vector<float> plank(16);
    for (float& p : plank) p = rng.range(0.75f, 1.0f);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float r, g, b;
            if (kind == 0) {  // wood planks
                float grain = 0.5f + 0.5f * std::sin((float)y * 0.9f + 3.0f * std::sin((float)x * 0.05f));
                float p = plank[(x * 16 / w) % 16];
                r = (0.45f + 0.15f * grain) * p; g = (0.28f + 0.10f * grain) * p; b = (0.14f + 0.05f * grain) * p;
            } else if (kind == 1) {  // checker carpet
                bool c = ((x * 8 / w) + (y * 8 / h)) % 2 == 0;
                r = c ? 0.55f : 0.25f; g = c ? 0.12f : 0.2f; b = c ? 0.1f : 0.45f;
            } else {  // speckled plaster
                float n = rng.range(0.85f, 1.0f);
                r = 0.72f * n; g = 0.7f * n; b = 0.65f * n;
            }
            // quantise to 8 bits so that the in-memory texture equals the PPM read back
            float* d = &t->mData[((size_t)y * w + x) * 4];
            d[0] = (float)(int)(r * 255.0f + 0.5f) / 255.0f;
            d[1] = (float)(int)(g * 255.0f + 0.5f) / 255.0f;
            d[2] = (float)(int)(b * 255.0f + 0.5f) / 255.0f;
            d[3] = 0.f;
        }
    return t;
}

inline int add_material(GenScene& g, const std::string& name, float kr, float kg, float kb, float sr, float sg, float sb, float ns,
                        const std::string& mapKd = "") {
    GenMaterial m;
    m.name = name; m.kd[0] = kr; m.kd[1] = kg; m.kd[2] = kb; m.ks[0] = sr; m.ks[1] = sg; m.ks[2] = sb; m.ns = ns; m.mapKd = mapKd;
    g.materials.push_back(m);
    return (int)g.materials.size() - 1;
}
inline void add_mesh(GenScene& g, const std::string& name, int mat, MeshBuilder& b) {
    g.meshMat.push_back({name, mat});
    g.meshes.push_back(b.mesh);
}

// A chair: seat, curved back, four legs.
inline void add_chair(MeshBuilder& b, Vec3 c, float yaw, int tess) {
    const float cs = std::cos(yaw), sn = std::sin(yaw);
    auto X = [=](Vec3 p) { return Vec3(c.x + p.x * cs - p.y * sn, c.y + p.x * sn + p.y * cs, c.z + p.z); };
    auto box = [&](Vec3 lo, Vec3 hi, int n) {
        Vec3 d = hi - lo;
        auto face = [&](Vec3 o, Vec3 du, Vec3 dv) { b.addSurface(n, n, [=](float s, float t) { return X(o + du * s + dv * t); }); };
        face(Vec3(lo.x, lo.y, hi.z), Vec3(d.x, 0, 0), Vec3(0, d.y, 0));
        face(Vec3(lo.x, hi.y, lo.z), Vec3(d.x, 0, 0), Vec3(0, -d.y, 0));
        face(Vec3(lo.x, lo.y, lo.z), Vec3(d.x, 0, 0), Vec3(0, 0, d.z));
        face(Vec3(hi.x, hi.y, lo.z), Vec3(-d.x, 0, 0), Vec3(0, 0, d.z));
        face(Vec3(hi.x, lo.y, lo.z), Vec3(0, d.y, 0), Vec3(0, 0, d.z));
        face(Vec3(lo.x, hi.y, lo.z), Vec3(0, -d.y, 0), Vec3(0, 0, d.z));
    };
    box(Vec3(-0.45f, -0.45f, 0.9f), Vec3(0.45f, 0.45f, 1.05f), std::max(1, tess / 4));
    for (int k = 0; k < 4; k++) {
        float lx = (k & 1) ? 0.35f : -0.43f, ly = (k & 2) ? 0.35f : -0.43f;
        box(Vec3(lx, ly, 0.0f), Vec3(lx + 0.08f, ly + 0.08f, 0.9f), 1);
    }
    // curved back rest (two-sided: front and rear surfaces)
    auto back = [=](float s, float t, float off) {
        float a = (s - 0.5f) * 1.2f;
        return Vec3(0.5f * std::sin(a), 0.45f - 0.18f * (1.0f - std::cos(a)) + off, 1.05f + t * 1.1f);
    };
    b.addSurface(tess, tess, [=](float s, float t) { return X(back(1.0f - s, t, 0.0f)); });
    b.addSurface(tess, tess, [=](float s, float t) { return X(back(s, t, 0.06f)); });
}

inline void set_camera(GenScene& g, float ox, float oy, float oz, float lx, float ly, float lz, float fovx) {
    g.camOrigin[0] = ox; g.camOrigin[1] = oy; g.camOrigin[2] = oz;
    g.camLookAt[0] = lx; g.camLookAt[1] = ly; g.camLookAt[2] = lz;
    g.camUp[0] = 0; g.camUp[1] = 0; g.camUp[2] = 1;
    g.fovx = fovx;
}

// conference-like: closed room, long table, chairs, slatted wall, ceiling light panels.
// detail = 8 gives ~0.33 M triangles (the real conference export is ~0.33 M).
inline GenScene GenerateConference(uint32_t seed, int detail) {
    GenScene g;
    g.name = "conference";
    Lcg rng(seed);
    const int wall = add_material(g, "wall", 0.72f, 0.70f, 0.65f, 0.02f, 0.02f, 0.02f, 8.f, "conference_plaster.ppm");
    const int floorM = add_material(g, "floor", 0.5f, 0.5f, 0.5f, 0.10f, 0.10f, 0.10f, 30.f, "conference_carpet.ppm");
    const int wood = add_material(g, "table", 0.45f, 0.28f, 0.14f, 0.25f, 0.25f, 0.25f, 60.f);
    const int fabric = add_material(g, "chair", 0.15f, 0.18f, 0.45f, 0.03f, 0.03f, 0.03f, 10.f);
    const int metal = add_material(g, "slats", 0.35f, 0.35f, 0.38f, 0.40f, 0.40f, 0.40f, 90.f);
    g.textures["conference_plaster.ppm"] = MakeProceduralTexture(64, 64, 2, seed + 1);
    g.textures["conference_carpet.ppm"] = MakeProceduralTexture(64, 64, 1, seed + 2);
    const Vec3 lo(-4.f, -9.f, 0.f), hi(20.f, 11.f, 8.f);
    const int n = 8 * detail;
    { MeshBuilder b; b.addRoom(lo, hi, n, 4 | 2); add_mesh(g, "walls", wall, b); }
    { MeshBuilder b; b.addRoom(lo, hi, n, 1); add_mesh(g, "floor", floorM, b); }
    {   // table: top slab + central plinth
        MeshBuilder b;
        b.addBox(Vec3(1.f, -1.f, 1.45f), Vec3(15.f, 3.f, 1.6f), 4 * detail);
        b.addBox(Vec3(3.f, 0.2f, 0.0f), Vec3(13.f, 1.8f, 1.45f), detail);
        add_mesh(g, "table", wood, b);
    }
    {   // chairs around the table
        MeshBuilder b;
        const int tess = 7 * detail;
        for (int k = 0; k < 8; k++) {
            float x = 2.0f + 1.7f * (float)k + rng.range(-0.1f, 0.1f);
            add_chair(b, Vec3(x, -2.0f + rng.range(-0.15f, 0.15f), 0.f), 3.14159265f + rng.range(-0.2f, 0.2f), tess);
            add_chair(b, Vec3(x, 4.0f + rng.range(-0.15f, 0.15f), 0.f), rng.range(-0.2f, 0.2f), tess);
        }
        add_chair(b, Vec3(0.0f, 1.0f, 0.f), 1.5708f, tess);
        add_chair(b, Vec3(16.0f, 1.0f, 0.f), -1.5708f, tess);
        add_mesh(g, "chairs", fabric, b);
    }
    {   // slatted panel along the back wall (many thin, long triangles: hard on the LBVH)
        MeshBuilder b;
        const int slats = 32 * detail;
        for (int k = 0; k < slats; k++) {
            float x = lo.x + 0.5f + (hi.x - lo.x - 1.0f) * ((float)k + 0.5f) / (float)slats;
            float tilt = rng.range(-0.03f, 0.03f);
            b.addBox(Vec3(x - 0.04f, hi.y - 0.35f + tilt, 1.0f), Vec3(x + 0.04f, hi.y - 0.25f + tilt, 7.0f), 1);
        }
        add_mesh(g, "slats", metal, b);
    }
    {   // ceiling light: 4 x 3 panels in one mesh, facing down
        MeshBuilder b;
        for (int j = 0; j < 3; j++)
            for (int i = 0; i < 4; i++) {
                float x = 0.5f + 4.2f * (float)i, y = -4.5f + 5.0f * (float)j;
                b.addGrid(Vec3(x, y + 1.2f, hi.z - 0.05f), Vec3(2.4f, 0, 0), Vec3(0, -1.2f, 0), 2, 1);
            }
        g.light = b.mesh;
    }
    g.lightIntensity[0] = 17.f; g.lightIntensity[1] = 12.f; g.lightIntensity[2] = 4.f; g.lightIntensity[3] = 0.f;
    set_camera(g, 15.56f, -4.79f, 4.37f, 1.15f, 2.28f, 1.76f, 70.f);  // scene/conference/*.json
    return g;
}

// livingroom-like: glossy floor and table (Phong exponents 10-200), sofa, shelf with books.
inline GenScene GenerateLivingroom(uint32_t seed, int detail) {
    GenScene g;
    g.name = "livingroom";
    Lcg rng(seed);
    const int wall = add_material(g, "wall", 0.75f, 0.72f, 0.68f, 0.0f, 0.0f, 0.0f, 1.f);
    const int floorM = add_material(g, "parquet", 0.5f, 0.5f, 0.5f, 0.30f, 0.30f, 0.30f, 120.f, "livingroom_wood.ppm");
    const int sofa = add_material(g, "sofa", 0.45f, 0.12f, 0.10f, 0.05f, 0.05f, 0.05f, 10.f);
    const int glass = add_material(g, "lacquer", 0.05f, 0.05f, 0.06f, 0.60f, 0.60f, 0.60f, 200.f);
    const int book = add_material(g, "books", 0.30f, 0.35f, 0.20f, 0.10f, 0.10f, 0.10f, 25.f);
    g.textures["livingroom_wood.ppm"] = MakeProceduralTexture(128, 128, 0, seed + 3);
    const Vec3 lo(-3.f, -2.5f, 0.f), hi(3.f, 5.5f, 3.f);
    const int n = 8 * detail;
    { MeshBuilder b; b.addRoom(lo, hi, n, 4 | 2); add_mesh(g, "walls", wall, b); }
    { MeshBuilder b; b.addRoom(lo, hi, n, 1); add_mesh(g, "floor", floorM, b); }
    {   // sofa with rounded cushions (displaced grids)
        MeshBuilder b;
        b.addBox(Vec3(-2.6f, 3.2f, 0.0f), Vec3(-1.6f, 5.2f, 0.45f), 2 * detail);
        b.addBox(Vec3(-2.9f, 3.2f, 0.0f), Vec3(-2.6f, 5.2f, 1.0f), 2 * detail);
        for (int k = 0; k < 2; k++) {
            float y0 = 3.25f + 1.0f * (float)k;
            b.addSurface(6 * detail, 6 * detail, [=](float s, float t) {
                float bump = 0.12f * std::sin(3.14159265f * s) * std::sin(3.14159265f * t);
                return Vec3(-2.55f + 0.9f * s, y0 + 0.9f * t, 0.45f + bump);
            });
        }
        add_mesh(g, "sofa", sofa, b);
    }
    {   // low lacquer table
        MeshBuilder b;
        b.addBox(Vec3(-0.6f, 2.6f, 0.38f), Vec3(0.8f, 3.6f, 0.44f), 2 * detail);
        for (int k = 0; k < 4; k++) {
            float x = (k & 1) ? 0.7f : -0.55f, y = (k & 2) ? 3.5f : 2.65f;
            b.addBox(Vec3(x, y, 0.f), Vec3(x + 0.06f, y + 0.06f, 0.38f), 1);
        }
        add_mesh(g, "table", glass, b);
    }
    {   // shelf with books on the right wall
        MeshBuilder b;
        for (int s = 0; s < 4; s++) {
            float z = 0.4f + 0.55f * (float)s;
            b.addBox(Vec3(2.55f, 1.0f, z), Vec3(2.98f, 4.5f, z + 0.04f), detail);
            float y = 1.05f;
            while (y < 4.4f) {
                float w = rng.range(0.03f, 0.09f), h = rng.range(0.28f, 0.45f);
                b.addBox(Vec3(2.62f, y, z + 0.04f), Vec3(2.95f, y + w, z + 0.04f + h), 1);
                y += w + 0.004f;
            }
        }
        add_mesh(g, "books", book, b);
    }
    {   // ceiling light
        MeshBuilder b;
        b.addGrid(Vec3(-0.8f, 3.0f, hi.z - 0.02f), Vec3(1.6f, 0, 0), Vec3(0, -1.6f, 0), 2, 2);
        g.light = b.mesh;
    }
    g.lightIntensity[0] = 68.f; g.lightIntensity[1] = 48.f; g.lightIntensity[2] = 16.f; g.lightIntensity[3] = 0.f;
    set_camera(g, 0.34f, -1.644f, 1.22f, 0.12f, 3.25f, 0.8957f, 71.f);  // scene/livingroom/*.json
    return g;
}

// buddha-like: one finely tessellated displaced statue (~detail^2 * 16 k triangles; detail 8 -> ~1.05 M)
// in a box, lit by a directional (exponent 50) emitter.
inline GenScene GenerateBuddha(uint32_t seed, int detail) {
    GenScene g;
    g.name = "buddha";
    Lcg rng(seed);
    const int wall = add_material(g, "wall", 0.6f, 0.6f, 0.6f, 0.0f, 0.0f, 0.0f, 1.f);
    const int gold = add_material(g, "statue", 0.55f, 0.40f, 0.12f, 0.30f, 0.25f, 0.10f, 40.f);
    const Vec3 lo(-5.f, -6.f, -1.2f), hi(5.f, 4.f, 6.f);
    { MeshBuilder b; b.addRoom(lo, hi, 4 * detail, 7); add_mesh(g, "room", wall, b); }
    {
        MeshBuilder b;
        float ph[8], fr[8];
        for (int k = 0; k < 8; k++) { ph[k] = rng.range(0.f, 6.28f); fr[k] = (float)(int)rng.range(3.f, 14.f); }
        const int nu = 128 * detail, nv = 64 * detail;
        b.addSurface(nu, nv, [=](float s, float t) {
            float th = 6.28318530718f * s, phi = 3.14159265359f * (1.0f - t);  // t: bottom -> top
            // body profile: wide base, waist, shoulders, head
            float z01 = t;
            float prof = 0.55f + 0.35f * std::sin(3.14159265359f * z01) - 0.18f * std::exp(-60.f * (z01 - 0.72f) * (z01 - 0.72f))
                       + 0.10f * std::exp(-90.f * (z01 - 0.86f) * (z01 - 0.86f));
            float disp = 0.f;
            for (int k = 0; k < 8; k++) disp += 0.012f * std::sin(fr[k] * th + ph[k]) * std::sin((fr[7 - k] + 2.f) * phi + ph[7 - k]);
            float r = (prof + disp) * std::sin(phi) * 1.15f + 0.001f;
            return Vec3(r * std::cos(th), r * std::sin(th), -1.2f + 3.4f * t + 0.02f * std::sin(9.f * th));
        });
        add_mesh(g, "statue", gold, b);
    }
    {
        MeshBuilder b;
        b.addGrid(Vec3(-1.5f, 0.5f, 5.5f), Vec3(3.0f, 0, 0), Vec3(0, -3.0f, 0), 2, 2);
        g.light = b.mesh;
    }
    g.lightIntensity[0] = 80.f; g.lightIntensity[1] = 70.f; g.lightIntensity[2] = 20.f; g.lightIntensity[3] = 50.f;
    set_camera(g, -1.4f, -3.88f, 0.10616f, -0.646f, -0.56f, -0.168f, 85.f);  // scene/buddha/*.json
    return g;
}

inline GenScene GenerateNamed(const std::string& name, uint32_t seed, int detail) {
    if (name == "conference") return GenerateConference(seed, detail);
    if (name == "livingroom") return GenerateLivingroom(seed, detail);
    if (name == "buddha") return GenerateBuddha(seed, detail);
    throw std::runtime_error("unknown scene name " + name + " (conference | livingroom | buddha)");
}

// Build the RtScene directly (identical to what LoadScene() reads back from ExportScene()).
inline shared_ptr<RtScene> ToRtScene(const GenScene& g, float aspect) {
    auto sc = make_shared<RtScene>();
    auto def = make_shared<RtMaterial>();
    def->mLambertReflectance = make_shared<RtTexture>(0.6f, 0.6f, 0.6f, 1.f);
    def->mPhongReflectance = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
    def->mPhongExponent = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
    sc->mMaterials.push_back(def);
    for (const GenMaterial& m : g.materials) {
        auto mat = make_shared<RtMaterial>();
        mat->mLambertReflectance = m.mapKd.empty() ? make_shared<RtTexture>(m.kd[0], m.kd[1], m.kd[2], 1.f) : g.textures.at(m.mapKd);
        mat->mPhongReflectance = make_shared<RtTexture>(m.ks[0], m.ks[1], m.ks[2], 1.f);
        mat->mPhongExponent = make_shared<RtTexture>(m.ns, m.ns, m.ns, 1.f);
        sc->mMaterials.push_back(mat);
    }
    for (size_t k = 0; k < g.meshes.size(); k++) {
        auto m = make_shared<RtMesh>(*g.meshes[k]);
        m->mMatIndex = 1 + g.meshMat[k].second;
        sc->mMeshes.push_back(m);
    }
    // addAreaLight (rtcommon.h:772-798)
    Vec4 li; li.x = g.lightIntensity[0]; li.y = g.lightIntensity[1]; li.z = g.lightIntensity[2]; li.w = g.lightIntensity[3];
    Vec4 pre = li;
    pre.x = li.x * Math::Pi; pre.y = li.y * Math::Pi; pre.z = li.z * Math::Pi;
    auto lm = make_shared<RtMaterial>();
    lm->mLambertReflectance = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
    lm->mPhongReflectance = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
    lm->mPhongExponent = make_shared<RtTexture>(0.f, 0.f, 0.f, 1.f);
    lm->mLightIntensity = pre;
    auto light = make_shared<RtMesh>(*g.light);
    light->mMatIndex = (int32_t)sc->mMaterials.size();
    sc->mMaterials.push_back(lm);
    sc->mMeshes.push_back(light);
    sc->mArealight = make_shared<RtAreaLight>();
    sc->mArealight->mMesh = light;
    sc->mArealight->mLightIntensity = li;
    sc->mArealight->mPrecomputedLightIntensity = pre;
    sc->isAlreadyHaveLightSource = true;
    // camera JSON -> RtStableCamera
    std::ostringstream cj;
    cj.precision(9);
    cj << "{\"origin\":[" << g.camOrigin[0] << "," << g.camOrigin[1] << "," << g.camOrigin[2] << "],\"direction\":[" << g.camLookAt[0] << ","
       << g.camLookAt[1] << "," << g.camLookAt[2] << "],\"up\":[" << g.camUp[0] << "," << g.camUp[1] << "," << g.camUp[2] << "],\"fovx\":" << g.fovx << "}";
    sc->setCamera(make_shared<RtStableCamera>(Json::parse(cj.str()), aspect));
    return sc;
}

inline void write_obj(const std::string& path, const std::string& mtlName, const std::vector<shared_ptr<RtMesh>>& meshes,
                      const std::vector<std::pair<std::string, std::string>>& nameAndMtl) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + path);
    fprintf(f, "# procedural stand-in generated by evplp_b200 (the reference asset is a git-LFS pointer)\n");
    if (!mtlName.empty()) fprintf(f, "mtllib %s\n", mtlName.c_str());
    size_t base = 1;
    for (size_t k = 0; k < meshes.size(); k++) {
        const RtMesh& m = *meshes[k];
        fprintf(f, "o %s\n", nameAndMtl[k].first.c_str());
        for (int32_t i = 0; i < m.mNumVertices; i++) fprintf(f, "v %.9g %.9g %.9g\n", m.mVertices[3 * i], m.mVertices[3 * i + 1], m.mVertices[3 * i + 2]);
        for (int32_t i = 0; i < m.mNumVertices; i++) fprintf(f, "vt %.9g %.9g\n", m.mTexCoords[2 * i], m.mTexCoords[2 * i + 1]);
        if (!nameAndMtl[k].second.empty()) fprintf(f, "usemtl %s\n", nameAndMtl[k].second.c_str());
        for (int32_t t = 0; t < m.mNumTriangles; t++) {
            size_t a = base + m.mTriIndices[3 * t], b = base + m.mTriIndices[3 * t + 1], c = base + m.mTriIndices[3 * t + 2];
            fprintf(f, "f %zu/%zu %zu/%zu %zu/%zu\n", a, a, b, b, c, c);
        }
        base += (size_t)m.mNumVertices;
    }
    fclose(f);
}

inline void write_ppm(const std::string& path, const RtTexture& t) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    fprintf(f, "P6\n%d %d\n255\n", t.mWidth, t.mHeight);
    for (int y = t.mHeight - 1; y >= 0; y--)  // file rows are top-down; mData row 0 is the bottom
        for (int x = 0; x < t.mWidth; x++)
            for (int c = 0; c < 3; c++) fputc((int)(t.mData[((size_t)y * t.mWidth + x) * 4 + c] * 255.0f + 0.5f), f);
    fclose(f);
}

// Technique JSON in the reference's format (scene/<name>/<name>_<variant>.json).
inline std::string TechniqueJson(const GenScene& g, const std::string& variant, int resX, int resY, const std::string& outPrefix) {
    struct V { const char* name; int paths, vplPaths; float radius; const char* mis; bool progressive; bool vsl; const char* clamp; };
    static const V variants[] = {
        {"ours", 300000, 30, 0.003f, nullptr, false, false, nullptr},
        {"ours_progressive", 300000, 30, 0.003f, nullptr, true, false, nullptr},
        {"vpl", 30, 30, 0.0f, "one", false, false, nullptr},
        {"vpl_progressive", 30, 30, 0.0f, "geometryClamp", true, false, "0.05"},
        {"vsl", 100, 100, 0.0f, "one", false, true, nullptr},
        {"vsl_progressive", 100, 100, 0.0f, "one", true, true, nullptr},
        {"pm", 300000, 0, 0.003f, nullptr, false, false, nullptr},
        {"pm_progressive", 300000, 0, 0.003f, nullptr, true, false, nullptr},
        {"ours_clamp", 300000, 30, 0.003f, "geometryClamp", false, false, nullptr},
    };
    const V* v = nullptr;
    for (const V& c : variants) if (variant == c.name) v = &c;
    if (!v) throw std::runtime_error("unknown technique variant " + variant);
    std::ostringstream o;
    o.precision(9);
    o << "{\n    \"resX\": " << resX << ",\n    \"resY\": " << resY << ",\n    \"scene\": [\"" << g.name << "_exported.obj\"],\n";
    o << "    \"arealight\": {\"obj\": \"" << g.name << "_exported_lights.obj\", \"intensity\": [" << g.lightIntensity[0] << ", "
      << g.lightIntensity[1] << ", " << g.lightIntensity[2] << ", " << g.lightIntensity[3] << "]},\n";
    o << "    \"camera\": {\"origin\": [" << g.camOrigin[0] << ", " << g.camOrigin[1] << ", " << g.camOrigin[2] << "], \"direction\": ["
      << g.camLookAt[0] << ", " << g.camLookAt[1] << ", " << g.camLookAt[2] << "], \"up\": [0.0, 0.0, 1.0], \"fovx\": " << g.fovx << "},\n";
    o << "    \"photonfam\": {\n        \"rngOffset\": 0,\n        \"numMaxIteration\": 8,\n        \"timeLimitMs\": 15000.0,\n"
      << "        \"frameMode\": \"accumulate\",\n        \"renderMode\": \"vplpm\",\n"
      << "        \"combinedFilename\": \"" << outPrefix << "_combined.pfm\",\n        \"weightedPhotonFilename\": \"" << outPrefix
      << "_weightedpm.pfm\",\n        \"weightedVplFilename\": \"" << outPrefix << "_weightedvpl.pfm\",\n"
      << "        \"statFilename\": \"" << outPrefix << "_stat.json\",\n        \"useJitter\": true,\n        \"useStat\": true,\n"
      << "        \"numLightPaths\": " << v->paths << ",\n        \"numVplLightPaths\": " << v->vplPaths << ",\n        \"numMaxBounces\": 3,\n"
      << "        \"radiusPercentage\": " << v->radius << ",\n";
    if (v->mis) o << "        \"misMode\": \"" << v->mis << "\",\n";
    if (v->clamp) o << "        \"clampingCoeff\": " << v->clamp << ",\n";
    if (v->vsl) o << "        \"forceVsl\": true,\n        \"vslRadiusPercentage\": 0.05,\n";
    o << "        \"DoProgressive\": " << (v->progressive ? "true" : "false") << ",\n        \"AlphaProgressive\": 0.7\n    }\n}\n";
    return o.str();
}

inline void ExportScene(const GenScene& g, const std::string& dir, int resX = 1280, int resY = 720) {
    mkdir(dir.c_str(), 0755);
    const std::string base = dir + "/" + g.name;
    {   // MTL
        FILE* f = fopen((base + "_exported.mtl").c_str(), "w");
        if (!f) throw std::runtime_error("cannot write " + base + "_exported.mtl");
        for (const GenMaterial& m : g.materials) {
            fprintf(f, "newmtl %s\nKd %.9g %.9g %.9g\nKs %.9g %.9g %.9g\nNs %.9g\n", m.name.c_str(), m.kd[0], m.kd[1], m.kd[2], m.ks[0], m.ks[1], m.ks[2], m.ns);
            if (!m.mapKd.empty()) fprintf(f, "map_Kd %s\n", m.mapKd.c_str());
            fprintf(f, "\n");
        }
        fclose(f);
    }
    for (auto& kv : g.textures) write_ppm(dir + "/" + kv.first, *kv.second);
    std::vector<std::pair<std::string, std::string>> names;
    for (auto& mm : g.meshMat) names.push_back({mm.first, g.materials[mm.second].name});
    write_obj(base + "_exported.obj", g.name + "_exported.mtl", g.meshes, names);
    write_obj(base + "_exported_lights.obj", "", {g.light}, {{"light", ""}});
    for (const char* variant : {"ours", "ours_progressive", "vpl", "vpl_progressive", "vsl", "vsl_progressive", "pm", "pm_progressive", "ours_clamp"}) {
        std::ofstream of(base + "_" + variant + ".json");
        of << TechniqueJson(g, variant, resX, resY, g.name + "_" + variant);
    }
    for (const char* variant : {"pt", "pt_progressive"}) {  // scene/*/*_pt.json: the path-traced reference images
        std::string text = TechniqueJson(g, "ours", resX, resY, g.name + "_" + variant);
        const size_t a = text.find("    \"photonfam\"");
        std::ostringstream o;
        o << text.substr(0, a) << "    \"pt\": {\n        \"rngOffset\": 0,\n        \"numMaxIteration\": 64,\n        \"timeLimitMs\": 15000.0,\n"
          << "        \"frameMode\": \"accumulate\",\n        \"outputFilename\": \"" << g.name << "_" << variant << ".pfm\",\n"
          << "        \"statFilename\": \"" << g.name << "_" << variant << "_stat.json\",\n        \"useJitter\": true,\n        \"useStat\": true,\n"
          << "        \"numSamplePerPixel\": 1,\n        \"numMaxBounces\": 3,\n        \"DoProgressive\": false,\n        \"AlphaProgressive\": 0.7\n    }\n}\n";
        std::ofstream of(base + "_" + variant + ".json");
        of << o.str();
    }
}

}  // namespace evplp_host
