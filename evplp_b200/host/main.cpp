// main.cpp -- the CLI of the reference (reflectcuts/main.cpp:87-124), headless.
//   evplp_render <scene.json> [--device N]                 render with the technique the JSON names
//   evplp_render --gen <conference|livingroom|buddha> <outdir> [--seed S] [--detail D] [--res W H]
//                                                          write the procedural stand-in assets + JSONs
#include <cstdlib>
#include <cstring>
#include <iostream>
#include "rtcomphoton.h"
#include "rtpt2.h"
#include "scenegen.h"

using namespace evplp_host;

int main(int numArg, const char* args[]) {
    try {
        if (numArg > 1 && strcmp(args[1], "--gen") == 0) {
            if (numArg < 4) { std::cerr << "usage: evplp_render --gen <name> <outdir> [--seed S] [--detail D] [--res W H]\n"; return 2; }
            uint32_t seed = 1; int detail = 8, resX = 1280, resY = 720;
            for (int i = 4; i < numArg; i++) {
                if (!strcmp(args[i], "--seed") && i + 1 < numArg) seed = (uint32_t)atoi(args[++i]);
                else if (!strcmp(args[i], "--detail") && i + 1 < numArg) detail = atoi(args[++i]);
                else if (!strcmp(args[i], "--res") && i + 2 < numArg) { resX = atoi(args[++i]); resY = atoi(args[++i]); }
            }
            GenScene g = GenerateNamed(args[2], seed, detail);
            ExportScene(g, args[3], resX, resY);
            size_t tris = 0;
            for (auto& m : g.meshes) tris += (size_t)m->mNumTriangles;
            std::cout << "wrote " << args[3] << "/" << g.name << "_exported.obj (" << tris << " triangles)\n";
            return 0;
        }
        std::string jsonFilename = numArg > 1 ? args[1] : "../scene/conference/conference_ours.json";  // main.cpp:89-98
        int device = 0;
        for (int i = 2; i < numArg; i++) if (!strcmp(args[i], "--device") && i + 1 < numArg) device = atoi(args[++i]);
        Json json = Json::parse_file(jsonFilename);
        shared_ptr<RtScene> scene = LoadScene(json, jsonFilename);
        if (!scene) { std::cerr << "no \"scene\" in " << jsonFilename << "\n"; return 1; }
        Vec2 res; res.x = json["resX"].as_float(); res.y = json["resY"].as_float();
        if (!json["pt"].is_null()) {  // main.cpp:105-109
            RtPt2 rtpt(device);
            rtpt.render(scene, res, json["pt"]);
            std::cout << "pt: " << rtpt.numIterations() << " iterations in " << rtpt.elapsedMs() << " ms\n";
        }
        if (!json["photonfam"].is_null()) {
            RtComPhoton rtcomp(device);
            rtcomp.render(scene, res, json["photonfam"]);
            std::cout << "photonfam: " << rtcomp.numIterations() << " iterations in " << rtcomp.elapsedMs() << " ms\n";
        }
        if (!json["lvcphotonfam"].is_null()) {
            RtLvcComPhoton rtcomp2(device);
            rtcomp2.render(scene, res, json["lvcphotonfam"]);
            std::cout << "lvcphotonfam: " << rtcomp2.numIterations() << " iterations in " << rtcomp2.elapsedMs() << " ms\n";
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}
