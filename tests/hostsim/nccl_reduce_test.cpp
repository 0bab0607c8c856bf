// nccl_reduce_test.cpp -- exercises evplp_reduce() with real ncclComm_t handles: one process, two GPUs
// (ncclCommInitAll), each handle holds different accumulation layers, after the grouped reduce both hold the sum.
// Built and run by tests/test_gpu_multi.py on boxes with >= 2 GPUs.
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/evplp.h"

#define CHECK(x) do { if (!(x)) { fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #x, evplp_last_error()); return 1; } } while (0)

int main() {
    int n = 0;
    cudaGetDeviceCount(&n);
    if (n < 2) { printf("SKIP: needs 2 GPUs\n"); return 0; }
    const int W = 64, H = 32;
    evplp_handle h[2];
    ncclComm_t comms[2];
    int devs[2] = {0, 1};
    CHECK(ncclCommInitAll(comms, 2, devs) == ncclSuccess);
    // a one-triangle scene + light so that the handles are fully set up
    const float verts[9] = {0, 0, 0, 1, 0, 0, 0, 1, 0};
    const int32_t idx[3] = {0, 1, 2};
    const float tex[4] = {0.5f, 0.5f, 0.5f, 0.f};
    EvplpMeshDesc mesh[2] = {{verts, nullptr, idx, 3, 1, 0}, {verts, nullptr, idx, 3, 1, 1}};
    EvplpMaterialDesc mat[2];
    for (int m = 0; m < 2; m++) {
        mat[m].lambertReflectance = tex; mat[m].lambertW = mat[m].lambertH = 1;
        mat[m].phongReflectance = tex; mat[m].phongW = mat[m].phongH = 1;
        mat[m].phongExponent = tex; mat[m].exponentW = mat[m].exponentH = 1;
        memset(mat[m].lightIntensity, 0, 16);
    }
    const float li[4] = {1, 1, 1, 0};
    std::vector<float> planes((size_t)4 * W * H * 4, 0.f);
    std::vector<int32_t> prims((size_t)W * H);
    for (int d = 0; d < 2; d++) {
        CHECK(evplp_create(d, W, H, &h[d]) == EVPLP_OK);
        CHECK(evplp_upload_scene(h[d], mesh, 2, mat, 2, 1, li, li) == EVPLP_OK);
        CHECK(evplp_build_bvh(h[d]) == EVPLP_OK);
        // light layer: rank d marks pixels whose index is a multiple of (d + 2) as "light" (primitive 1 = the light mesh)
        for (int i = 0; i < W * H; i++) prims[i] = (i % (d + 2) == 0) ? 1 : 0;
        CHECK(evplp_upload_gbuffer(h[d], planes.data(), prims.data()) == EVPLP_OK);
        CHECK(evplp_light_pass(h[d]) == EVPLP_OK);
        CHECK(evplp_synchronize(h[d]) == EVPLP_OK);
    }
    CHECK(ncclGroupStart() == ncclSuccess);
    for (int d = 0; d < 2; d++) CHECK(evplp_reduce(h[d], comms[d]) == EVPLP_OK);
    CHECK(ncclGroupEnd() == ncclSuccess);
    std::vector<uint32_t> light((size_t)W * H);
    std::vector<int64_t> vpl((size_t)W * H * 3), photon((size_t)W * H * 3);
    for (int d = 0; d < 2; d++) {
        CHECK(evplp_download_accum(h[d], vpl.data(), photon.data(), light.data()) == EVPLP_OK);
        for (int i = 0; i < W * H; i++) {
            const uint32_t want = (i % 2 == 0 ? 1u : 0u) + (i % 3 == 0 ? 1u : 0u);
            if (light[i] != want) { fprintf(stderr, "rank %d pixel %d: %u != %u\n", d, i, light[i], want); return 1; }
        }
        for (int64_t v : vpl) if (v != 0) return 1;
    }
    for (int d = 0; d < 2; d++) { evplp_destroy(h[d]); ncclCommDestroy(comms[d]); }
    printf("OK: evplp_reduce over 2 GPUs sums the layers\n");
    return 0;
}
