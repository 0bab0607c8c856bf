// nccl_reduce_test.cpp -- exercises evplp_reduce() with real ncclComm_t handles: one process, two GPUs
// (ncclCommInitAll), each handle holds different accumulation layers (written through the evplp_accum_layer pointers) and
// iteration counts; after the grouped reduce both hold the sums, exactly.
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
    const size_t px = (size_t)W * H;
    std::vector<int64_t> vplIn(px * 3), photonIn(px * 3);
    std::vector<uint32_t> lightIn(px);
    for (int d = 0; d < 2; d++) {
        CHECK(evplp_create(d, W, H, &h[d]) == EVPLP_OK);
        CHECK(evplp_upload_scene(h[d], mesh, 2, mat, 2, 1, li, li) == EVPLP_OK);
        CHECK(evplp_build_bvh(h[d]) == EVPLP_OK);
        CHECK(evplp_clear_accum(h[d]) == EVPLP_OK);
        CHECK(evplp_add_iterations(h[d], d + 3) == EVPLP_OK);   // the iteration counter travels with the layers
        CHECK(evplp_synchronize(h[d]) == EVPLP_OK);
        // rank d's layers: distinct, sign-mixed fixed-point values written straight into the device buffers
        // (evplp_accum_layer hands out the pointers a host that owns a communicator would reduce itself)
        for (size_t i = 0; i < px * 3; i++) {
            vplIn[i] = (int64_t)(i + 1) * (d == 0 ? 1000003ll : -7ll) + ((int64_t)d << 40);
            photonIn[i] = (int64_t)(i % 97) * (d + 1) - 11;
        }
        for (size_t i = 0; i < px; i++) lightIn[i] = (i % (size_t)(d + 2) == 0) ? 1u : 0u;
        void* p = nullptr; uint64_t n = 0;
        CHECK(cudaSetDevice(d) == cudaSuccess);
        CHECK(evplp_accum_layer(h[d], 0, &p, &n) == EVPLP_OK && n == px * 3);
        CHECK(cudaMemcpy(p, vplIn.data(), n * 8, cudaMemcpyHostToDevice) == cudaSuccess);
        CHECK(evplp_accum_layer(h[d], 1, &p, &n) == EVPLP_OK && n == px * 3);
        CHECK(cudaMemcpy(p, photonIn.data(), n * 8, cudaMemcpyHostToDevice) == cudaSuccess);
        CHECK(evplp_accum_layer(h[d], 2, &p, &n) == EVPLP_OK && n == px);
        CHECK(cudaMemcpy(p, lightIn.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess);
    }
    CHECK(ncclGroupStart() == ncclSuccess);
    for (int d = 0; d < 2; d++) CHECK(evplp_reduce(h[d], comms[d]) == EVPLP_OK);
    CHECK(ncclGroupEnd() == ncclSuccess);
    std::vector<uint32_t> light(px);
    std::vector<int64_t> vpl(px * 3), photon(px * 3);
    for (int d = 0; d < 2; d++) {
        CHECK(evplp_synchronize(h[d]) == EVPLP_OK);
        CHECK(evplp_download_accum(h[d], vpl.data(), photon.data(), light.data()) == EVPLP_OK);
        for (size_t i = 0; i < px * 3; i++) {
            const int64_t wantV = (int64_t)(i + 1) * 1000003ll + (int64_t)(i + 1) * -7ll + ((int64_t)1 << 40);
            const int64_t wantP = (int64_t)(i % 97) * 3 - 22;
            if (vpl[i] != wantV || photon[i] != wantP) { fprintf(stderr, "rank %d element %zu: layers differ from the sum\n", d, i); return 1; }
        }
        for (size_t i = 0; i < px; i++) {
            const uint32_t want = (i % 2 == 0 ? 1u : 0u) + (i % 3 == 0 ? 1u : 0u);
            if (light[i] != want) { fprintf(stderr, "rank %d pixel %zu: %u != %u\n", d, i, light[i], want); return 1; }
        }
        int64_t iters = 0;
        CHECK(evplp_iterations(h[d], &iters) == EVPLP_OK);
        if (iters != 7) { fprintf(stderr, "rank %d: iteration count %lld != 7\n", d, (long long)iters); return 1; }
    }
    for (int d = 0; d < 2; d++) { evplp_destroy(h[d]); ncclCommDestroy(comms[d]); }
    printf("OK: evplp_reduce over 2 GPUs sums the layers\n");
    return 0;
}
