// host_xorwow.cpp -- composes the single GF(2) skip matrix of one launch.
//
// curand_init(seed, subsequence, 0) (reference call sites lighttracing.cu:203, 711;
// lvclighttracing.cu:369) advances the seeded xorshift state by subsequence * 2^67 draws:
// cuRAND multiplies the 160-bit state by the precomputed matrix of base-4 digit k,
// `digit` times, for every digit of `subsequence` (curand_kernel.h,
// _skipahead_sequence_scratch).  On this path `subsequence` is the iteration number --
// uniform over the launch -- so the product of those matrices is formed ONCE here and each
// thread does one mat-vec (xorwow.h, xorwow_apply_matrix) instead of up to 48.
#include <stdint.h>
#include <string.h>
// cuRAND's own tables (CUDA toolkit).  The header declares __device__ copies too; a plain
// C++ translation unit only needs the host one.
#define __device__
#include <curand_precalc.h>
#undef __device__
#include "xorwow.h"

namespace evplp {

static void vecmat(const uint32_t* v, const uint32_t* m, uint32_t* out) {
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 32; j++)
            if (v[i] & (1u << j)) {
                const uint32_t* row = m + 5 * (i * 32 + j);
                for (int k = 0; k < 5; k++) r[k] ^= row[k];
            }
    memcpy(out, r, sizeof(r));
}

// c = a * b (row-vector convention: state' = state * M)
static void matmat(const uint32_t* a, const uint32_t* b, uint32_t* c) {
    uint32_t tmp[kSkipMatrixWords];
    for (int row = 0; row < 160; row++) vecmat(a + 5 * row, b, tmp + 5 * row);
    memcpy(c, tmp, sizeof(tmp));
}

void xorwow_compose_skip(uint32_t subsequence, uint32_t* out800) {
    // identity
    memset(out800, 0, sizeof(uint32_t) * kSkipMatrixWords);
    for (int row = 0; row < 160; row++) out800[5 * row + row / 32] = 1u << (row % 32);
    uint32_t p = subsequence;
    int matrixNum = 0;
    while (p) {  // 32-bit subsequence: at most 16 base-4 digits, all inside the precalc table
        for (uint32_t t = 0; t < (p & PRECALC_BLOCK_MASK); t++)
            matmat(out800, (const uint32_t*)precalc_xorwow_matrix_host[matrixNum], out800);
        p >>= PRECALC_BLOCK_SIZE;
        matrixNum++;
    }
}

}  // namespace evplp

// Host-only self-check of the table form (xorwow_seed_skip) against the matrix form (xorwow_seed + xorwow_apply_matrix):
// returns the number of seeds in [firstSeed, firstSeed + numSeeds) whose states differ.  No device is touched.
extern "C" int evplp_debug_xorwow_tables(uint32_t subsequence, uint32_t firstSeed, uint32_t numSeeds) {
    static thread_local uint32_t m[evplp::kSkipMatrixWords], tab[evplp::kSkipTableWords];
    evplp::xorwow_compose_skip(subsequence, m);
    evplp::xorwow_build_tables(m, tab);
    int bad = 0;
    for (uint32_t k = 0; k < numSeeds; k++) {
        evplp::Xorwow a = evplp::xorwow_seed(firstSeed + k);
        evplp::xorwow_apply_matrix(a, m);
        const evplp::Xorwow b = evplp::xorwow_seed_skip(firstSeed + k, tab);
        if (a.v0 != b.v0 || a.v1 != b.v1 || a.v2 != b.v2 || a.v3 != b.v3 || a.v4 != b.v4 || a.d != b.d) bad++;
    }
    return bad;
}
