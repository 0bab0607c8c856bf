// xorwow.h -- cuRAND XORWOW streams reproduced on device without curand_init's
// per-thread matrix walk.
//
// Reference semantics (lighttracing.cu:202-203, 710-711; lvclighttracing.cu:369):
//   curand_init(seed = launch index, subsequence = rngSeed, offset = 0)
//   curand_uniform(state) in (0, 1]
// cuRAND's XORWOW (CUDA toolkit curand_kernel.h, _curand_init_scratch / curand()):
// the seed is salted into five xorshift words + a Weyl counter d, then the state is
// advanced by subsequence * 2^67 draws by multiplying the 160-bit xorshift state with
// precomputed GF(2) matrices M^(2^67 * 4^k) once per base-4 digit of `subsequence`.
// Because `subsequence` is launch-uniform here (it is the iteration number), the host
// composes ONE 160x160 matrix per launch (evplp::xorwow_compose_skip in
// host_xorwow.cpp) and every thread applies a single branch-free mat-vec
// (SURVEY.md §7 H5).  d does not change (2^67 * anything is 0 mod 2^32).
#pragma once
#include "detmath.h"

namespace evplp {

struct Xorwow {
    uint32_t v0, v1, v2, v3, v4, d;
};

constexpr int kSkipMatrixWords = 800;  // 160 rows x 5 words, row-major (cuRAND layout)

EVPLP_HD Xorwow xorwow_seed(uint32_t seed) {
    // 64-bit seed with a zero high word (the reference passes a 32-bit launch index).
    uint32_t s0 = seed ^ 0xaad26b49u;
    uint32_t s1 = 0u ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    Xorwow s;
    s.d = 6615241u + t1 + t0;
    s.v0 = 123456789u + t0;
    s.v1 = 362436069u ^ t0;
    s.v2 = 521288629u + t1;
    s.v3 = 88675123u ^ t1;
    s.v4 = 5783321u + t0;
    return s;
}

// state <- state * M.  `m` may live in shared, constant or global memory; all threads
// read the same row at the same time (broadcast).
EVPLP_HD void xorwow_apply_matrix(Xorwow& s, const uint32_t* m) {
    uint32_t in[5] = {s.v0, s.v1, s.v2, s.v3, s.v4};
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    for (int i = 0; i < 5; i++) {
        uint32_t w = in[i];
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
        for (int j = 0; j < 32; j++) {
            uint32_t mask = 0u - ((w >> j) & 1u);
            const uint32_t* row = m + 5 * (i * 32 + j);
            r0 ^= row[0] & mask;
            r1 ^= row[1] & mask;
            r2 ^= row[2] & mask;
            r3 ^= row[3] & mask;
            r4 ^= row[4] & mask;
        }
    }
    s.v0 = r0; s.v1 = r1; s.v2 = r2; s.v3 = r3; s.v4 = r4;
}

// ---- the same mat-vec through 4-bit tables --------------------------------------------------------------------------
// Of the seeded state only v0, v1, v4 depend on the seed (xorwow_seed: v2, v3 come from the constant high seed word), so
// state * M = K ^ rows selected by the 96 varying bits, K = the contribution of the constant words.  The host folds the
// matrix into 24 tables (one per nibble of v0, v1, v4) of 16 entries x 5 words (padded to 8): 24 look-ups of 32 bytes
// replace 160 row selections (~240 instead of ~2000 instructions per path).  K is folded into table 0 (exactly one of
// its entries is always selected).  Layout: tab[((w * 8 + nibble) * 16 + value) * 8 + k], w = 0, 1, 2 for v0, v1, v4.
constexpr int kSkipTableWords = 24 * 16 * 8;

inline void xorwow_build_tables(const uint32_t* m /* 800 words */, uint32_t* tab /* kSkipTableWords */) {
    const Xorwow z = xorwow_seed(0u);   // v2, v3 do not depend on the seed
    uint32_t K[5] = {0, 0, 0, 0, 0};
    const uint32_t cw[2] = {z.v2, z.v3};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 32; j++)
            if ((cw[i] >> j) & 1u)
                for (int k = 0; k < 5; k++) K[k] ^= m[5 * ((2 + i) * 32 + j) + k];
    const int wordOf[3] = {0, 1, 4};
    for (int w = 0; w < 3; w++)
        for (int n = 0; n < 8; n++)
            for (int x = 0; x < 16; x++) {
                uint32_t* e = tab + ((w * 8 + n) * 16 + x) * 8;
                for (int k = 0; k < 8; k++) e[k] = 0u;
                for (int b = 0; b < 4; b++)
                    if ((x >> b) & 1)
                        for (int k = 0; k < 5; k++) e[k] ^= m[5 * (wordOf[w] * 32 + 4 * n + b) + k];
                if (w == 0 && n == 0)
                    for (int k = 0; k < 5; k++) e[k] ^= K[k];
            }
}

// xorwow_seed(seed) followed by xorwow_apply_matrix, through the tables (bit-identical)
EVPLP_HD Xorwow xorwow_seed_skip(uint32_t seed, const uint32_t* tab) {
    Xorwow s = xorwow_seed(seed);
    const uint32_t in[3] = {s.v0, s.v1, s.v4};
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int w = 0; w < 3; w++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int n = 0; n < 8; n++) {
            const uint32_t x = (in[w] >> (4 * n)) & 15u;
            const uint32_t* e = tab + ((w * 8 + n) * 16 + x) * 8;
#if defined(__CUDA_ARCH__)
            const uint4 a = *reinterpret_cast<const uint4*>(e);
            r0 ^= a.x; r1 ^= a.y; r2 ^= a.z; r3 ^= a.w; r4 ^= e[4];
#else
            r0 ^= e[0]; r1 ^= e[1]; r2 ^= e[2]; r3 ^= e[3]; r4 ^= e[4];
#endif
        }
    }
    s.v0 = r0; s.v1 = r1; s.v2 = r2; s.v3 = r3; s.v4 = r4;
    return s;
}

EVPLP_HD uint32_t xorwow_next(Xorwow& s) {
    uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1;
    s.v1 = s.v2;
    s.v2 = s.v3;
    s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return s.v4 + s.d;
}

// curand_uniform: x * 2^-32 + 2^-33 in float (the multiply is exact, so fused or not
// gives the same bits).
EVPLP_HD float xorwow_uniform(Xorwow& s) {
    uint32_t x = xorwow_next(s);
    return (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

}  // namespace evplp
