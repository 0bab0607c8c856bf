// oracle_scene.h -- TEST INFRASTRUCTURE (CPU oracle): RNG streams, scene arrays,
// software texture fetch, ray/triangle test and a plain median-split BVH.
// Parity status: pinned by the reference's own device code for the CUDA programs, see oracle_math.h.
#pragma once
#include <algorithm>
#include <vector>
#include "../include/evplp.h"  // POD descriptors only (EvplpRecord, EvplpParams, ...)
#include "oracle_math.h"

// cuRAND's host-side skip-ahead tables (CUDA toolkit header, same on both machines).
// (the header also declares __device__ copies; neutralise the qualifier for plain g++)
#ifndef __CUDACC__
#define __device__
#include <curand_precalc.h>
#undef __device__
#else
#include <curand_precalc.h>
#endif

namespace orc {

// ------------------------------------------------------------------------------
// cuRAND XORWOW, restated from the CUDA toolkit's curand_kernel.h
// (_curand_init_scratch, _skipahead_sequence_scratch, curand(), _curand_uniform).
// Reference call sites: lighttracing.cu:202-203 (paths), 710-711 (VSL pixels),
// lvclighttracing.cu:369 (LVC pixels): curand_init(launchId, rngSeed, 0, &state).
// ------------------------------------------------------------------------------
struct CurandState {
    unsigned int d, v[5];
};

inline void curand_matvec(const unsigned int* vector, const unsigned int* matrix, unsigned int* result, int n) {
    for (int i = 0; i < n; i++) result[i] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 32; j++)
            if (vector[i] & (1u << j))
                for (int k = 0; k < n; k++) result[k] ^= matrix[n * (i * 32 + j) + k];
}

inline void curand_init(unsigned long long seed, unsigned long long subsequence, unsigned long long offset,
                        CurandState* state) {
    unsigned int s0 = ((unsigned int)seed) ^ 0xaad26b49UL;
    unsigned int s1 = (unsigned int)(seed >> 32) ^ 0xf7dcefddUL;
    unsigned int t0 = 1099087573UL * s0;
    unsigned int t1 = 2591861531UL * s1;
    state->d = 6615241 + t1 + t0;
    state->v[0] = 123456789UL + t0;
    state->v[1] = 362436069UL ^ t0;
    state->v[2] = 521288629UL + t1;
    state->v[3] = 88675123UL ^ t1;
    state->v[4] = 5783321UL + t0;
    // skipahead_sequence: one table per base-4 digit of `subsequence`
    unsigned long long p = subsequence;
    int matrix_num = 0;
    unsigned int result[5];
    while (p && matrix_num < PRECALC_NUM_MATRICES) {
        for (unsigned int t = 0; t < (p & PRECALC_BLOCK_MASK); t++) {
            curand_matvec(state->v, precalc_xorwow_matrix_host[matrix_num], result, 5);
            for (int i = 0; i < 5; i++) state->v[i] = result[i];
        }
        p >>= PRECALC_BLOCK_SIZE;
        matrix_num++;
    }
    // (the oracle only supports subsequence < 4^32 and offset == 0, which is all the path uses)
    (void)offset;
}

inline unsigned int curand(CurandState* state) {
    unsigned int t = (state->v[0] ^ (state->v[0] >> 2));
    state->v[0] = state->v[1];
    state->v[1] = state->v[2];
    state->v[2] = state->v[3];
    state->v[3] = state->v[4];
    state->v[4] = (state->v[4] ^ (state->v[4] << 4)) ^ (t ^ (t << 1));
    state->d += 362437;
    return state->v[4] + state->d;
}

inline float curand_uniform(CurandState* state) {
    const float CURAND_2POW32_INV = 2.3283064e-10f;
    return curand(state) * CURAND_2POW32_INV + (CURAND_2POW32_INV / 2.0f);
}

// ------------------------------------------------------------------------------
// Scene arrays (the same inputs the product receives through evplp_upload_scene).
// ------------------------------------------------------------------------------
struct Texture {
    int w = 1, h = 1;
    std::vector<float> data;  // RGBA32F, row 0 = v near 0 (RtTexture::mData after stb flip)
};

struct Material {
    Texture lambert, phong, exponent;
    float lightIntensity[4];
};

struct Tri {
    F3 p0, p1, p2;
    F2 t0, t1, t2;
    int mat;
};

struct BvhNode {
    float lo[3], hi[3];
    int left, right;   // internal: child node indices; leaf: left = first, right = -count
};

struct Scene {
    std::vector<Tri> tris;        // global primitive order = mesh order, then triangle order
    std::vector<Material> mats;
    int lightFirst = 0, lightCount = 0;   // primitive range of the area-light mesh
    std::vector<float> lightCdf;          // rtcommon.h:501-531
    float lightArea = 0.f;
    float lightIntensity[4];              // pi-scaled (areaLightIntensity)
    float lightDisplay[4];
    std::vector<BvhNode> nodes;
    std::vector<int> order;               // primitive ids in BVH leaf order
    bool bruteForce = false;
};

// Software bilinear fetch with repeat wrap (replaces tex2D on RT_WRAP_REPEAT /
// RT_FILTER_LINEAR samplers, rtcommon.h:225-244, and GL_LINEAR/GL_REPEAT, 203-208).
// Defined in full float precision (hardware uses 8-bit weights).
inline void tex2D(const Texture& t, float u, float v, float out[4]) {
    if (t.w == 1 && t.h == 1) {
        for (int c = 0; c < 4; c++) out[c] = t.data[c];
        return;
    }
    float x = u * (float)t.w - 0.5f;
    float y = v * (float)t.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = (int)fx % t.w; if (i0 < 0) i0 += t.w;
    int j0 = (int)fy % t.h; if (j0 < 0) j0 += t.h;
    int i1 = i0 + 1; if (i1 == t.w) i1 = 0;
    int j1 = j0 + 1; if (j1 == t.h) j1 = 0;
    const float* t00 = &t.data[(size_t)(j0 * t.w + i0) * 4];
    const float* t10 = &t.data[(size_t)(j0 * t.w + i1) * 4];
    const float* t01 = &t.data[(size_t)(j1 * t.w + i0) * 4];
    const float* t11 = &t.data[(size_t)(j1 * t.w + i1) * 4];
    for (int c = 0; c < 4; c++) {
        float lo = t00[c] + a * (t10[c] - t00[c]);
        float hi = t01[c] + a * (t11[c] - t01[c]);
        out[c] = lo + b * (hi - lo);
    }
}

// optix::intersect_triangle_branchless (SDK header), as called from
// meshFineIntersect (triangleintersect.cu:17-41).
inline bool intersect_triangle_branchless(const F3& org, const F3& dir, float tmin, float tmax, const F3& p0,
                                          const F3& p1, const F3& p2, F3& n, float& t, float& beta, float& gamma) {
    const F3 e0 = p1 - p0;
    const F3 e1 = p0 - p2;
    n = cross(e1, e0);
    const F3 e2 = (1.0f / dot(n, dir)) * (p0 - org);
    const F3 i = cross(dir, e2);
    beta = dot(i, e1);
    gamma = dot(i, e0);
    t = dot(n, e2);
    return ((t < tmax) & (t > tmin) & (beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1));
}

struct Hit {
    int prim = -1;
    float t = 0.f, beta = 0.f, gamma = 0.f;
    F3 n;  // un-normalised cross(e1, e0)
};

inline bool box_hit(const BvhNode& nd, const F3& org, const F3& inv, float tmin, float tmax, float pad) {
    // conservative slab test in double with a padded box: never rejects a box whose
    // triangles the float triangle test could accept
    double t0 = tmin, t1 = tmax;
    const float o[3] = {org.x, org.y, org.z};
    const float iv[3] = {inv.x, inv.y, inv.z};
    for (int a = 0; a < 3; a++) {
        double lo = ((double)nd.lo[a] - pad - o[a]) * iv[a];
        double hi = ((double)nd.hi[a] + pad - o[a]) * iv[a];
        if (lo != lo || hi != hi) continue;  // 0 * inf: ray parallel and on the slab plane
        if (lo > hi) std::swap(lo, hi);
        if (lo > t0) t0 = lo;
        if (hi < t1) t1 = hi;
    }
    return t0 <= t1 * (1.0 + 1e-9) + 1e-30;
}

// Closest hit over ALL triangles passing the float test; ties -> smallest primitive id
// (the reference's Trbvh order is unpinned; SURVEY.md §A.5 last bullet).
inline Hit trace_closest(const Scene& s, const F3& org, const F3& dir, float tmin, float tmax) {
    Hit best;
    float bestT = tmax;
    auto test = [&](int prim) {
        const Tri& tr = s.tris[prim];
        F3 n; float t, b, g;
        if (intersect_triangle_branchless(org, dir, tmin, tmax, tr.p0, tr.p1, tr.p2, n, t, b, g)) {
            if (best.prim < 0 ? (t < bestT) : (t < bestT || (t == bestT && prim < best.prim))) {
                best.prim = prim; best.t = t; best.beta = b; best.gamma = g; best.n = n;
                bestT = t;
            }
        }
    };
    if (s.bruteForce || s.nodes.empty()) {
        for (int i = 0; i < (int)s.tris.size(); i++) test(i);
        return best;
    }
    F3 inv = mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    float pad = 0.f;
    {
        const BvhNode& r = s.nodes[0];
        float m = 0.f;
        for (int a = 0; a < 3; a++) m = fmaxf(m, fmaxf(fabsf(r.lo[a]), fabsf(r.hi[a])));
        pad = m * 1e-5f + 1e-20f;
    }
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BvhNode& nd = s.nodes[stack[--sp]];
        // bestT may only shrink; ties need t <= bestT, box_hit is inclusive with padding
        if (!box_hit(nd, org, inv, tmin, bestT, pad)) continue;
        if (nd.right < 0) {
            for (int k = 0; k < -nd.right; k++) test(s.order[nd.left + k]);
        } else {
            stack[sp++] = nd.left;
            stack[sp++] = nd.right;
        }
    }
    return best;
}

inline bool trace_any(const Scene& s, const F3& org, const F3& dir, float tmin, float tmax) {
    auto test = [&](int prim) {
        const Tri& tr = s.tris[prim];
        F3 n; float t, b, g;
        return intersect_triangle_branchless(org, dir, tmin, tmax, tr.p0, tr.p1, tr.p2, n, t, b, g);
    };
    if (s.bruteForce || s.nodes.empty()) {
        for (int i = 0; i < (int)s.tris.size(); i++) if (test(i)) return true;
        return false;
    }
    F3 inv = mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    float pad = 0.f;
    {
        const BvhNode& r = s.nodes[0];
        float m = 0.f;
        for (int a = 0; a < 3; a++) m = fmaxf(m, fmaxf(fabsf(r.lo[a]), fabsf(r.hi[a])));
        pad = m * 1e-5f + 1e-20f;
    }
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const BvhNode& nd = s.nodes[stack[--sp]];
        if (!box_hit(nd, org, inv, tmin, tmax, pad)) continue;
        if (nd.right < 0) {
            for (int k = 0; k < -nd.right; k++) if (test(s.order[nd.left + k])) return true;
        } else {
            stack[sp++] = nd.left;
            stack[sp++] = nd.right;
        }
    }
    return false;
}

// Median-split BVH over primitive boxes (independent of the product's LBVH: hits do not
// depend on the tree as long as culling is conservative).
inline void build_bvh(Scene& s) {
    const int n = (int)s.tris.size();
    s.nodes.clear(); s.order.resize(n);
    if (n == 0) return;
    std::vector<float> cx(n * 3), blo(n * 3), bhi(n * 3);
    for (int i = 0; i < n; i++) {
        s.order[i] = i;
        const Tri& t = s.tris[i];
        const float p[3][3] = {{t.p0.x, t.p0.y, t.p0.z}, {t.p1.x, t.p1.y, t.p1.z}, {t.p2.x, t.p2.y, t.p2.z}};
        for (int a = 0; a < 3; a++) {
            blo[i * 3 + a] = fminf(fminf(p[0][a], p[1][a]), p[2][a]);
            bhi[i * 3 + a] = fmaxf(fmaxf(p[0][a], p[1][a]), p[2][a]);
            cx[i * 3 + a] = 0.5f * (blo[i * 3 + a] + bhi[i * 3 + a]);
        }
    }
    struct Item { int node, first, count; };
    std::vector<Item> todo;
    s.nodes.push_back(BvhNode());
    todo.push_back({0, 0, n});
    while (!todo.empty()) {
        Item it = todo.back(); todo.pop_back();
        BvhNode nd;
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int a = 0; a < 3; a++) { nd.lo[a] = INFINITY; nd.hi[a] = -INFINITY; }
        for (int k = 0; k < it.count; k++) {
            int p = s.order[it.first + k];
            for (int a = 0; a < 3; a++) {
                nd.lo[a] = fminf(nd.lo[a], blo[p * 3 + a]); nd.hi[a] = fmaxf(nd.hi[a], bhi[p * 3 + a]);
                clo[a] = fminf(clo[a], cx[p * 3 + a]); chi[a] = fmaxf(chi[a], cx[p * 3 + a]);
            }
        }
        int axis = 0; float ext = chi[0] - clo[0];
        for (int a = 1; a < 3; a++) if (chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
        if (it.count <= 4 || !(ext > 0.f)) {
            nd.left = it.first; nd.right = -it.count;
            s.nodes[it.node] = nd;
            continue;
        }
        int mid = it.count / 2;
        std::nth_element(s.order.begin() + it.first, s.order.begin() + it.first + mid,
                         s.order.begin() + it.first + it.count,
                         [&](int a, int b) { return cx[a * 3 + axis] < cx[b * 3 + axis] || (cx[a * 3 + axis] == cx[b * 3 + axis] && a < b); });
        nd.left = (int)s.nodes.size(); nd.right = nd.left + 1;
        s.nodes[it.node] = nd;
        s.nodes.push_back(BvhNode()); s.nodes.push_back(BvhNode());
        todo.push_back({nd.left, it.first, mid});
        todo.push_back({nd.right, it.first + mid, it.count - mid});
    }
}

}  // namespace orc
