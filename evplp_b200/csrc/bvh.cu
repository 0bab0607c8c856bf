// bvh.cu -- Morton-code LBVH build on the device (replaces the OptiX "Trbvh" acceleration,
// reference rtcomphoton.h:705-707, and the bounds program meshBound, triangleintersect.cu:62-82).
//
// Pipeline (all on the context stream):
//   1 prim_bounds   per-triangle AABB + scene AABB (ordered-uint atomics)
//   2 morton        63-bit codes (21 bits/axis) of the box centre
//   3 CUB radix sort of (code, primId)         -- stable: equal codes keep primitive order
//   4 leaf_records  triangles rewritten in sorted order with e0/e1/n precomputed
//   5 karras        Karras 2012 binary radix tree, one thread per internal node
//   6 refit         bottom-up AABBs (second arriver at a node computes it)
//   7 collapse      binary tree -> BVH_WIDTH-ary nodes, level by level, surface-area guided
// Morton codes, sorted order and the binary topology are bit-identical to the CPU twin in
// the test oracle; the wide nodes are a product-only acceleration (hits do not depend on them).
#include <cub/device/device_radix_sort.cuh>
#include "context.h"

namespace evplp {

__device__ __forceinline__ uint32_t enc_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float dec_ordered(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

__global__ void prim_bounds_kernel(const float4* __restrict__ triVerts, int n, float* __restrict__ primLo,
                                   float* __restrict__ primHi, uint32_t* __restrict__ sceneEnc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (i < n) {
        float4 a = triVerts[3 * (size_t)i], b = triVerts[3 * (size_t)i + 1], c = triVerts[3 * (size_t)i + 2];
        lo[0] = fminf(fminf(a.x, b.x), c.x); hi[0] = fmaxf(fmaxf(a.x, b.x), c.x);
        lo[1] = fminf(fminf(a.y, b.y), c.y); hi[1] = fmaxf(fmaxf(a.y, b.y), c.y);
        lo[2] = fminf(fminf(a.z, b.z), c.z); hi[2] = fmaxf(fmaxf(a.z, b.z), c.z);
        for (int k = 0; k < 3; k++) { primLo[3 * (size_t)i + k] = lo[k]; primHi[3 * (size_t)i + k] = hi[k]; }
    }
    // warp reduce, then one atomic per warp
    for (int k = 0; k < 3; k++) {
        float l = lo[k], h = hi[k];
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (l != INFINITY) atomicMin(&sceneEnc[k], enc_ordered(l));
            if (h != -INFINITY) atomicMax(&sceneEnc[3 + k], enc_ordered(h));
        }
    }
}

__device__ __forceinline__ uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

struct Bounds3 {
    float lo[3], hi[3];
};

__global__ void morton_kernel(const float* __restrict__ primLo, const float* __restrict__ primHi, int n, Bounds3 scene,
                              uint64_t* __restrict__ codes, uint32_t* __restrict__ ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t g[3];
    for (int a = 0; a < 3; a++) {
        float c = (primLo[3 * (size_t)i + a] + primHi[3 * (size_t)i + a]) * 0.5f;
        float ext = scene.hi[a] - scene.lo[a];
        float q = ext > 0.0f ? det_div(c - scene.lo[a], ext) : 0.0f;
        g[a] = (uint32_t)fminf(fmaxf(q * 2097152.0f, 0.0f), 2097151.0f);
    }
    codes[i] = (expand21(g[0]) << 2) | (expand21(g[1]) << 1) | expand21(g[2]);
    ids[i] = (uint32_t)i;
}

__global__ void leaf_records_kernel(const float4* __restrict__ triVerts, const uint32_t* __restrict__ sorted, int n,
                                    float4* __restrict__ triLeaf) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t p = sorted[k];
    float4 a = triVerts[3 * (size_t)p], b = triVerts[3 * (size_t)p + 1], c = triVerts[3 * (size_t)p + 2];
    V3 p0 = v3(a.x, a.y, a.z), p1 = v3(b.x, b.y, b.z), p2 = v3(c.x, c.y, c.z);
    V3 e0 = p1 - p0, e1 = p0 - p2;
    V3 nn = cross(e1, e0);
    triLeaf[4 * (size_t)k + 0] = make_float4(p0.x, p0.y, p0.z, __int_as_float((int)p));
    triLeaf[4 * (size_t)k + 1] = make_float4(e0.x, e0.y, e0.z, a.w);  // a.w carries the material index bits
    triLeaf[4 * (size_t)k + 2] = make_float4(e1.x, e1.y, e1.z, 0.f);
    triLeaf[4 * (size_t)k + 3] = make_float4(nn.x, nn.y, nn.z, 0.f);
}

__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ sc, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = sc[i], b = sc[j];
    if (a == b) return 64 + __clz((int)((uint32_t)i ^ (uint32_t)j));
    return __clzll((long long)(a ^ b));
}

__global__ void karras_kernel(const uint64_t* __restrict__ sc, int n, int32_t* __restrict__ left,
                              int32_t* __restrict__ right, int32_t* __restrict__ parent,
                              int32_t* __restrict__ leafParent, int32_t* __restrict__ rangeFirst,
                              int32_t* __restrict__ rangeLast) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (lbvh_delta(sc, n, i, i + 1) - lbvh_delta(sc, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(sc, n, i, i - d);
    long long lmax = 2;
    while (true) {
        long long j = (long long)i + lmax * d;
        if (j < 0 || j >= n) break;
        if (!(lbvh_delta(sc, n, i, (int)j) > dmin)) break;
        lmax *= 2;
    }
    long long l = 0;
    for (long long t = lmax / 2; t >= 1; t /= 2) {
        long long j = (long long)i + (l + t) * d;
        if (j >= 0 && j < n && lbvh_delta(sc, n, i, (int)j) > dmin) l += t;
    }
    int j = i + (int)l * d;
    int dnode = lbvh_delta(sc, n, i, j);
    long long s = 0, t = l;
    do {
        t = (t + 1) / 2;
        long long q = (long long)i + (s + t) * d;
        if (q >= 0 && q < n && lbvh_delta(sc, n, i, (int)q) > dnode) s += t;
    } while (t > 1);
    int gamma = i + (int)s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    rangeFirst[i] = lo;
    rangeLast[i] = hi;
    int L = (lo == gamma) ? ~gamma : gamma;
    int R = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
    left[i] = L;
    right[i] = R;
    if (L >= 0) parent[L] = i; else leafParent[~L] = i;
    if (R >= 0) parent[R] = i; else leafParent[~R] = i;
    if (i == 0) parent[0] = -1;
}

__device__ __forceinline__ void child_bounds(int c, const float* __restrict__ nodeBounds, const float* __restrict__ primLo,
                                             const float* __restrict__ primHi, const uint32_t* __restrict__ sorted, float* b) {
    if (c >= 0) {
        for (int k = 0; k < 6; k++) b[k] = nodeBounds[6 * (size_t)c + k];
    } else {
        uint32_t p = sorted[~c];
        for (int k = 0; k < 3; k++) { b[k] = primLo[3 * (size_t)p + k]; b[3 + k] = primHi[3 * (size_t)p + k]; }
    }
}

__global__ void refit_kernel(int n, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                             const int32_t* __restrict__ parent, const int32_t* __restrict__ leafParent,
                             const float* __restrict__ primLo, const float* __restrict__ primHi,
                             const uint32_t* __restrict__ sorted, float* nodeBounds, uint32_t* flags) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int node = leafParent[k];
    while (node >= 0) {
        if (atomicAdd(&flags[node], 1u) == 0u) return;  // first arriver stops; the second sees both children
        __threadfence();
        float a[6], b[6];
        // children's bounds were written (st.cg) before their writer's fence + atomic: read them from L2
        const int ch[2] = {left[node], right[node]};
        for (int s = 0; s < 2; s++) {
            float* dst = s ? b : a;
            if (ch[s] >= 0) {
                for (int q = 0; q < 6; q++) dst[q] = __ldcg(&nodeBounds[6 * (size_t)ch[s] + q]);
            } else {
                uint32_t p = sorted[~ch[s]];
                for (int q = 0; q < 3; q++) { dst[q] = primLo[3 * (size_t)p + q]; dst[3 + q] = primHi[3 * (size_t)p + q]; }
            }
        }
        for (int q = 0; q < 3; q++) {
            __stcg(&nodeBounds[6 * (size_t)node + q], fminf(a[q], b[q]));
            __stcg(&nodeBounds[6 * (size_t)node + 3 + q], fmaxf(a[3 + q], b[3 + q]));
        }
        __threadfence();
        node = parent[node];
    }
}

__device__ __forceinline__ float half_area(const float* b) {
    float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
    return dx * dy + dy * dz + dz * dx;
}

// One level of the wide collapse.  Work item = (binary internal node, wide node slot).
template <int W>
__global__ void collapse_kernel(const uint32_t* __restrict__ queueIn, const uint32_t* __restrict__ counters, int level,
                                uint32_t* __restrict__ queueOut, uint32_t* countersOut,
                                const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                const int32_t* __restrict__ rangeFirst, const int32_t* __restrict__ rangeLast,
                                const float* __restrict__ nodeBounds, const float* __restrict__ primLo,
                                const float* __restrict__ primHi, const uint32_t* __restrict__ sorted, float pad, int leafMax,
                                WideNodeT<W>* __restrict__ nodes) {
    // counters: [0] = number of wide nodes allocated, [1 + (level&1)] = items in queueIn, [1 + (~level&1)] = out count
    const uint32_t numIn = counters[1 + (level & 1)];
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= numIn) return;
    const int bin = (int)queueIn[2 * (size_t)w];
    const uint32_t slot = queueIn[2 * (size_t)w + 1];
    int cand[W];
    int nc = 2;
    cand[0] = left[bin];
    cand[1] = right[bin];
    while (nc < W) {
        int best = -1;
        float bestArea = -1.f;
        for (int k = 0; k < nc; k++) {
            int c = cand[k];
            if (c < 0) continue;
            if (rangeLast[c] - rangeFirst[c] + 1 <= leafMax) continue;
            float a = half_area(nodeBounds + 6 * (size_t)c);
            if (a > bestArea) { bestArea = a; best = k; }
        }
        if (best < 0) break;
        int c = cand[best];
        cand[best] = left[c];
        cand[nc++] = right[c];
    }
    WideNodeT<W>& nd = nodes[slot];  // written in place (a 32-wide node does not fit in registers)
    for (int k = 0; k < W; k++) {
        if (k >= nc) {
            // empty slot: a huge FINITE box (no inf * 0 = NaN in the sign-masked slab test) that no ray ever enters
            nd.lox[k] = nd.loy[k] = nd.loz[k] = BVH_EMPTY_COORD;
            nd.hix[k] = nd.hiy[k] = nd.hiz[k] = BVH_EMPTY_COORD;
            nd.child[k] = BVH_EMPTY;
            continue;
        }
        int c = cand[k];
        float b[6];
        child_bounds(c, nodeBounds, primLo, primHi, sorted, b);
        nd.lox[k] = b[0] - pad; nd.loy[k] = b[1] - pad; nd.loz[k] = b[2] - pad;
        nd.hix[k] = b[3] + pad; nd.hiy[k] = b[4] + pad; nd.hiz[k] = b[5] + pad;
        if (c < 0) {
            nd.child[k] = bvh_make_leaf((uint32_t)(~c), 1u);
        } else {
            int cnt = rangeLast[c] - rangeFirst[c] + 1;
            if (cnt <= leafMax) {
                nd.child[k] = bvh_make_leaf((uint32_t)rangeFirst[c], (uint32_t)cnt);
            } else {
                uint32_t idx = atomicAdd(&countersOut[0], 1u);
                uint32_t q = atomicAdd(&countersOut[1 + ((level + 1) & 1)], 1u);
                queueOut[2 * (size_t)q] = (uint32_t)c;
                queueOut[2 * (size_t)q + 1] = idx;
                nd.child[k] = idx;
            }
        }
    }
}

// Warp-cooperative collapse for the 32-wide hierarchy: one warp per work item, lane k owns candidate k.  Every round the
// warp opens the expandable candidate with the largest box (arg-max by shuffles); the node is written with one child
// per lane (coalesced 128-byte rows).
__global__ void collapse32_kernel(const uint32_t* __restrict__ queueIn, const uint32_t* __restrict__ counters, int level,
                                  uint32_t* __restrict__ queueOut, uint32_t* countersOut,
                                  const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                  const int32_t* __restrict__ rangeFirst, const int32_t* __restrict__ rangeLast,
                                  const float* __restrict__ nodeBounds, const float* __restrict__ primLo,
                                  const float* __restrict__ primHi, const uint32_t* __restrict__ sorted, float pad, int leafMax,
                                  ShaftNode* __restrict__ nodes) {
    const unsigned full = 0xffffffffu;
    const uint32_t numIn = counters[1 + (level & 1)];
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= numIn) return;  // whole warp
    const int bin = (int)queueIn[2 * (size_t)w];
    const uint32_t slot = queueIn[2 * (size_t)w + 1];
    const int NONE = 0x7fffffff;
    int cand = lane == 0 ? left[bin] : lane == 1 ? right[bin] : NONE;
    int nc = 2;
    while (nc < SHAFT_WIDTH) {
        // key = half-area of expandable candidates, -1 otherwise
        float key = -1.f;
        if (cand != NONE && cand >= 0 && rangeLast[cand] - rangeFirst[cand] + 1 > leafMax) key = half_area(nodeBounds + 6 * (size_t)cand);
        float best = key;
        int bestLane = lane;
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(full, best, o);
            const int ol = __shfl_xor_sync(full, bestLane, o);
            if (ok > best || (ok == best && ol < bestLane)) { best = ok; bestLane = ol; }
        }
        if (best < 0.f) break;
        const int c = __shfl_sync(full, cand, bestLane);
        if (lane == bestLane) cand = left[c];
        if (lane == nc) cand = right[c];
        nc++;
    }
    ShaftNode& nd = nodes[slot];
    float b[6] = {BVH_EMPTY_COORD, BVH_EMPTY_COORD, BVH_EMPTY_COORD, BVH_EMPTY_COORD, BVH_EMPTY_COORD, BVH_EMPTY_COORD};
    uint32_t word = BVH_EMPTY;
    bool isNode = false;
    if (cand != NONE) {
        child_bounds(cand, nodeBounds, primLo, primHi, sorted, b);
        for (int k = 0; k < 3; k++) { b[k] -= pad; b[3 + k] += pad; }
        if (cand < 0) {
            word = bvh_make_leaf((uint32_t)(~cand), 1u);
        } else {
            const int cnt = rangeLast[cand] - rangeFirst[cand] + 1;
            if (cnt <= leafMax) word = bvh_make_leaf((uint32_t)rangeFirst[cand], (uint32_t)cnt);
            else isNode = true;
        }
    }
    // allocate node slots + queue entries for the child nodes with one atomic pair per warp
    const unsigned m = __ballot_sync(full, isNode);
    uint32_t baseIdx = 0, baseQ = 0;
    if (lane == 0 && m) {
        baseIdx = atomicAdd(&countersOut[0], (uint32_t)__popc(m));
        baseQ = atomicAdd(&countersOut[1 + ((level + 1) & 1)], (uint32_t)__popc(m));
    }
    baseIdx = __shfl_sync(full, baseIdx, 0);
    baseQ = __shfl_sync(full, baseQ, 0);
    if (isNode) {
        const uint32_t r = (uint32_t)__popc(m & ((1u << lane) - 1u));
        word = baseIdx + r;
        queueOut[2 * (size_t)(baseQ + r)] = (uint32_t)cand;
        queueOut[2 * (size_t)(baseQ + r) + 1] = word;
    }
    nd.lox[lane] = b[0]; nd.loy[lane] = b[1]; nd.loz[lane] = b[2];
    nd.hix[lane] = b[3]; nd.hiy[lane] = b[4]; nd.hiz[lane] = b[5];
    nd.child[lane] = word;
}

// Scene with <= BVH_LEAF_MAX triangles: one node, one leaf child.
template <int W>
__global__ void tiny_root_kernel(int n, const float* __restrict__ primLo, const float* __restrict__ primHi, float pad,
                                 WideNodeT<W>* nodes) {
    WideNodeT<W>& nd = nodes[0];
    float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) { b[k] = fminf(b[k], primLo[3 * i + k]); b[3 + k] = fmaxf(b[3 + k], primHi[3 * i + k]); }
    for (int k = 0; k < W; k++) {
        nd.lox[k] = nd.loy[k] = nd.loz[k] = BVH_EMPTY_COORD;
        nd.hix[k] = nd.hiy[k] = nd.hiz[k] = BVH_EMPTY_COORD;
        nd.child[k] = BVH_EMPTY;
    }
    nd.lox[0] = b[0] - pad; nd.loy[0] = b[1] - pad; nd.loz[0] = b[2] - pad;
    nd.hix[0] = b[3] + pad; nd.hiy[0] = b[4] + pad; nd.hiz[0] = b[5] + pad;
    nd.child[0] = bvh_make_leaf(0u, (uint32_t)n);
}


// Quantised twin of the 4-wide nodes (device_scene.h, CNode).  Every rounding goes outwards: the origin is the exact minimum of the
// child boxes, offsets are formed with directed subtractions, divided by a power-of-two scale (exact) and floored / ceiled; the
// scale is the smallest power of two that keeps every high plane within 255 units.
__global__ void compress_nodes_kernel(const WideNode* __restrict__ nodes, int numNodes, CNode* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) return;
    const WideNode nd = nodes[i];
    const float* lo[3] = {nd.lox, nd.loy, nd.loz};
    const float* hi[3] = {nd.hix, nd.hiy, nd.hiz};
    float org[3];
    uint32_t exps = 0u, ql[3] = {0u, 0u, 0u}, qh[3] = {0u, 0u, 0u};
    for (int a = 0; a < 3; a++) {
        float mn = INFINITY, mx = -INFINITY;
        for (int k = 0; k < BVH_WIDTH; k++)
            if (nd.child[k] != BVH_EMPTY) { mn = fminf(mn, lo[a][k]); mx = fmaxf(mx, hi[a][k]); }
        if (!(mn <= mx)) { mn = 0.f; mx = 0.f; }   // (a node without children does not occur)
        org[a] = mn;
        int e = 0;
        const float unit = __fdiv_ru(__fsub_ru(mx, mn), 255.0f);
        if (unit > 0.f) (void)frexpf(unit, &e);     // unit = m 2^e, m in [0.5, 1): 2^e >= unit
        int biased = e + 127;
        biased = biased < 1 ? 1 : (biased > 254 ? 254 : biased);
        for (;;) {
            const float s = __uint_as_float((uint32_t)biased << 23);
            bool ok = true;
            uint32_t wl = 0u, wh = 0u;
            for (int k = 0; k < BVH_WIDTH; k++) {
                uint32_t l = 255u, h = 0u;          // empty slot: inverted box (and the traversal checks the child word)
                if (nd.child[k] != BVH_EMPTY) {
                    const float fl = floorf(__fdiv_rd(__fsub_rd(lo[a][k], mn), s)), fh = ceilf(__fdiv_ru(__fsub_ru(hi[a][k], mn), s));
                    if (fh > 255.0f) ok = false;
                    l = (uint32_t)fmaxf(fl, 0.0f); h = (uint32_t)fminf(fmaxf(fh, 0.0f), 255.0f);
                }
                wl |= l << (8 * k); wh |= h << (8 * k);
            }
            if (ok || biased >= 254) { ql[a] = wl; qh[a] = wh; break; }
            biased++;
        }
        exps |= (uint32_t)biased << (8 * a);
    }
    CNode c;
    c.ox = org[0]; c.oy = org[1]; c.oz = org[2]; c.exps = exps;
    for (int k = 0; k < BVH_WIDTH; k++) c.child[k] = nd.child[k];
    c.qlx = ql[0]; c.qly = ql[1]; c.qlz = ql[2]; c.qhx = qh[0]; c.qhy = qh[1]; c.qhz = qh[2]; c.pad0 = 0u; c.pad1 = 0u;
    out[i] = c;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { if (err) *err = std::string(#x) + ": " + cudaGetErrorString(e_); return e_; } } while (0)


// Level-by-level collapse of the binary radix tree into W-ary nodes (surface-area guided: the child with
// the largest box is opened first), leaves of <= leafMax triangles.
template <int W>
static cudaError_t run_collapse(EvplpContext* c, WideNodeT<W>* nodes, int leafMax, int* numNodesOut, std::string* err) {
    cudaStream_t st = c->stream;
    uint32_t init[4] = {1u, 1u, 0u, 0u};  // one wide node (root), one item in queue A
    uint32_t rootItem[2] = {0u, 0u};
    CK(cudaMemcpyAsync(c->counters.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->queueA.p, rootItem, sizeof(rootItem), cudaMemcpyHostToDevice, st));
    // The item count of a level lives on the device; the grids are sized from an upper bound (W x the previous bound, at
    // most one item per binary internal node) and surplus threads exit, so the host only synchronises every 8 levels.
    const uint64_t cap = c->numPrims > 1 ? (uint64_t)c->numPrims - 1 : 1;
    uint64_t bound = 1;
    int level = 0;
    while (true) {
        uint32_t* qin = (level & 1) ? c->queueB.p : c->queueA.p;
        uint32_t* qout = (level & 1) ? c->queueA.p : c->queueB.p;
        CK(cudaMemsetAsync(c->counters.p + 1 + ((level + 1) & 1), 0, sizeof(uint32_t), st));
        if constexpr (W == SHAFT_WIDTH) {
            collapse32_kernel<<<(unsigned)((bound + 3) / 4), 128, 0, st>>>(qin, c->counters.p, level, qout, c->counters.p, c->left.p, c->right.p,
                                                                        c->rangeFirst.p, c->rangeLast.p, c->nodeBounds.p, c->primLo.p,
                                                                        c->primHi.p, c->primIdsSorted.p, c->boxPad, leafMax, nodes);
        } else {
            collapse_kernel<W><<<(unsigned)((bound + 63) / 64), 64, 0, st>>>(qin, c->counters.p, level, qout, c->counters.p, c->left.p, c->right.p,
                                                                          c->rangeFirst.p, c->rangeLast.p, c->nodeBounds.p, c->primLo.p,
                                                                          c->primHi.p, c->primIdsSorted.p, c->boxPad, leafMax, nodes);
        }
        c->launches++;
        bound = bound * W < cap ? bound * W : cap;
        level++;
        if (level % 8 == 0 || level > 4096) {
            uint32_t cnt[3];
            CK(cudaMemcpyAsync(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            *numNodesOut = (int)cnt[0];
            if (cnt[1 + (level & 1)] == 0) break;  // the queue the next level would read is empty
            if (level > 4096) { if (err) *err = "wide collapse did not terminate"; return cudaErrorUnknown; }
        }
    }
    return cudaGetLastError();
}

cudaError_t build_bvh_device(EvplpContext* c, std::string* err) {
    const int n = c->numPrims;
    cudaStream_t st = c->stream;
    c->numNodes = 0;
    c->numShaftNodes = 0;
    c->bvhBuilt = false;
    if (n == 0) { c->bvhBuilt = true; return cudaSuccess; }
    const int nInt = n > 1 ? n - 1 : 0;
    CK(c->primLo.reserve(3 * (size_t)n)); CK(c->primHi.reserve(3 * (size_t)n));
    CK(c->codes.reserve(n)); CK(c->codesSorted.reserve(n));
    CK(c->primIds.reserve(n)); CK(c->primIdsSorted.reserve(n));
    CK(c->triLeaf.reserve(4 * (size_t)n));
    CK(c->left.reserve(nInt)); CK(c->right.reserve(nInt)); CK(c->parent.reserve(nInt));
    CK(c->leafParent.reserve(n)); CK(c->rangeFirst.reserve(nInt)); CK(c->rangeLast.reserve(nInt));
    CK(c->nodeBounds.reserve(6 * (size_t)nInt)); CK(c->refitFlags.reserve(nInt));
    CK(c->nodes.reserve(nInt > 0 ? nInt : 1));
    // 32-wide nodes: every bottom-level node covers > leafMax triangles of its own and every node with node children is
    // full (32 children), so there are at most ~n / (leafMax + 1) * 32/31 of them
    CK(c->shaftNodes.reserve(nInt > 0 ? (size_t)nInt / (size_t)(c->opt.shaftLeafMax + 1) + (size_t)nInt / 16 + 64 : 1));
    CK(c->sceneBoundsEnc.reserve(6));
    CK(c->queueA.reserve(2 * (size_t)(nInt + 1))); CK(c->queueB.reserve(2 * (size_t)(nInt + 1)));
    CK(c->counters.reserve(4));

    size_t tempBytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, c->codes.p, c->codesSorted.p, c->primIds.p, c->primIdsSorted.p, n, 0, 63, st));
    CK(c->sortTemp.reserve(tempBytes));
    c->stageBegin(ST_BVH);  // device time of the build proper (allocations above are setup)

    const int TB = 256;
    const int gridN = (n + TB - 1) / TB;
    // 1 bounds
    uint32_t encInit[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CK(cudaMemcpyAsync(c->sceneBoundsEnc.p, encInit, sizeof(encInit), cudaMemcpyHostToDevice, st));
    prim_bounds_kernel<<<gridN, TB, 0, st>>>(c->triVerts.p, n, c->primLo.p, c->primHi.p, c->sceneBoundsEnc.p);
    c->launches++;
    uint32_t enc[6];
    CK(cudaMemcpyAsync(enc, c->sceneBoundsEnc.p, sizeof(enc), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    Bounds3 sb;
    float m = 0.f;
    for (int k = 0; k < 3; k++) {
        sb.lo[k] = dec_ordered(enc[k]); sb.hi[k] = dec_ordered(enc[3 + k]);
        c->sceneMin[k] = sb.lo[k]; c->sceneMax[k] = sb.hi[k];
        m = fmaxf(m, fmaxf(fabsf(sb.lo[k]), fabsf(sb.hi[k])));
    }
    c->boxPad = m * 1e-5f + 1e-20f;
    // 2 codes
    morton_kernel<<<gridN, TB, 0, st>>>(c->primLo.p, c->primHi.p, n, sb, c->codes.p, c->primIds.p);
    c->launches++;
    // 3 sort (stable LSD radix sort over the 63 used bits)
    CK(cub::DeviceRadixSort::SortPairs(c->sortTemp.p, tempBytes, c->codes.p, c->codesSorted.p, c->primIds.p, c->primIdsSorted.p, n, 0, 63, st));
    c->launches += 9;  // histogram + onesweep passes (CUB-internal; counted approximately)
    // 4 leaf records
    leaf_records_kernel<<<gridN, TB, 0, st>>>(c->triVerts.p, c->primIdsSorted.p, n, c->triLeaf.p);
    c->launches++;
    if (n <= c->opt.bvhLeafMax) {
        tiny_root_kernel<BVH_WIDTH><<<1, 1, 0, st>>>(n, c->primLo.p, c->primHi.p, c->boxPad, c->nodes.p);
        tiny_root_kernel<SHAFT_WIDTH><<<1, 1, 0, st>>>(n, c->primLo.p, c->primHi.p, c->boxPad, c->shaftNodes.p);
        CK(c->cnodes.reserve(1));
        compress_nodes_kernel<<<1, 32, 0, st>>>(c->nodes.p, 1, c->cnodes.p);
        c->launches += 2;
        CK(cudaStreamSynchronize(st));
        c->numNodes = 1;
        c->numShaftNodes = 1;
        c->bvhBuilt = true;
        return cudaSuccess;
    }
    // 5 topology
    const int gridI = (nInt + TB - 1) / TB;
    karras_kernel<<<gridI, TB, 0, st>>>(c->codesSorted.p, n, c->left.p, c->right.p, c->parent.p, c->leafParent.p,
                                        c->rangeFirst.p, c->rangeLast.p);
    c->launches++;
    // 6 refit
    CK(cudaMemsetAsync(c->refitFlags.p, 0, sizeof(uint32_t) * nInt, st));
    refit_kernel<<<gridN, TB, 0, st>>>(n, c->left.p, c->right.p, c->parent.p, c->leafParent.p, c->primLo.p, c->primHi.p,
                                       c->primIdsSorted.p, c->nodeBounds.p, c->refitFlags.p);
    c->launches++;
    // 7 collapse into the 4-wide per-ray hierarchy and the 32-wide shaft hierarchy
    {
        cudaError_t e = run_collapse<BVH_WIDTH>(c, c->nodes.p, c->opt.bvhLeafMax, &c->numNodes, err);
        if (e != cudaSuccess) return e;
        e = run_collapse<SHAFT_WIDTH>(c, c->shaftNodes.p, c->opt.shaftLeafMax, &c->numShaftNodes, err);
        if (e != cudaSuccess) return e;
        CK(c->cnodes.reserve(c->numNodes > 0 ? (size_t)c->numNodes : 1));
        if (c->numNodes > 0) {
            compress_nodes_kernel<<<(c->numNodes + 127) / 128, 128, 0, st>>>(c->nodes.p, c->numNodes, c->cnodes.p);
            c->launches++;
        }
    }
    CK(cudaGetLastError());
    c->bvhBuilt = true;
    return cudaSuccess;
}

}  // namespace evplp
