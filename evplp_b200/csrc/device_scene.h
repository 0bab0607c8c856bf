// device_scene.h -- HBM layout of the scene and the wide BVH, plus the traversal routines
// that replace rtTrace / meshFineIntersect / rtMaterialAnyHit
// (reference: triangleintersect.cu:17-41, lighttracing.cu:184-188,236,292; the OptiX
// "Trbvh" acceleration set up at rtcomphoton.h:705-707).
//
// Layout
//   triLeaf[4*k .. 4*k+3]  one triangle in BVH-leaf (Morton) order, 64 B:
//        {p0.xyz, primId} {e0.xyz, matIndex} {e1.xyz, -} {n.xyz, -}
//        e0 = p1-p0, e1 = p0-p2, n = cross(e1,e0): the ray-independent part of
//        optix::intersect_triangle_branchless, precomputed with the same roundings.
//   triVerts[3*p .. 3*p+2] original primitive order {p0.xyz, mat} {p1.xyz,-} {p2.xyz,-}
//   triUV[3*p .. 3*p+2]    original order texcoords t0,t1,t2
//   nodes[]                BVH_WIDTH-ary nodes, SoA child boxes (padded), 128 B aligned
//
// Hit semantics (identical to the oracle): a triangle is hit iff the branchless test
// accepts it; closest hit = smallest t, ties -> smallest primitive id; any hit = exists.
// Box tests are only a conservative cull: boxes are padded by 1e-5 * scene scale.
#pragma once
#include <cuda_runtime.h>
#include "shading.h"

namespace evplp {

constexpr int BVH_WIDTH = 4;
constexpr int BVH_LEAF_MAX = 2;             // default triangles per leaf child (tunable: bvh_leaf_max)
constexpr uint32_t BVH_EMPTY = 0x7fffffffu;
constexpr uint32_t BVH_LEAF_BIT = 0x80000000u;
constexpr int BVH_STACK = 96;
constexpr float BVH_EMPTY_COORD = 3.0e38f;   // box of an empty child slot (finite, beyond any scene)

template <int W>
struct alignas(128) WideNodeT {
    float lox[W], loy[W], loz[W];
    float hix[W], hiy[W], hiz[W];
    uint32_t child[W];  // EMPTY | LEAF_BIT | (count-1) << 27 | first   or   node index
};
using WideNode = WideNodeT<BVH_WIDTH>;   // 128 B: per-ray traversals (closest hit, per-lane / packet any hit)
constexpr int SHAFT_WIDTH = 32;          // one child per lane: warp-cooperative shaft traversal of the gather
using ShaftNode = WideNodeT<SHAFT_WIDTH>;  // 896 B, every plane array is one coalesced 128-byte row

// Quantised copy of a 4-wide node for the closest-hit traversal (one ray per thread: light paths, primary rays).  Incoherent rays
// fetch a different node per lane, and ncu shows that traversal bound by L1 requests (7 x 16 B per lane and node); this form is
// 4 x 16 B.  Child boxes are stored as 8-bit offsets from the node's own box origin in units of a per-axis power-of-two scale,
// rounded OUTWARDS at build time (bvh.cu, compress_nodes_kernel), so every decoded box contains the float box it came from:
// still only a conservative cull in front of the exact triangle test.
struct alignas(64) CNode {
    float ox, oy, oz;           // origin: the low corner of the union of the child boxes
    uint32_t exps;              // bytes 0..2: IEEE biased exponents of the x / y / z scales (scale = 2^(e - 127))
    uint32_t child[4];          // as in WideNode
    uint32_t qlx, qly, qlz, qhx;  // byte k = child k: low / high planes in scale units from the origin
    uint32_t qhy, qhz, pad0, pad1;
};

EVPLP_HD uint32_t bvh_make_leaf(uint32_t first, uint32_t count) { return BVH_LEAF_BIT | ((count - 1u) << 27) | first; }
EVPLP_HD uint32_t bvh_leaf_first(uint32_t c) { return c & 0x07ffffffu; }
EVPLP_HD uint32_t bvh_leaf_count(uint32_t c) { return ((c >> 27) & 0xfu) + 1u; }

struct DevScene {
    const float4* triLeaf;
    const float4* triVerts;
    const float2* triUV;
    const DevMaterial* mats;
    const float4* texPool;
    const float* lightCdf;
    const WideNode* nodes;
    const CNode* cnodes;          // quantised twin of `nodes` (same indices): closest-hit traversal
    const ShaftNode* shaftNodes;  // 32-wide hierarchy over the same triangles (same leaf order)
    int numShaftNodes;
    int numPrims;
    int numNodes;
    int lightFirst, lightCount;
    float lightArea;
    float lightIntensity[4];  // pi-scaled rgb, .w = emission exponent
    float lightDisplay[4];
};

struct RayHit {
    int prim;
    float t, beta, gamma;
    V3 n;  // un-normalised cross(e1, e0)
    int mat;
};

#if defined(__CUDACC__)

// Per-ray constants of the slab test: t = lo * inv - org*inv (one FMA per plane).
struct RaySlab {
    float ix, iy, iz, ox, oy, oz;
};
__device__ __forceinline__ RaySlab make_slab(V3 org, V3 dir) {
    RaySlab s;
    // a zero component would give inf * 0 = NaN products; a tiny one keeps the test finite
    float dx = fabsf(dir.x) < 1e-30f ? copysignf(1e-30f, dir.x) : dir.x;
    float dy = fabsf(dir.y) < 1e-30f ? copysignf(1e-30f, dir.y) : dir.y;
    float dz = fabsf(dir.z) < 1e-30f ? copysignf(1e-30f, dir.z) : dir.z;
    s.ix = __frcp_rn(dx); s.iy = __frcp_rn(dy); s.iz = __frcp_rn(dz);
    s.ox = org.x * s.ix; s.oy = org.y * s.iy; s.oz = org.z * s.iz;
    return s;
}

// entry distance of child c (clipped to [tmin, tmax]); returns whether the slab interval is non-empty
__device__ __forceinline__ bool slab_test(const RaySlab& s, const WideNode& nd, int c, float tmin, float tmax, float* tnear) {
    float x0 = __fmaf_rn(nd.lox[c], s.ix, -s.ox), x1 = __fmaf_rn(nd.hix[c], s.ix, -s.ox);
    float y0 = __fmaf_rn(nd.loy[c], s.iy, -s.oy), y1 = __fmaf_rn(nd.hiy[c], s.iy, -s.oy);
    float z0 = __fmaf_rn(nd.loz[c], s.iz, -s.oz), z1 = __fmaf_rn(nd.hiz[c], s.iz, -s.oz);
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    *tnear = tn;
    return tn <= tf;
}

// optix::intersect_triangle_branchless with the ray-independent terms precomputed.  Every operation is an explicit
// round-to-nearest intrinsic in the association order of vec.h, so the hit decision has the oracle's bits in ANY
// translation unit -- also in gather_fast.cu, which is compiled with FMA contraction.
__device__ __forceinline__ float dot_rn(V3 a, V3 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ V3 cross_rn(V3 a, V3 b) {
    return v3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
              __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ bool tri_test(V3 org, V3 dir, float tmin, float tmax, V3 p0, V3 e0, V3 e1, V3 n,
                                         float* t, float* beta, float* gamma) {
    const float inv = __frcp_rn(dot_rn(n, dir));   // = 1.0f / x correctly rounded (what the oracle's division gives), without the general divide
    const V3 e2 = v3(__fmul_rn(inv, __fsub_rn(p0.x, org.x)), __fmul_rn(inv, __fsub_rn(p0.y, org.y)), __fmul_rn(inv, __fsub_rn(p0.z, org.z)));
    const V3 i = cross_rn(dir, e2);
    *beta = dot_rn(i, e1);
    *gamma = dot_rn(i, e0);
    *t = dot_rn(n, e2);
    return (*t < tmax) & (*t > tmin) & (*beta >= 0.0f) & (*gamma >= 0.0f) & (__fadd_rn(*beta, *gamma) <= 1.0f);
}

__device__ __forceinline__ V3 ld3(const float4& v) { return v3(v.x, v.y, v.z); }

// Any hit for a whole warp at once: the 32 rays walk the tree together with ONE shared
// stack, so every node / triangle is fetched once per warp (uniform address -> broadcast)
// and there is no divergence.  Rays of the gather are coherent (neighbouring pixels, same
// VPL), so the union of their paths is barely larger than one ray's.  `active` lanes carry
// a ray; returns per lane whether it is occluded.  Must be called by all 32 lanes.
//
// The loop body is branch-free per lane: all four child slabs are tested with straight-line
// code, the per-lane 4-bit result is OR-reduced over the warp with one REDUX, and every lane
// performs the same (uniform) pushes, writing identical values to the warp's stack -- so no
// __syncwarp is needed (a lane only ever reads back what it wrote itself).
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ RaySlab make_slab_fast(V3 org, V3 dir) {
    RaySlab s;
    float dx = fabsf(dir.x) < 1e-30f ? copysignf(1e-30f, dir.x) : dir.x;
    float dy = fabsf(dir.y) < 1e-30f ? copysignf(1e-30f, dir.y) : dir.y;
    float dz = fabsf(dir.z) < 1e-30f ? copysignf(1e-30f, dir.z) : dir.z;
    // 1-ulp reciprocal is enough: boxes are padded by ~100x the slab test's rounding error
    s.ix = rcp_approx(dx); s.iy = rcp_approx(dy); s.iz = rcp_approx(dz);
    s.ox = org.x * s.ix; s.oy = org.y * s.iy; s.oz = org.z * s.iz;
    return s;
}

__device__ __forceinline__ bool slab4(const RaySlab& s, float lx, float ly, float lz, float hx, float hy, float hz,
                                      float tmin, float tmax) {
    const float x0 = __fmaf_rn(lx, s.ix, -s.ox), x1 = __fmaf_rn(hx, s.ix, -s.ox);
    const float y0 = __fmaf_rn(ly, s.iy, -s.oy), y1 = __fmaf_rn(hy, s.iy, -s.oy);
    const float z0 = __fmaf_rn(lz, s.iz, -s.oz), z1 = __fmaf_rn(hz, s.iz, -s.oz);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    return tn <= tf;
}

// ---- warp-cooperative any-hit traversal -------------------------------------------------------
// Shared-memory stack addressed by a 32-bit shared-window byte address kept in a register.
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Per-ray constants of the sign-masked slab test.  For each axis (a, b) = (inv, 0) when the
// direction component is >= 0 and (0, inv) otherwise, so that
//     t_entry = lo * a + hi * b - org * inv        t_exit = lo * b + hi * a - org * inv
// select the entry / exit plane with FMAs only (the FMA pipe is idle next to the ALU pipe that
// executes min/max), per lane, with no sortedness requirement and no address arithmetic.
struct RaySlabM {
    float ax, ay, az, bx, by, bz, ox, oy, oz;
};
__device__ __forceinline__ RaySlabM make_slab_masked(V3 org, V3 dir) {
    const RaySlab s = make_slab_fast(org, dir);
    RaySlabM m;
    m.ax = s.ix >= 0.f ? s.ix : 0.f; m.bx = s.ix >= 0.f ? 0.f : s.ix;
    m.ay = s.iy >= 0.f ? s.iy : 0.f; m.by = s.iy >= 0.f ? 0.f : s.iy;
    m.az = s.iz >= 0.f ? s.iz : 0.f; m.bz = s.iz >= 0.f ? 0.f : s.iz;
    m.ox = -s.ox; m.oy = -s.oy; m.oz = -s.oz;
    return m;
}
__device__ __forceinline__ bool slab_masked(const RaySlabM& s, float lx, float ly, float lz, float hx, float hy, float hz,
                                            float tmin, float tmax) {
    const float nx = __fmaf_rn(lx, s.ax, __fmaf_rn(hx, s.bx, s.ox)), fx = __fmaf_rn(lx, s.bx, __fmaf_rn(hx, s.ax, s.ox));
    const float ny = __fmaf_rn(ly, s.ay, __fmaf_rn(hy, s.by, s.oy)), fy = __fmaf_rn(ly, s.by, __fmaf_rn(hy, s.ay, s.oy));
    const float nz = __fmaf_rn(lz, s.az, __fmaf_rn(hz, s.bz, s.oz)), fz = __fmaf_rn(lz, s.bz, __fmaf_rn(hz, s.az, s.oz));
    const float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float tf = fminf(fminf(fx, fy), fminf(fz, tmax));
    return tn <= tf;
}

// entry distance of a child for the sign-masked slab test (+inf when the ray misses the box)
__device__ __forceinline__ float slab_entry_masked(const RaySlabM& s, float lx, float ly, float lz, float hx, float hy, float hz,
                                                   float tmin, float tmax) {
    const float nx = __fmaf_rn(lx, s.ax, __fmaf_rn(hx, s.bx, s.ox)), fx = __fmaf_rn(lx, s.bx, __fmaf_rn(hx, s.ax, s.ox));
    const float ny = __fmaf_rn(ly, s.ay, __fmaf_rn(hy, s.by, s.oy)), fy = __fmaf_rn(ly, s.by, __fmaf_rn(hy, s.ay, s.oy));
    const float nz = __fmaf_rn(lz, s.az, __fmaf_rn(hz, s.bz, s.oz)), fz = __fmaf_rn(lz, s.bz, __fmaf_rn(hz, s.az, s.oz));
    const float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float tf = fminf(fminf(fx, fy), fminf(fz, tmax));
    return tn <= tf ? tn : INFINITY;
}

// order (ta, ca), (tb, cb) so that ta >= tb
__device__ __forceinline__ void cswap_desc(float& ta, uint32_t& ca, float& tb, uint32_t& cb) {
    const bool sw = ta < tb;
    const float hi = fmaxf(ta, tb), lo = fminf(ta, tb);
    const uint32_t c_hi = sw ? cb : ca, c_lo = sw ? ca : cb;
    ta = hi; tb = lo; ca = c_hi; cb = c_lo;
}

// byte k of w as a float, exactly: 0x4B0000qq is 2^23 + qq
__device__ __forceinline__ float byte_as_float(uint32_t w, int k) {
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | (uint32_t)k)), 8388608.0f);
}
// entry distance of child k of a quantised node (+inf: missed or empty), sign-masked like slab_entry_masked
__device__ __forceinline__ float cnode_entry(int k, uint32_t child, uint4 qa, uint4 qb, float sax, float sbx, float say, float sby,
                                             float saz, float sbz, float bx0, float by0, float bz0, float tmin, float tmax) {
    const float lx = byte_as_float(qa.x, k), ly = byte_as_float(qa.y, k), lz = byte_as_float(qa.z, k);
    const float hx = byte_as_float(qa.w, k), hy = byte_as_float(qb.x, k), hz = byte_as_float(qb.y, k);
    const float nx = __fmaf_rn(lx, sax, __fmaf_rn(hx, sbx, bx0)), fx = __fmaf_rn(lx, sbx, __fmaf_rn(hx, sax, bx0));
    const float ny = __fmaf_rn(ly, say, __fmaf_rn(hy, sby, by0)), fy = __fmaf_rn(ly, sby, __fmaf_rn(hy, say, by0));
    const float nz = __fmaf_rn(lz, saz, __fmaf_rn(hz, sbz, bz0)), fz = __fmaf_rn(lz, sbz, __fmaf_rn(hz, saz, bz0));
    const float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, tmin));
    const float tf = fminf(fminf(fx, fy), fminf(fz, tmax));
    return (tn <= tf && child != BVH_EMPTY) ? tn : INFINITY;
}

// Closest hit, one ray per thread, private stack.  Children are visited front to back (sorting
// network on the entry distances); popped entries that start beyond the current best hit are skipped.
// QUANT: fetch the quantised nodes (4 x 16 B per visit instead of 7; incoherent rays: light paths, path-tracer bounces: -13..16 %)
// or the float nodes (coherent primary rays, whose node fetches coalesce anyway and which would only pay for the decode: +50 %).
template <bool QUANT>
__device__ inline RayHit trace_closest(const DevScene& sc, V3 org, V3 dir, float tmin, float tmax, int* overflow) {
    RayHit best;
    best.prim = -1; best.t = 0.f; best.beta = 0.f; best.gamma = 0.f; best.n = v3s(0.f); best.mat = 0;
    if (sc.numNodes == 0) return best;
    float bestT = tmax;
    const RaySlabM slab = make_slab_masked(org, dir);
    uint32_t stack[BVH_STACK];
    float stackT[BVH_STACK];
    int sp = 0;
    uint32_t cur = 0;  // root node
    // "while-while" order: a lane walks inner nodes until it holds a leaf (or runs out of work), and only then do the lanes of the
    // warp test triangles together.  With an if / else per step, a warp whose lanes sit in different states pays for a node visit
    // AND a leaf visit in every step; light paths are incoherent, so that is the common case.  The set of tests a ray performs is
    // still a superset of what decides its closest hit (entries beyond the best hit are skipped), and every test is exact, so the
    // result does not depend on the order.
    for (;;) {
        while (cur != BVH_EMPTY && !(cur & BVH_LEAF_BIT)) {
            float t0, t1, t2, t3;
            uint4 ch;
            if (QUANT) {
                const uint4* np = reinterpret_cast<const uint4*>(sc.cnodes + cur);
                const uint4 hd = __ldg(np), qa = __ldg(np + 2), qb = __ldg(np + 3);
                ch = __ldg(np + 1);
                // the node's frame in ray space: plane p at q scale units from the origin has t = q * (scale * inv) + (origin - org) * inv
                const float sx = __uint_as_float((hd.w & 0xffu) << 23), sy = __uint_as_float(((hd.w >> 8) & 0xffu) << 23),
                            sz = __uint_as_float(((hd.w >> 16) & 0xffu) << 23);
                const float sax = sx * slab.ax, sbx = sx * slab.bx, say = sy * slab.ay, sby = sy * slab.by, saz = sz * slab.az, sbz = sz * slab.bz;
                const float bx0 = __fmaf_rn(__uint_as_float(hd.x), slab.ax + slab.bx, slab.ox), by0 = __fmaf_rn(__uint_as_float(hd.y), slab.ay + slab.by, slab.oy),
                            bz0 = __fmaf_rn(__uint_as_float(hd.z), slab.az + slab.bz, slab.oz);
                t0 = cnode_entry(0, ch.x, qa, qb, sax, sbx, say, sby, saz, sbz, bx0, by0, bz0, tmin, bestT);
                t1 = cnode_entry(1, ch.y, qa, qb, sax, sbx, say, sby, saz, sbz, bx0, by0, bz0, tmin, bestT);
                t2 = cnode_entry(2, ch.z, qa, qb, sax, sbx, say, sby, saz, sbz, bx0, by0, bz0, tmin, bestT);
                t3 = cnode_entry(3, ch.w, qa, qb, sax, sbx, say, sby, saz, sbz, bx0, by0, bz0, tmin, bestT);
            } else {
                const float4* np = reinterpret_cast<const float4*>(sc.nodes + cur);
                const float4 lox = __ldg(np), loy = __ldg(np + 1), loz = __ldg(np + 2);
                const float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
                ch = __ldg(reinterpret_cast<const uint4*>(np + 6));
                t0 = slab_entry_masked(slab, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, tmin, bestT);
                t1 = slab_entry_masked(slab, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, tmin, bestT);
                t2 = slab_entry_masked(slab, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, tmin, bestT);
                t3 = slab_entry_masked(slab, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, tmin, bestT);
            }
            uint32_t c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
            cswap_desc(t0, c0, t1, c1); cswap_desc(t2, c2, t3, c3); cswap_desc(t0, c0, t2, c2);
            cswap_desc(t1, c1, t3, c3); cswap_desc(t1, c1, t2, c2);
            // (t0 >= t1 >= t2 >= t3; misses are +inf and sort to the front)
            if (sp + 3 > BVH_STACK) { *overflow = 1; return best; }
            if (t0 < INFINITY) { stack[sp] = c0; stackT[sp] = t0; sp++; }
            if (t1 < INFINITY) { stack[sp] = c1; stackT[sp] = t1; sp++; }
            if (t2 < INFINITY) { stack[sp] = c2; stackT[sp] = t2; sp++; }
            cur = t3 < INFINITY ? c3 : BVH_EMPTY;  // nearest child continues without a stack round trip
            while (cur == BVH_EMPTY && sp > 0) {
                --sp;
                if (stackT[sp] <= bestT) cur = stack[sp];  // entries that begin beyond the best hit cannot improve it (ties kept)
            }
        }
        if (cur == BVH_EMPTY) return best;
        {
            const uint32_t first = bvh_leaf_first(cur), count = bvh_leaf_count(cur);
            for (uint32_t k = 0; k < count; k++) {
                const float4* tp = sc.triLeaf + 4 * (size_t)(first + k);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2), d = __ldg(tp + 3);
                float t, be, ga;
                if (tri_test(org, dir, tmin, tmax, ld3(a), ld3(b), ld3(c), ld3(d), &t, &be, &ga)) {
                    const int prim = __float_as_int(a.w);
                    if (best.prim < 0 ? (t < bestT) : (t < bestT || (t == bestT && prim < best.prim))) {
                        best.prim = prim; best.t = t; best.beta = be; best.gamma = ga; best.n = ld3(d);
                        best.mat = __float_as_int(b.w);
                        bestT = t;
                    }
                }
            }
            cur = BVH_EMPTY;
            while (cur == BVH_EMPTY && sp > 0) {
                --sp;
                if (stackT[sp] <= bestT) cur = stack[sp];
            }
        }
    }
    return best;
}

// Any hit, one ray per thread, private stack.
__device__ inline bool trace_any(const DevScene& sc, V3 org, V3 dir, float tmin, float tmax, int* overflow) {
    if (sc.numNodes == 0) return false;
    const RaySlab slab = make_slab(org, dir);
    uint32_t stack[BVH_STACK];
    int sp = 0;
    uint32_t cur = 0;
    while (true) {
        if (cur & BVH_LEAF_BIT) {
            const uint32_t first = bvh_leaf_first(cur), count = bvh_leaf_count(cur);
            for (uint32_t k = 0; k < count; k++) {
                const float4* tp = sc.triLeaf + 4 * (size_t)(first + k);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2), d = __ldg(tp + 3);
                float t, be, ga;
                if (tri_test(org, dir, tmin, tmax, ld3(a), ld3(b), ld3(c), ld3(d), &t, &be, &ga)) return true;
            }
        } else {
            const WideNode& nd = sc.nodes[cur];
#pragma unroll
            for (int c = 0; c < BVH_WIDTH; c++) {
                float t;
                const uint32_t cd = nd.child[c];
                if (cd != BVH_EMPTY && slab_test(slab, nd, c, tmin, tmax, &t)) {
                    if (sp < BVH_STACK) stack[sp++] = cd; else *overflow = 1;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return false;
}

static_assert(BVH_WIDTH == 4, "trace_any_warp is written for 4-wide nodes");

static_assert(BVH_WIDTH == 4, "the warp traversal is written for 4-wide nodes");

// Any hit for a whole warp at once: the 32 rays walk the tree together with ONE shared
// stack, so every node / triangle is fetched once per warp (uniform address -> broadcast)
// and there is no divergence.  Rays of the gather are coherent (neighbouring pixels, same
// VPL), so the union of their paths is barely larger than one ray's.  `active` lanes carry
// a ray; returns per lane whether it is occluded.  Must be called by all 32 lanes.
//
// The loop body is branch-free per lane; the four per-child results are OR-reduced with votes
// and every lane performs the same (uniform) pushes, writing identical values to the warp's
// stack -- so no __syncwarp is needed (a lane only reads back slots it wrote itself; the votes
// keep the lanes within one iteration of each other).  Empty child slots hold huge finite
// boxes that are never entered, so the child words need no test.
__device__ inline bool trace_any_warp(const DevScene& sc, bool active, V3 org, V3 dir, float tmin, float tmax,
                                      uint32_t* warpStack /* BVH_STACK entries in smem */, int* overflow) {
    const unsigned full = 0xffffffffu;
    bool open = active;  // still needs an answer
    if (sc.numNodes == 0 || !__any_sync(full, active)) return false;
    const RaySlabM slab = make_slab_masked(org, dir);
    const uint32_t stackBase = (uint32_t)__cvta_generic_to_shared(warpStack);
    uint32_t sp = stackBase;
    uint32_t cur = 0;
    while (true) {
        if (cur & BVH_LEAF_BIT) {
            const uint32_t first = bvh_leaf_first(cur), count = bvh_leaf_count(cur);
            const float4* tp = sc.triLeaf + 4 * (size_t)first;
            bool hit = false;
            for (uint32_t k = 0; k < count; k++, tp += 4) {
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2), d = __ldg(tp + 3);
                float t, be, ga;
                hit |= tri_test(org, dir, tmin, tmax, ld3(a), ld3(b), ld3(c), ld3(d), &t, &be, &ga);
            }
            open = open && !hit;
            if (!__any_sync(full, open)) break;
            cur = BVH_EMPTY;
        } else {
            const float4* np = reinterpret_cast<const float4*>(sc.nodes + cur);
            const float4 lox = __ldg(np), loy = __ldg(np + 1), loz = __ldg(np + 2);
            const float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
            const uint4 ch = __ldg(reinterpret_cast<const uint4*>(np + 6));
            const bool h0 = slab_masked(slab, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, tmin, tmax);
            const bool h1 = slab_masked(slab, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, tmin, tmax);
            const bool h2 = slab_masked(slab, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, tmin, tmax);
            const bool h3 = slab_masked(slab, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, tmin, tmax);
            // warp-uniform from here on: every lane performs the same pushes (identical values);
            // the last entered child continues in a register (no stack round trip on the critical path)
            const bool p0 = __any_sync(full, h0 & open), p1 = __any_sync(full, h1 & open);
            const bool p2 = __any_sync(full, h2 & open), p3 = __any_sync(full, h3 & open);
            if (sp + 12u > stackBase + 4u * BVH_STACK) { *overflow = 1; break; }
            cur = BVH_EMPTY;
            if (p0) cur = ch.x;
            if (p1) { if (cur != BVH_EMPTY) { st_shared_u32(sp, cur); sp += 4u; } cur = ch.y; }
            if (p2) { if (cur != BVH_EMPTY) { st_shared_u32(sp, cur); sp += 4u; } cur = ch.z; }
            if (p3) { if (cur != BVH_EMPTY) { st_shared_u32(sp, cur); sp += 4u; } cur = ch.w; }
        }
        if (cur == BVH_EMPTY) {
            if (sp == stackBase) break;
            sp -= 4u;
            cur = ld_shared_u32(sp);
        }
    }
    return active && !open;
}

// Values the compiler must keep in a register instead of re-deriving them inside the traversal loop
// (it otherwise rematerialises selects / shared-window addresses on every node visit).
__device__ __forceinline__ int opaque(int x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ uint32_t opaque(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ float opaque(float x) { asm volatile("" : "+f"(x)); return x; }

// ---- shaft traversal of the 32-wide hierarchy ---------------------------------------------------
// All 32 shadow rays of a warp start at the same point (the VPL) and end inside the warp's pixel
// tile, so they lie in the shaft  S(t) = vpl + t * ([tileLo, tileHi] - vpl),  t in [tmin, tmax]
// (same parameter t as the rays).  Instead of 32 lanes x 4 child boxes of per-ray slab tests, ONE
// conservative shaft-vs-box test per child decides where the warp descends, and the 32 lanes test 32
// children of a 32-wide node at once (one coalesced 128-byte row per plane array).  The descent only
// COLLECTS candidate leaves; every lane then runs the exact triangle test of its own ray against the
// candidates' triangles, so the result per ray is the same "exists a triangle that the branchless test
// accepts" as everywhere else -- the shaft is just another conservative cull.
constexpr int SHAFT_CAND = 128;  // capacity of the candidate-leaf list; the launch-time limit (shaft_max_candidates) is <= this

struct Shaft {
    float ilx, ily, ilz, ihx, ihy, ihz;  // 1 / (tileLo - vpl), 1 / (tileHi - vpl) per axis
    V3 apex;
    unsigned signs;                      // bit a: tileLo_a - vpl_a > 0, bit 3 + a: tileHi_a - vpl_a > 0
    // fast path (no axis on which the tile straddles the apex): entry / exit plane of every axis selected by ADDRESS
    int sorted;                          // 1 = fast path usable
    int nOffX, nOffY, nOffZ, fOffX, fOffY, fOffZ;  // float offsets of the entry / exit plane rows inside a ShaftNode
    float nix, niy, niz, fix, fiy, fiz;  // matching reciprocals
};

__device__ __forceinline__ Shaft make_shaft(V3 apex, V3 tileLo, V3 tileHi) {
    Shaft s;
    s.apex = apex;
    float dl[3] = {tileLo.x - apex.x, tileLo.y - apex.y, tileLo.z - apex.z};
    float dh[3] = {tileHi.x - apex.x, tileHi.y - apex.y, tileHi.z - apex.z};
    unsigned sg = 0;
    float il[3], ih[3];
    int nOff[3], fOff[3];
    float ni[3], fi[3];
    int sorted = 1;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        // a zero extent keeps its constraint ("t * 0 <= A" / "t * 0 >= B") as a huge bound of the right sign
        if (fabsf(dl[a]) < 1e-30f) dl[a] = -1e-30f;
        if (fabsf(dh[a]) < 1e-30f) dh[a] = 1e-30f;
        const bool lp = dl[a] > 0.f, hp = dh[a] > 0.f;
        if (lp) sg |= 1u << a;
        if (hp) sg |= 8u << a;
        il[a] = rcp_approx(dl[a]);
        ih[a] = rcp_approx(dh[a]);
        // whole tile on the + side of the apex: t >= (blo - apex) / dhi and t <= (bhi - apex) / dlo;
        // whole tile on the - side:            t >= (bhi - apex) / dlo and t <= (blo - apex) / dhi
        if (lp != hp) sorted = 0;
        nOff[a] = (lp ? a : 3 + a) * SHAFT_WIDTH; fOff[a] = (lp ? 3 + a : a) * SHAFT_WIDTH;
        ni[a] = lp ? ih[a] : il[a]; fi[a] = lp ? il[a] : ih[a];
    }
    s.ilx = il[0]; s.ily = il[1]; s.ilz = il[2]; s.ihx = ih[0]; s.ihy = ih[1]; s.ihz = ih[2];
    s.signs = sg;
    s.sorted = sorted;
    s.nOffX = opaque(nOff[0]); s.nOffY = opaque(nOff[1]); s.nOffZ = opaque(nOff[2]);
    s.fOffX = opaque(fOff[0]); s.fOffY = opaque(fOff[1]); s.fOffZ = opaque(fOff[2]);
    s.nix = opaque(ni[0]); s.niy = opaque(ni[1]); s.niz = opaque(ni[2]); s.fix = opaque(fi[0]); s.fiy = opaque(fi[1]); s.fiz = opaque(fi[2]);
    return s;
}

// exists t in [tmin, tmax] with  apex + t * dlo <= bhi  and  apex + t * dhi >= blo  on every axis
__device__ __forceinline__ bool shaft_overlap(const Shaft& s, float blx, float bly, float blz, float bhx, float bhy, float bhz,
                                              float tmin, float tmax) {
    const float ax = (bhx - s.apex.x) * s.ilx, bx = (blx - s.apex.x) * s.ihx;
    const float ay = (bhy - s.apex.y) * s.ily, by = (bly - s.apex.y) * s.ihy;
    const float az = (bhz - s.apex.z) * s.ilz, bz = (blz - s.apex.z) * s.ihz;
    const float NI = -INFINITY, PI = INFINITY;
    // dlo > 0: a is an upper bound, else a lower bound;  dhi > 0: b is a lower bound, else an upper bound
    float lo = fmaxf(fmaxf((s.signs & 1u) ? NI : ax, (s.signs & 8u) ? bx : NI), fmaxf((s.signs & 2u) ? NI : ay, (s.signs & 16u) ? by : NI));
    lo = fmaxf(lo, fmaxf(fmaxf((s.signs & 4u) ? NI : az, (s.signs & 32u) ? bz : NI), tmin));
    float hi = fminf(fminf((s.signs & 1u) ? ax : PI, (s.signs & 8u) ? PI : bx), fminf((s.signs & 2u) ? ay : PI, (s.signs & 16u) ? PI : by));
    hi = fminf(hi, fminf(fminf((s.signs & 4u) ? az : PI, (s.signs & 32u) ? PI : bz), tmax));
    return lo <= hi;
}

// Returns per lane whether its ray (org = shaft apex, dir) is occluded.  Must be called by all 32 lanes.
// stackBase / candBase are shared-window byte addresses of the warp's stack (BVH_STACK words) and candidate list
// (SHAFT_CAND words); laneOff = lane index (node rows are indexed base + node * 224 + row offset + lane, in floats).
__device__ inline bool trace_any_warp_shaft(const DevScene& sc, bool active, V3 org, V3 dir, float tmin, float tmax, const Shaft& sh,
                                            uint32_t stackBase, uint32_t candBase, uint32_t* warpStack,
                                            int candMax /* more candidate leaves than this => per-ray packet traversal */, int* overflow,
                                            unsigned* counters /* [0] fallbacks [1] node visits [2] candidate leaves (per warp) */,
                                            bool forcePacket = false /* skip the descent: the caller expects it to overflow */) {
    const unsigned full = 0xffffffffu;
    if (sc.numShaftNodes == 0 || !__any_sync(full, active)) return false;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const float* base = reinterpret_cast<const float*>(sc.shaftNodes);
    constexpr uint32_t NODE_F = sizeof(ShaftNode) / 4;  // floats per node
    uint32_t sp = stackBase, cp = candBase;
    const uint32_t spEnd = stackBase + 4u * BVH_STACK, cpEnd = candBase + 4u * (uint32_t)candMax;
    uint32_t cur = 0;
    bool fallback = forcePacket;
    while (!forcePacket) {
        counters[1]++;
        const uint32_t idx = cur * NODE_F + lane;
        const uint32_t word = __float_as_uint(__ldg(base + idx + 6 * SHAFT_WIDTH));
        bool hit;
        if (sh.sorted) {
            const float tnx = (__ldg(base + idx + sh.nOffX) - sh.apex.x) * sh.nix, tfx = (__ldg(base + idx + sh.fOffX) - sh.apex.x) * sh.fix;
            const float tny = (__ldg(base + idx + sh.nOffY) - sh.apex.y) * sh.niy, tfy = (__ldg(base + idx + sh.fOffY) - sh.apex.y) * sh.fiy;
            const float tnz = (__ldg(base + idx + sh.nOffZ) - sh.apex.z) * sh.niz, tfz = (__ldg(base + idx + sh.fOffZ) - sh.apex.z) * sh.fiz;
            hit = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin)) <= fminf(fminf(tfx, tfy), fminf(tfz, tmax));
        } else {
            hit = shaft_overlap(sh, __ldg(base + idx), __ldg(base + idx + SHAFT_WIDTH), __ldg(base + idx + 2 * SHAFT_WIDTH),
                                __ldg(base + idx + 3 * SHAFT_WIDTH), __ldg(base + idx + 4 * SHAFT_WIDTH), __ldg(base + idx + 5 * SHAFT_WIDTH),
                                tmin, tmax);
        }
        hit = hit && word != BVH_EMPTY;
        const bool leaf = (word & BVH_LEAF_BIT) != 0u;
        const unsigned mi = __ballot_sync(full, hit && !leaf), ml = __ballot_sync(full, hit && leaf);
        const uint32_t spNew = sp + 4u * (uint32_t)__popc(mi), cpNew = cp + 4u * (uint32_t)__popc(ml);
        if (spNew > spEnd || cpNew > cpEnd) { fallback = true; break; }
        if (hit && !leaf) st_shared_u32(sp + 4u * (uint32_t)__popc(mi & lt), word);
        if (hit && leaf) st_shared_u32(cp + 4u * (uint32_t)__popc(ml & lt), (cur << 5) | lane);  // slot of the leaf: its word AND its box
        sp = spNew; cp = cpNew;
        __syncwarp();
        if (sp == stackBase) break;
        sp -= 4u;
        cur = ld_shared_u32(sp);
    }
    if (fallback) {  // fat shaft (tile across a depth edge) or cluttered region: per-ray packet traversal
        counters[0]++;
        __syncwarp();
        return trace_any_warp(sc, active, org, dir, tmin, tmax, warpStack, overflow);
    }
    const uint32_t cn = (cp - candBase) >> 2;
    counters[2] += cn;
    bool occ = false;
    if (cn) {
        // A leaf is a candidate when its box meets the SHAFT; most of them meet no individual ray of it.  One slab test
        // per ray against the (padded) leaf box -- the same conservative test the per-ray traversals descend by --
        // skips the exact triangle tests of such leaves for the whole warp.
        const RaySlabM rs = make_slab_masked(org, dir);
        for (uint32_t k = 0; k < cn; k++) {
            const uint32_t slot = ld_shared_u32(candBase + 4u * k);
            const float* nb = base + (slot >> 5) * NODE_F + (slot & 31u);
            const bool inBox = active && !occ &&
                               slab_masked(rs, __ldg(nb), __ldg(nb + SHAFT_WIDTH), __ldg(nb + 2 * SHAFT_WIDTH), __ldg(nb + 3 * SHAFT_WIDTH),
                                           __ldg(nb + 4 * SHAFT_WIDTH), __ldg(nb + 5 * SHAFT_WIDTH), tmin, tmax);
            if (!__any_sync(full, inBox)) continue;
            const uint32_t w = __float_as_uint(__ldg(nb + 6 * SHAFT_WIDTH));
            const uint32_t first = bvh_leaf_first(w), count = bvh_leaf_count(w);
            const float4* tp = sc.triLeaf + 4 * (size_t)first;
            for (uint32_t j = 0; j < count; j++, tp += 4) {
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2), d = __ldg(tp + 3);
                float t, be, ga;
                occ |= tri_test(org, dir, tmin, tmax, ld3(a), ld3(b), ld3(c), ld3(d), &t, &be, &ga);
            }
            if (!__any_sync(full, active && !occ)) break;
        }
    }
    __syncwarp();
    return active && occ;
}

#endif  // __CUDACC__

}  // namespace evplp
