// gather_fast.cu -- the VPL-cluster gather: splatColor + vplSplat + rtMaterialAnyHit (lighttracing.cu:184-188, 275-379)
// with the shadow-ray traversal amortised over CLUSTERS of VPLs.
//
// What changes against gather_vpl_kernel (stages.cu), which walks the hierarchy once per (8x4-pixel tile, VPL):
//   * the usable VPLs are sorted by the Morton code of their position and cut into runs of `gather_cluster_size` (16)
//     consecutive VPLs; a run whose box is large (the Morton order jumped inside it) is cut into halves / quarters
//     (cluster_layout_kernel), so a cluster's positions have a small bounding box;
//   * a warp descends the 32-wide hierarchy ONCE per (cluster, tile) with a double shaft -- every ray of the cluster starts
//     inside the cluster box and ends inside the tile's box of surface points, so at parameter t it lies inside the
//     interpolated box [clo + t (tlo - clo), chi + t (thi - chi)] -- and collects the candidate leaves, with their boxes,
//     in shared memory;
//   * per (VPL, tile) step only that short list is filtered, 32 candidates at a time (one per lane), against the thin
//     shaft (VPL point -> tile box); the survivors get the per-ray slab test of the leaf box, which builds each ray's OWN
//     list of boxes, and the exact triangle tests then run max-boxes-per-ray times per step;
//   * a cluster shaft that overflows its candidate batches (clutter between the cluster and the tile) falls back to one thin
//     descent per VPL; tiles with a long depth range are split into depth groups with compact boxes.
// Hit decisions are the same as everywhere else: tri_test rounds every operation explicitly (device_scene.h), the shafts
// and boxes are conservative culls.  What is NOT bit-identical to the oracle: each pixel sums its VPLs in Morton order
// instead of record order, and the shading tail below is compiled with FMA contraction, rsqrt / rcp approximations and a
// polynomial pow (relative error < 1e-5 per pair, measured; north_star tolerance for radiance: 1e-4 per pixel).  The
// bit-exact path stays available (gather_chunks = 1 or gather_algo = 0) and is what the parity tests compare bit for bit.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include "context.h"

namespace evplp {

constexpr int FG_CL_MAX = 16;      // staging capacity: VPLs per cluster
constexpr int FG_PV = 6;           // float4 per prepared VPL
constexpr int FG_CAND = 96;        // candidate leaves per batch of a descent (a node adds up to 32)
// (run-time knobs, FastParams: sharedBatches = candidate batches a cluster's shared descent may stream before its undecided rays go
//  per VPL; vplBatches = batches a single VPL streams before it switches to the packet traversal; skipMax)
constexpr int FG_SPLIT_ROUNDS = 3;  // depth groups per tile: up to 2^rounds
constexpr float FG_SPLIT_RATIO = 2.5f;   // a group is cut while its distance range exceeds this many tile widths
constexpr int FG_STACK = 96;       // inner-node stack of the descent (also the packet fallback's stack)
constexpr int FG_SHAFT_WORDS = 24;  // the cluster shaft's constants (words 0-12) and the depth group's tile box (16-21): warp-uniform, kept in shared memory, not in registers
constexpr int FG_WARP_WORDS = FG_CL_MAX * FG_PV * 4 + FG_STACK + 7 * FG_CAND + FG_SHAFT_WORDS;

struct FastParams {
    GatherParams g;
    int clusterSize;
    uint32_t count;
    const uint32_t* clusterList;    // packed (first << 5 | count) per cluster, Morton order
    const uint32_t* numClusters;    // how many (device: the layout is cut on the device)
    int tilesX, pitchX;             // 8x4-pixel tiles per row of the launch rectangle; row pitch of the tile numbering
    uint32_t ownedTiles;            // tiles t = offset + k * stride, k < ownedTiles (this handle's share of the image)
    uint32_t stride, offset;
    uint32_t numChunks;             // the cluster list is cut into numChunks ranges; work item = (tile, range)
    float tileAngle;                // width of an 8-pixel tile per unit of distance from the camera
    int sharedBatches, vplBatches, skipMax;
};

// ---- fast math of the shading tail ---------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// x^e for x in (1e-6, ~1], e >= 0: log2 by a degree-8 polynomial of m - 1 (m in [2/3, 4/3): relative error 1.4e-7, so the
// product e * log2 x keeps its accuracy next to x = 1 where large exponents matter), then ex2.approx.
__device__ __forceinline__ float fast_pow(float x, float e) {
    const int ix = __float_as_int(x);
    const int ex = (ix - 0x3f2aaaab) & 0xff800000;
    const float m = __int_as_float(ix - ex);
    const float fe = (float)(ex >> 23);
    const float t = m - 1.0f;
    float p = 0.2082485556602478f;
    p = fmaf(p, t, -0.22222177684307098f);
    p = fmaf(p, t, 0.20058539509773254f);
    p = fmaf(p, t, -0.23678641021251678f);
    p = fmaf(p, t, 0.2887956202030182f);
    p = fmaf(p, t, -0.36078956723213196f);
    p = fmaf(p, t, 0.4808940887451172f);
    p = fmaf(p, t, -0.7213465571403503f);
    p = fmaf(p, t, 1.4426950216293335f);
    return ex2_approx(e * fmaf(t, p, fe));
}

// ---- preparation: Morton order, prepared VPLs, cluster boxes ---------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void vpl_code_kernel(const EvplpRecord* __restrict__ records, const uint32_t* __restrict__ vplList, uint32_t count,
                                float3 smin, float3 scale, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t r = vplList[i];
    const float* p = records[r].position;
    const float qx = fminf(fmaxf((p[0] - smin.x) * scale.x, 0.f), 1023.f);
    const float qy = fminf(fmaxf((p[1] - smin.y) * scale.y, 0.f), 1023.f);
    const float qz = fminf(fmaxf((p[2] - smin.z) * scale.z, 0.f), 1023.f);
    keys[i] = (spread10((uint32_t)qx) << 2) | (spread10((uint32_t)qy) << 1) | spread10((uint32_t)qz);
    vals[i] = r;
}

// pv[6 j .. 6 j + 5] = {pos, e} {n, pSel} {flux, (e + 2) / 2pi} {refl, (e + 1) / 2pi * (1 - pSel) or 0} {kd / pi, pSel / pi} {ks, -}
__global__ void vpl_prepare_kernel(const EvplpRecord* __restrict__ records, const uint32_t* __restrict__ order, uint32_t count,
                                   float4* __restrict__ pv) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const float4* r = reinterpret_cast<const float4*>(records + order[j]);
    const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2), d = __ldg(r + 3), e = __ldg(r + 4), f = __ldg(r + 5);
    const float ex = f.w, pSel = b.w;
    // reflect(-fluxDir, n) = -fluxDir + 2 n (n . fluxDir)      (PhongEvalF / PhongPdfA, rtmaterial.cuh:88-119)
    const float nd = b.x * d.x + b.y * d.y + b.z * d.z;
    const float rx = 2.0f * nd * b.x - d.x, ry = 2.0f * nd * b.y - d.y, rz = 2.0f * nd * b.z - d.z;
    float4* o = pv + (size_t)j * FG_PV;
    o[0] = make_float4(a.x, a.y, a.z, ex);
    o[1] = make_float4(b.x, b.y, b.z, pSel);
    o[2] = make_float4(c.x, c.y, c.z, (ex + 2.0f) * 0.5f * kInvPi);
    o[3] = make_float4(rx, ry, rz, f.x <= 0.000001f ? 0.0f : (ex + 1.0f) * 0.5f * kInvPi * (1.0f - pSel));
    o[4] = make_float4(e.x * kInvPi, e.y * kInvPi, e.z * kInvPi, pSel * kInvPi);
    o[5] = make_float4(f.x, f.y, f.z, 0.0f);
}

// Cluster layout.  The Morton order jumps across the scene wherever a high bit of the code changes, and a run of `clusterSize`
// VPLs that spans such a jump has a box far larger than its VPLs need -- a fat double shaft for every tile.  Each run is therefore
// kept whole only while its box is small (longest edge <= maxExtent); otherwise it is cut into halves, and those once more into
// quarters (never below 4 VPLs).  Output: up to 4 packed words (first << 5 | count, count = 0: unused) per run.
__global__ void cluster_layout_kernel(const float4* __restrict__ pv, uint32_t count, int clusterSize, uint32_t numRuns, float maxExtent,
                                      uint32_t* __restrict__ slots) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= numRuns) return;
    const uint32_t first = r * (uint32_t)clusterSize, n = min((uint32_t)clusterSize, count - first);
    auto extent = [&](uint32_t a, uint32_t b) {   // longest edge of the box of VPLs [a, b)
        float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
        for (uint32_t j = a; j < b; j++) {
            const float4 p = pv[(size_t)j * FG_PV];
            lx = fminf(lx, p.x); ly = fminf(ly, p.y); lz = fminf(lz, p.z);
            hx = fmaxf(hx, p.x); hy = fmaxf(hy, p.y); hz = fmaxf(hz, p.z);
        }
        return fmaxf(hx - lx, fmaxf(hy - ly, hz - lz));
    };
    uint32_t out[4] = {0u, 0u, 0u, 0u};
    int k = 0;
    if (n < 8u || !(extent(first, first + n) > maxExtent)) {
        out[k++] = (first << 5) | n;
    } else {
        const uint32_t h = n / 2;
        for (int half = 0; half < 2; half++) {
            const uint32_t a = first + (half ? h : 0u), m = half ? n - h : h;
            if (m < 8u || !(extent(a, a + m) > maxExtent)) {
                out[k++] = (a << 5) | m;
            } else {
                out[k++] = (a << 5) | (m / 2);
                out[k++] = ((a + m / 2) << 5) | (m - m / 2);
            }
        }
    }
    for (int q = 0; q < 4; q++) slots[4 * (size_t)r + q] = out[q];
}

struct ClusterUsed {
    __device__ bool operator()(uint32_t w) const { return (w & 31u) != 0u; }
};

__global__ void cluster_bounds_kernel(const float4* __restrict__ pv, const uint32_t* __restrict__ clusterList,
                                      const uint32_t* __restrict__ numClusters, float4* __restrict__ cbox) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *numClusters) return;
    const uint32_t w = clusterList[c], first = w >> 5, last = first + (w & 31u);
    float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
    for (uint32_t j = first; j < last; j++) {
        const float4 p = pv[(size_t)j * FG_PV];
        lx = fminf(lx, p.x); ly = fminf(ly, p.y); lz = fminf(lz, p.z);
        hx = fmaxf(hx, p.x); hy = fmaxf(hy, p.y); hz = fmaxf(hz, p.z);
    }
    cbox[2 * c] = make_float4(lx, ly, lz, 0.f);
    cbox[2 * c + 1] = make_float4(hx, hy, hz, 0.f);
}

// ---- the double shaft ------------------------------------------------------------------------------------------------
// Rays start in [alo, ahi] (t = 0) and end in [tlo, thi] (t = 1).  A box [bl, bh] can meet one of them only if some t in
// [tmin, tmax] has  alo + t (tlo - alo) <= bh  and  ahi + t (thi - ahi) >= bl  on all three axes.  Per axis that is two
// half-lines in t: A = (bh - alo) / (tlo - alo) bounds t from below when tlo < alo and from above otherwise,
// B = (bl - ahi) / (thi - ahi) bounds t from below when thi > ahi and from above otherwise.
struct DShaft {
    float ilx, ily, ilz, ihx, ihy, ihz;   // 1 / (tlo - alo), 1 / (thi - ahi)
    float oAx, oAy, oAz, oBx, oBy, oBz;   // -alo * il, -ahi * ih
    unsigned signs;                       // bit a: A_a is a LOWER bound, bit 3 + a: B_a is a LOWER bound
};

__device__ __forceinline__ DShaft make_dshaft(V3 alo, V3 ahi, V3 tlo, V3 thi) {
    DShaft s;
    float dl[3] = {tlo.x - alo.x, tlo.y - alo.y, tlo.z - alo.z};
    float dh[3] = {thi.x - ahi.x, thi.y - ahi.y, thi.z - ahi.z};
    const float al[3] = {alo.x, alo.y, alo.z}, ah[3] = {ahi.x, ahi.y, ahi.z};
    float il[3], ih[3], oA[3], oB[3];
    unsigned sg = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        // a zero extent keeps its constraint ("t * 0 <= bh - alo" / "t * 0 >= bl - ahi") as a huge bound of the right sign
        if (fabsf(dl[a]) < 1e-30f) dl[a] = -1e-30f;
        if (fabsf(dh[a]) < 1e-30f) dh[a] = 1e-30f;
        if (dl[a] < 0.f) sg |= 1u << a;
        if (dh[a] > 0.f) sg |= 8u << a;
        il[a] = rcp_approx(dl[a]);
        ih[a] = rcp_approx(dh[a]);
        oA[a] = -al[a] * il[a];
        oB[a] = -ah[a] * ih[a];
    }
    s.ilx = il[0]; s.ily = il[1]; s.ilz = il[2]; s.ihx = ih[0]; s.ihy = ih[1]; s.ihz = ih[2];
    s.oAx = oA[0]; s.oAy = oA[1]; s.oAz = oA[2]; s.oBx = oB[0]; s.oBy = oB[1]; s.oBz = oB[2];
    s.signs = sg;
    return s;
}

__device__ __forceinline__ bool dshaft_overlap(const DShaft& s, float blx, float bly, float blz, float bhx, float bhy, float bhz,
                                               float tmin, float tmax) {
    const float ax = fmaf(bhx, s.ilx, s.oAx), bx = fmaf(blx, s.ihx, s.oBx);
    const float ay = fmaf(bhy, s.ily, s.oAy), by = fmaf(bly, s.ihy, s.oBy);
    const float az = fmaf(bhz, s.ilz, s.oAz), bz = fmaf(blz, s.ihz, s.oBz);
    const float NI = -INFINITY, PI = INFINITY;
    float lo = fmaxf(fmaxf((s.signs & 1u) ? ax : NI, (s.signs & 8u) ? bx : NI), fmaxf((s.signs & 2u) ? ay : NI, (s.signs & 16u) ? by : NI));
    lo = fmaxf(lo, fmaxf(fmaxf((s.signs & 4u) ? az : NI, (s.signs & 32u) ? bz : NI), tmin));
    float hi = fminf(fminf((s.signs & 1u) ? PI : ax, (s.signs & 8u) ? PI : bx), fminf((s.signs & 2u) ? PI : ay, (s.signs & 16u) ? PI : by));
    hi = fminf(hi, fminf(fminf((s.signs & 4u) ? PI : az, (s.signs & 32u) ? PI : bz), tmax));
    return lo <= hi;
}

__device__ __forceinline__ float ld_shared_f32(uint32_t addr) { return __uint_as_float(ld_shared_u32(addr)); }

// Resumable descent of the 32-wide hierarchy with the double shaft.  Pops nodes from the warp's stack (sp is kept by the
// caller between calls) and appends the candidate leaves (6 box planes + child word each) to the warp's shared-memory list,
// until the stack is empty (*done = true) or the next node could overflow the list (a node has up to 32 leaf children): the
// caller consumes the batch and calls again.  Nothing is ever thrown away.  Returns the batch size, or -1 when the STACK
// would overflow (the caller falls back to the per-ray packet traversal; not seen on the bundled scenes).
__device__ __forceinline__ int shaft_collect_batch(const DevScene& sc, uint32_t shaftBase, float tmin, float tmax, uint32_t stackBase,
                                                   uint32_t& sp, uint32_t candBase, bool* done, unsigned& nodeVisits) {
    const unsigned full = 0xffffffffu;
    DShaft sh;
    sh.ilx = ld_shared_f32(shaftBase); sh.ily = ld_shared_f32(shaftBase + 4u); sh.ilz = ld_shared_f32(shaftBase + 8u);
    sh.ihx = ld_shared_f32(shaftBase + 12u); sh.ihy = ld_shared_f32(shaftBase + 16u); sh.ihz = ld_shared_f32(shaftBase + 20u);
    sh.oAx = ld_shared_f32(shaftBase + 24u); sh.oAy = ld_shared_f32(shaftBase + 28u); sh.oAz = ld_shared_f32(shaftBase + 32u);
    sh.oBx = ld_shared_f32(shaftBase + 36u); sh.oBy = ld_shared_f32(shaftBase + 40u); sh.oBz = ld_shared_f32(shaftBase + 44u);
    sh.signs = ld_shared_u32(shaftBase + 48u);
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const float* base = reinterpret_cast<const float*>(sc.shaftNodes);
    constexpr uint32_t NODE_F = sizeof(ShaftNode) / 4;
    const uint32_t spEnd = stackBase + 4u * FG_STACK;
    uint32_t cn = 0;
    for (;;) {
        if (sp == stackBase) { *done = true; break; }
        if (cn > (uint32_t)(FG_CAND - SHAFT_WIDTH)) { *done = false; break; }
        sp -= 4u;
        const uint32_t cur = ld_shared_u32(sp);
        nodeVisits++;
        const uint32_t idx = cur * NODE_F + lane;
        const uint32_t word = __float_as_uint(__ldg(base + idx + 6 * SHAFT_WIDTH));
        const float lx = __ldg(base + idx), ly = __ldg(base + idx + SHAFT_WIDTH), lz = __ldg(base + idx + 2 * SHAFT_WIDTH);
        const float hx = __ldg(base + idx + 3 * SHAFT_WIDTH), hy = __ldg(base + idx + 4 * SHAFT_WIDTH), hz = __ldg(base + idx + 5 * SHAFT_WIDTH);
        const bool hit = dshaft_overlap(sh, lx, ly, lz, hx, hy, hz, tmin, tmax) && word != BVH_EMPTY;
        const bool leaf = (word & BVH_LEAF_BIT) != 0u;
        const unsigned mi = __ballot_sync(full, hit && !leaf), ml = __ballot_sync(full, hit && leaf);
        // push slot / candidate slot of this lane, pinned in registers BEFORE sp moves (nvcc 12.9 otherwise rebuilds the push
        // address from the updated stack pointer and gets it wrong for more than one push: seen in the SASS, caught by memcheck)
        const uint32_t pushAt = opaque(sp + 4u * (uint32_t)__popc(mi & lt));
        const uint32_t candAt = candBase + 4u * (cn + (uint32_t)__popc(ml & lt));
        const uint32_t spNew = sp + 4u * (uint32_t)__popc(mi);
        if (spNew > spEnd) return -1;
        __syncwarp();   // every lane has read `cur` before the slot is overwritten
        if (hit && !leaf) st_shared_u32(pushAt, word);
        if (hit && leaf) {
            st_shared_u32(candAt, __float_as_uint(lx)); st_shared_u32(candAt + 4u * FG_CAND, __float_as_uint(ly));
            st_shared_u32(candAt + 8u * FG_CAND, __float_as_uint(lz)); st_shared_u32(candAt + 12u * FG_CAND, __float_as_uint(hx));
            st_shared_u32(candAt + 16u * FG_CAND, __float_as_uint(hy)); st_shared_u32(candAt + 20u * FG_CAND, __float_as_uint(hz));
            st_shared_u32(candAt + 24u * FG_CAND, word);
        }
        sp = spNew; cn += (uint32_t)__popc(ml);
        __syncwarp();
    }
    return (int)cn;
}

// Exact visibility of one VPL's rays against the warp's current candidate batch.  FILTER: the batch came from a CLUSTER shaft,
// so it is first filtered, 32 candidates at a time (one per lane), against the thin shaft (VPL -> tile); without FILTER the
// batch came from that thin shaft itself.  Survivors: per-ray slab test of the leaf box, then the exact triangle tests.
// the depth group's box of surface points (the far end of every shaft), parked in the warp's shared-memory slice
__device__ __forceinline__ V3 ld_tile_lo(uint32_t shaftBase) { return v3(ld_shared_f32(shaftBase + 64u), ld_shared_f32(shaftBase + 68u), ld_shared_f32(shaftBase + 72u)); }
__device__ __forceinline__ V3 ld_tile_hi(uint32_t shaftBase) { return v3(ld_shared_f32(shaftBase + 76u), ld_shared_f32(shaftBase + 80u), ld_shared_f32(shaftBase + 84u)); }
// G-buffer texel re-read where it is needed (L1-resident) instead of living in registers across the visibility phase; volatile so
// that the compiler does not hoist the load back out of the cluster loop
__device__ __forceinline__ float4 ld_texel(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <bool FILTER>
__device__ __forceinline__ bool test_batch(const DevScene& sc, uint32_t candBase, int cn, bool active, V3 org, V3 dir, uint32_t shaftBase,
                                           float tmin, float tmax) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    DShaft ts;
    if (FILTER) ts = make_dshaft(org, org, ld_tile_lo(shaftBase), ld_tile_hi(shaftBase));
    RaySlabM rs;
    bool haveSlab = false, occ = false;
    for (int c0 = 0; c0 < cn; c0 += 32) {
        unsigned km;
        if (FILTER) {
            const int k = c0 + lane;
            const uint32_t ca = candBase + 4u * (uint32_t)(k < cn ? k : 0);
            const bool keep = k < cn && dshaft_overlap(ts, ld_shared_f32(ca), ld_shared_f32(ca + 4u * FG_CAND), ld_shared_f32(ca + 8u * FG_CAND),
                                                       ld_shared_f32(ca + 12u * FG_CAND), ld_shared_f32(ca + 16u * FG_CAND),
                                                       ld_shared_f32(ca + 20u * FG_CAND), tmin, tmax);
            km = __ballot_sync(full, keep);
        } else {
            km = cn - c0 >= 32 ? 0xffffffffu : ((1u << (cn - c0)) - 1u);
        }
        if (km && !haveSlab) { rs = make_slab_masked(org, dir); haveSlab = true; }
        // phase 1: per-ray slab test of every surviving leaf box -> this lane's own list (bit q - c0)
        unsigned mine = 0u;
        while (km) {
            const int b = __ffs((int)km) - 1;
            km &= km - 1u;
            const uint32_t qa = candBase + 4u * (uint32_t)(c0 + b);
            const bool inBox = slab_masked(rs, ld_shared_f32(qa), ld_shared_f32(qa + 4u * FG_CAND), ld_shared_f32(qa + 8u * FG_CAND),
                                           ld_shared_f32(qa + 12u * FG_CAND), ld_shared_f32(qa + 16u * FG_CAND),
                                           ld_shared_f32(qa + 20u * FG_CAND), tmin, tmax);
            if (inBox) mine |= 1u << b;
        }
        if (!active || occ) mine = 0u;
        // phase 2: every lane runs the exact triangle tests of ITS OWN boxes, so the warp iterates max-boxes-per-ray times
        // instead of once per leaf that any of its rays meets (near misses dominate: most boxes are met by a few rays only)
        while (__any_sync(full, mine != 0u)) {
            uint32_t tfirst = 0u, tcount = 0u;
            if (mine) {
                const int b = __ffs((int)mine) - 1;
                mine &= mine - 1u;
                const uint32_t w = ld_shared_u32(candBase + 4u * (uint32_t)(c0 + b) + 24u * FG_CAND);
                tfirst = bvh_leaf_first(w); tcount = bvh_leaf_count(w);
            }
            for (uint32_t u = 0; __any_sync(full, u < tcount); u++) {
                if (u < tcount) {
                    const float4* tp = sc.triLeaf + 4 * (size_t)(tfirst + u);
                    const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2), td = __ldg(tp + 3);
                    float tt, be, ga;
                    occ |= tri_test(org, dir, tmin, tmax, ld3(ta), ld3(tb), ld3(tc), ld3(td), &tt, &be, &ga);
                }
            }
            if (occ) mine = 0u;
        }
        if (!__any_sync(full, active && !occ)) return occ;
    }
    return occ;
}

__device__ __forceinline__ void store_shaft(uint32_t shaftBase, const DShaft& sh) {   // every lane writes the same words
    st_shared_u32(shaftBase, __float_as_uint(sh.ilx)); st_shared_u32(shaftBase + 4u, __float_as_uint(sh.ily)); st_shared_u32(shaftBase + 8u, __float_as_uint(sh.ilz));
    st_shared_u32(shaftBase + 12u, __float_as_uint(sh.ihx)); st_shared_u32(shaftBase + 16u, __float_as_uint(sh.ihy)); st_shared_u32(shaftBase + 20u, __float_as_uint(sh.ihz));
    st_shared_u32(shaftBase + 24u, __float_as_uint(sh.oAx)); st_shared_u32(shaftBase + 28u, __float_as_uint(sh.oAy)); st_shared_u32(shaftBase + 32u, __float_as_uint(sh.oAz));
    st_shared_u32(shaftBase + 36u, __float_as_uint(sh.oBx)); st_shared_u32(shaftBase + 40u, __float_as_uint(sh.oBy)); st_shared_u32(shaftBase + 44u, __float_as_uint(sh.oBz));
    st_shared_u32(shaftBase + 48u, sh.signs);
}

// cold path (the descent's stack would overflow): kept out of line so that its registers do not weigh on the main loop
__device__ __noinline__ bool packet_any_hit(const WideNode* nodes, int numNodes, const float4* triLeaf, bool active, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float tmin, float tmax, uint32_t* warpStack, int* overflow) {
    DevScene sc;
    sc.nodes = nodes; sc.numNodes = numNodes; sc.triLeaf = triLeaf;
    return trace_any_warp(sc, active, v3(ox, oy, oz), v3(dx, dy, dz), tmin, tmax, warpStack, overflow);
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
// MC (mode class): 0 = misMode one, 1 = balance / max / power2 (MIS weight), 2 = geometryClamp, 3 = geometryBrdfClamp
template <int MC, int MINB>
__global__ void __launch_bounds__(GATHER_WARPS * 32, MINB)
gather_cluster_kernel(DevScene sc, FastParams fp, const float4* __restrict__ gbuf, const float4* __restrict__ pv,
                      const float4* __restrict__ cbox, long long* __restrict__ acc, DevStats* stats,
                      uint32_t* __restrict__ tileCounter, const uint32_t* __restrict__ tileOrder, uint32_t* __restrict__ tileCost) {
    __shared__ __align__(16) uint32_t smemAll[GATHER_WARPS][FG_WARP_WORDS];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* spv = reinterpret_cast<float4*>(smemAll[warp]);
    uint32_t* stackPtr = smemAll[warp] + FG_CL_MAX * FG_PV * 4;
    const uint32_t stackBase = opaque((uint32_t)__cvta_generic_to_shared(stackPtr));
    const uint32_t candBase = opaque(stackBase + 4u * FG_STACK);
    const uint32_t shaftBase = opaque(candBase + 4u * 7u * FG_CAND);
    const GatherParams& gp = fp.g;
    const uint32_t vTotal = fp.ownedTiles * fp.numChunks;
    const float tmin = (float)0.0001, tmax = (float)(1 - 0.0001);   // Ray(vpl.pos, -v12, shadow, 0.0001, 1 - 0.0001) -- lighttracing.cu:292
    unsigned rays = 0, steps = 0, descents = 0, nodeVisits = 0, candTotal = 0, batches = 0, packets = 0, splits = 0;
#ifdef EVPLP_GATHER_PROF   // tuning build only: cycles per region of the item loop (clock64 deltas per warp)
    long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long profT = 0;
#define FG_PROF_BEGIN() profT = clock64()
#define FG_PROF_END(k) do { const long long now_ = clock64(); prof[k] += now_ - profT; profT = now_; } while (0)
#else
#define FG_PROF_BEGIN()
#define FG_PROF_END(k)
#endif
#ifdef EVPLP_GATHER_HIST   // tuning build only: the dynamically indexed histogram lives in local memory
    unsigned hist[12] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#define FG_HIST(x) x
#else
#define FG_HIST(x)
#endif
    int ovf = 0;
    for (;;) {
        uint32_t v;
        if (lane == 0) v = atomicAdd(tileCounter, 1u);
        v = __shfl_sync(full, v, 0);
        if (v >= vTotal) break;
        if (tileOrder) v = __ldg(tileOrder + v);  // most expensive items of the previous launch first
        const long long itemStart = clock64();
        FG_PROF_BEGIN();
        const uint32_t chunk = v / fp.ownedTiles, t = fp.offset + (v % fp.ownedTiles) * fp.stride;
        const int ty = (int)(t / (uint32_t)fp.pitchX), tx = (int)(t % (uint32_t)fp.pitchX);
        const int x = gp.x0 + tx * 8 + (lane & 7), y = gp.y0 + ty * 4 + (lane >> 3);
        const bool inside = tx < fp.tilesX && x < gp.x1 && y < gp.y1;
        const size_t n = (size_t)gp.W * gp.H;
        const size_t i = inside ? (size_t)y * gp.W + x : 0;
        const float4 g0 = gbuf[i];
        const bool valid = inside && g0.w != 0.0f;
        const float px = g0.x, py = g0.y, pz = g0.z;
        // A tile whose surface points spread over a long range of distances -- a depth edge inside it, or a floor seen at a
        // grazing angle -- has a long box, and every shaft towards it is a wide fan that meets far more geometry than its 32 rays
        // do.  Its lanes are split by distance from the camera, in up to three rounds of "cut the group's range in the middle
        // while it is longer than FG_SPLIT_RATIO tile widths", into at most eight depth groups, which are gathered one after the
        // other, each with its own compact box.  (More steps, but each of them cheap: the cost of a step grows much faster
        // than linearly with the fan's width, because wide fans also lose the shared cluster descents.)
        int grp = 0;
        {
            const float dx = px - gp.cameraPosition.x, dy = py - gp.cameraPosition.y, dz = pz - gp.cameraPosition.z;
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
            for (int round = 0; round < FG_SPLIT_ROUNDS; round++) {
                int newBit = 0;
                for (int g = 0; g < (1 << round); g++) {   // every lane walks every group: the shuffles stay warp-uniform
                    const bool in = valid && grp == g;
                    float lo = in ? dist : INFINITY, hi = in ? dist : -INFINITY;
                    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(full, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(full, hi, o)); }
                    const bool split = (hi - lo) > FG_SPLIT_RATIO * fp.tileAngle * lo;   // false for an empty group (inf - inf = NaN)
                    if (split && in && dist >= 0.5f * (lo + hi)) newBit = 1;
                }
                grp |= newBit << round;
            }
        }
        FG_PROF_END(8);
        unsigned groupsPresent = 0u;
        for (int g = 0; g < (1 << FG_SPLIT_ROUNDS); g++) if (__any_sync(full, valid && grp == g)) groupsPresent |= 1u << g;
        const uint32_t numClusters = __ldg(fp.numClusters);
        const uint32_t per = (numClusters + fp.numChunks - 1) / fp.numChunks;
        const uint32_t cBegin = min(numClusters, chunk * per);
        const uint32_t cEnd = min(numClusters, cBegin + per);
        float resx = 0.f, resy = 0.f, resz = 0.f;
        for (unsigned gmask = groupsPresent; gmask; gmask &= gmask - 1u) {
        const bool vg = valid && grp == (__ffs((int)gmask) - 1);
        // bounds of the group's surface points: the far end of every shaft
        {
            V3 tlo = vg ? v3(px, py, pz) : v3s(INFINITY), thi = vg ? v3(px, py, pz) : v3s(-INFINITY);
            for (int o = 16; o > 0; o >>= 1) {
                tlo = v3(fminf(tlo.x, __shfl_xor_sync(full, tlo.x, o)), fminf(tlo.y, __shfl_xor_sync(full, tlo.y, o)), fminf(tlo.z, __shfl_xor_sync(full, tlo.z, o)));
                thi = v3(fmaxf(thi.x, __shfl_xor_sync(full, thi.x, o)), fmaxf(thi.y, __shfl_xor_sync(full, thi.y, o)), fmaxf(thi.z, __shfl_xor_sync(full, thi.z, o)));
            }
            __syncwarp();   // (every lane writes the same words)
            st_shared_u32(shaftBase + 64u, __float_as_uint(tlo.x)); st_shared_u32(shaftBase + 68u, __float_as_uint(tlo.y)); st_shared_u32(shaftBase + 72u, __float_as_uint(tlo.z));
            st_shared_u32(shaftBase + 76u, __float_as_uint(thi.x)); st_shared_u32(shaftBase + 80u, __float_as_uint(thi.y)); st_shared_u32(shaftBase + 84u, __float_as_uint(thi.z));
            __syncwarp();
        }
        int skipLeft = 0, skipLen = 0;   // clusters that go straight to per-VPL descents after a fat cluster shaft
        for (uint32_t c = cBegin; c < cEnd; c++) {
            const uint32_t cw = __ldg(fp.clusterList + c);
            const uint32_t first = cw >> 5;
            const int nb = (int)(cw & 31u);
            __syncwarp();
            for (int k = lane; k < nb * FG_PV; k += 32) spv[k] = __ldg(pv + (size_t)first * FG_PV + k);
            __syncwarp();
            // which VPLs of the cluster light any pixel of the tile at all (cosine test of vplSplat, lighttracing.cu:282-288)
            // (actBits: bit j = THIS lane's pixel passes it; the visibility loops below only read bits, so the pixel's normal is not
            // live across them)
            unsigned live = 0u, actBits = 0u;
            {
                const float4 g1 = ld_texel(gbuf + n + i);
                for (int j = 0; j < nb; j++) {
                    const float4 a = spv[j * FG_PV], b = spv[j * FG_PV + 1];
                    const float vx = a.x - px, vy = a.y - py, vz = a.z - pz;
                    const float c1 = fmaxf(g1.x * vx + g1.y * vy + g1.z * vz, 0.0f);
                    const float c2 = fmaxf(-(b.x * vx + b.y * vy + b.z * vz), 0.0f);
                    const bool act = vg && !(c1 * c2 <= 0.000f);
                    if (act) actBits |= 1u << j;
                    if (__any_sync(full, act)) live |= 1u << j;
                }
            }
            FG_HIST(hist[8]++;)              // (cluster, tile) pairs with a valid pixel
            FG_PROF_END(1);
            if (!live) continue;
            FG_HIST(hist[9]++; hist[10] += (unsigned)__popc(live);)
            // ---- visibility: bit j of occBits = this lane's shadow ray to VPL j is occluded.
            // First choice: ONE descent for the whole cluster (double shaft from the cluster's box), whose candidate leaves every
            // VPL of the cluster then filters.  When that list does not fit one batch the shaft is fat -- the batch is dropped (a
            // bounded loss) and the cluster's VPLs descend one by one with their own thin shafts, streaming their candidates batch
            // by batch.  Fat shafts are a property of the tile (clutter, a depth edge), so after a failure the next `skip`
            // clusters do not even try, and `skip` doubles with every failure in a row.
            unsigned occBits = 0u;
            bool shared = nb > 1 && skipLeft == 0;
            if (skipLeft > 0) skipLeft--;
            if (shared) {
                const float4 bl = __ldg(cbox + 2 * (size_t)c), bh = __ldg(cbox + 2 * (size_t)c + 1);
                descents++;
                __syncwarp();
                store_shaft(shaftBase, make_dshaft(v3(bl.x, bl.y, bl.z), v3(bh.x, bh.y, bh.z), ld_tile_lo(shaftBase), ld_tile_hi(shaftBase)));
                st_shared_u32(stackBase, 0u);   // root (every lane writes the same word)
                uint32_t sp = stackBase + 4u;
                __syncwarp();
                bool done = false;
                int batchNo = 0;
                while (!done) {
                    const int cn = shaft_collect_batch(sc, shaftBase, tmin, tmax, stackBase, sp, candBase, &done, nodeVisits);
                    FG_PROF_END(2);
                    if (cn < 0) { done = false; break; }                      // the descent's stack would overflow
                    if (cn > 0) {
                        candTotal += (unsigned)cn;
                        batches++;
                        for (unsigned m = live; m; m &= m - 1u) {
                            const int j = __ffs((int)m) - 1;
                            const bool active = ((actBits & ~occBits) >> j) & 1u;
                            if (!__any_sync(full, active)) continue;
                            const float4 a = spv[j * FG_PV];
                            const float vx = a.x - px, vy = a.y - py, vz = a.z - pz;   // v12 = vpl.pos - x
                            if (test_batch<true>(sc, candBase, cn, active, v3(a.x, a.y, a.z), v3(-vx, -vy, -vz), shaftBase, tmin, tmax)) occBits |= 1u << j;
                        }
                        FG_PROF_END(3);
                    }
                    if (!done && ++batchNo >= fp.sharedBatches) break;   // a fat shaft: the occlusions found so far stand, the rest per VPL
                }
                if (!done) {
                    shared = false;
                    skipLen = skipLen ? min(fp.skipMax, skipLen * 2) : min(fp.skipMax, 1);
                    skipLeft = skipLen;
                    splits++;
                } else {
                    skipLen = 0;
                }
            }
            FG_PROF_END(3);
            if (!shared) {
                for (unsigned m = live; m; m &= m - 1u) {
                    const int j = __ffs((int)m) - 1;
                    const bool active = ((actBits & ~occBits) >> j) & 1u;
                    if (!__any_sync(full, active)) continue;
                    const float4 a = spv[j * FG_PV];
                    const float vx = a.x - px, vy = a.y - py, vz = a.z - pz;   // v12 = vpl.pos - x
                    const V3 org = v3(a.x, a.y, a.z), dir = v3(-vx, -vy, -vz);
                    descents++;
                    __syncwarp();
                    store_shaft(shaftBase, make_dshaft(org, org, ld_tile_lo(shaftBase), ld_tile_hi(shaftBase)));
                    st_shared_u32(stackBase, 0u);
                    uint32_t sp = stackBase + 4u;
                    __syncwarp();
                    bool done = false, occ = false, usePacket = false;
                    int batchNo = 0;
                    while (!done) {
                        const int cn = shaft_collect_batch(sc, shaftBase, tmin, tmax, stackBase, sp, candBase, &done, nodeVisits);
                        FG_PROF_END(4);
                        if (cn < 0) { usePacket = true; break; }                       // the descent's stack would overflow
                        if (cn > 0) {
                            candTotal += (unsigned)cn;
                            batches++;
                            occ |= test_batch<false>(sc, candBase, cn, active && !occ, org, dir, shaftBase, tmin, tmax);
                            FG_PROF_END(5);
                            if (!__any_sync(full, active && !occ)) break;                  // every ray has its answer
                        }
                        if (!done && ++batchNo >= fp.vplBatches) { usePacket = true; break; }   // far too many leaves: finish per ray
                    }
                    FG_PROF_END(4);
                    if (usePacket) {
                        packets++;
                        __syncwarp();
                        occ |= packet_any_hit(sc.nodes, sc.numNodes, sc.triLeaf, active && !occ, a.x, a.y, a.z, -vx, -vy, -vz, tmin, tmax, stackPtr, &ovf);
                        __syncwarp();
                        FG_PROF_END(6);
                    }
                    if (occ) occBits |= 1u << j;
                }
            }
            FG_PROF_END(4);
            // ---- shading of the unoccluded pairs.  The pixel's BRDF terms are re-derived per cluster (two L1-resident loads and
            // ~25 instructions per 16 VPLs) instead of occupying 11 registers during the visibility phase:
            // r1 = reflect(-wi10, n), so that dot(wi10, reflect(-wi12, n)) = dot(r1, wi12)
            const float4 g1 = ld_texel(gbuf + n + i), g2 = ld_texel(gbuf + 2 * n + i), g3 = ld_texel(gbuf + 3 * n + i);
            const float nx = g1.x, ny = g1.y, nz = g1.z;
            float r1x, r1y, r1z;
            {
                const float wx = gp.cameraPosition.x - px, wy = gp.cameraPosition.y - py, wz = gp.cameraPosition.z - pz;
                const float inv = rsqrtf(wx * wx + wy * wy + wz * wz);
                const float ux = wx * inv, uy = wy * inv, uz = wz * inv;
                const float nd = nx * ux + ny * uy + nz * uz;
                r1x = 2.0f * nd * nx - ux; r1y = 2.0f * nd * ny - uy; r1z = 2.0f * nd * nz - uz;
            }
            const float kd1x = g2.x * kInvPi, kd1y = g2.y * kInvPi, kd1z = g2.z * kInvPi;
            const float ks1x = g3.x, ks1y = g3.y, ks1z = g3.z, e1 = g3.w;
            const float cEval1 = (e1 + 2.0f) * 0.5f * kInvPi;
            const bool spec1 = ks1x > 0.f || ks1y > 0.f || ks1z > 0.f;
            for (unsigned m = live; m; m &= m - 1u) {
                const int j = __ffs((int)m) - 1;
                const float4 a = spv[j * FG_PV], b = spv[j * FG_PV + 1];
                const float vx = a.x - px, vy = a.y - py, vz = a.z - pz;   // v12 = vpl.pos - x
                const float c1 = fmaxf(nx * vx + ny * vy + nz * vz, 0.0f);
                const float c2 = fmaxf(-(b.x * vx + b.y * vy + b.z * vz), 0.0f);
                const float c1c2 = c1 * c2;
                const bool active = (actBits >> j) & 1u;
                rays += active ? 1u : 0u;
                steps++;
                if (active && !((occBits >> j) & 1u)) {
                    // shading tail of vplSplat (lighttracing.cu:296-345), FMA-contracted, with approximations of relative error < 1e-5
                    const float4 f2 = spv[j * FG_PV + 2], f3 = spv[j * FG_PV + 3], f4 = spv[j * FG_PV + 4], f5 = spv[j * FG_PV + 5];
                    const float dist2 = vx * vx + vy * vy + vz * vz;
                    const float inv = rsqrtf(dist2);
                    const float inv2 = inv * inv;
                    const float wx = vx * inv, wy = vy * inv, wz = vz * inv;             // wi12
                    const float g21 = c1c2 * inv2 * inv2;
                    // BRDF at the VPL: kd / pi + PhongEvalF(-wi12, fluxDir, n, e) ks
                    const float d2 = fmaxf(-(wx * f3.x + wy * f3.y + wz * f3.z), 0.0f);
                    const bool spec2 = f5.x > 0.f || f5.y > 0.f || f5.z > 0.f;        // warp-uniform
                    const float pow2 = (spec2 && d2 > 0.000001f) ? fast_pow(d2, a.w) : 0.0f;
                    const float s2 = f2.w * pow2;
                    const float b2x = fmaf(s2, f5.x, f4.x), b2y = fmaf(s2, f5.y, f4.y), b2z = fmaf(s2, f5.z, f4.z);
                    // BRDF at the pixel: kd / pi + PhongEvalF(wi10, wi12, n, e) ks
                    const float d1 = fmaxf(r1x * wx + r1y * wy + r1z * wz, 0.0f);
                    const float pow1 = (spec1 && d1 > 0.000001f) ? fast_pow(d1, e1) : 0.0f;
                    const float s1 = cEval1 * pow1;
                    const float b1x = fmaf(s1, ks1x, kd1x), b1y = fmaf(s1, ks1y, kd1y), b1z = fmaf(s1, ks1z, kd1z);
                    float cx = f2.x * b1x * b2x, cy = f2.y * b1y * b2y, cz = f2.z * b1z * b2z;   // flux * brdf1 * brdf2
                    if (MC == 0) {
                        cx *= g21; cy *= g21; cz *= g21;
                    } else if (MC == 1) {
                        // pdfDe = LambertPdfA pSel + PhongPdfA (1 - pSel), evaluated at the VPL towards x (lighttracing.cu:316-318):
                        // both share c1 c2 / d^4 and the cosine power already computed for BRDF 2
                        const float pdfDe = fmaf(g21, f4.w, f3.w * pow2 * (c1 * inv) * inv2);
                        float weight;
                        if (gp.misMode == 1) weight = __fdividef(gp.pdfMc, gp.pdfMc + pdfDe);
                        else if (gp.misMode == 2) weight = gp.pdfMc > pdfDe ? 1.0f : 0.0f;
                        else { const float a2 = gp.pdfMc * gp.pdfMc, d2e = pdfDe * pdfDe; weight = __fdividef(a2, a2 + d2e); }
                        const float wg = weight * g21;
                        cx *= wg; cy *= wg; cz *= wg;
                    } else if (MC == 2) {
                        const float gc = fminf(g21, gp.clampingValue);
                        cx *= gc; cy *= gc; cz *= gc;
                    } else {
                        cx = f2.x * fminf(g21 * b1x * b2x, gp.clampingValue);
                        cy = f2.y * fminf(g21 * b1y * b2y, gp.clampingValue);
                        cz = f2.z * fminf(g21 * b1z * b2z, gp.clampingValue);
                    }
                    resx += cx; resy += cy; resz += cz;
                }
            }
            FG_PROF_END(7);
        }
        }  // next depth group
        if (inside) {
            const long long q[3] = {to_fixed(resx * gp.invNumVpl), to_fixed(resy * gp.invNumVpl), to_fixed(resz * gp.invNumVpl)};
            if (fp.numChunks == 1) {
                for (int c = 0; c < 3; c++) acc[i * 3 + c] = gp.doAccumulate ? acc[i * 3 + c] + q[c] : q[c];
            } else {   // (the tile was cleared beforehand when doAccumulate == 0)
                for (int c = 0; c < 3; c++)
                    if (q[c]) atomicAdd(reinterpret_cast<unsigned long long*>(acc + i * 3 + c), (unsigned long long)q[c]);
            }
        }
        if (tileCost && lane == 0) {
            const long long dt = (clock64() - itemStart) >> 8;
            tileCost[v] = dt > 0xffffffffll ? 0xffffffffu : (uint32_t)dt;
        }
    }
    if (ovf) stats->stackOverflow = 1;
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(full, rays, o);
    if (lane == 0) {
        if (rays) atomicAdd(&stats->shadowRays, (unsigned long long)rays);
        if (steps) {
            atomicAdd(&stats->shaftSteps, (unsigned long long)steps); atomicAdd(&stats->shaftFallbacks, (unsigned long long)packets);
            atomicAdd(&stats->shaftNodeVisits, (unsigned long long)nodeVisits); atomicAdd(&stats->shaftCandLeaves, (unsigned long long)candTotal);
            atomicAdd(&stats->clusterDescents, (unsigned long long)descents); atomicAdd(&stats->clusterSplits, (unsigned long long)splits);
            atomicAdd(&stats->clusterHist[11], (unsigned long long)batches);
            FG_HIST(for (int k = 0; k < 12; k++) if (hist[k]) atomicAdd(&stats->clusterHist[k], (unsigned long long)hist[k]);)
#ifdef EVPLP_GATHER_PROF
            for (int k = 0; k < 10; k++) atomicAdd(&stats->clusterHist[k], (unsigned long long)prof[k]);
#endif
        }
    }
}

__global__ void clear_rect_kernel(long long* acc, int W, int x0, int y0, int x1, int y1) {
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x, y = y0 + blockIdx.y;
    if (x >= x1 || y >= y1) return;
    const size_t i = (size_t)y * W + x;
    acc[i * 3] = 0; acc[i * 3 + 1] = 0; acc[i * 3 + 2] = 0;
}

__global__ void iota_u32_kernel(uint32_t* out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

__global__ void debug_fast_pow_kernel(const float* x, const float* y, uint32_t n, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fast_pow(x[i], y[i]);
}

cudaError_t launch_debug_fast_pow(EvplpContext* c, const float* x, const float* y, uint32_t n, float* out) {
    if (n == 0) return cudaSuccess;
    debug_fast_pow_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(x, y, n, out);
    c->launches++;
    return cudaGetLastError();
}

#define FG_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

template <int MC>
static void launch_mc(EvplpContext* c, dim3 grid, const FastParams& fp, uint32_t* tileCounter, const uint32_t* tileOrder, uint32_t* tileCost) {
    const int mb = c->opt.gatherMinBlocks ? c->opt.gatherMinBlocks : 4;   // 64 registers, 32 warps / SM: the kernel is latency bound (measured: 4 > 3 > 2)
    if (mb >= 4)
        gather_cluster_kernel<MC, 4><<<grid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), fp, c->gbuf.p, c->vplPrepared.p, c->clusterBox.p, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost);
    else if (mb == 2)
        gather_cluster_kernel<MC, 2><<<grid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), fp, c->gbuf.p, c->vplPrepared.p, c->clusterBox.p, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost);
    else
        gather_cluster_kernel<MC, 3><<<grid, GATHER_WARPS * 32, 0, c->stream>>>(c->scene(), fp, c->gbuf.p, c->vplPrepared.p, c->clusterBox.p, c->accVpl.p, c->devStats.p, tileCounter, tileOrder, tileCost);
}

cudaError_t launch_gather_cluster(EvplpContext* c, EvplpTile t, GatherParams g, uint32_t count) {
    const EvplpParams& P = c->params;
    const int tw = t.x1 - t.x0, th = t.y1 - t.y0;
    if (tw <= 0 || th <= 0) return cudaSuccess;
    cudaStream_t st = c->stream;
    FastParams fp;
    fp.g = g;
    fp.count = count;
    fp.clusterSize = c->opt.clusterSize < 1 ? 1 : (c->opt.clusterSize > FG_CL_MAX ? FG_CL_MAX : c->opt.clusterSize);
    const uint32_t numRuns = (count + (uint32_t)fp.clusterSize - 1) / (uint32_t)fp.clusterSize;   // nominal clusters (before cuts)
    const TileShare share = tile_share(c, t);
    fp.stride = share.stride; fp.offset = share.offset; fp.tilesX = share.tilesX; fp.pitchX = share.pitchX; fp.ownedTiles = share.ownedTiles;
    fp.tileAngle = 8.0f * 2.0f * P.tanHalfFovX / (float)c->W;
    fp.sharedBatches = c->opt.sharedBatches; fp.vplBatches = c->opt.vplBatches; fp.skipMax = c->opt.clusterSkipMax;
    if (fp.ownedTiles == 0) return cudaSuccess;
    c->stats.gatherPairs += share.pixels * (uint64_t)count;
    if (count == 0) {
        // no usable VPL: the frame's contribution is zero (cleareveryframe still has to overwrite the layer)
        if (!P.doAccumulate) {
            dim3 cg((tw + 127) / 128, th);
            clear_rect_kernel<<<cg, 128, 0, st>>>(c->accVpl.p, c->W, t.x0, t.y0, t.x1, t.y1);
            c->launches++;
        }
        return cudaGetLastError();
    }
    // ---- Morton order of the usable VPLs, prepared VPLs, cluster boxes
    FG_CK(c->vplKeys.reserve(count)); FG_CK(c->vplKeysSorted.reserve(count)); FG_CK(c->vplOrder.reserve(count));
    FG_CK(c->vplVals.reserve(count));
    FG_CK(c->vplPrepared.reserve((size_t)count * FG_PV));
    FG_CK(c->clusterBox.reserve((size_t)numRuns * 4 * 2));
    FG_CK(c->clusterSlots.reserve((size_t)numRuns * 4)); FG_CK(c->clusterList.reserve((size_t)numRuns * 4 + 1));
    float3 smin = make_float3(c->sceneMin[0], c->sceneMin[1], c->sceneMin[2]), scale;
    scale.x = 1024.0f / fmaxf(c->sceneMax[0] - c->sceneMin[0], 1e-20f);
    scale.y = 1024.0f / fmaxf(c->sceneMax[1] - c->sceneMin[1], 1e-20f);
    scale.z = 1024.0f / fmaxf(c->sceneMax[2] - c->sceneMin[2], 1e-20f);
    const unsigned pb = (count + 255) / 256;
    vpl_code_kernel<<<pb, 256, 0, st>>>(c->records.p, c->vplList.p, count, smin, scale, c->vplKeys.p, c->vplVals.p);
    size_t tempBytes = 0;
    FG_CK(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, c->vplKeys.p, c->vplKeysSorted.p, c->vplVals.p, c->vplOrder.p, (int)count, 0, 30, st));
    FG_CK(c->sortTemp.reserve(tempBytes));
    FG_CK(cub::DeviceRadixSort::SortPairs(c->sortTemp.p, tempBytes, c->vplKeys.p, c->vplKeysSorted.p, c->vplVals.p, c->vplOrder.p, (int)count, 0, 30, st));
    vpl_prepare_kernel<<<pb, 256, 0, st>>>(c->records.p, c->vplOrder.p, count, c->vplPrepared.p);
    // cluster layout: runs of clusterSize, cut where the Morton order jumps (box edge > extent permille of the scene's longest edge)
    const float sceneEdge = fmaxf(c->sceneMax[0] - c->sceneMin[0], fmaxf(c->sceneMax[1] - c->sceneMin[1], c->sceneMax[2] - c->sceneMin[2]));
    const float maxExtent = c->opt.clusterExtentPermille > 0 ? sceneEdge * (float)c->opt.clusterExtentPermille * 0.001f : INFINITY;
    uint32_t* devNumClusters = c->clusterList.p + (size_t)numRuns * 4;
    cluster_layout_kernel<<<(numRuns + 127) / 128, 128, 0, st>>>(c->vplPrepared.p, count, fp.clusterSize, numRuns, maxExtent, c->clusterSlots.p);
    tempBytes = 0;
    FG_CK(cub::DeviceSelect::If(nullptr, tempBytes, c->clusterSlots.p, c->clusterList.p, devNumClusters, (int)(numRuns * 4), ClusterUsed(), st));
    FG_CK(c->sortTemp.reserve(tempBytes));
    FG_CK(cub::DeviceSelect::If(c->sortTemp.p, tempBytes, c->clusterSlots.p, c->clusterList.p, devNumClusters, (int)(numRuns * 4), ClusterUsed(), st));
    cluster_bounds_kernel<<<(numRuns * 4 + 127) / 128, 128, 0, st>>>(c->vplPrepared.p, c->clusterList.p, devNumClusters, c->clusterBox.p);
    fp.clusterList = c->clusterList.p; fp.numClusters = devNumClusters;
    c->launches += 8;
    // ---- work items: (owned tile, cluster range); enough of them per resident warp that the tail stays short
    const unsigned residentBlocks = 148u * 4u;
    const uint64_t residentWarps = (uint64_t)residentBlocks * GATHER_WARPS;
    unsigned chunks = 1;
    if (c->opt.gatherChunks > 1) {
        chunks = (unsigned)c->opt.gatherChunks;
    } else {
        const uint64_t want = residentWarps * 24u;
        if (fp.ownedTiles < want) chunks = (unsigned)((want + fp.ownedTiles - 1) / fp.ownedTiles);
        const unsigned maxChunks = (numRuns + 7) / 8;   // at least 8 (nominal) clusters per range
        if (chunks > maxChunks) chunks = maxChunks ? maxChunks : 1;
    }
    if (chunks > numRuns) chunks = numRuns;
    fp.numChunks = chunks;
    fp.g.numChunks = chunks;
    if (chunks > 1 && !P.doAccumulate) {
        dim3 cg((tw + 127) / 128, th);
        clear_rect_kernel<<<cg, 128, 0, st>>>(c->accVpl.p, c->W, t.x0, t.y0, t.x1, t.y1);
        c->launches++;
    }
    const uint32_t vTotal = fp.ownedTiles * chunks;
    uint32_t* tileCounter = c->counters.p + 2;
    FG_CK(cudaMemsetAsync(tileCounter, 0, sizeof(uint32_t), st));
    const uint32_t* tileOrder = nullptr;
    uint32_t* tileCost = nullptr;
    if (c->opt.gatherLpt) {
        // longest-processing-time-first: order the items by the cycles they took in the previous launch of the same item grid
        const uint64_t sig[4] = {((uint64_t)fp.ownedTiles << 32) | chunks, ((uint64_t)(uint32_t)t.x0 << 32) | (uint32_t)t.y0,
                                 ((uint64_t)(uint32_t)t.x1 << 32) | (uint32_t)t.y1, ((uint64_t)fp.stride << 32) | fp.offset | (1ull << 63)};
        const bool same = c->gatherCostValid && memcmp(sig, c->gatherSig, sizeof(sig)) == 0;
        if (!same) {
            FG_CK(c->gatherCost.reserve(vTotal)); FG_CK(c->gatherCostSorted.reserve(vTotal));
            FG_CK(c->gatherIota.reserve(vTotal)); FG_CK(c->gatherOrder.reserve(vTotal));
            iota_u32_kernel<<<(vTotal + 255) / 256, 256, 0, st>>>(c->gatherIota.p, vTotal);
            c->launches++;
            memcpy(c->gatherSig, sig, sizeof(sig));
            c->gatherCostValid = true;
        } else {
            tempBytes = 0;
            FG_CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tempBytes, c->gatherCost.p, c->gatherCostSorted.p, c->gatherIota.p, c->gatherOrder.p, (int)vTotal, 0, 32, st));
            FG_CK(c->sortTemp.reserve(tempBytes));
            FG_CK(cub::DeviceRadixSort::SortPairsDescending(c->sortTemp.p, tempBytes, c->gatherCost.p, c->gatherCostSorted.p, c->gatherIota.p, c->gatherOrder.p, (int)vTotal, 0, 32, st));
            c->launches += 2;
            tileOrder = c->gatherOrder.p;
        }
        tileCost = c->gatherCost.p;
    }
    const unsigned blocksWanted = (vTotal + GATHER_WARPS - 1) / GATHER_WARPS;
    dim3 grid(blocksWanted < residentBlocks ? blocksWanted : residentBlocks, 1, 1);
    c->stageBegin(ST_GATHER);
    const unsigned mode = P.misMode;
    if (mode == 0) launch_mc<0>(c, grid, fp, tileCounter, tileOrder, tileCost);
    else if (mode <= 3) launch_mc<1>(c, grid, fp, tileCounter, tileOrder, tileCost);
    else if (mode == 4) launch_mc<2>(c, grid, fp, tileCounter, tileOrder, tileCost);
    else launch_mc<3>(c, grid, fp, tileCounter, tileOrder, tileCost);
    c->stageEnd(ST_GATHER);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace evplp
