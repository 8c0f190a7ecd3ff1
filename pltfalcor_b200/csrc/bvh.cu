// bvh.cu — GPU BVH builder (replaces the driver BLAS/TLAS build, reference Source/Falcor/Scene/Scene.cpp:3092-3760)
// and the standalone ray-query kernels behind pltpt_trace_closest / pltpt_trace_any (replace TraceRay,
// pltpt_sample.rt.slang:174-183, ShadowRays.slangh:68-77,115-124).
//
// Build pipeline (all on the device; B200 has no RT cores so there is no driver structure to lean on):
//   1. k_world_tris    : instance transform -> world-space triangles with the contract's rounding, AABBs, scene bounds
//   2. k_morton        : 63-bit Morton code of the AABB centre; CUB radix sort (key, triangle)
//   3. k_radix_tree    : Karras 2012 binary radix tree over the sorted codes (one thread per internal node)
//   4. k_fit           : bottom-up AABB fit with per-node arrival counters
//   5. k_sah_costs     : bottom-up dynamic programme of Ylitie et al. 2017 (section 3.1) over the binary tree: C(n, i) = least SAH
//                        cost of representing the subtree of n by at most i children of a wide node (i = 1..7), with the choices
//                        "leaf of <= 3 triangles", "own wide node" (cost A_n c_node + best split of 8 slots) and the split of i
//                        slots between the two binary children recorded per node
//   6. k_collapse      : level-synchronous top-down pass that follows the recorded choices: every wide node takes the SAH-optimal
//                        set of <= 8 descendants of its binary node as children, puts each into the slot of the octant it lies in,
//                        and quantises the child boxes to 8 bits on a per-node power-of-two grid (80 B nodes) with conservative
//                        rounding verified in the kernel.  The level loop runs on the device (rotating queue counters); the host
//                        looks at the queue length once every 16 levels.
// Flat world-space BVH: static instances are pre-transformed exactly like Falcor does for static meshes
// (Scene.cpp:2712-2717), so one traversal handles BLAS+TLAS.
#include "bvh_traverse.cuh"
#include <cub/cub.cuh>
#include <cfloat>

namespace plt {
namespace {

constexpr uint32_t kMaxLeafTris = 3;
constexpr uint32_t kLeafBit = 0x80000000u;

// order-preserving float <-> uint mapping for atomicMin/Max
PLT_D uint32_t floatFlip(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
PLT_D float floatUnflip(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

PLT_D void fetchIndices(const DevScene& sc, uint32_t instanceID, uint32_t tri, uint32_t idx[3]) {  // Scene.slang getLocalIndices/getIndices
    const PltptGeometryInstance inst = sc.instances[instanceID];
    const PltptMeshDesc mesh = sc.meshes[inst.geometryID];
    if (mesh.indexCount == 0) { idx[0] = tri * 3 + inst.vbOffset; idx[1] = idx[0] + 1; idx[2] = idx[0] + 2; return; }
    const uint8_t* base = sc.indexData + size_t(inst.ibOffset) * 4;
    if (inst.flags & PLTPT_INSTANCE_FLAG_USE_16BIT_INDICES) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(base + size_t(tri) * 6);
        idx[0] = uint32_t(p[0]) + inst.vbOffset; idx[1] = uint32_t(p[1]) + inst.vbOffset; idx[2] = uint32_t(p[2]) + inst.vbOffset;
    } else {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(base + size_t(tri) * 12);
        idx[0] = p[0] + inst.vbOffset; idx[1] = p[1] + inst.vbOffset; idx[2] = p[2] + inst.vbOffset;
    }
}

// Intersection contract: p_w = ((m0*x + m1*y) + m2*z) + m3 per row, every operation rounded to fp32.
PLT_D float3 worldPoint(const float* __restrict__ M, const float* __restrict__ P) {
    float r[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        r[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[4 * i + 0], P[0]), __fmul_rn(M[4 * i + 1], P[1])), __fmul_rn(M[4 * i + 2], P[2])), M[4 * i + 3]);
    return mk3(r[0], r[1], r[2]);
}

__global__ void k_world_tris(DevScene sc, BvhTri* __restrict__ tris, float4* __restrict__ boxLo, float4* __restrict__ boxHi, uint32_t* __restrict__ sceneBounds) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sc.triCount) return;
    // instance of global triangle g: last i with instTriOffset[i] <= g
    uint32_t lo = 0, hi = sc.instanceCount;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sc.instTriOffset[mid] <= g) lo = mid; else hi = mid; }
    const uint32_t inst = lo, prim = g - sc.instTriOffset[inst];
    uint32_t idx[3];
    fetchIndices(sc, inst, prim, idx);
    const float* M = sc.world + 16 * size_t(sc.instances[inst].globalMatrixID);
    const float3 v0 = worldPoint(M, sc.vertices[idx[0]].position), v1 = worldPoint(M, sc.vertices[idx[1]].position), v2 = worldPoint(M, sc.vertices[idx[2]].position);
    BvhTri t;
    t.v0x = v0.x; t.v0y = v0.y; t.v0z = v0.z; t.instanceID = inst;
    t.v1x = v1.x; t.v1y = v1.y; t.v1z = v1.z; t.primitiveIndex = prim;
    t.v2x = v2.x; t.v2y = v2.y; t.v2z = v2.z; t.globalIndex = g;
    tris[g] = t;
    const float3 bl = vmin(v0, vmin(v1, v2)), bh = vmax(v0, vmax(v1, v2));
    boxLo[g] = mk4(bl.x, bl.y, bl.z, 0.f);
    boxHi[g] = mk4(bh.x, bh.y, bh.z, 0.f);
    atomicMin(sceneBounds + 0, floatFlip(bl.x)); atomicMin(sceneBounds + 1, floatFlip(bl.y)); atomicMin(sceneBounds + 2, floatFlip(bl.z));
    atomicMax(sceneBounds + 3, floatFlip(bh.x)); atomicMax(sceneBounds + 4, floatFlip(bh.y)); atomicMax(sceneBounds + 5, floatFlip(bh.z));
}

PLT_D uint64_t expandBits21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__global__ void k_morton(uint32_t n, const float4* __restrict__ boxLo, const float4* __restrict__ boxHi, const uint32_t* __restrict__ sceneBounds,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float3 slo = mk3(floatUnflip(sceneBounds[0]), floatUnflip(sceneBounds[1]), floatUnflip(sceneBounds[2]));
    const float3 shi = mk3(floatUnflip(sceneBounds[3]), floatUnflip(sceneBounds[4]), floatUnflip(sceneBounds[5]));
    const float4 l = boxLo[g], h = boxHi[g];
    const float3 c = mk3(.5f * (l.x + h.x), .5f * (l.y + h.y), .5f * (l.z + h.z));
    const float3 e = vmax(shi - slo, mk3(1e-30f));
    const float s = 2097151.f;
    const uint32_t qx = uint32_t(fminf(fmaxf((c.x - slo.x) / e.x, 0.f), 1.f) * s), qy = uint32_t(fminf(fmaxf((c.y - slo.y) / e.y, 0.f), 1.f) * s),
                   qz = uint32_t(fminf(fmaxf((c.z - slo.z) / e.z, 0.f), 1.f) * s);
    keys[g] = (expandBits21(qx) << 2) | (expandBits21(qy) << 1) | expandBits21(qz);
    vals[g] = g;
}

// ---- Karras 2012: "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" ----------------------
PLT_D int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(uint32_t(i) ^ uint32_t(j)) : __clzll((long long)(a ^ b));
}
// refs: internal node i -> i ; leaf (sorted position p) -> p | kLeafBit
__global__ void k_radix_tree(int n, const uint64_t* __restrict__ keys, uint32_t* __restrict__ left, uint32_t* __restrict__ right, uint32_t* __restrict__ first,
                             uint32_t* __restrict__ last, uint32_t* __restrict__ parentInner, uint32_t* __restrict__ parentLeaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t L = lo == gamma ? (uint32_t(gamma) | kLeafBit) : uint32_t(gamma);
    const uint32_t R = hi == gamma + 1 ? (uint32_t(gamma + 1) | kLeafBit) : uint32_t(gamma + 1);
    left[i] = L; right[i] = R; first[i] = uint32_t(lo); last[i] = uint32_t(hi);
    if (L & kLeafBit) parentLeaf[gamma] = uint32_t(i); else parentInner[gamma] = uint32_t(i);
    if (R & kLeafBit) parentLeaf[gamma + 1] = uint32_t(i); else parentInner[gamma + 1] = uint32_t(i);
    if (i == 0) parentInner[0] = 0xffffffffu;
}

// node boxes: index i (internal) in [0, n-1); leaves read the triangle boxes through `sorted`.
__global__ void k_fit(int n, const uint32_t* __restrict__ sorted, const float4* __restrict__ triLo, const float4* __restrict__ triHi, const uint32_t* __restrict__ left,
                      const uint32_t* __restrict__ right, const uint32_t* __restrict__ parentInner, const uint32_t* __restrict__ parentLeaf, float4* __restrict__ nodeLo,
                      float4* __restrict__ nodeHi, uint32_t* __restrict__ arrivals) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t node = parentLeaf[p];
    while (node != 0xffffffffu) {
        __threadfence();
        if (atomicAdd(arrivals + node, 1u) == 0u) return;  // the sibling subtree is not finished yet
        __threadfence();
        float4 lo[2], hi[2];
        const uint32_t c[2] = {left[node], right[node]};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (c[k] & kLeafBit) { const uint32_t t = sorted[c[k] & ~kLeafBit]; lo[k] = triLo[t]; hi[k] = triHi[t]; }
            else { lo[k] = __ldcg(nodeLo + c[k]); hi[k] = __ldcg(nodeHi + c[k]); }
        }
        __stcg(nodeLo + node, mk4(fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.f));
        __stcg(nodeHi + node, mk4(fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.f));
        node = parentInner[node];
    }
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) ---------------------------------------------------------------
// The binary tree is built bottom-up by agglomeration instead of top-down by Morton-code splits: the clusters (initially the triangles in
// Morton order) each look for the neighbour within kPlocRadius positions whose union with them has the smallest surface area; mutual
// nearest neighbours merge into a new node, the array is compacted, and the search repeats on the shorter array until one cluster is left.
// The merges follow the surface-area heuristic locally, where the radix tree only follows the bits of the codes.
#ifndef PLT_BVH_PLOC
#define PLT_BVH_PLOC 1  // 0: Karras radix tree (rounds 1-2a), kept for comparison (tools/variants.py: bvh_lbvh)
#endif
#ifndef PLT_PLOC_RADIUS
#define PLT_PLOC_RADIUS 8
#endif
constexpr int kPlocRadius = PLT_PLOC_RADIUS, kPlocThreads = 256;
PLT_D float unionArea(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    const float ex = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), ey = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), ez = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return ex * ey + ey * ez + ez * ex;
}
// first iteration input: clusters = leaves in Morton order
__global__ void k_ploc_init(uint32_t n, const uint32_t* __restrict__ sorted, const float4* __restrict__ triLo, const float4* __restrict__ triHi, uint32_t* __restrict__ cl,
                            float4* __restrict__ clLo, float4* __restrict__ clHi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = sorted[i];
    cl[i] = i | kLeafBit; clLo[i] = triLo[t]; clHi[i] = triHi[t];
}
// nearest neighbour of every cluster within the radius (shared-memory tile of the block's boxes plus a halo of kPlocRadius on both sides)
__global__ void __launch_bounds__(kPlocThreads) k_ploc_nearest(uint32_t c, const float4* __restrict__ clLo, const float4* __restrict__ clHi, uint32_t* __restrict__ nn) {
    __shared__ float4 sLo[kPlocThreads + 2 * kPlocRadius], sHi[kPlocThreads + 2 * kPlocRadius];
    const int base = int(blockIdx.x) * kPlocThreads - kPlocRadius;
    for (int k = threadIdx.x; k < kPlocThreads + 2 * kPlocRadius; k += kPlocThreads) {
        const int g = base + k;
        if (g >= 0 && g < int(c)) { sLo[k] = clLo[g]; sHi[k] = clHi[g]; }
    }
    __syncthreads();
    const int i = int(blockIdx.x) * kPlocThreads + int(threadIdx.x);
    if (i >= int(c)) return;
    const int li = int(threadIdx.x) + kPlocRadius;
    const float4 lo = sLo[li], hi = sHi[li];
    float best = FLT_MAX;
    int bestJ = -1;
#pragma unroll
    for (int d = -kPlocRadius; d <= kPlocRadius; ++d) {
        const int j = i + d;
        if (d == 0 || j < 0 || j >= int(c)) continue;
        const float a = unionArea(lo, hi, sLo[li + d], sHi[li + d]);
        if (a < best) { best = a; bestJ = j; }  // ties go to the lower index: the relation "mutual nearest" stays symmetric
    }
    nn[i] = uint32_t(bestJ);
}
// mutual nearest neighbours merge: the lower position keeps the new node, the upper one is dropped; everything else is kept
__global__ void k_ploc_merge(uint32_t c, const uint32_t* __restrict__ cl, const float4* __restrict__ clLo, const float4* __restrict__ clHi, const uint32_t* __restrict__ nn,
                             uint32_t* __restrict__ nodeCount, uint32_t* __restrict__ left, uint32_t* __restrict__ right, uint32_t* __restrict__ first, uint32_t* __restrict__ last,
                             uint32_t* __restrict__ parentInner, uint32_t* __restrict__ parentLeaf, float4* __restrict__ nodeLo, float4* __restrict__ nodeHi,
                             uint32_t* __restrict__ outRef, float4* __restrict__ outLo, float4* __restrict__ outHi, uint32_t* __restrict__ valid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const uint32_t j = nn[i];
    const bool mutual = j < c && nn[j] == i;
    if (mutual && i > j) { valid[i] = 0u; return; }
    valid[i] = 1u;
    if (!mutual) { outRef[i] = cl[i]; outLo[i] = clLo[i]; outHi[i] = clHi[i]; return; }
    const uint32_t a = cl[i], b = cl[j];
    const uint32_t id = atomicAdd(nodeCount, 1u);
    const float4 alo = clLo[i], ahi = clHi[i], blo = clLo[j], bhi = clHi[j];
    const float4 lo = mk4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f), hi = mk4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
    left[id] = a; right[id] = b;
    const uint32_t na = (a & kLeafBit) ? 1u : last[a] + 1u, nb = (b & kLeafBit) ? 1u : last[b] + 1u;
    first[id] = 0u; last[id] = na + nb - 1u;  // (first, last) only carry the triangle count of the subtree here: last - first + 1
    if (a & kLeafBit) parentLeaf[a & ~kLeafBit] = id; else parentInner[a] = id;
    if (b & kLeafBit) parentLeaf[b & ~kLeafBit] = id; else parentInner[b] = id;
    parentInner[id] = 0xffffffffu;  // until a later merge adopts this node: the last node created stays the root
    nodeLo[id] = lo; nodeHi[id] = hi;
    outRef[i] = id; outLo[i] = lo; outHi[i] = hi;
}
__global__ void k_ploc_compact(uint32_t c, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ offset, const uint32_t* __restrict__ ref, const float4* __restrict__ lo,
                               const float4* __restrict__ hi, uint32_t* __restrict__ cl, float4* __restrict__ clLo, float4* __restrict__ clHi, uint32_t* __restrict__ newCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    if (valid[i]) { const uint32_t o = offset[i]; cl[o] = ref[i]; clLo[o] = lo[i]; clHi[o] = hi[i]; }
    if (i == c - 1) *newCount = offset[i] + valid[i];
}

// ---- SAH-optimal collapse into wide nodes (Ylitie, Karras, Laine 2017, section 3.1) -------------------------------------------
constexpr float kCostNode = 1.f, kCostPrim = .3f;
// per binary internal node: cost[7] = C(n, 1..7); choice[8]: [0] = 1 when C(n,1) is the "own wide node" option (0 = leaf),
// [i-1] for i = 2..7 = slots given to the left child when i slots are split (0 = use C(n, i-1) instead), [7] = the same for 8 slots
struct SahTables { float* cost; uint8_t* choice; };
PLT_D float boxArea(float4 l, float4 h) { const float ex = h.x - l.x, ey = h.y - l.y, ez = h.z - l.z; return ex * ey + ey * ez + ez * ex; }
__global__ void k_sah_costs(int n, const uint32_t* __restrict__ sorted, const float4* __restrict__ triLo, const float4* __restrict__ triHi, const uint32_t* __restrict__ left,
                            const uint32_t* __restrict__ right, const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, const uint32_t* __restrict__ parentInner,
                            const uint32_t* __restrict__ parentLeaf, const float4* __restrict__ nodeLo, const float4* __restrict__ nodeHi, uint32_t* __restrict__ arrivals, SahTables sah) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t node = parentLeaf[p];
    while (node != 0xffffffffu) {
        __threadfence();
        if (atomicAdd(arrivals + node, 1u) == 0u) return;  // the sibling subtree is not finished yet
        __threadfence();
        float C[2][8];  // C[side][i], i = 1..7
        const uint32_t c[2] = {left[node], right[node]};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (c[k] & kLeafBit) {
                const uint32_t t = sorted[c[k] & ~kLeafBit];
                const float a = boxArea(triLo[t], triHi[t]) * kCostPrim;
                for (int i = 1; i <= 7; ++i) C[k][i] = a;
            } else
                for (int i = 1; i <= 7; ++i) C[k][i] = __ldcg(sah.cost + size_t(c[k]) * 7 + (i - 1));
        }
        const float A = boxArea(nodeLo[node], nodeHi[node]);
        const uint32_t N = last[node] - first[node] + 1u;
        float D[9];
        uint8_t Dk[9];
        for (int j = 2; j <= 8; ++j) {
            float best = FLT_MAX;
            int bk = 1;
            for (int k = (j - 7 > 1 ? j - 7 : 1); k <= j - 1 && k <= 7; ++k) {
                const float v = C[0][k] + C[1][j - k];
                if (v < best) { best = v; bk = k; }
            }
            D[j] = best; Dk[j] = uint8_t(bk);
        }
        float out[8];
        uint8_t ch[8];
        const float internal = A * kCostNode + D[8];
        const float leaf = N <= kMaxLeafTris ? A * float(N) * kCostPrim : FLT_MAX;
        out[1] = fminf(leaf, internal); ch[0] = internal < leaf ? 1 : 0; ch[7] = Dk[8];
        for (int i = 2; i <= 7; ++i) {
            if (D[i] < out[i - 1]) { out[i] = D[i]; ch[i - 1] = Dk[i]; }
            else { out[i] = out[i - 1]; ch[i - 1] = 0; }
        }
        for (int i = 1; i <= 7; ++i) __stcg(sah.cost + size_t(node) * 7 + (i - 1), out[i]);
        for (int i = 0; i < 8; ++i) sah.choice[size_t(node) * 8 + i] = ch[i];
        node = parentInner[node];
    }
}

struct CollapseItem { uint32_t bin, wide; };
struct TreeView {
    const uint32_t *left, *right, *first, *last, *sorted;
    const float4 *nodeLo, *nodeHi, *triLo, *triHi;
    PLT_D uint32_t count(uint32_t ref) const { return (ref & kLeafBit) ? 1u : last[ref] - first[ref] + 1u; }
    PLT_D uint32_t firstPos(uint32_t ref) const { return (ref & kLeafBit) ? (ref & ~kLeafBit) : first[ref]; }
    PLT_D void box(uint32_t ref, float3& lo, float3& hi) const {
        float4 l, h;
        if (ref & kLeafBit) { const uint32_t t = sorted[ref & ~kLeafBit]; l = triLo[t]; h = triHi[t]; }
        else { l = nodeLo[ref]; h = nodeHi[ref]; }
        lo = mk3(l.x, l.y, l.z); hi = mk3(h.x, h.y, h.z);
    }
    PLT_D float area(uint32_t ref) const {
        float3 lo, hi;
        box(ref, lo, hi);
        const float3 e = hi - lo;
        return e.x * e.y + e.y * e.z + e.z * e.x;
    }
};

// smallest biased exponent eb with origin + 255 * 2^(eb-127) >= hi (checked in fp32 exactly as the traversal dequantises)
PLT_D uint32_t gridExponent(float origin, float hi) {
    const float ext = hi - origin;
    int e = 0;
    if (ext > 0.f) frexpf(ext / 255.f, &e);  // ext/255 = m * 2^e, m in [0.5,1)  =>  2^e >= ext/255
    else e = -126;
    int eb = min(max(e + 127, 1), 254);
    while (eb < 254 && origin + 255.f * __uint_as_float(uint32_t(eb) << 23) < hi) ++eb;
    return uint32_t(eb);
}
PLT_D uint32_t quantLo(float origin, float s, float v) {
    int q = int(floorf((v - origin) / s));
    q = min(max(q, 0), 255);
    while (q > 0 && origin + float(q) * s > v) --q;
    return uint32_t(q);
}
PLT_D uint32_t quantHi(float origin, float s, float v) {
    int q = int(ceilf((v - origin) / s));
    q = min(max(q, 0), 255);
    while (q < 255 && origin + float(q) * s < v) ++q;
    return uint32_t(q);
}

// One level of the top-down pass: the items of this level are queue[level & 1][0 .. counts[level % 3]); inner children are appended to the
// other queue under counts[(level + 1) % 3]; counts[(level + 2) % 3] is re-zeroed for the level after that (nobody reads it meanwhile).
__global__ void __launch_bounds__(128) k_collapse(CollapseItem* __restrict__ queues, uint32_t queueStride, uint32_t level, uint32_t* __restrict__ levelCounts,
                                                  uint32_t* __restrict__ counters /*0 nodes,1 tris,2 max children seen*/, TreeView tv, SahTables sah, const BvhTri* __restrict__ srcTris,
                                                  BvhNode* __restrict__ outNodes, BvhTri* __restrict__ outTris, uint32_t* __restrict__ outParent, uint32_t rootIsLeaf) {
    const CollapseItem* items = queues + size_t(level & 1u) * queueStride;
    CollapseItem* nextItems = queues + size_t((level + 1u) & 1u) * queueStride;
    const uint32_t nItems = levelCounts[level % 3u];
    uint32_t* nextCount = levelCounts + (level + 1u) % 3u;
    if (blockIdx.x == 0 && threadIdx.x == 0) levelCounts[(level + 2u) % 3u] = 0;
  for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < nItems; it += gridDim.x * blockDim.x) {
    const CollapseItem item = items[it];
    // the SAH-optimal children of this wide node: follow the recorded slot splits down from the binary node
    uint32_t c[8];
    bool asLeaf[8];
    int cnt = 0;
    if (rootIsLeaf) { c[0] = item.bin; asLeaf[0] = true; cnt = 1; }
    else {
        uint32_t stRef[16];
        uint8_t stSlots[16];
        int sp = 0;
        const uint32_t k8 = sah.choice[size_t(item.bin) * 8 + 7];
        stRef[sp] = tv.right[item.bin]; stSlots[sp++] = uint8_t(8u - k8);
        stRef[sp] = tv.left[item.bin]; stSlots[sp++] = uint8_t(k8);
        while (sp > 0) {
            const uint32_t ref = stRef[--sp];
            uint32_t slots = stSlots[sp];
            if (ref & kLeafBit) { c[cnt] = ref; asLeaf[cnt++] = true; continue; }
            const uint8_t* ch = sah.choice + size_t(ref) * 8;
            while (slots > 1 && ch[slots - 1] == 0) --slots;  // C(n, i) = C(n, i-1): fewer slots do as well
            if (slots == 1) { c[cnt] = ref; asLeaf[cnt++] = ch[0] == 0; continue; }
            const uint32_t k = ch[slots - 1];
            stRef[sp] = tv.right[ref]; stSlots[sp++] = uint8_t(slots - k);
            stRef[sp] = tv.left[ref]; stSlots[sp++] = uint8_t(k);
        }
    }
    float3 nlo, nhi;
    tv.box(item.bin, nlo, nhi);
    for (int i = cnt; i < 8; ++i) c[i] = 0xffffffffu;  // empty slots
    constexpr uint32_t kEmpty = 0xffffffffu;
    uint32_t nInner = 0, nTris = 0;
    for (int i = 0; i < 8; ++i) {
        if (c[i] == kEmpty) continue;
        if (!asLeaf[i]) ++nInner; else nTris += tv.count(c[i]);
    }
    const uint32_t childBase = nInner ? atomicAdd(counters + 0, nInner) : 0u;
    const uint32_t triBase = nTris ? atomicAdd(counters + 1, nTris) : 0u;
    const uint32_t nextBase = nInner ? atomicAdd(nextCount, nInner) : 0u;

    BvhNode node;
    node.px = nlo.x; node.py = nlo.y; node.pz = nlo.z;
    const uint32_t ex = gridExponent(nlo.x, nhi.x), ey = gridExponent(nlo.y, nhi.y), ez = gridExponent(nlo.z, nhi.z);
    node.ex = uint8_t(ex); node.ey = uint8_t(ey); node.ez = uint8_t(ez);
    const float sx = __uint_as_float(ex << 23), sy = __uint_as_float(ey << 23), sz = __uint_as_float(ez << 23);
    node.childBase = childBase; node.triBase = triBase;
    uint32_t imask = 0, innerSeen = 0, triOff = 0;
    for (int i = 0; i < 8; ++i) {
        node.meta[i] = 0;
        node.qlox[i] = node.qloy[i] = node.qloz[i] = 0;
        node.qhix[i] = node.qhiy[i] = node.qhiz[i] = 0;
        if (c[i] == kEmpty) continue;
        float3 lo, hi;
        tv.box(c[i], lo, hi);
        node.qlox[i] = uint8_t(quantLo(nlo.x, sx, lo.x)); node.qhix[i] = uint8_t(quantHi(nlo.x, sx, hi.x));
        node.qloy[i] = uint8_t(quantLo(nlo.y, sy, lo.y)); node.qhiy[i] = uint8_t(quantHi(nlo.y, sy, hi.y));
        node.qloz[i] = uint8_t(quantLo(nlo.z, sz, lo.z)); node.qhiz[i] = uint8_t(quantHi(nlo.z, sz, hi.z));
        const uint32_t k = tv.count(c[i]);
        if (!asLeaf[i]) {
            imask |= 1u << i;
            node.meta[i] = uint8_t(0x20u | (24u + uint32_t(i)));
            nextItems[nextBase + innerSeen] = CollapseItem{c[i], childBase + innerSeen};
            outParent[childBase + innerSeen] = item.wide;  // for the bottom-up refit (k_refit_nodes)
            ++innerSeen;
        } else {
            node.meta[i] = uint8_t((k << 5) | triOff);
            // the <= 3 triangles under this binary node, by a walk over its subtree (a PLOC subtree is not a contiguous range of the Morton order)
            uint32_t st[4];
            int sp = 0;
            uint32_t emitted = 0;
            st[sp++] = c[i];
            while (sp > 0) {
                const uint32_t r = st[--sp];
                if (r & kLeafBit) { outTris[triBase + triOff + emitted] = srcTris[tv.sorted[r & ~kLeafBit]]; ++emitted; }
                else if (sp + 2 <= 4) { st[sp++] = tv.right[r]; st[sp++] = tv.left[r]; }
            }
            triOff += k;
        }
    }
    node.imask = uint8_t(imask);
    outNodes[item.wide] = node;
  }
}

// ---- refit (pltpt_update_transforms in PLTPT_BVH_UPDATE_REFIT mode; the reference's UpdateMode::Refit, Scene.cpp:3642-3647,3720) ----
// The tree keeps its topology: every leaf triangle is re-derived from its (instance, primitive) with the new matrices, then the node
// boxes are re-fitted bottom-up -- a thread starts at every node without inner children and climbs while it is the last of a node's
// inner children to arrive -- and the child boxes re-quantised on the node's new grid with the builder's conservative rounding.
__global__ void k_refit_tris(DevScene sc, BvhTri* __restrict__ tris) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sc.triCount) return;
    const uint32_t inst = tris[g].instanceID, prim = tris[g].primitiveIndex;
    uint32_t idx[3];
    fetchIndices(sc, inst, prim, idx);
    const float* M = sc.world + 16 * size_t(sc.instances[inst].globalMatrixID);
    const float3 v0 = worldPoint(M, sc.vertices[idx[0]].position), v1 = worldPoint(M, sc.vertices[idx[1]].position), v2 = worldPoint(M, sc.vertices[idx[2]].position);
    BvhTri& t = tris[g];
    t.v0x = v0.x; t.v0y = v0.y; t.v0z = v0.z;
    t.v1x = v1.x; t.v1y = v1.y; t.v1z = v1.z;
    t.v2x = v2.x; t.v2y = v2.y; t.v2z = v2.z;
}
__global__ void __launch_bounds__(128) k_refit_nodes(uint32_t nNodes, BvhNode* __restrict__ nodes, const BvhTri* __restrict__ tris, const uint32_t* __restrict__ parent,
                                                     uint32_t* __restrict__ arrivals, float4* boxLo, float4* boxHi) {
    uint32_t ni = blockIdx.x * blockDim.x + threadIdx.x;
    if (ni >= nNodes || nodes[ni].imask != 0) return;
    for (;;) {
        BvhNode node = nodes[ni];
        const uint32_t imask = node.imask;
        float3 clo[8], chi[8];
        float3 nlo = mk3(FLT_MAX, FLT_MAX, FLT_MAX), nhi = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        uint32_t used = 0;
        for (int i = 0; i < 8; ++i) {
            const bool inner = (imask >> i) & 1u;
            const uint32_t meta = node.meta[i];
            if (!inner && meta == 0) continue;
            used |= 1u << i;
            if (inner) {
                const uint32_t ci = node.childBase + __popc(imask & ((1u << i) - 1u));
                const float4 l = __ldcg(boxLo + ci), h = __ldcg(boxHi + ci);  // written by another thread of this launch: read through L2
                clo[i] = mk3(l.x, l.y, l.z); chi[i] = mk3(h.x, h.y, h.z);
            } else {
                float3 lo = mk3(FLT_MAX, FLT_MAX, FLT_MAX), hi = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
                const uint32_t cnt = meta >> 5, off = meta & 31u;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const BvhTri t = tris[node.triBase + off + k];
                    const float3 v0 = mk3(t.v0x, t.v0y, t.v0z), v1 = mk3(t.v1x, t.v1y, t.v1z), v2 = mk3(t.v2x, t.v2y, t.v2z);
                    lo = vmin(lo, vmin(v0, vmin(v1, v2))); hi = vmax(hi, vmax(v0, vmax(v1, v2)));
                }
                clo[i] = lo; chi[i] = hi;
            }
            nlo = vmin(nlo, clo[i]); nhi = vmax(nhi, chi[i]);
        }
        node.px = nlo.x; node.py = nlo.y; node.pz = nlo.z;
        const uint32_t ex = gridExponent(nlo.x, nhi.x), ey = gridExponent(nlo.y, nhi.y), ez = gridExponent(nlo.z, nhi.z);
        node.ex = uint8_t(ex); node.ey = uint8_t(ey); node.ez = uint8_t(ez);
        const float sx = __uint_as_float(ex << 23), sy = __uint_as_float(ey << 23), sz = __uint_as_float(ez << 23);
        for (int i = 0; i < 8; ++i) {
            if (!((used >> i) & 1u)) continue;
            node.qlox[i] = uint8_t(quantLo(nlo.x, sx, clo[i].x)); node.qhix[i] = uint8_t(quantHi(nlo.x, sx, chi[i].x));
            node.qloy[i] = uint8_t(quantLo(nlo.y, sy, clo[i].y)); node.qhiy[i] = uint8_t(quantHi(nlo.y, sy, chi[i].y));
            node.qloz[i] = uint8_t(quantLo(nlo.z, sz, clo[i].z)); node.qhiz[i] = uint8_t(quantHi(nlo.z, sz, chi[i].z));
        }
        nodes[ni] = node;
        if (ni == 0) return;
        __stcg(boxLo + ni, mk4(nlo.x, nlo.y, nlo.z, 0.f));
        __stcg(boxHi + ni, mk4(nhi.x, nhi.y, nhi.z, 0.f));
        __threadfence();
        const uint32_t p = parent[ni];
        const uint32_t need = uint32_t(__popc(uint32_t(nodes[p].imask)));
        if (atomicAdd(arrivals + p, 1u) + 1u < need) return;
        __threadfence();
        ni = p;
    }
}

template <bool ANY, bool kNearest>
__global__ void __launch_bounds__(128) k_trace(const BvhNode* __restrict__ nodes, const BvhTri* __restrict__ tris, uint32_t n, const float* __restrict__ rays, uint32_t* __restrict__ out,
                                               float* __restrict__ outT, uint32_t* __restrict__ overflow) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(rays) + 2 * size_t(i)), b = __ldg(reinterpret_cast<const float4*>(rays) + 2 * size_t(i) + 1);
        Hit h;
        const bool hit = traverseBvh<ANY, kNearest>(nodes, tris, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w, h, overflow);
        if (ANY) out[i] = hit ? 1u : 0u;
        else {
            reinterpret_cast<uint4*>(out)[i] = hit ? make_uint4(h.instanceID, h.primitiveIndex, __float_as_uint(h.u), __float_as_uint(h.v)) : make_uint4(0xffffffffu, 0u, 0u, 0u);
            if (outT) outT[i] = hit ? h.t : 0.f;
        }
    }
}

// The same query through the traversal driver the rendering kernels of non-trivial scenes use (travPersistent: persistent lanes
// that refill from a cursor, processNode per step), so that the bit-exact hit-ID checks cover that code path too.
template <bool ANY>
__global__ void __launch_bounds__(128) k_trace_lanes(const BvhNode* __restrict__ nodes, const BvhTri* __restrict__ tris, uint32_t n, const float* __restrict__ rays, uint32_t* __restrict__ out,
                                                     float* __restrict__ outT, uint32_t* __restrict__ cursor, uint32_t* __restrict__ overflow) {
    travPersistent<ANY>(nodes, tris, n, cursor, overflow, nullptr,
        [&](uint32_t i, float3& o, float3& d, float& tmin, float& tmax) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(rays) + 2 * size_t(i)), b = __ldg(reinterpret_cast<const float4*>(rays) + 2 * size_t(i) + 1);
            o = mk3(a.x, a.y, a.z); d = mk3(b.x, b.y, b.z); tmin = a.w; tmax = b.w;
            return true;
        },
        [&](uint32_t i, bool hit, const Hit& h) {
            if (ANY) out[i] = hit ? 1u : 0u;
            else {
                reinterpret_cast<uint4*>(out)[i] = hit ? make_uint4(h.instanceID, h.primitiveIndex, __float_as_uint(h.u), __float_as_uint(h.v)) : make_uint4(0xffffffffu, 0u, 0u, 0u);
                if (outT) outT[i] = hit ? h.t : 0.f;
            }
        });
}

inline uint32_t blocksFor(uint64_t n, uint32_t threads) { return uint32_t((n + threads - 1) / threads); }

}  // namespace

int buildBvh(PltptContext* c, cudaStream_t s) {
    DevScene& sc = c->scene;
    const uint32_t n = sc.triCount;
    cudaEventRecord(c->evA, s);
    DevBuf& srcTris = c->bvhScratch[0];
    DevBuf& triLo = c->bvhScratch[1];
    DevBuf& triHi = c->bvhScratch[2];
    DevBuf& keys = c->bvhScratch[3];   // keys in | keys out | vals in | vals out
    DevBuf& tree = c->bvhScratch[4];   // left,right,first,last,parentInner,arrivals (n each) + parentLeaf (n)
    DevBuf& nodeBox = c->bvhScratch[5];
    DevBuf& queue = c->bvhScratch[6];
    DevBuf& misc = c->bvhScratch[7];   // scene bounds (6) + counters (3) + overflow (1)
    PLT_CUDA(srcTris.ensure(size_t(n) * sizeof(BvhTri)));
    PLT_CUDA(triLo.ensure(size_t(n) * 16));
    PLT_CUDA(triHi.ensure(size_t(n) * 16));
    PLT_CUDA(keys.ensure(size_t(n) * (8 + 8 + 4 + 4)));
    PLT_CUDA(tree.ensure(size_t(n) * 7 * 4));
    PLT_CUDA(nodeBox.ensure(size_t(n) * 32));
    PLT_CUDA(queue.ensure(size_t(n) * 2 * sizeof(CollapseItem) + 16));
    PLT_CUDA(misc.ensure(64));
    PLT_CUDA(c->bvhTris.ensure(size_t(n) * sizeof(BvhTri)));
    // a wide node has >= 2 children unless it is the root of a <=3-triangle scene, so #nodes <= max(1, n - 1)
    PLT_CUDA(c->bvhNodes.ensure(size_t(n) * sizeof(BvhNode)));
    PLT_CUDA(c->bvhParent.ensure(size_t(n) * 4));

    uint32_t* miscU = misc.as<uint32_t>();
    const uint32_t initMisc[16] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, /*counters*/ 1u, 0u, 0u, /*overflow*/ 0u, /*query cursor*/ 0u, /*level counts*/ 0, 0, 0, 0, 0};
    PLT_CUDA(cudaMemcpyAsync(miscU, initMisc, sizeof(initMisc), cudaMemcpyHostToDevice, s));
    uint32_t* counters = miscU + 6;

    k_world_tris<<<blocksFor(n, 256), 256, 0, s>>>(sc, srcTris.as<BvhTri>(), triLo.as<float4>(), triHi.as<float4>(), miscU);
    uint64_t* keysIn = keys.as<uint64_t>();
    uint64_t* keysOut = keysIn + n;
    uint32_t* valsIn = reinterpret_cast<uint32_t*>(keysOut + n);
    uint32_t* valsOut = valsIn + n;
    k_morton<<<blocksFor(n, 256), 256, 0, s>>>(n, triLo.as<float4>(), triHi.as<float4>(), miscU, keysIn, valsIn);
    size_t tempBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, keysIn, keysOut, valsIn, valsOut, int(n), 0, 63, s);
    PLT_CUDA(c->lanes[0].work[24].ensure(tempBytes));
    PLT_CUDA(cub::DeviceRadixSort::SortPairs(c->lanes[0].work[24].p, tempBytes, keysIn, keysOut, valsIn, valsOut, int(n), 0, 63, s));
    c->stats.kernelLaunches += 2;

    uint32_t* left = tree.as<uint32_t>();
    uint32_t* right = left + n;
    uint32_t* first = right + n;
    uint32_t* last = first + n;
    uint32_t* parentInner = last + n;
    uint32_t* arrivals = parentInner + n;
    uint32_t* parentLeaf = arrivals + n;
    float4* nodeLo = nodeBox.as<float4>();
    float4* nodeHi = nodeLo + n;
    CollapseItem* q0 = queue.as<CollapseItem>();  // two halves of n items: levels alternate between them
    TreeView tv{left, right, first, last, valsOut, nodeLo, nodeHi, triLo.as<float4>(), triHi.as<float4>()};

    uint32_t levels = 0;
    // queue: two halves of n items each; level counts rotate through three words behind the build counters
    uint32_t* levelCounts = miscU + 11;
    const uint32_t levelInit[3] = {1u, 0u, 0u};
    PLT_CUDA(cudaMemcpyAsync(levelCounts, levelInit, sizeof(levelInit), cudaMemcpyHostToDevice, s));
    SahTables sah{nullptr, nullptr};
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    if (n == 1) {
        // no internal binary node exists: the root wide node gets the single triangle as its only leaf slot
        const CollapseItem root{0u | kLeafBit, 0u};
        PLT_CUDA(cudaMemcpyAsync(q0, &root, sizeof(root), cudaMemcpyHostToDevice, s));
        k_collapse<<<1, 128, 0, s>>>(q0, n, 0u, levelCounts, counters, tv, sah, srcTris.as<BvhTri>(), c->bvhNodes.as<BvhNode>(), c->bvhTris.as<BvhTri>(), c->bvhParent.as<uint32_t>(), 1u);
        c->stats.kernelLaunches++;
        levels = 1;
    } else {
        uint32_t rootId = 0;
#if PLT_BVH_PLOC
        {
            // cluster arrays (main + the merge pass's output before compaction), nearest-neighbour / keep / offset words, scan scratch
            DevBuf& ploc = c->lanes[0].work[39];
            size_t scanBytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), int(n), s);
            scanBytes = (scanBytes + 255) & ~size_t(255);
            PLT_CUDA(ploc.ensure(size_t(n) * (2 * 36 + 12) + scanBytes + 256));
            float4* clLo = ploc.as<float4>();
            float4* clHi = clLo + n;
            float4* tmpLo = clHi + n;
            float4* tmpHi = tmpLo + n;
            uint32_t* cl = reinterpret_cast<uint32_t*>(tmpHi + n);
            uint32_t* tmpRef = cl + n;
            uint32_t* nn = tmpRef + n;
            uint32_t* valid = nn + n;
            uint32_t* offset = valid + n;
            uint32_t* plocWords = offset + n;  // [0] nodes created, [1] clusters after the compaction
            void* scanTemp = reinterpret_cast<uint8_t*>(plocWords) + 256;
            PLT_CUDA(cudaMemsetAsync(plocWords, 0, 8, s));
            k_ploc_init<<<blocksFor(n, 256), 256, 0, s>>>(n, valsOut, triLo.as<float4>(), triHi.as<float4>(), cl, clLo, clHi);
            uint32_t clusters = n, iterations = 0;
            while (clusters > 1) {
                k_ploc_nearest<<<blocksFor(clusters, kPlocThreads), kPlocThreads, 0, s>>>(clusters, clLo, clHi, nn);
                k_ploc_merge<<<blocksFor(clusters, 256), 256, 0, s>>>(clusters, cl, clLo, clHi, nn, plocWords, left, right, first, last, parentInner, parentLeaf, nodeLo, nodeHi, tmpRef, tmpLo, tmpHi, valid);
                PLT_CUDA(cub::DeviceScan::ExclusiveSum(scanTemp, scanBytes, valid, offset, int(clusters), s));
                k_ploc_compact<<<blocksFor(clusters, 256), 256, 0, s>>>(clusters, valid, offset, tmpRef, tmpLo, tmpHi, cl, clLo, clHi, plocWords + 1);
                uint32_t next = 0;
                PLT_CUDA(cudaMemcpyAsync(&next, plocWords + 1, 4, cudaMemcpyDeviceToHost, s));
                PLT_CUDA(cudaStreamSynchronize(s));
                c->stats.kernelLaunches += 4;
                if (next == 0 || next >= clusters || ++iterations > 4096) { setError("PLOC did not converge: " + std::to_string(clusters) + " -> " + std::to_string(next) + " clusters"); return PLTPT_ERR_CUDA; }
                clusters = next;
            }
            PLT_CUDA(cudaMemcpyAsync(&rootId, cl, 4, cudaMemcpyDeviceToHost, s));
            PLT_CUDA(cudaStreamSynchronize(s));
            if (rootId & kLeafBit) { setError("PLOC root is a leaf"); return PLTPT_ERR_CUDA; }
            ploc.release();
        }
#else
        k_radix_tree<<<blocksFor(n - 1, 256), 256, 0, s>>>(int(n), keysOut, left, right, first, last, parentInner, parentLeaf);
        PLT_CUDA(cudaMemsetAsync(arrivals, 0, size_t(n) * 4, s));
        k_fit<<<blocksFor(n, 256), 256, 0, s>>>(int(n), valsOut, triLo.as<float4>(), triHi.as<float4>(), left, right, parentInner, parentLeaf, nodeLo, nodeHi, arrivals);
#endif
        // SAH tables of the dynamic programme (7 costs + 8 choices per binary node), bottom-up like the box fit
        DevBuf& sahBuf = c->lanes[0].work[25];
        PLT_CUDA(sahBuf.ensure(size_t(n) * (7 * 4 + 8)));
        sah.cost = sahBuf.as<float>();
        sah.choice = reinterpret_cast<uint8_t*>(sah.cost + size_t(n) * 7);
        PLT_CUDA(cudaMemsetAsync(arrivals, 0, size_t(n) * 4, s));
        k_sah_costs<<<blocksFor(n, 256), 256, 0, s>>>(int(n), valsOut, triLo.as<float4>(), triHi.as<float4>(), left, right, first, last, parentInner, parentLeaf, nodeLo, nodeHi, arrivals, sah);
        c->stats.kernelLaunches += 6;
        const CollapseItem root{rootId, 0u};
        PLT_CUDA(cudaMemcpyAsync(q0, &root, sizeof(root), cudaMemcpyHostToDevice, s));
        // top-down: one launch per level over a persistent grid; the host reads the queue length once per 16 levels only
        const uint32_t grid = uint32_t(sms) * 8;
        for (bool more = true; more;) {
            for (int k = 0; k < 16; ++k, ++levels)
                k_collapse<<<grid, 128, 0, s>>>(q0, n, levels, levelCounts, counters, tv, sah, srcTris.as<BvhTri>(), c->bvhNodes.as<BvhNode>(), c->bvhTris.as<BvhTri>(), c->bvhParent.as<uint32_t>(), 0u);
            c->stats.kernelLaunches += 16;
            uint32_t pending = 0;
            PLT_CUDA(cudaMemcpyAsync(&pending, levelCounts + levels % 3u, 4, cudaMemcpyDeviceToHost, s));
            PLT_CUDA(cudaStreamSynchronize(s));
            more = pending > 0;
            if (levels > 4096) { setError("BVH collapse did not terminate"); return PLTPT_ERR_CUDA; }
        }
        float rootCost = 0.f;
        float4 rl, rh;
        PLT_CUDA(cudaMemcpyAsync(&rootCost, sah.cost + size_t(rootId) * 7, 4, cudaMemcpyDeviceToHost, s));
        PLT_CUDA(cudaMemcpyAsync(&rl, nodeLo + rootId, 16, cudaMemcpyDeviceToHost, s));
        PLT_CUDA(cudaMemcpyAsync(&rh, nodeHi + rootId, 16, cudaMemcpyDeviceToHost, s));
        PLT_CUDA(cudaStreamSynchronize(s));
        const float ra = (rh.x - rl.x) * (rh.y - rl.y) + (rh.y - rl.y) * (rh.z - rl.z) + (rh.z - rl.z) * (rh.x - rl.x);
        c->stats.bvhSahCost = ra > 0.f ? double(rootCost) / double(ra) : 0.0;
        sahBuf.release();
    }
    uint32_t finalCounters[3] = {0, 0, 0};
    PLT_CUDA(cudaMemcpyAsync(finalCounters, counters, 12, cudaMemcpyDeviceToHost, s));
    cudaEventRecord(c->evB, s);
    PLT_CUDA(cudaStreamSynchronize(s));
    PLT_CUDA(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->evA, c->evB);
    if (finalCounters[1] != n) { setError("BVH build lost triangles: " + std::to_string(finalCounters[1]) + " of " + std::to_string(n)); return PLTPT_ERR_CUDA; }
    c->stats.bvhBuildMs = ms;
    c->stats.bvhNodes = finalCounters[0];
    c->stats.bvhLeafTriangles = finalCounters[1];
    c->stats.triangleCount = n;
    sc.bvhNodes = c->bvhNodes.as<BvhNode>();
    sc.bvhTris = c->bvhTris.as<BvhTri>();
    // build scratch is only needed again on the next set_scene
    for (int i = 0; i < 7; ++i) c->bvhScratch[i].release();
    return PLTPT_OK;
}

// Refit of the BVH built by buildBvh after the world matrices changed (same topology, new boxes).
int refitBvh(PltptContext* c, cudaStream_t s) {
    DevScene& sc = c->scene;
    const uint32_t n = sc.triCount, nNodes = uint32_t(c->stats.bvhNodes);
    if (!sc.bvhNodes || !c->bvhParent.p || nNodes == 0) { setError("refit without a built BVH"); return PLTPT_ERR_NO_SCENE; }
    DevBuf& scratch = c->bvhScratch[0];  // arrivals (4 B) + two float4 boxes per node
    PLT_CUDA(scratch.ensure(size_t(nNodes) * (4 + 32) + 32));
    float4* boxLo = scratch.as<float4>();
    float4* boxHi = boxLo + nNodes;
    uint32_t* arrivals = reinterpret_cast<uint32_t*>(boxHi + nNodes);
    cudaEventRecord(c->evA, s);
    PLT_CUDA(cudaMemsetAsync(arrivals, 0, size_t(nNodes) * 4, s));
    k_refit_tris<<<blocksFor(n, 256), 256, 0, s>>>(sc, c->bvhTris.as<BvhTri>());
    k_refit_nodes<<<blocksFor(nNodes, 128), 128, 0, s>>>(nNodes, c->bvhNodes.as<BvhNode>(), c->bvhTris.as<BvhTri>(), c->bvhParent.as<uint32_t>(), arrivals, boxLo, boxHi);
    cudaEventRecord(c->evB, s);
    PLT_CUDA(cudaStreamSynchronize(s));
    PLT_CUDA(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->evA, c->evB);
    c->stats.bvhRefitMs = ms;
    c->stats.kernelLaunches += 2;
    scratch.release();
    return PLTPT_OK;
}

static int traceLaunch(PltptContext* c, bool any, uint32_t n, const float* rays, uint32_t* out, float* outT, cudaStream_t s) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const uint32_t grid = std::min<uint32_t>(blocksFor(n, 128), uint32_t(sms) * 16);
    uint32_t* overflow = c->bvhScratch[7].as<uint32_t>() + 9;
    // PLTPT_TRACE_AUTO follows the renderer (launchTrace / launchShadowTrace): persistent lanes unless the tree has only a few nodes
    const bool lanes = c->traceMode == PLTPT_TRACE_PERSISTENT_LANES || (c->traceMode == PLTPT_TRACE_AUTO && c->stats.bvhNodes >= 256);
    if (lanes) {
        uint32_t* cursor = c->bvhScratch[7].as<uint32_t>() + 10;
        PLT_CUDA(cudaMemsetAsync(cursor, 0, 4, s));
        const uint32_t lgrid = std::min<uint32_t>(blocksFor(n, 128), uint32_t(sms) * 8);
        if (any) k_trace_lanes<true><<<lgrid, 128, 0, s>>>(c->scene.bvhNodes, c->scene.bvhTris, n, rays, out, nullptr, cursor, overflow);
        else k_trace_lanes<false><<<lgrid, 128, 0, s>>>(c->scene.bvhNodes, c->scene.bvhTris, n, rays, out, outT, cursor, overflow);
    } else if (c->traceMode == PLTPT_TRACE_ONE_RAY_PER_THREAD_NEAREST_FIRST) {  // what traceClosestHit() / occluded() inside the connect and manifold kernels run
        if (any) k_trace<true, true><<<grid, 128, 0, s>>>(c->scene.bvhNodes, c->scene.bvhTris, n, rays, out, nullptr, overflow);
        else k_trace<false, true><<<grid, 128, 0, s>>>(c->scene.bvhNodes, c->scene.bvhTris, n, rays, out, outT, overflow);
    } else if (any) k_trace<true, false><<<grid, 128, 0, s>>>(c->scene.bvhNodes, c->scene.bvhTris, n, rays, out, nullptr, overflow);  // the tiny-scene trace kernels
    else k_trace<false, false><<<grid, 128, 0, s>>>(c->scene.bvhNodes, c->scene.bvhTris, n, rays, out, outT, overflow);
    c->stats.kernelLaunches++;
    (any ? c->stats.raysAny : c->stats.raysClosest) += n;
    PLT_CUDA(cudaGetLastError());
    return PLTPT_OK;
}
int traceClosest(PltptContext* c, uint32_t n, const float* rays, uint32_t* hits, float* hitT, cudaStream_t s) { return traceLaunch(c, false, n, rays, hits, hitT, s); }
int traceAny(PltptContext* c, uint32_t n, const float* rays, uint32_t* occluded, cudaStream_t s) { return traceLaunch(c, true, n, rays, occluded, nullptr, s); }

}  // namespace plt
