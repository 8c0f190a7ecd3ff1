// bvh_traverse.cuh — ray / compressed-8-wide-BVH traversal and the watertight ray-triangle test (device side).
//
// Intersection contract (shared with the oracle, oracle/plt_scene.h; DESIGN.md "Intersection contract"):
// every arithmetic step of the triangle test is an individually rounded fp32 operation (CUDA _rn intrinsics, so
// nvcc can neither contract to FMA nor reassociate); a hit needs tmin < t < tmax; closest hit = min t, exact ties
// go to the lower global triangle index.  Box tests only have to be conservative.
#pragma once
#include "context.h"

namespace plt {

struct RayPrep {
    float ox, oy, oz;
    float Sx, Sy, Sz;
    int kx, ky, kz;
};
PLT_D float pick(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }
PLT_D RayPrep prepRay(float3 o, float3 d) {
    RayPrep r;
    r.ox = o.x; r.oy = o.y; r.oz = o.z;
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    const float dz = pick(d.x, d.y, d.z, r.kz);
    if (dz < 0.0f) { const int t = r.kx; r.kx = r.ky; r.ky = t; }
    r.Sx = __fdiv_rn(pick(d.x, d.y, d.z, r.kx), dz);
    r.Sy = __fdiv_rn(pick(d.x, d.y, d.z, r.ky), dz);
    r.Sz = __fdiv_rn(1.0f, dz);
    return r;
}
// Woop, Benthin, Wald: "Watertight Ray/Triangle Intersection", JCGT 2013 — same operation order as the oracle.
PLT_D bool intersectTri(const RayPrep& r, const float4 q0, const float4 q1, const float4 q2, float tmin, float tmax, float& t, float& u, float& v) {
    const float Ax_ = __fsub_rn(q0.x, r.ox), Ay_ = __fsub_rn(q0.y, r.oy), Az_ = __fsub_rn(q0.z, r.oz);
    const float Bx_ = __fsub_rn(q1.x, r.ox), By_ = __fsub_rn(q1.y, r.oy), Bz_ = __fsub_rn(q1.z, r.oz);
    const float Cx_ = __fsub_rn(q2.x, r.ox), Cy_ = __fsub_rn(q2.y, r.oy), Cz_ = __fsub_rn(q2.z, r.oz);
    const float Akz = pick(Ax_, Ay_, Az_, r.kz), Bkz = pick(Bx_, By_, Bz_, r.kz), Ckz = pick(Cx_, Cy_, Cz_, r.kz);
    const float Ax = __fsub_rn(pick(Ax_, Ay_, Az_, r.kx), __fmul_rn(r.Sx, Akz)), Ay = __fsub_rn(pick(Ax_, Ay_, Az_, r.ky), __fmul_rn(r.Sy, Akz));
    const float Bx = __fsub_rn(pick(Bx_, By_, Bz_, r.kx), __fmul_rn(r.Sx, Bkz)), By = __fsub_rn(pick(Bx_, By_, Bz_, r.ky), __fmul_rn(r.Sy, Bkz));
    const float Cx = __fsub_rn(pick(Cx_, Cy_, Cz_, r.kx), __fmul_rn(r.Sx, Ckz)), Cy = __fsub_rn(pick(Cx_, Cy_, Cz_, r.ky), __fmul_rn(r.Sy, Ckz));
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = __double2float_rn(__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx)));
        V = __double2float_rn(__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx)));
        W = __double2float_rn(__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax)));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    if (det == 0.0f) return false;
    const float Az = __fmul_rn(r.Sz, Akz), Bz = __fmul_rn(r.Sz, Bkz), Cz = __fmul_rn(r.Sz, Ckz);
    const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));
    const float rcpDet = __fdiv_rn(1.0f, det);
    t = __fmul_rn(T, rcpDet);
    if (!(t > tmin && t < tmax)) return false;
    u = __fmul_rn(V, rcpDet);
    v = __fmul_rn(W, rcpDet);
    return true;
}

struct Hit {
    float t, u, v;
    uint32_t instanceID, primitiveIndex, globalIndex;
    PLT_D bool valid() const { return instanceID != 0xffffffffu; }
};

constexpr int kTraversalStack = 160;


// Traversal of the compressed wide BVH.  processNode() handles ONE popped node (children slab tests, leaf triangles, ordered push
// of the inner children).  traverseBvh() runs a ray to completion in the calling thread; the trace kernels of non-trivial scenes
// drive the same code through travPersistent(), where a lane whose ray has finished takes the next ray of the queue instead
// of idling until the slowest ray of its warp is done.
// (Round 2 also measured a node-GROUP stack after Ylitie et al. 2017 -- one stack entry per node, children visited in octant
// order, branch-free slab tests: 9.2 -> 11.3 ms on the 10 M-triangle city, 5.9 -> 7.7 ms for the manifold walks; rejected,
// profiles/r2c_kernel_times.log.)
// ANY = true: finishes at the first accepted triangle.  `overflow` (optional) counts stack overflows (must stay 0; tests check).
struct TraversalCounters { uint32_t nodes, tris; };
struct RayConst {
    RayPrep rp;
    float3 o;
    float idx, idy, idz, tmin, tmax;
};
PLT_D void rayConstInit(RayConst& R, float3 o, float3 d, float tmin, float tmax) {
    R.rp = prepRay(o, d);
    const float tiny = 1e-30f;
    R.idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
    R.idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
    R.idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
    R.o = o; R.tmin = tmin; R.tmax = tmax;
}
// The triangles `triMask` (bits = slots from `triBase`) of the intersected leaf children of one node.  Returns true when ANY and a triangle
// was accepted.
template <bool ANY>
PLT_D bool testLeafTriangles(const BvhTri* __restrict__ tris, const RayConst& R, uint32_t triBase, uint32_t triMask, Hit& best, TraversalCounters* tc) {
    while (triMask) {
        const uint32_t k = uint32_t(__ffs(int(triMask))) - 1u;
        triMask &= triMask - 1u;
        const float4* tp = reinterpret_cast<const float4*>(tris + triBase + k);
        const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
        float t, u, v;
        if (tc) ++tc->tris;
        if (!intersectTri(R.rp, q0, q1, q2, R.tmin, R.tmax, t, u, v)) continue;
        if (ANY) { best.t = t; best.instanceID = __float_as_uint(q0.w); return true; }
        const uint32_t g = __float_as_uint(q2.w);
        if (best.valid() && (t > best.t || (t == best.t && g > best.globalIndex))) continue;
        best.t = t; best.u = u; best.v = v;
        best.instanceID = __float_as_uint(q0.w); best.primitiveIndex = __float_as_uint(q1.w); best.globalIndex = g;
    }
    return false;
}
// One popped node: slab tests of its children, ordered push of the intersected inner children; the triangles of its intersected leaf
// children are returned in (triBaseOut, triMaskOut) for testLeafTriangles() -- at once (traverseBvh) or batched over the lanes of a warp
// (travPersistent with PLT_DEFER_LEAVES).
template <bool ANY, bool kNearest>
PLT_D void processNode(const BvhNode* __restrict__ nodes, const RayConst& R, uint32_t ni, float bestT, uint32_t* stack, float* stackT, int& sp,
                       uint32_t* overflow, TraversalCounters* tc, uint32_t& triBaseOut, uint32_t& triMaskOut) {
    const float3 o = R.o;
    const float idx = R.idx, idy = R.idy, idz = R.idz, tmin = R.tmin;
    const bool negx = idx < 0.f, negy = idy < 0.f, negz = idz < 0.f;
    if (tc) ++tc->nodes;
        const uint4* np = reinterpret_cast<const uint4*>(nodes + ni);
        const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        const uint32_t imask = n0.w >> 24;
        const uint32_t childBase = n1.x, triBase = n1.y;
        const uint32_t metaLo = n1.z, metaHi = n1.w;
        // Slab distances in one FMA per plane: t = q * (s * id) + (p - o) * id.  The library is built with -fmad=false
        // (fp32 parity of the shading math with the oracle), so the traversal asks for its FMAs explicitly.
        // Error bound of a plane distance: 4u (|(p-o) id| + 255 |s id|), u = 2^-24; it is folded into the two per-axis
        // offsets (near planes pulled in, far planes pushed out), so the test stays conservative at no per-child cost.
        const float sxi = __uint_as_float((n0.w & 0xffu) << 23) * idx, syi = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * idy, szi = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * idz;
        const float pxi = (__uint_as_float(n0.x) - o.x) * idx, pyi = (__uint_as_float(n0.y) - o.y) * idy, pzi = (__uint_as_float(n0.z) - o.z) * idz;
        const float ex = 3e-7f * fmaf(255.f, fabsf(sxi), fabsf(pxi)), ey = 3e-7f * fmaf(255.f, fabsf(syi), fabsf(pyi)), ez = 3e-7f * fmaf(255.f, fabsf(szi), fabsf(pzi));
        const float pxn = pxi - ex, pxf = pxi + ex, pyn = pyi - ey, pyf = pyi + ey, pzn = pzi - ez, pzf = pzi + ez;
        // quantised planes: n2 = qlox[0..7] (x,y) qloy[0..7] (z,w); n3 = qloz, qhix; n4 = qhiy, qhiz.  Per axis the plane
        // the ray enters through is `lo` for a positive direction component and `hi` for a negative one.
        const uint32_t nxA = negx ? n3.z : n2.x, nxB = negx ? n3.w : n2.y, fxA = negx ? n2.x : n3.z, fxB = negx ? n2.y : n3.w;
        const uint32_t nyA = negy ? n4.x : n2.z, nyB = negy ? n4.y : n2.w, fyA = negy ? n2.z : n4.x, fyB = negy ? n2.w : n4.y;
        const uint32_t nzA = negz ? n4.z : n3.x, nzB = negz ? n4.w : n3.y, fzA = negz ? n3.x : n4.z, fzB = negz ? n3.y : n4.w;
        // kNearest: intersected inner children are pushed straight from the slab-test loop and only the nearest is then moved to the top of
        // the stack (the others keep slot order and are culled against the closest hit when popped); otherwise all of them are pushed in
        // order of distance, farthest first.  The ordered push was 28 % of k_trace's issue slots at 6 of 32 lanes on the 10 M-triangle city
        // (profiles/r2o_trace_sass_segments.txt); on a scene of a handful of nodes it is the faster of the two (profiles/r2o_variants.log).
        const int sp0 = sp;
        float nearT[8];
        uint32_t innerHits = 0;
        uint32_t triMask = 0;
        // exact-t ties in another leaf must still be visited: keep the far limit a few ulps beyond the best hit
        const float tfarLimit = bestT * 1.000002f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t meta = ((i < 4 ? metaLo : metaHi) >> ((i & 3) * 8)) & 0xffu;
            const bool inner = (imask >> i) & 1u;
            if (meta == 0 && !inner) continue;
            const int sh = (i & 3) * 8;
            const float tnx = fmaf(float(((i < 4 ? nxA : nxB) >> sh) & 0xffu), sxi, pxn), tfx = fmaf(float(((i < 4 ? fxA : fxB) >> sh) & 0xffu), sxi, pxf);
            const float tny = fmaf(float(((i < 4 ? nyA : nyB) >> sh) & 0xffu), syi, pyn), tfy = fmaf(float(((i < 4 ? fyA : fyB) >> sh) & 0xffu), syi, pyf);
            const float tnz = fmaf(float(((i < 4 ? nzA : nzB) >> sh) & 0xffu), szi, pzn), tfz = fmaf(float(((i < 4 ? fzA : fzB) >> sh) & 0xffu), szi, pzf);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tfarLimit));
            if (!(tn <= tf)) continue;
            if (inner) {
                if (kNearest) {
                    if (sp < kTraversalStack) {
                        if (!ANY) stackT[sp] = tn;
                        stack[sp++] = childBase + __popc(imask & ((1u << i) - 1u));
                    } else if (overflow) atomicAdd(overflow, 1u);
                } else { nearT[i] = tn; innerHits |= 1u << i; }
            }
            else triMask |= ((1u << (meta >> 5)) - 1u) << (meta & 31u);  // the leaf's triangles: slots [offset, offset + count) of this node's <= 24
        }
        // The triangles of all intersected leaf children are tested in ONE loop after the slab tests (testLeafTriangles): tested inside the
        // unrolled child loop, every child was a separate divergence point.
        triBaseOut = triBase; triMaskOut = triMask;
        // the nearest of the children just pushed goes on top (popped next); the others stay in slot order and are culled at pop time
        if (kNearest && !ANY && sp - sp0 > 1) {
            int bestK = sp0;
            float bt = stackT[sp0];
            for (int k = sp0 + 1; k < sp; ++k) { const float t = stackT[k]; if (t < bt) { bt = t; bestK = k; } }
            if (bestK != sp - 1) {
                const uint32_t n = stack[bestK]; stack[bestK] = stack[sp - 1]; stack[sp - 1] = n;
                stackT[bestK] = stackT[sp - 1]; stackT[sp - 1] = bt;
            }
        }
        // push the intersected inner children, farthest first, so the nearest is popped next
        while (!kNearest && innerHits) {
            int far = -1;
            float ft = -1.f;
            if (ANY) far = __ffs(innerHits) - 1;
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (((innerHits >> i) & 1u) && (far < 0 || nearT[i] > ft)) { far = i; ft = nearT[i]; }
            }
            innerHits &= ~(1u << far);
            if (sp < kTraversalStack) {
                if (!ANY) stackT[sp] = ft;
                stack[sp++] = childBase + __popc(imask & ((1u << far) - 1u));
            } else if (overflow) atomicAdd(overflow, 1u);
        }
}
// One ray to completion in the calling thread: the same processNode() the persistent-lane driver steps through (one node body, so the
// bit-exact hit-ID checks of either driver cover the arithmetic of both; measured equal to a hand-inlined copy, profiles/r2a_variants.log).
template <bool ANY, bool kNearest = true>
PLT_D bool traverseBvh(const BvhNode* __restrict__ nodes, const BvhTri* __restrict__ tris, float3 o, float3 d, float tmin, float tmax, Hit& best, uint32_t* overflow = nullptr,
                       TraversalCounters* tc = nullptr) {
    best.instanceID = 0xffffffffu;
    best.t = tmax;
    RayConst R;
    rayConstInit(R, o, d, tmin, tmax);
    uint32_t stack[kTraversalStack];
    float stackT[ANY ? 1 : kTraversalStack];  // entry distance of the pushed node: a closer hit found meanwhile culls it at pop time
    int sp = 0;
    stack[sp++] = 0;
    if (!ANY) stackT[0] = tmin;
    while (sp > 0) {
        const uint32_t ni = stack[--sp];
        if (!ANY && stackT[sp] > best.t * 1.000002f) continue;  // (same slack as tfarLimit in processNode: exact-t ties are still visited)
        uint32_t triBase, triMask;
        processNode<ANY, kNearest>(nodes, R, ni, best.t, stack, stackT, sp, overflow, tc, triBase, triMask);
        if (testLeafTriangles<ANY>(tris, R, triBase, triMask, best, tc)) return true;
    }
    return best.valid();
}

// Resumable form of the same loop for travPersistent().
template <bool ANY>
struct Trav {
    RayConst R;
    Hit best;
    int sp;
    uint32_t stack[kTraversalStack];
    float stackT[ANY ? 1 : kTraversalStack];
};
template <bool ANY>
PLT_D void travInit(Trav<ANY>& T, float3 o, float3 d, float tmin, float tmax) {
    T.best.instanceID = 0xffffffffu;
    T.best.t = tmax;
    rayConstInit(T.R, o, d, tmin, tmax);
    T.sp = 0;
    T.stack[T.sp++] = 0;
    if (!ANY) T.stackT[0] = tmin;
}
// One node; returns true when the ray is finished (stack empty, or ANY and a triangle was accepted).
template <bool ANY>
PLT_D bool travStep(const BvhNode* __restrict__ nodes, const BvhTri* __restrict__ tris, Trav<ANY>& T, uint32_t* overflow, TraversalCounters* tc) {
    if (T.sp <= 0) return true;
    const uint32_t ni = T.stack[--T.sp];
    if (!ANY && T.stackT[T.sp] > T.best.t * 1.000002f) return T.sp == 0;
    uint32_t triBase, triMask;
    processNode<ANY, false>(nodes, T.R, ni, T.best.t, T.stack, T.stackT, T.sp, overflow, tc, triBase, triMask);
    if (testLeafTriangles<ANY>(tris, T.R, triBase, triMask, T.best, tc)) { T.sp = 0; return true; }
    return T.sp == 0;
}
// The same without the triangle tests: the node's leaf triangles are handed back (triMask == 0: none).
template <bool ANY>
PLT_D void travStepDeferred(const BvhNode* __restrict__ nodes, Trav<ANY>& T, uint32_t* overflow, TraversalCounters* tc, uint32_t& triBase, uint32_t& triMask) {
    triMask = 0;
    if (T.sp <= 0) return;
    const uint32_t ni = T.stack[--T.sp];
    // (popping on until an entry survives the cull, so that the lane never idles through a node step, measured slower: city trace 8.2 -> 9.5 ms,
    // profiles/r2p_kernel_times.log against r2o_kernel_times.log)
    if (!ANY && T.stackT[T.sp] > T.best.t * 1.000002f) return;
    processNode<ANY, true>(nodes, T.R, ni, T.best.t, T.stack, T.stackT, T.sp, overflow, tc, triBase, triMask);
}

// Persistent-lane driver for a queue of `n` rays (Aila & Laine 2009, "Understanding the Efficiency of Ray Traversal on GPUs",
// dynamic fetch): fetch(i, o, d, tmin, tmax) loads ray i (false: nothing to trace for this entry), commit(i, hit, best) consumes its
// result.  Idle lanes refill from `cursor` (one warp-aggregated bump) once at least kRefillLanes lanes are idle; between refills
// every lane with a ray processes kStepsPerRound nodes.  All 32 lanes of a warp must call this together.
// PLT_DEFER_LEAVES = 1: the leaf triangles a lane meets during a round are queued (at most one set per node step) and tested after the
// round by all lanes together.  Tested inside the step, the triangle code ran for whichever 2-4 lanes of the warp had just met a leaf
// (a quarter of the kernel's issue slots at 1.8 of 32 lanes on incoherent rays, profiles/r2o_trace_sass_segments.txt).
#ifndef PLT_TRAV_REFILL
#define PLT_TRAV_REFILL 8
#endif
#ifndef PLT_TRAV_STEPS
#define PLT_TRAV_STEPS 4
#endif
#ifndef PLT_DEFER_LEAVES
#define PLT_DEFER_LEAVES 1  // 0: the round-1 driver (triangles tested inside the step, ordered push), kept for comparison (tools/variants.py: trav_legacy)
#endif
constexpr int kRefillLanes = PLT_TRAV_REFILL, kStepsPerRound = PLT_TRAV_STEPS;
template <bool ANY, class Fetch, class Commit>
PLT_D void travPersistent(const BvhNode* __restrict__ nodes, const BvhTri* __restrict__ tris, uint32_t n, uint32_t* cursor, uint32_t* overflow, TraversalCounters* tc,
                          Fetch fetch, Commit commit) {
    const int lane = threadIdx.x & 31;
    Trav<ANY> T;
    T.sp = 0;
    uint32_t item = 0;
    bool have = false, exhausted = n == 0;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !have);
        if (!exhausted && (__popc(idle) >= kRefillLanes)) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(cursor, uint32_t(__popc(idle)));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!have) {
                const uint32_t i = base + __popc(idle & ((1u << lane) - 1u));
                if (i < n) {
                    float3 o, d;
                    float tmin, tmax;
                    if (fetch(i, o, d, tmin, tmax)) { item = i; travInit(T, o, d, tmin, tmax); have = true; }
                }
            }
            exhausted = base + __popc(idle) >= n;  // warp-uniform
        }
        if (!__any_sync(0xffffffffu, have)) { if (exhausted) break; else continue; }
#if PLT_DEFER_LEAVES
        uint32_t pendBase[kStepsPerRound], pendMask[kStepsPerRound];
        int np = 0;
#pragma unroll 1
        for (int s = 0; s < kStepsPerRound; ++s) {
            if (have && T.sp > 0) {
                uint32_t tb, tm;
                travStepDeferred(nodes, T, overflow, tc, tb, tm);
                if (tm) { pendBase[np] = tb; pendMask[np] = tm; ++np; }
            }
        }
        bool accepted = false;
        while (np > 0) {
            --np;
            if (testLeafTriangles<ANY>(tris, T.R, pendBase[np], pendMask[np], T.best, tc)) { accepted = true; break; }
        }
        if (have && (accepted || T.sp <= 0)) { commit(item, T.best.valid(), T.best); have = false; T.sp = 0; }
#else
#pragma unroll 1
        for (int s = 0; s < kStepsPerRound; ++s) {
            if (have && travStep(nodes, tris, T, overflow, tc)) { commit(item, T.best.valid(), T.best); have = false; }
        }
#endif
    }
}

}  // namespace plt
