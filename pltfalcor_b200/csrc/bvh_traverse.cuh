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

// Walks the compressed wide BVH for one ray.  ANY = true: returns at the first accepted triangle.
// `overflow` (optional) counts rays whose traversal stack overflowed (must stay 0; checked by the tests).
template <bool ANY>
PLT_D bool traverseBvh(const BvhNode* __restrict__ nodes, const BvhTri* __restrict__ tris, float3 o, float3 d, float tmin, float tmax, Hit& best, uint32_t* overflow = nullptr) {
    best.instanceID = 0xffffffffu;
    best.t = tmax;
    const RayPrep rp = prepRay(o, d);
    const float tiny = 1e-30f;
    const float idx = 1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x));
    const float idy = 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y));
    const float idz = 1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z));
    uint32_t stack[kTraversalStack];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const uint32_t ni = stack[--sp];
        const uint4* np = reinterpret_cast<const uint4*>(nodes + ni);
        const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
        const float sx = __uint_as_float((n0.w & 0xffu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23), sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
        const uint32_t imask = n0.w >> 24;
        const uint32_t childBase = n1.x, triBase = n1.y;
        const uint32_t metaLo = n1.z, metaHi = n1.w;
        // quantised planes: n2 = qlox[0..7] (x,y) qloy[0..7] (z,w); n3 = qloz, qhix; n4 = qhiy, qhiz
        float nearT[8];
        uint32_t innerHits = 0;
        // the slab test is done against the ray segment [tmin, best.t] widened by a few ulps (conservative)
        const float tfarLimit = best.t * 1.0000004f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t meta = ((i < 4 ? metaLo : metaHi) >> ((i & 3) * 8)) & 0xffu;
            const bool inner = (imask >> i) & 1u;
            if (meta == 0 && !inner) continue;
            const int sh = (i & 3) * 8;
            const uint32_t wlx = i < 4 ? n2.x : n2.y, wly = i < 4 ? n2.z : n2.w, wlz = i < 4 ? n3.x : n3.y;
            const uint32_t whx = i < 4 ? n3.z : n3.w, why = i < 4 ? n4.x : n4.y, whz = i < 4 ? n4.z : n4.w;
            const float lox = px + float((wlx >> sh) & 0xffu) * sx, hix = px + float((whx >> sh) & 0xffu) * sx;
            const float loy = py + float((wly >> sh) & 0xffu) * sy, hiy = py + float((why >> sh) & 0xffu) * sy;
            const float loz = pz + float((wlz >> sh) & 0xffu) * sz, hiz = pz + float((whz >> sh) & 0xffu) * sz;
            const float ax = (lox - o.x) * idx, bx = (hix - o.x) * idx;
            const float ay = (loy - o.y) * idy, by = (hiy - o.y) * idy;
            const float az = (loz - o.z) * idz, bz = (hiz - o.z) * idz;
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin)) * 0.9999996f;
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tfarLimit)) * 1.0000004f;
            if (!(tn <= tf)) continue;
            if (inner) { nearT[i] = tn; innerHits |= 1u << i; }
            else {
                const uint32_t count = meta >> 5, first = triBase + (meta & 31u);
                for (uint32_t k = 0; k < count; ++k) {
                    const float4* tp = reinterpret_cast<const float4*>(tris + first + k);
                    const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
                    float t, u, v;
                    if (!intersectTri(rp, q0, q1, q2, tmin, tmax, t, u, v)) continue;
                    if (ANY) { best.t = t; best.instanceID = __float_as_uint(q0.w); return true; }
                    const uint32_t g = __float_as_uint(q2.w);
                    if (best.valid() && (t > best.t || (t == best.t && g > best.globalIndex))) continue;
                    best.t = t; best.u = u; best.v = v;
                    best.instanceID = __float_as_uint(q0.w); best.primitiveIndex = __float_as_uint(q1.w); best.globalIndex = g;
                }
            }
        }
        // push the intersected inner children, farthest first, so the nearest is popped next
        while (innerHits) {
            int far = -1;
            float ft = -1.f;
            if (ANY) far = __ffs(innerHits) - 1;
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (((innerHits >> i) & 1u) && (far < 0 || nearT[i] > ft)) { far = i; ft = nearT[i]; }
            }
            innerHits &= ~(1u << far);
            if (sp < kTraversalStack) stack[sp++] = childBase + __popc(imask & ((1u << far) - 1u));
            else if (overflow) atomicAdd(overflow, 1u);
        }
    }
    return best.valid();
}

}  // namespace plt
