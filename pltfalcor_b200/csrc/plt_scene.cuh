// plt_scene.cuh — device-side access to the Falcor scene buffers: index/vertex fetch, world transform, shading frame,
// hit packing, differential geometry.  Semantics follow Source/Falcor/Scene/{Scene,SceneTypes,HitInfo}.slang,
// Scene/Material/{MaterialFactory,ShadingUtils}.slang and Rendering/Utils/DifferentialShadingGeometry.slang (cited per function).
#pragma once
#include "context.h"

namespace plt {

struct TriHit {
    uint32_t instanceID, primitiveIndex;
    float u, v;
    PLT_D float3 weights() const { return mk3(1.f - u - v, u, v); }  // HitInfo.slang:76-79
};
struct StaticVertex { float3 position, normal; uint32_t packedTangent; float tangentSign; float2 texCrd; };  // tangent xyz decoded on demand (tangentOf)
struct VertexData { float3 posW, normalW, faceNormalW; float4 tangentW; float2 texC; };  // SceneTypes.slang:311-320

// Scene.slang getLocalIndices / getIndices
PLT_D void getIndices(const DevScene& sc, const PltptGeometryInstance& inst, uint32_t tri, uint32_t idx[3]) {
    const uint32_t indexCount = __ldg(&sc.meshes[inst.geometryID].indexCount);
    if (indexCount == 0) { idx[0] = tri * 3 + inst.vbOffset; idx[1] = idx[0] + 1; idx[2] = idx[0] + 2; return; }
    const uint8_t* base = sc.indexData + size_t(inst.ibOffset) * 4;
    if (inst.flags & PLTPT_INSTANCE_FLAG_USE_16BIT_INDICES) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(base + size_t(tri) * 6);
        idx[0] = uint32_t(__ldg(p)) + inst.vbOffset; idx[1] = uint32_t(__ldg(p + 1)) + inst.vbOffset; idx[2] = uint32_t(__ldg(p + 2)) + inst.vbOffset;
    } else {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(base + size_t(tri) * 12);
        idx[0] = __ldg(p) + inst.vbOffset; idx[1] = __ldg(p + 1) + inst.vbOffset; idx[2] = __ldg(p + 2) + inst.vbOffset;
    }
}
PLT_D PltptGeometryInstance loadInstance(const DevScene& sc, uint32_t instanceID) {
    PltptGeometryInstance g;
    const uint4* p = reinterpret_cast<const uint4*>(sc.instances + instanceID);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    g.flags = a.x; g.globalMatrixID = a.y; g.materialID = a.z; g.geometryID = a.w;
    g.vbOffset = b.x; g.ibOffset = b.y; g.instanceIndex = b.z; g.geometryIndex = b.w;
    return g;
}
// SceneTypes.slang:282-296 (PackedStaticVertexData::unpack): one 32 B vertex = two 16 B loads
PLT_D StaticVertex fetchVertex(const DevScene& sc, uint32_t index) {
    const float4* p = reinterpret_cast<const float4*>(sc.vertices + index);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    StaticVertex v;
    v.position = mk3(a.x, a.y, a.z);
    const uint32_t pk0 = __float_as_uint(a.w), pk1 = __float_as_uint(b.x), pk2 = __float_as_uint(b.y);
    v.texCrd = mk2(b.z, b.w);
    v.normal = normalize(mk3(h2f(pk0 & 0xffff), h2f(pk0 >> 16), h2f(pk1 & 0xffff)));
    v.packedTangent = pk2;
    v.tangentSign = signf(h2f(pk1 >> 16));
    return v;
}
PLT_D float4 tangentOf(const StaticVertex& v) { const float3 t = decodeNormal2x16(v.packedTangent); return mk4(t.x, t.y, t.z, v.tangentSign); }
PLT_D float3 mulPoint(const float* __restrict__ M, float3 p) {
    return mk3(__ldg(M + 0) * p.x + __ldg(M + 1) * p.y + __ldg(M + 2) * p.z + __ldg(M + 3), __ldg(M + 4) * p.x + __ldg(M + 5) * p.y + __ldg(M + 6) * p.z + __ldg(M + 7),
               __ldg(M + 8) * p.x + __ldg(M + 9) * p.y + __ldg(M + 10) * p.z + __ldg(M + 11));
}
PLT_D float3 mulVec(const float* __restrict__ M, float3 p) {
    return mk3(__ldg(M + 0) * p.x + __ldg(M + 1) * p.y + __ldg(M + 2) * p.z, __ldg(M + 4) * p.x + __ldg(M + 5) * p.y + __ldg(M + 6) * p.z,
               __ldg(M + 8) * p.x + __ldg(M + 9) * p.y + __ldg(M + 10) * p.z);
}

// Scene.slang computeVertexData / getVertexData (triangle meshes)
PLT_D VertexData getVertexData(const DevScene& sc, const TriHit& hit, uint32_t& materialID, StaticVertex* outVerts = nullptr) {
    const PltptGeometryInstance inst = loadInstance(sc, hit.instanceID);
    materialID = inst.materialID;
    uint32_t idx[3];
    getIndices(sc, inst, hit.primitiveIndex, idx);
    const StaticVertex v0 = fetchVertex(sc, idx[0]), v1 = fetchVertex(sc, idx[1]), v2 = fetchVertex(sc, idx[2]);
    if (outVerts) { outVerts[0] = v0; outVerts[1] = v1; outVerts[2] = v2; }
    const float* worldMat = sc.world + 16 * size_t(inst.globalMatrixID);
    const float* worldIT = sc.worldIT + 16 * size_t(inst.globalMatrixID);
    const bool isFrontFaceCW = (inst.flags & PLTPT_INSTANCE_FLAG_OBJECT_FRONT_FACE_CW) != 0;
    const float3 b = hit.weights();
    VertexData v;
    v.posW = v0.position * b.x; v.posW += v1.position * b.y; v.posW += v2.position * b.z;
    v.normalW = v0.normal * b.x; v.normalW += v1.normal * b.y; v.normalW += v2.normal * b.z;
    // A mesh without tangents stores handedness 0 in every vertex (SceneTypes.slang:252-266): the interpolated sign is then 0, the tangent
    // space is invented from the normal (prepareShadingData) and the tangent's direction is never read -- its three octahedral decodes, the
    // transform and the normalisation (a fifth of the shading kernel's instructions on such meshes) are skipped; the result is the same.
    const bool hasTangent = v0.tangentSign != 0.f || v1.tangentSign != 0.f || v2.tangentSign != 0.f;
    float4 tg = mk4(0.f);
    if (hasTangent) {
        tg = tangentOf(v0) * b.x; tg += tangentOf(v1) * b.y; tg += tangentOf(v2) * b.z;
        tg.w = signf(tg.w);
    }
    v.texC = v0.texCrd * b.x + v1.texCrd * b.y + v2.texCrd * b.z;
    const float3 N = normalize(cross(v1.position - v0.position, v2.position - v0.position));
    v.faceNormalW = isFrontFaceCW ? -N : N;
    v.posW = mulPoint(worldMat, v.posW);
    v.normalW = normalize(mulVec(worldIT, v.normalW));
    v.faceNormalW = normalize(mulVec(worldIT, v.faceNormalW));
    v.tangentW = mk4(0.f);
    if (hasTangent) {
        const float3 t = normalize(mulVec(worldMat, mk3(tg.x, tg.y, tg.z)));
        v.tangentW = mk4(t.x, t.y, t.z, tg.w);
    }
    return v;
}

// The material a hit sees: the scene's pre-derived record, or -- for a material with textured parameter slots -- the record derived from
// the slots sampled at the hit's texture coordinate (plt_material.cuh; every *Material.slang::setupMaterialInstance).
// kTex = false: the instantiation the hot stage kernels run on a scene without textured materials (no out-of-line call in their body:
// its mere presence cost k_connect_a 372 B of spills and 15 % of its time, profiles/r2k_variants.log); the launchers pick by scene.
template <bool kTex = true>
PLT_D DevMaterial loadMaterialAt(const DevScene& sc, uint32_t id, float2 uv) {
    DevMaterial m = loadMaterial(sc.materials, id);
    if (kTex && sc.tex.blobs && (m.hdrFlags & HF_Textured)) m = resolveTexturedMaterial(sc.tex, sc.profiles, id, uv, m.hdrFlags);
    return m;
}
// MaterialFactory.slang:79-131 (prepareShadingData) + ShadingUtils.slang:126-152 (tangent frame), :199-246 (normal mapping).
template <bool kTex = true>
PLT_D SD prepareShadingData(const DevScene& sc, const VertexData& v, uint32_t materialID, const DevMaterial& m, float3 viewDir) {
    SD sd;
    sd.posW = v.posW;
    sd.uv = v.texC;
    sd.V = viewDir;
    sd.N = v.normalW;
    const float3 tw = mk3(v.tangentW.x, v.tangentW.y, v.tangentW.z);
    const float NdotT = dot(tw, sd.N);
    const bool valid = fabsf(v.tangentW.w) == 1.0f && dot(tw, tw) > 0.f && fabsf(NdotT) < 0.9999f;
    if (valid) {
        sd.T = normalize(tw - sd.N * NdotT);
        sd.B = cross(sd.N, sd.T) * v.tangentW.w;
    } else {
        sd.T = perp_stark(sd.N);
        sd.B = cross(sd.N, sd.T);
    }
    sd.faceN = v.faceNormalW;
    sd.frontFacing = dot(sd.V, sd.faceN) >= 0.f;
    sd.materialID = materialID;
    if (kTex && sc.tex.blobs && valid && (m.hdrFlags & HF_NormalMap)) {  // (sc.tex.blobs: null for a scene without textured materials -- a uniform test)  // modifyTangentSpace: only with a valid tangent space (MaterialFactory.slang:114-120)
        const Frame3 f = applyNormalMapTextured(sc.tex, materialID, v.texC, v.tangentW, Frame3{sd.T, sd.B, sd.N});
        sd.T = f.T; sd.B = f.B; sd.N = f.N;
    }
    if (!sd.frontFacing && (m.hdrFlags & HF_DoubleSided)) sd.N = -sd.N;
    return sd;
}

// HitInfo.slang:459-487 / 96-130 (uncompressed 128-bit format, Triangle type = 1)
PLT_D uint4 packHit(const DevScene& sc, const TriHit& h) {
    const uint32_t typeOffset = 32u - sc.hitTypeBits;
    const uint32_t headerBits = sc.hitTypeBits + sc.hitInstBits + sc.hitPrimBits;
    uint4 o = make_uint4(0, 0, __float_as_uint(h.u), __float_as_uint(h.v));
    if (headerBits <= 32) o.x = (1u << typeOffset) | (h.instanceID << sc.hitPrimBits) | h.primitiveIndex;
    else { o.x = (1u << typeOffset) | h.instanceID; o.y = h.primitiveIndex; }
    return o;
}
PLT_D TriHit unpackHit(const DevScene& sc, uint4 in) {
    TriHit h;
    const uint32_t headerBits = sc.hitTypeBits + sc.hitInstBits + sc.hitPrimBits;
    if (headerBits <= 32) {
        h.instanceID = (in.x >> sc.hitPrimBits) & ((1u << sc.hitInstBits) - 1u);
        h.primitiveIndex = in.x & ((1u << sc.hitPrimBits) - 1u);
    } else {
        h.instanceID = in.x & ((1u << sc.hitInstBits) - 1u);
        h.primitiveIndex = in.y;
    }
    h.u = __uint_as_float(in.z);
    h.v = __uint_as_float(in.w);
    return h;
}
// LightCollection.slang:97-101
PLT_D uint32_t getEmissiveTriangleIndex(const DevScene& sc, uint32_t instanceID, uint32_t primitiveIndex) {
    if (!sc.perMeshInstanceOffset) return 0xffffffffu;
    const uint32_t offset = __ldg(sc.perMeshInstanceOffset + instanceID);
    return offset != 0xffffffffu ? offset + primitiveIndex : 0xffffffffu;
}

// ---- differential geometry: Rendering/Utils/DifferentialShadingGeometry.slang ---------------------------------------
struct V3pdv { float3 v, ddu, ddv; PLT_D void flip() { v = -v; ddu = -ddu; ddv = -ddv; } };
struct V3kpdv { float3 v, ddu, ddv, ddl; };
PLT_D V3kpdv zeroKpdv() { V3kpdv r; r.v = r.ddu = r.ddv = r.ddl = mk3(0.f); return r; }
struct DiffGeom {
    V3pdv p, n, s, t;
    PLT_D V3pdv pdw(float3 dst) const {  // :6-18
        V3pdv w;
        w.v = dst - p.v;
        const float ilo = 1.f / length(w.v);
        w.v *= ilo;
        w.ddu = -ilo * (p.ddu - w.v * dot(w.v, p.ddu));
        w.ddv = -ilo * (p.ddv - w.v * dot(w.v, p.ddv));
        return w;
    }
    PLT_D float3 fromLocal(float3 v) const { return s.v * v.x + t.v * v.y + n.v * v.z; }
};
PLT_D bool allZero(float2 a) { return a.x == 0.f && a.y == 0.f; }
PLT_D bool allZero(float3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
// :94-179 (object-space, like the reference: it never applies the instance transform here)
PLT_D DiffGeom differentialShadingGeometry(const DevScene& sc, const TriHit& hit) {
    const PltptGeometryInstance inst = loadInstance(sc, hit.instanceID);
    uint32_t idx[3];
    getIndices(sc, inst, hit.primitiveIndex, idx);
    const StaticVertex a = fetchVertex(sc, idx[0]), b = fetchVertex(sc, idx[1]), c = fetchVertex(sc, idx[2]);
    const float3 bary = hit.weights();
    float2 uv0 = a.texCrd, uv1 = b.texCrd, uv2 = c.texCrd;
    if (allZero(uv0) && allZero(uv1) && allZero(uv2)) { uv1 = mk2(1.0f, 0.0f); uv2 = mk2(0.0f, 1.0f); }
    const float3 p0 = a.position, p1 = b.position, p2 = c.position, n0 = a.normal, n1 = b.normal, n2 = c.normal;
    DiffGeom g;
    g.p.v = bary.x * p0 + bary.y * p1 + bary.z * p2;
    g.n.v = bary.x * n0 + bary.y * n1 + bary.z * n2;
    const float il = 1.0f / length(g.n.v);
    g.n.v *= il;
    g.p.ddu = perp_stark(g.n.v);
    g.p.ddv = cross(g.n.v, g.p.ddu);
    const float2 duv0 = uv1 - uv0, duv1 = uv2 - uv0;
    const float3 dp0 = p1 - p0, dp1 = p2 - p0;
    g.n.ddu = (n1 - n0) * il;
    g.n.ddv = (n2 - n0) * il;
    g.n.ddu -= g.n.v * dot(g.n.v, g.n.ddu);
    g.n.ddv -= g.n.v * dot(g.n.v, g.n.ddv);
    const float det = duv0.x * duv1.y - duv0.y * duv1.x;
    if (fabsf(det) > kFltEps) {
        const float invDet = 1.0f / det;
        g.p.ddu = (duv1.y * dp0 - duv0.y * dp1) * invDet;
        g.p.ddv = (-duv1.x * dp0 + duv0.x * dp1) * invDet;
        const float3 nddu = (duv1.y * g.n.ddu - duv0.y * g.n.ddv) * invDet;
        const float3 nddv = (-duv1.x * g.n.ddu + duv0.x * g.n.ddv) * invDet;
        g.n.ddu = nddu;
        g.n.ddv = nddv;
    } else { g.n.ddu = mk3(0.0f); g.n.ddv = mk3(0.0f); }
    g.s.v = g.p.ddu - g.n.v * dot(g.n.v, g.p.ddu);
    float ils = 1.0f / length(g.s.v);
    g.s.v *= ils;
    if (allZero(g.p.ddu)) { g.s.v = perp_stark(g.n.v); ils = 1.0f; }
    g.t.v = cross(g.n.v, g.s.v);
    g.s.ddu = ils * (-g.n.ddu * dot(g.n.v, g.p.ddu) - g.n.v * dot(g.n.ddu, g.p.ddu));
    g.s.ddv = ils * (-g.n.ddv * dot(g.n.v, g.p.ddu) - g.n.v * dot(g.n.ddv, g.p.ddu));
    g.s.ddu -= g.s.v * dot(g.s.ddu, g.s.v);
    g.s.ddv -= g.s.v * dot(g.s.ddv, g.s.v);
    g.t.ddu = cross(g.n.ddu, g.s.v) + cross(g.n.v, g.s.ddu);
    g.t.ddv = cross(g.n.ddv, g.s.v) + cross(g.n.v, g.s.ddv);
    return g;
}
// :181-222
PLT_D V3kpdv reflectDerivatives(const V3pdv& wi, const V3pdv& n) {
    const float wn = dot(wi.v, n.v);
    const float dwdun = dot(wi.ddu, n.v), dwdvn = dot(wi.ddv, n.v);
    const float dnduw = dot(n.ddu, wi.v), dndvw = dot(n.ddv, wi.v);
    V3kpdv ret;
    ret.v = 2.f * wn * n.v - wi.v;
    ret.ddu = 2.f * ((dwdun + dnduw) * n.v + wn * n.ddu) - wi.ddu;
    ret.ddv = 2.f * ((dwdvn + dndvw) * n.v + wn * n.ddv) - wi.ddv;
    ret.ddl = mk3(.0f);
    return ret;
}
PLT_D V3kpdv refractDerivatives(const V3pdv& wi, V3pdv n, float eta) {
    const float wn = dot(wi.v, n.v);
    eta = wn < .0f ? eta : 1.f / eta;
    if (wn < .0f) n.flip();
    const float eta2 = sqr(eta);
    const float dwdun = dot(wi.ddu, n.v), dwdvn = dot(wi.ddv, n.v);
    const float dnduw = dot(n.ddu, wi.v), dndvw = dot(n.ddv, wi.v);
    const float root = sqrtf(fmaxf(.0f, 1.f - eta2 * (1.0f - sqr(wn))));
    if (root == .0f) return zeroKpdv();
    const float a = eta - eta2 * wn / root;
    V3kpdv ret;
    ret.v = -eta * (wi.v - n.v * wn) - n.v * root;
    ret.ddu = a * (dnduw + dwdun) * n.v + (eta * wn - root) * n.ddu - eta * wi.ddu;
    ret.ddv = a * (dndvw + dwdvn) * n.v + (eta * wn - root) * n.ddv - eta * wi.ddv;
    ret.ddl = mk3(0.f);  // dEta/dLambda = 0 at every reference call site (PLTDielectricBSDF.slang partialDerivatives)
    return ret;
}
// DiffractionGrating.slang:183-265
PLT_D V3kpdv gratingDerivatives(const Grating& g, const DiffGeom& dsg, const V3pdv& wi, int lx, int ly, float wavelength) {
    const V3pdv n = dsg.n;
    V3pdv X;
    X.v = dsg.fromLocal(mk3(g.gx.x, g.gx.y, .0f));
    X.ddu = g.gx.x * dsg.s.ddu + g.gx.y * dsg.t.ddu;
    X.ddv = g.gx.x * dsg.s.ddv + g.gx.y * dsg.t.ddv;
    if (lx == 0 && ly == 0) return reflectDerivatives(wi, n);
    const float lambda = 1e-3f * wavelength;
    const float2 lp = mk2(lx * g.invPitch.x, ly * g.invPitch.y);
    const float3 Y = cross(n.v, X.v);
    const float x = dot(wi.v, X.v), y = dot(wi.v, Y), z = dot(wi.v, n.v);
    const float2 p = mk2(sqrtf(sqr(x) + sqr(z)), sqrtf(sqr(y) + sqr(z)));
    const float2 invp = mk2(p.x > kFltEps ? 1.f / p.x : .0f, p.y > kFltEps ? 1.f / p.y : .0f);
    const float2 sini = mk2(x * invp.x, y * invp.y);
    const float2 sino = mk2(lambda * lp.x - sini.x, lambda * lp.y - sini.y);
    const float a = sino.x;
    const float oneMinusA2 = fmaxf(.0f, 1 - sqr(a));
    const float b = sino.y;
    const float m = -oneMinusA2 / (sqr(a * b) - 1);
    const float qq = 1 - sqr(b) * m;
    if (sino.x > 1.f || sino.x < -1.f || sino.y > 1.f || sino.y < -1.f || qq < .0f || !(m >= .0f)) return zeroKpdv();
    const float sqrtOneM = sqrtf(qq), sqrtM = sqrtf(m);
    const float c = .5f / (sqrtOneM * sqrtf(oneMinusA2));
    const float dxdu = dot(wi.v, X.ddu) + dot(wi.ddu, X.v), dxdv = dot(wi.v, X.ddv) + dot(wi.ddv, X.v);
    const float dzdu = dot(wi.v, n.ddu) + dot(wi.ddu, n.v), dzdv = dot(wi.v, n.ddv) + dot(wi.ddv, n.v);
    const float2 invp3 = mk2(sqr(invp.x) * invp.x, sqr(invp.y) * invp.y);
    const float2 dsino_du = mk2(z * (x * dzdu - z * dxdu) * invp3.x, z * (y * dzdu - z * dxdu) * invp3.y);
    const float2 dsino_dv = mk2(z * (x * dzdv - z * dxdv) * invp3.x, z * (y * dzdv - z * dxdv) * invp3.y);
    const float dadu = dsino_du.x, dbdu = dsino_du.y, dadv = dsino_dv.x, dbdv = dsino_dv.y, dadl = lp.x, dbdl = lp.y;
    const float den = sqr(sqr(a * b) - 1);
    const float dmdu = 2 * a * (a * b * oneMinusA2 * dbdu - dadu * (1 - sqr(b))) / den;
    const float dmdv = 2 * a * (a * b * oneMinusA2 * dbdv - dadv * (1 - sqr(b))) / den;
    const float dmdl = 2 * a * (a * b * oneMinusA2 * dbdl - dadl * (1 - sqr(b))) / den;
    V3kpdv ret;
    ret.v = X.v * (a * sqrtOneM) + Y * (b * sqrtM) + n.v * sqrtf(fmaxf(.0f, 1 - sqr(a) * qq - sqr(b) * m));
#define PLT_GRATING_DERIV(da, db, dm)                                                                                   \
    (X.v * ((qq * (da) - a * b * (b * (dm) / 2.f + m * (db))) / sqrtOneM) + Y * ((b * (dm) / 2.f + m * (db)) / sqrtM) + \
     n.v * ((-2 * a * (da) * qq - b * (b * (dm) + 2.f * m * (db)) + sqr(a) * b * (b * (dm) + 2.f * m * (db))) * c))
    ret.ddu = PLT_GRATING_DERIV(dadu, dbdu, dmdu);
    ret.ddv = PLT_GRATING_DERIV(dadv, dbdv, dmdv);
    ret.ddl = PLT_GRATING_DERIV(dadl, dbdl, dmdl);
#undef PLT_GRATING_DERIV
    return ret;
}
// IMaterialInstance.partialDerivatives dispatch (per-material wrappers, e.g. PLTDielectricBSDF.slang partialDerivatives)
PLT_D V3kpdv bsdfPartialDerivatives(const DevMaterial& m, const SD& sd, const V3pdv& wi, const DiffGeom& dsg, uint32_t lobe, float wavelength) {
    switch (m.type) {
        case PLTPT_MATERIAL_DIFFUSE: return zeroKpdv();
        case PLTPT_MATERIAL_DIELECTRIC: {
            const float k = k2Pi / (1e-3f * wavelength);
            return (lobe & L_Transmission) != 0 ? refractDerivatives(wi, dsg.n, dielectricIor(m, k)) : reflectDerivatives(wi, dsg.n);
        }
        case PLTPT_MATERIAL_THIN_DIELECTRIC:
            if ((lobe & L_Transmission) != 0) { V3kpdv r; r.v = -wi.v; r.ddu = -wi.ddu; r.ddv = -wi.ddv; r.ddl = mk3(0.f); return r; }
            return reflectDerivatives(wi, dsg.n);
        case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
            int lx, ly;
            decodeLobeType(lobe, lx, ly);
            if (lx == 0 && ly == 0) return reflectDerivatives(wi, dsg.n);
            return gratingDerivatives(matGrating(m, sd.uv), dsg, wi, lx, ly, wavelength);
        }
        default: return reflectDerivatives(wi, dsg.n);
    }
}

}  // namespace plt
