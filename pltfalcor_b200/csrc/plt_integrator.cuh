// plt_integrator.cuh — device functions of the PLTPT sample and solve stages: light sampling, NEE / MNEE,
// forward replay of the generalized ray, measurement.  Semantics follow
//   Source/RenderPasses/PLTPT/{Path.slang, pltpt.slangh, pltpt_sample.rt.slang, pltpt_solve.rt.slang,
//   LightSampling.slangh, ShadowRays.slangh, Manifolds.slangh} and Source/Falcor/Rendering/Lights/*.slang (cited per function).
// Layout is ours: the bounce buffer is three SoA planes indexed [bounce][pixel] so that a warp of neighbouring
// pixels reads/writes full 128 B lines (the reference's 40 B AoS record is re-assembled only for parity dumps).
#pragma once
#include "bvh_traverse.cuh"
#include "plt_scene.cuh"

namespace plt {

// Path.slang:47-109
enum : uint32_t { RF_LastBounceWasDelta = 0x1, RF_HasDispersiveDeltaSegment = 0x2, RF_OnlyDelta = 0x4, RF_Miss = 0x10, RF_Terminated = 0x20, RF_MSMask = 0xffff0000u };

struct FrameParams {
    DevScene sc;
    PltptOptions opt;
    float3 camPos, camU, camV, camW;
    float jitterX, jitterY, polarizer, polarizerTheta;
    float viewRow0[3], viewRow1[3], viewRow2[3];
    uint32_t frameCount, width, height, pixelCount;
    // bounce planes: [ (maxBounces+1) * pixelCount ]
    uint4* bHit;       // PackedHitInfo
    float4* bDir;      // encodedWi.xy, asfloat(lobe), asfloat(flags)
    float2* bPdf;      // pathContribution, lastNonDeltaPdf
    float4* color;     // RGBA32F (may be null when accum is used)
    float4* albedo;
    float4* normal;
    double* accum;     // RGBA64F running sums (may be null)
    unsigned long long* counters;  // [0] bsdf evals x wavelengths, [1] recorded vertices, [2] traversal stack overflows, [3] mnee iterations,
                                   // [4] sum L(L-1)/2, [5] BVH nodes visited, [6] triangles tested
    // derived switches: pltpt.slangh:100-122 + PLTPT::getDefines (PLTPT.cpp:216-256)
    bool useEmissive, useAnalytic, doNEE, doMS, doMIS, perTileSG;
};

struct Bounce {
    uint4 hit;
    float3 wi;
    uint32_t lobe, flags;
    float pathContribution, lastNonDeltaPdf;
};
PLT_D void writeBounce(const FrameParams& P, uint32_t pixel, int p, uint4 hit, float3 wi, uint32_t lobe, uint32_t flags, float contribution, float lastNonDeltaPdf) {
    const size_t o = size_t(p) * P.pixelCount + pixel;
    const float2 e = ndir_to_oct_equal_area_snorm(wi);
    P.bHit[o] = hit;
    P.bDir[o] = mk4(e.x, e.y, __uint_as_float(lobe), __uint_as_float(flags));
    P.bPdf[o] = mk2(contribution, lastNonDeltaPdf);
}
PLT_D Bounce readBounce(const FrameParams& P, uint32_t pixel, int p) {
    const size_t o = size_t(p) * P.pixelCount + pixel;
    Bounce b;
    b.hit = P.bHit[o];
    const float4 d = P.bDir[o];
    const float2 q = P.bPdf[o];
    b.wi = oct_to_ndir_equal_area_snorm(mk2(d.x, d.y));
    b.lobe = __float_as_uint(d.z); b.flags = __float_as_uint(d.w);
    b.pathContribution = q.x; b.lastNonDeltaPdf = q.y;
    return b;
}

// ---- rays ------------------------------------------------------------------------------------------------------
constexpr float kShadowRayTMin = 1e-4f;
PLT_D bool traceClosestHit(const FrameParams& P, TraversalCounters& tc, float3 o, float3 d, float tmin, float tmax, TriHit& th, float& t) {
    Hit h;
    const bool hit = traverseBvh<false>(P.sc.bvhNodes, P.sc.bvhTris, o, d, tmin, tmax, h, reinterpret_cast<uint32_t*>(P.counters + 2), &tc);
    th.instanceID = h.instanceID; th.primitiveIndex = h.primitiveIndex; th.u = h.u; th.v = h.v;
    t = h.t;
    return hit;
}
PLT_D bool occluded(const FrameParams& P, TraversalCounters& tc, float3 o, float3 d, float tmin, float tmax) {
    Hit h;
    return traverseBvh<true>(P.sc.bvhNodes, P.sc.bvhTris, o, d, tmin, tmax, h, reinterpret_cast<uint32_t*>(P.counters + 2), &tc);
}
// ShadowRays.slangh:50-52
PLT_D float3 shadowRayOrigin(const SD& sd, float3 rayDir) { return computeRayOrigin(sd.posW, dot(sd.faceN, rayDir) >= .0f ? sd.faceN : -sd.faceN); }

// ---- sample stage ----------------------------------------------------------------------------------------------
struct PathState {  // Path.slang:146-196
    Rng sg;
    float3 origin, dir;
    float wavelength, weight, pdf, thp;
    uint32_t flags;
    int length;
    float lastNonDeltaPdf;
    TraversalCounters tc;
};
// Camera.slang:46-78 (computeRayPinhole)
PLT_D void cameraRay(const FrameParams& P, uint32_t px, uint32_t py, float3& origin, float3& dir) {
    float2 p = mk2((float(px) + 0.5f) / float(P.width), (float(py) + 0.5f) / float(P.height));
    p = p + mk2(-P.jitterX, P.jitterY);
    const float2 ndc = mk2(2.f * p.x - 1.f, -2.f * p.y + 1.f);
    origin = P.camPos;
    dir = normalize(ndc.x * P.camU + ndc.y * P.camV + P.camW);
}
// pltpt.slangh:170-176
PLT_D float russianRoulette(float rrVal, float u) {
    const float p = rrVal * rrVal;  // pow(rrVal, 2)
    return u < p ? .0f : 1.f - p;
}
// pltpt_sample.rt.slang:63-141 (scatterTriangleMeshClosestHit)
PLT_D void scatterClosestHit(const FrameParams& P, uint32_t pixel, PathState& path, const TriHit& th) {
    const uint4 hitInfo = packHit(P.sc, th);
    uint32_t materialID;
    const VertexData v = getVertexData(P.sc, th, materialID);
    const DevMaterial m = loadMaterial(P.sc.materials, materialID);
    const float3 wi = -path.dir;
    const SD sd = prepareShadingData(v, materialID, m, wi);
    if (P.opt.alphaMasking && m.opacity == .0f) {
        path.origin = sd.computeNewRayOrigin(false);
        return;
    }
    const float rcp_thp = P.opt.doRussianRoulette ? russianRoulette(1.f - path.weight, path.sg.next1D()) : 1.f;
    bool shouldEndPath = path.length == int(P.opt.maxBounces) || (rcp_thp == .0f && path.length > 0);
    Sample bs;
    bs.wo = mk3(0.f); bs.pdf = 0.f; bs.lobe = 0; bs.weight = 0.f;
    if (!shouldEndPath) {
        const bool valid = bsdfSample(m, sd, path.sg, bs, path.wavelength, L_All, mk3(0.f));
        shouldEndPath = !(valid && bs.pdf != .0f);
        shouldEndPath |= path.pdf * bs.pdf * path.thp * rcp_thp < 1e-6f;
    }
    const bool isDelta = (bs.lobe & L_Delta) != 0, isDispersive = (bs.lobe & L_DispersiveDelta) != 0, isReflection = (bs.lobe & L_Reflection) != 0;
    const bool isMSCompatible = (isReflection && (m.propFlags & P_CanMSReflective)) || (!isReflection && (m.propFlags & P_CanMSTransmissive));
    if (!isDelta && (path.flags & RF_OnlyDelta)) {
        if (P.albedo) P.albedo[pixel] = mk4(m.albx, m.alby, m.albz, 1.f);
        if (P.normal)
            P.normal[pixel] = mk4(P.viewRow0[0] * sd.N.x + P.viewRow0[1] * sd.N.y + P.viewRow0[2] * sd.N.z, P.viewRow1[0] * sd.N.x + P.viewRow1[1] * sd.N.y + P.viewRow1[2] * sd.N.z,
                                  P.viewRow2[0] * sd.N.x + P.viewRow2[1] * sd.N.y + P.viewRow2[2] * sd.N.z, 1.f);
    }
    if (shouldEndPath) {
        path.flags |= RF_Terminated;
        writeBounce(P, pixel, path.length, hitInfo, wi, 0u, path.flags, 1.f / (path.pdf * path.thp), path.lastNonDeltaPdf);
        return;
    }
    path.weight *= bs.weight;
    path.thp *= rcp_thp;
    path.pdf *= bs.pdf;
    if (isDispersive) path.flags |= RF_HasDispersiveDeltaSegment;
    writeBounce(P, pixel, path.length, hitInfo, wi, bs.lobe, path.flags, 1.f / (path.pdf * path.thp), path.lastNonDeltaPdf);
    ++path.length;
    path.lastNonDeltaPdf = !isDelta ? bs.pdf : path.lastNonDeltaPdf;
    const uint32_t ms = isMSCompatible ? (path.flags >> 16) + 1u : 0u;
    path.flags = (path.flags & ~RF_MSMask) | (ms << 16);
    path.flags = isDelta ? (path.flags | RF_LastBounceWasDelta) : (path.flags & ~RF_LastBounceWasDelta);
    if (!isDelta) path.flags &= ~RF_OnlyDelta;
    path.origin = sd.computeNewRayOrigin(isReflection);
    path.dir = bs.wo;
}
// pltpt_sample.rt.slang:146-185 (raygen)
PLT_D TraversalCounters samplePixel(const FrameParams& P, uint32_t px, uint32_t py) {
    const uint32_t pixel = py * P.width + px;
    PathState path;
    path.tc.nodes = path.tc.tris = 0;
    path.sg = Rng::seed(px, py, P.frameCount);
    path.wavelength = sampleHeroWavelength(path.sg.next1D());
    if (P.normal) P.normal[pixel] = mk4(0.f);
    if (P.albedo) P.albedo[pixel] = mk4(0.f);
    cameraRay(P, px, py, path.origin, path.dir);
    path.weight = path.pdf = path.thp = 1.f;
    path.flags = RF_OnlyDelta;
    path.lastNonDeltaPdf = 0.f;
    int guard = 0;  // bound on alpha-mask pass-throughs, which do not advance `length`
    for (path.length = 0; path.length < int(P.opt.maxBounces) + 1 && (path.flags & (RF_Miss | RF_Terminated)) == 0 && guard < 4096; ++guard) {
        TriHit th;
        float t;
        if (!traceClosestHit(P, path.tc, path.origin, path.dir, 0.f, kFltMax, th, t)) {
            path.flags |= RF_Miss;
            writeBounce(P, pixel, path.length, make_uint4(0, 0, 0, 0), -path.dir, 0u, path.flags, 1.f / (path.pdf * path.thp), path.lastNonDeltaPdf);
        } else scatterClosestHit(P, pixel, path, th);
    }
    return path.tc;
}

// ---- light sampling (LightSampling.slangh, Rendering/Lights/*.slang) ---------------------------------------------
enum { LT_EnvMap = 0, LT_Emissive = 1, LT_Analytic = 2 };
struct LightSample {
    float pdf;
    float3 origin;
    float distance;
    float3 dir;
    Beam beam;
    uint32_t lightType;
};
PLT_D bool anyPositive(float4 L, int hwss) { return L.x > 0.f || (hwss > 1 && L.y > 0.f) || (hwss > 2 && L.z > 0.f) || (hwss > 3 && L.w > 0.f); }
// LightSampling.slangh:40-76 (env maps are not supported: §8f rank 3)
PLT_D void lightTypeProbabilities(const FrameParams& P, float p[3]) {
    p[0] = 0.f; p[1] = P.useEmissive ? 1.f : 0.f; p[2] = P.useAnalytic ? 1.f : 0.f;
    const float sum = p[0] + p[1] + p[2];
    if (sum == 0.f) return;
    const float invSum = 1.f / sum;
    p[0] *= invSum; p[1] *= invSum; p[2] *= invSum;
}
PLT_D bool selectLightType(const FrameParams& P, uint32_t& lightType, float& pdf, Rng& sg) {
    float p[3];
    lightTypeProbabilities(P, p);
    float u = sg.next1D();
    for (lightType = 0; lightType < 3; ++lightType) {
        if (u < p[lightType]) { pdf = p[lightType]; return true; }
        u -= p[lightType];
    }
    return false;
}
struct EmissiveTri { float3 p0, p1, p2, normal; float area; uint32_t materialID; };
// LightCollectionShared.slang:118-135
PLT_D EmissiveTri loadEmissiveTriangle(const DevScene& sc, uint32_t idx) {
    const float4* p = reinterpret_cast<const float4*>(sc.etris + idx);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    EmissiveTri t;
    t.p0 = mk3(a.x, a.y, a.z); t.p1 = mk3(b.x, b.y, b.z); t.p2 = mk3(c.x, c.y, c.z);
    t.normal = decodeNormal2x16(__float_as_uint(d.x));
    t.area = d.y;  // stored as fp32 bits
    t.materialID = __float_as_uint(d.z);
    return t;
}
// EmissiveLightSamplerHelpers.slang:56-74: emission = profile spectrum x rgbToSpectrum(1,1,1)
PLT_D void materialEvalEmissive(const FrameParams& P, uint32_t materialID, float wavelength, bool importanceSample, Rng& sg, float4& Le, float4& wl) {
    const uint32_t id = __ldg(&P.sc.materials[materialID].emissionId);
    if (id != PLTPT_NO_SPECTRAL_PROFILE) {
        Le = sampleSpectrum(P.sc.profiles, id, wavelength, sg, importanceSample, int(P.opt.HWSS), wl);
        Le *= rgbToSpectrum4(mk3(1.f), wl);
    }
}
struct TriangleLightSample { float3 posW, normalW, dir; float distance; float4 Le, wavelengths; float pdf; };
// EmissiveLightSamplerHelpers.slang:86-128
PLT_D bool sampleTriangle(const FrameParams& P, float3 posW, uint32_t triangleIndex, Rng& sg, float wavelength, bool importanceSample, TriangleLightSample& ls) {
    ls.Le = mk4(0.f); ls.wavelengths = mk4(0.f); ls.pdf = 0.f;
    const EmissiveTri tri = loadEmissiveTriangle(P.sc, triangleIndex);
    const float2 u = sg.next2D();
    const float3 b = sample_triangle(u);
    ls.posW = tri.p0 * b.x + tri.p1 * b.y + tri.p2 * b.z;
    const float3 toLight = ls.posW - posW;
    const float distSqr = fmaxf(kFltMin, dot(toLight, toLight));
    ls.distance = sqrtf(distSqr);
    ls.dir = toLight / ls.distance;
    ls.normalW = tri.normal;
    const float cosTheta = dot(ls.normalW, -ls.dir);
    if (cosTheta <= 0.f) return false;
    materialEvalEmissive(P, tri.materialID, wavelength, importanceSample, sg, ls.Le, ls.wavelengths);
    ls.pdf = distSqr / fmaxf(kFltMin, cosTheta * tri.area);
    return true;
}
// EmissivePowerSampler.slang:64-90 / EmissiveUniformSampler.slang:56-72
PLT_D bool emissiveSampleLight(const FrameParams& P, float3 posW, Rng& sg, float wavelength, bool importanceSample, TriangleLightSample& ls) {
    const DevScene& sc = P.sc;
    uint32_t triangleIndex;
    float selectionPdf;
    if (P.opt.emissiveSampler == 1) {
        if (sc.etriCount == 0) return false;
        const float uLight = sg.next1D();
        triangleIndex = min(uint32_t(uLight * sc.etriCount), sc.etriCount - 1);
        const uint2 packed = __ldg(sc.alias + triangleIndex);
        const float threshold = h2f(packed.x >> 16u);
        const uint32_t selectAbove = ((packed.x & 0xFFFFu) << 8u) | ((packed.y >> 24u) & 0xFFu);
        const uint32_t selectBelow = packed.y & 0xFFFFFFu;
        triangleIndex = (sg.next1D() >= threshold) ? selectAbove : selectBelow;
        selectionPdf = __ldg(&sc.eflux[triangleIndex].flux) * sc.invWeightsSum;
    } else {
        if (sc.activeTriCount == 0) return false;
        const float uLight = sg.next1D();
        const uint32_t idx = min(uint32_t(uLight * sc.activeTriCount), sc.activeTriCount - 1);
        triangleIndex = __ldg(sc.activeTris + idx);
        selectionPdf = 1.f / float(sc.activeTriCount);
    }
    if (!sampleTriangle(P, posW, triangleIndex, sg, wavelength, importanceSample, ls)) return false;
    ls.pdf *= selectionPdf;
    return true;
}
// EmissivePowerSampler.slang:98-126, EmissiveLightSamplerHelpers.slang:135-151
PLT_D float emissiveEvalPdf(const FrameParams& P, float3 posW, uint32_t triangleIndex, float3 hitPosW, float3 hitNormalW) {
    const DevScene& sc = P.sc;
    if (triangleIndex == 0xffffffffu) return 0.f;
    const float sel = P.opt.emissiveSampler == 1 ? __ldg(&sc.eflux[triangleIndex].flux) * sc.invWeightsSum : (sc.activeTriCount > 0 ? 1.f / float(sc.activeTriCount) : 0.f);
    const float3 toLight = hitPosW - posW;
    const float distSqr = dot(toLight, toLight);
    if (distSqr <= kFltMin) return 0.f;
    const float3 L = toLight / sqrtf(distSqr);
    const float cosTheta = dot(hitNormalW, -L);
    if (cosTheta <= 0.f) return 0.f;
    const float area = __uint_as_float(__ldg(&sc.etris[triangleIndex].area));
    return sel * (distSqr / fmaxf(kFltMin, cosTheta * area));
}
// LightSampling.slangh:104-126
PLT_D bool generateEmissiveSample(const FrameParams& P, const SD& sd, Rng& sg, float wavelength, bool importanceSample, LightSample& ls) {
    if (!P.useEmissive) return false;
    TriangleLightSample tls;
    if (!emissiveSampleLight(P, sd.posW, sg, wavelength, importanceSample, tls)) return false;
    const float4 L = tls.pdf > 0.f ? tls.Le / tls.pdf : mk4(.0f);
    ls.pdf = tls.pdf;
    const float3 lightPos = computeRayOrigin(tls.posW, tls.normalW);
    ls.origin = shadowRayOrigin(sd, lightPos - sd.posW);
    const float3 toLight = lightPos - ls.origin;
    ls.distance = length(toLight);
    ls.dir = normalize(toLight);
    ls.beam = sourceBeamArea(tls.posW, -ls.dir, P.opt.sourcingAreaFromEmissiveGeometry, ls.distance, L, tls.wavelengths, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    return anyPositive(L, int(P.opt.HWSS));
}
// LightHelpers.slang:70-330 (analytic lights)
struct AnalyticSample { float3 posW, normalW, dir; float distance; float4 Li, wavelengths; float pdf; };
PLT_D bool sampleAnalyticLight(const FrameParams& P, float3 shadingPosW, const PltptLightData* light, Rng& sg, float wavelength, bool importanceSample, AnalyticSample& ls) {
    ls.Li = mk4(0.f); ls.wavelengths = mk4(0.f); ls.pdf = 0.f; ls.distance = 0.f;
    const float3 lightPos = mk3(__ldg(&light->posW[0]), __ldg(&light->posW[1]), __ldg(&light->posW[2]));
    const float3 lightDir = mk3(__ldg(&light->dirW[0]), __ldg(&light->dirW[1]), __ldg(&light->dirW[2]));
    const uint32_t type = __ldg(&light->type), specId = __ldg(&light->intensitySpectrumId);
    const int hwss = int(P.opt.HWSS);
    switch (type) {
        case 0: {  // Point (pdf 0: later rejected by ls.pdf > 1e-6, reference quirk SURVEY §8 a10)
            ls.posW = lightPos;
            ls.normalW = lightDir;
            const float3 toLight = ls.posW - shadingPosW;
            const float distSqr = fmaxf(dot(toLight, toLight), 1e-9f);
            ls.distance = sqrtf(distSqr);
            ls.dir = toLight / ls.distance;
            const float cosTheta = -dot(ls.dir, lightDir);
            float falloff = 1.f;
            const float cosOpening = __ldg(&light->cosOpeningAngle), penumbra = __ldg(&light->penumbraAngle);
            if (cosTheta < cosOpening) falloff = 0.f;
            else if (penumbra > 0.f) {
                const float deltaAngle = __ldg(&light->openingAngle) - acosf(cosTheta);
                const float tt = clampf(deltaAngle / penumbra, 0.f, 1.f);
                falloff = tt * tt * (3.f - 2.f * tt);
            }
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.Li *= falloff / distSqr;
            ls.pdf = 0.f;
            return true;
        }
        case 1:  // Directional
            ls.posW = mk3(INFINITY);
            ls.normalW = lightDir;
            ls.distance = kFltMax;
            ls.dir = -lightDir;
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.pdf = 0.f;
            return true;
        case 2: {  // Distant
            ls.posW = mk3(INFINITY);
            const float2 u = sg.next2D();
            const float cosSub = __ldg(&light->cosSubtendedAngle);
            const float3 dir = sample_cone(u, cosSub);
            ls.dir = normalize(mulVec(light->transMat, dir));
            ls.normalW = -ls.dir;
            ls.distance = kFltMax;
            const float solidAngle = k2Pi * (1.f - cosSub);
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.pdf = 1.f / solidAngle;
            return true;
        }
        case 3: case 4: case 5: {  // Rect, Disc, Sphere
            const float2 u = sg.next2D();
            float3 pos, nrm = mk3(0.f, 0.f, 1.f);
            if (type == 3) pos = mk3(u.x * 2.f - 1.f, u.y * 2.f - 1.f, 0.f);
            else if (type == 4) { const float2 d = sample_disk(u); pos = mk3(d.x, d.y, 0.f); }
            else { pos = sample_sphere(u); nrm = pos; }
            ls.posW = mulPoint(light->transMat, pos);
            ls.normalW = normalize(mulVec(light->transMatIT, nrm));
            // finalizeAreaLightSample
            const float3 toLight = ls.posW - shadingPosW;
            const float distSqr = fmaxf(dot(toLight, toLight), 1e-9f);
            ls.distance = sqrtf(distSqr);
            ls.dir = toLight / ls.distance;
            const float cosTheta = dot(ls.normalW, -ls.dir);
            if (cosTheta <= 0.f) return false;
            const float area = __ldg(&light->surfaceArea);
            ls.Li = sampleSpectrum(P.sc.profiles, specId, wavelength, sg, importanceSample, hwss, ls.wavelengths);
            ls.Li *= area * cosTheta / distSqr;
            ls.pdf = distSqr / (cosTheta * area);
            return true;
        }
        default: return false;
    }
}
// LightSampling.slangh:129-166
PLT_D bool generateAnalyticLightSample(const FrameParams& P, const SD& sd, Rng& sg, float wavelength, bool importanceSample, LightSample& ls) {
    const uint32_t lightCount = P.sc.lightCount;
    if (!P.useAnalytic || lightCount == 0) return false;
    const uint32_t lightIndex = min(uint32_t(sg.next1D() * lightCount), lightCount - 1);
    const PltptLightData* light = P.sc.lights + lightIndex;
    AnalyticSample s;
    if (!sampleAnalyticLight(P, sd.posW, light, sg, wavelength, importanceSample, s)) return false;
    const uint32_t type = __ldg(&light->type);
    const bool isDistantEmitter = type == 1 || type == 2;
    const float A = __ldg(&light->surfaceArea);
    const float Omega = 4.f * kPi * (1.f - sqr(__ldg(&light->cosSubtendedAngle)));
    const float4 L = s.Li * float(lightCount);
    ls.pdf = s.pdf / lightCount;
    ls.origin = shadowRayOrigin(sd, s.dir);
    ls.distance = s.distance;
    ls.dir = s.dir;
    if (isDistantEmitter) ls.beam = sourceBeamDistant(-ls.dir, Omega, L, s.wavelengths, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    else ls.beam = sourceBeamArea(mk3(__ldg(&light->posW[0]), __ldg(&light->posW[1]), __ldg(&light->posW[2])), -ls.dir, A, ls.distance, L, s.wavelengths, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    return anyPositive(L, int(P.opt.HWSS));
}
PLT_D void clearLightSample(LightSample& ls) {
    ls.pdf = 0.f; ls.origin = mk3(0.f); ls.distance = 0.f; ls.dir = mk3(0.f); ls.lightType = 0;
    ls.beam = makeBeam(mk4(0.f), 0.f, mk3(0.f), mk3(0.f, 0.f, 1.f), mk4(0.f), false, 4);
    ls.beam.r = mk3(0.f); ls.beam.t = mk3(0.f);
}
// LightSampling.slangh:168-206
PLT_D bool generateLightSample(const FrameParams& P, const SD& sd, bool sampleReflection, bool sampleTransmission, Rng& sg, Rng& tilesg, float wavelength, bool importanceSample, LightSample& ls) {
    clearLightSample(ls);
    uint32_t lightType;
    float selectionPdf;
    const bool success = P.perTileSG ? selectLightType(P, lightType, selectionPdf, tilesg) : selectLightType(P, lightType, selectionPdf, sg);
    if (!success) return false;
    bool valid = false;
    if (P.useEmissive && lightType == LT_Emissive) valid = generateEmissiveSample(P, sd, sg, wavelength, importanceSample, ls);
    if (P.useAnalytic && lightType == LT_Analytic) valid = generateAnalyticLightSample(P, sd, sg, wavelength, importanceSample, ls);
    ls.lightType = lightType;
    ls.pdf *= selectionPdf;
    ls.beam.scale(1.f / selectionPdf);
    const bool isReflection = dot(sd.faceN, ls.dir) * dot(sd.faceN, sd.V) > .0f;
    return valid && ls.pdf > 1e-6f && !(!sampleReflection && isReflection) && !(!sampleTransmission && !isReflection);
}

// ---- solve stage ---------------------------------------------------------------------------------------------
struct SolveState {
    Rng sg, tilesg;
    float3 Lrgb, pathOrigin;
    float wavelength;
    uint32_t evals;
    TraversalCounters tc;
};
// pltpt_solve.rt.slang:48-66 (forwardUnwindStep)
PLT_D void forwardUnwindStep(const FrameParams& P, SolveState& st, const VertexData& v, uint32_t materialID, float3 wo, uint32_t lobe, Beam& beam) {
    const DevMaterial m = loadMaterial(P.sc.materials, materialID);
    const SD sd = prepareShadingData(v, materialID, m, -beam.r);
    beam.propagate(sd.posW);
    const BsdfCtx cx = {P.sc.profiles, int(P.opt.HWSS)};
    bsdfEval(m, cx, sd, wo, lobe, beam);
    st.evals += uint32_t(beam.comps);
    if (P.opt.HWSSDoMIS) beam.spdf *= bsdfLambdaPdf(m, cx, sd, wo, lobe, beam, beam.wl);
}
// pltpt_solve.rt.slang:68-79
PLT_D void forwardUnwind(const FrameParams& P, SolveState& st, uint32_t pixel, int bounceIdx, Beam& beam) {
    for (int p = bounceIdx - 1; p >= 0; --p) {
        const Bounce b = readBounce(P, pixel, p);
        const TriHit th = unpackHit(P.sc, b.hit);
        uint32_t materialID;
        const VertexData v = getVertexData(P.sc, th, materialID);
        forwardUnwindStep(P, st, v, materialID, b.wi, b.lobe, beam);
    }
}
// pltpt_solve.rt.slang:96-147
PLT_D void measure(const FrameParams& P, SolveState& st, int bounceIdx, float pathContribution, Beam& beam) {
    if (!P.opt.useDirectLights && bounceIdx <= 1) return;
    beam.propagate(P.camPos);
    if (P.polarizer > .0f) {
        beam.rotateFrame(normalize(P.camU), -normalize(P.camW));
        float s, c;
        sincosf(P.polarizerTheta, &s, &c);
        const float a = 1.f - P.polarizer, h = .5f * P.polarizer;
        for (int j = 0; j < beam.comps; ++j) {
            const float4 SP = beam.getSP(j);
            // (1-p) I + p * MuellerPolarizer(theta), Mueller.slang linear polariser
            beam.setSP(j, mk4(a * SP.x + h * (SP.x + c * SP.y + s * SP.z), a * SP.y + h * (c * SP.x + c * c * SP.y + c * s * SP.z),
                              a * SP.z + h * (s * SP.x + c * s * SP.y + s * s * SP.z), a * SP.w));
        }
    }
    if (P.opt.HWSSDoMIS) {
        float sum = 0.f;
        for (int j = 0; j < int(P.opt.HWSS); ++j) sum += lane(beam.spdf, j);
        pathContribution *= sum > kFltEps ? beam.spdf.x / sum : .0f;
    } else pathContribution *= 1.f / float(beam.comps);
    const float3 Lrgb = vmax(mk3(.0f), spectrumToRgb(beam.L, beam.wl, beam.comps) * pathContribution);
    if (isfinite(Lrgb.x) && isfinite(Lrgb.y) && isfinite(Lrgb.z)) st.Lrgb += Lrgb;
}
// pltpt_solve.rt.slang:205-217
PLT_D bool wasNEEAvailable(const FrameParams& P, int bounceIdx, uint32_t flags) {
    return P.doNEE && bounceIdx > 0 && (!(flags & RF_LastBounceWasDelta) || (P.doMS && (flags >> 16) <= P.opt.MNEEMaxOccluders && bounceIdx > 1));
}
PLT_D float misWeight(float p0, float p1) { return p0 / (p0 + p1); }

// pltpt_solve.rt.slang:152-200
PLT_D void evalEmissive(const FrameParams& P, SolveState& st, uint32_t pixel, int bounceIdx, const Bounce& bounce, const DevMaterial& m, const SD& sd, const TriHit& th) {
    const bool didNEE = wasNEEAvailable(P, bounceIdx, bounce.flags);
    if (!(P.useEmissive && (m.propFlags & P_IsEmissive) && (P.doMIS || !didNEE))) return;
    const bool needMIS = P.doMIS && didNEE;
    const uint32_t triIdx = getEmissiveTriangleIndex(P.sc, th.instanceID, th.primitiveIndex);
    const float3 hitN = sd.frontFacing ? sd.faceN : -sd.faceN;
    float p[3];
    lightTypeProbabilities(P, p);
    const float lightPdf = p[1] * emissiveEvalPdf(P, st.pathOrigin, triIdx, sd.posW, hitN);
    const float w = needMIS ? misWeight(bounce.lastNonDeltaPdf, lightPdf) : 1.f;
    float4 wl = mk4(.0f);
    float4 Le = sampleSpectrum(P.sc.profiles, m.emissionId, st.wavelength, st.sg, P.opt.doImportanceSampleEmitters != 0, int(P.opt.HWSS), wl);
    Le *= w * rgbToSpectrum4(mk3(1.f), wl);
    if (!anyPositive(Le, int(P.opt.HWSS))) return;
    const float dist = length(st.pathOrigin - sd.posW);
    Beam beam = sourceBeamArea(sd.posW, bounce.wi, P.opt.sourcingAreaFromEmissiveGeometry, dist, Le, wl, P.opt.sourcingMaxBeamOmega, int(P.opt.HWSS));
    if (!P.opt.HWSSDoMIS && (bounce.flags & RF_HasDispersiveDeltaSegment)) beam.comps = 1;
    forwardUnwind(P, st, pixel, bounceIdx, beam);
    measure(P, st, bounceIdx, bounce.pathContribution, beam);
}

// ---- MNEE (Manifolds.slangh) ------------------------------------------------------------------------------------
constexpr int kMaxMSHits = 5;  // MNEEMaxOccluders (<= 4, validated in pltpt_set_options) + the current vertex
struct MSHit { uint4 hit; float3 posW; uint32_t lobe; };
struct ManifoldSampleData {
    MSHit hits[kMaxMSHits];
    float3 woLastHitToLight;
    int hitCount;
    uint32_t iterations;
    float wavelength, pdf;
    bool dispersive;
};
// Manifolds.slangh:95-108
PLT_D bool manifoldSampleVertex(uint4 hitInfo, const SD& sd, const DevMaterial& m, float3 lightDir, Rng& sg, ManifoldSampleData& data) {
    Sample s;
    if (!bsdfSample(m, sd, sg, s, data.wavelength, L_Delta, lightDir)) return false;
    if (data.hitCount >= kMaxMSHits) return false;
    data.pdf *= s.pdf;
    data.hits[data.hitCount].lobe = s.lobe;
    data.hits[data.hitCount].hit = hitInfo;
    data.hits[data.hitCount].posW = sd.posW;
    data.dispersive |= (s.lobe & L_DispersiveDelta) != 0;
    ++data.hitCount;
    return true;
}
// Manifolds.slangh:110-160
PLT_D bool traceMNEEOccluders(const FrameParams& P, TraversalCounters& tc, float3 rayOrigin, float3 rayDir, float rayTMin, float rayTMax, Rng& sg, int maxMSBounces, ManifoldSampleData& data) {
    const float3 wi = -rayDir;
    for (;;) {
        TriHit th;
        float t;
        if (!traceClosestHit(P, tc, rayOrigin, rayDir, rayTMin, rayTMax, th, t)) return true;
        if (data.hitCount >= int(P.opt.MNEEMaxOccluders) || data.hitCount >= maxMSBounces) return false;
        uint32_t materialID;
        VertexData v = getVertexData(P.sc, th, materialID);
        v.posW = rayOrigin + t * rayDir;
        const DevMaterial m = loadMaterial(P.sc.materials, materialID);
        const SD sd = prepareShadingData(v, materialID, m, wi);
        if (!(m.propFlags & P_CanMSTransmissive)) return false;
        if (!manifoldSampleVertex(packHit(P.sc, th), sd, m, -wi, sg, data) || (data.hits[data.hitCount - 1].lobe & L_Transmission) == 0) return false;
        rayOrigin = shadowRayOrigin(sd, rayDir);
        rayTMax -= t;
    }
}
// Manifolds.slangh:284-347
PLT_D float2 sphcoords(float3 w) {
    const float theta = acosf(clampf(w.z, -1.f, 1.f));
    float phi = atan2f(w.y, w.x);
    phi += phi < .0f ? k2Pi : .0f;
    return mk2(theta, phi);
}
PLT_D M22 sphDeriv(float3 v, float3 ddu, float3 ddv) {
    const float dAcos = -1.f / sqrtf(fmaxf(.0f, 1.f - sqr(v.z)));
    const float invNorm = 1.f / (v.x * v.x + v.y * v.y);
    return {dAcos * ddu.z, dAcos * ddv.z, invNorm * (v.x * ddu.y - v.y * ddu.x), invNorm * (v.x * ddv.y - v.y * ddv.x)};
}
PLT_D bool manifoldOneStep(const V3pdv& w, const V3kpdv& ws, const V3pdv& p, float& lenC, float3& positionStepDir) {
    const float2 spho = sphcoords(w.v), sphio = sphcoords(ws.v);
    const float dt = spho.x - sphio.x;
    float dp = spho.y - sphio.y;
    dp = fabsf(dp) > kPi ? dp - signf(dp) * k2Pi : dp;
    const float2 C = mk2(dt, dp);
    lenC = sqrtf(dot(C, C));
    const M22 dO = sphDeriv(w.v, w.ddu, w.ddv), dIO = sphDeriv(ws.v, ws.ddu, ws.ddv);
    const M22 dCdX = {dO.a - dIO.a, dO.b - dIO.b, dO.c - dIO.c, dO.d - dIO.d};
    M22 inv;
    if (!tryInvert(dCdX, inv)) return false;
    const float2 dX = mul(inv, C);
    positionStepDir = p.ddu * dX.x + p.ddv * dX.y;
    return true;
}
// Manifolds.slangh:165-281.  Kept out of line: its data-dependent Newton loop must not inflate the register budget of
// the straight NEE path.
__device__ __noinline__ bool manifoldNEESample(const FrameParams& P, TraversalCounters& tc, float3 visOrigin, float3 visDir, float visTMax, float3 origin, float3 lightPos, float3 lightDir,
                                               bool isDistantLight, int maxMSBounces, Rng& sg, ManifoldSampleData& data) {
    if (!traceMNEEOccluders(P, tc, visOrigin, visDir, 0.f, visTMax, sg, maxMSBounces, data)) return false;
    if (data.hitCount == 0) return true;
    float beta[kMaxMSHits];
    for (int i = 0; i < kMaxMSHits; ++i) beta[i] = 1.f;
    bool success = false;
    while (!success) {
        success = true;
        for (int idx = 0; idx < data.hitCount; ++idx) {
            const TriHit th = unpackHit(P.sc, data.hits[idx].hit);
            uint32_t materialID;
            const VertexData v = getVertexData(P.sc, th, materialID);
            const DiffGeom dsg = differentialShadingGeometry(P.sc, th);
            const float3 p0 = idx == 0 ? origin : data.hits[idx - 1].posW;
            const V3pdv pdvp0 = dsg.pdw(p0);
            const float3 p2 = idx + 1 == data.hitCount ? lightPos : data.hits[idx + 1].posW;
            V3pdv pdvp2;
            pdvp2.v = lightDir; pdvp2.ddu = pdvp2.ddv = mk3(0.f);
            if (!isDistantLight) pdvp2 = dsg.pdw(p2);
            const bool woIsIn = dot(pdvp2.v, dsg.n.v) < .0f;
            const V3pdv wiP = woIsIn ? pdvp0 : pdvp2;
            const V3pdv woP = woIsIn ? pdvp2 : pdvp0;
            const DevMaterial m = loadMaterial(P.sc.materials, materialID);
            const SD sd = prepareShadingData(v, materialID, m, wiP.v);
            const V3kpdv wos = bsdfPartialDerivatives(m, sd, wiP, dsg, data.hits[idx].lobe, data.wavelength);
            float lenC;
            float3 positionStepDir;
            if (allZero(wos.v)) return false;
            if (!manifoldOneStep(woP, wos, dsg.p, lenC, positionStepDir)) return false;
            if (lenC < P.opt.MNEESolverThreshold) continue;
            success = false;
            for (;;) {
                if (++data.iterations > P.opt.MNEEMaxIterations) return false;
                const float3 proposal = data.hits[idx].posW - beta[idx] * positionStepDir;
                float3 dirToProposal = proposal - p0;
                const float dist = length(dirToProposal);
                dirToProposal /= dist;
                TriHit ph;
                float pt;
                const bool isHit = traceClosestHit(P, tc, p0, dirToProposal, fmaxf(kShadowRayTMin, .99f * dist), 1.01f * dist, ph, pt);
                if (!isHit || th.instanceID != ph.instanceID) { beta[idx] /= 4.f; continue; }
                data.hits[idx].hit = packHit(P.sc, ph);
                data.hits[idx].posW = p0 + dirToProposal * pt;
                beta[idx] = fminf(1.0f, 2.f * beta[idx]);
                break;
            }
        }
    }
    const float3 p = data.hits[data.hitCount - 1].posW;
    float3 dir;
    float dist;
    if (!isDistantLight) { dir = lightPos - p; dist = length(dir); dir /= dist; }
    else { dir = lightDir; dist = kFltMax; }
    data.woLastHitToLight = dir;
    return !occluded(P, tc, p, dir, kShadowRayTMin, dist);
}
// Manifolds.slangh:80-93
PLT_D void manifoldForwardUnwind(const FrameParams& P, SolveState& st, const ManifoldSampleData& r, float3 origin, Beam& beam) {
    for (int idx = r.hitCount - 1; idx >= 0; --idx) {
        const TriHit th = unpackHit(P.sc, r.hits[idx].hit);
        uint32_t materialID;
        VertexData v = getVertexData(P.sc, th, materialID);
        v.posW = r.hits[idx].posW;
        const float3 p2 = idx == 0 ? origin : r.hits[idx - 1].posW;
        forwardUnwindStep(P, st, v, materialID, normalize(p2 - v.posW), r.hits[idx].lobe, beam);
    }
}

// pltpt_solve.rt.slang:219-366 (NEE + evalNEE)
PLT_D void nee(const FrameParams& P, SolveState& st, uint32_t pixel, int bounceIdx, const Bounce& bounce, const SD& sd, const VertexData& V, const DevMaterial& m) {
    const uint32_t lobes = m.lobes;
    const bool allowMSOnCurrentVertex = P.doMS && bounceIdx > 0 && !(bounce.flags & RF_LastBounceWasDelta) && (m.propFlags & P_CanMSReflective);
    const bool sampleReflection = (lobes & L_NonDeltaReflection) != 0 || allowMSOnCurrentVertex;
    const bool sampleTransmission = (lobes & L_NonDeltaTransmission) != 0;
    if (!sampleReflection && !sampleTransmission) return;
    LightSample ls;
    const bool valid = generateLightSample(P, sd, sampleReflection, sampleTransmission, st.sg, st.tilesg, st.wavelength, P.opt.doImportanceSampleEmitters != 0, ls);
    if (!valid || !anyPositive(ls.beam.L, int(P.opt.HWSS))) return;
    if (!P.opt.HWSSDoMIS && (bounce.flags & RF_HasDispersiveDeltaSegment)) ls.beam.comps = 1;
    const bool doMis = P.doMIS && ls.lightType != LT_Analytic;
    bool evalCurrentVertex = true;
    const int maxMSBounces = int(P.opt.maxBounces) - bounceIdx;
    if (!P.doMS || maxMSBounces <= 0) {
        if (occluded(P, st.tc, ls.origin, ls.dir, .0f, ls.distance)) return;
        if (doMis) ls.beam.scale(misWeight(ls.pdf, bsdfPdf(m, sd, ls.dir, ls.beam.wl.x)));
    } else {
        const float3 lightDir = -ls.beam.r;
        float MISBsdfPdf = 0.f;
        ManifoldSampleData ms;
        ms.hitCount = 0; ms.iterations = 0; ms.pdf = 1.f; ms.dispersive = false; ms.woLastHitToLight = mk3(0.f);
        ms.wavelength = ls.beam.wl.x;
        const bool isReflection = dot(sd.N, lightDir) * dot(sd.N, sd.V) > .0f;
        const bool currentVertexMS = isReflection && allowMSOnCurrentVertex;
        if (currentVertexMS) {
            if (!manifoldSampleVertex(bounce.hit, sd, m, lightDir, st.sg, ms) || (ms.hits[0].lobe & L_Reflection) == 0) return;
        }
        const float3 origin = !currentVertexMS ? V.posW : st.pathOrigin;
        const bool ok = manifoldNEESample(P, st.tc, ls.origin, ls.dir, ls.distance, origin, ls.beam.posW, lightDir, ls.beam.distant, maxMSBounces, st.sg, ms);
        if (P.counters && ms.iterations) atomicAdd(P.counters + 3, (unsigned long long)ms.iterations);
        if (!ok || ms.pdf < 1e-6f) return;
        if (ms.hitCount > 0) {
            if (!P.opt.HWSSDoMIS && ms.dispersive) ls.beam.comps = 1;
            if (!ls.beam.distant) {
                ls.dir = ms.woLastHitToLight;
                ls.beam.r = -ms.woLastHitToLight;
                ls.beam.t = perp_stark(ls.beam.r);
            }
            ls.beam.scale(1.f / ms.pdf);
            manifoldForwardUnwind(P, st, ms, origin, ls.beam);
            if (currentVertexMS) { evalCurrentVertex = false; MISBsdfPdf = bounce.lastNonDeltaPdf; }
        }
        if (doMis) {
            if (ms.hitCount == 0 || !currentVertexMS) MISBsdfPdf = bsdfPdf(m, sd, ls.dir, ls.beam.wl.x);
            ls.beam.scale(misWeight(ls.pdf, MISBsdfPdf * ms.pdf));
        }
    }
    // evalNEE (pltpt_solve.rt.slang:219-243)
    Beam& beam = ls.beam;
    if (evalCurrentVertex) {
        const SD forwardSd = prepareShadingData(V, sd.materialID, m, -beam.r);
        beam.propagate(forwardSd.posW);
        const BsdfCtx cx = {P.sc.profiles, int(P.opt.HWSS)};
        bsdfEval(m, cx, forwardSd, bounce.wi, L_NonDelta, beam);
        st.evals += uint32_t(beam.comps);
        if (P.opt.HWSSDoMIS) beam.spdf *= bsdfLambdaPdf(m, cx, forwardSd, bounce.wi, L_NonDelta, beam, beam.wl);
    }
    forwardUnwind(P, st, pixel, bounceIdx, beam);
    measure(P, st, bounceIdx, bounce.pathContribution, beam);
}

// pltpt_solve.rt.slang:401-476 (raygen).  Returns the number of recorded vertices; `evals` = eval calls x wavelengths.
PLT_D uint32_t solvePixel(const FrameParams& P, uint32_t px, uint32_t py, uint32_t& evals, TraversalCounters& tc) {
    const uint32_t pixel = py * P.width + px;
    SolveState st;
    st.sg = Rng::seed(px, py, P.frameCount);
    st.tilesg = Rng::seed(px / P.opt.tileSize, py / P.opt.tileSize, P.frameCount << 2);
    st.wavelength = sampleHeroWavelength(st.sg.next1D());
    st.Lrgb = mk3(0.f);
    st.pathOrigin = P.camPos;
    st.evals = 0;
    st.tc.nodes = st.tc.tris = 0;
    int p = 0;
    for (;; ++p) {
        const Bounce bounce = readBounce(P, pixel, p);
        if (bounce.flags & RF_Miss) break;  // onMiss handles env maps only (pltpt_solve.rt.slang:371-396); no env map support yet
        const TriHit th = unpackHit(P.sc, bounce.hit);
        uint32_t materialID;
        const VertexData v = getVertexData(P.sc, th, materialID);
        const DevMaterial m = loadMaterial(P.sc.materials, materialID);
        const SD sd = prepareShadingData(v, materialID, m, bounce.wi);
        if (P.doNEE) nee(P, st, pixel, p, bounce, sd, v, m);
        evalEmissive(P, st, pixel, p, bounce, m, sd, th);
        if ((bounce.flags & RF_Terminated) || p == int(P.opt.maxBounces)) { ++p; break; }
        st.pathOrigin = v.posW;
    }
    if (P.color) P.color[pixel] = mk4(st.Lrgb.x, st.Lrgb.y, st.Lrgb.z, 1.f);
    if (P.accum) {
        double* a = P.accum + 4 * size_t(pixel);
        a[0] += double(st.Lrgb.x); a[1] += double(st.Lrgb.y); a[2] += double(st.Lrgb.z); a[3] += 1.0;
    }
    evals = st.evals;
    tc = st.tc;
    return uint32_t(p);
}

}  // namespace plt
