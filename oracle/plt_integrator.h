// ORACLE — TEST INFRASTRUCTURE ONLY (see plt_math.h header).  Pinned to the reference's shader text function by
// function: camera rays, Russian roulette, light-type selection and whole NEE light samples (analytic, emissive uniform / power, environment),
// emissive-triangle and vertex decoding, shading data, measure / MISWeight, the manifold solver's geometry and Newton step (oracle/ref_compare.cpp,
// tests/test_oracle_ref.py).  The sample stage's closest-hit and miss programs are pinned WHOLE (path
// book-keeping, bounce records, guides).  The ORDER in which the solve stage's NEE / evalEmissive / manifold-walk programs call the pinned functions is
// hand-read (they interleave TraceRay with the bounce buffer) and is checked against closed-form optics and radiometric laws instead (tests/test_oracle_physics.py, _radiometry.py).
// plt_integrator.h — the PLTPT sample + solve stages, light sampling, NEE / MNEE, measurement.
//   Source/RenderPasses/PLTPT/{Path.slang, pltpt.slangh, pltpt_sample.rt.slang, pltpt_solve.rt.slang,
//   LightSampling.slangh, ShadowRays.slangh, Manifolds.slangh}
#pragma once
#include "plt_scene.h"
#include "plt_debug.h"

namespace orc {

// Path.slang:47-109
struct RayFlags {
    enum : uint32_t { lastBounceWasDeltaF = 0x1, hasDispersiveDeltaSegmentF = 0x2, onlyDeltaF = 0x4, missF = 0x10, terminatedF = 0x20, MSMask = 0xffff0000u };
    uint32_t flags = 0;
    uint32_t MSCompatibleBounces() const { return flags >> 16; }
    void setMSCompatibleBounces(uint32_t v) { flags &= ~MSMask; flags |= v << 16; }
    bool get(uint32_t f) const { return (flags & f) != 0; }
    void set(uint32_t f, bool v) { if (v) flags |= f; else flags &= ~f; }
    bool lastBounceWasDelta() const { return get(lastBounceWasDeltaF); }
    bool dispersiveDeltaSegment() const { return get(hasDispersiveDeltaSegmentF); }
    bool onlyDelta() const { return get(onlyDeltaF); }
    bool miss() const { return get(missF); }
    bool terminated() const { return get(terminatedF); }
    bool valid() const { return !miss() && !terminated(); }
};

struct Path {  // Path.slang:146-196
    SampleGenerator sg;
    float3 origin, dir, normal;
    float wavelength = 0, weight = 1, pdf = 1, thp_modifier = 1;
    uint32_t id = 0;
    int length = 0;
    RayFlags flags;
    float lastNonDeltaPdf = 0;
};

struct Integrator {
    const Scene* scene = nullptr;
    PltptOptions opt{};
    PltptCameraData cam{};
    uint32_t frameCount = 0, width = 0, height = 0;
    // outputs
    PltptBounce* bounces = nullptr;  // (maxBounces+1) per pixel, pixel-major (tile == frame)
    float* outColor = nullptr;       // RGBA32F
    float* outAlbedo = nullptr;
    float* outNormal = nullptr;
    uint64_t statEvals = 0;          // IMaterialInstance.eval calls x active wavelengths

    // ---- derived switches: pltpt.slangh:100-122 + PLTPT::getDefines (PLTPT.cpp:216-256)
    bool kUseEmissiveLights() const { return opt.useEmissiveLights && scene->d.emissiveTriangleCount > 0; }
    bool kUseEnvLight() const { return opt.useEnvLights && scene->env.valid(); }  // PLTPT.cpp:235
    bool kUseAnalyticLights() const { return opt.useAnalyticLights && scene->d.lightCount > 0; }
    bool kDoNEE() const { return opt.doNEE; }
    bool kDoMSonNEE() const { return opt.doNEE && opt.doMNEE; }
    bool kDoMIS() const { return opt.doNEE && opt.doMIS; }
    bool kNEEUsePerTileSG() const { return opt.doNEE && opt.NEEUsePerTileSG; }
    uint32_t maxBounces() const { return opt.maxBounces; }
    float3 camPos() const { return float3(cam.posW[0], cam.posW[1], cam.posW[2]); }
    float3 camW() const { return float3(cam.cameraW[0], cam.cameraW[1], cam.cameraW[2]); }
    float3 camU() const { return float3(cam.cameraU[0], cam.cameraU[1], cam.cameraU[2]); }
    float3 camV() const { return float3(cam.cameraV[0], cam.cameraV[1], cam.cameraV[2]); }

    PltptBounce& bounceAt(uint32_t px, uint32_t py, int p) const { return bounces[(size_t(py) * width + px) * (maxBounces() + 1) + p]; }

    // pltpt.slangh:138-165
    void appendBounce(const Path& path, const uint32_t hit[4], float3 wi, uint32_t lobe) const {
        PltptBounce& b = bounceAt(path.id & 0xffff, path.id >> 16, path.length);
        b.flags = path.flags.flags;
        for (int i = 0; i < 4; ++i) b.hit[i] = hit[i];
        const float2 e = ndir_to_oct_equal_area_snorm(wi);
        b.encodedWi[0] = e.x; b.encodedWi[1] = e.y;
        b.sampledBsdfLobe = lobe;
        b.lastNonDeltaPdf = path.lastNonDeltaPdf;
        b.pathContribution = 1.f / (path.pdf * path.thp_modifier);
    }
    static float3 bounceWi(const PltptBounce& b) { return oct_to_ndir_equal_area_snorm(float2(b.encodedWi[0], b.encodedWi[1])); }

    // ======================================================================== sample stage
    // Camera.slang:46-78
    void computeRayPinhole(uint32_t px, uint32_t py, float3& origin, float3& dir) const {
        float2 p = float2((float(px) + 0.5f) / float(width), (float(py) + 0.5f) / float(height));
        p += float2(-cam.jitterX, cam.jitterY);
        const float2 ndc = float2(2, -2) * p + float2(-1, 1);
        origin = camPos();
        dir = normalize(ndc.x * camU() + ndc.y * camV() + camW());
    }
    // pltpt.slangh:170-176, pltpt_sample.rt.slang:36-41
    static float RR(float rrVal, float rrE, float sample) {
        const float p = std::pow(rrVal, rrE);
        return sample < p ? .0f : 1.f - p;
    }
    // pltpt_sample.rt.slang:63-141
    void scatterClosestHit(Path& path, const TriangleHit& triangleHit) const {
        const uint32_t px = path.id & 0xffff, py = path.id >> 16;
        uint32_t hitInfo[4];
        scene->packHit(triangleHit, hitInfo);
        const VertexData v = scene->getVertexData(triangleHit);
        const uint32_t materialID = scene->getMaterialID(triangleHit.instanceID);
        const float3 wi = -path.dir;
        const ShadingData sd = scene->prepareShadingData(v, materialID, wi);
        const MaterialInstance bsdf = scene->getMaterialInstance(sd);
        const BSDFProperties bsdfProp = bsdf.getProperties();
        if (opt.alphaMasking && sd.opacity == .0f) {
            path.origin = sd.computeNewRayOrigin(false);
            return;
        }
        const float rcp_thp = opt.doRussianRoulette ? RR(1.f - path.weight, 2.f, sampleNext1D(path.sg)) : 1.f;
        bool shouldEndPath = path.length == int(maxBounces()) || (rcp_thp == .0f && path.length > 0);
        BSDFSample bsdfSample;
        if (!shouldEndPath) {
            const bool valid = bsdf.sample(sd, path.sg, bsdfSample, path.wavelength);
            shouldEndPath = !(valid && bsdfSample.pdf != .0f);
            const float new_contribution = path.pdf * bsdfSample.pdf * path.thp_modifier * rcp_thp;
            shouldEndPath |= new_contribution < 1e-6f;
        }
        const bool isDelta = bsdfSample.isLobe(L_Delta), isDispersive = bsdfSample.isLobe(L_DispersiveDelta), isReflection = bsdfSample.isLobe(L_Reflection);
        const bool isMSCompatible = (isReflection && bsdfProp.canManifoldWalkReflective()) || (!isReflection && bsdfProp.canManifoldWalkTransmissive());
        if (!isDelta && path.flags.onlyDelta()) {
            const size_t o = (size_t(py) * width + px) * 4;
            if (outAlbedo) { outAlbedo[o] = bsdfProp.albedo.x; outAlbedo[o + 1] = bsdfProp.albedo.y; outAlbedo[o + 2] = bsdfProp.albedo.z; outAlbedo[o + 3] = 1.f; }
            if (outNormal) {
                for (int r = 0; r < 3; ++r) outNormal[o + r] = cam.viewMat[4 * r] * sd.N.x + cam.viewMat[4 * r + 1] * sd.N.y + cam.viewMat[4 * r + 2] * sd.N.z;
                outNormal[o + 3] = 1.f;
            }
        }
        if (shouldEndPath) {
            path.flags.set(RayFlags::terminatedF, true);
            appendBounce(path, hitInfo, wi, 0);
            return;
        }
        path.weight *= bsdfSample.weight;
        path.thp_modifier *= rcp_thp;
        path.pdf *= bsdfSample.pdf;
        path.flags.set(RayFlags::hasDispersiveDeltaSegmentF, path.flags.dispersiveDeltaSegment() || isDispersive);
        appendBounce(path, hitInfo, wi, bsdfSample.lobe);
        ++path.length;
        path.lastNonDeltaPdf = !isDelta ? bsdfSample.pdf : path.lastNonDeltaPdf;
        path.flags.setMSCompatibleBounces(isMSCompatible ? path.flags.MSCompatibleBounces() + 1 : 0u);
        path.flags.set(RayFlags::lastBounceWasDeltaF, isDelta);
        path.flags.set(RayFlags::onlyDeltaF, path.flags.onlyDelta() && isDelta);
        path.normal = sd.N;
        path.origin = sd.computeNewRayOrigin(isReflection);
        path.dir = bsdfSample.wo;
    }
    // pltpt_sample.rt.slang:146-185
    void samplePixel(uint32_t px, uint32_t py) const {
        SampleGenerator sg(px, py, frameCount);
        const float samplingWavelength = SpectrumSampler::sampleOne(sampleNext1D(sg));
        const size_t o = (size_t(py) * width + px) * 4;
        if (outNormal) for (int i = 0; i < 4; ++i) outNormal[o + i] = 0.f;
        if (outAlbedo) for (int i = 0; i < 4; ++i) outAlbedo[o + i] = 0.f;
        Path path;
        path.id = (py << 16) | px;
        computeRayPinhole(px, py, path.origin, path.dir);
        path.normal = normalize(camW());
        path.sg = sg;
        path.wavelength = samplingWavelength;
        path.flags.set(RayFlags::onlyDeltaF, true);
        // Safety bound on alpha-mask pass-throughs (they do not advance `length`).
        int guard = 0;
        for (path.length = 0; path.length < int(maxBounces()) + 1 && path.flags.valid() && guard < 4096; ++guard) {
            const RayHit h = scene->traceClosest(path.origin, path.dir, 0.f, FLT_MAXf);
            if (!h.hit.valid()) {
                const uint32_t zero[4] = {0, 0, 0, 0};
                path.flags.set(RayFlags::missF, true);
                appendBounce(path, zero, -path.dir, 0);
            } else scatterClosestHit(path, h.hit);
        }
    }

    // ======================================================================== solve stage
    struct PLTSolve {
        SampleGenerator sg, tilesg;
        float3 Lrgb, pathOrigin, pathNormal;
        float wavelength = 0;
        float4 debugOutput = float4(0.f);  // Path.slang:204-220 (DO_DEBUG_VIEW)
    };
    // pltpt_debug.slangh:19-29
    void writeDebugOutput(uint32_t type, PLTSolve& solveData, float3 val) const {
        if (opt.debugView == type && all_finite(val)) solveData.debugOutput += float4(val.x, val.y, val.z, 1.f);
    }
    // pltpt_debug.slangh:31-51
    void markDebug(PLTSolve& solveData, int bounceIdx, const PLTBeam& beam) const {
        if (std::isnan(solveData.Lrgb.x) || std::isnan(solveData.Lrgb.y) || std::isnan(solveData.Lrgb.z)) writeDebugOutput(DV_NaNs, solveData, float3(5, 0, 3));
        if (std::isinf(solveData.Lrgb.x) || std::isinf(solveData.Lrgb.y) || std::isinf(solveData.Lrgb.z)) writeDebugOutput(DV_Inf, solveData, float3(2, 0, 5));
        writeDebugOutput(DV_PathLength, solveData, ccMagma(bounceIdx / float(int(maxBounces()) - 1)));
        writeDebugOutput(DV_HwssComps, solveData, ccMagma(beam.hwss_comps / float(hwss_samples())));
        const float4 S(beam.L[0], beam.L1[0], beam.L2[0], beam.L3[0]);
        const float cp = std::sqrt(sqr(S.w)) / S.x;
        const float lp = std::sqrt(sqr(S.y) + sqr(S.z)) / S.x;
        const float lpd = .5f * std::atan2(S.z, S.y);
        writeDebugOutput(DV_Polarization, solveData, lp * float3(sqr(std::cos(lpd)), 0, sqr(std::sin(lpd))) + cp * float3(0, 1, 0));
    }
    struct LightSample {  // LightsSamplingUtils.slang
        float pdf = 0;
        float3 origin;
        float distance = 0;
        float3 dir;
        PLTBeam beam;
        uint32_t lightType = 0;
    };
    enum { LT_EnvMap = 0, LT_Emissive = 1, LT_Analytic = 2 };
    static constexpr float kShadowRayTMin = 1e-4f;

    // ShadowRays.slangh:50-52
    static float3 computeNewRayOrigin(const ShadingData& sd, float3 rayDir) { return computeRayOrigin(sd.posW, dot(sd.faceN, rayDir) >= .0f ? sd.faceN : -sd.faceN); }
    bool visibilityRay(float3 origin, float3 dir, float distMax, float distMin = kShadowRayTMin) const { return !scene->traceAny(origin, dir, distMin, distMax); }
    RayHit shadowRay(float3 origin, float3 dir, float distMax, float distMin = kShadowRayTMin) const { return scene->traceClosest(origin, dir, distMin, distMax); }

    // LightSampling.slangh:40-76
    void lightTypeProbabilities(float p[3]) const {
        p[0] = kUseEnvLight() ? 1.f : 0.f;
        p[1] = kUseEmissiveLights() ? 1.f : 0.f;
        p[2] = kUseAnalyticLights() ? 1.f : 0.f;
        float sum = p[0] + p[1] + p[2];
        if (sum == 0.f) return;
        float invSum = 1.f / sum;
        p[0] *= invSum; p[1] *= invSum; p[2] *= invSum;
    }
    bool selectLightType(uint32_t& lightType, float& pdf, SampleGenerator& sg) const {
        float p[3];
        lightTypeProbabilities(p);
        float u = sampleNext1D(sg);
        for (lightType = 0; lightType < 3; ++lightType) {
            if (u < p[lightType]) { pdf = p[lightType]; return true; }
            u -= p[lightType];
        }
        return false;
    }
    // LightCollectionShared.slang:118-135
    struct EmissiveTriangle { float3 posW[3]; float2 texCoords[3]; float3 normal; float area; uint32_t materialID; };
    EmissiveTriangle getEmissiveTriangle(uint32_t idx) const {
        const PltptEmissiveTriangle& p = scene->d.emissiveTriangles[idx];
        EmissiveTriangle t;
        for (int k = 0; k < 3; ++k) {
            t.posW[k] = float3(p.posAndTexCoords[k][0], p.posAndTexCoords[k][1], p.posAndTexCoords[k][2]);
            const uint32_t uv = asuint(p.posAndTexCoords[k][3]);
            t.texCoords[k] = float2(f16tof32(uv & 0xffff), f16tof32(uv >> 16));
        }
        t.normal = decodeNormal2x16(p.normal);
        t.area = asfloat(p.area);
        t.materialID = p.materialID;
        return t;
    }
    // EmissiveLightSamplerHelpers.slang:56-74 — Le / sampledWavelengths are outputs, the bool result is unused
    void materialEvalEmissive(uint32_t materialID, float2 uv, float wavelength, bool importanceSample, SampleGenerator& sg, hwss_t& Le, hwss_t& sampledWavelengths) const {
        const PltptMaterialBlob& md = scene->d.materialData[materialID];
        const bool isBasic = (md.header[0] >> 31) & 1u;
        if (isBasic && md.emissionSpectralId != PLTPT_NO_SPECTRAL_PROFILE) {
            const float3 emissionRGB = scene->textures.sampleTexture(md.texEmissive, TextureSystem::samplerIDOf(md), uv, float4(1.f)).xyz();  // :68
            Le = scene->profiles.sampleSpectrum(md.emissionSpectralId, wavelength, sg, importanceSample, sampledWavelengths);
            Le *= SpectrumUtils::rgbToSpectrum(emissionRGB, sampledWavelengths);
        }
    }
    struct TriangleLightSample { uint32_t triangleIndex = 0; float3 posW, normalW, dir; float distance = 0; hwss_t Le, wavelengths; float pdf = 0; };
    // EmissiveLightSamplerHelpers.slang:86-128
    bool sampleTriangle(float3 posW, uint32_t triangleIndex, SampleGenerator& sg, float wavelength, bool importanceSample, TriangleLightSample& ls) const {
        ls = TriangleLightSample();
        ls.triangleIndex = triangleIndex;
        const EmissiveTriangle tri = getEmissiveTriangle(triangleIndex);
        const float2 u = sampleNext2D(sg);
        const float3 b = sample_triangle(u);
        ls.posW = tri.posW[0] * b[0] + tri.posW[1] * b[1] + tri.posW[2] * b[2];
        float3 toLight = ls.posW - posW;
        const float distSqr = hlsl_max(FLT_MINf, dot(toLight, toLight));
        ls.distance = std::sqrt(distSqr);
        ls.dir = toLight / ls.distance;
        ls.normalW = tri.normal;
        float cosTheta = dot(ls.normalW, -ls.dir);
        if (cosTheta <= 0.f) return false;
        ls.Le = hwss_t(.0f);
        const float2 uv = tri.texCoords[0] * b[0] + tri.texCoords[1] * b[1] + tri.texCoords[2] * b[2];  // EmissiveTriangle::getTexCoord, LightCollectionShared.slang
        materialEvalEmissive(tri.materialID, uv, wavelength, importanceSample, sg, ls.Le, ls.wavelengths);
        float denom = hlsl_max(FLT_MINf, cosTheta * tri.area);
        ls.pdf = distSqr / denom;
        return true;
    }
    // EmissivePowerSampler.slang:64-90 / EmissiveUniformSampler.slang:56-72
    bool emissiveSampleLight(float3 posW, SampleGenerator& sg, float wavelength, bool importanceSample, TriangleLightSample& ls) const {
        const PltptSceneDesc& d = scene->d;
        if (opt.emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER) {
            if (d.emissiveTriangleCount == 0) return false;
            float uLight = sampleNext1D(sg);
            uint32_t triangleCount = d.emissiveTriangleCount;
            uint32_t triangleIndex = hlsl_min((uint32_t)(uLight * triangleCount), triangleCount - 1);
            const uint32_t px = d.triangleAliasTable[2 * triangleIndex], py = d.triangleAliasTable[2 * triangleIndex + 1];
            float threshold = f16tof32(px >> 16u);
            uint32_t selectAbove = ((px & 0xFFFFu) << 8u) | ((py >> 24u) & 0xFFu);
            uint32_t selectBelow = py & 0xFFFFFFu;
            triangleIndex = (sampleNext1D(sg) >= threshold) ? selectAbove : selectBelow;
            float triangleSelectionPdf = d.emissiveFlux[triangleIndex].flux * d.invWeightsSum;
            if (!sampleTriangle(posW, triangleIndex, sg, wavelength, importanceSample, ls)) return false;
            ls.pdf *= triangleSelectionPdf;
            return true;
        } else {
            if (d.activeTriangleCount == 0) return false;
            float uLight = sampleNext1D(sg);
            uint32_t triangleCount = d.activeTriangleCount;
            uint32_t idx = hlsl_min((uint32_t)(uLight * triangleCount), triangleCount - 1);
            uint32_t triangleIndex = d.activeTriangles[idx];
            float triangleSelectionPdf = 1.f / (float)triangleCount;
            if (!sampleTriangle(posW, triangleIndex, sg, wavelength, importanceSample, ls)) return false;
            ls.pdf *= triangleSelectionPdf;
            return true;
        }
    }
    // EmissivePowerSampler.slang:98-126, EmissiveLightSamplerHelpers.slang:135-151
    float emissiveEvalPdf(float3 posW, uint32_t triangleIndex, float3 hitPosW, float3 hitNormalW) const {
        const PltptSceneDesc& d = scene->d;
        if (triangleIndex == 0xffffffffu) return 0.f;
        float sel;
        if (opt.emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER) sel = d.emissiveFlux[triangleIndex].flux * d.invWeightsSum;
        else sel = d.activeTriangleCount > 0 ? 1.f / (float)d.activeTriangleCount : 0.f;
        float3 toLight = hitPosW - posW;
        const float distSqr = dot(toLight, toLight);
        if (distSqr <= FLT_MINf) return 0.f;
        float3 L = toLight / std::sqrt(distSqr);
        float cosTheta = dot(hitNormalW, -L);
        if (cosTheta <= 0.f) return 0.f;
        float denom = hlsl_max(FLT_MINf, cosTheta * getEmissiveTriangle(triangleIndex).area);
        return sel * (distSqr / denom);
    }
    // LightSampling.slangh:86-102
    bool generateEnvMapSample(const ShadingData& sd, SampleGenerator& sg, float wavelength, LightSample& ls) const {
        ls = LightSample();
        if (!kUseEnvLight()) return false;
        EnvMap::Sample lightSample;
        if (!scene->env.sample(sampleNext2D(sg), lightSample)) return false;
        const hwss_t wavelengths = SpectrumSampler::hwssWavelengths(wavelength);
        const hwss_t Le = SpectrumUtils::rgbToSpectrum(lightSample.Le, wavelengths);
        const hwss_t L = lightSample.pdf > 0.f ? Le / lightSample.pdf : hwss_t(.0f);
        ls.beam = pltSourceBeamDistant(-lightSample.dir, opt.sourcingMaxBeamOmega, L, wavelengths, opt.sourcingMaxBeamOmega, false, hwss_samples());
        ls.pdf = lightSample.pdf;
        ls.origin = computeNewRayOrigin(sd, lightSample.dir);
        ls.distance = FLT_MAXf;  // kRayTMax
        ls.dir = lightSample.dir;
        return anyPositive(L);
    }
    // LightSampling.slangh:104-126
    bool generateEmissiveSample(const ShadingData& sd, SampleGenerator& sg, float wavelength, bool importanceSample, LightSample& ls) const {
        ls = LightSample();
        if (!kUseEmissiveLights()) return false;
        TriangleLightSample tls;
        if (!emissiveSampleLight(sd.posW, sg, wavelength, importanceSample, tls)) return false;
        const hwss_t L = tls.pdf > 0.f ? tls.Le / tls.pdf : hwss_t(.0f);
        ls.pdf = tls.pdf;
        const float3 lightPos = computeRayOrigin(tls.posW, tls.normalW);
        ls.origin = computeNewRayOrigin(sd, lightPos - sd.posW);
        const float3 toLight = lightPos - ls.origin;
        ls.distance = length(toLight);
        ls.dir = normalize(toLight);
        ls.beam = pltSourceBeamArea(tls.posW, -ls.dir, opt.sourcingAreaFromEmissiveGeometry, ls.distance, L, tls.wavelengths, opt.sourcingMaxBeamOmega, false, hwss_samples());
        return anyPositive(L);
    }
    static bool anyPositive(hwss_t L) {
        for (int j = 0; j < hwss_samples(); ++j) if (L[j] > 0.f) return true;
        return false;
    }
    // LightHelpers.slang:70-330 (analytic lights)
    struct AnalyticLightSample { float3 posW, normalW, dir; float distance = 0; hwss_t Li, wavelengths; float pdf = 0; };
    bool finalizeAreaLightSample(float3 shadingPosW, const PltptLightData& light, SampleGenerator& sg, float wavelength, bool importanceSample, AnalyticLightSample& ls) const {
        float3 toLight = ls.posW - shadingPosW;
        float distSqr = hlsl_max(dot(toLight, toLight), 1e-9f);
        ls.distance = std::sqrt(distSqr);
        ls.dir = toLight / ls.distance;
        float cosTheta = dot(ls.normalW, -ls.dir);
        if (cosTheta <= 0.f) return false;
        ls.Li = scene->profiles.sampleSpectrum(light.intensitySpectrumId, wavelength, sg, importanceSample, ls.wavelengths);
        ls.Li *= light.surfaceArea * cosTheta / distSqr;
        ls.pdf = distSqr / (cosTheta * light.surfaceArea);
        return true;
    }
    bool sampleAnalyticLight(float3 shadingPosW, const PltptLightData& light, SampleGenerator& sg, float wavelength, bool importanceSample, AnalyticLightSample& ls) const {
        ls = AnalyticLightSample();
        const float3 lightPos(light.posW[0], light.posW[1], light.posW[2]), lightDir(light.dirW[0], light.dirW[1], light.dirW[2]);
        switch (light.type) {
            case 0: {  // Point
                ls.posW = lightPos;
                ls.normalW = lightDir;
                float3 toLight = ls.posW - shadingPosW;
                float distSqr = hlsl_max(dot(toLight, toLight), 1e-9f);
                ls.distance = std::sqrt(distSqr);
                ls.dir = toLight / ls.distance;
                float cosTheta = -dot(ls.dir, lightDir);
                float falloff = 1.f;
                if (cosTheta < light.cosOpeningAngle) falloff = 0.f;
                else if (light.penumbraAngle > 0.f) {
                    float deltaAngle = light.openingAngle - std::acos(cosTheta);
                    float tt = clampf(deltaAngle / light.penumbraAngle, 0.f, 1.f);
                    falloff = tt * tt * (3.f - 2.f * tt);
                }
                ls.Li = scene->profiles.sampleSpectrum(light.intensitySpectrumId, wavelength, sg, importanceSample, ls.wavelengths);
                ls.Li *= falloff / distSqr;
                ls.pdf = 0.f;
                return true;
            }
            case 1:  // Directional
                ls.posW = float3(std::numeric_limits<float>::infinity());
                ls.normalW = lightDir;
                ls.distance = FLT_MAXf;
                ls.dir = -lightDir;
                ls.Li = scene->profiles.sampleSpectrum(light.intensitySpectrumId, wavelength, sg, importanceSample, ls.wavelengths);
                ls.pdf = 0.f;
                return true;
            case 2: {  // Distant
                ls.posW = float3(std::numeric_limits<float>::infinity());
                const float2 u = sampleNext2D(sg);
                float3 dir = sample_cone(u, light.cosSubtendedAngle);
                ls.dir = normalize(Scene::mulVec(light.transMat, dir));
                ls.normalW = -ls.dir;
                ls.distance = FLT_MAXf;
                float solidAngle = k2Pi * (1.f - light.cosSubtendedAngle);
                ls.Li = scene->profiles.sampleSpectrum(light.intensitySpectrumId, wavelength, sg, importanceSample, ls.wavelengths);
                ls.pdf = 1.f / solidAngle;
                return true;
            }
            case 3: case 4: case 5: {  // Rect, Disc, Sphere
                const float2 u = sampleNext2D(sg);
                float3 pos, nrm(0.f, 0.f, 1.f);
                if (light.type == 3) pos = float3(u.x * 2.f - 1.f, u.y * 2.f - 1.f, 0.f);
                else if (light.type == 4) pos = float3(sample_disk(u), 0.f);
                else { pos = sample_sphere(u); nrm = pos; }
                ls.posW = Scene::mulPoint(light.transMat, pos);
                ls.normalW = normalize(Scene::mulVec(light.transMatIT, nrm));
                return finalizeAreaLightSample(shadingPosW, light, sg, wavelength, importanceSample, ls);
            }
            default: return false;
        }
    }
    // LightSampling.slangh:129-166
    bool generateAnalyticLightSample(const ShadingData& sd, SampleGenerator& sg, float wavelength, bool importanceSample, LightSample& ls) const {
        ls = LightSample();
        uint32_t lightCount = scene->d.lightCount;
        if (!kUseAnalyticLights() || lightCount == 0) return false;
        uint32_t lightIndex = hlsl_min(uint32_t(sampleNext1D(sg) * lightCount), lightCount - 1);
        const PltptLightData& light = scene->d.lights[lightIndex];
        AnalyticLightSample lightSample;
        if (!sampleAnalyticLight(sd.posW, light, sg, wavelength, importanceSample, lightSample)) return false;
        const bool isDistantEmitter = light.type == 1 || light.type == 2;
        const float A = light.surfaceArea;
        const float Omega = 4.f * kPi * (1.f - sqr(light.cosSubtendedAngle));
        const hwss_t L = lightSample.Li * float(lightCount);
        ls.pdf = lightSample.pdf / lightCount;
        ls.origin = computeNewRayOrigin(sd, lightSample.dir);
        ls.distance = lightSample.distance;
        ls.dir = lightSample.dir;
        const hwss_t wavelengths = lightSample.wavelengths;
        if (isDistantEmitter) ls.beam = pltSourceBeamDistant(-ls.dir, Omega, L, wavelengths, opt.sourcingMaxBeamOmega, false, hwss_samples());
        else ls.beam = pltSourceBeamArea(float3(light.posW[0], light.posW[1], light.posW[2]), -ls.dir, A, ls.distance, L, wavelengths, opt.sourcingMaxBeamOmega, false, hwss_samples());
        return anyPositive(L);
    }
    // LightSampling.slangh:168-206
    bool generateLightSample(const ShadingData& sd, bool sampleReflection, bool sampleTransmission, SampleGenerator& sg, SampleGenerator& tilesg, float wavelength, bool importanceSample, LightSample& ls) const {
        ls = LightSample();
        uint32_t lightType;
        float selectionPdf;
        const bool success = kNEEUsePerTileSG() ? selectLightType(lightType, selectionPdf, tilesg) : selectLightType(lightType, selectionPdf, sg);
        if (!success) return false;
        bool valid = false;
        if (kUseEnvLight() && lightType == LT_EnvMap) valid = generateEnvMapSample(sd, sg, wavelength, ls);
        if (kUseEmissiveLights() && lightType == LT_Emissive) valid = generateEmissiveSample(sd, sg, wavelength, importanceSample, ls);
        if (kUseAnalyticLights() && lightType == LT_Analytic) valid = generateAnalyticLightSample(sd, sg, wavelength, importanceSample, ls);
        ls.lightType = lightType;
        ls.pdf *= selectionPdf;
        ls.beam.scale(1.f / selectionPdf);
        const bool isReflection = dot(sd.faceN, ls.dir) * dot(sd.faceN, sd.V) > .0f;
        return valid && ls.pdf > 1e-6f && !(!sampleReflection && isReflection) && !(!sampleTransmission && !isReflection);
    }

    // pltpt_solve.rt.slang:37-93
    void forwardUnwindStep(const TriangleHit& triangleHit, const VertexData& v, float3 wo, uint32_t lobe, PLTBeam& beam) {
        const uint32_t materialID = scene->getMaterialID(triangleHit.instanceID);
        const ShadingData sd = scene->prepareShadingData(v, materialID, -beam.r);
        beam.propagate(sd.posW);
        const MaterialInstance bsdf = scene->getMaterialInstance(sd);
        bsdf.eval(sd, wo, lobe, beam);
        statEvals += beam.hwss_comps;
        if (opt.HWSSDoMIS) beam.spectralPDFs *= bsdf.lambdaPdf(sd, wo, lobe, beam, beam.wavelengths);
    }
    void forwardUnwind(int bounceIdx, uint32_t px, uint32_t py, PLTBeam& beam) {
        for (int p = bounceIdx - 1; p >= 0; --p) {
            const PltptBounce& bounce = bounceAt(px, py, p);
            const TriangleHit th = scene->unpackHit(bounce.hit);
            const VertexData v = scene->getVertexData(th);
            forwardUnwindStep(th, v, bounceWi(bounce), bounce.sampledBsdfLobe, beam);
        }
    }
    // pltpt_solve.rt.slang:96-147
    void measure(PLTSolve& solveData, int bounceIdx, float pathContribution, PLTBeam& beam) const {
        if (!opt.useDirectLights && bounceIdx <= 1) return;
        if (opt.debugView) {  // pltpt_solve.rt.slang:100-106: coherence views before the propagation to the camera
            const float k = k2Pi / (1e-3f * beam.wavelengths.x);
            writeDebugOutput(DV_CoherenceArea, solveData, ccMagma(std::sqrt(beam.coherenceArea(k)) / 150.f));
            float2x2 U, S;
            svd_symmetric(beam.Theta.toInvCoherenceMat(k), U, S);
            writeDebugOutput(DV_CoherenceAnisotropy, solveData, ccInferno(1.f - S.m[0][0] / S.m[1][1]));
        }
        beam.propagate(camPos());
        if (cam.polarizer > .0f) {
            beam.rotateFrame(normalize(camU()), -normalize(camW()));
            const float4x4 P = (1.f - cam.polarizer) * ident4() + cam.polarizer * MuellerPolarizer(cam.polarizerTheta);
            for (uint32_t j = 0; j < beam.hwss_comps; ++j) beam.setSP(j, mul(P, beam.getSP(j)));
        }
        if (opt.HWSSDoMIS) {
            float sum = 0;
            for (int j = 0; j < hwss_samples(); ++j) sum += beam.spectralPDFs[j];
            pathContribution *= sum > FLT_EPS ? beam.spectralPDFs[0] / sum : .0f;
        } else pathContribution *= 1.f / float(beam.hwss_comps);
        const float3 Lrgb = vmax(float3(.0f), SpectrumUtils::spectrumToRgb(beam.L, beam.wavelengths, beam.hwss_comps) * pathContribution);
        if (all_finite(Lrgb)) solveData.Lrgb += Lrgb;
        if (opt.debugView) markDebug(solveData, bounceIdx, beam);
    }
    // pltpt_solve.rt.slang:205-217
    bool wasNEEAvailable(int bounceIdx, RayFlags f) const {
        return kDoNEE() && bounceIdx > 0 && (!f.lastBounceWasDelta() || (kDoMSonNEE() && f.MSCompatibleBounces() <= opt.MNEEMaxOccluders && bounceIdx > 1));
    }
    static float MISWeight(float p0, float p1) { return p0 / (p0 + p1); }

    // pltpt_solve.rt.slang:152-200
    void evalEmissive(PLTSolve& solveData, int bounceIdx, const PltptBounce& bounce, uint32_t px, uint32_t py, const BSDFProperties& bsdfProp, const ShadingData& sd, const TriangleHit& triangleHit, bool importanceSample) {
        RayFlags bf; bf.flags = bounce.flags;
        const bool didNEE = wasNEEAvailable(bounceIdx, bf);
        if (kUseEmissiveLights() && bsdfProp.isEmissive() && (kDoMIS() || !didNEE)) {
            const bool needMIS = kDoMIS() && didNEE;
            const uint32_t triIdx = scene->getEmissiveTriangleIndex(triangleHit.instanceID, triangleHit.primitiveIndex);
            const float3 hitN = sd.frontFacing ? sd.faceN : -sd.faceN;
            float p[3];
            lightTypeProbabilities(p);
            const float lightPdf = p[1] * emissiveEvalPdf(solveData.pathOrigin, triIdx, sd.posW, hitN);
            const float misWeight = needMIS ? MISWeight(bounce.lastNonDeltaPdf, lightPdf) : 1.f;
            hwss_t sampledWavelengths(.0f);
            hwss_t Le = scene->profiles.sampleSpectrum(bsdfProp.emissionSpectralId, solveData.wavelength, solveData.sg, importanceSample, sampledWavelengths);
            Le *= misWeight * SpectrumUtils::rgbToSpectrum(bsdfProp.emissionRGB, sampledWavelengths);
            if (anyPositive(Le)) {
                const float dist = length(solveData.pathOrigin - sd.posW);
                PLTBeam beam = pltSourceBeamArea(sd.posW, bounceWi(bounce), opt.sourcingAreaFromEmissiveGeometry, dist, Le, sampledWavelengths, opt.sourcingMaxBeamOmega, false, hwss_samples());
                if (!opt.HWSSDoMIS && bf.dispersiveDeltaSegment()) beam.collapseToHeroWavelengthOnly();
                forwardUnwind(bounceIdx, px, py, beam);
                measure(solveData, bounceIdx, bounce.pathContribution, beam);
            }
        }
    }

    // pltpt_solve.rt.slang:371-396
    void onMiss(PLTSolve& solveData, int bounceIdx, const PltptBounce& bounce, uint32_t px, uint32_t py) {
        const float3 wo = bounceWi(bounce);
        RayFlags bf; bf.flags = bounce.flags;
        const bool didNEE = wasNEEAvailable(bounceIdx, bf);
        if (kUseEnvLight() && (kDoMIS() || !didNEE)) {
            const bool needMIS = kDoMIS() && didNEE;
            float p[3];
            lightTypeProbabilities(p);
            const float lightPdf = p[0] * scene->env.evalPdf(-wo);
            const float misWeight = needMIS ? MISWeight(bounce.lastNonDeltaPdf, lightPdf) : 1.f;
            const hwss_t wavelengths = SpectrumSampler::hwssWavelengths(solveData.wavelength);
            const hwss_t Le = misWeight * SpectrumUtils::rgbToSpectrum(scene->env.eval(-wo), wavelengths);
            PLTBeam beam = pltSourceBeamDistant(wo, opt.sourcingMaxBeamOmega, Le, wavelengths, opt.sourcingMaxBeamOmega, false, hwss_samples());
            if (!opt.HWSSDoMIS && bf.dispersiveDeltaSegment()) beam.collapseToHeroWavelengthOnly();
            forwardUnwind(bounceIdx, px, py, beam);
            measure(solveData, bounceIdx, bounce.pathContribution, beam);
        }
    }

    // ---------------------------------------------------------------- MNEE (Manifolds.slangh)
    struct MSHitData { uint32_t hit[4] = {0, 0, 0, 0}; float3 posW; uint32_t lobe = 0; };
    struct ManifoldSampleData {
        MSHitData hits[8];
        float3 woLastHitToLight;
        int hitCount = 0;
        uint32_t iterations = 0;
        float wavelength = 0, pdf = 1.f;
        bool dispersive = false;
    };
    bool manifoldSampleVertex(const uint32_t hitInfo[4], const ShadingData& sd, const MaterialInstance& bsdf, float3 LightDir, SampleGenerator& sg, ManifoldSampleData& data) const {
        BSDFSample sample;
        if (!bsdf.sample(sd, sg, sample, data.wavelength, L_Delta, LightDir)) return false;
        data.pdf *= sample.pdf;
        data.hits[data.hitCount].lobe = sample.lobe;
        for (int i = 0; i < 4; ++i) data.hits[data.hitCount].hit[i] = hitInfo[i];
        data.hits[data.hitCount].posW = sd.posW;
        data.dispersive |= (sample.lobe & L_DispersiveDelta) != 0;
        ++data.hitCount;
        return true;
    }
    // Manifolds.slangh:110-160
    bool traceMNEEOccluders(float3 rayOrigin, float3 rayDir, float rayTMin, float rayTMax, SampleGenerator& sg, int maxMSBounces, ManifoldSampleData& data) const {
        const float3 wi = -rayDir;
        for (;;) {
            // shadowRay(RayDesc) keeps the desc's own TMin (0 for the NEE visibility ray, LightsSamplingUtils.slang getVisibilityRay)
            const RayHit hit = scene->traceClosest(rayOrigin, rayDir, rayTMin, rayTMax);
            if (!hit.hit.valid()) return true;
            if (data.hitCount >= int(opt.MNEEMaxOccluders) || data.hitCount >= maxMSBounces) return false;
            VertexData v = scene->getVertexData(hit.hit);
            v.posW = rayOrigin + hit.t * rayDir;
            const uint32_t materialID = scene->getMaterialID(hit.hit.instanceID);
            const ShadingData sd = scene->prepareShadingData(v, materialID, wi);
            const MaterialInstance bsdf = scene->getMaterialInstance(sd);
            const BSDFProperties bsdfProp = bsdf.getProperties();
            if (!bsdfProp.canManifoldWalkTransmissive()) return false;
            uint32_t hitInfo[4];
            scene->packHit(hit.hit, hitInfo);
            if (!manifoldSampleVertex(hitInfo, sd, bsdf, -wi, sg, data) || (data.hits[data.hitCount - 1].lobe & L_Transmission) == 0) return false;
            rayOrigin = computeNewRayOrigin(sd, rayDir);
            rayTMax -= hit.t;
        }
    }
    // DifferentialShadingGeometry.slang:94-179
    DifferentialShadingGeometry differentialShadingGeometry(const TriangleHit& hit) const {
        const float3 bary = hit.getBarycentricWeights();
        uint32_t idx[3];
        scene->getIndices(hit.instanceID, hit.primitiveIndex, idx);
        StaticVertexData vs[3] = {scene->getVertex(idx[0]), scene->getVertex(idx[1]), scene->getVertex(idx[2])};
        float2 uv0 = vs[0].texCrd, uv1 = vs[1].texCrd, uv2 = vs[2].texCrd;
        if (all_eq(uv0, 0.f) && all_eq(uv1, 0.f) && all_eq(uv2, 0.f)) { uv1 = float2(1.0f, 0.0f); uv2 = float2(0.0f, 1.0f); }
        const float3 p0 = vs[0].position, p1 = vs[1].position, p2 = vs[2].position;
        const float3 n0 = vs[0].normal, n1 = vs[1].normal, n2 = vs[2].normal;
        DifferentialShadingGeometry sdg;
        sdg.ngeo = normalize(cross(p1 - p0, p2 - p0));
        sdg.uv = bary[0] * uv0 + float2(bary[1]) + uv1 + bary[2] * uv2;  // sic (DifferentialShadingGeometry.slang:122)
        sdg.p.v = bary[0] * p0 + bary[1] * p1 + bary[2] * p2;
        sdg.n.v = bary[0] * n0 + bary[1] * n1 + bary[2] * n2;
        float il = 1.0f / length(sdg.n.v);
        sdg.n.v *= il;
        sdg.p.ddu = perp_stark(sdg.n.v);
        sdg.p.ddv = cross(sdg.n.v, sdg.p.ddu);
        float2 duv0 = uv1 - uv0, duv1 = uv2 - uv0;
        float3 dp0 = p1 - p0, dp1 = p2 - p0;
        sdg.n.ddu = (n1 - n0) * il;
        sdg.n.ddv = (n2 - n0) * il;
        sdg.n.ddu -= sdg.n.v * dot(sdg.n.v, sdg.n.ddu);
        sdg.n.ddv -= sdg.n.v * dot(sdg.n.v, sdg.n.ddv);
        float det = duv0.x * duv1.y - duv0.y * duv1.x;
        if (std::fabs(det) > FLT_EPS) {
            float invDet = 1.0f / det;
            sdg.p.ddu = (duv1.y * dp0 - duv0.y * dp1) * invDet;
            sdg.p.ddv = (-duv1.x * dp0 + duv0.x * dp1) * invDet;
            float3 nddu = (duv1.y * sdg.n.ddu - duv0.y * sdg.n.ddv) * invDet;
            float3 nddv = (-duv1.x * sdg.n.ddu + duv0.x * sdg.n.ddv) * invDet;
            sdg.n.ddu = nddu;
            sdg.n.ddv = nddv;
        } else { sdg.n.ddu = float3(0.0f); sdg.n.ddv = float3(0.0f); }
        sdg.s.v = sdg.p.ddu - sdg.n.v * dot(sdg.n.v, sdg.p.ddu);
        float ils = 1.0f / length(sdg.s.v);
        sdg.s.v *= ils;
        if (all_eq(sdg.p.ddu, 0.0f)) { sdg.s.v = perp_stark(sdg.n.v); ils = 1.0f; }
        sdg.t.v = cross(sdg.n.v, sdg.s.v);
        sdg.s.ddu = ils * (-sdg.n.ddu * dot(sdg.n.v, sdg.p.ddu) - sdg.n.v * dot(sdg.n.ddu, sdg.p.ddu));
        sdg.s.ddv = ils * (-sdg.n.ddv * dot(sdg.n.v, sdg.p.ddu) - sdg.n.v * dot(sdg.n.ddv, sdg.p.ddu));
        sdg.s.ddu -= sdg.s.v * dot(sdg.s.ddu, sdg.s.v);
        sdg.s.ddv -= sdg.s.v * dot(sdg.s.ddv, sdg.s.v);
        sdg.t.ddu = cross(sdg.n.ddu, sdg.s.v) + cross(sdg.n.v, sdg.s.ddu);
        sdg.t.ddv = cross(sdg.n.ddv, sdg.s.v) + cross(sdg.n.v, sdg.s.ddv);
        return sdg;
    }
    // Manifolds.slangh:284-347
    static float2 sphcoords(float3 w) {
        float theta = std::acos(clampf(w.z, -1.f, 1.f));
        float phi = std::atan2(w.y, w.x);
        phi += phi < .0f ? k2Pi : .0f;
        return float2(theta, phi);
    }
    static bool manifoldOneStep(const float3pdv& w, const float3kpdv& ws, const float3pdv& p, float& lenC, float3& positionStepDir) {
        const float2 spho = sphcoords(w.v), sphio = sphcoords(ws.v);
        const float dt = spho.x - sphio.x;
        float dp = spho.y - sphio.y;
        dp = std::fabs(dp) > kPi ? dp - signf(dp) * k2Pi : dp;
        const float2 C(dt, dp);
        lenC = length(C);
        auto deriv = [](float3 v, float3 ddu, float3 ddv) {
            const float dAcos = -1.f / std::sqrt(hlsl_max(.0f, 1.f - sqr(v.z)));
            const float invNorm = 1.f / dot(v.xy(), v.xy());
            return float2x2(dAcos * ddu.z, dAcos * ddv.z, invNorm * (v.x * ddu.y - v.y * ddu.x), invNorm * (v.x * ddv.y - v.y * ddv.x));
        };
        const float2x2 dO = deriv(w.v, w.ddu, w.ddv);
        const float2x2 dIO = deriv(ws.v, ws.ddu, ws.ddv);  // float2x2(float2x3) keeps the (u,v) columns
        const float2x2 dCdX = dO - dIO;
        float2x2 invdCdX;
        if (!tryInvert(dCdX, invdCdX)) return false;
        const float2 dX = mul(invdCdX, C);
        positionStepDir = p.ddu * dX.x + p.ddv * dX.y;
        return true;
    }
    // Manifolds.slangh:165-281
    bool manifoldNEESample(float3 visOrigin, float3 visDir, float visTMin, float visTMax, float3 origin, float3 lightPos, float3 lightDir, bool isDistantLight,
                           int maxMSBounces, SampleGenerator& sg, ManifoldSampleData& data) const {
        if (!traceMNEEOccluders(visOrigin, visDir, visTMin, visTMax, sg, maxMSBounces, data)) return false;
        if (data.hitCount == 0) return true;
        float beta[8];
        for (int i = 0; i < 8; ++i) beta[i] = 1.f;
        bool success = false;
        for (; !success;) {
            success = true;
            for (int idx = 0; idx < data.hitCount; ++idx) {
                const TriangleHit triangleHit = scene->unpackHit(data.hits[idx].hit);
                const VertexData v = scene->getVertexData(triangleHit);
                const uint32_t materialID = scene->getMaterialID(triangleHit.instanceID);
                const DifferentialShadingGeometry dsg = differentialShadingGeometry(triangleHit);
                const float3 p0 = idx == 0 ? origin : data.hits[idx - 1].posW;
                const float3pdv pdvp0 = dsg.pdw(p0);
                const float3 p2 = idx + 1 == data.hitCount ? lightPos : data.hits[idx + 1].posW;
                float3pdv pdvp2;
                pdvp2.v = lightDir;
                if (!isDistantLight) pdvp2 = dsg.pdw(p2);
                const bool woIsIn = dot(pdvp2.v, dsg.n.v) < .0f;
                const float3pdv wiP = woIsIn ? pdvp0 : pdvp2;
                const float3pdv woP = woIsIn ? pdvp2 : pdvp0;
                const ShadingData sd = scene->prepareShadingData(v, materialID, wiP.v);
                const MaterialInstance bsdf = scene->getMaterialInstance(sd);
                const float3kpdv wos = bsdf.partialDerivatives(wiP, dsg, data.hits[idx].lobe, data.wavelength);
                float lenC;
                float3 positionStepDir;
                if (wos.v.x == .0f && wos.v.y == .0f && wos.v.z == .0f) return false;
                if (!manifoldOneStep(woP, wos, dsg.p, lenC, positionStepDir)) return false;
                if (lenC < opt.MNEESolverThreshold) continue;
                success = false;
                for (;;) {
                    if (++data.iterations > opt.MNEEMaxIterations) return false;
                    const float3 proposal = data.hits[idx].posW - beta[idx] * positionStepDir;
                    float3 dirToProposal = proposal - p0;
                    const float dist = length(dirToProposal);
                    dirToProposal /= dist;
                    const RayHit closestHit = shadowRay(p0, dirToProposal, 1.01f * dist, hlsl_max(kShadowRayTMin, .99f * dist));
                    // (reference unpacks the hit before testing isHit; a miss unpacks to instance 0)
                    const uint32_t propInst = closestHit.hit.valid() ? closestHit.hit.instanceID : 0u;
                    if (!closestHit.hit.valid() || scene->unpackHit(data.hits[idx].hit).instanceID != propInst) {
                        beta[idx] /= 4.f;
                        continue;
                    }
                    scene->packHit(closestHit.hit, data.hits[idx].hit);
                    data.hits[idx].posW = p0 + dirToProposal * closestHit.t;
                    beta[idx] = hlsl_min(1.0f, 2.f * beta[idx]);
                    break;
                }
            }
        }
        if (success) {
            const float3 p = data.hits[data.hitCount - 1].posW;
            float3 dir;
            float dist;
            if (!isDistantLight) { dir = lightPos - p; dist = length(dir); dir /= dist; }
            else { dir = lightDir; dist = FLT_MAXf; }
            success = visibilityRay(p, dir, dist);
            data.woLastHitToLight = dir;
        }
        return success;
    }
    void manifoldForwardUnwind(const ManifoldSampleData& result, float3 origin, PLTBeam& beam) {
        for (int idx = result.hitCount - 1; idx >= 0; --idx) {
            const TriangleHit th = scene->unpackHit(result.hits[idx].hit);
            VertexData v = scene->getVertexData(th);
            v.posW = result.hits[idx].posW;
            const float3 p2 = idx == 0 ? origin : result.hits[idx - 1].posW;
            const float3 wo = normalize(p2 - v.posW);
            forwardUnwindStep(th, v, wo, result.hits[idx].lobe, beam);
        }
    }

    // pltpt_solve.rt.slang:219-366
    void evalNEE(PLTSolve& solveData, int bounceIdx, const PltptBounce& bounce, uint32_t px, uint32_t py, const VertexData& V, uint32_t materialID, const MaterialInstance& bsdf, PLTBeam& beam, bool NEEEvalCurrentVertex) {
        if (NEEEvalCurrentVertex) {
            const ShadingData forwardSd = scene->prepareShadingData(V, materialID, -beam.r);
            beam.propagate(forwardSd.posW);
            // NB: `bsdf` was instanced with the backward shading data; the grating frame depends only on sd.uv/T/N.
            bsdf.eval(forwardSd, bounceWi(bounce), L_NonDelta, beam);
            statEvals += beam.hwss_comps;
            if (opt.HWSSDoMIS) beam.spectralPDFs *= bsdf.lambdaPdf(forwardSd, bounceWi(bounce), L_NonDelta, beam, beam.wavelengths);
        }
        forwardUnwind(bounceIdx, px, py, beam);
        measure(solveData, bounceIdx, bounce.pathContribution, beam);
    }
    void NEE(PLTSolve& solveData, int bounceIdx, const PltptBounce& bounce, uint32_t px, uint32_t py, const TriangleHit& triangleHit, const ShadingData& sd, const VertexData& V, const MaterialInstance& bsdf, const BSDFProperties& bsdfProp) {
        const uint32_t lobes = bsdf.getLobes();
        RayFlags bf; bf.flags = bounce.flags;
        const bool allowMSOnCurrentVertex = kDoMSonNEE() && bounceIdx > 0 && !bf.lastBounceWasDelta() && bsdfProp.canManifoldWalkReflective();
        const bool sampleReflection = (lobes & L_NonDeltaReflection) != 0 || allowMSOnCurrentVertex;
        const bool sampleTransmission = (lobes & L_NonDeltaTransmission) != 0;
        if (!sampleReflection && !sampleTransmission) return;
        LightSample ls;
        const bool importanceSample = opt.doImportanceSampleEmitters;
        const bool valid = generateLightSample(sd, sampleReflection, sampleTransmission, solveData.sg, solveData.tilesg, solveData.wavelength, importanceSample, ls);
        if (!valid || !anyPositive(ls.beam.L)) return;
        if (!opt.HWSSDoMIS && bf.dispersiveDeltaSegment()) ls.beam.collapseToHeroWavelengthOnly();
        bool doMis = kDoMIS() && ls.lightType != LT_Analytic;
        bool evalCurrentVertex = true;
        const int maxMSBounces = int(maxBounces()) - bounceIdx;
        if (!kDoMSonNEE() || maxMSBounces <= 0) {
            // visibilityRay(RayDesc) with the light sample's own desc: TMin = 0 (LightsSamplingUtils.slang getVisibilityRay)
            if (scene->traceAny(ls.origin, ls.dir, .0f, ls.distance)) return;
            if (doMis) {
                const float bsdfPdf = bsdf.pdf(sd, ls.dir, ls.beam.wavelengths.x);
                ls.beam.scale(MISWeight(ls.pdf, bsdfPdf));
            }
        } else {
            const float3 lightDir = -ls.beam.r;
            float MISBsdfPdf = 0;
            ManifoldSampleData msresult;
            msresult.wavelength = ls.beam.wavelengths.x;
            const bool isReflection = dot(sd.N, lightDir) * dot(sd.N, sd.V) > .0f;
            const bool currentVertexMS = isReflection && allowMSOnCurrentVertex;
            if (currentVertexMS) {
                if (!manifoldSampleVertex(bounce.hit, sd, bsdf, lightDir, solveData.sg, msresult) || (msresult.hits[0].lobe & L_Reflection) == 0) return;
            }
            const float3 origin = !currentVertexMS ? V.posW : solveData.pathOrigin;
            const bool mneeOk = manifoldNEESample(ls.origin, ls.dir, .0f, ls.distance, origin, ls.beam.posW, lightDir, ls.beam.distant, maxMSBounces, solveData.sg, msresult);
            if (opt.debugView) {  // pltpt_solve.rt.slang:318-326
                writeDebugOutput(DV_MneeIterations, solveData, ccInferno(msresult.iterations / float(opt.MNEEMaxIterations)));
                if (msresult.hitCount > 0 && mneeOk) writeDebugOutput(DV_MneeBounces, solveData, msresult.hitCount == 1 ? float3(0, 0.3f, 1) : float3(1, 0, 0));
            }
            if (!mneeOk || msresult.pdf < 1e-6f) return;
            if (msresult.hitCount > 0) {
                if (!opt.HWSSDoMIS && msresult.dispersive) ls.beam.collapseToHeroWavelengthOnly();
                if (!ls.beam.distant) {
                    ls.dir = msresult.woLastHitToLight;
                    ls.beam.r = -msresult.woLastHitToLight;
                    ls.beam.t = perp_stark(ls.beam.r);
                }
                ls.beam.scale(1.f / msresult.pdf);
                manifoldForwardUnwind(msresult, origin, ls.beam);
                if (currentVertexMS) { evalCurrentVertex = false; MISBsdfPdf = bounce.lastNonDeltaPdf; }
            }
            if (doMis) {
                if (msresult.hitCount == 0 || !currentVertexMS) MISBsdfPdf = bsdf.pdf(sd, ls.dir, ls.beam.wavelengths.x);
                ls.beam.scale(MISWeight(ls.pdf, MISBsdfPdf * msresult.pdf));
            }
        }
        evalNEE(solveData, bounceIdx, bounce, px, py, V, sd.materialID, bsdf, ls.beam, evalCurrentVertex);
    }

    // pltpt_solve.rt.slang:401-476
    void solvePixel(uint32_t px, uint32_t py) {
        SampleGenerator sg(px, py, frameCount);
        SampleGenerator tilesg(px / opt.tileSize, py / opt.tileSize, frameCount << 2);
        const float samplingWavelength = SpectrumSampler::sampleOne(sampleNext1D(sg));
        PLTSolve solveData;
        solveData.sg = sg;
        solveData.tilesg = tilesg;
        solveData.pathOrigin = camPos();
        solveData.pathNormal = normalize(camW());
        solveData.wavelength = samplingWavelength;
        for (int p = 0;; ++p) {
            const PltptBounce bounce = bounceAt(px, py, p);
            RayFlags bf; bf.flags = bounce.flags;
            if (bf.miss()) { onMiss(solveData, p, bounce, px, py); break; }
            const TriangleHit triangleHit = scene->unpackHit(bounce.hit);
            const VertexData v = scene->getVertexData(triangleHit);
            const uint32_t materialID = scene->getMaterialID(triangleHit.instanceID);
            const ShadingData sd = scene->prepareShadingData(v, materialID, bounceWi(bounce));
            const MaterialInstance bsdf = scene->getMaterialInstance(sd);
            const BSDFProperties bsdfProp = bsdf.getProperties();
            if (kDoNEE()) NEE(solveData, p, bounce, px, py, triangleHit, sd, v, bsdf, bsdfProp);
            evalEmissive(solveData, p, bounce, px, py, bsdfProp, sd, triangleHit, opt.doImportanceSampleEmitters);
            if (opt.debugView && p == 0) {  // pltpt_solve.rt.slang:450-457
                writeDebugOutput(DV_Normals, solveData, sd.N);
                writeDebugOutput(DV_Albedo, solveData, bsdfProp.albedo);
                writeDebugOutput(DV_Roughness, solveData, float3(bsdfProp.roughness));
                writeDebugOutput(DV_UVs, solveData, float3(sd.uv.x, 0, sd.uv.y));
            }
            if (bf.terminated() || p == int(maxBounces())) break;
            solveData.pathOrigin = v.posW;
            solveData.pathNormal = sd.N;
        }
        const size_t o = (size_t(py) * width + px) * 4;
        float3 out = solveData.Lrgb;
        if (opt.debugView) {  // pltpt_solve.rt.slang:470-475
            const float4 dbg = solveData.debugOutput;
            const float3 dv = dbg.w > .0f ? float3(dbg.x, dbg.y, dbg.z) / dbg.w : float3(.0f);
            out = out + (dv - out) * opt.debugViewIntensity;
        }
        outColor[o] = out.x; outColor[o + 1] = out.y; outColor[o + 2] = out.z; outColor[o + 3] = 1.f;
    }
};

}  // namespace orc
