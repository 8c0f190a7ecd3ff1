// ORACLE — TEST INFRASTRUCTURE ONLY.
// ref_shims_end.h — included after oracle/_ref/ref_gen.h (the transpiled reference text) by both programs built on it (oracle/ref_compare.cpp,
// oracle/ref_frames.cpp): the services that need the transpiled types complete -- the texture unit, the emissive-sampler selection the pass
// makes by a define, and the shader tables of the two stages.  Hand-written plumbing; everything it calls is reference text.
#pragma once

namespace ref {
// the texture unit behind MaterialSystemStandIn::sampleTexture for Texture-mode handles: the oracle's filter over the oracle's texture table
static const orc::TextureSystem* gTextureService = nullptr;
inline float4 textureServiceFn(uint textureID, uint samplerID, float2 uv) {
    const orc::float4 v = gTextureService->sampleTexture((1u << 29) | textureID, samplerID, orc::float2(uv.x, uv.y), orc::float4(0.f));
    return float4(v.x, v.y, v.z, v.w);
}
// which transpiled emissive sampler LightSampling.slangh's `emissiveSampler` stands for: 0 = none (no sample drawn), 1 = uniform, 2 = power
static int gEmissiveKind = 0;
static EmissiveUniformSampler gUniformSampler;
static EmissivePowerSampler gPowerSampler;
template <class S> bool emissiveSamplerDispatch(float3 posW, float3 normalW, bool upperHemisphere, S& sg, float wavelength, bool importanceSampleSpectrum, TriangleLightSample& ls) {
    if (gEmissiveKind == 1) return gUniformSampler.sampleLight(posW, normalW, upperHemisphere, sg, wavelength, importanceSampleSpectrum, ls);
    if (gEmissiveKind == 2) return gPowerSampler.sampleLight(posW, normalW, upperHemisphere, sg, wavelength, importanceSampleSpectrum, ls);
    return false;
}
float emissiveEvalPdfDispatch(float3 posW, float3 normalW, bool upperHemisphere, TriangleLightHit hit) {
    if (gEmissiveKind == 1) return gUniformSampler.evalPdf(posW, normalW, upperHemisphere, hit);
    if (gEmissiveKind == 2) return gPowerSampler.evalPdf(posW, normalW, upperHemisphere, hit);
    return 0.f;
}
// the shader tables of the two stages (PLTPT.cpp:538-548), see TraceRay in ref_shims_late.h
inline void sbtClosestHit(Path& payload, BuiltInTriangleIntersectionAttributes attribs) { scatterTriangleMeshClosestHit(payload, attribs); }
inline void sbtMiss(Path& payload) { scatterMiss(payload); }
inline void sbtClosestHit(VisibilityRayData&, BuiltInTriangleIntersectionAttributes) {}   // no hit group for the visibility ray type: the payload stays "not visible"
inline void sbtMiss(VisibilityRayData& payload) { visibilityMiss(payload); }
inline void sbtClosestHit(ShadowRayHit& payload, BuiltInTriangleIntersectionAttributes attribs) { shadowTriangleMeshHit(payload, attribs); }
inline void sbtMiss(ShadowRayHit&) {}                                                      // no miss program for the shadow ray type: dist stays 0
}  // namespace ref
