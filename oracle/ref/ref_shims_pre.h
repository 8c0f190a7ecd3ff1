// ORACLE — TEST INFRASTRUCTURE ONLY.
// ref_shims_pre.h — what the transpiled reference modules (oracle/_ref/ref_gen.h) expect from the engine around them, declared before
// them: the compile-time HWSS width, the sample-generator free functions, and the two engine structs the BSDF code reads.
// Hand-written (engine plumbing, not PLT physics); every item cites the reference declaration it stands in for.
#pragma once
#include "slang_compat.h"

// Utils/HostDeviceShared.slangh:39-43,95-101 (shader branch; CONST_FUNCTION is dropped by the transpiler)
#define EXTRACT_BITS(bits, offset, value) (((value) >> (offset)) & ((1 << (bits)) - 1))
#define IS_BIT_SET(value, bitOffset) ((value & (1 << bitOffset)) != 0)
#define PACK_BITS(bits, offset, flags, value) ((((value) & ((1 << (bits)) - 1)) << (offset)) | ((flags) & (~(((1 << (bits)) - 1) << (offset)))))
#define PACK_BITS_UNSAFE(bits, offset, flags, value) (((value) << (offset)) | ((flags) & (~(((1 << (bits)) - 1) << (offset)))))
#define SETTER_DECL

namespace ref {

// HLSL float16_t and its vectors as storage types: 16-bit IEEE halves that convert to fp32 exactly when read
struct float16_t {
    uint16_t bits = 0;
    float16_t() = default;
    float16_t(float f) : bits(uint16_t(f32tof16(f))) {}
    operator float() const { return f16tof32(bits); }
};
inline uint16_t asuint16(float16_t h) { return h.bits; }
inline float16_t asfloat16(uint16_t b) { float16_t h; h.bits = b; return h; }
struct float16_t3 {
    float16_t x, y, z;
    float16_t3() = default;
    float16_t3(float f) : x(f), y(f), z(f) {}
    operator float3() const { return float3(float(x), float(y), float(z)); }
};
struct float16_t4 {
    float16_t x, y, z, w;
    float16_t4() = default;
    float16_t4(float f) : x(f), y(f), z(f), w(f) {}
    operator float4() const { return float4(float(x), float(y), float(z), float(w)); }
};

// Utils/HWSS.slang:33-37 with HWSS_SAMPLES = 4 (PLTPT's default `HWSS`, PLTPT.h:87) unless the translation unit is built for another count
// (-DREF_HWSS=n: the define PLTPT.cpp:226 passes)
#ifndef REF_HWSS
#define REF_HWSS 4
#endif
static const uint hwss_samples = REF_HWSS;
typedef vec<float, REF_HWSS> hwss_t;

// Utils/Sampling/SampleGeneratorInterface.slang:48-75: the generator only has to provide next() (32 random bits)
template <class S> float sampleNext1D(S& sg) {
    uint bits = sg.next();
    return (bits >> 8) * 0x1p-24f;
}
template <class S> float2 sampleNext2D(S& sg) {
    float2 sample;
    sample.x = sampleNext1D(sg);
    sample.y = sampleNext1D(sg);
    return sample;
}
template <class S> float3 sampleNext3D(S& sg) {
    float3 sample;
    sample.x = sampleNext1D(sg);
    sample.y = sampleNext1D(sg);
    sample.z = sampleNext1D(sg);
    return sample;
}

float3 computeRayOrigin(float3 pos, float3 normal);  // Utils/Geometry/GeometryHelpers.slang:47-61, transpiled further down
// Scene/ShadingData.slang:46-107: the fields, the two frame transforms and the ray-origin offset the path uses
struct ShadingData {
    float3 posW, V, N, T, B, faceN;
    float4 tangentW;
    float2 uv;
    bool frontFacing = true;
    uint materialID = 0;
    float opacity = 1.f;
    float3 computeNewRayOrigin(bool viewside = true) { return computeRayOrigin(posW, (frontFacing == viewside) ? faceN : -faceN); }  // :77-80
    float3 fromLocal(float3 v) const { return T * v.x + B * v.y + N * v.z; }
    float3 toLocal(float3 v) const { return float3(dot(v, T), dot(v, B), dot(v, N)); }
};

// Rendering/Materials/LobeType.slang:36-68
enum class LobeType : uint {
    None = 0x00, DiffuseReflection = 0x01, ScatterReflection = 0x02, DeltaReflection = 0x04, DispersiveDeltaReflection = 0x08,
    DiffuseTransmission = 0x10, ScatterTransmission = 0x20, DeltaTransmission = 0x40, DispersiveDeltaTransmission = 0x80,
    Diffuse = 0x11, Scatter = 0x22, NonDispersiveDelta = 0x44, DispersiveDelta = 0x88, Delta = 0xCC, NonDelta = 0x33,
    Reflection = 0x0f, Transmission = 0xf0, NonDeltaReflection = 0x03, NonDeltaTransmission = 0x30, ClientSide = 0xff, All = 0xffff, MaterialSpecifc = 0xff00,
};
static const uint LobeTypeMaterialSpecificOffset = 8;
// Rendering/Materials/IMaterialInstance.slang:44-60
enum class BSDFEvaluationFlags : uint { None = 0 };
static const float QcutoffForMS = .15f;

// Scene/Scene.slang:251-294 is transpiled (gScene.evalSpectrum); the table it reads is set by the caller of the comparison
struct SpectralProfile {  // Scene/Material/BasicMaterialData.slang:43-54
    float3 rgb;
    float intensity;
    uint bins;
    float minWavelength, maxWavelength;
    float pdf[128];
    float icdf[128];
};

// HLSL resource types and the few engine types LightCollection.slang / the emissive samplers name
template <class T> struct StructuredBuffer { const T* p = nullptr; T operator[](uint i) const { return p[i]; } };  // a load returns a copy
template <class T> struct Buffer { const T* p = nullptr; T operator[](uint i) const { return p[i]; } };
struct MeshLightData { uint instanceID = 0, triangleOffset = 0, triangleCount = 0, materialID = 0; };  // Scene/Lights/MeshLightData.slang:36-42 (not read by the path)
struct GeometryInstanceID { uint index = 0; };                                                          // Scene/SceneTypes.slang:57-70
struct SamplerState { uint id = 0; };  // the sampler's index in the material system's table, so that the texture service can honour it
// HLSL Texture2D<T>: texel loads and SampleLevel are services of the texture unit -- provided by the caller of the comparison (the oracle's
// importance-map mips and its fp32 bilinear filter, plt_envmap.h), not part of the text under comparison
template <class T> struct Texture2D {
    const void* ctx = nullptr;
    T (*loadFn)(const void* ctx, int x, int y, int mip) = nullptr;
    T (*sampleFn)(const void* ctx, float2 uv, float lod) = nullptr;
    uint width = 0, height = 0;
    T Load(int3 p) const { return loadFn(ctx, p.x, p.y, p.z); }
    T operator[](uint2 p) const { return loadFn(ctx, int(p.x), int(p.y), 0); }
    T SampleLevel(SamplerState, float2 uv, float lod) const { return sampleFn(ctx, uv, lod); }
    void GetDimensions(uint& w, uint& h) const { w = width; h = height; }
};
// HLSL ByteAddressBuffer (FinalizeIntegration.cs.slang reads the rasterised texel sums of textured emitters through it)
struct ByteAddressBuffer { const uint* p = nullptr; uint Load(uint address) const { return p[address / 4]; } uint4 Load4(uint address) const { return uint4(p[address / 4], p[address / 4 + 1], p[address / 4 + 2], p[address / 4 + 3]); } };
typedef vec<uint64_t, 4> uint64_t4;
struct ExplicitLodTextureSampler { float lod; ExplicitLodTextureSampler(float l = 0.f) : lod(l) {} };
typedef ExplicitLodTextureSampler ITextureSampler;  // the interface the loaders take; the pass passes ExplicitLodTextureSampler{0} (Scene/Material/TextureSampler.slang:74-87)

// ---- Rendering/Lights/EnvMapSamplerSetup.cs.slang:36-46: constant buffer and resources of the importance-map setup pass (bound by EnvMapSampler.cpp:101-112)
static uint2 outputDim = uint2(0u, 0u), outputDimInSamples = uint2(0u, 0u), numSamples = uint2(0u, 0u);
static float invSamples = 0.f;
static SamplerState gEnvSampler;
static Texture2D<float4> gEnvMap;
template <class T> struct RWTexture2D { T* p = nullptr; uint width = 0; T& operator[](uint2 px) { return p[size_t(px.y) * width + px.x]; } };
static RWTexture2D<float> gImportanceMap;
// ---- what RenderPasses/PLTPT/LightSampling.slangh expects from the pass around it
// pltpt.slangh:95-118: the pass's compile-time switches (defines set by PLTPT.cpp:222-260) -- variables here so that the comparison can
// walk through their combinations; pltpt.slangh:77-78: the two sourcing constants of the constant buffer
static bool kUseEnvLight = false, kUseEmissiveLights = false, kUseAnalyticLights = true, kNEEUsePerTileSelectorSG = false;
static const bool kForceFullyCoherent = false;
static float kSourcingAreaFromEmissiveGeometry = 1.f, kSourcingMaxBeamOmega = .025f;
static const float kRayTMax = 3.402823466e+38f;  // ShadowRays.slangh:47 (= FLT_MAX)
struct RayDesc { float3 Origin; float TMin = 0.f; float3 Direction; float TMax = 0.f; };  // HLSL built-in (LightSample::getVisibilityRay, Ray::toRayDesc)
// Rendering/Lights/EnvMapSampler.slang:46-52 and EmissiveLightSamplerInterface.slang:38-49: the sample records; the two samplers
// themselves need the texture system / the light collection and are outside this comparison -- they stand in as "no sample drawn"
struct TriangleLightSample { uint triangleIndex = 0; float3 posW, normalW, dir; float distance = 0.f; hwss_t Le, wavelengths; float pdf = 0.f; float2 uv; };
struct TriangleLightHit { uint triangleIndex = 0; float3 posW, normalW; };  // EmissiveLightSamplerInterface.slang:54-59
// `emissiveSampler` is EmissiveUniformSampler or EmissivePowerSampler by a define (EmissiveLightSampler.slang:36-52); both are transpiled
// further down, so the stand-in forwards to emissiveSamplerDispatch (ref_shims_post.h, found by argument-dependent lookup at instantiation),
// which calls the transpiled sampler the comparison has selected, or draws no sample
float emissiveEvalPdfDispatch(float3 posW, float3 normalW, bool upperHemisphere, struct TriangleLightHit hit);  // ref_shims_post.h
struct EmissiveSamplerStandIn {
    template <class S> bool sampleLight(float3 posW, float3 normalW, bool upperHemisphere, S& sg, float wavelength, bool importanceSampleSpectrum, TriangleLightSample& ls) {
        return emissiveSamplerDispatch(posW, normalW, upperHemisphere, sg, wavelength, importanceSampleSpectrum, ls);
    }
    float evalPdf(float3 posW, float3 normalW, bool upperHemisphere, TriangleLightHit hit) { return emissiveEvalPdfDispatch(posW, normalW, upperHemisphere, hit); }
};
static EmissiveSamplerStandIn emissiveSampler;

// ---- what `measure` (pltpt_solve.rt.slang:96-147) expects (the camera block is the transpiled CameraData.slang / Camera.slang): the solve payload (Path.slang:200-225, without its generators), two more switches of pltpt.slangh:95-104
static bool kUseDirectLights = true, mHWSSDoMIS = true;
// ---- what the closest-hit program of pltpt_sample.rt.slang expects: more switches of pltpt.slangh:90-118, the pass's resources
// (pltpt.slangh:84-88,123) as plain arrays, the DXR system values of the current hit, the per-path generator (Utils/Sampling/SampleGenerator:
// the pass's generator is the xoshiro the oracle restates and pins to PseudorandomTests.cpp; both sides step the same object type)
static uint mMaxBounces = 8, mTileSize = 512;
static bool kAlphaMasking = true, kDoRussianRoulette = true;
typedef uint4 PackedHitInfo;  // Scene/HitInfo.slang:63 (uncompressed hit info)
template <class T> struct RWStructuredBuffer { T* p = nullptr; T& operator[](uint i) { return p[i]; } };
static RWTexture2D<float4> gOutputColor, gOutputNormal, gOutputAlbedo;
struct BuiltInTriangleIntersectionAttributes { float2 barycentrics; };
// Utils/Sampling/SampleGenerator: SampleGenerator(pixel, sampleNumber) -- the pass's generator is the xoshiro the oracle restates and pins to
// PseudorandomTests.cpp; both sides step the same object type
struct SampleGenerator {
    orc::SampleGenerator g;
    SampleGenerator() = default;
    SampleGenerator(uint2 pixel, uint sampleNumber) : g(pixel.x, pixel.y, sampleNumber) {}
    uint next() { return g.next(); }
};
static bool kDoNEE = true, kDoMSonNEE = true, kDoMIS = true, kImportanceSampleEmissiveSpectra = true;   // pltpt.slangh:105-116
#ifndef REF_MNEE_OCCLUDERS
#define REF_MNEE_OCCLUDERS 2
#endif
static constexpr uint kMNEEMaxMSOccluders = REF_MNEE_OCCLUDERS;   // Manifolds.slangh:40-54 (MNEE_MAX_MS_OCCLUDERS: an array size, hence per build; default PLTPT.h:123)
static int kMNEEMaxIterations = 40;
static float kMNEESolverThreshold = 1e-5f;
static const float kShadowRayTMin = 1e-4f;                      // ShadowRays.slangh:46
// ---- the DXR runtime around the three programs of a stage (HLSL built-ins): launch index, constant buffer (pltpt.slangh:60-75), the
// acceleration structure handle.  TraceRay itself is in ref_shims_late.h.
static uint2 kTile = uint2(0u, 0u), kOutputSize = uint2(0u, 0u);
static uint gFrameCount = 0;
static uint3 gDispatchRaysIndex = uint3(0u, 0u, 0u);
inline uint3 DispatchRaysIndex() { return gDispatchRaysIndex; }
struct RtAccelStandIn {};
static constexpr uint RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH = 0x4u;
static constexpr uint defaultVisibilityRayId = 0, defaultShadowRayId = 1;   // ShadowRays.slangh:37-44 with PLTPT.cpp:60-61,221-222
static constexpr uint rayTypeCount = 2;                         // PLTPT.cpp:60-61: visibility and shadow ray types (the sample stage has one)
// ---- RenderPasses/AccumulatePass/Accumulate.cs.slang:38-68: constant buffer and resources of the accumulation pass (float input format)
static uint2 gResolution = uint2(0u, 0u);
static uint gAccumCount = 0;
static bool gAccumulate = true, gMovingAverageMode = false;
typedef float4 InputDataType;
static Texture2D<InputDataType> gCurFrame;
static RWTexture2D<float4> gOutputFrame, gLastFrameSum, gLastFrameCorr;
static RWTexture2D<uint4> gLastFrameSumLo, gLastFrameSumHi;
// DO_DEBUG_VIEW / DEBUG_VIEW_TYPE (PLTPT.cpp:230-231): the pass compiles the debug overlay in only when a view is selected; the two builds of
// oracle/ref_frames.cpp (-DREF_DEBUG_VIEW=0 / 1) are those two shader variants
#ifndef REF_DEBUG_VIEW
#define REF_DEBUG_VIEW 0
#endif
#define DO_DEBUG_VIEW REF_DEBUG_VIEW
static uint kDebugViewType = 0;
static float kDebugViewIntensity = 1.f;   // pltpt.slangh:85-87

}  // namespace ref
