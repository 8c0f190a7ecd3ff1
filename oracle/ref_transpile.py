#!/usr/bin/env python
"""ORACLE — TEST INFRASTRUCTURE ONLY.

ref_transpile.py — compiles the reference's own PLT shader text for the CPU.

`slangc` is not available offline, so the reference's Slang modules cannot be built with their own compiler.  Their
physics core is, however, plain C-like code.  This script reads the modules listed in SOURCES from /root/reference where they lie,
applies the purely TEXTUAL rewrite rules below (no statement of the reference is re-typed by hand, nothing is re-ordered inside a
function), and writes the result to oracle/_ref/ref_gen.h (git-ignored: reference text never enters the repository).  Together with
oracle/ref/slang_compat.h (HLSL vector / matrix types and intrinsics in C++) and oracle/ref/ref_shims_{pre,mid,late,end}.h (the engine services
the modules call: gScene.evalSpectrum on a caller-provided table, the sample generator, ShadingData, Slang's dynamic material dispatch written
out as a switch, the DXR runtime -- TraceRay over a caller-provided intersection search with the pass's shader tables --, constant buffers
and resources of the compute passes) it is compiled by `make -C oracle ref` into
  * oracle/_ref/ref_compare -- runs the oracle (namespace orc) and the transpiled reference (namespace ref) on the same random inputs, from
    single functions up to whole frames of random scenes, and prints the largest deviation per entry (tests/test_oracle_ref.py);
  * oracle/_ref/libref_frames*.so (oracle/ref_frames.cpp) -- the reference's two ray-generation programs, its AccumulatePass and its
    light-collection build passes over scenes in the C ABI's layouts, one build per shader variant (HWSS_SAMPLES 1-4, DO_DEBUG_VIEW,
    MNEE_MAX_MS_OCCLUDERS): the generator of the fixtures tests/golden/ref_*.npz (tools/make_golden_ref.py).

Rewrite rules (all mechanical; each is a regular expression or a brace-matched block move):
  R1  strip comments-only license headers, `import`, `__exported import`, `#include "*.slangh"`, `#pragma once`
  R2  floating literals get an `f` suffix (Slang literals are fp32; C++ ones would silently compute in double)
  R3  `__init(` -> constructor of the enclosing struct (+ a defaulted default constructor); `this.` -> `this->`
  R4  `[mutating]`, `[unroll]`, `[ForceInline]`... attributes dropped; `no_diff` dropped
  R5  `property NAME : T { get { BODY } }` -> `T NAME() { BODY }`; uses listed in PROPERTY_USES get `()`
  R6  generics `RET NAME<S : ISampleGenerator>(` -> `template <class S> RET NAME(`
  R7  `out T x` / `inout T x` parameters -> `T& x`
  R8  multi-component swizzles `.xy` -> `.xy()` (read-only in these modules); `rgb.r/.g/.b` -> `.x/.y/.z`; `.x` on scalars dropped
  R9  `A.B` static access -> `A::B` for the names in STATIC_SCOPES; `struct X : IFoo` loses the interface list
  R10 `static const float` -> `static constexpr float`; struct-typed `const` locals / parameters lose `const` (Slang methods are
      callable on them; C++ would need const-qualified methods), structs get their trailing `;`
      A property with a setter also gets `void set_NAME(T newValue) { BODY }`; in the files that use the RayFlags / BounceData properties,
      `x.flags.NAME = e;` -> `x.flags.set_NAME(e);`, `x.flags.NAME |= e;` -> `x.flags.set_NAME(x.flags.NAME() | (e));`, reads get `()`
  R12 `let x = e` -> `auto x = e`
  R13 `enum class X : T` -> `enum X : T` (Slang converts a scoped enumerator to its underlying type in `switch (uint)` / `uint(...)`; C++
      only does so for an unscoped enumeration, whose enumerators can still be written `X::Name`)
  R14 the entry points `void main(...)` of the sample shader, the solve shader and the environment-map setup pass -> `sampleRayGen()` /
      `solveRayGen()` / `envMapSamplerSetup(dispatchThreadID)`; `: SV_...` parameter semantics dropped
  R11 the handful of vector-condition ternaries (`c ? a : b` with bool2 `c`, DiffractionGrating.slang only) -> select(c, a, b); listed in PATCHES
      with the exact source text they replace -- the script fails if a patch no longer matches.
"""
import os
import re
import sys

REF = "/root/reference/Source/Falcor"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# (path, None = whole file | list of top-level items to extract by name)
SOURCES = [
    ("Utils/Math/MathConstants.slangh", "defines"),
    ("Utils/Math/MathHelpers.slang", ["sqr", "norm2", "factorial", "c_t", "cnorm2", "cabs", "csqrt", "cconj", "creal", "cimag", "cmul", "cdiv", "csqr", "carg", "sinc",
                                      "sample_disk", "sample_cone", "sample_sphere", "sample_disk_concentric", "sample_cosine_hemisphere_concentric", "cosine_hemisphere_concentric_pdf",
                                      "sample_triangle", "sample_boxmueller", "boxmueller_pdf", "sample_boxmueller_truncated", "boxmueller_truncated_pdf", "tryInvert", "inverse",
                                      "diag2", "ident2", "diag4", "ident4", "svd_symmetric", "perp_stark", "BesselJ_saa", "BesselJ_laa", "BesselJ", "ndir_to_oct_equal_area_snorm", "oct_to_ndir_equal_area_snorm",
                                      "ndir_to_oct_equal_area_unorm", "oct_to_ndir_equal_area_unorm", "oct_wrap", "ndir_to_oct_snorm", "oct_to_ndir_snorm", "world_to_latlong_map", "latlong_map_to_world"]),
    ("Utils/Math/FormatConversion.slang", ["floatToSnorm16", "unpackSnorm16", "packSnorm16", "unpackSnorm2x16", "packSnorm2x16"]),
    ("Utils/Math/PackedFormats.slang", ["encodeNormal2x16", "decodeNormal2x16"]),
    ("Scene/Lights/LightCollectionShared.slang", None),
    ("Scene/Lights/LightCollection.slang", None),
    ("Scene/Lights/EnvMapData.slang", None),
    ("Scene/Lights/EnvMap.slang", None),
    ("Scene/SceneTypes.slang", ["StaticVertexData", "PackedStaticVertexData"]),
    ("Utils/Color/ColorHelpers.slang", ["luminance", "XYZtoRGB_Rec709"]),
    ("Utils/Color/SpectrumUtils.slang", None),
    ("Rendering/Materials/Fresnel.slang", None),
    ("Rendering/Materials/PLT/Mueller.slang", None),
    ("Rendering/PLT.slang", None),
    ("Rendering/Materials/PLT/Interference.slang", None),
    ("Rendering/Materials/PLT/FractalSurface.slang", None),
    ("Rendering/Materials/PLT/GaussianSurface.slang", None),
    ("Rendering/Materials/PLT/Birefringence.slang", None),
    ("Utils/Color/SpectrumSampler.slang", None),
    ("Scene/Lights/LightData.slang", None),
    ("Scene/Camera/CameraData.slang", None),
    ("Utils/Math/Ray.slang", ["Ray"]),
    ("Scene/Camera/Camera.slang", None),
    ("Scene/Material/MaterialTypes.slang", None),
    ("Scene/Material/MaterialData.slang", None),
    ("Scene/Material/TextureHandle.slang", ["TextureHandle"]),
    ("Scene/Material/BasicMaterialData.slang", ["BasicMaterialData"]),
    ("literal", '#include "../ref/ref_shims_mid.h"'),
    ("Scene/Material/ShadingUtils.slang", ["ShadingFrame", "rgbToNormal", "rgToNormal", "computeShadingFrameFromNormalMap", "applyNormalMap", "computeTangentSpace"]),
    ("Scene/Scene.slang", ("members", "Scene", ["getLightCount", "evalSpectrumRGB", "evalSpectrum", "sampleSpectrum", "getLight", "getVertex", "getFaceNormalInObjectSpace", "computeVertexData",
                                                "computeFaceNormalAndAreaW", "getVertexPositionsW", "getVertexTexCoords"],
                           "RtAccelStandIn rtAccel; SpectralProfile* spectralProfiles = nullptr; uint lightCount = 0; LightData* lights = nullptr; Camera camera; LightCollection lightCollection; MaterialSystemStandIn materials; EnvMap envMap; "
                           "StructuredBuffer<PackedStaticVertexData> vertices; uint3 (*indexFn)(GeometryInstanceID, uint) = nullptr; "
                           "uint3 getIndices(GeometryInstanceID instanceID, uint triangleIndex) { return indexFn(instanceID, triangleIndex); }  /* index-buffer fetch: engine service */ "
                           "VertexData (*vertexDataFn)(TriangleHit) = nullptr; VertexData getVertexData(TriangleHit hit) { return vertexDataFn(hit); }  /* instance -> vertices, matrices: composed by the caller from getVertex / computeVertexData */ "
                           "uint (*materialIDFn)(GeometryInstanceID) = nullptr; uint getMaterialID(GeometryInstanceID instanceID) { return materialIDFn(instanceID); } "
                           "float4x4 (*worldMatrixFn)(GeometryInstanceID) = nullptr; float4x4 getWorldMatrix(GeometryInstanceID instanceID) { return worldMatrixFn(instanceID); }  /* Scene.slang:178-182: matrix buffer load */ "
                           "bool (*worldFrontFaceCWFn)(GeometryInstanceID) = nullptr; bool isWorldFrontFaceCW(GeometryInstanceID instanceID) { return worldFrontFaceCWFn(instanceID); }  /* Scene.slang:212-215: instance flag */")),
    ("Rendering/Materials/BSDFSample.slang", None),
    ("Rendering/Utils/DifferentialShadingGeometry.slang", ["float3pdv", "float3kpdv", "DifferentialShadingGeometry", "reflectDerivatives", "refractDerivatives", "differentialShadingGeometry"]),
    ("Rendering/Materials/PLT/DiffractionGrating.slangh", None),
    ("Rendering/Materials/PLT/DiffractionGrating.slang", None),
    ("Rendering/Materials/PLT/PLTDiffuseBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTConductorBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTOpaqueDielectricBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTDielectricBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTThinDielectricBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTCoatedConductorBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTCoatedOpaqueDielectricBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTDiffractionGratedConductorBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTDiffractionGratedOpaqueDielectricBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTMultiLayeredStackBSDF.slang", None),
    ("Rendering/Materials/PLT/PLTDiffuseMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTConductorMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTDielectricMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTOpaqueDielectricMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTThinDielectricMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTCoatedConductorMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTCoatedOpaqueDielectricMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTDiffractionGratedConductorMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTDiffractionGratedOpaqueDielectricMaterial.slang", None),
    ("Rendering/Materials/PLT/PLTMultiLayeredStackMaterial.slang", None),
    ("Rendering/Lights/LightHelpers.slang", None),
    ("Rendering/Lights/EmissiveLightSamplerHelpers.slang", None),
    ("Rendering/Lights/EmissiveUniformSampler.slang", None),
    ("Rendering/Lights/EmissivePowerSampler.slang", None),
    ("Rendering/Lights/EnvMapSampler.slang", None),
    ("literal", "static EnvMapSampler envMapSampler;  // cbuffer CB of the pass (pltpt.slangh:73)"),
    ("Rendering/Lights/EnvMapSamplerSetup.cs.slang", ["main"]),
    ("Utils/Geometry/GeometryHelpers.slang", ["computeRayOrigin"]),
    ("Rendering/Utils/LightsSamplingUtils.slang", None),
    ("../RenderPasses/PLTPT/ShadowRays.slangh", ["computeNewRayOrigin"]),
    ("../RenderPasses/PLTPT/LightSampling.slangh", None),
    ("../RenderPasses/PLTPT/Path.slang", ["RayFlags", "BounceData", "Path", "PLTSolve"]),
    ("../RenderPasses/PLTPT/DebugViewType.slangh", None),
    ("../RenderPasses/PLTPT/ColourCoding.slang", None),
    ("../RenderPasses/PLTPT/pltpt_debug.slangh", None),
    ("../RenderPasses/PLTPT/pltpt_solve.rt.slang", ["MISWeight", "measure"]),
    ("literal", '#include "../ref/ref_shims_late.h"'),
    ("../RenderPasses/PLTPT/ShadowRays.slangh", ["VisibilityRayData", "visibilityMiss", "visibilityRay", "ShadowRayHit", "shadowTriangleMeshHit", "shadowRay"]),
    ("../RenderPasses/PLTPT/pltpt.slangh", ["readBounce", "appendBounce", "RR"]),
    ("../RenderPasses/PLTPT/pltpt_sample.rt.slang", ["shouldTerminatePath", "pltSampleBSDF", "scatterMiss", "scatterTriangleMeshClosestHit", "main"]),
    ("../RenderPasses/PLTPT/pltpt_solve.rt.slang", ["forwardHWSSMis", "forwardUnwindStep", "forwardUnwind", "wasNEEAvailable", "evalEmissive"]),
    ("../RenderPasses/PLTPT/Manifolds.slangh", ["sphcoords", "sphcoordsDerivatives", "manifoldOneStep", "MSHitData", "ManifoldSampleData", "manifoldSampleVertex", "traceMNEEOccluders",
                                                "MNEEStep", "manifoldNEESample"]),
    ("../RenderPasses/PLTPT/pltpt_solve.rt.slang", ["manifoldForwardUnwind", "evalNEE", "NEE", "onMiss", "main"]),
    ("literal", "// BuildTriangleList.cs.slang:33-43 / FinalizeIntegration.cs.slang:33-42: constant buffers and resources of the two light-collection passes (one declaration per name)\n"
                "static uint gLightIdx = 0, gMaterialID = 0, gInstanceID = 0, gTriangleCount = 0, gTriangleOffset = 0;\n"
                "static RWStructuredBuffer<PackedEmissiveTriangle> gTriangleData;\nstatic RWStructuredBuffer<EmissiveFlux> gFluxData;\n"
                "static SamplerState gPointSampler;\nstatic ByteAddressBuffer gTexelMax, gTexelSum;   // results of the texture-space rasterisation of textured emitters (EmissiveIntegrator.3d.slang): not produced here"),
    ("Scene/Lights/BuildTriangleList.cs.slang", ["buildTriangleList"]),
    ("Scene/Lights/FinalizeIntegration.cs.slang", ["asuint64", "finalizeIntegration"]),
    ("../RenderPasses/AccumulatePass/Accumulate.cs.slang", ["accumulateSingle", "accumulateSingleCompensated", "accumulateDouble"]),
]

# the two ray-generation programs are both called `main`: renamed after their file (R14)
ENTRY_POINTS = {"PLTPT/pltpt_sample.rt.slang": "sampleRayGen", "PLTPT/pltpt_solve.rt.slang": "solveRayGen", "Lights/EnvMapSamplerSetup.cs.slang": "envMapSamplerSetup"}

STATIC_SCOPES = ["SpectrumUtils", "Birefringence", "SpectrumSampler", "BasicMaterialData", "BSDFProperties"]
# property reads outside the declaring struct: `roughness` is a property of both surface types, `q` of FractalSurface only (a plain
# field of GaussianSurface) -- decided per file by the surface type the file's BxDF declares (`typedef ... SurfaceType`)
PROPERTY_USES = [r"gfs\.roughness"]
PROPERTY_USES_FRACTAL = [r"gfs\.q"]
# exact-text patches: (file suffix, old, new, why)
PATCHES = [
    ("Utils/Color/SpectrumUtils.slang", "static const float3 rgbWavelengths = float3(630,532,467);", "static inline const float3 rgbWavelengths = float3(630,532,467);",
     "C++ needs `inline` for a class-type static data member defined in the class"),
    ("Utils/Color/ColorHelpers.slang", "static const float3x3 M =\n    {", "static const float3x3 M =\n    float3x3{", "list-initialisation of the matrix type"),
    # R11: vector-condition ternaries -> select()
    ("Utils/Color/SpectrumSampler.slang", "hwss -= hwss > hwss_t(maxWavelength) ? hwss_t(spanWavelength) : hwss_t(.0f);",
     "hwss -= select(hwss > hwss_t(maxWavelength), hwss_t(spanWavelength), hwss_t(.0f));", "R11"),
    ("Scene/Scene.slang", "return zeros ? hwss_t(.0f) : sp.intensity * lerp(a, b, w);", "return select(zeros, hwss_t(.0f), sp.intensity * lerp(a, b, w));", "R11"),
    ("PLT/DiffractionGrating.slang", "const float2 sini = li2 > FLT_EPSILON ?  wiLocal.xy / sqrt(li2) : .0f;", "const float2 sini = select(li2 > FLT_EPSILON,  wiLocal.xy / sqrt(li2), .0f);", "R11"),
    ("PLT/DiffractionGrating.slang", "const float2 sino = lo2 > FLT_EPSILON ? -woHint.xy  / sqrt(lo2) : .0f;", "const float2 sino = select(lo2 > FLT_EPSILON, -woHint.xy  / sqrt(lo2), .0f);", "R11"),
    ("PLT/DiffractionGrating.slang", "lobe = alpha==.0f && lobe==0 ? sign(l) : lobe;", "lobe = int2(select(bool2(alpha==.0f) && lobe==0, sign(l), lobe));", "R11 (+ the implicit float2 -> int2 conversion of the assignment)"),
    ("PLT/DiffractionGrating.slang", "const bool2 select = abs(rand)<float2(cdf) && lobe==0;\n            pdf = select ? float2(p) : pdf;\n            lobe = select ? int2(l+1) : lobe;",
     "const bool2 sel = abs(rand)<float2(cdf) && lobe==0;\n            pdf = select(sel, float2(p), pdf);\n            lobe = select(sel, int2(l+1), lobe);", "R11 (the local variable is called `select` in the reference)"),
    ("PLT/DiffractionGrating.slang", "lobe = rand<.0f ? -lobe : lobe;", "lobe = select(rand<.0f, -lobe, lobe);", "R11"),
    ("PLT/DiffractionGrating.slang", "const float2 sini = p>FLT_EPSILON ? wi.xy/p : .0f;", "const float2 sini = select(p>FLT_EPSILON, wi.xy/p, .0f);", "R11"),
    ("PLT/DiffractionGrating.slang", "const float2 invp = p>FLT_EPSILON ? 1.f/p : .0f;", "const float2 invp = select(p>FLT_EPSILON, 1.f/p, .0f);", "R11"),
    ("Utils/Math/MathHelpers.slang", "return (1.f - abs(v.yx)) * (v.xy >= 0.f ? 1.f : -1.f);", "return (1.f - abs(v.yx)) * select(v.xy >= 0.f, 1.f, -1.f);", "R11"),
    ("Utils/Math/MathHelpers.slang", "n.xy = (n.z < 0.0) ? oct_wrap(n.xy) : n.xy;", "{ const float2 nxy = (n.z < 0.0) ? oct_wrap(n.xy) : n.xy; n.x = nxy.x; n.y = nxy.y; }",
     "swizzle on the left-hand side (R8 models read-only swizzles)"),
    ("Lights/LightCollectionShared.slang", "        posAndTexCoords[0].xyz = tri.posW[0];\n        posAndTexCoords[1].xyz = tri.posW[1];\n        posAndTexCoords[2].xyz = tri.posW[2];",
     "        for (int i = 0; i < 3; i++) { posAndTexCoords[i].x = tri.posW[i].x; posAndTexCoords[i].y = tri.posW[i].y; posAndTexCoords[i].z = tri.posW[i].z; }",
     "swizzle on the left-hand side (R8 models read-only swizzles)"),
    ("Scene/SceneTypes.slang", "v.tangent.xyz = decodeNormal2x16(asuint(packedNormalTangentCurveRadius.z));",
     "{ const float3 txyz = decodeNormal2x16(asuint(packedNormalTangentCurveRadius.z)); v.tangent.x = txyz.x; v.tangent.y = txyz.y; v.tangent.z = txyz.z; }",
     "swizzle on the left-hand side"),
    ("Scene/SceneTypes.slang", "#else // !HOST_CODE\n    [mutating] void pack(const StaticVertexData v)", "#elif 0\n    [mutating] void pack(const StaticVertexData v)",
     "the shader-side `pack` is not used by the path (vertices arrive packed); only `unpack` is compared"),
    ("Scene/Lights/EnvMap.slang", "return envMap[coord].rgb * getIntensity();", "return envMap[coord].xyz * getIntensity();", "`.rgb` of a texel load (R8 knows xyzw swizzles)"),
    ("Scene/Lights/EnvMap.slang", "Texture2D       envMap;", "Texture2D<float4> envMap;", "HLSL's bare `Texture2D` is Texture2D<float4>; a C++ member needs the argument spelled out"),
    ("PLT/PLTDiffractionGratedConductorMaterial.slang", "const float2 invPitch = pitch>.0f ? 1.f/pitch : .0f;", "const float2 invPitch = select(pitch>.0f, 1.f/pitch, .0f);", "R11"),
    ("PLT/PLTDiffractionGratedOpaqueDielectricMaterial.slang", "const float2 invPitch = pitch>.0f ? 1.f/pitch : .0f;", "const float2 invPitch = select(pitch>.0f, 1.f/pitch, .0f);", "R11"),
    ("Scene/Material/ShadingUtils.slang", "    __init() {}\n", "", "the transpiler already emits a defaulted default constructor for a struct with __init (R3)"),
    ("Scene/Material/ShadingUtils.slang", "n.xy = rg * 2.f - 1.f;", "{ const float2 nxy = rg * 2.f - 1.f; n.x = nxy.x; n.y = nxy.y; }", "swizzle on the left-hand side"),
    ("Scene/Material/ShadingUtils.slang", "mapN = rgToNormal(encodedNormal.rg);", "mapN = rgToNormal(encodedNormal.xy);", "colour-named swizzle (R8)"),
    ("Scene/Scene.slang", "v.tangentW.xyz = mul((float3x3)worldMat, v.tangentW.xyz);", "{ const float3 txyz = mul((float3x3)worldMat, v.tangentW.xyz); v.tangentW.x = txyz.x; v.tangentW.y = txyz.y; v.tangentW.z = txyz.z; }",
     "swizzle on the left-hand side"),
    ("Scene/Scene.slang", "v.tangentW.xyz = normalize(v.tangentW.xyz);", "{ const float3 txyz = normalize(v.tangentW.xyz); v.tangentW.x = txyz.x; v.tangentW.y = txyz.y; v.tangentW.z = txyz.z; }", "swizzle on the left-hand side"),
    ("Scene/Scene.slang", "void getVertexPositionsW(const GeometryInstanceID instanceID, const uint triangleIndex, out float3 p[3])",
     "void getVertexPositionsW(const GeometryInstanceID instanceID, const uint triangleIndex, float3 p[3])", "an array parameter is already passed by reference in C++"),
    ("Scene/Scene.slang", "void getVertexTexCoords(const GeometryInstanceID instanceID, const uint triangleIndex, out float2 texC[3])",
     "void getVertexTexCoords(const GeometryInstanceID instanceID, const uint triangleIndex, float2 texC[3])", "an array parameter is already passed by reference in C++"),
    ("PLTPT/Manifolds.slangh", "const float2x2 do = sphcoordsDerivatives(w);", "const float2x2 do_ = sphcoordsDerivatives(w);", "`do` is a C++ keyword"),
    ("PLTPT/Manifolds.slangh", "const float2x2 dCdX = do - float2x2(dio);", "const float2x2 dCdX = do_ - float2x2(dio);", "`do` is a C++ keyword"),
    ("PLTPT/ColourCoding.slang", "const int l = clamp((int)val, 0, int(bins)-1);", "const int l = clamp(hlsl_ftoi(val), 0, int(bins)-1);",
     "HLSL's float -> int conversion saturates and maps NaN to 0; C++'s is undefined outside int's range (the views feed unbounded values)"),
    ("PLTPT/ColourCoding.slang", "const int h = clamp((int)(val + 1), 0, int(bins) - 1);", "const int h = clamp(hlsl_ftoi(val + 1), 0, int(bins) - 1);", "as above"),
    ("PLTPT/Path.slang", "    enum class Flags {", "    enum class Flags : uint {", "the mask 0xffff0000 needs an unsigned underlying type in C++ (Slang enums are 32-bit words)"),
    ("Materials/BSDFSample.slang", "    float   pdf;  ", "    float   pdf = 0.f;  ",
     "the closest-hit program reads bsdfSample.lobe / .pdf on paths that skipped sample() (pltpt_sample.rt.slang:87-103): a GPU compiler gives the "
     "unwritten local zeros, C++ an indeterminate value -- zero here, as the oracle defines it"),
    ("Materials/BSDFSample.slang", "    uint    lobe;  ", "    uint    lobe = 0;  ", "as above"),
    ("Materials/BSDFSample.slang", "    float   weight;", "    float   weight = 0.f;", "as above"),
    ("Geometry/GeometryHelpers.slang", "float3 iPos = asfloat(asint(pos) + (pos < 0.f ? -iOff : iOff));", "float3 iPos = asfloat(asint(pos) + select(pos < 0.f, -iOff, iOff));", "R11"),
    ("Geometry/GeometryHelpers.slang", "return abs(pos) < origin ? pos + fOff : iPos;", "return select(abs(pos) < origin, pos + fOff, iPos);", "R11"),
    ("PLTPT/LightSampling.slangh", "void getLightTypeSelectionProbabilities(out float p[3]) {", "void getLightTypeSelectionProbabilities(float p[3]) {",
     "an array parameter is already passed by reference in C++ (R7 would make it an array of references)"),
    ("PLT/PLTDiffractionGratedConductorBSDF.slang", "const int2 sign = lobe<0 ? int2(0) : int2(1);", "const int2 sign = select(lobe<0, int2(0), int2(1));", "R11"),
    ("PLT/PLTDiffractionGratedOpaqueDielectricBSDF.slang", "const int2 sign = lobe<0 ? int2(0) : int2(1);", "const int2 sign = select(lobe<0, int2(0), int2(1));", "R11"),
]


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def strip_comments_keep_lines(s):
    """Block comments and line comments removed (reference prose does not need to travel into the generated file)."""
    s = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), s, flags=re.S)
    s = re.sub(r"//[^\n]*", "", s)
    return s


def match_brace(s, i):
    """index just past the brace block that opens at s[i] == '{'"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "{":
            depth += 1
        elif s[j] == "}":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced braces")


def extract_items(s, names):
    """Top-level functions / structs with the given names (all overloads of a name in source order), emitted in the order of `names`
    (Slang resolves names regardless of declaration order, C++ needs the callee first)."""
    found = {n: [] for n in names}
    pat = re.compile(r"^(?:static\s+)?(?:inline\s+)?(?:const\s+)?[\w<>,:]+\s+(\w+)\s*(?:<[^>]*>)?\s*\([^;{]*\)\s*\{|^struct\s+(\w+)[^{;]*\{", re.M)
    pos = 0
    while True:
        m = pat.search(s, pos)
        if not m:
            break
        name = m.group(1) or m.group(2)
        end = match_brace(s, m.end() - 1)
        if name in found:
            found[name].append(s[m.start():end] + (";" if m.group(2) else ""))
        pos = end
    missing = [n for n in names if not found[n]]
    if missing:
        raise SystemExit(f"items not found in the reference source: {missing}")
    return "\n".join("\n".join(found[n]) for n in names)


def extract_members(s, names):
    """Member functions (indented, inside a struct) with the given names, all overloads, in source order."""
    out = []
    pat = re.compile(r"^[ \t]+(?:static\s+)?[\w<>,:]+\s+(\w+)\s*(?:<[^>]*>)?\s*\([^;{]*\)\s*\{", re.M)
    pos = 0
    while True:
        m = pat.search(s, pos)
        if not m:
            break
        end = match_brace(s, m.end() - 1)
        if m.group(1) in names:
            out.append(s[m.start():end])
            pos = end
        else:
            pos = m.end()
    found = {re.search(r"(\w+)\s*(?:<[^>]*>)?\s*\(", o).group(1) for o in out}
    if set(names) - found:
        raise SystemExit(f"members not found: {set(names) - found}")
    return "\n".join(out)


FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def rule_literals(s):
    return FLOAT_LIT.sub(lambda m: m.group(1) + "f", s)


def rule_defines(s):
    """MathConstants.slangh: keep the #define lines of the shader branch, as fp32 literals."""
    body = s[s.index("#else"):s.index("#endif // !HOST_CODE")]
    out = []
    for line in body.splitlines():
        m = re.match(r"#define\s+(\w+)\s+(\S+)", line)
        if m and (m.group(1).startswith("M_") or m.group(1).startswith("FLT_")):
            val = m.group(2)
            if re.fullmatch(r"[-+]?(\d+\.\d*|\.\d+)([eE][+-]?\d+)?", val):
                val += "f"
            out.append(f"#undef {m.group(1)}\n#define {m.group(1)} {val}")
    return "\n".join(out)


def rule_structs(s):
    """R3, R5, R9 (interface lists), R10 (trailing ';'): per struct, brace matched."""
    out, pos = [], 0
    pat = re.compile(r"\bstruct\s+(\w+)\s*(:[^{]*)?\{")
    while True:
        m = pat.search(s, pos)
        if not m:
            out.append(s[pos:])
            break
        name = m.group(1)
        end = match_brace(s, m.end() - 1)
        body = s[m.end():end - 1]
        had_init = "__init" in body
        body = re.sub(r"\b__init\s*\(", name + "(", body)
        if had_init:
            body = f"\n    {name}() = default;" + body
        # properties -> getters
        def prop(pm):
            pname, ptype = pm.group(1), pm.group(2)
            b0 = body_ref[0].index("{", pm.end() - 1)
            return None
        body_ref = [body]
        pp = re.compile(r"\bproperty\s+(\w+)\s*:\s*(\w+)\s*\{")
        props = []
        while True:
            pm = pp.search(body)
            if not pm:
                break
            pend = match_brace(body, pm.end() - 1)
            inner = body[pm.end():pend - 1]
            gm = re.search(r"\bget\s*\{", inner)
            gend = match_brace(inner, gm.end() - 1)
            getter = inner[gm.end():gend - 1]
            setter = ""
            sm = re.search(r"\bset\s*\{", inner)
            if sm:   # R5: `set { BODY }` (implicit parameter `newValue`) -> `void set_NAME(T newValue) { BODY }`
                send = match_brace(inner, sm.end() - 1)
                setter = f"\n    void set_{pm.group(1)}({pm.group(2)} newValue) {{{inner[sm.end():send - 1]}}}"
            body = body[:pm.start()] + f"{pm.group(2)} {pm.group(1)}() {{{getter}}}" + setter + body[pend:]
            props.append(pm.group(1))
        for p in props:  # bare uses inside the struct: not a member access of something else, not already a call, not a declaration
            body = re.sub(rf"(?<![\w.:]){p}\b(?!\s*\()(?!\s*=[^=])", p + "()", body)   # not `Enum::NAME`, not an enumerator `NAME = value`
            body = re.sub(rf"\b(\w+)\s+{p}\(\)\(\)\s*\{{", rf"\1 {p}() {{", body)
        body = rule_structs(body)  # nested structs
        tail = s[end:end + 1]
        base = " : MaterialBase" if m.group(2) and "MaterialBase" in m.group(2) else ""   # R9: interfaces go, a real base struct stays
        out.append(s[pos:m.start()] + f"struct {name}{base} {{" + body + "}" + ("" if tail == ";" else ";"))
        pos = end
    return "".join(out)


def scalar_swizzles(s):
    """R8c: `.x` applied to a variable declared as a plain float in the same function (`k.x` where `const float k`)."""
    floats = set(re.findall(r"\b(?:const\s+)?float\s+(\w+)\s*(?:=|;|,|\))", s))
    for name in floats:
        # only if no vector declaration of the same name exists in the text unit
        if re.search(rf"\b(?:float[234]|hwss_t|int2|uint2)\s+{name}\b", s):
            continue
        s = re.sub(rf"(?<![\w.]){name}\.x\b", name, s)
    return s


def per_function(s, fn):
    """applies fn to every outermost brace block that follows a ')' (function bodies), leaving the rest untouched"""
    out, pos = [], 0
    for m in re.finditer(r"\)\s*(?:const\s*)?\{", s):
        if m.start() < pos:
            continue
        b = m.end() - 1
        e = match_brace(s, b)
        # include the parameter list (walk back to the matching '(')
        depth, a = 0, m.start()
        while a >= 0:
            if s[a] == ")":
                depth += 1
            elif s[a] == "(":
                depth -= 1
                if depth == 0:
                    break
            a -= 1
        out.append(s[pos:a])
        out.append(fn(s[a:e]))
        pos = e
    out.append(s[pos:])
    return "".join(out)


def transpile(path, s, struct_names):
    for suffix, old, new, _why in PATCHES:
        if path.endswith(suffix):
            if old not in s:
                raise SystemExit(f"patch for {path} no longer matches: {old[:60]!r}")
            s = s.replace(old, new)
    s = strip_comments_keep_lines(s)
    for suffix, name in ENTRY_POINTS.items():   # R14
        if path.endswith(suffix):
            s = re.sub(r"\bvoid\s+main\s*\(", f"void {name}(", s)

    if path.endswith(".cs.slang"):
        s = re.sub(r"\s*:\s*SV_\w+", "", s)   # R14: HLSL system-value semantics of a compute entry point's parameters
    # R1
    s = re.sub(r"^\s*(__exported\s+)?import\s+[\w.]+\s*;\s*$", "", s, flags=re.M)
    s = re.sub(r'^\s*#include\s+"[^"]*"\s*$', "", s, flags=re.M)
    s = re.sub(r"^\s*#pragma once\s*$", "", s, flags=re.M)
    s = re.sub(r"^\s*(BEGIN|END)_NAMESPACE_FALCOR\s*$", "", s, flags=re.M)
    s = re.sub(r"\blet\s+", "auto ", s)   # R12: `let x = ...` -> `auto x = ...`
    s = re.sub(r"^enum\s+class\s+(\w+)\s*:", r"enum \1 :", s, flags=re.M)   # R13 (namespace-scope enums only: a nested one stays scoped)
    # R4
    s = re.sub(r"\[mutating\]\s*", "", s)
    s = re.sub(r"^\s*\[(?:unroll|ForceInline|Differentiable|BackwardDifferentiable|loop|branch|flatten|root)(?:\([^\]]*\))?\]\s*", "", s, flags=re.M)
    s = re.sub(r"\bno_diff\s+", "", s)
    s = re.sub(r"^\s*\[shader\([^\]]*\)\]\s*$", "", s, flags=re.M)   # DXR entry-point attributes
    s = re.sub(r"\bprecise\s+", "", s)          # every operation is individually rounded here anyway (-ffp-contract=off)
    s = re.sub(r"\s*\bCONST_FUNCTION\b", "", s)  # Utils/HostDeviceShared.slangh: empty in shader code
    # R2
    s = rule_literals(s)
    # R6
    s = re.sub(r"^(\s*)((?:static\s+)?[\w:<>]+\s+)(\w+)\s*<\s*(\w+)\s*:\s*\w+\s*>\s*\(", r"\1template <class \4> \2\3(", s, flags=re.M)
    # R7
    s = re.sub(r"\b(?:in)?out\s+((?:const\s+)?[\w:]+)\s+(\w+)", r"\1& \2", s)
    # R3 (this) + structs
    s = re.sub(r"\bthis\.", "this->", s)
    s = rule_structs(s)
    # R9
    for sc in STATIC_SCOPES:
        s = re.sub(rf"\b{sc}\.(?=\w)", sc + "::", s)
    s = re.sub(r"\b(\w+)\.Flags::", r"\1::Flags::", s)
    # R8
    s = re.sub(r"\)\.rgb\b", ").xyz()", s)   # `.rgb` of a call result (materialEvalEmissive, EnvMap.eval)
    if path.endswith("Material.slang"):   # R8: colour-named swizzles of the sampled texture values
        s = re.sub(r"\.rgb\b", ".xyz", s)
        s = re.sub(r"\.rg\b", ".xy", s)
        s = re.sub(r"\.a\b", ".w", s)
    s = re.sub(r"\brgb\.r\b", "rgb.x", s)
    s = re.sub(r"\brgb\.g\b", "rgb.y", s)
    s = re.sub(r"\brgb\.b\b", "rgb.z", s)
    s = re.sub(r"\.([xyzw]{2,4})\b(?!\s*\()", r".\1()", s)
    s = per_function(s, scalar_swizzles)
    if re.search(r"PLTPT/(Path\.slang|pltpt\.slangh|pltpt_sample\.rt\.slang|pltpt_solve\.rt\.slang)$", path):   # R5: RayFlags / BounceData properties with setters
        props = "MSCompatibleBounces|lastBounceWasDelta|dispersiveDeltaSegment|onlyDelta|miss|terminated"
        s = re.sub(rf"(\b[\w.]*flags)\.({props})\s*\|=\s*([^;]+);", r"\1.set_\2(\1.\2() | (\3));", s)
        s = re.sub(rf"(\b[\w.]*flags)\.({props})\s*=(?!=)\s*([^;]+);", r"\1.set_\2(\3);", s)
        s = re.sub(rf"(\b[\w.]*[fF]lags)\.({props})\b(?!\s*\()", r"\1.\2()", s)
        s = re.sub(r"(\b[\w.]*flags)\.valid\b(?!\s*\()", r"\1.valid()", s)
        s = re.sub(r"(\bdebugOutput|\])\.rgb\b", r"\1.xyz()", s)   # R8: colour-named swizzles of the debug overlay (DO_DEBUG_VIEW branch)
        s = re.sub(r"\bdebugOutput\.a\b", "debugOutput.w", s)
        s = re.sub(r"\bbounce\.wi\b(?!\s*[\(=])", "bounce.wi()", s)
        s = re.sub(r"\bbounce\.wi\s*=(?!=)\s*([^;]+);", r"bounce.set_wi(\1);", s)
        s = re.sub(r"\bbsdfProp\.(canManifoldWalk(?:Reflective|Transmissive)|isEmissive)\b(?!\s*\()", r"bsdfProp.\1()", s)
    if re.search(r"PLTPT/(Manifolds\.slangh|pltpt_solve\.rt\.slang)$", path):   # R5: ShadowRayHit.isHit, MSHitData.index read outside their structs
        s = re.sub(r"\.isHit\b(?!\s*\()", ".isHit()", s)
        s = re.sub(r"\bbsdfProp\.(canManifoldWalk(?:Reflective|Transmissive))\b(?!\s*\()", r"bsdfProp.\1()", s)
        s = re.sub(r"(\bhits\[[^\]]*\])\.index\b(?!\s*\()", r"\1.index()", s)
    for pu in PROPERTY_USES + (PROPERTY_USES_FRACTAL if re.search(r"typedef\s+FractalSurface\s+SurfaceType", s) else []):
        s = re.sub(rf"\b({pu})\b(?!\s*\()", r"\1()", s)
    # R10
    s = re.sub(r"\bstatic\s+const\s+(float|int|uint|bool)\b", r"static constexpr \1", s)
    if struct_names:
        s = re.sub(r"\bconst\s+(" + "|".join(sorted(struct_names, key=len, reverse=True)) + r")\b", r"\1", s)
    s = re.sub(r"\n{3,}", "\n\n", s)
    return s


def main():
    if not os.path.isdir(REF):
        raise SystemExit("/root/reference is not present: oracle/_ref can only be (re)generated where the reference lies")
    os.makedirs(OUT, exist_ok=True)
    texts = []
    for path, what in SOURCES:
        if path == "literal":   # a declaration the pass makes outside the transpiled files, placed between two of them
            texts.append((path, what, True))
            continue
        raw = read(path)
        if what == "defines":
            texts.append((path, rule_defines(raw), True))
            continue
        if isinstance(what, tuple) and what[0] == "members":
            _, sname, names, fields = what
            raw = f"struct {sname}\n{{\n    {fields}\n" + extract_members(raw, names) + "\n};\nstatic " + sname + " g" + sname + ";\n"
        elif what is not None:
            raw = extract_items(raw, what)
        texts.append((path, raw, False))
    struct_names = set()
    for _p, t, _d in texts:
        struct_names.update(re.findall(r"\bstruct\s+(\w+)", t))
    struct_names.update(["ShadingData", "DifferentialShadingGeometry", "BSDFProperties", "IMaterialInstance", "VertexData", "TriangleHit"])   # shim structs
    parts = ["// GENERATED by oracle/ref_transpile.py from the reference's Slang sources under /root/reference -- not part of the repository.",
             "#pragma once", '#include "../ref/slang_compat.h"', '#include "../ref/ref_shims_pre.h"', "namespace ref {"]
    for path, t, is_def in texts:
        parts.append(f"\n// ======== {path}")
        parts.append(t if is_def else transpile(path, t, struct_names))
    parts.append("}  // namespace ref")
    out = os.path.join(OUT, "ref_gen.h")
    with open(out, "w") as f:
        f.write("\n".join(parts) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
