// ORACLE -- TEST INFRASTRUCTURE ONLY.
// ref_shims_late.h -- included from oracle/_ref/ref_gen.h inside namespace ref, after the ten PLT materials and Path.slang: the dynamic
// dispatch Slang does through its IMaterial / IMaterialInstance existentials (MaterialFactory.slang:60-131, createDynamicObject by material
// type), written out as a switch over the ten transpiled types, and the pass's bounce buffer.  Hand-written plumbing: every BSDF, loader and
// ShadingUtils function called here is the transpiled reference text.

// IMaterialInstance (Rendering/Materials/IMaterialInstance.slang:90-150): whichever of the ten BSDFs the material set up
struct AnyMaterialInstance {
    MaterialType type = MaterialType::PLTDiffuse;
    PLTDiffuseBSDF diffuse;
    PLTConductorBSDF conductor;
    PLTDielectricBSDF dielectric;
    PLTOpaqueDielectricBSDF opaque;
    PLTThinDielectricBSDF thin;
    PLTCoatedConductorBSDF coatedC;
    PLTCoatedOpaqueDielectricBSDF coatedD;
    PLTDiffractionGratedConductorBSDF gratedC;
    PLTDiffractionGratedOpaqueDielectricBSDF gratedD;
    PLTMultiLayeredStackBSDF stack;
    template <class S> bool sample(ShadingData sd, S& sg, BSDFSample& result, const float wavelength, const uint lobeMask = (uint)LobeType::All, const float3 woHint = .0f) {
        switch (type) {case MaterialType::PLTDiffuse: return diffuse.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTConductor: return conductor.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTDielectric: return dielectric.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTOpaqueDielectric: return opaque.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTThinDielectric: return thin.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTCoatedConductor: return coatedC.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTCoatedOpaqueDielectric: return coatedD.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTDiffractionGratedConductor: return gratedC.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTDiffractionGratedOpaqueDielectric: return gratedD.sample(sd, sg, result, wavelength, lobeMask, woHint); case MaterialType::PLTMultiLayeredStack: return stack.sample(sd, sg, result, wavelength, lobeMask, woHint); default: return diffuse.sample(sd, sg, result, wavelength, lobeMask, woHint); }
    }
    void eval(ShadingData sd, const float3 toEye, const uint lobe, PLTBeam& beam) {
        switch (type) {case MaterialType::PLTDiffuse: diffuse.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTConductor: conductor.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTDielectric: dielectric.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTOpaqueDielectric: opaque.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTThinDielectric: thin.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTCoatedConductor: coatedC.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTCoatedOpaqueDielectric: coatedD.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTDiffractionGratedConductor: gratedC.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTDiffractionGratedOpaqueDielectric: gratedD.eval(sd, toEye, lobe, beam); break; case MaterialType::PLTMultiLayeredStack: stack.eval(sd, toEye, lobe, beam); break; default: diffuse.eval(sd, toEye, lobe, beam); break; }
    }
    hwss_t lambdaPdf(ShadingData sd, const float3 wo, const uint lobe, PLTBeam beam, const hwss_t wavelengths) {
        switch (type) {case MaterialType::PLTDiffuse: return diffuse.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTConductor: return conductor.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTDielectric: return dielectric.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTOpaqueDielectric: return opaque.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTThinDielectric: return thin.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTCoatedConductor: return coatedC.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTCoatedOpaqueDielectric: return coatedD.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTDiffractionGratedConductor: return gratedC.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTDiffractionGratedOpaqueDielectric: return gratedD.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; case MaterialType::PLTMultiLayeredStack: return stack.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; default: return diffuse.lambdaPdf(sd, wo, lobe, beam, wavelengths); break; }
        return hwss_t(.0f);
    }
    float pdf(ShadingData sd, const float3 toLight, const float wavelength) {
        switch (type) {case MaterialType::PLTDiffuse: return diffuse.pdf(sd, toLight, wavelength); case MaterialType::PLTConductor: return conductor.pdf(sd, toLight, wavelength); case MaterialType::PLTDielectric: return dielectric.pdf(sd, toLight, wavelength); case MaterialType::PLTOpaqueDielectric: return opaque.pdf(sd, toLight, wavelength); case MaterialType::PLTThinDielectric: return thin.pdf(sd, toLight, wavelength); case MaterialType::PLTCoatedConductor: return coatedC.pdf(sd, toLight, wavelength); case MaterialType::PLTCoatedOpaqueDielectric: return coatedD.pdf(sd, toLight, wavelength); case MaterialType::PLTDiffractionGratedConductor: return gratedC.pdf(sd, toLight, wavelength); case MaterialType::PLTDiffractionGratedOpaqueDielectric: return gratedD.pdf(sd, toLight, wavelength); case MaterialType::PLTMultiLayeredStack: return stack.pdf(sd, toLight, wavelength); default: return diffuse.pdf(sd, toLight, wavelength); }
    }
    uint getLobes(ShadingData sd) {
        switch (type) {case MaterialType::PLTDiffuse: return diffuse.getLobes(sd); case MaterialType::PLTConductor: return conductor.getLobes(sd); case MaterialType::PLTDielectric: return dielectric.getLobes(sd); case MaterialType::PLTOpaqueDielectric: return opaque.getLobes(sd); case MaterialType::PLTThinDielectric: return thin.getLobes(sd); case MaterialType::PLTCoatedConductor: return coatedC.getLobes(sd); case MaterialType::PLTCoatedOpaqueDielectric: return coatedD.getLobes(sd); case MaterialType::PLTDiffractionGratedConductor: return gratedC.getLobes(sd); case MaterialType::PLTDiffractionGratedOpaqueDielectric: return gratedD.getLobes(sd); case MaterialType::PLTMultiLayeredStack: return stack.getLobes(sd); default: return diffuse.getLobes(sd); }
    }
    float3kpdv partialDerivatives(float3pdv wi, DifferentialShadingGeometry dsg, const uint lobe, float wavelength) {
        switch (type) {case MaterialType::PLTDiffuse: return diffuse.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTConductor: return conductor.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTDielectric: return dielectric.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTOpaqueDielectric: return opaque.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTThinDielectric: return thin.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTCoatedConductor: return coatedC.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTCoatedOpaqueDielectric: return coatedD.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTDiffractionGratedConductor: return gratedC.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTDiffractionGratedOpaqueDielectric: return gratedD.partialDerivatives(wi, dsg, lobe, wavelength); case MaterialType::PLTMultiLayeredStack: return stack.partialDerivatives(wi, dsg, lobe, wavelength); default: return diffuse.partialDerivatives(wi, dsg, lobe, wavelength); }
    }
    BSDFProperties getProperties(ShadingData sd) {
        switch (type) {case MaterialType::PLTDiffuse: return diffuse.getProperties(sd); case MaterialType::PLTConductor: return conductor.getProperties(sd); case MaterialType::PLTDielectric: return dielectric.getProperties(sd); case MaterialType::PLTOpaqueDielectric: return opaque.getProperties(sd); case MaterialType::PLTThinDielectric: return thin.getProperties(sd); case MaterialType::PLTCoatedConductor: return coatedC.getProperties(sd); case MaterialType::PLTCoatedOpaqueDielectric: return coatedD.getProperties(sd); case MaterialType::PLTDiffractionGratedConductor: return gratedC.getProperties(sd); case MaterialType::PLTDiffractionGratedOpaqueDielectric: return gratedD.getProperties(sd); case MaterialType::PLTMultiLayeredStack: return stack.getProperties(sd); default: return diffuse.getProperties(sd); }
    }
};
typedef AnyMaterialInstance IMaterialInstance;

#define REF_FOR_MATERIAL(TYPE, STMT)                                                                                              \
    switch (TYPE) {                                                                                                               \
        case MaterialType::PLTDiffuse: { PLTDiffuseMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, diffuse) } break; \
        case MaterialType::PLTConductor: { PLTConductorMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, conductor) } break; \
        case MaterialType::PLTDielectric: { PLTDielectricMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, dielectric) } break; \
        case MaterialType::PLTOpaqueDielectric: { PLTOpaqueDielectricMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, opaque) } break; \
        case MaterialType::PLTThinDielectric: { PLTThinDielectricMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, thin) } break; \
        case MaterialType::PLTCoatedConductor: { PLTCoatedConductorMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, coatedC) } break; \
        case MaterialType::PLTCoatedOpaqueDielectric: { PLTCoatedOpaqueDielectricMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, coatedD) } break; \
        case MaterialType::PLTDiffractionGratedConductor: { PLTDiffractionGratedConductorMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, gratedC) } break; \
        case MaterialType::PLTDiffractionGratedOpaqueDielectric: { PLTDiffractionGratedOpaqueDielectricMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, gratedD) } break; \
        case MaterialType::PLTMultiLayeredStack: { PLTMultiLayeredStackMaterial material; std::memcpy(static_cast<void*>(&material), &blobs[materialID], sizeof material); STMT(material, stack) } break; \
        default: break;                                                                                                           \
    }

// MaterialFactory.slang:79-131, statement by statement
inline ShadingData MaterialSystemStandIn::prepareShadingData(VertexData v, uint materialID, float3 viewDir, ExplicitLodTextureSampler lod) const {
    ShadingData sd = {};
    sd.posW = v.posW;
    sd.uv = v.texC;
    sd.V = viewDir;
    sd.N = v.normalW;
    const bool validTangentSpace = computeTangentSpace(sd, v.tangentW);
    sd.tangentW = validTangentSpace ? v.tangentW : float4(sd.T, 1.f);
    sd.faceN = v.faceNormalW;
    sd.frontFacing = dot(sd.V, sd.faceN) >= 0.f;
    MaterialHeader header = blobs[materialID].header;
    sd.materialID = materialID;
#define REF_PREPARE(material, field) sd.opacity = material.evalOpacity(*this, v, lod); if (validTangentSpace) material.modifyTangentSpace(*this, lod, sd);
    REF_FOR_MATERIAL(header.getMaterialType(), REF_PREPARE)
#undef REF_PREPARE
    if (!sd.frontFacing && header.isDoubleSided()) sd.N = -sd.N;
    return sd;
}
// MaterialFactory.slang:63-67
inline AnyMaterialInstance MaterialSystemStandIn::getMaterialInstance(ShadingData sd, ExplicitLodTextureSampler lod) const {
    AnyMaterialInstance inst;
    const uint materialID = sd.materialID;
    MaterialHeader header = blobs[materialID].header;
    inst.type = header.getMaterialType();
#define REF_SETUP(material, field) inst.field = material.setupMaterialInstance(*this, sd, lod);
    REF_FOR_MATERIAL(inst.type, REF_SETUP)
#undef REF_SETUP
    return inst;
}
#undef REF_FOR_MATERIAL

// pltpt.slangh:123: the bounce buffer of the current tile
static RWStructuredBuffer<BounceData> bounceBuffer;

// ---- DXR's TraceRay for the pass's three ray types (HLSL built-in).  The intersection search itself is an engine / hardware service with no
// shader text behind it: the caller of the comparison supplies it (the oracle's ray caster over the same triangles).  What IS reference text
// is everything that runs on the result: the shader table of PLTPT.cpp:538-548 -- sample stage: miss 0 = scatterMiss, hit group 0 =
// scatterTriangleMeshClosestHit; solve stage: miss[visibility] = visibilityMiss, hit group[shadow] = shadowTriangleMeshHit, nothing else bound --
// spelt as overloads on the payload type (defined in ref_shims_post.h, after the programs).
struct RayCastResult { bool hit = false; uint instanceIndex = 0, primitiveIndex = 0; float2 barycentrics; float t = 0.f; };
static RayCastResult (*gCastClosest)(float3 origin, float3 dir, float tMin, float tMax) = nullptr;
static bool (*gCastAny)(float3 origin, float3 dir, float tMin, float tMax) = nullptr;
static float gRayTCurrent = 0.f;
static unsigned long long gWalks = 0, gProposals = 0;   // closest-hit queries of the shadow ray type (occluder search and manifold walk), for the comparison's own report
inline float RayTCurrent() { return gRayTCurrent; }
struct VisibilityRayData;
struct ShadowRayHit;
void sbtClosestHit(Path& payload, BuiltInTriangleIntersectionAttributes attribs);
void sbtMiss(Path& payload);
void sbtClosestHit(VisibilityRayData& payload, BuiltInTriangleIntersectionAttributes attribs);
void sbtMiss(VisibilityRayData& payload);
void sbtClosestHit(ShadowRayHit& payload, BuiltInTriangleIntersectionAttributes attribs);
void sbtMiss(ShadowRayHit& payload);
template <class Payload> void TraceRay(RtAccelStandIn, uint rayFlags, uint instanceInclusionMask, uint hitIdx, uint rayTypes, uint missIdx, RayDesc ray, Payload& payload) {
    if (rayFlags & RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH) {
        if (!gCastAny(ray.Origin, ray.Direction, ray.TMin, ray.TMax)) sbtMiss(payload);
        else sbtClosestHit(payload, BuiltInTriangleIntersectionAttributes());
        return;
    }
    if (hitIdx == defaultShadowRayId && rayTypes == rayTypeCount && sizeof(Payload) <= 32) { ++gWalks; if (ray.TMin > 2e-4f) ++gProposals; }
    const RayCastResult r = gCastClosest(ray.Origin, ray.Direction, ray.TMin, ray.TMax);
    if (!r.hit) { sbtMiss(payload); return; }
    gCurrentInstanceID.index = r.instanceIndex; gCurrentPrimitiveIndex = r.primitiveIndex; gRayTCurrent = r.t;
    BuiltInTriangleIntersectionAttributes attribs; attribs.barycentrics = r.barycentrics;
    sbtClosestHit(payload, attribs);
}
