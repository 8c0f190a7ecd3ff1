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
