// ORACLE -- TEST INFRASTRUCTURE ONLY.
// ref_shims_mid.h -- engine plumbing that needs the transpiled BasicMaterialData / MaterialHeader / TextureHandle (included from
// oracle/_ref/ref_gen.h, inside namespace ref, right after those): the material system as the PLT material loaders and
// materialEvalEmissive use it, the base of the material structs, BSDFProperties.  Hand-written; every item cites what it stands in for.

// Rendering/Materials/IMaterialInstance.slang:53-88 (the reference's `property` accessors are not used by the BSDF code)
struct BSDFProperties {
    uint emissionSpectralId = BasicMaterialData::no_spectral_profile;
    float3 emissionRGB = .0f;
    float3 albedo = .0f;
    float roughness = .0f;
    uint flags = 0;
    enum Flags : uint { IsTransmissive = 0x1, IsEmissive = 0x2, CanManifoldWalkReflective = 0x4, CanManifoldWalkTransmissive = 0x8 };
    bool isEmissive() { return (flags & (uint)Flags::IsEmissive) != 0; }
    bool canManifoldWalkReflective() { return (flags & (uint)Flags::CanManifoldWalkReflective) != 0; }     // :70-88 (property getters)
    bool canManifoldWalkTransmissive() { return (flags & (uint)Flags::CanManifoldWalkTransmissive) != 0; }
};

// Scene/Material/MaterialSystem.slang:120-291: what the loaders and materialEvalEmissive ask of the material system.  Material blobs are the
// 128-byte MaterialDataBlob records (header + payload reinterpreted as BasicMaterialData, MaterialSystem.slang:150-160).  A Uniform texture
// handle makes sampleTexture return the `uniformValue` argument (:283-291); a Texture handle is a texture-unit service: the value comes
// from `textureFn`, supplied by the caller of the comparison.
struct MaterialSystemStandIn {
    const MaterialDataBlob* blobs = nullptr;
    float4 (*textureFn)(uint textureID, uint samplerID, float2 uv) = nullptr;
    MaterialHeader getMaterialHeader(uint materialID) const { return blobs[materialID].header; }
    bool isBasicMaterial(uint materialID) const { return getMaterialHeader(materialID).isBasicMaterial(); }
    BasicMaterialData getBasicMaterialData(uint materialID) const {
        static_assert(sizeof(BasicMaterialData) <= sizeof(MaterialPayload), "BasicMaterialData must fit the payload");
        BasicMaterialData d;
        std::memcpy(static_cast<void*>(&d), &blobs[materialID].payload, sizeof d);
        return d;
    }
    bool isEmissiveTextured(uint materialID) const { return getBasicMaterialData(materialID).texEmissive.getMode() == TextureHandle::Mode::Texture; }   // MaterialSystem.slang: the emissive slot holds a texture
    SamplerState getTextureSampler(uint samplerID) const { SamplerState s; s.id = samplerID; return s; }
    SamplerState getDefaultTextureSampler(uint materialID) const { return getTextureSampler(getMaterialHeader(materialID).getDefaultTextureSamplerID()); }
    // MaterialFactory.slang:60-131: dynamic dispatch on the material type -- defined in ref_shims_late.h, once the material structs are complete
    struct ShadingData prepareShadingData(struct VertexData v, uint materialID, float3 viewDir, ExplicitLodTextureSampler lod) const;
    struct AnyMaterialInstance getMaterialInstance(struct ShadingData sd, ExplicitLodTextureSampler lod) const;
    float4 sampleTexture(TextureHandle handle, SamplerState s, float2 uv, float lod) const { return textureFn ? textureFn(handle.getTextureID(), s.id, uv) : float4(0.f); }   // MaterialSystem.slang: explicit-mip overload
    float4 sampleTexture(TextureHandle handle, SamplerState s, float2 uv, float4 uniformValue, ExplicitLodTextureSampler) const {
        if (handle.getMode() == TextureHandle::Mode::Texture && textureFn) return textureFn(handle.getTextureID(), s.id, uv);
        return uniformValue;
    }
};
typedef MaterialSystemStandIn MaterialSystem;

// Rendering/Materials/IMaterial.slang:91-94: the header every material struct starts with
struct MaterialBase { MaterialHeader header; };
// Scene/SceneTypes.slang:311-320 (evalOpacity reads texC)
struct VertexData { float3 posW, normalW; float4 tangentW; float2 texC; float3 faceNormalW; float curveRadius = 0.f; };
// Scene/HitInfo.slang:66-123: the three fields of a triangle hit and its barycentric weights (the packed form is the pass's own business)
struct TriangleHit {
    GeometryInstanceID instanceID;
    uint primitiveIndex = 0;
    float2 barycentrics;
    float3 getBarycentricWeights() { return float3(1.f - barycentrics.x - barycentrics.y, barycentrics.x, barycentrics.y); }
    // the packed form depends on the scene's bit allocation (HitInfo.slang:292-312, set by defines): supplied by the caller of the comparison
    static inline PackedHitInfo (*packFn)(const TriangleHit&) = nullptr;
    static inline void (*unpackFn)(PackedHitInfo, TriangleHit&) = nullptr;
    PackedHitInfo pack() { return packFn(*this); }
    TriangleHit() = default;
    TriangleHit(PackedHitInfo packed) { unpackFn(packed, *this); }
};
// DXR system values of the hit being shaded (Scene/Raytracing.slang getGeometryInstanceID, HLSL PrimitiveIndex)
static GeometryInstanceID gCurrentInstanceID;
static uint gCurrentPrimitiveIndex = 0;
inline GeometryInstanceID getGeometryInstanceID() { return gCurrentInstanceID; }
inline uint PrimitiveIndex() { return gCurrentPrimitiveIndex; }
