// plt_material.cuh — PLT material parameters from the Falcor material blob: the uniform case (unpacked once per scene on the host)
// and the textured case (parameter slots sampled per hit on the device), plus normal mapping.
//
// Reference: every Rendering/Materials/PLT/PLT*Material.slang::setupMaterialInstance samples its BasicMaterialData slots with
// `ms.sampleTexture(data.texX, s, sd.uv, data.X, lod)` (MaterialSystem.slang:283-291) at explicit LOD 0 (pltpt_sample.rt.slang:76,
// pltpt_solve.rt.slang:53) and derives the BSDF parameters from the result; modifyTangentSpace samples texNormalMap and calls
// applyNormalMap (Scene/Material/ShadingUtils.slang:199-246); evalOpacity reads the alpha of the base colour.
//
// B200 layout choice: a material whose handles are all Uniform (the common case of a PLT scene: spectral tables carry the detail) is
// derived ONCE at pltpt_set_scene into the 128 B DevMaterial the kernels load with eight 16 B loads; only a material with at least one
// Texture handle (flag HF_Textured) takes the per-hit path below -- an out-of-line call behind one predictable branch -- which
// re-derives the same DevMaterial from the sampled slots with the same function (deriveMaterial is __host__ __device__).
#pragma once
#include "plt_bsdf.cuh"

namespace plt {

enum : uint32_t { HF_DoubleSided = 1u, HF_Textured = 2u, HF_NormalMap = 4u };  // DevMaterial::hdrFlags

struct DevTexture { const void* texels; uint32_t width, height, format, pad; };           // mip 0, linear layout, in one device blob
struct DevSampler { uint32_t magFilter, addressU, addressV, pad; float4 border; };
struct DevTextures {
    const DevTexture* textures;
    const DevSampler* samplers;
    const PltptMaterialBlob* blobs;  // the material blobs as uploaded (read by the textured path only)
    uint32_t textureCount, samplerCount;
};

PLT_HD uint32_t texHandleMode(uint32_t h) { return (h >> 29) & 3u; }       // TextureHandle.slang:54-67
PLT_HD uint32_t texHandleID(uint32_t h) { return h & 0x1fffffffu; }        // :75
PLT_HD uint32_t blobSamplerID(const PltptMaterialBlob& m) { return (m.header[1] >> 17) & 0xffu; }  // MaterialData.slang:66,122
PLT_HD uint32_t blobNormalMapType(const PltptMaterialBlob& m) { return (m.flags >> 1) & 3u; }      // BasicMaterialData.slang:106,126

// ---- surface statistics from the roughness parameters ----------------------------------------------------------------
PLT_HD float roughnessToT(float alpha) {  // FractalSurface.slang:45-53
    const float meank = k2Pi / (kMeanWavelength * 1e-3f);
    const float a2 = sqr(fminf(fmaxf(alpha, .0f), .75f));
    return fminf(5000.f, (1 - a2) / (4.f * sqr(meank) * a2));
}
PLT_HD float TToRoughness(float T) {
    const float meank = k2Pi / (kMeanWavelength * 1e-3f);
    return 1.f / sqrtf(1.f + 4.f * sqr(meank) * T);
}
PLT_HD void setFractal(DevMaterial& d, float roughness, float gamma) {
    d.surfA = roughnessToT(roughness);
    d.surfGamma = fmaxf(1.1f, gamma);
    d.surfQ = TToRoughness(d.surfA);
    d.roughnessProp = d.surfQ;
}
PLT_HD float roughnessToSigma2(float r) { return powf(fminf(fmaxf(1.f - .5f * r, .0f), 1.f), 28.f) * 2500.f; }
PLT_HD void setGaussian(DevMaterial& d, float roughness) {  // GaussianSurface.slang:52-60
    const float s2 = roughnessToSigma2(roughness), s2max = roughnessToSigma2(0.1f), s2r0 = roughnessToSigma2(.0f);
    d.surfQ = (s2r0 - s2) / (s2r0 - s2max);
    d.surfA = fminf(s2max, s2);
    d.roughnessProp = 2.f * (1.f - powf(d.surfA / 2500.f, 1.f / 28.f));
}
PLT_HD float normalFresnelDielectric(float eta) {  // evalFresnelDielectric(eta, 1)
    if (eta == 1) return 0.f;
    const float rs = (1 - eta) / (1 + eta), rp = (eta - 1) / (eta + 1);
    return .5f * (rs * rs + rp * rp);
}

// The parameter slots of BasicMaterialData as setupMaterialInstance sees them: uniform values or texture samples.
struct MaterialSlots { float4 base, spec, data1, data2; float3 tran, emissionRGB; };

// Derives the DevMaterial of one blob from its slots exactly as the *Material.slang loaders + BSDF constructors do.
// `conductorF` = evalFresnelConductor of the material's n, k tables at the mean wavelength and normal incidence (albedo guide only).
// Returns false for a material type PLTPT cannot reach.
PLT_HD bool deriveMaterial(const PltptMaterialBlob& m, const MaterialSlots& s, float conductorF, DevMaterial& d) {
    const float4 base = s.base, spec = s.spec, data1 = s.data1, data2 = s.data2;
    const float3 tran = s.tran;
    const float IoR = h2f(m.IoR);
    d.type = m.header[0] & 0xffffu;
    d.c0x = base.x; d.c0y = base.y; d.c0z = base.z;
    d.c1x = d.c1y = d.c1z = 1.f;
    d.opacity = base.w;
    d.emissionId = PLTPT_NO_SPECTRAL_PROFILE;
    d.etaN = m.iorNSpectralId; d.etaK = m.iorKSpectralId;
    d.extEta = IoR;
    d.albx = base.x; d.alby = base.y; d.albz = base.z;
    d.surfA = d.surfGamma = d.surfQ = d.roughnessProp = 0.f;
    d.f0 = d.f1 = d.f2 = d.f3 = d.f4 = d.f5 = d.f6 = d.f7 = 0.f;
    d.u0 = d.u1 = 0u;
    d.lobes = d.propFlags = 0u;
    const float meanK = k2Pi / (kMeanWavelength * 1e-3f);
    switch (d.type) {
        case PLTPT_MATERIAL_DIFFUSE:  // PLTDiffuseMaterial.slang:40-50, PLTDiffuseBSDF.slang:136-155
            d.emissionId = m.emissionSpectralId;
            d.c1x = s.emissionRGB.x; d.c1y = s.emissionRGB.y; d.c1z = s.emissionRGB.z;  // emissionRGB (texEmissive, uniform value 1)
            d.lobes = L_DiffuseReflection;
            d.propFlags = m.emissionSpectralId != PLTPT_NO_SPECTRAL_PROFILE ? P_IsEmissive : 0;
            d.roughnessProp = 1.f;
            break;
        case PLTPT_MATERIAL_CONDUCTOR:  // PLTConductorMaterial.slang:40-57, PLTConductorBSDF.slang:218-241
            setFractal(d, spec.y, spec.x);
            d.lobes = d.surfQ == 0.f ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection);
            d.propFlags = d.surfQ < kQcutoffForMS ? P_CanMSReflective : 0;
            d.albx *= conductorF; d.alby *= conductorF; d.albz *= conductorF;
            break;
        case PLTPT_MATERIAL_DIELECTRIC: {  // PLTDielectricMaterial.slang:40-57, PLTDielectricBSDF.slang:335-360
            const float B = (tran.x > 0 || tran.y > 0 || tran.z > 0) ? spec.x : 0.f;
            d.c1x = tran.x; d.c1y = tran.y; d.c1z = tran.z;
            setGaussian(d, spec.y);
            d.f0 = fmaxf(1.f, IoR - 2.897f * B); d.f1 = spec.w; d.f2 = B;
            d.opacity = .5f * base.w;
            const uint32_t dT = B != 0.f ? L_DispersiveDeltaTransmission : L_DeltaTransmission;
            d.lobes = d.surfQ == 0.f ? (L_DeltaReflection | dT) : (L_DeltaReflection | dT | L_ScatterReflection | L_ScatterTransmission);
            d.propFlags = P_IsTransmissive | (d.surfQ < kQcutoffForMS ? P_CanMSTransmissive : 0);
            const float F = normalFresnelDielectric((d.f0 + B * sqr(meanK * k1_2Pi)) / d.f1);
            d.albx *= F; d.alby *= F; d.albz *= F;
            break;
        }
        case PLTPT_MATERIAL_OPAQUE_DIELECTRIC:  // PLTOpaqueDielectricMaterial.slang:40-52
            setFractal(d, spec.y, 3.f);
            d.lobes = d.surfQ == 0.f ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection);
            d.propFlags = d.surfQ < kQcutoffForMS ? P_CanMSReflective : 0;
            break;
        case PLTPT_MATERIAL_THIN_DIELECTRIC:  // PLTThinDielectricMaterial.slang:40-54
            d.f0 = IoR; d.f1 = spec.w; d.f2 = base.x * spec.x; d.f3 = base.y; d.f4 = base.z;
            d.opacity = .5f * base.w;
            d.lobes = L_DeltaReflection | L_DeltaTransmission;
            d.propFlags = P_IsTransmissive | P_CanMSReflective | P_CanMSTransmissive;
            d.albx = d.alby = d.albz = 1.f;
            d.roughnessProp = 0.f;
            break;
        case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC:  // PLTCoatedConductorMaterial.slang:40-60
            setFractal(d, spec.y, 3.f);
            d.f0 = data1.w; d.f1 = data1.x * spec.z;
            d.f2 = IoR + (d.f0 - IoR) * (fminf(fmaxf(d.f1, 0.f), .05f) / .05f);  // coatEta(), PLTCoatedConductorBSDF.slang:65-68
            d.u0 = d.type == PLTPT_MATERIAL_COATED_CONDUCTOR ? 1u : 0u;
            d.opacity = 1.f;
            d.lobes = d.surfQ == 0.f ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection);
            d.propFlags = d.surfQ < kQcutoffForMS ? P_CanMSReflective : 0;
            break;
        case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {  // PLTDiffractionGratedConductorMaterial.slang:41-73
            setFractal(d, data2.y, 3.f);
            d.f0 = data1.z;
            d.f1 = data1.x > 0.f ? 1.f / data1.x : 0.f; d.f2 = data1.y > 0.f ? 1.f / data1.y : 0.f;
            d.f3 = data2.x * spec.x; d.f4 = spec.y;
            d.f5 = cosf(d.f0); d.f6 = sinf(d.f0);
            const uint32_t typeAndLobes = m.data1[3];  // asuint16(data.data1.w): never textured
            d.u0 = min(max(typeAndLobes & 0xFFu, 1u), 8u);
            d.u1 = typeAndLobes >> 8;
            d.lobes = (d.surfQ == 0.f ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection)) | L_DispersiveDeltaReflection;
            d.propFlags = P_CanMSReflective;
            const float a = expf(-sqr(1.f * d.f3 * meanK));
            if (d.type == PLTPT_MATERIAL_GRATED_CONDUCTOR) { d.albx *= conductorF; d.alby *= conductorF; d.albz *= conductorF; }
            d.roughnessProp = (1 - a) * d.surfQ;
            break;
        }
        case PLTPT_MATERIAL_MULTILAYERED_STACK:  // PLTMultiLayeredStackMaterial.slang:40-56
            d.f0 = data2.x; d.f1 = data2.y; d.f2 = data1.x; d.f3 = data1.y; d.u0 = m.data1[3];
            d.lobes = L_DeltaReflection;
            d.propFlags = P_CanMSReflective;
            d.roughnessProp = 0.f;
            break;
        default:
            return false;
    }
    return true;
}
// the uniform values of a blob (fp16 slots); transmission carries w = 1 and the emissive slot's uniform value is 1 (PLTDiffuseMaterial.slang:45)
PLT_HD MaterialSlots uniformSlots(const PltptMaterialBlob& m) {
    MaterialSlots s;
    s.base = mk4(h2f(m.baseColor[0]), h2f(m.baseColor[1]), h2f(m.baseColor[2]), h2f(m.baseColor[3]));
    s.spec = mk4(h2f(m.specular[0]), h2f(m.specular[1]), h2f(m.specular[2]), h2f(m.specular[3]));
    s.data1 = mk4(h2f(m.data1[0]), h2f(m.data1[1]), h2f(m.data1[2]), h2f(m.data1[3]));
    s.data2 = mk4(h2f(m.data2[0]), h2f(m.data2[1]), h2f(m.data2[2]), h2f(m.data2[3]));
    s.tran = mk3(h2f(m.transmission[0]), h2f(m.transmission[1]), h2f(m.transmission[2]));
    s.emissionRGB = mk3(1.f);
    return s;
}

#ifdef __CUDACC__
#ifndef PLT_TEXCALL
#define PLT_TEXCALL static __device__ __noinline__
#endif
// ---- texture sampling (device): D3D texel addressing, sRGB decode per texel, exact fp32 bilinear weights ---------------------
PLT_D float srgbToLinear(float c) { return c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f); }
PLT_D float4 fetchTexel(const DevTexture& t, int x, int y) {
    const size_t i = size_t(y) * t.width + size_t(x);
    switch (t.format) {
        case PLTPT_TEXTURE_FORMAT_RGBA32_FLOAT: return __ldg(reinterpret_cast<const float4*>(t.texels) + i);
        case PLTPT_TEXTURE_FORMAT_RGBA16_FLOAT: {
            const uint2 p = __ldg(reinterpret_cast<const uint2*>(t.texels) + i);
            return mk4(h2f(p.x & 0xffffu), h2f(p.x >> 16), h2f(p.y & 0xffffu), h2f(p.y >> 16));
        }
        case PLTPT_TEXTURE_FORMAT_RGBA8_UNORM: case PLTPT_TEXTURE_FORMAT_RGBA8_UNORM_SRGB: {
            const uint32_t p = __ldg(reinterpret_cast<const uint32_t*>(t.texels) + i);
            const float4 c = mk4(float(p & 0xffu) / 255.f, float((p >> 8) & 0xffu) / 255.f, float((p >> 16) & 0xffu) / 255.f, float(p >> 24) / 255.f);
            return t.format == PLTPT_TEXTURE_FORMAT_RGBA8_UNORM ? c : mk4(srgbToLinear(c.x), srgbToLinear(c.y), srgbToLinear(c.z), c.w);
        }
        default: return mk4(0.f);
    }
}
PLT_D bool addressTexel(int& i, int n, uint32_t mode) {  // Core/API/Sampler.h:59-66; false = outside (Border)
    switch (mode) {
        case PLTPT_ADDRESS_WRAP: i %= n; if (i < 0) i += n; return true;
        case PLTPT_ADDRESS_MIRROR: { int p = i % (2 * n); if (p < 0) p += 2 * n; i = p < n ? p : 2 * n - 1 - p; return true; }
        case PLTPT_ADDRESS_CLAMP: i = i < 0 ? 0 : (i >= n ? n - 1 : i); return true;
        case PLTPT_ADDRESS_MIRROR_ONCE: if (i < 0) i = -1 - i; i = i >= n ? n - 1 : i; return true;
        default: return i >= 0 && i < n;
    }
}
PLT_D float4 addressedTexel(const DevTexture& t, const DevSampler& s, int x, int y) {
    if (!addressTexel(x, int(t.width), s.addressU) || !addressTexel(y, int(t.height), s.addressV)) return s.border;
    return fetchTexel(t, x, y);
}
PLT_D float4 sampleLevel0(const DevTexture& t, const DevSampler& s, float2 uv) {
    if (s.magFilter == PLTPT_FILTER_POINT) return addressedTexel(t, s, int(floorf(uv.x * float(t.width))), int(floorf(uv.y * float(t.height))));
    const float x = uv.x * float(t.width) - .5f, y = uv.y * float(t.height) - .5f;
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int ix = int(x0), iy = int(y0);
    const float4 c00 = addressedTexel(t, s, ix, iy), c10 = addressedTexel(t, s, ix + 1, iy), c01 = addressedTexel(t, s, ix, iy + 1), c11 = addressedTexel(t, s, ix + 1, iy + 1);
    const float4 top = c00 * (1.f - fx) + c10 * fx, bottom = c01 * (1.f - fx) + c11 * fx;
    return top * (1.f - fy) + bottom * fy;
}
// ms.sampleTexture(handle, s, uv, uniformValue, ExplicitLodTextureSampler(0)) -- MaterialSystem.slang:283-291
PLT_D float4 sampleTexture(const DevTextures& tx, uint32_t handle, uint32_t samplerID, float2 uv, float4 uniformValue) {
    if (texHandleMode(handle) == 0) return uniformValue;
    if (texHandleMode(handle) != 1 || texHandleID(handle) >= tx.textureCount) return mk4(0.f);
    DevSampler s;
    if (samplerID < tx.samplerCount) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(tx.samplers + samplerID));
        s.magFilter = a.x; s.addressU = a.y; s.addressV = a.z; s.pad = 0;
        s.border = __ldg(&tx.samplers[samplerID].border);
    } else { s.magFilter = PLTPT_FILTER_LINEAR; s.addressU = s.addressV = PLTPT_ADDRESS_WRAP; s.pad = 0; s.border = mk4(0.f); }  // MaterialSystem.cpp:76-79
    DevTexture t;
    const DevTexture* tp = tx.textures + texHandleID(handle);
    t.texels = tp->texels; t.width = tp->width; t.height = tp->height; t.format = tp->format; t.pad = 0;
    return sampleLevel0(t, s, uv);
}
PLT_D PltptMaterialBlob loadBlob(const DevTextures& tx, uint32_t materialID) {
    PltptMaterialBlob b;
    const uint4* src = reinterpret_cast<const uint4*>(tx.blobs + materialID);
    uint4* dst = reinterpret_cast<uint4*>(&b);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = __ldg(src + i);
    return b;
}
// evalFresnelConductor(eta, 1) of a material's n, k tables at the mean wavelength (albedo guide of the conductor materials)
PLT_D float conductorNormalFresnel(const DevProfile* profiles, uint32_t etaN, uint32_t etaK) {
    float2 rs, rp;
    fresnel_conductor(1.f, mk2(evalSpectrum(profiles, etaN, kMeanWavelength), evalSpectrum(profiles, etaK, kMeanWavelength)), rs, rp);
    return .5f * (cnorm2(rs) + cnorm2(rp));
}
// The textured path: the material's slots sampled at `uv`, then the same derivation as on the host.  Out of line: the untextured
// materials every benchmark scene uses never execute it, and inlined into the stage kernels it would raise their register demand.
PLT_TEXCALL DevMaterial resolveTexturedMaterial(const DevTextures& tx, const DevProfile* profiles, uint32_t materialID, float2 uv, uint32_t hdrFlags) {
    DevMaterial m;
    const PltptMaterialBlob b = loadBlob(tx, materialID);
    const uint32_t smp = blobSamplerID(b);
    const MaterialSlots u = uniformSlots(b);
    MaterialSlots s;
    s.base = sampleTexture(tx, b.texBaseColor, smp, uv, u.base);
    s.spec = sampleTexture(tx, b.texSpecular, smp, uv, u.spec);
    s.data1 = sampleTexture(tx, b.texData1, smp, uv, u.data1);
    s.data2 = sampleTexture(tx, b.texData2, smp, uv, u.data2);
    const float4 t = sampleTexture(tx, b.texTransmission, smp, uv, mk4(u.tran.x, u.tran.y, u.tran.z, 1.f));
    s.tran = mk3(t.x, t.y, t.z);
    const float4 e = sampleTexture(tx, b.texEmissive, smp, uv, mk4(1.f));
    s.emissionRGB = mk3(e.x, e.y, e.z);
    const uint32_t type = b.header[0] & 0xffffu;
    const bool cond = type == PLTPT_MATERIAL_CONDUCTOR || type == PLTPT_MATERIAL_GRATED_CONDUCTOR;
    deriveMaterial(b, s, cond ? conductorNormalFresnel(profiles, b.iorNSpectralId, b.iorKSpectralId) : 0.f, m);
    m.hdrFlags = hdrFlags;
    return m;
}
// material.modifyTangentSpace + applyNormalMap (ShadingUtils.slang:199-246); ShadingFrame(n, tangentW) = normalize + the unchecked
// orthonormalize (:61-65, :112-116).  `T`, `B`, `N`: the orthonormalised frame of the interpolated normal and tangent.
struct Frame3 { float3 T, B, N; };
PLT_TEXCALL Frame3 applyNormalMapTextured(const DevTextures& tx, uint32_t materialID, float2 uv, float4 tangentW, Frame3 f) {
    const PltptMaterialBlob b = loadBlob(tx, materialID);
    const float4 enc = sampleTexture(tx, b.texNormalMap, blobSamplerID(b), uv, mk4(0.f));
    const uint32_t type = blobNormalMapType(b);
    float3 mapN;
    if (type == 1) mapN = mk3(enc.x * 2.f - 1.f, enc.y * 2.f - 1.f, enc.z * 2.f - 1.f);  // NormalMapType::RGB, rgbToNormal
    else if (type == 2) mapN = mk3(enc.x * 2.f - 1.f, enc.y * 2.f - 1.f, sqrtf(1.f - fminf(fmaxf(enc.x * enc.x + enc.y * enc.y, 0.f), 1.f)));  // RG, rgToNormal
    else return f;
    f.N = normalize(f.T * mapN.x + f.B * mapN.y + f.N * mapN.z);
    const float3 tw = mk3(tangentW.x, tangentW.y, tangentW.z);
    f.T = normalize(tw - f.N * dot(tw, f.N));
    f.B = cross(f.N, f.T) * tangentW.w;
    return f;
}
#endif  // __CUDACC__

}  // namespace plt
