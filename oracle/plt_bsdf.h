// ORACLE — TEST INFRASTRUCTURE ONLY (see plt_math.h header).  Pinned to the reference's shader text: surface statistics,
// gratings, birefringence, eval / sample / pdf / lambdaPdf of the ten BSDFs, the ten material loaders on shared 128-byte records, texture-handle
// decoding (oracle/ref_compare.cpp, tests/test_oracle_ref.py).  The texture FILTER has no source in the reference (hardware): defined here.
// plt_bsdf.h — surface statistics, the ten PLT BSDFs and material-blob unpacking.
#pragma once
#include "plt_beam.h"

namespace orc {

enum Lobe : uint32_t {
    L_DiffuseReflection = 0x01, L_ScatterReflection = 0x02, L_DeltaReflection = 0x04, L_DispersiveDeltaReflection = 0x08,
    L_DiffuseTransmission = 0x10, L_ScatterTransmission = 0x20, L_DeltaTransmission = 0x40, L_DispersiveDeltaTransmission = 0x80,
    L_DispersiveDelta = 0x88, L_Delta = 0xCC, L_NonDelta = 0x33, L_Reflection = 0x0f, L_Transmission = 0xf0,
    L_NonDeltaReflection = 0x03, L_NonDeltaTransmission = 0x30, L_All = 0xffff, L_MaterialSpecific = 0xff00,
};
static constexpr int LobeTypeMaterialSpecificOffset = 8;
static constexpr float QcutoffForMS = .15f;  // IMaterialInstance.slang:44
// exp() of the delta-lobe energies `alpha`: the BSDFs go on to form 1 - alpha, which cancels catastrophically for
// near-smooth surfaces, so a 1-ulp difference between two exp implementations becomes an O(1) relative difference.
// The oracle therefore DEFINES this exp as the correctly rounded fp32 value (evaluated in double); the CUDA path does the same.
inline float exp_cr(float x) { return float(std::exp(double(x))); }

// IMaterialInstance.slang:50-95
struct BSDFProperties {
    uint32_t emissionSpectralId = PLTPT_NO_SPECTRAL_PROFILE;
    float3 emissionRGB, albedo;
    float roughness = 0;
    uint32_t flags = 0;
    enum { IsTransmissive = 1, IsEmissive = 2, CanManifoldWalkReflective = 4, CanManifoldWalkTransmissive = 8 };
    bool isEmissive() const { return flags & IsEmissive; }
    bool canManifoldWalkReflective() const { return flags & CanManifoldWalkReflective; }
    bool canManifoldWalkTransmissive() const { return flags & CanManifoldWalkTransmissive; }
};
// BSDFSample.slang:33-44. Zero-initialised here; the reference leaves it uninitialised when the path ends
// before sampling (pltpt_sample.rt.slang:95-106) — the oracle defines that case as lobe = 0.
struct BSDFSample {
    float3 wo;
    float pdf = 0;
    uint32_t lobe = 0;
    float weight = 0;
    bool isLobe(uint32_t type) const { return (lobe & type) != 0; }
};

// Scene/ShadingData.slang:56-130 (fields the PLT path reads)
struct ShadingData {
    float3 posW, V, N, T, B, faceN;
    float2 uv;
    float4 tangentW;
    bool frontFacing = true;
    uint32_t materialID = 0;
    float opacity = 1.f;
    uint32_t mtlHeader[2] = {0, 0};
    bool isDoubleSided() const { return mtlHeader[0] & (1u << 28); }
    float3 fromLocal(float3 v) const { return T * v.x + B * v.y + N * v.z; }
    float3 toLocal(float3 v) const { return float3(dot(v, T), dot(v, B), dot(v, N)); }
    float3 computeNewRayOrigin(bool viewside = true) const { return computeRayOrigin(posW, (frontFacing == viewside) ? faceN : -faceN); }
};

// Rendering/Utils/DifferentialShadingGeometry.slang:6-18
struct float3pdv {
    float3 v, ddu, ddv;
    void flip() { v = -v; ddu = -ddu; ddv = -ddv; }
};
struct float3kpdv {
    float3 v, ddu, ddv, ddl;
};
struct DifferentialShadingGeometry {
    float3pdv p, n, s, t;
    float2 uv;
    float3 ngeo;
    float3pdv pdw(float3 dst) const {
        float3pdv w;
        w.v = dst - p.v;
        const float ilo = 1.f / length(w.v);
        w.v *= ilo;
        w.ddu = -ilo * (p.ddu - w.v * dot(w.v, p.ddu));
        w.ddv = -ilo * (p.ddv - w.v * dot(w.v, p.ddv));
        return w;
    }
    float3 fromLocal(float3 v) const { return s.v * v.x + t.v * v.y + n.v * v.z; }
};
// DifferentialShadingGeometry.slang:181-222
inline float3kpdv reflectDerivatives(float3pdv wi, float3pdv n) {
    const float wn = dot(wi.v, n.v);
    const float dwdun = dot(wi.ddu, n.v), dwdvn = dot(wi.ddv, n.v);
    const float dnduw = dot(n.ddu, wi.v), dndvw = dot(n.ddv, wi.v);
    float3kpdv ret;
    ret.v = 2.f * wn * n.v - wi.v;
    ret.ddu = 2.f * ((dwdun + dnduw) * n.v + wn * n.ddu) - wi.ddu;
    ret.ddv = 2.f * ((dwdvn + dndvw) * n.v + wn * n.ddv) - wi.ddv;
    ret.ddl = float3(.0f);
    return ret;
}
inline float3kpdv refractDerivatives(float3pdv wi, float3pdv n, float eta, float dEta_dLambda = .0f, float dInvEta_dLamda = .0f) {
    const float wn = dot(wi.v, n.v);
    eta = wn < .0f ? eta : 1.f / eta;
    const float dEdl = wn < .0f ? dEta_dLambda : dInvEta_dLamda;
    if (wn < .0f) n.flip();
    const float eta2 = sqr(eta);
    const float dwdun = dot(wi.ddu, n.v), dwdvn = dot(wi.ddv, n.v);
    const float dnduw = dot(n.ddu, wi.v), dndvw = dot(n.ddv, wi.v);
    const float root = std::sqrt(hlsl_max(.0f, 1.f - eta2 * (1.0f - sqr(wn))));
    if (root == .0f) return float3kpdv();
    const float dRoot_dLambda = -(1.0f - sqr(wn)) * eta * dEdl / root;
    const float a = eta - eta2 * wn / root;
    float3kpdv ret;
    ret.v = -eta * (wi.v - n.v * wn) - n.v * root;
    ret.ddu = a * (dnduw + dwdun) * n.v + (eta * wn - root) * n.ddu - eta * wi.ddu;
    ret.ddv = a * (dndvw + dwdvn) * n.v + (eta * wn - root) * n.ddv - eta * wi.ddv;
    ret.ddl = -dEdl * (wi.v - n.v * wn) - n.v * dRoot_dLambda;
    return ret;
}

// ---------------------------------------------------------------------------------------------
// FractalSurface.slang:38-135 — K-correlation ("Gaussian-fractal") PSD
// ---------------------------------------------------------------------------------------------
static const float kMeanK = k2Pi / (SpectrumUtils::meanWavelength * 1e-3f);
struct FractalSurface {
    float T = 0, gamma = 3.f;
    static float roughnessToT(float alpha) {
        const float maxAlpha = .75f, maxT = 5000.f;
        const float alpha2 = sqr(clampf(alpha, .0f, maxAlpha));
        return hlsl_min(maxT, (1 - alpha2) / (4.f * sqr(kMeanK) * alpha2));
    }
    FractalSurface() = default;
    explicit FractalSurface(float PSDroughness, float g = 3.f) : T(roughnessToT(PSDroughness)), gamma(hlsl_max(1.1f, g)) {}
    float roughness() const { return 1.f / std::sqrt(1.f + 4.f * sqr(kMeanK) * T); }
    float q() const { return roughness() * 1.f; }
    float alpha(float3 wi, float3 wo, float k) const { return exp_cr(-sqr((std::fabs(wi.z) + std::fabs(wo.z)) * q() * k)); }
    hwss_t alpha(float3 wi, float3 wo, hwss_t k) const { hwss_t r; for (int j = 0; j < 4; ++j) r[j] = alpha(wi, wo, k[j]); return r; }
    float alpha(float3 wi, float k) const { return alpha(wi, wi, k); }
    bool isDeltaOnly() const { return q() == .0f; }
    float pGF(float2 z) const {
        const float f = 1.f / std::pow(1 + T * dot(z, z), (gamma + 1.f) / 2.f);
        return k1_2Pi * (gamma - 1) * T * f;
    }
    // Holzschuch & Pacanowski 2017 sampling (FractalSurface.slang:97-124); 4 random numbers
    float3 sample(float3 wi, SampleGenerator& sg, float k, float& pdf) const {
        const float s = std::sqrt(hlsl_max(.0f, 1 - sqr(wi.z)));
        const float phi_i = s > 0 ? std::atan2(wi.y, wi.x) : .0f;
        const float sqrtT = std::sqrt(T);
        const float2 u2d1 = sampleNext2D(sg);
        const float2 u2d2 = sampleNext2D(sg);
        (void)u2d2;
        const float M = 1.f - std::pow(1 + sqr(k) * T * sqr(1 + s), -(gamma - 1) / 2.f);
        const float f = std::sqrt(std::pow(1 - M * u2d1.x, -2.f / (gamma - 1)) - 1.f) / sqrtT;
        const float phi_max = (f == .0f || s == .0f) ? kPi : std::acos(clampf((sqr(f / k) + sqr(s) - 1) / (2.f * (f / k) * s), -1.f, 1.f));
        const float phi_f = phi_i + (2.f * u2d1.y - 1) * phi_max;
        const float2 vf = f * float2(std::cos(phi_f), std::sin(phi_f));
        const float2 zeta = vf;
        const float2 wo = zeta / k - float2(wi.x, wi.y);
        const float z = std::sqrt(hlsl_max(.0f, 1 - dot(wo, wo)));
        pdf = kPi / phi_max * z * pGF(zeta);
        return float3(wo, signf(wi.z) * z);
    }
    float pdf(float3 wi, float3 wo, float k) const { return pGF(k * (wi + wo).xy()); }
    float eval(float k, float3 h) const { return pGF(k * h.xy()); }
};

// ---------------------------------------------------------------------------------------------
// GaussianSurface.slang:37-110
// ---------------------------------------------------------------------------------------------
struct GaussianSurface {
    static constexpr float roughnessScaleToSigma2 = 2500.f;
    float sigma2 = 0, q = 0;
    static float roughnessToSigma2(float r) { return std::pow(clampf(1.f - .5f * r, .0f, 1.f), 28.f) * roughnessScaleToSigma2; }
    float roughness() const { return 2.f * (1.f - std::pow(sigma2 / roughnessScaleToSigma2, 1.f / 28.f)); }
    GaussianSurface() = default;
    explicit GaussianSurface(float PSDroughness) {
        const float s2 = roughnessToSigma2(PSDroughness);
        const float s2max = roughnessToSigma2(0.1f);
        const float s2r0 = roughnessToSigma2(.0f);
        q = lerpf(.0f, 1.f, (s2r0 - s2) / (s2r0 - s2max));
        sigma2 = hlsl_min(s2max, s2);
    }
    bool isDeltaOnly() const { return q == .0f; }
    float alpha(float3 wi, float3 wo, float k) const {
        const float a = sqr((std::fabs(wi.z) + std::fabs(wo.z)) * q * k);
        return isDeltaOnly() ? 1.f : exp_cr(-a);
    }
    hwss_t alpha(float3 wi, float3 wo, hwss_t k) const { hwss_t r; for (int j = 0; j < 4; ++j) r[j] = alpha(wi, wo, k[j]); return r; }
    float alpha(float3 wi, float k) const { return alpha(wi, wi, k); }
    float3 sample(float3 wi, SampleGenerator& sg, float k, float& pdf) const {
        const float w = 1.f / sigma2;
        const float s2 = w / sqr(k);
        const float2 mean = -wi.xy();
        const float2 wo = sample_boxmueller_truncated(sampleNext2D(sg), mean, s2, pdf);
        const float z = std::sqrt(hlsl_max(0.f, 1.f - dot(wo, wo)));
        return float3(wo, signf(wi.z) * z);
    }
    float pdf(float3 wi, float3 wo, float k) const {
        const float w = 1.f / sigma2;
        const float s2 = w / sqr(k);
        return boxmueller_truncated_pdf(wo.xy(), -wi.xy(), s2);
    }
    // coherence-aware evaluation (GaussianSurface.slang:97-109)
    float eval(float k, float2x2 invTheta, const float2x2& Q, float3 h) const {
        invTheta = mul(Q, mul(invTheta, transpose(Q)));
        const float2x2 S = ident2() / sigma2 + invTheta;
        float d = determinant(S);
        const float2x2 invS = float2x2(S.m[1][1], -S.m[0][1], -S.m[1][0], S.m[0][0]) / d;
        return std::exp(-.5f * sqr(k) * dot(h.xy(), mul(invS, h.xy()))) * (sqr(k) * k1_2Pi / std::sqrt(d));
    }
};

// Interference.slang:36-57
inline float phaseShift(float eta, float dist, float k, float2 r) { return std::sin(dist * k * eta + carg(r)); }
inline hwss_t phaseShift(float eta, float dist, hwss_t ks) { return vsin(dist * ks * eta); }
inline float4 interfere(float4 SP1, float4 SP2, float phasor_s, float phasor_p, float coh) {
    const float S1 = (SP1.x + SP1.y) / 2.f, P1 = (SP1.x - SP1.y) / 2.f;
    const float S2 = (SP2.x + SP2.y) / 2.f, P2 = (SP2.x - SP2.y) / 2.f;
    const float S = S1 + S2 + 2 * std::sqrt(hlsl_max(.0f, S1 * S2)) * phasor_s * coh;
    const float P = P1 + P2 + 2 * std::sqrt(hlsl_max(.0f, P1 * P2)) * phasor_p * coh;
    const float n = SP1.x + SP2.x > 1e-6f ? (S + P) / (SP1.x + SP2.x) : .0f;
    return float4(S + P, S - P, (SP1.z + SP2.z) * n, (SP1.w + SP2.w) * n);
}

inline float3 reflectZ(float3 wi) { return float3(-wi.x, -wi.y, wi.z); }
inline float2x2 scatterCoherenceTransform(float3 woLocal, uint32_t lobe) {
    const bool isDirect = (lobe & L_Delta) != 0;
    const float cosTheta = std::sqrt(hlsl_max(1e-2f, std::fabs(woLocal.z)));
    return isDirect ? ident2() : float2x2(cosTheta, .0f, .0f, 1.f / cosTheta);
}
inline hwss_t waveNumbers(hwss_t wavelengths) { return k2Pi / (1e-3f * wavelengths); }

// ---------------------------------------------------------------------------------------------
// Birefringence.slang:33-198
// ---------------------------------------------------------------------------------------------
namespace Birefringence {
inline float qo(float eo, float K) { const float k = 1; return std::sqrt(sqr(eo) * sqr(k) - sqr(K)); }
inline float2 qe(float eo, float ee, float3 A, float K) {
    const float e = sqr(ee) - sqr(eo);
    const float k = 1, a = A.x, b = A.y, c = A.z;
    const float G = sqr(b) * e + sqr(eo);
    const float root = std::sqrt(hlsl_max(0.f, sqr(K) * (sqr(a) * e - sqr(ee)) + sqr(k) * sqr(ee) * G));
    return float2((eo * root + c * b * K * e) / G, (-eo * root + c * b * K * e) / G);
}
inline float3 E_e(float eo, float3 A, float K, float q) {
    const float k = 1, a = A.x, b = A.y, c = A.z;
    return normalize(float3(a * sqr(k) * sqr(eo), K * c * q + b * (sqr(eo) * sqr(k) - sqr(q)), K * b * q + c * (sqr(eo) * sqr(k) - sqr(K))));
}
inline float3 E_ep(float eo, float ee, float3 A, float K) { return E_e(eo, A, K, qe(eo, ee, A, K).x); }
inline float3 E_em(float eo, float ee, float3 A, float K) { return E_e(eo, A, K, qe(eo, ee, A, K).y); }
inline float3 E_o(float3 A, float K, float q) { return normalize(float3(-A.y * K - A.z * q, A.x * K, A.x * q)); }
inline float3 E_op(float eo, float3 A, float K) { return E_o(A, K, qo(eo, K)); }
inline float3 E_om(float eo, float3 A, float K) { return E_o(A, K, -qo(eo, K)); }
inline float3 Poynting(float q, float3 E, float K) {
    return normalize(float3(E.x * (K * E.z - q * E.y), K * E.z * E.y + q * (sqr(E.x) + sqr(E.z)), -q * E.z * E.y - K * (sqr(E.x) + sqr(E.y))));
}
inline void fresnel_iso_aniso(float cos_theta, float sin_theta, float ei, float eo, float ee, float3 A,
                              float& rss, float& rsp, float& tso, float& tse, float& rps, float& rpp, float& tpo, float& tpe) {
    const float k = 1;
    const float K = k * ei * sin_theta;
    const float qi = -qo(ei, K);
    const float3 E2om = E_om(eo, A, K), E2em = E_em(eo, ee, A, K);
    const float q2om = -qo(eo, K);
    const float q2em = qe(eo, ee, A, K).y;
    const float A_1 = -E2em.x * E2om.y * K * cos_theta;
    const float A_2 = -E2em.y * E2om.x * K * cos_theta;
    const float C_1 = qi * cos_theta - K * sin_theta;
    const float C_2 = E2em.x * E2om.z * (C_1 + q2om * cos_theta) - A_1;
    const float C_3 = E2em.z * E2om.x * (C_1 + q2em * cos_theta) - A_2;
    const float Ns = C_2 * (qi + q2em) - C_3 * (qi + q2om);
    rps = (2 * E2em.x * E2om.x * cos_theta * C_1 * (q2om - q2em)) / Ns;
    rpp = (-E2em.x * E2om.z * (C_1 - q2om * cos_theta) * (qi + q2em) + E2em.z * E2om.x * (C_1 - q2em * cos_theta) * (qi + q2om) - A_1 * (q2em + qi) + A_2 * (q2om + qi)) / Ns;
    tpo = -(2 * E2em.x * cos_theta * (q2em + qi) * C_1) / Ns;
    tpe = (2 * E2om.x * cos_theta * (q2om + qi) * C_1) / Ns;
    rss = (C_2 * (qi - q2em) + C_3 * (q2om - qi)) / Ns;
    rsp = (2 * qi * (E2em.z * E2om.y * K - E2em.y * E2om.z * K + E2em.z * E2om.z * (q2om - q2em))) / Ns;
    tso = -(2 * qi * (E2em.y * K * cos_theta + E2em.z * C_1 + E2em.z * q2em * cos_theta)) / Ns;
    tse = (2 * qi * (E2om.y * K * cos_theta + E2om.z * C_1 + E2om.z * q2om * cos_theta)) / Ns;
}
inline void fresnel_aniso_iso(float cos_phi, float sin_phi, float ei, float eo, float ee, float3 A,
                              float& roo, float& roe, float& tos, float& top, float& reo, float& ree, float& tes, float& tep) {
    const float eo2 = ei;
    const float k = 1;
    const float K = k * ei * sin_phi;
    const float q2om = -qo(eo2, K);
    const float qop = qo(eo, K);
    const float2 _qe = qe(eo, ee, A, K);
    const float qep = _qe.x, qem = _qe.y;
    const float3 Eop = E_op(eo, A, K), Eep = E_ep(eo, ee, A, K), Eom = E_om(eo, A, K), Eem = E_em(eo, ee, A, K);
    const float B_1 = -Eep.x * K * cos_phi * (q2om - qep);
    const float B_2 = -Eom.x * K * cos_phi * (q2om + qop);
    const float B_3 = -Eop.x * K * cos_phi * (qop - q2om);
    const float D_1 = K * sin_phi - q2om * cos_phi;
    const float D_2 = Eep.z * (D_1 + qep * cos_phi);
    const float D_3 = Eem.z * (D_1 + qem * cos_phi);
    const float D_4 = Eop.z * (D_1 + qop * cos_phi);
    const float N2 = (Eep.x * (qep - q2om) * D_4 + Eop.x * (q2om - qop) * D_2 + Eop.y * B_1 + Eep.y * B_3);
    roo = -(Eep.x * Eom.z * (q2om - qep) * (qop * cos_phi - D_1) + Eom.x * (q2om + qop) * D_2 + Eom.y * B_1 - Eep.y * B_2) / N2;
    roe = -(-Eom.x * Eop.z * (qop * (qop * cos_phi + K * sin_phi) + q2om * D_1) + Eom.z * Eop.x * (qop * (qop * cos_phi - K * sin_phi) + q2om * D_1) + Eop.y * B_2 + Eom.y * B_3) / N2;
    tos = (Eep.x * qep * roe - Eom.x * qop + Eop.x * qop * roo) / q2om;
    top = -(Eom.z * qop - Eom.y * K - (Eep.y * K + Eep.z * qep) * roe - (Eop.y * K + Eop.z * qop) * roo) / D_1;
    reo = (Eem.x * (qem - q2om) * (D_2 + Eep.y * K * cos_phi) + Eep.x * (q2om - qep) * (D_3 + Eem.y * K * cos_phi)) / N2;
    ree = -(Eem.x * (qem - q2om) * D_4 - Eem.x * Eop.y * K * cos_phi * (q2om - qem) + Eem.y * B_3 + Eop.x * (q2om - qop) * D_3) / N2;
    tes = (Eem.x * qem + Eep.x * qep * ree + Eop.x * qop * reo) / q2om;
    tep = -(Eem.z + Eep.z * ree + Eop.z * reo) / cos_phi;
}
inline void vectors_in_slab(float3 I, float ei, float eo, float ee, float3 A, float3& Io, float3& Io2, float3& Ie, float3& Ie2, float& e_eff, float& e_eff2) {
    const float K = ei * I.z;
    const float qom = -qo(eo, K);
    const float qep = qe(eo, ee, A, K).x, qem = qe(eo, ee, A, K).y;
    const float3 Eom = E_om(eo, A, K), Eep = E_ep(eo, ee, A, K), Eem = E_em(eo, ee, A, K);
    Io = Poynting(qom, Eom, K);
    Io2 = Io;
    Io2.y *= -1.f;
    Ie = Poynting(qem, Eem, K);
    Ie2 = Poynting(qep, Eep, K);
    const float3 We = normalize(float3(0, qem, -K)), We2 = normalize(float3(0, qep, -K));
    e_eff = K / std::sqrt(hlsl_max(.0f, 1.f - sqr(We.y)));
    e_eff2 = K / std::sqrt(hlsl_max(.0f, 1.f - sqr(We2.y)));
}
}  // namespace Birefringence

// ---------------------------------------------------------------------------------------------
// DiffractionGrating.slang:40-266
// ---------------------------------------------------------------------------------------------
struct int2 { int x = 0, y = 0; };
struct DiffractionGrating {
    float2 gratingX, invPitch;
    float q = 0;
    uint32_t lobeCount = 5, type = 0;
    float intensityMultiplier = 1.f;
    enum { sinusoidal = 0, rectangular = 1, linear = 2, radial = 0x10, TypeMask = 0xF };

    DiffractionGrating() = default;
    DiffractionGrating(float gratingDir, const ShadingData& sd, float2 invPitch_, float q_, uint32_t lobes, uint32_t type_, float intensityMult) {
        type = type_;
        gratingX = float2(std::cos(gratingDir), std::sin(gratingDir));
        if ((type & radial) != 0) {
            const float2 rad = normalize(sd.uv - float2(.5f, .5f));
            gratingX = mul(float2x2(std::cos(gratingDir), std::sin(gratingDir), -std::sin(gratingDir), std::cos(gratingDir)), float2(rad.x, -rad.y));
        }
        invPitch = invPitch_;
        q = q_;
        lobeCount = lobes;
        intensityMultiplier = intensityMult;
    }
    bool is1DGrating() const { return invPitch.y == .0f; }
    float alpha(float3 wi, float k) const { return exp_cr(-sqr(wi.z * q * k)); }
    hwss_t alpha(float3 wi, hwss_t k) const { hwss_t r; for (int j = 0; j < 4; ++j) r[j] = alpha(wi, k[j]); return r; }
    ShadingData localGratingShadingFrame(const ShadingData& sd) const {
        ShadingData g;  // only T,B,N are set, as in the reference (DiffractionGrating.slang:84-91)
        g.T = sd.fromLocal(float3(gratingX, .0f));
        g.N = sd.N;
        g.B = cross(g.N, g.T);
        return g;
    }
    int2 closestLobeHint(float3 wiLocal, float3 woHint, float alpha, float lambda) const {
        int2 lobe;
        if (alpha == 1.f) return lobe;
        const float2 li2 = sqr(wiLocal.xy()) + float2(sqr(wiLocal.z));
        const float2 lo2 = sqr(woHint.xy()) + float2(sqr(woHint.z));
        float2 sini, sino, l;
        for (int i = 0; i < 2; ++i) {
            sini[i] = li2[i] > FLT_EPS ? wiLocal[i] / std::sqrt(li2[i]) : .0f;
            sino[i] = lo2[i] > FLT_EPS ? -woHint[i] / std::sqrt(lo2[i]) : .0f;
            l[i] = (sini[i] - sino[i]) / (invPitch[i] * lambda);
        }
        int lo[2];
        for (int i = 0; i < 2; ++i) {
            lo[i] = hlsl_min(hlsl_max((int)(l[i] + signf(l[i]) * .5f), -int(lobeCount)), int(lobeCount));
            lo[i] = (alpha == .0f && lo[i] == 0) ? int(signf(l[i])) : lo[i];
        }
        lobe.x = lo[0];
        lobe.y = is1DGrating() ? 0 : lo[1];
        return lobe;
    }
    float sincf(float x) const { return std::sin(x) / x; }
    float lobeIntensity(int2 lobe, float3 wi, float lambda) const {
        const uint32_t gratingtype = type & TypeMask;
        const float a = k4Pi * q / (lambda * std::fabs(wi.z));
        float ix, iy;
        if (gratingtype == sinusoidal) {
            iy = ix = lobe.x == 0 ? 1 : BesselJ(lobe.x, a);
            if (!is1DGrating()) iy = lobe.y == 0 ? 1 : BesselJ(lobe.y, a);
        } else if (gratingtype == rectangular) {
            iy = ix = lobe.x == 0 ? 1 : std::sin(a / 2.f) * sincf(kPi * lobe.x / 2.f);
            if (!is1DGrating()) iy = lobe.y == 0 ? 1 : std::sin(a / 2.f) * sincf(kPi * lobe.y / 2.f);
        } else {
            iy = ix = lobe.x == 0 ? 1 : 1.f / std::sqrt(std::fabs(float(lobe.x)));
            if (!is1DGrating()) iy = lobe.y == 0 ? 1 : 1.f / std::sqrt(std::fabs(float(lobe.y)));
        }
        return intensityMultiplier * ix * iy;
    }
    int2 sampleLobe(float3 wiLocal, float2& pdf, float lambda, SampleGenerator& sg) const {
        float intensity[16];
        float total = .0f;
        for (uint32_t l = 0; l < lobeCount; ++l) {
            int2 lb; lb.x = int(l) + 1; lb.y = 0;
            const float li = lobeIntensity(lb, wiLocal, lambda);
            total += li;
            intensity[l] = li;
        }
        const float2 rand = (sampleNext2D(sg) - .5f) * 2.f;
        float cdf = .0f;
        int lobe[2] = {0, 0};
        pdf = float2(.0f);
        for (uint32_t l = 0; l < lobeCount; ++l) {
            const float p = intensity[l] / total;
            cdf += p;
            for (int i = 0; i < 2; ++i) {
                const bool select = std::fabs(rand[i]) < cdf && lobe[i] == 0;
                pdf[i] = select ? p : pdf[i];
                lobe[i] = select ? int(l) + 1 : lobe[i];
            }
        }
        pdf /= 2.f;
        for (int i = 0; i < 2; ++i) lobe[i] = rand[i] < .0f ? -lobe[i] : lobe[i];
        if (is1DGrating()) { lobe[1] = 0; pdf.y = 1.f; }
        int2 r; r.x = lobe[0]; r.y = lobe[1];
        return r;
    }
    float3 diffract(float3 wi, int2 lobe, float lambda) const {
        float2 p, sini, sino;
        const float lb[2] = {float(lobe.x), float(lobe.y)};
        for (int i = 0; i < 2; ++i) {
            p[i] = std::sqrt(sqr(wi[i]) + sqr(wi.z));
            sini[i] = p[i] > FLT_EPS ? wi[i] / p[i] : .0f;
            sino[i] = lambda * lb[i] * invPitch[i] - sini[i];
        }
        const float a = sino.x, b = sino.y;
        const float m = (sqr(a) - 1) / (sqr(a * b) - 1);
        const float qq = 1 - sqr(b) * m;
        return (std::fabs(sino.x) <= 1.f && std::fabs(sino.y) <= 1.f)
                   ? float3(a * std::sqrt(hlsl_max(.0f, qq)), b * std::sqrt(m), std::sqrt(hlsl_max(.0f, 1 - sqr(a) * qq - sqr(b) * m)))
                   : float3(.0f);
    }
    float3pdv gratingXpdv(const DifferentialShadingGeometry& dsg) const {
        float3pdv X;
        X.v = dsg.fromLocal(float3(gratingX, .0f));
        X.ddu = gratingX.x * dsg.s.ddu + gratingX.y * dsg.t.ddu;
        X.ddv = gratingX.x * dsg.s.ddv + gratingX.y * dsg.t.ddv;
        return X;
    }
    // DiffractionGrating.slang:183-265
    float3kpdv partialDerivatives(const DifferentialShadingGeometry& dsg, const float3pdv& wi, int2 lobe, float wavelength) const {
        const float3pdv n = dsg.n;
        const float3pdv X = gratingXpdv(dsg);
        if (lobe.x == 0 && lobe.y == 0) return reflectDerivatives(wi, n);
        const float lambda = 1e-3f * wavelength;
        const float2 lp = float2(lobe.x * invPitch.x, lobe.y * invPitch.y);
        const float3 Y = cross(n.v, X.v);
        const float x = dot(wi.v, X.v), y = dot(wi.v, Y), z = dot(wi.v, n.v);
        const float2 p = float2(std::sqrt(sqr(x) + sqr(z)), std::sqrt(sqr(y) + sqr(z)));
        const float2 invp = float2(p.x > FLT_EPS ? 1.f / p.x : .0f, p.y > FLT_EPS ? 1.f / p.y : .0f);
        const float2 sini = float2(x, y) * invp;
        const float2 sino = lambda * lp - sini;
        const float a = sino.x;
        const float oneMinusA2 = hlsl_max(.0f, 1 - sqr(a));
        const float b = sino.y;
        const float m = -oneMinusA2 / (sqr(a * b) - 1);
        const float qq = 1 - sqr(b) * m;
        if (sino.x > 1.f || sino.x < -1.f || sino.y > 1.f || sino.y < -1.f || qq < .0f || !(m >= .0f)) return float3kpdv();
        const float sqrtOneM = std::sqrt(qq), sqrtM = std::sqrt(m);
        const float c = .5f / (sqrtOneM * std::sqrt(oneMinusA2));
        const float dxdu = dot(wi.v, X.ddu) + dot(wi.ddu, X.v), dxdv = dot(wi.v, X.ddv) + dot(wi.ddv, X.v);
        const float dzdu = dot(wi.v, n.ddu) + dot(wi.ddu, n.v), dzdv = dot(wi.v, n.ddv) + dot(wi.ddv, n.v);
        const float2 invp3 = sqr(invp) * invp;
        const float2 dsino_du = z * (float2(x, y) * dzdu - float2(z, z) * dxdu) * invp3;
        const float2 dsino_dv = z * (float2(x, y) * dzdv - float2(z, z) * dxdv) * invp3;
        const float2 dsino_dl = lp;
        const float dadu = dsino_du.x, dbdu = dsino_du.y, dadv = dsino_dv.x, dbdv = dsino_dv.y, dadl = dsino_dl.x, dbdl = dsino_dl.y;
        const float den = sqr(sqr(a * b) - 1);
        const float dmdu = 2 * a * (a * b * oneMinusA2 * dbdu - dadu * (1 - sqr(b))) / den;
        const float dmdv = 2 * a * (a * b * oneMinusA2 * dbdv - dadv * (1 - sqr(b))) / den;
        const float dmdl = 2 * a * (a * b * oneMinusA2 * dbdl - dadl * (1 - sqr(b))) / den;
        float3kpdv ret;
        ret.v = X.v * (a * sqrtOneM) + Y * (b * sqrtM) + n.v * std::sqrt(hlsl_max(.0f, 1 - sqr(a) * qq - sqr(b) * m));
        auto deriv = [&](float da, float db, float dm) {
            return X.v * ((qq * da - a * b * (b * dm / 2.f + m * db)) / sqrtOneM) + Y * ((b * dm / 2.f + m * db) / sqrtM) +
                   n.v * ((-2 * a * da * qq - b * (b * dm + 2.f * m * db) + sqr(a) * b * (b * dm + 2.f * m * db)) * c);
        };
        ret.ddu = deriv(dadu, dbdu, dmdu);
        ret.ddv = deriv(dadv, dbdv, dmdv);
        ret.ddl = deriv(dadl, dbdl, dmdl);
        return ret;
    }
};

// ---------------------------------------------------------------------------------------------
// BxDFs (local frame).  `toLight` = ShadingData.V in local space, `toEye` = outgoing direction.
// ---------------------------------------------------------------------------------------------
struct DiffuseBxDF {  // PLTDiffuseBSDF.slang:39-95
    float3 albedo;
    void eval(float3 toLight, float3, uint32_t lobe, PLTBeam& beam) const {
        const bool valid = (lobe & L_DiffuseReflection) != 0;
        const hwss_t colour = SpectrumUtils::rgbToSpectrum(albedo, beam.wavelengths);
        const hwss_t w = valid ? std::fabs(toLight.z) * k1_Pi * colour : hwss_t(.0f);
        beam.L *= w;
        beam.L1 = beam.L2 = beam.L3 = hwss_t(.0f);
    }
    bool sample(float3, float3& wo, float& pdf, float& weight, uint32_t& lobe, SampleGenerator& sg, float wavelength, uint32_t lobeMask) const {
        const bool has = (lobeMask & L_DiffuseReflection) != 0;
        weight = SpectrumUtils::rgbToSpectrum(albedo, wavelength);
        wo = sample_cosine_hemisphere_concentric(sampleNext2D(sg), pdf);
        lobe = L_DiffuseReflection;
        return has;
    }
    float pdf(float3, float3 woLocal, float) const { return cosine_hemisphere_concentric_pdf(woLocal); }
    hwss_t lambdaPdf(float3, float3, uint32_t, const PLTBeam&, hwss_t) const { return hwss_t(1.f); }
    float2x2 coherenceTransform(float3, float3 woLocal, uint32_t) const {
        const float cosTheta = std::sqrt(hlsl_max(1e-2f, std::fabs(woLocal.z)));
        return float2x2(cosTheta, .0f, .0f, 1.f / cosTheta);
    }
};

struct ConductorBxDF {  // PLTConductorBSDF.slang:44-170
    float3 specular;
    FractalSurface gfs;
    uint32_t etaN = PLTPT_NO_SPECTRAL_PROFILE, etaK = PLTPT_NO_SPECTRAL_PROFILE;
    float extEta = 1;
    const ProfileTable* prof = nullptr;
    void eval(float3 toLight, float3 toEye, uint32_t lobe, PLTBeam& beam) const {
        const bool isScatter = (lobe & L_ScatterReflection) != 0, isDirect = (lobe & L_DeltaReflection) != 0, hasReflection = (lobe & L_Reflection) != 0;
        if ((!isScatter && !isDirect) || !hasReflection) { beam.setZero(); return; }
        const hwss_t k = waveNumbers(beam.wavelengths);
        const hwss_t a = gfs.alpha(toLight, toEye, k);
        const hwss_t m00 = SpectrumUtils::rgbToSpectrum(specular, beam.wavelengths);
        const float3 h = toLight + toEye;
        const float3 m = normalize(h);
        const float LdotH = hlsl_max(.0f, dot(toLight, m));
        const hwss_t iorN = prof->evalSpectrum(etaN, beam.wavelengths), iorK = prof->evalSpectrum(etaK, beam.wavelengths);
        for (uint32_t j = 0; j < beam.hwss_comps; ++j) {
            const float4 SP = beam.getSP(j);
            float4 newSP(.0f);
            const float2 ior = float2(iorN[j], iorK[j]) / extEta;
            if (isScatter) {
                const float4x4 M1 = MuellerFresnelRConductor(LdotH, ior);
                const float D = gfs.eval(k[j], h);
                newSP = (1 - a[j]) * std::fabs(toLight.z) * D * mul(M1, SP);
            } else {
                const float4x4 M0 = MuellerFresnelRConductor(toLight.z, ior);
                newSP = a[j] * mul(M0, SP);
            }
            newSP *= m00[j];
            beam.setSP(j, newSP);
        }
    }
    bool sample(float3 wiLocal, float3& wo, float& pdf, float& weight, uint32_t& lobe, SampleGenerator& sg, float wavelength, uint32_t lobeMask) const {
        const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
        if (!hasReflection || (!hasDirect && !hasScatter)) return false;
        const float k = k2Pi / (1e-3f * wavelength);
        const float a = gfs.alpha(wiLocal, k);
        const float m00 = SpectrumUtils::rgbToSpectrum(specular, wavelength);
        pdf = 1.f;
        bool isDirect = hasDirect;
        if (hasDirect && hasScatter) {
            const float pdfDirect = a;
            isDirect = pdfDirect == 1.f || sampleNext1D(sg) < pdfDirect;
            pdf = isDirect ? pdfDirect : 1.f - pdfDirect;
        }
        weight = m00;
        if (isDirect) { wo = reflectZ(wiLocal); lobe = L_DeltaReflection; }
        else {
            float surfPdf;
            wo = gfs.sample(wiLocal, sg, k, surfPdf);
            pdf *= surfPdf;
            lobe = L_ScatterReflection;
        }
        return true;
    }
    float pdf(float3 wiLocal, float3 woLocal, float wavelength) const {
        const float k = k2Pi / (1e-3f * wavelength);
        const float a = gfs.alpha(wiLocal, k);
        return (1 - a) * gfs.pdf(wiLocal, woLocal, k);
    }
    float weight(float wavelength) const { return SpectrumUtils::rgbToSpectrum(specular, wavelength); }
    hwss_t lambdaPdf(float3 toLight, float3 toEye, uint32_t lobe, const PLTBeam&, hwss_t wavelengths) const {
        const bool isScatter = (lobe & L_NonDelta) != 0;
        const hwss_t k = waveNumbers(wavelengths);
        if (isScatter) {
            hwss_t ret(.0f);
            for (int j = 0; j < hwss_samples(); ++j) ret[j] = gfs.pdf(toEye, toLight, k[j]);
            return ret;
        }
        return hwss_t(1.f);
    }
    float2x2 coherenceTransform(float3, float3 woLocal, uint32_t lobe) const { return scatterCoherenceTransform(woLocal, lobe); }
    float3kpdv partialDerivatives(const float3pdv& wi, const float3pdv& n) const { return reflectDerivatives(wi, n); }
};

struct DielectricBxDF {  // PLTDielectricBSDF.slang:43-270
    float3 specular = float3(1.f), transmission = float3(1.f);
    GaussianSurface gfs;
    float etaInt = 1.5f, etaExt = 1.f, dispersionB = 0;
    float indexOfRefraction(float k) const { return (etaInt + dispersionB * sqr(k * k1_2Pi)) / etaExt; }
    hwss_t indexOfRefraction(hwss_t k) const { hwss_t r; for (int j = 0; j < 4; ++j) r[j] = indexOfRefraction(k[j]); return r; }
    bool isDispersive() const { return dispersionB != .0f; }
    static float3 refract(float3 wi, float cosThetaT, float eta) {
        const float scale = cosThetaT < 0 ? 1 / eta : eta;
        return float3(-scale * wi.x, -scale * wi.y, cosThetaT);
    }
    static float3 scattered_wo(float3 wo, float eta) {
        const float scale = wo.z > 0 ? 1 / eta : eta;
        const float2 xy = wo.xy() * scale;
        const float l2 = dot(xy, xy);
        return l2 > 1 ? float3(.0f) : float3(xy, -signf(wo.z) * std::sqrt(hlsl_max(0.f, 1 - l2)));
    }
    void eval(float3 toLight, float3 toEye, uint32_t lobe, PLTBeam& beam) const {
        const bool isReflection = toLight.z * toEye.z > .0f;
        const bool isScatter = (lobe & L_NonDelta) != 0, isDirect = (lobe & L_Delta) != 0;
        const bool hasTransmission = (lobe & L_Transmission) != 0, hasReflection = (lobe & L_Reflection) != 0;
        if (toLight.z * toEye.z == 0 || (!isScatter && !isDirect) || (!hasTransmission && !isReflection) || (!hasReflection && isReflection)) { beam.setZero(); return; }
        const hwss_t k = waveNumbers(beam.wavelengths);
        const hwss_t m00 = isReflection ? SpectrumUtils::rgbToSpectrum(specular, beam.wavelengths) : SpectrumUtils::rgbToSpectrum(transmission, beam.wavelengths);
        const hwss_t eta = indexOfRefraction(k);
        hwss_t relEta2;
        for (int j = 0; j < 4; ++j) relEta2[j] = sqr(toEye.z > .0f ? eta[j] : 1.f / eta[j]);
        const float3 toEyeUp = isReflection ? toEye : scattered_wo(toEye, eta.x);
        const float3 toLightBeam = beam.toLocal(toLight);
        const hwss_t a = gfs.alpha(toLight, toEyeUp, k);
        if (isScatter) {
            const float3 h = toLight + toEyeUp;
            const float3 m = normalize(h);
            const float LdotH = hlsl_max(0.f, dot(toLight, m));
            const float2x2 Q = beam.Q();
            const float2x2 invTheta = beam.invTheta(1.f);
            for (uint32_t j = 0; j < beam.hwss_comps; ++j) {
                const float4 SP = beam.getSP(j);
                const float J = isReflection ? 1.f : relEta2[j];
                const float D = gfs.eval(k[j], invTheta, Q, h);
                const float4x4 M1 = MuellerFresnelDielectric(signf(toLight.z) * LdotH, eta[j], isReflection);
                beam.setSP(j, m00[j] * (1 - a[j]) * J * std::fabs(toLight.z) * D * mul(M1, SP));
            }
        } else {
            for (uint32_t j = 0; j < beam.hwss_comps; ++j) {
                const float4 SP = beam.getSP(j);
                const float J = isReflection ? 1.f : relEta2[j];
                const float4x4 M0 = MuellerFresnelDielectric(toLight.z, eta[j], isReflection);
                float coh = 1.f;
                if (!isReflection) {
                    float cosThetaT;
                    evalFresnelDielectric(eta[j], toEye.z, cosThetaT);
                    const float3 lobeWo = beam.toLocal(refract(toEye, cosThetaT, eta[j]));
                    coh = j == 0 ? 1.f : beam.mutualCoherenceAngular(lobeWo, toLightBeam);
                }
                beam.setSP(j, m00[j] * a[j] * coh * J * mul(M0, SP));
            }
        }
    }
    float3kpdv partialDerivatives(const float3pdv& wi, const float3pdv& n, uint32_t lobe, float wavelength) const {
        const bool isTransmission = (lobe & L_Transmission) != 0;
        const float k = k2Pi / (1e-3f * wavelength);
        const float eta = indexOfRefraction(k);
        return isTransmission ? refractDerivatives(wi, n, eta) : reflectDerivatives(wi, n);
    }
    bool sample(float3 wiLocal, float3& wo, float& pdf, float& weight, uint32_t& lobe, SampleGenerator& sg, float wavelength, uint32_t lobeMask) const {
        const float k = k2Pi / (1e-3f * wavelength);
        const float a = gfs.alpha(wiLocal, k);
        const float eta = indexOfRefraction(k);
        const bool hasReflection = (lobeMask & L_Reflection) != 0, hasTransmission = (lobeMask & L_Transmission) != 0;
        const bool hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
        if ((!hasReflection && !hasTransmission) || (!hasDirect && !hasScatter)) return false;
        const float pdfDirect = hasDirect && hasScatter ? a : hasDirect ? 1.f : .0f;
        bool isReflection;
        const bool isDirect = pdfDirect == 1.f || sampleNext1D(sg) < pdfDirect;
        pdf = (hasDirect && hasScatter) ? (isDirect ? pdfDirect : 1.f - pdfDirect) : 1.f;
        if (isDirect) {
            float cosThetaT;
            const float Fd = evalFresnelDielectric(eta, wiLocal.z, cosThetaT);
            isReflection = hasReflection ? (Fd > 0 && sampleNext1D(sg) <= Fd) : false;
            wo = isReflection ? reflectZ(wiLocal) : refract(wiLocal, cosThetaT, eta);
            pdf *= (hasReflection && hasTransmission) ? (isReflection ? Fd : 1.f - Fd) : 1.f;
            lobe = isReflection ? L_DeltaReflection : (isDispersive() ? L_DispersiveDeltaTransmission : L_DeltaTransmission);
        } else {
            float surfPdf;
            wo = gfs.sample(wiLocal, sg, k, surfPdf);
            pdf *= surfPdf;
            const float3 h = wiLocal + wo;
            const float3 m = normalize(h);
            const float F = evalFresnelDielectric(eta, signf(wiLocal.z) * dot(wiLocal, m));
            isReflection = hasReflection ? (F > 0 && sampleNext1D(sg) <= F) : false;
            if (!isReflection) wo = scattered_wo(wo, eta);
            pdf *= (hasReflection && hasTransmission) ? (isReflection ? F : 1.f - F) : 1.f;
            lobe = isReflection ? L_ScatterReflection : L_ScatterTransmission;
        }
        weight = isReflection ? SpectrumUtils::rgbToSpectrum(specular, wavelength) : SpectrumUtils::rgbToSpectrum(transmission, wavelength);
        return (isReflection && wiLocal.z * wo.z <= 0.0f) ? false : (!isReflection && wiLocal.z * wo.z >= 0.0f) ? false : true;
    }
    float pdf(float3 wiLocal, float3 woLocal, float wavelength) const {
        const float k = k2Pi / (1e-3f * wavelength);
        const float a = gfs.alpha(wiLocal, k);
        const float eta = indexOfRefraction(k);
        const float pdfDirect = a;
        const bool isReflection = wiLocal.z * woLocal.z > 0.0f;
        const float3 wo = isReflection ? woLocal : scattered_wo(woLocal, eta);
        const float3 h = wiLocal + wo;
        const float3 m = normalize(h);
        const float F = evalFresnelDielectric(eta, signf(wiLocal.z) * dot(wiLocal, m));
        const float pdfF = isReflection ? F : 1.f - F;
        return (1.f - pdfDirect) * pdfF * gfs.pdf(wiLocal, wo, k);
    }
    hwss_t lambdaPdf(float3 toLight, float3 toEye, uint32_t lobe, const PLTBeam& beam, hwss_t wavelengths) const {
        const bool isReflection = toLight.z * toEye.z > .0f;
        const bool isScatter = (lobe & L_NonDelta) != 0, isDirect = (lobe & L_Delta) != 0;
        const hwss_t k = waveNumbers(wavelengths);
        const hwss_t eta = indexOfRefraction(k);
        const float3 toLightBeam = beam.toLocal(toLight);
        if (isDirect && !isReflection) {
            hwss_t ret(.0f);
            for (int j = 0; j < hwss_samples(); ++j) {
                float cosThetaT;
                evalFresnelDielectric(eta[j], toEye.z, cosThetaT);
                const float3 lobeWo = beam.toLocal(refract(toEye, cosThetaT, eta[j]));
                ret[j] = j == 0 ? 1.f : beam.mutualCoherenceAngular(lobeWo, toLightBeam);
            }
            return ret;
        }
        if (isScatter) {
            hwss_t ret(.0f);
            for (int j = 0; j < hwss_samples(); ++j) {
                const float3 wo = isReflection ? toLight : scattered_wo(toLight, eta[j]);
                ret[j] = gfs.pdf(toEye, wo, k[j]);
            }
            return ret;
        }
        return hwss_t(1.f);
    }
    float2x2 coherenceTransform(float3, float3 woLocal, uint32_t lobe) const { return scatterCoherenceTransform(woLocal, lobe); }
};

struct OpaqueDielectricBxDF {  // PLTOpaqueDielectricBSDF.slang:43-147
    float3 albedo = float3(1.f);
    FractalSurface gfs;
    void eval(float3 toLight, float3 toEye, uint32_t lobe, PLTBeam& beam) const {
        const bool hasReflection = (lobe & L_Reflection) != 0, isDirect = (lobe & L_Delta) != 0, isScatter = (lobe & L_NonDelta) != 0;
        if ((!isScatter && !isDirect) || !hasReflection) { beam.setZero(); return; }
        const hwss_t m00 = SpectrumUtils::rgbToSpectrum(albedo, beam.wavelengths);
        const float a = gfs.alpha(toLight, toEye, kMeanK);
        float M = a;
        if (isScatter) {
            const float3 h = toLight + toEye;
            const float D = gfs.eval(kMeanK, h);
            M = (1 - a) * std::fabs(toLight.z) * D;
        }
        for (uint32_t j = 0; j < beam.hwss_comps; ++j) beam.setSP(j, M * m00[j] * beam.getSP(j));
    }
    bool sample(float3 wiLocal, float3& wo, float& pdf, float& weight, uint32_t& lobe, SampleGenerator& sg, float wavelength, uint32_t lobeMask) const {
        const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
        if (!hasReflection || (!hasDirect && !hasScatter)) return false;
        const float a = gfs.alpha(wiLocal, kMeanK);
        const float m00 = SpectrumUtils::rgbToSpectrum(albedo, wavelength);
        pdf = 1.f;
        bool isDirect = hasDirect;
        if (hasDirect && hasScatter) {
            const float pdfDirect = a;
            isDirect = pdfDirect == 1.f || sampleNext1D(sg) < pdfDirect;
            pdf = isDirect ? pdfDirect : 1.f - pdfDirect;
        }
        weight = m00;
        if (isDirect) { wo = reflectZ(wiLocal); lobe = L_DeltaReflection; }
        else {
            float surfPdf;
            wo = gfs.sample(wiLocal, sg, kMeanK, surfPdf);
            pdf *= surfPdf;
            lobe = L_ScatterReflection;
        }
        return true;
    }
    float pdf(float3 wiLocal, float3 woLocal, float) const {
        const float a = gfs.alpha(wiLocal, kMeanK);
        return (1 - a) * gfs.pdf(wiLocal, woLocal, kMeanK);
    }
    float weight(float wavelength) const { return SpectrumUtils::rgbToSpectrum(albedo, wavelength); }
    hwss_t lambdaPdf(float3, float3, uint32_t, const PLTBeam&, hwss_t) const { return hwss_t(1.f); }
    float2x2 coherenceTransform(float3, float3 woLocal, uint32_t lobe) const { return scatterCoherenceTransform(woLocal, lobe); }
    float3kpdv partialDerivatives(const float3pdv& wi, const float3pdv& n) const { return reflectDerivatives(wi, n); }
};

struct ThinDielectricBxDF {  // PLTThinDielectricBSDF.slang:43-200
    static constexpr float etaExt = 1.f;
    float2 opticAxis_xy = float2(.15f, .04f);
    float etaAvg = 1.5f, tau = 0, birefringence = 0;
    float3 opticAxis() const {
        const float2 xy = opticAxis_xy;
        return float3(xy, std::sqrt(hlsl_max(.0f, 1.f - dot(xy, xy))));
    }
    void handle_birefringence(PLTBeam& beam, float3 wi, hwss_t k) const {
        const float3 OA = opticAxis();
        const float phii = std::atan2(wi.y, wi.x);
        const float phiA = std::atan2(OA.y, OA.x) - phii;
        const float3 I = float3(0, std::fabs(wi.z), std::sqrt(1.f - sqr(wi.z)));
        float3 A = float3(std::cos(phiA), 0, std::sin(phiA));
        A *= std::sqrt(1 - sqr(OA.z));
        A.y = signf(wi.z) * OA.z;
        const float3 A2 = float3(A.x, -A.y, A.z);
        const float3 Z = beam.toLocal(float3(std::cos(phii), std::sin(phii), 0));
        const float ei = etaExt, eo = etaAvg - .5f * birefringence, ee = etaAvg + .5f * birefringence;
        const float tauum = tau * 1e+3f;
        float3 Io, Io2, Ie, Ie2;
        float e_eff, e_eff2;
        Birefringence::vectors_in_slab(I, ei, eo, ee, A, Io, Io2, Ie, Ie2, e_eff, e_eff2);
        float rss, rsp, tso, tse, rps, rpp, tpo, tpe, roo, roe, tos, top, reo, ree, tes, tep;
        Birefringence::fresnel_iso_aniso(I.y, I.z, ei, eo, ee, A, rss, rsp, tso, tse, rps, rpp, tpo, tpe);
        Birefringence::fresnel_aniso_iso(I.y, I.z, ei, eo, ee, A2, roo, roe, tos, top, reo, ree, tes, tep);
        const float ss = tso * tos + tse * tes, sp = tso * top + tse * tep, ps = tpo * tos + tpe * tes, pp = tpo * top + tpe * tep;
        const float Ioz = std::fabs(tauum / Io.y) * Io.z, Iez = std::fabs(tauum / Ie.y) * Ie.z;
        const float OPLo = std::fabs(tauum / Io.y) * eo, OPLe = std::fabs(tauum / Ie.y) * e_eff;
        const float oez = Ioz - Iez;
        const hwss_t mutual_coh = beam.mutualCoherence(k, oez * Z);
        const hwss_t t = mutual_coh * 2.f * vsin(-1.f * k * (ei * oez * I.z + OPLo - OPLe));
        const hwss_t Lx = .5f * (beam.L + beam.L1), Ly = .5f * (beam.L - beam.L1);
        const hwss_t sqrtLxLy = vsqrt(Lx * Ly);
        const hwss_t nLx = vmax(hwss_t(.0f), sqr(ss) * Lx + sqr(ps) * Ly + ss * ps * t * sqrtLxLy);
        const hwss_t nLy = vmax(hwss_t(.0f), sqr(sp) * Lx + sqr(pp) * Ly + sp * pp * t * sqrtLxLy);
        const hwss_t newL = nLx + nLy;
        // all() over the hwss_samples lanes of hwss_t
        bool allOk = true, allFinite = true;
        for (int j = 0; j < hwss_samples(); ++j) {
            allFinite = allFinite && std::isfinite(newL[j]);
            allOk = allOk && (beam.L[j] > 1e-6f && std::isfinite(newL[j]));
        }
        const hwss_t SPnormalizer = allOk ? newL / beam.L : hwss_t(.0f);
        beam.L = allFinite ? newL : hwss_t(.0f);
        beam.L1 = allFinite ? nLx - nLy : hwss_t(.0f);
        beam.L2 *= SPnormalizer;
        beam.L3 *= SPnormalizer;
    }
    void eval(float3 toLight, float3 toEye, uint32_t lobe, PLTBeam& beam) const {
        const float eta = etaAvg / etaExt;
        const bool isReflection = toLight.z * toEye.z > 0.0f;
        const bool isDirect = (lobe & L_Delta) != 0, hasTransmission = (lobe & L_Transmission) != 0, hasReflection = (lobe & L_Reflection) != 0;
        if ((isReflection && eta == 1.0f) || toLight.z * toEye.z == 0 || !isDirect || (!hasTransmission && !isReflection) || (!hasReflection && isReflection)) { beam.setZero(); return; }
        const hwss_t k = waveNumbers(beam.wavelengths);
        const float costheta_i = std::fabs(toLight.z);
        const float4x4 R = MuellerFresnelRDielectric(costheta_i, eta), iR = invOneMinusMuellerFresnelRDielectric(costheta_i, eta), T = MuellerFresnelTDielectric(costheta_i, eta);
        const float4x4 M = isReflection ? (iR.m[0][0] > 0 ? (R + mul(mul(T, R), mul(iR, T))) : R) : (iR.m[0][0] > 0 ? mul(T, mul(iR, T)) : mul(T, T));
        if (birefringence > .0f && !isReflection) handle_birefringence(beam, toLight, k);
        for (uint32_t j = 0; j < beam.hwss_comps; ++j) beam.setSP(j, mul(M, beam.getSP(j)));
    }
    float3kpdv partialDerivatives(const float3pdv& wi, const float3pdv& n, uint32_t lobe) const {
        if ((lobe & L_Transmission) != 0) {
            float3pdv wo = wi;
            wo.flip();
            float3kpdv r; r.v = wo.v; r.ddu = wo.ddu; r.ddv = wo.ddv; r.ddl = float3(.0f);
            return r;
        }
        return reflectDerivatives(wi, n);
    }
    bool sample(float3 wiLocal, float3& wo, float& pdf, float& weight, uint32_t& lobe, SampleGenerator& sg, float, uint32_t lobeMask) const {
        const bool hasReflection = (lobeMask & L_Reflection) != 0, hasTransmission = (lobeMask & L_Transmission) != 0;
        if (!hasReflection && !hasTransmission) return false;
        const float eta = etaAvg / etaExt;
        const float Fd = evalFresnelDielectric(eta, std::fabs(wiLocal.z));
        const bool isReflection = hasReflection ? (Fd > 0 && sampleNext1D(sg) <= Fd) : false;
        wo = isReflection ? reflectZ(wiLocal) : -wiLocal;
        weight = 1.f;
        pdf = hasReflection && hasTransmission ? (isReflection ? Fd : 1.f - Fd) : 1.f;
        lobe = isReflection ? L_DeltaReflection : L_DeltaTransmission;
        return true;
    }
    float pdf(float3, float3, float) const { return .0f; }
    hwss_t lambdaPdf(float3, float3, uint32_t, const PLTBeam&, hwss_t) const { return hwss_t(1.f); }
    float2x2 coherenceTransform(float3, float3, uint32_t) const { return ident2(); }
};

// PLTCoatedConductorBSDF.slang:45-245 and PLTCoatedOpaqueDielectricBSDF.slang:44-244 (conductor == false)
struct CoatedBxDF {
    bool conductor = true;
    float3 specular;  // "albedo" for the opaque-dielectric variant
    FractalSurface gfs;
    uint32_t etaN = PLTPT_NO_SPECTRAL_PROFILE, etaK = PLTPT_NO_SPECTRAL_PROFILE;
    float extEta = 1, coatIOR = 1.5f, tau = 0;
    const ProfileTable* prof = nullptr;
    static float3 refractCoat(float3 wi, float eta, float cosThetaT) {
        const float scale = cosThetaT < 0 ? 1 / eta : eta;
        return float3(-scale * wi.x, -scale * wi.y, cosThetaT);
    }
    float coatEta() const {
        const float minTau = .05f;
        return lerpf(extEta, coatIOR, clampf(tau, 0, minTau) / minTau);
    }
    void eval(float3 toLight, float3 toEye, uint32_t lobe, PLTBeam& beam) const {
        const bool isReflection = toLight.z * toEye.z > 0.0f;
        const bool isScatter = (lobe & L_ScatterReflection) != 0, isDirect = (lobe & L_DeltaReflection) != 0;
        const float coatior = coatEta();
        const float eta1 = coatior / extEta;
        hwss_t iorN(.0f), iorK(.0f);
        if (conductor) { iorN = prof->evalSpectrum(etaN, beam.wavelengths); iorK = prof->evalSpectrum(etaK, beam.wavelengths); }
        float cosThetaTin;
        evalFresnelDielectric(eta1, toLight.z, cosThetaTin);
        const float3 wiIn = -refractCoat(toLight, eta1, cosThetaTin);
        float3 woIn = reflectZ(wiIn);
        if (!isDirect) {
            float cosThetaTout;
            evalFresnelDielectric(eta1, toEye.z, cosThetaTout);
            woIn = -refractCoat(toEye, eta1, cosThetaTout);
        }
        if (!isReflection || (!isScatter && !isDirect) || wiIn.z <= .0f || woIn.z <= .0f) { beam.setZero(); return; }
        const hwss_t k = waveNumbers(beam.wavelengths);
        const hwss_t a = gfs.alpha(wiIn, woIn, k);
        const hwss_t m00 = SpectrumUtils::rgbToSpectrum(specular, beam.wavelengths);
        const float3 h = wiIn + woIn;
        const float3 m = normalize(h);
        const float LdotH = hlsl_max(0.f, dot(wiIn, m));
        const float3 Z = beam.toLocal(float3(0, 0, 1));
        const hwss_t mutual_coh = beam.mutualCoherence(k, tau * Z);
        const float4x4 MiT = MuellerFresnelTDielectric(toLight.z, eta1);
        const float4x4 MoT = MuellerFresnelTDielectric(-woIn.z, eta1);
        if (conductor) {
            const float4x4 MiR = MuellerFresnelRDielectric(toLight.z, eta1);
            for (uint32_t j = 0; j < beam.hwss_comps; ++j) {
                float2 rs, rp;
                const float2 ior = float2(iorN[j], iorK[j]) / coatior;
                const float4x4 M2R = MuellerFresnelRConductor(LdotH, ior, rs, rp);
                const float4x4 M = mul(mul(MiT, M2R), MoT);
                const float4 SP = beam.getSP(j);
                float4 newSP = m00[j] * mul(M, SP);
                if (isScatter) {
                    const float D = gfs.eval(k[j], h);
                    newSP *= (1 - a[j]) * std::fabs(toLight.z) * D;
                } else {
                    const float pS = phaseShift(coatior, 2.f * tau / wiIn.z, k[j], rs);
                    const float pP = phaseShift(coatior, 2.f * tau / wiIn.z, k[j], rp);
                    newSP = interfere(a[j] * newSP, mul(MiR, SP), pS, pP, mutual_coh[j]);
                }
                beam.setSP(j, newSP);
            }
        } else {
            const hwss_t phases = phaseShift(coatior, 2.f * tau / wiIn.z, k);
            const float4x4 M = mul(MiT, MoT);
            for (uint32_t j = 0; j < beam.hwss_comps; ++j) {
                const float4 SP = beam.getSP(j);
                float4 newSP = m00[j] * mul(M, SP);
                float4 SPreflcoat;
                if (isScatter) {
                    const float4x4 MiR = MuellerFresnelRDielectric(LdotH, eta1);
                    const float D = gfs.eval(k[j], h);
                    newSP *= (1 - a[j]) * std::fabs(toLight.z) * D;
                    SPreflcoat = (1 - a[j]) * std::fabs(toLight.z) * D * mul(MiR, SP);
                } else {
                    const float4x4 MiR = MuellerFresnelRDielectric(toLight.z, eta1);
                    newSP *= a[j];
                    SPreflcoat = a[j] * mul(MiR, SP);
                }
                newSP = interfere(newSP, SPreflcoat, phases[j], phases[j], mutual_coh[j]);
                beam.setSP(j, newSP);
            }
        }
    }
    bool sample(float3 wiLocal, float3& wo, float& pdf, float& weight, uint32_t& lobe, SampleGenerator& sg, float wavelength, uint32_t lobeMask) const {
        const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
        if (!hasReflection || (!hasDirect && !hasScatter) || wiLocal.z <= .0f) return false;
        const float k = k2Pi / (1e-3f * wavelength);
        const float eta1 = coatEta() / extEta;
        float cosThetaT;
        const float Fd = evalFresnelDielectric(eta1, wiLocal.z, cosThetaT);
        const bool isReflCoat = Fd == .1f || sampleNext1D(sg) < Fd;  // sic (PLTCoatedConductorBSDF.slang:151)
        weight = 1.f;
        pdf = isReflCoat ? Fd : 1.f - Fd;
        if (isReflCoat) { wo = reflectZ(wiLocal); lobe = L_DeltaReflection; return true; }
        const float3 wiIn = -refractCoat(wiLocal, eta1, cosThetaT);
        if (wiIn.z == .0f) return false;
        weight = SpectrumUtils::rgbToSpectrum(specular, wavelength);
        const float a = gfs.alpha(wiIn, k);
        const float pdfDirect = hasDirect && hasScatter ? a : hasDirect ? 1.f : .0f;
        const bool isDirect = pdfDirect == 1.f || sampleNext1D(sg) < pdfDirect;
        if (hasDirect && hasScatter) pdf *= isDirect ? pdfDirect : 1.f - pdfDirect;
        if (isDirect) { wo = reflectZ(wiLocal); lobe = L_DeltaReflection; }
        else {
            float surfPdf;
            float3 woIn = gfs.sample(wiIn, sg, k, surfPdf);
            pdf *= surfPdf;
            float cosThetaT1;
            evalFresnelDielectric(eta1, -woIn.z, cosThetaT1);
            wo = refractCoat(-woIn, eta1, cosThetaT1);
            lobe = L_ScatterReflection;
        }
        return wiLocal.z * wo.z > .0f;
    }
    float pdf(float3 wiLocal, float3 woLocal, float wavelength) const {
        if (wiLocal.z <= .0f || woLocal.z <= .0f) return .0f;
        const float k = k2Pi / (1e-3f * wavelength);
        const float eta1 = coatEta() / extEta;
        float cosThetaT, cosThetaT1;
        const float Fd = evalFresnelDielectric(eta1, wiLocal.z, cosThetaT);
        evalFresnelDielectric(eta1, woLocal.z, cosThetaT1);
        const float3 wiIn = -refractCoat(wiLocal, eta1, cosThetaT);
        const float3 woIn = -refractCoat(woLocal, eta1, cosThetaT1);
        const float a = gfs.alpha(wiIn, k);
        return wiIn.z * woIn.z == .0f ? .0f : (1.f - a) * (1.f - Fd) * gfs.pdf(wiIn, woIn, k);
    }
    hwss_t lambdaPdf(float3 toLight, float3 toEye, uint32_t lobe, const PLTBeam&, hwss_t wavelengths) const {
        const bool isScatter = (lobe & L_NonDelta) != 0;
        const hwss_t k = waveNumbers(wavelengths);
        const float eta1 = coatEta() / extEta;
        if (isScatter) {
            float cosThetaT, cosThetaT1;
            evalFresnelDielectric(eta1, toEye.z, cosThetaT);
            evalFresnelDielectric(eta1, toLight.z, cosThetaT1);
            const float3 wiIn = -refractCoat(toEye, eta1, cosThetaT);
            const float3 woIn = -refractCoat(toLight, eta1, cosThetaT1);
            hwss_t ret(.0f);
            for (int j = 0; j < hwss_samples(); ++j) ret[j] = gfs.pdf(wiIn, woIn, k[j]);
            return ret;
        }
        return hwss_t(1.f);
    }
    float2x2 coherenceTransform(float3, float3 woLocal, uint32_t lobe) const { return scatterCoherenceTransform(woLocal, lobe); }
};

// PLTDiffractionGratedConductorBSDF.slang:45-195 / PLTDiffractionGratedOpaqueDielectricBSDF.slang:44-192
struct GratedBxDF {
    bool conductor = true;
    ConductorBxDF baseC;
    OpaqueDielectricBxDF baseD;
    DiffractionGrating grating;
    static void encodeLobeType(int2 lobe, uint32_t& bsdfLobe) {
        const int sx = lobe.x < 0 ? 0 : 1, sy = lobe.y < 0 ? 0 : 1;
        const int lx = hlsl_min(std::abs(lobe.x), 7), ly = hlsl_min(std::abs(lobe.y), 7);
        bsdfLobe &= ~uint32_t(L_MaterialSpecific);
        bsdfLobe += uint32_t((ly << 3) + lx + (sy << 7) + (sx << 6)) << LobeTypeMaterialSpecificOffset;
    }
    static int2 decodeLobeType(uint32_t bsdfLobe) {
        const int x = int(bsdfLobe) >> LobeTypeMaterialSpecificOffset;
        int2 lobe;
        lobe.x = (x & 0x7) * (2 * ((x >> 6) & 0x1) - 1);
        lobe.y = ((x >> 3) & 0x7) * (2 * ((x >> 7) & 0x1) - 1);
        return lobe;
    }
    float3 baseColour() const { return conductor ? baseC.specular : baseD.albedo; }
    float baseWeight(float wavelength) const { return conductor ? baseC.weight(wavelength) : baseD.weight(wavelength); }
    void eval(float3 toLight, float3 toEye, uint32_t bsdfLobe, PLTBeam& beam) const {
        const bool isDirect = (bsdfLobe & L_Delta) != 0;
        const hwss_t lambdas = 1e-3f * beam.wavelengths;
        const hwss_t k = k2Pi / lambdas;
        const hwss_t a = grating.alpha(toLight, k);
        const int2 lobe = decodeLobeType(bsdfLobe);
        const float3 toLightBeam = beam.toLocal(toLight);
        bool anyAPos = false, anyALt1 = false;
        for (int j = 0; j < hwss_samples(); ++j) { anyAPos = anyAPos || a[j] > .0f; anyALt1 = anyALt1 || a[j] < 1.f; }
        const bool lobeZero = lobe.x == 0 && lobe.y == 0;
        if (lobeZero && anyAPos) {
            if (conductor) baseC.eval(toLight, toEye, bsdfLobe, beam); else baseD.eval(toLight, toEye, bsdfLobe, beam);
            beam.scale(a);
        } else if (!lobeZero && anyALt1 && isDirect) {
            const hwss_t m00 = SpectrumUtils::rgbToSpectrum(baseColour(), beam.wavelengths);
            hwss_t iorN(.0f), iorK(.0f);
            if (conductor) { iorN = baseC.prof->evalSpectrum(baseC.etaN, beam.wavelengths); iorK = baseC.prof->evalSpectrum(baseC.etaK, beam.wavelengths); }
            for (uint32_t j = 0; j < beam.hwss_comps; ++j) {
                const float3 lobeWi = beam.toLocal(grating.diffract(toEye, lobe, lambdas[j]));
                const float coh = j == 0 ? 1.f : beam.mutualCoherenceAngular(lobeWi, toLightBeam);
                const float intensity = grating.lobeIntensity(lobe, toLight, lambdas[j]);
                if (conductor) {
                    const float2 ior = float2(iorN[j], iorK[j]) / baseC.extEta;
                    const float4x4 M0 = MuellerFresnelRConductor(toLight.z, ior);
                    beam.setSP(j, m00[j] * (1 - a[j]) * coh * intensity * mul(M0, beam.getSP(j)));
                } else {
                    beam.setSP(j, m00[j] * (1 - a[j]) * coh * intensity * beam.getSP(j));
                }
            }
        } else beam.setZero();
    }
    bool sample(float3 wiLocal, float3& wo, float& pdf, float& weight, uint32_t& bsdfLobe, SampleGenerator& sg, float wavelength, uint32_t lobeMask, float3 woHint) const {
        const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
        if (!hasReflection || (!hasDirect && !hasScatter)) return false;
        const float lambdas = 1e-3f * wavelength;
        const float k = k2Pi / lambdas;
        const float a = grating.alpha(wiLocal, k);
        const bool hasHint = hasDirect && (woHint.x != .0f || woHint.y != .0f || woHint.z != .0f);
        int2 lobe;
        if (hasHint && a < 1.f) lobe = grating.closestLobeHint(wiLocal, woHint, a, lambdas);
        bool isDirect = hasDirect;
        // NOTE: the reference leaves `pdf` unwritten when neither branch below runs; the oracle starts from 1.
        pdf = 1.f;
        if (hasHint) { isDirect = lobe.x == 0 && lobe.y == 0; pdf = 1.f; }
        else if (hasDirect && hasScatter) {
            const float pdfDirect = a;
            isDirect = pdfDirect == 1.f || sampleNext1D(sg) < pdfDirect;
            pdf = isDirect ? pdfDirect : 1.f - pdfDirect;
        }
        if (isDirect) {
            float basePdf = 0;
            if (conductor) baseC.sample(wiLocal, wo, basePdf, weight, bsdfLobe, sg, wavelength, lobeMask);
            else baseD.sample(wiLocal, wo, basePdf, weight, bsdfLobe, sg, wavelength, lobeMask);
            pdf *= basePdf;
            return true;
        }
        weight = baseWeight(wavelength);
        float2 lobePdf(1.f);
        if (!hasHint) lobe = grating.sampleLobe(wiLocal, lobePdf, lambdas, sg);
        wo = grating.diffract(wiLocal, lobe, lambdas);
        bsdfLobe = L_DispersiveDeltaReflection;
        encodeLobeType(lobe, bsdfLobe);
        pdf *= lobePdf.x * lobePdf.y;
        return true;
    }
    float pdf(float3 wiLocal, float3 woLocal, float wavelength) const {
        const float k = k2Pi / (1e-3f * wavelength);
        const float a = grating.alpha(wiLocal, k);
        return (conductor ? baseC.pdf(wiLocal, woLocal, wavelength) : baseD.pdf(wiLocal, woLocal, wavelength)) * a;
    }
    hwss_t lambdaPdf(float3 toLight, float3 toEye, uint32_t bsdfLobe, const PLTBeam& beam, hwss_t wavelengths) const {
        const bool isDirect = (bsdfLobe & L_Delta) != 0;
        const hwss_t lambdas = 1e-3f * wavelengths;
        const hwss_t k = k2Pi / lambdas;
        const hwss_t a = grating.alpha(toLight, k);
        const int2 lobe = decodeLobeType(bsdfLobe);
        const float3 toLightBeam = beam.toLocal(toLight);
        bool anyALt1 = false;
        for (int j = 0; j < hwss_samples(); ++j) anyALt1 = anyALt1 || a[j] < 1.f;
        if ((lobe.x != 0 || lobe.y != 0) && anyALt1 && isDirect) {
            hwss_t ret(.0f);
            for (int j = 0; j < hwss_samples(); ++j) {
                const float3 lobeWi = beam.toLocal(grating.diffract(toEye, lobe, lambdas[j]));
                ret[j] = j == 0 ? 1.f : beam.mutualCoherenceAngular(lobeWi, toLightBeam);
            }
            return ret;
        }
        return hwss_t(1.f);
    }
    float2x2 coherenceTransform(float3, float3 woLocal, uint32_t bsdfLobe) const {
        const int2 lobe = decodeLobeType(bsdfLobe);
        if (lobe.x == 0 && lobe.y == 0) return scatterCoherenceTransform(woLocal, bsdfLobe);
        return ident2();
    }
};

struct MultiLayeredStackBxDF {  // PLTMultiLayeredStackBSDF.slang:42-140
    float3 albedo = float3(1.f);
    float extEta = 1;
    float2 ior, thickness;
    uint32_t layers = 0;
    static float2 phasor(float cosThetaT, float thickness, float eta, float k) {
        const float d = 2.f * thickness / cosThetaT * k * eta;
        // Airy summation over the layers is an interference sum: near anti-reflection the result cancels to ~1e-4 of its terms,
        // so the phase factor is DEFINED as the correctly rounded fp32 cos/sin (evaluated in double), like exp_cr above.
        return float2(float(std::cos(double(d))), float(std::sin(double(d))));
    }
    void eval(float3 toLight, float3 toEye, uint32_t lobe, PLTBeam& beam) const {
        const bool isReflection = toLight.z * toEye.z > .0f;
        const bool isDirect = (lobe & L_Delta) != 0;
        if (!isDirect || !isReflection) { beam.setZero(); return; }
        const hwss_t k = waveNumbers(beam.wavelengths);
        const hwss_t m00 = SpectrumUtils::rgbToSpectrum(albedo, beam.wavelengths);
        const float eta1 = ior.x / extEta, eta2 = ior.y / ior.x;
        float2 cosThetaT, r;
        const float Rout = std::sqrt(evalFresnelDielectric(eta1, toLight.z, cosThetaT.x));
        r.x = std::sqrt(evalFresnelDielectric(eta2, -cosThetaT.x, cosThetaT.y));
        r.y = std::sqrt(evalFresnelDielectric(eta2, cosThetaT.y));
        cosThetaT = vabs(cosThetaT);
        for (uint32_t j = 0; j < beam.hwss_comps; ++j) {
            float2 G = c_t(std::sqrt(m00[j]));
            for (uint32_t l = 0; l < layers; ++l) {
                const uint32_t idx = l % 2;
                const float2 p = phasor(cosThetaT[idx], thickness[idx], ior[idx], k[j]);
                const float R = l == layers - 1 ? Rout : r[idx];
                G = cmul(G, p);
                G = cdiv(c_t(R) + G, c_t(1.f) + R * G);
            }
            beam.setSP(j, sqr(G.x) * beam.getSP(j));
        }
    }
    bool sample(float3 wiLocal, float3& wo, float& pdf, float& weight, uint32_t& lobe, SampleGenerator&, float wavelength, uint32_t lobeMask) const {
        const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0;
        if (!hasReflection || !hasDirect || wiLocal.z <= .0f) return false;
        pdf = 1.f;
        wo = reflectZ(wiLocal);
        lobe = L_DeltaReflection;
        weight = SpectrumUtils::rgbToSpectrum(albedo, wavelength);
        return wiLocal.z * wo.z > .0f;
    }
    float pdf(float3, float3, float) const { return .0f; }
    hwss_t lambdaPdf(float3, float3, uint32_t, const PLTBeam&, hwss_t) const { return hwss_t(1.f); }
    float2x2 coherenceTransform(float3, float3, uint32_t) const { return ident2(); }
};

// ---------------------------------------------------------------------------------------------
// Material instance: unpacks the 128 B blob (fp16 parameters) exactly as the *Material.slang modules do and
// wraps the BxDF with the common world<->local prologue/epilogue (e.g. PLTConductorBSDF.slang:179-196).
// ---------------------------------------------------------------------------------------------
inline float h2f(uint16_t h) { return f16tof32(h); }

// ---------------------------------------------------------------------------------------------
// Material textures: MaterialSystem.slang:283-291 (sampleTexture: Uniform handle -> the uniform value, Texture handle ->
// materialTextures[id].SampleLevel(sampler, uv, lod)) with the explicit LOD 0 every PLTPT call site passes
// (pltpt_sample.rt.slang:76, pltpt_solve.rt.slang:53,229,433, Manifolds.slangh:128,213).  The texture unit is hardware in the
// reference; the oracle DEFINES the filter the CUDA path must reproduce: D3D texel addressing (texel centres at (i + .5) / size),
// sRGB decode per texel before filtering, bilinear weights in exact fp32 (not the unit's 8-bit fixed point), address modes of
// Core/API/Sampler.h:59-66.
// ---------------------------------------------------------------------------------------------
struct TextureSystem {
    const PltptTextureDesc* textures = nullptr;
    uint32_t textureCount = 0;
    const PltptSamplerDesc* samplers = nullptr;
    uint32_t samplerCount = 0;

    static uint32_t handleMode(uint32_t h) { return (h >> 29) & 3u; }       // TextureHandle.slang:54-67
    static uint32_t handleTextureID(uint32_t h) { return h & 0x1fffffffu; }  // :75
    static uint32_t samplerIDOf(const PltptMaterialBlob& m) { return (m.header[1] >> 17) & 0xffu; }  // MaterialData.slang:66,122
    static uint32_t normalMapTypeOf(const PltptMaterialBlob& m) { return (m.flags >> 1) & 3u; }        // BasicMaterialData.slang:106,126

    static float srgbToLinear(float c) { return c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); }
    static float4 fetch(const PltptTextureDesc& t, int x, int y) {
        const size_t i = size_t(y) * t.width + size_t(x);
        switch (t.format) {
            case PLTPT_TEXTURE_FORMAT_RGBA32_FLOAT: { const float* p = (const float*)t.texels + 4 * i; return float4(p[0], p[1], p[2], p[3]); }
            case PLTPT_TEXTURE_FORMAT_RGBA16_FLOAT: { const uint16_t* p = (const uint16_t*)t.texels + 4 * i; return float4(f16tof32(p[0]), f16tof32(p[1]), f16tof32(p[2]), f16tof32(p[3])); }
            case PLTPT_TEXTURE_FORMAT_RGBA8_UNORM: { const uint8_t* p = (const uint8_t*)t.texels + 4 * i; return float4(p[0] / 255.f, p[1] / 255.f, p[2] / 255.f, p[3] / 255.f); }
            case PLTPT_TEXTURE_FORMAT_RGBA8_UNORM_SRGB: {
                const uint8_t* p = (const uint8_t*)t.texels + 4 * i;
                return float4(srgbToLinear(p[0] / 255.f), srgbToLinear(p[1] / 255.f), srgbToLinear(p[2] / 255.f), p[3] / 255.f);
            }
            default: return float4(0.f);
        }
    }
    // texel index under an address mode; false = outside (Border)
    static bool address(int& i, int n, uint32_t mode) {
        switch (mode) {
            case PLTPT_ADDRESS_WRAP: i %= n; if (i < 0) i += n; return true;
            case PLTPT_ADDRESS_MIRROR: { int p = i % (2 * n); if (p < 0) p += 2 * n; i = p < n ? p : 2 * n - 1 - p; return true; }
            case PLTPT_ADDRESS_CLAMP: i = i < 0 ? 0 : (i >= n ? n - 1 : i); return true;
            case PLTPT_ADDRESS_MIRROR_ONCE: if (i < 0) i = -1 - i; i = i >= n ? n - 1 : i; return true;
            default: return i >= 0 && i < n;  // Border
        }
    }
    float4 texel(const PltptTextureDesc& t, const PltptSamplerDesc& s, int x, int y) const {
        if (!address(x, int(t.width), s.addressU) || !address(y, int(t.height), s.addressV)) return float4(s.borderColor[0], s.borderColor[1], s.borderColor[2], s.borderColor[3]);
        return fetch(t, x, y);
    }
    float4 sampleLevel0(const PltptTextureDesc& t, const PltptSamplerDesc& s, float2 uv) const {
        if (s.magFilter == PLTPT_FILTER_POINT) return texel(t, s, int(std::floor(uv.x * float(t.width))), int(std::floor(uv.y * float(t.height))));
        const float x = uv.x * float(t.width) - .5f, y = uv.y * float(t.height) - .5f;
        const float x0 = std::floor(x), y0 = std::floor(y);
        const float fx = x - x0, fy = y - y0;
        const int ix = int(x0), iy = int(y0);
        const float4 c00 = texel(t, s, ix, iy), c10 = texel(t, s, ix + 1, iy), c01 = texel(t, s, ix, iy + 1), c11 = texel(t, s, ix + 1, iy + 1);
        const float4 top = c00 * (1.f - fx) + c10 * fx, bottom = c01 * (1.f - fx) + c11 * fx;
        return top * (1.f - fy) + bottom * fy;
    }
    // ms.sampleTexture(handle, s, uv, uniformValue, lod) with the material's default sampler
    float4 sampleTexture(uint32_t handle, uint32_t samplerID, float2 uv, float4 uniformValue) const {
        if (handleMode(handle) == 0) return uniformValue;                              // Mode::Uniform
        if (handleMode(handle) != 1 || handleTextureID(handle) >= textureCount) return float4(0.f);  // default: float4(0)
        PltptSamplerDesc def{};                                                          // MaterialSystem.cpp:76-79: linear, wrap
        def.magFilter = PLTPT_FILTER_LINEAR;
        return sampleLevel0(textures[handleTextureID(handle)], samplerID < samplerCount ? samplers[samplerID] : def, uv);
    }
};

struct MaterialInstance {
    uint32_t type = 0;
    DiffuseBxDF diffuse;
    ConductorBxDF conductor;
    DielectricBxDF dielectric;
    OpaqueDielectricBxDF opaque;
    ThinDielectricBxDF thin;
    CoatedBxDF coated;
    GratedBxDF grated;
    MultiLayeredStackBxDF stack;
    uint32_t emissionSpectralId = PLTPT_NO_SPECTRAL_PROFILE;
    float3 emissionRGB;
    const ProfileTable* prof = nullptr;

    static uint32_t materialType(const PltptMaterialBlob& m) { return m.header[0] & 0xffffu; }
    // evalOpacity of each *Material.slang (Dielectric/ThinDielectric halve it; coated return 1)
    static float evalOpacity(const PltptMaterialBlob& m, const TextureSystem& tex, float2 uv) {
        const float a = tex.sampleTexture(m.texBaseColor, TextureSystem::samplerIDOf(m), uv, float4(h2f(m.baseColor[0]), h2f(m.baseColor[1]), h2f(m.baseColor[2]), h2f(m.baseColor[3]))).w;
        switch (materialType(m)) {
            case PLTPT_MATERIAL_DIELECTRIC: case PLTPT_MATERIAL_THIN_DIELECTRIC: return .5f * a;
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: return 1.f;
            default: return a;
        }
    }
    MaterialInstance() = default;
    MaterialInstance(const PltptMaterialBlob& m, const ShadingData& sd, const ProfileTable* profiles, const TextureSystem& tex) : prof(profiles) {
        type = materialType(m);
        // every *Material.slang::setupMaterialInstance samples its parameter slots at sd.uv with the material's default sampler
        const uint32_t smp = TextureSystem::samplerIDOf(m);
        const float4 base = tex.sampleTexture(m.texBaseColor, smp, sd.uv, float4(h2f(m.baseColor[0]), h2f(m.baseColor[1]), h2f(m.baseColor[2]), h2f(m.baseColor[3])));
        const float4 spec = tex.sampleTexture(m.texSpecular, smp, sd.uv, float4(h2f(m.specular[0]), h2f(m.specular[1]), h2f(m.specular[2]), h2f(m.specular[3])));
        const float4 data1 = tex.sampleTexture(m.texData1, smp, sd.uv, float4(h2f(m.data1[0]), h2f(m.data1[1]), h2f(m.data1[2]), h2f(m.data1[3])));
        const float4 data2 = tex.sampleTexture(m.texData2, smp, sd.uv, float4(h2f(m.data2[0]), h2f(m.data2[1]), h2f(m.data2[2]), h2f(m.data2[3])));
        const float3 tran = tex.sampleTexture(m.texTransmission, smp, sd.uv, float4(h2f(m.transmission[0]), h2f(m.transmission[1]), h2f(m.transmission[2]), 1.f)).xyz();
        const float IoR = h2f(m.IoR);
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE:  // PLTDiffuseMaterial.slang:40-50
                diffuse.albedo = base.xyz();
                emissionSpectralId = m.emissionSpectralId;
                emissionRGB = tex.sampleTexture(m.texEmissive, smp, sd.uv, float4(1.f)).xyz();
                break;
            case PLTPT_MATERIAL_CONDUCTOR:  // PLTConductorMaterial.slang:40-57
                conductor.specular = base.xyz();
                conductor.gfs = FractalSurface(spec.y, spec.x);
                conductor.etaN = m.iorNSpectralId; conductor.etaK = m.iorKSpectralId; conductor.extEta = IoR; conductor.prof = prof;
                break;
            case PLTPT_MATERIAL_DIELECTRIC: {  // PLTDielectricMaterial.slang:40-57
                const float dispersionB = (tran.x > .0f || tran.y > .0f || tran.z > .0f) ? spec.x : .0f;
                dielectric.specular = base.xyz(); dielectric.transmission = tran;
                dielectric.gfs = GaussianSurface(spec.y);
                dielectric.etaInt = hlsl_max(1.f, IoR - 2.897f * dispersionB);
                dielectric.etaExt = spec.w; dielectric.dispersionB = dispersionB;
                break;
            }
            case PLTPT_MATERIAL_OPAQUE_DIELECTRIC:  // PLTOpaqueDielectricMaterial.slang:40-52
                opaque.albedo = base.xyz();
                opaque.gfs = FractalSurface(spec.y);
                break;
            case PLTPT_MATERIAL_THIN_DIELECTRIC:  // PLTThinDielectricMaterial.slang:40-54
                thin.opticAxis_xy = float2(base.y, base.z); thin.etaAvg = IoR; thin.tau = spec.w; thin.birefringence = base.x * spec.x;
                break;
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC:  // PLTCoatedConductorMaterial.slang:40-60
                coated.conductor = type == PLTPT_MATERIAL_COATED_CONDUCTOR;
                coated.specular = base.xyz(); coated.gfs = FractalSurface(spec.y);
                coated.etaN = m.iorNSpectralId; coated.etaK = m.iorKSpectralId;
                coated.extEta = IoR; coated.coatIOR = data1.w; coated.tau = data1.x * spec.z; coated.prof = prof;
                break;
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {  // PLTDiffractionGratedConductorMaterial.slang:41-73
                const float2 pitch(data1.x, data1.y);
                const float2 invPitch(pitch.x > .0f ? 1.f / pitch.x : .0f, pitch.y > .0f ? 1.f / pitch.y : .0f);
                const float q = data2.x * spec.x, powerMult = spec.y, roughness = data2.y;
                const uint32_t typeAndLobes = m.data1[3];
                const uint32_t lobes = hlsl_min(hlsl_max(typeAndLobes & 0xFFu, 1u), 8u);
                const uint32_t gtype = typeAndLobes >> 8;
                grated.conductor = type == PLTPT_MATERIAL_GRATED_CONDUCTOR;
                grated.baseC.specular = base.xyz(); grated.baseC.gfs = FractalSurface(roughness);
                grated.baseC.etaN = m.iorNSpectralId; grated.baseC.etaK = m.iorKSpectralId; grated.baseC.extEta = IoR; grated.baseC.prof = prof;
                grated.baseD.albedo = base.xyz(); grated.baseD.gfs = FractalSurface(roughness);
                grated.grating = DiffractionGrating(data1.z, sd, invPitch, q, lobes, gtype, powerMult);
                break;
            }
            case PLTPT_MATERIAL_MULTILAYERED_STACK:  // PLTMultiLayeredStackMaterial.slang:40-56
                stack.albedo = base.xyz(); stack.extEta = IoR; stack.ior = float2(data2.x, data2.y);
                stack.thickness = float2(data1.x, data1.y); stack.layers = m.data1[3];
                break;
            default: break;
        }
    }
    bool isGrated() const { return type == PLTPT_MATERIAL_GRATED_CONDUCTOR || type == PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC; }
    bool isCoated() const { return type == PLTPT_MATERIAL_COATED_CONDUCTOR || type == PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC; }
    const FractalSurface& baseGfs() const { return grated.conductor ? grated.baseC.gfs : grated.baseD.gfs; }

    void bxdfEval(float3 wi, float3 wo, uint32_t lobe, PLTBeam& beam) const {
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE: diffuse.eval(wi, wo, lobe, beam); break;
            case PLTPT_MATERIAL_CONDUCTOR: conductor.eval(wi, wo, lobe, beam); break;
            case PLTPT_MATERIAL_DIELECTRIC: dielectric.eval(wi, wo, lobe, beam); break;
            case PLTPT_MATERIAL_OPAQUE_DIELECTRIC: opaque.eval(wi, wo, lobe, beam); break;
            case PLTPT_MATERIAL_THIN_DIELECTRIC: thin.eval(wi, wo, lobe, beam); break;
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: coated.eval(wi, wo, lobe, beam); break;
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: grated.eval(wi, wo, lobe, beam); break;
            case PLTPT_MATERIAL_MULTILAYERED_STACK: stack.eval(wi, wo, lobe, beam); break;
            default: beam.setZero();
        }
    }
    float2x2 bxdfCoherenceTransform(float3 wi, float3 wo, uint32_t lobe) const {
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE: return diffuse.coherenceTransform(wi, wo, lobe);
            case PLTPT_MATERIAL_THIN_DIELECTRIC: case PLTPT_MATERIAL_MULTILAYERED_STACK: return ident2();
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: return grated.coherenceTransform(wi, wo, lobe);
            default: return scatterCoherenceTransform(wo, lobe);
        }
    }
    // IMaterialInstance.eval
    void eval(const ShadingData& sd, float3 toEye, uint32_t lobe, PLTBeam& beam) const {
        const ShadingData fr = isGrated() ? grated.grating.localGratingShadingFrame(sd) : sd;
        beam.r = fr.toLocal(beam.r);
        beam.t = fr.toLocal(beam.t);
        const float3 wiLocal = fr.toLocal(sd.V);
        const float3 woLocal = fr.toLocal(toEye);
        beam.constructLocalFrame(woLocal);
        bxdfEval(wiLocal, woLocal, lobe, beam);
        beam.transformCoherence(bxdfCoherenceTransform(wiLocal, woLocal, lobe));
        beam.r = fr.fromLocal(beam.r);
        beam.t = fr.fromLocal(beam.t);
    }
    // IMaterialInstance.sample
    bool sample(const ShadingData& sd, SampleGenerator& sg, BSDFSample& result, float wavelength, uint32_t lobeMask = L_All, float3 woHint = float3(.0f)) const {
        bool valid = false;
        const float3 wi = sd.toLocal(sd.V);
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE:
                valid = diffuse.sample(wi, result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobeMask);
                result.wo = sd.fromLocal(result.wo);  // unconditional (PLTDiffuseBSDF.slang:122-127)
                return valid;
            case PLTPT_MATERIAL_CONDUCTOR: valid = conductor.sample(wi, result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobeMask); break;
            case PLTPT_MATERIAL_OPAQUE_DIELECTRIC: valid = opaque.sample(wi, result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobeMask); break;
            case PLTPT_MATERIAL_DIELECTRIC: case PLTPT_MATERIAL_THIN_DIELECTRIC: {
                const float d = dot(sd.V, woHint);
                const uint32_t lobes = (d > .0f ? uint32_t(L_Reflection) : d < .0f ? uint32_t(L_Transmission) : uint32_t(L_All)) & lobeMask;
                if (type == PLTPT_MATERIAL_DIELECTRIC) {
                    valid = dielectric.sample(wi, result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobes);
                    result.wo = sd.fromLocal(result.wo);  // unconditional (PLTDielectricBSDF.slang:307-316)
                    return valid;
                }
                valid = thin.sample(wi, result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobes);
                break;
            }
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC:
                valid = coated.sample(wi, result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobeMask); break;
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
                const ShadingData g = grated.grating.localGratingShadingFrame(sd);
                valid = grated.sample(g.toLocal(sd.V), result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobeMask, g.toLocal(woHint));
                if (valid) result.wo = g.fromLocal(result.wo);
                return valid;
            }
            case PLTPT_MATERIAL_MULTILAYERED_STACK: valid = stack.sample(wi, result.wo, result.pdf, result.weight, result.lobe, sg, wavelength, lobeMask); break;
            default: return false;
        }
        if (valid) result.wo = sd.fromLocal(result.wo);
        return valid;
    }
    // IMaterialInstance.pdf — note the argument order bxdf.pdf(toLocal(toLight), toLocal(V)) of the wrappers
    float pdf(const ShadingData& sd, float3 toLight, float wavelength) const {
        const float3 a = sd.toLocal(toLight), b = sd.toLocal(sd.V);
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE: return diffuse.pdf(a, b, wavelength);
            case PLTPT_MATERIAL_CONDUCTOR: return conductor.pdf(a, b, wavelength);
            case PLTPT_MATERIAL_DIELECTRIC: return dielectric.pdf(a, b, wavelength);
            case PLTPT_MATERIAL_OPAQUE_DIELECTRIC: return opaque.pdf(a, b, wavelength);
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: return coated.pdf(a, b, wavelength);
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
                const ShadingData g = grated.grating.localGratingShadingFrame(sd);
                return grated.pdf(g.toLocal(toLight), g.toLocal(sd.V), wavelength);
            }
            default: return .0f;
        }
    }
    // IMaterialInstance.lambdaPdf — Dielectric and the grated materials first move the (copied) beam to the local frame
    hwss_t lambdaPdf(const ShadingData& sd, float3 wo, uint32_t lobe, PLTBeam beam, hwss_t wavelengths) const {
        const float3 a = sd.toLocal(sd.V), b = sd.toLocal(wo);
        switch (type) {
            case PLTPT_MATERIAL_CONDUCTOR: return conductor.lambdaPdf(a, b, lobe, beam, wavelengths);
            case PLTPT_MATERIAL_DIELECTRIC:
                beam.r = sd.toLocal(beam.r); beam.t = sd.toLocal(beam.t);
                return dielectric.lambdaPdf(a, b, lobe, beam, wavelengths);
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: return coated.lambdaPdf(a, b, lobe, beam, wavelengths);
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC:
                beam.r = sd.toLocal(beam.r); beam.t = sd.toLocal(beam.t);
                return grated.lambdaPdf(a, b, lobe, beam, wavelengths);
            default: return hwss_t(1.f);
        }
    }
    float3kpdv partialDerivatives(const float3pdv& wi, const DifferentialShadingGeometry& dsg, uint32_t lobe, float wavelength) const {
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE: return float3kpdv();
            case PLTPT_MATERIAL_DIELECTRIC: return dielectric.partialDerivatives(wi, dsg.n, lobe, wavelength);
            case PLTPT_MATERIAL_THIN_DIELECTRIC: return thin.partialDerivatives(wi, dsg.n, lobe);
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
                const int2 l = GratedBxDF::decodeLobeType(lobe);
                if (l.x == 0 && l.y == 0) return reflectDerivatives(wi, dsg.n);
                return grated.grating.partialDerivatives(dsg, wi, l, wavelength);
            }
            default: return reflectDerivatives(wi, dsg.n);
        }
    }
    BSDFProperties getProperties() const {
        BSDFProperties p;
        const float kMean = k2Pi / (SpectrumUtils::meanWavelength * 1e-3f);
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE: {
                const bool isEmissive = emissionSpectralId != PLTPT_NO_SPECTRAL_PROFILE;
                p.emissionSpectralId = emissionSpectralId;
                p.flags = isEmissive ? BSDFProperties::IsEmissive : 0;
                p.emissionRGB = isEmissive ? emissionRGB : float3(.0f);
                p.albedo = diffuse.albedo; p.roughness = 1.f;
                break;
            }
            case PLTPT_MATERIAL_CONDUCTOR: {
                if (conductor.gfs.q() < QcutoffForMS) p.flags = BSDFProperties::CanManifoldWalkReflective;
                const float n = prof->evalSpectrum(conductor.etaN, SpectrumUtils::meanWavelength), k = prof->evalSpectrum(conductor.etaK, SpectrumUtils::meanWavelength);
                p.albedo = conductor.specular * evalFresnelConductor(float2(n, k), 1.f);
                p.roughness = conductor.gfs.roughness();
                break;
            }
            case PLTPT_MATERIAL_DIELECTRIC:
                p.flags = BSDFProperties::IsTransmissive | (dielectric.gfs.q < QcutoffForMS ? BSDFProperties::CanManifoldWalkTransmissive : 0);
                p.albedo = dielectric.specular * evalFresnelDielectric(dielectric.indexOfRefraction(kMean), 1.f);
                p.roughness = dielectric.gfs.roughness();
                break;
            case PLTPT_MATERIAL_OPAQUE_DIELECTRIC:
                if (opaque.gfs.q() < QcutoffForMS) p.flags = BSDFProperties::CanManifoldWalkReflective;
                p.albedo = opaque.albedo; p.roughness = opaque.gfs.roughness();
                break;
            case PLTPT_MATERIAL_THIN_DIELECTRIC:
                p.flags = BSDFProperties::IsTransmissive | BSDFProperties::CanManifoldWalkReflective | BSDFProperties::CanManifoldWalkTransmissive;
                p.albedo = float3(1.f); p.roughness = .0f;
                break;
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC:
                if (coated.gfs.q() < QcutoffForMS) p.flags = BSDFProperties::CanManifoldWalkReflective;
                p.albedo = coated.specular; p.roughness = coated.gfs.roughness();
                break;
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
                p.flags = BSDFProperties::CanManifoldWalkReflective;
                const float a = grated.grating.alpha(float3(0, 0, 1), kMean);
                if (grated.conductor) {
                    const float n = prof->evalSpectrum(grated.baseC.etaN, SpectrumUtils::meanWavelength), k = prof->evalSpectrum(grated.baseC.etaK, SpectrumUtils::meanWavelength);
                    p.albedo = grated.baseC.specular * evalFresnelConductor(float2(n, k), 1.f);
                } else p.albedo = grated.baseD.albedo;
                p.roughness = (1 - a) * baseGfs().roughness();
                break;
            }
            case PLTPT_MATERIAL_MULTILAYERED_STACK:
                p.flags = BSDFProperties::CanManifoldWalkReflective;
                p.albedo = stack.albedo; p.roughness = 0;
                break;
            default: break;
        }
        return p;
    }
    uint32_t getLobes() const {
        switch (type) {
            case PLTPT_MATERIAL_DIFFUSE: return L_DiffuseReflection;
            case PLTPT_MATERIAL_CONDUCTOR: return conductor.gfs.isDeltaOnly() ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection);
            case PLTPT_MATERIAL_DIELECTRIC: {
                const uint32_t dT = dielectric.isDispersive() ? L_DispersiveDeltaTransmission : L_DeltaTransmission;
                return dielectric.gfs.isDeltaOnly() ? (L_DeltaReflection | dT) : (L_DeltaReflection | dT | L_ScatterReflection | L_ScatterTransmission);
            }
            case PLTPT_MATERIAL_OPAQUE_DIELECTRIC: return opaque.gfs.isDeltaOnly() ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection);
            case PLTPT_MATERIAL_THIN_DIELECTRIC: return L_DeltaReflection | L_DeltaTransmission;
            case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC:
                return coated.gfs.isDeltaOnly() ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection);
            case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC:
                return (baseGfs().isDeltaOnly() ? L_DeltaReflection : (L_DeltaReflection | L_ScatterReflection)) | L_DispersiveDeltaReflection;
            case PLTPT_MATERIAL_MULTILAYERED_STACK: return L_DeltaReflection;
            default: return 0;
        }
    }
};

}  // namespace orc
