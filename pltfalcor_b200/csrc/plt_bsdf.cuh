// plt_bsdf.cuh — generalized ray (PLT beam), Mueller/Fresnel algebra, surface statistics and the ten PLT BSDFs
// for sm_100a.  B200-first choices:
//   * Fresnel Mueller matrices are kept in their 4-parameter block form {a,b,c,d} (never expanded to 4x4);
//   * material blobs are unpacked ONCE at set_scene into 128 B `DevMaterial` records (fp32, surface statistics
//     pre-evaluated) that are fetched with eight 16 B loads;
//   * spectral profiles are repacked so that both interpolation end points come from one 8 B load.
// Reference semantics are cited per function (paths relative to Source/Falcor/Rendering/...).
#pragma once
#include "plt_math.cuh"

namespace plt {

enum : uint32_t {
    L_DiffuseReflection = 0x01, L_ScatterReflection = 0x02, L_DeltaReflection = 0x04, L_DispersiveDeltaReflection = 0x08,
    L_ScatterTransmission = 0x20, L_DeltaTransmission = 0x40, L_DispersiveDeltaTransmission = 0x80,
    L_DispersiveDelta = 0x88, L_Delta = 0xCC, L_NonDelta = 0x33, L_Reflection = 0x0f, L_Transmission = 0xf0,
    L_NonDeltaReflection = 0x03, L_NonDeltaTransmission = 0x30, L_All = 0xffff, L_MaterialSpecific = 0xff00,
};
enum : uint32_t { P_IsTransmissive = 1, P_IsEmissive = 2, P_CanMSReflective = 4, P_CanMSTransmissive = 8 };  // BSDFProperties::Flags
constexpr float kQcutoffForMS = .15f;

// ---- device-side tables --------------------------------------------------------------------
struct alignas(16) DevProfile {   // repacked SpectralProfile (BasicMaterialData.slang:43-54)
    float intensity, binsF, minWavelength, maxWavelength;
    float2 pair[128];             // (pdf[i], pdf[min(i+1, bins-1)])
    float icdf[128];
};
struct alignas(16) DevMaterial {  // unpacked MaterialDataBlob, 128 B
    uint32_t type, hdrFlags, lobes, propFlags;          // hdrFlags bit0 = doubleSided
    float opacity, roughnessProp; uint32_t emissionId, etaN;
    uint32_t etaK; float extEta, c0x, c0y;
    float c0z, c1x, c1y, c1z;
    float albx, alby, albz, surfA;                      // surfA = FractalSurface.T or GaussianSurface.sigma2
    float surfGamma, surfQ, f0, f1;
    float f2, f3, f4, f5;
    float f6, f7; uint32_t u0, u1;
};
static_assert(sizeof(DevMaterial) == 128, "DevMaterial must stay one 128 B line");
static_assert(sizeof(DevProfile) == 1552, "DevProfile layout");

struct SD {  // the fields of Scene/ShadingData.slang the PLT path reads
    float3 posW, V, N, T, B, faceN;
    float2 uv;
    uint32_t materialID;
    bool frontFacing;
    PLT_D float3 fromLocal(float3 v) const { return T * v.x + B * v.y + N * v.z; }
    PLT_D float3 toLocal(float3 v) const { return mk3(dot(v, T), dot(v, B), dot(v, N)); }
    PLT_D float3 computeNewRayOrigin(bool viewside) const { return computeRayOrigin(posW, (frontFacing == viewside) ? faceN : -faceN); }
};

PLT_D DevMaterial loadMaterial(const DevMaterial* mats, uint32_t id) {
    DevMaterial m;
    const float4* src = reinterpret_cast<const float4*>(mats + id);
    float4* dst = reinterpret_cast<float4*>(&m);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = __ldg(src + i);
    return m;
}

// Scene.slang:251-294 (evalSpectrum)
PLT_D float evalSpectrum(const DevProfile* profiles, uint32_t id, float lambda) {
    if (id == PLTPT_NO_SPECTRAL_PROFILE) return 0.f;
    const DevProfile* sp = profiles + id;
    const float4 hdr = __ldg(reinterpret_cast<const float4*>(sp));
    const float x = (lambda - hdr.z) / (hdr.w - hdr.z) * (hdr.y - 1.0f);
    const float i = floorf(x);
    const float w = x - i;
    if (lambda < hdr.z || lambda > hdr.w) return 0.f;
    const int bins = int(hdr.y);
    const int i0 = max(0, min(bins - 1, int(i)));
    const float2 p = __ldg(&sp->pair[i0]);
    // i1 = clamp(int(i+1)) equals i0+1 clamped whenever lambda is inside the table
    return hdr.x * lerpf(p.x, p.y, w);
}
PLT_D float4 evalSpectrum4(const DevProfile* profiles, uint32_t id, float4 l) {
    return mk4(evalSpectrum(profiles, id, l.x), evalSpectrum(profiles, id, l.y), evalSpectrum(profiles, id, l.z), evalSpectrum(profiles, id, l.w));
}
// Scene.slang:296-327 (sampleSpectrum)
PLT_D float4 sampleSpectrum(const DevProfile* profiles, uint32_t id, float hero, Rng& sg, bool importanceSample, int hwss, float4& sampledWavelengths) {
    if (id == PLTPT_NO_SPECTRAL_PROFILE) return mk4(0.f);
    float4 ret = mk4(0.f);
    if (importanceSample) {
        const DevProfile* sp = profiles + id;
        const float4 hdr = __ldg(reinterpret_cast<const float4*>(sp));
        const uint32_t bins = uint32_t(hdr.y);
        sampledWavelengths.x = hero;
        ret.x = evalSpectrum(profiles, id, hero);
        const float3 x = sg.next3D() * hdr.y;
        for (int j = 1; j < hwss; ++j) {
            const float xj = comp(x, j - 1);
            const uint32_t idx = uint32_t(xj);
            const float c0 = idx < bins ? __ldg(&sp->icdf[idx]) : 1.f, c1 = idx + 1 < bins ? __ldg(&sp->icdf[idx + 1]) : 1.f;
            const float cdf = lerpf(c0, c1, xj - floorf(xj));
            setLane(sampledWavelengths, j, lerpf(hdr.z, hdr.w, cdf));
            setLane(ret, j, hdr.x / hdr.y);
        }
    } else {
        sampledWavelengths = hwssWavelengths(hero, hwss);
        ret = evalSpectrum4(profiles, id, sampledWavelengths);
    }
    return ret;
}

// ---- Fresnel: Materials/Fresnel.slang:30-90 -------------------------------------------------
PLT_D void fresnel_dielectric(float cosThetaI_, float eta, float& rs, float& rp, float& ts, float& tp, float& cosThetaT) {
    if (eta == 1) { cosThetaT = -cosThetaI_; rs = rp = 0; ts = tp = 1; return; }
    eta = (cosThetaI_ > 0) ? eta : 1 / eta;
    const float cosThetaI = fabsf(cosThetaI_);
    const float cosThetaTSqr = 1 - (1 - sqr(cosThetaI)) / sqr(eta);
    cosThetaT = sqrtf(fmaxf(.0f, cosThetaTSqr));
    rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    const float imp = sqrtf(eta * cosThetaT / cosThetaI);
    ts = imp * (rs + 1);
    tp = imp * (rp + 1) / eta;
    cosThetaT *= -signf(cosThetaI_);
}
PLT_D void fresnel_conductor(float cosThetaI, float2 eta, float2& rs, float2& rp) {
    if ((eta.x == 0 && eta.y == 0) || cosThetaI < .0f) { rs = rp = mk2(0.f, 0.f); return; }
    const float2 ie = cdiv(mk2(1.f, 0.f), eta);
    const float2 t2 = mk2(1.f, 0.f) - (1 - sqr(cosThetaI)) * cmul(ie, ie);
    const float2 t = csqrt(t2);
    const float2 i = mk2(cosThetaI, 0.f);
    const float2 et = cmul(eta, t), ei = cmul(eta, i);
    rs = cdiv(i - et, i + et);
    rp = cdiv(ei - t, ei + t);
}
PLT_D float evalFresnelDielectric(float eta, float cosThetaI, float& cosThetaT) {
    float rs, rp, ts, tp;
    fresnel_dielectric(cosThetaI, eta, rs, rp, ts, tp, cosThetaT);
    return .5f * (sqr(rs) + sqr(rp));
}
PLT_D float evalFresnelDielectric(float eta, float cosThetaI) { float c; return evalFresnelDielectric(eta, cosThetaI, c); }

// ---- Mueller algebra in block form: Materials/PLT/Mueller.slang:34-144 ----------------------
// {a,b,c,d} == [[a,b,0,0],[b,a,0,0],[0,0,c,d],[0,0,-d,c]]; closed under products and sums.
struct FM { float a, b, c, d; };
// Reference behaviour (Mueller.slang:76-77): `fs*cconj(fp)` there is the component-wise float2 product, not the complex one, so the
// lower block is c = Re fs Re fp, d = -Im fs Im fp (equal to Re / Im of fs conj(fp) only for real coefficients, i.e. dielectrics).
// Found by compiling the reference's own text for the CPU (oracle/ref_compare.cpp); reproduced as it is.
PLT_D FM fmFresnel(float2 fs, float2 fp) {
    const float Rp = cnorm2(fp), Rs = cnorm2(fs);
    return {(Rp + Rs) / 2.f, (Rp - Rs) / 2.f, fs.x * fp.x, fs.y * -fp.y};
}
PLT_D FM fmMul(FM A, FM B) { return {A.a * B.a + A.b * B.b, A.a * B.b + A.b * B.a, A.c * B.c - A.d * B.d, A.c * B.d + A.d * B.c}; }
PLT_D FM fmAdd(FM A, FM B) { return {A.a + B.a, A.b + B.b, A.c + B.c, A.d + B.d}; }
PLT_D float4 fmApply(FM M, float4 SP) { return mk4(M.a * SP.x + M.b * SP.y, M.b * SP.x + M.a * SP.y, M.c * SP.z + M.d * SP.w, -M.d * SP.z + M.c * SP.w); }
PLT_D FM fmRConductor(float cosThetaI, float2 eta, float2& rs, float2& rp) { fresnel_conductor(cosThetaI, eta, rs, rp); return fmFresnel(rs, rp); }
PLT_D FM fmRConductor(float cosThetaI, float2 eta) { float2 rs, rp; return fmRConductor(cosThetaI, eta, rs, rp); }
PLT_D FM fmDielectric(float cosThetaI, float eta, bool reflection) {
    float rs, rp, ts, tp, c;
    fresnel_dielectric(cosThetaI, eta, rs, rp, ts, tp, c);
    return reflection ? fmFresnel(mk2(rs, 0.f), mk2(rp, 0.f)) : fmFresnel(mk2(ts, 0.f), mk2(tp, 0.f));
}
// invOneMinusMuellerFresnel (Mueller.slang:89-110); false if either 2x2 minor is singular
PLT_D bool fmInvOneMinus(FM R, FM& out) {
    M22 im1, im2;
    if (!tryInvert(M22{1.f - R.a, 0.f - R.b, 0.f - R.b, 1.f - R.a}, im1) || !tryInvert(M22{1.f - R.c, 0.f - R.d, 0.f + R.d, 1.f - R.c}, im2)) return false;
    out = {im1.a, im1.b, im2.a, im2.b};
    return true;
}

// ---- generalized ray: Rendering/PLT.slang:41-290 ---------------------------------------------
struct Beam {
    float4 L, L1, L2, L3, wl, spdf;
    float Dxx, Dxy, Dyy, r_m;
    float3 posW, r, t;
    int comps;
    bool distant;

    PLT_D float4 getSP(int j) const { return mk4(lane(L, j), lane(L1, j), lane(L2, j), lane(L3, j)); }
    PLT_D void setSP(int j, float4 s) { setLane(L, j, s.x); setLane(L1, j, s.y); setLane(L2, j, s.z); setLane(L3, j, s.w); }
    PLT_D void setZero() { L = L1 = L2 = L3 = mk4(0.f); }
    PLT_D void scale(float4 s) { L *= s; L1 *= s; L2 *= s; L3 *= s; }
    PLT_D void scale(float s) { L *= s; L1 *= s; L2 *= s; L3 *= s; }
    PLT_D float rmm() const { return r_m * 1e+3f; }
    PLT_D M22 invThetaK(float k) const { const float s = sqr((k / k2Pi) / rmm()); return {s * Dxx, s * Dxy, s * Dxy, s * Dyy}; }
    PLT_D M22 invTheta() const { const float s = sqr(1.f / rmm()); return {s * Dxx, s * Dxy, s * Dxy, s * Dyy}; }
    PLT_D M22 Q() const { const float3 b = cross(t, r); return {t.x, b.x, t.y, b.y}; }
    PLT_D void transformCoherence(M22 U) {
        const M22 nD = mul(transpose(U), mul(M22{Dxx, Dxy, Dxy, Dyy}, U));
        Dxx = nD.a; Dxy = nD.b; Dyy = nD.d;
    }
    PLT_D float4 mutualCoherence(float4 k, float3 v) const {
        const M22 invT = invTheta();
        const float2 v2 = mk2(v.x, v.y);
        const float q = dot(v2, mul(invT, v2));
        const float4 rl = k / k2Pi;
        const float4 x = rl * rl * q;
        return mk4(expf(-.5f * x.x), expf(-.5f * x.y), expf(-.5f * x.z), expf(-.5f * x.w));
    }
    PLT_D float mutualCoherenceAngular(float3 d1, float3 d2) const {
        const float3 dd = vabs(d1 - d2);
        const float s4 = sqrtf(k4Pi);
        const float2 v = mk2(1.f / fmaxf(s4 * dd.x, kFltEps), 1.f / fmaxf(s4 * dd.y, kFltEps));
        const M22 iT = invTheta();
        const float r2 = sqr(rmm());
        const M22 invT = {iT.a * r2, iT.b * r2, iT.c * r2, iT.d * r2};
        const float x = 1.f / dot(v, mul(invT, v));
        return expf(-.5f * x);
    }
    // rotateFrame: MuellerT (Mueller.slang:34-53) acts on (L1,L2) only; coherence rotates by R(θ).  Split in two so that a beam with
    // several Stokes columns (Beam4 below) rotates all of them with one evaluation of the angle.
    PLT_D void frameRotation(float3 newt, float& rx, float& ry, float& sint, float& cost) const {
#if PLT_FAST_DEVICE
        // sin and cos of atan2(y, x) are y / h and x / h: no inverse trigonometry, same rotation (atan2(0, 0) = 0 and a non-finite angle -> identity)
        const float y = dot(r, cross(t, newt)), x = dot(t, newt), h2 = x * x + y * y;
        const bool ok = h2 > 0.f && h2 < kFltMax;
        const float ih = rsqrtf(h2);
        sint = ok ? y * ih : 0.f; cost = ok ? x * ih : 1.f;
        rx = 2 * sint * cost; ry = 1 - 2 * sqr(sint);
        return;
#endif
        float theta = atan2f(dot(r, cross(t, newt)), dot(t, newt));
        theta = isfinite(theta) ? theta : .0f;
        sincosf(theta, &sint, &cost);
        rx = 2 * sint * cost; ry = 1 - 2 * sqr(sint);
    }
    PLT_D void rotateStokes(float4& l1, float4& l2, float rx, float ry) const {
        PLT_FOR_LAMBDA(j, comps) {
            const float a = lane(l1, j), b = lane(l2, j);
            setLane(l1, j, rx * a + -ry * b);
            setLane(l2, j, ry * a + rx * b);
        }
    }
    PLT_D void rotateFrame(float3 newt, float3 newr) {
        float rx, ry, sint, cost;
        frameRotation(newt, rx, ry, sint, cost);
        rotateStokes(L1, L2, rx, ry);
        transformCoherence(M22{cost, -sint, sint, cost});
        r = newr;
        t = newt;
    }
    PLT_D static float3 localFrameTangent(float3 newwo) {
        const float3 newt = cross(mk3(0.f, 0.f, 1.f), newwo);
        const float l2 = dot(newt, newt);
        return l2 > 1e-8f ? newt / sqrtf(l2) : perp_stark(newwo);
    }
    PLT_D void constructLocalFrame(float3 newwo) { rotateFrame(localFrameTangent(newwo), newwo); }
    PLT_D void propagate(float3 p) {
        if (!distant) r_m += length(p - posW);
        posW = p;
    }
    PLT_D float3 toLocal(float3 v) const { return mk3(dot(v, t), dot(v, cross(t, r)), dot(v, r)); }
};
// A generalized ray that carries FOUR Stokes vectors per wavelength through the same geometry, coherence and wavelengths: the columns
// of the Mueller matrix when they start as the basis vectors e0..e3 (config 1 of the benchmark, pltpt_bsdf_eval_mueller).  Everything a
// BSDF computes that does not depend on the Stokes vector -- local frame, rotation angle, Fresnel terms, surface statistics, coherence --
// is then evaluated once instead of four times.  Column 0 lives in the base class; the BSDFs below are templates over the beam type and
// reach the Stokes data only through nCols / colL.. / getSPc / setSPc, so the renderer's instantiation (Beam, one column) is unchanged.
struct Beam4 : Beam {
    float4 xL[3], xL1[3], xL2[3], xL3[3];
    PLT_D void setZero() { Beam::setZero(); for (int c = 0; c < 3; ++c) xL[c] = xL1[c] = xL2[c] = xL3[c] = mk4(0.f); }
    PLT_D void scale(float4 s) { Beam::scale(s); for (int c = 0; c < 3; ++c) { xL[c] *= s; xL1[c] *= s; xL2[c] *= s; xL3[c] *= s; } }
    PLT_D void rotateFrame(float3 newt, float3 newr) {
        float rx, ry, sint, cost;
        frameRotation(newt, rx, ry, sint, cost);
        rotateStokes(L1, L2, rx, ry);
        for (int c = 0; c < 3; ++c) rotateStokes(xL1[c], xL2[c], rx, ry);
        transformCoherence(M22{cost, -sint, sint, cost});
        r = newr;
        t = newt;
    }
    PLT_D void constructLocalFrame(float3 newwo) { rotateFrame(localFrameTangent(newwo), newwo); }
};
template <class B> struct BeamCols { static constexpr int n = 1; };
template <> struct BeamCols<Beam4> { static constexpr int n = 4; };
PLT_D float4& colL(Beam& b, int) { return b.L; }
PLT_D float4& colL1(Beam& b, int) { return b.L1; }
PLT_D float4& colL2(Beam& b, int) { return b.L2; }
PLT_D float4& colL3(Beam& b, int) { return b.L3; }
PLT_D float4& colL(Beam4& b, int c) { return c == 0 ? b.L : b.xL[c - 1]; }
PLT_D float4& colL1(Beam4& b, int c) { return c == 0 ? b.L1 : b.xL1[c - 1]; }
PLT_D float4& colL2(Beam4& b, int c) { return c == 0 ? b.L2 : b.xL2[c - 1]; }
PLT_D float4& colL3(Beam4& b, int c) { return c == 0 ? b.L3 : b.xL3[c - 1]; }
template <class B> PLT_D float4 getSPc(B& b, int j, int c) { return mk4(lane(colL(b, c), j), lane(colL1(b, c), j), lane(colL2(b, c), j), lane(colL3(b, c), j)); }
template <class B> PLT_D void setSPc(B& b, int j, int c, float4 s) { setLane(colL(b, c), j, s.x); setLane(colL1(b, c), j, s.y); setLane(colL2(b, c), j, s.z); setLane(colL3(b, c), j, s.w); }
#define PLT_FOR_COLS(B, c) _Pragma("unroll") for (int c = 0; c < BeamCols<B>::n; ++c)

PLT_D Beam makeBeam(float4 Le, float diffusivity, float3 posW, float3 dir, float4 wl, bool distant, int comps) {
    Beam b;
    b.L = Le; b.L1 = b.L2 = b.L3 = mk4(0.f);
    b.wl = wl; b.spdf = mk4(1.f);
    b.Dxx = diffusivity; b.Dxy = 0.f; b.Dyy = diffusivity;
    b.r_m = distant ? 1e-3f : 0.f;
    b.posW = posW; b.r = dir; b.t = perp_stark(dir);
    b.comps = comps; b.distant = distant;
    return b;
}
PLT_D Beam sourceBeamDistant(float3 dir, float solidAngle, float4 Le, float4 wl, float maxBeamOmega, int comps) {
    return makeBeam(Le, fminf(solidAngle, maxBeamOmega), mk3(0.f), dir, wl, true, comps);
}
PLT_D Beam sourceBeamArea(float3 posW, float3 dir, float A, float dist, float4 Le, float4 wl, float maxBeamOmega, int comps) {
    const float r2 = sqr(dist * 1e+3f);
    return makeBeam(Le, fminf(A, maxBeamOmega * r2), posW, dir, wl, false, comps);
}

// ---- surfaces ------------------------------------------------------------------------------
// Delta-lobe energies `alpha = exp(-x)` feed `1 - alpha`, which cancels catastrophically for near-smooth surfaces: a 1-ulp
// difference in exp becomes an O(1) relative difference of the scattered lobe.  Parity therefore needs the correctly rounded
// fp32 value here (double-precision exp rounded once); every other exp in the path uses the fast fp32 expf.
PLT_D float exp_cr(float x) { return float(exp(double(x))); }
// FractalSurface.slang:38-135 with T, gamma, q pre-evaluated
struct Fractal {
    float T, gamma, q;
    PLT_D float alpha(float3 wi, float3 wo, float k) const { return exp_cr(-sqr((fabsf(wi.z) + fabsf(wo.z)) * q * k)); }
    PLT_D float4 alpha4(float3 wi, float3 wo, float4 k) const { return mk4(alpha(wi, wo, k.x), alpha(wi, wo, k.y), alpha(wi, wo, k.z), alpha(wi, wo, k.w)); }
    PLT_D bool isDeltaOnly() const { return q == .0f; }
#if PLT_FAST_DEVICE
    // x^-p as exp2(-p log2 x) on the MUFU units: x = 1 + T |z|^2 lies in [1, ~1e7], so the absolute error of lg2.approx (2^-22) times
    // p <= ~3 keeps the result within 1e-6 relative -- the IEEE powf is ~40 instructions, this is 4
    PLT_D float pGF(float2 z) const { return k1_2Pi * (gamma - 1) * T * exp2f(-.5f * (gamma + 1.f) * __log2f(1 + T * dot(z, z))); }
#else
    PLT_D float pGF(float2 z) const { return k1_2Pi * (gamma - 1) * T * (1.f / powf(1 + T * dot(z, z), (gamma + 1.f) / 2.f)); }
#endif
    PLT_D float pdf(float3 wi, float3 wo, float k) const { return pGF(mk2(k * (wi.x + wo.x), k * (wi.y + wo.y))); }
    PLT_D float eval(float k, float3 h) const { return pGF(mk2(k * h.x, k * h.y)); }
    PLT_D float3 sample(float3 wi, Rng& sg, float k, float& pdf) const {
        const float s = sqrtf(fmaxf(.0f, 1 - sqr(wi.z)));
        const float phi_i = s > 0 ? atan2f(wi.y, wi.x) : .0f;
        const float sqrtT = sqrtf(T);
        const float2 u1 = sg.next2D();
        sg.next2D();  // the reference draws a second pair it never uses (FractalSurface.slang:107)
        const float M = 1.f - powf(1 + sqr(k) * T * sqr(1 + s), -(gamma - 1) / 2.f);
        const float f = sqrtf(powf(1 - M * u1.x, -2.f / (gamma - 1)) - 1.f) / sqrtT;
        const float fk = f / k;
        const float phi_max = (f == .0f || s == .0f) ? kPi : acosf(clampf((sqr(fk) + sqr(s) - 1) / (2.f * fk * s), -1.f, 1.f));
        const float phi_f = phi_i + (2.f * u1.y - 1) * phi_max;
        const float2 zeta = mk2(f * cosf(phi_f), f * sinf(phi_f));
        const float2 wo = mk2(zeta.x / k - wi.x, zeta.y / k - wi.y);
        const float z = sqrtf(fmaxf(.0f, 1 - dot(wo, wo)));
        pdf = kPi / phi_max * z * pGF(zeta);
        return mk3(wo.x, wo.y, signf(wi.z) * z);
    }
};
// GaussianSurface.slang:37-110 with sigma2, q pre-evaluated
struct Gaussian {
    float sigma2, q;
    PLT_D bool isDeltaOnly() const { return q == .0f; }
    PLT_D float alpha(float3 wi, float3 wo, float k) const { return isDeltaOnly() ? 1.f : exp_cr(-sqr((fabsf(wi.z) + fabsf(wo.z)) * q * k)); }
    PLT_D float4 alpha4(float3 wi, float3 wo, float4 k) const { return mk4(alpha(wi, wo, k.x), alpha(wi, wo, k.y), alpha(wi, wo, k.z), alpha(wi, wo, k.w)); }
    PLT_D float3 sample(float3 wi, Rng& sg, float k, float& pdf) const {
        const float s2 = (1.f / sigma2) / sqr(k);
        const float2 wo = sample_boxmueller_truncated(sg.next2D(), mk2(-wi.x, -wi.y), s2, pdf);
        return mk3(wo.x, wo.y, signf(wi.z) * sqrtf(fmaxf(0.f, 1.f - dot(wo, wo))));
    }
    PLT_D float pdf(float3 wi, float3 wo, float k) const { return boxmueller_truncated_pdf(mk2(wo.x, wo.y), mk2(-wi.x, -wi.y), (1.f / sigma2) / sqr(k)); }
    // coherence-aware evaluation: the beam's inverse-coherence matrix (rotated by Q) is folded into the Gaussian
    PLT_D float eval(float k, M22 invTheta, M22 Q, float3 h) const {
        invTheta = mul(Q, mul(invTheta, transpose(Q)));
        const float is = 1.f / sigma2;
        const M22 S = {is + invTheta.a, 0.f + invTheta.b, 0.f + invTheta.c, is + invTheta.d};
        const float d = det(S);
        const M22 invS = {S.d / d, -S.b / d, -S.c / d, S.a / d};
        const float2 h2 = mk2(h.x, h.y);
        return expf(-.5f * sqr(k) * dot(h2, mul(invS, h2))) * (sqr(k) * k1_2Pi / sqrtf(d));
    }
};

PLT_D float3 reflectZ(float3 wi) { return mk3(-wi.x, -wi.y, wi.z); }
PLT_D M22 scatterCoherenceTransform(float3 woLocal, uint32_t lobe) {
    const float c = sqrtf(fmaxf(1e-2f, fabsf(woLocal.z)));
    return (lobe & L_Delta) != 0 ? M22{1.f, 0.f, 0.f, 1.f} : M22{c, .0f, .0f, 1.f / c};
}
PLT_D float3 refractLocal(float3 wi, float cosThetaT, float eta) {  // PLTDielectricBSDF.slang:68-71 / refractCoat
    const float scale = cosThetaT < 0 ? 1 / eta : eta;
    return mk3(-scale * wi.x, -scale * wi.y, cosThetaT);
}
PLT_D float3 scattered_wo(float3 wo, float eta) {  // PLTDielectricBSDF.slang:72-78
    const float scale = wo.z > 0 ? 1 / eta : eta;
    const float2 xy = mk2(wo.x * scale, wo.y * scale);
    const float l2 = dot(xy, xy);
    return l2 > 1 ? mk3(0.f) : mk3(xy.x, xy.y, -signf(wo.z) * sqrtf(fmaxf(0.f, 1 - l2)));
}
// Interference.slang:36-57
PLT_D float4 interfere(float4 SP1, float4 SP2, float phasor_s, float phasor_p, float coh) {
    const float S1 = (SP1.x + SP1.y) / 2.f, P1 = (SP1.x - SP1.y) / 2.f, S2 = (SP2.x + SP2.y) / 2.f, P2 = (SP2.x - SP2.y) / 2.f;
    const float S = S1 + S2 + 2 * sqrtf(fmaxf(.0f, S1 * S2)) * phasor_s * coh;
    const float P = P1 + P2 + 2 * sqrtf(fmaxf(.0f, P1 * P2)) * phasor_p * coh;
    const float n = SP1.x + SP2.x > 1e-6f ? (S + P) / (SP1.x + SP2.x) : .0f;
    return mk4(S + P, S - P, (SP1.z + SP2.z) * n, (SP1.w + SP2.w) * n);
}

struct BsdfCtx {  // read-only tables every BSDF call may touch
    const DevProfile* profiles;
    int hwss;
};
struct Sample { float3 wo; float pdf; uint32_t lobe; float weight; };

PLT_D float3 matC0(const DevMaterial& m) { return mk3(m.c0x, m.c0y, m.c0z); }
PLT_D float3 matC1(const DevMaterial& m) { return mk3(m.c1x, m.c1y, m.c1z); }
PLT_D Fractal matFractal(const DevMaterial& m) { return {m.surfA, m.surfGamma, m.surfQ}; }
PLT_D Gaussian matGaussian(const DevMaterial& m) { return {m.surfA, m.surfQ}; }
constexpr float kMeanK = k2Pi / (kMeanWavelength * 1e-3f);

// =========================================================================== Diffuse (PLTDiffuseBSDF.slang:39-95)
template <class B>
PLT_D void diffuseEval(const DevMaterial& m, float3 toLight, uint32_t lobe, B& beam) {
    const bool valid = (lobe & L_DiffuseReflection) != 0;
    const float4 w = valid ? (fabsf(toLight.z) * k1_Pi) * rgbToSpectrum4(matC0(m), beam.wl) : mk4(0.f);
    PLT_FOR_COLS(B, c) {
        colL(beam, c) *= w;
        colL1(beam, c) = colL2(beam, c) = colL3(beam, c) = mk4(0.f);
    }
}
// =========================================================================== Conductor (PLTConductorBSDF.slang:44-170)
template <class B>
PLT_D void conductorEval(const DevMaterial& m, const BsdfCtx& cx, float3 toLight, float3 toEye, uint32_t lobe, B& beam) {
    const bool isScatter = (lobe & L_ScatterReflection) != 0, isDirect = (lobe & L_DeltaReflection) != 0;
    if ((!isScatter && !isDirect) || (lobe & L_Reflection) == 0) { beam.setZero(); return; }
    const Fractal gfs = matFractal(m);
    const float4 k = waveNumbers(beam.wl);
    const float4 a = gfs.alpha4(toLight, toEye, k);
    const float4 m00 = rgbToSpectrum4(matC0(m), beam.wl);
    const float3 h = toLight + toEye;
    const float LdotH = fmaxf(.0f, dot(toLight, normalize(h)));
    const float4 iorN = evalSpectrum4(cx.profiles, m.etaN, beam.wl), iorK = evalSpectrum4(cx.profiles, m.etaK, beam.wl);
    PLT_FOR_LAMBDA(j, beam.comps) {
        const float2 ior = mk2(lane(iorN, j) / m.extEta, lane(iorK, j) / m.extEta);
        const FM M = fmRConductor(isScatter ? LdotH : toLight.z, ior);
        const float w = isScatter ? (1 - lane(a, j)) * fabsf(toLight.z) * gfs.eval(lane(k, j), h) : lane(a, j);
        PLT_FOR_COLS(B, c) setSPc(beam, j, c, (w * fmApply(M, getSPc(beam, j, c))) * lane(m00, j));
    }
}
PLT_D bool fractalBaseSample(const DevMaterial& m, float kSample, bool perWavelengthK, float3 wi, Sample& s, Rng& sg, float wavelength, uint32_t lobeMask) {
    // shared by Conductor (k from wavelength) and OpaqueDielectric (k = mean k): PLTConductorBSDF.slang:101-138
    const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
    if (!hasReflection || (!hasDirect && !hasScatter)) return false;
    (void)perWavelengthK;
    const Fractal gfs = matFractal(m);
    const float a = gfs.alpha(wi, wi, kSample);
    s.pdf = 1.f;
    bool isDirect = hasDirect;
    if (hasDirect && hasScatter) {
        isDirect = a == 1.f || sg.next1D() < a;
        s.pdf = isDirect ? a : 1.f - a;
    }
    s.weight = rgbToSpectrum(matC0(m), wavelength);
    if (isDirect) { s.wo = reflectZ(wi); s.lobe = L_DeltaReflection; }
    else {
        float surfPdf;
        s.wo = gfs.sample(wi, sg, kSample, surfPdf);
        s.pdf *= surfPdf;
        s.lobe = L_ScatterReflection;
    }
    return true;
}
PLT_D float fractalBasePdf(const DevMaterial& m, float k, float3 wi, float3 wo) {
    const Fractal gfs = matFractal(m);
    return (1 - gfs.alpha(wi, wi, k)) * gfs.pdf(wi, wo, k);
}
PLT_D float4 conductorLambdaPdf(const DevMaterial& m, const BsdfCtx& cx, float3 toLight, float3 toEye, uint32_t lobe, float4 wl) {
    if ((lobe & L_NonDelta) == 0) return mk4(1.f);
    const Fractal gfs = matFractal(m);
    const float4 k = waveNumbers(wl);
    float4 ret = mk4(0.f);
    PLT_FOR_LAMBDA(j, cx.hwss) setLane(ret, j, gfs.pdf(toEye, toLight, lane(k, j)));
    return ret;
}
// =========================================================================== OpaqueDielectric (PLTOpaqueDielectricBSDF.slang:43-147)
template <class B>
PLT_D void opaqueEval(const DevMaterial& m, float3 toLight, float3 toEye, uint32_t lobe, B& beam) {
    const bool isDirect = (lobe & L_Delta) != 0, isScatter = (lobe & L_NonDelta) != 0;
    if ((!isScatter && !isDirect) || (lobe & L_Reflection) == 0) { beam.setZero(); return; }
    const Fractal gfs = matFractal(m);
    const float4 m00 = rgbToSpectrum4(matC0(m), beam.wl);
    const float a = gfs.alpha(toLight, toEye, kMeanK);
    float M = a;
    if (isScatter) M = (1 - a) * fabsf(toLight.z) * gfs.eval(kMeanK, toLight + toEye);
    PLT_FOR_LAMBDA(j, beam.comps) PLT_FOR_COLS(B, c) setSPc(beam, j, c, (M * lane(m00, j)) * getSPc(beam, j, c));
}
// =========================================================================== Dielectric (PLTDielectricBSDF.slang:43-270)
PLT_D float dielectricIor(const DevMaterial& m, float k) { return (m.f0 + m.f2 * sqr(k * k1_2Pi)) / m.f1; }
template <class B>
PLT_D void dielectricEval(const DevMaterial& m, const BsdfCtx& cx, float3 toLight, float3 toEye, uint32_t lobe, B& beam) {
    const float zz = toLight.z * toEye.z;
    const bool isReflection = zz > .0f;
    const bool isScatter = (lobe & L_NonDelta) != 0, isDirect = (lobe & L_Delta) != 0;
    const bool hasTransmission = (lobe & L_Transmission) != 0, hasReflection = (lobe & L_Reflection) != 0;
    if (zz == 0 || (!isScatter && !isDirect) || (!hasTransmission && !isReflection) || (!hasReflection && isReflection)) { beam.setZero(); return; }
    const Gaussian gfs = matGaussian(m);
    const float4 k = waveNumbers(beam.wl);
    const float4 m00 = rgbToSpectrum4(isReflection ? matC0(m) : matC1(m), beam.wl);
    const float4 eta = mk4(dielectricIor(m, k.x), dielectricIor(m, k.y), dielectricIor(m, k.z), dielectricIor(m, k.w));
    const float3 toEyeUp = isReflection ? toEye : scattered_wo(toEye, eta.x);
    const float3 toLightBeam = beam.toLocal(toLight);
    const float4 a = gfs.alpha4(toLight, toEyeUp, k);
    if (isScatter) {
        const float3 h = toLight + toEyeUp;
        const float LdotH = fmaxf(0.f, dot(toLight, normalize(h)));
        const M22 Q = beam.Q();
        const M22 invTheta = beam.invThetaK(1.f);
        PLT_FOR_LAMBDA(j, beam.comps) {
            const float e = lane(eta, j);
            const float J = isReflection ? 1.f : sqr(toEye.z > .0f ? e : 1.f / e);
            const float D = gfs.eval(lane(k, j), invTheta, Q, h);
            const FM M1 = fmDielectric(signf(toLight.z) * LdotH, e, isReflection);
            const float w = lane(m00, j) * (1 - lane(a, j)) * J * fabsf(toLight.z) * D;
            PLT_FOR_COLS(B, c) setSPc(beam, j, c, w * fmApply(M1, getSPc(beam, j, c)));
        }
    } else {
        PLT_FOR_LAMBDA(j, beam.comps) {
            const float e = lane(eta, j);
            const float J = isReflection ? 1.f : sqr(toEye.z > .0f ? e : 1.f / e);
            const FM M0 = fmDielectric(toLight.z, e, isReflection);
            float coh = 1.f;
            if (!isReflection && j != 0) {  // component 0 is the sampled lobe (PLTDielectricBSDF.slang:131-132)
                float cosThetaT;
                evalFresnelDielectric(e, toEye.z, cosThetaT);
                coh = beam.mutualCoherenceAngular(beam.toLocal(refractLocal(toEye, cosThetaT, e)), toLightBeam);
            }
            const float w = lane(m00, j) * lane(a, j) * coh * J;
            PLT_FOR_COLS(B, c) setSPc(beam, j, c, w * fmApply(M0, getSPc(beam, j, c)));
        }
    }
}
PLT_D bool dielectricSample(const DevMaterial& m, float3 wi, Sample& s, Rng& sg, float wavelength, uint32_t lobeMask) {
    const Gaussian gfs = matGaussian(m);
    const float k = k2Pi / (1e-3f * wavelength);
    const float a = gfs.alpha(wi, wi, k);
    const float eta = dielectricIor(m, k);
    const bool hasReflection = (lobeMask & L_Reflection) != 0, hasTransmission = (lobeMask & L_Transmission) != 0;
    const bool hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
    if ((!hasReflection && !hasTransmission) || (!hasDirect && !hasScatter)) return false;
    const float pdfDirect = hasDirect && hasScatter ? a : hasDirect ? 1.f : .0f;
    bool isReflection;
    const bool isDirect = pdfDirect == 1.f || sg.next1D() < pdfDirect;
    s.pdf = (hasDirect && hasScatter) ? (isDirect ? pdfDirect : 1.f - pdfDirect) : 1.f;
    if (isDirect) {
        float cosThetaT;
        const float Fd = evalFresnelDielectric(eta, wi.z, cosThetaT);
        isReflection = hasReflection ? (Fd > 0 && sg.next1D() <= Fd) : false;
        s.wo = isReflection ? reflectZ(wi) : refractLocal(wi, cosThetaT, eta);
        s.pdf *= (hasReflection && hasTransmission) ? (isReflection ? Fd : 1.f - Fd) : 1.f;
        s.lobe = isReflection ? L_DeltaReflection : (m.f2 != 0.f ? L_DispersiveDeltaTransmission : L_DeltaTransmission);
    } else {
        float surfPdf;
        s.wo = gfs.sample(wi, sg, k, surfPdf);
        s.pdf *= surfPdf;
        const float F = evalFresnelDielectric(eta, signf(wi.z) * dot(wi, normalize(wi + s.wo)));
        isReflection = hasReflection ? (F > 0 && sg.next1D() <= F) : false;
        if (!isReflection) s.wo = scattered_wo(s.wo, eta);
        s.pdf *= (hasReflection && hasTransmission) ? (isReflection ? F : 1.f - F) : 1.f;
        s.lobe = isReflection ? L_ScatterReflection : L_ScatterTransmission;
    }
    s.weight = rgbToSpectrum(isReflection ? matC0(m) : matC1(m), wavelength);
    const float zz = wi.z * s.wo.z;
    return (isReflection && zz <= 0.0f) ? false : (!isReflection && zz >= 0.0f) ? false : true;
}
PLT_D float dielectricPdf(const DevMaterial& m, float3 wi, float3 woLocal, float wavelength) {
    const Gaussian gfs = matGaussian(m);
    const float k = k2Pi / (1e-3f * wavelength);
    const float a = gfs.alpha(wi, wi, k);
    const float eta = dielectricIor(m, k);
    const bool isReflection = wi.z * woLocal.z > 0.0f;
    const float3 wo = isReflection ? woLocal : scattered_wo(woLocal, eta);
    const float F = evalFresnelDielectric(eta, signf(wi.z) * dot(wi, normalize(wi + wo)));
    return (1.f - a) * (isReflection ? F : 1.f - F) * gfs.pdf(wi, wo, k);
}
PLT_D float4 dielectricLambdaPdf(const DevMaterial& m, const BsdfCtx& cx, float3 toLight, float3 toEye, uint32_t lobe, const Beam& beam, float4 wl) {
    const bool isReflection = toLight.z * toEye.z > .0f;
    const bool isScatter = (lobe & L_NonDelta) != 0, isDirect = (lobe & L_Delta) != 0;
    const Gaussian gfs = matGaussian(m);
    const float4 k = waveNumbers(wl);
    if (isDirect && !isReflection) {
        const float3 toLightBeam = beam.toLocal(toLight);
        float4 ret = mk4(0.f);
        PLT_FOR_LAMBDA(j, cx.hwss) {
            if (j == 0) { ret.x = 1.f; continue; }
            const float e = dielectricIor(m, lane(k, j));
            float cosThetaT;
            evalFresnelDielectric(e, toEye.z, cosThetaT);
            setLane(ret, j, beam.mutualCoherenceAngular(beam.toLocal(refractLocal(toEye, cosThetaT, e)), toLightBeam));
        }
        return ret;
    }
    if (isScatter) {
        float4 ret = mk4(0.f);
        PLT_FOR_LAMBDA(j, cx.hwss) {
            const float3 wo = isReflection ? toLight : scattered_wo(toLight, dielectricIor(m, lane(k, j)));
            setLane(ret, j, gfs.pdf(toEye, wo, lane(k, j)));
        }
        return ret;
    }
    return mk4(1.f);
}
// =========================================================================== ThinDielectric (PLTThinDielectricBSDF.slang:43-200, Birefringence.slang)
namespace biref {
PLT_D float qo(float eo, float K) { return sqrtf(sqr(eo) - sqr(K)); }
PLT_D float2 qe(float eo, float ee, float3 A, float K) {
    const float e = sqr(ee) - sqr(eo);
    const float G = sqr(A.y) * e + sqr(eo);
    const float root = sqrtf(fmaxf(0.f, sqr(K) * (sqr(A.x) * e - sqr(ee)) + sqr(ee) * G));
    const float t = A.z * A.y * K * e;
    return mk2((eo * root + t) / G, (-eo * root + t) / G);
}
PLT_D float3 E_e(float eo, float3 A, float K, float q) {
    return normalize(mk3(A.x * sqr(eo), K * A.z * q + A.y * (sqr(eo) - sqr(q)), K * A.y * q + A.z * (sqr(eo) - sqr(K))));
}
PLT_D float3 E_o(float3 A, float K, float q) { return normalize(mk3(-A.y * K - A.z * q, A.x * K, A.x * q)); }
PLT_D float3 Poynting(float q, float3 E, float K) {
    return normalize(mk3(E.x * (K * E.z - q * E.y), K * E.z * E.y + q * (sqr(E.x) + sqr(E.z)), -q * E.z * E.y - K * (sqr(E.x) + sqr(E.y))));
}
}  // namespace biref
template <class B>
PLT_D void thinBirefringence(const DevMaterial& m, const BsdfCtx& cx, B& beam, float3 wi, float4 k) {
    using namespace biref;
    const float etaAvg = m.f0, tau = m.f1, bir = m.f2;
    const float2 oaxy = mk2(m.f3, m.f4);
    const float3 OA = mk3(oaxy.x, oaxy.y, sqrtf(fmaxf(.0f, 1.f - dot(oaxy, oaxy))));
    const float phii = atan2f(wi.y, wi.x);
    const float phiA = atan2f(OA.y, OA.x) - phii;
    const float3 I = mk3(0.f, fabsf(wi.z), sqrtf(1.f - sqr(wi.z)));
    float3 A = mk3(cosf(phiA), 0.f, sinf(phiA)) * sqrtf(1 - sqr(OA.z));
    A.y = signf(wi.z) * OA.z;
    const float3 A2 = mk3(A.x, -A.y, A.z);
    const float3 Z = beam.toLocal(mk3(cosf(phii), sinf(phii), 0.f));
    const float ei = 1.f, eo = etaAvg - .5f * bir, ee = etaAvg + .5f * bir;
    const float tauum = tau * 1e+3f;
    // vectors_in_slab (Birefringence.slang:168-196)
    const float Ks = ei * I.z;
    const float qom_s = -qo(eo, Ks);
    const float2 qe_s = qe(eo, ee, A, Ks);
    const float3 Io = Poynting(qom_s, E_o(A, Ks, qom_s), Ks);
    const float3 Ie = Poynting(qe_s.y, E_e(eo, A, Ks, qe_s.y), Ks);
    const float3 We = normalize(mk3(0.f, qe_s.y, -Ks));
    const float e_eff = Ks / sqrtf(fmaxf(.0f, 1.f - sqr(We.y)));
    // fresnel_iso_aniso(I.y, I.z, ..., A)  (Birefringence.slang:107-133)
    float tso, tse, tpo, tpe;
    {
        const float cos_t = I.y, sin_t = I.z;
        const float K = ei * sin_t;
        const float qi = -qo(ei, K);
        const float q2om = -qo(eo, K);
        const float q2em = qe(eo, ee, A, K).y;
        const float3 E2om = E_o(A, K, q2om), E2em = E_e(eo, A, K, q2em);
        const float A_1 = -E2em.x * E2om.y * K * cos_t, A_2 = -E2em.y * E2om.x * K * cos_t;
        const float C_1 = qi * cos_t - K * sin_t;
        const float C_2 = E2em.x * E2om.z * (C_1 + q2om * cos_t) - A_1;
        const float C_3 = E2em.z * E2om.x * (C_1 + q2em * cos_t) - A_2;
        const float Ns = C_2 * (qi + q2em) - C_3 * (qi + q2om);
        tpo = -(2 * E2em.x * cos_t * (q2em + qi) * C_1) / Ns;
        tpe = (2 * E2om.x * cos_t * (q2om + qi) * C_1) / Ns;
        tso = -(2 * qi * (E2em.y * K * cos_t + E2em.z * C_1 + E2em.z * q2em * cos_t)) / Ns;
        tse = (2 * qi * (E2om.y * K * cos_t + E2om.z * C_1 + E2om.z * q2om * cos_t)) / Ns;
    }
    // fresnel_aniso_iso(I.y, I.z, ..., A2)  (Birefringence.slang:134-165)
    float tos, top, tes, tep;
    {
        const float cos_p = I.y, sin_p = I.z;
        const float K = ei * sin_p;
        const float q2om = -qo(ei, K);
        const float qop = qo(eo, K);
        const float2 qq = qe(eo, ee, A2, K);
        const float qep = qq.x, qem = qq.y;
        const float3 Eop = E_o(A2, K, qop), Eep = E_e(eo, A2, K, qep), Eom = E_o(A2, K, -qop), Eem = E_e(eo, A2, K, qem);
        const float B_1 = -Eep.x * K * cos_p * (q2om - qep), B_2 = -Eom.x * K * cos_p * (q2om + qop), B_3 = -Eop.x * K * cos_p * (qop - q2om);
        const float D_1 = K * sin_p - q2om * cos_p;
        const float D_2 = Eep.z * (D_1 + qep * cos_p), D_3 = Eem.z * (D_1 + qem * cos_p), D_4 = Eop.z * (D_1 + qop * cos_p);
        const float N2 = (Eep.x * (qep - q2om) * D_4 + Eop.x * (q2om - qop) * D_2 + Eop.y * B_1 + Eep.y * B_3);
        const float roo = -(Eep.x * Eom.z * (q2om - qep) * (qop * cos_p - D_1) + Eom.x * (q2om + qop) * D_2 + Eom.y * B_1 - Eep.y * B_2) / N2;
        const float roe = -(-Eom.x * Eop.z * (qop * (qop * cos_p + K * sin_p) + q2om * D_1) + Eom.z * Eop.x * (qop * (qop * cos_p - K * sin_p) + q2om * D_1) + Eop.y * B_2 + Eom.y * B_3) / N2;
        tos = (Eep.x * qep * roe - Eom.x * qop + Eop.x * qop * roo) / q2om;
        top = -(Eom.z * qop - Eom.y * K - (Eep.y * K + Eep.z * qep) * roe - (Eop.y * K + Eop.z * qop) * roo) / D_1;
        const float reo = (Eem.x * (qem - q2om) * (D_2 + Eep.y * K * cos_p) + Eep.x * (q2om - qep) * (D_3 + Eem.y * K * cos_p)) / N2;
        const float ree = -(Eem.x * (qem - q2om) * D_4 - Eem.x * Eop.y * K * cos_p * (q2om - qem) + Eem.y * B_3 + Eop.x * (q2om - qop) * D_3) / N2;
        tes = (Eem.x * qem + Eep.x * qep * ree + Eop.x * qop * reo) / q2om;
        tep = -(Eem.z + Eep.z * ree + Eop.z * reo) / cos_p;
    }
    const float ss = tso * tos + tse * tes, sp = tso * top + tse * tep, ps = tpo * tos + tpe * tes, pp = tpo * top + tpe * tep;
    const float Ioz = fabsf(tauum / Io.y) * Io.z, Iez = fabsf(tauum / Ie.y) * Ie.z;
    const float OPLo = fabsf(tauum / Io.y) * eo, OPLe = fabsf(tauum / Ie.y) * e_eff;
    const float oez = Ioz - Iez;
    const float4 mutual_coh = beam.mutualCoherence(k, oez * Z);
    const float ph = ei * oez * I.z + OPLo - OPLe;
    const float4 t = mutual_coh * 2.f * mk4(sinf(-k.x * ph), sinf(-k.y * ph), sinf(-k.z * ph), sinf(-k.w * ph));
    PLT_FOR_COLS(B, c) {  // the Stokes update is not linear (sqrt of the product of the two linear intensities): column by column
        float4 &cL = colL(beam, c), &cL1 = colL1(beam, c);
        const float4 Lx = .5f * (cL + cL1), Ly = .5f * (cL - cL1);
        const float4 pr = Lx * Ly;
        const float4 sq = mk4(sqrtf(pr.x), sqrtf(pr.y), sqrtf(pr.z), sqrtf(pr.w));
        float4 nLx = sqr(ss) * Lx + sqr(ps) * Ly + (ss * ps) * t * sq;
        float4 nLy = sqr(sp) * Lx + sqr(pp) * Ly + (sp * pp) * t * sq;
        nLx = mk4(fmaxf(0.f, nLx.x), fmaxf(0.f, nLx.y), fmaxf(0.f, nLx.z), fmaxf(0.f, nLx.w));
        nLy = mk4(fmaxf(0.f, nLy.x), fmaxf(0.f, nLy.y), fmaxf(0.f, nLy.z), fmaxf(0.f, nLy.w));
        const float4 newL = nLx + nLy;
        bool allOk = true, allFinite = true;
        PLT_FOR_LAMBDA(j, cx.hwss) {
            const bool f = isfinite(lane(newL, j));
            allFinite = allFinite && f;
            allOk = allOk && (lane(cL, j) > 1e-6f && f);
        }
        const float4 norm = allOk ? newL / cL : mk4(0.f);
        cL = allFinite ? newL : mk4(0.f);
        cL1 = allFinite ? nLx - nLy : mk4(0.f);
        colL2(beam, c) *= norm;
        colL3(beam, c) *= norm;
    }
}
template <class B>
PLT_D void thinEval(const DevMaterial& m, const BsdfCtx& cx, float3 toLight, float3 toEye, uint32_t lobe, B& beam) {
    const float eta = m.f0 / 1.f;
    const float zz = toLight.z * toEye.z;
    const bool isReflection = zz > 0.0f;
    const bool isDirect = (lobe & L_Delta) != 0, hasTransmission = (lobe & L_Transmission) != 0, hasReflection = (lobe & L_Reflection) != 0;
    if ((isReflection && eta == 1.0f) || zz == 0 || !isDirect || (!hasTransmission && !isReflection) || (!hasReflection && isReflection)) { beam.setZero(); return; }
    const float4 k = waveNumbers(beam.wl);
    const float ci = fabsf(toLight.z);
    const FM R = fmDielectric(ci, eta, true), T = fmDielectric(ci, eta, false);
    FM iR;
    const bool inv = fmInvOneMinus(R, iR) && iR.a > 0;  // zero matrix on failure => iR[0][0] > 0 is false
    const FM M = isReflection ? (inv ? fmAdd(R, fmMul(fmMul(T, R), fmMul(iR, T))) : R) : (inv ? fmMul(T, fmMul(iR, T)) : fmMul(T, T));
    if (m.f2 > .0f && !isReflection) thinBirefringence(m, cx, beam, toLight, k);
    PLT_FOR_LAMBDA(j, beam.comps) PLT_FOR_COLS(B, c) setSPc(beam, j, c, fmApply(M, getSPc(beam, j, c)));
}
PLT_D bool thinSample(const DevMaterial& m, float3 wi, Sample& s, Rng& sg, uint32_t lobeMask) {
    const bool hasReflection = (lobeMask & L_Reflection) != 0, hasTransmission = (lobeMask & L_Transmission) != 0;
    if (!hasReflection && !hasTransmission) return false;
    const float Fd = evalFresnelDielectric(m.f0 / 1.f, fabsf(wi.z));
    const bool isReflection = hasReflection ? (Fd > 0 && sg.next1D() <= Fd) : false;
    s.wo = isReflection ? reflectZ(wi) : -wi;
    s.weight = 1.f;
    s.pdf = hasReflection && hasTransmission ? (isReflection ? Fd : 1.f - Fd) : 1.f;
    s.lobe = isReflection ? L_DeltaReflection : L_DeltaTransmission;
    return true;
}
// =========================================================================== Coated (PLTCoatedConductorBSDF.slang:45-245, PLTCoatedOpaqueDielectricBSDF.slang:44-244)
// f0 = coatIOR, f1 = tau, f2 = coatEta() (pre-evaluated), u0 = 1 for the conductor base
template <class B>
PLT_D void coatedEval(const DevMaterial& m, const BsdfCtx& cx, float3 toLight, float3 toEye, uint32_t lobe, B& beam) {
    const bool conductor = m.u0 != 0;
    const bool isReflection = toLight.z * toEye.z > 0.0f;
    const bool isScatter = (lobe & L_ScatterReflection) != 0, isDirect = (lobe & L_DeltaReflection) != 0;
    const float coatior = m.f2, tau = m.f1;
    const float eta1 = coatior / m.extEta;
    float cosThetaTin;
    evalFresnelDielectric(eta1, toLight.z, cosThetaTin);
    const float3 wiIn = -refractLocal(toLight, cosThetaTin, eta1);
    float3 woIn = reflectZ(wiIn);
    if (!isDirect) {
        float cosThetaTout;
        evalFresnelDielectric(eta1, toEye.z, cosThetaTout);
        woIn = -refractLocal(toEye, cosThetaTout, eta1);
    }
    if (!isReflection || (!isScatter && !isDirect) || wiIn.z <= .0f || woIn.z <= .0f) { beam.setZero(); return; }
    const Fractal gfs = matFractal(m);
    const float4 k = waveNumbers(beam.wl);
    const float4 a = gfs.alpha4(wiIn, woIn, k);
    const float4 m00 = rgbToSpectrum4(matC0(m), beam.wl);
    const float3 h = wiIn + woIn;
    const float LdotH = fmaxf(0.f, dot(wiIn, normalize(h)));
    const float3 Z = beam.toLocal(mk3(0.f, 0.f, 1.f));
    const float4 mutual_coh = beam.mutualCoherence(k, tau * Z);
    const FM MiT = fmDielectric(toLight.z, eta1, false), MoT = fmDielectric(-woIn.z, eta1, false);
    const float dist = 2.f * tau / wiIn.z;
    if (conductor) {
        const float4 iorN = evalSpectrum4(cx.profiles, m.etaN, beam.wl), iorK = evalSpectrum4(cx.profiles, m.etaK, beam.wl);
        const FM MiR = fmDielectric(toLight.z, eta1, true);
        PLT_FOR_LAMBDA(j, beam.comps) {
            float2 rs, rp;
            const FM M2R = fmRConductor(LdotH, mk2(lane(iorN, j) / coatior, lane(iorK, j) / coatior), rs, rp);
            const FM M = fmMul(fmMul(MiT, M2R), MoT);
            const float ws = isScatter ? (1 - lane(a, j)) * fabsf(toLight.z) * gfs.eval(lane(k, j), h) : 0.f;
            const float pS = isScatter ? 0.f : sinf(dist * lane(k, j) * coatior + atan2f(rs.y, rs.x));
            const float pP = isScatter ? 0.f : sinf(dist * lane(k, j) * coatior + atan2f(rp.y, rp.x));
            PLT_FOR_COLS(B, c) {
                const float4 SP = getSPc(beam, j, c);
                float4 nSP = lane(m00, j) * fmApply(M, SP);
                if (isScatter) nSP *= ws;
                else nSP = interfere(lane(a, j) * nSP, fmApply(MiR, SP), pS, pP, lane(mutual_coh, j));
                setSPc(beam, j, c, nSP);
            }
        }
    } else {
        const FM M = fmMul(MiT, MoT);
        PLT_FOR_LAMBDA(j, beam.comps) {
            const float phase = sinf(dist * lane(k, j) * coatior);
            const float w = isScatter ? (1 - lane(a, j)) * fabsf(toLight.z) * gfs.eval(lane(k, j), h) : lane(a, j);
            const FM Mr = fmDielectric(isScatter ? LdotH : toLight.z, eta1, true);
            PLT_FOR_COLS(B, c) {
                const float4 SP = getSPc(beam, j, c);
                float4 nSP = lane(m00, j) * fmApply(M, SP);
                nSP *= w;
                const float4 SPrefl = w * fmApply(Mr, SP);
                setSPc(beam, j, c, interfere(nSP, SPrefl, phase, phase, lane(mutual_coh, j)));
            }
        }
    }
}
PLT_D bool coatedSample(const DevMaterial& m, float3 wi, Sample& s, Rng& sg, float wavelength, uint32_t lobeMask) {
    const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
    if (!hasReflection || (!hasDirect && !hasScatter) || wi.z <= .0f) return false;
    const float k = k2Pi / (1e-3f * wavelength);
    const float eta1 = m.f2 / m.extEta;
    float cosThetaT;
    const float Fd = evalFresnelDielectric(eta1, wi.z, cosThetaT);
    const bool isReflCoat = Fd == .1f || sg.next1D() < Fd;  // sic, PLTCoatedConductorBSDF.slang:151
    s.weight = 1.f;
    s.pdf = isReflCoat ? Fd : 1.f - Fd;
    if (isReflCoat) { s.wo = reflectZ(wi); s.lobe = L_DeltaReflection; return true; }
    const float3 wiIn = -refractLocal(wi, cosThetaT, eta1);
    if (wiIn.z == .0f) return false;
    s.weight = rgbToSpectrum(matC0(m), wavelength);
    const Fractal gfs = matFractal(m);
    const float a = gfs.alpha(wiIn, wiIn, k);
    const float pdfDirect = hasDirect && hasScatter ? a : hasDirect ? 1.f : .0f;
    const bool isDirect = pdfDirect == 1.f || sg.next1D() < pdfDirect;
    if (hasDirect && hasScatter) s.pdf *= isDirect ? pdfDirect : 1.f - pdfDirect;
    if (isDirect) { s.wo = reflectZ(wi); s.lobe = L_DeltaReflection; }
    else {
        float surfPdf;
        const float3 woIn = gfs.sample(wiIn, sg, k, surfPdf);
        s.pdf *= surfPdf;
        float cosThetaT1;
        evalFresnelDielectric(eta1, -woIn.z, cosThetaT1);
        s.wo = refractLocal(-woIn, cosThetaT1, eta1);
        s.lobe = L_ScatterReflection;
    }
    return wi.z * s.wo.z > .0f;
}
PLT_D void coatedInner(const DevMaterial& m, float3 a, float3 b, float3& aIn, float3& bIn, float& Fd) {
    const float eta1 = m.f2 / m.extEta;
    float c0, c1;
    Fd = evalFresnelDielectric(eta1, a.z, c0);
    evalFresnelDielectric(eta1, b.z, c1);
    aIn = -refractLocal(a, c0, eta1);
    bIn = -refractLocal(b, c1, eta1);
}
PLT_D float coatedPdf(const DevMaterial& m, float3 wi, float3 wo, float wavelength) {
    if (wi.z <= .0f || wo.z <= .0f) return .0f;
    const float k = k2Pi / (1e-3f * wavelength);
    float3 wiIn, woIn;
    float Fd;
    coatedInner(m, wi, wo, wiIn, woIn, Fd);
    const Fractal gfs = matFractal(m);
    return wiIn.z * woIn.z == .0f ? .0f : (1.f - gfs.alpha(wiIn, wiIn, k)) * (1.f - Fd) * gfs.pdf(wiIn, woIn, k);
}
PLT_D float4 coatedLambdaPdf(const DevMaterial& m, const BsdfCtx& cx, float3 toLight, float3 toEye, uint32_t lobe, float4 wl) {
    if ((lobe & L_NonDelta) == 0) return mk4(1.f);
    float3 wiIn, woIn;
    float Fd;
    coatedInner(m, toEye, toLight, wiIn, woIn, Fd);
    const Fractal gfs = matFractal(m);
    const float4 k = waveNumbers(wl);
    float4 ret = mk4(0.f);
    PLT_FOR_LAMBDA(j, cx.hwss) setLane(ret, j, gfs.pdf(wiIn, woIn, lane(k, j)));
    return ret;
}
// =========================================================================== Diffraction grating (DiffractionGrating.slang:40-266)
struct Grating {
    float2 gx, invPitch;
    float q, mult;
    int lobeCount;
    uint32_t type;
    PLT_D bool is1D() const { return invPitch.y == .0f; }
    PLT_D float alpha(float3 wi, float k) const { return exp_cr(-sqr(wi.z * q * k)); }
    PLT_D float4 alpha4(float3 wi, float4 k) const { return mk4(alpha(wi, k.x), alpha(wi, k.y), alpha(wi, k.z), alpha(wi, k.w)); }
    PLT_D float intensity1(int l, float a) const {
        const uint32_t gt = type & 0xF;
        if (l == 0) return 1.f;
        if (gt == 0) return BesselJ(l, a);
        if (gt == 1) { const float x = kPi * l / 2.f; return sinf(a / 2.f) * (sinf(x) / x); }
        return 1.f / sqrtf(fabsf(float(l)));
    }
    PLT_D float lobeIntensity(int lx, int ly, float3 wi, float lambda) const {
        const float a = k4Pi * q / (lambda * fabsf(wi.z));
        const float ix = intensity1(lx, a);
        const float iy = is1D() ? ix : intensity1(ly, a);
        return mult * ix * iy;
    }
    PLT_D float3 diffract(float3 wi, int lx, int ly, float lambda) const {
        const float px = sqrtf(sqr(wi.x) + sqr(wi.z)), py = sqrtf(sqr(wi.y) + sqr(wi.z));
        const float six = px > kFltEps ? wi.x / px : .0f, siy = py > kFltEps ? wi.y / py : .0f;
        const float a = lambda * float(lx) * invPitch.x - six, b = lambda * float(ly) * invPitch.y - siy;
        const float mm = (sqr(a) - 1) / (sqr(a * b) - 1);
        const float qq = 1 - sqr(b) * mm;
        return (fabsf(a) <= 1.f && fabsf(b) <= 1.f) ? mk3(a * sqrtf(fmaxf(.0f, qq)), b * sqrtf(mm), sqrtf(fmaxf(.0f, 1 - sqr(a) * qq - sqr(b) * mm))) : mk3(0.f);
    }
    PLT_D void closestLobeHint(float3 wi, float3 woHint, float alpha, float lambda, int& lx, int& ly) const {
        lx = ly = 0;
        if (alpha == 1.f) return;
        const float w[2] = {wi.x, wi.y}, h[2] = {woHint.x, woHint.y}, ip[2] = {invPitch.x, invPitch.y};
        int lo[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float li2 = sqr(w[i]) + sqr(wi.z), lo2 = sqr(h[i]) + sqr(woHint.z);
            const float sini = li2 > kFltEps ? w[i] / sqrtf(li2) : .0f, sino = lo2 > kFltEps ? -h[i] / sqrtf(lo2) : .0f;
            const float l = (sini - sino) / (ip[i] * lambda);
            lo[i] = min(max((int)(l + signf(l) * .5f), -lobeCount), lobeCount);
            lo[i] = (alpha == .0f && lo[i] == 0) ? int(signf(l)) : lo[i];
        }
        lx = lo[0];
        ly = is1D() ? 0 : lo[1];
    }
    PLT_D void sampleLobe(float3 wi, float2& pdf, float lambda, Rng& sg, int& lx, int& ly) const {
        float inten[8];
        float total = .0f;
        for (int l = 0; l < lobeCount; ++l) { inten[l] = lobeIntensity(l + 1, 0, wi, lambda); total += inten[l]; }
        const float2 u = sg.next2D();
        const float rx = (u.x - .5f) * 2.f, ry = (u.y - .5f) * 2.f;
        float cdf = .0f;
        lx = ly = 0;
        pdf = mk2(0.f, 0.f);
        for (int l = 0; l < lobeCount; ++l) {
            const float p = inten[l] / total;
            cdf += p;
            if (fabsf(rx) < cdf && lx == 0) { pdf.x = p; lx = l + 1; }
            if (fabsf(ry) < cdf && ly == 0) { pdf.y = p; ly = l + 1; }
        }
        pdf = pdf / 2.f;
        lx = rx < .0f ? -lx : lx;
        ly = ry < .0f ? -ly : ly;
        if (is1D()) { ly = 0; pdf.y = 1.f; }
    }
};
// f0 dir, f1/f2 invPitch, f3 q, f4 intensity mult, f5/f6 cos/sin(dir), u0 lobes, u1 type
PLT_D Grating matGrating(const DevMaterial& m, float2 uv) {
    Grating g;
    g.gx = mk2(m.f5, m.f6);
    if ((m.u1 & 0x10) != 0) {  // radial (DiffractionGrating.slang:57-61)
        float2 rad = mk2(uv.x - .5f, uv.y - .5f);
        rad = rad / sqrtf(dot(rad, rad));
        g.gx = mk2(m.f5 * rad.x + m.f6 * -rad.y, -m.f6 * rad.x + m.f5 * -rad.y);
    }
    g.invPitch = mk2(m.f1, m.f2); g.q = m.f3; g.mult = m.f4; g.lobeCount = int(m.u0); g.type = m.u1;
    return g;
}
PLT_D void encodeLobeType(int lx, int ly, uint32_t& bsdfLobe) {
    const int sx = lx < 0 ? 0 : 1, sy = ly < 0 ? 0 : 1;
    lx = min(abs(lx), 7); ly = min(abs(ly), 7);
    bsdfLobe &= ~uint32_t(L_MaterialSpecific);
    bsdfLobe += uint32_t((ly << 3) + lx + (sy << 7) + (sx << 6)) << 8;
}
PLT_D void decodeLobeType(uint32_t bsdfLobe, int& lx, int& ly) {
    const int x = int(bsdfLobe) >> 8;
    lx = (x & 0x7) * (2 * ((x >> 6) & 0x1) - 1);
    ly = ((x >> 3) & 0x7) * (2 * ((x >> 7) & 0x1) - 1);
}
// =========================================================================== Grated (PLTDiffractionGrated{Conductor,OpaqueDielectric}BSDF.slang)
template <class B>
PLT_D void gratedEval(const DevMaterial& m, const BsdfCtx& cx, const Grating& g, float3 toLight, float3 toEye, uint32_t bsdfLobe, B& beam) {
    const bool conductor = m.type == PLTPT_MATERIAL_GRATED_CONDUCTOR;
    const bool isDirect = (bsdfLobe & L_Delta) != 0;
    const float4 lambdas = 1e-3f * beam.wl;
    const float4 k = k2Pi / lambdas;
    const float4 a = g.alpha4(toLight, k);
    int lx, ly;
    decodeLobeType(bsdfLobe, lx, ly);
    const float3 toLightBeam = beam.toLocal(toLight);
    bool anyPos = false, anyLt1 = false;
    PLT_FOR_LAMBDA(j, cx.hwss) { anyPos = anyPos || lane(a, j) > .0f; anyLt1 = anyLt1 || lane(a, j) < 1.f; }
    const bool lobeZero = lx == 0 && ly == 0;
    if (lobeZero && anyPos) {
        if (conductor) conductorEval(m, cx, toLight, toEye, bsdfLobe, beam); else opaqueEval(m, toLight, toEye, bsdfLobe, beam);
        beam.scale(a);
    } else if (!lobeZero && anyLt1 && isDirect) {
        const float4 m00 = rgbToSpectrum4(matC0(m), beam.wl);
        float4 iorN = mk4(0.f), iorK = mk4(0.f);
        if (conductor) { iorN = evalSpectrum4(cx.profiles, m.etaN, beam.wl); iorK = evalSpectrum4(cx.profiles, m.etaK, beam.wl); }
        PLT_FOR_LAMBDA(j, beam.comps) {
            const float lam = lane(lambdas, j);
            const float coh = j == 0 ? 1.f : beam.mutualCoherenceAngular(beam.toLocal(g.diffract(toEye, lx, ly, lam)), toLightBeam);
            const float w = lane(m00, j) * (1 - lane(a, j)) * coh * g.lobeIntensity(lx, ly, toLight, lam);
            FM M = {1.f, 0.f, 1.f, 0.f};
            if (conductor) M = fmRConductor(toLight.z, mk2(lane(iorN, j) / m.extEta, lane(iorK, j) / m.extEta));
            PLT_FOR_COLS(B, c) {
                float4 SP = getSPc(beam, j, c);
                if (conductor) SP = fmApply(M, SP);
                setSPc(beam, j, c, w * SP);
            }
        }
    } else beam.setZero();
}
PLT_D bool gratedSample(const DevMaterial& m, const Grating& g, float3 wi, Sample& s, Rng& sg, float wavelength, uint32_t lobeMask, float3 woHint) {
    const bool conductor = m.type == PLTPT_MATERIAL_GRATED_CONDUCTOR;
    const bool hasReflection = (lobeMask & L_Reflection) != 0, hasDirect = (lobeMask & L_Delta) != 0, hasScatter = (lobeMask & L_NonDelta) != 0;
    if (!hasReflection || (!hasDirect && !hasScatter)) return false;
    const float lambdas = 1e-3f * wavelength;
    const float k = k2Pi / lambdas;
    const float a = g.alpha(wi, k);
    const bool hasHint = hasDirect && (woHint.x != .0f || woHint.y != .0f || woHint.z != .0f);
    int lx = 0, ly = 0;
    if (hasHint && a < 1.f) g.closestLobeHint(wi, woHint, a, lambdas, lx, ly);
    bool isDirect = hasDirect;
    s.pdf = 1.f;
    if (hasHint) isDirect = lx == 0 && ly == 0;
    else if (hasDirect && hasScatter) {
        isDirect = a == 1.f || sg.next1D() < a;
        s.pdf = isDirect ? a : 1.f - a;
    }
    if (isDirect) {
        const float pdf0 = s.pdf;
        fractalBaseSample(m, conductor ? k : kMeanK, conductor, wi, s, sg, wavelength, lobeMask);
        s.pdf *= pdf0;
        return true;
    }
    s.weight = rgbToSpectrum(matC0(m), wavelength);
    float2 lobePdf = mk2(1.f, 1.f);
    if (!hasHint) g.sampleLobe(wi, lobePdf, lambdas, sg, lx, ly);
    s.wo = g.diffract(wi, lx, ly, lambdas);
    s.lobe = L_DispersiveDeltaReflection;
    encodeLobeType(lx, ly, s.lobe);
    s.pdf *= lobePdf.x * lobePdf.y;
    return true;
}
PLT_D float4 gratedLambdaPdf(const DevMaterial& m, const BsdfCtx& cx, const Grating& g, float3 toLight, float3 toEye, uint32_t bsdfLobe, const Beam& beam, float4 wl) {
    const bool isDirect = (bsdfLobe & L_Delta) != 0;
    const float4 lambdas = 1e-3f * wl;
    const float4 a = g.alpha4(toLight, k2Pi / lambdas);
    int lx, ly;
    decodeLobeType(bsdfLobe, lx, ly);
    bool anyLt1 = false;
    PLT_FOR_LAMBDA(j, cx.hwss) anyLt1 = anyLt1 || lane(a, j) < 1.f;
    if ((lx != 0 || ly != 0) && anyLt1 && isDirect) {
        const float3 toLightBeam = beam.toLocal(toLight);
        float4 ret = mk4(0.f);
        PLT_FOR_LAMBDA(j, cx.hwss)
            setLane(ret, j, j == 0 ? 1.f : beam.mutualCoherenceAngular(beam.toLocal(g.diffract(toEye, lx, ly, lane(lambdas, j))), toLightBeam));
        return ret;
    }
    return mk4(1.f);
}
// =========================================================================== MultiLayeredStack (PLTMultiLayeredStackBSDF.slang:42-140)
// f0/f1 = ior, f2/f3 = thickness, u0 = layers
template <class B>
PLT_D void stackEval(const DevMaterial& m, float3 toLight, float3 toEye, uint32_t lobe, B& beam) {
    if ((lobe & L_Delta) == 0 || !(toLight.z * toEye.z > .0f)) { beam.setZero(); return; }
    const float4 k = waveNumbers(beam.wl);
    const float4 m00 = rgbToSpectrum4(matC0(m), beam.wl);
    const float eta1 = m.f0 / m.extEta, eta2 = m.f1 / m.f0;
    float ct0, ct1, r0, r1;
    const float Rout = sqrtf(evalFresnelDielectric(eta1, toLight.z, ct0));
    r0 = sqrtf(evalFresnelDielectric(eta2, -ct0, ct1));
    r1 = sqrtf(evalFresnelDielectric(eta2, ct1));
    ct0 = fabsf(ct0); ct1 = fabsf(ct1);
    const uint32_t layers = m.u0;
    PLT_FOR_LAMBDA(j, beam.comps) {
        float2 G = mk2(sqrtf(lane(m00, j)), 0.f);
        for (uint32_t l = 0; l < layers; ++l) {
            const bool odd = (l & 1) != 0;
            const float d = 2.f * (odd ? m.f3 : m.f2) / (odd ? ct1 : ct0) * lane(k, j) * (odd ? m.f1 : m.f0);
            // interference sum: near anti-reflection it cancels to ~1e-4 of its terms, so the phase factor is the correctly
            // rounded fp32 cos/sin (evaluated in double), as in the oracle
            double sdd, cdd;
            sincos(double(d), &sdd, &cdd);
            const float sd = float(sdd), cd = float(cdd);
            const float R = l == layers - 1 ? Rout : (odd ? r1 : r0);
            G = cmul(G, mk2(cd, sd));
            G = cdiv(mk2(R + G.x, G.y), mk2(1.f + R * G.x, R * G.y));
        }
        PLT_FOR_COLS(B, c) setSPc(beam, j, c, sqr(G.x) * getSPc(beam, j, c));
    }
}

// =========================================================================== IMaterialInstance dispatch
PLT_D bool isGrated(uint32_t t) { return t == PLTPT_MATERIAL_GRATED_CONDUCTOR || t == PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC; }
struct LocalFrame { float3 T, B, N; PLT_D float3 toLocal(float3 v) const { return mk3(dot(v, T), dot(v, B), dot(v, N)); } PLT_D float3 fromLocal(float3 v) const { return T * v.x + B * v.y + N * v.z; } };
PLT_D LocalFrame gratingFrame(const SD& sd, const Grating& g) {  // DiffractionGrating.slang:84-91
    LocalFrame f;
    f.T = sd.fromLocal(mk3(g.gx.x, g.gx.y, .0f));
    f.N = sd.N;
    f.B = cross(f.N, f.T);
    return f;
}

// IMaterialInstance.eval: local-frame prologue/epilogue (e.g. PLTConductorBSDF.slang:179-196)
template <class B>
PLT_D void bsdfEval(const DevMaterial& m, const BsdfCtx& cx, const SD& sd, float3 toEye, uint32_t lobe, B& beam) {
    LocalFrame fr = {sd.T, sd.B, sd.N};
    Grating g;
    const bool grated = isGrated(m.type);
    if (grated) { g = matGrating(m, sd.uv); fr = gratingFrame(sd, g); }
    beam.r = fr.toLocal(beam.r);
    beam.t = fr.toLocal(beam.t);
    const float3 wi = fr.toLocal(sd.V), wo = fr.toLocal(toEye);
    beam.constructLocalFrame(wo);
    M22 U = scatterCoherenceTransform(wo, lobe);
    switch (m.type) {
        case PLTPT_MATERIAL_DIFFUSE: diffuseEval(m, wi, lobe, beam); { const float c = sqrtf(fmaxf(1e-2f, fabsf(wo.z))); U = {c, .0f, .0f, 1.f / c}; } break;
        case PLTPT_MATERIAL_CONDUCTOR: conductorEval(m, cx, wi, wo, lobe, beam); break;
        case PLTPT_MATERIAL_DIELECTRIC: dielectricEval(m, cx, wi, wo, lobe, beam); break;
        case PLTPT_MATERIAL_OPAQUE_DIELECTRIC: opaqueEval(m, wi, wo, lobe, beam); break;
        case PLTPT_MATERIAL_THIN_DIELECTRIC: thinEval(m, cx, wi, wo, lobe, beam); U = {1.f, 0.f, 0.f, 1.f}; break;
        case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: coatedEval(m, cx, wi, wo, lobe, beam); break;
        case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
            gratedEval(m, cx, g, wi, wo, lobe, beam);
            int lx, ly;
            decodeLobeType(lobe, lx, ly);
            if (lx != 0 || ly != 0) U = {1.f, 0.f, 0.f, 1.f};
            break;
        }
        case PLTPT_MATERIAL_MULTILAYERED_STACK: stackEval(m, wi, wo, lobe, beam); U = {1.f, 0.f, 0.f, 1.f}; break;
        default: beam.setZero();
    }
    beam.transformCoherence(U);
    beam.r = fr.fromLocal(beam.r);
    beam.t = fr.fromLocal(beam.t);
}

// IMaterialInstance.sample
PLT_D bool bsdfSample(const DevMaterial& m, const SD& sd, Rng& sg, Sample& s, float wavelength, uint32_t lobeMask, float3 woHint) {
    s.wo = mk3(0.f); s.pdf = 0.f; s.lobe = 0; s.weight = 0.f;
    const float3 wi = sd.toLocal(sd.V);
    bool valid = false;
    switch (m.type) {
        case PLTPT_MATERIAL_DIFFUSE:
            s.weight = rgbToSpectrum(matC0(m), wavelength);
            s.wo = sd.fromLocal(sample_cosine_hemisphere_concentric(sg.next2D(), s.pdf));
            s.lobe = L_DiffuseReflection;
            return (lobeMask & L_DiffuseReflection) != 0;
        case PLTPT_MATERIAL_CONDUCTOR: valid = fractalBaseSample(m, k2Pi / (1e-3f * wavelength), true, wi, s, sg, wavelength, lobeMask); break;
        case PLTPT_MATERIAL_OPAQUE_DIELECTRIC: valid = fractalBaseSample(m, kMeanK, false, wi, s, sg, wavelength, lobeMask); break;
        case PLTPT_MATERIAL_DIELECTRIC: case PLTPT_MATERIAL_THIN_DIELECTRIC: {
            const float d = dot(sd.V, woHint);
            const uint32_t lobes = (d > .0f ? uint32_t(L_Reflection) : d < .0f ? uint32_t(L_Transmission) : uint32_t(L_All)) & lobeMask;
            if (m.type == PLTPT_MATERIAL_DIELECTRIC) {
                valid = dielectricSample(m, wi, s, sg, wavelength, lobes);
                s.wo = sd.fromLocal(s.wo);
                return valid;
            }
            valid = thinSample(m, wi, s, sg, lobes);
            break;
        }
        case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: valid = coatedSample(m, wi, s, sg, wavelength, lobeMask); break;
        case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
            const Grating g = matGrating(m, sd.uv);
            const LocalFrame fr = gratingFrame(sd, g);
            valid = gratedSample(m, g, fr.toLocal(sd.V), s, sg, wavelength, lobeMask, fr.toLocal(woHint));
            if (valid) s.wo = fr.fromLocal(s.wo);
            return valid;
        }
        case PLTPT_MATERIAL_MULTILAYERED_STACK:
            if ((lobeMask & L_Reflection) == 0 || (lobeMask & L_Delta) == 0 || wi.z <= .0f) return false;
            s.pdf = 1.f; s.wo = reflectZ(wi); s.lobe = L_DeltaReflection; s.weight = rgbToSpectrum(matC0(m), wavelength);
            valid = wi.z * s.wo.z > .0f;
            break;
        default: return false;
    }
    if (valid) s.wo = sd.fromLocal(s.wo);
    return valid;
}

// IMaterialInstance.pdf(sd, toLight, wavelength): the wrappers call bxdf.pdf(toLocal(toLight), toLocal(sd.V), ...)
PLT_D float bsdfPdf(const DevMaterial& m, const SD& sd, float3 toLight, float wavelength) {
    float3 a = sd.toLocal(toLight), b = sd.toLocal(sd.V);
    switch (m.type) {
        case PLTPT_MATERIAL_DIFFUSE: return b.z * k1_Pi;
        case PLTPT_MATERIAL_CONDUCTOR: return fractalBasePdf(m, k2Pi / (1e-3f * wavelength), a, b);
        case PLTPT_MATERIAL_OPAQUE_DIELECTRIC: return fractalBasePdf(m, kMeanK, a, b);
        case PLTPT_MATERIAL_DIELECTRIC: return dielectricPdf(m, a, b, wavelength);
        case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: return coatedPdf(m, a, b, wavelength);
        case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC: {
            const Grating g = matGrating(m, sd.uv);
            const LocalFrame fr = gratingFrame(sd, g);
            a = fr.toLocal(toLight); b = fr.toLocal(sd.V);
            const float k = k2Pi / (1e-3f * wavelength);
            return fractalBasePdf(m, m.type == PLTPT_MATERIAL_GRATED_CONDUCTOR ? k : kMeanK, a, b) * g.alpha(a, k);
        }
        default: return 0.f;
    }
}

// IMaterialInstance.lambdaPdf
PLT_D float4 bsdfLambdaPdf(const DevMaterial& m, const BsdfCtx& cx, const SD& sd, float3 wo, uint32_t lobe, Beam beam, float4 wl) {
    const float3 a = sd.toLocal(sd.V), b = sd.toLocal(wo);
    switch (m.type) {
        case PLTPT_MATERIAL_CONDUCTOR: return conductorLambdaPdf(m, cx, a, b, lobe, wl);
        case PLTPT_MATERIAL_DIELECTRIC:
            beam.r = sd.toLocal(beam.r); beam.t = sd.toLocal(beam.t);
            return dielectricLambdaPdf(m, cx, a, b, lobe, beam, wl);
        case PLTPT_MATERIAL_COATED_CONDUCTOR: case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC: return coatedLambdaPdf(m, cx, a, b, lobe, wl);
        case PLTPT_MATERIAL_GRATED_CONDUCTOR: case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC:
            beam.r = sd.toLocal(beam.r); beam.t = sd.toLocal(beam.t);
            return gratedLambdaPdf(m, cx, matGrating(m, sd.uv), a, b, lobe, beam, wl);
        default: return mk4(1.f);
    }
}

}  // namespace plt
