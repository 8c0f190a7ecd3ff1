// ORACLE — TEST INFRASTRUCTURE ONLY (see plt_math.h header).  Pinned to the reference's shader text (Fresnel.slang, Mueller.slang,
// PLT.slang, Interference.slang: oracle/ref_compare.cpp, tests/test_oracle_ref.py).
// plt_beam.h — Fresnel, Mueller matrices and the generalized ray ("PLT beam").
#pragma once
#include "plt_spectrum.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// Rendering/Materials/Fresnel.slang:30-90
// ---------------------------------------------------------------------------------------------
inline void fresnel_dielectric(float cosThetaI_, float eta, float& rs, float& rp, float& ts, float& tp, float& cosThetaT) {
    if (eta == 1) {
        cosThetaT = -cosThetaI_;
        rs = rp = 0;
        ts = tp = 1;
        return;
    }
    eta = (cosThetaI_ > 0) ? eta : 1 / eta;
    const float cosThetaI = std::fabs(cosThetaI_);
    const float cosThetaTSqr = 1 - (1 - sqr(cosThetaI)) / sqr(eta);
    cosThetaT = std::sqrt(hlsl_max(.0f, cosThetaTSqr));
    rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    const float imp = std::sqrt(eta * cosThetaT / cosThetaI);
    ts = imp * (rs + 1);
    tp = imp * (rp + 1) / eta;
    cosThetaT *= -signf(cosThetaI_);
}
inline void fresnel_conductor(float cosThetaI, float2 eta, float2& rs, float2& rp) {
    if ((eta.x == 0 && eta.y == 0) || cosThetaI < .0f) {
        rs = rp = float2(.0f);
        return;
    }
    const float2 t2 = c_t(1) - (1 - sqr(cosThetaI)) * csqr(cdiv(c_t(1), eta));
    const float2 t = csqrt(t2);
    const float2 i = c_t(cosThetaI);
    rs = cdiv((i - cmul(eta, t)), (i + cmul(eta, t)));
    rp = cdiv((cmul(eta, i) - t), (cmul(eta, i) + t));
}
inline float evalFresnelDielectric(float eta, float cosThetaI, float& cosThetaT) {
    float rs, rp, ts, tp;
    fresnel_dielectric(cosThetaI, eta, rs, rp, ts, tp, cosThetaT);
    return .5f * (sqr(rs) + sqr(rp));
}
inline float evalFresnelDielectric(float eta, float cosThetaI) {
    float c;
    return evalFresnelDielectric(eta, cosThetaI, c);
}
inline float evalFresnelConductor(float2 eta, float cosThetaI) {
    float2 rs, rp;
    fresnel_conductor(cosThetaI, eta, rs, rp);
    return .5f * (cnorm2(rs) + cnorm2(rp));
}

// ---------------------------------------------------------------------------------------------
// Rendering/Materials/PLT/Mueller.slang:34-144
// ---------------------------------------------------------------------------------------------
inline float4x4 MuellerT(float3 n1, float3 t1, float3 /*n2*/, float3 t2, float& outsint, float& outcost) {
    float theta = std::atan2(dot(n1, cross(t1, t2)), dot(t1, t2));
    theta = std::isfinite(theta) ? theta : .0f;
    const float sint = std::sin(theta), cost = std::cos(theta);
    const float rx = 2 * sint * cost;
    const float ry = 1 - 2 * sqr(sint);
    outsint = sint;
    outcost = cost;
    float4x4 T;
    T.m[0][0] = T.m[3][3] = 1;
    T.m[1][1] = T.m[2][2] = rx;   // sic: the reference puts sin(2θ) on the diagonal (Mueller.slang:48-51)
    T.m[1][2] = -ry;
    T.m[2][1] = ry;
    return T;
}
inline float4x4 MuellerPolarizer(float theta) {
    float s = std::sin(theta), c = std::cos(theta);
    float4x4 P;
    P.m[0][0] = 1;
    P.m[0][1] = P.m[1][0] = c;
    P.m[0][2] = P.m[2][0] = s;
    P.m[1][2] = P.m[2][1] = c * s;
    P.m[1][1] = c * c;
    P.m[2][2] = s * s;
    P.m[3][3] = 0;
    return .5f * P;
}
inline float4x4 make4x4(float a00, float a01, float a02, float a03, float a10, float a11, float a12, float a13,
                        float a20, float a21, float a22, float a23, float a30, float a31, float a32, float a33) {
    float4x4 M;
    float v[16] = {a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33};
    for (int i = 0; i < 16; ++i) M.m[i / 4][i % 4] = v[i];
    return M;
}
inline float4x4 mul(const float4x4& A, const float4x4& B) {
    float4x4 R;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = A.m[i][0] * B.m[0][j];
            for (int k = 1; k < 4; ++k) s += A.m[i][k] * B.m[k][j];
            R.m[i][j] = s;
        }
    return R;
}
inline float4x4 MuellerFresnel(float2 fs, float2 fp) {
    const float Rp = cnorm2(fp), Rs = cnorm2(fs);
    const float m00 = (Rp + Rs) / 2.f, m01 = (Rp - Rs) / 2.f;
    // Reference behaviour (Mueller.slang:76-77, found by the comparison with the transpiled reference text, oracle/ref_compare.cpp):
    // `fs*cconj(fp)` is the COMPONENT-WISE float2 product, not the complex one, so m22 = Re fs Re fp and m23 = -Im fs Im fp
    // (the physical values would be Re / Im of fs conj(fp)).  Identical for real coefficients (dielectrics); differs for conductors.
    const float2 sp = fs * cconj(fp);
    const float m22 = sp.x, m23 = sp.y;
    return make4x4(m00, m01, 0, 0, m01, m00, 0, 0, 0, 0, m22, m23, 0, 0, -m23, m22);
}
inline float4x4 invOneMinusMuellerFresnel(float2 fs, float2 fp) {
    const float Rp = cnorm2(fp), Rs = cnorm2(fs);
    const float m00 = (Rp + Rs) / 2, m01 = (Rp - Rs) / 2;
    const float2 sp = fs * cconj(fp);  // component-wise, as in MuellerFresnel above (Mueller.slang:87-88)
    const float m22 = sp.x, m23 = sp.y;
    const float2x2 minor1 = float2x2(1, 0, 0, 1) - float2x2(m00, m01, m01, m00);
    const float2x2 minor2 = float2x2(1, 0, 0, 1) - float2x2(m22, m23, -m23, m22);
    float2x2 im1, im2;
    if (!tryInvert(minor1, im1) || !tryInvert(minor2, im2)) return float4x4();
    // [3][3] slot takes im2[0][0] as in the reference (Mueller.slang:106-109); valid for the [[a,b],[-b,a]] block.
    return make4x4(im1.m[0][0], im1.m[0][1], 0, 0, im1.m[0][1], im1.m[0][0], 0, 0, 0, 0, im2.m[0][0], im2.m[0][1], 0, 0, im2.m[1][0], im2.m[0][0]);
}
inline float4x4 MuellerFresnelRConductor(float cosThetaI, float2 eta, float2& rs, float2& rp) {
    fresnel_conductor(cosThetaI, eta, rs, rp);
    return MuellerFresnel(rs, rp);
}
inline float4x4 MuellerFresnelRConductor(float cosThetaI, float2 eta) {
    float2 rs, rp;
    return MuellerFresnelRConductor(cosThetaI, eta, rs, rp);
}
inline float4x4 MuellerFresnelRDielectric(float cosThetaI, float eta) {
    float rs, rp, ts, tp, c;
    fresnel_dielectric(cosThetaI, eta, rs, rp, ts, tp, c);
    return MuellerFresnel(float2(rs, 0), float2(rp, 0));
}
inline float4x4 invOneMinusMuellerFresnelRDielectric(float cosThetaI, float eta) {
    float rs, rp, ts, tp, c;
    fresnel_dielectric(cosThetaI, eta, rs, rp, ts, tp, c);
    return invOneMinusMuellerFresnel(float2(rs, 0), float2(rp, 0));
}
inline float4x4 MuellerFresnelTDielectric(float cosThetaI, float eta) {
    float rs, rp, ts, tp, c;
    fresnel_dielectric(cosThetaI, eta, rs, rp, ts, tp, c);
    return MuellerFresnel(float2(ts, 0), float2(tp, 0));
}
inline float4x4 MuellerFresnelDielectric(float cosThetaI, float eta, bool reflection) {
    return reflection ? MuellerFresnelRDielectric(cosThetaI, eta) : MuellerFresnelTDielectric(cosThetaI, eta);
}

// ---------------------------------------------------------------------------------------------
// Rendering/PLT.slang:41-245
// ---------------------------------------------------------------------------------------------
struct PLTCoherence {
    float3 D;   // diffusivity matrix (xx, xy, yy)
    float r_m = 0;  // distance travelled from source (m)
    PLTCoherence() = default;
    PLTCoherence(const float2x2& Dm, float r) : D(Dm.m[0][0], Dm.m[0][1], Dm.m[1][1]), r_m(r) {}
    float2x2 toDiffusivityMat() const { return float2x2(D.x, D.y, D.y, D.z); }
    float r() const { return r_m * 1e+3f; }
    float2x2 toInvCoherenceMat(float k) const {
        const float recp_lambda = k / k2Pi;
        return sqr(recp_lambda / r()) * toDiffusivityMat();
    }
    float2x2 toInvCoherenceMat() const { return sqr(1.f / r()) * toDiffusivityMat(); }
    float det(float k) const { return 1.f / determinant(toInvCoherenceMat(k)); }
    void propagate(float dist_m) { r_m += dist_m; }
    void transform(const float2x2& U) {
        const float2x2 Dm = toDiffusivityMat();
        const float2x2 newD = mul(transpose(U), mul(Dm, U));
        D = float3(newD.m[0][0], newD.m[0][1], newD.m[1][1]);
    }
};

struct PLTBeam {
    hwss_t L, L1, L2, L3;
    hwss_t wavelengths, spectralPDFs;
    uint32_t hwss_comps = 4;
    PLTCoherence Theta;
    float3 posW, r, t;
    bool distant = false;

    PLTBeam() = default;
    PLTBeam(hwss_t Le, const float2x2& Diffusivity, float3 posW_, float3 dir, float3 tangent, hwss_t wl, bool distant_, uint32_t comps) {
        hwss_comps = comps;
        wavelengths = wl;
        spectralPDFs = hwss_t(1.f);
        L = Le;
        L1 = L2 = L3 = hwss_t(.0f);
        Theta = PLTCoherence(Diffusivity, distant_ ? 1e-3f : .0f);
        posW = posW_;
        r = dir;
        t = tangent;
        distant = distant_;
    }
    float4 getSP(int j) const { return float4(L[j], L1[j], L2[j], L3[j]); }
    void setSP(int j, float4 SP) { L[j] = SP.x; L1[j] = SP.y; L2[j] = SP.z; L3[j] = SP.w; }
    void setZero() { L = L1 = L2 = L3 = hwss_t(.0f); }
    void scale(hwss_t s) { L *= s; L1 *= s; L2 *= s; L3 *= s; }
    void scale(float s) { L *= s; L1 *= s; L2 *= s; L3 *= s; }

    float2x2 Q() const {
        const float2 t2 = t.xy();
        const float2 b2 = cross(t, r).xy();
        return transpose(float2x2(t2.x, t2.y, b2.x, b2.y));
    }
    float2x2 invTheta(float k) const { return Theta.toInvCoherenceMat(k); }
    float2x2 invTheta() const { return Theta.toInvCoherenceMat(); }
    hwss_t mutualCoherence(hwss_t k, float3 v) const {
        const hwss_t recp_lambda = k / k2Pi;
        const float2x2 invT = invTheta();
        const hwss_t x = sqr(recp_lambda) * dot(v.xy(), mul(invT, v.xy()));
        return vexp(-.5f * x);
    }
    float mutualCoherenceAngular(float3 d1, float3 d2) const {
        const float2 v = 1.f / vmax(std::sqrt(k4Pi) * vabs(d1 - d2).xy(), float2(FLT_EPS));
        const float2x2 invT = invTheta() * sqr(Theta.r());
        const float x = 1.f / dot(v, mul(invT, v));
        return std::exp(-.5f * x);
    }
    float coherenceArea(float k) const { return kPi * std::sqrt(Theta.det(k)); }
    void transformCoherence(const float2x2& U) { Theta.transform(U); }
    void rotateFrame(float3 newt, float3 newr) {
        float sint, cost;
        float4x4 T = MuellerT(r, t, newr, newt, sint, cost);
        for (uint32_t j = 0; j < hwss_comps; ++j) setSP(j, mul(T, getSP(j)));
        transformCoherence(float2x2(cost, -sint, sint, cost));
        r = newr;
        t = newt;
    }
    void constructLocalFrame(float3 newwo) {
        float3 newt = cross(float3(0, 0, 1), newwo);
        const float l2 = dot(newt, newt);
        newt = l2 > 1e-8 ? newt / std::sqrt(l2) : perp_stark(newwo);
        rotateFrame(newt, newwo);
    }
    void propagate(float3 p) {
        if (!distant) Theta.propagate(length(p - posW));
        posW = p;
    }
    float3 fromLocal(float3 v) const { return t * v.x + cross(t, r) * v.y + r * v.z; }
    float3 toLocal(float3 v) const { return float3(dot(v, t), dot(v, cross(t, r)), dot(v, r)); }
    void collapseToHeroWavelengthOnly() { hwss_comps = 1; }
};

// PLT.slang:253-290
inline PLTBeam pltSourceBeamDistant(float3 dir, float solidAngle, hwss_t Le, hwss_t wavelengths, float maxBeamOmega, bool forceFullyCoherent, uint32_t comps) {
    solidAngle = hlsl_min(solidAngle, maxBeamOmega);
    const float2x2 Diffusivity = forceFullyCoherent ? 1e-9f * ident2() : solidAngle * ident2();
    return PLTBeam(Le, Diffusivity, float3(.0f), dir, perp_stark(dir), wavelengths, true, comps);
}
inline PLTBeam pltSourceBeamArea(float3 posW, float3 dir, float A, float dist, hwss_t Le, hwss_t wavelengths, float maxBeamOmega, bool forceFullyCoherent, uint32_t comps) {
    const float r2 = sqr(dist * 1e+3f);
    A = hlsl_min(A, maxBeamOmega * r2);
    const float2x2 Diffusivity = forceFullyCoherent ? 1e-7f * ident2() : A * ident2();
    return PLTBeam(Le, Diffusivity, posW, dir, perp_stark(dir), wavelengths, false, comps);
}

}  // namespace orc
