// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// ref_compare.cpp — pins the oracle (namespace orc: the hand-written C++ restatement of the reference's Slang path, oracle/*.h) to
// the reference's OWN shader text (namespace ref: oracle/_ref/ref_gen.h, produced from /root/reference by the purely textual
// rewrite of oracle/ref_transpile.py and compiled against oracle/ref/slang_compat.h).  Both sides are called on the same seeded
// random inputs; per function the program prints the number of comparisons, the largest distance in units in the last place and
// the largest relative deviation as one JSON object (tests/test_oracle_ref.py asserts on it).
// Build: make -C oracle ref      (g++ -O2 -ffp-contract=off, like the oracle itself)
#include <cinttypes>
#include <cstdio>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "plt_integrator.h"   // the oracle: orc::
#include "_ref/ref_gen.h"     // the transpiled reference: ref::

namespace {
struct Stat { uint64_t n = 0, nonfiniteMismatch = 0, identical = 0, units = 0; double maxUlp = 0, maxRel = 0, maxAbs = 0; };
std::map<std::string, Stat> g_stats;
std::vector<std::string> g_order;

double ulpDist(float a, float b) {
    if (a == b) return 0;
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4); std::memcpy(&ib, &b, 4);
    if (ia < 0) ia = int32_t(0x80000000u) - ia;
    if (ib < 0) ib = int32_t(0x80000000u) - ib;
    return std::fabs(double(int64_t(ia) - int64_t(ib)));
}
struct Cmp {
    Stat& s;
    explicit Cmp(const std::string& name) : s(g_stats[name]) { if (s.n == 0 && std::find(g_order.begin(), g_order.end(), name) == g_order.end()) g_order.push_back(name); }
    void add(float a, float b) {
        ++s.n;
        const bool fa = std::isfinite(a), fb = std::isfinite(b);
        if (!fa || !fb) { if (fa != fb || (std::isnan(a) != std::isnan(b)) || (std::isinf(a) && a != b)) ++s.nonfiniteMismatch; return; }
        s.maxUlp = std::max(s.maxUlp, ulpDist(a, b));
        s.maxAbs = std::max(s.maxAbs, std::fabs(double(a) - double(b)));
        const double m = std::max(std::fabs(double(a)), std::fabs(double(b)));
        if (m > 1e-30) s.maxRel = std::max(s.maxRel, std::fabs(double(a) - double(b)) / m);
    }
    void add(bool a, bool b) { ++s.n; if (a != b) { ++s.nonfiniteMismatch; } }
    void add(uint32_t a, uint32_t b) { ++s.n; if (a != b) { ++s.nonfiniteMismatch; } }
    void add(int a, int b) { ++s.n; if (a != b) { ++s.nonfiniteMismatch; } }
    void add(orc::float2 a, ref::float2 b) { add(a.x, b.x); add(a.y, b.y); }
    void add(orc::float3 a, ref::float3 b) { add(a.x, b.x); add(a.y, b.y); add(a.z, b.z); }
    void add(orc::float4 a, ref::float4 b) { add(a.x, b.x); add(a.y, b.y); add(a.z, b.z); add(a.w, b.w); }
    void add(const orc::float2x2& a, const ref::float2x2& b) { for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) add(a.m[i][j], b[i][j]); }
    void add(const orc::float4x4& a, const ref::float4x4& b) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) add(a.m[i][j], b[i][j]); }
    void add(const orc::PLTBeam& a, const ref::PLTBeam& b) {
        add(a.L, b.L); add(a.L1, b.L1); add(a.L2, b.L2); add(a.L3, b.L3); add(a.wavelengths, b.wavelengths); add(a.spectralPDFs, b.spectralPDFs);
        add(uint32_t(a.hwss_comps), uint32_t(b.hwss_comps)); add(a.Theta.D, b.Theta.D); add(a.Theta.r_m, b.Theta.r_m);
        add(a.posW, b.posW); add(a.r, b.r); add(a.t, b.t); add(a.distant, b.distant);
    }
};
ref::float2 R(orc::float2 v) { return ref::float2(v.x, v.y); }
ref::float3 R(orc::float3 v) { return ref::float3(v.x, v.y, v.z); }
ref::float4 R(orc::float4 v) { return ref::float4(v.x, v.y, v.z, v.w); }
ref::float2x2 R(const orc::float2x2& m) { return ref::float2x2(m.m[0][0], m.m[0][1], m.m[1][0], m.m[1][1]); }

struct Rnd {
    std::mt19937 g;
    explicit Rnd(uint32_t seed) : g(seed) {}
    float u() { return float(g() >> 8) * 0x1p-24f; }
    float range(float a, float b) { return a + (b - a) * u(); }
    orc::float2 f2(float a, float b) { return orc::float2(range(a, b), range(a, b)); }
    orc::float3 f3(float a, float b) { return orc::float3(range(a, b), range(a, b), range(a, b)); }
    orc::float4 f4(float a, float b) { return orc::float4(range(a, b), range(a, b), range(a, b), range(a, b)); }
    orc::float3 dir() { orc::float3 v; do { v = f3(-1.f, 1.f); } while (orc::dot(v, v) < 1e-3f || orc::dot(v, v) > 1.f); return orc::normalize(v); }
    orc::float3 dirUp() { orc::float3 v = dir(); v.z = std::fabs(v.z); return v; }
    orc::float4 wavelengths() { return orc::SpectrumSampler::hwssWavelengths(range(380.f, 720.f)); }
    orc::float4 ks() { const orc::float4 w = wavelengths(); return orc::float4(orc::k2Pi / (1e-3f * w.x), orc::k2Pi / (1e-3f * w.y), orc::k2Pi / (1e-3f * w.z), orc::k2Pi / (1e-3f * w.w)); }
};
orc::PLTBeam randomBeam(Rnd& r, ref::PLTBeam& rb) {
    const orc::float3 pos = r.f3(-2.f, 2.f), dir = r.dir();
    const orc::float4 Le = r.f4(0.f, 3.f), wl = r.wavelengths();
    const bool distant = r.u() < 0.3f;
    orc::PLTBeam b;
    if (distant) {
        const float omega = r.range(1e-5f, 0.05f);
        b = orc::pltSourceBeamDistant(dir, omega, Le, wl, 0.025f, false, 4);
        rb = ref::pltSourceBeamDistant(R(dir), omega, R(Le), R(wl), 0.025f, false, 4);
    } else {
        const float A = r.range(0.01f, 30.f), dist = r.range(0.05f, 10.f);
        b = orc::pltSourceBeamArea(pos, dir, A, dist, Le, wl, 0.025f, false, 4);
        rb = ref::pltSourceBeamArea(R(pos), R(dir), A, dist, R(Le), R(wl), 0.025f, false, 4);
    }
    // give the beam some polarisation and anisotropic coherence so that every field is exercised
    b.L1 = b.L * r.range(-.5f, .5f); b.L2 = b.L * r.range(-.5f, .5f); b.L3 = b.L * r.range(-.3f, .3f);
    rb.L1 = R(b.L1); rb.L2 = R(b.L2); rb.L3 = R(b.L3);
    return b;
}
// the reference's generic code takes any generator with next(); wrapped so that argument-dependent lookup stays inside namespace ref
struct RefSG { orc::SampleGenerator g; ref::uint next() { return g.next(); } };
}  // namespace
#include "ref/ref_shims_post.h"

int main(int argc, char** argv) {
    const int N = argc > 1 ? std::atoi(argv[1]) : 100000;
    orc::hwss_cfg().samples = 4;
    Rnd r(12345u);
    // ---- MathHelpers.slang
    for (int i = 0; i < N; ++i) {
        const orc::float2 z = r.f2(-3.f, 3.f), w = r.f2(-3.f, 3.f), u = r.f2(0.f, 1.f);
        { Cmp c("MathHelpers.complex"); c.add(orc::csqrt(z), ref::csqrt(R(z))); c.add(orc::cdiv(z, w), ref::cdiv(R(z), R(w))); c.add(orc::cmul(z, w), ref::cmul(R(z), R(w))); c.add(orc::carg(z), ref::carg(R(z))); c.add(orc::cabs(z), ref::cabs(R(z))); }
        { Cmp c("MathHelpers.sample_shapes"); c.add(orc::sample_disk(u), ref::sample_disk(R(u))); c.add(orc::sample_sphere(u), ref::sample_sphere(R(u))); c.add(orc::sample_triangle(u), ref::sample_triangle(R(u)));
          const float ct = r.range(0.5f, 0.99999f); c.add(orc::sample_cone(u, ct), ref::sample_cone(R(u), ct));
          float p0, p1; c.add(orc::sample_cosine_hemisphere_concentric(u, p0), ref::sample_cosine_hemisphere_concentric(R(u), p1)); c.add(p0, p1); }
        { Cmp c("MathHelpers.boxmueller_truncated"); const orc::float2 mean = r.f2(-.7f, .7f); const float s2 = r.range(1e-5f, 0.2f); float p0, p1;
          const orc::float2 a = orc::sample_boxmueller_truncated(u, mean, s2, p0); const ref::float2 b = ref::sample_boxmueller_truncated(R(u), R(mean), s2, p1);
          c.add(a, b); c.add(p0, p1); c.add(orc::boxmueller_truncated_pdf(a, mean, s2), ref::boxmueller_truncated_pdf(b, R(mean), s2)); }
        { Cmp c("MathHelpers.matrices"); const orc::float2x2 M(r.range(-2.f, 2.f), r.range(-2.f, 2.f), r.range(-2.f, 2.f), r.range(-2.f, 2.f)); orc::float2x2 i0, U0, S0; ref::float2x2 i1, U1, S1;
          const bool ok0 = orc::tryInvert(M, i0), ok1 = ref::tryInvert(R(M), i1); c.add(ok0, ok1); if (ok0 && ok1) c.add(i0, i1);
          orc::svd_symmetric(M, U0, S0); ref::svd_symmetric(R(M), U1, S1); c.add(U0, U1); c.add(S0, S1); }
        { Cmp c("MathHelpers.perp_stark_oct"); const orc::float3 d = r.dir(); c.add(orc::perp_stark(d), ref::perp_stark(R(d)));
          const orc::float2 e = orc::ndir_to_oct_equal_area_snorm(d); c.add(e, ref::ndir_to_oct_equal_area_snorm(R(d))); c.add(orc::oct_to_ndir_equal_area_snorm(e), ref::oct_to_ndir_equal_area_snorm(R(e))); }
        { Cmp c("MathHelpers.BesselJ"); const int nu = int(r.range(-6.99f, 6.99f)); const float x = r.range(0.01f, 40.f); c.add(orc::BesselJ(nu, x), ref::BesselJ(nu, x)); }
    }
    // ---- SpectrumUtils.slang / ColorHelpers.slang
    for (int i = 0; i < N; ++i) {
        const float l = r.range(380.f, 720.f);
        const orc::float4 wl = r.wavelengths(), sp = r.f4(0.f, 2.f);
        const orc::float3 rgb = r.f3(0.f, 1.f);
        { Cmp c("SpectrumUtils.wavelengthToXYZ_D65"); c.add(orc::SpectrumUtils::wavelengthToXYZ(l), ref::SpectrumUtils::wavelengthToXYZ(l)); c.add(orc::SpectrumUtils::wavelengthToD65(l), ref::SpectrumUtils::wavelengthToD65(l)); }
        { Cmp c("SpectrumUtils.rgbToSpectrum"); c.add(orc::SpectrumUtils::rgbToSpectrum(rgb, l), ref::SpectrumUtils::rgbToSpectrum(R(rgb), l)); c.add(orc::SpectrumUtils::rgbToSpectrum(rgb, wl), ref::SpectrumUtils::rgbToSpectrum(R(rgb), R(wl))); }
        { Cmp c("SpectrumUtils.spectrumToRgb"); const uint32_t comps = 1 + uint32_t(r.u() * 3.999f); c.add(orc::SpectrumUtils::spectrumToRgb(sp, wl, comps), ref::SpectrumUtils::spectrumToRgb(R(sp), R(wl), comps)); }
    }
    // ---- Fresnel.slang / Mueller.slang
    for (int i = 0; i < N; ++i) {
        const float ci = r.range(-1.f, 1.f), eta = r.u() < 0.05f ? 1.f : r.range(1.01f, 2.6f);
        const orc::float2 ceta(r.range(0.05f, 3.f), r.range(0.f, 6.f));
        { Cmp c("Fresnel.fresnel_dielectric"); float a[5], b[5]; orc::fresnel_dielectric(ci, eta, a[0], a[1], a[2], a[3], a[4]); ref::fresnel_dielectric(ci, eta, b[0], b[1], b[2], b[3], b[4]); for (int k = 0; k < 5; ++k) c.add(a[k], b[k]);
          c.add(orc::evalFresnelDielectric(eta, ci), ref::evalFresnelDielectric(eta, ci)); }
        { Cmp c("Fresnel.fresnel_conductor"); orc::float2 rs0, rp0; ref::float2 rs1, rp1; const float cc = r.range(-0.1f, 1.f); orc::fresnel_conductor(cc, ceta, rs0, rp0); ref::fresnel_conductor(cc, R(ceta), rs1, rp1); c.add(rs0, rs1); c.add(rp0, rp1);
          c.add(orc::evalFresnelConductor(ceta, std::fabs(cc)), ref::evalFresnelConductor(R(ceta), std::fabs(cc))); }
        { Cmp c("Mueller.MuellerT"); const orc::float3 n1 = r.dir(); const orc::float3 t1 = orc::perp_stark(n1); const orc::float3 n2 = r.u() < .5f ? n1 : r.dir(); const orc::float3 t2 = orc::normalize(orc::cross(n2, r.dir()));
          float s0, c0, s1, c1; c.add(orc::MuellerT(n1, t1, n2, t2, s0, c0), ref::MuellerT(R(n1), R(t1), R(n2), R(t2), s1, c1)); c.add(s0, s1); c.add(c0, c1); }
        { Cmp c("Mueller.MuellerPolarizer"); const float th = r.range(-7.f, 7.f); c.add(orc::MuellerPolarizer(th), ref::MuellerPolarizer(th)); }
        { Cmp c("Mueller.MuellerFresnel"); const orc::float2 fs = r.f2(-1.f, 1.f), fp = r.f2(-1.f, 1.f); c.add(orc::MuellerFresnel(fs, fp), ref::MuellerFresnel(R(fs), R(fp))); c.add(orc::invOneMinusMuellerFresnel(fs, fp), ref::invOneMinusMuellerFresnel(R(fs), R(fp))); }
        { Cmp c("Mueller.MuellerFresnelRConductor"); const float cc = r.range(0.f, 1.f); c.add(orc::MuellerFresnelRConductor(cc, ceta), ref::MuellerFresnelRConductor(cc, R(ceta))); }
        { Cmp c("Mueller.MuellerFresnelDielectric"); c.add(orc::MuellerFresnelDielectric(ci, eta, true), ref::MuellerFresnelDielectric(ci, eta, true)); c.add(orc::MuellerFresnelDielectric(ci, eta, false), ref::MuellerFresnelDielectric(ci, eta, false));
          c.add(orc::invOneMinusMuellerFresnelRDielectric(ci, eta), ref::invOneMinusMuellerFresnelRDielectric(ci, eta)); }
    }
    // ---- PLT.slang (generalized ray)
    for (int i = 0; i < N; ++i) {
        ref::PLTBeam rb;
        orc::PLTBeam b = randomBeam(r, rb);
        { Cmp c("PLT.sourceBeam"); c.add(b, rb); }
        const orc::float3 p = r.f3(-3.f, 3.f);
        b.propagate(p); rb.propagate(R(p));
        { Cmp c("PLT.propagate"); c.add(b, rb); }
        // into a local frame (as the BSDF wrappers do), rotate to an exitant direction, transform coherence
        const orc::float3 N_ = r.dir(), T_ = orc::perp_stark(N_), B_ = orc::cross(N_, T_);
        auto toLocal = [&](orc::float3 v) { return orc::float3(orc::dot(v, T_), orc::dot(v, B_), orc::dot(v, N_)); };
        b.r = toLocal(b.r); b.t = toLocal(b.t); rb.r = R(b.r); rb.t = R(b.t);
        const orc::float3 wo = r.u() < 0.02f ? orc::float3(0.f, 0.f, 1.f) : r.dir();
        b.constructLocalFrame(wo); rb.constructLocalFrame(R(wo));
        { Cmp c("PLT.constructLocalFrame_rotateFrame"); c.add(b, rb); }
        const orc::float4 k = r.ks();
        const orc::float3 v = r.f3(-2.f, 2.f), d1 = r.dir(), d2 = orc::normalize(d1 + r.f3(-.05f, .05f));
        { Cmp c("PLT.coherence"); c.add(b.mutualCoherence(k, v), rb.mutualCoherence(R(k), R(v))); c.add(b.mutualCoherenceAngular(d1, d2), rb.mutualCoherenceAngular(R(d1), R(d2)));
          c.add(b.Q(), rb.Q()); c.add(b.invTheta(k.x), rb.invTheta(k.x)); c.add(b.invTheta(), rb.invTheta()); c.add(b.toLocal(v), rb.toLocal(R(v))); c.add(b.fromLocal(v), rb.fromLocal(R(v)));
          c.add(b.coherenceArea(k.x), rb.coherenceArea(k.x)); }
        const float ct = std::sqrt(std::fmax(1e-2f, std::fabs(wo.z)));
        const orc::float2x2 U(ct, 0.f, 0.f, 1.f / ct);
        b.transformCoherence(U); rb.transformCoherence(R(U));
        { Cmp c("PLT.transformCoherence"); c.add(b, rb); }
    }
    // ---- Interference.slang
    for (int i = 0; i < N; ++i) {
        Cmp c("Interference.interfere_phaseShift");
        const orc::float4 a = r.f4(0.f, 2.f) * orc::float4(1.f, .5f, .5f, .3f), b = r.f4(0.f, 2.f) * orc::float4(1.f, -.5f, .4f, .2f);
        const float ps = r.range(-1.f, 1.f), pp = r.range(-1.f, 1.f), coh = r.u();
        c.add(orc::interfere(a, b, ps, pp, coh), ref::interfere(R(a), R(b), ps, pp, coh));
        const float eta = r.range(1.f, 2.5f), dist = r.range(0.f, 20.f);
        const orc::float2 ph = r.f2(-1.f, 1.f);
        const orc::float4 k = r.ks();
        c.add(orc::phaseShift(eta, dist, k.x, ph), ref::phaseShift(eta, dist, k.x, R(ph)));
        c.add(orc::phaseShift(eta, dist, k), ref::phaseShift(eta, dist, R(k)));
    }
    // ---- FractalSurface.slang / GaussianSurface.slang
    for (int i = 0; i < N; ++i) {
        const float rough = r.u() < 0.05f ? 0.f : r.range(0.f, 0.9f), gamma = r.range(1.f, 4.f);
        const orc::float3 wi = r.dirUp(), wo = r.dir(), h = wi + wo;
        const orc::float4 k = r.ks();
        {
            const orc::FractalSurface a(rough, gamma); ref::FractalSurface b(rough, gamma);
            { Cmp c("FractalSurface.parameters"); c.add(a.T, b.T); c.add(a.gamma, b.gamma); c.add(a.roughness(), b.roughness()); c.add(a.q(), b.q()); c.add(a.isDeltaOnly(), b.isDeltaOnly()); }
            { Cmp c("FractalSurface.alpha"); c.add(a.alpha(wi, wo, k.x), b.alpha(R(wi), R(wo), k.x)); c.add(a.alpha(wi, wo, k), b.alpha(R(wi), R(wo), R(k))); c.add(a.alpha(wi, k.y), b.alpha(R(wi), k.y)); }
            if (rough > 0.f) {
                Cmp c("FractalSurface.eval_pdf_sample");
                c.add(a.eval(k.x, h), b.eval(k.x, R(h))); c.add(a.pdf(wi, wo, k.x), b.pdf(R(wi), R(wo), k.x));
                orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 7u); RefSG g1{g0};
                float p0, p1;
                c.add(a.sample(wi, g0, k.z, p0), b.sample(R(wi), g1, k.z, p1)); c.add(p0, p1); c.add(g0.next(), g1.next());
            }
        }
        {
            const orc::GaussianSurface a(rough); ref::GaussianSurface b(rough);
            { Cmp c("GaussianSurface.parameters"); c.add(a.sigma2, b.sigma2); c.add(a.q, b.q); c.add(a.roughness(), b.roughness()); }
            { Cmp c("GaussianSurface.alpha"); c.add(a.alpha(wi, wo, k.x), b.alpha(R(wi), R(wo), k.x)); c.add(a.alpha(wi, wo, k), b.alpha(R(wi), R(wo), R(k))); }
            if (rough > 0.f) {
                Cmp c("GaussianSurface.eval_pdf_sample");
                ref::PLTBeam rb; const orc::PLTBeam beam = randomBeam(r, rb);
                c.add(a.eval(k.x, beam.invTheta(1.f), beam.Q(), h), b.eval(k.x, rb.invTheta(1.f), rb.Q(), R(h)));
                c.add(a.pdf(wi, wo, k.x), b.pdf(R(wi), R(wo), k.x));
                orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 9u); RefSG g1{g0};
                float p0, p1;
                c.add(a.sample(wi, g0, k.w, p0), b.sample(R(wi), g1, k.w, p1)); c.add(p0, p1);
            }
        }
    }
    // ---- Birefringence.slang
    for (int i = 0; i < N; ++i) {
        const float eo = r.range(1.3f, 1.7f), ee = eo + r.range(-.15f, .15f), ei = 1.f;
        const orc::float3 A = r.dir();
        const float ct = r.range(0.05f, 1.f), st = std::sqrt(1.f - ct * ct), K = ei * st;
        { Cmp c("Birefringence.modes_fields"); c.add(orc::Birefringence::qo(eo, K), ref::Birefringence::qo(eo, K)); c.add(orc::Birefringence::qe(eo, ee, A, K), ref::Birefringence::qe(eo, ee, R(A), K));
          c.add(orc::Birefringence::E_ep(eo, ee, A, K), ref::Birefringence::E_ep(eo, ee, R(A), K)); c.add(orc::Birefringence::E_em(eo, ee, A, K), ref::Birefringence::E_em(eo, ee, R(A), K));
          c.add(orc::Birefringence::E_op(eo, A, K), ref::Birefringence::E_op(eo, R(A), K)); c.add(orc::Birefringence::E_om(eo, A, K), ref::Birefringence::E_om(eo, R(A), K));
          const orc::float3 E = r.dir(); const float q = r.range(-1.7f, 1.7f); c.add(orc::Birefringence::Poynting(q, E, K), ref::Birefringence::Poynting(q, R(E), K)); }
        { Cmp c("Birefringence.fresnel_iso_aniso"); float a[8], b[8];
          orc::Birefringence::fresnel_iso_aniso(ct, st, ei, eo, ee, A, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); ref::Birefringence::fresnel_iso_aniso(ct, st, ei, eo, ee, R(A), b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
          for (int k = 0; k < 8; ++k) c.add(a[k], b[k]); }
        { Cmp c("Birefringence.fresnel_aniso_iso"); float a[8], b[8];
          orc::Birefringence::fresnel_aniso_iso(ct, st, ei, eo, ee, A, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); ref::Birefringence::fresnel_aniso_iso(ct, st, ei, eo, ee, R(A), b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
          for (int k = 0; k < 8; ++k) c.add(a[k], b[k]); }
        { Cmp c("Birefringence.vectors_in_slab"); const orc::float3 I(0.f, ct, st); orc::float3 a[4]; ref::float3 b[4]; float e0[2], e1[2];
          orc::Birefringence::vectors_in_slab(I, ei, eo, ee, A, a[0], a[1], a[2], a[3], e0[0], e0[1]); ref::Birefringence::vectors_in_slab(R(I), ei, eo, ee, R(A), b[0], b[1], b[2], b[3], e1[0], e1[1]);
          for (int k = 0; k < 4; ++k) c.add(a[k], b[k]); c.add(e0[0], e1[0]); c.add(e0[1], e1[1]); }
    }
    ref_compare_more(N, r);  // gratings, spectral profiles, BSDFs (ref_shims_post.h)
    std::printf("{");
    for (size_t i = 0; i < g_order.size(); ++i) {
        const Stat& s = g_stats[g_order[i]];
        std::printf("%s\n \"%s\": {\"n\": %" PRIu64 ", \"max_ulp\": %.0f, \"max_rel\": %.3e, \"max_abs\": %.3e, \"mismatches\": %" PRIu64, i ? "," : "", g_order[i].c_str(), s.n, s.maxUlp, s.maxRel, s.maxAbs, s.nonfiniteMismatch);
        if (s.units) std::printf(", \"units\": %" PRIu64 ", \"identical_units\": %" PRIu64, s.units, s.identical);
        std::printf("}");
    }
    std::printf("\n}\n");
    return 0;
}
