// ORACLE — TEST INFRASTRUCTURE ONLY.
// ref_shims_post.h — second half of oracle/ref_compare.cpp (included after both implementations and the helpers are visible):
// spectral-profile lookups, diffraction gratings, differential reflect / refract, and eval / sample / pdf / lambdaPdf of the ten PLT
// BSDFs through their IMaterialInstance wrappers -- oracle (orc::) against the transpiled reference text (ref::).
#pragma once

#include "ref_shims_end.h"

namespace {
ref::ShadingData R(const orc::ShadingData& s) {
    ref::ShadingData o;
    o.posW = R(s.posW); o.V = R(s.V); o.N = R(s.N); o.T = R(s.T); o.B = R(s.B); o.faceN = R(s.faceN); o.uv = R(s.uv); o.frontFacing = s.frontFacing; o.materialID = s.materialID; o.tangentW = R(s.tangentW); o.opacity = s.opacity;
    return o;
}
ref::PLTBeam R(const orc::PLTBeam& b) {
    ref::PLTBeam o;
    o.L = R(b.L); o.L1 = R(b.L1); o.L2 = R(b.L2); o.L3 = R(b.L3); o.wavelengths = R(b.wavelengths); o.spectralPDFs = R(b.spectralPDFs); o.hwss_comps = b.hwss_comps;
    o.Theta.D = R(b.Theta.D); o.Theta.r_m = b.Theta.r_m; o.posW = R(b.posW); o.r = R(b.r); o.t = R(b.t); o.distant = b.distant;
    return o;
}
ref::float3pdv R(const orc::float3pdv& v) { ref::float3pdv o; o.v = R(v.v); o.ddu = R(v.ddu); o.ddv = R(v.ddv); return o; }
void addK(Cmp& c, const orc::float3kpdv& a, const ref::float3kpdv& b) { c.add(a.v, b.v); c.add(a.ddu, b.ddu); c.add(a.ddv, b.ddv); c.add(a.ddl, b.ddl); }

orc::ShadingData randomShadingData(Rnd& r) {
    orc::ShadingData sd;
    sd.N = r.dir(); sd.T = orc::perp_stark(sd.N); sd.B = orc::cross(sd.N, sd.T);
    sd.posW = r.f3(-1.f, 1.f); sd.faceN = sd.N; sd.uv = r.f2(0.f, 1.f);
    orc::float3 v = r.dirUp();
    if (v.z < 0.02f) v = orc::normalize(orc::float3(v.x, v.y, 0.02f));
    sd.V = sd.fromLocal(v);
    return sd;
}
// a lobe word the wrappers accept for the material (the sample() of the material itself provides realistic ones)
const uint32_t kLobeChoices[] = {orc::L_DiffuseReflection, orc::L_ScatterReflection, orc::L_DeltaReflection, orc::L_ScatterTransmission, orc::L_DeltaTransmission,
                                 orc::L_DispersiveDeltaTransmission, orc::L_NonDelta, orc::L_Delta, orc::L_All};

// One material of each kind with random parameters, set field by field on both sides (the *Material.slang loaders that unpack the
// 128-byte blob need the texture system and are outside this comparison; tests/test_ingest.py covers the packing).
struct Pair {
    orc::MaterialInstance o;
    int kind;
    ref::PLTDiffuseBSDF diffuse; ref::PLTConductorBSDF conductor; ref::PLTDielectricBSDF dielectric; ref::PLTOpaqueDielectricBSDF opaque; ref::PLTThinDielectricBSDF thin;
    ref::PLTCoatedConductorBSDF coatedC; ref::PLTCoatedOpaqueDielectricBSDF coatedD; ref::PLTDiffractionGratedConductorBSDF gratedC; ref::PLTDiffractionGratedOpaqueDielectricBSDF gratedD;
    ref::PLTMultiLayeredStackBSDF stack;
};
const char* kKindNames[10] = {"PLTDiffuseBSDF", "PLTConductorBSDF", "PLTDielectricBSDF", "PLTOpaqueDielectricBSDF", "PLTThinDielectricBSDF", "PLTCoatedConductorBSDF",
                              "PLTCoatedOpaqueDielectricBSDF", "PLTDiffractionGratedConductorBSDF", "PLTDiffractionGratedOpaqueDielectricBSDF", "PLTMultiLayeredStackBSDF"};
void makePair(Pair& p, int kind, Rnd& r, const orc::ProfileTable* prof, const orc::ShadingData& sd) {
    p.kind = kind;
    p.o = orc::MaterialInstance();
    p.o.prof = prof;
    const orc::float3 col = r.f3(0.05f, 1.f);
    const float rough = r.u() < 0.15f ? 0.f : r.range(0.01f, 0.6f), gamma = r.range(1.5f, 4.f), extEta = r.u() < .5f ? 1.f : r.range(1.f, 1.4f);
    const uint32_t n = uint32_t(r.u() * 1.999f) * 2u, k = n + 1u;  // profile ids (n, k) of one of the two tabulated conductors
    switch (kind) {
        case 0: p.o.type = PLTPT_MATERIAL_DIFFUSE; p.o.diffuse.albedo = col; p.diffuse.bxdf.albedo = R(col); break;
        case 1: p.o.type = PLTPT_MATERIAL_CONDUCTOR;
            p.o.conductor.specular = col; p.o.conductor.gfs = orc::FractalSurface(rough, gamma); p.o.conductor.etaN = n; p.o.conductor.etaK = k; p.o.conductor.extEta = extEta; p.o.conductor.prof = prof;
            p.conductor.bxdf.specular = R(col); p.conductor.bxdf.gfs = ref::FractalSurface(rough, gamma); p.conductor.bxdf.etaN = n; p.conductor.bxdf.etaK = k; p.conductor.bxdf.extEta = extEta; break;
        case 2: { p.o.type = PLTPT_MATERIAL_DIELECTRIC;
            const orc::float3 tr = r.f3(0.3f, 1.f); const float etaInt = r.range(1.2f, 2.2f), B = r.u() < .5f ? 0.f : r.range(0.f, 0.02f);
            p.o.dielectric.specular = col; p.o.dielectric.transmission = tr; p.o.dielectric.gfs = orc::GaussianSurface(rough); p.o.dielectric.etaInt = etaInt; p.o.dielectric.etaExt = extEta; p.o.dielectric.dispersionB = B;
            p.dielectric.bxdf.specular = R(col); p.dielectric.bxdf.transmission = R(tr); p.dielectric.bxdf.gfs = ref::GaussianSurface(rough); p.dielectric.bxdf.etaInt = etaInt; p.dielectric.bxdf.etaExt = extEta; p.dielectric.bxdf.dispersionB = B; break; }
        case 3: p.o.type = PLTPT_MATERIAL_OPAQUE_DIELECTRIC; p.o.opaque.albedo = col; p.o.opaque.gfs = orc::FractalSurface(rough);
            p.opaque.bxdf.albedo = R(col); p.opaque.bxdf.gfs = ref::FractalSurface(rough); break;
        case 4: { p.o.type = PLTPT_MATERIAL_THIN_DIELECTRIC;
            const orc::float2 oa = r.f2(-.5f, .5f); const float eta = r.range(1.2f, 1.9f), tau = r.range(0.01f, 1.f), bir = r.u() < .3f ? 0.f : r.range(0.f, 0.2f);
            p.o.thin.opticAxis_xy = oa; p.o.thin.etaAvg = eta; p.o.thin.tau = tau; p.o.thin.birefringence = bir;
            p.thin.bxdf.opticAxis_xy = R(oa); p.thin.bxdf.etaAvg = eta; p.thin.bxdf.tau = tau; p.thin.bxdf.birefringence = bir; break; }
        case 5: case 6: { const bool cond = kind == 5; p.o.type = cond ? PLTPT_MATERIAL_COATED_CONDUCTOR : PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC;
            const float coat = r.range(1.2f, 1.8f), tau = r.u() < .2f ? r.range(0.f, .05f) : r.range(0.05f, 5.f);
            p.o.coated.conductor = cond; p.o.coated.specular = col; p.o.coated.gfs = orc::FractalSurface(rough); p.o.coated.etaN = n; p.o.coated.etaK = k; p.o.coated.extEta = extEta; p.o.coated.coatIOR = coat; p.o.coated.tau = tau; p.o.coated.prof = prof;
            if (cond) { p.coatedC.bxdf.specular = R(col); p.coatedC.bxdf.gfs = ref::FractalSurface(rough); p.coatedC.bxdf.etaN = n; p.coatedC.bxdf.etaK = k; p.coatedC.bxdf.extEta = extEta; p.coatedC.bxdf.coatIOR = coat; p.coatedC.bxdf.tau = tau; }
            else { p.coatedD.bxdf.albedo = R(col); p.coatedD.bxdf.gfs = ref::FractalSurface(rough); p.coatedD.bxdf.extEta = extEta; p.coatedD.bxdf.coatIOR = coat; p.coatedD.bxdf.tau = tau; }
            break; }
        case 7: case 8: { const bool cond = kind == 7; p.o.type = cond ? PLTPT_MATERIAL_GRATED_CONDUCTOR : PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC;
            const float dir = r.range(-3.f, 3.f), q = r.range(0.02f, 0.5f), mult = r.range(.5f, 1.5f);
            const orc::float2 ip(1.f / r.range(1.f, 6.f), r.u() < .5f ? 0.f : 1.f / r.range(1.f, 6.f));
            const uint32_t lobes = 1u + uint32_t(r.u() * 7.99f), gtype = uint32_t(r.u() * 2.99f) | (r.u() < .2f ? 0x10u : 0u);
            p.o.grated.conductor = cond;
            p.o.grated.baseC.specular = col; p.o.grated.baseC.gfs = orc::FractalSurface(rough); p.o.grated.baseC.etaN = n; p.o.grated.baseC.etaK = k; p.o.grated.baseC.extEta = extEta; p.o.grated.baseC.prof = prof;
            p.o.grated.baseD.albedo = col; p.o.grated.baseD.gfs = orc::FractalSurface(rough);
            p.o.grated.grating = orc::DiffractionGrating(dir, sd, ip, q, lobes, gtype, mult);
            ref::ShadingData rsd = R(sd);
            if (cond) { p.gratedC.bxdf.base.specular = R(col); p.gratedC.bxdf.base.gfs = ref::FractalSurface(rough); p.gratedC.bxdf.base.etaN = n; p.gratedC.bxdf.base.etaK = k; p.gratedC.bxdf.base.extEta = extEta;
                        p.gratedC.bxdf.grating = ref::DiffractionGrating(dir, rsd, R(ip), q, lobes, gtype, mult); }
            else { p.gratedD.bxdf.base.albedo = R(col); p.gratedD.bxdf.base.gfs = ref::FractalSurface(rough); p.gratedD.bxdf.grating = ref::DiffractionGrating(dir, rsd, R(ip), q, lobes, gtype, mult); }
            break; }
        default: { p.o.type = PLTPT_MATERIAL_MULTILAYERED_STACK;
            const orc::float2 ior(r.range(1.2f, 1.6f), r.range(1.8f, 2.6f)), th(r.range(0.05f, 0.3f), r.range(0.05f, 0.3f)); const uint32_t layers = 1u + uint32_t(r.u() * 20.f);
            p.o.stack.albedo = col; p.o.stack.extEta = extEta; p.o.stack.ior = ior; p.o.stack.thickness = th; p.o.stack.layers = layers;
            p.stack.bxdf.albedo = R(col); p.stack.bxdf.extEta = extEta; p.stack.bxdf.ior = R(ior); p.stack.bxdf.thickness = R(th); p.stack.bxdf.layers = layers; break; }
    }
}
#define REF_DISPATCH(p, CALL)                                                                                                           \
    switch ((p).kind) { case 0: (p).diffuse.CALL; break; case 1: (p).conductor.CALL; break; case 2: (p).dielectric.CALL; break; case 3: (p).opaque.CALL; break; case 4: (p).thin.CALL; break; \
                        case 5: (p).coatedC.CALL; break; case 6: (p).coatedD.CALL; break; case 7: (p).gratedC.CALL; break; case 8: (p).gratedD.CALL; break; default: (p).stack.CALL; break; }
#define REF_DISPATCH_R(p, OUT, CALL)                                                                                                    \
    switch ((p).kind) { case 0: OUT = (p).diffuse.CALL; break; case 1: OUT = (p).conductor.CALL; break; case 2: OUT = (p).dielectric.CALL; break; case 3: OUT = (p).opaque.CALL; break; case 4: OUT = (p).thin.CALL; break; \
                        case 5: OUT = (p).coatedC.CALL; break; case 6: OUT = (p).coatedD.CALL; break; case 7: OUT = (p).gratedC.CALL; break; case 8: OUT = (p).gratedD.CALL; break; default: OUT = (p).stack.CALL; break; }
}  // namespace

static void ref_compare_frames(int N, Rnd& r, PltptSpectralProfile* prof);
static void ref_compare_more(int N, Rnd& r) {
    // ---- spectral profiles: two tabulated "conductors" (n, k tables, no icdf) and one emitter (pdf + icdf), random but well formed
    static PltptSpectralProfile prof[5];
    static ref::SpectralProfile rprof[5];
    for (int i = 0; i < 5; ++i) {
        PltptSpectralProfile& p = prof[i];
        p.bins = i == 4 ? 128u : 20u + uint32_t(r.u() * 100.f);
        p.minWavelength = i == 4 ? 380.f : r.range(300.f, 400.f); p.maxWavelength = i == 4 ? 720.f : r.range(700.f, 900.f);
        p.intensity = i == 4 ? r.range(.5f, 5.f) : 1.f;
        float acc = 0.f;
        for (uint32_t b = 0; b < 128; ++b) { p.pdf[b] = b < p.bins ? (i % 2 == 0 ? r.range(0.1f, 1.6f) : r.range(1.5f, 6.f)) : 0.f; acc += p.pdf[b]; }
        float c = 0.f;
        for (uint32_t b = 0; b < 128; ++b) { c += p.pdf[b] / acc; p.icdf[b] = b < p.bins ? std::fmin(c, 1.f) : 1.f; }
        p.rgb[0] = p.rgb[1] = p.rgb[2] = 1.f;
        ref::SpectralProfile& q = rprof[i];
        q.rgb = ref::float3(1.f); q.intensity = p.intensity; q.bins = p.bins; q.minWavelength = p.minWavelength; q.maxWavelength = p.maxWavelength;
        for (int b = 0; b < 128; ++b) { q.pdf[b] = p.pdf[b]; q.icdf[b] = p.icdf[b]; }
    }
    static orc::ProfileTable table;
    table.profiles = prof; table.count = 5;
    ref::gScene.spectralProfiles = rprof;
    for (int i = 0; i < N; ++i) {
        const uint32_t id = uint32_t(r.u() * 4.999f);
        const float l = r.range(350.f, 760.f);
        const orc::float4 wl = r.wavelengths();
        { Cmp c("Scene.evalSpectrum"); c.add(table.evalSpectrum(id, l), ref::gScene.evalSpectrum(id, l)); c.add(table.evalSpectrum(id, wl), ref::gScene.evalSpectrum(id, R(wl)));
          c.add(table.evalSpectrum(PLTPT_NO_SPECTRAL_PROFILE, l), ref::gScene.evalSpectrum(ref::BasicMaterialData::no_spectral_profile, l)); }
        { Cmp c("Scene.sampleSpectrum"); const bool imp = r.u() < .5f; orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 3u); RefSG g1{g0};
          orc::float4 w0; ref::float4 w1; const float hero = r.range(380.f, 720.f);
          c.add(table.sampleSpectrum(4u, hero, g0, imp, w0), ref::gScene.sampleSpectrum(4u, hero, g1, imp, w1)); c.add(w0, w1); }
        { Cmp c("SpectrumSampler.hwssWavelengths"); const float hero = r.range(380.f, 720.f); c.add(orc::SpectrumSampler::hwssWavelengths(hero), ref::SpectrumSampler::hwssWavelengths(hero));
          const float u = r.u(); c.add(orc::SpectrumSampler::sampleOne(u), ref::SpectrumSampler::sampleOne(u)); }
    }
    // ---- DifferentialShadingGeometry.slang: reflect / refract derivatives
    for (int i = 0; i < N; ++i) {
        Cmp c("DifferentialShadingGeometry.reflect_refractDerivatives");
        orc::float3pdv wi, n;
        n.v = r.dir(); n.ddu = r.f3(-1.f, 1.f); n.ddv = r.f3(-1.f, 1.f);
        wi.v = r.dir(); wi.ddu = r.f3(-1.f, 1.f); wi.ddv = r.f3(-1.f, 1.f);
        addK(c, orc::reflectDerivatives(wi, n), ref::reflectDerivatives(R(wi), R(n)));
        const float eta = r.range(1.1f, 2.f), de = r.range(-1e-3f, 1e-3f);
        addK(c, orc::refractDerivatives(wi, n, eta, de, -de), ref::refractDerivatives(R(wi), R(n), eta, de, -de));
    }
    // ---- DiffractionGrating.slang
    for (int i = 0; i < N; ++i) {
        const orc::ShadingData sd = randomShadingData(r);
        ref::ShadingData rsd = R(sd);
        const float dir = r.range(-3.f, 3.f), q = r.range(0.02f, 0.5f), mult = r.range(.5f, 1.5f);
        const orc::float2 ip(1.f / r.range(1.f, 6.f), r.u() < .5f ? 0.f : 1.f / r.range(1.f, 6.f));
        const uint32_t lobes = 1u + uint32_t(r.u() * 7.99f), gtype = uint32_t(r.u() * 2.99f) | (r.u() < .2f ? 0x10u : 0u);
        const orc::DiffractionGrating a(dir, sd, ip, q, lobes, gtype, mult); ref::DiffractionGrating b(dir, rsd, R(ip), q, lobes, gtype, mult);
        const orc::float3 wi = r.dirUp(), hint = r.dirUp();
        const orc::float4 k = r.ks();
        const float lambda = r.range(0.38f, 0.72f);
        { Cmp c("DiffractionGrating.setup_alpha"); c.add(a.gratingX, b.gratingX); c.add(a.alpha(wi, k.x), b.alpha(R(wi), k.x)); c.add(a.alpha(wi, k), b.alpha(R(wi), R(k))); c.add(a.is1DGrating(), b.is1DGrating());
          const orc::ShadingData g0 = a.localGratingShadingFrame(sd); ref::ShadingData g1 = b.localGratingShadingFrame(rsd); c.add(g0.T, g1.T); c.add(g0.B, g1.B); c.add(g0.N, g1.N); }
        orc::int2 lb; lb.x = int(r.range(-7.99f, 7.99f)); lb.y = ip.y == 0.f ? 0 : int(r.range(-3.99f, 3.99f));
        const ref::int2 rlb(lb.x, lb.y);
        { Cmp c("DiffractionGrating.lobeIntensity_diffract"); c.add(a.lobeIntensity(lb, wi, lambda), b.lobeIntensity(rlb, R(wi), lambda)); c.add(a.diffract(wi, lb, lambda), b.diffract(R(wi), rlb, lambda)); }
        { Cmp c("DiffractionGrating.closestLobeHint"); const float al = r.u() < .1f ? 0.f : (r.u() < .1f ? 1.f : r.u()); const orc::int2 h0 = a.closestLobeHint(wi, hint, al, lambda); const ref::int2 h1 = b.closestLobeHint(R(wi), R(hint), al, lambda);
          c.add(h0.x, h1.x); c.add(h0.y, h1.y); }
        { Cmp c("DiffractionGrating.sampleLobe"); orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 5u); RefSG g1{g0}; orc::float2 p0; ref::float2 p1; ref::float3 woUnused;
          const orc::int2 s0 = a.sampleLobe(wi, p0, lambda, g0); const ref::int2 s1 = b.sampleLobe(R(wi), woUnused, p1, lambda, g1); c.add(s0.x, s1.x); c.add(s0.y, s1.y); c.add(p0, p1); }
    }
    // ---- the ten BSDFs through their IMaterialInstance wrappers
    for (int i = 0; i < N; ++i) {
        const int kind = i % 10;
        const orc::ShadingData sd = randomShadingData(r);
        ref::ShadingData rsd = R(sd);
        static Pair p;
        makePair(p, kind, r, &table, sd);
        const float wavelength = r.range(380.f, 720.f);
        // sample(): the generator streams must stay in step as well
        orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 11u); RefSG g1{g0};
        orc::BSDFSample s0; ref::BSDFSample s1;
        // The grated BSDFs leave the `out float pdf` of their sample() unwritten when the lobe mask excludes either the delta or the scatter
        // lobes and no hint is given (PLTDiffractionGratedConductorBSDF.slang:132-141) and then multiply it -- undefined in the reference,
        // never reached by the integrator (its masks are All, or Delta with a hint).  The oracle defines the start value as 1; same here.
        s1.pdf = 1.f;
        const uint32_t mask = r.u() < .7f ? uint32_t(orc::L_All) : kLobeChoices[uint32_t(r.u() * 8.99f)];
        orc::float3 hint = r.u() < .5f ? orc::float3(0.f) : r.dir();
        const bool v0 = p.o.sample(sd, g0, s0, wavelength, mask, hint);
        bool v1 = false;
        REF_DISPATCH_R(p, v1, sample(rsd, g1, s1, wavelength, mask, R(hint)));
        { Cmp c(std::string(kKindNames[kind]) + ".sample"); c.add(v0, v1); c.add(g0.next(), g1.next());
          if (v0 && v1) { c.add(s0.wo, s1.wo); c.add(s0.pdf, s1.pdf); c.add(s0.lobe, s1.lobe); c.add(s0.weight, s1.weight); } }
        // eval(): a sampled direction with its own lobe (the case the integrator evaluates), or a random direction and lobe
        const bool useSampled = v0 && v1 && r.u() < .6f;
        const orc::float3 toEye = useSampled ? s0.wo : sd.fromLocal(r.dir());
        const uint32_t lobe = useSampled ? s0.lobe : kLobeChoices[uint32_t(r.u() * 8.99f)];
        ref::PLTBeam rb;
        orc::PLTBeam b = randomBeam(r, rb);
        b.r = -sd.V; b.t = orc::perp_stark(b.r); b.propagate(sd.posW);   // arrives along -V, like forwardUnwindStep
        if (r.u() < .2f) b.hwss_comps = 1;
        rb = R(b);
        { Cmp c(std::string(kKindNames[kind]) + ".lambdaPdf"); ref::float4 l1; REF_DISPATCH_R(p, l1, lambdaPdf(rsd, R(toEye), lobe, rb, rb.wavelengths)); c.add(p.o.lambdaPdf(sd, toEye, lobe, b, b.wavelengths), l1); }
        p.o.eval(sd, toEye, lobe, b);
        REF_DISPATCH(p, eval(rsd, R(toEye), lobe, rb));
        { Cmp c(std::string(kKindNames[kind]) + ".eval"); c.add(b, rb); }
        { Cmp c(std::string(kKindNames[kind]) + ".pdf"); float q1 = 0.f; const orc::float3 toLight = sd.fromLocal(r.dirUp()); REF_DISPATCH_R(p, q1, pdf(rsd, R(toLight), wavelength)); c.add(p.o.pdf(sd, toLight, wavelength), q1); }
    }
    // ---- LightHelpers.slang: the six analytic light samplers behind sampleLight() (LightData.slang is transpiled too: same struct layout)
    {
        static orc::Scene osc;
        osc.profiles = table;
        orc::Integrator integ;
        integ.scene = &osc;
        const char* kLightNames[6] = {"Point", "Directional", "Distant", "Rect", "Disc", "Sphere"};
        for (int i = 0; i < N; ++i) {
            PltptLightData L{};
            ref::LightData Lr;
            L.type = uint32_t(i % 6);
            const orc::float3 pos = r.f3(-3.f, 3.f), dir = r.dir();
            for (int k = 0; k < 3; ++k) { L.posW[k] = pos[k]; L.dirW[k] = dir[k]; }
            L.openingAngle = r.u() < .4f ? orc::kPi : r.range(0.2f, 1.5f);
            L.cosOpeningAngle = std::cos(L.openingAngle);
            L.penumbraAngle = r.u() < .5f ? 0.f : r.range(0.01f, 0.3f);
            L.cosSubtendedAngle = r.u() < .3f ? 0.9999893f : std::cos(r.range(0.001f, 0.4f));
            L.intensitySpectrumId = 4u;
            L.surfaceArea = r.range(0.05f, 8.f);
            // local-to-world: rotation (random orthonormal frame) * anisotropic scale + translation; IT = inverse transpose of it
            const orc::float3 ez = r.dir(), ex = orc::perp_stark(ez), ey = orc::cross(ez, ex);
            const orc::float3 sc = r.f3(0.2f, 2.f);
            const orc::float3 cols[3] = {ex * sc.x, ey * sc.y, ez * sc.z}, colsIT[3] = {ex / sc.x, ey / sc.y, ez / sc.z};
            for (int row = 0; row < 3; ++row) {
                for (int c = 0; c < 3; ++c) { L.transMat[4 * row + c] = cols[c][row]; L.transMatIT[4 * row + c] = colsIT[c][row]; }
                L.transMat[4 * row + 3] = pos[row]; L.transMatIT[4 * row + 3] = 0.f;
            }
            L.transMat[15] = 1.f; L.transMatIT[15] = 1.f;
            Lr.posW = R(pos); Lr.type = L.type; Lr.dirW = R(dir); Lr.openingAngle = L.openingAngle; Lr.cosOpeningAngle = L.cosOpeningAngle; Lr.cosSubtendedAngle = L.cosSubtendedAngle;
            Lr.penumbraAngle = L.penumbraAngle; Lr.intensitySpectrumId = L.intensitySpectrumId; Lr.surfaceArea = L.surfaceArea;
            for (int row = 0; row < 4; ++row) for (int c = 0; c < 4; ++c) { Lr.transMat[row][c] = L.transMat[4 * row + c]; Lr.transMatIT[row][c] = L.transMatIT[4 * row + c]; }
            const orc::float3 shadingPos = r.f3(-4.f, 4.f);
            const float hero = r.range(380.f, 720.f);
            const bool imp = r.u() < .5f;
            orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 17u); RefSG g1{g0};
            orc::Integrator::AnalyticLightSample a;
            ref::AnalyticLightSample b;
            const bool v0 = integ.sampleAnalyticLight(shadingPos, L, g0, hero, imp, a);
            const bool v1 = ref::sampleLight(R(shadingPos), Lr, g1, hero, imp, b);
            Cmp c(std::string("LightHelpers.sample") + kLightNames[L.type]);
            c.add(v0, v1); c.add(g0.next(), g1.next());
            if (v0 && v1) { c.add(a.posW, b.posW); c.add(a.normalW, b.normalW); c.add(a.dir, b.dir); c.add(a.distance, b.distance); c.add(a.Li, b.Li); c.add(a.wavelengths, b.wavelengths); c.add(a.pdf, b.pdf); }
        }
    }
    // ---- LightSampling.slangh (the pass's NEE light sampling) with LightsSamplingUtils.slang, ShadowRays.slangh computeNewRayOrigin and
    // GeometryHelpers.slang computeRayOrigin: light-type selection under every combination of the three switches, and whole analytic-light
    // samples (light choice, radiance scaling, beam sourcing, ray origin, reflection / transmission filter).  The environment-map and
    // emissive-triangle samplers need the texture system / the light collection and stand in as "no sample" on the reference side; the
    // oracle side is never asked for one (those types are switched off in the whole-sample runs).
    {
        static orc::Scene osc;
        osc.profiles = table;
        static PltptLightData lights[4];
        static ref::LightData rlights[4];
        static const float oneTexel[4] = {1.f, 1.f, 1.f, 1.f};
        orc::Integrator integ;
        integ.scene = &osc;
        integ.opt = PltptOptions{};  // every field the light sampling reads is set below
        for (int i = 0; i < N; ++i) {
            { Cmp c("GeometryHelpers.computeRayOrigin");
              const float scale = r.u() < .3f ? 0.05f : (r.u() < .5f ? 3.f : 500.f);
              const orc::float3 pos = r.f3(-scale, scale), nrm = r.dir();
              c.add(orc::computeRayOrigin(pos, nrm), ref::computeRayOrigin(R(pos), R(nrm))); }
            const orc::ShadingData sd = randomShadingData(r);
            ref::ShadingData rsd = R(sd);
            { Cmp c("ShadowRays.computeNewRayOrigin"); const orc::float3 d = r.dir(); c.add(orc::Integrator::computeNewRayOrigin(sd, d), ref::computeNewRayOrigin(rsd, R(d))); }
            // light-type selection: the oracle derives its switches from options AND scene content, like PLTPT.cpp:235-243
            const bool env = r.u() < .5f, emi = r.u() < .5f, ana = r.u() < .5f;
            integ.opt.useEnvLights = 1; integ.opt.useEmissiveLights = 1; integ.opt.useAnalyticLights = 1;
            osc.env.texels = env ? oneTexel : nullptr;
            osc.d.emissiveTriangleCount = emi ? 1u : 0u;
            osc.d.lightCount = ana ? 1u : 0u;
            ref::kUseEnvLight = env; ref::kUseEmissiveLights = emi; ref::kUseAnalyticLights = ana;
            { Cmp c("LightSampling.selectLightType"); orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 19u); RefSG g1{g0};
              uint32_t t0 = 7u; ref::uint t1 = 7u; float p0 = -1.f, p1 = -1.f;
              const bool v0 = integ.selectLightType(t0, p0, g0), v1 = ref::selectLightType(t1, p1, g1);
              c.add(v0, v1); c.add(t0, uint32_t(t1)); if (v0 && v1) c.add(p0, p1); c.add(g0.next(), g1.next()); }
            // whole samples: analytic lights only
            osc.env.texels = nullptr; osc.d.emissiveTriangleCount = 0;
            ref::kUseEnvLight = false; ref::kUseEmissiveLights = false; ref::kUseAnalyticLights = true;
            const uint32_t count = 1u + uint32_t(r.u() * 3.999f);
            for (uint32_t k = 0; k < count; ++k) {
                PltptLightData& L = lights[k];
                L = PltptLightData{};
                L.type = uint32_t(r.u() * 5.999f);
                const orc::float3 pos = r.f3(-3.f, 3.f), dir = r.dir();
                for (int a = 0; a < 3; ++a) { L.posW[a] = pos[a]; L.dirW[a] = dir[a]; }
                L.openingAngle = r.u() < .4f ? orc::kPi : r.range(0.2f, 1.5f);
                L.cosOpeningAngle = std::cos(L.openingAngle);
                L.penumbraAngle = r.u() < .5f ? 0.f : r.range(0.01f, 0.3f);
                L.cosSubtendedAngle = r.u() < .3f ? 0.9999893f : std::cos(r.range(0.001f, 0.4f));
                L.intensitySpectrumId = 4u;
                L.surfaceArea = r.range(0.05f, 8.f);
                const orc::float3 ez = r.dir(), ex = orc::perp_stark(ez), ey = orc::cross(ez, ex), sc = r.f3(0.2f, 2.f);
                const orc::float3 cols[3] = {ex * sc.x, ey * sc.y, ez * sc.z}, colsIT[3] = {ex / sc.x, ey / sc.y, ez / sc.z};
                for (int row = 0; row < 3; ++row) {
                    for (int cc = 0; cc < 3; ++cc) { L.transMat[4 * row + cc] = cols[cc][row]; L.transMatIT[4 * row + cc] = colsIT[cc][row]; }
                    L.transMat[4 * row + 3] = pos[row];
                }
                L.transMat[15] = 1.f; L.transMatIT[15] = 1.f;
                ref::LightData& Lr = rlights[k];
                Lr = ref::LightData();
                Lr.posW = R(pos); Lr.type = L.type; Lr.dirW = R(dir); Lr.openingAngle = L.openingAngle; Lr.cosOpeningAngle = L.cosOpeningAngle; Lr.cosSubtendedAngle = L.cosSubtendedAngle;
                Lr.penumbraAngle = L.penumbraAngle; Lr.intensitySpectrumId = L.intensitySpectrumId; Lr.surfaceArea = L.surfaceArea;
                for (int row = 0; row < 4; ++row) for (int cc = 0; cc < 4; ++cc) { Lr.transMat[row][cc] = L.transMat[4 * row + cc]; Lr.transMatIT[row][cc] = L.transMatIT[4 * row + cc]; }
            }
            osc.d.lights = lights; osc.d.lightCount = count;
            ref::gScene.lights = rlights; ref::gScene.lightCount = count;
            const float area = r.range(0.01f, 4.f), omega = r.range(0.001f, 0.1f);
            integ.opt.sourcingAreaFromEmissiveGeometry = area; integ.opt.sourcingMaxBeamOmega = omega;
            ref::kSourcingAreaFromEmissiveGeometry = area; ref::kSourcingMaxBeamOmega = omega;
            const bool perTile = r.u() < .5f, reflOk = r.u() < .8f, transOk = r.u() < .8f, imp = r.u() < .5f;
            integ.opt.doNEE = 1; integ.opt.NEEUsePerTileSG = perTile ? 1 : 0;
            ref::kNEEUsePerTileSelectorSG = perTile;
            const float hero = r.range(380.f, 720.f);
            orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 23u), t0(uint32_t(i & 0xff), uint32_t(i >> 8), 29u); RefSG g1{g0}, t1{t0};
            orc::Integrator::LightSample a;
            ref::LightSample b;
            const bool v0 = integ.generateLightSample(sd, reflOk, transOk, g0, t0, hero, imp, a);
            const bool v1 = ref::generateLightSample(rsd, reflOk, transOk, g1, t1, hero, imp, b);
            Cmp c("LightSampling.generateLightSample_analytic");
            c.add(v0, v1); c.add(g0.next(), g1.next()); c.add(t0.next(), t1.next());
            if (v0 && v1) { c.add(a.pdf, b.pdf); c.add(a.origin, b.origin); c.add(a.distance, b.distance); c.add(a.dir, b.dir); c.add(a.lightType, uint32_t(b.lightType)); c.add(a.beam, b.beam); }
        }
        osc.d.lights = nullptr; osc.d.lightCount = 0;
    }
    // ---- pltpt_solve.rt.slang: `measure` (propagation to the camera, camera polariser, spectral MIS, XYZ fit -> RGB, accumulation) and MISWeight
    {
        static orc::Scene osc;
        orc::Integrator integ;
        integ.scene = &osc;
        integ.opt = PltptOptions{};
        for (int i = 0; i < N; ++i) {
            const orc::float3 camPos = r.f3(-5.f, 5.f), w = r.dir() * r.range(0.5f, 3.f);
            const orc::float3 u = orc::perp_stark(orc::normalize(w)) * r.range(0.3f, 2.f), v = orc::cross(orc::normalize(w), orc::normalize(u)) * r.range(0.3f, 2.f);
            const float pol = r.u() < .4f ? 0.f : r.u(), theta = r.range(-3.2f, 3.2f);
            for (int a = 0; a < 3; ++a) { integ.cam.posW[a] = camPos[a]; integ.cam.cameraU[a] = u[a]; integ.cam.cameraV[a] = v[a]; integ.cam.cameraW[a] = w[a]; }
            integ.cam.polarizer = pol; integ.cam.polarizerTheta = theta;
            ref::gScene.camera.data.posW = R(camPos); ref::gScene.camera.data.cameraU = R(u); ref::gScene.camera.data.cameraV = R(v); ref::gScene.camera.data.cameraW = R(w);
            ref::gScene.camera.data.polarizer = pol; ref::gScene.camera.data.polarizerTheta = theta;
            const bool direct = r.u() < .8f, mis = r.u() < .6f;
            integ.opt.useDirectLights = direct; integ.opt.HWSSDoMIS = mis; integ.opt.debugView = 0;
            ref::kUseDirectLights = direct; ref::mHWSSDoMIS = mis;
            ref::PLTBeam rb;
            orc::PLTBeam b = randomBeam(r, rb);
            b.spectralPDFs = r.f4(0.f, 2.f);
            if (r.u() < .1f) b.spectralPDFs = orc::float4(0.f);
            if (r.u() < .2f) b.hwss_comps = 1;
            rb = R(b);
            const int bounceIdx = int(r.u() * 8.99f);
            const float contribution = r.range(0.01f, 20.f);
            orc::Integrator::PLTSolve s0; ref::PLTSolve s1;
            const orc::float3 start = r.f3(0.f, 2.f);
            s0.Lrgb = start; s1.Lrgb = R(start);
            integ.measure(s0, bounceIdx, contribution, b);
            ref::measure(s1, bounceIdx, contribution, rb);
            { Cmp c("pltpt_solve.measure"); c.add(s0.Lrgb, s1.Lrgb); c.add(b, rb); }
            { Cmp c("pltpt_solve.MISWeight"); const float p0 = r.range(0.f, 50.f), p1 = r.range(1e-6f, 50.f); c.add(orc::Integrator::MISWeight(p0, p1), ref::MISWeight(p0, p1)); }
        }
    }
    // ---- packed scene records as the path decodes them: snorm / octahedral normals (FormatConversion.slang, PackedFormats.slang,
    // MathHelpers.slang oct_to_ndir_snorm), emissive triangles (LightCollectionShared.slang PackedEmissiveTriangle::unpack and the two
    // interpolators), vertices (SceneTypes.slang PackedStaticVertexData::unpack)
    {
        static orc::Scene osc;
        orc::Integrator integ;
        integ.scene = &osc;
        static PltptEmissiveTriangle etri;
        static PltptPackedVertex vert;
        osc.d.emissiveTriangles = &etri; osc.d.emissiveTriangleCount = 1;
        osc.d.vertices = &vert; osc.d.vertexCount = 1;
        auto halfBits = [&](float lo, float hi) { return ref::f32tof16(r.range(lo, hi)); };
        for (int i = 0; i < N; ++i) {
            const uint32_t packed = r.g(), edge = (r.u() < .1f) ? (r.u() < .5f ? 0x80008000u : 0x7fff8001u) : packed;
            { Cmp c("PackedFormats.decodeNormal2x16"); c.add(orc::decodeNormal2x16(edge), ref::decodeNormal2x16(edge)); c.add(orc::unpackSnorm16(packed & 0xffffu), ref::unpackSnorm16(packed & 0xffffu));
              const orc::float2 o = r.f2(-1.f, 1.f); c.add(orc::oct_to_ndir_snorm(o), ref::oct_to_ndir_snorm(R(o))); }
            { Cmp c("LightCollectionShared.PackedEmissiveTriangle_unpack");
              ref::PackedEmissiveTriangle pt;
              for (int k = 0; k < 3; ++k) {
                  const orc::float3 pw = r.f3(-10.f, 10.f);
                  const uint32_t uv = halfBits(-2.f, 3.f) | (halfBits(-2.f, 3.f) << 16);
                  float uvf; std::memcpy(&uvf, &uv, 4);
                  for (int a = 0; a < 3; ++a) etri.posAndTexCoords[k][a] = pw[a];
                  etri.posAndTexCoords[k][3] = uvf;
                  pt.posAndTexCoords[k] = ref::float4(pw.x, pw.y, pw.z, uvf);
              }
              const float area = r.range(1e-4f, 20.f);
              etri.normal = r.g(); std::memcpy(&etri.area, &area, 4); etri.materialID = r.g() & 0xffffu; etri.lightIdx = r.g() & 0xffu;
              pt.normal = etri.normal; pt.area = etri.area; pt.materialID = etri.materialID; pt.lightIdx = etri.lightIdx;
              const orc::Integrator::EmissiveTriangle a = integ.getEmissiveTriangle(0);
              ref::EmissiveTriangle b = pt.unpack();
              for (int k = 0; k < 3; ++k) { c.add(a.posW[k], b.posW[k]); c.add(a.texCoords[k], b.texCoords[k]); }
              c.add(a.normal, b.normal); c.add(a.area, b.area); c.add(a.materialID, uint32_t(b.materialID));
              const orc::float3 bary = orc::sample_triangle(r.f2(0.f, 1.f));
              c.add(a.posW[0] * bary[0] + a.posW[1] * bary[1] + a.posW[2] * bary[2], b.getPosition(R(bary)));          // as sampleTriangle interpolates (plt_integrator.h)
              c.add(a.texCoords[0] * bary[0] + a.texCoords[1] * bary[1] + a.texCoords[2] * bary[2], b.getTexCoord(R(bary))); }
            { Cmp c("SceneTypes.PackedStaticVertexData_unpack");
              const orc::float3 pos = r.f3(-50.f, 50.f); const orc::float2 tc = r.f2(-4.f, 4.f);
              const uint32_t w0 = halfBits(-1.f, 1.f) | (halfBits(-1.f, 1.f) << 16);
              const float sgn = r.u() < .2f ? 0.f : (r.u() < .5f ? -1.f : 1.f);
              const uint32_t w1 = halfBits(-1.f, 1.f) | (ref::f32tof16(sgn) << 16), w2 = r.g();
              for (int a = 0; a < 3; ++a) vert.position[a] = pos[a];
              vert.texCrd[0] = tc.x; vert.texCrd[1] = tc.y;
              vert.packedNormalTangentCurveRadius[0] = w0; vert.packedNormalTangentCurveRadius[1] = w1; vert.packedNormalTangentCurveRadius[2] = w2;
              ref::PackedStaticVertexData pv;
              pv.position = R(pos); pv.texCrd = R(tc);
              pv.packedNormalTangentCurveRadius = ref::float3(ref::asfloat(ref::uint(w0)), ref::asfloat(ref::uint(w1)), ref::asfloat(ref::uint(w2)));
              const orc::StaticVertexData a = osc.getVertex(0);
              ref::StaticVertexData b = pv.unpack();
              c.add(a.position, b.position); c.add(a.normal, b.normal); c.add(a.tangent, b.tangent); c.add(a.texCrd, b.texCrd); }
        }
        osc.d.emissiveTriangles = nullptr; osc.d.emissiveTriangleCount = 0; osc.d.vertices = nullptr; osc.d.vertexCount = 0;
    }
    // ---- emissive-triangle sampling: EmissiveLightSamplerHelpers.slang (materialEvalEmissive, sampleTriangle, evalTrianglePdf),
    // EmissiveUniformSampler.slang and EmissivePowerSampler.slang (sampleLight, evalPdf) over LightCollection.slang, on small random light
    // collections (packed triangles, flux table, active list, alias table) with untextured emitters
    {
        static orc::Scene osc;
        osc.profiles = table;
        orc::Integrator integ;
        integ.scene = &osc;
        integ.opt = PltptOptions{};
        constexpr uint32_t kMax = 6;
        static PltptEmissiveTriangle tris[kMax];
        static PltptEmissiveFlux flux[kMax];
        static uint32_t active[kMax], alias[2 * kMax];
        static ref::uint2 ralias[kMax];
        static PltptMaterialBlob mats[2];
        static ref::PackedEmissiveTriangle rtris[kMax];
        static ref::EmissiveFlux rflux[kMax];
        mats[0] = PltptMaterialBlob{}; mats[1] = PltptMaterialBlob{};
        mats[0].header[0] = 1u << 31; mats[0].emissionSpectralId = 4u;                             // emissive basic material
        mats[1].header[0] = 1u << 31; mats[1].emissionSpectralId = PLTPT_NO_SPECTRAL_PROFILE;      // basic, not emissive: Le stays 0
        osc.d.materialData = mats; osc.d.materialCount = 2;
        static_assert(sizeof(ref::MaterialDataBlob) == sizeof(PltptMaterialBlob) && sizeof(PltptMaterialBlob) == 128, "both sides read the same 128-byte material records");
        ref::gScene.materials.blobs = reinterpret_cast<const ref::MaterialDataBlob*>(mats);
        ref::EmissiveUniformSampler& uniformSampler = ref::gUniformSampler;
        ref::EmissivePowerSampler& powerSampler = ref::gPowerSampler;
        static PltptLightData pointLight;
        static ref::LightData rpointLight;
        for (int i = 0; i < N; ++i) {
            const uint32_t n = 1u + uint32_t(r.u() * (kMax - 0.001f));
            float total = 0.f;
            for (uint32_t t = 0; t < n; ++t) {
                const orc::float3 v0 = r.f3(-3.f, 3.f), v1 = v0 + r.f3(-1.f, 1.f), v2 = v0 + r.f3(-1.f, 1.f);
                const orc::float3 vs[3] = {v0, v1, v2};
                for (int k = 0; k < 3; ++k) {
                    const uint32_t uv = ref::f32tof16(r.u()) | (ref::f32tof16(r.u()) << 16);
                    float uvf; std::memcpy(&uvf, &uv, 4);
                    for (int a = 0; a < 3; ++a) tris[t].posAndTexCoords[k][a] = vs[k][a];
                    tris[t].posAndTexCoords[k][3] = uvf;
                    rtris[t].posAndTexCoords[k] = ref::float4(vs[k].x, vs[k].y, vs[k].z, uvf);
                }
                const float area = 0.5f * orc::length(orc::cross(v1 - v0, v2 - v0));
                tris[t].normal = r.g(); std::memcpy(&tris[t].area, &area, 4); tris[t].materialID = r.u() < .85f ? 0u : 1u; tris[t].lightIdx = 0;
                rtris[t].normal = tris[t].normal; rtris[t].area = tris[t].area; rtris[t].materialID = tris[t].materialID; rtris[t].lightIdx = 0;
                flux[t].flux = r.range(0.01f, 5.f); total += flux[t].flux;
                rflux[t].flux = flux[t].flux;
                active[t] = n - 1u - t;  // a permutation: the active list is an indirection
                const uint32_t above = uint32_t(r.u() * (n - 0.001f)), below = uint32_t(r.u() * (n - 0.001f));
                const uint32_t px = (ref::f32tof16(r.u()) << 16) | ((above >> 8) & 0xffffu), py = ((above & 0xffu) << 24) | (below & 0xffffffu);
                alias[2 * t] = px; alias[2 * t + 1] = py; ralias[t] = ref::uint2(px, py);
            }
            const float invSum = 1.f / total;
            osc.d.emissiveTriangles = tris; osc.d.emissiveTriangleCount = n; osc.d.emissiveFlux = flux; osc.d.activeTriangles = active; osc.d.activeTriangleCount = n;
            osc.d.triangleAliasTable = alias; osc.d.invWeightsSum = invSum;
            ref::LightCollection& lc = ref::gScene.lightCollection;
            lc.triangleCount = n; lc.activeTriangleCount = n; lc.triangleData.p = rtris; lc.activeTriangles.p = active; lc.fluxData.p = rflux;
            powerSampler._emissivePower.invWeightsSum = invSum; powerSampler._emissivePower.triangleAliasTable.p = ralias;
            const bool power = r.u() < .5f, imp = r.u() < .5f;
            integ.opt.emissiveSampler = power ? PLTPT_EMISSIVE_SAMPLER_POWER : PLTPT_EMISSIVE_SAMPLER_UNIFORM;
            const orc::float3 shadingPos = r.f3(-4.f, 4.f), shadingN = r.dir();
            const float hero = r.range(380.f, 720.f);
            orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 31u); RefSG g1{g0};
            orc::Integrator::TriangleLightSample a;
            ref::TriangleLightSample b;
            const bool v0 = integ.emissiveSampleLight(shadingPos, g0, hero, imp, a);
            const bool v1 = power ? powerSampler.sampleLight(R(shadingPos), R(shadingN), false, g1, hero, imp, b) : uniformSampler.sampleLight(R(shadingPos), R(shadingN), false, g1, hero, imp, b);
            { Cmp c(power ? "EmissivePowerSampler.sampleLight" : "EmissiveUniformSampler.sampleLight");
              c.add(v0, v1); c.add(g0.next(), g1.next()); c.add(a.triangleIndex, uint32_t(b.triangleIndex));
              if (v0 && v1) { c.add(a.posW, b.posW); c.add(a.normalW, b.normalW); c.add(a.dir, b.dir); c.add(a.distance, b.distance); c.add(a.Le, b.Le); c.add(a.wavelengths, b.wavelengths); c.add(a.pdf, b.pdf); } }
            { Cmp c(power ? "EmissivePowerSampler.evalPdf" : "EmissiveUniformSampler.evalPdf");
              ref::TriangleLightHit hit;
              const uint32_t ti = r.u() < .05f ? 0xffffffffu : uint32_t(r.u() * (n - 0.001f));
              const orc::float3 hp = r.f3(-3.f, 3.f), hn = r.dir();
              hit.triangleIndex = ti; hit.posW = R(hp); hit.normalW = R(hn);
              const float p0 = integ.emissiveEvalPdf(shadingPos, ti, hp, hn);
              const float p1 = power ? powerSampler.evalPdf(R(shadingPos), R(shadingN), false, hit) : uniformSampler.evalPdf(R(shadingPos), R(shadingN), false, hit);
              c.add(p0, p1); }
            // the pass's whole NEE sample with emissive triangles in the mix (LightSampling.slangh generateEmissiveSample + two-way selection
            // with one analytic point light)
            {
                const bool withAnalytic = r.u() < .5f;
                pointLight = PltptLightData{};
                const orc::float3 lp = r.f3(-3.f, 3.f), ld = r.dir();
                for (int k = 0; k < 3; ++k) { pointLight.posW[k] = lp[k]; pointLight.dirW[k] = ld[k]; }
                pointLight.type = 0; pointLight.openingAngle = orc::kPi; pointLight.cosOpeningAngle = -1.f; pointLight.intensitySpectrumId = 4u; pointLight.cosSubtendedAngle = 0.9999893f;
                rpointLight = ref::LightData();
                rpointLight.posW = R(lp); rpointLight.dirW = R(ld); rpointLight.type = 0; rpointLight.openingAngle = pointLight.openingAngle; rpointLight.cosOpeningAngle = -1.f;
                rpointLight.intensitySpectrumId = 4u; rpointLight.cosSubtendedAngle = pointLight.cosSubtendedAngle;
                osc.d.lights = &pointLight; osc.d.lightCount = withAnalytic ? 1u : 0u;
                ref::gScene.lights = &rpointLight; ref::gScene.lightCount = withAnalytic ? 1u : 0u;
                integ.opt.useEmissiveLights = 1; integ.opt.useAnalyticLights = 1; integ.opt.useEnvLights = 1; integ.opt.doNEE = 1; integ.opt.NEEUsePerTileSG = 0;
                ref::kUseEnvLight = false; ref::kUseEmissiveLights = true; ref::kUseAnalyticLights = withAnalytic; ref::kNEEUsePerTileSelectorSG = false;
                const float area = r.range(0.01f, 4.f), omega = r.range(0.001f, 0.1f);
                integ.opt.sourcingAreaFromEmissiveGeometry = area; integ.opt.sourcingMaxBeamOmega = omega;
                ref::kSourcingAreaFromEmissiveGeometry = area; ref::kSourcingMaxBeamOmega = omega;
                ref::gEmissiveKind = power ? 2 : 1;
                const orc::ShadingData sd = randomShadingData(r);
                ref::ShadingData rsd = R(sd);
                orc::SampleGenerator h0(uint32_t(i & 0xffff), uint32_t(i >> 16), 37u), t0(uint32_t(i & 0xff), uint32_t(i >> 8), 41u); RefSG h1{h0}, t1{t0};
                orc::Integrator::LightSample la;
                ref::LightSample lb;
                const bool w0 = integ.generateLightSample(sd, true, true, h0, t0, hero, imp, la);
                const bool w1 = ref::generateLightSample(rsd, true, true, h1, t1, hero, imp, lb);
                Cmp c("LightSampling.generateLightSample_emissive");
                c.add(w0, w1); c.add(h0.next(), h1.next());
                if (w0 && w1) { c.add(la.pdf, lb.pdf); c.add(la.origin, lb.origin); c.add(la.distance, lb.distance); c.add(la.dir, lb.dir); c.add(la.lightType, uint32_t(lb.lightType)); c.add(la.beam, lb.beam); }
                ref::gEmissiveKind = 0;
                osc.d.lights = nullptr; osc.d.lightCount = 0; ref::gScene.lightCount = 0;
            }
        }
        osc.d.emissiveTriangles = nullptr; osc.d.emissiveTriangleCount = 0; osc.d.activeTriangleCount = 0; osc.d.materialData = nullptr; osc.d.materialCount = 0;
    }
    // ---- environment map: EnvMap.slang (eval, toLocal / toWorld, lat-long mapping), EnvMapSampler.slang (hierarchical importance sampling,
    // evalPdf) and the pass's generateEnvMapSample.  Texel loads and the two texture filters are services of the texture unit: both sides
    // read the oracle's lat-long texels, its importance-map mips and its fp32 filters (plt_envmap.h)
    {
        static orc::Scene osc;
        osc.profiles = table;
        orc::Integrator integ;
        integ.scene = &osc;
        integ.opt = PltptOptions{};
        constexpr uint32_t W = 32, H = 16;
        static float texels[W * H * 4];
        for (uint32_t t = 0; t < W * H; ++t) {
            const float sun = (t == 5 * W + 20 || t == 5 * W + 21) ? 80.f : 1.f;  // a bright spot, so that the importance map is far from uniform
            texels[4 * t] = sun * r.range(0.05f, 1.f); texels[4 * t + 1] = sun * r.range(0.05f, 1.f); texels[4 * t + 2] = sun * r.range(0.05f, 1.f); texels[4 * t + 3] = 1.f;
        }
        PltptSceneDesc d{};
        d.envMapTexels = texels; d.envMapWidth = W; d.envMapHeight = H;
        const float ang = 0.7f, ca = std::cos(ang), sa = std::sin(ang);
        const float rot[12] = {ca, 0.f, sa, 0.f, 0.f, 1.f, 0.f, 0.f, -sa, 0.f, ca, 0.f}, inv[12] = {ca, 0.f, -sa, 0.f, 0.f, 1.f, 0.f, 0.f, sa, 0.f, ca, 0.f};
        for (int k = 0; k < 12; ++k) { d.envMap.transform[k] = rot[k]; d.envMap.invTransform[k] = inv[k]; }
        d.envMap.tint[0] = 0.9f; d.envMap.tint[1] = 1.f; d.envMap.tint[2] = 0.8f; d.envMap.intensity = 1.7f;
        osc.env.init(d);
        const orc::EnvMap& env = osc.env;
        ref::EnvMap& renv = ref::gScene.envMap;
        for (int row = 0; row < 3; ++row) for (int c = 0; c < 4; ++c) { renv.data.transform[row][c] = rot[4 * row + c]; renv.data.invTransform[row][c] = inv[4 * row + c]; }
        renv.data.tint = ref::float3(0.9f, 1.f, 0.8f); renv.data.intensity = 1.7f;
        renv.envMap.ctx = &env; renv.envMap.width = W; renv.envMap.height = H;
        renv.envMap.loadFn = [](const void* c, int x, int y, int) { const orc::float3 t = static_cast<const orc::EnvMap*>(c)->texel(x, y); return ref::float4(t.x, t.y, t.z, 1.f); };
        renv.envMap.sampleFn = [](const void* c, ref::float2 uv, float) { const orc::float3 t = static_cast<const orc::EnvMap*>(c)->sampleBilinear(orc::float2(uv.x, uv.y)); return ref::float4(t.x, t.y, t.z, 1.f); };
        ref::EnvMapSampler& smp = ref::envMapSampler;
        smp.importanceMap.ctx = &env; smp.importanceMap.width = smp.importanceMap.height = orc::EnvMap::kDimension;
        smp.importanceMap.loadFn = [](const void* c, int x, int y, int mip) { return static_cast<const orc::EnvMap*>(c)->load(uint32_t(x), uint32_t(y), uint32_t(mip)); };
        smp.importanceMap.sampleFn = [](const void* c, ref::float2 uv, float) {  // point sampling, clamp to edge
            const int dim = int(orc::EnvMap::kDimension);
            const int x = std::min(std::max(int(std::floor(uv.x * float(dim))), 0), dim - 1), y = std::min(std::max(int(std::floor(uv.y * float(dim))), 0), dim - 1);
            return static_cast<const orc::EnvMap*>(c)->load(uint32_t(x), uint32_t(y), 0u); };
        smp.importanceInvDim = ref::float2(1.f / float(orc::EnvMap::kDimension));
        smp.importanceBaseMip = env.baseMip;
        // the importance map's base level: the reference's setup pass (EnvMapSamplerSetup.cs.slang main, bound as EnvMapSampler.cpp:101-112 binds it: 512 x 512
        // texels, 8 x 8 samples each) against the oracle's EnvMap::init.  The mip chain above it is the driver's generateMips (a 2 x 2 box filter).
        {
            static float importance[orc::EnvMap::kDimension * orc::EnvMap::kDimension];
            ref::outputDim = ref::uint2(orc::EnvMap::kDimension, orc::EnvMap::kDimension);
            ref::outputDimInSamples = ref::uint2(orc::EnvMap::kDimension * orc::EnvMap::kSamplesX, orc::EnvMap::kDimension * orc::EnvMap::kSamplesY);
            ref::numSamples = ref::uint2(orc::EnvMap::kSamplesX, orc::EnvMap::kSamplesY);
            ref::invSamples = 1.f / (orc::EnvMap::kSamplesX * orc::EnvMap::kSamplesY);
            ref::gEnvMap = renv.envMap;
            ref::gImportanceMap.p = importance; ref::gImportanceMap.width = orc::EnvMap::kDimension;
            const uint32_t step = N >= 50000 ? 1u : 4u;   // every texel in the full run
            Cmp c("EnvMapSamplerSetup.main");
            for (uint32_t y = 0; y < orc::EnvMap::kDimension; y += step) for (uint32_t x = 0; x < orc::EnvMap::kDimension; x += step) {
                ref::envMapSamplerSetup(ref::uint3(x, y, 0u));
                c.add(env.load(x, y, 0u), importance[size_t(y) * orc::EnvMap::kDimension + x]);
            }
        }
        integ.opt.useEnvLights = 1; integ.opt.useEmissiveLights = 1; integ.opt.useAnalyticLights = 1; integ.opt.doNEE = 1;
        ref::kUseEnvLight = true; ref::kUseEmissiveLights = false; ref::kUseAnalyticLights = false; ref::kNEEUsePerTileSelectorSG = false; ref::gEmissiveKind = 0;
        for (int i = 0; i < N; ++i) {
            const orc::float3 dir = r.dir();
            { Cmp c("EnvMap.eval_toLocal_toWorld"); c.add(env.eval(dir), renv.eval(R(dir))); c.add(env.toLocal(dir), renv.toLocal(R(dir))); c.add(env.toWorld(dir), renv.toWorld(R(dir)));
              c.add(orc::world_to_latlong_map(dir), ref::world_to_latlong_map(R(dir))); }
            const orc::float2 rnd = r.f2(0.f, 1.f);
            const float hero = r.range(380.f, 720.f);
            { Cmp c("EnvMapSampler.sample"); orc::EnvMap::Sample a; ref::EnvMapSample b;
              const bool v0 = env.sample(rnd, a), v1 = smp.sample(R(rnd), hero, b);
              c.add(v0, v1); c.add(a.dir, b.dir); c.add(a.pdf, b.pdf);
              const orc::float4 wl = orc::SpectrumSampler::hwssWavelengths(hero);
              c.add(wl, b.wavelengths); c.add(orc::SpectrumUtils::rgbToSpectrum(a.Le, wl), b.Le); }
            { Cmp c("EnvMapSampler.evalPdf"); c.add(env.evalPdf(dir), smp.evalPdf(R(dir))); }
            {
                const float omega = r.range(0.001f, 0.1f);
                integ.opt.sourcingMaxBeamOmega = omega; ref::kSourcingMaxBeamOmega = omega;
                const orc::ShadingData sd = randomShadingData(r);
                ref::ShadingData rsd = R(sd);
                orc::SampleGenerator h0(uint32_t(i & 0xffff), uint32_t(i >> 16), 43u), t0(1u, 2u, 47u); RefSG h1{h0}, t1{t0};
                orc::Integrator::LightSample la;
                ref::LightSample lb;
                const bool w0 = integ.generateLightSample(sd, true, true, h0, t0, hero, false, la);
                const bool w1 = ref::generateLightSample(rsd, true, true, h1, t1, hero, false, lb);
                Cmp c("LightSampling.generateLightSample_env");
                c.add(w0, w1); c.add(h0.next(), h1.next());
                if (w0 && w1) { c.add(la.pdf, lb.pdf); c.add(la.origin, lb.origin); c.add(la.distance, lb.distance); c.add(la.dir, lb.dir); c.add(la.lightType, uint32_t(lb.lightType)); c.add(la.beam, lb.beam); }
            }
        }
        ref::kUseEnvLight = false;
    }
    // ---- the ten PLT material loaders (PLT*Material.slang::setupMaterialInstance over MaterialData.slang / BasicMaterialData.slang /
    // TextureHandle.slang): both sides unpack the SAME random 128-byte material record; the resulting BSDFs are then compared through
    // sample / eval / pdf / lambdaPdf (the BSDF code itself is pinned above, so a difference here is a difference in unpacking)
    {
        static PltptMaterialBlob blob;
        static_assert(sizeof(ref::PLTConductorMaterial) <= 128 && sizeof(ref::MaterialHeader) == 8, "material struct = header + BasicMaterialData inside the 128-byte record");
        ref::gScene.materials.blobs = reinterpret_cast<const ref::MaterialDataBlob*>(&blob);
        ref::gScene.materials.textureFn = nullptr;
        orc::TextureSystem noTextures;
        const uint32_t kTypes[10] = {PLTPT_MATERIAL_DIFFUSE, PLTPT_MATERIAL_CONDUCTOR, PLTPT_MATERIAL_DIELECTRIC, PLTPT_MATERIAL_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_THIN_DIELECTRIC,
                                     PLTPT_MATERIAL_COATED_CONDUCTOR, PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_GRATED_CONDUCTOR, PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC,
                                     PLTPT_MATERIAL_MULTILAYERED_STACK};
        auto half = [&](float lo, float hi) { return uint16_t(ref::f32tof16(r.range(lo, hi))); };
        for (int i = 0; i < N; ++i) {
            const int kind = i % 10;
            blob = PltptMaterialBlob{};
            blob.header[0] = kTypes[kind] | (1u << 31);
            for (int k = 0; k < 4; ++k) { blob.baseColor[k] = half(0.05f, 1.f); blob.specular[k] = half(0.02f, 2.5f); blob.data1[k] = half(0.05f, 3.f); blob.data2[k] = half(0.02f, 2.5f); }
            if (r.u() < .15f) blob.specular[1] = 0;   // roughness 0 where the material keeps it there
            if (r.u() < .15f) blob.data2[1] = 0;
            for (int k = 0; k < 3; ++k) blob.transmission[k] = half(0.2f, 1.f);
            blob.IoR = half(1.f, 2.2f);
            const uint32_t n = uint32_t(r.u() * 1.999f) * 2u;
            blob.iorNSpectralId = n; blob.iorKSpectralId = n + 1u;
            blob.emissionSpectralId = r.u() < .2f ? 4u : PLTPT_NO_SPECTRAL_PROFILE;
            if (kind == 7 || kind == 8) blob.data1[3] = uint16_t(((uint32_t(r.u() * 2.99f) | (r.u() < .2f ? 0x10u : 0u)) << 8) | (1u + uint32_t(r.u() * 7.99f)));  // grating type | lobes, raw bits
            if (kind == 9) blob.data1[3] = uint16_t(1u + uint32_t(r.u() * 11.99f));                                                                             // layer count, raw bits
            const orc::ShadingData sd = randomShadingData(r);
            ref::ShadingData rsd = R(sd);
            static Pair p;
            p.kind = kind;
            p.o = orc::MaterialInstance(blob, sd, &table, noTextures);
            const ref::ExplicitLodTextureSampler lod = {0.f};
#define REF_LOAD(FIELD, TYPE) { ref::TYPE m; std::memcpy(static_cast<void*>(&m), &blob, sizeof m); p.FIELD = m.setupMaterialInstance(ref::gScene.materials, rsd, lod); }
            switch (kind) {
                case 0: REF_LOAD(diffuse, PLTDiffuseMaterial) break;
                case 1: REF_LOAD(conductor, PLTConductorMaterial) break;
                case 2: REF_LOAD(dielectric, PLTDielectricMaterial) break;
                case 3: REF_LOAD(opaque, PLTOpaqueDielectricMaterial) break;
                case 4: REF_LOAD(thin, PLTThinDielectricMaterial) break;
                case 5: REF_LOAD(coatedC, PLTCoatedConductorMaterial) break;
                case 6: REF_LOAD(coatedD, PLTCoatedOpaqueDielectricMaterial) break;
                case 7: REF_LOAD(gratedC, PLTDiffractionGratedConductorMaterial) break;
                case 8: REF_LOAD(gratedD, PLTDiffractionGratedOpaqueDielectricMaterial) break;
                default: REF_LOAD(stack, PLTMultiLayeredStackMaterial) break;
            }
#undef REF_LOAD
            const std::string name = std::string(kKindNames[kind]).replace(std::string(kKindNames[kind]).size() - 4, 4, "Material") + ".setupMaterialInstance";
            Cmp c(name);
            const float wavelength = r.range(380.f, 720.f);
            orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 53u); RefSG g1{g0};
            orc::BSDFSample s0; ref::BSDFSample s1;
            s1.pdf = 1.f;  // see the BSDF loop above (grated BSDFs)
            const bool v0 = p.o.sample(sd, g0, s0, wavelength, uint32_t(orc::L_All), orc::float3(0.f));
            bool v1 = false;
            REF_DISPATCH_R(p, v1, sample(rsd, g1, s1, wavelength, uint32_t(orc::L_All), ref::float3(0.f)));
            c.add(v0, v1); c.add(g0.next(), g1.next());
            if (v0 && v1) { c.add(s0.wo, s1.wo); c.add(s0.pdf, s1.pdf); c.add(s0.lobe, s1.lobe); c.add(s0.weight, s1.weight); }
            const bool useSampled = v0 && v1 && r.u() < .6f;
            const orc::float3 toEye = useSampled ? s0.wo : sd.fromLocal(r.dir());
            const uint32_t lobe = useSampled ? s0.lobe : kLobeChoices[uint32_t(r.u() * 8.99f)];
            ref::PLTBeam rb;
            orc::PLTBeam b = randomBeam(r, rb);
            b.r = -sd.V; b.t = orc::perp_stark(b.r); b.propagate(sd.posW);
            rb = R(b);
            ref::float4 l1; REF_DISPATCH_R(p, l1, lambdaPdf(rsd, R(toEye), lobe, rb, rb.wavelengths)); c.add(p.o.lambdaPdf(sd, toEye, lobe, b, b.wavelengths), l1);
            p.o.eval(sd, toEye, lobe, b);
            REF_DISPATCH(p, eval(rsd, R(toEye), lobe, rb));
            c.add(b, rb);
            float q1 = 0.f; const orc::float3 toLight = sd.fromLocal(r.dirUp()); REF_DISPATCH_R(p, q1, pdf(rsd, R(toLight), wavelength)); c.add(p.o.pdf(sd, toLight, wavelength), q1);
        }
    }
    // ---- shading data at a hit and textured material slots: MaterialFactory.slang:79-131 prepareShadingData, composed here from the
    // transpiled pieces in the reference's order -- ShadingUtils.slang computeTangentSpace (ShadingFrame::orthonormalizeSafe), the loaders'
    // evalOpacity and modifyTangentSpace (applyNormalMap / computeShadingFrameFromNormalMap, RGB and RG maps), the double-sided flip -- and the
    // loaders' setupMaterialInstance on that shading data with Texture-mode handles (the filter is a texture-unit service: the oracle's)
    {
        static orc::Scene osc;
        osc.profiles = table;
        static PltptMaterialBlob blob;
        constexpr int kTex = 3;
        static float texels[kTex][4 * 4 * 4];
        static PltptTextureDesc tex[kTex];
        static PltptSamplerDesc smp[2];
        for (int t = 0; t < kTex; ++t) {
            for (int k = 0; k < 64; ++k) texels[t][k] = r.range(0.05f, 1.f);
            tex[t] = PltptTextureDesc{texels[t], 4u, 4u, PLTPT_TEXTURE_FORMAT_RGBA32_FLOAT, 0u};
        }
        smp[0] = PltptSamplerDesc{PLTPT_FILTER_LINEAR, 0u, 0u, 0u, {0.f, 0.f, 0.f, 0.f}};   // linear, wrap
        smp[1] = PltptSamplerDesc{PLTPT_FILTER_POINT, 2u, 1u, 0u, {0.f, 0.f, 0.f, 0.f}};    // point, clamp / mirror
        osc.textures.textures = tex; osc.textures.textureCount = kTex; osc.textures.samplers = smp; osc.textures.samplerCount = 2;
        osc.d.materialData = &blob; osc.d.materialCount = 1;
        ref::gTextureService = &osc.textures;
        ref::gScene.materials.blobs = reinterpret_cast<const ref::MaterialDataBlob*>(&blob);
        ref::gScene.materials.textureFn = ref::textureServiceFn;
        const uint32_t kTypes[10] = {PLTPT_MATERIAL_DIFFUSE, PLTPT_MATERIAL_CONDUCTOR, PLTPT_MATERIAL_DIELECTRIC, PLTPT_MATERIAL_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_THIN_DIELECTRIC,
                                     PLTPT_MATERIAL_COATED_CONDUCTOR, PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_GRATED_CONDUCTOR, PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC,
                                     PLTPT_MATERIAL_MULTILAYERED_STACK};
        auto half = [&](float lo, float hi) { return uint16_t(ref::f32tof16(r.range(lo, hi))); };
        auto handle = [&]() { return r.u() < .35f ? ((1u << 29) | uint32_t(r.u() * (kTex - 0.001f))) : 0u; };
        for (int i = 0; i < N; ++i) {
            const int kind = i % 10;
            blob = PltptMaterialBlob{};
            blob.header[0] = kTypes[kind] | (1u << 31) | (r.u() < .5f ? (1u << 28) : 0u);                 // basic material, double-sided or not
            blob.header[1] = (r.u() < .5f ? 1u : 0u) << 17;                                                // default texture sampler 0 or 1
            blob.flags = uint32_t(r.u() * 2.999f) << 1;                                                    // NormalMapType None / RGB / RG
            for (int k = 0; k < 4; ++k) { blob.baseColor[k] = half(0.05f, 1.f); blob.specular[k] = half(0.02f, 2.5f); blob.data1[k] = half(0.05f, 3.f); blob.data2[k] = half(0.02f, 2.5f); }
            for (int k = 0; k < 3; ++k) blob.transmission[k] = half(0.2f, 1.f);
            blob.IoR = half(1.f, 2.2f);
            const uint32_t n = uint32_t(r.u() * 1.999f) * 2u;
            blob.iorNSpectralId = n; blob.iorKSpectralId = n + 1u; blob.emissionSpectralId = PLTPT_NO_SPECTRAL_PROFILE;
            if (kind == 7 || kind == 8) blob.data1[3] = uint16_t(((uint32_t(r.u() * 2.99f) | (r.u() < .2f ? 0x10u : 0u)) << 8) | (1u + uint32_t(r.u() * 7.99f)));
            if (kind == 9) blob.data1[3] = uint16_t(1u + uint32_t(r.u() * 11.99f));
            blob.texBaseColor = handle(); blob.texSpecular = handle(); blob.texData1 = handle(); blob.texData2 = handle(); blob.texTransmission = handle(); blob.texNormalMap = handle();
            // a hit: interpolated vertex data (unit normal, tangent valid / zero sign / parallel to the normal), view direction on either side
            orc::VertexData v;
            v.posW = r.f3(-3.f, 3.f); v.normalW = r.dir(); v.texC = r.f2(-1.5f, 2.5f);
            v.faceNormalW = orc::normalize(v.normalW + r.f3(-.3f, .3f));
            const float tsel = r.u();
            const orc::float3 tdir = tsel < .1f ? v.normalW : (tsel < .15f ? orc::float3(0.f) : r.dir());
            v.tangentW = orc::float4(tdir.x, tdir.y, tdir.z, tsel < .8f ? (r.u() < .5f ? 1.f : -1.f) : 0.f);
            const orc::float3 viewDir = r.dir();
            const orc::ShadingData sd = osc.prepareShadingData(v, 0u, viewDir);
            // the reference's sequence (MaterialFactory.slang:79-131)
            ref::VertexData rv;
            rv.posW = R(v.posW); rv.normalW = R(v.normalW); rv.faceNormalW = R(v.faceNormalW); rv.tangentW = R(v.tangentW); rv.texC = R(v.texC);
            ref::ShadingData rsd = {};
            rsd.posW = rv.posW; rsd.uv = rv.texC; rsd.V = R(viewDir); rsd.N = rv.normalW;
            const bool validTangentSpace = ref::computeTangentSpace(rsd, rv.tangentW);
            rsd.tangentW = validTangentSpace ? rv.tangentW : ref::float4(rsd.T, 1.f);
            rsd.faceN = rv.faceNormalW;
            rsd.frontFacing = ref::dot(rsd.V, rsd.faceN) >= 0.f;
            rsd.materialID = 0;
            const ref::ExplicitLodTextureSampler lod = {0.f};
            ref::MaterialHeader header = ref::gScene.materials.getMaterialHeader(0);
            static Pair p;
            p.kind = kind;
#define REF_MATERIAL(FIELD, TYPE) { ref::TYPE m; std::memcpy(static_cast<void*>(&m), &blob, sizeof m); rsd.opacity = m.evalOpacity(ref::gScene.materials, rv, lod); \
                if (validTangentSpace) m.modifyTangentSpace(ref::gScene.materials, lod, rsd); \
                if (!rsd.frontFacing && header.isDoubleSided()) rsd.N = -rsd.N; \
                p.FIELD = m.setupMaterialInstance(ref::gScene.materials, rsd, lod); }
            switch (kind) {
                case 0: REF_MATERIAL(diffuse, PLTDiffuseMaterial) break;
                case 1: REF_MATERIAL(conductor, PLTConductorMaterial) break;
                case 2: REF_MATERIAL(dielectric, PLTDielectricMaterial) break;
                case 3: REF_MATERIAL(opaque, PLTOpaqueDielectricMaterial) break;
                case 4: REF_MATERIAL(thin, PLTThinDielectricMaterial) break;
                case 5: REF_MATERIAL(coatedC, PLTCoatedConductorMaterial) break;
                case 6: REF_MATERIAL(coatedD, PLTCoatedOpaqueDielectricMaterial) break;
                case 7: REF_MATERIAL(gratedC, PLTDiffractionGratedConductorMaterial) break;
                case 8: REF_MATERIAL(gratedD, PLTDiffractionGratedOpaqueDielectricMaterial) break;
                default: REF_MATERIAL(stack, PLTMultiLayeredStackMaterial) break;
            }
#undef REF_MATERIAL
            { Cmp c("MaterialFactory.prepareShadingData");
              c.add(sd.posW, rsd.posW); c.add(sd.V, rsd.V); c.add(sd.N, rsd.N); c.add(sd.T, rsd.T); c.add(sd.B, rsd.B); c.add(sd.tangentW, rsd.tangentW); c.add(sd.faceN, rsd.faceN);
              c.add(sd.uv, rsd.uv); c.add(sd.frontFacing, rsd.frontFacing); c.add(sd.opacity, rsd.opacity); }
            p.o = osc.getMaterialInstance(sd);
            Cmp c("PLTMaterials.setupMaterialInstance_textured");
            const float wavelength = r.range(380.f, 720.f);
            orc::SampleGenerator g0(uint32_t(i & 0xffff), uint32_t(i >> 16), 59u); RefSG g1{g0};
            orc::BSDFSample s0; ref::BSDFSample s1;
            s1.pdf = 1.f;
            const bool v0 = p.o.sample(sd, g0, s0, wavelength, uint32_t(orc::L_All), orc::float3(0.f));
            bool v1 = false;
            REF_DISPATCH_R(p, v1, sample(rsd, g1, s1, wavelength, uint32_t(orc::L_All), ref::float3(0.f)));
            c.add(v0, v1); c.add(g0.next(), g1.next());
            if (v0 && v1) { c.add(s0.wo, s1.wo); c.add(s0.pdf, s1.pdf); c.add(s0.lobe, s1.lobe); c.add(s0.weight, s1.weight); }
            const bool useSampled = v0 && v1 && r.u() < .6f;
            const orc::float3 toEye = useSampled ? s0.wo : sd.fromLocal(r.dir());
            const uint32_t lobe = useSampled ? s0.lobe : kLobeChoices[uint32_t(r.u() * 8.99f)];
            if (kind == 9) continue;   // the multilayer stack's eval has its own relative bound (above); its unpacking is covered by sample() and by the untextured run
            ref::PLTBeam rb;
            orc::PLTBeam b = randomBeam(r, rb);
            b.r = -sd.V; b.t = orc::perp_stark(b.r); b.propagate(sd.posW);
            rb = R(b);
            p.o.eval(sd, toEye, lobe, b);
            REF_DISPATCH(p, eval(rsd, R(toEye), lobe, rb));
            c.add(b, rb);
        }
        ref::gScene.materials.textureFn = nullptr;
        osc.d.materialData = nullptr; osc.d.materialCount = 0;
    }
    // ---- interpolated vertex data at a hit: Scene.slang computeVertexData + getFaceNormalInObjectSpace (barycentric interpolation of the
    // unpacked vertices, object -> world by the instance's matrices, face normal by winding) against the oracle's getVertexData on a
    // one-triangle scene (16- and 32-bit index buffers and non-indexed meshes are the oracle's own fetch: test_gpu_trace / test_ingest)
    {
        static orc::Scene osc;
        static PltptPackedVertex verts[3];
        static PltptGeometryInstance inst;
        static PltptMeshDesc mesh;
        static float world[16], worldIT[16];
        osc.d.vertices = verts; osc.d.vertexCount = 3; osc.d.geometryInstances = &inst; osc.d.instanceCount = 1; osc.d.meshes = &mesh; osc.d.meshCount = 1;
        osc.d.worldMatrices = world; osc.d.inverseTransposeWorldMatrices = worldIT; osc.d.matrixCount = 1;
        auto halfBits = [&](float lo, float hi) { return ref::f32tof16(r.range(lo, hi)); };
        for (int i = 0; i < N; ++i) {
            ref::StaticVertexData rvs[3];
            for (int k = 0; k < 3; ++k) {
                const orc::float3 pos = r.f3(-5.f, 5.f); const orc::float2 tc = r.f2(-2.f, 3.f);
                const float sgn = r.u() < .2f ? 0.f : (r.u() < .5f ? -1.f : 1.f);
                const uint32_t w0 = halfBits(-1.f, 1.f) | (halfBits(-1.f, 1.f) << 16), w1 = halfBits(-1.f, 1.f) | (ref::f32tof16(sgn) << 16), w2 = r.g();
                for (int a = 0; a < 3; ++a) verts[k].position[a] = pos[a];
                verts[k].texCrd[0] = tc.x; verts[k].texCrd[1] = tc.y;
                verts[k].packedNormalTangentCurveRadius[0] = w0; verts[k].packedNormalTangentCurveRadius[1] = w1; verts[k].packedNormalTangentCurveRadius[2] = w2;
                ref::PackedStaticVertexData pv;
                pv.position = R(pos); pv.texCrd = R(tc);
                pv.packedNormalTangentCurveRadius = ref::float3(ref::asfloat(ref::uint(w0)), ref::asfloat(ref::uint(w1)), ref::asfloat(ref::uint(w2)));
                rvs[k] = pv.unpack();
            }
            inst = PltptGeometryInstance{}; mesh = PltptMeshDesc{};
            const bool cw = r.u() < .5f;
            inst.flags = cw ? PLTPT_INSTANCE_FLAG_OBJECT_FRONT_FACE_CW : 0u;
            const orc::float3 ez = r.dir(), ex = orc::perp_stark(ez), ey = orc::cross(ez, ex), sc = r.f3(0.3f, 2.5f), tr = r.f3(-4.f, 4.f);
            const orc::float3 cols[3] = {ex * sc.x, ey * sc.y, ez * sc.z}, colsIT[3] = {ex / sc.x, ey / sc.y, ez / sc.z};
            ref::float4x4 wm; ref::float3x3 wit;
            for (int row = 0; row < 3; ++row) {
                for (int c = 0; c < 3; ++c) { world[4 * row + c] = cols[c][row]; worldIT[4 * row + c] = colsIT[c][row]; wm[row][c] = world[4 * row + c]; wit[row][c] = worldIT[4 * row + c]; }
                world[4 * row + 3] = tr[row]; worldIT[4 * row + 3] = 0.f; wm[row][3] = tr[row];
            }
            for (int c = 0; c < 4; ++c) { world[12 + c] = worldIT[12 + c] = c == 3 ? 1.f : 0.f; wm[3][c] = world[12 + c]; }
            orc::TriangleHit hit;
            hit.instanceID = 0; hit.primitiveIndex = 0;
            const orc::float3 bary = orc::sample_triangle(r.f2(0.f, 1.f));
            hit.barycentrics = orc::float2(bary.y, bary.z);
            const orc::float3 w = hit.getBarycentricWeights();
            const orc::VertexData a = osc.getVertexData(hit);
            ref::VertexData b = ref::gScene.computeVertexData(rvs, wm, wit, cw, R(w));
            { Cmp c("Scene.computeVertexData");
              c.add(a.posW, b.posW); c.add(a.normalW, b.normalW); c.add(a.faceNormalW, b.faceNormalW); c.add(a.tangentW, b.tangentW); c.add(a.texC, b.texC); }
            // DifferentialShadingGeometry.slang differentialShadingGeometry: positions, normals and tangent frame with their uv-derivatives at the
            // hit (the manifold solver's geometry); one run in five with all texture coordinates zero (its fallback parameterisation)
            {
                const bool zeroUV = (i % 5) == 0;
                static ref::PackedStaticVertexData rpacked[3];
                for (int k = 0; k < 3; ++k) {
                    if (zeroUV) { verts[k].texCrd[0] = 0.f; verts[k].texCrd[1] = 0.f; }
                    rpacked[k].position = ref::float3(verts[k].position[0], verts[k].position[1], verts[k].position[2]);
                    rpacked[k].texCrd = ref::float2(verts[k].texCrd[0], verts[k].texCrd[1]);
                    const uint32_t* pk = verts[k].packedNormalTangentCurveRadius;
                    rpacked[k].packedNormalTangentCurveRadius = ref::float3(ref::asfloat(ref::uint(pk[0])), ref::asfloat(ref::uint(pk[1])), ref::asfloat(ref::uint(pk[2])));
                }
                ref::gScene.vertices.p = rpacked;
                ref::gScene.indexFn = [](ref::GeometryInstanceID, ref::uint) { return ref::uint3(0u, 1u, 2u); };
                orc::Integrator integ;
                integ.scene = &osc;
                const orc::DifferentialShadingGeometry g0 = integ.differentialShadingGeometry(hit);
                ref::TriangleHit rhit;
                rhit.instanceID.index = 0; rhit.primitiveIndex = 0; rhit.barycentrics = R(hit.barycentrics);
                ref::DifferentialShadingGeometry g1 = ref::differentialShadingGeometry(ref::gScene, rhit);
                Cmp c("DifferentialShadingGeometry.differentialShadingGeometry");
                c.add(g0.p.v, g1.p.v); c.add(g0.p.ddu, g1.p.ddu); c.add(g0.p.ddv, g1.p.ddv); c.add(g0.n.v, g1.n.v); c.add(g0.n.ddu, g1.n.ddu); c.add(g0.n.ddv, g1.n.ddv);
                c.add(g0.s.v, g1.s.v); c.add(g0.s.ddu, g1.s.ddu); c.add(g0.s.ddv, g1.s.ddv); c.add(g0.t.v, g1.t.v); c.add(g0.t.ddu, g1.t.ddu); c.add(g0.t.ddv, g1.t.ddv);
                c.add(g0.uv, g1.uv); c.add(g0.ngeo, g1.ngeo);
            }
        }
        osc.d.vertices = nullptr; osc.d.vertexCount = 0; osc.d.instanceCount = 0; osc.d.meshCount = 0; osc.d.matrixCount = 0;
    }
    // ---- primary rays and Russian roulette: Camera.slang computeRayPinhole over CameraData.slang (both sides read the same 656-byte camera
    // block, jitter included), pltpt.slangh RR
    {
        static orc::Scene osc;
        orc::Integrator integ;
        integ.scene = &osc;
        integ.opt = PltptOptions{};
        static_assert(sizeof(ref::CameraData) == sizeof(PltptCameraData) && sizeof(PltptCameraData) == 656, "CameraData is the 656-byte block of include/pltpt.h");
        for (int i = 0; i < N; ++i) {
            PltptCameraData cam{};
            float* words = reinterpret_cast<float*>(&cam);
            for (size_t k = 0; k < sizeof(cam) / 4; ++k) words[k] = r.range(-2.f, 2.f);   // every field random: a misplaced one would be picked up
            cam.jitterX = r.range(-.5f, .5f) / 1920.f; cam.jitterY = r.range(-.5f, .5f) / 1080.f;
            integ.cam = cam;
            const uint32_t w = 16u + uint32_t(r.u() * 4000.f), h = 16u + uint32_t(r.u() * 2200.f);
            integ.width = w; integ.height = h;
            std::memcpy(static_cast<void*>(&ref::gScene.camera.data), &cam, sizeof cam);
            const uint32_t px = uint32_t(r.u() * (w - 0.001f)), py = uint32_t(r.u() * (h - 0.001f));
            orc::float3 o, d;
            integ.computeRayPinhole(px, py, o, d);
            ref::Ray ray = ref::gScene.camera.computeRayPinhole(ref::uint2(px, py), ref::uint2(w, h));
            { Cmp c("Camera.computeRayPinhole"); c.add(o, ray.origin); c.add(d, ray.dir); }
            { Cmp c("pltpt.RR"); const float v = r.u(), e = r.range(0.2f, 3.f), u = r.u(); c.add(orc::Integrator::RR(v, e, u), ref::RR(v, e, u)); }
        }
    }
    // ---- the manifold solver's Newton step (Manifolds.slangh manifoldOneStep, sphcoords, sphcoordsDerivatives): spherical-coordinate
    // constraint, its Jacobian from the direction differentials, 2x2 inverse, position step
    for (int i = 0; i < N; ++i) {
        Cmp c("Manifolds.manifoldOneStep");
        orc::float3pdv w, p; orc::float3kpdv ws;
        w.v = r.dir(); w.ddu = r.f3(-2.f, 2.f); w.ddv = r.f3(-2.f, 2.f);
        ws.v = r.u() < .5f ? orc::normalize(w.v + r.f3(-.05f, .05f)) : r.dir(); ws.ddu = r.f3(-2.f, 2.f); ws.ddv = r.f3(-2.f, 2.f); ws.ddl = r.f3(-2.f, 2.f);
        p.v = r.f3(-3.f, 3.f); p.ddu = r.f3(-1.f, 1.f); p.ddv = r.f3(-1.f, 1.f);
        ref::float3kpdv rws; rws.v = R(ws.v); rws.ddu = R(ws.ddu); rws.ddv = R(ws.ddv); rws.ddl = R(ws.ddl);
        float l0 = 0.f, l1 = 0.f; orc::float3 s0(0.f); ref::float3 s1(0.f);
        const bool v0 = orc::Integrator::manifoldOneStep(w, ws, p, l0, s0), v1 = ref::manifoldOneStep(R(w), rws, R(p), l1, s1);
        c.add(v0, v1); c.add(l0, l1); if (v0 && v1) c.add(s0, s1);
        c.add(orc::Integrator::sphcoords(w.v), ref::sphcoords(R(w.v)));
    }
    // ---- the closest-hit program of the sample stage, whole: pltpt_sample.rt.slang scatterTriangleMeshClosestHit (with shouldTerminatePath,
    // pltSampleBSDF), pltpt.slangh appendBounce, Path.slang RayFlags / BounceData / Path -- hit -> vertex data -> shading data -> material
    // instance -> alpha pass-through / Russian roulette / BSDF sample -> guide outputs -> path book-keeping and the bounce record, on a
    // four-triangle scene with random materials and random incoming path states.  Slang's dynamic material dispatch is a switch over the ten
    // transpiled types (ref_shims_late.h); the DXR system values, the packed hit word and the index / matrix fetch are supplied by the caller
    {
        constexpr uint32_t kTri = 4, kMaxBounces = 8, kW = 4, kH = 4;
        static orc::Scene osc;
        osc.profiles = table;
        static PltptPackedVertex verts[3 * kTri];
        static ref::PackedStaticVertexData rverts[3 * kTri];
        static PltptGeometryInstance insts[kTri];
        static PltptMeshDesc meshes[kTri];
        static float world[16 * kTri], worldIT[16 * kTri];
        static ref::float4x4 rworld[kTri];
        static ref::float3x3 rworldIT[kTri];
        static bool cw[kTri];
        static PltptMaterialBlob blobs[kTri];
        osc.d.vertices = verts; osc.d.vertexCount = 3 * kTri; osc.d.geometryInstances = insts; osc.d.instanceCount = kTri; osc.d.meshes = meshes; osc.d.meshCount = kTri;
        osc.d.worldMatrices = world; osc.d.inverseTransposeWorldMatrices = worldIT; osc.d.matrixCount = kTri; osc.d.materialData = blobs; osc.d.materialCount = kTri;
        osc.d.hitInfoTypeBits = 3; osc.d.hitInfoInstanceIDBits = 2; osc.d.hitInfoPrimitiveIndexBits = 1;
        orc::Integrator integ;
        integ.scene = &osc;
        integ.opt = PltptOptions{};
        integ.opt.maxBounces = kMaxBounces; integ.opt.HWSS = 4;
        integ.width = kW; integ.height = kH;
        static PltptBounce obounces[kW * kH * (kMaxBounces + 1)];
        static float oalbedo[kW * kH * 4], onormal[kW * kH * 4];
        integ.bounces = obounces; integ.outAlbedo = oalbedo; integ.outNormal = onormal;
        // the reference side
        static ref::BounceData rbounces[kW * kH * (kMaxBounces + 1)];
        static ref::float4 ralbedo[kW * kH], rnormal[kW * kH];
        ref::mMaxBounces = kMaxBounces; ref::mTileSize = kW;
        ref::bounceBuffer.p = rbounces;
        ref::gOutputAlbedo.p = ralbedo; ref::gOutputAlbedo.width = kW; ref::gOutputNormal.p = rnormal; ref::gOutputNormal.width = kW;
        ref::gScene.vertices.p = rverts;
        ref::gScene.materials.blobs = reinterpret_cast<const ref::MaterialDataBlob*>(blobs);
        ref::gScene.materials.textureFn = nullptr;
        ref::gScene.indexFn = [](ref::GeometryInstanceID id, ref::uint tri) { return ref::uint3(3u * id.index + 3u * tri, 3u * id.index + 3u * tri + 1u, 3u * id.index + 3u * tri + 2u); };
        ref::gScene.materialIDFn = [](ref::GeometryInstanceID id) { return id.index; };
        static const orc::Scene* packScene = &osc;
        ref::TriangleHit::packFn = [](const ref::TriangleHit& h) {
            orc::TriangleHit oh; oh.instanceID = h.instanceID.index; oh.primitiveIndex = h.primitiveIndex; oh.barycentrics = orc::float2(h.barycentrics.x, h.barycentrics.y);
            uint32_t w[4]; packScene->packHit(oh, w); return ref::PackedHitInfo(w[0], w[1], w[2], w[3]); };
        static ref::float4x4* gW = rworld; static ref::float3x3* gWIT = rworldIT; static bool* gCW = cw;
        ref::gScene.vertexDataFn = [](ref::TriangleHit hit) {   // Scene.slang:576-585
            const ref::uint3 vtxIndices = ref::gScene.getIndices(hit.instanceID, hit.primitiveIndex);
            ref::StaticVertexData vertices[3] = {ref::gScene.getVertex(vtxIndices[0]), ref::gScene.getVertex(vtxIndices[1]), ref::gScene.getVertex(vtxIndices[2])};
            return ref::gScene.computeVertexData(vertices, gW[hit.instanceID.index], gWIT[hit.instanceID.index], gCW[hit.instanceID.index], hit.getBarycentricWeights()); };
        const uint32_t kTypes[10] = {PLTPT_MATERIAL_DIFFUSE, PLTPT_MATERIAL_CONDUCTOR, PLTPT_MATERIAL_DIELECTRIC, PLTPT_MATERIAL_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_THIN_DIELECTRIC,
                                     PLTPT_MATERIAL_COATED_CONDUCTOR, PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_GRATED_CONDUCTOR, PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC,
                                     PLTPT_MATERIAL_MULTILAYERED_STACK};
        auto half = [&](float lo, float hi) { return uint16_t(ref::f32tof16(r.range(lo, hi))); };
        auto halfBits = [&](float lo, float hi) { return ref::f32tof16(r.range(lo, hi)); };
        for (int i = 0; i < N / 4; ++i) {
            // geometry and materials of the four instances
            for (uint32_t t = 0; t < kTri; ++t) {
                const orc::float3 c0 = r.f3(-3.f, 3.f);
                for (int k = 0; k < 3; ++k) {
                    const orc::float3 pos = c0 + r.f3(-1.f, 1.f); const orc::float2 tc = r.f2(-1.f, 2.f);
                    const float sgn = r.u() < .2f ? 0.f : (r.u() < .5f ? -1.f : 1.f);
                    const uint32_t w0 = halfBits(-1.f, 1.f) | (halfBits(-1.f, 1.f) << 16), w1 = halfBits(-1.f, 1.f) | (ref::f32tof16(sgn) << 16), w2 = r.g();
                    PltptPackedVertex& v = verts[3 * t + k];
                    for (int a = 0; a < 3; ++a) v.position[a] = pos[a];
                    v.texCrd[0] = tc.x; v.texCrd[1] = tc.y;
                    v.packedNormalTangentCurveRadius[0] = w0; v.packedNormalTangentCurveRadius[1] = w1; v.packedNormalTangentCurveRadius[2] = w2;
                    rverts[3 * t + k].position = R(pos); rverts[3 * t + k].texCrd = R(tc);
                    rverts[3 * t + k].packedNormalTangentCurveRadius = ref::float3(ref::asfloat(ref::uint(w0)), ref::asfloat(ref::uint(w1)), ref::asfloat(ref::uint(w2)));
                }
                insts[t] = PltptGeometryInstance{}; meshes[t] = PltptMeshDesc{};
                cw[t] = r.u() < .5f;
                insts[t].flags = cw[t] ? PLTPT_INSTANCE_FLAG_OBJECT_FRONT_FACE_CW : 0u;
                insts[t].globalMatrixID = t; insts[t].materialID = t; insts[t].geometryID = t; insts[t].vbOffset = 3 * t;
                const orc::float3 ez = r.dir(), ex = orc::perp_stark(ez), ey = orc::cross(ez, ex), sc = r.f3(0.5f, 2.f), tr = r.f3(-2.f, 2.f);
                const orc::float3 cols[3] = {ex * sc.x, ey * sc.y, ez * sc.z}, colsIT[3] = {ex / sc.x, ey / sc.y, ez / sc.z};
                float* wm = world + 16 * t; float* wit = worldIT + 16 * t;
                for (int row = 0; row < 3; ++row) {
                    for (int c = 0; c < 3; ++c) { wm[4 * row + c] = cols[c][row]; wit[4 * row + c] = colsIT[c][row]; rworld[t][row][c] = wm[4 * row + c]; rworldIT[t][row][c] = wit[4 * row + c]; }
                    wm[4 * row + 3] = tr[row]; wit[4 * row + 3] = 0.f; rworld[t][row][3] = tr[row];
                }
                for (int c = 0; c < 4; ++c) { wm[12 + c] = wit[12 + c] = c == 3 ? 1.f : 0.f; rworld[t][3][c] = wm[12 + c]; }
                const int kind = int(r.u() * 9.999f);
                PltptMaterialBlob& blob = blobs[t];
                blob = PltptMaterialBlob{};
                blob.header[0] = kTypes[kind] | (1u << 31) | (r.u() < .5f ? (1u << 28) : 0u);
                for (int k = 0; k < 4; ++k) { blob.baseColor[k] = half(0.05f, 1.f); blob.specular[k] = half(0.02f, 2.5f); blob.data1[k] = half(0.05f, 3.f); blob.data2[k] = half(0.02f, 2.5f); }
                if (r.u() < .1f) blob.baseColor[3] = 0;   // fully transparent: the alpha pass-through
                if (r.u() < .15f) blob.specular[1] = 0;
                for (int k = 0; k < 3; ++k) blob.transmission[k] = half(0.2f, 1.f);
                blob.IoR = half(1.f, 2.2f);
                const uint32_t n = uint32_t(r.u() * 1.999f) * 2u;
                blob.iorNSpectralId = n; blob.iorKSpectralId = n + 1u; blob.emissionSpectralId = r.u() < .2f ? 4u : PLTPT_NO_SPECTRAL_PROFILE;
                if (kind == 7 || kind == 8) blob.data1[3] = uint16_t(((uint32_t(r.u() * 2.99f) | (r.u() < .2f ? 0x10u : 0u)) << 8) | (1u + uint32_t(r.u() * 7.99f)));
                if (kind == 9) blob.data1[3] = uint16_t(1u + uint32_t(r.u() * 11.99f));
            }
            PltptCameraData cam{};
            float* words = reinterpret_cast<float*>(&cam);
            for (size_t k = 0; k < sizeof(cam) / 4; ++k) words[k] = r.range(-2.f, 2.f);
            integ.cam = cam;
            std::memcpy(static_cast<void*>(&ref::gScene.camera.data), &cam, sizeof cam);
            for (int q = 0; q < 4; ++q) {
                const bool alpha = r.u() < .8f, rr = r.u() < .7f;
                integ.opt.alphaMasking = alpha; integ.opt.doRussianRoulette = rr;
                ref::kAlphaMasking = alpha; ref::kDoRussianRoulette = rr;
                const uint32_t px = uint32_t(r.u() * (kW - 0.001f)), py = uint32_t(r.u() * (kH - 0.001f));
                // an incoming path state
                orc::Path po;
                po.sg = orc::SampleGenerator(uint32_t(i), uint32_t(q), 61u);
                po.origin = r.f3(-5.f, 5.f); po.dir = r.dir(); po.normal = r.dir();
                po.wavelength = r.range(380.f, 720.f); po.weight = r.u() < .2f ? 1.f : r.range(0.01f, 1.f); po.pdf = r.range(0.01f, 5.f); po.thp_modifier = r.range(1.f, 3.f);
                if (r.u() < .05f) po.pdf = 1e-7f;   // the 1e-6 contribution cut
                po.id = (py << 16) | px;
                po.length = r.u() < .15f ? int(kMaxBounces) : int(r.u() * (kMaxBounces - 0.001f));
                po.flags.flags = (r.g() & 0x7u) | (uint32_t(r.u() * 3.999f) << 16);
                po.lastNonDeltaPdf = r.range(0.f, 4.f);
                ref::Path pr;
                pr.sg.g = po.sg; pr.origin = R(po.origin); pr.dir = R(po.dir); pr.normal = R(po.normal); pr.wavelength = po.wavelength; pr.weight = po.weight; pr.pdf = po.pdf;
                pr.thp_modifier = po.thp_modifier; pr.id = po.id; pr.length = po.length; pr.flags.flags = po.flags.flags; pr.lastNonDeltaPdf = po.lastNonDeltaPdf;
                // sentinels in every output both sides may or may not write
                const int len0 = po.length;
                PltptBounce& ob = obounces[(size_t(py) * kW + px) * (kMaxBounces + 1) + len0];
                ref::BounceData& rbnc = rbounces[((px % kW) + (py % kW) * kW) * (kMaxBounces + 1) + len0];
                std::memset(&ob, 0xcd, sizeof ob); std::memset(static_cast<void*>(&rbnc), 0xcd, sizeof rbnc);
                for (int k = 0; k < 4; ++k) { oalbedo[(py * kW + px) * 4 + k] = -7.f; onormal[(py * kW + px) * 4 + k] = -7.f; }
                ralbedo[py * kW + px] = ref::float4(-7.f); rnormal[py * kW + px] = ref::float4(-7.f);
                // a hit
                orc::TriangleHit hit;
                hit.instanceID = uint32_t(r.u() * (kTri - 0.001f)); hit.primitiveIndex = 0;
                const orc::float3 bary = orc::sample_triangle(r.f2(0.f, 1.f));
                hit.barycentrics = orc::float2(bary.y, bary.z);
                integ.scatterClosestHit(po, hit);
                ref::gCurrentInstanceID.index = hit.instanceID; ref::gCurrentPrimitiveIndex = 0;
                ref::BuiltInTriangleIntersectionAttributes attribs; attribs.barycentrics = R(hit.barycentrics);
                ref::scatterTriangleMeshClosestHit(pr, attribs);
                if (std::getenv("REF_DEBUG")) {
                    Cmp("dbg.origin").add(po.origin, pr.origin); Cmp("dbg.dir").add(po.dir, pr.dir); Cmp("dbg.normal").add(po.normal, pr.normal); Cmp("dbg.weight").add(po.weight, pr.weight);
                    Cmp("dbg.pdf").add(po.pdf, pr.pdf); Cmp("dbg.thp").add(po.thp_modifier, pr.thp_modifier); Cmp("dbg.lnd").add(po.lastNonDeltaPdf, pr.lastNonDeltaPdf);
                    Cmp("dbg.wi").add(ob.encodedWi[0], rbnc.encodedWi.x); Cmp("dbg.contrib").add(ob.pathContribution, rbnc.pathContribution); Cmp("dbg.blnd").add(ob.lastNonDeltaPdf, rbnc.lastNonDeltaPdf);
                    for (int k = 0; k < 4; ++k) { Cmp("dbg.albedo").add(oalbedo[(py * kW + px) * 4 + k], ralbedo[py * kW + px][k]); Cmp("dbg.gnormal").add(onormal[(py * kW + px) * 4 + k], rnormal[py * kW + px][k]); }
                }
                Cmp c("pltpt_sample.scatterTriangleMeshClosestHit");
                c.add(po.origin, pr.origin); c.add(po.dir, pr.dir); c.add(po.normal, pr.normal); c.add(po.weight, pr.weight); c.add(po.pdf, pr.pdf); c.add(po.thp_modifier, pr.thp_modifier);
                c.add(po.length, pr.length); c.add(po.flags.flags, uint32_t(pr.flags.flags)); c.add(po.lastNonDeltaPdf, pr.lastNonDeltaPdf); c.add(po.sg.next(), pr.sg.next());
                for (int k = 0; k < 4; ++k) c.add(ob.hit[k], uint32_t(rbnc.hit[k]));
                c.add(ob.encodedWi[0], rbnc.encodedWi.x); c.add(ob.encodedWi[1], rbnc.encodedWi.y); c.add(ob.sampledBsdfLobe, uint32_t(rbnc.sampledBsdfLobe)); c.add(ob.flags, uint32_t(rbnc.flags.flags));
                c.add(ob.pathContribution, rbnc.pathContribution); c.add(ob.lastNonDeltaPdf, rbnc.lastNonDeltaPdf);
                for (int k = 0; k < 4; ++k) { c.add(oalbedo[(py * kW + px) * 4 + k], ralbedo[py * kW + px][k]); c.add(onormal[(py * kW + px) * 4 + k], rnormal[py * kW + px][k]); }
                // the miss program on the state the hit left behind (scatterMiss: flag + bounce record with an empty hit word), when the path goes on
                if (po.flags.valid() && po.length <= int(kMaxBounces)) {
                    PltptBounce& ob2 = obounces[(size_t(py) * kW + px) * (kMaxBounces + 1) + po.length];
                    ref::BounceData& rb2 = rbounces[((px % kW) + (py % kW) * kW) * (kMaxBounces + 1) + po.length];
                    std::memset(&ob2, 0xcd, sizeof ob2); std::memset(static_cast<void*>(&rb2), 0xcd, sizeof rb2);
                    const uint32_t zero[4] = {0, 0, 0, 0};
                    po.flags.set(orc::RayFlags::missF, true);                 // plt_integrator.h samplePixel, the miss branch
                    integ.appendBounce(po, zero, -po.dir, 0);
                    ref::scatterMiss(pr);
                    Cmp m("pltpt_sample.scatterMiss");
                    m.add(po.flags.flags, uint32_t(pr.flags.flags));
                    for (int k = 0; k < 4; ++k) m.add(ob2.hit[k], uint32_t(rb2.hit[k]));
                    m.add(ob2.encodedWi[0], rb2.encodedWi.x); m.add(ob2.encodedWi[1], rb2.encodedWi.y); m.add(ob2.sampledBsdfLobe, uint32_t(rb2.sampledBsdfLobe)); m.add(ob2.flags, uint32_t(rb2.flags.flags));
                    m.add(ob2.pathContribution, rb2.pathContribution); m.add(ob2.lastNonDeltaPdf, rb2.lastNonDeltaPdf);
                }
            }
        }
        osc.d = PltptSceneDesc{};
    }
    ref_compare_frames(N, r, prof);
}

// ---- WHOLE FRAMES: both ray-generation programs of the pass as the reference wrote them (pltpt_sample.rt.slang main -> TraceRay -> closest-hit /
// miss programs; pltpt_solve.rt.slang main -> NEE with generateLightSample, visibilityRay / manifoldNEESample (Manifolds.slangh: traceMNEEOccluders,
// the Newton walk with its line search, manifoldForwardUnwind), evalNEE, evalEmissive, onMiss, forwardUnwind, measure), next to the oracle's
// samplePixel / solvePixel, on random closed-ish scenes of large triangles with random materials of all ten types, emissive triangles in a light
// collection, analytic lights of all six kinds and an environment map, under random option sets.  The only thing supplied to the reference text is
// the intersection search behind TraceRay (ref_shims_late.h): DXR hardware in the reference, the oracle's ray caster here, the same for both sides.
static void ref_compare_frames(int N, Rnd& r, PltptSpectralProfile* prof) {
    constexpr uint32_t kTri = 20, kW = 4, kH = 4, kMaxB = 8;
    static orc::Scene osc;
    static PltptPackedVertex verts[3 * kTri];
    static ref::PackedStaticVertexData rverts[3 * kTri];
    static PltptGeometryInstance insts[kTri];
    static PltptMeshDesc meshes[kTri];
    static float world[16 * kTri], worldIT[16 * kTri];
    static ref::float4x4 rworld[kTri];
    static ref::float3x3 rworldIT[kTri];
    static bool cw[kTri];
    static PltptMaterialBlob blobs[kTri];
    static PltptEmissiveTriangle etris[kTri];
    static PltptEmissiveFlux eflux[kTri];
    static uint32_t eactive[kTri], ealias[2 * kTri], perInst[kTri];
    static ref::uint2 ralias[kTri];
    static ref::PackedEmissiveTriangle retris[kTri];
    static ref::EmissiveFlux rflux[kTri];
    static PltptLightData lights[2];
    static ref::LightData rlights[2];
    static PltptBounce obounces[kW * kH * (kMaxB + 1)];
    static ref::BounceData rbounces[kW * kH * (kMaxB + 1)];
    static float ocolor[kW * kH * 4], oalbedo[kW * kH * 4], onormal[kW * kH * 4];
    static ref::float4 rcolor[kW * kH], ralbedo[kW * kH], rnormal[kW * kH];
    // one environment map for all scenes (its importance map is expensive to build): swapped into the scene when the scene uses it
    constexpr uint32_t EW = 16, EH = 8;
    static float envTexels[EW * EH * 4];
    static orc::EnvMap sharedEnv;
    PltptSceneDesc envDesc{};
    {
        for (uint32_t t = 0; t < EW * EH; ++t) {
            const float sun = (t == 2 * EW + 5) ? 60.f : 1.f;
            envTexels[4 * t] = sun * r.range(0.05f, 1.f); envTexels[4 * t + 1] = sun * r.range(0.05f, 1.f); envTexels[4 * t + 2] = sun * r.range(0.05f, 1.f); envTexels[4 * t + 3] = 1.f;
        }
        envDesc.envMapTexels = envTexels; envDesc.envMapWidth = EW; envDesc.envMapHeight = EH;
        const float ang = -0.4f, ca = std::cos(ang), sa = std::sin(ang);
        const float rot[12] = {ca, 0.f, sa, 0.f, 0.f, 1.f, 0.f, 0.f, -sa, 0.f, ca, 0.f}, inv[12] = {ca, 0.f, -sa, 0.f, 0.f, 1.f, 0.f, 0.f, sa, 0.f, ca, 0.f};
        for (int k = 0; k < 12; ++k) { envDesc.envMap.transform[k] = rot[k]; envDesc.envMap.invTransform[k] = inv[k]; }
        envDesc.envMap.tint[0] = 1.f; envDesc.envMap.tint[1] = 0.9f; envDesc.envMap.tint[2] = 0.8f; envDesc.envMap.intensity = 0.6f;
        sharedEnv.init(envDesc);
        ref::EnvMap& renv = ref::gScene.envMap;
        for (int row = 0; row < 3; ++row) for (int c = 0; c < 4; ++c) { renv.data.transform[row][c] = rot[4 * row + c]; renv.data.invTransform[row][c] = inv[4 * row + c]; }
        renv.data.tint = ref::float3(1.f, 0.9f, 0.8f); renv.data.intensity = 0.6f;
        renv.envMap.ctx = &osc.env; renv.envMap.width = EW; renv.envMap.height = EH;
        renv.envMap.loadFn = [](const void* c, int x, int y, int) { const orc::float3 t = static_cast<const orc::EnvMap*>(c)->texel(x, y); return ref::float4(t.x, t.y, t.z, 1.f); };
        renv.envMap.sampleFn = [](const void* c, ref::float2 uv, float) { const orc::float3 t = static_cast<const orc::EnvMap*>(c)->sampleBilinear(orc::float2(uv.x, uv.y)); return ref::float4(t.x, t.y, t.z, 1.f); };
        ref::EnvMapSampler& smp = ref::envMapSampler;
        smp.importanceMap.ctx = &osc.env; smp.importanceMap.width = smp.importanceMap.height = orc::EnvMap::kDimension;
        smp.importanceMap.loadFn = [](const void* c, int x, int y, int mip) { return static_cast<const orc::EnvMap*>(c)->load(uint32_t(x), uint32_t(y), uint32_t(mip)); };
        smp.importanceMap.sampleFn = [](const void* c, ref::float2 uv, float) {
            const int dim = int(orc::EnvMap::kDimension);
            const int x = std::min(std::max(int(std::floor(uv.x * float(dim))), 0), dim - 1), y = std::min(std::max(int(std::floor(uv.y * float(dim))), 0), dim - 1);
            return static_cast<const orc::EnvMap*>(c)->load(uint32_t(x), uint32_t(y), 0u); };
        smp.importanceInvDim = ref::float2(1.f / float(orc::EnvMap::kDimension));
        smp.importanceBaseMip = sharedEnv.baseMip;
    }
    // the reference side's scene services
    ref::mTileSize = kW;
    ref::bounceBuffer.p = rbounces;
    ref::gOutputColor.p = rcolor; ref::gOutputColor.width = kW; ref::gOutputAlbedo.p = ralbedo; ref::gOutputAlbedo.width = kW; ref::gOutputNormal.p = rnormal; ref::gOutputNormal.width = kW;
    ref::gScene.vertices.p = rverts;
    ref::gScene.materials.blobs = reinterpret_cast<const ref::MaterialDataBlob*>(blobs);
    ref::gScene.materials.textureFn = nullptr;
    ref::gScene.indexFn = [](ref::GeometryInstanceID id, ref::uint tri) { return ref::uint3(3u * id.index + 3u * tri, 3u * id.index + 3u * tri + 1u, 3u * id.index + 3u * tri + 2u); };
    ref::gScene.materialIDFn = [](ref::GeometryInstanceID id) { return id.index; };
    static const orc::Scene* sc = &osc;
    ref::TriangleHit::packFn = [](const ref::TriangleHit& h) {
        orc::TriangleHit oh; oh.instanceID = h.instanceID.index; oh.primitiveIndex = h.primitiveIndex; oh.barycentrics = orc::float2(h.barycentrics.x, h.barycentrics.y);
        uint32_t w[4]; sc->packHit(oh, w); return ref::PackedHitInfo(w[0], w[1], w[2], w[3]); };
    ref::TriangleHit::unpackFn = [](ref::PackedHitInfo p, ref::TriangleHit& h) {
        const uint32_t w[4] = {p.x, p.y, p.z, p.w}; const orc::TriangleHit oh = sc->unpackHit(w);
        h.instanceID.index = oh.instanceID; h.primitiveIndex = oh.primitiveIndex; h.barycentrics = ref::float2(oh.barycentrics.x, oh.barycentrics.y); };
    static ref::float4x4* gW = rworld; static ref::float3x3* gWIT = rworldIT; static bool* gCW = cw;
    ref::gScene.vertexDataFn = [](ref::TriangleHit hit) {   // Scene.slang:576-585
        const ref::uint3 vtxIndices = ref::gScene.getIndices(hit.instanceID, hit.primitiveIndex);
        ref::StaticVertexData vertices[3] = {ref::gScene.getVertex(vtxIndices[0]), ref::gScene.getVertex(vtxIndices[1]), ref::gScene.getVertex(vtxIndices[2])};
        return ref::gScene.computeVertexData(vertices, gW[hit.instanceID.index], gWIT[hit.instanceID.index], gCW[hit.instanceID.index], hit.getBarycentricWeights()); };
    ref::gCastClosest = [](ref::float3 o, ref::float3 d, float tMin, float tMax) {
        const orc::RayHit h = sc->traceClosest(orc::float3(o.x, o.y, o.z), orc::float3(d.x, d.y, d.z), tMin, tMax);
        ref::RayCastResult res; res.hit = h.hit.valid();
        if (res.hit) { res.instanceIndex = h.hit.instanceID; res.primitiveIndex = h.hit.primitiveIndex; res.barycentrics = ref::float2(h.hit.barycentrics.x, h.hit.barycentrics.y); res.t = h.t; }
        return res; };
    ref::gCastAny = [](ref::float3 o, ref::float3 d, float tMin, float tMax) { return sc->traceAny(orc::float3(o.x, o.y, o.z), orc::float3(d.x, d.y, d.z), tMin, tMax); };
    const uint32_t kTypes[10] = {PLTPT_MATERIAL_DIFFUSE, PLTPT_MATERIAL_CONDUCTOR, PLTPT_MATERIAL_DIELECTRIC, PLTPT_MATERIAL_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_THIN_DIELECTRIC,
                                 PLTPT_MATERIAL_COATED_CONDUCTOR, PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC, PLTPT_MATERIAL_GRATED_CONDUCTOR, PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC,
                                 PLTPT_MATERIAL_MULTILAYERED_STACK};
    auto half = [&](float lo, float hi) { return uint16_t(ref::f32tof16(r.range(lo, hi))); };
    auto halfBits = [&](float lo, float hi) { return ref::f32tof16(r.range(lo, hi)); };
    const int scenes = std::max(4, N / 64);
    uint64_t pixels = 0, identicalPixels = 0, vertices = 0, mneePixels = 0, evals = 0;
    for (int i = 0; i < scenes; ++i) {
        // ---- geometry: large triangles around the origin, one instance (own matrix, own material) each
        const bool smoothGlass = r.u() < .35f;   // a scene rich in smooth dielectrics: manifold walks
        for (uint32_t t = 0; t < kTri; ++t) {
            const orc::float3 c0 = r.dir() * r.range(1.5f, 4.f);
            for (int k = 0; k < 3; ++k) {
                const orc::float3 pos = c0 + r.f3(-5.f, 5.f); const orc::float2 tc = r.f2(-1.f, 2.f);
                const float sgn = r.u() < .2f ? 0.f : (r.u() < .5f ? -1.f : 1.f);
                const uint32_t w0 = halfBits(-1.f, 1.f) | (halfBits(-1.f, 1.f) << 16), w1 = halfBits(-1.f, 1.f) | (ref::f32tof16(sgn) << 16), w2 = r.g();
                PltptPackedVertex& v = verts[3 * t + k];
                for (int a = 0; a < 3; ++a) v.position[a] = pos[a];
                v.texCrd[0] = tc.x; v.texCrd[1] = tc.y;
                v.packedNormalTangentCurveRadius[0] = w0; v.packedNormalTangentCurveRadius[1] = w1; v.packedNormalTangentCurveRadius[2] = w2;
                rverts[3 * t + k].position = R(pos); rverts[3 * t + k].texCrd = R(tc);
                rverts[3 * t + k].packedNormalTangentCurveRadius = ref::float3(ref::asfloat(ref::uint(w0)), ref::asfloat(ref::uint(w1)), ref::asfloat(ref::uint(w2)));
            }
            insts[t] = PltptGeometryInstance{}; meshes[t] = PltptMeshDesc{};
            meshes[t].vertexCount = 3; meshes[t].indexCount = 0;
            cw[t] = r.u() < .5f;
            insts[t].flags = cw[t] ? PLTPT_INSTANCE_FLAG_OBJECT_FRONT_FACE_CW : 0u;
            insts[t].globalMatrixID = t; insts[t].materialID = t; insts[t].geometryID = t; insts[t].vbOffset = 3 * t;
            const orc::float3 ez = r.dir(), ex = orc::perp_stark(ez), ey = orc::cross(ez, ex), scl = r.f3(0.7f, 1.4f), tr = r.f3(-.5f, .5f);
            const orc::float3 cols[3] = {ex * scl.x, ey * scl.y, ez * scl.z}, colsIT[3] = {ex / scl.x, ey / scl.y, ez / scl.z};
            float* wm = world + 16 * t; float* wit = worldIT + 16 * t;
            for (int row = 0; row < 3; ++row) {
                for (int c = 0; c < 3; ++c) { wm[4 * row + c] = cols[c][row]; wit[4 * row + c] = colsIT[c][row]; rworld[t][row][c] = wm[4 * row + c]; rworldIT[t][row][c] = wit[4 * row + c]; }
                wm[4 * row + 3] = tr[row]; wit[4 * row + 3] = 0.f; rworld[t][row][3] = tr[row];
            }
            for (int c = 0; c < 4; ++c) { wm[12 + c] = wit[12 + c] = c == 3 ? 1.f : 0.f; rworld[t][3][c] = wm[12 + c]; }
            int kind = int(r.u() * 9.999f);
            if (smoothGlass && r.u() < .6f) kind = r.u() < .5f ? 2 : 4;
            const bool emitter = r.u() < .25f;   // only the diffuse material emits (PLTDiffuseMaterial.slang:44-47); a few non-diffuse records keep an emission id on purpose
            if (emitter && r.u() < .9f) kind = 0;
            PltptMaterialBlob& blob = blobs[t];
            blob = PltptMaterialBlob{};
            blob.header[0] = kTypes[kind] | (1u << 31) | (r.u() < .6f ? (1u << 28) : 0u);
            for (int k = 0; k < 4; ++k) { blob.baseColor[k] = half(0.05f, 1.f); blob.specular[k] = half(0.02f, 2.5f); blob.data1[k] = half(0.05f, 3.f); blob.data2[k] = half(0.02f, 2.5f); }
            if (r.u() < .08f) blob.baseColor[3] = 0;   // fully transparent: the alpha pass-through
            if (r.u() < .15f || (smoothGlass && (kind == 2 || kind == 4))) blob.specular[1] = 0;
            for (int k = 0; k < 3; ++k) blob.transmission[k] = half(0.2f, 1.f);
            blob.IoR = half(1.05f, 2.2f);
            const uint32_t n = uint32_t(r.u() * 1.999f) * 2u;
            blob.iorNSpectralId = n; blob.iorKSpectralId = n + 1u; blob.emissionSpectralId = emitter ? 4u : PLTPT_NO_SPECTRAL_PROFILE;
            if (kind == 7 || kind == 8) blob.data1[3] = uint16_t(((uint32_t(r.u() * 2.99f) | (r.u() < .2f ? 0x10u : 0u)) << 8) | (1u + uint32_t(r.u() * 7.99f)));
            if (kind == 9) blob.data1[3] = uint16_t(1u + uint32_t(r.u() * 11.99f));
        }
        PltptSceneDesc d{};
        d.vertices = verts; d.vertexCount = 3 * kTri; d.geometryInstances = insts; d.instanceCount = kTri; d.meshes = meshes; d.meshCount = kTri;
        d.worldMatrices = world; d.inverseTransposeWorldMatrices = worldIT; d.matrixCount = kTri; d.materialData = blobs; d.materialCount = kTri;
        d.spectralProfiles = prof; d.profileCount = 5;
        d.hitInfoTypeBits = 3; d.hitInfoInstanceIDBits = 5; d.hitInfoPrimitiveIndexBits = 1;
        // ---- the light collection over the emissive instances (LightCollection.slang: packed world-space triangles, flux, active list,
        // per-instance offsets; EmissivePowerSampler's alias table)
        uint32_t ne = 0; float total = 0.f;
        for (uint32_t t = 0; t < kTri; ++t) {
            perInst[t] = 0xffffffffu;
            if (blobs[t].emissionSpectralId == PLTPT_NO_SPECTRAL_PROFILE) continue;
            perInst[t] = ne;
            const float* wm = world + 16 * t;
            orc::float3 vs[3];
            for (int k = 0; k < 3; ++k) {
                const float* P = verts[3 * t + k].position;
                for (int row = 0; row < 3; ++row) vs[k][row] = ((wm[4 * row + 0] * P[0] + wm[4 * row + 1] * P[1]) + wm[4 * row + 2] * P[2]) + wm[4 * row + 3];
                const uint32_t uv = ref::f32tof16(verts[3 * t + k].texCrd[0]) | (ref::f32tof16(verts[3 * t + k].texCrd[1]) << 16);
                float uvf; std::memcpy(&uvf, &uv, 4);
                for (int a = 0; a < 3; ++a) etris[ne].posAndTexCoords[k][a] = vs[k][a];
                etris[ne].posAndTexCoords[k][3] = uvf;
                retris[ne].posAndTexCoords[k] = ref::float4(vs[k].x, vs[k].y, vs[k].z, uvf);
            }
            const orc::float3 fn = orc::cross(vs[1] - vs[0], vs[2] - vs[0]);
            const float area = 0.5f * orc::length(fn);
            etris[ne].normal = orc::encodeNormal2x16(orc::normalize(cw[t] ? -fn : fn)); std::memcpy(&etris[ne].area, &area, 4); etris[ne].materialID = t; etris[ne].lightIdx = 0;
            retris[ne].normal = etris[ne].normal; retris[ne].area = etris[ne].area; retris[ne].materialID = t; retris[ne].lightIdx = 0;
            eflux[ne].flux = area * r.range(0.5f, 2.f); total += eflux[ne].flux; rflux[ne].flux = eflux[ne].flux;
            ++ne;
        }
        for (uint32_t t = 0; t < ne; ++t) {
            eactive[t] = t;
            const uint32_t above = uint32_t(r.u() * (ne - 0.001f)), below = uint32_t(r.u() * (ne - 0.001f));
            const uint32_t px = (ref::f32tof16(r.u()) << 16) | ((above >> 8) & 0xffffu), py = ((above & 0xffu) << 24) | (below & 0xffffffu);
            ealias[2 * t] = px; ealias[2 * t + 1] = py; ralias[t] = ref::uint2(px, py);
        }
        const float invSum = ne ? 1.f / total : 0.f;
        d.emissiveTriangles = ne ? etris : nullptr; d.emissiveTriangleCount = ne; d.emissiveFlux = eflux; d.activeTriangles = eactive; d.activeTriangleCount = ne;
        d.perMeshInstanceOffset = perInst; d.triangleAliasTable = ealias; d.invWeightsSum = invSum;
        ref::LightCollection& lc = ref::gScene.lightCollection;
        lc.triangleCount = ne; lc.activeTriangleCount = ne; lc.triangleData.p = retris; lc.activeTriangles.p = eactive; lc.fluxData.p = rflux; lc.perMeshInstanceOffset.p = perInst;
        ref::gPowerSampler._emissivePower.invWeightsSum = invSum; ref::gPowerSampler._emissivePower.triangleAliasTable.p = ralias;
        // ---- analytic lights: none, one or two, of any of the six kinds
        const uint32_t nl = std::min(2u, uint32_t(r.u() * 3.999f));
        for (uint32_t l = 0; l < nl; ++l) {
            PltptLightData& L = lights[l]; ref::LightData& Lr = rlights[l];
            L = PltptLightData{}; Lr = ref::LightData();
            L.type = uint32_t(r.u() * 5.999f);
            const orc::float3 pos = r.f3(-2.f, 2.f), dir = r.dir();
            for (int k = 0; k < 3; ++k) { L.posW[k] = pos[k]; L.dirW[k] = dir[k]; }
            L.openingAngle = r.u() < .5f ? orc::kPi : r.range(0.4f, 1.5f);
            L.cosOpeningAngle = std::cos(L.openingAngle);
            L.penumbraAngle = r.u() < .5f ? 0.f : r.range(0.01f, 0.3f);
            L.cosSubtendedAngle = r.u() < .3f ? 0.9999893f : std::cos(r.range(0.001f, 0.4f));
            L.intensitySpectrumId = 4u;
            L.surfaceArea = r.range(0.05f, 8.f);
            const orc::float3 ez = r.dir(), ex = orc::perp_stark(ez), ey = orc::cross(ez, ex), scl = r.f3(0.2f, 1.5f);
            const orc::float3 cols[3] = {ex * scl.x, ey * scl.y, ez * scl.z}, colsIT[3] = {ex / scl.x, ey / scl.y, ez / scl.z};
            for (int row = 0; row < 3; ++row) {
                for (int c = 0; c < 3; ++c) { L.transMat[4 * row + c] = cols[c][row]; L.transMatIT[4 * row + c] = colsIT[c][row]; }
                L.transMat[4 * row + 3] = pos[row]; L.transMatIT[4 * row + 3] = 0.f;
            }
            L.transMat[15] = 1.f; L.transMatIT[15] = 1.f;
            Lr.posW = R(pos); Lr.type = L.type; Lr.dirW = R(dir); Lr.openingAngle = L.openingAngle; Lr.cosOpeningAngle = L.cosOpeningAngle; Lr.cosSubtendedAngle = L.cosSubtendedAngle;
            Lr.penumbraAngle = L.penumbraAngle; Lr.intensitySpectrumId = L.intensitySpectrumId; Lr.surfaceArea = L.surfaceArea;
            for (int row = 0; row < 4; ++row) for (int c = 0; c < 4; ++c) { Lr.transMat[row][c] = L.transMat[4 * row + c]; Lr.transMatIT[row][c] = L.transMatIT[4 * row + c]; }
        }
        d.lights = nl ? lights : nullptr; d.lightCount = nl;
        ref::gScene.lights = rlights; ref::gScene.lightCount = nl;
        osc.init(d);
        const bool withEnv = r.u() < .5f;
        if (withEnv) std::swap(osc.env, sharedEnv);
        // ---- camera: a pinhole at a random place looking at the origin
        PltptCameraData cam{};
        {
            const orc::float3 pos = r.dir() * r.range(0.2f, 2.5f), target = r.f3(-.5f, .5f), up0 = r.dir();
            const orc::float3 w = orc::normalize(target - pos), u = orc::normalize(orc::cross(w, up0)), v = orc::cross(u, w);
            const float fd = r.range(0.5f, 3.f), halfH = r.range(0.3f, 1.2f) * fd, halfW = halfH * float(kW) / float(kH);
            for (int k = 0; k < 3; ++k) { cam.posW[k] = pos[k]; cam.cameraW[k] = w[k] * fd; cam.cameraU[k] = u[k] * halfW; cam.cameraV[k] = v[k] * halfH; }
            cam.jitterX = r.range(-.5f, .5f) / float(kW); cam.jitterY = r.range(-.5f, .5f) / float(kH);
            cam.nearZ = 0.1f; cam.farZ = 1000.f; cam.focalDistance = fd;
            if (r.u() < .3f) { cam.polarizer = r.range(0.2f, 1.f); cam.polarizerTheta = r.range(0.f, 3.f); }
        }
        std::memcpy(static_cast<void*>(&ref::gScene.camera.data), &cam, sizeof cam);
        // ---- options
        orc::Integrator integ;
        integ.scene = &osc;
        PltptOptions& o = integ.opt;
        o = PltptOptions{}; o.sampleGenerator = 1; o.debugViewIntensity = 1.f;
        o.maxBounces = 2u + uint32_t(r.u() * (kMaxB - 2 + 0.999f)); o.tileSize = kW; o.HWSS = 4; o.HWSSDoMIS = r.u() < .7f;
        o.sourcingAreaFromEmissiveGeometry = r.range(0.01f, 4.f); o.sourcingMaxBeamOmega = r.range(0.001f, 0.1f);
        o.useDirectLights = r.u() < .85f; o.useEnvLights = r.u() < .9f; o.useEmissiveLights = r.u() < .9f; o.useAnalyticLights = r.u() < .9f;
        o.alphaMasking = r.u() < .8f; o.doNEE = r.u() < .9f; o.doMIS = r.u() < .8f; o.emissiveSampler = r.u() < .5f ? PLTPT_EMISSIVE_SAMPLER_POWER : PLTPT_EMISSIVE_SAMPLER_UNIFORM;
        o.doRussianRoulette = r.u() < .6f; o.NEEUsePerTileSG = r.u() < .3f; o.doImportanceSampleEmitters = r.u() < .6f; o.doMNEE = r.u() < .75f;
        o.MNEEMaxOccluders = 2; o.MNEEMaxIterations = r.u() < .5f ? 40u : 12u; o.MNEESolverThreshold = r.u() < .5f ? 1e-5f : 5e-5f;
        integ.cam = cam; integ.width = kW; integ.height = kH;
        integ.bounces = obounces; integ.outColor = ocolor; integ.outAlbedo = oalbedo; integ.outNormal = onormal;
        ref::mMaxBounces = o.maxBounces; ref::mHWSSDoMIS = o.HWSSDoMIS; ref::kUseDirectLights = o.useDirectLights;
        ref::kUseEnvLight = integ.kUseEnvLight(); ref::kUseEmissiveLights = integ.kUseEmissiveLights(); ref::kUseAnalyticLights = integ.kUseAnalyticLights();   // PLTPT.cpp:232-238
        ref::kAlphaMasking = o.alphaMasking; ref::kImportanceSampleEmissiveSpectra = o.doImportanceSampleEmitters;
        ref::kDoNEE = integ.kDoNEE(); ref::kDoMSonNEE = integ.kDoMSonNEE(); ref::kDoMIS = integ.kDoMIS(); ref::kNEEUsePerTileSelectorSG = integ.kNEEUsePerTileSG();
        ref::kDoRussianRoulette = o.doRussianRoulette;
        ref::kMNEEMaxIterations = int(o.MNEEMaxIterations); ref::kMNEESolverThreshold = o.MNEESolverThreshold;
        ref::kSourcingAreaFromEmissiveGeometry = o.sourcingAreaFromEmissiveGeometry; ref::kSourcingMaxBeamOmega = o.sourcingMaxBeamOmega;
        ref::gEmissiveKind = o.emissiveSampler == PLTPT_EMISSIVE_SAMPLER_POWER ? 2 : 1;
        ref::kOutputSize = ref::uint2(kW, kH); ref::kTile = ref::uint2(0u, 0u);
        for (uint32_t frame = 0; frame < 2; ++frame) {
            integ.frameCount = 7u * uint32_t(i) + frame; ref::gFrameCount = integ.frameCount;
            std::memset(obounces, 0xcd, sizeof obounces); std::memset(static_cast<void*>(rbounces), 0xcd, sizeof rbounces);
            for (uint32_t py = 0; py < kH; ++py) for (uint32_t px = 0; px < kW; ++px) {
                integ.samplePixel(px, py);
                ref::gDispatchRaysIndex = ref::uint3(px, py, 0u);
                ref::sampleRayGen();
            }
            Cmp cb("pltpt_sample.main_bounce_records"), cg("pltpt_sample.main_guides"), cc("pltpt_solve.main_color");
            for (uint32_t py = 0; py < kH; ++py) for (uint32_t px = 0; px < kW; ++px) {
                const uint32_t pix = py * kW + px;
                // the records of this path, up to the one that ends it (the rest of the slot is never written by either side)
                for (uint32_t p = 0; p <= o.maxBounces; ++p) {
                    const PltptBounce& ob = obounces[size_t(pix) * (o.maxBounces + 1) + p];
                    const ref::BounceData& rb = rbounces[size_t(pix) * (o.maxBounces + 1) + p];
                    for (int k = 0; k < 4; ++k) cb.add(ob.hit[k], uint32_t(rb.hit[k]));
                    cb.add(ob.encodedWi[0], rb.encodedWi.x); cb.add(ob.encodedWi[1], rb.encodedWi.y); cb.add(ob.sampledBsdfLobe, uint32_t(rb.sampledBsdfLobe)); cb.add(ob.flags, uint32_t(rb.flags.flags));
                    cb.add(ob.pathContribution, rb.pathContribution); cb.add(ob.lastNonDeltaPdf, rb.lastNonDeltaPdf);
                    ++vertices;
                    orc::RayFlags f; f.flags = ob.flags;
                    if (f.miss() || f.terminated()) break;
                }
                for (int k = 0; k < 4; ++k) { cg.add(oalbedo[pix * 4 + k], ralbedo[pix][k]); cg.add(onormal[pix * 4 + k], rnormal[pix][k]); }
            }
            for (uint32_t py = 0; py < kH; ++py) for (uint32_t px = 0; px < kW; ++px) {
                const uint32_t pix = py * kW + px;
                integ.solvePixel(px, py);
                ref::gDispatchRaysIndex = ref::uint3(px, py, 0u);
                ref::solveRayGen();
                bool same = true;
                for (int k = 0; k < 4; ++k) { cc.add(ocolor[pix * 4 + k], rcolor[pix][k]); same = same && ocolor[pix * 4 + k] == rcolor[pix][k]; }
                ++pixels; identicalPixels += same ? 1u : 0u;
                mneePixels += (ocolor[pix * 4] > 0.f || ocolor[pix * 4 + 1] > 0.f || ocolor[pix * 4 + 2] > 0.f) ? 1u : 0u;
            }
        }
        evals += integ.statEvals;
        if (withEnv) std::swap(osc.env, sharedEnv);
    }
    { Stat& s = g_stats["pltpt_solve.main_color"]; s.identical = identicalPixels; s.units = pixels; }
    if (std::getenv("REF_DEBUG")) std::fprintf(stderr, "frames: %llu pixels, %llu lit, %.2f records per path, %llu unwind evals, %llu shadow-type queries of which %llu line-search proposals\n", (unsigned long long)pixels, (unsigned long long)mneePixels, double(vertices) / double(pixels), (unsigned long long)evals, (unsigned long long)ref::gWalks, (unsigned long long)ref::gProposals);
    ref::gEmissiveKind = 0;
    osc.d = PltptSceneDesc{};
}
