// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the CPU restatement (ctypes-loadable).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include "plt_integrator.h"
#include <thread>
#include <atomic>
#include <cstdio>

using namespace orc;

namespace {
template <typename F>
void parallel_for(uint32_t n, int threads, F&& f) {
    if (threads <= 1 || n < 2) { for (uint32_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<uint32_t> next{0};
    const uint32_t chunk = std::max<uint32_t>(1, n / (uint32_t(threads) * 16));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                uint32_t b = next.fetch_add(chunk);
                if (b >= n) break;
                for (uint32_t i = b; i < std::min(n, b + chunk); ++i) f(i);
            }
        });
    for (auto& t : pool) t.join();
}
}  // namespace

extern "C" {

// ---- leaves --------------------------------------------------------------------------------
void oracle_prng(uint32_t px, uint32_t py, uint32_t sampleNumber, uint32_t n, uint32_t* out) {
    SampleGenerator sg(px, py, sampleNumber);
    for (uint32_t i = 0; i < n; ++i) out[i] = sg.next();
}
void oracle_xoshiro(const uint32_t seed[4], uint32_t n, uint32_t* out) {
    SampleGenerator sg;
    for (int i = 0; i < 4; ++i) sg.s[i] = seed[i];
    for (uint32_t i = 0; i < n; ++i) out[i] = sg.next();
}
void oracle_splitmix64(uint64_t seed, uint32_t n, uint64_t* out) {
    for (uint32_t i = 0; i < n; ++i) out[i] = splitmix64_next(seed);
}
void oracle_sample_next_1d(uint32_t px, uint32_t py, uint32_t sampleNumber, uint32_t n, float* out) {
    SampleGenerator sg(px, py, sampleNumber);
    for (uint32_t i = 0; i < n; ++i) out[i] = sampleNext1D(sg);
}
void oracle_compute_ray_origin(uint32_t n, const float* pos, const float* normal, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        float3 r = computeRayOrigin(float3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), float3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
float oracle_sampled_spectrum_eval(float lambdaStart, float lambdaEnd, uint32_t n, const float* samples, float lambda) {
    return SampledSpectrum(lambdaStart, lambdaEnd, std::vector<float>(samples, samples + n)).eval(lambda);
}
void oracle_wavelength_to_xyz(uint32_t n, const float* lambda, float* out) {
    for (uint32_t i = 0; i < n; ++i) { float3 r = SpectrumUtils::wavelengthToXYZ(lambda[i]); out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z; }
}
void oracle_rgb_to_spectrum(const float rgb[3], uint32_t n, const float* lambda, float* out) {
    for (uint32_t i = 0; i < n; ++i) out[i] = SpectrumUtils::rgbToSpectrum(float3(rgb[0], rgb[1], rgb[2]), lambda[i]);
}
void oracle_spectrum_to_rgb(const float spectrum[4], const float wavelengths[4], uint32_t samples, float out[3]) {
    float3 r = SpectrumUtils::spectrumToRgb(float4(spectrum[0], spectrum[1], spectrum[2], spectrum[3]), float4(wavelengths[0], wavelengths[1], wavelengths[2], wavelengths[3]), samples);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void oracle_hwss_wavelengths(float hero, int hwss, float out[4]) {
    hwss_cfg().samples = hwss;
    hwss_t h = SpectrumSampler::hwssWavelengths(hero);
    for (int i = 0; i < 4; ++i) out[i] = h[i];
}
// Piecewise-linear table (.spd rows) -> SampledSpectrum -> SpectralProfile (SceneBuilder.cpp:911-999)
void oracle_gen_spectral_profile(uint32_t n, const float* wavelengths, const float* values, float scale, int computeICdf, PltptSpectralProfile* out) {
    PiecewiseLinearSpectrum pls;
    pls.wavelengths.assign(wavelengths, wavelengths + n);
    pls.values.assign(values, values + n);
    for (auto& v : pls.values) v *= scale;
    *out = genSpectralProfile(SampledSpectrum(pls), computeICdf != 0);
}
float oracle_eval_spectrum(const PltptSpectralProfile* p, float lambda) {
    ProfileTable t;
    t.profiles = p; t.count = 1;
    return t.evalSpectrum(0, lambda);
}
void oracle_oct_equal_area(uint32_t n, const float* dirs, float* enc, float* dec) {
    for (uint32_t i = 0; i < n; ++i) {
        float2 e = ndir_to_oct_equal_area_snorm(float3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        float3 d = oct_to_ndir_equal_area_snorm(e);
        enc[2 * i] = e.x; enc[2 * i + 1] = e.y;
        dec[3 * i] = d.x; dec[3 * i + 1] = d.y; dec[3 * i + 2] = d.z;
    }
}
void oracle_f16_roundtrip(uint32_t n, const float* in, uint16_t* h, float* back) {
    for (uint32_t i = 0; i < n; ++i) { h[i] = uint16_t(f32tof16(in[i])); back[i] = f16tof32(h[i]); }
}
uint32_t oracle_encode_normal2x16(const float n[3]) { return encodeNormal2x16(float3(n[0], n[1], n[2])); }

// ---- scene handle --------------------------------------------------------------------------
void* oracle_scene_create(const PltptSceneDesc* desc) {
    Scene* s = new Scene();
    s->init(*desc);
    return s;
}
void oracle_scene_destroy(void* s) { delete (Scene*)s; }
uint32_t oracle_scene_triangle_count(void* s) { return uint32_t(((Scene*)s)->tris.size()); }

// rays: float[8] {ox,oy,oz,tmin,dx,dy,dz,tmax}; hits: uint4 {instance, prim, asuint(u), asuint(v)}; miss => instance = ~0
void oracle_trace_closest(void* sp, uint32_t n, const float* rays, uint32_t* hits, float* hitT, int brute, int threads) {
    const Scene* s = (const Scene*)sp;
    parallel_for(n, threads, [&](uint32_t i) {
        const float* r = rays + 8 * size_t(i);
        const float3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
        RayHit h = brute ? s->traceClosestBrute(o, d, r[3], r[7]) : s->traceClosest(o, d, r[3], r[7]);
        hits[4 * size_t(i)] = h.hit.instanceID;
        hits[4 * size_t(i) + 1] = h.hit.valid() ? h.hit.primitiveIndex : 0;
        hits[4 * size_t(i) + 2] = h.hit.valid() ? asuint(h.hit.barycentrics.x) : 0;
        hits[4 * size_t(i) + 3] = h.hit.valid() ? asuint(h.hit.barycentrics.y) : 0;
        if (hitT) hitT[i] = h.hit.valid() ? h.t : 0.f;
    });
}
void oracle_trace_any(void* sp, uint32_t n, const float* rays, uint32_t* occluded, int brute, int threads) {
    const Scene* s = (const Scene*)sp;
    parallel_for(n, threads, [&](uint32_t i) {
        const float* r = rays + 8 * size_t(i);
        const float3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
        occluded[i] = (brute ? s->traceAnyBrute(o, d, r[3], r[7]) : s->traceAny(o, d, r[3], r[7])) ? 1u : 0u;
    });
}
// All exact-t ties of the closest hit (SURVEY: "exact ties are listed"): returns the count of triangles sharing the minimum t.
uint32_t oracle_count_ties(void* sp, const float* ray) {
    const Scene* s = (const Scene*)sp;
    const float3 o(ray[0], ray[1], ray[2]), d(ray[4], ray[5], ray[6]);
    RayHit best = s->traceClosestBrute(o, d, ray[3], ray[7]);
    if (!best.hit.valid()) return 0;
    const Scene::RayPrep r = Scene::prepRay(o, d);
    uint32_t c = 0;
    for (uint32_t g = 0; g < s->tris.size(); ++g) {
        float t, u, v;
        if (Scene::intersectTri(r, s->tris[g], ray[3], ray[7], t, u, v) && t == best.t) ++c;
    }
    return c;
}

// ---- BSDF sweeps (same contract as pltpt_bsdf_eval_mueller / pltpt_bsdf_sample in include/pltpt.h) ----
static ShadingData canonicalSd(const Scene* s, uint32_t materialID, float3 V) {
    ShadingData sd;
    sd.posW = float3(0.f); sd.V = V;
    sd.N = float3(0, 0, 1); sd.T = float3(1, 0, 0); sd.B = float3(0, 1, 0);
    sd.faceN = sd.N; sd.frontFacing = true; sd.materialID = materialID;
    sd.mtlHeader[0] = s->d.materialData[materialID].header[0];
    sd.mtlHeader[1] = s->d.materialData[materialID].header[1];
    return sd;
}
void oracle_bsdf_eval_mueller(void* sp, uint32_t materialID, uint32_t lobe, uint32_t n, uint32_t groups, const float* wi, const float* wo,
                              const float* hero, float* mueller, int hwss, float maxBeamOmega, int threads) {
    const Scene* s = (const Scene*)sp;
    parallel_for(n, threads, [&](uint32_t i) {
        hwss_cfg().samples = hwss;
        const float3 Wi(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), Wo(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        const ShadingData sd = canonicalSd(s, materialID, Wi);
        const MaterialInstance bsdf = s->getMaterialInstance(sd);
        for (uint32_t g = 0; g < groups; ++g) {
            const hwss_t wl = SpectrumSampler::hwssWavelengths(hero[size_t(g) * n + i]);
            for (int c = 0; c < 4; ++c) {
                PLTBeam beam = pltSourceBeamArea(Wi * 1.f, -Wi, 1.f, 1.f, hwss_t(1.f), wl, maxBeamOmega, false, uint32_t(hwss));
                float4 e(0.f);
                e[c] = 1.f;
                for (int j = 0; j < hwss; ++j) beam.setSP(j, e);
                beam.propagate(sd.posW);
                bsdf.eval(sd, Wo, lobe, beam);
                for (int j = 0; j < hwss; ++j) {
                    const float4 SP = beam.getSP(j);
                    float* o = mueller + ((size_t(i) * groups + g) * hwss + j) * 16;
                    for (int r = 0; r < 4; ++r) o[4 * r + c] = SP[r];
                }
            }
        }
    });
}
void oracle_bsdf_sample(void* sp, uint32_t materialID, uint32_t lobeMask, uint32_t n, const float* wi, const float* wavelengths, uint32_t seed, float* out, int hwss, int threads) {
    const Scene* s = (const Scene*)sp;
    parallel_for(n, threads, [&](uint32_t i) {
        hwss_cfg().samples = hwss;
        const float3 Wi(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        const ShadingData sd = canonicalSd(s, materialID, Wi);
        const MaterialInstance bsdf = s->getMaterialInstance(sd);
        SampleGenerator sg(i & 0xffff, i >> 16, seed);
        BSDFSample r;
        const bool valid = bsdf.sample(sd, sg, r, wavelengths[i], lobeMask);
        float* o = out + 8 * size_t(i);
        o[0] = r.wo.x; o[1] = r.wo.y; o[2] = r.wo.z; o[3] = r.pdf; o[4] = asfloat(r.lobe); o[5] = r.weight; o[6] = valid ? 1.f : 0.f;
        // NB argument naming of IMaterialInstance.pdf(sd, toLight, ...): sd.V is the eye side (PLTConductorBSDF.slang:205-208)
        o[7] = valid ? bsdf.pdf(sd, r.wo, wavelengths[i]) : 0.f;
    });
}

// ---- rendering -----------------------------------------------------------------------------
// Renders one frame (sample + solve stages). Any output pointer may be null except color when doSolve != 0.
// Returns the number of IMaterialInstance.eval x wavelength evaluations performed.
uint64_t oracle_render_frame(void* sp, const PltptOptions* opt, const PltptCameraData* cam, uint32_t frameCount, uint32_t width, uint32_t height,
                             float* color, float* albedo, float* normal, PltptBounce* bounces, int doSolve, int threads) {
    const Scene* s = (const Scene*)sp;
    std::vector<PltptBounce> localBounces;
    if (!bounces) { localBounces.resize(size_t(width) * height * (opt->maxBounces + 1)); bounces = localBounces.data(); }
    std::memset(bounces, 0, sizeof(PltptBounce) * size_t(width) * height * (opt->maxBounces + 1));
    std::atomic<uint64_t> evals{0};
    parallel_for(height, threads, [&](uint32_t y) {
        hwss_cfg().samples = int(opt->HWSS);
        Integrator I;
        I.scene = s; I.opt = *opt; I.cam = *cam; I.frameCount = frameCount; I.width = width; I.height = height;
        I.bounces = bounces; I.outColor = color; I.outAlbedo = albedo; I.outNormal = normal;
        for (uint32_t x = 0; x < width; ++x) {
            I.samplePixel(x, y);
            if (doSolve) I.solvePixel(x, y);
        }
        evals += I.statEvals;
    });
    return evals.load();
}

// ---- environment map (plt_envmap.h): dirs float3 -> out float4 {eval rgb, evalPdf}; rnd float2 -> out float8 {dir, pdf, Le rgb, 0}
void oracle_envmap_eval(void* sp, uint32_t n, const float* dirs, float* out) {
    const Scene& sc = *(Scene*)sp;
    for (uint32_t i = 0; i < n; ++i) {
        const float3 d(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        const float3 c = sc.env.eval(d);
        out[4 * i] = c.x; out[4 * i + 1] = c.y; out[4 * i + 2] = c.z; out[4 * i + 3] = sc.env.evalPdf(d);
    }
}
void oracle_envmap_sample(void* sp, uint32_t n, const float* rnd, float* out) {
    const Scene& sc = *(Scene*)sp;
    for (uint32_t i = 0; i < n; ++i) {
        EnvMap::Sample s;
        sc.env.sample(float2(rnd[2 * i], rnd[2 * i + 1]), s);
        float* o = out + 8 * i;
        o[0] = s.dir.x; o[1] = s.dir.y; o[2] = s.dir.z; o[3] = s.pdf; o[4] = s.Le.x; o[5] = s.Le.y; o[6] = s.Le.z; o[7] = 0.f;
    }
}
void oracle_envmap_importance(void* sp, float* out) {
    const Scene& sc = *(Scene*)sp;
    if (!sc.env.valid()) return;
    std::memcpy(out, sc.env.mips[0].data(), sc.env.mips[0].size() * sizeof(float));
}

// material-texture filter (plt_bsdf.h TextureSystem::sampleTexture): uv float2 -> float4
void oracle_texture_sample(void* sp, uint32_t handle, uint32_t samplerID, const float uniformValue[4], uint32_t n, const float* uv, float* out) {
    const Scene& sc = *(Scene*)sp;
    for (uint32_t i = 0; i < n; ++i) {
        const float4 c = sc.textures.sampleTexture(handle, samplerID, float2(uv[2 * i], uv[2 * i + 1]), float4(uniformValue[0], uniformValue[1], uniformValue[2], uniformValue[3]));
        out[4 * i] = c.x; out[4 * i + 1] = c.y; out[4 * i + 2] = c.z; out[4 * i + 3] = c.w;
    }
}

void oracle_colour_map(int which, uint32_t n, const float* vals, float* out) {  // 0 = ccMagma, 1 = ccInferno (plt_debug.h)
    for (uint32_t i = 0; i < n; ++i) { const float3 c = which == 0 ? ccMagma(vals[i]) : ccInferno(vals[i]); out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z; }
}

int oracle_hardware_threads() { return int(std::thread::hardware_concurrency()); }

}  // extern "C"
