// ingest.cu — scene ingestion without Falcor: what Falcor's SceneBuilder / LightCollection / EmissivePowerSampler
// produce on the host and in compute passes, built here behind the C ABI so that a host shell (or the test harness) can go from
// `.spd` tables, PLT material parameters and emissive meshes to the buffers pltpt_set_scene takes.
//   .spd file -> PiecewiseLinearSpectrum      Source/Falcor/Utils/Color/Spectrum.cpp:89-112, Spectrum.h:88-106
//   -> SampledSpectrum -> SpectralProfile      Utils/Color/SampledSpectrum.h:57-68,303-317; Scene/SceneBuilder.cpp:911-960
//   PLT material parameters -> MaterialDataBlob Scene/Material/PLT/*Material.{cpp,slang} (fp16 slots, see oracle/plt_bsdf.h MaterialInstance)
//   emissive meshes -> PackedEmissiveTriangle + flux (GPU)   Scene/Lights/BuildTriangleList.cs.slang, FinalizeIntegration.cs.slang:64-110
//   flux -> alias table                         Rendering/Lights/EmissivePowerSampler.cpp:86-170 (without the final random permutation)
#include "context.h"
#include "plt_scene.cuh"
#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

using namespace plt;

namespace {
int ifail(int code, const std::string& msg) { setError(msg); return code; }

// ---- spectra ---------------------------------------------------------------------------------------------------------------
// Falcor's HOST lerp is (1 - t) * a + t * b (Utils/Math/Common.h:54-57), not the shader form a + t * (b - a): the profile tables are held to the
// reference's own C++ byte for byte (tests/golden/ref_spectral_profiles.npz), so the form matters
inline float hostLerp(float a, float b, float t) { return (1.f - t) * a + t * b; }
struct Pls { const float* w; const float* v; uint32_t n; };
float plsEval(const Pls& s, float lam) {  // PiecewiseLinearSpectrum::eval
    if (s.n == 0 || lam < s.w[0] || lam > s.w[s.n - 1]) return 0.f;
    const uint32_t it = uint32_t(std::lower_bound(s.w, s.w + s.n, lam) - s.w);
    if (it == 0) return s.v[0];
    const uint32_t i = it - 1;
    const float t = (lam - s.w[i]) / (s.w[i + 1] - s.w[i]);
    return hostLerp(s.v[i], s.v[i + 1], t);
}
struct Sampled { float start, end; std::vector<float> samples; };
Sampled sampledFromPls(const Pls& s) {  // SampledSpectrum(PiecewiseLinearSpectrum), bin >= 5 nm
    Sampled r;
    r.start = s.w[0]; r.end = s.w[s.n - 1];
    float least = r.end - r.start;
    for (uint32_t i = 1; i < s.n; ++i) least = std::min(least, s.w[i] - s.w[i - 1]);
    const float bin = std::max(5.f, least / 2.f);
    const uint32_t n = std::max<uint32_t>(1u, uint32_t((r.end - r.start) / bin + 0.5f));
    r.samples.resize(n);
    for (uint32_t i = 0; i < n; ++i) r.samples[i] = plsEval(s, n > 1 ? r.start + ((r.end - r.start) / float(n - 1)) * float(i) : r.start);
    return r;
}
float sampledEval(const Sampled& s, float lam) {  // SampledSpectrum::eval
    if (lam < s.start || lam > s.end) return 0.f;
    const float x = ((lam - s.start) / (s.end - s.start)) * (float(s.samples.size()) - 1.0f);
    const int i = int(std::floor(x));
    if (size_t(i) + 1 >= s.samples.size()) return s.samples.back();
    const float w = x - float(i);
    return hostLerp(s.samples[i], s.samples[i + 1], w);
}
// SpectrumUtils::toRGB_D65 (SpectrumUtils.h:78-151): trapezoid sum of XYZ(lambda) D65(lambda) s(lambda) -> Rec.709 / Y_D65.
// DEVIATION (same as the harness, pltfalcor_b200/scene.py): the reference integrates against the CIE-1931 1 nm tables; this uses the
// analytic fit the device code uses (SpectrumUtils.slang:43-49).  `rgb` only feeds the per-triangle flux, i.e. the light-selection
// pdf, which the estimator divides out again.
void profileRgb(const Sampled& s, float rgb[3]) {
    static const double d65[69] = {49.9755, 52.3118, 54.6482, 68.7015, 82.7549, 87.1204, 91.486, 92.4589, 93.4318, 90.057, 86.6823, 95.7736, 104.865, 110.936,
        117.008, 117.41, 117.812, 116.336, 114.861, 115.392, 115.923, 112.367, 108.811, 109.082, 109.354, 108.578, 107.802, 106.296,
        104.79, 106.239, 107.689, 106.047, 104.405, 104.225, 104.046, 102.023, 100., 98.1671, 96.3342, 96.0611, 95.788, 92.2368, 88.6856,
        89.3459, 90.0062, 89.8026, 89.5991, 88.6489, 87.6987, 85.4936, 83.2886, 83.4939, 83.6992, 81.863, 80.0268, 80.1207, 80.2146,
        81.2462, 82.2778, 80.281, 78.2842, 74.0027, 69.7213, 70.6652, 71.6091, 72.979, 74.349, 67.9765, 61.604};
    rgb[0] = rgb[1] = rgb[2] = 0.f;
    const size_t n = s.samples.size();
    if (n < 2) return;
    auto G = [](double x, double mu, double s1, double s2) { const double t = (x - mu) * (x < mu ? s1 : s2); return std::exp(-0.5 * t * t); };
    const double delta = (double(s.end) - double(s.start)) / (double(n) - 1.0);
    double acc[3] = {0, 0, 0};
    for (size_t q = 0; q < n; ++q) {
        const double lam = std::min(double(s.start) + delta * double(q), double(s.end));
        if (lam < 380.0 || lam > 720.0) continue;
        const double x = (lam - 380.0) / 5.0;
        const int i0 = std::min(67, int(x));
        const double d = d65[i0] + (d65[i0 + 1] - d65[i0]) * (x - i0);
        const double xyz[3] = {0.362 * G(lam, 442.0, 0.0624, 0.0374) + 1.056 * G(lam, 599.8, 0.0264, 0.0323) - 0.065 * G(lam, 501.1, 0.0490, 0.0382),
                               0.821 * G(lam, 568.8, 0.0213, 0.0247) + 0.286 * G(lam, 530.9, 0.0613, 0.0322),
                               1.217 * G(lam, 437.0, 0.0845, 0.0278) + 0.681 * G(lam, 459.0, 0.0385, 0.0725)};
        const double wq = d * double(s.samples[q]) * delta * ((q == 0 || q == n - 1) ? 0.5 : 1.0);
        for (int c = 0; c < 3; ++c) acc[c] += xyz[c] * wq;
    }
    static const double M[3][3] = {{3.240969941904522, -1.537383177570094, -0.4986107602930032}, {-0.9692436362808803, 1.875967501507721, 0.04155505740717569},
                                   {0.05563007969699373, -0.2039769588889765, 1.056971514242878}};
    for (int r = 0; r < 3; ++r) rgb[r] = float((M[r][0] * acc[0] + M[r][1] * acc[1] + M[r][2] * acc[2]) / 10567.0762);
}

// ---- fp16 (round to nearest even; numpy / f32tof16 semantics) ------------------------------------------------------------------
uint16_t f2h(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t e = int32_t((x >> 23) & 0xff) - 127 + 15;
    uint32_t m = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return uint16_t(sign | 0x7c00u | (m ? 0x200u : 0u));
    if (e >= 31) return uint16_t(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return uint16_t(sign);
        m |= 0x800000u;
        const uint32_t shift = uint32_t(14 - e);
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) ++h;
        return uint16_t(sign | h);
    }
    uint32_t h = (uint32_t(e) << 10) | (m >> 13);
    const uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return uint16_t(sign | h);
}
void half4(uint16_t dst[4], float a, float b, float c, float d) { dst[0] = f2h(a); dst[1] = f2h(b); dst[2] = f2h(c); dst[3] = f2h(d); }

// ---- light collection ------------------------------------------------------------------------------------------------------------
__device__ uint32_t dF2h(float f) { return uint32_t(__half_as_ushort(__float2half_rn(f))); }
__device__ uint32_t dPackSnorm16(float v) {  // packSnorm16 (Utils/Math/FormatConversion.slang): round(clamp(v,-1,1) * 32767)
    v = fminf(fmaxf(isfinite(v) ? v : 0.f, -1.f), 1.f);
    return uint32_t(int(rintf(v * 32767.f))) & 0xffffu;
}
__device__ uint32_t dEncodeNormal2x16(float3 n) {  // ndir_to_oct_snorm + packSnorm2x16 (MathHelpers.slang:161-170)
    const float inv = 1.f / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    float2 p = mk2(n.x * inv, n.y * inv);
    if (n.z < 0.f) p = mk2((1.f - fabsf(p.y)) * (p.x >= 0.f ? 1.f : -1.f), (1.f - fabsf(p.x)) * (p.y >= 0.f ? 1.f : -1.f));
    return dPackSnorm16(p.x) | (dPackSnorm16(p.y) << 16);
}
// One thread per (emissive instance, primitive): BuildTriangleList.cs.slang + FinalizeIntegration.cs.slang for untextured emitters.
__global__ void __launch_bounds__(128) k_build_emissive(DevScene sc, const uint32_t* __restrict__ instList, const uint32_t* __restrict__ instFirst, uint32_t instCount,
                                                        uint32_t triCount, const float* __restrict__ rgbOfMaterial, PltptEmissiveTriangle* __restrict__ tris,
                                                        PltptEmissiveFlux* __restrict__ flux) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    // which emissive instance: the last one whose first triangle is <= t
    uint32_t lo = 0, hi = instCount;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (instFirst[mid] <= t) lo = mid; else hi = mid; }
    const uint32_t inst = instList[lo], prim = t - instFirst[lo];
    const PltptGeometryInstance gi = sc.instances[inst];
    uint32_t idx[3];
    getIndices(sc, gi, prim, idx);
    const float* M = sc.world + 16 * size_t(gi.globalMatrixID);
    float3 pw[3];
    float2 uv[3];
    for (int k = 0; k < 3; ++k) {
        const PltptPackedVertex& v = sc.vertices[idx[k]];
        pw[k] = mulPoint(M, mk3(v.position[0], v.position[1], v.position[2]));
        uv[k] = mk2(v.texCrd[0], v.texCrd[1]);
    }
    float3 n = cross(pw[1] - pw[0], pw[2] - pw[0]);
    const float len = length(n);
    if (gi.flags & PLTPT_INSTANCE_FLAG_WORLD_FRONT_FACE_CW) n = mk3(-n.x, -n.y, -n.z);   // Scene.slang:493-494 (computeFaceNormalAndAreaW): winding in world space
    PltptEmissiveTriangle e;
    for (int k = 0; k < 3; ++k) {
        e.posAndTexCoords[k][0] = pw[k].x; e.posAndTexCoords[k][1] = pw[k].y; e.posAndTexCoords[k][2] = pw[k].z;
        e.posAndTexCoords[k][3] = __uint_as_float(dF2h(uv[k].x) | (dF2h(uv[k].y) << 16));
    }
    const float area = 0.5f * len;
    e.normal = dEncodeNormal2x16(n / len);
    e.area = __float_as_uint(area);
    e.materialID = gi.materialID;
    e.lightIdx = lo;   // index of the mesh light (LightCollection.cpp:481-485: one dispatch per emissive instance, in instance order)
    tris[t] = e;
    const float* rgb = rgbOfMaterial + 3 * size_t(gi.materialID);
    const float r = fmaxf(rgb[0], 0.f), g = fmaxf(rgb[1], 0.f), b = fmaxf(rgb[2], 0.f);
    PltptEmissiveFlux f;
    f.flux = ((r * 0.2126f + g * 0.7152f) + b * 0.0722f) * area * kPi;
    f.averageRadiance[0] = r; f.averageRadiance[1] = g; f.averageRadiance[2] = b;
    flux[t] = f;
}
}  // namespace

extern "C" int pltpt_spd_load(const char* path, float* wavelengths, float* values, uint32_t capacity, uint32_t* count) {
    if (!path || !count) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    FILE* f = std::fopen(path, "r");
    if (!f) return ifail(PLTPT_ERR_INVALID_ARGUMENT, std::string("Path could not be read '") + path + "'.");  // Spectrum.cpp:93
    char line[512];
    uint32_t n = 0;
    while (std::fgets(line, sizeof line, f)) {
        const char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '#' || *p == '\n' || *p == '\r' || *p == 0) continue;
        float a, b;
        if (std::sscanf(p, "%f %f", &a, &b) != 2) break;  // stop at the first malformed row
        if (n < capacity && wavelengths && values) { wavelengths[n] = a; values[n] = b; }
        ++n;
    }
    std::fclose(f);
    *count = n;
    if (n == 0) return ifail(PLTPT_ERR_INVALID_ARGUMENT, std::string("Path could not be read '") + path + "'.");
    return PLTPT_OK;
}

extern "C" int pltpt_gen_spectral_profile(uint32_t n, const float* wavelengths, const float* values, float scale, int computeICdf, PltptSpectralProfile* out) {
    if (!wavelengths || !values || !out || n == 0) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "null or empty spectrum");
    std::vector<float> v(values, values + n);
    for (float& x : v) x *= scale;  // PiecewiseLinearSpectrum::scale
    const Sampled ss = sampledFromPls(Pls{wavelengths, v.data(), n});
    PltptSpectralProfile sp;
    std::memset(&sp, 0, sizeof sp);
    profileRgb(ss, sp.rgb);
    sp.intensity = 1.f;
    sp.minWavelength = std::max(ss.start, 380.f);
    sp.maxWavelength = std::min(ss.end, 720.f);
    sp.bins = std::min<uint32_t>(128u, uint32_t(float(ss.samples.size()) * (sp.maxWavelength - sp.minWavelength) / (ss.end - ss.start) + .5f));
    *out = sp;
    if (sp.bins == 0) return PLTPT_OK;
    const float r = sp.bins > 1 ? 1.f / float(sp.bins - 1) : 1.f;
    float sum = 0.f;
    for (uint32_t i = 0; i < sp.bins; ++i) {
        sp.pdf[i] = sampledEval(ss, hostLerp(sp.minWavelength, sp.maxWavelength, float(i) * r));
        sum += sp.pdf[i];
    }
    if (sum < 1.192092896e-07f || !computeICdf) { *out = sp; return PLTPT_OK; }
    for (uint32_t i = 0; i < sp.bins; ++i) sp.pdf[i] /= sum;
    sp.intensity = sum;
    float cdfa = 0.f, cdfb = sp.pdf[0];
    for (uint32_t pdfi = 0, icdfi = 0; icdfi < sp.bins; ++icdfi) {
        const float cdf = float(icdfi) * r;
        while (cdf >= cdfb && pdfi + 1 < sp.bins) { cdfa = cdfb; ++pdfi; cdfb += sp.pdf[pdfi]; }
        const float l = std::min(1.f, (cdf - cdfa) / (cdfb - cdfa));
        const float a = float(std::max<uint32_t>(1u, pdfi)) - 1.f;
        sp.icdf[icdfi] = hostLerp(a, float(pdfi), l) * r;
    }
    *out = sp;
    return PLTPT_OK;
}

extern "C" int pltpt_pack_material(const PltptMaterialParams* p, PltptMaterialBlob* out) {
    if (!p || !out) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (p->type < PLTPT_MATERIAL_DIFFUSE || p->type > PLTPT_MATERIAL_MULTILAYERED_STACK) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "not a PLT material type");
    PltptMaterialBlob m;
    std::memset(&m, 0, sizeof m);
    const bool emissive = p->type == PLTPT_MATERIAL_DIFFUSE && p->emissionSpectralId != PLTPT_NO_SPECTRAL_PROFILE;
    m.header[0] = (p->type & 0xffffu) | (0xffu << 20) | (p->doubleSided ? 1u << 28 : 0u) | (emissive ? 1u << 30 : 0u) | (1u << 31);  // MaterialHeader packedData.x
    m.header[1] = f2h(0.5f);  // alpha threshold 0.5, AlphaMode::Opaque (Material.cpp:63-65)
    half4(m.baseColor, 1.f, 1.f, 1.f, 1.f);
    m.transmission[0] = m.transmission[1] = m.transmission[2] = f2h(1.f);
    m.IoR = f2h(1.5f);
    m.emissionSpectralId = m.iorNSpectralId = m.iorKSpectralId = PLTPT_NO_SPECTRAL_PROFILE;
    const float* c = p->color;
    switch (p->type) {
        case PLTPT_MATERIAL_DIFFUSE:
            half4(m.baseColor, c[0], c[1], c[2], 1.f);
            m.emissionSpectralId = p->emissionSpectralId;
            break;
        case PLTPT_MATERIAL_CONDUCTOR:
            half4(m.baseColor, c[0], c[1], c[2], 1.f);
            half4(m.specular, p->gamma, p->roughness, 1.f, 1.f);
            m.IoR = f2h(p->extEta);
            m.iorNSpectralId = p->iorNSpectralId; m.iorKSpectralId = p->iorKSpectralId;
            break;
        case PLTPT_MATERIAL_DIELECTRIC:
            half4(m.baseColor, c[0], c[1], c[2], 1.f);
            half4(m.specular, p->dispersionB, p->roughness, 0.f, p->extEta);
            m.transmission[0] = f2h(p->transmission[0]); m.transmission[1] = f2h(p->transmission[1]); m.transmission[2] = f2h(p->transmission[2]);
            m.IoR = f2h(p->ior);
            break;
        case PLTPT_MATERIAL_OPAQUE_DIELECTRIC:
            half4(m.baseColor, c[0], c[1], c[2], 1.f);
            half4(m.specular, 0.f, p->roughness, 0.f, 0.f);
            break;
        case PLTPT_MATERIAL_THIN_DIELECTRIC:
            half4(m.baseColor, p->birefringence, p->opticAxis[0], p->opticAxis[1], 1.f);
            half4(m.specular, p->birefringenceScale, 0.f, 0.f, p->tau);
            m.IoR = f2h(p->ior);
            break;
        case PLTPT_MATERIAL_COATED_CONDUCTOR:
        case PLTPT_MATERIAL_COATED_OPAQUE_DIELECTRIC:
            half4(m.baseColor, c[0], c[1], c[2], 1.f);
            half4(m.specular, 0.f, p->roughness, 1.f, 0.f);
            half4(m.data1, p->tau, 0.f, 0.f, p->coatIor);
            m.IoR = f2h(p->extEta);
            if (p->type == PLTPT_MATERIAL_COATED_CONDUCTOR) { m.iorNSpectralId = p->iorNSpectralId; m.iorKSpectralId = p->iorKSpectralId; }
            break;
        case PLTPT_MATERIAL_GRATED_CONDUCTOR:
        case PLTPT_MATERIAL_GRATED_OPAQUE_DIELECTRIC:
            half4(m.baseColor, c[0], c[1], c[2], 1.f);
            half4(m.specular, p->gratingQScale, p->gratingPowerMultiplier, 0.f, 0.f);
            half4(m.data1, p->gratingPitch[0], p->gratingPitch[1], p->gratingDirection, 0.f);
            m.data1[3] = uint16_t(((p->gratingType & 0xffu) << 8) | (p->gratingLobes & 0xffu));
            half4(m.data2, p->gratingQ, p->roughness, 0.f, 0.f);
            if (p->type == PLTPT_MATERIAL_GRATED_CONDUCTOR) { m.IoR = f2h(p->extEta); m.iorNSpectralId = p->iorNSpectralId; m.iorKSpectralId = p->iorKSpectralId; }
            break;
        case PLTPT_MATERIAL_MULTILAYERED_STACK:
            half4(m.baseColor, c[0], c[1], c[2], 1.f);
            half4(m.data1, p->layerThickness[0], p->layerThickness[1], 0.f, 0.f);
            m.data1[3] = uint16_t(p->layers & 0xffffu);
            half4(m.data2, p->layerIor[0], p->layerIor[1], 0.f, 0.f);
            m.IoR = f2h(p->extEta);
            break;
        default: break;
    }
    *out = m;
    return PLTPT_OK;
}

// EmissivePowerSampler::generateAliasTable (EmissivePowerSampler.cpp:86-170) without the final random permutation: the table is *a*
// valid alias table either way; parity is stated on the induced pdf flux / sum.
static void generateAliasTable(std::vector<float> w, std::vector<uint32_t>& table, float& weightSum) {
    const uint32_t N = uint32_t(w.size());
    double sum = 0.0;
    for (float f : w) sum += f;
    for (float& f : w) f *= float(N) / float(sum);
    std::vector<uint32_t> perm(N);
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return w[a] < w[b]; });
    std::vector<float> thresholds(N, 0.f);
    std::vector<uint32_t> redirect(N, 0u);
    uint32_t head = 0, tail = N - 1;
    if (N == 1) { thresholds[0] = 1.f; redirect[0] = 0; }
    while (head != tail) {
        const uint32_t i = perm[head], j = perm[tail];
        thresholds[i] = w[i];
        redirect[i] = j;
        w[j] -= 1.0f - w[i];
        if (head == tail - 1) { thresholds[j] = 1.0f; redirect[j] = j; break; }
        else if (w[j] < 1.0f) { std::swap(perm[head], perm[tail]); tail--; }
        else head++;
    }
    table.resize(size_t(N) * 2);
    for (uint32_t i = 0; i < N; ++i) {
        const uint32_t prob = uint32_t(f2h(thresholds[i])) << 16;
        const uint32_t lx = redirect[i] & 0xFFFFFFu, ly = i & 0xFFFFFFu;
        table[2 * size_t(i)] = prob | ((lx >> 8) & 0xFFFFu);
        table[2 * size_t(i) + 1] = ((lx & 0xFFu) << 24) | ly;
    }
    weightSum = float(sum);
}

// Builds the LightCollection + EmissivePowerSampler tables of the scene held by `h` (pltpt_set_scene must have run; the scene may
// have been given without them) and installs them.  The results can be read back with pltpt_get_light_collection.
extern "C" int pltpt_build_light_collection(PltptHandle h) {
    if (!h) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->hasScene) return ifail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    PLT_CUDA(cudaSetDevice(h->device));
    if (h->waitHostAsync()) return ifail(PLTPT_ERR_CUDA, "a frame queued by pltpt_render_frame_host_async failed");
    if (int r = syncFrameStats(h)) return r;  // frames in flight read the light tables that are rebuilt here
    PLT_CUDA(cudaDeviceSynchronize());
    const uint32_t instCount = uint32_t(h->hInstances.size());
    std::vector<uint32_t> instList, instFirst, offsets(instCount, 0xffffffffu);
    uint32_t total = 0;
    for (uint32_t i = 0; i < instCount; ++i) {
        const uint32_t mat = h->hInstances[i].materialID;
        if (mat >= h->hMaterialEmission.size() || h->hMaterialEmission[mat] == PLTPT_NO_SPECTRAL_PROFILE) continue;
        offsets[i] = total;
        instList.push_back(i);
        instFirst.push_back(total);
        total += h->hInstTriOffset[i + 1] - h->hInstTriOffset[i];
    }
    h->scene.etriCount = h->scene.activeTriCount = total;
    h->emissiveTriangleCount = total;
    PLT_CUDA(h->perMeshOffset.ensure(std::max<size_t>(4, size_t(instCount) * 4)));
    PLT_CUDA(cudaMemcpy(h->perMeshOffset.p, offsets.data(), size_t(instCount) * 4, cudaMemcpyHostToDevice));
    h->scene.perMeshInstanceOffset = h->perMeshOffset.as<uint32_t>();
    if (total == 0) { h->scene.invWeightsSum = 0.f; return PLTPT_OK; }
    DevBuf dList, dFirst, dRgb;
    PLT_CUDA(dList.ensure(instList.size() * 4));
    PLT_CUDA(dFirst.ensure(instFirst.size() * 4));
    PLT_CUDA(dRgb.ensure(h->hMaterialRgb.size() * 4));
    PLT_CUDA(cudaMemcpy(dList.p, instList.data(), instList.size() * 4, cudaMemcpyHostToDevice));
    PLT_CUDA(cudaMemcpy(dFirst.p, instFirst.data(), instFirst.size() * 4, cudaMemcpyHostToDevice));
    PLT_CUDA(cudaMemcpy(dRgb.p, h->hMaterialRgb.data(), h->hMaterialRgb.size() * 4, cudaMemcpyHostToDevice));
    PLT_CUDA(h->etris.ensure(size_t(total) * sizeof(PltptEmissiveTriangle)));
    PLT_CUDA(h->eflux.ensure(size_t(total) * sizeof(PltptEmissiveFlux)));
    PLT_CUDA(h->activeTris.ensure(size_t(total) * 4));
    k_build_emissive<<<(total + 127) / 128, 128>>>(h->scene, dList.as<uint32_t>(), dFirst.as<uint32_t>(), uint32_t(instList.size()), total, dRgb.as<float>(),
                                                   h->etris.as<PltptEmissiveTriangle>(), h->eflux.as<PltptEmissiveFlux>());
    h->stats.kernelLaunches++;
    PLT_CUDA(cudaGetLastError());
    std::vector<PltptEmissiveFlux> flux(total);
    PLT_CUDA(cudaMemcpy(flux.data(), h->eflux.p, size_t(total) * sizeof(PltptEmissiveFlux), cudaMemcpyDeviceToHost));
    std::vector<float> w(total);
    for (uint32_t i = 0; i < total; ++i) w[i] = flux[i].flux;
    std::vector<uint32_t> table, active(total);
    std::iota(active.begin(), active.end(), 0u);
    float weightSum = 0.f;
    generateAliasTable(w, table, weightSum);
    PLT_CUDA(h->alias.ensure(table.size() * 4));
    PLT_CUDA(cudaMemcpy(h->alias.p, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
    PLT_CUDA(cudaMemcpy(h->activeTris.p, active.data(), size_t(total) * 4, cudaMemcpyHostToDevice));
    h->scene.etris = h->etris.as<PltptEmissiveTriangle>();
    h->scene.eflux = h->eflux.as<PltptEmissiveFlux>();
    h->scene.activeTris = h->activeTris.as<uint32_t>();
    h->scene.alias = h->alias.as<uint2>();
    h->scene.invWeightsSum = 1.0f / weightSum;
    h->ownLightCollection = true;
    return PLTPT_OK;
}

// Animated instance transforms (the reference refits / rebuilds its TLAS on demand, Source/Falcor/Scene/Scene.cpp:3624-3700; geometry
// here lives in ONE world-space BVH; pltpt_set_bvh_update_mode picks a full GPU rebuild or a bottom-up refit of it): new world and
// inverse-transpose matrices for the scene held by `h`, BVH rebuilt, and -- when the library built the light collection itself --
// the emissive triangles re-derived.  A light collection supplied by the caller is the caller's to refresh (pltpt_set_scene).
extern "C" int pltpt_update_transforms(PltptHandle h, const float* worldMatrices, const float* inverseTransposeWorldMatrices, uint32_t matrixCount, uint32_t pointersAreDevice) {
    if (!h || !worldMatrices || !inverseTransposeWorldMatrices) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (!h->hasScene) return ifail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    if (matrixCount != h->matrixCount) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "matrixCount differs from the scene's matrix table (" + std::to_string(h->matrixCount) + ")");
    for (const auto& gi : h->hInstances)
        if (gi.globalMatrixID >= matrixCount) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "an instance refers to a matrix beyond matrixCount");
    PLT_CUDA(cudaSetDevice(h->device));
    if (h->waitHostAsync()) return ifail(PLTPT_ERR_CUDA, "a frame queued by pltpt_render_frame_host_async failed");
    if (int r = syncFrameStats(h)) return r;  // frames in flight read the old BVH
    PLT_CUDA(cudaDeviceSynchronize());
    const cudaMemcpyKind kind = pointersAreDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    PLT_CUDA(cudaMemcpy(h->world.p, worldMatrices, size_t(matrixCount) * 64, kind));
    PLT_CUDA(cudaMemcpy(h->worldIT.p, inverseTransposeWorldMatrices, size_t(matrixCount) * 64, kind));
    if (int r = h->bvhUpdateMode == PLTPT_BVH_UPDATE_REFIT ? refitBvh(h, 0) : buildBvh(h, 0)) return r;
    if (h->ownLightCollection) return pltpt_build_light_collection(h);
    return PLTPT_OK;
}

extern "C" int pltpt_get_light_collection(PltptHandle h, PltptEmissiveTriangle* tris, PltptEmissiveFlux* flux, uint32_t* perMeshInstanceOffset, uint32_t* aliasTable,
                                          float* invWeightsSum, uint32_t capacity, uint32_t* count) {
    if (!h || !count) return ifail(PLTPT_ERR_INVALID_ARGUMENT, "null argument");
    if (!h->hasScene) return ifail(PLTPT_ERR_NO_SCENE, "pltpt_set_scene has not been called");
    PLT_CUDA(cudaSetDevice(h->device));
    const uint32_t n = h->scene.etriCount;
    *count = n;
    if (invWeightsSum) *invWeightsSum = h->scene.invWeightsSum;
    if (perMeshInstanceOffset && h->scene.perMeshInstanceOffset) PLT_CUDA(cudaMemcpy(perMeshInstanceOffset, h->scene.perMeshInstanceOffset, h->hInstances.size() * 4, cudaMemcpyDeviceToHost));
    if (n > capacity) return (tris || flux || aliasTable) ? ifail(PLTPT_ERR_INVALID_ARGUMENT, "capacity too small") : PLTPT_OK;
    if (tris && n) PLT_CUDA(cudaMemcpy(tris, h->scene.etris, size_t(n) * sizeof(PltptEmissiveTriangle), cudaMemcpyDeviceToHost));
    if (flux && n) PLT_CUDA(cudaMemcpy(flux, h->scene.eflux, size_t(n) * sizeof(PltptEmissiveFlux), cudaMemcpyDeviceToHost));
    if (aliasTable && n && h->scene.alias) PLT_CUDA(cudaMemcpy(aliasTable, h->scene.alias, size_t(n) * 8, cudaMemcpyDeviceToHost));
    return PLTPT_OK;
}
