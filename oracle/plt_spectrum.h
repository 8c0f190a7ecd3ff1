// ORACLE — TEST INFRASTRUCTURE ONLY (see plt_math.h header).
// plt_spectrum.h — sample generator, hero-wavelength spectral sampling, colour conversion, spectral profiles.
#pragma once
#include "plt_math.h"
#include "../include/pltpt.h"
#include <vector>
#include <string>

namespace orc {

// ---------------------------------------------------------------------------------------------
// PRNG: SplitMix64-seeded xoshiro128** (the 2018 variant multiplying state[0]).
//   Utils/Sampling/UniformSampleGenerator.slang:50-77, Pseudorandom/Xoshiro.slang:41-68,
//   Pseudorandom/SplitMix64.slang:45-77, Math/BitTricks.slang:45-61, SampleGeneratorInterface.slang:48-75
// ---------------------------------------------------------------------------------------------
inline uint32_t interleave_32bit(uint32_t vx, uint32_t vy) {
    uint32_t x = vx & 0x0000ffff, y = vy & 0x0000ffff;
    x = (x | (x << 8)) & 0x00FF00FF; x = (x | (x << 4)) & 0x0F0F0F0F; x = (x | (x << 2)) & 0x33333333; x = (x | (x << 1)) & 0x55555555;
    y = (y | (y << 8)) & 0x00FF00FF; y = (y | (y << 4)) & 0x0F0F0F0F; y = (y | (y << 2)) & 0x33333333; y = (y | (y << 1)) & 0x55555555;
    return x | (y << 1);
}
inline uint64_t splitmix64_next(uint64_t& state) {
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

struct SampleGenerator {
    uint32_t s[4];
    SampleGenerator() : s{0, 0, 0, 0} {}
    SampleGenerator(uint32_t px, uint32_t py, uint32_t sampleNumber) {
        uint64_t st = (uint64_t(sampleNumber) << 32) | uint64_t(interleave_32bit(px, py));
        uint64_t s0 = splitmix64_next(st), s1 = splitmix64_next(st);
        s[0] = uint32_t(s0); s[1] = uint32_t(s0 >> 32); s[2] = uint32_t(s1); s[3] = uint32_t(s1 >> 32);
    }
    uint32_t next() {
        const uint32_t result = rotl32(s[0] * 5, 7) * 9;
        const uint32_t t = s[1] << 9;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl32(s[3], 11);
        return result;
    }
};
inline float sampleNext1D(SampleGenerator& sg) { return float(sg.next() >> 8) * 0x1p-24f; }
inline float2 sampleNext2D(SampleGenerator& sg) { float2 r; r.x = sampleNext1D(sg); r.y = sampleNext1D(sg); return r; }
inline float3 sampleNext3D(SampleGenerator& sg) { float3 r; r.x = sampleNext1D(sg); r.y = sampleNext1D(sg); r.z = sampleNext1D(sg); return r; }

// ---------------------------------------------------------------------------------------------
// HWSS: hwss_t = vector<float, HWSS_SAMPLES> (Utils/HWSS.slang). The oracle keeps 4 lanes and a
// run-time count `g_hwss` (set per call from the options); unused lanes are computed but ignored.
// ---------------------------------------------------------------------------------------------
typedef float4 hwss_t;
struct HwssConfig { int samples = 4; };
inline HwssConfig& hwss_cfg() { static thread_local HwssConfig c; return c; }
inline int hwss_samples() { return hwss_cfg().samples; }

namespace SpectrumUtils {
static constexpr float minWavelength = 380.0f, maxWavelength = 720.0f, meanWavelength = 550.f;

// SpectrumUtils.slang:66-70
inline float G(float x, float lambda, float invSigma1, float invSigma2) {
    float t = (x - lambda) * (x < lambda ? invSigma1 : invSigma2);
    return std::exp(-0.5f * t * t);
}
// SpectrumUtils.slang:43-49 (Wyman et al. 2013 multi-lobe fit)
inline float3 wavelengthToXYZ(float lambda) {
    float x = 0.362f * G(lambda, 442.0f, 0.0624f, 0.0374f) + 1.056f * G(lambda, 599.8f, 0.0264f, 0.0323f) - 0.065f * G(lambda, 501.1f, 0.0490f, 0.0382f);
    float y = 0.821f * G(lambda, 568.8f, 0.0213f, 0.0247f) + 0.286f * G(lambda, 530.9f, 0.0613f, 0.0322f);
    float z = 1.217f * G(lambda, 437.0f, 0.0845f, 0.0278f) + 0.681f * G(lambda, 459.0f, 0.0385f, 0.0725f);
    return float3(x, y, z);
}
static const float D65_5nm[69] = {49.9755f, 52.3118f, 54.6482f, 68.7015f, 82.7549f, 87.1204f, 91.486f, 92.4589f, 93.4318f, 90.057f, 86.6823f, 95.7736f, 104.865f, 110.936f,
    117.008f, 117.41f, 117.812f, 116.336f, 114.861f, 115.392f, 115.923f, 112.367f, 108.811f, 109.082f, 109.354f, 108.578f, 107.802f, 106.296f,
    104.79f, 106.239f, 107.689f, 106.047f, 104.405f, 104.225f, 104.046f, 102.023f, 100.f, 98.1671f, 96.3342f, 96.0611f, 95.788f, 92.2368f, 88.6856f,
    89.3459f, 90.0062f, 89.8026f, 89.5991f, 88.6489f, 87.6987f, 85.4936f, 83.2886f, 83.4939f, 83.6992f, 81.863f, 80.0268f, 80.1207f, 80.2146f,
    81.2462f, 82.2778f, 80.281f, 78.2842f, 74.0027f, 69.7213f, 70.6652f, 71.6091f, 72.979f, 74.349f, 67.9765f, 61.604f};
// SpectrumUtils.slang:72-85
inline float wavelengthToD65(float lambda) {
    const float wavelengthResize = (lambda - minWavelength) * 0.2f;
    const int wavelength_t = int(std::ceil(wavelengthResize) + .5f);
    const int wavelength_b = int(std::floor(wavelengthResize) + .5f);
    const float weight = wavelengthResize - wavelength_b;
    // (index clamp is an oracle-only guard; the integrator only produces wavelengths in [380,720])
    return D65_5nm[hlsl_min(hlsl_max(wavelength_t, 0), 68)] * weight + D65_5nm[hlsl_min(hlsl_max(wavelength_b, 0), 68)] * (1.0f - weight);
}
// ColorHelpers.slang:61-70
inline float3 XYZtoRGB_Rec709(float3 c) {
    return float3(3.240969941904522f * c.x + -1.537383177570094f * c.y + -0.4986107602930032f * c.z,
                  -0.9692436362808803f * c.x + 1.875967501507721f * c.y + 0.04155505740717569f * c.z,
                  0.05563007969699373f * c.x + -0.2039769588889765f * c.y + 1.056971514242878f * c.z);
}
// SpectrumUtils.slang:99-112
inline float3 spectrumToRgb(hwss_t spectrum, hwss_t wavelengths, uint32_t samples) {
    float3 result(.0f);
    for (uint32_t i = 0; i < samples; i++) {
        float3 xyz = wavelengthToXYZ(wavelengths[i]) * wavelengthToD65(wavelengths[i]) * spectrum[i];
        result += XYZtoRGB_Rec709(xyz);
    }
    const float Y_D65 = 10.5670762f;
    return result / Y_D65;
}
static const float white_sd[11] = {1.0000f, 1.0000f, 0.9999f, 0.9993f, 0.9992f, 0.9998f, 1.0000f, 1.0000f, 1.0000f, 1.0000f, 0.0f};
static const float cyan_sd[11] = {0.9710f, 0.9426f, 1.0007f, 1.0007f, 1.0007f, 1.0007f, 0.1564f, 0.0000f, 0.0000f, 0.0000f, 0.0f};
static const float magenta_sd[11] = {1.0000f, 1.0000f, 0.968f, 0.22295f, 0.0000f, 0.0458f, 0.8369f, 1.0000f, 1.0000f, 0.9959f, 0.0f};
static const float yellow_sd[11] = {0.0001f, 0.0000f, 0.1088f, 0.6651f, 1.0000f, 1.0000f, 0.9996f, 0.9586f, 0.9685f, 0.9840f, 0.0f};
static const float red_sd[11] = {0.1012f, 0.0515f, 0.0000f, 0.0000f, 0.0000f, 0.0000f, 0.8325f, 1.0149f, 1.0149f, 1.014f, 0.0f};
static const float green_sd[11] = {0.0000f, 0.0000f, 0.0273f, 0.7937f, 1.0000f, 0.9418f, 0.1719f, 0.0000f, 0.0000f, 0.0025f, 0.0f};
static const float blue_sd[11] = {1.0000f, 1.0000f, 0.8916f, 0.3323f, 0.0000f, 0.0000f, 0.0003f, 0.0369f, 0.0483f, 0.0496f, 0.0f};
// SpectrumUtils.slang:114-176 (Smits-style 10-bin basis; bin 10 — exactly 720 nm — is zero)
inline float rgbToSpectrum(float3 rgb, float wavelength) {
    float sd = .0f;
    const int b = (int)(clampf((wavelength - minWavelength) / (maxWavelength - minWavelength), .0f, 1.f) * 10.f);
    const float white_s = white_sd[b], cyan_s = cyan_sd[b], magenta_s = magenta_sd[b], yellow_s = yellow_sd[b];
    const float red_s = red_sd[b], green_s = green_sd[b], blue_s = blue_sd[b];
    const float r = rgb.x, g = rgb.y, bl = rgb.z;
    if (r <= g && r <= bl) {
        sd += white_s * r;
        if (g <= bl) { sd += cyan_s * (g - r); sd += blue_s * (bl - g); }
        else { sd += cyan_s * (bl - r); sd += green_s * (g - bl); }
    } else if (g <= r && g <= bl) {
        sd += white_s * g;
        if (r <= bl) { sd += magenta_s * (r - g); sd += blue_s * (bl - r); }
        else { sd += magenta_s * (bl - g); sd += red_s * (r - bl); }
    } else {
        sd += white_s * bl;
        if (r <= g) { sd += yellow_s * (r - bl); sd += green_s * (g - r); }
        else { sd += yellow_s * (g - bl); sd += red_s * (r - g); }
    }
    return sd;
}
// SpectrumUtils.slang:177-239 — the vector overload applies the scalar rule per lane.
inline hwss_t rgbToSpectrum(float3 rgb, hwss_t wavelengths) {
    hwss_t r;
    for (int i = 0; i < 4; ++i) r[i] = rgbToSpectrum(rgb, wavelengths[i]);
    return r;
}
}  // namespace SpectrumUtils

namespace SpectrumSampler {
// SpectrumSampler.slang:38-41
inline float sampleOne(float rnd) { return SpectrumUtils::minWavelength + rnd * (SpectrumUtils::maxWavelength - SpectrumUtils::minWavelength); }
// SpectrumSampler.slang:44-54
inline hwss_t hwssWavelengths(float hero) {
    const float span = SpectrumUtils::maxWavelength - SpectrumUtils::minWavelength;
    hwss_t h;
    const int n = hwss_samples();
    for (int i = 0; i < 4; ++i) h[i] = float(i);
    h = hwss_t(hero) + h * (span / float(n));
    for (int i = 0; i < 4; ++i) h[i] -= h[i] > SpectrumUtils::maxWavelength ? span : .0f;
    return h;
}
}  // namespace SpectrumSampler

// ---------------------------------------------------------------------------------------------
// Spectral profiles on the device side: Scene.slang:251-327
// ---------------------------------------------------------------------------------------------
struct ProfileTable {
    const PltptSpectralProfile* profiles = nullptr;
    uint32_t count = 0;

    float evalSpectrum(uint32_t id, float lambda) const {
        if (id == PLTPT_NO_SPECTRAL_PROFILE) return .0f;
        const PltptSpectralProfile& sp = profiles[id];
        const uint32_t bins = sp.bins;
        const float minl = sp.minWavelength, maxl = sp.maxWavelength;
        const float x = (lambda - minl) / (maxl - minl) * (bins - 1.0f);
        const float i = std::floor(x);
        const float w = x - i;
        if (lambda < minl || lambda > maxl) return .0f;
        const int i0 = hlsl_max(0, hlsl_min(int(bins) - 1, int(i)));
        const int i1 = hlsl_max(0, hlsl_min(int(bins) - 1, int(i + 1)));
        return sp.intensity * lerpf(sp.pdf[i0], sp.pdf[i1], w);
    }
    hwss_t evalSpectrum(uint32_t id, hwss_t lambdas) const {
        hwss_t r;
        for (int j = 0; j < 4; ++j) r[j] = evalSpectrum(id, lambdas[j]);
        return r;
    }
    // Scene.slang:296-327
    hwss_t sampleSpectrum(uint32_t id, float heroWavelength, SampleGenerator& sg, bool importanceSample, hwss_t& sampledWavelengths) const {
        // NOTE (reference quirk): for id == no_spectral_profile the out parameter is left unwritten.
        if (id == PLTPT_NO_SPECTRAL_PROFILE) return hwss_t(.0f);
        hwss_t ret(.0f);
        if (importanceSample) {
            const PltptSpectralProfile& sp = profiles[id];
            sampledWavelengths.x = heroWavelength;
            ret.x = evalSpectrum(id, heroWavelength);
            const float3 x = float(sp.bins) * sampleNext3D(sg);
            for (int j = 1; j < hwss_samples(); ++j) {
                const uint32_t idx = uint32_t(x[j - 1]);
                const float cdf = lerpf(idx < sp.bins ? sp.icdf[idx] : 1.f, idx + 1 < sp.bins ? sp.icdf[idx + 1] : 1.f, fracf(x[j - 1]));
                sampledWavelengths[j] = lerpf(sp.minWavelength, sp.maxWavelength, cdf);
                ret[j] = sp.intensity / (float)sp.bins;
            }
        } else {
            sampledWavelengths = SpectrumSampler::hwssWavelengths(heroWavelength);
            ret = evalSpectrum(id, sampledWavelengths);
        }
        return ret;
    }
};

// ---------------------------------------------------------------------------------------------
// Host-side table ingestion: Utils/Color/Spectrum.{h,cpp}, SampledSpectrum.h, SceneBuilder.cpp:911-960
// ---------------------------------------------------------------------------------------------
// Falcor's HOST lerp, Utils/Math/Common.h:54-57: (1 - t) * a + t * b -- not the shader form (lerpf): found by compiling the reference's own C++
// (oracle/ref_spectrum.cpp); the profile tables are held to it byte for byte (tests/golden/ref_spectral_profiles.npz)
inline float hostLerp(float a, float b, float t) { return (1.f - t) * a + t * b; }
struct PiecewiseLinearSpectrum {
    std::vector<float> wavelengths, values;
    // Spectrum.h:88-106
    float eval(float wavelength) const {
        if (wavelengths.empty() || wavelength < wavelengths.front() || wavelength > wavelengths.back()) return 0.f;
        auto it = std::lower_bound(wavelengths.begin(), wavelengths.end(), wavelength);
        if (it == wavelengths.begin()) return values.front();
        size_t index = size_t(it - wavelengths.begin()) - 1;
        float t = (wavelength - wavelengths[index]) / (wavelengths[index + 1] - wavelengths[index]);
        return hostLerp(values[index], values[index + 1], t);
    }
    float leastBinSize() const {
        float size = wavelengths.back() - wavelengths.front();
        for (size_t i = 1; i < wavelengths.size(); ++i) size = hlsl_min(size, wavelengths[i] - wavelengths[i - 1]);
        return size;
    }
};
struct SampledSpectrum {
    float lambdaStart = 380.f, lambdaEnd = 720.f;
    std::vector<float> samples;
    SampledSpectrum() = default;
    SampledSpectrum(float a, float b, std::vector<float> v) : lambdaStart(a), lambdaEnd(b), samples(std::move(v)) {}
    // SampledSpectrum.h:57-68
    explicit SampledSpectrum(const PiecewiseLinearSpectrum& pls, float lambdaBinMinSize = 5.f) : lambdaStart(pls.wavelengths.front()), lambdaEnd(pls.wavelengths.back()) {
        const float binSize = hlsl_max(lambdaBinMinSize, pls.leastBinSize() / 2.f);
        samples.resize(std::max<size_t>(1, size_t((lambdaEnd - lambdaStart) / float(binSize) + .5f)));
        for (size_t idx = 0; idx < samples.size(); ++idx) samples[idx] = pls.eval(lambdaStart + (lambdaEnd - lambdaStart) / float(samples.size() - 1) * idx);
    }
    // SampledSpectrum.h:303-317
    float eval(float lambda) const {
        if (lambda < lambdaStart || lambda > lambdaEnd) return 0.f;
        float x = ((lambda - lambdaStart) / (lambdaEnd - lambdaStart)) * (samples.size() - 1.0f);
        size_t i = (size_t)std::floor(x);
        if (i + 1 >= samples.size()) return samples[samples.size() - 1];
        float w = x - (float)i;
        return hostLerp(samples[i], samples[i + 1], w);
    }
};
// SceneBuilder.cpp:911-960. sp.rgb is left to the caller (needs the CIE 1931 1-nm tables, SpectrumUtils.h:143-151).
inline PltptSpectralProfile genSpectralProfile(const SampledSpectrum& ss, bool computeICdf) {
    PltptSpectralProfile sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.intensity = 1.f;
    sp.minWavelength = hlsl_max(ss.lambdaStart, SpectrumUtils::minWavelength);
    sp.maxWavelength = hlsl_min(ss.lambdaEnd, SpectrumUtils::maxWavelength);
    sp.bins = std::min<uint32_t>(128u, uint32_t(ss.samples.size() * (sp.maxWavelength - sp.minWavelength) / (ss.lambdaEnd - ss.lambdaStart) + .5f));
    if (sp.bins == 0) return sp;
    const float r = sp.bins > 1 ? 1.f / float(sp.bins - 1) : 1.f;
    float sum = .0f;
    for (size_t i = 0; i < sp.bins; ++i) {
        sp.pdf[i] = ss.eval(hostLerp(sp.minWavelength, sp.maxWavelength, i * r));
        sum += sp.pdf[i];
    }
    if (sum < FLT_EPS || !computeICdf) return sp;
    for (size_t i = 0; i < sp.bins; ++i) sp.pdf[i] /= sum;
    sp.intensity = sum;
    float cdfa = .0f, cdfb = sp.pdf[0];
    for (size_t pdfi = 0, icdfi = 0; icdfi < sp.bins; ++icdfi) {
        const float cdf = icdfi * r;
        while (cdf >= cdfb && pdfi + 1 < sp.bins) { cdfa = cdfb; ++pdfi; cdfb += sp.pdf[pdfi]; }
        const float l = hlsl_min(1.f, (cdf - cdfa) / (cdfb - cdfa));
        sp.icdf[icdfi] = hostLerp((float)std::max<uint32_t>(1, uint32_t(pdfi)) - 1.f, (float)pdfi, l) * r;
    }
    return sp;
}

}  // namespace orc
