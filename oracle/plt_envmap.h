// ORACLE — TEST INFRASTRUCTURE ONLY (see plt_math.h header).  Pinned to the reference's shader text: EnvMap.slang eval / toLocal /
// toWorld, EnvMapSampler.slang sample / evalPdf (oracle/ref_compare.cpp).  The importance-map construction (a compute shader) and the bilinear filter
// (hardware) are restated here and checked on their properties (tests/test_envmap.py).
// plt_envmap.h — environment map access and hierarchical importance sampling.
//   Source/Falcor/Scene/Lights/EnvMap.slang:36-134, EnvMapData.slang:36-43, EnvMap.cpp:121-126 (sampler: linear, wrap u / clamp v)
//   Source/Falcor/Rendering/Lights/EnvMapSampler.slang:53-165, EnvMapSamplerSetup.cs.slang:47-75, EnvMapSampler.cpp:82-121
// Texture filtering is the D3D/Vulkan fixed-function unit in the reference (implementation-defined sub-texel precision); here it is
// restated as exact fp32 bilinear interpolation, which the CUDA path follows operation by operation.
#pragma once
#include <vector>
#include "plt_math.h"
#include "../include/pltpt.h"

namespace orc {

// MathHelpers.slang:106-113
inline float2 world_to_latlong_map(float3 dir) {
    const float3 p = normalize(dir);
    return float2(std::atan2(p.x, -p.z) * k1_2Pi + 0.5f, std::acos(p.y) * k1_Pi);
}
inline float3 oct_to_ndir_equal_area_unorm(float2 p) { return oct_to_ndir_equal_area_snorm(p * 2.f - 1.f); }   // MathHelpers.slang:241-243
inline float2 ndir_to_oct_equal_area_unorm(float3 n) { return ndir_to_oct_equal_area_snorm(n) * 0.5f + 0.5f; } // MathHelpers.slang:214-217
inline float luminance709(float3 rgb) { return dot(rgb, float3(0.2126f, 0.7152f, 0.0722f)); }                  // ColorHelpers.slang:39-42

struct EnvMap {
    const float* texels = nullptr;  // RGBA32F, row-major
    uint32_t width = 0, height = 0;
    PltptEnvMapData data{};
    // importance map: mip 0 = dim x dim, ..., mip baseMip = 1 x 1
    static constexpr uint32_t kDimension = 512, kSamplesX = 8, kSamplesY = 8;  // EnvMapSampler.cpp:40-42 (512 x 512 @ 64 spp)
    std::vector<std::vector<float>> mips;
    uint32_t baseMip = 0;

    bool valid() const { return texels != nullptr; }

    float3 texel(int x, int y) const {
        const float* t = texels + 4 * (size_t(y) * width + size_t(x));
        return float3(t[0], t[1], t[2]);
    }
    // SampleLevel(envSampler, uv, 0): bilinear, address U = Wrap, V = Clamp
    float3 sampleBilinear(float2 uv) const {
        const float x = uv.x * float(width) - 0.5f, y = uv.y * float(height) - 0.5f;
        const float fx0 = std::floor(x), fy0 = std::floor(y);
        const float wx = x - fx0, wy = y - fy0;
        const int W = int(width), H = int(height);
        int x0 = int(fx0) % W; if (x0 < 0) x0 += W;
        int x1 = (int(fx0) + 1) % W; if (x1 < 0) x1 += W;
        const int y0 = hlsl_min(hlsl_max(int(fy0), 0), H - 1), y1 = hlsl_min(hlsl_max(int(fy0) + 1, 0), H - 1);
        const float3 top = texel(x0, y0) + (texel(x1, y0) - texel(x0, y0)) * wx;
        const float3 bot = texel(x0, y1) + (texel(x1, y1) - texel(x0, y1)) * wx;
        return top + (bot - top) * wy;
    }
    static float3 mul3x4(const float* M, float3 d) {  // mul((float3x3)M, d)
        return float3((M[0] * d.x + M[1] * d.y) + M[2] * d.z, (M[4] * d.x + M[5] * d.y) + M[6] * d.z, (M[8] * d.x + M[9] * d.y) + M[10] * d.z);
    }
    float3 toWorld(float3 d) const { return mul3x4(data.transform, d); }
    float3 toLocal(float3 d) const { return mul3x4(data.invTransform, d); }
    float3 intensity() const { return float3(data.tint[0], data.tint[1], data.tint[2]) * data.intensity; }
    // EnvMap.slang:57-74
    float3 eval(float3 dirW) const { return sampleBilinear(world_to_latlong_map(toLocal(dirW))) * intensity(); }

    // EnvMapSamplerSetup.cs.slang:47-75 + Texture::generateMips (2x2 box filter)
    void init(const PltptSceneDesc& d) {
        texels = d.envMapTexels; width = d.envMapWidth; height = d.envMapHeight; data = d.envMap;
        mips.clear();
        if (!texels) return;
        const uint32_t dim = kDimension;
        const float invSamples = 1.f / float(kSamplesX * kSamplesY);
        std::vector<float> base(size_t(dim) * dim);
        for (uint32_t py = 0; py < dim; ++py)
            for (uint32_t px = 0; px < dim; ++px) {
                float L = 0.f;
                for (uint32_t y = 0; y < kSamplesY; ++y)
                    for (uint32_t x = 0; x < kSamplesX; ++x) {
                        const float2 p((float(px * kSamplesX + x) + 0.5f) / float(dim * kSamplesX), (float(py * kSamplesY + y) + 0.5f) / float(dim * kSamplesY));
                        const float3 dir = oct_to_ndir_equal_area_unorm(p);
                        L += luminance709(sampleBilinear(world_to_latlong_map(dir)));
                    }
                base[size_t(py) * dim + px] = L * invSamples;
            }
        mips.push_back(std::move(base));
        for (uint32_t n = dim / 2; n >= 1; n /= 2) {
            const std::vector<float>& src = mips.back();
            std::vector<float> dst(size_t(n) * n);
            for (uint32_t y = 0; y < n; ++y)
                for (uint32_t x = 0; x < n; ++x) {
                    const size_t o = size_t(2 * y) * (2 * n) + 2 * x;
                    dst[size_t(y) * n + x] = ((src[o] + src[o + 1]) + (src[o + 2 * n] + src[o + 2 * n + 1])) * 0.25f;
                }
            mips.push_back(std::move(dst));
            if (n == 1) break;
        }
        baseMip = uint32_t(mips.size()) - 1;
    }
    float load(uint32_t x, uint32_t y, uint32_t mip) const { return mips[mip][size_t(y) * (kDimension >> mip) + x]; }

    struct Sample { float3 dir; float pdf = 0; float3 Le; };
    // EnvMapSampler.slang:71-146 (Le returned as RGB; the caller upsamples to the HWSS wavelengths)
    bool sample(float2 rnd, Sample& result) const {
        float2 p = rnd;
        uint32_t posx = 0, posy = 0;
        for (int mip = int(baseMip) - 1; mip >= 0; --mip) {
            posx *= 2; posy *= 2;
            const float w0 = load(posx, posy, mip), w1 = load(posx + 1, posy, mip), w2 = load(posx, posy + 1, mip), w3 = load(posx + 1, posy + 1, mip);
            const float q0 = w0 + w2, q1 = w1 + w3;
            uint32_t offx, offy;
            const float d = q0 / (q0 + q1);
            if (p.x < d) { offx = 0; p.x = p.x / d; }
            else { offx = 1; p.x = (p.x - d) / (1.f - d); }
            const float e = offx == 0 ? (w0 / q0) : (w1 / q1);
            if (p.y < e) { offy = 0; p.y = p.y / e; }
            else { offy = 1; p.y = (p.y - e) / (1.f - e); }
            posx += offx; posy += offy;
        }
        const float invDim = 1.f / float(kDimension);
        const float2 uv((float(posx) + p.x) * invDim, (float(posy) + p.y) * invDim);
        const float3 dir = oct_to_ndir_equal_area_unorm(uv);
        const float avg_w = load(0, 0, baseMip);
        const float pdf = load(posx, posy, 0) / avg_w;
        result.dir = toWorld(dir);
        result.pdf = pdf * 0.079577471545947667884f;  // M_1_4PI
        result.Le = eval(result.dir);
        return true;
    }
    // EnvMapSampler.slang:156-162 (importanceSampler: point, clamp)
    float evalPdf(float3 dirW) const {
        const float2 uv = ndir_to_oct_equal_area_unorm(toLocal(dirW));
        const float avg_w = load(0, 0, baseMip);
        const int dim = int(kDimension);
        const int x = hlsl_min(hlsl_max(int(std::floor(uv.x * float(dim))), 0), dim - 1), y = hlsl_min(hlsl_max(int(std::floor(uv.y * float(dim))), 0), dim - 1);
        const float pdf = load(uint32_t(x), uint32_t(y), 0) / avg_w;
        return pdf * (1.f / k4Pi);
    }
};

}  // namespace orc
