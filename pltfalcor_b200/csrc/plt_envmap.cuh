// plt_envmap.cuh — environment-map lighting of the PLTPT path: texel access, bilinear evaluation, hierarchical importance
// sampling and the pdf of a direction.  Follows
//   Source/Falcor/Scene/Lights/EnvMap.slang:36-134 (eval / toWorld / toLocal), EnvMap.cpp:121-126 (linear, wrap u / clamp v),
//   Source/Falcor/Rendering/Lights/EnvMapSampler.slang:53-165 (sample, evalPdf),
//   Source/Falcor/Rendering/Lights/EnvMapSamplerSetup.cs.slang:47-75 + EnvMapSampler.cpp:82-121 (importance map, 512^2 @ 64 spp).
// The reference samples through the texture unit; here the bilinear filter is spelled out in fp32 (same operation order as
// oracle/plt_envmap.h) and the importance-map mip chain lives in one linear buffer (mip m starts at envMipOffset(m)).
#pragma once
#include "plt_math.cuh"

namespace plt {

constexpr uint32_t kEnvImportanceDim = 512, kEnvImportanceSamples = 8, kEnvImportanceBaseMip = 9;  // 512 = 2^9: mips 0..9
__host__ __device__ constexpr uint32_t envMipOffset(uint32_t mip) {  // sum_{k<mip} (512 >> k)^2
    uint32_t o = 0;
    for (uint32_t k = 0; k < mip; ++k) o += (kEnvImportanceDim >> k) * (kEnvImportanceDim >> k);
    return o;
}
constexpr uint32_t kEnvImportanceFloats = envMipOffset(kEnvImportanceBaseMip + 1);

struct DevEnvMap {
    const float4* texels;      // lat-long RGBA32F, row-major (null = no env map)
    const float* importance;   // mip chain of the octahedral importance map
    uint32_t width, height;
    float toWorld[9], toLocal[9];  // (float3x3) of EnvMapData::transform / invTransform, row-major
    float3 intensity;              // intensity * tint
};

// MathHelpers.slang:106-113
PLT_D float2 world_to_latlong_map(float3 dir) {
    const float3 p = normalize(dir);
    return mk2(atan2f(p.x, -p.z) * k1_2Pi + 0.5f, acosf(p.y) * k1_Pi);
}
PLT_D float3 oct_to_ndir_equal_area_unorm(float2 p) { return oct_to_ndir_equal_area_snorm(mk2(p.x * 2.f - 1.f, p.y * 2.f - 1.f)); }
PLT_D float2 ndir_to_oct_equal_area_unorm(float3 n) {
    const float2 s = ndir_to_oct_equal_area_snorm(n);
    return mk2(s.x * 0.5f + 0.5f, s.y * 0.5f + 0.5f);
}
PLT_D float luminance709(float3 c) { return (c.x * 0.2126f + c.y * 0.7152f) + c.z * 0.0722f; }
PLT_D float3 mul33(const float* M, float3 d) {
    return mk3((M[0] * d.x + M[1] * d.y) + M[2] * d.z, (M[3] * d.x + M[4] * d.y) + M[5] * d.z, (M[6] * d.x + M[7] * d.y) + M[8] * d.z);
}
PLT_D float3 envTexel(const DevEnvMap& e, int x, int y) {
    const float4 t = __ldg(e.texels + size_t(y) * e.width + size_t(x));
    return mk3(t.x, t.y, t.z);
}
// SampleLevel(envSampler, uv, 0)
PLT_D float3 envSampleBilinear(const DevEnvMap& e, float2 uv) {
    const float x = uv.x * float(e.width) - 0.5f, y = uv.y * float(e.height) - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float wx = x - fx0, wy = y - fy0;
    const int W = int(e.width), H = int(e.height);
    int x0 = int(fx0) % W; if (x0 < 0) x0 += W;
    int x1 = (int(fx0) + 1) % W; if (x1 < 0) x1 += W;
    const int y0 = min(max(int(fy0), 0), H - 1), y1 = min(max(int(fy0) + 1, 0), H - 1);
    const float3 t00 = envTexel(e, x0, y0), t10 = envTexel(e, x1, y0), t01 = envTexel(e, x0, y1), t11 = envTexel(e, x1, y1);
    const float3 top = t00 + (t10 - t00) * wx;
    const float3 bot = t01 + (t11 - t01) * wx;
    return top + (bot - top) * wy;
}
// EnvMap.slang:57-74
PLT_D float3 envEval(const DevEnvMap& e, float3 dirW) { return envSampleBilinear(e, world_to_latlong_map(mul33(e.toLocal, dirW))) * e.intensity; }
PLT_D float envImportance(const DevEnvMap& e, uint32_t x, uint32_t y, uint32_t mip) { return __ldg(e.importance + envMipOffset(mip) + size_t(y) * (kEnvImportanceDim >> mip) + x); }

struct EnvSample { float3 dir; float pdf; float3 Le; };
// EnvMapSampler.slang:71-146
PLT_D void envSample(const DevEnvMap& e, float2 rnd, EnvSample& r) {
    float2 p = rnd;
    uint32_t posx = 0, posy = 0, off = envMipOffset(kEnvImportanceBaseMip);
#pragma unroll 1
    for (int mip = int(kEnvImportanceBaseMip) - 1; mip >= 0; --mip) {
        posx *= 2; posy *= 2;
        const uint32_t n = kEnvImportanceDim >> mip;
        off -= n * n;  // offset of `mip`
        const float* row = e.importance + off + size_t(posy) * n + posx;
        const float2 a = __ldg(reinterpret_cast<const float2*>(row)), b = __ldg(reinterpret_cast<const float2*>(row + n));  // posx even, n even: 8 B aligned
        const float w0 = a.x, w1 = a.y, w2 = b.x, w3 = b.y;
        const float q0 = w0 + w2, q1 = w1 + w3;
        uint32_t offx, offy;
        const float d = q0 / (q0 + q1);
        if (p.x < d) { offx = 0; p.x = p.x / d; }
        else { offx = 1; p.x = (p.x - d) / (1.f - d); }
        const float ee = offx == 0 ? (w0 / q0) : (w1 / q1);
        if (p.y < ee) { offy = 0; p.y = p.y / ee; }
        else { offy = 1; p.y = (p.y - ee) / (1.f - ee); }
        posx += offx; posy += offy;
    }
    const float invDim = 1.f / float(kEnvImportanceDim);
    const float2 uv = mk2((float(posx) + p.x) * invDim, (float(posy) + p.y) * invDim);
    const float3 dir = oct_to_ndir_equal_area_unorm(uv);
    const float avg_w = envImportance(e, 0, 0, kEnvImportanceBaseMip);
    const float pdf = envImportance(e, posx, posy, 0) / avg_w;
    r.dir = mul33(e.toWorld, dir);
    r.pdf = pdf * 0.079577471545947667884f;  // M_1_4PI
    r.Le = envEval(e, r.dir);
}
// EnvMapSampler.slang:156-162 (point sampling, clamp)
PLT_D float envEvalPdf(const DevEnvMap& e, float3 dirW) {
    const float2 uv = ndir_to_oct_equal_area_unorm(mul33(e.toLocal, dirW));
    const float avg_w = envImportance(e, 0, 0, kEnvImportanceBaseMip);
    const int dim = int(kEnvImportanceDim);
    const int x = min(max(int(floorf(uv.x * float(dim))), 0), dim - 1), y = min(max(int(floorf(uv.y * float(dim))), 0), dim - 1);
    return (envImportance(e, uint32_t(x), uint32_t(y), 0) / avg_w) * (1.f / k4Pi);
}

}  // namespace plt
