// plt_math.cuh — device vector math, PRNG, spectral helpers for the sm_100a PLTPT path.
// Semantics follow the reference Slang modules cited per function; layout/structure is our own.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/pltpt.h"

#define PLT_HD __host__ __device__ __forceinline__
#define PLT_D __device__ __forceinline__

// Loops over the (at most four) wavelength samples of a generalized ray.  PLT_LAMBDA_UNROLL = 1 spells them as four unrolled iterations
// guarded by the runtime count: lane indices become compile-time constants (no select chains for float4 components) and the four
// wavelengths form independent instruction streams the scheduler can interleave; 0 keeps the plain runtime-bounded loop.
#ifndef PLT_LAMBDA_UNROLL
#define PLT_LAMBDA_UNROLL 0
#endif
#if PLT_LAMBDA_UNROLL
#define PLT_FOR_LAMBDA(j, n) _Pragma("unroll") for (int j = 0; j < 4; ++j) if (j < int(n))
#else
#define PLT_FOR_LAMBDA(j, n) for (int j = 0; j < int(n); ++j)
#endif

// PLT_FAST_MATH = 1 (set per translation unit, pltfalcor_b200/build.py): the replay stage (stage_unwind.cu) evaluates BSDFs on paths
// that are already fixed -- no ray is generated and no branch of the path depends on its arithmetic -- so it only has to meet the
// 1e-4 relative BSDF tolerance, not the operation-by-operation rounding the sample stage keeps for identical paths.  There the
// IEEE division / square-root sequences (8-12 instructions each plus a slow-path call) become MUFU approximations (<= 2 ulp),
// normalisation is one rsqrt, and rotations are built from the sine and cosine directly instead of atan2 -> sincos.
#ifndef PLT_FAST_MATH
#define PLT_FAST_MATH 0
#endif
#if PLT_FAST_MATH && defined(__CUDA_ARCH__)
#define PLT_FAST_DEVICE 1
#else
#define PLT_FAST_DEVICE 0
#endif

namespace plt {

constexpr float kPi = 3.14159265358979323846f;
constexpr float k2Pi = 6.28318530717958647693f;
constexpr float k4Pi = 12.5663706143591729539f;
constexpr float kPi_2 = 1.57079632679489661923f;
constexpr float kPi_4 = 0.785398163397448309616f;
constexpr float k1_Pi = 0.318309886183790671538f;
constexpr float k2_Pi = 0.636619772367581343076f;
constexpr float k1_2Pi = 0.159154943091895335769f;
constexpr float kFltEps = 1.192092896e-07F;
constexpr float kFltMin = 1.175494351e-38F;
constexpr float kFltMax = 3.402823466e+38F;

// ---- float2/3/4 operators -----------------------------------------------------------------
PLT_HD float2 mk2(float x, float y) { return make_float2(x, y); }
PLT_HD float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
PLT_HD float3 mk3(float s) { return make_float3(s, s, s); }
PLT_HD float4 mk4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
PLT_HD float4 mk4(float s) { return make_float4(s, s, s, s); }

PLT_HD float2 operator+(float2 a, float2 b) { return mk2(a.x + b.x, a.y + b.y); }
PLT_HD float2 operator-(float2 a, float2 b) { return mk2(a.x - b.x, a.y - b.y); }
PLT_HD float2 operator*(float2 a, float s) { return mk2(a.x * s, a.y * s); }
PLT_HD float2 operator*(float s, float2 a) { return mk2(a.x * s, a.y * s); }
PLT_HD float2 operator/(float2 a, float s) { return mk2(a.x / s, a.y / s); }
PLT_HD float2 operator-(float2 a) { return mk2(-a.x, -a.y); }
PLT_HD float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }

PLT_HD float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PLT_HD float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PLT_HD float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PLT_HD float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PLT_HD float3 operator*(float s, float3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
PLT_HD float3 operator/(float3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
PLT_HD float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
PLT_HD float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
PLT_HD float3& operator-=(float3& a, float3 b) { a = a - b; return a; }
PLT_HD float3& operator*=(float3& a, float s) { a = a * s; return a; }
PLT_HD float3& operator/=(float3& a, float s) { a = a / s; return a; }
PLT_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PLT_HD float3 cross(float3 a, float3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PLT_HD float length(float3 a) { return sqrtf(dot(a, a)); }
#if PLT_FAST_DEVICE
PLT_HD float3 normalize(float3 a) { return a * rsqrtf(dot(a, a)); }
#else
// The oracle's definition of the HLSL intrinsic (oracle/plt_math.h): one correctly rounded square root, one correctly rounded reciprocal,
// three products.  (Rounds 1-2 used v / length(v): three IEEE divisions per call were a sixth of k_shade's instructions -- 1.22 -> 1.04 ms,
// k_connect_a 1.29 -> 1.19 ms per Cornell frame, profiles/r2m_variants.log.)
PLT_HD float3 normalize(float3 a) { return a * (1.f / length(a)); }
#endif
PLT_HD float3 vabs(float3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
PLT_HD float3 vmax(float3 a, float3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PLT_HD float3 vmin(float3 a, float3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PLT_HD float comp(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

PLT_HD float4 operator+(float4 a, float4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
PLT_HD float4 operator-(float4 a, float4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
PLT_HD float4 operator*(float4 a, float4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
PLT_HD float4 operator/(float4 a, float4 b) { return mk4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
PLT_HD float4 operator*(float4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
PLT_HD float4 operator*(float s, float4 a) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
PLT_HD float4 operator/(float4 a, float s) { return mk4(a.x / s, a.y / s, a.z / s, a.w / s); }
PLT_HD float4 operator/(float s, float4 a) { return mk4(s / a.x, s / a.y, s / a.z, s / a.w); }
PLT_HD float4& operator*=(float4& a, float4 b) { a = a * b; return a; }
PLT_HD float4& operator*=(float4& a, float s) { a = a * s; return a; }
PLT_HD float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
PLT_HD float lane(const float4& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : (i == 2 ? a.z : a.w)); }
PLT_HD void setLane(float4& a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else if (i == 2) a.z = v; else a.w = v; }

PLT_HD float sqr(float v) { return v * v; }
PLT_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
PLT_HD float lerpf(float a, float b, float t) { return a + (b - a) * t; }
PLT_HD float signf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
#ifdef __CUDA_ARCH__
PLT_D uint32_t asuint(float f) { return __float_as_uint(f); }
PLT_D float asfloat(uint32_t u) { return __uint_as_float(u); }
#else
inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif

// fp16 -> fp32 (f16tof32): exact for every finite value, subnormals included.  On the device it is the hardware conversion (one
// instruction; the bit manipulation below was a tenth of k_shade's instructions: three vertex normals + texture coordinates per hit,
// profiles/r2t_shade_lines.txt); the host keeps the portable form.
PLT_HD float h2f(uint32_t h) {
#ifdef __CUDA_ARCH__
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"((unsigned short)(h & 0xffffu)));
    return f;
#endif
    h &= 0xffffu;
    const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ffu, out;
    if (e == 0) {
        if (man == 0) out = sign;
        else {
            int sh = -1;
            do { ++sh; man <<= 1; } while ((man & 0x400u) == 0);
            out = sign | uint32_t(127 - 15 - sh) << 23 | (man & 0x3ffu) << 13;
        }
    } else if (e == 31) out = sign | 0x7f800000u | man << 13;
    else out = sign | (e + 112) << 23 | man << 13;
    return asfloat(out);
}

// 2x2 matrix, row-major
struct M22 { float a, b, c, d; };
PLT_HD M22 mul(M22 A, M22 B) { return {A.a * B.a + A.b * B.c, A.a * B.b + A.b * B.d, A.c * B.a + A.d * B.c, A.c * B.b + A.d * B.d}; }
PLT_HD M22 transpose(M22 A) { return {A.a, A.c, A.b, A.d}; }
PLT_HD float2 mul(M22 A, float2 v) { return mk2(A.a * v.x + A.b * v.y, A.c * v.x + A.d * v.y); }
PLT_HD float det(M22 A) { return A.a * A.d - A.b * A.c; }
PLT_HD M22 operator*(float s, M22 A) { return {s * A.a, s * A.b, s * A.c, s * A.d}; }
// MathHelpers.slang:484-496
PLT_HD bool tryInvert(M22 M, M22& inv) {
    const float dt = det(M);
    if (fabsf(dt) < 4e-6f) return false;
    const float id = 1.0f / dt;
    inv = {M.d * id, -M.b * id, -M.c * id, M.a * id};
    return true;
}

// MathHelpers.slang:609-621
PLT_HD float3 perp_stark(float3 u) {
    const float3 a = vabs(u);
    const uint32_t uyx = (a.x - a.y) < 0 ? 1 : 0, uzx = (a.x - a.z) < 0 ? 1 : 0, uzy = (a.y - a.z) < 0 ? 1 : 0;
    const uint32_t xm = uyx & uzx, ym = (1 ^ xm) & uzy, zm = 1 ^ (xm | ym);
    return normalize(cross(u, mk3(float(xm), float(ym), float(zm))));
}

// complex helpers — MathHelpers.slang:822-837
PLT_HD float cnorm2(float2 z) { return dot(z, z); }
PLT_HD float2 cmul(float2 a, float2 b) { return mk2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PLT_HD float2 cconj(float2 z) { return mk2(z.x, -z.y); }
PLT_HD float2 cdiv(float2 a, float2 b) { return cmul(a, cconj(b)) / cnorm2(b); }
PLT_HD float2 csqrt(float2 z) {
    const float r = sqrtf(cnorm2(z));
    return mk2(sqrtf(fmaxf(.0f, r + z.x) / 2.f), signf(z.y) * sqrtf(fmaxf(.0f, r - z.x) / 2.f));
}

// Octahedral maps — MathHelpers.slang:149-240
PLT_HD float2 oct_wrap(float2 v) { return mk2((1.f - fabsf(v.y)) * (v.x >= 0.f ? 1.f : -1.f), (1.f - fabsf(v.x)) * (v.y >= 0.f ? 1.f : -1.f)); }
PLT_HD float3 oct_to_ndir_snorm(float2 p) {
    float3 n = mk3(p.x, p.y, 1.f - fabsf(p.x) - fabsf(p.y));
    if (n.z < 0.f) { const float2 w = oct_wrap(mk2(n.x, n.y)); n.x = w.x; n.y = w.y; }
    return normalize(n);
}
PLT_HD float unpackSnorm16(uint32_t v) { return fmaxf(float(int32_t(int16_t(v & 0xffff))) / 32767.f, -1.f); }
PLT_HD float3 decodeNormal2x16(uint32_t p) { return oct_to_ndir_snorm(mk2(unpackSnorm16(p & 0xffff), unpackSnorm16(p >> 16))); }
PLT_HD float2 ndir_to_oct_equal_area_snorm(float3 n) {
    const float r = sqrtf(1.f - fabsf(n.z));
    const float phi = atan2f(fabsf(n.y), fabsf(n.x));
    float2 p;
    p.y = r * phi * k2_Pi;
    p.x = r - p.y;
    if (n.z < 0.f) p = mk2(1.f - p.y, 1.f - p.x);
    return mk2(p.x * signf(n.x), p.y * signf(n.y));
}
PLT_HD float3 oct_to_ndir_equal_area_snorm(float2 p) {
    const float d = 1.f - (fabsf(p.x) + fabsf(p.y));
    const float r = 1.f - fabsf(d);
    const float phi = (r > 0.f) ? ((fabsf(p.y) - fabsf(p.x)) / r + 1.f) * kPi_4 : 0.f;
    const float f = r * sqrtf(2.f - r * r);
#if PLT_FAST_DEVICE
    return mk3(f * signf(p.x) * __cosf(phi), f * signf(p.y) * __sinf(phi), signf(d) * (1.f - r * r));  // phi in [0, pi/2]: abs error < 1e-6
#else
    return mk3(f * signf(p.x) * cosf(phi), f * signf(p.y) * sinf(phi), signf(d) * (1.f - r * r));
#endif
}

// GeometryHelpers.slang:47-61
PLT_HD float3 computeRayOrigin(float3 pos, float3 normal) {
    const float origin = 1.f / 16.f, fScale = 3.f / 65536.f, iScale = 3 * 256.f;
    float r[3];
    const float p[3] = {pos.x, pos.y, pos.z}, n[3] = {normal.x, normal.y, normal.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int32_t iOff = int32_t(n[i] * iScale);
        const float iPos = asfloat(uint32_t(int32_t(asuint(p[i])) + (p[i] < 0.f ? -iOff : iOff)));
        r[i] = fabsf(p[i]) < origin ? p[i] + n[i] * fScale : iPos;
    }
    return mk3(r[0], r[1], r[2]);
}

// ---- PRNG: UniformSampleGenerator.slang:50-77, Xoshiro.slang:41-68, SplitMix64.slang:45-77 -------------
PLT_HD uint32_t interleave_32bit(uint32_t vx, uint32_t vy) {
    uint32_t x = vx & 0xffffu, y = vy & 0xffffu;
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    y = (y | (y << 8)) & 0x00FF00FFu; y = (y | (y << 4)) & 0x0F0F0F0Fu; y = (y | (y << 2)) & 0x33333333u; y = (y | (y << 1)) & 0x55555555u;
    return x | (y << 1);
}
PLT_HD uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    uint32_t s0, s1, s2, s3;
    PLT_HD static Rng seed(uint32_t px, uint32_t py, uint32_t sampleNumber) {
        uint64_t st = (uint64_t(sampleNumber) << 32) | uint64_t(interleave_32bit(px, py));
        const uint64_t a = splitmix64(st), b = splitmix64(st);
        return {uint32_t(a), uint32_t(a >> 32), uint32_t(b), uint32_t(b >> 32)};
    }
    PLT_HD uint32_t next() {
        const uint32_t m = s0 * 5u;
        const uint32_t result = ((m << 7) | (m >> 25)) * 9u;
        const uint32_t t = s1 << 9;
        s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3;
        s2 ^= t;
        s3 = (s3 << 11) | (s3 >> 21);
        return result;
    }
    PLT_HD float next1D() { return float(next() >> 8) * 0x1p-24f; }
    PLT_HD float2 next2D() { float2 r; r.x = next1D(); r.y = next1D(); return r; }
    PLT_HD float3 next3D() { float3 r; r.x = next1D(); r.y = next1D(); r.z = next1D(); return r; }
};

// ---- sampling helpers — MathHelpers.slang:255-449 ------------------------------------------
PLT_D float2 sample_disk(float2 u) { const float r = sqrtf(u.x), phi = k2Pi * u.y; return mk2(r * cosf(phi), r * sinf(phi)); }
PLT_D float3 sample_cone(float2 u, float cosTheta) {
    const float z = u.x * (1.f - cosTheta) + cosTheta, r = sqrtf(1.f - z * z), phi = k2Pi * u.y;
    return mk3(r * cosf(phi), r * sinf(phi), z);
}
PLT_D float3 sample_sphere(float2 u) {
    const float phi = k2Pi * u.y, c = 1.0f - 2.0f * u.x, s = sqrtf(fmaxf(0.0f, 1.0f - c * c));
    return mk3(s * cosf(phi), s * sinf(phi), c);
}
PLT_D float2 sample_disk_concentric(float2 u) {
    u = mk2(2.f * u.x - 1.f, 2.f * u.y - 1.f);
    if (u.x == 0.f && u.y == 0.f) return u;
    float phi, r;
    if (fabsf(u.x) > fabsf(u.y)) { r = u.x; phi = (u.y / u.x) * kPi_4; }
    else { r = u.y; phi = kPi_2 - (u.x / u.y) * kPi_4; }
    return mk2(r * cosf(phi), r * sinf(phi));
}
PLT_D float3 sample_cosine_hemisphere_concentric(float2 u, float& pdf) {
    const float2 d = sample_disk_concentric(u);
    const float z = sqrtf(fmaxf(0.f, 1.f - dot(d, d)));
    pdf = z * k1_Pi;
    return mk3(d.x, d.y, z);
}
PLT_D float3 sample_triangle(float2 u) {
    const float su = sqrtf(u.x);
    const float2 b = mk2(1.f - su, u.y * su);
    return mk3(1.f - b.x - b.y, b.x, b.y);
}
PLT_D float truncated_max_phi(float r, float l) {
    return (r < kFltEps || l < kFltEps) ? kPi : acosf(clampf((sqr(r) + sqr(l) - 1) / (2.f * r * l), -1.f, 1.f));
}
PLT_D float2 sample_boxmueller_truncated(float2 smp, float2 mean, float sigma2, float& pdf) {
    const float mm = dot(mean, mean);
    const float l = sqrtf(fminf(1.f, mm)), coso = sqrtf(fmaxf(.0f, 1 - mm));
    const float phi_i = (mean.x != .0f || mean.y != .0f) ? atan2f(mean.y, mean.x) : .0f;
    const float s = expf(-.5f * sqr(1 + l) / sigma2);
    const float x = (1.f - s) * fmaxf(kFltEps, smp.x) + s;
    const float r = sqrtf(-2 * sigma2 * logf(x));
    const float max_phi = truncated_max_phi(r, l);
    const float phi = phi_i + kPi + max_phi * (2.f * smp.y - 1.f);
    pdf = .5f * x / (max_phi * sigma2) * coso;
    return mk2(r * cosf(phi) + mean.x, r * sinf(phi) + mean.y);
}
PLT_D float boxmueller_truncated_pdf(float2 wo, float2 mean, float sigma2) {
    const float mm = dot(mean, mean);
    const float l = sqrtf(fminf(1.f, mm)), coso = sqrtf(fmaxf(.0f, 1 - mm));
    wo = wo - mean;
    const float r2 = dot(wo, wo);
    const float x = expf(-.5f * r2 / sigma2);
    return .5f * x / (truncated_max_phi(sqrtf(r2), l) * sigma2) * coso;
}

// Bessel J — MathHelpers.slang:808-885
PLT_D float factorialf(int n) { float r = 1.f; for (int i = 2; i <= n; ++i) r *= float(i); return r; }
PLT_D float BesselJ(int nu, float x) {
    const float t = (nu > 0 || (-nu) % 2 == 0) ? 1.f : -1.f;
    nu = abs(nu);
    const float cutoff = float(nu) + 1.f, interp = cutoff / 10.f;
    const float laa = sqrtf(2.f / kPi / x) * cosf(x - nu * kPi / 2.f - kPi / 4.f);
    float saa = nanf("");
    if (nu < 7) {
        const float x2 = sqr(x), x4 = sqr(x2), x6 = x2 * x4, x8 = sqr(x4);
        const float s = 6.f * powf(x / 2.f, (float)nu);
        const float a = 1.f / (6.f * factorialf(nu));
        const float b = -x2 / (24.f * factorialf(nu + 1));
        const float c1 = 11.f * x8 - 864.f * (6 + nu) * (x6 - 12.f * (5.f + nu) * (3.f * x4 + 32.f * (4 + nu) * (-2.f * x2 + 27.f * (3 + nu))));
        const float c = c1 / factorialf(6 + nu) * x4 * 5.652695401144601e-10f;
        saa = s * (a + b + c);
    }
    return (nu >= 7 || fabsf(x) >= cutoff + interp) ? laa * t : lerpf(saa, laa, fmaxf(.0f, fminf(1.f, .5f * (x - cutoff - interp) / interp))) * t;
}

// ---- spectrum: SpectrumUtils.slang:43-239, SpectrumSampler.slang:38-54 ------------------------
constexpr float kMinWavelength = 380.f, kMaxWavelength = 720.f, kMeanWavelength = 550.f;
__device__ const float kD65_5nm[69] = {49.9755f, 52.3118f, 54.6482f, 68.7015f, 82.7549f, 87.1204f, 91.486f, 92.4589f, 93.4318f, 90.057f, 86.6823f, 95.7736f, 104.865f, 110.936f,
    117.008f, 117.41f, 117.812f, 116.336f, 114.861f, 115.392f, 115.923f, 112.367f, 108.811f, 109.082f, 109.354f, 108.578f, 107.802f, 106.296f,
    104.79f, 106.239f, 107.689f, 106.047f, 104.405f, 104.225f, 104.046f, 102.023f, 100.f, 98.1671f, 96.3342f, 96.0611f, 95.788f, 92.2368f, 88.6856f,
    89.3459f, 90.0062f, 89.8026f, 89.5991f, 88.6489f, 87.6987f, 85.4936f, 83.2886f, 83.4939f, 83.6992f, 81.863f, 80.0268f, 80.1207f, 80.2146f,
    81.2462f, 82.2778f, 80.281f, 78.2842f, 74.0027f, 69.7213f, 70.6652f, 71.6091f, 72.979f, 74.349f, 67.9765f, 61.604f};
// 7 basis spectra x 11 bins (white, cyan, magenta, yellow, red, green, blue), SpectrumUtils.slang:116-122
__device__ const float kRgbBasis[7][11] = {
    {1.0000f, 1.0000f, 0.9999f, 0.9993f, 0.9992f, 0.9998f, 1.0000f, 1.0000f, 1.0000f, 1.0000f, 0.0f},
    {0.9710f, 0.9426f, 1.0007f, 1.0007f, 1.0007f, 1.0007f, 0.1564f, 0.0000f, 0.0000f, 0.0000f, 0.0f},
    {1.0000f, 1.0000f, 0.968f, 0.22295f, 0.0000f, 0.0458f, 0.8369f, 1.0000f, 1.0000f, 0.9959f, 0.0f},
    {0.0001f, 0.0000f, 0.1088f, 0.6651f, 1.0000f, 1.0000f, 0.9996f, 0.9586f, 0.9685f, 0.9840f, 0.0f},
    {0.1012f, 0.0515f, 0.0000f, 0.0000f, 0.0000f, 0.0000f, 0.8325f, 1.0149f, 1.0149f, 1.014f, 0.0f},
    {0.0000f, 0.0000f, 0.0273f, 0.7937f, 1.0000f, 0.9418f, 0.1719f, 0.0000f, 0.0000f, 0.0025f, 0.0f},
    {1.0000f, 1.0000f, 0.8916f, 0.3323f, 0.0000f, 0.0000f, 0.0003f, 0.0369f, 0.0483f, 0.0496f, 0.0f}};

PLT_D float xyzG(float x, float lambda, float s1, float s2) {
    const float t = (x - lambda) * (x < lambda ? s1 : s2);
    // ex2.approx: 28 of these per measured connection (7 lobes x 4 wavelengths) were 12 % of k_unwind_step's instructions with the
    // full-precision expf (profiles/r1e_unwind_lines.txt); its ~2e-6 relative error is far inside the 1e-4 colour tolerance and
    // the fit itself is only good to ~1e-2 of the CIE curves (SpectrumUtilsTests.cpp:40-80)
    return __expf(-0.5f * t * t);
}
#if PLT_FAST_DEVICE
// exp(-t^2 / 2) = 2^-(c t)^2 with c = sqrt(log2(e) / 2) folded into the two widths at compile time: subtract, select, multiply, square, ex2
PLT_D float xyzGFolded(float x, float lambda, float c1, float c2) {
    const float t = (x - lambda) * (x < lambda ? c1 : c2);
    return exp2f(-(t * t));
}
#define PLT_XYZG(l, lambda, s1, s2) xyzGFolded(l, lambda, 0.84932180028801904272f * (s1), 0.84932180028801904272f * (s2))
#else
#define PLT_XYZG(l, lambda, s1, s2) xyzG(l, lambda, s1, s2)
#endif
PLT_D float3 wavelengthToXYZ(float l) {
    return mk3(0.362f * PLT_XYZG(l, 442.0f, 0.0624f, 0.0374f) + 1.056f * PLT_XYZG(l, 599.8f, 0.0264f, 0.0323f) - 0.065f * PLT_XYZG(l, 501.1f, 0.0490f, 0.0382f),
               0.821f * PLT_XYZG(l, 568.8f, 0.0213f, 0.0247f) + 0.286f * PLT_XYZG(l, 530.9f, 0.0613f, 0.0322f),
               1.217f * PLT_XYZG(l, 437.0f, 0.0845f, 0.0278f) + 0.681f * PLT_XYZG(l, 459.0f, 0.0385f, 0.0725f));
}
PLT_D float wavelengthToD65(float lambda) {
    const float w = (lambda - kMinWavelength) * 0.2f;
    const int t = min(max(int(ceilf(w) + .5f), 0), 68), b = min(max(int(floorf(w) + .5f), 0), 68);
    const float weight = w - b;
    return kD65_5nm[t] * weight + kD65_5nm[b] * (1.0f - weight);
}
PLT_D float3 XYZtoRGB_Rec709(float3 c) {
    return mk3(3.240969941904522f * c.x + -1.537383177570094f * c.y + -0.4986107602930032f * c.z,
               -0.9692436362808803f * c.x + 1.875967501507721f * c.y + 0.04155505740717569f * c.z,
               0.05563007969699373f * c.x + -0.2039769588889765f * c.y + 1.056971514242878f * c.z);
}
PLT_D float3 spectrumToRgb(float4 spectrum, float4 wavelengths, int samples) {
    float3 result = mk3(0.f);
    PLT_FOR_LAMBDA(i, samples) {
        const float l = lane(wavelengths, i);
        const float3 xyz = wavelengthToXYZ(l) * wavelengthToD65(l) * lane(spectrum, i);
        result += XYZtoRGB_Rec709(xyz);
    }
    return result / 10.5670762f;
}
// SpectrumUtils.slang:106-160.  Which two chromatic basis spectra an RGB triple mixes depends on the ordering of r, g, b only, so
// it is resolved once per triple; per wavelength that leaves three table reads (white + two) instead of seven.  The sum keeps the
// reference's operation order ((white * w + first * c1) + second * c2), so the result is bit-identical to the seven-read form.
struct RgbBasisSel { int i1, i2; float w, c1, c2; };
PLT_D RgbBasisSel rgbBasisSelect(float3 rgb) {
    const float r = rgb.x, g = rgb.y, bl = rgb.z;
    RgbBasisSel s;
    if (r <= g && r <= bl) {
        s.w = r; s.i1 = 1;                                        // cyan
        if (g <= bl) { s.c1 = g - r; s.i2 = 6; s.c2 = bl - g; }   // blue
        else { s.c1 = bl - r; s.i2 = 5; s.c2 = g - bl; }          // green
    } else if (g <= r && g <= bl) {
        s.w = g; s.i1 = 2;                                        // magenta
        if (r <= bl) { s.c1 = r - g; s.i2 = 6; s.c2 = bl - r; }
        else { s.c1 = bl - g; s.i2 = 4; s.c2 = r - bl; }          // red
    } else {
        s.w = bl; s.i1 = 3;                                       // yellow
        if (r <= g) { s.c1 = r - bl; s.i2 = 5; s.c2 = g - r; }
        else { s.c1 = g - bl; s.i2 = 4; s.c2 = r - g; }
    }
    return s;
}
PLT_D float rgbBasisEval(const RgbBasisSel& s, float wavelength) {
    const int b = (int)(clampf((wavelength - kMinWavelength) / (kMaxWavelength - kMinWavelength), .0f, 1.f) * 10.f);
    float sd = 0.f;
    sd += kRgbBasis[0][b] * s.w;
    sd += kRgbBasis[s.i1][b] * s.c1;
    sd += kRgbBasis[s.i2][b] * s.c2;
    return sd;
}
PLT_D float rgbToSpectrum(float3 rgb, float wavelength) { return rgbBasisEval(rgbBasisSelect(rgb), wavelength); }
PLT_D float4 rgbToSpectrum4(float3 rgb, float4 wl) {
    const RgbBasisSel s = rgbBasisSelect(rgb);
    return mk4(rgbBasisEval(s, wl.x), rgbBasisEval(s, wl.y), rgbBasisEval(s, wl.z), rgbBasisEval(s, wl.w));
}
PLT_D float sampleHeroWavelength(float u) { return kMinWavelength + u * (kMaxWavelength - kMinWavelength); }
PLT_D float4 hwssWavelengths(float hero, int hwss) {
    const float span = kMaxWavelength - kMinWavelength;
    const float step = span / float(hwss);
    float4 h = mk4(hero + 0.f * step, hero + 1.f * step, hero + 2.f * step, hero + 3.f * step);
    h.x -= h.x > kMaxWavelength ? span : 0.f; h.y -= h.y > kMaxWavelength ? span : 0.f;
    h.z -= h.z > kMaxWavelength ? span : 0.f; h.w -= h.w > kMaxWavelength ? span : 0.f;
    return h;
}
PLT_D float4 waveNumbers(float4 wl) { return k2Pi / (1e-3f * wl); }

}  // namespace plt
