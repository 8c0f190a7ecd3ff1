// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference algorithm; never shipped,
// never linked into the product library.  Parity status: PINNED in two ways.  (1) PRNG / computeRayOrigin / SampledSpectrum /
// XYZ-fit leaves by the reference's own known-answer procedures (tests/test_oracle_leaves.py).  (2) The PLT-specific code -- for which
// the reference holds no tests -- by the reference's OWN SHADER TEXT: oracle/ref_transpile.py rewrites the Slang modules
// textually into C++ where /root/reference lies, and oracle/ref_compare.cpp runs them next to these headers on 10^5 seeded inputs per
// function, and renders WHOLE FRAMES with the reference's two ray-generation programs (sample and solve stage, NEE / MNEE / emissive / miss /
// replay, only the intersection search behind TraceRay supplied) beside samplePixel / solvePixel: 127 entries, 100 of them bit-exact, the others
// inside documented ulp bounds; 49 974 of 49 984 frame pixels equal to the last bit (tests/test_oracle_ref.py).
// What stays outside the reference's text: ray casting itself (DXR hardware, no source in the reference: the oracle defines the contract).
//
// plt_math.h — vector types and the subset of Utils/Math/MathHelpers.slang the PLT path calls.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>

namespace orc {

// normalize(): an HLSL intrinsic whose precision the reference leaves to the GPU (DXIL lowers it to v * rsqrt(dot(v, v)) with an approximate
// rsqrt).  The oracle DEFINES it as v * (1 / sqrt(dot(v, v))) with a correctly rounded square root and reciprocal: three products instead
// of three divisions -- exactly reproducible on the CPU and on the GPU, and a sixth of the sample stage's instructions cheaper there than
// the v / length(v) form rounds 1 and 2 started with.  Divisions the Slang text spells out stay divisions.
//
// HLSL / DXIL min and max return the non-NaN operand (IEEE 754-2008 minNum / maxNum), exactly like CUDA fminf / fmaxf and
// unlike std::min / std::max.  Every min / max of the restated Slang code goes through these.
inline float hlsl_max(float a, float b) { return std::fmax(a, b); }
inline float hlsl_min(float a, float b) { return std::fmin(a, b); }
inline double hlsl_max(double a, double b) { return std::fmax(a, b); }
inline double hlsl_min(double a, double b) { return std::fmin(a, b); }
inline int hlsl_max(int a, int b) { return a < b ? b : a; }
inline int hlsl_min(int a, int b) { return b < a ? b : a; }
inline uint32_t hlsl_max(uint32_t a, uint32_t b) { return a < b ? b : a; }
inline uint32_t hlsl_min(uint32_t a, uint32_t b) { return b < a ? b : a; }
inline size_t hlsl_max(size_t a, size_t b) { return a < b ? b : a; }
inline size_t hlsl_min(size_t a, size_t b) { return b < a ? b : a; }


static constexpr float kPi = 3.14159265358979323846f;
static constexpr float k2Pi = 6.28318530717958647693f;
static constexpr float k4Pi = 12.5663706143591729539f;
static constexpr float kPi_2 = 1.57079632679489661923f;
static constexpr float kPi_4 = 0.785398163397448309616f;
static constexpr float k1_Pi = 0.318309886183790671538f;
static constexpr float k2_Pi = 0.636619772367581343076f;
static constexpr float k1_2Pi = 0.159154943091895335769f;
static constexpr float FLT_EPS = 1.192092896e-07F;
static constexpr float FLT_MINf = 1.175494351e-38F;
static constexpr float FLT_MAXf = 3.402823466e+38F;

struct float2 {
    float x = 0, y = 0;
    float2() = default;
    float2(float a, float b) : x(a), y(b) {}
    explicit float2(float a) : x(a), y(a) {}
    float& operator[](int i) { return i == 0 ? x : y; }
    float operator[](int i) const { return i == 0 ? x : y; }
};
struct float3 {
    float x = 0, y = 0, z = 0;
    float3() = default;
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit float3(float a) : x(a), y(a), z(a) {}
    float3(float2 a, float c) : x(a.x), y(a.y), z(c) {}
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float2 xy() const { return float2(x, y); }
};
struct float4 {
    float x = 0, y = 0, z = 0, w = 0;
    float4() = default;
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    explicit float4(float a) : x(a), y(a), z(a), w(a) {}
    float4(float3 a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    float3 xyz() const { return float3(x, y, z); }
};

#define ORC_VEC_OPS(T, N)                                                                                      \
    inline T operator+(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }            \
    inline T operator-(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }            \
    inline T operator*(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }            \
    inline T operator/(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }            \
    inline T operator*(T a, float s) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }           \
    inline T operator*(float s, T a) { T r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }           \
    inline T operator/(T a, float s) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }           \
    inline T operator/(float s, T a) { T r; for (int i = 0; i < N; ++i) r[i] = s / a[i]; return r; }           \
    inline T operator+(T a, float s) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }           \
    inline T operator-(T a, float s) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }           \
    inline T operator-(float s, T a) { T r; for (int i = 0; i < N; ++i) r[i] = s - a[i]; return r; }           \
    inline T operator-(T a) { T r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                       \
    inline T& operator+=(T& a, T b) { a = a + b; return a; }                                                   \
    inline T& operator-=(T& a, T b) { a = a - b; return a; }                                                   \
    inline T& operator*=(T& a, T b) { a = a * b; return a; }                                                   \
    inline T& operator*=(T& a, float s) { a = a * s; return a; }                                               \
    inline T& operator/=(T& a, float s) { a = a / s; return a; }                                               \
    inline float dot(T a, T b) { float r = a[0] * b[0]; for (int i = 1; i < N; ++i) r += a[i] * b[i]; return r; } \
    inline T vabs(T a) { T r; for (int i = 0; i < N; ++i) r[i] = std::fabs(a[i]); return r; }                  \
    inline T vmax(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = hlsl_max(a[i], b[i]); return r; }        \
    inline T vmin(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = hlsl_min(a[i], b[i]); return r; }        \
    inline T vsqrt(T a) { T r; for (int i = 0; i < N; ++i) r[i] = std::sqrt(a[i]); return r; }                 \
    inline T vexp(T a) { T r; for (int i = 0; i < N; ++i) r[i] = std::exp(a[i]); return r; }                   \
    inline T vsin(T a) { T r; for (int i = 0; i < N; ++i) r[i] = std::sin(a[i]); return r; }                   \
    inline bool any_gt(T a, float s) { for (int i = 0; i < N; ++i) if (a[i] > s) return true; return false; }  \
    inline bool all_eq(T a, float s) { for (int i = 0; i < N; ++i) if (!(a[i] == s)) return false; return true; } \
    inline bool all_finite(T a) { for (int i = 0; i < N; ++i) if (!std::isfinite(a[i])) return false; return true; } \
    inline float length(T a) { return std::sqrt(dot(a, a)); }                                                  \
    inline T normalize(T a) { return a * (1.f / length(a)); }

ORC_VEC_OPS(float2, 2)
ORC_VEC_OPS(float3, 3)
ORC_VEC_OPS(float4, 4)

inline float3 cross(float3 a, float3 b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float sqr(float v) { return v * v; }
inline float2 sqr(float2 v) { return v * v; }
inline float4 sqr(float4 v) { return v * v; }
inline float clampf(float v, float lo, float hi) { return hlsl_min(hlsl_max(v, lo), hi); }
inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }  // HLSL lerp: x + s*(y-x)
inline float4 lerp4(float4 a, float4 b, float4 t) { return a + (b - a) * t; }
inline float signf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }  // HLSL sign()
inline float fracf(float v) { return v - std::floor(v); }
inline uint32_t asuint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float asfloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int32_t asint(float f) { int32_t u; std::memcpy(&u, &f, 4); return u; }
inline float asfloat_i(int32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// IEEE half <-> float (f16tof32 / f32tof16, round-to-nearest-even as glm::packHalf / DXIL do).
inline float f16tof32(uint32_t h) {
    h &= 0xffffu;
    uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu;
    uint32_t out;
    if (exp == 0) {
        if (man == 0) out = sign;
        else {
            int e = -1;
            do { ++e; man <<= 1; } while ((man & 0x400u) == 0);
            out = sign | uint32_t(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
        }
    } else if (exp == 31) out = sign | 0x7f800000u | man << 13;
    else out = sign | (exp + 112) << 23 | man << 13;
    return asfloat(out);
}
inline uint32_t f32tof16(float f) {
    uint32_t x = asuint(f), sign = (x >> 16) & 0x8000u;
    int32_t exp = int32_t((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return sign | 0x7c00u | (man ? 0x200u : 0);
    if (exp >= 31) return sign | 0x7c00u;
    if (exp <= 0) {
        if (exp < -10) return sign;
        man |= 0x800000u;
        uint32_t shift = uint32_t(14 - exp);
        uint32_t h = man >> shift, rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return sign | h;
    }
    uint32_t h = uint32_t(exp) << 10 | man >> 13, rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
    return sign | h;
}

// ---- 2x2 / 4x4 matrices, row-major m[r][c]; mul(M, v) = M v as in Slang/HLSL ----
struct float2x2 {
    float m[2][2] = {{0, 0}, {0, 0}};
    float2x2() = default;
    float2x2(float a, float b, float c, float d) { m[0][0] = a; m[0][1] = b; m[1][0] = c; m[1][1] = d; }
};
inline float2x2 operator*(float s, const float2x2& A) { return float2x2(s * A.m[0][0], s * A.m[0][1], s * A.m[1][0], s * A.m[1][1]); }
inline float2x2 operator*(const float2x2& A, float s) { return s * A; }
inline float2x2 operator/(const float2x2& A, float s) { return float2x2(A.m[0][0] / s, A.m[0][1] / s, A.m[1][0] / s, A.m[1][1] / s); }
inline float2x2 operator+(const float2x2& A, const float2x2& B) { return float2x2(A.m[0][0] + B.m[0][0], A.m[0][1] + B.m[0][1], A.m[1][0] + B.m[1][0], A.m[1][1] + B.m[1][1]); }
inline float2x2 operator-(const float2x2& A, const float2x2& B) { return float2x2(A.m[0][0] - B.m[0][0], A.m[0][1] - B.m[0][1], A.m[1][0] - B.m[1][0], A.m[1][1] - B.m[1][1]); }
inline float2x2 mul(const float2x2& A, const float2x2& B) {
    float2x2 R;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) R.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j];
    return R;
}
inline float2 mul(const float2x2& A, float2 v) { return float2(A.m[0][0] * v.x + A.m[0][1] * v.y, A.m[1][0] * v.x + A.m[1][1] * v.y); }
inline float2x2 transpose(const float2x2& A) { return float2x2(A.m[0][0], A.m[1][0], A.m[0][1], A.m[1][1]); }
inline float determinant(const float2x2& A) { return A.m[0][0] * A.m[1][1] - A.m[0][1] * A.m[1][0]; }
inline float2x2 ident2(float d = 1.f) { return float2x2(d, 0, 0, d); }

struct float4x4 {
    float m[4][4] = {};
};
inline float4 mul(const float4x4& A, float4 v) {
    float4 r;
    for (int i = 0; i < 4; ++i) r[i] = A.m[i][0] * v.x + A.m[i][1] * v.y + A.m[i][2] * v.z + A.m[i][3] * v.w;
    return r;
}
inline float4x4 operator*(float s, float4x4 A) { for (auto& r : A.m) for (auto& v : r) v *= s; return A; }
inline float4x4 operator+(float4x4 A, const float4x4& B) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) A.m[i][j] += B.m[i][j]; return A; }
inline float4x4 ident4() { float4x4 I; for (int i = 0; i < 4; ++i) I.m[i][i] = 1.f; return I; }

// MathHelpers.slang:484-496
inline bool tryInvert(const float2x2& M, float2x2& inv) {
    float det = determinant(M);
    if (std::fabs(det) < 4e-6f) return false;
    float invdet = 1.0f / det;
    inv.m[0][0] = M.m[1][1] * invdet;
    inv.m[1][1] = M.m[0][0] * invdet;
    inv.m[0][1] = -M.m[0][1] * invdet;
    inv.m[1][0] = -M.m[1][0] * invdet;
    return true;
}

// MathHelpers.slang:583-602
inline void svd_symmetric(const float2x2& A, float2x2& U, float2x2& Sigma) {
    const float a = A.m[0][0], b = A.m[1][0], c = A.m[0][1], d = A.m[1][1];
    const float a2 = sqr(a), b2 = sqr(b), c2 = sqr(c), d2 = sqr(d);
    const float y = 2 * a * c + 2 * b * d;
    const float x = a2 + b2 - c2 - d2;
    const float theta = (x != .0f || y != .0f) ? .5f * std::atan2(y, x) : .0f;
    const float s1 = a2 + b2 + c2 + d2;
    const float s2 = std::sqrt(sqr(x) + sqr(y));
    U = float2x2(std::cos(theta), -std::sin(theta), std::sin(theta), std::cos(theta));
    Sigma = float2x2(std::sqrt((s1 + s2) / 2.f), 0, 0, std::sqrt((s1 - s2) / 2.f));
}

// MathHelpers.slang:609-621
inline float3 perp_stark(float3 u) {
    float3 a = vabs(u);
    uint32_t uyx = (a.x - a.y) < 0 ? 1 : 0;
    uint32_t uzx = (a.x - a.z) < 0 ? 1 : 0;
    uint32_t uzy = (a.y - a.z) < 0 ? 1 : 0;
    uint32_t xm = uyx & uzx;
    uint32_t ym = (1 ^ xm) & uzy;
    uint32_t zm = 1 ^ (xm | ym);
    return normalize(cross(u, float3(float(xm), float(ym), float(zm))));
}

// complex numbers as float2 — MathHelpers.slang:822-837
inline float2 c_t(float re, float im = .0f) { return float2(re, im); }
inline float cnorm2(float2 z) { return dot(z, z); }
inline float cabs(float2 z) { return std::sqrt(cnorm2(z)); }
inline float2 csqrt(float2 z) {
    const float r = cabs(z);
    const float a = std::sqrt(hlsl_max(.0f, r + z.x) / 2.f);
    const float b = std::sqrt(hlsl_max(.0f, r - z.x) / 2.f);
    return c_t(a, signf(z.y) * b);
}
inline float2 cconj(float2 z) { return c_t(z.x, -z.y); }
inline float2 cmul(float2 a, float2 b) { return c_t(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
inline float2 cdiv(float2 a, float2 b) { return cmul(a, cconj(b)) / cnorm2(b); }
inline float2 csqr(float2 z) { return cmul(z, z); }
inline float carg(float2 z) { return std::atan2(z.y, z.x); }

// Octahedral maps — MathHelpers.slang:149-240
inline float2 oct_wrap(float2 v) {
    return float2((1.f - std::fabs(v.y)) * (v.x >= 0.f ? 1.f : -1.f), (1.f - std::fabs(v.x)) * (v.y >= 0.f ? 1.f : -1.f));
}
inline float2 ndir_to_oct_snorm(float3 n) {
    float2 p = n.xy() * (1.f / (std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z)));
    p = (n.z < 0.f) ? oct_wrap(p) : p;
    return p;
}
inline float3 oct_to_ndir_snorm(float2 p) {
    float3 n = float3(p.x, p.y, 1.f - std::fabs(p.x) - std::fabs(p.y));
    float2 tmp = (n.z < 0.0f) ? oct_wrap(float2(n.x, n.y)) : float2(n.x, n.y);
    n.x = tmp.x;
    n.y = tmp.y;
    return normalize(n);
}
inline float2 ndir_to_oct_equal_area_snorm(float3 n) {
    float r = std::sqrt(1.f - std::fabs(n.z));
    float phi = std::atan2(std::fabs(n.y), std::fabs(n.x));
    float2 p;
    p.y = r * phi * k2_Pi;
    p.x = r - p.y;
    if (n.z < 0.f) p = float2(1.f - p.y, 1.f - p.x);
    return float2(p.x * signf(n.x), p.y * signf(n.y));
}
inline float3 oct_to_ndir_equal_area_snorm(float2 p) {
    float d = 1.f - (std::fabs(p.x) + std::fabs(p.y));
    float r = 1.f - std::fabs(d);
    float phi = (r > 0.f) ? ((std::fabs(p.y) - std::fabs(p.x)) / r + 1.f) * kPi_4 : 0.f;
    float f = r * std::sqrt(2.f - r * r);
    float x = f * signf(p.x) * std::cos(phi);
    float y = f * signf(p.y) * std::sin(phi);
    float z = signf(d) * (1.f - r * r);
    return float3(x, y, z);
}
// Utils/Math/FormatConversion.slang unpackSnorm2x16 / packSnorm2x16 (used by decodeNormal2x16, PackedFormats.slang:50-62)
inline float unpackSnorm16(uint32_t v) {
    int32_t i = int32_t(int16_t(v & 0xffff));
    return hlsl_max(float(i) / 32767.f, -1.f);
}
inline uint32_t packSnorm16(float v) {
    v = std::isnan(v) ? 0.f : clampf(v, -1.f, 1.f);
    return uint32_t(int32_t(std::trunc(v * 32767.f + (v >= 0.f ? 0.5f : -0.5f)))) & 0xffffu;
}
inline float3 decodeNormal2x16(uint32_t packed) {
    return oct_to_ndir_snorm(float2(unpackSnorm16(packed & 0xffff), unpackSnorm16(packed >> 16)));
}
inline uint32_t encodeNormal2x16(float3 n) {
    float2 o = ndir_to_oct_snorm(n);
    return packSnorm16(o.x) | (packSnorm16(o.y) << 16);
}

// Sampling helpers — MathHelpers.slang:255-377
inline float2 sample_disk(float2 u) {
    float r = std::sqrt(u.x), phi = k2Pi * u.y;
    return float2(r * std::cos(phi), r * std::sin(phi));
}
inline float3 sample_cone(float2 u, float cosTheta) {
    float z = u.x * (1.f - cosTheta) + cosTheta;
    float r = std::sqrt(1.f - z * z);
    float phi = k2Pi * u.y;
    return float3(r * std::cos(phi), r * std::sin(phi), z);
}
inline float3 sample_sphere(float2 u) {
    float phi = k2Pi * u.y;
    float cosTheta = 1.0f - 2.0f * u.x;
    float sinTheta = std::sqrt(hlsl_max(0.0f, 1.0f - cosTheta * cosTheta));
    return float3(sinTheta * std::cos(phi), sinTheta * std::sin(phi), cosTheta);
}
inline float2 sample_disk_concentric(float2 u) {
    u = 2.f * u - 1.f;
    if (u.x == 0.f && u.y == 0.f) return u;
    float phi, r;
    if (std::fabs(u.x) > std::fabs(u.y)) { r = u.x; phi = (u.y / u.x) * kPi_4; }
    else { r = u.y; phi = kPi_2 - (u.x / u.y) * kPi_4; }
    return r * float2(std::cos(phi), std::sin(phi));
}
inline float3 sample_cosine_hemisphere_concentric(float2 u, float& pdf) {
    float2 d = sample_disk_concentric(u);
    float z = std::sqrt(hlsl_max(0.f, 1.f - dot(d, d)));
    pdf = z * k1_Pi;
    return float3(d, z);
}
inline float cosine_hemisphere_concentric_pdf(float3 wo) { return wo.z * k1_Pi; }
inline float3 sample_triangle(float2 u) {
    float su = std::sqrt(u.x);
    float2 b = float2(1.f - su, u.y * su);
    return float3(1.f - b.x - b.y, b.x, b.y);
}
// Truncated Box-Mueller — MathHelpers.slang:416-449
inline float2 sample_boxmueller_truncated(float2 sample, float2 mean, float sigma2, float& pdf) {
    const float l = std::sqrt(hlsl_min(1.f, dot(mean, mean)));
    const float coso = std::sqrt(hlsl_max(.0f, 1 - dot(mean, mean)));
    const float phi_i = (mean.x != .0f || mean.y != .0f) ? std::atan2(mean.y, mean.x) : .0f;
    const float s = std::exp(-.5f * sqr(1 + l) / sigma2);
    const float x = (1.f - s) * hlsl_max(FLT_EPS, sample.x) + s;
    const float r = std::sqrt(-2 * sigma2 * std::log(x));
    const float max_phi = (r < FLT_EPS || l < FLT_EPS) ? kPi : std::acos(clampf((sqr(r) + sqr(l) - 1) / (2.f * r * l), -1.f, 1.f));
    const float phi = phi_i + kPi + max_phi * (2.f * sample.y - 1.f);
    const float2 p = r * float2(std::cos(phi), std::sin(phi));
    pdf = .5f * x / (max_phi * sigma2) * coso;
    return p + mean;
}
inline float boxmueller_truncated_pdf(float2 wo, float2 mean, float sigma2) {
    const float l = std::sqrt(hlsl_min(1.f, dot(mean, mean)));
    const float coso = std::sqrt(hlsl_max(.0f, 1 - dot(mean, mean)));
    wo -= mean;
    const float r2 = dot(wo, wo);
    const float x = std::exp(-.5f * r2 / sigma2);
    const float r = std::sqrt(r2);
    const float max_phi = (r < FLT_EPS || l < FLT_EPS) ? kPi : std::acos(clampf((sqr(r) + sqr(l) - 1) / (2.f * r * l), -1.f, 1.f));
    return .5f * x / (max_phi * sigma2) * coso;
}

// Bessel J — MathHelpers.slang:808-885 (used by the sinusoidal diffraction grating)
inline uint32_t factorial(uint32_t n) {
    uint32_t r = 1;
    for (uint32_t i = 2; i <= n; ++i) r *= i;
    return r;
}
inline float BesselJ_saa(int nu, float x) {
    if (nu >= 7) return std::numeric_limits<float>::quiet_NaN();
    const float x2 = sqr(x), x4 = sqr(x2), x6 = x2 * x4, x8 = sqr(x4);
    const float s = 6.f * std::pow(x / 2.f, (float)nu);
    const float a = 1.f / (6.f * factorial(nu));
    const float b = -x2 / (24.f * factorial(nu + 1));
    const float c1 = 11.f * x8 - 864.f * (6 + nu) * (x6 - 12.f * (5.f + nu) * (3.f * x4 + 32.f * (4 + nu) * (-2.f * x2 + 27.f * (3 + nu))));
    const float c = c1 / float(factorial(6 + nu)) * x4 * 5.652695401144601e-10f;
    return s * (a + b + c);
}
inline float BesselJ_laa(int nu, float x) { return std::sqrt(2.f / kPi / x) * std::cos(x - nu * kPi / 2.f - kPi / 4.f); }
inline float BesselJ(int nu, float x) {
    const float t = (nu > 0 || (-nu) % 2 == 0) ? 1.f : -1.f;
    nu = std::abs(nu);
    float cutoff = float(nu) + 1.f;
    float interp = cutoff / 10.f;
    float laa = BesselJ_laa(nu, x);
    float saa = BesselJ_saa(nu, x);
    return (nu >= 7 || std::fabs(x) >= cutoff + interp) ? laa * t : lerpf(saa, laa, hlsl_max(.0f, hlsl_min(1.f, .5f * (x - cutoff - interp) / interp))) * t;
}

// GeometryHelpers.slang:47-61 (Wächter & Binder, Ray Tracing Gems ch. 6)
inline float3 computeRayOrigin(float3 pos, float3 normal) {
    const float origin = 1.f / 16.f;
    const float fScale = 3.f / 65536.f;
    const float iScale = 3 * 256.f;
    float3 r;
    for (int i = 0; i < 3; ++i) {
        int32_t iOff = int32_t(normal[i] * iScale);
        float iPos = asfloat_i(asint(pos[i]) + (pos[i] < 0.f ? -iOff : iOff));
        float fOff = normal[i] * fScale;
        r[i] = std::fabs(pos[i]) < origin ? pos[i] + fOff : iPos;
    }
    return r;
}

}  // namespace orc
