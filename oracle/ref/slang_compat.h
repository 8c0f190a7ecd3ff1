// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// slang_compat.h — the part of the Slang / HLSL core language the reference's PLT shader modules use, as C++17, so that those
// modules can be compiled for the CPU after the purely textual rewrite of oracle/ref_transpile.py (slangc is not available
// offline; DESIGN.md "Oracle").  Nothing in here is PLT-specific: vector / matrix types with HLSL's component-wise operators,
// broadcasting scalars, comparison masks, and the intrinsic functions.  All arithmetic is fp32 (the transpiler suffixes every
// floating literal with `f`; build with -ffp-contract=off).  The generated code lives in namespace `ref`.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <type_traits>

namespace ref {

typedef uint32_t uint;

template <class T, int N> struct vec;
template <class T> struct vec<T, 1> {   // vector<float, 1>: hwss_t at HWSS = 1
    T x{};
    vec() = default;
    template <class U, class = std::enable_if_t<std::is_arithmetic<U>::value>> vec(U s) : x(T(s)) {}
    template <class U> explicit vec(const vec<U, 1>& o) : x(T(o.x)) {}
    T& operator[](int) { return x; }
    const T& operator[](int) const { return x; }
};
template <class T> struct vec<T, 2> {
    T x{}, y{};
    vec() = default;
    vec(T a, T b) : x(a), y(b) {}
    template <class U, class = std::enable_if_t<std::is_arithmetic<U>::value>> vec(U s) : x(T(s)), y(T(s)) {}  // HLSL scalar broadcast
    template <class U> explicit vec(const vec<U, 2>& o) : x(T(o.x)), y(T(o.y)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    vec<T, 2> xy() const { return {x, y}; }
    vec<T, 2> yx() const { return {y, x}; }
    vec<T, 2> xx() const { return {x, x}; }
    vec<T, 2> yy() const { return {y, y}; }
};
template <class T> struct vec<T, 3> {
    T x{}, y{}, z{};
    vec() = default;
    vec(T a, T b, T c) : x(a), y(b), z(c) {}
    vec(vec<T, 2> a, T c) : x(a.x), y(a.y), z(c) {}
    template <class U, class S, class = std::enable_if_t<!std::is_same<U, T>::value && std::is_arithmetic<S>::value>> vec(vec<U, 2> a, S c) : x(T(a.x)), y(T(a.y)), z(T(c)) {}  // int3(uint2, int)
    vec(T a, vec<T, 2> b) : x(a), y(b.x), z(b.y) {}
    template <class U, class = std::enable_if_t<std::is_arithmetic<U>::value>> vec(U s) : x(T(s)), y(T(s)), z(T(s)) {}
    template <class U> explicit vec(const vec<U, 3>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    vec<T, 2> xy() const { return {x, y}; }
    vec<T, 2> xz() const { return {x, z}; }
    vec<T, 2> yz() const { return {y, z}; }
    vec<T, 2> zz() const { return {z, z}; }
    vec<T, 3> xyz() const { return {x, y, z}; }
};
template <class T> struct vec<T, 4> {
    T x{}, y{}, z{}, w{};
    vec() = default;
    vec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    vec(vec<T, 3> a, T d) : x(a.x), y(a.y), z(a.z), w(d) {}
    vec(vec<T, 2> a, vec<T, 2> b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    vec(vec<T, 2> a, T c, T d) : x(a.x), y(a.y), z(c), w(d) {}
    template <class U, class = std::enable_if_t<std::is_arithmetic<U>::value>> vec(U s) : x(T(s)), y(T(s)), z(T(s)), w(T(s)) {}
    template <class U> explicit vec(const vec<U, 4>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    vec<T, 2> xy() const { return {x, y}; }
    vec<T, 2> yz() const { return {y, z}; }
    vec<T, 2> zw() const { return {z, w}; }
    vec<T, 3> xyz() const { return {x, y, z}; }
};
typedef vec<float, 1> float1;
typedef vec<float, 2> float2;
typedef vec<float, 3> float3;
typedef vec<float, 4> float4;
typedef vec<int, 2> int2;
typedef vec<int, 3> int3;
typedef vec<int, 4> int4;
typedef vec<uint, 2> uint2;
typedef vec<uint, 3> uint3;
typedef vec<uint, 4> uint4;
typedef vec<bool, 2> bool2;
typedef vec<bool, 3> bool3;
typedef vec<bool, 4> bool4;

// usual arithmetic conversions of HLSL: int (op) float -> float, bool counts as int
template <class A, class B> using arith_t = decltype(std::declval<std::conditional_t<std::is_same<A, bool>::value, int, A>>() + std::declval<std::conditional_t<std::is_same<B, bool>::value, int, B>>());
template <class T> using if_scalar = std::enable_if_t<std::is_arithmetic<T>::value>;

#define REF_BINOP(OP)                                                                                                                              \
    template <class A, class B, int N> vec<arith_t<A, B>, N> operator OP(const vec<A, N>& a, const vec<B, N>& b) {                                \
        vec<arith_t<A, B>, N> r; for (int i = 0; i < N; ++i) r[i] = arith_t<A, B>(a[i]) OP arith_t<A, B>(b[i]); return r; }                        \
    template <class A, class B, int N, class = if_scalar<B>> vec<arith_t<A, B>, N> operator OP(const vec<A, N>& a, B b) {                         \
        vec<arith_t<A, B>, N> r; for (int i = 0; i < N; ++i) r[i] = arith_t<A, B>(a[i]) OP arith_t<A, B>(b); return r; }                           \
    template <class A, class B, int N, class = if_scalar<A>> vec<arith_t<A, B>, N> operator OP(A a, const vec<B, N>& b) {                         \
        vec<arith_t<A, B>, N> r; for (int i = 0; i < N; ++i) r[i] = arith_t<A, B>(a) OP arith_t<A, B>(b[i]); return r; }
REF_BINOP(+) REF_BINOP(-) REF_BINOP(*) REF_BINOP(/)
#undef REF_BINOP
template <class A, int N> vec<A, N> operator-(const vec<A, N>& a) { vec<A, N> r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }
#define REF_ASSIGNOP(OP)                                                                                                                           \
    template <class A, class B, int N> vec<A, N>& operator OP##=(vec<A, N>& a, const vec<B, N>& b) { for (int i = 0; i < N; ++i) a[i] = A(a[i] OP b[i]); return a; } \
    template <class A, class B, int N, class = if_scalar<B>> vec<A, N>& operator OP##=(vec<A, N>& a, B b) { for (int i = 0; i < N; ++i) a[i] = A(a[i] OP b); return a; }
REF_ASSIGNOP(+) REF_ASSIGNOP(-) REF_ASSIGNOP(*) REF_ASSIGNOP(/)
#undef REF_ASSIGNOP
#define REF_CMPOP(OP)                                                                                                                              \
    template <class A, class B, int N> vec<bool, N> operator OP(const vec<A, N>& a, const vec<B, N>& b) { vec<bool, N> r; for (int i = 0; i < N; ++i) r[i] = a[i] OP b[i]; return r; } \
    template <class A, class B, int N, class = if_scalar<B>> vec<bool, N> operator OP(const vec<A, N>& a, B b) { vec<bool, N> r; for (int i = 0; i < N; ++i) r[i] = a[i] OP b; return r; } \
    template <class A, class B, int N, class = if_scalar<A>> vec<bool, N> operator OP(A a, const vec<B, N>& b) { vec<bool, N> r; for (int i = 0; i < N; ++i) r[i] = a OP b[i]; return r; }
REF_CMPOP(<) REF_CMPOP(<=) REF_CMPOP(>) REF_CMPOP(>=) REF_CMPOP(==) REF_CMPOP(!=)
#undef REF_CMPOP
template <int N> vec<bool, N> operator&&(const vec<bool, N>& a, const vec<bool, N>& b) { vec<bool, N> r; for (int i = 0; i < N; ++i) r[i] = a[i] && b[i]; return r; }
template <int N> vec<bool, N> operator||(const vec<bool, N>& a, const vec<bool, N>& b) { vec<bool, N> r; for (int i = 0; i < N; ++i) r[i] = a[i] || b[i]; return r; }
template <int N> vec<bool, N> operator!(const vec<bool, N>& a) { vec<bool, N> r; for (int i = 0; i < N; ++i) r[i] = !a[i]; return r; }
template <int N> bool all(const vec<bool, N>& a) { bool r = true; for (int i = 0; i < N; ++i) r = r && a[i]; return r; }
template <int N> bool any(const vec<bool, N>& a) { bool r = false; for (int i = 0; i < N; ++i) r = r || a[i]; return r; }
inline bool all(bool a) { return a; }
inline bool any(bool a) { return a; }
// Slang's `cond ? a : b` with a vector condition selects per component (the transpiler rewrites those few ternaries to select())
template <class A, class B, int N> vec<arith_t<A, B>, N> select(const vec<bool, N>& c, const vec<A, N>& a, const vec<B, N>& b) {
    vec<arith_t<A, B>, N> r; for (int i = 0; i < N; ++i) r[i] = c[i] ? arith_t<A, B>(a[i]) : arith_t<A, B>(b[i]); return r; }
template <class A, class B, int N, class = if_scalar<B>> vec<arith_t<A, B>, N> select(const vec<bool, N>& c, const vec<A, N>& a, B b) { return select(c, a, vec<B, N>(b)); }
template <class A, class B, int N, class = if_scalar<A>> vec<arith_t<A, B>, N> select(const vec<bool, N>& c, A a, const vec<B, N>& b) { return select(c, vec<A, N>(a), b); }
template <class A, class B, int N, class = if_scalar<A>, class = if_scalar<B>> vec<arith_t<A, B>, N> select(const vec<bool, N>& c, A a, B b) { return select(c, vec<A, N>(a), vec<B, N>(b)); }
// shifts of integer vectors by a scalar count
template <class T, int N, class = std::enable_if_t<std::is_integral<T>::value>> vec<T, N> operator>>(const vec<T, N>& a, int s) { vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = T(a[i] >> s); return r; }
template <class T, int N, class = std::enable_if_t<std::is_integral<T>::value>> vec<T, N> operator<<(const vec<T, N>& a, int s) { vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = T(a[i] << s); return r; }

// ---- scalar intrinsics (fp32; min / max follow IEEE minNum / maxNum like DXIL and like the oracle's hlsl_min / hlsl_max)
inline float abs(float v) { return std::fabs(v); }
inline int abs(int v) { return v < 0 ? -v : v; }
inline float sqrt(float v) { return std::sqrt(v); }
inline float rsqrt(float v) { return 1.0f / std::sqrt(v); }
inline float exp(float v) { return std::exp(v); }
inline float exp2(float v) { return std::exp2(v); }
inline float log(float v) { return std::log(v); }
inline float log2(float v) { return std::log2(v); }
inline float sin(float v) { return std::sin(v); }
inline float cos(float v) { return std::cos(v); }
inline float tan(float v) { return std::tan(v); }
inline float asin(float v) { return std::asin(v); }
inline float acos(float v) { return std::acos(v); }
inline float atan(float v) { return std::atan(v); }
inline float atan2(float y, float x) { return std::atan2(y, x); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float floor(float v) { return std::floor(v); }
inline float ceil(float v) { return std::ceil(v); }
inline float round(float v) { return std::nearbyint(v); }
inline float trunc(float v) { return std::trunc(v); }
inline float frac(float v) { return v - std::floor(v); }
inline int hlsl_ftoi(float v) { return v != v ? 0 : (v >= 2147483648.f ? 2147483647 : (v <= -2147483648.f ? int(-2147483647 - 1) : int(v))); }   // HLSL (int)float: saturating, NaN -> 0
inline float fmod(float a, float b) { return std::fmod(a, b); }
inline float sign(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
inline int sign(int v) { return v > 0 ? 1 : (v < 0 ? -1 : 0); }
inline bool isfinite(float v) { return std::isfinite(v); }
inline bool isnan(float v) { return std::isnan(v); }
inline bool isinf(float v) { return std::isinf(v); }
inline float saturate(float v) { return std::fmin(std::fmax(v, 0.f), 1.f); }
inline float smoothstep(float a, float b, float x) { const float t = saturate((x - a) / (b - a)); return t * t * (3.f - 2.f * t); }  // HLSL intrinsic
inline float lerp(float a, float b, float t) { return a + t * (b - a); }  // HLSL lerp: x + s*(y-x)
inline float step(float e, float x) { return x >= e ? 1.f : 0.f; }
inline float radians(float d) { return d * 0.0174532925199432957692f; }
template <class A, class B, class = if_scalar<A>, class = if_scalar<B>> arith_t<A, B> max(A a, B b) {
    typedef arith_t<A, B> R;
    if constexpr (std::is_floating_point<R>::value) return R(std::fmax(R(a), R(b))); else return R(a) < R(b) ? R(b) : R(a);
}
template <class A, class B, class = if_scalar<A>, class = if_scalar<B>> arith_t<A, B> min(A a, B b) {
    typedef arith_t<A, B> R;
    if constexpr (std::is_floating_point<R>::value) return R(std::fmin(R(a), R(b))); else return R(b) < R(a) ? R(b) : R(a);
}
template <class A, class B, class C, class = if_scalar<A>, class = if_scalar<B>, class = if_scalar<C>> auto clamp(A v, B lo, C hi) { return min(max(v, lo), hi); }
inline uint asuint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline uint asuint(uint u) { return u; }
inline float asfloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float asfloat(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float asfloat(float f) { return f; }
inline double asdouble(uint lo, uint hi) { const uint64_t b = (uint64_t(hi) << 32) | lo; double d; std::memcpy(&d, &b, 8); return d; }   // HLSL asdouble(lowbits, highbits)
inline void asuint(double d, uint& lo, uint& hi) { uint64_t b; std::memcpy(&b, &d, 8); lo = uint(b); hi = uint(b >> 32); }          // HLSL asuint(double, out lowbits, out highbits)
inline int asint(float f) { int i; std::memcpy(&i, &f, 4); return i; }
// HLSL f16tof32 / f32tof16 (the low 16 bits are an IEEE binary16; conversion to half rounds to nearest even)
inline float f16tof32(uint h) {
    const uint s = (h >> 15) & 1u, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    float v;
    if (e == 0) v = std::ldexp(float(m), -24);
    else if (e == 31) v = m ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity();
    else v = std::ldexp(float(m | 0x400u), int(e) - 25);
    return s ? -v : v;
}
inline uint f32tof16(float f) {
    uint x; std::memcpy(&x, &f, 4);
    const uint s = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return s | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u);
    if (x >= 0x477ff000u) return s | 0x7c00u;                       // rounds to infinity
    if (x < 0x33000001u) return s;                                  // rounds to zero
    int e = int(x >> 23) - 127;
    uint m = (x & 0x7fffffu) | 0x800000u;
    int shift = e < -14 ? 13 + (-14 - e) : 13;
    uint h = m >> shift;
    const uint rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    if (e < -14) return s | h;                                      // subnormal half (h may carry into the smallest normal)
    return s | ((uint(e + 15) << 10) + (h - 0x400u));
}

// ---- component-wise vector intrinsics
#define REF_MAP1(F) template <class T, int N> auto F(const vec<T, N>& a) { vec<decltype(F(a[0])), N> r; for (int i = 0; i < N; ++i) r[i] = F(a[i]); return r; }
REF_MAP1(abs) REF_MAP1(sqrt) REF_MAP1(rsqrt) REF_MAP1(exp) REF_MAP1(exp2) REF_MAP1(log) REF_MAP1(log2) REF_MAP1(sin) REF_MAP1(cos) REF_MAP1(tan) REF_MAP1(asin) REF_MAP1(acos) REF_MAP1(atan)
REF_MAP1(floor) REF_MAP1(ceil) REF_MAP1(round) REF_MAP1(trunc) REF_MAP1(frac) REF_MAP1(sign) REF_MAP1(isfinite) REF_MAP1(isnan) REF_MAP1(isinf) REF_MAP1(saturate)
REF_MAP1(asint) REF_MAP1(asuint) REF_MAP1(asfloat)
#undef REF_MAP1
#define REF_MAP2(F)                                                                                                                                 \
    template <class A, class B, int N> auto F(const vec<A, N>& a, const vec<B, N>& b) { vec<decltype(F(a[0], b[0])), N> r; for (int i = 0; i < N; ++i) r[i] = F(a[i], b[i]); return r; } \
    template <class A, class B, int N, class = if_scalar<B>> auto F(const vec<A, N>& a, B b) { vec<decltype(F(a[0], b)), N> r; for (int i = 0; i < N; ++i) r[i] = F(a[i], b); return r; }   \
    template <class A, class B, int N, class = if_scalar<A>> auto F(A a, const vec<B, N>& b) { vec<decltype(F(a, b[0])), N> r; for (int i = 0; i < N; ++i) r[i] = F(a, b[i]); return r; }
REF_MAP2(max) REF_MAP2(min) REF_MAP2(pow) REF_MAP2(atan2) REF_MAP2(fmod) REF_MAP2(step)
#undef REF_MAP2
template <class T, int N, class L, class H> vec<T, N> clamp(const vec<T, N>& v, const L& lo, const H& hi) { return vec<T, N>(min(max(v, lo), hi)); }
template <class T, int N, class S> vec<T, N> lerp(const vec<T, N>& a, const vec<T, N>& b, const S& t) { return a + t * (b - a); }
template <class T, int N> T dot(const vec<T, N>& a, const vec<T, N>& b) { T r = a[0] * b[0]; for (int i = 1; i < N; ++i) r += a[i] * b[i]; return r; }
inline float dot(float a, float b) { return a * b; }
template <int N> float length(const vec<float, N>& a) { return std::sqrt(dot(a, a)); }
template <int N> float distance(const vec<float, N>& a, const vec<float, N>& b) { return length(a - b); }
template <int N> vec<float, N> normalize(const vec<float, N>& a) { return a * (1.f / length(a)); }  // the oracle's definition of the intrinsic (oracle/plt_math.h)
inline float3 cross(const float3& a, const float3& b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <int N> vec<float, N> reflect(const vec<float, N>& i, const vec<float, N>& n) { return i - 2.f * dot(n, i) * n; }

// ---- matrices (row-major: M[row][col], HLSL mul(M, v) = row . v)
template <int R, int C> struct mat {
    vec<float, C> rows[R];
    mat() = default;
    template <class... S, class = std::enable_if_t<sizeof...(S) == size_t(R * C) && (std::is_arithmetic<S>::value && ...)>> mat(S... s) {
        const float v[] = {float(s)...};
        for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) rows[r][c] = v[r * C + c];
    }
    template <class... V, class = std::enable_if_t<sizeof...(V) == size_t(R) && (std::is_same<V, vec<float, C>>::value && ...)>, class = void> mat(V... v) {
        const vec<float, C> t[] = {v...};
        for (int r = 0; r < R; ++r) rows[r] = t[r];
    }
    // (float3x3)M of a float4x4: the upper-left block (HLSL matrix truncation cast; LightHelpers.slang sampleDistantLight)
    template <int R2, int C2, class = std::enable_if_t<(R2 >= R && C2 >= C && (R2 != R || C2 != C))>> explicit mat(const mat<R2, C2>& M) {
        for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) rows[r][c] = M[r][c];
    }
    vec<float, C>& operator[](int r) { return rows[r]; }
    const vec<float, C>& operator[](int r) const { return rows[r]; }
};
typedef mat<2, 2> float2x2;
typedef mat<3, 3> float3x3;
typedef mat<4, 4> float4x4;
typedef mat<3, 4> float3x4;
typedef mat<2, 3> float2x3;
template <int R, int C> vec<float, R> mul(const mat<R, C>& M, const vec<float, C>& v) { vec<float, R> r; for (int i = 0; i < R; ++i) r[i] = dot(M[i], v); return r; }
template <int R, int C> vec<float, C> mul(const vec<float, R>& v, const mat<R, C>& M) {
    vec<float, C> r; for (int c = 0; c < C; ++c) { float s = v[0] * M[0][c]; for (int i = 1; i < R; ++i) s += v[i] * M[i][c]; r[c] = s; } return r; }
template <int R, int K, int C> mat<R, C> mul(const mat<R, K>& A, const mat<K, C>& B) {
    mat<R, C> r;
    for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) { float s = A[i][0] * B[0][j]; for (int k = 1; k < K; ++k) s += A[i][k] * B[k][j]; r[i][j] = s; }
    return r;
}
template <int R, int C> mat<C, R> transpose(const mat<R, C>& M) { mat<C, R> r; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) r[j][i] = M[i][j]; return r; }
inline float determinant(const float2x2& M) { return M[0][0] * M[1][1] - M[0][1] * M[1][0]; }
inline float determinant(const float3x3& M) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}
#define REF_MATOP(OP)                                                                                                                               \
    template <int R, int C> mat<R, C> operator OP(const mat<R, C>& A, const mat<R, C>& B) { mat<R, C> r; for (int i = 0; i < R; ++i) r[i] = A[i] OP B[i]; return r; } \
    template <int R, int C, class S, class = if_scalar<S>> mat<R, C> operator OP(const mat<R, C>& A, S s) { mat<R, C> r; for (int i = 0; i < R; ++i) r[i] = A[i] OP float(s); return r; } \
    template <int R, int C, class S, class = if_scalar<S>> mat<R, C> operator OP(S s, const mat<R, C>& A) { mat<R, C> r; for (int i = 0; i < R; ++i) r[i] = float(s) OP A[i]; return r; }
REF_MATOP(+) REF_MATOP(-) REF_MATOP(*) REF_MATOP(/)
#undef REF_MATOP
template <int R, int C> mat<R, C> operator-(const mat<R, C>& A) { mat<R, C> r; for (int i = 0; i < R; ++i) r[i] = -A[i]; return r; }

}  // namespace ref
