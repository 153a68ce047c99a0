// TEST INFRASTRUCTURE -- not product code.
//
// Header-only stand-in for the subset of glm 0.9.9.8 (g-truc/glm@bf71a834, the commit the
// reference pins in /root/reference/CMakeLists.txt:13-28) that the reference's translation
// units use.  glm itself is not vendored in the reference tree and cannot be fetched here, so
// the oracle build (oracle/Makefile) compiles the reference's own .cpp files, where they lie,
// against this file.
//
// Only glm's scalar code paths are restated (the reference never defines
// GLM_FORCE_INTRINSICS), with the expression order of glm's published sources, because the
// parity contract is bit-exact fp32:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z          (glm/detail/func_geometric.inl compute_dot<vec<3>>)
//   cross(a,b)    = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
//   length(v)     = sqrt(dot(v,v))
//   normalize(v)  = v * (1 / sqrt(dot(v,v)))               (v * inversesqrt(dot(v,v)))
//   reflect(I,N)  = I - N * dot(N,I) * 2
//   min(x,y)      = (y < x) ? y : x ;  max(x,y) = (x < y) ? y : x
//   clamp(x,a,b)  = min(max(x,a),b)
//   mat3 * v      = (m0.x*v.x + m1.x*v.y) + m2.x*v.z  per row
//   v / s, v /= s = true per-component division
//   glm::pow(float,int) is not a glm template match; it resolves to std::pow (double).
// Used at /root/reference/src/main.cpp:163-178,184-186,304,332; triangle.cpp:171-174;
// sphere.cpp:157; cylinder.cpp:78-116; box.cpp:12-16; kdtree.cpp:271; utils.h:60-124.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>

namespace glm {

typedef int length_t;
enum qualifier { packed_highp, packed_mediump, packed_lowp, defaultp = packed_highp };

template <length_t L, typename T, qualifier Q = defaultp> struct vec;

template <typename T, qualifier Q> struct vec<3, T, Q> {
    T x, y, z;

    constexpr vec() = default;
    constexpr vec(const vec &) = default;
    constexpr vec &operator=(const vec &) = default;
    template <typename S> constexpr explicit vec(S s) : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)) {}
    template <typename A, typename B, typename C>
    constexpr vec(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    template <typename U, qualifier P>
    constexpr vec(const vec<3, U, P> &o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}

    static constexpr length_t length() { return 3; }

    constexpr T &operator[](length_t i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T &operator[](length_t i) const { return i == 0 ? x : (i == 1 ? y : z); }

    constexpr vec &operator+=(const vec &o) { x += o.x; y += o.y; z += o.z; return *this; }
    constexpr vec &operator-=(const vec &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    constexpr vec &operator*=(const vec &o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    constexpr vec &operator/=(const vec &o) { x /= o.x; y /= o.y; z /= o.z; return *this; }
    template <typename S> constexpr vec &operator+=(S s) { x += static_cast<T>(s); y += static_cast<T>(s); z += static_cast<T>(s); return *this; }
    template <typename S> constexpr vec &operator-=(S s) { x -= static_cast<T>(s); y -= static_cast<T>(s); z -= static_cast<T>(s); return *this; }
    template <typename S> constexpr vec &operator*=(S s) { x *= static_cast<T>(s); y *= static_cast<T>(s); z *= static_cast<T>(s); return *this; }
    template <typename S> constexpr vec &operator/=(S s) { x /= static_cast<T>(s); y /= static_cast<T>(s); z /= static_cast<T>(s); return *this; }
};

template <typename T, qualifier Q> constexpr vec<3, T, Q> operator+(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return vec<3, T, Q>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator-(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return vec<3, T, Q>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator*(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return vec<3, T, Q>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator/(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return vec<3, T, Q>(a.x / b.x, a.y / b.y, a.z / b.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator-(const vec<3, T, Q> &a) { return vec<3, T, Q>(-a.x, -a.y, -a.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator+(const vec<3, T, Q> &a, T s) { return vec<3, T, Q>(a.x + s, a.y + s, a.z + s); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator-(const vec<3, T, Q> &a, T s) { return vec<3, T, Q>(a.x - s, a.y - s, a.z - s); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator*(const vec<3, T, Q> &a, T s) { return vec<3, T, Q>(a.x * s, a.y * s, a.z * s); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator/(const vec<3, T, Q> &a, T s) { return vec<3, T, Q>(a.x / s, a.y / s, a.z / s); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator+(T s, const vec<3, T, Q> &a) { return vec<3, T, Q>(s + a.x, s + a.y, s + a.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator-(T s, const vec<3, T, Q> &a) { return vec<3, T, Q>(s - a.x, s - a.y, s - a.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator*(T s, const vec<3, T, Q> &a) { return vec<3, T, Q>(s * a.x, s * a.y, s * a.z); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> operator/(T s, const vec<3, T, Q> &a) { return vec<3, T, Q>(s / a.x, s / a.y, s / a.z); }
template <typename T, qualifier Q> constexpr bool operator==(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T, qualifier Q> constexpr bool operator!=(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return !(a == b); }

typedef vec<3, float, defaultp> vec3;
typedef vec<3, std::uint8_t, defaultp> u8vec3;

using std::pow;
using std::sqrt;

template <typename T> constexpr T min(T x, T y) { return (y < x) ? y : x; }
template <typename T> constexpr T max(T x, T y) { return (x < y) ? y : x; }
template <typename T> constexpr T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }

template <typename T, qualifier Q> constexpr vec<3, T, Q> min(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return vec<3, T, Q>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> max(const vec<3, T, Q> &a, const vec<3, T, Q> &b) { return vec<3, T, Q>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
template <typename T, qualifier Q> constexpr vec<3, T, Q> clamp(const vec<3, T, Q> &v, const vec<3, T, Q> &lo, const vec<3, T, Q> &hi) { return min(max(v, lo), hi); }

template <typename T, qualifier Q> constexpr T dot(const vec<3, T, Q> &a, const vec<3, T, Q> &b)
{
    const vec<3, T, Q> tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
template <typename T, qualifier Q> constexpr vec<3, T, Q> cross(const vec<3, T, Q> &a, const vec<3, T, Q> &b)
{
    return vec<3, T, Q>(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
template <typename T, qualifier Q> inline T length(const vec<3, T, Q> &v) { return std::sqrt(dot(v, v)); }
template <typename T> inline T inversesqrt(T x) { return static_cast<T>(1) / std::sqrt(x); }
template <typename T, qualifier Q> inline vec<3, T, Q> normalize(const vec<3, T, Q> &v) { return v * inversesqrt(dot(v, v)); }
template <typename T, qualifier Q> inline vec<3, T, Q> reflect(const vec<3, T, Q> &I, const vec<3, T, Q> &N) { return I - N * dot(N, I) * static_cast<T>(2); }

template <length_t C, length_t R, typename T, qualifier Q = defaultp> struct mat;
template <typename T, qualifier Q> struct mat<3, 3, T, Q> {
    vec<3, T, Q> col[3];
    constexpr mat() = default;
    constexpr mat(const vec<3, T, Q> &c0, const vec<3, T, Q> &c1, const vec<3, T, Q> &c2) : col{c0, c1, c2} {}
    constexpr const vec<3, T, Q> &operator[](length_t i) const { return col[i]; }
    constexpr vec<3, T, Q> &operator[](length_t i) { return col[i]; }
};
typedef mat<3, 3, float, defaultp> mat3;

template <typename T, qualifier Q> constexpr vec<3, T, Q> operator*(const mat<3, 3, T, Q> &m, const vec<3, T, Q> &v)
{
    return vec<3, T, Q>(m[0].x * v.x + m[1].x * v.y + m[2].x * v.z,
                        m[0].y * v.x + m[1].y * v.y + m[2].y * v.z,
                        m[0].z * v.x + m[1].z * v.y + m[2].z * v.z);
}

} // namespace glm
