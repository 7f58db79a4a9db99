// glm.hpp — minimal stand-in for the GLM math library, ONLY so that the reference's in-tree
// rasterizer (apps/gsrast/gscuda/{GSCuda.cu,AuxBuffer.cu}) can be compiled UNMODIFIED where it lies
// into oracle/_ref (GLM itself is not installed in this image; see oracle/Makefile).
// TEST INFRASTRUCTURE: nothing in the product includes this.
//
// It implements exactly the subset those two files use, with GLM's documented semantics and
// association order for the floating-point operators (column-major storage, m[col][row];
// mat3*mat3 and mat4*mat4 accumulate left to right; mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w);
// dot(vec4) = (x+y) + (z+w); normalize = v * (1/sqrt(dot)); min/max are the comparison forms).
#pragma once

#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>

#if defined(__CUDACC__)
#define GLMS_FN __host__ __device__ inline
#else
#define GLMS_FN inline
#endif

namespace glm {

typedef int int32;
typedef unsigned int uint32;

template <typename T> struct tvec3;
template <typename T> struct tvec4;

template <typename T>
struct tvec2 {
    union { T x, r; };
    union { T y, g; };
    GLMS_FN tvec2() : x(0), y(0) {}
    GLMS_FN tvec2(T s) : x(s), y(s) {}
    template <typename A, typename B> GLMS_FN tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
    template <typename U> GLMS_FN tvec2(const tvec2<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
    template <typename U> GLMS_FN tvec2(const tvec3<U>& v);
    GLMS_FN T& operator[](int i) { return i == 0 ? x : y; }
    GLMS_FN const T& operator[](int i) const { return i == 0 ? x : y; }
};

template <typename T>
struct tvec3 {
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    GLMS_FN tvec3() : x(0), y(0), z(0) {}
    GLMS_FN tvec3(T s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    GLMS_FN tvec3(A a, B b_, C c) : x(static_cast<T>(a)), y(static_cast<T>(b_)), z(static_cast<T>(c)) {}
    template <typename U> GLMS_FN tvec3(const tvec3<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
    template <typename U> GLMS_FN tvec3(const tvec4<U>& v);
    GLMS_FN T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    GLMS_FN const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    GLMS_FN tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};

template <typename T>
struct tvec4 {
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    union { T w, a; };
    GLMS_FN tvec4() : x(0), y(0), z(0), w(0) {}
    GLMS_FN tvec4(T s) : x(s), y(s), z(s), w(s) {}
    template <typename A, typename B, typename C, typename D>
    GLMS_FN tvec4(A a_, B b_, C c, D d) : x(static_cast<T>(a_)), y(static_cast<T>(b_)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
    template <typename U, typename S> GLMS_FN tvec4(const tvec3<U>& v, S s) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(s)) {}
    GLMS_FN T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    GLMS_FN const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

template <typename T> template <typename U> GLMS_FN tvec2<T>::tvec2(const tvec3<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
template <typename T> template <typename U> GLMS_FN tvec3<T>::tvec3(const tvec4<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;
typedef tvec2<unsigned int> uvec2;

// ---- component-wise operators -------------------------------------------------------------------
template <typename T> GLMS_FN tvec2<T> operator+(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x + b.x, a.y + b.y); }
template <typename T> GLMS_FN tvec2<T> operator-(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x - b.x, a.y - b.y); }
template <typename T> GLMS_FN tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <typename T> GLMS_FN tvec2<T> operator*(const tvec2<T>& a, T s) { return tvec2<T>(a.x * s, a.y * s); }
template <typename T> GLMS_FN tvec2<T> operator+(const tvec2<T>& a, T s) { return tvec2<T>(a.x + s, a.y + s); }

template <typename T> GLMS_FN tvec3<T> operator+(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> GLMS_FN tvec3<T> operator-(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> GLMS_FN tvec3<T> operator*(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> GLMS_FN tvec3<T> operator*(const tvec3<T>& a, T s) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> GLMS_FN tvec3<T> operator*(T s, const tvec3<T>& a) { return tvec3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T> GLMS_FN tvec3<T> operator+(const tvec3<T>& a, T s) { return tvec3<T>(a.x + s, a.y + s, a.z + s); }
template <typename T> GLMS_FN tvec3<T> operator+(T s, const tvec3<T>& a) { return tvec3<T>(s + a.x, s + a.y, s + a.z); }

template <typename T> GLMS_FN tvec4<T> operator+(const tvec4<T>& a, const tvec4<T>& b) { return tvec4<T>(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
template <typename T> GLMS_FN tvec4<T> operator*(const tvec4<T>& a, const tvec4<T>& b) { return tvec4<T>(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
template <typename T> GLMS_FN tvec4<T> operator*(const tvec4<T>& a, T s) { return tvec4<T>(a.x * s, a.y * s, a.z * s, a.w * s); }
template <typename T> GLMS_FN tvec4<T> operator*(T s, const tvec4<T>& a) { return tvec4<T>(s * a.x, s * a.y, s * a.z, s * a.w); }

// ---- scalar / vector min, max (GLM: min(x,y) = (y < x) ? y : x ; max(x,y) = (x < y) ? y : x) ---------
template <typename T> GLMS_FN T min(T x, T y) { return (y < x) ? y : x; }
template <typename T> GLMS_FN T max(T x, T y) { return (x < y) ? y : x; }
template <typename T> GLMS_FN tvec2<T> min(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(min(a.x, b.x), min(a.y, b.y)); }
template <typename T> GLMS_FN tvec2<T> max(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(max(a.x, b.x), max(a.y, b.y)); }

// ---- geometric ------------------------------------------------------------------------------------------
GLMS_FN float dot(const vec3& a, const vec3& b) { vec3 t = a * b; return t.x + t.y + t.z; }
GLMS_FN float dot(const vec4& a, const vec4& b) { vec4 t = a * b; return (t.x + t.y) + (t.z + t.w); }
GLMS_FN float inversesqrt(float x) { return 1.0f / sqrtf(x); }
GLMS_FN float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLMS_FN vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
GLMS_FN vec4 normalize(const vec4& v) { return v * inversesqrt(dot(v, v)); }

// ---- matrices (column-major: m[col][row]) -------------------------------------------------------------------
struct mat4;

struct mat3 {
    vec3 value[3];
    GLMS_FN mat3() {}
    GLMS_FN mat3(float s) { value[0] = vec3(s, 0, 0); value[1] = vec3(0, s, 0); value[2] = vec3(0, 0, s); }
    template <typename X1, typename Y1, typename Z1, typename X2, typename Y2, typename Z2, typename X3, typename Y3, typename Z3>
    GLMS_FN mat3(X1 x1, Y1 y1, Z1 z1, X2 x2, Y2 y2, Z2 z2, X3 x3, Y3 y3, Z3 z3) {
        value[0] = vec3(static_cast<float>(x1), static_cast<float>(y1), static_cast<float>(z1));
        value[1] = vec3(static_cast<float>(x2), static_cast<float>(y2), static_cast<float>(z2));
        value[2] = vec3(static_cast<float>(x3), static_cast<float>(y3), static_cast<float>(z3));
    }
    GLMS_FN mat3(const mat4& m);
    GLMS_FN vec3& operator[](int i) { return value[i]; }
    GLMS_FN const vec3& operator[](int i) const { return value[i]; }
};

struct mat4 {
    vec4 value[4];
    GLMS_FN mat4() {}
    GLMS_FN mat4(float s) { value[0] = vec4(s, 0, 0, 0); value[1] = vec4(0, s, 0, 0); value[2] = vec4(0, 0, s, 0); value[3] = vec4(0, 0, 0, s); }
    GLMS_FN vec4& operator[](int i) { return value[i]; }
    GLMS_FN const vec4& operator[](int i) const { return value[i]; }
};

GLMS_FN mat3::mat3(const mat4& m) { value[0] = vec3(m[0]); value[1] = vec3(m[1]); value[2] = vec3(m[2]); }

GLMS_FN mat3 transpose(const mat3& m) {
    mat3 r;
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 3; ++k) r[c][k] = m[k][c];
    return r;
}
GLMS_FN mat4 transpose(const mat4& m) {
    mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 4; ++k) r[c][k] = m[k][c];
    return r;
}

// Result[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]
GLMS_FN mat3 operator*(const mat3& A, const mat3& B) {
    mat3 R;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) R[c][r] = A[0][r] * B[c][0] + A[1][r] * B[c][1] + A[2][r] * B[c][2];
    return R;
}
// (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)
GLMS_FN vec4 operator*(const mat4& m, const vec4& v) {
    const vec4 Mul0 = m[0] * v.x, Mul1 = m[1] * v.y;
    const vec4 Add0 = Mul0 + Mul1;
    const vec4 Mul2 = m[2] * v.z, Mul3 = m[3] * v.w;
    const vec4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}

}  // namespace glm
