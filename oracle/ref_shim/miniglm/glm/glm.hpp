// miniglm -- the small part of GLM 0.9.x that the reference's octree / seam code uses, so that
// leven/src/octree.cpp compiles for the host where it lies (oracle/Makefile: ref).  The reference
// does not vendor GLM.  TEST INFRASTRUCTURE (oracle/): never included by leven_b200/.
// Only integer / float vectors with component-wise arithmetic: nothing here decides a result
// beyond IEEE / two's-complement arithmetic on the components.
#pragma once
#define GLM_VERSION 93   // GLM 0.9.3.x defines it (core/setup.hpp); include/leven_compute.hpp keys its glm support on it
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace glm {

template <class T> struct tvec2 {
    T x, y;
    tvec2() : x(0), y(0) {}
    explicit tvec2(T s) : x(s), y(s) {}
    tvec2(T a, T b) : x(a), y(b) {}
    T &operator[](int i) { return (&x)[i]; }
    const T &operator[](int i) const { return (&x)[i]; }
};
template <class T> struct tvec4;
template <class T> struct tvec3 {
    T x, y, z;
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T s) : x(s), y(s), z(s) {}
    template <class A, class B, class C> tvec3(A a, B b, C c) : x((T)a), y((T)b), z((T)c) {}
    template <class U> explicit tvec3(const tvec3<U> &o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <class U> explicit tvec3(const tvec4<U> &o);
    T &operator[](int i) { return (&x)[i]; }
    const T &operator[](int i) const { return (&x)[i]; }
};
template <class T> struct alignas(16) tvec4 {   // 16-byte aligned: the simplifier _mm_load_ps's MeshVertex members
    T x, y, z, w;
    tvec4() : x(0), y(0), z(0), w(0) {}
    explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    template <class A, class B, class C, class D> tvec4(A a, B b, C c, D d) : x((T)a), y((T)b), z((T)c), w((T)d) {}
    template <class U, class D> tvec4(const tvec3<U> &o, D d) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)d) {}
    template <class U> explicit tvec4(const tvec4<U> &o) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)o.w) {}
    T &operator[](int i) { return (&x)[i]; }
    const T &operator[](int i) const { return (&x)[i]; }
};
template <class T> template <class U> tvec3<T>::tvec3(const tvec4<U> &o) : x((T)o.x), y((T)o.y), z((T)o.z) {}

typedef tvec2<float> vec2; typedef tvec3<float> vec3; typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;  typedef tvec3<int> ivec3;  typedef tvec4<int> ivec4;
typedef tvec3<unsigned> uvec3; typedef tvec4<unsigned> uvec4;

#define MINIGLM_OP3(op) \
    template <class T> tvec3<T> operator op(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x op b.x, a.y op b.y, a.z op b.z); } \
    template <class T> tvec3<T> operator op(const tvec3<T> &a, T b) { return tvec3<T>(a.x op b, a.y op b, a.z op b); } \
    template <class T> tvec3<T> operator op(T a, const tvec3<T> &b) { return tvec3<T>(a op b.x, a op b.y, a op b.z); } \
    template <class T> tvec3<T> &operator op##=(tvec3<T> &a, const tvec3<T> &b) { a = a op b; return a; } \
    template <class T> tvec3<T> &operator op##=(tvec3<T> &a, T b) { a = a op b; return a; }
MINIGLM_OP3(+) MINIGLM_OP3(-) MINIGLM_OP3(*) MINIGLM_OP3(/)
#undef MINIGLM_OP3
#define MINIGLM_OP4(op) \
    template <class T> tvec4<T> operator op(const tvec4<T> &a, const tvec4<T> &b) { return tvec4<T>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    template <class T> tvec4<T> operator op(const tvec4<T> &a, T b) { return tvec4<T>(a.x op b, a.y op b, a.z op b, a.w op b); } \
    template <class T> tvec4<T> operator op(T a, const tvec4<T> &b) { return tvec4<T>(a op b.x, a op b.y, a op b.z, a op b.w); } \
    template <class T> tvec4<T> &operator op##=(tvec4<T> &a, const tvec4<T> &b) { a = a op b; return a; } \
    template <class T> tvec4<T> &operator op##=(tvec4<T> &a, T b) { a = a op b; return a; }
MINIGLM_OP4(+) MINIGLM_OP4(-) MINIGLM_OP4(*) MINIGLM_OP4(/)
#undef MINIGLM_OP4
// integer-only operators
inline ivec3 operator%(const ivec3 &a, int b) { return ivec3(a.x % b, a.y % b, a.z % b); }
inline ivec3 operator%(const ivec3 &a, const ivec3 &b) { return ivec3(a.x % b.x, a.y % b.y, a.z % b.z); }
inline ivec3 operator&(const ivec3 &a, int b) { return ivec3(a.x & b, a.y & b, a.z & b); }
inline ivec3 operator>>(const ivec3 &a, int b) { return ivec3(a.x >> b, a.y >> b, a.z >> b); }
inline ivec3 operator<<(const ivec3 &a, int b) { return ivec3(a.x << b, a.y << b, a.z << b); }
template <class T> tvec3<T> operator-(const tvec3<T> &a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <class T> tvec4<T> operator-(const tvec4<T> &a) { return tvec4<T>(-a.x, -a.y, -a.z, -a.w); }
template <class T> bool operator==(const tvec2<T> &a, const tvec2<T> &b) { return a.x == b.x && a.y == b.y; }
template <class T> bool operator==(const tvec3<T> &a, const tvec3<T> &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <class T> bool operator!=(const tvec3<T> &a, const tvec3<T> &b) { return !(a == b); }
template <class T> bool operator==(const tvec4<T> &a, const tvec4<T> &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
template <class T> bool operator!=(const tvec4<T> &a, const tvec4<T> &b) { return !(a == b); }

template <class T> T min(T a, T b) { return b < a ? b : a; }
template <class T> T max(T a, T b) { return a < b ? b : a; }
template <class T> T abs(T a) { return a < 0 ? -a : a; }
template <class T> T clamp(T v, T lo, T hi) { return min(max(v, lo), hi); }
template <class T> tvec3<T> min(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
template <class T> tvec3<T> max(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4 &a, const vec4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(const vec3 &a) { return std::sqrt(dot(a, a)); }
inline float length2(const vec3 &a) { return dot(a, a); }   // gtx/norm.inl
inline float length2(const vec4 &a) { return dot(a, a); }
inline vec3 normalize(const vec3 &a) { return a * (1.f / length(a)); }
inline vec3 cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline int log2(int v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }

}  // namespace glm
