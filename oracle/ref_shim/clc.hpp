// clc.hpp -- just enough of the OpenCL C 1.2 language environment, as C++14, to compile the
// reference's kernel sources (leven/cl/*.cl) for the host CPU where they lie.
//
// TEST INFRASTRUCTURE (oracle/): only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs may execute what is built from this.  Nothing under leven_b200/ includes it.
//
// The kernels' control flow, tables, indexing, constants and operand order come from the
// reference text itself.  What this header has to DEFINE are the OpenCL built-ins, which the
// reference built with -cl-fast-relaxed-math on a vendor runtime (compute.cpp:258) and which are
// therefore implementation-defined to a few ulp.  They follow the arithmetic spec of DESIGN.md 2
// (the same choices the C restatement oracle/lvn_oracle.c and the CUDA kernels make):
//   dot(float2/3/4)  = fma chain: fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))          (simplex.cl dot products)
//   normalize(v)     = v * (1 / sqrt((x*x + y*y) + z*z)), zero vector -> itself
//   length(v)        = sqrt((x*x + y*y) + z*z)        rsqrt(x) = 1 / sqrt(x)
//   pow(x, -1.f)     = 1 / x                          mix(a,b,t) = a + (b - a) * t
//   clamp(x,lo,hi)   = fmin(fmax(x, lo), hi)          step(e, x) = x < e ? 0 : 1
//   sin / cos        = libm sinf / cosf               every other operation: IEEE binary32, no contraction
//   read_imagef      = CLK_FILTER_NEAREST | CLK_ADDRESS_REPEAT | CLK_NORMALIZED_COORDS_TRUE per the
//                      OpenCL 1.2 spec 8.2: u = (s - floor(s)) * w, i = (int)floor(u), i > w-1 -> i - w;
//                      CL_UNORM_INT8 -> byte / 255.f
// Compile with -ffp-contract=off -Wno-narrowing (OpenCL C allows int -> float in vector literals).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace clc {

typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned char uchar;
typedef unsigned short ushort;

template <class T, int N> struct vec;

// A swizzle: a view of an N-component vector's storage selecting components I...
template <class T, int N, int... I> struct swz {
    T d[N];
    static constexpr int K = sizeof...(I);
    operator vec<T, K>() const { return vec<T, K>(d[I]...); }
    swz &operator=(const vec<T, K> &v) { const int idx[] = {I...}; for (int k = 0; k < K; k++) d[idx[k]] = v.s[k]; return *this; }
    swz &operator=(const swz &o) { return *this = (vec<T, K>)o; }
    template <int M, int... J> swz &operator=(const swz<T, M, J...> &o) { return *this = (vec<T, K>)o; }
    swz &operator+=(const vec<T, K> &v) { return *this = (vec<T, K>)(*this) + v; }
    swz &operator-=(const vec<T, K> &v) { return *this = (vec<T, K>)(*this) - v; }
    swz &operator*=(const vec<T, K> &v) { return *this = (vec<T, K>)(*this) * v; }
    swz &operator*=(T v) { return *this = (vec<T, K>)(*this) * v; }
    swz &operator+=(T v) { return *this = (vec<T, K>)(*this) + v; }
    swz &operator-=(T v) { return *this = (vec<T, K>)(*this) - v; }
};

template <class T> struct vec<T, 2> {
    union {
        struct { T x, y; };
        T s[2];
        swz<T,2,0,0> xx; swz<T,2,0,1> xy; swz<T,2,1,0> yx; swz<T,2,1,1> yy; swz<T,2,0,0,0> xxx;
        swz<T,2,0,0,1> xxy; swz<T,2,0,1,0> xyx; swz<T,2,0,1,1> xyy; swz<T,2,1,0,0> yxx; swz<T,2,1,0,1> yxy;
        swz<T,2,1,1,0> yyx; swz<T,2,1,1,1> yyy;
    };
    vec() : s{} {}
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    explicit vec(S v) : s{(T)v, (T)v} {}
    vec(T a, T b) : s{a, b} {}
    vec(const vec &o) : s{o.s[0], o.s[1]} {}
    vec &operator=(const vec &o) { s[0] = o.s[0]; s[1] = o.s[1]; return *this; }
};

template <class T> struct vec<T, 3> {
    union {
        struct { T x, y, z; };
        T s[3];
        swz<T,3,0,0> xx; swz<T,3,0,1> xy; swz<T,3,0,2> xz; swz<T,3,1,0> yx; swz<T,3,1,1> yy; swz<T,3,1,2> yz;
        swz<T,3,2,0> zx; swz<T,3,2,1> zy; swz<T,3,2,2> zz; swz<T,3,0,0,0> xxx; swz<T,3,0,0,1> xxy;
        swz<T,3,0,0,2> xxz; swz<T,3,0,1,0> xyx; swz<T,3,0,1,1> xyy; swz<T,3,0,1,2> xyz; swz<T,3,0,2,0> xzx;
        swz<T,3,0,2,1> xzy; swz<T,3,0,2,2> xzz; swz<T,3,1,0,0> yxx; swz<T,3,1,0,1> yxy; swz<T,3,1,0,2> yxz;
        swz<T,3,1,1,0> yyx; swz<T,3,1,1,1> yyy; swz<T,3,1,1,2> yyz; swz<T,3,1,2,0> yzx; swz<T,3,1,2,1> yzy;
        swz<T,3,1,2,2> yzz; swz<T,3,2,0,0> zxx; swz<T,3,2,0,1> zxy; swz<T,3,2,0,2> zxz; swz<T,3,2,1,0> zyx;
        swz<T,3,2,1,1> zyy; swz<T,3,2,1,2> zyz; swz<T,3,2,2,0> zzx; swz<T,3,2,2,1> zzy; swz<T,3,2,2,2> zzz;
    };
    vec() : s{} {}
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    explicit vec(S v) : s{(T)v, (T)v, (T)v} {}
    vec(T a, T b, T c) : s{a, b, c} {}
    vec(const vec<T, 2> &a, T c) : s{a.s[0], a.s[1], c} {}
    vec(T a, const vec<T, 2> &b) : s{a, b.s[0], b.s[1]} {}
    vec(const vec &o) : s{o.s[0], o.s[1], o.s[2]} {}
    vec &operator=(const vec &o) { s[0] = o.s[0]; s[1] = o.s[1]; s[2] = o.s[2]; return *this; }
};

template <class T> struct vec<T, 4> {
    union {
        struct { T x, y, z, w; };
        T s[4];
        swz<T,4,0,0> xx; swz<T,4,0,1> xy; swz<T,4,0,2> xz; swz<T,4,0,3> xw; swz<T,4,1,0> yx; swz<T,4,1,1> yy;
        swz<T,4,1,2> yz; swz<T,4,1,3> yw; swz<T,4,2,0> zx; swz<T,4,2,1> zy; swz<T,4,2,2> zz; swz<T,4,2,3> zw;
        swz<T,4,3,0> wx; swz<T,4,3,1> wy; swz<T,4,3,2> wz; swz<T,4,3,3> ww; swz<T,4,0,0,0> xxx;
        swz<T,4,0,0,1> xxy; swz<T,4,0,0,2> xxz; swz<T,4,0,0,3> xxw; swz<T,4,0,1,0> xyx; swz<T,4,0,1,1> xyy;
        swz<T,4,0,1,2> xyz; swz<T,4,0,1,3> xyw; swz<T,4,0,2,0> xzx; swz<T,4,0,2,1> xzy; swz<T,4,0,2,2> xzz;
        swz<T,4,0,2,3> xzw; swz<T,4,0,3,0> xwx; swz<T,4,0,3,1> xwy; swz<T,4,0,3,2> xwz; swz<T,4,0,3,3> xww;
        swz<T,4,1,0,0> yxx; swz<T,4,1,0,1> yxy; swz<T,4,1,0,2> yxz; swz<T,4,1,0,3> yxw; swz<T,4,1,1,0> yyx;
        swz<T,4,1,1,1> yyy; swz<T,4,1,1,2> yyz; swz<T,4,1,1,3> yyw; swz<T,4,1,2,0> yzx; swz<T,4,1,2,1> yzy;
        swz<T,4,1,2,2> yzz; swz<T,4,1,2,3> yzw; swz<T,4,1,3,0> ywx; swz<T,4,1,3,1> ywy; swz<T,4,1,3,2> ywz;
        swz<T,4,1,3,3> yww; swz<T,4,2,0,0> zxx; swz<T,4,2,0,1> zxy; swz<T,4,2,0,2> zxz; swz<T,4,2,0,3> zxw;
        swz<T,4,2,1,0> zyx; swz<T,4,2,1,1> zyy; swz<T,4,2,1,2> zyz; swz<T,4,2,1,3> zyw; swz<T,4,2,2,0> zzx;
        swz<T,4,2,2,1> zzy; swz<T,4,2,2,2> zzz; swz<T,4,2,2,3> zzw; swz<T,4,2,3,0> zwx; swz<T,4,2,3,1> zwy;
        swz<T,4,2,3,2> zwz; swz<T,4,2,3,3> zww; swz<T,4,3,0,0> wxx; swz<T,4,3,0,1> wxy; swz<T,4,3,0,2> wxz;
        swz<T,4,3,0,3> wxw; swz<T,4,3,1,0> wyx; swz<T,4,3,1,1> wyy; swz<T,4,3,1,2> wyz; swz<T,4,3,1,3> wyw;
        swz<T,4,3,2,0> wzx; swz<T,4,3,2,1> wzy; swz<T,4,3,2,2> wzz; swz<T,4,3,2,3> wzw; swz<T,4,3,3,0> wwx;
        swz<T,4,3,3,1> wwy; swz<T,4,3,3,2> wwz; swz<T,4,3,3,3> www; swz<T,4,0,1,2,3> xyzw;
    };
    vec() : s{} {}
    template <class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    explicit vec(S v) : s{(T)v, (T)v, (T)v, (T)v} {}
    vec(T a, T b, T c, T d) : s{a, b, c, d} {}
    vec(const vec<T, 3> &a, T d) : s{a.s[0], a.s[1], a.s[2], d} {}
    vec(const vec<T, 2> &a, T c, T d) : s{a.s[0], a.s[1], c, d} {}
    vec(const vec<T, 2> &a, const vec<T, 2> &b) : s{a.s[0], a.s[1], b.s[0], b.s[1]} {}
    vec(const vec &o) : s{o.s[0], o.s[1], o.s[2], o.s[3]} {}
    vec &operator=(const vec &o) { s[0] = o.s[0]; s[1] = o.s[1]; s[2] = o.s[2]; s[3] = o.s[3]; return *this; }
};

typedef vec<float, 2> float2; typedef vec<float, 3> float3; typedef vec<float, 4> float4;
typedef vec<int, 2> int2;     typedef vec<int, 3> int3;     typedef vec<int, 4> int4;
typedef vec<uint, 2> uint2;   typedef vec<uint, 3> uint3;   typedef vec<uint, 4> uint4;

// ---- "vector-like": a vec or a swizzle of one ----
template <class A> struct vl { static const bool is = false; };
template <class T, int N> struct vl<vec<T, N>> {
    static const bool is = true; typedef T elem; static const int n = N;
    static const vec<T, N> &get(const vec<T, N> &a) { return a; }
};
template <class T, int N, int... I> struct vl<swz<T, N, I...>> {
    static const bool is = true; typedef T elem; static const int n = sizeof...(I);
    static vec<T, n> get(const swz<T, N, I...> &a) { return (vec<T, n>)a; }
};
#define CLC_VV template <class A, class B, class = typename std::enable_if<vl<A>::is && vl<B>::is && vl<A>::n == vl<B>::n>::type>
#define CLC_VS template <class A, class S, class = typename std::enable_if<vl<A>::is && std::is_arithmetic<S>::value>::type>
#define CLC_SV template <class S, class A, class = typename std::enable_if<std::is_arithmetic<S>::value && vl<A>::is>::type, class = void>
#define CLC_V  template <class A, class = typename std::enable_if<vl<A>::is>::type>
#define CLC_RV vec<typename vl<A>::elem, vl<A>::n>

// component-wise operators; a scalar operand is converted to the element type (OpenCL C 6.3)
#define CLC_BINOP(op) \
    CLC_VV CLC_RV operator op(const A &a, const B &b) { auto x = vl<A>::get(a); auto y = vl<B>::get(b); CLC_RV r; \
        for (int i = 0; i < vl<A>::n; i++) r.s[i] = x.s[i] op y.s[i]; return r; } \
    CLC_VS CLC_RV operator op(const A &a, S b) { auto x = vl<A>::get(a); const typename vl<A>::elem y = (typename vl<A>::elem)b; CLC_RV r; \
        for (int i = 0; i < vl<A>::n; i++) r.s[i] = x.s[i] op y; return r; } \
    CLC_SV CLC_RV operator op(S a, const A &b) { const typename vl<A>::elem x = (typename vl<A>::elem)a; auto y = vl<A>::get(b); CLC_RV r; \
        for (int i = 0; i < vl<A>::n; i++) r.s[i] = x op y.s[i]; return r; }
CLC_BINOP(+) CLC_BINOP(-) CLC_BINOP(*) CLC_BINOP(/) CLC_BINOP(&) CLC_BINOP(|) CLC_BINOP(^) CLC_BINOP(<<) CLC_BINOP(>>)
#undef CLC_BINOP
CLC_V CLC_RV operator-(const A &a) { auto x = vl<A>::get(a); CLC_RV r; for (int i = 0; i < vl<A>::n; i++) r.s[i] = -x.s[i]; return r; }
#define CLC_ASSIGNOP(op) \
    template <class T, int N, class B, class = typename std::enable_if<vl<B>::is && vl<B>::n == N>::type> \
    vec<T, N> &operator op##=(vec<T, N> &a, const B &b) { a = a op b; return a; } \
    template <class T, int N, class S, class = typename std::enable_if<std::is_arithmetic<S>::value>::type, class = void> \
    vec<T, N> &operator op##=(vec<T, N> &a, S b) { a = a op b; return a; }
CLC_ASSIGNOP(+) CLC_ASSIGNOP(-) CLC_ASSIGNOP(*) CLC_ASSIGNOP(/)
#undef CLC_ASSIGNOP

// ---- scalar built-ins (float) ----
inline float sqrt(float x) { return ::sqrtf(x); }
inline float rsqrt(float x) { return 1.f / ::sqrtf(x); }
inline float fabs(float x) { return ::fabsf(x); }
inline float floor(float x) { return ::floorf(x); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float atan(float x) { return ::atanf(x); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline float fmod(float x, float y) { return ::fmodf(x, y); }
inline float pow(float x, float y) { return y == -1.f ? 1.f / x : ::powf(x, y); }
inline float radians(float d) { return d * 0.017453292519943295f; }
inline float sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
inline float step(float edge, float x) { return x < edge ? 0.f : 1.f; }
inline float mix(float a, float b, float t) { return a + (b - a) * t; }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline uint max(uint a, uint b) { return a > b ? a : b; }
inline float clamp(float x, float lo, float hi) { return ::fminf(::fmaxf(x, lo), hi); }
inline double clamp(double x, double lo, double hi) { return ::fmin(::fmax(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
inline int clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline uint clz(uint x) { return x == 0 ? 32u : (uint)__builtin_clz(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
// double forms, for expressions the reference writes with unsuffixed literals
inline double sqrt(double x) { return ::sqrt(x); }
inline double fabs(double x) { return ::fabs(x); }

// ---- vector built-ins ----
#define CLC_MAP1(name) CLC_V CLC_RV name(const A &a) { auto x = vl<A>::get(a); CLC_RV r; \
        for (int i = 0; i < vl<A>::n; i++) r.s[i] = name(x.s[i]); return r; }
CLC_MAP1(floor) CLC_MAP1(fabs) CLC_MAP1(sqrt) CLC_MAP1(sin) CLC_MAP1(cos) CLC_MAP1(sign)
#undef CLC_MAP1
#define CLC_MAP2(name) \
    CLC_VV CLC_RV name(const A &a, const B &b) { auto x = vl<A>::get(a); auto y = vl<B>::get(b); CLC_RV r; \
        for (int i = 0; i < vl<A>::n; i++) r.s[i] = name(x.s[i], y.s[i]); return r; } \
    CLC_VS CLC_RV name(const A &a, S b) { auto x = vl<A>::get(a); CLC_RV r; \
        for (int i = 0; i < vl<A>::n; i++) r.s[i] = name(x.s[i], (typename vl<A>::elem)b); return r; }
CLC_MAP2(min) CLC_MAP2(max) CLC_MAP2(fmod) CLC_MAP2(step)
#undef CLC_MAP2
CLC_SV CLC_RV step(S e, const A &b) { auto y = vl<A>::get(b); CLC_RV r; for (int i = 0; i < vl<A>::n; i++) r.s[i] = step((float)e, y.s[i]); return r; }
CLC_V CLC_RV clamp(const A &a, typename vl<A>::elem lo, typename vl<A>::elem hi) { auto x = vl<A>::get(a); CLC_RV r;
    for (int i = 0; i < vl<A>::n; i++) r.s[i] = clamp(x.s[i], lo, hi); return r; }
CLC_VV CLC_RV mix(const A &a, const B &b, float t) { return vl<A>::get(a) + (vl<B>::get(b) - vl<A>::get(a)) * t; }

inline float dot(float a, float b) { return a * b; }
inline float dot(const float2 &a, const float2 &b) { return ::fmaf(a.y, b.y, a.x * b.x); }
inline float dot(const float3 &a, const float3 &b) { return ::fmaf(a.z, b.z, ::fmaf(a.y, b.y, a.x * b.x)); }
inline float dot(const float4 &a, const float4 &b) { return ::fmaf(a.w, b.w, ::fmaf(a.z, b.z, ::fmaf(a.y, b.y, a.x * b.x))); }
inline float length(float a) { return ::fabsf(a); }
inline float length(const float2 &v) { return ::sqrtf(v.x * v.x + v.y * v.y); }
inline float length(const float3 &v) { return ::sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z); }
inline float length(const float4 &v) { return ::sqrtf(((v.x * v.x + v.y * v.y) + v.z * v.z) + v.w * v.w); }
inline float2 normalize(const float2 &v) { const float l = v.x * v.x + v.y * v.y; return l == 0.f ? v : v * (1.f / ::sqrtf(l)); }
inline float3 normalize(const float3 &v) { const float l = (v.x * v.x + v.y * v.y) + v.z * v.z; return l == 0.f ? v : v * (1.f / ::sqrtf(l)); }
inline float4 normalize(const float4 &v) { const float l = ((v.x * v.x + v.y * v.y) + v.z * v.z) + v.w * v.w; return l == 0.f ? v : v * (1.f / ::sqrtf(l)); }
inline float3 cross(const float3 &a, const float3 &b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

inline float4 convert_float4(const int4 &v) { return float4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
inline float3 convert_float3(const int3 &v) { return float3((float)v.x, (float)v.y, (float)v.z); }
inline int4 convert_int4(const float4 &v) { return int4((int)v.x, (int)v.y, (int)v.z, (int)v.w); }

// ---- images ----
struct image2d { const unsigned char *rgba; int width, height; };
typedef const image2d *image2d_t;
typedef int sampler_t;
enum { CLK_FILTER_NEAREST = 1, CLK_ADDRESS_REPEAT = 2, CLK_NORMALIZED_COORDS_TRUE = 4, CLK_LOCAL_MEM_FENCE = 1, CLK_GLOBAL_MEM_FENCE = 2 };
inline int clc_repeat_nearest(float s, int w)
{
    const float u = (s - ::floorf(s)) * (float)w;
    int i = (int)::floorf(u);
    if (i > w - 1) i -= w;
    return i;
}
inline float4 read_imagef(image2d_t img, sampler_t, const float2 &c)
{
    const int i = clc_repeat_nearest(c.x, img->width), j = clc_repeat_nearest(c.y, img->height);
    const unsigned char *p = img->rgba + 4 * ((size_t)j * img->width + i);
    return float4(p[0] / 255.f, p[1] / 255.f, p[2] / 255.f, p[3] / 255.f);
}

// ---- work-items: an NDRange is a loop nest in the launcher (clc_launch) ----
struct work_item { size_t gid[3], gsize[3]; };
extern thread_local work_item clc_wi;
inline size_t get_global_id(uint d) { return clc_wi.gid[d]; }
inline size_t get_global_size(uint d) { return clc_wi.gsize[d]; }

// ---- 64-bit atomics (cl_khr_int64_base_atomics) ----
inline ulong atom_xchg(ulong *p, ulong v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline ulong atom_cmpxchg(ulong *p, ulong cmp, ulong v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }

}  // namespace clc

// address-space and access qualifiers vanish on the host (define these AFTER every standard header)
#define kernel
#define __kernel
#define global
#define __global
#define constant const
#define __constant const
#define local
#define __local
#define read_only
#define write_only
#define __read_only
#define __write_only
