// Device density functions: simplex noise, fractals, terrain, stress field.
//
// Arithmetic spec: IEEE binary32, round-to-nearest-even, no
// contraction -- this library is compiled with -fmad=false -- except the
// __fmaf_rn calls written out in the simplex dot products.  Division and
// square root are the correctly rounded ones (nvcc defaults -prec-div=true
// -prec-sqrt=true).  The reference functions restated here:
//   snoise2 / snoise3      leven/cl/simplex.cl:72-230
//   BasicFractal           leven/cl/noise.cl:8-36
//   RidgedMultiFractal     leven/cl/noise.cl:42-81
//   Terrain / DensityFunc  leven/cl/noise.cl:205-268
#pragma once

#include "lvn_internal.h"

namespace lvn {

#define LVN_F2 0.366025403784f
#define LVN_G2 0.211324865405f
#define LVN_F3 0.333333333333f
#define LVN_G3 0.166666666667f

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// snoise2's gradient table is 257 x 257 (row pitch LVN_G2PITCH): texel (col, row) of the 256 x 256
// REPEAT image plus one wrapped column and row, so the "+1" corners need no second wrap.
#define LVN_G2PITCH 257

// One simplex corner: t = 0.5 - |P|^2; n = t < 0 ? 0 : t^4 * dot(grad, P)
__device__ __forceinline__ float corner2(const float2 g, float x, float y)
{
    const float t0 = 0.5f - __fmaf_rn(y, y, x * x);
    const float d = __fmaf_rn(g.y, y, g.x * x);
    const float t = t0 * t0;
    const float n = t * t * d;
    return (t0 < 0.f) ? 0.f : n;   // "if (t0 < 0.f) n0 = 0.f" tests the un-squared t0
}

__device__ __forceinline__ float snoise2(const float2 *__restrict__ grad, float px, float py)
{
    const float s = (px + py) * LVN_F2;
    const float ix = floorf(px + s), iy = floorf(py + s);
    const float t = (ix + iy) * LVN_G2;
    const float x0 = px - (ix - t), y0 = py - (iy - t);
    const int ii = __float2int_rz(ix) & 255, jj = __float2int_rz(iy) & 255;
    const bool xy = x0 > y0;
    const float o1x = xy ? 1.f : 0.f, o1y = xy ? 0.f : 1.f;
    const float2 *g = grad + (jj * LVN_G2PITCH + ii);
    const float2 g0 = __ldg(g), g1 = __ldg(g + (xy ? 1 : LVN_G2PITCH)), g2 = __ldg(g + (LVN_G2PITCH + 1));

    const float n0 = corner2(g0, x0, y0);
    const float n1 = corner2(g1, x0 - o1x + LVN_G2, y0 - o1y + LVN_G2);
    const float n2 = corner2(g2, x0 - (1.f - 2.f * LVN_G2), y0 - (1.f - 2.f * LVN_G2));
    return 70.f * (n0 + n1 + n2);
}

// the column of a corner's second lookup, one byte per texel behind the 65536 gradients: the first lookup of a
// corner used to fetch a whole float4 for its w (4 data-pipe wavefronts per warp where the bytes of a row share
// one: the L1 data pipe was the generic Hermite kernel's busiest unit, 80 %)
__device__ __forceinline__ const unsigned char *snoise3_columns(const float4 *grad3)
{
    return reinterpret_cast<const unsigned char *>(grad3 + 65536);
}

__device__ __forceinline__ float corner3(const float4 *__restrict__ grad3, int ci, int cj, int ck,
                                         float x, float y, float z)
{
    const int col = __ldg(&snoise3_columns(grad3)[((cj & 255) << 8) | (ci & 255)]);
    const float4 g = __ldg(&grad3[((ck & 255) << 8) | col]);
    const float r = __fmaf_rn(z, z, __fmaf_rn(y, y, x * x));
    // "0.6 - dot(Pf, Pf)": the unsuffixed 0.6 is a double (simplex.cl:184,196,208,220) -- subtract in
    // double, round once; 0.6f != 0.6 (snoise2's 0.5 is exact either way)
    float t = (float)(0.6 - (double)r);
    if (t < 0.f) return 0.f;
    t *= t;
    return t * t * __fmaf_rn(g.z, z, __fmaf_rn(g.y, y, g.x * x));
}

__device__ __forceinline__ float snoise3(const float4 *__restrict__ grad3, float px, float py, float pz)
{
    const float s = (px + py + pz) * LVN_F3;
    const float ix = floorf(px + s), iy = floorf(py + s), iz = floorf(pz + s);
    const float t = (ix + iy + iz) * LVN_G3;
    const float x0 = px - (ix - t), y0 = py - (iy - t), z0 = pz - (iz - t);
    const int ii = __float2int_rz(ix), jj = __float2int_rz(iy), kk = __float2int_rz(iz);
    const float isXy = (x0 < y0) ? 0.f : 1.f;
    const float isXz = (x0 < z0) ? 0.f : 1.f;
    const float isY = (y0 < z0) ? 0.f : 1.f;
    float ox = isXy + isXz, oy = 1.f - isXy, oz = 1.f - isXz;
    oy += isY;
    oz += 1.f - isY;
    const float o2x = clamp01(ox), o2y = clamp01(oy), o2z = clamp01(oz);
    const float o1x = clamp01(ox - 1.f), o1y = clamp01(oy - 1.f), o1z = clamp01(oz - 1.f);

    const float n0 = corner3(grad3, ii, jj, kk, x0, y0, z0);
    const float n1 = corner3(grad3, ii + (int)o1x, jj + (int)o1y, kk + (int)o1z,
                             x0 - o1x + LVN_G3, y0 - o1y + LVN_G3, z0 - o1z + LVN_G3);
    const float n2 = corner3(grad3, ii + (int)o2x, jj + (int)o2y, kk + (int)o2z,
                             x0 - o2x + 2.f * LVN_G3, y0 - o2y + 2.f * LVN_G3, z0 - o2z + 2.f * LVN_G3);
    const float n3 = corner3(grad3, ii + 1, jj + 1, kk + 1,
                             x0 - (1.f - 3.f * LVN_G3), y0 - (1.f - 3.f * LVN_G3), z0 - (1.f - 3.f * LVN_G3));
    return 32.f * (n0 + n1 + n2 + n3);
}

template <int OCTAVES>
__device__ __forceinline__ float basic_fractal(const float2 *__restrict__ grad, float frequency,
                                               float lacunarity, float persistence, float px, float py)
{
    float noise = 0.f, amplitude = 1.f;
    px *= frequency;
    py *= frequency;
#pragma unroll
    for (int i = 0; i < OCTAVES; i++) {
        noise += snoise2(grad, px, py) * amplitude;
        px *= lacunarity;
        py *= lacunarity;
        amplitude *= persistence;
    }
    return noise;
}

template <int OCTAVES>
__device__ __forceinline__ float ridged_multifractal(const float2 *__restrict__ grad, float lacunarity,
                                                     float gain, float offset, float px, float py)
{
    float signal = snoise2(grad, px, py);
    signal = fabsf(signal);
    signal = offset - signal;
    signal *= signal;
    float noise = signal;
    float frequency = 1.f;
#pragma unroll
    for (int i = 0; i < OCTAVES; i++) {
        px *= lacunarity;
        py *= lacunarity;
        float weight = signal * gain;
        weight = clamp01(weight);
        signal = snoise2(grad, px, py);
        signal = fabsf(signal);
        signal = offset - signal;
        signal *= weight;
        const float exponent = 1.f / frequency;   // pow(frequency, -1.f)
        frequency *= lacunarity;
        noise += signal * exponent;
    }
    noise *= (1.f / (float)OCTAVES);
    return noise;
}

// Terrain (noise.cl:205-223): depends on x and z only.
__device__ __forceinline__ float terrain(const float2 *__restrict__ grad, float x, float z)
{
    const float px = x * (1.f / 2000.f), py = z * (1.f / 2000.f);
    float ridged = 0.8f * ridged_multifractal<7>(grad, 2.114352f, 1.5241f, 1.f, px, py);
    ridged = clamp01(ridged);
    float billow = 0.6f * basic_fractal<4>(grad, 0.24f, 1.8754f, 0.433f, -4.33f * px, 7.98f * py);
    billow = (0.5f * billow) + 0.5f;
    float noise = billow * ridged;
    float b2 = 0.6f * basic_fractal<2>(grad, 0.63f, 2.2f, 0.15f, px, py);
    b2 = (b2 * 0.5f) + 0.5f;
    noise += b2;
    return noise;
}

// MAX_TERRAIN_HEIGHT * Terrain (noise.cl:266, compute.cpp:254): the surface height of a column.
// density = y - height; "density < 0" is exactly "y < height".
__device__ __forceinline__ float terrain_height(const float2 *__restrict__ grad, float x, float z)
{
    return 900.f * terrain(grad, x, z);
}

// ---------------------------------------------------------------------------
// Two Terrain evaluations per thread on sm_100's packed FP32 pipe (FADD2 / FMUL2 / FFMA2).
//
// The scalar evaluation above is bound by instruction issue: without contraction
// every flop is its own instruction.  Here a float2 holds the SAME quantity of two independent
// positions (.x = position A, .y = position B), so one packed instruction does the work of two
// and each half is the IEEE-754 binary32 round-to-nearest result of the scalar operation --
// bit-identical to terrain() by construction, and checked against it and the CPU checker in
// tests/test_parity_gpu.py.
//
// What keeps it exact:
//  * ptxas 12.9 contracts a packed multiply that feeds a packed add into one FFMA2 even for
//    mul.rn.f32x2 / add.rn.f32x2 under --fmad=false (it never does that to scalar mul.rn.f32).
//    Every packed product is therefore written as fma(a, b, nz) with nz = (-0, -0) handed in as
//    a kernel parameter: a*b + (-0) rounds exactly like a*b (and keeps the sign of a zero
//    product), costs the same single FFMA2, and ptxas cannot fold an addend it does not know.
//    tests/test_abi.py checks the SASS: no FMUL2 at all in the library.
//  * floor(v) is fl_rd(v + 1.5*2^23) - 1.5*2^23: exact for |v| < 2^22 (the sum lies in
//    [2^23, 2^24) where the floats are the integers, rounding down gives floor(v) + 1.5*2^23 and
//    the subtraction is exact), and the low mantissa bits of the biased sum are floor(v) in two's
//    complement, which gives the table index without a float->int conversion.
//  * "t0 < 0 ? 0 : t^4 d" becomes max(t0, 0)^4 d: a culled corner contributes +-0 instead of +0.
//    x + (-0) == x + (+0) unless every term is -0; a simplex point always has a live corner, and
//    a live corner that is exactly zero has the same sign in both forms, so only a sum of
//    (+-0, -0, -0) can differ: -0 here, +0 in the reference.  70*(...) keeps it a zero and every
//    caller either takes fabsf() or adds it to a running sum that starts at +0 (0 + -0 = +0).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 rep2(float v) { return make_float2(v, v); }
// packed product, uncontractable (see above)
__device__ __forceinline__ float2 mul2(float2 a, float2 b, float2 nz) { return __ffma2_rn(a, b, nz); }

// Gradient operands of a packed corner: (.x of position A, .x of position B) and the same for .y,
// so that the corner's dot product is two packed instructions.  With the gradient fetched as one
// float2 per position (LVN_X2_SPLIT_GRAD=0) the dot product has to stay scalar -- four scalar FP
// instructions per corner between packed ones, and every switch between packed and scalar FP costs
// the FMA pipe about a cycle (profiles/micro/ffma2_rate.cu).  The split tables (two float arrays, same
// 257-pitch layout) deliver the components directly as register pairs.
#ifndef LVN_X2_SPLIT_GRAD
#define LVN_X2_SPLIT_GRAD 0   // measured (B200, ring): 0 = 185.2 us, 1 = 191.4 us (twice the gathers), 2 = float2 loads + MOV transposes
#endif
struct GradTables {
    const float2 *g;     // [257*257] (x, y) per texel
    const float *gx;     // [257*257]
    const float *gy;     // [257*257]
};
__device__ __forceinline__ GradTables grad_tables(const DensityParams &dp)
{
    GradTables t; t.g = dp.grad2; t.gx = dp.grad2x; t.gy = dp.grad2x + LVN_G2PITCH * LVN_G2PITCH; return t;
}

__device__ __forceinline__ float2 corner2x2(const float2 gx, const float2 gy, float2 x, float2 y, float2 nz)
{
    const float2 t0 = sub2(rep2(0.5f), fma2(y, y, mul2(x, x, nz)));
#if LVN_X2_SPLIT_GRAD
    const float2 d = fma2(gy, y, mul2(gx, x, nz));      // fma(g.y, y, g.x * x) per position, as in corner2()
#else
    const float2 d = make_float2(__fmaf_rn(gy.x, y.x, gx.x * x.x), __fmaf_rn(gy.y, y.y, gx.y * x.y));
#endif
    const float2 tc = make_float2(fmaxf(t0.x, 0.f), fmaxf(t0.y, 0.f));
    const float2 t = mul2(tc, tc, nz);
    return mul2(mul2(t, t, nz), d, nz);
}

__device__ __forceinline__ float2 snoise2x2(const GradTables grad, float2 px, float2 py, float2 nz)
{
#ifndef LVN_X2_XU_FLOOR
    const float2 M = rep2(12582912.f);   // 1.5 * 2^23
    const float2 s = mul2(add2(px, py), rep2(LVN_F2), nz);
    const float2 mx = __fadd2_rd(add2(px, s), M), my = __fadd2_rd(add2(py, s), M);
    const float2 ix = sub2(mx, M), iy = sub2(my, M);
    const int iiA = __float_as_int(mx.x), jjA = __float_as_int(my.x), iiB = __float_as_int(mx.y), jjB = __float_as_int(my.y);
#else   // floor and float->int on the XU pipe instead of four packed adds on the FMA pipe
    const float2 s = mul2(add2(px, py), rep2(LVN_F2), nz);
    const float2 vx = add2(px, s), vy = add2(py, s);
    const float2 ix = make_float2(floorf(vx.x), floorf(vx.y)), iy = make_float2(floorf(vy.x), floorf(vy.y));
    const int iiA = __float2int_rz(ix.x), jjA = __float2int_rz(iy.x), iiB = __float2int_rz(ix.y), jjB = __float2int_rz(iy.y);
#endif
    const float2 t = mul2(add2(ix, iy), rep2(LVN_G2), nz);
    const float2 x0 = sub2(px, sub2(ix, t)), y0 = sub2(py, sub2(iy, t));
    const bool xyA = x0.x > y0.x, xyB = x0.y > y0.y;
    const float2 o1x = make_float2(xyA ? 1.f : 0.f, xyB ? 1.f : 0.f);
#ifndef LVN_X2_XU_FLOOR
    const float2 o1y = sub2(rep2(1.f), o1x);
#else
    const float2 o1y = make_float2(xyA ? 0.f : 1.f, xyB ? 0.f : 1.f);
#endif
    const int oA = (jjA & 255) * LVN_G2PITCH + (iiA & 255), oB = (jjB & 255) * LVN_G2PITCH + (iiB & 255);
    const int o1A = oA + (xyA ? 1 : LVN_G2PITCH), o1B = oB + (xyB ? 1 : LVN_G2PITCH);
#if LVN_X2_SPLIT_GRAD == 1
    const float2 g0x = make_float2(__ldg(grad.gx + oA), __ldg(grad.gx + oB)), g0y = make_float2(__ldg(grad.gy + oA), __ldg(grad.gy + oB));
    const float2 g1x = make_float2(__ldg(grad.gx + o1A), __ldg(grad.gx + o1B)), g1y = make_float2(__ldg(grad.gy + o1A), __ldg(grad.gy + o1B));
    const float2 g2x = make_float2(__ldg(grad.gx + oA + (LVN_G2PITCH + 1)), __ldg(grad.gx + oB + (LVN_G2PITCH + 1)));
    const float2 g2y = make_float2(__ldg(grad.gy + oA + (LVN_G2PITCH + 1)), __ldg(grad.gy + oB + (LVN_G2PITCH + 1)));
#else
    const float2 g0A = __ldg(grad.g + oA), g1A = __ldg(grad.g + o1A), g2A = __ldg(grad.g + oA + (LVN_G2PITCH + 1));
    const float2 g0B = __ldg(grad.g + oB), g1B = __ldg(grad.g + o1B), g2B = __ldg(grad.g + oB + (LVN_G2PITCH + 1));
    const float2 g0x = make_float2(g0A.x, g0B.x), g0y = make_float2(g0A.y, g0B.y);
    const float2 g1x = make_float2(g1A.x, g1B.x), g1y = make_float2(g1A.y, g1B.y);
    const float2 g2x = make_float2(g2A.x, g2B.x), g2y = make_float2(g2A.y, g2B.y);
#endif

    const float2 n0 = corner2x2(g0x, g0y, x0, y0, nz);
    const float2 n1 = corner2x2(g1x, g1y, add2(sub2(x0, o1x), rep2(LVN_G2)), add2(sub2(y0, o1y), rep2(LVN_G2)), nz);
    const float2 n2 = corner2x2(g2x, g2y, sub2(x0, rep2(1.f - 2.f * LVN_G2)), sub2(y0, rep2(1.f - 2.f * LVN_G2)), nz);
    return mul2(rep2(70.f), add2(add2(n0, n1), n2), nz);
}

// Per-octave constants of the fractals, folded in IEEE binary32 exactly like the unrolled scalar
// code folds them: amplitude_i = amplitude_(i-1) * persistence, frequency_i = frequency_(i-1) *
// lacunarity, exponent_i = 1 / frequency_i.
template <int N> struct OctaveTable { float v[N]; };
template <int N> __host__ __device__ constexpr OctaveTable<N> amplitude_table(float persistence)
{
    OctaveTable<N> t = {};
    float a = 1.f;
    for (int i = 0; i < N; i++) { t.v[i] = a; a *= persistence; }
    return t;
}
template <int N> __host__ __device__ constexpr OctaveTable<N> exponent_table(float lacunarity)
{
    OctaveTable<N> t = {};
    float f = 1.f;
    for (int i = 0; i < N; i++) { t.v[i] = 1.f / f; f *= lacunarity; }
    return t;
}

// Terrain's three fractals (noise.cl:207-222)
static __constant__ OctaveTable<7> c_x2_ridgedExp = exponent_table<7>(2.114352f);
static __constant__ OctaveTable<4> c_x2_billowAmp = amplitude_table<4>(0.433f);
static __constant__ OctaveTable<2> c_x2_b2Amp = amplitude_table<2>(0.15f);

// Measured on B200 (512-chunk ring, Hermite kernel): octave loops unrolled 187 us (64 registers) /
// 195 us (48); rolled with the tables above (4 snoise2 bodies instead of 14, 40 registers) 190 us;
// the evaluation out of line (__noinline__) 188 us; floor() and float->int on the XU pipe instead
// of the packed adds (LVN_X2_XU_FLOOR) 190 us.  All within 4 %: the kernel is bound by the FMA
// pipe while it evaluates noise, not by instruction fetch or occupancy.
#ifndef LVN_X2_UNROLL
#define LVN_X2_UNROLL 1
#endif

template <int OCTAVES>
__device__ __forceinline__ float2 basic_fractal_x2(const GradTables grad, float frequency, float lacunarity,
                                                   float persistence, const float *amp, float2 px, float2 py, float2 nz)
{
    float2 noise = rep2(0.f);
    px = mul2(px, rep2(frequency), nz);
    py = mul2(py, rep2(frequency), nz);
#if LVN_X2_UNROLL
    float amplitude = 1.f;
#pragma unroll
    for (int i = 0; i < OCTAVES; i++) {
        noise = add2(noise, mul2(snoise2x2(grad, px, py, nz), rep2(amplitude), nz));
        px = mul2(px, rep2(lacunarity), nz);
        py = mul2(py, rep2(lacunarity), nz);
        amplitude *= persistence;
    }
#else
#pragma unroll 1
    for (int i = 0; i < OCTAVES; i++) {
        noise = add2(noise, mul2(snoise2x2(grad, px, py, nz), rep2(amp[i]), nz));
        px = mul2(px, rep2(lacunarity), nz);
        py = mul2(py, rep2(lacunarity), nz);
    }
#endif
    return noise;
}

template <int OCTAVES>
__device__ __forceinline__ float2 ridged_multifractal_x2(const GradTables grad, float lacunarity, float gain,
                                                         float offset, const float *expo, float2 px, float2 py, float2 nz)
{
    float2 signal = snoise2x2(grad, px, py, nz);
    signal = sub2(rep2(offset), make_float2(fabsf(signal.x), fabsf(signal.y)));
    signal = mul2(signal, signal, nz);
    float2 noise = signal;
#if LVN_X2_UNROLL
    float frequency = 1.f;
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int i = 0; i < OCTAVES; i++) {
        px = mul2(px, rep2(lacunarity), nz);
        py = mul2(py, rep2(lacunarity), nz);
        const float2 weight = make_float2(clamp01(signal.x * gain), clamp01(signal.y * gain));
        signal = snoise2x2(grad, px, py, nz);
        signal = sub2(rep2(offset), make_float2(fabsf(signal.x), fabsf(signal.y)));
        signal = mul2(signal, weight, nz);
#if LVN_X2_UNROLL
        const float exponent = 1.f / frequency;   // pow(frequency, -1.f)
        frequency *= lacunarity;
#else
        const float exponent = expo[i];
#endif
        noise = add2(noise, mul2(signal, rep2(exponent), nz));
    }
    return mul2(noise, rep2(1.f / (float)OCTAVES), nz);
}

// terrain_height() at (xA, zA) and (xB, zB): x = (xA, xB), z = (zA, zB); negZero must be -0.f
#ifdef LVN_X2_NOINLINE
__device__ __noinline__ float2 terrain_height_x2(const GradTables grad, float negZero, float2 x, float2 z)
#else
__device__ __forceinline__ float2 terrain_height_x2(const GradTables grad, float negZero, float2 x, float2 z)
#endif
{
    const float2 nz = rep2(negZero);
    const float2 px = mul2(x, rep2(1.f / 2000.f), nz), py = mul2(z, rep2(1.f / 2000.f), nz);
    const float2 r = ridged_multifractal_x2<7>(grad, 2.114352f, 1.5241f, 1.f, c_x2_ridgedExp.v, px, py, nz);
    const float2 bi = basic_fractal_x2<4>(grad, 0.24f, 1.8754f, 0.433f, c_x2_billowAmp.v, mul2(px, rep2(-4.33f), nz), mul2(py, rep2(7.98f), nz), nz);
    const float2 b2 = basic_fractal_x2<2>(grad, 0.63f, 2.2f, 0.15f, c_x2_b2Amp.v, px, py, nz);
    float h[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {   // the tail of Terrain, scalar per position exactly as terrain()
        const float ridged = clamp01(0.8f * (k ? r.y : r.x));
        float billow = 0.6f * (k ? bi.y : bi.x);
        billow = (0.5f * billow) + 0.5f;
        float noise = billow * ridged;
        float b = 0.6f * (k ? b2.y : b2.x);
        b = (b * 0.5f) + 0.5f;
        noise += b;
        h[k] = 900.f * noise;
    }
    return make_float2(h[0], h[1]);
}

// BASELINE config 4 stress field: threshold - ridged fBm(snoise3), 4 octaves.
__device__ __forceinline__ float stress_density(const DensityParams &dp, float x, float y, float z)
{
    float qx = x * (1.f / 16.f), qy = y * (1.f / 16.f), qz = z * (1.f / 16.f);
    float f = 0.f, amp = 1.f;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        float r = 1.f - fabsf(snoise3(dp.grad3, qx, qy, qz));
        r *= r;
        f += r * amp;
        qx *= 2.f;
        qy *= 2.f;
        qz *= 2.f;
        amp *= 0.5f;
    }
    f *= (1.f / 1.875f);
    return dp.param - f;
}

// ---------------------------------------------------------------------------
// The stress density at two positions per thread (packed FP32, like terrain_height_x2): .x = position A,
// .y = position B; every packed op is the IEEE binary32 operation of snoise3 / stress_density on each half,
// so both halves are bit-identical to the scalar functions above (checked by the stress-field parity tests).
// What stays per position: the simplex ordering (compares and selects), the two dependent gradient
// fetches per corner, the gradient dot products (their operands come from different table rows) and
// "0.6 - dot(P, P)", which the reference evaluates in double (simplex.cl:184-220).  A culled corner is
// selected to +0 exactly as the scalar code returns it (no max() rewrite here).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float corner3_tail(const float4 *__restrict__ grad3, int ci, int cj, int ck, float r, float x, float y, float z,
                                              float &t)
{
    // returns the gradient dot; t = (float)(0.6 - (double)r), culled corners are handled by the caller
    const int col = __ldg(&snoise3_columns(grad3)[((cj & 255) << 8) | (ci & 255)]);
    const float4 g = __ldg(&grad3[((ck & 255) << 8) | col]);
    t = (float)(0.6 - (double)r);
    return __fmaf_rn(g.z, z, __fmaf_rn(g.y, y, g.x * x));
}

__device__ __forceinline__ float2 corner3x2(const float4 *__restrict__ grad3, int ciA, int cjA, int ckA, int ciB, int cjB, int ckB,
                                            float2 x, float2 y, float2 z, float2 nz)
{
    const float2 r = fma2(z, z, fma2(y, y, mul2(x, x, nz)));
    float tA, tB;
    const float dA = corner3_tail(grad3, ciA, cjA, ckA, r.x, x.x, y.x, z.x, tA);
    const float dB = corner3_tail(grad3, ciB, cjB, ckB, r.y, x.y, y.y, z.y, tB);
    const float2 t = make_float2(tA, tB);
    const float2 t2 = mul2(t, t, nz);
    const float2 n = mul2(mul2(t2, t2, nz), make_float2(dA, dB), nz);
    return make_float2(tA < 0.f ? 0.f : n.x, tB < 0.f ? 0.f : n.y);
}

__device__ __forceinline__ float2 snoise3x2(const float4 *__restrict__ grad3, float2 px, float2 py, float2 pz, float2 nz)
{
    const float2 M = rep2(12582912.f);   // 1.5 * 2^23: floor(v) = fl_rd(v + M) - M for |v| < 2^22, see snoise2x2
    const float2 s = mul2(add2(add2(px, py), pz), rep2(LVN_F3), nz);
    const float2 mx = __fadd2_rd(add2(px, s), M), my = __fadd2_rd(add2(py, s), M), mz = __fadd2_rd(add2(pz, s), M);
    const float2 ix = sub2(mx, M), iy = sub2(my, M), iz = sub2(mz, M);
    // the low mantissa bits of the biased sums are floor(v) in two's complement (only the low 8 bits are used)
    const int iiA = __float_as_int(mx.x), jjA = __float_as_int(my.x), kkA = __float_as_int(mz.x);
    const int iiB = __float_as_int(mx.y), jjB = __float_as_int(my.y), kkB = __float_as_int(mz.y);
    const float2 t = mul2(add2(add2(ix, iy), iz), rep2(LVN_G3), nz);
    const float2 x0 = sub2(px, sub2(ix, t)), y0 = sub2(py, sub2(iy, t)), z0 = sub2(pz, sub2(iz, t));
    // simplex ordering per position (simplex.cl:169-180), as the reference writes it; the offsets are exactly 0 or 1
    // (twice rewritten in integers -- same truth table, checked by brute force on the host -- and twice the
    // stress field came out different and the kernel slower: profiles/r02_notes.md 6; this form stays)
    int o1[2][3], o2[2][3];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float xx = k ? x0.y : x0.x, yy = k ? y0.y : y0.x, zz = k ? z0.y : z0.x;
        const float isXy = (xx < yy) ? 0.f : 1.f, isXz = (xx < zz) ? 0.f : 1.f, isY = (yy < zz) ? 0.f : 1.f;
        float ox = isXy + isXz, oy = 1.f - isXy, oz = 1.f - isXz;
        oy += isY;
        oz += 1.f - isY;
        o2[k][0] = (int)clamp01(ox); o2[k][1] = (int)clamp01(oy); o2[k][2] = (int)clamp01(oz);
        o1[k][0] = (int)clamp01(ox - 1.f); o1[k][1] = (int)clamp01(oy - 1.f); o1[k][2] = (int)clamp01(oz - 1.f);
    }
    const float2 o1x = make_float2(o1[0][0] ? 1.f : 0.f, o1[1][0] ? 1.f : 0.f), o1y = make_float2(o1[0][1] ? 1.f : 0.f, o1[1][1] ? 1.f : 0.f),
                 o1z = make_float2(o1[0][2] ? 1.f : 0.f, o1[1][2] ? 1.f : 0.f);
    const float2 o2x = make_float2(o2[0][0] ? 1.f : 0.f, o2[1][0] ? 1.f : 0.f), o2y = make_float2(o2[0][1] ? 1.f : 0.f, o2[1][1] ? 1.f : 0.f),
                 o2z = make_float2(o2[0][2] ? 1.f : 0.f, o2[1][2] ? 1.f : 0.f);
    const float2 g1 = rep2(LVN_G3), g2 = rep2(2.f * LVN_G3), g3 = rep2(1.f - 3.f * LVN_G3);
    const float2 n0 = corner3x2(grad3, iiA, jjA, kkA, iiB, jjB, kkB, x0, y0, z0, nz);
    const float2 n1 = corner3x2(grad3, iiA + o1[0][0], jjA + o1[0][1], kkA + o1[0][2], iiB + o1[1][0], jjB + o1[1][1], kkB + o1[1][2],
                                add2(sub2(x0, o1x), g1), add2(sub2(y0, o1y), g1), add2(sub2(z0, o1z), g1), nz);
    const float2 n2 = corner3x2(grad3, iiA + o2[0][0], jjA + o2[0][1], kkA + o2[0][2], iiB + o2[1][0], jjB + o2[1][1], kkB + o2[1][2],
                                add2(sub2(x0, o2x), g2), add2(sub2(y0, o2y), g2), add2(sub2(z0, o2z), g2), nz);
    const float2 n3 = corner3x2(grad3, iiA + 1, jjA + 1, kkA + 1, iiB + 1, jjB + 1, kkB + 1,
                                sub2(x0, g3), sub2(y0, g3), sub2(z0, g3), nz);
    return mul2(rep2(32.f), add2(add2(add2(n0, n1), n2), n3), nz);
}

// stress_density at (xA, yA, zA) and (xB, yB, zB)
__device__ __forceinline__ float2 stress_density_x2(const DensityParams &dp, float2 x, float2 y, float2 z)
{
    const float2 nz = rep2(dp.negZero);
    float2 qx = mul2(x, rep2(1.f / 16.f), nz), qy = mul2(y, rep2(1.f / 16.f), nz), qz = mul2(z, rep2(1.f / 16.f), nz);
    float2 f = rep2(0.f);
    float amp = 1.f;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const float2 n = snoise3x2(dp.grad3, qx, qy, qz, nz);
        float2 r = sub2(rep2(1.f), make_float2(fabsf(n.x), fabsf(n.y)));
        r = mul2(r, r, nz);
        f = add2(f, mul2(r, rep2(amp), nz));
        qx = mul2(qx, rep2(2.f), nz);
        qy = mul2(qy, rep2(2.f), nz);
        qz = mul2(qz, rep2(2.f), nz);
        amp *= 0.5f;
    }
    f = mul2(f, rep2(1.f / 1.875f), nz);
    return sub2(rep2(dp.param), f);
}

__device__ __forceinline__ float density3(const DensityParams &dp, float x, float y, float z)
{
    if (dp.kind == 1) return stress_density(dp, x, y, z);
    return y - terrain_height(dp.grad2, x, z);
}

// density3 at two positions
__device__ __forceinline__ float2 density3_x2(const DensityParams &dp, float2 x, float2 y, float2 z)
{
    if (dp.kind == 1) return stress_density_x2(dp, x, y, z);
    const float2 h = terrain_height_x2(grad_tables(dp), dp.negZero, x, z);
    return make_float2(y.x - h.x, y.y - h.y);
}

__device__ __forceinline__ float mixf(float a, float b, float t) { return a + (b - a) * t; }

// normalize(float3) := v * (1 / sqrt((x*x + y*y) + z*z)); the zero vector maps to itself
__device__ __forceinline__ void normalize3(float &x, float &y, float &z)
{
    const float lenSq = (x * x + y * y) + z * z;
    if (lenSq == 0.f) return;
    const float inv = 1.f / sqrtf(lenSq);
    x *= inv;
    y *= inv;
    z *= inv;
}

}  // namespace lvn
