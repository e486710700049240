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

__device__ __forceinline__ float corner3(const float4 *__restrict__ grad3, int ci, int cj, int ck,
                                         float x, float y, float z)
{
    const int col = __float_as_int(__ldg(&grad3[((cj & 255) << 8) | (ci & 255)]).w);
    const float4 g = __ldg(&grad3[((ck & 255) << 8) | col]);
    const float r = __fmaf_rn(z, z, __fmaf_rn(y, y, x * x));
    float t = 0.6f - r;
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

__device__ __forceinline__ float density3(const DensityParams &dp, float x, float y, float z)
{
    if (dp.kind == 1) return stress_density(dp, x, y, z);
    return y - terrain_height(dp.grad2, x, z);
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
