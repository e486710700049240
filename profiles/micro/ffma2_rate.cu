// ffma2_rate.cu -- issue / pipe rates of sm_100's packed FP32 (FFMA2) alone and mixed with scalar
// FFMA and integer ALU work.  8 independent chains per thread, 8 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float2 p[8];
    float s[8];
    int n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { p[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); s[i] = i + 0.25f; n[i] = threadIdx.x + i; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { s[i] = __fmaf_rn(s[i], a, b); }                                   // FFMA
            if (MODE == 1) { p[i] = __ffma2_rn(p[i], a2, b2); }                                // FFMA2
            if (MODE == 2) { p[i] = __ffma2_rn(p[i], a2, b2); s[i] = __fmaf_rn(s[i], a, b); }  // 1 : 1
            if (MODE == 3) { p[i] = __ffma2_rn(p[i], a2, b2); s[i] = __fmaf_rn(s[i], a, b); s[i] = __fmaf_rn(s[i], b, a); }   // 1 : 2
            if (MODE == 4) { p[i] = __ffma2_rn(p[i], a2, b2); n[i] = (n[i] ^ it) + i; }        // FFMA2 + 2 ALU ops
            if (MODE == 5) { p[i] = __fadd2_rn(p[i], a2); }                                    // FADD2
            if (MODE == 6) { s[i] = __fmaf_rn(s[i], a, b); n[i] = (n[i] ^ it) + i; }           // FFMA + 2 ALU ops
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += p[i].x + p[i].y + s[i] + (float)n[i];
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE> int run(const char *name, int fp_per_iter, int packed_per_iter, int other_per_iter)
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    float *out; CK(cudaMalloc(&out, 4));
    const int iters = 4096, blocks = prop.multiProcessorCount * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 64, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = prop.clockRate * 1e3;     // Hz (nominal max)
    const double cycles = ms * 1e-3 * clk;
    const double warpsPerSmsp = 4.0 * 8 / 4;     // 4 blocks x 8 warps per SM, 4 sub-partitions
    const double instPerSmsp = (double)iters * 8 * (packed_per_iter + fp_per_iter + other_per_iter) * warpsPerSmsp;
    const double flopLanes = (double)iters * 8 * (2.0 * packed_per_iter + fp_per_iter) * warpsPerSmsp * 32;   // fma lane-ops per SMSP
    printf("%-26s %7.3f ms   %5.2f inst/clk/SMSP   %5.1f fma-lane-ops/clk/SMSP (32 = scalar peak)\n", name, ms, instPerSmsp / cycles, flopLanes / cycles);
    return 0;
}

int main()
{
    run<0>("FFMA", 1, 0, 0);
    run<1>("FFMA2", 0, 1, 0);
    run<5>("FADD2", 0, 1, 0);
    run<2>("FFMA2 : FFMA = 1:1", 1, 1, 0);
    run<3>("FFMA2 : FFMA = 1:2", 2, 1, 0);
    run<4>("FFMA2 + 2 ALU", 0, 1, 2);
    run<6>("FFMA + 2 ALU", 1, 0, 2);
    return 0;
}
