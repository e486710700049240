// Can SM stores fill PCIe as well as the copy engine?  Writes N bytes into mapped pinned host
// memory with coalesced float4 stores (full 128 B lines), and with 48 B-strided float4 stores
// (the natural MeshVertex pattern), against cudaMemcpyAsync D2H of the same bytes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o zero_copy_bw zero_copy_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_coalesced(float4 *dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_float4((float)i, 1.f, 2.f, 3.f);
}
__global__ void k_strided(float4 *dst, size_t nrec)   // 3 float4 per 48 B record, one record per thread
{
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrec; r += (size_t)gridDim.x * blockDim.x) {
        dst[3 * r + 0] = make_float4((float)r, 1.f, 2.f, 3.f);
        dst[3 * r + 1] = make_float4(4.f, 5.f, 6.f, 7.f);
        dst[3 * r + 2] = make_float4(8.f, 9.f, 10.f, 11.f);
    }
}
int main()
{
    const size_t bytes = 48ull << 20, n = bytes / 16;
    float4 *h, *hd, *d;
    cudaHostAlloc(&h, bytes, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&hd, h, 0);
    cudaMalloc(&d, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("memcpy D2H      : %7.1f us  %6.1f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    int grids[] = {8, 32, 148, 592, 2368};
    for (int g : grids)
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0); k_coalesced<<<g, 256>>>(hd, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); printf("coalesced g=%4d : %7.1f us  %6.1f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
        }
    for (int g : grids)
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0); k_strided<<<g, 256>>>(hd, n / 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); printf("strided48 g=%4d : %7.1f us  %6.1f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
