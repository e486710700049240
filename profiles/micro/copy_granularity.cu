// What do the lane copies cost by themselves?  34.1 MB device -> pinned host as one memcpy, as
// 12 memcpys (4 lanes x {vertices 5.35 MB, triangles 2.58 MB, seam nodes 0.58 MB}) on one
// stream, on two streams, and through cudaMemcpyBatchAsync.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o copy_granularity copy_granularity.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
int main()
{
    const size_t sz[3] = {5350000, 2580000, 580000};
    size_t total = 0;
    for (int l = 0; l < 4; l++) for (int a = 0; a < 3; a++) total += sz[a];
    char *h, *d;
    cudaHostAlloc(&h, total, cudaHostAllocDefault);
    cudaMalloc(&d, total);
    cudaStream_t s0, s1; cudaStreamCreate(&s0); cudaStreamCreate(&s1);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, s0); cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost, s0); cudaEventRecord(e1, s0); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("1 memcpy              : %7.1f us  %5.1f GB/s\n", ms * 1e3, total / ms / 1e6);
    }
    for (int rep = 0; rep < 3; rep++) {
        size_t off = 0;
        cudaEventRecord(e0, s0);
        for (int l = 0; l < 4; l++) for (int a = 0; a < 3; a++) { cudaMemcpyAsync(h + off, d + off, sz[a], cudaMemcpyDeviceToHost, s0); off += sz[a]; }
        cudaEventRecord(e1, s0); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("12 memcpys, 1 stream  : %7.1f us  %5.1f GB/s\n", ms * 1e3, total / ms / 1e6);
    }
    for (int rep = 0; rep < 3; rep++) {
        size_t off = 0;
        cudaEventRecord(e0, s0); cudaStreamWaitEvent(s1, e0, 0);
        for (int l = 0; l < 4; l++) for (int a = 0; a < 3; a++) { cudaMemcpyAsync(h + off, d + off, sz[a], cudaMemcpyDeviceToHost, a == 0 ? s0 : s1); off += sz[a]; }
        cudaEventRecord(e2, s1); cudaStreamWaitEvent(s0, e2, 0);
        cudaEventRecord(e1, s0); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("12 memcpys, 2 streams : %7.1f us  %5.1f GB/s\n", ms * 1e3, total / ms / 1e6);
    }
#if CUDART_VERSION >= 12080
    for (int rep = 0; rep < 3; rep++) {
        std::vector<void *> dst, src; std::vector<size_t> n;
        size_t off = 0;
        for (int l = 0; l < 4; l++) for (int a = 0; a < 3; a++) { dst.push_back(h + off); src.push_back(d + off); n.push_back(sz[a]); off += sz[a]; }
        cudaMemcpyAttributes attr = {};
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t attrIdx = 0, fail = 0;
        cudaEventRecord(e0, s0);
        cudaError_t err = cudaMemcpyBatchAsync(dst.data(), src.data(), n.data(), dst.size(), &attr, &attrIdx, 1, &fail, s0);
        cudaEventRecord(e1, s0); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemcpyBatchAsync  : %7.1f us  %5.1f GB/s  (%s)\n", ms * 1e3, total / ms / 1e6, cudaGetErrorString(err));
    }
#endif
    return 0;
}
