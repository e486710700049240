// boundary_cost.cu -- what does a kernel boundary cost while a bulk device->host copy is in flight,
// and which launch mechanism avoids it?  (profiles/r01c_pipeline.md measured ~30 us per boundary.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o boundary_cost boundary_cost.cu && ./boundary_cost
//
// A chain of K kernels (148 blocks x 256 threads, each spinning ~SPIN_US) on one stream, timed by
// wall clock, alone and with a 256 MB D2H copy running on a second stream.  Variants:
//   plain        ordinary launches
//   domain       cudaLaunchAttributeMemSyncDomain = remote on the compute kernels
//   pdl          programmatic stream serialization, griddepcontrol.wait at the top of each kernel
//   pdl+trigger  the same, each kernel also triggers its dependents at its start
//   graph        the chain captured into a CUDA graph
//   coop         ONE cooperative kernel, K phases separated by grid.sync()
//   flag         ONE ordinary kernel of 148 resident blocks, K phases separated by a hand-written
//                global-memory barrier (atomic counter + spin), no cooperative launch
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void spin(long long cycles)
{
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
}

__global__ void k_plain(long long cycles, int *sink) { spin(cycles); if (sink && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sink, 1); }
__global__ void k_pdl(long long cycles, int *sink, int trigger)
{
    if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    spin(cycles);
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sink, 1);
}
__global__ void k_coop(long long cycles, int phases, int *sink)
{
    cg::grid_group g = cg::this_grid();
    for (int p = 0; p < phases; p++) {
        spin(cycles);
        if (sink && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sink, 1);
        g.sync();
    }
}
__global__ void k_flag(long long cycles, int phases, int *sink, unsigned int *bar)
{
    for (int p = 0; p < phases; p++) {
        spin(cycles);
        if (sink && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sink, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            const unsigned int want = (unsigned int)(p + 1) * gridDim.x;
            while (*(volatile unsigned int *)bar < want) { }
            __threadfence();
        }
        __syncthreads();
    }
}

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
    const int K = 40, BLOCKS = 148, THREADS = 256, REPS = 10;
    const double SPIN_US = 10.0;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const long long cycles = (long long)(SPIN_US * 1e-6 * prop.clockRate * 1e3);
    printf("%s, %d SMs, chain of %d kernels x %.0f us\n", prop.name, prop.multiProcessorCount, K, SPIN_US);

    cudaStream_t s, side; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    int *sink; CK(cudaMalloc(&sink, 4)); CK(cudaMemset(sink, 0, 4));
    unsigned int *bar; CK(cudaMalloc(&bar, 4));
    const size_t NB = 256u << 20;
    char *dsrc, *hdst; CK(cudaMalloc(&dsrc, NB)); CK(cudaMallocHost(&hdst, NB));

    cudaGraphExec_t gexec = nullptr;
    {
        cudaGraph_t graph;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < K; i++) k_plain<<<BLOCKS, THREADS, 0, s>>>(cycles, sink);
        CK(cudaStreamEndCapture(s, &graph));
        CK(cudaGraphInstantiate(&gexec, graph, 0));
    }

    auto chain = [&](int variant) {
        if (variant == 4) { CK(cudaGraphLaunch(gexec, s)); return; }
        if (variant == 5) {
            int phases = K; long long c = cycles; int *sk = sink;
            void *args[] = {&c, &phases, &sk};
            CK(cudaLaunchCooperativeKernel((void *)k_coop, dim3(BLOCKS), dim3(THREADS), args, 0, s));
            return;
        }
        if (variant == 6) {
            CK(cudaMemsetAsync(bar, 0, 4, s));
            k_flag<<<BLOCKS, THREADS, 0, s>>>(cycles, K, sink, bar);
            return;
        }
        for (int i = 0; i < K; i++) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(BLOCKS); cfg.blockDim = dim3(THREADS); cfg.stream = s;
            cudaLaunchAttribute at[2]; int na = 0;
            if (variant == 1) { at[na].id = cudaLaunchAttributeMemSyncDomain; at[na].val.memSyncDomain = cudaLaunchMemSyncDomainRemote; na++; }
            if (variant == 2 || variant == 3) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; na++; }
            cfg.attrs = at; cfg.numAttrs = na;
            if (variant == 2 || variant == 3) CK(cudaLaunchKernelEx(&cfg, k_pdl, cycles, sink, variant == 3 ? 1 : 0));
            else CK(cudaLaunchKernelEx(&cfg, k_plain, cycles, sink));
        }
    };
    const char *names[] = {"plain", "domain", "pdl", "pdl+trigger", "graph", "coop", "flag"};
    for (int variant = 0; variant < 7; variant++) {
        double t[2] = {0, 0};
        for (int copying = 0; copying < 2; copying++) {
            for (int r = -2; r < REPS; r++) {
                CK(cudaDeviceSynchronize());
                if (copying) CK(cudaMemcpyAsync(hdst, dsrc, NB, cudaMemcpyDeviceToHost, side));   // ~4.5 ms in flight
                const double t0 = now_us();
                chain(variant);
                CK(cudaStreamSynchronize(s));
                const double dt = now_us() - t0;
                if (r >= 0) t[copying] += dt / REPS;
            }
        }
        printf("%-12s alone %8.1f us   with D2H %8.1f us   extra per boundary %6.2f us\n", names[variant], t[0], t[1], (t[1] - t[0]) / K);
    }
    // the same question when every kernel also writes a word into MAPPED HOST memory (the header
    // mirror of api.cu): plain / pdl / graph / one-kernel-with-flags, and a memset between kernels
    int *hsink; CK(cudaHostAlloc(&hsink, 4096, cudaHostAllocMapped)); *hsink = 0;
    int *hsink_dev; CK(cudaHostGetDevicePointer(&hsink_dev, hsink, 0));
    cudaGraphExec_t gexec2 = nullptr;
    {
        cudaGraph_t graph;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < K; i++) k_plain<<<BLOCKS, THREADS, 0, s>>>(cycles, hsink_dev);
        CK(cudaStreamEndCapture(s, &graph));
        CK(cudaGraphInstantiate(&gexec2, graph, 0));
    }
    auto chain2 = [&](int variant) {
        if (variant == 2) { CK(cudaGraphLaunch(gexec2, s)); return; }
        if (variant == 3) { CK(cudaMemsetAsync(bar, 0, 4, s)); k_flag<<<BLOCKS, THREADS, 0, s>>>(cycles, K, hsink_dev, bar); return; }
        for (int i = 0; i < K; i++) {
            if (variant == 4) { CK(cudaMemsetAsync(bar, 0, 4, s)); k_plain<<<BLOCKS, THREADS, 0, s>>>(cycles, sink); continue; }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(BLOCKS); cfg.blockDim = dim3(THREADS); cfg.stream = s;
            cudaLaunchAttribute at[1]; int na = 0;
            if (variant == 1) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; na++; }
            cfg.attrs = at; cfg.numAttrs = na;
            if (variant == 1) CK(cudaLaunchKernelEx(&cfg, k_pdl, cycles, hsink_dev, 0));
            else CK(cudaLaunchKernelEx(&cfg, k_plain, cycles, hsink_dev));
        }
    };
    const char *names2[] = {"plain+hostwr", "pdl+hostwr", "graph+hostwr", "flag+hostwr", "memset+plain"};
    for (int variant = 0; variant < 5; variant++) {
        double t[2] = {0, 0};
        for (int copying = 0; copying < 2; copying++) {
            for (int r = -2; r < REPS; r++) {
                CK(cudaDeviceSynchronize());
                if (copying) CK(cudaMemcpyAsync(hdst, dsrc, NB, cudaMemcpyDeviceToHost, side));
                const double t0 = now_us();
                chain2(variant);
                CK(cudaStreamSynchronize(s));
                const double dt = now_us() - t0;
                if (r >= 0) t[copying] += dt / REPS;
            }
        }
        printf("%-12s alone %8.1f us   with D2H %8.1f us   extra per boundary %6.2f us\n", names2[variant], t[0], t[1], (t[1] - t[0]) / K);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
