// Micro-benchmarks that size the two synchronisation-bound kernels (sync-free triangular solve, persistent sweep):
//   A. flag-chain hop latency: thread i waits for x[i-1] (NaN-payload sentinel) and publishes x[i]
//   B. grid barrier latency over a co-resident grid (atomic arrive + acquire polling), with / without __threadfence
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_latency sync_latency.cu && ./sync_latency
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#define SENT 0xFFF8DEADBEEF0001ull
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long ldr(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned int lda(const unsigned int *p) { unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// hop i is done by lane 0 of warp (i % warps_total); each warp handles hops i, i+W, i+2W ... ; stride spreads consecutive
// hops over CTAs: warp index = (i * stride) % W
__global__ void chain_kernel(unsigned long long *x, int hops, int W, int stride) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (threadIdx.x & 31) return;
    for (int i = 0; i < hops; i++) {
        if ((int)(((long long)i * stride) % W) != w) continue;
        unsigned long long prev = 0;
        if (i > 0) { do { prev = ldr(x + (size_t)(i - 1) * 16); } while (prev == SENT); }
        str(x + (size_t)i * 16, prev + 1);
    }
}

// many pollers variant: every thread of the grid polls its own parent chain concurrently: `lanes` independent chains
__global__ void chains_kernel(unsigned long long *x, int hops, int W, int stride, int lanes) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (lane >= lanes) return;
    for (int i = 0; i < hops; i++) {
        if ((int)(((long long)i * stride) % W) != w) continue;
        unsigned long long prev = 0;
        unsigned long long *row = x + ((size_t)lane * hops) * 16;
        if (i > 0) { do { prev = ldr(row + (size_t)(i - 1) * 16); } while (prev == SENT); }
        str(row + (size_t)i * 16, prev + 1);
    }
}

template <bool FENCE>
__global__ void barrier_kernel(unsigned int *bar, int iters, double *sink) {
    unsigned int arrivals = 0;
    double acc = 0;
    for (int it = 0; it < iters; it++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (FENCE) __threadfence();
            atomicAdd(bar, 1u);
            arrivals += gridDim.x;
            while (lda(bar) < arrivals) { }
            if (FENCE) __threadfence();
        }
        __syncthreads();
        acc += 1.0;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int hops = 20000;
    unsigned long long *x; CK(cudaMalloc(&x, (size_t)hops * 16 * 8 * 32));
    std::vector<unsigned long long> init((size_t)hops * 16 * 32, SENT);
    float ms;
    struct Cfg { int ctas, threads, stride; } cfgs[] = {{1, 64, 1}, {1, 256, 1}, {148, 32, 1}, {148, 256, 1}, {296, 256, 1}, {148, 256, 7}, {592, 256, 13}};
    for (auto c : cfgs) {
        CK(cudaMemcpy(x, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
        int W = c.ctas * c.threads / 32;
        CK(cudaEventRecord(e0));
        chain_kernel<<<c.ctas, c.threads>>>(x, hops, W, c.stride);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("chain: ctas=%4d threads=%3d stride=%2d : %.3f us/hop\n", c.ctas, c.threads, c.stride, ms * 1e3 / hops);
    }
    for (int lanes : {1, 8, 32}) {
        CK(cudaMemcpy(x, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
        int ctas = 296, threads = 256, W = ctas * threads / 32;
        CK(cudaEventRecord(e0));
        chains_kernel<<<ctas, threads>>>(x, hops, W, 7, lanes);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("chains: %2d independent chains per warp, 296x256 : %.3f us/hop\n", lanes, ms * 1e3 / hops);
    }
    unsigned int *bar; double *sink; CK(cudaMalloc(&bar, 4)); CK(cudaMalloc(&sink, 8));
    for (int per_sm : {1, 2, 4}) {
        for (int fence = 0; fence < 2; fence++) {
            int grid = p.multiProcessorCount * per_sm, iters = 2000;
            CK(cudaMemset(bar, 0, 4));
            void *args[] = {&bar, &iters, &sink};
            CK(cudaEventRecord(e0));
            if (fence) CK(cudaLaunchCooperativeKernel((void *)barrier_kernel<true>, dim3(grid), dim3(256), args, 0, 0));
            else CK(cudaLaunchCooperativeKernel((void *)barrier_kernel<false>, dim3(grid), dim3(256), args, 0, 0));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("grid barrier: %4d CTAs x256, fence=%d : %.3f us/barrier\n", grid, fence, ms * 1e3 / iters);
        }
    }
    return 0;
}
