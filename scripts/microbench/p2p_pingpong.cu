// Two GPUs of one box, one process: kernel on GPU0 and kernel on GPU1 ping-pong a flag through peer-mapped memory.
// Measures the one-way flag latency for several store/load flavours (sizes the per-colour halo exchange).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_pingpong p2p_pingpong.cu && ./p2p_pingpong
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void st_rel_sys(unsigned long long *p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acq_sys(const unsigned long long *p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rlx_sys(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_rlx_sys(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

// mode 0: release/acquire.sys ; 1: relaxed.sys ; 2: volatile ; 3: fence.sys + relaxed, with a 2 KB payload written first
__global__ void pingpong(unsigned long long *mine, unsigned long long *theirs, double *their_payload, int iters, int who, int mode) {
    for (int i = 1; i <= iters; i++) {
        if (who == 0) {
            if (mode == 3) { for (int k = threadIdx.x; k < 256; k += blockDim.x) their_payload[k] = (double)i; __syncthreads(); if (threadIdx.x == 0) __threadfence_system(); }
            if (threadIdx.x == 0) {
                if (mode == 0) st_rel_sys(theirs, i); else if (mode == 2) *(volatile unsigned long long *)theirs = i; else st_rlx_sys(theirs, i);
                if (mode == 0) while (ld_acq_sys(mine) < (unsigned long long)i) {} else if (mode == 2) while (*(volatile unsigned long long *)mine < (unsigned long long)i) {} else while (ld_rlx_sys(mine) < (unsigned long long)i) {}
            }
            __syncthreads();
        } else {
            if (threadIdx.x == 0) {
                if (mode == 0) while (ld_acq_sys(mine) < (unsigned long long)i) {} else if (mode == 2) while (*(volatile unsigned long long *)mine < (unsigned long long)i) {} else while (ld_rlx_sys(mine) < (unsigned long long)i) {}
            }
            __syncthreads();
            if (mode == 3) { for (int k = threadIdx.x; k < 256; k += blockDim.x) their_payload[k] = (double)i; __syncthreads(); if (threadIdx.x == 0) __threadfence_system(); }
            if (threadIdx.x == 0) {
                if (mode == 0) st_rel_sys(theirs, i); else if (mode == 2) *(volatile unsigned long long *)theirs = i; else st_rlx_sys(theirs, i);
            }
            __syncthreads();
        }
    }
}

int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    unsigned long long *f[2]; double *pay[2]; cudaStream_t st[2]; cudaEvent_t e0, e1;
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&f[d], 64)); CK(cudaMalloc(&pay[d], 4096)); CK(cudaStreamCreate(&st[d]));
    }
    CK(cudaSetDevice(0)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    const char *names[4] = {"st.release.sys / ld.acquire.sys", "st.relaxed.sys / ld.relaxed.sys", "volatile", "2 KB payload + __threadfence_system + relaxed flag"};
    for (int mode = 0; mode < 4; mode++) {
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaMemset(f[d], 0, 64)); CK(cudaDeviceSynchronize()); }
        CK(cudaSetDevice(0)); CK(cudaEventRecord(e0, st[0]));
        pingpong<<<1, 256, 0, st[0]>>>(f[0], f[1], pay[1], iters, 0, mode);
        CK(cudaEventRecord(e1, st[0]));
        CK(cudaSetDevice(1));
        pingpong<<<1, 256, 0, st[1]>>>(f[1], f[0], pay[0], iters, 1, mode);
        CK(cudaSetDevice(0)); CK(cudaEventSynchronize(e1));
        CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-55s : round trip %.2f us  (one way %.2f us)\n", names[mode], ms * 1e3 / iters, ms * 1e3 / iters / 2);
    }
    return 0;
}
