// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu   (run on a B200: ./barrier_bench)
// Micro-benchmark: grid-barrier variants for the persistent ICP kernel (cycles per barrier incl. 29 data atomics).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel(unsigned *p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_rlx(unsigned *p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

template <int V>
__global__ void k(unsigned *bar, unsigned long long *acc, long long *out, int iters, int data) {
    long long t_tot = 0;
    for (int it = 0; it < iters; it++) {
        __syncthreads();
        const long long c0 = clock64();
        unsigned long long *a = acc + (it % 3) * 32;
        if (data && threadIdx.x < 29) atomicAdd(a + threadIdx.x, (unsigned long long)(threadIdx.x + 1));
        const unsigned target = (unsigned)(it + 1) * gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (V == 0) { __threadfence(); atomicAdd(bar, 1u); while (ld_acq(bar) < target) {} __threadfence(); }
            if (V == 1) { red_rel(bar, 1u); while (ld_acq(bar) < target) {} }
            if (V == 2) { __threadfence(); red_rlx(bar, 1u); while (ld_rlx(bar) < target) {} __threadfence(); }
            if (V == 3) { red_rel(bar, 1u); while (ld_rlx(bar) < target) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
            if (V == 4) { __threadfence(); red_rlx(bar, 1u); while (*(volatile unsigned *)bar < target) {} }
        }
        __syncthreads();
        unsigned long long v = 0;
        if (threadIdx.x < 29) v = __ldcg(a + threadIdx.x);
        if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 61) acc[((it + 2) % 3) * 32 + threadIdx.x - 32] = 0;
        __syncthreads();
        const long long c1 = clock64();
        t_tot += c1 - c0;
        if (threadIdx.x < 29 && data && v != (unsigned long long)(threadIdx.x + 1) * gridDim.x) out[1] = it + 1;  // check
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t_tot / iters;
}
template <int V> void run(int blocks, int data) {
    unsigned *bar; unsigned long long *acc; long long *out;
    cudaMalloc(&bar, 4); cudaMalloc(&acc, 96 * 8); cudaMalloc(&out, 16);
    cudaMemset(bar, 0, 4); cudaMemset(acc, 0, 96 * 8); cudaMemset(out, 0, 16);
    int iters = 200;
    void *args[] = {&bar, &acc, &out, &iters, &data};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k<V>, dim3(blocks), dim3(256), args, 0, 0);
    cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("variant %d blocks %4d data %d: %lld cycles/barrier  (err %s, bad %lld)\n", V, blocks, data, h[0], cudaGetErrorString(e), h[1]);
    cudaFree(bar); cudaFree(acc); cudaFree(out);
}
int main() {
    for (int data = 0; data <= 1; data++)
        for (int b : {8, 148, 444}) { run<0>(b, data); run<1>(b, data); run<2>(b, data); run<3>(b, data); run<4>(b, data); }
    return 0;
}
