// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench2 barrier_bench2.cu   (run on a B200)
// Micro-benchmark, round 2: end-of-pass exchange of the persistent ICP kernel = 29 64-bit sums from every CTA + one grid
// barrier + every CTA reading the 29 totals.  Variants of WHERE the sums and the arrivals go:
//   CL      cluster size (1: none; 4 / 8: sums folded over the cluster through distributed shared memory, one global
//           arrival and one poller per cluster, hardware cluster barriers around them)
//   STRIDE  distance between two sums in 8-byte words (1: two 128-byte lines as today; 32: one 256-byte block per sum)
//   K       replicas of the accumulator (CTA or cluster b adds into replica b % K; readers add the K copies)
//   FLAGS   0: everybody polls the arrival counter; F > 0: the last arriver (atom with return) raises F flag lines, CTA b polls
//           flag b % F
// Bounded spins: a variant that would hang reports "TIMEOUT" instead.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel(unsigned *p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_acqrel(unsigned *p, unsigned v) { unsigned o; asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void st_rel(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

constexpr int NS = 29;
constexpr long long SPIN_MAX = 1ll << 22;

template <int CL, int STRIDE, int K, int FLAGS>
__global__ void __launch_bounds__(256) k(unsigned *bar, unsigned *flags, unsigned long long *acc, long long *out, int iters) {
    __shared__ unsigned long long part[32];
    __shared__ unsigned long long tot[32];
    long long t_tot = 0;
    const int bufsz = K * NS * STRIDE;
    unsigned crank = 0;
    if (CL > 1) crank = cg::this_cluster().block_rank();
    const unsigned unit = CL > 1 ? blockIdx.x / CL : blockIdx.x;   // arrival unit: CTA or cluster
    const unsigned nunits = CL > 1 ? gridDim.x / CL : gridDim.x;
    int bad = 0;
    for (int it = 0; it < iters; it++) {
        __syncthreads();
        const long long c0 = clock64();
        unsigned long long *a = acc + (size_t)(it % 3) * bufsz;
        unsigned long long *ak = a + (size_t)(unit % K) * NS * STRIDE;
        if (threadIdx.x < NS) part[threadIdx.x] = (unsigned long long)(threadIdx.x + 1);
        if (CL > 1) {
            cg::this_cluster().sync();
            if (crank == 0 && threadIdx.x < NS) {
                unsigned long long s = 0;
#pragma unroll
                for (int r = 0; r < CL; r++) s += *cg::this_cluster().map_shared_rank(&part[threadIdx.x], r);
                atomicAdd(ak + threadIdx.x * STRIDE, s);
            }
        } else {
            __syncthreads();
            if (threadIdx.x < NS) atomicAdd(ak + threadIdx.x * STRIDE, part[threadIdx.x]);
        }
        __syncthreads();
        if (threadIdx.x == 0 && (CL == 1 || crank == 0)) {
            long long spin = 0;
            if (FLAGS == 0) {
                const unsigned target = (unsigned)(it + 1) * nunits;
                red_rel(bar, 1u);
                while (ld_acq(bar) < target && ++spin < SPIN_MAX) {}
            } else {
                const unsigned old = atom_acqrel(bar, 1u);
                if (old == (unsigned)(it + 1) * nunits - 1) {
#pragma unroll
                    for (int f = 0; f < FLAGS; f++) st_rel(flags + f * 64, (unsigned)(it + 1));
                } else {
                    while (ld_acq(flags + (unit % FLAGS) * 64) < (unsigned)(it + 1) && ++spin < SPIN_MAX) {}
                }
            }
            if (spin >= SPIN_MAX) bad = -1;
        }
        if (CL > 1) cg::this_cluster().sync();
        else __syncthreads();
        // every CTA reads the totals (K replicas each)
        if (threadIdx.x < NS) {
            unsigned long long v = 0;
#pragma unroll
            for (int r = 0; r < K; r++) v += __ldcg(a + (size_t)r * NS * STRIDE + threadIdx.x * STRIDE);
            tot[threadIdx.x] = v;
            if (v != (unsigned long long)(threadIdx.x + 1) * gridDim.x) bad = it + 1;
        }
        if (blockIdx.x == 0) {
            unsigned long long *z = acc + (size_t)((it + 2) % 3) * bufsz;
            for (int i = threadIdx.x; i < K * NS; i += 256) z[(size_t)(i / NS) * NS * STRIDE + (i % NS) * STRIDE] = 0;
        }
        __syncthreads();
        const long long c1 = clock64();
        t_tot += c1 - c0;
    }
    if (bad) out[1] = bad;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t_tot / iters;
}

template <int CL, int STRIDE, int K, int FLAGS> void run(int blocks) {
    unsigned *bar, *flags; unsigned long long *acc; long long *out;
    const size_t accsz = (size_t)3 * K * NS * STRIDE * 8;
    cudaMalloc(&bar, 256); cudaMalloc(&flags, 64 * 4 * 16); cudaMalloc(&acc, accsz); cudaMalloc(&out, 16);
    cudaMemset(bar, 0, 256); cudaMemset(flags, 0, 64 * 4 * 16); cudaMemset(acc, 0, accsz); cudaMemset(out, 0, 16);
    int iters = 200;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; na++;
    if (CL > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; na++; }
    cfg.attrs = at; cfg.numAttrs = na;
    int maxcl = -1;
    if (CL > 1) cudaOccupancyMaxActiveClusters(&maxcl, k<CL, STRIDE, K, FLAGS>, &cfg);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<CL, STRIDE, K, FLAGS>, bar, flags, acc, out, iters);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("CL %d stride %2d K %d flags %d blocks %4d: %6lld cycles  (launch %s, sync %s, bad %lld, max clusters %d)\n", CL, STRIDE, K, FLAGS,
           blocks, h[0], cudaGetErrorString(e), cudaGetErrorString(e2), h[1], maxcl);
    cudaFree(bar); cudaFree(flags); cudaFree(acc); cudaFree(out);
}

int main() {
    for (int b : {148, 392, 592}) {
        run<1, 1, 1, 0>(b);
        run<1, 32, 1, 0>(b);
        run<1, 32, 4, 0>(b);
        run<1, 32, 8, 0>(b);
        run<1, 1, 1, 8>(b);
        run<1, 32, 4, 8>(b);
        run<4, 1, 1, 0>(b);
        run<4, 32, 1, 0>(b);
        run<4, 32, 4, 0>(b);
        run<4, 32, 4, 8>(b);
        run<8, 1, 1, 0>(b);
        run<8, 32, 4, 0>(b);
    }
    return 0;
}
