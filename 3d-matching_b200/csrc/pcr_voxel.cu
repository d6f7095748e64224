// pcr_voxel.cu — voxel-grid down-sampling (K1), replaces pcd.voxel_down_sample(voxel) at
// src/ply/ply.py:106 (origin = min_bound - voxel/2, index = floor((p - origin)/voxel),
// one averaged point per voxel).  Determinism: output order = ascending linear voxel id (D3); per-voxel sums
// are int64 fixed point (D5) so the atomic accumulation order cannot change a bit.
//
// HBM roofline: 16 n read (x2: count pass + accumulate pass) + 16 m write + 4 (cells) table traffic.
#include "pcr_common.cuh"

typedef unsigned long long u64;

struct VoxDims {
    double ox, oy, oz, voxel;
    int nx, ny, nz;
};

__device__ __forceinline__ uint32_t vox_cell(const VoxDims &g, const float4 &p) {
    // fp64 division, as the reference's double arithmetic (A.1); the oracle evaluates the same expression
    int cx = (int)floor(((double)p.x - g.ox) / g.voxel);
    int cy = (int)floor(((double)p.y - g.oy) / g.voxel);
    int cz = (int)floor(((double)p.z - g.oz) / g.voxel);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    cz = min(max(cz, 0), g.nz - 1);
    return (uint32_t)((cz * g.ny + cy) * g.nx + cx);
}

__global__ void __launch_bounds__(256) k_vox_count(const float4 *__restrict__ pts, int n, VoxDims g,
                                                   uint32_t *__restrict__ cell, u64 *__restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = vox_cell(g, __ldg(pts + i));
    cell[i] = c;
    // low 32 bits: point count; bit 32 set once by the first arrival -> after the scan the high word is the
    // rank of the voxel among occupied voxels (ascending id)
    const u64 old = atomicAdd(packed + c, 1ull);
    if ((old & 0xffffffffull) == 0ull) atomicAdd(packed + c, 1ull << 32);
}

__global__ void __launch_bounds__(256) k_vox_accum(const float4 *__restrict__ pts, int n,
                                                   const uint32_t *__restrict__ cell, const u64 *__restrict__ scanned,
                                                   double scale, long long *__restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(pts + i);
    const uint32_t r = (uint32_t)(scanned[cell[i]] >> 32);
    atomicAdd((u64 *)(sums + 3 * (size_t)r + 0), (u64)fixed_ll((double)p.x, scale));
    atomicAdd((u64 *)(sums + 3 * (size_t)r + 1), (u64)fixed_ll((double)p.y, scale));
    atomicAdd((u64 *)(sums + 3 * (size_t)r + 2), (u64)fixed_ll((double)p.z, scale));
}

__global__ void __launch_bounds__(256) k_vox_finalize(const u64 *__restrict__ scanned, long long ncells,
                                                      const long long *__restrict__ sums, double inv_scale,
                                                      float4 *__restrict__ out) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (long long)gridDim.x * blockDim.x) {
        const u64 a = scanned[c], b = scanned[c + 1];
        const uint32_t cnt = (uint32_t)(b & 0xffffffffull) - (uint32_t)(a & 0xffffffffull);
        if (cnt == 0) continue;
        const uint32_t r = (uint32_t)(a >> 32);
        const double k = (double)cnt;
        out[r] = make_float4((float)(((double)sums[3 * (size_t)r + 0] * inv_scale) / k),
                             (float)(((double)sums[3 * (size_t)r + 1] * inv_scale) / k),
                             (float)(((double)sums[3 * (size_t)r + 2] * inv_scale) / k), 0.0f);
    }
}

// Enqueues the whole down-sampling of one cloud on ctx->stream WITHOUT a host synchronisation: the (rank << 32 | count)
// total of the scan is copied to *h_total (pinned, the caller's slot); once the stream has been synchronised the number of
// occupied voxels is *h_total >> 32.  pcr_align enqueues its two clouds on two streams and waits once.
int pcr_voxel_enqueue(pcr_ctx *ctx, const float4 *pts, int n, double voxel, float4 *out, u64 *h_total) {
    *h_total = 0;
    if (!(voxel > 0.0)) return pcr_fail(ctx, PCR_ERR_INVALID, "voxel_size must be > 0");
    if (n == 0) return PCR_OK;
    float lo[3], hi[3];
    PCR_TRY(pcr_bounds(ctx, pts, n, lo, hi));
    VoxDims g;
    g.voxel = voxel;
    double org[3];
    long long dims[3];
    float amax = 0.0f;
    for (int d = 0; d < 3; d++) {
        if (!(lo[d] <= hi[d]) || isinf(lo[d]) || isinf(hi[d]))
            return pcr_fail(ctx, PCR_ERR_INVALID, "voxel_down_sample: non-finite coordinates");
        org[d] = (double)lo[d] - voxel * 0.5;
        dims[d] = (long long)floor(((double)hi[d] - org[d]) / voxel) + 1;
        amax = fmaxf(amax, fmaxf(fabsf(lo[d]), fabsf(hi[d])));
    }
    if ((double)dims[0] * (double)dims[1] * (double)dims[2] > (double)PCR_MAX_GRID_CELLS)
        return pcr_fail(ctx, PCR_ERR_TOO_LARGE, "voxel grid %lld x %lld x %lld exceeds the dense-grid budget; voxel_size is too small",
                        dims[0], dims[1], dims[2]);
    g.ox = org[0]; g.oy = org[1]; g.oz = org[2];
    g.nx = (int)dims[0]; g.ny = (int)dims[1]; g.nz = (int)dims[2];
    const long long ncells = dims[0] * dims[1] * dims[2];
    const int E = amax > 0.0f ? pcr_pow2ceil_exp((double)amax) : 0;
    const int k = 62 - E - pcr_ilog2ceil(n > 1 ? n : 1);
    PCR_ALLOC(cell, uint32_t, (size_t)n);
    PCR_ALLOC(packed, u64, (size_t)ncells + 1);
    PCR_ALLOC(sums, long long, 3 * (size_t)n);
    KScope *ks = new KScope(ctx, KC_VOXEL, 40.0 * n + 24.0 * (double)ncells + 16.0 * n);
    PCR_CUDA(cudaMemsetAsync(packed, 0, sizeof(u64) * ((size_t)ncells + 1), ctx->stream));
    k_vox_count<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, g, cell, packed);
    PCR_LAUNCHED();
    PCR_TRY(pcr_exclusive_scan_u64(ctx, packed, ncells));
    PCR_CUDA(cudaMemcpyAsync(h_total, packed + ncells, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaMemsetAsync(sums, 0, sizeof(long long) * 3 * (size_t)n, ctx->stream));
    k_vox_accum<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, cell, packed, ldexp(1.0, k), sums);
    PCR_LAUNCHED();
    const int blocks = (int)min((long long)div_up(ncells, 256), (long long)ctx->sm_count * 16);
    k_vox_finalize<<<blocks, 256, 0, ctx->stream>>>(packed, ncells, sums, ldexp(1.0, -k), out);
    PCR_LAUNCHED();
    delete ks;
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

int pcr_voxel_impl(pcr_ctx *ctx, const float4 *pts, int n, double voxel, float4 *out, int *m_host) {
    *m_host = 0;
    u64 *h_total = (u64 *)ctx->pinned;
    PCR_TRY(pcr_voxel_enqueue(ctx, pts, n, voxel, out, h_total));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    *m_host = (int)(*h_total >> 32);
    return PCR_OK;
}
