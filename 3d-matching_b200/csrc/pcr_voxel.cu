// pcr_voxel.cu — voxel-grid down-sampling (K1), replaces pcd.voxel_down_sample(voxel) at
// src/ply/ply.py:106 (origin = min_bound - voxel/2, index = floor((p - origin)/voxel),
// one averaged point per voxel).  Determinism: output order = ascending linear voxel id (D3); per-voxel sums
// are int64 fixed point (D5) so the atomic accumulation order cannot change a bit.
//
// HBM roofline: 16 n read (x2: count pass + accumulate pass) + 16 m write + 4 (cells) table traffic.
#include <cub/device/device_radix_sort.cuh>

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
// ---- sparse path: voxel grids beyond the dense-table budget (a 200-unit scene at the reference's default voxel 0.3, a lidar
// sweep at 5 cm).  Open3D hashes the voxel index; here the 64-bit linear voxel id of every point is radix-sorted
// (cub::DeviceRadixSort — library code on a fallback path; the dense path above is the hot one), heads of equal-key runs are
// ranked with the scan the dense path uses, and the same int64 fixed-point sums (D5) are accumulated per rank.  Output order
// = ascending voxel id (D3) and sums are order-free, so the result equals the dense path's and the oracle's bit for bit.
struct VoxDims64 {
    double ox, oy, oz, voxel;
    long long nx, ny, nz;
};

__global__ void __launch_bounds__(256) k_vox_keys(const float4 *__restrict__ pts, int n, VoxDims64 g, u64 *__restrict__ keys,
                                                  uint32_t *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(pts + i);
    long long cx = (long long)floor(((double)p.x - g.ox) / g.voxel);
    long long cy = (long long)floor(((double)p.y - g.oy) / g.voxel);
    long long cz = (long long)floor(((double)p.z - g.oz) / g.voxel);
    cx = min(max(cx, 0LL), g.nx - 1);
    cy = min(max(cy, 0LL), g.ny - 1);
    cz = min(max(cz, 0LL), g.nz - 1);
    keys[i] = (u64)((cz * g.ny + cy) * g.nx + cx);
    idx[i] = (uint32_t)i;
}

// flags[i] = 1 where a run of equal keys starts (exclusive scan -> rank of the run); flags[n] = 0 receives the total
__global__ void __launch_bounds__(256) k_vox_heads(const u64 *__restrict__ keys, int n, uint32_t *__restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    flags[i] = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
}

// sorted position i belongs to run rank[i + 1] - 1 (inclusive count of heads up to i)
__global__ void __launch_bounds__(256) k_vox_accum_sorted(const float4 *__restrict__ pts, int n, const uint32_t *__restrict__ idx,
                                                          const uint32_t *__restrict__ rank, const u64 *__restrict__ keys,
                                                          double scale, long long *__restrict__ sums, uint32_t *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = i == 0 || keys[i] != keys[i - 1];
    const uint32_t r = rank[i] - (head ? 0u : 1u);  // rank[] = exclusive scan of the head flags
    const float4 p = __ldg(pts + idx[i]);
    atomicAdd((u64 *)(sums + 3 * (size_t)r + 0), (u64)fixed_ll((double)p.x, scale));
    atomicAdd((u64 *)(sums + 3 * (size_t)r + 1), (u64)fixed_ll((double)p.y, scale));
    atomicAdd((u64 *)(sums + 3 * (size_t)r + 2), (u64)fixed_ll((double)p.z, scale));
    atomicAdd(cnt + r, 1u);
}

__global__ void __launch_bounds__(256) k_vox_finalize_sorted(const uint32_t *__restrict__ rank, int n, const long long *__restrict__ sums,
                                                             const uint32_t *__restrict__ cnt, double inv_scale, float4 *__restrict__ out,
                                                             u64 *__restrict__ total) {
    const uint32_t m = rank[n];
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) *total = ((u64)m) << 32;
    if (r >= m) return;
    const double k = (double)cnt[r];
    out[r] = make_float4((float)(((double)sums[3 * (size_t)r + 0] * inv_scale) / k), (float)(((double)sums[3 * (size_t)r + 1] * inv_scale) / k),
                         (float)(((double)sums[3 * (size_t)r + 2] * inv_scale) / k), 0.0f);
}

static int voxel_enqueue_sparse(pcr_ctx *ctx, const float4 *pts, int n, const VoxDims64 &g, double cells, int k, float4 *out, u64 *h_total) {
    PCR_ALLOC(keys, u64, 2 * (size_t)n);
    PCR_ALLOC(idx, uint32_t, 2 * (size_t)n);
    PCR_ALLOC(flags, uint32_t, (size_t)n + 1);
    PCR_ALLOC(sums, long long, 3 * (size_t)n);
    PCR_ALLOC(cnt, uint32_t, (size_t)n);
    PCR_ALLOC(total, u64, 1);
    int end_bit = 1;
    while (end_bit < 64 && ldexp(1.0, end_bit) < cells) end_bit++;
    size_t tmp_bytes = 0;
    PCR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys + n, idx, idx + n, n, 0, end_bit, ctx->stream));
    unsigned char *tmp = arena<unsigned char>(ctx, tmp_bytes ? tmp_bytes : 1);
    if (!tmp) return PCR_ERR_OOM;
    KScope ks(ctx, KC_VOXEL, 16.0 * n + 8.0 * 12.0 * n + 40.0 * n, 5);
    k_vox_keys<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, g, keys, idx);
    PCR_LAUNCHED();
    PCR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys + n, idx, idx + n, n, 0, end_bit, ctx->stream));
    k_vox_heads<<<div_up(n + 1, 256), 256, 0, ctx->stream>>>(keys + n, n, flags);
    PCR_LAUNCHED();
    PCR_TRY(pcr_exclusive_scan_u32(ctx, flags, n));
    PCR_CUDA(cudaMemsetAsync(sums, 0, sizeof(long long) * 3 * (size_t)n, ctx->stream));
    PCR_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (size_t)n, ctx->stream));
    k_vox_accum_sorted<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, idx + n, flags, keys + n, ldexp(1.0, k), sums, cnt);
    PCR_LAUNCHED();
    k_vox_finalize_sorted<<<div_up(n, 256), 256, 0, ctx->stream>>>(flags, n, sums, cnt, ldexp(1.0, -k), out, total);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    PCR_CUDA(cudaMemcpyAsync(h_total, total, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    return PCR_OK;
}

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
    const int E = amax > 0.0f ? pcr_pow2ceil_exp((double)amax) : 0;
    const int k = 62 - E - pcr_ilog2ceil(n > 1 ? n : 1);
    const double cells = (double)dims[0] * (double)dims[1] * (double)dims[2];
    if (cells > (double)PCR_MAX_GRID_CELLS) {
        // as the oracle: a dimension beyond int32 or a grid beyond 2^63 ids is an error, everything else is down-sampled
        if (dims[0] > 2147483647LL || dims[1] > 2147483647LL || dims[2] > 2147483647LL || cells > 9.0e18)
            return pcr_fail(ctx, PCR_ERR_TOO_LARGE, "voxel grid %lld x %lld x %lld: voxel_size is too small for this extent", dims[0], dims[1],
                            dims[2]);
        VoxDims64 g64{org[0], org[1], org[2], voxel, dims[0], dims[1], dims[2]};
        return voxel_enqueue_sparse(ctx, pts, n, g64, cells, k, out, h_total);
    }
    g.ox = org[0]; g.oy = org[1]; g.oz = org[2];
    g.nx = (int)dims[0]; g.ny = (int)dims[1]; g.nz = (int)dims[2];
    const long long ncells = dims[0] * dims[1] * dims[2];
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
    PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
    *m_host = (int)(*h_total >> 32);
    return PCR_OK;
}
