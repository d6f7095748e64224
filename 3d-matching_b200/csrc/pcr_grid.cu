// pcr_grid.cu — device uniform-grid build (K2): bounds reduction, cell counting, exclusive scan, scatter.
// Replaces the KD-tree builds inside every Open3D call of the reference (KDTreeFlann).
//
// HBM roofline: algorithmic bytes = 16 n (read) + 4 n (cell ids) + 16 n (sorted write) + 4 (ncells+1).
#include <cooperative_groups.h>

#include "pcr_common.cuh"

namespace cg = cooperative_groups;

// Bounding boxes of one or two clouds in ONE launch with no initialisation pass: every block reduces its share to six
// floats, the last block to finish (ticket) folds the per-block results of each cloud and re-arms the ticket.  (Round 1:
// an init kernel + a reduction with six same-address atomics per warp, per cloud, and a host synchronisation for each
// of the four clouds an alignment looks at.)
__global__ void __launch_bounds__(256) k_bounds2(const float4 *__restrict__ pa, int na, int blocks_a, const float4 *__restrict__ pb,
                                                 int nb, float *__restrict__ part, unsigned int *__restrict__ ticket,
                                                 float *__restrict__ out) {
    __shared__ float sm[8][6];
    __shared__ bool last;
    const bool second = (int)blockIdx.x >= blocks_a;
    const float4 *__restrict__ pts = second ? pb : pa;
    const int n = second ? nb : na;
    const int blk = second ? blockIdx.x - blocks_a : blockIdx.x, nblk = second ? gridDim.x - blocks_a : blocks_a;
    float v[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int i = blk * blockDim.x + threadIdx.x; i < n; i += nblk * blockDim.x) {
        const float4 p = __ldg(pts + i);
        v[0] = fminf(v[0], p.x); v[3] = fmaxf(v[3], p.x);
        v[1] = fminf(v[1], p.y); v[4] = fmaxf(v[4], p.y);
        v[2] = fminf(v[2], p.z); v[5] = fmaxf(v[5], p.z);
    }
    auto fold = [&]() {  // CTA-wide min / max of v[] into thread 0
#pragma unroll
        for (int d = 0; d < 6; d++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float t = __shfl_xor_sync(0xffffffffu, v[d], o);
                v[d] = d < 3 ? fminf(v[d], t) : fmaxf(v[d], t);
            }
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int d = 0; d < 6; d++) sm[threadIdx.x >> 5][d] = v[d];
        __syncthreads();
        if (threadIdx.x < 32) {
#pragma unroll
            for (int d = 0; d < 6; d++) {
                float t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x][d] : (d < 3 ? INFINITY : -INFINITY);
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    const float u = __shfl_xor_sync(0xffffffffu, t, o);
                    t = d < 3 ? fminf(t, u) : fmaxf(t, u);
                }
                v[d] = t;
            }
        }
        __syncthreads();
    };
    fold();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < 6; d++) part[(size_t)blockIdx.x * 6 + d] = v[d];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int cloud = 0; cloud < 2; cloud++) {
        const int b0 = cloud ? blocks_a : 0, b1 = cloud ? (int)gridDim.x : blocks_a;
#pragma unroll
        for (int d = 0; d < 6; d++) v[d] = d < 3 ? INFINITY : -INFINITY;
        for (int b = b0 + threadIdx.x; b < b1; b += blockDim.x)
#pragma unroll
            for (int d = 0; d < 6; d++) {
                const float t = __ldcg(part + (size_t)b * 6 + d);
                v[d] = d < 3 ? fminf(v[d], t) : fmaxf(v[d], t);
            }
        fold();
        if (threadIdx.x == 0)
#pragma unroll
            for (int d = 0; d < 6; d++) out[cloud * 6 + d] = v[d];
    }
    if (threadIdx.x == 0) *ticket = 0u;  // re-armed for the next launch on this context's stream
}

static bool bounds_cached(pcr_ctx *ctx, const float4 *pts, int n, float lo[3], float hi[3]) {
    for (const auto &e : ctx->bounds_cache)
        if (e.ptr == (const void *)pts && e.n == n) {  // same buffer earlier in this call: no second reduction + sync
            if (lo)
                for (int d = 0; d < 3; d++) { lo[d] = e.lo[d]; hi[d] = e.hi[d]; }
            return true;
        }
    return false;
}

// bounding boxes of two clouds (the second may be empty: nb = 0) with one launch and one host synchronisation; both go
// into the context's per-call cache
int pcr_bounds_pair(pcr_ctx *ctx, const float4 *pa, int na, const float4 *pb, int nb) {
    const bool need_a = na > 0 && !bounds_cached(ctx, pa, na, nullptr, nullptr);
    const bool need_b = nb > 0 && !bounds_cached(ctx, pb, nb, nullptr, nullptr);
    if (!need_a && !need_b) return PCR_OK;
    if (!need_a) { pa = pb; na = nb; pb = nullptr; nb = 0; }
    else if (!need_b) { pb = nullptr; nb = 0; }
    if (!ctx->bounds_ticket) {
        PCR_CUDA(cudaMalloc(&ctx->bounds_ticket, sizeof(unsigned int)));
        PCR_CUDA(cudaMemsetAsync(ctx->bounds_ticket, 0, sizeof(unsigned int), ctx->stream));
    }
    const int blocks_a = min(div_up(na, 256), ctx->sm_count * 4);
    const int blocks_b = nb > 0 ? min(div_up(nb, 256), ctx->sm_count * 4) : 0;
    PCR_ALLOC(part, float, (size_t)(blocks_a + blocks_b) * 6 + 12);
    float *out = part + (size_t)(blocks_a + blocks_b) * 6;
    {
        KScope ks(ctx, KC_BOUNDS, 16.0 * ((double)na + (double)nb));
        k_bounds2<<<blocks_a + blocks_b, 256, 0, ctx->stream>>>(pa, na, blocks_a, pb, nb, part, ctx->bounds_ticket, out);
        PCR_LAUNCHED();
    }
    float *hb = (float *)ctx->pinned;
    PCR_CUDA(cudaMemcpyAsync(hb, out, 12 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
    for (int cloud = 0; cloud < (nb > 0 ? 2 : 1); cloud++) {
        pcr_ctx::BoundsEntry e;
        e.ptr = cloud ? (const void *)pb : (const void *)pa;
        e.n = cloud ? nb : na;
        for (int d = 0; d < 3; d++) {
            e.lo[d] = hb[cloud * 6 + d];
            e.hi[d] = hb[cloud * 6 + 3 + d];
        }
        ctx->bounds_cache.push_back(e);
    }
    return PCR_OK;
}

int pcr_bounds(pcr_ctx *ctx, const float4 *pts, int n, float lo[3], float hi[3]) {
    if (bounds_cached(ctx, pts, n, lo, hi)) return PCR_OK;
    PCR_TRY(pcr_bounds_pair(ctx, pts, n, nullptr, 0));
    if (!bounds_cached(ctx, pts, n, lo, hi)) return pcr_fail(ctx, PCR_ERR_INVALID, "bounds: empty cloud");
    return PCR_OK;
}

// ---- single-pass exclusive scan (decoupled look-back) ----------------------------------------------------------------
// One kernel instead of three (tile sums / scan of sums / tiles): a tile takes a ticket (so every earlier tile is
// already running), publishes its aggregate, and warp 0 walks back over the 32 nearest predecessors at a time until
// it meets an inclusive prefix.  state[t] = (flag << 62) | value, flag 1 = aggregate, 2 = inclusive prefix; values stay
// below 2^62 (they are point counts).  state and the ticket are zeroed before the launch.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <typename T>
__global__ void __launch_bounds__(1024) k_scan_onepass(T *__restrict__ data, long long n, unsigned long long *__restrict__ state,
                                                       unsigned int *__restrict__ ticket) {
    constexpr int ITEMS = 4;
    typedef unsigned long long u64s;
    __shared__ T ws[32];
    __shared__ unsigned int s_tile;
    __shared__ T s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const long long tile = s_tile;
    const long long base = tile * (1024 * ITEMS) + (long long)threadIdx.x * ITEMS;
    T v[ITEMS];
    T s = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        v[k] = (base + k < n) ? data[base + k] : (T)0;
        s += v[k];
    }
    T x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        T w = ws[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T y = __shfl_up_sync(0xffffffffu, w, o);
            if (threadIdx.x >= o) w += y;
        }
        ws[threadIdx.x] = w;  // inclusive over warps; ws[31] = tile aggregate
        const T aggregate = __shfl_sync(0xffffffffu, w, 31);
        const int lane = threadIdx.x;
        T prefix = 0;
        if (tile == 0) {
            if (lane == 0) {
                __threadfence();
                atomicExch(state + tile, (2ull << 62) | (u64s)aggregate);
            }
        } else {
            if (lane == 0) {
                __threadfence();
                atomicExch(state + tile, (1ull << 62) | (u64s)aggregate);
            }
            long long j = tile - 1;
            for (;;) {
                const long long idx = j - lane;
                u64s w2;
                for (;;) {  // wait until the 32 predecessors of this window have published something
                    w2 = idx >= 0 ? ld_acquire_u64(state + idx) : (2ull << 62);  // before tile 0: inclusive prefix 0
                    if (__ballot_sync(0xffffffffu, (w2 >> 62) == 0ull) == 0u) break;
                }
                const unsigned int incl = __ballot_sync(0xffffffffu, (w2 >> 62) == 2ull);
                const int first = incl ? (__ffs(incl) - 1) : 32;  // nearest predecessor with an inclusive prefix
                T part = (lane <= first) ? (T)(w2 & ((1ull << 62) - 1ull)) : (T)0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (incl) break;
                j -= 32;
            }
            if (lane == 0) {
                __threadfence();
                atomicExch(state + tile, (2ull << 62) | (u64s)(prefix + aggregate));
            }
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    T run = (x - s) + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : (T)0) + s_prefix;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    // data[n] receives the total: the thread that owns element n - 1 has it
    if (base <= n - 1 && n - 1 < base + ITEMS) data[n] = run;
}

template <typename T>
static int exclusive_scan_impl(pcr_ctx *ctx, T *data, long long n) {
    // data has n+1 slots; data[n] receives the total
    if (n <= 0) {
        if (n == 0) PCR_CUDA(cudaMemsetAsync(data, 0, sizeof(T), ctx->stream));
        return PCR_OK;
    }
    const int tiles = div_up(n, 4096);
    PCR_ALLOC(state, unsigned long long, (size_t)tiles + 1);  // [tiles] flags + values, then the ticket
    PCR_CUDA(cudaMemsetAsync(state, 0, sizeof(unsigned long long) * ((size_t)tiles + 1), ctx->stream));
    k_scan_onepass<T><<<tiles, 1024, 0, ctx->stream>>>(data, n, state, (unsigned int *)(state + tiles));
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

int pcr_exclusive_scan_u32(pcr_ctx *ctx, uint32_t *data, long long n) { return exclusive_scan_impl<uint32_t>(ctx, data, n); }

// Counting-sort tables in ONE allocation cleared by ONE memset: [ncells + 1 counters / starts][scan state: tiles flags + ticket].
// (Round 1 cleared `start`, a second per-cell `fill` array and the scan state separately: three memsets per grid, ~10 grids
// per alignment.)
struct SortTables {
    uint32_t *start;
    unsigned long long *state;
    int tiles;
};
static int sort_tables_alloc(pcr_ctx *ctx, long long ncells, SortTables *t) {
    t->tiles = div_up(ncells, 4096);
    const size_t start_bytes = ((sizeof(uint32_t) * ((size_t)ncells + 1)) + 15) & ~(size_t)15;
    const size_t bytes = start_bytes + sizeof(unsigned long long) * ((size_t)t->tiles + 1);
    unsigned char *base = arena<unsigned char>(ctx, bytes);
    if (!base) return PCR_ERR_OOM;
    t->start = reinterpret_cast<uint32_t *>(base);
    t->state = reinterpret_cast<unsigned long long *>(base + start_bytes);
    PCR_CUDA(cudaMemsetAsync(base, 0, bytes, ctx->stream));
    return PCR_OK;
}
static int sort_tables_scan(pcr_ctx *ctx, const SortTables &t, long long ncells) {
    k_scan_onepass<uint32_t><<<t.tiles, 1024, 0, ctx->stream>>>(t.start, ncells, t.state, (unsigned int *)(t.state + t.tiles));
    PCR_LAUNCHED();
    return PCR_OK;
}
int pcr_exclusive_scan_u64(pcr_ctx *ctx, unsigned long long *data, long long n) {
    return exclusive_scan_impl<unsigned long long>(ctx, data, n);
}

// ---- cell counting and scatter ---------------------------------------------------------------------------
struct GridDims {
    double ox, oy, oz, inv_h;
    int nx, ny, nz;
};

__device__ __forceinline__ uint32_t cell_of(const GridDims &g, const float4 &p) {
    int cx = (int)floor(((double)p.x - g.ox) * g.inv_h);
    int cy = (int)floor(((double)p.y - g.oy) * g.inv_h);
    int cz = (int)floor(((double)p.z - g.oz) * g.inv_h);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    cz = min(max(cz, 0), g.nz - 1);
    return (uint32_t)((cz * g.ny + cy) * g.nx + cx);
}

// cell[i] = cell of point i, rank[i] = its arrival rank inside the cell (the value the counter had)
__global__ void __launch_bounds__(256) k_cell_count(const float4 *__restrict__ pts, int n, GridDims g,
                                                    uint32_t *__restrict__ cell, uint32_t *__restrict__ rank,
                                                    uint32_t *__restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_of(g, __ldg(pts + i));
    cell[i] = c;
    rank[i] = atomicAdd(count + c, 1u);
}

// pos = start[c] + rank: the slot order inside a cell follows the atomic arrival order of the counting pass, which no
// result depends on (ties are broken by original index, sums are fixed point).
__global__ void __launch_bounds__(256) k_cell_scatter(const float4 *__restrict__ pts, int n,
                                                      const uint32_t *__restrict__ cell, const uint32_t *__restrict__ rank,
                                                      const uint32_t *__restrict__ start, float4 *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell[i];
    const uint32_t pos = start[c] + rank[i];
    float4 p = __ldg(pts + i);
    p.w = __int_as_float(i);
    sorted[pos] = p;
}

// ---- small clouds: the whole counting sort in ONE launch of one thread-block cluster ------------------------------------
// The down-sampled clouds (~9k points) get six search structures per alignment on the critical path (normals grid, FPFH
// grid per cloud; RANSAC grid; Morton order), each a memset + three tiny kernels whose cost was launch latency, not work
// (~35 us per build, 0.2 ms per alignment).  Eight CTAs of 1024 threads on one GPC do all four phases between hardware
// cluster barriers (~0.2 us each): clear the table, count (cell id and arrival rank stay in registers), exclusive scan
// (per-thread runs of cells, CTA scan, the eight CTA totals exchanged through distributed shared memory), scatter.
constexpr int GC_CTAS = 8, GC_THREADS = 1024, GC_NT = GC_CTAS * GC_THREADS, GC_PPT = 4;
constexpr int GC_MAX_POINTS = GC_NT * GC_PPT;          // 32,768
constexpr long long GC_MAX_CELLS = (1LL << 20);        // 4 MB table: 128 cells per thread in the scan

struct MortonDims {
    double ox, oy, oz, inv_c;
    int lim;  // 2^L - 1
};

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 8 bits -> every third bit
    v &= 0xffu;
    v = (v | (v << 8)) & 0x00f00fu;
    v = (v | (v << 4)) & 0x0c30c3u;
    v = (v | (v << 2)) & 0x249249u;
    return v;
}

__device__ __forceinline__ uint32_t morton_of(const MortonDims &g, const float4 &p) {
    const int cx = min(max((int)floor(((double)p.x - g.ox) * g.inv_c), 0), g.lim);
    const int cy = min(max((int)floor(((double)p.y - g.oy) * g.inv_c), 0), g.lim);
    const int cz = min(max((int)floor(((double)p.z - g.oz) * g.inv_c), 0), g.lim);
    return spread3((uint32_t)cx) | (spread3((uint32_t)cy) << 1) | (spread3((uint32_t)cz) << 2);
}

__device__ __forceinline__ uint32_t gc_cell(const GridDims &g, const float4 &p) { return cell_of(g, p); }
__device__ __forceinline__ uint32_t gc_cell(const MortonDims &g, const float4 &p) { return morton_of(g, p); }

// start has ncells + 1 entries padded to a multiple of 4 (the pad is written, never read)
template <typename Dims>
__global__ void __cluster_dims__(GC_CTAS, 1, 1) __launch_bounds__(GC_THREADS) k_grid_cluster(const float4 *__restrict__ pts, int n, Dims g,
                                                                                           long long ncells, uint32_t *__restrict__ start,
                                                                                           float4 *__restrict__ sorted) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t s_total;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int t = rank * GC_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long len = (ncells + 1 + 3) & ~3LL;  // entries, a multiple of 4
    // phase 0: clear
    for (long long k = 4LL * t; k < len; k += 4LL * GC_NT) *reinterpret_cast<uint4 *>(start + k) = make_uint4(0u, 0u, 0u, 0u);
    cluster.sync();
    // phase 1: count
    uint32_t c[GC_PPT], r[GC_PPT];
#pragma unroll
    for (int k = 0; k < GC_PPT; k++) {
        const int i = t + k * GC_NT;
        c[k] = r[k] = 0u;
        if (i < n) {
            c[k] = gc_cell(g, __ldg(pts + i));
            r[k] = atomicAdd(start + c[k], 1u);
        }
    }
    cluster.sync();
    // phase 2: exclusive scan over `len` entries; thread t owns the run [t S, (t + 1) S), S a multiple of 4
    const long long S = (((len + GC_NT - 1) / GC_NT) + 3) & ~3LL;
    const long long k0 = min((long long)t * S, len), k1 = min(k0 + S, len);
    uint32_t mine = 0;
    for (long long k = k0; k < k1; k += 4) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(start + k));
        mine += (v.x + v.y) + (v.z + v.w);
    }
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = ws[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        ws[lane] = w;  // inclusive over warps
        if (lane == 31) s_total = w;
    }
    __syncthreads();
    uint32_t run = (x - mine) + (warp ? ws[warp - 1] : 0u);  // exclusive inside this CTA
    cluster.sync();
    for (int q = 0; q < rank; q++) run += *cluster.map_shared_rank(&s_total, q);
    for (long long k = k0; k < k1; k += 4) {
        uint4 v = __ldcg(reinterpret_cast<const uint4 *>(start + k));
        uint4 o;
        o.x = run; run += v.x;
        o.y = run; run += v.y;
        o.z = run; run += v.z;
        o.w = run; run += v.w;
        *reinterpret_cast<uint4 *>(start + k) = o;
    }
    cluster.sync();  // also keeps every CTA's s_total alive until all have read it
    // phase 3: scatter
#pragma unroll
    for (int k = 0; k < GC_PPT; k++) {
        const int i = t + k * GC_NT;
        if (i < n) {
            float4 p = __ldg(pts + i);
            p.w = __int_as_float(i);
            sorted[__ldcg(start + c[k]) + r[k]] = p;
        }
    }
}

static inline bool grid_small(int n, long long ncells) {
    const char *env = getenv("PCR_GRID_CLUSTER");  // 0: always the multi-kernel build (tests compare the two)
    if (env && atoi(env) == 0) return false;
    return n <= GC_MAX_POINTS && ncells <= GC_MAX_CELLS;
}

template <typename Dims>
static int grid_cluster_launch(pcr_ctx *ctx, const float4 *pts, int n, const Dims &d, long long ncells, uint32_t **start_out,
                               float4 *sorted) {
    const size_t len = ((size_t)ncells + 1 + 3) & ~(size_t)3;
    uint32_t *start = arena<uint32_t>(ctx, len);
    if (!start) return PCR_ERR_OOM;
    k_grid_cluster<Dims><<<GC_CTAS, GC_THREADS, 0, ctx->stream>>>(pts, n, d, ncells, start, sorted);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    *start_out = start;
    return PCR_OK;
}

int pcr_grid_build(pcr_ctx *ctx, const float4 *pts, int n, double radius, const float *lo_in, const float *hi_in,
                   Grid *g) {
    return pcr_grid_build_rings(ctx, pts, n, radius, 1, lo_in, hi_in, g);
}

int pcr_grid_build_rings(pcr_ctx *ctx, const float4 *pts, int n, double radius, int rings, const float *lo_in,
                         const float *hi_in, Grid *g) {
    if (n <= 0 || !(radius > 0.0) || rings < 1 || rings > 2)
        return pcr_fail(ctx, PCR_ERR_INVALID, "grid build: n=%d radius=%g rings=%d", n, radius, rings);
    float lo[3], hi[3];
    if (lo_in && hi_in) {
        for (int d = 0; d < 3; d++) { lo[d] = lo_in[d]; hi[d] = hi_in[d]; }
    } else {
        PCR_TRY(pcr_bounds(ctx, pts, n, lo, hi));
    }
    for (int d = 0; d < 3; d++)
        if (!(lo[d] <= hi[d]) || isinf(lo[d]) || isinf(hi[d]))
            return pcr_fail(ctx, PCR_ERR_INVALID, "grid build: non-finite coordinates");
    // cell size: radius with a 2^-10 margin, enlarged (x1.25 steps) until the dense table fits the budget
    const double reach = radius * (1.0 + 1.0 / 1024.0);
    double h = reach / rings;
    long long nx, ny, nz;
    for (;;) {
        nx = (long long)floor(((double)hi[0] - (double)lo[0]) / h) + 1;
        ny = (long long)floor(((double)hi[1] - (double)lo[1]) / h) + 1;
        nz = (long long)floor(((double)hi[2] - (double)lo[2]) / h) + 1;
        if ((double)nx * (double)ny * (double)nz <= (double)PCR_MAX_GRID_CELLS) break;
        h *= 1.25;
    }
    const long long ncells = nx * ny * nz;
    GridDims gd{(double)lo[0], (double)lo[1], (double)lo[2], 1.0 / h, (int)nx, (int)ny, (int)nz};
    PCR_ALLOC(sorted, float4, (size_t)n);
    uint32_t *start = nullptr;
    if (grid_small(n, ncells)) {
        KScope ks(ctx, KC_GRID_BUILD, 56.0 * n + 8.0 * (double)ncells, 1);
        PCR_TRY(grid_cluster_launch(ctx, pts, n, gd, ncells, &start, sorted));
    } else {
        PCR_ALLOC(cell, uint32_t, 2 * (size_t)n);
        uint32_t *rank = cell + n;
        KScope ks(ctx, KC_GRID_BUILD, 56.0 * n + 8.0 * (double)ncells, 3);
        SortTables tb;
        PCR_TRY(sort_tables_alloc(ctx, ncells, &tb));
        start = tb.start;
        k_cell_count<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, gd, cell, rank, start);
        PCR_LAUNCHED();
        PCR_TRY(sort_tables_scan(ctx, tb, ncells));
        k_cell_scatter<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, cell, rank, start, sorted);
        PCR_LAUNCHED();
        PCR_CUDA(cudaGetLastError());
    }
    g->sorted = sorted;
    g->start = start;
    g->ox = gd.ox; g->oy = gd.oy; g->oz = gd.oz;
    g->inv_h = gd.inv_h;
    g->h = h;
    g->nx = gd.nx; g->ny = gd.ny; g->nz = gd.nz;
    g->n = n;
    g->R = 1;
    g->big = (size_t)ncells * sizeof(uint32_t) > ((size_t)48 << 20) ? 1 : 0;
    g->blk = nullptr;
    g->cstart = nullptr;
    while ((double)g->R * h < reach) g->R++;  // R h >= radius (1 + 2^-10): the block covers the radius with the margin
    return PCR_OK;
}

// ---- compact grid (two-level table, see Grid) ----------------------------------------------------------------------
// Same cells, same sorted order as the dense grid (ascending cell id; the order inside a cell is arbitrary there too), but
// no per-cell table: a counting sort over BLOCKS of 32 consecutive cell ids, then every point ranks itself among the few
// members of its block.  Tables: cnt/bstart[nblk + 1] and mask/popc[nblk + 1] (one allocation, one memset).
__global__ void __launch_bounds__(256) k_cg_count(const float4 *__restrict__ pts, int n, GridDims g, uint32_t *__restrict__ cell,
                                                  uint32_t *__restrict__ rank, uint32_t *__restrict__ cnt, uint32_t *__restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell_of(g, __ldg(pts + i));
    cell[i] = c;
    rank[i] = atomicAdd(cnt + (c >> 5), 1u);
    atomicOr(mask + (c >> 5), 1u << (c & 31u));
}

// pc[b] = popc(mask[b]) (scanned in place afterwards: number of occupied cells before block b)
__global__ void __launch_bounds__(256) k_cg_popc(const uint32_t *__restrict__ mask, long long nblk, uint32_t *__restrict__ pc) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblk) pc[b] = __popc(mask[b]);
}

// block-ordered copies of (cell id, original index): slot = bstart[block] + arrival rank
__global__ void __launch_bounds__(256) k_cg_scatter(int n, const uint32_t *__restrict__ cell, const uint32_t *__restrict__ rank,
                                                    const uint32_t *__restrict__ bstart, uint2 *__restrict__ slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cell[i];
    slot[bstart[c >> 5] + rank[i]] = make_uint2(c, (uint32_t)i);
}

// every slot ranks itself among the members of its block by (cell id, slot); the first slot of a cell publishes the
// cell's start; blk[b] = (occupied cells before b, mask) is written by the thread of the block's first slot, the entries
// of empty blocks by k_cg_pack below
__global__ void __launch_bounds__(256) k_cg_rank(const float4 *__restrict__ pts, int n, const uint2 *__restrict__ slot,
                                                 const uint32_t *__restrict__ bstart, const uint32_t *__restrict__ pc,
                                                 const uint32_t *__restrict__ mask, float4 *__restrict__ sorted,
                                                 uint32_t *__restrict__ cstart) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint2 me = slot[k];
    const uint32_t b = me.x >> 5;
    const uint32_t s0 = bstart[b], s1 = bstart[b + 1];
    uint32_t below = 0, same_before = 0;
    for (uint32_t m = s0; m < s1; m++) {
        const uint32_t cm = slot[m].x;
        below += cm < me.x;
        same_before += (cm == me.x) & (m < (uint32_t)k);
    }
    float4 p = __ldg(pts + me.y);
    p.w = __int_as_float((int)me.y);
    sorted[s0 + below + same_before] = p;
    if (same_before == 0) cstart[pc[b] + __popc(mask[b] & ((1u << (me.x & 31u)) - 1u))] = s0 + below;
}

__global__ void __launch_bounds__(256) k_cg_pack(const uint32_t *__restrict__ pc, const uint32_t *__restrict__ mask, long long nblk,
                                                 uint2 *__restrict__ blk, uint32_t *__restrict__ cstart, int n) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= nblk) blk[b] = make_uint2(pc[b], b < nblk ? mask[b] : 0u);  // entry nblk: (n_occ, 0), for c = ncells
    if (b == 0) cstart[pc[nblk]] = (uint32_t)n;                            // sentinel after the last occupied cell
}

int pcr_grid_build_compact(pcr_ctx *ctx, const float4 *pts, int n, double radius, const float *lo_in, const float *hi_in, Grid *g) {
    if (n <= 0 || !(radius > 0.0)) return pcr_fail(ctx, PCR_ERR_INVALID, "grid build: n=%d radius=%g", n, radius);
    float lo[3], hi[3];
    if (lo_in && hi_in) {
        for (int d = 0; d < 3; d++) { lo[d] = lo_in[d]; hi[d] = hi_in[d]; }
    } else {
        PCR_TRY(pcr_bounds(ctx, pts, n, lo, hi));
    }
    for (int d = 0; d < 3; d++)
        if (!(lo[d] <= hi[d]) || isinf(lo[d]) || isinf(hi[d]))
            return pcr_fail(ctx, PCR_ERR_INVALID, "grid build: non-finite coordinates");
    // small tables: the dense form is one dependent load shorter per lookup and its build is cheap
    {
        const double h0 = radius * (1.0 + 1.0 / 1024.0);
        const double cells0 = (floor(((double)hi[0] - lo[0]) / h0) + 1) * (floor(((double)hi[1] - lo[1]) / h0) + 1) *
                              (floor(((double)hi[2] - lo[2]) / h0) + 1);
        const char *env = getenv("PCR_GRID_COMPACT");  // read per call: the tests compare the two forms in one process
        const bool off = env && atoi(env) == 0;
        if (off || cells0 <= (double)(1 << 22)) return pcr_grid_build_rings(ctx, pts, n, radius, 1, lo, hi, g);
    }
    const double reach = radius * (1.0 + 1.0 / 1024.0);
    double h = reach;
    long long nx, ny, nz;
    for (;;) {  // cell ids are 32-bit and the block table takes a quarter byte per cell: up to 2^30 cells
        nx = (long long)floor(((double)hi[0] - (double)lo[0]) / h) + 1;
        ny = (long long)floor(((double)hi[1] - (double)lo[1]) / h) + 1;
        nz = (long long)floor(((double)hi[2] - (double)lo[2]) / h) + 1;
        if ((double)nx * (double)ny * (double)nz <= (double)(1LL << 30)) break;
        h *= 1.25;
    }
    const long long ncells = nx * ny * nz;
    const long long nblk = (ncells >> 5) + 1;  // blocks that hold a cell id 0 .. ncells (c = ncells is looked up as an end)
    GridDims gd{(double)lo[0], (double)lo[1], (double)lo[2], 1.0 / h, (int)nx, (int)ny, (int)nz};
    PCR_ALLOC(cell, uint32_t, 2 * (size_t)n);
    uint32_t *rank = cell + n;
    PCR_ALLOC(slot, uint2, (size_t)n);
    PCR_ALLOC(sorted, float4, (size_t)n);
    PCR_ALLOC(cstart, uint32_t, (size_t)n + 1);
    PCR_ALLOC(blk, uint2, (size_t)nblk + 1);
    KScope ks(ctx, KC_GRID_BUILD, 80.0 * n + 40.0 * (double)nblk, 7);
    // [cnt -> bstart: nblk + 1][mask: nblk + 1][pc: nblk + 1][scan state x 2], one memset
    const int tiles = div_up(nblk, 4096);
    const size_t tab = (((size_t)nblk + 1) * sizeof(uint32_t) + 15) & ~(size_t)15;
    const size_t st_bytes = sizeof(unsigned long long) * ((size_t)tiles + 1);
    unsigned char *base = arena<unsigned char>(ctx, 3 * tab + 2 * st_bytes);
    if (!base) return PCR_ERR_OOM;
    PCR_CUDA(cudaMemsetAsync(base, 0, 3 * tab + 2 * st_bytes, ctx->stream));
    uint32_t *cnt = (uint32_t *)base, *mask = (uint32_t *)(base + tab), *pc = (uint32_t *)(base + 2 * tab);
    unsigned long long *st_a = (unsigned long long *)(base + 3 * tab), *st_b = (unsigned long long *)(base + 3 * tab + st_bytes);
    k_cg_count<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, gd, cell, rank, cnt, mask);
    PCR_LAUNCHED();
    k_scan_onepass<uint32_t><<<tiles, 1024, 0, ctx->stream>>>(cnt, nblk, st_a, (unsigned int *)(st_a + tiles));
    PCR_LAUNCHED();
    k_cg_popc<<<div_up(nblk, 256), 256, 0, ctx->stream>>>(mask, nblk, pc);
    PCR_LAUNCHED();
    k_scan_onepass<uint32_t><<<tiles, 1024, 0, ctx->stream>>>(pc, nblk, st_b, (unsigned int *)(st_b + tiles));
    PCR_LAUNCHED();
    k_cg_scatter<<<div_up(n, 256), 256, 0, ctx->stream>>>(n, cell, rank, cnt, slot);
    PCR_LAUNCHED();
    k_cg_rank<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, slot, cnt, pc, mask, sorted, cstart);
    PCR_LAUNCHED();
    k_cg_pack<<<div_up(nblk + 1, 256), 256, 0, ctx->stream>>>(pc, mask, nblk, blk, cstart, n);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    g->sorted = sorted;
    g->start = nullptr;
    g->blk = blk;
    g->cstart = cstart;
    g->ox = gd.ox; g->oy = gd.oy; g->oz = gd.oz;
    g->inv_h = gd.inv_h;
    g->h = h;
    g->nx = gd.nx; g->ny = gd.ny; g->nz = gd.nz;
    g->n = n;
    g->R = 1;
    while ((double)g->R * h < reach) g->R++;
    g->big = 0;
    return PCR_OK;
}

// ---- Morton-order sort ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_morton_count(const float4 *__restrict__ pts, int n, MortonDims g,
                                                      uint32_t *__restrict__ cell, uint32_t *__restrict__ rank,
                                                      uint32_t *__restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = morton_of(g, __ldg(pts + i));
    cell[i] = c;
    rank[i] = atomicAdd(count + c, 1u);
}

int pcr_morton_sort(pcr_ctx *ctx, const float4 *pts, int n, const float4 **sorted_out) {
    if (n <= 0) return pcr_fail(ctx, PCR_ERR_INVALID, "morton sort: empty cloud");
    float lo[3], hi[3];
    PCR_TRY(pcr_bounds(ctx, pts, n, lo, hi));
    double ext = 0.0;
    for (int d = 0; d < 3; d++) {
        if (!(lo[d] <= hi[d]) || isinf(lo[d]) || isinf(hi[d])) return pcr_fail(ctx, PCR_ERR_INVALID, "morton sort: non-finite coordinates");
        ext = fmax(ext, (double)hi[d] - (double)lo[d]);
    }
    int L = 4;
    while (L < 8 && (1LL << (3 * L)) < 4LL * n) L++;
    const long long ncells = 1LL << (3 * L);
    MortonDims md{(double)lo[0], (double)lo[1], (double)lo[2], ext > 0.0 ? (double)(1 << L) / (ext * (1.0 + 1e-6)) : 0.0, (1 << L) - 1};
    PCR_ALLOC(sorted, float4, (size_t)n);
    if (grid_small(n, ncells)) {
        KScope ks(ctx, KC_GRID_BUILD, 56.0 * n + 8.0 * (double)ncells, 1);
        uint32_t *start = nullptr;
        PCR_TRY(grid_cluster_launch(ctx, pts, n, md, ncells, &start, sorted));
    } else {
        PCR_ALLOC(cell, uint32_t, 2 * (size_t)n);
        uint32_t *rank = cell + n;
        KScope ks(ctx, KC_GRID_BUILD, 56.0 * n + 8.0 * (double)ncells, 3);
        SortTables tb;
        PCR_TRY(sort_tables_alloc(ctx, ncells, &tb));
        uint32_t *start = tb.start;
        k_morton_count<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, md, cell, rank, start);
        PCR_LAUNCHED();
        PCR_TRY(sort_tables_scan(ctx, tb, ncells));
        k_cell_scatter<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts, n, cell, rank, start, sorted);
        PCR_LAUNCHED();
        PCR_CUDA(cudaGetLastError());
    }
    *sorted_out = sorted;
    return PCR_OK;
}
