// pcr_common.cuh — shared declarations of the B200 registration engine (sm_100a only).
//
// Arithmetic specification — every translation unit is compiled with -fmad=false so that
// fp32/fp64 products and sums round exactly as the CPU oracle's; FMA is used only where written explicitly.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pcr.h"
#include "../../include/pcr_detmath.h"

#define PCR_MAX_GRID_CELLS (1LL << 27)

// kernel classes for the optional per-kernel timing (pcr_set_profiling / pcr_kernel_stats)
enum KClass {
    KC_PACK = 0, KC_BOUNDS, KC_GRID_BUILD, KC_SCAN, KC_VOXEL, KC_KNN_LIST, KC_KNN_COV, KC_NORMALS_SOLVE, KC_SPFH, KC_FPFH,
    KC_NN_FEATURES, KC_MATCH_MISC, KC_RANSAC_GENERATE, KC_RANSAC_VALIDATE, KC_RANSAC_STEP, KC_ICP_PASS, KC_NN1, KC_COUNT
};

struct KPending {
    int id;
    cudaEvent_t a, b;
    double bytes;   // algorithmic bytes per launch
    double flops;   // algorithmic flops per launch (descriptor GEMM only)
    long long launches;
};

// ---- uniform grid -----------------------------------------------------------------------------------
// Dense grid over the bounding box of the indexed cloud; cells ordered x-fastest so that the three
// x-neighbours of a cell are one contiguous range of the sorted point array.
struct Grid {
    const float4 *sorted;   // points sorted by cell; .w = original index (int bits)
    const uint32_t *start;  // ncells+1 exclusive prefix
    double ox, oy, oz;      // origin (min bound)
    double inv_h;           // 1 / cell size
    double h;
    int nx, ny, nz;
    int n;
    int R;                  // rings of cells that cover the search radius (1 unless built by pcr_grid_build_rings)
    int big;                // the `start` table does not fit L2 comfortably: searches prefetch the rows of their block
    // COMPACT form (pcr_grid_build_compact; blk != nullptr, start == nullptr): the dense `start` table of a fine grid over
    // a surface is almost all empty cells (1M points in 437^3 cells: 340 MB, beyond L2, and its memset + scan dominated
    // the ICP preparation).  Two levels instead: per block of 32 consecutive cell ids (an x-run) blk[b] = (number of
    // OCCUPIED cells before the block, occupancy mask of the block), and cstart[r] = first sorted position of the r-th
    // occupied cell (cstart[n_occ] = n).  start(c) = cstart[blk[c >> 5].x + popc(blk[c >> 5].y & lowbits(c & 31))]:
    // 8 bytes per 32 cells + 4 bytes per occupied cell, one more dependent load than the dense table, all of it in L2.
    const uint2 *blk;
    const uint32_t *cstart;
};


// prepared ICP work: everything that depends only on the two clouds (pcr_icp.cu); may be built on another context
struct IcpPrep {
    Grid g;
    const float4 *src_sorted = nullptr;
    float amax = 0.0f;
    bool valid = false;
};

// EXPERIMENTAL (off unless PCR_VAL_LISTS=1; pcr_celllists.cu): per-fine-cell candidate lists over the target cloud.
// For a fine cell C (side c = radius / (1.5 div)) the list holds every target point t with
// dmin(t, C) <= min(min_t' dmax(t', C), r) (1 + 1e-5): the nearest neighbour of EVERY query inside C is on the list, so a
// radius-limited 1-NN tests 3-6 points instead of walking 27 coarse cells.  The winner is chosen with the same fp32
// distance and (d2, index) key as grid_nn1, hence identical results (NumPy prototype + proof:
// tools/proto/cell_candidate_lists.py).  Not yet measured on the GPU.
struct CellLists {
    const uint32_t *head;  // per fine cell: (offset << 4) | count; count 15 = list too long -> full search
    const float4 *items;   // candidate points, .w = original index (int bits).  (4-byte positions into the cell-sorted cloud
                           // were measured: the pool shrinks 4x but the dependent load costs more: validate 0.40 -> 0.43 ms)
    double ox, oy, oz, inv_c;
    int nx, ny, nz;
    unsigned int cap;      // capacity of `items` (host bookkeeping)
};

// prepared RANSAC work (target grid + spatially sorted source), see pcr_ransac.cu
struct RansacWork {
    Grid g;
    const float4 *src_sorted;
    float r2;
    int k_d;
    CellLists cl;
    bool use_lists;
};

// One RANSAC wave in two steps (pcr_ransac.cu): generation does not depend on the best result so far, validation does — so the
// next wave can be generated (on another stream) while this one is validated and, on several GPUs, exchanged.
struct Survivor;
struct WaveWork {
    Survivor *surv = nullptr;          // set by the caller: use these buffers; nullptr: wave_generate takes arena scratch
    unsigned char *hdr = nullptr;      // 4 counters (survivors, records, ticket, -) + the record buffer (cap records)
    int *bucket_best = nullptr;
    long long hyp_begin = 0, count = 0;
    int cap = 0;
};
size_t pcr_wave_survivor_bytes(long long count);

struct pcr_ctx {
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<KPending> pending;
    double k_ms[KC_COUNT] = {0};
    double k_ms_helper[KC_COUNT] = {0};  // share of k_ms measured on the helper context
    double k_bytes[KC_COUNT] = {0};
    double k_flops[KC_COUNT] = {0};
    long long k_launches[KC_COUNT] = {0};
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    // bump arena for per-call scratch (reset at the start of every exported call)
    std::vector<void *> blocks;
    std::vector<size_t> block_sizes;
    size_t cur_block = 0, cur_off = 0;
    // pinned host staging for small result records
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    // pinned staging for file input (pcr_align_files): grow-only, freed at destroy
    void *stage = nullptr;
    size_t stage_bytes = 0;
    std::atomic<bool> busy{false};  // one exported call at a time per context: the second caller gets PCR_ERR_BUSY
    // per-context (hence per-device and per-owner-thread) launch-configuration caches: occupancy and function
    // attributes are properties of (function, device), and contexts may be created from several host threads
    int occ_icp = 0, occ_val256 = 0, occ_val512 = 0;
    bool match_tc_attr_set = false;
    // bounding boxes already reduced during the current exported call, keyed by (pointer, n); cleared on entry
    struct BoundsEntry { const void *ptr; int n; float lo[3], hi[3]; };
    std::vector<BoundsEntry> bounds_cache;
    unsigned int *bounds_ticket = nullptr;  // device word: completion ticket of k_bounds2 (zero between launches)
    // RANSAC session (pcr_ransac_session_begin/end): prepared work in buffers that outlive the per-call arena
    struct RansacSession {
        bool active = false;
        const void *src = nullptr, *tgt = nullptr;
        int ms = 0, mt = 0;
        double max_dist = 0.0;
        const void *corr_ok = nullptr;  // correspondence buffer already validated against (ms, mt) in this session
        int corr_ok_c = 0;
        RansacWork w;
        // grid.sorted, grid.start, src_sorted, and (experimental candidate lists) head, items: grow-only, freed at destroy
        void *bufs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        size_t cap[5] = {0, 0, 0, 0, 0};
    } rsess;
    // pcr_align overlaps independent stages (the two clouds' preprocessing; full-resolution normals next to RANSAC) on
    // a second context with its own stream, arena and pinned page, driven by one persistent host thread (the stages
    // contain host synchronisations, so a second stream alone would not overlap them)
    pcr_ctx *helper = nullptr;
    struct Worker *worker = nullptr;
    bool owns_stream = false;
    int yielding = 0;  // > 0: this context runs beside a critical path; its long kernels use CTAs of `yielding` queries per warp
    void *dist = nullptr;              // multi-GPU state (pcr_dist.cu): NCCL communicator, exchange buffers, worker contexts
    cudaStream_t hp_stream = nullptr;  // highest-priority stream: pcr_align's critical path runs here while the helper works
    cudaStream_t aux_stream = nullptr; // second highest-priority stream: work pcr_align issues next to its critical path
    // Host waits: spinning cudaStreamSynchronize by default (lowest latency for one alignment at a time).  pcr_align_batch
    // switches its worker contexts to a blocking event wait when there are more host threads than cores (several ranks x
    // workers x helper threads on one box): spinning threads then only steal the cores the launching threads need.
    bool blocking_sync = false;
    cudaEvent_t sync_ev = nullptr;
    cudaStream_t aux2_stream = nullptr; // third: the target -> source direction of the descriptor matching (pcr_match.cu)
};

// one persistent host thread executing one task at a time
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> task;
    bool has_task = false, done = true, quit = false;
    int rc = 0;
    Worker();
    ~Worker();
    void submit(std::function<int()> f);  // returns immediately
    int wait();                           // joins the task, returns its status
};

extern std::atomic<long long> g_pcr_launches;
#define PCR_LAUNCHED() (g_pcr_launches++)

// RAII scope: when profiling is on, brackets the kernels launched inside it with a CUDA event pair on the
// context's stream; pcr_kernel_stats() later resolves the pairs.  `launches` kernels of the class are inside.
struct KScope {
    pcr_ctx *ctx;
    KPending p;
    bool on;
    KScope(pcr_ctx *c, int id, double bytes_per_launch, long long launches = 1, double flops_per_launch = 0.0);
    ~KScope();
    void set_launches(long long n) { p.launches = n; }
};

cudaError_t pcr_sync_stream(pcr_ctx *ctx, cudaStream_t s);  // cudaStreamSynchronize, or a blocking event wait (ctx->blocking_sync)

// ---- error handling -------------------------------------------------------------------------------
int pcr_fail(pcr_ctx *ctx, int code, const char *fmt, ...);
#define PCR_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return pcr_fail(ctx, PCR_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)
#define PCR_TRY(call)            \
    do {                         \
        int r__ = (call);        \
        if (r__ != PCR_OK) return r__; \
    } while (0)

// ---- arena ----------------------------------------------------------------------------------------
void pcr_arena_reset(pcr_ctx *ctx);
void *pcr_arena_alloc(pcr_ctx *ctx, size_t bytes);  // 256-byte aligned; nullptr on failure (err set)
template <typename T>
static inline T *arena(pcr_ctx *ctx, size_t n) {
    return reinterpret_cast<T *>(pcr_arena_alloc(ctx, n * sizeof(T)));
}
#define PCR_ALLOC(ptr, T, n)                                  \
    T *ptr = arena<T>(ctx, (n));                              \
    if (!ptr) return PCR_ERR_OOM

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- host-side spec helpers (restated independently in oracle/pcr_oracle.c) ---------------------------
static inline int pcr_ilog2ceil(long long n) {
    int e = 0;
    while ((1LL << e) < n) e++;
    return e;
}
int pcr_pow2ceil_exp(double x);  // smallest e with 2^e >= x

// bounds: lo/hi (host) of a float4 cloud; one device reduction + one D2H sync
int pcr_bounds(pcr_ctx *ctx, const float4 *pts, int n, float lo[3], float hi[3]);
// both boxes with ONE launch and ONE synchronisation, into the per-call cache (either cloud may already be cached)
int pcr_bounds_pair(pcr_ctx *ctx, const float4 *pa, int na, const float4 *pb, int nb);
// build a search grid supporting radius-`radius` queries with a 3x3x3 cell probe.
// If bounds are already known pass them (have_bounds), else they are computed.
int pcr_grid_build(pcr_ctx *ctx, const float4 *pts, int n, double radius, const float *lo, const float *hi, Grid *g);
// same, with cells of size radius / rings (rings = 1 or 2): a radius query then probes a (2 R + 1)^3 block, R = g->R.
// Finer cells pay off for k-nearest queries on dense clouds, where the k-th neighbour is much closer than the radius.
int pcr_grid_build_rings(pcr_ctx *ctx, const float4 *pts, int n, double radius, int rings, const float *lo, const float *hi,
                         Grid *g);
// same search structure in the COMPACT form (see Grid) when the dense table would be large, else the dense grid; only
// for kernels that read the table through grid_start_at / grid_range (grid_nn1, grid_nn1_cert)
int pcr_grid_build_compact(pcr_ctx *ctx, const float4 *pts, int n, double radius, const float *lo, const float *hi, Grid *g);
// Morton-order copy of a cloud (dense counting sort on interleaved cell ids; .w = original index): consecutive
// points are spatially compact, which the warp-cooperative joins rely on
int pcr_morton_sort(pcr_ctx *ctx, const float4 *pts, int n, const float4 **sorted_out);
// EXPERIMENTAL candidate lists over the points of a grid (pcr_celllists.cu); *ok = false: keep the full search
int pcr_celllists_build(pcr_ctx *ctx, const Grid &g, double r, int div, CellLists *out, bool *ok);
// exclusive scan helpers (in pcr_grid.cu)
int pcr_exclusive_scan_u32(pcr_ctx *ctx, uint32_t *data, long long n);  // in place, data[n] must exist (total)
int pcr_exclusive_scan_u64(pcr_ctx *ctx, unsigned long long *data, long long n);

// ---- device helpers ---------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float dist2f(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));  // D1
}

// p' = fp32(R p + t), fp64 products/sums individually rounded (D7)
__device__ __forceinline__ float3 xform_pt(const double *__restrict__ T, float x, float y, float z) {
    const double dx = x, dy = y, dz = z;
    float3 o;
    o.x = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], dx), __dmul_rn(T[1], dy)), __dmul_rn(T[2], dz)), T[3]);
    o.y = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], dx), __dmul_rn(T[5], dy)), __dmul_rn(T[6], dz)), T[7]);
    o.z = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], dx), __dmul_rn(T[9], dy)), __dmul_rn(T[10], dz)), T[11]);
    return o;
}

// cell coordinate in fp64 (exact to ~1e-13 cells; the cell size carries a 2^-10 margin over the radius so a
// point with fp32 d2 < r2 is always within one cell of the query).  Clamped to [-2, n+1] before the int cast.
__device__ __forceinline__ int grid_cell(double v, double o, double inv_h, int n) {
    double c = floor((v - o) * inv_h);
    c = fmin(fmax(c, -2.0), (double)n + 1.0);
    return (int)c;
}

// L2 prefetch of the cache line that holds *p (no register, no dependency): the searches below issue it for the `start`
// entries of all nine rows of the 3x3x3 block before they walk the first one, so that on a table larger than L2 (the
// dense ICP grid at 1M points is 340 MB) the later rows cost an L2 round trip instead of a DRAM one.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

typedef unsigned long long pcr_u64k;

// number of points whose cell id is below c (0 <= c <= ncells): the dense table entry, or the two-level lookup
__device__ __forceinline__ uint32_t grid_start_at(const Grid &g, long long c) {
    if (!g.blk) return __ldg(g.start + c);
    const uint2 b = __ldg(g.blk + (c >> 5));
    return __ldg(g.cstart + (b.x + __popc(b.y & ((1u << (int)(c & 31)) - 1u))));
}
// [*b, *e) = sorted positions of the cells c0 .. c1 (inclusive, c0 <= c1 in one row); one block entry serves both ends when
// they share a block
__device__ __forceinline__ void grid_range(const Grid &g, long long c0, long long c1, uint32_t *b, uint32_t *e) {
    if (!g.blk) {
        *b = __ldg(g.start + c0);
        *e = __ldg(g.start + c1 + 1);
        return;
    }
    const long long c2 = c1 + 1;
    const uint2 b0 = __ldg(g.blk + (c0 >> 5));
    const uint2 b1 = (c2 >> 5) == (c0 >> 5) ? b0 : __ldg(g.blk + (c2 >> 5));
    *b = __ldg(g.cstart + (b0.x + __popc(b0.y & ((1u << (int)(c0 & 31)) - 1u))));
    *e = __ldg(g.cstart + (b1.x + __popc(b1.y & ((1u << (int)(c2 & 31)) - 1u))));
}

// smallest (d2, index) key over the contiguous range [b, e) of the cell-sorted cloud, FOUR loads in flight per lane
// (the search is a chain of dependent loads: one candidate per round trip was ~700 cycles each, profiles/r2_summary.md).
// Indices past the end are clamped to the last point: examining a point twice cannot change a minimum.
__device__ __forceinline__ pcr_u64k grid_scan_min(const float4 *__restrict__ sorted, uint32_t b, uint32_t e, float qx, float qy,
                                                  float qz, pcr_u64k bkey) {
#pragma unroll 1
    for (uint32_t k = b; k < e; k += 4) {
        const uint32_t last = e - 1;
        const float4 p0 = __ldg(sorted + k), p1 = __ldg(sorted + min(k + 1, last));
        const float4 p2 = __ldg(sorted + min(k + 2, last)), p3 = __ldg(sorted + min(k + 3, last));
        const pcr_u64k k0 = (((pcr_u64k)__float_as_uint(dist2f(qx, qy, qz, p0.x, p0.y, p0.z))) << 32) | (uint32_t)__float_as_int(p0.w);
        const pcr_u64k k1 = (((pcr_u64k)__float_as_uint(dist2f(qx, qy, qz, p1.x, p1.y, p1.z))) << 32) | (uint32_t)__float_as_int(p1.w);
        const pcr_u64k k2 = (((pcr_u64k)__float_as_uint(dist2f(qx, qy, qz, p2.x, p2.y, p2.z))) << 32) | (uint32_t)__float_as_int(p2.w);
        const pcr_u64k k3 = (((pcr_u64k)__float_as_uint(dist2f(qx, qy, qz, p3.x, p3.y, p3.z))) << 32) | (uint32_t)__float_as_int(p3.w);
        const pcr_u64k ka = k0 < k1 ? k0 : k1, kb = k2 < k3 ? k2 : k3;
        const pcr_u64k kc = ka < kb ? ka : kb;
        bkey = kc < bkey ? kc : bkey;
    }
    return bkey;
}

// radius-limited 1-NN in a grid: best (d2, idx) under the (d2, idx) lexicographic order, d2 < r2 strictly.
// Returns original index or -1.  `seed` (optional, an index into `tgt_orig`, e.g. the previous ICP pass's
// correspondence) only tightens the initial bound: the result is identical with or without it, because every
// cell that intersects the closed ball of the current best distance is still visited (rows by slab distance,
// cells of a row by x distance; both bounds carry a 1e-4-cell slack that dominates the fp32 rounding involved,
// and cells at exactly the best distance are kept since ties are decided by index).  Visiting MORE cells of the block
// than necessary never changes the result, so the row ranges of a z-slab are fetched together (one round trip for up
// to three rows) with the x window the bound allowed when the slab was entered.
__device__ __forceinline__ int grid_nn1_seeded(const Grid &g, float qx, float qy, float qz, float r2, int seed,
                                               const float4 *__restrict__ tgt_orig, float *d2_out) {
    const double fx = ((double)qx - g.ox) * g.inv_h, fy = ((double)qy - g.oy) * g.inv_h, fz = ((double)qz - g.oz) * g.inv_h;
    const int cx = (int)fmin(fmax(floor(fx), -2.0), (double)g.nx + 1.0);
    const int cy = (int)fmin(fmax(floor(fy), -2.0), (double)g.ny + 1.0);
    const int cz = (int)fmin(fmax(floor(fz), -2.0), (double)g.nz + 1.0);
    // best (d2, idx) as ONE 64-bit key (fp32 bits of d2 >= 0 order like the value; index in the low word): the
    // lexicographic tie rule D2 becomes a branch-free unsigned compare.  Start key = (r2, 0): only d2 < r2 beats it.
    typedef pcr_u64k u64k;
    const uint32_t r2bits = __float_as_uint(r2);
    u64k bkey = ((u64k)r2bits) << 32;
    if (seed >= 0) {
        const float4 p = __ldg(tgt_orig + seed);
        const float d2 = dist2f(qx, qy, qz, p.x, p.y, p.z);
        const u64k k = (((u64k)__float_as_uint(d2)) << 32) | (uint32_t)seed;
        bkey = k < bkey ? k : bkey;
    }
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    if (x0 <= x1) {
        const int y0 = max(cy - 1, 0), y1 = min(cy + 1, g.ny - 1);
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.nz - 1);
        const float fxf = (float)fmin(fmax(fx, -4.0), (double)g.nx + 4.0);
        const float fyf = (float)fmin(fmax(fy, -4.0), (double)g.ny + 4.0), fzf = (float)fmin(fmax(fz, -4.0), (double)g.nz + 4.0);
        const float h2 = (float)(g.h * g.h), inv_hf = (float)g.inv_h;
        if (g.big) {
            for (int z = z0; z <= z1; z++)
                for (int y = y0; y <= y1; y++) prefetch_l2(g.start + ((long long)z * g.ny + y) * g.nx + x0);
        }
        // home row first: it usually holds the nearest point
        if (cy >= y0 && cy <= y1 && cz >= z0 && cz <= z1) {
            const float best = __uint_as_float((uint32_t)(bkey >> 32));
            const float rc = sqrtf(best) * inv_hf * 1.0001f + 1e-4f;
            const int xa = max(x0, (int)floorf(fxf - rc)), xb = min(x1, (int)floorf(fxf + rc));
            if (xa <= xb) {
                const long long row = ((long long)cz * g.ny + cy) * g.nx;
                uint32_t b, e;
                grid_range(g, row + xa, row + xb, &b, &e);
                bkey = grid_scan_min(g.sorted, b, e, qx, qy, qz, bkey);
            }
        }
        // then only the rows inside the window that the current best distance reaches in y and z (typically 1-3 of the
        // 8): rows outside it have a slab distance above the best, exactly the rows the per-row test would skip
        {
            const float best0 = __uint_as_float((uint32_t)(bkey >> 32));
            const float rw = sqrtf(best0) * inv_hf * 1.0001f + 1e-4f;
            const int ya = max(y0, (int)floorf(fyf - rw)), yb = min(y1, (int)floorf(fyf + rw));
            const int za = max(z0, (int)floorf(fzf - rw)), zb = min(z1, (int)floorf(fzf + rw));
            const int xa = max(x0, (int)floorf(fxf - rw)), xb = min(x1, (int)floorf(fxf + rw));
            if (xa <= xb) {
#pragma unroll 1
                for (int z = za; z <= zb; z++) {
                    uint32_t rb[3], re[3];  // ya..yb is at most three rows (a subrange of y0..y1)
#pragma unroll
                    for (int o = 0; o < 3; o++) {
                        const int y = ya + o;
                        const bool ok = y <= yb && !(y == cy && z == cz);
                        const long long row = ((long long)z * g.ny + (ok ? y : ya)) * g.nx;
                        rb[o] = re[o] = 0u;
                        if (ok) grid_range(g, row + xa, row + xb, &rb[o], &re[o]);
                    }
#pragma unroll
                    for (int o = 0; o < 3; o++) {
                        if (rb[o] >= re[o]) continue;
                        const int y = ya + o;
                        const float best = __uint_as_float((uint32_t)(bkey >> 32));
                        const float ey = fmaxf(fmaxf((float)y - fyf, fyf - (float)(y + 1)), 0.0f);
                        const float ez = fmaxf(fmaxf((float)z - fzf, fzf - (float)(z + 1)), 0.0f);
                        const float sy = fmaxf(ey - 1e-4f, 0.0f), sz = fmaxf(ez - 1e-4f, 0.0f);
                        if ((sy * sy + sz * sz) * h2 > best) continue;
                        bkey = grid_scan_min(g.sorted, rb[o], re[o], qx, qy, qz, bkey);
                    }
                }
            }
        }
    }
    const uint32_t hi = (uint32_t)(bkey >> 32);
    *d2_out = __uint_as_float(hi);
    return hi < r2bits ? (int)(uint32_t)(bkey & 0xffffffffull) : -1;
}

// Search variant that also returns a CERTIFICATE for later reuse.  It examines EVERY point of the 3x3x3 cell block
// around the query (no pruning: the row ranges are fetched three rows per round trip, the candidates four per round
// trip), i.e. every target point within one cell size h of the query, and keeps the two smallest keys (d2, index) and
// the third smallest d2:
//   *d2_out   smallest examined d2 (+inf if the block is empty) — also when it is not below r2, so that a
//             "no correspondence" result carries a certificate too;
//   *other_lb lower bound of the squared distance from the query to every target point OTHER than the nearest
//             examined one: min(second smallest examined d2, h^2 (1 - 2e-4));
//   *j2, *third_lb the second nearest examined point (-1 if none) and the same kind of bound for every point other
//             than the nearest two.
// The nearest neighbour itself is found exactly as by grid_nn1 (same key order, same radius rule).
__device__ __forceinline__ int grid_nn1_cert(const Grid &g, float qx, float qy, float qz, float r2, float *d2_out,
                                             float *other_lb, int *j2, float *third_lb) {
    typedef pcr_u64k u64k;
    const double fx = ((double)qx - g.ox) * g.inv_h, fy = ((double)qy - g.oy) * g.inv_h, fz = ((double)qz - g.oz) * g.inv_h;
    const int cx = (int)fmin(fmax(floor(fx), -2.0), (double)g.nx + 1.0);
    const int cy = (int)fmin(fmax(floor(fy), -2.0), (double)g.ny + 1.0);
    const int cz = (int)fmin(fmax(floor(fz), -2.0), (double)g.nz + 1.0);
    const uint32_t r2bits = __float_as_uint(r2);
    const u64k KINF = ((u64k)0x7f800000u) << 32;  // +inf: keys over ALL examined points (the radius rule is applied at the end)
    u64k k1 = KINF, k2 = KINF;
    float third = INFINITY;
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    if (x0 <= x1) {
        if (g.big) {
            for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); z++)
                for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); y++) prefetch_l2(g.start + ((long long)z * g.ny + y) * g.nx + x0);
        }
#pragma unroll 1
        for (int dz = -1; dz <= 1; dz++) {
            const int z = cz + dz;
            if (z < 0 || z >= g.nz) continue;
            uint32_t rb[3], re[3];
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const int y = cy + dy;
                const bool ok = y >= 0 && y < g.ny;
                const long long row = ((long long)z * g.ny + (ok ? y : 0)) * g.nx;
                rb[dy + 1] = re[dy + 1] = 0u;
                if (ok) grid_range(g, row + x0, row + x1, &rb[dy + 1], &re[dy + 1]);
            }
#pragma unroll
            for (int o = 0; o < 3; o++) {
#pragma unroll 1
                for (uint32_t k = rb[o]; k < re[o]; k += 4) {
                    const uint32_t last = re[o] - 1;
                    float4 p[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) p[u] = __ldg(g.sorted + min(k + u, last));
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        // a slot past the end examines nothing: key and distance +inf leave k1, k2 and third unchanged
                        const bool live = k + u <= last;
                        const float d2 = live ? dist2f(qx, qy, qz, p[u].x, p[u].y, p[u].z) : INFINITY;
                        const u64k key = live ? ((((u64k)__float_as_uint(d2)) << 32) | (uint32_t)__float_as_int(p[u].w)) : ~0ull;
                        const bool lt1 = key < k1, lt2 = key < k2;
                        third = lt2 ? __uint_as_float((uint32_t)(k2 >> 32)) : fminf(third, d2);
                        k2 = lt1 ? k1 : (lt2 ? key : k2);
                        k1 = lt1 ? key : k1;
                    }
                }
            }
        }
    }
    const uint32_t hi = (uint32_t)(k1 >> 32);
    *d2_out = __uint_as_float(hi);
    const float hlim = (float)(g.h * g.h) * (1.0f - 2e-4f);
    *other_lb = fminf(__uint_as_float((uint32_t)(k2 >> 32)), hlim);
    *j2 = k2 < KINF ? (int)(uint32_t)(k2 & 0xffffffffull) : -1;
    *third_lb = fminf(third, hlim);
    return hi < r2bits ? (int)(uint32_t)(k1 & 0xffffffffull) : -1;
}

__device__ __forceinline__ int grid_nn1(const Grid &g, float qx, float qy, float qz, float r2, float *d2_out) {
    return grid_nn1_seeded(g, qx, qy, qz, r2, -1, nullptr, d2_out);
}

// radius-limited 1-NN through the candidate lists: same result as grid_nn1 (see CellLists)
__device__ __forceinline__ int lists_nn1(const CellLists &L, const Grid &g, float qx, float qy, float qz, float r2, float *d2_out) {
    typedef unsigned long long u64k;
    const double fx = ((double)qx - L.ox) * L.inv_c, fy = ((double)qy - L.oy) * L.inv_c, fz = ((double)qz - L.oz) * L.inv_c;
    // outside the lattice (it pads the cloud by more than the radius), or NaN: no target point within the radius
    if (!(fx >= 0.0 && fy >= 0.0 && fz >= 0.0 && fx < (double)L.nx && fy < (double)L.ny && fz < (double)L.nz)) {
        *d2_out = r2;
        return -1;
    }
    const uint32_t h = __ldg(L.head + ((long long)(int)fz * L.ny + (int)fy) * L.nx + (int)fx);
    const uint32_t cnt = h & 15u;
    if (cnt == 15u) return grid_nn1(g, qx, qy, qz, r2, d2_out);
    const float4 *__restrict__ it = L.items + (h >> 4);
    const uint32_t r2bits = __float_as_uint(r2);
    u64k bkey = ((u64k)r2bits) << 32;
    for (uint32_t k = 0; k < cnt; k++) {
        const float4 p = __ldg(it + k);
        const float d2 = dist2f(qx, qy, qz, p.x, p.y, p.z);
        const u64k key = (((u64k)__float_as_uint(d2)) << 32) | (uint32_t)__float_as_int(p.w);
        bkey = key < bkey ? key : bkey;
    }
    const uint32_t hi = (uint32_t)(bkey >> 32);
    *d2_out = __uint_as_float(hi);
    return hi < r2bits ? (int)(uint32_t)(bkey & 0xffffffffull) : -1;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// llrint(x * 2^k): exact power-of-two scaling (scale = 2^k precomputed) then round-to-nearest-even (D5)
__device__ __forceinline__ long long fixed_ll(double x, double scale) { return __double2ll_rn(__dmul_rn(x, scale)); }

#endif  // __CUDACC__
