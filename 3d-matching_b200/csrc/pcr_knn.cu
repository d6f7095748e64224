// pcr_knn.cu — hybrid kNN (K3/K4 search), normal estimation (K3) and SPFH/FPFH (K4, K5).
//
// Replaces, on the reference's path:
//   KDTreeFlann::SearchHybrid                           -> k_knn
//   pcd.estimate_normals(Hybrid(2v,30))  src/ply/ply.py:110-112,133-135 (A.2) -> k_knn<MODE_NORMAL> + k_normals_solve
//   compute_fpfh_feature(Hybrid(5v,100)) src/ply/ply.py:117-120         (A.4) -> k_knn<MODE_LIST> + k_spfh + k_fpfh
//
// Search design: one warp per query.  The 3x3x3 cell probe is 9 contiguous row ranges of the cell-sorted
// cloud (x-fastest cell order), so lanes read consecutive float4s (coalesced).  In-radius candidates are
// appended to a per-warp shared-memory buffer as 64-bit keys (fp32 d2 bits << 32 | index): key order IS the
// (d2, index) order of determinism rule D2.  A warp-wide bitonic sort of the buffer yields the neighbour
// list; when the buffer fills it is sorted and truncated to max_nn and the admission threshold tightens.
#include <cstdio>
#include <cstdlib>

#include "pcr_common.cuh"

constexpr int KNN_WARPS = 4;
constexpr int KNN_CAP = 1024;  // keys per warp (8 KB)
typedef unsigned long long u64;

__device__ __forceinline__ void warp_bitonic_sort(u64 *buf, int n_pow2, int lane) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (n_pow2 >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const u64 a = buf[i], b = buf[l];
                const bool up = ((i & k) == 0);
                if ((a > b) == up) {
                    buf[i] = b;
                    buf[l] = a;
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// sort buf[0..cnt) ascending; returns nothing.  Pads to a power of two with ~0.
__device__ __forceinline__ void warp_sort_prefix(u64 *buf, int cnt, int lane) {
    const int p = next_pow2(cnt < 2 ? 2 : cnt);
    for (int t = cnt + lane; t < p; t += 32) buf[t] = ~0ull;
    __syncwarp();
    warp_bitonic_sort(buf, p, lane);
}

// Collect the hybrid neighbour list of query (qx,qy,qz) into buf (sorted ascending); returns its length (<= max_nn).
__device__ __forceinline__ int warp_knn_hybrid(const Grid &g, float qx, float qy, float qz, float r2, int max_nn,
                                               u64 *buf, int lane) {
    const int cx = grid_cell((double)qx, g.ox, g.inv_h, g.nx);
    const int cy = grid_cell((double)qy, g.oy, g.inv_h, g.ny);
    const int cz = grid_cell((double)qz, g.oz, g.inv_h, g.nz);
    u64 thr = ((u64)__float_as_uint(r2)) << 32;  // d2 < r2  <=>  key < thr  (d2 >= 0)
    int cnt = 0;
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    if (x0 <= x1) {
        const int y0 = max(cy - 1, 0), y1 = min(cy + 1, g.ny - 1);
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.nz - 1);
        for (int z = z0; z <= z1; z++) {
            for (int y = y0; y <= y1; y++) {
                const long long row = ((long long)z * g.ny + y) * g.nx;
                const uint32_t b = __ldg(g.start + row + x0);
                const uint32_t e = __ldg(g.start + row + x1 + 1);
                for (uint32_t base = b; base < e; base += 32) {
                    const uint32_t k = base + lane;
                    bool pred = false;
                    u64 key = 0;
                    if (k < e) {
                        const float4 p = __ldg(g.sorted + k);
                        const float d2 = dist2f(qx, qy, qz, p.x, p.y, p.z);
                        key = (((u64)__float_as_uint(d2)) << 32) | (uint32_t)__float_as_int(p.w);
                        pred = key < thr;
                    }
                    if (cnt + 32 > KNN_CAP) {  // warp-uniform: make room
                        __syncwarp();
                        warp_sort_prefix(buf, cnt, lane);
                        if (cnt > max_nn) cnt = max_nn;
                        if (cnt == max_nn) thr = buf[max_nn - 1];
                        __syncwarp();
                        pred = pred && key < thr;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, pred);
                    if (pred) buf[cnt + __popc(m & ((1u << lane) - 1u))] = key;
                    cnt += __popc(m);
                }
            }
        }
    }
    __syncwarp();
    if (cnt > 1) warp_sort_prefix(buf, cnt, lane);
    __syncwarp();
    return cnt < max_nn ? cnt : max_nn;
}

// ---- register-resident variant for max_nn <= 32 ----------------------------------------------------------------------
// The current best keys live one per lane, sorted ascending by lane.  A 32-wide scan step with many admissible
// candidates is bitonic-sorted across the warp and merged (min(best[i], cand[31-i]) + bitonic merge); a step with
// few candidates inserts them one by one (ballot-rank + shuffle-up).  Rows are visited home row first and a row
// whose slab is provably farther than the current max_nn-th distance is skipped (same bound as grid_nn1).
__device__ __forceinline__ u64 warp_bitonic_sort32(u64 v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const u64 o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
            v = keep_min ? (o < v ? o : v) : (o > v ? o : v);
        }
    }
    return v;
}

__device__ __forceinline__ u64 warp_bitonic_merge32(u64 v, int lane) {  // v bitonic -> ascending
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const u64 o = __shfl_xor_sync(0xffffffffu, v, j);
        v = ((lane & j) == 0) ? (o < v ? o : v) : (o > v ? o : v);
    }
    return v;
}

// (dy, dz) offsets of the rows of a (2R+1)^2 block, nearest rows first; the first 9 entries are the R = 1 block
__constant__ signed char KNN_ROW_DY[25] = {0, -1, 1, 0, 0, -1, -1, 1, 1, -2, 2, 0, 0, -2, -2, 2, 2, -1, 1, -1, 1, -2, -2, 2, 2};
__constant__ signed char KNN_ROW_DZ[25] = {0, 0, 0, -1, 1, -1, 1, -1, 1, 0, 0, -2, 2, -1, 1, -1, 1, -2, -2, 2, 2, -2, 2, -2, 2};

// returns the number of neighbours (<= max_nn <= 32); *best_out = this lane's key of the ascending list.
// The grid may have cells finer than the radius (g.R = 1 or 2 rings): rows are visited nearest first, a row whose slab
// is provably farther than the current admission threshold (radius, then the max_nn-th distance) is skipped, and so
// are the cells of a row beyond that distance in x (both bounds carry a 1e-4-cell slack that dominates the fp32
// rounding involved; cells at exactly the threshold distance are kept).  On a dense cloud the list fills up in the
// first few rows of fine cells and most of the block is never touched; the result does not depend on R.
__device__ __forceinline__ int warp_knn_top32(const Grid &g, float qx, float qy, float qz, float r2, int max_nn, int lane,
                                              u64 *best_out) {
    const double fx = ((double)qx - g.ox) * g.inv_h, fy = ((double)qy - g.oy) * g.inv_h, fz = ((double)qz - g.oz) * g.inv_h;
    const int R = g.R;
    const int cx = (int)fmin(fmax(floor(fx), -1.0 - R), (double)g.nx + R);
    const int cy = (int)fmin(fmax(floor(fy), -1.0 - R), (double)g.ny + R);
    const int cz = (int)fmin(fmax(floor(fz), -1.0 - R), (double)g.nz + R);
    const u64 INF = ~0ull;
    u64 best = INF;
    u64 thr = ((u64)__float_as_uint(r2)) << 32;  // admissible: key < thr
    const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
    if (x0 <= x1) {
        const int y0 = max(cy - R, 0), y1 = min(cy + R, g.ny - 1);
        const int z0 = max(cz - R, 0), z1 = min(cz + R, g.nz - 1);
        const float fxf = (float)fmin(fmax(fx, -4.0 - R), (double)g.nx + 4.0 + R);
        const float fyf = (float)fmin(fmax(fy, -4.0 - R), (double)g.ny + 4.0 + R), fzf = (float)fmin(fmax(fz, -4.0 - R), (double)g.nz + 4.0 + R);
        const float h2 = (float)(g.h * g.h), inv_hf = (float)g.inv_h;
        const int nrows = (2 * R + 1) * (2 * R + 1);
#pragma unroll 1
        for (int o = 0; o < nrows; o++) {
            const int y = cy + KNN_ROW_DY[o], z = cz + KNN_ROW_DZ[o];
            if (y < y0 || y > y1 || z < z0 || z > z1) continue;
            const float tdist2 = __uint_as_float((uint32_t)(thr >> 32));
            if (o > 0) {
                const float ey = fmaxf(fmaxf((float)y - fyf, fyf - (float)(y + 1)), 0.0f);
                const float ez = fmaxf(fmaxf((float)z - fzf, fzf - (float)(z + 1)), 0.0f);
                const float sy = fmaxf(ey - 1e-4f, 0.0f), sz = fmaxf(ez - 1e-4f, 0.0f);
                // skip only if strictly farther than the admission threshold's distance (ties are kept)
                if ((sy * sy + sz * sz) * h2 > tdist2) continue;
            }
            // cells of the row that can hold an admissible point (in cell units, padded)
            const float rc = sqrtf(tdist2) * inv_hf * 1.0001f + 1e-4f;
            const int xa = max(x0, (int)floorf(fxf - rc)), xb = min(x1, (int)floorf(fxf + rc));
            if (xa > xb) continue;
            const long long row = ((long long)z * g.ny + y) * g.nx;
            const uint32_t b = __ldg(g.start + row + xa);
            const uint32_t e = __ldg(g.start + row + xb + 1);
            for (uint32_t base = b; base < e; base += 32) {
                const uint32_t k = base + lane;
                u64 key = INF;
                if (k < e) {
                    const float4 p = __ldg(g.sorted + k);
                    const float d2 = dist2f(qx, qy, qz, p.x, p.y, p.z);
                    key = (((u64)__float_as_uint(d2)) << 32) | (uint32_t)__float_as_int(p.w);
                }
                unsigned m = __ballot_sync(0xffffffffu, key < thr);
                if (m == 0) continue;
                if (__popc(m) > 12) {  // an insertion is ~8 dependent instructions, a sort + merge ~160
                    u64 c = key < thr ? key : INF;
                    c = warp_bitonic_sort32(c, lane);
                    const u64 rev = __shfl_sync(0xffffffffu, c, 31 - lane);
                    best = warp_bitonic_merge32(rev < best ? rev : best, lane);
                } else {
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const u64 ck = __shfl_sync(0xffffffffu, key, src);
                        if (!(ck < thr)) continue;  // the threshold may have tightened since the ballot
                        const int pos = __popc(__ballot_sync(0xffffffffu, best < ck));
                        const u64 up = __shfl_up_sync(0xffffffffu, best, 1);
                        best = lane < pos ? best : (lane == pos ? ck : up);
                        const u64 t = __shfl_sync(0xffffffffu, best, max_nn - 1);
                        if (t < thr) thr = t;
                    }
                    continue;
                }
                const u64 t = __shfl_sync(0xffffffffu, best, max_nn - 1);
                if (t < thr) thr = t;
            }
        }
    }
    *best_out = best;
    const int found = __popc(__ballot_sync(0xffffffffu, best != INF));
    return found < max_nn ? found : max_nn;
}

// ---- histogram selection for max_nn <= 32 (round 2) -------------------------------------------------------------------
// On a dense cloud (full-resolution normals: ~760 candidates in the 27 cells of a radius-sized grid, ~280 inside the
// radius, 30 wanted) the streaming list above costs ~2,350 warp instructions per query, half of them spent looking at
// candidates and the rest on 32-wide sort + merge steps.  This path
//   (1) works on a grid of HALF-radius cells (2 rings) and first looks only at the central 3x3x3 block, which contains
//       every point closer than one cell size h to the query: if max_nn of the keys seen there are closer than
//       rho = h (1 - 5e-5), the max_nn nearest neighbours are among them and the other 98 cells are never touched
//       (~150 candidates instead of ~760 on the dense cloud); otherwise the rest of the 5x5x5 block is appended;
//   (2) fetches all row ranges of a phase in ONE round trip (a lane per row) and walks them as one flat index space,
//       128 candidates per step with the 4 loads of a lane in flight together — the search is a chain of dependent
//       loads, and what bounds it is round trips per query;
//   (3) appends the in-radius keys to shared memory without ordering them, locates with a 256-bin histogram of d2 / r2
//       the bin B of the max_nn-th smallest key (on a surface the neighbour count grows like d^2: ~1 key per bin), and
//       sorts the <= 32 keys of bins <= B once across the warp.  The bin index is monotone in d2, so those keys contain the
//       max_nn smallest; the (d2, index) order of rule D2 is the key order.
// Exactly the list warp_knn_top32 returns.  Queries this path cannot take (more than KSEL_CAP keys inside the radius, more
// than 32 keys up to bin B — a plateau of exact ties —, a grid with other ring counts) return false and take it instead.
constexpr int KSEL_CAP = 512;   // in-radius keys per warp (4 KB)
constexpr int KSEL_BINS = 256;

// One phase of the candidate stream: lane l owns the range [beg, beg + len) of the cell-sorted cloud (len may be 0); the
// 32 ranges are walked as one flat index space.  scratch: 65 words of shared memory.  Appends the keys < thr to skeys,
// counts the keys < thr_in.  Returns false on overflow.
__device__ __forceinline__ bool knn_stream_ranges(const Grid &g, float qx, float qy, float qz, u64 thr, u64 thr_in, uint32_t beg,
                                                  uint32_t len, int lane, uint32_t *scratch, u64 *skeys, int &cnt, int &n_in) {
    const u64 INF = ~0ull;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return true;
    __syncwarp();
    scratch[lane] = incl - len;  // exclusive offset of range `lane`; scratch[32] = total is the sentinel
    scratch[33 + lane] = beg;
    if (lane == 0) scratch[32] = total;
    __syncwarp();
    // this lane's current range, kept in registers: its flat index only grows, so the cursor only moves forward, and the
    // common "no boundary crossed" case costs one compare (the first version re-read the tables for every candidate and
    // spent 20 % of the kernel's instructions in this walk)
    int r = 0;
    uint32_t cur_off = scratch[0], nxt_off = scratch[1], cur_beg = scratch[33];
#pragma unroll 1
    for (uint32_t base = 0; base < total; base += 128) {
        float4 p[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t f = base + 32 * i + lane;
            p[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (f < total) {
                while (nxt_off <= f) {
                    r++;
                    cur_off = nxt_off;
                    nxt_off = scratch[r + 1];
                    cur_beg = scratch[33 + r];
                }
                p[i] = __ldg(g.sorted + cur_beg + (f - cur_off));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (base + 32 * i >= total) break;  // warp-uniform
            u64 key = INF;
            if (base + 32 * i + lane < total) {
                const float d2 = dist2f(qx, qy, qz, p[i].x, p[i].y, p[i].z);
                key = (((u64)__float_as_uint(d2)) << 32) | (uint32_t)__float_as_int(p[i].w);
            }
            const bool pred = key < thr;
            const unsigned m = __ballot_sync(0xffffffffu, pred);
            if (cnt + 32 > KSEL_CAP) return false;  // warp-uniform
            if (pred) skeys[cnt + __popc(m & lt_mask)] = key;
            cnt += __popc(m);
            n_in += __popc(__ballot_sync(0xffffffffu, key < thr_in));
        }
    }
    return true;
}

__device__ __forceinline__ bool warp_knn_select32(const Grid &g, float qx, float qy, float qz, float r2, int max_nn, int lane,
                                                  u64 *skeys, uint32_t *hist, u64 *best_out, int *count_out) {
    const int R = g.R;
    if (R < 1 || R > 2) return false;
    const double fx = ((double)qx - g.ox) * g.inv_h, fy = ((double)qy - g.oy) * g.inv_h, fz = ((double)qz - g.oz) * g.inv_h;
    const int cx = (int)fmin(fmax(floor(fx), -4.0), (double)g.nx + 3.0);
    const int cy = (int)fmin(fmax(floor(fy), -4.0), (double)g.ny + 3.0);
    const int cz = (int)fmin(fmax(floor(fz), -4.0), (double)g.nz + 3.0);
    const u64 INF = ~0ull;
    const u64 thr = ((u64)__float_as_uint(r2)) << 32;  // d2 < r2  <=>  key < thr
    const unsigned lt_mask = (1u << lane) - 1u;
    // inner bound (2 rings only): a point with fp32 d2 < rho2 = h^2 (1 - 1e-4) is closer than h (1 - 4e-5) to the query and
    // therefore lies in the central 3x3x3 block (cell coordinates are fp64, exact to ~1e-13 cells)
    const float rho2 = fminf((float)(g.h * g.h * (1.0 - 1e-4)), r2);
    const u64 thr_in = R == 2 ? ((u64)__float_as_uint(rho2)) << 32 : thr;
    // row ranges: lane l < (2R+1)^2 owns row (dy, dz); S0..S3 = start[] at the x cells  cx-R | cx-1 | cx+2 | cx+R+1
    const int W = 2 * R + 1;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    bool central = false;
    if (lane < W * W) {
        const int dy = lane % W - R, dz = lane / W - R;
        const int y = cy + dy, z = cz + dz;
        central = dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1;
        const int xa = max(cx - R, 0), xb = min(cx + R, g.nx - 1);      // whole row, clamped
        const int xc = max(cx - 1, 0), xd = min(cx + 1, g.nx - 1);      // central cells, clamped
        if (y >= 0 && y < g.ny && z >= 0 && z < g.nz && xa <= xb) {
            const uint32_t *st = g.start + ((long long)z * g.ny + y) * g.nx;
            s0 = __ldg(st + xa);
            s3 = __ldg(st + xb + 1);
            if (xc <= xd) {
                s1 = __ldg(st + xc);
                s2 = __ldg(st + xd + 1);
            } else {
                s1 = s2 = s0;  // no central cell inside the grid: everything belongs to the outer phase
            }
        }
    }
    int cnt = 0, n_in = 0;
    // phase 1: the central block (1 ring: the whole block)
    if (!knn_stream_ranges(g, qx, qy, qz, thr, thr_in, R == 2 ? s1 : s0, central ? (R == 2 ? s2 - s1 : s3 - s0) : 0u, lane, hist, skeys,
                           cnt, n_in))
        return false;
    if (R == 2 && n_in < max_nn) {
        // phase 2: the rest of the 5x5x5 block — left and right parts of the central rows, whole outer rows
        if (!knn_stream_ranges(g, qx, qy, qz, thr, thr, s0, central ? s1 - s0 : s3 - s0, lane, hist, skeys, cnt, n_in)) return false;
        if (!knn_stream_ranges(g, qx, qy, qz, thr, thr, s2, central ? s3 - s2 : 0u, lane, hist, skeys, cnt, n_in)) return false;
    }
    __syncwarp();
    if (cnt <= 32) {  // sparse neighbourhood: one key per lane, one sort
        const u64 v = warp_bitonic_sort32(lane < cnt ? skeys[lane] : INF, lane);
        __syncwarp();
        *best_out = v;
        *count_out = cnt < max_nn ? cnt : max_nn;
        return true;
    }
    // cnt > 32 >= max_nn: locate the bin of the max_nn-th smallest key
    for (int b = lane; b < KSEL_BINS; b += 32) hist[b] = 0u;
    __syncwarp();
    const float scale = (float)KSEL_BINS / r2;
    for (int k = lane; k < cnt; k += 32) {
        const float d2 = __uint_as_float((uint32_t)(skeys[k] >> 32));
        atomicAdd(&hist[min(KSEL_BINS - 1, (int)(d2 * scale))], 1u);
    }
    __syncwarp();
    uint32_t h8[8], mine = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        h8[i] = hist[8 * lane + i];
        mine += h8[i];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - mine;
    // the lane whose 8 bins contain the max_nn-th key
    const bool holder = excl < (uint32_t)max_nn && incl >= (uint32_t)max_nn;
    int B = 0;
    uint32_t upto = 0;  // keys in bins <= B
    if (holder) {
        uint32_t p = excl;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const bool here = p < (uint32_t)max_nn && p + h8[i] >= (uint32_t)max_nn;
            if (here) {
                B = 8 * lane + i;
                upto = p + h8[i];
            }
            p += h8[i];
        }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, holder)) - 1;
    B = __shfl_sync(0xffffffffu, B, src);
    upto = __shfl_sync(0xffffffffu, upto, src);
    if (upto > 32u) return false;  // a plateau of (near-)ties around the max_nn-th key: the streaming list handles it
    __syncwarp();
    // gather the keys of bins <= B, one per lane (the histogram words are reused as the staging area)
    u64 *out = reinterpret_cast<u64 *>(hist);
    int sel = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int k = base + lane;
        u64 key = INF;
        bool pred = false;
        if (k < cnt) {
            key = skeys[k];
            pred = min(KSEL_BINS - 1, (int)(__uint_as_float((uint32_t)(key >> 32)) * scale)) <= B;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pred);
        if (pred) out[sel + __popc(m & lt_mask)] = key;
        sel += __popc(m);
    }
    __syncwarp();
    const u64 v = warp_bitonic_sort32(lane < sel ? out[lane] : INF, lane);
    __syncwarp();
    *best_out = v;
    *count_out = max_nn;
    return true;
}

// ---- MODE_LIST: neighbour lists to global memory ------------------------------------------------------------
// self_order: the queries ARE the indexed cloud: take them in cell order (g.sorted, .w = original index) so that
// consecutive warps probe the same rows of cells
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn_list(const float4 *__restrict__ queries, int nq, Grid g,
                                                             float r2, int max_nn, int *__restrict__ idx,
                                                             float *__restrict__ d2, int *__restrict__ cnt_out, int self_order) {
    __shared__ u64 sbuf[KNN_WARPS][KNN_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *buf = sbuf[warp];
    for (int qs = blockIdx.x * KNN_WARPS + warp; qs < nq; qs += gridDim.x * KNN_WARPS) {
        const float4 p = self_order ? __ldg(g.sorted + qs) : __ldg(queries + qs);
        const int q = self_order ? __float_as_int(p.w) : qs;
        const int c = warp_knn_hybrid(g, p.x, p.y, p.z, r2, max_nn, buf, lane);
        for (int k = lane; k < max_nn; k += 32) {
            const bool v = k < c;
            const u64 key = v ? buf[k] : 0ull;
            idx[(size_t)q * max_nn + k] = v ? (int)(uint32_t)(key & 0xffffffffull) : -1;
            d2[(size_t)q * max_nn + k] = v ? __uint_as_float((uint32_t)(key >> 32)) : 0.0f;
        }
        if (lane == 0) cnt_out[q] = c;
        __syncwarp();
    }
}

// ---- MODE_NORMAL: covariance of the neighbourhood (A.2 cumulants, sequential in neighbour order) -----------------
// cov_out: 6 doubles per query (c00 c01 c02 c11 c12 c22); the eigen-solve runs one query per thread afterwards.
// TOP32 = true: max_nn <= 32, register-resident list; shared memory only stages the 32 x 9 products of a warp (2.3 KB),
// so the kernel is limited by its 61 registers (8 CTAs / SM) instead of by the 8 KB-per-warp list buffer (6 CTAs / SM).
template <bool TOP32>
__global__ void __launch_bounds__(KNN_WARPS * 32, TOP32 ? 8 : 4) k_knn_cov(const float4 *__restrict__ pts, int n, Grid g, float r2,
                                                                            int max_nn, double *__restrict__ cov_out,
                                                                            unsigned int *__restrict__ stats) {
    __shared__ u64 sbuf[KNN_WARPS][TOP32 ? KSEL_CAP : KNN_CAP];
    __shared__ uint32_t shist[TOP32 ? KNN_WARPS : 1][KSEL_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *buf = sbuf[warp];
    // Queries are taken in CELL order (the grid's sorted copy; .w = original index): consecutive warps then probe the
    // same rows of cells, which the input order (arbitrary) does not give (L1 hit rate 6 % before)
    for (int qs = blockIdx.x * KNN_WARPS + warp; qs < n; qs += gridDim.x * KNN_WARPS) {
        const float4 p = __ldg(g.sorted + qs);
        const int q = __float_as_int(p.w);
        // lane l < 9 accumulates cumulant l = A*B with A, B in {1, x, y, z}; operands are chosen with fp32 value
        // selects (no divergent switch) and x*1.0 is exact, so lanes 0..2 still add the bare coordinate
        const int ia = lane < 3 ? lane + 1 : (lane < 6 ? 1 : (lane < 8 ? 2 : 3));
        const int ib = lane < 3 ? 0 : (lane < 6 ? lane - 2 : (lane < 8 ? lane - 4 : 3));
        int c;
        double cu = 0.0;
        if (TOP32) {
            u64 mine;
            const bool fast = warp_knn_select32(g, p.x, p.y, p.z, r2, max_nn, lane, buf, shist[warp], &mine, &c);
            if (!fast) c = warp_knn_top32(g, p.x, p.y, p.z, r2, max_nn, lane, &mine);
            if (stats && lane == 0) atomicAdd(stats + (fast ? 0 : 1), 1u);  // PCR_DEBUG only
            // lane k fetches neighbour k (all gathers in flight together) and writes its nine products to shared
            // memory; lane l < 9 then adds column l in neighbour order (the reference's sequential cumulants)
            double *prod = reinterpret_cast<double *>(buf);
            if (lane < c) {
                const float4 pk = __ldg(pts + (int)(uint32_t)(mine & 0xffffffffull));
                const double x = pk.x, y = pk.y, z = pk.z;
                double *d = prod + lane * 9;
                d[0] = x; d[1] = y; d[2] = z;
                d[3] = x * x; d[4] = x * y; d[5] = x * z;
                d[6] = y * y; d[7] = y * z; d[8] = z * z;
            }
            __syncwarp();
            if (c >= 3 && lane < 9) {
                for (int k = 0; k < c; k++) cu = cu + prod[k * 9 + lane];
                cu = cu / (double)c;
            }
            __syncwarp();
        } else {
            c = warp_knn_hybrid(g, p.x, p.y, p.z, r2, max_nn, buf, lane);
            if (c >= 3 && lane < 9) {
                for (int k = 0; k < c; k++) {
                    const float4 pj = __ldg(pts + (int)(uint32_t)(buf[k] & 0xffffffffull));
                    const float af = ia == 1 ? pj.x : (ia == 2 ? pj.y : pj.z);
                    const float bf = ib == 0 ? 1.0f : (ib == 1 ? pj.x : (ib == 2 ? pj.y : pj.z));
                    cu = cu + (double)af * (double)bf;
                }
                cu = cu / (double)c;
            }
        }
        double m[9];
#pragma unroll
        for (int i = 0; i < 9; i++) m[i] = __shfl_sync(0xffffffffu, cu, i);
        if (lane == 0) {
            double *o = cov_out + (size_t)q * 6;
            if (c < 3) {
                o[0] = 1.0; o[1] = 0.0; o[2] = 0.0; o[3] = 1.0; o[4] = 0.0; o[5] = 1.0;
            } else {
                o[0] = m[3] - m[0] * m[0];
                o[1] = m[4] - m[0] * m[1];
                o[2] = m[5] - m[0] * m[2];
                o[3] = m[6] - m[1] * m[1];
                o[4] = m[7] - m[1] * m[2];
                o[5] = m[8] - m[2] * m[2];
            }
        }
        __syncwarp();
    }
}

// ---- robust symmetric 3x3 eigen-solver (A.2 FastEigen3x3 restated) ------------------------------------------------
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3_cross(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double d3_dot(D3 a, D3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ D3 d3_scale(D3 a, double s) { return D3{a.x * s, a.y * s, a.z * s}; }

__device__ D3 eigvec0(const double *A, double ev) {
    const D3 r0{A[0] - ev, A[1], A[2]};
    const D3 r1{A[1], A[3] - ev, A[4]};
    const D3 r2{A[2], A[4], A[5] - ev};
    const D3 c01 = d3_cross(r0, r1), c02 = d3_cross(r0, r2), c12 = d3_cross(r1, r2);
    const double d0 = d3_dot(c01, c01), d1 = d3_dot(c02, c02), d2 = d3_dot(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    if (imax == 0) return d3_scale(c01, 1.0 / sqrt(d0));
    if (imax == 1) return d3_scale(c02, 1.0 / sqrt(d1));
    return d3_scale(c12, 1.0 / sqrt(d2));
}

__device__ D3 eigvec1(const double *A, D3 e0, double ev1) {
    D3 U, V;
    if (fabs(e0.x) > fabs(e0.y)) {
        const double inv = 1.0 / sqrt(e0.x * e0.x + e0.z * e0.z);
        U = D3{-e0.z * inv, 0.0, e0.x * inv};
    } else {
        const double inv = 1.0 / sqrt(e0.y * e0.y + e0.z * e0.z);
        U = D3{0.0, e0.z * inv, -e0.y * inv};
    }
    V = d3_cross(e0, U);
    const D3 AU{(A[0] * U.x + A[1] * U.y) + A[2] * U.z, (A[1] * U.x + A[3] * U.y) + A[4] * U.z,
                (A[2] * U.x + A[4] * U.y) + A[5] * U.z};
    const D3 AV{(A[0] * V.x + A[1] * V.y) + A[2] * V.z, (A[1] * V.x + A[3] * V.y) + A[4] * V.z,
                (A[2] * V.x + A[4] * V.y) + A[5] * V.z};
    double m00 = d3_dot(U, AU) - ev1;
    double m01 = d3_dot(U, AV);
    double m11 = d3_dot(V, AV) - ev1;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        const double mx = a00 > a01 ? a00 : a01;
        if (mx > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1.0 / sqrt(1.0 + m00 * m00); m00 *= m01; }
            return D3{m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z};
        }
        return U;
    } else {
        const double mx = a11 > a01 ? a11 : a01;
        if (mx > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m11; }
            else { m11 /= m01; m01 = 1.0 / sqrt(1.0 + m11 * m11); m11 *= m01; }
            return D3{m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z};
        }
        return U;
    }
}

__device__ D3 fast_eigen3x3(const double *C) {
    double A[6];
    double mc = C[0];
#pragma unroll
    for (int i = 1; i < 6; i++) if (C[i] > mc) mc = C[i];
    if (mc == 0.0) return D3{0, 0, 0};
#pragma unroll
    for (int i = 0; i < 6; i++) A[i] = C[i] / mc;
    const double norm = (A[1] * A[1] + A[2] * A[2]) + A[4] * A[4];
    if (norm > 0.0) {
        const double q = ((A[0] + A[3]) + A[5]) / 3.0;
        const double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        const double p = sqrt((((b00 * b00 + b11 * b11) + b22 * b22) + norm * 2.0) / 6.0);
        const double c00 = b11 * b22 - A[4] * A[4];
        const double c01 = A[1] * b22 - A[4] * A[2];
        const double c02 = A[1] * A[4] - b11 * A[2];
        const double det = ((b00 * c00 - A[1] * c01) + A[2] * c02) / ((p * p) * p);
        double half_det = det * 0.5;
        if (half_det < -1.0) half_det = -1.0;
        if (half_det > 1.0) half_det = 1.0;
        const double angle = pcr_acos(half_det) / 3.0;
        const double two_thirds_pi = 2.09439510239319549;
        const double beta2 = pcr_cos(angle) * 2.0;
        const double beta0 = pcr_cos(angle + two_thirds_pi) * 2.0;
        const double beta1 = -(beta0 + beta2);
        const double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (half_det >= 0.0) {
            const D3 v2 = eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            const D3 v1 = eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return d3_cross(v1, v2);
        } else {
            const D3 v0 = eigvec0(A, e0);
            if (e0 < e1 && e0 < e2) return v0;
            const D3 v1 = eigvec1(A, v0, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return d3_cross(v0, v1);
        }
    }
    D3 r{0, 0, 1};
    if (C[0] < C[3] && C[0] < C[5]) { r.x = 1; r.z = 0; }
    else if (C[3] < C[0] && C[3] < C[5]) { r.y = 1; r.z = 0; }
    return r;
}

__global__ void __launch_bounds__(128) k_normals_solve(const double *__restrict__ cov, int n, float4 *__restrict__ normals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double C[6];
#pragma unroll
    for (int k = 0; k < 6; k++) C[k] = cov[(size_t)i * 6 + k];
    D3 nr = fast_eigen3x3(C);
    const double len = sqrt(d3_dot(nr, nr));
    if (len == 0.0 || isnan(len)) nr = D3{0, 0, 1};
    normals[i] = make_float4((float)nr.x, (float)nr.y, (float)nr.z, 0.0f);
}

// ---- FPFH (A.4) -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_features(float4 p1f, float4 n1f, float4 p2f, float4 n2f, double *f) {
    D3 n1{n1f.x, n1f.y, n1f.z}, n2{n2f.x, n2f.y, n2f.z};
    D3 d{(double)p2f.x - (double)p1f.x, (double)p2f.y - (double)p1f.y, (double)p2f.z - (double)p1f.z};
    f[0] = f[1] = f[2] = f[3] = 0.0;
    const double len = sqrt(d3_dot(d, d));
    if (len == 0.0) return;
    f[3] = len;
    const double a1 = d3_dot(n1, d) / len;
    const double a2 = d3_dot(n2, d) / len;
    if (pcr_acos(fabs(a1)) > pcr_acos(fabs(a2))) {
        const D3 t = n1; n1 = n2; n2 = t;
        d.x = -d.x; d.y = -d.y; d.z = -d.z;
        f[2] = -a2;
    } else {
        f[2] = a1;
    }
    D3 v = d3_cross(d, n1);
    const double vn = sqrt(d3_dot(v, v));
    if (vn == 0.0) { f[0] = f[1] = f[2] = f[3] = 0.0; return; }
    v.x /= vn; v.y /= vn; v.z /= vn;
    const D3 w = d3_cross(n1, v);
    f[1] = d3_dot(v, n2);
    f[0] = pcr_atan2(d3_dot(w, n2), d3_dot(n1, n2));
}

__device__ __forceinline__ int clamp_bin(double x) {
    int h = (int)floor(x);
    if (h < 0) h = 0;
    if (h >= 11) h = 10;
    return h;
}

// one warp per point: lanes over neighbours 1..c-1 (position 0, the query itself, is skipped)
__global__ void __launch_bounds__(128) k_spfh(const float4 *__restrict__ pts, const float4 *__restrict__ nrm, int n,
                                              const int *__restrict__ idx, const int *__restrict__ cnt, int max_nn,
                                              double *__restrict__ spfh, const float4 *__restrict__ order) {
    __shared__ int hist[4][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int is = blockIdx.x * 4 + warp; is < n; is += gridDim.x * 4) {
        const int i = __float_as_int(__ldg(order + is).w);  // points in cell order: neighbouring warps gather the same rows
        hist[warp][lane] = 0;
        if (lane == 0) hist[warp][32] = 0;
        __syncwarp();
        const int c = cnt[i];
        const float4 p1 = __ldg(pts + i), n1 = __ldg(nrm + i);
        for (int k = 1 + lane; k < c; k += 32) {
            const int j = idx[(size_t)i * max_nn + k];
            double f[4];
            pair_features(p1, n1, __ldg(pts + j), __ldg(nrm + j), f);
            atomicAdd(&hist[warp][clamp_bin(11.0 * (f[0] + PCR_PI) / (2.0 * PCR_PI))], 1);
            atomicAdd(&hist[warp][11 + clamp_bin(11.0 * (f[1] + 1.0) * 0.5)], 1);
            atomicAdd(&hist[warp][22 + clamp_bin(11.0 * (f[2] + 1.0) * 0.5)], 1);
        }
        __syncwarp();
        const double inc = c > 1 ? 100.0 / (double)(c - 1) : 0.0;
        for (int j = lane; j < 33; j += 32) {
            const int h = hist[warp][j];
            double v = 0.0;
            for (int t = 0; t < h; t++) v = v + inc;  // the reference adds hist_incr once per neighbour
            spfh[(size_t)i * 33 + j] = v;
        }
        __syncwarp();
    }
}

// One warp per point.  FPFH[j] = (sum_k SPFH[nb_k][j] / d2_k) * (100 / S_b) + SPFH[i][j]  with the block normaliser
// S_b of rule D8.  Lane j owns bin j (lane 0 also bin 32) and adds the quotients of its bin in neighbour order; a chunk of
// FPFH_CHUNK neighbours is gathered at once — lane j reads element j of every neighbour's SPFH row (one coalesced 256-byte
// row per load instruction, all of them in flight before the first is used), lane k reads bin 32 of neighbour k — so a
// chunk costs ONE L2 round trip.  (Until round 2 the elements went through a shared-memory staging array one dependent
// gather at a time: ~10k cycles per chunk on a kernel that is a single wave of warps, i.e. a pure latency chain.)
constexpr int FPFH_CHUNK = 8;
constexpr int FPFH_TILE = 128;  // neighbours staged per round (4 per lane); max_nn = 100 fits in one
__global__ void __launch_bounds__(128) k_fpfh(int n, const int *__restrict__ idx, const float *__restrict__ d2,
                                              const int *__restrict__ cnt, int max_nn,
                                              const double *__restrict__ spfh, float *__restrict__ out,
                                              const float4 *__restrict__ order) {
    // per warp: distance, its reciprocal, neighbour index and the finished quotient of bin 32, for one tile of neighbours
    __shared__ double s_b[4][FPFH_TILE], s_r[4][FPFH_TILE], s_q32[4][FPFH_TILE];
    __shared__ int s_nb[4][FPFH_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int is = blockIdx.x * 4 + warp; is < n; is += gridDim.x * 4) {
        const int i = __float_as_int(__ldg(order + is).w);  // cell order, as in k_spfh
        const int c = cnt[i];
        double F0 = 0.0, F1 = 0.0;  // bins `lane` and (lane 0 only) 32
        double sum = 0.0;           // lanes 0..2: normaliser of block `lane`
        for (int t0 = 1; t0 < c; t0 += FPFH_TILE) {
            const int nt = min(FPFH_TILE, c - t0);
            // One correctly rounded reciprocal per neighbour, then every quotient SPFH/d2 by Markstein's sequence
            // q0 = a*r, rem = fma(-b, q0, a), q = fma(rem, r, q0): with r = RN(1/b) this IS the correctly rounded a/b
            // unless b's significand is all ones — impossible here, b is an fp32 value widened to fp64.  33 divisions
            // per neighbour become 1 division + 33 x 3 multiply-adds with identical bits (checked by the parity tests).
            // Staging: every lane takes 4 neighbours of the tile — index, distance, reciprocal (the divisions of the whole
            // tile run side by side instead of one per chunk) and bin 32, which lane 0 will only have to add.
            __syncwarp();
            int nb4[FPFH_TILE / 32];
            double b4[FPFH_TILE / 32];
#pragma unroll
            for (int g = 0; g < FPFH_TILE / 32; g++) {
                const int k = lane + 32 * g;
                nb4[g] = k < nt ? idx[(size_t)i * max_nn + t0 + k] : 0;
                b4[g] = k < nt ? (double)d2[(size_t)i * max_nn + t0 + k] : 0.0;
            }
#pragma unroll
            for (int g = 0; g < FPFH_TILE / 32; g++) {
                const int k = lane + 32 * g;
                if (k < nt) {
                    const double a32 = __ldg(spfh + (size_t)nb4[g] * 33 + 32);
                    const double r = b4[g] == 0.0 ? 0.0 : 1.0 / b4[g];
                    const double q0 = a32 * r;
                    const double rem = __fma_rn(-b4[g], q0, a32);
                    s_nb[warp][k] = nb4[g];
                    s_b[warp][k] = b4[g];
                    s_r[warp][k] = r;
                    // a zero distance (duplicate point) is skipped by the reference; adding +0.0 is the same bits
                    s_q32[warp][k] = b4[g] == 0.0 ? 0.0 : __fma_rn(rem, r, q0);
                }
            }
            __syncwarp();
            for (int k0 = 0; k0 < nt; k0 += FPFH_CHUNK) {
                const int nk = min(FPFH_CHUNK, nt - k0);
                double a[FPFH_CHUNK];
#pragma unroll
                for (int kk = 0; kk < FPFH_CHUNK; kk++)
                    a[kk] = kk < nk ? __ldg(spfh + (size_t)s_nb[warp][k0 + kk] * 33 + lane) : 0.0;
#pragma unroll
                for (int kk = 0; kk < FPFH_CHUNK; kk++) {
                    if (kk < nk) {
                        const double b = s_b[warp][k0 + kk], r = s_r[warp][k0 + kk];
                        const double q0 = a[kk] * r;
                        const double rem = __fma_rn(-b, q0, a[kk]);
                        F0 = F0 + (b == 0.0 ? 0.0 : __fma_rn(rem, r, q0));
                        F1 = F1 + s_q32[warp][k0 + kk];
                    }
                }
            }
        }
        // D8: the normaliser of block b is the sum of its 11 accumulated bins in bin order (lane b gathers them by
        // shuffles; bin 32 lives in lane 0's F1).  A running sum over (neighbour, bin), as Open3D keeps it, is 1100
        // dependent additions on 3 lanes per point — it was 34 % of this kernel's instructions.
        const double F32 = __shfl_sync(0xffffffffu, F1, 0);
#pragma unroll
        for (int j = 0; j < 11; j++) {
            double vv = __shfl_sync(0xffffffffu, F0, (11 * lane + j) & 31);
            if (lane == 2 && j == 10) vv = F32;
            sum = sum + vv;
        }
        if (lane < 3 && sum != 0.0) sum = 100.0 / sum;
        const double s0 = __shfl_sync(0xffffffffu, sum, 0);
        const double s1 = __shfl_sync(0xffffffffu, sum, 1);
        const double s2 = __shfl_sync(0xffffffffu, sum, 2);
        const double *hi = spfh + (size_t)i * 33;
        if (c > 1) {
            const double sj = lane < 11 ? s0 : (lane < 22 ? s1 : s2);
            out[(size_t)i * 33 + lane] = (float)(F0 * sj + hi[lane]);
            if (lane == 0) out[(size_t)i * 33 + 32] = (float)(F1 * s2 + hi[32]);
        } else {
            out[(size_t)i * 33 + lane] = 0.0f;
            if (lane == 0) out[(size_t)i * 33 + 32] = 0.0f;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// A context that works NEXT TO a critical path (pcr_align's helper) launches short-lived CTAs, a few queries each, instead
// of a persistent grid: CTA slots then free up all the time and the block scheduler hands them to the higher-priority
// stream first — a persistent grid never retires a CTA, so stream priorities had nothing to act on and the descriptor
// matching ran at half speed beside the full-resolution normals (tools/gpu_timeline.py).
static int knn_blocks(pcr_ctx *ctx, int nq) {
    if (ctx->yielding) return div_up(nq, KNN_WARPS * ctx->yielding);
    return min(div_up(nq, KNN_WARPS), ctx->sm_count * 16);
}

int pcr_knn_impl(pcr_ctx *ctx, const float4 *pts, int n, const float4 *q, int nq, double radius, int max_nn, int *idx,
                 float *d2, int *cnt) {
    if (!(radius > 0.0) || max_nn < 1 || max_nn > KNN_CAP / 2)
        return pcr_fail(ctx, PCR_ERR_INVALID, "knn: radius must be > 0 and 1 <= max_nn <= %d", KNN_CAP / 2);
    if (nq == 0) return PCR_OK;
    if (n == 0) {
        PCR_CUDA(cudaMemsetAsync(idx, 0xff, sizeof(int) * (size_t)nq * max_nn, ctx->stream));
        PCR_CUDA(cudaMemsetAsync(d2, 0, sizeof(float) * (size_t)nq * max_nn, ctx->stream));
        PCR_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)nq, ctx->stream));
        return PCR_OK;
    }
    Grid g;
    PCR_TRY(pcr_grid_build(ctx, pts, n, radius, nullptr, nullptr, &g));
    KScope ks(ctx, KC_KNN_LIST, 16.0 * n + 16.0 * nq + 8.0 * (double)nq * max_nn);
    k_knn_list<<<knn_blocks(ctx, nq), KNN_WARPS * 32, 0, ctx->stream>>>(q, nq, g, (float)(radius * radius), max_nn, idx,
                                                                        d2, cnt, (q == pts && nq == n) ? 1 : 0);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

int pcr_normals_impl(pcr_ctx *ctx, const float4 *pts, int n, double radius, int max_nn, float4 *normals) {
    if (!(radius > 0.0) || max_nn < 1 || max_nn > KNN_CAP / 2)
        return pcr_fail(ctx, PCR_ERR_INVALID, "normals: radius must be > 0 and 1 <= max_nn <= %d", KNN_CAP / 2);
    if (n == 0) return PCR_OK;
    Grid g;
    // Half-radius cells (2 rings): warp_knn_select32 usually finds the max_nn nearest neighbours inside the central 3x3x3
    // block and never touches the other 98 cells.  (With the round-1 streaming list alone 2 rings were slower — one round
    // trip per row, 25 rows; the new path fetches all row ranges at once.)  PCR_KNN_RINGS=1 keeps radius-sized cells.
    static const int rings = getenv("PCR_KNN_RINGS") ? atoi(getenv("PCR_KNN_RINGS")) : 2;
    PCR_TRY(pcr_grid_build_rings(ctx, pts, n, radius, (max_nn <= 32 && rings == 2) ? 2 : 1, nullptr, nullptr, &g));
    PCR_ALLOC(cov, double, (size_t)n * 6);
    {
        KScope ks(ctx, KC_KNN_COV, 32.0 * n + 48.0 * n);
        unsigned int *stats = nullptr;
        static const bool dbg = getenv("PCR_DEBUG") != nullptr;
        if (dbg) {
            stats = arena<unsigned int>(ctx, 2);
            if (stats) cudaMemsetAsync(stats, 0, 8, ctx->stream);
        }
        if (max_nn <= 32)
            k_knn_cov<true><<<knn_blocks(ctx, n), KNN_WARPS * 32, 0, ctx->stream>>>(pts, n, g, (float)(radius * radius), max_nn, cov, stats);
        else
            k_knn_cov<false><<<knn_blocks(ctx, n), KNN_WARPS * 32, 0, ctx->stream>>>(pts, n, g, (float)(radius * radius), max_nn, cov, stats);
        PCR_LAUNCHED();
        if (stats) {
            unsigned int h[2] = {0, 0};
            cudaMemcpyAsync(h, stats, 8, cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            fprintf(stderr, "[pcr] knn_cov n=%d max_nn=%d: histogram-select %u, streaming list %u\n", n, max_nn, h[0], h[1]);
        }
    }
    {
        KScope ks(ctx, KC_NORMALS_SOLVE, 48.0 * n + 16.0 * n);
        k_normals_solve<<<div_up(n, 128), 128, 0, ctx->stream>>>(cov, n, normals);
        PCR_LAUNCHED();
    }
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

int pcr_fpfh_impl(pcr_ctx *ctx, const float4 *pts, const float4 *nrm, int n, double radius, int max_nn, float *out) {
    if (!(radius > 0.0) || max_nn < 1 || max_nn > KNN_CAP / 2)
        return pcr_fail(ctx, PCR_ERR_INVALID, "fpfh: radius must be > 0 and 1 <= max_nn <= %d", KNN_CAP / 2);
    if (n == 0) return PCR_OK;
    Grid g;
    PCR_TRY(pcr_grid_build(ctx, pts, n, radius, nullptr, nullptr, &g));
    PCR_ALLOC(idx, int, (size_t)n * max_nn);
    PCR_ALLOC(d2, float, (size_t)n * max_nn);
    PCR_ALLOC(cnt, int, (size_t)n);
    PCR_ALLOC(spfh, double, (size_t)n * 33);
    {
        KScope ks(ctx, KC_KNN_LIST, 32.0 * n + 8.0 * (double)n * max_nn);
        k_knn_list<<<knn_blocks(ctx, n), KNN_WARPS * 32, 0, ctx->stream>>>(pts, n, g, (float)(radius * radius), max_nn,
                                                                           idx, d2, cnt, 1);
        PCR_LAUNCHED();
    }
    const int blocks = min(div_up(n, 4), ctx->sm_count * 16);
    {
        KScope ks(ctx, KC_SPFH, 32.0 * n + 4.0 * (double)n * max_nn + 264.0 * n);
        k_spfh<<<blocks, 128, 0, ctx->stream>>>(pts, nrm, n, idx, cnt, max_nn, spfh, g.sorted);
        PCR_LAUNCHED();
    }
    {
        KScope ks(ctx, KC_FPFH, 264.0 * n + 8.0 * (double)n * max_nn + 132.0 * n);
        k_fpfh<<<blocks, 128, 0, ctx->stream>>>(n, idx, d2, cnt, max_nn, spfh, out, g.sorted);
        PCR_LAUNCHED();
    }
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
