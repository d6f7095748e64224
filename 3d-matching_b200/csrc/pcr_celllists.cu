// pcr_celllists.cu — per-fine-cell candidate lists over a target cloud: the search structure of k_ransac_validate (the
// default since round 2; PCR_VAL_LISTS=0 keeps the 27-cell grid walk, which also serves cells flagged 15).
//
// Motivation (profiles/r1_summary.md): the validation of a RANSAC survivor runs a radius-limited nearest-neighbour query
// for every source point — tens of millions of queries per run, ~300 instructions each on the grid walk (27 coarse cells,
// ~20 candidates).  With lists a query computes its fine cell, reads one header word and tests the 3-6 points that can be
// nearest to ANY position inside that cell.  Measured on B200: k_ransac_validate 0.84 -> 0.40 ms per alignment, 10M
// hypotheses 162 -> 338 M hyp/s, results identical.
//
// Exactness.  For the cube C of a fine cell (side c, enlarged by 2^-20 for the rounding of the cell mapping) let
//   dmin(t, C) / dmax(t, C) = the smallest / largest distance from target point t to C,
//   bound = min( min_t' dmax(t', C), r ) * (1 + 1e-5).
// The list is K(C) = { t : dmin(t, C) <= bound }.  For a query q in C whose nearest point (by the fp32 rule D1, ties by
// index, D2) is t*, and the minimiser t' of dmax:  dmin(t*, C) <= |q - t*| <= |q - t'| (1 + 5e-7) <= dmax(t', C) (1 + 5e-7),
// and only points closer than r matter; 5e-7 bounds the relative rounding of two fp32 squared distances.  So t* is on
// the list, every point that ties it is too, and the final (d2, index) key over the list returns exactly what
// grid_nn1 returns.  Cells whose list would exceed 14 entries, or that do not fit the item pool, are flagged and fall
// back to the full search.  The NumPy prototype tools/proto/cell_candidate_lists.py and the host check
// tests/c/celllists_host_check.cu (the rule as stated by celllists_build_cell in pcr_celllists.cuh, against brute force)
// verify the claim on the CPU; tests/test_gpu_ransac_lists.py and the 10M-hypothesis oracle golden verify the device build.
#include "pcr_common.cuh"
#include "pcr_celllists.cuh"

namespace {

// POINT-DRIVEN build (round 2).  The cell-driven kernel of round 1 (one thread per fine cell, two sweeps over the points of
// the 3-4 coarse cells per axis the cell can reach) evaluated ~30 M (cell, point) pairs in fp64 at 11-12 of 32 lanes:
// 186 us on the bench pair, as much as the validation it accelerates.  A target point only matters to the fine cells whose
// cube comes within R = r (1 + 1e-5) of it — a window of W^3 cells, W = floor(2R / c) + 2 (8 for c = r / 3) — so the build
// walks (point, window cell) pairs instead: ~4.7 M pairs, three light passes, no divergence to speak of.
//   pass 0  bound[C] = min over the points t with dmin(t, C) <= R of dmax(t, C)^2          (64-bit atomicMin on the fp64 bits)
//   pass 1  cnt[C]   = #{ t : dmin(t, C)^2 <= min(bound[C], r^2) (1 + 1e-5)^2 }            (atomicAdd)
//   alloc   head[C]  = 0 | 15 (more than PCR_LIST_MAX entries, or pool full) | (offset << 4) | cnt
//   pass 2  the same test again; the points are written behind the cell's offset
// Same lists as celllists_build_cell (pcr_celllists.cuh, the host-checked statement of the rule): its `best` also runs over
// points with dmin > R, but those have dmax > R >= r and cannot lower min(best, r^2).  All cube distances in fp64, same
// formulas.  The ORDER of a list depends on the atomics; the query result does not (the (d2, index) key is order-free).
struct FineLattice {
    double fox, foy, foz, c, half, R, r2lim;  // r2lim = r^2
    int fnx, fny, fnz, W;
};

// one block per target point; thread l walks the window cells l, l + 256, ...
template <int PASS>
__global__ void __launch_bounds__(256) k_celllists_pairs(const float4 *__restrict__ pts, int n, FineLattice L,
                                                         unsigned long long *__restrict__ bound, uint32_t *cnt,
                                                         const uint32_t *head, uint32_t *__restrict__ fill,
                                                         float4 *__restrict__ items) {
    __shared__ int s_base[3];
    const float4 p = __ldg(pts + blockIdx.x);
    const double px = p.x, py = p.y, pz = p.z;
    if (threadIdx.x < 3) {
        // first window cell along the axis: the lowest cell whose (2^-20-enlarged) cube can be within R of the point
        const double reach = L.R + 1e-5 * L.c;
        const double v = threadIdx.x == 0 ? px - L.fox : (threadIdx.x == 1 ? py - L.foy : pz - L.foz);
        s_base[threadIdx.x] = (int)floor((v - reach) / L.c);
    }
    __syncthreads();
    const int bx = s_base[0], by = s_base[1], bz = s_base[2];
    const int W = L.W, W3 = W * W * W;
    const double RR = L.R * L.R;
    for (int l = threadIdx.x; l < W3; l += 256) {
        int a, b, cc;
        if (W == 8) { a = l & 7; b = (l >> 3) & 7; cc = l >> 6; }
        else { a = l % W; b = (l / W) % W; cc = l / (W * W); }
        const int ix = bx + a, iy = by + b, iz = bz + cc;
        if (ix < 0 || iy < 0 || iz < 0 || ix >= L.fnx || iy >= L.fny || iz >= L.fnz) continue;
        const double ex = fabs(px - (L.fox + ((double)ix + 0.5) * L.c)), ey = fabs(py - (L.foy + ((double)iy + 0.5) * L.c)),
                     ez = fabs(pz - (L.foz + ((double)iz + 0.5) * L.c));
        const double nx = fmax(ex - L.half, 0.0), ny = fmax(ey - L.half, 0.0), nz = fmax(ez - L.half, 0.0);
        const double dmin2 = (nx * nx + ny * ny) + nz * nz;
        if (dmin2 > RR) continue;  // cannot be on the list of this cell (lim <= r^2 (1 + 1e-5)^2 = R^2), nor lower its bound below r^2
        const int cell = (iz * L.fny + iy) * L.fnx + ix;
        if (PASS == 0) {
            const double fx = ex + L.half, fy = ey + L.half, fz = ez + L.half;
            atomicMin(bound + cell, (unsigned long long)__double_as_longlong((fx * fx + fy * fy) + fz * fz));  // positive doubles order as integers
        } else {
            const unsigned long long bb = bound[cell];
            if (bb == ~0ull) continue;
            const double lim = fmin(__longlong_as_double((long long)bb), L.r2lim) * ((1.0 + 1e-5) * (1.0 + 1e-5));
            if (!(dmin2 <= lim)) continue;
            if (PASS == 1) {
                atomicAdd(cnt + cell, 1u);
            } else {
                const uint32_t h = head[cell];
                if ((h & 15u) == 15u || h == 0u) continue;
                items[(h >> 4) + atomicAdd(fill + cell, 1u)] = p;
            }
        }
    }
}

// cnt (in head[]) -> header word; list space is claimed with one atomicAdd per non-empty cell
__global__ void __launch_bounds__(256) k_celllists_alloc(uint32_t *__restrict__ head, long long ncell, unsigned int *__restrict__ total,
                                                         unsigned int cap) {
    const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
    const uint32_t c = id < ncell ? head[id] : 0u;
    const bool want = c != 0u && c <= (uint32_t)PCR_LIST_MAX;
    // one atomicAdd per warp (200k same-address atomics took 129 us): inclusive scan of the requests over the lanes
    const int lane = threadIdx.x & 31;
    uint32_t incl = want ? c : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t sum = __shfl_sync(0xffffffffu, incl, 31);
    unsigned int base = 0;
    if (lane == 31 && sum) base = atomicAdd(total, sum);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (c == 0u) return;
    uint32_t h = 15u;
    if (want) {
        const unsigned int off = base + incl - c;
        if (off + c <= cap) h = (off << 4) | c;
    }
    head[id] = h;
}

}  // namespace

// Builds the lists in the context's arena (they live until the exported call returns).  `*ok` = false when the lattice
// would be too large; the caller then keeps the full search.
int pcr_celllists_build(pcr_ctx *ctx, const Grid &g, double r, int div, CellLists *out, bool *ok) {
    *ok = false;
    if (!(r > 0.0) || div < 1 || g.n <= 0) return PCR_OK;
    const CellListsDims d = celllists_dims(g, r, div);
    const double c = d.c, pad = d.pad;
    const double *fn = d.fn;
    if (fn[0] * fn[1] * fn[2] > (double)(1 << 25)) return PCR_OK;  // 32M header words = 128 MB: beyond that, not worth it
    const int fnx = (int)fn[0], fny = (int)fn[1], fnz = (int)fn[2];
    const size_t ncell = (size_t)fnx * fny * fnz;
    // measured on the bench pair (tests/c/celllists_host_check.cu): 144 items per point at div 2, 305 at div 3
    const double cap_d = fmin((40.0 * div * div + 32.0) * (double)g.n + 4096.0, (double)(1u << 27));
    const unsigned int cap = (unsigned int)cap_d;
    PCR_ALLOC(head, uint32_t, ncell);
    PCR_ALLOC(items, float4, (size_t)cap);
    PCR_ALLOC(total, unsigned int, 1);
    PCR_ALLOC(fill, uint32_t, ncell);
    PCR_ALLOC(bound, unsigned long long, ncell);
    FineLattice L;
    L.fox = g.ox - pad; L.foy = g.oy - pad; L.foz = g.oz - pad;
    L.c = c;
    L.half = 0.5 * c * (1.0 + 9.5367431640625e-07);
    L.R = r * (1.0 + 1e-5);
    L.r2lim = r * r;
    L.fnx = fnx; L.fny = fny; L.fnz = fnz;
    L.W = (int)floor((2.0 * L.R + 2e-5 * c) / c) + 2;
    if (ncell > (size_t)0x7fffffff) return PCR_OK;  // 32-bit cell ids
    KScope ks(ctx, KC_GRID_BUILD, 16.0 * (double)ncell + 16.0 * (double)g.n, 4);
    PCR_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned int), ctx->stream));
    PCR_CUDA(cudaMemsetAsync(head, 0, sizeof(uint32_t) * ncell, ctx->stream));
    PCR_CUDA(cudaMemsetAsync(fill, 0, sizeof(uint32_t) * ncell, ctx->stream));
    PCR_CUDA(cudaMemsetAsync(bound, 0xff, sizeof(unsigned long long) * ncell, ctx->stream));
    const int pb = g.n;
    k_celllists_pairs<0><<<pb, 256, 0, ctx->stream>>>(g.sorted, g.n, L, bound, head, head, fill, items);
    PCR_LAUNCHED();
    k_celllists_pairs<1><<<pb, 256, 0, ctx->stream>>>(g.sorted, g.n, L, bound, head, head, fill, items);
    PCR_LAUNCHED();
    k_celllists_alloc<<<div_up((long long)ncell, 256), 256, 0, ctx->stream>>>(head, (long long)ncell, total, cap);
    PCR_LAUNCHED();
    k_celllists_pairs<2><<<pb, 256, 0, ctx->stream>>>(g.sorted, g.n, L, bound, head, head, fill, items);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    out->head = head;
    out->items = items;
    out->ox = g.ox - pad;
    out->oy = g.oy - pad;
    out->oz = g.oz - pad;
    out->inv_c = 1.0 / c;
    out->nx = fnx;
    out->ny = fny;
    out->nz = fnz;
    out->cap = cap;
    *ok = true;
    return PCR_OK;
}
