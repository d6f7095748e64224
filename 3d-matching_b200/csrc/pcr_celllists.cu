// pcr_celllists.cu — EXPERIMENTAL, off unless PCR_VAL_LISTS=1 (pcr_ransac.cu reads the switch).
//
// Per-fine-cell candidate lists over a target cloud, built from its uniform grid.  Motivation (profiles/r1_summary.md):
// k_ransac_validate is instruction-bound at ~300 instructions per radius-limited nearest-neighbour query (27 coarse
// cells, ~20 candidates) and runs tens of millions of queries per RANSAC.  With lists a query computes its fine cell,
// reads one header word and tests the 3-6 points that can be nearest to ANY position inside that cell.
//
// Exactness.  For the cube C of a fine cell (side c, enlarged by 2^-20 for the rounding of the cell mapping) let
//   dmin(t, C) / dmax(t, C) = the smallest / largest distance from target point t to C,
//   bound = min( min_t' dmax(t', C), r ) * (1 + 1e-5).
// The list is K(C) = { t : dmin(t, C) <= bound }.  For a query q in C whose nearest point (by the fp32 rule D1, ties by
// index, D2) is t*, and the minimiser t' of dmax:  dmin(t*, C) <= |q - t*| <= |q - t'| (1 + 5e-7) <= dmax(t', C) (1 + 5e-7),
// and only points closer than r matter; 5e-7 bounds the relative rounding of two fp32 squared distances.  So t* is on
// the list, every point that ties it is too, and the final (d2, index) key over the list returns exactly what
// grid_nn1 returns.  Cells whose list would exceed 14 entries, or that do not fit the item pool, are flagged and fall
// back to the full search.  The NumPy prototype tools/proto/cell_candidate_lists.py checks the claim against the CPU
// oracle (18,000 queries under good and bad hypotheses: identical indices and fp32 distances; 4.7 candidates per query
// at c = v/2, 2.8 at c = v/3 on the bench pair).
//
// Build cost: one thread per fine cell, two sweeps over the target points of the coarse cells within r of the cube
// (fp64 cube distances); list space is claimed with one atomicAdd per non-empty cell — list ORDER is therefore not
// deterministic, the query RESULT is (the key is order-free).
//
// Status: compiled, NOT yet run on a GPU (the round's GPU budget was spent when it was written).  The default path
// does not touch it.
#include "pcr_common.cuh"
#include "pcr_celllists.cuh"

namespace {

__global__ void __launch_bounds__(256) k_celllists_build(Grid g, double fox, double foy, double foz, double c, int fnx, int fny,
                                                         int fnz, double r, uint32_t *__restrict__ head,
                                                         float4 *__restrict__ items, unsigned int *__restrict__ total,
                                                         unsigned int cap) {
    const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
    if (id >= (long long)fnx * fny * fnz) return;
    head[id] = celllists_build_cell(g, fox, foy, foz, c, fnx, fny, r, id, items, total, cap);
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
    KScope ks(ctx, KC_GRID_BUILD, 4.0 * (double)ncell + 16.0 * (double)g.n);
    PCR_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned int), ctx->stream));
    k_celllists_build<<<div_up((long long)ncell, 256), 256, 0, ctx->stream>>>(g, g.ox - pad, g.oy - pad, g.oz - pad, c, fnx, fny, fnz, r,
                                                                            head, items, total, cap);
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
