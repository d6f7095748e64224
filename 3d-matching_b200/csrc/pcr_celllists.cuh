// pcr_celllists.cuh — the per-cell logic of the EXPERIMENTAL candidate lists (see pcr_celllists.cu), written
// __host__ __device__ so that tests/c/celllists_host_check.cu can run the very same code on the CPU against brute force.
#pragma once
#include "pcr_common.cuh"

constexpr int PCR_LIST_MAX = 14;

// Candidate list of fine cell `id`: returns the header word ((offset << 4) | count; 0 = empty; 15 = use the full search)
// and appends the points to `items` (space claimed with one atomic add on *total).
__host__ __device__ inline uint32_t celllists_build_cell(const Grid &g, double fox, double foy, double foz, double c, int fnx,
                                                         int fny, double r, long long id, float4 *items,
                                                         unsigned int *total, unsigned int cap) {
    const int ix = (int)(id % fnx), iy = (int)((id / fnx) % fny), iz = (int)(id / ((long long)fnx * fny));
    const double half = 0.5 * c * (1.0 + 9.5367431640625e-07);
    const double cx = fox + ((double)ix + 0.5) * c, cy = foy + ((double)iy + 0.5) * c, cz = foz + ((double)iz + 0.5) * c;
    const double R = r * (1.0 + 1e-5);
    // coarse cells that can hold a point within R of the cube
    const int x0 = max((int)floor((cx - half - R - g.ox) * g.inv_h), 0), x1 = min((int)floor((cx + half + R - g.ox) * g.inv_h), g.nx - 1);
    const int y0 = max((int)floor((cy - half - R - g.oy) * g.inv_h), 0), y1 = min((int)floor((cy + half + R - g.oy) * g.inv_h), g.ny - 1);
    const int z0 = max((int)floor((cz - half - R - g.oz) * g.inv_h), 0), z1 = min((int)floor((cz + half + R - g.oz) * g.inv_h), g.nz - 1);
    if (x0 > x1 || y0 > y1 || z0 > z1) return 0u;
    // sweep 1: the smallest "farthest corner" distance
    double best = 1.0e300;
    for (int z = z0; z <= z1; z++)
        for (int y = y0; y <= y1; y++) {
            const long long row = ((long long)z * g.ny + y) * g.nx;
            const uint32_t b = g.start[row + x0], e = g.start[row + x1 + 1];
            for (uint32_t k = b; k < e; k++) {
                const float4 p = g.sorted[k];
                const double ax = fabs((double)p.x - cx) + half, ay = fabs((double)p.y - cy) + half, az = fabs((double)p.z - cz) + half;
                best = fmin(best, (ax * ax + ay * ay) + az * az);
            }
        }
    if (best >= 1.0e300) return 0u;
    const double lim = fmin(best, r * r) * ((1.0 + 1e-5) * (1.0 + 1e-5));
    // sweep 2: every point whose nearest cube point is within the bound
    uint32_t loc[PCR_LIST_MAX];
    int cnt = 0;
    bool over = false;
    for (int z = z0; z <= z1; z++)
        for (int y = y0; y <= y1; y++) {
            const long long row = ((long long)z * g.ny + y) * g.nx;
            const uint32_t b = g.start[row + x0], e = g.start[row + x1 + 1];
            for (uint32_t k = b; k < e; k++) {
                const float4 p = g.sorted[k];
                const double ax = fmax(fabs((double)p.x - cx) - half, 0.0), ay = fmax(fabs((double)p.y - cy) - half, 0.0),
                             az = fmax(fabs((double)p.z - cz) - half, 0.0);
                if ((ax * ax + ay * ay) + az * az <= lim) {
                    if (cnt < PCR_LIST_MAX) loc[cnt++] = k;
                    else over = true;
                }
            }
        }
    if (cnt == 0) return 0u;
    if (over) return 15u;
#ifdef __CUDA_ARCH__
    const unsigned int off = atomicAdd(total, (unsigned int)cnt);
#else
    const unsigned int off = *total;
    *total += (unsigned int)cnt;
#endif
    if (off + (unsigned int)cnt > cap) return 15u;
    for (int i = 0; i < cnt; i++) items[off + i] = g.sorted[loc[i]];
    return (off << 4) | (unsigned int)cnt;
}

// geometry of the fine lattice over a grid: origin = grid origin - pad, cell c = r / (1.5 div)
struct CellListsDims {
    double c, pad;
    double fn[3];
};
static inline CellListsDims celllists_dims(const Grid &g, double r, int div) {
    CellListsDims d;
    d.c = r / (1.5 * div);  // r = 1.5 v for RANSAC validation: c = v / div
    d.pad = r * (1.0 + 1e-3) + d.c;
    d.fn[0] = ceil((g.nx * g.h + 2.0 * d.pad) / d.c);
    d.fn[1] = ceil((g.ny * g.h + 2.0 * d.pad) / d.c);
    d.fn[2] = ceil((g.nz * g.h + 2.0 * d.pad) / d.c);
    return d;
}
