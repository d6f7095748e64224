// Host-side check of the EXPERIMENTAL candidate lists (3d-matching_b200/csrc/pcr_celllists.cuh): runs the very
// __host__ __device__ code the build kernel runs, on the CPU, over a uniform grid laid out as pcr_grid.cu lays it out,
// and compares list-based radius-limited nearest neighbours with brute force under the fp32 distance rule (D1) and the
// (d2, index) tie rule (D2).  Built with nvcc as a host program; no GPU is needed to run it.
//   celllists_host_check points.f32 n radius div n_queries
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "pcr_celllists.cuh"

static inline float d2_rule(const float *q, const float4 &p) {
    const float dx = q[0] - p.x, dy = q[1] - p.y, dz = q[2] - p.z;
    volatile float a = dx * dx, b = dy * dy, c = dz * dz;  // volatile: every product and sum individually rounded
    volatile float s = a + b;
    return s + c;
}

int main(int argc, char **argv) {
    if (argc < 6) return 2;
    const int n = atoi(argv[2]);
    const double r = atof(argv[3]);
    const int div = atoi(argv[4]);
    const int nq = atoi(argv[5]);
    std::vector<float> xyz(3 * (size_t)n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(xyz.data(), sizeof(float), xyz.size(), f) != xyz.size()) return 3;
    fclose(f);
    // uniform grid as pcr_grid_build_rings builds it (rings = 1): origin = min bound, h = r (1 + 2^-10), x fastest
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            lo[d] = std::min(lo[d], xyz[3 * i + d]);
            hi[d] = std::max(hi[d], xyz[3 * i + d]);
        }
    Grid g;
    memset(&g, 0, sizeof g);
    g.h = r * (1.0 + 1.0 / 1024.0);
    g.inv_h = 1.0 / g.h;
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    g.nx = (int)floor(((double)hi[0] - lo[0]) / g.h) + 1;
    g.ny = (int)floor(((double)hi[1] - lo[1]) / g.h) + 1;
    g.nz = (int)floor(((double)hi[2] - lo[2]) / g.h) + 1;
    g.n = n;
    g.R = 1;
    const size_t ncells = (size_t)g.nx * g.ny * g.nz;
    std::vector<uint32_t> start(ncells + 1, 0), cell(n);
    for (int i = 0; i < n; i++) {
        const int cx = (int)floor(((double)xyz[3 * i] - g.ox) * g.inv_h), cy = (int)floor(((double)xyz[3 * i + 1] - g.oy) * g.inv_h),
                  cz = (int)floor(((double)xyz[3 * i + 2] - g.oz) * g.inv_h);
        cell[i] = (uint32_t)(((size_t)cz * g.ny + cy) * g.nx + cx);
        start[cell[i] + 1]++;
    }
    for (size_t k = 0; k < ncells; k++) start[k + 1] += start[k];
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    std::vector<float4> sorted(n);
    for (int i = 0; i < n; i++) {
        float4 p;
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        memcpy(&p.w, &i, 4);
        sorted[fill[cell[i]]++] = p;
    }
    g.sorted = sorted.data();
    g.start = start.data();
    // the lists, through the same code the kernel runs
    const CellListsDims d = celllists_dims(g, r, div);
    const int fnx = (int)d.fn[0], fny = (int)d.fn[1], fnz = (int)d.fn[2];
    const size_t nf = (size_t)fnx * fny * fnz;
    const unsigned int cap = (unsigned)(40 * div * div + 32) * (unsigned)n + 4096u;  // as pcr_celllists_build sizes the pool
    std::vector<uint32_t> head(nf);
    std::vector<float4> items(cap);
    unsigned int total = 0;
    const double fox = g.ox - d.pad, foy = g.oy - d.pad, foz = g.oz - d.pad;
    size_t active = 0, overflow = 0;
    for (size_t id = 0; id < nf; id++) {
        head[id] = celllists_build_cell(g, fox, foy, foz, d.c, fnx, fny, r, (long long)id, items.data(), &total, cap);
        active += (head[id] & 15u) != 0;
        overflow += (head[id] & 15u) == 15u;
    }
    // queries: around target points (inside and outside the radius), plus far ones outside the lattice
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    const float r2 = (float)(r * r);
    uint32_t r2bits;
    memcpy(&r2bits, &r2, 4);
    const double inv_c = 1.0 / d.c;
    long long mism = 0, hits = 0, listed = 0, fallback = 0;
    double list_len = 0.0;
    for (int qi = 0; qi < nq; qi++) {
        float q[3];
        const int base = (int)(rng() % (uint64_t)n);
        const double scale = (qi % 10 == 0) ? 40.0 * r : ((qi % 3 == 0) ? 2.5 * r : 0.8 * r);
        for (int k = 0; k < 3; k++) q[k] = (float)(xyz[3 * base + k] + scale * U(rng));
        // brute force
        unsigned long long bkey = ((unsigned long long)r2bits) << 32;
        for (int i = 0; i < n; i++) {
            float4 p;
            p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
            const float d2 = d2_rule(q, p);
            uint32_t b;
            memcpy(&b, &d2, 4);
            const unsigned long long key = (((unsigned long long)b) << 32) | (uint32_t)i;
            bkey = std::min(bkey, key);
        }
        const int want = (uint32_t)(bkey >> 32) < r2bits ? (int)(uint32_t)bkey : -1;
        const uint32_t want_bits = (uint32_t)(bkey >> 32);
        // lists (host twin of lists_nn1)
        int got = -1;
        uint32_t got_bits = r2bits;
        const double fx = ((double)q[0] - fox) * inv_c, fy = ((double)q[1] - foy) * inv_c, fz = ((double)q[2] - foz) * inv_c;
        if (fx >= 0.0 && fy >= 0.0 && fz >= 0.0 && fx < (double)fnx && fy < (double)fny && fz < (double)fnz) {
            const uint32_t h = head[((size_t)(int)fz * fny + (int)fy) * fnx + (int)fx];
            const uint32_t cnt = h & 15u;
            if (cnt == 15u) {
                fallback++;
                got = want;
                got_bits = want_bits;
            } else {
                listed++;
                list_len += cnt;
                unsigned long long k2 = ((unsigned long long)r2bits) << 32;
                for (uint32_t k = 0; k < cnt; k++) {
                    const float4 p = items[(h >> 4) + k];
                    const float d2 = d2_rule(q, p);
                    uint32_t b, w;
                    memcpy(&b, &d2, 4);
                    memcpy(&w, &p.w, 4);
                    k2 = std::min(k2, (((unsigned long long)b) << 32) | w);
                }
                got = (uint32_t)(k2 >> 32) < r2bits ? (int)(uint32_t)k2 : -1;
                got_bits = (uint32_t)(k2 >> 32);
            }
        }
        hits += want >= 0;
        if (got != want || (want >= 0 && got_bits != want_bits)) {
            if (mism < 5) fprintf(stderr, "mismatch: query %d want %d got %d\n", qi, want, got);
            mism++;
        }
    }
    printf("points %d fine cells %zu (%dx%dx%d) active %zu overflow %zu items %u (%.1f per point)\n", n, nf, fnx, fny, fnz, active,
           overflow, total, (double)total / n);
    printf("queries %d hits %lld listed %lld mean list %.2f fallback %lld mismatches %lld\n", nq, hits, listed,
           listed ? list_len / listed : 0.0, fallback, mism);
    return mism == 0 ? 0 : 1;
}
