/*
 * pcr_oracle_literal.c — APPENDIX-A-LITERAL fp64 restatement of the Open3D 0.19.0 routines the reference calls.
 * TEST INFRASTRUCTURE ONLY (same rule as pcr_oracle.c: only tests/ and bench.py's CPU legs may load it).
 *
 * Purpose ("next" #2): pcr_oracle.c and the CUDA kernels share an arithmetic specification (DESIGN.md §3,
 * rules D1-D9) chosen so that both produce the same bits: fp32 distances, fixed-point sums, a cumulative ICP transform,
 * a per-block FPFH normaliser, Umeyama through Sigma^T Sigma, LDL^T without pivoting, polynomial sin/cos/atan2/acos.
 * THIS file follows SURVEY.md Appendix A to the letter instead, in double precision throughout and with libm:
 *   A.1  voxel means accumulated in double, in input order                       (src/ply/ply.py:106)
 *   A.2  covariance from 9 running cumulants, Geometric-Tools 3x3 solver, libm   (src/ply/ply.py:110-112, 133-135)
 *   A.3  hybrid search on fp64 squared distances
 *   A.4  FPFH with the RUNNING normaliser over (neighbour, bin), libm atan2/acos (src/ply/ply.py:117-120)
 *   A.5  33-D distances accumulated in groups of four (nanoflann L2 adaptor)     (src/matcher/ransac.py:85)
 *   A.6  Eigen::umeyama through a Jacobi SVD of Sigma; fp64 validation distances,
 *        fitness / RMSE compared as doubles                                      (src/matcher/ransac.py:42-59)
 *   A.7  ICP with the INCREMENTAL in-place pcd.Transform(update), double J^T J
 *        sums, LDL^T with Eigen's diagonal pivoting, libm sin/cos                (src/matcher/icp.py:42-48)
 * It shares NO code with pcr_oracle.c (own grid search instead of the KD-tree, own SVD, own solver).  The only rules it
 * keeps are the ones Open3D leaves undefined: output order of the voxel map (ascending voxel id), tie order (lowest
 * index), the RANSAC sample stream (Philox, as D6 — Open3D's is an unseeded mt19937) evaluated as the sequential loop.
 * tests/test_oracle_literal.py reports, on the cfg1 and cfg2 clouds, how far the D-rule results are from these:
 * neighbour-index / correspondence-set agreement and max |dT| against the north-star tolerance (1e-5 rotation,
 * 1e-5 x extent translation).  PARITY REMAINS UNPINNED against real Open3D (absent); this bounds what the D rules changed.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LIT_API __attribute__((visibility("default")))
#define LIT_PI 3.14159265358979323846

/* ------------------------------------------------------------------------------------------------------------ */
/* uniform grid over (n,3) double points: exact radius / k-nearest search by scanning the 27 neighbouring cells   */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const double *p;
    int n, nx, ny, nz;
    double ox, oy, oz, h;
    int *start, *idx;
} lgrid;

static int lg_cell(double v, double o, double h, int n) {
    int c = (int)floor((v - o) / h);
    if (c < 0) c = 0;
    if (c >= n) c = n - 1;
    return c;
}

static lgrid *lg_build(const double *p, int n, double radius) {
    lgrid *g = (lgrid *)calloc(1, sizeof(lgrid));
    g->p = p;
    g->n = n;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            const double v = p[3 * i + d];
            if (i == 0 || v < lo[d]) lo[d] = v;
            if (i == 0 || v > hi[d]) hi[d] = v;
        }
    double h = radius * (1.0 + 1e-9);
    for (;;) { /* at most 2^24 cells: larger cells still cover the radius with a 27-cell probe */
        const double c = (floor((hi[0] - lo[0]) / h) + 1) * (floor((hi[1] - lo[1]) / h) + 1) * (floor((hi[2] - lo[2]) / h) + 1);
        if (c <= 16777216.0) break;
        h *= 1.26;
    }
    g->h = h;
    g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
    g->nx = (int)floor((hi[0] - lo[0]) / h) + 1;
    g->ny = (int)floor((hi[1] - lo[1]) / h) + 1;
    g->nz = (int)floor((hi[2] - lo[2]) / h) + 1;
    const size_t nc = (size_t)g->nx * g->ny * g->nz;
    g->start = (int *)calloc(nc + 1, sizeof(int));
    g->idx = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *cell = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        const int cx = lg_cell(p[3 * i], g->ox, h, g->nx), cy = lg_cell(p[3 * i + 1], g->oy, h, g->ny), cz = lg_cell(p[3 * i + 2], g->oz, h, g->nz);
        cell[i] = (cz * g->ny + cy) * g->nx + cx;
        g->start[cell[i] + 1]++;
    }
    for (size_t c = 0; c < nc; c++) g->start[c + 1] += g->start[c];
    int *fill = (int *)malloc(sizeof(int) * nc);
    memcpy(fill, g->start, sizeof(int) * nc);
    for (int i = 0; i < n; i++) g->idx[fill[cell[i]]++] = i; /* ascending original index inside a cell */
    free(fill);
    free(cell);
    return g;
}

static void lg_free(lgrid *g) {
    if (!g) return;
    free(g->start);
    free(g->idx);
    free(g);
}

typedef struct { double d2; int i; } lcand;
static int lc_less(lcand a, lcand b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.i < b.i); }

/* A.3 SearchHybrid: the k nearest points with d2 < r2 (strict), ascending (d2, index).  buf holds k entries. */
static int lg_hybrid(const lgrid *g, const double *q, double r2, int k, lcand *buf) {
    int cnt = 0;
    if (g->n == 0 || k <= 0) return 0;
    /* the query may lie outside the box: cells are clamped, and a query farther than h from the box finds nothing */
    const int cx = (int)floor((q[0] - g->ox) / g->h), cy = (int)floor((q[1] - g->oy) / g->h), cz = (int)floor((q[2] - g->oz) / g->h);
    for (int z = cz - 1; z <= cz + 1; z++) {
        if (z < 0 || z >= g->nz) continue;
        for (int y = cy - 1; y <= cy + 1; y++) {
            if (y < 0 || y >= g->ny) continue;
            for (int x = cx - 1; x <= cx + 1; x++) {
                if (x < 0 || x >= g->nx) continue;
                const size_t c = ((size_t)z * g->ny + y) * g->nx + x;
                for (int s = g->start[c]; s < g->start[c + 1]; s++) {
                    const int j = g->idx[s];
                    const double dx = q[0] - g->p[3 * j], dy = q[1] - g->p[3 * j + 1], dz = q[2] - g->p[3 * j + 2];
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    if (!(d2 < r2)) continue;
                    lcand c1 = {d2, j};
                    if (cnt == k && !lc_less(c1, buf[k - 1])) continue;
                    int pos = cnt < k ? cnt++ : k - 1;
                    while (pos > 0 && lc_less(c1, buf[pos - 1])) { buf[pos] = buf[pos - 1]; pos--; }
                    buf[pos] = c1;
                }
            }
        }
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* A.1 VoxelDownSample                                                                                           */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct { long long key; int idx; } lvk;
static int lvk_cmp(const void *a, const void *b) {
    const lvk *x = (const lvk *)a, *y = (const lvk *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

LIT_API int lit_voxel_downsample(const double *p, int n, double voxel, double *out, int *m_out) {
    *m_out = 0;
    if (!(voxel > 0.0)) return -1;
    if (n == 0) return 0;
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) lo[d] = hi[d] = p[d];
    for (int i = 1; i < n; i++)
        for (int d = 0; d < 3; d++) {
            if (p[3 * i + d] < lo[d]) lo[d] = p[3 * i + d];
            if (p[3 * i + d] > hi[d]) hi[d] = p[3 * i + d];
        }
    double org[3];
    long long dim[3];
    for (int d = 0; d < 3; d++) {
        org[d] = lo[d] - voxel * 0.5;
        dim[d] = (long long)floor((hi[d] - org[d]) / voxel) + 1;
    }
    lvk *keys = (lvk *)malloc(sizeof(lvk) * (size_t)n);
    for (int i = 0; i < n; i++) {
        long long c[3];
        for (int d = 0; d < 3; d++) c[d] = (long long)floor((p[3 * i + d] - org[d]) / voxel);
        keys[i].key = (c[2] * dim[1] + c[1]) * dim[0] + c[0];
        keys[i].idx = i;
    }
    qsort(keys, (size_t)n, sizeof(lvk), lvk_cmp);
    int m = 0;
    for (int i = 0; i < n;) {
        int j = i;
        double s[3] = {0, 0, 0};
        while (j < n && keys[j].key == keys[i].key) { /* AccumulatedPoint: double sums in input order */
            for (int d = 0; d < 3; d++) s[d] += p[3 * keys[j].idx + d];
            j++;
        }
        for (int d = 0; d < 3; d++) out[3 * m + d] = s[d] / (double)(j - i);
        m++;
        i = j;
    }
    free(keys);
    *m_out = m;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* A.2 EstimateNormals (fast_normal_computation = true): Geometric-Tools robust symmetric 3x3 eigen-solver         */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct { double v[3]; } l3;
static l3 l_cross(l3 a, l3 b) { l3 r = {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}}; return r; }
static double l_dot(l3 a, l3 b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }

static l3 gt_eigvec0(const double A[3][3], double ev) {
    l3 r0 = {{A[0][0] - ev, A[0][1], A[0][2]}}, r1 = {{A[0][1], A[1][1] - ev, A[1][2]}}, r2 = {{A[0][2], A[1][2], A[2][2] - ev}};
    l3 c[3] = {l_cross(r0, r1), l_cross(r0, r2), l_cross(r1, r2)};
    double d[3] = {l_dot(c[0], c[0]), l_dot(c[1], c[1]), l_dot(c[2], c[2])};
    int im = 0;
    double dm = d[0];
    if (d[1] > dm) { dm = d[1]; im = 1; }
    if (d[2] > dm) { im = 2; }
    const double s = 1.0 / sqrt(d[im]);
    l3 r = {{c[im].v[0] * s, c[im].v[1] * s, c[im].v[2] * s}};
    return r;
}

static l3 gt_eigvec1(const double A[3][3], l3 e0, double ev1) {
    l3 U, V;
    if (fabs(e0.v[0]) > fabs(e0.v[1])) {
        const double inv = 1.0 / sqrt(e0.v[0] * e0.v[0] + e0.v[2] * e0.v[2]);
        U.v[0] = -e0.v[2] * inv; U.v[1] = 0; U.v[2] = e0.v[0] * inv;
    } else {
        const double inv = 1.0 / sqrt(e0.v[1] * e0.v[1] + e0.v[2] * e0.v[2]);
        U.v[0] = 0; U.v[1] = e0.v[2] * inv; U.v[2] = -e0.v[1] * inv;
    }
    V = l_cross(e0, U);
    l3 AU, AV;
    for (int i = 0; i < 3; i++) {
        AU.v[i] = A[i][0] * U.v[0] + A[i][1] * U.v[1] + A[i][2] * U.v[2];
        AV.v[i] = A[i][0] * V.v[0] + A[i][1] * V.v[1] + A[i][2] * V.v[2];
    }
    double m00 = l_dot(U, AU) - ev1, m01 = l_dot(U, AV), m11 = l_dot(V, AV) - ev1;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    l3 r;
    if (a00 >= a11) {
        if (fmax(a00, a01) > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            for (int i = 0; i < 3; i++) r.v[i] = m01 * U.v[i] - m00 * V.v[i];
            return r;
        }
        return U;
    }
    if (fmax(a11, a01) > 0) {
        if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
        else { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
        for (int i = 0; i < 3; i++) r.v[i] = m11 * U.v[i] - m01 * V.v[i];
        return r;
    }
    return U;
}

static l3 gt_smallest_eigvec(const double C[3][3]) {
    double mc = C[0][0];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) if (C[i][j] > mc) mc = C[i][j];
    l3 zero = {{0, 0, 0}};
    if (mc == 0.0) return zero;
    double A[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i][j] = C[i][j] / mc;
    const double norm = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (norm > 0) {
        const double q = (A[0][0] + A[1][1] + A[2][2]) / 3;
        const double b00 = A[0][0] - q, b11 = A[1][1] - q, b22 = A[2][2] - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        const double c00 = b11 * b22 - A[1][2] * A[1][2], c01 = A[0][1] * b22 - A[1][2] * A[0][2], c02 = A[0][1] * A[1][2] - b11 * A[0][2];
        const double det = (b00 * c00 - A[0][1] * c01 + A[0][2] * c02) / (p * p * p);
        double hd = det * 0.5;
        if (hd < -1) hd = -1;
        if (hd > 1) hd = 1;
        const double ang = acos(hd) / 3;
        const double beta2 = cos(ang) * 2, beta0 = cos(ang + 2.09439510239319549) * 2, beta1 = -(beta0 + beta2);
        const double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (hd >= 0) {
            const l3 v2 = gt_eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            const l3 v1 = gt_eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return l_cross(v1, v2);
        }
        const l3 v0 = gt_eigvec0(A, e0);
        if (e0 < e1 && e0 < e2) return v0;
        const l3 v1 = gt_eigvec1(A, v0, e1);
        if (e1 < e0 && e1 < e2) return v1;
        return l_cross(v0, v1);
    }
    l3 r = {{0, 0, 1}};
    if (C[0][0] < C[1][1] && C[0][0] < C[2][2]) { r.v[0] = 1; r.v[2] = 0; }
    else if (C[1][1] < C[0][0] && C[1][1] < C[2][2]) { r.v[1] = 1; r.v[2] = 0; }
    return r;
}

/* nbr_out (optional): n x max_nn int32 neighbour indices (-1 padded), for the index-agreement report */
LIT_API int lit_estimate_normals(const double *p, int n, double radius, int max_nn, double *normals, int *nbr_out) {
    lgrid *g = lg_build(p, n, radius);
    const double r2 = radius * radius;
#pragma omp parallel
    {
        lcand *buf = (lcand *)malloc(sizeof(lcand) * (size_t)(max_nn > 0 ? max_nn : 1));
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            const int c = lg_hybrid(g, p + 3 * i, r2, max_nn, buf);
            if (nbr_out) for (int k = 0; k < max_nn; k++) nbr_out[(size_t)i * max_nn + k] = k < c ? buf[k].i : -1;
            double C[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            if (c >= 3) {
                double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                for (int k = 0; k < c; k++) {
                    const double *q = p + 3 * buf[k].i;
                    cu[0] += q[0]; cu[1] += q[1]; cu[2] += q[2];
                    cu[3] += q[0] * q[0]; cu[4] += q[0] * q[1]; cu[5] += q[0] * q[2];
                    cu[6] += q[1] * q[1]; cu[7] += q[1] * q[2]; cu[8] += q[2] * q[2];
                }
                for (int k = 0; k < 9; k++) cu[k] /= (double)c;
                C[0][0] = cu[3] - cu[0] * cu[0]; C[1][1] = cu[6] - cu[1] * cu[1]; C[2][2] = cu[8] - cu[2] * cu[2];
                C[0][1] = C[1][0] = cu[4] - cu[0] * cu[1];
                C[0][2] = C[2][0] = cu[5] - cu[0] * cu[2];
                C[1][2] = C[2][1] = cu[7] - cu[1] * cu[2];
            }
            l3 nr = gt_smallest_eigvec(C);
            const double len = sqrt(l_dot(nr, nr));
            if (len == 0.0 || len != len) { nr.v[0] = 0; nr.v[1] = 0; nr.v[2] = 1; }
            for (int d = 0; d < 3; d++) normals[3 * i + d] = nr.v[d];
        }
        free(buf);
    }
    lg_free(g);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* A.4 FPFH                                                                                                      */
/* ------------------------------------------------------------------------------------------------------------ */
static void lit_pair_features(const double *p1, const double *n1in, const double *p2, const double *n2in, double *f) {
    l3 n1 = {{n1in[0], n1in[1], n1in[2]}}, n2 = {{n2in[0], n2in[1], n2in[2]}};
    l3 d = {{p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}};
    f[0] = f[1] = f[2] = f[3] = 0;
    const double len = sqrt(l_dot(d, d));
    if (len == 0.0) return;
    f[3] = len;
    const double a1 = l_dot(n1, d) / len, a2 = l_dot(n2, d) / len;
    if (acos(fabs(a1)) > acos(fabs(a2))) {
        const l3 t = n1; n1 = n2; n2 = t;
        for (int k = 0; k < 3; k++) d.v[k] = -d.v[k];
        f[2] = -a2;
    } else {
        f[2] = a1;
    }
    l3 v = l_cross(d, n1);
    const double vn = sqrt(l_dot(v, v));
    if (vn == 0.0) { f[0] = f[1] = f[2] = f[3] = 0; return; }
    for (int k = 0; k < 3; k++) v.v[k] /= vn;
    const l3 w = l_cross(n1, v);
    f[1] = l_dot(v, n2);
    f[0] = atan2(l_dot(w, n2), l_dot(n1, n2));
}

static int lit_bin(double x) {
    int h = (int)floor(x);
    return h < 0 ? 0 : (h >= 11 ? 10 : h);
}

/* fpfh: (n,33) double, row = point */
LIT_API int lit_fpfh(const double *p, const double *nrm, int n, double radius, int max_nn, double *fpfh) {
    lgrid *g = lg_build(p, n, radius);
    const double r2 = radius * radius;
    const int K = max_nn > 0 ? max_nn : 1;
    int *ni = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1) * K), *nc = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    double *nd = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * K);
    double *spfh = (double *)calloc((size_t)(n > 0 ? n : 1) * 33, sizeof(double));
#pragma omp parallel
    {
        lcand *buf = (lcand *)malloc(sizeof(lcand) * (size_t)K);
#pragma omp for schedule(dynamic, 128)
        for (int i = 0; i < n; i++) {
            const int c = lg_hybrid(g, p + 3 * i, r2, max_nn, buf);
            nc[i] = c;
            for (int k = 0; k < c; k++) { ni[(size_t)i * K + k] = buf[k].i; nd[(size_t)i * K + k] = buf[k].d2; }
            if (c > 1) {
                const double inc = 100.0 / (double)(c - 1);
                double *h = spfh + (size_t)i * 33;
                for (int k = 1; k < c; k++) {
                    double f[4];
                    const int j = buf[k].i;
                    lit_pair_features(p + 3 * i, nrm + 3 * i, p + 3 * j, nrm + 3 * j, f);
                    h[lit_bin(11 * (f[0] + LIT_PI) / (2.0 * LIT_PI))] += inc;
                    h[11 + lit_bin(11 * (f[1] + 1.0) * 0.5)] += inc;
                    h[22 + lit_bin(11 * (f[2] + 1.0) * 0.5)] += inc;
                }
            }
        }
        free(buf);
    }
#pragma omp parallel for schedule(dynamic, 128)
    for (int i = 0; i < n; i++) {
        double F[33], sum[3] = {0, 0, 0};
        for (int j = 0; j < 33; j++) F[j] = 0;
        const int c = nc[i];
        if (c > 1) {
            for (int k = 1; k < c; k++) {
                const double dist = nd[(size_t)i * K + k];
                if (dist == 0.0) continue;
                const double *hs = spfh + (size_t)ni[(size_t)i * K + k] * 33;
                for (int j = 0; j < 33; j++) { /* Open3D: val added to the feature AND to the running block sum */
                    const double val = hs[j] / dist;
                    sum[j / 11] += val;
                    F[j] += val;
                }
            }
            for (int b = 0; b < 3; b++) if (sum[b] != 0.0) sum[b] = 100.0 / sum[b];
            const double *hi = spfh + (size_t)i * 33;
            for (int j = 0; j < 33; j++) { F[j] *= sum[j / 11]; F[j] += hi[j]; }
        }
        for (int j = 0; j < 33; j++) fpfh[(size_t)i * 33 + j] = F[j];
    }
    free(ni); free(nc); free(nd); free(spfh);
    lg_free(g);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* A.5 CorrespondencesFromFeatures: exact 1-NN in 33-D; distance as nanoflann's L2 adaptor accumulates it         */
/* ------------------------------------------------------------------------------------------------------------ */
static double lit_dist33(const double *a, const double *b) {
    double r = 0;
    int k = 0;
    for (; k + 4 <= 33; k += 4) {
        const double d0 = a[k] - b[k], d1 = a[k + 1] - b[k + 1], d2 = a[k + 2] - b[k + 2], d3 = a[k + 3] - b[k + 3];
        r += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; k < 33; k++) { const double d = a[k] - b[k]; r += d * d; }
    return r;
}

LIT_API int lit_nn_features(const double *fq, int nq, const double *fb, int nb, int *nn) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nq; i++) {
        double best = INFINITY;
        int bi = -1;
        for (int j = 0; j < nb; j++) {
            const double d = lit_dist33(fq + (size_t)i * 33, fb + (size_t)j * 33);
            if (d < best) { best = d; bi = j; }
        }
        nn[i] = bi;
    }
    return 0;
}

LIT_API int lit_match_features(const double *fs, int ms, const double *ft, int mt, int mutual, double ratio, int *corr, int *c_out) {
    *c_out = 0;
    if (ms == 0 || mt == 0) return 0;
    int *ns = (int *)malloc(sizeof(int) * (size_t)ms);
    lit_nn_features(fs, ms, ft, mt, ns);
    int c = 0;
    if (mutual) {
        int *nt = (int *)malloc(sizeof(int) * (size_t)mt);
        lit_nn_features(ft, mt, fs, ms, nt);
        for (int i = 0; i < ms; i++)
            if (ns[i] >= 0 && nt[ns[i]] == i) { corr[2 * c] = i; corr[2 * c + 1] = ns[i]; c++; }
        free(nt);
        if (c >= (int)(ratio * ms)) { free(ns); *c_out = c; return 0; }
        c = 0;
    }
    for (int i = 0; i < ms; i++) if (ns[i] >= 0) { corr[2 * c] = i; corr[2 * c + 1] = ns[i]; c++; }
    free(ns);
    *c_out = c;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* A.6 RANSAC: Eigen::umeyama (no scaling) through a one-sided Jacobi SVD of Sigma                                */
/* ------------------------------------------------------------------------------------------------------------ */
static double det3(const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

/* A = U diag(s) V^T, one-sided (Hestenes) Jacobi on the columns; null columns of U completed to an orthonormal basis */
static void svd3(const double A[3][3], double U[3][3], double s[3], double V[3][3]) {
    double W[3][3];
    memcpy(W, A, sizeof(W));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = i == j;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double a = 0, b = 0, g = 0;
                for (int k = 0; k < 3; k++) { a += W[k][p] * W[k][p]; b += W[k][q] * W[k][q]; g += W[k][p] * W[k][q]; }
                if (g == 0.0 || fabs(g) <= 1e-300) continue;
                off = fmax(off, fabs(g) / sqrt(a * b + 1e-300));
                const double zeta = (b - a) / (2.0 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int k = 0; k < 3; k++) {
                    const double wp = W[k][p], wq = W[k][q];
                    W[k][p] = c * wp - sn * wq;
                    W[k][q] = sn * wp + c * wq;
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - sn * vq;
                    V[k][q] = sn * vp + c * vq;
                }
            }
        if (off < 1e-16) break;
    }
    int ord[3] = {0, 1, 2};
    double nrm[3];
    for (int j = 0; j < 3; j++) nrm[j] = sqrt(W[0][j] * W[0][j] + W[1][j] * W[1][j] + W[2][j] * W[2][j]);
    for (int a = 0; a < 2; a++) for (int b = a + 1; b < 3; b++) if (nrm[ord[b]] > nrm[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double Vs[3][3];
    int rank = 0;
    for (int j = 0; j < 3; j++) {
        const int o = ord[j];
        s[j] = nrm[o];
        for (int k = 0; k < 3; k++) Vs[k][j] = V[k][o];
        if (nrm[o] > 1e-14 * (nrm[ord[0]] > 0 ? nrm[ord[0]] : 1.0) && nrm[o] > 0) {
            for (int k = 0; k < 3; k++) U[k][j] = W[k][o] / nrm[o];
            rank = j + 1;
        } else {
            for (int k = 0; k < 3; k++) U[k][j] = 0;
        }
    }
    memcpy(V, Vs, sizeof(Vs));
    /* complete U (the completion is what a full SVD returns up to a rotation inside the null space; for rank 2 it is unique
       up to sign, and the sign is absorbed by umeyama's determinant rule) */
    if (rank == 0) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) U[i][j] = i == j; return; }
    if (rank == 1) {
        l3 u0 = {{U[0][0], U[1][0], U[2][0]}}, e = {{0, 0, 0}};
        int m = fabs(u0.v[0]) <= fabs(u0.v[1]) ? (fabs(u0.v[0]) <= fabs(u0.v[2]) ? 0 : 2) : (fabs(u0.v[1]) <= fabs(u0.v[2]) ? 1 : 2);
        e.v[m] = 1;
        l3 u1 = l_cross(u0, e);
        const double n1 = sqrt(l_dot(u1, u1));
        for (int k = 0; k < 3; k++) U[k][1] = u1.v[k] / n1;
        rank = 2;
    }
    if (rank == 2) {
        l3 u0 = {{U[0][0], U[1][0], U[2][0]}}, u1 = {{U[0][1], U[1][1], U[2][1]}};
        l3 u2 = l_cross(u0, u1);
        for (int k = 0; k < 3; k++) U[k][2] = u2.v[k];
    }
}

/* src3 / tgt3: three points each (rows); T row-major 4x4 */
static void lit_umeyama3(const double s[3][3], const double t[3][3], double *T) {
    double ms[3], mt[3];
    for (int d = 0; d < 3; d++) { ms[d] = (s[0][d] + s[1][d] + s[2][d]) / 3.0; mt[d] = (t[0][d] + t[1][d] + t[2][d]) / 3.0; }
    double Sg[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += (t[k][i] - mt[i]) * (s[k][j] - ms[j]);
            Sg[i][j] = a / 3.0;
        }
    double U[3][3], sv[3], V[3][3];
    svd3(Sg, U, sv, V);
    double S[3] = {1, 1, 1};
    if (det3(U) * det3(V) < 0) S[2] = -1;
    double R[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i][j] = U[i][0] * S[0] * V[j][0] + U[i][1] * S[1] * V[j][1] + U[i][2] * S[2] * V[j][2];
    for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T[4 * i + j] = R[i][j];
        T[4 * i + 3] = mt[i] - (R[i][0] * ms[0] + R[i][1] * ms[1] + R[i][2] * ms[2]);
    }
}

LIT_API void lit_umeyama(const double *src3, const double *tgt3, double *T) {
    double s[3][3], t[3][3];
    memcpy(s, src3, sizeof(s));
    memcpy(t, tgt3, sizeof(t));
    lit_umeyama3(s, t, T);
}

static void lit_philox(uint64_t ctr, uint64_t key, uint32_t out[4]) { /* Philox4x32-10, counter = (ctr, 0), as rule D6 */
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0, k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void lit_apply(const double *T, const double *p, double *o) {
    /* Open3D: homogeneous multiply (Eigen 4x4 * 4-vector) and divide by w (w = 1 for a rigid transform) */
    for (int i = 0; i < 3; i++) o[i] = T[4 * i] * p[0] + T[4 * i + 1] * p[1] + T[4 * i + 2] * p[2] + T[4 * i + 3];
}

typedef struct {
    double T[16];
    double fitness, inlier_rmse;
    long long best_hyp, inlier_count, hyp_evaluated, survivors, est_k;
} lit_ransac_result;

typedef struct { int ok; double T[16]; long long cnt; double err2; int cin; } lit_eval;

LIT_API int lit_ransac(const double *src, int ms, const double *tgt, int mt, const int *corr, int c, double max_dist,
                       double edge_sim, long long max_iter, double confidence, uint64_t seed, lit_ransac_result *res) {
    memset(res, 0, sizeof(*res));
    for (int i = 0; i < 16; i++) res->T[i] = (i % 5 == 0);
    res->best_hyp = -1;
    res->est_k = max_iter;
    if (c < 3 || !(max_dist > 0.0) || ms == 0 || mt == 0) return 0;
    lgrid *g = lg_build(tgt, mt, max_dist);
    const double r2 = max_dist * max_dist;
    long long est_k = max_iter;
    double best_fit = 0.0, best_rmse = 0.0;
    const int CH = 512;
    lit_eval *E = (lit_eval *)malloc(sizeof(lit_eval) * CH);
    long long itr = 0;
    int stop = 0;
    while (itr < est_k && itr < max_iter && !stop) {
        const int nch = (int)((itr + CH < max_iter ? itr + CH : max_iter) - itr);
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < nch; k++) {
            lit_eval *e = &E[k];
            e->ok = 0;
            uint32_t r[4];
            lit_philox((uint64_t)(itr + k), seed, r);
            double s[3][3], t[3][3];
            for (int a = 0; a < 3; a++) {
                const int id = (int)(((uint64_t)r[a] * (uint64_t)c) >> 32);
                for (int d = 0; d < 3; d++) { s[a][d] = src[3 * corr[2 * id] + d]; t[a][d] = tgt[3 * corr[2 * id + 1] + d]; }
            }
            int pass = 1;
            for (int a = 0; a < 3 && pass; a++) /* CorrespondenceCheckerBasedOnEdgeLength */
                for (int b = a + 1; b < 3; b++) {
                    double ds = 0, dt = 0;
                    for (int d = 0; d < 3; d++) { ds += (s[a][d] - s[b][d]) * (s[a][d] - s[b][d]); dt += (t[a][d] - t[b][d]) * (t[a][d] - t[b][d]); }
                    ds = sqrt(ds); dt = sqrt(dt);
                    if (ds < dt * edge_sim || dt < ds * edge_sim) { pass = 0; break; }
                }
            if (!pass) continue;
            lit_umeyama3(s, t, e->T);
            for (int a = 0; a < 3 && pass; a++) { /* CorrespondenceCheckerBasedOnDistance */
                double q[3], d2 = 0;
                lit_apply(e->T, s[a], q);
                for (int d = 0; d < 3; d++) d2 += (q[d] - t[a][d]) * (q[d] - t[a][d]);
                if (sqrt(d2) > max_dist) pass = 0;
            }
            if (!pass) continue;
            long long cnt = 0;
            double err2 = 0;
            for (int i = 0; i < ms; i++) { /* GetRegistrationResultAndCorrespondences */
                double q[3];
                lcand b;
                lit_apply(e->T, src + 3 * i, q);
                if (lg_hybrid(g, q, r2, 1, &b)) { cnt++; err2 += b.d2; }
            }
            int cin = 0;
            for (int i = 0; i < c; i++) {
                double q[3], d2 = 0;
                lit_apply(e->T, src + 3 * corr[2 * i], q);
                for (int d = 0; d < 3; d++) d2 += (q[d] - tgt[3 * corr[2 * i + 1] + d]) * (q[d] - tgt[3 * corr[2 * i + 1] + d]);
                if (sqrt(d2) < max_dist) cin++;
            }
            e->cnt = cnt; e->err2 = err2; e->cin = cin; e->ok = 1;
        }
        for (int k = 0; k < nch; k++) {
            const long long h = itr + k;
            if (h >= est_k) { stop = 1; break; }
            res->hyp_evaluated++;
            if (!E[k].ok) continue;
            res->survivors++;
            const double fit = (double)E[k].cnt / (double)ms;
            const double rmse = E[k].cnt > 0 ? sqrt(E[k].err2 / (double)E[k].cnt) : 0.0;
            if (!(fit > best_fit || (fit == best_fit && rmse < best_rmse))) continue; /* IsBetterRANSACThan */
            best_fit = fit; best_rmse = rmse;
            res->best_hyp = h;
            res->inlier_count = E[k].cnt;
            memcpy(res->T, E[k].T, sizeof(res->T));
            const double ratio = (double)E[k].cin / (double)c;
            const double est = log(1.0 - confidence) / log(1.0 - pow(ratio, 3.0));
            if (est >= 0.0 && est < (double)est_k) est_k = (long long)ceil(est);
        }
        itr += nch;
    }
    res->fitness = best_fit;
    res->inlier_rmse = best_rmse;
    res->est_k = est_k;
    free(E);
    lg_free(g);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* A.7 RegistrationICP, TransformationEstimationPointToPlane                                                     */
/* ------------------------------------------------------------------------------------------------------------ */
/* Eigen::LDLT: symmetric diagonal pivoting (largest remaining |diagonal|), then the triangular solves */
static int lit_ldlt6(double A[6][6], const double *b, double *x) {
    int perm[6];
    double M[6][6];
    memcpy(M, A, sizeof(M));
    for (int i = 0; i < 6; i++) perm[i] = i;
    for (int k = 0; k < 6; k++) {
        int piv = k;
        for (int i = k + 1; i < 6; i++) if (fabs(M[i][i]) > fabs(M[piv][piv])) piv = i;
        if (piv != k) {
            for (int j = 0; j < 6; j++) { const double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
            for (int j = 0; j < 6; j++) { const double t = M[j][k]; M[j][k] = M[j][piv]; M[j][piv] = t; }
            const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        if (M[k][k] == 0.0) return -1;
        for (int i = k + 1; i < 6; i++) M[i][k] /= M[k][k];
        for (int i = k + 1; i < 6; i++)
            for (int j = k + 1; j <= i; j++) { M[i][j] -= M[i][k] * M[k][k] * M[j][k]; M[j][i] = M[i][j]; }
    }
    double y[6];
    for (int i = 0; i < 6; i++) { y[i] = b[perm[i]]; for (int k = 0; k < i; k++) y[i] -= M[i][k] * y[k]; }
    for (int i = 0; i < 6; i++) y[i] /= M[i][i];
    for (int i = 5; i >= 0; i--) for (int k = i + 1; k < 6; k++) y[i] -= M[k][i] * y[k];
    for (int i = 0; i < 6; i++) x[perm[i]] = y[i];
    for (int i = 0; i < 6; i++) if (x[i] != x[i] || isinf(x[i])) return -1;
    return 0;
}

typedef struct {
    double T[16];
    double fitness, inlier_rmse;
    long long inlier_count;
    int iterations, converged;
} lit_icp_result;

static void lit_nn_pass(const lgrid *g, const double *pts, int ns, double r2, int *corr, long long *cnt, double *err2) {
    long long c = 0;
    double e = 0;
#pragma omp parallel for schedule(static) reduction(+ : c, e)
    for (int i = 0; i < ns; i++) {
        lcand b;
        if (lg_hybrid(g, pts + 3 * i, r2, 1, &b)) { corr[i] = b.i; c++; e += b.d2; }
        else corr[i] = -1;
    }
    *cnt = c;
    *err2 = e;
}

LIT_API int lit_icp_point_to_plane(const double *src, int ns, const double *tgt, const double *tn, int nt, double max_dist,
                                   const double *T_init, int max_iter, double rel_fit, double rel_rmse, lit_icp_result *res,
                                   int *corr_out) {
    memset(res, 0, sizeof(*res));
    memcpy(res->T, T_init, 16 * sizeof(double));
    if (!(max_dist > 0.0)) return -1;
    if (ns == 0 || nt == 0) { if (corr_out) for (int i = 0; i < ns; i++) corr_out[i] = -1; return 0; }
    lgrid *g = lg_build(tgt, nt, max_dist);
    const double r2 = max_dist * max_dist;
    double *pcd = (double *)malloc(sizeof(double) * 3 * (size_t)ns);
    for (int i = 0; i < ns; i++) lit_apply(T_init, src + 3 * i, pcd + 3 * i); /* pcd.Transform(init) */
    int *corr = (int *)malloc(sizeof(int) * (size_t)ns);
    double T[16];
    memcpy(T, T_init, sizeof(T));
    long long cnt;
    double err2;
    lit_nn_pass(g, pcd, ns, r2, corr, &cnt, &err2);
    double fit = (double)cnt / ns, rmse = cnt > 0 ? sqrt(err2 / (double)cnt) : 0.0;
    for (int it = 0; it < max_iter; it++) {
        double U[16];
        for (int i = 0; i < 16; i++) U[i] = (i % 5 == 0);
        if (cnt > 0) {
            double A[6][6], b[6], x[6];
            memset(A, 0, sizeof(A));
            memset(b, 0, sizeof(b));
            for (int i = 0; i < ns; i++) { /* sequential double sums (Open3D: OpenMP-reduced; order is not part of A.7) */
                const int j = corr[i];
                if (j < 0) continue;
                const double *s = pcd + 3 * i, *t = tgt + 3 * j, *n = tn + 3 * j;
                const double r = (s[0] - t[0]) * n[0] + (s[1] - t[1]) * n[1] + (s[2] - t[2]) * n[2];
                const double J[6] = {s[1] * n[2] - s[2] * n[1], s[2] * n[0] - s[0] * n[2], s[0] * n[1] - s[1] * n[0], n[0], n[1], n[2]};
                for (int a = 0; a < 6; a++) { for (int c2 = 0; c2 < 6; c2++) A[a][c2] += J[a] * J[c2]; b[a] -= J[a] * r; }
            }
            if (lit_ldlt6(A, b, x) == 0) { /* TransformVector6dToMatrix4d */
                const double sa = sin(x[0]), ca = cos(x[0]), sb = sin(x[1]), cb = cos(x[1]), sg = sin(x[2]), cg = cos(x[2]);
                U[0] = cb * cg; U[1] = sa * sb * cg - ca * sg; U[2] = ca * sb * cg + sa * sg; U[3] = x[3];
                U[4] = cb * sg; U[5] = sa * sb * sg + ca * cg; U[6] = ca * sb * sg - sa * cg; U[7] = x[4];
                U[8] = -sb; U[9] = sa * cb; U[10] = ca * cb; U[11] = x[5];
            }
        }
        double Tn[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double a = 0; for (int k = 0; k < 4; k++) a += U[4 * i + k] * T[4 * k + j]; Tn[4 * i + j] = a; }
        memcpy(T, Tn, sizeof(T));
        for (int i = 0; i < ns; i++) { double q[3]; lit_apply(U, pcd + 3 * i, q); memcpy(pcd + 3 * i, q, sizeof(q)); } /* in place */
        const double pf = fit, pr = rmse;
        lit_nn_pass(g, pcd, ns, r2, corr, &cnt, &err2);
        fit = (double)cnt / ns;
        rmse = cnt > 0 ? sqrt(err2 / (double)cnt) : 0.0;
        res->iterations = it + 1;
        if (fabs(pf - fit) < rel_fit && fabs(pr - rmse) < rel_rmse) { res->converged = 1; break; }
    }
    memcpy(res->T, T, sizeof(T));
    res->fitness = fit;
    res->inlier_rmse = rmse;
    res->inlier_count = cnt;
    if (corr_out) memcpy(corr_out, corr, sizeof(int) * (size_t)ns);
    free(pcd);
    free(corr);
    lg_free(g);
    return 0;
}

/* radius-limited 1-NN on fp64 distances (index agreement reports) */
LIT_API int lit_nn1(const double *tgt, int nt, const double *q, int nq, double radius, int *idx) {
    lgrid *g = lg_build(tgt, nt, radius);
    const double r2 = radius * radius;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; i++) {
        lcand b;
        idx[i] = lg_hybrid(g, q + 3 * i, r2, 1, &b) ? b.i : -1;
    }
    lg_free(g);
    return 0;
}
