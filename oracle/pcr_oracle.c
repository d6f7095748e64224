/*
 * pcr_oracle.c — CPU ORACLE for the registration hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (3d-matching_b200/) never links, imports or calls it.
 *
 * What it restates: the arithmetic that the reference (KTC-Security-Circle/3d-matching) reaches
 * through the third-party wheel open3d==0.19.0 (pinned in the reference's uv.lock:741-742), which
 * is absent from /root/reference and cannot be installed here.  The algorithm statements follow
 * SURVEY.md Appendix A (Open3D v0.19.0 semantics) and the reference's call sites:
 *   voxel_down_sample        src/ply/ply.py:106            -> orc_voxel_downsample   (A.1)
 *   estimate_normals         src/ply/ply.py:110-112,133-135-> orc_estimate_normals   (A.2, A.3)
 *   compute_fpfh_feature     src/ply/ply.py:117-120        -> orc_fpfh               (A.4)
 *   correspondences_from_features  src/matcher/ransac.py:85-> orc_match_features     (A.5)
 *   registration_ransac_based_on_feature_matching  src/matcher/ransac.py:42-59 -> orc_ransac (A.6)
 *   registration_icp (point-to-plane)  src/matcher/icp.py:42-48 -> orc_icp_point_to_plane (A.7)
 *   compute_step_transformation   src/matcher/ransac.py:104-192 -> orc_kabsch3 / orc_ransac_step
 *   evaluate_inlier_ratio(_fast)  src/matcher/ransac.py:195-277 -> orc_inlier_ratio
 *
 * PARITY UNPINNED for the Open3D-backed functions: the reference ships no golden vectors, no
 * value assertions and no data; Open3D itself cannot run here.  The NumPy
 * functions (Kabsch / inlier ratio) ARE pinned: tests/golden/ransac_numpy_golden.npz is produced by
 * importing the reference's own functions (tests/golden/make_ransac_numpy_golden.py).
 *
 * Determinism rules (each replaces an Open3D nondeterminism):
 *   D1 inputs are quantised to fp32; 3-D squared distances are fp32: (dx*dx + dy*dy) + dz*dz.
 *   D2 neighbour order and ties: ascending (d2, index).
 *   D3 voxel output order: ascending linear voxel id (z, y, x lexicographic); sums in input order.
 *   D4 elementary functions: include/pcr_detmath.h.
 *   D5 order-free sums are int64: the inlier sum of d2 is fixed point with a power-of-two scale; the ICP normal
 *      equations quantise J (6 entries) and r once per correspondence to kq-bit integers and sum exact integer
 *      products (kq = min(30, (62 - ceil(log2 n)) / 2)).
 *   D6 RANSAC sampling: Philox4x32-10 keyed by seed, counter = global hypothesis index; result is
 *      that of the sequential (single-thread) Open3D loop.
 *   D7 ICP transforms the ORIGINAL fp32 source by the cumulative fp64 transform each pass.
 *   D8 (part) the 6x6 point-to-plane system is solved by block elimination with closed-form 3x3 inverses (solve6_block).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/pcr_detmath.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */

static inline float dist2f(const float *a, const float *b) {
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz; /* D1 */
}

static int ilog2ceil_i(long long n) { /* smallest e with 2^e >= n, n >= 1 */
    int e = 0;
    while (((long long)1 << e) < n) e++;
    return e;
}

static int pow2ceil_exp(double x) { /* smallest e with 2^e >= x, x > 0 finite */
    int e;
    const double m = frexp(x, &e); /* x = m*2^e, m in [0.5,1) */
    return (m == 0.5) ? e - 1 : e;
}

/* D7 / A.6: p' = fp32( R p + t ), products and sums individually rounded in fp64 */
static inline void xform_pt(const double *T, const float *p, float *o) {
    const double x = p[0], y = p[1], z = p[2];
    o[0] = (float)(((T[0] * x + T[1] * y) + T[2] * z) + T[3]);
    o[1] = (float)(((T[4] * x + T[5] * y) + T[6] * z) + T[7]);
    o[2] = (float)(((T[8] * x + T[9] * y) + T[10] * z) + T[11]);
}

static void mat4_identity(double *T) {
    memset(T, 0, 16 * sizeof(double));
    T[0] = T[5] = T[10] = T[15] = 1.0;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* KD-tree over (n,3) fp32 points (the reference's Open3D uses nanoflann; this is an            */
/* independent exact search structure — results are unique given rule D2)                      */
/* ------------------------------------------------------------------------------------------ */

#define KD_LEAF 12

typedef struct {
    int n;
    const float *pts;
    int *idx;     /* permutation */
    int *nlo;     /* node: first element */
    int *nhi;     /* node: one past last */
    int *left;    /* child ids, -1 for leaf */
    int *right;
    int *dim;
    float *split;
    int nnodes;
} kdtree;

static void kd_select(const float *pts, int *idx, int lo, int hi, int k, int d) {
    /* quickselect: place the k-th smallest (by coordinate d, ties by index) at position k */
    while (hi - lo > 1) {
        const int mid = lo + (hi - lo) / 2;
        const int pi = idx[mid];
        const float pv = pts[3 * pi + d];
        int i = lo, j = hi - 1;
        while (i <= j) {
            while (pts[3 * idx[i] + d] < pv || (pts[3 * idx[i] + d] == pv && idx[i] < pi)) i++;
            while (pts[3 * idx[j] + d] > pv || (pts[3 * idx[j] + d] == pv && idx[j] > pi)) j--;
            if (i <= j) {
                const int t = idx[i];
                idx[i] = idx[j];
                idx[j] = t;
                i++;
                j--;
            }
        }
        if (k <= j) hi = j + 1;
        else if (k >= i) lo = i;
        else return;
    }
}

static int kd_build_rec(kdtree *t, int lo, int hi) {
    const int id = t->nnodes++;
    t->nlo[id] = lo;
    t->nhi[id] = hi;
    t->left[id] = t->right[id] = -1;
    if (hi - lo <= KD_LEAF) return id;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = lo; i < hi; i++) {
        const float *p = t->pts + 3 * t->idx[i];
        for (int d = 0; d < 3; d++) {
            if (p[d] < mn[d]) mn[d] = p[d];
            if (p[d] > mx[d]) mx[d] = p[d];
        }
    }
    int d = 0;
    if (mx[1] - mn[1] > mx[d] - mn[d]) d = 1;
    if (mx[2] - mn[2] > mx[d] - mn[d]) d = 2;
    const int mid = lo + (hi - lo) / 2;
    kd_select(t->pts, t->idx, lo, hi, mid, d);
    t->dim[id] = d;
    t->split[id] = t->pts[3 * t->idx[mid] + d];
    /* left: [lo,mid) all <= split ; right: [mid,hi) all >= split */
    const int l = kd_build_rec(t, lo, mid);
    const int r = kd_build_rec(t, mid, hi);
    t->left[id] = l;
    t->right[id] = r;
    return id;
}

static kdtree *kd_build(const float *pts, int n) {
    kdtree *t = (kdtree *)calloc(1, sizeof(kdtree));
    t->n = n;
    t->pts = pts;
    const int cap = 2 * (n / (KD_LEAF / 2) + 2) + 4;
    t->idx = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    t->nlo = (int *)malloc(sizeof(int) * cap);
    t->nhi = (int *)malloc(sizeof(int) * cap);
    t->left = (int *)malloc(sizeof(int) * cap);
    t->right = (int *)malloc(sizeof(int) * cap);
    t->dim = (int *)malloc(sizeof(int) * cap);
    t->split = (float *)malloc(sizeof(float) * cap);
    for (int i = 0; i < n; i++) t->idx[i] = i;
    t->nnodes = 0;
    if (n > 0) kd_build_rec(t, 0, n);
    return t;
}

static void kd_free(kdtree *t) {
    if (!t) return;
    free(t->idx); free(t->nlo); free(t->nhi); free(t->left); free(t->right); free(t->dim); free(t->split);
    free(t);
}

/* bounded max-heap of (d2, idx), ordered by (d2, idx) lexicographic */
typedef struct { float d2; int idx; } cand;
static inline int cand_less(cand a, cand b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

typedef struct {
    cand *h; int size, cap; float r2; /* only candidates with d2 < r2 are eligible */
} knn_heap;

static void heap_push(knn_heap *H, cand c) {
    if (H->size < H->cap) {
        int i = H->size++;
        H->h[i] = c;
        while (i > 0) {
            const int p = (i - 1) / 2;
            if (cand_less(H->h[p], H->h[i])) { cand t = H->h[p]; H->h[p] = H->h[i]; H->h[i] = t; i = p; }
            else break;
        }
    } else if (cand_less(c, H->h[0])) {
        H->h[0] = c;
        int i = 0;
        for (;;) {
            int l = 2 * i + 1, r = l + 1, m = i;
            if (l < H->size && cand_less(H->h[m], H->h[l])) m = l;
            if (r < H->size && cand_less(H->h[m], H->h[r])) m = r;
            if (m == i) break;
            cand t = H->h[m]; H->h[m] = H->h[i]; H->h[i] = t; i = m;
        }
    }
}

static void kd_search_rec(const kdtree *t, int node, const float *q, knn_heap *H) {
    if (t->left[node] < 0) {
        for (int i = t->nlo[node]; i < t->nhi[node]; i++) {
            const int j = t->idx[i];
            const float d2 = dist2f(q, t->pts + 3 * j);
            if (d2 < H->r2) { cand c = {d2, j}; heap_push(H, c); }
        }
        return;
    }
    const int d = t->dim[node];
    const float diff = q[d] - t->split[node];
    const int near = diff <= 0.0f ? t->left[node] : t->right[node];
    const int far = diff <= 0.0f ? t->right[node] : t->left[node];
    kd_search_rec(t, near, q, H);
    const float pd2 = diff * diff; /* fp32 lower bound of any fp32 d2 on the far side */
    if (pd2 < H->r2 && (H->size < H->cap || pd2 <= H->h[0].d2)) kd_search_rec(t, far, q, H);
}

static int cand_cmp(const void *a, const void *b) {
    const cand *x = (const cand *)a, *y = (const cand *)b;
    if (cand_less(*x, *y)) return -1;
    if (cand_less(*y, *x)) return 1;
    return 0;
}

/* A.3 SearchHybrid: k nearest with d2 < r2, ascending (d2, idx).  buf must hold k cands. */
static int kd_hybrid(const kdtree *t, const float *q, float r2, int k, cand *buf) {
    knn_heap H = {buf, 0, k, r2};
    if (t->n > 0 && k > 0) kd_search_rec(t, 0, q, &H);
    qsort(buf, (size_t)H.size, sizeof(cand), cand_cmp);
    return H.size;
}

static inline float radius2_f(double radius) { return (float)(radius * radius); }

ORC_API int orc_knn_hybrid(const float *pts, int n, const float *queries, int nq, double radius, int max_nn,
                           int *out_idx, float *out_d2, int *out_cnt) {
    kdtree *t = kd_build(pts, n);
    const float r2 = radius2_f(radius);
#pragma omp parallel
    {
        cand *buf = (cand *)malloc(sizeof(cand) * (size_t)(max_nn > 0 ? max_nn : 1));
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < nq; i++) {
            const int c = kd_hybrid(t, queries + 3 * i, r2, max_nn, buf);
            out_cnt[i] = c;
            for (int k = 0; k < max_nn; k++) {
                out_idx[(size_t)i * max_nn + k] = k < c ? buf[k].idx : -1;
                out_d2[(size_t)i * max_nn + k] = k < c ? buf[k].d2 : 0.0f;
            }
        }
        free(buf);
    }
    kd_free(t);
    return 0;
}

/* radius-limited 1-NN: idx = -1 when no target point has d2 < radius^2 */
ORC_API int orc_nn1(const float *tgt, int nt, const float *queries, int nq, double radius, int *out_idx,
                    float *out_d2) {
    kdtree *t = kd_build(tgt, nt);
    const float r2 = radius2_f(radius);
#pragma omp parallel for schedule(dynamic, 512)
    for (int i = 0; i < nq; i++) {
        cand b;
        const int c = kd_hybrid(t, queries + 3 * i, r2, 1, &b);
        out_idx[i] = c ? b.idx : -1;
        out_d2[i] = c ? b.d2 : 0.0f;
    }
    kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.1 voxel down-sample (src/ply/ply.py:106)                                                  */
/* ------------------------------------------------------------------------------------------ */

typedef struct { long long key; int idx; } vkey;
static int vkey_cmp(const void *a, const void *b) {
    const vkey *x = (const vkey *)a, *y = (const vkey *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* out must hold n*3 floats.  Returns 0, or -1 on invalid voxel size / too large grid. */
ORC_API int orc_voxel_downsample(const float *pts, int n, double voxel, float *out, int *m_out) {
    *m_out = 0;
    if (!(voxel > 0.0)) return -1;
    if (n == 0) return 0;
    float mn[3] = {pts[0], pts[1], pts[2]}, mx[3] = {pts[0], pts[1], pts[2]};
    for (int i = 1; i < n; i++)
        for (int d = 0; d < 3; d++) {
            const float v = pts[3 * i + d];
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    double org[3];
    long long dims[3];
    for (int d = 0; d < 3; d++) {
        org[d] = (double)mn[d] - voxel * 0.5;
        dims[d] = (long long)floor(((double)mx[d] - org[d]) / voxel) + 1;
        if (dims[d] > 2147483647LL) return -1;
    }
    if ((double)dims[0] * (double)dims[1] * (double)dims[2] > 9.0e18) return -1;
    vkey *keys = (vkey *)malloc(sizeof(vkey) * (size_t)n);
    for (int i = 0; i < n; i++) {
        long long c[3];
        for (int d = 0; d < 3; d++) c[d] = (long long)floor(((double)pts[3 * i + d] - org[d]) / voxel);
        keys[i].key = (c[2] * dims[1] + c[1]) * dims[0] + c[0]; /* D3 */
        keys[i].idx = i;
    }
    qsort(keys, (size_t)n, sizeof(vkey), vkey_cmp);
    /* D5: per-voxel coordinate sums are int64 fixed point, scale 2^k with k = 62 - E - ceil(log2 n) where
       2^E bounds every |coordinate|; exact for all but sub-2^-k bits, and independent of summation order */
    float amax = 0.0f;
    for (int d = 0; d < 3; d++) {
        if (fabsf(mn[d]) > amax) amax = fabsf(mn[d]);
        if (fabsf(mx[d]) > amax) amax = fabsf(mx[d]);
    }
    const int E = amax > 0.0f ? pow2ceil_exp((double)amax) : 0;
    const int k = 62 - E - ilog2ceil_i(n > 1 ? n : 1);
    int m = 0;
    for (int i = 0; i < n;) {
        int j = i;
        long long s[3] = {0, 0, 0};
        while (j < n && keys[j].key == keys[i].key) {
            const float *p = pts + 3 * keys[j].idx;
            s[0] += llrint(ldexp((double)p[0], k));
            s[1] += llrint(ldexp((double)p[1], k));
            s[2] += llrint(ldexp((double)p[2], k));
            j++;
        }
        const double cnt = (double)(j - i);
        out[3 * m + 0] = (float)(ldexp((double)s[0], -k) / cnt);
        out[3 * m + 1] = (float)(ldexp((double)s[1], -k) / cnt);
        out[3 * m + 2] = (float)(ldexp((double)s[2], -k) / cnt);
        m++;
        i = j;
    }
    free(keys);
    *m_out = m;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.2 normals (src/ply/ply.py:110-112, 133-135)                                               */
/* ------------------------------------------------------------------------------------------ */

typedef struct { double x, y, z; } v3;
static inline v3 v3_cross(v3 a, v3 b) {
    v3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static inline double v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 v3_scale(v3 a, double s) { v3 r = {a.x * s, a.y * s, a.z * s}; return r; }

/* A = symmetric (a00 a01 a02 a11 a12 a22) */
static v3 eigvec0(const double *A, double ev) {
    const v3 r0 = {A[0] - ev, A[1], A[2]};
    const v3 r1 = {A[1], A[3] - ev, A[4]};
    const v3 r2 = {A[2], A[4], A[5] - ev};
    const v3 c01 = v3_cross(r0, r1), c02 = v3_cross(r0, r2), c12 = v3_cross(r1, r2);
    const double d0 = v3_dot(c01, c01), d1 = v3_dot(c02, c02), d2 = v3_dot(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    if (imax == 0) return v3_scale(c01, 1.0 / sqrt(d0));
    if (imax == 1) return v3_scale(c02, 1.0 / sqrt(d1));
    return v3_scale(c12, 1.0 / sqrt(d2));
}

static v3 eigvec1(const double *A, v3 e0, double ev1) {
    v3 U, V;
    if (fabs(e0.x) > fabs(e0.y)) {
        const double inv = 1.0 / sqrt(e0.x * e0.x + e0.z * e0.z);
        U.x = -e0.z * inv; U.y = 0.0; U.z = e0.x * inv;
    } else {
        const double inv = 1.0 / sqrt(e0.y * e0.y + e0.z * e0.z);
        U.x = 0.0; U.y = e0.z * inv; U.z = -e0.y * inv;
    }
    V = v3_cross(e0, U);
    const v3 AU = {(A[0] * U.x + A[1] * U.y) + A[2] * U.z, (A[1] * U.x + A[3] * U.y) + A[4] * U.z,
                   (A[2] * U.x + A[4] * U.y) + A[5] * U.z};
    const v3 AV = {(A[0] * V.x + A[1] * V.y) + A[2] * V.z, (A[1] * V.x + A[3] * V.y) + A[4] * V.z,
                   (A[2] * V.x + A[4] * V.y) + A[5] * V.z};
    double m00 = v3_dot(U, AU) - ev1;
    double m01 = v3_dot(U, AV);
    double m11 = v3_dot(V, AV) - ev1;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        const double mx = a00 > a01 ? a00 : a01;
        if (mx > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1.0 / sqrt(1.0 + m00 * m00); m00 *= m01; }
            v3 r = {m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z};
            return r;
        }
        return U;
    } else {
        const double mx = a11 > a01 ? a11 : a01;
        if (mx > 0) {
            if (a11 >= a01) { m01 /= m11; m11 = 1.0 / sqrt(1.0 + m01 * m01); m01 *= m11; }
            else { m11 /= m01; m01 = 1.0 / sqrt(1.0 + m11 * m11); m11 *= m01; }
            v3 r = {m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z};
            return r;
        }
        return U;
    }
}

/* smallest-eigenvalue eigenvector of the symmetric covariance C (c00 c01 c02 c11 c12 c22) */
static v3 fast_eigen3x3(const double *C) {
    double A[6];
    double mc = C[0];
    for (int i = 1; i < 6; i++) if (C[i] > mc) mc = C[i];
    if (mc == 0.0) { v3 z = {0, 0, 0}; return z; }
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
            const v3 v2 = eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            const v3 v1 = eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return v3_cross(v1, v2);
        } else {
            const v3 v0 = eigvec0(A, e0);
            if (e0 < e1 && e0 < e2) return v0;
            const v3 v1 = eigvec1(A, v0, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return v3_cross(v0, v1);
        }
    } else {
        v3 r = {0, 0, 1};
        if (C[0] < C[3] && C[0] < C[5]) { r.x = 1; r.z = 0; }
        else if (C[3] < C[0] && C[3] < C[5]) { r.y = 1; r.z = 0; }
        return r;
    }
}

static v3 normal_from_neighbours(const float *pts, const cand *nb, int cnt) {
    double C[6];
    if (cnt < 3) { C[0] = C[3] = C[5] = 1.0; C[1] = C[2] = C[4] = 0.0; }
    else {
        double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < cnt; k++) {
            const float *p = pts + 3 * nb[k].idx;
            const double x = p[0], y = p[1], z = p[2];
            cu[0] += x; cu[1] += y; cu[2] += z;
            cu[3] += x * x; cu[4] += x * y; cu[5] += x * z;
            cu[6] += y * y; cu[7] += y * z; cu[8] += z * z;
        }
        const double n = (double)cnt;
        for (int i = 0; i < 9; i++) cu[i] /= n;
        C[0] = cu[3] - cu[0] * cu[0];
        C[3] = cu[6] - cu[1] * cu[1];
        C[5] = cu[8] - cu[2] * cu[2];
        C[1] = cu[4] - cu[0] * cu[1];
        C[2] = cu[5] - cu[0] * cu[2];
        C[4] = cu[7] - cu[1] * cu[2];
    }
    v3 nrm = fast_eigen3x3(C);
    const double len = sqrt(v3_dot(nrm, nrm));
    if (len == 0.0 || !(len == len)) { nrm.x = 0; nrm.y = 0; nrm.z = 1; } /* zero or NaN -> +z */
    return nrm;
}

ORC_API int orc_estimate_normals(const float *pts, int n, double radius, int max_nn, float *normals) {
    kdtree *t = kd_build(pts, n);
    const float r2 = radius2_f(radius);
#pragma omp parallel
    {
        cand *buf = (cand *)malloc(sizeof(cand) * (size_t)(max_nn > 0 ? max_nn : 1));
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            const int c = kd_hybrid(t, pts + 3 * i, r2, max_nn, buf);
            const v3 nr = normal_from_neighbours(pts, buf, c);
            normals[3 * i + 0] = (float)nr.x;
            normals[3 * i + 1] = (float)nr.y;
            normals[3 * i + 2] = (float)nr.z;
        }
        free(buf);
    }
    kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.4 FPFH (src/ply/ply.py:117-120)                                                           */
/* ------------------------------------------------------------------------------------------ */

static void pair_features(const float *p1f, const float *n1f, const float *p2f, const float *n2f, double *f) {
    v3 p1 = {p1f[0], p1f[1], p1f[2]}, p2 = {p2f[0], p2f[1], p2f[2]};
    v3 n1 = {n1f[0], n1f[1], n1f[2]}, n2 = {n2f[0], n2f[1], n2f[2]};
    v3 d = {p2.x - p1.x, p2.y - p1.y, p2.z - p1.z};
    f[0] = f[1] = f[2] = f[3] = 0.0;
    const double len = sqrt(v3_dot(d, d));
    if (len == 0.0) return;
    f[3] = len;
    const double a1 = v3_dot(n1, d) / len;
    const double a2 = v3_dot(n2, d) / len;
    if (pcr_acos(fabs(a1)) > pcr_acos(fabs(a2))) {
        const v3 t = n1; n1 = n2; n2 = t;
        d.x = -d.x; d.y = -d.y; d.z = -d.z;
        f[2] = -a2;
    } else {
        f[2] = a1;
    }
    v3 v = v3_cross(d, n1);
    const double vn = sqrt(v3_dot(v, v));
    if (vn == 0.0) { f[0] = f[1] = f[2] = f[3] = 0.0; return; }
    v.x /= vn; v.y /= vn; v.z /= vn;
    const v3 w = v3_cross(n1, v);
    f[1] = v3_dot(v, n2);
    f[0] = pcr_atan2(v3_dot(w, n2), v3_dot(n1, n2));
}

static inline int clamp_bin(double x) {
    int h = (int)floor(x);
    if (h < 0) h = 0;
    if (h >= 11) h = 10;
    return h;
}

/* fpfh: (n,33) fp32, row = point (the reference's Feature.data is its (33,n) transpose) */
ORC_API int orc_fpfh(const float *pts, const float *normals, int n, double radius, int max_nn, float *fpfh) {
    kdtree *t = kd_build(pts, n);
    const float r2 = radius2_f(radius);
    const int K = max_nn > 0 ? max_nn : 1;
    int *nb_idx = (int *)malloc(sizeof(int) * (size_t)n * K);
    float *nb_d2 = (float *)malloc(sizeof(float) * (size_t)n * K);
    int *nb_cnt = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    double *spfh = (double *)calloc((size_t)(n > 0 ? n : 1) * 33, sizeof(double));
#pragma omp parallel
    {
        cand *buf = (cand *)malloc(sizeof(cand) * (size_t)K);
#pragma omp for schedule(dynamic, 128)
        for (int i = 0; i < n; i++) {
            const int c = kd_hybrid(t, pts + 3 * i, r2, max_nn, buf);
            nb_cnt[i] = c;
            for (int k = 0; k < c; k++) { nb_idx[(size_t)i * K + k] = buf[k].idx; nb_d2[(size_t)i * K + k] = buf[k].d2; }
            if (c > 1) {
                const double inc = 100.0 / (double)(c - 1);
                double *h = spfh + (size_t)i * 33;
                for (int k = 1; k < c; k++) { /* position 0 (the query itself) is skipped */
                    const int j = buf[k].idx;
                    double f[4];
                    pair_features(pts + 3 * i, normals + 3 * i, pts + 3 * j, normals + 3 * j, f);
                    h[clamp_bin(11.0 * (f[0] + PCR_PI) / (2.0 * PCR_PI))] += inc;
                    h[11 + clamp_bin(11.0 * (f[1] + 1.0) * 0.5)] += inc;
                    h[22 + clamp_bin(11.0 * (f[2] + 1.0) * 0.5)] += inc;
                }
            }
        }
        free(buf);
    }
#pragma omp parallel for schedule(dynamic, 128)
    for (int i = 0; i < n; i++) {
        double F[33], sum[3] = {0, 0, 0};
        for (int j = 0; j < 33; j++) F[j] = 0.0;
        const int c = nb_cnt[i];
        if (c > 1) {
            for (int k = 1; k < c; k++) {
                const double dist = (double)nb_d2[(size_t)i * K + k];
                if (dist == 0.0) continue;
                const double *hs = spfh + (size_t)nb_idx[(size_t)i * K + k] * 33;
                for (int j = 0; j < 33; j++) F[j] += hs[j] / dist;
            }
            /* D8: the normaliser of a block is the sum of its 11 accumulated bins, in bin order (Open3D keeps a running
             * sum over (neighbour, bin) instead: the same real number, rounded along another path) */
            for (int b = 0; b < 3; b++)
                for (int j = 0; j < 11; j++) sum[b] += F[11 * b + j];
            for (int b = 0; b < 3; b++) if (sum[b] != 0.0) sum[b] = 100.0 / sum[b];
            const double *hi = spfh + (size_t)i * 33;
            for (int j = 0; j < 33; j++) F[j] = F[j] * sum[j / 11] + hi[j];
        }
        for (int j = 0; j < 33; j++) fpfh[(size_t)i * 33 + j] = (float)F[j];
    }
    free(nb_idx); free(nb_d2); free(nb_cnt); free(spfh);
    kd_free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.5 feature matching (src/matcher/ransac.py:42-47, :85)                                     */
/* ------------------------------------------------------------------------------------------ */

/* 33-D distance: fp64 sequential accumulation over fp32 features; ties -> lowest target index */
ORC_API int orc_nn_features(const float *fq, int nq, const float *fb, int nb, int *nn) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < nq; i++) {
        const float *a = fq + (size_t)i * 33;
        double best = INFINITY;
        int bi = -1;
        for (int j = 0; j < nb; j++) {
            const float *b = fb + (size_t)j * 33;
            double acc = 0.0;
            int k = 0;
            for (; k < 33; k++) {
                const double df = (double)a[k] - (double)b[k];
                acc += df * df;
                if (acc > best) break; /* monotone partial sums: cannot win (ties keep lower j) */
            }
            if (k == 33 && acc < best) { best = acc; bi = j; }
        }
        nn[i] = bi;
    }
    return 0;
}

/* corr: (ms,2) int32 capacity.  mutual != 0: keep (i,j) with nn_t(j)==i; if fewer than
   mutual_ratio*ms survive, fall back to the one-directional set (A.5). */
ORC_API int orc_match_features(const float *fs, int ms, const float *ft, int mt, int mutual, double mutual_ratio,
                               int *corr, int *c_out) {
    *c_out = 0;
    if (ms == 0 || mt == 0) return 0;
    int *nn_s = (int *)malloc(sizeof(int) * (size_t)ms);
    orc_nn_features(fs, ms, ft, mt, nn_s);
    int c = 0;
    if (mutual) {
        int *nn_t = (int *)malloc(sizeof(int) * (size_t)mt);
        orc_nn_features(ft, mt, fs, ms, nn_t);
        for (int i = 0; i < ms; i++) {
            const int j = nn_s[i];
            if (j >= 0 && nn_t[j] == i) { corr[2 * c] = i; corr[2 * c + 1] = j; c++; }
        }
        free(nn_t);
        /* Open3D: int(corres_mutual.size()) >= int(mutual_consistency_ratio * num_src) — both sides truncated */
        if (c >= (int)(mutual_ratio * (double)ms)) { free(nn_s); *c_out = c; return 0; }
        c = 0;
    }
    for (int i = 0; i < ms; i++) /* a descriptor holding a NaN has no nearest neighbour (nn = -1): no pair */
        if (nn_s[i] >= 0) { corr[2 * c] = i; corr[2 * c + 1] = nn_s[i]; c++; }
    free(nn_s);
    *c_out = c;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011) — D6                                                     */
/* ------------------------------------------------------------------------------------------ */

static void philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

ORC_API void orc_philox(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key, uint32_t *out) {
    philox4x32_10(ctr_lo, ctr_hi, key, out);
}

/* ------------------------------------------------------------------------------------------ */
/* 3-point rigid estimate (Umeyama without scale, A.6; Kabsch src/matcher/ransac.py:151-181)   */
/* ------------------------------------------------------------------------------------------ */

/* cyclic Jacobi on symmetric 3x3 K (full storage), 8 sweeps; V columns = eigenvectors */
static void jacobi3(double K[3][3], double V[3][3]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    static const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int sweep = 0; sweep < 8; sweep++) {
        for (int e = 0; e < 3; e++) {
            const int p = PQ[e][0], q = PQ[e][1];
            const double apq = K[p][q];
            if (apq == 0.0) continue;
            const double theta = (K[q][q] - K[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0);
            const double s = t * c;
            /* K <- G^T K G on rows/cols p,q */
            for (int k = 0; k < 3; k++) {
                const double kp = K[k][p], kq = K[k][q];
                K[k][p] = c * kp - s * kq;
                K[k][q] = s * kp + c * kq;
            }
            for (int k = 0; k < 3; k++) {
                const double pk = K[p][k], qk = K[q][k];
                K[p][k] = c * pk - s * qk;
                K[q][k] = s * pk + c * qk;
            }
            for (int k = 0; k < 3; k++) {
                const double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - s * vq;
                V[k][q] = s * vp + c * vq;
            }
        }
    }
}

/* s[3][3], t[3][3]: three source / target points (rows).  T: 4x4 row-major.  Always finite. */
static void rigid3(const double s[3][3], const double t[3][3], double *T) {
    double ms[3], mt[3];
    for (int d = 0; d < 3; d++) {
        ms[d] = ((s[0][d] + s[1][d]) + s[2][d]) / 3.0;
        mt[d] = ((t[0][d] + t[1][d]) + t[2][d]) / 3.0;
    }
    double a[3][3], b[3][3];
    for (int k = 0; k < 3; k++) for (int d = 0; d < 3; d++) { a[k][d] = s[k][d] - ms[d]; b[k][d] = t[k][d] - mt[d]; }
    double S[3][3]; /* Sigma = (1/3) sum_k b_k a_k^T  (dst x src^T) */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) S[i][j] = ((b[0][i] * a[0][j] + b[1][i] * a[1][j]) + b[2][i] * a[2][j]) / 3.0;
    double K[3][3], V[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K[i][j] = (S[0][i] * S[0][j] + S[1][i] * S[1][j]) + S[2][i] * S[2][j];
    jacobi3(K, V);
    /* order eigenvalues descending, ties by index */
    int o0 = 0, o1 = 1, o2 = 2;
    double l0 = K[0][0], l1 = K[1][1], l2 = K[2][2];
#define SWAP_IF(la, lb, oa, ob) if (lb > la) { double tl = la; la = lb; lb = tl; int to = oa; oa = ob; ob = to; }
    SWAP_IF(l0, l1, o0, o1)
    SWAP_IF(l1, l2, o1, o2)
    SWAP_IF(l0, l1, o0, o1)
#undef SWAP_IF
    (void)l2; (void)o2;
    const v3 v1 = {V[0][o0], V[1][o0], V[2][o0]};
    const v3 v2 = {V[0][o1], V[1][o1], V[2][o1]};
    const v3 v3v = v3_cross(v1, v2);
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    v3 w1 = {(S[0][0] * v1.x + S[0][1] * v1.y) + S[0][2] * v1.z, (S[1][0] * v1.x + S[1][1] * v1.y) + S[1][2] * v1.z,
             (S[2][0] * v1.x + S[2][1] * v1.y) + S[2][2] * v1.z};
    const double n1 = sqrt(v3_dot(w1, w1));
    if (n1 > 0.0 && n1 == n1 && n1 < INFINITY) {
        const v3 u1 = v3_scale(w1, 1.0 / n1);
        v3 w2 = {(S[0][0] * v2.x + S[0][1] * v2.y) + S[0][2] * v2.z, (S[1][0] * v2.x + S[1][1] * v2.y) + S[1][2] * v2.z,
                 (S[2][0] * v2.x + S[2][1] * v2.y) + S[2][2] * v2.z};
        double pr = v3_dot(w2, u1);
        w2.x -= pr * u1.x; w2.y -= pr * u1.y; w2.z -= pr * u1.z;
        double n2 = sqrt(v3_dot(w2, w2));
        if (!(n2 > n1 * 1e-10)) { /* rank <= 1: complete the frame without twist */
            w2 = v2;
            pr = v3_dot(w2, u1);
            w2.x -= pr * u1.x; w2.y -= pr * u1.y; w2.z -= pr * u1.z;
            n2 = sqrt(v3_dot(w2, w2));
            if (!(n2 > 1e-6)) {
                w2 = v3v;
                pr = v3_dot(w2, u1);
                w2.x -= pr * u1.x; w2.y -= pr * u1.y; w2.z -= pr * u1.z;
                n2 = sqrt(v3_dot(w2, w2));
            }
        }
        const v3 u2 = v3_scale(w2, 1.0 / n2);
        const v3 u3 = v3_cross(u1, u2);
        const double U[3][3] = {{u1.x, u2.x, u3.x}, {u1.y, u2.y, u3.y}, {u1.z, u2.z, u3.z}};
        const double W[3][3] = {{v1.x, v2.x, v3v.x}, {v1.y, v2.y, v3v.y}, {v1.z, v2.z, v3v.z}};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[i][j] = (U[i][0] * W[j][0] + U[i][1] * W[j][1]) + U[i][2] * W[j][2];
    }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T[4 * i + j] = R[i][j];
        T[4 * i + 3] = mt[i] - ((R[i][0] * ms[0] + R[i][1] * ms[1]) + R[i][2] * ms[2]);
    }
    T[12] = T[13] = T[14] = 0.0;
    T[15] = 1.0;
    for (int i = 0; i < 12; i++)
        if (!(T[i] == T[i]) || T[i] == INFINITY || T[i] == -INFINITY) { mat4_identity(T); break; } /* ransac.py:184-185 */
}

ORC_API void orc_kabsch3(const double *src3, const double *tgt3, double *T) {
    double s[3][3], t[3][3];
    memcpy(s, src3, sizeof(s));
    memcpy(t, tgt3, sizeof(t));
    rigid3(s, t, T);
}

/* sample without replacement for the manual-step path (ransac.py:143): 3 distinct of c */
static void sample3_distinct(uint64_t seed, uint64_t h, int c, int out[3]) {
    uint32_t r[4];
    philox4x32_10(h, 1, seed, r);
    int i0 = (int)(((uint64_t)r[0] * (uint64_t)c) >> 32);
    int i1 = (int)(((uint64_t)r[1] * (uint64_t)(c - 1)) >> 32);
    int i2 = (int)(((uint64_t)r[2] * (uint64_t)(c - 2)) >> 32);
    if (i1 >= i0) i1++;
    int lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
    if (i2 >= lo) i2++;
    if (i2 >= hi) i2++;
    out[0] = i0; out[1] = i1; out[2] = i2;
}

/* manual-step twin of compute_step_transformation: hypothesis h of stream `seed` */
ORC_API int orc_ransac_step(const float *src, const float *tgt, const int *corr, int c, uint64_t seed, uint64_t h,
                            double *T, int *sample_out) {
    mat4_identity(T);
    if (c < 3) return 0; /* ransac.py:138-140 */
    int id[3];
    sample3_distinct(seed, h, c, id);
    double s[3][3], t[3][3];
    for (int k = 0; k < 3; k++)
        for (int d = 0; d < 3; d++) {
            s[k][d] = src[3 * corr[2 * id[k]] + d];
            t[k][d] = tgt[3 * corr[2 * id[k] + 1] + d];
        }
    if (sample_out) { sample_out[0] = id[0]; sample_out[1] = id[1]; sample_out[2] = id[2]; }
    rigid3(s, t, T);
    return 0;
}

/* evaluate_inlier_ratio (ransac.py:195-236): count of pairs with ||T s - t|| < thresh, fp64 */
ORC_API int orc_inlier_count(const float *src, const float *tgt, const int *corr, int c, const double *T, double thresh) {
    int cnt = 0;
    for (int i = 0; i < c; i++) {
        const float *p = src + 3 * corr[2 * i], *q = tgt + 3 * corr[2 * i + 1];
        const double x = p[0], y = p[1], z = p[2];
        const double dx = (((T[0] * x + T[1] * y) + T[2] * z) + T[3]) - (double)q[0];
        const double dy = (((T[4] * x + T[5] * y) + T[6] * z) + T[7]) - (double)q[1];
        const double dz = (((T[8] * x + T[9] * y) + T[10] * z) + T[11]) - (double)q[2];
        if (sqrt((dx * dx + dy * dy) + dz * dz) < thresh) cnt++;
    }
    return cnt;
}

/* evaluate_inlier_ratio_fast (ransac.py:239-277): squared distances < thresh_sq */
ORC_API int orc_inlier_count_sq(const float *src, const float *tgt, const int *corr, int c, const double *T,
                                double thresh_sq) {
    int cnt = 0;
    for (int i = 0; i < c; i++) {
        const float *p = src + 3 * corr[2 * i], *q = tgt + 3 * corr[2 * i + 1];
        const double x = p[0], y = p[1], z = p[2];
        const double dx = (((T[0] * x + T[1] * y) + T[2] * z) + T[3]) - (double)q[0];
        const double dy = (((T[4] * x + T[5] * y) + T[6] * z) + T[7]) - (double)q[1];
        const double dz = (((T[8] * x + T[9] * y) + T[10] * z) + T[11]) - (double)q[2];
        if ((dx * dx + dy * dy) + dz * dz < thresh_sq) cnt++;
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* A.6 RANSAC (src/matcher/ransac.py:20-59)                                                    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    double T[16];
    double fitness;
    double inlier_rmse;
    long long best_hyp;      /* global index of the winning hypothesis, -1 if none */
    long long inlier_count;
    long long sum_d2_fixed;  /* D5: sum of llrint(d2 * 2^k_d) */
    int k_d;
    long long hyp_evaluated; /* hypotheses consumed by the sequential loop (itr < est_k) */
    long long survivors;     /* passed both checkers among those */
    long long est_k;
} orc_ransac_result;

typedef struct { int ok; double T[16]; long long cnt, sumq; int corr_inl; } hyp_eval;

static int ransac_k_d(double max_dist, int nq) { return 62 - 2 * pow2ceil_exp(max_dist) - ilog2ceil_i(nq > 1 ? nq : 1); }

static void eval_hypothesis(long long h, uint64_t seed, const float *src, int ms, const float *tgt, const kdtree *tt,
                            const int *corr, int c, double max_dist, double edge_sim, float r2, int k_d, hyp_eval *E) {
    E->ok = 0;
    uint32_t r[4];
    philox4x32_10((uint64_t)h, 0, seed, r);
    int id[3];
    for (int k = 0; k < 3; k++) id[k] = (int)(((uint64_t)r[k] * (uint64_t)c) >> 32); /* with replacement */
    double s[3][3], t[3][3];
    for (int k = 0; k < 3; k++)
        for (int d = 0; d < 3; d++) {
            s[k][d] = src[3 * corr[2 * id[k]] + d];
            t[k][d] = tgt[3 * corr[2 * id[k] + 1] + d];
        }
    /* CorrespondenceCheckerBasedOnEdgeLength(edge_sim) */
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++) {
            const double ax = s[i][0] - s[j][0], ay = s[i][1] - s[j][1], az = s[i][2] - s[j][2];
            const double bx = t[i][0] - t[j][0], by = t[i][1] - t[j][1], bz = t[i][2] - t[j][2];
            const double ds = sqrt((ax * ax + ay * ay) + az * az), dt = sqrt((bx * bx + by * by) + bz * bz);
            if (ds < dt * edge_sim || dt < ds * edge_sim) return;
        }
    rigid3(s, t, E->T);
    const double *T = E->T;
    /* CorrespondenceCheckerBasedOnDistance(max_dist) */
    for (int k = 0; k < 3; k++) {
        const double dx = (((T[0] * s[k][0] + T[1] * s[k][1]) + T[2] * s[k][2]) + T[3]) - t[k][0];
        const double dy = (((T[4] * s[k][0] + T[5] * s[k][1]) + T[6] * s[k][2]) + T[7]) - t[k][1];
        const double dz = (((T[8] * s[k][0] + T[9] * s[k][1]) + T[10] * s[k][2]) + T[11]) - t[k][2];
        if (sqrt((dx * dx + dy * dy) + dz * dz) > max_dist) return;
    }
    /* GetRegistrationResultAndCorrespondences on the transformed source */
    long long cnt = 0, sumq = 0;
    for (int i = 0; i < ms; i++) {
        float q[3];
        xform_pt(T, src + 3 * i, q);
        cand b;
        if (kd_hybrid(tt, q, r2, 1, &b)) { cnt++; sumq += llrint(ldexp((double)b.d2, k_d)); }
    }
    E->cnt = cnt;
    E->sumq = sumq;
    E->corr_inl = orc_inlier_count(src, tgt, corr, c, T, max_dist);
    E->ok = 1;
}

ORC_API int orc_ransac(const float *src, int ms, const float *tgt, int mt, const int *corr, int c, double max_dist,
                       double edge_sim, long long max_iter, double confidence, uint64_t seed, orc_ransac_result *res) {
    memset(res, 0, sizeof(*res));
    mat4_identity(res->T);
    res->best_hyp = -1;
    res->est_k = max_iter;
    if (c < 3 || !(max_dist > 0.0) || ms == 0 || mt == 0) return 0;
    kdtree *tt = kd_build(tgt, mt);
    const float r2 = radius2_f(max_dist);
    const int k_d = ransac_k_d(max_dist, ms);
    res->k_d = k_d;
    long long est_k = max_iter, best_cnt = 0, best_sumq = 0;
    const int CH = 256;
    hyp_eval *E = (hyp_eval *)malloc(sizeof(hyp_eval) * CH);
    long long itr = 0;
    int stop = 0;
    while (itr < est_k && itr < max_iter && !stop) {
        const long long hi = (itr + CH < max_iter) ? itr + CH : max_iter;
        const int nch = (int)(hi - itr);
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < nch; k++)
            eval_hypothesis(itr + k, seed, src, ms, tgt, tt, corr, c, max_dist, edge_sim, r2, k_d, &E[k]);
        for (int k = 0; k < nch; k++) { /* the sequential loop, replayed in order */
            const long long h = itr + k;
            if (h >= est_k) { stop = 1; break; }
            res->hyp_evaluated++;
            if (!E[k].ok) continue;
            res->survivors++;
            /* IsBetterRANSACThan: fitness greater, or equal and rmse smaller (same ms -> compare
               counts; equal counts -> compare fixed-point sums).  Default best has fitness 0. */
            const int better = E[k].cnt > best_cnt || (E[k].cnt == best_cnt && best_cnt > 0 && E[k].sumq < best_sumq);
            if (!better) continue;
            best_cnt = E[k].cnt;
            best_sumq = E[k].sumq;
            res->best_hyp = h;
            memcpy(res->T, E[k].T, sizeof(res->T));
            const double ratio = (double)E[k].corr_inl / (double)c;
            const double est = log(1.0 - confidence) / log(1.0 - ratio * ratio * ratio);
            if (est >= 0.0 && est < (double)est_k) est_k = (long long)ceil(est);
        }
        itr = hi;
    }
    res->inlier_count = best_cnt;
    res->sum_d2_fixed = best_sumq;
    res->fitness = (double)best_cnt / (double)ms;
    res->inlier_rmse = best_cnt > 0 ? sqrt(ldexp((double)best_sumq, -k_d) / (double)best_cnt) : 0.0;
    res->est_k = est_k;
    free(E);
    kd_free(tt);
    return 0;
}

/* diagnostic: evaluate a single hypothesis (used by tests to compare per-hypothesis records) */
ORC_API int orc_ransac_eval_one(const float *src, int ms, const float *tgt, int mt, const int *corr, int c,
                                double max_dist, double edge_sim, uint64_t seed, long long h, double *T,
                                long long *cnt, long long *sumq, int *corr_inl) {
    kdtree *tt = kd_build(tgt, mt);
    hyp_eval E;
    eval_hypothesis(h, seed, src, ms, tgt, tt, corr, c, max_dist, edge_sim, radius2_f(max_dist),
                    ransac_k_d(max_dist, ms), &E);
    kd_free(tt);
    if (!E.ok) return 0;
    memcpy(T, E.T, sizeof(E.T));
    *cnt = E.cnt; *sumq = E.sumq; *corr_inl = E.corr_inl;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* A.7 point-to-plane ICP (src/matcher/icp.py:17-48)                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    double T[16];
    double fitness;
    double inlier_rmse;
    long long inlier_count;
    long long sum_d2_fixed;
    int k_d;
    int iterations;   /* number of update steps applied */
    int converged;
} orc_icp_result;

/* 6x6 SPD solve, rule D8 (round 2): block elimination over the rotation / translation 3x3 blocks with closed-form
 * symmetric 3x3 inverses (adjugate over determinant), every operation written out and individually rounded:
 *     A = [P Q^T; Q S],  Pi = P^-1,  W = Q Pi,  Sc = S - W Q^T,  x2 = Sc^-1 (b2 - W b1),  x1 = Pi (b1 - Q^T x2).
 * It replaces the unpivoted LDL^T of round 1: on the device that was a chain of six pivots, each behind an fp64
 * division (21 divisions, ~3,900 cycles per ICP pass on one thread while every SM waits); this form has TWO divisions
 * on its critical path and wide instruction-level parallelism.  Open3D itself calls Eigen's pivoted LDL^T
 * (oracle/pcr_oracle_literal.c does); tests/test_oracle_literal.py bounds the difference (final transform ~1e-9 against
 * the 1e-5 tolerance).  Returns 0 on success, -1 when a block is not positive (det <= 0) or the result is not finite. */
static int inv3_sym(const double m[6] /* 00 01 02 11 12 22 */, double o[6]) {
    const double c00 = m[3] * m[5] - m[4] * m[4];
    const double c01 = m[2] * m[4] - m[1] * m[5];
    const double c02 = m[1] * m[4] - m[2] * m[3];
    const double det = (m[0] * c00 + m[1] * c01) + m[2] * c02;
    if (!(det > 0.0) || det == INFINITY) return -1;
    const double c11 = m[0] * m[5] - m[2] * m[2];
    const double c12 = m[1] * m[2] - m[0] * m[4];
    const double c22 = m[0] * m[3] - m[1] * m[1];
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id; o[3] = c11 * id; o[4] = c12 * id; o[5] = c22 * id;
    return 0;
}

static int solve6_block(const double A[6][6], const double *b, double *x) {
    const double P[6] = {A[0][0], A[0][1], A[0][2], A[1][1], A[1][2], A[2][2]};
    double Pi[6];
    if (inv3_sym(P, Pi) != 0) return -1;
    const double PiF[3][3] = {{Pi[0], Pi[1], Pi[2]}, {Pi[1], Pi[3], Pi[4]}, {Pi[2], Pi[4], Pi[5]}};
    double W[3][3]; /* W = Q Pi, Q[i][k] = A[3 + i][k] */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) W[i][j] = (A[3 + i][0] * PiF[0][j] + A[3 + i][1] * PiF[1][j]) + A[3 + i][2] * PiF[2][j];
    double Sc[6]; /* upper triangle of S - W Q^T */
    int e = 0;
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++)
            Sc[e++] = A[3 + i][3 + j] - ((W[i][0] * A[3 + j][0] + W[i][1] * A[3 + j][1]) + W[i][2] * A[3 + j][2]);
    double Si[6];
    if (inv3_sym(Sc, Si) != 0) return -1;
    double r2[3];
    for (int i = 0; i < 3; i++) r2[i] = b[3 + i] - ((W[i][0] * b[0] + W[i][1] * b[1]) + W[i][2] * b[2]);
    x[3] = (Si[0] * r2[0] + Si[1] * r2[1]) + Si[2] * r2[2];
    x[4] = (Si[1] * r2[0] + Si[3] * r2[1]) + Si[4] * r2[2];
    x[5] = (Si[2] * r2[0] + Si[4] * r2[1]) + Si[5] * r2[2];
    double r1[3]; /* b1 - Q^T x2 */
    for (int k = 0; k < 3; k++) r1[k] = b[k] - ((A[3][k] * x[3] + A[4][k] * x[4]) + A[5][k] * x[5]);
    x[0] = (PiF[0][0] * r1[0] + PiF[0][1] * r1[1]) + PiF[0][2] * r1[2];
    x[1] = (PiF[1][0] * r1[0] + PiF[1][1] * r1[1]) + PiF[1][2] * r1[2];
    x[2] = (PiF[2][0] * r1[0] + PiF[2][1] * r1[1]) + PiF[2][2] * r1[2];
    for (int i = 0; i < 6; i++) if (!(x[i] == x[i]) || x[i] == INFINITY || x[i] == -INFINITY) return -1;
    return 0;
}

/* TransformVector6dToMatrix4d: R = Rz(x2) Ry(x1) Rx(x0), t = (x3,x4,x5) */
static void vec6_to_mat4(const double *x, double *U) {
    double sa, ca, sb, cb, sg, cg;
    pcr_sincos(x[0], &sa, &ca);
    pcr_sincos(x[1], &sb, &cb);
    pcr_sincos(x[2], &sg, &cg);
    U[0] = cb * cg;  U[1] = (sa * sb) * cg - ca * sg;  U[2] = (ca * sb) * cg + sa * sg;  U[3] = x[3];
    U[4] = cb * sg;  U[5] = (sa * sb) * sg + ca * cg;  U[6] = (ca * sb) * sg - sa * cg;  U[7] = x[4];
    U[8] = -sb;      U[9] = sa * cb;                   U[10] = ca * cb;                  U[11] = x[5];
    U[12] = U[13] = U[14] = 0.0; U[15] = 1.0;
}

static void mat4_mul_affine(const double *U, const double *T, double *O) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) {
            double v = (U[4 * i] * T[j] + U[4 * i + 1] * T[4 + j]) + U[4 * i + 2] * T[8 + j];
            if (j == 3) v = v + U[4 * i + 3];
            O[4 * i + j] = v;
        }
    O[12] = O[13] = O[14] = 0.0; O[15] = 1.0;
}

typedef struct { long long JJ[21], Jr[6], cnt, sumq; } icp_acc;

ORC_API int orc_icp_point_to_plane(const float *src, int ns, const float *tgt, const float *tgt_normals, int nt,
                                   double max_dist, const double *T_init, int max_iter, double rel_fitness,
                                   double rel_rmse, orc_icp_result *res, int *corr_out) {
    memset(res, 0, sizeof(*res));
    memcpy(res->T, T_init, 16 * sizeof(double));
    if (!(max_dist > 0.0)) return -1;
    if (ns == 0 || nt == 0) { if (corr_out) for (int i = 0; i < ns; i++) corr_out[i] = -1; return 0; }
    kdtree *tt = kd_build(tgt, nt);
    const float r2 = radius2_f(max_dist);
    /* D5 scales */
    float amax = 0.0f;
    for (int i = 0; i < 3 * nt; i++) { const float a = fabsf(tgt[i]); if (a > amax) amax = a; }
    const int lg = ilog2ceil_i(ns > 1 ? ns : 1);
    const int e_r = pow2ceil_exp(max_dist);
    const int e_J = pow2ceil_exp(2.0 * ((double)amax + max_dist) + 1.0);
    const int e_R = e_r + 2;
    const int k_d = 62 - 2 * e_r - lg;
    /* D5 (ICP): J and r are quantised once per correspondence to kq-bit integers; the sums are exact integer products */
    const int kq = (62 - lg) / 2 < 30 ? (62 - lg) / 2 : 30;
    const int s_J = kq - e_J, s_R = kq - e_R;
    res->k_d = k_d;
    double T[16];
    memcpy(T, T_init, sizeof(T));
    int *corr = (int *)malloc(sizeof(int) * (size_t)ns);
    double prev_fit = 0.0, prev_rmse = 0.0;
    const int nth = orc_num_threads();
    icp_acc *accs = (icp_acc *)malloc(sizeof(icp_acc) * (size_t)nth);
    for (int pass = 0;; pass++) {
        memset(accs, 0, sizeof(icp_acc) * (size_t)nth);
#pragma omp parallel
        {
#ifdef _OPENMP
            icp_acc *A = &accs[omp_get_thread_num()];
#else
            icp_acc *A = &accs[0];
#endif
#pragma omp for schedule(static)
            for (int i = 0; i < ns; i++) {
                float q[3];
                xform_pt(T, src + 3 * i, q);
                cand b;
                if (!kd_hybrid(tt, q, r2, 1, &b)) { corr[i] = -1; continue; }
                corr[i] = b.idx;
                const float *tp = tgt + 3 * b.idx, *np = tgt_normals + 3 * b.idx;
                /* point-to-plane row in fp32 (every operation individually rounded, no FMA): J and r are quantised to
                 * kq <= 30 bits right after, so fp64 here bought nothing but fp64-pipe time on the device */
                const float sx = q[0], sy = q[1], sz = q[2];
                const float nx = np[0], ny = np[1], nz = np[2];
                const float ex = sx - tp[0], ey = sy - tp[1], ez = sz - tp[2];
                const float r = ((ex * nx) + (ey * ny)) + (ez * nz);
                const float J[6] = {(sy * nz) - (sz * ny), (sz * nx) - (sx * nz), (sx * ny) - (sy * nx), nx, ny, nz};
                long long qJ[6];
                for (int a = 0; a < 6; a++) qJ[a] = llrintf(ldexpf(J[a], s_J));
                const long long qr = llrintf(ldexpf(r, s_R));
                int e = 0;
                for (int a = 0; a < 6; a++)
                    for (int c = a; c < 6; c++) A->JJ[e++] += qJ[a] * qJ[c];
                for (int a = 0; a < 6; a++) A->Jr[a] += qJ[a] * qr;
                A->cnt++;
                A->sumq += llrint(ldexp((double)b.d2, k_d));
            }
        }
        icp_acc S;
        memset(&S, 0, sizeof(S));
        for (int t = 0; t < nth; t++) {
            for (int e = 0; e < 21; e++) S.JJ[e] += accs[t].JJ[e];
            for (int e = 0; e < 6; e++) S.Jr[e] += accs[t].Jr[e];
            S.cnt += accs[t].cnt;
            S.sumq += accs[t].sumq;
        }
        const double fit = (double)S.cnt / (double)ns;
        const double rmse = S.cnt > 0 ? sqrt(ldexp((double)S.sumq, -k_d) / (double)S.cnt) : 0.0;
        res->fitness = fit;
        res->inlier_rmse = rmse;
        res->inlier_count = S.cnt;
        res->sum_d2_fixed = S.sumq;
        memcpy(res->T, T, sizeof(T));
        if (pass > 0 && fabs(prev_fit - fit) < rel_fitness && fabs(prev_rmse - rmse) < rel_rmse) { res->converged = 1; break; }
        if (pass >= max_iter) break;
        prev_fit = fit;
        prev_rmse = rmse;
        /* ComputeTransformation: solve JtJ x = -Jtr; empty set or failed solve -> identity update */
        double U[16];
        mat4_identity(U);
        if (S.cnt > 0) {
            double Am[6][6], bv[6], x[6];
            int e = 0;
            for (int a = 0; a < 6; a++)
                for (int c = a; c < 6; c++) { Am[a][c] = Am[c][a] = ldexp((double)S.JJ[e], -2 * s_J); e++; }
            for (int a = 0; a < 6; a++) bv[a] = -ldexp((double)S.Jr[a], -(s_J + s_R));
            if (solve6_block(Am, bv, x) == 0) vec6_to_mat4(x, U);
        }
        double Tn[16];
        mat4_mul_affine(U, T, Tn);
        memcpy(T, Tn, sizeof(T));
        res->iterations = pass + 1;
    }
    if (corr_out) memcpy(corr_out, corr, sizeof(int) * (size_t)ns);
    free(accs);
    free(corr);
    kd_free(tt);
    return 0;
}

/* expose the spec helpers so tests can check the device-side copies */
ORC_API int orc_pow2ceil_exp(double x) { return pow2ceil_exp(x); }
ORC_API int orc_ilog2ceil(long long n) { return ilog2ceil_i(n); }
ORC_API void orc_detmath(double x, double y, double *out) {
    out[0] = pcr_sin(x); out[1] = pcr_cos(x); out[2] = pcr_atan2(y, x); out[3] = pcr_acos(x);
}
ORC_API void orc_fast_eigen3x3(const double *C6, double *out3) {
    const v3 r = fast_eigen3x3(C6);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
ORC_API void orc_pair_features(const float *p1, const float *n1, const float *p2, const float *n2, double *f4) {
    pair_features(p1, n1, p2, n2, f4);
}
ORC_API void orc_transform_points(const double *T, const float *pts, int n, float *out) {
    for (int i = 0; i < n; i++) xform_pt(T, pts + 3 * i, out + 3 * i);
}
