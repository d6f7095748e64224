// pcr_ransac.cu — batched RANSAC (K7 + K8), replaces registration_ransac_based_on_feature_matching as called
// from src/matcher/ransac.py:42-59, and the NumPy manual-step functions
// compute_step_transformation / evaluate_inlier_ratio(_fast) (src/matcher/ransac.py:104-277).
//
// Determinism (D6): hypothesis h draws its three correspondences from Philox4x32-10(counter = h, key = seed),
// so the hypothesis stream is independent of wave size and GPU count; the winner is the one the sequential
// single-thread Open3D loop would return (prefix-maximum replay in pcr_ransac_scan, on the host).
//
//   k_ransac_generate : one thread per hypothesis — sample, edge-length checker, 3-point Umeyama (Jacobi
//                       eigen-solve of Sigma^T Sigma in fp64), distance checker, survivors appended.
//   k_ransac_validate : one block per survivor — transform all source points (fp64 -> fp32), radius-limited
//                       1-NN in the target grid, inlier count + fixed-point sum d2 (D5), correspondence
//                       inlier count; survivors better than the running best are emitted as records.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "pcr_common.cuh"

typedef unsigned long long u64;

__device__ __forceinline__ void philox4x32_10(u64 ctr_lo, u64 ctr_hi, u64 key, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v_cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double v_dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 v_scale(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }

// cyclic Jacobi, 8 sweeps, on the symmetric 3x3 K; V columns = eigenvectors (same operation order as the oracle)
__device__ void jacobi3(double K[3][3], double V[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 8; sweep++) {
#pragma unroll
        for (int e = 0; e < 3; e++) {
            const int p = (e == 2) ? 1 : 0, q = (e == 0) ? 1 : 2;
            const double apq = K[p][q];
            if (apq == 0.0) continue;
            const double theta = (K[q][q] - K[p][p]) / (2.0 * apq);
            const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0);
            const double s = t * c;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double kp = K[k][p], kq = K[k][q];
                K[k][p] = c * kp - s * kq;
                K[k][q] = s * kp + c * kq;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double pk = K[p][k], qk = K[q][k];
                K[p][k] = c * pk - s * qk;
                K[q][k] = s * pk + c * qk;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - s * vq;
                V[k][q] = s * vp + c * vq;
            }
        }
    }
}

// three source / target points (rows) -> T (12 doubles, rows 0..2 of the 4x4).  Always finite.
__device__ void rigid3(const double s[3][3], const double t[3][3], double *T) {
    double ms[3], mt[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        ms[d] = ((s[0][d] + s[1][d]) + s[2][d]) / 3.0;
        mt[d] = ((t[0][d] + t[1][d]) + t[2][d]) / 3.0;
    }
    double a[3][3], b[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            a[k][d] = s[k][d] - ms[d];
            b[k][d] = t[k][d] - mt[d];
        }
    double S[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S[i][j] = ((b[0][i] * a[0][j] + b[1][i] * a[1][j]) + b[2][i] * a[2][j]) / 3.0;
    double K[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) K[i][j] = (S[0][i] * S[0][j] + S[1][i] * S[1][j]) + S[2][i] * S[2][j];
    jacobi3(K, V);
    int o0 = 0, o1 = 1, o2 = 2;
    double l0 = K[0][0], l1 = K[1][1], l2 = K[2][2];
    if (l1 > l0) { double tl = l0; l0 = l1; l1 = tl; int to = o0; o0 = o1; o1 = to; }
    if (l2 > l1) { double tl = l1; l1 = l2; l2 = tl; int to = o1; o1 = o2; o2 = to; }
    if (l1 > l0) { double tl = l0; l0 = l1; l1 = tl; int to = o0; o0 = o1; o1 = to; }
    // select columns o0, o1 of V without dynamic register indexing
    V3 v1, v2;
    v1.x = o0 == 0 ? V[0][0] : (o0 == 1 ? V[0][1] : V[0][2]);
    v1.y = o0 == 0 ? V[1][0] : (o0 == 1 ? V[1][1] : V[1][2]);
    v1.z = o0 == 0 ? V[2][0] : (o0 == 1 ? V[2][1] : V[2][2]);
    v2.x = o1 == 0 ? V[0][0] : (o1 == 1 ? V[0][1] : V[0][2]);
    v2.y = o1 == 0 ? V[1][0] : (o1 == 1 ? V[1][1] : V[1][2]);
    v2.z = o1 == 0 ? V[2][0] : (o1 == 1 ? V[2][1] : V[2][2]);
    const V3 v3 = v_cross(v1, v2);
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    V3 w1{(S[0][0] * v1.x + S[0][1] * v1.y) + S[0][2] * v1.z, (S[1][0] * v1.x + S[1][1] * v1.y) + S[1][2] * v1.z,
          (S[2][0] * v1.x + S[2][1] * v1.y) + S[2][2] * v1.z};
    const double n1 = sqrt(v_dot(w1, w1));
    if (n1 > 0.0 && !isnan(n1) && !isinf(n1)) {
        const V3 u1 = v_scale(w1, 1.0 / n1);
        V3 w2{(S[0][0] * v2.x + S[0][1] * v2.y) + S[0][2] * v2.z, (S[1][0] * v2.x + S[1][1] * v2.y) + S[1][2] * v2.z,
              (S[2][0] * v2.x + S[2][1] * v2.y) + S[2][2] * v2.z};
        double pr = v_dot(w2, u1);
        w2.x -= pr * u1.x; w2.y -= pr * u1.y; w2.z -= pr * u1.z;
        double n2 = sqrt(v_dot(w2, w2));
        if (!(n2 > n1 * 1e-10)) {
            w2 = v2;
            pr = v_dot(w2, u1);
            w2.x -= pr * u1.x; w2.y -= pr * u1.y; w2.z -= pr * u1.z;
            n2 = sqrt(v_dot(w2, w2));
            if (!(n2 > 1e-6)) {
                w2 = v3;
                pr = v_dot(w2, u1);
                w2.x -= pr * u1.x; w2.y -= pr * u1.y; w2.z -= pr * u1.z;
                n2 = sqrt(v_dot(w2, w2));
            }
        }
        const V3 u2 = v_scale(w2, 1.0 / n2);
        const V3 u3 = v_cross(u1, u2);
        const double U[3][3] = {{u1.x, u2.x, u3.x}, {u1.y, u2.y, u3.y}, {u1.z, u2.z, u3.z}};
        const double W[3][3] = {{v1.x, v2.x, v3.x}, {v1.y, v2.y, v3.y}, {v1.z, v2.z, v3.z}};
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) R[i][j] = (U[i][0] * W[j][0] + U[i][1] * W[j][1]) + U[i][2] * W[j][2];
    }
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) T[4 * i + j] = R[i][j];
        T[4 * i + 3] = mt[i] - ((R[i][0] * ms[0] + R[i][1] * ms[1]) + R[i][2] * ms[2]);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) bad = bad || isnan(T[i]) || isinf(T[i]);
    if (bad) {  // src/matcher/ransac.py:184-185
#pragma unroll
        for (int i = 0; i < 12; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    }
}

struct Survivor {
    long long hyp;
    double T[12];
    long long f_cnt, f_sumq;  // totals of a completed evaluation (valid once published through bucket_best)
};

constexpr int VAL_BUCKETS = 64;  // hypothesis-index buckets of a wave for the in-wave sharing of bests

__device__ __forceinline__ void gather3(const float4 *__restrict__ src, const float4 *__restrict__ tgt,
                                        const int2 *__restrict__ corr, const int id[3], double s[3][3], double t[3][3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int2 c = __ldg(corr + id[k]);
        const float4 p = __ldg(src + c.x), q = __ldg(tgt + c.y);
        s[k][0] = p.x; s[k][1] = p.y; s[k][2] = p.z;
        t[k][0] = q.x; t[k][1] = q.y; t[k][2] = q.z;
    }
}

__global__ void __launch_bounds__(128) k_ransac_generate(const float4 *__restrict__ src, const float4 *__restrict__ tgt,
                                                         const int2 *__restrict__ corr, int c, double max_dist,
                                                         double edge_sim, long long hyp_begin, long long count, u64 seed,
                                                         Survivor *__restrict__ surv, unsigned int *__restrict__ n_surv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const long long h = hyp_begin + i;
    uint32_t r[4];
    philox4x32_10((u64)h, 0ull, seed, r);
    int id[3];
#pragma unroll
    for (int k = 0; k < 3; k++) id[k] = (int)__umulhi(r[k], (uint32_t)c);  // (r * c) >> 32, with replacement
    double s[3][3], t[3][3];
    gather3(src, tgt, corr, id, s, t);
    // CorrespondenceCheckerBasedOnEdgeLength
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = a + 1; b < 3; b++) {
            const double ax = s[a][0] - s[b][0], ay = s[a][1] - s[b][1], az = s[a][2] - s[b][2];
            const double bx = t[a][0] - t[b][0], by = t[a][1] - t[b][1], bz = t[a][2] - t[b][2];
            const double ds = sqrt((ax * ax + ay * ay) + az * az), dt = sqrt((bx * bx + by * by) + bz * bz);
            if (ds < dt * edge_sim || dt < ds * edge_sim) return;
        }
    double T[12];
    rigid3(s, t, T);
    // CorrespondenceCheckerBasedOnDistance
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double dx = (((T[0] * s[k][0] + T[1] * s[k][1]) + T[2] * s[k][2]) + T[3]) - t[k][0];
        const double dy = (((T[4] * s[k][0] + T[5] * s[k][1]) + T[6] * s[k][2]) + T[7]) - t[k][1];
        const double dz = (((T[8] * s[k][0] + T[9] * s[k][1]) + T[10] * s[k][2]) + T[11]) - t[k][2];
        if (sqrt((dx * dx + dy * dy) + dz * dz) > max_dist) return;
    }
    const unsigned int slot = atomicAdd(n_surv, 1u);
    surv[slot].hyp = h;
#pragma unroll
    for (int k = 0; k < 12; k++) surv[slot].T[k] = T[k];
}

// VAL_THREADS = 256 for pruned waves (many survivors, most of them rejected after a few chunks), 1024 for the first
// ("blind") wave, where nothing can be pruned yet and the latency of one full evaluation is what matters.
// The 256-point chunks of the Morton-ordered source are visited in a strided order (stride coprime with the number
// of chunks), so that the first few chunks already sample the whole cloud: the misses of a slightly-off hypothesis,
// which cluster at the far ends of the object, are seen early and the exact pruning rule fires after a few chunks
// instead of most of them.  The pruning rule itself does not depend on the order.
// IsBetterRANSACThan: (c1, s1) strictly better than (c0, s0)
__device__ __forceinline__ bool ransac_better(long long c1, long long s1, long long c0, long long s0) {
    return c1 > c0 || (c1 == c0 && c0 > 0 && s1 < s0);
}

// Best score among (a) the running best at wave start and (b) the completed survivors of EARLIER hypothesis buckets of
// this wave (warp-cooperative; every lane returns the result).  A hypothesis h may be pruned against any hypothesis
// g < h: when the sequential loop reaches h it holds a best at least as good as g's, so h is a prefix maximum only if it
// beats g.  bucket_best[b] is the index of the best completed survivor of bucket b (-1: none yet); a survivor is
// published (totals, fence, compare-and-swap of the index) only after its full evaluation, so a reader always sees
// consistent totals.
__device__ __forceinline__ void ransac_effective_best(const Survivor *surv, const int *bucket_best, int my_bucket, int lane,
                                                      long long best_cnt, long long best_sumq, long long *oc, long long *os) {
    long long c = best_cnt, q = best_sumq;
    for (int b = lane; b < my_bucket; b += 32) {
        const int idx = *(volatile const int *)(bucket_best + b);
        if (idx >= 0) {
            const long long c1 = *(volatile const long long *)&surv[idx].f_cnt;
            const long long q1 = *(volatile const long long *)&surv[idx].f_sumq;
            if (ransac_better(c1, q1, c, q)) { c = c1; q = q1; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long c1 = __shfl_xor_sync(0xffffffffu, c, o), q1 = __shfl_xor_sync(0xffffffffu, q, o);
        if (ransac_better(c1, q1, c, q)) { c = c1; q = q1; }
    }
    *oc = c;
    *os = q;
}

template <int VAL_THREADS, bool LISTS = false>
__global__ void __launch_bounds__(VAL_THREADS) k_ransac_validate(
    const float4 *__restrict__ src, const float4 *__restrict__ src_orig, int ms, const float4 *__restrict__ tgt, Grid g, const int2 *__restrict__ corr, int c,
    double max_dist, float r2, double sc_d, Survivor *__restrict__ surv, const unsigned int *__restrict__ n_surv,
    long long best_cnt, long long best_sumq, pcr_hyp_record *__restrict__ recs, unsigned int *__restrict__ n_recs,
    unsigned int rec_cap, int chunk_stride, unsigned int *__restrict__ next_surv, int *__restrict__ bucket_best,
    long long hyp_begin, long long hyp_count, CellLists L) {
    __shared__ double sT[12];
    __shared__ long long red[VAL_THREADS / 32][3];
    __shared__ long long s_best[2];
    __shared__ long long s_tot;
    __shared__ unsigned int s_next;
    const unsigned int ns = *n_surv;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // survivors are handed out dynamically: evaluations range from one chunk (pruned) to the whole cloud
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = atomicAdd(next_surv, 1u);
        __syncthreads();
        const unsigned int sidx = s_next;
        if (sidx >= ns) break;
        if (threadIdx.x < 12) sT[threadIdx.x] = surv[sidx].T[threadIdx.x];
        const int my_bucket = (int)(((surv[sidx].hyp - hyp_begin) * VAL_BUCKETS) / hyp_count);
        __syncthreads();
        const double *T = sT;  // broadcast shared-memory reads
        long long cnt = 0, sumq = 0, cin = 0;
        // Exact pruning against the effective best (above): stop as soon as the survivor can no longer be an
        // improvement — its count cannot reach the best count, or it can at most tie the count while its sum of
        // squared distances (which only grows) already reaches the best sum.  All quantities are block-uniform.
        int found = 0;
        long long partial = 0;
        bool pruned = false;
        const int nchunks = (ms + VAL_THREADS - 1) / VAL_THREADS;
        int done_pts = 0, chunk = 0;
        for (int kc = 0; kc < nchunks; kc++) {
            if ((kc & 3) == 0 && warp == 0) {  // refresh the effective best every 4 chunks (visible after the barriers below)
                long long ec, es;
                ransac_effective_best(surv, bucket_best, my_bucket, lane, best_cnt, best_sumq, &ec, &es);
                if (lane == 0) { s_best[0] = ec; s_best[1] = es; }
            }
            const int base = chunk * VAL_THREADS;
            chunk += chunk_stride;
            if (chunk >= nchunks) chunk -= nchunks;
            done_pts += min(VAL_THREADS, ms - base);
            const int i = base + threadIdx.x;
            bool hit = false;
            long long q_add = 0;
            if (i < ms) {
                const float4 p = __ldg(src + i);
                const float3 q = xform_pt(T, p.x, p.y, p.z);
                float d2;
                if ((LISTS ? lists_nn1(L, g, q.x, q.y, q.z, r2, &d2) : grid_nn1(g, q.x, q.y, q.z, r2, &d2)) >= 0) {
                    hit = true;
                    cnt++;
                    q_add = fixed_ll((double)d2, sc_d);
                    sumq += q_add;
                }
            }
            const long long w = warp_sum_ll(q_add);
            if (lane == 0) red[warp][0] = w;
            found += __syncthreads_count(hit);
            // the block total of this chunk: ONE warp adds the warp partials (every thread adding all of them was 8 % of the
            // kernel's instructions); everybody picks it up after the barrier that the round needs anyway
            if (warp == 0) {
                long long v = lane < VAL_THREADS / 32 ? red[lane][0] : 0;
                v = warp_sum_ll(v);
                if (lane == 0) s_tot = v;
            }
            const long long e_cnt = s_best[0], e_sumq = s_best[1];
            __syncthreads();  // s_tot is visible; red / s_best are rewritten in the next round
            partial += s_tot;
            const long long reachable = (long long)found + (ms - done_pts);
            if (reachable < e_cnt || (e_cnt > 0 && reachable == e_cnt && partial >= e_sumq)) {
                pruned = true;
                break;
            }
        }
        if (pruned) continue;  // block-uniform
        cnt = warp_sum_ll(cnt);
        sumq = warp_sum_ll(sumq);
        if (lane == 0) {
            red[warp][0] = cnt;
            red[warp][1] = sumq;
        }
        if (warp == 0) {
            long long ec, es;
            ransac_effective_best(surv, bucket_best, my_bucket, lane, best_cnt, best_sumq, &ec, &es);
            if (lane == 0) { s_best[0] = ec; s_best[1] = es; }
        }
        __syncthreads();
        long long a = 0, b = 0;
#pragma unroll
        for (int w2 = 0; w2 < VAL_THREADS / 32; w2++) { a += red[w2][0]; b += red[w2][1]; }
        // a necessary condition for being a prefix maximum of the sequential loop (block-uniform)
        const bool better = ransac_better(a, b, s_best[0], s_best[1]);
        if (threadIdx.x == 0) {
            // publish the completed evaluation to this wave's later buckets
            surv[sidx].f_cnt = a;
            surv[sidx].f_sumq = b;
            __threadfence();
            int cur = *(volatile int *)(bucket_best + my_bucket);
            for (;;) {
                if (cur >= 0 && !ransac_better(a, b, *(volatile long long *)&surv[cur].f_cnt, *(volatile long long *)&surv[cur].f_sumq)) break;
                const int old = atomicCAS(bucket_best + my_bucket, cur, (int)sidx);
                if (old == cur) break;
                cur = old;
            }
        }
        if (!better) continue;
        // correspondence-set inliers (only records need them: the early-exit estimate of the host replay)
        for (int i = threadIdx.x; i < c; i += VAL_THREADS) {
            const int2 cc = __ldg(corr + i);
            const float4 p = __ldg(src_orig + cc.x), q = __ldg(tgt + cc.y);
            const double x = p.x, y = p.y, z = p.z;
            const double dx = (((T[0] * x + T[1] * y) + T[2] * z) + T[3]) - (double)q.x;
            const double dy = (((T[4] * x + T[5] * y) + T[6] * z) + T[7]) - (double)q.y;
            const double dz = (((T[8] * x + T[9] * y) + T[10] * z) + T[11]) - (double)q.z;
            if (sqrt((dx * dx + dy * dy) + dz * dz) < max_dist) cin++;
        }
        cin = warp_sum_ll(cin);
        __syncthreads();
        if (lane == 0) red[warp][2] = cin;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long d = 0;
#pragma unroll
            for (int w2 = 0; w2 < VAL_THREADS / 32; w2++) d += red[w2][2];
            const unsigned int slot = atomicAdd(n_recs, 1u);
            if (slot < rec_cap) {
                pcr_hyp_record *o = recs + slot;
                o->hyp = surv[sidx].hyp;
                o->inlier_count = a;
                o->sum_d2_fixed = b;
                o->corr_inliers = (int)d;
                o->reserved = 0;
#pragma unroll
                for (int i = 0; i < 12; i++) o->transformation[i] = sT[i];
            }
        }
    }
}

// ---- manual-step twins -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ransac_step(const float4 *__restrict__ src, const float4 *__restrict__ tgt,
                                                     const int2 *__restrict__ corr, int c, u64 seed, long long h_begin,
                                                     int count, double *__restrict__ T_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double *o = T_out + (size_t)i * 16;
    if (c < 3) {  // src/matcher/ransac.py:138-140
        for (int k = 0; k < 16; k++) o[k] = (k % 5 == 0) ? 1.0 : 0.0;
        return;
    }
    uint32_t r[4];
    philox4x32_10((u64)(h_begin + i), 1ull, seed, r);
    int i0 = (int)__umulhi(r[0], (uint32_t)c);
    int i1 = (int)__umulhi(r[1], (uint32_t)(c - 1));
    int i2 = (int)__umulhi(r[2], (uint32_t)(c - 2));
    if (i1 >= i0) i1++;
    const int lo = min(i0, i1), hi = max(i0, i1);
    if (i2 >= lo) i2++;
    if (i2 >= hi) i2++;
    const int id[3] = {i0, i1, i2};
    double s[3][3], t[3][3], T[12];
    gather3(src, tgt, corr, id, s, t);
    rigid3(s, t, T);
    for (int k = 0; k < 12; k++) o[k] = T[k];
    o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
}

__global__ void __launch_bounds__(256) k_inlier_count(const float4 *__restrict__ src, const float4 *__restrict__ tgt,
                                                      const int2 *__restrict__ corr, int c, const double *__restrict__ Ts,
                                                      double thresh, int squared, int *__restrict__ counts) {
    __shared__ int red[8];
    const double *T = Ts + (size_t)blockIdx.x * 16;
    int cnt = 0;
    for (int i = threadIdx.x; i < c; i += 256) {
        const int2 cc = __ldg(corr + i);
        const float4 p = __ldg(src + cc.x), q = __ldg(tgt + cc.y);
        const double x = p.x, y = p.y, z = p.z;
        const double dx = (((T[0] * x + T[1] * y) + T[2] * z) + T[3]) - (double)q.x;
        const double dy = (((T[4] * x + T[5] * y) + T[6] * z) + T[7]) - (double)q.y;
        const double dz = (((T[8] * x + T[9] * y) + T[10] * z) + T[11]) - (double)q.z;
        const double d2 = (dx * dx + dy * dy) + dz * dz;
        if (squared ? (d2 < thresh) : (sqrt(d2) < thresh)) cnt++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; w++) s += red[w];
        counts[blockIdx.x] = s;
    }
}

// every correspondence must index into the two clouds (the reference raises IndexError for a bad pair; an
// out-of-range index here would be an out-of-bounds device read): flag = 1 + index of one offending pair
__global__ void __launch_bounds__(256) k_corr_check(const int2 *__restrict__ corr, int c, int ms, int mt,
                                                    unsigned int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const int2 cc = __ldg(corr + i);
    if ((unsigned int)cc.x >= (unsigned int)ms || (unsigned int)cc.y >= (unsigned int)mt) atomicMax(flag, (unsigned int)i + 1u);
}

// ---- host side ---------------------------------------------------------------------------------------------------
int pcr_corr_check_impl(pcr_ctx *ctx, const int *corr, int c, int ms, int mt) {
    if (c <= 0) return PCR_OK;
    if (!corr) return pcr_fail(ctx, PCR_ERR_INVALID, "correspondences: null pointer");
    PCR_ALLOC(flag, unsigned int, 1);
    PCR_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned int), ctx->stream));
    k_corr_check<<<div_up(c, 256), 256, 0, ctx->stream>>>((const int2 *)corr, c, ms, mt, flag);
    PCR_LAUNCHED();
    unsigned int *h = (unsigned int *)ctx->pinned;
    PCR_CUDA(cudaMemcpyAsync(h, flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*h)
        return pcr_fail(ctx, PCR_ERR_INVALID, "correspondence %u indexes outside the clouds (source has %d points, target %d)",
                        *h - 1u, ms, mt);
    return PCR_OK;
}

extern "C" int pcr_ransac_k_d(double max_dist, int ms) {
    return 62 - 2 * pcr_pow2ceil_exp(max_dist) - pcr_ilog2ceil(ms > 1 ? ms : 1);
}

int pcr_ransac_prepare(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, double max_dist,
                       RansacWork *w) {
    PCR_TRY(pcr_grid_build(ctx, tgt, mt, max_dist, nullptr, nullptr, &w->g));
    // source in cell order of a grid over itself: lanes of a warp then probe neighbouring target cells
    PCR_TRY(pcr_morton_sort(ctx, src, ms, &w->src_sorted));
    w->r2 = (float)(max_dist * max_dist);
    w->k_d = pcr_ransac_k_d(max_dist, ms);
    // validation through per-fine-cell candidate lists (pcr_celllists.cu; measured on B200: k_ransac_validate 0.84 -> 0.40 ms
    // per alignment, 10M hypotheses 162 -> 338 M hyp/s, results identical).  PCR_VAL_LISTS=0 keeps the 27-cell grid walk,
    // =3 uses cells of v/3 instead of v/2.
    w->use_lists = false;
    w->cl = CellLists{};
    static const int lists_div = getenv("PCR_VAL_LISTS") ? atoi(getenv("PCR_VAL_LISTS")) : 2;
    if (lists_div > 0) {
        bool ok = false;
        PCR_TRY(pcr_celllists_build(ctx, w->g, max_dist, lists_div >= 2 ? lists_div : 2, &w->cl, &ok));
        w->use_lists = ok;
    }
    return PCR_OK;
}

// ---- session: prepared work copied out of the per-call arena ------------------------------------------------------
int pcr_ransac_session_end_impl(pcr_ctx *ctx) {
    ctx->rsess.corr_ok = nullptr;
    ctx->rsess.active = false;  // the buffers are kept for the next session (cudaFree would synchronise the device)
    return PCR_OK;
}

int pcr_ransac_session_begin_impl(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, double max_dist) {
    pcr_ransac_session_end_impl(ctx);
    RansacWork w;
    PCR_TRY(pcr_ransac_prepare(ctx, src, ms, tgt, mt, max_dist, &w));
    const size_t ncells = (size_t)w.g.nx * w.g.ny * w.g.nz;
    // the experimental candidate lists (PCR_VAL_LISTS) are carried like the grid: header words + the whole item pool
    const size_t lcells = w.use_lists ? (size_t)w.cl.nx * w.cl.ny * w.cl.nz : 0;
    const size_t bytes[5] = {sizeof(float4) * (size_t)mt, sizeof(uint32_t) * (ncells + 1), sizeof(float4) * (size_t)ms,
                             sizeof(uint32_t) * lcells, w.use_lists ? sizeof(float4) * (size_t)w.cl.cap : 0};
    const void *from[5] = {w.g.sorted, w.g.start, w.src_sorted, w.cl.head, w.cl.items};
    for (int i = 0; i < (w.use_lists ? 5 : 3); i++) {
        if (ctx->rsess.cap[i] < bytes[i]) {
            if (ctx->rsess.bufs[i]) PCR_CUDA(cudaFree(ctx->rsess.bufs[i]));
            ctx->rsess.bufs[i] = nullptr;
            ctx->rsess.cap[i] = 0;
            PCR_CUDA(cudaMalloc(&ctx->rsess.bufs[i], bytes[i] + bytes[i] / 4));
            ctx->rsess.cap[i] = bytes[i] + bytes[i] / 4;
        }
        PCR_CUDA(cudaMemcpyAsync(ctx->rsess.bufs[i], from[i], bytes[i], cudaMemcpyDeviceToDevice, ctx->stream));
    }
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));  // the arena copies may be recycled by the next call
    w.g.sorted = (const float4 *)ctx->rsess.bufs[0];
    w.g.start = (const uint32_t *)ctx->rsess.bufs[1];
    w.src_sorted = (const float4 *)ctx->rsess.bufs[2];
    if (w.use_lists) {
        w.cl.head = (const uint32_t *)ctx->rsess.bufs[3];
        w.cl.items = (const float4 *)ctx->rsess.bufs[4];
    }
    ctx->rsess.w = w;
    ctx->rsess.src = src; ctx->rsess.tgt = tgt; ctx->rsess.ms = ms; ctx->rsess.mt = mt; ctx->rsess.max_dist = max_dist;
    ctx->rsess.active = true;
    return PCR_OK;
}

size_t pcr_wave_survivor_bytes(long long count) { return sizeof(Survivor) * (size_t)count; }

// step 1 (on ctx->stream, scratch from the arena — the caller releases it): survivors of [hyp_begin, hyp_end)
int pcr_wave_generate(pcr_ctx *ctx, const float4 *src, const float4 *tgt, const int *corr, int c, double max_dist,
                         double edge_sim, long long hyp_begin, long long hyp_end, u64 seed, int cap, WaveWork *ww) {
    const long long count = hyp_end - hyp_begin;
    ww->hyp_begin = hyp_begin;
    ww->count = count;
    ww->cap = cap;
    if (count > (1LL << 30)) return pcr_fail(ctx, PCR_ERR_INVALID, "wave too large");
    // the caller's buffers (ww->surv set: room for `count` survivors, 16 + cap records, VAL_BUCKETS ints), else arena scratch;
    // counters and records share one buffer so that the common case (few records) needs ONE D2H copy + sync
    Survivor *surv = ww->surv;
    unsigned char *hdr = ww->hdr;
    int *bucket_best = ww->bucket_best;
    if (!surv) {
        surv = arena<Survivor>(ctx, (size_t)count);
        hdr = arena<unsigned char>(ctx, 16 + sizeof(pcr_hyp_record) * (size_t)cap);
        bucket_best = arena<int>(ctx, VAL_BUCKETS);
        if (!surv || !hdr || !bucket_best) return PCR_ERR_OOM;
    }
    unsigned int *counters = (unsigned int *)hdr;
    PCR_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned int), ctx->stream));
    PCR_CUDA(cudaMemsetAsync(bucket_best, 0xff, VAL_BUCKETS * sizeof(int), ctx->stream));
    {
        KScope ks(ctx, KC_RANSAC_GENERATE, 120.0 * (double)count);
        k_ransac_generate<<<div_up(count, 128), 128, 0, ctx->stream>>>(src, tgt, (const int2 *)corr, c, max_dist, edge_sim,
                                                                       hyp_begin, count, seed, surv, counters);
        PCR_LAUNCHED();
    }
    PCR_CUDA(cudaGetLastError());
    ww->surv = surv;
    ww->hdr = hdr;
    ww->bucket_best = bucket_best;
    return PCR_OK;
}

// step 2: validation against (best_cnt, best_sumq), records to the host (sorted by hypothesis)
int pcr_wave_validate(pcr_ctx *ctx, const RansacWork &w, const float4 *src, int ms, const float4 *tgt, const int *corr, int c,
                         double max_dist, const WaveWork &ww, long long best_cnt, long long best_sumq, pcr_hyp_record *recs_host,
                         int *n_recs_host, long long *n_surv_host) {
    const long long count = ww.count, hyp_begin = ww.hyp_begin, hyp_end = ww.hyp_begin + ww.count;
    const int cap = ww.cap;
    Survivor *surv = ww.surv;
    unsigned char *hdr = ww.hdr;
    unsigned int *counters = (unsigned int *)hdr;
    pcr_hyp_record *recs = (pcr_hyp_record *)(hdr + 16);
    int *bucket_best = ww.bucket_best;
    constexpr int FIRST = 64;
    *n_recs_host = 0;
    *n_surv_host = 0;
    if (count <= 0) return PCR_OK;
    // CTA size: the duration of a wave is bounded below by ONE full evaluation (ceil(ms / threads) sequential rounds of
    // ~8 us: 290 us for 9k points with 256 threads — measured: the waves ran at 22 % warp occupancy, waiting for such
    // tails), so larger CTAs shorten every wave; smaller CTAs prune at a finer grain and pack better.
    static const int vt_env = getenv("PCR_VAL_THREADS") ? atoi(getenv("PCR_VAL_THREADS")) : 0;
    const bool blind = best_cnt <= 0;  // nothing to prune against yet: few survivors, full evaluations
    // measured: 512 threads win on small waves (latency of the few full evaluations), 256 on large, throughput-bound
    // waves (finer pruning granularity, better packing)
    const int vthreads = blind ? 1024 : (vt_env ? vt_env : (count >= 65536 ? 256 : 512));
    const int nchunks = div_up(ms, vthreads);
    int stride = (int)(nchunks * 0.618);  // stride coprime with the chunk count (1 when there are < 3 chunks)
    if (stride < 1) stride = 1;
    while (stride > 1 && std::__gcd(stride, nchunks) != 1) stride--;
    int &occ256 = ctx->occ_val256, &occ512 = ctx->occ_val512;  // per context (one thread at a time), hence per device
    if (!occ256) PCR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ256, k_ransac_validate<256, true>, 256, 0));
    if (!occ512) PCR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ512, k_ransac_validate<512, true>, 512, 0));
    const int per_sm = vthreads == 1024 ? 1 : (vthreads == 512 ? (occ512 > 0 ? occ512 : 2) : (occ256 > 0 ? occ256 : 4));
    const int vblocks = (int)std::min<long long>(count, (long long)ctx->sm_count * per_sm);
    const size_t pend_idx = ctx->pending.size();
    static const bool dbg = getenv("PCR_DEBUG") != nullptr;
    cudaEvent_t dbg_a = nullptr, dbg_b = nullptr;
    if (dbg) {
        cudaEventCreate(&dbg_a);
        cudaEventCreate(&dbg_b);
        cudaEventRecord(dbg_a, ctx->stream);
    }
    {
        KScope ks(ctx, KC_RANSAC_VALIDATE, 0.0);
#define PCR_VAL_LAUNCH(NT, LS)                                                                                                   \
        k_ransac_validate<NT, LS><<<vblocks, NT, 0, ctx->stream>>>(w.src_sorted, src, ms, tgt, w.g, (const int2 *)corr, c, max_dist,  \
                                                                   w.r2, ldexp(1.0, w.k_d), surv, counters, best_cnt, best_sumq, \
                                                                   recs, counters + 1, (unsigned int)cap, stride, counters + 2,   \
                                                                   bucket_best, hyp_begin, count, w.cl)
        if (w.use_lists) {  // EXPERIMENTAL (PCR_VAL_LISTS): same kernel, nearest neighbours through the candidate lists
            if (vthreads == 1024) PCR_VAL_LAUNCH(1024, true);
            else if (vthreads == 512) PCR_VAL_LAUNCH(512, true);
            else PCR_VAL_LAUNCH(256, true);
        } else if (vthreads == 1024) PCR_VAL_LAUNCH(1024, false);
        else if (vthreads == 512) PCR_VAL_LAUNCH(512, false);
        else PCR_VAL_LAUNCH(256, false);
#undef PCR_VAL_LAUNCH
        PCR_LAUNCHED();
    }
    PCR_CUDA(cudaGetLastError());
    if (dbg) cudaEventRecord(dbg_b, ctx->stream);
    unsigned char *hp = (unsigned char *)ctx->pinned;
    const int first = cap < FIRST ? cap : FIRST;
    PCR_CUDA(cudaMemcpyAsync(hp, hdr, 16 + sizeof(pcr_hyp_record) * (size_t)first, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
    const unsigned int *hc = (const unsigned int *)hp;
    *n_surv_host = hc[0];
    if (dbg) {
        float ms_v = 0.0f;
        cudaEventElapsedTime(&ms_v, dbg_a, dbg_b);
        fprintf(stderr, "[pcr ransac] wave [%lld, %lld): %u survivors, %u records, validate %.1f us (%d threads, %d CTAs, best_cnt %lld)\n",
                hyp_begin, hyp_end, hc[0], hc[1], ms_v * 1e3, vthreads, vblocks, best_cnt);
        cudaEventDestroy(dbg_a);
        cudaEventDestroy(dbg_b);
    }
    // algorithmic bytes of the validation launch: per survivor 16 M_s (source) + 16 M_t (target) + 8 C (pairs)
    if (ctx->profiling && pend_idx < ctx->pending.size())
        ctx->pending[pend_idx].bytes = (double)hc[0] * (16.0 * ms + 16.0 * w.g.n + 8.0 * c);
    const unsigned int nrec = hc[1];
    if (nrec > (unsigned int)cap)
        return pcr_fail(ctx, PCR_ERR_INVALID, "ransac wave produced %u records, capacity %d", nrec, cap);
    if (nrec) {
        const unsigned int got = nrec < (unsigned int)first ? nrec : (unsigned int)first;
        memcpy(recs_host, hp + 16, sizeof(pcr_hyp_record) * got);
        if (nrec > got) {
            PCR_CUDA(cudaMemcpyAsync(recs_host + got, recs + got, sizeof(pcr_hyp_record) * (nrec - got), cudaMemcpyDeviceToHost,
                                     ctx->stream));
            PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
        }
        std::sort(recs_host, recs_host + nrec, [](const pcr_hyp_record &a, const pcr_hyp_record &b) { return a.hyp < b.hyp; });
    }
    *n_recs_host = (int)nrec;
    return PCR_OK;
}

// scores [hyp_begin, hyp_end): leaves up to cap records (sorted by hypothesis) in recs_host; counts returned
int pcr_ransac_wave_impl(pcr_ctx *ctx, const RansacWork &w, const float4 *src, int ms, const float4 *tgt,
                         const int *corr, int c, double max_dist, double edge_sim, long long hyp_begin,
                         long long hyp_end, u64 seed, long long best_cnt, long long best_sumq, pcr_hyp_record *recs_host,
                         int cap, int *n_recs_host, long long *n_surv_host) {
    *n_recs_host = 0;
    *n_surv_host = 0;
    if (hyp_end - hyp_begin <= 0) return PCR_OK;
    const size_t mark_block = ctx->cur_block, mark_off = ctx->cur_off;  // wave scratch is released on return
    WaveWork ww;
    int rc = pcr_wave_generate(ctx, src, tgt, corr, c, max_dist, edge_sim, hyp_begin, hyp_end, seed, cap, &ww);
    if (rc == PCR_OK) rc = pcr_wave_validate(ctx, w, src, ms, tgt, corr, c, max_dist, ww, best_cnt, best_sumq, recs_host, n_recs_host, n_surv_host);
    ctx->cur_block = mark_block;
    ctx->cur_off = mark_off;
    return rc;
}

extern "C" int pcr_ransac_scan(const pcr_hyp_record *recs, int n, int64_t hyp_begin, int64_t hyp_end, int c, int ms,
                               double confidence, int32_t k_d, pcr_reg_result *st, int *stop) {
    // Replays `for itr in [hyp_begin, hyp_end): if itr >= est_k: break; ...` of A.6 over the survivors that can
    // change the state (records sorted by hyp).  st->est_k must be initialised to max_iter by the caller.
    *stop = 0;
    int64_t h_set = -1;
    for (int i = 0; i < n; i++) {
        const pcr_hyp_record &r = recs[i];
        if (r.hyp < hyp_begin || r.hyp >= hyp_end) continue;
        if (r.hyp >= st->est_k) break;
        const bool better = r.inlier_count > st->inlier_count ||
                            (r.inlier_count == st->inlier_count && st->inlier_count > 0 && r.sum_d2_fixed < st->sum_d2_fixed);
        if (!better) continue;
        st->inlier_count = r.inlier_count;
        st->sum_d2_fixed = r.sum_d2_fixed;
        st->best_hyp = r.hyp;
        for (int k = 0; k < 12; k++) st->transformation[k] = r.transformation[k];
        st->transformation[12] = st->transformation[13] = st->transformation[14] = 0.0;
        st->transformation[15] = 1.0;
        const double ratio = (double)r.corr_inliers / (double)c;
        const double est = log(1.0 - confidence) / log(1.0 - ratio * ratio * ratio);
        if (est >= 0.0 && est < (double)st->est_k) {
            st->est_k = (int64_t)ceil(est);
            h_set = r.hyp;
        }
    }
    if (st->est_k <= hyp_end) {
        *stop = 1;
        int64_t ev = st->est_k;
        if (h_set + 1 > ev) ev = h_set + 1;
        if (hyp_begin > ev) ev = hyp_begin;
        st->hyp_evaluated = ev;
    } else {
        st->hyp_evaluated = hyp_end;
    }
    st->k_d = k_d;
    st->fitness = ms > 0 ? (double)st->inlier_count / (double)ms : 0.0;
    st->inlier_rmse = st->inlier_count > 0 ? sqrt(ldexp((double)st->sum_d2_fixed, -k_d) / (double)st->inlier_count) : 0.0;
    return PCR_OK;
}

static void reg_result_init(pcr_reg_result *r, int64_t max_iter) {
    memset(r, 0, sizeof(*r));
    r->transformation[0] = r->transformation[5] = r->transformation[10] = r->transformation[15] = 1.0;
    r->best_hyp = -1;
    r->est_k = max_iter;
}

int pcr_ransac_impl(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, const int *corr, int c,
                    double max_dist, double edge_sim, int64_t max_iter, double confidence, u64 seed,
                    pcr_reg_result *res, const RansacWork *prepared) {
    reg_result_init(res, max_iter);
    // Open3D returns the default result for ransac_n > |corr| or a non-positive threshold (A.6)
    if (c < 3 || !(max_dist > 0.0) || ms == 0 || mt == 0 || max_iter <= 0) return PCR_OK;
    RansacWork w;
    if (prepared) w = *prepared;  // built ahead of time by the caller (pcr_align: next to the descriptor matching)
    else PCR_TRY(pcr_ransac_prepare(ctx, src, ms, tgt, mt, max_dist, &w));
    std::vector<pcr_hyp_record> recs;
    static const int wave_first = getenv("PCR_WAVE_FIRST") ? atoi(getenv("PCR_WAVE_FIRST")) : 2048;
    static const int wave_growth = getenv("PCR_WAVE_GROWTH") ? atoi(getenv("PCR_WAVE_GROWTH")) : 64;
    int64_t begin = 0, wave = wave_first;  // small blind first wave (no best to prune against yet), then growing to fill the GPU
    int64_t survivors = 0;
    // The hypotheses of the SECOND wave are generated on another stream while the first wave's survivors are validated
    // (generation needs no result of the first wave; its validation does): ~55 us of the 100k-hypothesis run leave the
    // critical path.  Only with confidence 1.0, where the run cannot stop early and the second wave is certain to be needed:
    // with the reference's 0.999 the first wave usually ends the run, and the discarded generation (which the call must
    // still wait for) was measured to cost 0.04 ms per alignment.  PCR_RANSAC_SPECULATE=0: never, =2: always.
    static const int spec_mode = getenv("PCR_RANSAC_SPECULATE") ? atoi(getenv("PCR_RANSAC_SPECULATE")) : 1;
    const bool speculate = spec_mode == 2 || (spec_mode == 1 && confidence >= 1.0);
    const size_t mark_block = ctx->cur_block, mark_off = ctx->cur_off;
    WaveWork spec;
    cudaEvent_t spec_done = nullptr;
    int wave_idx = 0;
    int rc_all = PCR_OK;
    while (begin < max_iter && begin < res->est_k) {
        const int64_t end = std::min<int64_t>(max_iter, begin + wave);
        const int64_t next_wave = std::min<int64_t>(wave * wave_growth, 1 << 20);
        int cap = 4096;
        int nrec = 0;
        long long nsurv = 0;
        recs.resize((size_t)cap);
        int rc;
        if (wave_idx == 0 && speculate && ctx->aux2_stream && ctx->aux2_stream != ctx->stream && end < max_iter) {
            WaveWork w0;
            rc = pcr_wave_generate(ctx, src, tgt, corr, c, max_dist, edge_sim, begin, end, seed, cap, &w0);
            if (rc == PCR_OK) {
                cudaEvent_t in_ready;
                cudaEventCreateWithFlags(&in_ready, cudaEventDisableTiming);
                cudaEventRecord(in_ready, ctx->stream);  // the correspondences and clouds are complete at this point of the stream
                cudaStreamWaitEvent(ctx->aux2_stream, in_ready, 0);
                cudaEventDestroy(in_ready);
                cudaStream_t keep = ctx->stream;
                ctx->stream = ctx->aux2_stream;
                const int rcs = pcr_wave_generate(ctx, src, tgt, corr, c, max_dist, edge_sim, end, std::min<int64_t>(max_iter, end + next_wave), seed,
                                              4096, &spec);
                ctx->stream = keep;
                if (rcs == PCR_OK && cudaEventCreateWithFlags(&spec_done, cudaEventDisableTiming) == cudaSuccess)
                    cudaEventRecord(spec_done, ctx->aux2_stream);
                else cudaStreamSynchronize(ctx->aux2_stream);  // no speculation: the wave is generated again in its turn
                rc = pcr_wave_validate(ctx, w, src, ms, tgt, corr, c, max_dist, w0, res->inlier_count, res->sum_d2_fixed, recs.data(), &nrec, &nsurv);
            }
        } else if (wave_idx == 1 && spec_done) {
            cudaStreamWaitEvent(ctx->stream, spec_done, 0);
            cudaEventDestroy(spec_done);
            spec_done = nullptr;
            rc = pcr_wave_validate(ctx, w, src, ms, tgt, corr, c, max_dist, spec, res->inlier_count, res->sum_d2_fixed, recs.data(), &nrec, &nsurv);
        } else {
            rc = pcr_ransac_wave_impl(ctx, w, src, ms, tgt, corr, c, max_dist, edge_sim, begin, end, seed, res->inlier_count,
                                      res->sum_d2_fixed, recs.data(), cap, &nrec, &nsurv);
        }
        // a wave whose records did not fit: generation + validation again with a larger buffer
        while (rc == PCR_ERR_INVALID && cap < (1 << 24) && (int64_t)cap < end - begin) {
            cap *= 16;
            recs.resize((size_t)cap);
            rc = pcr_ransac_wave_impl(ctx, w, src, ms, tgt, corr, c, max_dist, edge_sim, begin, end, seed, res->inlier_count,
                                      res->sum_d2_fixed, recs.data(), cap, &nrec, &nsurv);
        }
        if (rc != PCR_OK) {
            rc_all = rc;
            break;
        }
        if (wave_idx >= 1 && !spec_done) {  // the scratch of the first two waves is released once both are done
            ctx->cur_block = mark_block;
            ctx->cur_off = mark_off;
        }
        survivors += nsurv;
        int stop = 0;
        pcr_ransac_scan(recs.data(), nrec, begin, end, c, ms, confidence, w.k_d, res, &stop);
        begin = end;
        wave_idx++;
        if (stop) break;
        // completed evaluations are shared inside a wave (bucket_best), so the second wave can be large: measured
        // (2048, x8) 1.22 ms, (2048, x64) 1.09 ms, (4096, x64) 1.12 ms, (1024, x128) 1.07 ms for 100k hypotheses
        wave = next_wave;
    }
    if (spec_done) {  // a speculated wave that is not needed (early stop, error): its kernels still use this call's arena
        cudaStreamWaitEvent(ctx->stream, spec_done, 0);
        cudaEventDestroy(spec_done);
    }
    if (rc_all != PCR_OK) return rc_all;
    res->survivors = survivors;
    res->k_d = w.k_d;
    if (res->hyp_evaluated > max_iter) res->hyp_evaluated = max_iter;
    return PCR_OK;
}

int pcr_ransac_step_impl(pcr_ctx *ctx, const float4 *src, const float4 *tgt, const int *corr, int c, u64 seed,
                         long long h_begin, int count, double *T) {
    if (count <= 0) return PCR_OK;
    KScope ks(ctx, KC_RANSAC_STEP, 120.0 * (double)count);
    k_ransac_step<<<div_up(count, 128), 128, 0, ctx->stream>>>(src, tgt, (const int2 *)corr, c, seed, h_begin, count, T);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

int pcr_inlier_count_impl(pcr_ctx *ctx, const float4 *src, const float4 *tgt, const int *corr, int c, const double *T,
                          int count, double thresh, int squared, int *counts) {
    if (count <= 0) return PCR_OK;
    KScope ks(ctx, KC_RANSAC_STEP, 40.0 * (double)c * count);
    k_inlier_count<<<count, 256, 0, ctx->stream>>>(src, tgt, (const int2 *)corr, c, T, thresh, squared, counts);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
