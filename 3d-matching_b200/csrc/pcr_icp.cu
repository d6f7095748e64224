// pcr_icp.cu — point-to-plane ICP (K9 + K10 + K11 fused), replaces open3d registration_icp as called from
// src/matcher/icp.py:42-48.
//
// Two kernel launches per NN pass, no host synchronisation between passes:
//   k_icp_nn    — every (Morton-ordered) source point is transformed by the cumulative fp64 transform (D7) and its
//                 radius-limited nearest target point is found in the uniform grid.  The search is a chain of
//                 dependent loads (cell range -> candidates), i.e. memory-latency bound, so this kernel is kept
//                 lean (few registers, maximum occupancy) to keep many loads in flight.
//   k_icp_accum — the point-to-plane normal equations (21 + 6 sums), the inlier count and the sum of squared
//                 distances are accumulated as int64 fixed point (D5: order-free, so warp shuffles and atomics
//                 give bit-reproducible sums); the last block to finish solves the 6x6 system (LDL^T), composes
//                 the Euler-ZYX update, applies the convergence test and publishes the next transform.
//
// HBM roofline: 16 B source + 16 B target + 16 B normal + 4 B index = 52 B per point and pass.
#include "pcr_common.cuh"

struct IcpState {
    double T[16];
    long long acc[32];  // 0..20 JtJ upper triangle (row-major), 21..26 Jtr, 27 count, 28 sum d2
    unsigned int ticket;
    int pass;
    int done;
    int iterations;
    int converged;
    int max_iter;
    double prev_fit, prev_rmse;
    double fitness, rmse;
    long long count, sumq;
    double rel_fit, rel_rmse;
    double sc_JJ, sc_Jr, sc_d;       // 2^k scales
    double isc_JJ, isc_Jr, isc_d;    // 2^-k
    double T_out[16];
};

__device__ int ldlt6_solve_dev(const double A[6][6], const double *b, double *x) {
    double L[6][6], d[6], y[6];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) L[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double dj = A[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) dj = dj - (L[j][k] * L[j][k]) * d[k];
        bad = bad || !(dj > 0.0) || isinf(dj);
        d[j] = dj;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double v = A[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) v = v - (L[i][k] * L[j][k]) * d[k];
            L[i][j] = v / dj;
        }
    }
    if (bad) return -1;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) v = v - L[i][k] * y[k];
        y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = y[i] / d[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double v = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) v = v - L[k][i] * x[k];
        x[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (isnan(x[i]) || isinf(x[i])) return -1;
    return 0;
}

// runs in ONE thread of the last block of a pass
__device__ void icp_finish_pass(IcpState *S, int ns) {
    const long long cnt = S->acc[27], sumq = S->acc[28];
    const double fit = (double)cnt / (double)ns;
    const double rmse = cnt > 0 ? sqrt(((double)sumq * S->isc_d) / (double)cnt) : 0.0;
    S->fitness = fit;
    S->rmse = rmse;
    S->count = cnt;
    S->sumq = sumq;
    for (int i = 0; i < 16; i++) S->T_out[i] = S->T[i];
    const int pass = S->pass;
    bool stop = false;
    if (pass > 0 && fabs(S->prev_fit - fit) < S->rel_fit && fabs(S->prev_rmse - rmse) < S->rel_rmse) {
        S->converged = 1;
        stop = true;
    }
    if (!stop && pass >= S->max_iter) stop = true;
    if (stop) {
        S->done = 1;
    } else {
        S->prev_fit = fit;
        S->prev_rmse = rmse;
        double U[16];
        for (int i = 0; i < 16; i++) U[i] = (i % 5 == 0) ? 1.0 : 0.0;
        if (cnt > 0) {
            double A[6][6], b[6], x[6];
            int e = 0;
            for (int a = 0; a < 6; a++)
                for (int c = a; c < 6; c++) {
                    const double v = (double)S->acc[e] * S->isc_JJ;
                    A[a][c] = v;
                    A[c][a] = v;
                    e++;
                }
            for (int a = 0; a < 6; a++) b[a] = -((double)S->acc[21 + a] * S->isc_Jr);
            if (ldlt6_solve_dev(A, b, x) == 0) {
                double sa, ca, sb, cb, sg, cg;
                pcr_sincos(x[0], &sa, &ca);
                pcr_sincos(x[1], &sb, &cb);
                pcr_sincos(x[2], &sg, &cg);
                U[0] = cb * cg;  U[1] = (sa * sb) * cg - ca * sg;  U[2] = (ca * sb) * cg + sa * sg;  U[3] = x[3];
                U[4] = cb * sg;  U[5] = (sa * sb) * sg + ca * cg;  U[6] = (ca * sb) * sg - sa * cg;  U[7] = x[4];
                U[8] = -sb;      U[9] = sa * cb;                   U[10] = ca * cb;                  U[11] = x[5];
            }
        }
        double Tn[16];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++) {
                double v = (U[4 * i] * S->T[j] + U[4 * i + 1] * S->T[4 + j]) + U[4 * i + 2] * S->T[8 + j];
                if (j == 3) v = v + U[4 * i + 3];
                Tn[4 * i + j] = v;
            }
        Tn[12] = Tn[13] = Tn[14] = 0.0;
        Tn[15] = 1.0;
        for (int i = 0; i < 16; i++) S->T[i] = Tn[i];
        S->iterations = pass + 1;
    }
    S->pass = pass + 1;
    for (int i = 0; i < 32; i++) S->acc[i] = 0;
    S->ticket = 0;
}

constexpr int ICP_THREADS = 256;

// Correspondence of one transformed source point q (index i in Morton order): certified reuse or full search; keeps
// seed[i] / cert[i] up to date (see k_icp_nn).  Returns the target index or -1 and the fp32 squared distance.
__device__ __forceinline__ int icp_point_nn(const Grid &g, const float4 *__restrict__ tgt, float3 q, float r2, float rc2, int pass,
                                            int i, int *__restrict__ seed, float4 *__restrict__ cert, float *d2_out) {
    const int j_old = seed[i];
    if (j_old >= 0 && pass >= 3) {
        const float4 c = cert[i];
        if (c.w > 0.0f) {
            const float4 t = __ldg(tgt + j_old);
            const float d1 = dist2f(q.x, q.y, q.z, t.x, t.y, t.z);
            const float mx = q.x - c.x, my = q.y - c.y, mz = q.z - c.z;
            const float delta = sqrtf((mx * mx + my * my) + mz * mz);
            if ((sqrtf(d1) + delta) * 1.00002f < sqrtf(c.w) * 0.99998f) {
                const int j = d1 < r2 ? j_old : -1;
                seed[i] = j;
                *d2_out = d1;
                return j;
            }
        }
    }
    float d2, second = 0.0f;
    int j;
    if (pass >= 2) j = grid_nn1_cert(g, q.x, q.y, q.z, r2, rc2, &d2, &second);
    else j = grid_nn1(g, q.x, q.y, q.z, r2, &d2);
    seed[i] = j;
    if (pass >= 2) cert[i] = make_float4(q.x, q.y, q.z, j >= 0 ? second : 0.0f);
    *d2_out = d2;
    return j;
}

// point-to-plane normal equations of one correspondence, int64 fixed point (rule D5)
__device__ __forceinline__ void icp_accumulate(long long *acc, float3 q, float d2, float4 tp, float4 np, double scJJ, double scJr,
                                               double scd) {
    const double sx = q.x, sy = q.y, sz = q.z;
    const double nx = np.x, ny = np.y, nz = np.z;
    const double r = ((sx - (double)tp.x) * nx + (sy - (double)tp.y) * ny) + (sz - (double)tp.z) * nz;
    double J[6];
    J[0] = sy * nz - sz * ny;
    J[1] = sz * nx - sx * nz;
    J[2] = sx * ny - sy * nx;
    J[3] = nx;
    J[4] = ny;
    J[5] = nz;
    int e = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int c = a; c < 6; c++) acc[e++] += fixed_ll(J[a] * J[c], scJJ);
#pragma unroll
    for (int a = 0; a < 6; a++) acc[21 + a] += fixed_ll(J[a] * r, scJr);
    acc[27] += 1;
    acc[28] += fixed_ll((double)d2, scd);
}

// NN half of a pass.  Per source point (Morton order) the state is seed[i] = current correspondence (-1: none) and
// cert[i] = (query position at the last full search, lower bound of the squared distance to every OTHER target point).
// A pass first tries to CERTIFY the previous correspondence j: if dist(q', t_j) + |q' - q_ref| is below the distance
// bound of all other points (triangle inequality, 2e-5 relative slack >> fp32 rounding), j is still the unique
// nearest neighbour, so d2 = fp32 dist2(q', t_j) is exactly what the search would return and the search is skipped.
// Otherwise the full grid search runs (from pass 2 on it also refreshes the certificate).  Results are identical to
// searching every pass; late ICP passes, where the update is tiny, certify nearly every point.
__global__ void __launch_bounds__(256) k_icp_nn(const float4 *__restrict__ src, int ns, Grid g, const float4 *__restrict__ tgt,
                                                float r2, const IcpState *__restrict__ S, int *__restrict__ seed,
                                                float *__restrict__ d2s, float4 *__restrict__ cert) {
    if (S->done) return;
    __shared__ double sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = S->T[threadIdx.x];
    const int pass = S->pass;
    __syncthreads();
    const float rc2 = 0.64f * (float)(g.h * g.h);  // certificate radius 0.8 cell
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(src + i);
        const float3 q = xform_pt(sT, p.x, p.y, p.z);
        float d2;
        const int j = icp_point_nn(g, tgt, q, r2, rc2, pass, i, seed, cert, &d2);
        d2s[i] = d2;
    }
}

__global__ void __launch_bounds__(ICP_THREADS) k_icp_accum(const float4 *__restrict__ src, int ns,
                                                           const float4 *__restrict__ tgt, const float4 *__restrict__ nrm,
                                                           IcpState *__restrict__ S, const int *__restrict__ seed,
                                                           const float *__restrict__ d2s) {
    if (S->done) return;
    __shared__ double sT[12];
    __shared__ long long red[ICP_THREADS / 32][29];
    __shared__ bool is_last;
    if (threadIdx.x < 12) sT[threadIdx.x] = S->T[threadIdx.x];
    const double scJJ = S->sc_JJ, scJr = S->sc_Jr, scd = S->sc_d;
    __syncthreads();
    const double *T = sT;  // broadcast shared-memory reads keep 24 registers free for the accumulators

    long long acc[29];
#pragma unroll
    for (int i = 0; i < 29; i++) acc[i] = 0;

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x) {
        const int j = seed[i];
        if (j >= 0) {
            const float4 p = __ldg(src + i);
            const float3 q = xform_pt(T, p.x, p.y, p.z);
            const float d2 = d2s[i];
            icp_accumulate(acc, q, d2, __ldg(tgt + j), __ldg(nrm + j), scJJ, scJr, scd);
        }
    }
    // block reduction (integer sums: any order gives the same bits)
#pragma unroll
    for (int i = 0; i < 29; i++) acc[i] = warp_sum_ll(acc[i]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 29; i++) red[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 29) {
        long long s = 0;
#pragma unroll
        for (int w = 0; w < ICP_THREADS / 32; w++) s += red[w][threadIdx.x];
        if (s != 0) atomicAdd((unsigned long long *)&S->acc[threadIdx.x], (unsigned long long)s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&S->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        icp_finish_pass(S, ns);
    }
}

// correspondences from Morton order back to the caller's source order
__global__ void __launch_bounds__(256) k_unpermute_corr(const float4 *__restrict__ src_sorted, const int *__restrict__ seed, int ns,
                                                        int *__restrict__ corr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ns) corr[__float_as_int(__ldg(src_sorted + i).w)] = seed[i];
}

__global__ void __launch_bounds__(256) k_nn1(const float4 *__restrict__ q, int nq, Grid g, float r2,
                                             int *__restrict__ idx, float *__restrict__ d2o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float4 p = __ldg(q + i);
    float d2;
    const int j = grid_nn1(g, p.x, p.y, p.z, r2, &d2);
    idx[i] = j;
    if (d2o) d2o[i] = j >= 0 ? d2 : 0.0f;
}

__global__ void __launch_bounds__(256) k_absmax(const float4 *__restrict__ pts, int n, unsigned int *__restrict__ out) {
    float m = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(pts + i);
        m = fmaxf(m, fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order as uints
}

// sort the source by the cells of a grid over itself so that neighbouring threads probe neighbouring cells
int pcr_sort_cloud_spatially(pcr_ctx *ctx, const float4 *pts, int n, double cell, const float4 **sorted_out) {
    Grid g;
    PCR_TRY(pcr_grid_build(ctx, pts, n, cell, nullptr, nullptr, &g));
    *sorted_out = g.sorted;
    return PCR_OK;
}

int pcr_icp_impl(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, const float4 *nrm, int nt,
                 double max_dist, const double *init, int max_iter, double rel_fit, double rel_rmse,
                 pcr_reg_result *res, int *corr, bool sync_result) {
    memset(res, 0, sizeof(*res));
    for (int i = 0; i < 16; i++) res->transformation[i] = init[i];
    res->best_hyp = -1;
    if (!(max_dist > 0.0)) return pcr_fail(ctx, PCR_ERR_INVALID, "max_correspondence_distance must be > 0");
    if (max_iter < 0) return pcr_fail(ctx, PCR_ERR_INVALID, "max_iteration must be >= 0");
    if (ns == 0 || nt == 0) {
        if (corr && ns > 0) PCR_CUDA(cudaMemsetAsync(corr, 0xff, sizeof(int) * (size_t)ns, ctx->stream));
        return PCR_OK;
    }
    // target grid (bounds needed anyway for the fixed-point scale)
    float lo[3], hi[3];
    PCR_TRY(pcr_bounds(ctx, tgt, nt, lo, hi));
    float amax = 0.0f;
    for (int d = 0; d < 3; d++) amax = fmaxf(amax, fmaxf(fabsf(lo[d]), fabsf(hi[d])));
    Grid g;
    PCR_TRY(pcr_grid_build(ctx, tgt, nt, max_dist, lo, hi, &g));
    const float4 *src_sorted = nullptr;
    PCR_TRY(pcr_morton_sort(ctx, src, ns, &src_sorted));

    const int lg = pcr_ilog2ceil(ns > 1 ? ns : 1);
    const int e_r = pcr_pow2ceil_exp(max_dist);
    const int e_J = pcr_pow2ceil_exp(2.0 * ((double)amax + max_dist) + 1.0);
    const int e_R = e_r + 2;
    const int k_d = 62 - 2 * e_r - lg, k_JJ = 62 - 2 * e_J - lg, k_Jr = 62 - e_J - e_R - lg;

    IcpState *hS = (IcpState *)ctx->pinned;
    memset(hS, 0, sizeof(IcpState));
    for (int i = 0; i < 16; i++) hS->T[i] = init[i];
    hS->max_iter = max_iter;
    hS->rel_fit = rel_fit;
    hS->rel_rmse = rel_rmse;
    hS->sc_JJ = ldexp(1.0, k_JJ); hS->isc_JJ = ldexp(1.0, -k_JJ);
    hS->sc_Jr = ldexp(1.0, k_Jr); hS->isc_Jr = ldexp(1.0, -k_Jr);
    hS->sc_d = ldexp(1.0, k_d);   hS->isc_d = ldexp(1.0, -k_d);
    PCR_ALLOC(dS, IcpState, 1);
    PCR_ALLOC(seed, int, (size_t)ns);
    PCR_ALLOC(d2s, float, (size_t)ns);
    PCR_ALLOC(cert, float4, (size_t)ns);
    PCR_CUDA(cudaMemsetAsync(seed, 0xff, sizeof(int) * (size_t)ns, ctx->stream));
    PCR_CUDA(cudaMemcpyAsync(dS, hS, sizeof(IcpState), cudaMemcpyHostToDevice, ctx->stream));
    const float r2 = (float)(max_dist * max_dist);
    const int blocks = min(div_up(ns, ICP_THREADS), ctx->sm_count * 4);
    const int nn_blocks = min(div_up(ns, 256), ctx->sm_count * 8);
    const size_t pend_idx = ctx->pending.size();
    {
        KScope ks(ctx, KC_ICP_PASS, 16.0 * ns + 32.0 * nt + 4.0 * ns, max_iter + 1);
        for (int pass = 0; pass <= max_iter; pass++) {
            k_icp_nn<<<nn_blocks, 256, 0, ctx->stream>>>(src_sorted, ns, g, tgt, r2, dS, seed, d2s, cert);
            PCR_LAUNCHED();
            k_icp_accum<<<blocks, ICP_THREADS, 0, ctx->stream>>>(src_sorted, ns, tgt, nrm, dS, seed, d2s);
            PCR_LAUNCHED();
        }
    }
    if (corr) {
        k_unpermute_corr<<<div_up(ns, 256), 256, 0, ctx->stream>>>(src_sorted, seed, ns, corr);
        PCR_LAUNCHED();
    }
    PCR_CUDA(cudaGetLastError());
    PCR_CUDA(cudaMemcpyAsync(hS, dS, sizeof(IcpState), cudaMemcpyDeviceToHost, ctx->stream));
    if (sync_result) {
        PCR_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < 16; i++) res->transformation[i] = hS->T_out[i];
        res->fitness = hS->fitness;
        res->inlier_rmse = hS->rmse;
        res->inlier_count = hS->count;
        res->sum_d2_fixed = hS->sumq;
        res->k_d = k_d;
        res->iterations = hS->iterations;
        res->converged = hS->converged;
        // passes that did work (the rest returned at the `done` check): iterations + 1
        if (ctx->profiling && pend_idx < ctx->pending.size()) ctx->pending[pend_idx].launches = hS->pass;
    }
    return PCR_OK;
}

int pcr_nn1_impl(pcr_ctx *ctx, const float4 *tgt, int nt, const float4 *q, int nq, double radius, int *idx, float *d2) {
    if (!(radius > 0.0)) return pcr_fail(ctx, PCR_ERR_INVALID, "radius must be > 0");
    if (nq == 0) return PCR_OK;
    if (nt == 0) {
        PCR_CUDA(cudaMemsetAsync(idx, 0xff, sizeof(int) * (size_t)nq, ctx->stream));
        if (d2) PCR_CUDA(cudaMemsetAsync(d2, 0, sizeof(float) * (size_t)nq, ctx->stream));
        return PCR_OK;
    }
    Grid g;
    PCR_TRY(pcr_grid_build(ctx, tgt, nt, radius, nullptr, nullptr, &g));
    KScope ks(ctx, KC_NN1, 16.0 * nt + 24.0 * nq);
    k_nn1<<<div_up(nq, 256), 256, 0, ctx->stream>>>(q, nq, g, (float)(radius * radius), idx, d2);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
