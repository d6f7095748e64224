// pcr_icp.cu — point-to-plane ICP (K9 + K10 + K11 fused), replaces open3d registration_icp as called from
// src/matcher/icp.py:42-48.
//
// ONE persistent cooperative kernel (k_icp_persist) runs the whole loop: per pass, every (Morton-ordered) source point
// is transformed by the cumulative fp64 transform (D7), its radius-limited nearest target point is found (certified
// reuse of the previous correspondence where provable, else a search in the uniform grid), its point-to-plane row is
// quantised and accumulated into the 21 + 6 + 2 integer sums (D5: order-free, so shuffles and atomics give
// bit-reproducible sums); one grid barrier later every CTA solves the 6x6 system (LDL^T), composes the Euler-ZYX
// update, applies the convergence test and holds the next transform.  No host synchronisation between passes.
//
// HBM roofline: 16 B source + 16 B target + 16 B normal + 4 B index = 52 B per point and pass.
#include <cstdio>
#include <cstdlib>

#include "pcr_common.cuh"

struct IcpState {
    double T[16];
    unsigned int bar;       // grid-barrier arrival counter
    int pass;
    int done;
    int iterations;
    int converged;
    int max_iter;
    int cert_pass;         // first pass whose searches also produce certificates (PCR_ICP_CERT_PASS, default ICP_CERT_PASS)
    double prev_fit, prev_rmse;
    double fitness, rmse;
    long long count, sumq;
    double rel_fit, rel_rmse;
    double sc_J, sc_R, sc_d;         // 2^(kq - e_J), 2^(kq - e_R), 2^k_d
    double isc_JJ, isc_Jr, isc_d;    // 2^-2(kq - e_J), 2^-((kq - e_J) + (kq - e_R)), 2^-k_d
    double T_out[16];
    int trace;             // PCR_ICP_TRACE=1: collect the cycle counters below (two extra atomics per CTA and pass)
    // triple-buffered sums: 0..20 JtJ upper triangle (row-major), 21..26 Jtr, 27 count, 28 sum d2.  One 256-byte block per
    // sum (ICP_ACC_STRIDE words): every CTA adds its 29 partial sums once per pass, and atomics on one address — and on the two
    // 128-byte lines of a pair, which share an L2 slice — are applied one after the other (tools/micro/barrier_bench2.cu:
    // exchange + barrier of 592 CTAs 7,736 -> 5,883 cycles, 392 CTAs 5,595 -> 4,524)
    // (placed after the header: the host copies header + sums to the device, and only the header back)
    long long acc3[3][29 * 32 /* ICP_ACC_STRIDE */];
    long long dbgsearch[64];  // per-pass number of points that took the search path / the second tier (PCR_ICP_TRACE)
    long long dbgtier2[64];
    long long dbgmax[64];  // per-pass slowest CTA loop
    long long dbgfin[64];  // per-pass slowest CTA finish
    long long dbgp[64];  // per-pass loop cycles of CTA 0 (first 64 passes)
    long long dbg2[4];   // PCR_ICP_TRACE: inside the end-of-pass logic of CTA 0 — LDL^T, sin/cos + update, compose
    long long dbg[4];  // clock cycles of CTA 0: point loop, reduction + barrier, end-of-pass logic (PCR_ICP_TRACE=1 prints them)
    unsigned long long dbgcta[1024];  // PCR_ICP_TRACE: per CTA, loop cycles of pass 20 << 16 | SM id (copied back only with the trace on)
};

// 6x6 SPD solve, rule D8 (round 2; the same operations in the same order as oracle/pcr_oracle.c: solve6_block): block
// elimination over the rotation / translation 3x3 blocks with closed-form symmetric 3x3 inverses.  Two fp64 divisions on
// the critical path instead of the six pivots (21 divisions) of the unpivoted LDL^T it replaces — the solve runs on one
// thread per CTA while the whole GPU waits for the next transform.
__device__ __forceinline__ int inv3_sym_dev(const double *m /* 00 01 02 11 12 22 */, double *o) {
    const double c00 = m[3] * m[5] - m[4] * m[4];
    const double c01 = m[2] * m[4] - m[1] * m[5];
    const double c02 = m[1] * m[4] - m[2] * m[3];
    const double det = (m[0] * c00 + m[1] * c01) + m[2] * c02;
    if (!(det > 0.0) || isinf(det)) return -1;
    const double c11 = m[0] * m[5] - m[2] * m[2];
    const double c12 = m[1] * m[2] - m[0] * m[4];
    const double c22 = m[0] * m[3] - m[1] * m[1];
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id; o[3] = c11 * id; o[4] = c12 * id; o[5] = c22 * id;
    return 0;
}

__device__ __forceinline__ int solve6_block_dev(const double (*A)[6], const double *b, double *x) {
    const double P[6] = {A[0][0], A[0][1], A[0][2], A[1][1], A[1][2], A[2][2]};
    double Pi[6];
    if (inv3_sym_dev(P, Pi) != 0) return -1;
    const double PiF[3][3] = {{Pi[0], Pi[1], Pi[2]}, {Pi[1], Pi[3], Pi[4]}, {Pi[2], Pi[4], Pi[5]}};
    double W[3][3];  // W = Q Pi, Q[i][k] = A[3 + i][k]
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) W[i][j] = (A[3 + i][0] * PiF[0][j] + A[3 + i][1] * PiF[1][j]) + A[3 + i][2] * PiF[2][j];
    double Sc[6];  // upper triangle of S - W Q^T
    Sc[0] = A[3][3] - ((W[0][0] * A[3][0] + W[0][1] * A[3][1]) + W[0][2] * A[3][2]);
    Sc[1] = A[3][4] - ((W[0][0] * A[4][0] + W[0][1] * A[4][1]) + W[0][2] * A[4][2]);
    Sc[2] = A[3][5] - ((W[0][0] * A[5][0] + W[0][1] * A[5][1]) + W[0][2] * A[5][2]);
    Sc[3] = A[4][4] - ((W[1][0] * A[4][0] + W[1][1] * A[4][1]) + W[1][2] * A[4][2]);
    Sc[4] = A[4][5] - ((W[1][0] * A[5][0] + W[1][1] * A[5][1]) + W[1][2] * A[5][2]);
    Sc[5] = A[5][5] - ((W[2][0] * A[5][0] + W[2][1] * A[5][1]) + W[2][2] * A[5][2]);
    double Si[6];
    if (inv3_sym_dev(Sc, Si) != 0) return -1;
    double r2[3];
#pragma unroll
    for (int i = 0; i < 3; i++) r2[i] = b[3 + i] - ((W[i][0] * b[0] + W[i][1] * b[1]) + W[i][2] * b[2]);
    x[3] = (Si[0] * r2[0] + Si[1] * r2[1]) + Si[2] * r2[2];
    x[4] = (Si[1] * r2[0] + Si[3] * r2[1]) + Si[4] * r2[2];
    x[5] = (Si[2] * r2[0] + Si[4] * r2[1]) + Si[5] * r2[2];
    double r1[3];  // b1 - Q^T x2
#pragma unroll
    for (int k = 0; k < 3; k++) r1[k] = b[k] - ((A[3][k] * x[3] + A[4][k] * x[4]) + A[5][k] * x[5]);
    x[0] = (PiF[0][0] * r1[0] + PiF[0][1] * r1[1]) + PiF[0][2] * r1[2];
    x[1] = (PiF[1][0] * r1[0] + PiF[1][1] * r1[1]) + PiF[1][2] * r1[2];
    x[2] = (PiF[2][0] * r1[0] + PiF[2][1] * r1[1]) + PiF[2][2] * r1[2];
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (isnan(x[i]) || isinf(x[i])) return -1;
    return 0;
}

// Per-CTA replica of the loop state.  Every CTA runs the (deterministic) end-of-pass logic redundantly on the same
// 29 integer sums, so all CTAs hold bit-identical transforms and stop flags without a second grid barrier.
struct IcpLocal {
    double T[16];
    double prev_fit, prev_rmse, fitness, rmse;
    long long count, sumq;
    int pass, done, iterations, converged;
    long long tdbg[3];
};

// End-of-pass logic, run redundantly by every CTA on the same 29 integer sums.  It is a latency chain on the fp64 pipe
// (PCR_ICP_TRACE, cycles per pass at 100k points, round 1: an unpivoted LDL^T 3,936 — six pivots each behind an fp64
// division, 21 divisions in all —, the three sin/cos 930, compose 304, statistics ~700), not an instruction-count problem.
// The solve is now the block form above (rule D8), and the INDEPENDENT pieces
// run on different warps of the CTA instead of one after the other on one thread (round 2; same arithmetic, same bits):
//   step 1   warp 0: 6x6 solve -> x              ||  warp 1: fitness, RMSE, convergence / iteration-limit decision
//   step 2   warps 0, 1, 2: sin/cos of x[0], x[1], x[2]
//   step 3   thread 0: Euler-ZYX update U, T <- U T, bookkeeping
// A warp-parallel LDL^T (row i on lane i, shuffles) was measured slower than the serial one (7,400 vs 6,000 cycles).
struct IcpEnd {
    double x[6], sn[3], cs[3];
    int solved, stop;
};

// step 1a (one thread): x = solve(A, b); solved = 0 when there is no correspondence or the factorisation fails
__device__ __noinline__ void icp_end_solve(IcpEnd *E, const long long *acc, const double (*fA)[6], const double *fb) {
    int ok = 0;
    if (acc[27] > 0) {
        double x[6];
        if (solve6_block_dev(fA, fb, x) == 0) {
            ok = 1;
#pragma unroll
            for (int i = 0; i < 6; i++) E->x[i] = x[i];
        }
    }
    E->solved = ok;
}

// step 1b (one thread): statistics of the pass that just ended and the stop decision
__device__ __forceinline__ void icp_end_stats(IcpLocal *L, IcpEnd *E, const long long *acc, double isc_d, double rel_fit, double rel_rmse,
                                              int max_iter, int ns) {
    const long long cnt = acc[27], sumq = acc[28];
    const double fit = (double)cnt / (double)ns;
    const double rmse = cnt > 0 ? sqrt(((double)sumq * isc_d) / (double)cnt) : 0.0;
    const int pass = L->pass;
    bool stop = false;
    if (pass > 0 && fabs(L->prev_fit - fit) < rel_fit && fabs(L->prev_rmse - rmse) < rel_rmse) {
        L->converged = 1;
        stop = true;
    }
    if (!stop && pass >= max_iter) stop = true;
    L->fitness = fit;
    L->rmse = rmse;
    L->count = cnt;
    L->sumq = sumq;
    if (!stop) {
        L->prev_fit = fit;
        L->prev_rmse = rmse;
    }
    E->stop = stop ? 1 : 0;
}

// step 3 (warp 0): T <- U T, one entry of the new transform per lane (12 lanes; the same products and sums in the same
// order as the serial form, so the same bits), then the bookkeeping on lane 0
__device__ __forceinline__ void icp_end_compose(IcpLocal *L, const IcpEnd *E, int lane) {
    const int pass = L->pass;
    const bool upd = !E->stop;
    double tn = 0.0;
    if (upd && lane < 12) {
        const int i = lane >> 2, j = lane & 3;
        double u0 = i == 0 ? 1.0 : 0.0, u1 = i == 1 ? 1.0 : 0.0, u2 = i == 2 ? 1.0 : 0.0, u3 = 0.0;
        if (E->solved) {
            // branch-free over the three rows (a - b c == a + (-b) c bit for bit; multiplying by +-1 is exact):
            //   row 0: cb cg, (sa sb) cg - ca sg, (ca sb) cg + sa sg      row 1: cb sg, (sa sb) sg + ca cg, (ca sb) sg - sa cg
            //   row 2: -sb, sa cb, ca cb
            const double sa = E->sn[0], ca = E->cs[0], sb = E->sn[1], cb = E->cs[1], sg = E->sn[2], cg = E->cs[2];
            const double m = i == 0 ? cg : sg, n = i == 0 ? sg : cg, s1 = i == 0 ? -1.0 : 1.0;
            const double r0 = cb * m, r1 = (sa * sb) * m + (s1 * ca) * n, r2 = (ca * sb) * m + (-s1 * sa) * n;
            u0 = i == 2 ? -sb : r0;
            u1 = i == 2 ? sa * cb : r1;
            u2 = i == 2 ? ca * cb : r2;
            u3 = E->x[3 + i];
        }
        double v = (u0 * L->T[j] + u1 * L->T[4 + j]) + u2 * L->T[8 + j];
        if (j == 3) v = v + u3;
        tn = v;
    }
    __syncwarp();
    if (upd && lane < 16) L->T[lane] = lane < 12 ? tn : (lane == 15 ? 1.0 : 0.0);
    if (lane == 0) {
        if (E->stop) L->done = 1;
        else L->iterations = pass + 1;
        L->pass = pass + 1;
    }
}

#ifndef PCR_ICP_THREADS
#define PCR_ICP_THREADS 256
#endif
constexpr int ICP_THREADS = PCR_ICP_THREADS;  // a multiple of 128 (the accumulate phase works on groups of 4 warps x 128 rows)
#ifndef PCR_ICP_CTAS
#define PCR_ICP_CTAS (1024 / PCR_ICP_THREADS)
#endif
constexpr int ICP_CTAS_PER_SM = PCR_ICP_CTAS;
constexpr int ICP_WARPS = ICP_THREADS / 32;
// first pass whose searches also produce certificates (plain pruned searches before it, certified reuse after it)
#ifndef PCR_ICP_CERT_PASS
#define PCR_ICP_CERT_PASS 2
#endif
constexpr int ICP_CERT_PASS = PCR_ICP_CERT_PASS;  // default; PCR_ICP_CERT_PASS in the environment overrides it per call
constexpr int ICP_ACC_STRIDE = 32;                 // 8-byte words between two of the 29 sums (one 256-byte block each)

// Result of the correspondence step for one source point: target index (-1: none), fp32 squared distance, and the
// matched target point and normal.
struct IcpMatch {
    int j;
    float d2;
    float tx, ty, tz, nx, ny, nz;
};

// Correspondence of one transformed source point q (index i in Morton order): certified reuse or full search.
//
// Per-point state (written and read only by the thread that owns the point, see k_icp_persist), denormalised so that
// a certified pass is pure streaming — four coalesced 16-byte loads, no dependent gather:
//   st0[i] = (q_ref, w)   query position at the last certificate search (grid_nn1_cert, which examines every target
//                         point within one cell size h > max_dist of q_ref) and the certificate w (below)
//   st1[i] = (t_j, j)     current correspondence: target point and index (index -1: none)
//   st2[i] = (n_j, w3)    its normal; w3 belongs to the second tier
//   cert2[i] = (t_j2, j2) second tier (below): the second nearest point and its index
//   w > 0 : a correspondence j exists; w is a lower bound of the squared distance from q_ref to every OTHER target
//           point.  For the new position q, with delta = |q - q_ref|: every other point is at least sqrt(w) - delta
//           away, so if dist(q, t_j) < sqrt(w) - delta then j is still the unique nearest point, and
//           d2 = fp32 dist2(q, t_j) is exactly what the search would return (the radius rule is re-applied).
//   w < 0 : no correspondence; -w is a lower bound of the squared distance from q_ref to EVERY target point, so if
//           sqrt(-w) - delta > max_dist there is still none.
// Near-ties (second nearest almost as close as the nearest: a fraction ~1e-4 of the points, and persistently so once
// the cloud has stopped moving — some even flip back and forth under the last-bit jitter of the converged transform)
// would fail that test in every pass, and one such point stalls its whole CTA in a full search.  They take a second
// tier: cert2[i] = (t_j2, j2) holds the second nearest point and st2[i].w = w3 a lower bound for every point other than
// the nearest two; the winner of the exact (d2, index) key comparison between j and j2 is kept if it beats w3 as above.
// Passes 0..1 use the cheaper pruned search while the cloud still moves (no certificate is read before pass 3).
// All bound comparisons carry a 2e-5 relative slack on both sides, orders of magnitude above the fp32 rounding of
// the distances involved, so a certified answer is always the exact answer (icp_cert_ok: the comparison in squared form).
// a + b < s with a 2e-5 relative slack on both sides, from the SQUARES a2, b2, s2 (all >= 0), without a square root:
//   a + b < s  <=>  R := s^2 - a^2 - b^2 > 0  and  4 a^2 b^2 < R^2.
// With A = 1.00004 a2, B = 1.00004 b2, S = 0.99996 s2 the computed R = (S - A) - B is off by at most 3e-7 S, while the
// un-slacked s^2 - a^2 - b^2 exceeds the exact S - A - B by 4e-5 (S + A + B): a passing test implies the exact
// inequality.  S > 1e-16 keeps R^2 and 4AB clear of the denormal range (distances of 1e-8: never met; such a point takes
// the search).  The test only decides between reuse and search — both give the same answer — so it may be as strict as
// it likes; it replaces three correctly rounded sqrtf (MUFU.RSQ + Newton step + slow-path branch each) per point and pass.
__device__ __forceinline__ bool icp_cert_ok(float a2, float b2, float s2) {
    const float A = __fmul_rn(a2, 1.00004f), B = __fmul_rn(b2, 1.00004f), S = __fmul_rn(s2, 0.99996f);
    const float R = __fsub_rn(__fsub_rn(S, A), B);
    return R > 0.0f && S > 1e-16f && __fmul_rn(__fmul_rn(4.0f, A), B) < __fmul_rn(R, R);
}

__device__ __forceinline__ void icp_point_nn(const Grid &g, const float4 *__restrict__ tgt, const float4 *__restrict__ nrm, float3 q,
                                             float r2, float max_dist2_f, int cert_pass, int pass, int i, IcpState *__restrict__ trc,
                                             float4 *__restrict__ st0, float4 *__restrict__ st1, float4 *__restrict__ st2,
                                             float4 *__restrict__ cert2, IcpMatch &o) {
    if (pass > cert_pass) {
        const float4 c = st0[i], a = st1[i], b = st2[i];
        // second tier: ONE coalesced 16-byte record (the second candidate's position and index; its bound w3 rides in
        // st2.w) read on demand — it used to be (j2, w3) + a gather of tgt[j2], two chained misses for the 0.5 % of the
        // points that take this tier in every steady pass
        const float mx = q.x - c.x, my = q.y - c.y, mz = q.z - c.z;
        const float m2 = __fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz));  // |q - q_ref|^2
        if (c.w > 0.0f) {
            const int j_old = __float_as_int(a.w);
            const float d1 = dist2f(q.x, q.y, q.z, a.x, a.y, a.z);
            o.tx = a.x; o.ty = a.y; o.tz = a.z; o.nx = b.x; o.ny = b.y; o.nz = b.z;
            // (a certified nearest point that has drifted out of the radius takes the search path, which then
            // records a "no correspondence" certificate)
            if (d1 < r2 && icp_cert_ok(d1, m2, c.w)) {
                o.j = j_old;
                o.d2 = d1;
                return;
            }
            const float4 t2 = cert2[i];
            if (trc && pass < 64) atomicAdd((unsigned long long *)&trc->dbgtier2[pass], 1ull);
            const int j2 = __float_as_int(t2.w);
            if (j2 >= 0) {
                typedef unsigned long long u64k;
                const float w3 = b.w;
                const float d2b = dist2f(q.x, q.y, q.z, t2.x, t2.y, t2.z);
                const u64k ka = (((u64k)__float_as_uint(d1)) << 32) | (uint32_t)j_old;
                const u64k kb = (((u64k)__float_as_uint(d2b)) << 32) | (uint32_t)j2;
                const bool a_wins = ka < kb;  // exactly the comparison the search makes between these two
                const float dw = a_wins ? d1 : d2b;
                if (dw < r2 && icp_cert_ok(dw, m2, w3)) {
                    if (!a_wins) {
                        // the pair swaps roles: w no longer bounds "all points but the nearest", so it is set to a
                        // value that always defers to this tier; w3 bounds every point outside the pair as before
                        const float4 n2 = __ldg(nrm + j2);
                        st0[i] = make_float4(c.x, c.y, c.z, 1.17549435e-38f);
                        st1[i] = make_float4(t2.x, t2.y, t2.z, __int_as_float(j2));
                        st2[i] = make_float4(n2.x, n2.y, n2.z, w3);
                        cert2[i] = make_float4(a.x, a.y, a.z, __int_as_float(j_old));
                        o.tx = t2.x; o.ty = t2.y; o.tz = t2.z; o.nx = n2.x; o.ny = n2.y; o.nz = n2.z;
                    }
                    o.j = a_wins ? j_old : j2;
                    o.d2 = dw;
                    return;
                }
            }
        } else if (c.w < 0.0f) {
            if (icp_cert_ok(max_dist2_f, m2, -c.w)) {
                o.j = -1;
                o.d2 = 0.0f;
                return;
            }
        }
    }
    float d2, other = 0.0f, third = 0.0f;
    int j, j2 = -1;
    if (trc && pass < 64) atomicAdd((unsigned long long *)&trc->dbgsearch[pass], 1ull);
    if (pass >= cert_pass) j = grid_nn1_cert(g, q.x, q.y, q.z, r2, &d2, &other, &j2, &third);
    else j = grid_nn1(g, q.x, q.y, q.z, r2, &d2);
    o.j = j;
    o.d2 = d2;
    if (j >= 0) {
        const float4 tp = __ldg(tgt + j), np = __ldg(nrm + j);
        o.tx = tp.x; o.ty = tp.y; o.tz = tp.z; o.nx = np.x; o.ny = np.y; o.nz = np.z;
        st1[i] = make_float4(tp.x, tp.y, tp.z, __int_as_float(j));
        if (pass >= cert_pass) {
            float4 t2 = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
            if (j2 >= 0) {
                t2 = __ldg(tgt + j2);
                t2.w = __int_as_float(j2);
            }
            st0[i] = make_float4(q.x, q.y, q.z, other);
            st2[i] = make_float4(np.x, np.y, np.z, third);
            cert2[i] = t2;
        }
    } else {
        st1[i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
        // nothing inside the radius: every target point is at least min(nearest examined, h) away
        if (pass >= cert_pass) st0[i] = make_float4(q.x, q.y, q.z, -fminf(d2, other));
    }
}

// Accumulate phase of one warp and chunk: the sums of entry group GRP (7 / 6 / 7 / 7 of the 27 products) over 4 of the
// 128 staged rows of the warp's group (see k_icp_persist).
template <int GRP>
__device__ __forceinline__ void icp_accumulate(const int (*__restrict__ rw)[ICP_THREADS], int row0, long long (&acc)[7]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int r = row0 + 32 * k;
        const long long v3 = rw[3][r], v4 = rw[4][r], v5 = rw[5][r], v6 = rw[6][r];
        if (GRP == 0) {
            const long long v0 = rw[0][r], v1 = rw[1][r], v2 = rw[2][r];
            acc[0] += v0 * v0; acc[1] += v0 * v1; acc[2] += v0 * v2; acc[3] += v0 * v3;
            acc[4] += v0 * v4; acc[5] += v0 * v5; acc[6] += v0 * v6;
        } else if (GRP == 1) {
            const long long v1 = rw[1][r], v2 = rw[2][r];
            acc[0] += v1 * v1; acc[1] += v1 * v2; acc[2] += v1 * v3; acc[3] += v1 * v4;
            acc[4] += v1 * v5; acc[5] += v1 * v6;
        } else if (GRP == 2) {
            const long long v2 = rw[2][r];
            acc[0] += v2 * v2; acc[1] += v2 * v3; acc[2] += v2 * v4; acc[3] += v2 * v5;
            acc[4] += v2 * v6; acc[5] += v5 * v5; acc[6] += v5 * v6;
        } else {
            acc[0] += v3 * v3; acc[1] += v3 * v4; acc[2] += v3 * v5; acc[3] += v3 * v6;
            acc[4] += v4 * v4; acc[5] += v4 * v5; acc[6] += v4 * v6;
        }
    }
}

// barrier of one group of 4 warps (named barriers 1..6, immediate ids so that ptxas reserves only those)
__device__ __forceinline__ void icp_group_barrier(int gid) {
    switch (gid) {
        case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
        case 3: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
        case 4: asm volatile("bar.sync 5, 128;" ::: "memory"); break;
        default: asm volatile("bar.sync 6, 128;" ::: "memory"); break;
    }
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid-wide barrier of a cooperative launch (all CTAs co-resident): monotonically increasing arrival counter
__device__ __forceinline__ void icp_grid_barrier(unsigned int *bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // release-arrive / acquire-poll (measured 20 % faster than fence + atomic + fence, tools/micro/barrier_bench.cu)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        while (ld_acquire_u32(bar) < target) {
        }
    }
    __syncthreads();
}

// The WHOLE ICP loop in one persistent cooperative kernel (one launch per pcr_icp call, no host round trips).
//
// Per pass and source point (Morton order; a thread owns the same points in every pass, so the per-point state needs
// no inter-CTA coherence): transform by the cumulative fp64 transform (D7), nearest target point (certified reuse of
// the previous correspondence where provable, else the grid search — identical results, see icp_point_nn), then the
// point-to-plane row (J, r) in fp64.
//
// Normal equations (rule D5): J (6 entries) and r are quantised ONCE per correspondence to kq-bit integers
// (kq = min(30, (62 - ceil(log2 n)) / 2), scales 2^(kq - e_J), 2^(kq - e_R)); the 21 + 6 sums are sums of exact
// int32 x int32 -> int64 products, so they are order-free and bit-reproducible, and one product costs ONE
// IMAD.WIDE instead of DMUL + DMUL + F2I + 64-bit add.  The quantised rows of a 256-point chunk are staged in
// shared memory (SoA, double-buffered, one group barrier per chunk) and the 27 products are split over the warps:
// warp w forms the sums of entry group w & 3 (7 / 6 / 7 / 7 entries) for 4 rows per lane, so a thread carries
// 7 accumulators (14 registers) instead of 27 (54) — the NN search keeps its occupancy.  The inlier count and
// the fixed-point sum of d2 stay with the thread that owns the point.
//
// End of pass: one 64-bit atomic per CTA and sum into a triple-buffered accumulator, ONE grid barrier, then every
// CTA solves the 6x6 system redundantly (bit-identical), so no second barrier / broadcast is needed.
__global__ void __launch_bounds__(ICP_THREADS, ICP_CTAS_PER_SM) k_icp_persist(const float4 *__restrict__ src, int ns, Grid g,
                                                                const float4 *__restrict__ tgt, const float4 *__restrict__ nrm,
                                                                float r2, IcpState *__restrict__ S, float4 *__restrict__ st0,
                                                                float4 *__restrict__ st1, float4 *__restrict__ st2,
                                                                float4 *__restrict__ cert2, int *__restrict__ corr) {
    __shared__ int rows[2][7][ICP_THREADS];
    __shared__ long long red[ICP_WARPS][9];
    __shared__ long long tot[29];
    __shared__ double fA[6][6], fb[6];
    __shared__ IcpLocal L;
    __shared__ IcpEnd E;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 16) L.T[threadIdx.x] = S->T[threadIdx.x];
    if (threadIdx.x == 0) {
        L.prev_fit = L.prev_rmse = L.fitness = L.rmse = 0.0;
        L.count = L.sumq = 0;
        L.pass = L.done = L.iterations = L.converged = 0;
    }
    const double scd = S->sc_d;
    const float scJf = (float)S->sc_J, scRf = (float)S->sc_R;  // powers of two: exact in fp32
    const double isc_JJ = S->isc_JJ, isc_Jr = S->isc_Jr, isc_d = S->isc_d, rel_fit = S->rel_fit, rel_rmse = S->rel_rmse;
    const int max_iter = S->max_iter, trace = S->trace, cert_pass = S->cert_pass;
    const float max_dist2_f = r2 * 1.000002f;  // >= max_dist^2 (r2 is the fp32 rounding of max_dist^2)
    const int grp = warp & 3;                       // entry group of this warp
    const int row0 = (warp >> 2) * 128 + lane;      // rows row0 + 32 k, k = 0..3 (each group of 4 warps owns 128 rows)
    __syncthreads();

    for (int pass = 0;; pass++) {
        const double *T = L.T;  // broadcast shared-memory reads
        long long acc[7];
#pragma unroll
        for (int e = 0; e < 7; e++) acc[e] = 0;
        long long sumq = 0;
        int cnt = 0, buf = 0;
        const long long c0 = clock64();
        for (int base = blockIdx.x * ICP_THREADS; base < ns; base += gridDim.x * ICP_THREADS, buf ^= 1) {
            const int i = base + threadIdx.x;
            int q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0, q5 = 0, q6 = 0;
            if (i < ns) {
                const float4 p = __ldg(src + i);
                const float3 q = xform_pt(T, p.x, p.y, p.z);
                IcpMatch m;
                icp_point_nn(g, tgt, nrm, q, r2, max_dist2_f, cert_pass, pass, i, trace ? S : nullptr, st0, st1, st2, cert2, m);
                if (m.j >= 0) {
                    // point-to-plane row in fp32, every operation individually rounded (as the oracle): J and r are
                    // quantised to kq <= 30 bits right here, so fp64 bought nothing but fp64-pipe time
                    const float sx = q.x, sy = q.y, sz = q.z;
                    const float nx = m.nx, ny = m.ny, nz = m.nz;
                    const float ex = __fsub_rn(sx, m.tx), ey = __fsub_rn(sy, m.ty), ez = __fsub_rn(sz, m.tz);
                    const float r = __fadd_rn(__fadd_rn(__fmul_rn(ex, nx), __fmul_rn(ey, ny)), __fmul_rn(ez, nz));
                    q0 = __float2int_rn(__fmul_rn(__fsub_rn(__fmul_rn(sy, nz), __fmul_rn(sz, ny)), scJf));
                    q1 = __float2int_rn(__fmul_rn(__fsub_rn(__fmul_rn(sz, nx), __fmul_rn(sx, nz)), scJf));
                    q2 = __float2int_rn(__fmul_rn(__fsub_rn(__fmul_rn(sx, ny), __fmul_rn(sy, nx)), scJf));
                    q3 = __float2int_rn(__fmul_rn(nx, scJf));
                    q4 = __float2int_rn(__fmul_rn(ny, scJf));
                    q5 = __float2int_rn(__fmul_rn(nz, scJf));
                    q6 = __float2int_rn(__fmul_rn(r, scRf));
                    cnt++;
                    sumq += fixed_ll((double)m.d2, scd);
                }
            }
            int(*rw)[ICP_THREADS] = rows[buf];
            rw[0][threadIdx.x] = q0; rw[1][threadIdx.x] = q1; rw[2][threadIdx.x] = q2; rw[3][threadIdx.x] = q3;
            rw[4][threadIdx.x] = q4; rw[5][threadIdx.x] = q5; rw[6][threadIdx.x] = q6;
            // the accumulate phase of a warp reads only the 128 rows of its own group of 4 warps: a named barrier per
            // group instead of __syncthreads halves the set of warps a fast warp waits for (the search time varies).
            // (Handing the 128-point units out dynamically, per group, by an atomic ticket was measured too: -1.5 % at
            // 1M points, +10 % at 100k, where every group has a single unit anyway — not kept.)
            icp_group_barrier(warp >> 2);
            // (the entry group is uniform over the warp: one switch per chunk, not one per row)
            switch (grp) {
                case 0: icp_accumulate<0>(rw, row0, acc); break;
                case 1: icp_accumulate<1>(rw, row0, acc); break;
                case 2: icp_accumulate<2>(rw, row0, acc); break;
                default: icp_accumulate<3>(rw, row0, acc); break;
            }
        }
        // CTA reduction, then one atomic per sum and CTA into the pass's accumulator
        long long *gacc = S->acc3[pass % 3];
        const long long c1 = clock64();
        // Warp reduction of the eight 64-bit sums as a transposing butterfly: at offsets 16, 8, 4 a lane keeps half of its
        // values and hands the other half to its partner (4 + 2 + 1 shuffles), then offsets 2 and 1 finish the one value
        // left — 9 64-bit shuffles instead of 8 x 5, after which lane 4 m holds the warp total of value m.  The plain
        // per-value reduction was 17 % of the kernel's instructions at 100k points, all warps of the SM hitting the
        // 32-lane/clock shuffle pipe at the same moment at the end of every pass (profiles/r2_summary.md).
        long long tv;
        {
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            long long w4[4], w2[2];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long long lo = acc[k], hi = k + 4 < 7 ? acc[k + 4] : sumq;
                w4[k] = (h16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, h16 ? lo : hi, 16);
            }
#pragma unroll
            for (int k = 0; k < 2; k++) w2[k] = (h8 ? w4[k + 2] : w4[k]) + __shfl_xor_sync(0xffffffffu, h8 ? w4[k] : w4[k + 2], 8);
            tv = (h4 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? w2[0] : w2[1], 4);
            tv += __shfl_xor_sync(0xffffffffu, tv, 2);
            tv += __shfl_xor_sync(0xffffffffu, tv, 1);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        __syncthreads();  // the accumulate phase of the last chunk is done in every warp before `red` is reused
        // value m = 4 [lane & 16] + 2 [lane & 8] + [lane & 4]: acc[0..6] in slots 0..6, the sum of d2 (value 7) in slot 8
        if ((lane & 3) == 0) {
            const int m = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            red[warp][m == 7 ? 8 : m] = tv;
        }
        if (lane == 1) red[warp][7] = cnt;
        __syncthreads();
        if (threadIdx.x < 29) {
            // sum e lives in slot `sl` of the two warps of group `gr` (see the accumulate phase above)
            const int e = threadIdx.x;
            int gr, sl;
            if (e < 6) { gr = 0; sl = e; }                       // (0,0..5)
            else if (e < 11) { gr = 1; sl = e - 6; }             // (1,1..5)
            else if (e < 15) { gr = 2; sl = e - 11; }            // (2,2..5)
            else if (e < 18) { gr = 3; sl = e - 15; }            // (3,3..5)
            else if (e < 20) { gr = 3; sl = e - 18 + 4; }        // (4,4), (4,5)
            else if (e == 20) { gr = 2; sl = 5; }                // (5,5)
            else if (e == 21) { gr = 0; sl = 6; }                // J0 r
            else if (e == 22) { gr = 1; sl = 5; }                // J1 r
            else if (e == 23) { gr = 2; sl = 4; }                // J2 r
            else if (e == 24) { gr = 3; sl = 3; }                // J3 r
            else if (e == 25) { gr = 3; sl = 6; }                // J4 r
            else if (e == 26) { gr = 2; sl = 6; }                // J5 r
            else { gr = -1; sl = e - 20; }                       // 27: count (slot 7), 28: sum d2 (slot 8), all warps
            long long sum = 0;
            if (gr >= 0) {
#pragma unroll
                for (int w = gr; w < ICP_WARPS; w += 4) sum += red[w][sl];
            } else {
#pragma unroll
                for (int w = 0; w < ICP_WARPS; w++) sum += red[w][sl];
            }
            if (sum != 0) atomicAdd((unsigned long long *)&gacc[e * ICP_ACC_STRIDE], (unsigned long long)sum);
        }
        const long long cb = clock64();
        if (trace && threadIdx.x == 0 && pass < 64) atomicMax((unsigned long long *)&S->dbgmax[pass], (unsigned long long)(c1 - c0));
        if (trace && threadIdx.x == 0 && pass == 20 && blockIdx.x < 1024) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            S->dbgcta[blockIdx.x] = ((unsigned long long)(c1 - c0) << 16) | smid;
        }
        icp_grid_barrier(&S->bar, (unsigned int)(pass + 1) * gridDim.x);
        const long long ce = clock64();
        if (threadIdx.x < 29) {
            const long long v = __ldcg(&gacc[threadIdx.x * ICP_ACC_STRIDE]);
            tot[threadIdx.x] = v;
            const int e = threadIdx.x;
            if (e < 21) {
                int a = 0, rem = e;
                while (rem >= 6 - a) { rem -= 6 - a; a++; }
                const double d = (double)v * isc_JJ;
                fA[a][a + rem] = d;
                fA[a + rem][a] = d;
            } else if (e < 27) {
                fb[e - 21] = -((double)v * isc_Jr);
            }
        }
        // the buffer of pass - 1 was read by every CTA before it arrived at this barrier; it is next used in pass + 2
        if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 32 + 29) S->acc3[(pass + 2) % 3][(threadIdx.x - 32) * ICP_ACC_STRIDE] = 0;
        __syncthreads();
        const long long c2 = clock64();
        if (threadIdx.x == 0) icp_end_solve(&E, tot, fA, fb);
        else if (threadIdx.x == 32) icp_end_stats(&L, &E, tot, isc_d, rel_fit, rel_rmse, max_iter, ns);
        __syncthreads();
        const long long c3 = clock64();
        if (!E.stop && E.solved && lane == 0 && warp < 3) pcr_sincos(E.x[warp], &E.sn[warp], &E.cs[warp]);
        __syncthreads();
        const long long c4 = clock64();
        if (warp == 0) {
            icp_end_compose(&L, &E, lane);
            if (lane == 0) { L.tdbg[0] = c3 - c2; L.tdbg[1] = c4 - c3; L.tdbg[2] = clock64() - c4; }
        }
        __syncthreads();
        if (trace && threadIdx.x == 0 && pass < 64) atomicMax((unsigned long long *)&S->dbgfin[pass], (unsigned long long)(clock64() - c2));
        if (trace && blockIdx.x == 0 && threadIdx.x == 0) {
            S->dbg[0] += c1 - c0;
            if (pass < 63) S->dbgp[pass] = c1 - c0;
            S->dbg[1] += c2 - c1;
            S->dbg[3] += ce - cb;
            S->dbg[2] += clock64() - c2;
            for (int i = 0; i < 3; i++) S->dbg2[i] += L.tdbg[i];
        }
        if (L.done) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < 16; i++) S->T_out[i] = L.T[i];
        S->fitness = L.fitness;
        S->rmse = L.rmse;
        S->count = L.count;
        S->sumq = L.sumq;
        S->pass = L.pass;
        S->iterations = L.iterations;
        S->converged = L.converged;
        S->done = 1;
    }
    // Morton order back to the caller's source order (a thread reads only the seeds it wrote itself)
    if (corr)
        for (int base = blockIdx.x * ICP_THREADS; base < ns; base += gridDim.x * ICP_THREADS) {
            const int i = base + threadIdx.x;
            if (i < ns) corr[__float_as_int(__ldg(src + i).w)] = __float_as_int(st1[i].w);
        }
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

// target grid for radius max_dist, Morton-sorted source, largest |target coordinate| (for the fixed-point scale)
int pcr_icp_prepare(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, int nt, double max_dist, IcpPrep *prep) {
    prep->valid = false;
    if (!(max_dist > 0.0) || ns <= 0 || nt <= 0) return PCR_OK;  // pcr_icp_impl reports / handles these cases
    float lo[3], hi[3];
    PCR_TRY(pcr_bounds(ctx, tgt, nt, lo, hi));
    prep->amax = 0.0f;
    for (int d = 0; d < 3; d++) prep->amax = fmaxf(prep->amax, fmaxf(fabsf(lo[d]), fabsf(hi[d])));
    PCR_TRY(pcr_grid_build_compact(ctx, tgt, nt, max_dist, lo, hi, &prep->g));
    PCR_TRY(pcr_morton_sort(ctx, src, ns, &prep->src_sorted));
    prep->valid = true;
    return PCR_OK;
}

int pcr_icp_impl(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, const float4 *nrm, int nt,
                 double max_dist, const double *init, int max_iter, double rel_fit, double rel_rmse,
                 pcr_reg_result *res, int *corr, bool sync_result, const IcpPrep *prepared) {
    memset(res, 0, sizeof(*res));
    for (int i = 0; i < 16; i++) res->transformation[i] = init[i];
    res->best_hyp = -1;
    if (!(max_dist > 0.0)) return pcr_fail(ctx, PCR_ERR_INVALID, "max_correspondence_distance must be > 0");
    if (max_iter < 0) return pcr_fail(ctx, PCR_ERR_INVALID, "max_iteration must be >= 0");
    if (ns == 0 || nt == 0) {
        if (corr && ns > 0) PCR_CUDA(cudaMemsetAsync(corr, 0xff, sizeof(int) * (size_t)ns, ctx->stream));
        return PCR_OK;
    }
    IcpPrep own;
    if (!(prepared && prepared->valid)) {
        PCR_TRY(pcr_icp_prepare(ctx, src, ns, tgt, nt, max_dist, &own));
        prepared = &own;
    }
    const float amax = prepared->amax;
    Grid g = prepared->g;
    const float4 *src_sorted = prepared->src_sorted;

    const int lg = pcr_ilog2ceil(ns > 1 ? ns : 1);
    const int e_r = pcr_pow2ceil_exp(max_dist);
    const int e_J = pcr_pow2ceil_exp(2.0 * ((double)amax + max_dist) + 1.0);
    const int e_R = e_r + 2;
    const int k_d = 62 - 2 * e_r - lg;
    const int kq = (62 - lg) / 2 < 30 ? (62 - lg) / 2 : 30;  // bits per quantised factor
    const int s_J = kq - e_J, s_R = kq - e_R;

    IcpState *hS = (IcpState *)ctx->pinned;
    memset(hS, 0, sizeof(IcpState));
    for (int i = 0; i < 16; i++) hS->T[i] = init[i];
    hS->max_iter = max_iter;
    hS->trace = getenv("PCR_ICP_TRACE") ? 1 : 0;
    hS->cert_pass = getenv("PCR_ICP_CERT_PASS") ? atoi(getenv("PCR_ICP_CERT_PASS")) : ICP_CERT_PASS;
    if (hS->cert_pass < 0) hS->cert_pass = 0;
    hS->rel_fit = rel_fit;
    hS->rel_rmse = rel_rmse;
    hS->sc_J = ldexp(1.0, s_J); hS->isc_JJ = ldexp(1.0, -2 * s_J);
    hS->sc_R = ldexp(1.0, s_R); hS->isc_Jr = ldexp(1.0, -(s_J + s_R));
    hS->sc_d = ldexp(1.0, k_d);   hS->isc_d = ldexp(1.0, -k_d);
    PCR_ALLOC(dS, IcpState, 1);
    PCR_ALLOC(st0, float4, (size_t)ns);
    PCR_ALLOC(st1, float4, (size_t)ns);
    PCR_ALLOC(st2, float4, (size_t)ns);
    PCR_ALLOC(cert2, float4, (size_t)ns);
    const size_t state_bytes = hS->trace ? sizeof(IcpState) : offsetof(IcpState, acc3);                 // device -> host
    const size_t init_bytes = hS->trace ? sizeof(IcpState) : offsetof(IcpState, dbgsearch);             // host -> device: + zeroed sums
    PCR_CUDA(cudaMemcpyAsync(dS, hS, init_bytes, cudaMemcpyHostToDevice, ctx->stream));
    float r2 = (float)(max_dist * max_dist);
    // cooperative launch: every CTA must be resident (the kernel synchronises the grid once per pass)
    int &occ = ctx->occ_icp;
    if (!occ) PCR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_icp_persist, ICP_THREADS, 0));
    if (occ < 1) return pcr_fail(ctx, PCR_ERR_CUDA, "k_icp_persist does not fit on an SM");
    int blocks = min(div_up(ns, ICP_THREADS), ctx->sm_count * occ);
    // diagnostic: a small grid gives a small cloud the many-chunks-per-CTA shape of a large one (tests/test_gpu_parity.py)
    if (getenv("PCR_ICP_MAX_CTAS")) blocks = max(1, min(blocks, atoi(getenv("PCR_ICP_MAX_CTAS"))));
    const size_t pend_idx = ctx->pending.size();
    {
        // ONE launch runs all passes; the per-pass figure (52 B per point) is what the kernel statistics report,
        // with the number of passes actually run filled in below
        KScope ks(ctx, KC_ICP_PASS, 16.0 * ns + 32.0 * nt + 4.0 * ns, max_iter + 1);
        void *args[] = {(void *)&src_sorted, (void *)&ns, (void *)&g, (void *)&tgt, (void *)&nrm, (void *)&r2, (void *)&dS,
                        (void *)&st0, (void *)&st1, (void *)&st2, (void *)&cert2, (void *)&corr};
        PCR_CUDA(cudaLaunchCooperativeKernel((const void *)k_icp_persist, dim3(blocks), dim3(ICP_THREADS), args, 0, ctx->stream));
        PCR_LAUNCHED();
    }
    PCR_CUDA(cudaGetLastError());
    PCR_CUDA(cudaMemcpyAsync(hS, dS, state_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (sync_result) {
        PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
        for (int i = 0; i < 16; i++) res->transformation[i] = hS->T_out[i];
        res->fitness = hS->fitness;
        res->inlier_rmse = hS->rmse;
        res->inlier_count = hS->count;
        res->sum_d2_fixed = hS->sumq;
        res->k_d = k_d;
        res->iterations = hS->iterations;
        res->converged = hS->converged;
        if (getenv("PCR_ICP_TRACE"))
            fprintf(stderr, "[pcr icp] ns %d passes %d blocks %d cycles/pass: loop %.0f  reduce+barrier %.0f  finish %.0f\n", ns, hS->pass,
                    blocks, (double)hS->dbg[0] / hS->pass, (double)hS->dbg[1] / hS->pass, (double)hS->dbg[2] / hS->pass);
        if (getenv("PCR_ICP_TRACE")) {
            fprintf(stderr, "[pcr icp] loop cycles per pass:");
            for (int i = 0; i < hS->pass && i < 12; i++) fprintf(stderr, " %lld", hS->dbgp[i]);
            fprintf(stderr, "\n[pcr icp] slowest CTA loop per pass:");
            for (int i = 0; i < hS->pass && i < 24; i++) fprintf(stderr, " %lld", hS->dbgmax[i]);
            fprintf(stderr, "\n[pcr icp] points on the search path per pass:");
            for (int i = 0; i < hS->pass && i < 24; i++) fprintf(stderr, " %lld", hS->dbgsearch[i]);
            fprintf(stderr, "\n[pcr icp] points on the second tier per pass:");
            for (int i = 0; i < hS->pass && i < 24; i++) fprintf(stderr, " %lld", hS->dbgtier2[i]);
            fprintf(stderr, "\n[pcr icp] slowest CTA finish per pass:");
            for (int i = 0; i < hS->pass && i < 24; i++) fprintf(stderr, " %lld", hS->dbgfin[i]);
            fprintf(stderr, "  barrier alone %.0f\n", (double)hS->dbg[3] / hS->pass);
            if (hS->pass > 20) {
                fprintf(stderr, "[pcr icp] pass 20, per CTA (index:sm:loop cycles):");
                for (int i = 0; i < blocks && i < 1024; i++) fprintf(stderr, " %d:%llu:%llu", i, hS->dbgcta[i] & 0xffffull, hS->dbgcta[i] >> 16);
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "[pcr icp] end-of-pass logic of CTA 0, cycles per pass: solve || statistics %.0f  sin/cos %.0f  compose %.0f\n",
                    (double)hS->dbg2[0] / hS->pass, (double)hS->dbg2[1] / hS->pass, (double)hS->dbg2[2] / hS->pass);
        }
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
    PCR_TRY(pcr_grid_build_compact(ctx, tgt, nt, radius, nullptr, nullptr, &g));
    KScope ks(ctx, KC_NN1, 16.0 * nt + 24.0 * nq);
    k_nn1<<<div_up(nq, 256), 256, 0, ctx->stream>>>(q, nq, g, (float)(radius * radius), idx, d2);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
