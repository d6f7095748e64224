// pcr_capi.cu — extern "C" entry points declared in include/pcr.h and the end-to-end align driver.
#include <cstdlib>

#include "pcr_common.cuh"
#include <thread>

typedef unsigned long long u64;

// implementation functions (one per translation unit)
int pcr_icp_impl(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, const float4 *nrm, int nt,
                 double max_dist, const double *init, int max_iter, double rel_fit, double rel_rmse,
                 pcr_reg_result *res, int *corr, bool sync_result, const IcpPrep *prepared = nullptr);
int pcr_icp_prepare(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, int nt, double max_dist, IcpPrep *prep);
int pcr_nn1_impl(pcr_ctx *ctx, const float4 *tgt, int nt, const float4 *q, int nq, double radius, int *idx, float *d2);
int pcr_knn_impl(pcr_ctx *ctx, const float4 *pts, int n, const float4 *q, int nq, double radius, int max_nn, int *idx,
                 float *d2, int *cnt);
int pcr_normals_impl(pcr_ctx *ctx, const float4 *pts, int n, double radius, int max_nn, float4 *normals);
int pcr_fpfh_impl(pcr_ctx *ctx, const float4 *pts, const float4 *nrm, int n, double radius, int max_nn, float *out);
int pcr_voxel_impl(pcr_ctx *ctx, const float4 *pts, int n, double voxel, float4 *out, int *m_host);
int pcr_voxel_enqueue(pcr_ctx *ctx, const float4 *pts, int n, double voxel, float4 *out, unsigned long long *h_total);
int pcr_nn_features_impl(pcr_ctx *ctx, const float *fq, int nq, const float *fb, int nb, int *nn);
int pcr_match_impl(pcr_ctx *ctx, const float *fs, int ms, const float *ft, int mt, int mutual, double mutual_ratio,
                   int *corr, int *c_host);
int pcr_ransac_impl(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, const int *corr, int c,
                    double max_dist, double edge_sim, int64_t max_iter, double confidence, u64 seed,
                    pcr_reg_result *res, const RansacWork *prepared = nullptr);
int pcr_ransac_prepare(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, double max_dist,
                       RansacWork *w);
int pcr_ransac_session_begin_impl(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, double max_dist);
int pcr_ransac_session_end_impl(pcr_ctx *ctx);
int pcr_ransac_wave_impl(pcr_ctx *ctx, const RansacWork &w, const float4 *src, int ms, const float4 *tgt,
                         const int *corr, int c, double max_dist, double edge_sim, long long hyp_begin,
                         long long hyp_end, u64 seed, long long best_cnt, long long best_sumq, pcr_hyp_record *recs_host,
                         int cap, int *n_recs_host, long long *n_surv_host);
int pcr_ransac_step_impl(pcr_ctx *ctx, const float4 *src, const float4 *tgt, const int *corr, int c, u64 seed,
                         long long h_begin, int count, double *T);
int pcr_inlier_count_impl(pcr_ctx *ctx, const float4 *src, const float4 *tgt, const int *corr, int c, const double *T,
                          int count, double thresh, int squared, int *counts);

int pcr_corr_check_impl(pcr_ctx *ctx, const int *corr, int c, int ms, int mt);
int pcr_pack_impl(pcr_ctx *ctx, const float *xyz, int n, float4 *out);

struct CallGuard {
    pcr_ctx *ctx;
    bool ok;
    explicit CallGuard(pcr_ctx *c) : ctx(c), ok(false) {
        if (!ctx) return;
        bool expected = false;
        if (!ctx->busy.compare_exchange_strong(expected, true, std::memory_order_acquire)) return;
        ok = true;
        cudaSetDevice(ctx->device);
        pcr_arena_reset(ctx);
    }
    ~CallGuard() {
        if (ok) ctx->busy.store(false, std::memory_order_release);
    }
};
#define PCR_ENTER()                   \
    if (!ctx) return PCR_ERR_INVALID; \
    CallGuard guard__(ctx);           \
    if (!guard__.ok) return PCR_ERR_BUSY /* the message buffer belongs to the call in progress: not touched */
#define PCR_ARG(cond) \
    if (!(cond)) return pcr_fail(ctx, PCR_ERR_INVALID, "invalid argument: %s", #cond)

extern "C" {

int pcr_voxel_downsample(pcr_ctx *ctx, const float *xyzw, int n, double voxel, float *out, int *m_host) {
    PCR_ENTER();
    PCR_ARG(n >= 0 && m_host);
    return pcr_voxel_impl(ctx, (const float4 *)xyzw, n, voxel, (float4 *)out, m_host);
}

int pcr_estimate_normals(pcr_ctx *ctx, const float *xyzw, int n, double radius, int max_nn, float *normals) {
    PCR_ENTER();
    PCR_ARG(n >= 0);
    return pcr_normals_impl(ctx, (const float4 *)xyzw, n, radius, max_nn, (float4 *)normals);
}

int pcr_compute_fpfh(pcr_ctx *ctx, const float *xyzw, const float *normals, int n, double radius, int max_nn,
                     float *fpfh) {
    PCR_ENTER();
    PCR_ARG(n >= 0);
    return pcr_fpfh_impl(ctx, (const float4 *)xyzw, (const float4 *)normals, n, radius, max_nn, fpfh);
}

int pcr_knn_hybrid(pcr_ctx *ctx, const float *xyzw, int n, const float *q, int nq, double radius, int max_nn, int *idx,
                   float *d2, int *cnt) {
    PCR_ENTER();
    PCR_ARG(n >= 0 && nq >= 0);
    return pcr_knn_impl(ctx, (const float4 *)xyzw, n, (const float4 *)q, nq, radius, max_nn, idx, d2, cnt);
}

int pcr_nn1(pcr_ctx *ctx, const float *tgt, int nt, const float *q, int nq, double radius, int *idx, float *d2) {
    PCR_ENTER();
    PCR_ARG(nt >= 0 && nq >= 0);
    return pcr_nn1_impl(ctx, (const float4 *)tgt, nt, (const float4 *)q, nq, radius, idx, d2);
}

int pcr_match_features(pcr_ctx *ctx, const float *fs, int ms, const float *ft, int mt, int mutual, double mutual_ratio,
                       int *corr, int *c_host) {
    PCR_ENTER();
    PCR_ARG(ms >= 0 && mt >= 0 && c_host);
    return pcr_match_impl(ctx, fs, ms, ft, mt, mutual, mutual_ratio, corr, c_host);
}

int pcr_nn_features(pcr_ctx *ctx, const float *fq, int nq, const float *fb, int nb, int *nn) {
    PCR_ENTER();
    PCR_ARG(nq >= 0 && nb >= 0);
    return pcr_nn_features_impl(ctx, fq, nq, fb, nb, nn);
}

int pcr_ransac(pcr_ctx *ctx, const float *src, int ms, const float *tgt, int mt, const int *corr, int c,
               double max_dist, double edge_sim, int64_t max_iter, double confidence, uint64_t seed,
               pcr_reg_result *result) {
    PCR_ENTER();
    PCR_ARG(ms >= 0 && mt >= 0 && c >= 0 && result);
    PCR_TRY(pcr_corr_check_impl(ctx, corr, c, ms, mt));
    return pcr_ransac_impl(ctx, (const float4 *)src, ms, (const float4 *)tgt, mt, corr, c, max_dist, edge_sim, max_iter,
                           confidence, seed, result);
}

int pcr_ransac_wave(pcr_ctx *ctx, const float *src, int ms, const float *tgt, int mt, const int *corr, int c,
                    double max_dist, double edge_sim, int64_t hyp_begin, int64_t hyp_end, uint64_t seed,
                    int64_t best_count, int64_t best_sum, pcr_hyp_record *records, int cap, int *n_records,
                    int64_t *n_survivors) {
    PCR_ENTER();
    PCR_ARG(ms > 0 && mt > 0 && c >= 3 && max_dist > 0.0 && records && cap > 0 && n_records && n_survivors);
    RansacWork w;
    auto &rs = ctx->rsess;
    if (rs.active && rs.src == (const void *)src && rs.tgt == (const void *)tgt && rs.ms == ms && rs.mt == mt && rs.max_dist == max_dist) {
        w = rs.w;  // inside a session on these clouds: the grid and the sorted source are already there
        if (rs.corr_ok != (const void *)corr || rs.corr_ok_c != c) {  // validated once per session and buffer
            PCR_TRY(pcr_corr_check_impl(ctx, corr, c, ms, mt));
            rs.corr_ok = corr;
            rs.corr_ok_c = c;
        }
    } else {
        PCR_TRY(pcr_corr_check_impl(ctx, corr, c, ms, mt));
        PCR_TRY(pcr_ransac_prepare(ctx, (const float4 *)src, ms, (const float4 *)tgt, mt, max_dist, &w));
    }
    long long ns = 0;
    const int rc = pcr_ransac_wave_impl(ctx, w, (const float4 *)src, ms, (const float4 *)tgt, corr, c, max_dist, edge_sim,
                                        hyp_begin, hyp_end, seed, best_count, best_sum, records, cap, n_records, &ns);
    *n_survivors = ns;
    return rc;
}

int pcr_ransac_session_begin(pcr_ctx *ctx, const float *src, int ms, const float *tgt, int mt, double max_dist) {
    PCR_ENTER();
    PCR_ARG(ms > 0 && mt > 0 && max_dist > 0.0);
    return pcr_ransac_session_begin_impl(ctx, (const float4 *)src, ms, (const float4 *)tgt, mt, max_dist);
}

int pcr_ransac_session_end(pcr_ctx *ctx) {
    PCR_ENTER();
    return pcr_ransac_session_end_impl(ctx);
}

int pcr_ransac_step(pcr_ctx *ctx, const float *src, int ms, const float *tgt, int mt, const int *corr, int c, uint64_t seed,
                    int64_t h_begin, int count, double *T) {
    PCR_ENTER();
    PCR_ARG(ms >= 0 && mt >= 0 && c >= 0 && count >= 0);
    PCR_TRY(pcr_corr_check_impl(ctx, corr, c, ms, mt));
    return pcr_ransac_step_impl(ctx, (const float4 *)src, (const float4 *)tgt, corr, c, seed, h_begin, count, T);
}

int pcr_inlier_count(pcr_ctx *ctx, const float *src, int ms, const float *tgt, int mt, const int *corr, int c, const double *T,
                     int count, double thresh, int squared, int *counts) {
    PCR_ENTER();
    PCR_ARG(ms >= 0 && mt >= 0 && c >= 0 && count >= 0);
    PCR_TRY(pcr_corr_check_impl(ctx, corr, c, ms, mt));
    return pcr_inlier_count_impl(ctx, (const float4 *)src, (const float4 *)tgt, corr, c, T, count, thresh, squared,
                                 counts);
}

int pcr_icp_point_to_plane(pcr_ctx *ctx, const float *src, int ns, const float *tgt, const float *nrm, int nt,
                           double max_dist, const double *init, int max_iter, double rel_fitness, double rel_rmse,
                           pcr_reg_result *result, int *corr) {
    PCR_ENTER();
    PCR_ARG(ns >= 0 && nt >= 0 && init && result);
    return pcr_icp_impl(ctx, (const float4 *)src, ns, (const float4 *)tgt, (const float4 *)nrm, nt, max_dist, init,
                        max_iter, rel_fitness, rel_rmse, result, corr, true);
}

void pcr_align_default_params(pcr_align_params *p) {
    if (!p) return;
    p->voxel_size = 0.3;            // Ply default, src/ply/ply.py:32
    p->ransac_max_iter = 30;        // src/matcher/ransac.py:24
    p->ransac_confidence = 0.999;   // src/matcher/ransac.py:58
    p->seed = 0;
    p->icp_max_iter = 30;           // Open3D ICPConvergenceCriteria default (A.7)
    p->icp_rel_fitness = 1e-6;
    p->icp_rel_rmse = 1e-6;
    p->source_normals = 1;          // Ply.__init__ estimates full-resolution normals on every cloud (ply.py:65)
    p->reserved = 0;
}

}  // extern "C"

// ---- end to end --------------------------------------------------------------------------------------------------
struct StageTimer {
    cudaEvent_t ev[9];
    int n = 0;
    cudaStream_t s;
    explicit StageTimer(cudaStream_t st) : s(st) {
        for (auto &e : ev) cudaEventCreate(&e);
    }
    ~StageTimer() {
        for (auto &e : ev) cudaEventDestroy(e);
    }
    void mark() {
        if (n < 9) cudaEventRecord(ev[n++], s);
    }
};

int pcr_helper_get(pcr_ctx *ctx, pcr_ctx **out);

// Ply._preprocess of one cloud (src/ply/ply.py:106-120) on `c`, in two parts: the voxel grid (its size comes back through
// one pinned word, *h_total >> 32, valid once c->stream has been synchronised), then normals + FPFH of the down-sampled
// cloud (fully asynchronous)
static int preprocess_voxel_enqueue(pcr_ctx *c, const float4 *pts, int n, double v, float4 **down, unsigned long long *h_total) {
    pcr_ctx *ctx = c;
    PCR_ALLOC(d, float4, (size_t)n);
    PCR_TRY(pcr_voxel_enqueue(ctx, pts, n, v, d, h_total));
    *down = d;
    return PCR_OK;
}

static int preprocess_voxel_finish(pcr_ctx *c, const float4 *pts, int n, const float4 *d, int m) {
    pcr_ctx *ctx = c;
    // the voxel centroids lie inside the bounding box of the cloud they average: that box (cached by the voxel
    // stage) serves as the box of the search grids over the down-sampled cloud — grids are only search structures,
    // no result depends on their origin — and saves a reduction + host synchronisation per cloud
    float lo[3], hi[3];
    PCR_TRY(pcr_bounds(ctx, pts, n, lo, hi));  // cache hit
    pcr_ctx::BoundsEntry e;
    e.ptr = d;
    e.n = m;
    for (int k = 0; k < 3; k++) { e.lo[k] = lo[k]; e.hi[k] = hi[k]; }
    ctx->bounds_cache.push_back(e);
    return PCR_OK;
}

static int preprocess_features(pcr_ctx *c, const float4 *d, int m, double v, float4 **nrm, float **fpfh) {
    pcr_ctx *ctx = c;
    PCR_ALLOC(nn, float4, (size_t)m);
    PCR_TRY(pcr_normals_impl(ctx, d, m, 2.0 * v, 30, nn));
    PCR_ALLOC(f, float, (size_t)m * 33);
    PCR_TRY(pcr_fpfh_impl(ctx, d, nn, m, 5.0 * v, 100, f));
    *nrm = nn;
    *fpfh = f;
    return PCR_OK;
}

// The full-resolution normals (needed only by ICP) run on the helper context (own stream + host thread) NEXT TO matching
// and RANSAC, whose validation kernels leave issue slots free:
//   preprocessing of both clouds (main)  ->  matching + RANSAC (main) || full-resolution normals (helper)  ->  ICP (main)
// Every stage computes exactly what it computes alone, so results do not depend on the overlap (PCR_ALIGN_OVERLAP=0
// runs the stages one after the other on the main context).  stage_ms are the main thread's stage times: 0 voxel +
// normals + FPFH of both clouds, 3 matching, 4 RANSAC, 5 waiting for the full-resolution normals, 6 ICP, 7 total.
static int align_device(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, int nt, const pcr_align_params *p,
                        pcr_align_result *res) {
    memset(res, 0, sizeof(*res));
    const double v = p->voxel_size;
    if (!(v > 0.0)) return pcr_fail(ctx, PCR_ERR_INVALID, "voxel_size must be > 0");
    if (ns <= 0 || nt <= 0) return pcr_fail(ctx, PCR_ERR_INVALID, "Point cloud is empty");  // src/ply/ply.py:81-84
    static const bool overlap = !(getenv("PCR_ALIGN_OVERLAP") && atoi(getenv("PCR_ALIGN_OVERLAP")) == 0);
    static const bool use_prio = !(getenv("PCR_ALIGN_PRIORITY") && atoi(getenv("PCR_ALIGN_PRIORITY")) == 0);
    pcr_ctx *h = nullptr;
    if (overlap) PCR_TRY(pcr_helper_get(ctx, &h));
    // While the helper context works, the critical path (this thread's kernels) runs on a highest-priority stream, so
    // that its CTAs are dispatched ahead of the helper's queued ones; the call ends with a full synchronisation, so
    // the caller's stream sees no difference.  The guard restores ctx->stream on every return path.
    struct StreamSwap {
        pcr_ctx *c;
        cudaStream_t saved;
        bool on;
        ~StreamSwap() {
            if (on) {
                pcr_sync_stream(c, c->stream);
                c->stream = saved;
            }
        }
    } swap{ctx, ctx->stream, false};
    if (overlap && use_prio && ctx->hp_stream) {
        cudaEvent_t in_ready;
        PCR_CUDA(cudaEventCreateWithFlags(&in_ready, cudaEventDisableTiming));
        PCR_CUDA(cudaEventRecord(in_ready, ctx->stream));
        PCR_CUDA(cudaStreamWaitEvent(ctx->hp_stream, in_ready, 0));
        PCR_CUDA(cudaEventDestroy(in_ready));
        ctx->stream = ctx->hp_stream;
        swap.on = true;
    }
    StageTimer tm(ctx->stream);
    tm.mark();
    float4 *sd = nullptr, *td = nullptr, *sn = nullptr, *tn = nullptr, *tfn = nullptr;
    float *sf = nullptr, *tf = nullptr;
    int ms = 0, mt = 0;
    cudaEvent_t ready = nullptr;
    if (overlap) {
        // the clouds were produced on the main stream: the helper stream waits for them
        PCR_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        PCR_CUDA(cudaEventRecord(ready, ctx->stream));
        PCR_CUDA(cudaStreamWaitEvent(h->stream, ready, 0));
        pcr_arena_reset(h);
        h->bounds_cache.clear();
    }
    // the boxes of the two full clouds: one launch, one synchronisation, shared with the helper context (round 1: four
    // reductions, each with its own init kernel and host synchronisation)
    PCR_TRY(pcr_bounds_pair(ctx, src, ns, tgt, nt));
    if (overlap) h->bounds_cache = ctx->bounds_cache;
    // Both voxel grids first (each ends with a host synchronisation), then normals + FPFH of the two down-sampled clouds
    // side by side on the main and the auxiliary stream: they are chains of small, latency-bound kernels over ~9k points
    // (one to two waves of CTAs each) whose tails leave most of the GPU idle.  PCR_PRE_CONCURRENT=0: one after the other.
    static const bool pre_conc = !(getenv("PCR_PRE_CONCURRENT") && atoi(getenv("PCR_PRE_CONCURRENT")) == 0);
    // The two voxel grids run side by side as well (main + auxiliary stream; one host wait for both sizes): each is a chain
    // of five small kernels (~50 us) that used to end in its own synchronisation.
    int rc;
    {
        unsigned long long *h_tot = (unsigned long long *)ctx->pinned;  // two slots
        const bool two = overlap && pre_conc && ctx->aux_stream;
        if (two) {
            cudaEvent_t ev;
            cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            cudaEventRecord(ev, ctx->stream);  // the inputs are complete on the main stream (the bounds were read back)
            cudaStreamWaitEvent(ctx->aux_stream, ev, 0);
            cudaEventDestroy(ev);
        }
        rc = preprocess_voxel_enqueue(ctx, src, ns, v, &sd, &h_tot[0]);
        if (rc == PCR_OK) {
            cudaStream_t keep = ctx->stream;
            if (two) ctx->stream = ctx->aux_stream;
            rc = preprocess_voxel_enqueue(ctx, tgt, nt, v, &td, &h_tot[1]);
            ctx->stream = keep;
            if (two && pcr_sync_stream(ctx, ctx->aux_stream) != cudaSuccess && rc == PCR_OK)
                rc = pcr_fail(ctx, PCR_ERR_CUDA, "voxel grid (auxiliary stream): %s", cudaGetErrorString(cudaGetLastError()));
        }
        if (pcr_sync_stream(ctx, ctx->stream) != cudaSuccess && rc == PCR_OK)
            rc = pcr_fail(ctx, PCR_ERR_CUDA, "voxel grid: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc == PCR_OK) {
            ms = (int)(h_tot[0] >> 32);
            mt = (int)(h_tot[1] >> 32);
            rc = preprocess_voxel_finish(ctx, src, ns, sd, ms);
        }
        if (rc == PCR_OK) rc = preprocess_voxel_finish(ctx, tgt, nt, td, mt);
    }
    if (rc == PCR_OK) {
        if (overlap && pre_conc && ctx->aux_stream) {
            cudaEvent_t ev;
            cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            cudaEventRecord(ev, ctx->stream);
            cudaStreamWaitEvent(ctx->aux_stream, ev, 0);
            cudaStream_t keep = ctx->stream;
            ctx->stream = ctx->aux_stream;
            rc = preprocess_features(ctx, td, mt, v, &tn, &tf);
            ctx->stream = keep;
            cudaEventRecord(ev, ctx->aux_stream);
            if (rc == PCR_OK) rc = preprocess_features(ctx, sd, ms, v, &sn, &sf);
            cudaStreamWaitEvent(ctx->stream, ev, 0);
            cudaEventDestroy(ev);
        } else {
            rc = preprocess_features(ctx, sd, ms, v, &sn, &sf);
            if (rc == PCR_OK) rc = preprocess_features(ctx, td, mt, v, &tn, &tf);
        }
    }
    if (rc != PCR_OK) {
        if (ready) cudaEventDestroy(ready);
        return rc;
    }
    tm.mark();
    tm.mark();
    tm.mark();
    // Ply._add_normals on the full-resolution clouds (src/ply/ply.py:65,133-135): next to matching + RANSAC
    IcpPrep icp_prep;
    pcr_ctx *const main_ctx = ctx;
    // ICP needs the target normals and its search structures, not the source normals (Ply.__init__ estimates them on every
    // cloud, point-to-plane ICP never reads them): the helper raises `icp_ready` (a CUDA event on its stream + a host flag
    // once the event has been recorded) after the former, and the main stream waits for THAT — on the device, without a
    // host wake-up — while the source normals finish beside the first ICP passes; the call still ends only when the
    // helper is done.
    std::atomic<int> *phase = new std::atomic<int>(0);  // 0: not yet, 1: icp_ready recorded, 2: failed before it
    cudaEvent_t icp_ready = nullptr;
    struct PhaseGuard {  // destroyed after the helper has been joined (declared before the join guard)
        std::atomic<int> *p;
        cudaEvent_t *ev;
        ~PhaseGuard() {
            delete p;
            if (*ev) cudaEventDestroy(*ev);
        }
    } phase_guard{phase, &icp_ready};
    if (overlap) PCR_CUDA(cudaEventCreateWithFlags(&icp_ready, cudaEventDisableTiming));
    auto full_normals = [=, &tfn, &icp_prep](pcr_ctx *c) -> int {
        pcr_ctx *ctx = c;
        auto part1 = [&]() -> int {
            PCR_ALLOC(t, float4, (size_t)nt);
            PCR_TRY(pcr_normals_impl(ctx, tgt, nt, 2.0 * v, 30, t));
            tfn = t;
            // the ICP search structures depend only on the clouds: built here, off the critical path
            if (c != main_ctx) PCR_TRY(pcr_icp_prepare(ctx, src, ns, tgt, nt, 0.4 * v, &icp_prep));
            return PCR_OK;
        };
        const int r1 = part1();
        if (c != main_ctx) {
            if (r1 == PCR_OK && cudaEventRecord(icp_ready, ctx->stream) == cudaSuccess) phase->store(1, std::memory_order_release);
            else phase->store(2, std::memory_order_release);
        }
        if (r1 != PCR_OK) return r1;
        if (p->source_normals) {
            PCR_ALLOC(sfn, float4, (size_t)ns);
            PCR_TRY(pcr_normals_impl(ctx, src, ns, 2.0 * v, 30, sfn));
        }
        return PCR_OK;
    };
    if (overlap) {
        ctx->worker->submit([=]() -> int {
            cudaSetDevice(h->device);
            const int r = full_normals(h);
            if (phase->load(std::memory_order_acquire) == 0) phase->store(2, std::memory_order_release);
            if (r != PCR_OK) return r;
            return pcr_sync_stream(h, h->stream) == cudaSuccess ? PCR_OK : PCR_ERR_CUDA;
        });
    }
    // the helper's job refers to locals of this call: it is joined on EVERY return path
    struct JoinGuard {
        Worker *w;
        bool armed;
        ~JoinGuard() {
            if (armed) w->wait();
        }
    } join_guard{overlap ? ctx->worker : nullptr, overlap};
    // The RANSAC search structures (target grid, Morton-ordered source, candidate lists: ~150 us of small kernels with no
    // host synchronisation) depend only on the down-sampled clouds: they are built on an auxiliary stream NEXT TO the
    // descriptor matching instead of between matching and RANSAC (timeline: tools/gpu_timeline.py).
    RansacWork rwork;
    bool rwork_ok = false;
    cudaEvent_t rprep_done = nullptr;
    if (overlap && ctx->aux_stream && ms > 0 && mt > 0) {
        cudaEvent_t down_ready;
        PCR_CUDA(cudaEventCreateWithFlags(&down_ready, cudaEventDisableTiming));
        PCR_CUDA(cudaEventRecord(down_ready, ctx->stream));
        PCR_CUDA(cudaStreamWaitEvent(ctx->aux_stream, down_ready, 0));
        PCR_CUDA(cudaEventDestroy(down_ready));
        cudaStream_t keep = ctx->stream;
        ctx->stream = ctx->aux_stream;
        const int rcp = pcr_ransac_prepare(ctx, sd, ms, td, mt, 1.5 * v, &rwork);
        ctx->stream = keep;
        if (rcp != PCR_OK) {
            if (ready) cudaEventDestroy(ready);
            return rcp;
        }
        PCR_CUDA(cudaEventCreateWithFlags(&rprep_done, cudaEventDisableTiming));
        PCR_CUDA(cudaEventRecord(rprep_done, ctx->aux_stream));
        rwork_ok = true;
    }
    // global_registration (src/matcher/ransac.py:41-59): mutual filter True, threshold 1.5 v
    int c = 0;
    int *corr = arena<int>(ctx, 2 * (size_t)ms);
    rc = corr ? pcr_match_impl(ctx, sf, ms, tf, mt, 1, 0.1, corr, &c) : PCR_ERR_OOM;
    tm.mark();
    if (rprep_done) {
        cudaStreamWaitEvent(ctx->stream, rprep_done, 0);
        cudaEventDestroy(rprep_done);
    }
    if (rc == PCR_OK)
        rc = pcr_ransac_impl(ctx, sd, ms, td, mt, corr, c, 1.5 * v, 0.9, p->ransac_max_iter, p->ransac_confidence, p->seed, &res->ransac,
                             rwork_ok ? &rwork : nullptr);
    tm.mark();
    static const bool icp_early = !(getenv("PCR_ICP_EARLY") && atoi(getenv("PCR_ICP_EARLY")) == 0);
    auto join_helper = [&]() -> int {
        if (!join_guard.armed) return PCR_OK;
        join_guard.armed = false;
        const int rch = ctx->worker->wait();
        return rch == PCR_OK ? PCR_OK : pcr_fail(ctx, rch, "full-resolution normals: %s", h->err.c_str());
    };
    if (overlap) {
        if (rc == PCR_OK && icp_early) {
            int ph;
            while ((ph = phase->load(std::memory_order_acquire)) == 0) std::this_thread::yield();  // host-side enqueueing only
            if (ph == 1) PCR_CUDA(cudaStreamWaitEvent(ctx->stream, icp_ready, 0));
            else {
                const int rj = join_helper();
                rc = rj != PCR_OK ? rj : pcr_fail(ctx, PCR_ERR_CUDA, "full-resolution normals: the helper could not record its event");
            }
        } else {
            const int rj = join_helper();
            if (rc == PCR_OK) rc = rj;
        }
    } else if (rc == PCR_OK) {
        rc = full_normals(ctx);
    }
    tm.mark();
    // refine_registration (src/matcher/icp.py:41-48): full-resolution clouds, threshold 0.4 v
    if (rc == PCR_OK)
        rc = pcr_icp_impl(ctx, src, ns, tgt, tfn, nt, 0.4 * v, res->ransac.transformation, p->icp_max_iter, p->icp_rel_fitness,
                          p->icp_rel_rmse, &res->icp, nullptr, true, &icp_prep);
    tm.mark();
    {
        const int rj = join_helper();
        if (rc == PCR_OK) rc = rj;
    }
    if (ready) cudaEventDestroy(ready);
    if (rc != PCR_OK) return rc;
    PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
    res->n_src_down = ms;
    res->n_tgt_down = mt;
    res->n_corr = c;
    for (int i = 0; i < 7; i++) cudaEventElapsedTime(&res->stage_ms[i], tm.ev[i], tm.ev[i + 1]);
    cudaEventElapsedTime(&res->stage_ms[7], tm.ev[0], tm.ev[7]);
    return PCR_OK;
}

int pcr_align_device_impl(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, int nt, const pcr_align_params *p,
                          pcr_align_result *res) {
    return align_device(ctx, src, ns, tgt, nt, p, res);
}

extern "C" {

int pcr_align(pcr_ctx *ctx, const float *src, int ns, const float *tgt, int nt, const pcr_align_params *p,
              pcr_align_result *result) {
    PCR_ENTER();
    PCR_ARG(p && result);
    return align_device(ctx, (const float4 *)src, ns, (const float4 *)tgt, nt, p, result);
}

int pcr_align_host(pcr_ctx *ctx, const float *src_xyz, int ns, const float *tgt_xyz, int nt, const pcr_align_params *p,
                   pcr_align_result *result) {
    PCR_ENTER();
    PCR_ARG(p && result && ns >= 0 && nt >= 0);
    if (ns == 0 || nt == 0) return pcr_fail(ctx, PCR_ERR_INVALID, "Point cloud is empty");
    PCR_ALLOC(s3, float, 3 * (size_t)ns);
    PCR_ALLOC(t3, float, 3 * (size_t)nt);
    PCR_ALLOC(s4, float4, (size_t)ns);
    PCR_ALLOC(t4, float4, (size_t)nt);
    PCR_CUDA(cudaMemcpyAsync(s3, src_xyz, sizeof(float) * 3 * (size_t)ns, cudaMemcpyHostToDevice, ctx->stream));
    PCR_CUDA(cudaMemcpyAsync(t3, tgt_xyz, sizeof(float) * 3 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    // pack without leaving the call (pcr_pack_xyz_f32 would re-enter the guard)
    PCR_TRY(pcr_pack_impl(ctx, s3, ns, s4));
    PCR_TRY(pcr_pack_impl(ctx, t3, nt, t4));
    return align_device(ctx, s4, ns, t4, nt, p, result);
}

int pcr_align_files(pcr_ctx *ctx, const char *src_path, const char *tgt_path, const pcr_align_params *p,
                    pcr_align_result *result) {
    PCR_ENTER();
    PCR_ARG(src_path && tgt_path && p && result);
    pcr_ply_info is, it;
    char err_s[256], err_t[256];
    int rc = pcr_ply_probe(src_path, &is, err_s, (int)sizeof err_s);
    if (rc != PCR_OK) return pcr_fail(ctx, rc, "%s: %s", src_path, err_s);
    rc = pcr_ply_probe(tgt_path, &it, err_t, (int)sizeof err_t);
    if (rc != PCR_OK) return pcr_fail(ctx, rc, "%s: %s", tgt_path, err_t);
    if (is.n_vertex == 0 || it.n_vertex == 0) return pcr_fail(ctx, PCR_ERR_INVALID, "Point cloud is empty");
    PCR_ARG(is.n_vertex <= 0x7fffffff && it.n_vertex <= 0x7fffffff);
    const int ns = (int)is.n_vertex, nt = (int)it.n_vertex;
    const size_t need = sizeof(float4) * ((size_t)ns + (size_t)nt);
    if (ctx->stage_bytes < need) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr;
        ctx->stage_bytes = 0;
        const size_t cap = need + need / 4;
        if (cudaMallocHost(&ctx->stage, cap) != cudaSuccess) {
            cudaGetLastError();
            ctx->stage = nullptr;
            return pcr_fail(ctx, PCR_ERR_OOM, "cannot pin %zu bytes of host memory for the file staging buffer", cap);
        }
        ctx->stage_bytes = cap;
    }
    float *hs = (float *)ctx->stage, *ht = hs + 4 * (size_t)ns;
    // the two files are decoded concurrently (each decode is itself multi-threaded for large files)
    int rc_t = PCR_OK;
    bool threaded = false;
    std::thread th;
    try {
        th = std::thread([&] { rc_t = pcr_ply_read(tgt_path, nt, ht, nullptr, nullptr, 0, nullptr, err_t, (int)sizeof err_t); });
        threaded = true;
    } catch (...) {
    }
    rc = pcr_ply_read(src_path, ns, hs, nullptr, nullptr, 0, nullptr, err_s, (int)sizeof err_s);
    if (threaded) th.join();
    else rc_t = pcr_ply_read(tgt_path, nt, ht, nullptr, nullptr, 0, nullptr, err_t, (int)sizeof err_t);
    if (rc != PCR_OK) return pcr_fail(ctx, rc, "%s: %s", src_path, err_s);
    if (rc_t != PCR_OK) return pcr_fail(ctx, rc_t, "%s: %s", tgt_path, err_t);
    PCR_ALLOC(s4, float4, (size_t)ns);
    PCR_ALLOC(t4, float4, (size_t)nt);
    PCR_CUDA(cudaMemcpyAsync(s4, hs, sizeof(float4) * (size_t)ns, cudaMemcpyHostToDevice, ctx->stream));
    PCR_CUDA(cudaMemcpyAsync(t4, ht, sizeof(float4) * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    rc = align_device(ctx, s4, ns, t4, nt, p, result);
    if (rc != PCR_OK) cudaStreamSynchronize(ctx->stream);  // the staging buffer is reused by the next call
    return rc;
}

}  // extern "C"
