// pcr_dist.cu — the two shardings of the path on the C side of the boundary (SURVEY.md §8b "pcr_nccl_select_best",
// "pcr_align_batch"; §8e), so that a C / non-PyTorch caller can use more than one GPU and the per-wave exchange does not
// run through an interpreter:
//   pcr_ransac_multi   RANSAC hypotheses sharded over the ranks of an NCCL communicator: every wave [begin, end) of the
//                      global hypothesis stream is cut into contiguous per-rank slices; each rank scores its slice
//                      (pcr_ransac_wave_impl), keeps the chain of prefix maxima, ONE ncclAllGather of a fixed 3 KB block per
//                      rank on the compute stream exchanges the chains, and every rank replays the sequential loop
//                      (pcr_ransac_scan).  Philox is keyed by the global hypothesis index, so the result is identical for
//                      every world size (and to pcr_ransac).
//   pcr_align_batch    batches of independent pairs, pair i -> rank i mod world, aligned by `workers` host threads per
//                      rank (own context + stream each); no communication until a final ncclAllGather of 18 doubles per pair.
// One process per GPU.  NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch
// already loaded, in a plain C program the system one), so the library itself has no link-time NCCL dependency and
// single-GPU users never touch it.  world == 1 needs no communicator at all.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <sched.h>

#include "pcr_common.cuh"

typedef unsigned long long u64;

int pcr_ransac_prepare(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, double max_dist, RansacWork *w);
int pcr_ransac_session_begin_impl(pcr_ctx *ctx, const float4 *src, int ms, const float4 *tgt, int mt, double max_dist);
int pcr_ransac_session_end_impl(pcr_ctx *ctx);
int pcr_ransac_wave_impl(pcr_ctx *ctx, const RansacWork &w, const float4 *src, int ms, const float4 *tgt, const int *corr, int c,
                         double max_dist, double edge_sim, long long hyp_begin, long long hyp_end, u64 seed, long long best_cnt,
                         long long best_sumq, pcr_hyp_record *recs_host, int cap, int *n_recs_host, long long *n_surv_host);
int pcr_corr_check_impl(pcr_ctx *ctx, const int *corr, int c, int ms, int mt);
int pcr_wave_generate(pcr_ctx *ctx, const float4 *src, const float4 *tgt, const int *corr, int c, double max_dist, double edge_sim,
                      long long hyp_begin, long long hyp_end, u64 seed, int cap, WaveWork *ww);
int pcr_wave_validate(pcr_ctx *ctx, const RansacWork &w, const float4 *src, int ms, const float4 *tgt, const int *corr, int c,
                      double max_dist, const WaveWork &ww, long long best_cnt, long long best_sumq, pcr_hyp_record *recs_host,
                      int *n_recs_host, long long *n_surv_host);
int pcr_align_device_impl(pcr_ctx *ctx, const float4 *src, int ns, const float4 *tgt, int nt, const pcr_align_params *p,
                          pcr_align_result *res);

// ---- NCCL, bound at run time ------------------------------------------------------------------------------------------
namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclChar = 0 };  // ncclInt8 / ncclChar = 0 in nccl.h (all 2.x releases)

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("PCR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.err = "cannot load libnccl.so.2 (set PCR_NCCL_LIB)";
            return;
        }
        api.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
        api.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllGather");
        api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) {
            api.err = "libnccl is missing a required symbol";
            api.lib = nullptr;
        }
    });
    return &api;
}

struct DistState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    void *dev_buf = nullptr;   // send block followed by the gathered blocks
    size_t dev_bytes = 0;
    void *host_buf = nullptr;  // pinned mirror
    size_t host_bytes = 0;
    std::vector<pcr_ctx *> workers;  // pcr_align_batch: extra contexts of this rank (worker 0 is the caller's context)
    cudaStream_t spec_stream = nullptr;  // pcr_ransac_multi: the next wave's hypotheses are generated here
};

DistState *dist_of(pcr_ctx *ctx) {
    if (!ctx->dist) ctx->dist = new DistState();
    return (DistState *)ctx->dist;
}

int dist_buffers(pcr_ctx *ctx, DistState *d, size_t bytes) {
    if (d->dev_bytes < bytes) {
        if (d->dev_buf) cudaFree(d->dev_buf);
        if (d->host_buf) cudaFreeHost(d->host_buf);
        d->dev_buf = d->host_buf = nullptr;
        d->dev_bytes = d->host_bytes = 0;
        PCR_CUDA(cudaMalloc(&d->dev_buf, bytes));
        PCR_CUDA(cudaMallocHost(&d->host_buf, bytes));
        d->dev_bytes = d->host_bytes = bytes;
    }
    return PCR_OK;
}

// every rank contributes `bytes` from send_host; recv_host receives world x bytes (rank order).  world == 1: a copy.
int dist_allgather(pcr_ctx *ctx, DistState *d, const void *send_host, size_t bytes, void *recv_host) {
    if (d->world == 1 || !d->comm) {
        memcpy(recv_host, send_host, bytes);
        return PCR_OK;
    }
    NcclApi *api = nccl_api();
    PCR_TRY(dist_buffers(ctx, d, bytes * (size_t)(d->world + 1)));
    char *dsend = (char *)d->dev_buf, *drecv = dsend + bytes;
    char *hsend = (char *)d->host_buf, *hrecv = hsend + bytes;
    memcpy(hsend, send_host, bytes);
    PCR_CUDA(cudaMemcpyAsync(dsend, hsend, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = api->AllGather(dsend, drecv, bytes, ncclChar, d->comm, ctx->stream);
    if (rc != ncclSuccess)
        return pcr_fail(ctx, PCR_ERR_CUDA, "ncclAllGather: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
    PCR_CUDA(cudaMemcpyAsync(hrecv, drecv, bytes * (size_t)d->world, cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(recv_host, hrecv, bytes * (size_t)d->world);
    return PCR_OK;
}

struct Guard {  // the same one-call-at-a-time rule as every other export
    pcr_ctx *ctx;
    bool ok = false;
    explicit Guard(pcr_ctx *c) : ctx(c) {
        bool expected = false;
        if (ctx && ctx->busy.compare_exchange_strong(expected, true, std::memory_order_acquire)) {
            ok = true;
            cudaSetDevice(ctx->device);
            pcr_arena_reset(ctx);
            ctx->bounds_cache.clear();
        }
    }
    ~Guard() {
        if (ok) ctx->busy.store(false, std::memory_order_release);
    }
};

constexpr int FIXED_CAP = 24;  // records per rank carried by the per-wave all-gather (a chain of prefix maxima is ~ln n long)

}  // namespace

void pcr_dist_free(pcr_ctx *ctx) {  // called by pcr_destroy
    DistState *d = (DistState *)ctx->dist;
    if (!d) return;
    for (pcr_ctx *w : d->workers) pcr_destroy(w);
    if (d->comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy(d->comm);
    if (d->dev_buf) cudaFree(d->dev_buf);
    if (d->host_buf) cudaFreeHost(d->host_buf);
    delete d;
    ctx->dist = nullptr;
}

extern "C" {

int pcr_comm_unique_id(void *id_out, int bytes) {
    if (!id_out || bytes < (int)sizeof(ncclUniqueId)) return PCR_ERR_INVALID;
    NcclApi *api = nccl_api();
    if (!api->lib) return PCR_ERR_IO;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return PCR_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return PCR_OK;
}

int pcr_comm_init(pcr_ctx *ctx, const void *id, int bytes, int rank, int world) {
    if (!ctx) return PCR_ERR_INVALID;
    Guard g(ctx);
    if (!g.ok) return PCR_ERR_BUSY;
    if (world < 1 || rank < 0 || rank >= world) return pcr_fail(ctx, PCR_ERR_INVALID, "comm init: rank %d of %d", rank, world);
    DistState *d = dist_of(ctx);
    if (d->comm) {
        nccl_api()->CommDestroy(d->comm);
        d->comm = nullptr;
    }
    d->rank = rank;
    d->world = world;
    if (world == 1) return PCR_OK;
    if (!id || bytes < (int)sizeof(ncclUniqueId)) return pcr_fail(ctx, PCR_ERR_INVALID, "comm init: a %zu-byte id is required", sizeof(ncclUniqueId));
    NcclApi *api = nccl_api();
    if (!api->lib) return pcr_fail(ctx, PCR_ERR_IO, "%s", api->err.c_str());
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const int rc = api->CommInitRank(&d->comm, world, uid, rank);
    if (rc != ncclSuccess) {
        d->comm = nullptr;
        return pcr_fail(ctx, PCR_ERR_CUDA, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
    }
    return PCR_OK;
}

int pcr_comm_destroy(pcr_ctx *ctx) {
    if (!ctx) return PCR_ERR_INVALID;
    Guard g(ctx);
    if (!g.ok) return PCR_ERR_BUSY;
    DistState *d = (DistState *)ctx->dist;
    if (d && d->comm) {
        nccl_api()->CommDestroy(d->comm);
        d->comm = nullptr;
    }
    if (d) {
        d->world = 1;
        d->rank = 0;
    }
    return PCR_OK;
}

int pcr_ransac_multi(pcr_ctx *ctx, const float *src_, int ms, const float *tgt_, int mt, const int *corr, int c, double max_dist,
                     double edge_sim, int64_t max_iter, double confidence, uint64_t seed, int64_t first_wave, int growth,
                     pcr_reg_result *res, int *n_waves) {
    if (!ctx) return PCR_ERR_INVALID;
    Guard g(ctx);
    if (!g.ok) return PCR_ERR_BUSY;
    if (!res || ms < 0 || mt < 0 || c < 0) return pcr_fail(ctx, PCR_ERR_INVALID, "ransac_multi: invalid argument");
    const float4 *src = (const float4 *)src_, *tgt = (const float4 *)tgt_;
    DistState *d = dist_of(ctx);
    const int world = d->comm ? d->world : 1, rank = d->comm ? d->rank : 0;
    memset(res, 0, sizeof(*res));
    res->transformation[0] = res->transformation[5] = res->transformation[10] = res->transformation[15] = 1.0;
    res->best_hyp = -1;
    res->est_k = max_iter;
    if (n_waves) *n_waves = 0;
    if (c < 3 || !(max_dist > 0.0) || ms == 0 || mt == 0 || max_iter <= 0) return PCR_OK;
    PCR_TRY(pcr_corr_check_impl(ctx, corr, c, ms, mt));
    RansacWork w;
    PCR_TRY(pcr_ransac_prepare(ctx, src, ms, tgt, mt, max_dist, &w));
    // measured at 10M hypotheses (tools/gpu_dist_check.py), ms on 1 / 2 / 8 GPUs: (4096, x4) 32.6 / 18.2 / -,
    // (2048, x8) 28.6 / 15.0 / 4.88, (16384, x4) 28.7 / 15.0 / 4.43
    if (first_wave <= 0) first_wave = 16384;
    if (growth < 2) growth = 4;
    const int64_t max_wave = (int64_t)1 << 22;
    int64_t begin = 0, wave = first_wave * world, survivors = 0;
    int waves = 0;
    constexpr int REC_WORDS = sizeof(pcr_hyp_record) / 8;
    const size_t blk_words = 2 + (size_t)FIXED_CAP * REC_WORDS;
    std::vector<pcr_hyp_record> recs, merged;
    std::vector<long long> sendb(blk_words), recvb(blk_words * (size_t)world);
    // The hypotheses of wave k + 1 are generated on a second stream while wave k is validated, read back, exchanged and
    // replayed (generation needs nothing of that; it is ~20 % of a run's GPU work): the GPU no longer idles while the ranks
    // wait for the slowest of them in the all-gather.  Two survivor buffers, sized for the largest slice of the schedule,
    // alternate.  PCR_DIST_SPECULATE=0: one wave after the other.
    static const bool spec_on = !(getenv("PCR_DIST_SPECULATE") && atoi(getenv("PCR_DIST_SPECULATE")) == 0);
    if (spec_on && !d->spec_stream && cudaStreamCreateWithFlags(&d->spec_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        d->spec_stream = nullptr;
    }
    const bool speculate = spec_on && d->spec_stream != nullptr;
    constexpr int CAP0 = 4096;
    WaveWork bufs[2];
    if (speculate) {
        int64_t b = 0, wv = wave, max_slice = 1;
        while (b < max_iter) {
            const int64_t e = std::min<int64_t>(max_iter, b + wv), nn = e - b;
            max_slice = std::max<int64_t>(max_slice, (nn + world - 1) / world + 1);
            b = e;
            if (wv < max_wave * world) wv *= growth;
        }
        for (int k = 0; k < 2; k++) {
            bufs[k].surv = (Survivor *)arena<unsigned char>(ctx, pcr_wave_survivor_bytes(max_slice));
            bufs[k].hdr = arena<unsigned char>(ctx, 16 + sizeof(pcr_hyp_record) * (size_t)CAP0);
            bufs[k].bucket_best = arena<int>(ctx, 64);
            if (!bufs[k].surv || !bufs[k].hdr || !bufs[k].bucket_best) return PCR_ERR_OOM;
        }
    }
    cudaEvent_t spec_done = nullptr;  // non-null: bufs[cur] holds the generated survivors of the wave that starts at `begin`
    int cur = 0;
    struct SpecGuard {  // a speculated wave that is never used still runs on this call's arena: wait for it on every path
        cudaEvent_t *ev;
        cudaStream_t s;
        ~SpecGuard() {
            if (*ev) {
                cudaEventDestroy(*ev);
                cudaStreamSynchronize(s);
            }
        }
    } spec_guard{&spec_done, d->spec_stream};
    while (begin < max_iter && begin < res->est_k) {
        const int64_t end = std::min<int64_t>(max_iter, begin + wave);
        const int64_t n = end - begin;
        const int64_t lo = begin + n * rank / world, hi = begin + n * (rank + 1) / world;
        const int64_t wave_next = wave < max_wave * world ? wave * growth : wave;
        int nrec = 0;
        long long nsurv = 0;
        if (hi > lo) {
            int cap = CAP0;
            recs.resize((size_t)cap);
            int rc;
            if (speculate) {
                WaveWork &ww = bufs[cur];
                if (spec_done) {  // generated during the previous wave
                    PCR_CUDA(cudaStreamWaitEvent(ctx->stream, spec_done, 0));
                    PCR_CUDA(cudaEventDestroy(spec_done));
                    spec_done = nullptr;
                    rc = PCR_OK;
                } else {
                    rc = pcr_wave_generate(ctx, src, tgt, corr, c, max_dist, edge_sim, lo, hi, seed, cap, &ww);
                }
                // this rank's slice of the next wave, on the other stream and in the other buffer
                const int64_t end2 = std::min<int64_t>(max_iter, end + wave_next), n2 = end2 - end;
                const int64_t lo2 = end + n2 * rank / world, hi2 = end + n2 * (rank + 1) / world;
                if (rc == PCR_OK && hi2 > lo2) {
                    cudaEvent_t in_ready;
                    PCR_CUDA(cudaEventCreateWithFlags(&in_ready, cudaEventDisableTiming));
                    PCR_CUDA(cudaEventRecord(in_ready, ctx->stream));  // inputs complete; the other buffer's last reader (wave k - 1) is done
                    PCR_CUDA(cudaStreamWaitEvent(d->spec_stream, in_ready, 0));
                    PCR_CUDA(cudaEventDestroy(in_ready));
                    cudaStream_t keep = ctx->stream;
                    ctx->stream = d->spec_stream;
                    const int rcs = pcr_wave_generate(ctx, src, tgt, corr, c, max_dist, edge_sim, lo2, hi2, seed, CAP0, &bufs[cur ^ 1]);
                    ctx->stream = keep;
                    if (rcs == PCR_OK && cudaEventCreateWithFlags(&spec_done, cudaEventDisableTiming) == cudaSuccess) {
                        cudaEventRecord(spec_done, d->spec_stream);
                    } else {
                        spec_done = nullptr;
                        cudaStreamSynchronize(d->spec_stream);
                    }
                }
                if (rc == PCR_OK)
                    rc = pcr_wave_validate(ctx, w, src, ms, tgt, corr, c, max_dist, ww, res->inlier_count, res->sum_d2_fixed, recs.data(), &nrec, &nsurv);
            } else {
                rc = pcr_ransac_wave_impl(ctx, w, src, ms, tgt, corr, c, max_dist, edge_sim, lo, hi, seed, res->inlier_count,
                                          res->sum_d2_fixed, recs.data(), cap, &nrec, &nsurv);
            }
            // a slice whose records did not fit: generated and validated again with a larger record buffer (arena scratch)
            while (rc == PCR_ERR_INVALID && cap < (1 << 24) && (int64_t)cap < hi - lo) {
                cap *= 16;
                recs.resize((size_t)cap);
                rc = pcr_ransac_wave_impl(ctx, w, src, ms, tgt, corr, c, max_dist, edge_sim, lo, hi, seed, res->inlier_count,
                                          res->sum_d2_fixed, recs.data(), cap, &nrec, &nsurv);
            }
            if (rc != PCR_OK) return rc;
        }
        // chain of prefix maxima of this rank's slice (records are sorted by hypothesis index)
        std::vector<pcr_hyp_record> chain;
        {
            long long bc = res->inlier_count, bs = res->sum_d2_fixed;
            for (int i = 0; i < nrec; i++) {
                const pcr_hyp_record &r = recs[(size_t)i];
                if (r.inlier_count > bc || (r.inlier_count == bc && bc > 0 && r.sum_d2_fixed < bs)) {
                    chain.push_back(r);
                    bc = r.inlier_count;
                    bs = r.sum_d2_fixed;
                }
            }
        }
        merged.clear();
        if (world > 1) {
            std::fill(sendb.begin(), sendb.end(), 0);
            sendb[0] = (long long)chain.size();
            sendb[1] = nsurv;
            const size_t k = std::min<size_t>(chain.size(), FIXED_CAP);
            if (k) memcpy(&sendb[2], chain.data(), k * sizeof(pcr_hyp_record));
            PCR_TRY(dist_allgather(ctx, d, sendb.data(), blk_words * 8, recvb.data()));
            long long maxc = 0;
            for (int r = 0; r < world; r++) {
                maxc = std::max(maxc, recvb[(size_t)r * blk_words]);
                survivors += recvb[(size_t)r * blk_words + 1];
            }
            if (maxc <= FIXED_CAP) {
                for (int r = 0; r < world; r++) {
                    const long long cr = recvb[(size_t)r * blk_words];
                    const pcr_hyp_record *p = (const pcr_hyp_record *)&recvb[(size_t)r * blk_words + 2];
                    merged.insert(merged.end(), p, p + cr);
                }
            } else {  // a chain did not fit: second exchange with the common capacity
                std::vector<pcr_hyp_record> mine((size_t)maxc), all((size_t)maxc * world);
                memset(mine.data(), 0, sizeof(pcr_hyp_record) * (size_t)maxc);
                if (!chain.empty()) memcpy(mine.data(), chain.data(), chain.size() * sizeof(pcr_hyp_record));
                PCR_TRY(dist_allgather(ctx, d, mine.data(), sizeof(pcr_hyp_record) * (size_t)maxc, all.data()));
                for (int r = 0; r < world; r++)
                    merged.insert(merged.end(), all.begin() + (size_t)r * maxc, all.begin() + (size_t)r * maxc + recvb[(size_t)r * blk_words]);
            }
        } else {
            merged = chain;
            survivors += nsurv;
        }
        int stop = 0;
        pcr_ransac_scan(merged.data(), (int)merged.size(), begin, end, c, ms, confidence, w.k_d, res, &stop);
        waves++;
        begin = end;
        if (stop) break;
        wave = wave_next;
        cur ^= 1;
    }
    res->survivors = survivors;
    res->k_d = w.k_d;
    if (res->hyp_evaluated > max_iter) res->hyp_evaluated = max_iter;
    if (n_waves) *n_waves = waves;
    return PCR_OK;
}

int pcr_align_batch(pcr_ctx *ctx, int n_local, const float *const *src_dev, const int *ns, const float *const *tgt_dev, const int *nt,
                    const pcr_align_params *p, int workers, int n_total, double *out_host) {
    if (!ctx) return PCR_ERR_INVALID;
    Guard g(ctx);
    if (!g.ok) return PCR_ERR_BUSY;
    if (n_local < 0 || n_total < 0 || !p || (n_total > 0 && !out_host)) return pcr_fail(ctx, PCR_ERR_INVALID, "align_batch: invalid argument");
    DistState *d = dist_of(ctx);
    const int world = d->comm ? d->world : 1, rank = d->comm ? d->rank : 0;
    const int per = (n_total + world - 1) / world;
    const int expect = n_total > rank ? (n_total - rank + world - 1) / world : 0;
    if (n_local != expect) return pcr_fail(ctx, PCR_ERR_INVALID, "align_batch: rank %d of %d holds %d pairs, expected %d of %d", rank, world, n_local, expect, n_total);
    if (workers < 1) workers = 1;
    if (workers > 16) workers = 16;
    if (workers > n_local) workers = n_local > 0 ? n_local : 1;
    while ((int)d->workers.size() < workers - 1) {
        pcr_ctx *wctx = nullptr;
        const int rc = pcr_create(ctx->device, &wctx);
        if (rc != PCR_OK) return pcr_fail(ctx, rc, "align_batch: cannot create a worker context");
        if (cudaStreamCreateWithFlags(&wctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            pcr_destroy(wctx);
            return pcr_fail(ctx, PCR_ERR_CUDA, "align_batch: cannot create a worker stream");
        }
        wctx->owns_stream = true;
        d->workers.push_back(wctx);
    }
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));  // the inputs were produced on the caller's stream
    // Host waits of the workers: every alignment waits ~7 times for the GPU, and a spinning wait holds a core.  With one rank
    // per GPU, `workers` contexts per rank and a helper thread per context there can be more host threads than cores (8 ranks
    // x 6 workers x 2 on a 32-core box): the threads then block on an event instead (PCR_BATCH_BLOCKING=0/1 overrides).
    {
        static const int bl_env = getenv("PCR_BATCH_BLOCKING") ? atoi(getenv("PCR_BATCH_BLOCKING")) : -1;
        cpu_set_t set;
        int cores = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        if (cores < 1) cores = 1;
        // measured on one B200 with the affinity cut to 2 / 4 / 16 cores (tools/gpu_batch_workers.py, 6 workers = 12 threads):
        // spinning 1,238 / 1,513 / 1,734 pairs/s, blocking 1,615 / 1,620 / 1,678 — and 1,256 -> 534 with a single worker,
        // so blocking only when the threads outnumber the cores at least two to one
        const bool blocking = bl_env >= 0 ? bl_env != 0 : (long long)world * workers * 2 >= 2LL * cores;
        ctx->blocking_sync = blocking;
        if (ctx->helper) ctx->helper->blocking_sync = blocking;
        for (pcr_ctx *wc : d->workers) {
            wc->blocking_sync = blocking;
            if (wc->helper) wc->helper->blocking_sync = blocking;
        }
    }
    struct SyncRestore {  // single alignments on this context spin again
        pcr_ctx *c;
        ~SyncRestore() {
            c->blocking_sync = false;
            if (c->helper) c->helper->blocking_sync = false;
        }
    } sync_restore{ctx};
    std::vector<double> mine((size_t)per * 18, 0.0);
    std::vector<int> rcs((size_t)workers, PCR_OK);
    auto work = [&](int wi) {
        pcr_ctx *c = wi == 0 ? ctx : d->workers[(size_t)wi - 1];
        cudaSetDevice(c->device);
        for (int k = wi; k < n_local; k += workers) {
            if (wi != 0) {
                pcr_arena_reset(c);
                c->bounds_cache.clear();
            }
            pcr_align_result r;
            const int rc = pcr_align_device_impl(c, (const float4 *)src_dev[k], ns[k], (const float4 *)tgt_dev[k], nt[k], p, &r);
            if (rc != PCR_OK) {
                rcs[(size_t)wi] = rc;
                if (c != ctx) ctx->err = c->err;
                return;
            }
            double *o = &mine[(size_t)k * 18];
            for (int i = 0; i < 16; i++) o[i] = r.icp.transformation[i];
            o[16] = r.icp.fitness;
            o[17] = r.icp.inlier_rmse;
            if (wi == 0) {  // the caller's context keeps its arena for the whole call otherwise: release per pair
                pcr_arena_reset(c);
                c->bounds_cache.clear();
            }
        }
    };
    std::vector<std::thread> th;
    for (int wi = 1; wi < workers; wi++) th.emplace_back(work, wi);
    work(0);
    for (auto &t : th) t.join();
    for (int rc : rcs)
        if (rc != PCR_OK) return rc;
    std::vector<double> all((size_t)per * 18 * world, 0.0);
    PCR_TRY(dist_allgather(ctx, d, mine.data(), sizeof(double) * (size_t)per * 18, all.data()));
    for (int r = 0; r < world; r++)
        for (int k = 0; r + k * world < n_total; k++)
            memcpy(out_host + (size_t)(r + k * world) * 18, &all[((size_t)r * per + k) * 18], 18 * sizeof(double));
    return PCR_OK;
}

}  // extern "C"
