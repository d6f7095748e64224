// pcr_ctx.cu — context, scratch arena, error plumbing, layout helpers.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "pcr_common.cuh"

std::atomic<long long> g_pcr_launches{0};

int pcr_fail(pcr_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

static cudaEvent_t ev_get(pcr_ctx *ctx) {
    if (!ctx->ev_pool.empty()) {
        cudaEvent_t e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

KScope::KScope(pcr_ctx *c, int id, double bytes_per_launch, long long launches, double flops_per_launch)
    : ctx(c), on(c && c->profiling) {
    p.id = id;
    p.bytes = bytes_per_launch;
    p.flops = flops_per_launch;
    p.launches = launches;
    if (on) {
        p.a = ev_get(ctx);
        p.b = ev_get(ctx);
        cudaEventRecord(p.a, ctx->stream);
    }
}

KScope::~KScope() {
    if (on) {
        cudaEventRecord(p.b, ctx->stream);
        ctx->pending.push_back(p);
    }
}

int pcr_pow2ceil_exp(double x) {
    int e;
    const double m = frexp(x, &e);
    return (m == 0.5) ? e - 1 : e;
}

void pcr_arena_reset(pcr_ctx *ctx) {
    ctx->bounds_cache.clear();  // caller buffers may have changed since the previous call
    // keep only the largest block; free the rest so the arena converges to one block
    if (ctx->blocks.size() > 1) {
        size_t total = 0;
        for (size_t s : ctx->block_sizes) total += s;
        for (void *b : ctx->blocks) cudaFree(b);
        ctx->blocks.clear();
        ctx->block_sizes.clear();
        void *p = nullptr;
        if (cudaMalloc(&p, total) == cudaSuccess) {
            ctx->blocks.push_back(p);
            ctx->block_sizes.push_back(total);
        } else {
            cudaGetLastError();
        }
    }
    ctx->cur_block = 0;
    ctx->cur_off = 0;
}

void *pcr_arena_alloc(pcr_ctx *ctx, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    while (ctx->cur_block < ctx->blocks.size()) {
        if (ctx->cur_off + bytes <= ctx->block_sizes[ctx->cur_block]) {
            void *p = (char *)ctx->blocks[ctx->cur_block] + ctx->cur_off;
            ctx->cur_off += bytes;
            return p;
        }
        ctx->cur_block++;
        ctx->cur_off = 0;
    }
    size_t sz = bytes > ((size_t)64 << 20) ? bytes : ((size_t)64 << 20);
    void *p = nullptr;
    // a fresh cudaMalloc is synchronous with respect to the device; it only happens while the arena warms up
    if (cudaMalloc(&p, sz) != cudaSuccess) {
        cudaGetLastError();
        pcr_fail(ctx, PCR_ERR_OOM, "device allocation of %zu bytes failed", sz);
        return nullptr;
    }
    ctx->blocks.push_back(p);
    ctx->block_sizes.push_back(sz);
    ctx->cur_block = ctx->blocks.size() - 1;
    ctx->cur_off = bytes;
    return p;
}

// ---- worker thread ------------------------------------------------------------------------------------
Worker::Worker() {
    th = std::thread([this] {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [this] { return has_task || quit; });
            if (quit) return;
            std::function<int()> f = std::move(task);
            has_task = false;
            lk.unlock();
            const int r = f();
            lk.lock();
            rc = r;
            done = true;
            cv.notify_all();
        }
    });
}
Worker::~Worker() {
    {
        std::lock_guard<std::mutex> lk(m);
        quit = true;
    }
    cv.notify_all();
    if (th.joinable()) th.join();
}
void Worker::submit(std::function<int()> f) {
    std::lock_guard<std::mutex> lk(m);
    task = std::move(f);
    has_task = true;
    done = false;
    cv.notify_all();
}
int Worker::wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return done; });
    return rc;
}

cudaError_t pcr_sync_stream(pcr_ctx *ctx, cudaStream_t s) {
    if (!ctx->blocking_sync) return cudaStreamSynchronize(s);
    if (!ctx->sync_ev && cudaEventCreateWithFlags(&ctx->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        ctx->sync_ev = nullptr;
        return cudaStreamSynchronize(s);
    }
    const cudaError_t e = cudaEventRecord(ctx->sync_ev, s);
    return e != cudaSuccess ? e : cudaEventSynchronize(ctx->sync_ev);
}

// helper context of `ctx` (created on first use): same device, own non-blocking stream
int pcr_helper_get(pcr_ctx *ctx, pcr_ctx **out) {
    if (!ctx->helper) {
        pcr_ctx *h = nullptr;
        const int rc = pcr_create(ctx->device, &h);
        if (rc != PCR_OK) return pcr_fail(ctx, rc, "cannot create the helper context");
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
            pcr_destroy(h);
            return pcr_fail(ctx, PCR_ERR_CUDA, "cannot create the helper stream");
        }
        h->owns_stream = true;
        h->yielding = getenv("PCR_HELPER_YIELD") ? atoi(getenv("PCR_HELPER_YIELD")) : 2;
        int prio_low = 0, prio_high = 0;
        cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high);
        if (cudaStreamCreateWithPriority(&ctx->hp_stream, cudaStreamNonBlocking, prio_high) != cudaSuccess) {
            cudaGetLastError();
            ctx->hp_stream = nullptr;
        }
        if (cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_high) != cudaSuccess) {
            cudaGetLastError();
            ctx->aux_stream = nullptr;
        }
        if (cudaStreamCreateWithPriority(&ctx->aux2_stream, cudaStreamNonBlocking, prio_high) != cudaSuccess) {
            cudaGetLastError();
            ctx->aux2_stream = nullptr;
        }
        ctx->helper = h;
        ctx->worker = new Worker();
    }
    ctx->helper->profiling = ctx->profiling;
    ctx->helper->blocking_sync = ctx->blocking_sync;
    *out = ctx->helper;
    return PCR_OK;
}

void pcr_dist_free(pcr_ctx *ctx);  // pcr_dist.cu

extern "C" {

int pcr_version(void) { return 100; }

int pcr_create(int device, pcr_ctx **out) {
    if (!out) return PCR_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return PCR_ERR_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess) return PCR_ERR_CUDA;
    pcr_ctx *ctx = new pcr_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        delete ctx;
        return PCR_ERR_CUDA;  // sm_100a only: there is no fallback path
    }
    ctx->pinned_bytes = 1 << 20;
    if (cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) != cudaSuccess) {
        delete ctx;
        return PCR_ERR_OOM;
    }
    *out = ctx;
    return PCR_OK;
}

int pcr_destroy(pcr_ctx *ctx) {
    if (!ctx) return PCR_OK;
    cudaSetDevice(ctx->device);
    if (ctx->worker) {
        ctx->worker->wait();
        delete ctx->worker;
    }
    if (ctx->helper) pcr_destroy(ctx->helper);
    pcr_dist_free(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (void *b : ctx->blocks) cudaFree(b);
    for (void *b : ctx->rsess.bufs)
        if (b) cudaFree(b);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (const KPending &p : ctx->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    if (ctx->bounds_ticket) cudaFree(ctx->bounds_ticket);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->hp_stream) cudaStreamDestroy(ctx->hp_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux2_stream) cudaStreamDestroy(ctx->aux2_stream);
    if (ctx->sync_ev) cudaEventDestroy(ctx->sync_ev);
    delete ctx;
    return PCR_OK;
}

const char *pcr_last_error(pcr_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pcr_set_stream(pcr_ctx *ctx, void *cuda_stream) {
    if (!ctx) return PCR_ERR_INVALID;
    ctx->stream = (cudaStream_t)cuda_stream;
    return PCR_OK;
}

int64_t pcr_launch_count(pcr_ctx *ctx) {
    (void)ctx;
    return g_pcr_launches.load();
}

int pcr_set_profiling(pcr_ctx *ctx, int enabled) {
    if (!ctx) return PCR_ERR_INVALID;
    ctx->profiling = enabled != 0;
    return PCR_OK;
}

static const char *const KNAMES[KC_COUNT] = {
    "pack", "bounds", "grid_build", "scan", "voxel", "knn_list", "knn_cov", "normals_solve", "spfh", "fpfh",
    "nn_features", "match_misc", "ransac_generate", "ransac_validate", "ransac_step", "icp_pass", "nn1"};

int pcr_kernel_class_count(void) { return KC_COUNT; }
const char *pcr_kernel_class_name(int id) { return (id >= 0 && id < KC_COUNT) ? KNAMES[id] : ""; }

int pcr_kernel_stats(pcr_ctx *ctx, pcr_kernel_stat *out, int cap, int reset) {
    if (!ctx || !out || cap < KC_COUNT) return PCR_ERR_INVALID;
    cudaSetDevice(ctx->device);
    PCR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (getenv("PCR_TIMELINE") && !ctx->pending.empty()) {
        // diagnostic: start offset and duration of every timed scope since the last call, both contexts, relative to the
        // first scope of this context (there is no nsys in the image; this is the poor man's timeline)
        if (ctx->helper) cudaStreamSynchronize(ctx->helper->stream);
        const cudaEvent_t t0 = ctx->pending.front().a;
        for (int which = 0; which < 2; which++) {
            pcr_ctx *c = which ? ctx->helper : ctx;
            if (!c) continue;
            for (const KPending &p : c->pending) {
                float off = 0.0f, dur = 0.0f;
                if (cudaEventElapsedTime(&off, t0, p.a) == cudaSuccess && cudaEventElapsedTime(&dur, p.a, p.b) == cudaSuccess)
                    fprintf(stderr, "[pcr timeline] %s %-16s start %8.1f us  dur %7.1f us  (%lld launches)\n", which ? "helper" : "main  ",
                            KNAMES[p.id], off * 1e3, dur * 1e3, p.launches);
                else
                    cudaGetLastError();
            }
        }
    }
    for (const KPending &p : ctx->pending) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            ctx->k_ms[p.id] += ms;
            ctx->k_launches[p.id] += p.launches;
            ctx->k_bytes[p.id] += p.bytes * (double)p.launches;
            ctx->k_flops[p.id] += p.flops * (double)p.launches;
        } else {
            cudaGetLastError();
        }
        ctx->ev_pool.push_back(p.a);
        ctx->ev_pool.push_back(p.b);
    }
    ctx->pending.clear();
    if (ctx->helper) {  // kernels that pcr_align ran on the helper context count for this context
        pcr_ctx *h = ctx->helper;
        cudaStreamSynchronize(h->stream);
        for (const KPending &p : h->pending) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
                ctx->k_ms[p.id] += ms;
                ctx->k_ms_helper[p.id] += ms;
                ctx->k_launches[p.id] += p.launches;
                ctx->k_bytes[p.id] += p.bytes * (double)p.launches;
                ctx->k_flops[p.id] += p.flops * (double)p.launches;
            } else {
                cudaGetLastError();
            }
            h->ev_pool.push_back(p.a);
            h->ev_pool.push_back(p.b);
        }
        h->pending.clear();
    }
    for (int i = 0; i < KC_COUNT; i++) {
        out[i].total_ms = ctx->k_ms[i];
        out[i].overlapped_ms = ctx->k_ms_helper[i];
        out[i].launches = ctx->k_launches[i];
        out[i].bytes = ctx->k_bytes[i];
        out[i].flops = ctx->k_flops[i];
        if (reset) {
            ctx->k_ms_helper[i] = 0;
            ctx->k_ms[i] = 0;
            ctx->k_launches[i] = 0;
            ctx->k_bytes[i] = 0;
            ctx->k_flops[i] = 0;
        }
    }
    return PCR_OK;
}

}  // extern "C"

// ---- layout helpers -----------------------------------------------------------------------------------
template <typename T>
__global__ void k_pack_xyz(const T *__restrict__ xyz, int n, float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4((float)xyz[3 * i], (float)xyz[3 * i + 1], (float)xyz[3 * i + 2], 0.0f);
}
__global__ void k_unpack_xyz(const float4 *__restrict__ in, int n, float *__restrict__ xyz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 p = in[i];
        xyz[3 * i] = p.x;
        xyz[3 * i + 1] = p.y;
        xyz[3 * i + 2] = p.z;
    }
}

struct Mat12 { double m[12]; };
__global__ void k_transform_points(const float4 *__restrict__ in, int n, Mat12 T, float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    const float3 q = xform_pt(T.m, p.x, p.y, p.z);
    out[i] = make_float4(q.x, q.y, q.z, p.w);
}

int pcr_pack_impl(pcr_ctx *ctx, const float *xyz, int n, float4 *out) {
    if (n <= 0) return PCR_OK;
    k_pack_xyz<float><<<div_up(n, 256), 256, 0, ctx->stream>>>(xyz, n, out);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

extern "C" {
int pcr_pack_xyz_f32(pcr_ctx *ctx, const float *xyz, int n, float *xyzw) {
    if (!ctx || n < 0) return PCR_ERR_INVALID;
    if (n == 0) return PCR_OK;
    k_pack_xyz<float><<<div_up(n, 256), 256, 0, ctx->stream>>>(xyz, n, (float4 *)xyzw);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
int pcr_pack_xyz_f64(pcr_ctx *ctx, const double *xyz, int n, float *xyzw) {
    if (!ctx || n < 0) return PCR_ERR_INVALID;
    if (n == 0) return PCR_OK;
    k_pack_xyz<double><<<div_up(n, 256), 256, 0, ctx->stream>>>(xyz, n, (float4 *)xyzw);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
int pcr_unpack_xyz_f32(pcr_ctx *ctx, const float *xyzw, int n, float *xyz) {
    if (!ctx || n < 0) return PCR_ERR_INVALID;
    if (n == 0) return PCR_OK;
    k_unpack_xyz<<<div_up(n, 256), 256, 0, ctx->stream>>>((const float4 *)xyzw, n, xyz);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
int pcr_transform_points(pcr_ctx *ctx, const float *xyzw, int n, const double *T, float *out) {
    if (!ctx || n < 0 || !T) return PCR_ERR_INVALID;
    if (n == 0) return PCR_OK;
    Mat12 m;
    for (int i = 0; i < 12; i++) m.m[i] = T[i];
    k_transform_points<<<div_up(n, 256), 256, 0, ctx->stream>>>((const float4 *)xyzw, n, m, (float4 *)out);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}
}
