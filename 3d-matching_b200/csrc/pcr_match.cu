// pcr_match.cu — FPFH feature matching (K6), replaces open3d CorrespondencesFromFeatures as reached from
// src/matcher/ransac.py:85 and from registration_ransac_based_on_feature_matching (src/matcher/ransac.py:42-47)
// (1-NN in 33-D per source descriptor, optional mutual filter with the 10 % fall-back).
//
// Distance specification: fp64 sequential accumulation over the fp32 descriptors, ties -> lowest index.
// k_nn_features_exact evaluates it for every pair (CUDA-core fp64, smem-staged tiles, broadcast reads).
#include "pcr_common.cuh"

constexpr int NNF_THREADS = 128;
constexpr int NNF_TILE = 64;

__global__ void __launch_bounds__(NNF_THREADS) k_nn_features_exact(const float *__restrict__ fq, int nq,
                                                                   const float *__restrict__ fb, int nb,
                                                                   int *__restrict__ nn) {
    __shared__ float tile[NNF_TILE][33];
    const int q = blockIdx.x * NNF_THREADS + threadIdx.x;
    float a[33];
#pragma unroll
    for (int k = 0; k < 33; k++) a[k] = (q < nq) ? fq[(size_t)q * 33 + k] : 0.0f;
    double best = INFINITY;
    int bi = -1;
    for (int base = 0; base < nb; base += NNF_TILE) {
        const int tn = min(NNF_TILE, nb - base);
        __syncthreads();
        for (int t = threadIdx.x; t < tn * 33; t += NNF_THREADS) tile[t / 33][t % 33] = fb[(size_t)base * 33 + t];
        __syncthreads();
        for (int j = 0; j < tn; j++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 33; k++) {
                const double df = (double)a[k] - (double)tile[j][k];
                acc = acc + df * df;
            }
            if (acc < best) {
                best = acc;
                bi = base + j;
            }
        }
    }
    if (q < nq) nn[q] = bi;
}

__global__ void __launch_bounds__(256) k_mutual_flags(const int *__restrict__ nn_s, const int *__restrict__ nn_t, int ms,
                                                      uint32_t *__restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ms) return;
    const int j = nn_s[i];
    flags[i] = (j >= 0 && nn_t[j] == i) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_corr_compact(const int *__restrict__ nn_s, const uint32_t *__restrict__ pos,
                                                      int ms, int *__restrict__ corr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ms) return;
    if (pos[i + 1] != pos[i]) {
        corr[2 * (size_t)pos[i]] = i;
        corr[2 * (size_t)pos[i] + 1] = nn_s[i];
    }
}

// one-directional set: every source descriptor with a nearest target descriptor (a descriptor holding a NaN has none,
// nn = -1, and contributes no pair — an index of -1 would be an out-of-bounds read in every consumer)
__global__ void __launch_bounds__(256) k_valid_flags(const int *__restrict__ nn_s, int ms, uint32_t *__restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ms) return;
    flags[i] = nn_s[i] >= 0 ? 1u : 0u;
}

int pcr_nn_features_tc_impl(pcr_ctx *ctx, const float *fq, int nq, const float *fb, int nb, int *nn);

// PCR_MATCH_EXACT=1 forces the CUDA-core exact kernel for every size (bring-up / A-B comparison aid)
static bool match_use_tensor_cores() {
    static const bool v = [] {  // initialised once, thread-safely
        const char *e = getenv("PCR_MATCH_EXACT");
        return !(e && e[0] == '1');
    }();
    return v;
}

int pcr_nn_features_impl(pcr_ctx *ctx, const float *fq, int nq, const float *fb, int nb, int *nn) {
    if (nq == 0) return PCR_OK;
    // tensor-core filter + exact certificate (pcr_match_tc.cu) once the problem fills at least a few tiles
    if (nb >= 512 && nq >= 128 && match_use_tensor_cores()) return pcr_nn_features_tc_impl(ctx, fq, nq, fb, nb, nn);
    if (nb == 0) {
        PCR_CUDA(cudaMemsetAsync(nn, 0xff, sizeof(int) * (size_t)nq, ctx->stream));
        return PCR_OK;
    }
    KScope ks(ctx, KC_NN_FEATURES, 132.0 * ((double)nq + nb) + 4.0 * nq, 1, 2.0 * 33.0 * (double)nq * (double)nb);
    k_nn_features_exact<<<div_up(nq, NNF_THREADS), NNF_THREADS, 0, ctx->stream>>>(fq, nq, fb, nb, nn);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

int pcr_match_impl(pcr_ctx *ctx, const float *fs, int ms, const float *ft, int mt, int mutual, double mutual_ratio,
                   int *corr, int *c_host) {
    *c_host = 0;
    if (ms == 0 || mt == 0) return PCR_OK;
    PCR_ALLOC(nn_s, int, (size_t)ms);
    PCR_ALLOC(nn_t, int, mutual ? (size_t)mt : 1);
    // The two directions of the mutual filter are independent chains of six launches each (operand preparation, two
    // tensor-core passes, pick, final choice, fallback): side by side on two streams when the context has them (created
    // with pcr_align's helper; PCR_MATCH_CONCURRENT=0: one after the other).  The tensor-core kernels still take turns —
    // a CTA holds all 512 TMEM columns of its SM — but each direction's small kernels run beside the other's GEMM.
    static const bool conc = !(getenv("PCR_MATCH_CONCURRENT") && atoi(getenv("PCR_MATCH_CONCURRENT")) == 0);
    cudaEvent_t t_done = nullptr;
    if (mutual && conc && ctx->aux2_stream && ctx->aux2_stream != ctx->stream) {
        cudaEvent_t in_ready;
        PCR_CUDA(cudaEventCreateWithFlags(&in_ready, cudaEventDisableTiming));
        PCR_CUDA(cudaEventRecord(in_ready, ctx->stream));
        PCR_CUDA(cudaStreamWaitEvent(ctx->aux2_stream, in_ready, 0));
        PCR_CUDA(cudaEventDestroy(in_ready));
        cudaStream_t keep = ctx->stream;
        ctx->stream = ctx->aux2_stream;
        const int rc2 = pcr_nn_features_impl(ctx, ft, mt, fs, ms, nn_t);
        ctx->stream = keep;
        PCR_TRY(rc2);
        PCR_CUDA(cudaEventCreateWithFlags(&t_done, cudaEventDisableTiming));
        PCR_CUDA(cudaEventRecord(t_done, ctx->aux2_stream));
    }
    const bool two_streams = t_done != nullptr;
    {
        const int rc1 = pcr_nn_features_impl(ctx, fs, ms, ft, mt, nn_s);
        if (t_done) {  // joined on every path: the other stream works on this call's arena
            cudaStreamWaitEvent(ctx->stream, t_done, 0);
            cudaEventDestroy(t_done);
        }
        PCR_TRY(rc1);
    }
    PCR_ALLOC(pos, uint32_t, (size_t)ms + 1);
    uint32_t *h = (uint32_t *)ctx->pinned;
    if (mutual) {
        if (!two_streams) PCR_TRY(pcr_nn_features_impl(ctx, ft, mt, fs, ms, nn_t));
        k_mutual_flags<<<div_up(ms, 256), 256, 0, ctx->stream>>>(nn_s, nn_t, ms, pos);
        PCR_LAUNCHED();
        PCR_TRY(pcr_exclusive_scan_u32(ctx, pos, ms));
        k_corr_compact<<<div_up(ms, 256), 256, 0, ctx->stream>>>(nn_s, pos, ms, corr);
        PCR_LAUNCHED();
        PCR_CUDA(cudaMemcpyAsync(h, pos + ms, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
        const int c = (int)*h;
        // Open3D: int(corres_mutual.size()) >= int(mutual_consistency_ratio * num_src) — both sides truncated (A.5)
        if (c >= (int)(mutual_ratio * (double)ms)) {
            *c_host = c;
            return PCR_OK;
        }
        // too few mutual pairs: fall back to the one-directional set (A.5)
    }
    k_valid_flags<<<div_up(ms, 256), 256, 0, ctx->stream>>>(nn_s, ms, pos);
    PCR_LAUNCHED();
    PCR_TRY(pcr_exclusive_scan_u32(ctx, pos, ms));
    k_corr_compact<<<div_up(ms, 256), 256, 0, ctx->stream>>>(nn_s, pos, ms, corr);
    PCR_LAUNCHED();
    PCR_CUDA(cudaMemcpyAsync(h, pos + ms, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PCR_CUDA(pcr_sync_stream(ctx, ctx->stream));
    *c_host = (int)*h;
    return PCR_OK;
}
