// pcr_match_tc.cu — descriptor matching on the 5th-generation tensor cores (K6, tcgen05 + TMEM).
//
// Replaces the 33-D 1-NN search of open3d CorrespondencesFromFeatures (src/matcher/ransac.py:42-47, :85;
// SURVEY.md A.5) — the one dense contraction on the path: ||a - b||^2 = ||a||^2 + ||b||^2 - 2 a.b.
//
//   k_feat_prep       fp32 descriptors -> split-bf16 operands (a = hi + lo) laid out in the UMMA canonical
//                     K-major no-swizzle layout, one contiguous block per tile, + ||b||^2 per row.
//                     Query role columns [hi | lo | hi], base role [hi | hi | lo]  =>  A.B^T = hi.hi + lo.hi + hi.lo
//                     (K = 99 padded to 112 = 7 MMAs of K = 16).
//   k_match_tc<0>     GEMM pass 1: per query row and column slice, the column with the smallest approximate score
//                     s~ = ||b||^2 - 2 a.b (branch-free min-tree, the column's position folded into the low mantissa bits).
//   k_match_pick      exact distance d* (rule D9) of the best of those columns — an UPPER bound of the row's minimum —
//                     and the row threshold thr = d* - ||a||^2 + eps.
//   k_match_tc<1>     GEMM pass 2: every column with s~ <= thr is appended to the row's candidate list.
//   k_match_final     exact distances of the candidates, minimum by (distance, index); rows whose list overflowed go to
//                     the exact scan k_match_fallback.
// Exactness.  Let e_j = |s~_j - s_j| be the error of the approximate score of column j against the true
// s_j = ||b_j||^2 - 2 a.b_j.  A true minimiser j0 has ||a - b_j0||^2 <= d*, i.e. s_j0 <= d* - ||a||^2, hence
// s~_j0 <= d* - ||a||^2 + e_j0 <= thr as soon as eps >= max_j e_j: EVERY minimiser (ties included) is a candidate, and
// the final choice among the candidates is made with exact arithmetic — the result equals k_nn_features_exact bit for
// bit, whatever pass 1 returned (any column gives a valid upper bound).  Error bound: x = hi + lo' + r with
// |r| <= 2^-17 |x| (hi = bf16(x): |x - hi| <= 2^-9 |x|; lo' = bf16(x - hi): error <= 2^-9 * 2^-9 |x| / 2 ... <= 2^-17 |x|),
// so a.b - (hi.hi + lo.hi + hi.lo) = lo_a.lo_b + cross terms in r: |.| <= (2^-18 + 2 * 2^-17 + ...) sum |a_k b_k|
// <= 1.2e-5 ||a|| ||b||; fp32 accumulation of the 112 products in the tensor core: <= 112 * 2^-23 sum |terms| <=
// 1.4e-5 ||a|| ||b||; ||b||^2 in fp32 FMA: <= 33 * 2^-24 ||b||^2; the final FMA: 2^-24 |s~|.  Doubled for the factor 2:
// e_j <= 5.2e-5 ||a|| ||b_j|| + 2e-6 ||b_j||^2 + 6e-8 (||b_j||^2 + 2 ||a|| ||b_j||).  eps = 1.5e-4 ||a|| max||b|| +
// 1e-6 (max||b||^2 + ||a||^2) keeps a factor ~3 over that; tests/test_gpu_parity.py attacks it with large-norm
// descriptors that differ in the last bit.
#include <cuda_bf16.h>
#include <stdio.h>

#include "pcr_common.cuh"

constexpr int TC_K = 112;           // 3 x 33 = 99, padded to a multiple of 16
constexpr int TC_KCH = TC_K / 8;    // 16-byte K chunks
constexpr int TC_M = 128;           // query rows per tile (UMMA M)
constexpr int TC_N = 256;           // base rows per tile (UMMA N)
constexpr int TC_A_BYTES = TC_M * TC_K * 2;   // 28672
constexpr int TC_B_BYTES = TC_N * TC_K * 2;   // 57344
constexpr int TC_CAP = 16;  // candidates kept per (row, column slice) in pass 2; more -> exact fallback for that row

// ---- operand preparation ---------------------------------------------------------------------------------------
// tile layout (rows_per_tile = R): byte offset(r, c) = (c / 8) * (R * 16) + r * 16 + (c % 8) * 2
__global__ void __launch_bounds__(128) k_feat_prep(const float *__restrict__ f, int n, int rows_per_tile, int role,
                                                   __nv_bfloat16 *__restrict__ out, float *__restrict__ nrm2,
                                                   unsigned int *__restrict__ max_nrm2_bits) {
    // one thread per (row, 16-byte K chunk); rows are the fast index so the 16-byte stores of a warp are contiguous
    const int n_pad = ((n + rows_per_tile - 1) / rows_per_tile) * rows_per_tile;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_pad * TC_KCH) return;
    const int ch = (int)(gid / n_pad), r = (int)(gid - (long long)ch * n_pad);
    const int tile = r / rows_per_tile, rr = r - tile * rows_per_tile;
    char *base = (char *)out + (size_t)tile * rows_per_tile * TC_K * 2;
    __align__(16) __nv_bfloat16 v8[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int c = ch * 8 + e;
        float v = 0.0f;
        if (r < n && c < 99) {
            const int seg = c / 33, j = c - seg * 33;
            const float x = __ldg(f + (size_t)r * 33 + j);
            const __nv_bfloat16 hi = __float2bfloat16_rn(x);
            const float lo = x - __bfloat162float(hi);
            // query role: [hi | lo | hi]; base role: [hi | hi | lo]
            const bool want_lo = (role == 0) ? (seg == 1) : (seg == 2);
            v = want_lo ? lo : __bfloat162float(hi);
        }
        v8[e] = __float2bfloat16_rn(v);
    }
    *(uint4 *)(base + (size_t)ch * rows_per_tile * 16 + (size_t)rr * 16) = *(const uint4 *)v8;
    if (nrm2 && ch == TC_KCH - 1) {  // the all-padding chunk's thread also produces ||b||^2
        float acc = 0.0f;
        if (r < n)
            for (int j = 0; j < 33; j++) {
                const float x = __ldg(f + (size_t)r * 33 + j);
                acc = fmaf(x, x, acc);
            }
        nrm2[r] = (r < n) ? acc : INFINITY;  // padded base rows can never be selected
        if (r < n && max_nrm2_bits) atomicMax(max_nrm2_bits, __float_as_uint(acc));
    }
}

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error must surface as a launch failure (trap -> cudaErrorLaunchFailure), never as a hung GPU.
// 2^26 polls are seconds; the longest legitimate wait in this kernel is a few microseconds.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 26)) {
            printf("[pcr] k_match_tc: mbarrier wait timed out (block %d thread %d barrier +%u parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, smem_u32(bar) & 0xffu, parity);
            __trap();
        }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance between core matrices adjacent in K, SBO = between core matrices adjacent in M/N.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;  // descriptor version (sm_100)
    return d;         // base offset 0, layout type 0 = SWIZZLE_NONE
}

// instruction descriptor: D = F32, A = B = BF16, both K-major, N = 256, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- the tensor-core kernel ---------------------------------------------------------------------------------------
// Warp-specialised and pipelined (round 2; the round-1 kernel issued MMA -> commit -> all warps wait -> epilogue ->
// __syncthreads per tile and ran the tensor pipe at 9 %):
//   warp 8 (one elected lane) = producer: brings the base tiles in with cp.async.bulk through a 3-stage shared-memory
//       ring (+ a 4-slot ring for the ||b||^2 vectors) and issues the 7 tcgen05.mma of a tile into one of TWO TMEM
//       accumulators (2 x 256 columns = all of TMEM);
//   warps 0..7 = epilogue: tcgen05.ld of their lane quadrant / column half, score, min-tree, rare top-4 insertion.
// MMA(t + 1) runs while the epilogue of tile t reads the other accumulator; no CTA-wide barrier inside the tile loop.
// mbarriers (all single-phase-bit, tracked by tile counters):
//   b_full[s]     TMA bytes of stage s landed                      (producer waits)
//   nb_full[n]    ||b||^2 vector of slot n landed                  (epilogue waits)
//   b_free[s]     the MMAs reading stage s completed               (tcgen05.commit; producer waits before reloading)
//   tmem_full[a]  accumulator a holds tile t                       (tcgen05.commit; epilogue waits)
//   tmem_empty[a] all 8 epilogue warps are done with accumulator a (producer waits before overwriting)
//   a_full / a_free  query tile of the item loaded / no longer read
// No waiter can see a phase bit flip twice (the failure mode of parity waits; it DID happen in an earlier version whose
// epilogue also waited on b_full: for the first two tiles the producer's tmem_empty waits pass trivially, so stage 0 could
// be filled by tile 0 AND tile 3 before an epilogue warp delayed by a co-resident kernel had looked at it once):
//   b_full[s], b_free[s]: waited on only by the producer, which causes the next flip itself, later in program order;
//   tmem_full[a] of tile t flips again with MMA(t + 2), issued after tmem_empty[a] of tile t = all 8 warps past their wait;
//   nb_full[n] of tile t flips again with the load of tile t + 4, issued in producer iteration t + 2 after the tmem_empty
//   wait of tile t + 2 — a real wait for every t >= 0 — i.e. after every warp finished the epilogue of tile t, its last reader
//   (which is also why slot t % 4 may be overwritten then).
constexpr int TC_STAGES = 3;
constexpr int TC_NB_SLOTS = 4;
constexpr int TC_THREADS = 288;  // 8 epilogue warps + 1 producer warp

struct TcSmem {
    uint64_t a_full, a_free, b_full[TC_STAGES], b_free[TC_STAGES], tmem_full[2], tmem_empty[2], nb_full[TC_NB_SLOTS];
    uint32_t tmem_base;
    uint32_t pad[3];
    float nb[TC_NB_SLOTS][TC_N];
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MODE 0: pass 1 (best column per row and slice -> best_col).  MODE 1: pass 2 (columns with score <= thr[row] -> cand / cand_cnt).
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_match_tc(const __nv_bfloat16 *__restrict__ a_tiles, int n_a_tiles, int nq,
                                                            const __nv_bfloat16 *__restrict__ b_tiles, const float *__restrict__ b_nrm2,
                                                            int n_b_tiles, int n_split, int *__restrict__ best_col,
                                                            const float *__restrict__ thr_arr, int *__restrict__ cand,
                                                            int *__restrict__ cand_cnt) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [A tile][B stage 0..2][TcSmem]
    unsigned char *sA = smem_raw;
    unsigned char *sB = smem_raw + TC_A_BYTES;
    TcSmem *S = (TcSmem *)(sB + (size_t)TC_STAGES * TC_B_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S->tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(&S->a_full, 1);
        mbar_init(&S->a_free, 1);
        for (int i = 0; i < TC_STAGES; i++) {
            mbar_init(&S->b_full[i], 1);
            mbar_init(&S->b_free[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&S->tmem_full[i], 1);
            mbar_init(&S->tmem_empty[i], 8);
        }
        for (int i = 0; i < TC_NB_SLOTS; i++) mbar_init(&S->nb_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;

    // chunk-major tiles: K-adjacent core matrices are R*16 bytes apart (LBO), M/N-adjacent ones 128 bytes (SBO)
    const uint32_t lbo_a = (uint32_t)TC_M * 16u, sbo_a = 128u;
    const uint32_t lbo_b = (uint32_t)TC_N * 16u, sbo_b = 128u;

    if (warp == 8) {
        if (lane == 0) {
            uint32_t g_load = 0, g_mma = 0, n_items = 0;
            for (int item = blockIdx.x; item < n_a_tiles * n_split; item += gridDim.x) {
                const int at = item / n_split, sp = item - at * n_split;
                const int t0 = (int)((long long)n_b_tiles * sp / n_split), t1 = (int)((long long)n_b_tiles * (sp + 1) / n_split);
                const int nt = t1 - t0;
                if (nt <= 0) continue;
                // the previous item's MMAs no longer read sA (first use: parity 1 passes at once)
                mbar_wait(&S->a_free, (n_items & 1u) ^ 1u);
                mbar_expect_tx(&S->a_full, TC_A_BYTES);
                bulk_g2s(sA, (const char *)a_tiles + (size_t)at * TC_A_BYTES, TC_A_BYTES, &S->a_full);
                int loaded = 0;
                auto load_tile = [&](int t) {
                    const uint32_t st = g_load % TC_STAGES, ph = (g_load / TC_STAGES) & 1u;
                    mbar_wait(&S->b_free[st], ph ^ 1u);
                    mbar_expect_tx(&S->b_full[st], TC_B_BYTES);
                    bulk_g2s(sB + (size_t)st * TC_B_BYTES, (const char *)b_tiles + (size_t)t * TC_B_BYTES, TC_B_BYTES, &S->b_full[st]);
                    mbar_expect_tx(&S->nb_full[g_load % TC_NB_SLOTS], TC_N * 4);
                    bulk_g2s(S->nb[g_load % TC_NB_SLOTS], b_nrm2 + (size_t)t * TC_N, TC_N * 4, &S->nb_full[g_load % TC_NB_SLOTS]);
                    g_load++;
                };
                // two tiles ahead at the start of an item (their stages / slots were released by the previous item's tail)
                while (loaded < nt && loaded < 2) load_tile(t0 + loaded++);
                mbar_wait(&S->a_full, n_items & 1u);
                for (int k = 0; k < nt; k++) {
                    const uint32_t st = g_mma % TC_STAGES, ph = (g_mma / TC_STAGES) & 1u;
                    const uint32_t acc = g_mma & 1u, pha = (g_mma >> 1) & 1u;
                    mbar_wait(&S->b_full[st], ph);
                    mbar_wait(&S->tmem_empty[acc], pha ^ 1u);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + (size_t)st * TC_B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TC_K / 16; kk++) {
                        // one MMA consumes K = 16 = two 16-byte K chunks
                        umma_bf16(tmem + acc * (uint32_t)TC_N, make_smem_desc(a0 + 2u * kk * lbo_a, lbo_a, sbo_a),
                                  make_smem_desc(b0 + 2u * kk * lbo_b, lbo_b, sbo_b), TC_IDESC, kk > 0 ? 1u : 0u);
                    }
                    umma_commit(&S->tmem_full[acc]);
                    umma_commit(&S->b_free[st]);
                    if (k == nt - 1) umma_commit(&S->a_free);
                    g_mma++;
                    // tile k + 2 goes into the stage of tile k - 1 while MMA(k) runs
                    if (loaded < nt) load_tile(t0 + loaded++);
                }
                n_items++;
            }
        }
        __syncwarp();
    } else {
        uint32_t g = 0;
        for (int item = blockIdx.x; item < n_a_tiles * n_split; item += gridDim.x) {
            const int at = item / n_split, sp = item - at * n_split;
            const int t0 = (int)((long long)n_b_tiles * sp / n_split), t1 = (int)((long long)n_b_tiles * (sp + 1) / n_split);
            if (t1 <= t0) continue;
            const int row = at * TC_M + (warp & 3) * 32 + lane;
            const int list = sp * 2 + (warp >> 2);  // two lists (column halves of a tile) per base slice
            float best = INFINITY;                  // MODE 0: smallest packed score, and the chunk it came from
            int bchunk = -1;
            float thr = -INFINITY;                  // MODE 1
            int cnt = 0;
            int *out = nullptr;
            if (MODE == 1 && row < nq) {
                thr = __ldg(thr_arr + row);
                out = cand + ((size_t)list * nq + row) * TC_CAP;
            }
            for (int t = t0; t < t1; t++) {
                const uint32_t acc = g & 1u, pha = (g >> 1) & 1u;
                mbar_wait(&S->nb_full[g % TC_NB_SLOTS], (g / TC_NB_SLOTS) & 1u);
                mbar_wait(&S->tmem_full[acc], pha);
                tc_fence_after();
                // warp w reads TMEM lane quadrant w % 4 (query row 32*(w%4) + lane) and the column half w / 4 of the tile
                // (4 chunks of 32 columns)
                const float *nb = S->nb[g % TC_NB_SLOTS];
#pragma unroll 1
                for (int cch = (warp >> 2) * 4; cch < (warp >> 2) * 4 + 4; cch++) {
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + acc * (uint32_t)TC_N + (uint32_t)(cch * 32), v);
                    float sc[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) sc[j] = fmaf(-2.0f, __uint_as_float(v[j]), nb[cch * 32 + j]);
                    if (MODE == 0) {
                        // position inside the chunk in the 5 low mantissa bits (a relative change of 2^-18: pass 1 only
                        // has to return SOME good column), then one branch-free min-tree: ~3 instructions per score.
                        // A padded base row scores +inf, whose packed form is a NaN: fminf drops it.
#pragma unroll
                        for (int j = 0; j < 32; j++) sc[j] = __uint_as_float((__float_as_uint(sc[j]) & 0xffffffe0u) | (uint32_t)j);
                    }
                    float m16[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) m16[j] = fminf(sc[j], sc[j + 16]);
#pragma unroll
                    for (int j = 0; j < 8; j++) m16[j] = fminf(m16[j], m16[j + 8]);
#pragma unroll
                    for (int j = 0; j < 4; j++) m16[j] = fminf(m16[j], m16[j + 4]);
                    const float cmin = fminf(fminf(m16[0], m16[1]), fminf(m16[2], m16[3]));
                    if (MODE == 0) {
                        const bool lt = cmin < best;
                        best = lt ? cmin : best;
                        bchunk = lt ? t * (TC_N / 32) + cch : bchunk;
                    } else if (cmin <= thr) {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (sc[j] <= thr) {
                                if (cnt < TC_CAP) out[cnt] = t * TC_N + cch * 32 + j;
                                cnt++;
                            }
                    }
                }
                tc_fence_before();
                g++;
                __syncwarp();
                if (lane == 0) mbar_arrive(&S->tmem_empty[acc]);
            }
            if (row < nq) {
                if (MODE == 0) best_col[(size_t)list * nq + row] = bchunk >= 0 ? bchunk * 32 + (int)(__float_as_uint(best) & 31u) : -1;
                else cand_cnt[(size_t)list * nq + row] = cnt;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- exact re-check -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double exact_dist(const float *__restrict__ a, const float *__restrict__ b) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 33; k++) {
        const double df = (double)a[k] - (double)b[k];
        acc = acc + df * df;
    }
    return acc;
}

// after pass 1: exact distance of the best provisional column = an upper bound d* of the row's minimum, and the
// pass-2 threshold thr = d* - ||a||^2 + eps, rounded up (see the header).  One thread per row.
__global__ void __launch_bounds__(128) k_match_pick(const float *__restrict__ fq, int nq, const float *__restrict__ fb, int nb,
                                                    const int *__restrict__ best_col, int n_lists,
                                                    const unsigned int *__restrict__ max_nrm2_bits, float *__restrict__ thr) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    float a[33];
    double na = 0.0;
#pragma unroll
    for (int k = 0; k < 33; k++) {
        a[k] = __ldg(fq + (size_t)q * 33 + k);
        na += (double)a[k] * (double)a[k];
    }
    double best = INFINITY;
    for (int l = 0; l < n_lists; l++) {
        const int j = best_col[(size_t)l * nq + q];
        if (j < 0 || j >= nb) continue;
        const double d = exact_dist(a, fb + (size_t)j * 33);
        if (d < best) best = d;
    }
    const double bmax = sqrt((double)__uint_as_float(*max_nrm2_bits));
    const double eps = 1.5e-4 * sqrt(na) * bmax + 1e-6 * (bmax * bmax + na) + 1e-30;
    // no finite upper bound (NaN descriptor, or no column at all): thr = -inf, no candidates, the row ends with nn = -1
    thr[q] = (best < INFINITY && na == na) ? __double2float_ru((best - na) + eps) : -INFINITY;
}

// after pass 2: one warp per query row scores its candidates exactly; minimum by (distance, index).  A row with an
// overflowed list is scanned exactly by k_match_fallback instead.
__global__ void __launch_bounds__(128) k_match_final(const float *__restrict__ fq, int nq, const float *__restrict__ fb, int nb,
                                                     const int *__restrict__ cand, const int *__restrict__ cand_cnt, int n_lists,
                                                     const float *__restrict__ thr, int *__restrict__ nn,
                                                     int *__restrict__ fallback_rows, unsigned int *__restrict__ n_fallback) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    float a[33];
#pragma unroll
    for (int k = 0; k < 33; k++) a[k] = __ldg(fq + (size_t)q * 33 + k);
    double best = INFINITY;
    int bi = 0x7fffffff;
    bool over = false;
    for (int l = 0; l < n_lists; l++) {
        const int c = cand_cnt[(size_t)l * nq + q];
        over = over || c > TC_CAP;
        for (int k = lane; k < min(c, TC_CAP); k += 32) {
            const int j = cand[((size_t)l * nq + q) * TC_CAP + k];
            if (j < 0 || j >= nb) continue;
            const double d = exact_dist(a, fb + (size_t)j * 33);
            if (d < best || (d == best && j < bi)) { best = d; bi = j; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
    if (lane != 0) return;
    if (over) {
        nn[q] = -2;
        fallback_rows[atomicAdd(n_fallback, 1u)] = q;
    } else {
        // bi stays unset only when the row has no finite distance at all (thr = -inf): no neighbour, as the exact kernel
        nn[q] = bi == 0x7fffffff ? -1 : bi;
    }
}

// descriptors [n][33] -> [33][n]: the fallback scan then reads 32 consecutive candidates per load instruction (the
// row-major layout costs 32 L1 tag look-ups per load and made the scan L1-bound: 142 us for 3.5 % of the rows)
__global__ void __launch_bounds__(256) k_feat_transpose(const float *__restrict__ f, int n, float *__restrict__ ft) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * 33) return;
    const int j = (int)(i / 33), k = (int)(i % 33);
    ft[(size_t)k * n + j] = __ldg(f + i);
}

// Exact scan for the rows that failed the certificate (device-side count, no host sync).
// Work item = (row, part): the candidates are cut into FB_PARTS contiguous ranges so that a few hundred rows still fill
// the machine; the query descriptor is read from shared memory (broadcast) so that the kernel runs at 40 registers.
//
// fp32 filter: sum (a_k - b_k)^2 in fp32 has a RELATIVE error below 36 * 2^-24 = 2.2e-6 (no cancellation in this form),
// so the exact minimiser j* satisfies d32(j*) <= (1 + 7e-6) * d32(j) for every j.  Pass 1 finds the minimum of d32 over
// the part (coalesced loads from the transposed descriptors, fp32 only); pass 2 re-scores in fp64 (rule D9) only the
// candidates within 1e-4 of it — a handful — so the fp64 dependency chains, which made the all-fp64 scan latency-bound
// (142 us for 3.5 % of the rows), all but disappear.  The part minimum is >= the row minimum, so the filter of a part
// keeps every candidate the row-wide filter would keep; the part that finishes last merges the FB_PARTS exact results
// with the (distance, index) tie rule.  (A per-thread running minimum does not work: among the 32 lanes of a warp some
// lane sets a new record in nearly every step, and the warp pays the fp64 path every time.)
// Pass 1 parks its fp32 distances in shared memory, so pass 2 reads no descriptors at all; the candidates of a row are
// cut into >= 8 parts so that a few hundred rows give a few thousand work items.
constexpr int FB_THREADS = 128;
constexpr int FB_ROWS = 1;  // rows sharing one descriptor sweep (4 was measured slower: too few work items, LDS-heavy)
constexpr int FB_PER_MAX = 2560;  // candidates per part: FB_ROWS * FB_PER_MAX floats of dynamic shared memory (40 KB)
struct FbPart {
    double d;
    int j;
    int pad;
};

__global__ void __launch_bounds__(FB_THREADS) k_match_fallback(const float *__restrict__ fq, const float *__restrict__ fb,
                                                               const float *__restrict__ fbT, int nb, int nparts, int per,
                                                               const int *__restrict__ rows, const unsigned int *__restrict__ n_rows,
                                                               FbPart *__restrict__ parts, unsigned int *__restrict__ tickets,
                                                               int *__restrict__ nn) {
    extern __shared__ float sd32[];  // [FB_ROWS][per]
    __shared__ float sa[FB_ROWS][33];
    __shared__ double sd[FB_THREADS / 32];
    __shared__ int si[FB_THREADS / 32];
    __shared__ float sf[FB_ROWS][FB_THREADS / 32];
    const unsigned int n = *n_rows;
    const unsigned int n_groups = (n + FB_ROWS - 1) / FB_ROWS;
    const unsigned int n_items = n_groups * (unsigned int)nparts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const unsigned int grp = item / (unsigned int)nparts;
        const int part = (int)(item % (unsigned int)nparts);
        const int j0 = part * per, j1 = min(nb, j0 + per);
        __syncthreads();
        for (int t = threadIdx.x; t < FB_ROWS * 33; t += FB_THREADS) {
            const unsigned int r = grp * FB_ROWS + t / 33;
            sa[t / 33][t % 33] = r < n ? __ldg(fq + (size_t)rows[r] * 33 + t % 33) : 0.0f;
        }
        __syncthreads();
        // pass 1: fp32 distances of FB_ROWS rows to every candidate of the part, one descriptor sweep
        float m32[FB_ROWS];
#pragma unroll
        for (int u = 0; u < FB_ROWS; u++) m32[u] = INFINITY;
        for (int j = j0 + threadIdx.x; j < j1; j += FB_THREADS) {
            float acc[FB_ROWS];
#pragma unroll
            for (int u = 0; u < FB_ROWS; u++) acc[u] = 0.0f;
#pragma unroll
            for (int k = 0; k < 33; k++) {
                const float bk = __ldg(fbT + (size_t)k * nb + j);
#pragma unroll
                for (int u = 0; u < FB_ROWS; u++) {
                    const float df = sa[u][k] - bk;
                    acc[u] = __fmaf_rn(df, df, acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < FB_ROWS; u++) {
                sd32[u * per + (j - j0)] = acc[u];
                m32[u] = fminf(m32[u], acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < FB_ROWS; u++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m32[u] = fminf(m32[u], __shfl_xor_sync(0xffffffffu, m32[u], o));
            if (lane == 0) sf[u][warp] = m32[u];
        }
        __syncthreads();
        // pass 2, row by row: exact fp64 re-score of the candidates within 1e-4 of the part minimum
        for (int u = 0; u < FB_ROWS; u++) {
            const unsigned int r = grp * FB_ROWS + u;
            if (r >= n) break;  // block-uniform
            float m = sf[u][0];
#pragma unroll
            for (int w = 1; w < FB_THREADS / 32; w++) m = fminf(m, sf[u][w]);
            const float lim = m * 1.0001f;
            double best = INFINITY;
            int bi = 0x7fffffff;
            for (int j = j0 + threadIdx.x; j < j1; j += FB_THREADS) {
                if (sd32[u * per + (j - j0)] <= lim) {
                    const double d = exact_dist(sa[u], fb + (size_t)j * 33);
                    if (d < best) { best = d; bi = j; }  // ascending j per thread: strict < keeps the lowest index
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
            }
            __syncthreads();
            if (lane == 0) { sd[warp] = best; si[warp] = bi; }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < FB_THREADS / 32; w++)
                    if (sd[w] < best || (sd[w] == best && si[w] < bi)) { best = sd[w]; bi = si[w]; }
                FbPart *o = parts + (size_t)r * nparts + part;
                o->d = best;
                o->j = bi;
                __threadfence();
                const unsigned int t = atomicAdd(tickets + r, 1u);
                if (t == (unsigned int)nparts - 1u) {  // last part of this row: merge with the (distance, index) rule
                    __threadfence();
                    const volatile FbPart *v = parts + (size_t)r * nparts;
                    double bd = v[0].d;
                    int bj = v[0].j;
                    for (int p2 = 1; p2 < nparts; p2++) {
                        const double d2 = v[p2].d;
                        const int jj = v[p2].j;
                        if (d2 < bd || (d2 == bd && jj < bj)) { bd = d2; bj = jj; }
                    }
                    nn[rows[r]] = bj == 0x7fffffff ? -1 : bj;  // no finite distance (NaN descriptor): no neighbour, as the exact kernel
                }
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
struct TcOperand {
    __nv_bfloat16 *tiles;
    float *nrm2;
    unsigned int *max_bits;
    int n, n_tiles;
};

static int tc_prep(pcr_ctx *ctx, const float *f, int n, int role, TcOperand *op) {
    const int rpt = role == 0 ? TC_M : TC_N;
    op->n = n;
    op->n_tiles = div_up(n, rpt);
    const size_t n_pad = (size_t)op->n_tiles * rpt;
    op->tiles = arena<__nv_bfloat16>(ctx, n_pad * TC_K);
    op->nrm2 = role == 1 ? arena<float>(ctx, n_pad) : nullptr;
    op->max_bits = role == 1 ? arena<unsigned int>(ctx, 4) : nullptr;
    if (!op->tiles || (role == 1 && (!op->nrm2 || !op->max_bits))) return PCR_ERR_OOM;
    if (role == 1) PCR_CUDA(cudaMemsetAsync(op->max_bits, 0, 16, ctx->stream));
    k_feat_prep<<<div_up((long long)n_pad * TC_KCH, 128), 128, 0, ctx->stream>>>(f, n, rpt, role, op->tiles, op->nrm2, op->max_bits);
    PCR_LAUNCHED();
    PCR_CUDA(cudaGetLastError());
    return PCR_OK;
}

int pcr_nn_features_tc_impl(pcr_ctx *ctx, const float *fq, int nq, const float *fb, int nb, int *nn) {
    if (nq == 0) return PCR_OK;
    if (nb == 0) {
        PCR_CUDA(cudaMemsetAsync(nn, 0xff, sizeof(int) * (size_t)nq, ctx->stream));
        return PCR_OK;
    }
    TcOperand A, B;
    {
        KScope ks(ctx, KC_MATCH_MISC, 132.0 * ((double)nq + nb) + 224.0 * ((double)nq + nb), 2);
        PCR_TRY(tc_prep(ctx, fq, nq, 0, &A));
        PCR_TRY(tc_prep(ctx, fb, nb, 1, &B));
    }
    PCR_ALLOC(fbT, float, (size_t)nb * 33);
    int fb_nparts = div_up(nb, FB_PER_MAX);
    if (fb_nparts < 8) fb_nparts = 8;
    const int fb_per = div_up(nb, fb_nparts);
    PCR_ALLOC(fb_parts, FbPart, (size_t)nq * fb_nparts);
    PCR_ALLOC(fb_tickets, unsigned int, (size_t)nq);
    PCR_CUDA(cudaMemsetAsync(fb_tickets, 0, sizeof(unsigned int) * (size_t)nq, ctx->stream));
    {
        KScope ks(ctx, KC_MATCH_MISC, 264.0 * nb);
        k_feat_transpose<<<div_up((long long)nb * 33, 256), 256, 0, ctx->stream>>>(fb, nb, fbT);
        PCR_LAUNCHED();
    }
    // slices of the base rows so that the grid covers the machine
    int n_split = 1;
    while (A.n_tiles * n_split * 2 <= ctx->sm_count && n_split * 2 <= B.n_tiles && n_split < 4) n_split *= 2;
    const int n_lists = 2 * n_split;  // each CTA keeps two lists per row (column halves of a tile)
    PCR_ALLOC(best_col, int, (size_t)n_lists * nq);
    PCR_ALLOC(thr, float, (size_t)nq);
    PCR_ALLOC(cand, int, (size_t)n_lists * nq * TC_CAP);
    PCR_ALLOC(cand_cnt, int, (size_t)n_lists * nq);
    PCR_ALLOC(fb_rows, int, (size_t)nq);
    PCR_ALLOC(n_fb, unsigned int, 4);
    PCR_CUDA(cudaMemsetAsync(n_fb, 0, 16, ctx->stream));
    const size_t smem = (size_t)TC_A_BYTES + (size_t)TC_STAGES * TC_B_BYTES + sizeof(TcSmem) + 1024;
    if (!ctx->match_tc_attr_set) {  // per device: set once per context
        PCR_CUDA(cudaFuncSetAttribute(k_match_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PCR_CUDA(cudaFuncSetAttribute(k_match_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->match_tc_attr_set = true;
    }
    const int items = A.n_tiles * n_split;
    const double gemm_flops = 2.0 * TC_K * (double)A.n_tiles * TC_M * (double)B.n_tiles * TC_N;
    const double gemm_bytes = 224.0 * ((double)nq + (double)nb * A.n_tiles);
    {
        KScope ks(ctx, KC_NN_FEATURES, gemm_bytes + 4.0 * nq * n_lists, 1, gemm_flops);
        k_match_tc<0><<<min(items, ctx->sm_count), TC_THREADS, smem, ctx->stream>>>(A.tiles, A.n_tiles, nq, B.tiles, B.nrm2, B.n_tiles, n_split,
                                                                             best_col, nullptr, nullptr, nullptr);
        PCR_LAUNCHED();
    }
    {
        KScope ks(ctx, KC_MATCH_MISC, 132.0 * nq * (1 + n_lists));
        k_match_pick<<<div_up(nq, 128), 128, 0, ctx->stream>>>(fq, nq, fb, nb, best_col, n_lists, B.max_bits, thr);
        PCR_LAUNCHED();
    }
    {
        KScope ks(ctx, KC_NN_FEATURES, gemm_bytes + 8.0 * nq * n_lists, 1, gemm_flops);
        k_match_tc<1><<<min(items, ctx->sm_count), TC_THREADS, smem, ctx->stream>>>(A.tiles, A.n_tiles, nq, B.tiles, B.nrm2, B.n_tiles, n_split,
                                                                             nullptr, thr, cand, cand_cnt);
        PCR_LAUNCHED();
    }
    {
        KScope ks(ctx, KC_MATCH_MISC, 132.0 * nq * 4, 2);
        k_match_final<<<div_up((long long)nq * 32, 128), 128, 0, ctx->stream>>>(fq, nq, fb, nb, cand, cand_cnt, n_lists, thr, nn, fb_rows, n_fb);
        PCR_LAUNCHED();
        k_match_fallback<<<ctx->sm_count * 12, FB_THREADS, sizeof(float) * FB_ROWS * (size_t)fb_per, ctx->stream>>>(
            fq, fb, fbT, nb, fb_nparts, fb_per, fb_rows, n_fb, fb_parts, fb_tickets, nn);
        PCR_LAUNCHED();
    }
    PCR_CUDA(cudaGetLastError());
    if (getenv("PCR_DEBUG")) {  // bring-up aid: how many rows needed the exact fallback
        unsigned int h = 0;
        cudaMemcpyAsync(&h, n_fb, 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[pcr] match tc: nq=%d nb=%d n_split=%d fallback rows=%u\n", nq, nb, n_split, h);
    }
    return PCR_OK;
}
