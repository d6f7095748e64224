// pcr_match_tc.cu — descriptor matching on the 5th-generation tensor cores (K6, tcgen05 + TMEM).
//
// Replaces the 33-D 1-NN search of open3d CorrespondencesFromFeatures (src/matcher/ransac.py:42-47, :85;
// SURVEY.md A.5) — the one dense contraction on the path: ||a - b||^2 = ||a||^2 + ||b||^2 - 2 a.b.
//
//   k_feat_prep     fp32 descriptors -> split-bf16 operands (a = hi + lo) laid out in the UMMA canonical
//                   K-major no-swizzle layout, one contiguous block per tile, + ||b||^2 per row.
//                   Query role columns [hi | lo | hi], base role [hi | hi | lo]  =>  A.B^T = hi.hi + lo.hi + hi.lo
//                   (K = 99 padded to 112 = 7 MMAs of K = 16); relative error of a.b about 2^-15.
//   k_match_tc      one CTA per (128-query tile, slice of the base rows): the operand blocks arrive by
//                   cp.async.bulk (TMA engine, mbarrier complete_tx), one elected thread issues
//                   tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, fp32 accumulators in TMEM), the four warps
//                   read their TMEM lane quadrant with tcgen05.ld and keep, per query row, the 4 smallest
//                   scores ||b||^2 - 2 a.b and the 4th score.  The distance matrix is never materialised.
//   k_match_recheck exact re-check (rule D9: fp64 sequential over the fp32 descriptors, ties -> lowest index) of
//                   the candidates; a row is accepted only if its best exact distance is below the proven lower
//                   bound of every non-candidate (4th score - error bound); other rows go to the exact kernel.
// The result is therefore bit-identical to k_nn_features_exact for every row.
#include <cuda_bf16.h>
#include <stdio.h>

#include "pcr_common.cuh"

constexpr int TC_K = 112;           // 3 x 33 = 99, padded to a multiple of 16
constexpr int TC_KCH = TC_K / 8;    // 16-byte K chunks
constexpr int TC_M = 128;           // query rows per tile (UMMA M)
constexpr int TC_N = 256;           // base rows per tile (UMMA N)
constexpr int TC_A_BYTES = TC_M * TC_K * 2;   // 28672
constexpr int TC_B_BYTES = TC_N * TC_K * 2;   // 57344
constexpr int TC_NCAND = 4;

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
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
struct TcSmem {
    uint64_t bar_a, bar_b[2], bar_mma;
    uint32_t tmem_base;
    uint32_t pad[3];
    float nb[2][TC_N];
};

__global__ void __launch_bounds__(256, 1) k_match_tc(const __nv_bfloat16 *__restrict__ a_tiles, int n_a_tiles, int nq,
                                                     const __nv_bfloat16 *__restrict__ b_tiles, const float *__restrict__ b_nrm2,
                                                     int n_b_tiles, int n_split,
                                                     int *__restrict__ cand_idx, float *__restrict__ cand_kth) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [A tile][B tile 0][B tile 1][TcSmem]
    unsigned char *sA = smem_raw;
    unsigned char *sB0 = smem_raw + TC_A_BYTES;
    unsigned char *sB1 = sB0 + TC_B_BYTES;
    TcSmem *S = (TcSmem *)(sB1 + TC_B_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S->tmem_base)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(&S->bar_a, 1);
        mbar_init(&S->bar_b[0], 1);
        mbar_init(&S->bar_b[1], 1);
        mbar_init(&S->bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S->tmem_base;
    uint32_t par_a = 0, par_b[2] = {0, 0}, par_mma = 0;

    // chunk-major tiles: K-adjacent core matrices are R*16 bytes apart (LBO), M/N-adjacent ones 128 bytes (SBO)
    const uint32_t lbo_a = (uint32_t)TC_M * 16u, sbo_a = 128u;
    const uint32_t lbo_b = (uint32_t)TC_N * 16u, sbo_b = 128u;

    for (int item = blockIdx.x; item < n_a_tiles * n_split; item += gridDim.x) {
        const int at = item / n_split, sp = item - at * n_split;
        const int t0 = (int)((long long)n_b_tiles * sp / n_split), t1 = (int)((long long)n_b_tiles * (sp + 1) / n_split);
        if (threadIdx.x == 0) {
            mbar_expect_tx(&S->bar_a, TC_A_BYTES);
            bulk_g2s(sA, (const char *)a_tiles + (size_t)at * TC_A_BYTES, TC_A_BYTES, &S->bar_a);
            if (t0 < t1) {
                mbar_expect_tx(&S->bar_b[0], TC_B_BYTES + TC_N * 4);
                bulk_g2s(sB0, (const char *)b_tiles + (size_t)t0 * TC_B_BYTES, TC_B_BYTES, &S->bar_b[0]);
                bulk_g2s(S->nb[0], b_nrm2 + (size_t)t0 * TC_N, TC_N * 4, &S->bar_b[0]);
            }
        }
        float s0 = INFINITY, s1 = INFINITY, s2 = INFINITY, s3 = INFINITY;
        int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
        mbar_wait(&S->bar_a, par_a);
        par_a ^= 1;
        for (int t = t0; t < t1; t++) {
            const int buf = (t - t0) & 1;
            unsigned char *sB = buf ? sB1 : sB0;
            if (threadIdx.x == 0) {
                if (t + 1 < t1) {  // prefetch the next base tile into the other buffer (its MMAs completed last iteration)
                    unsigned char *sBn = buf ? sB0 : sB1;
                    mbar_expect_tx(&S->bar_b[buf ^ 1], TC_B_BYTES + TC_N * 4);
                    bulk_g2s(sBn, (const char *)b_tiles + (size_t)(t + 1) * TC_B_BYTES, TC_B_BYTES, &S->bar_b[buf ^ 1]);
                    bulk_g2s(S->nb[buf ^ 1], b_nrm2 + (size_t)(t + 1) * TC_N, TC_N * 4, &S->bar_b[buf ^ 1]);
                }
            }
            mbar_wait(&S->bar_b[buf], par_b[buf]);
            par_b[buf] ^= 1;
            if (threadIdx.x == 0) {
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
                for (int k = 0; k < TC_K / 16; k++) {
                    // one MMA consumes K = 16 = two 16-byte K chunks
                    const uint32_t koff_a = 2u * k * lbo_a;
                    const uint32_t koff_b = 2u * k * lbo_b;
                    umma_bf16(tmem, make_smem_desc(a0 + koff_a, lbo_a, sbo_a), make_smem_desc(b0 + koff_b, lbo_b, sbo_b), TC_IDESC,
                              k > 0 ? 1u : 0u);
                }
                umma_commit(&S->bar_mma);
            }
            mbar_wait(&S->bar_mma, par_mma);
            par_mma ^= 1;
            tc_fence_after();
            // epilogue: 8 warps; warp w reads TMEM lane quadrant w % 4 (query row 32*(w%4) + lane) and the column half
            // w / 4 of the tile (4 chunks of 32 columns), keeping its own top-4 list
            const float *nb = S->nb[buf];
#pragma unroll 1
            for (int cch = (warp >> 2) * 4; cch < (warp >> 2) * 4 + 4; cch++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cch * 32), v);
                // scores of the 32 columns, then one min-tree: a chunk that cannot enter the top-4 costs ~2 instr/value
                float sc[32];
#pragma unroll
                for (int j = 0; j < 32; j++) sc[j] = fmaf(-2.0f, __uint_as_float(v[j]), nb[cch * 32 + j]);
                float m16[16];
#pragma unroll
                for (int j = 0; j < 16; j++) m16[j] = fminf(sc[j], sc[j + 16]);
#pragma unroll
                for (int j = 0; j < 8; j++) m16[j] = fminf(m16[j], m16[j + 8]);
#pragma unroll
                for (int j = 0; j < 4; j++) m16[j] = fminf(m16[j], m16[j + 4]);
                const float cmin = fminf(fminf(m16[0], m16[1]), fminf(m16[2], m16[3]));
                if (cmin < s3) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float s = sc[j];
                        if (s < s3) {
                            const int col = t * TC_N + cch * 32 + j;
                            if (s < s2) {
                                s3 = s2; i3 = i2;
                                if (s < s1) {
                                    s2 = s1; i2 = i1;
                                    if (s < s0) { s1 = s0; i1 = i0; s0 = s; i0 = col; }
                                    else { s1 = s; i1 = col; }
                                } else { s2 = s; i2 = col; }
                            } else { s3 = s; i3 = col; }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncthreads();  // TMEM and this base buffer are free again
        }
        const int row = at * TC_M + (warp & 3) * 32 + lane;
        if (row < nq) {
            const int list = sp * 2 + (warp >> 2);  // two candidate lists (column halves) per base slice
            int *ci = cand_idx + ((size_t)list * nq + row) * TC_NCAND;
            ci[0] = i0; ci[1] = i1; ci[2] = i2; ci[3] = i3;
            cand_kth[(size_t)list * nq + row] = s3;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
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

// one warp per query row: lane l re-scores candidate l (n_split * 4 <= 64 candidates, two rounds at most)
__global__ void __launch_bounds__(128) k_match_recheck(const float *__restrict__ fq, int nq, const float *__restrict__ fb, int nb,
                                                       const int *__restrict__ cand_idx, const float *__restrict__ cand_kth,
                                                       int n_split, const unsigned int *__restrict__ max_nrm2_bits,
                                                       int *__restrict__ nn, int *__restrict__ fallback_rows,
                                                       unsigned int *__restrict__ n_fallback) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    float a[33];
    double na = 0.0;
#pragma unroll
    for (int k = 0; k < 33; k++) {
        a[k] = __ldg(fq + (size_t)q * 33 + k);
        na += (double)a[k] * (double)a[k];
    }
    double best = INFINITY;
    int bi = 0x7fffffff;
    float kth = INFINITY;
    const int ncand = n_split * TC_NCAND;
    for (int c = lane; c < ncand; c += 32) {
        const int sp = c / TC_NCAND, k = c - sp * TC_NCAND;
        if (k == 0) kth = fminf(kth, cand_kth[(size_t)sp * nq + q]);
        const int j = cand_idx[((size_t)sp * nq + q) * TC_NCAND + k];
        if (j < 0 || j >= nb) continue;
        const double d = exact_dist(a, fb + (size_t)j * 33);
        if (d < best || (d == best && j < bi)) { best = d; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const float ok = __shfl_xor_sync(0xffffffffu, kth, o);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
        kth = fminf(kth, ok);
    }
    if (lane != 0) return;
    // every non-candidate j has score_j >= kth, and true ||a-b_j||^2 >= ||a||^2 + score_j - eps
    const double bmax = sqrt((double)__uint_as_float(*max_nrm2_bits));
    const double eps = 1.5e-4 * sqrt(na) * bmax + 1e-6 * (bmax * bmax + na) + 1e-30;
    const bool ok = (bi != 0x7fffffff) && (kth == INFINITY || best < (na + (double)kth) - eps);
    if (ok) {
        nn[q] = bi;
    } else {
        nn[q] = -2;
        fallback_rows[atomicAdd(n_fallback, 1u)] = q;
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
    // long column streams keep the top-4 insertions rare: only as many slices as needed to occupy the SMs
    int n_split = 1;
    while (A.n_tiles * n_split * 2 <= ctx->sm_count && n_split * 2 <= B.n_tiles && n_split < 4) n_split *= 2;
    const int n_lists = 2 * n_split;  // each CTA keeps two lists per row (column halves of a tile)
    PCR_ALLOC(cand_idx, int, (size_t)n_lists * nq * TC_NCAND);
    PCR_ALLOC(cand_kth, float, (size_t)n_lists * nq);
    PCR_ALLOC(fb_rows, int, (size_t)nq);
    PCR_ALLOC(n_fb, unsigned int, 4);
    PCR_CUDA(cudaMemsetAsync(n_fb, 0, 16, ctx->stream));
    const size_t smem = (size_t)TC_A_BYTES + 2 * (size_t)TC_B_BYTES + sizeof(TcSmem) + 1024;
    if (!ctx->match_tc_attr_set) {  // per device: set once per context
        PCR_CUDA(cudaFuncSetAttribute(k_match_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->match_tc_attr_set = true;
    }
    const int items = A.n_tiles * n_split;
    {
        KScope ks(ctx, KC_NN_FEATURES, 224.0 * ((double)nq + (double)nb * A.n_tiles) + 20.0 * nq * n_lists, 1,
                  2.0 * TC_K * (double)A.n_tiles * TC_M * (double)B.n_tiles * TC_N);
        k_match_tc<<<min(items, ctx->sm_count), 256, smem, ctx->stream>>>(A.tiles, A.n_tiles, nq, B.tiles, B.nrm2, B.n_tiles, n_split,
                                                                          cand_idx, cand_kth);
        PCR_LAUNCHED();
    }
    {
        KScope ks(ctx, KC_MATCH_MISC, 132.0 * nq * (1 + TC_NCAND * n_lists), 2);
        k_match_recheck<<<div_up((long long)nq * 32, 128), 128, 0, ctx->stream>>>(fq, nq, fb, nb, cand_idx, cand_kth, n_lists, B.max_bits, nn, fb_rows,
                                                                  n_fb);
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
