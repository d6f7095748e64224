/*
 * pcr.h — C ABI of the B200 point-cloud registration engine (libpcr_b200.so).
 *
 * Drop-in boundary for the reference's src/matcher hot path (KTC-Security-Circle/3d-matching).  The
 * reference has no FFI of its own: its boundary is the Python surface of src/matcher/ + the Ply attribute
 * protocol, and every heavy call goes to the open3d==0.19.0 wheel.  Each export below names the reference
 * call site (file:line under /root/reference) it replaces.  The Python mirror that binds these symbols with
 * ctypes is 3d-matching_b200/{matcher,ply,pcr_b200}; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every function returns PCR_OK (0) or a negative pcr_status; it never throws; pcr_last_error(ctx)
 *     returns the message of the last failure on that context.
 *   - *_dev pointers are device pointers on the context's device; *_host pointers are host memory.
 *   - clouds on the device are packed float4 (x, y, z, unused) — "xyzw"; normals likewise.
 *   - FPFH descriptors are (n, 33) fp32 row-major (row = point; the reference's Feature.data is the (33, n)
 *     transpose); correspondences are (c, 2) int32 [source index, target index]; transforms are 4x4
 *     row-major fp64.
 *   - work is enqueued on the stream given to pcr_set_stream (default: the legacy default stream); calls
 *     that return host results synchronise that stream before returning.
 *   - one context may be used by one thread at a time (PCR_ERR_BUSY otherwise); different contexts are
 *     independent.  The library owns only per-context scratch; results go to caller-provided buffers.
 */
#ifndef PCR_H
#define PCR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PCR_API __attribute__((visibility("default")))
#else
#define PCR_API
#endif

typedef enum pcr_status {
    PCR_OK = 0,
    PCR_ERR_INVALID = -1, /* bad argument (maps to ValueError) */
    PCR_ERR_CUDA = -2,    /* CUDA runtime failure (RuntimeError) */
    PCR_ERR_OOM = -3,     /* device allocation failed (MemoryError) */
    PCR_ERR_BUSY = -4,    /* context used concurrently */
    PCR_ERR_TOO_LARGE = -5, /* a voxel grid dimension beyond int32 or more than 2^63 voxel ids (grids beyond the dense-table
                               budget are down-sampled through sorted 64-bit keys; search grids enlarge their cells) */
    PCR_ERR_IO = -6         /* pcr_ply_*: file missing / unreadable / short write (OSError) */
} pcr_status;

typedef struct pcr_ctx pcr_ctx;

/* open3d RegistrationResult as plain data (src/matcher/ransac.py:134-136, benchmark_ransac.py:199-200) */
typedef struct pcr_reg_result {
    double transformation[16];
    double fitness;
    double inlier_rmse;
    int64_t inlier_count;
    int64_t sum_d2_fixed; /* sum of llrint(d2 * 2^k_d) over inliers (determinism rule D5) */
    int32_t k_d;
    int32_t iterations;   /* ICP: update steps applied; RANSAC: unused */
    int32_t converged;    /* ICP: 1 when the relative criteria stopped the loop */
    int32_t reserved;
    int64_t best_hyp;      /* RANSAC: global index of the winning hypothesis (-1: none) */
    int64_t hyp_evaluated; /* RANSAC: hypotheses consumed by the sequential-equivalent loop */
    int64_t survivors;     /* RANSAC: of those, how many passed both checkers */
    int64_t est_k;         /* RANSAC: final early-exit bound */
} pcr_reg_result;

/* one scored RANSAC hypothesis (a survivor of both checkers) */
typedef struct pcr_hyp_record {
    int64_t hyp;          /* global hypothesis index */
    int64_t inlier_count; /* source points with a target neighbour closer than max_dist */
    int64_t sum_d2_fixed;
    int32_t corr_inliers; /* correspondences with ||T s - t|| < max_dist (drives est_k) */
    int32_t reserved;
    double transformation[12]; /* rows 0..2 of the 4x4 */
} pcr_hyp_record;

/* ---- context ------------------------------------------------------------------------------------ */
PCR_API int pcr_create(int device, pcr_ctx **out);
PCR_API int pcr_destroy(pcr_ctx *ctx);
PCR_API const char *pcr_last_error(pcr_ctx *ctx);
PCR_API int pcr_set_stream(pcr_ctx *ctx, void *cuda_stream);
PCR_API int pcr_version(void);
/* number of this library's kernels launched on ctx since creation (bench.py's gpu_launches) */
PCR_API int64_t pcr_launch_count(pcr_ctx *ctx);

/* Optional per-kernel timing (the CUDA-event twin of the reference's src/utils/profiler.py): when enabled,
 * every kernel class is bracketed by an event pair on the context's stream.  pcr_kernel_stats synchronises the
 * stream and returns, per class, accumulated device ms, launches and ALGORITHMIC bytes / flops. */
typedef struct pcr_kernel_stat {
    double total_ms;
    int64_t launches;
    double bytes;
    double flops;
    double overlapped_ms; /* part of total_ms spent on pcr_align's helper context, i.e. off the critical path and
                           * time-sliced with the critical path's kernels */
} pcr_kernel_stat;
PCR_API int pcr_set_profiling(pcr_ctx *ctx, int enabled);
PCR_API int pcr_kernel_class_count(void);
PCR_API const char *pcr_kernel_class_name(int id);
PCR_API int pcr_kernel_stats(pcr_ctx *ctx, pcr_kernel_stat *out, int cap, int reset);

/* ---- layout helpers ------------------------------------------------------------------------------- */
/* (n,3) fp32 or fp64 device array -> packed float4 cloud (quantisation to fp32, rule D1) */
PCR_API int pcr_pack_xyz_f32(pcr_ctx *ctx, const float *xyz_dev, int n, float *xyzw_dev);
PCR_API int pcr_pack_xyz_f64(pcr_ctx *ctx, const double *xyz_dev, int n, float *xyzw_dev);
PCR_API int pcr_unpack_xyz_f32(pcr_ctx *ctx, const float *xyzw_dev, int n, float *xyz_dev);
/* pcd.transform(T): out = fp32(R p + t) with the specified fp64 operation order (rule D7); in-place allowed.
 * Used for RegistrationResult.correspondence_set and by _visualize_matcher.py:502-503-style callers. */
PCR_API int pcr_transform_points(pcr_ctx *ctx, const float *xyzw_dev, int n, const double *T_host /* 16 */,
                                 float *out_xyzw_dev);

/* ---- preprocessing (Ply._preprocess, src/ply/ply.py:87-135) ---------------------------------------- */
/* pcd.voxel_down_sample(voxel)            src/ply/ply.py:106.  out_xyzw_dev capacity n; *m_host = #voxels */
PCR_API int pcr_voxel_downsample(pcr_ctx *ctx, const float *xyzw_dev, int n, double voxel, float *out_xyzw_dev,
                                 int *m_host);
/* pcd.estimate_normals(KDTreeSearchParamHybrid(radius, max_nn))   src/ply/ply.py:110-112, 133-135 */
PCR_API int pcr_estimate_normals(pcr_ctx *ctx, const float *xyzw_dev, int n, double radius, int max_nn,
                                 float *normals_xyzw_dev);
/* compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius, max_nn))   src/ply/ply.py:117-120 */
PCR_API int pcr_compute_fpfh(pcr_ctx *ctx, const float *xyzw_dev, const float *normals_xyzw_dev, int n, double radius,
                             int max_nn, float *fpfh_dev /* n x 33 */);
/* KDTreeFlann.search_hybrid_vector_3d for every query (diagnostic / tests): idx, d2 are (nq, max_nn) */
PCR_API int pcr_knn_hybrid(pcr_ctx *ctx, const float *xyzw_dev, int n, const float *queries_xyzw_dev, int nq,
                           double radius, int max_nn, int *idx_dev, float *d2_dev, int *cnt_dev);
/* radius-limited 1-NN (the primitive inside registration_icp / RANSAC validation) */
PCR_API int pcr_nn1(pcr_ctx *ctx, const float *tgt_xyzw_dev, int nt, const float *queries_xyzw_dev, int nq,
                    double radius, int *idx_dev, float *d2_dev);

/* ---- feature matching ----------------------------------------------------------------------------------- */
/* correspondences_from_features(src_fpfh, tgt_fpfh, mutual_filter)   src/matcher/ransac.py:85 (and the
 * matching stage inside registration_ransac_based_on_feature_matching, src/matcher/ransac.py:42-47).
 * corr_dev capacity (ms, 2) int32; *c_host = number of pairs. */
PCR_API int pcr_match_features(pcr_ctx *ctx, const float *fs_dev, int ms, const float *ft_dev, int mt, int mutual,
                               double mutual_ratio, int *corr_dev, int *c_host);
/* one-directional exact 33-D 1-NN (nn_dev[i] = argmin_j ||fq_i - fb_j||, ties -> lowest j) */
PCR_API int pcr_nn_features(pcr_ctx *ctx, const float *fq_dev, int nq, const float *fb_dev, int nb, int *nn_dev);

/* ---- RANSAC (registration_ransac_based_on_feature_matching, src/matcher/ransac.py:42-59) ---------------- */
/* Full single-GPU loop over hypotheses [0, max_iter) with the sequential-equivalent early exit. */
PCR_API int pcr_ransac(pcr_ctx *ctx, const float *src_xyzw_dev, int ms, const float *tgt_xyzw_dev, int mt,
                       const int *corr_dev, int c, double max_dist, double edge_sim, int64_t max_iter,
                       double confidence, uint64_t seed, pcr_reg_result *result_host);
/* One wave: score hypotheses [hyp_begin, hyp_end) and return the survivors that are prefix-maxima of this
 * range — a superset: every survivor better than the running best (best_count, best_sum_d2_fixed; pass 0,0
 * at the start), sorted by hypothesis index — to records_host (capacity cap; PCR_ERR_INVALID if it does not
 * fit); *n_records_host = how many; *n_survivors_host = survivors in the range.
 * Which non-maximal survivors the superset also holds depends on kernel timing (survivors prune against the completed
 * evaluations of earlier hypotheses of the same wave); the prefix maxima themselves, and therefore everything
 * pcr_ransac_scan derives, are deterministic.
 * Used by the multi-GPU driver: each rank scores its slice, records are all-gathered, pcr_ransac_scan merges. */
PCR_API int pcr_ransac_wave(pcr_ctx *ctx, const float *src_xyzw_dev, int ms, const float *tgt_xyzw_dev, int mt,
                            const int *corr_dev, int c, double max_dist, double edge_sim, int64_t hyp_begin,
                            int64_t hyp_end, uint64_t seed, int64_t best_count, int64_t best_sum_d2_fixed,
                            pcr_hyp_record *records_host, int cap, int *n_records_host,
                            int64_t *n_survivors_host);
/* Optional session around a series of pcr_ransac_wave calls on the SAME clouds: the target search grid and the
 * spatially sorted source are built once (into buffers owned by the context) instead of once per wave.  The caller
 * must not modify the two clouds between begin and end; a wave on other buffers / sizes / max_dist simply prepares
 * its own work as before.  pcr_ransac_session_end (or pcr_destroy) releases the buffers. */
PCR_API int pcr_ransac_session_begin(pcr_ctx *ctx, const float *src_xyzw_dev, int ms, const float *tgt_xyzw_dev, int mt,
                                     double max_dist);
PCR_API int pcr_ransac_session_end(pcr_ctx *ctx);
/* Host-only: replay the sequential loop over records sorted by hypothesis index, updating *state.
 * Returns 1 in *stop_host when the early-exit bound was reached inside [hyp_begin, hyp_end). */
PCR_API int pcr_ransac_scan(const pcr_hyp_record *records_host, int n, int64_t hyp_begin, int64_t hyp_end, int c,
                            int ms, double confidence, int32_t k_d, pcr_reg_result *state, int *stop_host);
PCR_API int pcr_ransac_k_d(double max_dist, int ms);

/* ---- manual-step twins (src/matcher/ransac.py:104-277) -------------------------------------------------- */
/* Every function that takes correspondences (pcr_ransac, pcr_ransac_wave, and the two below) first checks on the
 * device that each pair indexes inside [0, ms) x [0, mt) and returns PCR_ERR_INVALID otherwise (the reference raises
 * IndexError for such a pair); inside a pcr_ransac_session a buffer is checked once. */
/* compute_step_transformation for hypotheses [h_begin, h_begin + count): 3 distinct correspondences drawn by
 * Philox(seed, h), Kabsch.  T_dev: count x 16 fp64. */
PCR_API int pcr_ransac_step(pcr_ctx *ctx, const float *src_xyzw_dev, int ms, const float *tgt_xyzw_dev, int mt,
                            const int *corr_dev, int c, uint64_t seed, int64_t h_begin, int count, double *T_dev);
/* evaluate_inlier_ratio / _fast for `count` transforms over c correspondences.  squared != 0: test
 * d^2 < thresh (the _fast variant), else ||.|| < thresh.  counts_dev: count int32 inlier counts. */
PCR_API int pcr_inlier_count(pcr_ctx *ctx, const float *src_xyzw_dev, int ms, const float *tgt_xyzw_dev, int mt,
                             const int *corr_dev, int c, const double *T_dev, int count, double thresh, int squared,
                             int *counts_dev);

/* ---- ICP (registration_icp + TransformationEstimationPointToPlane, src/matcher/icp.py:42-48) ------------- */
/* corr_dev: optional (ns) int32, target index per source point or -1. */
PCR_API int pcr_icp_point_to_plane(pcr_ctx *ctx, const float *src_xyzw_dev, int ns, const float *tgt_xyzw_dev,
                                   const float *tgt_normals_xyzw_dev, int nt, double max_dist,
                                   const double *init_host /* 16 */, int max_iter, double rel_fitness,
                                   double rel_rmse, pcr_reg_result *result_host, int *corr_dev);

/* ---- end-to-end (the north-star call: clouds + voxel size -> 4x4, fitness, inlier RMSE) ------------------ */
typedef struct pcr_align_params {
    double voxel_size;
    int64_t ransac_max_iter; /* reference default 30 (src/matcher/ransac.py:24) */
    double ransac_confidence; /* 0.999 (src/matcher/ransac.py:58) */
    uint64_t seed;
    int32_t icp_max_iter;     /* Open3D default 30 (src/matcher/icp.py:42-48 passes no criteria) */
    double icp_rel_fitness;   /* 1e-6 */
    double icp_rel_rmse;      /* 1e-6 */
    int32_t source_normals;   /* 1: also estimate full-resolution source normals, as Ply.__init__ does */
    int32_t reserved;
} pcr_align_params;

typedef struct pcr_align_result {
    pcr_reg_result ransac;
    pcr_reg_result icp;
    int32_t n_src_down, n_tgt_down, n_corr, reserved;
    float stage_ms[8]; /* voxel, normals_down, fpfh, match, ransac, normals_full, icp, total (device time) */
} pcr_align_result;

PCR_API void pcr_align_default_params(pcr_align_params *p);
/* device-resident inputs (packed float4) */
PCR_API int pcr_align(pcr_ctx *ctx, const float *src_xyzw_dev, int ns, const float *tgt_xyzw_dev, int nt,
                      const pcr_align_params *p, pcr_align_result *result_host);
/* host inputs: (n,3) fp32 host arrays; H2D copies, the full path and the D2H of the result are inside */
PCR_API int pcr_align_host(pcr_ctx *ctx, const float *src_xyz_host, int ns, const float *tgt_xyz_host, int nt,
                           const pcr_align_params *p, pcr_align_result *result_host);

/* PLY paths (src/main.py:26-31 hands paths to Ply()): both files are decoded concurrently by pcr_ply_read (below)
 * into the context's pinned staging buffer, copied to the device once each, and aligned; the timed region of a caller
 * therefore covers file -> result.  File errors return the pcr_ply_* status (PCR_ERR_IO / PCR_ERR_INVALID). */
PCR_API int pcr_align_files(pcr_ctx *ctx, const char *src_path, const char *tgt_path, const pcr_align_params *p,
                            pcr_align_result *result_host);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink (nothing in the reference: it is single-process) ----
 * The hypothesis space of one RANSAC and batches of independent pairs shard over the GPUs of a box; single-pair ICP does
 * not ("replicas only").  NCCL is bound at run time (dlopen of libnccl.so.2; PCR_NCCL_LIB overrides), so programs that
 * stay on one GPU need no NCCL at all, and with world == 1 / no communicator every call below is the single-GPU path.
 *   rank 0:      pcr_comm_unique_id(id, 128)      -> hand the 128 bytes to the other ranks (MPI, torch.distributed, a file ...)
 *   every rank:  pcr_comm_init(ctx, id, 128, rank, world)
 * pcr_ransac_multi: every rank passes the SAME clouds and correspondences; hypotheses [0, max_iter) are scored in waves
 * cut into per-rank slices, one fixed-size ncclAllGather per wave exchanges the prefix maxima, and every rank replays the
 * sequential loop: the result is bit-identical on every rank, for every world size, and to pcr_ransac.
 * first_wave (hypotheses per rank of the first, unpruned wave; <= 0: 16384) and growth (wave size factor; < 2: 4) only
 * change the schedule, never the result.  *n_waves_host (optional) = waves executed.
 * pcr_align_batch: pair i of n_total belongs to rank i % world; the caller passes ITS n_local pairs in that order
 * (device clouds).  The rank aligns them with `workers` host threads (own context and stream each; 1 = sequential) and
 * one final all-gather fills out_host (n_total x 18 doubles: 16 transform entries, fitness, inlier RMSE) on every rank. */
#define PCR_COMM_ID_BYTES 128
PCR_API int pcr_comm_unique_id(void *id_out_host, int bytes);
PCR_API int pcr_comm_init(pcr_ctx *ctx, const void *id_host, int bytes, int rank, int world);
PCR_API int pcr_comm_destroy(pcr_ctx *ctx);
PCR_API int pcr_ransac_multi(pcr_ctx *ctx, const float *src_xyzw_dev, int ms, const float *tgt_xyzw_dev, int mt,
                             const int *corr_dev, int c, double max_dist, double edge_sim, int64_t max_iter,
                             double confidence, uint64_t seed, int64_t first_wave, int growth,
                             pcr_reg_result *result_host, int *n_waves_host);
PCR_API int pcr_align_batch(pcr_ctx *ctx, int n_local, const float *const *src_xyzw_dev, const int *ns,
                            const float *const *tgt_xyzw_dev, const int *nt, const pcr_align_params *p, int workers,
                            int n_total, double *out_host);

/* ---- PLY files (host only; no context, no device work) ------------------------------------------------------
 * Replaces o3d.io.read_point_cloud (src/ply/ply.py:80) and o3d.io.write_point_cloud (trim_ply.py:40; the
 * reference's converter writes ASCII PLY, convert_stl-ply.py:8).  Formats: ascii, binary_little_endian,
 * binary_big_endian; vertex properties x y z of any scalar type, optional nx ny nz; other vertex properties and
 * other elements (faces) are skipped; list properties on the vertex element are PCR_ERR_INVALID.  A file without
 * vertices reads as n_vertex = 0 (the mirror raises the reference's ValueError, src/ply/ply.py:81-84).
 * Errors are returned as pcr_status plus a message in `err` (may be NULL).  The functions are thread-safe. */
typedef struct pcr_ply_info {
    int64_t n_vertex;
    int64_t data_offset;   /* byte offset of the first element record */
    int32_t format;        /* 0 ascii, 1 binary_little_endian, 2 binary_big_endian */
    int32_t has_normals;   /* nx ny nz present */
    int32_t has_colors;    /* red green blue present */
    int32_t n_props;       /* properties of the vertex element */
    int32_t vertex_stride; /* bytes per vertex record (binary formats; 0 for ascii) */
    int32_t reserved;
} pcr_ply_info;
PCR_API int pcr_ply_probe(const char *path, pcr_ply_info *info, char *err, int err_cap);
/* Decodes the vertices into caller-provided HOST buffers (each may be NULL): xyzw_host (n,4) fp32 packed float4
 * with w = 0 — the device layout, so pinned memory goes to the GPU with one cudaMemcpy; normals_xyzw_host (n,4)
 * fp32, written only when the file has normals (see info->has_normals); xyz64_host (n,3) fp64, the values before
 * the fp32 quantisation of rule D1.  n_cap = capacity of the buffers in vertices.  threads <= 0: automatic. */
PCR_API int pcr_ply_read(const char *path, int64_t n_cap, float *xyzw_host, float *normals_xyzw_host,
                         double *xyz64_host, int threads, pcr_ply_info *info, char *err, int err_cap);
/* Writes n vertices from packed float4 host arrays; normals (n,4) fp32 and rgb (n,3) uint8 are optional.
 * binary != 0: binary_little_endian; 0: ascii with the shortest decimal text that reads back to the same fp32. */
PCR_API int pcr_ply_write(const char *path, const float *xyzw_host, int64_t n, const float *normals_xyzw_host,
                          const unsigned char *rgb_host, int binary, char *err, int err_cap);

#ifdef __cplusplus
}
#endif
#endif /* PCR_H */
