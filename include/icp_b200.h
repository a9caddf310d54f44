/* icp_b200.h -- C ABI of the B200-native ICP hot path (libicp_b200.so).
 *
 * Drop-in boundary for ONE path of B1AnKAlpha/IterativeClosestPoint: the ICP iteration loop of the
 * reference's core/ engine (target octree build -> exact nearest neighbour per source point -> 3-sigma
 * rejection + centroid / cross-covariance -> 3x3 SVD Kabsch solve -> apply).  Plain pointers and sizes
 * only; no C++/torch types.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - Point arrays are AoS `double xyz[n][3]`, i.e. exactly `std::vector<Point3D>::data()`
 *     (PointCloudRegistration/core/pointcloud.h:12-23: struct Point3D { double x, y, z; }).
 *   - 4x4 transforms are ROW-major double[16] everywhere, in this header and in the C++ adapter (icp_b200_engine.hpp:
 *     Mat4 = std::array<double,16>, row-major).  The reference's Eigen::Matrix4d is column-major: a caller that wants one
 *     maps it with  Eigen::Map<const Eigen::Matrix<double,4,4,Eigen::RowMajor>>(m.data())  (assigning the 16 doubles to a
 *     Matrix4d unchanged would give the transpose).
 *   - All functions return an icp_status; ICP_OK == 0.  Nothing throws, nothing falls back to the CPU:
 *     without a usable CUDA device every call fails with ICP_CUDA_ERROR.
 *   - A handle is bound to one CUDA device and is not re-entrant (the reference engine is not either:
 *     services/registrationservice.cpp:188-191 guards it with m_isRegistering).
 */
#ifndef ICP_B200_H
#define ICP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICP_B200_ABI_VERSION 2

typedef struct icp_b200_ctx* icp_handle;

/* Status codes map 1:1 to the reference's failure exits. */
typedef enum icp_status {
    ICP_OK = 0,
    ICP_EMPTY_INPUT = 1,      /* core/icpengine.cpp:26-34  finished(false, "点云数据为空" / "源点云或目标点云为空") */
    ICP_CANCELLED = 2,        /* core/icpengine.cpp:160-164 finished(false, "用户取消")                              */
    ICP_TOO_FEW_INLIERS = 3,  /* core/icpengine.cpp:319-323 finished(false, "有效点对不足")                          */
    ICP_INVALID_ARGUMENT = 4,
    ICP_CUDA_ERROR = 5,
    ICP_NCCL_ERROR = 6,
    ICP_NO_OCTREE = 7,        /* stage call issued before icp_octree_build */
    ICP_IO_ERROR = 8,         /* core/lasio.cpp:9-12,23-27,134-137: file cannot be opened / read / created */
    ICP_BAD_FORMAT = 9        /* core/lasio.cpp:30-34 "不是有效的LAS文件"; icp_registration.cpp:291-295 implausible point count */
} icp_status;

/* The engine (PointCloudRegistration/core/) and the CLI program (icp_registration.cpp) implement the same
 * loop with the differences tabulated in SURVEY.md 3.3; `variant` selects which one is reproduced. */
typedef enum icp_variant {
    ICP_VARIANT_ENGINE = 0,   /* core/icpengine.cpp:117-394 */
    ICP_VARIANT_CLI = 1       /* icp_registration.cpp:443-622 */
} icp_variant;

/* struct ICPParameters (core/icpengine.h:13-19) + variant. */
typedef struct icp_params {
    int32_t max_iterations;     /* maxIterations   = 50   */
    int32_t octree_max_points;  /* octreeMaxPoints = 10   */
    int32_t octree_max_depth;   /* octreeMaxDepth  = 20 ; 0..63 accepted (the reference's GUI offers 10..50) */
    int32_t variant;            /* icp_variant */
    double tolerance;           /* tolerance       = 1e-6 */
    double sigma_multiplier;    /* sigmaMultiplier = 3.0  */
} icp_params;

/* struct IterationResult (core/icpengine.h:24-32). */
typedef struct icp_iteration {
    int32_t iteration;          /* 1-based */
    int32_t valid_points;
    int32_t outlier_points;
    int32_t has_angles;         /* 0 for the convergence record, whose angle fields the reference leaves unset (icpengine.cpp:294-303) */
    double rmse;
    double transform[16];       /* cumulative, row-major */
    double rotation_angle;      /* degrees, acos((tr R - 1)/2) without clamping (icpengine.cpp:361) */
    double translation_distance;
    double nn_ms;               /* device time of this iteration's NN stage (not in the reference) */
    double iter_ms;             /* device time of the whole iteration */
} icp_iteration;

/* Per-iteration statistics of stages a9-a11 (core/icpengine.cpp:187-278). */
typedef struct icp_stats {
    double min_distance, max_distance;  /* over finite distances */
    double mean, std_dev, threshold, rmse, sum_sq;
    int64_t problem_count, valid_count, outlier_count;
} icp_stats;

/* struct ICPResult (core/icpengine.h:37-44) + what the CLI's ICP() returns through its out-parameters. */
typedef struct icp_result {
    int32_t status;             /* icp_status of the run */
    int32_t success;            /* ICPResult::success */
    int32_t total_iterations;   /* ICPResult::totalIterations (= history length on the success exits, 0 otherwise) */
    int32_t loop_iterations;    /* NN passes executed */
    int32_t history_len;        /* records written to `history` (kept on failure exits too) */
    int32_t history_cap;        /* in: capacity of `history` */
    double final_rmse;          /* ICPResult::finalRMSE */
    double final_R[9];          /* ICPResult::finalR ; CLI variant: the LAST incremental R (icp_registration.cpp:616-621) */
    double final_t[3];
    double cumulative_T[16];
    double last_T[16];
    icp_iteration* history;     /* in: caller-owned array of history_cap records, may be NULL */
    /* device-side timing of the run (CUDA events on the handle's stream), milliseconds */
    float ms_h2d, ms_build, ms_loop, ms_d2h, ms_nn_total, ms_nn_first;
} icp_result;

typedef struct icp_octree_info {
    int64_t n_points, n_nodes, n_leaves, node_bytes, point_bytes;
    int32_t depth, max_points, max_depth, pad_;
    double root_lo[3], root_hi[3];
    float build_ms, pad2_;      /* device time of the last icp_octree_build (keys + sort + node table) */
    /* the search structure built beside the reference tree: isotropic tree + pyramid of entry grids */
    int64_t search_nodes, search_node_bytes, grid_bytes;
    int32_t search_depth, grid_base_level, grid_fine_level, pad3_;
    double grid_base_cell;      /* edge of a base-level cell, metres */
} icp_octree_info;

/* Callbacks replacing the engine's Qt signals (core/icpengine.h:70-75).  They fire on the calling thread,
 * once per iteration, in the reference's order: log -> iterationCompleted -> progressUpdated
 * (icpengine.cpp:364-367). */
typedef void (*icp_iteration_cb)(const icp_iteration* it, void* user);
typedef void (*icp_progress_cb)(int iteration, int total, double rmse, void* user);
typedef void (*icp_log_cb)(const char* utf8_message, void* user);

/* ---- lifetime ------------------------------------------------------------------------------------ */
int icp_create(icp_handle* out, int device_id);                 /* ICPEngine::ICPEngine (icpengine.cpp:7-13) */
void icp_destroy(icp_handle h);                                 /* ICPEngine::~ICPEngine */
const char* icp_last_error(icp_handle h);                       /* detail for CUDA/NCCL/argument failures */
int icp_abi_version(void);

int icp_set_params(icp_handle h, const icp_params* p);          /* ICPEngine::setParameters (icpengine.cpp:19-22) */
int icp_get_params(icp_handle h, icp_params* p);                /* ICPEngine::getParameters (icpengine.h:61) */
void icp_default_params(icp_params* p);                         /* ICPParameters defaults (icpengine.h:13-19) */
int icp_set_callbacks(icp_handle h, icp_iteration_cb on_iteration, icp_progress_cb on_progress, icp_log_cb on_log,
                      void* user);
/* Tuning knobs that never change results (every search mode returns the reference's indices):
 *   "nn_mode"  0 literal reference traversal from the root, one query per thread; 3 per-thread walk over the entry-grid
 *              cells the search ball touches (a pruned tree search where the walk does not apply); 4 the same walk with the
 *              candidate scan balanced over the warp; 5 keep / collect: candidates and a lower bound carried between
 *              iterations settle a query without a search; 6 (default) mode 4 while the registration moves, mode 5 once it
 *              has nearly converged -- all with the literal traversal as fallback.  (1 and 2, the climbing search as a
 *              mode of its own and the warp-tile kernel, lost and were retired: ICP_INVALID_ARGUMENT.)
 *   "keep_k" "keep_alpha" "keep_rcap" "keep_bias" "keep_enter" "keep_exit"   the keep / collect path (modes 5, 6);
 *   "grid_levels" "grid_coarse" "grid_max_cells" "grid_shift" "base_occupancy" "range_max"   the entry-grid pyramid;
 *   "search_leaf" "search_depth" "walk_bias" "walk_max_cells" "nn_chunks" "temporal_skip"   search details;
 *   "order_queries" (Morton-order the source internally, default 1), "write_mask" (keep the inlier mask),
 *   "lookahead" (iterations enqueued before the host reads their records, modes 0 and 3 without callbacks / stop flag),
 *   "redistribute" "shard_target" (sharded runs: spatial re-deal of the source shards and 1/R target upload, default 1),
 *   "batch_small" "batch_workers" (icp_register_batch: one-block kernel for small pairs, worker streams), "count" (profiling
 *   counters). */
int icp_set_option(icp_handle h, const char* key, double value);

/* ---- the whole hot path, HOST buffers in and out ------------------------------------------------- */
/* ICPEngine::registerPointClouds (icpengine.cpp:24-60 -> runICP :117-394) and the CLI's ICP()
 * (icp_registration.cpp:443-446).  `src_xyz` is updated in place on the exits where the reference writes
 * the source back (success / divergence / max iterations; CLI also on <3 inliers) and left untouched
 * otherwise.  `stop_flag` (may be NULL) is polled once per iteration like m_shouldStop (icpengine.cpp:160). */
int icp_register(icp_handle h, double* src_xyz, int64_t n_src, const double* tgt_xyz, int64_t n_tgt,
                 icp_result* out, const volatile int* stop_flag);

/* Same loop on inputs that are already resident in this handle's device memory: the target uploaded by
 * icp_octree_build and the source by icp_source_upload.  Used by bench.py for the device-resident figure
 * and by the sharded driver.  `src_out_xyz` may be NULL (no write-back copy). */
int icp_source_upload(icp_handle h, const double* src_xyz, int64_t n_src);
int icp_register_resident(icp_handle h, int64_t n_src_global, icp_result* out, double* src_out_xyz,
                          const volatile int* stop_flag);

/* ---- stages, individually callable (parity tests, benchmarks) -------------------------------------- */
/* Octree::Octree (core/octree.cpp:41-77 -> buildTree :86-126; CLI twin icp_registration.cpp:155-185). */
int icp_octree_build(icp_handle h, const double* tgt_xyz, int64_t n_tgt, int max_points, int max_depth);
int icp_octree_get_info(icp_handle h, icp_octree_info* info);
/* Pre-order dump (children in octant order) for structure parity; same layout as the oracle's dump.
 * Call with NULL arrays to get the sizes. */
int icp_octree_dump(icp_handle h, int64_t* n_nodes, int64_t* n_leaf_points, int32_t* depth, uint64_t* key,
                    uint8_t* is_leaf, int32_t* count, double* box6, int32_t* leaf_idx);
/* Octree::findNearest for n queries (core/octree.cpp:175-184; loop core/icpengine.cpp:172-184).
 * idx_out receives indices into the ORIGINAL target order; dist_out (may be NULL) the distances of
 * computeDistance (icpengine.cpp:68-74).  `kernel_ms` (may be NULL) receives the NN kernel's device time. */
int icp_nn_query(icp_handle h, const double* q_xyz, int64_t n, int32_t* idx_out, double* dist_out, float* kernel_ms);
/* Stages a9-a11 on caller-supplied correspondences (core/icpengine.cpp:187-278; CLI :499-541). */
int icp_iteration_stats(icp_handle h, const double* src_xyz, int64_t n, const int32_t* idx, int iteration,
                        double* dist_out, uint8_t* inlier_mask_out, icp_stats* stats_out);
/* ICPEngine::computeBestFitTransform (icpengine.cpp:76-115) / best_fit_transform (icp_registration.cpp:389-440)
 * on n matched pairs; T_out row-major. */
int icp_best_fit_transform(icp_handle h, const double* a_xyz, const double* b_xyz, int64_t n, double* T_out);
/* The single-warp SVD + reflection fix + translation on a caller-supplied H and centroids (icpengine.cpp:93-112). */
int icp_solve_from_H(icp_handle h, const double* H9, const double* cA3, const double* cB3, double* T_out,
                     double* U9, double* S3, double* V9);
/* `src = T * src` (icpengine.cpp:345-346); also PointCloud::applyTransform's consumer path. */
int icp_apply_transform(icp_handle h, const double* T16, double* xyz, int64_t n);

/* ---- multi-GPU: source sharded by point range, target octree replicated ------------ */
/* One handle per process/GPU.  `unique_id` is NCCL's 128-byte ncclUniqueId created by rank 0
 * (icp_comm_unique_id) and distributed by the host plumbing (torch.distributed in this repo). */
int icp_comm_unique_id(icp_handle h, void* unique_id_128);
int icp_comm_init(icp_handle h, int rank, int n_ranks, const void* unique_id_128);
int icp_comm_destroy(icp_handle h);
/* icp_register on this rank's shard of the source; `n_src_global` is the total source size (the N of
 * the mean / variance).  All ranks return identical results (rank-ordered summation of the gathered partials).
 * COLLECTIVE: every rank calls it the same number of times, with the same target and parameters.  Inside, each rank
 * uploads 1/R of the target (an all-gather over NVLink hands every replica the rest) and the ranks re-deal the source
 * points by region over NVLink (a range of the caller's order says nothing about where its points are); the moved points
 * travel home before the write-back, so the caller gets ITS shard back in ITS order.  A stop flag raised on any rank ends
 * the run on every rank in the same iteration (ICP_CANCELLED, sources untouched).  icp_source_upload /
 * icp_register_resident on a handle with a communicator behave the same way.  A rank's shard may be EMPTY (n_shard = 0, the
 * pointer may then be NULL): the rank still takes part in every collective and returns the common result. */
int icp_register_sharded(icp_handle h, double* src_shard_xyz, int64_t n_shard, int64_t n_src_global,
                         const double* tgt_xyz, int64_t n_tgt, icp_result* out, const volatile int* stop_flag);

/* ---- many small independent registrations (BASELINE.json config #5) ---------------------------------- */
int icp_register_batch(icp_handle h, int32_t n_pairs, double* const* src_xyz, const int64_t* n_src,
                       const double* const* tgt_xyz, const int64_t* n_tgt, icp_result* results);

/* How many queries since the last reset were answered by the order-independent fast path and how many
 * had to be re-run through the literal reference traversal (exact ties / duplicates / 1-ulp near ties). */
int icp_nn_counters(icp_handle h, int64_t* fast_path, int64_t* literal_fallback, int reset);
/* Balanced kernels (nn_mode 4 - 6): queries handed to the per-thread search, and the number of candidates scanned
 * (profiling counters, maintained only with the "count" option). */
int icp_nn_tile_counters(icp_handle h, int64_t* per_thread_lanes, int64_t* candidates_scanned, int reset);

/* ---- data formats either side of the loop ------------------------------------ */
#define ICP_LAS_HEADER_BYTES 227   /* LAS 1.2 public header block as both reference readers/writers use it */
#define ICP_LAS_RECORD_BYTES 20    /* point data record format 0, what both reference writers emit */

/* The header fields the reference reads (core/lasio.cpp:38-48; icp_registration.cpp:282-305) and writes (:140-184). */
typedef struct icp_las_header {
    uint32_t offset_to_data;    /* byte 96  */
    uint32_t n_points;          /* byte 107 */
    uint16_t record_length;     /* byte 105 */
    uint16_t pad_[3];
    double scale[3];            /* bytes 131, 139, 147 */
    double offset[3];           /* bytes 155, 163, 171 */
    double min[3], max[3];      /* bytes 187/203/219 and 179/195/211 */
} icp_las_header;

/* LAS point records still in file form: n records `record_length` bytes apart, X/Y/Z = the first three int32. */
typedef struct icp_las_points {
    const uint8_t* records;
    int64_t n;
    int32_t record_length;
    int32_t pad_;
    double scale[3], offset[3];
} icp_las_points;

/* Header parsing; ICP_BAD_FORMAT unless the block starts with "LASF" (lasio.cpp:30-34).  Host only, no handle. */
int icp_las_parse_header(const uint8_t* header227, icp_las_header* out);
/* The point loop of LASIO::readLAS / readLASBatch (lasio.cpp:86-104, 268-287) and of readLASFile
 * (icp_registration.cpp:347-362) on the device: p = raw * scale + offset per axis.  12 B read, 24 B written per point. */
int icp_las_decode(icp_handle h, const uint8_t* records, int64_t n, int32_t record_length, const double* scale3,
                   const double* offset3, double* xyz_out);
/* The point loop of LASIO::writeLAS (lasio.cpp:192-204) and saveResultAsLAS (icp_registration.cpp:783-810):
 * (int32)((p - offset) / scale) with the reference's truncating cast, 8 zero bytes; 20-byte records out. */
int icp_las_encode(icp_handle h, const double* xyz, int64_t n, const double* scale3, const double* offset3,
                   uint8_t* records_out);
/* PointCloud::computeBounds (core/pointcloud.cpp:24-45); all zero for an empty cloud. */
int icp_cloud_bounds(icp_handle h, const double* xyz, int64_t n, double* min3, double* max3);
/* The bytes LASIO::writeLAS (variant ENGINE: scale 0.001, offset = cloud minimum; scale3/offset3 ignored) or
 * saveResultAsLAS (variant CLI: the cloud's own scale/offset) would put in the file.  `bytes_out` always receives the
 * size needed (227 + 20 n); ICP_EMPTY_INPUT for an empty cloud (lasio.cpp:128-131). */
int icp_las_file_image(icp_handle h, const double* xyz, int64_t n, int variant, const double* scale3, const double* offset3,
                       uint8_t* image_out, int64_t cap, int64_t* bytes_out);
/* LASIO::writeLAS(filename, cloud) (lasio.h:31) / saveResultAsLAS(cloud, filename) (icp_registration.cpp:698). */
int icp_las_write(icp_handle h, const char* path, const double* xyz, int64_t n, int variant, const double* scale3,
                  const double* offset3);
/* LASIO::readLAS(filename, cloud, maxPoints) (lasio.h:23; variant ENGINE) / readLASFile(filename, cloud)
 * (icp_registration.cpp:248; variant CLI: no signature check, point count must be 1..1e8, max_points ignored).
 * With xyz_out == NULL only the header is read and *n_out receives the number of points a full call returns.
 * A truncated point block is ICP_IO_ERROR (the reference would parse stale buffer contents). */
int icp_las_read(icp_handle h, const char* path, int64_t max_points, int variant, icp_las_header* header_out,
                 double* xyz_out, int64_t cap, int64_t* n_out);
/* PointCloud::downsample(targetSize) (core/pointcloud.cpp:107-128): point (int)(i * size / target) for i < target, the
 * whole cloud when it is not larger than target; ICP_EMPTY_INPUT where the reference returns nullptr. */
int icp_downsample(icp_handle h, const double* xyz, int64_t n, int32_t target_size, double* xyz_out, int64_t* n_out);
/* The CLI's 1-in-k sampling (icp_registration.cpp:857,877-882): points 0, k, 2k, ... */
int icp_downsample_stride(icp_handle h, const double* xyz, int64_t n, int64_t stride, double* xyz_out, int64_t* n_out);
/* Iteration replay of the viewer (widgets/pointcloudviewer.cpp:86-116): xyz_out = original cloud moved by one
 * IterationResult::transform through PointCloud::applyTransform (core/pointcloud.cpp:73-86); T16 == NULL is the
 * viewer's index -1 (the untouched original).  xyz_out may alias original_xyz. */
int icp_replay_iteration(icp_handle h, const double* original_xyz, int64_t n, const double* T16, double* xyz_out);
/* saveTransformation(R, t, filename, &iteration_transforms) (icp_registration.cpp:625-695): the CLI's text report,
 * byte for byte (ostream precision 10).  iteration_T16: n_iterations row-major 4x4 matrices, may be NULL.  Host only. */
int icp_save_transformation(const char* path, const double* R9, const double* t3, const double* iteration_T16,
                            int32_t n_iterations);
/* icp_register on clouds that are still LAS point records: the records go to the device as they are (12-34 B per point
 * instead of 24), are decoded there (lasio.cpp:92-99) and feed the loop directly.  The registered source is written to
 * src_out_xyz (n x 3, may be NULL) on the exits where the reference writes its source back. */
int icp_register_las(icp_handle h, const icp_las_points* src, const icp_las_points* tgt, icp_result* out,
                     double* src_out_xyz, const volatile int* stop_flag);

/* Number of kernels this library has launched on the handle since creation (bench.py's gpu_launches). */
int64_t icp_kernel_launches(icp_handle h);

#ifdef __cplusplus
}
#endif
#endif /* ICP_B200_H */
