// Host-side internals of libicp_b200.so shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/icp_b200.h"
#include <nccl.h>

namespace icpb {

struct DeviceOctree {
    Node* nodes = nullptr;
    uint32_t* parent = nullptr;  // parent node index per node (root: 0xFFFFFFFF)
    int64_t n_nodes = 0, cap_nodes = 0, cap_cell = 0, cap_pts = 0, cap_inv = 0, cap_grid = 0;
    bool inv_valid = false;
    bool full_keys = false;  // an earlier build of this tree needed octant keys down to max_depth (build_tree)
    TPoint* pts = nullptr;  // Morton-sorted target points (xyz + original index)
    int64_t n_pts = 0;
    int64_t n_leaves = 0;
    int depth = 0;          // deepest node level
    int max_pts = 10, max_depth = 20;
    double root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
    uint32_t pos_of_idx0 = 0;  // sorted position of original point 0 (findNearest's default answer)
    uint32_t* inv_perm = nullptr;  // original index -> sorted position (built lazily for the stage API)
    // search tree only: integer cell coordinates per node and the dense entry grid at level grid_level
    bool want_cell = false;
    uint64_t* cell = nullptr;
    uint2* grid = nullptr;            // pyramid of dense entry grids, levels glev_min .. glev_min + glev_n - 1
    int glev_min = 0, glev_n = 0;
    int gbase = 0;                    // index (within the pyramid) of the base level; the levels before it are coarser ones for wide balls
    double spacing = 0.0;             // typical distance between neighbouring points (base cell edge / sqrt(points per occupied base cell))
    long long goff[4] = {0, 0, 0, 0};
    int gdim[4][3] = {};
    double cube = 0.0;                // edge of the cubic root
    bool valid = false;
};

// Reusable device buffer that only grows.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time so that the library shares the process's already-loaded libnccl.so.2
// (torch's bundled copy when driven from Python) and has no link-time dependency for 1-GPU users.
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // source upload, overlapped with the target's tree build
    cudaStream_t stream_hi = nullptr;  // higher priority: the work-list kernel of one chunk of queries runs beside the next chunk's walk
    cudaEvent_t ev_chunk[9] = {};
    int opt_nn_chunks = 1;           // modes 4 / 6: walk the queries in this many chunks (>= 2^20 queries), work lists overlapped; measured
                                     // neutral on the 10^7-point scenes (helps once the registration has nearly converged, costs before)
    cudaEvent_t ev_src = nullptr;
    cudaEvent_t ev[12] = {};
    DevBuf pin_a, pin_b;       // pinned host staging (grow-only), batch path
    std::string err;
    icp_params params;
    icp_iteration_cb on_iteration = nullptr;
    icp_progress_cb on_progress = nullptr;
    icp_log_cb on_log = nullptr;
    void* user = nullptr;
    int64_t launches = 0;
    int sm_count = 148;

    DeviceOctree tree;  // the reference's octree (structure parity, literal traversal)
    DeviceOctree fast;  // isotropic search tree over the same points; match positions index ITS point order
    int opt_search_leaf = 4;         // leaf capacity of the search tree
    int opt_search_depth = 16;       // depth cap of the search tree
    int opt_grid_shift = 0;          // entry grid level relative to the median leaf depth
    long long opt_grid_max_cells = 1ll << 32;  // entries (8 B each) over the whole pyramid: up to 32 GiB of the 180 GB
    int opt_grid_levels = 3;         // pyramid height (base level + finer ones)
    int opt_grid_coarse = 1;         // levels coarser than the base one (balanced walk only: balls wider than a base cell)
    double opt_base_occupancy = 4.0; // mean points per occupied cell the base level must still have
    int opt_range_max = 96;          // inner cells up to this many points are entered as plain point ranges (swept: tools/opt_sweep.py)
    int opt_walk_bias = -100;        // cell walk: levels finer (+) or coarser (-) than 'cell >= ball box'; -100 = per mode
                                     // (measured best: -2 for the per-thread walk, 0 for the balanced one)
    int opt_walk_max_cells = 27;     // cell walk gives way to the climbing search beyond this many cells
    DevBuf tgt_raw;  // original-order target AoS (kept for the stage API)
    int64_t n_tgt = 0;

    // resident source in internal (query-coherent) order
    DevBuf sx, sy, sz, sperm;  // SoA coordinates + original index of each internal slot
    int64_t n_src = 0;
    DevBuf pos, dist, mask;    // per-query NN result (sorted target position), distance, inlier mask
    DevBuf node_io;            // per-query leaf of the last match (temporal start of the next search); mode 4: the work list
    DevBuf lb;                 // mode 4: per-query lower bound on the distance to every non-matched target point (float)
    bool opt_temporal_skip = true;  // mode 4: keep a match without searching when that bound proves it
    DevBuf cand;               // mode 5: per query, the K nearest target points of its last search (uint4; NONE = unused)
    DevBuf work2;              // mode 5: second work list (queries handed to the per-thread kernel)
    bool keep_valid = false;   // mode 5: cand / lb describe the resident source as it is now
    int opt_keep_k = 4;        // mode 5: candidates carried per query (1..4)
    double opt_keep_alpha = 2.0;  // mode 5: search ball = seed radius x alpha (a wider ball records a better bound)
    int opt_keep_bias = 0;     // mode 5: pyramid level relative to 'cell >= ball box' (+1: up to three cells per axis)
    double opt_keep_enter = 0.2;    // mode 6: the keep / collect kernels take over once the RMSE is below this fraction of the point spacing ...
    double opt_keep_exit = 0.7;     // ... and hand back to the balanced walk above this one
    double last_rmse = -1.0;        // RMSE of the last iteration of the last run over the resident source (-1: unknown)
    double opt_keep_rcap = 0.4;  // mode 5: the ball is widened up to this fraction of the base-level cell edge at most
    DevBuf part_a, part_b;     // per-block partials
    DevBuf scratch0, scratch1, scratch2, scratch3, scratch_src;
    DevBuf scratch_keys;       // octrees deeper than 21 levels: the extra key words (build.cu)
    DevBuf las_src, las_tgt;   // raw LAS point records of icp_register_las, decoded on the device (cloudio.cu)
    bool src_identity_perm = false;  // resident source is in caller order (no permutation)
    bool prev_valid = false;         // pos / node_io hold last run's matches of the resident source against the current tree
    float last_build_ms = 0.f;
    // tuning knobs (icp_set_option)
    int opt_nn_mode = 6;             // 0: literal traversal from the root; 1: climb; 2: warp tiles; 3: cell walk; 4: balanced cell walk;
                                     // 5: keep / collect (nn_keep.cu); 6: 4 while the registration moves, 5 once it has nearly converged
    bool opt_order_queries = true;   // Morton-order the source for traversal coherence
    bool opt_write_mask = false;     // keep the per-point inlier mask of the last iteration
    bool opt_count = false;          // maintain the NN path counters (same-address atomics: profiling / tests only)
    LoopState* d_state = nullptr;
    // peer-to-peer exchange of the per-iteration records (common.cuh: Mailbox); falls back to the two NCCL all-gathers when
    // the mailboxes cannot be opened on every rank
    Mailbox* mail = nullptr;
    Mailbox* peer_mail[MAIL_RANKS] = {};
    bool p2p = false;
    size_t nn_smem_opt_in = 0;    // nn_kernel's dynamic shared memory opt-in done up to this size (deep octrees)
    bool rs_smem_opt_in = false;  // radix_scatter_kernel's dynamic shared memory opt-in done for this handle's device
    int lv_grid = 0;              // blocks of the cooperative octree level kernel (all resident)
    unsigned int mail_epoch = 0;
    unsigned int mail_run = 0;       // sharded runs started on this handle (the same number on every rank): high part of the epoch
    unsigned long long* d_counters = nullptr;  // see NNArgs::counters (4 entries)
    unsigned int* d_work_count = nullptr;      // mode 4/5: lengths of the two work lists (node_io, work2)
    IterRecord* h_rec = nullptr;  // pinned, device-mapped: ring of REC_RING records, one per iteration enqueued ahead
    IterRecord* d_rec = nullptr;
    static constexpr int REC_RING = 8;
    cudaEvent_t ev_it[3 * REC_RING] = {};  // per ring slot: NN stage begins / NN stage ends / iteration ends
    int opt_lookahead = 4;           // iterations enqueued before the host looks at the records (1 when a callback, a stop flag
                                     // or the per-iteration debug counters need every record as it is produced)

    // batch of small registrations: pool of worker handles (own stream each) on this device
    std::vector<Ctx*> workers;
    int opt_batch_workers = 8;
    bool opt_batch_small = true;     // pairs of <= 2048 / 4096 points go through the one-block kernel (batch.cu)

    // multi-GPU
    NcclApi* nccl = nullptr;
    void* comm = nullptr;
    int rank = 0, n_ranks = 1;
    DevBuf gather_a, gather_b;
    // spatial redistribution of the source shards over NVLink (shard.cu): rank r ends up with the r-th slice of the cloud's
    // spatial order whatever range of the caller's order it was handed; the moved points travel back the same way
    bool opt_redistribute = true;
    bool opt_shard_target = true;    // sharded runs: each rank uploads 1/R of the target, an all-gather over NVLink does the rest
    bool rd_active = false;
    int64_t rd_n_in = 0, rd_n_recv = 0;
    int64_t rd_cnt_s[MAIL_RANKS] = {}, rd_cnt_r[MAIL_RANKS] = {};
    DevBuf rd_perm, rd_send, rd_recv, rd_tmp, rd_back;
};

#define ICPB_CUDA(ctx, call)                                                                      \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                         std::to_string(__LINE__) + ")";                                          \
            return ICP_CUDA_ERROR;                                                                \
        }                                                                                         \
    } while (0)

#define ICPB_TRY(expr)              \
    do {                            \
        int s__ = (expr);           \
        if (s__ != ICP_OK) return s__; \
    } while (0)

int devbuf_reserve(Ctx* c, DevBuf& b, size_t bytes);
int pinned_reserve(Ctx* c, DevBuf& b, size_t bytes);
void devbuf_free(DevBuf& b);

#define ICPB_NCCL(ctx, call)                                                                   \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess) {                                                              \
            (ctx)->err = std::string(#call) + ": " + (ctx)->nccl->GetErrorString(r__);         \
            return ICP_NCCL_ERROR;                                                             \
        }                                                                                      \
    } while (0)

// shard.cu: multi-GPU data movement
NcclApi* nccl_load(Ctx* c);
int target_upload_sharded(Ctx* c, const double* host_tgt_xyz, int64_t n_tgt);  // 1/R over PCIe per rank, all-gather over NVLink
int redistribute_source(Ctx* c, const double* d_xyz, int64_t n_in, const double** d_out, int64_t* n_out);
int redistribute_return(Ctx* c, const double* d_recv_order_xyz, double* d_caller_order_xyz);
int sort_pairs_u64_u32(Ctx* c, uint64_t*& keys, uint64_t*& keys_alt, uint32_t*& vals, uint32_t*& vals_alt, int64_t n, int key_bits);

// build.cu
int octree_build_device(Ctx* c, const double* d_tgt_xyz, int64_t m, int max_pts, int max_depth);
void octree_free(Ctx* c);
// Morton-orders n query points (AoS, device) for traversal coherence: fills SoA coordinates and the
// permutation (internal slot -> original index).
int order_queries(Ctx* c, const double* d_q_xyz, int64_t n, double* sx, double* sy, double* sz, uint32_t* perm);
int build_inv_perm(Ctx* c);

// nn.cu
struct NNLaunch {
    const double* sx;
    const double* sy;
    const double* sz;      // queries, SoA (may alias the outputs when apply_pending)
    double* ox;
    double* oy;
    double* oz;            // transformed queries written back when apply_pending (may be null otherwise)
    int64_t n;
    uint32_t* pos_out;     // sorted target position of the NN
    double* dist_out;      // distance
    const uint32_t* prev_pos;  // last iteration's match per query, seeds the search (may be null)
    uint32_t* node_io;         // in: node the previous search started from; out: this one's (may be null)
    float* lb_io = nullptr;         // mode 4/5: temporal bounds, valid for the positions the queries have on entry (may be null)
    uint4* cand_io = nullptr;       // mode 5: candidates of the last search per query (may be null)
    StatA* part_a;         // per-block partial (may be null: no statistics)
    const LoopState* state;  // may be null (stateless query)
    int apply_pending;     // read state->have_T / T_pending and transform on load
    int mode;              // 0: literal traversal from the root; 1: per-thread fast path; 2: warp tiles
    double init_best;      // DBL_MAX (engine) or 1e20 (CLI)
};
int nn_launch(Ctx* c, const NNLaunch& L);
int nn_grid_blocks(int64_t n);

// cloudio.cu
int las_decode_launch(Ctx* c, cudaStream_t st, const uint8_t* d_rec, int64_t n, int rl, const double* scale, const double* offset,
                      double* d_xyz);

// iter.cu
int apply_aos_launch(Ctx* c, const double* d_T16, double* xyz, int64_t n);
int unsort_launch(Ctx* c, const double* sx, const double* sy, const double* sz, const uint32_t* perm, int64_t n,
                  double* out_xyz);

}  // namespace icpb
