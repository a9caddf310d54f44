// C ABI of libicp_b200.so (include/icp_b200.h): context, host orchestration of the iteration loop
// (replaces ICPEngine::registerPointClouds / runICP, core/icpengine.cpp:24-60,117-394, and the CLI's ICP(),
// icp_registration.cpp:443-622) and the sharded multi-GPU driver.
#include "internal.h"
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>

namespace icpb {

// iter.cu / build.cu entry points not in internal.h
int stat_a_blocks(Ctx* c, int64_t n);
int stat_a_launch(Ctx* c, const double* dist, int64_t n, StatA* part, StatA* rank_slot, const PeerMail* pm = nullptr);
int stage_a_reduce_launch(Ctx* c, const StatA* part, int n_part, StatA* rank_part_slot);
int dist_from_idx_launch(Ctx* c, const double* sx, const double* sy, const double* sz, const int32_t* idx, int64_t n,
                         uint32_t* pos_out, double* dist_out, StatA* part, int* n_part);
int stage_b_blocks(Ctx* c, int64_t n);
int stage_b_launch(Ctx* c, const double* sx, const double* sy, const double* sz, const uint32_t* pos, const double* dist,
                   int64_t n, int iter, const StatA* rank_a, uint8_t* mask_out, double* part, double* rank_b, const PeerMail* pm = nullptr,
                   IterRecord* rec = nullptr, int stop_req = 0);
int pairs_b_launch(Ctx* c, const double* a_xyz, const double* b_xyz, int64_t n, double* part, double* out17);
int solve_launch(Ctx* c, const double* rank_parts, int n_ranks, IterRecord* rec = nullptr);
int bestfit_launch(Ctx* c, const double* b17, const double* a0, const double* b0, double* T_out);
int solve_from_H_launch(Ctx* c, const double* in15, double* out37);
int apply_pending_launch(Ctx* c, double* x, double* y, double* z, int64_t n, float* lb = nullptr);
int unsort_results_launch(Ctx* c, const uint32_t* pos, const double* dist, const uint32_t* perm, int64_t n, int32_t* idx_out,
                          double* dist_out);
int aos_to_soa_launch(Ctx* c, const double* xyz, int64_t n, double* sx, double* sy, double* sz);
bool small_pair_eligible(int64_t n_src, int64_t n_tgt);
int small_batch_run(Ctx* c, const std::vector<int32_t>& which, double* const* src_xyz, const int64_t* n_src,
                    const double* const* tgt_xyz, const int64_t* n_tgt, std::vector<IterRecord>& recs, int rec_cap,
                    std::vector<int>& n_rec, std::vector<int>& exit_code, std::vector<char>& flagged,
                    const double*& moved_ptr, std::vector<long long>& src_off);

// ------------------------------------------------------------------------------------------------
static void log_msg(Ctx* c, const char* fmt, ...) {
    if (!c->on_log) return;
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    c->on_log(buf, c->user);
}

static int ensure_run_buffers(Ctx* c, int64_t n) {
    ICPB_TRY(devbuf_reserve(c, c->pos, (size_t)n * sizeof(uint32_t)));
    ICPB_TRY(devbuf_reserve(c, c->dist, (size_t)n * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->mask, (size_t)n));
    ICPB_TRY(devbuf_reserve(c, c->node_io, (size_t)n * sizeof(uint32_t)));
    ICPB_TRY(devbuf_reserve(c, c->lb, (size_t)n * sizeof(float)));
    if (c->opt_nn_mode >= 5) {
        ICPB_TRY(devbuf_reserve(c, c->cand, (size_t)n * sizeof(uint4)));
        ICPB_TRY(devbuf_reserve(c, c->work2, (size_t)n * sizeof(uint32_t)));
    }
    const size_t nbA = (size_t)std::max<int64_t>(stat_a_blocks(c, n), (n + 255) / 256 / 4) + 1024;
    ICPB_TRY(devbuf_reserve(c, c->part_a, nbA * sizeof(StatA)));
    ICPB_TRY(devbuf_reserve(c, c->part_b, (size_t)(stage_b_blocks(c, n) + 8) * STATB_DOUBLES * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->gather_a, (size_t)std::max(c->n_ranks, 1) * sizeof(StatA) + 64));
    ICPB_TRY(devbuf_reserve(c, c->gather_b, (size_t)std::max(c->n_ranks, 1) * STATB_DOUBLES * sizeof(double) + 64));
    return ICP_OK;
}

static int ensure_source_buffers(Ctx* c, int64_t n) {
    ICPB_TRY(devbuf_reserve(c, c->sx, (size_t)n * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->sy, (size_t)n * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->sz, (size_t)n * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->sperm, (size_t)n * sizeof(uint32_t)));
    return ICP_OK;
}

static void identity16(double* T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

// rotationAngle / translationDistance of a cumulative transform (icpengine.cpp:356-362); evaluated on the host
// with the C library's acos, as the reference does.  Summation orders follow the reference build's Eigen
// reductions, pinned by tests/golden/engine_*.npz: trace = a0 + (a1 + a2) (unrolled scalar redux), squared norm of
// the translation block = (t0^2 + t1^2) + t2^2 (packet first, scalar tail).
static void angles_of(const double* T, double* angle_deg, double* trans) {
    volatile double tr = T[0] + (T[5] + T[10]);
    *angle_deg = std::acos((tr - 1.0) / 2.0) * 180.0 / M_PI;
    volatile double a = T[3] * T[3], b = T[7] * T[7], cc = T[11] * T[11];
    volatile double s = a + b;
    *trans = std::sqrt(s + cc);
}

// Upload a host AoS cloud into `stage` (device) -- pinned staging is left to the caller's allocator; plain
// cudaMemcpyAsync from pageable memory is what a drop-in caller with a std::vector will hit.
static int upload(Ctx* c, DevBuf& stage, const double* host_xyz, int64_t n) {
    ICPB_TRY(devbuf_reserve(c, stage, (size_t)n * 3 * sizeof(double)));
    ICPB_CUDA(c, cudaMemcpyAsync(stage.p, host_xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return ICP_OK;
}

static int source_from_device_aos(Ctx* c, const double* d_xyz, int64_t n) {
    c->prev_valid = false;
    ICPB_TRY(ensure_source_buffers(c, n));
    c->n_src = n;
    if (c->opt_order_queries && n > 1)
        return order_queries(c, d_xyz, n, (double*)c->sx.p, (double*)c->sy.p, (double*)c->sz.p, (uint32_t*)c->sperm.p);
    c->src_identity_perm = true;
    return aos_to_soa_launch(c, d_xyz, n, (double*)c->sx.p, (double*)c->sy.p, (double*)c->sz.p);
}

// Turns the per-iteration device records into the reference's result structures, applying the reference's own
// bookkeeping (history, totalIterations, finalRMSE, which transform is "final", which exits write the source back).
// Shared by the iteration loop below and by the batched small-problem path (batch.cu).
struct RunAcc {
    Ctx* c;
    icp_result* out;
    int variant, max_iterations;
    long long n_global;
    int n_hist = 0;
    double last_rmse = 0.0, prev_error = 1e10;
    double T_last[16], T_cum[16];
    bool write_back = true;

    RunAcc(Ctx* ctx, icp_result* o, int var, int max_it, long long ng) : c(ctx), out(o), variant(var), max_iterations(max_it), n_global(ng) {
        identity16(T_last);
        identity16(T_cum);
        out->history_len = 0;
        out->loop_iterations = 0;
        out->status = ICP_OK;
    }
    void push(const icp_iteration& it, bool notify) {
        if (out->history && n_hist < out->history_cap) out->history[n_hist] = it;
        ++n_hist;
        last_rmse = it.rmse;
        if (notify && c->on_iteration) c->on_iteration(&it, c->user);                                       // icpengine.cpp:366
        if (notify && c->on_progress) c->on_progress(it.iteration, max_iterations, it.rmse, c->user);       // :367
    }
    // returns true if the loop goes on
    bool consume(const IterRecord& rec, int iter, float nn_ms, float iter_ms, bool notify) {
        out->loop_iterations = iter + 1;
        if (rec.exit_code == 4) {  // a rank of a sharded run was asked to stop (icpengine.cpp:160-164): nothing of this iteration counts
            log_msg(c, u8"配准已停止");
            out->status = ICP_CANCELLED;
            out->loop_iterations = iter;
            write_back = false;
            return false;
        }
        // the reference's own texts (core/icpengine.cpp:227-232,257-260,280-284; CLI: one console line per iteration,
        // icp_registration.cpp:478,543-545)
        if (variant == ICP_VARIANT_ENGINE) {
            if (rec.problems > 0.0) log_msg(c, u8"警告: 发现 %.0f 个异常距离值", rec.problems);
            log_msg(c, u8"  距离范围: 最小=%.6f, 最大=%.6f", rec.dmin, rec.dmax);
            log_msg(c, u8"  距离统计: 均值=%.6f, 标准差=%.6f, 阈值=%.6f", rec.mean, rec.std_dev, rec.threshold);
            log_msg(c, u8"  RMSE = %.6f (有效点: %d/%lld, 剔除离群点: %d)", rec.rmse, rec.valid_points, n_global, rec.outlier_points);
        } else {
            log_msg(c, u8"迭代 %d/%d ... RMSE = %g (有效点: %d/%lld, 剔除离群点: %d)", iter + 1, max_iterations, rec.rmse, rec.valid_points,
                    n_global, rec.outlier_points);
        }
        std::memcpy(T_cum, rec.T_cum, sizeof T_cum);
        std::memcpy(T_last, rec.T_last, sizeof T_last);
        if (rec.exit_code == 0 || rec.exit_code == 3) prev_error = rec.rmse;
        icp_iteration it;
        std::memset(&it, 0, sizeof it);
        it.iteration = iter + 1;
        it.rmse = rec.rmse;
        it.valid_points = rec.valid_points;
        it.outlier_points = rec.outlier_points;
        it.nn_ms = nn_ms;
        it.iter_ms = iter_ms;
        std::memcpy(it.transform, rec.T_cum, sizeof it.transform);
        if (rec.exit_code == 1) {  // converged: icpengine.cpp:291-305 ; CLI :551-554
            log_msg(c, u8"收敛达到! 迭代次数: %d", iter + 1);  // icpengine.cpp:291 ; icp_registration.cpp:552
            if (variant == ICP_VARIANT_ENGINE) {
                it.has_angles = 0;  // the reference leaves the angle fields of this record unset (:294-303)
                push(it, notify);
            }
            return false;
        }
        if (rec.exit_code == 2) {  // error grew: icpengine.cpp:311-314
            log_msg(c, variant == ICP_VARIANT_ENGINE ? u8"警告: 误差增加，停止迭代" : u8"警告: 误差增加,停止迭代。");  // :312 ; CLI :560
            return false;
        }
        if (rec.exit_code == 3) {  // < 3 inliers: icpengine.cpp:319-323 ; CLI :567-570 breaks and writes back
            log_msg(c, variant == ICP_VARIANT_ENGINE ? u8"错误: 有效点对不足，无法计算变换"
                                                     : u8"警告: 有效点对太少（< 3），无法计算变换，停止迭代。");  // :320 ; CLI :568
            if (variant == ICP_VARIANT_ENGINE) {
                out->status = ICP_TOO_FEW_INLIERS;
                write_back = false;
            }
            return false;
        }
        it.has_angles = 1;
        angles_of(rec.T_cum, &it.rotation_angle, &it.translation_distance);
        push(it, notify);
        return true;
    }
    void finish() {
        std::memcpy(out->cumulative_T, T_cum, sizeof T_cum);
        std::memcpy(out->last_T, T_last, sizeof T_last);
        if (write_back) {
            out->success = 1;
            out->total_iterations = n_hist;                                  // icpengine.cpp:386
            out->final_rmse = (variant == ICP_VARIANT_CLI) ? prev_error : (n_hist ? last_rmse : 0.0);
            const double* F = (variant == ICP_VARIANT_CLI) ? T_last : T_cum;  // CLI: last incremental T (:616-621)
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) out->final_R[3 * i + j] = F[4 * i + j];
                out->final_t[i] = F[4 * i + 3];
            }
        } else {
            out->success = 0;
            out->total_iterations = 0;
            out->final_rmse = 0.0;
            std::memset(out->final_R, 0, sizeof out->final_R);
            std::memset(out->final_t, 0, sizeof out->final_t);
        }
        out->history_len = n_hist < (out->history ? out->history_cap : 0) ? n_hist : (out->history ? out->history_cap : 0);
    }
};

// The iteration loop on the resident source/target.  n_global = N of the mean/variance.
static int run_loop(Ctx* c, int64_t n_global, icp_result* out, const volatile int* stop_flag, bool* write_back) {
    const icp_params& P = c->params;
    const int64_t n = c->n_src;
    const int variant = P.variant;
    ICPB_TRY(ensure_run_buffers(c, std::max<int64_t>(n, 1)));

    LoopState hs;
    std::memset(&hs, 0, sizeof hs);
    hs.prev_error = 1e10;  // icpengine.cpp:156
    hs.no_improve = 0;
    identity16(hs.T_pending);
    identity16(hs.T_last);
    identity16(hs.T_cum);
    for (int a = 0; a < 3; ++a) hs.pivot_a[a] = hs.pivot_b[a] = 0.5 * (c->tree.root_lo[a] + c->tree.root_hi[a]);
    hs.tolerance = P.tolerance;
    hs.sigma = (variant == ICP_VARIANT_CLI) ? 3.0 : P.sigma_multiplier;  // icp_registration.cpp:523
    hs.variant = variant;
    hs.max_iterations = P.max_iterations;
    hs.n_global = n_global;
    ICPB_CUDA(c, cudaMemcpyAsync(c->d_state, &hs, sizeof hs, cudaMemcpyHostToDevice, c->stream));

    RunAcc acc(c, out, variant, P.max_iterations, (long long)n_global);
    out->ms_nn_total = 0.f;
    out->ms_nn_first = 0.f;
    *write_back = true;

    StatA* part_a = (StatA*)c->part_a.p + 64;  // first 64 entries are scratch of the build
    StatA* rank_a = (StatA*)c->gather_a.p;
    double* rank_b = (double*)c->gather_b.p;
    double* part_b = (double*)c->part_b.p;
    const double init_best = (variant == ICP_VARIANT_CLI) ? 1e20 : DBL_MAX;  // octree.cpp:180 vs icp_registration.cpp:201

    const bool resume = c->prev_valid;  // same resident source and tree as the last run: its matches seed this one
    c->prev_valid = false;
    if (c->opt_nn_mode == 4)  // temporal bounds belong to one run: the source moves between runs
        ICPB_CUDA(c, cudaMemsetAsync(c->lb.p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(float), c->stream));
    // modes 5 / 6: candidates and bounds survive between runs over the same resident source and tree (the loop's last
    // transform is applied by apply_pending_launch, which shrinks the bounds by the distance each point moves)
    const bool keep_state = c->opt_nn_mode >= 5 && resume && c->keep_valid;
    if (c->opt_nn_mode == 5 && !keep_state) {
        ICPB_CUDA(c, cudaMemsetAsync(c->lb.p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(float), c->stream));
        ICPB_CUDA(c, cudaMemsetAsync(c->cand.p, 0xFF, (size_t)std::max<int64_t>(n, 1) * sizeof(uint4), c->stream));
    }
    c->keep_valid = false;
    // mode 6: the balanced walk while the registration still moves by a good fraction of the point spacing per iteration
    // (few matches survive an iteration, recording bounds only costs), keep / collect once it has (nearly) converged:
    // phase 0 walk; 1 walk that also records bounds (the hand-over iteration); 2 keep / collect
    int phase = (c->opt_nn_mode == 6 && keep_state) ? 2 : 0;
    double prev_rmse = (c->opt_nn_mode == 6 && resume) ? c->last_rmse : -1.0;
    if (c->n_ranks > 1 && ++c->mail_run >= (1u << 20)) c->mail_run = 1u;
    ICPB_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    // The device decides when the loop ends (solve_step: LoopState::exit_code); every kernel of an iteration returns at once
    // when it already has.  So the host may enqueue several iterations before it reads their records: no round trip per
    // iteration.  A callback, a stop flag, the host-side mode schedule of nn_mode 6 or the per-iteration debug counters need
    // every record as it is produced (the reference calls back and polls stop() once per iteration, icpengine.cpp:160-164,364-367).
    const bool debug_iter = c->opt_count && getenv("ICP_B200_DEBUG_ITER");
    const bool each = stop_flag || c->on_iteration || c->on_progress || c->on_log || debug_iter ||
                      !(c->opt_nn_mode == 0 || c->opt_nn_mode == 3);
    const int ahead = each ? 1 : std::min(std::max(c->opt_lookahead, 1), (int)Ctx::REC_RING);
    bool go_on = true;
    for (int iter0 = 0; iter0 < P.max_iterations && go_on; iter0 += ahead) {
    const int n_batch = std::min(ahead, P.max_iterations - iter0);
    for (int slot = 0; slot < n_batch; ++slot) {
        const int iter = iter0 + slot;
        // A sharded run must leave on every rank in the same iteration: the request travels with the stage-B records instead
        // (solve_step, exit code 4) -- a rank that broke out here would leave its peers waiting for its records.
        const int stop_req = (variant == ICP_VARIANT_ENGINE && stop_flag && *stop_flag) ? 1 : 0;
        if (stop_req && c->n_ranks <= 1) {  // icpengine.cpp:160-164
            log_msg(c, u8"配准已停止");
            out->status = ICP_CANCELLED;
            acc.write_back = false;
            go_on = false;
            break;
        }
        if (variant == ICP_VARIANT_ENGINE) log_msg(c, u8"迭代 %d/%d ...", iter + 1, P.max_iterations);  // icpengine.cpp:166
        c->h_rec[slot].iteration = 0;  // (an iteration that found the loop ended leaves its record untouched)

        NNLaunch L;
        L.sx = (double*)c->sx.p; L.sy = (double*)c->sy.p; L.sz = (double*)c->sz.p;
        L.ox = (double*)c->sx.p; L.oy = (double*)c->sy.p; L.oz = (double*)c->sz.p;
        L.n = n;
        L.pos_out = (uint32_t*)c->pos.p;
        L.dist_out = (double*)c->dist.p;
        L.prev_pos = ((iter > 0 || resume) && c->opt_nn_mode >= 1) ? (uint32_t*)c->pos.p : nullptr;
        L.node_io = nullptr;
        L.lb_io = (float*)c->lb.p;
        L.cand_io = (c->opt_nn_mode >= 5) ? (uint4*)c->cand.p : nullptr;
        L.part_a = nullptr;
        L.state = c->d_state;
        L.apply_pending = 1;
        L.mode = c->opt_nn_mode;
        if (c->opt_nn_mode == 6) {
            const double sp = c->fast.spacing;
            const bool near = prev_rmse >= 0.0 && prev_rmse <= c->opt_keep_enter * sp;
            const bool stay = prev_rmse >= 0.0 && prev_rmse <= c->opt_keep_exit * sp;
            L.mode = 4;
            if (!L.prev_pos || !c->opt_temporal_skip || !(phase == 0 ? near : stay)) {
                phase = 0;
                L.lb_io = nullptr;
                L.cand_io = nullptr;
            } else if (phase == 0) {
                phase = 1;
                ICPB_CUDA(c, cudaMemsetAsync(c->lb.p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(float), c->stream));
                ICPB_CUDA(c, cudaMemsetAsync(c->cand.p, 0xFF, (size_t)std::max<int64_t>(n, 1) * sizeof(uint4), c->stream));
            } else {
                phase = 2;
                L.mode = 5;
            }
        }
        L.init_best = init_best;
        ICPB_CUDA(c, cudaEventRecord(c->ev_it[3 * slot], c->stream));
        ICPB_TRY(nn_launch(c, L));
        ICPB_CUDA(c, cudaEventRecord(c->ev_it[3 * slot + 1], c->stream));

        // stage A: this rank's Chan partial of the distances; ranks exchange partials, every rank merges them in rank order
        PeerMail pm;
        pm.epoch = 0u;
        pm.n_ranks = c->n_ranks;
        pm.rank = c->rank;
        const bool p2p = c->n_ranks > 1 && c->p2p && P.max_iterations < 4000;
        if (p2p) {
            // epoch = (run, iteration): every rank enters a run with the same run number (icp_register_sharded is a collective
            // call), so a rank that left the previous run early -- an error on its side -- cannot be satisfied by stale flags
            for (int r = 0; r < MAIL_RANKS; ++r) pm.peer[r] = c->peer_mail[r];
            pm.epoch = c->mail_run * 4096u + (unsigned int)iter + 1u;
        }
        ICPB_TRY(stat_a_launch(c, L.dist_out, n, part_a, rank_a + c->rank, p2p ? &pm : nullptr));
        if (c->n_ranks > 1 && !p2p)
            ICPB_NCCL(c, c->nccl->AllGather(rank_a + c->rank, rank_a, sizeof(StatA) / sizeof(double), ncclFloat64,
                                            (ncclComm_t)c->comm, c->stream));
        // stage B (+ the solve on a single rank)
        ICPB_TRY(stage_b_launch(c, L.sx, L.sy, L.sz, L.pos_out, L.dist_out, n, iter, rank_a,
                                c->opt_write_mask ? (uint8_t*)c->mask.p : nullptr, part_b, rank_b, p2p ? &pm : nullptr, c->d_rec + slot,
                                c->n_ranks > 1 ? stop_req : 0));
        if (c->n_ranks > 1 && !p2p) {
            ICPB_NCCL(c, c->nccl->AllGather(rank_b + (size_t)c->rank * STATB_DOUBLES, rank_b, STATB_DOUBLES, ncclFloat64,
                                            (ncclComm_t)c->comm, c->stream));
            ICPB_TRY(solve_launch(c, rank_b, c->n_ranks, c->d_rec + slot));
        }
        ICPB_CUDA(c, cudaEventRecord(c->ev_it[3 * slot + 2], c->stream));
    }
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int slot = 0; slot < n_batch && go_on; ++slot) {
        const int iter = iter0 + slot;
        const IterRecord rec = c->h_rec[slot];
        if (rec.iteration != iter + 1) {  // cannot happen while go_on: the record of the ending iteration stops the host first
            c->err = "iteration record missing";
            return ICP_CUDA_ERROR;
        }
        float nn_ms = 0.f, iter_ms = 0.f;
        cudaEventElapsedTime(&nn_ms, c->ev_it[3 * slot], c->ev_it[3 * slot + 1]);
        cudaEventElapsedTime(&iter_ms, c->ev_it[3 * slot], c->ev_it[3 * slot + 2]);
        out->ms_nn_total += nn_ms;
        if (iter == 0) out->ms_nn_first = nn_ms;
        out->loop_iterations = iter + 1;
        if (debug_iter) {  // profiling aid: counters and work-list lengths of this iteration
            unsigned long long w[16];
            unsigned int wc[16];
            cudaMemcpy(w, c->d_counters, sizeof w, cudaMemcpyDeviceToHost);
            cudaMemcpy(wc, c->d_work_count, sizeof wc, cudaMemcpyDeviceToHost);
            for (int k = 2; k < 10; ++k) wc[0] += wc[k];  // the balanced walk keeps one list per chunk of queries
            fprintf(stderr, "[icp_b200] iter %d nn %.3f ms: settled=%llu literal=%llu slow=%llu candidates=%llu items=%llu kept=%llu | list1=%u list2=%u"
                            " | left the walk: no seed %llu, wide ball %llu, crowded cell %llu, queue overflow %llu, no unique minimum %llu\n",
                    iter, nn_ms, w[0], w[1], w[2], w[3], w[4], w[5], wc[0], wc[1], w[9], w[10], w[11], w[12], w[13]);
            cudaMemset(c->d_counters, 0, sizeof w);
        }
        prev_rmse = rec.rmse;
        if (!acc.consume(rec, iter, nn_ms, iter_ms, true)) go_on = false;
    }
    }
    ICPB_CUDA(c, cudaEventRecord(c->ev[3], c->stream));

    c->prev_valid = out->loop_iterations > 0 && c->opt_nn_mode >= 1;
    *write_back = acc.write_back;
    if (acc.write_back)  // the last T, if any
        ICPB_TRY(apply_pending_launch(c, (double*)c->sx.p, (double*)c->sy.p, (double*)c->sz.p, n,
                                      (c->opt_nn_mode == 5 || phase >= 1) ? (float*)c->lb.p : nullptr));
    c->keep_valid = c->prev_valid && acc.write_back && (c->opt_nn_mode == 5 || phase >= 1) && c->opt_temporal_skip;
    c->last_rmse = prev_rmse;
    acc.finish();
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
        out->ms_loop = ms;
    }
    return ICP_OK;
}

// The resident source (internal order) back into the caller's array.  After a spatial redistribution (shard.cu) the rank holds
// other ranks' points: they first travel home over NVLink.  `n_caller` = points of the caller's array on this rank.
static int write_back_source(Ctx* c, double* host_xyz, int64_t n_caller) {
    const int64_t n = c->n_src;  // resident points
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)std::max<int64_t>(n, 1) * 3 * sizeof(double)));
    ICPB_TRY(unsort_launch(c, (double*)c->sx.p, (double*)c->sy.p, (double*)c->sz.p,
                           c->src_identity_perm ? nullptr : (uint32_t*)c->sperm.p, n, (double*)c->scratch1.p));
    const double* d_out = (const double*)c->scratch1.p;
    if (c->rd_active) {
        if (n_caller != c->rd_n_in) {
            c->err = "write-back: the caller's array does not have the size of the shard that was uploaded";
            return ICP_INVALID_ARGUMENT;
        }
        ICPB_TRY(devbuf_reserve(c, c->scratch2, (size_t)std::max<int64_t>(n_caller, 1) * 3 * sizeof(double)));
        ICPB_TRY(redistribute_return(c, (const double*)c->scratch1.p, (double*)c->scratch2.p));
        d_out = (const double*)c->scratch2.p;
    } else if (n_caller != n) {
        c->err = "write-back: size mismatch";
        return ICP_INVALID_ARGUMENT;
    }
    if (n_caller > 0 && host_xyz)
        ICPB_CUDA(c, cudaMemcpyAsync(host_xyz, d_out, (size_t)n_caller * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

// Device AoS points of this rank -> resident source; on a sharded handle the ranks first re-deal the points by region.
static int source_from_shard(Ctx* c, const double* d_xyz, int64_t n) {
    const double* d_use = d_xyz;
    int64_t n_use = n;
    c->rd_active = false;
    if (c->n_ranks > 1) ICPB_TRY(redistribute_source(c, d_xyz, n, &d_use, &n_use));
    c->n_src = n_use;
    if (n_use <= 0) return ICP_OK;
    return source_from_device_aos(c, d_use, n_use);
}

static void init_result(icp_result* out) {
    icp_iteration* h = out->history;
    int cap = out->history_cap;
    std::memset(out, 0, sizeof *out);
    out->history = h;
    out->history_cap = h ? cap : 0;
    identity16(out->cumulative_T);
    identity16(out->last_T);
}

// Stages one cloud on the device as AoS doubles: either a plain copy of xyz, or -- when `las` is given -- a copy of the
// raw LAS records followed by the decode kernel (cloudio.cu), so only the records cross PCIe.
static int stage_cloud(Ctx* c, cudaStream_t st, DevBuf& dst, DevBuf& rec_stage, const double* xyz, const icp_las_points* las, int64_t n) {
    ICPB_TRY(devbuf_reserve(c, dst, (size_t)n * 3 * sizeof(double)));
    if (!las) {
        ICPB_CUDA(c, cudaMemcpyAsync(dst.p, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
        return ICP_OK;
    }
    const size_t bytes = (size_t)n * (size_t)las->record_length;
    ICPB_TRY(devbuf_reserve(c, rec_stage, bytes));
    ICPB_CUDA(c, cudaMemcpyAsync(rec_stage.p, las->records, bytes, cudaMemcpyHostToDevice, st));
    return las_decode_launch(c, st, (const uint8_t*)rec_stage.p, n, las->record_length, las->scale, las->offset, (double*)dst.p);
}

// "八叉树测试" (core/icpengine.cpp:127-137): the reference looks up the nearest target point of the FIRST source point before the
// loop and logs it.  One query through the literal traversal, only when a log callback is installed.
static int log_octree_test(Ctx* c, const double* src_xyz, const double* tgt_xyz) {
    ICPB_TRY(devbuf_reserve(c, c->scratch3, 256));
    double* d = (double*)c->scratch3.p;  // x, y, z, dist, pos
    ICPB_CUDA(c, cudaMemcpyAsync(d, src_xyz, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NNLaunch L;
    L.sx = d; L.sy = d + 1; L.sz = d + 2;
    L.ox = L.oy = L.oz = nullptr;
    L.n = 1;
    L.pos_out = (uint32_t*)(d + 4);
    L.dist_out = d + 3;
    L.prev_pos = nullptr;
    L.node_io = nullptr;
    L.part_a = nullptr;
    L.state = nullptr;
    L.apply_pending = 0;
    L.mode = 0;
    L.init_best = DBL_MAX;
    ICPB_TRY(nn_launch(c, L));
    double dist = 0.0;
    uint32_t pos = 0;
    ICPB_CUDA(c, cudaMemcpyAsync(&dist, d + 3, sizeof dist, cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaMemcpyAsync(&pos, d + 4, sizeof pos, cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    TPoint tp;
    ICPB_CUDA(c, cudaMemcpy(&tp, c->fast.pts + pos, sizeof tp, cudaMemcpyDeviceToHost));
    const long long idx = tp.idx;
    log_msg(c, u8"八叉树测试: 查询点(%.3f,%.3f,%.3f) -> 最近点[%lld](%.3f,%.3f,%.3f), 距离=%.3f", src_xyz[0], src_xyz[1], src_xyz[2], idx,
            tgt_xyz[3 * idx], tgt_xyz[3 * idx + 1], tgt_xyz[3 * idx + 2], dist);
    return ICP_OK;
}

static int register_impl(Ctx* c, double* src_xyz, int64_t n_src, int64_t n_src_global, const double* tgt_xyz, int64_t n_tgt,
                         icp_result* out, const volatile int* stop_flag, const icp_las_points* src_las = nullptr,
                         const icp_las_points* tgt_las = nullptr) {
    init_result(out);
    // (an EMPTY shard of a sharded source is a supported case: the rank still takes part in every collective)
    const bool no_source = c->n_ranks > 1 ? (n_src < 0 || (n_src > 0 && !src_xyz)) : ((!src_xyz && !src_las) || n_src <= 0);
    if (no_source || (!tgt_xyz && !tgt_las) || n_src_global <= 0 || n_tgt <= 0) {  // icpengine.cpp:26-34
        out->status = ICP_EMPTY_INPUT;
        return ICP_EMPTY_INPUT;
    }
    ICPB_CUDA(c, cudaSetDevice(c->device));
    const icp_params& P = c->params;
    const int leaf = (P.variant == ICP_VARIANT_CLI) ? 10 : P.octree_max_points;  // icp_registration.cpp:454
    const int depth = (P.variant == ICP_VARIANT_CLI) ? 20 : P.octree_max_depth;
    const bool engine_log = P.variant == ICP_VARIANT_ENGINE;
    log_msg(c, engine_log ? u8"========== 开始ICP配准 ==========" : u8"开始ICP精匹配...");  // icpengine.cpp:43-45 ; CLI :448-450
    log_msg(c, u8"源点云: %lld 个点", (long long)n_src_global);
    log_msg(c, u8"目标点云: %lld 个点", (long long)n_tgt);
    if (engine_log && c->n_ranks <= 1) {  // icpengine.cpp:48-57
        if (src_xyz && n_src > 0) log_msg(c, u8"源点云第一个点: (%.3f, %.3f, %.3f)", src_xyz[0], src_xyz[1], src_xyz[2]);
        if (tgt_xyz && n_tgt > 0) log_msg(c, u8"目标点云第一个点: (%.3f, %.3f, %.3f)", tgt_xyz[0], tgt_xyz[1], tgt_xyz[2]);
    }

    ICPB_CUDA(c, cudaEventRecord(c->ev[4], c->stream));
    if (c->n_ranks > 1 && c->comm && c->opt_shard_target && tgt_xyz && !tgt_las)
        ICPB_TRY(target_upload_sharded(c, tgt_xyz, n_tgt));  // 1/R of it over this rank's PCIe link, the rest over NVLink
    else
        ICPB_TRY(stage_cloud(c, c->stream, c->tgt_raw, c->las_tgt, tgt_xyz, tgt_las, n_tgt));
    c->n_tgt = n_tgt;
    // the source goes up on a second stream so that (from pinned memory) its copy overlaps the target's tree build
    DevBuf& src_stage = c->scratch_src;
    ICPB_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    if (n_src > 0) {
        ICPB_TRY(devbuf_reserve(c, src_stage, (size_t)n_src * 3 * sizeof(double)));
        if (src_las) ICPB_TRY(devbuf_reserve(c, c->las_src, (size_t)n_src * (size_t)src_las->record_length));
        ICPB_CUDA(c, cudaStreamWaitEvent(c->stream2, c->ev[5], 0));  // after the target copy: one copy engine direction, FIFO
        ICPB_TRY(stage_cloud(c, c->stream2, src_stage, c->las_src, src_xyz, src_las, n_src));
        ICPB_CUDA(c, cudaEventRecord(c->ev_src, c->stream2));
    }
    if (engine_log) log_msg(c, u8"构建目标点云八叉树索引...");  // icpengine.cpp:119
    ICPB_TRY(octree_build_device(c, (const double*)c->tgt_raw.p, n_tgt, leaf, depth));
    log_msg(c, engine_log ? u8"八叉树构建完成!" : u8"构建八叉树索引... 完成!");  // :124 ; CLI :453-455
    if (engine_log && c->on_log && c->n_ranks <= 1 && src_xyz && tgt_xyz && n_src > 0) ICPB_TRY(log_octree_test(c, src_xyz, tgt_xyz));
    c->src_identity_perm = false;
    c->n_src = n_src;
    if (n_src > 0) ICPB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_src, 0));
    if (n_src > 0 || c->n_ranks > 1) ICPB_TRY(source_from_shard(c, (const double*)src_stage.p, n_src));  // (collective on a sharded handle)
    ICPB_CUDA(c, cudaEventRecord(c->ev[6], c->stream));

    bool write_back = true;
    ICPB_TRY(run_loop(c, n_src_global, out, stop_flag, &write_back));
    ICPB_CUDA(c, cudaEventRecord(c->ev[7], c->stream));
    if (write_back && (c->rd_active || (n_src > 0 && src_xyz))) ICPB_TRY(write_back_source(c, src_xyz, n_src));  // icpengine.cpp:371-375
    cudaEvent_t end;
    ICPB_CUDA(c, cudaEventCreate(&end));
    ICPB_CUDA(c, cudaEventRecord(end, c->stream));
    ICPB_CUDA(c, cudaEventSynchronize(end));
    cudaEventElapsedTime(&out->ms_h2d, c->ev[4], c->ev[5]);
    cudaEventElapsedTime(&out->ms_build, c->ev[5], c->ev[6]);
    cudaEventElapsedTime(&out->ms_d2h, c->ev[7], end);
    cudaEventDestroy(end);
    if (out->status == ICP_OK) {
        if (engine_log) {  // icpengine.cpp:389-391
            log_msg(c, u8"========== 配准完成 ==========");
            log_msg(c, u8"总迭代次数: %d", out->total_iterations);
            log_msg(c, u8"最终RMSE: %.6f", out->final_rmse);
        } else {
            log_msg(c, u8"最终RMSE: %g", out->final_rmse);  // icp_registration.cpp:606
        }
    }
    return out->status;
}

}  // namespace icpb

using namespace icpb;

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int icp_abi_version(void) { return ICP_B200_ABI_VERSION; }

void icp_default_params(icp_params* p) {
    if (!p) return;
    p->max_iterations = 50;
    p->tolerance = 1e-6;
    p->sigma_multiplier = 3.0;
    p->octree_max_points = 10;
    p->octree_max_depth = 20;
    p->variant = ICP_VARIANT_ENGINE;
}

int icp_create(icp_handle* out, int device_id) {
    if (!out) return ICP_INVALID_ARGUMENT;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device_id < 0 || device_id >= n_dev) {
        cudaGetLastError();
        return ICP_CUDA_ERROR;  // no CPU fallback: without a CUDA device the library refuses to work
    }
    Ctx* c = new (std::nothrow) Ctx();
    if (!c) return ICP_CUDA_ERROR;
    c->device = device_id;
    icp_default_params(&c->params);
    if (cudaSetDevice(device_id) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    if (cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    if (cudaEventCreateWithFlags(&c->ev_src, cudaEventDisableTiming) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    {
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        if (cudaStreamCreateWithPriority(&c->stream_hi, cudaStreamNonBlocking, hi_prio) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
        for (auto& e : c->ev_chunk)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    }
    for (auto& e : c->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    if (cudaMalloc(&c->d_state, sizeof(LoopState)) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    if (cudaMalloc(&c->d_counters, 16 * sizeof(unsigned long long)) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    cudaMemset(c->d_counters, 0, 16 * sizeof(unsigned long long));
    if (cudaMalloc(&c->d_work_count, 16 * sizeof(unsigned int)) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    if (cudaHostAlloc(&c->h_rec, Ctx::REC_RING * sizeof(IterRecord), cudaHostAllocMapped) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    for (auto& e : c->ev_it)
        if (cudaEventCreate(&e) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    if (cudaHostGetDevicePointer(&c->d_rec, c->h_rec, 0) != cudaSuccess) { delete c; return ICP_CUDA_ERROR; }
    if (getenv("ICP_B200_DEBUG_COUNTERS")) c->opt_count = true;
    const char* m = getenv("ICP_B200_NN_MODE");
    if (m && (atoi(m) == 0 || (atoi(m) >= 3 && atoi(m) <= 6))) c->opt_nn_mode = atoi(m);
    *out = (icp_handle)c;
    return ICP_OK;
}

void icp_destroy(icp_handle h) {
    Ctx* c = (Ctx*)h;
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    icp_comm_destroy(h);
    for (Ctx* w : c->workers) icp_destroy((icp_handle)w);
    c->workers.clear();
    octree_free(c);
    DevBuf* bufs[] = {&c->tgt_raw, &c->sx, &c->sy, &c->sz, &c->sperm, &c->pos, &c->dist, &c->mask, &c->part_a, &c->part_b,
                      &c->scratch0, &c->scratch1, &c->scratch2, &c->scratch3, &c->scratch_src, &c->gather_a, &c->gather_b, &c->node_io, &c->las_src, &c->las_tgt, &c->lb, &c->cand, &c->work2, &c->rd_perm, &c->rd_send, &c->rd_recv, &c->rd_tmp, &c->rd_back, &c->scratch_keys};
    for (DevBuf* b : bufs) devbuf_free(*b);
    if (c->pin_a.p) cudaFreeHost(c->pin_a.p);
    if (c->pin_b.p) cudaFreeHost(c->pin_b.p);
    if (c->d_state) cudaFree(c->d_state);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->d_work_count) cudaFree(c->d_work_count);
    if (c->h_rec) cudaFreeHost(c->h_rec);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_it)
        if (e) cudaEventDestroy(e);
    if (c->ev_src) cudaEventDestroy(c->ev_src);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream_hi) cudaStreamDestroy(c->stream_hi);
    for (auto& e : c->ev_chunk)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c->nccl;
    delete c;
}

const char* icp_last_error(icp_handle h) {
    Ctx* c = (Ctx*)h;
    return c ? c->err.c_str() : "null handle";
}

int icp_set_params(icp_handle h, const icp_params* p) {
    Ctx* c = (Ctx*)h;
    if (!c || !p) return ICP_INVALID_ARGUMENT;
    if (p->variant != ICP_VARIANT_ENGINE && p->variant != ICP_VARIANT_CLI) {
        c->err = "params: unknown variant";
        return ICP_INVALID_ARGUMENT;
    }
    if (p->octree_max_depth < 0 || p->octree_max_depth > 63) {  // (the reference's GUI offers 10 .. 50, settingspage.cpp:76)
        c->err = "params: octree_max_depth must be in [0,63]";
        return ICP_INVALID_ARGUMENT;
    }
    c->params = *p;
    return ICP_OK;
}

int icp_get_params(icp_handle h, icp_params* p) {
    Ctx* c = (Ctx*)h;
    if (!c || !p) return ICP_INVALID_ARGUMENT;
    *p = c->params;
    return ICP_OK;
}

int icp_set_callbacks(icp_handle h, icp_iteration_cb on_iteration, icp_progress_cb on_progress, icp_log_cb on_log, void* user) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    c->on_iteration = on_iteration;
    c->on_progress = on_progress;
    c->on_log = on_log;
    c->user = user;
    return ICP_OK;
}

int icp_set_option(icp_handle h, const char* key, double value) {
    Ctx* c = (Ctx*)h;
    if (!c || !key) return ICP_INVALID_ARGUMENT;
    if (!strcmp(key, "nn_mode")) {
        const int m = (int)(value + 0.5);
        if (!(m == 0 || (m >= 3 && m <= 6))) {
            c->err = "nn_mode: 0 literal traversal, 3 per-thread cell walk, 4 balanced walk, 5 keep / collect, 6 automatic (default)";
            return ICP_INVALID_ARGUMENT;
        }
        c->opt_nn_mode = m;
        c->prev_valid = false;
    }
    else if (!strcmp(key, "count")) c->opt_count = value != 0.0;
    else if (!strcmp(key, "batch_small")) c->opt_batch_small = value != 0.0;
    else if (!strcmp(key, "batch_workers")) c->opt_batch_workers = std::min(std::max((int)value, 1), 64);
    else if (!strcmp(key, "base_occupancy")) c->opt_base_occupancy = std::max(value, 1.0);
    else if (!strcmp(key, "grid_levels")) c->opt_grid_levels = std::min(std::max((int)value, 1), 4);
    else if (!strcmp(key, "range_max")) c->opt_range_max = std::max((int)value, 0);
    else if (!strcmp(key, "walk_bias")) c->opt_walk_bias = (int)value;
    else if (!strcmp(key, "search_depth")) c->opt_search_depth = std::min(std::max((int)value, 0), 21);
    else if (!strcmp(key, "grid_coarse")) c->opt_grid_coarse = std::min(std::max((int)value, 0), 3);
    else if (!strcmp(key, "grid_shift")) c->opt_grid_shift = (int)value;
    else if (!strcmp(key, "grid_max_cells")) c->opt_grid_max_cells = std::max((long long)value, 1ll);
    else if (!strcmp(key, "walk_max_cells")) c->opt_walk_max_cells = std::max((int)value, 1);
    else if (!strcmp(key, "search_leaf")) c->opt_search_leaf = std::min(std::max((int)value, 1), 1024);
    else if (!strcmp(key, "order_queries")) c->opt_order_queries = value != 0.0;
    else if (!strcmp(key, "nn_chunks")) c->opt_nn_chunks = std::min(std::max((int)value, 1), 8);
    else if (!strcmp(key, "temporal_skip")) c->opt_temporal_skip = value != 0.0;
    else if (!strcmp(key, "keep_k")) { c->opt_keep_k = std::min(std::max((int)value, 1), 4); c->keep_valid = false; }
    else if (!strcmp(key, "keep_alpha")) c->opt_keep_alpha = std::min(std::max(value, 1.0), 16.0);
    else if (!strcmp(key, "keep_enter")) c->opt_keep_enter = std::max(value, 0.0);
    else if (!strcmp(key, "keep_exit")) c->opt_keep_exit = std::max(value, 0.0);
    else if (!strcmp(key, "keep_rcap")) c->opt_keep_rcap = std::min(std::max(value, 0.0), 8.0);
    else if (!strcmp(key, "keep_bias")) c->opt_keep_bias = std::min(std::max((int)value, -4), 4);
    else if (!strcmp(key, "write_mask")) c->opt_write_mask = value != 0.0;
    else if (!strcmp(key, "redistribute")) c->opt_redistribute = value != 0.0;
    else if (!strcmp(key, "shard_target")) c->opt_shard_target = value != 0.0;
    else if (!strcmp(key, "lookahead")) c->opt_lookahead = std::min(std::max((int)value, 1), Ctx::REC_RING);
    else {
        c->err = std::string("unknown option ") + key;
        return ICP_INVALID_ARGUMENT;
    }
    return ICP_OK;
}

int icp_nn_counters(icp_handle h, int64_t* fast_path, int64_t* literal_fallback, int reset) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    unsigned long long v[4] = {0, 0, 0, 0};
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    ICPB_CUDA(c, cudaMemcpy(v, c->d_counters, sizeof v, cudaMemcpyDeviceToHost));
    if (fast_path) *fast_path = (int64_t)v[0];
    if (literal_fallback) *literal_fallback = (int64_t)v[1];
    if (reset) ICPB_CUDA(c, cudaMemset(c->d_counters, 0, 2 * sizeof(unsigned long long)));
    return ICP_OK;
}

int icp_nn_tile_counters(icp_handle h, int64_t* per_thread_lanes, int64_t* candidates_scanned, int reset) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    unsigned long long v[4] = {0, 0, 0, 0};
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    ICPB_CUDA(c, cudaMemcpy(v, c->d_counters, sizeof v, cudaMemcpyDeviceToHost));
    if (per_thread_lanes) *per_thread_lanes = (int64_t)v[2];
    if (candidates_scanned) *candidates_scanned = (int64_t)v[3];
    if (getenv("ICP_B200_DEBUG_COUNTERS")) {
        unsigned long long w[8];
        ICPB_CUDA(c, cudaMemcpy(w, c->d_counters, sizeof w, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[icp_b200] tile counters: slow_lanes=%llu candidates=%llu rounds(mode 4: scan items)=%llu passes(mode 4: matches kept without a search)=%llu nodes=%llu start_steps=%llu\n", w[2], w[3],
                w[4], w[5], w[6], w[7]);
    }
    if (reset) ICPB_CUDA(c, cudaMemset(c->d_counters + 2, 0, 6 * sizeof(unsigned long long)));
    return ICP_OK;
}

int64_t icp_kernel_launches(icp_handle h) {
    Ctx* c = (Ctx*)h;
    return c ? c->launches : 0;
}

int icp_register(icp_handle h, double* src_xyz, int64_t n_src, const double* tgt_xyz, int64_t n_tgt, icp_result* out,
                 const volatile int* stop_flag) {
    Ctx* c = (Ctx*)h;
    if (!c || !out) return ICP_INVALID_ARGUMENT;
    if (c->n_ranks > 1) {
        c->err = "icp_register on a handle with an initialised communicator: use icp_register_sharded";
        return ICP_INVALID_ARGUMENT;
    }
    return register_impl(c, src_xyz, n_src, n_src, tgt_xyz, n_tgt, out, stop_flag);
}

int icp_register_las(icp_handle h, const icp_las_points* src, const icp_las_points* tgt, icp_result* out, double* src_out_xyz,
                     const volatile int* stop_flag) {
    Ctx* c = (Ctx*)h;
    if (!c || !out) return ICP_INVALID_ARGUMENT;
    if (c->n_ranks > 1) {
        c->err = "icp_register_las on a handle with an initialised communicator";
        return ICP_INVALID_ARGUMENT;
    }
    if (!src || !tgt || !src->records || !tgt->records || src->n <= 0 || tgt->n <= 0) {  // icpengine.cpp:26-34
        init_result(out);
        out->status = ICP_EMPTY_INPUT;
        return ICP_EMPTY_INPUT;
    }
    if (src->record_length < 12 || tgt->record_length < 12) return ICP_INVALID_ARGUMENT;
    return register_impl(c, src_out_xyz, src->n, src->n, nullptr, tgt->n, out, stop_flag, src, tgt);
}

int icp_register_sharded(icp_handle h, double* src_shard_xyz, int64_t n_shard, int64_t n_src_global, const double* tgt_xyz,
                         int64_t n_tgt, icp_result* out, const volatile int* stop_flag) {
    Ctx* c = (Ctx*)h;
    if (!c || !out) return ICP_INVALID_ARGUMENT;
    return register_impl(c, src_shard_xyz, n_shard, n_src_global, tgt_xyz, n_tgt, out, stop_flag);
}

int icp_octree_build(icp_handle h, const double* tgt_xyz, int64_t n_tgt, int max_points, int max_depth) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    if (!tgt_xyz || n_tgt <= 0) return ICP_EMPTY_INPUT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload(c, c->tgt_raw, tgt_xyz, n_tgt));
    c->n_tgt = n_tgt;
    ICPB_CUDA(c, cudaEventRecord(c->ev[5], c->stream));
    ICPB_TRY(octree_build_device(c, (const double*)c->tgt_raw.p, n_tgt, max_points, max_depth));
    ICPB_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->last_build_ms, c->ev[5], c->ev[6]);
    return ICP_OK;
}

int icp_octree_get_info(icp_handle h, icp_octree_info* info) {
    Ctx* c = (Ctx*)h;
    if (!c || !info) return ICP_INVALID_ARGUMENT;
    if (!c->tree.valid) return ICP_NO_OCTREE;
    const DeviceOctree& t = c->tree;
    std::memset(info, 0, sizeof *info);
    info->n_points = t.n_pts;
    info->n_nodes = t.n_nodes;
    info->n_leaves = t.n_leaves;
    info->node_bytes = t.n_nodes * (int64_t)sizeof(Node);
    info->point_bytes = t.n_pts * (int64_t)sizeof(TPoint);
    info->depth = t.depth;
    info->max_points = t.max_pts;
    info->max_depth = t.max_depth;
    for (int a = 0; a < 3; ++a) {
        info->root_lo[a] = t.root_lo[a];
        info->root_hi[a] = t.root_hi[a];
    }
    info->build_ms = c->last_build_ms;
    const DeviceOctree& f = c->fast;
    info->search_nodes = f.n_nodes;
    info->search_node_bytes = f.n_nodes * (int64_t)sizeof(Node);
    info->search_depth = f.depth;
    info->grid_base_level = f.glev_min + f.gbase;
    info->grid_fine_level = f.glev_min + f.glev_n - 1;
    if (f.glev_n > 0) {
        const int k = f.glev_n - 1;
        info->grid_bytes = (f.goff[k] + (int64_t)f.gdim[k][0] * f.gdim[k][1] * f.gdim[k][2]) * (int64_t)sizeof(uint2);
        info->grid_base_cell = f.cube / (double)(1ll << (f.glev_min + f.gbase));
    }
    return ICP_OK;
}

int icp_octree_dump(icp_handle h, int64_t* n_nodes_out, int64_t* n_leaf_points_out, int32_t* depth, uint64_t* key,
                    uint8_t* is_leaf, int32_t* count, double* box6, int32_t* leaf_idx) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    if (!c->tree.valid) return ICP_NO_OCTREE;
    const DeviceOctree& t = c->tree;
    if (n_nodes_out) *n_nodes_out = t.n_nodes;
    if (n_leaf_points_out) *n_leaf_points_out = t.n_pts;
    if (!depth) return ICP_OK;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    std::vector<Node> nodes((size_t)t.n_nodes);
    std::vector<TPoint> pts((size_t)t.n_pts);
    ICPB_CUDA(c, cudaMemcpy(nodes.data(), t.nodes, nodes.size() * sizeof(Node), cudaMemcpyDeviceToHost));
    ICPB_CUDA(c, cudaMemcpy(pts.data(), t.pts, pts.size() * sizeof(TPoint), cudaMemcpyDeviceToHost));
    // pre-order walk, children in octant order; leaf indices ascending (the reference's leaf lists are)
    struct Item { uint32_t node; uint64_t key; };
    std::vector<Item> stack;
    stack.push_back({0u, 0ull});
    int64_t slot = 0, nidx = 0;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        const Node& nd = nodes[it.node];
        const uint32_t mask = nd.meta & 0xFFu;
        depth[slot] = (int32_t)((nd.meta >> 8) & 0xFFu);
        key[slot] = it.key;
        is_leaf[slot] = mask == 0 ? 1 : 0;
        count[slot] = mask == 0 ? (int32_t)nd.npts : 0;
        double* b = box6 + 6 * slot;
        b[0] = nd.lo[0]; b[1] = nd.hi[0]; b[2] = nd.lo[1]; b[3] = nd.hi[1]; b[4] = nd.lo[2]; b[5] = nd.hi[2];
        ++slot;
        if (mask == 0) {
            for (uint32_t k = 0; k < nd.npts; ++k) leaf_idx[nidx + k] = (int32_t)pts[nd.pt0 + k].idx;
            std::sort(leaf_idx + nidx, leaf_idx + nidx + nd.npts);
            nidx += nd.npts;
        } else {
            int nch = __builtin_popcount(mask);
            for (int o = 7; o >= 0; --o)
                if ((mask >> o) & 1u) {
                    --nch;
                    stack.push_back({nd.child0 + (uint32_t)nch, (it.key << 3) | (uint64_t)o});
                }
        }
    }
    return ICP_OK;
}

int icp_nn_query(icp_handle h, const double* q_xyz, int64_t n, int32_t* idx_out, double* dist_out, float* kernel_ms) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    if (!c->tree.valid) return ICP_NO_OCTREE;
    if (n <= 0) return ICP_OK;
    if (!q_xyz || !idx_out) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload(c, c->scratch_src, q_xyz, n));
    c->src_identity_perm = false;
    ICPB_TRY(source_from_device_aos(c, (const double*)c->scratch_src.p, n));
    ICPB_TRY(ensure_run_buffers(c, n));
    NNLaunch L;
    L.sx = (double*)c->sx.p; L.sy = (double*)c->sy.p; L.sz = (double*)c->sz.p;
    L.ox = L.oy = L.oz = nullptr;
    L.n = n;
    L.pos_out = (uint32_t*)c->pos.p;
    L.dist_out = (double*)c->dist.p;
    L.prev_pos = nullptr;
    L.node_io = nullptr;
    L.part_a = nullptr;
    L.state = nullptr;
    L.apply_pending = 0;
    L.mode = c->opt_nn_mode;
    L.init_best = (c->params.variant == ICP_VARIANT_CLI) ? 1e20 : DBL_MAX;
    ICPB_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    ICPB_TRY(nn_launch(c, L));
    ICPB_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    // results back in caller order
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)n * (sizeof(int32_t) + sizeof(double))));
    double* d_dist = (double*)c->scratch1.p;
    int32_t* d_idx = (int32_t*)(d_dist + n);
    ICPB_TRY(unsort_results_launch(c, L.pos_out, L.dist_out, c->src_identity_perm ? nullptr : (uint32_t*)c->sperm.p, n, d_idx, d_dist));
    ICPB_CUDA(c, cudaMemcpyAsync(idx_out, d_idx, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (dist_out) ICPB_CUDA(c, cudaMemcpyAsync(dist_out, d_dist, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (kernel_ms) cudaEventElapsedTime(kernel_ms, c->ev[0], c->ev[1]);
    return ICP_OK;
}

int icp_iteration_stats(icp_handle h, const double* src_xyz, int64_t n, const int32_t* idx, int iteration, double* dist_out,
                        uint8_t* inlier_mask_out, icp_stats* stats_out) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    if (!c->tree.valid) return ICP_NO_OCTREE;
    if (!src_xyz || !idx || n <= 0) return ICP_EMPTY_INPUT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(build_inv_perm(c));
    c->prev_valid = false;
    ICPB_TRY(upload(c, c->scratch_src, src_xyz, n));
    ICPB_TRY(ensure_source_buffers(c, n));
    ICPB_TRY(ensure_run_buffers(c, n));
    c->n_src = n;
    c->src_identity_perm = true;
    double *sx = (double*)c->sx.p, *sy = (double*)c->sy.p, *sz = (double*)c->sz.p;
    ICPB_TRY(aos_to_soa_launch(c, (const double*)c->scratch_src.p, n, sx, sy, sz));
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)n * sizeof(int32_t)));
    ICPB_CUDA(c, cudaMemcpyAsync(c->scratch1.p, idx, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    LoopState hs;
    std::memset(&hs, 0, sizeof hs);
    hs.prev_error = 1e10;
    identity16(hs.T_pending); identity16(hs.T_last); identity16(hs.T_cum);
    for (int a = 0; a < 3; ++a) hs.pivot_a[a] = hs.pivot_b[a] = 0.5 * (c->tree.root_lo[a] + c->tree.root_hi[a]);
    hs.tolerance = c->params.tolerance;
    hs.sigma = (c->params.variant == ICP_VARIANT_CLI) ? 3.0 : c->params.sigma_multiplier;
    hs.variant = c->params.variant;
    hs.max_iterations = c->params.max_iterations;
    hs.n_global = n;
    ICPB_CUDA(c, cudaMemcpyAsync(c->d_state, &hs, sizeof hs, cudaMemcpyHostToDevice, c->stream));
    StatA* part_a = (StatA*)c->part_a.p + 64;
    StatA* rank_a = (StatA*)c->gather_a.p;
    double* rank_b = (double*)c->gather_b.p;
    int n_part = 0;
    ICPB_TRY(dist_from_idx_launch(c, sx, sy, sz, (const int32_t*)c->scratch1.p, n, (uint32_t*)c->pos.p, (double*)c->dist.p, part_a,
                                  &n_part));
    ICPB_TRY(stage_a_reduce_launch(c, part_a, n_part, rank_a));
    const int keep_ranks = c->n_ranks, keep_rank = c->rank;
    c->n_ranks = 1;  // the stage API is per handle: no exchange
    c->rank = 0;
    int sb = stage_b_launch(c, sx, sy, sz, (uint32_t*)c->pos.p, (double*)c->dist.p, n, iteration, rank_a, (uint8_t*)c->mask.p,
                            (double*)c->part_b.p, rank_b);
    c->n_ranks = keep_ranks;
    c->rank = keep_rank;
    ICPB_TRY(sb);
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    const IterRecord rec = *c->h_rec;
    if (dist_out) ICPB_CUDA(c, cudaMemcpy(dist_out, c->dist.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    if (inlier_mask_out) ICPB_CUDA(c, cudaMemcpy(inlier_mask_out, c->mask.p, (size_t)n, cudaMemcpyDeviceToHost));
    if (stats_out) {
        double b17[STATB_DOUBLES];
        ICPB_CUDA(c, cudaMemcpy(b17, rank_b, sizeof b17, cudaMemcpyDeviceToHost));
        stats_out->min_distance = rec.dmin;
        stats_out->max_distance = rec.dmax;
        stats_out->mean = rec.mean;
        stats_out->std_dev = rec.std_dev;
        stats_out->threshold = rec.threshold;
        stats_out->rmse = rec.rmse;
        stats_out->sum_sq = b17[1];
        stats_out->problem_count = (int64_t)rec.problems;
        stats_out->valid_count = rec.valid_points;
        stats_out->outlier_count = rec.outlier_points;
    }
    return ICP_OK;
}

int icp_best_fit_transform(icp_handle h, const double* a_xyz, const double* b_xyz, int64_t n, double* T_out) {
    Ctx* c = (Ctx*)h;
    if (!c || !T_out) return ICP_INVALID_ARGUMENT;
    if (!a_xyz || !b_xyz || n <= 0) return ICP_EMPTY_INPUT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)n * 6 * sizeof(double) + 1024));
    double* da = (double*)c->scratch1.p;
    double* db = da + 3 * n;
    ICPB_CUDA(c, cudaMemcpyAsync(da, a_xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ICPB_CUDA(c, cudaMemcpyAsync(db, b_xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ICPB_TRY(devbuf_reserve(c, c->part_b, (size_t)(stage_b_blocks(c, n) + 8) * STATB_DOUBLES * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->gather_b, (size_t)(STATB_DOUBLES + 16) * sizeof(double) + 64));
    double* sums = (double*)c->gather_b.p;
    ICPB_TRY(pairs_b_launch(c, da, db, n, (double*)c->part_b.p, sums));
    ICPB_TRY(bestfit_launch(c, sums, da, db, sums + STATB_DOUBLES));
    ICPB_CUDA(c, cudaMemcpyAsync(T_out, sums + STATB_DOUBLES, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

int icp_solve_from_H(icp_handle h, const double* H9, const double* cA3, const double* cB3, double* T_out, double* U9, double* S3,
                     double* V9) {
    Ctx* c = (Ctx*)h;
    if (!c || !H9 || !cA3 || !cB3 || !T_out) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(devbuf_reserve(c, c->scratch0, 64 * sizeof(double)));
    double in[15], outv[37];
    std::memcpy(in, H9, 9 * sizeof(double));
    std::memcpy(in + 9, cA3, 3 * sizeof(double));
    std::memcpy(in + 12, cB3, 3 * sizeof(double));
    double* d = (double*)c->scratch0.p;
    ICPB_CUDA(c, cudaMemcpyAsync(d, in, sizeof in, cudaMemcpyHostToDevice, c->stream));
    ICPB_TRY(solve_from_H_launch(c, d, d + 16));
    ICPB_CUDA(c, cudaMemcpyAsync(outv, d + 16, sizeof outv, cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    std::memcpy(T_out, outv, 16 * sizeof(double));
    if (U9) std::memcpy(U9, outv + 16, 9 * sizeof(double));
    if (S3) std::memcpy(S3, outv + 25, 3 * sizeof(double));
    if (V9) std::memcpy(V9, outv + 28, 9 * sizeof(double));
    return ICP_OK;
}

int icp_apply_transform(icp_handle h, const double* T16, double* xyz, int64_t n) {
    Ctx* c = (Ctx*)h;
    if (!c || !T16) return ICP_INVALID_ARGUMENT;
    if (n <= 0) return ICP_OK;
    if (!xyz) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload(c, c->scratch_src, xyz, n));
    ICPB_TRY(devbuf_reserve(c, c->scratch0, 64 * sizeof(double)));
    ICPB_CUDA(c, cudaMemcpyAsync(c->scratch0.p, T16, 16 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ICPB_TRY(apply_aos_launch(c, (const double*)c->scratch0.p, (double*)c->scratch_src.p, n));
    ICPB_CUDA(c, cudaMemcpyAsync(xyz, c->scratch_src.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

int icp_source_upload(icp_handle h, const double* src_xyz, int64_t n_src) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    if (c->n_ranks > 1 ? (n_src < 0 || (n_src > 0 && !src_xyz)) : (!src_xyz || n_src <= 0)) return ICP_EMPTY_INPUT;  // (empty shards allowed)
    ICPB_CUDA(c, cudaSetDevice(c->device));
    if (n_src > 0) ICPB_TRY(upload(c, c->scratch_src, src_xyz, n_src));
    c->src_identity_perm = false;
    ICPB_TRY(source_from_shard(c, (const double*)c->scratch_src.p, n_src));  // (collective on a sharded handle)
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

int icp_register_resident(icp_handle h, int64_t n_src_global, icp_result* out, double* src_out_xyz, const volatile int* stop_flag) {
    Ctx* c = (Ctx*)h;
    if (!c || !out) return ICP_INVALID_ARGUMENT;
    init_result(out);
    if (!c->tree.valid) return ICP_NO_OCTREE;
    if (c->n_src <= 0 && c->n_ranks <= 1) return ICP_EMPTY_INPUT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    bool write_back = true;
    ICPB_TRY(run_loop(c, n_src_global > 0 ? n_src_global : c->n_src, out, stop_flag, &write_back));
    if (write_back && src_out_xyz && (c->rd_active || c->n_src > 0))
        ICPB_TRY(write_back_source(c, src_out_xyz, c->rd_active ? c->rd_n_in : c->n_src));
    return out->status;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------
int icp_comm_unique_id(icp_handle h, void* unique_id_128) {
    Ctx* c = (Ctx*)h;
    if (!c || !unique_id_128) return ICP_INVALID_ARGUMENT;
    NcclApi* a = nccl_load(c);
    if (!a) return ICP_NCCL_ERROR;
    ncclUniqueId id;
    ICPB_NCCL(c, a->GetUniqueId(&id));
    std::memcpy(unique_id_128, &id, sizeof id);
    return ICP_OK;
}

static void mailbox_close(Ctx* c) {
    for (int r = 0; r < MAIL_RANKS; ++r) {
        if (c->peer_mail[r] && c->peer_mail[r] != c->mail) cudaIpcCloseMemHandle(c->peer_mail[r]);
        c->peer_mail[r] = nullptr;
    }
    if (c->mail) cudaFree(c->mail);
    c->mail = nullptr;
    c->p2p = false;
}

// One mailbox per rank, opened by every peer through CUDA IPC; the 64-byte handles travel through the NCCL communicator that
// was just created.  Every rank then reports whether it could open all of them: the mailboxes are used only if all could
// (otherwise every rank keeps the two NCCL all-gathers per iteration).
static int mailbox_open(Ctx* c, NcclApi* a) {
    mailbox_close(c);
    const int n = c->n_ranks;
    ICPB_CUDA(c, cudaMalloc(&c->mail, sizeof(Mailbox)));
    ICPB_CUDA(c, cudaMemset(c->mail, 0, sizeof(Mailbox)));
    c->mail_epoch = 0u;
    cudaIpcMemHandle_t mine;
    bool ok = cudaIpcGetMemHandle(&mine, c->mail) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        std::memset(&mine, 0, sizeof mine);
    }
    unsigned char* d_h = nullptr;
    ICPB_CUDA(c, cudaMalloc(&d_h, (size_t)n * sizeof mine + (size_t)n));
    ICPB_CUDA(c, cudaMemcpyAsync(d_h + (size_t)c->rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    ICPB_NCCL(c, a->AllGather(d_h + (size_t)c->rank * sizeof mine, d_h, sizeof mine, ncclUint8, (ncclComm_t)c->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all(n);
    ICPB_CUDA(c, cudaMemcpyAsync(all.data(), d_h, (size_t)n * sizeof mine, cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < n && ok; ++r) {
        if (r == c->rank) {
            c->peer_mail[r] = c->mail;
            continue;
        }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
        } else {
            c->peer_mail[r] = (Mailbox*)p;
        }
    }
    // agree: one byte per rank
    unsigned char* d_ok = d_h + (size_t)n * sizeof mine;
    const unsigned char flag = ok ? 1 : 0;
    ICPB_CUDA(c, cudaMemcpyAsync(d_ok + c->rank, &flag, 1, cudaMemcpyHostToDevice, c->stream));
    ICPB_NCCL(c, a->AllGather(d_ok + c->rank, d_ok, 1, ncclUint8, (ncclComm_t)c->comm, c->stream));
    std::vector<unsigned char> oks(n);
    ICPB_CUDA(c, cudaMemcpyAsync(oks.data(), d_ok, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_h);
    bool all_ok = true;
    for (int r = 0; r < n; ++r) all_ok = all_ok && oks[r] != 0;
    if (!all_ok) {
        log_msg(c, "peer mailboxes unavailable on some rank: per-iteration records go through NCCL all-gathers");
        mailbox_close(c);
        return ICP_OK;
    }
    c->p2p = true;
    return ICP_OK;
}

int icp_comm_init(icp_handle h, int rank, int n_ranks, const void* unique_id_128) {
    Ctx* c = (Ctx*)h;
    if (!c || !unique_id_128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ICP_INVALID_ARGUMENT;
    NcclApi* a = nccl_load(c);
    if (!a) return ICP_NCCL_ERROR;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id_128, sizeof id);
    ncclComm_t comm = nullptr;
    ICPB_NCCL(c, a->CommInitRank(&comm, n_ranks, id, rank));
    c->comm = comm;
    c->rank = rank;
    c->n_ranks = n_ranks;
    c->p2p = false;
    if (n_ranks > 1 && n_ranks <= MAIL_RANKS && !getenv("ICP_B200_NO_P2P")) ICPB_TRY(mailbox_open(c, a));
    return ICP_OK;
}

int icp_comm_destroy(icp_handle h) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    mailbox_close(c);
    if (c->comm && c->nccl) c->nccl->CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    c->rank = 0;
    c->n_ranks = 1;
    return ICP_OK;
}

// ---- batch (BASELINE.json config #5) -------------------------------------------------------------------
// Many small independent registrations: each is launch- and latency-bound on its own (a 2k-point pair keeps a
// fraction of one SM busy), so the batch is spread over a pool of worker handles -- one CUDA stream each on the
// same device, driven by one host thread each -- and the GPU overlaps their kernels.  Every pair goes through the
// ordinary single-pair path (register_impl), so batch results are the single-pair results by construction.
// everything icp_set_params / icp_set_option can change, from the handle that owns a batch to one of its workers
static void copy_options(Ctx* w, const Ctx* c) {
    w->params = c->params;
    w->opt_nn_mode = c->opt_nn_mode;
    w->opt_order_queries = c->opt_order_queries;
    w->opt_write_mask = c->opt_write_mask;
    w->opt_count = c->opt_count;
    w->opt_temporal_skip = c->opt_temporal_skip;
    w->opt_nn_chunks = c->opt_nn_chunks;
    w->opt_keep_k = c->opt_keep_k;
    w->opt_keep_alpha = c->opt_keep_alpha;
    w->opt_keep_bias = c->opt_keep_bias;
    w->opt_keep_enter = c->opt_keep_enter;
    w->opt_keep_exit = c->opt_keep_exit;
    w->opt_keep_rcap = c->opt_keep_rcap;
    w->opt_search_leaf = c->opt_search_leaf;
    w->opt_search_depth = c->opt_search_depth;
    w->opt_grid_shift = c->opt_grid_shift;
    w->opt_grid_max_cells = c->opt_grid_max_cells;
    w->opt_grid_levels = c->opt_grid_levels;
    w->opt_grid_coarse = c->opt_grid_coarse;
    w->opt_base_occupancy = c->opt_base_occupancy;
    w->opt_range_max = c->opt_range_max;
    w->opt_walk_bias = c->opt_walk_bias;
    w->opt_walk_max_cells = c->opt_walk_max_cells;
    w->opt_lookahead = c->opt_lookahead;
}

int icp_register_batch(icp_handle h, int32_t n_pairs, double* const* src_xyz, const int64_t* n_src, const double* const* tgt_xyz,
                       const int64_t* n_tgt, icp_result* results) {
    Ctx* c = (Ctx*)h;
    if (!c || n_pairs < 0 || (n_pairs > 0 && (!src_xyz || !n_src || !tgt_xyz || !n_tgt || !results))) return ICP_INVALID_ARGUMENT;
    if (n_pairs == 0) return ICP_OK;
    auto ensure_workers = [&](int want) -> int {
    while ((int)c->workers.size() < want) {
        icp_handle w = nullptr;
        if (icp_create(&w, c->device) != ICP_OK) {
            c->err = "batch: could not create a worker handle";
            return ICP_CUDA_ERROR;
        }
        c->workers.push_back((Ctx*)w);
    }
    return ICP_OK;
    };
    // ---- small pairs: one thread block each, all in one launch (batch.cu) ---------------------------------------------
    std::vector<int32_t> todo;  // pairs left for the general path
    std::atomic<int> fatal{ICP_OK}, worst{ICP_OK};
    {
        std::vector<int32_t> small;
        for (int32_t p = 0; p < n_pairs; ++p) {
            if (c->opt_batch_small && src_xyz[p] && tgt_xyz[p] && small_pair_eligible(n_src[p], n_tgt[p]) && !c->on_iteration &&
                !c->on_progress && !c->on_log)
                small.push_back(p);
            else
                todo.push_back(p);
        }
        if (!small.empty()) {
            // chunks of a few waves of blocks, spread over up to three lanes (this handle + two workers, one host
            // thread and one stream each): while one lane's kernel runs, the others pack / copy / unpack
            const int rec_cap = std::max(c->params.max_iterations, 0) + 1;
            const int chunk = std::max(c->sm_count * 2, 64);
            const int n_chunks = ((int)small.size() + chunk - 1) / chunk;
            const int lanes = std::min(3, n_chunks);
            if (lanes > 1) ICPB_TRY(ensure_workers(lanes - 1));
            std::vector<std::vector<int32_t>> redo((size_t)lanes);
            std::atomic<int> next_chunk{0}, fatal_lane{-1};
            auto lane_fn = [&](int lane) {
                Ctx* w = (lane == 0) ? c : c->workers[(size_t)lane - 1];
                cudaSetDevice(w->device);
                if (w != c) w->params = c->params;
                std::vector<IterRecord> recs;
                std::vector<int> n_rec, exit_code;
                std::vector<char> flagged;
                std::vector<long long> src_off;
                for (;;) {
                    const int ci = next_chunk.fetch_add(1);
                    if (ci >= n_chunks || fatal.load() != ICP_OK) break;
                    const std::vector<int32_t> part(small.begin() + (size_t)ci * chunk,
                                                    small.begin() + std::min(small.size(), (size_t)(ci + 1) * chunk));
                    const double* moved = nullptr;
                    const int s = small_batch_run(w, part, src_xyz, n_src, tgt_xyz, n_tgt, recs, rec_cap, n_rec, exit_code, flagged, moved,
                                                  src_off);
                    if (s != ICP_OK) {
                        int none = ICP_OK;
                        if (fatal.compare_exchange_strong(none, s)) fatal_lane.store(lane);  // the first failure reports
                        break;
                    }
                    for (size_t k = 0; k < part.size(); ++k) {
                        const int32_t p = part[k];
                        if (flagged[k]) {
                            redo[(size_t)lane].push_back(p);
                            continue;
                        }
                        init_result(&results[p]);
                        RunAcc acc(w, &results[p], c->params.variant, c->params.max_iterations, (long long)n_src[p]);
                        for (int r = 0; r < n_rec[k]; ++r)
                            if (!acc.consume(recs[k * (size_t)rec_cap + (size_t)r], r, 0.f, 0.f, false)) break;
                        acc.finish();
                        if (acc.write_back)  // icpengine.cpp:371-375
                            std::memcpy(src_xyz[p], moved + 3 * src_off[k], (size_t)n_src[p] * 3 * sizeof(double));
                        int expect = ICP_OK;
                        if (results[p].status != ICP_OK) worst.compare_exchange_strong(expect, results[p].status);
                    }
                }
            };
            std::vector<std::thread> lane_threads;
            for (int l = 1; l < lanes; ++l) lane_threads.emplace_back(lane_fn, l);
            lane_fn(0);
            for (auto& t : lane_threads) t.join();
            if (fatal.load() != ICP_OK) {
                const int fl = fatal_lane.load();
                if (fl > 0) c->err = c->workers[(size_t)fl - 1]->err;  // (lane 0 is this handle: its message is already here)
                return fatal.load();
            }
            for (auto& r : redo) todo.insert(todo.end(), r.begin(), r.end());
        }
    }
    if (todo.empty()) return worst.load();
    const int n_todo = (int)todo.size();
    const int want = std::max(1, std::min(std::min(c->opt_batch_workers, n_todo), 64));
    ICPB_TRY(ensure_workers(want));
    std::atomic<int32_t> next{0};
    std::atomic<int> fatal_worker{-1};

    auto run = [&](Ctx* w, int wi) {
        cudaSetDevice(w->device);
        copy_options(w, c);  // a pair that takes the general path runs with this handle's tuning
        for (;;) {
            const int32_t q = next.fetch_add(1);
            if (q >= n_todo || fatal.load() != ICP_OK) break;
            const int32_t p = todo[(size_t)q];
            const int s = register_impl(w, src_xyz[p], n_src[p], n_src[p], tgt_xyz[p], n_tgt[p], &results[p], nullptr);
            if (s == ICP_CUDA_ERROR || s == ICP_NCCL_ERROR) {
                int none = ICP_OK;
                if (fatal.compare_exchange_strong(none, s)) fatal_worker.store(wi);
                break;
            }
            int expect = ICP_OK;
            if (s != ICP_OK) worst.compare_exchange_strong(expect, s);
        }
    };
    std::vector<std::thread> threads;
    for (int k = 1; k < want; ++k) threads.emplace_back(run, c->workers[(size_t)k], k);
    run(c->workers[0], 0);
    for (auto& t : threads) t.join();
    for (Ctx* w : c->workers) c->launches += w->launches, w->launches = 0;
    if (fatal.load() != ICP_OK) {
        c->err = "batch: a worker failed: " + c->workers[(size_t)std::max(fatal_worker.load(), 0)]->err;
        return fatal.load();
    }
    return worst.load();
}

}  // extern "C"
