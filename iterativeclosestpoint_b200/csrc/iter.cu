// The per-iteration reductions and the Kabsch solve (replaces core/icpengine.cpp:187-346,
// ICPEngine::computeBestFitTransform :76-115 and Eigen::JacobiSVD<Matrix3d>; CLI twin
// icp_registration.cpp:389-440,499-603).
//
//   stage A   distances -> (count, mean, M2, min, max, problems) by Chan merges in a FIXED tree order
//             -> mean, population std, threshold (icpengine.cpp:235-255)
//   stage B   inlier mask d <= thr, inlier count, sum d^2, and first/second moments of the matched pairs
//             about pivots (icpengine.cpp:263-278, 325-337, 82-90)
//   solve     RMSE, loop control (icpengine.cpp:287-323), H -> 3x3 two-sided Jacobi SVD in Eigen's exact
//             operation order -> R = V U^T with reflection fix -> t -> T, T_cum = T * T_cum
//   apply     src = T * src (icpengine.cpp:345), normally fused into the next NN kernel's load
//
// All partial results are combined in an order that depends only on the launch geometry, never on
// scheduling, so a run is bit-reproducible and every rank of a sharded run derives identical totals.
#include "internal.h"
#include "solve.cuh"
#include <algorithm>

namespace icpb {

constexpr int RED_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// stage A
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ StatA stat_empty() {
    StatA s;
    s.n = 0.0; s.mean = 0.0; s.m2 = 0.0; s.dmin = DBL_MAX; s.dmax = 0.0; s.problems = 0.0;
    return s;
}

// Welford update == stat_merge(a, {1, d, 0}) (the stage API's kernel below; the loop's own pass does not divide per element)
__device__ __forceinline__ void stat_push(StatA& a, double d) {
    a.n += 1.0;
    const double delta = d - a.mean;
    a.mean += delta / a.n;
    a.m2 += delta * (d - a.mean);
    if (isfinite(d)) {
        a.dmin = fmin(a.dmin, d);
        a.dmax = fmax(a.dmax, d);
    } else {
        a.problems += 1.0;
    }
}

__device__ __forceinline__ StatA stat_block_merge(StatA acc, StatA* sm /* RED_THREADS */) {
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 1; s < RED_THREADS; s <<= 1) {
        if ((threadIdx.x % (2 * s)) == 0) sm[threadIdx.x] = stat_merge(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    return sm[0];
}

// Distances -> this rank's (count, mean, M2, min, max, problems), streaming at memory speed: every thread keeps the pivoted
// moments  s1 = sum (d - p),  s2 = sum (d - p)^2  of a fixed set of elements (p = last iteration's mean, 0 at the first: no
// division per element and no cancellation in  M2 = s2 - s1^2 / n  once p is near the mean), the block adds them in a fixed
// shuffle tree, the last block adds the block partials in index order.  mean = p + s1 / n.  Independent of scheduling.
struct MomA {
    double n, s1, s2, dmin, dmax, problems;
};

__device__ __forceinline__ void moma_push(MomA& a, double d, double p) {
    const double t = d - p;
    a.s1 += t;
    a.s2 += t * t;
    if (isfinite(d)) {
        a.dmin = d < a.dmin ? d : a.dmin;
        a.dmax = d > a.dmax ? d : a.dmax;
    } else {
        a.problems += 1.0;
    }
}

__device__ __forceinline__ MomA moma_add(const MomA& a, const MomA& b) {
    MomA r;
    r.n = a.n + b.n;
    r.s1 = a.s1 + b.s1;
    r.s2 = a.s2 + b.s2;
    r.dmin = fmin(a.dmin, b.dmin);
    r.dmax = fmax(a.dmax, b.dmax);
    r.problems = a.problems + b.problems;
    return r;
}

__device__ __forceinline__ MomA moma_block_sum(MomA v, MomA* sm /* RED_THREADS / 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MomA w;
        w.n = __shfl_xor_sync(0xffffffffu, v.n, o);
        w.s1 = __shfl_xor_sync(0xffffffffu, v.s1, o);
        w.s2 = __shfl_xor_sync(0xffffffffu, v.s2, o);
        w.dmin = __shfl_xor_sync(0xffffffffu, v.dmin, o);
        w.dmax = __shfl_xor_sync(0xffffffffu, v.dmax, o);
        w.problems = __shfl_xor_sync(0xffffffffu, v.problems, o);
        v = moma_add(v, w);  // commutative term by term: every lane ends with the same bits
    }
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    MomA r = sm[0];
#pragma unroll
    for (int w = 1; w < RED_THREADS / 32; ++w) r = moma_add(r, sm[w]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(RED_THREADS) stat_a_kernel(const double* __restrict__ dist, int64_t n,
                                                             StatA* __restrict__ part, LoopState* __restrict__ st,
                                                             StatA* __restrict__ rank_slot, const PeerMail pm) {
    __shared__ MomA sm[RED_THREADS / 32];
    __shared__ bool is_last;
    if (st->exit_code != 0) return;  // the loop has ended: iterations enqueued ahead of the host's knowledge do nothing
    const double p = isfinite(st->mean) ? st->mean : 0.0;
    // block-contiguous chunk, thread-strided inside it (coalesced, fixed assignment), four loads in flight
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t b = (int64_t)blockIdx.x * chunk, e = min(n, b + chunk);
    MomA acc;
    acc.n = 0.0; acc.s1 = 0.0; acc.s2 = 0.0; acc.dmin = DBL_MAX; acc.dmax = 0.0; acc.problems = 0.0;
    int64_t i = b + threadIdx.x;
    for (; i + 7 * RED_THREADS < e; i += 8 * RED_THREADS) {  // eight loads in flight per thread
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(dist + i + u * RED_THREADS);
#pragma unroll
        for (int u = 0; u < 8; ++u) moma_push(acc, v[u], p);
        acc.n += 8.0;
    }
    for (; i < e; i += RED_THREADS) {
        moma_push(acc, dist[i], p);
        acc.n += 1.0;
    }
    const MomA tot = moma_block_sum(acc, sm);
    MomA* mpart = reinterpret_cast<MomA*>(part);
    if (threadIdx.x == 0) {
        mpart[blockIdx.x] = tot;
        __threadfence();
        is_last = atomicAdd(&st->ticket_a, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int n_part = (int)gridDim.x;
    const int per = (n_part + RED_THREADS - 1) / RED_THREADS;
    const int pb = threadIdx.x * per, pe = min(n_part, pb + per);
    MomA a2;
    a2.n = 0.0; a2.s1 = 0.0; a2.s2 = 0.0; a2.dmin = DBL_MAX; a2.dmax = 0.0; a2.problems = 0.0;
    for (int k = pb; k < pe; ++k) {
        const double* q = reinterpret_cast<const double*>(mpart + k);
        MomA v;
        v.n = __ldcg(q); v.s1 = __ldcg(q + 1); v.s2 = __ldcg(q + 2); v.dmin = __ldcg(q + 3); v.dmax = __ldcg(q + 4); v.problems = __ldcg(q + 5);
        a2 = moma_add(a2, v);
    }
    const MomA all_m = moma_block_sum(a2, sm);
    StatA all;
    all.n = all_m.n;
    all.mean = all_m.n > 0.0 ? p + all_m.s1 / all_m.n : 0.0;
    all.m2 = all_m.n > 0.0 ? fmax(all_m.s2 - all_m.s1 * (all_m.s1 / all_m.n), 0.0) : 0.0;
    if (!(all.m2 == all.m2)) all.m2 = all_m.s2;  // inf - inf: keep the infinity the reference would carry
    all.dmin = all_m.dmin;
    all.dmax = all_m.dmax;
    all.problems = all_m.problems;
    if (threadIdx.x == 0) {
        *rank_slot = all;
        st->ticket_a = 0u;
    }
    if (pm.epoch && threadIdx.x < pm.n_ranks) {
        // this rank's record into every rank's mailbox (its own included), then the epoch: one thread per destination
        Mailbox* m = pm.peer[threadIdx.x];
        m->a[pm.rank] = all;
        __threadfence_system();
        st_release_sys(&m->flag_a[pm.rank], pm.epoch);
    }
}

int stat_a_blocks(Ctx* c, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + RED_THREADS - 1) / RED_THREADS, (int64_t)c->sm_count * 8));
}

int stat_a_launch(Ctx* c, const double* dist, int64_t n, StatA* part, StatA* rank_slot, const PeerMail* pm) {
    PeerMail none;
    none.epoch = 0u;
    none.n_ranks = 1;
    none.rank = 0;
    stat_a_kernel<<<stat_a_blocks(c, n), RED_THREADS, 0, c->stream>>>(dist, n, part, c->d_state, rank_slot, pm ? *pm : none);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

// one block: merges `n_part` partials into out[0] (fixed order); used by the stage API
__global__ void __launch_bounds__(RED_THREADS) stage_a_reduce_kernel(const StatA* __restrict__ part, int n_part,
                                                                     StatA* __restrict__ out) {
    __shared__ StatA sm[RED_THREADS];
    const int per = (n_part + RED_THREADS - 1) / RED_THREADS;
    const int b = threadIdx.x * per, e = min(n_part, b + per);
    StatA acc = stat_empty();
    for (int k = b; k < e; ++k) acc = stat_merge(acc, part[k]);
    const StatA tot = stat_block_merge(acc, sm);
    if (threadIdx.x == 0) out[0] = tot;
}

int stage_a_reduce_launch(Ctx* c, const StatA* part, int n_part, StatA* rank_part_slot) {
    stage_a_reduce_kernel<<<1, RED_THREADS, 0, c->stream>>>(part, n_part, rank_part_slot);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

// distances for caller-supplied correspondences (stage API): d_i = |src_i - tgt[idx_i]|, out-of-range idx ->
// DBL_MAX and a problem (icpengine.cpp:199-204, engine variant only)
__global__ void __launch_bounds__(RED_THREADS) dist_from_idx_kernel(const double* __restrict__ sx, const double* __restrict__ sy,
                                                                    const double* __restrict__ sz, const int32_t* __restrict__ idx,
                                                                    int64_t n, const uint32_t* __restrict__ inv_perm, int64_t m,
                                                                    const TPoint* __restrict__ pts, int variant,
                                                                    uint32_t* __restrict__ pos_out, double* __restrict__ dist_out,
                                                                    StatA* __restrict__ part) {
    __shared__ StatA sm[RED_THREADS];
    StatA acc = stat_empty();
    const int64_t per = (n + (int64_t)gridDim.x * RED_THREADS - 1) / ((int64_t)gridDim.x * RED_THREADS);
    const int64_t b = ((int64_t)blockIdx.x * RED_THREADS + threadIdx.x) * per;
    const int64_t e = min(n, b + per);
    for (int64_t i = b; i < e; ++i) {
        const int32_t j = idx[i];
        StatA one = stat_empty();
        one.n = 1.0;
        double d;
        uint32_t pos = 0xFFFFFFFFu;
        if (j < 0 || (int64_t)j >= m) {
            d = DBL_MAX;
            if (variant == ICP_VARIANT_ENGINE) one.problems = 1.0;
        } else {
            pos = inv_perm[j];
            const TPoint p = pts[pos];
            d = dsqrt(sumsq3(dsub(sx[i], p.x), dsub(sy[i], p.y), dsub(sz[i], p.z)));
            if (!isfinite(d)) one.problems = 1.0;
        }
        one.mean = d;
        if (isfinite(d) && !(j < 0 || (int64_t)j >= m)) {
            one.dmin = d;
            one.dmax = d;
        }
        pos_out[i] = pos;
        dist_out[i] = d;
        acc = stat_merge(acc, one);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 1; s < RED_THREADS; s <<= 1) {
        if ((threadIdx.x % (2 * s)) == 0) sm[threadIdx.x] = stat_merge(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}

int dist_from_idx_launch(Ctx* c, const double* sx, const double* sy, const double* sz, const int32_t* idx, int64_t n,
                         uint32_t* pos_out, double* dist_out, StatA* part, int* n_part) {
    const int blocks = (int)std::min<int64_t>((n + RED_THREADS - 1) / RED_THREADS, (int64_t)c->sm_count * 4);
    dist_from_idx_kernel<<<blocks, RED_THREADS, 0, c->stream>>>(sx, sy, sz, idx, n, c->fast.inv_perm, c->fast.n_pts,
                                                                c->fast.pts, c->params.variant, pos_out, dist_out, part);
    c->launches++;
    *n_part = blocks;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// stage B
// ------------------------------------------------------------------------------------------------
struct AccB {
    double v[STATB_DOUBLES];
};

__device__ __forceinline__ void accb_zero(AccB& a) {
#pragma unroll
    for (int k = 0; k < STATB_DOUBLES; ++k) a.v[k] = 0.0;
}

__device__ __forceinline__ void accb_add_pair(AccB& acc, double d, double ax, double ay, double az, double bx, double by,
                                              double bz, const double* pa, const double* pb) {
    acc.v[0] += 1.0;
    acc.v[1] += d * d;  // distances[idx] * distances[idx] (icpengine.cpp:271-273)
    const double a[3] = {ax - pa[0], ay - pa[1], az - pa[2]};
    const double b[3] = {bx - pb[0], by - pb[1], bz - pb[2]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        acc.v[2 + r] += a[r];
        acc.v[5 + r] += b[r];
#pragma unroll
        for (int q = 0; q < 3; ++q) acc.v[8 + 3 * r + q] += a[r] * b[q];
    }
}

__device__ __forceinline__ void accb_block_reduce(AccB& acc, double* out /* 17 doubles */) {
    __shared__ double sm[RED_THREADS / 32][STATB_DOUBLES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < STATB_DOUBLES; ++k) {
        double v = acc.v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < STATB_DOUBLES) {
        double v = sm[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < RED_THREADS / 32; ++w) v += sm[w][threadIdx.x];
        out[threadIdx.x] = v;
        __threadfence();
    }
}

// forward declarations (defined below)
__device__ __forceinline__ void sum_partials_fixed(const double* part, int n_part, double* out, double* sm);

// Inlier test + accumulation.  Every block derives the threshold from the (already gathered) stage-A rank
// partials -- identical arithmetic in every block and on every rank -- then owns block-strided tiles (fixed
// geometry => fixed summation order).  The last block to finish sums the block partials in index order into this
// rank's slot and, on a single-rank run, goes straight on to the solve.
#ifndef STAGEB_UNR
#define STAGEB_UNR 4
#define STAGEB_BLOCKS 2
#endif
__global__ void __launch_bounds__(RED_THREADS, STAGEB_BLOCKS) stage_b_kernel(const double* __restrict__ sx, const double* __restrict__ sy,
                                                              const double* __restrict__ sz, const uint32_t* __restrict__ pos,
                                                              const double* __restrict__ dist, int64_t n,
                                                              const TPoint* __restrict__ pts, LoopState* __restrict__ st,
                                                              const StatA* __restrict__ rank_a, int n_ranks, int rank, int iter,
                                                              uint8_t* __restrict__ mask_out, double* __restrict__ part,
                                                              double* __restrict__ rank_b, IterRecord* __restrict__ rec,
                                                              const PeerMail pm, int stop_req) {
    __shared__ double s_thr;
    __shared__ double s_rb[MAIL_RANKS * STATB_DOUBLES];
    __shared__ double sm_red[RED_THREADS];
    __shared__ bool is_last;
    if (st->exit_code != 0) return;  // the loop has ended: iterations enqueued ahead of the host's knowledge do nothing
    StatA a_all;
    double mean = 0.0, sd = 0.0;
    if (threadIdx.x == 0) {
        double thr;
        if (pm.epoch) {
            // the stage-A records arrive in this rank's own mailbox; read them past the caches once every epoch is in
            const Mailbox* own = pm.peer[pm.rank];
            const bool arrived = mail_wait(own->flag_a, n_ranks, pm.epoch);
            stat_a_finalize(st, own->a, n_ranks, iter, a_all, mean, sd, thr);  // (loads past L1: ld.global.cg)
            if (!arrived) thr = __longlong_as_double(0x7FF8000000000000LL);  // a peer is gone: no inliers, the run ends
        } else {
            stat_a_finalize(st, rank_a, n_ranks, iter, a_all, mean, sd, thr);
        }
        s_thr = thr;
    }
    __syncthreads();
    const double thr = s_thr;
    double pa[3], pb[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        pa[a] = st->pivot_a[a];
        pb[a] = st->pivot_b[a];
    }
    AccB acc;
    accb_zero(acc);
    // UNR tiles per trip: the (distance, match) loads, then the gathers of the matched points and the query loads, are in
    // flight together; the order in which a thread adds its queries is still fixed by the launch geometry alone.  Everything
    // stays in registers: the loads are unconditional (an outlier's slot reads point 0 and is not added), so no predicated
    // array element forces the compiler into local memory.
    constexpr int UNR = STAGEB_UNR;
    const int64_t stride = (int64_t)gridDim.x * RED_THREADS;
    for (int64_t base = (int64_t)blockIdx.x * RED_THREADS; base < n; base += stride * UNR) {
        double d[UNR];
        uint32_t p[UNR];
        bool ok[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = base + u * stride + threadIdx.x;
            d[u] = (i < n) ? __ldg(dist + i) : 0.0;
            p[u] = (i < n) ? __ldg(pos + i) : 0xFFFFFFFFu;
        }
        double ax[UNR], ay[UNR], az[UNR], bx[UNR], by[UNR], bz[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = base + u * stride + threadIdx.x;
            ok[u] = (i < n) && (d[u] <= thr) && (p[u] != 0xFFFFFFFFu);  // icpengine.cpp:264-268 (NaN => outlier)
            if (mask_out && i < n) mask_out[i] = ok[u] ? 1 : 0;
            const int64_t iq = ok[u] ? i : 0;
            const uint32_t ip = ok[u] ? p[u] : 0u;
            long long w;
            asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=d"(bx[u]), "=d"(by[u]), "=d"(bz[u]), "=l"(w) : "l"(pts + ip));
            ax[u] = __ldg(sx + iq);
            ay[u] = __ldg(sy + iq);
            az[u] = __ldg(sz + iq);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            if (ok[u]) accb_add_pair(acc, d[u], ax[u], ay[u], az[u], bx[u], by[u], bz[u], pa, pb);
    }
    accb_block_reduce(acc, part + (int64_t)blockIdx.x * STATB_DOUBLES);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(&st->ticket_b, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    sum_partials_fixed(part, (int)gridDim.x, rank_b + (int64_t)rank * STATB_DOUBLES, sm_red);
    if (threadIdx.x == 0) {
        rank_b[(int64_t)rank * STATB_DOUBLES + STATB_STOP] = stop_req ? 1.0 : 0.0;
        st->ticket_b = 0u;
        st->a = a_all;
        st->mean = mean;
        st->std_dev = sd;
        st->threshold = thr;
        st->iter = iter;
        if (n_ranks == 1) {
            __threadfence();
            solve_step(st, rank_b, 1, rec);
        }
    }
    if (pm.epoch) {
        // this rank's 17 sums into every mailbox, then the epoch; then wait for everyone's and solve right here --
        // every rank sums the records in rank order with the same arithmetic, so all ranks hold the same transform
        __syncthreads();
        __threadfence();
        const double* mine = rank_b + (int64_t)rank * STATB_DOUBLES;
        for (int t = threadIdx.x; t < n_ranks * STATB_DOUBLES; t += RED_THREADS)
            pm.peer[t / STATB_DOUBLES]->b[rank][t % STATB_DOUBLES] = __ldcg(mine + t % STATB_DOUBLES);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < n_ranks) st_release_sys(&pm.peer[threadIdx.x]->flag_b[rank], pm.epoch);
        const Mailbox* own = pm.peer[rank];
        __shared__ bool s_arrived;
        if (threadIdx.x == 0) s_arrived = mail_wait(own->flag_b, n_ranks, pm.epoch);
        __syncthreads();
        for (int t = threadIdx.x; t < n_ranks * STATB_DOUBLES; t += RED_THREADS)
            s_rb[t] = s_arrived ? __ldcg(&own->b[t / STATB_DOUBLES][t % STATB_DOUBLES]) : 0.0;  // nothing => fewer than 3 inliers
        __syncthreads();
        if (threadIdx.x == 0) solve_step(st, s_rb, n_ranks, rec);
    }
}

// explicit pairs (best-fit stage API): all pairs are inliers, pivots = first pair
__global__ void __launch_bounds__(RED_THREADS) pairs_b_kernel(const double* __restrict__ a_xyz, const double* __restrict__ b_xyz,
                                                              int64_t n, double* __restrict__ part) {
    double pa[3], pb[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        pa[a] = a_xyz[a];
        pb[a] = b_xyz[a];
    }
    AccB acc;
    accb_zero(acc);
    for (int64_t base = (int64_t)blockIdx.x * RED_THREADS; base < n; base += (int64_t)gridDim.x * RED_THREADS) {
        const int64_t i = base + threadIdx.x;
        if (i >= n) break;
        accb_add_pair(acc, 0.0, a_xyz[3 * i], a_xyz[3 * i + 1], a_xyz[3 * i + 2], b_xyz[3 * i], b_xyz[3 * i + 1],
                      b_xyz[3 * i + 2], pa, pb);
    }
    accb_block_reduce(acc, part + (int64_t)blockIdx.x * STATB_DOUBLES);
}

// one block: sums n_part partial records, all 17 components at once.  Thread t adds component t % 17 over slice t / 17 of
// the records (15 contiguous slices, records in index order), then 17 threads add the 15 slice sums in slice order: a fixed
// order, independent of scheduling.  (One component after the other with a block-wide tree each cost ~40 us of pure
// latency per iteration -- half of stage B on an eighth of the cloud.)
__device__ __forceinline__ void sum_partials_fixed(const double* part, int n_part, double* out, double* sm) {
    constexpr int SLICES = RED_THREADS / STATB_DOUBLES;
    const int k = threadIdx.x % STATB_DOUBLES, s = threadIdx.x / STATB_DOUBLES;
    double v = 0.0;
    if (s < SLICES) {
        const int per = (n_part + SLICES - 1) / SLICES;
        const int b = s * per, e = min(n_part, b + per);
#pragma unroll 8
        for (int j = b; j < e; ++j) v += __ldcg(part + (int64_t)j * STATB_DOUBLES + k);
    }
    sm[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x < STATB_DOUBLES) {
        double a = 0.0;
        for (int q = 0; q < SLICES; ++q) a += sm[q * STATB_DOUBLES + threadIdx.x];
        out[threadIdx.x] = a;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RED_THREADS) stage_b_reduce_kernel(const double* __restrict__ part, int n_part,
                                                                     double* __restrict__ out) {
    __shared__ double sm[RED_THREADS];
    sum_partials_fixed(part, n_part, out, sm);
}

int stage_b_blocks(Ctx* c, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + RED_THREADS - 1) / RED_THREADS, (int64_t)c->sm_count * 8));
}

int stage_b_launch(Ctx* c, const double* sx, const double* sy, const double* sz, const uint32_t* pos, const double* dist,
                   int64_t n, int iter, const StatA* rank_a, uint8_t* mask_out, double* part, double* rank_b, const PeerMail* pm,
                   IterRecord* rec, int stop_req) {
    PeerMail none;
    none.epoch = 0u;
    none.n_ranks = 1;
    none.rank = 0;
    const int blocks = stage_b_blocks(c, n);
    stage_b_kernel<<<blocks, RED_THREADS, 0, c->stream>>>(sx, sy, sz, pos, dist, n, c->fast.pts, c->d_state, rank_a, c->n_ranks,
                                                          c->rank, iter, mask_out, part, rank_b, rec ? rec : c->d_rec, pm ? *pm : none, stop_req);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int pairs_b_launch(Ctx* c, const double* a_xyz, const double* b_xyz, int64_t n, double* part, double* out17) {
    const int blocks = stage_b_blocks(c, n);
    pairs_b_kernel<<<blocks, RED_THREADS, 0, c->stream>>>(a_xyz, b_xyz, n, part);
    stage_b_reduce_kernel<<<1, RED_THREADS, 0, c->stream>>>(part, blocks, out17);
    c->launches += 2;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

__global__ void solve_kernel(LoopState* __restrict__ st, const double* __restrict__ rank_parts, int n_ranks,
                             IterRecord* __restrict__ rec) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st->exit_code != 0) return;  // the loop has ended
    solve_step(st, rank_parts, n_ranks, rec);
}

int solve_launch(Ctx* c, const double* rank_parts, int n_ranks, IterRecord* rec) {
    solve_kernel<<<1, 32, 0, c->stream>>>(c->d_state, rank_parts, n_ranks, rec ? rec : c->d_rec);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

// stage API: best-fit transform from summed moments / from a caller-supplied H
__global__ void bestfit_kernel(const double* __restrict__ b17, const double* __restrict__ a0, const double* __restrict__ b0,
                               double* __restrict__ T_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double pa[3] = {a0[0], a0[1], a0[2]}, pb[3] = {b0[0], b0[1], b0[2]};
    double cA[3], cB[3], H[9], T[16];
    moments_to_H(b17, pa, pb, cA, cB, H);
    solve_from_H(H, cA, cB, T, nullptr, nullptr, nullptr);
    for (int i = 0; i < 16; ++i) T_out[i] = T[i];
}

__global__ void solve_from_H_kernel(const double* __restrict__ in /* H9, cA3, cB3 */, double* __restrict__ out /* T16,U9,S3,V9 */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double T[16], U[9], S[3], V[9];
    solve_from_H(in, in + 9, in + 12, T, U, S, V);
    for (int i = 0; i < 16; ++i) out[i] = T[i];
    for (int i = 0; i < 9; ++i) {
        out[16 + i] = U[i];
        out[28 + i] = V[i];
    }
    for (int i = 0; i < 3; ++i) out[25 + i] = S[i];
}

int bestfit_launch(Ctx* c, const double* b17, const double* a0, const double* b0, double* T_out) {
    bestfit_kernel<<<1, 32, 0, c->stream>>>(b17, a0, b0, T_out);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int solve_from_H_launch(Ctx* c, const double* in15, double* out37) {
    solve_from_H_kernel<<<1, 32, 0, c->stream>>>(in15, out37);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// apply / layout kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void apply_T(const double* T, double& x, double& y, double& z) {
    const double a = x, b = y, c = z;
    x = dadd(dadd(dadd(dmul(T[0], a), dmul(T[1], b)), dmul(T[2], c)), T[3]);
    y = dadd(dadd(dadd(dmul(T[4], a), dmul(T[5], b)), dmul(T[6], c)), T[7]);
    z = dadd(dadd(dadd(dmul(T[8], a), dmul(T[9], b)), dmul(T[10], c)), T[11]);
}

// applies state->T_pending if state->have_T (used once at loop exit)
__global__ void __launch_bounds__(256) apply_pending_kernel(const LoopState* __restrict__ st, double* __restrict__ x,
                                                            double* __restrict__ y, double* __restrict__ z, int64_t n,
                                                            float* __restrict__ lb) {
    if (!st->have_T) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = x[i], b = y[i], c = z[i];
    const double oa = a, ob = b, oc = c;
    apply_T(st->T_pending, a, b, c);
    x[i] = a;
    y[i] = b;
    z[i] = c;
    if (lb) {  // nn_keep.cu: the bound holds for where the point was; it moved by at most this much
        const double moved = dmul(dsqrt(sumsq3(dsub(a, oa), dsub(b, ob), dsub(c, oc))), 1.0 + 1e-9);
        lb[i] = __double2float_rd(dsub((double)lb[i], moved));
    }
}

__global__ void clear_pending_kernel(LoopState* st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) st->have_T = 0;
}

__global__ void __launch_bounds__(256) apply_aos_kernel(const double* __restrict__ T, double* __restrict__ xyz, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = xyz[3 * i], b = xyz[3 * i + 1], c = xyz[3 * i + 2];
    apply_T(T, a, b, c);
    xyz[3 * i] = a;
    xyz[3 * i + 1] = b;
    xyz[3 * i + 2] = c;
}

__global__ void __launch_bounds__(256) unsort_kernel(const double* __restrict__ sx, const double* __restrict__ sy,
                                                     const double* __restrict__ sz, const uint32_t* __restrict__ perm,
                                                     int64_t n, double* __restrict__ out_xyz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t j = perm ? (int64_t)perm[i] : i;
    out_xyz[3 * j] = sx[i];
    out_xyz[3 * j + 1] = sy[i];
    out_xyz[3 * j + 2] = sz[i];
}

// scatter per-query results back to the caller's order: idx (original target index) and distance
__global__ void __launch_bounds__(256) unsort_results_kernel(const uint32_t* __restrict__ pos, const double* __restrict__ dist,
                                                             const uint32_t* __restrict__ perm, const TPoint* __restrict__ pts,
                                                             int64_t n, int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t j = perm ? (int64_t)perm[i] : i;
    idx_out[j] = (int32_t)pts[pos[i]].idx;
    if (dist_out) dist_out[j] = dist[i];
}

__global__ void __launch_bounds__(256) aos_to_soa_kernel(const double* __restrict__ xyz, int64_t n, double* __restrict__ sx,
                                                         double* __restrict__ sy, double* __restrict__ sz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sx[i] = xyz[3 * i];
    sy[i] = xyz[3 * i + 1];
    sz[i] = xyz[3 * i + 2];
}

static inline int nblk(int64_t n) { return (int)((n + 255) / 256); }

int apply_pending_launch(Ctx* c, double* x, double* y, double* z, int64_t n, float* lb) {
    if (n > 0) apply_pending_kernel<<<nblk(n), 256, 0, c->stream>>>(c->d_state, x, y, z, n, lb);
    clear_pending_kernel<<<1, 32, 0, c->stream>>>(c->d_state);
    c->launches += 2;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int apply_aos_launch(Ctx* c, const double* d_T16, double* xyz, int64_t n) {
    if (n <= 0) return ICP_OK;
    apply_aos_kernel<<<nblk(n), 256, 0, c->stream>>>(d_T16, xyz, n);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int unsort_launch(Ctx* c, const double* sx, const double* sy, const double* sz, const uint32_t* perm, int64_t n,
                  double* out_xyz) {
    if (n <= 0) return ICP_OK;
    unsort_kernel<<<nblk(n), 256, 0, c->stream>>>(sx, sy, sz, perm, n, out_xyz);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int unsort_results_launch(Ctx* c, const uint32_t* pos, const double* dist, const uint32_t* perm, int64_t n, int32_t* idx_out,
                          double* dist_out) {
    if (n <= 0) return ICP_OK;
    unsort_results_kernel<<<nblk(n), 256, 0, c->stream>>>(pos, dist, perm, c->fast.pts, n, idx_out, dist_out);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int aos_to_soa_launch(Ctx* c, const double* xyz, int64_t n, double* sx, double* sy, double* sz) {
    if (n <= 0) return ICP_OK;
    aos_to_soa_kernel<<<nblk(n), 256, 0, c->stream>>>(xyz, n, sx, sy, sz);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
