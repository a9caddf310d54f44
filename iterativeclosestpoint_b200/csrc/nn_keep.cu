// Exact nearest-neighbour query, temporal form (mode 5): keep what can be proven, search only the rest.
// Same job and same answers as nn.cu (replaces Octree::findNearest / searchNearest, core/octree.cpp:128-184, inside the
// per-point loop of core/icpengine.cpp:172-184); the exactness argument is the one at the top of nn.cu: the reference
// returns P whenever s(P)(1 + 2^-40) < s(p) for every other target point p, whatever order it visits things in.
//
// ICP asks the same N questions again and again while the source creeps towards the target, so most answers are already
// known.  Every settled query carries, from one iteration to the next,
//     cand[K]   the K nearest target points its last search found (positions in the search tree's point order), and
//     lb        a lower bound (float, rounded down) on its distance to EVERY target point that is not in cand.
// Next iteration the query has moved by `moved`, so every point outside cand is at least lb - moved away (triangle
// inequality).  nn_keep_kernel evaluates the K candidates (K point loads, no traversal): if the nearest of them is
// closer than lb - moved by a margin far above 2^-40, and the runner-up among the candidates is 2^-40 away as well, it is
// the unique nearest neighbour of the whole cloud -- the match (possibly a different one of the K than last time) is
// written out, lb shrinks by `moved`, done.  On the synthetic aerial scenes this settles ~75 % of the queries while the
// registration is still moving by centimetres per iteration and ~100 % once it has converged; that part of the NN stage
// is a pure streaming pass (HBM bound).
// The queries it cannot settle are compacted onto a work list.  nn_collect_kernel runs over that list with full warps
// (see there): ball around the query with the radius of its best candidate, widened so that the new bound is worth
// something; every point of every touched grid cell, the scan balanced over the warp; every point inside the ball is
// collected, the K nearest become the new candidates.  Crowded cells (entered through the tree), oversized blocks,
// missing seeds, overfull lists and non-unique minima go on a second list for the per-thread kernel of nn.cu (cell walk
// with tree descent, climbing search, literal traversal); those record no bound (lb = 0).
#include <algorithm>
#include "nn_common.cuh"

namespace icpb {

constexpr int KP_THREADS = 256;
constexpr double MARGIN40 = 1.0 + 9.094947017729282e-13;  // 1 + 2^-40

// warp-aggregated append of `v` (for the lanes with `p`) to a device list; every lane of the warp must call it
__device__ __forceinline__ void list_push(uint32_t* list, unsigned int* count, const bool p, const uint32_t v) {
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    unsigned int at = 0;
    if (lane == (__ffs(m) - 1)) at = atomicAdd(count, (unsigned int)__popc(m));
    at = __shfl_sync(0xffffffffu, at, __ffs(m) - 1);
    if (p) list[at + __popc(m & ((1u << lane) - 1u))] = v;
}

template <int K>
__global__ void __launch_bounds__(KP_THREADS, 8) nn_keep_kernel(const NNArgs A) {
    const long long i = (long long)blockIdx.x * KP_THREADS + threadIdx.x;
    const bool active = i < A.n;
    bool kept = false, finite_q = false;
    if (active) {
        double qx = A.sx[i], qy = A.sy[i], qz = A.sz[i];
        const uint4 cv = A.cand_io[i];
        const float lbf = A.lb_io[i];
        double moved = 0.0;  // upper bound on how far this query moved since lb was recorded
        if (A.apply_pending && A.state->have_T) {
            const double ox = qx, oy = qy, oz = qz;
            apply_T_point(A.state->T_pending, qx, qy, qz);
            A.ox[i] = qx;
            A.oy[i] = qy;
            A.oz[i] = qz;
            moved = dmul(dsqrt(sumsq3(dsub(qx, ox), dsub(qy, oy), dsub(qz, oz))), 1.0 + 1e-9);
        }
        finite_q = isfinite(qx) && isfinite(qy) && isfinite(qz);
        uint32_t c[4] = {cv.x, cv.y, cv.z, cv.w};
        if (c[0] == NONE && A.prev_pos) c[0] = A.prev_pos[i];  // settled by another kernel: its match is the one candidate
        double best = ICPB_INF, second = ICPB_INF;
        uint32_t bpos = NONE;
        if (finite_q) {
            double px[K], py[K], pz[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                px[j] = py[j] = pz[j] = 0.0;
                uint32_t pidx;
                if (c[j] != NONE) load_point(A.pts, c[j], px[j], py[j], pz[j], pidx);
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double s = sumsq3(dsub(px[j], qx), dsub(py[j], qy), dsub(pz[j], qz));
                if (c[j] != NONE) {
                    if (s < best) {
                        second = best;
                        best = s;
                        bpos = c[j];
                    } else if (s < second) {
                        second = s;
                    }
                }
            }
        }
        if (bpos != NONE) {
            const double d = dsqrt(best);
            const double lbn = dsub((double)lbf, moved);  // every target point outside cand is at least this far away now
            if (second > dmul(best, MARGIN40) && d < dmul(lbn, 1.0 - 1e-6)) {
                kept = true;
                A.pos_out[i] = bpos;
                A.dist_out[i] = d;  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
                A.lb_io[i] = __double2float_rd(lbn);
            }
        }
        if (!kept) A.dist_out[i] = best;  // hand-over to the search: s of a real target point (+inf if none)
    }
    list_push(A.worklist, A.work_count, active && !kept && finite_q, (uint32_t)i);
    list_push(A.worklist2, A.work_count + 1, active && !kept && !finite_q, (uint32_t)i);
    if (A.counters) {  // profiling / tests only
        const unsigned kp = __ballot_sync(0xffffffffu, kept);
        if ((threadIdx.x & 31) == 0 && kp) {
            atomicAdd(&A.counters[0], (unsigned long long)__popc(kp));
            atomicAdd(&A.counters[5], (unsigned long long)__popc(kp));
        }
    }
}

// the K + 1 smallest s seen so far, ascending, with the positions of the first K
template <int K>
struct TopK {
    double t[K + 1];
    uint32_t p[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j <= K; ++j) t[j] = ICPB_INF;
#pragma unroll
        for (int j = 0; j < K; ++j) p[j] = NONE;
    }
    // requires s < t[K]
    __device__ __forceinline__ void insert(const double s, const uint32_t pos) {
        bool c[K + 1];
#pragma unroll
        for (int j = 0; j <= K; ++j) c[j] = s < t[j];
#pragma unroll
        for (int j = K; j >= 1; --j) {
            t[j] = c[j - 1] ? t[j - 1] : (c[j] ? s : t[j]);
            if (j < K) p[j] = c[j - 1] ? p[j - 1] : (c[j] ? pos : p[j]);
        }
        t[0] = c[0] ? s : t[0];
        p[0] = c[0] ? pos : p[0];
    }
};


// ---------------------------------------------------------------------------------------------------
// The search over the work list, balanced over the warp like nn_group.cu: every lane sets up one query (ball, pyramid
// level, the block of <= 3 x 3 x 3 cells the ball touches), the non-empty cells are cut into scan items of at most
// GW_SUB consecutive points, and the items are dealt out round-robin -- every lane scans one item per trip, whichever
// query it belongs to.  What comes back is not a per-item minimum but every point within the query's radius R:
// the scanning lane appends (s, position) to the owner's list in shared memory when s <= R^2.  R is at least the
// distance to the seed (a real point), so the nearest neighbour is always on the list; the owner sorts its list
// (K + 1 smallest), proves the minimum unique, keeps the K nearest as candidates and records
//     lb = min(R, clearance inside the block, (K+1)-th smallest on the list).
// ---------------------------------------------------------------------------------------------------
constexpr int GC_THREADS = 128;
constexpr int GC_WARPS = GC_THREADS / 32;
constexpr int GC_QCAP = 256;  // scan items per warp and round
constexpr int GC_SUB = 8;     // points per scan item (two batches of four loads in flight)
constexpr int GC_LIST = 16;   // collected points per query; more => the per-thread kernel takes the query

struct __align__(16) GcSlot {
    double qx, qy, qz, tau;
};

template <int K>
__global__ void __launch_bounds__(GC_THREADS) nn_collect_kernel(const NNArgs A) {
    __shared__ GcSlot slot_all[GC_WARPS][32];
    __shared__ uint2 queue_all[GC_WARPS][GC_QCAP];        // x = first point, y = count | owner lane << 8
    __shared__ double ls_all[GC_WARPS][GC_LIST][32];      // collected s, [entry][owner lane]
    __shared__ uint32_t lp_all[GC_WARPS][GC_LIST][32];    // collected positions
    __shared__ unsigned int lc_all[GC_WARPS][32];         // entries per owner (may exceed GC_LIST: overflow)
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    GcSlot* slot = slot_all[w];
    uint2* queue = queue_all[w];
    double (*ls)[32] = ls_all[w];
    uint32_t (*lp)[32] = lp_all[w];
    unsigned int* lc = lc_all[w];
    const unsigned int cnt = *A.work_count;
    unsigned long long n_ok = 0, n_cand = 0, n_items = 0;

    for (unsigned int base = blockIdx.x * GC_THREADS + (threadIdx.x & ~31); base < cnt; base += gridDim.x * GC_THREADS) {
        const unsigned int t = base + (unsigned int)lane;
        const bool act = t < cnt;
        // ---- A. own query ----
        uint32_t i = 0;
        double qx = 0.0, qy = 0.0, qz = 0.0, clear = 0.0, R = 0.0;
        bool elig = false;
        int x0 = 0, y0 = 0, z0 = 0, nx = 0, ny = 0, nz = 0;
        GridView V = grid_view(A, 0);
        if (act) {
            i = A.worklist[t];
            qx = A.sx[i];
            qy = A.sy[i];
            qz = A.sz[i];
            double Sd = A.dist_out[i];  // s of the best old candidate (nn_keep_kernel)
            if (!(Sd < 1e19)) Sd = walk_seed(A, qx, qy, qz);
            if (Sd < 1e19) {
                // the ball that must be searched has the radius of the seed; it is widened (x alpha, up to rcap) so that the
                // bound recorded for the next iterations reaches beyond the match
                const double r = dmul(dsqrt(Sd), 1.0 + 9.5367431640625e-07);  // sqrt(S) (1 + 2^-20)
                R = fmax(r, fmin(dmul(r, A.walk_alpha), A.walk_rcap));
                const double e = dadd(R, A.geps);
                const double wd = dmul(e, A.walk_wmul);
                int k = A.gnlev - 1;
                while (k > 0 && A.gedge[k] < wd) --k;
                V = grid_view(A, k);
                int x1 = grid_cell_index(A, V, dadd(qx, e), 0, V.nx), y1 = grid_cell_index(A, V, dadd(qy, e), 1, V.ny),
                    z1 = grid_cell_index(A, V, dadd(qz, e), 2, V.nz);
                x0 = grid_cell_index(A, V, dsub(qx, e), 0, V.nx);
                y0 = grid_cell_index(A, V, dsub(qy, e), 1, V.ny);
                z0 = grid_cell_index(A, V, dsub(qz, e), 2, V.nz);
                x0 = max(x0, 0); y0 = max(y0, 0); z0 = max(z0, 0);
                x1 = min(x1, V.nx - 1); y1 = min(y1, V.ny - 1); z1 = min(z1, V.nz - 1);
                nx = x1 - x0 + 1; ny = y1 - y0 + 1; nz = z1 - z0 + 1;
                if (nx >= 1 && ny >= 1 && nz >= 1 && nx <= 3 && ny <= 3 && nz <= 3) {
                    elig = true;
                    // clearance of q inside the block of cells: every point outside the block is farther than this
                    // (2 geps: the cells' bisection boundaries versus org + k * edge)
                    const double ed = A.gedge[k];
                    const double cx = fmin(dsub(qx, dadd(A.gorg[0], dmul((double)x0, ed))), dsub(dadd(A.gorg[0], dmul((double)(x1 + 1), ed)), qx));
                    const double cy = fmin(dsub(qy, dadd(A.gorg[1], dmul((double)y0, ed))), dsub(dadd(A.gorg[1], dmul((double)(y1 + 1), ed)), qy));
                    const double cz = fmin(dsub(qz, dadd(A.gorg[2], dmul((double)z0, ed))), dsub(dadd(A.gorg[2], dmul((double)(z1 + 1), ed)), qz));
                    clear = dsub(fmin(cx, fmin(cy, cz)), dmul(A.geps, 2.0));
                }
            }
        }
        {
            GcSlot s;
            s.qx = qx; s.qy = qy; s.qz = qz;
            s.tau = dmul(R, R);  // >= S (1 + 2^-21): the seed point itself is collected
            slot[lane] = s;
            lc[lane] = 0u;
        }
        const int ncell = elig ? nx * ny * nz : 0;                       // <= 27
        const int nchunk = __reduce_max_sync(FULL, (ncell + 7) >> 3);    // cells are taken eight at a time
        const int nxy = nx * ny;
        // c -> (dx, dy, dz) without integer division: (c * ceil(256 / d)) >> 8 == c / d for c < 27, d in {1, 2, 3, 4, 6, 9}
        const int mxy = elig ? (256 + nxy - 1) / nxy : 256, mx = elig ? (256 + nx - 1) / nx : 256;
        __syncwarp();

        for (int ch = 0; ch < nchunk; ++ch) {
            // ---- B1. this round's cells: all entries loaded at once, then decoded ----
            uint32_t ept[8], ecnt[8];
            uint2 en[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = ch * 8 + j;
                en[j] = make_uint2(0u, 0u);
                if (c < ncell) {
                    const int dz = (c * mxy) >> 8, rem = c - dz * nxy, dy = (rem * mx) >> 8, dx = rem - dy * nx;
                    en[j] = grid_entry(V, x0 + dx, y0 + dy, z0 + dz);
                }
            }
            bool crowded = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ept[j] = ecnt[j] = 0u;
                const uint32_t kind = en[j].y >> 30;
                if (kind == 3u) crowded = true;  // entered through the search tree: per-thread kernel
                bool take = kind == 1u;
                if (kind == 2u) {
                    // a leaf above the grid level owns an aligned block of cells: scan it once, from the first cell that the
                    // block and this query's range have in common
                    const int c = ch * 8 + j;
                    const int dz = (c * mxy) >> 8, rem = c - dz * nxy, dy = (rem * mx) >> 8, dx = rem - dy * nx;
                    const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
                    const int sh = V.level - (int)((en[j].y >> 24) & 0x3Fu);
                    take = x == max((x >> sh) << sh, x0) && y == max((y >> sh) << sh, y0) && z == max((z >> sh) << sh, z0);
                }
                if (take) {
                    ept[j] = en[j].x;
                    ecnt[j] = en[j].y & 0xFFFFFFu;
                }
            }
            if (crowded) elig = false;

            // ---- B2. queue the scan items (a lane's items are consecutive), then scan them round-robin ----
            uint32_t nsub = 0;
            if (elig) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    nsub += (ecnt[j] + GC_SUB - 1) / GC_SUB;
                    n_cand += ecnt[j];
                }
            }
            uint32_t off_end = nsub;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, off_end, o);
                if (lane >= o) off_end += v;
            }
            const bool fits = off_end <= (uint32_t)GC_QCAP;
            if (!fits) elig = false;
            const uint32_t total = __reduce_max_sync(FULL, fits ? off_end : 0u);
            n_items += (lane == 0) ? total : 0u;
            if (elig) {
                uint32_t o = off_end - nsub;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    for (uint32_t k = 0; k < ecnt[j]; k += GC_SUB) {
                        const uint32_t m = ecnt[j] - k;
                        queue[o++] = make_uint2(ept[j] + k, (m < (uint32_t)GC_SUB ? m : (uint32_t)GC_SUB) | ((uint32_t)lane << 8));
                    }
                }
            }
            __syncwarp();
            for (uint32_t b0 = 0; b0 < total; b0 += 32) {
                const uint32_t j = b0 + lane;
                const bool has = j < total;
                const uint2 it = has ? queue[j] : make_uint2(0u, 0u);
                const uint32_t pc = it.y & 0xFFu, owner = (it.y >> 8) & 31u;
                const GcSlot s = slot[owner];
                const bool second_batch = __any_sync(FULL, pc > 4u);
#pragma unroll
                for (int b = 0; b < GC_SUB; b += 4) {
                    if (b == 0 || second_batch) {
                        double px[4], py[4], pz[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            px[u] = py[u] = pz[u] = 0.0;
                            uint32_t pidx;
                            if ((uint32_t)(b + u) < pc) load_point(A.pts, it.x + (uint32_t)(b + u), px[u], py[u], pz[u], pidx);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double v = sumsq3(dsub(px[u], s.qx), dsub(py[u], s.qy), dsub(pz[u], s.qz));
                            if ((uint32_t)(b + u) < pc && v <= s.tau) {
                                const unsigned int at = atomicAdd(&lc[owner], 1u);
                                if (at < (unsigned int)GC_LIST) {
                                    ls[at][owner] = v;
                                    lp[at][owner] = it.x + (uint32_t)(b + u);
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();  // the queue is rewritten by the next round; the lists are complete after the last one
        }

        // ---- C. the owner sorts its list: unique minimum with margin 2^-40 => the reference's answer ----
        bool settled = false;
        if (act) {
            const unsigned int nl = lc[lane];
            if (elig && nl >= 1u && nl <= (unsigned int)GC_LIST) {
                TopK<K> top;
                top.init();
                for (unsigned int k = 0; k < nl; ++k) {
                    const double v = ls[k][lane];
                    if (v < top.t[K]) top.insert(v, lp[k][lane]);
                }
                if (top.t[1] > dmul(top.t[0], MARGIN40)) {
                    settled = true;
                    A.pos_out[i] = top.p[0];
                    A.dist_out[i] = dsqrt(top.t[0]);  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
                    uint4 cv;
                    cv.x = top.p[0];
                    cv.y = K > 1 ? top.p[K > 1 ? 1 : 0] : NONE;
                    cv.z = K > 2 ? top.p[K > 2 ? 2 : 0] : NONE;
                    cv.w = K > 3 ? top.p[K > 3 ? 3 : 0] : NONE;
                    A.cand_io[i] = cv;
                    // points not on the list: inside the block they have s > R^2, outside they are beyond the clearance;
                    // points on the list that are not candidates: at least the (K+1)-th smallest s away
                    const double others = fmin(fmin(dmul(R, 1.0 - 1e-9), clear), dsqrt(top.t[K]));
                    A.lb_io[i] = others > 0.0 ? __double2float_rd(dmul(others, 1.0 - 1e-9)) : 0.0f;
                    ++n_ok;
                }
            }
            if (!settled) {
                A.cand_io[i] = make_uint4(NONE, NONE, NONE, NONE);
                A.lb_io[i] = 0.0f;
            }
        }
        list_push(A.worklist2, A.work_count + 1, act && !settled, i);
        __syncwarp();  // slots, lists and counters are reused by the next trip
    }
    if (A.counters) {  // profiling / tests only
        const unsigned ok = __reduce_add_sync(FULL, (unsigned)n_ok);
        const unsigned nc = __reduce_add_sync(FULL, (unsigned)n_cand);
        if (lane == 0) {
            if (ok) atomicAdd(&A.counters[0], (unsigned long long)ok);
            atomicAdd(&A.counters[3], (unsigned long long)nc);
            atomicAdd(&A.counters[4], n_items);
        }
    }
}

template <int K>
static int keep_launch_k(Ctx* c, const NNArgs& A) {
    nn_keep_kernel<K><<<(int)((A.n + KP_THREADS - 1) / KP_THREADS), KP_THREADS, 0, c->stream>>>(A);
    const int blocks = (int)std::min<long long>((A.n + GC_THREADS - 1) / GC_THREADS, (long long)c->sm_count * 7);
    nn_collect_kernel<K><<<blocks, GC_THREADS, 0, c->stream>>>(A);
    c->launches += 2;
    return ICP_OK;
}

// keep + search; the caller runs the per-thread kernel over the second list afterwards
int nn_keep_launch(Ctx* c, const NNArgs& A, int k) {
    if (k <= 1) keep_launch_k<1>(c, A);
    else if (k == 2) keep_launch_k<2>(c, A);
    else if (k == 3) keep_launch_k<3>(c, A);
    else keep_launch_k<4>(c, A);
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
