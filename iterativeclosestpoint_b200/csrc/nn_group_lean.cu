// Exact nearest-neighbour query, cell walk with the candidate scan balanced over the warp (mode 4) -- the lean form, used
// while no temporal bound is recorded (mode 6 while the registration still moves, stateless queries): balls of at most
// 2 x 2 x 2 cells, one round, candidates above min(seed, best)(1 + 2^-39) skipped.  nn_group.cu is the general form
// (up to 3 x 3 x 3 cells in rounds of eight, every scanned point counted for the runner-up that feeds the bound).
// Same job and same answers as nn.cu (replaces Octree::findNearest / searchNearest, core/octree.cpp:128-184, inside the
// per-point loop of core/icpengine.cpp:172-184); the exactness argument is the one at the top of nn.cu:
// find the exact minimum of s over every point the search ball can contain, prove that it is unique by a margin of
// 2^-40, otherwise leave the query to the literal traversal.
//
// Why a second kernel: in the one-thread-per-query walk (nn_common.cuh: cell_walk) a lane scans 1..8 cells of 1..64
// points, so a warp runs as long as its unluckiest lane (ncu: 14 of 32 lanes active on average).  Here the warp
//   A. lets every lane set up its own query (move it by the pending transform, radius from last iteration's match,
//      grid level, the <= 2 x 2 x 2 cells its ball touches, all eight cell entries loaded at once),
//   B. cuts the non-empty cells into scan items of at most GW_SUB consecutive points, queued in shared memory,
//      and deals the items out round-robin: every lane scans one item per trip, whichever query it belongs to,
//   C. hands each item's (best, position, tie flag) back; the owning lane merges its own items in order.
// Queries it cannot settle (no seed, a ball over more than two cells along an axis, a crowded cell that is entered
// through the search tree, no unique minimum, queue overflow) go on a work list for the per-thread kernel (nn.cu).
#include "nn_common.cuh"

namespace icpb {

constexpr int GW_THREADS = 128;
constexpr int GW_WARPS = GW_THREADS / 32;
constexpr int GW_QCAP = 256;  // scan items per warp
constexpr int GW_SUB = 8;     // points per scan item (two batches of four loads in flight)

struct __align__(16) GlSlot {
    double qx, qy, qz, bound;
};

// SEED: queries without a previous match look for a seed in their own base cell (first iteration, stateless queries).  Later
// iterations run the instance without that code (a fifth of the kernel's instructions, never executed there: less pressure on
// the instruction cache); a query that has no match then simply goes on the work list.
template <bool SEED>
__global__ void __launch_bounds__(GW_THREADS) nn_group_lean_kernel(const NNArgs A) {
    __shared__ GlSlot slot_all[GW_WARPS][32];
    __shared__ uint2 queue_all[GW_WARPS][GW_QCAP];   // item: x = first point, y = count | owner lane << 8; result: x = position, y = tie
    __shared__ double rbest_all[GW_WARPS][GW_QCAP];  // result: smallest s of the item
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    GlSlot* slot = slot_all[w];
    uint2* queue = queue_all[w];
    double* rbest = rbest_all[w];
    const long long i = (long long)blockIdx.x * GW_THREADS + threadIdx.x;
    const bool active = i < A.n;

    // ---- A. own query ----
    double qx = 0.0, qy = 0.0, qz = 0.0, bound = 0.0;
    bool elig = false;
#ifdef ICPB_WHY
    int why = 0;  // profiling build only (make EXTRA=-DICPB_WHY): 1 no seed, 2 ball over more than two cells along an axis, 3 crowded cell, 4 queue overflow, 5 no unique minimum
#define ICPB_WHY_SET(v) why = (v)
#else
#define ICPB_WHY_SET(v)
#endif
    uint32_t ept[8], ecnt[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) ept[c] = ecnt[c] = 0u;
    if (active) {
        qx = A.sx[i];
        qy = A.sy[i];
        qz = A.sz[i];
        if (A.apply_pending && A.state->have_T) {
            apply_T_point(A.state->T_pending, qx, qy, qz);
            A.ox[i] = qx;
            A.oy[i] = qy;
            A.oz[i] = qz;
        }
        double Sd = ICPB_INF;
        const uint32_t pp = A.prev_pos ? A.prev_pos[i] : NONE;
        const bool finite_q = isfinite(qx) && isfinite(qy) && isfinite(qz);
        if (finite_q) {
            if (pp != NONE) {
                double px, py, pz;
                uint32_t pidx;
                load_point(A.pts, pp, px, py, pz, pidx);
                Sd = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
            }
            if (SEED && !(Sd < 1e19)) Sd = walk_seed(A, qx, qy, qz);
        }
        if (Sd < 1e19) {
            // the same ball, level and cell range as cell_walk (nn_common.cuh)
            const double r = sqrt_upper(Sd);  // >= sqrt(S): any box that holds the ball will do
            const double e = dadd(r, A.geps);
            const GridView V = grid_view(A, grid_level_for_width(A, dmul(e, 2.0), A.gbias));
            const double ec = dmul(e, V.inv);
            int x0, x1, y0, y1, z0, z1;
            grid_cell_span(A, V, qx, ec, 0, V.nx, x0, x1);
            grid_cell_span(A, V, qy, ec, 1, V.ny, y0, y1);
            grid_cell_span(A, V, qz, ec, 2, V.nz, z0, z1);
            x0 = max(x0, 0); y0 = max(y0, 0); z0 = max(z0, 0);
            x1 = min(x1, V.nx - 1); y1 = min(y1, V.ny - 1); z1 = min(z1, V.nz - 1);
            const int bx = x1 - x0, by = y1 - y0, bz = z1 - z0;  // cells per axis - 1
            if (bx >= 0 && by >= 0 && bz >= 0 && bx <= 1 && by <= 1 && bz <= 1) {
                elig = true;
                bound = dmul(Sd, 1.0 + 1.8189894035458565e-12);  // S (1 + 2^-39)
                const int ncell = 1 << (bx + by + bz);
                uint2 en[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    en[c] = make_uint2(0u, 0u);
                    if (c < ncell) {
                        const int c1 = c >> bx;
                        en[c] = grid_entry(V, x0 + (c & bx), y0 + (c1 & by), z0 + (c1 >> by));
                    }
                }
                bool crowded = false;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t kind = en[c].y >> 30;
                    if (kind == 3u) crowded = true;  // entered through the search tree: per-thread kernel
                    bool take = kind == 1u;
                    if (kind == 2u) {
                        // a leaf above the grid level owns an aligned block of cells: scan it once, from the first
                        // cell that the block and this query's range have in common
                        const int c1 = c >> bx;
                        const int x = x0 + (c & bx), y = y0 + (c1 & by), z = z0 + (c1 >> by);
                        const int sh = V.level - (int)((en[c].y >> 24) & 0x3Fu);
                        take = x == max((x >> sh) << sh, x0) && y == max((y >> sh) << sh, y0) && z == max((z >> sh) << sh, z0);
                    }
                    if (take) {
                        ept[c] = en[c].x;
                        ecnt[c] = en[c].y & 0xFFFFFFu;
                    }
                }
                if (crowded) {
                    elig = false;
                    ICPB_WHY_SET(3);
                }
            } else {
                ICPB_WHY_SET(2);
            }
        } else {
            ICPB_WHY_SET(1);
        }
    }

    // ---- B. queue the scan items (a lane's items are consecutive), then scan them round-robin ----
    uint32_t nsub = 0;
    if (elig) {
#pragma unroll
        for (int c = 0; c < 8; ++c) nsub += (ecnt[c] + GW_SUB - 1) / GW_SUB;
    }
    uint32_t off_end = nsub;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, off_end, o);
        if (lane >= o) off_end += v;
    }
    const uint32_t off_begin = off_end - nsub;
    const bool fits = off_end <= (uint32_t)GW_QCAP;
    if (!fits) {
        if (elig) ICPB_WHY_SET(4);
        elig = false;
    }
    const uint32_t total = __reduce_max_sync(FULL, fits ? off_end : 0u);
    if (elig) {
        uint32_t o = off_begin;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            for (uint32_t k = 0; k < ecnt[c]; k += GW_SUB) {
                const uint32_t m = ecnt[c] - k;
                queue[o++] = make_uint2(ept[c] + k, (m < (uint32_t)GW_SUB ? m : (uint32_t)GW_SUB) | ((uint32_t)lane << 8));
            }
        }
        GlSlot s;
        s.qx = qx; s.qy = qy; s.qz = qz; s.bound = bound;
        slot[lane] = s;
    }
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t j = base + lane;
        const bool has = j < total;
        const uint2 it = has ? queue[j] : make_uint2(0u, 0u);
        const uint32_t cnt = it.y & 0xFFu;
        const GlSlot s = slot[(it.y >> 8) & 31u];
        double best = ICPB_INF, second = ICPB_INF;
        uint32_t bpos = NONE;
        const bool second_batch = __any_sync(FULL, cnt > 4u);
#pragma unroll
        for (int b = 0; b < GW_SUB; b += 4) {
            if (b == 0 || second_batch) {
                double px[4], py[4], pz[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    px[t] = py[t] = pz[t] = 0.0;
                    uint32_t pidx;
                    if ((uint32_t)(b + t) < cnt) load_point(A.pts, it.x + (uint32_t)(b + t), px[t], py[t], pz[t], pidx);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double v = sumsq3(dsub(px[t], s.qx), dsub(py[t], s.qy), dsub(pz[t], s.qz));
                    if ((uint32_t)(b + t) < cnt && v <= s.bound) {
                        if (v < best) {
                            second = best;
                            best = v;
                            bpos = it.x + (uint32_t)(b + t);
                        } else if (v < second) {
                            second = v;
                        }
                    }
                }
            }
        }
        if (has) {
            rbest[j] = best;
            queue[j] = make_uint2(bpos, (bpos != NONE && !(second > dmul(best, 1.0 + 9.094947017729282e-13))) ? 1u : 0u);
        }
    }
    __syncwarp();

    // ---- C. merge the own items; unique minimum with margin 2^-40 => the reference's answer ----
    bool settled = false;
    if (elig) {
        double gb = ICPB_INF, gs = ICPB_INF;
        uint32_t gpos = NONE, gtie = 0u;
        for (uint32_t k = off_begin; k < off_end; ++k) {
            const double b = rbest[k];
            const uint2 r = queue[k];
            if (b < gb) {
                gs = gb;
                gb = b;
                gpos = r.x;
                gtie = r.y;
            } else if (b < gs) {
                gs = b;
            }
        }
        if (gpos != NONE && gtie == 0u && gs > dmul(gb, 1.0 + 9.094947017729282e-13)) {
            settled = true;
            A.pos_out[i] = gpos;
            A.dist_out[i] = dsqrt(gb);  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
        }
    }
    const unsigned pend = __ballot_sync(FULL, active && !settled);
    if (pend) {
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(A.work_count, (unsigned int)__popc(pend));
        at = __shfl_sync(FULL, at, 0);
        if (active && !settled) A.worklist[at + __popc(pend & ((1u << lane) - 1u))] = (uint32_t)i;
    }
    if (A.counters) {  // profiling / tests only
        const unsigned ok = __ballot_sync(FULL, settled);
        uint32_t cand = 0;
        if (elig) {
#pragma unroll
            for (int c = 0; c < 8; ++c) cand += ecnt[c];
        }
        cand = __reduce_add_sync(FULL, cand);
        if (lane == 0) {
            if (ok) atomicAdd(&A.counters[0], (unsigned long long)__popc(ok));
            if (pend) atomicAdd(&A.counters[2], (unsigned long long)__popc(pend));
            atomicAdd(&A.counters[3], (unsigned long long)cand);
            atomicAdd(&A.counters[4], (unsigned long long)total);
        }
#ifdef ICPB_WHY
        if (active && !settled && why == 0) why = 5;
#pragma unroll
        for (int r = 1; r <= 5; ++r) {
            const unsigned m = __ballot_sync(FULL, active && !settled && why == r);
            if (lane == 0 && m) atomicAdd(&A.counters[8 + r], (unsigned long long)__popc(m));
        }
#endif
    }
}

int nn_group_lean_launch(Ctx* c, const NNArgs& A) {
    const int blocks = (int)((A.n + GW_THREADS - 1) / GW_THREADS);
    if (A.prev_pos)
        nn_group_lean_kernel<false><<<blocks, GW_THREADS, 0, c->stream>>>(A);
    else
        nn_group_lean_kernel<true><<<blocks, GW_THREADS, 0, c->stream>>>(A);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
