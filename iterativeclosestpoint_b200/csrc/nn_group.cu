// Exact nearest-neighbour query, cell walk with the candidate scan balanced over the warp (mode 4).
// Same job and same answers as nn.cu (replaces Octree::findNearest / searchNearest, core/octree.cpp:128-184, inside the
// per-point loop of core/icpengine.cpp:172-184); the exactness argument is the one at the top of nn.cu:
// find the exact minimum of s over every point the search ball can contain, prove that it is unique by a margin of
// 2^-40, otherwise leave the query to the literal traversal.
//
// Why a second kernel: in the one-thread-per-query walk (nn_common.cuh: cell_walk) a lane scans 1..8 cells of 1..64
// points, so a warp runs as long as its unluckiest lane (ncu: 14 of 32 lanes active on average).  Here the warp
//   A. lets every lane set up its own query (move it by the pending transform, radius from last iteration's match,
//      grid level, the block of <= 3 x 3 x 3 cells its ball touches; the cells are taken in rounds of eight -- one
//      round unless a lane's ball is wider than a cell -- with the round's eight cell entries loaded at once),
//   B. cuts the non-empty cells into scan items of at most GW_SUB consecutive points, queued in shared memory,
//      and deals the items out round-robin: every lane scans one item per trip, whichever query it belongs to,
//   C. hands each item's (best, position, tie flag) back; the owning lane merges its own items in order.
// Temporal skip: every settled query also records lb = a lower bound (rounded down) on its distance to every target point
// OTHER than its match -- the smaller of the runner-up found by the scan and the query's clearance inside the block of cells
// that was scanned (all other points lie outside that block).  Next iteration the query has moved by delta, so every other
// point is at least lb - delta away (triangle inequality); if the old match is closer than that, by a margin far above the
// 2^-40 the exactness argument needs, it is still the unique nearest neighbour and the search is skipped altogether.
// In the converged regime (delta -> 0, match distance << point spacing) most queries take this exit.
// Queries it cannot settle (no seed, a ball over more than three cells along an axis, a crowded cell that is entered
// through the search tree, no unique minimum, queue overflow) go on a work list for the per-thread kernel (nn.cu).
#include "nn_common.cuh"

namespace icpb {

constexpr int GW_THREADS = 128;
constexpr int GW_WARPS = GW_THREADS / 32;
constexpr int GW_QCAP = 256;  // scan items per warp
constexpr int GW_SUB = 8;     // points per scan item (two batches of four loads in flight)

struct __align__(16) GwSlot {
    double qx, qy, qz, pad;
};

__global__ void __launch_bounds__(GW_THREADS) nn_group_kernel(const NNArgs A) {
    __shared__ GwSlot slot_all[GW_WARPS][32];
    __shared__ uint2 queue_all[GW_WARPS][GW_QCAP];   // item: x = first point, y = count | owner lane << 8; result: x = position, y = tie
    __shared__ double rbest_all[GW_WARPS][GW_QCAP];  // result: smallest s of the item
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    GwSlot* slot = slot_all[w];
    uint2* queue = queue_all[w];
    double* rbest = rbest_all[w];
    const long long i = (long long)blockIdx.x * GW_THREADS + threadIdx.x;
    const bool active = i < A.n;

    // ---- A. own query ----
    double qx = 0.0, qy = 0.0, qz = 0.0, clear = 0.0;
    bool elig = false, kept = false;
    int x0 = 0, y0 = 0, z0 = 0, nx = 0, ny = 0, nz = 0;  // block of cells the ball touches (<= 3 per axis)
    GridView V = grid_view(A, 0);
    if (active) {
        qx = A.sx[i];
        qy = A.sy[i];
        qz = A.sz[i];
        double moved = 0.0;  // upper bound on how far this query moved since its lb was recorded
        if (A.apply_pending && A.state->have_T) {
            const double ox = qx, oy = qy, oz = qz;
            apply_T_point(A.state->T_pending, qx, qy, qz);
            A.ox[i] = qx;
            A.oy[i] = qy;
            A.oz[i] = qz;
            moved = dmul(dsqrt(sumsq3(dsub(qx, ox), dsub(qy, oy), dsub(qz, oz))), 1.0 + 1e-9);
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
                if (A.lb_io) {
                    // every other target point is at least lb - moved away; the old match is closer => it stays
                    const double lbn = dsub((double)A.lb_io[i], moved);
                    if (dsqrt(Sd) < dmul(lbn, 1.0 - 1e-6)) {
                        kept = true;
                        if (A.pos_out != A.prev_pos) A.pos_out[i] = pp;
                        A.dist_out[i] = dsqrt(Sd);  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
                        A.lb_io[i] = __double2float_rd(lbn);
                    }
                }
            }
            if (!(Sd < 1e19)) Sd = walk_seed(A, qx, qy, qz);
        }
        if (Sd < 1e19 && !kept) {
            // the same ball, level and cell range as cell_walk (nn_common.cuh)
            const double r = dmul(dsqrt(Sd), 1.0 + 9.5367431640625e-07);  // sqrt(S) (1 + 2^-20)
            const double e = dadd(r, A.geps);
            V = grid_view(A, grid_level_for_width(A, dmul(e, 2.0), A.gbias));
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
                // clearance of q inside the block of cells [x0..x1] x [y0..y1] x [z0..z1]: every point outside the block is
                // farther than this (2 geps: the cells' bisection boundaries versus org + k * edge)
                const double ed = A.gedge[V.level - A.glmin];
                const double cx = fmin(dsub(qx, dadd(A.gorg[0], dmul((double)x0, ed))), dsub(dadd(A.gorg[0], dmul((double)(x1 + 1), ed)), qx));
                const double cy = fmin(dsub(qy, dadd(A.gorg[1], dmul((double)y0, ed))), dsub(dadd(A.gorg[1], dmul((double)(y1 + 1), ed)), qy));
                const double cz = fmin(dsub(qz, dadd(A.gorg[2], dmul((double)z0, ed))), dsub(dadd(A.gorg[2], dmul((double)(z1 + 1), ed)), qz));
                clear = dsub(fmin(cx, fmin(cy, cz)), dmul(A.geps, 2.0));
            }
        }
    }
    if (elig) {
        GwSlot s;
        s.qx = qx; s.qy = qy; s.qz = qz; s.pad = 0.0;
        slot[lane] = s;
    }
    const int ncell = elig ? nx * ny * nz : 0;                       // <= 27
    const int nchunk = __reduce_max_sync(FULL, (ncell + 7) >> 3);    // cells are taken eight at a time (usually one round)
    const int nxy = nx * ny;
    // c -> (dx, dy, dz) without integer division: (c * ceil(256 / d)) >> 8 == c / d for c < 27, d in {1, 2, 3, 4, 6, 9}
    const int mxy = elig ? (256 + nxy - 1) / nxy : 256, mx = elig ? (256 + nx - 1) / nx : 256;

    // running result of the own query over all rounds
    double gb = ICPB_INF, gs = ICPB_INF;
    uint32_t gpos = NONE, gtie = 0x80000000u;
    uint32_t cand_count = 0, item_count = 0;

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
                nsub += (ecnt[j] + GW_SUB - 1) / GW_SUB;
                cand_count += ecnt[j];
            }
        }
        uint32_t off_end = nsub;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, off_end, o);
            if (lane >= o) off_end += v;
        }
        const uint32_t off_begin = off_end - nsub;
        const bool fits = off_end <= (uint32_t)GW_QCAP;
        if (!fits) elig = false;
        const uint32_t total = __reduce_max_sync(FULL, fits ? off_end : 0u);
        item_count += total;
        if (elig) {
            uint32_t o = off_begin;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                for (uint32_t k = 0; k < ecnt[j]; k += GW_SUB) {
                    const uint32_t m = ecnt[j] - k;
                    queue[o++] = make_uint2(ept[j] + k, (m < (uint32_t)GW_SUB ? m : (uint32_t)GW_SUB) | ((uint32_t)lane << 8));
                }
            }
        }
        __syncwarp();
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t j = base + lane;
            const bool has = j < total;
            const uint2 it = has ? queue[j] : make_uint2(0u, 0u);
            const uint32_t cnt = it.y & 0xFFu;
            const GwSlot s = slot[(it.y >> 8) & 31u];
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
                        if ((uint32_t)(b + t) < cnt) {  // every scanned point counts: the runner-up feeds the temporal bound
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
                // y: the item's runner-up rounded down to a float (sign bit free) | bit 31: runner-up within 2^-40 of the best
                const uint32_t tie = (bpos != NONE && !(second > dmul(best, 1.0 + 9.094947017729282e-13))) ? 0x80000000u : 0u;
                queue[j] = make_uint2(bpos, __float_as_uint(__double2float_rd(second)) | tie);
            }
        }
        __syncwarp();

        // ---- C1. fold the own items of this round into the running result ----
        if (elig) {
            for (uint32_t k = off_begin; k < off_end; ++k) {
                const double b = rbest[k];
                const uint2 r = queue[k];
                // runner-up over everything scanned = min(second-best item, every better item's own runner-up)
                const double rs = (double)__uint_as_float(r.y & 0x7FFFFFFFu);
                if (b < gb) {
                    gs = fmin(gb, rs);
                    gb = b;
                    gpos = r.x;
                    gtie = r.y;
                } else {
                    gs = fmin(gs, b);
                }
            }
        }
        __syncwarp();  // the queue is rewritten by the next round
    }

    // ---- C2. unique minimum with margin 2^-40 => the reference's answer ----
    bool settled = kept;
    if (elig && gpos != NONE && (gtie & 0x80000000u) == 0u && gs > dmul(gb, 1.0 + 9.094947017729282e-13)) {
        settled = true;
        A.pos_out[i] = gpos;
        A.dist_out[i] = dsqrt(gb);  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
        if (A.lb_io) {
            const double others = fmin(dsqrt(gs), clear);
            A.lb_io[i] = others > 0.0 ? __double2float_rd(dmul(others, 1.0 - 1e-9)) : 0.0f;
        }
    }
    if (A.lb_io && active && !settled) A.lb_io[i] = 0.0f;  // the per-thread kernel records no bound
    const unsigned pend = __ballot_sync(FULL, active && !settled);
    if (pend) {
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(A.work_count, (unsigned int)__popc(pend));
        at = __shfl_sync(FULL, at, 0);
        if (active && !settled) A.worklist[at + __popc(pend & ((1u << lane) - 1u))] = (uint32_t)i;
    }
    if (A.counters) {  // profiling / tests only
        const unsigned ok = __ballot_sync(FULL, settled);
        const unsigned kp = __ballot_sync(FULL, kept);
        const uint32_t cand = __reduce_add_sync(FULL, elig ? cand_count : 0u);
        if (lane == 0) {
            if (ok) atomicAdd(&A.counters[0], (unsigned long long)__popc(ok));
            if (pend) atomicAdd(&A.counters[2], (unsigned long long)__popc(pend));
            atomicAdd(&A.counters[3], (unsigned long long)cand);
            atomicAdd(&A.counters[4], (unsigned long long)item_count);
            if (kp) atomicAdd(&A.counters[5], (unsigned long long)__popc(kp));
        }
    }
}

int nn_group_launch(Ctx* c, const NNArgs& A) {
    const int blocks = (int)((A.n + GW_THREADS - 1) / GW_THREADS);
    nn_group_kernel<<<blocks, GW_THREADS, 0, c->stream>>>(A);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
