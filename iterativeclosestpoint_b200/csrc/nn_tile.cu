// Exact nearest neighbour, one TILE of 32 consecutive (Morton-ordered) queries per warp.
// Replaces the per-point loop around Octree::findNearest (core/icpengine.cpp:172-184, octree.cpp:128-184).
//
// Why tiles: a per-thread tree descent keeps ~6 of 32 lanes busy (ncu, profiles/r01_nn_kernel_thread.md); the
// queries of one warp are neighbours in space, so their searches touch almost the same leaves.  The warp therefore
//   1. applies the pending transform, takes the bounding box B of its 32 queries and a radius r that is known to
//      reach every lane's nearest neighbour (the distance to last iteration's match, or a guess that is verified);
//   2. collects, cooperatively, every octree leaf whose box meets E = B (+) r  (stack of node ids in shared memory,
//      one node per lane per round, child boxes rebuilt by the reference's bisection) and stages the leaves' points
//      in shared memory (32-byte records, coalesced 16-byte loads);
//   3. lets every lane scan ALL staged points -- uniform control flow, shared-memory broadcast reads -- keeping the
//      smallest and second smallest value of the reference's squared-distance expression;
//   4. proves each lane's answer: every target point that was not scanned lies in a box disjoint from E, hence is
//      farther than the lane's clearance c inside E; with c^2 > best (1 + 2^-38) no unscanned point can come within
//      best (1 + 2^-39), and with second > best (1 + 2^-40) the minimum is unique with margin, which makes the
//      reference's own traversal return the same point whatever order it visits things in (argument in nn.cu).
// Lanes that cannot be proven (radius too small, exact or 1-ulp ties, non-finite input) first get another pass with
// the radius their own scan result calls for, then fall back to the per-thread search / the literal reference
// traversal of nn_common.cuh.  Results are therefore identical to the reference's for every query.
#include "nn_common.cuh"

namespace icpb {

constexpr int TW = 4;                      // tiles (warps) per CTA
constexpr int TILE_THREADS = TW * 32;
constexpr int CAND_CAP = 192;              // staged candidates per scan (32 B each): 6 KB per warp
constexpr int STK_SOFT = 224;              // batch pops keep the node stack below this ...
constexpr int STK_ALLOC = 384;             // ... single pops can add 7 per level on top (7 * 22 = 154)
constexpr int TERMINAL_PTS = 16;           // subtrees with at most this many points are staged whole
constexpr int MAX_PASSES = 3;
constexpr int TILE_MAX_CELLS = 160;        // boxes meeting more grid cells are collected through the tree
static_assert(TILE_MAX_CELLS + 32 <= STK_SOFT, "every over-full cell of a tile must fit on the node stack");
constexpr int CAND_BUDGET = 8192;          // a pass that would stage more than this gives the tile up
constexpr unsigned FULL = 0xffffffffu;

struct __align__(16) Cand {
    double x, y, z;
    unsigned long long pos;  // position in the sorted target
};
static_assert(sizeof(Cand) == 32, "Cand must be 32 bytes");
static_assert(sizeof(Cand) * CAND_CAP >= sizeof(uint2) * NN_MAX_LEVELS * 32, "slow-path stack aliases the candidate buffer");

__device__ __forceinline__ double wmin(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int wscan_incl(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// existing octants of `nd` whose box meets the closed box [elo, ehi]; child boxes by the reference's bisection
// (lower half [lo, mid], upper half [mid, hi], octree.cpp:97-99,115-120)
__device__ __forceinline__ uint32_t overlap_octants(const NodeRegs& nd, const double* elo, const double* ehi) {
    uint32_t m = nd.meta & 0xFFu;
    const uint32_t LOW[3] = {0x55u, 0x33u, 0x0Fu};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double mid = dmul(dadd(nd.lo[a], nd.hi[a]), 0.5);
        const bool low_ok = elo[a] <= mid && ehi[a] >= nd.lo[a];
        const bool high_ok = ehi[a] >= mid && elo[a] <= nd.hi[a];
        m &= (low_ok ? LOW[a] : 0u) | (high_ok ? (~LOW[a] & 0xFFu) : 0u);
    }
    return m;
}

__device__ __forceinline__ bool box_inside(const NodeRegs& nd, const double* elo, const double* ehi) {
    return nd.lo[0] >= elo[0] && nd.hi[0] <= ehi[0] && nd.lo[1] >= elo[1] && nd.hi[1] <= ehi[1] && nd.lo[2] >= elo[2] &&
           nd.hi[2] <= ehi[2];
}

__global__ void __launch_bounds__(TILE_THREADS, 6) nn_tile_kernel(const NNArgs A) {
    __shared__ Cand s_cand[TW][CAND_CAP];
    __shared__ uint32_t s_stk[TW][STK_ALLOC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tile = (long long)blockIdx.x * TW + warp;
    if (tile * 32 >= A.n) return;  // warp-uniform; the kernel has no block-wide barrier
    Cand* cand = s_cand[warp];
    uint32_t* stk = s_stk[warp];
    const long long i = tile * 32 + lane;
    const bool active = i < A.n;

    double qx = 0.0, qy = 0.0, qz = 0.0;
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
    }
    const bool finite_q = active && isfinite(qx) && isfinite(qy) && isfinite(qz);

    // ---- seeds: squared distance to last iteration's match (a real target point bounds the answer) -------------
    uint32_t pp = NONE;
    double Sd = ICPB_INF;
    if (finite_q && A.prev_pos) {
        pp = A.prev_pos[i];
        if (pp != NONE) {
            double px, py, pz;
            uint32_t pidx;
            load_point(A.pts, pp, px, py, pz, pidx);
            Sd = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
        }
    }
    const bool seeded = Sd < 1e19;

    // ---- the tile's box --------------------------------------------------------------------------------------
    double bl[3], bh[3];
    bl[0] = wmin(finite_q ? qx : ICPB_INF);
    bl[1] = wmin(finite_q ? qy : ICPB_INF);
    bl[2] = wmin(finite_q ? qz : ICPB_INF);
    bh[0] = wmax(finite_q ? qx : -ICPB_INF);
    bh[1] = wmax(finite_q ? qy : -ICPB_INF);
    bh[2] = wmax(finite_q ? qz : -ICPB_INF);
    const int n_fin = __popc(__ballot_sync(FULL, finite_q));
    const double ex = bh[0] - bl[0], ey = bh[1] - bl[1], ez = bh[2] - bl[2];
    const double diag = sqrt(ex * ex + ey * ey + ez * ez);
    const double mag = fmax(fmax(fmax(fabs(bl[0]), fabs(bh[0])), fmax(fabs(bl[1]), fabs(bh[1]))), fmax(fabs(bl[2]), fabs(bh[2])));
    // smallest padding that still leaves every lane a strictly positive clearance (exact copies have best == 0)
    const double pad = fmax(diag * 1e-6, fmax(mag * 9.094947017729282e-13, 1e-30));

    const double GROW20 = 1.0 + 9.5367431640625e-07;  // 1 + 2^-20: absorbs the float round-down of the clearance
    double r_tile = 0.0;
    {
        const double r_i = seeded ? dmul(dsqrt(Sd), GROW20) : 0.0;
        const int n_seed = __popc(__ballot_sync(FULL, seeded));
        if (n_seed > 0) {
            const double mean_r = wsum(seeded ? r_i : 0.0) / (double)n_seed;
            const double r_cap = fmax(diag, 3.0 * mean_r);  // lanes far beyond the tile's typical radius go the slow way
            r_tile = wmax((seeded && r_i <= r_cap) ? r_i : 0.0);
            r_tile = fmax(r_tile, pad);
        } else {
            r_tile = 0.5 * fmax(ex, fmax(ey, ez));  // no seeds yet: a guess, verified below
            if (r_tile > 0.0) r_tile = fmax(r_tile, pad);
        }
    }

    bool resolved = !finite_q;  // non-finite queries: the reference accepts nothing (index 0)
    bool tie = false;
    uint32_t result = NONE;
    double best = ICPB_INF, second = ICPB_INF;
    uint32_t bpos = NONE;
    uint32_t start_node = (A.tile_node != nullptr) ? A.tile_node[tile] : 0u;
    unsigned long long scanned = 0;
    unsigned dbg_rounds = 0, dbg_passes = 0, dbg_nodes = 0, dbg_steps = 0;
    unsigned unresolved = __ballot_sync(FULL, !resolved);

    for (int pass = 0; pass < MAX_PASSES && unresolved != 0u && r_tile > 0.0 && n_fin > 0; ++pass) {
        double elo[3], ehi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            elo[a] = bl[a] - r_tile;
            ehi[a] = bh[a] + r_tile;
        }
        ++dbg_passes;
        // ---- where the candidates come from: the entry-grid cells that E meets (one 8-byte entry per lane per round);
        //      a box that meets too many cells is collected through the tree instead, from the smallest subtree
        //      that holds every leaf meeting E (all lanes walk together)
        int gx0 = 0, gy0 = 0, gz0 = 0, gnxr = 0, gnyr = 0, ncell = 0;
        bool use_grid = false;
        GridView V = grid_view(A, 0);
        if (A.grid != nullptr) {
            // finest level at which E still meets few enough cells
            for (int k = A.gnlev - 1; k >= 0 && !use_grid; --k) {
                V = grid_view(A, k);
                int x0 = grid_cell_index(A, V, elo[0] - A.geps, 0, V.nx), x1 = grid_cell_index(A, V, ehi[0] + A.geps, 0, V.nx);
                int y0 = grid_cell_index(A, V, elo[1] - A.geps, 1, V.ny), y1 = grid_cell_index(A, V, ehi[1] + A.geps, 1, V.ny);
                int z0 = grid_cell_index(A, V, elo[2] - A.geps, 2, V.nz), z1 = grid_cell_index(A, V, ehi[2] + A.geps, 2, V.nz);
                x0 = max(x0, 0); y0 = max(y0, 0); z0 = max(z0, 0);
                x1 = min(x1, V.nx - 1); y1 = min(y1, V.ny - 1); z1 = min(z1, V.nz - 1);
                const long long nc = (long long)max(x1 - x0 + 1, 0) * (long long)max(y1 - y0 + 1, 0) * (long long)max(z1 - z0 + 1, 0);
                if (nc <= (long long)TILE_MAX_CELLS) {
                    use_grid = true;
                    gx0 = x0; gy0 = y0; gz0 = z0;
                    gnxr = x1 - x0 + 1;
                    gnyr = y1 - y0 + 1;
                    ncell = (int)nc;
                }
            }
        }
        if (!use_grid) {
            uint32_t n = start_node;
            NodeRegs nd;
            for (;;) {
                nd = load_node(A.nodes, n);
                ++dbg_steps;
                const bool inside = elo[0] >= nd.lo[0] && ehi[0] <= nd.hi[0] && elo[1] >= nd.lo[1] && ehi[1] <= nd.hi[1] &&
                                    elo[2] >= nd.lo[2] && ehi[2] <= nd.hi[2];
                if (inside || n == 0u) break;
                n = __ldg(A.parent + n);
            }
            for (;;) {
                const uint32_t mask = nd.meta & 0xFFu;
                if (mask == 0u || nd.npts <= (uint32_t)A.terminal_pts) break;
                const uint32_t ov = overlap_octants(nd, elo, ehi);
                if (__popc(ov) != 1) break;
                const uint32_t o = (uint32_t)(__ffs(ov) - 1);
                n = nd.child0 + __popc(mask & ((1u << o) - 1u));
                nd = load_node(A.nodes, n);
                ++dbg_steps;
            }
            start_node = n;
        }
        // ---- collect + scan ------------------------------------------------------------------------------------
        best = ICPB_INF;
        second = ICPB_INF;
        bpos = NONE;
        int S = use_grid ? 0 : 1, count = 0, total = 0, cell_base = 0;
        bool aborted = false;
        uint32_t pend_pt0 = 0;
        int pend_n = 0;
        if (lane == 0 && !use_grid) stk[0] = start_node;
        __syncwarp();
        for (;;) {
            if (!__any_sync(FULL, pend_n > 0)) {
                if (cell_base < ncell) {
                    // next 32 grid cells, one per lane
                    const int cc = cell_base + lane;
                    cell_base += 32;
                    uint32_t push = NONE;
                    if (cc < ncell) {
                        const int x = gx0 + cc % gnxr, y = gy0 + (cc / gnxr) % gnyr, z = gz0 + cc / (gnxr * gnyr);
                        const uint2 en = grid_entry(V, x, y, z);
                        const uint32_t kind = en.y >> 30;
                        bool take = kind == 1u;
                        if (kind == 2u) {  // a shallower leaf owns a block of cells: take it at the first cell shared with the range
                            const int sh = V.level - (int)((en.y >> 24) & 0x3Fu);
                            take = x == max((x >> sh) << sh, gx0) && y == max((y >> sh) << sh, gy0) && z == max((z >> sh) << sh, gz0);
                        }
                        if (take) {
                            pend_pt0 = en.x;
                            pend_n = (int)(en.y & 0xFFFFFFu);
                        }
                        if (kind == 3u) push = en.x;  // over-full cell: expand its node below
                    }
                    const unsigned pm = __ballot_sync(FULL, push != NONE);
                    if (push != NONE) stk[S + __popc(pm & ((1u << lane) - 1u))] = push;
                    S += __popc(pm);
                    ++dbg_rounds;
                    const int pincl = wscan_incl(pend_n, lane);
                    total += __shfl_sync(FULL, pincl, 31);
                    __syncwarp();
                    if (total > CAND_BUDGET) {
                        aborted = true;
                        break;
                    }
                } else if (S == 0) {
                    if (count == 0) break;
                } else {
                    // pop up to 32 nodes, one per lane
                    int k = (STK_SOFT - S) / 7;
                    k = k < 1 ? 1 : k;
                    k = k > 32 ? 32 : k;
                    k = k > S ? S : k;
                    const uint32_t node = (lane < k) ? stk[S - 1 - lane] : NONE;
                    S -= k;
                    ++dbg_rounds;
                    dbg_nodes += (unsigned)k;
                    __syncwarp();
                    uint32_t ov = 0, mask = 0, child0 = 0;
                    if (node != NONE) {
                        const NodeRegs nd = load_node(A.nodes, node);
                        mask = nd.meta & 0xFFu;
                        if (mask == 0u || nd.npts <= (uint32_t)A.terminal_pts || box_inside(nd, elo, ehi)) {
                            pend_pt0 = nd.pt0;
                            pend_n = (int)nd.npts;
                        } else {
                            ov = overlap_octants(nd, elo, ehi);
                            child0 = nd.child0;
                        }
                    }
                    const int cnt = __popc(ov);
                    const int incl = wscan_incl(cnt, lane);
                    int off = S + incl - cnt;
                    while (ov) {
                        const uint32_t o = (uint32_t)(__ffs(ov) - 1);
                        ov &= ov - 1u;
                        stk[off++] = child0 + __popc(mask & ((1u << o) - 1u));
                    }
                    S += __shfl_sync(FULL, incl, 31);
                    const int pincl = wscan_incl(pend_n, lane);
                    total += __shfl_sync(FULL, pincl, 31);
                    __syncwarp();
                    if (total > CAND_BUDGET) {
                        aborted = true;
                        break;
                    }
                }
            }
            // stage as much of the pending ranges as fits
            {
                const int incl = wscan_incl(pend_n, lane);
                const int room = CAND_CAP - count;
                const int before = incl - pend_n;
                int fit = room - before;
                fit = fit < 0 ? 0 : (fit > pend_n ? pend_n : fit);
                const int4* src = reinterpret_cast<const int4*>(A.pts + pend_pt0);
                int4* dst = reinterpret_cast<int4*>(cand + count + before);
                for (int k = 0; k < fit; ++k) {
                    const int4 a = __ldg(src + 2 * k);
                    int4 b = __ldg(src + 2 * k + 1);
                    b.z = (int)(pend_pt0 + (uint32_t)k);  // the staged record carries the sorted position
                    b.w = 0;
                    dst[2 * k] = a;
                    dst[2 * k + 1] = b;
                }
                pend_pt0 += (uint32_t)fit;
                pend_n -= fit;
                const int tot = __shfl_sync(FULL, incl, 31);
                count += tot < room ? tot : room;
                __syncwarp();
            }
            if (__any_sync(FULL, pend_n > 0) || (S == 0 && cell_base >= ncell)) {
                // every lane scans every staged point: uniform control flow, broadcast reads
#pragma unroll 4
                for (int c = 0; c < count; ++c) {
                    const double2 a = *reinterpret_cast<const double2*>(&cand[c].x);
                    const double2 b = *reinterpret_cast<const double2*>(&cand[c].z);
                    const double s = sumsq3(dsub(a.x, qx), dsub(a.y, qy), dsub(b.x, qz));
                    const uint32_t p = (uint32_t)__double_as_longlong(b.y);
                    if (s < best) {
                        second = best;
                        best = s;
                        bpos = p;
                    } else if (s < second) {
                        second = s;
                    }
                }
                scanned += (unsigned long long)count;
                count = 0;
                __syncwarp();
            }
        }
        if (aborted) break;

        // ---- proof per lane ----------------------------------------------------------------------------------
        if (!resolved && bpos != NONE && best < 1e19) {
            const double c = fmin(fmin(dsub(qx, elo[0]), dsub(ehi[0], qx)),
                                  fmin(fmin(dsub(qy, elo[1]), dsub(ehi[1], qy)), fmin(dsub(qz, elo[2]), dsub(ehi[2], qz))));
            const double cf = (c > 0.0) ? (double)__double2float_rd(c) : 0.0;
            if (cf * cf > dmul(best, 1.0 + 3.637978807091713e-12)) {  // best (1 + 2^-38)
                resolved = true;
                if (second > dmul(best, 1.0 + 9.094947017729282e-13))  // unique with margin 2^-40
                    result = bpos;
                else
                    tie = true;  // exact / 1-ulp tie: only the literal traversal knows the reference's pick
            }
        }
        unresolved = __ballot_sync(FULL, !resolved);
        if (unresolved == 0u) break;
        // ---- radius the unresolved lanes ask for -----------------------------------------------------------------
        const bool has = !resolved && bpos != NONE && best < 1e19;
        const double r_need = has ? dmul(dsqrt(best), GROW20) : 0.0;
        const int n_has = __popc(__ballot_sync(FULL, has));
        if (n_has == 0) break;
        const double mean_r = wsum(has ? r_need : 0.0) / (double)n_has;
        const double r_cap = fmax(diag, 3.0 * mean_r);
        const double r_next = wmax((has && r_need <= r_cap) ? r_need : 0.0);
        if (!(r_next > r_tile)) break;
        r_tile = r_next;
    }
    if (A.tile_node != nullptr && lane == 0) A.tile_node[tile] = start_node;

    // ---- slow path: per-thread search / literal traversal for what the tile could not prove ------------------------
    const bool slow = finite_q && (!resolved || tie);
    bool fell_back = false;
    if (__any_sync(FULL, slow)) {
        __syncwarp();
        if (slow) {
            uint2* tstk = reinterpret_cast<uint2*>(cand) + lane;
            uint32_t rn;
            const double extra = (!tie && bpos != NONE) ? best : ICPB_INF;
            result = per_thread_query<32>(A, qx, qy, qz, true, tie ? NONE : pp, NONE, extra, tstk, rn, fell_back, tie);
        }
    }

    if (active) {
        // findNearest returns index 0 when nothing was accepted (best_idx = 0 initially, octree.cpp:179)
        const uint32_t pos = (result == NONE) ? A.pos_of_idx0 : result;
        double px, py, pz;
        uint32_t pidx;
        load_point(A.pts, pos, px, py, pz, pidx);
        A.pos_out[i] = pos;
        A.dist_out[i] = dsqrt(sumsq3(dsub(qx, px), dsub(qy, py), dsub(qz, pz)));  // computeDistance (icpengine.cpp:68-74)
        if (A.node_io) A.node_io[i] = NONE;
    }
    if (A.counters) {
        const unsigned fb = __ballot_sync(FULL, fell_back);
        const unsigned sl = __ballot_sync(FULL, slow && !fell_back);
        const unsigned ac = __ballot_sync(FULL, active);
        if (lane == 0) {
            if (fb) atomicAdd(&A.counters[1], (unsigned long long)__popc(fb));
            atomicAdd(&A.counters[0], (unsigned long long)__popc(ac & ~fb));
            if (sl) atomicAdd(&A.counters[2], (unsigned long long)__popc(sl));
            atomicAdd(&A.counters[3], scanned);
            atomicAdd(&A.counters[4], (unsigned long long)dbg_rounds);
            atomicAdd(&A.counters[5], (unsigned long long)dbg_passes);
            atomicAdd(&A.counters[6], (unsigned long long)dbg_nodes);
            atomicAdd(&A.counters[7], (unsigned long long)dbg_steps);
        }
    }
}

int nn_tile_launch(Ctx* c, const NNArgs& A) {
    const long long tiles = (A.n + 31) / 32;
    const int blocks = (int)((tiles + TW - 1) / TW);
    nn_tile_kernel<<<blocks, TILE_THREADS, 0, c->stream>>>(A);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
