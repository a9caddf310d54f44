// Exact nearest-neighbour query, BOX SEARCH with carried candidate lists (mode 7, the default).
// Same job and same answers as nn.cu (replaces Octree::findNearest / searchNearest, core/octree.cpp:128-184, inside the
// per-point loop of core/icpengine.cpp:172-184); the exactness argument is the one at the top of nn.cu: find the exact
// minimum of s (the reference's squared-distance expression) over every target point the search ball can contain, prove that
// it is unique by a margin of 2^-40, otherwise leave the query to the literal traversal.
//
// Why: in the per-query walks (nn_common.cuh: cell_walk, and the balanced form of it) every lane pulls its own 32-byte
// sectors -- ~13 L1 wavefronts per query -- and the kernel is bound by L1 lookups and issue slots at a sixth of HBM speed.
// Here the source is cut into GROUPS of up to 8 queries that are consecutive in the internal order and start out inside one
// cell of the cell grid (build.cu).  A group owns ONE candidate list = every target point inside the group's BOX, recentred
// on the box and stored as FP32 (x, y, z, |p|^2) plus the points' positions.  The target never changes and the source moves
// by centimetres per iteration, so the lists live in device memory and are used again for as long as they cover the queries:
//
//   nn_list_kernel   one query per thread, streaming.  Moves the query by the pending transform (core/icpengine.cpp:345, fused
//       into the load), scans its group's list -- the lanes of a group read the same 16-byte candidates in the same
//       instruction, so the scan runs out of L1 at one sector per group and step -- with s' = |p|^2 - 2 q.p in FP32 (3 FMA per
//       candidate), tracking the smallest, the second smallest and the winner's slot; evaluates the winner in FP64 with the
//       reference's expression; and accepts it if the ball of that radius around the query lies inside the stored box: then
//       no point outside the list can be nearer.  A query whose ball leaves the box asks for its group to be rebuilt.
//   nn_build_kernel  one GROUP per thread (a group is cell-sized: a handful of cells, a few dozen points).  Box = union of the
//       queries' balls (radius = distance to last iteration's match, a real target point, plus a skin that keeps the list valid
//       while the cloud moves; a guess for a query without a match), cell-grid entries under it, points inside it -> list.
//       Then nn_list_kernel runs over the queries of the rebuilt groups.
//   FP32 with FP64 recheck: |(s'_j - s'_k) - (s_j - s_k)| <= 114 * 2^-24 * W^2 (W = largest half box edge), so a gap
//       above 3 * 2^-16 * W^2 proves the winner unique in FP64 by far more than 2^-40.  A smaller gap re-evaluates in FP64 every
//       candidate within that margin of the best and applies the 2^-40 test to the exact values.
// Queries that stay open (no unique minimum, a ball wider than `emax`, a box over too many cells or points, a seedless query
// whose nearest point lies outside the box) go on the query work list for the per-thread kernel (nn.cu).
#include "nn_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace icpb {

constexpr int LS_THREADS = 256;
constexpr int BD_THREADS = 128;
constexpr int BX_G = 8;               // queries per group at most (build.cu: QGROUP_MAX)
constexpr int BX_CELLCAP = 64;        // cell-grid entries under one group box
constexpr uint32_t BX_NOLIST = 0xFFFFFFFFu;

// one candidate against the running (best, second, slot): s' = |p|^2 - 2 q.p
#define BX_SCAN_STEP(c, k)                                                              \
    {                                                                                   \
        const float s_ = fmaf(m2x, (c).x, fmaf(m2y, (c).y, fmaf(m2z, (c).z, (c).w)));   \
        second = fminf(second, fmaxf(s_, best));                                        \
        if (s_ < best) slot = (int)(k);                                                 \
        best = fminf(best, s_);                                                         \
    }

// ---------------------------------------------------------------------------------------------------------------------------
// the streaming pass over the carried lists: one query per thread
// ---------------------------------------------------------------------------------------------------------------------------
// INDIRECT: the queries are those of the groups on the group work list (8 threads per listed group), already moved.
// A.list_final: a query the list cannot settle goes to the per-thread kernel (its group has just been rebuilt, or cannot be);
// otherwise it asks for its group to be rebuilt (once per group: the flag word holds the iteration's epoch).
template <bool INDIRECT>
__global__ void __launch_bounds__(LS_THREADS) nn_list_kernel(const NNArgs A) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float INF32 = __int_as_float(0x7f800000);
    if (A.state && A.state->exit_code != 0) return;  // the loop has ended: iterations enqueued ahead do nothing
    const long long t = (long long)blockIdx.x * LS_THREADS + threadIdx.x;
    long long i = t;
    bool active;
    uint32_t gid = 0;
    if (INDIRECT) {
        const long long slot = t >> 3;
        active = slot < (long long)*A.group_count;
        if (active) {
            gid = A.group_list[slot];
            const uint32_t qb = __ldg(A.gstart + gid), qe = __ldg(A.gstart + gid + 1);
            i = (long long)qb + (t & 7);
            active = i < (long long)qe;
        }
    } else {
        active = i < A.n;
        if (active) gid = __ldg(A.gidx + i);
    }
    // solve_step: the registration has closed in since the lists were built, tighter ones pay -> everything is rebuilt
    const bool rebuild_all = !INDIRECT && !A.list_final && A.state && A.state->rebuild_all != 0;

    bool want_rebuild = false, open = false;
    if (active) {
        double qx = A.sx[i], qy = A.sy[i], qz = A.sz[i];
        const uint32_t pp = A.prev_pos ? A.prev_pos[i] : NONE;
        const int4* hp = reinterpret_cast<const int4*>(A.lhdr + gid);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
        double ppx = 0.0, ppy = 0.0, ppz = 0.0;  // last iteration's match, fetched early: most often it wins again
        if (pp != NONE) {
            uint32_t pidx;
            load_point(A.pts, pp, ppx, ppy, ppz, pidx);
        }
        const double og0 = __hiloint2double(h0.y, h0.x), og1 = __hiloint2double(h0.w, h0.z), og2 = __hiloint2double(h1.y, h1.x);
        const float hf0 = __int_as_float(h1.z), hf1 = __int_as_float(h1.w), hf2 = __int_as_float(h2.x);
        const uint32_t hcount = (uint32_t)h2.y;
        if (!INDIRECT && A.apply_pending && A.state->have_T) {
            apply_T_point(A.state->T_pending, qx, qy, qz);
            A.ox[i] = qx;
            A.oy[i] = qy;
            A.oz[i] = qz;
        }
        // clearance of the query inside the stored box (negative: outside; NaN for a non-finite query)
        const double dx0 = dsub(qx, og0), dx1 = dsub(qy, og1), dx2 = dsub(qz, og2);
        const double clr = fmin(fmin((double)hf0 - fabs(dx0), (double)hf1 - fabs(dx1)), (double)hf2 - fabs(dx2));
        const bool use_list = hcount <= (uint32_t)BOX_LIST_CAP && clr > 0.0 && !rebuild_all;
        open = true;
        if (use_list) {
            // ---- scan the group's list ----
            const float vx = (float)dx0, vy = (float)dx1, vz = (float)dx2;
            const float4* lc = A.lcand + (long long)gid * BOX_LIST_CAP;
            const float m2x = -2.0f * vx, m2y = -2.0f * vy, m2z = -2.0f * vz;
            float best = INF32, second = INF32;
            int slot = -1;
#pragma unroll 4
            for (unsigned int k = 0; k < hcount; ++k) {
                const float4 c = __ldg(lc + k);
                BX_SCAN_STEP(c, k);
            }
            if (slot >= 0) {
                const float hmax = fmaxf(fmaxf(hf0, hf1), hf2);
                const float margin = 4.57763671875e-05f * hmax * hmax;  // 3 * 2^-16 * W^2
                const uint32_t* pos_of = A.lpos + (long long)gid * BOX_LIST_CAP;
                uint32_t win = NONE;
                double s_win = 0.0;
                bool tie = false;
                if (second - best > margin) {
                    win = __ldg(pos_of + slot);
                    double px = ppx, py = ppy, pz = ppz;
                    if (win != pp) {
                        uint32_t pidx;
                        load_point(A.pts, win, px, py, pz, pidx);
                    }
                    s_win = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
                } else {
                    // FP64 recheck of every candidate within the margin of the FP32 minimum
                    const float lim = best + margin;
                    double b64 = ICPB_INF, s64 = ICPB_INF;
                    for (unsigned int k = 0; k < hcount; ++k) {
                        const float4 c = __ldg(lc + k);
                        const float s = fmaf(m2x, c.x, fmaf(m2y, c.y, fmaf(m2z, c.z, c.w)));
                        if (s <= lim) {
                            const uint32_t pos = __ldg(pos_of + k);
                            double px, py, pz;
                            uint32_t pidx;
                            load_point(A.pts, pos, px, py, pz, pidx);
                            const double v = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
                            if (v < b64) {
                                s64 = b64;
                                b64 = v;
                                win = pos;
                            } else if (v < s64) {
                                s64 = v;
                            }
                        }
                    }
                    s_win = b64;
                    tie = !(win != NONE && s64 > dmul(b64, 1.0 + 9.094947017729282e-13));
                }
                if (win != NONE) {
                    // The list holds every target point within `clr` of the query along each axis, so the answer stands if the
                    // ball of the distance found lies inside.  (sqrt_upper >= sqrt(s) (1 + 2^-18): a point that ties with the
                    // winner within 2^-40 is inside too.)
                    const double rho = dadd(sqrt_upper(s_win), A.geps);
                    if (rho <= clr) {
                        if (!tie) {
                            open = false;
                            A.pos_out[i] = win;
                            A.dist_out[i] = dsqrt(s_win);  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
                        }
                        // (a tie inside a covering list: a rebuild cannot help, the literal traversal decides)
                    } else {
                        want_rebuild = true;
                    }
                } else {
                    want_rebuild = true;
                }
            } else {
                want_rebuild = true;  // an empty list: the box holds no target point
            }
        } else {
            want_rebuild = true;
        }
        if (A.list_final) want_rebuild = false;
        if (want_rebuild) open = false;  // the builder's
    }
    // groups to rebuild: the first query to ask puts the group on the list (one global counter update per block)
    {
        __shared__ unsigned int s_n, s_base;
        if (threadIdx.x == 0) s_n = 0u;
        __syncthreads();
        unsigned int mine = 0xFFFFFFFFu;
        if (want_rebuild && atomicExch(A.gflag + gid, A.epoch) != A.epoch) mine = atomicAdd(&s_n, 1u);
        __syncthreads();
        if (threadIdx.x == 0 && s_n) s_base = atomicAdd(A.group_count, s_n);
        __syncthreads();
        if (mine != 0xFFFFFFFFu) A.group_list[s_base + mine] = gid;
    }
    // open queries go to the per-thread kernel
    const unsigned pend = __ballot_sync(FULL, open);
    if (pend) {
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(A.work_count, (unsigned int)__popc(pend));
        at = __shfl_sync(FULL, at, 0);
        if (open) A.worklist[at + __popc(pend & lt_mask)] = (uint32_t)i;
    }
    if (A.counters) {  // profiling / tests only
        const unsigned okb = __ballot_sync(FULL, active && !want_rebuild && !open);
        if (lane == 0) {
            if (okb) atomicAdd(&A.counters[0], (unsigned long long)__popc(okb));
            if (okb && !INDIRECT && !A.list_final) atomicAdd(&A.counters[5], (unsigned long long)__popc(okb));
            if (pend) atomicAdd(&A.counters[2], (unsigned long long)__popc(pend));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// the builder: one group per thread
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BD_THREADS) nn_build_kernel(const NNArgs A) {
    if (A.state && A.state->exit_code != 0) return;  // the loop has ended: iterations enqueued ahead do nothing
    const long long t = (long long)blockIdx.x * BD_THREADS + threadIdx.x;
    const long long n_slots = A.group_list ? (long long)*A.group_count : A.n_groups;
    if (t >= n_slots) return;
    const long long gid = A.group_list ? (long long)A.group_list[t] : t;
    const uint32_t qb = __ldg(A.gstart + gid), qe = __ldg(A.gstart + gid + 1);
    const float INF32 = __int_as_float(0x7f800000);

    // ---- the box: union of the queries' balls, relative to the grid origin, FP32 rounded outwards ----
    float flo[3] = {INF32, INF32, INF32}, fhi[3] = {-INF32, -INF32, -INF32};
    bool any = false;
    for (uint32_t j = qb; j < qe; ++j) {
        const double qx = A.sx[j], qy = A.sy[j], qz = A.sz[j];
        if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) continue;
        const uint32_t pp = A.prev_pos ? A.prev_pos[j] : NONE;
        double e = A.box_guess;  // no match yet: a guess that nn_list_kernel checks against what it finds
        if (pp != NONE) {
            // ball radius >= the distance to a real target point (then the box certainly holds the nearest point), plus the
            // skin that keeps the list usable while the query moves
            double px, py, pz;
            uint32_t pidx;
            load_point(A.pts, pp, px, py, pz, pidx);
            const double Sd = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
            e = (Sd < 1e19) ? dadd(dadd(sqrt_upper(Sd), dmul(A.geps, 2.0)), A.box_skin) : A.box_guess;
        }
        if (!(e <= A.box_emax)) continue;  // a wide ball: the per-thread search takes this query
        const double a3[3] = {dsub(qx, A.gorg[0]), dsub(qy, A.gorg[1]), dsub(qz, A.gorg[2])};
        const double far = A.gcube;
        if (!(a3[0] > -far && a3[0] < far + far && a3[1] > -far && a3[1] < far + far && a3[2] > -far && a3[2] < far + far)) continue;
        any = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            flo[c] = fminf(flo[c], __double2float_rd(dsub(a3[c], e)));
            fhi[c] = fmaxf(fhi[c], __double2float_ru(dadd(a3[c], e)));
        }
    }
    BoxListHdr H;
    H.og[0] = H.og[1] = H.og[2] = 0.0;
    H.hf[0] = H.hf[1] = H.hf[2] = 0.f;
    H.count = BX_NOLIST;
    H.pad[0] = H.pad[1] = 0u;
    if (any) {
        float cf[3], hf[3];
        int c0[3], cn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cf[c] = 0.5f * (flo[c] + fhi[c]);
            hf[c] = __fmul_ru(fmaxf(__fsub_ru(fhi[c], cf[c]), __fsub_ru(cf[c], flo[c])), 1.0000005f);
            // cells under [cf - hf, cf + hf]: FP32 with outward rounding (a cell too many is harmless)
            const float lo = __fmul_rd(__fsub_rd(cf[c], hf[c]), A.cinv_lo), hi = __fmul_ru(__fadd_ru(cf[c], hf[c]), A.cinv_hi);
            const float dimc = (float)A.cdim[c];
            const int i0 = (int)fminf(fmaxf(floorf(lo), 0.0f), dimc);        // dim => past the grid
            const int i1 = (int)fminf(fmaxf(floorf(hi), -1.0f), dimc - 1.0f);
            c0[c] = i0;
            cn[c] = i1 - i0 + 1;
            H.og[c] = dadd(A.gorg[c], (double)cf[c]);
            H.hf[c] = hf[c];
        }
        const bool cells_ok = cn[0] > 0 && cn[1] > 0 && cn[2] > 0 && (long long)cn[0] * cn[1] * cn[2] <= (long long)BX_CELLCAP;
        if (cn[0] <= 0 || cn[1] <= 0 || cn[2] <= 0) {
            H.count = 0u;  // the box misses the grid: an empty list (the queries go to the per-thread search)
        } else if (cells_ok) {
            // ---- every target point of the cells under the box that lies inside the box ----
            float4* lc = A.lcand + gid * BOX_LIST_CAP;
            uint32_t* lp = A.lpos + gid * BOX_LIST_CAP;
            unsigned int count = 0;
            for (int z = c0[2]; z < c0[2] + cn[2]; ++z)
                for (int y = c0[1]; y < c0[1] + cn[1]; ++y) {
                    const uint2* row = A.cells + ((long long)z * A.cdim[1] + y) * A.cdim[0] + c0[0];
                    for (int x = 0; x < cn[0]; ++x) {
                        const uint2 en = __ldg(row + x);
                        if (en.y - en.x >= (1u << 24)) count = BOX_LIST_CAP + 1;  // (a crowded cell: no list)
                        for (uint32_t p = en.x; p < en.y && count <= (unsigned)BOX_LIST_CAP; ++p) {
                            double px, py, pz;
                            uint32_t pidx;
                            load_point(A.pts, p, px, py, pz, pidx);
                            const float vx = (float)dsub(px, H.og[0]), vy = (float)dsub(py, H.og[1]), vz = (float)dsub(pz, H.og[2]);
                            if (fabsf(vx) <= hf[0] && fabsf(vy) <= hf[1] && fabsf(vz) <= hf[2]) {
                                if (count < (unsigned)BOX_LIST_CAP) {
                                    lc[count] = make_float4(vx, vy, vz, fmaf(vz, vz, fmaf(vy, vy, vx * vx)));
                                    lp[count] = p;
                                }
                                ++count;
                            }
                        }
                    }
                }
            H.count = count <= (unsigned)BOX_LIST_CAP ? count : BX_NOLIST;
        }
    }
    A.lhdr[gid] = H;
    if (A.counters && H.count != BX_NOLIST) atomicAdd(&A.counters[3], (unsigned long long)H.count);
    if (A.counters) atomicAdd(&A.counters[4], 1ull);
}

// One NN stage.  with_lists: the lists are valid -> list pass over every query, builder over the groups that asked, list pass
// over those groups.  Otherwise (first iteration of a run: nothing is pending, every list is built): builder over every group,
// then the list pass.  Queries left open are on the query work list afterwards.
int nn_box_launch(Ctx* c, const NNArgs& A_in, bool with_lists) {
    NNArgs A = A_in;
    if (A.n <= 0 || A.n_groups <= 0) return ICP_OK;
    static const bool dbg = getenv("ICP_B200_DEBUG_SYNC") != nullptr;
    auto check = [&](const char* what) {
        if (!dbg) return;
        cudaError_t e = cudaStreamSynchronize(c->stream);
        unsigned int wc[2] = {0, 0};
        cudaMemcpy(wc, c->d_work_count, sizeof wc, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[icp_b200] %s (%s): %u queries open, %u of %lld groups listed\n", what, cudaGetErrorString(e), wc[0], wc[1], A.n_groups);
    };
    const int q_blocks = (int)((A.n + LS_THREADS - 1) / LS_THREADS);
    const int g_blocks = (int)((A.n_groups + BD_THREADS - 1) / BD_THREADS);
    const int gq_blocks = (int)((A.n_groups * BX_G + LS_THREADS - 1) / LS_THREADS);
    uint32_t* const glist = A.group_list;
    if (with_lists) {
        A.list_final = 0;
        nn_list_kernel<false><<<q_blocks, LS_THREADS, 0, c->stream>>>(A);
        check("list pass");
        A.apply_pending = 0;  // the list pass moved every query
        nn_build_kernel<<<g_blocks, BD_THREADS, 0, c->stream>>>(A);  // (surplus threads leave at once)
        check("builder");
        A.list_final = 1;
        nn_list_kernel<true><<<gq_blocks, LS_THREADS, 0, c->stream>>>(A);
        check("list pass over the rebuilt groups");
        c->launches += 3;
    } else {
        A.group_list = nullptr;
        nn_build_kernel<<<g_blocks, BD_THREADS, 0, c->stream>>>(A);
        check("builder (all groups)");
        A.group_list = glist;
        A.list_final = 1;
        nn_list_kernel<false><<<q_blocks, LS_THREADS, 0, c->stream>>>(A);
        check("list pass");
        c->launches += 2;
    }
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
