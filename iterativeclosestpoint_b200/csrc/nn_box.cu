// Exact nearest-neighbour query, BOX SEARCH with carried candidate lists (mode 7, the default): one quarter-warp per GROUP of
// up to 8 neighbouring queries.  Same job and same answers as nn.cu (replaces Octree::findNearest / searchNearest,
// core/octree.cpp:128-184, inside the per-point loop of core/icpengine.cpp:172-184); the exactness argument is the one at the
// top of nn.cu: find the exact minimum of s (the reference's squared-distance expression) over every target point the search
// ball can contain, prove that it is unique by a margin of 2^-40, otherwise leave the query to the literal traversal.
//
// Why: in the per-query walks (nn_common.cuh: cell_walk, and the balanced form of it) every lane pulls its own 32-byte
// sectors -- ~13 L1 wavefronts per query -- and the kernel is bound by L1 lookups and issue slots at a sixth of HBM speed.
// Here the 8 queries of a group (consecutive in the internal order, all inside one cell of the cell grid when the source was
// ordered, build.cu) share ONE candidate list = every target point inside the group's BOX, recentred on the box and stored as
// FP32 (x, y, z, |p|^2).  The target never changes and the source moves by centimetres per iteration, so the list is kept in
// device memory and used again for as long as it provably still covers the group:
//
//   nn_list_kernel   (every iteration, streaming)  each lane moves its query by the pending transform (core/icpengine.cpp:345,
//       fused into the load).  The nearest point is no farther than  eb + |movement|  (eb = the distance found last time,
//       triangle inequality); if that ball lies inside the group's stored box for every query of the group, the group's list
//       holds every point that can win: the quarter-warp copies the list to shared memory with coalesced loads and every lane
//       scans it with broadcast reads, s' = |p|^2 - 2 q.p in FP32 (3 FMA per candidate), tracking the smallest, the second
//       smallest and the winner's slot.  Groups that fail the test (or carry no list) go on a group work list.
//   nn_box_kernel    (first iteration, and the groups on the work list)  builds the box = union of the queries' balls (radius
//       from last iteration's match, a real target point, plus a skin that keeps the list valid while the cloud moves),
//       enumerates the cell-grid entries under it with coalesced loads, streams the points of the non-empty cells 32 at a time
//       (a balanced search maps a trip's lanes onto the cells' ranges), keeps those inside the box, scans them the same way
//       and writes the list out for the iterations to come.
//   FP32 with FP64 recheck: |(s'_j - s'_k) - (s_j - s_k)| <= 114 * 2^-24 * W^2 (W = largest half box edge), so a gap
//       above 3 * 2^-16 * W^2 proves the winner unique in FP64 by far more than 2^-40; the winner's s and the distance handed on
//       are then evaluated in FP64 with the reference's expression.  A smaller gap re-evaluates in FP64 every candidate within
//       that margin of the best and applies the 2^-40 test to the exact values.
// Queries that stay open (no unique minimum, a ball wider than `emax`, a box over too many cells, a seedless query whose nearest
// point turns out to lie outside the box that was staged) go on the query work list for the per-thread kernel (nn.cu).
#include "nn_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace icpb {

constexpr int BX_THREADS = 128;
constexpr int BX_WARPS = BX_THREADS / 32;
constexpr int BX_G = 8;               // lanes per group
constexpr int BX_NG = 32 / BX_G;      // groups per warp
constexpr int BX_LCAP = 96;           // builder: candidates per group list in shared memory; scanned (flushed) before a trip could overflow it
constexpr int BX_CELLCAP = 64;        // builder: cell-grid entries under one group box
constexpr uint32_t BX_NOLIST = 0xFFFFFFFFu;
static_assert(BOX_LIST_CAP <= BX_LCAP && BOX_LIST_CAP % BX_G == 0, "carried lists must fit the builder's shared list");

struct __align__(16) BoxGroup {
    double og[3];       // box centre (absolute coordinates)
    float hf[3];        // half extents
    int c0[3];          // first cell per axis
    int nx, ny;         // cells along x, y
    float rnx, rny;     // 1 / nx, 1 / ny
    unsigned int count; // candidates in the list
};

// min / max over the 8 lanes of a group
__device__ __forceinline__ float group_min(float v) {
    v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return fminf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float group_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}

// one candidate against the running (best, second, slot): s' = |p|^2 - 2 q.p
#define BX_SCAN_STEP(c, k)                                                              \
    {                                                                                   \
        const float s_ = fmaf(m2x, (c).x, fmaf(m2y, (c).y, fmaf(m2z, (c).z, (c).w)));   \
        second = fminf(second, fmaxf(s_, best));                                        \
        if (s_ < best) slot = (int)(k);                                                 \
        best = fminf(best, s_);                                                         \
    }

// FP64 recheck of every candidate of a (whole) shared list within `lim` of the FP32 minimum: exact minimum and runner-up.
__device__ __forceinline__ uint32_t box_recheck(const NNArgs& A, const float4* s_cand, int g, unsigned int cnt, const uint32_t* pos_of,
                                                int pos_stride, float m2x, float m2y, float m2z, float lim, double qx, double qy,
                                                double qz, double& s_win) {
    double b64 = ICPB_INF, s64 = ICPB_INF;
    uint32_t win = NONE;
    for (unsigned int k = 0; k < cnt; ++k) {
        const float4 c = s_cand[k * BX_NG + g];
        const float s = fmaf(m2x, c.x, fmaf(m2y, c.y, fmaf(m2z, c.z, c.w)));
        if (s <= lim) {
            const uint32_t pos = pos_of[k * pos_stride];
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
    if (win != NONE && s64 > dmul(b64, 1.0 + 9.094947017729282e-13)) {
        s_win = b64;
        return win;
    }
    return NONE;
}

// ---------------------------------------------------------------------------------------------------------------------------
// the builder
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BX_THREADS) nn_box_kernel(const NNArgs A) {
    __shared__ float4 s_cand_all[BX_WARPS][BX_LCAP * BX_NG];     // slot k of group g at [k * BX_NG + g]
    __shared__ uint32_t s_cpos_all[BX_WARPS][BX_LCAP * BX_NG];   // its position in the sorted target
    __shared__ uint2 s_item_all[BX_WARPS][32];                   // non-empty cells of one enumeration trip: x = first point, y = count | group << 24
    __shared__ BoxGroup s_gb_all[BX_WARPS][BX_NG];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 3, sub = lane & 7;
    const unsigned gmask = 0xFFu << (8 * g);
    const unsigned lt_mask = (1u << lane) - 1u;
    float4* s_cand = s_cand_all[w];
    uint32_t* s_cpos = s_cpos_all[w];
    uint2* s_item = s_item_all[w];
    BoxGroup* s_gb = s_gb_all[w];
    const float INF32 = __int_as_float(0x7f800000);

    // group slots: either all groups, or the entries of the group work list that nn_list_kernel left
    const long long n_slots = A.group_list ? (long long)*A.group_count : A.n_groups;
    const long long warp_stride = (long long)gridDim.x * BX_WARPS * BX_NG;
    for (long long slot0 = ((long long)blockIdx.x * BX_WARPS + w) * BX_NG; slot0 < n_slots; slot0 += warp_stride) {
    const long long gslot = slot0 + g;
    long long gid = -1;
    if (gslot < n_slots) gid = A.group_list ? (long long)A.group_list[gslot] : gslot;
    uint32_t qb = 0, qn_count = 0;
    if (gid >= 0) {
        qb = __ldg(A.gstart + gid);
        qn_count = __ldg(A.gstart + gid + 1) - qb;
    }
    const bool active = (uint32_t)sub < qn_count;
    const long long i = (long long)qb + sub;

    // ---- A. own query ----
    double qx = 0.0, qy = 0.0, qz = 0.0, e = 0.0;
    bool elig = false;
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
        if (finite_q && pp != NONE) {
            double px, py, pz;
            uint32_t pidx;
            load_point(A.pts, pp, px, py, pz, pidx);
            Sd = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
        }
        // ball radius: >= the distance to a real target point when there is one (then the box certainly holds the nearest
        // point) plus the skin that keeps the list usable while the query moves, else a guess that step F checks against
        // what was found
        e = (Sd < 1e19) ? dadd(dadd(sqrt_upper(Sd), dmul(A.geps, 2.0)), A.box_skin) : A.box_guess;
        elig = finite_q && e <= A.box_emax;
    }
    // coordinates relative to the grid origin; queries far outside the target's cube are left to the per-thread search
    const double ax = dsub(qx, A.gorg[0]), ay = dsub(qy, A.gorg[1]), az = dsub(qz, A.gorg[2]);
    {
        const double far = A.gcube;
        if (!(ax > -far && ax < far + far && ay > -far && ay < far + far && az > -far && az < far + far)) elig = false;
    }

    // ---- B. the group's box ----
    float flo[3], fhi[3];
    {
        const double a3[3] = {ax, ay, az};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            flo[c] = group_min(elig ? __double2float_rd(dsub(a3[c], e)) : INF32);
            fhi[c] = group_max(elig ? __double2float_ru(dadd(a3[c], e)) : -INF32);
        }
    }
    const bool group_any = (__ballot_sync(FULL, elig) & gmask) != 0u;
    float cf[3], hf[3];
    int c0[3], cn[3];
    bool group_ok = group_any;
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
        cn[c] = group_any ? i1 - i0 + 1 : 0;
    }
    unsigned int ncell = (cn[0] > 0 && cn[1] > 0 && cn[2] > 0) ? (unsigned)cn[0] * (unsigned)cn[1] * (unsigned)cn[2] : 0u;
    if (ncell > (unsigned)BX_CELLCAP) {
        group_ok = false;  // a jump of the ordering curve inside the group, or a huge ball: the per-thread search takes these queries
        ncell = 0u;
    }
    const double ogx = dadd(A.gorg[0], (double)cf[0]), ogy = dadd(A.gorg[1], (double)cf[1]), ogz = dadd(A.gorg[2], (double)cf[2]);
    if (sub == 0) {
        BoxGroup B;
        B.og[0] = ogx; B.og[1] = ogy; B.og[2] = ogz;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            B.hf[c] = hf[c];
            B.c0[c] = c0[c];
        }
        B.nx = max(cn[0], 1);
        B.ny = max(cn[1], 1);
        B.rnx = 1.0f / (float)B.nx;
        B.rny = 1.0f / (float)B.ny;
        B.count = 0u;
        s_gb[g] = B;
    }
    // own query relative to the box centre, FP32
    const float qfx = (float)dsub(qx, ogx), qfy = (float)dsub(qy, ogy), qfz = (float)dsub(qz, ogz);
    const float m2x = -2.0f * qfx, m2y = -2.0f * qfy, m2z = -2.0f * qfz;
    const float hmax = fmaxf(fmaxf(hf[0], hf[1]), hf[2]);
    const float margin = 4.57763671875e-05f * hmax * hmax;  // 3 * 2^-16 * W^2
    __syncwarp();

    // running result of the scans
    float best = INF32, second = INF32;
    uint32_t bpos = NONE;
    int flushes = 0;
    auto scan_list = [&]() {
        const unsigned int cnt = s_gb[g].count;
        int slot = -1;
#pragma unroll 4
        for (unsigned int k = 0; k < cnt; ++k) {
            const float4 c = s_cand[k * BX_NG + g];
            BX_SCAN_STEP(c, k);
        }
        if (slot >= 0) bpos = s_cpos[slot * BX_NG + g];
    };

    // ---- C. + D. cells under the four boxes -> items -> points -> lists ----
    const unsigned int n0 = __shfl_sync(FULL, ncell, 0), n1 = __shfl_sync(FULL, ncell, 8), n2 = __shfl_sync(FULL, ncell, 16),
                       n3 = __shfl_sync(FULL, ncell, 24);
    const unsigned int pre1 = n0, pre2 = n0 + n1, pre3 = pre2 + n2, T = pre3 + n3;
    bool crowded = false;  // a cell of this lane's enumeration holds too many points for an item
    for (unsigned int f0 = 0; f0 < T; f0 += 32) {
        // one cell per lane
        const unsigned int f = f0 + lane;
        uint2 en = make_uint2(0u, 0u);
        unsigned int cg = 0;
        if (f < T) {
            cg = (f >= pre1 ? 1u : 0u) + (f >= pre2 ? 1u : 0u) + (f >= pre3 ? 1u : 0u);
            const unsigned int l = f - (cg == 0 ? 0u : (cg == 1 ? pre1 : (cg == 2 ? pre2 : pre3)));
            const BoxGroup& B = s_gb[cg];
            const int t = (int)(((float)l + 0.5f) * B.rnx);
            const int dx = (int)l - t * B.nx;
            const int dz = (int)(((float)t + 0.5f) * B.rny);
            const int dy = t - dz * B.ny;
            en = __ldg(A.cells + ((long long)(B.c0[2] + dz) * A.cdim[1] + (B.c0[1] + dy)) * A.cdim[0] + (B.c0[0] + dx));
        }
        unsigned int cnt = en.y - en.x;
        if (cnt >= (1u << 24)) {
            crowded = true;
            cnt = 0u;
        }
        const unsigned nonempty = __ballot_sync(FULL, cnt > 0u);
        const int n_items = __popc(nonempty);
        if (cnt > 0u) s_item[__popc(nonempty & lt_mask)] = make_uint2(en.x, cnt | (cg << 24));
        __syncwarp();
        // item per lane, inclusive scan of the point counts
        uint2 it = (lane < n_items) ? s_item[lane] : make_uint2(0u, 0u);
        const unsigned int icnt = it.y & 0xFFFFFFu;
        unsigned int iend = icnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int v = __shfl_up_sync(FULL, iend, o);
            if (lane >= o) iend += v;
        }
        const unsigned int istart = iend - icnt;
        const unsigned int P = __shfl_sync(FULL, iend, 31);
        int base_item = 0;  // items that started before this trip's window
        for (unsigned int t0 = 0; t0 < P; t0 += 32) {
            // which item does point t0 + lane belong to: items starting inside the window mark their start lane
            const unsigned int rel = istart - t0;
            const unsigned int word = __reduce_or_sync(FULL, (icnt > 0u && rel < 32u) ? (1u << rel) : 0u);
            const unsigned int pf = t0 + lane;
            const bool valid = pf < P;
            int k = base_item + __popc(word & (lt_mask | (1u << lane))) - 1;
            base_item += __popc(word);
            k = valid ? k : 0;
            const unsigned int k_start = __shfl_sync(FULL, istart, k);
            const unsigned int k_pt0 = __shfl_sync(FULL, it.x, k);
            const unsigned int k_g = __shfl_sync(FULL, it.y, k) >> 24;
            const uint32_t pos = k_pt0 + (pf - k_start);
            bool inside = false;
            float vx = 0.f, vy = 0.f, vz = 0.f;
            if (valid) {
                double px, py, pz;
                uint32_t pidx;
                load_point(A.pts, pos, px, py, pz, pidx);
                const BoxGroup& B = s_gb[k_g];
                vx = (float)dsub(px, B.og[0]);
                vy = (float)dsub(py, B.og[1]);
                vz = (float)dsub(pz, B.og[2]);
                inside = fabsf(vx) <= B.hf[0] && fabsf(vy) <= B.hf[1] && fabsf(vz) <= B.hf[2];
            }
            // append to the owning group's list: rank among the trip's lanes of the same group
            const unsigned same = __match_any_sync(FULL, valid ? k_g : 0xFFu);
            const unsigned ins = __ballot_sync(FULL, inside) & same;
            const unsigned int base = valid ? s_gb[k_g].count : 0u;
            __syncwarp();
            if (inside) {
                const unsigned int slot = base + __popc(ins & lt_mask);
                s_cand[slot * BX_NG + k_g] = make_float4(vx, vy, vz, fmaf(vz, vz, fmaf(vy, vy, vx * vx)));
                s_cpos[slot * BX_NG + k_g] = pos;
                if ((ins & lt_mask) == 0u) s_gb[k_g].count = base + __popc(ins);
            }
            __syncwarp();
            // a list that could not take another full trip is scanned now and emptied
            if (__any_sync(FULL, s_gb[g].count > (unsigned)(BX_LCAP - 32))) {
                scan_list();
                ++flushes;
                __syncwarp();
                if (sub == 0) s_gb[g].count = 0u;
                __syncwarp();
            }
        }
        __syncwarp();
    }
    scan_list();
    crowded = (__ballot_sync(FULL, crowded) != 0u);  // (rare: any crowded cell under the warp's boxes sends the warp's groups on)

    // ---- E. the list, for the iterations to come ----
    if (A.lhdr && gid >= 0) {
        const unsigned int cnt = s_gb[g].count;
        const bool keep = group_ok && !crowded && flushes == 0 && cnt <= (unsigned)BOX_LIST_CAP;
        if (keep) {
            for (unsigned int k = sub; k < cnt; k += BX_G) {
                A.lcand[gid * BOX_LIST_CAP + k] = s_cand[k * BX_NG + g];
                A.lpos[gid * BOX_LIST_CAP + k] = s_cpos[k * BX_NG + g];
            }
        }
        if (sub == 0) {
            BoxListHdr H;
            H.og[0] = ogx; H.og[1] = ogy; H.og[2] = ogz;
            H.hf[0] = hf[0]; H.hf[1] = hf[1]; H.hf[2] = hf[2];
            H.count = keep ? cnt : BX_NOLIST;
            H.pad[0] = H.pad[1] = 0u;
            A.lhdr[gid] = H;
        }
    }

    // ---- F. decide ----
    bool settled = false;
    if (elig && group_ok && !crowded && bpos != NONE) {
        uint32_t win = NONE;
        double s_win = 0.0;
        if (second - best > margin) {
            win = bpos;
            double px, py, pz;
            uint32_t pidx;
            load_point(A.pts, win, px, py, pz, pidx);
            s_win = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
        } else if (flushes == 0) {
            win = box_recheck(A, s_cand, g, s_gb[g].count, s_cpos + g, BX_NG, m2x, m2y, m2z, best + margin, qx, qy, qz, s_win);
        }
        if (win != NONE) {
            // the staged box holds every target point within `clr` of the query along each axis; the answer stands if the
            // ball of the distance found lies inside (always true for a seeded query, checked for all)
            // (sqrt_upper >= sqrt(s) (1 + 2^-18): a point that ties with the winner within 2^-40 is inside too)
            const double rho = dadd(sqrt_upper(s_win), A.geps);
            const double clr = fmin(fmin((double)hf[0] - fabs(dsub(qx, ogx)), (double)hf[1] - fabs(dsub(qy, ogy))),
                                    (double)hf[2] - fabs(dsub(qz, ogz)));
            if (rho <= clr) {
                settled = true;
                A.pos_out[i] = win;
                const double d = dsqrt(s_win);  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
                A.dist_out[i] = d;
                if (A.ebound) A.ebound[i] = __double2float_ru(d);
            }
        }
    }
    const unsigned pend = __ballot_sync(FULL, active && !settled);
    if (pend) {
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(A.work_count, (unsigned int)__popc(pend));
        at = __shfl_sync(FULL, at, 0);
        if (active && !settled) A.worklist[at + __popc(pend & lt_mask)] = (uint32_t)i;
    }
    if (A.counters) {  // profiling / tests only
        const unsigned ok = __ballot_sync(FULL, settled);
        const unsigned int cand = __reduce_add_sync(FULL, (sub == 0 && flushes == 0) ? s_gb[g].count : 0u);
        if (lane == 0) {
            if (ok) atomicAdd(&A.counters[0], (unsigned long long)__popc(ok));
            if (pend) atomicAdd(&A.counters[2], (unsigned long long)__popc(pend));
            atomicAdd(&A.counters[3], (unsigned long long)cand);
            atomicAdd(&A.counters[4], (unsigned long long)T);
        }
    }
    __syncwarp();
    }  // group slots
}

// ---------------------------------------------------------------------------------------------------------------------------
// the streaming pass over the carried lists
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BX_THREADS) nn_list_kernel(const NNArgs A) {
    __shared__ float4 s_cand_all[BX_WARPS][BOX_LIST_CAP * BX_NG];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 3, sub = lane & 7;
    const unsigned gmask = 0xFFu << (8 * g);
    const unsigned lt_mask = (1u << lane) - 1u;
    float4* s_cand = s_cand_all[w];
    const float INF32 = __int_as_float(0x7f800000);

    const long long gid = ((long long)blockIdx.x * BX_WARPS + w) * BX_NG + g;
    const bool have_group = gid < A.n_groups;
    uint32_t qb = 0, qn_count = 0;
    BoxListHdr H;
    H.count = BX_NOLIST;
    H.og[0] = H.og[1] = H.og[2] = 0.0;
    H.hf[0] = H.hf[1] = H.hf[2] = 0.f;
    if (have_group) {
        qb = __ldg(A.gstart + gid);
        qn_count = __ldg(A.gstart + gid + 1) - qb;
        const int4* hp = reinterpret_cast<const int4*>(A.lhdr + gid);
        const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
        H.og[0] = __hiloint2double(h0.y, h0.x);
        H.og[1] = __hiloint2double(h0.w, h0.z);
        H.og[2] = __hiloint2double(h1.y, h1.x);
        H.hf[0] = __int_as_float(h1.z);
        H.hf[1] = __int_as_float(h1.w);
        H.hf[2] = __int_as_float(h2.x);
        H.count = (uint32_t)h2.y;
    }
    const bool active = (uint32_t)sub < qn_count;
    const long long i = (long long)qb + sub;

    // ---- own query: move it, bound its nearest-neighbour distance, test the ball against the stored box ----
    double qx = 0.0, qy = 0.0, qz = 0.0;
    float vx = 0.f, vy = 0.f, vz = 0.f, slack = INF32;
    bool ok = true;
    if (active) {
        qx = A.sx[i];
        qy = A.sy[i];
        qz = A.sz[i];
        float move = 0.f;
        if (A.apply_pending && A.state->have_T) {
            const double x0 = qx, y0 = qy, z0 = qz;
            apply_T_point(A.state->T_pending, qx, qy, qz);
            A.ox[i] = qx;
            A.oy[i] = qy;
            A.oz[i] = qz;
            const float mx = (float)dsub(qx, x0), my = (float)dsub(qy, y0), mz = (float)dsub(qz, z0);
            move = __fmul_ru(__fsqrt_ru(fmaf(mz, mz, fmaf(my, my, mx * mx))), 1.000002f);
        }
        // the point matched last time is a real target point at distance <= eb from where the query was
        const float e_pred = __fadd_ru(__fadd_ru(A.ebound[i], move), A.geps2_f);
        vx = (float)dsub(qx, H.og[0]);
        vy = (float)dsub(qy, H.og[1]);
        vz = (float)dsub(qz, H.og[2]);
        const float sx_ = __fsub_rd(H.hf[0], __fadd_ru(__fmul_ru(fabsf(vx), 1.0000002f), e_pred));
        const float sy_ = __fsub_rd(H.hf[1], __fadd_ru(__fmul_ru(fabsf(vy), 1.0000002f), e_pred));
        const float sz_ = __fsub_rd(H.hf[2], __fadd_ru(__fmul_ru(fabsf(vz), 1.0000002f), e_pred));
        slack = fminf(sx_, fminf(sy_, sz_));
        ok = slack >= 0.0f;  // false for NaN (non-finite query or bound)
    }
    // (the ballot on its own line: inside a short-circuited && the lanes whose left operand is false would skip it and the
    // others would wait for them for ever)
    const unsigned not_ok = __ballot_sync(FULL, !ok);
    bool use_list = have_group && H.count <= (uint32_t)BOX_LIST_CAP && (not_ok & gmask) == 0u;
    if (A.box_tighten < 0.f) use_list = false;  // (bring-up switch: every group goes to the builder)
    // a list far wider than the balls need (the registration has closed in since it was built) is rebuilt tighter
    if (A.box_tighten > 0.f) {
        const float gslack = group_min(slack);
        if (gslack > A.box_tighten && H.count > 12u) use_list = false;
    }
    {
        const unsigned defer = __ballot_sync(FULL, have_group && sub == 0 && !use_list);
        if (defer) {
            unsigned int at = 0;
            if (lane == 0) at = atomicAdd(A.group_count, (unsigned int)__popc(defer));
            at = __shfl_sync(FULL, at, 0);
            if (have_group && sub == 0 && !use_list) A.group_list[at + __popc(defer & lt_mask)] = (uint32_t)gid;
        }
    }
    const unsigned int cnt = use_list ? H.count : 0u;
    for (unsigned int k = sub; k < cnt; k += BX_G) s_cand[k * BX_NG + g] = __ldg(A.lcand + gid * BOX_LIST_CAP + k);
    __syncwarp();

    // ---- scan ----
    const float m2x = -2.0f * vx, m2y = -2.0f * vy, m2z = -2.0f * vz;
    float best = INF32, second = INF32;
    int slot = -1;
#pragma unroll 4
    for (unsigned int k = 0; k < cnt; ++k) {
        const float4 c = s_cand[k * BX_NG + g];
        BX_SCAN_STEP(c, k);
    }

    // ---- decide ----
    bool settled = false;
    if (active && use_list && slot >= 0) {
        const float hmax = fmaxf(fmaxf(H.hf[0], H.hf[1]), H.hf[2]);
        const float margin = 4.57763671875e-05f * hmax * hmax;  // 3 * 2^-16 * W^2
        uint32_t win = NONE;
        double s_win = 0.0;
        const uint32_t* pos_of = A.lpos + gid * BOX_LIST_CAP;
        if (second - best > margin) {
            win = __ldg(pos_of + slot);
            double px, py, pz;
            uint32_t pidx;
            load_point(A.pts, win, px, py, pz, pidx);
            s_win = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
        } else {
            win = box_recheck(A, s_cand, g, cnt, pos_of, 1, m2x, m2y, m2z, best + margin, qx, qy, qz, s_win);
        }
        if (win != NONE) {
            settled = true;
            A.pos_out[i] = win;
            const double d = dsqrt(s_win);  // computeDistance (icpengine.cpp:68-74): sqrt of the same sum of squares
            A.dist_out[i] = d;
            A.ebound[i] = __double2float_ru(d);
        }
    }
    // open queries of a group that used its list (no unique minimum) go to the per-thread kernel; the queries of a deferred
    // group are the builder's
    const unsigned pend = __ballot_sync(FULL, active && use_list && !settled);
    if (pend) {
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(A.work_count, (unsigned int)__popc(pend));
        at = __shfl_sync(FULL, at, 0);
        if (active && use_list && !settled) A.worklist[at + __popc(pend & lt_mask)] = (uint32_t)i;
    }
    if (A.counters) {  // profiling / tests only
        const unsigned okb = __ballot_sync(FULL, settled);
        const unsigned int cand = __reduce_add_sync(FULL, sub == 0 ? cnt : 0u);
        if (lane == 0) {
            if (okb) atomicAdd(&A.counters[0], (unsigned long long)__popc(okb));
            if (okb) atomicAdd(&A.counters[5], (unsigned long long)__popc(okb));
            if (pend) atomicAdd(&A.counters[2], (unsigned long long)__popc(pend));
            atomicAdd(&A.counters[3], (unsigned long long)cand);
        }
    }
}

// `with_lists`: run nn_list_kernel over every group first, then the builder over the groups it deferred; otherwise the builder
// over every group (first iteration, or no list storage).
int nn_box_launch(Ctx* c, const NNArgs& A_in, bool with_lists) {
    NNArgs A = A_in;
    const long long warps = (A.n_groups + BX_NG - 1) / BX_NG;
    const int blocks_all = (int)((warps + BX_WARPS - 1) / BX_WARPS);
    if (blocks_all <= 0) return ICP_OK;
    static const bool dbg = getenv("ICP_B200_DEBUG_SYNC") != nullptr;
    if (dbg) fprintf(stderr, "[icp_b200] nn_box_launch: %lld groups, %d blocks, with_lists=%d lhdr=%p ebound=%p\n", A.n_groups, blocks_all, (int)with_lists, (void*)A.lhdr, (void*)A.ebound);
    if (with_lists) {
        nn_list_kernel<<<blocks_all, BX_THREADS, 0, c->stream>>>(A);
        c->launches++;
        if (dbg) {
            cudaError_t e = cudaStreamSynchronize(c->stream);
            unsigned int wc[2] = {0, 0};
            cudaMemcpy(wc, c->d_work_count, sizeof wc, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[icp_b200] list kernel done (%s): %u queries open, %u of %lld groups deferred\n", cudaGetErrorString(e), wc[0], wc[1], A.n_groups);
        }
        A.apply_pending = 0;  // the list kernel moved every query, also those of the groups it deferred
        nn_box_kernel<<<std::min(blocks_all, c->sm_count * 16), BX_THREADS, 0, c->stream>>>(A);
        if (dbg) {
            cudaError_t e = cudaStreamSynchronize(c->stream);
            fprintf(stderr, "[icp_b200] builder over the deferred groups done (%s)\n", cudaGetErrorString(e));
        }
    } else {
        A.group_list = nullptr;
        nn_box_kernel<<<std::min(blocks_all, c->sm_count * 32), BX_THREADS, 0, c->stream>>>(A);
    }
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
