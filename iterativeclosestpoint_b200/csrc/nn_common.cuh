// Device-side pieces shared by the NN kernels (nn.cu: one query per thread; nn_tile.cu: one tile of queries per warp).
#pragma once
#include "internal.h"

namespace icpb {


constexpr int NN_MAX_LEVELS = 22;           // octree_max_depth <= 21  => at most 22 levels on a path
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int SEED_SCAN = 8;                // points inspected to seed a query
#define ICPB_INF __longlong_as_double(0x7FF0000000000000LL)

struct NNArgs {
    const Node* __restrict__ nodes;    // search tree (isotropic cells): every fast path; match positions index `pts`
    const TPoint* __restrict__ pts;
    const Node* __restrict__ rnodes;   // the reference's octree and its point order: literal traversal only
    const TPoint* __restrict__ rpts;
    const uint32_t* __restrict__ inv;  // original target index -> position in `pts`
    // entry grid of the search tree (build.cu): cell -> owning node, NONE where empty
    const uint2* __restrict__ grid;    // pyramid: levels glmin .. glmin + gnlev - 1, level k at grid + goff[k]
    int glmin, gnlev, gmax_cells, gbias;
    int gbase, gkmin;                  // pyramid index of the base level (seeds); coarsest index the walk may use
    long long goff[4];
    int gdim[4][3];
    double ginv[4];                    // 1 / cell edge per level
    double gorg[3], geps, gcube;
    const double* sx;
    const double* sy;
    const double* sz;
    double* ox;
    double* oy;
    double* oz;
    long long n;
    uint32_t* pos_out;
    double* dist_out;
    const uint32_t* prev_pos;  // last iteration's match per query (may be null)
    uint32_t* node_io;         // in: leaf that held last iteration's match; out: this iteration's (may be null)
    const uint32_t* __restrict__ parent;
    StatA* part_a;
    const LoopState* state;
    unsigned long long* counters;  // [0] fast-path answers, [1] literal fallbacks, [2] tile lanes sent to the per-thread search,
                                   // [3] candidates scanned by tiles (may be null)
    int apply_pending;
    int mode;                  // 0: literal traversal from the root; 1: per-thread, climb from the last leaf; 2: warp tiles;
                               // 3: per-thread, entry through the grid cells the search ball touches;
                               // 4: as 3, the candidate scan balanced over the warp (nn_group.cu)
    float* lb_io;              // mode 4: per query, a lower bound on the distance to every target point other than its match
                               // (rounded down; 0 = unknown) -- lets a later iteration keep the match without a search
    double gedge[4];           // cell edge per pyramid level
    double gbias_mul;          // 2^-gbias
    uint4* cand_io;            // mode 5 (nn_keep.cu): the K nearest target points of the query's last search
    uint32_t* worklist2;       // mode 5: queries handed on to the per-thread kernel (count at work_count[1])
    double walk_alpha, walk_wmul, walk_rcap;  // mode 5: search ball = seed radius x alpha, widened up to rcap at most; level = finest
                               // with cell edge >= ball radius x wmul
    long long resident_threads; // threads of nn_kernel the GPU holds at once (work-list spreading)
    uint32_t* worklist;        // mode 4: queries the balanced kernel hands to the per-thread kernel ...
    unsigned int* work_count;  // ... and how many; the per-thread kernel runs over that list when worklist != null
    double init_best;
    uint32_t pos_of_idx0;
};

struct NodeRegs {
    double lo[3], hi[3];
    uint32_t child0, pt0, npts, meta;
};

__device__ __forceinline__ NodeRegs load_node(const Node* __restrict__ nodes, uint32_t i) {
    const int4* p = reinterpret_cast<const int4*>(nodes + i);
    int4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    NodeRegs r;
    r.lo[0] = __hiloint2double(a.y, a.x);
    r.lo[1] = __hiloint2double(a.w, a.z);
    r.lo[2] = __hiloint2double(b.y, b.x);
    r.hi[0] = __hiloint2double(b.w, b.z);
    r.hi[1] = __hiloint2double(c.y, c.x);
    r.hi[2] = __hiloint2double(c.w, c.z);
    r.child0 = (uint32_t)d.x;
    r.pt0 = (uint32_t)d.y;
    r.npts = (uint32_t)d.z;
    r.meta = (uint32_t)d.w;
    return r;
}

// One 256-bit load per target point (sm_100a: LDG.E.256): a point is one 32-byte sector, so one request to L1 instead of two.
__device__ __forceinline__ void load_point(const TPoint* __restrict__ pts, uint32_t i, double& x, double& y, double& z,
                                           uint32_t& idx) {
    long long w;
    asm("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=l"(w) : "l"(pts + i));
    idx = (uint32_t)w;
}

// OctreeNode::minDistanceTo's per-axis term: max(0, max(lo - q, q - hi))   (octree.cpp:34-36)
__device__ __forceinline__ double axis_dist(double lo, double hi, double q) {
    return stdmax(0.0, stdmax(dsub(lo, q), dsub(q, hi)));
}

struct Search {
    double best;       // best squared distance so far
    uint32_t pos;      // its position in the sorted target, NONE while nothing accepted
    uint32_t idx;      // its original index
};

// ---------------------------------------------------------------------------------------------------
// The reference traversal from the root (whose own prune test the caller has already made).
// stk: this thread's column of the shared stack (row stride STRIDE).
// ---------------------------------------------------------------------------------------------------
template <int STRIDE>
__device__ __forceinline__ void dfs_literal(const Node* __restrict__ nodes, const TPoint* __restrict__ pts, const double qx,
                                            const double qy, const double qz, Search& S, uint2* stk) {
    int sp = 0;
    uint32_t cur = 0;
    bool have_cur = true;
    for (;;) {
        if (have_cur) {
            const NodeRegs nd = load_node(nodes, cur);
            const uint32_t mask = nd.meta & 0xFFu;
            if (mask == 0) {
                // leaf: octree.cpp:139-150.  The reference scans ascending original index with a strict <,
                // i.e. within one leaf the smallest distance wins and equal distances go to the lowest index;
                // points here are in key order, so that rule is applied explicitly.
                bool from_this_leaf = false;
                for (uint32_t k = 0; k < nd.npts; ++k) {
                    double px, py, pz;
                    uint32_t pidx;
                    load_point(pts, nd.pt0 + k, px, py, pz, pidx);
                    const double d2 = sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz));
                    if (d2 < S.best || (from_this_leaf && d2 == S.best && pidx < S.idx)) {
                        S.best = d2;
                        S.pos = nd.pt0 + k;
                        S.idx = pidx;
                        from_this_leaf = true;
                    }
                }
                have_cur = false;
            } else {
                // inner node: octree.cpp:152-171
                double dl[3], dh[3];
                const double q[3] = {qx, qy, qz};
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const double mid = dmul(dadd(nd.lo[a], nd.hi[a]), 0.5);
                    const double l = axis_dist(nd.lo[a], mid, q[a]);
                    const double h = axis_dist(mid, nd.hi[a], q[a]);
                    dl[a] = dmul(l, l);
                    dh[a] = dmul(h, h);
                }
                double md[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    const double s = dadd(dadd((o & 1) ? dh[0] : dl[0], (o & 2) ? dh[1] : dl[1]), (o & 4) ? dh[2] : dl[2]);
                    md[o] = ((mask >> o) & 1u) ? dsqrt(s) : ICPB_INF;
                }
                // rank = position after a stable ascending sort by md (std::sort on <= 8 items == insertion sort)
                uint32_t word = 0;
                double mdmin = md[0];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t r = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < i) r += (md[j] <= md[i]) ? 1u : 0u;
                        if (j > i) r += (md[j] < md[i]) ? 1u : 0u;
                    }
                    word |= (uint32_t)i << (3 * r);
                    if (i > 0) mdmin = fmin(mdmin, md[i]);
                }
                const uint32_t cnt = __popc(mask);
                const double m0 = dmul(mdmin, mdmin);
                if (m0 >= S.best) {
                    have_cur = false;  // nearest child pruned => all children pruned (octree.cpp:134-135)
                } else {
                    const uint32_t oct = word & 7u;
                    if (cnt > 1) {
                        // entry: x = node, y = order word (24 bits) | next position (4 bits) | count (4 bits)
                        stk[sp * STRIDE] = make_uint2(cur, (word & 0xFFFFFFu) | (1u << 24) | (cnt << 28));
                        ++sp;
                    }
                    cur = nd.child0 + __popc(mask & ((1u << oct) - 1u));
                }
            }
        } else {
            if (sp == 0) break;
            uint2 top = stk[(sp - 1) * STRIDE];
            const uint32_t k = (top.y >> 24) & 0xFu, cnt = top.y >> 28;
            const uint32_t oct = (top.y >> (3 * k)) & 7u;
            const NodeRegs pn = load_node(nodes, top.x);
            const double q[3] = {qx, qy, qz};
            double d[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double mid = dmul(dadd(pn.lo[a], pn.hi[a]), 0.5);
                const bool up = (oct >> a) & 1u;
                d[a] = axis_dist(up ? mid : pn.lo[a], up ? pn.hi[a] : mid, q[a]);
            }
            const double mdc = dsqrt(sumsq3(d[0], d[1], d[2]));
            const double m = dmul(mdc, mdc);
            if (m >= S.best) {
                --sp;  // this sibling and every later one (larger distance) are pruned
                continue;
            }
            if (k + 1 >= cnt) {
                --sp;
            } else {
                top.y = (top.y & ~(0xFu << 24)) | ((k + 1) << 24);
                stk[(sp - 1) * STRIDE] = top;
            }
            const uint32_t mask = pn.meta & 0xFFu;
            cur = pn.child0 + __popc(mask & ((1u << oct) - 1u));
            have_cur = true;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Fast path: exact minimum of s and the runner-up, un-rooted bounds, any visiting order.
// Stack entries: x = child0 of the parent, y = parent's child mask | (octants still to visit << 8).
// ---------------------------------------------------------------------------------------------------
struct Fast {
    double best, second, bound;
    uint32_t pos, node;
};

__device__ __forceinline__ void fast_take(Fast& F, double s, uint32_t pos, uint32_t node) {
    const double grow = 1.0 + 1.8189894035458565e-12;  // 1 + 2^-39
    if (s < F.best) {
        F.second = F.best;
        F.best = s;
        F.pos = pos;
        F.node = node;
        const double b = dmul(s, grow);
        F.bound = b < F.bound ? b : F.bound;
    } else if (s < F.second) {
        F.second = s;
    }
}

// Scans the points [pt0, pt0 + npts) of one leaf, four loads in flight at a time.  Points whose s exceeds the running
// bound (<= min(seed, best) (1 + 2^-39)) can be neither the minimum nor a near-tie of it and are skipped.
__device__ __forceinline__ void scan_leaf_points(const TPoint* __restrict__ pts, const uint32_t pt0, const uint32_t npts,
                                                 const double qx, const double qy, const double qz, Fast& F,
                                                 const uint32_t tag) {
    for (uint32_t k = 0; k < npts; k += 4) {
        double px[4], py[4], pz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t kk = (k + j < npts) ? k + j : npts - 1u;
            uint32_t pidx;
            load_point(pts, pt0 + kk, px[j], py[j], pz[j], pidx);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double s = sumsq3(dsub(px[j], qx), dsub(py[j], qy), dsub(pz[j], qz));
            if (k + j < npts && s <= F.bound) fast_take(F, s, pt0 + k + j, tag);
        }
    }
}

template <int STRIDE>
__device__ __forceinline__ void fast_search(const Node* __restrict__ nodes, const TPoint* __restrict__ pts, const double qx,
                                            const double qy, const double qz, uint32_t start, Fast& F, uint2* stk) {
    uint2* top = stk;  // one past the last entry of this thread's column
    uint32_t cur = start;
    for (;;) {
        const NodeRegs nd = load_node(nodes, cur);
        const uint32_t mask = nd.meta & 0xFFu;
        bool descend = false;
        if (mask == 0) {
            const double sb = sumsq3(axis_dist(nd.lo[0], nd.hi[0], qx), axis_dist(nd.lo[1], nd.hi[1], qy),
                                     axis_dist(nd.lo[2], nd.hi[2], qz));
            if (sb <= F.bound) scan_leaf_points(pts, nd.pt0, nd.npts, qx, qy, qz, F, cur);
        } else {
            double dl[3], dh[3];
            const double q[3] = {qx, qy, qz};
            uint32_t oq = 0;  // octant of the query inside this node (visited first: it tightens the bound)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double mid = dmul(dadd(nd.lo[a], nd.hi[a]), 0.5);
                const double l = axis_dist(nd.lo[a], mid, q[a]);
                const double h = axis_dist(mid, nd.hi[a], q[a]);
                dl[a] = dmul(l, l);
                dh[a] = dmul(h, h);
                oq |= (q[a] > mid ? 1u : 0u) << a;
            }
            uint32_t surv = 0;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const double s = dadd(dadd((o & 1) ? dh[0] : dl[0], (o & 2) ? dh[1] : dl[1]), (o & 4) ? dh[2] : dl[2]);
                surv |= (s <= F.bound ? 1u : 0u) << o;
            }
            surv &= mask;
            if (surv) {
                const uint32_t o = ((surv >> oq) & 1u) ? oq : (uint32_t)(__ffs(surv) - 1);
                surv &= ~(1u << o);
                if (surv) {
                    *top = make_uint2(nd.child0, mask | (surv << 8));
                    top += STRIDE;
                }
                cur = nd.child0 + __popc(mask & ((1u << o) - 1u));
                descend = true;
            }
        }
        if (descend) continue;
        // next pending sibling
        if (top == stk) break;
        uint2 e = *(top - STRIDE);
        uint32_t surv = e.y >> 8;
        const uint32_t o = (uint32_t)(__ffs(surv) - 1);
        surv &= surv - 1u;
        if (surv) {
            e.y = (e.y & 0xFFu) | (surv << 8);
            *(top - STRIDE) = e;
        } else {
            top -= STRIDE;
        }
        cur = e.x + __popc((e.y & 0xFFu) & ((1u << o) - 1u));
    }
}


// ---------------------------------------------------------------------------------------------------
// Cell walk (mode 3): the search ball around q (radius = distance to a known target point) touches a handful of
// entry-grid cells; the exact minimum over the subtrees that own those cells is the exact minimum over the cloud,
// because every point of any other cell differs from q by more than the radius along some axis.
// Grid entries (uint2, build.cu): y >> 30 = kind: 0 empty; 1 leaf at the grid level (x = first point, y & 0xFFFFFF =
// count); 2 leaf above the grid level owning a block of cells (same, depth in bits 24-29); 3 inner node (x = node).
// Returns false if the walk does not apply (no seed, or too many cells) -- the caller then uses the climbing search.
// ---------------------------------------------------------------------------------------------------
struct GridView {
    const uint2* g;
    int nx, ny, nz, level;
    double inv;
};

__device__ __forceinline__ GridView grid_view(const NNArgs& A, int k) {
    GridView V;
    V.g = A.grid + A.goff[k];
    V.nx = A.gdim[k][0];
    V.ny = A.gdim[k][1];
    V.nz = A.gdim[k][2];
    V.level = A.glmin + k;
    V.inv = A.ginv[k];
    return V;
}

// finest pyramid level whose cell edge is at least w (so a box of width w meets at most 2 cells per axis), + bias
// (a heuristic only: the callers check the number of cells the box really meets)
__device__ __forceinline__ int grid_level_for_width(const NNArgs& A, double w, int bias) {
    const double ws = w * A.gbias_mul;  // w / 2^bias
    int k = A.gnlev - 1;
#pragma unroll
    for (int t = 0; t < 3; ++t)
        if (k > A.gkmin && A.gedge[k] < ws) --k;
    return k;
}

__device__ __forceinline__ int grid_cell_index(const NNArgs& A, const GridView& V, double v, int a, int n) {
    double f = floor(dmul(dsub(v, A.gorg[a]), V.inv));
    f = fmin(fmax(f, -1.0), (double)n);
    return (int)f;
}

// cells [lo, hi] that the interval [v - e, v + e] meets along axis a, clamped to [-1, n] like grid_cell_index; one
// subtraction and two products per axis (the roundings, ~2^-40 of a cell, are far inside the 2^-20 cell that e carries)
__device__ __forceinline__ void grid_cell_span(const NNArgs& A, const GridView& V, double v, double e_cells, int a, int n, int& lo,
                                               int& hi) {
    const double f = dmul(dsub(v, A.gorg[a]), V.inv);
    lo = (int)fmin(fmax(floor(dsub(f, e_cells)), -1.0), (double)n);
    hi = (int)fmin(fmax(floor(dadd(f, e_cells)), -1.0), (double)n);
}

// an upper bound of sqrt(s) within 2^-17 relative, from the single-precision reciprocal square root (two instructions
// instead of the correctly rounded double-precision sequence); for search radii only -- distances handed to the caller
// are always dsqrt
__device__ __forceinline__ double sqrt_upper(double s) {
    const float a = fmaxf(__double2float_ru(s), 1e-30f);
    return dmul((double)(a * rsqrtf(a)), 1.0 + 7.62939453125e-06);
}

__device__ __forceinline__ uint2 grid_entry(const GridView& V, int x, int y, int z) {
    return __ldg(V.g + ((long long)z * V.ny + y) * V.nx + x);
}

// Seed for a query without a previous match: squared distance to a real target point of q's own base-level cell
// (clamped into the grid, so queries outside the cloud's box get a seed too) or of a face neighbour; +inf if none.
__device__ __forceinline__ double walk_seed(const NNArgs& A, const double qx, const double qy, const double qz) {
    double Sd = ICPB_INF;  // any real target point is a valid seed -- a closer one only makes the walk cheaper
    const GridView V0 = grid_view(A, A.gbase);
    int ix = grid_cell_index(A, V0, qx, 0, V0.nx), iy = grid_cell_index(A, V0, qy, 1, V0.ny), iz = grid_cell_index(A, V0, qz, 2, V0.nz);
    ix = min(max(ix, 0), V0.nx - 1);
    iy = min(max(iy, 0), V0.ny - 1);
    iz = min(max(iz, 0), V0.nz - 1);
    uint2 e = make_uint2(0u, 0u);
#pragma unroll 1
    for (int t = 0; t < 7; ++t) {
        const int dz = (t == 1) ? -1 : (t == 2) ? 1 : 0, dy = (t == 3) ? -1 : (t == 4) ? 1 : 0, dx = (t == 5) ? -1 : (t == 6) ? 1 : 0;
        const int x = ix + dx, y = iy + dy, z = iz + dz;
        if (x < 0 || y < 0 || z < 0 || x >= V0.nx || y >= V0.ny || z >= V0.nz) continue;
        e = grid_entry(V0, x, y, z);
        if ((e.y >> 30) != 0u) break;
    }
    const uint32_t kind = e.y >> 30;
    if (kind == 0u) return ICPB_INF;
    uint32_t pt0 = e.x, npts = e.y & 0xFFFFFFu;
    if (kind == 3u) {
        uint32_t n = e.x;
        NodeRegs nd = load_node(A.nodes, n);
        for (;;) {
            const uint32_t mask = nd.meta & 0xFFu;
            if (mask == 0u || nd.npts <= 32u) break;
            uint32_t oct = 0;
            oct |= (qx > dmul(dadd(nd.lo[0], nd.hi[0]), 0.5)) ? 1u : 0u;
            oct |= (qy > dmul(dadd(nd.lo[1], nd.hi[1]), 0.5)) ? 2u : 0u;
            oct |= (qz > dmul(dadd(nd.lo[2], nd.hi[2]), 0.5)) ? 4u : 0u;
            if (!((mask >> oct) & 1u)) break;
            n = nd.child0 + __popc(mask & ((1u << oct) - 1u));
            nd = load_node(A.nodes, n);
        }
        pt0 = nd.pt0;
        npts = nd.npts;
    }
    const uint32_t ns = npts < 32u ? npts : 32u;
    for (uint32_t k = 0; k < ns; k += 4) {
        double px[4], py[4], pz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t kk = (k + j < ns) ? k + j : ns - 1u;
            uint32_t pidx;
            load_point(A.pts, pt0 + kk, px[j], py[j], pz[j], pidx);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Sd = fmin(Sd, sumsq3(dsub(px[j], qx), dsub(py[j], qy), dsub(pz[j], qz)));
    }
    if (!(Sd < 1e19)) return ICPB_INF;
    return Sd;
}

template <int STRIDE>
__device__ __forceinline__ bool cell_walk(const NNArgs& A, const double qx, const double qy, const double qz, double Sd,
                                          uint2* stk, Fast& F) {
    if (!(Sd < 1e19)) {
        Sd = walk_seed(A, qx, qy, qz);
        if (!(Sd < 1e19)) return false;
    }
    const double r = dmul(dsqrt(Sd), 1.0 + 9.5367431640625e-07);  // sqrt(S) (1 + 2^-20)
    const double e = dadd(r, A.geps);
    const GridView V = grid_view(A, grid_level_for_width(A, dmul(e, 2.0), A.gbias));
    int x0 = grid_cell_index(A, V, dsub(qx, e), 0, V.nx), x1 = grid_cell_index(A, V, dadd(qx, e), 0, V.nx);
    int y0 = grid_cell_index(A, V, dsub(qy, e), 1, V.ny), y1 = grid_cell_index(A, V, dadd(qy, e), 1, V.ny);
    int z0 = grid_cell_index(A, V, dsub(qz, e), 2, V.nz), z1 = grid_cell_index(A, V, dadd(qz, e), 2, V.nz);
    x0 = max(x0, 0); y0 = max(y0, 0); z0 = max(z0, 0);
    x1 = min(x1, V.nx - 1); y1 = min(y1, V.ny - 1); z1 = min(z1, V.nz - 1);
    if ((long long)(x1 - x0 + 1) * (long long)(y1 - y0 + 1) * (long long)(z1 - z0 + 1) > (long long)A.gmax_cells) return false;
    F.best = ICPB_INF;
    F.second = ICPB_INF;
    F.pos = NONE;
    F.node = NONE;
    F.bound = dmul(Sd, 1.0 + 1.8189894035458565e-12);  // S (1 + 2^-39)
    // One flat loop over (cell, point batch): a lane that has used up its leaf advances to its next non-empty cell
    // while the others wait, then all lanes scan a batch together -- the warp's trip count follows the lane with the
    // most POINTS, not (most cells) x (largest leaf).
    if (x1 < x0 || y1 < y0 || z1 < z0) return true;  // the ball misses the grid: nothing found, the caller falls back
    int x = x0 - 1, y = y0, z = z0;       // cell cursor (x fastest)
    bool more = true;
    uint32_t pt0 = 0, npts = 0, k = 0;
    for (;;) {
        while (k >= npts) {
            if (++x > x1) {
                x = x0;
                if (++y > y1) {
                    y = y0;
                    if (++z > z1) {
                        more = false;
                        break;
                    }
                }
            }
            const uint2 en = grid_entry(V, x, y, z);
            const uint32_t kind = en.y >> 30;
            if (kind == 0u) continue;
            if (kind == 3u) {
                fast_search<STRIDE>(A.nodes, A.pts, qx, qy, qz, en.x, F, stk);
                continue;
            }
            if (kind == 2u) {
                // a leaf above the grid level owns an aligned block of cells: scan it once, from the first cell
                // that the block and this query's range have in common
                const int sh = V.level - (int)((en.y >> 24) & 0x3Fu);
                if (x != max((x >> sh) << sh, x0) || y != max((y >> sh) << sh, y0) || z != max((z >> sh) << sh, z0)) continue;
            }
            pt0 = en.x;
            npts = en.y & 0xFFFFFFu;
            k = 0;
        }
        if (!more) break;
        double px[4], py[4], pz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t kk = (k + j < npts) ? k + j : npts - 1u;
            uint32_t pidx;
            load_point(A.pts, pt0 + kk, px[j], py[j], pz[j], pidx);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double s = sumsq3(dsub(px[j], qx), dsub(py[j], qy), dsub(pz[j], qz));
            if (k + j < npts && s <= F.bound) fast_take(F, s, pt0 + k + j, NONE);
        }
        k += 4;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// One query, one thread: the fast path (cell walk, else a pruned tree search from the deepest node around the query that
// certainly holds the answer) with the literal reference traversal
// as fallback.  Returns the sorted target position of the answer (NONE if the reference accepts no point).
//   pp          last iteration's match (NONE if unknown)
//   extra_seed  squared distance of some real target point already known (or +inf)
// ---------------------------------------------------------------------------------------------------
template <int STRIDE>
__device__ __forceinline__ uint32_t per_thread_query(const NNArgs& A, const double qx, const double qy, const double qz,
                                                     const bool finite_q, const uint32_t pp,
                                                     const double extra_seed, uint2* stk, uint32_t& result_node,
                                                     bool& fell_back, const bool skip_fast = false, double* best_s = nullptr) {
    uint32_t result = NONE;
    result_node = NONE;
    if (best_s) *best_s = -1.0;  // set to s(answer) when a fast path produced the answer (computeDistance = sqrt of it)
    bool need_literal = finite_q;
    bool walked = false;
    if (A.mode == 3 && finite_q && !skip_fast) {
        double Sd = extra_seed;
        if (pp != NONE) {
            double px, py, pz;
            uint32_t pidx;
            load_point(A.pts, pp, px, py, pz, pidx);
            Sd = fmin(Sd, sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz)));
        }
        Fast F;
        walked = cell_walk<STRIDE>(A, qx, qy, qz, Sd, stk, F);
        if (walked && F.pos != NONE && F.second > dmul(F.best, 1.0 + 9.094947017729282e-13)) {
            result = F.pos;
            result_node = F.node;
            need_literal = false;
            if (best_s) *best_s = F.best;
        }
    }
    if (A.mode >= 1 && finite_q && !skip_fast && !walked) {
        double Sd = ICPB_INF;
        uint32_t start = 0;
        {
            // ---- point location: walk down the cell path of q, remembering each level's clearance ----
            uint32_t n = 0;
            int level = 0;
            NodeRegs nd;
            // Running pointer rather than stk[level * STRIDE]: ptxas 12.9 (sm_100a) mis-addressed the indexed
            // form of this store by two rows in the rotated loop (seen in SASS and on the device), PTX was correct.
            uint2* path = stk;
            for (;;) {
                nd = load_node(A.nodes, n);
                double c = fmin(fmin(dsub(qx, nd.lo[0]), dsub(nd.hi[0], qx)),
                                fmin(fmin(dsub(qy, nd.lo[1]), dsub(nd.hi[1], qy)), fmin(dsub(qz, nd.lo[2]), dsub(nd.hi[2], qz))));
                float cf = (c > 0.0) ? __double2float_rd(c) : 0.0f;
                *path = make_uint2(n, __float_as_uint(cf));
                const uint32_t mask = nd.meta & 0xFFu;
                if (mask == 0) break;
                uint32_t oct = 0;
                oct |= (qx > dmul(dadd(nd.lo[0], nd.hi[0]), 0.5)) ? 1u : 0u;
                oct |= (qy > dmul(dadd(nd.lo[1], nd.hi[1]), 0.5)) ? 2u : 0u;
                oct |= (qz > dmul(dadd(nd.lo[2], nd.hi[2]), 0.5)) ? 4u : 0u;
                if (!((mask >> oct) & 1u)) break;
                n = nd.child0 + __popc(mask & ((1u << oct) - 1u));
                ++level;
                path += STRIDE;
            }
            // ---- seed: squared distance of a real target point of the located cell ----
            const uint32_t ns = nd.npts < (uint32_t)SEED_SCAN ? nd.npts : (uint32_t)SEED_SCAN;
            for (uint32_t k = 0; k < ns; ++k) {
                double px, py, pz;
                uint32_t pidx;
                load_point(A.pts, nd.pt0 + k, px, py, pz, pidx);
                Sd = fmin(Sd, sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz)));
            }
            if (pp != NONE) {
                double px, py, pz;
                uint32_t pidx;
                load_point(A.pts, pp, px, py, pz, pidx);
                Sd = fmin(Sd, sumsq3(dsub(px, qx), dsub(py, qy), dsub(pz, qz)));
            }
            Sd = fmin(Sd, extra_seed);
            // ---- subtree start: deepest path node whose clearance^2 exceeds the seed bound ----
            const double clear_req = dmul(Sd, 1.0 + 3.637978807091713e-12);  // S (1 + 2^-38)
            for (int l = level; l > 0; --l) {
                const uint2 e = *path;
                path -= STRIDE;
                const double cf = (double)__uint_as_float(e.y);
                if (cf * cf > clear_req) {
                    start = e.x;
                    break;
                }
            }
        }
        if (Sd < 1e19) {  // also false for inf/NaN; keeps clear of the CLI's initial best 1e20
            Fast F;
            F.best = ICPB_INF;
            F.second = ICPB_INF;
            F.pos = NONE;
            F.node = NONE;
            F.bound = dmul(Sd, 1.0 + 1.8189894035458565e-12);                // S (1 + 2^-39)
            fast_search<STRIDE>(A.nodes, A.pts, qx, qy, qz, start, F, stk);
            // unique minimum with margin 2^-40 => order-independent => the reference's answer
            if (F.pos != NONE && F.second > dmul(F.best, 1.0 + 9.094947017729282e-13)) {
                result = F.pos;
                result_node = F.node;
                need_literal = false;
                if (best_s) *best_s = F.best;
            }
        }
    }
    if (need_literal) {
        // ---- the reference traversal from the root (octree.cpp:175-184) ----
        Search S;
        S.best = A.init_best;
        S.pos = NONE;
        S.idx = 0;
        const NodeRegs root = load_node(A.rnodes, 0);
        const double mdr = dsqrt(sumsq3(axis_dist(root.lo[0], root.hi[0], qx), axis_dist(root.lo[1], root.hi[1], qy),
                                        axis_dist(root.lo[2], root.hi[2], qz)));
        if (!(dmul(mdr, mdr) >= S.best)) dfs_literal<STRIDE>(A.rnodes, A.rpts, qx, qy, qz, S, stk);
        result = (S.pos == NONE) ? NONE : __ldg(A.inv + S.idx);  // the reference's pick, as a search-tree position
        fell_back = true;
    }
    return result;
}

// src = T * src (icpengine.cpp:345): ((T0 x + T1 y) + T2 z) + T3 * 1.0, no contraction
__device__ __forceinline__ void apply_T_point(const double* T, double& qx, double& qy, double& qz) {
    const double x = qx, y = qy, z = qz;
    qx = dadd(dadd(dadd(dmul(T[0], x), dmul(T[1], y)), dmul(T[2], z)), T[3]);
    qy = dadd(dadd(dadd(dmul(T[4], x), dmul(T[5], y)), dmul(T[6], z)), T[7]);
    qz = dadd(dadd(dadd(dmul(T[8], x), dmul(T[9], y)), dmul(T[10], z)), T[11]);
}

}  // namespace icpb
