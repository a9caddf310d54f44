// Exact nearest-neighbour query, one source point per thread (replaces Octree::findNearest /
// Octree::searchNearest, PointCloudRegistration/core/octree.cpp:128-184, and the per-point loop of
// core/icpengine.cpp:172-184; CLI twin icp_registration.cpp:108-151,197-205).
//
// The answer of the reference is defined by its traversal: depth-first, children ordered by the ROOTED box
// distance (ties keep octant order), a node skipped when fl(fl(sqrt(s))^2) >= best, points accepted on a
// strict <.  `dfs_literal` executes exactly that traversal with an explicit per-thread stack in shared
// memory, so its indices agree with the reference bit for bit including every tie.
//
// Fast path (mode 1) -- result-preserving by the following argument (DESIGN.md "NN query"):
//   Let s(p) be the reference's squared-distance expression and suppose one target point P has
//   s(P) * (1 + 2^-40) < s(p) for every other point p.  Then the reference returns P whatever its visiting
//   order: an ancestor box N of P has fl(fl(sqrt(sbox))^2) <= s(P)(1+u)^3, which is below any `best` the
//   reference can hold before it has scanned P (the initial value or some other point's s), so P's leaf is
//   never pruned, P is accepted on the strict <, and nothing can replace it afterwards.
//   The fast path therefore only has to find the exact minimum of s and PROVE the margin.  It searches the
//   same octree with an un-rooted box bound (sbox <= s(p) for every p in the box, by monotonicity of the
//   rounded operations), keeps the two smallest s it sees, prunes only boxes with sbox > min(best, seed) *
//   (1 + 2^-39), and starts at the deepest node around the query whose clearance squared exceeds the seed bound
//   (every box outside that node is farther than the bound).  Any point it did not look at has
//   s > best (1 + 2^-39).  If second <= best (1 + 2^-40) (exact ties, duplicates, 1-ulp near ties) the query
//   is re-run through `dfs_literal`; otherwise the unique minimum is the reference's answer.
#include <algorithm>
#include <cmath>
#include "nn_common.cuh"

namespace icpb {

constexpr int NN_THREADS = 128;

__device__ __forceinline__ StatA warp_merge(StatA v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        StatA w;
        w.n = __shfl_xor_sync(0xffffffffu, v.n, o);
        w.mean = __shfl_xor_sync(0xffffffffu, v.mean, o);
        w.m2 = __shfl_xor_sync(0xffffffffu, v.m2, o);
        w.dmin = __shfl_xor_sync(0xffffffffu, v.dmin, o);
        w.dmax = __shfl_xor_sync(0xffffffffu, v.dmax, o);
        w.problems = __shfl_xor_sync(0xffffffffu, v.problems, o);
        // merge in a lane-independent order so every lane ends with the same bits
        const bool lower = (threadIdx.x & o) == 0;
        v = lower ? stat_merge(v, w) : stat_merge(w, v);
    }
    return v;
}

// One query through the per-thread search.  `moved`: the coordinates in A.sx are final (no pending transform).
__device__ __forceinline__ void nn_one_query(const NNArgs& A, const long long i, const bool moved, uint2* stk, StatA& st,
                                             bool& fell_back) {
    double qx = A.sx[i], qy = A.sy[i], qz = A.sz[i];
    if (!moved && A.apply_pending && A.state->have_T) {
        apply_T_point(A.state->T_pending, qx, qy, qz);
        A.ox[i] = qx;
        A.oy[i] = qy;
        A.oz[i] = qz;
    }
    const bool finite_q = isfinite(qx) && isfinite(qy) && isfinite(qz);
    uint32_t pp = NONE;
    if (A.mode == 3 && A.prev_pos) {
        pp = A.prev_pos[i];
    }
    uint32_t result_node = NONE;
    double best_s;
    const uint32_t result = per_thread_query<NN_THREADS>(A, qx, qy, qz, finite_q, pp, ICPB_INF, stk, result_node, fell_back,
                                                         false, &best_s);
    // findNearest returns index 0 when nothing was accepted (best_idx = 0 initially, octree.cpp:179)
    const uint32_t pos = (result == NONE) ? A.pos_of_idx0 : result;
    // computeDistance(p_src, p_tgt) (icpengine.cpp:68-74) = sqrt(dx*dx + dy*dy + dz*dz) with d = src - tgt; the search
    // evaluated the same sum with d = tgt - src, whose squares are the same doubles, so its value is reused.
    double d;
    if (best_s >= 0.0) {
        d = dsqrt(best_s);
    } else {
        double px, py, pz;
        uint32_t pidx;
        load_point(A.pts, pos, px, py, pz, pidx);
        d = dsqrt(sumsq3(dsub(qx, px), dsub(qy, py), dsub(qz, pz)));
    }
    A.pos_out[i] = pos;
    if (A.node_io && !A.worklist) A.node_io[i] = result_node;
    A.dist_out[i] = d;
    st.n = 1.0;
    st.mean = d;
    if (isfinite(d)) {
        st.dmin = d;
        st.dmax = d;
    } else {
        st.problems = 1.0;
    }
}

__global__ void __launch_bounds__(NN_THREADS) nn_kernel(const NNArgs A) {
    // one stack row per tree level (dynamic: the reference's octree may be up to 63 levels deep, most are 10 - 15)
    extern __shared__ __align__(16) uint2 stack[];
    __shared__ StatA warp_part[NN_THREADS / 32];
    const long long i = (long long)blockIdx.x * NN_THREADS + threadIdx.x;
    uint2* stk = stack + threadIdx.x;
    if (A.state && A.state->exit_code != 0) return;  // the loop has ended: iterations enqueued ahead do nothing

    StatA st;
    st.n = 0.0; st.mean = 0.0; st.m2 = 0.0; st.dmin = DBL_MAX; st.dmax = 0.0; st.problems = 0.0;
    bool fell_back = false;

    if (A.worklist) {
        // second half of mode 4: the queries nn_group_kernel could not settle (already moved, matches untouched).
        // One query per thread while the list is shorter than the grid; the surplus blocks exit at once.
        // These are the hard queries and every one is a chain of dependent loads, so a short list is spread thin: only the
        // first `act` lanes of a warp take entries while that still leaves the list within the warps the GPU holds at once
        // (a quarter as many queries per warp = a quarter of the divergent work in front of its slowest lane).
        const long long cnt = (long long)*A.work_count;
        const int lane = threadIdx.x & 31;
        const long long n_warps = ((long long)gridDim.x * NN_THREADS) >> 5;
        int act = 32;
        while (act > 4 && cnt * (64 / act) <= A.resident_threads && cnt * 2 <= n_warps * act) act >>= 1;
        if (lane < act) {
            for (long long t = (i >> 5) * act + lane; t < cnt; t += n_warps * act) {
                fell_back = false;
                nn_one_query(A, (long long)A.worklist[t], true, stk, st, fell_back);
                if (A.counters) atomicAdd(&A.counters[fell_back ? 1 : 0], 1ull);
            }
        }
        return;
    }
    const bool active = i < A.n;
    if (active) nn_one_query(A, i, false, stk, st, fell_back);
    if (A.counters && A.mode >= 1) {
        const unsigned fb = __ballot_sync(0xffffffffu, fell_back);
        const unsigned ac = __ballot_sync(0xffffffffu, active);
        if ((threadIdx.x & 31) == 0) {
            if (fb) atomicAdd(&A.counters[1], (unsigned long long)__popc(fb));
            atomicAdd(&A.counters[0], (unsigned long long)__popc(ac & ~fb));
        }
    }
    if (A.part_a) {
        st = warp_merge(st);
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = st;
        __syncthreads();
        if (threadIdx.x == 0) {
            StatA r = warp_part[0];
#pragma unroll
            for (int w = 1; w < NN_THREADS / 32; ++w) r = stat_merge(r, warp_part[w]);
            A.part_a[blockIdx.x] = r;
        }
    }
}

int nn_grid_blocks(int64_t n) { return (int)((n + NN_THREADS - 1) / NN_THREADS); }

int nn_group_launch(Ctx* c, const NNArgs& A);  // nn_group.cu
int nn_group_lean_launch(Ctx* c, const NNArgs& A);  // nn_group_lean.cu
int nn_keep_launch(Ctx* c, const NNArgs& A, int k);  // nn_keep.cu

// shared memory of nn_kernel: one row of NN_THREADS stack entries per level of the deeper of the two trees
static int nn_stack_bytes(Ctx* c, size_t* bytes) {
    const int rows = std::max(std::max(c->tree.depth, c->fast.depth) + 3, 8);
    *bytes = (size_t)rows * NN_THREADS * sizeof(uint2);
    if (*bytes > 48 * 1024 && *bytes > c->nn_smem_opt_in) {
        ICPB_CUDA(c, cudaFuncSetAttribute(nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*bytes));
        c->nn_smem_opt_in = *bytes;
    }
    return ICP_OK;
}

int nn_launch(Ctx* c, const NNLaunch& L_in) {
    NNLaunch L = L_in;
    size_t stack_bytes = 0;
    ICPB_TRY(nn_stack_bytes(c, &stack_bytes));
    if (L.mode == 6) L.mode = 4;  // the per-iteration choice between 4 and 5 is run_loop's (api.cu); a stateless query is a plain walk
    if (L.n <= 0) return ICP_OK;
    NNArgs A;
    A.nodes = c->fast.nodes;
    A.pts = c->fast.pts;
    A.rnodes = c->tree.nodes;
    A.rpts = c->tree.pts;
    A.inv = c->fast.inv_perm;
    A.grid = c->fast.grid;
    A.glmin = c->fast.glev_min;
    A.gnlev = c->fast.glev_n;
    A.resident_threads = (long long)c->sm_count * 7 * NN_THREADS;  // nn_kernel: 7 blocks per SM (registers)
    A.gbase = c->fast.gbase;
    A.gkmin = c->fast.gbase;  // the per-thread walks stay on the base level and finer; the balanced kernels may go coarser
    A.gmax_cells = c->opt_walk_max_cells;
    A.gbias = c->opt_walk_bias != -100 ? c->opt_walk_bias : (L.mode == 4 ? 0 : -2);
    A.gcube = c->fast.cube;
    A.gbias_mul = std::ldexp(1.0, -A.gbias);
    for (int k = 0; k < 4; ++k) {
        A.goff[k] = c->fast.goff[k];
        for (int a = 0; a < 3; ++a) A.gdim[k][a] = c->fast.gdim[k][a];
        A.ginv[k] = (k < c->fast.glev_n) ? (double)(1ll << (c->fast.glev_min + k)) / c->fast.cube : 0.0;
        A.gedge[k] = (k < c->fast.glev_n) ? c->fast.cube / (double)(1ll << (c->fast.glev_min + k)) : 0.0;
    }
    for (int a = 0; a < 3; ++a) A.gorg[a] = c->fast.root_lo[a];
    // finest cell * 2^-20 >> any rounding of the bisection boundaries
    A.geps = c->fast.cube / (double)(1ll << (c->fast.glev_min + c->fast.glev_n - 1)) * 9.5367431640625e-07;
    {
        // ... and above the rounding of the coordinates themselves (large offsets, tiny cells)
        double mag = 0.0;
        for (int a = 0; a < 3; ++a) mag = std::max(mag, std::max(std::fabs(c->fast.root_lo[a]), std::fabs(c->fast.root_hi[a])));
        A.geps += 64.0 * 2.220446049250313e-16 * mag;
    }
    A.sx = L.sx; A.sy = L.sy; A.sz = L.sz;
    A.ox = L.ox; A.oy = L.oy; A.oz = L.oz;
    A.n = L.n;
    A.pos_out = L.pos_out;
    A.dist_out = L.dist_out;
    A.prev_pos = L.prev_pos;
    A.node_io = L.node_io;
    A.parent = c->fast.parent;
    A.part_a = L.part_a;
    A.state = L.state;
    A.counters = c->opt_count ? c->d_counters : nullptr;
    A.apply_pending = L.apply_pending;
    A.mode = L.mode;
    A.init_best = L.init_best;
    A.pos_of_idx0 = c->fast.pos_of_idx0;
    A.worklist = nullptr;
    A.work_count = nullptr;
    A.lb_io = (L.mode >= 4 && c->opt_temporal_skip) ? L.lb_io : nullptr;
    A.cand_io = nullptr;
    A.worklist2 = nullptr;
    A.walk_alpha = std::max(c->opt_keep_alpha, 1.0);
    A.walk_wmul = std::ldexp(2.0, -c->opt_keep_bias);
    A.walk_rcap = c->opt_keep_rcap * A.gedge[A.gbase];
    const bool in_place5 = L.ox == L.sx && L.oy == L.sy && L.oz == L.sz;
    if (L.mode == 5 && L.prev_pos && A.lb_io && L.cand_io && in_place5 && L.apply_pending && c->d_work_count && c->node_io.p && c->work2.p) {
        // keep what last iteration's candidates and bound prove; search the rest; the per-thread kernel takes what is left
        A.cand_io = L.cand_io;
        A.worklist = (uint32_t*)c->node_io.p;
        A.worklist2 = (uint32_t*)c->work2.p;
        A.work_count = c->d_work_count;
        ICPB_CUDA(c, cudaMemsetAsync(c->d_work_count, 0, 2 * sizeof(unsigned int), c->stream));
        A.gkmin = 0;
        ICPB_TRY(nn_keep_launch(c, A, c->opt_keep_k));
        A.gkmin = A.gbase;
        A.mode = 3;
        A.gbias = c->opt_walk_bias != -100 ? c->opt_walk_bias : -2;
        A.gbias_mul = std::ldexp(1.0, -A.gbias);
        A.apply_pending = 0;
        A.node_io = nullptr;
        A.worklist = A.worklist2;
        A.work_count = c->d_work_count + 1;
        nn_kernel<<<std::min(nn_grid_blocks(L.n), c->sm_count * 64), NN_THREADS, stack_bytes, c->stream>>>(A);
        c->launches++;
        ICPB_CUDA(c, cudaGetLastError());
        return ICP_OK;
    }
    if (L.mode == 5) A.gbias = c->opt_walk_bias != -100 ? c->opt_walk_bias : 0;
    A.gbias_mul = std::ldexp(1.0, -A.gbias);
    if (L.mode >= 4) {
        // the balanced kernel settles what it can; the per-thread kernel (cell walk, then climb / literal) takes the rest
        const bool in_place = !L.apply_pending || (L.ox == L.sx && L.oy == L.sy && L.oz == L.sz);
        if (in_place && c->d_work_count && c->node_io.p) {
            // Optionally (nn_chunks > 1) the queries are walked in chunks: while the balanced kernel walks chunk k + 1, the
            // per-thread kernel works off chunk k's list on a higher-priority stream, so that its few, long searches fill
            // issue slots the walk leaves idle instead of standing alone at the end of the stage.
            const int nchunk = (L.n >= (1 << 20) && !A.counters) ? std::min(std::max(c->opt_nn_chunks, 1), 8) : 1;
            ICPB_CUDA(c, cudaMemsetAsync(c->d_work_count, 0, 16 * sizeof(unsigned int), c->stream));
            const int walk_bias = A.gbias;
            for (int k = 0; k < nchunk; ++k) {
                const long long lo = ((L.n * k / nchunk) + 127) / 128 * 128, hi = (k + 1 == nchunk) ? L.n : ((L.n * (k + 1) / nchunk) + 127) / 128 * 128;
                if (hi <= lo) continue;
                NNArgs B = A;
                B.sx += lo; B.sy += lo; B.sz += lo;
                if (B.ox) { B.ox += lo; B.oy += lo; B.oz += lo; }
                B.pos_out += lo; B.dist_out += lo;
                if (B.prev_pos) B.prev_pos += lo;
                if (B.lb_io) B.lb_io += lo;
                B.n = hi - lo;
                B.worklist = (uint32_t*)c->node_io.p + lo;  // entries are indices relative to the chunk
                B.work_count = c->d_work_count + 2 + k;
                B.gkmin = 0;
                B.gbias = walk_bias;
                B.gbias_mul = std::ldexp(1.0, -B.gbias);
                ICPB_TRY(B.lb_io ? nn_group_launch(c, B) : nn_group_lean_launch(c, B));
                cudaStream_t fs = c->stream;
                if (nchunk > 1) {
                    ICPB_CUDA(c, cudaEventRecord(c->ev_chunk[k], c->stream));
                    ICPB_CUDA(c, cudaStreamWaitEvent(c->stream_hi, c->ev_chunk[k], 0));
                    fs = c->stream_hi;
                }
                B.gkmin = B.gbase;
                B.mode = 3;
                B.gbias = c->opt_walk_bias != -100 ? c->opt_walk_bias : -2;
                B.gbias_mul = std::ldexp(1.0, -B.gbias);
                B.apply_pending = 0;
                B.node_io = nullptr;
                // (the list is walked with a grid stride: a few waves of blocks are enough however long the chunk is)
                nn_kernel<<<std::min(nn_grid_blocks(B.n), c->sm_count * 32), NN_THREADS, stack_bytes, fs>>>(B);
                c->launches++;
                ICPB_CUDA(c, cudaGetLastError());
            }
            if (nchunk > 1) {
                ICPB_CUDA(c, cudaEventRecord(c->ev_chunk[8], c->stream_hi));
                ICPB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_chunk[8], 0));
            }
            return ICP_OK;
        }
        A.mode = 3;
    }
    nn_kernel<<<nn_grid_blocks(L.n), NN_THREADS, stack_bytes, c->stream>>>(A);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
