// Many small independent registrations (BASELINE.json config #5: 4096 pairs of 2k points): ONE thread block runs
// one whole registration -- the complete loop of core/icpengine.cpp:117-394 -- without leaving the SM.
//
//   target    SoA in shared memory (<= 4096 points, 96 KB)
//   source    in registers, 4 points per thread (<= 2048 points), moved in place each iteration
//   NN        exhaustive: every thread walks the whole target (shared-memory broadcast reads) keeping, per query, the
//             smallest and second smallest value of the reference's squared-distance expression.  A unique minimum
//             (margin 2^-40) is the reference's answer whatever its octree does (argument in nn.cu); a query without
//             one (exact tie, duplicate target points) or with non-finite coordinates flags the PAIR, which the host
//             then sends through the general path (octrees, literal traversal).
//   stats     Welford per thread -> fixed-tree Chan merge; inlier mask; pivoted moments; fixed-order block sums
//   solve     solve_step() of solve.cuh on thread 0: the same SVD / Kabsch / loop-control code as the large path
// No host round trip per iteration: the block writes one IterRecord per iteration and the host rebuilds the
// reference's result structures from them (RunAcc in api.cu).
#include "internal.h"
#include "solve.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace icpb {

constexpr int SB_THREADS = 512;
constexpr int SB_QPT = 4;                          // source points per thread
constexpr int SB_MAX_SRC = SB_THREADS * SB_QPT;    // 2048
constexpr int SB_MAX_TGT = 4096;
constexpr unsigned SB_FULL = 0xffffffffu;
#define SB_INF __longlong_as_double(0x7FF0000000000000LL)

struct SmallPair {
    long long src_off, tgt_off;  // first point of the pair in the packed source / target arrays
    int n_src, n_tgt;
};

struct SmallOut {
    int flagged;       // 1: needs the general path (tie / non-finite / nothing accepted)
    int n_records;     // IterRecords written
    int exit_code;     // of the last iteration (0 = ran out of iterations)
    int pad;
};

struct SmallParams {
    double tolerance, sigma;
    int variant, max_iterations, rec_cap;
};

__device__ __forceinline__ StatA sb_stat_empty() {
    StatA s;
    s.n = 0.0; s.mean = 0.0; s.m2 = 0.0; s.dmin = DBL_MAX; s.dmax = 0.0; s.problems = 0.0;
    return s;
}

__global__ void __launch_bounds__(SB_THREADS, 1) icp_small_kernel(const SmallPair* __restrict__ pairs, int n_pairs,
                                                                   double* __restrict__ src_xyz, const double* __restrict__ tgt_xyz,
                                                                   const SmallParams P, IterRecord* __restrict__ recs,
                                                                   SmallOut* __restrict__ outs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tx = reinterpret_cast<double*>(smem_raw);
    double* ty = tx + SB_MAX_TGT;
    double* tz = ty + SB_MAX_TGT;
    StatA* s_stat = reinterpret_cast<StatA*>(tz + SB_MAX_TGT);           // [SB_THREADS]
    double* s_red = reinterpret_cast<double*>(s_stat + SB_THREADS);      // [SB_THREADS / 32][STATB_DOUBLES] + 32
    LoopState* st = reinterpret_cast<LoopState*>(s_red + (SB_THREADS / 32) * STATB_DOUBLES + 32);
    IterRecord* rec = reinterpret_cast<IterRecord*>(st + 1);
    __shared__ int s_flag;
    __shared__ double s_thr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        const SmallPair pr = pairs[p];
        const int n = pr.n_src, m = pr.n_tgt;
        __syncthreads();
        // ---- target -> shared memory; bounding box for the first pivot ----------------------------------------
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int j = tid; j < m; j += SB_THREADS) {
            const double* t = tgt_xyz + 3 * (pr.tgt_off + j);
            const double x = t[0], y = t[1], z = t[2];
            tx[j] = x; ty[j] = y; tz[j] = z;
            lo[0] = fmin(lo[0], x); hi[0] = fmax(hi[0], x);
            lo[1] = fmin(lo[1], y); hi[1] = fmax(hi[1], y);
            lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fmin(lo[a], __shfl_xor_sync(SB_FULL, lo[a], o));
                hi[a] = fmax(hi[a], __shfl_xor_sync(SB_FULL, hi[a], o));
            }
            if (lane == 0) {
                s_red[warp * 6 + a] = lo[a];
                s_red[warp * 6 + 3 + a] = hi[a];
            }
        }
        if (tid == 0) s_flag = 0;
        __syncthreads();
        if (tid == 0) {
            double l[3], h[3];
            for (int a = 0; a < 3; ++a) {
                l[a] = s_red[a];
                h[a] = s_red[3 + a];
                for (int w = 1; w < SB_THREADS / 32; ++w) {
                    l[a] = fmin(l[a], s_red[w * 6 + a]);
                    h[a] = fmax(h[a], s_red[w * 6 + 3 + a]);
                }
            }
            LoopState z;
            z.a = sb_stat_empty();
            z.mean = z.std_dev = z.threshold = 0.0;
            for (int k = 0; k < 9; ++k) z.b.sab[k] = 0.0;
            z.b.n = z.b.sumsq = 0.0;
            z.rmse = 0.0;
            z.prev_error = 1e10;  // icpengine.cpp:156
            z.no_improve = 0;
            z.iter = 0;
            z.exit_code = 0;
            z.have_T = 0;
            for (int k = 0; k < 16; ++k) z.T_pending[k] = z.T_last[k] = z.T_cum[k] = (k % 5 == 0) ? 1.0 : 0.0;
            for (int a = 0; a < 3; ++a) {
                z.b.sa[a] = z.b.sb[a] = 0.0;
                z.pivot_a[a] = z.pivot_b[a] = 0.5 * ((l[a] - 0.001) + (h[a] + 0.001));  // centre of the reference's root box
            }
            z.tolerance = P.tolerance;
            z.sigma = P.sigma;
            z.variant = P.variant;
            z.max_iterations = P.max_iterations;
            z.n_global = n;
            z.ticket_a = z.ticket_b = 0;
            *st = z;
        }
        // ---- source -> registers ------------------------------------------------------------------------------------
        double qx[SB_QPT], qy[SB_QPT], qz[SB_QPT];
        bool valid[SB_QPT];
#pragma unroll
        for (int k = 0; k < SB_QPT; ++k) {
            const int i = tid + k * SB_THREADS;
            valid[k] = i < n;
            qx[k] = qy[k] = qz[k] = 0.0;
            if (valid[k]) {
                const double* sp = src_xyz + 3 * (pr.src_off + i);
                qx[k] = sp[0]; qy[k] = sp[1]; qz[k] = sp[2];
                if (!(isfinite(qx[k]) && isfinite(qy[k]) && isfinite(qz[k]))) s_flag = 1;
            }
        }
        __syncthreads();

        int n_rec = 0, exit_code = 0;
        for (int iter = 0; iter < P.max_iterations && s_flag == 0; ++iter) {
            // ---- pending transform (icpengine.cpp:345) ----------------------------------------------------------------
            if (st->have_T) {
                double T[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) T[k] = st->T_pending[k];
#pragma unroll
                for (int k = 0; k < SB_QPT; ++k) {
                    const double x = qx[k], y = qy[k], z = qz[k];
                    qx[k] = dadd(dadd(dadd(dmul(T[0], x), dmul(T[1], y)), dmul(T[2], z)), T[3]);
                    qy[k] = dadd(dadd(dadd(dmul(T[4], x), dmul(T[5], y)), dmul(T[6], z)), T[7]);
                    qz[k] = dadd(dadd(dadd(dmul(T[8], x), dmul(T[9], y)), dmul(T[10], z)), T[11]);
                }
            }
            // ---- exhaustive NN ---------------------------------------------------------------------------------------------
            double best[SB_QPT], second[SB_QPT];
            int bi[SB_QPT];
#pragma unroll
            for (int k = 0; k < SB_QPT; ++k) {
                best[k] = SB_INF;
                second[k] = SB_INF;
                bi[k] = -1;
            }
#pragma unroll 2
            for (int j = 0; j < m; ++j) {
                const double x = tx[j], y = ty[j], z = tz[j];
#pragma unroll
                for (int k = 0; k < SB_QPT; ++k) {
                    const double s = sumsq3(dsub(x, qx[k]), dsub(y, qy[k]), dsub(z, qz[k]));
                    if (s < second[k]) {
                        if (s < best[k]) {
                            second[k] = best[k];
                            best[k] = s;
                            bi[k] = j;
                        } else {
                            second[k] = s;
                        }
                    }
                }
            }
            // ---- distances, stage A -----------------------------------------------------------------------------------------------
            StatA acc = sb_stat_empty();
            double d[SB_QPT];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < SB_QPT; ++k) {
                d[k] = 0.0;
                if (!valid[k]) continue;
                // unique minimum with margin, and a value the reference's initial best (DBL_MAX / 1e20) would accept
                if (bi[k] < 0 || !(best[k] < 1e19) || !(second[k] > dmul(best[k], 1.0 + 9.094947017729282e-13))) bad = true;
                d[k] = dsqrt(best[k]);  // computeDistance (icpengine.cpp:68-74): same sum of the same squares
                acc.n += 1.0;
                const double delta = d[k] - acc.mean;
                acc.mean += delta / acc.n;
                acc.m2 += delta * (d[k] - acc.mean);
                if (isfinite(d[k])) {
                    acc.dmin = fmin(acc.dmin, d[k]);
                    acc.dmax = fmax(acc.dmax, d[k]);
                } else {
                    acc.problems += 1.0;
                }
            }
            if (bad) s_flag = 1;
            s_stat[tid] = acc;
            __syncthreads();
            for (int s = 1; s < SB_THREADS; s <<= 1) {
                if ((tid % (2 * s)) == 0) s_stat[tid] = stat_merge(s_stat[tid], s_stat[tid + s]);
                __syncthreads();
            }
            if (s_flag) break;  // block-uniform: read after the barrier
            if (tid == 0) {
                const StatA a = s_stat[0];
                const double N = (double)n;
                const double mean = a.mean, sd = dsqrt(ddiv(a.m2, N));
                double thr;
                if (P.variant == ICP_VARIANT_ENGINE && iter == 0)
                    thr = dadd(mean, stdmax(dmul(P.sigma, sd), dmul(mean, 0.5)));  // icpengine.cpp:250-252
                else
                    thr = dadd(mean, dmul(P.sigma, sd));                            // :254 ; CLI :523
                st->a = a;
                st->mean = mean;
                st->std_dev = sd;
                st->threshold = thr;
                st->iter = iter;
                s_thr = thr;
            }
            __syncthreads();
            // ---- stage B: inliers, pivoted moments ----------------------------------------------------------------------------------
            const double thr = s_thr;
            double pa[3], pb[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pa[a] = st->pivot_a[a];
                pb[a] = st->pivot_b[a];
            }
            double v[STATB_DOUBLES];
#pragma unroll
            for (int k = 0; k < STATB_DOUBLES; ++k) v[k] = 0.0;
#pragma unroll
            for (int k = 0; k < SB_QPT; ++k) {
                if (!valid[k] || !(d[k] <= thr)) continue;  // icpengine.cpp:264-268 (NaN => outlier)
                const double a3[3] = {qx[k] - pa[0], qy[k] - pa[1], qz[k] - pa[2]};
                const double b3[3] = {tx[bi[k]] - pb[0], ty[bi[k]] - pb[1], tz[bi[k]] - pb[2]};
                v[0] += 1.0;
                v[1] += d[k] * d[k];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    v[2 + r] += a3[r];
                    v[5 + r] += b3[r];
#pragma unroll
                    for (int q = 0; q < 3; ++q) v[8 + 3 * r + q] += a3[r] * b3[q];
                }
            }
#pragma unroll
            for (int k = 0; k < STATB_DOUBLES; ++k) {
                double x = v[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(SB_FULL, x, o);
                if (lane == 0) s_red[warp * STATB_DOUBLES + k] = x;
            }
            __syncthreads();
            if (tid == 0) {
                double b17[STATB_DOUBLES];
                for (int k = 0; k < STATB_DOUBLES; ++k) {
                    double x = s_red[k];
                    for (int w = 1; w < SB_THREADS / 32; ++w) x += s_red[w * STATB_DOUBLES + k];
                    b17[k] = x;
                }
                // reuse the large path's solve: it reads its input through a pointer
                double* slot = s_red + (SB_THREADS / 32) * STATB_DOUBLES - STATB_DOUBLES;  // last row: already consumed
                for (int k = 0; k < STATB_DOUBLES; ++k) slot[k] = b17[k];
                solve_step(st, slot, 1, rec);
                if (n_rec < P.rec_cap) recs[(long long)p * P.rec_cap + n_rec] = *rec;
            }
            ++n_rec;
            __syncthreads();
            exit_code = st->exit_code;
            if (exit_code != 0) break;
        }
        __syncthreads();
        const bool flagged = s_flag != 0;
        // ---- write the moved source back (the host decides whether the caller gets to see it) ------------------------------------
        if (!flagged) {
            if (st->have_T) {  // ran out of iterations with a transform still pending
                double T[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) T[k] = st->T_pending[k];
#pragma unroll
                for (int k = 0; k < SB_QPT; ++k) {
                    const double x = qx[k], y = qy[k], z = qz[k];
                    qx[k] = dadd(dadd(dadd(dmul(T[0], x), dmul(T[1], y)), dmul(T[2], z)), T[3]);
                    qy[k] = dadd(dadd(dadd(dmul(T[4], x), dmul(T[5], y)), dmul(T[6], z)), T[7]);
                    qz[k] = dadd(dadd(dadd(dmul(T[8], x), dmul(T[9], y)), dmul(T[10], z)), T[11]);
                }
            }
#pragma unroll
            for (int k = 0; k < SB_QPT; ++k) {
                const int i = tid + k * SB_THREADS;
                if (i < n) {
                    double* sp = src_xyz + 3 * (pr.src_off + i);
                    sp[0] = qx[k]; sp[1] = qy[k]; sp[2] = qz[k];
                }
            }
        }
        if (tid == 0) {
            SmallOut o;
            o.flagged = flagged ? 1 : 0;
            o.n_records = n_rec < P.rec_cap ? n_rec : P.rec_cap;
            o.exit_code = exit_code;
            o.pad = 0;
            outs[p] = o;
        }
    }
}

static size_t small_smem_bytes() {
    return (size_t)3 * SB_MAX_TGT * sizeof(double) + (size_t)SB_THREADS * sizeof(StatA) +
           ((size_t)(SB_THREADS / 32) * STATB_DOUBLES + 32) * sizeof(double) + sizeof(LoopState) + sizeof(IterRecord) + 64;
}

bool small_pair_eligible(int64_t n_src, int64_t n_tgt) {
    return n_src >= 1 && n_tgt >= 1 && n_src <= SB_MAX_SRC && n_tgt <= SB_MAX_TGT;
}

// Runs the eligible pairs listed in `which` through the one-block kernel.  On return flagged[k] tells whether pair
// which[k] must be redone by the general path; for the others recs / n_rec / exit hold the iteration records and the
// moved sources have been copied into `moved` (packed like the uploads).
int small_batch_run(Ctx* c, const std::vector<int32_t>& which, double* const* src_xyz, const int64_t* n_src,
                    const double* const* tgt_xyz, const int64_t* n_tgt, std::vector<IterRecord>& recs, int rec_cap,
                    std::vector<int>& n_rec, std::vector<int>& exit_code, std::vector<char>& flagged,
                    const double*& moved_ptr, std::vector<long long>& src_off) {
    const int np = (int)which.size();
    if (np == 0) return ICP_OK;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    std::vector<SmallPair> pairs((size_t)np);
    long long so = 0, to = 0;
    src_off.resize((size_t)np);
    for (int k = 0; k < np; ++k) {
        const int32_t p = which[(size_t)k];
        pairs[(size_t)k] = {so, to, (int)n_src[p], (int)n_tgt[p]};
        src_off[(size_t)k] = so;
        so += n_src[p];
        to += n_tgt[p];
    }
    ICPB_TRY(devbuf_reserve(c, c->scratch_src, (size_t)so * 3 * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->tgt_raw, (size_t)to * 3 * sizeof(double)));
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)np * (sizeof(SmallPair) + sizeof(SmallOut)) + 256));
    ICPB_TRY(devbuf_reserve(c, c->scratch2, (size_t)np * rec_cap * sizeof(IterRecord)));
    double* d_src = (double*)c->scratch_src.p;
    double* d_tgt = (double*)c->tgt_raw.p;
    SmallPair* d_pairs = (SmallPair*)c->scratch1.p;
    SmallOut* d_outs = (SmallOut*)(d_pairs + np);
    IterRecord* d_recs = (IterRecord*)c->scratch2.p;
    c->tree.valid = false;  // tgt_raw is reused
    c->fast.valid = false;
    c->prev_valid = false;
    // pack on the host into pinned staging, several threads, one copy each way (thousands of small pageable copies
    // would cost more than the kernel)
    auto t_start = std::chrono::steady_clock::now();
    ICPB_TRY(pinned_reserve(c, c->pin_a, (size_t)so * 3 * sizeof(double)));
    ICPB_TRY(pinned_reserve(c, c->pin_b, (size_t)to * 3 * sizeof(double)));
    double* h_src = (double*)c->pin_a.p;
    double* h_tgt = (double*)c->pin_b.p;
    const int n_thr = std::max(1, std::min(4, np / 64));
    auto parallel_for_pairs = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < n_thr; ++t)
            th.emplace_back([&, t] { for (int k = t; k < np; k += n_thr) fn(k); });
        for (int k = 0; k < np; k += n_thr) fn(k);
        for (auto& x : th) x.join();
    };
    parallel_for_pairs([&](int k) {
        const int32_t p = which[(size_t)k];
        std::memcpy(h_src + 3 * pairs[(size_t)k].src_off, src_xyz[p], (size_t)n_src[p] * 3 * sizeof(double));
        std::memcpy(h_tgt + 3 * pairs[(size_t)k].tgt_off, tgt_xyz[p], (size_t)n_tgt[p] * 3 * sizeof(double));
    });
    auto t_packed = std::chrono::steady_clock::now();
    ICPB_CUDA(c, cudaEventRecord(c->ev[9], c->stream));
    ICPB_CUDA(c, cudaMemcpyAsync(d_src, h_src, (size_t)so * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ICPB_CUDA(c, cudaMemcpyAsync(d_tgt, h_tgt, (size_t)to * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ICPB_CUDA(c, cudaMemcpyAsync(d_pairs, pairs.data(), pairs.size() * sizeof(SmallPair), cudaMemcpyHostToDevice, c->stream));
    SmallParams P;
    P.tolerance = c->params.tolerance;
    P.sigma = (c->params.variant == ICP_VARIANT_CLI) ? 3.0 : c->params.sigma_multiplier;
    P.variant = c->params.variant;
    P.max_iterations = c->params.max_iterations;
    P.rec_cap = rec_cap;
    const size_t smem = small_smem_bytes();
    ICPB_CUDA(c, cudaFuncSetAttribute(icp_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = std::min(np, c->sm_count * 4);
    ICPB_CUDA(c, cudaEventRecord(c->ev[10], c->stream));
    icp_small_kernel<<<blocks, SB_THREADS, smem, c->stream>>>(d_pairs, np, d_src, d_tgt, P, d_recs, d_outs);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    ICPB_CUDA(c, cudaEventRecord(c->ev[11], c->stream));
    std::vector<SmallOut> outs((size_t)np);
    recs.resize((size_t)np * rec_cap);
    ICPB_CUDA(c, cudaMemcpyAsync(outs.data(), d_outs, outs.size() * sizeof(SmallOut), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaMemcpyAsync(recs.data(), d_recs, recs.size() * sizeof(IterRecord), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaMemcpyAsync(h_src, d_src, (size_t)so * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));  // moved sources
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    moved_ptr = h_src;
    if (getenv("ICP_B200_DEBUG_COUNTERS")) {
        float ms_h2d = 0.f, ms_k = 0.f;
        cudaEventElapsedTime(&ms_h2d, c->ev[9], c->ev[10]);
        cudaEventElapsedTime(&ms_k, c->ev[10], c->ev[11]);
        const double ms_pack = std::chrono::duration<double, std::milli>(t_packed - t_start).count();
        const double ms_all = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
        fprintf(stderr, "[icp_b200] small batch: %d pairs, pack %.1f ms, h2d %.1f ms, kernel %.1f ms, total %.1f ms\n", np, ms_pack, ms_h2d,
                ms_k, ms_all);
    }
    n_rec.resize((size_t)np);
    exit_code.resize((size_t)np);
    flagged.resize((size_t)np);
    for (int k = 0; k < np; ++k) {
        n_rec[(size_t)k] = outs[(size_t)k].n_records;
        exit_code[(size_t)k] = outs[(size_t)k].exit_code;
        flagged[(size_t)k] = (char)outs[(size_t)k].flagged;
    }
    return ICP_OK;
}

}  // namespace icpb
