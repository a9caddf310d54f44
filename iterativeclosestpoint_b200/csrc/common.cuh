// Shared device-side types and helpers of libicp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace icpb {

// ---------------------------------------------------------------------------------------------------
// FP64 arithmetic without FMA contraction.  The reference is built for baseline x86-64 (no FMA), so every
// product-sum on the parity path must round twice.  The translation units are compiled with -fmad=false
// as well; these wrappers make the intent explicit where it decides parity.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
// std::max(a,b) = (a < b) ? b : a  (NaN handling differs from fmax; the reference uses std::max)
__device__ __forceinline__ double stdmax(double a, double b) { return (a < b) ? b : a; }

// (x*x + y*y) + z*z exactly as `dx*dx + dy*dy + dz*dz` evaluates in C++ without contraction.
__device__ __forceinline__ double sumsq3(double x, double y, double z) {
    return dadd(dadd(dmul(x, x), dmul(y, y)), dmul(z, z));
}

// ---------------------------------------------------------------------------------------------------
// Linear octree node, 64 bytes, one per 64-byte line.  Nodes are numbered level by level (BFS); the
// existing children of an inner node are consecutive, in octant order, starting at `child0`.
// Replaces class OctreeNode (core/octree.h:10-22): box, leaf flag, child pointers, leaf index list.
// ---------------------------------------------------------------------------------------------------
struct __align__(64) Node {
    double lo[3];     // min_x, min_y, min_z
    double hi[3];     // max_x, max_y, max_z
    uint32_t child0;  // inner: node index of the first existing child
    uint32_t pt0;     // first position (Morton-sorted target order) of the points below this node
    uint32_t npts;    // number of points below this node (leaf: the leaf's point count)
    uint32_t meta;    // bits 0-7: mask of existing octants (0 => leaf); bits 8-15: depth
};
static_assert(sizeof(Node) == 64, "Node must be 64 bytes");

// Target point in Morton-sorted order: xyz + original index (as the low 32 bits of the 4th lane).
struct __align__(32) TPoint {
    double x, y, z;
    long long idx;
};
static_assert(sizeof(TPoint) == 32, "TPoint must be 32 bytes");

// Chan/Welford partial of the distance statistics (stage A): count, mean, M2, min, max over finite
// values, number of problem (NaN/Inf/out-of-range) distances.
struct StatA {
    double n, mean, m2, dmin, dmax, problems;
};

__device__ __forceinline__ StatA stat_merge(const StatA& a, const StatA& b) {
    StatA r;
    r.n = a.n + b.n;
    if (r.n == 0.0) {
        r.mean = 0.0;
        r.m2 = 0.0;
    } else {
        double delta = b.mean - a.mean;
        double f = b.n / r.n;
        r.mean = a.mean + delta * f;
        r.m2 = a.m2 + b.m2 + delta * delta * a.n * f;
    }
    r.dmin = fmin(a.dmin, b.dmin);
    r.dmax = fmax(a.dmax, b.dmax);
    r.problems = a.problems + b.problems;
    return r;
}

// Stage-B partial: inlier count, sum d^2 over inliers, pivoted first and second moments.
struct StatB {
    double n;       // inlier count
    double sumsq;   // sum d_i^2 over inliers
    double sa[3];   // sum (a - pa)
    double sb[3];   // sum (b - pb)
    double sab[9];  // sum (a - pa)(b - pb)^T, row-major
};
// 17 sums + one control word: ranks that were asked to stop put 1 there, so the sum over ranks tells every rank, in the same
// iteration, that the run is cancelled (ICPEngine::stop(), core/icpengine.cpp:62-66,160-164, on a sharded run)
static constexpr int STATB_DOUBLES = 18;
static constexpr int STATB_STOP = 17;

// Per-run device state shared by the iteration kernels (one instance per handle, in device memory).
struct LoopState {
    // stage A result
    StatA a;
    double mean, std_dev, threshold;
    // stage B result
    StatB b;
    double rmse;
    // loop control (core/icpengine.cpp:156-157, 287-323)
    double prev_error;
    int no_improve;
    int iter;        // loop index of the iteration being processed
    int exit_code;   // 0 continue, 1 converged, 2 diverged, 3 too few inliers
    int have_T;      // the NN kernel must apply T_pending on load
    double T_pending[16];
    double T_last[16];
    double T_cum[16];
    double pivot_a[3], pivot_b[3];
    // parameters
    double tolerance, sigma;
    int variant, max_iterations;
    long long n_global;  // N of mean / variance (global source size)
    unsigned int ticket_a, ticket_b;
};

// Host-visible record written once per iteration by the solve step.
// Peer-to-peer exchange of the two per-iteration records (multi-GPU, one process per GPU): every rank owns one Mailbox in
// device memory, opened by all its peers through CUDA IPC.  The producing kernel's last block stores its record straight
// into every peer's mailbox over NVLink and then the epoch into that peer's flag; consumers spin on their OWN mailbox's
// flags (local memory) until every rank's epoch has arrived.  No collective launch, no host round trip.
constexpr int MAIL_RANKS = 8;
struct Mailbox {
    unsigned int flag_a[MAIL_RANKS];  // epoch of the stage-A record last written by rank r
    unsigned int flag_b[MAIL_RANKS];  // ... stage-B record
    StatA a[MAIL_RANKS];
    double b[MAIL_RANKS][STATB_DOUBLES];
};
struct PeerMail {
    Mailbox* peer[MAIL_RANKS];  // peer[r] = rank r's mailbox as mapped into this process (peer[rank] = the own one)
    int n_ranks, rank;
    unsigned int epoch;         // > 0: exchange through the mailboxes; 0: records are gathered by the caller (NCCL) or single rank
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until every rank's epoch in `flags` has reached `epoch` (wrap-safe).  A peer that died would leave this kernel spinning
// for ever, so the wait gives up after ~4 s of GPU time and reports it (the callers then let the run fail cleanly).
__device__ __forceinline__ bool mail_wait(const unsigned int* flags, int n_ranks, unsigned int epoch) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int r = 0; r < n_ranks; ++r) {
        unsigned int spins = 0;
        while ((int)(ld_acquire_sys(flags + r) - epoch) < 0) {
            if ((++spins & 0xFFFFu) == 0u) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 4000000000ull) return false;
            }
        }
    }
    return true;
}

struct IterRecord {
    int iteration;
    int valid_points;
    int outlier_points;
    int exit_code;
    double rmse, mean, std_dev, threshold, dmin, dmax, problems;
    double T_cum[16];
    double T_last[16];
};

}  // namespace icpb
