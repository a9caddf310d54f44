// Device-side solve step shared by iter.cu (the per-iteration kernels) and batch.cu (whole small registrations in one
// CTA): threshold from stage-A partials, moments -> H -> 3x3 two-sided Jacobi SVD in Eigen's operation order ->
// Kabsch transform with reflection fix -> loop control of core/icpengine.cpp:287-323.
#pragma once
#include "internal.h"

namespace icpb {

__device__ __forceinline__ StatA stat_load_cg(const StatA* p) {
    StatA r;
    const double* d = reinterpret_cast<const double*>(p);
    r.n = __ldcg(d); r.mean = __ldcg(d + 1); r.m2 = __ldcg(d + 2);
    r.dmin = __ldcg(d + 3); r.dmax = __ldcg(d + 4); r.problems = __ldcg(d + 5);
    return r;
}

// mean / std / threshold from the rank partials merged in rank order (icpengine.cpp:235-255)
__device__ __forceinline__ void stat_a_finalize(const LoopState* st, const StatA* rank_part, int n_ranks, int iter, StatA& a,
                                                double& mean, double& sd, double& thr) {
    a = stat_load_cg(rank_part);
    for (int r = 1; r < n_ranks; ++r) a = stat_merge(a, stat_load_cg(rank_part + r));
    const double N = (double)st->n_global;
    mean = a.mean;                      // = (sum d) / N
    sd = dsqrt(ddiv(a.m2, N));          // population std (icpengine.cpp:241-245)
    if (st->variant == ICP_VARIANT_ENGINE && iter == 0) {
        thr = dadd(mean, stdmax(dmul(st->sigma, sd), dmul(mean, 0.5)));  // icpengine.cpp:250-252
    } else {
        thr = dadd(mean, dmul(st->sigma, sd));                            // :254 ; CLI :523
    }
}


// ------------------------------------------------------------------------------------------------
// 3x3 two-sided Jacobi SVD in Eigen 3.3.4's operation order (JacobiSVD.h:663-786, RealSvd2x2.h:19-50,
// Jacobi.h:83-114,428-440).  Row-major 3x3.  One thread; every operation is IEEE double with no
// contraction, so given the same H the result equals the reference's bit for bit.
// ------------------------------------------------------------------------------------------------
struct Rot {
    double c, s;
};

__device__ __forceinline__ void rot_apply(double* x, double* y, int n, int stride, Rot j) {
    if (j.c == 1.0 && j.s == 0.0) return;  // Jacobi.h:308
    for (int i = 0; i < n; ++i) {
        const double xi = x[i * stride], yi = y[i * stride];
        x[i * stride] = dadd(dmul(j.c, xi), dmul(j.s, yi));
        y[i * stride] = dadd(dmul(-j.s, xi), dmul(j.c, yi));
    }
}

__device__ __forceinline__ Rot make_jacobi(double x, double y, double z) {  // Jacobi.h:83-114
    Rot r;
    const double deno = dmul(2.0, fabs(y));
    if (deno < DBL_MIN) {
        r.c = 1.0;
        r.s = 0.0;
        return r;
    }
    const double tau = ddiv(dsub(x, z), deno);
    const double w = dsqrt(dadd(dmul(tau, tau), 1.0));
    double t;
    if (tau > 0.0)
        t = ddiv(1.0, dadd(tau, w));
    else
        t = ddiv(1.0, dsub(tau, w));
    const double sign_t = t > 0.0 ? 1.0 : -1.0;
    const double n = ddiv(1.0, dsqrt(dadd(dmul(t, t), 1.0)));
    r.s = dmul(dmul(dmul(-sign_t, ddiv(y, fabs(y))), fabs(t)), n);
    r.c = n;
    return r;
}

static __device__ void svd3(const double* H, double* U, double* S, double* V) {
    const double precision = 2.0 * DBL_EPSILON;
    const double consider_as_zero = DBL_MIN;
    double W[9];
    double scale = 0.0;
    for (int i = 0; i < 9; ++i) {
        const double a = fabs(H[i]);
        if (a > scale) scale = a;
    }
    if (scale == 0.0) scale = 1.0;
    for (int i = 0; i < 9; ++i) {
        W[i] = ddiv(H[i], scale);
        U[i] = V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    double max_diag = fabs(W[0]);
    if (fabs(W[4]) > max_diag) max_diag = fabs(W[4]);
    if (fabs(W[8]) > max_diag) max_diag = fabs(W[8]);
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 1000) {
        finished = true;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                const double pm = dmul(precision, max_diag);
                const double thr = (consider_as_zero < pm) ? pm : consider_as_zero;
                if (fabs(W[3 * p + q]) > thr || fabs(W[3 * q + p]) > thr) {
                    finished = false;
                    // real_2x2_jacobi_svd on the (p,q) block
                    double m[4] = {W[3 * p + p], W[3 * p + q], W[3 * q + p], W[3 * q + q]};
                    Rot rot1;
                    const double t = dadd(m[0], m[3]);
                    const double d = dsub(m[2], m[1]);
                    if (fabs(d) < DBL_MIN) {
                        rot1.s = 0.0;
                        rot1.c = 1.0;
                    } else {
                        const double u = ddiv(t, d);
                        const double tmp = dsqrt(dadd(1.0, dmul(u, u)));
                        rot1.s = ddiv(1.0, tmp);
                        rot1.c = ddiv(u, tmp);
                    }
                    rot_apply(&m[0], &m[2], 2, 1, rot1);
                    const Rot jr = make_jacobi(m[0], m[1], m[3]);
                    Rot jl;  // rot1 * jr.transpose()
                    jl.c = dsub(dmul(rot1.c, jr.c), dmul(rot1.s, -jr.s));
                    jl.s = dadd(dmul(rot1.c, -jr.s), dmul(rot1.s, jr.c));
                    rot_apply(&W[3 * p], &W[3 * q], 3, 1, jl);
                    rot_apply(&U[p], &U[q], 3, 3, jl);
                    Rot jrt;
                    jrt.c = jr.c;
                    jrt.s = -jr.s;
                    rot_apply(&W[p], &W[q], 3, 3, jrt);
                    rot_apply(&V[p], &V[q], 3, 3, jrt);
                    const double a = fabs(W[3 * p + p]), b = fabs(W[3 * q + q]);
                    const double mx = a < b ? b : a;
                    max_diag = max_diag < mx ? mx : max_diag;
                }
            }
    }
    for (int i = 0; i < 3; ++i) {
        const double a = W[3 * i + i];
        S[i] = fabs(a);
        if (a < 0.0)
            for (int r = 0; r < 3; ++r) U[3 * r + i] = -U[3 * r + i];
    }
    for (int i = 0; i < 3; ++i) S[i] = dmul(S[i], scale);
    for (int i = 0; i < 3; ++i) {
        int pos = 0;
        double mx = S[i];
        for (int k = i + 1; k < 3; ++k)
            if (S[k] > mx) {
                mx = S[k];
                pos = k - i;
            }
        if (mx == 0.0) break;
        if (pos) {
            pos += i;
            double tmp = S[i]; S[i] = S[pos]; S[pos] = tmp;
            for (int r = 0; r < 3; ++r) {
                tmp = U[3 * r + i]; U[3 * r + i] = U[3 * r + pos]; U[3 * r + pos] = tmp;
                tmp = V[3 * r + i]; V[3 * r + i] = V[3 * r + pos]; V[3 * r + pos] = tmp;
            }
        }
    }
}

// length-3 inner product in the order the reference build evaluates it: rows 0-1 of a fixed-size product
// accumulate left to right (SSE2 packet path), row 2 goes through the unrolled reduction a0 + (a1 + a2)
__device__ __forceinline__ double dot3_row(int row, double a0, double b0, double a1, double b1, double a2, double b2) {
    if (row < 2) return dadd(dadd(dmul(a0, b0), dmul(a1, b1)), dmul(a2, b2));
    return dadd(dmul(a0, b0), dadd(dmul(a1, b1), dmul(a2, b2)));
}

// icpengine.cpp:93-112: R = V U^T ; det < 0 -> negate V's last column ; t = cB - R cA ; T row-major
static __device__ void solve_from_H(const double* H, const double* cA, const double* cB, double* T, double* Uo, double* So,
                             double* Vo) {
    double U[9], S[3], V[9], R[9];
    svd3(H, U, S, V);
    if (Uo)
        for (int i = 0; i < 9; ++i) {
            Uo[i] = U[i];
            Vo[i] = V[i];
        }
    if (So)
        for (int i = 0; i < 3; ++i) So[i] = S[i];
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                R[3 * i + j] = dot3_row(i, V[3 * i], U[3 * j], V[3 * i + 1], U[3 * j + 1], V[3 * i + 2], U[3 * j + 2]);
        if (pass == 1) break;
        const double det = dadd(dsub(dmul(R[0], dsub(dmul(R[4], R[8]), dmul(R[5], R[7]))),
                                     dmul(R[1], dsub(dmul(R[3], R[8]), dmul(R[5], R[6])))),
                                dmul(R[2], dsub(dmul(R[3], R[7]), dmul(R[4], R[6]))));
        if (!(det < 0.0)) break;
        for (int r = 0; r < 3; ++r) V[3 * r + 2] = dmul(V[3 * r + 2], -1.0);
    }
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = dsub(cB[i], dot3_row(i, R[3 * i], cA[0], R[3 * i + 1], cA[1], R[3 * i + 2], cA[2]));
    }
}

__device__ __forceinline__ void mat4_mul(const double* A, const double* B, double* C) {  // icpengine.cpp:342
    double out[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            out[4 * i + j] = dadd(dadd(dadd(dmul(A[4 * i], B[j]), dmul(A[4 * i + 1], B[4 + j])), dmul(A[4 * i + 2], B[8 + j])),
                                  dmul(A[4 * i + 3], B[12 + j]));
    for (int i = 0; i < 16; ++i) C[i] = out[i];
}

// centroids and H from pivoted sums: cA = pa + sa/n ; H = sab - sa sb^T / n
__device__ __forceinline__ void moments_to_H(const double* b17, const double* pa, const double* pb, double* cA, double* cB,
                                             double* H) {
    const double n = b17[0];
    double ma[3], mb[3];
    for (int r = 0; r < 3; ++r) {
        ma[r] = b17[2 + r] / n;
        mb[r] = b17[5 + r] / n;
        cA[r] = pa[r] + ma[r];
        cB[r] = pb[r] + mb[r];
    }
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) H[3 * r + q] = b17[8 + 3 * r + q] - b17[2 + r] * mb[q];
}

// Sums the rank partials in rank order, then RMSE, loop control and the Kabsch solve.
static __device__ __noinline__ void solve_step(LoopState* st, const double* rank_parts, int n_ranks, IterRecord* rec) {
    double b[STATB_DOUBLES];
    for (int k = 0; k < STATB_DOUBLES; ++k) {
        double v = rank_parts[k];  // plain loads: written by this thread, an earlier kernel, or shared memory (batch.cu)
        for (int r = 1; r < n_ranks; ++r) v += rank_parts[(int64_t)r * STATB_DOUBLES + k];
        b[k] = v;
    }
    if (b[STATB_STOP] > 0.0) {
        // some rank was asked to stop: every rank leaves here, in the same iteration, with nothing of this iteration recorded
        // (the reference polls its flag before the iteration's work, core/icpengine.cpp:160-164)
        st->exit_code = 4;
        st->have_T = 0;
        rec->iteration = st->iter + 1;
        rec->exit_code = 4;
        __threadfence_system();
        return;
    }
    const double valid = b[0];
    const double rmse = valid > 0.0 ? dsqrt(ddiv(b[1], valid)) : 0.0;  // icpengine.cpp:274
    st->rmse = rmse;
    int exit_code = 0;
    const double improvement = dsub(st->prev_error, rmse);  // :288
    if (fabs(improvement) < st->tolerance) {
        st->no_improve++;
        if (st->no_improve >= 3) exit_code = 1;  // converged (:291-305)
    } else {
        st->no_improve = 0;
    }
    if (exit_code == 0 && rmse > dmul(st->prev_error, 1.1)) exit_code = 2;  // diverged (:311-314)
    if (exit_code == 0) {
        st->prev_error = rmse;                    // :316
        if (valid < 3.0) exit_code = 3;           // :319-323
    }
    st->have_T = 0;
    if (exit_code == 0) {
        double cA[3], cB[3], H[9], T[16];
        moments_to_H(b, st->pivot_a, st->pivot_b, cA, cB, H);
        solve_from_H(H, cA, cB, T, nullptr, nullptr, nullptr);
        double Tc[16];
        mat4_mul(T, st->T_cum, Tc);
        for (int i = 0; i < 16; ++i) {
            st->T_cum[i] = Tc[i];
            st->T_last[i] = T[i];
            st->T_pending[i] = T[i];
        }
        st->have_T = 1;
        // after the move the inlier centroid of the source coincides with cB: pivot both sides there
        for (int r = 0; r < 3; ++r) {
            st->pivot_a[r] = cB[r];
            st->pivot_b[r] = cB[r];
        }
    }
    st->exit_code = exit_code;
    rec->iteration = st->iter + 1;
    rec->valid_points = (int)valid;
    rec->outlier_points = (int)((double)st->n_global - valid);
    rec->exit_code = exit_code;
    rec->rmse = rmse;
    rec->mean = st->mean;
    rec->std_dev = st->std_dev;
    rec->threshold = st->threshold;
    rec->dmin = st->a.dmin;
    rec->dmax = st->a.dmax;
    rec->problems = st->a.problems;
    for (int i = 0; i < 16; ++i) {
        rec->T_cum[i] = st->T_cum[i];
        rec->T_last[i] = st->T_last[i];
    }
    __threadfence_system();
}


}  // namespace icpb
