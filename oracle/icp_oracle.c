/* TEST INFRASTRUCTURE ONLY -- the product path (iterativeclosestpoint_b200/csrc, libicp_b200.so) never
 * includes, links or calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so, and only as the checker / the CPU arm.
 *
 * Plain-C CPU restatement of the reference's ICP hot path, written from the algorithm's description:
 *   octree build          PointCloudRegistration/core/octree.cpp:41-126   (CLI twin icp_registration.cpp:66-185)
 *   exact 1-NN search     core/octree.cpp:32-38,128-184                   (CLI twin :50-55,108-151,197-205)
 *   distances/stats/mask  core/icpengine.cpp:187-278                      (CLI twin :499-541)
 *   loop control          core/icpengine.cpp:156-164,287-323              (CLI twin :548-570)
 *   Kabsch solve          core/icpengine.cpp:76-115                       (CLI twin :389-440)
 *   3x3 two-sided Jacobi  Eigen/src/SVD/JacobiSVD.h:663-786, Eigen/src/misc/RealSvd2x2.h:19-50,
 *                         Eigen/src/Jacobi/Jacobi.h:83-114,294-299,308-314,428-440
 *   accumulate + apply    core/icpengine.cpp:342-346
 *   finalisation          core/icpengine.cpp:349-393                      (CLI twin :609-621)
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this restatement is
 * pinned against the reference ITSELF, compiled unmodified into oracle/_ref/ (see oracle/Makefile and
 * tests/test_oracle_vs_ref.py) and against the vectors that build produced (tests/golden/).
 * Everything that is a plain loop in the reference is reproduced bit-for-bit (tree, NN indices, distances,
 * mean/std/threshold, masks, RMSE, SVD given H, apply).  The one stage that cannot be is the 3xN.Nx3
 * cross-covariance, which the reference sends through Eigen's cache-size-dependent blocked GEMM; here it is
 * a sequential sum in inlier order, so H (and from iteration 2 on everything downstream) agrees to ~1e-15
 * relative, not bitwise.  Compile with -O2 -ffp-contract=off (no FMA contraction), as the reference build.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define ORC_VARIANT_ENGINE 0
#define ORC_VARIANT_CLI 1

/* ------------------------------------------------------------------------------------------------------
 * Octree.  Nodes live in one growable array; children are stored as array slots (-1 = absent).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    double lo[3], hi[3]; /* box: lo = min_x,min_y,min_z ; hi = max_x,max_y,max_z */
    int32_t child[8];    /* node slot per octant or -1 */
    int64_t first;       /* leaf: offset into leaf_idx */
    int32_t count;       /* leaf: number of point indices */
    int32_t is_leaf;
} orc_node;

typedef struct orc_tree {
    const double* xyz; /* borrowed, n x 3 */
    int64_t n;
    int max_pts, max_depth;
    orc_node* nodes;
    int64_t n_nodes, cap_nodes;
    int32_t* leaf_idx;
    int64_t n_leaf_idx;
    int32_t* scratch; /* partition scratch, n entries */
} orc_tree;

static int64_t orc_new_node(orc_tree* t, const double lo[3], const double hi[3]) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
        t->nodes = (orc_node*)realloc(t->nodes, (size_t)t->cap_nodes * sizeof(orc_node));
    }
    orc_node* nd = &t->nodes[t->n_nodes];
    for (int a = 0; a < 3; ++a) {
        nd->lo[a] = lo[a];
        nd->hi[a] = hi[a];
    }
    for (int c = 0; c < 8; ++c) nd->child[c] = -1;
    nd->first = 0;
    nd->count = 0;
    nd->is_leaf = 1;
    return t->n_nodes++;
}

/* octree.cpp:86-126.  `idx[0..cnt)` holds this node's point indices in ascending order; it is partitioned
 * stably (in place, through t->scratch) into the eight octants, so every child list stays ascending. */
static void orc_build_rec(orc_tree* t, int64_t me, int32_t* idx, int64_t cnt, int depth) {
    if (cnt <= (int64_t)t->max_pts || depth >= t->max_depth) { /* :88 */
        orc_node* nd = &t->nodes[me];
        nd->is_leaf = 1;
        nd->first = t->n_leaf_idx;
        nd->count = (int32_t)cnt;
        memcpy(t->leaf_idx + t->n_leaf_idx, idx, (size_t)cnt * sizeof(int32_t));
        t->n_leaf_idx += cnt;
        return;
    }
    double lo[3], hi[3], mid[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = t->nodes[me].lo[a];
        hi[a] = t->nodes[me].hi[a];
        mid[a] = (lo[a] + hi[a]) / 2; /* :97-99 */
    }
    t->nodes[me].is_leaf = 0;
    int64_t bucket[9];
    memset(bucket, 0, sizeof bucket);
    for (int64_t k = 0; k < cnt; ++k) {
        const double* p = t->xyz + 3 * (int64_t)idx[k];
        int oct = 0;
        if (p[0] > mid[0]) oct |= 1; /* :106-108, strict */
        if (p[1] > mid[1]) oct |= 2;
        if (p[2] > mid[2]) oct |= 4;
        bucket[oct + 1]++;
    }
    for (int c = 0; c < 8; ++c) bucket[c + 1] += bucket[c];
    {
        /* stable counting partition through the shared scratch (its use ends before any recursion) */
        int64_t pos[8];
        for (int c = 0; c < 8; ++c) pos[c] = bucket[c];
        int32_t* tmp = t->scratch;
        for (int64_t k = 0; k < cnt; ++k) {
            const double* p = t->xyz + 3 * (int64_t)idx[k];
            int oct = 0;
            if (p[0] > mid[0]) oct |= 1;
            if (p[1] > mid[1]) oct |= 2;
            if (p[2] > mid[2]) oct |= 4;
            tmp[pos[oct]++] = idx[k];
        }
        memcpy(idx, tmp, (size_t)cnt * sizeof(int32_t));
    }
    for (int c = 0; c < 8; ++c) { /* :113-125, only non-empty octants get a node */
        int64_t ccnt = bucket[c + 1] - bucket[c];
        if (ccnt == 0) continue;
        double clo[3], chi[3];
        for (int a = 0; a < 3; ++a) {
            if ((c >> a) & 1) {
                clo[a] = mid[a];
                chi[a] = hi[a];
            } else {
                clo[a] = lo[a];
                chi[a] = mid[a];
            }
        }
        int64_t ch = orc_new_node(t, clo, chi);
        t->nodes[me].child[c] = (int32_t)ch;
        orc_build_rec(t, ch, idx + bucket[c], ccnt, depth + 1);
    }
}

orc_tree* orc_octree_build(const double* xyz, int64_t n, int max_pts, int max_depth) {
    orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
    t->xyz = xyz;
    t->n = n;
    t->max_pts = max_pts;
    t->max_depth = max_depth;
    if (n <= 0) return t; /* octree.cpp:45, empty cloud -> no root */
    double lo[3], hi[3];
    for (int a = 0; a < 3; ++a) lo[a] = hi[a] = xyz[a]; /* :47-49 */
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) { /* :51-58, strict comparisons (NaN never replaces) */
            double v = xyz[3 * i + a];
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    const double eps = 0.001; /* :61-64 */
    for (int a = 0; a < 3; ++a) {
        lo[a] -= eps;
        hi[a] += eps;
    }
    t->leaf_idx = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    int32_t* all = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    t->scratch = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) all[i] = (int32_t)i; /* :70-73 */
    int64_t root = orc_new_node(t, lo, hi);
    orc_build_rec(t, root, all, n, 0);
    free(all);
    free(t->scratch);
    t->scratch = NULL;
    return t;
}

void orc_octree_free(orc_tree* t) {
    if (!t) return;
    free(t->nodes);
    free(t->leaf_idx);
    free(t);
}

/* octree.cpp:32-38 -- note the sqrt: pruning and child ordering both use the ROOTED distance. */
static double orc_box_dist(const orc_node* nd, const double q[3]) {
    double d[3];
    for (int a = 0; a < 3; ++a) {
        double below = nd->lo[a] - q[a];
        double above = q[a] - nd->hi[a];
        double m = (below < above) ? above : below; /* std::max(a,b) = (a<b)?b:a */
        d[a] = (0.0 < m) ? m : 0.0;
    }
    return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

typedef struct {
    int64_t box_tests, leaves, point_tests;
} orc_counters;

/* octree.cpp:128-173 */
static void orc_search_rec(const orc_tree* t, int64_t me, const double q[3], int32_t* best_idx, double* best_d2,
                           orc_counters* ctr) {
    const orc_node* nd = &t->nodes[me];
    double md = orc_box_dist(nd, q);
    if (ctr) ctr->box_tests++;
    if (md * md >= *best_d2) return; /* :134-135, rooted then re-squared */
    if (nd->is_leaf) {
        if (ctr) ctr->leaves++;
        for (int32_t k = 0; k < nd->count; ++k) { /* :139-150, ascending index, strict < */
            int32_t idx = t->leaf_idx[nd->first + k];
            const double* p = t->xyz + 3 * (int64_t)idx;
            double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
            double d2 = dx * dx + dy * dy + dz * dz;
            if (ctr) ctr->point_tests++;
            if (d2 < *best_d2) {
                *best_d2 = d2;
                *best_idx = idx;
            }
        }
        return;
    }
    /* :152-171: existing children in octant order with their rooted box distance, sorted ascending by that
     * distance.  std::sort on <= 8 elements is an insertion sort that moves an element left only past strictly
     * greater ones, i.e. it is stable: equal distances keep octant order. */
    int order[8];
    double dist[8];
    int m = 0;
    for (int c = 0; c < 8; ++c) {
        if (nd->child[c] < 0) continue;
        double dc = orc_box_dist(&t->nodes[nd->child[c]], q);
        if (ctr) ctr->box_tests++;
        int pos = m++;
        while (pos > 0 && dc < dist[pos - 1]) {
            dist[pos] = dist[pos - 1];
            order[pos] = order[pos - 1];
            --pos;
        }
        dist[pos] = dc;
        order[pos] = c;
    }
    for (int k = 0; k < m; ++k) orc_search_rec(t, nd->child[order[k]], q, best_idx, best_d2, ctr);
}

/* octree.cpp:175-184.  init_best = DBL_MAX (engine) or 1e20 (CLI, icp_registration.cpp:201). */
static int32_t orc_find_one(const orc_tree* t, const double q[3], double init_best, orc_counters* ctr) {
    if (t->n_nodes == 0 || t->n == 0) return 0;
    int32_t best_idx = 0;
    double best = init_best;
    /* the child loop re-tests each child on entry exactly as the reference's recursion does; the extra
     * box test has no effect on the result, so it is counted but not repeated here */
    orc_search_rec(t, 0, q, &best_idx, &best, ctr);
    return best_idx;
}

typedef struct {
    const orc_tree* t;
    const double* q;
    int64_t nq;
    int32_t* out;
    double init_best;
    atomic_llong* next;
} orc_nn_job;

static void* orc_nn_worker(void* arg) {
    orc_nn_job* j = (orc_nn_job*)arg;
    const long long chunk = 256;
    for (;;) {
        long long b = atomic_fetch_add(j->next, chunk);
        if (b >= j->nq) break;
        long long e = b + chunk < j->nq ? b + chunk : j->nq;
        for (long long i = b; i < e; ++i) j->out[i] = orc_find_one(j->t, j->q + 3 * i, j->init_best, NULL);
    }
    return NULL;
}

void orc_octree_find_nearest(const orc_tree* t, const double* q, int64_t nq, int32_t* idx_out, int variant, int nthreads) {
    double init_best = (variant == ORC_VARIANT_CLI) ? 1e20 : DBL_MAX;
    if (nthreads > 1) {
        atomic_llong next;
        atomic_init(&next, 0);
        orc_nn_job job = {t, q, nq, idx_out, init_best, &next};
        pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
        for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, orc_nn_worker, &job);
        for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
        free(th);
        return;
    }
    for (int64_t i = 0; i < nq; ++i) idx_out[i] = orc_find_one(t, q + 3 * i, init_best, NULL);
}

/* Traversal work counters (SURVEY.md section 6 "traversal work per query"), summed over the queries. */
void orc_octree_count_work(const orc_tree* t, const double* q, int64_t nq, int variant, int64_t* box_tests,
                           int64_t* leaves, int64_t* point_tests) {
    double init_best = (variant == ORC_VARIANT_CLI) ? 1e20 : DBL_MAX;
    orc_counters c = {0, 0, 0};
    for (int64_t i = 0; i < nq; ++i) (void)orc_find_one(t, q + 3 * i, init_best, &c);
    *box_tests = c.box_tests;
    *leaves = c.leaves;
    *point_tests = c.point_tests;
}

/* Pre-order dump in the same format as oracle/ref_engine_wrap.cpp:ref_octree_dump. */
static void orc_dump_rec(const orc_tree* t, int64_t me, int depth, uint64_t key, int64_t* n_nodes, int64_t* n_idx,
                         int32_t* o_depth, uint64_t* o_key, uint8_t* o_leaf, int32_t* o_count, double* o_box,
                         int32_t* o_idx) {
    const orc_node* nd = &t->nodes[me];
    int64_t slot = (*n_nodes)++;
    if (o_depth) {
        o_depth[slot] = depth;
        o_key[slot] = key;
        o_leaf[slot] = (uint8_t)(nd->is_leaf ? 1 : 0);
        o_count[slot] = nd->is_leaf ? nd->count : 0;
        double* b = o_box + 6 * slot;
        b[0] = nd->lo[0]; b[1] = nd->hi[0]; b[2] = nd->lo[1]; b[3] = nd->hi[1]; b[4] = nd->lo[2]; b[5] = nd->hi[2];
    }
    if (nd->is_leaf) {
        for (int32_t k = 0; k < nd->count; ++k) {
            if (o_idx) o_idx[*n_idx] = t->leaf_idx[nd->first + k];
            (*n_idx)++;
        }
        return;
    }
    for (int c = 0; c < 8; ++c)
        if (nd->child[c] >= 0)
            orc_dump_rec(t, nd->child[c], depth + 1, (key << 3) | (uint64_t)c, n_nodes, n_idx, o_depth, o_key, o_leaf,
                         o_count, o_box, o_idx);
}

int64_t orc_octree_dump(const orc_tree* t, int64_t* n_leaf_pts, int32_t* o_depth, uint64_t* o_key, uint8_t* o_leaf,
                        int32_t* o_count, double* o_box, int32_t* o_idx) {
    int64_t n_nodes = 0, n_idx = 0;
    if (t->n_nodes > 0) orc_dump_rec(t, 0, 0, 0, &n_nodes, &n_idx, o_depth, o_key, o_leaf, o_count, o_box, o_idx);
    if (n_leaf_pts) *n_leaf_pts = n_idx;
    return n_nodes;
}

/* ------------------------------------------------------------------------------------------------------
 * 3x3 two-sided Jacobi SVD, restating Eigen 3.3.4 JacobiSVD<Matrix3d>::compute for a real square matrix
 * (no QR preconditioner is run when rows == cols, JacobiSVD.h:683-695).  Row-major 3x3 arrays.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    double c, s;
} orc_rot;

/* Jacobi.h:428-440 (scalar path): x' = c x + s y ; y' = -s x + c y ; skipped when (c,s) == (1,0) (:308). */
static void orc_rot_apply(double* x, double* y, int n, int stride, orc_rot j) {
    if (j.c == 1.0 && j.s == 0.0) return;
    for (int i = 0; i < n; ++i) {
        double xi = x[i * stride], yi = y[i * stride];
        x[i * stride] = j.c * xi + j.s * yi;
        y[i * stride] = -j.s * xi + j.c * yi;
    }
}

/* Jacobi.h:83-114 */
static orc_rot orc_make_jacobi(double x, double y, double z) {
    orc_rot r;
    double deno = 2.0 * fabs(y);
    if (deno < DBL_MIN) {
        r.c = 1.0;
        r.s = 0.0;
        return r;
    }
    double tau = (x - z) / deno;
    double w = sqrt(tau * tau + 1.0);
    double t;
    if (tau > 0.0)
        t = 1.0 / (tau + w);
    else
        t = 1.0 / (tau - w);
    double sign_t = t > 0.0 ? 1.0 : -1.0;
    double n = 1.0 / sqrt(t * t + 1.0);
    r.s = -sign_t * (y / fabs(y)) * fabs(t) * n;
    r.c = n;
    return r;
}

/* RealSvd2x2.h:19-50 on the (p,q) sub-block of W (row-major 3x3). */
static void orc_svd2x2(const double* W, int p, int q, orc_rot* j_left, orc_rot* j_right) {
    double m[4] = {W[3 * p + p], W[3 * p + q], W[3 * q + p], W[3 * q + q]};
    orc_rot rot1;
    double t = m[0] + m[3];
    double d = m[2] - m[1];
    if (fabs(d) < DBL_MIN) {
        rot1.s = 0.0;
        rot1.c = 1.0;
    } else {
        double u = t / d;
        double tmp = sqrt(1.0 + u * u);
        rot1.s = 1.0 / tmp;
        rot1.c = u / tmp;
    }
    orc_rot_apply(&m[0], &m[2], 2, 1, rot1); /* m.applyOnTheLeft(0,1,rot1): rows 0 and 1 */
    *j_right = orc_make_jacobi(m[0], m[1], m[3]);
    /* *j_left = rot1 * j_right->transpose()  (Jacobi.h:50-56 with other = (c, -s)) */
    orc_rot o = {j_right->c, -j_right->s};
    j_left->c = rot1.c * o.c - rot1.s * o.s;
    j_left->s = rot1.c * o.s + rot1.s * o.c;
}

void orc_svd3(const double* H, double* U, double* S, double* V) {
    const double precision = 2.0 * DBL_EPSILON; /* JacobiSVD.h:672 */
    const double consider_as_zero = DBL_MIN;    /* :675 */
    double W[9];
    double scale = 0.0; /* :678-679 */
    for (int i = 0; i < 9; ++i) {
        double a = fabs(H[i]);
        if (a > scale) scale = a;
    }
    if (scale == 0.0) scale = 1.0;
    for (int i = 0; i < 9; ++i) W[i] = H[i] / scale; /* :691 */
    for (int i = 0; i < 9; ++i) U[i] = V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    double max_diag = 0.0; /* :699; maxCoeff keeps the first of equal maxima, NaN never wins */
    max_diag = fabs(W[0]);
    if (fabs(W[4]) > max_diag) max_diag = fabs(W[4]);
    if (fabs(W[8]) > max_diag) max_diag = fabs(W[8]);
    int finished = 0;
    while (!finished) { /* :702-737 */
        finished = 1;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                double pm = precision * max_diag; /* :713, numext::maxi(a,b) = a<b ? b : a */
                double thr = (consider_as_zero < pm) ? pm : consider_as_zero;
                if (fabs(W[3 * p + q]) > thr || fabs(W[3 * q + p]) > thr) {
                    finished = 0;
                    orc_rot jl, jr;
                    orc_svd2x2(W, p, q, &jl, &jr);
                    orc_rot_apply(&W[3 * p], &W[3 * q], 3, 1, jl);   /* W.applyOnTheLeft(p,q,j_left): rows */
                    orc_rot_apply(&U[p], &U[q], 3, 3, jl);           /* U.applyOnTheRight(p,q,j_left^T): cols, rotation j_left */
                    orc_rot jrt = {jr.c, -jr.s};
                    orc_rot_apply(&W[p], &W[q], 3, 3, jrt);          /* W.applyOnTheRight(p,q,j_right): cols, rotation j_right^T */
                    orc_rot_apply(&V[p], &V[q], 3, 3, jrt);          /* V.applyOnTheRight(p,q,j_right) */
                    double a = fabs(W[3 * p + p]), b = fabs(W[3 * q + q]);
                    double mx = a < b ? b : a;
                    max_diag = max_diag < mx ? mx : max_diag; /* :732 */
                }
            }
    }
    for (int i = 0; i < 3; ++i) { /* :741-759 */
        double a = W[3 * i + i];
        S[i] = fabs(a);
        if (a < 0.0)
            for (int r = 0; r < 3; ++r) U[3 * r + i] = -U[3 * r + i];
    }
    for (int i = 0; i < 3; ++i) S[i] *= scale; /* :761 */
    for (int i = 0; i < 3; ++i) {              /* :765-782 */
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

/* Order in which Eigen 3.3.4 sums a length-3 inner product of fixed-size operands when built for baseline
 * x86-64 (SSE2 packets of two doubles), as the reference is: rows 0 and 1 of a 3x3 * 3x3 or 3x3 * 3x1
 * product go through the packet path and accumulate left to right, row 2 is the scalar remainder and goes
 * through the unrolled reduction a0 + (a1 + a2).  Established against oracle/_ref element by element
 * (tests/test_oracle_vs_ref.py::test_solve_from_H_bit_exact). */
static double orc_dot3(int row, double a0, double b0, double a1, double b1, double a2, double b2) {
    if (row < 2) return (a0 * b0 + a1 * b1) + a2 * b2;
    return a0 * b0 + (a1 * b1 + a2 * b2);
}

/* icpengine.cpp:93-112 (and the CLI's :414-437, which negates row 2 of V^T == column 2 of V):
 * R = V U^T, reflection fix, t = cB - R cA, 4x4 row-major T. */
void orc_solve_from_H(const double* H, const double* cA, const double* cB, double* T) {
    double U[9], S[3], V[9], R[9];
    orc_svd3(H, U, S, V);
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                R[3 * i + j] = orc_dot3(i, V[3 * i + 0], U[3 * j + 0], V[3 * i + 1], U[3 * j + 1], V[3 * i + 2], U[3 * j + 2]);
        if (pass == 1) break;
        /* Eigen determinant_impl<.,3>: det = h(0,1,2) - h(1,0,2) + h(2,0,1), h(a,b,c) = m(0,a)*(m(1,b)*m(2,c) - m(1,c)*m(2,b)) */
        double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (!(det < 0.0)) break;
        for (int r = 0; r < 3; ++r) V[3 * r + 2] *= -1.0;
    }
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = cB[i] - orc_dot3(i, R[3 * i + 0], cA[0], R[3 * i + 1], cA[1], R[3 * i + 2], cA[2]);
    }
}

/* icpengine.cpp:82-90: centroids (sequential sum / n) and H = sum (a - cA)(b - cB)^T, sequential in pair order. */
void orc_centroids_H(const double* a, const double* b, int64_t n, double* cA, double* cB, double* H) {
    double sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) {
            sa[r] += a[3 * i + r];
            sb[r] += b[3 * i + r];
        }
    for (int r = 0; r < 3; ++r) {
        cA[r] = sa[r] / (double)n;
        cB[r] = sb[r] / (double)n;
    }
    for (int i = 0; i < 9; ++i) H[i] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double da[3], db[3];
        for (int r = 0; r < 3; ++r) {
            da[r] = a[3 * i + r] - cA[r];
            db[r] = b[3 * i + r] - cB[r];
        }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) H[3 * r + c] += da[r] * db[c];
    }
}

void orc_kabsch(const double* a, const double* b, int64_t n, double* T) {
    double cA[3], cB[3], H[9];
    orc_centroids_H(a, b, n, cA, cB, H);
    orc_solve_from_H(H, cA, cB, T);
}

/* icpengine.cpp:345: src = T * src on homogeneous columns; element r = ((T(r,0)x + T(r,1)y) + T(r,2)z) + T(r,3)*1. */
void orc_apply(const double* T, double* xyz, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        for (int r = 0; r < 3; ++r) xyz[3 * i + r] = ((T[4 * r] * x + T[4 * r + 1] * y) + T[4 * r + 2] * z) + T[4 * r + 3] * 1.0;
    }
}

/* icpengine.cpp:342: T_cum = T * T_cum (fixed-size 4x4 product, sequential inner sum). */
void orc_mat4_mul(const double* A, const double* B, double* C) {
    double out[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            out[4 * i + j] = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
    memcpy(C, out, sizeof out);
}

/* icpengine.cpp:356-362 */
void orc_angles(const double* T, double* angle_deg, double* trans_dist) {
    double trace = T[0] + (T[5] + T[10]);
    *angle_deg = acos((trace - 1.0) / 2.0) * 180.0 / M_PI;
    /* Block<Matrix4d,3,1>::norm(): vectorised redux, packet (t0,t1) first, then the scalar tail -- pinned by
     * tests/golden/engine_*.npz (the trace above is the unrolled scalar redux a0 + (a1 + a2)) */
    *trans_dist = sqrt((T[3] * T[3] + T[7] * T[7]) + T[11] * T[11]);
}

/* ------------------------------------------------------------------------------------------------------
 * One iteration's statistics on given correspondences (icpengine.cpp:187-278 ; CLI :499-541).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    double min_distance, max_distance; /* over finite distances (engine only) */
    double mean, std_dev, threshold, rmse, sum_sq;
    int64_t problem_count, valid_count, outlier_count;
} orc_stats;

void orc_iteration_stats(const double* src, int64_t n, const double* tgt, int64_t m, const int32_t* idx, int iter,
                         double sigma, int variant, double* dist, uint8_t* mask, orc_stats* st) {
    double mn = DBL_MAX, mx = 0.0;
    int64_t problems = 0;
    for (int64_t i = 0; i < n; ++i) {
        int32_t j = idx[i];
        if (variant == ORC_VARIANT_ENGINE && (j < 0 || (int64_t)j >= m)) { /* :199-204 */
            problems++;
            dist[i] = DBL_MAX;
            continue;
        }
        double dx = src[3 * i] - tgt[3 * (int64_t)j];
        double dy = src[3 * i + 1] - tgt[3 * (int64_t)j + 1];
        double dz = src[3 * i + 2] - tgt[3 * (int64_t)j + 2];
        dist[i] = sqrt(dx * dx + dy * dy + dz * dz); /* :68-74 */
        if (variant == ORC_VARIANT_ENGINE) {
            if (isnan(dist[i]) || isinf(dist[i])) problems++; /* :208-218 */
            if (isfinite(dist[i])) {                            /* :220-223 */
                if (dist[i] < mn) mn = dist[i];
                if (dist[i] > mx) mx = dist[i];
            }
        }
    }
    double mean = 0.0;
    for (int64_t i = 0; i < n; ++i) mean += dist[i]; /* :235-239 */
    mean /= (double)n;
    double var = 0.0;
    for (int64_t i = 0; i < n; ++i) var += (dist[i] - mean) * (dist[i] - mean); /* :241-244 */
    double sd = sqrt(var / (double)n);
    double thr;
    if (variant == ORC_VARIANT_ENGINE && iter == 0) { /* :249-252 */
        double a = sigma * sd, b = mean * 0.5;
        thr = mean + (a < b ? b : a);
    } else {
        thr = mean + sigma * sd; /* :254 ; CLI :523 (sigma = 3.0) */
    }
    int64_t valid = 0;
    double sum_sq = 0.0;
    for (int64_t i = 0; i < n; ++i) { /* :263-274 */
        int ok = dist[i] <= thr;
        mask[i] = (uint8_t)ok;
        if (ok) {
            valid++;
            sum_sq += dist[i] * dist[i];
        }
    }
    st->min_distance = mn;
    st->max_distance = mx;
    st->mean = mean;
    st->std_dev = sd;
    st->threshold = thr;
    st->sum_sq = sum_sq;
    st->rmse = valid > 0 ? sqrt(sum_sq / (double)valid) : 0.0; /* :274 */
    st->problem_count = problems;
    st->valid_count = valid;
    st->outlier_count = n - valid;
}

/* ------------------------------------------------------------------------------------------------------
 * The whole loop (icpengine.cpp:117-394 ; CLI icp_registration.cpp:443-622).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t max_iterations;
    int32_t octree_max_points;
    int32_t octree_max_depth;
    int32_t variant;
    double tolerance;
    double sigma_multiplier;
} orc_params;

typedef struct {
    int32_t iteration, valid_points, outlier_points, has_angles;
    double rmse;
    double transform[16]; /* cumulative, row-major */
    double rotation_angle, translation_distance;
} orc_iter;

#define ORC_OK 0
#define ORC_EMPTY_INPUT 1
#define ORC_CANCELLED 2
#define ORC_TOO_FEW_INLIERS 3

typedef struct {
    int32_t status, success, total_iterations, loop_iterations; /* loop_iterations = NN passes executed */
    int32_t history_len, pad_;                                  /* records pushed (kept even on failure exits) */
    double final_rmse;
    double final_R[9], final_t[3];
    double last_T[16], cum_T[16];
} orc_result;

/* Optional per-iteration trace for parity tests: for loop iteration k < trace_iters the arrays receive
 * idx[k*n..], dist[k*n..], mask[k*n..] and stats[k]. */
typedef struct {
    int32_t trace_iters;
    int32_t* idx;
    double* dist;
    uint8_t* mask;
    orc_stats* stats;
    double* src_before; /* source coordinates entering iteration k, n*3 each */
} orc_trace;

int orc_icp_run(double* src_xyz, int64_t n, const double* tgt_xyz, int64_t m, const orc_params* p, int stop_after,
                orc_result* res, orc_iter* hist, int hist_cap, orc_trace* tr, int nthreads) {
    memset(res, 0, sizeof *res);
    for (int i = 0; i < 16; ++i) res->last_T[i] = res->cum_T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (!src_xyz || !tgt_xyz || n <= 0 || m <= 0) { /* icpengine.cpp:26-34 */
        res->status = ORC_EMPTY_INPUT;
        return res->status;
    }
    int variant = p->variant;
    int leaf = variant == ORC_VARIANT_CLI ? 10 : p->octree_max_points; /* CLI hard-codes 10/20/3.0, :454,:523 */
    int depth = variant == ORC_VARIANT_CLI ? 20 : p->octree_max_depth;
    double sigma = variant == ORC_VARIANT_CLI ? 3.0 : p->sigma_multiplier;
    orc_tree* tree = orc_octree_build(tgt_xyz, m, leaf, depth);
    double* cur = (double*)malloc((size_t)n * 3 * sizeof(double));
    memcpy(cur, src_xyz, (size_t)n * 3 * sizeof(double));
    int32_t* idx = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    double* dist = (double*)malloc((size_t)n * sizeof(double));
    uint8_t* mask = (uint8_t*)malloc((size_t)n);
    double* va = (double*)malloc((size_t)n * 3 * sizeof(double));
    double* vb = (double*)malloc((size_t)n * 3 * sizeof(double));
    double T[16], Tc[16];
    for (int i = 0; i < 16; ++i) T[i] = Tc[i] = (i % 5 == 0) ? 1.0 : 0.0;
    double prev_error = 1e10; /* :156-157 */
    int no_improve = 0;
    int n_hist = 0, write_back = 1, signals = 0;
    double last_rmse = 0.0;
    res->status = ORC_OK;
    for (int iter = 0; iter < p->max_iterations; ++iter) {
        if (variant == ORC_VARIANT_ENGINE && stop_after >= 0 && signals >= stop_after && signals > 0) { /* :160-164 */
            res->status = ORC_CANCELLED;
            write_back = 0;
            break;
        }
        res->loop_iterations = iter + 1;
        if (tr && iter < tr->trace_iters && tr->src_before) memcpy(tr->src_before + (size_t)iter * n * 3, cur, (size_t)n * 3 * sizeof(double));
        orc_octree_find_nearest(tree, cur, n, idx, variant, nthreads); /* :172-184 */
        orc_stats st;
        orc_iteration_stats(cur, n, tgt_xyz, m, idx, iter, sigma, variant, dist, mask, &st);
        if (tr && iter < tr->trace_iters) {
            if (tr->idx) memcpy(tr->idx + (size_t)iter * n, idx, (size_t)n * sizeof(int32_t));
            if (tr->dist) memcpy(tr->dist + (size_t)iter * n, dist, (size_t)n * sizeof(double));
            if (tr->mask) memcpy(tr->mask + (size_t)iter * n, mask, (size_t)n);
            if (tr->stats) tr->stats[iter] = st;
        }
        double rmse = st.rmse;
        double improvement = prev_error - rmse; /* :288 */
        if (fabs(improvement) < p->tolerance) {
            no_improve++;
            if (no_improve >= 3) {
                if (variant == ORC_VARIANT_ENGINE) { /* :294-303: extra record, angle fields unset */
                    if (n_hist < hist_cap) {
                        orc_iter* h = &hist[n_hist];
                        memset(h, 0, sizeof *h);
                        h->iteration = iter + 1;
                        h->rmse = rmse;
                        h->valid_points = (int32_t)st.valid_count;
                        h->outlier_points = (int32_t)st.outlier_count;
                        memcpy(h->transform, Tc, sizeof Tc);
                        h->has_angles = 0;
                    }
                    n_hist++;
                    signals++;
                    last_rmse = rmse;
                }
                break;
            }
        } else {
            no_improve = 0;
        }
        if (rmse > prev_error * 1.1) break; /* :311-314 */
        prev_error = rmse;                  /* :316 */
        if (st.valid_count < 3) {           /* :319-323 ; CLI :567-570 just breaks */
            if (variant == ORC_VARIANT_ENGINE) {
                res->status = ORC_TOO_FEW_INLIERS;
                write_back = 0;
            }
            break;
        }
        int64_t nv = 0;
        for (int64_t i = 0; i < n; ++i) { /* :325-337 */
            if (!mask[i]) continue;
            for (int r = 0; r < 3; ++r) {
                va[3 * nv + r] = cur[3 * i + r];
                vb[3 * nv + r] = tgt_xyz[3 * (int64_t)idx[i] + r];
            }
            nv++;
        }
        orc_kabsch(va, vb, nv, T);   /* :339 */
        orc_mat4_mul(T, Tc, Tc);     /* :342 */
        orc_apply(T, cur, n);        /* :345-346 */
        if (n_hist < hist_cap) {     /* :349-362 ; CLI :593-595 keeps only the transform */
            orc_iter* h = &hist[n_hist];
            memset(h, 0, sizeof *h);
            h->iteration = iter + 1;
            h->rmse = rmse;
            h->valid_points = (int32_t)st.valid_count;
            h->outlier_points = (int32_t)st.outlier_count;
            memcpy(h->transform, Tc, sizeof Tc);
            orc_angles(Tc, &h->rotation_angle, &h->translation_distance);
            h->has_angles = 1;
        }
        n_hist++;
        signals++;
        last_rmse = rmse;
    }
    if (write_back) { /* :371-375 */
        memcpy(src_xyz, cur, (size_t)n * 3 * sizeof(double));
        res->success = 1;
        res->total_iterations = n_hist; /* :386 */
        res->final_rmse = n_hist ? last_rmse : 0.0;
        const double* F = variant == ORC_VARIANT_CLI ? T : Tc; /* CLI returns the LAST incremental T, :616-621 */
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) res->final_R[3 * i + j] = F[4 * i + j];
            res->final_t[i] = F[4 * i + 3];
        }
        if (variant == ORC_VARIANT_CLI) res->final_rmse = prev_error; /* :606 prints prev_error */
    }
    res->history_len = n_hist;
    memcpy(res->last_T, T, sizeof T);
    memcpy(res->cum_T, Tc, sizeof Tc);
    free(cur); free(idx); free(dist); free(mask); free(va); free(vb);
    orc_octree_free(tree);
    return res->status;
}

int orc_hw_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
