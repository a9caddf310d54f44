// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// extern "C" wrapper around the UNMODIFIED reference engine sources, compiled in place from
// /root/reference/PointCloudRegistration/core/{octree,pointcloud,icpengine}.cpp against the
// tiny Qt stand-in headers in oracle/qt_shim/ (no Qt in this image).  The output
// (oracle/_ref/libref_engine.so) is git-ignored; no reference source is copied into the repo.
//
// This file supplies (a) the bodies of the five Qt signals that moc would have generated
// (core/icpengine.h:70-75) and (b) plain-C entry points so Python tests (ctypes) can drive
// Octree::findNearest (core/octree.cpp:175-184), ICPEngine::registerPointClouds
// (core/icpengine.cpp:24-60 -> runICP :117-394), ICPEngine::computeBestFitTransform (:76-115)
// and Eigen::JacobiSVD<Matrix3d>.
#include <vector>
#include <string>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <limits>
#include <numeric>
#include "Eigen/Eigen"

// The reference keeps the tree root and the Kabsch solve private; the wrapper needs to read the
// former (tree dump for structure parity) and call the latter.  Access control only.
#define private public
#include "octree.h"
#include "icpengine.h"
#undef private

#include <thread>
#include <atomic>

// ---------------------------------------------------------------------------------------------
// Signal bodies (what moc would emit).  They record into a per-thread capture block.
// ---------------------------------------------------------------------------------------------
namespace {
struct Capture {
    std::vector<IterationResult> iters;
    std::vector<std::string> logs;
    std::vector<int> progress_iter;
    std::vector<double> progress_rmse;
    bool got_started = false;
    bool got_finished = false;
    bool finished_ok = false;
    std::string finished_msg;
    int stop_after = -1;       // call engine->stop() after this many iterationCompleted signals
    ICPEngine* engine = nullptr;
    std::vector<std::string> order;  // sequence of signal names (callback-order parity)
};
thread_local Capture* g_cap = nullptr;
}  // namespace

void ICPEngine::started() { if (g_cap) { g_cap->got_started = true; g_cap->order.push_back("started"); } }
void ICPEngine::progressUpdated(int iteration, int total, double rmse) {
    (void)total;
    if (!g_cap) return;
    g_cap->progress_iter.push_back(iteration);
    g_cap->progress_rmse.push_back(rmse);
    g_cap->order.push_back("progress");
}
void ICPEngine::iterationCompleted(const IterationResult& r) {
    if (!g_cap) return;
    g_cap->iters.push_back(r);
    g_cap->order.push_back("iteration");
    if (g_cap->stop_after >= 0 && (int)g_cap->iters.size() >= g_cap->stop_after && g_cap->engine)
        g_cap->engine->stop();
}
void ICPEngine::finished(bool success, const QString& message) {
    if (!g_cap) return;
    g_cap->got_finished = true;
    g_cap->finished_ok = success;
    g_cap->finished_msg = message.str();
    g_cap->order.push_back("finished");
}
void ICPEngine::logMessage(const QString& message) {
    if (!g_cap) return;
    g_cap->logs.push_back(message.str());
}

// ---------------------------------------------------------------------------------------------
extern "C" {

struct ref_octree {
    std::vector<Point3D> pts;
    Octree* tree;
};

void* ref_octree_create(const double* xyz, int64_t n, int max_pts, int max_depth) {
    ref_octree* h = new ref_octree();
    h->pts.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) h->pts[(size_t)i] = Point3D(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    h->tree = new Octree(h->pts, max_pts, max_depth);
    return h;
}

void ref_octree_destroy(void* hv) {
    ref_octree* h = (ref_octree*)hv;
    if (!h) return;
    delete h->tree;
    delete h;
}

// nthreads <= 1: the reference as shipped (serial).  nthreads > 1: std::thread workers AROUND the
// reference's own const findNearest (read-only => thread-safe), pulling 256-query chunks off an atomic
// counter -- "reference x T cores".
void ref_octree_find_nearest(void* hv, const double* q, int64_t nq, int32_t* out, int nthreads) {
    ref_octree* h = (ref_octree*)hv;
    const Octree* t = h->tree;
    if (nthreads > 1) {
        std::atomic<int64_t> next(0);
        const int64_t chunk = 256;
        auto work = [&]() {
            for (;;) {
                int64_t b = next.fetch_add(chunk);
                if (b >= nq) break;
                int64_t e = std::min(nq, b + chunk);
                for (int64_t i = b; i < e; ++i) {
                    Point3D p(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
                    out[i] = t->findNearest(p);
                }
            }
        };
        std::vector<std::thread> th;
        for (int k = 0; k < nthreads; ++k) th.emplace_back(work);
        for (auto& x : th) x.join();
        return;
    }
    for (int64_t i = 0; i < nq; ++i) {
        Point3D p(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        out[i] = t->findNearest(p);
    }
}

int ref_max_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

// Pre-order dump of the pointer tree (children visited 0..7) for structure parity:
//   per node: depth, octant path key (3 bits/level, root split most significant), is_leaf, point count,
//   box {min_x,max_x,min_y,max_y,min_z,max_z}; leaf point indices concatenated in node order.
// Call with null outputs to size the arrays; returns node count, *n_leaf_pts gets index count.
static void dump_rec(const OctreeNode* nd, int depth, uint64_t key, int64_t* n_nodes, int64_t* n_idx,
                     int32_t* o_depth, uint64_t* o_key, uint8_t* o_leaf, int32_t* o_count, double* o_box,
                     int32_t* o_idx) {
    int64_t me = (*n_nodes)++;
    if (o_depth) {
        o_depth[me] = depth;
        o_key[me] = key;
        o_leaf[me] = nd->is_leaf ? 1 : 0;
        o_count[me] = (int32_t)nd->point_indices.size();
        double* b = o_box + 6 * me;
        b[0] = nd->min_x; b[1] = nd->max_x; b[2] = nd->min_y; b[3] = nd->max_y; b[4] = nd->min_z; b[5] = nd->max_z;
    }
    if (nd->is_leaf) {
        for (int idx : nd->point_indices) {
            if (o_idx) o_idx[*n_idx] = idx;
            (*n_idx)++;
        }
        return;
    }
    for (int c = 0; c < 8; ++c)
        if (nd->children[c])
            dump_rec(nd->children[c], depth + 1, (key << 3) | (uint64_t)c, n_nodes, n_idx, o_depth, o_key, o_leaf,
                     o_count, o_box, o_idx);
}

int64_t ref_octree_dump(void* hv, int64_t* n_leaf_pts, int32_t* o_depth, uint64_t* o_key, uint8_t* o_leaf,
                        int32_t* o_count, double* o_box, int32_t* o_idx) {
    ref_octree* h = (ref_octree*)hv;
    int64_t n_nodes = 0, n_idx = 0;
    if (h->tree->root) dump_rec(h->tree->root, 0, 0, &n_nodes, &n_idx, o_depth, o_key, o_leaf, o_count, o_box, o_idx);
    if (n_leaf_pts) *n_leaf_pts = n_idx;
    return n_nodes;
}

struct ref_iter {
    int32_t iteration;
    int32_t validPoints;
    int32_t outlierPoints;
    int32_t has_angles;  // 0 for the convergence record, whose angle fields the reference leaves unset
    double rmse;
    double transform[16];  // row-major
    double rotationAngle;
    double translationDistance;
};

struct ref_result {
    int32_t success;
    int32_t totalIterations;
    int32_t got_finished;
    int32_t finished_ok;
    int32_t n_iter_signals;
    int32_t n_logs;
    double finalRMSE;
    double finalR[9];
    double finalT[3];
    char finished_msg[128];
    char signal_order[4096];  // compact: s=started i=iteration p=progress f=finished
};

static std::vector<std::string> g_last_logs;

// Runs ICPEngine::registerPointClouds on copies of the inputs held in reference PointCloud objects;
// src_xyz receives the (possibly updated) source points afterwards.  A null src/tgt pointer passes a
// null PointCloud* to the engine; n == 0 passes an empty cloud.
int ref_engine_run(double* src_xyz, int64_t n_src, const double* tgt_xyz, int64_t n_tgt, int max_iterations,
                   double tolerance, double sigma_multiplier, int octree_max_points, int octree_max_depth,
                   int stop_after, ref_result* out, ref_iter* hist, int hist_cap, int print_logs) {
    PointCloud src, tgt;
    if (src_xyz) {
        src.points.resize((size_t)n_src);
        for (int64_t i = 0; i < n_src; ++i) src.points[(size_t)i] = Point3D(src_xyz[3 * i], src_xyz[3 * i + 1], src_xyz[3 * i + 2]);
    }
    if (tgt_xyz) {
        tgt.points.resize((size_t)n_tgt);
        for (int64_t i = 0; i < n_tgt; ++i) tgt.points[(size_t)i] = Point3D(tgt_xyz[3 * i], tgt_xyz[3 * i + 1], tgt_xyz[3 * i + 2]);
    }
    ICPEngine engine;
    ICPParameters p;
    p.maxIterations = max_iterations;
    p.tolerance = tolerance;
    p.sigmaMultiplier = sigma_multiplier;
    p.octreeMaxPoints = octree_max_points;
    p.octreeMaxDepth = octree_max_depth;
    engine.setParameters(p);

    Capture cap;
    cap.stop_after = stop_after;
    cap.engine = &engine;
    g_cap = &cap;
    if (stop_after == 0) {
        // stop() before the run is reset by registerPointClouds (m_shouldStop=false, icpengine.cpp:38);
        // the earliest observable cancel is therefore after the first iteration signal.
    }
    engine.registerPointClouds(src_xyz ? &src : nullptr, tgt_xyz ? &tgt : nullptr);
    g_cap = nullptr;

    std::memset(out, 0, sizeof(*out));
    out->got_finished = cap.got_finished;
    out->finished_ok = cap.finished_ok;
    out->n_iter_signals = (int)cap.iters.size();
    out->n_logs = (int)cap.logs.size();
    std::strncpy(out->finished_msg, cap.finished_msg.c_str(), sizeof(out->finished_msg) - 1);
    {
        size_t k = 0;
        for (const std::string& s : cap.order) {
            if (k + 1 >= sizeof(out->signal_order)) break;
            out->signal_order[k++] = s[0];
        }
        out->signal_order[k] = 0;
    }
    bool ran = cap.got_started;
    if (ran) {
        ICPResult r = engine.getResult();
        // success/totalIterations/finalRMSE/finalR/finalT are only assigned on the exits that reach
        // icpengine.cpp:371-393; on the early-return exits they keep whatever ICPResult() held.
        out->success = (cap.got_finished && cap.finished_ok) ? (r.success ? 1 : 0) : 0;
        if (cap.got_finished && cap.finished_ok) {
            out->totalIterations = r.totalIterations;
            out->finalRMSE = r.finalRMSE;
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) out->finalR[3 * i + j] = r.finalR[i][j];
                out->finalT[i] = r.finalT[i];
            }
        }
        int n = (int)r.iterationHistory.size();
        for (int k = 0; k < n && k < hist_cap; ++k) {
            const IterationResult& it = r.iterationHistory[(size_t)k];
            ref_iter& o = hist[k];
            o.iteration = it.iteration;
            o.validPoints = it.validPoints;
            o.outlierPoints = it.outlierPoints;
            o.rmse = it.rmse;
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) o.transform[4 * i + j] = it.transform(i, j);
            o.rotationAngle = it.rotationAngle;
            o.translationDistance = it.translationDistance;
            o.has_angles = 1;
        }
        // The convergence record is the one that repeats the previous cumulative transform and is the
        // last entry after a "收敛达到" log line; flag it so callers ignore its unset angle fields.
        bool converged = false;
        for (const std::string& s : cap.logs)
            if (s.find("\xE6\x94\xB6\xE6\x95\x9B\xE8\xBE\xBE\xE5\x88\xB0") != std::string::npos) converged = true;
        if (converged && n > 0 && n <= hist_cap) hist[n - 1].has_angles = 0;
        if (src_xyz && src.points.size() == (size_t)n_src)
            for (int64_t i = 0; i < n_src; ++i) {
                src_xyz[3 * i] = src.points[(size_t)i].x;
                src_xyz[3 * i + 1] = src.points[(size_t)i].y;
                src_xyz[3 * i + 2] = src.points[(size_t)i].z;
            }
    }
    if (print_logs)
        for (const std::string& s : cap.logs) std::printf("%s\n", s.c_str());
    g_last_logs = cap.logs;
    return ran ? 0 : 1;
}

// The logMessage texts of the last ref_engine_run, joined with '\n' (UTF-8).  Returns the number of bytes needed.
int64_t ref_engine_last_logs(char* buf, int64_t cap) {
    std::string all;
    for (const std::string& s : g_last_logs) {
        all += s;
        all += '\n';
    }
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(all.size(), (size_t)cap - 1);
        std::memcpy(buf, all.data(), n);
        buf[n] = 0;
    }
    return (int64_t)all.size() + 1;
}

// ICPEngine::computeBestFitTransform on n matched pairs (AoS n x 3 each); T_out row-major 4x4.
void ref_engine_kabsch(const double* a_xyz, const double* b_xyz, int64_t n, double* T_out) {
    Eigen::MatrixXd A(3, n), B(3, n);
    for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) {
            A(r, i) = a_xyz[3 * i + r];
            B(r, i) = b_xyz[3 * i + r];
        }
    ICPEngine engine;
    Eigen::Matrix4d T = engine.computeBestFitTransform(A, B);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T_out[4 * i + j] = T(i, j);
}

// Centroids and the cross-covariance exactly as computeBestFitTransform forms them (icpengine.cpp:82-90),
// exposed so tests can feed the SAME H to every SVD restatement.
void ref_engine_centroids_H(const double* a_xyz, const double* b_xyz, int64_t n, double* cA, double* cB, double* H_out) {
    Eigen::MatrixXd A(3, n), B(3, n);
    for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) {
            A(r, i) = a_xyz[3 * i + r];
            B(r, i) = b_xyz[3 * i + r];
        }
    Eigen::Vector3d centroid_A = A.rowwise().mean();
    Eigen::Vector3d centroid_B = B.rowwise().mean();
    Eigen::MatrixXd AA = A.colwise() - centroid_A;
    Eigen::MatrixXd BB = B.colwise() - centroid_B;
    Eigen::Matrix3d H = AA * BB.transpose();
    for (int i = 0; i < 3; ++i) {
        cA[i] = centroid_A(i);
        cB[i] = centroid_B(i);
        for (int j = 0; j < 3; ++j) H_out[3 * i + j] = H(i, j);
    }
}

// Everything after H in computeBestFitTransform (icpengine.cpp:93-112), on a caller-supplied H and centroids.
void ref_engine_solve_from_H(const double* H_in, const double* cA, const double* cB, double* T_out, double* U_out,
                             double* S_out, double* V_out) {
    Eigen::Matrix3d H;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) H(i, j) = H_in[3 * i + j];
    Eigen::JacobiSVD<Eigen::Matrix3d> svd(H, Eigen::ComputeFullU | Eigen::ComputeFullV);
    Eigen::Matrix3d U = svd.matrixU();
    Eigen::Matrix3d V = svd.matrixV();
    Eigen::Vector3d S = svd.singularValues();
    for (int i = 0; i < 3; ++i) {
        if (S_out) S_out[i] = S(i);
        for (int j = 0; j < 3; ++j) {
            if (U_out) U_out[3 * i + j] = U(i, j);
            if (V_out) V_out[3 * i + j] = V(i, j);
        }
    }
    Eigen::Matrix3d R = V * U.transpose();
    if (R.determinant() < 0) {
        V.col(2) *= -1;
        R = V * U.transpose();
    }
    Eigen::Vector3d ca(cA[0], cA[1], cA[2]), cb(cB[0], cB[1], cB[2]);
    Eigen::Vector3d t = cb - R * ca;
    Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
    T.block<3, 3>(0, 0) = R;
    T.block<3, 1>(0, 3) = t;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T_out[4 * i + j] = T(i, j);
}

// `src = T * src` on a 4 x N homogeneous MatrixXd and `T_cum = T * T_cum`, the two Eigen products of
// icpengine.cpp:342-346, so the restatement's scalar formula can be checked bit-for-bit.
void ref_engine_apply(const double* T_in, double* xyz, int64_t n) {
    Eigen::Matrix4d T;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T(i, j) = T_in[4 * i + j];
    Eigen::MatrixXd src = Eigen::MatrixXd::Ones(4, n);
    for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) src(r, i) = xyz[3 * i + r];
    src = T * src;
    Eigen::MatrixXd src3d = src.topRows(3);
    for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) xyz[3 * i + r] = src3d(r, i);
}

void ref_engine_mat4_mul(const double* A_in, const double* B_in, double* C_out) {
    Eigen::Matrix4d A, B;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            A(i, j) = A_in[4 * i + j];
            B(i, j) = B_in[4 * i + j];
        }
    Eigen::Matrix4d C = A * B;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) C_out[4 * i + j] = C(i, j);
}

// rotationAngle / translationDistance as icpengine.cpp:356-362 derives them from a cumulative transform.
void ref_engine_angles(const double* T_in, double* angle_deg, double* trans_dist) {
    Eigen::Matrix4d T;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T(i, j) = T_in[4 * i + j];
    Eigen::Matrix3d R = T.block<3, 3>(0, 0);
    Eigen::Vector3d t = T.block<3, 1>(0, 3);
    double trace = R.trace();
    *angle_deg = std::acos((trace - 1.0) / 2.0) * 180.0 / M_PI;
    *trans_dist = t.norm();
}

}  // extern "C"
