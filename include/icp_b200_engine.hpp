// icp_b200_engine.hpp -- header-only C++17 adapters over the C ABI (icp_b200.h) that keep the reference's entry
// points, so its callers compile against libicp_b200.so by changing one include:
//
//   icpb200::ICPEngine      same methods as the reference's ICPEngine (PointCloudRegistration/core/icpengine.h:51-87):
//                           setParameters / getParameters / registerPointClouds / stop / getResult; the five Qt signals
//                           (icpengine.h:70-75) become std::function members that fire on the calling thread in the
//                           reference's order (log -> iterationCompleted -> progressUpdated, icpengine.cpp:364-367).
//   icpb200::ICP(...)       the CLI's ICP() (icp_registration.cpp:443-446), same argument list.
//   icpb200::Octree         Octree(points, max_pts, max_d) / findNearest (core/octree.h:27-43).
//
// `Cloud` is any type with a public `std::vector<P> points` whose P is three consecutive doubles x, y, z -- both the
// engine's PointCloud (core/pointcloud.h:30-65, Point3D at :12-23) and the CLI's (icp_registration.cpp:16-21) qualify.
// 4x4 matrices are returned as row-major std::array<double,16>; with Eigen available use
//   Eigen::Map<const Eigen::Matrix<double,4,4,Eigen::RowMajor>>(m.data()).
// No CPU fallback: without a CUDA device the constructors throw std::runtime_error.
#pragma once
#include <array>
#include <atomic>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "icp_b200.h"

namespace icpb200 {

using Mat4 = std::array<double, 16>;  // row-major

struct ICPParameters {  // core/icpengine.h:13-19
    int maxIterations = 50;
    double tolerance = 1e-6;
    double sigmaMultiplier = 3.0;
    int octreeMaxPoints = 10;
    int octreeMaxDepth = 20;
};

struct IterationResult {  // core/icpengine.h:24-32
    int iteration = 0;
    double rmse = 0.0;
    int validPoints = 0;
    int outlierPoints = 0;
    Mat4 transform{};            // cumulative
    double rotationAngle = 0.0;  // degrees; not set on the convergence record (the reference leaves it unset too)
    double translationDistance = 0.0;
};

struct ICPResult {  // core/icpengine.h:37-44
    bool success = false;
    int totalIterations = 0;
    double finalRMSE = 0.0;
    double finalR[3][3] = {};
    double finalT[3] = {};
    std::vector<IterationResult> iterationHistory;
};

namespace detail {
template <class P>
inline double* as_xyz(std::vector<P>& v) {
    static_assert(sizeof(P) == 3 * sizeof(double), "point type must be three packed doubles (x, y, z)");
    return v.empty() ? nullptr : reinterpret_cast<double*>(v.data());
}
template <class P>
inline const double* as_xyz(const std::vector<P>& v) {
    static_assert(sizeof(P) == 3 * sizeof(double), "point type must be three packed doubles (x, y, z)");
    return v.empty() ? nullptr : reinterpret_cast<const double*>(v.data());
}
inline IterationResult from_c(const icp_iteration& it) {
    IterationResult r;
    r.iteration = it.iteration;
    r.rmse = it.rmse;
    r.validPoints = it.valid_points;
    r.outlierPoints = it.outlier_points;
    for (int k = 0; k < 16; ++k) r.transform[k] = it.transform[k];
    if (it.has_angles) {
        r.rotationAngle = it.rotation_angle;
        r.translationDistance = it.translation_distance;
    }
    return r;
}
}  // namespace detail

class ICPEngine {
public:
    // replacements for the Qt signals (core/icpengine.h:70-75)
    std::function<void()> started;
    std::function<void(int, int, double)> progressUpdated;
    std::function<void(const IterationResult&)> iterationCompleted;
    std::function<void(bool, const std::string&)> finished;
    std::function<void(const std::string&)> logMessage;

    explicit ICPEngine(int device = 0) {
        if (icp_create(&h_, device) != ICP_OK) throw std::runtime_error("icp_b200: no usable CUDA device (there is no CPU fallback)");
        icp_set_callbacks(h_, &ICPEngine::on_iteration, &ICPEngine::on_progress, &ICPEngine::on_log, this);
    }
    ~ICPEngine() { icp_destroy(h_); }
    ICPEngine(const ICPEngine&) = delete;
    ICPEngine& operator=(const ICPEngine&) = delete;

    void setParameters(const ICPParameters& p) { params_ = p; }  // icpengine.cpp:19-22
    ICPParameters getParameters() const { return params_; }
    void stop() {                                                // icpengine.cpp:62-66 (atomic here)
        stop_.store(1);
        if (logMessage) logMessage(u8"用户请求停止配准...");
    }
    ICPResult getResult() const { return result_; }
    icp_handle handle() const { return h_; }

    // icpengine.cpp:24-60.  `source` is updated in place exactly where the reference writes it back.
    template <class Cloud>
    void registerPointClouds(Cloud* source, const Cloud* target) {
        if (!source || !target) return emit_finished(false, u8"源点云或目标点云为空");
        if (source->points.empty() || target->points.empty()) return emit_finished(false, u8"点云数据为空");
        stop_.store(0);
        result_ = ICPResult();
        if (started) started();
        icp_params cp;
        icp_default_params(&cp);
        cp.max_iterations = params_.maxIterations;
        cp.tolerance = params_.tolerance;
        cp.sigma_multiplier = params_.sigmaMultiplier;
        cp.octree_max_points = params_.octreeMaxPoints;
        cp.octree_max_depth = params_.octreeMaxDepth;
        cp.variant = ICP_VARIANT_ENGINE;
        if (icp_set_params(h_, &cp) != ICP_OK) return emit_finished(false, icp_last_error(h_));
        std::vector<icp_iteration> hist((size_t)params_.maxIterations + 2);
        icp_result res{};
        res.history = hist.data();
        res.history_cap = (int32_t)hist.size();
        const int st = icp_register(h_, detail::as_xyz(source->points), (int64_t)source->points.size(),
                                    detail::as_xyz(target->points), (int64_t)target->points.size(), &res,
                                    reinterpret_cast<const volatile int*>(&stop_));
        result_.success = res.success != 0;
        result_.totalIterations = res.total_iterations;
        result_.finalRMSE = res.final_rmse;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) result_.finalR[i][j] = res.final_R[3 * i + j];
            result_.finalT[i] = res.final_t[i];
        }
        for (int k = 0; k < res.history_len; ++k) result_.iterationHistory.push_back(detail::from_c(hist[(size_t)k]));
        switch (st) {
            case ICP_OK: return emit_finished(true, u8"配准成功");
            case ICP_CANCELLED: return emit_finished(false, u8"用户取消");
            case ICP_TOO_FEW_INLIERS: return emit_finished(false, u8"有效点对不足");
            case ICP_EMPTY_INPUT: return emit_finished(false, u8"点云数据为空");
            default: return emit_finished(false, icp_last_error(h_));
        }
    }

private:
    static void on_iteration(const icp_iteration* it, void* u) {
        auto* self = static_cast<ICPEngine*>(u);
        if (self->iterationCompleted) self->iterationCompleted(detail::from_c(*it));
    }
    static void on_progress(int i, int n, double rmse, void* u) {
        auto* self = static_cast<ICPEngine*>(u);
        if (self->progressUpdated) self->progressUpdated(i, n, rmse);
    }
    static void on_log(const char* m, void* u) {
        auto* self = static_cast<ICPEngine*>(u);
        if (self->logMessage) self->logMessage(m);
    }
    void emit_finished(bool ok, const std::string& msg) {
        if (finished) finished(ok, msg);
    }
    icp_handle h_ = nullptr;
    ICPParameters params_;
    ICPResult result_;
    std::atomic<int> stop_{0};
    static_assert(sizeof(std::atomic<int>) == sizeof(int), "stop flag is read by the library as a plain int");
};

// The CLI's ICP() (icp_registration.cpp:443-446): `source` is moved in place, final_R / final_t receive the LAST
// incremental transform (:616-621), iteration_transforms the cumulative transform of every iteration (:593-595).
template <class Cloud>
inline void ICP(Cloud& source, const Cloud& target, int max_iterations, double tolerance, double final_R[3][3],
                double final_t[3], std::vector<Mat4>* iteration_transforms = nullptr, int device = 0) {
    icp_handle h = nullptr;
    if (icp_create(&h, device) != ICP_OK) throw std::runtime_error("icp_b200: no usable CUDA device (there is no CPU fallback)");
    icp_params cp;
    icp_default_params(&cp);
    cp.max_iterations = max_iterations;
    cp.tolerance = tolerance;
    cp.variant = ICP_VARIANT_CLI;
    icp_set_params(h, &cp);
    std::vector<icp_iteration> hist((size_t)max_iterations + 2);
    icp_result res{};
    res.history = hist.data();
    res.history_cap = (int32_t)hist.size();
    icp_register(h, detail::as_xyz(source.points), (int64_t)source.points.size(), detail::as_xyz(target.points),
                 (int64_t)target.points.size(), &res, nullptr);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) final_R[i][j] = res.final_R[3 * i + j];
        final_t[i] = res.final_t[i];
    }
    if (iteration_transforms) {
        iteration_transforms->clear();
        for (int k = 0; k < res.history_len; ++k) {
            Mat4 m;
            for (int e = 0; e < 16; ++e) m[(size_t)e] = hist[(size_t)k].transform[e];
            iteration_transforms->push_back(m);
        }
    }
    icp_destroy(h);
}

// Octree(points, max_pts, max_d) / findNearest (core/octree.h:27-43).  findNearest of many points at once is the
// loop of core/icpengine.cpp:172-184.
class Octree {
public:
    template <class P>
    explicit Octree(const std::vector<P>& pts, int max_pts = 10, int max_d = 20, int device = 0) {
        if (icp_create(&h_, device) != ICP_OK) throw std::runtime_error("icp_b200: no usable CUDA device (there is no CPU fallback)");
        empty_ = pts.empty();
        if (!empty_ && icp_octree_build(h_, detail::as_xyz(pts), (int64_t)pts.size(), max_pts, max_d) != ICP_OK) {
            const std::string e = icp_last_error(h_);
            icp_destroy(h_);
            throw std::runtime_error("icp_b200: octree build failed: " + e);
        }
    }
    ~Octree() { icp_destroy(h_); }
    Octree(const Octree&) = delete;
    Octree& operator=(const Octree&) = delete;

    template <class P>
    int findNearest(const P& query) const {
        static_assert(sizeof(P) == 3 * sizeof(double), "point type must be three packed doubles (x, y, z)");
        if (empty_) return 0;  // octree.cpp:177
        int32_t idx = 0;
        icp_nn_query(h_, reinterpret_cast<const double*>(&query), 1, &idx, nullptr, nullptr);
        return idx;
    }
    template <class P>
    std::vector<int32_t> findNearest(const std::vector<P>& queries) const {
        std::vector<int32_t> idx(queries.size(), 0);
        if (!empty_ && !queries.empty())
            icp_nn_query(h_, detail::as_xyz(queries), (int64_t)queries.size(), idx.data(), nullptr, nullptr);
        return idx;
    }

private:
    icp_handle h_ = nullptr;
    bool empty_ = true;
};

}  // namespace icpb200
