// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// extern "C" wrapper around the UNMODIFIED reference CLI program /root/reference/icp_registration.cpp
// (its own Octree :23-206, best_fit_transform :389-440, ICP :443-622), compiled in place into
// oracle/_ref/libref_cli.so.  The whole translation unit is pulled into a namespace (and its main()
// renamed) so that its Point3D/Octree/PointCloud classes cannot collide with the engine's.
#include <iostream>
#include <fstream>
#include <sstream>
#include <vector>
#include <cmath>
#include <limits>
#include <string>
#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <numeric>
#include "Eigen/Eigen"

#define main ref_cli_program_main
namespace refcli {
#include "icp_registration.cpp"
}
#undef main

extern "C" {

struct ref_cli_octree {
    std::vector<refcli::Point3D> pts;
    refcli::Octree* tree;
};

void* ref_cli_octree_create(const double* xyz, int64_t n, int max_pts, int max_depth) {
    ref_cli_octree* h = new ref_cli_octree();
    h->pts.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) h->pts[(size_t)i] = refcli::Point3D(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    h->tree = new refcli::Octree(h->pts, max_pts, max_depth);
    return h;
}

void ref_cli_octree_destroy(void* hv) {
    ref_cli_octree* h = (ref_cli_octree*)hv;
    if (!h) return;
    delete h->tree;
    delete h;
}

void ref_cli_octree_find_nearest(void* hv, const double* q, int64_t nq, int32_t* out) {
    ref_cli_octree* h = (ref_cli_octree*)hv;
    for (int64_t i = 0; i < nq; ++i) {
        refcli::Point3D p(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        out[i] = h->tree->findNearest(p);
    }
}

// ICP(source, target, max_iterations, tolerance, final_R, final_t, &iteration_transforms)
// (icp_registration.cpp:443-446).  Returns the number of per-iteration cumulative transforms; the first
// `cap` of them are written row-major to iter_T.  stdout chatter of the reference is swallowed unless
// print != 0.
int ref_cli_icp(double* src_xyz, int64_t n_src, const double* tgt_xyz, int64_t n_tgt, int max_iterations,
                double tolerance, double* final_R9, double* final_t3, double* iter_T, int cap, int print) {
    refcli::PointCloud src, tgt;
    src.points.resize((size_t)n_src);
    tgt.points.resize((size_t)n_tgt);
    for (int64_t i = 0; i < n_src; ++i) src.points[(size_t)i] = refcli::Point3D(src_xyz[3 * i], src_xyz[3 * i + 1], src_xyz[3 * i + 2]);
    for (int64_t i = 0; i < n_tgt; ++i) tgt.points[(size_t)i] = refcli::Point3D(tgt_xyz[3 * i], tgt_xyz[3 * i + 1], tgt_xyz[3 * i + 2]);
    double R[3][3], t[3];
    std::vector<Eigen::Matrix4d> its;
    std::streambuf* old = nullptr;
    std::ostringstream sink;
    if (!print) old = std::cout.rdbuf(sink.rdbuf());
    refcli::ICP(src, tgt, max_iterations, tolerance, R, t, &its);
    if (!print) std::cout.rdbuf(old);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) final_R9[3 * i + j] = R[i][j];
        final_t3[i] = t[i];
    }
    for (size_t k = 0; k < its.size() && (int)k < cap; ++k)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) iter_T[16 * k + 4 * i + j] = its[k](i, j);
    for (int64_t i = 0; i < n_src; ++i) {
        src_xyz[3 * i] = src.points[(size_t)i].x;
        src_xyz[3 * i + 1] = src.points[(size_t)i].y;
        src_xyz[3 * i + 2] = src.points[(size_t)i].z;
    }
    return (int)its.size();
}

// best_fit_transform(A /*N x 3*/, B /*N x 3*/) (icp_registration.cpp:389-440); T_out row-major.
void ref_cli_best_fit_transform(const double* a_xyz, const double* b_xyz, int64_t n, double* T_out) {
    Eigen::MatrixXd A(n, 3), B(n, 3);
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            A(i, c) = a_xyz[3 * i + c];
            B(i, c) = b_xyz[3 * i + c];
        }
    Eigen::Matrix4d T = refcli::best_fit_transform(A, B);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T_out[4 * i + j] = T(i, j);
}

// saveTransformation (icp_registration.cpp:625-695) text format, for the "next" row f1 adapter test.
void ref_cli_save_transformation(const double* R9, const double* t3, const double* iter_T, int n_iter,
                                 const char* filename) {
    double R[3][3], t[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i][j] = R9[3 * i + j];
        t[i] = t3[i];
    }
    std::vector<Eigen::Matrix4d> its;
    for (int k = 0; k < n_iter; ++k) {
        Eigen::Matrix4d T;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) T(i, j) = iter_T[16 * k + 4 * i + j];
        its.push_back(T);
    }
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    refcli::saveTransformation(R, t, std::string(filename), n_iter > 0 ? &its : nullptr);
    std::cout.rdbuf(old);
}

// readLASFile (icp_registration.cpp:248-378): returns the point count (-1 on failure), the file's scale / offset as the
// CLI keeps them on the cloud (:307-312), and up to cap points.
int64_t ref_cli_read_las(const char* filename, double* xyz_out, int64_t cap, double* scale3, double* offset3) {
    refcli::PointCloud c;
    std::streambuf *old = std::cout.rdbuf(), *olde = std::cerr.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    std::cerr.rdbuf(sink.rdbuf());
    const bool ok = refcli::readLASFile(std::string(filename), c);
    std::cout.rdbuf(old);
    std::cerr.rdbuf(olde);
    if (!ok) return -1;
    scale3[0] = c.x_scale; scale3[1] = c.y_scale; scale3[2] = c.z_scale;
    offset3[0] = c.x_offset; offset3[1] = c.y_offset; offset3[2] = c.z_offset;
    for (size_t i = 0; i < c.points.size() && (int64_t)i < cap; ++i) {
        xyz_out[3 * i] = c.points[i].x; xyz_out[3 * i + 1] = c.points[i].y; xyz_out[3 * i + 2] = c.points[i].z;
    }
    return (int64_t)c.points.size();
}

// saveResultAsLAS (icp_registration.cpp:698-815) with the cloud's scale / offset set as main() does (:864-875).
void ref_cli_save_las(const double* xyz, int64_t n, const double* scale3, const double* offset3, const char* filename) {
    refcli::PointCloud c;
    c.points.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) c.points[(size_t)i] = refcli::Point3D(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    c.x_scale = scale3[0]; c.y_scale = scale3[1]; c.z_scale = scale3[2];
    c.x_offset = offset3[0]; c.y_offset = offset3[1]; c.z_offset = offset3[2];
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    refcli::saveResultAsLAS(c, std::string(filename));
    std::cout.rdbuf(old);
}

}  // extern "C"
