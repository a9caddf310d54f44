// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// extern "C" wrapper around the UNMODIFIED reference data-format code next to the ICP loop:
// PointCloudRegistration/core/lasio.cpp (LASIO::readLAS / writeLAS / readLASBatch) and core/pointcloud.cpp
// (PointCloud::computeBounds / applyTransform / downsample), compiled in place (against oracle/qt_shim) into
// oracle/_ref/libref_io.so.  It pins oracle/cloudio_oracle.c and produces tests/golden/io_*.npz.
#include <cstdint>
#include <iostream>
#include <sstream>
#include <vector>
#include "lasio.h"
#include "pointcloud.h"

namespace {
struct Quiet {  // the reference prints progress to stdout / stderr
    std::streambuf *o, *e;
    std::ostringstream sink;
    Quiet() : o(std::cout.rdbuf(sink.rdbuf())), e(std::cerr.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(o); std::cerr.rdbuf(e); }
};
void fill(PointCloud& c, const double* xyz, int64_t n) {
    c.points.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) c.points[(size_t)i] = Point3D(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
int64_t spill(const std::vector<Point3D>& p, double* out, int64_t cap) {
    for (size_t i = 0; i < p.size() && (int64_t)i < cap; ++i) {
        out[3 * i] = p[i].x; out[3 * i + 1] = p[i].y; out[3 * i + 2] = p[i].z;
    }
    return (int64_t)p.size();
}
}  // namespace

extern "C" {

// computeBounds() then LASIO::writeLAS(filename, cloud) (lasio.cpp:126-209); returns the bool.
int ref_io_write_las(const char* filename, const double* xyz, int64_t n) {
    Quiet q;
    PointCloud c;
    fill(c, xyz, n);
    c.computeBounds();
    return LASIO::writeLAS(filename, c) ? 1 : 0;
}

// LASIO::readLAS(filename, cloud, maxPoints) (lasio.cpp:6-125): returns the point count (or -1 on failure), writes up to
// cap points and the cloud's bounds (min xyz, max xyz).
int64_t ref_io_read_las(const char* filename, int64_t max_points, double* xyz_out, int64_t cap, double* bounds6) {
    Quiet q;
    PointCloud c;
    if (!LASIO::readLAS(filename, c, (size_t)max_points)) return -1;
    if (bounds6) {
        bounds6[0] = c.minX; bounds6[1] = c.minY; bounds6[2] = c.minZ;
        bounds6[3] = c.maxX; bounds6[4] = c.maxY; bounds6[5] = c.maxZ;
    }
    return spill(c.points, xyz_out, cap);
}

// LASIO::readLASBatch (lasio.cpp:211-300): concatenated points and the size of every batch handed to the callback.
int64_t ref_io_read_las_batch(const char* filename, int64_t batch_size, double* xyz_out, int64_t cap, int64_t* batch_sizes,
                              int64_t batch_cap, int64_t* n_batches) {
    Quiet q;
    int64_t at = 0, nb = 0;
    size_t total = LASIO::readLASBatch(filename, (size_t)batch_size, [&](const std::vector<Point3D>& b) {
        for (const auto& p : b) {
            if (at < cap) { xyz_out[3 * at] = p.x; xyz_out[3 * at + 1] = p.y; xyz_out[3 * at + 2] = p.z; }
            ++at;
        }
        if (nb < batch_cap) batch_sizes[nb] = (int64_t)b.size();
        ++nb;
    });
    *n_batches = nb;
    return (int64_t)total;
}

void ref_io_bounds(const double* xyz, int64_t n, double* min3, double* max3) {
    PointCloud c;
    fill(c, xyz, n);
    c.computeBounds();
    min3[0] = c.minX; min3[1] = c.minY; min3[2] = c.minZ;
    max3[0] = c.maxX; max3[1] = c.maxY; max3[2] = c.maxZ;
}

// PointCloud::downsample(targetSize) (pointcloud.cpp:107-128): -1 for the nullptr exits.
int64_t ref_io_downsample(const double* xyz, int64_t n, int target, double* out, int64_t cap) {
    PointCloud c;
    fill(c, xyz, n);
    PointCloud* s = c.downsample(target);
    if (!s) return -1;
    const int64_t m = spill(s->points, out, cap);
    delete s;
    return m;
}

// PointCloud::applyTransform(R, t) (pointcloud.cpp:73-86), in place.
void ref_io_apply_transform(const double* R9, const double* t3, double* xyz, int64_t n) {
    PointCloud c;
    fill(c, xyz, n);
    double R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = R9[3 * i + j];
    c.applyTransform(R, t3);
    spill(c.points, xyz, n);
}

}  // extern "C"
