/* TEST INFRASTRUCTURE ONLY -- the product path (iterativeclosestpoint_b200/csrc, libicp_b200.so) never includes,
 * links or calls this file; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load liboracle.so.
 *
 * Plain-C CPU restatement of the data-format steps either side of the ICP loop:
 *   LAS 1.2 point decode          PointCloudRegistration/core/lasio.cpp:38-48,92-99   (CLI twin icp_registration.cpp:282-359)
 *   LAS 1.2 writer (engine)       core/lasio.cpp:126-209      header fields :140-184, truncating cast :193-195
 *   LAS 1.2 writer (CLI)          icp_registration.cpp:698-815  header :706-777, truncating cast :785-787
 *   PointCloud::computeBounds     core/pointcloud.cpp:24-45
 *   PointCloud::downsample        core/pointcloud.cpp:107-128    (index = (int)(i * step), step = size / target)
 *   CLI stride sampling           icp_registration.cpp:877-882   (every sample_rate-th point)
 *   PointCloud::applyTransform    core/pointcloud.cpp:73-86      (viewer replay, widgets/pointcloudviewer.cpp:86-116)
 *   saveTransformation            icp_registration.cpp:625-695   (text report, ostream precision 10 == "%.10g")
 *
 * Parity pinning: the reference has no tests for these either; the functions below are pinned against the UNMODIFIED
 * reference sources compiled into oracle/_ref/ (libref_io.so: core/lasio.cpp + core/pointcloud.cpp; libref_cli.so: the
 * CLI's readLASFile / saveResultAsLAS / saveTransformation) by tests/test_cloudio_oracle.py and through the vectors that
 * build wrote to tests/golden/io_*.npz (tools/make_golden_io.py).  Compile with -O2 -ffp-contract=off.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_LAS_HEADER 227
#define ORC_LAS_RECORD 20

/* static_cast<int32_t>(double) as x86-64 evaluates it (cvttsd2si): truncation toward zero, 0x80000000 when the
 * value is NaN or does not fit.  The C standard leaves that case undefined, so it is spelled out. */
static int32_t orc_trunc_i32(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN;
    return (int32_t)v;
}

/* lasio.cpp:92-99 / icp_registration.cpp:351-359: first three int32 of each record, p = raw * scale + offset. */
void orc_las_decode(const uint8_t* records, int64_t n, int32_t record_length, const double* scale, const double* offset,
                    double* xyz_out) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t raw[3];
        memcpy(raw, records + i * (int64_t)record_length, 12);
        for (int a = 0; a < 3; ++a) xyz_out[3 * i + a] = raw[a] * scale[a] + offset[a];
    }
}

/* lasio.cpp:192-204 / icp_registration.cpp:783-810: (int32)((p - offset) / scale), then 8 zero bytes. */
void orc_las_encode(const double* xyz, int64_t n, const double* scale, const double* offset, uint8_t* records_out) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t raw[3];
        for (int a = 0; a < 3; ++a) raw[a] = orc_trunc_i32((xyz[3 * i + a] - offset[a]) / scale[a]);
        memcpy(records_out + i * ORC_LAS_RECORD, raw, 12);
        memset(records_out + i * ORC_LAS_RECORD + 12, 0, 8);
    }
}

/* pointcloud.cpp:24-45: std::min / std::max from +-DBL_MAX; all zero for an empty cloud. */
void orc_bounds(const double* xyz, int64_t n, double* min3, double* max3) {
    if (n <= 0) {
        for (int a = 0; a < 3; ++a) min3[a] = max3[a] = 0.0;
        return;
    }
    for (int a = 0; a < 3; ++a) {
        min3[a] = DBL_MAX;
        max3[a] = -DBL_MAX;
    }
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            const double v = xyz[3 * i + a];
            if (v < min3[a]) min3[a] = v; /* std::min(m, v) = (v < m) ? v : m */
            if (max3[a] < v) max3[a] = v; /* std::max(m, v) = (m < v) ? v : m */
        }
}

/* icp_registration.cpp:750-762: bounds seeded from point 0, strict comparisons (the CLI writer's header). */
static void orc_bounds_cli(const double* xyz, int64_t n, double* min3, double* max3) {
    for (int a = 0; a < 3; ++a) min3[a] = max3[a] = xyz[a];
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            const double v = xyz[3 * i + a];
            if (v < min3[a]) min3[a] = v;
            if (v > max3[a]) max3[a] = v;
        }
}

static void put_u16(uint8_t* h, int at, uint16_t v) { memcpy(h + at, &v, 2); }
static void put_u32(uint8_t* h, int at, uint32_t v) { memcpy(h + at, &v, 4); }
static void put_f64(uint8_t* h, int at, double v) { memcpy(h + at, &v, 8); }

/* The 227-byte header either writer emits.  variant 0: LASIO::writeLAS (scale 0.001, offset = cloud minimum, lasio.cpp:
 * 140-184); variant 1: saveResultAsLAS (system id / software / date fields, the cloud's own scale and offset,
 * icp_registration.cpp:706-777).  min3/max3 are the bounds the header records. */
void orc_las_header(int variant, int64_t n, const double* scale, const double* offset, const double* min3, const double* max3,
                    uint8_t* header227) {
    uint8_t* h = header227;
    memset(h, 0, ORC_LAS_HEADER);
    memcpy(h, "LASF", 4);
    h[24] = 1;
    h[25] = 2;
    if (variant == 1) {
        memcpy(h + 26, "ICP Registration", 16);
        memcpy(h + 58, "Custom ICP", 10);
        put_u16(h, 90, 307);
        put_u16(h, 92, 2025);
    }
    put_u16(h, 94, ORC_LAS_HEADER);
    put_u32(h, 96, ORC_LAS_HEADER);
    h[104] = 0;
    put_u16(h, 105, ORC_LAS_RECORD);
    put_u32(h, 107, (uint32_t)n);
    for (int a = 0; a < 3; ++a) {
        put_f64(h, 131 + 8 * a, scale[a]);
        put_f64(h, 155 + 8 * a, offset[a]);
        put_f64(h, 179 + 16 * a, max3[a]);
        put_f64(h, 187 + 16 * a, min3[a]);
    }
}

/* Whole file image (header + records) of LASIO::writeLAS (variant 0) or saveResultAsLAS (variant 1, scale/offset given).
 * out must hold 227 + 20 n bytes.  Returns the byte count, 0 for an empty cloud (lasio.cpp:128-131). */
int64_t orc_las_file_image(int variant, const double* xyz, int64_t n, const double* scale_in, const double* offset_in,
                           uint8_t* out) {
    double mn[3], mx[3], scale[3], offset[3];
    if (n <= 0) return 0;
    if (variant == 0) {
        orc_bounds(xyz, n, mn, mx);
        for (int a = 0; a < 3; ++a) {
            scale[a] = 0.001;
            offset[a] = mn[a];
        }
    } else {
        orc_bounds_cli(xyz, n, mn, mx);
        for (int a = 0; a < 3; ++a) {
            scale[a] = scale_in[a];
            offset[a] = offset_in[a];
        }
    }
    orc_las_header(variant, n, scale, offset, mn, mx, out);
    orc_las_encode(xyz, n, scale, offset, out + ORC_LAS_HEADER);
    return ORC_LAS_HEADER + ORC_LAS_RECORD * n;
}

/* Header fields both readers use (lasio.cpp:38-48): returns 0 on a bad signature. */
int orc_las_parse_header(const uint8_t* header227, uint32_t* offset_to_data, uint32_t* n_points, uint16_t* record_length,
                         double* scale, double* offset) {
    if (memcmp(header227, "LASF", 4) != 0) return 0;
    memcpy(offset_to_data, header227 + 96, 4);
    memcpy(record_length, header227 + 105, 2);
    memcpy(n_points, header227 + 107, 4);
    for (int a = 0; a < 3; ++a) {
        memcpy(scale + a, header227 + 131 + 8 * a, 8);
        memcpy(offset + a, header227 + 155 + 8 * a, 8);
    }
    return 1;
}

/* pointcloud.cpp:107-128.  Returns the number of points written (0 for an empty cloud or target <= 0). */
int64_t orc_downsample(const double* xyz, int64_t n, int32_t target, double* out) {
    if (n <= 0 || target <= 0) return 0;
    if ((int32_t)n <= target) {
        memcpy(out, xyz, (size_t)n * 24);
        return n;
    }
    const double step = (double)n / target;
    for (int32_t i = 0; i < target; ++i) {
        const int32_t idx = (int32_t)(i * step);
        memcpy(out + 3 * (int64_t)i, xyz + 3 * (int64_t)idx, 24);
    }
    return target;
}

/* icp_registration.cpp:877-882: for (i = 0; i < size; i += sample_rate). */
int64_t orc_downsample_stride(const double* xyz, int64_t n, int64_t stride, double* out) {
    int64_t k = 0;
    if (stride <= 0) return 0;
    for (int64_t i = 0; i < n; i += stride, ++k) memcpy(out + 3 * k, xyz + 3 * i, 24);
    return k;
}

/* pointcloud.cpp:73-86 with R, t taken from a row-major 4x4 as the viewer does (pointcloudviewer.cpp:100-110). */
void orc_cloud_apply(const double* T16, const double* xyz_in, int64_t n, double* xyz_out) {
    for (int64_t i = 0; i < n; ++i) {
        const double x = xyz_in[3 * i], y = xyz_in[3 * i + 1], z = xyz_in[3 * i + 2];
        for (int r = 0; r < 3; ++r) xyz_out[3 * i + r] = T16[4 * r] * x + T16[4 * r + 1] * y + T16[4 * r + 2] * z + T16[4 * r + 3];
    }
}

/* icp_registration.cpp:625-695.  iter_T: n_iter row-major 4x4 cumulative transforms (may be NULL / 0).  Writes the text
 * into buf (cap bytes) and returns its length (or the length needed when it does not fit). */
int64_t orc_transformation_text(const double* R9, const double* t3, const double* iter_T, int32_t n_iter, char* buf, int64_t cap) {
    int64_t len = 0;
#define EMIT(...)                                                                              \
    do {                                                                                       \
        int w__ = snprintf(len < cap ? buf + len : NULL, len < cap ? (size_t)(cap - len) : 0, __VA_ARGS__); \
        len += w__;                                                                            \
    } while (0)
    EMIT("ICP配准变换参数\n==================\n\n");
    EMIT("说明: 将源点云变换到目标点云坐标系下的变换矩阵\n");
    EMIT("变换公式: P_target = R * P_source + t\n\n");
    if (iter_T && n_iter > 0) {
        EMIT("==================\n迭代过程变换参数\n==================\n\n");
        for (int32_t k = 0; k < n_iter; ++k) {
            const double* T = iter_T + 16 * (int64_t)k;
            EMIT("--- 迭代 %d ---\n旋转矩阵 R:\n", k + 1);
            for (int i = 0; i < 3; ++i) EMIT("  [%.10g, %.10g, %.10g]\n", T[4 * i], T[4 * i + 1], T[4 * i + 2]);
            EMIT("平移向量 t:\n  [%.10g, %.10g, %.10g]\n\n", T[3], T[7], T[11]);
        }
        EMIT("\n");
    }
    EMIT("==================\n最终变换参数\n==================\n\n");
    EMIT("旋转矩阵 R (3x3):\n");
    for (int i = 0; i < 3; ++i) EMIT("  [%.10g, %.10g, %.10g]\n", R9[3 * i], R9[3 * i + 1], R9[3 * i + 2]);
    EMIT("\n平移向量 t (3x1):\n  [%.10g, %.10g, %.10g]\n", t3[0], t3[1], t3[2]);
    EMIT("\n变换矩阵 (齐次坐标形式 4x4):\n");
    for (int i = 0; i < 3; ++i) EMIT("  [%.10g, %.10g, %.10g, %.10g]\n", R9[3 * i], R9[3 * i + 1], R9[3 * i + 2], t3[i]);
    EMIT("  [0, 0, 0, 1]\n");
#undef EMIT
    return len;
}
