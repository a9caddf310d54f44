// The data-format steps either side of the ICP loop, on the device:
//   LAS 1.2 point decode  raw int32 x scale + offset           PointCloudRegistration/core/lasio.cpp:92-99
//                                                              (CLI twin icp_registration.cpp:351-359)
//   LAS 1.2 point encode  (int32)((p - offset) / scale)        core/lasio.cpp:192-204 (CLI :783-810)
//   bounds                PointCloud::computeBounds            core/pointcloud.cpp:24-45
//   downsampling          PointCloud::downsample               core/pointcloud.cpp:107-128; CLI stride icp_registration.cpp:877-882
//   replay                copy + PointCloud::applyTransform    core/pointcloud.cpp:73-86, widgets/pointcloudviewer.cpp:86-116
// and the host-side file framing around them (227-byte header, text report).  All streaming, HBM-bound passes:
// decode 12 B in (20 B record) / 24 B out per point, encode 24 B in / 20 B out, downsample 24 B gathered / 24 B out.
// Same arithmetic as the reference: products and sums round separately (-fmad=false + explicit __dmul_rn / __dadd_rn),
// the division is a division, the cast truncates toward zero.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "internal.h"

namespace icpb {

struct Vec3 {
    double v[3];
};

// One thread per point.  Records are `rl` bytes apart; only the first 12 bytes (X, Y, Z as little-endian int32) are read.
// Byte loads when the stride is not a multiple of four (point formats 2 and 3 are 26 and 34 bytes).
__global__ void __launch_bounds__(256) las_decode_kernel(const uint8_t* __restrict__ rec, int64_t n, int rl, Vec3 scale, Vec3 offset,
                                                         double* __restrict__ xyz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = rec + i * (int64_t)rl;
    int32_t raw[3];
    if ((rl & 3) == 0) {
        const int32_t* q = reinterpret_cast<const int32_t*>(p);
        raw[0] = __ldg(q);
        raw[1] = __ldg(q + 1);
        raw[2] = __ldg(q + 2);
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const uint32_t b0 = __ldg(p + 4 * a), b1 = __ldg(p + 4 * a + 1), b2 = __ldg(p + 4 * a + 2), b3 = __ldg(p + 4 * a + 3);
            raw[a] = (int32_t)(b0 | (b1 << 8) | (b2 << 16) | (b3 << 24));
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) xyz[3 * i + a] = dadd(dmul((double)raw[a], scale.v[a]), offset.v[a]);  // x * x_scale + x_offset
}

// static_cast<int32_t>(double) as the reference's x86-64 build evaluates it: toward zero, 0x80000000 when out of range / NaN.
__device__ __forceinline__ int32_t trunc_i32(double v) {
    return (v > -2147483649.0 && v < 2147483648.0) ? __double2int_rz(v) : (int32_t)0x80000000;
}

__global__ void __launch_bounds__(256) las_encode_kernel(const double* __restrict__ xyz, int64_t n, Vec3 scale, Vec3 offset,
                                                         uint32_t* __restrict__ rec /* 5 words per point */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t* o = rec + 5 * i;
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = (uint32_t)trunc_i32(ddiv(dsub(xyz[3 * i + a], offset.v[a]), scale.v[a]));
    o[3] = 0u;  // intensity, flags, classification
    o[4] = 0u;  // scan angle, user data, point source id
}

// Per-block min / max of every axis; block partials are folded by the host (at most a few thousand values).
// mode 0: std::min / std::max from +-DBL_MAX (pointcloud.cpp:30-41); NaN coordinates never win either comparison.
__global__ void __launch_bounds__(256) bounds_kernel(const double* __restrict__ xyz, int64_t n, double* __restrict__ part /* 6 per block */) {
    __shared__ double sm[6][8];
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = xyz[3 * i + a];
            mn[a] = (v < mn[a]) ? v : mn[a];
            mx[a] = (mx[a] < v) ? v : mx[a];
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, mn[a], o), w = __shfl_xor_sync(0xffffffffu, mx[a], o);
            mn[a] = (u < mn[a]) ? u : mn[a];
            mx[a] = (mx[a] < w) ? w : mx[a];
        }
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            sm[a][wp] = mn[a];
            sm[3 + a][wp] = mx[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const bool is_min = threadIdx.x < 3;
        double r = sm[threadIdx.x][0];
        for (int k = 1; k < 8; ++k) {
            const double u = sm[threadIdx.x][k];
            r = is_min ? ((u < r) ? u : r) : ((r < u) ? u : r);
        }
        part[6 * blockIdx.x + threadIdx.x] = r;
    }
}

// out[i] = in[(int)(i * step)] (pointcloud.cpp:119-123) or in[i * stride] (icp_registration.cpp:877-879)
__global__ void __launch_bounds__(256) gather_step_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t n_out, double step,
                                                          int64_t stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int64_t j = stride > 0 ? i * stride : (int64_t)__double2int_rz(dmul((double)(int)i, step));
    out[3 * i] = in[3 * j];
    out[3 * i + 1] = in[3 * j + 1];
    out[3 * i + 2] = in[3 * j + 2];
}

static inline int nblk(int64_t n) { return (int)((n + 255) / 256); }

int las_decode_launch(Ctx* c, cudaStream_t st, const uint8_t* d_rec, int64_t n, int rl, const double* scale, const double* offset,
                      double* d_xyz) {
    if (n <= 0) return ICP_OK;
    Vec3 s, o;
    for (int a = 0; a < 3; ++a) {
        s.v[a] = scale[a];
        o.v[a] = offset[a];
    }
    las_decode_kernel<<<nblk(n), 256, 0, st>>>(d_rec, n, rl, s, o, d_xyz);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

static int bounds_device(Ctx* c, const double* d_xyz, int64_t n, double* mn, double* mx) {
    const int blocks = std::min(nblk(n), c->sm_count * 8);
    ICPB_TRY(devbuf_reserve(c, c->scratch0, (size_t)blocks * 6 * sizeof(double)));
    bounds_kernel<<<blocks, 256, 0, c->stream>>>(d_xyz, n, (double*)c->scratch0.p);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    std::vector<double> part((size_t)blocks * 6);
    ICPB_CUDA(c, cudaMemcpyAsync(part.data(), c->scratch0.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int a = 0; a < 3; ++a) {
        mn[a] = DBL_MAX;
        mx[a] = -DBL_MAX;
    }
    for (int b = 0; b < blocks; ++b)
        for (int a = 0; a < 3; ++a) {
            mn[a] = std::min(mn[a], part[(size_t)6 * b + a]);
            mx[a] = std::max(mx[a], part[(size_t)6 * b + 3 + a]);
        }
    return ICP_OK;
}

static int upload_xyz(Ctx* c, DevBuf& b, const double* xyz, int64_t n) {
    ICPB_TRY(devbuf_reserve(c, b, (size_t)n * 3 * sizeof(double)));
    ICPB_CUDA(c, cudaMemcpyAsync(b.p, xyz, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return ICP_OK;
}

// encode n device points into c->scratch2 (20 B records)
static int encode_device(Ctx* c, const double* d_xyz, int64_t n, const double* scale, const double* offset) {
    ICPB_TRY(devbuf_reserve(c, c->scratch2, (size_t)n * ICP_LAS_RECORD_BYTES));
    Vec3 s, o;
    for (int a = 0; a < 3; ++a) {
        s.v[a] = scale[a];
        o.v[a] = offset[a];
    }
    las_encode_kernel<<<nblk(n), 256, 0, c->stream>>>(d_xyz, n, s, o, (uint32_t*)c->scratch2.p);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

static void put16(uint8_t* h, int at, uint16_t v) { std::memcpy(h + at, &v, 2); }
static void put32(uint8_t* h, int at, uint32_t v) { std::memcpy(h + at, &v, 4); }
static void put64(uint8_t* h, int at, double v) { std::memcpy(h + at, &v, 8); }

// The 227-byte LAS 1.2 header of LASIO::writeLAS (lasio.cpp:140-184) or saveResultAsLAS (icp_registration.cpp:706-777).
static void las_header_bytes(int variant, int64_t n, const double* scale, const double* offset, const double* mn, const double* mx,
                             uint8_t* h) {
    std::memset(h, 0, ICP_LAS_HEADER_BYTES);
    std::memcpy(h, "LASF", 4);
    h[24] = 1;
    h[25] = 2;
    if (variant == ICP_VARIANT_CLI) {
        std::memcpy(h + 26, "ICP Registration", 16);
        std::memcpy(h + 58, "Custom ICP", 10);
        put16(h, 90, 307);
        put16(h, 92, 2025);
    }
    put16(h, 94, ICP_LAS_HEADER_BYTES);
    put32(h, 96, ICP_LAS_HEADER_BYTES);
    put16(h, 105, ICP_LAS_RECORD_BYTES);
    put32(h, 107, (uint32_t)n);
    for (int a = 0; a < 3; ++a) {
        put64(h, 131 + 8 * a, scale[a]);
        put64(h, 155 + 8 * a, offset[a]);
        put64(h, 179 + 16 * a, mx[a]);
        put64(h, 187 + 16 * a, mn[a]);
    }
}

}  // namespace icpb

using namespace icpb;

extern "C" {

int icp_las_parse_header(const uint8_t* header227, icp_las_header* out) {
    if (!header227 || !out) return ICP_INVALID_ARGUMENT;
    std::memset(out, 0, sizeof *out);
    if (std::memcmp(header227, "LASF", 4) != 0) return ICP_BAD_FORMAT;  // lasio.cpp:30-34
    uint32_t off, npt;
    uint16_t rl;
    std::memcpy(&off, header227 + 96, 4);
    std::memcpy(&rl, header227 + 105, 2);
    std::memcpy(&npt, header227 + 107, 4);
    out->offset_to_data = off;
    out->n_points = npt;
    out->record_length = rl;
    for (int a = 0; a < 3; ++a) {
        std::memcpy(&out->scale[a], header227 + 131 + 8 * a, 8);
        std::memcpy(&out->offset[a], header227 + 155 + 8 * a, 8);
        std::memcpy(&out->max[a], header227 + 179 + 16 * a, 8);
        std::memcpy(&out->min[a], header227 + 187 + 16 * a, 8);
    }
    return ICP_OK;
}

int icp_las_decode(icp_handle h, const uint8_t* records, int64_t n, int32_t record_length, const double* scale3,
                   const double* offset3, double* xyz_out) {
    Ctx* c = (Ctx*)h;
    if (!c || !scale3 || !offset3 || record_length < 12) return ICP_INVALID_ARGUMENT;
    if (n <= 0) return ICP_OK;
    if (!records || !xyz_out) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)n * (size_t)record_length;
    ICPB_TRY(devbuf_reserve(c, c->scratch2, bytes));
    ICPB_TRY(devbuf_reserve(c, c->scratch_src, (size_t)n * 3 * sizeof(double)));
    ICPB_CUDA(c, cudaMemcpyAsync(c->scratch2.p, records, bytes, cudaMemcpyHostToDevice, c->stream));
    ICPB_TRY(las_decode_launch(c, c->stream, (const uint8_t*)c->scratch2.p, n, record_length, scale3, offset3, (double*)c->scratch_src.p));
    ICPB_CUDA(c, cudaMemcpyAsync(xyz_out, c->scratch_src.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

int icp_las_encode(icp_handle h, const double* xyz, int64_t n, const double* scale3, const double* offset3, uint8_t* records_out) {
    Ctx* c = (Ctx*)h;
    if (!c || !scale3 || !offset3) return ICP_INVALID_ARGUMENT;
    if (n <= 0) return ICP_OK;
    if (!xyz || !records_out) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload_xyz(c, c->scratch_src, xyz, n));
    ICPB_TRY(encode_device(c, (const double*)c->scratch_src.p, n, scale3, offset3));
    ICPB_CUDA(c, cudaMemcpyAsync(records_out, c->scratch2.p, (size_t)n * ICP_LAS_RECORD_BYTES, cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

int icp_cloud_bounds(icp_handle h, const double* xyz, int64_t n, double* min3, double* max3) {
    Ctx* c = (Ctx*)h;
    if (!c || !min3 || !max3) return ICP_INVALID_ARGUMENT;
    if (n <= 0 || !xyz) {  // pointcloud.cpp:26-30: an empty cloud has all-zero bounds
        for (int a = 0; a < 3; ++a) min3[a] = max3[a] = 0.0;
        return ICP_OK;
    }
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload_xyz(c, c->scratch_src, xyz, n));
    return bounds_device(c, (const double*)c->scratch_src.p, n, min3, max3);
}

int icp_las_file_image(icp_handle h, const double* xyz, int64_t n, int variant, const double* scale3, const double* offset3,
                       uint8_t* image_out, int64_t cap, int64_t* bytes_out) {
    Ctx* c = (Ctx*)h;
    if (!c || !bytes_out) return ICP_INVALID_ARGUMENT;
    *bytes_out = 0;
    if (n <= 0 || !xyz) return ICP_EMPTY_INPUT;  // lasio.cpp:128-131 "点云为空，无法写入"
    if (n > 0xFFFFFFFFll) return ICP_INVALID_ARGUMENT;
    if (variant == ICP_VARIANT_CLI && (!scale3 || !offset3)) return ICP_INVALID_ARGUMENT;
    const int64_t need = ICP_LAS_HEADER_BYTES + (int64_t)ICP_LAS_RECORD_BYTES * n;
    *bytes_out = need;
    if (!image_out || cap < need) return ICP_INVALID_ARGUMENT;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload_xyz(c, c->scratch_src, xyz, n));
    double mn[3], mx[3], scale[3], offset[3];
    ICPB_TRY(bounds_device(c, (const double*)c->scratch_src.p, n, mn, mx));
    for (int a = 0; a < 3; ++a) {
        scale[a] = (variant == ICP_VARIANT_CLI) ? scale3[a] : 0.001;   // lasio.cpp:166-168 vs icp_registration.cpp:765-767
        offset[a] = (variant == ICP_VARIANT_CLI) ? offset3[a] : mn[a];  // lasio.cpp:171-173 vs :769-771
    }
    las_header_bytes(variant, n, scale, offset, mn, mx, image_out);
    ICPB_TRY(encode_device(c, (const double*)c->scratch_src.p, n, scale, offset));
    ICPB_CUDA(c, cudaMemcpyAsync(image_out + ICP_LAS_HEADER_BYTES, c->scratch2.p, (size_t)n * ICP_LAS_RECORD_BYTES, cudaMemcpyDeviceToHost,
                                 c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

int icp_las_write(icp_handle h, const char* path, const double* xyz, int64_t n, int variant, const double* scale3,
                  const double* offset3) {
    Ctx* c = (Ctx*)h;
    if (!c || !path) return ICP_INVALID_ARGUMENT;
    if (n <= 0 || !xyz) return ICP_EMPTY_INPUT;
    std::vector<uint8_t> image;
    try {  // (nothing may throw across the C boundary)
        image.resize((size_t)(ICP_LAS_HEADER_BYTES + (int64_t)ICP_LAS_RECORD_BYTES * n));
    } catch (...) {
        c->err = "out of host memory for the LAS file image";
        return ICP_IO_ERROR;
    }
    int64_t bytes = 0;
    ICPB_TRY(icp_las_file_image(h, xyz, n, variant, scale3, offset3, image.data(), (int64_t)image.size(), &bytes));
    FILE* f = std::fopen(path, "wb");
    if (!f) {
        c->err = std::string("cannot create ") + path;  // lasio.cpp:134-137 "无法创建文件"
        return ICP_IO_ERROR;
    }
    const size_t w = std::fwrite(image.data(), 1, (size_t)bytes, f);
    const int rc = std::fclose(f);
    if (w != (size_t)bytes || rc != 0) {
        c->err = std::string("short write to ") + path;
        return ICP_IO_ERROR;
    }
    return ICP_OK;
}

int icp_las_read(icp_handle h, const char* path, int64_t max_points, int variant, icp_las_header* header_out, double* xyz_out,
                 int64_t cap, int64_t* n_out) {
    Ctx* c = (Ctx*)h;
    if (!c || !path || !n_out) return ICP_INVALID_ARGUMENT;
    *n_out = 0;
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        c->err = std::string("cannot open ") + path;  // lasio.cpp:9-12 "无法打开文件"
        return ICP_IO_ERROR;
    }
    uint8_t hdr[ICP_LAS_HEADER_BYTES];
    icp_las_header H;
    if (std::fread(hdr, 1, sizeof hdr, f) != sizeof hdr) {  // lasio.cpp:23-27 "无法读取文件头"
        std::fclose(f);
        c->err = std::string("cannot read the LAS header of ") + path;
        return ICP_IO_ERROR;
    }
    int st = icp_las_parse_header(hdr, &H);
    if (st == ICP_BAD_FORMAT && variant == ICP_VARIANT_CLI) {
        // readLASFile prints the signature but does not check it (icp_registration.cpp:274-278)
        std::memcpy(hdr, "LASF", 4);
        st = icp_las_parse_header(hdr, &H);
    }
    if (st != ICP_OK) {
        std::fclose(f);
        c->err = std::string("not a LAS file: ") + path;
        return st;
    }
    if (header_out) *header_out = H;
    if (variant == ICP_VARIANT_CLI && (H.n_points == 0 || H.n_points > 100000000u)) {  // icp_registration.cpp:291-295
        std::fclose(f);
        c->err = "implausible point count in the LAS header";
        return ICP_BAD_FORMAT;
    }
    int64_t n = (int64_t)H.n_points;
    if (variant == ICP_VARIANT_ENGINE && max_points > 0 && max_points < n) n = max_points;  // lasio.cpp:59-63
    *n_out = n;
    if (!xyz_out) {  // header / size query
        std::fclose(f);
        return ICP_OK;
    }
    if (cap < n || H.record_length < 12) {
        std::fclose(f);
        return ICP_INVALID_ARGUMENT;
    }
    // the header's word is not taken for the file's size: a corrupt count must not drive the allocation
    bool ok = std::fseek(f, 0, SEEK_END) == 0;
    const long long file_bytes = ok ? (long long)std::ftell(f) : -1;
    if (file_bytes < 0 || (long long)H.offset_to_data + (long long)n * (long long)H.record_length > file_bytes) {
        std::fclose(f);
        *n_out = 0;
        c->err = std::string("truncated LAS point data in ") + path;
        return ICP_IO_ERROR;
    }
    std::vector<uint8_t> rec;
    try {
        rec.resize((size_t)n * H.record_length);
    } catch (...) {
        std::fclose(f);
        *n_out = 0;
        c->err = "out of host memory for the LAS point records";
        return ICP_IO_ERROR;
    }
    ok = std::fseek(f, (long)H.offset_to_data, SEEK_SET) == 0;
    ok = ok && std::fread(rec.data(), 1, rec.size(), f) == rec.size();
    std::fclose(f);
    if (!ok) {
        // The reference parses whatever its buffer holds after a short read; a truncated file is reported instead.
        *n_out = 0;
        c->err = std::string("truncated LAS point data in ") + path;
        return ICP_IO_ERROR;
    }
    return icp_las_decode(h, rec.data(), n, (int32_t)H.record_length, H.scale, H.offset, xyz_out);
}

int icp_downsample(icp_handle h, const double* xyz, int64_t n, int32_t target_size, double* xyz_out, int64_t* n_out) {
    Ctx* c = (Ctx*)h;
    if (!c || !n_out) return ICP_INVALID_ARGUMENT;
    *n_out = 0;
    if (n <= 0 || !xyz || target_size <= 0) return ICP_EMPTY_INPUT;  // pointcloud.cpp:109-111 returns nullptr
    if (!xyz_out) return ICP_INVALID_ARGUMENT;
    if (n > 0x7fffffffLL) {  // the reference's sizes are ints (pointcloud.cpp:113)
        c->err = "downsample: more than 2^31 - 1 points";
        return ICP_INVALID_ARGUMENT;
    }
    if (n <= (int64_t)target_size) {  // pointcloud.cpp:117-118: copy
        std::memcpy(xyz_out, xyz, (size_t)n * 3 * sizeof(double));
        *n_out = n;
        return ICP_OK;
    }
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload_xyz(c, c->scratch_src, xyz, n));
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)target_size * 3 * sizeof(double)));
    const double step = (double)n / (double)target_size;
    gather_step_kernel<<<nblk(target_size), 256, 0, c->stream>>>((const double*)c->scratch_src.p, (double*)c->scratch1.p, target_size, step, 0);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    ICPB_CUDA(c, cudaMemcpyAsync(xyz_out, c->scratch1.p, (size_t)target_size * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_out = target_size;
    return ICP_OK;
}

int icp_downsample_stride(icp_handle h, const double* xyz, int64_t n, int64_t stride, double* xyz_out, int64_t* n_out) {
    Ctx* c = (Ctx*)h;
    if (!c || !n_out || stride <= 0) return ICP_INVALID_ARGUMENT;
    *n_out = 0;
    if (n <= 0) return ICP_OK;
    if (!xyz || !xyz_out) return ICP_INVALID_ARGUMENT;
    const int64_t m = (n + stride - 1) / stride;
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload_xyz(c, c->scratch_src, xyz, n));
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)m * 3 * sizeof(double)));
    gather_step_kernel<<<nblk(m), 256, 0, c->stream>>>((const double*)c->scratch_src.p, (double*)c->scratch1.p, m, 0.0, stride);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    ICPB_CUDA(c, cudaMemcpyAsync(xyz_out, c->scratch1.p, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_out = m;
    return ICP_OK;
}

int icp_replay_iteration(icp_handle h, const double* original_xyz, int64_t n, const double* T16, double* xyz_out) {
    Ctx* c = (Ctx*)h;
    if (!c) return ICP_INVALID_ARGUMENT;
    if (n <= 0) return ICP_OK;
    if (!original_xyz || !xyz_out) return ICP_INVALID_ARGUMENT;
    if (!T16) {  // index -1 in the viewer: the untouched original (pointcloudviewer.cpp:95-97)
        std::memmove(xyz_out, original_xyz, (size_t)n * 3 * sizeof(double));
        return ICP_OK;
    }
    ICPB_CUDA(c, cudaSetDevice(c->device));
    ICPB_TRY(upload_xyz(c, c->scratch_src, original_xyz, n));
    ICPB_TRY(devbuf_reserve(c, c->scratch0, 64 * sizeof(double)));
    ICPB_CUDA(c, cudaMemcpyAsync(c->scratch0.p, T16, 16 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ICPB_TRY(apply_aos_launch(c, (const double*)c->scratch0.p, (double*)c->scratch_src.p, n));
    ICPB_CUDA(c, cudaMemcpyAsync(xyz_out, c->scratch_src.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
    return ICP_OK;
}

// Host-only text report; ostream << double at precision 10 is "%.10g".
int icp_save_transformation(const char* path, const double* R9, const double* t3, const double* iteration_T16, int32_t n_iterations) {
    if (!path || !R9 || !t3) return ICP_INVALID_ARGUMENT;
    FILE* f = std::fopen(path, "wb");
    if (!f) return ICP_IO_ERROR;  // "无法创建变换参数文件"
    std::fprintf(f, "ICP配准变换参数\n==================\n\n");
    std::fprintf(f, "说明: 将源点云变换到目标点云坐标系下的变换矩阵\n");
    std::fprintf(f, "变换公式: P_target = R * P_source + t\n\n");
    if (iteration_T16 && n_iterations > 0) {
        std::fprintf(f, "==================\n迭代过程变换参数\n==================\n\n");
        for (int32_t k = 0; k < n_iterations; ++k) {
            const double* T = iteration_T16 + 16 * (int64_t)k;
            std::fprintf(f, "--- 迭代 %d ---\n旋转矩阵 R:\n", k + 1);
            for (int i = 0; i < 3; ++i) std::fprintf(f, "  [%.10g, %.10g, %.10g]\n", T[4 * i], T[4 * i + 1], T[4 * i + 2]);
            std::fprintf(f, "平移向量 t:\n  [%.10g, %.10g, %.10g]\n\n", T[3], T[7], T[11]);
        }
        std::fprintf(f, "\n");
    }
    std::fprintf(f, "==================\n最终变换参数\n==================\n\n");
    std::fprintf(f, "旋转矩阵 R (3x3):\n");
    for (int i = 0; i < 3; ++i) std::fprintf(f, "  [%.10g, %.10g, %.10g]\n", R9[3 * i], R9[3 * i + 1], R9[3 * i + 2]);
    std::fprintf(f, "\n平移向量 t (3x1):\n  [%.10g, %.10g, %.10g]\n", t3[0], t3[1], t3[2]);
    std::fprintf(f, "\n变换矩阵 (齐次坐标形式 4x4):\n");
    for (int i = 0; i < 3; ++i) std::fprintf(f, "  [%.10g, %.10g, %.10g, %.10g]\n", R9[3 * i], R9[3 * i + 1], R9[3 * i + 2], t3[i]);
    std::fprintf(f, "  [0, 0, 0, 1]\n");
    return std::fclose(f) == 0 ? ICP_OK : ICP_IO_ERROR;
}

}  // extern "C"
