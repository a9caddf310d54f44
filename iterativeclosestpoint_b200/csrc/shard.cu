// Multi-GPU data movement of the sharded path (BASELINE.json configs 3 and 4: the source cloud shards by point
// range, the target octree is replicated).  One process per GPU; NCCL over NVLink carries the bulk data, the two tiny
// per-iteration records travel through the peer mailboxes (iter.cu).
//
//   target_upload_sharded   every rank needs the whole target for its replica of the octree, but it does not need to pull all
//                           of it over its own PCIe link: rank r uploads the r-th slice and an all-gather over NVLink hands
//                           every rank the rest (N x 240 MB over PCIe was 2/3 of the 8-GPU end-to-end time).
//   redistribute_source     a rank is handed one contiguous RANGE of the caller's order, which says nothing about where those
//                           points are: its queries then touch the whole replicated structure.  The ranks therefore re-deal
//                           the points by region: a coarse Morton histogram (32 x 32 x 32 bins of the target's cube) is summed
//                           over the ranks, every rank cuts the same R contiguous key ranges of equal population out of it,
//                           partitions its points by destination (stable, so the result does not depend on scheduling) and
//                           one grouped send / receive moves them.  The rank's resident source is what it received.
//   redistribute_return     the moved points travel back the same way and are scattered to where the caller had them.
#include "internal.h"
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <vector>

namespace icpb {

NcclApi* nccl_load(Ctx* c) {
    if (c->nccl) return c->nccl;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        c->err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return nullptr;
    }
    NcclApi* a = new NcclApi();
    a->lib = lib;
    a->GetUniqueId = (decltype(a->GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    a->CommInitRank = (decltype(a->CommInitRank))dlsym(lib, "ncclCommInitRank");
    a->CommDestroy = (decltype(a->CommDestroy))dlsym(lib, "ncclCommDestroy");
    a->AllGather = (decltype(a->AllGather))dlsym(lib, "ncclAllGather");
    a->AllReduce = (decltype(a->AllReduce))dlsym(lib, "ncclAllReduce");
    a->Send = (decltype(a->Send))dlsym(lib, "ncclSend");
    a->Recv = (decltype(a->Recv))dlsym(lib, "ncclRecv");
    a->GroupStart = (decltype(a->GroupStart))dlsym(lib, "ncclGroupStart");
    a->GroupEnd = (decltype(a->GroupEnd))dlsym(lib, "ncclGroupEnd");
    a->GetErrorString = (decltype(a->GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!a->GetUniqueId || !a->CommInitRank || !a->CommDestroy || !a->AllGather || !a->AllReduce || !a->Send || !a->Recv ||
        !a->GroupStart || !a->GroupEnd || !a->GetErrorString) {
        c->err = "libnccl.so.2 lacks a required symbol";
        delete a;
        return nullptr;
    }
    c->nccl = a;
    return a;
}

// ---------------------------------------------------------------------------------------------------------------------------
int target_upload_sharded(Ctx* c, const double* host_tgt_xyz, int64_t n_tgt) {
    const int R = c->n_ranks;
    const int64_t chunk = (n_tgt + R - 1) / R;  // points per rank (the last slice may be short; the buffer is padded)
    ICPB_TRY(devbuf_reserve(c, c->tgt_raw, (size_t)chunk * R * 3 * sizeof(double)));
    double* d = (double*)c->tgt_raw.p;
    const int64_t lo = std::min<int64_t>(chunk * c->rank, n_tgt), hi = std::min<int64_t>(lo + chunk, n_tgt);
    if (hi > lo)
        ICPB_CUDA(c, cudaMemcpyAsync(d + 3 * lo, host_tgt_xyz + 3 * lo, (size_t)(hi - lo) * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ICPB_NCCL(c, c->nccl->AllGather(d + 3 * chunk * c->rank, d, (size_t)chunk * 3, ncclFloat64, (ncclComm_t)c->comm, c->stream));
    return ICP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
constexpr int RD_BITS = 5;                       // bins per axis = 32
constexpr int RD_BINS = 1 << (3 * RD_BITS);      // 32768

__global__ void __launch_bounds__(256) rd_bin_kernel(const double* __restrict__ xyz, int64_t n, double lx, double ly, double lz,
                                                     double inv, uint64_t* __restrict__ bin_out, uint32_t* __restrict__ idx_out,
                                                     unsigned int* __restrict__ hist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    const double l[3] = {lx, ly, lz};
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double f = (p[a] - l[a]) * inv;
        if (!(f > 0.0)) f = 0.0;  // also catches NaN
        if (f > (double)((1 << RD_BITS) - 1)) f = (double)((1 << RD_BITS) - 1);
        q[a] = (uint32_t)f;
    }
    uint32_t key = 0;
    for (int b = RD_BITS - 1; b >= 0; --b) key = (key << 3) | ((q[0] >> b) & 1u) | (((q[1] >> b) & 1u) << 1) | (((q[2] >> b) & 1u) << 2);
    bin_out[i] = key;
    idx_out[i] = (uint32_t)i;
    atomicAdd(hist + key, 1u);
}

// bin -> destination rank, in place; counts per destination
__global__ void __launch_bounds__(256) rd_dest_kernel(uint64_t* __restrict__ key, int64_t n, const unsigned char* __restrict__ dest_of_bin,
                                                      unsigned int* __restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int d = dest_of_bin[key[i]];
    key[i] = d;
    atomicAdd(cnt + d, 1u);
}

__global__ void __launch_bounds__(256) rd_gather_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ idx, int64_t n,
                                                        double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t i = idx[j];
    out[3 * j] = xyz[3 * i];
    out[3 * j + 1] = xyz[3 * i + 1];
    out[3 * j + 2] = xyz[3 * i + 2];
}

__global__ void __launch_bounds__(256) rd_scatter_kernel(const double* __restrict__ in, const uint32_t* __restrict__ idx, int64_t n,
                                                         double* __restrict__ xyz) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t i = idx[j];
    xyz[3 * i] = in[3 * j];
    xyz[3 * i + 1] = in[3 * j + 1];
    xyz[3 * i + 2] = in[3 * j + 2];
}

static int exchange(Ctx* c, const double* send, const int64_t* cnt_s, double* recv, const int64_t* cnt_r) {
    NcclApi* a = c->nccl;
    ICPB_NCCL(c, a->GroupStart());
    int64_t os = 0, orr = 0;
    for (int p = 0; p < c->n_ranks; ++p) {
        if (cnt_s[p] > 0) ICPB_NCCL(c, a->Send(send + 3 * os, (size_t)cnt_s[p] * 3, ncclFloat64, p, (ncclComm_t)c->comm, c->stream));
        if (cnt_r[p] > 0) ICPB_NCCL(c, a->Recv(recv + 3 * orr, (size_t)cnt_r[p] * 3, ncclFloat64, p, (ncclComm_t)c->comm, c->stream));
        os += cnt_s[p];
        orr += cnt_r[p];
    }
    ICPB_NCCL(c, a->GroupEnd());
    return ICP_OK;
}

// d_xyz: this rank's points (AoS, device, n_in of them, caller order).  On return *d_out / *n_out are the points this rank
// owns from now on (AoS, device, arrival order: by sending rank, each rank's points in its own caller order).
int redistribute_source(Ctx* c, const double* d_xyz, int64_t n_in, const double** d_out, int64_t* n_out) {
    c->rd_active = false;
    *d_out = d_xyz;
    *n_out = n_in;
    const int R = c->n_ranks;
    if (R <= 1 || R > MAIL_RANKS || !c->opt_redistribute || !c->comm || !c->fast.valid) return ICP_OK;
    cudaStream_t s = c->stream;
    const int64_t n1 = std::max<int64_t>(n_in, 1);
    // scratch: keys (2 x u64) + indices (2 x u32) for the stable partition, histogram, counts, destination table
    ICPB_TRY(devbuf_reserve(c, c->rd_tmp, (size_t)n1 * (2 * sizeof(uint64_t) + 2 * sizeof(uint32_t)) + (RD_BINS + 64 + 64) * sizeof(unsigned int) + RD_BINS));
    uint64_t* keys = (uint64_t*)c->rd_tmp.p;
    uint64_t* keys_alt = keys + n1;
    uint32_t* idx = (uint32_t*)(keys_alt + n1);
    uint32_t* idx_alt = idx + n1;
    unsigned int* hist = (unsigned int*)(idx_alt + n1);
    unsigned int* cnt = hist + RD_BINS;          // [64] this rank's counts per destination
    unsigned int* d_mat = cnt + 64;              // [64] everyone's counts: row r = what rank r sends to each destination
    unsigned char* dest_of_bin = (unsigned char*)(d_mat + 64);  // [RD_BINS]
    ICPB_CUDA(c, cudaMemsetAsync(hist, 0, (RD_BINS + 64) * sizeof(unsigned int), s));
    const int kb = (int)((n_in + 255) / 256);
    const double inv = (double)(1 << RD_BITS) / c->fast.cube;
    if (n_in > 0) rd_bin_kernel<<<kb, 256, 0, s>>>(d_xyz, n_in, c->fast.root_lo[0], c->fast.root_lo[1], c->fast.root_lo[2], inv, keys, idx, hist);
    c->launches++;
    ICPB_NCCL(c, c->nccl->AllReduce(hist, hist, RD_BINS, ncclUint32, ncclSum, (ncclComm_t)c->comm, s));
    std::vector<unsigned int> h_hist(RD_BINS);
    ICPB_CUDA(c, cudaMemcpyAsync(h_hist.data(), hist, RD_BINS * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    ICPB_CUDA(c, cudaStreamSynchronize(s));
    // R contiguous key ranges of (nearly) equal population: identical on every rank (same histogram, same arithmetic)
    std::vector<unsigned char> h_dest(RD_BINS);
    {
        unsigned long long total = 0;
        for (unsigned int v : h_hist) total += v;
        unsigned long long run = 0;
        int d = 0;
        for (int b = 0; b < RD_BINS; ++b) {
            // bin b goes to the rank whose share its midpoint falls into
            const unsigned long long mid = run + h_hist[b] / 2;
            while (d < R - 1 && total > 0 && mid * (unsigned long long)R >= (unsigned long long)(d + 1) * total) ++d;
            h_dest[b] = (unsigned char)d;
            run += h_hist[b];
        }
    }
    ICPB_CUDA(c, cudaMemcpyAsync(dest_of_bin, h_dest.data(), RD_BINS, cudaMemcpyHostToDevice, s));
    if (n_in > 0) rd_dest_kernel<<<kb, 256, 0, s>>>(keys, n_in, dest_of_bin, cnt);
    c->launches++;
    // stable partition by destination: one radix pass over the 8-bit destination
    if (n_in > 0) ICPB_TRY(sort_pairs_u64_u32(c, keys, keys_alt, idx, idx_alt, n_in, 8));
    ICPB_TRY(devbuf_reserve(c, c->rd_perm, (size_t)n1 * sizeof(uint32_t)));
    ICPB_CUDA(c, cudaMemcpyAsync(c->rd_perm.p, idx, (size_t)n_in * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    ICPB_TRY(devbuf_reserve(c, c->rd_send, (size_t)n1 * 3 * sizeof(double)));
    if (n_in > 0) rd_gather_kernel<<<kb, 256, 0, s>>>(d_xyz, (const uint32_t*)c->rd_perm.p, n_in, (double*)c->rd_send.p);
    c->launches++;
    ICPB_CUDA(c, cudaMemcpyAsync(d_mat + (size_t)c->rank * R, cnt, (size_t)R * sizeof(unsigned int), cudaMemcpyDeviceToDevice, s));
    ICPB_NCCL(c, c->nccl->AllGather(d_mat + (size_t)c->rank * R, d_mat, (size_t)R, ncclUint32, (ncclComm_t)c->comm, s));
    std::vector<unsigned int> mat((size_t)R * R);
    ICPB_CUDA(c, cudaMemcpyAsync(mat.data(), d_mat, mat.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    ICPB_CUDA(c, cudaStreamSynchronize(s));
    int64_t n_recv = 0;
    for (int p = 0; p < R; ++p) {
        c->rd_cnt_s[p] = mat[(size_t)c->rank * R + p];
        c->rd_cnt_r[p] = mat[(size_t)p * R + c->rank];
        n_recv += c->rd_cnt_r[p];
    }
    ICPB_TRY(devbuf_reserve(c, c->rd_recv, (size_t)std::max<int64_t>(n_recv, 1) * 3 * sizeof(double)));
    ICPB_TRY(exchange(c, (const double*)c->rd_send.p, c->rd_cnt_s, (double*)c->rd_recv.p, c->rd_cnt_r));
    ICPB_CUDA(c, cudaGetLastError());
    c->rd_active = true;
    c->rd_n_in = n_in;
    c->rd_n_recv = n_recv;
    *d_out = (const double*)c->rd_recv.p;
    *n_out = n_recv;
    return ICP_OK;
}

// d_recv_order_xyz: the rank's resident points in arrival order (n_recv).  d_caller_order_xyz receives this rank's ORIGINAL
// points (n_in), moved, in the caller's order.
int redistribute_return(Ctx* c, const double* d_recv_order_xyz, double* d_caller_order_xyz) {
    if (!c->rd_active) return ICP_INVALID_ARGUMENT;
    ICPB_TRY(devbuf_reserve(c, c->rd_back, (size_t)std::max<int64_t>(c->rd_n_in, 1) * 3 * sizeof(double)));
    ICPB_TRY(exchange(c, d_recv_order_xyz, c->rd_cnt_r, (double*)c->rd_back.p, c->rd_cnt_s));
    const int64_t n = c->rd_n_in;
    if (n > 0) rd_scatter_kernel<<<(int)((n + 255) / 256), 256, 0, c->stream>>>((const double*)c->rd_back.p, (const uint32_t*)c->rd_perm.p, n, d_caller_order_xyz);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

}  // namespace icpb
