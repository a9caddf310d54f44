// Target octree construction as a LINEAR octree (replaces Octree::Octree + Octree::buildTree,
// PointCloudRegistration/core/octree.cpp:41-126; CLI twin icp_registration.cpp:66-104,155-185).
//
//   K0  bbox        strict min/max of the target + the 0.001 expansion            (octree.cpp:47-64)
//   K1  keys        per point, the octant path of buildTree's recursion, obtained by the SAME FP64
//                   bisection mid = (lo+hi)/2, bit = p > mid, 3 bits per level      (octree.cpp:97-110)
//   K2  sort        stable LSD radix sort of (key, index), 8 bits per pass, one kernel per pass (chained scan); the keys are
//                   taken 16 levels deep first and to the full depth only if some node has to split below that
//   K3  nodes       all levels in one cooperative launch: a prefix with more than max_pts points and depth < max_depth is an
//                   inner node, otherwise a leaf (octree.cpp:88); breadth-first numbering
//
// Cell membership is decided by comparisons against bisection midpoints, never by scaling, so every
// point lands in exactly the reference's leaf and every box is bit-identical to the reference's.
#include "internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace icpb {

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
int devbuf_reserve(Ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return ICP_OK;
    if (b.p) ICPB_CUDA(c, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    ICPB_CUDA(c, cudaMalloc(&b.p, want));
    b.cap = want;
    return ICP_OK;
}

int pinned_reserve(Ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return ICP_OK;
    if (b.p) ICPB_CUDA(c, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    ICPB_CUDA(c, cudaHostAlloc(&b.p, want, cudaHostAllocDefault));
    b.cap = want;
    return ICP_OK;
}

void devbuf_free(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (three kernels; sizes here are at most a few 10^7)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (int)(blockDim.x >> 5)) ? smem_warp[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem_warp[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem_warp[32] = winc;
    }
    __syncthreads();
    uint32_t res = smem_warp[warp] + inc - v;
    *total = smem_warp[32];
    __syncthreads();
    return res;
}

// ------------------------------------------------------------------------------------------------
// K0: bounding box (octree.cpp:47-64).  min/max are order-independent, so a parallel reduction gives the
// reference's values exactly; NaN coordinates never replace the running bound there (strict < / >) and
// are skipped here by fmin/fmax, except that a NaN in point 0 seeds the bound in the reference -- handled
// in the finishing step.
// ------------------------------------------------------------------------------------------------
constexpr int BBOX_THREADS = 256;

__global__ void __launch_bounds__(BBOX_THREADS) bbox_partial_kernel(const double* __restrict__ xyz, int64_t n,
                                                                    double* __restrict__ part /* [grid][6] */) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double v = xyz[3 * i + a];
            lo[a] = fmin(lo[a], v);
            hi[a] = fmax(hi[a], v);
        }
    }
    __shared__ double sm[6][BBOX_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) {
            sm[a][warp] = lo[a];
            sm[3 + a][warp] = hi[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[threadIdx.x][0];
        for (int w = 1; w < BBOX_THREADS / 32; ++w) v = (threadIdx.x < 3) ? fmin(v, sm[threadIdx.x][w]) : fmax(v, sm[threadIdx.x][w]);
        part[(int64_t)blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(192) bbox_finish_kernel(const double* __restrict__ part, int n_part, const double* __restrict__ xyz,
                                                          double* __restrict__ root /* lo[3], hi[3] */) {
    // warp a (0..5) reduces component a of the block partials; min/max are order-independent
    const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = (a < 3) ? DBL_MAX : -DBL_MAX;
    for (int k = lane; k < n_part; k += 32) v = (a < 3) ? fmin(v, part[(int64_t)k * 6 + a]) : fmax(v, part[(int64_t)k * 6 + a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (a < 3) ? fmin(v, w) : fmax(v, w);
    }
    if (lane != 0) return;
    const double p0 = xyz[a % 3];
    if (p0 != p0) v = p0;  // pts[0] seeds the scan (octree.cpp:47-49): a NaN there sticks
    const double eps = 0.001;
    root[a] = (a < 3) ? dsub(v, eps) : dadd(v, eps);  // octree.cpp:61-64
}

// Search tree only: make the root a cube (one edge length for all axes) so that its cells are isotropic.
__global__ void cube_root_kernel(double* __restrict__ root) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double ext = fmax(fmax(root[3] - root[0], root[4] - root[1]), root[5] - root[2]) * (1.0 + 1e-12);
    for (int a = 0; a < 3; ++a) root[3 + a] = fmax(root[3 + a], root[a] + ext);
}

// ------------------------------------------------------------------------------------------------
// K1: octant-path keys by FP64 bisection (octree.cpp:97-110 applied max_depth times)
// ------------------------------------------------------------------------------------------------
// A 64-bit word holds 21 levels; a deeper tree (the reference accepts octreeMaxDepth up to 50, settingspage.cpp:76) takes
// ceil(max_depth / 21) words per point: word w of point i at keys[w * n + i], most significant word first.
constexpr int KEY_LEVELS = 21;
constexpr int KEY_DEPTH_FIRST = 16;  // levels of octant keys a build takes first (build_tree)

__global__ void __launch_bounds__(256) morton_keys_kernel(const double* __restrict__ xyz, int64_t n,
                                                          const double* __restrict__ root, int max_depth,
                                                          uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lo[3], hi[3], p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = root[a];
        hi[a] = root[3 + a];
        p[a] = xyz[3 * i + a];
    }
    uint64_t key = 0;
    int word = 0;
    for (int d = 0; d < max_depth; ++d) {
        uint32_t oct = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double mid = dmul(dadd(lo[a], hi[a]), 0.5);  // (min+max)/2, exact halving
            const bool up = p[a] > mid;                        // strict, octree.cpp:106-108
            oct |= (up ? 1u : 0u) << a;
            lo[a] = up ? mid : lo[a];
            hi[a] = up ? hi[a] : mid;
        }
        key = (key << 3) | oct;
        if ((d + 1) % KEY_LEVELS == 0 && d + 1 < max_depth) {
            keys[(int64_t)word * n + i] = key;
            key = 0;
            ++word;
        }
    }
    keys[(int64_t)word * n + i] = key;
    idx[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------------------------------
// K2: stable LSD radix sort, 8 bits per pass, (uint64 key, uint32 payload)
// Each block owns a contiguous tile; each warp a contiguous chunk of the tile, walked in rounds of 32
// consecutive elements, so ranks follow input order (stability).
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef ICPB_RS_ROUNDS
#define ICPB_RS_ROUNDS 16
#endif
constexpr int RS_ROUNDS = ICPB_RS_ROUNDS;                            // elements per thread
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;          // 4096 elements per block
constexpr int RS_CHUNK = 32 * RS_ROUNDS;                 // per warp

// One pass = ONE kernel (chained scan with decoupled look-back): the digit histograms of ALL passes are taken up front in a
// single read of the keys (they do not depend on the order), so a pass only has to find, for each of its tiles, how many
// keys of every digit lie in the tiles before it.  A tile publishes its own digit counts at once (AGGREGATE), then walks back
// over its predecessors' words until it meets one that already holds an inclusive PREFIX, and publishes its own prefix.
// Tiles are handed out by an atomic ticket, so every predecessor of a waiting tile is running or done: no deadlock.
constexpr uint64_t OS_AGG = 1ull << 62, OS_INCL = 2ull << 62, OS_VAL = OS_AGG - 1;
constexpr int OS_MAX_PASSES = 8;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ghist[p][d] += number of keys whose digit of pass p is d.  A thread keeps a run (digit, count) per pass in registers and
// only touches shared memory when the digit changes: clouds in scan order have long runs in the upper digits, which would
// otherwise serialise the shared-memory atomics of a warp on one address.
template <int P>
__global__ void __launch_bounds__(256) radix_hist_all_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[P][256];
    for (int k = threadIdx.x; k < P * 256; k += 256) (&h[0][0])[k] = 0;
    __syncthreads();
    uint32_t cur[P], cnt[P];
#pragma unroll
    for (int p = 0; p < P; ++p) cur[p] = cnt[p] = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const uint64_t key = __ldg(keys + i);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const uint32_t d = (uint32_t)(key >> (8 * p)) & 0xFFu;
            if (d == cur[p]) {
                ++cnt[p];
            } else {
                if (cnt[p]) atomicAdd(&h[p][cur[p]], cnt[p]);
                cur[p] = d;
                cnt[p] = 1;
            }
        }
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (cnt[p]) atomicAdd(&h[p][cur[p]], cnt[p]);
    __syncthreads();
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const uint32_t v = h[p][threadIdx.x];
        if (v) atomicAdd(&ghist[p * 256 + threadIdx.x], v);
    }
}

// block p: gstart[p][d] = number of keys whose digit of pass p is smaller than d
__global__ void __launch_bounds__(256) radix_digit_scan_kernel(const uint32_t* __restrict__ ghist, uint32_t* __restrict__ gstart) {
    __shared__ uint32_t sw[33];
    const uint32_t v = ghist[blockIdx.x * 256 + threadIdx.x];
    uint32_t total;
    gstart[blockIdx.x * 256 + threadIdx.x] = block_exclusive_scan(v, sw, &total);
}

// The scatter of one pass.  Each warp owns a contiguous chunk of the tile and walks it in rounds of 32 consecutive keys (kept
// in registers), ranking equal digits by __match_any_sync, so ranks follow input order (stability).  The tile is first put in
// digit order in shared memory, then copied out: neighbouring threads write neighbouring addresses of the same digit's run.
constexpr size_t RS_SCATTER_SMEM = (size_t)RS_TILE * (sizeof(uint64_t) + sizeof(uint32_t)) + (RS_WARPS * 256 + 512 + 16) * sizeof(uint32_t);

__global__ void __launch_bounds__(RS_THREADS, 3) radix_onesweep_kernel(const uint64_t* __restrict__ keys_in,
                                                                       const uint32_t* __restrict__ vals_in, int64_t n, int shift,
                                                                       const uint32_t* __restrict__ gstart /* [256] of this pass */,
                                                                       uint64_t* __restrict__ status /* [tiles][256], zeroed */,
                                                                       uint32_t* __restrict__ ticket /* zeroed */,
                                                                       uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(rs_smem);              // [RS_TILE]
    uint32_t* svals = reinterpret_cast<uint32_t*>(skeys + RS_TILE);      // [RS_TILE]
    uint32_t (*wh)[256] = reinterpret_cast<uint32_t (*)[256]>(svals + RS_TILE);  // [RS_WARPS][256] per-warp counts, then offsets
    uint32_t* dstart = &wh[0][0] + RS_WARPS * 256;                       // [256] start of digit d inside the staged tile
    uint32_t* gbase = dstart + 256;                                      // [256] start of this tile's digit-d run in the output
    uint32_t* wtot = gbase + 256;                                        // [8]
    uint32_t* s_tile = wtot + 8;
    if (threadIdx.x == 0) *s_tile = atomicAdd(ticket, 1u);
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&wh[0][0])[k] = 0;
    __syncthreads();
    const uint32_t tile = *s_tile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t block_base = (int64_t)tile * RS_TILE;
    const int64_t chunk_base = block_base + (int64_t)warp * RS_CHUNK;
    uint64_t key[RS_ROUNDS];
    uint32_t rank2[RS_ROUNDS / 2];  // position among the keys of the same digit in this warp's chunk (< 512): two per word
#pragma unroll
    for (int r = 0; r < RS_ROUNDS / 2; ++r) rank2[r] = 0;
    {
        uint32_t* hist = wh[warp];
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; ++r) {
            const int64_t i = chunk_base + (int64_t)r * 32 + lane;
            const bool ok = i < n;
            key[r] = ok ? keys_in[i] : ~0ull;
            const uint32_t d = ok ? (uint32_t)((key[r] >> shift) & 0xFF) : 0x100u;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t before = __popc(peers & lt);
            uint32_t old = 0;
            if (ok && before == 0) {
                old = hist[d];
                hist[d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
            rank2[r >> 1] |= (old + before) << (16 * (r & 1));
            __syncwarp();
        }
    }
    __syncthreads();
    uint32_t my_cnt = 0;
    {
        // thread d: this tile's count of digit d; publish it; scan over the digits (staging order) and per-warp offsets
        const int d = threadIdx.x;
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) cnt += wh[w][d];
        uint64_t* my = status + (size_t)tile * 256 + d;
        st_relaxed_u64(my, (uint64_t)cnt | (tile == 0 ? OS_INCL : OS_AGG));
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t before_warps = 0;
        for (int w = 0; w < warp; ++w) before_warps += wtot[w];
        uint32_t run = before_warps + incl - cnt;
        dstart[d] = run;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
        my_cnt = cnt;
    }
    __syncthreads();
    {
        const uint32_t* off = wh[warp];
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; ++r) {
            const int64_t i = chunk_base + (int64_t)r * 32 + lane;
            if (i < n) {
                const uint32_t d = (uint32_t)((key[r] >> shift) & 0xFF);
                const uint32_t dst = off[d] + ((rank2[r >> 1] >> (16 * (r & 1))) & 0xFFFFu);  // position inside the staged tile
                skeys[dst] = key[r];
                svals[dst] = vals_in[i];
            }
        }
    }
    {
        // look back over the tiles before this one (after the staging: the predecessors had that long to publish their prefixes)
        const int d = threadIdx.x;
        const uint32_t cnt = my_cnt;
        uint64_t* my = status + (size_t)tile * 256 + d;
        uint64_t excl = 0;
        if (tile > 0) {
            const uint64_t* look = my - 256;
            for (;;) {
                uint64_t v;
                do {
                    v = ld_relaxed_u64(look);
                } while ((v >> 62) == 0);
                excl += v & OS_VAL;
                if ((v >> 62) == 2) break;
                look -= 256;
            }
            st_relaxed_u64(my, (excl + cnt) | OS_INCL);
        }
        gbase[d] = gstart[d] + (uint32_t)excl;
    }
    __syncthreads();
    const int count = (int)min((int64_t)RS_TILE, n - block_base);
    for (int j = threadIdx.x; j < count; j += RS_THREADS) {
        const uint64_t k = skeys[j];
        const uint32_t d = (uint32_t)((k >> shift) & 0xFF);
        const uint32_t dst = gbase[d] + ((uint32_t)j - dstart[d]);
        keys_out[dst] = k;
        vals_out[dst] = svals[j];
    }
}

template <int P>
static void hist_all_launch(Ctx* c, const uint64_t* keys, int64_t n, uint32_t* ghist) {
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
    radix_hist_all_kernel<P><<<blocks, 256, 0, c->stream>>>(keys, n, ghist);
}

static int radix_sort_pairs(Ctx* c, uint64_t*& keys, uint64_t*& keys_alt, uint32_t*& vals, uint32_t*& vals_alt,
                            int64_t n, int key_bits) {
    if (n <= 0 || key_bits <= 0) return ICP_OK;
    const int n_pass = (key_bits + 7) / 8;
    if (n_pass > OS_MAX_PASSES) {
        c->err = "radix sort: more than 64 key bits";
        return ICP_INVALID_ARGUMENT;
    }
    const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    if (!c->rs_smem_opt_in) {  // more than 48 KB of dynamic shared memory: opt in once per handle (= per device context)
        ICPB_CUDA(c, cudaFuncSetAttribute(radix_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SCATTER_SMEM));
        c->rs_smem_opt_in = true;
    }
    // scratch: ghist[8][256] | gstart[8][256] | ticket (64 B) | status[n_tiles][256]
    const size_t head = (size_t)2 * OS_MAX_PASSES * 256 * sizeof(uint32_t);
    const size_t status_bytes = 64 + (size_t)n_tiles * 256 * sizeof(uint64_t);
    ICPB_TRY(devbuf_reserve(c, c->scratch2, head + status_bytes));
    uint32_t* ghist = (uint32_t*)c->scratch2.p;
    uint32_t* gstart = ghist + OS_MAX_PASSES * 256;
    uint32_t* ticket = (uint32_t*)((char*)c->scratch2.p + head);
    uint64_t* status = (uint64_t*)((char*)ticket + 64);
    ICPB_CUDA(c, cudaMemsetAsync(ghist, 0, OS_MAX_PASSES * 256 * sizeof(uint32_t), c->stream));
    switch (n_pass) {
        case 1: hist_all_launch<1>(c, keys, n, ghist); break;
        case 2: hist_all_launch<2>(c, keys, n, ghist); break;
        case 3: hist_all_launch<3>(c, keys, n, ghist); break;
        case 4: hist_all_launch<4>(c, keys, n, ghist); break;
        case 5: hist_all_launch<5>(c, keys, n, ghist); break;
        case 6: hist_all_launch<6>(c, keys, n, ghist); break;
        case 7: hist_all_launch<7>(c, keys, n, ghist); break;
        default: hist_all_launch<8>(c, keys, n, ghist); break;
    }
    radix_digit_scan_kernel<<<n_pass, 256, 0, c->stream>>>(ghist, gstart);
    c->launches += 2;
    for (int p = 0; p < n_pass; ++p) {
        ICPB_CUDA(c, cudaMemsetAsync(ticket, 0, status_bytes, c->stream));
        radix_onesweep_kernel<<<(unsigned)n_tiles, RS_THREADS, RS_SCATTER_SMEM, c->stream>>>(keys, vals, n, 8 * p, gstart + p * 256, status,
                                                                                             ticket, keys_alt, vals_alt);
        c->launches++;
        std::swap(keys, keys_alt);
        std::swap(vals, vals_alt);
    }
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int sort_pairs_u64_u32(Ctx* c, uint64_t*& keys, uint64_t*& keys_alt, uint32_t*& vals, uint32_t*& vals_alt, int64_t n, int key_bits) {
    return radix_sort_pairs(c, keys, keys_alt, vals, vals_alt, n, key_bits);
}

__global__ void __launch_bounds__(256) gather_u64_kernel(const uint64_t* __restrict__ in, const uint32_t* __restrict__ idx, int64_t n,
                                                         uint64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

// gather the target into Morton order: (x, y, z, original index)
__global__ void __launch_bounds__(256) gather_points_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ idx,
                                                            int64_t n, TPoint* __restrict__ out, uint32_t* __restrict__ pos_of_idx0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = idx[i];
    TPoint p;
    p.x = xyz[3 * (int64_t)j];
    p.y = xyz[3 * (int64_t)j + 1];
    p.z = xyz[3 * (int64_t)j + 2];
    p.idx = (long long)j;
    out[i] = p;
    if (j == 0) *pos_of_idx0 = (uint32_t)i;
}

// ------------------------------------------------------------------------------------------------
// K3: node table, one level per step.  A node of level d covers the sorted range [pt0, pt0+npts) whose
// keys share their top 3d bits.  (octree.cpp:86-126)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lower_bound_octant(const uint64_t* __restrict__ keys, uint32_t lo, uint32_t hi,
                                                       int shift, uint32_t oct) {
    // first position in [lo, hi) whose octant at `shift` is >= oct
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        uint32_t o = (uint32_t)((keys[mid] >> shift) & 7u);
        if (o < oct) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// All levels in ONE cooperative launch (every block resident, a grid barrier between levels).  A block takes tiles of
// LV_TILE nodes of the level in increasing order.  One thread per node first sorts out the leaves (most nodes of the deep
// levels) and lists the nodes that split, in order.  Eight lanes per listed node then find its children: lane `oct` looks up
// where octant `oct` starts in the node's sorted range, its neighbour's start is where it ends.  The first-child index of a
// node is the number of children of all nodes before it, found without a second pass: a tile publishes its own child count
// at once and looks back over the earlier tiles' words (chained scan, as in the sort above; a block never waits for a later
// tile, so it cannot deadlock).  The node numbering is the breadth-first one a count / scan / emit sequence per level gives.
constexpr int LV_THREADS = 256, LV_ROUNDS = 2, LV_TILE = LV_THREADS * LV_ROUNDS, LV_WARPS = LV_THREADS / 32, LV_UNROLL = 4;
constexpr uint32_t LV_FLAG_CAPACITY = 1, LV_FLAG_DEEPER = 2;

struct LevelState {  // zeroed before the launch
    uint32_t level_total[66];  // children emitted from level L (written by that level's last tile)
    uint32_t level_flags[66];  // flags raised while level L was processed: what the blocks decide on after that level's barrier
                               // (a block still reading after barrier L must not see what a faster block raises in L + 1)
    uint32_t n_leaves;
    uint32_t flags;            // LV_FLAG_CAPACITY: the node table is too small; LV_FLAG_DEEPER: a node must split below key_depth
    uint32_t needed;           // node slots wanted so far
    uint32_t depth, n_nodes;   // the result
    uint32_t barrier;
#ifdef ICPB_LVTIME
    unsigned long long t_level[68];  // profiling build: globaltimer when block 0 left each level's barrier
    uint32_t n_level[68];
#endif
};

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire_u32(bar) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(LV_THREADS) octree_levels_kernel(Node* nodes, uint32_t* parent, uint64_t* cell, uint32_t cap_nodes,
                                                                   const uint64_t* __restrict__ keys /* [n_words][m], sorted */,
                                                                   int64_t m, int n_words, int key_depth, int max_pts, int max_depth,
                                                                   uint64_t* status, LevelState* st) {
    __shared__ uint32_t s_list[LV_TILE];      // the tile's nodes that split (index within the level), in order
    __shared__ uint32_t s_cnt[LV_TILE];       // children per listed node, then their exclusive prefix
    __shared__ uint32_t s_b[LV_TILE * 8];     // where octant o of listed node g starts: s_b[g * 8 + o]
    __shared__ uint32_t s_w[LV_WARPS];
    __shared__ uint32_t s_excl, s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t oct = threadIdx.x & 7u, group = threadIdx.x >> 3;
    const uint32_t lt = (1u << lane) - 1u;
#ifdef ICPB_LVTIME
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        st->t_level[0] = t0;
    }
#endif
    uint32_t first = 0, count = 1, epoch = 0;
    for (int level = 0;; ++level) {
        // the key word and the bit position of this level's octant
        const int w = min(level / KEY_LEVELS, n_words - 1);
        const int wl = min(KEY_LEVELS, key_depth - KEY_LEVELS * w);
        const int shift = level < key_depth ? 3 * (wl - 1 - (level - KEY_LEVELS * w)) : 0;
        const uint64_t* lk = keys + (size_t)w * m;
        const uint32_t next_first = first + count;
        // tile size of this level: LV_TILE nodes, less when the level has too few nodes to give every block one (the upper
        // levels: every node splits, and its searches run over long ranges -- spread them over the blocks)
        const uint32_t tn = min((uint32_t)LV_TILE, max(32u, ((count + gridDim.x - 1) / gridDim.x + 31u) & ~31u));
        const uint32_t n_tiles = (count + tn - 1) / tn;
        const uint64_t tag = (uint64_t)(level + 1) << 34;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // ---- one thread per node: leaves are finished here, the nodes that split are listed in order ----
            uint32_t n_split = 0, leaves = 0;
#pragma unroll
            for (int r = 0; r < LV_ROUNDS; ++r) {
                const uint32_t in_tile = r * LV_THREADS + threadIdx.x;
                const uint32_t t = tile * tn + in_tile;
                bool split = false;
                if (in_tile < tn && t < count) {
                    Node* nd = nodes + first + t;
                    const uint4 tail = __ldcg(reinterpret_cast<const uint4*>(nd) + 3);  // child0, pt0, npts, meta
                    split = !(tail.z <= (uint32_t)max_pts || level >= max_depth);       // octree.cpp:88
                    if (split && level >= key_depth) {  // its octant bits are not in the keys: the caller re-sorts
                        atomicOr(&st->level_flags[level], LV_FLAG_DEEPER);
                        atomicOr(&st->flags, LV_FLAG_DEEPER);
                        split = false;
                    }
                    if (!split) {  // leaf: empty child mask
                        nd->child0 = 0;
                        nd->meta = ((uint32_t)level << 8);
                        ++leaves;
                    }
                }
                const uint32_t mine = __ballot_sync(0xffffffffu, split);
                if (lane == 0) s_w[warp] = (uint32_t)__popc(mine);
                __syncthreads();
                uint32_t before = n_split, all = 0;
#pragma unroll
                for (int k = 0; k < LV_WARPS; ++k) {
                    const uint32_t v = s_w[k];
                    if (k < warp) before += v;
                    all += v;
                }
                if (split) s_list[before + __popc(mine & lt)] = t;
                n_split += all;
                __syncthreads();
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) leaves += __shfl_xor_sync(0xffffffffu, leaves, o);
            if (lane == 0 && leaves) atomicAdd(&st->n_leaves, leaves);

            // ---- eight lanes per listed node: octant starts, child count ----
            for (uint32_t g0 = 0; g0 < n_split; g0 += LV_UNROLL * (LV_THREADS / 8)) {
                uint32_t bb[LV_UNROLL], hi[LV_UNROLL], end[LV_UNROLL];
#pragma unroll
                for (int u = 0; u < LV_UNROLL; ++u) {
                    const uint32_t g = g0 + u * (LV_THREADS / 8) + group;
                    bb[u] = hi[u] = end[u] = 0;
                    if (g < n_split) {
                        const uint4 tail = __ldcg(reinterpret_cast<const uint4*>(nodes + first + s_list[g]) + 3);
                        end[u] = tail.y + tail.z;
                        bb[u] = tail.y;
                        hi[u] = (oct == 0) ? tail.y : end[u];
                    }
                }
                // LV_UNROLL searches side by side (independent loads in flight): first position in [bb, hi) whose octant
                // at `shift` is >= oct
                for (bool more = true; more;) {
                    more = false;
#pragma unroll
                    for (int u = 0; u < LV_UNROLL; ++u) {
                        if (bb[u] < hi[u]) {
                            const uint32_t mid = bb[u] + ((hi[u] - bb[u]) >> 1);
                            const uint32_t o = (uint32_t)((lk[mid] >> shift) & 7u);
                            if (o < oct) bb[u] = mid + 1; else hi[u] = mid;
                            more = true;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < LV_UNROLL; ++u) {
                    const uint32_t g = g0 + u * (LV_THREADS / 8) + group;
                    if (g < n_split) s_b[g * 8 + oct] = bb[u];
                    const uint32_t next_b = __shfl_down_sync(0xffffffffu, bb[u], 1);
                    const uint32_t ee = (oct == 7) ? end[u] : next_b;
                    const uint32_t any = __ballot_sync(0xffffffffu, g < n_split && ee > bb[u]);
                    if (g < n_split && oct == 0) s_cnt[g] = (uint32_t)__popc((any >> (lane & ~7)) & 0xFFu);
                }
            }
            __syncthreads();

            // ---- exclusive prefix of the child counts over the list; the tile's total ----
            uint32_t total = 0;
            {
                uint32_t v[LV_ROUNDS], sum = 0;
#pragma unroll
                for (int k = 0; k < LV_ROUNDS; ++k) {
                    const uint32_t g = threadIdx.x * LV_ROUNDS + k;
                    v[k] = g < n_split ? s_cnt[g] : 0u;
                    sum += v[k];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += u;
                }
                if (lane == 31) s_w[warp] = incl;
                __syncthreads();
                uint32_t run = incl - sum;
#pragma unroll
                for (int k = 0; k < LV_WARPS; ++k) {
                    const uint32_t u = s_w[k];
                    if (k < warp) run += u;
                    total += u;
                }
#pragma unroll
                for (int k = 0; k < LV_ROUNDS; ++k) {
                    const uint32_t g = threadIdx.x * LV_ROUNDS + k;
                    if (g < n_split) s_cnt[g] = run;
                    run += v[k];
                }
            }

            // ---- children of all tiles before this one (warp 0 looks back, 32 tiles at a time) ----
            if (warp == 0) {
                if (lane == 0) st_relaxed_u64(status + tile, tag | ((tile == 0 ? 2ull : 1ull) << 32) | total);
                uint32_t excl = 0;
                if (tile > 0) {
                    int64_t base = (int64_t)tile - 1;
                    for (;;) {
                        const int64_t j = base - lane;
                        uint64_t v = tag | (2ull << 32);  // before tile 0: an empty inclusive prefix
                        if (j >= 0) {
                            do {
                                v = ld_relaxed_u64(status + j);
                            } while ((v >> 34) != (uint64_t)(level + 1) || ((v >> 32) & 3u) == 0);
                        }
                        const uint32_t incl_lanes = __ballot_sync(0xffffffffu, ((v >> 32) & 3u) == 2u);
                        const int stop = incl_lanes ? __ffs(incl_lanes) - 1 : 31;
                        uint32_t add = lane <= stop ? (uint32_t)v : 0u;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
                        excl += add;
                        if (incl_lanes) break;
                        base -= 32;
                    }
                    if (lane == 0) st_relaxed_u64(status + tile, tag | (2ull << 32) | (excl + total));
                }
                if (lane == 0) {
                    s_excl = excl;
                    s_total = total;
                    if (tile == n_tiles - 1) st->level_total[level] = excl + total;
                }
            }
            __syncthreads();
            const uint32_t excl = s_excl;
            const uint64_t want = (uint64_t)next_first + excl + s_total;
            const bool fits = want <= cap_nodes;
            if (threadIdx.x == 0) {
                atomicMax(&st->needed, (uint32_t)min(want, (uint64_t)0xFFFFFFFFu));
                if (!fits) {
                    atomicOr(&st->level_flags[level], LV_FLAG_CAPACITY);
                    atomicOr(&st->flags, LV_FLAG_CAPACITY);
                }
            }

            // ---- eight lanes per listed node: the children ----
            if (fits) {
                for (uint32_t g0 = 0; g0 < n_split; g0 += LV_THREADS / 8) {  // (same trip count for every lane: ballot inside)
                    const uint32_t g = g0 + group;
                    const bool have = g < n_split;
                    const uint32_t t = have ? s_list[g] : 0u;
                    Node* nd = nodes + first + t;
                    uint32_t bb = 0, ee = 0;
                    if (have) {
                        const uint4 tail = __ldcg(reinterpret_cast<const uint4*>(nd) + 3);
                        bb = s_b[g * 8 + oct];
                        ee = (oct == 7) ? tail.y + tail.z : s_b[g * 8 + oct + 1];
                    }
                    const uint32_t any = __ballot_sync(0xffffffffu, ee > bb);
                    if (!have) continue;
                    const uint32_t mask = (any >> (lane & ~7)) & 0xFFu;
                    const uint32_t c0 = next_first + excl + s_cnt[g];
                    if (ee > bb) {
                        const uint32_t k = (uint32_t)__popc(mask & ((1u << oct) - 1u));
                        const double2* box = reinterpret_cast<const double2*>(nd);
                        const double2 q0 = __ldcg(box), q1 = __ldcg(box + 1), q2 = __ldcg(box + 2);
                        const double lo[3] = {q0.x, q0.y, q1.x}, hi[3] = {q1.y, q2.x, q2.y};
                        Node ch;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {  // octree.cpp:97-99, 115-120
                            const double mid = dmul(dadd(lo[a], hi[a]), 0.5);
                            const bool up = (oct >> a) & 1u;
                            ch.lo[a] = up ? mid : lo[a];
                            ch.hi[a] = up ? hi[a] : mid;
                        }
                        ch.child0 = 0;
                        ch.pt0 = bb;
                        ch.npts = ee - bb;
                        ch.meta = ((uint32_t)(level + 1) << 8);
                        nodes[c0 + k] = ch;
                        parent[c0 + k] = first + t;
                        if (cell) {  // integer cell coordinates of the child at its depth, 21 bits per axis
                            const uint64_t pc = __ldcg(cell + first + t);
                            const uint64_t cx = ((pc & 0x1FFFFFull) << 1) | (oct & 1u);
                            const uint64_t cy = (((pc >> 21) & 0x1FFFFFull) << 1) | ((oct >> 1) & 1u);
                            const uint64_t cz = (((pc >> 42) & 0x1FFFFFull) << 1) | ((oct >> 2) & 1u);
                            cell[c0 + k] = cx | (cy << 21) | (cz << 42);
                        }
                    }
                    if (oct == 0) {
                        nd->child0 = c0;
                        nd->meta = ((uint32_t)level << 8) | mask;
                    }
                }
            }
            __syncthreads();  // the shared lists are reused by the next tile
        }
        grid_barrier(&st->barrier, gridDim.x * (++epoch));
#ifdef ICPB_LVTIME
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            st->t_level[level + 1] = t1;
            st->n_level[level] = count;
        }
#endif
        const uint32_t total = ld_relaxed_u32(&st->level_total[level]);
        const uint32_t flags = ld_relaxed_u32(&st->level_flags[level]);
        if (flags != 0 || total == 0) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                st->depth = (uint32_t)level;
                st->n_nodes = next_first;
            }
            break;
        }
        first = next_first;
        count = total;
    }
}

__global__ void root_node_kernel(Node* __restrict__ nodes, const double* __restrict__ root, uint32_t n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Node r;
    for (int a = 0; a < 3; ++a) {
        r.lo[a] = root[a];
        r.hi[a] = root[3 + a];
    }
    r.child0 = 0;
    r.pt0 = 0;
    r.npts = n;
    r.meta = 0;
    nodes[0] = r;
}

__global__ void inv_perm_kernel(const TPoint* __restrict__ pts, int64_t n, uint32_t* __restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[(uint32_t)pts[i].idx] = (uint32_t)i;
}

// forget a tree's contents but keep its device buffers for the next build of this handle
static void tree_reset(DeviceOctree& t) {
    DeviceOctree keep;
    keep.nodes = t.nodes; keep.parent = t.parent; keep.cell = t.cell; keep.cap_nodes = t.cap_nodes; keep.cap_cell = t.cap_cell;
    keep.pts = t.pts; keep.cap_pts = t.cap_pts;
    keep.inv_perm = t.inv_perm; keep.cap_inv = t.cap_inv;
    keep.grid = t.grid; keep.cap_grid = t.cap_grid;
    keep.full_keys = t.full_keys;
    t = keep;
}

static void tree_free(DeviceOctree& t) {
    if (t.nodes) cudaFree(t.nodes);
    if (t.parent) cudaFree(t.parent);
    if (t.pts) cudaFree(t.pts);
    if (t.inv_perm) cudaFree(t.inv_perm);
    if (t.cell) cudaFree(t.cell);
    if (t.grid) cudaFree(t.grid);
    t = DeviceOctree();
}

void octree_free(Ctx* c) {
    tree_free(c->tree);
    tree_free(c->fast);
}

static int grow_nodes(Ctx* c, DeviceOctree& t, int64_t need) {
    if (need <= t.cap_nodes && (!t.want_cell || need <= t.cap_cell)) return ICP_OK;
    int64_t cap = std::max<int64_t>(need + need / 2, 1024);
    Node* nn = nullptr;
    uint32_t* np = nullptr;
    uint64_t* nc = nullptr;
    ICPB_CUDA(c, cudaMalloc(&nn, (size_t)cap * sizeof(Node)));
    ICPB_CUDA(c, cudaMalloc(&np, (size_t)cap * sizeof(uint32_t)));
    if (t.want_cell) ICPB_CUDA(c, cudaMalloc(&nc, (size_t)cap * sizeof(uint64_t)));
    if (t.nodes) {
        ICPB_CUDA(c, cudaMemcpyAsync(nn, t.nodes, (size_t)t.n_nodes * sizeof(Node), cudaMemcpyDeviceToDevice, c->stream));
        ICPB_CUDA(c, cudaMemcpyAsync(np, t.parent, (size_t)t.n_nodes * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        if (nc) ICPB_CUDA(c, cudaMemcpyAsync(nc, t.cell, (size_t)t.n_nodes * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
        ICPB_CUDA(c, cudaStreamSynchronize(c->stream));
        ICPB_CUDA(c, cudaFree(t.nodes));
        ICPB_CUDA(c, cudaFree(t.parent));
        if (t.cell) ICPB_CUDA(c, cudaFree(t.cell));
    }
    t.nodes = nn;
    t.parent = np;
    t.cell = nc;
    t.cap_nodes = cap;
    t.cap_cell = nc ? cap : 0;
    return ICP_OK;
}

// Builds one linear octree over the m target points at d_xyz.  cubic == false: the reference's tree (root = the
// cloud's bounding box, octree.cpp:41-126).  cubic == true: the SEARCH tree -- same construction over a cubic root,
// so its cells are cubes; it only has to bound its points (every point lies in the closed box of its leaf), not
// to match anything in the reference.
static int build_tree(Ctx* c, DeviceOctree& t, const double* d_xyz, int64_t m, int max_pts, int max_depth, bool cubic) {
    tree_reset(t);  // keeps the device allocations of an earlier build (grow-only), forgets its contents
    t.want_cell = cubic;
    t.max_pts = max_pts;
    t.max_depth = max_depth;
    t.n_pts = m;
    cudaStream_t s = c->stream;

    // K0
    const int bb_blocks = (int)std::min<int64_t>((m + BBOX_THREADS - 1) / BBOX_THREADS, (int64_t)c->sm_count * 8);
    ICPB_TRY(devbuf_reserve(c, c->scratch0, (size_t)(bb_blocks + 1) * 6 * sizeof(double) + 64));
    double* d_part = (double*)c->scratch0.p;
    double* d_root = d_part + (size_t)bb_blocks * 6;
    bbox_partial_kernel<<<bb_blocks, BBOX_THREADS, 0, s>>>(d_xyz, m, d_part);
    bbox_finish_kernel<<<1, 192, 0, s>>>(d_part, bb_blocks, d_xyz, d_root);
    c->launches += 2;
    if (cubic) {
        cube_root_kernel<<<1, 32, 0, s>>>(d_root);
        c->launches++;
    }

    uint32_t* d_misc = nullptr;  // [0] pos_of_idx0
    ICPB_TRY(devbuf_reserve(c, c->part_a, 4096));
    d_misc = (uint32_t*)c->part_a.p;
    LevelState* d_state = (LevelState*)(d_misc + 64);
    LevelState h_state;
    if (c->lv_grid == 0) {  // every block of the level kernel has to be resident
        int per_sm = 0;
        ICPB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, octree_levels_kernel, LV_THREADS, 0));
        c->lv_grid = std::max(1, per_sm) * c->sm_count;
    }

    // K1
    uint64_t *keys = nullptr, *keys_alt = nullptr;
    uint32_t *idx = nullptr, *idx_alt = nullptr;
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)m * (2 * sizeof(uint64_t) + 2 * sizeof(uint32_t)) + 1024));
    const int kb = (int)((m + 255) / 256);
    const int n_words = std::max(1, (max_depth + KEY_LEVELS - 1) / KEY_LEVELS);
    // Octant keys are taken (and sorted) only KEY_DEPTH_FIRST levels deep at first: real clouds stop splitting well above
    // max_depth, the points of a leaf need no order among themselves (leaf scans break ties by original index), and every
    // 8 key bits cost a pass over the points.  If some node still has to split below that depth the build starts over with
    // keys of the full depth, and the tree remembers it for its next build.
    int key_depth = (n_words == 1 && !t.full_keys) ? std::min(max_depth, KEY_DEPTH_FIRST) : max_depth;
    int level = 0;
    for (;;) {
        keys = (uint64_t*)c->scratch1.p;
        keys_alt = keys + m;
        idx = (uint32_t*)(keys_alt + m);
        idx_alt = idx + m;
        auto word_levels = [&](int w) { return std::min(KEY_LEVELS, key_depth - KEY_LEVELS * w); };
        uint64_t* words_sorted = nullptr;  // deep trees: [n_words][m], sorted order
        if (n_words == 1) {
            morton_keys_kernel<<<kb, 256, 0, s>>>(d_xyz, m, d_root, key_depth, keys, idx);
            c->launches++;
            // K2
            ICPB_TRY(radix_sort_pairs(c, keys, keys_alt, idx, idx_alt, m, 3 * key_depth));
        } else {
            // deeper than one key word: least significant word first, each pass stable, the permutation carried along
            ICPB_TRY(devbuf_reserve(c, c->scratch_keys, (size_t)m * sizeof(uint64_t) * 2 * n_words));
            uint64_t* words = (uint64_t*)c->scratch_keys.p;  // [n_words][m], caller order
            words_sorted = words + (size_t)n_words * m;
            morton_keys_kernel<<<kb, 256, 0, s>>>(d_xyz, m, d_root, max_depth, words, idx);
            c->launches++;
            for (int w = n_words - 1; w >= 0; --w) {
                gather_u64_kernel<<<kb, 256, 0, s>>>(words + (size_t)w * m, idx, m, keys);
                c->launches++;
                ICPB_TRY(radix_sort_pairs(c, keys, keys_alt, idx, idx_alt, m, 3 * word_levels(w)));
            }
            for (int w = 0; w < n_words; ++w) {
                gather_u64_kernel<<<kb, 256, 0, s>>>(words + (size_t)w * m, idx, m, words_sorted + (size_t)w * m);
                c->launches++;
            }
        }

        // sorted points
        if (t.cap_pts < m) {
            if (t.pts) ICPB_CUDA(c, cudaFree(t.pts));
            t.pts = nullptr;
            t.cap_pts = 0;
            ICPB_CUDA(c, cudaMalloc(&t.pts, (size_t)m * sizeof(TPoint)));
            t.cap_pts = m;
        }
        ICPB_CUDA(c, cudaMemsetAsync(d_misc, 0, 64, s));
        gather_points_kernel<<<kb, 256, 0, s>>>(d_xyz, idx, m, t.pts, d_misc);
        c->launches++;

        // K3
        ICPB_TRY(grow_nodes(c, t, std::max<int64_t>(m / 2, 1024)));
        ICPB_CUDA(c, cudaMemsetAsync(t.parent, 0xFF, sizeof(uint32_t), s));
        if (t.cell) ICPB_CUDA(c, cudaMemsetAsync(t.cell, 0, sizeof(uint64_t), s));
        root_node_kernel<<<1, 32, 0, s>>>(t.nodes, d_root, (uint32_t)m);
        c->launches++;
        t.n_nodes = 1;
        // the level loop, on the device.  A node table that turns out too small is grown and the levels run again.
        const uint64_t* level_keys = words_sorted ? words_sorted : keys;
        const size_t status_bytes = ((size_t)m / LV_TILE + (size_t)c->lv_grid + 2) * sizeof(uint64_t);  // one word per tile of a level
        ICPB_TRY(devbuf_reserve(c, c->scratch3, status_bytes));
        uint64_t* status = (uint64_t*)c->scratch3.p;
        bool deeper_keys = false;
        for (;;) {
            ICPB_CUDA(c, cudaMemsetAsync(d_state, 0, sizeof(LevelState), s));
            ICPB_CUDA(c, cudaMemsetAsync(status, 0, status_bytes, s));
            Node* a_nodes = t.nodes;
            uint32_t* a_parent = t.parent;
            uint64_t* a_cell = t.cell;
            uint32_t a_cap = (uint32_t)std::min<int64_t>(t.cap_nodes, 0xFFFFFFFFll);
            int64_t a_m = m;
            int a_words = n_words, a_kd = key_depth, a_maxpts = max_pts, a_maxdepth = max_depth;
            void* args[] = {&a_nodes, &a_parent, &a_cell, &a_cap, &level_keys, &a_m, &a_words, &a_kd, &a_maxpts, &a_maxdepth, &status, &d_state};
            ICPB_CUDA(c, cudaLaunchCooperativeKernel((const void*)octree_levels_kernel, dim3(c->lv_grid), dim3(LV_THREADS), args, 0, s));
            c->launches++;
            ICPB_CUDA(c, cudaMemcpyAsync(&h_state, d_state, sizeof(LevelState), cudaMemcpyDeviceToHost, s));
            ICPB_CUDA(c, cudaStreamSynchronize(s));
            if (h_state.flags & LV_FLAG_DEEPER) {
                deeper_keys = true;
                break;
            }
            if (!(h_state.flags & LV_FLAG_CAPACITY)) break;
            if (t.cap_nodes >= 0xFFFFFFFFll) {
                c->err = "octree: more than 2^32 nodes";
                return ICP_INVALID_ARGUMENT;
            }
            ICPB_TRY(grow_nodes(c, t, std::max<int64_t>((int64_t)h_state.needed, t.cap_nodes * 2)));
            ICPB_CUDA(c, cudaMemsetAsync(t.parent, 0xFF, sizeof(uint32_t), s));
            if (t.cell) ICPB_CUDA(c, cudaMemsetAsync(t.cell, 0, sizeof(uint64_t), s));
            root_node_kernel<<<1, 32, 0, s>>>(t.nodes, d_root, (uint32_t)m);
        }
        level = (int)h_state.depth;
        t.n_nodes = h_state.n_nodes;
#ifdef ICPB_LVTIME
        for (int l = 0; l <= level; ++l)
            fprintf(stderr, "[icp_b200] level %2d: %8u nodes %8.1f us\n", l, h_state.n_level[l], (double)(h_state.t_level[l + 1] - h_state.t_level[l]) / 1000.0);
#endif
        if (!deeper_keys) break;
        t.full_keys = true;
        key_depth = max_depth;
    }
    t.depth = level;
    uint32_t misc[1];
    double root[6];
    ICPB_CUDA(c, cudaMemcpyAsync(misc, d_misc, sizeof misc, cudaMemcpyDeviceToHost, s));
    ICPB_CUDA(c, cudaMemcpyAsync(root, d_root, sizeof root, cudaMemcpyDeviceToHost, s));
    ICPB_CUDA(c, cudaStreamSynchronize(s));
    ICPB_CUDA(c, cudaGetLastError());
    t.pos_of_idx0 = misc[0];
    t.n_leaves = h_state.n_leaves;
    for (int a = 0; a < 3; ++a) {
        t.root_lo[a] = root[a];
        t.root_hi[a] = root[3 + a];
    }
    t.valid = true;
    return ICP_OK;
}

static int build_inv_perm_of(Ctx* c, DeviceOctree& t) {
    if (t.inv_valid) return ICP_OK;
    if (t.cap_inv < t.n_pts) {
        if (t.inv_perm) ICPB_CUDA(c, cudaFree(t.inv_perm));
        t.inv_perm = nullptr;
        t.cap_inv = 0;
        ICPB_CUDA(c, cudaMalloc(&t.inv_perm, (size_t)t.n_pts * sizeof(uint32_t)));
        t.cap_inv = t.n_pts;
    }
    t.inv_valid = true;
    inv_perm_kernel<<<(int)((t.n_pts + 255) / 256), 256, 0, c->stream>>>(t.pts, t.n_pts, t.inv_perm);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// Entry grid of the search tree: a dense array over the cloud's bounding box whose cells are the search tree's
// cells at one level L; each entry holds the node that owns the cell (the depth-L node, or the shallower leaf
// covering it), NONE where there is no point.  Lets a query jump to the few cells its search ball touches
// instead of walking down from the root or up from a previous leaf.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) leaf_depth_hist_kernel(const Node* __restrict__ nodes, int64_t n_nodes,
                                                              unsigned long long* __restrict__ hist /* [96] */) {
    // [0,32): points in leaves per depth; [32,64): nodes per depth; [64,96): points that reach each depth.
    // Nodes are numbered level by level, so a thread that owns a CONTIGUOUS run of nodes sees one depth (two at a level
    // boundary): it sums in registers and touches the block's shared histogram once per depth, the block touches the global
    // one once per bin.  (Three shared 64-bit atomics per node on the same address cost 0.74 ms at 6.6 M nodes.)
    __shared__ unsigned long long sh[96];
    if (threadIdx.x < 96) sh[threadIdx.x] = 0ull;
    __syncthreads();
    const int64_t n_thr = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (n_nodes + n_thr - 1) / n_thr;
    const int64_t b = tid * per, e = min(n_nodes, b + per);
    uint32_t cur = 0xFFFFFFFFu;
    unsigned long long leaf_pts = 0ull, cnt = 0ull, pts = 0ull;
    for (int64_t i = b; i < e; ++i) {
        const uint32_t meta = nodes[i].meta;
        const uint32_t npts = nodes[i].npts;
        const uint32_t d = (meta >> 8) & 0x1Fu;
        if (d != cur) {
            if (cur != 0xFFFFFFFFu) {
                if (leaf_pts) atomicAdd(&sh[cur], leaf_pts);
                atomicAdd(&sh[32 + cur], cnt);
                atomicAdd(&sh[64 + cur], pts);
            }
            cur = d;
            leaf_pts = cnt = pts = 0ull;
        }
        if ((meta & 0xFFu) == 0u) leaf_pts += npts;
        cnt += 1ull;
        pts += npts;
    }
    if (cur != 0xFFFFFFFFu) {
        if (leaf_pts) atomicAdd(&sh[cur], leaf_pts);
        atomicAdd(&sh[32 + cur], cnt);
        atomicAdd(&sh[64 + cur], pts);
    }
    __syncthreads();
    if (threadIdx.x < 96 && sh[threadIdx.x] != 0ull) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void __launch_bounds__(128) grid_fill_kernel(const Node* __restrict__ nodes, const uint64_t* __restrict__ cell,
                                                        int64_t n_nodes, int level, int nx, int ny, int nz, int range_max,
                                                        uint2* __restrict__ grid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const Node nd = nodes[i];
    const int d = (int)((nd.meta >> 8) & 0xFFu);
    const bool leaf = (nd.meta & 0xFFu) == 0u;
    if (!(d == level || (leaf && d < level))) return;
    // entry kinds, see nn_common.cuh (cell walk).  The points below any node are contiguous in the sorted order, so a
    // cell with few points is entered as a plain range whether its node is a leaf or not; crowded cells keep the node.
    uint2 e;
    if (nd.npts < (1u << 24) && (leaf || nd.npts <= (uint32_t)range_max))
        e = make_uint2(nd.pt0, nd.npts | ((uint32_t)d << 24) | ((d == level ? 1u : 2u) << 30));
    else if (d == level)
        e = make_uint2((uint32_t)i, 3u << 30);
    else
        return;
    const uint64_t pc = cell[i];
    const int sh = level - d;
    const long long x0 = (long long)(pc & 0x1FFFFFull) << sh, y0 = (long long)((pc >> 21) & 0x1FFFFFull) << sh,
                    z0 = (long long)((pc >> 42) & 0x1FFFFFull) << sh;
    const long long bs = 1ll << sh;
    const long long x1 = min(x0 + bs, (long long)nx), y1 = min(y0 + bs, (long long)ny), z1 = min(z0 + bs, (long long)nz);
    for (long long z = z0; z < z1; ++z)
        for (long long y = y0; y < y1; ++y)
            for (long long x = x0; x < x1; ++x) grid[(z * ny + y) * nx + x] = e;
}

static int build_grid(Ctx* c, DeviceOctree& t) {
    cudaStream_t s = c->stream;
    ICPB_TRY(devbuf_reserve(c, c->scratch0, 96 * sizeof(unsigned long long)));
    unsigned long long* d_hist = (unsigned long long*)c->scratch0.p;
    ICPB_CUDA(c, cudaMemsetAsync(d_hist, 0, 96 * sizeof(unsigned long long), s));
    leaf_depth_hist_kernel<<<(int)std::min<int64_t>((t.n_nodes + 255) / 256, (int64_t)c->sm_count * 8), 256, 0, s>>>(t.nodes, t.n_nodes, d_hist);
    c->launches++;
    unsigned long long hist[96];
    ICPB_CUDA(c, cudaMemcpyAsync(hist, d_hist, sizeof hist, cudaMemcpyDeviceToHost, s));
    ICPB_CUDA(c, cudaStreamSynchronize(s));
    // Levels by mean occupancy (points that reach a depth / nodes at that depth), among depths most points reach:
    // the BASE level is the deepest one whose cells still hold several points (cheapest when the search ball is a good
    // fraction of a cell); finer levels, down to ~1.5 points per cell, serve small balls (a converged registration).
    int base = 0, fine = 0;
    for (int d = 0; d <= 21; ++d) {
        if (hist[32 + d] == 0) break;
        if (2 * hist[64 + d] < (unsigned long long)t.n_pts) break;  // most points sit in shallower leaves
        const double occ = (double)hist[64 + d] / (double)hist[32 + d];
        if (occ >= c->opt_base_occupancy) base = d;
        if (occ >= 1.5) fine = d;
    }
    base = std::max(std::min(base + c->opt_grid_shift, 21), 0);
    const int want_levels = std::min(std::max(c->opt_grid_levels, 1), 4);
    fine = std::max(std::min(fine, base + want_levels - 1), base);
    // ... all lowered until the pyramid of dense arrays over the bounding box stays within the entry budget
    t.cube = t.root_hi[0] - t.root_lo[0];
    double ext[3];
    for (int a = 0; a < 3; ++a) ext[a] = c->tree.root_hi[a] - t.root_lo[a];  // the cloud's own extent (reference root box)
    auto dims = [&](int level, long long* n) {
        const double g = t.cube / (double)(1ll << level);
        double tot = 1.0;
        for (int a = 0; a < 3; ++a) {
            n[a] = (long long)(ext[a] / g) + 1;
            tot *= (double)n[a];
        }
        return tot;
    };
    // entry budget: 2^28 (2 GiB) for small clouds, 64 entries per point for large ones (a 10^8-point aerial tile needs
    // 3.2 G entries = 26 GB for the two levels its density calls for), never above grid_max_cells
    const double budget = std::min((double)c->opt_grid_max_cells, std::max(268435456.0, 64.0 * (double)t.n_pts));
    int nlev = 1;
    for (;;) {
        nlev = fine - base + 1;
        double tot = 0.0;
        long long n[3];
        for (int k = 0; k < nlev; ++k) tot += dims(base + k, n);
        if (tot <= budget || fine == 0) break;
        if (fine > base) --fine; else { --fine; --base; }
    }
    base = std::max(base, 0);
    fine = std::max(fine, base);
    // coarser levels below the base one cost an eighth of it each; the pyramid holds four levels at most
    const int coarse = std::min(std::min(std::max(c->opt_grid_coarse, 0), base), 4 - (fine - base + 1));
    nlev = fine - base + 1 + coarse;
    t.glev_n = nlev;
    t.glev_min = base - coarse;
    t.gbase = coarse;
    {
        const double occ = hist[32 + base] ? (double)hist[64 + base] / (double)hist[32 + base] : 1.0;
        t.spacing = t.cube / (double)(1ll << base) / std::sqrt(std::max(occ, 1.0));
    }
    long long total = 0;
    for (int k = 0; k < nlev; ++k) {
        long long n[3];
        dims(t.glev_min + k, n);
        t.goff[k] = total;
        for (int a = 0; a < 3; ++a) t.gdim[k][a] = (int)n[a];
        total += n[0] * n[1] * n[2];
    }
    if (t.cap_grid < total) {
        if (t.grid) ICPB_CUDA(c, cudaFree(t.grid));
        t.grid = nullptr;
        t.cap_grid = 0;
        ICPB_CUDA(c, cudaMalloc(&t.grid, (size_t)total * sizeof(uint2)));
        t.cap_grid = total;
    }
    ICPB_CUDA(c, cudaMemsetAsync(t.grid, 0, (size_t)total * sizeof(uint2), s));
    for (int k = 0; k < nlev; ++k) {
        grid_fill_kernel<<<(int)((t.n_nodes + 127) / 128), 128, 0, s>>>(t.nodes, t.cell, t.n_nodes, t.glev_min + k, t.gdim[k][0],
                                                                        t.gdim[k][1], t.gdim[k][2], c->opt_range_max,
                                                                        t.grid + t.goff[k]);
        c->launches++;
    }
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

// The reference's tree (c->tree: structure parity, literal traversal) and the isotropic search tree (c->fast: every
// fast search path, and the canonical point order that match positions refer to).
int octree_build_device(Ctx* c, const double* d_xyz, int64_t m, int max_pts, int max_depth) {
    c->tree.valid = false;
    c->fast.valid = false;
    c->prev_valid = false;
    if (m <= 0) return ICP_EMPTY_INPUT;
    if (max_depth < 0 || max_depth > 63 || m > 0x7fffffffLL) {
        c->err = "octree: max_depth must be in [0,63] and n_tgt < 2^31";
        return ICP_INVALID_ARGUMENT;
    }
    ICPB_TRY(build_tree(c, c->tree, d_xyz, m, max_pts, max_depth, false));
    ICPB_TRY(build_tree(c, c->fast, d_xyz, m, c->opt_search_leaf, c->opt_search_depth, true));
    ICPB_TRY(build_inv_perm_of(c, c->fast));  // original index -> search-tree position (literal results, stage API)
    ICPB_TRY(build_grid(c, c->fast));
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// Query ordering: the NN kernel runs one query per thread, so neighbouring threads should walk the same
// nodes.  Queries are ordered by a 3x21-bit Morton code of their own bounding box (this is only a
// permutation for locality -- any order gives the same per-query answer).
// ------------------------------------------------------------------------------------------------
// 13 bits per axis (8192^3 cells: centimetres on a 100 m tile, ~0.4 m on a 3 km one) order the queries finely enough for
// neighbouring threads to share cells, and keep the sort at 5 radix passes instead of 8.
constexpr int QKEY_BITS = 13;
constexpr uint32_t QKEY_MAX = (1u << QKEY_BITS) - 1u;

__global__ void __launch_bounds__(256) query_keys_kernel(const double* __restrict__ xyz, int64_t n,
                                                         const double* __restrict__ box, uint64_t* __restrict__ keys,
                                                         uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // ISOTROPIC cells: one scale (the largest extent) for all three axes, so that 32 consecutive queries form a
    // compact, roughly cubic clump (a per-axis scale would slice 2.5-D scenes into thin height slabs whose tiles
    // follow contour lines).
    const double ext = fmax(fmax(box[3] - box[0], box[4] - box[1]), box[5] - box[2]);
    const double inv = ext > 0.0 ? (double)QKEY_MAX / ext : 0.0;
    uint64_t key = 0;
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double f = (xyz[3 * i + a] - box[a]) * inv;
        if (!(f > 0.0)) f = 0.0;  // also catches NaN
        if (f > (double)QKEY_MAX) f = (double)QKEY_MAX;
        q[a] = (uint32_t)f;
    }
    for (int b = QKEY_BITS - 1; b >= 0; --b)
        key = (key << 3) | (((q[0] >> b) & 1u)) | (((q[1] >> b) & 1u) << 1) | (((q[2] >> b) & 1u) << 2);
    keys[i] = key;
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) gather_soa_kernel(const double* __restrict__ xyz, const uint32_t* __restrict__ idx,
                                                         int64_t n, double* __restrict__ sx, double* __restrict__ sy,
                                                         double* __restrict__ sz, uint32_t* __restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = idx[i];
    sx[i] = xyz[3 * (int64_t)j];
    sy[i] = xyz[3 * (int64_t)j + 1];
    sz[i] = xyz[3 * (int64_t)j + 2];
    perm[i] = j;
}

int order_queries(Ctx* c, const double* d_q, int64_t n, double* sx, double* sy, double* sz, uint32_t* perm) {
    cudaStream_t s = c->stream;
    const int bb_blocks = (int)std::min<int64_t>((n + BBOX_THREADS - 1) / BBOX_THREADS, (int64_t)c->sm_count * 8);
    ICPB_TRY(devbuf_reserve(c, c->scratch0, (size_t)(bb_blocks + 1) * 6 * sizeof(double) + 64));
    double* d_part = (double*)c->scratch0.p;
    double* d_box = d_part + (size_t)bb_blocks * 6;
    bbox_partial_kernel<<<bb_blocks, BBOX_THREADS, 0, s>>>(d_q, n, d_part);
    bbox_finish_kernel<<<1, 192, 0, s>>>(d_part, bb_blocks, d_q, d_box);
    ICPB_TRY(devbuf_reserve(c, c->scratch1, (size_t)n * (2 * sizeof(uint64_t) + 2 * sizeof(uint32_t)) + 1024));
    uint64_t* keys = (uint64_t*)c->scratch1.p;
    uint64_t* keys_alt = keys + n;
    uint32_t* idx = (uint32_t*)(keys_alt + n);
    uint32_t* idx_alt = idx + n;
    const int kb = (int)((n + 255) / 256);
    query_keys_kernel<<<kb, 256, 0, s>>>(d_q, n, d_box, keys, idx);
    c->launches += 3;
    ICPB_TRY(radix_sort_pairs(c, keys, keys_alt, idx, idx_alt, n, 3 * QKEY_BITS));
    gather_soa_kernel<<<kb, 256, 0, s>>>(d_q, idx, n, sx, sy, sz, perm);
    c->launches++;
    ICPB_CUDA(c, cudaGetLastError());
    return ICP_OK;
}

int build_inv_perm(Ctx* c) { return build_inv_perm_of(c, c->fast); }

}  // namespace icpb
