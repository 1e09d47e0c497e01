// prims.cu -- HBM-bound integer primitives of the rating-matrix build: histogram, exclusive scan,
// stable LSD radix sort (8 bits per pass), gather. All hand-written; no CUB/Thrust.
//
// These replace the reference's single-threaded passes over IList<int>
// (Data/DataSet.cs:134-191 count / index build, MultiCore.cs:58-66 block bucketing).
#include "common.cuh"
#include <cstdarg>
#include <thread>

namespace mml {

// ---- thread-local error text --------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

int32_t on_ranks(int n, const std::function<int32_t(int)>& fn)
{
    if (n == 1) return fn(0);
    std::vector<int32_t> st((size_t)n, MML_OK);
    std::vector<std::string> msg((size_t)n);
    std::vector<std::thread> th;
    th.reserve((size_t)n);
    for (int r = 0; r < n; r++)
        th.emplace_back([&, r] {
            st[(size_t)r] = fn(r);
            if (st[(size_t)r] != MML_OK) msg[(size_t)r] = last_error();
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < n; r++)
        if (st[(size_t)r] != MML_OK) { set_error("[gpu %d of %d] %s", r, n, msg[(size_t)r].c_str()); return st[(size_t)r]; }
    return MML_OK;
}

int bits_for(uint32_t max_value)
{
    int b = 1;
    while (b < 32 && (max_value >> b) != 0) b++;
    return b;
}

// ---- iota / gather / histogram ----------------------------------------------------------------
__global__ void iota_kernel(uint32_t* __restrict__ v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) v[i] = (uint32_t)i;
}

__global__ void gather_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx,
                              uint32_t* __restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = src[idx[i]];
}

__global__ void histogram_kernel(const int32_t* __restrict__ ids, int64_t n, uint32_t* __restrict__ counts)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) atomicAdd(&counts[ids[i]], 1u);
}

static inline int grid_for(int64_t n, int threads, int max_blocks = 148 * 16)
{
    int64_t b = ceil_div(n, threads);
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

int32_t iota_u32(uint32_t* vals, int64_t n, cudaStream_t s)
{
    iota_kernel<<<grid_for(n, 256), 256, 0, s>>>(vals, n);
    MML_CUDA(cudaGetLastError());
    return MML_OK;
}

int32_t gather_u32(const uint32_t* src, const uint32_t* idx, uint32_t* out, int64_t n, cudaStream_t s)
{
    gather_kernel<<<grid_for(n, 256), 256, 0, s>>>(src, idx, out, n);
    MML_CUDA(cudaGetLastError());
    return MML_OK;
}

int32_t histogram_i32(const int32_t* ids, int64_t n, uint32_t* counts, cudaStream_t s)
{
    histogram_kernel<<<grid_for(n, 256), 256, 0, s>>>(ids, n, counts);
    MML_CUDA(cudaGetLastError());
    return MML_OK;
}

// ---- exclusive scan ----------------------------------------------------------------------------
// Three-phase scan: per-tile sums -> (recursive) scan of the sums -> per-tile scan + offset.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 2048

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the block total in *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = SCAN_THREADS / 32;
        uint32_t w = (lane < NW) ? warp_sums[lane] : 0u;
        uint32_t winc = warp_inclusive_scan(w);
        if (lane < NW) warp_sums[lane] = winc - w;   // exclusive prefix of the warp totals
        if (lane == NW - 1) *total = winc;
    }
    __syncthreads();
    return warp_sums[warp] + inc - v;
}

__global__ void scan_tile_sums_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ sums)
{
    __shared__ uint32_t total;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n) local += in[i];
    }
    block_exclusive_scan(local, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void scan_tile_apply_kernel(const uint32_t* __restrict__ in, int64_t n,
                                       const uint32_t* __restrict__ tile_offsets, uint32_t* __restrict__ out)
{
    __shared__ uint32_t total;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        v[k] = (i < n) ? in[i] : 0u;
        local += v[k];
    }
    uint32_t run = block_exclusive_scan(local, &total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
    // out[n] = grand total, written by the thread that owns position n
    if (base <= n && n < base + SCAN_ITEMS) out[n] = run;
}

// single-block scan for n <= SCAN_TILE (also the recursion floor)
__global__ void scan_small_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out)
{
    __shared__ uint32_t total;
    const int64_t base = (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        v[k] = (i < n) ? in[i] : 0u;
        local += v[k];
    }
    uint32_t run = block_exclusive_scan(local, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) out[n] = total;
}

int32_t exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s)
{
    if (n <= 0) {
        MML_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t), s));
        return MML_OK;
    }
    if (n <= SCAN_TILE) {
        scan_small_kernel<<<1, SCAN_THREADS, 0, s>>>(in, n, out);
        MML_CUDA(cudaGetLastError());
        return MML_OK;
    }
    int64_t tiles = ceil_div(n + 1, SCAN_TILE);   // +1 so that position n (the total) has an owner tile
    StreamBuf<uint32_t> sums, offsets;          // stream-ordered scratch: no host synchronisation in here
    MML_TRY(sums.alloc((size_t)tiles, s));
    MML_TRY(offsets.alloc((size_t)tiles + 1, s));
    scan_tile_sums_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, n, sums.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(exclusive_scan_u32(sums.p, offsets.p, tiles, s));
    scan_tile_apply_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, n, offsets.p, out);
    MML_CUDA(cudaGetLastError());
    return MML_OK;
}

// ---- stable LSD radix sort (pairs) -----------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;                         // elements per lane
constexpr int RS_STRIP = 32 * RS_ROUNDS;              // contiguous elements per warp
constexpr int RS_TILE = RS_STRIP * RS_WARPS;          // 4096 elements per block

// hist[d * nblk + b] = number of keys of tile b whose digit is d
__global__ void rs_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                               uint32_t* __restrict__ hist, int nblk)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_TILE / RS_THREADS; k++) {
        int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblk + blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                  int64_t n, int shift, const uint32_t* __restrict__ hist_scanned, int nblk)
{
    __shared__ uint32_t cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < RS_WARPS * 256; t += RS_THREADS) (&cnt[0][0])[t] = 0;
    __syncthreads();

    const int64_t strip = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_STRIP;
    uint32_t key[RS_ROUNDS];
    uint16_t rank[RS_ROUNDS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = strip + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0u;
        const uint32_t d = (key[r] >> shift) & 255u;
        // lanes with the same digit form a group; invalid lanes get private groups
        const uint32_t tag = valid ? d : (0x100u | (uint32_t)lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, tag);
        const uint32_t before = __popc(peers & lt_mask);
        uint32_t base = 0;
        if (valid) base = cnt[warp][d];
        __syncwarp();
        if (valid && before == 0) cnt[warp][d] = base + __popc(peers);
        __syncwarp();
        rank[r] = (uint16_t)(base + before);
    }
    __syncthreads();
    // digit d = threadIdx.x: turn per-warp counts into bases (global base of this tile + warps before)
    {
        const int d = threadIdx.x;
        uint32_t running = hist_scanned[(int64_t)d * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = cnt[w][d];
            cnt[w][d] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int64_t i = strip + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & 255u;
            const uint32_t pos = cnt[warp][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = vals_in[i];
        }
    }
}

int32_t radix_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp, uint32_t* vals_tmp,
                         int64_t n, int key_bits, cudaStream_t s)
{
    if (n <= 1) return MML_OK;
    if (n >= ((int64_t)1 << 32)) { set_error("radix_sort_pairs: n too large"); return MML_ERR_ARG; }
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    const int nblk = (int)ceil_div(n, RS_TILE);
    StreamBuf<uint32_t> hist, hist_scanned;
    MML_TRY(hist.alloc((size_t)256 * nblk, s));
    MML_TRY(hist_scanned.alloc((size_t)256 * nblk + 1, s));
    uint32_t *kin = keys, *vin = vals, *kout = keys_tmp, *vout = vals_tmp;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        rs_hist_kernel<<<nblk, RS_THREADS, 0, s>>>(kin, n, shift, hist.p, nblk);
        MML_CUDA(cudaGetLastError());
        MML_TRY(exclusive_scan_u32(hist.p, hist_scanned.p, (int64_t)256 * nblk, s));
        rs_scatter_kernel<<<nblk, RS_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, hist_scanned.p, nblk);
        MML_CUDA(cudaGetLastError());
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (kin != keys) {   // odd number of passes: result sits in the tmp buffers
        MML_CUDA(cudaMemcpyAsync(keys, kin, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        MML_CUDA(cudaMemcpyAsync(vals, vin, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    }
    return MML_OK;
}

}  // namespace mml
