// ratings.cu -- rating-matrix build on device: COO upload, counts, stats, CSR, block partition.
//
// Reference: Data/StaticRatings.cs:47-84 (COO store), Data/DataSet.cs:134-191 (counts, ByUser/ByItem),
// Data/Ratings.cs:76-84 (Average), Data/RatingScale.cs:104-117 (min/max), MultiCore.cs:43-73 (blocks).
#include "common.cuh"
#include <algorithm>
#include <cfloat>
#include <new>

namespace mml {

// ---- stats: double sum + min/max, deterministic two-level reduction ---------------------------
constexpr int ST_THREADS = 256;

__global__ void stats_partial_kernel(const float* __restrict__ v, int64_t n,
                                     double* __restrict__ psum, float* __restrict__ pmin, float* __restrict__ pmax)
{
    __shared__ double ssum[ST_THREADS];
    __shared__ float smin[ST_THREADS], smax[ST_THREADS];
    double sum = 0.0;
    float mn = FLT_MAX, mx = -FLT_MAX;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float x = v[i];
        sum += (double)x;
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    ssum[threadIdx.x] = sum; smin[threadIdx.x] = mn; smax[threadIdx.x] = mx;
    __syncthreads();
    for (int s = ST_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            ssum[threadIdx.x] += ssum[threadIdx.x + s];
            smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + s]);
            smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { psum[blockIdx.x] = ssum[0]; pmin[blockIdx.x] = smin[0]; pmax[blockIdx.x] = smax[0]; }
}

static int32_t compute_stats(Ratings* r)
{
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(r->n, ST_THREADS * 8), 1), 1184);
    DevBuf<double> psum; DevBuf<float> pmin, pmax;
    MML_TRY(psum.alloc(blocks)); MML_TRY(pmin.alloc(blocks)); MML_TRY(pmax.alloc(blocks));
    stats_partial_kernel<<<blocks, ST_THREADS, 0, r->ctx->stream>>>(r->values.p, r->n, psum.p, pmin.p, pmax.p);
    MML_CUDA(cudaGetLastError());
    std::vector<double> hs(blocks); std::vector<float> hmin(blocks), hmax(blocks);
    MML_CUDA(cudaMemcpyAsync(hs.data(), psum.p, sizeof(double) * blocks, cudaMemcpyDeviceToHost, r->ctx->stream));
    MML_CUDA(cudaMemcpyAsync(hmin.data(), pmin.p, sizeof(float) * blocks, cudaMemcpyDeviceToHost, r->ctx->stream));
    MML_CUDA(cudaMemcpyAsync(hmax.data(), pmax.p, sizeof(float) * blocks, cudaMemcpyDeviceToHost, r->ctx->stream));
    MML_CUDA(cudaStreamSynchronize(r->ctx->stream));
    double sum = 0; float mn = FLT_MAX, mx = -FLT_MAX;
    for (int b = 0; b < blocks; b++) { sum += hs[b]; mn = std::min(mn, hmin[b]); mx = std::max(mx, hmax[b]); }
    // Data/Ratings.cs:82: (float) sum / Count -- the cast binds to sum
    r->average = r->n > 0 ? (float)sum / (float)r->n : 0.f;
    r->min_rating = mn; r->max_rating = mx;
    return MML_OK;
}

// ---- block keys for MultiCore.PartitionUsersAndItems -----------------------------------------
__global__ void block_key_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t n,
                                 const int32_t* __restrict__ user_perm, const int32_t* __restrict__ item_perm,
                                 int32_t g, uint32_t* __restrict__ key)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride)
        key[t] = (uint32_t)(user_perm[users[t]] % g) * (uint32_t)g + (uint32_t)(item_perm[items[t]] % g);
}

__global__ void u32_to_i64_kernel(const uint32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) out[t] = (int64_t)in[t];
}

static inline int grid_n(int64_t n) { return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n, 256), 1), 148 * 16); }

// sorted rating indices by `key` (stable) + bucket pointers; shared by CSR and block partition
static int32_t sort_indices_by_key(Ctx* ctx, uint32_t* d_key /* n, clobbered */, int64_t n, uint32_t n_buckets,
                                   int64_t* h_ptr /* n_buckets+1 */, int32_t* h_idx /* n */)
{
    cudaStream_t s = ctx->stream;
    // bucket pointers
    DevBuf<uint32_t> cnt, ptr;
    MML_TRY(cnt.alloc(n_buckets)); MML_TRY(ptr.alloc((size_t)n_buckets + 1));
    MML_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
    MML_TRY(histogram_i32((const int32_t*)d_key, n, cnt.p, s));
    MML_TRY(exclusive_scan_u32(cnt.p, ptr.p, n_buckets, s));
    if (h_ptr) {
        DevBuf<int64_t> ptr64;
        MML_TRY(ptr64.alloc((size_t)n_buckets + 1));
        u32_to_i64_kernel<<<grid_n(n_buckets + 1), 256, 0, s>>>(ptr.p, ptr64.p, (int64_t)n_buckets + 1);
        MML_CUDA(cudaGetLastError());
        MML_CUDA(cudaMemcpyAsync(h_ptr, ptr64.p, sizeof(int64_t) * ((size_t)n_buckets + 1), cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaStreamSynchronize(s));
    }
    if (h_idx && n > 0) {
        DevBuf<uint32_t> vals, ktmp, vtmp;
        MML_TRY(vals.alloc(n)); MML_TRY(ktmp.alloc(n)); MML_TRY(vtmp.alloc(n));
        MML_TRY(iota_u32(vals.p, n, s));
        MML_TRY(radix_sort_pairs(d_key, vals.p, ktmp.p, vtmp.p, n, bits_for(n_buckets > 0 ? n_buckets - 1 : 0), s));
        MML_CUDA(cudaMemcpyAsync(h_idx, vals.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaStreamSynchronize(s));
    }
    return MML_OK;
}

}  // namespace mml

using namespace mml;

// =================================== C ABI: context ==============================================
struct mml_ctx { Ctx c; };
struct mml_ratings { Ratings r; };

extern "C" const char* mml_last_error(void) { return mml::last_error(); }
extern "C" const char* mml_version(void) { return "mmlb200 0.1 sm_100a"; }

namespace mml {
int32_t ctx_create_on_device(int dev, mml_ctx** out)
{
    int count = 0;
    MML_CUDA(cudaGetDeviceCount(&count));
    MML_CHECK(count > 0, MML_ERR_CUDA, "mml_ctx_create: no CUDA device (this library has no CPU fallback)");
    MML_CHECK(dev >= 0 && dev < count, MML_ERR_ARG, "mml_ctx_create: device %d not in [0,%d)", dev, count);
    MML_CUDA(cudaSetDevice(dev));
    mml_ctx* c = new (std::nothrow) mml_ctx();
    MML_CHECK(c != nullptr, MML_ERR_ARG, "out of host memory");
    c->c.device = dev;
    cudaDeviceProp prop;
    MML_CUDA(cudaGetDeviceProperties(&prop, dev));
    c->c.sm_count = prop.multiProcessorCount;
    {   // the primitives' scratch is stream-ordered (StreamBuf): keep freed blocks in the pool instead of returning them
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    MML_CUDA(cudaStreamCreateWithFlags(&c->c.stream, cudaStreamNonBlocking));
    MML_CUDA(cudaStreamCreateWithFlags(&c->c.copy_stream, cudaStreamNonBlocking));
    MML_CUDA(cudaStreamCreateWithFlags(&c->c.out_stream, cudaStreamNonBlocking));
    MML_CUDA(cudaStreamCreateWithFlags(&c->c.aux_stream, cudaStreamNonBlocking));
    MML_CUDA(cudaEventCreateWithFlags(&c->c.copy_done, cudaEventDisableTiming));
    *out = c;
    return MML_OK;
}
}

extern "C" int32_t mml_ctx_create(int32_t n_gpus, const int32_t* device_ids, mml_ctx** out)
{
    MML_CHECK(out != nullptr, MML_ERR_ARG, "mml_ctx_create: out is NULL");
    MML_CHECK(n_gpus >= 1 && n_gpus <= 64, MML_ERR_ARG, "mml_ctx_create: n_gpus=%d", n_gpus);
    if (n_gpus == 1) return ctx_create_on_device(device_ids ? device_ids[0] : 0, out);
    // One process, n_gpus GPUs (the NumGpus property of the host classes): a root context over one rank context per GPU.
    int count = 0;
    MML_CUDA(cudaGetDeviceCount(&count));
    MML_CHECK(count >= n_gpus, MML_ERR_CUDA, "mml_ctx_create: %d GPUs asked for, %d visible (no CPU fallback)", n_gpus, count);
    mml_ctx* root = new (std::nothrow) mml_ctx();
    MML_CHECK(root != nullptr, MML_ERR_ARG, "out of host memory");
    root->c.device = device_ids ? device_ids[0] : 0;
    root->c.n_gpus = n_gpus;
    int32_t st = MML_OK;
    for (int r = 0; r < n_gpus && st == MML_OK; r++) {
        mml_ctx* p = nullptr;
        st = ctx_create_on_device(device_ids ? device_ids[r] : r, &p);
        if (st == MML_OK) root->c.peers.push_back(p);
    }
    if (st == MML_OK) {
        std::vector<Ctx*> pc;
        for (mml_ctx* p : root->c.peers) pc.push_back(&p->c);
        st = dist_init_all(pc);
        root->c.sm_count = pc[0]->sm_count;
    }
    if (st != MML_OK) { mml_ctx_destroy(root); return st; }
    *out = root;
    return MML_OK;
}

extern "C" int32_t mml_ctx_destroy(mml_ctx* ctx)
{
    if (!ctx) return MML_OK;
    if (ctx->c.is_root()) {
        for (mml_ctx* p : ctx->c.peers) mml_ctx_destroy(p);
        delete ctx;
        return MML_OK;
    }
    cudaSetDevice(ctx->c.device);
    dist_destroy(&ctx->c);
    topn_cache_destroy(&ctx->c);
    topn_exact_cache_destroy(&ctx->c);
    if (ctx->c.out_stream) cudaStreamDestroy(ctx->c.out_stream);
    if (ctx->c.aux_stream) cudaStreamDestroy(ctx->c.aux_stream);
    if (ctx->c.stream) cudaStreamDestroy(ctx->c.stream);
    if (ctx->c.copy_stream) cudaStreamDestroy(ctx->c.copy_stream);
    if (ctx->c.copy_done) cudaEventDestroy(ctx->c.copy_done);
    if (ctx->c.flush_buf) cudaFree(ctx->c.flush_buf);
    delete ctx;
    return MML_OK;
}

extern "C" int32_t mml_ctx_synchronize(mml_ctx* ctx)
{
    MML_LOCK(mml::ctx_of(ctx));
    MML_CHECK(ctx != nullptr, MML_ERR_ARG, "ctx is NULL");
    if (ctx->c.is_root()) {
        for (mml_ctx* p : ctx->c.peers) MML_TRY(mml_ctx_synchronize(p));
        return MML_OK;
    }
    MML_CUDA(cudaSetDevice(ctx->c.device));
    MML_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return MML_OK;
}

/* Writes a buffer larger than the 126 MB L2 so that the next timed kernel starts cold. */
extern "C" int32_t mml_ctx_flush_l2(mml_ctx* ctx)
{
    MML_LOCK(mml::ctx_of(ctx));
    MML_CHECK(ctx != nullptr, MML_ERR_ARG, "ctx is NULL");
    if (ctx->c.is_root()) {
        for (mml_ctx* p : ctx->c.peers) MML_TRY(mml_ctx_flush_l2(p));
        return MML_OK;
    }
    MML_CUDA(cudaSetDevice(ctx->c.device));
    const size_t bytes = (size_t)384 << 20;
    if (ctx->c.flush_buf == nullptr) MML_CUDA(cudaMalloc(&ctx->c.flush_buf, bytes));
    ctx->c.flush_val ^= 0xA5;
    MML_CUDA(cudaMemsetAsync(ctx->c.flush_buf, ctx->c.flush_val, bytes, ctx->c.stream));
    MML_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return MML_OK;
}

// ---- L2 probe (mml_ctx_probe_l2): the SGD epoch kernel's item-row traffic without the arithmetic ---------------------------
namespace mml {
template <int MODE>
__global__ void __launch_bounds__(512) l2_probe_kernel(float* __restrict__ tab, const uint32_t n_rows, const int row_f4, const int touches,
                                                       float* __restrict__ sink)
{
    const int lane = threadIdx.x & 7;                                   // lane inside the 8-lane worker
    const uint32_t worker = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    uint32_t x = worker * 2654435761u + 12345u;
    float acc = 0.f;
    const int pieces = row_f4 / 8;                                      // 128-bit pieces per lane
    for (int t = 0; t < touches; t += 2) {                              // two independent rows in flight per worker
        x = x * 1664525u + 1013904223u;
        const uint32_t r0 = (x >> 8) % n_rows;
        x = x * 1664525u + 1013904223u;
        const uint32_t r1 = (x >> 8) % n_rows;
        float4* p0 = reinterpret_cast<float4*>(tab) + (size_t)r0 * row_f4 + lane;
        float4* p1 = reinterpret_cast<float4*>(tab) + (size_t)r1 * row_f4 + lane;
        for (int v = 0; v < pieces; v++) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (MODE != MML_L2_RED) {
                asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p0 + 8 * v) : "memory");
                asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p1 + 8 * v) : "memory");
                acc += (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
            }
            if (MODE != MML_L2_READ) {
                const float d = MODE == MML_L2_READ_RED ? 1e-30f * a.x : 1e-30f;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(p0 + 8 * v), "f"(d) : "memory");
                asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" :: "l"(p1 + 8 * v), "f"(d) : "memory");
            }
        }
    }
    if (acc == 123.456f) sink[0] = acc;                                 // keeps the loads alive
}
}  // namespace mml

extern "C" int32_t mml_ctx_probe_l2(mml_ctx* ctx, int32_t mode, int32_t n_rows, int32_t row_floats, int32_t reps, double* rows_per_s)
{
    MML_LOCK(mml::ctx_of(ctx));
    MML_CHECK(ctx != nullptr && rows_per_s != nullptr, MML_ERR_ARG, "NULL argument");
    MML_CHECK(mode >= MML_L2_READ && mode <= MML_L2_READ_RED && n_rows > 0 && row_floats >= 32 && row_floats % 32 == 0 && reps > 0,
              MML_ERR_ARG, "mml_ctx_probe_l2: mode 0..2, row_floats a multiple of 32");
    if (ctx->c.is_root()) return mml_ctx_probe_l2(ctx->c.peers[0], mode, n_rows, row_floats, reps, rows_per_s);
    MML_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.stream;
    mml::DevBuf<float> tab, sink;
    MML_TRY(tab.alloc((size_t)n_rows * row_floats));
    MML_TRY(sink.alloc(1));
    MML_CUDA(cudaMemsetAsync(tab.p, 0, tab.bytes(), s));
    const int touches = 2048, grid = ctx->c.sm_count, threads = 512;
    cudaEvent_t e0, e1;
    MML_CUDA(cudaEventCreate(&e0)); MML_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps + 1; r++) {                                // launch 0 warms the L2
        cudaEventRecord(e0, s);
        if (mode == MML_L2_READ) mml::l2_probe_kernel<MML_L2_READ><<<grid, threads, 0, s>>>(tab.p, (uint32_t)n_rows, row_floats / 4, touches, sink.p);
        else if (mode == MML_L2_RED) mml::l2_probe_kernel<MML_L2_RED><<<grid, threads, 0, s>>>(tab.p, (uint32_t)n_rows, row_floats / 4, touches, sink.p);
        else mml::l2_probe_kernel<MML_L2_READ_RED><<<grid, threads, 0, s>>>(tab.p, (uint32_t)n_rows, row_floats / 4, touches, sink.p);
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    MML_CUDA(cudaGetLastError());
    *rows_per_s = (double)grid * (threads / 8) * touches / ((double)best * 1e-3);
    return MML_OK;
}

extern "C" int32_t mml_ctx_sm_count(mml_ctx* ctx, int32_t* out)
{
    MML_LOCK(mml::ctx_of(ctx));
    MML_CHECK(ctx != nullptr && out != nullptr, MML_ERR_ARG, "NULL argument");
    *out = ctx->c.sm_count;
    return MML_OK;
}

// =================================== C ABI: ratings ==============================================
extern "C" int32_t mml_ratings_create(mml_ctx* ctx, const int32_t* users, const int32_t* items, const float* values,
                                      int64_t n, int32_t max_user, int32_t max_item, mml_ratings** out)
{
    MML_LOCK(mml::ctx_of(ctx));
    MML_CHECK(ctx && out, MML_ERR_ARG, "mml_ratings_create: NULL argument");
    MML_CHECK(n >= 0 && n < ((int64_t)1 << 31), MML_ERR_ARG, "mml_ratings_create: n=%lld out of range", (long long)n);
    MML_CHECK(n == 0 || (users && items && values), MML_ERR_ARG, "mml_ratings_create: NULL data");
    MML_CHECK(max_user >= -1 && max_item >= -1, MML_ERR_ARG, "mml_ratings_create: bad max ids");
    if (ctx->c.is_root()) {
        // one shard per GPU: the users with u % N == r (the reference's block rule, MultiCore.cs:64, lifted to GPUs), with
        // the whole id space; Average and the scale are those of the whole set
        const int N = (int)ctx->c.peers.size();
        mml_ratings* root = new (std::nothrow) mml_ratings();
        MML_CHECK(root != nullptr, MML_ERR_ARG, "out of host memory");
        root->r.ctx = &ctx->c; root->r.n = n; root->r.max_user = max_user; root->r.max_item = max_item;
        root->r.shards.assign((size_t)N, nullptr);
        std::vector<double> sums((size_t)N, 0.0);
        const int32_t st = on_ranks(N, [&](int r) -> int32_t {
            std::vector<int32_t> su, si; std::vector<float> sv;
            su.reserve((size_t)(n / N + 1024)); si.reserve((size_t)(n / N + 1024)); sv.reserve((size_t)(n / N + 1024));
            for (int64_t t = 0; t < n; t++)
                if (users[t] >= 0 && users[t] % N == r) { su.push_back(users[t]); si.push_back(items[t]); sv.push_back(values[t]); }
                else if (users[t] < 0 && r == 0) { set_error("mml_ratings_create: rating %lld has a negative user id", (long long)t); return MML_ERR_ARG; }
            double sum = 0;
            for (float x : sv) sum += (double)x;
            sums[(size_t)r] = sum;
            return mml_ratings_create(ctx->c.peers[(size_t)r], su.data(), si.data(), sv.data(), (int64_t)su.size(), max_user, max_item,
                                      &root->r.shards[(size_t)r]);
        });
        if (st != MML_OK) { mml_ratings_destroy(root); return st; }
        double sum = 0; float mn = FLT_MAX, mx = -FLT_MAX;
        for (int r = 0; r < N; r++) {
            sum += sums[(size_t)r];
            if (root->r.shards[(size_t)r]->r.n > 0) {
                mn = std::min(mn, root->r.shards[(size_t)r]->r.min_rating); mx = std::max(mx, root->r.shards[(size_t)r]->r.max_rating);
            }
        }
        root->r.average = n > 0 ? (float)sum / (float)n : 0.f;
        root->r.min_rating = mn; root->r.max_rating = mx;
        *out = root;
        return MML_OK;
    }
    MML_CUDA(cudaSetDevice(ctx->c.device));
    mml_ratings* h = new (std::nothrow) mml_ratings();
    MML_CHECK(h != nullptr, MML_ERR_ARG, "out of host memory");
    Ratings& r = h->r;
    r.ctx = &ctx->c; r.n = n; r.max_user = max_user; r.max_item = max_item;
    cudaStream_t s = ctx->c.stream;
    int32_t st = MML_OK;
    do {
        if ((st = r.users.alloc(n)) || (st = r.items.alloc(n)) || (st = r.values.alloc(n))) break;
        if ((st = r.count_by_user.alloc(r.n_users())) || (st = r.count_by_item.alloc(r.n_items()))) break;
        // range check on the host while the copies are in flight would need a second pass; the ids
        // are validated on device by the histogram below only implicitly, so check here.
        for (int64_t t = 0; t < n; t++) {
            if ((uint32_t)users[t] > (uint32_t)max_user || (uint32_t)items[t] > (uint32_t)max_item) {
                set_error("mml_ratings_create: rating %lld has id out of range (user %d, item %d)", (long long)t, users[t], items[t]);
                st = MML_ERR_ARG; break;
            }
        }
        if (st) break;
        if (n > 0) {
            if (cudaMemcpyAsync(r.users.p, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s) != cudaSuccess ||
                cudaMemcpyAsync(r.items.p, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s) != cudaSuccess ||
                cudaMemcpyAsync(r.values.p, values, sizeof(float) * n, cudaMemcpyHostToDevice, s) != cudaSuccess) {
                set_error("mml_ratings_create: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
                st = MML_ERR_CUDA; break;
            }
        }
        cudaMemsetAsync(r.count_by_user.p, 0, r.count_by_user.bytes(), s);
        cudaMemsetAsync(r.count_by_item.p, 0, r.count_by_item.bytes(), s);
        if ((st = histogram_i32(r.users.p, n, r.count_by_user.p, s))) break;
        if ((st = histogram_i32(r.items.p, n, r.count_by_item.p, s))) break;
        if ((st = compute_stats(&r))) break;
    } while (0);
    if (st) { delete h; return st; }
    *out = h;
    return MML_OK;
}

extern "C" int32_t mml_ratings_destroy(mml_ratings* r)
{
    MML_LOCK((r ? mml::ratings_of(r)->ctx : nullptr));
    if (!r) return MML_OK;
    if (!r->r.shards.empty() || r->r.ctx->is_root()) {
        for (mml_ratings* s : r->r.shards) mml_ratings_destroy(s);
        delete r;
        return MML_OK;
    }
    cudaSetDevice(r->r.ctx->device);
    delete r;
    return MML_OK;
}

extern "C" int32_t mml_ratings_counts(mml_ratings* h, int32_t by_item, int32_t* counts_out)
{
    MML_LOCK((h ? mml::ratings_of(h)->ctx : nullptr));
    MML_CHECK(h && counts_out, MML_ERR_ARG, "mml_ratings_counts: NULL argument");
    Ratings& r = h->r;
    if (!r.shards.empty()) {   // every shard counts over the whole id space: add them up
        const int32_t rows = by_item ? r.n_items() : r.n_users();
        std::vector<int32_t> part((size_t)rows);
        for (int32_t t = 0; t < rows; t++) counts_out[t] = 0;
        for (mml_ratings* s : r.shards) {
            MML_TRY(mml_ratings_counts(s, by_item, part.data()));
            for (int32_t t = 0; t < rows; t++) counts_out[t] += part[(size_t)t];
        }
        return MML_OK;
    }
    MML_CUDA(cudaSetDevice(r.ctx->device));
    DevBuf<uint32_t>& c = by_item ? r.count_by_item : r.count_by_user;
    const int32_t rows = by_item ? r.n_items() : r.n_users();
    MML_CUDA(cudaMemcpyAsync(counts_out, c.p, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost, r.ctx->stream));
    MML_CUDA(cudaStreamSynchronize(r.ctx->stream));
    return MML_OK;
}

extern "C" int32_t mml_ratings_csr(mml_ratings* h, int32_t by_item, int64_t* row_ptr, int32_t* idx)
{
    MML_LOCK((h ? mml::ratings_of(h)->ctx : nullptr));
    MML_CHECK(h && row_ptr && idx, MML_ERR_ARG, "mml_ratings_csr: NULL argument");
    Ratings& r = h->r;
    MML_CHECK(r.shards.empty(), MML_ERR_UNSUPPORTED, "mml_ratings_csr: not available on a multi-GPU context (the shards index their own ratings)");
    MML_CUDA(cudaSetDevice(r.ctx->device));
    const int32_t rows = by_item ? r.n_items() : r.n_users();
    DevBuf<uint32_t> key;
    MML_TRY(key.alloc(r.n));
    MML_CUDA(cudaMemcpyAsync(key.p, by_item ? r.items.p : r.users.p, sizeof(uint32_t) * (size_t)r.n,
                             cudaMemcpyDeviceToDevice, r.ctx->stream));
    return sort_indices_by_key(r.ctx, key.p, r.n, (uint32_t)rows, row_ptr, idx);
}

extern "C" int32_t mml_ratings_stats(mml_ratings* h, float* average, float* min_rating, float* max_rating)
{
    MML_LOCK((h ? mml::ratings_of(h)->ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "mml_ratings_stats: NULL argument");
    if (average) *average = h->r.average;
    if (min_rating) *min_rating = h->r.min_rating;
    if (max_rating) *max_rating = h->r.max_rating;
    return MML_OK;
}

extern "C" int32_t mml_partition_blocks(mml_ratings* h, const int32_t* user_perm, const int32_t* item_perm, int32_t g,
                                        int64_t* block_ptr, int32_t* idx)
{
    MML_LOCK((h ? mml::ratings_of(h)->ctx : nullptr));
    MML_CHECK(h && user_perm && item_perm && block_ptr && idx, MML_ERR_ARG, "mml_partition_blocks: NULL argument");
    Ratings& r = h->r;
    MML_CHECK(r.shards.empty(), MML_ERR_UNSUPPORTED, "mml_partition_blocks: not available on a multi-GPU context");
    MML_CHECK(g >= 1 && (int64_t)g * g < ((int64_t)1 << 31), MML_ERR_ARG, "mml_partition_blocks: bad g=%d", g);
    MML_CUDA(cudaSetDevice(r.ctx->device));
    cudaStream_t s = r.ctx->stream;
    DevBuf<int32_t> up, ip; DevBuf<uint32_t> key;
    MML_TRY(up.alloc(r.n_users())); MML_TRY(ip.alloc(r.n_items())); MML_TRY(key.alloc(r.n));
    MML_CUDA(cudaMemcpyAsync(up.p, user_perm, sizeof(int32_t) * r.n_users(), cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(ip.p, item_perm, sizeof(int32_t) * r.n_items(), cudaMemcpyHostToDevice, s));
    block_key_kernel<<<grid_n(r.n), 256, 0, s>>>(r.users.p, r.items.p, r.n, up.p, ip.p, g, key.p);
    MML_CUDA(cudaGetLastError());
    return sort_indices_by_key(r.ctx, key.p, r.n, (uint32_t)(g * g), block_ptr, idx);
}

// MultiCore.PartitionIndices (MultiCore.cs:79-92): element t of RandomIndex goes to list t % g, at position t / g
namespace mml {
__global__ void partition_indices_kernel(const int32_t* __restrict__ ri, int64_t n, int32_t g, int32_t* __restrict__ idx)
{
    const int64_t base = n / g, rem = n % g;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const int64_t l = t % g, pos = t / g;
        idx[l * base + (l < rem ? l : rem) + pos] = ri[t];
    }
}
}

extern "C" int32_t mml_partition_indices(mml_ctx* hctx, const int32_t* random_index, int64_t n, int32_t num_groups,
                                         int64_t* list_ptr, int32_t* idx)
{
    MML_LOCK(mml::ctx_of(hctx));
    MML_CHECK(hctx && list_ptr && (n == 0 || (random_index && idx)), MML_ERR_ARG, "mml_partition_indices: NULL argument");
    MML_CHECK(num_groups >= 1 && n >= 0 && n < ((int64_t)1 << 31), MML_ERR_ARG, "mml_partition_indices: bad sizes");
    Ctx* ctx = ctx_of(hctx);
    if (ctx->is_root()) return mml_partition_indices(ctx->peers[0], random_index, n, num_groups, list_ptr, idx);
    const int32_t g = (int32_t)std::min<int64_t>(num_groups, n);          // :81
    for (int32_t l = 0; l <= num_groups; l++) {
        if (g == 0) { list_ptr[l] = 0; continue; }
        const int64_t ll = std::min<int64_t>(l, g), base = n / g, rem = n % g;
        list_ptr[l] = ll * base + std::min<int64_t>(ll, rem);
    }
    if (n == 0) return MML_OK;
    MML_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevBuf<int32_t> d_ri, d_idx;
    MML_TRY(d_ri.alloc(n)); MML_TRY(d_idx.alloc(n));
    MML_CUDA(cudaMemcpyAsync(d_ri.p, random_index, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    partition_indices_kernel<<<grid_n(n), 256, 0, s>>>(d_ri.p, n, g, d_idx.p);
    MML_CUDA(cudaGetLastError());
    MML_CUDA(cudaMemcpyAsync(idx, d_idx.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

namespace mml {
Ratings* ratings_of(mml_ratings* h) { return h ? &h->r : nullptr; }
Ctx* ctx_of(mml_ctx* h) { return h ? &h->c : nullptr; }
}
