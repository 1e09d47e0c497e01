// common.cuh -- shared declarations of libmmlb200 (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <functional>
#include <algorithm>

#include "../../include/mmlb200.h"

namespace mml {

// ---- error handling: every ABI entry point returns a status, never throws -----------------
void set_error(const char* fmt, ...);
const char* last_error();

#define MML_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mml::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return MML_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define MML_CHECK(cond, code, ...)                                                             \
    do {                                                                                       \
        if (!(cond)) { mml::set_error(__VA_ARGS__); return (code); }                          \
    } while (0)

#define MML_TRY(expr)                                                                          \
    do { int32_t _s = (expr); if (_s != MML_OK) return _s; } while (0)

// ---- a device buffer that frees itself --------------------------------------------------------
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    int32_t alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        MML_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
        return MML_OK;
    }
    // grow-only: keeps the buffer when it already holds `count` elements (no cudaFree / cudaMalloc on the hot path)
    int32_t ensure(size_t count) { return (p && n >= std::max<size_t>(count, 1)) ? MML_OK : alloc(count); }
    size_t bytes() const { return n * sizeof(T); }
};

// ---- a stream-ordered temporary (cudaMallocAsync / cudaFreeAsync on `s`): scratch of the primitives, freed in stream order,
//      so neither the allocation nor the release makes the host wait for the device ---------------
template <typename T>
struct StreamBuf {
    T* p = nullptr;
    cudaStream_t s = nullptr;
    StreamBuf() = default;
    StreamBuf(const StreamBuf&) = delete;
    StreamBuf& operator=(const StreamBuf&) = delete;
    ~StreamBuf() { if (p) cudaFreeAsync(p, s); }
    int32_t alloc(size_t count, cudaStream_t stream) {
        s = stream;
        if (count == 0) count = 1;
        MML_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), s));
        return MML_OK;
    }
};

// ---- a page-locked host buffer (staging for asynchronous copies), grow-only ---------------------
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t n = 0;
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    int32_t ensure(size_t count) {
        if (count == 0) count = 1;
        if (p && n >= count) return MML_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        MML_CUDA(cudaHostAlloc((void**)&p, count * sizeof(T), cudaHostAllocDefault));
        n = count;
        return MML_OK;
    }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- primitives (prims.cu) ---------------------------------------------------------------------
// counts[ids[t]]++ for t < n (counts must be zeroed by the caller)
int32_t histogram_i32(const int32_t* ids, int64_t n, uint32_t* counts, cudaStream_t s);
// out[i] = sum_{j<i} in[j]; out has n+1 entries (out[n] = total). in/out may not alias.
int32_t exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s);
// Stable LSD radix sort of (key,value) pairs on key bits [0, key_bits). Result ends in keys/vals
// (ping-pong through keys_tmp/vals_tmp is handled inside; all four buffers hold n entries).
int32_t radix_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp, uint32_t* vals_tmp,
                         int64_t n, int key_bits, cudaStream_t s);
// vals[i] = i
int32_t iota_u32(uint32_t* vals, int64_t n, cudaStream_t s);
// out[i] = src[idx[i]]
int32_t gather_u32(const uint32_t* src, const uint32_t* idx, uint32_t* out, int64_t n, cudaStream_t s);
int bits_for(uint32_t max_value);

// ---- objects -----------------------------------------------------------------------------------
struct Ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host -> device staging that may overlap kernels on `stream`
    cudaEvent_t copy_done = nullptr;
    int n_gpus = 1;              // world size: one process (context) per GPU
    int rank = 0;
    void* comm = nullptr;        // ncclComm_t when n_gpus > 1
    void* flush_buf = nullptr;   // mml_ctx_flush_l2
    int flush_val = 0;
    void* topn_cache = nullptr;  // Recommend() workspace of this context, grow-only (topn_tc.cu); freed by topn_cache_destroy
    void* topn_exact_cache = nullptr;     // the same for the exact CUDA-core path (topn.cu)
    cudaStream_t out_stream = nullptr;    // device -> host result staging that may overlap the next batch's kernels
    cudaStream_t aux_stream = nullptr;    // a second kernel stream (WRMF: the next batch's Gram sums under this batch's solves)
    // One-process multi-GPU (mml_ctx_create with n_gpus > 1, the NumGpus property of the host classes): this context is
    // then only the root of `peers`, one ordinary rank context per GPU (rank r of n_gpus, NCCL communicators from
    // ncclCommInitAll); every entry point on the root or on a handle created from it fans out to the peers, one host
    // thread per GPU (on_ranks), exactly the calls a one-process-per-GPU host would make.
    std::vector<mml_ctx*> peers;
    bool is_root() const { return !peers.empty(); }
    // Every ABI call on this context or on a handle created from it holds this lock for its duration: the handles share the
    // context's streams and their own grow-only scratch buffers, and the reference calls Predict / Recommend concurrently from
    // TPL threads on one object (Eval/Items.cs:147-164). Recursive: some entry points are built on others.
    std::recursive_mutex mu;
};

#define MML_LOCK(ctxptr)                                                                       \
    std::unique_lock<std::recursive_mutex> _mml_lock;                                          \
    do { mml::Ctx* _lc = (ctxptr); if (_lc) _mml_lock = std::unique_lock<std::recursive_mutex>(_lc->mu); } while (0)

struct Ratings {
    Ctx* ctx = nullptr;
    int64_t n = 0;
    int32_t max_user = -1, max_item = -1;
    DevBuf<int32_t> users, items;
    DevBuf<float> values;
    DevBuf<uint32_t> count_by_user, count_by_item;
    float average = 0.f, min_rating = 0.f, max_rating = 0.f;
    std::vector<mml_ratings*> shards;   // root of a one-process multi-GPU rating set: shard r holds the users with u % N == r
    int32_t n_users() const { return max_user + 1; }
    int32_t n_items() const { return max_item + 1; }
};

Ratings* ratings_of(mml_ratings* h);
int32_t dist_destroy(Ctx* c);
void topn_cache_destroy(Ctx* c);
void topn_exact_cache_destroy(Ctx* c);
int32_t dist_allreduce_u32(Ctx* c, uint32_t* d_buf, size_t n);
int32_t dist_allreduce_f64(Ctx* c, double* d_buf, size_t n);
int32_t dist_allreduce_f64_max(Ctx* c, double* d_buf, size_t n);
int32_t dist_ring_exchange(Ctx* c, cudaStream_t stream, const float* send_a, size_t n_send_a, const float* send_b, size_t n_send_b, int to,
                           float* recv_a, size_t n_recv_a, float* recv_b, size_t n_recv_b, int from);
int32_t dist_broadcast_f32(Ctx* c, float* d_buf, size_t n, int root);
int32_t dist_group_start();
int32_t dist_group_end();
Ctx* ctx_of(mml_ctx* h);

// Runs fn(r) for r = 0..n-1, one host thread per rank (each thread drives one GPU: collectives inside fn meet their
// peers), joins, and returns the first non-zero status with that thread's error message made the caller's.
int32_t on_ranks(int n, const std::function<int32_t(int)>& fn);
// ncclCommInitAll for the peers of a root context (dist.cu)
int32_t dist_init_all(std::vector<Ctx*>& peers);

// Fan-out of an entry point on a root handle: `call` is evaluated once per shard with `s` bound to shard r's handle.
#define MML_FORWARD_ALL(handle, call)                                                          \
    do {                                                                                       \
        if ((handle) && !(handle)->shards.empty()) {                                           \
            auto& _sh = (handle)->shards;                                                      \
            return mml::on_ranks((int)_sh.size(), [&](int _r) -> int32_t { auto* s = _sh[_r]; (void)s; return (call); }); \
        }                                                                                      \
    } while (0)

}  // namespace mml
