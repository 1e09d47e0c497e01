// shuffle.cu -- applies a Fisher-Yates swap-target array bit-exactly, in parallel.
//
// Reference: Utils.Shuffle (Utils.cs:52-64) -- for i = n-1 .. 0: swap(a[i], a[H[i]]) with
// H[i] = random.Next(i + 1) -- as used for DataSet.RandomIndex (Data/DataSet.cs:193-202) and the
// permutations of MultiCore.PartitionUsersAndItems (MultiCore.cs:51-52, 68-70). The RNG stream is
// sequential and stays on the host; this kernel sequence reproduces the loop's result exactly.
//
// Method (deterministic reservations, Shun et al., SODA 2015): an iteration may run as soon as no
// EARLIER pending iteration (= larger i, the loop counts down) touches one of its two cells.
// Round: every pending iteration writes its index into both cells with atomicMax; those that read
// their own index back from both cells own them, swap, and retire; the rest retry. Swaps of one
// round touch disjoint cells; the earliest pending iteration always wins, so the loop terminates,
// in O(log n) rounds for random targets.
#include "common.cuh"
#include <algorithm>

namespace mml {

__global__ void shuf_reserve_kernel(const int32_t* __restrict__ pending, int64_t m,
                                    const int32_t* __restrict__ H, int32_t* __restrict__ R)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < m; t += stride) {
        const int32_t i = pending[t];
        atomicMax(&R[i], i);
        atomicMax(&R[H[i]], i);
    }
}

__global__ void shuf_commit_kernel(const int32_t* __restrict__ pending, int64_t m, const int32_t* __restrict__ H,
                                   const int32_t* __restrict__ R, int32_t* __restrict__ a, uint8_t* __restrict__ done)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < m; t += stride) {
        const int32_t i = pending[t], h = H[i];
        const bool own = R[i] == i && R[h] == i;
        if (own) { const int32_t x = a[i]; a[i] = a[h]; a[h] = x; }
        done[t] = own ? 1 : 0;
    }
}

__global__ void shuf_retire_kernel(const int32_t* __restrict__ pending, int64_t m, const int32_t* __restrict__ H,
                                   int32_t* __restrict__ R, const uint8_t* __restrict__ done,
                                   int32_t* __restrict__ next, unsigned long long* __restrict__ next_count)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < m; t += stride) {
        const int32_t i = pending[t];
        R[i] = -1; R[H[i]] = -1;
        if (!done[t]) next[atomicAdd(next_count, 1ull)] = i;
    }
}

__global__ void fill_i32_kernel(int32_t* __restrict__ p, int64_t n, int32_t v)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) p[t] = v;
}

int32_t shuffle_apply_device(int32_t* d_a, const int32_t* d_H, int64_t n, cudaStream_t s, int* rounds_out)
{
    if (n <= 1) { if (rounds_out) *rounds_out = 0; return MML_OK; }
    DevBuf<int32_t> R, pend0, pend1; DevBuf<uint8_t> done; DevBuf<unsigned long long> cnt;
    MML_TRY(R.alloc(n)); MML_TRY(pend0.alloc(n)); MML_TRY(pend1.alloc(n)); MML_TRY(done.alloc(n)); MML_TRY(cnt.alloc(1));
    const int T = 256;
    auto grid = [&](int64_t m) { return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(m, T), 1), 148 * 16); };
    fill_i32_kernel<<<grid(n), T, 0, s>>>(R.p, n, -1);
    MML_TRY(iota_u32((uint32_t*)pend0.p, n, s));
    int32_t* cur = pend0.p; int32_t* nxt = pend1.p;
    int64_t m = n;
    int rounds = 0;
    while (m > 0) {
        MML_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), s));
        shuf_reserve_kernel<<<grid(m), T, 0, s>>>(cur, m, d_H, R.p);
        shuf_commit_kernel<<<grid(m), T, 0, s>>>(cur, m, d_H, R.p, d_a, done.p);
        shuf_retire_kernel<<<grid(m), T, 0, s>>>(cur, m, d_H, R.p, done.p, nxt, cnt.p);
        MML_CUDA(cudaGetLastError());
        unsigned long long left = 0;
        MML_CUDA(cudaMemcpyAsync(&left, cnt.p, sizeof(left), cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaStreamSynchronize(s));
        MML_CHECK((int64_t)left < m, MML_ERR_STATE, "shuffle_apply made no progress");
        m = (int64_t)left;
        std::swap(cur, nxt);
        rounds++;
    }
    if (rounds_out) *rounds_out = rounds;
    return MML_OK;
}

}  // namespace mml

using namespace mml;

extern "C" int32_t mml_shuffle_apply(mml_ctx* hctx, int32_t* perm, const int32_t* H, int64_t n)
{
    MML_LOCK(mml::ctx_of(hctx));
    MML_CHECK(hctx && (n == 0 || (perm && H)), MML_ERR_ARG, "mml_shuffle_apply: NULL argument");
    MML_CHECK(n >= 0 && n < ((int64_t)1 << 31), MML_ERR_ARG, "mml_shuffle_apply: n out of range");
    Ctx* ctx = ctx_of(hctx);
    if (ctx->is_root()) return mml_shuffle_apply(ctx->peers[0], perm, H, n);   // one permutation: one GPU
    for (int64_t i = 0; i < n; i++)
        MML_CHECK(H[i] >= 0 && H[i] <= i, MML_ERR_ARG, "mml_shuffle_apply: H[%lld]=%d is not in [0,%lld]", (long long)i, H[i], (long long)i);
    if (n == 0) return MML_OK;
    MML_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevBuf<int32_t> a, h;
    MML_TRY(a.alloc(n)); MML_TRY(h.alloc(n));
    MML_CUDA(cudaMemcpyAsync(a.p, perm, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(h.p, H, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    MML_TRY(shuffle_apply_device(a.p, h.p, n, s, nullptr));
    MML_CUDA(cudaMemcpyAsync(perm, a.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}
