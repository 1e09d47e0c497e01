// topn.cu -- Recommend() for a batch of users on an item-MF model (WRMF / MF).
//
// Reference: Recommender.Recommend (Recommender.cs:52-103): score every candidate that is not in ignore_items with
// Predict, keep those with score > float.MinValue; n > 0 -> the n best (C5 IntervalHeap), n = -1 -> all, by
// descending score (stable sort: ties keep candidate-list order). Predict of item-MF = RowScalarProduct
// (ItemRecommendation/MF.cs:151-157 -> DataType/MatrixExtensions.cs:224-241: sequential fp32 multiply then add,
// float.MinValue for ids outside the model).
//
// Result order here: (score desc, candidate position asc) -- the reference's order for n = -1 always, and for
// n > 0 whenever the best n + 1 scores differ (C5's tie order is not pinned by anything in the reference).
// Scores are bit-exact: every (user, candidate) dot product is accumulated f = 0 .. k-1 in fp32 with separate
// multiply and add, exactly as RowScalarProduct does.
//
// Round-1 status: exact scoring on the CUDA cores (tiled, shared-memory staged) into a score buffer, then a
// per-user selection kernel. The tcgen05 scoring GEMM (approximate scores to pick a candidate superset, exact
// re-scoring of the finalists) named by the north star is not written yet ("what comes next").
#include "common.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <new>

namespace mml {

static inline int grid_n(int64_t n, int threads = 256)
{
    return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n, threads), 1), 148 * 16);
}

// ---- exact scores of a user batch against all candidates -------------------------------------------------
constexpr int TS = 64;      // users per tile = candidates per tile
constexpr int TK = 32;      // factors per staging step

// scores[b * n_cand + c] = dot(U[users[b]], V[cand[c]]) (sequential fp32 mul/add), or -FLT_MAX for ids outside the model
__global__ void __launch_bounds__(256) score_tile_kernel(const float* __restrict__ U, int32_t n_model_users,
                                                         const float* __restrict__ V, int32_t n_model_items, int32_t k,
                                                         const int32_t* __restrict__ users, int32_t n_batch,
                                                         const int32_t* __restrict__ cand, int64_t n_cand,
                                                         float* __restrict__ scores)
{
    __shared__ float Us[TK][TS + 1];    // transposed: [f][user]
    __shared__ float Vs[TK][TS + 1];
    __shared__ int32_t su[TS], sc[TS];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;     // 16 x 16 threads, 4 x 4 results each
    const int64_t c0 = (int64_t)blockIdx.x * TS;
    const int b0 = blockIdx.y * TS;
    if (threadIdx.x < TS) {
        const int b = b0 + threadIdx.x;
        int32_t uid = b < n_batch ? users[b] : -1;
        if (uid >= n_model_users) uid = -1;
        su[threadIdx.x] = uid;
    } else if (threadIdx.x < 2 * TS) {
        const int64_t c = c0 + (threadIdx.x - TS);
        int32_t iid = c < n_cand ? (cand ? cand[c] : (int32_t)c) : -1;
        if (iid >= n_model_items) iid = -1;
        sc[threadIdx.x - TS] = iid;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
    for (int f0 = 0; f0 < k; f0 += TK) {
        // stage TS rows x TK factors of both sides (coalesced along f)
        for (int t = threadIdx.x; t < TS * TK; t += 256) {
            const int row = t / TK, f = t % TK;
            const int32_t uid = su[row], iid = sc[row];
            Us[f][row] = (uid >= 0 && f0 + f < k) ? U[(size_t)uid * k + f0 + f] : 0.f;
            Vs[f][row] = (iid >= 0 && f0 + f < k) ? V[(size_t)iid * k + f0 + f] : 0.f;
        }
        __syncthreads();
        const int fmax = min(TK, k - f0);
        for (int f = 0; f < fmax; f++) {
            float a4[4], b4[4];
#pragma unroll
            for (int a = 0; a < 4; a++) a4[a] = Us[f][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; b++) b4[b] = Vs[f][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = __fadd_rn(acc[a][b], __fmul_rn(a4[a], b4[b]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int b = b0 + ty * 4 + a;
        if (b >= n_batch) continue;
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
            const int64_t c = c0 + tx * 4 + bb;
            if (c >= n_cand) continue;
            const bool ok = su[ty * 4 + a] >= 0 && sc[tx * 4 + bb] >= 0;
            scores[(size_t)b * n_cand + c] = ok ? acc[a][bb] : -FLT_MAX;
        }
    }
}

// scores of the user's ignore_items -> -inf. pos_of[item] = first candidate position of item (or -1),
// next_same[pos] = next position holding the same item (or -1).
__global__ void mask_ignored_kernel(const int64_t* __restrict__ ign_ptr, const int32_t* __restrict__ ign_idx,
                                    int32_t b_lo, int32_t n_batch, const int32_t* __restrict__ pos_of, int32_t n_pos_of,
                                    const int32_t* __restrict__ next_same, int64_t n_cand, float* __restrict__ scores)
{
    const int b = blockIdx.x;
    if (b >= n_batch) return;
    const int64_t lo = ign_ptr[b_lo + b], hi = ign_ptr[b_lo + b + 1];
    for (int64_t t = lo + threadIdx.x; t < hi; t += blockDim.x) {
        const int32_t item = ign_idx[t];
        if (item < 0 || item >= n_pos_of) continue;
        for (int32_t pos = pos_of[item]; pos >= 0; pos = next_same[pos]) scores[(size_t)b * n_cand + pos] = -INFINITY;
    }
}

// ---- per-user selection of the n best (n <= MAXN) --------------------------------------------------------
constexpr int SEL_T = 128;
constexpr int MAXN = 32;

__device__ __forceinline__ bool better(float sa, int pa, float sb, int pb) { return sa > sb || (sa == sb && pa < pb); }

// one CTA per user: every thread keeps the n best of its strided share (sorted, in shared memory), then n rounds of
// block-wide arg-best over the list heads.
__global__ void __launch_bounds__(SEL_T) select_topn_kernel(const float* __restrict__ scores, int64_t n_cand, int32_t n,
                                                            const int32_t* __restrict__ cand,
                                                            int32_t* __restrict__ out_items, float* __restrict__ out_scores,
                                                            int32_t* __restrict__ out_counts)
{
    extern __shared__ unsigned char sel_smem[];
    float* ls = reinterpret_cast<float*>(sel_smem);                  // [SEL_T][n]
    int32_t* lp = reinterpret_cast<int32_t*>(ls + (size_t)SEL_T * n);  // [SEL_T][n]
    __shared__ float rs[SEL_T / 32]; __shared__ int rp[SEL_T / 32]; __shared__ int rt[SEL_T / 32];
    __shared__ int win_t;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* sc = scores + (size_t)b * n_cand;
    float* my_s = ls + (size_t)tid * n; int32_t* my_p = lp + (size_t)tid * n;
    int cnt = 0;
    for (int64_t c = tid; c < n_cand; c += SEL_T) {
        const float s = sc[c];
        if (!(s > -FLT_MAX)) continue;                               // float.MinValue and masked entries never qualify
        if (cnt == n && !better(s, (int)c, my_s[n - 1], my_p[n - 1])) continue;
        int pos = cnt < n ? cnt : n - 1;                             // insertion into the sorted list
        while (pos > 0 && better(s, (int)c, my_s[pos - 1], my_p[pos - 1])) { my_s[pos] = my_s[pos - 1]; my_p[pos] = my_p[pos - 1]; pos--; }
        my_s[pos] = s; my_p[pos] = (int32_t)c;
        if (cnt < n) cnt++;
    }
    int head = 0, produced = 0;
    for (int r = 0; r < n; r++) {
        float s = head < cnt ? my_s[head] : -INFINITY; int p = head < cnt ? my_p[head] : 0x7fffffff; int t = tid;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float s2 = __shfl_xor_sync(0xffffffffu, s, d); const int p2 = __shfl_xor_sync(0xffffffffu, p, d);
            const int t2 = __shfl_xor_sync(0xffffffffu, t, d);
            if (better(s2, p2, s, p)) { s = s2; p = p2; t = t2; }
        }
        if ((tid & 31) == 0) { rs[tid >> 5] = s; rp[tid >> 5] = p; rt[tid >> 5] = t; }
        __syncthreads();
        if (tid == 0) {
            float bs = rs[0]; int bp = rp[0], bt = rt[0];
            for (int w = 1; w < SEL_T / 32; w++) if (better(rs[w], rp[w], bs, bp)) { bs = rs[w]; bp = rp[w]; bt = rt[w]; }
            if (bp != 0x7fffffff) {
                out_items[(size_t)b * n + r] = cand ? cand[bp] : bp;
                out_scores[(size_t)b * n + r] = bs;
                win_t = bt;
            } else {
                win_t = -1;
            }
        }
        __syncthreads();
        if (win_t < 0) break;
        if (win_t == tid) head++;
        produced++;
        __syncthreads();
    }
    if (tid == 0) out_counts[b] = produced;
}

// ---- full ranking / large n: sort the whole score buffer ----------------------------------------------------
// key ascending == score descending (total order on floats; -0 and +0 compare equal in the reference, so map -0 to +0)
__global__ void rank_key_kernel(const float* __restrict__ scores, int64_t total, uint32_t* __restrict__ key, uint32_t* __restrict__ val)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        float s = scores[t];
        if (s == 0.f) s = 0.f;
        uint32_t u = __float_as_uint(s);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // ascending in s
        key[t] = ~u;                                          // descending in s
        val[t] = (uint32_t)t;
    }
}

__global__ void rank_user_key_kernel(const uint32_t* __restrict__ val, int64_t total, int64_t n_cand, uint32_t* __restrict__ key)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) key[t] = (uint32_t)(val[t] / n_cand);
}

// after sorting by (user, score desc, position asc): row b occupies [b * n_cand, (b + 1) * n_cand)
__global__ void rank_emit_kernel(const uint32_t* __restrict__ val, const float* __restrict__ scores, int64_t n_cand, int32_t n_out,
                                 const int32_t* __restrict__ cand, int32_t* __restrict__ out_items, float* __restrict__ out_scores,
                                 int32_t* __restrict__ out_counts)
{
    const int b = blockIdx.x;
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int local = 0;
    for (int64_t r = threadIdx.x; r < n_out; r += blockDim.x) {
        const uint32_t src = val[(size_t)b * n_cand + r];
        const float s = scores[src];
        if (s > -FLT_MAX) {
            const int64_t pos = (int64_t)src - (int64_t)b * n_cand;
            out_items[(size_t)b * n_out + r] = cand ? cand[pos] : (int32_t)pos;
            out_scores[(size_t)b * n_out + r] = s;
            local++;
        }
    }
    atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) out_counts[b] = s_cnt;     // qualifying entries sort first, so they are rows 0 .. count-1
}

// Buffers of the exact path, kept by the context between calls (grow-only).
struct ExactWork {
    DevBuf<int32_t> d_cand, d_pos_of, d_next, d_users, d_ign_idx, d_out_i, d_out_c;
    DevBuf<int64_t> d_ign_ptr;
    DevBuf<float> d_scores, d_out_s;
};
void topn_exact_cache_destroy(Ctx* ctx)
{
    delete reinterpret_cast<ExactWork*>(ctx->topn_exact_cache);
    ctx->topn_exact_cache = nullptr;
}


// d_U / d_V: model on the device. All other pointers: host. Exact scoring on the CUDA cores.
static int32_t topn_exact_device(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                    const int32_t* users, int64_t n_users, int32_t n,
                    const int32_t* candidates, int64_t n_cand,
                    const int64_t* ignore_ptr, const int32_t* ignore_idx,
                    int32_t* out_items, float* out_scores, int32_t* out_counts, int64_t* launches)
{
    cudaStream_t s = ctx->stream;
    if (n_users == 0) return MML_OK;
    if (!candidates) n_cand = n_model_items;
    const int32_t n_out = n < 0 ? (int32_t)n_cand : (int32_t)std::min<int64_t>(n, n_cand);
    if (n_cand == 0 || n_out == 0) { for (int64_t b = 0; b < n_users; b++) out_counts[b] = 0; return MML_OK; }
    MML_CHECK(n_cand < ((int64_t)1 << 31), MML_ERR_ARG, "topn: too many candidates");
    // buffers: grow-only members of the context (the tensor path sends its few undecided users here on every call: a dozen
    // cudaMalloc / cudaFree pairs per call cost 30-150 ms of wall clock on a busy box against 104 ms of device time)
    if (!ctx->topn_exact_cache) ctx->topn_exact_cache = new (std::nothrow) ExactWork();
    MML_CHECK(ctx->topn_exact_cache != nullptr, MML_ERR_ARG, "out of host memory");
    ExactWork& xw = *reinterpret_cast<ExactWork*>(ctx->topn_exact_cache);
    DevBuf<int32_t>& d_cand = xw.d_cand; DevBuf<int32_t>& d_pos_of = xw.d_pos_of; DevBuf<int32_t>& d_next = xw.d_next;
    DevBuf<int32_t>& d_users = xw.d_users; DevBuf<int32_t>& d_ign_idx = xw.d_ign_idx; DevBuf<int64_t>& d_ign_ptr = xw.d_ign_ptr;
    // candidate list and the position maps for ignore_items
    if (candidates) {
        MML_TRY(d_cand.ensure(n_cand));
        MML_CUDA(cudaMemcpyAsync(d_cand.p, candidates, sizeof(int32_t) * n_cand, cudaMemcpyHostToDevice, s));
    }
    int32_t n_pos_of = 0;
    if (ignore_ptr && ignore_idx && ignore_ptr[n_users] > 0) {
        int32_t max_id = n_model_items - 1;
        if (candidates) for (int64_t c = 0; c < n_cand; c++) max_id = std::max(max_id, candidates[c]);
        n_pos_of = max_id + 1;
        std::vector<int32_t> pos_of(std::max(n_pos_of, 1), -1), next(n_cand, -1);
        for (int64_t c = n_cand - 1; c >= 0; c--) {
            const int32_t item = candidates ? candidates[c] : (int32_t)c;
            if (item < 0) continue;
            next[c] = pos_of[item]; pos_of[item] = (int32_t)c;
        }
        MML_TRY(d_pos_of.ensure(pos_of.size())); MML_TRY(d_next.ensure(n_cand));
        MML_TRY(d_ign_ptr.ensure((size_t)n_users + 1)); MML_TRY(d_ign_idx.ensure((size_t)ignore_ptr[n_users]));
        MML_CUDA(cudaMemcpyAsync(d_pos_of.p, pos_of.data(), sizeof(int32_t) * pos_of.size(), cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaMemcpyAsync(d_next.p, next.data(), sizeof(int32_t) * n_cand, cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaMemcpyAsync(d_ign_ptr.p, ignore_ptr, sizeof(int64_t) * ((size_t)n_users + 1), cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaMemcpyAsync(d_ign_idx.p, ignore_idx, sizeof(int32_t) * (size_t)ignore_ptr[n_users], cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaStreamSynchronize(s));
    }
    MML_TRY(d_users.ensure(n_users));
    MML_CUDA(cudaMemcpyAsync(d_users.p, users, sizeof(int32_t) * n_users, cudaMemcpyHostToDevice, s));

    const bool fast = n > 0 && n_out <= MAXN;
    // users per pass: score buffer <= 1 GiB; the sorting path also needs 32-bit positions
    int64_t B = std::max<int64_t>(1, ((int64_t)1 << 28) / n_cand);
    if (!fast) B = std::max<int64_t>(1, std::min<int64_t>(B, (((int64_t)1 << 31) - 1) / n_cand));
    B = std::min<int64_t>(B, n_users);
    B = std::min<int64_t>(B, 65535 * (int64_t)TS);
    DevBuf<float>& d_scores = xw.d_scores; DevBuf<float>& d_out_s = xw.d_out_s; DevBuf<int32_t>& d_out_i = xw.d_out_i; DevBuf<int32_t>& d_out_c = xw.d_out_c;
    MML_TRY(d_scores.ensure((size_t)B * n_cand));
    MML_TRY(d_out_s.ensure((size_t)B * n_out)); MML_TRY(d_out_i.ensure((size_t)B * n_out)); MML_TRY(d_out_c.ensure(B));
    DevBuf<uint32_t> key, val, k2, v2;          // the n = -1 (full ranking) path sorts B x n_cand keys: per call
    if (!fast) { MML_TRY(key.alloc((size_t)B * n_cand)); MML_TRY(val.alloc((size_t)B * n_cand)); MML_TRY(k2.alloc((size_t)B * n_cand)); MML_TRY(v2.alloc((size_t)B * n_cand)); }
    for (int64_t b_lo = 0; b_lo < n_users; b_lo += B) {
        const int32_t nb = (int32_t)std::min<int64_t>(B, n_users - b_lo);
        dim3 grid((unsigned)ceil_div(n_cand, TS), (unsigned)ceil_div(nb, TS));
        score_tile_kernel<<<grid, 256, 0, s>>>(d_U, n_model_users, d_V, n_model_items, k, d_users.p + b_lo, nb,
                                               candidates ? d_cand.p : nullptr, n_cand, d_scores.p);
        if (n_pos_of > 0)
            mask_ignored_kernel<<<nb, 128, 0, s>>>(d_ign_ptr.p, d_ign_idx.p, (int32_t)b_lo, nb, d_pos_of.p, n_pos_of, d_next.p, n_cand, d_scores.p);
        MML_CUDA(cudaGetLastError());
        MML_CUDA(cudaMemsetAsync(d_out_i.p, 0, sizeof(int32_t) * (size_t)nb * n_out, s));
        MML_CUDA(cudaMemsetAsync(d_out_s.p, 0, sizeof(float) * (size_t)nb * n_out, s));
        if (fast) {
            const size_t smem = (size_t)SEL_T * n_out * 8;
            select_topn_kernel<<<nb, SEL_T, smem, s>>>(d_scores.p, n_cand, n_out, candidates ? d_cand.p : nullptr,
                                                       d_out_i.p, d_out_s.p, d_out_c.p);
            MML_CUDA(cudaGetLastError());
            if (launches) *launches += 3;
        } else {
            const int64_t total = (int64_t)nb * n_cand;
            rank_key_kernel<<<grid_n(total), 256, 0, s>>>(d_scores.p, total, key.p, val.p);
            MML_CUDA(cudaGetLastError());
            MML_TRY(radix_sort_pairs(key.p, val.p, k2.p, v2.p, total, 32, s));          // by score desc, stable in position
            rank_user_key_kernel<<<grid_n(total), 256, 0, s>>>(val.p, total, n_cand, key.p);
            MML_CUDA(cudaGetLastError());
            MML_TRY(radix_sort_pairs(key.p, val.p, k2.p, v2.p, total, bits_for((uint32_t)std::max(nb - 1, 1)), s));   // by user, stable
            rank_emit_kernel<<<nb, 256, 0, s>>>(val.p, d_scores.p, n_cand, n_out, candidates ? d_cand.p : nullptr,
                                                d_out_i.p, d_out_s.p, d_out_c.p);
            MML_CUDA(cudaGetLastError());
            if (launches) *launches += 16;
        }
        MML_CUDA(cudaMemcpyAsync(out_items + (size_t)b_lo * n_out, d_out_i.p, sizeof(int32_t) * (size_t)nb * n_out, cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaMemcpyAsync(out_scores + (size_t)b_lo * n_out, d_out_s.p, sizeof(float) * (size_t)nb * n_out, cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaMemcpyAsync(out_counts + b_lo, d_out_c.p, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaStreamSynchronize(s));
    }
    return MML_OK;
}

bool topn_tc_eligible(int32_t k, int32_t n, int64_t n_cand);
void topn_tc_set_filter(int kind);
int32_t topn_tc_run(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                    const int32_t* users, int64_t n_users, int32_t n, int32_t n_out, const int32_t* d_cand, int32_t n_cand,
                    bool has_invalid_cand, const int64_t* ignore_ptr, const int32_t* ignore_idx,
                    int32_t* out_items, float* out_scores, int32_t* out_counts, std::vector<int64_t>& redo_users, int64_t* launches);

static int g_topn_mode = 0;                         // MML_TOPN_AUTO
static thread_local int64_t g_stat_tc = 0, g_stat_exact = 0;     // users served by each path in this thread's last call
static thread_local float g_stat_ms = 0.f;

// Recommend() for a user batch. Tensor-core path (topn_tc.cu) when the request qualifies; users whose candidate
// superset it cannot prove complete, and requests outside its envelope, run on the exact CUDA-core path.
int32_t topn_device(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                    const int32_t* users, int64_t n_users, int32_t n,
                    const int32_t* candidates, int64_t n_cand,
                    const int64_t* ignore_ptr, const int32_t* ignore_idx,
                    int32_t* out_items, float* out_scores, int32_t* out_counts, int64_t* launches)
{
    g_stat_tc = 0; g_stat_exact = 0; g_stat_ms = 0.f;
    if (n_users == 0) return MML_OK;
    if (!candidates) n_cand = n_model_items;
    bool tc = g_topn_mode != MML_TOPN_EXACT && topn_tc_eligible(k, n, n_cand) && n_cand < ((int64_t)1 << 31);
    bool has_invalid = false;
    if (tc && candidates) {   // a candidate listed twice is scored (and may be returned) twice: exact path only
        std::vector<uint8_t> seen((size_t)std::max(n_model_items, 1), 0);
        for (int64_t c = 0; c < n_cand && tc; c++) {
            const int32_t id = candidates[c];
            if (id < 0 || id >= n_model_items) { has_invalid = true; continue; }
            if (seen[id]) tc = false;
            seen[id] = 1;
        }
    }
    MML_CHECK(tc || g_topn_mode != MML_TOPN_TENSOR, MML_ERR_UNSUPPORTED,
              "topn: request is outside the tensor-core path (n <= 16, num_factors <= 128, distinct candidates)");
    if (!tc) {
        g_stat_exact = n_users;
        return topn_exact_device(ctx, d_U, n_model_users, d_V, n_model_items, k, users, n_users, n, candidates, n_cand,
                                 ignore_ptr, ignore_idx, out_items, out_scores, out_counts, launches);
    }
    cudaStream_t s = ctx->stream;
    const int32_t n_out = (int32_t)std::min<int64_t>(n, n_cand);
    const int64_t n_ign = (ignore_ptr && ignore_idx) ? ignore_ptr[n_users] : 0;
    DevBuf<int32_t> d_cand;
    if (candidates) {
        MML_TRY(d_cand.alloc(n_cand));
        MML_CUDA(cudaMemcpyAsync(d_cand.p, candidates, sizeof(int32_t) * n_cand, cudaMemcpyHostToDevice, s));
    }
    cudaEvent_t e0, e1;
    MML_CUDA(cudaEventCreate(&e0)); MML_CUDA(cudaEventCreate(&e1));
    MML_CUDA(cudaEventRecord(e0, s));
    std::vector<int64_t> redo_users;
    MML_TRY(topn_tc_run(ctx, d_U, n_model_users, d_V, n_model_items, k, users, n_users, n, n_out, candidates ? d_cand.p : nullptr,
                        (int32_t)n_cand, has_invalid, ignore_ptr, ignore_idx, out_items, out_scores, out_counts, redo_users, launches));
    MML_CUDA(cudaEventRecord(e1, s));
    MML_CUDA(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&g_stat_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    g_stat_tc = n_users - (int64_t)redo_users.size();
    g_stat_exact = (int64_t)redo_users.size();
    if (redo_users.empty()) return MML_OK;
    // users the filter could not decide: exact path on the sub-batch, results scattered back
    const size_t nr = redo_users.size();
    std::vector<int32_t> ru(nr), ri(nr * (size_t)n_out), rc(nr), rign;
    std::vector<float> rs(nr * (size_t)n_out);
    std::vector<int64_t> rptr(nr + 1, 0);
    for (size_t t = 0; t < nr; t++) {
        ru[t] = users[redo_users[t]];
        if (n_ign > 0) {
            const int64_t lo = ignore_ptr[redo_users[t]], hi = ignore_ptr[redo_users[t] + 1];
            rign.insert(rign.end(), ignore_idx + lo, ignore_idx + hi);
            rptr[t + 1] = rptr[t] + (hi - lo);
        }
    }
    if (rign.empty()) rign.push_back(0);
    MML_TRY(topn_exact_device(ctx, d_U, n_model_users, d_V, n_model_items, k, ru.data(), (int64_t)nr, n, candidates, n_cand,
                              n_ign > 0 ? rptr.data() : nullptr, n_ign > 0 ? rign.data() : nullptr,
                              ri.data(), rs.data(), rc.data(), launches));
    for (size_t t = 0; t < nr; t++) {
        const size_t b = (size_t)redo_users[t];
        memcpy(out_items + b * n_out, ri.data() + t * n_out, sizeof(int32_t) * n_out);
        memcpy(out_scores + b * n_out, rs.data() + t * n_out, sizeof(float) * n_out);
        out_counts[b] = rc[t];
    }
    return MML_OK;
}

// =================================================================================================
// Eval.Items.Evaluate on the device: ranking measures straight from the score rows
// =================================================================================================
// Reference: Eval/Items.cs:126-209 and Eval/Measures/{AUC,PrecisionAndRecall,NDCG,ReciprocalRank}.cs. For one test user
// the reference builds Recommend(user, candidates, n, ignore = training items) and walks that list once per measure.
// Every measure only needs the ranks of the user's correct items (test items among the candidates) in that list, so
// the list is never materialised: the correct items are ordered by (score desc, candidate position asc), every
// listed candidate finds by binary search how many correct items precede it, and a prefix sum over those counts gives
// each correct item's rank. One CTA per user; scores are the exact ones of score_tile_kernel.
struct EvalArgs {
    const float* scores; int64_t n_cand; int32_t n;          // scores: [batch][n_cand], -inf = ignored, -FLT_MAX = dropped
    const int32_t* pos_of; int32_t n_pos_of;                 // item -> candidate position or -1
    const int64_t* test_ptr; const int32_t* test_idx;        // correct-item rows of ALL test users
    int64_t b_lo;                                            // first test user of this batch
    float* c_score; int32_t* c_pos;                          // scratch, one slot per test entry: listed correct items
    float* s_score; int32_t* s_pos;                          //   ... sorted by rank
    uint32_t* bucket;                                        // scratch [test nnz + users]: bucket of user u starts at test_ptr[u] + u
    float* out; int32_t* used;                               // [users][8], [users]
};

__device__ __forceinline__ bool ranks_before(float sa, int32_t pa, float sb, int32_t pb)
{
    return sa > sb || (sa == sb && pa < pb);                 // OrderByDescending is stable: ties keep candidate order
}

constexpr int EVAL_T = 256;
__global__ void __launch_bounds__(EVAL_T) items_eval_kernel(const EvalArgs a)
{
    __shared__ int sh_m, sh_ncorrect, sh_listed, sh_ignored;
    const int b = blockIdx.x;
    const int64_t u = a.b_lo + b;
    const float* row = a.scores + (size_t)b * a.n_cand;
    const int64_t lo = a.test_ptr[u], hi = a.test_ptr[u + 1];
    float* cs = a.c_score + lo; int32_t* cp = a.c_pos + lo;
    float* ss = a.s_score + lo; int32_t* sp = a.s_pos + lo;
    uint32_t* bucket = a.bucket + lo + u;
    if (threadIdx.x == 0) { sh_m = 0; sh_ncorrect = 0; sh_listed = 0; sh_ignored = 0; }
    __syncthreads();
    // correct_items = test row  intersected with the candidates (Items.cs:152-153); those with a listed score are ranked
    for (int64_t e = lo + threadIdx.x; e < hi; e += EVAL_T) {
        const int32_t item = a.test_idx[e];
        const int32_t pos = (item >= 0 && item < a.n_pos_of) ? a.pos_of[item] : -1;
        if (pos < 0) continue;
        atomicAdd(&sh_ncorrect, 1);
        const float sc = row[pos];
        if (sc > -FLT_MAX) { const int slot = atomicAdd(&sh_m, 1); cs[slot] = sc; cp[slot] = pos; }
    }
    __syncthreads();
    const int m = sh_m;
    for (int i = threadIdx.x; i < m; i += EVAL_T) {
        int r = 0;
        const float si = cs[i]; const int32_t pi = cp[i];
        for (int j = 0; j < m; j++) r += ranks_before(cs[j], cp[j], si, pi) ? 1 : 0;
        ss[r] = si; sp[r] = pi;
    }
    for (int i = threadIdx.x; i <= m; i += EVAL_T) bucket[i] = 0u;
    __syncthreads();
    int listed = 0, ignored = 0;
    for (int64_t c = threadIdx.x; c < a.n_cand; c += EVAL_T) {
        const float sc = row[c];
        if (sc == -INFINITY) { ignored++; continue; }
        if (!(sc > -FLT_MAX)) continue;                      // Recommender.cs:70: only scores > float.MinValue are listed
        listed++;
        if (m == 0) continue;
        int l = 0, h = m;                                    // first correct item that does not precede candidate c
        while (l < h) {
            const int mid = (l + h) >> 1;
            if (ranks_before(ss[mid], sp[mid], sc, (int32_t)c)) l = mid + 1; else h = mid;
        }
        atomicAdd(&bucket[l], 1u);
    }
    atomicAdd(&sh_listed, listed); atomicAdd(&sh_ignored, ignored);
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int64_t L = sh_listed;
    const int64_t n_correct = sh_ncorrect;
    const int64_t num_cand_user = a.n_cand - sh_ignored;     // Items.cs:161-162
    float* o = a.out + (size_t)u * 8;
    for (int x = 0; x < 8; x++) o[x] = 0.f;
    if (n_correct == 0 || n_correct == num_cand_user) { a.used[u] = 0; return; }   // :154-155, :163-164
    const int64_t Lp = (a.n > 0) ? min((int64_t)a.n, L) : L; // prediction.Count
    const int64_t dropped = num_cand_user - Lp;              // :169
    int64_t hits = 0, h5 = 0, h10 = 0, after_sum = 0;
    double ap = 0.0, dcg = 0.0, rr = 0.0;
    int64_t rank = 0;
    const double ln2 = log(2.0);
    for (int t = 0; t < m; t++) {
        rank += bucket[t];                                   // listed candidates that do not come after correct item t
        if (rank > Lp) break;
        hits++;
        ap += (double)hits / (double)rank;                   // PrecisionAndRecall.AP
        dcg += 1.0 / (log((double)(rank + 1)) / ln2);        // NDCG: 1 / Math.Log(rank + 1, 2)
        if (hits == 1) rr = 1.0 / (double)rank;              // ReciprocalRank
        if (rank <= 5) h5++;
        if (rank <= 10) h10++;
        after_sum += Lp - rank;                              // list entries after this hit (relevant ones removed below)
    }
    double idcg = 0.0;
    for (int64_t i = 0; i < n_correct; i++) idcg += 1.0 / (log((double)(i + 2)) / ln2);
    // AUC.Compute: pairs (relevant, non-relevant) in the right order
    const int64_t missing = n_correct - hits;
    const int64_t eval_pairs = (num_cand_user - hits) * hits;
    double auc;
    if (eval_pairs == 0) auc = 0.5;                          // AUC.cs: checked before the consistency test below
    else if (dropped - missing < 0) { a.used[u] = -1; return; }   // the reference throws "Should not happen."
    else {
        const int64_t correct_pairs = after_sum - hits * (hits - 1) / 2 + hits * (dropped - missing);
        auc = (double)correct_pairs / (double)eval_pairs;
    }
    o[0] = (float)auc;
    o[1] = (float)(hits ? ap / (double)n_correct : 0.0);
    o[2] = (float)(dcg / idcg);
    o[3] = (float)rr;
    o[4] = (float)((double)h5 / 5.0);  o[5] = (float)((double)h10 / 10.0);
    o[6] = (float)((double)h5 / (double)n_correct); o[7] = (float)((double)h10 / (double)n_correct);
    a.used[u] = 1;
}

// d_U / d_V on the device, everything else on the host. out_measures: n_users x 8 = {AUC, MAP, NDCG, MRR, prec@5,
// prec@10, recall@5, recall@10} of each user; out_used: 1 = counted, 0 = skipped by the reference's rules.
int32_t items_eval_device(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                          const int32_t* users, int64_t n_users, const int32_t* candidates, int64_t n_cand,
                          const int64_t* test_ptr, const int32_t* test_idx,
                          const int64_t* ignore_ptr, const int32_t* ignore_idx, int32_t n,
                          float* out_measures, int32_t* out_used, int64_t* launches)
{
    cudaStream_t s = ctx->stream;
    if (n_users == 0) return MML_OK;
    MML_CHECK(n_cand > 0 && n_cand < ((int64_t)1 << 31), MML_ERR_ARG, "items_evaluate: bad candidate count");
    int32_t max_id = n_model_items - 1;
    for (int64_t c = 0; c < n_cand; c++) max_id = std::max(max_id, candidates[c]);
    const int32_t n_pos_of = max_id + 1;
    std::vector<int32_t> pos_of((size_t)std::max(n_pos_of, 1), -1), next((size_t)n_cand, -1);
    for (int64_t c = 0; c < n_cand; c++) {
        const int32_t item = candidates[c];
        MML_CHECK(item >= 0, MML_ERR_ARG, "items_evaluate: negative candidate id");
        MML_CHECK(pos_of[item] < 0, MML_ERR_ARG, "items_evaluate: candidate %d is listed twice", item);
        pos_of[item] = (int32_t)c;
    }
    const int64_t n_test = test_ptr[n_users];
    const int64_t n_ign = (ignore_ptr && ignore_idx) ? ignore_ptr[n_users] : 0;
    DevBuf<int32_t> d_cand, d_pos_of, d_next, d_users, d_ign_idx, d_test_idx, d_cp, d_sp, d_used;
    DevBuf<int64_t> d_ign_ptr, d_test_ptr;
    DevBuf<float> d_cs, d_ss, d_out;
    DevBuf<uint32_t> d_bucket;
    MML_TRY(d_cand.alloc(n_cand)); MML_TRY(d_pos_of.alloc(pos_of.size())); MML_TRY(d_next.alloc(n_cand));
    MML_TRY(d_users.alloc(n_users)); MML_TRY(d_test_ptr.alloc(n_users + 1)); MML_TRY(d_test_idx.alloc(n_test));
    MML_TRY(d_cp.alloc(n_test)); MML_TRY(d_sp.alloc(n_test)); MML_TRY(d_cs.alloc(n_test)); MML_TRY(d_ss.alloc(n_test));
    MML_TRY(d_bucket.alloc(n_test + n_users)); MML_TRY(d_out.alloc((size_t)n_users * 8)); MML_TRY(d_used.alloc(n_users));
    MML_CUDA(cudaMemcpyAsync(d_cand.p, candidates, sizeof(int32_t) * n_cand, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(d_pos_of.p, pos_of.data(), sizeof(int32_t) * pos_of.size(), cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(d_next.p, next.data(), sizeof(int32_t) * n_cand, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(d_users.p, users, sizeof(int32_t) * n_users, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(d_test_ptr.p, test_ptr, sizeof(int64_t) * (n_users + 1), cudaMemcpyHostToDevice, s));
    if (n_test > 0) MML_CUDA(cudaMemcpyAsync(d_test_idx.p, test_idx, sizeof(int32_t) * n_test, cudaMemcpyHostToDevice, s));
    if (n_ign > 0) {
        MML_TRY(d_ign_ptr.alloc(n_users + 1)); MML_TRY(d_ign_idx.alloc(n_ign));
        MML_CUDA(cudaMemcpyAsync(d_ign_ptr.p, ignore_ptr, sizeof(int64_t) * (n_users + 1), cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaMemcpyAsync(d_ign_idx.p, ignore_idx, sizeof(int32_t) * n_ign, cudaMemcpyHostToDevice, s));
    }
    int64_t B = std::max<int64_t>(1, ((int64_t)1 << 28) / n_cand);      // score buffer <= 1 GiB
    B = std::min<int64_t>(std::min<int64_t>(B, n_users), 65535 * (int64_t)TS);
    DevBuf<float> d_scores;
    MML_TRY(d_scores.alloc((size_t)B * n_cand));
    EvalArgs a{};
    a.scores = d_scores.p; a.n_cand = n_cand; a.n = n; a.pos_of = d_pos_of.p; a.n_pos_of = n_pos_of;
    a.test_ptr = d_test_ptr.p; a.test_idx = d_test_idx.p;
    a.c_score = d_cs.p; a.c_pos = d_cp.p; a.s_score = d_ss.p; a.s_pos = d_sp.p; a.bucket = d_bucket.p;
    a.out = d_out.p; a.used = d_used.p;
    for (int64_t b_lo = 0; b_lo < n_users; b_lo += B) {
        const int32_t nb = (int32_t)std::min<int64_t>(B, n_users - b_lo);
        dim3 grid((unsigned)ceil_div(n_cand, TS), (unsigned)ceil_div(nb, TS));
        score_tile_kernel<<<grid, 256, 0, s>>>(d_U, n_model_users, d_V, n_model_items, k, d_users.p + b_lo, nb, d_cand.p, n_cand, d_scores.p);
        if (n_ign > 0)
            mask_ignored_kernel<<<nb, 128, 0, s>>>(d_ign_ptr.p, d_ign_idx.p, (int32_t)b_lo, nb, d_pos_of.p, n_pos_of, d_next.p, n_cand, d_scores.p);
        a.b_lo = b_lo;
        items_eval_kernel<<<nb, EVAL_T, 0, s>>>(a);
        MML_CUDA(cudaGetLastError());
        if (launches) *launches += 3;
    }
    MML_CUDA(cudaMemcpyAsync(out_measures, d_out.p, sizeof(float) * (size_t)n_users * 8, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaMemcpyAsync(out_used, d_used.p, sizeof(int32_t) * n_users, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    for (int64_t u = 0; u < n_users; u++)
        MML_CHECK(out_used[u] >= 0, MML_ERR_ARG, "items_evaluate: user %d has test items among its ignored (training) items "
                  "(AUC.Compute throws \"Should not happen.\")", users[u]);
    return MML_OK;
}

}  // namespace mml

using namespace mml;

extern "C" int32_t mml_topn_set_mode(int32_t mode)
{
    MML_CHECK(mode >= MML_TOPN_AUTO && mode <= MML_TOPN_TENSOR, MML_ERR_ARG, "mml_topn_set_mode: unknown mode %d", mode);
    g_topn_mode = mode;
    return MML_OK;
}

extern "C" int32_t mml_topn_set_filter(int32_t kind)
{
    MML_CHECK(kind == MML_TOPN_FILTER_BF16 || kind == MML_TOPN_FILTER_TF32, MML_ERR_ARG, "mml_topn_set_filter: unknown kind %d", kind);
    topn_tc_set_filter(kind);
    return MML_OK;
}

extern "C" int32_t mml_topn_last_stats(int64_t* users_tensor_path, int64_t* users_exact_path, float* tensor_path_ms)
{
    if (users_tensor_path) *users_tensor_path = g_stat_tc;
    if (users_exact_path) *users_exact_path = g_stat_exact;
    if (tensor_path_ms) *tensor_path_ms = g_stat_ms;
    return MML_OK;
}

extern "C" int32_t mml_topn_mf(mml_ctx* hctx, const float* user_factors, int32_t n_model_users,
                               const float* item_factors, int32_t n_model_items, int32_t k,
                               const int32_t* users, int64_t n_users, int32_t n,
                               const int32_t* candidates, int64_t n_cand,
                               const int64_t* ignore_ptr, const int32_t* ignore_idx,
                               int32_t* out_items, float* out_scores, int32_t* out_counts)
{
    MML_LOCK(mml::ctx_of(hctx));
    MML_CHECK(hctx && user_factors && item_factors && (n_users == 0 || (users && out_items && out_scores && out_counts)),
              MML_ERR_ARG, "mml_topn_mf: NULL argument");
    MML_CHECK(k >= 1 && n_model_users >= 0 && n_model_items >= 0 && n_users >= 0 && (n > 0 || n == -1), MML_ERR_ARG,
              "mml_topn_mf: bad sizes (n must be > 0 or -1)");
    Ctx* ctx = ctx_of(hctx);
    if (ctx->is_root()) {   // users sharded over the GPUs (contiguous ranges of the list), factors uploaded to each
        const int64_t N = (int64_t)ctx->peers.size();
        const int64_t nc = candidates ? n_cand : n_model_items;
        const int64_t n_out = n < 0 ? nc : std::min<int64_t>(n, nc);
        return on_ranks((int)N, [&](int x) -> int32_t {
            const int64_t lo = n_users * x / N, hi = n_users * (x + 1) / N;
            if (hi <= lo) return MML_OK;
            std::vector<int64_t> ip;
            const bool ign = ignore_ptr && ignore_idx;
            if (ign) { ip.resize((size_t)(hi - lo + 1)); for (int64_t t = lo; t <= hi; t++) ip[(size_t)(t - lo)] = ignore_ptr[t] - ignore_ptr[lo]; }
            return mml_topn_mf(ctx->peers[(size_t)x], user_factors, n_model_users, item_factors, n_model_items, k, users + lo, hi - lo, n,
                               candidates, n_cand, ign ? ip.data() : nullptr, ign ? ignore_idx + ignore_ptr[lo] : nullptr,
                               out_items + lo * n_out, out_scores + lo * n_out, out_counts + lo);
        });
    }
    MML_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevBuf<float> dU, dV;
    MML_TRY(dU.alloc((size_t)n_model_users * k)); MML_TRY(dV.alloc((size_t)n_model_items * k));
    MML_CUDA(cudaMemcpyAsync(dU.p, user_factors, sizeof(float) * (size_t)n_model_users * k, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(dV.p, item_factors, sizeof(float) * (size_t)n_model_items * k, cudaMemcpyHostToDevice, s));
    return topn_device(ctx, dU.p, n_model_users, dV.p, n_model_items, k, users, n_users, n, candidates, n_cand,
                       ignore_ptr, ignore_idx, out_items, out_scores, out_counts, nullptr);
}

extern "C" int32_t mml_items_evaluate_mf(mml_ctx* hctx, const float* user_factors, int32_t n_model_users,
                                         const float* item_factors, int32_t n_model_items, int32_t k,
                                         const int32_t* test_users, int64_t n_test_users,
                                         const int32_t* candidates, int64_t n_cand,
                                         const int64_t* test_ptr, const int32_t* test_idx,
                                         const int64_t* ignore_ptr, const int32_t* ignore_idx, int32_t n,
                                         float* out_measures, int32_t* out_used)
{
    MML_LOCK(mml::ctx_of(hctx));
    MML_CHECK(hctx && user_factors && item_factors && candidates && test_ptr &&
              (n_test_users == 0 || (test_users && out_measures && out_used)), MML_ERR_ARG, "mml_items_evaluate_mf: NULL argument");
    MML_CHECK(k >= 1 && n_model_users >= 0 && n_model_items >= 0 && n_test_users >= 0 && (n > 0 || n == -1), MML_ERR_ARG,
              "mml_items_evaluate_mf: bad sizes (n must be > 0 or -1)");
    MML_CHECK(test_ptr[n_test_users] == 0 || test_idx, MML_ERR_ARG, "mml_items_evaluate_mf: NULL test_idx");
    Ctx* ctx = ctx_of(hctx);
    if (ctx->is_root())
        return mml_items_evaluate_mf(ctx->peers[0], user_factors, n_model_users, item_factors, n_model_items, k, test_users, n_test_users,
                                     candidates, n_cand, test_ptr, test_idx, ignore_ptr, ignore_idx, n, out_measures, out_used);
    MML_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevBuf<float> dU, dV;
    MML_TRY(dU.alloc((size_t)n_model_users * k)); MML_TRY(dV.alloc((size_t)n_model_items * k));
    MML_CUDA(cudaMemcpyAsync(dU.p, user_factors, sizeof(float) * (size_t)n_model_users * k, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(dV.p, item_factors, sizeof(float) * (size_t)n_model_items * k, cudaMemcpyHostToDevice, s));
    return items_eval_device(ctx, dU.p, n_model_users, dV.p, n_model_items, k, test_users, n_test_users, candidates, n_cand,
                             test_ptr, test_idx, ignore_ptr, ignore_idx, n, out_measures, out_used, nullptr);
}
