// topn_tc.cu -- Recommend() scoring on the tcgen05 tensor cores with the per-user top-k selection fused
// into the epilogue (K9 of SURVEY.md section 2d).
//
// Reference: Recommender.Recommend (Recommender.cs:52-103) over ItemRecommendation/MF.cs:151-157 scores
// (DataType/MatrixExtensions.cs:224-241: sequential fp32 multiply-then-add dot product).
//
// The reference's result is defined by EXACT fp32 scores, so the tensor cores are used as a filter:
//   1. score_select_kernel: S = U_tile * V_tile^T as TF32 tcgen05.mma (fp32 factor rows are fed as they are;
//      the tensor core uses the upper 19 bits), 256 users x 128 candidates per step, accumulators in TMEM,
//      operands staged in shared memory by TMA (128-byte swizzle). The epilogue warps read the accumulators with
//      tcgen05.ld; every thread owns one user and keeps the CAP best approximate scores it has seen (not in the
//      user's ignore list) -- the 400 GB score matrix of config 5 never exists.
//   2. finalize_kernel: with a_n the n-th best approximate score and d >= |approximate - exact| (a bound from
//      the TF32 input truncation: d = c * |u| * max|v|), every member of the exact top n has an approximate
//      score >= a_n - 2d. If the kept list reaches below that cut-off the superset is complete: its members
//      are re-scored exactly (the reference's arithmetic) and ordered by (score desc, candidate position asc).
//      Otherwise the user is flagged and the caller re-runs it on the exact CUDA-core path (topn.cu).
// Results are therefore bit-identical to the exact path, whatever the tensor cores round.
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <cfloat>
#include <cmath>

namespace mml {

constexpr int TC_ROWS = 256;                  // users per CTA: two 128-row MMA tiles sharing every V tile
constexpr int TC_N = 128;                     // candidates per MMA tile
constexpr int TC_KC = 32;                     // floats per K chunk = one 128-byte swizzle atom
constexpr int TC_CHUNK_BYTES = 128 * 128;     // 128 rows x 128 bytes
constexpr int TC_CAP = 32;                    // approximate scores kept per user (and per candidate split)
constexpr int TC_MAX_N = 16;                  // largest n served by this path
constexpr int TC_THREADS = 320;               // warp 0: TMA, warp 1: MMA issue + TMEM, warps 2-9: epilogue
constexpr float TC_ERR_C = 0.0025f;           // |tf32 score - exact| <= TC_ERR_C * |u| * |v| (2^-9 truncation + slack)
constexpr int TC_MAX_SPLITS = 32;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// Bounded wait: a protocol bug must end as an error, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* err)
{
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) { atomicExch(err, 1u); __threadfence(); asm volatile("trap;"); }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (sm_100 version 1).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128.
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((128u >> 4) << 24);

struct TcArgs {
    int32_t n_rows;                 // users in the batch
    int32_t n_cand;                 // candidates (positions)
    int32_t n_model_items;
    int32_t kc;                     // K chunks (kp / 32)
    int32_t stages;                 // V ring depth
    int32_t n_tiles, tiles_per_split, splits;
    const uint8_t* row_ok;          // [n_rows] user id inside the model
    const int32_t* cand;            // [n_cand] item id per position or NULL (identity)
    const int64_t* ign_ptr;         // [n_rows + 1] or NULL
    const int32_t* ign_idx;         // item ids, ascending inside a row
    float* list_s; int32_t* list_p; // [n_rows][splits][CAP]
    int32_t* list_n;                // [n_rows][splits]
    uint32_t* err;
};

struct TcIns { float thr; int cnt; };

// Slow path of the epilogue: candidate at position pos passed the threshold of this user's list.
__device__ __noinline__ TcIns tc_consider(const TcArgs& a, float s, int pos, float* ls, int32_t* lp, float thr, int cnt,
                                          int64_t ig_lo, int64_t ig_hi)
{
    TcIns r; r.thr = thr; r.cnt = cnt;
    if (pos >= a.n_cand) return r;
    const int32_t item = a.cand ? a.cand[pos] : pos;
    if ((uint32_t)item >= (uint32_t)a.n_model_items) return r;      // Predict = float.MinValue: never qualifies
    while (ig_lo < ig_hi) {                                          // ignore_items.Contains(item)
        const int64_t mid = (ig_lo + ig_hi) >> 1;
        const int32_t x = a.ign_idx[mid];
        if (x == item) return r;
        if (x < item) ig_lo = mid + 1; else ig_hi = mid;
    }
    int j = cnt < TC_CAP ? cnt : TC_CAP - 1;
    while (j > 0 && ls[j - 1] < s) { ls[j] = ls[j - 1]; lp[j] = lp[j - 1]; j--; }
    ls[j] = s; lp[j] = pos;
    if (cnt < TC_CAP) r.cnt = cnt + 1;
    if (r.cnt == TC_CAP) r.thr = ls[TC_CAP - 1];
    return r;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
score_select_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_v, const TcArgs a)
{
    extern __shared__ uint8_t tc_smem_raw[];
    __shared__ uint64_t bars[2 * 8 + 1 + 4];      // full[8], empty[8], u_full, tmem_full[2], tmem_empty[2]
    __shared__ uint32_t tmem_base_s;
    const uint32_t smem0 = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_u = smem0;                                               // [2][kc] chunks
    const uint32_t smem_v = smem0 + 2u * a.kc * TC_CHUNK_BYTES;                  // [stages] chunks
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * 8, bar_u = bar_full + 16 * 8;
    const uint32_t bar_tfull = bar_full + 17 * 8, bar_tempty = bar_full + 19 * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp = blockIdx.y;
    const int t_begin = sp * a.tiles_per_split, t_end = min(t_begin + a.tiles_per_split, a.n_tiles);
    const int row0 = blockIdx.x * TC_ROWS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_u, 1);
        for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {    // TMEM: 512 columns = 2 buffers x 2 row halves x 128 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {    // ===== TMA producer =====
            mbar_expect_tx(bar_u, 2u * a.kc * TC_CHUNK_BYTES);
            for (int h = 0; h < 2; h++)
                for (int c = 0; c < a.kc; c++)
                    tma_load_2d(smem_u + (uint32_t)(h * a.kc + c) * TC_CHUNK_BYTES, &map_u, bar_u, c * TC_KC, row0 + h * 128);
            int stage = 0; uint32_t phase = 0;
            for (int t = t_begin; t < t_end; t++)
                for (int c = 0; c < a.kc; c++) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u, a.err);
                    mbar_expect_tx(bar_full + 8 * stage, TC_CHUNK_BYTES);
                    tma_load_2d(smem_v + (uint32_t)stage * TC_CHUNK_BYTES, &map_v, bar_full + 8 * stage, c * TC_KC, t * TC_N);
                    if (++stage == a.stages) { stage = 0; phase ^= 1u; }
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {    // ===== MMA issuer =====
            mbar_wait(bar_u, 0, a.err);
            tc_fence_after();
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int t = t_begin; t < t_end; t++, it++) {
                const int buf = it & 1;
                const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(bar_tempty + 8 * buf, aphase ^ 1u, a.err);
                tc_fence_after();
                for (int c = 0; c < a.kc; c++) {
                    mbar_wait(bar_full + 8 * stage, phase, a.err);
                    tc_fence_after();
                    const uint32_t vb = smem_v + (uint32_t)stage * TC_CHUNK_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t ub = smem_u + (uint32_t)(h * a.kc + c) * TC_CHUNK_BYTES;
                        const uint32_t d = tmem_base + (uint32_t)((buf * 2 + h) * TC_N);
#pragma unroll
                        for (int kk = 0; kk < 4; kk++)
                            tc_mma_tf32(d, tc_smem_desc(ub + kk * 32), tc_smem_desc(vb + kk * 32), TC_IDESC, (c | kk) != 0 ? 1u : 0u);
                    }
                    tc_commit(bar_empty + 8 * stage);          // frees the V chunk when these MMAs have read it
                    if (++stage == a.stages) { stage = 0; phase ^= 1u; }
                }
                tc_commit(bar_tfull + 8 * buf);                // accumulators of this tile complete
            }
        }
    } else {
        // ===== epilogue: 8 warps, thread <-> user row =====
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const int h = (warp - 2) >> 2;                         // row half
        const int row = row0 + h * 128 + q * 32 + lane;
        const bool ok = row < a.n_rows && a.row_ok[row];
        float thr = ok ? -INFINITY : INFINITY;
        int cnt = 0;
        const size_t lbase = ((size_t)(ok ? row : 0) * a.splits + sp) * TC_CAP;
        float* ls = a.list_s + lbase; int32_t* lp = a.list_p + lbase;
        int64_t ig_lo = 0, ig_hi = 0;
        if (ok && a.ign_ptr) { ig_lo = a.ign_ptr[row]; ig_hi = a.ign_ptr[row + 1]; }
        int it = 0;
        for (int t = t_begin; t < t_end; t++, it++) {
            const int buf = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(bar_tfull + 8 * buf, aphase, a.err);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + h) * TC_N);
#pragma unroll 1
            for (int j = 0; j < TC_N / 32; j++) {
                uint32_t v[32];
                __syncwarp();
                tc_ld32(taddr + j * 32, v);
                tc_ld_wait();
                float m = __uint_as_float(v[0]);
#pragma unroll
                for (int c = 1; c < 32; c++) m = fmaxf(m, __uint_as_float(v[c]));
                if (m > thr) {
                    const int pos0 = t * TC_N + j * 32;
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const float s = __uint_as_float(v[c]);
                        if (s > thr) {
                            const TcIns r = tc_consider(a, s, pos0 + c, ls, lp, thr, cnt, ig_lo, ig_hi);
                            thr = r.thr; cnt = r.cnt;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
        if (ok) a.list_n[(size_t)row * a.splits + sp] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- staging: dense, zero-padded operand panels + row norms --------------------------------------------------
// dst[r][0..kp) = src[id(r)][0..k) (zero beyond k, zero row for ids outside the model); norm[r] = |row|;
// ok[r] = id inside the model; *max_norm_bits = max over rows (float bits; NaN sorts above +inf).
__global__ void tc_stage_rows_kernel(const float* __restrict__ src, int32_t n_src_rows, int32_t k,
                                     const int32_t* __restrict__ ids, int32_t n, int32_t kp,
                                     float* __restrict__ dst, float* __restrict__ norm, uint8_t* __restrict__ ok,
                                     uint32_t* __restrict__ max_norm_bits)
{
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n; r += stride) {
        const int32_t id = ids ? ids[r] : (int32_t)r;
        const bool valid = (uint32_t)id < (uint32_t)n_src_rows;
        float ss = 0.f;
        for (int f = lane; f < kp; f += 32) {
            const float x = (valid && f < k) ? src[(size_t)id * k + f] : 0.f;
            dst[(size_t)r * kp + f] = x;
            ss = fmaf(x, x, ss);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
        if (lane == 0) {
            const float nr = sqrtf(ss);
            if (norm) norm[r] = nr;
            if (ok) ok[r] = valid ? 1 : 0;
            if (max_norm_bits) atomicMax(max_norm_bits, __float_as_uint(nr));
        }
    }
}

// row of every ignore entry (binary search in the CSR pointers) + "rows are ascending" check
__global__ void tc_ignore_rows_kernel(const int64_t* __restrict__ ptr, int32_t n_rows, const int32_t* __restrict__ idx, int64_t total,
                                      uint32_t* __restrict__ row_of, uint32_t* __restrict__ unsorted)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        int32_t lo = 0, hi = n_rows;                  // last row with ptr[row] <= t
        while (hi - lo > 1) { const int32_t mid = (lo + hi) >> 1; if (ptr[mid] <= t) lo = mid; else hi = mid; }
        row_of[t] = (uint32_t)lo;
        if (t + 1 < ptr[lo + 1] && idx[t] > idx[t + 1]) atomicExch(unsorted, 1u);
    }
}

__global__ void tc_bias_items_kernel(const int32_t* __restrict__ idx, int64_t total, uint32_t* __restrict__ key)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) key[t] = (uint32_t)idx[t] ^ 0x80000000u;      // order-preserving for negative ids
}
__global__ void tc_unbias_items_kernel(const uint32_t* __restrict__ key, int64_t total, int32_t* __restrict__ idx)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) idx[t] = (int32_t)(key[t] ^ 0x80000000u);
}

// ---- finalize: exact re-scoring of the candidate superset, one warp per user ---------------------------------
constexpr int FIN_WARPS = 4;

struct FinArgs {
    const float* U; const float* V; int32_t k;            // original model matrices
    const int32_t* users;                                   // [n_rows] user ids
    const int32_t* cand;                                    // or NULL
    const float* list_s; const int32_t* list_p; const int32_t* list_n;
    const float* unorm; const uint32_t* vmax_bits;
    const uint8_t* row_ok;
    int32_t n_rows, splits, n, n_out;
    int32_t* out_items; float* out_scores; int32_t* out_counts; uint8_t* redo;
};

__global__ void __launch_bounds__(FIN_WARPS * 32) tc_finalize_kernel(const FinArgs a)
{
    extern __shared__ uint8_t fin_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int E_max = a.splits * TC_CAP;
    float* ap = reinterpret_cast<float*>(fin_smem) + (size_t)warp * E_max * 3;      // approximate scores
    int32_t* pp = reinterpret_cast<int32_t*>(ap + E_max);                            // positions
    float* ex = reinterpret_cast<float*>(pp + E_max);                                // exact scores
    const int b = blockIdx.x * FIN_WARPS + warp;
    if (b >= a.n_rows) return;
    if (!a.row_ok[b]) { if (lane == 0) { a.out_counts[b] = 0; a.redo[b] = 0; } return; }
    // 1. gather the per-split lists; full_min = largest "smallest kept score" over the lists that are full
    int E = 0;
    float full_min = -INFINITY;
    bool any_full = false;
    for (int s = 0; s < a.splits; s++) {
        const int c = a.list_n[(size_t)b * a.splits + s];
        const size_t base = ((size_t)b * a.splits + s) * TC_CAP;
        for (int e = lane; e < c; e += 32) { ap[E + e] = a.list_s[base + e]; pp[E + e] = a.list_p[base + e]; }
        if (c == TC_CAP) { any_full = true; full_min = fmaxf(full_min, a.list_s[base + TC_CAP - 1]); }
        E += c;
    }
    __syncwarp();
    if (E == 0) { if (lane == 0) { a.out_counts[b] = 0; a.redo[b] = 0; } return; }
    const float delta = TC_ERR_C * a.unorm[b] * __uint_as_float(*a.vmax_bits) + 1e-30f;
    // 2. a_n = n-th best approximate score (rank by score desc, list index asc)
    float a_n = -INFINITY;
    if (E >= a.n) {
        for (int e = lane; e < E; e += 32) {
            const float se = ap[e];
            int rank = 0;
            for (int o = 0; o < E; o++) { const float so = ap[o]; rank += (so > se || (so == se && o < e)) ? 1 : 0; }
            if (rank == a.n - 1) a_n = se;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a_n = fmaxf(a_n, __shfl_xor_sync(0xffffffffu, a_n, d));
    }
    const float cut = a_n - 2.f * delta;        // -inf when fewer than n candidates were seen at all
    // 3. is the superset complete? (a full list may have dropped candidates scoring up to its smallest entry)
    if (!(delta < INFINITY) || (any_full && full_min >= cut)) { if (lane == 0) { a.redo[b] = 1; a.out_counts[b] = 0; } return; }
    // 4. exact scores of the finalists: sequential fp32 multiply, then add (MatrixExtensions.cs:234-238)
    const float* urow = a.U + (size_t)a.users[b] * a.k;
    for (int e = lane; e < E; e += 32) {
        float s = -INFINITY;
        if (ap[e] >= cut) {
            const int32_t item = a.cand ? a.cand[pp[e]] : pp[e];
            const float* vrow = a.V + (size_t)item * a.k;
            s = 0.f;
            for (int f = 0; f < a.k; f++) s = __fadd_rn(s, __fmul_rn(urow[f], vrow[f]));
            if (!(s > -FLT_MAX)) s = -INFINITY;       // score > float.MinValue (Recommender.cs:72,86)
        }
        ex[e] = s;
    }
    __syncwarp();
    // 5. order by (exact score desc, candidate position asc); the best n go out
    int qualified = 0;
    for (int e = lane; e < E; e += 32) {
        const float se = ex[e];
        if (!(se > -INFINITY)) continue;
        qualified++;
        const int pe = pp[e];
        int rank = 0;
        for (int o = 0; o < E; o++) { const float so = ex[o]; rank += (so > se || (so == se && pp[o] < pe)) ? 1 : 0; }
        if (rank < a.n_out) {
            a.out_items[(size_t)b * a.n_out + rank] = a.cand ? a.cand[pe] : pe;
            a.out_scores[(size_t)b * a.n_out + rank] = se;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) qualified += __shfl_xor_sync(0xffffffffu, qualified, d);
    if (lane == 0) { a.out_counts[b] = min(qualified, a.n_out); a.redo[b] = 0; }
}

// ---- host ------------------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int32_t make_panel_map(CUtensorMap* map, float* base, int64_t rows, int32_t kp)
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        MML_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
        MML_CHECK(p != nullptr && qr == cudaDriverEntryPointSuccess, MML_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        fn = (encode_tiled_fn)p;
    }
    const cuuint64_t dims[2] = { (cuuint64_t)kp, (cuuint64_t)rows };
    const cuuint64_t strides[1] = { (cuuint64_t)kp * sizeof(float) };
    const cuuint32_t box[2] = { TC_KC, 128 };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MML_CHECK(r == CUDA_SUCCESS, MML_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MML_OK;
}

static inline int tc_grid(int64_t n, int threads = 256)
{
    return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n, threads), 1), 148 * 16);
}

bool topn_tc_eligible(int32_t k, int32_t n, int64_t n_cand)
{
    return n >= 1 && n <= TC_MAX_N && k >= 1 && k <= 128 && n_cand >= 1;
}

// Top n of a user batch on the tensor-core path. d_users / d_cand (or NULL) / d_ign_ptr / d_ign_idx (or NULL): device.
// d_redo[b] = 1 for users whose candidate superset could not be proven complete (caller re-runs them exactly).
// Outputs (device): d_out_items/d_out_scores [n_users x n_out], d_out_counts [n_users].
int32_t topn_tc_batch(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                      const int32_t* d_users, int32_t n_users, int32_t n, int32_t n_out,
                      const int32_t* d_cand, int32_t n_cand,
                      const int64_t* d_ign_ptr, int32_t* d_ign_idx, int64_t n_ign,
                      int32_t* d_out_items, float* d_out_scores, int32_t* d_out_counts, uint8_t* d_redo,
                      int64_t* launches)
{
    cudaStream_t s = ctx->stream;
    const int32_t kp = (int32_t)ceil_div(k, TC_KC) * TC_KC, kc = kp / TC_KC;
    const int64_t rows_pad = ceil_div(n_users, TC_ROWS) * TC_ROWS, cand_pad = ceil_div(n_cand, TC_N) * TC_N;
    DevBuf<float> Ub, Vb, unorm; DevBuf<uint8_t> row_ok; DevBuf<uint32_t> vmax, err;
    MML_TRY(Ub.alloc((size_t)rows_pad * kp)); MML_TRY(Vb.alloc((size_t)cand_pad * kp));
    MML_TRY(unorm.alloc(n_users)); MML_TRY(row_ok.alloc(n_users)); MML_TRY(vmax.alloc(1)); MML_TRY(err.alloc(1));
    MML_CUDA(cudaMemsetAsync(vmax.p, 0, sizeof(uint32_t), s));
    MML_CUDA(cudaMemsetAsync(err.p, 0, sizeof(uint32_t), s));
    if (rows_pad > n_users) MML_CUDA(cudaMemsetAsync(Ub.p + (size_t)n_users * kp, 0, sizeof(float) * (size_t)(rows_pad - n_users) * kp, s));
    if (cand_pad > n_cand) MML_CUDA(cudaMemsetAsync(Vb.p + (size_t)n_cand * kp, 0, sizeof(float) * (size_t)(cand_pad - n_cand) * kp, s));
    tc_stage_rows_kernel<<<tc_grid((int64_t)n_users * 32), 256, 0, s>>>(d_U, n_model_users, k, d_users, n_users, kp, Ub.p, unorm.p, row_ok.p, nullptr);
    tc_stage_rows_kernel<<<tc_grid((int64_t)n_cand * 32), 256, 0, s>>>(d_V, n_model_items, k, d_cand, n_cand, kp, Vb.p, nullptr, nullptr, vmax.p);
    MML_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
    // ignore lists must be ascending inside a row for the epilogue's binary search
    if (d_ign_ptr && n_ign > 0) {
        DevBuf<uint32_t> row_of, flag, key, t1, t2;
        MML_TRY(row_of.alloc(n_ign)); MML_TRY(flag.alloc(1));
        MML_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(uint32_t), s));
        tc_ignore_rows_kernel<<<tc_grid(n_ign), 256, 0, s>>>(d_ign_ptr, n_users, d_ign_idx, n_ign, row_of.p, flag.p);
        MML_CUDA(cudaGetLastError());
        uint32_t unsorted = 0;
        MML_CUDA(cudaMemcpyAsync(&unsorted, flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaStreamSynchronize(s));
        if (launches) *launches += 1;
        if (unsorted) {
            MML_TRY(key.alloc(n_ign)); MML_TRY(t1.alloc(n_ign)); MML_TRY(t2.alloc(n_ign));
            tc_bias_items_kernel<<<tc_grid(n_ign), 256, 0, s>>>(d_ign_idx, n_ign, key.p);
            MML_TRY(radix_sort_pairs(key.p, row_of.p, t1.p, t2.p, n_ign, 32, s));                                   // by item
            MML_TRY(radix_sort_pairs(row_of.p, key.p, t1.p, t2.p, n_ign, bits_for((uint32_t)std::max(n_users - 1, 1)), s));   // by row, stable
            tc_unbias_items_kernel<<<tc_grid(n_ign), 256, 0, s>>>(key.p, n_ign, d_ign_idx);
            MML_CUDA(cudaGetLastError());
            if (launches) *launches += 10;
        }
    }
    CUtensorMap map_u, map_v;
    MML_TRY(make_panel_map(&map_u, Ub.p, rows_pad, kp));
    MML_TRY(make_panel_map(&map_v, Vb.p, cand_pad, kp));
    // candidate splits: enough CTAs to fill the GPU when the batch has few user tiles
    const int row_tiles = (int)(rows_pad / TC_ROWS), n_tiles = (int)(cand_pad / TC_N);
    int splits = (int)std::min<int64_t>(std::min<int64_t>(ceil_div(ctx->sm_count, row_tiles), TC_MAX_SPLITS), n_tiles);
    splits = std::max(splits, 1);
    const int tps = (int)ceil_div(n_tiles, splits);
    splits = (int)ceil_div(n_tiles, tps);
    DevBuf<float> list_s; DevBuf<int32_t> list_p, list_n;
    MML_TRY(list_s.alloc((size_t)n_users * splits * TC_CAP)); MML_TRY(list_p.alloc((size_t)n_users * splits * TC_CAP));
    MML_TRY(list_n.alloc((size_t)n_users * splits));
    MML_CUDA(cudaMemsetAsync(list_n.p, 0, list_n.bytes(), s));
    int max_optin = 0;
    MML_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    const int64_t fixed = 2ll * kc * TC_CHUNK_BYTES + 1024 + 512;      // U panels + alignment slack + static barriers
    const int stages = (int)std::min<int64_t>(8, (max_optin - fixed) / TC_CHUNK_BYTES);
    MML_CHECK(stages >= 2, MML_ERR_UNSUPPORTED, "topn: shared memory too small for the tcgen05 path");
    const size_t smem = (size_t)(2ll * kc + stages) * TC_CHUNK_BYTES + 1024;
    MML_CUDA(cudaFuncSetAttribute((const void*)score_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TcArgs a{};
    a.n_rows = n_users; a.n_cand = n_cand; a.n_model_items = n_model_items; a.kc = kc; a.stages = stages;
    a.n_tiles = n_tiles; a.tiles_per_split = tps; a.splits = splits;
    a.row_ok = row_ok.p; a.cand = d_cand; a.ign_ptr = (d_ign_ptr && n_ign > 0) ? d_ign_ptr : nullptr; a.ign_idx = d_ign_idx;
    a.list_s = list_s.p; a.list_p = list_p.p; a.list_n = list_n.p; a.err = err.p;
    score_select_kernel<<<dim3(row_tiles, splits), TC_THREADS, smem, s>>>(map_u, map_v, a);
    MML_CUDA(cudaGetLastError());
    FinArgs f{};
    f.U = d_U; f.V = d_V; f.k = k; f.users = d_users; f.cand = d_cand;
    f.list_s = list_s.p; f.list_p = list_p.p; f.list_n = list_n.p; f.unorm = unorm.p; f.vmax_bits = vmax.p; f.row_ok = row_ok.p;
    f.n_rows = n_users; f.splits = splits; f.n = n; f.n_out = n_out;
    f.out_items = d_out_items; f.out_scores = d_out_scores; f.out_counts = d_out_counts; f.redo = d_redo;
    const size_t fsmem = (size_t)FIN_WARPS * splits * TC_CAP * 12;
    MML_CUDA(cudaFuncSetAttribute((const void*)tc_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    tc_finalize_kernel<<<(unsigned)ceil_div(n_users, FIN_WARPS), FIN_WARPS * 32, fsmem, s>>>(f);
    MML_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
    MML_CUDA(cudaStreamSynchronize(s));      // staging buffers are released on return
    uint32_t h_err = 0;
    MML_CUDA(cudaMemcpy(&h_err, err.p, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    MML_CHECK(h_err == 0, MML_ERR_CUDA, "topn: tcgen05 pipeline timed out");
    return MML_OK;
}

}  // namespace mml
