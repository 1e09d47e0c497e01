// wrmf.cu -- WRMF (Hu / Koren / Volinsky) ALS epoch on sm_100a.
//
// Reference: ItemRecommendation/WRMF.cs:68-156 (Iterate, Optimize, ComputeSquareMatrix),
// ItemRecommendation/MF.cs:51-67 (InitModel/Train), Data/PosOnlyFeedback.cs:35-83 and
// DataType/SparseBooleanMatrix.cs:37-91 (user / item matrices: sets, duplicate events collapse).
//
// Per half-sweep (W <- argmin given H):
//   HH = sum_i h_i h_i^T                       (k x k; float products accumulated in double, :94-108)
//   for every row u:  A = HH + alpha * sum_{i in S_u} h_i h_i^T + lambda * I   (double, :113-145)
//                     b = (1 + alpha) * sum_{i in S_u} h_i
//                     w_u = (float)(A^-1 b)    (the reference inverts with MathNet's LU and multiplies;
//                                               A is SPD, so a Cholesky solve gives the same w_u)
// The arithmetic is kept in the reference's precision -- float products, double sums, double solve -- so
// the factor rows agree with the oracle far inside the 1e-4 gate (B200 has full-rate FP64 on the SMs).
// One CTA per row; rows are handed out through an atomic counter, longest rows first.
//
// Round-1 status: the per-row accumulation runs on the FP64 CUDA cores; the tcgen05 (TF32 split) Gram /
// gather-SYRK variant named by the north star is not written yet ("what comes next").
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <new>
#include <numeric>
#include <vector>

namespace mml {

struct Feedback {
    Ctx* ctx = nullptr;
    int32_t max_user = -1, max_item = -1;
    int64_t n_events = 0, nnz = 0;
    DevBuf<uint32_t> user_ptr, item_ptr;   // [n_users + 1], [n_items + 1]
    DevBuf<int32_t> user_cols, item_rows;  // [nnz] items of each user (ascending), users of each item (ascending)
    int32_t n_users() const { return max_user + 1; }
    int32_t n_items() const { return max_item + 1; }
};

struct Wrmf {
    Ctx* ctx = nullptr;
    Feedback* fb = nullptr;
    int32_t k = 0;
    double alpha = 1.0, reg = 0.015;
    DevBuf<float> U, V;                    // [n_users x k], [n_items x k] row-major, as the reference's Matrix<float>
    DevBuf<double> HH, HH_part;
    DevBuf<int32_t> order_u, order_i;      // this rank's rows by descending nnz (work queue order)
    int32_t n_local_u = 0, n_local_i = 0;  // rows this rank solves in each half-sweep
    std::vector<int32_t> range_u, range_i; // [world + 1] multi-GPU: rank r solves rows [range[r], range[r + 1])
    DevBuf<unsigned> counter;
    bool has_model = false;
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    struct WrmfTcWork* tc_work = nullptr;  // tensor-path buffers (wrmf_tc.cu), grow-only
};

static inline int grid_n(int64_t n, int threads = 256)
{
    return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n, threads), 1), 148 * 16);
}

// ---- feedback build -------------------------------------------------------------------------------
__global__ void pair_head_kernel(const uint32_t* __restrict__ u_sorted, const uint32_t* __restrict__ i_sorted, int64_t n,
                                 uint32_t* __restrict__ head)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride)
        head[t] = (t == 0 || u_sorted[t] != u_sorted[t - 1] || i_sorted[t] != i_sorted[t - 1]) ? 1u : 0u;
}

__global__ void pair_compact_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ pos,
                                    const uint32_t* __restrict__ u_sorted, const uint32_t* __restrict__ i_sorted, int64_t n,
                                    uint32_t* __restrict__ u_out, uint32_t* __restrict__ i_out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride)
        if (head[t]) { u_out[pos[t]] = u_sorted[t]; i_out[pos[t]] = i_sorted[t]; }
}

__global__ void nnz_key_kernel(const uint32_t* __restrict__ ptr, int32_t n_rows, uint32_t* __restrict__ key, uint32_t* __restrict__ val)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n_rows; t += stride) { key[t] = 0xFFFFFFFFu - (ptr[t + 1] - ptr[t]); val[t] = (uint32_t)t; }
}

// sorts (a, b) pairs by (a, b): stable LSD over b then over a. Result in a/b.
static int32_t sort_pairs_lex(Ctx* ctx, uint32_t* a, uint32_t* b, int64_t n, uint32_t max_a, uint32_t max_b)
{
    cudaStream_t s = ctx->stream;
    DevBuf<uint32_t> t1, t2;
    MML_TRY(t1.alloc(n)); MML_TRY(t2.alloc(n));
    MML_TRY(radix_sort_pairs(b, a, t1.p, t2.p, n, bits_for(max_b), s));   // by b, carrying a
    MML_TRY(radix_sort_pairs(a, b, t1.p, t2.p, n, bits_for(max_a), s));   // by a (stable), carrying b
    return MML_OK;
}

static int32_t build_feedback(Feedback& f, const int32_t* h_users, const int32_t* h_items, int64_t n)
{
    cudaStream_t s = f.ctx->stream;
    DevBuf<uint32_t> u, i, head, pos, uu, ii, cnt;
    MML_TRY(u.alloc(n)); MML_TRY(i.alloc(n)); MML_TRY(head.alloc(n)); MML_TRY(pos.alloc((size_t)n + 1));
    if (n > 0) {
        MML_CUDA(cudaMemcpyAsync(u.p, h_users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaMemcpyAsync(i.p, h_items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    }
    MML_TRY(sort_pairs_lex(f.ctx, u.p, i.p, n, (uint32_t)std::max(f.max_user, 0), (uint32_t)std::max(f.max_item, 0)));
    pair_head_kernel<<<grid_n(n), 256, 0, s>>>(u.p, i.p, n, head.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(exclusive_scan_u32(head.p, pos.p, n, s));
    uint32_t nnz = 0;
    MML_CUDA(cudaMemcpyAsync(&nnz, pos.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    f.nnz = nnz;
    MML_TRY(uu.alloc(nnz)); MML_TRY(ii.alloc(nnz));
    pair_compact_kernel<<<grid_n(n), 256, 0, s>>>(head.p, pos.p, u.p, i.p, n, uu.p, ii.p);
    MML_CUDA(cudaGetLastError());
    // CSR by user: pairs are sorted by (user, item)
    MML_TRY(cnt.alloc(std::max(f.n_users(), f.n_items())));
    MML_TRY(f.user_ptr.alloc((size_t)f.n_users() + 1)); MML_TRY(f.item_ptr.alloc((size_t)f.n_items() + 1));
    MML_TRY(f.user_cols.alloc(nnz)); MML_TRY(f.item_rows.alloc(nnz));
    MML_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
    MML_TRY(histogram_i32((const int32_t*)uu.p, nnz, cnt.p, s));
    MML_TRY(exclusive_scan_u32(cnt.p, f.user_ptr.p, f.n_users(), s));
    MML_CUDA(cudaMemcpyAsync(f.user_cols.p, ii.p, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToDevice, s));
    // CSC: the same pairs sorted by (item, user)
    MML_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
    MML_TRY(histogram_i32((const int32_t*)ii.p, nnz, cnt.p, s));
    MML_TRY(exclusive_scan_u32(cnt.p, f.item_ptr.p, f.n_items(), s));
    MML_TRY(sort_pairs_lex(f.ctx, ii.p, uu.p, nnz, (uint32_t)std::max(f.max_item, 0), (uint32_t)std::max(f.max_user, 0)));
    MML_CUDA(cudaMemcpyAsync(f.item_rows.p, uu.p, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToDevice, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

// rows in descending nnz order
static int32_t rows_by_nnz(Ctx* ctx, const uint32_t* ptr, int32_t n_rows, DevBuf<int32_t>& out)
{
    cudaStream_t s = ctx->stream;
    DevBuf<uint32_t> key, val, t1, t2;
    MML_TRY(key.alloc(n_rows)); MML_TRY(val.alloc(n_rows)); MML_TRY(t1.alloc(n_rows)); MML_TRY(t2.alloc(n_rows));
    MML_TRY(out.alloc(n_rows));
    nnz_key_kernel<<<grid_n(n_rows), 256, 0, s>>>(ptr, n_rows, key.p, val.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(radix_sort_pairs(key.p, val.p, t1.p, t2.p, n_rows, 32, s));
    MML_CUDA(cudaMemcpyAsync(out.p, val.p, sizeof(int32_t) * (size_t)n_rows, cudaMemcpyDeviceToDevice, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

// Multi-GPU: rows of a half-sweep are independent given the other factor matrix, so rank r solves the
// contiguous row range [range[r], range[r + 1]) and the ranks all-gather the solved rows. Ranges are balanced by
// cost = nnz + row_cost (row_cost stands for the solve, which every row pays, empty ones included: WRMF.cs:79-92
// solves all rows). out = the rank's rows in descending nnz order.
static int32_t shard_rows(Ctx* ctx, const uint32_t* d_ptr, int32_t n_rows, int64_t row_cost, std::vector<int32_t>& range,
                          DevBuf<int32_t>& out, int32_t* n_local)
{
    cudaStream_t s = ctx->stream;
    const int R = std::max(ctx->n_gpus, 1);
    std::vector<uint32_t> ptr((size_t)n_rows + 1);
    MML_CUDA(cudaMemcpyAsync(ptr.data(), d_ptr, sizeof(uint32_t) * ptr.size(), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    const int64_t total = (int64_t)ptr[n_rows] + row_cost * n_rows;
    range.assign((size_t)R + 1, n_rows);
    range[0] = 0;
    int64_t cum = 0;
    int r = 1;
    for (int32_t row = 0; row < n_rows && r < R; row++) {
        cum += (int64_t)(ptr[row + 1] - ptr[row]) + row_cost;
        while (r < R && cum * R >= total * r) range[r++] = row + 1;
    }
    const int32_t lo = range[ctx->rank], hi = range[ctx->rank + 1];
    std::vector<int32_t> rows((size_t)(hi - lo));
    std::iota(rows.begin(), rows.end(), lo);
    std::stable_sort(rows.begin(), rows.end(), [&](int32_t a, int32_t b) { return ptr[a + 1] - ptr[a] > ptr[b + 1] - ptr[b]; });
    MML_TRY(out.alloc(std::max<size_t>(rows.size(), 1)));
    if (!rows.empty()) MML_CUDA(cudaMemcpyAsync(out.p, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaStreamSynchronize(s));
    *n_local = hi - lo;
    return MML_OK;
}

// ---- Gram matrix: HH = sum_i (float)(h_i[r] * h_i[c]) in double (WRMF.cs:94-108) -------------------
constexpr int AT = 512;                 // threads of the Gram / ALS kernels: 16 row lanes x 32 column lanes
constexpr int GRAM_ROWS = 1024;         // rows per partial
constexpr int HB = 8;                   // factor rows staged per step

template <int RA, int CB>               // RA = ceil(k / 16), CB = ceil(k / 32)
__global__ void __launch_bounds__(AT) gram_partial_kernel(const float* __restrict__ H, int32_t n_rows, int32_t k,
                                                          double* __restrict__ part)
{
    extern __shared__ float sh[];       // [HB][k]
    const int tr = threadIdx.x / 32, tc = threadIdx.x % 32;
    double acc[RA][CB];
#pragma unroll
    for (int a = 0; a < RA; a++)
#pragma unroll
        for (int b = 0; b < CB; b++) acc[a][b] = 0.0;
    const int32_t lo = blockIdx.x * GRAM_ROWS, hi = min(lo + GRAM_ROWS, n_rows);
    for (int32_t base = lo; base < hi; base += HB) {
        const int nb = min(HB, hi - base);
        for (int t = threadIdx.x; t < nb * k; t += AT) sh[t] = H[(size_t)base * k + t];
        __syncthreads();
        for (int x = 0; x < nb; x++) {
            const float* h = sh + x * k;
            float hr[RA], hc[CB];
#pragma unroll
            for (int a = 0; a < RA; a++) { const int r = tr + 16 * a; hr[a] = r < k ? h[r] : 0.f; }
#pragma unroll
            for (int b = 0; b < CB; b++) { const int c = tc + 32 * b; hc[b] = c < k ? h[c] : 0.f; }
#pragma unroll
            for (int a = 0; a < RA; a++)
#pragma unroll
                for (int b = 0; b < CB; b++) acc[a][b] += (double)__fmul_rn(hr[a], hc[b]);
        }
        __syncthreads();
    }
    double* out = part + (size_t)blockIdx.x * k * k;
#pragma unroll
    for (int a = 0; a < RA; a++)
#pragma unroll
        for (int b = 0; b < CB; b++) {
            const int r = tr + 16 * a, c = tc + 32 * b;
            if (r < k && c < k) out[(size_t)r * k + c] = acc[a][b];
        }
}

__global__ void gram_reduce_kernel(const double* __restrict__ part, int n_part, int32_t kk, double* __restrict__ HH)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kk) return;
    double s = 0;
    for (int p = 0; p < n_part; p++) s += part[(size_t)p * kk + idx];   // fixed order: deterministic
    HH[idx] = s;
}

// ---- one ALS row per CTA (WRMF.cs:110-156) ----------------------------------------------------------
template <int RA, int CB>
__global__ void __launch_bounds__(AT) als_rows_kernel(const uint32_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                                      const int32_t* __restrict__ order, int32_t n_rows,
                                                      float* __restrict__ W, const float* __restrict__ H, int32_t k,
                                                      const double* __restrict__ HH, double alpha, double reg,
                                                      unsigned* __restrict__ counter)
{
    extern __shared__ double smem_d[];
    double* A = smem_d;                          // [k][k] lower triangle used
    double* bvec = A + (size_t)k * k;            // [k]
    float* sh = reinterpret_cast<float*>(bvec + k);   // [HB][k]
    __shared__ int s_row;
    const int tr = threadIdx.x / 32, tc = threadIdx.x % 32;
    for (;;) {
        if (threadIdx.x == 0) s_row = (int)atomicAdd(counter, 1u);
        __syncthreads();
        const int q = s_row;
        if (q >= n_rows) break;
        const int u = order[q];
        const uint32_t beg = row_ptr[u], end = row_ptr[u + 1];
        if (beg == end) {   // HCp = 0 => w = 0 exactly
            for (int f = threadIdx.x; f < k; f += AT) W[(size_t)u * k + f] = 0.f;
            __syncthreads();
            continue;
        }
        double acc[RA][CB];
#pragma unroll
        for (int a = 0; a < RA; a++)
#pragma unroll
            for (int b = 0; b < CB; b++) acc[a][b] = 0.0;
        double bsum = 0.0;
        for (uint32_t base = beg; base < end; base += HB) {
            const int nb = (int)min((uint32_t)HB, end - base);
            for (int t = threadIdx.x; t < nb * k; t += AT) {
                const int x = t / k, f = t - x * k;
                sh[t] = H[(size_t)cols[base + x] * k + f];
            }
            __syncthreads();
            for (int x = 0; x < nb; x++) {
                const float* h = sh + x * k;
                float hr[RA], hc[CB];
#pragma unroll
                for (int a = 0; a < RA; a++) { const int r = tr + 16 * a; hr[a] = r < k ? h[r] : 0.f; }
#pragma unroll
                for (int b = 0; b < CB; b++) { const int c = tc + 32 * b; hc[b] = c < k ? h[c] : 0.f; }
#pragma unroll
                for (int a = 0; a < RA; a++)
#pragma unroll
                    for (int b = 0; b < CB; b++) acc[a][b] += (double)__fmul_rn(hr[a], hc[b]);
                if (threadIdx.x < k) bsum += (double)h[threadIdx.x];
            }
            __syncthreads();
        }
        // m = HH + alpha * acc + reg * I ; HCp = (1 + alpha) * sum h
#pragma unroll
        for (int a = 0; a < RA; a++)
#pragma unroll
            for (int b = 0; b < CB; b++) {
                const int r = tr + 16 * a, c = tc + 32 * b;
                if (r < k && c < k) A[(size_t)r * k + c] = HH[(size_t)r * k + c] + acc[a][b] * alpha + (r == c ? reg : 0.0);
            }
        if (threadIdx.x < k) bvec[threadIdx.x] = bsum * (1.0 + alpha);
        __syncthreads();
        // Cholesky A = L L^T (lower, in place), right-looking
        for (int j = 0; j < k; j++) {
            if (threadIdx.x == 0) A[(size_t)j * k + j] = sqrt(A[(size_t)j * k + j]);
            __syncthreads();
            const double d = A[(size_t)j * k + j];
            for (int i = j + 1 + threadIdx.x; i < k; i += AT) A[(size_t)i * k + j] /= d;
            __syncthreads();
            // trailing update of the lower triangle: A[i][c] -= L[i][j] * L[c][j], j < c <= i
            const int m = k - j - 1;
            for (int t = threadIdx.x; t < m * m; t += AT) {
                const int i = j + 1 + t / m, c = j + 1 + t % m;
                if (c <= i) A[(size_t)i * k + c] -= A[(size_t)i * k + j] * A[(size_t)c * k + j];
            }
            __syncthreads();
        }
        // forward substitution L y = b, then L^T w = y (column oriented)
        for (int j = 0; j < k; j++) {
            if (threadIdx.x == 0) bvec[j] /= A[(size_t)j * k + j];
            __syncthreads();
            const double y = bvec[j];
            for (int i = j + 1 + threadIdx.x; i < k; i += AT) bvec[i] -= A[(size_t)i * k + j] * y;
            __syncthreads();
        }
        for (int j = k - 1; j >= 0; j--) {
            if (threadIdx.x == 0) bvec[j] /= A[(size_t)j * k + j];
            __syncthreads();
            const double w = bvec[j];
            for (int i = threadIdx.x; i < j; i += AT) bvec[i] -= A[(size_t)j * k + i] * w;
            __syncthreads();
        }
        for (int f = threadIdx.x; f < k; f += AT) W[(size_t)u * k + f] = (float)bvec[f];
        __syncthreads();
    }
}

struct WrmfTcWork;
WrmfTcWork* wrmf_tc_work_create();
void wrmf_tc_work_destroy(WrmfTcWork* w);
bool wrmf_tc_eligible(int32_t k);
int32_t wrmf_tc_half_sweep(Ctx* ctx, WrmfTcWork* work, const uint32_t* row_ptr, const int32_t* cols, const int32_t* order, int32_t n_rows,
                           float* W, const float* H, int32_t k, const double* HH, double alpha, double reg, int64_t* launches,
                           float* debug_G_row0, int first_solver);
static int g_wrmf_mode = 0;            // MML_WRMF_AUTO
static float* g_wrmf_debug_G = nullptr;

typedef void (*gram_fn_t)(const float*, int32_t, int32_t, double*);
typedef void (*als_fn_t)(const uint32_t*, const int32_t*, const int32_t*, int32_t, float*, const float*, int32_t,
                         const double*, double, double, unsigned*);

static int32_t pick(int32_t k, gram_fn_t* g, als_fn_t* a)
{
    if (k <= 32) { *g = gram_partial_kernel<2, 1>; *a = als_rows_kernel<2, 1>; }
    else if (k <= 64) { *g = gram_partial_kernel<4, 2>; *a = als_rows_kernel<4, 2>; }
    else if (k <= 96) { *g = gram_partial_kernel<6, 3>; *a = als_rows_kernel<6, 3>; }
    else if (k <= 128) { *g = gram_partial_kernel<8, 4>; *a = als_rows_kernel<8, 4>; }
    else if (k <= 160) { *g = gram_partial_kernel<10, 5>; *a = als_rows_kernel<10, 5>; }
    else { set_error("WRMF: num_factors=%d > 160 is not supported", k); return MML_ERR_UNSUPPORTED; }
    return MML_OK;
}

// W <- optimum given H (one half-sweep, WRMF.cs:79-92)
static int32_t half_sweep(Wrmf& m, const uint32_t* row_ptr, const int32_t* cols, const int32_t* order, int32_t n_rows,
                          float* W, const float* H, int32_t n_h_rows)
{
    cudaStream_t s = m.ctx->stream;
    const int32_t k = m.k;
    gram_fn_t gf; als_fn_t af;
    MML_TRY(pick(k, &gf, &af));
    const int n_part = (int)std::max<int64_t>(ceil_div(n_h_rows, GRAM_ROWS), 1);
    if (m.HH_part.n < (size_t)n_part * k * k) MML_TRY(m.HH_part.alloc((size_t)n_part * k * k));
    gf<<<n_part, AT, sizeof(float) * HB * k, s>>>(H, n_h_rows, k, m.HH_part.p);
    gram_reduce_kernel<<<(unsigned)ceil_div(k * k, 256), 256, 0, s>>>(m.HH_part.p, n_part, k * k, m.HH.p);
    MML_CUDA(cudaGetLastError());
    // AUTO takes the tensor path only when HH sums many more rows than it has columns: with n_h_rows ~ k the system is so
    // ill-conditioned that the fp32-level rounding of the per-row Gram sums shows above the 1e-4 gate (double does not).
    const bool forced = g_wrmf_mode == MML_WRMF_TENSOR || g_wrmf_mode == MML_WRMF_TENSOR_F64 || g_wrmf_mode == MML_WRMF_TENSOR_PCG;
    const bool tc = g_wrmf_mode != MML_WRMF_FP64 && wrmf_tc_eligible(k) && (forced || (int64_t)n_h_rows >= 16ll * k);
    MML_CHECK(tc || !forced, MML_ERR_UNSUPPORTED, "WRMF: num_factors=%d is outside the tensor-core path (multiple of 4, <= 128)", k);
    if (tc) {
        m.launches += 2;
        if (!m.tc_work) m.tc_work = wrmf_tc_work_create();
        return wrmf_tc_half_sweep(m.ctx, m.tc_work, row_ptr, cols, order, n_rows, W, H, k, m.HH.p, m.alpha, m.reg, &m.launches, g_wrmf_debug_G,
                                  g_wrmf_mode == MML_WRMF_TENSOR_F64 ? 2 : (g_wrmf_mode == MML_WRMF_TENSOR_PCG ? 0 : 1));
    }
    const size_t smem = sizeof(double) * ((size_t)k * k + k) + sizeof(float) * HB * k;
    MML_CUDA(cudaFuncSetAttribute((const void*)af, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MML_CUDA(cudaMemsetAsync(m.counter.p, 0, sizeof(unsigned), s));
    const int grid = std::min(n_rows, m.ctx->sm_count * (smem <= 100 * 1024 ? 2 : 1));
    af<<<std::max(grid, 1), AT, smem, s>>>(row_ptr, cols, order, n_rows, W, H, k, m.HH.p, m.alpha, m.reg, m.counter.p);
    MML_CUDA(cudaGetLastError());
    m.launches += 3;
    return MML_OK;
}

Feedback* feedback_of(mml_feedback* h);
int32_t items_eval_device(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                          const int32_t* users, int64_t n_users, const int32_t* candidates, int64_t n_cand,
                          const int64_t* test_ptr, const int32_t* test_idx,
                          const int64_t* ignore_ptr, const int32_t* ignore_idx, int32_t n,
                          float* out_measures, int32_t* out_used, int64_t* launches);
int32_t topn_device(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                    const int32_t* users, int64_t n_users, int32_t n, const int32_t* candidates, int64_t n_cand,
                    const int64_t* ignore_ptr, const int32_t* ignore_idx,
                    int32_t* out_items, float* out_scores, int32_t* out_counts, int64_t* launches);

}  // namespace mml

using namespace mml;

// On a one-process multi-GPU context both handles are roots over one ordinary handle per GPU (shards[r] on rank r's
// context): the feedback and the model are replicated, every half-sweep solves its own row range on every GPU and the rows
// are all-gathered (as in the one-process-per-GPU mode); Recommend() / Evaluate() split their user lists over the GPUs.
struct mml_feedback { Feedback f; std::vector<mml_feedback*> shards; };
struct mml_wrmf { Wrmf m; std::vector<mml_wrmf*> shards; };
namespace mml { Feedback* feedback_of(mml_feedback* h) { return h ? &h->f : nullptr; } }

extern "C" int32_t mml_feedback_create(mml_ctx* hctx, const int32_t* users, const int32_t* items, int64_t n,
                                       int32_t max_user, int32_t max_item, mml_feedback** out)
{
    MML_LOCK(mml::ctx_of(hctx));
    MML_CHECK(hctx && out && (n == 0 || (users && items)), MML_ERR_ARG, "mml_feedback_create: NULL argument");
    MML_CHECK(n >= 0 && n < ((int64_t)1 << 31), MML_ERR_ARG, "mml_feedback_create: n out of range");
    for (int64_t t = 0; t < n; t++)
        MML_CHECK((uint32_t)users[t] <= (uint32_t)max_user && (uint32_t)items[t] <= (uint32_t)max_item, MML_ERR_ARG,
                  "mml_feedback_create: event %lld has id out of range (user %d, item %d)", (long long)t, users[t], items[t]);
    Ctx* ctx = ctx_of(hctx);
    if (ctx->is_root()) {
        mml_feedback* root = new (std::nothrow) mml_feedback();
        MML_CHECK(root != nullptr, MML_ERR_ARG, "out of host memory");
        root->f.ctx = ctx; root->f.max_user = max_user; root->f.max_item = max_item; root->f.n_events = n;
        root->shards.assign(ctx->peers.size(), nullptr);
        const int32_t st = on_ranks((int)ctx->peers.size(), [&](int x) -> int32_t {
            return mml_feedback_create(ctx->peers[(size_t)x], users, items, n, max_user, max_item, &root->shards[(size_t)x]);
        });
        if (st != MML_OK) { mml_feedback_destroy(root); return st; }
        root->f.nnz = root->shards[0]->f.nnz;
        *out = root;
        return MML_OK;
    }
    MML_CUDA(cudaSetDevice(ctx->device));
    mml_feedback* h = new (std::nothrow) mml_feedback();
    MML_CHECK(h != nullptr, MML_ERR_ARG, "out of host memory");
    h->f.ctx = ctx; h->f.max_user = max_user; h->f.max_item = max_item; h->f.n_events = n;
    const int32_t st = build_feedback(h->f, users, items, n);
    if (st) { delete h; return st; }
    *out = h;
    return MML_OK;
}

extern "C" int32_t mml_feedback_destroy(mml_feedback* f)
{
    MML_LOCK((f ? f->f.ctx : nullptr));
    if (!f) return MML_OK;
    if (f->f.ctx->is_root()) {
        for (mml_feedback* s : f->shards) mml_feedback_destroy(s);
        delete f;
        return MML_OK;
    }
    cudaSetDevice(f->f.ctx->device);
    delete f;
    return MML_OK;
}

extern "C" int32_t mml_feedback_nnz(mml_feedback* f, int64_t* nnz)
{
    MML_LOCK((f ? f->f.ctx : nullptr));
    MML_CHECK(f && nnz, MML_ERR_ARG, "NULL argument");
    *nnz = f->f.nnz;
    return MML_OK;
}

extern "C" int32_t mml_feedback_csr(mml_feedback* h, int32_t by_item, int64_t* row_ptr, int32_t* cols)
{
    MML_LOCK((h ? h->f.ctx : nullptr));
    MML_CHECK(h && row_ptr && cols, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) return mml_feedback_csr(h->shards[0], by_item, row_ptr, cols);
    Feedback& f = h->f;
    MML_CUDA(cudaSetDevice(f.ctx->device));
    cudaStream_t s = f.ctx->stream;
    const int32_t rows = by_item ? f.n_items() : f.n_users();
    std::vector<uint32_t> p((size_t)rows + 1);
    MML_CUDA(cudaMemcpyAsync(p.data(), (by_item ? f.item_ptr : f.user_ptr).p, sizeof(uint32_t) * p.size(), cudaMemcpyDeviceToHost, s));
    if (f.nnz > 0)
        MML_CUDA(cudaMemcpyAsync(cols, (by_item ? f.item_rows : f.user_cols).p, sizeof(int32_t) * (size_t)f.nnz, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    for (size_t t = 0; t < p.size(); t++) row_ptr[t] = p[t];
    return MML_OK;
}

extern "C" int32_t mml_wrmf_create(mml_ctx* hctx, mml_feedback* hf, const mml_wrmf_params* p, mml_wrmf** out)
{
    MML_LOCK(mml::ctx_of(hctx));
    MML_CHECK(hctx && hf && p && out, MML_ERR_ARG, "mml_wrmf_create: NULL argument");
    MML_CHECK(p->num_factors >= 1 && p->num_factors <= 160, MML_ERR_UNSUPPORTED, "mml_wrmf_create: num_factors=%d not in [1,160]", p->num_factors);
    Ctx* ctx = ctx_of(hctx);
    if (ctx->is_root()) {
        MML_CHECK(hf->shards.size() == ctx->peers.size(), MML_ERR_ARG, "mml_wrmf_create: the feedback was not created on this multi-GPU context");
        mml_wrmf* root = new (std::nothrow) mml_wrmf();
        MML_CHECK(root != nullptr, MML_ERR_ARG, "out of host memory");
        root->m.ctx = ctx; root->m.fb = &hf->f; root->m.k = p->num_factors; root->m.alpha = p->alpha; root->m.reg = p->regularization;
        root->shards.assign(ctx->peers.size(), nullptr);
        const int32_t st = on_ranks((int)ctx->peers.size(), [&](int x) -> int32_t {
            return mml_wrmf_create(ctx->peers[(size_t)x], hf->shards[(size_t)x], p, &root->shards[(size_t)x]);
        });
        if (st != MML_OK) { mml_wrmf_destroy(root); return st; }
        *out = root;
        return MML_OK;
    }
    MML_CUDA(cudaSetDevice(ctx->device));
    mml_wrmf* h = new (std::nothrow) mml_wrmf();
    MML_CHECK(h != nullptr, MML_ERR_ARG, "out of host memory");
    Wrmf& m = h->m;
    m.ctx = ctx; m.fb = &hf->f; m.k = p->num_factors; m.alpha = p->alpha; m.reg = p->regularization;
    int32_t st = MML_OK;
    do {
        if ((st = m.U.alloc((size_t)m.fb->n_users() * m.k)) || (st = m.V.alloc((size_t)m.fb->n_items() * m.k))) break;
        if ((st = m.HH.alloc((size_t)m.k * m.k)) || (st = m.counter.alloc(1))) break;
        if (ctx->n_gpus > 1) {
            const int64_t row_cost = 8 * (int64_t)m.k;   // measured at k = 128: one solve costs about as much as 1100 Gram-sum entries
            if ((st = shard_rows(ctx, m.fb->user_ptr.p, m.fb->n_users(), row_cost, m.range_u, m.order_u, &m.n_local_u))) break;
            if ((st = shard_rows(ctx, m.fb->item_ptr.p, m.fb->n_items(), row_cost, m.range_i, m.order_i, &m.n_local_i))) break;
        } else {
            if ((st = rows_by_nnz(ctx, m.fb->user_ptr.p, m.fb->n_users(), m.order_u))) break;
            if ((st = rows_by_nnz(ctx, m.fb->item_ptr.p, m.fb->n_items(), m.order_i))) break;
            m.n_local_u = m.fb->n_users(); m.n_local_i = m.fb->n_items();
            m.range_u = {0, m.n_local_u}; m.range_i = {0, m.n_local_i};
        }
        if (cudaEventCreate(&m.ev0) != cudaSuccess || cudaEventCreate(&m.ev1) != cudaSuccess) { set_error("cudaEventCreate failed"); st = MML_ERR_CUDA; break; }
    } while (0);
    if (st) { delete h; return st; }
    *out = h;
    return MML_OK;
}

extern "C" int32_t mml_wrmf_destroy(mml_wrmf* h)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    if (!h) return MML_OK;
    if (h->m.ctx->is_root()) {
        for (mml_wrmf* s : h->shards) mml_wrmf_destroy(s);
        delete h;
        return MML_OK;
    }
    cudaSetDevice(h->m.ctx->device);
    cudaStreamSynchronize(h->m.ctx->stream);
    if (h->m.ev0) cudaEventDestroy(h->m.ev0);
    if (h->m.ev1) cudaEventDestroy(h->m.ev1);
    if (h->m.tc_work) wrmf_tc_work_destroy(h->m.tc_work);
    delete h;
    return MML_OK;
}

extern "C" int32_t mml_wrmf_set_model(mml_wrmf* h, const float* user_factors, const float* item_factors)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && user_factors && item_factors, MML_ERR_ARG, "mml_wrmf_set_model: NULL argument");
    MML_FORWARD_ALL(h, mml_wrmf_set_model(s, user_factors, item_factors));
    Wrmf& m = h->m;
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    MML_CUDA(cudaMemcpyAsync(m.U.p, user_factors, sizeof(float) * (size_t)m.fb->n_users() * m.k, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(m.V.p, item_factors, sizeof(float) * (size_t)m.fb->n_items() * m.k, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaStreamSynchronize(s));
    m.has_model = true;
    return MML_OK;
}

__global__ void wrmf_init_kernel(float* __restrict__ rows, int64_t n, uint64_t seed, uint64_t stream, float mean, float stddev)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        uint64_t x = seed ^ (stream << 56);
        x += 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1);
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
        const float u1 = ((float)(uint32_t)(x >> 40) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(uint32_t)((x >> 8) & 0xFFFFFFu) + 0.5f) * (1.0f / 16777216.0f);
        rows[t] = mean + stddev * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    }
}

extern "C" int32_t mml_wrmf_init_model(mml_wrmf* h, uint64_t seed, double init_mean, double init_stddev)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    MML_FORWARD_ALL(h, mml_wrmf_init_model(s, seed, init_mean, init_stddev));   // counter-based: the same model on every GPU
    Wrmf& m = h->m;
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    const int64_t nu = (int64_t)m.fb->n_users() * m.k, ni = (int64_t)m.fb->n_items() * m.k;
    wrmf_init_kernel<<<grid_n(nu), 256, 0, s>>>(m.U.p, nu, seed, 1, (float)init_mean, (float)init_stddev);
    wrmf_init_kernel<<<grid_n(ni), 256, 0, s>>>(m.V.p, ni, seed, 2, (float)init_mean, (float)init_stddev);
    MML_CUDA(cudaGetLastError());
    MML_CUDA(cudaStreamSynchronize(s));
    m.launches += 2;
    m.has_model = true;
    return MML_OK;
}

extern "C" int32_t mml_wrmf_get_model(mml_wrmf* h, float* user_factors, float* item_factors)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) return mml_wrmf_get_model(h->shards[0], user_factors, item_factors);   // replicated
    Wrmf& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_wrmf_get_model: no model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    if (user_factors) MML_CUDA(cudaMemcpyAsync(user_factors, m.U.p, sizeof(float) * (size_t)m.fb->n_users() * m.k, cudaMemcpyDeviceToHost, s));
    if (item_factors) MML_CUDA(cudaMemcpyAsync(item_factors, m.V.p, sizeof(float) * (size_t)m.fb->n_items() * m.k, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

// All-gather of the solved rows: every rank's range is broadcast from it (ranges differ in length), one grouped call.
static int32_t gather_rows(Wrmf& m, float* W, const std::vector<int32_t>& range)
{
    if (m.ctx->n_gpus <= 1) return MML_OK;
    MML_TRY(dist_group_start());
    for (int r = 0; r < m.ctx->n_gpus; r++)
        MML_TRY(dist_broadcast_f32(m.ctx, W + (size_t)range[r] * m.k, (size_t)(range[r + 1] - range[r]) * m.k, r));
    MML_TRY(dist_group_end());
    return MML_OK;
}

extern "C" int32_t mml_wrmf_shard(mml_wrmf* h, int32_t by_item, int32_t* ranges)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && ranges, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) return mml_wrmf_shard(h->shards[0], by_item, ranges);
    const std::vector<int32_t>& r = by_item ? h->m.range_i : h->m.range_u;
    for (size_t t = 0; t < r.size(); t++) ranges[t] = r[t];
    return MML_OK;
}

// WRMF.Iterate (WRMF.cs:68-73): user half-sweep, then item half-sweep
extern "C" int32_t mml_wrmf_iterate(mml_wrmf* h)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    MML_FORWARD_ALL(h, mml_wrmf_iterate(s));
    Wrmf& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_wrmf_iterate: no model (call set_model / init_model first)");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    Feedback& f = *m.fb;
    MML_CUDA(cudaEventRecord(m.ev0, s));
    MML_TRY(half_sweep(m, f.user_ptr.p, f.user_cols.p, m.order_u.p, m.n_local_u, m.U.p, m.V.p, f.n_items()));
    MML_TRY(gather_rows(m, m.U.p, m.range_u));
    MML_TRY(half_sweep(m, f.item_ptr.p, f.item_rows.p, m.order_i.p, m.n_local_i, m.V.p, m.U.p, f.n_users()));
    MML_TRY(gather_rows(m, m.V.p, m.range_i));
    MML_CUDA(cudaEventRecord(m.ev1, s));
    m.timed = true;
    return MML_OK;
}

extern "C" int32_t mml_wrmf_retrain(mml_wrmf* h, int32_t by_item, const int32_t* ids, int64_t n)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && (n == 0 || ids), MML_ERR_ARG, "mml_wrmf_retrain: NULL argument");
    MML_FORWARD_ALL(h, mml_wrmf_retrain(s, by_item, ids, n));   // replicated model: every GPU re-solves the rows itself
    Wrmf& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_wrmf_retrain: no model");
    MML_CHECK(n >= 0 && n < ((int64_t)1 << 31), MML_ERR_ARG, "mml_wrmf_retrain: bad count");
    if (n == 0) return MML_OK;
    Feedback& f = *m.fb;
    const int32_t n_rows = by_item ? f.n_items() : f.n_users();
    for (int64_t t = 0; t < n; t++)
        MML_CHECK(ids[t] >= 0 && ids[t] < n_rows, MML_ERR_ARG, "mml_wrmf_retrain: id %d out of range", ids[t]);
    MML_CUDA(cudaSetDevice(m.ctx->device));
    DevBuf<int32_t> d_ids;
    MML_TRY(d_ids.alloc(n));
    MML_CUDA(cudaMemcpyAsync(d_ids.p, ids, sizeof(int32_t) * n, cudaMemcpyHostToDevice, m.ctx->stream));
    // WRMF.cs:159-170: ComputeSquareMatrix of the other side, then Optimize for the row -- a half-sweep over the given rows
    if (by_item) MML_TRY(half_sweep(m, f.item_ptr.p, f.item_rows.p, d_ids.p, (int32_t)n, m.V.p, m.U.p, f.n_users()));
    else MML_TRY(half_sweep(m, f.user_ptr.p, f.user_cols.p, d_ids.p, (int32_t)n, m.U.p, m.V.p, f.n_items()));
    MML_CUDA(cudaStreamSynchronize(m.ctx->stream));
    return MML_OK;
}

extern "C" int32_t mml_wrmf_set_mode(int32_t mode)
{
    MML_CHECK(mode >= MML_WRMF_AUTO && mode <= MML_WRMF_TENSOR_PCG, MML_ERR_ARG, "mml_wrmf_set_mode: unknown mode %d", mode);
    g_wrmf_mode = mode;
    return MML_OK;
}

// Diagnostic: one user half-sweep on the tensor path; out_gram (128 x 128 floats, row-major, zero beyond k) receives
// sum_{i in S_u} h_i h_i^T of the user with the most feedback events (the first row of the work queue).
extern "C" int32_t mml_wrmf_debug_gram(mml_wrmf* h, float* out_gram, int32_t* out_user)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && out_gram && out_user, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) return mml_wrmf_debug_gram(h->shards[0], out_gram, out_user);
    Wrmf& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_wrmf_debug_gram: no model");
    MML_CHECK(wrmf_tc_eligible(m.k), MML_ERR_UNSUPPORTED, "mml_wrmf_debug_gram: num_factors outside the tensor-core path");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    Feedback& f = *m.fb;
    MML_CHECK(f.n_users() > 0, MML_ERR_STATE, "no users");
    DevBuf<float> scratch;                      // the sweep's result is discarded: the model is left untouched
    MML_TRY(scratch.alloc((size_t)f.n_users() * m.k));
    const int old_mode = g_wrmf_mode;
    g_wrmf_mode = MML_WRMF_TENSOR; g_wrmf_debug_G = out_gram;
    const int32_t st = half_sweep(m, f.user_ptr.p, f.user_cols.p, m.order_u.p, f.n_users(), scratch.p, m.V.p, f.n_items());
    g_wrmf_mode = old_mode; g_wrmf_debug_G = nullptr;
    MML_TRY(st);
    MML_CUDA(cudaMemcpy(out_user, m.order_u.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return MML_OK;
}

extern "C" int32_t mml_wrmf_stats(mml_wrmf* h, int64_t* kernel_launches, float* last_iterate_ms)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) {   // launches of all GPUs, device time of the slowest
        int64_t total = 0; float worst = 0.f;
        for (mml_wrmf* s : h->shards) {
            int64_t l = 0; float ms = 0.f;
            MML_TRY(mml_wrmf_stats(s, &l, &ms));
            total += l; worst = std::max(worst, ms);
        }
        if (kernel_launches) *kernel_launches = total;
        if (last_iterate_ms) *last_iterate_ms = worst;
        return MML_OK;
    }
    Wrmf& m = h->m;
    if (kernel_launches) *kernel_launches = m.launches;
    if (last_iterate_ms) {
        *last_iterate_ms = 0.f;
        if (m.timed) {
            MML_CUDA(cudaSetDevice(m.ctx->device));
            MML_CUDA(cudaEventSynchronize(m.ev1));
            MML_CUDA(cudaEventElapsedTime(last_iterate_ms, m.ev0, m.ev1));
        }
    }
    return MML_OK;
}

// Recommend() on the device-resident WRMF model (Recommender.cs:52-103 with ItemRecommendation/MF.cs:151-157 scores)
extern "C" int32_t mml_wrmf_evaluate(mml_wrmf* h, const int32_t* test_users, int64_t n_test_users,
                                     const int32_t* candidates, int64_t n_cand,
                                     const int64_t* test_ptr, const int32_t* test_idx,
                                     const int64_t* ignore_ptr, const int32_t* ignore_idx, int32_t n,
                                     float* out_measures, int32_t* out_used)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && candidates && test_ptr && (n_test_users == 0 || (test_users && out_measures && out_used)), MML_ERR_ARG,
              "mml_wrmf_evaluate: NULL argument");
    MML_CHECK((n > 0 || n == -1) && n_test_users >= 0, MML_ERR_ARG, "mml_wrmf_evaluate: n must be > 0 or -1");
    MML_CHECK(test_ptr[n_test_users] == 0 || test_idx, MML_ERR_ARG, "mml_wrmf_evaluate: NULL test_idx");
    if (!h->shards.empty()) {   // contiguous ranges of the test users, one per GPU; rows of the outputs are disjoint
        const int64_t N = (int64_t)h->shards.size();
        return on_ranks((int)N, [&](int x) -> int32_t {
            const int64_t lo = n_test_users * x / N, hi = n_test_users * (x + 1) / N;
            if (hi <= lo) return MML_OK;
            std::vector<int64_t> tp((size_t)(hi - lo + 1)), ip;
            for (int64_t t = lo; t <= hi; t++) tp[(size_t)(t - lo)] = test_ptr[t] - test_ptr[lo];
            const bool ign = ignore_ptr && ignore_idx;
            if (ign) { ip.resize((size_t)(hi - lo + 1)); for (int64_t t = lo; t <= hi; t++) ip[(size_t)(t - lo)] = ignore_ptr[t] - ignore_ptr[lo]; }
            return mml_wrmf_evaluate(h->shards[(size_t)x], test_users + lo, hi - lo, candidates, n_cand, tp.data(),
                                     test_idx ? test_idx + test_ptr[lo] : nullptr, ign ? ip.data() : nullptr,
                                     ign ? ignore_idx + ignore_ptr[lo] : nullptr, n, out_measures + lo * 8, out_used + lo);
        });
    }
    Wrmf& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_wrmf_evaluate: no model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    return items_eval_device(m.ctx, m.U.p, m.fb->n_users(), m.V.p, m.fb->n_items(), m.k, test_users, n_test_users,
                             candidates, n_cand, test_ptr, test_idx, ignore_ptr, ignore_idx, n, out_measures, out_used,
                             &m.launches);
}

extern "C" int32_t mml_wrmf_recommend(mml_wrmf* h, const int32_t* users, int64_t n_users, int32_t n,
                                      const int32_t* candidates, int64_t n_cand,
                                      const int64_t* ignore_ptr, const int32_t* ignore_idx,
                                      int32_t* out_items, float* out_scores, int32_t* out_counts)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && (n_users == 0 || (users && out_items && out_scores && out_counts)), MML_ERR_ARG, "mml_wrmf_recommend: NULL argument");
    MML_CHECK(n > 0 || n == -1, MML_ERR_ARG, "mml_wrmf_recommend: n must be > 0 or -1");
    if (!h->shards.empty()) {   // Top-N shards users with no collective: contiguous ranges of the user list, one per GPU
        const int64_t N = (int64_t)h->shards.size();
        const int64_t nc = candidates ? n_cand : h->m.fb->n_items();
        const int64_t n_out = n < 0 ? nc : std::min<int64_t>(n, nc);
        return on_ranks((int)N, [&](int x) -> int32_t {
            const int64_t lo = n_users * x / N, hi = n_users * (x + 1) / N;
            if (hi <= lo) return MML_OK;
            std::vector<int64_t> ip;
            const bool ign = ignore_ptr && ignore_idx;
            if (ign) { ip.resize((size_t)(hi - lo + 1)); for (int64_t t = lo; t <= hi; t++) ip[(size_t)(t - lo)] = ignore_ptr[t] - ignore_ptr[lo]; }
            return mml_wrmf_recommend(h->shards[(size_t)x], users + lo, hi - lo, n, candidates, n_cand, ign ? ip.data() : nullptr,
                                      ign ? ignore_idx + ignore_ptr[lo] : nullptr, out_items + lo * n_out, out_scores + lo * n_out,
                                      out_counts + lo);
        });
    }
    Wrmf& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_wrmf_recommend: no model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    return topn_device(m.ctx, m.U.p, m.fb->n_users(), m.V.p, m.fb->n_items(), m.k, users, n_users, n, candidates, n_cand,
                       ignore_ptr, ignore_idx, out_items, out_scores, out_counts, &m.launches);
}
