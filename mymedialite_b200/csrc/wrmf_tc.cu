// wrmf_tc.cu -- the per-row normal equations of a WRMF half-sweep on the tcgen05 tensor cores (K7 of SURVEY.md 2d).
//
// Reference: WRMF.Optimize(u, ...) (ItemRecommendation/WRMF.cs:110-156):
//     A_u = HH + alpha * sum_{i in S_u} h_i h_i^T + lambda I,   b_u = (1 + alpha) * sum_{i in S_u} h_i,   w_u = A_u^-1 b_u
//
// Two kernels per batch of rows:
//   wrmf_syrk_kernel   G_u = sum_{i in S_u} h_i h_i^T  (k x k) as tcgen05.mma kind::tf32 with BOTH operands taken from one
//                      shared-memory tile of gathered factor rows (MN-major: a gathered row is contiguous along M = N).
//                      fp32 accuracy comes from the split h = hi + lo (hi = upper 19 bits, lo = h - hi, exact):
//                      G = hi hi^T + hi lo^T + lo hi^T  (3 MMAs per 8 rows; lo lo^T ~ 2^-22 relative is dropped).
//                      Accumulators live in TMEM (4 x 128 columns: four rows in flight); producer warps gather rows
//                      with 128-bit loads, split them and store both tiles in the 128-byte-swizzled canonical layout;
//                      epilogue warps read TMEM with tcgen05.ld and write G_u (fp32) to HBM. b_u is accumulated in
//                      double by the producer lanes on the way.
//   wrmf_solve_kernel  A_u = HH (fp64, CUDA cores) + alpha * G_u + lambda I in fp64, blocked Cholesky (16-wide panels,
//                      b_u carried as an extra row so the forward substitution is free), back substitution, w_u -> fp32.
// Only G_u -- the sum over the row's own few hundred entries, a small part of A_u next to HH -- sees fp32 rounding;
// HH, the assembly and the solve stay in the reference's double precision (the gate is 1e-4 relative on w_u).
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

namespace mml {

constexpr int WS_THREADS = 288;          // warp 0: MMA issue + TMEM, warps 1-4: gather producers, warps 5-8: epilogue
constexpr int WS_KCH = 32;               // gathered rows per stage
constexpr int WS_TILE = WS_KCH * 512;    // one tile: 32 rows x 128 floats = 16 KB
constexpr int WS_STAGES = 4;             // stage = hi tile + lo tile = 32 KB
constexpr int WS_KP = 128;

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint32_t bar, uint32_t parity, uint32_t* err)
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
__device__ __forceinline__ bool ws_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void ws_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ws_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ws_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ws_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void ws_ld32(uint32_t taddr, uint32_t (&v)[32])
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
__device__ __forceinline__ void ws_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of one 8-row group of a gathered tile. MN-major 32-bit operands have exactly one legal
// layout, "128-byte swizzle with 32-byte base" (layout type 1): an atom is 4 K-rows x 128 bytes (32 floats along M/N), the
// 32-byte chunks of row r are XOR-ed with r; atoms along M/N are LBO = 512 bytes apart, 4-row groups along K SBO = 2048.
__device__ __forceinline__ uint64_t ws_smem_desc(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(512 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// D = F32, A = B = TF32, A and B MN-major (bits 15, 16), N = 128, M = 128
constexpr uint32_t WS_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct SyrkArgs {
    const uint32_t* row_ptr; const int32_t* cols;      // CSR of the side being solved
    const int32_t* order;                              // rows by descending length
    int32_t q_lo, q_hi;                                // this batch = order[q_lo .. q_hi)
    const float* H; int32_t k;                         // the fixed factor matrix [*, k], k % 4 == 0, k <= 128
    float* G;                                          // [q_hi - q_lo][128][128]
    double* bsum;                                      // [q_hi - q_lo][128], zeroed by the caller
    uint32_t* err;
};

__global__ void __launch_bounds__(WS_THREADS, 1) wrmf_syrk_kernel(const SyrkArgs a)
{
    extern __shared__ uint8_t ws_smem_raw[];
    __shared__ uint64_t bars[2 * WS_STAGES + 8];      // full[4], empty[4], tfull[4], tempty[4]
    __shared__ uint32_t tmem_base_s;
    const uint32_t smem0 = (ws_smem_u32(ws_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = ws_smem_raw + (smem0 - ws_smem_u32(ws_smem_raw));
    const uint32_t bar_full = ws_smem_u32(bars), bar_empty = bar_full + 8 * WS_STAGES;
    const uint32_t bar_tfull = bar_full + 16 * WS_STAGES, bar_tempty = bar_tfull + 8 * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_STAGES; s++) { ws_mbar_init(bar_full + 8 * s, 4); ws_mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 4; b++) { ws_mbar_init(bar_tfull + 8 * b, 1); ws_mbar_init(bar_tempty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(ws_smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ws_fence_before();
    __syncthreads();
    ws_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    // every role walks the same row sequence: batch-local rows blockIdx.x, blockIdx.x + gridDim.x, ...; empty rows are skipped
    if (warp == 0) {
        // ===== MMA issuer: the whole warp walks the rows in uniform control flow, one elected lane issues (inside
        //       `if (lane == 0)` ptxas wraps every tcgen05.mma in an ELECT / R2UR / BRA.U.ANY loop: see topn_tc.cu) =====
        int stage = 0; uint32_t phase = 0; int it = 0;
        for (int q = a.q_lo + blockIdx.x; q < a.q_hi; q += gridDim.x) {
            const int u = a.order[q];
            const uint32_t K = a.row_ptr[u + 1] - a.row_ptr[u];
            if (K == 0) continue;
            const int buf = it & 3;
            ws_mbar_wait(bar_tempty + 8 * buf, ((uint32_t)(it >> 2) & 1u) ^ 1u, a.err);
            ws_fence_after();
            const uint32_t d = tmem_base + (uint32_t)(buf * 128);
            for (uint32_t base = 0; base < K; base += WS_KCH) {
                ws_mbar_wait(bar_full + 8 * stage, phase, a.err);
                ws_fence_after();
                const uint32_t hi = smem0 + (uint32_t)stage * 2 * WS_TILE, lo = hi + WS_TILE;
                const int groups = (int)min((uint32_t)4, (K - base + 7) / 8);
                if (ws_elect_one()) {
                    for (int g = 0; g < groups; g++) {
                        const uint64_t dh = ws_smem_desc(hi + g * 4096), dl = ws_smem_desc(lo + g * 4096);
                        ws_mma_tf32(d, dh, dh, WS_IDESC, (base | (uint32_t)g) != 0u ? 1u : 0u);
                        ws_mma_tf32(d, dh, dl, WS_IDESC, 1u);
                        ws_mma_tf32(d, dl, dh, WS_IDESC, 1u);
                    }
                    ws_commit(bar_empty + 8 * stage);
                    if (base + WS_KCH >= K) ws_commit(bar_tfull + 8 * buf);     // the row's last stage: its accumulators are complete
                }
                __syncwarp();
                if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
            }
            it++;
        }
    } else if (warp <= 4) {
        // ===== gather producers: warp pw takes rows pw, pw + 4, ... of every stage =====
        // (Keeping two stages in flight per warp -- rows of stage t + 1 and item ids of stage t + 2 requested while stage t is
        // split and stored -- was measured: 48.9 against 48.2 ms per epoch at config 3, GPU call ai; a ring of FOUR stages in
        // registers the same, call aj: the stage's fence.proxy.async compiles to MEMBAR.ALL.CTA, which waits for every load the
        // warp has in flight, so a deeper request queue drains at each stage anyway. Writing the tiles with st.async (async
        // proxy, bytes counted on the `full` barrier, no fence; needs a launch with a 1 x 1 x 1 cluster attribute -- without
        // one STAS is an illegal instruction on sm_100a, scripts/probe/st_async_probe.cu) kept the requests in flight but was
        // slower still: 51.8 ms, call am. One warp per WHOLE stage (four stages in flight per CTA, each warp's fence draining
        // only its own loads): 51.9 ms, call an. More requests in flight do not help this kernel.)
        const int pw = warp - 1;
        const int j = lane >> 3, c = lane & 7;           // 128-byte block along M/N and 16-byte chunk inside it
        const bool col_ok = 4 * lane < a.k;
        int stage = 0; uint32_t phase = 0;
        int qi = 0;
        for (int q = a.q_lo + blockIdx.x; q < a.q_hi; q += gridDim.x, qi++) {
            const int u = a.order[q];
            const uint32_t beg = a.row_ptr[u], end = a.row_ptr[u + 1];
            if (beg == end) continue;
            double b0 = 0, b1 = 0, b2 = 0, b3 = 0;
            for (uint32_t base = beg; base < end; base += WS_KCH) {
                // ids of this warp's 8 rows, then all 8 row loads in flight
                int32_t idv = -1;
                if (lane < 8) { const uint32_t e = base + pw + 4 * lane; if (e < end) idv = a.cols[e]; }
                float4 x[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int32_t id = __shfl_sync(0xffffffffu, idv, i);
                    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (id >= 0 && col_ok) x[i] = __ldg(reinterpret_cast<const float4*>(a.H + (size_t)id * a.k) + lane);
                }
                ws_mbar_wait(bar_empty + 8 * stage, phase ^ 1u, a.err);
                uint8_t* hi = smem_gen + (size_t)stage * 2 * WS_TILE; uint8_t* lo = hi + WS_TILE;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = pw + 4 * i, kg = r >> 2, rr = r & 3;
                    const uint32_t off = (uint32_t)(kg * 2048 + j * 512 + rr * 128 + (((c >> 1) ^ rr) << 5) + ((c & 1) << 4));
                    float4 h4, l4;
                    h4.x = __uint_as_float(__float_as_uint(x[i].x) & 0xFFFFE000u); l4.x = x[i].x - h4.x;
                    h4.y = __uint_as_float(__float_as_uint(x[i].y) & 0xFFFFE000u); l4.y = x[i].y - h4.y;
                    h4.z = __uint_as_float(__float_as_uint(x[i].z) & 0xFFFFE000u); l4.z = x[i].z - h4.z;
                    h4.w = __uint_as_float(__float_as_uint(x[i].w) & 0xFFFFE000u); l4.w = x[i].w - h4.w;
                    *reinterpret_cast<float4*>(hi + off) = h4;
                    *reinterpret_cast<float4*>(lo + off) = l4;
                    b0 += (double)x[i].x; b1 += (double)x[i].y; b2 += (double)x[i].z; b3 += (double)x[i].w;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA's async proxy
                __syncwarp();
                if (lane == 0) ws_mbar_arrive(bar_full + 8 * stage);
                if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
            }
            if (col_ok) {
                double* bs = a.bsum + (size_t)(q - a.q_lo) * WS_KP + 4 * lane;
                atomicAdd(bs, b0); atomicAdd(bs + 1, b1); atomicAdd(bs + 2, b2); atomicAdd(bs + 3, b3);
            }
        }
    } else {
        // ===== epilogue: thread <-> row m of G_u =====
        const int qd = warp & 3;
        const int m = qd * 32 + lane;
        int it = 0;
        for (int q = a.q_lo + blockIdx.x; q < a.q_hi; q += gridDim.x) {
            const int u = a.order[q];
            if (a.row_ptr[u + 1] == a.row_ptr[u]) continue;
            const int buf = it & 3;
            ws_mbar_wait(bar_tfull + 8 * buf, (uint32_t)(it >> 2) & 1u, a.err);
            ws_fence_after();
            float* out = a.G + ((size_t)(q - a.q_lo) * WS_KP + m) * WS_KP;
#pragma unroll 1
            for (int jj = 0; jj < 4; jj++) {
                uint32_t v[32];
                __syncwarp();
                ws_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * 128 + jj * 32), v);
                ws_ld_wait();
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++)
                    reinterpret_cast<float4*>(out + jj * 32)[c4] = make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]),
                                                                               __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
            }
            ws_fence_before();
            __syncwarp();
            if (lane == 0) ws_mbar_arrive(bar_tempty + 8 * buf);
            it++;
        }
    }
    ws_fence_before();
    __syncthreads();
    if (warp == 0) {
        ws_fence_after();
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- assembly + blocked Cholesky solve + iterative refinement, one CTA per row --------------------------------------
// A_u is held as its lower triangle in 16 x 16 blocks (rows padded to 17 elements: conflict-free row and column access),
// block (bi, bj), bj <= bi, at index bi (bi + 1) / 2 + bj.
//   1. A~ = HH + alpha G~ + lambda I with the tensor cores' fp32-accurate G~;   A~ = L L^T (right-looking, 16-wide panels)
//   2. w = L^-T L^-1 b
//   3. refinement: r = b - A w with A applied EXACTLY in double straight from the factor rows
//      (A w = HH w + lambda w + alpha sum_i h_i (h_i . w): 2 k flops per entry instead of k^2), w += L^-T L^-1 r,
//      until |r| <= 1e-9 |b|. The factor is therefore only a preconditioner: its rounding -- and that of G~ -- slows the
//      convergence but does not enter the result, which is the double-precision solution.
// T = float (default): the factor, its block inverses and the triangular solves are single precision -- 39 KB for the
// factor at k = 128, three rows in flight per SM, one shuffle per exchanged value, hardware rsqrt; on WRMF systems
// (cond(A) ~ 10..100, lambda I keeps them far from singular) it converges in the same two residual evaluations as the
// double factor. A row that has not converged after SV_MAX_REFINE<T> steps raises *fail and the host repeats the
// half-sweep with T = double (78 KB, two rows per SM).
constexpr int SV_THREADS = 256;
constexpr int SV_NB = 16;
constexpr int SV_LD = 17;
constexpr int SV_BLK = SV_NB * SV_LD;
template <typename T> struct SvCfg;
template <> struct SvCfg<double> { static constexpr int max_refine = 3; static constexpr int ctas = 2; static constexpr int red_bufs = 8; };
// float: FOUR rows per SM (round 2, late): 64 registers (20 bytes of spill) and 56 KB -- the gather's per-warp partial sums
// share four buffers, warps 4..7 adding to what warps 0..3 wrote.
template <> struct SvCfg<float> { static constexpr int max_refine = 8; static constexpr int ctas = 4; static constexpr int red_bufs = 4; };

struct SolveArgs {
    const uint32_t* row_ptr; const int32_t* cols; const int32_t* order; int32_t q_lo, q_hi;
    const float* G; const double* bsum; const double* HH;   // HH [k][k] fp64
    const float* H;                                         // the fixed factor matrix (refinement gathers it again)
    double alpha, reg;
    int32_t k;
    float* W;                                               // [*, k] the factor matrix being solved
    uint32_t* fail;                                         // rows whose refinement did not converge
};

template <typename T>
__device__ __forceinline__ T* sv_blk(T* L, int bi, int bj) { return L + (size_t)(bi * (bi + 1) / 2 + bj) * SV_BLK; }

__device__ __forceinline__ double sv_rsqrt(double d)
{   // single-precision seed, two Newton steps in double (2^-22 -> 2^-43 -> full precision)
    double rd = (double)rsqrtf((float)d);
    rd = rd * fma(-0.5 * d, rd * rd, 1.5);
    return rd * fma(-0.5 * d, rd * rd, 1.5);
}
__device__ __forceinline__ float sv_rsqrt(float d)
{
    const float rd = rsqrtf(d);
    return rd * fmaf(-0.5f * d, rd * rd, 1.5f);
}

// Triangular solves with the blocked factor. The 16 x 16 diagonal blocks are applied through their explicit inverses
// Dinv[p] = L(p,p)^-1 (computed once per factorisation), so a panel step is a small matrix-vector product instead of a
// 16-step dependent chain: 2 barriers per panel.
// x <- L^-1 x (forward) for the vector x[0 .. kp); all threads of the CTA call it
template <typename T>
__device__ __forceinline__ void sv_forward(T* L, const T* Dinv, T* x, int np)
{
    const int tid = threadIdx.x;
    for (int p = 0; p < np; p++) {
        const int j0 = p * SV_NB;
        T y = 0;
        if (tid < SV_NB) {
            const T* D = Dinv + (size_t)p * SV_BLK + tid * SV_LD;
#pragma unroll
            for (int t = 0; t < SV_NB; t++) y += (t <= tid) ? D[t] * x[j0 + t] : (T)0;
        }
        if (tid < 32) __syncwarp();
        if (tid < SV_NB) x[j0 + tid] = y;
        __syncthreads();
        for (int i = j0 + SV_NB + tid; i < np * SV_NB; i += SV_THREADS) {
            const T* B = sv_blk(L, i / SV_NB, p) + (i % SV_NB) * SV_LD;
            T s = 0;
#pragma unroll
            for (int t = 0; t < SV_NB; t++) s += B[t] * x[j0 + t];
            x[i] -= s;
        }
        __syncthreads();
    }
}

// x <- L^-T x (backward)
template <typename T>
__device__ __forceinline__ void sv_backward(T* L, const T* Dinv, T* x, int np)
{
    const int tid = threadIdx.x;
    for (int p = np - 1; p >= 0; p--) {
        const int j0 = p * SV_NB;
        T y = 0;
        if (tid < SV_NB) {
            const T* D = Dinv + (size_t)p * SV_BLK + tid;
#pragma unroll
            for (int t = 0; t < SV_NB; t++) y += (t >= tid) ? D[t * SV_LD] * x[j0 + t] : (T)0;
        }
        if (tid < 32) __syncwarp();
        if (tid < SV_NB) x[j0 + tid] = y;
        __syncthreads();
        for (int j = tid; j < j0; j += SV_THREADS) {
            const T* B = sv_blk(L, p, j / SV_NB) + (j % SV_NB);
            T s = 0;
#pragma unroll
            for (int t = 0; t < SV_NB; t++) s += B[t * SV_LD] * x[j0 + t];
            x[j] -= s;
        }
        __syncthreads();
    }
}

// shared memory of one row: doubles first (alignment), then the T-typed factor
template <typename T>
__host__ __device__ inline size_t sv_smem_bytes(int np)
{
    const size_t kp = (size_t)np * SV_NB, nblk = (size_t)np * (np + 1) / 2;
    return sizeof(double) * kp * (3 + SvCfg<T>::red_bufs) + sizeof(T) * (nblk * SV_BLK + (size_t)np * SV_BLK + 2 * kp);
}

template <typename T>
__global__ void __launch_bounds__(SV_THREADS, SvCfg<T>::ctas) wrmf_solve_kernel(const SolveArgs a)
{
    extern __shared__ double sv_smem[];
    const int k = a.k, np = (k + SV_NB - 1) / SV_NB, kp = np * SV_NB;
    const int q = a.q_lo + blockIdx.x;
    if (q >= a.q_hi) return;
    const int u = a.order[q];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t beg = a.row_ptr[u], end = a.row_ptr[u + 1];
    if (beg == end) {                                 // HCp = 0 => w = 0 exactly
        for (int f = tid; f < k; f += SV_THREADS) a.W[(size_t)u * k + f] = 0.f;
        return;
    }
    const int nblk = np * (np + 1) / 2;
    double* b0 = sv_smem;                             // [kp] right-hand side
    double* wv = b0 + kp;                             // [kp] solution
    double* rv = wv + kp;                             // [kp] residual
    constexpr int RB = SvCfg<T>::red_bufs;
    double* red = rv + kp;                            // [RB][kp] partial sums of the refinement gather
    T* L = reinterpret_cast<T*>(red + (size_t)RB * kp);   // [nblk][16][17]
    T* Dinv = L + (size_t)nblk * SV_BLK;              // [np][16][17] inverses of the diagonal blocks
    T* rdiag = Dinv + (size_t)np * SV_BLK;            // [kp] 1 / L[i][i]
    T* xs = rdiag + kp;                               // [kp] right-hand side / solution of a triangular solve
    __shared__ double s_norm[2];
    const float* G = a.G + (size_t)blockIdx.x * WS_KP * WS_KP;
    // ---- assemble the lower triangle (in double, rounded to T); padding rows/columns: identity. One 16 x 16 block per
    //      step (256 threads = its 256 entries), block indices advanced incrementally; four blocks' loads in flight.
    {
        const int r = tid >> 4, c = tid & 15;
        int bi = 0, bj = 0;
#pragma unroll 4
        for (int blk = 0; blk < nblk; blk++) {
            const int i = bi * SV_NB + r, j = bj * SV_NB + c;
            double v = 0.0;
            if (i < k && j < k && j <= i) v = a.HH[(size_t)i * k + j] + a.alpha * (double)G[(size_t)i * WS_KP + j] + (i == j ? a.reg : 0.0);
            else if (i == j) v = 1.0;
            L[(size_t)blk * SV_BLK + r * SV_LD + c] = (T)v;
            if (++bj > bi) { bi++; bj = 0; }
        }
    }
    for (int f = tid; f < kp; f += SV_THREADS) {
        const double bb = f < k ? a.bsum[(size_t)blockIdx.x * WS_KP + f] * (1.0 + a.alpha) : 0.0;
        b0[f] = bb; xs[f] = (T)bb;
    }
    __syncthreads();
    // ---- blocked right-looking Cholesky
    for (int p = 0; p < np; p++) {
        const int j0 = p * SV_NB;
        if (warp == 0) {                              // diagonal block in registers: lane l owns row l
            T* D = sv_blk(L, p, p);
            const int row = lane & (SV_NB - 1);       // lanes 16..31 mirror 0..15 (their results are discarded)
            T rr[SV_NB];
#pragma unroll
            for (int c = 0; c < SV_NB; c++) rr[c] = (c <= row) ? D[row * SV_LD + c] : (T)0;
#pragma unroll
            for (int c = 0; c < SV_NB; c++) {
                const T dcc = __shfl_sync(0xffffffffu, rr[c], c);
                const T rd = sv_rsqrt(dcc);           // 1 / L[c][c]
                const T lc = rr[c] * rd;              // L[row][c] for row >= c (row == c: sqrt(dcc))
                rr[c] = lc;
                if (lane == c) rdiag[j0 + c] = rd;
#pragma unroll
                for (int c2 = c + 1; c2 < SV_NB; c2++) {
                    const T l2 = __shfl_sync(0xffffffffu, lc, c2);
                    rr[c2] -= lc * ((row >= c2) ? l2 : (T)0);
                }
            }
            if (lane < SV_NB) {
#pragma unroll
                for (int c = 0; c < SV_NB; c++) if (c <= row) D[row * SV_LD + c] = rr[c];
            }
        }
        __syncthreads();
        if (warp == 0) {
            // explicit inverse of the diagonal block for the triangular solves (not needed by the factorisation itself, so
            // it runs beside the panel): lane c solves L11 x = e_c (column c of L11^-1)
            const T* D = sv_blk(L, p, p);
            T* Di = Dinv + (size_t)p * SV_BLK;
            if (lane < SV_NB) {
                T x[SV_NB];
#pragma unroll
                for (int r = 0; r < SV_NB; r++) {
                    T sacc = (r == lane) ? (T)1 : (T)0;
#pragma unroll
                    for (int t = 0; t < r; t++) sacc -= D[r * SV_LD + t] * x[t];
                    x[r] = (r >= lane) ? sacc * rdiag[j0 + r] : (T)0;
                }
#pragma unroll
                for (int r = 0; r < SV_NB; r++) Di[r * SV_LD + lane] = x[r];
            }
        } else {
            // panel below the diagonal block: X L11^T = A21 by substitution, one thread of warps 1..7 per row
            const T* D = sv_blk(L, p, p);
            for (int i = j0 + SV_NB + (tid - 32); i < kp; i += SV_THREADS - 32) {
                T* Ai = sv_blk(L, i / SV_NB, p) + (i % SV_NB) * SV_LD;
                T x[SV_NB];
#pragma unroll
                for (int c = 0; c < SV_NB; c++) {
                    T sacc = Ai[c];
#pragma unroll
                    for (int t = 0; t < c; t++) sacc -= x[t] * D[c * SV_LD + t];
                    x[c] = sacc * rdiag[j0 + c];
                }
#pragma unroll
                for (int c = 0; c < SV_NB; c++) Ai[c] = x[c];
            }
        }
        __syncthreads();
        // trailing update: C(bi, bj) -= L(bi, p) L(bj, p)^T for p < bj <= bi; 4 x 4 register tiles, 16 threads per block pair
        const int mb = np - p - 1;
        const int npair = mb * (mb + 1) / 2;
        for (int t = tid; t < npair * 16; t += SV_THREADS) {
            const int pair = t >> 4, tx = (t >> 2) & 3, ty = t & 3;
            int xi = (int)((sqrtf(8.f * pair + 1.f) - 1.f) * 0.5f);
            while (xi * (xi + 1) / 2 > pair) xi--;
            while ((xi + 1) * (xi + 2) / 2 <= pair) xi++;
            const int xj = pair - xi * (xi + 1) / 2;
            const int bi = p + 1 + xi, bj = p + 1 + xj;
            const T* Ra = sv_blk(L, bi, p) + (4 * tx) * SV_LD;
            const T* Cb = sv_blk(L, bj, p) + (4 * ty) * SV_LD;
            T acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y] = 0;
#pragma unroll 4
            for (int sidx = 0; sidx < SV_NB; sidx++) {
                T ra[4], cb[4];
#pragma unroll
                for (int x = 0; x < 4; x++) { ra[x] = Ra[x * SV_LD + sidx]; cb[x] = Cb[x * SV_LD + sidx]; }
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y] += ra[x] * cb[y];
            }
            T* C = sv_blk(L, bi, bj) + (4 * tx) * SV_LD + 4 * ty;
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) C[x * SV_LD + y] -= acc[x][y];
        }
        __syncthreads();
    }
    // ---- w = L^-T L^-1 b
    sv_forward(L, Dinv, xs, np);
    sv_backward(L, Dinv, xs, np);
    for (int f = tid; f < kp; f += SV_THREADS) wv[f] = (double)xs[f];
    __syncthreads();
    // ---- refinement against the exact operator
    bool converged = false;
    for (int iter = 0; iter <= SvCfg<T>::max_refine; iter++) {
        // r = b - HH w - lambda w   (HH is symmetric: column reads are coalesced)
        for (int f = tid; f < kp; f += SV_THREADS) {
            double sacc = 0.0;
            if (f < k) {
                sacc = b0[f] - a.reg * wv[f];
                for (int g = 0; g < k; g++) sacc -= a.HH[(size_t)g * k + f] * wv[g];
            }
            rv[f] = sacc;
        }
        // - alpha sum_i h_i (h_i . w): a warp per entry, lane l holds factors 4 l .. 4 l + 3
        double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        const bool col_ok = 4 * lane < k;
        double w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        __syncthreads();
        if (col_ok) { w0 = wv[4 * lane]; w1 = wv[4 * lane + 1]; w2 = wv[4 * lane + 2]; w3 = wv[4 * lane + 3]; }
        for (uint32_t e = beg + warp; e < end; e += SV_THREADS / 32) {
            const int32_t id = a.cols[e];
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col_ok) h = __ldg(reinterpret_cast<const float4*>(a.H + (size_t)id * k) + lane);
            double d = (double)h.x * w0 + (double)h.y * w1 + (double)h.z * w2 + (double)h.w * w3;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            r0 += d * (double)h.x; r1 += d * (double)h.y; r2 += d * (double)h.z; r3 += d * (double)h.w;
        }
        if (col_ok && warp < RB) {
            double* rw = red + (size_t)warp * kp + 4 * lane;
            rw[0] = r0; rw[1] = r1; rw[2] = r2; rw[3] = r3;
        }
        __syncthreads();
        if (RB < SV_THREADS / 32) {                   // the upper warps add to the buffers of the lower ones
            if (col_ok && warp >= RB) {
                double* rw = red + (size_t)(warp - RB) * kp + 4 * lane;
                rw[0] += r0; rw[1] += r1; rw[2] += r2; rw[3] += r3;
            }
            __syncthreads();
        }
        double rmax = 0.0, bmax = 0.0;
        for (int f = tid; f < k; f += SV_THREADS) {
            double sacc = 0.0;
            for (int wi = 0; wi < RB; wi++) sacc += red[(size_t)wi * kp + f];
            const double r = rv[f] - a.alpha * sacc;
            rv[f] = r;
            rmax = fmax(rmax, fabs(r)); bmax = fmax(bmax, fabs(b0[f]));
        }
        // |r|_inf and |b|_inf over the CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o)); bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o)); }
        if (tid == 0) { s_norm[0] = 0.0; s_norm[1] = 0.0; }
        __syncthreads();
        if (lane == 0) {
            atomicMax(reinterpret_cast<unsigned long long*>(&s_norm[0]), (unsigned long long)__double_as_longlong(rmax));
            atomicMax(reinterpret_cast<unsigned long long*>(&s_norm[1]), (unsigned long long)__double_as_longlong(bmax));
        }
        __syncthreads();
        if (!(s_norm[0] > 1e-9 * s_norm[1])) { converged = true; break; }   // uniform across the CTA
        if (iter == SvCfg<T>::max_refine) break;
        for (int f = tid; f < kp; f += SV_THREADS) xs[f] = (T)rv[f];
        __syncthreads();
        sv_forward(L, Dinv, xs, np);
        sv_backward(L, Dinv, xs, np);
        for (int f = tid; f < k; f += SV_THREADS) wv[f] += (double)xs[f];
        __syncthreads();
    }
    if (!converged && tid == 0) atomicAdd(a.fail, 1u);
    for (int f = tid; f < k; f += SV_THREADS) a.W[(size_t)u * k + f] = (float)wv[f];
}

// ---- preconditioned conjugate gradients + one refinement against the exact operator (MML_WRMF_TENSOR_PCG) ----------------
// The Cholesky kernel above spends its time in dependent panel steps behind CTA barriers (ncu, round 1: barrier stall 10.7
// per issue; ~150 barriers and 86 us per row): the 165k rows of config 3 cost 32 ms of a 52 ms epoch. Every system of a
// half-sweep is the SAME matrix A0 = HH + lambda I plus a row-specific alpha G_u that is a small correction for almost every
// row (G_u sums a few hundred of the tens of thousands of rows that make up HH). So M = A0^-1, computed once per half-sweep
// (wrmf_precond_kernel), is a preconditioner under which A_u has its spectrum clustered at 1 + [0, small] and conjugate
// gradients converge in a handful of iterations, each two dense matrix-vector products and four CTA barriers:
//   * thread t keeps row t of A~ = fl32(HH + alpha G~ + lambda I) in 128 registers (read as column t of the symmetric
//     inputs: coalesced); M sits in shared memory (fp32, column access: conflict-free); vectors are exchanged through
//     shared memory and read as broadcast 128-bit loads;
//   * w1 = PCG(A~, b); then ONE residual against the exact operator in double, r = b - HH w1 - lambda w1 - alpha sum_i h_i (h_i.w1)
//     straight from the factor rows (as the Cholesky kernel's refinement), d = PCG(A~, r), w = w1 + d. The fp32 rounding of A~
//     and of the tensor cores' G~ only enters d, i.e. at relative size |d| / |w| ~ 1e-5 of itself;
//   * a row whose inner solves do not converge in CG_MAX_IT iterations, or whose correction is not small (|d| > 1e-3 |w|: the
//     contraction was not what the argument assumes), raises *fail and the half-sweep is redone by the Cholesky kernels.
// Measured at config 3 (profiles/r2_wrmf_solvers.md): rows within 1.2e-7 of the double-precision solve (the Cholesky path:
// ~1e-6), but 40.7 ms of solves per epoch against 32 ms -- ~30 CG iterations x 5 barrier-separated steps per row are as
// latency bound as the ~150 barriers of the blocked factorisation -- so the Cholesky kernel stays the default and this one a
// selectable mode. A plain (Jacobi-preconditioned) CG was measured first and rejected outright: hundreds of iterations
// (58 ms per epoch) and row errors up to 2e-4 without the exact refinement.
constexpr int CG_THREADS = 256;           // two threads per matrix row: thread 2 t + h holds columns 64 h .. 64 h + 63 of row t
constexpr int CG_HALF = WS_KP / 2;
constexpr int CG_MAX_IT = 60;
constexpr int CG_WARPS = CG_THREADS / 32;
constexpr int PRE_THREADS = 128;

// M = (HH + lambda I)^-1 as fp32, one CTA: Cholesky in double in shared memory, then column t of the inverse by thread t
// (forward and backward substitution; the columns live in `scratch`, k x k doubles, [row][t]: coalesced).
__global__ void __launch_bounds__(PRE_THREADS) wrmf_precond_kernel(const double* __restrict__ HH, double reg, int k,
                                                                   double* __restrict__ scratch, float* __restrict__ M)
{
    extern __shared__ double Lp[];                   // [128][129]
    constexpr int LD = WS_KP + 1;
    const int t = threadIdx.x;
    for (int j = 0; j < WS_KP; j++) {
        double v = (j == t) ? 1.0 : 0.0;             // padding: identity
        if (t < k && j < k) v = HH[(size_t)j * k + t] + (j == t ? reg : 0.0);
        Lp[j * LD + t] = v;                          // column t of row j (symmetric)
    }
    __syncthreads();
    for (int j = 0; j < WS_KP; j++) {
        if (t == j) Lp[j * LD + j] = sqrt(Lp[j * LD + j]);
        __syncthreads();
        if (t > j) Lp[t * LD + j] /= Lp[j * LD + j];
        __syncthreads();
        if (t > j) {
            const double ltj = Lp[t * LD + j];
            for (int c = j + 1; c <= t; c++) Lp[t * LD + c] -= ltj * Lp[c * LD + j];
        }
        __syncthreads();
    }
    // column t of the inverse: L y = e_t, L^T x = y
    double* col = scratch + t;                       // element i at col[i * 128]
    for (int i = 0; i < WS_KP; i++) {
        double sacc = (i == t) ? 1.0 : 0.0;
        for (int c = 0; c < i; c++) sacc -= Lp[i * LD + c] * col[(size_t)c * WS_KP];
        col[(size_t)i * WS_KP] = sacc / Lp[i * LD + i];
    }
    for (int i = WS_KP - 1; i >= 0; i--) {
        double sacc = col[(size_t)i * WS_KP];
        for (int c = i + 1; c < WS_KP; c++) sacc -= Lp[c * LD + i] * col[(size_t)c * WS_KP];
        col[(size_t)i * WS_KP] = sacc / Lp[i * LD + i];
    }
    for (int i = 0; i < WS_KP; i++) M[(size_t)i * WS_KP + t] = (float)col[(size_t)i * WS_KP];
}

struct CgShared {
    float vec[2][WS_KP];            // the vector being multiplied (search direction / residual)
    float red[4][2][CG_WARPS];
    double wd[WS_KP];               // w1 in double (exact residual)
    double part[CG_WARPS][WS_KP];   // per-warp partial sums of the exact residual's gather
    double redd[2][CG_WARPS];
};

// sums v0 and v1 over the CTA (callers zero the contribution of the odd thread of a row pair); `red` is one of four rotating
// buffers: a barrier separates write and read, three more barriers pass before the buffer is written again
__device__ __forceinline__ void cg_sum2(float& v0, float& v1, float (*red)[CG_WARPS], int lane, int warp)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
    if (lane == 0) { red[0][warp] = v0; red[1][warp] = v1; }
    __syncthreads();
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int x = 0; x < CG_WARPS; x++) { s0 += red[0][x]; s1 += red[1][x]; }
    v0 = s0; v1 = s1;
}

__device__ __forceinline__ double cg_max(double v, double* red, int lane, int warp)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int x = 1; x < CG_WARPS; x++) m = fmax(m, red[x]);
    return m;
}

// y[t] = sum_j X[t][j] v[j] with this thread's half of row t in registers (A) or shared memory (M, column access)
__device__ __forceinline__ float cg_mv_regs(const float (&A)[CG_HALF], const float* v, int h)
{
    const float4* v4 = reinterpret_cast<const float4*>(v + CG_HALF * h);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < CG_HALF / 4; j++) {
        const float4 x = v4[j];
        s0 = fmaf(A[4 * j], x.x, s0); s1 = fmaf(A[4 * j + 1], x.y, s1); s2 = fmaf(A[4 * j + 2], x.z, s2); s3 = fmaf(A[4 * j + 3], x.w, s3);
    }
    const float s = (s0 + s1) + (s2 + s3);
    return s + __shfl_xor_sync(0xffffffffu, s, 1);
}
__device__ __forceinline__ float cg_mv_smem(const float* Ms, const float* v, int t, int h)
{
    const float4* v4 = reinterpret_cast<const float4*>(v + CG_HALF * h);
    const float* Mc = Ms + (size_t)(CG_HALF * h) * WS_KP + t;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
    for (int j = 0; j < CG_HALF / 4; j++) {
        const float4 x = v4[j];
        s0 = fmaf(Mc[(4 * j) * WS_KP], x.x, s0); s1 = fmaf(Mc[(4 * j + 1) * WS_KP], x.y, s1);
        s2 = fmaf(Mc[(4 * j + 2) * WS_KP], x.z, s2); s3 = fmaf(Mc[(4 * j + 3) * WS_KP], x.w, s3);
    }
    const float s = (s0 + s1) + (s2 + s3);
    return s + __shfl_xor_sync(0xffffffffu, s, 1);
}

// d = PCG(A~, rhs): both threads of row t hold rhs[t] on entry and d[t] on return; false if it did not converge.
__device__ __forceinline__ bool cg_pcg(const float (&A)[CG_HALF], const float* Ms, CgShared& sh, float rhs, float& x_out,
                                       int t, int h, int lane, int warp)
{
    const float mine = h == 0 ? 1.f : 0.f;          // a row's two threads carry the same vector element: count it once
    float r = rhs, x = 0.f;
    if (h == 0) sh.vec[1][t] = r;
    __syncthreads();
    float z = cg_mv_smem(Ms, sh.vec[1], t, h);
    float pv = z;
    float rz = mine * r * z, rr = mine * r * r;
    cg_sum2(rz, rr, sh.red[0], lane, warp);
    const float rr0 = rr;
    bool ok = !(rr0 > 0.f);
    for (int it = 0; it < CG_MAX_IT && !ok; it++) {
        if (h == 0) sh.vec[0][t] = pv;
        __syncthreads();
        const float Ap = cg_mv_regs(A, sh.vec[0], h);
        float pAp = mine * pv * Ap, dummy = 0.f;
        cg_sum2(pAp, dummy, sh.red[1], lane, warp);
        const float al = rz / pAp;
        x = fmaf(al, pv, x);
        r = fmaf(-al, Ap, r);
        if (h == 0) sh.vec[1][t] = r;
        __syncthreads();
        z = cg_mv_smem(Ms, sh.vec[1], t, h);
        float rz_new = mine * r * z, rr_new = mine * r * r;
        cg_sum2(rz_new, rr_new, sh.red[2 + (it & 1)], lane, warp);
        if (!(rr_new > 1e-10f * rr0)) { ok = true; break; }           // |r| <= 1e-5 |r0|
        pv = fmaf(rz_new / rz, pv, z);
        rz = rz_new;
    }
    x_out = x;
    return ok;
}

__global__ void __launch_bounds__(CG_THREADS, 2) wrmf_pcg_kernel(const SolveArgs a, const float* __restrict__ Mg)
{
    extern __shared__ float cg_smem[];
    float* Ms = cg_smem;                                             // [128][128] fp32
    CgShared& sh = *reinterpret_cast<CgShared*>(cg_smem + WS_KP * WS_KP);
    const int k = a.k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = tid >> 1, h = tid & 1;
    for (int x = tid; x < WS_KP * WS_KP / 4; x += CG_THREADS) reinterpret_cast<float4*>(Ms)[x] = reinterpret_cast<const float4*>(Mg)[x];
    __syncthreads();
    const bool row_ok = t < k;
    const bool col_ok = 4 * lane < k;
    for (int qb = blockIdx.x; a.q_lo + qb < a.q_hi; qb += gridDim.x) {
        const int u = a.order[a.q_lo + qb];
        const uint32_t beg = a.row_ptr[u], end = a.row_ptr[u + 1];
        if (beg == end) {                                           // HCp = 0 => w = 0 exactly
            if (row_ok && h == 0) a.W[(size_t)u * k + t] = 0.f;
            continue;
        }
        const float* G = a.G + (size_t)qb * WS_KP * WS_KP;
        // this thread's half of row t of A~ (read as column t: the inputs are symmetric), padding rows / columns: identity
        float A[CG_HALF];
#pragma unroll
        for (int jj = 0; jj < CG_HALF; jj++) {
            const int j = CG_HALF * h + jj;
            double v = 0.0;
            if (row_ok && j < k) v = a.HH[(size_t)j * k + t] + a.alpha * (double)G[(size_t)j * WS_KP + t] + (j == t ? a.reg : 0.0);
            else if (j == t) v = 1.0;
            A[jj] = (float)v;
            if ((jj & 15) == 15) asm volatile("" ::: "memory");     // 16 rows of loads in flight at a time, not all 64 (registers)
        }
        const double bd = row_ok ? a.bsum[(size_t)qb * WS_KP + t] * (1.0 + a.alpha) : 0.0;
        const double bmax = cg_max(fabs(bd), sh.redd[0], lane, warp);
        float x1 = 0.f;
        bool ok = cg_pcg(A, Ms, sh, (float)(bd / bmax), x1, t, h, lane, warp);
        const double w1 = (double)x1 * bmax;
        // r = b - HH w1 - lambda w1 - alpha sum_i h_i (h_i . w1): the exact operator, in double
        if (h == 0) sh.wd[t] = w1;
        __syncthreads();
        double rd = 0.0;
        if (row_ok) {
            const int g0 = CG_HALF * h, g1 = min(k, g0 + CG_HALF);
            for (int g = g0; g < g1; g++) rd -= a.HH[(size_t)g * k + t] * sh.wd[g];
        }
        rd += __shfl_xor_sync(0xffffffffu, rd, 1);
        if (row_ok) rd += bd - a.reg * w1;
        {
            double r0 = 0, r1 = 0, r2 = 0, r3 = 0, w0 = 0, w1r = 0, w2 = 0, w3 = 0;
            if (col_ok) { w0 = sh.wd[4 * lane]; w1r = sh.wd[4 * lane + 1]; w2 = sh.wd[4 * lane + 2]; w3 = sh.wd[4 * lane + 3]; }
            // a warp per entry, lane l holds factors 4 l .. 4 l + 3; two entries in flight
            uint32_t e = beg + warp;
            for (; e + CG_WARPS < end; e += 2 * CG_WARPS) {
                const int32_t ida = a.cols[e], idb = a.cols[e + CG_WARPS];
                float4 ha = make_float4(0.f, 0.f, 0.f, 0.f), hb = ha;
                if (col_ok) {
                    ha = __ldg(reinterpret_cast<const float4*>(a.H + (size_t)ida * k) + lane);
                    hb = __ldg(reinterpret_cast<const float4*>(a.H + (size_t)idb * k) + lane);
                }
                double da = (double)ha.x * w0 + (double)ha.y * w1r + (double)ha.z * w2 + (double)ha.w * w3;
                double db = (double)hb.x * w0 + (double)hb.y * w1r + (double)hb.z * w2 + (double)hb.w * w3;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { da += __shfl_xor_sync(0xffffffffu, da, o); db += __shfl_xor_sync(0xffffffffu, db, o); }
                r0 += da * (double)ha.x + db * (double)hb.x; r1 += da * (double)ha.y + db * (double)hb.y;
                r2 += da * (double)ha.z + db * (double)hb.z; r3 += da * (double)ha.w + db * (double)hb.w;
            }
            for (; e < end; e += CG_WARPS) {
                const int32_t id = a.cols[e];
                float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col_ok) hv = __ldg(reinterpret_cast<const float4*>(a.H + (size_t)id * k) + lane);
                double d = (double)hv.x * w0 + (double)hv.y * w1r + (double)hv.z * w2 + (double)hv.w * w3;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                r0 += d * (double)hv.x; r1 += d * (double)hv.y; r2 += d * (double)hv.z; r3 += d * (double)hv.w;
            }
            double* pw = sh.part[warp] + 4 * lane;
            pw[0] = r0; pw[1] = r1; pw[2] = r2; pw[3] = r3;
        }
        __syncthreads();
        if (row_ok) {
            double acc = 0.0;
#pragma unroll
            for (int x = 0; x < CG_WARPS; x++) acc += sh.part[x][t];
            rd -= a.alpha * acc;
        }
        const double rmax = cg_max(fabs(rd), sh.redd[1], lane, warp);
        double w = w1;
        if (rmax > 1e-9 * bmax) {                                   // uniform across the CTA
            float d1 = 0.f;
            ok = cg_pcg(A, Ms, sh, (float)(rd / rmax), d1, t, h, lane, warp) && ok;
            const double dd = (double)d1 * rmax;
            w = w1 + dd;
            // the correction must be small next to the solution, or the one-step argument does not hold
            const double dm = cg_max(fabs(dd), sh.redd[0], lane, warp);
            const double wm = cg_max(fabs(w), sh.redd[1], lane, warp);
            if (dm > 1e-3 * wm) ok = false;
        }
        if (!ok && tid == 0) atomicAdd(a.fail, 1u);
        if (row_ok && h == 0) a.W[(size_t)u * k + t] = (float)w;
        __syncthreads();                                            // shared scratch is reused by the next row
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------
bool wrmf_tc_eligible(int32_t k) { return k >= 4 && k <= 128 && (k % 4) == 0; }

struct WrmfTcWork {
    DevBuf<float> G; DevBuf<double> bsum; DevBuf<uint32_t> err;
    DevBuf<float> M; DevBuf<double> Mscratch;          // preconditioner of the PCG solver and its work space
    int32_t cap_rows = 0;
};
WrmfTcWork* wrmf_tc_work_create() { return new (std::nothrow) WrmfTcWork(); }
void wrmf_tc_work_destroy(WrmfTcWork* w) { delete w; }

// One half-sweep's per-row systems: W[u] <- solve for every row of `order`. HH (fp64, k x k) is on the device.
static int32_t half_sweep_impl(Ctx* ctx, WrmfTcWork* work, const uint32_t* row_ptr, const int32_t* cols, const int32_t* order, int32_t n_rows,
                               float* W, const float* H, int32_t k, const double* HH, double alpha, double reg, int64_t* launches,
                               float* debug_G_row0, const int solver, uint32_t* not_converged)
{
    cudaStream_t s = ctx->stream;
    MML_CHECK(work != nullptr, MML_ERR_STATE, "wrmf: no tensor-path workspace");
    WrmfTcWork& w = *work;
    // Rows per batch: G 1 GiB. (Running the next batch's Gram sums on a second stream under this batch's solves was tried:
    // the solve kernel's three CTAs per SM hold 61k of the 64k registers, so the Gram kernel's CTAs only get an SM once the
    // solves have drained -- no overlap, 52.0 ms either way.)
    const int32_t B = 16384;
    const int32_t cap = std::min(B, std::max(n_rows, 1));
    if (w.cap_rows < cap) {
        MML_TRY(w.G.alloc((size_t)cap * WS_KP * WS_KP)); MML_TRY(w.bsum.alloc((size_t)cap * WS_KP));
        if (!w.err.p) MML_TRY(w.err.alloc(2));
        w.cap_rows = cap;
    }
    MML_CUDA(cudaMemsetAsync(w.err.p, 0, 2 * sizeof(uint32_t), s));
    const size_t smem_syrk = (size_t)WS_STAGES * 2 * WS_TILE + 1024;
    MML_CUDA(cudaFuncSetAttribute((const void*)wrmf_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_syrk));
    const int np = (k + SV_NB - 1) / SV_NB;
    // solver: 0 = preconditioned CG + one exact refinement (default), 1 = Cholesky with the single-precision factor, 2 = double
    const bool f64 = solver == 2;
    const size_t smem_solve = solver == 0 ? sizeof(float) * WS_KP * WS_KP + sizeof(CgShared)
                                          : (f64 ? sv_smem_bytes<double>(np) : sv_smem_bytes<float>(np));
    void (*chol_fn)(const SolveArgs) = f64 ? wrmf_solve_kernel<double> : wrmf_solve_kernel<float>;
    if (solver != 0) MML_CUDA(cudaFuncSetAttribute((const void*)chol_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
    else {
        // the shared preconditioner M = (HH + lambda I)^-1 of this half-sweep
        MML_TRY(w.M.ensure((size_t)WS_KP * WS_KP)); MML_TRY(w.Mscratch.ensure((size_t)WS_KP * WS_KP));
        const size_t smem_pre = sizeof(double) * WS_KP * (WS_KP + 1);
        MML_CUDA(cudaFuncSetAttribute((const void*)wrmf_precond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pre));
        MML_CUDA(cudaFuncSetAttribute((const void*)wrmf_pcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve));
        wrmf_precond_kernel<<<1, PRE_THREADS, smem_pre, s>>>(HH, reg, k, w.Mscratch.p, w.M.p);
        MML_CUDA(cudaGetLastError());
        if (launches) *launches += 1;
    }
    for (int32_t q_lo = 0; q_lo < n_rows; q_lo += B) {
        const int32_t q_hi = std::min(n_rows, q_lo + B), nb = q_hi - q_lo;
        MML_CUDA(cudaMemsetAsync(w.bsum.p, 0, sizeof(double) * (size_t)nb * WS_KP, s));
        SyrkArgs sa{};
        sa.row_ptr = row_ptr; sa.cols = cols; sa.order = order; sa.q_lo = q_lo; sa.q_hi = q_hi; sa.H = H; sa.k = k;
        sa.G = w.G.p; sa.bsum = w.bsum.p; sa.err = w.err.p;
        wrmf_syrk_kernel<<<std::min(nb, ctx->sm_count), WS_THREADS, smem_syrk, s>>>(sa);
        MML_CUDA(cudaGetLastError());
        if (debug_G_row0 && q_lo == 0)
            MML_CUDA(cudaMemcpyAsync(debug_G_row0, w.G.p, sizeof(float) * WS_KP * WS_KP, cudaMemcpyDeviceToHost, s));
        SolveArgs va{};
        va.row_ptr = row_ptr; va.cols = cols; va.order = order; va.q_lo = q_lo; va.q_hi = q_hi; va.G = w.G.p; va.bsum = w.bsum.p; va.HH = HH;
        va.H = H; va.alpha = alpha; va.reg = reg; va.k = k; va.W = W; va.fail = w.err.p + 1;
        if (solver == 0) wrmf_pcg_kernel<<<std::min(nb, 2 * ctx->sm_count), CG_THREADS, smem_solve, s>>>(va, w.M.p);
        else chol_fn<<<nb, SV_THREADS, smem_solve, s>>>(va);
        MML_CUDA(cudaGetLastError());
        if (launches) *launches += 2;
    }
    uint32_t h_err[2] = {0, 0};
    MML_CUDA(cudaMemcpyAsync(h_err, w.err.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    MML_CHECK(h_err[0] == 0, MML_ERR_CUDA, "wrmf: tcgen05 pipeline timed out");
    *not_converged = h_err[1];
    return MML_OK;
}

int32_t wrmf_tc_half_sweep(Ctx* ctx, WrmfTcWork* work, const uint32_t* row_ptr, const int32_t* cols, const int32_t* order, int32_t n_rows,
                           float* W, const float* H, int32_t k, const double* HH, double alpha, double reg, int64_t* launches,
                           float* debug_G_row0, int first_solver)
{
    // Solvers: 0 = preconditioned conjugate gradients + one exact refinement (MML_WRMF_TENSOR_PCG), 1 = Cholesky with the
    // single-precision factor (default), 2 = with the double-precision factor (MML_WRMF_TENSOR_F64). A half-sweep that leaves
    // a row unconverged is redone by the next solver down the ladder (W is output only -- the PCG kernel does not read it --
    // and H is untouched, so a repeat is safe). MMLB200_WRMF_SOLVER = pcg | chol | chol64 overrides the first solver.
    static const int env_first = [] {
        const char* e = getenv("MMLB200_WRMF_SOLVER");
        if (e && strcmp(e, "chol64") == 0) return 2;
        if (e && strcmp(e, "chol") == 0) return 1;
        if (e && strcmp(e, "pcg") == 0) return 0;
        const char* f = getenv("MMLB200_WRMF_FACTOR");
        return (f && strcmp(f, "fp64") == 0) ? 2 : -1;
    }();
    uint32_t bad = 0;
    for (int solver = env_first >= 0 ? env_first : first_solver; solver <= 2; solver++) {
        MML_TRY(half_sweep_impl(ctx, work, row_ptr, cols, order, n_rows, W, H, k, HH, alpha, reg, launches, debug_G_row0, solver, &bad));
        if (bad == 0) return MML_OK;
    }
    MML_CHECK(bad == 0, MML_ERR_CUDA, "wrmf: %u rows did not converge (system too ill-conditioned for the 1e-9 residual bound)", bad);
    return MML_OK;
}

}  // namespace mml
