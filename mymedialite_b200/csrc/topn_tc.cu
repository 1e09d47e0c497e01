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
#include <chrono>
#include <string>
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>

namespace mml {

constexpr int TC_ROWS = 256;                  // users per CTA: two 128-row MMA tiles sharing every V tile
constexpr int TC_N = 128;                     // candidates per MMA tile
constexpr int TC_KC = 32;                     // TF32 filter: floats per K chunk = one 128-byte swizzle atom
constexpr int TC_KC_BF16 = 64;                // BF16 filter: elements per K chunk (same 128 bytes)
constexpr int TC_CHUNK_BYTES = 128 * 128;     // 128 rows x 128 bytes
constexpr int TC_CAP = 128;                   // candidates collected per epilogue thread (2 threads per user and candidate split) before the user goes to the exact path
constexpr int TC_SAMPLE = 16384;              // target size of the candidate sample the threshold comes from
constexpr int TC_MAX_N = 16;                  // largest n served by this path
constexpr int TC_THREADS = 576;               // warp 0: TMA, warp 1: MMA issue + TMEM, warps 2-17: epilogue (2 threads per user row)
constexpr float TC_ERR_C = 0.0025f;           // |tf32 score - exact| <= TC_ERR_C * |u| * |v| (2^-9 truncation + slack)
// BF16 filter: both operands are rounded to nearest bf16 (relative error <= 2^-9 each), so every product is off by at most
// (2^-8 + 2^-18) |u_f v_f|; the fp32 accumulation of <= 128 products and the fp32 rounding of the exact sequential sum add
// < 4e-5 sum|u_f v_f|; with Cauchy-Schwarz |approx - exact| <= 0.00395 |u| |v|. Slack on top:
constexpr float TC_ERR_C_BF16 = 0.0045f;
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
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of a converged warp (elect.sync). The single-thread roles run as WHOLE warps in uniform control flow with the
// issuing instructions under this predicate: inside `if (lane == 0)` ptxas wraps every tcgen05.mma / TMA instruction (their
// operands live in uniform registers) in an ELECT / R2UR / BRA.U.ANY loop -- ~16 instructions and ~130 cycles per MMA in the
// round-1 kernel, i.e. the issuing thread, not the tensor pipe, set the pace (ncu: the issuer busy all the time, 49 % of its
// samples fixed-latency waits, tensor pipe 38 % active).
__device__ __forceinline__ bool tc_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
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
// The same descriptor as two words: the low one carries the (address >> 4) field, so the descriptor of the same tile 32 bytes
// further along K is low + 2; the high word is a constant. The issuing thread builds its descriptors this way: one add per
// operand and MMA instead of a shift / mask / or chain per descriptor (the issuing thread's instruction stream is what the
// tensor pipe waits for: ncu after the elect.sync change still showed it busy 90 % of the time at 7 cycles per instruction).
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t TC_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ void tc_mma_tf32_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(TC_DESC_HI) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(TC_DESC_HI) : "memory");
}
// Instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128.
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((128u >> 4) << 24);
// The same with A = B = BF16 (kind::f16; K = 16 per instruction = the same 32 bytes of a swizzled row as 8 TF32 values).
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((128u >> 4) << 24);

struct TcArgs {
    int32_t n_rows;                 // users in the batch
    int32_t stride;                 // column j of the V panel is candidate position j * stride (1 when collecting)
    int32_t kc;                     // K chunks (kp / elements per chunk)
    int32_t epc;                    // elements per K chunk: 32 (TF32 panels) or 64 (BF16 panels)
    int32_t bf16;                   // operand panels hold bf16 (kind::f16 MMAs) instead of fp32 (kind::tf32)
    int32_t stages;                 // V ring depth
    int32_t list_off;               // byte offset of the sampling lists behind the aligned operand area
    int32_t n_tiles, tiles_per_split, splits;
    int32_t n_cols;                 // real columns of the V panel (the rest is zero padding)
    int32_t m;                      // sampling pass: length of the per-user sorted list (>= n)
    int32_t dbg;                    // experiments (MMLB200_TC_DBG): 1 = epilogue does not examine, 2 = no MMAs are issued
    const uint8_t* row_ok;          // [n_rows] user id inside the model
    const uint32_t* bad;            // NULL, or [n_tiles * 4]: bit c of word g set = column 32 g + c is an id outside the model
    const int64_t* ign_ptr;         // [n_rows + 1] or NULL
    const int32_t* ign_pos;         // candidate POSITIONS of the user's ignore_items, ascending inside a row
    float* thr;                     // [n_rows] sampling pass: out, m-th best sampled score; collecting pass: in, threshold
    float* buf_s; int32_t* buf_p;   // collecting pass: [n_rows][splits][2][TC_CAP] approximate score, position
    int32_t* buf_n;                 // [n_rows][splits][2] entries wanted (> TC_CAP = overflow)
    uint32_t* err;
};

// Sampling pass slow path: s enters the sorted list ls[0 .. m) (stride 512 floats: one column of shared memory per thread).
__device__ __noinline__ float tc_insert(float* ls, int m, float s)
{
    int j = m - 1;
    while (j > 0 && ls[(j - 1) * 2 * TC_ROWS] < s) { ls[j * 2 * TC_ROWS] = ls[(j - 1) * 2 * TC_ROWS]; j--; }
    ls[j * 2 * TC_ROWS] = s;
    return ls[(m - 1) * 2 * TC_ROWS];
}

// MODE 0 (sampling): every thread keeps the m best approximate scores of its user over a strided sample of the candidates;
//                    the m-th best, minus twice the error bound, is a threshold no member of the exact top n can fall below.
// MODE 1 (collecting): every thread appends the candidates of its user that reach the threshold (a few dozen of 10^5).
// ignore_items are skipped with a cursor into the user's position-sorted ignore list: one register compare per 32 scores.
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
score_select_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_v, const TcArgs a)
{
    extern __shared__ uint8_t tc_smem_raw[];
    __shared__ uint64_t bars[2 * 8 + 1 + 4];      // full[8], empty[8], u_full, tmem_full[2], tmem_empty[2]
    __shared__ uint32_t tmem_base_s;
    const uint32_t smem0 = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_u = smem0;                                               // [2][kc] chunks
    const uint32_t smem_v = smem0 + 2u * a.kc * TC_CHUNK_BYTES;                  // [stages] chunks
    float* lists = reinterpret_cast<float*>(tc_smem_raw + (smem0 - smem_u32(tc_smem_raw)) + (size_t)a.list_off);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * 8, bar_u = bar_full + 16 * 8;
    const uint32_t bar_tfull = bar_full + 17 * 8, bar_tempty = bar_full + 19 * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp = blockIdx.y;
    const int t_begin = sp * a.tiles_per_split, t_end = min(t_begin + a.tiles_per_split, a.n_tiles);
    const int row0 = blockIdx.x * TC_ROWS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_u, 1);
        for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 16); }
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
        // ===== TMA producer (whole warp, one elected lane issues) =====
        if (tc_elect_one()) {
            mbar_expect_tx(bar_u, 2u * a.kc * TC_CHUNK_BYTES);
            for (int h = 0; h < 2; h++)
                for (int c = 0; c < a.kc; c++)
                    tma_load_2d(smem_u + (uint32_t)(h * a.kc + c) * TC_CHUNK_BYTES, &map_u, bar_u, c * a.epc, row0 + h * 128);
        }
        __syncwarp();
        int stage = 0; uint32_t phase = 0;
        for (int t = t_begin; t < t_end; t++)
            for (int c = 0; c < a.kc; c++) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u, a.err);
                if (tc_elect_one()) {
                    mbar_expect_tx(bar_full + 8 * stage, TC_CHUNK_BYTES);
                    tma_load_2d(smem_v + (uint32_t)stage * TC_CHUNK_BYTES, &map_v, bar_full + 8 * stage, c * a.epc, t * TC_N);
                }
                __syncwarp();
                if (++stage == a.stages) { stage = 0; phase ^= 1u; }
            }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp, one elected lane issues) =====
        mbar_wait(bar_u, 0, a.err);
        tc_fence_after();
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        const int kc = a.kc, n_stages = a.stages;
        const bool use_bf16 = a.bf16 != 0, no_mma = (a.dbg & 2) != 0;
        const uint32_t u_lo = tc_desc_lo(smem_u), v_lo = tc_desc_lo(smem_v);
        const uint32_t half_lo = (uint32_t)kc * (TC_CHUNK_BYTES >> 4);      // descriptor distance of the two row halves of U
        for (int t = t_begin; t < t_end; t++, it++) {
            const int buf = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(bar_tempty + 8 * buf, aphase ^ 1u, a.err);
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(buf * 2 * TC_N), d1 = d0 + (uint32_t)TC_N;
            for (int c = 0; c < kc; c++) {
                mbar_wait(bar_full + 8 * stage, phase, a.err);
                tc_fence_after();
                if (tc_elect_one()) {
                    const uint32_t a0 = u_lo + (uint32_t)c * (TC_CHUNK_BYTES >> 4), a1 = a0 + half_lo;
                    const uint32_t b0 = v_lo + (uint32_t)stage * (TC_CHUNK_BYTES >> 4);
                    if (!no_mma) {
                        if (use_bf16) {
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) tc_mma_bf16_lo(d0, a0 + 2 * kk, b0 + 2 * kk, TC_IDESC_BF16, (c | kk) != 0 ? 1u : 0u);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) tc_mma_bf16_lo(d1, a1 + 2 * kk, b0 + 2 * kk, TC_IDESC_BF16, (c | kk) != 0 ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) tc_mma_tf32_lo(d0, a0 + 2 * kk, b0 + 2 * kk, TC_IDESC, (c | kk) != 0 ? 1u : 0u);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) tc_mma_tf32_lo(d1, a1 + 2 * kk, b0 + 2 * kk, TC_IDESC, (c | kk) != 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(bar_empty + 8 * stage);          // frees the V chunk when these MMAs have read it
                    if (c + 1 == kc) tc_commit(bar_tfull + 8 * buf);     // ... and, after the tile's last chunk, its accumulators are complete
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: 16 warps; two threads per user row, each examines two of the four 32-column groups of a tile =====
        const int e = warp - 2;
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const int h = (e >> 2) & 1;                            // row half
        const int part = e >> 3;                               // column groups 2 part, 2 part + 1 of every tile
        const int rloc = h * 128 + q * 32 + lane;
        const int row = row0 + rloc;
        const bool ok = row < a.n_rows && a.row_ok[row];
        float* ls = lists + part * TC_ROWS + rloc;             // entry j at ls[j * 2 * TC_ROWS]
        float thr;
        if (MODE == 0) {
            for (int j = 0; j < a.m; j++) ls[j * 2 * TC_ROWS] = -INFINITY;
            thr = ok ? -INFINITY : INFINITY;
        } else {
            thr = ok ? a.thr[row] : INFINITY;
        }
        int cnt = 0;
        const size_t bbase = (((size_t)(ok ? row : 0) * a.splits + sp) * 2 + part) * TC_CAP;
        int64_t ig = 0, ig_end = 0;
        int32_t ig_next = 0x7fffffff;
        if (ok && a.ign_ptr) {
            ig = a.ign_ptr[row]; ig_end = a.ign_ptr[row + 1];
            if (ig < ig_end) ig_next = a.ign_pos[ig];
        }
        const long long st = a.stride;
        const int n_groups = (t_end - t_begin) * 2;
        // group gi = 32 consecutive columns; its accumulators are fetched while group gi - 1 is being examined
        uint32_t va[32], vb[32];
        auto fetch = [&](int gi, uint32_t (&v)[32]) {
            const int it = gi >> 1, j = part * 2 + (gi & 1);
            const int buf = it & 1;
            if ((gi & 1) == 0) {      // one lane polls: 512 threads spinning on the barrier unit delay the MMA / TMA hand-overs
                if (lane == 0) mbar_wait(bar_tfull + 8 * buf, (uint32_t)(it >> 1) & 1u, a.err);
                __syncwarp();
                tc_fence_after();
            }
            __syncwarp();
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + h) * TC_N + j * 32), v);
        };
        auto examine = [&](int gi, uint32_t (&v)[32]) {
            if (a.dbg & 1) return;
            const int g = (t_begin + (gi >> 1)) * (TC_N / 32) + part * 2 + (gi & 1);
            const int col0 = g * 32;
            uint32_t mask = a.bad ? a.bad[g] : 0u;
            if (col0 + 32 > a.n_cols) mask |= col0 >= a.n_cols ? 0xffffffffu : (0xffffffffu << (a.n_cols - col0));
            if ((long long)ig_next < (long long)(col0 + 32) * st) {     // some ignore_items fall into these 32 columns
                do {
                    const long long p = ig_next;
                    if (p >= (long long)col0 * st && p % st == 0) mask |= 1u << (int)(p / st - col0);
                    ++ig;
                    ig_next = ig < ig_end ? a.ign_pos[ig] : 0x7fffffff;
                } while ((long long)ig_next < (long long)(col0 + 32) * st);
            }
            if (mask) {
#pragma unroll
                for (int c = 0; c < 32; c++) if ((mask >> c) & 1u) v[c] = 0x7fc00000u;      // NaN: fails every comparison, fmaxf skips it
            }
            float m8[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                m8[i] = __uint_as_float(v[8 * i]);
#pragma unroll
                for (int c = 1; c < 8; c++) m8[i] = fmaxf(m8[i], __uint_as_float(v[8 * i + c]));
            }
            const float m = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
            if (MODE == 0) {
                if (m > thr) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (m8[i] > thr) {
#pragma unroll
                            for (int c = 8 * i; c < 8 * i + 8; c++) {
                                const float s = __uint_as_float(v[c]);
                                if (s > thr) thr = tc_insert(ls, a.m, s);
                            }
                        }
                    }
                }
            } else {
                if (m >= thr) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (m8[i] >= thr) {
#pragma unroll
                            for (int c = 8 * i; c < 8 * i + 8; c++) {
                                const float s = __uint_as_float(v[c]);
                                if (s >= thr) {
                                    if (cnt < TC_CAP) { a.buf_s[bbase + cnt] = s; a.buf_p[bbase + cnt] = col0 + c; }
                                    cnt++;
                                }
                            }
                        }
                    }
                }
            }
        };
        auto release = [&](int gi) {      // last group of a tile examined: its TMEM buffer may be overwritten
            if (gi & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * ((gi >> 1) & 1));
            }
        };
        // n_groups is even (two groups per tile and thread). A tile's TMEM buffer is handed back as soon as its second
        // group sits in registers -- before this warp blocks on the next tile's accumulators and before it examines the
        // group -- so that the MMAs of tile t + 2 start while the second half of tile t is still being examined.
        // (Round 2: both groups of a tile are now in registers BEFORE the first is examined -- the hand-back used to wait for
        // examine(gi), which sat between the second fetch and its wait; with two TMEM buffers the MMAs of tile t + 2 wait for
        // exactly that hand-back.)
        if (n_groups > 0) fetch(0, va);
        for (int gi = 0; gi < n_groups; gi += 2) {
            tc_ld_wait();                                   // va holds group gi (first group of its tile)
            fetch(gi + 1, vb);                              // same tile: no barrier
            tc_ld_wait();                                   // vb holds group gi + 1: the tile's buffer has been read
            release(gi + 1);
            examine(gi, va);
            if (gi + 2 < n_groups) fetch(gi + 2, va);       // waits for the next tile's MMAs
            examine(gi + 1, vb);
        }
        if (MODE == 0) {
            // the user's two threads hold the m best of their halves of the sample (sorted): m-th best of the union
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (ok && part == 0) {
                const float* la = lists + rloc; const float* lb = lists + TC_ROWS + rloc;
                int ia = 0, ib = 0;
                float t = -INFINITY;
                for (int x = 0; x < a.m; x++) {
                    const float fa = la[ia * 2 * TC_ROWS], fb = lb[ib * 2 * TC_ROWS];
                    if (fa >= fb) { t = fa; ia++; } else { t = fb; ib++; }
                }
                a.thr[row] = t;
            }
        } else if (ok) {
            a.buf_n[((size_t)row * a.splits + sp) * 2 + part] = cnt;
        }
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
// dst[r][0..kp) = src[id(r)][0..k) (zero beyond k, zero row for ids outside the model), id(r) = ids ? ids[r * id_stride] :
// r * id_stride; norm[r] = |row|; ok[r] = id inside the model; *max_norm_bits = max over rows (float bits; NaN sorts
// above +inf); bad: bit (r % 32) of word r / 32 set for rows whose id is outside the model.
__global__ void tc_stage_rows_kernel(const float* __restrict__ src, int32_t n_src_rows, int32_t k,
                                     const int32_t* __restrict__ ids, int32_t id_stride, int32_t n, int32_t kp,
                                     float* __restrict__ dst, float* __restrict__ norm, uint8_t* __restrict__ ok,
                                     uint32_t* __restrict__ max_norm_bits, uint32_t* __restrict__ bad, int bf16)
{
    __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(dst);      // bf16 panels: kp elements of 2 bytes per row
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n; r += stride) {
        const int32_t id = ids ? ids[r * id_stride] : (int32_t)(r * id_stride);
        const bool valid = (uint32_t)id < (uint32_t)n_src_rows;
        float ss = 0.f;
        for (int f = lane; f < kp; f += 32) {
            const float x = (valid && f < k) ? src[(size_t)id * k + f] : 0.f;
            if (bf16) dst16[(size_t)r * kp + f] = __float2bfloat16_rn(x);
            else dst[(size_t)r * kp + f] = x;
            ss = fmaf(x, x, ss);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
        if (lane == 0) {
            const float nr = sqrtf(ss);
            if (norm) norm[r] = nr;
            if (ok) ok[r] = valid ? 1 : 0;
            if (max_norm_bits) atomicMax(max_norm_bits, __float_as_uint(nr));
            if (bad && !valid) atomicOr(bad + (r >> 5), 1u << (r & 31));
        }
    }
}

// pos_of[item] = candidate position (candidates are distinct)
__global__ void tc_pos_of_kernel(const int32_t* __restrict__ cand, int32_t n_cand, int32_t n_items, int32_t* __restrict__ pos_of)
{
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cand) { const int32_t id = cand[c]; if ((uint32_t)id < (uint32_t)n_items) pos_of[id] = c; }
}

// ignore item ids -> candidate positions (INT_MAX: not a candidate), row of every entry (binary search in the CSR
// pointers), and an "ascending inside every row" check
__global__ void tc_ignore_pos_kernel(const int64_t* __restrict__ ptr, int32_t n_rows, int32_t* __restrict__ idx, int64_t total,
                                     const int32_t* __restrict__ pos_of, int32_t n_items, int32_t n_cand,
                                     uint32_t* __restrict__ row_of)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        int32_t lo = 0, hi = n_rows;                  // last row with ptr[row] <= t
        while (hi - lo > 1) { const int32_t mid = (lo + hi) >> 1; if (ptr[mid] <= t) lo = mid; else hi = mid; }
        row_of[t] = (uint32_t)lo;
        const int32_t item = idx[t];
        int32_t pos = 0x7fffffff;
        if (pos_of) { if ((uint32_t)item < (uint32_t)n_items && pos_of[item] >= 0) pos = pos_of[item]; }
        else if ((uint32_t)item < (uint32_t)n_cand) pos = item;
        idx[t] = pos;
    }
}

__global__ void tc_ignore_sorted_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ row_of,
                                        const int32_t* __restrict__ pos, int64_t total, uint32_t* __restrict__ unsorted)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride)
        if (t + 1 < ptr[row_of[t] + 1] && pos[t] > pos[t + 1]) *unsorted = 1u;   // every writer stores the same value
}

__global__ void tc_copy_u32_kernel(const uint32_t* __restrict__ src, int64_t total, uint32_t* __restrict__ dst)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < total; t += stride) dst[t] = src[t];
}

// thr[b] = (m-th best sampled score) - 2 d, d2[b] = 2 d, d = err_c * |u_b| * max|v| (+ underflow slack)
__global__ void tc_threshold_kernel(float* __restrict__ thr, float* __restrict__ d2, const float* __restrict__ unorm,
                                    const uint32_t* __restrict__ vmax_bits, int32_t n_rows, float err_c)
{
    const int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_rows) return;
    const float dd = 2.f * (err_c * unorm[b] * __uint_as_float(*vmax_bits) + 1e-30f);
    d2[b] = dd;
    thr[b] = (dd < INFINITY) ? thr[b] - dd : -INFINITY;
}

// ---- finalize: exact re-scoring of the candidate superset, one warp per user ---------------------------------
struct FinArgs {
    const float* U; const float* V; int32_t k;            // original model matrices
    const int32_t* users;                                   // [n_rows] user ids
    const int32_t* cand;                                    // or NULL
    const float* buf_s; const int32_t* buf_p; const int32_t* buf_n;
    const float* thr; const float* d2;
    const uint8_t* row_ok;
    int32_t n_rows, splits, n, n_out, warps;
    int32_t* out_items; float* out_scores; int32_t* out_counts; uint8_t* redo;
};

__global__ void tc_finalize_kernel(const FinArgs a)
{
    extern __shared__ uint8_t fin_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int parts = a.splits * 2;                                                  // two collecting threads per user and split
    const int E_max = parts * TC_CAP;
    float* ap = reinterpret_cast<float*>(fin_smem) + (size_t)warp * E_max * 3;      // approximate scores
    int32_t* pp = reinterpret_cast<int32_t*>(ap + E_max);                            // positions
    float* ex = reinterpret_cast<float*>(pp + E_max);                                // exact scores
    const int b = blockIdx.x * a.warps + warp;
    if (b >= a.n_rows) return;
    if (!a.row_ok[b]) { if (lane == 0) { a.out_counts[b] = 0; a.redo[b] = 0; } return; }
    // 1. gather what the splits collected; a split that wanted more than TC_CAP entries lost some
    int E = 0;
    bool overflow = false;
    for (int s = 0; s < parts; s++) {
        const int c = a.buf_n[(size_t)b * parts + s];
        if (c > TC_CAP) { overflow = true; break; }
        const size_t base = ((size_t)b * parts + s) * TC_CAP;
        for (int e = lane; e < c; e += 32) { ap[E + e] = a.buf_s[base + e]; pp[E + e] = a.buf_p[base + e]; }
        E += c;
    }
    __syncwarp();
    const float d2 = a.d2[b], thr = a.thr[b];
    // thr = -inf: the sample held fewer than m scorable candidates and everything scorable was collected.
    // thr > -inf: at least m >= n candidates reach it, anything else is a broken invariant -> exact path.
    if (overflow || !(d2 < INFINITY) || (thr > -INFINITY && E < a.n)) { if (lane == 0) { a.redo[b] = 1; a.out_counts[b] = 0; } return; }
    if (E == 0) { if (lane == 0) { a.out_counts[b] = 0; a.redo[b] = 0; } return; }
    // 2. a_n = n-th best approximate score (rank by score desc, buffer index asc)
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
    const float cut = a_n - d2;        // every member of the exact top n scores >= cut approximately; cut >= thr
    // 3. exact scores of the finalists: sequential fp32 multiply, then add (MatrixExtensions.cs:234-238)
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
    // 4. order by (exact score desc, candidate position asc); the best n go out
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

static int32_t make_panel_map(CUtensorMap* map, float* base, int64_t rows, int32_t kp, bool bf16)
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
    const cuuint64_t strides[1] = { (cuuint64_t)kp * (bf16 ? 2 : 4) };
    const cuuint32_t box[2] = { (cuuint32_t)(bf16 ? TC_KC_BF16 : TC_KC), 128 };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MML_CHECK(r == CUDA_SUCCESS, MML_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MML_OK;
}

static inline int tc_grid(int64_t n, int threads = 256)
{
    return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n, threads), 1), 148 * 16);
}

// Operand precision of the filter GEMM: TF32 (default) or BF16 (twice the tensor rate, half the operand bytes; the wider
// error bound lengthens the exactly re-scored candidate list and sends more users to the exact path). Measured on config 5
// (gpurun_out/pp_topn.csv, qq_topn_trace.log): the collecting pass is bound by operand staging, not by the MMA rate, so
// BF16 gains 6 % there and loses it again in the finalize kernel. (Round 2, after the MMA-issuing thread stopped being the
// bound: 90.9 ms per call with BF16 against 92.0 ms with TF32, 665 instead of 29 users on the exact path -- still no reason to
// switch.) mml_topn_set_filter / MMLB200_TC_FILTER=bf16|tf32.
static int g_tc_filter = -1;
void topn_tc_set_filter(int kind) { g_tc_filter = kind; }
bool topn_tc_filter_bf16()
{
    if (g_tc_filter < 0) {
        const char* e = getenv("MMLB200_TC_FILTER");
        g_tc_filter = (e && strcmp(e, "bf16") == 0) ? MML_TOPN_FILTER_BF16 : MML_TOPN_FILTER_TF32;
    }
    return g_tc_filter == MML_TOPN_FILTER_BF16;
}

bool topn_tc_eligible(int32_t k, int32_t n, int64_t n_cand)
{
    return n >= 1 && n <= TC_MAX_N && k >= 1 && k <= 128 && n_cand >= 1 && n_cand < ((int64_t)1 << 30);
}

// Candidate-side state of one Recommend() call, shared by all user batches.
struct TcCandidates {
    int32_t n_cand = 0, kp = 0, kc = 0, stride = 1, n_samp = 0;
    bool bf16 = false;                    // panels hold bf16 (kp elements of 2 bytes per row) instead of fp32
    size_t row_floats() const { return bf16 ? (size_t)kp / 2 : (size_t)kp; }      // panel row length in 4-byte units
    int64_t cand_pad = 0, samp_pad = 0;
    bool has_bad = false;                 // some candidate id lies outside the model
    DevBuf<float> Vb, Vs;                 // all candidates / the strided sample, zero padded panels
    DevBuf<uint32_t> bad_b, bad_s, vmax;  // never-qualifying columns of either panel; max |v| (float bits)
    DevBuf<int32_t> pos_of;               // explicit candidate lists only: item -> position
    CUtensorMap map_b, map_s;
};

static int32_t tc_prepare_candidates(Ctx* ctx, TcCandidates& c, const float* d_V, int32_t n_model_items, int32_t k,
                                     const int32_t* d_cand, int32_t n_cand, bool has_invalid, int32_t n, int64_t* launches)
{
    cudaStream_t s = ctx->stream;
    c.n_cand = n_cand; c.has_bad = has_invalid;
    const int32_t epc = c.bf16 ? TC_KC_BF16 : TC_KC;
    c.kp = (int32_t)ceil_div(k, epc) * epc; c.kc = c.kp / epc;
    const size_t rf = c.row_floats();
    c.cand_pad = ceil_div(n_cand, TC_N) * TC_N;
    // sample: every stride-th position; about n * stride candidates reach the threshold it yields
    c.stride = (int32_t)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_cand, TC_SAMPLE), std::max(1, 96 / n)));
    c.n_samp = (int32_t)ceil_div(n_cand, c.stride);
    c.samp_pad = ceil_div(c.n_samp, TC_N) * TC_N;
    MML_TRY(c.Vb.ensure((size_t)c.cand_pad * rf)); MML_TRY(c.Vs.ensure((size_t)c.samp_pad * rf));
    MML_TRY(c.bad_b.ensure((size_t)c.cand_pad / 32)); MML_TRY(c.bad_s.ensure((size_t)c.samp_pad / 32)); MML_TRY(c.vmax.ensure(1));
    MML_CUDA(cudaMemsetAsync(c.bad_b.p, 0, c.bad_b.bytes(), s)); MML_CUDA(cudaMemsetAsync(c.bad_s.p, 0, c.bad_s.bytes(), s));
    MML_CUDA(cudaMemsetAsync(c.vmax.p, 0, sizeof(uint32_t), s));
    if (c.cand_pad > n_cand) MML_CUDA(cudaMemsetAsync(c.Vb.p + (size_t)n_cand * rf, 0, sizeof(float) * (size_t)(c.cand_pad - n_cand) * rf, s));
    if (c.samp_pad > c.n_samp) MML_CUDA(cudaMemsetAsync(c.Vs.p + (size_t)c.n_samp * rf, 0, sizeof(float) * (size_t)(c.samp_pad - c.n_samp) * rf, s));
    tc_stage_rows_kernel<<<tc_grid((int64_t)n_cand * 32), 256, 0, s>>>(d_V, n_model_items, k, d_cand, 1, n_cand, c.kp, c.Vb.p, nullptr, nullptr, c.vmax.p, c.bad_b.p, c.bf16 ? 1 : 0);
    tc_stage_rows_kernel<<<tc_grid((int64_t)c.n_samp * 32), 256, 0, s>>>(d_V, n_model_items, k, d_cand, c.stride, c.n_samp, c.kp, c.Vs.p, nullptr, nullptr, nullptr, c.bad_s.p, c.bf16 ? 1 : 0);
    if (d_cand) {
        MML_TRY(c.pos_of.ensure(std::max(n_model_items, 1)));
        MML_CUDA(cudaMemsetAsync(c.pos_of.p, 0xff, c.pos_of.bytes(), s));
        tc_pos_of_kernel<<<(unsigned)ceil_div(n_cand, 256), 256, 0, s>>>(d_cand, n_cand, n_model_items, c.pos_of.p);
    }
    MML_CUDA(cudaGetLastError());
    if (launches) *launches += 3;
    MML_TRY(make_panel_map(&c.map_b, c.Vb.p, c.cand_pad, c.kp, c.bf16));
    MML_TRY(make_panel_map(&c.map_s, c.Vs.p, c.samp_pad, c.kp, c.bf16));
    return MML_OK;
}

// Per-batch buffers of Recommend(): grow-only members of the context's cache (no cudaMalloc / cudaFree per call -- a fresh
// 1.3 GB workspace per call cost 3 to 80 ms on round 1's boxes). The per-batch inputs (users, ignore CSR) and outputs
// (lists, counts, redo flags, the pipeline error word) exist twice: batch b + 1 is uploaded and batch b - 1 downloaded
// through page-locked staging on the copy streams while batch b's kernels run.
struct TcWork {
    int32_t cap_users = 0, splits = 1, tps = 1, stages = 2, stages0 = 2;
    size_t smem = 0, smem0 = 0;
    DevBuf<float> Ub, unorm, thr, d2, buf_s;
    DevBuf<uint8_t> row_ok;
    DevBuf<int32_t> buf_p, buf_n;
    DevBuf<uint32_t> row_of, key, t1, t2;
    // double-buffered per batch
    DevBuf<int32_t> users[2], ign_idx[2], out_i[2], out_c[2];
    DevBuf<int64_t> ign_ptr[2];
    DevBuf<float> out_s[2];
    DevBuf<uint8_t> redo[2];
    DevBuf<uint32_t> err[2];
    PinBuf<int32_t> h_users[2], h_ign_idx[2], h_out_i[2], h_out_c[2];
    PinBuf<int64_t> h_ign_ptr[2];
    PinBuf<float> h_out_s[2];
    PinBuf<uint8_t> h_redo[2];
    PinBuf<uint32_t> h_err[2];
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

struct TopnCache {
    TcCandidates c;
    TcWork w;
};

static TopnCache* topn_cache(Ctx* ctx)
{
    if (!ctx->topn_cache) ctx->topn_cache = new (std::nothrow) TopnCache();
    return reinterpret_cast<TopnCache*>(ctx->topn_cache);
}

void topn_cache_destroy(Ctx* ctx)
{
    TopnCache* t = reinterpret_cast<TopnCache*>(ctx->topn_cache);
    if (!t) return;
    for (int x = 0; x < 2; x++) {
        if (t->w.ev_in[x]) cudaEventDestroy(t->w.ev_in[x]);
        if (t->w.ev_done[x]) cudaEventDestroy(t->w.ev_done[x]);
        if (t->w.ev_out[x]) cudaEventDestroy(t->w.ev_out[x]);
    }
    delete t;
    ctx->topn_cache = nullptr;
}

static int32_t tc_alloc_work(Ctx* ctx, TcWork& w, const TcCandidates& c, int32_t cap_users, int32_t n_out, int64_t max_ign)
{
    w.cap_users = cap_users;
    const int64_t rows_pad = ceil_div(cap_users, TC_ROWS) * TC_ROWS;
    const int row_tiles = (int)(rows_pad / TC_ROWS), n_tiles = (int)(c.cand_pad / TC_N);
    // candidate splits fill the GPU when the batch has few user tiles
    int splits = (int)std::min<int64_t>(std::min<int64_t>(ceil_div(ctx->sm_count, row_tiles), TC_MAX_SPLITS), n_tiles);
    splits = std::max(splits, 1);
    w.tps = (int)ceil_div(n_tiles, splits);
    w.splits = (int)ceil_div(n_tiles, w.tps);
    MML_TRY(w.Ub.ensure((size_t)rows_pad * c.row_floats()));
    MML_TRY(w.unorm.ensure(cap_users)); MML_TRY(w.thr.ensure(cap_users)); MML_TRY(w.d2.ensure(cap_users));
    MML_TRY(w.row_ok.ensure(cap_users));
    MML_TRY(w.buf_s.ensure((size_t)cap_users * w.splits * 2 * TC_CAP)); MML_TRY(w.buf_p.ensure((size_t)cap_users * w.splits * 2 * TC_CAP));
    MML_TRY(w.buf_n.ensure((size_t)cap_users * w.splits * 2));
    if (max_ign > 0) {
        MML_TRY(w.row_of.ensure(max_ign)); MML_TRY(w.key.ensure(max_ign)); MML_TRY(w.t1.ensure(max_ign)); MML_TRY(w.t2.ensure(max_ign));
    }
    for (int x = 0; x < 2; x++) {
        MML_TRY(w.users[x].ensure(cap_users)); MML_TRY(w.out_i[x].ensure((size_t)cap_users * n_out)); MML_TRY(w.out_s[x].ensure((size_t)cap_users * n_out));
        MML_TRY(w.out_c[x].ensure(cap_users)); MML_TRY(w.redo[x].ensure(cap_users)); MML_TRY(w.err[x].ensure(1));
        MML_TRY(w.h_users[x].ensure(cap_users)); MML_TRY(w.h_out_i[x].ensure((size_t)cap_users * n_out)); MML_TRY(w.h_out_s[x].ensure((size_t)cap_users * n_out));
        MML_TRY(w.h_out_c[x].ensure(cap_users)); MML_TRY(w.h_redo[x].ensure(cap_users)); MML_TRY(w.h_err[x].ensure(1));
        if (max_ign > 0) {
            MML_TRY(w.ign_ptr[x].ensure((size_t)cap_users + 1)); MML_TRY(w.ign_idx[x].ensure(max_ign));
            MML_TRY(w.h_ign_ptr[x].ensure((size_t)cap_users + 1)); MML_TRY(w.h_ign_idx[x].ensure(max_ign));
        }
        if (!w.ev_in[x]) {
            MML_CUDA(cudaEventCreateWithFlags(&w.ev_in[x], cudaEventDisableTiming));
            MML_CUDA(cudaEventCreateWithFlags(&w.ev_done[x], cudaEventDisableTiming));
            MML_CUDA(cudaEventCreateWithFlags(&w.ev_out[x], cudaEventDisableTiming));
        }
    }
    int max_optin = 0;
    MML_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    // the sampling pass keeps 2 x 256 sorted lists of up to TC_MAX_N floats behind the operand area, the collecting pass does not
    const int64_t list_bytes = 2ll * TC_ROWS * TC_MAX_N * sizeof(float);
    const int64_t fixed = 2ll * c.kc * TC_CHUNK_BYTES + 1024 + 512;      // U panels, alignment slack, static barriers
    w.stages0 = (int)std::min<int64_t>(8, (max_optin - fixed - list_bytes) / TC_CHUNK_BYTES);
    w.stages = (int)std::min<int64_t>(8, (max_optin - fixed) / TC_CHUNK_BYTES);
    MML_CHECK(w.stages0 >= 2, MML_ERR_UNSUPPORTED, "topn: shared memory too small for the tcgen05 path");
    w.smem0 = (size_t)(2ll * c.kc + w.stages0) * TC_CHUNK_BYTES + 1024 + list_bytes;
    w.smem = (size_t)(2ll * c.kc + w.stages) * TC_CHUNK_BYTES + 1024;
    MML_CUDA(cudaFuncSetAttribute((const void*)score_select_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem0));
    MML_CUDA(cudaFuncSetAttribute((const void*)score_select_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
    return MML_OK;
}

// Top n of a user batch on the tensor-core path; users and the ignore CSR (item ids) are already in w.users[x] / w.ign_*[x].
// w.redo[x][b] = 1 for users whose candidate superset could not be proven complete (caller re-runs them exactly).
// Results: w.out_i[x] / w.out_s[x] [n_users x n_out], w.out_c[x] [n_users]. Nothing here waits for the device.
static int32_t topn_tc_batch(Ctx* ctx, TcCandidates& c, TcWork& w, int x, const float* d_U, int32_t n_model_users, const float* d_V,
                             int32_t n_model_items, int32_t k, int32_t n_users, int32_t n, int32_t n_out, const int32_t* d_cand,
                             int64_t n_ign, int64_t* launches)
{
    cudaStream_t s = ctx->stream;
    const int32_t kp = c.kp, kc = c.kc;
    const int64_t rows_pad = ceil_div(n_users, TC_ROWS) * TC_ROWS;
    MML_CUDA(cudaMemsetAsync(w.err[x].p, 0, sizeof(uint32_t), s));
    MML_CUDA(cudaMemsetAsync(w.out_i[x].p, 0, sizeof(int32_t) * (size_t)n_users * n_out, s));
    MML_CUDA(cudaMemsetAsync(w.out_s[x].p, 0, sizeof(float) * (size_t)n_users * n_out, s));
    const size_t rf = c.row_floats();
    if (rows_pad > n_users) MML_CUDA(cudaMemsetAsync(w.Ub.p + (size_t)n_users * rf, 0, sizeof(float) * (size_t)(rows_pad - n_users) * rf, s));
    tc_stage_rows_kernel<<<tc_grid((int64_t)n_users * 32), 256, 0, s>>>(d_U, n_model_users, k, w.users[x].p, 1, n_users, kp, w.Ub.p, w.unorm.p, w.row_ok.p, nullptr, nullptr, c.bf16 ? 1 : 0);
    MML_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    // ignore_items as candidate positions, ascending inside a row (the epilogue walks them with a cursor): always sorted on
    // the device by (row, position) -- half a millisecond per 5M entries, and no round trip to the host to ask whether the
    // caller's lists were already in order
    const bool have_ign = n_ign > 0;
    if (have_ign) {
        tc_ignore_pos_kernel<<<tc_grid(n_ign), 256, 0, s>>>(w.ign_ptr[x].p, n_users, w.ign_idx[x].p, n_ign, d_cand ? c.pos_of.p : nullptr,
                                                           n_model_items, c.n_cand, w.row_of.p);
        tc_copy_u32_kernel<<<tc_grid(n_ign), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(w.ign_idx[x].p), n_ign, w.key.p);
        MML_CUDA(cudaGetLastError());
        MML_TRY(radix_sort_pairs(w.key.p, w.row_of.p, w.t1.p, w.t2.p, n_ign, 31, s));                                              // by position
        MML_TRY(radix_sort_pairs(w.row_of.p, w.key.p, w.t1.p, w.t2.p, n_ign, bits_for((uint32_t)std::max(n_users - 1, 1)), s));   // by row, stable
        tc_copy_u32_kernel<<<tc_grid(n_ign), 256, 0, s>>>(w.key.p, n_ign, reinterpret_cast<uint32_t*>(w.ign_idx[x].p));
        MML_CUDA(cudaGetLastError());
        if (launches) *launches += 13;
    }
    CUtensorMap map_u;
    MML_TRY(make_panel_map(&map_u, w.Ub.p, rows_pad, kp, c.bf16));
    const int row_tiles = (int)(rows_pad / TC_ROWS);
    TcArgs a{};
    a.n_rows = n_users; a.kc = kc; a.m = n;
    a.bf16 = c.bf16 ? 1 : 0; a.epc = c.bf16 ? TC_KC_BF16 : TC_KC;
    { static const int dbg = [] { const char* e = getenv("MMLB200_TC_DBG"); return e ? atoi(e) : 0; }(); a.dbg = dbg; }
    a.row_ok = w.row_ok.p; a.ign_ptr = have_ign ? w.ign_ptr[x].p : nullptr; a.ign_pos = w.ign_idx[x].p;
    a.thr = w.thr.p; a.err = w.err[x].p;
    // pass 1: threshold from the candidate sample
    a.stride = c.stride; a.n_tiles = (int)(c.samp_pad / TC_N); a.tiles_per_split = a.n_tiles; a.splits = 1;
    a.stages = w.stages0; a.list_off = (2 * kc + w.stages0) * TC_CHUNK_BYTES;
    a.n_cols = c.n_samp; a.bad = c.has_bad ? c.bad_s.p : nullptr;
    score_select_kernel<0><<<dim3(row_tiles, 1), TC_THREADS, w.smem0, s>>>(map_u, c.map_s, a);
    MML_CUDA(cudaGetLastError());
    tc_threshold_kernel<<<(unsigned)ceil_div(n_users, 256), 256, 0, s>>>(w.thr.p, w.d2.p, w.unorm.p, c.vmax.p, n_users, c.bf16 ? TC_ERR_C_BF16 : TC_ERR_C);
    MML_CUDA(cudaGetLastError());
    // pass 2: collect every candidate that reaches it
    MML_CUDA(cudaMemsetAsync(w.buf_n.p, 0, sizeof(int32_t) * (size_t)n_users * w.splits * 2, s));
    a.stages = w.stages; a.list_off = 0;
    a.stride = 1; a.n_tiles = (int)(c.cand_pad / TC_N); a.tiles_per_split = w.tps; a.splits = w.splits;
    a.n_cols = c.n_cand; a.bad = c.has_bad ? c.bad_b.p : nullptr;
    a.buf_s = w.buf_s.p; a.buf_p = w.buf_p.p; a.buf_n = w.buf_n.p;
    score_select_kernel<1><<<dim3(row_tiles, w.splits), TC_THREADS, w.smem, s>>>(map_u, c.map_b, a);
    MML_CUDA(cudaGetLastError());
    FinArgs f{};
    f.U = d_U; f.V = d_V; f.k = k; f.users = w.users[x].p; f.cand = d_cand;
    f.buf_s = w.buf_s.p; f.buf_p = w.buf_p.p; f.buf_n = w.buf_n.p; f.thr = w.thr.p; f.d2 = w.d2.p; f.row_ok = w.row_ok.p;
    f.n_rows = n_users; f.splits = w.splits; f.n = n; f.n_out = n_out;
    f.out_items = w.out_i[x].p; f.out_scores = w.out_s[x].p; f.out_counts = w.out_c[x].p; f.redo = w.redo[x].p;
    const size_t per_warp = (size_t)w.splits * 2 * TC_CAP * 12;
    f.warps = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)96 * 1024 / per_warp));
    const size_t fsmem = per_warp * f.warps;
    MML_CUDA(cudaFuncSetAttribute((const void*)tc_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    tc_finalize_kernel<<<(unsigned)ceil_div(n_users, f.warps), f.warps * 32, fsmem, s>>>(f);
    MML_CUDA(cudaGetLastError());
    if (launches) *launches += 4;
    return MML_OK;
}

// MMLB200_TOPN_TIMES=1: host time stamps of one call (no synchronisation added), printed when the call returns
struct TcTimes {
    bool on; std::chrono::steady_clock::time_point t0; std::string log;
    TcTimes() { static const bool e = [] { const char* v = getenv("MMLB200_TOPN_TIMES"); return v && *v && *v != '0'; }(); on = e; t0 = std::chrono::steady_clock::now(); }
    void mark(const char* what) {
        if (!on) return;
        char buf[96];
        snprintf(buf, sizeof(buf), " %s@%.1f", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        log += buf;
    }
    ~TcTimes() { if (on) fprintf(stderr, "[mmlb200 topn host ms]%s\n", log.c_str()); }
};

static bool tc_trace() { static int t = -1; if (t < 0) { const char* e = getenv("MMLB200_TRACE"); t = (e && *e && *e != '0') ? 1 : 0; } return t == 1; }
struct TcPhase {
    cudaStream_t s; std::chrono::steady_clock::time_point t0;
    explicit TcPhase(cudaStream_t st) : s(st) { if (tc_trace()) { cudaStreamSynchronize(s); t0 = std::chrono::steady_clock::now(); } }
    void mark(const char* what) {
        if (!tc_trace()) return;
        cudaStreamSynchronize(s);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[mmlb200 topn] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// All user batches of one Recommend() call. users / ignore CSR / outputs: host (pageable or not). d_cand: device or NULL.
// Three streams: batch b + 1's inputs go up (copy_stream) and batch b - 1's results come down (out_stream) through
// page-locked staging while batch b's kernels run (stream); the host only waits for staging it is about to reuse and for
// results it is about to hand to the caller.
int32_t topn_tc_run(Ctx* ctx, const float* d_U, int32_t n_model_users, const float* d_V, int32_t n_model_items, int32_t k,
                    const int32_t* users, int64_t n_users, int32_t n, int32_t n_out, const int32_t* d_cand, int32_t n_cand,
                    bool has_invalid_cand, const int64_t* ignore_ptr, const int32_t* ignore_idx,
                    int32_t* out_items, float* out_scores, int32_t* out_counts, std::vector<int64_t>& redo_users, int64_t* launches)
{
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream, os = ctx->out_stream;
    TcPhase ph(s);
    TcTimes tt;
    TopnCache* cache = topn_cache(ctx);
    MML_CHECK(cache != nullptr, MML_ERR_ARG, "out of host memory");
    TcCandidates& c = cache->c;
    TcWork& w = cache->w;
    c.bf16 = topn_tc_filter_bf16();
    MML_TRY(tc_prepare_candidates(ctx, c, d_V, n_model_items, k, d_cand, n_cand, has_invalid_cand, n, launches));
    ph.mark("candidate panels");
    const int64_t n_ign = (ignore_ptr && ignore_idx) ? ignore_ptr[n_users] : 0;
    const int64_t B = 1 << 18;                       // users per pass (staging panel 128 MB, collect buffers 512 MB)
    int64_t max_ign = 0;
    if (n_ign > 0)
        for (int64_t b_lo = 0; b_lo < n_users; b_lo += B)
            max_ign = std::max(max_ign, ignore_ptr[std::min(b_lo + B, n_users)] - ignore_ptr[b_lo]);
    MML_TRY(tc_alloc_work(ctx, w, c, (int32_t)std::min<int64_t>(B, n_users), n_out, max_ign));
    ph.mark("workspace");
    tt.mark("prepared");
    // batch bookkeeping for the retire step
    struct Pending { int64_t b_lo; int32_t nb; bool live; } pend[2] = {{0, 0, false}, {0, 0, false}};
    auto retire = [&](int x) -> int32_t {
        if (!pend[x].live) return MML_OK;
        tt.mark("wait");
        MML_CUDA(cudaEventSynchronize(w.ev_out[x]));
        tt.mark("got");
        const int64_t b_lo = pend[x].b_lo; const int32_t nb = pend[x].nb;
        MML_CHECK(w.h_err[x].p[0] == 0, MML_ERR_CUDA, "topn: tcgen05 pipeline timed out");
        memcpy(out_items + (size_t)b_lo * n_out, w.h_out_i[x].p, sizeof(int32_t) * (size_t)nb * n_out);
        memcpy(out_scores + (size_t)b_lo * n_out, w.h_out_s[x].p, sizeof(float) * (size_t)nb * n_out);
        memcpy(out_counts + b_lo, w.h_out_c[x].p, sizeof(int32_t) * (size_t)nb);
        const uint8_t* rd = w.h_redo[x].p;
        for (int32_t t = 0; t < nb; t++) if (rd[t]) redo_users.push_back(b_lo + t);
        pend[x].live = false;
        tt.mark("retired");
        return MML_OK;
    };
    int bi = 0;
    for (int64_t b_lo = 0; b_lo < n_users; b_lo += B, bi++) {
        const int x = bi & 1;
        const int32_t nb = (int32_t)std::min<int64_t>(B, n_users - b_lo);
        // staging x was last read by the upload of batch bi - 2, device inputs x by its kernels: both finished once its
        // results were retired (previous iteration)
        memcpy(w.h_users[x].p, users + b_lo, sizeof(int32_t) * (size_t)nb);
        MML_CUDA(cudaMemcpyAsync(w.users[x].p, w.h_users[x].p, sizeof(int32_t) * (size_t)nb, cudaMemcpyHostToDevice, cs));
        int64_t nib = 0;
        if (n_ign > 0) {
            const int64_t i_lo = ignore_ptr[b_lo];
            nib = ignore_ptr[b_lo + nb] - i_lo;
            int64_t* pl = w.h_ign_ptr[x].p;
            for (int32_t t = 0; t <= nb; t++) pl[t] = ignore_ptr[b_lo + t] - i_lo;
            MML_CUDA(cudaMemcpyAsync(w.ign_ptr[x].p, pl, sizeof(int64_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, cs));
            if (nib > 0) {
                memcpy(w.h_ign_idx[x].p, ignore_idx + i_lo, sizeof(int32_t) * (size_t)nib);
                MML_CUDA(cudaMemcpyAsync(w.ign_idx[x].p, w.h_ign_idx[x].p, sizeof(int32_t) * (size_t)nib, cudaMemcpyHostToDevice, cs));
            }
        }
        MML_CUDA(cudaEventRecord(w.ev_in[x], cs));
        MML_CUDA(cudaStreamWaitEvent(s, w.ev_in[x], 0));
        ph.mark("batch H2D");
        tt.mark("staged");
        MML_TRY(topn_tc_batch(ctx, c, w, x, d_U, n_model_users, d_V, n_model_items, k, nb, n, n_out, d_cand, nib, launches));
        MML_CUDA(cudaEventRecord(w.ev_done[x], s));
        ph.mark("batch kernels");
        tt.mark("issued");
        MML_CUDA(cudaStreamWaitEvent(os, w.ev_done[x], 0));
        MML_CUDA(cudaMemcpyAsync(w.h_out_i[x].p, w.out_i[x].p, sizeof(int32_t) * (size_t)nb * n_out, cudaMemcpyDeviceToHost, os));
        MML_CUDA(cudaMemcpyAsync(w.h_out_s[x].p, w.out_s[x].p, sizeof(float) * (size_t)nb * n_out, cudaMemcpyDeviceToHost, os));
        MML_CUDA(cudaMemcpyAsync(w.h_out_c[x].p, w.out_c[x].p, sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, os));
        MML_CUDA(cudaMemcpyAsync(w.h_redo[x].p, w.redo[x].p, (size_t)nb, cudaMemcpyDeviceToHost, os));
        MML_CUDA(cudaMemcpyAsync(w.h_err[x].p, w.err[x].p, sizeof(uint32_t), cudaMemcpyDeviceToHost, os));
        MML_CUDA(cudaEventRecord(w.ev_out[x], os));
        pend[x] = {b_lo, nb, true};
        MML_TRY(retire(x ^ 1));      // the previous batch, while this one computes
        ph.mark("previous batch retired");
    }
    MML_TRY(retire(bi & 1));            // batches bi - 2 (already retired in the loop: a no-op) and bi - 1
    MML_TRY(retire((bi & 1) ^ 1));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

}  // namespace mml
