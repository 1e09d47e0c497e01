// sgd.cu -- the SGD epoch of MatrixFactorization / BiasedMatrixFactorization on sm_100a.
//
// Reference: RatingPrediction/BiasedMatrixFactorization.cs:161-325 (InitModel, Train, Iterate, the
// per-rating update, Predict), :496-552 (objective); RatingPrediction/MatrixFactorization.cs:99-259;
// MultiCore.cs:43-73 (the user x item block stratification the DSGD schedule mirrors).
//
// Schedule. The reference's DSGD mode cuts the rating matrix into g x g blocks by
// (user_perm[u] % g, item_perm[i] % g) and runs g sub-epochs of g mutually disjoint blocks. Here the
// same stratification is applied at two levels on one GPU:
//   level 1: G worker groups = CTAs.   CTA j owns user group j for the whole epoch; in sub-epoch
//            ("slot") s it holds item group (s + j) mod G, staged in shared memory.
//   level 2: W sub-groups  = warps.    Inside block (j, b) warp w owns user sub-group w; in step t
//            it works on item sub-group (t + w) mod W; __syncthreads() separates the steps.
// At any moment no two warps of the GPU touch the same user row or the same item row, so the
// parallel epoch equals a serial pass over the ratings in the order (slot, step, j, w, entry) --
// the order mml_sgd_schedule_dump returns and the tests replay through the CPU oracle.
//
// Data layout in HBM. Factor rows are renumbered group-major (all rows of one group are contiguous)
// and padded with zeros to kp = 32 * kpl floats, so a warp moves a row with one 128-bit (64/32-bit)
// access per lane and an item group is one contiguous region (bulk-copied to shared memory).
// Ratings are stored as entries (internal user row, internal item row, value) sorted by the
// consumption order above; sub_ptr[] delimits the (j, slot, w, step) sub-blocks.
#include "sgd.cuh"
#include <algorithm>
#include <cmath>
#include <new>
#include <numeric>
#include <cooperative_groups.h>

namespace mml {

// =================================================================================================
// host: group maps
// =================================================================================================
// Deals ids to n_blocks * G * W groups and numbers them group-major.
//   level 0 (GPU block)  : perm[id] % R          (the reference rule, MultiCore.cs:64)
//   level 1/2, PERM_MOD  : (perm[id] / R) % (G*W) -> g = x % G, w = x / G
//   level 1/2, BALANCED  : ids of a block sorted by rating count (desc) and dealt boustrophedon
static void build_group_map(GroupMap& m, int32_t n_ext, const uint32_t* counts, const int32_t* perm,
                            int32_t R, int32_t only_block /* -1 = all blocks */, int32_t G, int32_t W, int32_t rule)
{
    m.n_ext = n_ext;
    m.n_blocks = only_block >= 0 ? 1 : R;
    const int32_t T = G * W;
    m.grp.assign(n_ext, -1);
    std::vector<std::vector<int32_t>> by_block(m.n_blocks);
    for (int32_t id = 0; id < n_ext; id++) {
        const int32_t pid = perm ? perm[id] : id;
        const int32_t blk = pid % R;
        if (only_block >= 0 && blk != only_block) continue;
        const int32_t bi = only_block >= 0 ? 0 : blk;
        if (rule == MML_GROUPS_PERM_MOD) {
            const int32_t x = (pid / R) % T;
            m.grp[id] = (bi * G + (x % G)) * W + (x / G);
        } else {
            by_block[bi].push_back(id);
        }
    }
    if (rule != MML_GROUPS_PERM_MOD) {
        for (int32_t bi = 0; bi < m.n_blocks; bi++) {
            auto& ids = by_block[bi];
            std::stable_sort(ids.begin(), ids.end(), [&](int32_t a, int32_t b) { return counts[a] > counts[b]; });
            for (size_t pos = 0; pos < ids.size(); pos++) {
                const size_t round = pos / T, r = pos % T;
                const int32_t x = (int32_t)((round & 1) ? (T - 1 - r) : r);
                m.grp[ids[pos]] = (bi * G + (x % G)) * W + (x / G);
            }
        }
    }
    // group-major numbering (counting sort, ascending id inside a group)
    const int32_t n_grp = m.n_blocks * T;
    m.grp_ptr.assign((size_t)n_grp + 1, 0);
    for (int32_t id = 0; id < n_ext; id++) if (m.grp[id] >= 0) m.grp_ptr[m.grp[id] + 1]++;
    for (int32_t g = 0; g < n_grp; g++) m.grp_ptr[g + 1] += m.grp_ptr[g];
    m.n_int = m.grp_ptr[n_grp];
    m.to_int.assign(n_ext, -1);
    m.to_ext.assign(std::max(m.n_int, 1), 0);
    std::vector<int32_t> cursor(m.grp_ptr.begin(), m.grp_ptr.end() - 1);
    for (int32_t id = 0; id < n_ext; id++) {
        if (m.grp[id] < 0) continue;
        const int32_t r = cursor[m.grp[id]]++;
        m.to_int[id] = r;
        m.to_ext[r] = id;
    }
}

static int32_t upload_group_map(GroupMap& m, cudaStream_t s)
{
    MML_TRY(m.d_grp.alloc(m.n_ext)); MML_TRY(m.d_to_int.alloc(m.n_ext)); MML_TRY(m.d_to_ext.alloc(m.n_int));
    if (m.n_ext > 0) {
        MML_CUDA(cudaMemcpyAsync(m.d_grp.p, m.grp.data(), sizeof(int32_t) * m.n_ext, cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaMemcpyAsync(m.d_to_int.p, m.to_int.data(), sizeof(int32_t) * m.n_ext, cudaMemcpyHostToDevice, s));
    }
    if (m.n_int > 0)
        MML_CUDA(cudaMemcpyAsync(m.d_to_ext.p, m.to_ext.data(), sizeof(int32_t) * m.n_int, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

// =================================================================================================
// device: strata build
// =================================================================================================
static inline int grid_n(int64_t n, int threads = 256)
{
    return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n, threads), 1), 148 * 16);
}

// key = ((((B*G + j)*G + slot)*W + w)*W + step ; bad[0] counts ratings whose user/item is not mapped
__global__ void strata_key_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t n,
                                  const int32_t* __restrict__ user_grp, const int32_t* __restrict__ item_grp,
                                  int32_t G, int32_t W, uint32_t* __restrict__ key, uint32_t* __restrict__ bad)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const int32_t ug = user_grp[users[t]], ig = item_grp[items[t]];
        if (ug < 0 || ig < 0) { atomicAdd(bad, 1u); key[t] = 0; continue; }
        const int32_t j = ug / W, w = ug % W;
        const int32_t Bb = ig / W, c = ig % W;
        const int32_t B = Bb / G, b = Bb % G;
        int32_t slot = b - j; if (slot < 0) slot += G;
        int32_t step = c - w; if (step < 0) step += W;
        key[t] = (uint32_t)((((B * G + j) * G + slot) * W + w) * W + step);
    }
}

__global__ void strata_entries_kernel(const uint32_t* __restrict__ order, const int32_t* __restrict__ users,
                                      const int32_t* __restrict__ items, const float* __restrict__ values, int64_t n,
                                      const int32_t* __restrict__ user_int, const int32_t* __restrict__ item_int,
                                      int32_t* __restrict__ ent_u, int32_t* __restrict__ ent_i,
                                      float* __restrict__ ent_v, int32_t* __restrict__ ent_idx)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const uint32_t src = order[t];
        ent_u[t] = user_int[users[src]];
        ent_i[t] = item_int[items[src]];
        ent_v[t] = values[src];
        ent_idx[t] = (int32_t)src;
    }
}

static int32_t build_strata(Sgd& m)
{
    Ratings& r = *m.ratings;
    cudaStream_t s = m.ctx->stream;
    const int64_t n = r.n;
    m.n_sub = (int64_t)m.R * m.G * m.G * m.W * m.W;
    MML_CHECK(m.n_sub < ((int64_t)1 << 31), MML_ERR_ARG, "strata: %lld sub-blocks is too many (G=%d W=%d)",
              (long long)m.n_sub, m.G, m.W);
    DevBuf<uint32_t> key, vals, ktmp, vtmp, cnt, bad;
    MML_TRY(key.alloc(n)); MML_TRY(vals.alloc(n)); MML_TRY(ktmp.alloc(n)); MML_TRY(vtmp.alloc(n));
    MML_TRY(cnt.alloc(m.n_sub)); MML_TRY(bad.alloc(1));
    MML_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(uint32_t), s));
    MML_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
    strata_key_kernel<<<grid_n(n), 256, 0, s>>>(r.users.p, r.items.p, n, m.users.d_grp.p, m.items.d_grp.p,
                                                m.G, m.W, key.p, bad.p);
    MML_CUDA(cudaGetLastError());
    uint32_t h_bad = 0;
    MML_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    MML_CHECK(h_bad == 0, MML_ERR_ARG,
              "strata: %u ratings belong to users of another GPU block (pass each rank the ratings of its own users)", h_bad);
    MML_TRY(histogram_i32((const int32_t*)key.p, n, cnt.p, s));
    MML_TRY(m.sub_ptr.alloc((size_t)m.n_sub + 1));
    MML_TRY(exclusive_scan_u32(cnt.p, m.sub_ptr.p, m.n_sub, s));
    MML_TRY(iota_u32(vals.p, n, s));
    MML_TRY(radix_sort_pairs(key.p, vals.p, ktmp.p, vtmp.p, n, bits_for((uint32_t)(m.n_sub - 1)), s));
    MML_TRY(m.ent_u.alloc(n)); MML_TRY(m.ent_i.alloc(n)); MML_TRY(m.ent_v.alloc(n)); MML_TRY(m.ent_idx.alloc(n));
    strata_entries_kernel<<<grid_n(n), 256, 0, s>>>(vals.p, r.users.p, r.items.p, r.values.p, n,
                                                    m.users.d_to_int.p, m.items.d_to_int.p,
                                                    m.ent_u.p, m.ent_i.p, m.ent_v.p, m.ent_idx.p);
    MML_CUDA(cudaGetLastError());
    MML_CUDA(cudaStreamSynchronize(s));
    m.launches += 8;
    return MML_OK;
}

// =================================================================================================
// device: the per-rating update
// =================================================================================================
struct SgdArgs {
    float* P; float* Q; float* bu; float* bi;
    const int32_t* ent_u; const int32_t* ent_i; const float* ent_v;
    const uint32_t* sub_ptr;      // offset to the GPU-level item block B
    const int32_t* item_ptr;      // offset to B: [G + 1] internal item row range per item group
    const float* regw_u; const float* regw_i;   // frequency regularisation weights or NULL
    uint32_t* flags;              // persistent kernel: progress counter per CTA
    const int32_t* seq;           // persistent kernel: sub-epoch sequence [G] (device)
    uint32_t epoch_base;
    int32_t G, W;
    float lr, gb, minr, range, reg_u, reg_i, blr, breg;
    int32_t loss;
};

// A factor row as seen by one lane: kpl floats, 128-bit accesses where the row is long enough.
template <int KPL> struct Row;
template <> struct Row<1> {
    static __device__ __forceinline__ void load(float (&r)[1], const float* row, int lane) { r[0] = row[lane]; }
    static __device__ __forceinline__ void store(const float (&r)[1], float* row, int lane) { row[lane] = r[0]; }
};
template <> struct Row<2> {
    static __device__ __forceinline__ void load(float (&r)[2], const float* row, int lane)
    { const float2 v = reinterpret_cast<const float2*>(row)[lane]; r[0] = v.x; r[1] = v.y; }
    static __device__ __forceinline__ void store(const float (&r)[2], float* row, int lane)
    { reinterpret_cast<float2*>(row)[lane] = make_float2(r[0], r[1]); }
};
template <> struct Row<4> {
    static __device__ __forceinline__ void load(float (&r)[4], const float* row, int lane)
    { const float4 v = reinterpret_cast<const float4*>(row)[lane]; r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }
    static __device__ __forceinline__ void store(const float (&r)[4], float* row, int lane)
    { reinterpret_cast<float4*>(row)[lane] = make_float4(r[0], r[1], r[2], r[3]); }
};
template <> struct Row<8> {
    static __device__ __forceinline__ void load(float (&r)[8], const float* row, int lane)
    {
        const float4 a = reinterpret_cast<const float4*>(row)[lane];
        const float4 b = reinterpret_cast<const float4*>(row)[32 + lane];
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
    }
    static __device__ __forceinline__ void store(const float (&r)[8], float* row, int lane)
    {
        reinterpret_cast<float4*>(row)[lane] = make_float4(r[0], r[1], r[2], r[3]);
        reinterpret_cast<float4*>(row)[32 + lane] = make_float4(r[4], r[5], r[6], r[7]);
    }
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// BiasedMatrixFactorization.cs:264-310 (BIASED) / MatrixFactorization.cs:166-196 for one rating, fp32.
// p and q are updated from their pre-update values; the biases before the factors.
template <int KPL, bool BIASED>
__device__ __forceinline__ void sgd_update(const SgdArgs& a, float (&p)[KPL], float (&q)[KPL],
                                           float& bu, float& bi, float v, float regu, float regi)
{
    float dot = 0.f;
#pragma unroll
    for (int f = 0; f < KPL; f++) dot = fmaf(p[f], q[f], dot);
    dot = warp_sum(dot);
    float gc;
    if (BIASED) {
        const float score = ((a.gb + bu) + bi) + dot;
        const float sig = 1.f / (1.f + expf(-score));
        const float err = v - (a.minr + sig * a.range);
        if (a.loss == MML_LOSS_RMSE) gc = err * sig * (1.f - sig) * a.range;
        else if (a.loss == MML_LOSS_MAE) gc = (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * sig * (1.f - sig) * a.range;
        else gc = err;
        const float step = a.blr * a.lr;
        bu += step * (gc - a.breg * regu * bu);
        bi += step * (gc - a.breg * regi * bi);
    } else {
        gc = v - (a.gb + dot);
    }
#pragma unroll
    for (int f = 0; f < KPL; f++) {
        const float pf = p[f], qf = q[f];
        p[f] = pf + a.lr * (gc * qf - regu * pf);
        q[f] = qf + a.lr * (gc * pf - regi * qf);
    }
}

// =================================================================================================
// device: DSGD kernels
// =================================================================================================
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cg_f(const float* p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// Work of CTA j on block (j, b): stage item group b, W conflict-free steps, write the group back.
// STAGE = item group lives in shared memory while the CTA holds it; otherwise item rows are
// updated in global memory (groups too large for shared memory, e.g. very small G).
template <int KPL, bool BIASED, bool STAGE>
__device__ __forceinline__ void sgd_block(const SgdArgs& a, const int j, const int slot, float* smem)
{
    constexpr int KP = 32 * KPL;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int b = slot + j; if (b >= a.G) b -= a.G;
    const int i_lo = a.item_ptr[b], i_hi = a.item_ptr[b + 1];
    const int n_it = i_hi - i_lo;
    float* sQ = smem;
    float* sB = smem + (size_t)n_it * KP;
    if (STAGE) {
        const float4* src = reinterpret_cast<const float4*>(a.Q + (size_t)i_lo * KP);
        float4* dst = reinterpret_cast<float4*>(sQ);
        const int n4 = n_it * (KP / 4);
        for (int t = threadIdx.x; t < n4; t += blockDim.x) dst[t] = ld_cg_f4(src + t);
        if (BIASED) for (int t = threadIdx.x; t < n_it; t += blockDim.x) sB[t] = ld_cg_f(a.bi + i_lo + t);
        __syncthreads();
    }
    const uint32_t* sp = a.sub_ptr + ((size_t)(j * a.G + slot) * a.W + w) * a.W;

    int cur_u = -1;
    float p[KPL];
    float bu_v = 0.f, regu = a.reg_u;
    for (int step = 0; step < a.W; step++) {
        const uint32_t beg = sp[step], end = sp[step + 1];
        for (uint32_t base = beg; base < end; base += 32) {
            const int cnt = min(32u, end - base);
            int mu = 0, mi = 0; float mv = 0.f;
            if (lane < cnt) { mu = a.ent_u[base + lane]; mi = a.ent_i[base + lane]; mv = a.ent_v[base + lane]; }
            for (int e = 0; e < cnt; e++) {
                const int u = __shfl_sync(0xffffffffu, mu, e);
                const int i = __shfl_sync(0xffffffffu, mi, e);
                const float v = __shfl_sync(0xffffffffu, mv, e);
                if (u != cur_u) {   // warp-uniform; the user row is owned by this warp for the whole block
                    if (cur_u >= 0) {
                        Row<KPL>::store(p, a.P + (size_t)cur_u * KP, lane);
                        if (BIASED && lane == 0) a.bu[cur_u] = bu_v;
                    }
                    Row<KPL>::load(p, a.P + (size_t)u * KP, lane);
                    if (BIASED) bu_v = a.bu[u];
                    if (a.regw_u) regu = a.regw_u[u];
                    cur_u = u;
                }
                float* qrow = STAGE ? (sQ + (size_t)(i - i_lo) * KP) : (a.Q + (size_t)i * KP);
                float* bip = STAGE ? (sB + (i - i_lo)) : (a.bi + i);
                float q[KPL];
                Row<KPL>::load(q, qrow, lane);
                float bi_v = BIASED ? *bip : 0.f;
                const float regi = a.regw_i ? a.regw_i[i] : a.reg_i;
                sgd_update<KPL, BIASED>(a, p, q, bu_v, bi_v, v, regu, regi);
                Row<KPL>::store(q, qrow, lane);
                if (BIASED && lane == 0) *bip = bi_v;
                __syncwarp();
            }
        }
        __syncthreads();
    }
    if (cur_u >= 0) {
        Row<KPL>::store(p, a.P + (size_t)cur_u * KP, lane);
        if (BIASED && lane == 0) a.bu[cur_u] = bu_v;
    }
    if (STAGE) {
        // the last __syncthreads() above made every warp's shared-memory updates visible
        float4* dst = reinterpret_cast<float4*>(a.Q + (size_t)i_lo * KP);
        const float4* src = reinterpret_cast<const float4*>(sQ);
        const int n4 = n_it * (KP / 4);
        for (int t = threadIdx.x; t < n4; t += blockDim.x) dst[t] = src[t];
        if (BIASED) for (int t = threadIdx.x; t < n_it; t += blockDim.x) a.bi[i_lo + t] = sB[t];
    }
}

// One launch per sub-epoch: grid = G CTAs, block = W warps.
template <int KPL, bool BIASED, bool STAGE>
__global__ void __launch_bounds__(1024) sgd_slot_kernel(const SgdArgs a, const int slot)
{
    extern __shared__ float4 smem4[];
    sgd_block<KPL, BIASED, STAGE>(a, blockIdx.x, slot, reinterpret_cast<float*>(smem4));
}

// One cooperative launch per epoch: CTA j walks the sub-epoch sequence; before it takes item group b
// it waits (acquire) until the CTA that held b in the previous sub-epoch has published it (release).
// All G CTAs are co-resident (cooperative launch), so the waits cannot deadlock.
template <int KPL, bool BIASED, bool STAGE>
__global__ void __launch_bounds__(1024) sgd_epoch_kernel(const SgdArgs a)
{
    extern __shared__ float4 smem4[];
    const int j = blockIdx.x;
    for (int t = 0; t < a.G; t++) {
        const int slot = a.seq[t];
        if (t > 0) {
            // item group b = (slot + j) % G was held in sub-epoch t-1 by CTA jp with (seq[t-1] + jp) % G == b
            int b = slot + j; if (b >= a.G) b -= a.G;
            int jp = b - a.seq[t - 1]; if (jp < 0) jp += a.G;
            if (threadIdx.x == 0) {
                const uint32_t want = a.epoch_base + (uint32_t)t;
                while ((int32_t)(ld_acquire_u32(a.flags + jp) - want) < 0) __nanosleep(20);
            }
            __syncthreads();
        }
        sgd_block<KPL, BIASED, STAGE>(a, j, slot, reinterpret_cast<float*>(smem4));
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_u32(a.flags + j, a.epoch_base + (uint32_t)t + 1u);
    }
}

// =================================================================================================
// device: reference-order serial pass (MaxThreads = 1 semantics, exact mixed precision)
// =================================================================================================
// One warp walks `indices` in order. Arithmetic follows the reference line by line: fp32 sequential
// dot (mul then add), double score / sigmoid / error, float gradient_common, double deltas,
// `+= (float)` increments (DataType/MatrixExtensions.cs:76-79, 224-241).
template <bool BIASED>
__global__ void sgd_serial_kernel(const SgdArgs a, const int32_t* __restrict__ indices, int64_t n_idx,
                                  const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                                  const float* __restrict__ values,
                                  const int32_t* __restrict__ user_int, const int32_t* __restrict__ item_int,
                                  int32_t k, int32_t kp, int update_user, int update_item)
{
    const int lane = threadIdx.x;
    const int nslot = kp / 32;
    for (int64_t t = 0; t < n_idx; t++) {
        const int32_t idx = indices[t];
        const int32_t u = user_int[users[idx]], i = item_int[items[idx]];
        const float r = values[idx];
        volatile float* prow = a.P + (size_t)u * kp;
        volatile float* qrow = a.Q + (size_t)i * kp;
        float pv[8], qv[8], prod[8];
        for (int s = 0; s < nslot; s++) {
            pv[s] = prow[s * 32 + lane]; qv[s] = qrow[s * 32 + lane];
            prod[s] = __fmul_rn(pv[s], qv[s]);
        }
        float dot = 0.f;
        for (int f = 0; f < k; f++) dot = __fadd_rn(dot, __shfl_sync(0xffffffffu, prod[f >> 5], f & 31));
        float gc, regu = a.reg_u, regi = a.reg_i;
        if (a.regw_u) regu = a.regw_u[u];
        if (a.regw_i) regi = a.regw_i[i];
        if (BIASED) {
            volatile float* pbu = a.bu + u; volatile float* pbi = a.bi + i;
            const float bu = *pbu, bi = *pbi;
            const double score = (double)__fadd_rn(__fadd_rn(__fadd_rn(a.gb, bu), bi), dot);
            const double sig = 1.0 / (1.0 + exp(-score));
            const double pred = (double)a.minr + sig * (double)a.range;
            const double err = (double)r - pred;
            if (a.loss == MML_LOSS_RMSE) gc = (float)(err * sig * (1.0 - sig) * (double)a.range);
            else if (a.loss == MML_LOSS_MAE) gc = (float)((err > 0 ? 1.0 : (err < 0 ? -1.0 : 0.0)) * sig * (1.0 - sig) * (double)a.range);
            else gc = (float)err;
            __syncwarp();
            if (lane == 0) {
                if (update_user) *pbu = __fadd_rn(bu, __fmul_rn(__fmul_rn(a.blr, a.lr), __fsub_rn(gc, __fmul_rn(__fmul_rn(a.breg, regu), bu))));
                if (update_item) *pbi = __fadd_rn(bi, __fmul_rn(__fmul_rn(a.blr, a.lr), __fsub_rn(gc, __fmul_rn(__fmul_rn(a.breg, regi), bi))));
            }
        } else {
            gc = __fsub_rn(r, __fadd_rn(a.gb, dot));
        }
        for (int s = 0; s < nslot; s++) {
            const double uf = pv[s], vf = qv[s];
            if (BIASED) {
                if (update_user) prow[s * 32 + lane] = __fadd_rn(pv[s], (float)((double)a.lr * ((double)gc * vf - (double)regu * uf)));
                if (update_item) qrow[s * 32 + lane] = __fadd_rn(qv[s], (float)((double)a.lr * ((double)gc * uf - (double)regi * vf)));
            } else {
                // MatrixFactorization.cs:181-191: err * i_f and Regularization * u_f are float products
                if (update_user) prow[s * 32 + lane] = __fadd_rn(pv[s], (float)((double)a.lr * (double)__fsub_rn(__fmul_rn(gc, qv[s]), __fmul_rn(regu, pv[s]))));
                if (update_item) qrow[s * 32 + lane] = __fadd_rn(qv[s], (float)((double)a.lr * (double)__fsub_rn(__fmul_rn(gc, pv[s]), __fmul_rn(regi, qv[s]))));
            }
        }
        __syncwarp();
    }
}

// =================================================================================================
// device: model in/out, init, predict, evaluate, objective
// =================================================================================================
// internal[r][f] = (f < k && count[ext] > 0) ? external[ext][f] : 0      (one warp per row)
__global__ void rows_in_kernel(const float* __restrict__ ext_rows, const int32_t* __restrict__ to_ext,
                               const uint32_t* __restrict__ counts, int32_t n_int, int32_t k, int32_t kp,
                               float* __restrict__ int_rows)
{
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n_int; r += stride) {
        const int32_t e = to_ext[r];
        const bool keep = counts[e] > 0;
        for (int f = lane; f < kp; f += 32)
            int_rows[r * kp + f] = (keep && f < k) ? ext_rows[(int64_t)e * k + f] : 0.f;
    }
}

__global__ void rows_out_kernel(const float* __restrict__ int_rows, const int32_t* __restrict__ to_ext,
                                int32_t n_int, int32_t k, int32_t kp, float* __restrict__ ext_rows)
{
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n_int; r += stride) {
        const int32_t e = to_ext[r];
        for (int f = lane; f < k; f += 32) ext_rows[(int64_t)e * k + f] = int_rows[r * kp + f];
    }
}

__global__ void vec_in_kernel(const float* __restrict__ ext, const int32_t* __restrict__ to_ext, int32_t n_int,
                              float* __restrict__ in)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; r < n_int; r += stride) in[r] = ext[to_ext[r]];
}

__global__ void vec_out_kernel(const float* __restrict__ in, const int32_t* __restrict__ to_ext, int32_t n_int,
                               float* __restrict__ ext)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; r < n_int; r += stride) ext[to_ext[r]] = in[r];
}

// regw[r] = (float)(reg / sqrt(count))  (BiasedMatrixFactorization.cs:281-282)
__global__ void regw_kernel(const int32_t* __restrict__ to_ext, const uint32_t* __restrict__ counts, int32_t n_int,
                            float reg, float* __restrict__ regw)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; r < n_int; r += stride) regw[r] = (float)((double)reg / sqrt((double)counts[to_ext[r]]));
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Counter-based N(mean, stddev): value (ext row e, factor f) depends only on (seed, stream, e, f), so the
// model is the same for every group layout. Rows of entities without ratings are zero.
__global__ void init_rows_kernel(const int32_t* __restrict__ to_ext, const uint32_t* __restrict__ counts,
                                 int32_t n_int, int32_t k, int32_t kp, uint64_t seed, uint64_t stream,
                                 float mean, float stddev, float* __restrict__ int_rows)
{
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n_int; r += stride) {
        const int32_t e = to_ext[r];
        const bool keep = counts[e] > 0;
        for (int f = lane; f < kp; f += 32) {
            float val = 0.f;
            if (keep && f < k) {
                const uint64_t h = splitmix64(splitmix64(seed ^ (stream << 56)) + (uint64_t)e * 1024ull + (uint64_t)f);
                const float u1 = ((float)(uint32_t)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
                const float u2 = ((float)(uint32_t)((h >> 8) & 0xFFFFFFu) + 0.5f) * (1.0f / 16777216.0f);
                val = mean + stddev * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
            }
            int_rows[r * kp + f] = val;
        }
    }
}

struct PredArgs {
    const float* P; const float* Q; const float* bu; const float* bi;
    const int32_t* user_int; const int32_t* item_int;
    int32_t n_users_ext, n_items_ext, kp, biased;
    float gb, minr, maxr, range;
};

// BiasedMatrixFactorization.cs:313-325 / MatrixFactorization.cs:205-217,251-259 -- one warp per pair
__device__ __forceinline__ float predict_pair(const PredArgs& a, int32_t u, int32_t i, int lane)
{
    const bool ku = u >= 0 && u < a.n_users_ext && a.user_int[u] >= 0;
    const bool ki = i >= 0 && i < a.n_items_ext && a.item_int[i] >= 0;
    float dot = 0.f;
    if (ku && ki) {
        const float* prow = a.P + (size_t)a.user_int[u] * a.kp;
        const float* qrow = a.Q + (size_t)a.item_int[i] * a.kp;
        for (int f = lane; f < a.kp; f += 32) dot = fmaf(prow[f], qrow[f], dot);
        dot = warp_sum(dot);
    }
    if (a.biased) {
        double score = a.gb;
        if (ku) score += a.bu[a.user_int[u]];
        if (ki) score += a.bi[a.item_int[i]];
        if (ku && ki) score += dot;
        return (float)((double)a.minr + (1.0 / (1.0 + exp(-score))) * (double)a.range);
    }
    if (!ku || !ki) return a.gb;
    float res = a.gb + dot;
    if (res > a.maxr) res = a.maxr;
    if (res < a.minr) res = a.minr;
    return res;
}

__global__ void predict_kernel(const PredArgs a, const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                               int64_t n, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; t < n; t += stride) {
        const float pr = predict_pair(a, users[t], items[t], lane);
        if (lane == 0) out[t] = pr;
    }
}

// Eval/Ratings.cs:96-162. part[blk*4 + {0,1,2,3}] = sum err^2, sum |err|, sum CBD, sum objective loss
constexpr int EV_THREADS = 256;
__global__ void evaluate_kernel(const PredArgs a, const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                                const float* __restrict__ values, int64_t n, int32_t loss, double* __restrict__ part)
{
    __shared__ double sh[EV_THREADS / 32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; t < n; t += stride) {
        const float pr = predict_pair(a, users[t], items[t], lane);
        const float r = values[t];
        const float err = __fsub_rn(pr, r);
        s0 += (double)__fmul_rn(err, err);
        s1 += (double)fabsf(err);
        double pn = ((double)pr - (double)a.minr) / ((double)a.maxr - (double)a.minr);
        const double an = ((double)r - (double)a.minr) / ((double)a.maxr - (double)a.minr);
        double pc = pn < 0.01 ? 0.01 : (pn > 0.99 ? 0.99 : pn);
        s2 += -(an * log10(pc) + (1.0 - an) * log10(1.0 - pc));
        if (loss == MML_LOSS_MAE) s3 += (double)fabsf(err);
        else if (loss == MML_LOSS_RMSE) { const double d = (double)err; s3 += d * d; }
        else {
            double pl = pn < 0.0 ? 0.0 : (pn > 1.0 ? 1.0 : pn);
            s3 -= an * log(pl);
            s3 -= (1.0 - an) * log(1.0 - pl);
        }
    }
    if (lane == 0) { sh[warp][0] = s0; sh[warp][1] = s1; sh[warp][2] = s2; sh[warp][3] = s3; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double acc = 0;
        for (int wv = 0; wv < EV_THREADS / 32; wv++) acc += sh[wv][threadIdx.x];
        part[(int64_t)blockIdx.x * 4 + threadIdx.x] = acc;
    }
}

// BiasedMatrixFactorization.cs:518-549: sum over rows of weight(count) * (|row|^2 + bias_reg * bias^2).
// mode 0: weight = count * reg ; mode 1 (frequency regularisation): weight = reg / sqrt(count) [count > 0]
__global__ void regterm_kernel(const float* __restrict__ rows, const float* __restrict__ bias,
                               const int32_t* __restrict__ to_ext, const uint32_t* __restrict__ counts,
                               int32_t n_int, int32_t kp, float reg, float bias_reg, int mode, double* __restrict__ part)
{
    __shared__ double sh[EV_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n_int; r += stride) {
        const uint32_t c = counts[to_ext[r]];
        if (c == 0) continue;
        double nrm = 0;
        for (int f = lane; f < kp; f += 32) { const double x = rows[r * kp + f]; nrm += x * x; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, d);
        const double b = bias ? (double)bias[r] : 0.0;
        const double wgt = mode == 0 ? (double)c * (double)reg : (double)reg / sqrt((double)c);
        acc += wgt * (nrm + (double)bias_reg * b * b);
    }
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int wv = 0; wv < EV_THREADS / 32; wv++) t += sh[wv];
        part[blockIdx.x] = t;
    }
}

// =================================================================================================
// host: launch helpers
// =================================================================================================
static SgdArgs make_args(Sgd& m, int32_t B)
{
    SgdArgs a{};
    a.P = m.P.p; a.Q = m.Q.p; a.bu = m.bu.p; a.bi = m.bi.p;
    a.ent_u = m.ent_u.p; a.ent_i = m.ent_i.p; a.ent_v = m.ent_v.p;
    a.sub_ptr = m.sub_ptr.p ? m.sub_ptr.p + (size_t)B * m.G * m.G * m.W * m.W : nullptr;
    a.item_ptr = m.d_item_ptr.p ? m.d_item_ptr.p + (size_t)B * m.G : nullptr;
    a.regw_u = m.p.frequency_regularization ? m.regw_u.p : nullptr;
    a.regw_i = m.p.frequency_regularization ? m.regw_i.p : nullptr;
    a.flags = m.flags.p; a.seq = nullptr; a.epoch_base = m.epoch_base;
    a.G = m.G; a.W = m.W;
    a.lr = m.lr; a.gb = m.global_bias; a.minr = m.min_rating; a.range = m.range;
    if (m.p.biased) { a.reg_u = m.p.reg_u; a.reg_i = m.p.reg_i; }
    else { a.reg_u = m.p.regularization; a.reg_i = m.p.regularization; }
    a.blr = m.p.bias_learn_rate; a.breg = m.p.bias_reg; a.loss = m.p.loss;
    return a;
}

typedef void (*slot_fn_t)(const SgdArgs, const int);
typedef void (*epoch_fn_t)(const SgdArgs);

template <int KPL>
static void pick_kernels(bool biased, bool stage, slot_fn_t* sf, epoch_fn_t* ef)
{
    if (biased) {
        if (stage) { *sf = sgd_slot_kernel<KPL, true, true>; *ef = sgd_epoch_kernel<KPL, true, true>; }
        else { *sf = sgd_slot_kernel<KPL, true, false>; *ef = sgd_epoch_kernel<KPL, true, false>; }
    } else {
        if (stage) { *sf = sgd_slot_kernel<KPL, false, true>; *ef = sgd_epoch_kernel<KPL, false, true>; }
        else { *sf = sgd_slot_kernel<KPL, false, false>; *ef = sgd_epoch_kernel<KPL, false, false>; }
    }
}

static int32_t get_kernels(Sgd& m, slot_fn_t* sf, epoch_fn_t* ef)
{
    const bool stage = m.stage_bytes > 0;
    switch (m.kpl) {
        case 1: pick_kernels<1>(m.p.biased != 0, stage, sf, ef); break;
        case 2: pick_kernels<2>(m.p.biased != 0, stage, sf, ef); break;
        case 4: pick_kernels<4>(m.p.biased != 0, stage, sf, ef); break;
        case 8: pick_kernels<8>(m.p.biased != 0, stage, sf, ef); break;
        default: set_error("unsupported num_factors"); return MML_ERR_UNSUPPORTED;
    }
    if (stage) {
        MML_CUDA(cudaFuncSetAttribute((const void*)*sf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.stage_bytes));
        MML_CUDA(cudaFuncSetAttribute((const void*)*ef, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.stage_bytes));
    }
    return MML_OK;
}

static PredArgs make_pred_args(Sgd& m)
{
    PredArgs a{};
    a.P = m.P.p; a.Q = m.Q.p; a.bu = m.bu.p; a.bi = m.bi.p;
    a.user_int = m.users.d_to_int.p; a.item_int = m.items.d_to_int.p;
    a.n_users_ext = m.users.n_ext; a.n_items_ext = m.items.n_ext; a.kp = m.kp; a.biased = m.p.biased;
    a.gb = m.global_bias; a.minr = m.min_rating; a.maxr = m.max_rating; a.range = m.range;
    return a;
}

// sums of evaluate_kernel over device-resident (users, items, values)
static int32_t evaluate_device(Sgd& m, const int32_t* d_u, const int32_t* d_i, const float* d_v, int64_t n, double* sums4)
{
    cudaStream_t s = m.ctx->stream;
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n * 32, EV_THREADS * 4), 1), 148 * 8);
    DevBuf<double> part;
    MML_TRY(part.alloc((size_t)blocks * 4));
    evaluate_kernel<<<blocks, EV_THREADS, 0, s>>>(make_pred_args(m), d_u, d_i, d_v, n, m.p.loss, part.p);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    std::vector<double> h((size_t)blocks * 4);
    MML_CUDA(cudaMemcpyAsync(h.data(), part.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < 4; c++) sums4[c] = 0;
    for (int b = 0; b < blocks; b++) for (int c = 0; c < 4; c++) sums4[c] += h[(size_t)b * 4 + c];
    return MML_OK;
}

static void sums_to_measures(const Sgd& m, const double* sums4, int64_t n, float* out4)
{
    // Eval/Ratings.cs:129-138
    const double mae = sums4[1] / (double)n, rmse = std::sqrt(sums4[0] / (double)n), cbd = sums4[2] / (double)n;
    out4[0] = (float)rmse; out4[1] = (float)mae;
    out4[2] = (float)mae / (m.max_rating - m.min_rating);
    out4[3] = (float)cbd;
}

static int32_t regterm(Sgd& m, bool user_side, double* out)
{
    cudaStream_t s = m.ctx->stream;
    GroupMap& gm = user_side ? m.users : m.items;
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>(ceil_div((int64_t)gm.n_int * 32, EV_THREADS), 1), 148 * 8);
    DevBuf<double> part;
    MML_TRY(part.alloc(blocks));
    const float reg = m.p.biased ? (user_side ? m.p.reg_u : m.p.reg_i) : m.p.regularization;
    regterm_kernel<<<blocks, EV_THREADS, 0, s>>>(user_side ? m.P.p : m.Q.p,
                                                 m.p.biased ? (user_side ? m.bu.p : m.bi.p) : nullptr,
                                                 gm.d_to_ext.p,
                                                 user_side ? m.ratings->count_by_user.p : m.ratings->count_by_item.p,
                                                 gm.n_int, m.kp, reg, m.p.bias_reg,
                                                 m.p.frequency_regularization ? 1 : 0, part.p);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    std::vector<double> h(blocks);
    MML_CUDA(cudaMemcpyAsync(h.data(), part.p, sizeof(double) * blocks, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    double t = 0; for (double x : h) t += x;
    *out = t;
    return MML_OK;
}

// ComputeObjective (BiasedMatrixFactorization.cs:515-552)
static int32_t objective(Sgd& m, double* out)
{
    double sums[4];
    MML_TRY(evaluate_device(m, m.ratings->users.p, m.ratings->items.p, m.ratings->values.p, m.ratings->n, sums));
    double ru = 0, ri = 0;
    MML_TRY(regterm(m, true, &ru));
    MML_TRY(regterm(m, false, &ri));
    *out = sums[3] + ru + ri;
    return MML_OK;
}

// UpdateLearnRate (BiasedMatrixFactorization.cs:225-244, MatrixFactorization.cs:129-132)
static int32_t update_learnrate(Sgd& m)
{
    if (m.p.biased && m.p.bold_driver) {
        double loss = 0;
        MML_TRY(objective(m, &loss));
        loss = (double)(float)loss;   // ComputeObjective returns float (:515)
        if (loss > m.last_loss) m.lr *= 0.5f;
        else if (loss < m.last_loss) m.lr *= 1.05f;
        m.last_loss = loss;
    } else {
        m.lr *= m.p.decay;
    }
    return MML_OK;
}

static int32_t run_serial(Sgd& m, const int32_t* d_idx, int64_t n, int update_user, int update_item)
{
    if (n <= 0) return MML_OK;
    Ratings& r = *m.ratings;
    SgdArgs a = make_args(m, 0);
    if (m.p.biased)
        sgd_serial_kernel<true><<<1, 32, 0, m.ctx->stream>>>(a, d_idx, n, r.users.p, r.items.p, r.values.p,
                                                             m.users.d_to_int.p, m.items.d_to_int.p, m.k, m.kp, update_user, update_item);
    else
        sgd_serial_kernel<false><<<1, 32, 0, m.ctx->stream>>>(a, d_idx, n, r.users.p, r.items.p, r.values.p,
                                                              m.users.d_to_int.p, m.items.d_to_int.p, m.k, m.kp, update_user, update_item);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    return MML_OK;
}

static int32_t run_dsgd_epoch(Sgd& m, const int32_t* h_seq)
{
    cudaStream_t s = m.ctx->stream;
    slot_fn_t sf; epoch_fn_t ef;
    MML_TRY(get_kernels(m, &sf, &ef));
    std::vector<int32_t> seq(m.G);
    for (int t = 0; t < m.G; t++) seq[t] = h_seq ? h_seq[t] : t;
    std::vector<char> seen(m.G, 0);
    for (int t = 0; t < m.G; t++) {
        MML_CHECK(seq[t] >= 0 && seq[t] < m.G && !seen[seq[t]], MML_ERR_ARG, "subepoch_sequence is not a permutation of 0..%d", m.G - 1);
        seen[seq[t]] = 1;
    }
    const int threads = m.W * 32;
    for (int B = 0; B < m.R; B++) {   // R = 1 unless the ring driver moves item blocks between GPUs
        SgdArgs a = make_args(m, B);
        if (m.p.persistent) {
            DevBuf<int32_t>& dseq = m.d_index;   // reuse: serial index cache is unused in DSGD mode
            if ((int64_t)dseq.n < m.G) MML_TRY(dseq.alloc(m.G));
            MML_CUDA(cudaMemcpyAsync(dseq.p, seq.data(), sizeof(int32_t) * m.G, cudaMemcpyHostToDevice, s));
            a.seq = dseq.p;
            void* kargs[] = { (void*)&a };
            MML_CUDA(cudaLaunchCooperativeKernel((const void*)ef, dim3(m.G), dim3(threads), kargs, m.stage_bytes, s));
            m.epoch_base += (uint32_t)m.G + 1u;
            m.launches++;
        } else {
            for (int t = 0; t < m.G; t++) {
                sf<<<m.G, threads, m.stage_bytes, s>>>(a, seq[t]);
                m.launches++;
            }
            MML_CUDA(cudaGetLastError());
        }
    }
    return MML_OK;
}

}  // namespace mml

using namespace mml;

// =================================================================================================
// C ABI
// =================================================================================================
struct mml_sgd { Sgd m; };

extern "C" void mml_mf_params_default(mml_mf_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->biased = 1;
    p->num_factors = 10;
    p->learn_rate = 0.01f;
    p->decay = 1.0f;
    p->regularization = 0.015f;
    p->bias_learn_rate = 1.0f;
    p->bias_reg = 0.01f;
    p->reg_u = 0.015f; p->reg_i = 0.015f;
    p->frequency_regularization = 0;
    p->loss = MML_LOSS_RMSE;
    p->bold_driver = 0;
    p->max_threads = 1;
    p->schedule = MML_SCHEDULE_DSGD;
    p->num_groups = 0; p->num_subgroups = 0;
    p->group_rule = MML_GROUPS_BALANCED;
    p->persistent = -1;
}

extern "C" int32_t mml_sgd_create(mml_ctx* hctx, mml_ratings* hr, const mml_mf_params* p,
                                  const int32_t* user_perm, const int32_t* item_perm, mml_sgd** out)
{
    MML_CHECK(hctx && hr && p && out, MML_ERR_ARG, "mml_sgd_create: NULL argument");
    Ctx* ctx = ctx_of(hctx); Ratings* r = ratings_of(hr);
    MML_CHECK(p->num_factors >= 1 && p->num_factors <= 256, MML_ERR_UNSUPPORTED,
              "mml_sgd_create: num_factors=%d not in [1,256]", p->num_factors);
    MML_CHECK(r->n_users() > 0 && r->n_items() > 0, MML_ERR_ARG, "mml_sgd_create: empty id space");
    MML_CUDA(cudaSetDevice(ctx->device));
    mml_sgd* h = new (std::nothrow) mml_sgd();
    MML_CHECK(h != nullptr, MML_ERR_ARG, "out of host memory");
    Sgd& m = h->m;
    m.ctx = ctx; m.ratings = r; m.p = *p;
    m.k = p->num_factors;
    m.kpl = m.k <= 32 ? 1 : (m.k <= 64 ? 2 : (m.k <= 128 ? 4 : 8));
    m.kp = 32 * m.kpl;
    m.R = 1; m.rank = 0;
    cudaStream_t s = ctx->stream;
    int32_t st = MML_OK;
    do {
        // group shape
        if (p->schedule == MML_SCHEDULE_DSGD) {
            int32_t G = p->num_groups > 0 ? p->num_groups : ctx->sm_count;
            G = std::min(G, std::min(r->n_users(), r->n_items()));
            G = std::max(G, 1);
            int32_t W = p->num_subgroups;
            if (W <= 0) W = (r->n / ((int64_t)G * G) >= 2048) ? 16 : 8;
            W = std::min(W, 32);
            W = std::min(W, std::max(1, std::min(r->n_users(), r->n_items()) / G));
            m.G = G; m.W = std::max(W, 1);
        } else {
            m.G = 1; m.W = 1;
        }
        // counts -> host (group balancing, zero rows)
        std::vector<uint32_t> cu(r->n_users()), ci(r->n_items());
        if (cudaMemcpyAsync(cu.data(), r->count_by_user.p, sizeof(uint32_t) * cu.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaMemcpyAsync(ci.data(), r->count_by_item.p, sizeof(uint32_t) * ci.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("mml_sgd_create: count download failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = MML_ERR_CUDA; break;
        }
        const int32_t rule = p->schedule == MML_SCHEDULE_DSGD ? p->group_rule : MML_GROUPS_PERM_MOD;
        build_group_map(m.users, r->n_users(), cu.data(), p->schedule == MML_SCHEDULE_DSGD ? user_perm : nullptr,
                        m.R, m.rank, m.G, m.W, rule);
        build_group_map(m.items, r->n_items(), ci.data(), p->schedule == MML_SCHEDULE_DSGD ? item_perm : nullptr,
                        m.R, -1, m.G, m.W, rule);
        if ((st = upload_group_map(m.users, s)) || (st = upload_group_map(m.items, s))) break;
        // item group ranges (level 1): group (B, b) covers packed groups (B*G + b)*W .. +W
        m.h_item_ptr.resize((size_t)m.R * m.G + 1);
        int32_t max_items = 0;
        for (int32_t g = 0; g <= m.R * m.G; g++) m.h_item_ptr[g] = m.items.grp_ptr[(size_t)g * m.W];
        for (int32_t g = 0; g < m.R * m.G; g++) max_items = std::max(max_items, m.h_item_ptr[g + 1] - m.h_item_ptr[g]);
        if ((st = m.d_item_ptr.alloc(m.h_item_ptr.size()))) break;
        if (cudaMemcpyAsync(m.d_item_ptr.p, m.h_item_ptr.data(), sizeof(int32_t) * m.h_item_ptr.size(), cudaMemcpyHostToDevice, s) != cudaSuccess) {
            set_error("mml_sgd_create: H2D failed"); st = MML_ERR_CUDA; break;
        }
        // model storage
        if ((st = m.P.alloc((size_t)m.users.n_int * m.kp)) || (st = m.Q.alloc((size_t)m.items.n_int * m.kp)) ||
            (st = m.bu.alloc(m.users.n_int)) || (st = m.bi.alloc(m.items.n_int))) break;
        cudaMemsetAsync(m.P.p, 0, m.P.bytes(), s); cudaMemsetAsync(m.Q.p, 0, m.Q.bytes(), s);
        cudaMemsetAsync(m.bu.p, 0, m.bu.bytes(), s); cudaMemsetAsync(m.bi.p, 0, m.bi.bytes(), s);
        if (p->frequency_regularization) {
            if ((st = m.regw_u.alloc(m.users.n_int)) || (st = m.regw_i.alloc(m.items.n_int))) break;
            const float ru = p->biased ? p->reg_u : p->regularization, ri = p->biased ? p->reg_i : p->regularization;
            regw_kernel<<<grid_n(m.users.n_int), 256, 0, s>>>(m.users.d_to_ext.p, r->count_by_user.p, m.users.n_int, ru, m.regw_u.p);
            regw_kernel<<<grid_n(m.items.n_int), 256, 0, s>>>(m.items.d_to_ext.p, r->count_by_item.p, m.items.n_int, ri, m.regw_i.p);
            m.launches += 2;
        }
        // scale and global bias (BiasedMatrixFactorization.cs:186-190 / MatrixFactorization.cs:124)
        m.min_rating = r->min_rating; m.max_rating = r->max_rating;
        m.range = m.max_rating - m.min_rating;
        if (p->biased) {
            const double avg = (double)(r->average - m.min_rating) / (double)m.range;
            m.global_bias = (float)std::log(avg / (1 - avg));
        } else {
            m.global_bias = r->average;
        }
        m.lr = p->learn_rate;
        // strata
        if (p->schedule == MML_SCHEDULE_DSGD) {
            if ((st = build_strata(m))) break;
            const size_t need = (size_t)max_items * (m.kp + 1) * sizeof(float);
            int max_optin = 0;
            cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
            m.stage_bytes = (need <= (size_t)max_optin) ? ((need + 15) / 16) * 16 : 0;
            if (m.p.persistent < 0) m.p.persistent = 1;
            // item rows that stay in global memory are read through L1, which is only coherent across
            // launches: the single-launch epoch needs the staged (shared-memory) item groups
            if (m.stage_bytes == 0) m.p.persistent = 0;
            if (m.p.persistent) {
                // all G CTAs must be co-resident
                slot_fn_t sf; epoch_fn_t ef;
                if ((st = get_kernels(m, &sf, &ef))) break;
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)ef, m.W * 32, m.stage_bytes) != cudaSuccess) per_sm = 0;
                if ((int64_t)per_sm * ctx->sm_count < m.G) m.p.persistent = 0;
            }
            if ((st = m.flags.alloc(m.G))) break;
            cudaMemsetAsync(m.flags.p, 0, m.flags.bytes(), s);
            m.epoch_base = 0;
        }
        if (cudaEventCreate(&m.ev0) != cudaSuccess || cudaEventCreate(&m.ev1) != cudaSuccess) {
            set_error("cudaEventCreate failed"); st = MML_ERR_CUDA; break;
        }
        if (cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("mml_sgd_create: %s", cudaGetErrorString(cudaGetLastError())); st = MML_ERR_CUDA; break;
        }
    } while (0);
    if (st) { delete h; return st; }
    *out = h;
    return MML_OK;
}

extern "C" int32_t mml_sgd_destroy(mml_sgd* h)
{
    if (!h) return MML_OK;
    cudaSetDevice(h->m.ctx->device);
    cudaStreamSynchronize(h->m.ctx->stream);
    if (h->m.ev0) cudaEventDestroy(h->m.ev0);
    if (h->m.ev1) cudaEventDestroy(h->m.ev1);
    delete h;
    return MML_OK;
}

static int32_t rows_from_host(Sgd& m, GroupMap& gm, const uint32_t* d_counts, const float* h_rows, float* d_int)
{
    cudaStream_t s = m.ctx->stream;
    DevBuf<float> tmp;
    MML_TRY(tmp.alloc((size_t)gm.n_ext * m.k));
    MML_CUDA(cudaMemcpyAsync(tmp.p, h_rows, sizeof(float) * (size_t)gm.n_ext * m.k, cudaMemcpyHostToDevice, s));
    rows_in_kernel<<<grid_n((int64_t)gm.n_int * 32), 256, 0, s>>>(tmp.p, gm.d_to_ext.p, d_counts, gm.n_int, m.k, m.kp, d_int);
    MML_CUDA(cudaGetLastError());
    MML_CUDA(cudaStreamSynchronize(s));
    m.launches++;
    return MML_OK;
}

static int32_t vec_from_host(Sgd& m, GroupMap& gm, const float* h_vec, float* d_int)
{
    cudaStream_t s = m.ctx->stream;
    if (!h_vec) { MML_CUDA(cudaMemsetAsync(d_int, 0, sizeof(float) * std::max(gm.n_int, 1), s)); return MML_OK; }
    DevBuf<float> tmp;
    MML_TRY(tmp.alloc(gm.n_ext));
    MML_CUDA(cudaMemcpyAsync(tmp.p, h_vec, sizeof(float) * gm.n_ext, cudaMemcpyHostToDevice, s));
    vec_in_kernel<<<grid_n(gm.n_int), 256, 0, s>>>(tmp.p, gm.d_to_ext.p, gm.n_int, d_int);
    MML_CUDA(cudaGetLastError());
    MML_CUDA(cudaStreamSynchronize(s));
    m.launches++;
    return MML_OK;
}

static int32_t after_init(Sgd& m)
{
    m.lr = m.p.learn_rate;
    m.has_model = true;
    if (m.p.biased && m.p.bold_driver) {   // BiasedMatrixFactorization.cs:168-169
        MML_TRY(objective(m, &m.last_loss));
        m.last_loss = (double)(float)m.last_loss;
    }
    return MML_OK;
}

extern "C" int32_t mml_sgd_set_model(mml_sgd* h, const float* user_factors, const float* item_factors,
                                     const float* user_bias, const float* item_bias)
{
    MML_CHECK(h && user_factors && item_factors, MML_ERR_ARG, "mml_sgd_set_model: NULL argument");
    Sgd& m = h->m;
    MML_CUDA(cudaSetDevice(m.ctx->device));
    MML_TRY(rows_from_host(m, m.users, m.ratings->count_by_user.p, user_factors, m.P.p));
    MML_TRY(rows_from_host(m, m.items, m.ratings->count_by_item.p, item_factors, m.Q.p));
    MML_TRY(vec_from_host(m, m.users, user_bias, m.bu.p));
    MML_TRY(vec_from_host(m, m.items, item_bias, m.bi.p));
    MML_CUDA(cudaStreamSynchronize(m.ctx->stream));
    return after_init(m);
}

extern "C" int32_t mml_sgd_init_model(mml_sgd* h, uint64_t seed, double init_mean, double init_stddev)
{
    MML_CHECK(h, MML_ERR_ARG, "mml_sgd_init_model: NULL argument");
    Sgd& m = h->m;
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    init_rows_kernel<<<grid_n((int64_t)m.users.n_int * 32), 256, 0, s>>>(m.users.d_to_ext.p, m.ratings->count_by_user.p,
        m.users.n_int, m.k, m.kp, seed, 1, (float)init_mean, (float)init_stddev, m.P.p);
    init_rows_kernel<<<grid_n((int64_t)m.items.n_int * 32), 256, 0, s>>>(m.items.d_to_ext.p, m.ratings->count_by_item.p,
        m.items.n_int, m.k, m.kp, seed, 2, (float)init_mean, (float)init_stddev, m.Q.p);
    MML_CUDA(cudaGetLastError());
    MML_CUDA(cudaMemsetAsync(m.bu.p, 0, m.bu.bytes(), s));
    MML_CUDA(cudaMemsetAsync(m.bi.p, 0, m.bi.bytes(), s));
    MML_CUDA(cudaStreamSynchronize(s));
    m.launches += 2;
    return after_init(m);
}

extern "C" int32_t mml_sgd_get_model(mml_sgd* h, float* user_factors, float* item_factors,
                                     float* user_bias, float* item_bias, float* global_bias, float* current_learnrate)
{
    MML_CHECK(h, MML_ERR_ARG, "mml_sgd_get_model: NULL argument");
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_get_model: no model (call set_model / init_model first)");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    for (int side = 0; side < 2; side++) {
        GroupMap& gm = side ? m.items : m.users;
        float* h_rows = side ? item_factors : user_factors;
        float* h_vec = side ? item_bias : user_bias;
        if (h_rows) {
            DevBuf<float> tmp;
            MML_TRY(tmp.alloc((size_t)gm.n_ext * m.k));
            MML_CUDA(cudaMemsetAsync(tmp.p, 0, tmp.bytes(), s));
            rows_out_kernel<<<grid_n((int64_t)gm.n_int * 32), 256, 0, s>>>(side ? m.Q.p : m.P.p, gm.d_to_ext.p, gm.n_int, m.k, m.kp, tmp.p);
            MML_CUDA(cudaGetLastError());
            MML_CUDA(cudaMemcpyAsync(h_rows, tmp.p, sizeof(float) * (size_t)gm.n_ext * m.k, cudaMemcpyDeviceToHost, s));
            MML_CUDA(cudaStreamSynchronize(s));
            m.launches++;
        }
        if (h_vec) {
            DevBuf<float> tmp;
            MML_TRY(tmp.alloc(gm.n_ext));
            MML_CUDA(cudaMemsetAsync(tmp.p, 0, tmp.bytes(), s));
            vec_out_kernel<<<grid_n(gm.n_int), 256, 0, s>>>(side ? m.bi.p : m.bu.p, gm.d_to_ext.p, gm.n_int, tmp.p);
            MML_CUDA(cudaGetLastError());
            MML_CUDA(cudaMemcpyAsync(h_vec, tmp.p, sizeof(float) * gm.n_ext, cudaMemcpyDeviceToHost, s));
            MML_CUDA(cudaStreamSynchronize(s));
            m.launches++;
        }
    }
    if (global_bias) *global_bias = m.global_bias;
    if (current_learnrate) *current_learnrate = m.lr;
    return MML_OK;
}

extern "C" int32_t mml_sgd_set_learnrate(mml_sgd* h, float lr)
{
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    h->m.lr = lr;
    return MML_OK;
}

extern "C" int32_t mml_sgd_invalidate_index(mml_sgd* h)
{
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    h->m.n_index = -1;
    return MML_OK;
}

extern "C" int32_t mml_sgd_iterate(mml_sgd* h, const int32_t* subepoch_sequence, const int32_t* random_index, int64_t n_index)
{
    MML_CHECK(h, MML_ERR_ARG, "mml_sgd_iterate: NULL argument");
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_iterate: no model (call set_model / init_model first)");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    MML_CUDA(cudaEventRecord(m.ev0, s));
    if (m.p.schedule == MML_SCHEDULE_DSGD) {
        MML_TRY(run_dsgd_epoch(m, subepoch_sequence));
    } else {
        MML_CHECK(n_index == m.ratings->n, MML_ERR_ARG, "mml_sgd_iterate: serial schedule needs RandomIndex of length %lld (got %lld)",
                  (long long)m.ratings->n, (long long)n_index);
        if (m.n_index != n_index) {
            MML_CHECK(random_index != nullptr, MML_ERR_ARG, "mml_sgd_iterate: random_index is NULL");
            for (int64_t t = 0; t < n_index; t++)
                MML_CHECK(random_index[t] >= 0 && random_index[t] < m.ratings->n, MML_ERR_ARG, "random_index[%lld] out of range", (long long)t);
            MML_TRY(m.d_index.alloc(n_index));
            MML_CUDA(cudaMemcpyAsync(m.d_index.p, random_index, sizeof(int32_t) * n_index, cudaMemcpyHostToDevice, s));
            m.n_index = n_index;
        }
        MML_TRY(run_serial(m, m.d_index.p, n_index, 1, 1));
    }
    MML_CUDA(cudaEventRecord(m.ev1, s));
    m.timed = true;
    // UpdateLearnRate: once in single-thread mode, twice in the reference's multi-threaded mode
    // (BiasedMatrixFactorization.cs:216 and :221); plain MF always once (MatrixFactorization.cs:195)
    MML_TRY(update_learnrate(m));
    if (m.p.biased && m.p.max_threads > 1) MML_TRY(update_learnrate(m));
    return MML_OK;
}

extern "C" int32_t mml_sgd_iterate_indices(mml_sgd* h, const int32_t* indices, int64_t n, int32_t update_user, int32_t update_item)
{
    MML_CHECK(h && (indices || n == 0), MML_ERR_ARG, "mml_sgd_iterate_indices: NULL argument");
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_iterate_indices: no model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    for (int64_t t = 0; t < n; t++)
        MML_CHECK(indices[t] >= 0 && indices[t] < m.ratings->n, MML_ERR_ARG, "indices[%lld] out of range", (long long)t);
    DevBuf<int32_t> d;
    MML_TRY(d.alloc(n));
    if (n > 0) MML_CUDA(cudaMemcpyAsync(d.p, indices, sizeof(int32_t) * n, cudaMemcpyHostToDevice, m.ctx->stream));
    MML_TRY(run_serial(m, d.p, n, update_user, update_item));
    MML_CUDA(cudaStreamSynchronize(m.ctx->stream));
    if (!m.p.biased) MML_TRY(update_learnrate(m));   // MatrixFactorization.cs:195
    return MML_OK;
}

extern "C" int32_t mml_sgd_predict(mml_sgd* h, const int32_t* users, const int32_t* items, int64_t n, float* out)
{
    MML_CHECK(h && (n == 0 || (users && items && out)), MML_ERR_ARG, "mml_sgd_predict: NULL argument");
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_predict: no model");
    if (n == 0) return MML_OK;
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    DevBuf<int32_t> du, di; DevBuf<float> dout;
    MML_TRY(du.alloc(n)); MML_TRY(di.alloc(n)); MML_TRY(dout.alloc(n));
    MML_CUDA(cudaMemcpyAsync(du.p, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(di.p, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    predict_kernel<<<grid_n(n * 32), 256, 0, s>>>(make_pred_args(m), du.p, di.p, n, dout.p);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    MML_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

extern "C" int32_t mml_sgd_evaluate(mml_sgd* h, const int32_t* users, const int32_t* items, const float* values, int64_t n, float* out4)
{
    MML_CHECK(h && out4 && (n == 0 || (users && items && values)), MML_ERR_ARG, "mml_sgd_evaluate: NULL argument");
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_evaluate: no model");
    MML_CHECK(n > 0, MML_ERR_ARG, "mml_sgd_evaluate: empty test set");   // Eval/Ratings.cs:98-99 returns null
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    DevBuf<int32_t> du, di; DevBuf<float> dv;
    MML_TRY(du.alloc(n)); MML_TRY(di.alloc(n)); MML_TRY(dv.alloc(n));
    MML_CUDA(cudaMemcpyAsync(du.p, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(di.p, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(dv.p, values, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    double sums[4];
    MML_TRY(evaluate_device(m, du.p, di.p, dv.p, n, sums));
    sums_to_measures(m, sums, n, out4);
    return MML_OK;
}

extern "C" int32_t mml_sgd_evaluate_train(mml_sgd* h, float* out4)
{
    MML_CHECK(h && out4, MML_ERR_ARG, "mml_sgd_evaluate_train: NULL argument");
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_evaluate_train: no model");
    MML_CHECK(m.ratings->n > 0, MML_ERR_ARG, "mml_sgd_evaluate_train: empty training set");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    double sums[4];
    MML_TRY(evaluate_device(m, m.ratings->users.p, m.ratings->items.p, m.ratings->values.p, m.ratings->n, sums));
    sums_to_measures(m, sums, m.ratings->n, out4);
    return MML_OK;
}

extern "C" int32_t mml_sgd_objective(mml_sgd* h, double* out)
{
    MML_CHECK(h && out, MML_ERR_ARG, "mml_sgd_objective: NULL argument");
    MML_CHECK(h->m.has_model, MML_ERR_STATE, "mml_sgd_objective: no model");
    MML_CUDA(cudaSetDevice(h->m.ctx->device));
    return objective(h->m, out);
}

extern "C" int32_t mml_sgd_stats(mml_sgd* h, int64_t* kernel_launches, float* last_iterate_ms)
{
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    Sgd& m = h->m;
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

extern "C" int32_t mml_sgd_strata_info(mml_sgd* h, int32_t* G, int32_t* W, int64_t* n_subblocks, int64_t* staged_bytes)
{
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    if (G) *G = h->m.G;
    if (W) *W = h->m.W;
    if (n_subblocks) *n_subblocks = h->m.n_sub;
    if (staged_bytes) *staged_bytes = (int64_t)h->m.stage_bytes;
    return MML_OK;
}

extern "C" int32_t mml_sgd_schedule_dump(mml_sgd* h, const int32_t* subepoch_sequence, int32_t* order)
{
    MML_CHECK(h && order, MML_ERR_ARG, "mml_sgd_schedule_dump: NULL argument");
    Sgd& m = h->m;
    MML_CHECK(m.p.schedule == MML_SCHEDULE_DSGD, MML_ERR_STATE, "mml_sgd_schedule_dump: not a DSGD model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    const int64_t n = m.ratings->n;
    std::vector<uint32_t> sp((size_t)m.n_sub + 1);
    std::vector<int32_t> idx(std::max<int64_t>(n, 1));
    MML_CUDA(cudaMemcpyAsync(sp.data(), m.sub_ptr.p, sizeof(uint32_t) * sp.size(), cudaMemcpyDeviceToHost, s));
    if (n > 0) MML_CUDA(cudaMemcpyAsync(idx.data(), m.ent_idx.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    int64_t pos = 0;
    const int G = m.G, W = m.W;
    for (int B = 0; B < m.R; B++)
        for (int t = 0; t < G; t++) {
            const int slot = subepoch_sequence ? subepoch_sequence[t] : t;
            MML_CHECK(slot >= 0 && slot < G, MML_ERR_ARG, "subepoch_sequence[%d] out of range", t);
            for (int step = 0; step < W; step++)
                for (int j = 0; j < G; j++)
                    for (int w = 0; w < W; w++) {
                        const size_t sb = ((((size_t)B * G + j) * G + slot) * W + w) * W + step;
                        for (uint32_t e = sp[sb]; e < sp[sb + 1]; e++) order[pos++] = idx[e];
                    }
        }
    MML_CHECK(pos == n, MML_ERR_STATE, "schedule covers %lld of %lld ratings", (long long)pos, (long long)n);
    return MML_OK;
}
