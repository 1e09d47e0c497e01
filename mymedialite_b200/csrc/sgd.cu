// sgd.cu -- the SGD epoch of MatrixFactorization / BiasedMatrixFactorization on sm_100a.
//
// Reference: RatingPrediction/BiasedMatrixFactorization.cs:161-325 (InitModel, Train, Iterate, the
// per-rating update, Predict), :496-552 (objective); RatingPrediction/MatrixFactorization.cs:99-259;
// MultiCore.cs:43-73 (the user x item block stratification the DSGD schedule mirrors).
//
// Schedule. The reference's DSGD mode cuts the rating matrix into g x g blocks by
// (user_perm[u] % g, item_perm[i] % g) and runs g sub-epochs of g mutually disjoint blocks, one
// thread per block. Here:
//   level 1: G worker groups = CTAs (the reference's threads). CTA j owns user group j for the
//            whole epoch; in sub-epoch ("slot") s it holds item group (s + j) mod G, staged in
//            shared memory -- exactly the reference's block schedule with g = G.
//   level 2: inside block (j, b), where the reference's thread walks the block serially, the CTA
//            walks it in ROUNDS: a round is a set of ratings with pairwise distinct users and
//            distinct item rows (a matching of the block's bipartite graph, found at build time
//            by greedy edge colouring). The ratings of a round are independent, so the CTA's
//            workers (sub-warps of L lanes, one rating each) take them in parallel;
//            __syncthreads() separates rounds.
// No two workers of the GPU ever touch the same user row or item row at the same time, so the
// parallel epoch equals a serial pass over the ratings in the order (slot, block, round, entry)
// -- the order mml_sgd_schedule_dump returns and the tests replay through the CPU oracle.
// Hot items (an item with d ratings in a block forces d rounds) get `hot_copies` private copies of
// their row per block; the copies' deltas are summed when the block ends (see sgd_block).
//
// Data layout in HBM. Factor rows are renumbered group-major (all rows of one group are contiguous)
// and padded with zeros to kp floats (32, 64, 128 or 256), so a worker moves a row with 128-bit
// accesses and an item group is one contiguous region. Ratings are stored as entries (internal
// user row, item row inside the staged group image, value) sorted by (block, round);
// round_ptr[] delimits the rounds and blk_round_ptr[] the rounds of each block.
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
//   level 0 (GPU block)  : perm[id] % R          (the reference rule, MultiCore.cs:64; users always: the host shards
//                          the ratings by it) -- items under the BALANCED rule: balanced like the levels below
//   level 1/2, PERM_MOD  : (perm[id] / R) % (G*W) -> g = x % G, w = x / G
//   level 1/2, BALANCED  : ids of a block in descending rating count, each to the lightest group so far
// Hot ids (items only, hot_min > 0): an id with at least hot_min ratings would serialise its whole
// group (its updates form one dependence chain), so it gets no sub-group: it belongs to CTA-level
// group g only and every warp works on a private copy of its row (see sgd_block). At most
// MAX_HOT per CTA-level group; hot rows are numbered first inside their group.
constexpr int32_t HOT_BIT = 1 << 30;
constexpr int32_t MAX_HOT = 8;

static void build_group_map(GroupMap& m, int32_t n_ext, const uint32_t* counts, const int32_t* perm,
                            int32_t R, int32_t only_block /* -1 = all blocks */, int32_t G, int32_t W, int32_t rule,
                            int64_t hot_min, std::vector<int32_t>* hot_cnt /* [n_blocks * G] or NULL */)
{
    m.n_ext = n_ext;
    m.n_blocks = only_block >= 0 ? 1 : R;
    const int32_t T = G * W;
    m.grp.assign(n_ext, -1);
    if (hot_cnt) hot_cnt->assign((size_t)m.n_blocks * G, 0);
    std::vector<std::vector<int32_t>> by_block(m.n_blocks);
    if (only_block < 0 && R > 1 && rule == MML_GROUPS_BALANCED) {
        // Items under the BALANCED rule: the GPU-level blocks are balanced too -- ids in descending rating count, each to
        // the lightest block so far, so the R hottest items land in R different blocks. A ring step lasts as long as its
        // slowest rank, and a block's time grows with the popularity of its hottest rows (their atomics queue up at the L2),
        // so blocks of equal size AND equal hotness keep the ranks in step. Every rank computes the same map (global counts).
        std::vector<int32_t> ids(n_ext);
        std::iota(ids.begin(), ids.end(), 0);
        std::stable_sort(ids.begin(), ids.end(), [&](int32_t a, int32_t b) { return counts[a] > counts[b]; });
        std::vector<int64_t> bload(R, 0);
        for (int32_t id : ids) {
            int32_t best = 0;
            for (int32_t b = 1; b < R; b++) if (bload[b] < bload[best]) best = b;
            bload[best] += std::max<uint32_t>(counts[id], 1u);
            by_block[best].push_back(id);
        }
    } else {
        for (int32_t id = 0; id < n_ext; id++) {
            const int32_t pid = perm ? perm[id] : id;
            const int32_t blk = pid % R;
            if (only_block >= 0 && blk != only_block) continue;
            by_block[only_block >= 0 ? 0 : blk].push_back(id);
        }
    }
    for (int32_t bi = 0; bi < m.n_blocks; bi++) {
        auto& ids = by_block[bi];
        if (rule == MML_GROUPS_PERM_MOD) {
            for (int32_t id : ids) {
                const int32_t pid = perm ? perm[id] : id;
                const int32_t x = (pid / R) % T, g = x % G;
                if (hot_cnt && hot_min > 0 && (int64_t)counts[id] >= hot_min && (*hot_cnt)[(size_t)bi * G + g] < MAX_HOT) {
                    (*hot_cnt)[(size_t)bi * G + g]++;
                    m.grp[id] = ((bi * G + g) * W) | HOT_BIT;
                } else {
                    m.grp[id] = (bi * G + g) * W + (x / G);
                }
            }
            continue;
        }
        std::stable_sort(ids.begin(), ids.end(), [&](int32_t a, int32_t b) { return counts[a] > counts[b]; });
        // loads: per CTA-level group and per packed group; a min-heap of packed groups by (load, index)
        std::vector<int64_t> gload(G, 0), load(T, 0);
        size_t pos = 0;
        if (hot_cnt && hot_min > 0) {
            // hot ids, heaviest first, each to the lightest CTA-level group that still has a hot slot
            for (; pos < ids.size() && (int64_t)counts[ids[pos]] >= hot_min; pos++) {
                int32_t best = -1;
                for (int32_t g = 0; g < G; g++)
                    if ((*hot_cnt)[(size_t)bi * G + g] < MAX_HOT && (best < 0 || gload[g] < gload[best])) best = g;
                if (best < 0) break;
                (*hot_cnt)[(size_t)bi * G + best]++;
                gload[best] += counts[ids[pos]];
                m.grp[ids[pos]] = ((bi * G + best) * W) | HOT_BIT;
            }
            for (int32_t x = 0; x < T; x++) load[x] = gload[x % G] / W;
        }
        std::vector<std::pair<int64_t, int32_t>> heap(T);
        for (int32_t x = 0; x < T; x++) heap[x] = std::make_pair(load[x], x);
        auto cmp = [](const std::pair<int64_t, int32_t>& a, const std::pair<int64_t, int32_t>& b) { return a > b; };
        std::make_heap(heap.begin(), heap.end(), cmp);
        for (; pos < ids.size(); pos++) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            auto& top = heap.back();
            const int32_t x = top.second;
            m.grp[ids[pos]] = (bi * G + (x % G)) * W + (x / G);
            top.first += std::max<uint32_t>(counts[ids[pos]], 1u);   // ids without ratings still spread evenly
            std::push_heap(heap.begin(), heap.end(), cmp);
        }
    }
    // group-major numbering (counting sort, ascending id inside a group); hot rows first in their CTA group
    const int32_t n_grp = m.n_blocks * T;
    m.grp_ptr.assign((size_t)n_grp + 1, 0);
    auto slot_of = [&](int32_t g) { return (g & HOT_BIT) ? (g & ~HOT_BIT) : g; };   // hot rows share slot (bi*G+g)*W
    // two keys per id: (packed slot, hot first). Count hot separately so they precede the cold rows of sub-group 0.
    std::vector<int32_t> n_hot_in((size_t)m.n_blocks * G, 0);
    for (int32_t id = 0; id < n_ext; id++) {
        const int32_t g = m.grp[id];
        if (g < 0) continue;
        m.grp_ptr[slot_of(g) + 1]++;
        if (g & HOT_BIT) n_hot_in[slot_of(g) / W]++;
    }
    for (int32_t g = 0; g < n_grp; g++) m.grp_ptr[g + 1] += m.grp_ptr[g];
    m.n_int = m.grp_ptr[n_grp];
    m.to_int.assign(n_ext, -1);
    m.to_ext.assign(std::max(m.n_int, 1), 0);
    std::vector<int32_t> cur_hot((size_t)m.n_blocks * G), cur_cold(n_grp);
    for (int32_t g = 0; g < n_grp; g++) cur_cold[g] = m.grp_ptr[g] + ((g % W) == 0 ? n_hot_in[g / W] : 0);
    for (int32_t cg = 0; cg < m.n_blocks * G; cg++) cur_hot[cg] = m.grp_ptr[(size_t)cg * W];
    for (int32_t id = 0; id < n_ext; id++) {
        const int32_t g = m.grp[id];
        if (g < 0) continue;
        const int32_t r = (g & HOT_BIT) ? cur_hot[slot_of(g) / W]++ : cur_cold[g]++;
        m.to_int[id] = r;
        m.to_ext[r] = id;
    }
}

// Owned-users mode of the async schedule: inside every user group the users are dealt to the group's `nw` workers, heaviest
// first, each to the lightest worker so far (a worker's load = the ratings of its users, summed over ALL item groups), and the
// group's internal rows are renumbered worker-major (ascending id inside a worker). The entries of a block, sorted by user row,
// then fall into one contiguous slice per worker, and a worker meets the same users in every block of the epoch.
static void pin_users_to_workers(GroupMap& m, const uint32_t* counts, int32_t G, int32_t nw, std::vector<int32_t>& worker_ptr)
{
    worker_ptr.assign((size_t)G * nw + 1, 0);
    std::vector<int32_t> ids, owner;
    std::vector<std::pair<int64_t, int32_t>> heap;
    auto cmp = [](const std::pair<int64_t, int32_t>& a, const std::pair<int64_t, int32_t>& b) { return a > b; };
    for (int32_t g = 0; g < G; g++) {
        const int32_t lo = m.grp_ptr[g], hi = m.grp_ptr[g + 1];
        ids.assign(m.to_ext.begin() + lo, m.to_ext.begin() + hi);
        std::stable_sort(ids.begin(), ids.end(), [&](int32_t a, int32_t b) { return counts[a] > counts[b]; });
        heap.resize(nw);
        for (int32_t w = 0; w < nw; w++) heap[w] = std::make_pair((int64_t)0, w);
        std::make_heap(heap.begin(), heap.end(), cmp);
        owner.assign(ids.size(), 0);
        std::vector<int32_t> n_of(nw, 0);
        for (size_t x = 0; x < ids.size(); x++) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            auto& top = heap.back();
            owner[x] = top.second;
            n_of[top.second]++;
            top.first += std::max<uint32_t>(counts[ids[x]], 1u);
            std::push_heap(heap.begin(), heap.end(), cmp);
        }
        std::vector<int32_t> first(nw + 1, 0);
        for (int32_t w = 0; w < nw; w++) first[w + 1] = first[w] + n_of[w];
        for (int32_t w = 0; w < nw; w++) worker_ptr[(size_t)g * nw + w] = lo + first[w];
        // worker-major, ascending id inside a worker: walk the ids in ascending order and append to their worker's range
        std::vector<int32_t> of_id_sorted(ids.size());
        std::vector<std::pair<int32_t, int32_t>> io(ids.size());
        for (size_t x = 0; x < ids.size(); x++) io[x] = std::make_pair(ids[x], owner[x]);
        std::sort(io.begin(), io.end());
        std::vector<int32_t> cur(first.begin(), first.end() - 1);
        for (auto& pr : io) {
            const int32_t r = lo + cur[pr.second]++;
            m.to_int[pr.first] = r;
            m.to_ext[r] = pr.first;
        }
    }
    worker_ptr[(size_t)G * nw] = m.grp_ptr[G];
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

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// blk = (B*G + j)*G + slot ; bad[0] counts ratings whose user/item is not mapped to this rank
__global__ void strata_block_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t n,
                                    const int32_t* __restrict__ user_grp, const int32_t* __restrict__ item_grp,
                                    int32_t G, uint32_t* __restrict__ key, uint32_t* __restrict__ bad)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const int32_t j = user_grp[users[t]];
        int32_t ig = item_grp[items[t]];
        if (j < 0 || ig < 0) { atomicAdd(bad, 1u); key[t] = 0; continue; }
        ig &= ~HOT_BIT;
        const int32_t B = ig / G, b = ig % G;
        int32_t slot = b - j; if (slot < 0) slot += G;
        key[t] = (uint32_t)((B * G + j) * G + slot);
    }
}

// Per entry (in block order): internal user row, item row inside the staged image of its group (absolute_rows: the internal
// item row itself -- what the async mode and the evaluation of the training set index Q with)
// (cold items and copy 0 of hot items: internal row - first row of the group; copy c >= 1 of hot item x:
// n_it + x * (C - 1) + (c - 1)), value, source index, copy.
__global__ void strata_entries_kernel(const uint32_t* __restrict__ order, const int32_t* __restrict__ users,
                                      const int32_t* __restrict__ items, const float* __restrict__ values, int64_t n,
                                      const int32_t* __restrict__ user_int, const int32_t* __restrict__ item_int,
                                      const int32_t* __restrict__ item_grp, const int32_t* __restrict__ item_ptr,
                                      int32_t C, bool absolute_rows, int32_t* __restrict__ ent_u, int32_t* __restrict__ ent_i,
                                      float* __restrict__ ent_v, int32_t* __restrict__ ent_idx, int8_t* __restrict__ ent_copy)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const uint32_t src = order[t];
        const int32_t u = users[src], it = items[src];
        const int32_t ig = item_grp[it];
        const int32_t Bb = ig & ~HOT_BIT;
        const int32_t i_lo = item_ptr[Bb], n_it = item_ptr[Bb + 1] - i_lo;
        int32_t row = item_int[it] - (absolute_rows ? 0 : i_lo);   // async mode: item rows stay in global memory
        int8_t copy = -1;
        if (ig & HOT_BIT) {
            const int32_t c = (int32_t)(mix32(src) % (uint32_t)C);
            copy = (int8_t)c;
            if (c > 0) row = n_it + row * (C - 1) + (c - 1);
        }
        ent_u[t] = user_int[u];
        ent_i[t] = row;
        ent_v[t] = values[src];
        ent_idx[t] = (int32_t)src;
        ent_copy[t] = copy;
    }
}

// Greedy edge colouring, one thread per block: entry e gets the smallest round not yet used by its user or its
// item row inside this block (128-bit masks per node in `scratch`); entries beyond 128 rounds go to rounds of
// their own. Rounds are therefore matchings: no user and no item row appears twice in a round.
constexpr int COLOR_BITS = 12;
__global__ void strata_color_kernel(const uint32_t* __restrict__ blk_ptr, int32_t n_blk, int32_t G,
                                    const int32_t* __restrict__ ent_u, const int32_t* __restrict__ ent_i,
                                    const int32_t* __restrict__ user_ptr, int32_t nu_max, int32_t nr_max,
                                    unsigned long long* __restrict__ scratch, uint32_t* __restrict__ color,
                                    uint32_t* __restrict__ overflow_flag)
{
    const int32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= n_blk) return;
    const int32_t j = (blk / G) % G;
    const int32_t u_lo = user_ptr[j];
    unsigned long long* mu = scratch + (size_t)blk * (size_t)(nu_max + nr_max) * 2;
    unsigned long long* mi = mu + (size_t)nu_max * 2;
    uint32_t extra = 128;
    for (uint32_t e = blk_ptr[blk]; e < blk_ptr[blk + 1]; e++) {
        const int32_t u = ent_u[e] - u_lo, r = ent_i[e];
        const unsigned long long lo = mu[2 * u] | mi[2 * r];
        uint32_t c;
        if (~lo) {
            c = (uint32_t)__ffsll((long long)~lo) - 1u;
            mu[2 * u] |= 1ull << c; mi[2 * r] |= 1ull << c;
        } else {
            const unsigned long long hi = mu[2 * u + 1] | mi[2 * r + 1];
            if (~hi) {
                const uint32_t b = (uint32_t)__ffsll((long long)~hi) - 1u;
                c = 64u + b;
                mu[2 * u + 1] |= 1ull << b; mi[2 * r + 1] |= 1ull << b;
            } else {
                c = extra++;
            }
        }
        if (c >= (1u << COLOR_BITS)) { atomicExch(overflow_flag, 1u); c = (1u << COLOR_BITS) - 1u; }
        color[e] = c;
    }
}

__global__ void strata_key2_kernel(const uint32_t* __restrict__ blk_sorted, const uint32_t* __restrict__ color, int64_t n,
                                   uint32_t* __restrict__ key2)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) key2[t] = (blk_sorted[t] << COLOR_BITS) | color[t];
}

// head[t] = 1 where a new round starts; rounds_in_blk[blk]++ for every head
__global__ void strata_heads_kernel(const uint32_t* __restrict__ key2, int64_t n, uint32_t* __restrict__ head,
                                    uint32_t* __restrict__ rounds_in_blk)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const bool h = t == 0 || key2[t] != key2[t - 1];
        head[t] = h ? 1u : 0u;
        if (h) atomicAdd(&rounds_in_blk[key2[t] >> COLOR_BITS], 1u);
    }
}

__global__ void strata_round_ptr_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ round_id, int64_t n,
                                        uint32_t n_rounds, uint32_t* __restrict__ round_ptr)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) if (head[t]) round_ptr[round_id[t]] = (uint32_t)t;
    if (blockIdx.x == 0 && threadIdx.x == 0) round_ptr[n_rounds] = (uint32_t)n;
}

__global__ void gather_entries_kernel(const uint32_t* __restrict__ perm, int64_t n,
                                      const int32_t* __restrict__ u0, const int32_t* __restrict__ i0, const float* __restrict__ v0,
                                      const int32_t* __restrict__ x0, const int8_t* __restrict__ c0,
                                      int32_t* __restrict__ u1, int32_t* __restrict__ i1, float* __restrict__ v1,
                                      int32_t* __restrict__ x1, int8_t* __restrict__ c1)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const uint32_t s = perm[t];
        u1[t] = u0[s]; i1[t] = i0[s]; v1[t] = v0[s]; x1[t] = x0[s]; c1[t] = c0[s];
    }
}

// Async mode: worker w of the CTA gets the w-th of n_workers equal slices of the block's entries (sorted by
// user), cut at user boundaries so that every user row of the block belongs to exactly one worker.
// (A two-level cut balancing entries + 2 x user runs per warp was measured and dropped: hand-over waits stayed at 12 % of
// the CTA time, 14.59 ms against 14.50 ms per epoch -- the waits are not caused by uneven run counts.)
__global__ void strata_split_kernel(const uint32_t* __restrict__ blk_ptr, int32_t n_blk, int32_t n_workers,
                                    const int32_t* __restrict__ ent_u, uint32_t* __restrict__ wptr)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_blk * (n_workers + 1)) return;
    const int32_t blk = (int32_t)(t / (n_workers + 1)), w = (int32_t)(t % (n_workers + 1));
    const uint32_t beg = blk_ptr[blk], end = blk_ptr[blk + 1];
    uint32_t cut = beg + (uint32_t)(((uint64_t)(end - beg) * (uint64_t)w) / (uint64_t)n_workers);
    if (w == n_workers) cut = end;
    while (cut > beg && cut < end && ent_u[cut] == ent_u[cut - 1]) cut++;
    wptr[t] = cut;
}

// Owned-users mode: worker w's slice of a block = the entries of its own users = [first entry with user row >= worker_ptr[j nw + w],
// the same for w + 1) -- binary search in the block's entries (sorted by user row).
__global__ void strata_split_owned_kernel(const uint32_t* __restrict__ blk_ptr, int32_t n_blk, int32_t G, int32_t n_workers,
                                          const int32_t* __restrict__ ent_u, const int32_t* __restrict__ worker_ptr, uint32_t* __restrict__ wptr)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_blk * (n_workers + 1)) return;
    const int32_t blk = (int32_t)(t / (n_workers + 1)), w = (int32_t)(t % (n_workers + 1));
    const int32_t j = (blk / G) % G;
    uint32_t lo = blk_ptr[blk], hi = blk_ptr[blk + 1];
    if (w == n_workers) { wptr[t] = hi; return; }
    const int32_t first = worker_ptr[(size_t)j * n_workers + w];
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (ent_u[mid] < first) lo = mid + 1; else hi = mid;
    }
    wptr[t] = lo;
}

__global__ void gather_key_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, int64_t n, uint32_t* __restrict__ out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) out[t] = src[idx[t]];
}

__global__ void user_key_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ user_int, int64_t n, uint32_t* __restrict__ key)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) { const int32_t r = user_int[users[t]]; key[t] = r < 0 ? 0u : (uint32_t)r; }
}

static int32_t build_strata_async(Sgd& m, int32_t n_workers)
{
    Ratings& r = *m.ratings;
    cudaStream_t s = m.ctx->stream;
    const int64_t n = r.n;
    const int64_t n_blk64 = (int64_t)m.RB * m.G * m.G;
    MML_CHECK(n_blk64 * (n_workers + 1) < ((int64_t)1 << 31), MML_ERR_ARG, "strata: %lld blocks is too many (G=%d)", (long long)n_blk64, m.G);
    const int32_t n_blk = (int32_t)n_blk64;
    m.n_blk = n_blk;
    m.n_workers = n_workers;
    // the five n-sized temporaries come from the stream-ordered pool (kept between calls: a second Train() in the same
    // process does not pay cudaMalloc / cudaFree of 2 GB again)
    StreamBuf<uint32_t> blk, key, vals, ktmp, vtmp;
    DevBuf<uint32_t> cnt, bad, blk_ptr;
    MML_TRY(blk.alloc(n, s)); MML_TRY(key.alloc(n, s)); MML_TRY(vals.alloc(n, s)); MML_TRY(ktmp.alloc(n, s)); MML_TRY(vtmp.alloc(n, s));
    MML_TRY(cnt.alloc(n_blk)); MML_TRY(bad.alloc(1)); MML_TRY(blk_ptr.alloc((size_t)n_blk + 1));
    MML_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(uint32_t), s));
    MML_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
    strata_block_kernel<<<grid_n(n), 256, 0, s>>>(r.users.p, r.items.p, n, m.users.d_grp.p, m.items.d_grp.p, m.G, blk.p, bad.p);
    MML_CUDA(cudaGetLastError());
    uint32_t h_bad = 0;
    MML_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    MML_CHECK(h_bad == 0, MML_ERR_ARG,
              "strata: %u ratings belong to users of another GPU block (pass each rank the ratings of its own users)", h_bad);
    MML_TRY(histogram_i32((const int32_t*)blk.p, n, cnt.p, s));
    MML_TRY(exclusive_scan_u32(cnt.p, blk_ptr.p, n_blk, s));
    // order by (block, user row, rating index): stable LSD over the user key, then over the block key
    user_key_kernel<<<grid_n(n), 256, 0, s>>>(r.users.p, m.users.d_to_int.p, n, key.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(iota_u32(vals.p, n, s));
    MML_TRY(radix_sort_pairs(key.p, vals.p, ktmp.p, vtmp.p, n, bits_for((uint32_t)std::max(m.users.n_int - 1, 1)), s));
    gather_key_kernel<<<grid_n(n), 256, 0, s>>>(blk.p, vals.p, n, key.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(radix_sort_pairs(key.p, vals.p, ktmp.p, vtmp.p, n, bits_for((uint32_t)std::max(n_blk - 1, 1)), s));
    MML_TRY(m.ent_u.alloc(n)); MML_TRY(m.ent_i.alloc(n)); MML_TRY(m.ent_v.alloc(n)); MML_TRY(m.ent_idx.alloc(n));
    MML_TRY(m.ent_copy.alloc(n));
    strata_entries_kernel<<<grid_n(n), 256, 0, s>>>(vals.p, r.users.p, r.items.p, r.values.p, n,
                                                    m.users.d_to_int.p, m.items.d_to_int.p, m.items.d_grp.p, m.d_item_ptr.p,
                                                    1, true, m.ent_u.p, m.ent_i.p, m.ent_v.p, m.ent_idx.p, m.ent_copy.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(m.wptr.alloc((size_t)n_blk * (n_workers + 1)));
    const int64_t nt = (int64_t)n_blk * (n_workers + 1);
    if (m.owned) {
        DevBuf<int32_t> d_wp;
        MML_TRY(d_wp.alloc(m.h_worker_ptr.size()));
        MML_CUDA(cudaMemcpyAsync(d_wp.p, m.h_worker_ptr.data(), sizeof(int32_t) * m.h_worker_ptr.size(), cudaMemcpyHostToDevice, s));
        strata_split_owned_kernel<<<(unsigned)ceil_div(nt, 256), 256, 0, s>>>(blk_ptr.p, n_blk, m.G, n_workers, m.ent_u.p, d_wp.p, m.wptr.p);
        MML_CUDA(cudaGetLastError());
        MML_CUDA(cudaStreamSynchronize(s));
    } else {
        strata_split_kernel<<<(unsigned)ceil_div(nt, 256), 256, 0, s>>>(blk_ptr.p, n_blk, n_workers, m.ent_u.p, m.wptr.p);
        MML_CUDA(cudaGetLastError());
        MML_CUDA(cudaStreamSynchronize(s));
    }
    m.launches += 14;
    return MML_OK;
}

static int32_t build_strata(Sgd& m, int32_t nu_max, int32_t nr_max)
{
    Ratings& r = *m.ratings;
    cudaStream_t s = m.ctx->stream;
    const int64_t n = r.n;
    const int64_t n_blk64 = (int64_t)m.RB * m.G * m.G;
    MML_CHECK(n_blk64 < ((int64_t)1 << (32 - COLOR_BITS)), MML_ERR_ARG, "strata: %lld blocks is too many (G=%d)", (long long)n_blk64, m.G);
    const int32_t n_blk = (int32_t)n_blk64;
    m.n_blk = n_blk;
    DevBuf<uint32_t> key, vals, ktmp, vtmp, cnt, bad, color;
    MML_TRY(key.alloc(n)); MML_TRY(vals.alloc(n)); MML_TRY(ktmp.alloc(n)); MML_TRY(vtmp.alloc(n));
    MML_TRY(cnt.alloc(n_blk)); MML_TRY(bad.alloc(2)); MML_TRY(color.alloc(n));
    MML_CUDA(cudaMemsetAsync(bad.p, 0, 2 * sizeof(uint32_t), s));
    MML_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
    // 1. entries in block order (stable: ascending rating index inside a block)
    strata_block_kernel<<<grid_n(n), 256, 0, s>>>(r.users.p, r.items.p, n, m.users.d_grp.p, m.items.d_grp.p, m.G, key.p, bad.p);
    MML_CUDA(cudaGetLastError());
    uint32_t h_bad[2] = {0, 0};
    MML_CUDA(cudaMemcpyAsync(h_bad, bad.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    MML_CHECK(h_bad[0] == 0, MML_ERR_ARG,
              "strata: %u ratings belong to users of another GPU block (pass each rank the ratings of its own users)", h_bad[0]);
    DevBuf<uint32_t> blk_ptr;
    MML_TRY(blk_ptr.alloc((size_t)n_blk + 1));
    MML_TRY(histogram_i32((const int32_t*)key.p, n, cnt.p, s));
    MML_TRY(exclusive_scan_u32(cnt.p, blk_ptr.p, n_blk, s));
    MML_TRY(iota_u32(vals.p, n, s));
    MML_TRY(radix_sort_pairs(key.p, vals.p, ktmp.p, vtmp.p, n, bits_for((uint32_t)std::max(n_blk - 1, 1)), s));
    DevBuf<int32_t> u0, i0, x0; DevBuf<float> v0; DevBuf<int8_t> c0;
    MML_TRY(u0.alloc(n)); MML_TRY(i0.alloc(n)); MML_TRY(x0.alloc(n)); MML_TRY(v0.alloc(n)); MML_TRY(c0.alloc(n));
    strata_entries_kernel<<<grid_n(n), 256, 0, s>>>(vals.p, r.users.p, r.items.p, r.values.p, n,
                                                    m.users.d_to_int.p, m.items.d_to_int.p, m.items.d_grp.p, m.d_item_ptr.p,
                                                    m.hot_copies, false, u0.p, i0.p, v0.p, x0.p, c0.p);
    MML_CUDA(cudaGetLastError());
    // 2. rounds: greedy edge colouring per block
    {
        DevBuf<unsigned long long> scratch;
        const size_t words = (size_t)n_blk * (size_t)(nu_max + nr_max) * 2;
        MML_TRY(scratch.alloc(words));
        MML_CUDA(cudaMemsetAsync(scratch.p, 0, scratch.bytes(), s));
        strata_color_kernel<<<(unsigned)ceil_div(n_blk, 64), 64, 0, s>>>(blk_ptr.p, n_blk, m.G, u0.p, i0.p, m.d_user_ptr.p,
                                                                          nu_max, nr_max, scratch.p, color.p, bad.p + 1);
        MML_CUDA(cudaGetLastError());
        MML_CUDA(cudaMemcpyAsync(h_bad + 1, bad.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaStreamSynchronize(s));
        MML_CHECK(h_bad[1] == 0, MML_ERR_UNSUPPORTED, "strata: a block needs more than %d rounds (one user or item dominates it); use more groups", 1 << COLOR_BITS);
    }
    // 3. entries in (block, round) order; key currently holds the sorted block ids
    strata_key2_kernel<<<grid_n(n), 256, 0, s>>>(key.p, color.p, n, ktmp.p);
    MML_CUDA(cudaGetLastError());
    MML_CUDA(cudaMemcpyAsync(key.p, ktmp.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    MML_TRY(iota_u32(vals.p, n, s));
    MML_TRY(radix_sort_pairs(key.p, vals.p, ktmp.p, vtmp.p, n, bits_for((uint32_t)std::max(n_blk - 1, 1)) + COLOR_BITS, s));
    MML_TRY(m.ent_u.alloc(n)); MML_TRY(m.ent_i.alloc(n)); MML_TRY(m.ent_v.alloc(n)); MML_TRY(m.ent_idx.alloc(n));
    MML_TRY(m.ent_copy.alloc(n));
    gather_entries_kernel<<<grid_n(n), 256, 0, s>>>(vals.p, n, u0.p, i0.p, v0.p, x0.p, c0.p,
                                                    m.ent_u.p, m.ent_i.p, m.ent_v.p, m.ent_idx.p, m.ent_copy.p);
    MML_CUDA(cudaGetLastError());
    // 4. round boundaries
    DevBuf<uint32_t> head, round_id, rounds_in_blk;
    MML_TRY(head.alloc(n)); MML_TRY(round_id.alloc((size_t)n + 1)); MML_TRY(rounds_in_blk.alloc(n_blk));
    MML_CUDA(cudaMemsetAsync(rounds_in_blk.p, 0, rounds_in_blk.bytes(), s));
    strata_heads_kernel<<<grid_n(n), 256, 0, s>>>(key.p, n, head.p, rounds_in_blk.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(exclusive_scan_u32(head.p, round_id.p, n, s));
    uint32_t n_rounds = 0;
    MML_CUDA(cudaMemcpyAsync(&n_rounds, round_id.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    m.n_rounds = n_rounds;
    MML_TRY(m.round_ptr.alloc((size_t)n_rounds + 1));
    strata_round_ptr_kernel<<<grid_n(n), 256, 0, s>>>(head.p, round_id.p, n, n_rounds, m.round_ptr.p);
    MML_CUDA(cudaGetLastError());
    MML_TRY(m.blk_round_ptr.alloc((size_t)n_blk + 1));
    MML_TRY(exclusive_scan_u32(rounds_in_blk.p, m.blk_round_ptr.p, n_blk, s));
    MML_CUDA(cudaStreamSynchronize(s));
    m.launches += 16;
    return MML_OK;
}

// =================================================================================================
// device: the per-rating update
// =================================================================================================
struct SgdArgs {
    float* P; float* Q; float* bu; float* bi;
    const int32_t* ent_u; const int32_t* ent_i; const float* ent_v;
    const uint32_t* wptr;           // async mode, offset to B: [G*G][n_workers + 1] worker slices of each block
    const uint32_t* round_ptr;      // [n_rounds + 1]
    const uint32_t* blk_round_ptr;  // offset to the GPU-level item block B: [G*G + 1]
    const int32_t* item_ptr;        // offset to B: [G + 1] internal item row range per item group
    const int32_t* hot_cnt;         // offset to B: [G] hot items per item group (their rows come first)
    const float* regw_u; const float* regw_i;   // frequency regularisation weights or NULL
    unsigned long long* wait_stats; // diagnostic (MMLB200_SGD_WAITSTATS): [0] cycles CTAs spent in hand-over waits, [1] CTA cycles
    uint32_t* flags;                // persistent kernel: progress counter per CTA
    const int32_t* seq;             // persistent kernel: sub-epoch sequence [G] (device)
    uint32_t epoch_base;
    int32_t G, C;                   // worker groups, copies per hot item
    int32_t n_workers;              // async mode: workers per worker group the slices were cut for
    int32_t cpg;                    // async mode: CTAs per worker group (they share the group's blocks)
    float hot_scale;                // merge of hot-item copies: 1 = sum of the chains' steps, 1/C = average
    float lr, gb, minr, range, reg_u, reg_i, blr, breg;
    int32_t loss;
};

// A factor row as seen by one lane of an L-lane worker: KPL = kp / L floats as KPL/4 128-bit pieces;
// piece v of lane sl is float4 number v*L + sl of the row (consecutive lanes -> consecutive 16 bytes).
template <int L, int KPL>
struct Row {
    static_assert(KPL % 4 == 0, "rows move in 128-bit pieces");
    static __device__ __forceinline__ void load(float (&r)[KPL], const float* row, int sl)
    {
#pragma unroll
        for (int v = 0; v < KPL / 4; v++) {
            const float4 x = reinterpret_cast<const float4*>(row)[v * L + sl];
            r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
        }
    }
    static __device__ __forceinline__ void store(const float (&r)[KPL], float* row, int sl)
    {
#pragma unroll
        for (int v = 0; v < KPL / 4; v++)
            reinterpret_cast<float4*>(row)[v * L + sl] = make_float4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
    }
    // L1-bypassing load (rows other SMs or the L2 atomic unit modify)
    static __device__ __forceinline__ void load_cg(float (&r)[KPL], const float* row, int sl)
    {
#pragma unroll
        for (int v = 0; v < KPL / 4; v++) {
            float4 x;
            asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                         : "l"(reinterpret_cast<const float4*>(row) + (v * L + sl)) : "memory");
            r[4 * v] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
        }
    }
};

template <int L>
__device__ __forceinline__ float worker_sum(float v)
{
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) { return worker_sum<32>(v); }

// BiasedMatrixFactorization.cs:264-310 (BIASED) / MatrixFactorization.cs:166-196 for one rating, fp32.
// p and q are updated from their pre-update values; the biases before the factors.
template <int L, int KPL, bool BIASED>
__device__ __forceinline__ void sgd_update(const SgdArgs& a, float (&p)[KPL], float (&q)[KPL],
                                           float& bu, float& bi, float v, float regu, float regi)
{
    float dot = 0.f;
#pragma unroll
    for (int f = 0; f < KPL; f++) dot = fmaf(p[f], q[f], dot);
    dot = worker_sum<L>(dot);
    float gc;
    if (BIASED) {
        const float score = ((a.gb + bu) + bi) + dot;
        const float sig = __fdividef(1.f, 1.f + __expf(-score));   // ex2.approx / rcp.approx: ~1e-6 relative
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

// The same update for the async mode, which applies the item step as an atomic add: returns the item row's step
// dq = lr (g p - reg_i q) (and dbi for the bias) instead of the new row, and folds the learn rate into the
// coefficients: p <- (1 - lr reg_u) p + (lr g) q -- two instructions per factor and row instead of three.
template <int L, int KPL, bool BIASED>
__device__ __forceinline__ void sgd_update_delta(const SgdArgs& a, float (&p)[KPL], const float (&q)[KPL], float (&dq)[KPL],
                                                 float& bu, const float bi, float& dbi, float v, float regu, float regi)
{
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int f = 0; f < KPL; f += 2) { d0 = fmaf(p[f], q[f], d0); d1 = fmaf(p[f + 1], q[f + 1], d1); }
    const float dot = worker_sum<L>(d0 + d1);
    float gc;
    dbi = 0.f;
    if (BIASED) {
        const float score = ((a.gb + bu) + bi) + dot;
        const float sig = __fdividef(1.f, 1.f + __expf(-score));
        const float err = v - (a.minr + sig * a.range);
        if (a.loss == MML_LOSS_RMSE) gc = err * sig * (1.f - sig) * a.range;
        else if (a.loss == MML_LOSS_MAE) gc = (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * sig * (1.f - sig) * a.range;
        else gc = err;
        const float step = a.blr * a.lr;
        bu += step * (gc - a.breg * regu * bu);
        dbi = step * (gc - a.breg * regi * bi);
    } else {
        gc = v - (a.gb + dot);
    }
    const float lg = a.lr * gc, cu = fmaf(-a.lr, regu, 1.f), ci = a.lr * regi;
#pragma unroll
    for (int f = 0; f < KPL; f++) {
        const float pf = p[f], qf = q[f];
        p[f] = fmaf(lg, qf, cu * pf);
        dq[f] = fmaf(lg, pf, -(ci * qf));
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
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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

// Stage item group rows [i_lo, i_lo + n_it) (+ biases) into shared memory; hot items (the first n_hot rows)
// additionally get copies 1..C-1 at rows n_it + x*(C-1) + (c-1) and a snapshot at rows n_it + n_hot*(C-1) + x.
// Lock words (async mode) are cleared. Ends with a barrier.
template <int KP, bool BIASED>
__device__ __forceinline__ void stage_group(const SgdArgs& a, int i_lo, int n_it, int n_hot, int C,
                                            float* sQ, float* sB, unsigned* sLock)
{
    const float4* src = reinterpret_cast<const float4*>(a.Q + (size_t)i_lo * KP);
    float4* dst = reinterpret_cast<float4*>(sQ);
    const int n4 = n_it * (KP / 4);
    for (int t = threadIdx.x; t < n4; t += blockDim.x) dst[t] = ld_cg_f4(src + t);
    if (BIASED) for (int t = threadIdx.x; t < n_it; t += blockDim.x) sB[t] = ld_cg_f(a.bi + i_lo + t);
    if (sLock) for (int t = threadIdx.x; t < n_it + n_hot * C; t += blockDim.x) sLock[t] = 0u;
    if (n_hot > 0) {
        __syncthreads();
        constexpr int per = KP / 4;
        for (int t = threadIdx.x; t < n_hot * C * per; t += blockDim.x) {
            const int x = t / (C * per), c = (t / per) % C, f = t % per;
            const int row = (c < C - 1) ? (n_it + x * (C - 1) + c) : (n_it + n_hot * (C - 1) + x);
            dst[(size_t)row * per + f] = dst[(size_t)x * per + f];
        }
        if (BIASED) for (int t = threadIdx.x; t < n_hot * C; t += blockDim.x) {
            const int x = t / C, c = t % C;
            const int row = (c < C - 1) ? (n_it + x * (C - 1) + c) : (n_it + n_hot * (C - 1) + x);
            sB[row] = sB[x];
        }
    }
    __syncthreads();
}

// Merge the hot items' copies (row = old + scale * sum_c (copy_c - old), copy 0 being the row itself; scale = 1
// sums the chains' steps, scale = 1/C averages them) and write the group back. Call after a barrier.
template <int KP, bool BIASED>
__device__ __forceinline__ void unstage_group(const SgdArgs& a, int i_lo, int n_it, int n_hot, int C,
                                              float* sQ, float* sB)
{
    if (n_hot > 0) {
        const int old0 = n_it + n_hot * (C - 1);
        const float scale = a.hot_scale;
        for (int t = threadIdx.x; t < n_hot * KP; t += blockDim.x) {
            const int x = t / KP, f = t % KP;
            const float old = sQ[(size_t)(old0 + x) * KP + f];
            float acc = sQ[(size_t)x * KP + f] - old;
            for (int c = 1; c < C; c++) acc += sQ[(size_t)(n_it + x * (C - 1) + (c - 1)) * KP + f] - old;
            sQ[(size_t)x * KP + f] = old + scale * acc;
        }
        if (BIASED) for (int x = threadIdx.x; x < n_hot; x += blockDim.x) {
            const float old = sB[old0 + x];
            float acc = sB[x] - old;
            for (int c = 1; c < C; c++) acc += sB[n_it + x * (C - 1) + (c - 1)] - old;
            sB[x] = old + scale * acc;
        }
        __syncthreads();
    }
    float4* dst = reinterpret_cast<float4*>(a.Q + (size_t)i_lo * KP);
    const float4* src = reinterpret_cast<const float4*>(sQ);
    const int n4 = n_it * (KP / 4);
    for (int t = threadIdx.x; t < n4; t += blockDim.x) dst[t] = src[t];
    if (BIASED) for (int t = threadIdx.x; t < n_it; t += blockDim.x) a.bi[i_lo + t] = sB[t];
}

// Work of CTA j on block (j, b): stage item group b, walk the block's rounds, write the group back.
//
// Shared-memory image of the item group (STAGE): rows [0, n_it) are the group's item rows (hot items
// first); rows [n_it, n_it + h*(C-1)) are private copies 1..C-1 of the h hot items (copy 0 is the row
// itself); rows [n_it + h*(C-1), n_it + h*C) keep the hot rows as they were at block start. Each copy is
// a separate node for the round colouring, so a hot item's updates in this block run as C independent
// chains, merged at the end: row = copy0 + sum_{c>=1} (copy_c - old). The item biases mirror the rows in sB.
// Without STAGE (item group too large for shared memory) item rows are updated in global memory
// and there are no hot items.
//
// A worker is L consecutive lanes (kp = L * KPL): 32 / L ratings per warp instruction.
template <int L, int KPL, bool BIASED, bool STAGE>
__device__ __forceinline__ void sgd_block(const SgdArgs& a, const int j, const int slot, float* smem)
{
    constexpr int KP = L * KPL;
    constexpr int WPW = 32 / L;                        // workers per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sl = lane % L;                           // lane inside the worker
    const int wid = warp * WPW + lane / L;             // worker inside the CTA
    const int n_workers = (blockDim.x >> 5) * WPW;
    const int C = a.C;
    int b = slot + j; if (b >= a.G) b -= a.G;
    const int i_lo = a.item_ptr[b], i_hi = a.item_ptr[b + 1];
    const int n_it = i_hi - i_lo;
    const int n_hot = STAGE ? a.hot_cnt[b] : 0;
    const int n_rows = n_it + n_hot * C;
    float* sQ = smem;
    float* sB = smem + (size_t)n_rows * KP;
    const uint32_t r0 = a.blk_round_ptr[j * a.G + slot], r1 = a.blk_round_ptr[j * a.G + slot + 1];
    if (STAGE) stage_group<KP, BIASED>(a, i_lo, n_it, n_hot, C, sQ, sB, nullptr);

    // round boundaries 32 at a time: lane l holds round_ptr[rbase + l]
    uint32_t rbase = r0;
    uint32_t my_ptr = (rbase + lane <= r1) ? a.round_ptr[rbase + lane] : 0u;
    // entry of this worker for the first sub-round of the next round, fetched one round ahead
    int nu = 0, ni = 0; float nv = 0.f;
    if (r0 < r1) {
        const uint32_t e = __shfl_sync(0xffffffffu, my_ptr, 0) + (uint32_t)wid;
        const uint32_t eb = __shfl_sync(0xffffffffu, my_ptr, 1);
        if (e < eb) { nu = a.ent_u[e]; ni = a.ent_i[e]; nv = a.ent_v[e]; }
    }
    for (uint32_t r = r0; r < r1; r++) {
        int ridx = (int)(r - rbase);
        if (ridx == 31) {   // slot 31 holds this round's start; reload so that both bounds are in registers
            rbase = r; ridx = 0;
            my_ptr = (rbase + lane <= r1) ? a.round_ptr[rbase + lane] : 0u;
        }
        const uint32_t ea = __shfl_sync(0xffffffffu, my_ptr, ridx);
        const uint32_t eb = __shfl_sync(0xffffffffu, my_ptr, ridx + 1);
        int cu = nu, ci = ni; float cv = nv;
        if (r + 1 < r1) {   // next round's first entry (read-only data: safe before the barrier)
            const uint32_t en_b = (ridx + 2 <= 31) ? __shfl_sync(0xffffffffu, my_ptr, ridx + 2) : a.round_ptr[r + 2];
            const uint32_t en = eb + (uint32_t)wid;
            if (en < en_b) { nu = a.ent_u[en]; ni = a.ent_i[en]; nv = a.ent_v[en]; }
        }
        // sub-rounds: the warp iterates while its first worker still has an entry
        for (uint32_t e0 = ea + (uint32_t)(warp * WPW); e0 < eb; e0 += (uint32_t)n_workers) {
            const uint32_t e = e0 + (uint32_t)(lane / L);
            const bool active = e < eb;
            if (e0 != ea + (uint32_t)(warp * WPW)) {   // later sub-rounds fetch their entry here
                if (active) { cu = a.ent_u[e]; ci = a.ent_i[e]; cv = a.ent_v[e]; }
            }
            float p[KPL], q[KPL];
            float bu_v = 0.f, bi_v = 0.f, regu = a.reg_u, regi = a.reg_i;
            float* prow = a.P + (size_t)(active ? cu : 0) * KP;
            float* qrow = STAGE ? (sQ + (size_t)(active ? ci : 0) * KP) : (a.Q + (size_t)(i_lo + (active ? ci : 0)) * KP);
            float* bip = STAGE ? (sB + (active ? ci : 0)) : (a.bi + i_lo + (active ? ci : 0));
            Row<L, KPL>::load(p, prow, sl);
            Row<L, KPL>::load(q, qrow, sl);
            if (BIASED) { bu_v = a.bu[active ? cu : 0]; bi_v = *bip; }
            if (a.regw_u) {
                regu = a.regw_u[active ? cu : 0];
                const int it = active ? ci : 0;
                regi = a.regw_i[i_lo + (it < n_it ? it : (it - n_it) / (C - 1))];
            }
            sgd_update<L, KPL, BIASED>(a, p, q, bu_v, bi_v, cv, regu, regi);
            if (active) {
                Row<L, KPL>::store(p, prow, sl);
                Row<L, KPL>::store(q, qrow, sl);
                if (BIASED && sl == 0) { a.bu[cu] = bu_v; *bip = bi_v; }
            }
        }
        __syncthreads();
    }
    // the last __syncthreads() above (or the staging barrier, for an empty block) ordered all shared-memory updates
    if (STAGE) unstage_group<KP, BIASED>(a, i_lo, n_it, n_hot, C, sQ, sB);
}

__device__ __forceinline__ void red_add_f4(float* p, float x, float y, float z, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void red_add_f(float* p, float x)
{
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(x) : "memory");
}

// Async mode of block (j, b): the block's entries are sorted by user and cut into one slice per worker (whole
// user runs), so user rows stay exclusive to a worker -- kept in registers across a user's run, the next
// user's row fetched while the current rating is computed. Item rows stay in global memory (they are
// L2-resident: n_items * kp * 4 bytes is a few MB): a worker reads the row (L1 bypassed, one rating ahead),
// takes the gradient, and applies its step  q += lr * (g * p - reg * q)  as a vector atomic add executed by
// the L2 (red.global.add.v4.f32). Steps of different workers on the same item row are therefore all applied,
// in some order, none lost -- only the gradient may have been taken at a q that is a few updates old, as in
// the reference's lock-free NaiveParallelization mode (BiasedMatrixFactorization.cs:136-141, :201-204), but
// confined to a block: blocks of a sub-epoch stay disjoint as in the reference's DSGD mode. A popular item's
// chain of updates is serialised by the L2 atomic unit, not by the SM. No barriers inside the block.
// The read-only head of a worker's slice of block (j, slot): its entry range and first two entries. It does not depend on
// the model, so the persistent kernel fetches it one sub-epoch ahead (three dependent global-memory latencies off the
// critical path of every block hand-over).
struct AsyncHead {
    uint32_t e0, e1;
    int u1, i1, u2, i2;
    float v1, v2;
};

template <int L>
__device__ __forceinline__ AsyncHead async_head(const SgdArgs& a, const int j, const int sub, const int slot)
{
    constexpr int WPW = 32 / L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = sub * (int)(blockDim.x >> 5) * WPW + warp * WPW + lane / L;   // worker inside the group
    const int n_workers = a.n_workers;
    AsyncHead h;
    h.u1 = 0; h.i1 = 0; h.u2 = 0; h.i2 = 0; h.v1 = 0.f; h.v2 = 0.f;
    const uint32_t* wp = a.wptr + (size_t)(j * a.G + slot) * (n_workers + 1) + min(wid, n_workers - 1);
    h.e0 = wid < n_workers ? wp[0] : 0u; h.e1 = wid < n_workers ? wp[1] : 0u;
    if (h.e0 < h.e1) { h.u1 = a.ent_u[h.e0]; h.i1 = a.ent_i[h.e0]; h.v1 = a.ent_v[h.e0]; }
    if (h.e0 + 1 < h.e1) { h.u2 = a.ent_u[h.e0 + 1]; h.i2 = a.ent_i[h.e0 + 1]; h.v2 = a.ent_v[h.e0 + 1]; }
    return h;
}

template <int L, int KPL, bool BIASED>
__device__ __forceinline__ void sgd_block_async(const SgdArgs& a, const int j, const int slot, const AsyncHead& head)
{
    constexpr int KP = L * KPL;
    const int lane = threadIdx.x & 31;
    const int sl = lane % L;
    float* Qg = a.Q;            // entries hold internal item rows (strata_entries_kernel, absolute_rows)
    float* Bg = a.bi;
    const uint32_t e0 = head.e0, e1 = head.e1;
    // entries two ahead, user row and item row one ahead
    int u1 = head.u1, i1 = head.i1, u2 = head.u2, i2 = head.i2; float v1 = head.v1, v2 = head.v2;
    float p[KPL], pn[KPL], qn[KPL];
    float bu_v = 0.f, bun = 0.f, bin = 0.f, regu = a.reg_u, regun = a.reg_u;
#pragma unroll
    for (int f = 0; f < KPL; f++) { p[f] = 0.f; pn[f] = 0.f; qn[f] = 0.f; }
    if (e0 < e1) {
        Row<L, KPL>::load(pn, a.P + (size_t)u1 * KP, sl);
        Row<L, KPL>::load_cg(qn, Qg + (size_t)i1 * KP, sl);
        if (BIASED) { bun = a.bu[u1]; bin = ld_cg_f(Bg + i1); }
        if (a.regw_u) regun = a.regw_u[u1];
    }
    // the warp iterates until its longest worker is done
    uint32_t len = e1 - e0;
#pragma unroll
    for (int d = L; d < 32; d <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
    int cur_u = -1;
    for (uint32_t t = 0; t < len; t++) {
        const uint32_t e = e0 + t;
        const bool active = e < e1;
        const int u = u1, i = i1; const float v = v1;
        u1 = u2; i1 = i2; v1 = v2;
        if (e + 2 < e1) { u2 = a.ent_u[e + 2]; i2 = a.ent_i[e + 2]; v2 = a.ent_v[e + 2]; }
        if (active && u != cur_u) {   // new user run: flush the previous row, adopt the fetched one
            if (cur_u >= 0) {
                Row<L, KPL>::store(p, a.P + (size_t)cur_u * KP, sl);
                if (BIASED) a.bu[cur_u] = bu_v;
            }
#pragma unroll
            for (int f = 0; f < KPL; f++) p[f] = pn[f];
            bu_v = bun; regu = regun;
            cur_u = u;
        }
        float q[KPL];
#pragma unroll
        for (int f = 0; f < KPL; f++) q[f] = qn[f];
        const float bi0 = bin;
        const bool same_item = active && e + 1 < e1 && i1 == i;   // next rating hits the same item: forward the new row
        if (e + 1 < e1) {   // next entry: its item row, and its user row if a new run starts (runs are contiguous)
            if (!same_item) {
                Row<L, KPL>::load_cg(qn, Qg + (size_t)i1 * KP, sl);
                if (BIASED) bin = ld_cg_f(Bg + i1);
            }
            if (u1 != cur_u) {
                Row<L, KPL>::load(pn, a.P + (size_t)u1 * KP, sl);
                if (BIASED) bun = a.bu[u1];
                if (a.regw_u) regun = a.regw_u[u1];
            }
        }
        const float regi = a.regw_i ? a.regw_i[active ? i : 0] : a.reg_i;
        float pw[KPL], dq[KPL];
#pragma unroll
        for (int f = 0; f < KPL; f++) pw[f] = p[f];
        float buw = bu_v, dbi;
        sgd_update_delta<L, KPL, BIASED>(a, pw, q, dq, buw, bi0, dbi, v, regu, regi);
        if (active) {
#pragma unroll
            for (int f = 0; f < KPL; f++) p[f] = pw[f];
            bu_v = buw;
            float* qrow = Qg + (size_t)i * KP;
#pragma unroll
            for (int vv = 0; vv < KPL / 4; vv++)
                red_add_f4(qrow + 4 * (vv * L + sl), dq[4 * vv], dq[4 * vv + 1], dq[4 * vv + 2], dq[4 * vv + 3]);
            if (BIASED && sl == 0) red_add_f(Bg + i, dbi);
            if (same_item) {
#pragma unroll
                for (int f = 0; f < KPL; f++) qn[f] = q[f] + dq[f];
                bin = bi0 + dbi;
            }
        }
    }
    if (cur_u >= 0) {
        Row<L, KPL>::store(p, a.P + (size_t)cur_u * KP, sl);
        if (BIASED) a.bu[cur_u] = bu_v;
    }
}

// ---- the async block, second form ------------------------------------------------------------------------------------------
// Same schedule and semantics as sgd_block_async; what changed is the instruction stream (ncu of round 1: ~200 warp
// instructions per warp iteration, 16 warps per SM, issue slots half idle on scoreboard waits):
//   * factor arithmetic on packed pairs (fma.rn.f32x2 / mul.f32x2 of sm_100: one instruction per two factors);
//   * a worker past the end of its slice keeps running with zero coefficients (p <- 1 p + 0 q, no stores) instead of
//     computing into copies that are committed under a predicate;
//   * a just-updated item row is not forwarded through registers to a next rating on the same item (it only happens at run
//     boundaries); that rating's row is simply loaded after the step was added instead of one rating ahead;
// Measured on config 4 (profiles/r2_sgd_variants.log): 16.17 ms (first form) -> 15.78 ms (this form, 16 lanes x 8 floats) ->
// 14.69 ms (8 lanes x 16 floats: four ratings per warp instruction share the scalar part); cutting the registers to 64 for two
// CTAs per SM (no next-user row held ahead) gave 14.79 ms and was dropped. At k = 64 the 8 x 8 shape stays (4 x 16 is slower).
// A third form was measured and dropped (GPU calls o, p of round 2): rows requested TWO ratings ahead with cp.async.cg into a
// per-worker ring in shared memory (no row registers held ahead, no scoreboard shared with the entry loads; 96 registers,
// 198 KB of rings). Parity-green, but 15.27 ms against 14.34 ms on config 4 and 2.93 against 2.60 ms on one GPU-level
// sub-epoch of the 8-GPU ring: the loop is not waiting for a row that a longer request distance would have brought earlier,
// it queues on the SM's request path to the L2 (l1tex__m_l1tex2xbar_req_cycles_active 67 % of the whole kernel, hand-overs
// included), and the staged form adds requests (two 16-byte bias chunks per rating) to it.
template <int KPL>
struct Row2 {
    float2 r[KPL / 2];
};

template <int L, int KPL>
__device__ __forceinline__ void row2_load(Row2<KPL>& d, const float* row, int sl)
{
#pragma unroll
    for (int v = 0; v < KPL / 4; v++) {
        const float4 x = reinterpret_cast<const float4*>(row)[v * L + sl];
        d.r[2 * v] = make_float2(x.x, x.y); d.r[2 * v + 1] = make_float2(x.z, x.w);
    }
}
template <int L, int KPL>
__device__ __forceinline__ void row2_load_cg(Row2<KPL>& d, const float* row, int sl)
{
#pragma unroll
    for (int v = 0; v < KPL / 4; v++) {
        const float4 x = ld_cg_f4(reinterpret_cast<const float4*>(row) + (v * L + sl));
        d.r[2 * v] = make_float2(x.x, x.y); d.r[2 * v + 1] = make_float2(x.z, x.w);
    }
}
template <int L, int KPL>
__device__ __forceinline__ void row2_store(const Row2<KPL>& d, float* row, int sl)
{
#pragma unroll
    for (int v = 0; v < KPL / 4; v++)
        reinterpret_cast<float4*>(row)[v * L + sl] = make_float4(d.r[2 * v].x, d.r[2 * v].y, d.r[2 * v + 1].x, d.r[2 * v + 1].y);
}

// The item-row step as ONE bulk reduction per rating (BULK): the worker writes dq to a shared-memory buffer and lane 0 hands
// it to the TMA unit (cp.reduce.async.bulk ... .add.f32: the L2 adds 4 kp bytes to the row), instead of kp/4 vector atomics
// per lane. Why: REDG costs the SM's load/store path ~1.3 cycles PER LANE (B300 microarchitecture notes: 0.85 single address,
// 1.29 spread), i.e. ~41 cycles for each of the 4 warp-wide red.v4 of a 4-rating warp iteration = ~41 cycles per rating and
// SM -- which is the measured speed of the loop (20.3 ns = 39 cycles per rating and SM at config 4). Two buffers per worker:
// the bulk operation of iteration t reads its buffer while iteration t + 1 fills the other.
// Measured (GPU call r): parity-green, 14.67 ms against 14.34 ms at config 4, 2.68 against 2.60 ms on a ring sub-epoch -- no
// gain, so the per-lane atomics stay the default (MMLB200_SGD_VARIANT=2 selects this form): the loop is not bound by the SM's
// store path either. What binds is the L2 itself: a row read plus a row atomic (read-modify-write in the slice) per rating,
// see mml_ctx_probe_l2 and bench.py's roofline.binding. Of the ~11 L2 requests of a rating two are the 4-byte item-bias read
// and the 4-byte item-bias atomic (plain MF without biases: 12.6 ms against 14.2 ms, call ab). Keeping the biases of the
// block's item group in shared memory and adding the CTA's net change at the block end (call ac) gave 12.98 ms -- and a
// WRONG model (train RMSE of epoch 1 at 10M ratings: 0.814 against 0.735): with 8 CTAs per group every CTA pushes its own
// copy of a popular item's bias towards the target and the eight net changes are summed. The bias has to be shared by the
// whole group at every rating, i.e. stay in the L2. (Also measured and dropped, GPU call u: prefetch.global.L2 of the
// entry arrays 64 ratings ahead and of the next run's user row two ratings ahead -- 14.50 ms with, 14.26 ms without.)
__device__ __forceinline__ void bulk_red_add_f32(float* dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void sts_f4(uint32_t addr, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <int L, int KPL, bool BIASED, bool FREQW, bool BULK>
__device__ __forceinline__ void sgd_block_async2(const SgdArgs& a, const AsyncHead& head, const uint32_t stage)
{
    constexpr int KP = L * KPL;
    const int lane = threadIdx.x & 31;
    const int sl = lane % L;
    float* const Qg = a.Q;
    float* const Bg = a.bi;
    const uint32_t e0 = head.e0, e1 = head.e1;
    int u1 = head.u1, i1 = head.i1, u2 = head.u2, i2 = head.i2; float v1 = head.v1, v2 = head.v2;
    Row2<KPL> p, pn, qn;
    float bu_v = 0.f, bun = 0.f, bin = 0.f, regu = a.reg_u, regun = a.reg_u;
#pragma unroll
    for (int f = 0; f < KPL / 2; f++) { p.r[f] = make_float2(0.f, 0.f); pn.r[f] = p.r[f]; qn.r[f] = p.r[f]; }
    if (e0 < e1) {
        row2_load<L, KPL>(pn, a.P + (size_t)u1 * KP, sl);
        if (BIASED) bun = a.bu[u1];
        if (FREQW) regun = a.regw_u[u1];
        row2_load_cg<L, KPL>(qn, Qg + (size_t)i1 * KP, sl);
        if (BIASED) bin = ld_cg_f(Bg + i1);
    }
    uint32_t len = e1 - e0;
#pragma unroll
    for (int d = L; d < 32; d <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
    int cur_u = -1;
    for (uint32_t t = 0; t < len; t++) {
        const uint32_t e = e0 + t;
        const bool active = e < e1;
        const int u = u1, i = i1; const float v = v1;
        u1 = u2; i1 = i2; v1 = v2;
        if (e + 2 < e1) { u2 = a.ent_u[e + 2]; i2 = a.ent_i[e + 2]; v2 = a.ent_v[e + 2]; }
        if (active && u != cur_u) {   // new user run: flush the previous row, take the next one
            if (cur_u >= 0) {
                row2_store<L, KPL>(p, a.P + (size_t)cur_u * KP, sl);
                if (BIASED) a.bu[cur_u] = bu_v;
            }
#pragma unroll
            for (int f = 0; f < KPL / 2; f++) p.r[f] = pn.r[f];
            bu_v = bun; regu = regun;
            cur_u = u;
        }
        Row2<KPL> q;
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) q.r[f] = qn.r[f];
        const float bi0 = bin;
        // the next rating hits the same item row (a run boundary: consecutive users sharing an item): its row is read after
        // this rating's step has been added (same thread, same address: the red and the load stay in order)
        const bool same_item = active && e + 1 < e1 && i1 == i;
        if (e + 1 < e1) {   // next entry: its item row, and its user row if a new run starts
            if (!same_item) {
                row2_load_cg<L, KPL>(qn, Qg + (size_t)i1 * KP, sl);
                if (BIASED) bin = ld_cg_f(Bg + i1);
            }
            if (u1 != cur_u) {
                row2_load<L, KPL>(pn, a.P + (size_t)u1 * KP, sl);
                if (BIASED) bun = a.bu[u1];
                if (FREQW) regun = a.regw_u[u1];
            }
        }
        // (a template flag, not a run-time test of the pointer: a predicated-off load still ties its register to the scoreboard
        // it shares with the row loads above, and the first use of regi then waits for those)
        const float regi = FREQW ? a.regw_i[active ? i : 0] : a.reg_i;
        // dot product on packed pairs
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) acc = __ffma2_rn(p.r[f], q.r[f], acc);
        const float dot = worker_sum<L>(acc.x + acc.y);
        float gc, dbi = 0.f, bu_new = bu_v;
        if (BIASED) {
            const float score = ((a.gb + bu_v) + bi0) + dot;
            const float sig = __fdividef(1.f, 1.f + __expf(-score));
            const float err = v - (a.minr + sig * a.range);
            if (a.loss == MML_LOSS_RMSE) gc = err * sig * (1.f - sig) * a.range;
            else if (a.loss == MML_LOSS_MAE) gc = (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * sig * (1.f - sig) * a.range;
            else gc = err;
            const float step = a.blr * a.lr;
            bu_new = bu_v + step * (gc - a.breg * regu * bu_v);
            dbi = step * (gc - a.breg * regi * bi0);
        } else {
            gc = v - (a.gb + dot);
        }
        // p <- cu p + lg q ; dq = lg p - ci q   (pre-update p). Inactive workers: cu = 1, lg = 0 leave p as it is.
        const float lg = active ? a.lr * gc : 0.f;
        const float cu = active ? fmaf(-a.lr, regu, 1.f) : 1.f;
        const float nci = -(a.lr * regi);
        const float2 lg2 = make_float2(lg, lg), cu2 = make_float2(cu, cu), nci2 = make_float2(nci, nci);
        Row2<KPL> dq;
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) {
            dq.r[f] = __ffma2_rn(lg2, p.r[f], __fmul2_rn(nci2, q.r[f]));
            p.r[f] = __ffma2_rn(lg2, q.r[f], __fmul2_rn(cu2, p.r[f]));
        }
        if (BULK) {
            const uint32_t buf = stage + (t & 1u) * (uint32_t)(KP * 4);
            if (sl == 0) bulk_wait_read<1>();      // the operation that read this buffer two iterations ago is done with it
            __syncwarp();
#pragma unroll
            for (int vv = 0; vv < KPL / 4; vv++)
                sts_f4(buf + 16u * (uint32_t)(vv * L + sl), dq.r[2 * vv].x, dq.r[2 * vv].y, dq.r[2 * vv + 1].x, dq.r[2 * vv + 1].y);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (sl == 0) {
                if (active) bulk_red_add_f32(Qg + (size_t)i * KP, buf, KP * 4);
                bulk_commit();
            }
            if (active) {
                bu_v = bu_new;
                if (BIASED && sl == 0) red_add_f(Bg + i, dbi);
            }
            if (__any_sync(0xffffffffu, same_item)) {      // the next rating reads this row: after the step has landed
                if (sl == 0) bulk_wait<0>();
                __syncwarp();
                if (same_item) {
                    row2_load_cg<L, KPL>(qn, Qg + (size_t)i * KP, sl);
                    if (BIASED) bin = ld_cg_f(Bg + i);
                }
            }
        } else if (active) {
            bu_v = bu_new;
            float* qrow = Qg + (size_t)i * KP;
#pragma unroll
            for (int vv = 0; vv < KPL / 4; vv++)
                red_add_f4(qrow + 4 * (vv * L + sl), dq.r[2 * vv].x, dq.r[2 * vv].y, dq.r[2 * vv + 1].x, dq.r[2 * vv + 1].y);
            if (BIASED && sl == 0) red_add_f(Bg + i, dbi);
            if (same_item) {
                row2_load_cg<L, KPL>(qn, Qg + (size_t)i * KP, sl);
                if (BIASED) bin = ld_cg_f(Bg + i);
            }
        }
    }
    if (cur_u >= 0) {
        row2_store<L, KPL>(p, a.P + (size_t)cur_u * KP, sl);
        if (BIASED) a.bu[cur_u] = bu_v;
    }
    if (BULK) {     // every step of this block has been added before the block is handed over
        if (sl == 0) bulk_wait<0>();
        __syncwarp();
    }
}

// ---- NaiveParallelization (BiasedMatrixFactorization.cs:136-141, :201-204) -----------------------------------------------------
// The reference deals RandomIndex round-robin into MaxThreads lists (MultiCore.PartitionIndices) and lets every thread walk
// its list with no coordination: both rows of a rating may be under update by other threads. Here a list belongs to a worker
// (L lanes); the GPU runs n_lists = all its workers at once. A rating reads both rows L1-bypassed, and applies BOTH steps as
// vector atomic adds (the reference's racing read-modify-writes lose updates; the atomics do not). Entries are laid out list
// after list at build time (naive_entries_kernel), item and user rows as internal row numbers.
__global__ void naive_entries_kernel(const int32_t* __restrict__ ri, int64_t n, int32_t n_lists,
                                     const int32_t* __restrict__ users, const int32_t* __restrict__ items, const float* __restrict__ values,
                                     const int32_t* __restrict__ user_int, const int32_t* __restrict__ item_int,
                                     int32_t* __restrict__ ent_u, int32_t* __restrict__ ent_i, float* __restrict__ ent_v)
{
    const int64_t base = n / n_lists, rem = n % n_lists;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        const int64_t l = t % n_lists, pos = t / n_lists;           // MultiCore.cs:88-89
        const int64_t e = l * base + (l < rem ? l : rem) + pos;
        const int32_t src = ri[t];
        ent_u[e] = user_int[users[src]]; ent_i[e] = item_int[items[src]]; ent_v[e] = values[src];
    }
}

template <int L, int KPL, bool BIASED>
__global__ void __launch_bounds__(512) sgd_naive_kernel(const SgdArgs a, const int64_t n, const int32_t n_lists)
{
    constexpr int KP = L * KPL;
    const int lane = threadIdx.x & 31, sl = lane % L;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;          // list of this worker
    const int64_t base = n / n_lists, rem = n % n_lists;
    int64_t e0 = 0, e1 = 0;
    if (w < n_lists) { e0 = w * base + (w < rem ? w : rem); e1 = e0 + base + (w < rem ? 1 : 0); }
    uint32_t len = (uint32_t)(e1 - e0);
#pragma unroll
    for (int d = L; d < 32; d <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
    Row2<KPL> pn, qn;
    int un = 0, in = 0; float vn = 0.f, bun = 0.f, bin = 0.f;
#pragma unroll
    for (int f = 0; f < KPL / 2; f++) { pn.r[f] = make_float2(0.f, 0.f); qn.r[f] = pn.r[f]; }
    if (e0 < e1) {
        un = a.ent_u[e0]; in = a.ent_i[e0]; vn = a.ent_v[e0];
        row2_load_cg<L, KPL>(pn, a.P + (size_t)un * KP, sl);
        row2_load_cg<L, KPL>(qn, a.Q + (size_t)in * KP, sl);
        if (BIASED) { bun = ld_cg_f(a.bu + un); bin = ld_cg_f(a.bi + in); }
    }
    for (uint32_t t = 0; t < len; t++) {
        const int64_t e = e0 + t;
        const bool active = e < e1;
        const int u = un, i = in; const float v = vn;
        Row2<KPL> p, q;
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) { p.r[f] = pn.r[f]; q.r[f] = qn.r[f]; }
        const float bu0 = bun, bi0 = bin;
        if (e + 1 < e1) {                                         // the next rating's rows, one rating ahead
            un = a.ent_u[e + 1]; in = a.ent_i[e + 1]; vn = a.ent_v[e + 1];
            row2_load_cg<L, KPL>(pn, a.P + (size_t)un * KP, sl);
            row2_load_cg<L, KPL>(qn, a.Q + (size_t)in * KP, sl);
            if (BIASED) { bun = ld_cg_f(a.bu + un); bin = ld_cg_f(a.bi + in); }
        }
        const float regu = a.regw_u ? a.regw_u[active ? u : 0] : a.reg_u, regi = a.regw_i ? a.regw_i[active ? i : 0] : a.reg_i;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) acc = __ffma2_rn(p.r[f], q.r[f], acc);
        const float dot = worker_sum<L>(acc.x + acc.y);
        float gc, dbu = 0.f, dbi = 0.f;
        if (BIASED) {
            const float score = ((a.gb + bu0) + bi0) + dot;
            const float sig = __fdividef(1.f, 1.f + __expf(-score));
            const float err = v - (a.minr + sig * a.range);
            if (a.loss == MML_LOSS_RMSE) gc = err * sig * (1.f - sig) * a.range;
            else if (a.loss == MML_LOSS_MAE) gc = (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * sig * (1.f - sig) * a.range;
            else gc = err;
            const float step = a.blr * a.lr;
            dbu = step * (gc - a.breg * regu * bu0);
            dbi = step * (gc - a.breg * regi * bi0);
        } else {
            gc = v - (a.gb + dot);
        }
        if (!active) continue;
        const float lg = a.lr * gc, ncu = -(a.lr * regu), nci = -(a.lr * regi);
        const float2 lg2 = make_float2(lg, lg), ncu2 = make_float2(ncu, ncu), nci2 = make_float2(nci, nci);
        float* prow = a.P + (size_t)u * KP;
        float* qrow = a.Q + (size_t)i * KP;
#pragma unroll
        for (int vv = 0; vv < KPL / 4; vv++) {
            const float2 dp0 = __ffma2_rn(lg2, q.r[2 * vv], __fmul2_rn(ncu2, p.r[2 * vv])), dp1 = __ffma2_rn(lg2, q.r[2 * vv + 1], __fmul2_rn(ncu2, p.r[2 * vv + 1]));
            const float2 dq0 = __ffma2_rn(lg2, p.r[2 * vv], __fmul2_rn(nci2, q.r[2 * vv])), dq1 = __ffma2_rn(lg2, p.r[2 * vv + 1], __fmul2_rn(nci2, q.r[2 * vv + 1]));
            red_add_f4(prow + 4 * (vv * L + sl), dp0.x, dp0.y, dp1.x, dp1.y);
            red_add_f4(qrow + 4 * (vv * L + sl), dq0.x, dq0.y, dq1.x, dq1.y);
        }
        if (BIASED && sl == 0) { red_add_f(a.bu + u, dbu); red_add_f(a.bi + i, dbi); }
    }
}

// One launch per sub-epoch: grid = G CTAs.
template <int L, int KPL, bool BIASED, bool STAGE, bool ASYNC>
__global__ void __launch_bounds__(512) sgd_slot_kernel(const SgdArgs a, const int slot)
{
    extern __shared__ float4 smem4[];
    if (ASYNC) {
        const int j = blockIdx.x / a.cpg, sub = blockIdx.x % a.cpg;
        sgd_block_async<L, KPL, BIASED>(a, j, slot, async_head<L>(a, j, sub, slot));
    }
    else sgd_block<L, KPL, BIASED, STAGE>(a, blockIdx.x, slot, reinterpret_cast<float*>(smem4));
}

// One cooperative launch per epoch. A worker group is cpg CTAs (cpg = 1 unless async mode asks for more); group j
// walks the sub-epoch sequence; before it takes item group b every one of its CTAs waits (acquire) until all cpg
// CTAs of the group that held b in the previous sub-epoch have published it (release) -- one progress counter
// per CTA. All G * cpg CTAs are co-resident (cooperative launch), so the waits cannot deadlock. G = 1: no
// hand-over at all, the GPU is one worker group (the reference's lock-free NaiveParallelization inside the GPU).
template <int L, int KPL, bool BIASED, bool STAGE, bool ASYNC>
__global__ void __launch_bounds__(512) sgd_epoch_kernel(const SgdArgs a)
{
    extern __shared__ float4 smem4[];
    const int cpg = ASYNC ? a.cpg : 1;
    const int j = blockIdx.x / cpg, sub = blockIdx.x % cpg;
    AsyncHead head;
    if (ASYNC) head = async_head<L>(a, j, sub, a.seq[0]);
    for (int t = 0; t < a.G; t++) {
        const int slot = a.seq[t];
        AsyncHead next = head;
        if (ASYNC && t + 1 < a.G) next = async_head<L>(a, j, sub, a.seq[t + 1]);
        if (t > 0) {
            // item group b = (slot + j) % G was held in sub-epoch t-1 by group jp with (seq[t-1] + jp) % G == b
            int b = slot + j; if (b >= a.G) b -= a.G;
            int jp = b - a.seq[t - 1]; if (jp < 0) jp += a.G;
            // ... and, with several CTAs per group, by any CTA of this group: the worker slices of a block are cut by entry
            // count, so a user row this CTA takes now may have belonged to a sister CTA in the previous sub-epoch.
            const uint32_t want = a.epoch_base + (uint32_t)t;
            for (int x = (int)threadIdx.x; x < (cpg > 1 ? 2 * cpg : cpg); x += (int)blockDim.x) {
                const uint32_t* f = a.flags + (x < cpg ? jp * cpg + x : j * cpg + (x - cpg));
                while ((int32_t)(ld_relaxed_u32(f) - want) < 0) __nanosleep(20);
                __threadfence();   // acquire
            }
            __syncthreads();
        }
        if (ASYNC) sgd_block_async<L, KPL, BIASED>(a, j, slot, head);
        else sgd_block<L, KPL, BIASED, STAGE>(a, j, slot, reinterpret_cast<float*>(smem4));
        head = next;
        if (a.G > 1) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release_u32(a.flags + blockIdx.x, a.epoch_base + (uint32_t)t + 1u);
        }
    }
}

// The same two kernels on sgd_block_async2 (async mode only).
// shared address of this worker's two dq buffers (BULK)
template <int L, int KPL>
__device__ __forceinline__ uint32_t bulk_stage_base()
{
    extern __shared__ float4 smem4[];
    return (uint32_t)__cvta_generic_to_shared(smem4) + (uint32_t)(threadIdx.x / L) * (uint32_t)(2 * L * KPL * 4);
}

template <int L, int KPL, bool BIASED, bool FREQW, bool BULK>
__global__ void __launch_bounds__(512) sgd_slot2_kernel(const SgdArgs a, const int slot)
{
    const int j = blockIdx.x / a.cpg, sub = blockIdx.x % a.cpg;
    sgd_block_async2<L, KPL, BIASED, FREQW, BULK>(a, async_head<L>(a, j, sub, slot), BULK ? bulk_stage_base<L, KPL>() : 0u);
}

template <int L, int KPL, bool BIASED, bool FREQW, bool BULK>
__global__ void __launch_bounds__(512) sgd_epoch2_kernel(const SgdArgs a)
{
    const uint32_t stage = BULK ? bulk_stage_base<L, KPL>() : 0u;
    const int cpg = a.cpg;
    const int j = blockIdx.x / cpg, sub = blockIdx.x % cpg;
    const long long k0 = a.wait_stats ? clock64() : 0;
    long long waited = 0;
    AsyncHead head = async_head<L>(a, j, sub, a.seq[0]);
    for (int t = 0; t < a.G; t++) {
        const int slot = a.seq[t];
        AsyncHead next = head;
        if (t + 1 < a.G) next = async_head<L>(a, j, sub, a.seq[t + 1]);
        if (t > 0) {   // hand-over: see sgd_epoch_kernel
            int b = slot + j; if (b >= a.G) b -= a.G;
            int jp = b - a.seq[t - 1]; if (jp < 0) jp += a.G;
            const uint32_t want = a.epoch_base + (uint32_t)t;
            const long long w0 = a.wait_stats ? clock64() : 0;
            for (int x = (int)threadIdx.x; x < (cpg > 1 ? 2 * cpg : cpg); x += (int)blockDim.x) {
                const uint32_t* f = a.flags + (x < cpg ? jp * cpg + x : j * cpg + (x - cpg));
                while ((int32_t)(ld_relaxed_u32(f) - want) < 0) __nanosleep(20);
                __threadfence();   // acquire
            }
            __syncthreads();
            if (a.wait_stats) waited += clock64() - w0;
        }
        sgd_block_async2<L, KPL, BIASED, FREQW, BULK>(a, head, stage);
        head = next;
        if (a.G > 1) {
            const long long w0 = a.wait_stats ? clock64() : 0;
            __threadfence();
            __syncthreads();     // the CTA's slowest warp: also idle time of the others
            if (a.wait_stats && threadIdx.x == 0) waited += clock64() - w0;
            if (threadIdx.x == 0) st_release_u32(a.flags + blockIdx.x, a.epoch_base + (uint32_t)t + 1u);
        }
    }
    if (a.wait_stats && threadIdx.x == 0) {
        atomicAdd(a.wait_stats, (unsigned long long)waited);
        atomicAdd(a.wait_stats + 1, (unsigned long long)(clock64() - k0));
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
            const double pred = __dadd_rn((double)a.minr, __dmul_rn(sig, (double)a.range));   // no FMA contraction: the CLR does not fuse
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
                if (update_user) prow[s * 32 + lane] = __fadd_rn(pv[s], (float)__dmul_rn((double)a.lr, __dsub_rn(__dmul_rn((double)gc, vf), __dmul_rn((double)regu, uf))));
                if (update_item) qrow[s * 32 + lane] = __fadd_rn(qv[s], (float)__dmul_rn((double)a.lr, __dsub_rn(__dmul_rn((double)gc, uf), __dmul_rn((double)regi, vf))));
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

// BiasedMatrixFactorization.cs:313-325 / MatrixFactorization.cs:205-217,251-259.
// A warp takes 32 pairs at a time: the ids are read coalesced (lane = pair); the 32 dot products are formed by sub-warps of
// L lanes, each lane moving V 128-bit pieces of either row (a whole row is requested by one or two instructions per lane, so
// the 512 bytes of a k = 128 row reach the memory system together), two steps (2 * 32 / L pairs) in flight; and the scalar
// part -- the double-precision link, clipping and, in evaluate_kernel, the logarithms of the measures -- runs once per pair on
// the pair's own lane instead of L times. ROWS = true: users[] / items[] already hold internal factor rows (the strata
// entries, sorted by block and user: consecutive pairs share their user row, which then comes from the L1).
struct PairDots {
    int32_t urow, irow;     // internal rows of this lane's pair or -1 (unknown id)
    float dot;              // p_u . q_i of this lane's pair (0 unless both rows are known)
};

template <int L, int V, bool ROWS>
__device__ __forceinline__ PairDots pair_dots(const PredArgs& a, const int32_t* __restrict__ users,
                                              const int32_t* __restrict__ items, int64_t base, int64_t n, int lane)
{
    constexpr int NP = 32 / L;          // pairs per step
    constexpr int KP = 4 * L * V;
    PairDots r;
    r.urow = -1; r.irow = -1; r.dot = 0.f;
    if (base + lane < n) {
        const int32_t u = users[base + lane], i = items[base + lane];
        if (ROWS) { r.urow = u; r.irow = i; }
        else {
            if (u >= 0 && u < a.n_users_ext) r.urow = a.user_int[u];
            if (i >= 0 && i < a.n_items_ext) r.irow = a.item_int[i];
        }
    }
    const int cnt = (int)min((int64_t)32, n - base);
    const int sl = lane % L, g = lane / L;
    const int steps = (cnt + NP - 1) / NP;
    for (int s = 0; s < steps; s += 2) {
        float4 pa[2][V], qa[2][V];
#pragma unroll
        for (int x = 0; x < 2; x++) {
            const int pi = (s + x) * NP + g;                 // the pair this sub-warp takes in step s + x
            const int32_t ur = __shfl_sync(0xffffffffu, r.urow, pi & 31), ir = __shfl_sync(0xffffffffu, r.irow, pi & 31);
            const bool ok = pi < cnt && ur >= 0 && ir >= 0;
            const float4* pr = reinterpret_cast<const float4*>(a.P + (size_t)(ok ? ur : 0) * KP);
            const float4* qr = reinterpret_cast<const float4*>(a.Q + (size_t)(ok ? ir : 0) * KP);
#pragma unroll
            for (int v = 0; v < V; v++) {
                pa[x][v] = ok ? pr[v * L + sl] : make_float4(0.f, 0.f, 0.f, 0.f);
                qa[x][v] = ok ? qr[v * L + sl] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int x = 0; x < 2; x++) {
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int v = 0; v < V; v++) {
                d0 = fmaf(pa[x][v].x, qa[x][v].x, d0); d1 = fmaf(pa[x][v].y, qa[x][v].y, d1);
                d0 = fmaf(pa[x][v].z, qa[x][v].z, d0); d1 = fmaf(pa[x][v].w, qa[x][v].w, d1);
            }
            float d = d0 + d1;
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            // pair l was taken in step l / NP by sub-warp l % NP
            const float t = __shfl_sync(0xffffffffu, d, (lane % NP) * L);
            if (lane / NP == s + x && r.urow >= 0 && r.irow >= 0) r.dot = t;
        }
    }
    return r;
}

__device__ __forceinline__ float predict_from_dot(const PredArgs& a, const PairDots& r)
{
    const bool ku = r.urow >= 0, ki = r.irow >= 0;
    if (a.biased) {
        double score = a.gb;
        if (ku) score += a.bu[r.urow];
        if (ki) score += a.bi[r.irow];
        if (ku && ki) score += r.dot;
        return (float)__dadd_rn((double)a.minr, __dmul_rn(1.0 / (1.0 + exp(-score)), (double)a.range));
    }
    if (!ku || !ki) return a.gb;
    float res = a.gb + r.dot;
    if (res > a.maxr) res = a.maxr;
    if (res < a.minr) res = a.minr;
    return res;
}

template <int L, int V>
__global__ void predict_kernel(const PredArgs a, const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                               int64_t n, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    const int64_t stride = (((int64_t)gridDim.x * blockDim.x) >> 5) * 32;
    for (; base < n; base += stride) {
        const PairDots r = pair_dots<L, V, false>(a, users, items, base, n, lane);
        if (base + lane < n) out[base + lane] = predict_from_dot(a, r);
    }
}

// Eval/Ratings.cs:96-162. part[blk*4 + {0,1,2,3}] = sum err^2, sum |err|, sum CBD, sum objective loss
constexpr int EV_THREADS = 256;
template <int L, int V, bool ROWS>
__global__ void __launch_bounds__(EV_THREADS) evaluate_kernel(const PredArgs a, const int32_t* __restrict__ users,
                                                              const int32_t* __restrict__ items, const float* __restrict__ values,
                                                              int64_t n, int32_t loss, double* __restrict__ part)
{
    __shared__ double sh[EV_THREADS / 32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    const int64_t stride = (((int64_t)gridDim.x * blockDim.x) >> 5) * 32;
    for (; base < n; base += stride) {
        const PairDots pd = pair_dots<L, V, ROWS>(a, users, items, base, n, lane);
        if (base + lane >= n) continue;
        const float pr = predict_from_dot(a, pd);
        const float r = values[base + lane];
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
            // Eval/Measures/LogisticLoss.cs:37-57 divides by rating_range_size as the model holds it: 0 while InitModel
            // computes the bold driver's first last_loss (BiasedMatrixFactorization.cs:168-169 runs before :186), which makes
            // that loss NaN and the first UpdateLearnRate comparison a no-op -- kept by using a.range, not max - min
            double pl = ((double)pr - (double)a.minr) / (double)a.range;
            const double al = ((double)r - (double)a.minr) / (double)a.range;
            if (pl < 0.0) pl = 0.0;
            if (pl > 1.0) pl = 1.0;
            s3 -= al * log(pl);
            s3 -= (1.0 - al) * log(1.0 - pl);
        }
    }
    // lanes hold their own pairs' sums: butterfly over the warp, then over the block's warps in a fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
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
    a.round_ptr = m.round_ptr.p;
    a.wptr = m.wptr.p ? m.wptr.p + (size_t)B * m.G * m.G * (m.n_workers + 1) : nullptr;
    a.blk_round_ptr = m.blk_round_ptr.p ? m.blk_round_ptr.p + (size_t)B * m.G * m.G : nullptr;
    a.item_ptr = m.d_item_ptr.p ? m.d_item_ptr.p + (size_t)B * m.G : nullptr;
    a.hot_cnt = m.d_hot_cnt.p ? m.d_hot_cnt.p + (size_t)B * m.G : nullptr;
    a.regw_u = m.p.frequency_regularization ? m.regw_u.p : nullptr;
    a.regw_i = m.p.frequency_regularization ? m.regw_i.p : nullptr;
    a.flags = m.flags.p; a.seq = nullptr; a.epoch_base = m.epoch_base;
    a.wait_stats = m.wait_stats.p;
    a.G = m.G; a.C = m.hot_copies; a.n_workers = m.n_workers; a.cpg = std::max(m.cpg, 1);
    a.hot_scale = m.p.hot_merge_average ? 1.f / (float)m.hot_copies : 1.f;
    a.lr = m.lr; a.gb = m.global_bias; a.minr = m.min_rating; a.range = m.range;
    if (m.p.biased) { a.reg_u = m.p.reg_u; a.reg_i = m.p.reg_i; }
    else { a.reg_u = m.p.regularization; a.reg_i = m.p.regularization; }
    a.blr = m.p.bias_learn_rate; a.breg = m.p.bias_reg; a.loss = m.p.loss;
    return a;
}

typedef void (*slot_fn_t)(const SgdArgs, const int);
typedef void (*epoch_fn_t)(const SgdArgs);

// ---- owned-users epoch: no block hand-overs ------------------------------------------------------------------------------------
// In sgd_epoch2_kernel 12-14 % of the CTA time (19 % on the short sub-epochs of the multi-GPU ring) goes to the hand-over
// between sub-epochs: every CTA waits for the CTAs that held its next item group and for its sister CTAs (a user row may
// move between sister CTAs from block to block), then for its own slowest warp. None of that protects anything the async mode
// needs: item rows are only ever touched by atomics, so two groups on one item group cost nothing but a little more staleness;
// what must stay exclusive is a USER row. Here users are pinned to workers for the whole epoch (pin_users_to_workers): worker
// w of group j owns the same users in every block (j, *), its slices of the G blocks -- taken in the epoch's sub-epoch order --
// form ONE stream of ratings, and the kernel is a single loop over that stream: no flags, no barriers, no cooperative launch.
// The groups still walk the item groups in the DSGD order (s + j) mod G, in step only as far as equal loads keep them in step
// (a worker's load is balanced over the whole epoch, not per block): the "blocks of a sub-epoch are disjoint" of the reference
// becomes approximate, the per-block run structure (what the RMSE gate is sensitive to) is unchanged.
// Shared memory: per worker the G slice starts and the G + 1 prefix sums of their lengths.
template <int L, int KPL, bool BIASED, bool FREQW>
__global__ void __launch_bounds__(512) sgd_free_kernel(const SgdArgs a)
{
    extern __shared__ float4 smem4[];
    constexpr int KP = L * KPL;
    constexpr int WPW = 32 / L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sl = lane % L;
    const int cpg = a.cpg, G = a.G;
    const int j = blockIdx.x / cpg, sub = blockIdx.x % cpg;
    const int wl = warp * WPW + lane / L;                                   // worker inside the CTA
    const int wid = sub * (int)(blockDim.x >> 5) * WPW + wl;                 // worker inside the group
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem4) + (size_t)wl * (2 * G + 1);
    uint32_t* cum = tab;                                                     // [G + 1]
    uint32_t* base = tab + G + 1;                                            // [G]
    const bool has = wid < a.n_workers;
    for (int t = sl; t < G; t += L) {
        const uint32_t* wp = a.wptr + (size_t)(j * G + a.seq[t]) * (a.n_workers + 1) + (has ? wid : 0);
        const uint32_t b0 = wp[0], b1 = has ? wp[1] : b0;
        base[t] = b0; cum[t + 1] = b1 - b0;
    }
    __syncwarp();
    if (sl == 0) {
        uint32_t acc = 0;
        cum[0] = 0;
        for (int t = 0; t < G; t++) { acc += cum[t + 1]; cum[t + 1] = acc; }
    }
    __syncwarp();
    const uint32_t total = cum[G];
    // cursor of the entry fetches (two ratings ahead of the one being computed)
    int kf = 0;
    uint32_t c0 = 0, c1 = cum[1], bs = base[0];
    auto entry_of = [&](uint32_t pos) -> uint32_t {                          // pos < total, non-decreasing from call to call
        while (pos >= c1) { kf++; c0 = c1; c1 = cum[kf + 1]; bs = base[kf]; }
        return bs + (pos - c0);
    };
    float* const Qg = a.Q;
    float* const Bg = a.bi;
    int u1 = 0, i1 = 0, u2 = 0, i2 = 0; float v1 = 0.f, v2 = 0.f;
    if (0 < total) { const uint32_t e = entry_of(0); u1 = a.ent_u[e]; i1 = a.ent_i[e]; v1 = a.ent_v[e]; }
    if (1 < total) { const uint32_t e = entry_of(1); u2 = a.ent_u[e]; i2 = a.ent_i[e]; v2 = a.ent_v[e]; }
    Row2<KPL> p, pn, qn;
    float bu_v = 0.f, bun = 0.f, bin = 0.f, regu = a.reg_u, regun = a.reg_u;
#pragma unroll
    for (int f = 0; f < KPL / 2; f++) { p.r[f] = make_float2(0.f, 0.f); pn.r[f] = p.r[f]; qn.r[f] = p.r[f]; }
    if (0 < total) {
        row2_load<L, KPL>(pn, a.P + (size_t)u1 * KP, sl);
        if (BIASED) bun = a.bu[u1];
        if (FREQW) regun = a.regw_u[u1];
        row2_load_cg<L, KPL>(qn, Qg + (size_t)i1 * KP, sl);
        if (BIASED) bin = ld_cg_f(Bg + i1);
    }
    uint32_t len = total;
#pragma unroll
    for (int d = L; d < 32; d <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
    int cur_u = -1;
    for (uint32_t t = 0; t < len; t++) {
        const bool active = t < total;
        const int u = u1, i = i1; const float v = v1;
        u1 = u2; i1 = i2; v1 = v2;
        if (t + 2 < total) { const uint32_t e = entry_of(t + 2); u2 = a.ent_u[e]; i2 = a.ent_i[e]; v2 = a.ent_v[e]; }
        if (active && u != cur_u) {   // new user run: flush the previous row, take the next one
            if (cur_u >= 0) {
                row2_store<L, KPL>(p, a.P + (size_t)cur_u * KP, sl);
                if (BIASED) a.bu[cur_u] = bu_v;
            }
#pragma unroll
            for (int f = 0; f < KPL / 2; f++) p.r[f] = pn.r[f];
            bu_v = bun; regu = regun;
            cur_u = u;
        }
        Row2<KPL> q;
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) q.r[f] = qn.r[f];
        const float bi0 = bin;
        const bool same_item = active && t + 1 < total && i1 == i;     // see sgd_block_async2
        if (t + 1 < total) {
            if (!same_item) {
                row2_load_cg<L, KPL>(qn, Qg + (size_t)i1 * KP, sl);
                if (BIASED) bin = ld_cg_f(Bg + i1);
            }
            if (u1 != cur_u) {
                row2_load<L, KPL>(pn, a.P + (size_t)u1 * KP, sl);
                if (BIASED) bun = a.bu[u1];
                if (FREQW) regun = a.regw_u[u1];
            }
        }
        const float regi = FREQW ? a.regw_i[active ? i : 0] : a.reg_i;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) acc = __ffma2_rn(p.r[f], q.r[f], acc);
        const float dot = worker_sum<L>(acc.x + acc.y);
        float gc, dbi = 0.f, bu_new = bu_v;
        if (BIASED) {
            const float score = ((a.gb + bu_v) + bi0) + dot;
            const float sig = __fdividef(1.f, 1.f + __expf(-score));
            const float err = v - (a.minr + sig * a.range);
            if (a.loss == MML_LOSS_RMSE) gc = err * sig * (1.f - sig) * a.range;
            else if (a.loss == MML_LOSS_MAE) gc = (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * sig * (1.f - sig) * a.range;
            else gc = err;
            const float step = a.blr * a.lr;
            bu_new = bu_v + step * (gc - a.breg * regu * bu_v);
            dbi = step * (gc - a.breg * regi * bi0);
        } else {
            gc = v - (a.gb + dot);
        }
        const float lg = active ? a.lr * gc : 0.f;
        const float cu = active ? fmaf(-a.lr, regu, 1.f) : 1.f;
        const float nci = -(a.lr * regi);
        const float2 lg2 = make_float2(lg, lg), cu2 = make_float2(cu, cu), nci2 = make_float2(nci, nci);
        Row2<KPL> dq;
#pragma unroll
        for (int f = 0; f < KPL / 2; f++) {
            dq.r[f] = __ffma2_rn(lg2, p.r[f], __fmul2_rn(nci2, q.r[f]));
            p.r[f] = __ffma2_rn(lg2, q.r[f], __fmul2_rn(cu2, p.r[f]));
        }
        if (active) {
            bu_v = bu_new;
            float* qrow = Qg + (size_t)i * KP;
#pragma unroll
            for (int vv = 0; vv < KPL / 4; vv++)
                red_add_f4(qrow + 4 * (vv * L + sl), dq.r[2 * vv].x, dq.r[2 * vv].y, dq.r[2 * vv + 1].x, dq.r[2 * vv + 1].y);
            if (BIASED && sl == 0) red_add_f(Bg + i, dbi);
            if (same_item) {
                row2_load_cg<L, KPL>(qn, Qg + (size_t)i * KP, sl);
                if (BIASED) bin = ld_cg_f(Bg + i);
            }
        }
    }
    if (cur_u >= 0) {
        row2_store<L, KPL>(p, a.P + (size_t)cur_u * KP, sl);
        if (BIASED) a.bu[cur_u] = bu_v;
    }
}

template <int L, int KPL>
static epoch_fn_t pick_free_kernel(bool biased, bool freqw)
{
    if (biased) return freqw ? sgd_free_kernel<L, KPL, true, true> : sgd_free_kernel<L, KPL, true, false>;
    return freqw ? sgd_free_kernel<L, KPL, false, true> : sgd_free_kernel<L, KPL, false, false>;
}
static epoch_fn_t get_free_kernel(const Sgd& m)
{
    const bool b = m.p.biased != 0, fw = m.p.frequency_regularization != 0;
    switch (m.kp) {
        case 32: return pick_free_kernel<8, 4>(b, fw);
        case 64: return pick_free_kernel<8, 8>(b, fw);
        case 128: return pick_free_kernel<8, 16>(b, fw);
        case 256: return pick_free_kernel<32, 8>(b, fw);
    }
    return nullptr;
}

template <int L, int KPL, bool ASYNC>
static void pick_kernels2(bool biased, bool stage, slot_fn_t* sf, epoch_fn_t* ef)
{
    if (biased) {
        if (stage) { *sf = sgd_slot_kernel<L, KPL, true, true, ASYNC>; *ef = sgd_epoch_kernel<L, KPL, true, true, ASYNC>; }
        else { *sf = sgd_slot_kernel<L, KPL, true, false, ASYNC>; *ef = sgd_epoch_kernel<L, KPL, true, false, ASYNC>; }
    } else {
        if (stage) { *sf = sgd_slot_kernel<L, KPL, false, true, ASYNC>; *ef = sgd_epoch_kernel<L, KPL, false, true, ASYNC>; }
        else { *sf = sgd_slot_kernel<L, KPL, false, false, ASYNC>; *ef = sgd_epoch_kernel<L, KPL, false, false, ASYNC>; }
    }
}

template <int L, int KPL>
static void pick_kernels(bool async, bool biased, bool stage, slot_fn_t* sf, epoch_fn_t* ef)
{
    if (async) pick_kernels2<L, KPL, true>(biased, stage, sf, ef);
    else pick_kernels2<L, KPL, false>(biased, stage, sf, ef);
}

template <int L, int KPL, bool BULK>
static void pick_kernels_v2(bool biased, bool freqw, slot_fn_t* sf, epoch_fn_t* ef)
{
    if (biased) {
        if (freqw) { *sf = sgd_slot2_kernel<L, KPL, true, true, BULK>; *ef = sgd_epoch2_kernel<L, KPL, true, true, BULK>; }
        else { *sf = sgd_slot2_kernel<L, KPL, true, false, BULK>; *ef = sgd_epoch2_kernel<L, KPL, true, false, BULK>; }
    } else {
        if (freqw) { *sf = sgd_slot2_kernel<L, KPL, false, true, BULK>; *ef = sgd_epoch2_kernel<L, KPL, false, true, BULK>; }
        else { *sf = sgd_slot2_kernel<L, KPL, false, false, BULK>; *ef = sgd_epoch2_kernel<L, KPL, false, false, BULK>; }
    }
}

// Async epoch kernel: which loop (Sgd::variant: 0 = sgd_block_async, 1 = sgd_block_async2) and how many lanes per worker
// for padded row length kp. Default: the second form everywhere, 8 lanes x 16 floats at kp = 128.
static int async_lanes(int kp, int variant)
{
    if (variant == 0) return kp <= 64 ? 8 : (kp == 128 ? 16 : 32);
    return kp <= 128 ? 8 : 32;
}

// kp = 32: 8 lanes x 4 floats (4 ratings per warp) ; 64: 8 x 8 ; 128: 16 x 8 (first form) or 8 x 16 ; 256: 32 x 8
static int32_t get_kernels(Sgd& m, slot_fn_t* sf, epoch_fn_t* ef)
{
    const bool stage = m.stage_bytes > 0, async = m.p.intra_block == MML_INTRA_ASYNC;
    if (async && m.variant > 0) {
        const bool b = m.p.biased != 0, fw = m.p.frequency_regularization != 0;
        if (m.variant >= 2) {
            switch (m.kp) {
                case 32: pick_kernels_v2<8, 4, true>(b, fw, sf, ef); break;
                case 64: pick_kernels_v2<8, 8, true>(b, fw, sf, ef); break;
                case 128: pick_kernels_v2<8, 16, true>(b, fw, sf, ef); break;
                case 256: pick_kernels_v2<32, 8, true>(b, fw, sf, ef); break;
                default: set_error("unsupported num_factors"); return MML_ERR_UNSUPPORTED;
            }
            MML_CUDA(cudaFuncSetAttribute((const void*)*sf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.stage_bytes));
            MML_CUDA(cudaFuncSetAttribute((const void*)*ef, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.stage_bytes));
            return MML_OK;
        }
        switch (m.kp) {
            case 32: pick_kernels_v2<8, 4, false>(b, fw, sf, ef); break;
            case 64: pick_kernels_v2<8, 8, false>(b, fw, sf, ef); break;
            case 128: pick_kernels_v2<8, 16, false>(b, fw, sf, ef); break;
            case 256: pick_kernels_v2<32, 8, false>(b, fw, sf, ef); break;
            default: set_error("unsupported num_factors"); return MML_ERR_UNSUPPORTED;
        }
        return MML_OK;
    }
    switch (m.kp) {
        case 32: pick_kernels<8, 4>(async, m.p.biased != 0, stage, sf, ef); break;
        case 64: pick_kernels<8, 8>(async, m.p.biased != 0, stage, sf, ef); break;
        case 128: pick_kernels<16, 8>(async, m.p.biased != 0, stage, sf, ef); break;
        case 256: pick_kernels<32, 8>(async, m.p.biased != 0, stage, sf, ef); break;
        default: set_error("unsupported num_factors"); return MML_ERR_UNSUPPORTED;
    }
    if (stage) {
        MML_CUDA(cudaFuncSetAttribute((const void*)*sf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.stage_bytes));
        MML_CUDA(cudaFuncSetAttribute((const void*)*ef, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.stage_bytes));
    }
    return MML_OK;
}

// Multi-GPU: after an epoch every rank holds the current rows of its home item block only; prediction and model
// download need all of them -> every home block is broadcast from its rank (one grouped NCCL call).
static int32_t sync_items(Sgd& m)
{
    if (m.R <= 1 || !m.items_dirty) return MML_OK;
    MML_TRY(dist_group_start());
    for (int B = 0; B < m.RB; B++) {        // block B is at home on rank B / split after an epoch
        const int32_t lo = m.h_item_ptr[(size_t)B * m.G], hi = m.h_item_ptr[(size_t)(B + 1) * m.G];
        MML_TRY(dist_broadcast_f32(m.ctx, m.Q.p + (size_t)lo * m.kp, (size_t)(hi - lo) * m.kp, B / m.split));
        MML_TRY(dist_broadcast_f32(m.ctx, m.bi.p + lo, (size_t)(hi - lo), B / m.split));
    }
    MML_TRY(dist_group_end());
    m.items_dirty = false;
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
template <bool ROWS>
static void launch_evaluate(Sgd& m, int blocks, const int32_t* d_u, const int32_t* d_i, const float* d_v, int64_t n, double* part)
{
    cudaStream_t s = m.ctx->stream;
    const PredArgs a = make_pred_args(m);
    switch (m.kp) {
        case 32: evaluate_kernel<8, 1, ROWS><<<blocks, EV_THREADS, 0, s>>>(a, d_u, d_i, d_v, n, m.p.loss, part); break;
        case 64: evaluate_kernel<16, 1, ROWS><<<blocks, EV_THREADS, 0, s>>>(a, d_u, d_i, d_v, n, m.p.loss, part); break;
        case 128: evaluate_kernel<16, 2, ROWS><<<blocks, EV_THREADS, 0, s>>>(a, d_u, d_i, d_v, n, m.p.loss, part); break;
        default: evaluate_kernel<32, 2, ROWS><<<blocks, EV_THREADS, 0, s>>>(a, d_u, d_i, d_v, n, m.p.loss, part); break;
    }
}

// rows = true: d_u / d_i hold internal factor rows (the async strata entries) instead of ids
static int32_t evaluate_device(Sgd& m, const int32_t* d_u, const int32_t* d_i, const float* d_v, int64_t n, double* sums4,
                               bool rows = false)
{
    cudaStream_t s = m.ctx->stream;
    MML_TRY(sync_items(m));
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(n * 32, EV_THREADS * 4), 1), 148 * 8);
    DevBuf<double>& part = m.scr_part;
    if (part.n < (size_t)blocks * 4) MML_TRY(part.alloc((size_t)148 * 8 * 4));
    if (rows) launch_evaluate<true>(m, blocks, d_u, d_i, d_v, n, part.p);
    else launch_evaluate<false>(m, blocks, d_u, d_i, d_v, n, part.p);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    std::vector<double> h((size_t)blocks * 4);
    MML_CUDA(cudaMemcpyAsync(h.data(), part.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < 4; c++) sums4[c] = 0;
    for (int b = 0; b < blocks; b++) for (int c = 0; c < 4; c++) sums4[c] += h[(size_t)b * 4 + c];
    return MML_OK;
}

// The training set: through the async strata entries when there are any -- (internal user row, internal item row, value)
// sorted by block and user, i.e. no id translation and a user row shared by consecutive pairs -- else in COO order.
static int32_t evaluate_train_device(Sgd& m, double* sums4)
{
    const bool rows = m.p.schedule == MML_SCHEDULE_DSGD && m.p.intra_block == MML_INTRA_ASYNC && m.ent_u.p != nullptr;
    if (rows) return evaluate_device(m, m.ent_u.p, m.ent_i.p, m.ent_v.p, m.ratings->n, sums4, true);
    return evaluate_device(m, m.ratings->users.p, m.ratings->items.p, m.ratings->values.p, m.ratings->n, sums4);
}

// Sum over ranks of per-rank partial sums (and of the rating count n)
static int32_t reduce_over_ranks(Sgd& m, double* vals, int count)
{
    if (m.R <= 1) return MML_OK;
    cudaStream_t s = m.ctx->stream;
    DevBuf<double> d;
    MML_TRY(d.alloc(count));
    MML_CUDA(cudaMemcpyAsync(d.p, vals, sizeof(double) * count, cudaMemcpyHostToDevice, s));
    MML_TRY(dist_allreduce_f64(m.ctx, d.p, count));
    MML_CUDA(cudaMemcpyAsync(vals, d.p, sizeof(double) * count, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
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
                                                 user_side ? m.ratings->count_by_user.p : m.item_counts.p,
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
    MML_TRY(evaluate_train_device(m, sums));
    double ru = 0, ri = 0;
    MML_TRY(regterm(m, true, &ru));
    MML_TRY(regterm(m, false, &ri));
    double local[1] = { sums[3] + ru };     // loss and user terms are per rank, item terms are global
    MML_TRY(reduce_over_ranks(m, local, 1));
    *out = local[0] + ri;
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

// NaiveParallelization epoch: lists cut from RandomIndex (d_index), one per worker of the GPU
static int32_t run_naive_epoch(Sgd& m, bool rebuild)
{
    Ratings& r = *m.ratings;
    cudaStream_t s = m.ctx->stream;
    const int64_t n = r.n;
    if (n <= 0) return MML_OK;
    const int lanes = async_lanes(m.kp, 1);
    const int64_t workers = (int64_t)m.ctx->sm_count * (512 / lanes);
    // MultiCore.cs:81: min(num_groups, Count) lists; here num_groups = the GPU's workers, but no list shorter than 256
    // ratings: a small data set walked by thousands of lists at once is far staler than the reference's handful of threads
    const int32_t n_lists = (int32_t)std::max<int64_t>(1, std::min<int64_t>(workers, n / 256));
    if (rebuild || m.ent_u.p == nullptr) {
        MML_TRY(m.ent_u.ensure(n)); MML_TRY(m.ent_i.ensure(n)); MML_TRY(m.ent_v.ensure(n));
        naive_entries_kernel<<<grid_n(n), 256, 0, s>>>(m.d_index.p, n, n_lists, r.users.p, r.items.p, r.values.p,
                                                       m.users.d_to_int.p, m.items.d_to_int.p, m.ent_u.p, m.ent_i.p, m.ent_v.p);
        MML_CUDA(cudaGetLastError());
        m.launches++;
    }
    const SgdArgs a = make_args(m, 0);
    const int blocks = (int)ceil_div((int64_t)n_lists * lanes, 512);
    const bool b = m.p.biased != 0;
    switch (m.kp) {
        case 32: if (b) sgd_naive_kernel<8, 4, true><<<blocks, 512, 0, s>>>(a, n, n_lists); else sgd_naive_kernel<8, 4, false><<<blocks, 512, 0, s>>>(a, n, n_lists); break;
        case 64: if (b) sgd_naive_kernel<8, 8, true><<<blocks, 512, 0, s>>>(a, n, n_lists); else sgd_naive_kernel<8, 8, false><<<blocks, 512, 0, s>>>(a, n, n_lists); break;
        case 128: if (b) sgd_naive_kernel<8, 16, true><<<blocks, 512, 0, s>>>(a, n, n_lists); else sgd_naive_kernel<8, 16, false><<<blocks, 512, 0, s>>>(a, n, n_lists); break;
        default: if (b) sgd_naive_kernel<32, 8, true><<<blocks, 512, 0, s>>>(a, n, n_lists); else sgd_naive_kernel<32, 8, false><<<blocks, 512, 0, s>>>(a, n, n_lists); break;
    }
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
    // GPU-level sub-epochs -- the reference's block schedule (BiasedMatrixFactorization.cs:213-214) with GPUs in place of
    // threads. The item matrix is cut into RB = split x R blocks (split = 1, or 2 with MMLB200_RING_SPLIT=2); rank r starts the epoch holding its
    // home blocks split*r .. split*r + split-1. In sub-epoch S' it trains on block B = (S' + split*r) mod RB (one persistent
    // launch) and then, ON A SECOND STREAM, sends B to rank r - 1 and receives block B + split from rank r + 1 -- the block it
    // needs in sub-epoch S' + split, which rank r + 1 has just finished. With split = 2 a block therefore has a whole
    // sub-epoch to travel (and a late neighbour a whole sub-epoch of slack) while this rank trains on its other block; the
    // launch of S' + 2 waits for the exchange of S' only. After RB sub-epochs every block is back on its home rank.
    // (split = 1: the block needed next is finished by the neighbour at the same moment -- the exchange cannot overlap.)
    static const bool trace = [] { const char* e = getenv("MMLB200_TRACE"); return e && *e && *e != '0'; }();
    const int RB = m.RB, split = m.split;
    cudaStream_t xs = m.R > 1 ? m.ctx->aux_stream : s;
    std::vector<cudaEvent_t> tev;
    if (trace && m.R > 1) { tev.resize((size_t)2 * RB + 1); for (auto& e : tev) cudaEventCreate(&e); }
    if (!tev.empty()) cudaEventRecord(tev[2 * RB], s);
    for (int S = 0; S < RB; S++) {
        const int B = (S + split * m.rank) % RB;
        SgdArgs a = make_args(m, B);
        if (m.R > 1 && S >= split) MML_CUDA(cudaStreamWaitEvent(s, m.ev_xchg[S % split], 0));     // block B has arrived
        if (!tev.empty()) cudaEventRecord(tev[2 * S], s);
        if (m.owned) {                       // owned-users epoch: one plain launch, no hand-overs
            DevBuf<int32_t>& dseq = m.d_index;
            if ((int64_t)dseq.n < m.G) MML_TRY(dseq.alloc(m.G));
            MML_CUDA(cudaMemcpyAsync(dseq.p, seq.data(), sizeof(int32_t) * m.G, cudaMemcpyHostToDevice, s));
            a.seq = dseq.p;
            epoch_fn_t fk = get_free_kernel(m);
            fk<<<m.G * m.cpg, threads, m.free_smem, s>>>(a);
            MML_CUDA(cudaGetLastError());
            m.launches++;
        } else if (m.p.persistent) {
            DevBuf<int32_t>& dseq = m.d_index;   // reuse: serial index cache is unused in DSGD mode
            if ((int64_t)dseq.n < m.G) MML_TRY(dseq.alloc(m.G));
            MML_CUDA(cudaMemcpyAsync(dseq.p, seq.data(), sizeof(int32_t) * m.G, cudaMemcpyHostToDevice, s));
            a.seq = dseq.p;
            void* kargs[] = { (void*)&a };
            MML_CUDA(cudaLaunchCooperativeKernel((const void*)ef, dim3(m.G * m.cpg), dim3(threads), kargs, m.stage_bytes, s));
            m.epoch_base += (uint32_t)m.G + 1u;
            m.launches++;
        } else {
            for (int t = 0; t < m.G; t++) {
                sf<<<m.G * m.cpg, threads, m.stage_bytes, s>>>(a, seq[t]);
                m.launches++;
            }
            MML_CUDA(cudaGetLastError());
        }
        if (!tev.empty()) cudaEventRecord(tev[2 * S + 1], s);
        if (m.R > 1) {
            MML_CUDA(cudaEventRecord(m.ev_kern[S % split], s));
            MML_CUDA(cudaStreamWaitEvent(xs, m.ev_kern[S % split], 0));
            const int Bn = (B + split) % RB;
            const int32_t s_lo = m.h_item_ptr[(size_t)B * m.G], s_hi = m.h_item_ptr[(size_t)(B + 1) * m.G];
            const int32_t r_lo = m.h_item_ptr[(size_t)Bn * m.G], r_hi = m.h_item_ptr[(size_t)(Bn + 1) * m.G];
            MML_TRY(dist_ring_exchange(m.ctx, xs, m.Q.p + (size_t)s_lo * m.kp, (size_t)(s_hi - s_lo) * m.kp, m.bi.p + s_lo, (size_t)(s_hi - s_lo),
                                       (m.rank + m.R - 1) % m.R,
                                       m.Q.p + (size_t)r_lo * m.kp, (size_t)(r_hi - r_lo) * m.kp, m.bi.p + r_lo, (size_t)(r_hi - r_lo),
                                       (m.rank + 1) % m.R));
            MML_CUDA(cudaEventRecord(m.ev_xchg[S % split], xs));
        }
    }
    if (m.R > 1)       // the last exchanges bring the home blocks back: the epoch (and its timing) ends when they have landed
        for (int x = 0; x < split; x++) MML_CUDA(cudaStreamWaitEvent(s, m.ev_xchg[x], 0));
    if (!tev.empty()) {
        cudaEvent_t e_end;
        cudaEventCreate(&e_end);
        cudaEventRecord(e_end, s);
        cudaStreamSynchronize(s);
        float tk = 0.f, tot = 0.f;
        for (int S = 0; S < RB; S++) { float x = 0.f; cudaEventElapsedTime(&x, tev[2 * S], tev[2 * S + 1]); tk += x; }
        cudaEventElapsedTime(&tot, tev[2 * RB], e_end);
        fprintf(stderr, "[mmlb200 sgd rank %d] epoch %.3f ms: %d sub-epoch kernels %.3f ms, waiting for item blocks %.3f ms\n",
                m.rank, tot, RB, tk, tot - tk);
        cudaEventDestroy(e_end);
        for (auto& e : tev) cudaEventDestroy(e);
    }
    if (m.R > 1) m.items_dirty = true;
    if (m.wait_stats.p) {
        unsigned long long h[2] = {0, 0};
        cudaStreamSynchronize(s);
        cudaMemcpy(h, m.wait_stats.p, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemset(m.wait_stats.p, 0, sizeof(h));
        fprintf(stderr, "[mmlb200 sgd rank %d] epoch kernel: %.1f %% of the CTA time at block hand-overs (flag waits + CTA barriers), G=%d x %d\n",
                m.rank, h[1] ? 100.0 * (double)h[0] / (double)h[1] : 0.0, m.G, m.cpg);
    }
    return MML_OK;
}

}  // namespace mml

using namespace mml;

// =================================================================================================
// C ABI
// =================================================================================================
// A model on a one-process multi-GPU context is a root over one ordinary model per GPU (shards[r] on rank r's context,
// created from shard r of the ratings): every entry point below fans out to the shards, one host thread per GPU, which is
// exactly what the ranks of a one-process-per-GPU host would call.
struct mml_sgd { Sgd m; std::vector<mml_sgd*> shards; };

// pairs (users[t], items[t]) dealt to the rank that owns the user (u % N; ids outside the model go to rank 0)
static void deal_pairs(int N, const int32_t* users, int64_t n, std::vector<std::vector<int64_t>>& pos)
{
    pos.assign((size_t)N, std::vector<int64_t>());
    for (auto& p : pos) p.reserve((size_t)(n / N + 16));
    for (int64_t t = 0; t < n; t++) pos[(size_t)(users[t] >= 0 ? users[t] % N : 0)].push_back(t);
}

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
    p->hot_item_factor = 0.0f;
    p->hot_copies = 8;
    p->intra_block = MML_INTRA_ASYNC;
    p->async_workers = 0;
    p->hot_merge_average = 1;
}

extern "C" int32_t mml_sgd_create(mml_ctx* hctx, mml_ratings* hr, const mml_mf_params* p,
                                  const int32_t* user_perm, const int32_t* item_perm, mml_sgd** out)
{
    MML_LOCK(mml::ctx_of(hctx));
    MML_CHECK(hctx && hr && p && out, MML_ERR_ARG, "mml_sgd_create: NULL argument");
    Ctx* ctx = ctx_of(hctx); Ratings* r = ratings_of(hr);
    MML_CHECK(p->num_factors >= 1 && p->num_factors <= 256, MML_ERR_UNSUPPORTED,
              "mml_sgd_create: num_factors=%d not in [1,256]", p->num_factors);
    MML_CHECK(r->n_users() > 0 && r->n_items() > 0, MML_ERR_ARG, "mml_sgd_create: empty id space");
    MML_CHECK(p->schedule >= MML_SCHEDULE_SERIAL && p->schedule <= MML_SCHEDULE_NAIVE, MML_ERR_ARG, "mml_sgd_create: unknown schedule %d", p->schedule);
    if (ctx->is_root()) {
        MML_CHECK(r->shards.size() == ctx->peers.size(), MML_ERR_ARG, "mml_sgd_create: the ratings were not created on this multi-GPU context");
        MML_CHECK(p->schedule == MML_SCHEDULE_DSGD, MML_ERR_UNSUPPORTED,
                  "mml_sgd_create: the serial (MaxThreads = 1 order) schedule runs on one GPU; a multi-GPU context needs MML_SCHEDULE_DSGD");
        MML_CHECK(user_perm == nullptr, MML_ERR_UNSUPPORTED,
                  "mml_sgd_create: a multi-GPU context shards users by id (u %% N) when the ratings are created; user_perm must be NULL");
        mml_sgd* root = new (std::nothrow) mml_sgd();
        MML_CHECK(root != nullptr, MML_ERR_ARG, "out of host memory");
        root->m.ctx = ctx; root->m.ratings = r; root->m.p = *p; root->m.k = p->num_factors;
        root->shards.assign(ctx->peers.size(), nullptr);
        const int32_t st = on_ranks((int)ctx->peers.size(), [&](int x) -> int32_t {
            return mml_sgd_create(ctx->peers[(size_t)x], r->shards[(size_t)x], p, nullptr, item_perm, &root->shards[(size_t)x]);
        });
        if (st != MML_OK) { mml_sgd_destroy(root); return st; }
        *out = root;
        return MML_OK;
    }
    MML_CUDA(cudaSetDevice(ctx->device));
    mml_sgd* h = new (std::nothrow) mml_sgd();
    MML_CHECK(h != nullptr, MML_ERR_ARG, "out of host memory");
    Sgd& m = h->m;
    m.ctx = ctx; m.ratings = r; m.p = *p;
    m.k = p->num_factors;
    m.kp = m.k <= 32 ? 32 : (m.k <= 64 ? 64 : (m.k <= 128 ? 128 : 256));
    m.kpl = m.kp / 32;
    m.R = std::max(ctx->n_gpus, 1); m.rank = ctx->rank;
    {   // item blocks per rank. One by default: with two (MMLB200_RING_SPLIT=2) the ring step of one block overlaps the training
        // of the other, but the sub-epoch kernels lose more on the halved item blocks (shorter user runs) than the overlap wins
        // -- 25.2 ms against 17.2 ms per epoch on 2 GPUs (profiles/r2_ring_split_2gpu.log).
        const char* e = getenv("MMLB200_RING_SPLIT");
        m.split = (m.R > 1 && e && *e == '2') ? 2 : 1;
        m.RB = m.R * m.split;
    }
    cudaStream_t s = ctx->stream;
    int32_t st = MML_OK;
    do {
        // group shape: G worker groups (CTAs), W warps per CTA, C private copies per hot item
        if (p->schedule == MML_SCHEDULE_DSGD) {
            // async mode: a worker group may be several CTAs (ctas_per_group) sharing the group's blocks
            // Default (both 0): groups of 8 CTAs, SMs / 8 groups (18 x 8 on 148 SMs: with the packed loop config 4 runs 14.50 ms
            // against 14.89 ms at 37 x 4, 15.5 ms at 9 x 16 and 16.2 ms at 4 x 37; one GPU-level sub-epoch of the 8-GPU ring
            // 2.70 / 2.72 / 3.20 / 3.02 ms: profiles/r2_sgd_grid_waits.log); it also leaves four SMs to the ring's transfers.
            // CTA size (both 0, one GPU, rows of 128+ factors): 4 warps and 32 CTAs per group -- four CTAs per SM, so that an SM
            // whose CTA waits at a hand-over has three others to run: 13.73 ms against 14.2 ms for 18 x 8 CTAs of 16 warps and
            // 13.82 ms for 18 x 16 of 8 (profiles/r2_sgd_cta_size.log). Rows of up to 64 factors: 8 warps, 18 x 8 (config 2: 1.66 ms
            // against 1.75 ms with 16 warps). Ring sub-epochs (several GPUs): 16 warps, 18 x 8 (2.60 against 2.68 ms).
            const bool all_default = p->num_groups <= 0 && p->ctas_per_group <= 0 && p->num_subgroups <= 0 && ctx->sm_count >= 16;
            const bool small_ctas = all_default && p->intra_block == MML_INTRA_ASYNC && m.R == 1 && m.kp >= 128;
            m.cpg = 1;
            if (p->intra_block == MML_INTRA_ASYNC) {
                if (p->ctas_per_group > 0) m.cpg = std::min(p->ctas_per_group, ctx->sm_count);
                else if (p->num_groups <= 0 && ctx->sm_count >= 16) m.cpg = small_ctas ? 32 : 8;
            }
            int32_t G = p->num_groups > 0 ? p->num_groups : (small_ctas ? ctx->sm_count / 8 : ctx->sm_count / m.cpg);
            G = std::min(G, std::min(r->n_users(), r->n_items()));
            m.G = std::max(G, 1);
            int32_t W = p->num_subgroups > 0 ? p->num_subgroups : (small_ctas ? 4 : (all_default && m.R > 1 ? 16 : 8));
            m.W = std::max(1, std::min(W, 16));
            m.hot_copies = p->hot_copies > 0 ? std::min(p->hot_copies, 64) : 8;
        } else {
            m.G = 1; m.W = 1; m.hot_copies = 1; m.cpg = 1;
        }
        // counts -> host (group balancing, zero rows)
        // item counts, rating sum and scale are global quantities: reduce them over the ranks
        std::vector<uint32_t> cu(r->n_users()), ci(r->n_items());
        if ((st = m.item_counts.alloc(r->n_items()))) break;
        cudaMemcpyAsync(m.item_counts.p, r->count_by_item.p, sizeof(uint32_t) * ci.size(), cudaMemcpyDeviceToDevice, s);
        if ((st = dist_allreduce_u32(ctx, m.item_counts.p, ci.size()))) break;
        double g_avg = r->average; float g_min = r->min_rating, g_max = r->max_rating;
        if (m.R > 1) {
            DevBuf<double> red;
            if ((st = red.alloc(4))) break;
            // sum of ratings and their number (sum), then -min and max (max)
            double h4[4] = { (double)r->average * (double)r->n, (double)r->n, -(double)r->min_rating, (double)r->max_rating };
            if (r->n == 0) { h4[2] = -1e30; h4[3] = -1e30; }
            cudaMemcpyAsync(red.p, h4, sizeof(h4), cudaMemcpyHostToDevice, s);
            if ((st = dist_allreduce_f64(ctx, red.p, 2))) break;
            if ((st = dist_allreduce_f64_max(ctx, red.p + 2, 2))) break;
            cudaMemcpyAsync(h4, red.p, sizeof(h4), cudaMemcpyDeviceToHost, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("stats reduction failed"); st = MML_ERR_CUDA; break; }
            g_avg = h4[1] > 0 ? (double)((float)h4[0] / (float)h4[1]) : 0.0;
            g_min = (float)-h4[2]; g_max = (float)h4[3];
        }
        if (cudaMemcpyAsync(cu.data(), r->count_by_user.p, sizeof(uint32_t) * cu.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaMemcpyAsync(ci.data(), m.item_counts.p, sizeof(uint32_t) * ci.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            set_error("mml_sgd_create: count download failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = MML_ERR_CUDA; break;
        }
        const bool dsgd = p->schedule == MML_SCHEDULE_DSGD;
        const int32_t rule = dsgd ? p->group_rule : MML_GROUPS_PERM_MOD;
        build_group_map(m.users, r->n_users(), cu.data(), dsgd ? user_perm : nullptr, m.R, m.rank, m.G, 1, rule, 0, nullptr);
        // Items: an item with d ratings in a block needs d rounds there. Items whose per-block share
        // count / G reaches hot_item_factor x the block's ideal round count (ratings per block / workers) are
        // "hot" and get C private copies per block -- which needs the staged (shared-memory) item group, so
        // fall back step by step when that does not fit.
        int max_optin = 0;
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
        {   // diagnostic knob: MMLB200_SGD_VARIANT=0 selects the first form of the async loop (sgd_block_async)
            const char* ev = getenv("MMLB200_SGD_VARIANT");   // default 1: second form, per-lane vector atomics; 2: bulk reductions
            m.variant = (ev && *ev == '0') ? 0 : ((ev && *ev == '2') ? 2 : 1);
        }
        const int lanes = async_lanes(m.kp, m.variant);
        const int n_workers = m.W * (32 / lanes);
        int64_t hot_min = 0;
        const bool async = dsgd && p->intra_block == MML_INTRA_ASYNC;
        if (dsgd && !async && m.hot_copies > 1 && p->hot_item_factor > 0.f) {
            const double ideal_rounds = std::max((double)r->n / ((double)m.G * m.G) / n_workers, 4.0);
            hot_min = std::max<int64_t>((int64_t)((double)p->hot_item_factor * ideal_rounds * m.G), 2);
        }
        size_t need = 0;
        int64_t max_rows = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            build_group_map(m.items, r->n_items(), ci.data(), dsgd ? item_perm : nullptr, m.RB, -1, m.G, 1, rule, hot_min, &m.h_hot_cnt);
            m.h_item_ptr.assign(m.items.grp_ptr.begin(), m.items.grp_ptr.end());   // [R * G + 1]
            max_rows = 0;
            m.n_hot = 0;
            for (int32_t g = 0; g < m.RB * m.G; g++) {
                max_rows = std::max<int64_t>(max_rows, (int64_t)(m.h_item_ptr[g + 1] - m.h_item_ptr[g]) + (int64_t)m.h_hot_cnt[g] * m.hot_copies);
                m.n_hot += m.h_hot_cnt[g];
            }
            need = (size_t)max_rows * (m.kp + 2) * sizeof(float);   // row + bias + lock word
            if (need <= (size_t)max_optin || hot_min == 0) break;
            hot_min = 0;   // retry without hot items
        }
        m.stage_bytes = (dsgd && !async && need <= (size_t)max_optin) ? ((need + 15) / 16) * 16 : 0;
        if (async && m.variant >= 2) m.stage_bytes = (size_t)n_workers * 2 * m.kp * sizeof(float);   // two dq buffers per worker
        int32_t nu_max = 1;
        for (int32_t g = 0; g < m.G; g++) nu_max = std::max(nu_max, m.users.grp_ptr[g + 1] - m.users.grp_ptr[g]);
        {   // owned-users mode, see sgd_free_kernel. Default for rows of up to 64 factors: there the epoch is bound by the
            // hand-overs (config 2, k = 64: 1.75 ms against 2.07 ms per epoch); at k = 128 the L2 is the limit either way and
            // the barrier-free loop's extra bookkeeping costs 5 % (config 4: 15.0 against 14.2 ms; ring sub-epoch 2.86 against
            // 2.60 ms) -- profiles/r2_sgd_owned_users.log. MMLB200_SGD_OWNED = 0 | 1 overrides.
            // Not when a single user outweighs the average worker several times over: the workers of such users are still
            // running long after everyone else has finished, and an epoch that ends with a few heavy users alone on the
            // popular item rows leaves those rows fitted to them -- seen on a set with a 37k-rating user among 10M ratings
            // (34 x the average worker load; tests/cpp host_test multi): test RMSE jumping between 0.58 and 0.68 from epoch to
            // epoch. Config 2's heaviest user is ~6 x the average load and stays within 0.05 % of the oracle; the line is drawn at 8.
            const char* eo = getenv("MMLB200_SGD_OWNED");
            uint32_t cu_max = 0;
            for (uint32_t c : cu) cu_max = std::max(cu_max, c);
            const double avg_load = (double)r->n / std::max<double>((double)m.G * n_workers * m.cpg, 1.0);
            const bool want = eo && (*eo == '0' || *eo == '1') ? *eo == '1' : (m.kp <= 64 && (double)cu_max <= 8.0 * avg_load);
            m.free_smem = (size_t)n_workers * (2 * (size_t)m.G + 1) * sizeof(uint32_t);
            m.owned = async && m.variant == 1 && p->async_workers == 0 && p->persistent != 0 && want &&
                      m.free_smem <= (size_t)max_optin && get_free_kernel(m) != nullptr;
            {
                const char* tr = getenv("MMLB200_TRACE");
                if (async && tr && *tr && *tr != '0')
                    fprintf(stderr, "[mmlb200 sgd rank %d] async schedule: %s (heaviest user %u ratings, average worker load %.0f)\n", m.rank,
                            m.owned ? "owned-users loop" : "block schedule with hand-overs", cu_max, avg_load);
            }
            if (m.owned) {
                pin_users_to_workers(m.users, cu.data(), m.G, n_workers * m.cpg, m.h_worker_ptr);
                if (cudaFuncSetAttribute((const void*)get_free_kernel(m), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m.free_smem) != cudaSuccess) {
                    set_error("mml_sgd_create: cudaFuncSetAttribute failed"); st = MML_ERR_CUDA; break;
                }
            }
        }
        if ((st = upload_group_map(m.users, s)) || (st = upload_group_map(m.items, s))) break;
        if ((st = m.d_item_ptr.alloc(m.h_item_ptr.size())) || (st = m.d_hot_cnt.alloc(m.h_hot_cnt.size())) ||
            (st = m.d_user_ptr.alloc(m.users.grp_ptr.size()))) break;
        if (cudaMemcpyAsync(m.d_item_ptr.p, m.h_item_ptr.data(), sizeof(int32_t) * m.h_item_ptr.size(), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(m.d_hot_cnt.p, m.h_hot_cnt.data(), sizeof(int32_t) * m.h_hot_cnt.size(), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(m.d_user_ptr.p, m.users.grp_ptr.data(), sizeof(int32_t) * m.users.grp_ptr.size(), cudaMemcpyHostToDevice, s) != cudaSuccess) {
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
            regw_kernel<<<grid_n(m.items.n_int), 256, 0, s>>>(m.items.d_to_ext.p, m.item_counts.p, m.items.n_int, ri, m.regw_i.p);
            m.launches += 2;
        }
        // scale and global bias (BiasedMatrixFactorization.cs:186-190 / MatrixFactorization.cs:124)
        m.min_rating = g_min; m.max_rating = g_max;
        m.range = m.max_rating - m.min_rating;
        if (p->biased) {
            const double avg = (double)((float)g_avg - m.min_rating) / (double)m.range;
            m.global_bias = (float)std::log(avg / (1 - avg));
        } else {
            m.global_bias = (float)g_avg;
        }
        m.lr = p->learn_rate;
        // strata
        if (p->schedule == MML_SCHEDULE_DSGD) {
            if (m.p.intra_block == MML_INTRA_ASYNC) {
                const int32_t nw = p->async_workers > 0 ? std::min(p->async_workers, n_workers * m.cpg) : n_workers * m.cpg;
                if ((st = build_strata_async(m, nw))) break;
            }
            else if ((st = build_strata(m, nu_max, (int32_t)std::max<int64_t>(max_rows, 1)))) break;
            if (m.p.persistent < 0) m.p.persistent = 1;
            // item rows that stay in global memory are read through L1, which is only coherent across
            // launches: the single-launch epoch needs the staged (shared-memory) item groups
            if (m.stage_bytes == 0 && m.p.intra_block != MML_INTRA_ASYNC) m.p.persistent = 0;
            if (m.p.persistent) {
                // all G CTAs must be co-resident
                slot_fn_t sf; epoch_fn_t ef;
                if ((st = get_kernels(m, &sf, &ef))) break;
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)ef, m.W * 32, m.stage_bytes) != cudaSuccess) per_sm = 0;
                if ((int64_t)per_sm * ctx->sm_count < (int64_t)m.G * m.cpg) m.p.persistent = 0;
            }
            if ((st = m.flags.alloc((size_t)m.G * m.cpg))) break;
            cudaMemsetAsync(m.flags.p, 0, m.flags.bytes(), s);
            {   // diagnostic: share of the epoch kernel's CTA time spent waiting at block hand-overs (thread 0 of every CTA)
                const char* ev = getenv("MMLB200_SGD_WAITSTATS");
                if (ev && *ev && *ev != '0') {
                    if ((st = m.wait_stats.alloc(2))) break;
                    cudaMemsetAsync(m.wait_stats.p, 0, m.wait_stats.bytes(), s);
                }
            }
            m.epoch_base = 0;
        }
        for (int x = 0; x < 2 && m.R > 1; x++)
            if (cudaEventCreateWithFlags(&m.ev_kern[x], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&m.ev_xchg[x], cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); st = MML_ERR_CUDA; }
        if (st) break;
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
    MML_LOCK((h ? h->m.ctx : nullptr));
    if (!h) return MML_OK;
    if (h->m.ctx->is_root()) {
        for (mml_sgd* s : h->shards) mml_sgd_destroy(s);
        delete h;
        return MML_OK;
    }
    cudaSetDevice(h->m.ctx->device);
    cudaStreamSynchronize(h->m.ctx->stream);
    if (h->m.ev0) cudaEventDestroy(h->m.ev0);
    if (h->m.ev1) cudaEventDestroy(h->m.ev1);
    for (int x = 0; x < 2; x++) { if (h->m.ev_kern[x]) cudaEventDestroy(h->m.ev_kern[x]); if (h->m.ev_xchg[x]) cudaEventDestroy(h->m.ev_xchg[x]); }
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
    if (m.p.biased && m.p.bold_driver) {
        // BiasedMatrixFactorization.cs:168-169: InitModel computes last_loss BEFORE Train() sets rating_range_size and
        // global_bias (:186-190), i.e. with both still 0 -- every prediction is min_rating. Same here, so that the first
        // bold-driver decision compares against the number the reference compares against.
        const float gb = m.global_bias, range = m.range;
        m.global_bias = 0.f; m.range = 0.f;
        const int32_t st = objective(m, &m.last_loss);
        m.global_bias = gb; m.range = range;
        MML_TRY(st);
        m.last_loss = (double)(float)m.last_loss;
    }
    return MML_OK;
}

extern "C" int32_t mml_sgd_set_model(mml_sgd* h, const float* user_factors, const float* item_factors,
                                     const float* user_bias, const float* item_bias)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && user_factors && item_factors, MML_ERR_ARG, "mml_sgd_set_model: NULL argument");
    MML_FORWARD_ALL(h, mml_sgd_set_model(s, user_factors, item_factors, user_bias, item_bias));   // every shard keeps its own user rows
    Sgd& m = h->m;
    MML_CUDA(cudaSetDevice(m.ctx->device));
    MML_TRY(rows_from_host(m, m.users, m.ratings->count_by_user.p, user_factors, m.P.p));
    MML_TRY(rows_from_host(m, m.items, m.item_counts.p, item_factors, m.Q.p));
    MML_TRY(vec_from_host(m, m.users, user_bias, m.bu.p));
    MML_TRY(vec_from_host(m, m.items, item_bias, m.bi.p));
    MML_CUDA(cudaStreamSynchronize(m.ctx->stream));
    return after_init(m);
}

extern "C" int32_t mml_sgd_init_model(mml_sgd* h, uint64_t seed, double init_mean, double init_stddev)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "mml_sgd_init_model: NULL argument");
    MML_FORWARD_ALL(h, mml_sgd_init_model(s, seed, init_mean, init_stddev));   // counter-based: a row's values do not depend on the shard
    Sgd& m = h->m;
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    init_rows_kernel<<<grid_n((int64_t)m.users.n_int * 32), 256, 0, s>>>(m.users.d_to_ext.p, m.ratings->count_by_user.p,
        m.users.n_int, m.k, m.kp, seed, 1, (float)init_mean, (float)init_stddev, m.P.p);
    init_rows_kernel<<<grid_n((int64_t)m.items.n_int * 32), 256, 0, s>>>(m.items.d_to_ext.p, m.item_counts.p,
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
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "mml_sgd_get_model: NULL argument");
    if (!h->shards.empty()) {
        // user rows and biases come from the shard that owns the user (u % N), the item side from shard 0 (every shard
        // holds all item rows after the home blocks were broadcast, which get_model does on every shard: a collective)
        const int N = (int)h->shards.size();
        const int64_t nu = h->m.ratings->n_users(), k = h->m.k;
        return on_ranks(N, [&](int x) -> int32_t {
            std::vector<float> U, bu;
            if (user_factors) U.assign((size_t)(nu * k), 0.f);
            if (user_bias) bu.assign((size_t)nu, 0.f);
            MML_TRY(mml_sgd_get_model(h->shards[(size_t)x], user_factors ? U.data() : nullptr, x == 0 ? item_factors : nullptr,
                                      user_bias ? bu.data() : nullptr, x == 0 ? item_bias : nullptr,
                                      x == 0 ? global_bias : nullptr, x == 0 ? current_learnrate : nullptr));
            for (int64_t u = x; u < nu; u += N) {
                if (user_factors) memcpy(user_factors + u * k, U.data() + u * k, sizeof(float) * (size_t)k);
                if (user_bias) user_bias[u] = bu[(size_t)u];
            }
            return MML_OK;
        });
    }
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_get_model: no model (call set_model / init_model first)");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    MML_TRY(sync_items(m));
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
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    for (mml_sgd* s : h->shards) s->m.lr = lr;
    h->m.lr = lr;
    return MML_OK;
}

extern "C" int32_t mml_sgd_set_scale(mml_sgd* h, float min_rating, float max_rating, float global_bias)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    for (mml_sgd* s : h->shards) MML_TRY(mml_sgd_set_scale(s, min_rating, max_rating, global_bias));
    h->m.min_rating = min_rating; h->m.max_rating = max_rating; h->m.range = max_rating - min_rating;
    h->m.global_bias = global_bias;
    return MML_OK;
}

extern "C" int32_t mml_sgd_invalidate_index(mml_sgd* h)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    for (mml_sgd* s : h->shards) s->m.n_index = -1;
    h->m.n_index = -1;
    return MML_OK;
}

extern "C" int32_t mml_sgd_iterate(mml_sgd* h, const int32_t* subepoch_sequence, const int32_t* random_index, int64_t n_index)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "mml_sgd_iterate: NULL argument");
    MML_FORWARD_ALL(h, mml_sgd_iterate(s, subepoch_sequence, random_index, n_index));
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_iterate: no model (call set_model / init_model first)");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    MML_CUDA(cudaEventRecord(m.ev0, s));
    if (m.p.schedule == MML_SCHEDULE_DSGD) {
        MML_TRY(run_dsgd_epoch(m, subepoch_sequence));
    } else {
        MML_CHECK(n_index == m.ratings->n, MML_ERR_ARG, "mml_sgd_iterate: this schedule needs RandomIndex of length %lld (got %lld)",
                  (long long)m.ratings->n, (long long)n_index);
        const bool fresh_index = m.n_index != n_index;
        if (m.n_index != n_index) {
            MML_CHECK(random_index != nullptr, MML_ERR_ARG, "mml_sgd_iterate: random_index is NULL");
            for (int64_t t = 0; t < n_index; t++)
                MML_CHECK(random_index[t] >= 0 && random_index[t] < m.ratings->n, MML_ERR_ARG, "random_index[%lld] out of range", (long long)t);
            MML_TRY(m.d_index.alloc(n_index));
            MML_CUDA(cudaMemcpyAsync(m.d_index.p, random_index, sizeof(int32_t) * n_index, cudaMemcpyHostToDevice, s));
            m.n_index = n_index;
        }
        if (m.p.schedule == MML_SCHEDULE_NAIVE) MML_TRY(run_naive_epoch(m, fresh_index));
        else MML_TRY(run_serial(m, m.d_index.p, n_index, 1, 1));
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
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && (indices || n == 0), MML_ERR_ARG, "mml_sgd_iterate_indices: NULL argument");
    MML_CHECK(h->shards.empty(), MML_ERR_UNSUPPORTED, "mml_sgd_iterate_indices: rating indices are per GPU on a multi-GPU context");
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

extern "C" int32_t mml_sgd_learn_factors(mml_sgd* h, const int32_t* indices, int64_t n, int32_t update_user, int32_t update_item,
                                         int32_t num_iter)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && (indices || n == 0), MML_ERR_ARG, "mml_sgd_learn_factors: NULL argument");
    MML_CHECK(num_iter >= 0, MML_ERR_ARG, "mml_sgd_learn_factors: num_iter = %d", num_iter);
    MML_CHECK(h->m.p.schedule >= MML_SCHEDULE_SERIAL && h->m.p.schedule <= MML_SCHEDULE_NAIVE, MML_ERR_ARG, "unknown schedule");
    MML_CHECK(h->shards.empty(), MML_ERR_UNSUPPORTED, "mml_sgd_learn_factors: rating indices are per GPU on a multi-GPU context");
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_learn_factors: no model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    for (int64_t t = 0; t < n; t++)
        MML_CHECK(indices[t] >= 0 && indices[t] < m.ratings->n, MML_ERR_ARG, "indices[%lld] out of range", (long long)t);
    DevBuf<int32_t> d;
    MML_TRY(d.alloc(n));
    if (n > 0) MML_CUDA(cudaMemcpyAsync(d.p, indices, sizeof(int32_t) * n, cudaMemcpyHostToDevice, m.ctx->stream));
    // MatrixFactorization.cs:198-202: NumIter passes of Iterate(list, ...); plain MF's Iterate(list) ends with UpdateLearnRate
    // (:195), the biased override (BiasedMatrixFactorization.cs:264-310) does not touch the learn rate
    for (int32_t it = 0; it < num_iter; it++) {
        MML_TRY(run_serial(m, d.p, n, update_user, update_item));
        if (!m.p.biased) MML_TRY(update_learnrate(m));
    }
    MML_CUDA(cudaStreamSynchronize(m.ctx->stream));
    return MML_OK;
}

extern "C" int32_t mml_sgd_predict(mml_sgd* h, const int32_t* users, const int32_t* items, int64_t n, float* out)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && (n == 0 || (users && items && out)), MML_ERR_ARG, "mml_sgd_predict: NULL argument");
    if (!h->shards.empty()) {   // every pair is predicted by the GPU that owns the user's row
        std::vector<std::vector<int64_t>> pos;
        deal_pairs((int)h->shards.size(), users, n, pos);
        return on_ranks((int)h->shards.size(), [&](int x) -> int32_t {
            const std::vector<int64_t>& p = pos[(size_t)x];
            std::vector<int32_t> su(p.size()), si(p.size()); std::vector<float> so(p.size());
            for (size_t t = 0; t < p.size(); t++) { su[t] = users[p[t]]; si[t] = items[p[t]]; }
            MML_TRY(mml_sgd_predict(h->shards[(size_t)x], su.data(), si.data(), (int64_t)p.size(), so.data()));
            for (size_t t = 0; t < p.size(); t++) out[p[t]] = so[t];
            return MML_OK;
        });
    }
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_predict: no model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    MML_TRY(sync_items(m));   // collective on a multi-GPU context: every rank calls predict
    if (n == 0) return MML_OK;
    cudaStream_t s = m.ctx->stream;
    DevBuf<int32_t>& du = m.scr_u; DevBuf<int32_t>& di = m.scr_i; DevBuf<float>& dout = m.scr_v;
    if (du.n < (size_t)n) MML_TRY(du.alloc(n));
    if (di.n < (size_t)n) MML_TRY(di.alloc(n));
    if (dout.n < (size_t)n) MML_TRY(dout.alloc(n));
    MML_CUDA(cudaMemcpyAsync(du.p, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(di.p, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    {
        const PredArgs a = make_pred_args(m);
        const int g = grid_n(n * 32);
        switch (m.kp) {
            case 32: predict_kernel<8, 1><<<g, 256, 0, s>>>(a, du.p, di.p, n, dout.p); break;
            case 64: predict_kernel<16, 1><<<g, 256, 0, s>>>(a, du.p, di.p, n, dout.p); break;
            case 128: predict_kernel<16, 2><<<g, 256, 0, s>>>(a, du.p, di.p, n, dout.p); break;
            default: predict_kernel<32, 2><<<g, 256, 0, s>>>(a, du.p, di.p, n, dout.p); break;
        }
    }
    MML_CUDA(cudaGetLastError());
    m.launches++;
    MML_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

extern "C" int32_t mml_sgd_evaluate(mml_sgd* h, const int32_t* users, const int32_t* items, const float* values, int64_t n, float* out4)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && out4 && (n == 0 || (users && items && values)), MML_ERR_ARG, "mml_sgd_evaluate: NULL argument");
    if (!h->shards.empty()) {   // every GPU evaluates the pairs of its own users; the sums are all-reduced inside
        const int N = (int)h->shards.size();
        std::vector<std::vector<int64_t>> pos;
        deal_pairs(N, users, n, pos);
        std::vector<float> res((size_t)N * 4, 0.f);
        MML_TRY(on_ranks(N, [&](int x) -> int32_t {
            const std::vector<int64_t>& p = pos[(size_t)x];
            std::vector<int32_t> su(p.size() + 1), si(p.size() + 1); std::vector<float> sv(p.size() + 1);
            for (size_t t = 0; t < p.size(); t++) { su[t] = users[p[t]]; si[t] = items[p[t]]; sv[t] = values[p[t]]; }
            return mml_sgd_evaluate(h->shards[(size_t)x], su.data(), si.data(), sv.data(), (int64_t)p.size(), res.data() + 4 * x);
        }));
        for (int c = 0; c < 4; c++) out4[c] = res[(size_t)c];
        return MML_OK;
    }
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_evaluate: no model");
    MML_CHECK(n > 0 || m.R > 1, MML_ERR_ARG, "mml_sgd_evaluate: empty test set");   // Eval/Ratings.cs:98-99 returns null
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    DevBuf<int32_t>& du = m.scr_u; DevBuf<int32_t>& di = m.scr_i; DevBuf<float>& dv = m.scr_v;
    if (du.n < (size_t)n) MML_TRY(du.alloc(n));
    if (di.n < (size_t)n) MML_TRY(di.alloc(n));
    if (dv.n < (size_t)n) MML_TRY(dv.alloc(n));
    // The test ratings travel on the copy stream: in the find-iter loop (Iterate(); Evaluate(test)) the upload runs
    // under the epoch kernel still in flight on `s`; the scratch buffers are free (the previous Evaluate synchronised).
    cudaStream_t cs = m.ctx->copy_stream;
    MML_CUDA(cudaMemcpyAsync(du.p, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, cs));
    MML_CUDA(cudaMemcpyAsync(di.p, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, cs));
    MML_CUDA(cudaMemcpyAsync(dv.p, values, sizeof(float) * n, cudaMemcpyHostToDevice, cs));
    MML_CUDA(cudaEventRecord(m.ctx->copy_done, cs));
    MML_CUDA(cudaStreamWaitEvent(s, m.ctx->copy_done, 0));
    double sums[5];
    MML_TRY(evaluate_device(m, du.p, di.p, dv.p, n, sums));
    sums[4] = (double)n;
    MML_TRY(reduce_over_ranks(m, sums, 5));   // multi-GPU: every rank passes its own users' test ratings
    MML_CHECK(sums[4] > 0, MML_ERR_ARG, "mml_sgd_evaluate: empty test set");
    sums_to_measures(m, sums, (int64_t)sums[4], out4);
    return MML_OK;
}

extern "C" int32_t mml_sgd_evaluate_train(mml_sgd* h, float* out4)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && out4, MML_ERR_ARG, "mml_sgd_evaluate_train: NULL argument");
    if (!h->shards.empty()) {
        const int N = (int)h->shards.size();
        std::vector<float> res((size_t)N * 4, 0.f);
        MML_TRY(on_ranks(N, [&](int x) -> int32_t { return mml_sgd_evaluate_train(h->shards[(size_t)x], res.data() + 4 * x); }));
        for (int c = 0; c < 4; c++) out4[c] = res[(size_t)c];
        return MML_OK;
    }
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_evaluate_train: no model");
    MML_CHECK(m.ratings->n > 0 || m.R > 1, MML_ERR_ARG, "mml_sgd_evaluate_train: empty training set");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    double sums[5];
    MML_TRY(evaluate_train_device(m, sums));
    sums[4] = (double)m.ratings->n;
    MML_TRY(reduce_over_ranks(m, sums, 5));
    sums_to_measures(m, sums, (int64_t)sums[4], out4);
    return MML_OK;
}

extern "C" int32_t mml_sgd_objective(mml_sgd* h, double* out)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && out, MML_ERR_ARG, "mml_sgd_objective: NULL argument");
    if (!h->shards.empty()) {
        const int N = (int)h->shards.size();
        std::vector<double> res((size_t)N, 0.0);
        MML_TRY(on_ranks(N, [&](int x) -> int32_t { return mml_sgd_objective(h->shards[(size_t)x], &res[(size_t)x]); }));
        *out = res[0];
        return MML_OK;
    }
    MML_CHECK(h->m.has_model, MML_ERR_STATE, "mml_sgd_objective: no model");
    MML_CUDA(cudaSetDevice(h->m.ctx->device));
    return objective(h->m, out);
}

extern "C" int32_t mml_sgd_stats(mml_sgd* h, int64_t* kernel_launches, float* last_iterate_ms)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) {   // launches of all GPUs, device time of the slowest
        int64_t total = 0; float worst = 0.f;
        for (mml_sgd* s : h->shards) {
            int64_t l = 0; float ms = 0.f;
            MML_TRY(mml_sgd_stats(s, &l, &ms));
            total += l; worst = std::max(worst, ms);
        }
        if (kernel_launches) *kernel_launches = total;
        if (last_iterate_ms) *last_iterate_ms = worst;
        return MML_OK;
    }
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

extern "C" int32_t mml_sgd_strata_info(mml_sgd* h, int32_t* G, int32_t* W, int64_t* n_rounds, int64_t* staged_bytes)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) return mml_sgd_strata_info(h->shards[0], G, W, n_rounds, staged_bytes);
    if (G) *G = h->m.G;
    if (W) *W = h->m.W;
    if (n_rounds) *n_rounds = h->m.n_rounds;
    if (staged_bytes) *staged_bytes = h->m.p.intra_block == MML_INTRA_ASYNC ? 0 : (int64_t)h->m.stage_bytes;   // async: dq buffers, not item groups
    return MML_OK;
}

extern "C" int32_t mml_sgd_grid(mml_sgd* h, int32_t* G, int32_t* ctas_per_group)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) return mml_sgd_grid(h->shards[0], G, ctas_per_group);
    if (G) *G = h->m.G;
    if (ctas_per_group) *ctas_per_group = h->m.cpg;
    return MML_OK;
}

extern "C" int32_t mml_sgd_hot_items(mml_sgd* h, int64_t* n_hot)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && n_hot, MML_ERR_ARG, "NULL argument");
    if (!h->shards.empty()) return mml_sgd_hot_items(h->shards[0], n_hot);
    *n_hot = h->m.n_hot;
    return MML_OK;
}

extern "C" int32_t mml_sgd_schedule_dump(mml_sgd* h, const int32_t* subepoch_sequence, int32_t* order,
                                        int32_t* block, int32_t* copy, int32_t* round)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && order, MML_ERR_ARG, "mml_sgd_schedule_dump: NULL argument");
    MML_CHECK(h->shards.empty(), MML_ERR_UNSUPPORTED, "mml_sgd_schedule_dump: per GPU on a multi-GPU context");
    Sgd& m = h->m;
    MML_CHECK(m.p.schedule == MML_SCHEDULE_DSGD, MML_ERR_STATE, "mml_sgd_schedule_dump: not a DSGD model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    cudaStream_t s = m.ctx->stream;
    const int64_t n = m.ratings->n;
    const bool async = m.p.intra_block == MML_INTRA_ASYNC;
    std::vector<uint32_t> brp((size_t)m.n_blk + 1), rp((size_t)m.n_rounds + 1), wp;
    std::vector<int32_t> idx(std::max<int64_t>(n, 1));
    if (async) {
        wp.resize((size_t)m.n_blk * (m.n_workers + 1));
        MML_CUDA(cudaMemcpyAsync(wp.data(), m.wptr.p, sizeof(uint32_t) * wp.size(), cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaStreamSynchronize(s));
        // one "round" per block: its entries in worker-slice order (a serial-equivalent order only with 1 worker)
        rp.assign((size_t)m.n_blk + 1, 0);
        for (int32_t bk = 0; bk < m.n_blk; bk++) { brp[bk] = bk; rp[bk] = wp[(size_t)bk * (m.n_workers + 1)]; }
        brp[m.n_blk] = m.n_blk; rp[m.n_blk] = (uint32_t)n;
    }
    std::vector<int8_t> cp(std::max<int64_t>(n, 1));
    if (!async) {
        MML_CUDA(cudaMemcpyAsync(brp.data(), m.blk_round_ptr.p, sizeof(uint32_t) * brp.size(), cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaMemcpyAsync(rp.data(), m.round_ptr.p, sizeof(uint32_t) * rp.size(), cudaMemcpyDeviceToHost, s));
    }
    if (n > 0) {
        MML_CUDA(cudaMemcpyAsync(idx.data(), m.ent_idx.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
        MML_CUDA(cudaMemcpyAsync(cp.data(), m.ent_copy.p, sizeof(int8_t) * n, cudaMemcpyDeviceToHost, s));
    }
    MML_CUDA(cudaStreamSynchronize(s));
    int64_t pos = 0;
    const int G = m.G;
    for (int S = 0; S < m.RB; S++)          // execution order of the GPU-level sub-epochs on this rank
        for (int t = 0; t < G; t++) {
            const int B = (S + m.split * m.rank) % m.RB;
            const int slot = subepoch_sequence ? subepoch_sequence[t] : t;
            MML_CHECK(slot >= 0 && slot < G, MML_ERR_ARG, "subepoch_sequence[%d] out of range", t);
            for (int j = 0; j < G; j++) {
                const size_t blk = ((size_t)B * G + j) * G + slot;
                for (uint32_t rd = brp[blk]; rd < brp[blk + 1]; rd++)
                    for (uint32_t e = rp[rd]; e < rp[rd + 1]; e++) {
                        order[pos] = idx[e];
                        if (block) block[pos] = (S * G + t) * G + j;
                        if (copy) copy[pos] = cp[e];
                        if (round) round[pos] = (int32_t)rd;
                        pos++;
                    }
            }
        }
    MML_CHECK(pos == n, MML_ERR_STATE, "schedule covers %lld of %lld ratings", (long long)pos, (long long)n);
    return MML_OK;
}

// =================================================================================================
// fold-in / incremental operations
// =================================================================================================
namespace mml {

struct FoldArgs {
    int32_t k, num_iter, freq_reg, loss;
    float learn_rate, decay, reg, blr, breg;
};

// FoldIn for a batch of users that are not part of the model, one warp per user (the users are independent; inside a
// user the reference's loop is a serial chain over its ratings). Arithmetic follows MatrixFactorization.cs:323-347 /
// BiasedMatrixFactorization.cs:445-492 operation by operation: sequential fp32 dot (mul then add), double link, the
// per-factor step as an fp32 expression widened to double, `+= (float)`. Lane l holds factors l, l + 32, ...
template <bool BIASED>
__global__ void __launch_bounds__(128) fold_in_kernel(const PredArgs a, const FoldArgs fa,
                                                      const int64_t* __restrict__ rated_ptr,
                                                      const int32_t* __restrict__ rated_items,
                                                      const float* __restrict__ rated_values, int64_t n_users,
                                                      const float* __restrict__ init, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_users) return;
    const int k = fa.k, nslot = a.kp / 32;
    float pv[8], qv[8], prod[8];
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const int f = s * 32 + lane;
        pv[s] = (s < nslot && f < k) ? init[w * k + f] : 0.f;
    }
    const int64_t lo = rated_ptr[w], hi = rated_ptr[w + 1];
    float regw = fa.reg;
    if (BIASED && fa.freq_reg) regw = (float)((double)fa.reg / sqrt((double)(hi - lo)));   // :453
    float ub = 0.f;
    double lr = (double)fa.learn_rate;
    for (int it = 0; it < fa.num_iter; it++) {
        for (int64_t t = lo; t < hi; t++) {
            const int32_t row = a.item_int[rated_items[t]];
            const float r = rated_values[t];
            const float* qrow = a.Q + (size_t)row * a.kp;
#pragma unroll
            for (int s = 0; s < 8; s++)
                if (s < nslot) { qv[s] = qrow[s * 32 + lane]; prod[s] = __fmul_rn(qv[s], pv[s]); }
            float dot = 0.f;
#pragma unroll
            for (int s = 0; s < 8; s++)
                if (s < nslot) {
                    const int cnt = min(32, k - s * 32);
                    for (int j = 0; j < cnt; j++) dot = __fadd_rn(dot, __shfl_sync(0xffffffffu, prod[s], j));
                }
            if (BIASED) {
                const double score = (double)__fadd_rn(__fadd_rn(__fadd_rn(a.gb, ub), a.bi[row]), dot);
                const double sig = 1.0 / (1.0 + exp(-score));
                const double pred = __dadd_rn((double)a.minr, __dmul_rn(sig, (double)a.range));
                const double err = (double)r - pred;
                float gc;
                if (fa.loss == MML_LOSS_RMSE) gc = (float)(err * sig * (1.0 - sig) * (double)a.range);
                else if (fa.loss == MML_LOSS_MAE) gc = (float)((err > 0 ? 1.0 : (err < 0 ? -1.0 : 0.0)) * sig * (1.0 - sig) * (double)a.range);
                else gc = (float)err;
                ub = __fadd_rn(ub, __fmul_rn(__fmul_rn(fa.blr, fa.learn_rate), __fsub_rn(gc, __fmul_rn(__fmul_rn(fa.breg, regw), ub))));
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if (s < nslot) {
                        const float d = __fsub_rn(__fmul_rn(gc, qv[s]), __fmul_rn(regw, pv[s]));
                        pv[s] = __fadd_rn(pv[s], (float)((double)fa.learn_rate * (double)d));
                    }
            } else {
                const float err = __fsub_rn(r, __fadd_rn(a.gb, dot));                    // Predict(vector, item, false)
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if (s < nslot) {
                        const float d = __fsub_rn(__fmul_rn(err, qv[s]), __fmul_rn(regw, pv[s]));
                        pv[s] = __fadd_rn(pv[s], (float)(lr * (double)d));
                    }
            }
        }
        lr *= (double)fa.decay;                                                          // MatrixFactorization.cs:345
    }
    const int stride = BIASED ? k + 1 : k;
    float* o = out + w * stride + (BIASED ? 1 : 0);                                      // FOLD_IN_FACTORS_START
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const int f = s * 32 + lane;
        if (s < nslot && f < k) o[f] = pv[s];
    }
    if (BIASED && lane == 0) out[w * stride] = ub;                                       // FOLD_IN_BIAS_INDEX
}

// Predict(float[] user_vector, int item_id) for every (vector, candidate) pair, one thread per pair, sequential fp32 dot.
// MatrixFactorization.cs:223-241 (clipped to the rating scale), BiasedMatrixFactorization.cs:328-336 (item terms only for
// items the model knows).
__global__ void score_vectors_kernel(const PredArgs a, int32_t k, const float* __restrict__ vectors, int64_t n_users,
                                     const int32_t* __restrict__ cand, int64_t n_cand, float* __restrict__ out)
{
    const int64_t total = n_users * n_cand;
    const int stride = a.biased ? k + 1 : k;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = t / n_cand;
        const int32_t item = cand[t - u * n_cand];
        const float* v = vectors + u * stride;
        const bool known = item >= 0 && item < a.n_items_ext && a.item_int[item] >= 0;
        const int32_t row = known ? a.item_int[item] : 0;
        const float* q = a.Q + (size_t)row * a.kp;
        const float* vf = v + (a.biased ? 1 : 0);
        float dot = 0.f;
        for (int f = 0; f < k; f++) dot = __fadd_rn(dot, __fmul_rn(q[f], vf[f]));
        float res;
        if (a.biased) {
            double score = (double)__fadd_rn(a.gb, v[0]);
            if (known) score += (double)__fadd_rn(a.bi[row], dot);
            res = (float)__dadd_rn((double)a.minr, __dmul_rn(1.0 / (1.0 + exp(-score)), (double)a.range));
        } else {
            res = __fadd_rn(a.gb, dot);
            if (res > a.maxr) res = a.maxr;
            if (res < a.minr) res = a.minr;
        }
        out[t] = res;
    }
}

// rows[ids[j]] = given row (padded with zeros to kp), bias likewise; ids this rank does not hold are skipped
__global__ void set_rows_kernel(const int32_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ to_int,
                                const float* __restrict__ rows, const float* __restrict__ biases,
                                int32_t k, int32_t kp, float* __restrict__ dst, float* __restrict__ dst_bias)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const int32_t r = to_int[ids[w]];
    if (r < 0) return;
    if (rows)
        for (int f = lane; f < kp; f += 32) dst[(size_t)r * kp + f] = f < k ? rows[w * k + f] : 0.f;
    if (biases && dst_bias && lane == 0) dst_bias[r] = biases[w];
}

}  // namespace mml

extern "C" int32_t mml_sgd_fold_in(mml_sgd* h, const int64_t* rated_ptr, const int32_t* rated_items,
                                   const float* rated_values, int64_t n_users, const float* init_factors,
                                   int32_t num_iter, float* out_vectors)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && (n_users == 0 || (rated_ptr && init_factors && out_vectors)), MML_ERR_ARG, "mml_sgd_fold_in: NULL argument");
    if (!h->shards.empty()) {   // needs the item side only: every GPU synchronises its item rows (a collective), GPU 0 answers
        return on_ranks((int)h->shards.size(), [&](int x) -> int32_t {
            if (x == 0) return mml_sgd_fold_in(h->shards[0], rated_ptr, rated_items, rated_values, n_users, init_factors, num_iter, out_vectors);
            return mml_sgd_fold_in(h->shards[(size_t)x], nullptr, nullptr, nullptr, 0, nullptr, num_iter, nullptr);
        });
    }
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_fold_in: no model");
    MML_CHECK(num_iter >= 0 && n_users >= 0, MML_ERR_ARG, "mml_sgd_fold_in: negative count");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    MML_TRY(sync_items(m));
    if (n_users == 0) return MML_OK;
    MML_CHECK(rated_ptr[0] == 0, MML_ERR_ARG, "mml_sgd_fold_in: rated_ptr[0] must be 0");
    for (int64_t u = 0; u < n_users; u++)
        MML_CHECK(rated_ptr[u + 1] >= rated_ptr[u], MML_ERR_ARG, "mml_sgd_fold_in: rated_ptr is not ascending at %lld", (long long)u);
    const int64_t nnz = rated_ptr[n_users];
    MML_CHECK(nnz == 0 || (rated_items && rated_values), MML_ERR_ARG, "mml_sgd_fold_in: NULL rated arrays");
    for (int64_t t = 0; t < nnz; t++)   // the reference indexes item_factors with the id: out of range throws there
        MML_CHECK(rated_items[t] >= 0 && rated_items[t] < m.items.n_ext && m.items.to_int[rated_items[t]] >= 0, MML_ERR_ARG,
                  "mml_sgd_fold_in: rated item %d is not part of the model", rated_items[t]);
    cudaStream_t s = m.ctx->stream;
    const int stride = m.p.biased ? m.k + 1 : m.k;
    DevBuf<int64_t> d_ptr; DevBuf<int32_t> d_it; DevBuf<float> d_val, d_init, d_out;
    MML_TRY(d_ptr.alloc(n_users + 1)); MML_TRY(d_it.alloc(nnz)); MML_TRY(d_val.alloc(nnz));
    MML_TRY(d_init.alloc((size_t)n_users * m.k)); MML_TRY(d_out.alloc((size_t)n_users * stride));
    MML_CUDA(cudaMemcpyAsync(d_ptr.p, rated_ptr, sizeof(int64_t) * (n_users + 1), cudaMemcpyHostToDevice, s));
    if (nnz > 0) {
        MML_CUDA(cudaMemcpyAsync(d_it.p, rated_items, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, s));
        MML_CUDA(cudaMemcpyAsync(d_val.p, rated_values, sizeof(float) * nnz, cudaMemcpyHostToDevice, s));
    }
    MML_CUDA(cudaMemcpyAsync(d_init.p, init_factors, sizeof(float) * (size_t)n_users * m.k, cudaMemcpyHostToDevice, s));
    FoldArgs fa{};
    fa.k = m.k; fa.num_iter = num_iter; fa.freq_reg = m.p.biased ? m.p.frequency_regularization : 0; fa.loss = m.p.loss;
    fa.learn_rate = m.p.learn_rate; fa.decay = m.p.decay;
    fa.reg = m.p.biased ? m.p.reg_u : m.p.regularization;
    fa.blr = m.p.bias_learn_rate; fa.breg = m.p.bias_reg;
    const int blocks = (int)ceil_div(n_users * 32, 128);
    if (m.p.biased) fold_in_kernel<true><<<blocks, 128, 0, s>>>(make_pred_args(m), fa, d_ptr.p, d_it.p, d_val.p, n_users, d_init.p, d_out.p);
    else fold_in_kernel<false><<<blocks, 128, 0, s>>>(make_pred_args(m), fa, d_ptr.p, d_it.p, d_val.p, n_users, d_init.p, d_out.p);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    MML_CUDA(cudaMemcpyAsync(out_vectors, d_out.p, sizeof(float) * (size_t)n_users * stride, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

extern "C" int32_t mml_sgd_score_items(mml_sgd* h, const float* user_vectors, int64_t n_users,
                                       const int32_t* candidates, int64_t n_cand, float* out_scores)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h, MML_ERR_ARG, "mml_sgd_score_items: NULL argument");
    if (!h->shards.empty()) {
        return on_ranks((int)h->shards.size(), [&](int x) -> int32_t {
            if (x == 0) return mml_sgd_score_items(h->shards[0], user_vectors, n_users, candidates, n_cand, out_scores);
            return mml_sgd_score_items(h->shards[(size_t)x], nullptr, 0, nullptr, 0, nullptr);
        });
    }
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_score_items: no model");
    MML_CHECK(n_users >= 0 && n_cand >= 0, MML_ERR_ARG, "mml_sgd_score_items: negative count");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    MML_TRY(sync_items(m));
    if (n_users == 0 || n_cand == 0) return MML_OK;
    MML_CHECK(user_vectors && candidates && out_scores, MML_ERR_ARG, "mml_sgd_score_items: NULL argument");
    if (!m.p.biased)   // MatrixFactorization.Predict(vector, item) has no range check: RowScalarProduct throws
        for (int64_t c = 0; c < n_cand; c++)
            MML_CHECK(candidates[c] >= 0 && candidates[c] < m.items.n_ext && m.items.to_int[candidates[c]] >= 0, MML_ERR_ARG,
                      "i too big: %d, dim1 is %d", candidates[c], m.items.n_ext);
    cudaStream_t s = m.ctx->stream;
    const int stride = m.p.biased ? m.k + 1 : m.k;
    DevBuf<float> d_vec, d_out; DevBuf<int32_t> d_cand;
    MML_TRY(d_vec.alloc((size_t)n_users * stride)); MML_TRY(d_cand.alloc(n_cand)); MML_TRY(d_out.alloc((size_t)n_users * n_cand));
    MML_CUDA(cudaMemcpyAsync(d_vec.p, user_vectors, sizeof(float) * (size_t)n_users * stride, cudaMemcpyHostToDevice, s));
    MML_CUDA(cudaMemcpyAsync(d_cand.p, candidates, sizeof(int32_t) * n_cand, cudaMemcpyHostToDevice, s));
    const int64_t total = n_users * n_cand;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)m.ctx->sm_count * 32);
    score_vectors_kernel<<<blocks, 256, 0, s>>>(make_pred_args(m), m.k, d_vec.p, n_users, d_cand.p, n_cand, d_out.p);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    MML_CUDA(cudaMemcpyAsync(out_scores, d_out.p, sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost, s));
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}

extern "C" int32_t mml_sgd_set_rows(mml_sgd* h, int32_t by_item, const int32_t* ids, int64_t n,
                                    const float* factors, const float* biases)
{
    MML_LOCK((h ? h->m.ctx : nullptr));
    MML_CHECK(h && (n == 0 || ids), MML_ERR_ARG, "mml_sgd_set_rows: NULL argument");
    MML_FORWARD_ALL(h, mml_sgd_set_rows(s, by_item, ids, n, factors, biases));   // a shard skips the user rows it does not hold
    Sgd& m = h->m;
    MML_CHECK(m.has_model, MML_ERR_STATE, "mml_sgd_set_rows: no model");
    MML_CUDA(cudaSetDevice(m.ctx->device));
    MML_TRY(sync_items(m));
    if (n <= 0 || (!factors && !biases)) return MML_OK;
    GroupMap& gm = by_item ? m.items : m.users;
    for (int64_t j = 0; j < n; j++)
        MML_CHECK(ids[j] >= 0 && ids[j] < gm.n_ext, MML_ERR_ARG, "mml_sgd_set_rows: id %d out of range", ids[j]);
    cudaStream_t s = m.ctx->stream;
    DevBuf<int32_t> d_ids; DevBuf<float> d_rows, d_b;
    MML_TRY(d_ids.alloc(n));
    MML_CUDA(cudaMemcpyAsync(d_ids.p, ids, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    if (factors) {
        MML_TRY(d_rows.alloc((size_t)n * m.k));
        MML_CUDA(cudaMemcpyAsync(d_rows.p, factors, sizeof(float) * (size_t)n * m.k, cudaMemcpyHostToDevice, s));
    }
    if (biases) {
        MML_TRY(d_b.alloc(n));
        MML_CUDA(cudaMemcpyAsync(d_b.p, biases, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    }
    set_rows_kernel<<<(int)ceil_div(n * 32, 256), 256, 0, s>>>(d_ids.p, n, gm.d_to_int.p, factors ? d_rows.p : nullptr,
                                                              biases ? d_b.p : nullptr, m.k, m.kp,
                                                              by_item ? m.Q.p : m.P.p, by_item ? m.bi.p : m.bu.p);
    MML_CUDA(cudaGetLastError());
    m.launches++;
    MML_CUDA(cudaStreamSynchronize(s));
    return MML_OK;
}
