// sgd.cuh -- state of one MatrixFactorization / BiasedMatrixFactorization model on one GPU.
#pragma once
#include "common.cuh"

namespace mml {

// How the ids of one side (users or items) are dealt to the stratification groups and where
// their rows live in the internal (group-major, padded) factor matrix.
struct GroupMap {
    int32_t n_ext = 0;                 // MaxID + 1
    int32_t n_int = 0;                 // rows held on this rank
    int32_t n_blocks = 1;              // GPU-level blocks covered (users: 1 = own block, items: R)
    std::vector<int32_t> grp;          // [n_ext] packed group ((blk * G + g) * W + w) or -1 (not on this rank)
    std::vector<int32_t> to_int;       // [n_ext] internal row or -1
    std::vector<int32_t> to_ext;       // [n_int]
    std::vector<int32_t> grp_ptr;      // [n_blocks * G * W + 1] internal row range of each packed group
    DevBuf<int32_t> d_grp, d_to_int, d_to_ext;
};

struct Sgd {
    Ctx* ctx = nullptr;
    Ratings* ratings = nullptr;
    mml_mf_params p{};
    int32_t k = 0, kp = 0, kpl = 0;    // factors, padded row length (32 * kpl), floats per lane
    int32_t R = 1, rank = 0;           // world size (= GPU-level user blocks) and own rank
    int32_t split = 1, RB = 1;         // item blocks per rank (2 when R > 1: ring overlap) and GPU-level item blocks R * split
    cudaEvent_t ev_kern[2] = {nullptr, nullptr}, ev_xchg[2] = {nullptr, nullptr};   // sub-epoch kernel done / its ring exchange done
    int32_t G = 1, W = 1;              // worker groups and warps per CTA
    int32_t cpg = 1;                   // CTAs per worker group (async mode; 1 otherwise)
    int32_t variant = 1;               // async epoch kernel: 0 = sgd_block_async, 1 = sgd_block_async2 (get_kernels)
    int32_t hot_copies = 1;            // private copies of a hot item row inside a block
    GroupMap users, items;
    std::vector<int32_t> h_item_ptr;   // [R * G + 1] internal item row range of CTA-level item group (B, b)
    DevBuf<int32_t> d_item_ptr;
    std::vector<int32_t> h_hot_cnt;    // [R * G] hot items of CTA-level item group (B, b) (rows first in the group)
    DevBuf<int32_t> d_hot_cnt;
    int64_t n_hot = 0;

    // model, internal order, rows padded with zeros to kp floats
    DevBuf<float> P, Q, bu, bi;
    DevBuf<float> regw_u, regw_i;      // per-row regularisation weights (frequency regularisation only)
    float global_bias = 0.f, lr = 0.f, min_rating = 0.f, max_rating = 0.f, range = 0.f;
    bool has_model = false;
    bool items_dirty = false;          // multi-GPU: item blocks away from home were updated since the last sync
    DevBuf<uint32_t> item_counts;      // CountByItem over all ranks
    double last_loss = 0.0;            // bold driver

    // strata: entries ordered by (block = (B, j, slot), round), see sgd.cu
    int32_t n_blk = 0;                 // R * G * G blocks
    int64_t n_rounds = 0;
    int32_t n_workers = 0;             // async mode: workers per worker group the slices were cut for
    DevBuf<uint32_t> wptr;             // async mode: [n_blk][n_workers + 1]
    bool owned = false;                // async mode: users pinned to workers for the whole epoch (no block hand-overs)
    std::vector<int32_t> h_worker_ptr; // owned: [G * n_workers + 1] first internal user row of every worker
    size_t free_smem = 0;              // owned: dynamic shared memory of sgd_free_kernel (the workers' slice tables)
    DevBuf<uint32_t> round_ptr;        // [n_rounds + 1] first entry of each round
    DevBuf<uint32_t> blk_round_ptr;    // [n_blk + 1] first round of each block
    DevBuf<int32_t> d_user_ptr;        // [G + 1] internal user row range of each user group
    DevBuf<int32_t> ent_u, ent_i, ent_idx;
    DevBuf<int8_t> ent_copy;           // -1 = cold item, else the private copy of the hot item row the entry updates
    DevBuf<float> ent_v;
    size_t stage_bytes = 0;            // shared memory for the largest item group; 0 = not staged
    DevBuf<uint32_t> flags;            // persistent kernel: per-CTA progress counters
    DevBuf<unsigned long long> wait_stats;   // MMLB200_SGD_WAITSTATS diagnostic: {hand-over wait cycles, CTA cycles}
    uint32_t epoch_base = 0;

    // grow-only scratch of Evaluate()/Predict() (no cudaMalloc/cudaFree in the per-epoch find-iter loop)
    DevBuf<int32_t> scr_u, scr_i;
    DevBuf<float> scr_v;
    DevBuf<double> scr_part;

    // serial schedule: cached RandomIndex
    DevBuf<int32_t> d_index;
    int64_t n_index = -1;

    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

}  // namespace mml
