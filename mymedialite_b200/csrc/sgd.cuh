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
    int32_t R = 1, rank = 0;           // GPU-level blocks (world size) and own block
    int32_t G = 1, W = 1;              // CTA-level and warp-level groups
    GroupMap users, items;
    std::vector<int32_t> h_item_ptr;   // [R * G + 1] internal item row range of CTA-level item group (B, b)
    DevBuf<int32_t> d_item_ptr;

    // model, internal order, rows padded with zeros to kp floats
    DevBuf<float> P, Q, bu, bi;
    DevBuf<float> regw_u, regw_i;      // per-row regularisation weights (frequency regularisation only)
    float global_bias = 0.f, lr = 0.f, min_rating = 0.f, max_rating = 0.f, range = 0.f;
    bool has_model = false;
    double last_loss = 0.0;            // bold driver

    // strata: entries ordered by (B, j, slot, w, step), see sgd.cu
    int64_t n_sub = 0;                 // R * G * G * W * W
    DevBuf<int32_t> ent_u, ent_i, ent_idx;
    DevBuf<float> ent_v;
    DevBuf<uint32_t> sub_ptr;          // [n_sub + 1]
    size_t stage_bytes = 0;            // shared memory for the largest item group; 0 = not staged
    DevBuf<uint32_t> flags;            // persistent kernel: per-CTA progress counters
    uint32_t epoch_base = 0;

    // serial schedule: cached RandomIndex
    DevBuf<int32_t> d_index;
    int64_t n_index = -1;

    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

}  // namespace mml
