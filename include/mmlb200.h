/*
 * mmlb200.h -- C ABI of libmmlb200.so, the B200-native matrix-factorization engine that the
 * CUDA-backed MyMediaLite recommender classes (CudaBiasedMatrixFactorization,
 * CudaMatrixFactorization, CudaWRMF) bind through P/Invoke.
 *
 * Conventions
 *   - plain C: pointers and sizes only; every pointer argument is HOST memory owned by the caller
 *     and is only read/written during the call (the library copies to/from the device);
 *   - every function returns an int32 status (MML_OK = 0) and never throws;
 *     mml_last_error() returns a thread-local message for the last failure;
 *   - handles are library-owned, released by the matching *_destroy;
 *   - there is NO CPU fallback: without a CUDA device mml_ctx_create fails with MML_ERR_CUDA.
 *
 * Each entry point cites the reference code (relative to src/MyMediaLite/ of
 * jordansilva/MyMediaLite) whose work it takes over.
 */
#ifndef MMLB200_H
#define MMLB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MML_OK            0
#define MML_ERR_CUDA      1   /* CUDA runtime / launch failure (sticky for the context) */
#define MML_ERR_ARG       2   /* invalid argument */
#define MML_ERR_STATE     3   /* call order violated (e.g. iterate before a model exists) */
#define MML_ERR_NCCL      4
#define MML_ERR_UNSUPPORTED 5
#define MML_ERR_FORMAT    6   /* malformed input line: the reference's FormatException */
#define MML_ERR_IO        7   /* file cannot be opened / mapped */

typedef struct mml_ctx     mml_ctx;      /* device, stream (and NCCL communicator when n_gpus > 1) */
typedef struct mml_ratings mml_ratings;  /* COO rating set resident in HBM */
typedef struct mml_sgd     mml_sgd;      /* MatrixFactorization / BiasedMatrixFactorization model */
typedef struct mml_feedback mml_feedback;/* implicit feedback (CSR by user and by item) in HBM */
typedef struct mml_wrmf    mml_wrmf;     /* WRMF model */
typedef struct mml_ingest  mml_ingest;   /* a parsed rating / feedback file: COO in pinned host memory + id tables */

const char* mml_last_error(void);
/* Library build info: "mmlb200 <version> sm_100a". */
const char* mml_version(void);

/* ---- context ------------------------------------------------------------------------------ */
/* device_ids may be NULL (device 0..n_gpus-1). One context per process per GPU. */
int32_t mml_ctx_create(int32_t n_gpus, const int32_t* device_ids, mml_ctx** out);
/* Multi-GPU: one process per GPU. Rank 0 calls mml_dist_unique_id (128 bytes = ncclUniqueId), hands the bytes to
 * the other ranks (the host's own plumbing: torch.distributed, MPI, a file ...), then every rank calls
 * mml_ctx_create_dist. Models created on such a context shard users over ranks (user block = user_perm[u] % world,
 * MultiCore.cs:64 lifted to GPUs) and rotate item blocks round the ring after every GPU-level sub-epoch. */
int32_t mml_dist_unique_id(uint8_t* out128);
int32_t mml_ctx_create_dist(int32_t rank, int32_t world, int32_t device, const uint8_t* unique_id128, mml_ctx** out);
int32_t mml_ctx_destroy(mml_ctx* ctx);
int32_t mml_ctx_synchronize(mml_ctx* ctx);
/* Benchmark hygiene: overwrites a 384 MB scratch buffer so the 126 MB L2 holds none of the model. */
int32_t mml_ctx_flush_l2(mml_ctx* ctx);
/* SM count of the device (used by hosts to pick the number of worker groups). */
int32_t mml_ctx_sm_count(mml_ctx* ctx, int32_t* out);
/* Measurement aid for the roofline of the SGD epoch kernel (bench.py): the rate at which this GPU's L2 serves the kernel's
 * own access pattern on a table of n_rows rows of row_floats floats that fits the L2 -- 8-lane workers touching random rows,
 * rows read L1-bypassed in 128-bit pieces (mode 0), updated with red.global.add.v4.f32 (mode 1), or read and then updated
 * (mode 2, what the kernel does to an item row per rating). *rows_per_s = rows touched per second, best of `reps` launches. */
enum { MML_L2_READ = 0, MML_L2_RED = 1, MML_L2_READ_RED = 2 };
int32_t mml_ctx_probe_l2(mml_ctx* ctx, int32_t mode, int32_t n_rows, int32_t row_floats, int32_t reps, double* rows_per_s);

/* ---- ingest: text file -> id mapping -> COO in pinned host memory (host threads; the step before the path) ------- */
enum {
    MML_FILE_RATINGS = 0,          /* user item rating [...]   IO/StaticRatingData.cs:74-117, IO/RatingData.cs:57-88 */
    MML_FILE_RATINGS_NO_VALUE = 1, /* user item [...]          TestRatingFileFormat.WITHOUT_RATINGS (value 0) */
    MML_FILE_FEEDBACK = 2          /* user item [...]          IO/ItemData.cs:59-93 */
};
enum {
    MML_MAP_IDENTITY = 0,          /* internal id = int.Parse(token)               Data/IdentityMapping.cs:62-67 */
    MML_MAP_FIRST_SEEN = 1         /* internal id = order of first appearance      Data/Mapping.cs:75-85 */
};
/* Parses `path` as the reference's readers do: lines end at "\n", "\r" or "\r\n"; empty lines are skipped (feedback
 * files: lines that are empty after Trim()); every other line is split at tab, space and comma (IO/Constants.cs:25, empty
 * tokens kept) and must have at least 3 (2) columns, else MML_ERR_FORMAT with the reference's message; column 2 is read as
 * float.Parse(InvariantCulture) does. `prior` (may be NULL) is an earlier ingest whose id tables this one continues -- the
 * reference passes the same IMapping objects to the training and the test file reader. n_threads = 0: all host threads.
 * The result is independent of n_threads (first-seen ids are assigned in file order). */
int32_t mml_ingest_file(const char* path, int32_t kind, int32_t user_mapping, int32_t item_mapping,
                        int32_t ignore_first_line, int32_t n_threads, const mml_ingest* prior, mml_ingest** out);
/* Same on a buffer already in memory (TextReader overloads, StaticRatingData.cs:82-117). */
int32_t mml_ingest_text(const char* text, int64_t len, int32_t kind, int32_t user_mapping, int32_t item_mapping,
                        int32_t ignore_first_line, int32_t n_threads, const mml_ingest* prior, mml_ingest** out);
int32_t mml_ingest_destroy(mml_ingest* h);
/* Count, MaxUserID, MaxItemID (of this data set), sizes of the id tables (0 for identity mappings), and whether the
 * COO arrays live in cudaHostAlloc'ed memory. Any output pointer may be NULL. */
int32_t mml_ingest_info(const mml_ingest* h, int64_t* n, int32_t* max_user, int32_t* max_item,
                        int32_t* n_user_ids, int32_t* n_item_ids, int32_t* pinned);
/* Copies the triples out (any pointer may be NULL; values only for MML_FILE_RATINGS). */
int32_t mml_ingest_copy(const mml_ingest* h, int32_t* users, int32_t* items, float* values);
/* Mapping.ToOriginalID (Data/Mapping.cs:63-69) for internal ids first .. first+count-1 of the user (which = 0) or item
 * (which = 1) table: the strings are written back to back into buf, offsets[count + 1] delimit them; *needed = bytes
 * required (call with buf = offsets = NULL to size the buffer). */
int32_t mml_ingest_original_ids(const mml_ingest* h, int32_t which, int32_t first, int32_t count,
                                char* buf, int64_t buf_len, int64_t* offsets, int64_t* needed);
/* mml_ratings_create / mml_feedback_create straight from the pinned COO arrays. */
int32_t mml_ingest_to_ratings(mml_ctx* ctx, const mml_ingest* h, mml_ratings** out);
int32_t mml_ingest_to_feedback(mml_ctx* ctx, const mml_ingest* h, mml_feedback** out);

/* ---- rating matrix build -------------------------------------------------------------------- */
/* Replaces StaticRatings/Ratings storage (Data/StaticRatings.cs:47-84, Data/Ratings.cs:150-190):
 * uploads the COO triples, computes CountByUser/CountByItem (Data/DataSet.cs:134-169), Average
 * (Data/Ratings.cs:76-84) and the rating scale min/max (Data/RatingScale.cs:104-117) on device. */
int32_t mml_ratings_create(mml_ctx* ctx, const int32_t* users, const int32_t* items, const float* values,
                           int64_t n, int32_t max_user, int32_t max_item, mml_ratings** out);
int32_t mml_ratings_destroy(mml_ratings* r);
/* CountByUser (by_item = 0) / CountByItem (by_item = 1): counts_out has max_id + 1 entries. */
int32_t mml_ratings_counts(mml_ratings* r, int32_t by_item, int32_t* counts_out);
/* ByUser / ByItem (Data/DataSet.cs:171-191) as CSR: row_ptr[max_id + 2], idx[n] holds rating
 * indices in ascending order inside each row (bit-exact with the reference's forward pass). */
int32_t mml_ratings_csr(mml_ratings* r, int32_t by_item, int64_t* row_ptr, int32_t* idx);
/* Average, Scale.Min, Scale.Max. */
int32_t mml_ratings_stats(mml_ratings* r, float* average, float* min_rating, float* max_rating);
/* Applies the Fisher-Yates swap targets H (H[i] = random.Next(i+1), drawn for i = n-1..0 by the
 * host's RNG) to perm in place: for i = n-1..0 swap(perm[i], perm[H[i]]) -- the permutation
 * Utils.Shuffle (Utils.cs:52-64) produces for DataSet.RandomIndex (Data/DataSet.cs:193-202). */
int32_t mml_shuffle_apply(mml_ctx* ctx, int32_t* perm, const int32_t* H, int64_t n);
/* MultiCore.PartitionUsersAndItems (MultiCore.cs:43-73) without the per-block shuffle: block
 * (user_perm[u] % g, item_perm[i] % g), rating indices ascending inside a block.
 * block_ptr[g*g + 1] (row-major), idx[n]. */
int32_t mml_partition_blocks(mml_ratings* r, const int32_t* user_perm, const int32_t* item_perm, int32_t g,
                             int64_t* block_ptr, int32_t* idx);

/* MultiCore.PartitionIndices (MultiCore.cs:79-92): random_index (= DataSet.RandomIndex, n entries) dealt round-robin into
 * g = min(num_groups, n) lists: list_ptr[num_groups + 1] (lists beyond g are empty), idx[n] = list 0, list 1, ...;
 * list l holds random_index[l], random_index[l + g], random_index[l + 2 g], ... in that order. */
int32_t mml_partition_indices(mml_ctx* ctx, const int32_t* random_index, int64_t n, int32_t num_groups,
                              int64_t* list_ptr, int32_t* idx);

/* ---- MatrixFactorization / BiasedMatrixFactorization ------------------------------------------ */
enum { MML_LOSS_RMSE = 0, MML_LOSS_MAE = 1, MML_LOSS_LOGISTIC = 2 };
enum {
    MML_SCHEDULE_SERIAL = 0,   /* MaxThreads = 1: one pass over RandomIndex in the reference's order */
    MML_SCHEDULE_DSGD   = 1,   /* stratified DSGD block schedule (BiasedMatrixFactorization.cs:205-215) */
    MML_SCHEDULE_NAIVE  = 2    /* NaiveParallelization (BiasedMatrixFactorization.cs:136-141, :201-204): RandomIndex dealt round-robin
                                  into lists (MultiCore.PartitionIndices, MultiCore.cs:79-92), one list per worker, every list walked
                                  in order, all of them at once with no exclusivity at all -- the reference's lock-free mode; its
                                  lost updates become atomic adds here (red.global.add on both rows) */
};
enum {
    MML_GROUPS_PERM_MOD = 0,   /* group = perm[id] % groups, the reference rule (MultiCore.cs:64) */
    MML_GROUPS_BALANCED = 1    /* same stratification, ids dealt to groups balanced by rating count */
};

enum {
    MML_INTRA_ROUNDS = 0,      /* inside a block: rounds of mutually independent ratings (conflict-free, deterministic:
                                  the epoch equals a serial pass in the order mml_sgd_schedule_dump reports) */
    MML_INTRA_ASYNC  = 1       /* inside a block: user rows exclusive per worker, item rows updated without ordering
                                  (lock-free, as the reference's NaiveParallelization, but confined to a block) */
};

/* Hyper-parameters: names and defaults are those of the reference properties
 * (MatrixFactorization.cs:87-96, BiasedMatrixFactorization.cs:85-141). */
typedef struct mml_mf_params {
    int32_t biased;               /* 1 = BiasedMatrixFactorization, 0 = MatrixFactorization */
    int32_t num_factors;          /* NumFactors = 10 */
    float   learn_rate;           /* LearnRate = 0.01 */
    float   decay;                /* Decay = 1.0 */
    float   regularization;       /* Regularization = 0.015 (plain MF) */
    float   bias_learn_rate;      /* BiasLearnRate = 1.0 */
    float   bias_reg;             /* BiasReg = 0.01 */
    float   reg_u, reg_i;         /* RegU = RegI = Regularization */
    int32_t frequency_regularization;
    int32_t loss;                 /* MML_LOSS_* */
    int32_t bold_driver;
    int32_t max_threads;          /* MaxThreads: 1 = UpdateLearnRate once per Iterate(); > 1 = the reference's
                                     multi-threaded semantics (learn rate updated twice per epoch, :216/:221) */
    /* engine knobs (not reference options) */
    int32_t schedule;             /* MML_SCHEDULE_* */
    int32_t num_groups;           /* G: DSGD worker groups on this GPU; 0 = SMs / ctas_per_group */
    int32_t num_subgroups;        /* warps per worker group (a warp runs 32/L ratings of a round at once); 0 = 8 */
    int32_t group_rule;           /* MML_GROUPS_* */
    int32_t persistent;           /* 1 = one cooperative launch per epoch with neighbour flags instead of
                                     one launch per sub-epoch; 0 = per-sub-epoch launches; -1 = auto */
    float   hot_item_factor;      /* an item whose ratings per block reach factor x (ratings per block / workers)
                                     is "hot": inside a block its updates run on hot_copies private copies of the
                                     row and the copies' deltas are summed at the end of the block (they would
                                     otherwise form one serial chain); 0 = off (default): the copies trade RMSE
                                     parity for speed, see DESIGN.md */
    int32_t hot_copies;           /* private copies per hot item and block; 0 = 8 */
    int32_t intra_block;          /* MML_INTRA_*; default MML_INTRA_ASYNC */
    int32_t hot_merge_average;    /* hot-item copies are merged by averaging (1, default) or summing (0) their steps */
    int32_t async_workers;        /* async mode: workers per worker group that take part; 0 = all, 1 = serial
                                     inside a block (deterministic; used by the parity tests) */
    int32_t ctas_per_group;       /* async mode: CTAs that make up one worker group and share its blocks
                                     (num_groups = 0 then means SMs / ctas_per_group groups); 0 = 4 when
                                     num_groups is 0 too, else 1.
                                     num_groups = 1 makes the whole GPU one worker group: no block hand-over,
                                     the reference's NaiveParallelization inside the GPU */
} mml_mf_params;

void mml_mf_params_default(mml_mf_params* p);

/* Train() part 1 (BiasedMatrixFactorization.cs:173-190 / MatrixFactorization.cs:119-126): allocates
 * the model for `ratings`, sets rating_range_size and global_bias, builds the DSGD strata when
 * schedule = MML_SCHEDULE_DSGD. user_perm/item_perm (host, may be NULL = identity) are the
 * permutations PartitionUsersAndItems would draw (MultiCore.cs:51-52). */
int32_t mml_sgd_create(mml_ctx* ctx, mml_ratings* ratings, const mml_mf_params* p,
                       const int32_t* user_perm, const int32_t* item_perm, mml_sgd** out);
int32_t mml_sgd_destroy(mml_sgd* m);
/* InitModel (MatrixFactorization.cs:99-116, BiasedMatrixFactorization.cs:161-170) with factors
 * supplied by the host's RNG (user matrix first, then item matrix, row-major, num_factors columns).
 * Rows of entities without ratings are zeroed; biases (may be NULL) default to zero;
 * current_learnrate = LearnRate. */
int32_t mml_sgd_set_model(mml_sgd* m, const float* user_factors, const float* item_factors,
                          const float* user_bias, const float* item_bias);
/* Same InitModel, factors drawn on device ~ N(mean, stddev) from a counter-based generator. */
int32_t mml_sgd_init_model(mml_sgd* m, uint64_t seed, double init_mean, double init_stddev);
/* Any output pointer may be NULL. */
int32_t mml_sgd_get_model(mml_sgd* m, float* user_factors, float* item_factors,
                          float* user_bias, float* item_bias, float* global_bias, float* current_learnrate);
int32_t mml_sgd_set_learnrate(mml_sgd* m, float current_learnrate);
/* LoadModel (BiasedMatrixFactorization.cs:353-402): the rating scale and global bias come from the model file. */
int32_t mml_sgd_set_scale(mml_sgd* m, float min_rating, float max_rating, float global_bias);
/* Iterate() (BiasedMatrixFactorization.cs:197-222 / MatrixFactorization.cs:135-138): one epoch +
 * UpdateLearnRate. DSGD schedule: subepoch_sequence (host, G entries, may be NULL = 0..G-1) is the
 * shuffled sub-epoch order of :210-211. Serial schedule: random_index (host, n_index entries) is
 * ratings.RandomIndex; it is uploaded once and cached until a different length is passed or
 * mml_sgd_invalidate_index is called. */
/* Naive schedule: random_index as for the serial schedule; the lists (one per worker of the GPU) are cut from it on the
 * device the first time and cached like the serial index. */
int32_t mml_sgd_iterate(mml_sgd* m, const int32_t* subepoch_sequence, const int32_t* random_index, int64_t n_index);
int32_t mml_sgd_invalidate_index(mml_sgd* m);
/* Iterate(IList<int>, bool, bool) (BiasedMatrixFactorization.cs:264-310, MatrixFactorization.cs:166-196)
 * in the reference's exact order and mixed precision; no learn-rate update for the biased model,
 * plain MF decays as the reference does (:195). */
int32_t mml_sgd_iterate_indices(mml_sgd* m, const int32_t* indices, int64_t n, int32_t update_user, int32_t update_item);
/* LearnFactors (MatrixFactorization.cs:198-202): num_iter (= NumIter) passes of Iterate(list, update_user, update_item) --
 * what RetrainUser / RetrainItem run over ByUser[u] / ByItem[i] after re-drawing the row (:141-160). Plain MF decays the
 * learn rate after every pass (:195), the biased model leaves it alone. */
int32_t mml_sgd_learn_factors(mml_sgd* m, const int32_t* indices, int64_t n, int32_t update_user, int32_t update_item,
                              int32_t num_iter);
/* Predict (BiasedMatrixFactorization.cs:313-325 / MatrixFactorization.cs:251-259), batched. */
int32_t mml_sgd_predict(mml_sgd* m, const int32_t* users, const int32_t* items, int64_t n, float* out);
/* Eval.Ratings.Evaluate (Eval/Ratings.cs:96-139): out4 = {RMSE, MAE, NMAE, CBD}. */
int32_t mml_sgd_evaluate(mml_sgd* m, const int32_t* users, const int32_t* items, const float* values, int64_t n, float* out4);
/* Same on the training set already resident in HBM (ComputeFit, Eval/Ratings.cs:164-170). */
int32_t mml_sgd_evaluate_train(mml_sgd* m, float* out4);
/* ComputeObjective (BiasedMatrixFactorization.cs:515-552). */
int32_t mml_sgd_objective(mml_sgd* m, double* out);
/* Number of kernels this model launched so far and the device time of the last iterate (ms). */
int32_t mml_sgd_stats(mml_sgd* m, int64_t* kernel_launches, float* last_iterate_ms);
/* Strata shape: G, warps per group, number of rounds, staged item block bytes (0 = item rows stay in global memory). */
int32_t mml_sgd_strata_info(mml_sgd* m, int32_t* G, int32_t* W, int64_t* n_rounds, int64_t* staged_bytes);
/* Grid of the DSGD epoch kernel: G worker groups x ctas_per_group CTAs (as resolved from the params' defaults). */
int32_t mml_sgd_grid(mml_sgd* m, int32_t* G, int32_t* ctas_per_group);
/* Strata shape: ..., number of hot items. */
int32_t mml_sgd_hot_items(mml_sgd* m, int64_t* n_hot);
/* The serial-equivalent order of one DSGD epoch with the given sub-epoch sequence (NULL = 0..G-1):
 * order[n] receives rating indices such that processing them one after the other gives the same
 * model as the conflict-free parallel schedule (tests replay it through the oracle).
 * block[n] (may be NULL) = id of the (sub-epoch, worker group) block each entry belongs to;
 * copy[n] (may be NULL) = -1 for ordinary items, else the private copy of the hot item's row the entry
 * updates (copies start from the row at block start and their deltas are summed at block end);
 * round[n] (may be NULL) = id of the round (set of mutually independent ratings) the entry runs in. */
int32_t mml_sgd_schedule_dump(mml_sgd* m, const int32_t* subepoch_sequence, int32_t* order,
                              int32_t* block, int32_t* copy, int32_t* round);

/* ---- fold-in and incremental updates (IFoldInRatingPredictor / IncrementalRatingPredictor) -------------------------- */
/* FoldIn (MatrixFactorization.cs:323-347, BiasedMatrixFactorization.cs:445-492) for a batch of users described by
 * ratings, one warp per user, the reference's arithmetic operation by operation. rated_ptr[n_users + 1] delimits each
 * user's (item, rating) pairs in rated_items / rated_values, in the order AFTER rated_items.Shuffle() (the host's RNG draws
 * it, as it draws init_factors[n_users x num_factors] = the InitNormal vector of each user). num_iter = NumIter.
 * out_vectors: n_users x num_factors for MatrixFactorization; n_users x (num_factors + 1) for the biased model, entry 0
 * the user bias (FOLD_IN_BIAS_INDEX = 0, FOLD_IN_FACTORS_START = 1, BiasedMatrixFactorization.cs:80-82).
 * A rated item outside the model is MML_ERR_ARG (the reference throws from the matrix indexer). Model unchanged. */
int32_t mml_sgd_fold_in(mml_sgd* m, const int64_t* rated_ptr, const int32_t* rated_items, const float* rated_values,
                        int64_t n_users, const float* init_factors, int32_t num_iter, float* out_vectors);
/* Predict(float[] user_vector, int item_id) (MatrixFactorization.cs:223-241, BiasedMatrixFactorization.cs:328-336) for
 * every fold-in vector and every candidate -- the scoring half of ScoreItems (MatrixFactorization.cs:350-363).
 * out_scores: n_users x n_cand, row-major. */
int32_t mml_sgd_score_items(mml_sgd* m, const float* user_vectors, int64_t n_users,
                            const int32_t* candidates, int64_t n_cand, float* out_scores);
/* Overwrites the factor rows (and biases; either pointer may be NULL = unchanged) of the given users (by_item = 0) or
 * items (1): the row re-initialisation of RetrainUser / RetrainItem (MatrixFactorization.cs:141-160,
 * BiasedMatrixFactorization.cs:419-431; the host draws RowInitNormal and then calls mml_sgd_iterate_indices on
 * ByUser[u] / ByItem[i] with update_user / update_item set accordingly) and the zeroing of RemoveUser / RemoveItem
 * (MatrixFactorization.cs:300-316, BiasedMatrixFactorization.cs:433-445). factors: n x num_factors. */
int32_t mml_sgd_set_rows(mml_sgd* m, int32_t by_item, const int32_t* ids, int64_t n, const float* factors, const float* biases);

/* ---- Top-N Recommend() ---------------------------------------------------------------------- */
/* Recommender.Recommend (Recommender.cs:52-103) for a batch of users on an item-MF model
 * (score = RowScalarProduct, ItemRecommendation/MF.cs:151-157): for every user the top n
 * candidates not in its ignore list, ordered by (score desc, candidate position asc).
 * n > 0 or n = -1 (all candidates with a score > float.MinValue); out arrays hold n_users * n_out entries,
 * n_out = n_cand for n = -1, else min(n, n_cand).
 * candidates: n_cand item ids (NULL = 0..n_model_items-1; the reference's own default, Enumerable.Range(0, MaxItemID - 1),
 * drops the last two items -- every in-tree caller passes an explicit list, and so should hosts of this library). ignore CSR: ignore_ptr[n_users+1], ignore_idx
 * (may be NULL). out_items/out_scores: n_users * n entries; out_counts[n_users]. */
int32_t mml_topn_mf(mml_ctx* ctx, const float* user_factors, int32_t n_model_users,
                    const float* item_factors, int32_t n_model_items, int32_t k,
                    const int32_t* users, int64_t n_users, int32_t n,
                    const int32_t* candidates, int64_t n_cand,
                    const int64_t* ignore_ptr, const int32_t* ignore_idx,
                    int32_t* out_items, float* out_scores, int32_t* out_counts);

/* Engine knob (not a reference option): which scoring path Recommend() uses. AUTO = tcgen05 TF32 scoring with the
 * top-k selection fused into the GEMM epilogue, finalists re-scored exactly (n <= 16, num_factors <= 128, distinct
 * candidates), exact CUDA-core scoring otherwise and for users whose candidate superset cannot be proven complete.
 * Both paths return bit-identical results; EXACT / TENSOR force one of them (tests, benchmarks). */
enum { MML_TOPN_AUTO = 0, MML_TOPN_EXACT = 1, MML_TOPN_TENSOR = 2 };
int32_t mml_topn_set_mode(int32_t mode);
/* Engine knob: operand precision of the tensor-core filter GEMM: TF32 (default) or BF16 (MMAs at twice the rate on half
 * the operand bytes; its wider error bound, 0.0045 |u| max|v| against 0.0025, lengthens the list of finalists that are
 * re-scored exactly). The results are the same bits either way. */
enum { MML_TOPN_FILTER_BF16 = 0, MML_TOPN_FILTER_TF32 = 1 };
int32_t mml_topn_set_filter(int32_t kind);
/* Users served by each path in the last Recommend() call of this process and the device time of the tensor path. */
int32_t mml_topn_last_stats(int64_t* users_tensor_path, int64_t* users_exact_path, float* tensor_path_ms);

/* ---- item-ranking evaluation ------------------------------------------------------------------------------ */
/* Eval.Items.Evaluate (Eval/Items.cs:126-209) for an item-MF model: for every test user the ranking of `candidates`
 * (distinct ids; the host draws Items.Candidates' shuffle) without the user's ignore row (its training items, NULL =
 * RepeatedEvents.Yes), cut to the n best (n = -1: whole list), is compared with the user's test row. The list is never
 * materialised: the ranks of the test items are counted from the exact score rows (score desc, candidate position asc).
 * test_ptr[n_test_users + 1] / test_idx: test_user_matrix rows aligned with test_users; ignore_ptr / ignore_idx likewise.
 * out_measures: n_test_users x 8 = {AUC, MAP, NDCG, MRR, prec@5, prec@10, recall@5, recall@10} as the (float) values the
 * reference adds up (Eval/Measures/AUC.cs, PrecisionAndRecall.cs, NDCG.cs, ReciprocalRank.cs); out_used[u] = 1 when the
 * user counts (num_users), 0 when the reference skips it (no test item among the candidates, or nothing but test items).
 * The host sums the rows with out_used = 1 and divides by their number. */
int32_t mml_items_evaluate_mf(mml_ctx* ctx, const float* user_factors, int32_t n_model_users,
                              const float* item_factors, int32_t n_model_items, int32_t k,
                              const int32_t* test_users, int64_t n_test_users,
                              const int32_t* candidates, int64_t n_cand,
                              const int64_t* test_ptr, const int32_t* test_idx,
                              const int64_t* ignore_ptr, const int32_t* ignore_idx, int32_t n,
                              float* out_measures, int32_t* out_used);

/* ---- WRMF ------------------------------------------------------------------------------------- */
/* PosOnlyFeedback.UserMatrix / ItemMatrix (Data/PosOnlyFeedback.cs:35-83): duplicates collapse. */
int32_t mml_feedback_create(mml_ctx* ctx, const int32_t* users, const int32_t* items, int64_t n,
                            int32_t max_user, int32_t max_item, mml_feedback** out);
int32_t mml_feedback_destroy(mml_feedback* f);
int32_t mml_feedback_nnz(mml_feedback* f, int64_t* nnz);
/* UserMatrix (by_item = 0) / ItemMatrix (by_item = 1) as CSR: row_ptr[rows + 1], cols[nnz] ascending inside a row. */
int32_t mml_feedback_csr(mml_feedback* f, int32_t by_item, int64_t* row_ptr, int32_t* cols);

typedef struct mml_wrmf_params {
    int32_t num_factors;      /* NumFactors = 10 (ItemRecommendation/MF.cs:43-45) */
    double  alpha;            /* Alpha = 1 (WRMF.cs:56) */
    double  regularization;   /* Regularization = 0.015 (WRMF.cs:59) */
} mml_wrmf_params;

int32_t mml_wrmf_create(mml_ctx* ctx, mml_feedback* f, const mml_wrmf_params* p, mml_wrmf** out);
int32_t mml_wrmf_destroy(mml_wrmf* m);
int32_t mml_wrmf_set_model(mml_wrmf* m, const float* user_factors, const float* item_factors);
int32_t mml_wrmf_init_model(mml_wrmf* m, uint64_t seed, double init_mean, double init_stddev);
int32_t mml_wrmf_get_model(mml_wrmf* m, float* user_factors, float* item_factors);
/* WRMF.Iterate (WRMF.cs:68-73): user half-sweep then item half-sweep. */
int32_t mml_wrmf_iterate(mml_wrmf* m);
int32_t mml_wrmf_stats(mml_wrmf* m, int64_t* kernel_launches, float* last_iterate_ms);
/* RetrainUser / RetrainItem (WRMF.cs:159-170): the Gram matrix of the other side, then Optimize() for each given row --
 * a half-sweep restricted to `ids` (users when by_item = 0, items when 1). On a multi-GPU context every rank solves the
 * given rows itself (the model is replicated), no collective. */
int32_t mml_wrmf_retrain(mml_wrmf* m, int32_t by_item, const int32_t* ids, int64_t n);
/* Multi-GPU contexts (mml_ctx_create_dist): every rank passes the same feedback and model; in each half-sweep
 * (WRMF.cs:79-92, a Parallel.For over independent rows) rank r solves the contiguous rows
 * [ranges[r], ranges[r + 1]) -- balanced by events per row -- and the ranks all-gather the solved rows (NCCL), so
 * every rank holds the whole model after mml_wrmf_iterate. ranges receives world + 1 entries. */
int32_t mml_wrmf_shard(mml_wrmf* m, int32_t by_item, int32_t* ranges);
/* Engine knob: AUTO = per-row Gram sums sum_{i in S_u} h_i h_i^T on the tcgen05 tensor cores (3 x TF32 split, fp32-accurate)
 * with HH and the assembly in double, a blocked Cholesky factor in single precision used as the preconditioner of an
 * iterative refinement against the exact double-precision operator (num_factors a multiple of 4, <= 128; a half-sweep
 * with a row that does not converge is repeated with the double-precision factor), the all-double CUDA-core kernels
 * otherwise; FP64 / TENSOR force one of them, TENSOR_F64 = TENSOR with the double-precision factor, TENSOR_PCG = TENSOR with the
 * row systems solved by conjugate gradients preconditioned with (HH + lambda I)^-1 and one refinement against the exact operator
 * (rows within 2e-7 of the double solve at config 3, but 62 ms per epoch against 52 ms: measured, not the default). */
enum { MML_WRMF_AUTO = 0, MML_WRMF_FP64 = 1, MML_WRMF_TENSOR = 2, MML_WRMF_TENSOR_F64 = 3, MML_WRMF_TENSOR_PCG = 4 };
int32_t mml_wrmf_set_mode(int32_t mode);
/* Diagnostic for the parity tests: the tensor-core Gram sum (128 x 128 floats, zero beyond num_factors) of the user with
 * the most events, and that user's id. The model is not modified. */
int32_t mml_wrmf_debug_gram(mml_wrmf* m, float* out_gram, int32_t* out_user);
/* mml_items_evaluate_mf on the device-resident model (no factor upload). */
int32_t mml_wrmf_evaluate(mml_wrmf* m, const int32_t* test_users, int64_t n_test_users,
                          const int32_t* candidates, int64_t n_cand,
                          const int64_t* test_ptr, const int32_t* test_idx,
                          const int64_t* ignore_ptr, const int32_t* ignore_idx, int32_t n,
                          float* out_measures, int32_t* out_used);
/* mml_topn_mf on the device-resident model (no factor upload). */
int32_t mml_wrmf_recommend(mml_wrmf* m, const int32_t* users, int64_t n_users, int32_t n,
                           const int32_t* candidates, int64_t n_cand,
                           const int64_t* ignore_ptr, const int32_t* ignore_idx,
                           int32_t* out_items, float* out_scores, int32_t* out_counts);

#ifdef __cplusplus
}
#endif
#endif /* MMLB200_H */
