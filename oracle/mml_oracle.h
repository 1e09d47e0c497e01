/*
 * mml_oracle.h -- CPU restatement of MyMediaLite's matrix-factorization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / the timed CPU arm.  The product (libmmlb200.so) never
 * links, loads or calls this code.
 *
 * Every function cites the reference file:line (relative to /root/reference/src/MyMediaLite/)
 * whose arithmetic and ordering it follows.
 *
 * PARITY PIN STATUS
 *   - System.Random (BCL, not in the reference tree): pinned by public known-answer values
 *     (seed 0/1/42 first outputs) in tests/test_oracle_rng.py.
 *   - Dot / Inc / counts / by-user / partition shapes / learn-rate schedule: pinned by the
 *     reference's own NUnit known answers (src/Tests/...), restated in tests/test_oracle_*.py.
 *   - SGD / ALS / top-N numeric outputs: the reference's tests record none and no .NET runtime
 *     exists in this image, so those are "parity unpinned": this restatement is the only pin.
 *   - MathNet Normal.Sample (polar Box-Muller) and DenseMatrix.Inverse, C5 IntervalHeap tie
 *     order: third-party binaries without source in the tree -> "parity unpinned".
 */
#ifndef MML_ORACLE_H
#define MML_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- System.Random (Knuth subtractive), used through MyMediaLite.Random (Random.cs:23-64) ---- */
typedef struct mo_rng {
    int32_t seed_array[56];
    int32_t inext, inextp;
} mo_rng;

void    mo_rng_init(mo_rng* r, int32_t seed);
int32_t mo_rng_next(mo_rng* r);                    /* Random.Next()            */
double  mo_rng_next_double(mo_rng* r);             /* Random.NextDouble()      */
int32_t mo_rng_next_max(mo_rng* r, int32_t max);   /* Random.Next(maxValue)    */

/* Utils.Shuffle (Utils.cs:52-64): for i=n-1..0 { r=Next(i+1); swap(a[i],a[r]); } */
void mo_shuffle_i32(mo_rng* r, int32_t* a, int64_t n);
/* Only the swap targets of the same loop: H[i] = Next(i+1), drawn for i = n-1 .. 0. */
void mo_shuffle_targets(mo_rng* r, int32_t* H, int64_t n);
/* Apply swap targets sequentially (the definition K3 shuffle_apply must reproduce bit-exactly). */
void mo_shuffle_apply(int32_t* a, const int32_t* H, int64_t n);

/* MathNet.Numerics 3.15 Normal.Sample (polar Box-Muller, two NextDouble per trial);
 * call sites DataType/MatrixExtensions.cs:62-69. */
double mo_normal_sample(mo_rng* r, double mean, double stddev);
void   mo_init_normal(mo_rng* r, float* data, int64_t n, double mean, double stddev);

/* ---- data set (Data/DataSet.cs:134-202, Data/Ratings.cs:76-84, Data/RatingScale.cs:104-117) ---- */
void  mo_count_by(const int32_t* ids, int64_t n, int32_t max_id, int32_t* counts);
/* CSR analogue of ByUser/ByItem: row_ptr[max_id+2], idx[n] = rating indices ascending per row. */
void  mo_build_index(const int32_t* ids, int64_t n, int32_t max_id, int64_t* row_ptr, int32_t* idx);
float mo_average(const float* values, int64_t n);
void  mo_scale(const float* values, int64_t n, float* min_out, float* max_out);
void  mo_random_index(mo_rng* r, int32_t* index, int64_t n);      /* DataSet.cs:193-202 */

/* MultiCore.PartitionUsersAndItems (MultiCore.cs:43-73). Returns g (clamped num_groups).
 * block_ptr has g*g+1 entries (row-major [i*g+j]); idx has n entries; draws the RNG in the
 * reference order (user perm, item perm, then block (0,0)..(g-1,g-1) shuffles).
 * user_perm/item_perm (may be NULL) receive the permutations used. */
int32_t mo_partition_users_and_items(mo_rng* r, const int32_t* users, const int32_t* items, int64_t n,
                                     int32_t max_user, int32_t max_item, int32_t num_groups,
                                     int64_t* block_ptr, int32_t* idx,
                                     int32_t* user_perm, int32_t* item_perm);
/* Same block membership/order as above but with given permutations and NO shuffling of the
 * blocks (ascending rating index inside each block): the bit-exact twin of K2 block_partition. */
void mo_partition_blocks_given(const int32_t* users, const int32_t* items, int64_t n,
                               const int32_t* user_perm, const int32_t* item_perm, int32_t g,
                               int64_t* block_ptr, int32_t* idx);
/* MultiCore.PartitionIndices (MultiCore.cs:79-92): returns number of groups. */
int32_t mo_partition_indices(const int32_t* random_index, int64_t n, int32_t num_groups,
                             int64_t* group_ptr, int32_t* idx);

/* ---- rating prediction: MatrixFactorization / BiasedMatrixFactorization ---- */
enum { MO_LOSS_RMSE = 0, MO_LOSS_MAE = 1, MO_LOSS_LOGISTIC = 2 };

typedef struct mo_mf_params {
    int32_t num_factors;        /* NumFactors = 10            MatrixFactorization.cs:87-96 */
    float   learn_rate;         /* LearnRate = 0.01f */
    float   decay;              /* Decay = 1.0f */
    float   regularization;     /* Regularization = 0.015f (plain MF) */
    int32_t num_iter;           /* NumIter = 30 */
    double  init_mean;          /* InitMean = 0 */
    double  init_stddev;        /* InitStdDev = 0.1 */
    /* BiasedMatrixFactorization.cs:85-141 */
    float   bias_learn_rate;    /* 1.0f */
    float   bias_reg;           /* 0.01f */
    float   reg_u, reg_i;       /* = Regularization */
    int32_t frequency_regularization;
    int32_t loss;               /* MO_LOSS_* */
    int32_t max_threads;        /* 1 */
    int32_t bold_driver;
    int32_t naive_parallelization;
    int32_t omp_threads;        /* oracle-only: >1 runs DSGD blocks of a sub-epoch on that many OpenMP threads */
} mo_mf_params;

void mo_mf_params_default(mo_mf_params* p);

typedef struct mo_model mo_model;   /* opaque: holds factors, biases, schedule state */

/* biased != 0 -> BiasedMatrixFactorization, else MatrixFactorization. The data arrays are borrowed. */
mo_model* mo_model_create(int biased, const mo_mf_params* p,
                          const int32_t* users, const int32_t* items, const float* values, int64_t n,
                          int32_t max_user, int32_t max_item);
void  mo_model_destroy(mo_model* m);
void  mo_model_init(mo_model* m, mo_rng* r);        /* InitModel (+partition and global bias of Train) */
void  mo_model_train(mo_model* m, mo_rng* r);       /* Train() */
void  mo_model_iterate(mo_model* m, mo_rng* r);     /* Iterate() */
float mo_model_predict(const mo_model* m, int32_t u, int32_t i);
void  mo_model_predict_many(const mo_model* m, const int32_t* u, const int32_t* i, int64_t n, float* out);
/* Eval/Ratings.cs:96-139: out[0]=RMSE out[1]=MAE out[2]=NMAE out[3]=CBD */
void  mo_model_evaluate(const mo_model* m, const int32_t* u, const int32_t* i, const float* v, int64_t n, float* out4);
float mo_model_objective(const mo_model* m);        /* BiasedMatrixFactorization.cs:515-552 */
float mo_model_learnrate(const mo_model* m);
float mo_model_global_bias(const mo_model* m);
float* mo_model_user_factors(mo_model* m);
float* mo_model_item_factors(mo_model* m);
float* mo_model_user_bias(mo_model* m);
float* mo_model_item_bias(mo_model* m);
const int32_t* mo_model_random_index(mo_model* m);  /* NULL until built */
/* One pass of Iterate(IList<int>,bool,bool) over an explicit index list (no lr update for BMF;
 * plain MF applies its UpdateLearnRate as the reference does, MatrixFactorization.cs:195). */
void  mo_model_fold_in(const mo_model* m, const int32_t* items, const float* values, int64_t n, const float* init, float* out);
float mo_model_predict_vector(const mo_model* m, const float* user_vector, int32_t item_id);
void  mo_model_iterate_indices(mo_model* m, const int32_t* idx, int64_t n_idx, int update_user, int update_item);
/* Mini-batch replay of the GPU stratum schedule, used ONLY to check the CUDA kernel's arithmetic:
 * ratings (user,value) of one item run are consumed `batch` at a time against the same q_i/b_i. */
void  mo_bmf_replay_runs(mo_model* m, const int32_t* run_item, const int64_t* run_ptr, int64_t n_runs,
                         const int32_t* ent_user, const float* ent_value, int32_t batch);

/* ---- item recommendation: WRMF (ItemRecommendation/WRMF.cs:56-156, MF.cs:51-67,151-157) ---- */
/* feedback is given as CSR by user and CSR by item (duplicates already collapsed). */
void mo_wrmf_optimize(const int64_t* row_ptr, const int32_t* cols, int32_t n_rows,
                      float* W, const float* H, int32_t n_h_rows, int32_t k,
                      double alpha, double regularization, int omp_threads);
void mo_wrmf_gram(const float* H, int32_t n_rows, int32_t k, double* HH);
/* Collapse duplicate (row,col) events keeping first-seen order: SparseBooleanMatrix (HashSet rows). */
int64_t mo_feedback_csr(const int32_t* rows, const int32_t* cols, int64_t n, int32_t max_row,
                        int64_t* row_ptr, int32_t* out_cols);

/* ---- Recommender.Recommend (Recommender.cs:52-103) with (score desc, candidate position asc) ---- */
/* scores of item MF: dot(U[u],V[i]) sequential fp32 mul-then-add (MatrixExtensions.cs:224-241).
 * n < 0 -> full ranking. Returns number of results written (<= n_cand). */
int64_t mo_recommend_mf(const float* U, const float* V, int32_t k, int32_t n_items_model, int32_t n_users_model,
                        int32_t user, int32_t n,
                        const int32_t* candidates, int64_t n_cand,
                        const int32_t* ignore, int64_t n_ignore,
                        int32_t* out_items, float* out_scores);
/* Same, but scores come from a rating predictor model (rating_based_ranking). */
int64_t mo_recommend_model(const mo_model* m, int32_t user, int32_t n,
                           const int32_t* candidates, int64_t n_cand,
                           const int32_t* ignore, int64_t n_ignore,
                           int32_t* out_items, float* out_scores);

float mo_row_scalar_product(const float* a, const float* b, int32_t k);  /* MatrixExtensions.cs:224-241 */

#ifdef __cplusplus
}
#endif
#endif
