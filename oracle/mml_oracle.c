/*
 * mml_oracle.c -- CPU restatement of MyMediaLite's matrix-factorization hot path.
 * TEST INFRASTRUCTURE ONLY (see mml_oracle.h). Build: see oracle/Makefile
 * (-O2 -ffp-contract=off: fp32 operations stay fp32 and are never fused, as on the CLR x64 JIT).
 *
 * Citations are relative to /root/reference/src/MyMediaLite/.
 */
#include "mml_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * System.Random -- .NET BCL (not in the reference tree; call sites Random.cs:23-35, Utils.cs:54-57).
 * Knuth subtractive generator as published in the .NET reference source.
 * ------------------------------------------------------------------------------------------ */
#define MO_MBIG  2147483647
#define MO_MSEED 161803398

void mo_rng_init(mo_rng* r, int32_t seed)
{
    int32_t subtraction = (seed == INT32_MIN) ? INT32_MAX : (seed < 0 ? -seed : seed);
    int32_t mj = MO_MSEED - subtraction;
    int32_t mk = 1;
    memset(r->seed_array, 0, sizeof(r->seed_array));
    r->seed_array[55] = mj;
    for (int i = 1; i < 55; i++) {
        int ii = (21 * i) % 55;
        r->seed_array[ii] = mk;
        mk = mj - mk;
        if (mk < 0) mk += MO_MBIG;
        mj = r->seed_array[ii];
    }
    for (int k = 1; k < 5; k++)
        for (int i = 1; i < 56; i++) {
            /* C# int arithmetic wraps silently */
            r->seed_array[i] = (int32_t)((uint32_t)r->seed_array[i] - (uint32_t)r->seed_array[1 + (i + 30) % 55]);
            if (r->seed_array[i] < 0) r->seed_array[i] += MO_MBIG;
        }
    r->inext = 0;
    r->inextp = 21;
}

int32_t mo_rng_next(mo_rng* r)
{
    int32_t inext = r->inext, inextp = r->inextp;
    if (++inext >= 56) inext = 1;
    if (++inextp >= 56) inextp = 1;
    int32_t v = r->seed_array[inext] - r->seed_array[inextp];
    if (v == MO_MBIG) v--;
    if (v < 0) v += MO_MBIG;
    r->seed_array[inext] = v;
    r->inext = inext;
    r->inextp = inextp;
    return v;
}

double mo_rng_next_double(mo_rng* r) { return mo_rng_next(r) * (1.0 / MO_MBIG); }

int32_t mo_rng_next_max(mo_rng* r, int32_t max) { return (int32_t)(mo_rng_next_double(r) * max); }

/* Utils.cs:52-64 */
void mo_shuffle_i32(mo_rng* r, int32_t* a, int64_t n)
{
    for (int64_t i = n - 1; i >= 0; i--) {
        int32_t j = mo_rng_next_max(r, (int32_t)(i + 1));
        int32_t t = a[i]; a[i] = a[j]; a[j] = t;
    }
}

void mo_shuffle_targets(mo_rng* r, int32_t* H, int64_t n)
{
    for (int64_t i = n - 1; i >= 0; i--) H[i] = mo_rng_next_max(r, (int32_t)(i + 1));
}

void mo_shuffle_apply(int32_t* a, const int32_t* H, int64_t n)
{
    for (int64_t i = n - 1; i >= 0; i--) {
        int32_t j = H[i];
        int32_t t = a[i]; a[i] = a[j]; a[j] = t;
    }
}

/* MathNet.Numerics 3.15.0 Distributions.Normal.Sample -> SampleUnchecked -> PolarTransform
 * (binary only in the tree: src/packages/MathNet.Numerics.3.15.0; restated from the published
 * algorithm: polar Box-Muller, second variate discarded). */
double mo_normal_sample(mo_rng* r, double mean, double stddev)
{
    for (;;) {
        double a = mo_rng_next_double(r);
        double b = mo_rng_next_double(r);
        double v1 = (2.0 * a) - 1.0;
        double v2 = (2.0 * b) - 1.0;
        double rr = (v1 * v1) + (v2 * v2);
        if (rr >= 1.0 || rr == 0.0) continue;
        double fac = sqrt(-2.0 * log(rr) / rr);
        return mean + stddev * (v1 * fac);
    }
}

/* DataType/MatrixExtensions.cs:62-69 */
void mo_init_normal(mo_rng* r, float* data, int64_t n, double mean, double stddev)
{
    for (int64_t i = 0; i < n; i++) data[i] = (float)mo_normal_sample(r, mean, stddev);
}

/* ------------------------------------------------------------------------------------------
 * Data set
 * ------------------------------------------------------------------------------------------ */
/* Data/DataSet.cs:134-169 */
void mo_count_by(const int32_t* ids, int64_t n, int32_t max_id, int32_t* counts)
{
    memset(counts, 0, sizeof(int32_t) * (size_t)(max_id + 1));
    for (int64_t i = 0; i < n; i++) counts[ids[i]]++;
}

/* Data/DataSet.cs:171-191: one forward pass appending -> ascending rating index per row */
void mo_build_index(const int32_t* ids, int64_t n, int32_t max_id, int64_t* row_ptr, int32_t* idx)
{
    int64_t rows = (int64_t)max_id + 1;
    memset(row_ptr, 0, sizeof(int64_t) * (size_t)(rows + 1));
    for (int64_t i = 0; i < n; i++) row_ptr[ids[i] + 1]++;
    for (int64_t r = 0; r < rows; r++) row_ptr[r + 1] += row_ptr[r];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(rows > 0 ? rows : 1));
    memcpy(cur, row_ptr, sizeof(int64_t) * (size_t)rows);
    for (int64_t i = 0; i < n; i++) idx[cur[ids[i]]++] = (int32_t)i;
    free(cur);
}

/* Data/Ratings.cs:76-84: (float) sum / Count -- the cast binds to sum */
float mo_average(const float* values, int64_t n)
{
    double sum = 0;
    for (int64_t i = 0; i < n; i++) sum += values[i];
    return (float)sum / (float)n;
}

/* Data/RatingScale.cs:104-117 (Min/Max of the sorted distinct levels) */
void mo_scale(const float* values, int64_t n, float* min_out, float* max_out)
{
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int64_t i = 0; i < n; i++) {
        if (values[i] < mn) mn = values[i];
        if (values[i] > mx) mx = values[i];
    }
    *min_out = mn; *max_out = mx;
}

/* Data/DataSet.cs:193-202 */
void mo_random_index(mo_rng* r, int32_t* index, int64_t n)
{
    for (int64_t i = 0; i < n; i++) index[i] = (int32_t)i;
    mo_shuffle_i32(r, index, n);
}

/* MultiCore.cs:43-73 */
void mo_partition_blocks_given(const int32_t* users, const int32_t* items, int64_t n,
                               const int32_t* user_perm, const int32_t* item_perm, int32_t g,
                               int64_t* block_ptr, int32_t* idx)
{
    int64_t nb = (int64_t)g * g;
    memset(block_ptr, 0, sizeof(int64_t) * (size_t)(nb + 1));
    for (int64_t t = 0; t < n; t++) {
        int64_t b = (int64_t)(user_perm[users[t]] % g) * g + (item_perm[items[t]] % g);
        block_ptr[b + 1]++;
    }
    for (int64_t b = 0; b < nb; b++) block_ptr[b + 1] += block_ptr[b];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
    memcpy(cur, block_ptr, sizeof(int64_t) * (size_t)nb);
    for (int64_t t = 0; t < n; t++) {
        int64_t b = (int64_t)(user_perm[users[t]] % g) * g + (item_perm[items[t]] % g);
        idx[cur[b]++] = (int32_t)t;
    }
    free(cur);
}

int32_t mo_partition_users_and_items(mo_rng* r, const int32_t* users, const int32_t* items, int64_t n,
                                     int32_t max_user, int32_t max_item, int32_t num_groups,
                                     int64_t* block_ptr, int32_t* idx,
                                     int32_t* user_perm_out, int32_t* item_perm_out)
{
    int32_t g = num_groups;
    if (g > max_user + 1) g = max_user + 1;
    if (g > max_item + 1) g = max_item + 1;
    int32_t* up = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_user + 1));
    int32_t* ip = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_item + 1));
    for (int32_t u = 0; u <= max_user; u++) up[u] = u;
    for (int32_t i = 0; i <= max_item; i++) ip[i] = i;
    mo_shuffle_i32(r, up, max_user + 1);
    mo_shuffle_i32(r, ip, max_item + 1);
    mo_partition_blocks_given(users, items, n, up, ip, g, block_ptr, idx);
    for (int64_t b = 0; b < (int64_t)g * g; b++)
        mo_shuffle_i32(r, idx + block_ptr[b], block_ptr[b + 1] - block_ptr[b]);
    if (user_perm_out) memcpy(user_perm_out, up, sizeof(int32_t) * (size_t)(max_user + 1));
    if (item_perm_out) memcpy(item_perm_out, ip, sizeof(int32_t) * (size_t)(max_item + 1));
    free(up); free(ip);
    return g;
}

/* MultiCore.cs:79-92 */
int32_t mo_partition_indices(const int32_t* random_index, int64_t n, int32_t num_groups,
                             int64_t* group_ptr, int32_t* idx)
{
    int32_t g = num_groups;
    if ((int64_t)g > n) g = (int32_t)n;
    for (int32_t k = 0; k <= g; k++) group_ptr[k] = 0;
    for (int64_t t = 0; t < n; t++) group_ptr[(t % g) + 1]++;
    for (int32_t k = 0; k < g; k++) group_ptr[k + 1] += group_ptr[k];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(g > 0 ? g : 1));
    memcpy(cur, group_ptr, sizeof(int64_t) * (size_t)g);
    for (int64_t t = 0; t < n; t++) idx[cur[t % g]++] = random_index[t];
    free(cur);
    return g;
}

/* DataType/MatrixExtensions.cs:224-241: sequential fp32, multiply then add (never fused) */
float mo_row_scalar_product(const float* a, const float* b, int32_t k)
{
    float result = 0;
    for (int32_t c = 0; c < k; c++) result += a[c] * b[c];
    return result;
}

/* ------------------------------------------------------------------------------------------
 * MatrixFactorization / BiasedMatrixFactorization
 * ------------------------------------------------------------------------------------------ */
struct mo_model {
    int biased;
    mo_mf_params p;
    const int32_t* users; const int32_t* items; const float* values;
    int64_t n; int32_t max_user, max_item;
    int32_t* count_by_user; int32_t* count_by_item;
    float min_rating, max_rating, rating_range_size;
    float global_bias;
    float current_learnrate;
    double last_loss;
    float* U; float* V; float* bu; float* bi;
    int32_t* random_index;        /* lazily built, DataSet.cs:100-108 */
    /* MaxThreads > 1 */
    int32_t g;
    int64_t* block_ptr; int32_t* block_idx;
    int32_t n_lists; int64_t* list_ptr; int32_t* list_idx;
};

void mo_mf_params_default(mo_mf_params* p)
{
    memset(p, 0, sizeof(*p));
    p->num_factors = 10; p->learn_rate = 0.01f; p->decay = 1.0f; p->regularization = 0.015f;
    p->num_iter = 30; p->init_mean = 0; p->init_stddev = 0.1;
    p->bias_learn_rate = 1.0f; p->bias_reg = 0.01f; p->reg_u = 0.015f; p->reg_i = 0.015f;
    p->frequency_regularization = 0; p->loss = MO_LOSS_RMSE; p->max_threads = 1;
    p->bold_driver = 0; p->naive_parallelization = 0; p->omp_threads = 1;
}

mo_model* mo_model_create(int biased, const mo_mf_params* p,
                          const int32_t* users, const int32_t* items, const float* values, int64_t n,
                          int32_t max_user, int32_t max_item)
{
    mo_model* m = (mo_model*)calloc(1, sizeof(mo_model));
    m->biased = biased; m->p = *p;
    m->users = users; m->items = items; m->values = values; m->n = n;
    m->max_user = max_user; m->max_item = max_item;
    m->count_by_user = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_user + 1));
    m->count_by_item = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_item + 1));
    mo_count_by(users, n, max_user, m->count_by_user);
    mo_count_by(items, n, max_item, m->count_by_item);
    /* RatingPrediction/RatingPredictor.cs:39-49 */
    mo_scale(values, n, &m->min_rating, &m->max_rating);
    m->last_loss = -INFINITY;
    return m;
}

void mo_model_destroy(mo_model* m)
{
    if (!m) return;
    free(m->count_by_user); free(m->count_by_item);
    free(m->U); free(m->V); free(m->bu); free(m->bi);
    free(m->random_index); free(m->block_ptr); free(m->block_idx); free(m->list_ptr); free(m->list_idx);
    free(m);
}

float mo_model_learnrate(const mo_model* m)   { return m->current_learnrate; }
float mo_model_global_bias(const mo_model* m) { return m->global_bias; }
float* mo_model_user_factors(mo_model* m) { return m->U; }
float* mo_model_item_factors(mo_model* m) { return m->V; }
float* mo_model_user_bias(mo_model* m)    { return m->bu; }
float* mo_model_item_bias(mo_model* m)    { return m->bi; }
const int32_t* mo_model_random_index(mo_model* m) { return m->random_index; }

/* MatrixFactorization.cs:205-217, 251-259 ; BiasedMatrixFactorization.cs:313-325 */
float mo_model_predict(const mo_model* m, int32_t u, int32_t i)
{
    const int32_t k = m->p.num_factors;
    if (!m->biased) {
        if (u > m->max_user) return m->global_bias;
        if (i > m->max_item) return m->global_bias;
        float result = m->global_bias + mo_row_scalar_product(m->U + (int64_t)u * k, m->V + (int64_t)i * k, k);
        if (result > m->max_rating) return m->max_rating;
        if (result < m->min_rating) return m->min_rating;
        return result;
    }
    double score = m->global_bias;
    if (u <= m->max_user) score += m->bu[u];
    if (i <= m->max_item) score += m->bi[i];
    if (u <= m->max_user && i <= m->max_item)
        score += mo_row_scalar_product(m->U + (int64_t)u * k, m->V + (int64_t)i * k, k);
    return (float)(m->min_rating + (1 / (1 + exp(-score))) * m->rating_range_size);
}

void mo_model_predict_many(const mo_model* m, const int32_t* u, const int32_t* i, int64_t n, float* out)
{
    for (int64_t t = 0; t < n; t++) out[t] = mo_model_predict(m, u[t], i[t]);
}

/* Eval/Ratings.cs:96-162 */
static double mo_cbd(double actual, double prediction, double mn, double mx)
{
    prediction = (prediction - mn) / (mx - mn);
    actual = (actual - mn) / (mx - mn);
    if (prediction < 0.01) prediction = 0.01;
    if (prediction > 0.99) prediction = 0.99;
    return -(actual * log10(prediction) + (1 - actual) * log10(1 - prediction));
}

void mo_model_evaluate(const mo_model* m, const int32_t* u, const int32_t* i, const float* v, int64_t n, float* out4)
{
    double rmse = 0, mae = 0, cbd = 0;
    for (int64_t t = 0; t < n; t++) {
        float prediction = mo_model_predict(m, u[t], i[t]);
        float error = prediction - v[t];
        rmse += error * error;                 /* fp32 product widened */
        mae  += fabsf(error);
        cbd  += mo_cbd(v[t], prediction, m->min_rating, m->max_rating);
    }
    mae = mae / n; rmse = sqrt(rmse / n); cbd = cbd / n;
    out4[0] = (float)rmse; out4[1] = (float)mae;
    out4[2] = (float)mae / (m->max_rating - m->min_rating);
    out4[3] = (float)cbd;
}

/* BiasedMatrixFactorization.cs:496-552 (+ Eval/Measures/{RMSE,MAE,LogisticLoss}.cs) */
static double mo_norm_sq(const float* row, int32_t k)
{
    double sum = 0;
    for (int32_t f = 0; f < k; f++) { double v = row[f]; sum += pow(v, 2); }
    return pow(sqrt(sum), 2);
}

float mo_model_objective(const mo_model* m)
{
    const int32_t k = m->p.num_factors;
    double loss = 0;
    for (int64_t t = 0; t < m->n; t++) {
        float pred = mo_model_predict(m, m->users[t], m->items[t]);
        if (m->p.loss == MO_LOSS_MAE) loss += fabsf(pred - m->values[t]);
        else if (m->p.loss == MO_LOSS_RMSE) loss += pow(pred - m->values[t], 2);
        else {
            double prediction = pred;
            prediction = (prediction - m->min_rating) / m->rating_range_size;
            if (prediction < 0.0) prediction = 0.0;
            if (prediction > 1.0) prediction = 1.0;
            double actual = (m->values[t] - m->min_rating) / m->rating_range_size;
            loss -= actual * log(prediction);
            loss -= (1 - actual) * log(1 - prediction);
        }
    }
    double complexity = 0;
    if (m->p.frequency_regularization) {
        for (int32_t u = 0; u <= m->max_user; u++)
            if (m->count_by_user[u] > 0) {
                complexity += (m->p.reg_u / sqrt((double)m->count_by_user[u])) * mo_norm_sq(m->U + (int64_t)u * k, k);
                complexity += (m->p.reg_u / sqrt((double)m->count_by_user[u])) * m->p.bias_reg * pow(m->bu[u], 2);
            }
        for (int32_t i = 0; i <= m->max_item; i++)
            if (m->count_by_item[i] > 0) {
                complexity += (m->p.reg_i / sqrt((double)m->count_by_item[i])) * mo_norm_sq(m->V + (int64_t)i * k, k);
                complexity += (m->p.reg_i / sqrt((double)m->count_by_item[i])) * m->p.bias_reg * pow(m->bi[i], 2);
            }
    } else {
        for (int32_t u = 0; u <= m->max_user; u++) {
            complexity += m->count_by_user[u] * m->p.reg_u * mo_norm_sq(m->U + (int64_t)u * k, k);
            complexity += m->count_by_user[u] * m->p.reg_u * m->p.bias_reg * pow(m->bu[u], 2);
        }
        for (int32_t i = 0; i <= m->max_item; i++) {
            complexity += m->count_by_item[i] * m->p.reg_i * mo_norm_sq(m->V + (int64_t)i * k, k);
            complexity += m->count_by_item[i] * m->p.reg_i * m->p.bias_reg * pow(m->bi[i], 2);
        }
    }
    return (float)(loss + complexity);
}

/* BiasedMatrixFactorization.cs:225-244 ; MatrixFactorization.cs:129-132 */
static void mo_update_learnrate(mo_model* m)
{
    if (m->biased && m->p.bold_driver) {
        double loss = mo_model_objective(m);
        if (loss > m->last_loss) m->current_learnrate *= 0.5f;
        else if (loss < m->last_loss) m->current_learnrate *= 1.05f;
        m->last_loss = loss;
    } else {
        m->current_learnrate *= m->p.decay;
    }
}

/* MatrixFactorization.cs:166-196 */
static void mo_mf_iterate_list(mo_model* m, const int32_t* idx, int64_t n_idx, int update_user, int update_item)
{
    const int32_t k = m->p.num_factors;
    const float reg = m->p.regularization;
    for (int64_t t = 0; t < n_idx; t++) {
        int32_t index = idx[t];
        int32_t u = m->users[index], i = m->items[index];
        float* pu = m->U + (int64_t)u * k;
        float* qi = m->V + (int64_t)i * k;
        float err = m->values[index] - (m->global_bias + mo_row_scalar_product(pu, qi, k));
        for (int32_t f = 0; f < k; f++) {
            float u_f = pu[f], i_f = qi[f];
            if (update_user) {
                double delta_u = err * i_f - reg * u_f;          /* fp32 expression widened */
                pu[f] += (float)(m->current_learnrate * delta_u);
            }
            if (update_item) {
                double delta_i = err * u_f - reg * i_f;
                qi[f] += (float)(m->current_learnrate * delta_i);
            }
        }
    }
    mo_update_learnrate(m);   /* MatrixFactorization.cs:195 */
}

/* BiasedMatrixFactorization.cs:247-261 */
static inline float mo_gradient_common(int loss, double sig_score, double err, float range)
{
    switch (loss) {
    case MO_LOSS_MAE: {
        int s = (err > 0) - (err < 0);
        return (float)(s * sig_score * (1 - sig_score) * range);
    }
    case MO_LOSS_LOGISTIC: return (float)err;
    default: return (float)(err * sig_score * (1 - sig_score) * range);
    }
}

/* BiasedMatrixFactorization.cs:264-310 -- THE hot loop */
static void mo_bmf_iterate_list(mo_model* m, const int32_t* idx, int64_t n_idx, int update_user, int update_item)
{
    const int32_t k = m->p.num_factors;
    const float lr = m->current_learnrate;
    for (int64_t t = 0; t < n_idx; t++) {
        int32_t index = idx[t];
        int32_t u = m->users[index], i = m->items[index];
        float* pu = m->U + (int64_t)u * k;
        float* qi = m->V + (int64_t)i * k;

        double score = m->global_bias + m->bu[u] + m->bi[i] + mo_row_scalar_product(pu, qi, k); /* fp32 sum */
        double sig_score = 1 / (1 + exp(-score));
        double prediction = m->min_rating + sig_score * m->rating_range_size;
        double err = m->values[index] - prediction;
        float gradient_common = mo_gradient_common(m->p.loss, sig_score, err, m->rating_range_size);

        float user_reg_weight = m->p.frequency_regularization ? (float)(m->p.reg_u / sqrt((double)m->count_by_user[u])) : m->p.reg_u;
        float item_reg_weight = m->p.frequency_regularization ? (float)(m->p.reg_i / sqrt((double)m->count_by_item[i])) : m->p.reg_i;

        if (update_user)
            m->bu[u] += m->p.bias_learn_rate * lr * (gradient_common - m->p.bias_reg * user_reg_weight * m->bu[u]);
        if (update_item)
            m->bi[i] += m->p.bias_learn_rate * lr * (gradient_common - m->p.bias_reg * item_reg_weight * m->bi[i]);

        for (int32_t f = 0; f < k; f++) {
            double u_f = pu[f], i_f = qi[f];
            if (update_user) {
                double delta_u = gradient_common * i_f - user_reg_weight * u_f;
                pu[f] += (float)(lr * delta_u);
            }
            if (update_item) {
                double delta_i = gradient_common * u_f - item_reg_weight * i_f;
                qi[f] += (float)(lr * delta_i);
            }
        }
    }
}

void mo_model_iterate_indices(mo_model* m, const int32_t* idx, int64_t n_idx, int update_user, int update_item)
{
    if (m->biased) mo_bmf_iterate_list(m, idx, n_idx, update_user, update_item);
    else mo_mf_iterate_list(m, idx, n_idx, update_user, update_item);
}

/* FoldIn: MatrixFactorization.cs:323-347 (plain; out has k entries) and BiasedMatrixFactorization.cs:445-492
 * (biased; out[0] = user bias, out[1..k] = factors: FOLD_IN_BIAS_INDEX / FOLD_IN_FACTORS_START, :80-82).
 * `items`/`values` are rated_items AFTER rated_items.Shuffle(), `init` is the vector InitNormal drew -- both draws belong
 * to the caller's RNG stream (vector first, then the shuffle). The model is not modified. */
void mo_model_fold_in(const mo_model* m, const int32_t* items, const float* values, int64_t n, const float* init, float* out)
{
    const int32_t k = m->p.num_factors;
    if (!m->biased) {
        float* user_vector = out;
        memcpy(user_vector, init, sizeof(float) * (size_t)k);
        const float reg = m->p.regularization;
        double lr = m->p.learn_rate;
        for (int32_t it = 0; it < m->p.num_iter; it++) {
            for (int64_t t = 0; t < n; t++) {
                const float* qi = m->V + (int64_t)items[t] * k;
                float err = values[t] - (m->global_bias + mo_row_scalar_product(qi, user_vector, k));  /* Predict(v, i, false) */
                for (int32_t f = 0; f < k; f++) {
                    float u_f = user_vector[f], i_f = qi[f];
                    double delta_u = err * i_f - reg * u_f;       /* fp32 expression widened */
                    user_vector[f] += (float)(lr * delta_u);
                }
            }
            lr *= m->p.decay;
        }
        return;
    }
    float user_bias = 0;
    float* factors = out + 1;
    memcpy(factors, init, sizeof(float) * (size_t)k);
    const float reg_weight = m->p.frequency_regularization ? (float)(m->p.reg_u / sqrt((double)n)) : m->p.reg_u;
    const float lrate = m->p.learn_rate;                          /* LearnRate, not current_learnrate (:466,:477) */
    for (int32_t it = 0; it < m->p.num_iter; it++)
        for (int64_t t = 0; t < n; t++) {
            const int32_t item_id = items[t];
            const float* qi = m->V + (int64_t)item_id * k;
            double score = m->global_bias + user_bias + m->bi[item_id] + mo_row_scalar_product(qi, factors, k);
            double sig_score = 1 / (1 + exp(-score));
            double prediction = m->min_rating + sig_score * m->rating_range_size;
            double err = values[t] - prediction;
            float gradient_common = mo_gradient_common(m->p.loss, sig_score, err, m->rating_range_size);
            user_bias += m->p.bias_learn_rate * lrate * (gradient_common - m->p.bias_reg * reg_weight * user_bias);
            for (int32_t f = 0; f < k; f++) {
                float u_f = factors[f], i_f = qi[f];
                double delta_u = gradient_common * i_f - reg_weight * u_f;   /* fp32 expression widened (:474-476) */
                factors[f] += (float)(lrate * delta_u);
            }
        }
    out[0] = user_bias;
}

/* Predict(float[] user_vector, int item_id): MatrixFactorization.cs:223-241 (bounded), BiasedMatrixFactorization.cs:328-336 */
float mo_model_predict_vector(const mo_model* m, const float* user_vector, int32_t item_id)
{
    const int32_t k = m->p.num_factors;
    if (!m->biased) {
        float result = m->global_bias + mo_row_scalar_product(m->V + (int64_t)item_id * k, user_vector, k);
        if (result > m->max_rating) return m->max_rating;
        if (result < m->min_rating) return m->min_rating;
        return result;
    }
    double score = m->global_bias + user_vector[0];
    if (item_id <= m->max_item)
        score += m->bi[item_id] + mo_row_scalar_product(m->V + (int64_t)item_id * k, user_vector + 1, k);
    return (float)(m->min_rating + 1 / (1 + exp(-score)) * m->rating_range_size);
}

/* MatrixFactorization.cs:99-116 ; BiasedMatrixFactorization.cs:161-190 */
void mo_model_init(mo_model* m, mo_rng* r)
{
    const int32_t k = m->p.num_factors;
    int64_t nu = (int64_t)m->max_user + 1, ni = (int64_t)m->max_item + 1;
    free(m->U); free(m->V); free(m->bu); free(m->bi);
    m->U = (float*)malloc(sizeof(float) * (size_t)(nu * k));
    m->V = (float*)malloc(sizeof(float) * (size_t)(ni * k));
    mo_init_normal(r, m->U, nu * k, m->p.init_mean, m->p.init_stddev);
    mo_init_normal(r, m->V, ni * k, m->p.init_mean, m->p.init_stddev);
    for (int64_t u = 0; u < nu; u++)
        if (m->count_by_user[u] == 0) for (int32_t f = 0; f < k; f++) m->U[u * k + f] = 0;
    for (int64_t i = 0; i < ni; i++)
        if (m->count_by_item[i] == 0) for (int32_t f = 0; f < k; f++) m->V[i * k + f] = 0;
    m->current_learnrate = m->p.learn_rate;
    m->bu = (float*)calloc((size_t)nu, sizeof(float));
    m->bi = (float*)calloc((size_t)ni, sizeof(float));
    m->rating_range_size = m->max_rating - m->min_rating;
    if (m->biased) {
        /* the reference computes last_loss inside InitModel with global_bias still unset (0) and
         * rating_range_size unset (0): BiasedMatrixFactorization.cs:161-170 precedes :186-190 */
        if (m->p.bold_driver) {
            float keep = m->rating_range_size;
            m->rating_range_size = 0; m->global_bias = 0;
            m->last_loss = mo_model_objective(m);
            m->rating_range_size = keep;
        }
        if (m->p.max_threads > 1) {
            if (m->p.naive_parallelization) {
                if (!m->random_index) {
                    m->random_index = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->n);
                    mo_random_index(r, m->random_index, m->n);
                }
                free(m->list_ptr); free(m->list_idx);
                m->list_ptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m->p.max_threads + 1));
                m->list_idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->n);
                m->n_lists = mo_partition_indices(m->random_index, m->n, m->p.max_threads, m->list_ptr, m->list_idx);
            } else {
                free(m->block_ptr); free(m->block_idx);
                int64_t gmax = m->p.max_threads;
                m->block_ptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(gmax * gmax + 1));
                m->block_idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->n);
                m->g = mo_partition_users_and_items(r, m->users, m->items, m->n, m->max_user, m->max_item,
                                                    m->p.max_threads, m->block_ptr, m->block_idx, NULL, NULL);
            }
        }
        double avg = (mo_average(m->values, m->n) - m->min_rating) / m->rating_range_size; /* fp32 expr widened */
        m->global_bias = (float)log(avg / (1 - avg));
    } else {
        m->global_bias = mo_average(m->values, m->n);     /* MatrixFactorization.cs:124 */
    }
}

static const int32_t* mo_get_random_index(mo_model* m, mo_rng* r)
{
    if (!m->random_index) {
        m->random_index = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->n);
        mo_random_index(r, m->random_index, m->n);
    }
    return m->random_index;
}

/* BiasedMatrixFactorization.cs:197-222 ; MatrixFactorization.cs:135-138 */
void mo_model_iterate(mo_model* m, mo_rng* r)
{
    if (!m->biased) {
        mo_mf_iterate_list(m, mo_get_random_index(m, r), m->n, 1, 1);
        return;
    }
    if (m->p.max_threads > 1) {
        if (m->p.naive_parallelization) {
            /* Parallel.For over thread_lists: racy in the reference; the oracle runs the lists one
             * after another (one of the legal interleavings). */
            for (int32_t l = 0; l < m->n_lists; l++)
                mo_bmf_iterate_list(m, m->list_idx + m->list_ptr[l], m->list_ptr[l + 1] - m->list_ptr[l], 1, 1);
        } else {
            int32_t g = m->g;
            int32_t* seq = (int32_t*)malloc(sizeof(int32_t) * (size_t)g);
            for (int32_t s = 0; s < g; s++) seq[s] = s;
            mo_shuffle_i32(r, seq, g);
            for (int32_t s = 0; s < g; s++) {
                int32_t i = seq[s];
                /* blocks of one sub-epoch touch disjoint users and items: any execution order,
                 * sequential or parallel, gives bit-identical results */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(m->p.omp_threads > 1 ? m->p.omp_threads : 1)
#endif
                for (int32_t j = 0; j < g; j++) {
                    int64_t b = (int64_t)j * g + ((i + j) % g);
                    mo_bmf_iterate_list(m, m->block_idx + m->block_ptr[b], m->block_ptr[b + 1] - m->block_ptr[b], 1, 1);
                }
            }
            free(seq);
        }
        mo_update_learnrate(m);   /* :216 */
    } else {
        mo_bmf_iterate_list(m, mo_get_random_index(m, r), m->n, 1, 1);
    }
    mo_update_learnrate(m);       /* :221 */
}

void mo_model_train(mo_model* m, mo_rng* r)
{
    mo_model_init(m, r);
    for (int32_t it = 0; it < m->p.num_iter; it++) mo_model_iterate(m, r);
}

/* Replay of the GPU stratum schedule (mini-batch per item run) in the reference's arithmetic style
 * but fp32 (the product kernel is fp32 throughout); used only to check the kernel itself. */
void mo_bmf_replay_runs(mo_model* m, const int32_t* run_item, const int64_t* run_ptr, int64_t n_runs,
                        const int32_t* ent_user, const float* ent_value, int32_t batch)
{
    const int32_t k = m->p.num_factors;
    const float lr = m->current_learnrate;
    float* gq = (float*)malloc(sizeof(float) * (size_t)k);
    float* g = (float*)malloc(sizeof(float) * (size_t)batch);
    for (int64_t rr = 0; rr < n_runs; rr++) {
        int32_t i = run_item[rr];
        float* qi = m->V + (int64_t)i * k;
        float regi = m->p.frequency_regularization ? (float)(m->p.reg_i / sqrt((double)m->count_by_item[i])) : m->p.reg_i;
        for (int64_t s = run_ptr[rr]; s < run_ptr[rr + 1]; s += batch) {
            int32_t nb = (int32_t)((run_ptr[rr + 1] - s) < batch ? (run_ptr[rr + 1] - s) : batch);
            float gsum = 0;
            for (int32_t f = 0; f < k; f++) gq[f] = 0;
            for (int32_t b = 0; b < nb; b++) {
                int32_t u = ent_user[s + b];
                float* pu = m->U + (int64_t)u * k;
                float dot = 0;
                for (int32_t f = 0; f < k; f++) dot += pu[f] * qi[f];
                float score = m->global_bias + m->bu[u] + m->bi[i] + dot;
                float sig = 1.0f / (1.0f + expf(-score));
                float err = ent_value[s + b] - (m->min_rating + sig * m->rating_range_size);
                float gc;
                if (m->p.loss == MO_LOSS_MAE) gc = ((err > 0) - (err < 0)) * sig * (1 - sig) * m->rating_range_size;
                else if (m->p.loss == MO_LOSS_LOGISTIC) gc = err;
                else gc = err * sig * (1 - sig) * m->rating_range_size;
                g[b] = gc; gsum += gc;
                float regu = m->p.frequency_regularization ? (float)(m->p.reg_u / sqrt((double)m->count_by_user[u])) : m->p.reg_u;
                m->bu[u] += m->p.bias_learn_rate * lr * (gc - m->p.bias_reg * regu * m->bu[u]);
                for (int32_t f = 0; f < k; f++) {
                    float pf = pu[f];
                    gq[f] += gc * pf;
                    pu[f] = pf + lr * (gc * qi[f] - regu * pf);
                }
            }
            m->bi[i] += m->p.bias_learn_rate * lr * (gsum - nb * m->p.bias_reg * regi * m->bi[i]);
            for (int32_t f = 0; f < k; f++) qi[f] += lr * (gq[f] - nb * regi * qi[f]);
        }
    }
    free(gq); free(g);
}

/* ------------------------------------------------------------------------------------------
 * WRMF (ItemRecommendation/WRMF.cs:68-156)
 * ------------------------------------------------------------------------------------------ */
/* WRMF.cs:94-108: fp32 products accumulated in double, row index ascending per (f1,f2) */
void mo_wrmf_gram(const float* H, int32_t n_rows, int32_t k, double* HH)
{
    for (int32_t a = 0; a < k * k; a++) HH[a] = 0;
    for (int32_t i = 0; i < n_rows; i++) {
        const float* h = H + (int64_t)i * k;
        for (int32_t f1 = 0; f1 < k; f1++)
            for (int32_t f2 = f1; f2 < k; f2++)
                HH[f1 * k + f2] += h[f1] * h[f2];
    }
    for (int32_t f1 = 0; f1 < k; f1++)
        for (int32_t f2 = f1 + 1; f2 < k; f2++) HH[f2 * k + f1] = HH[f1 * k + f2];
}

/* Solve the SPD system m x = b in double (the reference forms m^-1 with MathNet's LU-based
 * DenseMatrix.Inverse and multiplies, WRMF.cs:137-155: identical up to double rounding). */
static int mo_chol_solve(double* A, double* b, int32_t k)
{
    for (int32_t j = 0; j < k; j++) {
        double d = A[j * k + j];
        for (int32_t t = 0; t < j; t++) d -= A[j * k + t] * A[j * k + t];
        if (d <= 0) return -1;
        d = sqrt(d);
        A[j * k + j] = d;
        for (int32_t i = j + 1; i < k; i++) {
            double s = A[i * k + j];
            for (int32_t t = 0; t < j; t++) s -= A[i * k + t] * A[j * k + t];
            A[i * k + j] = s / d;
        }
    }
    for (int32_t i = 0; i < k; i++) {
        double s = b[i];
        for (int32_t t = 0; t < i; t++) s -= A[i * k + t] * b[t];
        b[i] = s / A[i * k + i];
    }
    for (int32_t i = k - 1; i >= 0; i--) {
        double s = b[i];
        for (int32_t t = i + 1; t < k; t++) s -= A[t * k + i] * b[t];
        b[i] = s / A[i * k + i];
    }
    return 0;
}

/* WRMF.cs:79-92, 110-156 */
void mo_wrmf_optimize(const int64_t* row_ptr, const int32_t* cols, int32_t n_rows,
                      float* W, const float* H, int32_t n_h_rows, int32_t k,
                      double alpha, double regularization, int omp_threads)
{
    double* HH = (double*)malloc(sizeof(double) * (size_t)(k * k));
    mo_wrmf_gram(H, n_h_rows, k, HH);
#ifdef _OPENMP
#pragma omp parallel num_threads(omp_threads > 1 ? omp_threads : 1)
#endif
    {
        double* A = (double*)malloc(sizeof(double) * (size_t)(k * k));
        double* b = (double*)malloc(sizeof(double) * (size_t)k);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int32_t u = 0; u < n_rows; u++) {
            for (int32_t a = 0; a < k * k; a++) A[a] = 0;
            for (int32_t f = 0; f < k; f++) b[f] = 0;
            for (int64_t e = row_ptr[u]; e < row_ptr[u + 1]; e++) {
                const float* h = H + (int64_t)cols[e] * k;
                for (int32_t f1 = 0; f1 < k; f1++) {
                    for (int32_t f2 = f1; f2 < k; f2++) A[f1 * k + f2] += h[f1] * h[f2];  /* fp32 product */
                    b[f1] += h[f1];
                }
            }
            for (int32_t f1 = 0; f1 < k; f1++) {
                for (int32_t f2 = f1; f2 < k; f2++) {
                    double d = HH[f1 * k + f2] + A[f1 * k + f2] * alpha;
                    if (f1 == f2) d += regularization;
                    A[f1 * k + f2] = d; A[f2 * k + f1] = d;
                }
                b[f1] = b[f1] * (1 + alpha);
            }
            mo_chol_solve(A, b, k);
            for (int32_t f = 0; f < k; f++) W[(int64_t)u * k + f] = (float)b[f];
        }
        free(A); free(b);
    }
    free(HH);
}

/* Data/PosOnlyFeedback.cs:68-83 + DataType/SparseBooleanMatrix.cs:37-91 (rows are HashSets) */
int64_t mo_feedback_csr(const int32_t* rows, const int32_t* cols, int64_t n, int32_t max_row,
                        int64_t* row_ptr, int32_t* out_cols)
{
    int64_t nr = (int64_t)max_row + 1;
    int64_t* tmp_ptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nr + 1));
    int32_t* tmp_idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    mo_build_index(rows, n, max_row, tmp_ptr, tmp_idx);
    int64_t w = 0;
    for (int64_t r = 0; r < nr; r++) {
        row_ptr[r] = w;
        int64_t start = w;
        for (int64_t e = tmp_ptr[r]; e < tmp_ptr[r + 1]; e++) {
            int32_t c = cols[tmp_idx[e]];
            int dup = 0;
            for (int64_t q = start; q < w; q++) if (out_cols[q] == c) { dup = 1; break; }
            if (!dup) out_cols[w++] = c;
        }
    }
    row_ptr[nr] = w;
    free(tmp_ptr); free(tmp_idx);
    return w;
}

/* ------------------------------------------------------------------------------------------
 * Recommender.Recommend (Recommender.cs:52-103)
 * ------------------------------------------------------------------------------------------ */
typedef struct { float score; int64_t pos; int32_t item; } mo_scored;

static int mo_scored_cmp(const void* a, const void* b)
{
    const mo_scored* x = (const mo_scored*)a; const mo_scored* y = (const mo_scored*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->pos > y->pos) - (x->pos < y->pos);   /* stable: OrderByDescending keeps list order */
}

static int mo_in_sorted(const int32_t* a, int64_t n, int32_t v)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo < n && a[lo] == v;
}

static int mo_i32_cmp(const void* a, const void* b)
{
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

typedef float (*mo_score_fn)(const void* ctx, int32_t user, int32_t item);

static int64_t mo_recommend_generic(mo_score_fn fn, const void* ctx, int32_t user, int32_t n,
                                    const int32_t* candidates, int64_t n_cand,
                                    const int32_t* ignore, int64_t n_ignore,
                                    int32_t* out_items, float* out_scores)
{
    int32_t* ign = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_ignore > 0 ? n_ignore : 1));
    if (n_ignore > 0) memcpy(ign, ignore, sizeof(int32_t) * (size_t)n_ignore);
    qsort(ign, (size_t)n_ignore, sizeof(int32_t), mo_i32_cmp);
    mo_scored* sc = (mo_scored*)malloc(sizeof(mo_scored) * (size_t)(n_cand > 0 ? n_cand : 1));
    int64_t cnt = 0;
    for (int64_t c = 0; c < n_cand; c++) {
        int32_t item = candidates[c];
        if (mo_in_sorted(ign, n_ignore, item)) continue;
        float s = fn(ctx, user, item);
        if (s > -FLT_MAX) { sc[cnt].score = s; sc[cnt].pos = c; sc[cnt].item = item; cnt++; }
    }
    qsort(sc, (size_t)cnt, sizeof(mo_scored), mo_scored_cmp);
    int64_t out = (n < 0 || cnt < n) ? cnt : n;
    for (int64_t t = 0; t < out; t++) { out_items[t] = sc[t].item; out_scores[t] = sc[t].score; }
    free(sc); free(ign);
    return out;
}

typedef struct { const float* U; const float* V; int32_t k, n_users, n_items; } mo_mf_ctx;

/* ItemRecommendation/MF.cs:151-157 */
static float mo_mf_score(const void* c, int32_t user, int32_t item)
{
    const mo_mf_ctx* x = (const mo_mf_ctx*)c;
    if (user >= x->n_users || item >= x->n_items) return -FLT_MAX;
    return mo_row_scalar_product(x->U + (int64_t)user * x->k, x->V + (int64_t)item * x->k, x->k);
}

int64_t mo_recommend_mf(const float* U, const float* V, int32_t k, int32_t n_items_model, int32_t n_users_model,
                        int32_t user, int32_t n,
                        const int32_t* candidates, int64_t n_cand,
                        const int32_t* ignore, int64_t n_ignore,
                        int32_t* out_items, float* out_scores)
{
    mo_mf_ctx c = { U, V, k, n_users_model, n_items_model };
    return mo_recommend_generic(mo_mf_score, &c, user, n, candidates, n_cand, ignore, n_ignore, out_items, out_scores);
}

static float mo_model_score(const void* c, int32_t user, int32_t item)
{
    return mo_model_predict((const mo_model*)c, user, item);
}

int64_t mo_recommend_model(const mo_model* m, int32_t user, int32_t n,
                           const int32_t* candidates, int64_t n_cand,
                           const int32_t* ignore, int64_t n_ignore,
                           int32_t* out_items, float* out_scores)
{
    return mo_recommend_generic(mo_model_score, m, user, n, candidates, n_cand, ignore, n_ignore, out_items, out_scores);
}
