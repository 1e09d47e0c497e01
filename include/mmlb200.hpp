// mmlb200.hpp -- C++17 host layer over the C ABI of libmmlb200.so (include/mmlb200.h), header only.
//
// The reference is compiled code (C#); this image has no .NET / Mono toolchain, so next to the C# sources in csharp/ the host
// side is given in C++ as well: the same three recommender classes with the reference's surface -- property names and
// defaults, Train() / Iterate() / Predict() / Recommend() / SaveModel() / LoadModel() / ToString(), the same RNG draw
// order, the same text model layout, errors as exceptions carrying mml_last_error(). Nothing here computes on the CPU:
// every call that touches ratings, factors or scores goes to the library; without a CUDA device Context() throws.
//
//   MyMediaLite.Random / System.Random              Random.cs:23-64 (+ BCL)                -> mymedialite::Random
//   Utils.Shuffle                                   Utils.cs:52-64                         -> Random::Shuffle
//   MatrixExtensions.InitNormal (MathNet Normal)    DataType/MatrixExtensions.cs:62-69     -> Random::InitNormal
//   IRatings / StaticRatings                        Data/StaticRatings.cs:47-84            -> Ratings
//   IPosOnlyFeedback                                Data/PosOnlyFeedback.cs:35-83          -> PosOnlyFeedback
//   RatingPrediction.MatrixFactorization            MatrixFactorization.cs:35-418          -> MatrixFactorization
//   RatingPrediction.BiasedMatrixFactorization      BiasedMatrixFactorization.cs:61-563    -> BiasedMatrixFactorization
//   ItemRecommendation.WRMF                         ItemRecommendation/WRMF.cs, MF.cs      -> WRMF
//   IO.Model / MatrixExtensions / VectorExtensions  IO/Model.cs:85-114, IO/*.cs            -> modelio::*
//   IO.StaticRatingData / ItemData                  IO/StaticRatingData.cs:36-117, ...     -> StaticRatingData / ItemData
#ifndef MMLB200_HPP
#define MMLB200_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mmlb200.h"

namespace mymedialite {

// ---- errors: non-zero status -> exception (the C# wrapper throws InvalidOperationException / FormatException) ---------
struct MmlError : std::runtime_error {
    int32_t status;
    MmlError(int32_t st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};
struct FormatException : MmlError { using MmlError::MmlError; };

inline void Check(int32_t status)
{
    if (status == MML_OK) return;
    const char* m = mml_last_error();
    if (status == MML_ERR_FORMAT) throw FormatException(status, m ? m : "");
    throw MmlError(status, m ? m : "");
}

// ---- System.Random (Knuth's subtractive generator, MSEED 161803398) + the draws the recommenders make -------------------
class Random {
    int32_t sa_[56];
    int inext_ = 0, inextp_ = 21;
    static constexpr int32_t MBIG = 2147483647, MSEED = 161803398;

    int32_t Sample()
    {
        if (++inext_ >= 56) inext_ = 1;
        if (++inextp_ >= 56) inextp_ = 1;
        int32_t r = sa_[inext_] - sa_[inextp_];
        if (r == MBIG) r--;
        if (r < 0) r += MBIG;
        sa_[inext_] = r;
        return r;
    }

public:
    explicit Random(int32_t seed)
    {
        const int32_t sub = seed == std::numeric_limits<int32_t>::min() ? MBIG : std::abs(seed);
        int32_t mj = MSEED - sub, mk = 1;
        for (int i = 0; i < 56; i++) sa_[i] = 0;
        sa_[55] = mj;
        for (int i = 1; i < 55; i++) {
            const int ii = (21 * i) % 55;
            sa_[ii] = mk;
            mk = mj - mk;
            if (mk < 0) mk += MBIG;
            mj = sa_[ii];
        }
        for (int k = 1; k < 5; k++)
            for (int i = 1; i < 56; i++) {
                sa_[i] = (int32_t)((uint32_t)sa_[i] - (uint32_t)sa_[1 + (i + 30) % 55]);   // wraps like C# unchecked int
                if (sa_[i] < 0) sa_[i] += MBIG;
            }
    }

    int32_t Next() { return Sample(); }
    double NextDouble() { return Sample() * (1.0 / MBIG); }
    int32_t Next(int32_t max_value) { return (int32_t)(NextDouble() * max_value); }

    // Utils.Shuffle (Utils.cs:52-64): for i = n-1 .. 0: r = Next(i + 1); swap(a[i], a[r]) -- the i = 0 step draws too
    template <typename T>
    void Shuffle(std::vector<T>& a)
    {
        for (int64_t i = (int64_t)a.size() - 1; i >= 0; i--) {
            const int32_t r = Next((int32_t)i + 1);
            std::swap(a[(size_t)i], a[(size_t)r]);
        }
    }
    // the swap targets alone (applied on the device by mml_shuffle_apply)
    std::vector<int32_t> ShuffleTargets(int64_t n)
    {
        std::vector<int32_t> H((size_t)n);
        for (int64_t i = n - 1; i >= 0; i--) H[(size_t)i] = Next((int32_t)i + 1);
        return H;
    }
    // MathNet.Numerics Normal.Sample: polar Box-Muller, two NextDouble per trial, the first variate is returned
    double Normal(double mean, double stddev)
    {
        double v1, r;
        do {
            v1 = 2.0 * NextDouble() - 1.0;
            const double v2 = 2.0 * NextDouble() - 1.0;
            r = v1 * v1 + v2 * v2;
        } while (r >= 1.0 || r == 0.0);
        return mean + stddev * v1 * std::sqrt(-2.0 * std::log(r) / r);
    }
    // MatrixExtensions.InitNormal (DataType/MatrixExtensions.cs:62-69): row-major, cast to float
    std::vector<float> InitNormal(int64_t count, double mean, double stddev)
    {
        std::vector<float> out((size_t)count);
        for (auto& x : out) x = (float)Normal(mean, stddev);
        return out;
    }

    // MyMediaLite.Random (Random.cs:23-64): one instance per thread, re-created by the Seed setter
    static std::unique_ptr<Random>& Slot() { static thread_local std::unique_ptr<Random> inst; return inst; }
    static void Seed(int32_t seed) { Slot().reset(new Random(seed)); }
    static Random& GetInstance()
    {
        if (!Slot()) Slot().reset(new Random((int32_t)(std::time(nullptr) & 0x7FFFFFFF)));
        return *Slot();
    }
};

// ---- library context: one per process and GPU count (mml_ctx); throws without a CUDA device (there is no CPU path) -----
class Context {
    mml_ctx* h_ = nullptr;
public:
    // GPUs 0 .. n_gpus - 1 driven from this process (mml_ctx_create with n_gpus > 1 = the NumGpus property of the classes)
    explicit Context(int n_gpus = 1) { Check(mml_ctx_create((int32_t)std::max(n_gpus, 1), nullptr, &h_)); }
    ~Context() { if (h_) mml_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    mml_ctx* get() const { return h_; }
    static Context& Default() { return ForGpus(1); }
    static Context& ForGpus(uint32_t n_gpus)
    {
        static std::map<uint32_t, std::unique_ptr<Context>> all;
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        auto& c = all[std::max<uint32_t>(n_gpus, 1)];
        if (!c) c.reset(new Context((int)std::max<uint32_t>(n_gpus, 1)));
        return *c;
    }
};

// ---- the one engine knob (process-wide; not a recommender option: the option set stays the reference's + NumGpus) -------
//   Order::Auto (default)  Iterate() runs the parallel epoch kernel (the DSGD block schedule) whatever MaxThreads says --
//                          MaxThreads keeps its other meaning, UpdateLearnRate twice per epoch when > 1 -- except on data
//                          sets below SerialBelow ratings, where the exact single-threaded order costs nothing;
//   Order::Reference       MaxThreads = 1 walks RandomIndex in the reference's order on one warp (parity runs);
//   Order::Parallel        always the parallel kernel.
//   DeviceInit             InitModel draws on the device (counter-based) instead of MyMediaLite.Random on the host.
// Environment: MMLB200_ORDER = auto | reference | parallel, MMLB200_INIT = host | device.
struct Engine {
    enum class Order { Auto, Reference, Parallel };
    static constexpr int64_t SerialBelow = 20000;
    static Order& order()
    {
        static Order o = [] {
            const char* e = std::getenv("MMLB200_ORDER");
            const std::string v = e ? e : "auto";
            return v == "reference" ? Order::Reference : (v == "parallel" ? Order::Parallel : Order::Auto);
        }();
        return o;
    }
    static bool& device_init()
    {
        static bool d = [] { const char* e = std::getenv("MMLB200_INIT"); return e && std::string(e) == "device"; }();
        return d;
    }
};

// ---- data sets -------------------------------------------------------------------------------------------------------------
struct Ratings {                                     // IRatings: COO triples + MaxUserID / MaxItemID
    std::vector<int32_t> Users, Items;
    std::vector<float> Values;
    int32_t MaxUserID = -1, MaxItemID = -1;
    int64_t Count() const { return (int64_t)Users.size(); }
    void Add(int32_t user, int32_t item, float value)
    {
        Users.push_back(user); Items.push_back(item); Values.push_back(value);
        MaxUserID = std::max(MaxUserID, user); MaxItemID = std::max(MaxItemID, item);
        random_index_.clear();
    }
    // DataSet.RandomIndex (Data/DataSet.cs:100-109, 193-202): shuffled once, rebuilt when Count changes; the swap targets come
    // from MyMediaLite.Random, the permutation is applied on the device
    const std::vector<int32_t>& RandomIndex()
    {
        if ((int64_t)random_index_.size() != Count()) {
            random_index_.resize((size_t)Count());
            for (int64_t t = 0; t < Count(); t++) random_index_[(size_t)t] = (int32_t)t;
            if (Count() > 0) {
                const std::vector<int32_t> H = Random::GetInstance().ShuffleTargets(Count());
                Check(mml_shuffle_apply(Context::Default().get(), random_index_.data(), H.data(), Count()));
            }
        }
        return random_index_;
    }
private:
    std::vector<int32_t> random_index_;
};

struct PosOnlyFeedback {                              // IPosOnlyFeedback: (user, item) events
    std::vector<int32_t> Users, Items;
    int32_t MaxUserID = -1, MaxItemID = -1;
    int64_t Count() const { return (int64_t)Users.size(); }
    void Add(int32_t user, int32_t item)
    {
        Users.push_back(user); Items.push_back(item);
        MaxUserID = std::max(MaxUserID, user); MaxItemID = std::max(MaxItemID, item);
    }
};

// ---- readers (IO/StaticRatingData.cs, IO/ItemData.cs) on the native parallel parser ---------------------------------------
namespace detail {
struct IngestHandle {
    mml_ingest* h = nullptr;
    ~IngestHandle() { if (h) mml_ingest_destroy(h); }
};
}
struct StaticRatingData {
    static Ratings Read(const std::string& filename, bool ignore_first_line = false, int n_threads = 0)
    {
        detail::IngestHandle g;
        Check(mml_ingest_file(filename.c_str(), MML_FILE_RATINGS, MML_MAP_IDENTITY, MML_MAP_IDENTITY, ignore_first_line ? 1 : 0,
                              n_threads, nullptr, &g.h));
        Ratings r;
        int64_t n = 0;
        Check(mml_ingest_info(g.h, &n, &r.MaxUserID, &r.MaxItemID, nullptr, nullptr, nullptr));
        r.Users.resize((size_t)n); r.Items.resize((size_t)n); r.Values.resize((size_t)n);
        Check(mml_ingest_copy(g.h, r.Users.data(), r.Items.data(), r.Values.data()));
        return r;
    }
};
struct ItemData {
    static PosOnlyFeedback Read(const std::string& filename, bool ignore_first_line = false, int n_threads = 0)
    {
        detail::IngestHandle g;
        Check(mml_ingest_file(filename.c_str(), MML_FILE_FEEDBACK, MML_MAP_IDENTITY, MML_MAP_IDENTITY, ignore_first_line ? 1 : 0,
                              n_threads, nullptr, &g.h));
        PosOnlyFeedback f;
        int64_t n = 0;
        Check(mml_ingest_info(g.h, &n, &f.MaxUserID, &f.MaxItemID, nullptr, nullptr, nullptr));
        f.Users.resize((size_t)n); f.Items.resize((size_t)n);
        Check(mml_ingest_copy(g.h, f.Users.data(), f.Items.data(), nullptr));
        return f;
    }
};

// ---- text model format (IO/Model.cs:85-114, IO/MatrixExtensions.cs:31-89, IO/VectorExtensions.cs:40-60) --------------
namespace modelio {
// float.ToString(CultureInfo.InvariantCulture) on the .NET Framework / Mono: 7 significant digits ("G7"), "E+XX" exponents
inline std::string Fmt(float x)
{
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
    if (x == 0) return "0";
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.7G", (double)x);
    std::string s(buf);
    const size_t e = s.find('E');
    if (e == std::string::npos) return s;
    std::string mant = s.substr(0, e), exp = s.substr(e + 1);
    if (mant.find('.') != std::string::npos) {
        while (!mant.empty() && mant.back() == '0') mant.pop_back();
        if (!mant.empty() && mant.back() == '.') mant.pop_back();
    }
    const char sign = (exp[0] == '-') ? '-' : '+';
    size_t p = 0;
    while (p < exp.size() && (exp[p] == '+' || exp[p] == '-')) p++;
    while (p + 1 < exp.size() && exp[p] == '0') p++;
    std::string digits = exp.substr(p);
    if (digits.size() < 2) digits = std::string(2 - digits.size(), '0') + digits;
    return mant + "E" + sign + digits;
}
inline void WriteHeader(std::ostream& w, const std::string& type_name) { w << type_name << "\n" << "2.99" << "\n"; }
inline std::string ReadHeader(std::istream& r)
{
    std::string type_name, version;
    if (!std::getline(r, type_name) || type_name.empty()) throw std::runtime_error("Unexpected end of file");
    std::getline(r, version);                        // ignored by the reference too
    return type_name;
}
inline void WriteVector(std::ostream& w, const std::vector<float>& v)
{
    w << v.size() << "\n";
    for (float x : v) w << Fmt(x) << "\n";
}
inline std::vector<float> ReadVector(std::istream& r)
{
    std::string line;
    std::getline(r, line);
    std::vector<float> v((size_t)std::stoll(line));
    for (auto& x : v) { std::getline(r, line); x = (float)std::stod(line); }
    return v;
}
inline void WriteMatrix(std::ostream& w, const std::vector<float>& m, int64_t rows, int64_t cols)
{
    w << rows << " " << cols << "\n";
    for (int64_t i = 0; i < rows; i++)
        for (int64_t j = 0; j < cols; j++) w << i << " " << j << " " << Fmt(m[(size_t)(i * cols + j)]) << "\n";
    w << "\n";
}
inline std::vector<float> ReadMatrix(std::istream& r, int64_t* rows, int64_t* cols)
{
    std::string line;
    std::getline(r, line);
    std::istringstream head(line);
    int64_t d1 = 0, d2 = 0;
    head >> d1 >> d2;
    std::vector<float> m((size_t)(d1 * d2), 0.f);
    while (std::getline(r, line)) {
        std::istringstream row(line);
        int64_t i, j; double v;
        if (!(row >> i >> j >> v)) break;            // the empty line that ends a matrix
        if (i >= d1) throw std::runtime_error("i = " + std::to_string(i) + " >= " + std::to_string(d1));
        if (j >= d2) throw std::runtime_error("j = " + std::to_string(j) + " >= " + std::to_string(d2));
        m[(size_t)(i * d2 + j)] = (float)v;
    }
    *rows = d1; *cols = d2;
    return m;
}
}  // namespace modelio

inline const char* NetBool(bool b) { return b ? "True" : "False"; }

// ---- RatingPrediction.MatrixFactorization on the GPU ------------------------------------------------------------------------
class MatrixFactorization {
public:
    // MatrixFactorization.cs:87-96
    float Regularization = 0.015f, LearnRate = 0.01f, Decay = 1.0f;
    uint32_t NumIter = 30, NumFactors = 10;
    float InitStdDev = 0.1f, InitMean = 0.f;
    uint32_t NumGpus = 1;                              // the one added property
    float MinRating = 1.f, MaxRating = 5.f;
    int32_t MaxUserID = -1, MaxItemID = -1;
    Ratings* ratings = nullptr;                        // the Ratings property (not owned)

    virtual ~MatrixFactorization() { Release(); }
    MatrixFactorization() = default;
    MatrixFactorization(const MatrixFactorization&) = delete;
    MatrixFactorization& operator=(const MatrixFactorization&) = delete;

    virtual std::string TypeName() const { return "MyMediaLite.RatingPrediction.CudaMatrixFactorization"; }
    virtual std::string ClassName() const { return "MatrixFactorization"; }

    bool CanPredict(int32_t user_id, int32_t item_id) const { return user_id <= MaxUserID && item_id <= MaxItemID; }
    float current_learnrate() const
    {
        float lr = 0;
        Check(mml_sgd_get_model(Model(), nullptr, nullptr, nullptr, nullptr, nullptr, &lr));
        return lr;
    }

    // InitModel (MatrixFactorization.cs:99-116): Train() always builds a fresh device model
    virtual void InitModel()
    {
        if (!ratings) throw std::invalid_argument("Ratings is not set");
        Release();
        MaxUserID = ratings->MaxUserID; MaxItemID = ratings->MaxItemID;
        mml_ctx* ctx = Context::ForGpus(NumGpus).get();
        Check(mml_ratings_create(ctx, ratings->Users.data(), ratings->Items.data(), ratings->Values.data(), ratings->Count(),
                                 MaxUserID, MaxItemID, &dev_ratings_));
        if (ratings->Count() > 0) { float avg; Check(mml_ratings_stats(dev_ratings_, &avg, &MinRating, &MaxRating)); }
        mml_mf_params p = Params();
        parallel_ = p.schedule == MML_SCHEDULE_DSGD;
        Check(mml_sgd_create(ctx, dev_ratings_, &p, nullptr, nullptr, &model_));
        if (Engine::device_init()) {
            Check(mml_sgd_init_model(model_, (uint64_t)Random::GetInstance().Next(), InitMean, InitStdDev));
            return;
        }
        Random& rng = Random::GetInstance();           // user matrix first, then the item matrix
        const std::vector<float> U = rng.InitNormal((int64_t)(MaxUserID + 1) * NumFactors, InitMean, InitStdDev);
        const std::vector<float> V = rng.InitNormal((int64_t)(MaxItemID + 1) * NumFactors, InitMean, InitStdDev);
        Check(mml_sgd_set_model(model_, U.data(), V.data(), nullptr, nullptr));
    }
    virtual void Train()
    {
        InitModel();
        for (uint32_t it = 0; it < NumIter; it++) Iterate();
    }
    virtual void Iterate()
    {
        if (parallel_) {
            int32_t G = 0, W = 0; int64_t rounds = 0, staged = 0;
            Check(mml_sgd_strata_info(Model(), &G, &W, &rounds, &staged));
            std::vector<int32_t> subepoch_sequence((size_t)G);
            for (int32_t g = 0; g < G; g++) subepoch_sequence[(size_t)g] = g;
            Random::GetInstance().Shuffle(subepoch_sequence);                       // BiasedMatrixFactorization.cs:210-211
            Check(mml_sgd_iterate(Model(), subepoch_sequence.data(), nullptr, 0));
            return;
        }
        const std::vector<int32_t>& index = ratings->RandomIndex();
        Check(mml_sgd_iterate(Model(), nullptr, index.data(), (int64_t)index.size()));
    }
    // whether Iterate() runs the parallel epoch kernel (see Engine above); several GPUs always do
    virtual int32_t Threads() const { return 1; }
    bool Parallel() const
    {
        if (NumGpus > 1 || Engine::order() == Engine::Order::Parallel) return true;
        if (Engine::order() == Engine::Order::Reference) return Threads() > 1;
        return Threads() > 1 || (ratings && ratings->Count() >= Engine::SerialBelow);
    }
    float Predict(int32_t user_id, int32_t item_id) const
    {
        float out = 0;
        Check(mml_sgd_predict(Model(), &user_id, &item_id, 1, &out));
        return out;
    }
    std::vector<float> Predict(const std::vector<int32_t>& users, const std::vector<int32_t>& items) const
    {
        std::vector<float> out(users.size());
        Check(mml_sgd_predict(Model(), users.data(), items.data(), (int64_t)users.size(), out.data()));
        return out;
    }
    struct Measures { float RMSE, MAE, NMAE, CBD; };
    // Eval.Ratings.Evaluate (Eval/Ratings.cs:96-139) in one device pass
    Measures Evaluate(const Ratings& test) const
    {
        float r[4];
        Check(mml_sgd_evaluate(Model(), test.Users.data(), test.Items.data(), test.Values.data(), test.Count(), r));
        return Measures{r[0], r[1], r[2], r[3]};
    }
    double ComputeObjective() const { double v = 0; Check(mml_sgd_objective(Model(), &v)); return v; }

    // Recommender.Recommend (Recommender.cs:52-103) with Predict as the score
    std::vector<std::pair<int32_t, float>> Recommend(int32_t user_id, int n = -1, const std::vector<int32_t>* ignore_items = nullptr,
                                                     const std::vector<int32_t>* candidate_items = nullptr) const
    {
        std::vector<int32_t> cand;
        if (candidate_items) cand = *candidate_items;
        else for (int32_t i = 0; i < MaxItemID - 1; i++) cand.push_back(i);      // :57-58, the reference's own default
        if (ignore_items)
            cand.erase(std::remove_if(cand.begin(), cand.end(), [&](int32_t c) {
                           return std::find(ignore_items->begin(), ignore_items->end(), c) != ignore_items->end(); }), cand.end());
        const std::vector<float> scores = Predict(std::vector<int32_t>(cand.size(), user_id), cand);
        std::vector<size_t> order(cand.size());
        for (size_t t = 0; t < order.size(); t++) order[t] = t;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return scores[a] > scores[b]; });
        if (n >= 0 && (size_t)n < order.size()) order.resize((size_t)n);
        std::vector<std::pair<int32_t, float>> out;
        for (size_t t : order) out.emplace_back(cand[t], scores[t]);
        return out;
    }

    // RetrainUser / RetrainItem (MatrixFactorization.cs:141-160, BiasedMatrixFactorization.cs:419-431)
    void RetrainUser(int32_t user_id) { Retrain(user_id, false); }
    void RetrainItem(int32_t item_id) { Retrain(item_id, true); }
    // FoldIn (MatrixFactorization.cs:323-347, BiasedMatrixFactorization.cs:445-492): vector and shuffle drawn here, passes on the device
    std::vector<float> FoldIn(std::vector<std::pair<int32_t, float>> rated_items) const
    {
        Random& rng = Random::GetInstance();
        const std::vector<float> init = rng.InitNormal(NumFactors, InitMean, InitStdDev);
        rng.Shuffle(rated_items);
        std::vector<int32_t> items; std::vector<float> values;
        for (auto& t : rated_items) { items.push_back(t.first); values.push_back(t.second); }
        const int64_t ptr[2] = {0, (int64_t)items.size()};
        std::vector<float> out(NumFactors + (Biased() ? 1 : 0));
        Check(mml_sgd_fold_in(Model(), ptr, items.data(), values.data(), 1, init.data(), (int32_t)NumIter, out.data()));
        return out;
    }
    // ScoreItems (MatrixFactorization.cs:350-363)
    std::vector<std::pair<int32_t, float>> ScoreItems(const std::vector<std::pair<int32_t, float>>& rated_items,
                                                      const std::vector<int32_t>& candidate_items) const
    {
        const std::vector<float> v = FoldIn(rated_items);
        std::vector<float> s(candidate_items.size());
        Check(mml_sgd_score_items(Model(), v.data(), 1, candidate_items.data(), (int64_t)candidate_items.size(), s.data()));
        std::vector<std::pair<int32_t, float>> out;
        for (size_t t = 0; t < s.size(); t++) out.emplace_back(candidate_items[t], s[t]);
        return out;
    }

    virtual void SaveModel(const std::string& filename) const
    {
        std::vector<float> U((size_t)(MaxUserID + 1) * NumFactors), V((size_t)(MaxItemID + 1) * NumFactors);
        float gb = 0;
        Check(mml_sgd_get_model(Model(), U.data(), V.data(), nullptr, nullptr, &gb, nullptr));
        std::ofstream w(filename);
        modelio::WriteHeader(w, TypeName());
        w << modelio::Fmt(gb) << "\n";
        modelio::WriteMatrix(w, U, MaxUserID + 1, NumFactors);
        modelio::WriteMatrix(w, V, MaxItemID + 1, NumFactors);
    }
    virtual void LoadModel(const std::string& filename)
    {
        std::ifstream r(filename);
        if (!r) throw std::runtime_error("cannot open " + filename);
        modelio::ReadHeader(r);
        std::string line;
        std::getline(r, line);
        const float bias = (float)std::stod(line);
        int64_t nu, ku, ni, ki;
        const std::vector<float> U = modelio::ReadMatrix(r, &nu, &ku);
        const std::vector<float> V = modelio::ReadMatrix(r, &ni, &ki);
        Adopt(U, nu, ku, V, ni, ki, nullptr, nullptr, bias, MinRating, MaxRating);
    }
    virtual std::string ToString() const
    {
        std::ostringstream s;
        s << ClassName() << " num_factors=" << NumFactors << " regularization=" << modelio::Fmt(Regularization)
          << " learn_rate=" << modelio::Fmt(LearnRate) << " learn_rate_decay=" << modelio::Fmt(Decay) << " num_iter=" << NumIter;
        return s.str();
    }

protected:
    mml_ratings* dev_ratings_ = nullptr;
    mml_sgd* model_ = nullptr;
    bool parallel_ = false;

    virtual bool Biased() const { return false; }
    mml_sgd* Model() const
    {
        if (!model_) throw std::logic_error("the recommender has no model: call Train() or LoadModel() first");
        return model_;
    }
    void Release()
    {
        if (model_) { mml_sgd_destroy(model_); model_ = nullptr; }
        if (dev_ratings_) { mml_ratings_destroy(dev_ratings_); dev_ratings_ = nullptr; }
    }
    virtual mml_mf_params Params() const
    {
        mml_mf_params p;
        mml_mf_params_default(&p);
        p.biased = 0; p.num_factors = (int32_t)NumFactors; p.learn_rate = LearnRate; p.decay = Decay; p.regularization = Regularization;
        p.schedule = Parallel() ? MML_SCHEDULE_DSGD : MML_SCHEDULE_SERIAL;
        return p;
    }
    void Retrain(int32_t id, bool by_item)
    {
        const std::vector<float> row = Random::GetInstance().InitNormal(NumFactors, InitMean, InitStdDev);
        const float zero = 0.f;
        Check(mml_sgd_set_rows(Model(), by_item ? 1 : 0, &id, 1, row.data(), Biased() ? &zero : nullptr));
        std::vector<int32_t> idx;                      // ByUser[u] / ByItem[i]: rating indices in ascending order
        const std::vector<int32_t>& ids = by_item ? ratings->Items : ratings->Users;
        for (int64_t t = 0; t < ratings->Count(); t++) if (ids[(size_t)t] == id) idx.push_back((int32_t)t);
        // LearnFactors (MatrixFactorization.cs:198-202): NumIter passes over the list
        Check(mml_sgd_learn_factors(Model(), idx.data(), (int64_t)idx.size(), by_item ? 0 : 1, by_item ? 1 : 0, (int32_t)NumIter));
    }
    // a model without training data: one pseudo rating per id keeps every row (InitModel zeroes rows without ratings only)
    void Adopt(const std::vector<float>& U, int64_t nu, int64_t ku, const std::vector<float>& V, int64_t ni, int64_t ki,
               const std::vector<float>* bu, const std::vector<float>* bi, float bias, float min_rating, float max_rating)
    {
        if (ku != ki)
            throw std::runtime_error("Number of user and item factors must match: " + std::to_string(ku) + " != " + std::to_string(ki));
        Release();
        MaxUserID = (int32_t)nu - 1; MaxItemID = (int32_t)ni - 1; NumFactors = (uint32_t)ku;
        MinRating = min_rating; MaxRating = max_rating;
        const int64_t n = std::max(nu, ni);
        std::vector<int32_t> uu((size_t)n), ii((size_t)n);
        std::vector<float> vv((size_t)n, min_rating);
        for (int64_t t = 0; t < n; t++) { uu[(size_t)t] = (int32_t)(t % nu); ii[(size_t)t] = (int32_t)(t % ni); }
        mml_ctx* ctx = Context::Default().get();
        Check(mml_ratings_create(ctx, uu.data(), ii.data(), vv.data(), n, MaxUserID, MaxItemID, &dev_ratings_));
        mml_mf_params p = Params();
        p.schedule = MML_SCHEDULE_SERIAL;
        parallel_ = false;
        Check(mml_sgd_create(ctx, dev_ratings_, &p, nullptr, nullptr, &model_));
        Check(mml_sgd_set_model(model_, U.data(), V.data(), bu ? bu->data() : nullptr, bi ? bi->data() : nullptr));
        Check(mml_sgd_set_scale(model_, min_rating, max_rating, bias));
    }
};

// ---- RatingPrediction.BiasedMatrixFactorization on the GPU -----------------------------------------------------------------
class BiasedMatrixFactorization : public MatrixFactorization {
public:
    enum class OptimizationTarget { RMSE, MAE, LogisticLoss };
    // BiasedMatrixFactorization.cs:85-141
    float BiasReg = 0.01f, BiasLearnRate = 1.0f, RegU = 0.015f, RegI = 0.015f;
    bool FrequencyRegularization = false, BoldDriver = false, NaiveParallelization = false;
    OptimizationTarget Loss = OptimizationTarget::RMSE;
    int32_t MaxThreads = 1;

    void SetRegularization(float v) { Regularization = v; RegU = v; RegI = v; }     // the setter of :97-104 fans out

    std::string TypeName() const override { return "MyMediaLite.RatingPrediction.CudaBiasedMatrixFactorization"; }
    std::string ClassName() const override { return "BiasedMatrixFactorization"; }

    int32_t Threads() const override { return MaxThreads; }
    void SaveModel(const std::string& filename) const override
    {
        const int64_t nu = MaxUserID + 1, ni = MaxItemID + 1;
        std::vector<float> U((size_t)nu * NumFactors), V((size_t)ni * NumFactors), bu((size_t)nu), bi((size_t)ni);
        float gb = 0;
        Check(mml_sgd_get_model(Model(), U.data(), V.data(), bu.data(), bi.data(), &gb, nullptr));
        std::ofstream w(filename);                                                  // layout of :339-351
        modelio::WriteHeader(w, TypeName());
        w << modelio::Fmt(gb) << "\n" << modelio::Fmt(MinRating) << "\n" << modelio::Fmt(MaxRating) << "\n";
        modelio::WriteVector(w, bu);
        modelio::WriteMatrix(w, U, nu, NumFactors);
        modelio::WriteVector(w, bi);
        modelio::WriteMatrix(w, V, ni, NumFactors);
    }
    void LoadModel(const std::string& filename) override
    {
        std::ifstream r(filename);
        if (!r) throw std::runtime_error("cannot open " + filename);
        modelio::ReadHeader(r);
        std::string line;
        std::getline(r, line); const float bias = (float)std::stod(line);
        std::getline(r, line); const float mn = (float)std::stod(line);
        std::getline(r, line); const float mx = (float)std::stod(line);
        int64_t nu, ku, ni, ki;
        const std::vector<float> bu = modelio::ReadVector(r);
        const std::vector<float> U = modelio::ReadMatrix(r, &nu, &ku);
        const std::vector<float> bi = modelio::ReadVector(r);
        const std::vector<float> V = modelio::ReadMatrix(r, &ni, &ki);
        if ((int64_t)bu.size() != nu)
            throw std::runtime_error("Number of users must be the same for biases and factors: " + std::to_string(bu.size()) + " != " + std::to_string(nu));
        if ((int64_t)bi.size() != ni)
            throw std::runtime_error("Number of items must be the same for biases and factors: " + std::to_string(bi.size()) + " != " + std::to_string(ni));
        Adopt(U, nu, ku, V, ni, ki, &bu, &bi, bias, mn, mx);
    }
    std::string ToString() const override
    {
        static const char* loss_names[] = {"RMSE", "MAE", "LogisticLoss"};
        std::ostringstream s;
        s << ClassName() << " num_factors=" << NumFactors << " bias_reg=" << modelio::Fmt(BiasReg) << " reg_u=" << modelio::Fmt(RegU)
          << " reg_i=" << modelio::Fmt(RegI) << " frequency_regularization=" << NetBool(FrequencyRegularization)
          << " learn_rate=" << modelio::Fmt(LearnRate) << " bias_learn_rate=" << modelio::Fmt(BiasLearnRate)
          << " learn_rate_decay=" << modelio::Fmt(Decay) << " num_iter=" << NumIter << " bold_driver=" << NetBool(BoldDriver)
          << " loss=" << loss_names[(int)Loss] << " max_threads=" << MaxThreads << " naive_parallelization=" << NetBool(NaiveParallelization);
        return s.str();
    }

protected:
    bool Biased() const override { return true; }
    mml_mf_params Params() const override
    {
        mml_mf_params p = MatrixFactorization::Params();
        p.biased = 1; p.bias_learn_rate = BiasLearnRate; p.bias_reg = BiasReg; p.reg_u = RegU; p.reg_i = RegI;
        p.frequency_regularization = FrequencyRegularization ? 1 : 0;
        p.loss = Loss == OptimizationTarget::MAE ? MML_LOSS_MAE : (Loss == OptimizationTarget::LogisticLoss ? MML_LOSS_LOGISTIC : MML_LOSS_RMSE);
        p.bold_driver = BoldDriver ? 1 : 0; p.max_threads = MaxThreads;
        // MaxThreads > 1 selects the reference's DSGD block schedule (:178-184); on the GPU the worker groups are CTAs, and the
        // parallel kernel is also what MaxThreads = 1 runs unless the engine order says Reference (Engine above)
        // NaiveParallelization (:136-141, :201-204): the list schedule of MultiCore.PartitionIndices (one GPU only)
        if (p.schedule == MML_SCHEDULE_DSGD && NaiveParallelization && NumGpus <= 1) p.schedule = MML_SCHEDULE_NAIVE;
        return p;
    }
};

// ---- ItemRecommendation.WRMF on the GPU ---------------------------------------------------------------------------------------
class WRMF {
public:
    // WRMF.cs:56-65, MF.cs:37-48
    uint32_t NumFactors = 10, NumIter = 15;
    double Alpha = 1.0, Regularization = 0.015, InitMean = 0.0, InitStdDev = 0.1;
    uint32_t NumGpus = 1;
    int32_t MaxUserID = -1, MaxItemID = -1;
    PosOnlyFeedback* Feedback = nullptr;               // not owned

    WRMF() = default;
    WRMF(const WRMF&) = delete;
    WRMF& operator=(const WRMF&) = delete;
    ~WRMF() { Release(); }

    std::string TypeName() const { return "MyMediaLite.ItemRecommendation.CudaWRMF"; }

    void InitModel()
    {
        if (!Feedback) throw std::invalid_argument("Feedback is not set");
        MaxUserID = Feedback->MaxUserID; MaxItemID = Feedback->MaxItemID;
        NewModel(MaxUserID + 1, MaxItemID + 1, Feedback->Users, Feedback->Items);
        Random& rng = Random::GetInstance();           // MF.cs:56-57: user matrix first, no zeroing of empty rows
        const std::vector<float> U = rng.InitNormal((int64_t)(MaxUserID + 1) * NumFactors, InitMean, InitStdDev);
        const std::vector<float> V = rng.InitNormal((int64_t)(MaxItemID + 1) * NumFactors, InitMean, InitStdDev);
        Check(mml_wrmf_set_model(model_, U.data(), V.data()));
    }
    void Train()
    {
        InitModel();
        for (uint32_t it = 0; it < NumIter; it++) Iterate();
    }
    void Iterate() { Check(mml_wrmf_iterate(Model())); cache_.valid = false; }          // WRMF.cs:68-73
    void RetrainUser(int32_t user_id) { Check(mml_wrmf_retrain(Model(), 0, &user_id, 1)); cache_.valid = false; }   // WRMF.cs:159-163
    void RetrainItem(int32_t item_id) { Check(mml_wrmf_retrain(Model(), 1, &item_id, 1)); cache_.valid = false; }   // WRMF.cs:166-170

    float Predict(int32_t user_id, int32_t item_id) const          // MF.cs:151-157
    {
        if (user_id > MaxUserID || item_id > MaxItemID || user_id < 0 || item_id < 0) return std::numeric_limits<float>::lowest();
        const std::vector<int32_t> cand{item_id};
        const auto r = Recommend(user_id, 1, nullptr, &cand);
        return r.empty() ? std::numeric_limits<float>::lowest() : r[0].second;
    }
    std::vector<std::pair<int32_t, float>> Recommend(int32_t user_id, int n = -1, const std::vector<int32_t>* ignore_items = nullptr,
                                                     const std::vector<int32_t>* candidate_items = nullptr) const
    {
        // Eval.Items.Evaluate (Eval/Items.cs:147-164) and WritePredictions (ItemRecommendation/Extensions.cs:65-128) call this
        // once per user with ignore_items = the user's training items, from several threads. The first such call after the
        // model changed computes the lists of ALL users in one device call and keeps them; later calls with the same n and
        // candidates whose ignore list is the user's training row are lookups.
        std::vector<std::pair<int32_t, float>> hit;
        if (FromCache(user_id, n, ignore_items, candidate_items, &hit)) return hit;
        std::vector<std::vector<int32_t>> ign;
        if (ignore_items) ign.push_back(*ignore_items);
        return RecommendMany({user_id}, n, ignore_items ? &ign : nullptr, candidate_items)[0];
    }
    // the all-users loop of ItemRecommendation/Extensions.WritePredictions (:65-128) in one device call
    std::vector<std::vector<std::pair<int32_t, float>>> RecommendMany(const std::vector<int32_t>& users, int n,
            const std::vector<std::vector<int32_t>>* ignore_items = nullptr, const std::vector<int32_t>* candidate_items = nullptr) const
    {
        std::vector<int32_t> cand;
        if (candidate_items) cand = *candidate_items;
        else for (int32_t i = 0; i < MaxItemID - 1; i++) cand.push_back(i);      // Recommender.cs:57-58
        const int64_t n_out = n < 0 ? (int64_t)cand.size() : std::min<int64_t>(n, (int64_t)cand.size());
        std::vector<int64_t> ptr; std::vector<int32_t> idx;
        if (ignore_items) {
            ptr.assign(users.size() + 1, 0);
            for (size_t b = 0; b < users.size(); b++) {
                idx.insert(idx.end(), (*ignore_items)[b].begin(), (*ignore_items)[b].end());
                ptr[b + 1] = (int64_t)idx.size();
            }
            if (idx.empty()) idx.push_back(0);
        }
        std::vector<int32_t> items(std::max<size_t>(users.size() * (size_t)n_out, 1)), counts(std::max<size_t>(users.size(), 1));
        std::vector<float> scores(items.size());
        Check(mml_wrmf_recommend(Model(), users.data(), (int64_t)users.size(), n, cand.data(), (int64_t)cand.size(),
                                 ignore_items ? ptr.data() : nullptr, ignore_items ? idx.data() : nullptr,
                                 items.data(), scores.data(), counts.data()));
        std::vector<std::vector<std::pair<int32_t, float>>> out(users.size());
        for (size_t b = 0; b < users.size(); b++)
            for (int32_t r = 0; r < counts[b]; r++) out[b].emplace_back(items[b * (size_t)n_out + r], scores[b * (size_t)n_out + r]);
        return out;
    }
    void SaveModel(const std::string& filename) const              // ItemRecommendation/MF.cs:160-170
    {
        std::vector<float> U((size_t)(MaxUserID + 1) * NumFactors), V((size_t)(MaxItemID + 1) * NumFactors);
        Check(mml_wrmf_get_model(Model(), U.data(), V.data()));
        std::ofstream w(filename);
        modelio::WriteHeader(w, TypeName());
        modelio::WriteMatrix(w, U, MaxUserID + 1, NumFactors);
        modelio::WriteMatrix(w, V, MaxItemID + 1, NumFactors);
    }
    void LoadModel(const std::string& filename)                    // ItemRecommendation/MF.cs:173-195
    {
        std::ifstream r(filename);
        if (!r) throw std::runtime_error("cannot open " + filename);
        modelio::ReadHeader(r);
        int64_t nu, ku, ni, ki;
        const std::vector<float> U = modelio::ReadMatrix(r, &nu, &ku);
        const std::vector<float> V = modelio::ReadMatrix(r, &ni, &ki);
        if (ku != ki)
            throw std::runtime_error("Number of user and item factors must match: " + std::to_string(ku) + " != " + std::to_string(ki));
        MaxUserID = (int32_t)nu - 1; MaxItemID = (int32_t)ni - 1; NumFactors = (uint32_t)ku;
        NewModel((int32_t)nu, (int32_t)ni, {}, {});
        Check(mml_wrmf_set_model(model_, U.data(), V.data()));
    }
    std::string ToString() const
    {
        std::ostringstream s;
        s << "WRMF num_factors=" << NumFactors << " regularization=" << modelio::Fmt((float)Regularization)
          << " alpha=" << modelio::Fmt((float)Alpha) << " num_iter=" << NumIter;
        return s.str();
    }

private:
    struct ListCache {
        bool valid = false;
        int n = 0;
        std::vector<int32_t> cand;
        std::vector<int64_t> ptr; std::vector<int32_t> idx;      // training rows (sets, ascending) of all users
        std::vector<int32_t> items, counts; std::vector<float> scores;
    };
    mutable ListCache cache_;
    mutable std::mutex cache_mu_;

    bool FromCache(int32_t user_id, int n, const std::vector<int32_t>* ignore_items, const std::vector<int32_t>* candidate_items,
                   std::vector<std::pair<int32_t, float>>* out) const
    {
        if (n <= 0 || user_id < 0 || user_id > MaxUserID || !Feedback || Feedback->Users.empty()) return false;
        std::lock_guard<std::mutex> lock(cache_mu_);
        std::vector<int32_t> cand;
        if (candidate_items) cand = *candidate_items;
        else for (int32_t i = 0; i < MaxItemID - 1; i++) cand.push_back(i);
        ListCache& c = cache_;
        if (c.ptr.empty()) {                                     // training rows: the user matrix is a set
            const size_t nu = (size_t)MaxUserID + 1;
            std::vector<std::pair<int32_t, int32_t>> ev(Feedback->Users.size());
            for (size_t t = 0; t < ev.size(); t++) ev[t] = {Feedback->Users[t], Feedback->Items[t]};
            std::sort(ev.begin(), ev.end());
            ev.erase(std::unique(ev.begin(), ev.end()), ev.end());
            c.ptr.assign(nu + 1, 0);
            for (auto& e : ev) c.ptr[(size_t)e.first + 1]++;
            for (size_t u = 0; u < nu; u++) c.ptr[u + 1] += c.ptr[u];
            c.idx.resize(std::max<size_t>(ev.size(), 1));
            for (size_t t = 0; t < ev.size(); t++) c.idx[t] = ev[t].second;
        }
        std::vector<int32_t> given = ignore_items ? *ignore_items : std::vector<int32_t>();
        std::sort(given.begin(), given.end());
        given.erase(std::unique(given.begin(), given.end()), given.end());
        const int64_t lo = c.ptr[(size_t)user_id], hi = c.ptr[(size_t)user_id + 1];
        if ((int64_t)given.size() != hi - lo || !std::equal(given.begin(), given.end(), c.idx.begin() + lo)) return false;
        const int64_t n_out = std::min<int64_t>(n, (int64_t)cand.size());
        if (!c.valid || c.n != n || c.cand != cand) {
            const size_t nu = (size_t)MaxUserID + 1;
            std::vector<int32_t> users(nu);
            for (size_t u = 0; u < nu; u++) users[u] = (int32_t)u;
            c.items.assign(std::max<size_t>(nu * (size_t)n_out, 1), 0); c.scores.assign(c.items.size(), 0.f); c.counts.assign(nu, 0);
            Check(mml_wrmf_recommend(Model(), users.data(), (int64_t)nu, n, cand.data(), (int64_t)cand.size(), c.ptr.data(), c.idx.data(),
                                     c.items.data(), c.scores.data(), c.counts.data()));
            c.valid = true; c.n = n; c.cand = cand;
        }
        out->clear();
        for (int32_t r = 0; r < c.counts[(size_t)user_id]; r++)
            out->emplace_back(c.items[(size_t)user_id * (size_t)n_out + r], c.scores[(size_t)user_id * (size_t)n_out + r]);
        return true;
    }

    mml_feedback* fb_ = nullptr;
    mml_wrmf* model_ = nullptr;
    mml_wrmf* Model() const
    {
        if (!model_) throw std::logic_error("the recommender has no model: call Train() or LoadModel() first");
        return model_;
    }
    void Release()
    {
        if (model_) { mml_wrmf_destroy(model_); model_ = nullptr; }
        if (fb_) { mml_feedback_destroy(fb_); fb_ = nullptr; }
    }
    void NewModel(int32_t n_users, int32_t n_items, const std::vector<int32_t>& users, const std::vector<int32_t>& items)
    {
        Release();
        cache_ = ListCache();
        mml_ctx* ctx = Context::ForGpus(NumGpus).get();
        Check(mml_feedback_create(ctx, users.data(), items.data(), (int64_t)users.size(), n_users - 1, n_items - 1, &fb_));
        mml_wrmf_params p;
        p.num_factors = (int32_t)NumFactors; p.alpha = Alpha; p.regularization = Regularization;
        Check(mml_wrmf_create(ctx, fb_, &p, &model_));
    }
};

}  // namespace mymedialite
#endif  // MMLB200_HPP
