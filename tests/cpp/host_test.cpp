// host_test.cpp -- drives the C++ host layer (include/mmlb200.hpp) the way the reference's NUnit tests and CLI drive the
// C# classes. Built and run by tests/test_cpp_host.py:
//   host_test cpu                 checks that need no device: System.Random known answers, number formatting, ToString(),
//                                 hyper-parameter defaults, and that creating a context without a CUDA device THROWS (no CPU path)
//   host_test gpu TRAIN TEST DIR  config 1 on the device: prints one "key value..." line per result for the Python test to
//                                 hold against tests/golden/oracle_config1.json
//   host_test multi N [RATINGS]   one process, N GPUs (the NumGpus property): BiasedMatrixFactorization on a synthetic set of
//                                 the config 2 shape (71.5k x 10.7k, RATINGS ratings, default 10M, k = 64) trained with
//                                 NumGpus = 1 and NumGpus = N from the same initial model; WRMF with NumGpus = N against
//                                 NumGpus = 1; and 10^4 per-user Recommend() calls served from the cached all-users result
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "mmlb200.hpp"

using namespace mymedialite;

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

static int run_cpu()
{
    // System.Random: public known answers
    { Random r(0); EXPECT(r.Next() == 1559595546); EXPECT(r.Next() == 1755192844); EXPECT(r.Next() == 1649316166); }
    { Random r(1); EXPECT(r.Next() == 534011718); EXPECT(r.Next() == 237820880); EXPECT(r.Next() == 1002897798); }
    { Random r(42); EXPECT(r.Next() == 1434747710); EXPECT(r.Next() == 302596119); EXPECT(r.Next() == 269548474); }
    {   // Shuffle = applying the swap targets one by one (Utils.cs:52-64); both draw n numbers
        Random a(7), b(7);
        std::vector<int32_t> v{0, 1, 2, 3, 4, 5, 6, 7, 8};
        a.Shuffle(v);
        std::vector<int32_t> w{0, 1, 2, 3, 4, 5, 6, 7, 8};
        const std::vector<int32_t> H = b.ShuffleTargets(9);
        for (int i = 8; i >= 0; i--) std::swap(w[i], w[H[i]]);
        EXPECT(v == w);
        EXPECT(a.Next() == b.Next());
    }
    {   // Normal.Sample consumes pairs of uniforms; values are finite and centred
        Random r(1);
        double s = 0;
        for (int i = 0; i < 20000; i++) s += r.Normal(0.0, 0.1);
        EXPECT(std::fabs(s / 20000) < 0.005);
    }
    // float.ToString(InvariantCulture) as the .NET Framework prints it
    EXPECT(modelio::Fmt(0.015f) == "0.015"); EXPECT(modelio::Fmt(1.0f) == "1"); EXPECT(modelio::Fmt(0.f) == "0");
    EXPECT(modelio::Fmt(1e-5f) == "1E-05"); EXPECT(modelio::Fmt(1.5e10f) == "1.5E+10"); EXPECT(modelio::Fmt(3.14159274f) == "3.141593");
    EXPECT(modelio::Fmt(-2.5f) == "-2.5"); EXPECT(modelio::Fmt(1234567.f) == "1234567"); EXPECT(modelio::Fmt(12345678.f) == "1.234568E+07");
    // defaults and ToString() of the reference classes (MatrixFactorization.cs:87-96, 411-417; BiasedMatrixFactorization.cs:85-141, 555-561)
    MatrixFactorization mf;
    EXPECT(mf.NumFactors == 10 && mf.NumIter == 30 && mf.LearnRate == 0.01f && mf.Regularization == 0.015f && mf.Decay == 1.0f);
    EXPECT(mf.ToString() == "MatrixFactorization num_factors=10 regularization=0.015 learn_rate=0.01 learn_rate_decay=1 num_iter=30");
    BiasedMatrixFactorization bmf;
    EXPECT(bmf.BiasReg == 0.01f && bmf.BiasLearnRate == 1.0f && bmf.MaxThreads == 1 && bmf.RegU == 0.015f && bmf.RegI == 0.015f);
    bmf.SetRegularization(0.05f);
    EXPECT(bmf.RegU == 0.05f && bmf.RegI == 0.05f);
    bmf.SetRegularization(0.015f);
    EXPECT(bmf.ToString() == "BiasedMatrixFactorization num_factors=10 bias_reg=0.01 reg_u=0.015 reg_i=0.015 frequency_regularization=False "
                             "learn_rate=0.01 bias_learn_rate=1 learn_rate_decay=1 num_iter=30 bold_driver=False loss=RMSE max_threads=1 "
                             "naive_parallelization=False");
    WRMF w;
    EXPECT(w.NumFactors == 10 && w.NumIter == 15 && w.Alpha == 1.0 && w.Regularization == 0.015);
    EXPECT(w.ToString() == "WRMF num_factors=10 regularization=0.015 alpha=1 num_iter=15");
    // model text format round trip (no device)
    {
        std::stringstream s;
        const std::vector<float> m{1.5f, -2.f, 0.f, 1e-5f, 3.25f, 7.f};
        modelio::WriteMatrix(s, m, 2, 3);
        modelio::WriteVector(s, {0.5f, 2.f});
        int64_t r, c;
        const std::vector<float> back = modelio::ReadMatrix(s, &r, &c);
        EXPECT(r == 2 && c == 3 && back == m);
        EXPECT((modelio::ReadVector(s) == std::vector<float>{0.5f, 2.f}));
    }
    // calls before Train() fail loudly
    try { mf.Predict(0, 0); EXPECT(false); } catch (const std::logic_error&) {}
    // reader errors arrive as FormatException with the reference's message
    {
        mml_ingest* g = nullptr;
        const char* text = "1\t2\n";
        const int32_t st = mml_ingest_text(text, 4, MML_FILE_RATINGS, MML_MAP_IDENTITY, MML_MAP_IDENTITY, 0, 1, nullptr, &g);
        try { Check(st); EXPECT(false); }
        catch (const FormatException& e) { EXPECT(std::string(e.what()) == "Expected at least 3 columns: 1\t2"); }
    }
    // no CPU fallback: without a device the context cannot exist
    bool have_device = true;
    try { Context c(0); } catch (const MmlError& e) { have_device = false; EXPECT(e.status == MML_ERR_CUDA); }
    std::printf("device %s\n", have_device ? "present" : "absent: Context() threw, as it must");
    std::printf(failures ? "FAILED %d\n" : "OK\n", failures);
    return failures ? 1 : 0;
}

static int run_gpu(const char* train_file, const char* test_file, const char* dir)
{
    Ratings train = StaticRatingData::Read(train_file);
    Ratings test = StaticRatingData::Read(test_file);
    std::printf("counts %lld %lld max %d %d\n", (long long)train.Count(), (long long)test.Count(), train.MaxUserID, train.MaxItemID);
    for (int biased = 1; biased >= 0; biased--) {
        // config 1: num_factors=10 num_iter=30 --random-seed=1; Train() = InitModel + NumIter x Iterate
        Random::Seed(1);
        std::unique_ptr<MatrixFactorization> rec(biased ? new BiasedMatrixFactorization() : new MatrixFactorization());
        Ratings data = train;                            // RandomIndex is cached per data set
        rec->ratings = &data;
        rec->InitModel();
        const char* name = biased ? "BiasedMatrixFactorization" : "MatrixFactorization";
        for (uint32_t it = 0; it < rec->NumIter; it++) {
            rec->Iterate();
            std::printf("epoch %s %u %.9g %.9g\n", name, it, rec->Evaluate(data).RMSE, rec->Evaluate(test).RMSE);
        }
        std::printf("predict %s", name);
        for (float p : rec->Predict(test.Users, test.Items)) std::printf(" %.9g", p);
        std::printf("\n");
        // SaveModel / LoadModel round trip into a fresh object: predictions within the text format's 7 digits
        const std::string path = std::string(dir) + "/" + name + ".model";
        rec->SaveModel(path);
        std::unique_ptr<MatrixFactorization> back(biased ? new BiasedMatrixFactorization() : new MatrixFactorization());
        back->LoadModel(path);
        double worst = 0;
        const std::vector<float> p0 = rec->Predict(test.Users, test.Items), p1 = back->Predict(test.Users, test.Items);
        for (size_t t = 0; t < p0.size(); t++) worst = std::max(worst, (double)std::fabs(p0[t] - p1[t]));
        std::printf("roundtrip %s %.3g %d %d\n", name, worst, back->MaxUserID, back->MaxItemID);
        // fold-in: three results, descending, taken from the candidates (FoldInRatingPredictorExtensionsTest.cs:43-62)
        const std::vector<int32_t> cand{0, 1, 2};
        auto scored = rec->ScoreItems({{0, 1.0f}, {1, 4.0f}}, cand);
        std::printf("scoreitems %s %zu %.9g %.9g %.9g\n", name, scored.size(), scored[0].second, scored[1].second, scored[2].second);
        rec->RetrainUser(0);
        std::printf("retrain %s %.9g\n", name, rec->Predict(0, 0));
    }
    // WRMF on the same pairs read as implicit feedback: k = 4, 3 epochs, top-2 without the training items
    PosOnlyFeedback fb;
    for (int64_t t = 0; t < train.Count(); t++) fb.Add(train.Users[(size_t)t], train.Items[(size_t)t]);
    Random::Seed(1);
    WRMF w;
    w.NumFactors = 4; w.NumIter = 3; w.Feedback = &fb;
    w.Train();
    for (int32_t u = 0; u <= fb.MaxUserID; u++) {
        std::vector<int32_t> ign, cand;
        for (int64_t t = 0; t < fb.Count(); t++) if (fb.Users[(size_t)t] == u) ign.push_back(fb.Items[(size_t)t]);
        for (int32_t i = 0; i <= fb.MaxItemID; i++) cand.push_back(i);
        std::printf("wrmf_top2 %d", u);
        for (auto& r : w.Recommend(u, 2, &ign, &cand)) std::printf(" %d %.9g", r.first, r.second);
        std::printf("\n");
    }
    std::printf("tostring %s\n", w.ToString().c_str());
    std::printf("OK\n");
    return 0;
}

// xorshift64*: test data only
struct Xs {
    uint64_t s;
    explicit Xs(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 1) {}
    uint64_t next() { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 0x2545F4914F6CDD1Dull; }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double gauss() { double a = 0; for (int i = 0; i < 6; i++) a += uni(); return (a - 3.0) * 1.41421356; }   // ~N(0,1)
};

static int run_multi(uint32_t n_gpus, int64_t n_ratings)
{
    using clk = std::chrono::steady_clock;
    Engine::device_init() = true;                        // both runs start from the same counter-based model
    // planted model of the config 2 shape: skewed user activity and item popularity, half-star levels
    const int32_t nu = 71500, ni = 10700;
    Xs g(20260102);
    std::vector<float> bu((size_t)nu), bi((size_t)ni), pu((size_t)nu * 4), qi((size_t)ni * 4);
    for (auto& x : bu) x = 0.3f * (float)g.gauss();
    for (auto& x : bi) x = 0.3f * (float)g.gauss();
    for (auto& x : pu) x = 0.35f * (float)g.gauss();
    for (auto& x : qi) x = 0.35f * (float)g.gauss();
    Ratings train, test;
    for (int64_t t = 0; t < n_ratings + n_ratings / 10; t++) {
        const double a = g.uni(), b = g.uni();
        const int32_t u = std::min<int32_t>(nu - 1, (int32_t)(nu * a * a)), i = std::min<int32_t>(ni - 1, (int32_t)(ni * b * b * b));
        float v = 3.6f + bu[(size_t)u] + bi[(size_t)i] + 0.5f * (float)g.gauss();
        for (int f = 0; f < 4; f++) v += pu[(size_t)u * 4 + f] * qi[(size_t)i * 4 + f];
        v = std::min(5.0f, std::max(0.5f, std::round(v * 2.f) / 2.f));
        (t < n_ratings ? train : test).Add(u, i, v);
    }
    train.MaxUserID = test.MaxUserID = nu - 1; train.MaxItemID = test.MaxItemID = ni - 1;
    std::printf("multi data %lld %lld\n", (long long)train.Count(), (long long)test.Count());
    for (uint32_t gpus : {1u, n_gpus}) {
        Random::Seed(1);
        BiasedMatrixFactorization rec;
        rec.NumFactors = 64; rec.NumGpus = gpus; rec.ratings = &train;
        const auto t0 = clk::now();
        rec.InitModel();
        const auto t1 = clk::now();
        for (int it = 0; it < 6; it++) {
            rec.Iterate();
            std::printf("multi epoch %u %d %.6f\n", gpus, it, rec.Evaluate(test).RMSE);
        }
        const auto t2 = clk::now();
        std::printf("multi time %u init_s %.3f epochs_s %.3f predict %.6f\n", gpus, std::chrono::duration<double>(t1 - t0).count(),
                    std::chrono::duration<double>(t2 - t1).count(), rec.Predict(5, 7));
    }
    // WRMF: NumGpus = N against NumGpus = 1 (rows are solved on different GPUs, the arithmetic is the same)
    Engine::device_init() = false;
    PosOnlyFeedback fb;
    const int32_t wu = 20000, wi = 5000;
    for (int64_t t = 0; t < 1000000; t++) {
        const double a = g.uni(), b = g.uni();
        fb.Add(std::min<int32_t>(wu - 1, (int32_t)(wu * a * a)), std::min<int32_t>(wi - 1, (int32_t)(wi * b * b)));
    }
    fb.Add(wu - 1, wi - 1);
    std::vector<std::vector<std::pair<int32_t, float>>> lists[2];
    std::vector<int32_t> cand;
    for (int32_t i = 0; i <= fb.MaxItemID; i++) cand.push_back(i);
    std::vector<std::vector<int32_t>> train_rows((size_t)wu);
    for (int64_t t = 0; t < fb.Count(); t++) train_rows[(size_t)fb.Users[(size_t)t]].push_back(fb.Items[(size_t)t]);
    int x = 0;
    for (uint32_t gpus : {1u, n_gpus}) {
        Random::Seed(3);
        WRMF w;
        w.NumFactors = 32; w.NumIter = 2; w.NumGpus = gpus; w.Feedback = &fb;
        w.Train();
        const auto t0 = clk::now();
        auto first = w.Recommend(0, 10, &train_rows[0], &cand);              // computes and caches the lists of all users
        const auto t1 = clk::now();
        lists[x].push_back(first);
        for (int32_t u = 1; u < 10001; u++) lists[x].push_back(w.Recommend(u, 10, &train_rows[(size_t)u], &cand));
        const auto t2 = clk::now();
        std::printf("multi wrmf %u first_call_ms %.2f next_10000_calls_ms %.2f\n", gpus,
                    std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
        x++;
    }
    size_t same = 0;
    for (size_t u = 0; u < lists[0].size(); u++) same += lists[0][u] == lists[1][u] ? 1 : 0;
    std::printf("multi wrmf_lists_identical %zu of %zu\n", same, lists[0].size());
    std::printf("OK\n");
    return 0;
}

int main(int argc, char** argv)
{
    try {
        if (argc >= 2 && std::strcmp(argv[1], "cpu") == 0) return run_cpu();
        if (argc >= 3 && std::strcmp(argv[1], "multi") == 0)
            return run_multi((uint32_t)std::atoi(argv[2]), argc >= 4 ? std::atoll(argv[3]) : 10000000);
        if (argc >= 5 && std::strcmp(argv[1], "gpu") == 0) return run_gpu(argv[2], argv[3], argv[4]);
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
    std::printf("usage: host_test cpu | host_test gpu TRAIN TEST DIR | host_test multi N [RATINGS]\n");
    return 64;
}
