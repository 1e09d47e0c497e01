"""Oracle vs. the reference's NUnit known answers for the learn-rate schedule
(src/Tests/RatingPrediction/BiasedMatrixFactorizationTest.cs:30-62, MatrixFactorizationTest.cs:29-61)
plus internal consistency of the restated SGD loops."""
import numpy as np
import pytest

from oracle import oracle as O

ONE = (np.array([0], np.int32), np.array([0], np.int32), np.array([0.0], np.float32))


@pytest.mark.parametrize("biased", [True, False])
def test_current_learnrate_after_init(biased):
    m = O.Model(*ONE, biased=biased, learn_rate=1.1)
    m.init(O.Random(1))
    assert m.learnrate == np.float32(1.1)


@pytest.mark.parametrize("biased", [True, False])
def test_default_is_no_decay(biased):
    m = O.Model(*ONE, biased=biased, learn_rate=1.1, num_iter=10)
    m.train(O.Random(1))
    assert m.learnrate == np.float32(1.1)


@pytest.mark.parametrize("biased", [True, False])
def test_decay(biased):
    m = O.Model(*ONE, biased=biased, learn_rate=1.0, decay=0.5, num_iter=1)
    rng = O.Random(1)
    m.train(rng)
    assert m.learnrate == 0.5
    m.iterate(rng)
    assert m.learnrate == 0.25


def test_decay_applied_twice_when_multithreaded():
    # BiasedMatrixFactorization.cs:216 and :221
    u = np.array([0, 1, 2, 3], np.int32); i = np.array([0, 1, 2, 3], np.int32)
    v = np.array([1, 2, 3, 4], np.float32)
    m = O.Model(u, i, v, biased=True, learn_rate=1.0, decay=0.5, num_iter=1, max_threads=2)
    m.train(O.Random(1))
    assert m.learnrate == 0.25


def test_rows_without_ratings_are_zeroed(example_data):
    u, i, v, *_ = example_data
    m = O.Model(u, i, v, biased=True, max_user=6, max_item=5)
    m.init(O.Random(1))
    assert np.all(m.user_factors[5:] == 0) and np.all(m.item_factors[4:] == 0)
    assert np.all(m.user_factors[:5] != 0)
    assert np.all(m.user_bias == 0) and np.all(m.item_bias == 0)


def test_init_draw_order_user_then_item(example_data):
    u, i, v, *_ = example_data
    m = O.Model(u, i, v, biased=True, num_factors=10)
    m.init(O.Random(1))
    r = O.Random(1)
    exp_u = r.init_normal(5 * 10)
    exp_i = r.init_normal(4 * 10)
    assert np.array_equal(m.user_factors.ravel(), exp_u)
    assert np.array_equal(m.item_factors.ravel(), exp_i)


def test_global_bias_is_logit_of_scaled_average(example_data):
    u, i, v, *_ = example_data
    m = O.Model(u, i, v, biased=True)
    m.init(O.Random(1))
    avg = np.float32(np.float32(v.astype(np.float64).sum()) / np.float32(v.size))
    a = float(np.float32((avg - np.float32(1.0)) / np.float32(4.0)))
    assert m.global_bias == np.float32(np.log(a / (1 - a)))
    m2 = O.Model(u, i, v, biased=False)
    m2.init(O.Random(1))
    assert m2.global_bias == avg


def test_config1_training_reduces_error_and_is_deterministic(example_data):
    u, i, v, tu, ti, tv = example_data
    runs = []
    for _ in range(2):
        m = O.Model(u, i, v, biased=True, num_factors=10, num_iter=30)
        m.train(O.Random(1))
        runs.append((m.user_factors.copy(), m.evaluate(u, i, v)["RMSE"], m.evaluate(tu, ti, tv)["RMSE"]))
    assert np.array_equal(runs[0][0], runs[1][0])
    m0 = O.Model(u, i, v, biased=True, num_factors=10, num_iter=0)
    m0.train(O.Random(1))
    assert runs[0][1] < m0.evaluate(u, i, v)["RMSE"]


def test_dsgd_blocks_cover_every_rating_once():
    # one DSGD epoch with g threads touches each rating exactly once: with lr chosen so that
    # updates are visible, compare against running the same blocks by hand
    rng = np.random.default_rng(0)
    n = 500
    u = rng.integers(0, 40, n).astype(np.int32); i = rng.integers(0, 30, n).astype(np.int32)
    v = rng.integers(1, 6, n).astype(np.float32)
    m = O.Model(u, i, v, biased=True, num_factors=8, num_iter=1, max_threads=4)
    r = O.Random(7)
    m.init(r)
    U0, V0 = m.user_factors.copy(), m.item_factors.copy()
    m.iterate(r)
    # replay by hand
    m2 = O.Model(u, i, v, biased=True, num_factors=8, num_iter=1, max_threads=1)
    r2 = O.Random(7)
    m2.init(r2)           # same init draws (partition draws come after init)
    assert np.array_equal(m2.user_factors, U0) and np.array_equal(m2.item_factors, V0)
    g, ptr, idx, up, ip = O.partition_users_and_items(r2, u, i, int(u.max()), int(i.max()), 4)
    seq = r2.shuffle(np.arange(g))
    for s in seq:
        for j in range(g):
            b = j * g + (s + j) % g
            m2.iterate_indices(idx[ptr[b]:ptr[b + 1]])
    assert np.array_equal(m2.user_factors, m.user_factors)
    assert np.array_equal(m2.item_factors, m.item_factors)
    assert np.array_equal(m2.user_bias, m.user_bias)


def test_omp_dsgd_is_bit_identical_to_sequential_dsgd():
    rng = np.random.default_rng(1)
    n = 5000
    u = rng.integers(0, 300, n).astype(np.int32); i = rng.integers(0, 200, n).astype(np.int32)
    v = rng.integers(1, 6, n).astype(np.float32)
    res = []
    for th in (1, 4):
        m = O.Model(u, i, v, biased=True, num_factors=16, num_iter=3, max_threads=8, omp_threads=th)
        m.train(O.Random(3))
        res.append((m.user_factors.copy(), m.item_factors.copy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_recommend_order_and_ignore():
    U = np.array([[1.0, 0.0], [0.0, 1.0]], np.float32)
    V = np.array([[0.5, 0.1], [0.9, 0.2], [0.5, 0.3], [0.1, 0.9]], np.float32)
    items, scores = O.recommend_mf(U, V, 0)
    assert items.tolist() == [1, 0, 2, 3]           # tie 0/2 keeps candidate order (stable sort)
    items, scores = O.recommend_mf(U, V, 0, n=2, ignore=[1])
    assert items.tolist() == [0, 2]
    items, _ = O.recommend_mf(U, V, 0, n=-1, candidates=[3, 2, 0])
    assert items.tolist() == [2, 0, 3]
    items, _ = O.recommend_mf(U, V, 5, n=-1)        # unknown user: float.MinValue scores are dropped
    assert items.size == 0


def test_wrmf_half_sweep_solves_normal_equations():
    rng = np.random.default_rng(0)
    nu, ni, k = 30, 20, 6
    H = rng.normal(0, 0.1, (ni, k)).astype(np.float32)
    W = rng.normal(0, 0.1, (nu, k)).astype(np.float32)
    rows = rng.integers(0, nu, 150).astype(np.int32); cols = rng.integers(0, ni, 150).astype(np.int32)
    rows[rows == 7] = 8          # leave user 7 empty
    ptr, c = O.feedback_csr(rows, cols, nu - 1)
    O.wrmf_optimize(ptr, c, W, H, alpha=1.0, regularization=0.015)
    assert np.all(W[7] == 0)     # HCp = 0 -> exact zero row
    Hd = H.astype(np.float64)
    for u in (0, 3, 29):
        S = c[ptr[u]:ptr[u + 1]]
        A = Hd.T @ Hd + 1.0 * Hd[S].T @ Hd[S] + 0.015 * np.eye(k)
        b = 2.0 * Hd[S].sum(0)
        assert np.allclose(W[u], np.linalg.solve(A, b), rtol=1e-5, atol=1e-7)
    assert np.allclose(O.wrmf_gram(H), Hd.T @ Hd, rtol=1e-6)


# ---- FoldIn (MatrixFactorization.cs:323-347, BiasedMatrixFactorization.cs:445-492) -------------------------------------
def _fold_in_numpy(m, items, values, init, biased, num_iter):
    """Second, independent restatement in numpy scalars (float32 for C# float, float for double)."""
    import math
    f32 = np.float32
    V, k = m.item_factors, m.k
    p = m.params
    gb = f32(m.global_bias)

    def dot(a, b):
        r = f32(0)
        for x, y in zip(a, b):
            r = f32(r + f32(x * y))
        return r
    vec = init.astype(np.float32).copy()
    if not biased:
        lr = float(p.learn_rate)
        for _ in range(num_iter):
            for it, r in zip(items, values):
                err = f32(f32(r) - f32(gb + dot(V[it], vec)))
                for f in range(k):
                    d = f32(f32(err * V[it, f]) - f32(f32(p.regularization) * vec[f]))
                    vec[f] = f32(vec[f] + f32(lr * float(d)))
            lr *= float(f32(p.decay))
        return vec
    mn, rng_size = f32(m.values.min()), f32(f32(m.values.max()) - f32(m.values.min()))
    ub = f32(0)
    regw = f32(p.reg_u)
    for _ in range(num_iter):
        for it, r in zip(items, values):
            score = float(f32(f32(f32(gb + ub) + m.item_bias[it]) + dot(V[it], vec)))
            sig = 1 / (1 + math.exp(-score))
            err = float(f32(r)) - (float(mn) + sig * float(rng_size))
            gc = f32(err * sig * (1 - sig) * float(rng_size))
            ub = f32(ub + f32(f32(f32(p.bias_learn_rate) * f32(p.learn_rate)) * f32(gc - f32(f32(f32(p.bias_reg) * regw) * ub))))
            for f in range(k):
                d = f32(f32(gc * V[it, f]) - f32(regw * vec[f]))
                vec[f] = f32(vec[f] + f32(float(f32(p.learn_rate)) * float(d)))
    return np.concatenate([[ub], vec]).astype(np.float32)


@pytest.mark.parametrize("biased", [True, False])
def test_fold_in_two_restatements_agree(biased):
    rng = np.random.default_rng(2)
    n_users, n_items, n = 40, 25, 600
    u = rng.integers(0, n_users, n).astype(np.int32); i = rng.integers(0, n_items, n).astype(np.int32)
    u[:n_users] = np.arange(n_users); i[:n_items] = np.arange(n_items)
    v = (rng.integers(1, 11, n) / 2).astype(np.float32)
    m = O.Model(u, i, v, biased=biased, num_factors=7, num_iter=4, decay=0.95 if not biased else 1.0)
    m.init(O.Random(5))
    m.iterate(O.Random(6))
    items = rng.integers(0, n_items, 9).astype(np.int32)
    values = (rng.integers(1, 11, 9) / 2).astype(np.float32)
    init = (rng.standard_normal(7) * 0.1).astype(np.float32)
    with np.errstate(over="ignore"):
        want = _fold_in_numpy(m, items, values, init, biased, 4)
    got = m.fold_in(items, values, init)
    if biased:
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)   # libm exp vs math.exp: at most a last-bit difference
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # no ratings: the vector is the drawn one
    empty = m.fold_in([], [], init)
    assert np.array_equal(empty[-7:], init) and (not biased or empty[0] == 0)
    # Predict(vector, item): clipped for the plain model, sigmoid link for the biased one
    s = m.predict_vector(got, 3)
    assert v.min() <= s <= v.max()


def test_bold_driver_rule_in_the_oracle():
    """BiasedMatrixFactorization.cs:225-244: lr * 0.5 when the objective grew, * 1.05 when it shrank; the first comparison
    is against the loss of InitModel, computed while rating_range_size and global_bias are still 0 (:161-170 vs :186-190)."""
    rng = np.random.default_rng(3)
    n_users, n_items, n = 60, 30, 2500
    u = rng.integers(0, n_users, n).astype(np.int32); i = rng.integers(0, n_items, n).astype(np.int32)
    v = (rng.integers(1, 11, n) / 2).astype(np.float32)
    for lr0, expect_halving in ((0.01, False), (0.9, True)):
        m = O.Model(u, i, v, biased=True, num_factors=6, bold_driver=1, learn_rate=lr0)
        r = O.Random(1)
        m.init(r)
        seq = [np.float32(lr0)]
        for _ in range(8):
            m.iterate(r)
            seq.append(np.float32(m.learnrate))
        ratios = {round(float(b / a), 4) for a, b in zip(seq[:-1], seq[1:])}
        assert ratios <= {0.5, 1.05}, seq
        assert (0.5 in ratios) == expect_halving, seq
        if not expect_halving:
            assert round(float(seq[1] / seq[0]), 4) == 1.05   # a sane first epoch beats "every prediction = min_rating"
