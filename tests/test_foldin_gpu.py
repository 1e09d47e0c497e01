"""Fold-in and incremental operations on the device (mml_sgd_fold_in / score_items / set_rows + iterate_indices) against
the oracle's restatement of MatrixFactorization.cs:141-160, 323-363 and BiasedMatrixFactorization.cs:419-492, and the
properties the reference's own test checks (src/Tests/RatingPrediction/FoldInRatingPredictorExtensionsTest.cs:43-81)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = 2e-6   # same operations in the same order; only exp() of the double-precision link may differ in its last bit


@pytest.fixture(scope="module")
def eng():
    from mymedialite_b200 import engine
    ctx = engine.Context(0)
    yield engine, ctx
    ctx.close()


def trained_pair(eng, biased, k, **kw):
    """A model trained for two serial epochs on both sides (so that factors and biases are not trivial)."""
    from mymedialite_b200 import synthetic
    engine, ctx = eng
    d = synthetic.ratings(200, 90, 9000, "half", 21)
    u, i, v = d["train"]
    om = O.Model(u, i, v, biased=biased, num_factors=k, **kw)
    rng = O.Random(3)
    om.init(rng)
    r = engine.DeviceRatings(ctx, u, i, v)
    gm = engine.SgdModel(ctx, r, engine.default_params(biased=int(biased), num_factors=k,
                                                        **{a: b for a, b in kw.items() if a != "num_iter"}))
    gm.set_model(om.user_factors.copy(), om.item_factors.copy())
    idx = np.arange(u.size, dtype=np.int32)
    for _ in range(2):
        om.iterate_indices(idx)
        gm.iterate_indices(idx)
    return (u, i, v), om, gm, r


def fold_in_cases(rng, n_items, k, n_users):
    items, values = [], []
    for j in range(n_users):
        cnt = [0, 1, 2, 7, 40, 150][j % 6]
        items.append(rng.integers(0, n_items, cnt).astype(np.int32))
        values.append((rng.integers(1, 11, cnt) / 2).astype(np.float32))
    init = (rng.standard_normal((n_users, k)) * 0.1).astype(np.float32)
    return items, values, init


@pytest.mark.parametrize("biased,k,kw", [
    (True, 10, {}), (True, 64, {"frequency_regularization": 1}), (True, 128, {"loss": 1}), (True, 40, {"loss": 2}),
    (False, 10, {}), (False, 96, {"decay": 0.9}),
])
def test_fold_in_matches_oracle(eng, biased, k, kw):
    (u, i, v), om, gm, r = trained_pair(eng, biased, k, num_iter=7, **kw)
    rng = np.random.default_rng(k)
    items, values, init = fold_in_cases(rng, int(i.max()) + 1, k, 13)
    got = gm.fold_in(items, values, init, 7)
    assert got.shape == (13, k + (1 if biased else 0))
    for j in range(13):
        want = om.fold_in(items[j], values[j], init[j])
        np.testing.assert_allclose(got[j], want, rtol=TOL, atol=TOL, err_msg="user %d (%d ratings)" % (j, len(items[j])))
    # a user without ratings keeps the drawn vector (and a zero bias)
    np.testing.assert_array_equal(got[0][-k:], init[0])
    # the model itself is untouched
    g = gm.get_model()
    np.testing.assert_allclose(g["V"], om.item_factors, rtol=5e-5, atol=5e-5)


@pytest.mark.parametrize("biased", [True, False])
def test_score_items_matches_oracle(eng, biased):
    k = 24
    (u, i, v), om, gm, r = trained_pair(eng, biased, k, num_iter=5)
    rng = np.random.default_rng(8)
    items, values, init = fold_in_cases(rng, int(i.max()) + 1, k, 6)
    vec = gm.fold_in(items, values, init, 5)
    n_items = int(i.max()) + 1
    cand = rng.permutation(n_items).astype(np.int32)[:50]
    if biased:
        cand = np.concatenate([cand, np.array([n_items, n_items + 7], np.int32)])   # unknown items: bias terms only
    got = gm.score_items(vec, cand)
    for j in range(6):
        ov = om.fold_in(items[j], values[j], init[j])
        want = np.array([om.predict_vector(ov, c) for c in cand], np.float32)
        np.testing.assert_allclose(got[j], want, rtol=TOL, atol=TOL)
    if not biased:
        from mymedialite_b200._capi import MmlError
        with pytest.raises(MmlError):
            gm.score_items(vec, np.array([n_items], np.int32))     # RowScalarProduct throws "i too big" in the reference
        assert got.min() >= v.min() and got.max() <= v.max()        # Predict(vector, item) is clipped to the scale


def test_fold_in_rejects_unknown_items(eng):
    from mymedialite_b200._capi import MmlError
    (u, i, v), om, gm, r = trained_pair(eng, True, 8)
    with pytest.raises(MmlError):
        gm.fold_in([[int(i.max()) + 1]], [[3.0]], np.zeros((1, 8), np.float32), 1)
    assert gm.fold_in([], [], np.zeros((0, 8), np.float32), 3).shape == (0, 9)


@pytest.mark.parametrize("biased", [True, False])
def test_retrain_user_and_item_match_oracle(eng, biased):
    """RetrainUser / RetrainItem (MatrixFactorization.cs:141-160): row re-drawn by the host RNG, bias zeroed, then
    LearnFactors (:198-202) = NumIter passes over ByUser / ByItem updating that side only (plain MF decays the learn rate
    after every pass, :195); RemoveUser zeroes the row."""
    k = 16
    num_iter = 7
    (u, i, v), om, gm, r = trained_pair(eng, biased, k)
    rng = np.random.default_rng(4)
    for ent, by_item in ((5, False), (17, True), (0, False)):
        row = (rng.standard_normal(k) * 0.1).astype(np.float32)
        if by_item:
            om.item_factors[ent] = row
            if biased:
                om.item_bias[ent] = 0
            idx = np.nonzero(i == ent)[0].astype(np.int32)
        else:
            om.user_factors[ent] = row
            if biased:
                om.user_bias[ent] = 0
            idx = np.nonzero(u == ent)[0].astype(np.int32)
        gm.set_rows([ent], row, [0.0] if biased else None, by_item=by_item)
        for _ in range(num_iter):
            om.iterate_indices(idx, update_user=not by_item, update_item=by_item)
        gm.learn_factors(idx, num_iter, update_user=not by_item, update_item=by_item)
    g = gm.get_model()
    np.testing.assert_allclose(g["U"], om.user_factors, rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(g["V"], om.item_factors, rtol=5e-5, atol=5e-5)
    if biased:
        np.testing.assert_allclose(g["bu"], om.user_bias, rtol=5e-5, atol=5e-5)
        np.testing.assert_allclose(g["bi"], om.item_bias, rtol=5e-5, atol=5e-5)
    assert abs(gm.learnrate - om.learnrate) < 1e-9
    gm.set_rows([3], np.zeros(k, np.float32), [0.0])
    g = gm.get_model()
    assert not g["U"][3].any() and g["bu"][3] == 0


@pytest.mark.parametrize("cls", ["MatrixFactorization", "BiasedMatrixFactorization"])
def test_recommend_items_like_the_reference_test(cls):
    """FoldInRatingPredictorExtensionsTest.TestTopNWithCandidates / TestTopNWithoutCandidates (NumFactors 4, NumIter 5)
    on synthetic data of the ml-100k shape: 3 results, scores descending, items taken from the candidates."""
    from mymedialite_b200 import recommenders as R, synthetic, sysrandom
    sysrandom.seed(7)
    d = synthetic.ratings(943, 1682, 100000, "int", 31)
    rec = getattr(R, cls)()
    rec.Ratings = R.Ratings(*d["train"])
    rec.NumFactors, rec.NumIter = 4, 5
    rec.Train()
    rated = [(1, 1.0), (2, 4.0), (3, 4.5)]
    cand = [4, 5, 6, 7, 8]
    res = rec.RecommendItems(rated, cand, 3)
    assert len(res) == 3 and res[0][1] >= res[1][1] >= res[2][1] and all(it in cand for it, _ in res)
    res = rec.RecommendItems(rated, None, 3)
    assert len(res) == 3 and res[0][1] >= res[1][1] >= res[2][1]
    assert len(rec.ScoreItems(rated)) == rec.MaxItemID - 1
    # RetrainUser / RetrainItem keep the model usable and change only what they should
    before = rec._model.get_model()
    rec.RetrainUser(10)
    rec.RetrainItem(20)
    after = rec._model.get_model()
    changed_u = np.nonzero((before["U"] != after["U"]).any(axis=1))[0]
    changed_i = np.nonzero((before["V"] != after["V"]).any(axis=1))[0]
    assert changed_u.tolist() == [10] and changed_i.tolist() == [20]
    assert np.isfinite(rec.Predict(10, 20))
