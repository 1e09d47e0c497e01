"""Eval.Items.Evaluate on the device (mml_items_evaluate_mf / mml_wrmf_evaluate) against the oracle's restatement of
Eval/Items.cs:126-209 + Eval/Measures/*.cs driven by the oracle's Recommend(), and against the reference's own known
answer (src/Tests/Eval/ItemsTest.cs:34-77)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = 2e-6     # measures are ratios of small integers and sums of 1/log2 terms computed in double, cast to float


@pytest.fixture(scope="module")
def eng():
    from mymedialite_b200 import engine
    ctx = engine.Context(0)
    yield engine, ctx
    ctx.close()


def oracle_recommender(U, V):
    def recommend(user, n, ignore, candidates):
        items, scores = O.recommend_mf(U, V, user, n, candidates, sorted(ignore))
        return list(zip(items.tolist(), scores.tolist()))
    return recommend


def random_case(seed, n_users=60, n_items=200, k=12, quantize=False):
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((n_users, k)) * 0.3).astype(np.float32)
    V = (rng.standard_normal((n_items, k)) * 0.3).astype(np.float32)
    if quantize:                                   # few distinct scores -> many ties, resolved by candidate position
        U = np.round(U * 2).astype(np.float32); V = np.round(V * 2).astype(np.float32)
    n_train, n_test = 1500, 500
    tr_u = rng.integers(0, n_users, n_train); tr_i = rng.integers(0, n_items, n_train)
    te_u = rng.integers(0, n_users + 3, n_test); te_i = rng.integers(0, n_items + 5, n_test)   # ids outside the model too
    keep = ~np.isin(te_u * 100000 + te_i, tr_u * 100000 + tr_i)   # the reference requires test and training to be disjoint
    return U, V, tr_u, tr_i, te_u[keep], te_i[keep], rng


def rows_for(users, items, wanted):
    return [sorted(set(items[users == u].tolist())) for u in wanted]


@pytest.mark.parametrize("seed,quantize,n,repeated", [(1, False, -1, False), (2, True, -1, False), (3, False, 5, False),
                                                      (4, True, 20, False), (5, False, -1, True), (6, True, 3, True)])
def test_items_evaluate_matches_oracle(eng, seed, quantize, n, repeated):
    engine, ctx = eng
    U, V, tr_u, tr_i, te_u, te_i, rng = random_case(seed, quantize=quantize)
    cand = rng.permutation(V.shape[0] + 5).astype(np.int32)[:150]          # shuffled subset, some ids outside the model
    test_users = np.array(O.first_seen(te_u) + [57, 58], np.int32)          # plus users that may have no test items
    test_users = np.array(O.first_seen(test_users), np.int32)
    want, want_rows = O.items_evaluate(oracle_recommender(U, V), te_u, te_i, tr_u, tr_i, test_users=test_users.tolist(),
                                       candidate_items=cand.tolist(), repeated_events=repeated, n=n)
    rows, used = engine.items_evaluate_mf(ctx, U, V, test_users, cand, rows_for(te_u, te_i, test_users),
                                          None if repeated else rows_for(tr_u, tr_i, test_users), n)
    assert int((used == 1).sum()) == want["num_users"] and want["num_users"] > 20
    for b, u in enumerate(test_users.tolist()):
        assert (used[b] == 1) == (u in want_rows), u
        if used[b] == 1:
            np.testing.assert_allclose(rows[b], want_rows[u], rtol=TOL, atol=TOL, err_msg="user %d" % u)
    acc = np.zeros(8, np.float32)
    for b in np.nonzero(used == 1)[0]:
        acc = (acc + rows[b]).astype(np.float32)
    got = acc / np.float32(want["num_users"])
    np.testing.assert_allclose(got, [want[m] for m in O.ITEM_MEASURES], rtol=1e-5, atol=1e-6)


def test_reference_items_test_known_answer(eng):
    """ItemsTest.TestEvalDefault with MostPopular expressed as a rank-1 model (score = number of training events):
    one evaluated list, AUC 0.5, prec@5 0."""
    engine, ctx = eng
    train_u, train_i = np.array([1, 1, 2, 2, 3, 3]), np.array([1, 2, 2, 3, 1, 2])
    test_u, test_i = np.array([2, 2, 4]), np.array([3, 4, 4])
    counts = np.bincount(train_i, minlength=4).astype(np.float32)
    U = np.ones((5, 1), np.float32); V = counts.reshape(-1, 1)
    cand = np.array([3], np.int32)                                          # OVERLAP of test and training items
    test_users = np.array([2, 4], np.int32)
    rows, used = engine.items_evaluate_mf(ctx, U, V, test_users, cand, rows_for(test_u, test_i, test_users),
                                          rows_for(train_u, train_i, test_users), -1)
    assert used.tolist() == [1, 0] and rows[0, 0] == 0.5 and rows[0, 4] == 0.0
    # TestEvalDefaultGivenUserAndItems: all users 1..4
    test_users = np.array([1, 2, 3, 4], np.int32)
    rows, used = engine.items_evaluate_mf(ctx, U, V, test_users, cand, rows_for(test_u, test_i, test_users),
                                          rows_for(train_u, train_i, test_users), -1)
    assert used.tolist() == [0, 1, 0, 0] and rows[1, 0] == 0.5


def test_skip_rules_and_overlap_error(eng):
    engine, ctx = eng
    from mymedialite_b200._capi import MmlError
    U = np.ones((3, 2), np.float32); V = np.arange(12, dtype=np.float32).reshape(6, 2)
    cand = np.arange(6, dtype=np.int32)
    # user 0: no test item among the candidates; user 1: every non-ignored candidate is a test item; user 2: regular
    test_rows = [[9], [0, 1, 2], [5]]
    ignore_rows = [[], [3, 4, 5], [0]]
    rows, used = engine.items_evaluate_mf(ctx, U, V, [0, 1, 2], cand, test_rows, ignore_rows, -1)
    assert used.tolist() == [0, 0, 1]
    assert rows[2].tolist() == [1.0, 1.0, 1.0, 1.0, np.float32(1 / 5), np.float32(1 / 10), 1.0, 1.0]
    # a test item that is also a training item of the user: AUC.Compute throws "Should not happen."
    with pytest.raises(MmlError):
        engine.items_evaluate_mf(ctx, U, V, [2], cand, [[5, 0]], [[0]], -1)
    with pytest.raises(MmlError):
        engine.items_evaluate_mf(ctx, U, V, [2], np.array([1, 1], np.int32), [[1]], None, -1)   # duplicate candidate


def test_host_evaluate_on_a_trained_wrmf():
    """Items.Evaluate through the host mirror on a trained WRMF (model resident on the device) vs the oracle driven by
    the same factors; candidate list = OVERLAP + the reference's shuffle."""
    from mymedialite_b200 import evalitems, recommenders as R, synthetic, sysrandom
    sysrandom.seed(3)
    u, i = synthetic.implicit(400, 150, 12000, seed=8)
    rng = np.random.default_rng(0)
    is_test = rng.random(u.size) < 0.2
    rec = R.WRMF()
    rec.NumFactors, rec.NumIter = 16, 3
    rec.Feedback = R.PosOnlyFeedback(u[~is_test], i[~is_test])
    rec.Train()
    test = R.PosOnlyFeedback(u[is_test], i[is_test])
    sysrandom.seed(11)
    res = evalitems.Evaluate(rec, test, rec.Feedback, n=-1)
    sysrandom.seed(11)
    cand = evalitems.Candidates(None, evalitems.OVERLAP, test, rec.Feedback)
    Uh, Vh = rec._model.get_model()
    want, _ = O.items_evaluate(oracle_recommender(Uh, Vh), test.Users, test.Items, rec.Feedback.Users, rec.Feedback.Items,
                               candidate_items=cand.tolist())
    assert res["num_users"] == want["num_users"] and res["num_items"] == cand.size
    for m in O.ITEM_MEASURES:
        assert abs(res[m] - want[m]) < 1e-5, (m, res[m], want[m])
    assert 0.5 < res["AUC"] <= 1.0
    fit = evalitems.ComputeFit(rec)
    assert 0.5 < fit <= 1.0
