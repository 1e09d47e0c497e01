"""CUDA SGD path (through the C ABI) against the CPU oracle.

Serial schedule: the reference's order and mixed precision -> factors equal to fp32 rounding.
DSGD schedule: the parallel epoch is conflict-free, hence equal to a serial pass in the order
mml_sgd_schedule_dump reports; the oracle replays that order (tolerance = fp32 dot/update rounding).
Statistical gate (north_star): per-epoch RMSE within 0.5 % of the oracle's single-threaded run."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from mymedialite_b200 import engine
    ctx = engine.Context(0)
    yield engine, ctx
    ctx.close()


def small_data(n_users=300, n_items=120, n=20000, seed=5):
    from mymedialite_b200 import synthetic
    d = synthetic.ratings(n_users, n_items, n, "half", seed)
    return d


def oracle_model(u, i, v, biased, k, rng_seed=1, **kw):
    m = O.Model(u, i, v, biased=biased, num_factors=k, **kw)
    m.init(O.Random(rng_seed))
    return m


def gpu_model(eng, u, i, v, biased, k, om, **kw):
    engine, ctx = eng
    r = engine.DeviceRatings(ctx, u, i, v)
    p = engine.default_params(biased=int(biased), num_factors=k, **kw)
    m = engine.SgdModel(ctx, r, p)
    m.set_model(om.user_factors.copy(), om.item_factors.copy())
    return r, m


def assert_model_close(gm, om, biased, tol):
    g = gm.get_model()
    np.testing.assert_allclose(g["U"], om.user_factors, rtol=tol, atol=tol)
    np.testing.assert_allclose(g["V"], om.item_factors, rtol=tol, atol=tol)
    if biased:
        np.testing.assert_allclose(g["bu"], om.user_bias, rtol=tol, atol=tol)
        np.testing.assert_allclose(g["bi"], om.item_bias, rtol=tol, atol=tol)


@pytest.mark.parametrize("biased", [True, False])
def test_serial_example_train_matches_oracle(eng, example_data, biased):
    """Config 1: tests/example.train, k=10, 30 epochs, MaxThreads=1 order."""
    engine, ctx = eng
    u, i, v, tu, ti, tv = example_data
    rng = O.Random(1)
    om = O.Model(u, i, v, biased=biased, num_factors=10, num_iter=30)
    om.init(rng)
    r, gm = gpu_model(eng, u, i, v, biased, 10, om, schedule=engine._capi.SCHEDULE_SERIAL)
    assert abs(gm.get_model(False, False)["global_bias"] - om.global_bias) < 1e-6
    om.train(rng)                      # draws RandomIndex on the first Iterate, 30 epochs
    ri = om.random_index.copy()
    for _ in range(30):
        gm.iterate(random_index=ri)
    assert_model_close(gm, om, biased, 2e-5)
    np.testing.assert_allclose(gm.predict(tu, ti), om.predict_many(tu, ti), rtol=1e-5, atol=1e-5)
    ge, oe = gm.evaluate(tu, ti, tv), om.evaluate(tu, ti, tv)
    for key in ("RMSE", "MAE", "NMAE", "CBD"):
        assert abs(ge[key] - oe[key]) <= 1e-5 * max(1.0, abs(oe[key])), key
    assert gm.learnrate == pytest.approx(om.learnrate)


def test_learnrate_schedule_known_answers(eng):
    """src/Tests/RatingPrediction/BiasedMatrixFactorizationTest.cs:30-62 on the CUDA class."""
    engine, ctx = eng
    u = np.array([0, 1, 2, 3], np.int32); i = np.array([0, 1, 2, 3], np.int32); v = np.array([1, 2, 3, 4], np.float32)
    r = engine.DeviceRatings(ctx, u, i, v)
    for max_threads, after_one in ((1, 0.5), (2, 0.25)):
        p = engine.default_params(num_factors=4, learn_rate=1.0, decay=0.5, max_threads=max_threads,
                                  schedule=engine._capi.SCHEDULE_SERIAL)
        m = engine.SgdModel(ctx, r, p)
        m.set_model(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32))
        assert m.learnrate == 1.0
        m.iterate(random_index=np.arange(4, dtype=np.int32))
        assert m.learnrate == after_one
    p = engine.default_params(num_factors=4, learn_rate=1.1, schedule=engine._capi.SCHEDULE_SERIAL)
    m = engine.SgdModel(ctx, r, p)
    m.set_model(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32))
    for _ in range(10):
        m.iterate(random_index=np.arange(4, dtype=np.int32))
    assert m.learnrate == np.float32(1.1)


CASES = [
    # k, biased, G, W, persistent, loss, freq_reg, n_items
    (10, True, 4, 4, 0, 0, 0, 120),
    (64, True, 4, 4, 1, 0, 0, 120),
    (128, True, 6, 2, 1, 0, 1, 120),
    (40, False, 4, 4, 1, 0, 0, 120),
    (64, True, 3, 8, 0, 1, 0, 120),
    (32, True, 5, 1, 1, 2, 0, 120),
    (200, True, 2, 2, 1, 0, 0, 120),
    (128, True, 2, 4, -1, 0, 0, 1500),    # item groups too large for shared memory -> global item rows
]


@pytest.mark.parametrize("k,biased,G,W,persistent,loss,freq,n_items", CASES)
def test_dsgd_epoch_equals_oracle_replay(eng, k, biased, G, W, persistent, loss, freq, n_items):
    engine, ctx = eng
    d = small_data(n_items=n_items, n=20000 if n_items < 1000 else 40000)
    u, i, v = d["train"]
    kw = dict(loss=loss, frequency_regularization=freq) if biased else {}
    om = oracle_model(u, i, v, biased, k, **kw)
    r, gm = gpu_model(eng, u, i, v, biased, k, om, num_groups=G, num_subgroups=W, persistent=persistent, **kw)
    info = gm.strata_info()
    assert info["G"] == G and info["W"] == W
    if n_items >= 1000:
        assert info["staged_bytes"] == 0
    rs = np.random.RandomState(3)
    for epoch in range(2):
        seq = rs.permutation(G).astype(np.int32)
        order = gm.schedule(seq)
        assert np.array_equal(np.sort(order), np.arange(u.size))
        gm.iterate(subepoch_sequence=seq)
        om.iterate_indices(order)
    assert_model_close(gm, om, biased, 5e-5)


def test_dsgd_reference_group_rule_and_schedule_is_stratified(eng):
    """PERM_MOD rule: group = perm[id] % groups (MultiCore.cs:64); concurrent sub-blocks share no user and no item."""
    engine, ctx = eng
    d = small_data()
    u, i, v = d["train"]
    G, W = 4, 2
    rng = O.Random(7)
    up = rng.shuffle(np.arange(u.max() + 1)); ip = rng.shuffle(np.arange(i.max() + 1))
    r = engine.DeviceRatings(ctx, u, i, v)
    p = engine.default_params(num_factors=16, num_groups=G, num_subgroups=W, group_rule=engine._capi.GROUPS_PERM_MOD)
    gm = engine.SgdModel(ctx, r, p, up, ip)
    order = gm.schedule()
    T = G * W
    ug, ig = up[u[order]] % T, ip[i[order]] % T
    j, w = ug % G, ug // G
    b, c = ig % G, ig // G
    slot, step = (b - j) % G, (c - w) % W
    key = slot * W + step
    assert np.all(np.diff(key) >= 0), "schedule is ordered by (slot, step)"
    for s in np.unique(key):
        sel = key == s
        workers = j[sel] * W + w[sel]
        # inside one (slot, step) every user and every item belongs to exactly one worker
        for ids in (u[order][sel], i[order][sel]):
            owner = {}
            for x, wk in zip(ids, workers):
                assert owner.setdefault(x, wk) == wk


def test_dsgd_rmse_tracks_single_threaded_oracle(eng):
    """north_star gate: per-epoch train/test RMSE within 0.5 % of the reference's own (MaxThreads=1) run."""
    engine, ctx = eng
    from mymedialite_b200 import synthetic
    d = synthetic.ratings(3000, 800, 300000, "half", 11)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    k = 32
    rng = O.Random(1)
    om = O.Model(u, i, v, biased=True, num_factors=k)
    om.init(rng)
    r, gm = gpu_model(eng, u, i, v, True, k, om, num_groups=16, num_subgroups=4)
    for epoch in range(8):
        om.iterate(rng)
        gm.iterate()
        o_tr, o_te = om.evaluate(u, i, v)["RMSE"], om.evaluate(tu, ti, tv)["RMSE"]
        g_tr, g_te = gm.evaluate_train()["RMSE"], gm.evaluate(tu, ti, tv)["RMSE"]
        assert abs(g_tr - o_tr) / o_tr < 0.005, (epoch, g_tr, o_tr)
        assert abs(g_te - o_te) / o_te < 0.005, (epoch, g_te, o_te)


def test_predict_evaluate_objective_match_oracle(eng):
    engine, ctx = eng
    d = small_data()
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    for biased, kw in ((True, dict(loss=0)), (True, dict(loss=2, frequency_regularization=1)), (False, {})):
        om = oracle_model(u, i, v, biased, 24, **kw)
        om.iterate_indices(np.arange(u.size, dtype=np.int32))
        r, gm = gpu_model(eng, u, i, v, biased, 24, om, num_groups=4, num_subgroups=2, **kw)
        if biased:
            gm_state = dict(bu=om.user_bias.copy(), bi=om.item_bias.copy())
            gm.set_model(om.user_factors.copy(), om.item_factors.copy(), gm_state["bu"], gm_state["bi"])
        # unknown ids are legal in Predict (BiasedMatrixFactorization.cs:313-325, MatrixFactorization.cs:251-259)
        pu = np.concatenate([tu, [u.max() + 5, 0]]).astype(np.int32)
        pi = np.concatenate([ti, [0, i.max() + 9]]).astype(np.int32)
        np.testing.assert_allclose(gm.predict(pu, pi), om.predict_many(pu, pi), rtol=2e-6, atol=2e-6)
        ge, oe = gm.evaluate(tu, ti, tv), om.evaluate(tu, ti, tv)
        for key in ("RMSE", "MAE", "NMAE", "CBD"):
            assert abs(ge[key] - oe[key]) <= 2e-6 * max(1.0, abs(oe[key])), (key, ge, oe)
        if biased:
            assert gm.objective() == pytest.approx(om.objective(), rel=1e-5)


def test_device_init_is_layout_independent_and_zeroes_empty_rows(eng):
    engine, ctx = eng
    u = np.array([0, 2, 2, 5], np.int32); i = np.array([1, 1, 3, 0], np.int32); v = np.array([1, 2, 3, 4], np.float32)
    r = engine.DeviceRatings(ctx, u, i, v, max_user=7, max_item=4)
    models = []
    for G in (1, 2):
        p = engine.default_params(num_factors=20, num_groups=G, num_subgroups=1)
        m = engine.SgdModel(ctx, r, p)
        m.init_model(1234, 0.0, 0.1)
        models.append(m.get_model())
    np.testing.assert_array_equal(models[0]["U"], models[1]["U"])
    np.testing.assert_array_equal(models[0]["V"], models[1]["V"])
    U, V = models[0]["U"], models[0]["V"]
    assert np.all(U[[1, 3, 4, 6, 7]] == 0) and np.all(V[[2, 4]] == 0)
    assert np.all(U[[0, 2, 5]] != 0) and abs(U[[0, 2, 5]].std() - 0.1) < 0.03


def test_errors_are_reported_not_swallowed(eng):
    engine, ctx = eng
    u = np.array([0, 1], np.int32); i = np.array([0, 1], np.int32); v = np.array([1, 2], np.float32)
    with pytest.raises(engine._capi.MmlError):
        engine.DeviceRatings(ctx, u, i, v, max_user=0, max_item=1)        # id out of range
    r = engine.DeviceRatings(ctx, u, i, v)
    with pytest.raises(engine._capi.MmlError):
        engine.SgdModel(ctx, r, engine.default_params(num_factors=1000))  # unsupported width
    m = engine.SgdModel(ctx, r, engine.default_params(num_factors=8))
    with pytest.raises(engine._capi.MmlError):
        m.iterate()                                                       # no model yet
