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
    for _ in range(30):                # Train() minus InitModel: RandomIndex is drawn on the first Iterate
        om.iterate(rng)
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


@pytest.mark.parametrize("intra", ["rounds", "async1"])
@pytest.mark.parametrize("k,biased,G,W,persistent,loss,freq,n_items", CASES)
def test_dsgd_epoch_equals_oracle_replay(eng, k, biased, G, W, persistent, loss, freq, n_items, intra):
    """Deterministic schedules: `rounds` (conflict-free parallel rounds) and `async` restricted to one worker per
    block. Both equal a serial pass in the dumped order, replayed by the oracle."""
    engine, ctx = eng
    mode = dict(intra_block=engine._capi.INTRA_ROUNDS) if intra == "rounds" else \
        dict(intra_block=engine._capi.INTRA_ASYNC, async_workers=1)
    d = small_data(n_items=n_items, n=20000 if n_items < 1000 else 40000)
    u, i, v = d["train"]
    kw = dict(loss=loss, frequency_regularization=freq) if biased else {}
    om = oracle_model(u, i, v, biased, k, **kw)
    kw.update(mode)
    r, gm = gpu_model(eng, u, i, v, biased, k, om, num_groups=G, num_subgroups=W, persistent=persistent,
                      hot_item_factor=0.0, **kw)
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
    # async: the step is applied as an fp32 delta by the L2 atomic unit -> one more rounding per update
    assert_model_close(gm, om, biased, 5e-5 if intra == "rounds" else 3e-4)


@pytest.mark.parametrize("k,G,cpg,W,persistent", [(64, 3, 2, 4, 1), (128, 4, 3, 2, 1), (32, 2, 4, 2, 0), (128, 1, 6, 4, 1)])
def test_dsgd_multi_cta_groups_equal_oracle_replay(eng, k, G, cpg, W, persistent):
    """A worker group made of several CTAs (ctas_per_group): with one active worker per group the epoch is
    still a serial pass in the dumped order; the hand-over waits on every CTA of the previous holder."""
    engine, ctx = eng
    d = small_data()
    u, i, v = d["train"]
    om = oracle_model(u, i, v, True, k)
    r, gm = gpu_model(eng, u, i, v, True, k, om, num_groups=G, ctas_per_group=cpg, num_subgroups=W, persistent=persistent,
                      intra_block=engine._capi.INTRA_ASYNC, async_workers=1)
    assert gm.strata_info()["G"] == G
    rs = np.random.RandomState(4)
    for epoch in range(2):
        seq = rs.permutation(G).astype(np.int32)
        order = gm.schedule(seq)
        assert np.array_equal(np.sort(order), np.arange(u.size))
        gm.iterate(subepoch_sequence=seq)
        om.iterate_indices(order)
    assert_model_close(gm, om, True, 3e-4)


def _emulate_epoch(model, u, i, v, order, block, copy, W, hp, average=False):
    """fp32 numpy restatement of the kernel's semantics INCLUDING hot-item copies: inside a block a hot item's
    entries update private copies (all starting from the row at block start); at block end
    row = copy0 + sum_{c>=1} (copy_c - old). Everything else is the reference update (:264-310)."""
    U, V, bu, bi = model["U"], model["V"], model["bu"], model["bi"]
    f = np.float32
    lr, gb, minr, rng = f(hp["lr"]), f(hp["gb"]), f(hp["minr"]), f(hp["range"])
    reg, blr, breg = f(0.015), f(1.0), f(0.01)
    pos_sorted = np.argsort(block, kind="stable")
    bounds = np.flatnonzero(np.diff(block[pos_sorted])) + 1
    for chunk in np.split(pos_sorted, bounds):
        hot = {}
        for pos in chunk:
            t = order[pos]; uu, ii, c = u[t], i[t], copy[pos]
            if c >= 0:
                if ii not in hot:
                    hot[ii] = ([V[ii].copy() for _ in range(W)], [f(bi[ii]) for _ in range(W)], V[ii].copy(), f(bi[ii]))
                q, qb = hot[ii][0][c], hot[ii][1][c]
            else:
                q, qb = V[ii], bi[ii]
            p = U[uu]
            score = f(f(f(gb + bu[uu]) + qb) + f(np.dot(p, q)))
            sig = f(1.0) / (f(1.0) + np.exp(-score, dtype=f))
            err = f(v[t]) - (minr + sig * rng)
            gc = f(err * sig * (f(1.0) - sig) * rng)
            bu[uu] = bu[uu] + blr * lr * (gc - breg * reg * bu[uu])
            qb_new = qb + blr * lr * (gc - breg * reg * qb)
            p_new = p + lr * (gc * q - reg * p)
            q_new = q + lr * (gc * p - reg * q)
            U[uu] = p_new
            if c >= 0:
                hot[ii][0][c][:] = q_new; hot[ii][1][c] = f(qb_new)
            else:
                V[ii] = q_new; bi[ii] = qb_new
        for ii, (rows, biases, old, old_b) in hot.items():
            scale = f(1.0 / W) if average else f(1.0)
            acc, accb = rows[0] - old, f(biases[0] - old_b)
            for c in range(1, W):
                acc += rows[c] - old; accb = f(accb + f(biases[c] - old_b))
            V[ii] = old + scale * acc; bi[ii] = f(old_b + scale * accb)


@pytest.mark.parametrize("k,G,W,persistent,hot,avg", [(64, 4, 4, 1, 0.5, 1), (128, 3, 8, 0, 0.5, 1), (10, 5, 2, 1, 0.1, 1)])
def test_dsgd_hot_item_copies_match_emulation(eng, k, G, W, persistent, hot, avg):
    """Hot items (whose updates would serialise a block) run as W private chains per block, merged by summing
    deltas; the device result equals an fp32 emulation of exactly that rule."""
    engine, ctx = eng
    d = small_data(n_users=400, n_items=60, n=12000)
    u, i, v = d["train"]
    om = oracle_model(u, i, v, True, k)
    r, gm = gpu_model(eng, u, i, v, True, k, om, num_groups=G, num_subgroups=W, persistent=persistent, hot_item_factor=hot,
                      hot_copies=W, intra_block=engine._capi.INTRA_ROUNDS, hot_merge_average=avg)
    assert gm.hot_items() > 0
    state = dict(U=om.user_factors.copy(), V=om.item_factors.copy(), bu=om.user_bias.copy(), bi=om.item_bias.copy())
    g0 = gm.get_model(False, False)
    avg, mn, mx = r.stats()
    hp = dict(lr=0.01, gb=g0["global_bias"], minr=mn, range=mx - mn)
    order, block, copy = gm.schedule(detail=True)
    assert (copy >= 0).sum() > 0 and copy.max() < W
    gm.iterate()
    _emulate_epoch(state, u, i, v, order, block, copy, W, hp, average=bool(avg))
    g = gm.get_model()
    for key in ("U", "V", "bu", "bi"):
        np.testing.assert_allclose(g[key], state[key], rtol=1e-4, atol=1e-4, err_msg=key)


def test_dsgd_reference_group_rule_and_schedule_is_stratified(eng):
    """PERM_MOD rule: group = perm[id] % G (MultiCore.cs:64); the blocks of a sub-epoch share no user and no item,
    and the dumped order walks sub-epochs in sequence, the reference's DSGD schedule with g = G."""
    engine, ctx = eng
    d = small_data()
    u, i, v = d["train"]
    G = 6
    rng = O.Random(7)
    up = rng.shuffle(np.arange(u.max() + 1)); ip = rng.shuffle(np.arange(i.max() + 1))
    r = engine.DeviceRatings(ctx, u, i, v)
    p = engine.default_params(num_factors=16, num_groups=G, num_subgroups=2, group_rule=engine._capi.GROUPS_PERM_MOD,
                              hot_item_factor=0.0, intra_block=engine._capi.INTRA_ROUNDS)
    gm = engine.SgdModel(ctx, r, p, up, ip)
    seq = np.array([3, 0, 5, 1, 4, 2], np.int32)
    order, block, copy = gm.schedule(seq, detail=True)
    assert np.all(copy == -1)
    j, b = up[u[order]] % G, ip[i[order]] % G
    slot = (b - j) % G
    t = block // G
    assert np.array_equal(slot, seq[t]) and np.array_equal(block % G, j)
    assert np.all(np.diff(t) >= 0)
    # the reference's partition: same membership as MultiCore.PartitionUsersAndItems with these permutations
    optr, oidx = O.partition_blocks_given(u, i, up, ip, G)
    for jj in range(G):
        for bb in range(G):
            mine = np.sort(order[(j == jj) & (b == bb)])
            assert np.array_equal(mine, oidx[optr[jj * G + bb]:optr[jj * G + bb + 1]])


def test_dsgd_rounds_are_matchings(eng):
    """Inside a block the CTA runs rounds of ratings with pairwise distinct users and item rows."""
    engine, ctx = eng
    d = small_data(n_users=200, n_items=50, n=6000)
    u, i, v = d["train"]
    r = engine.DeviceRatings(ctx, u, i, v)
    gm = engine.SgdModel(ctx, r, engine.default_params(num_factors=8, num_groups=3, num_subgroups=2, hot_item_factor=1.0,
                                                       intra_block=engine._capi.INTRA_ROUNDS))
    order, block, copy = gm.schedule(detail=True)
    rounds = gm.round_sizes()
    assert rounds.sum() == u.size
    pos = 0
    for sz in rounds:
        sel = order[pos:pos + sz]
        assert len(set(u[sel])) == sz, "a user appears twice in a round"
        rows = list(zip(i[sel], copy[pos:pos + sz]))
        assert len(set(rows)) == sz, "an item row appears twice in a round"
        assert len(set(block[pos:pos + sz])) == 1
        pos += sz


@pytest.mark.parametrize("intra", ["async", "rounds", "async_4x4", "async_1x16"])
def test_dsgd_rmse_tracks_single_threaded_oracle(eng, intra):
    """north_star gate: per-epoch train/test RMSE within 0.5 % of the reference's own (MaxThreads=1) run, for the
    default lock-free intra-block mode and for the conflict-free rounds."""
    engine, ctx = eng
    from mymedialite_b200 import synthetic
    d = synthetic.ratings(3000, 800, 300000, "half", 11)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    k = 32
    rng = O.Random(1)
    om = O.Model(u, i, v, biased=True, num_factors=k)
    om.init(rng)
    shape = dict(num_groups=16)
    if "x" in intra:     # worker groups of several CTAs; 1 x 16 = the whole grid is one lock-free group
        g, c = intra.split("_")[1].split("x")
        shape = dict(num_groups=int(g), ctas_per_group=int(c))
    r, gm = gpu_model(eng, u, i, v, True, k, om, num_subgroups=4, hot_item_factor=0.0,
                      intra_block=engine._capi.INTRA_ROUNDS if intra == "rounds" else engine._capi.INTRA_ASYNC, **shape)
    for epoch in range(8):
        om.iterate(rng)
        gm.iterate()
        o_tr, o_te = om.evaluate(u, i, v)["RMSE"], om.evaluate(tu, ti, tv)["RMSE"]
        g_tr, g_te = gm.evaluate_train()["RMSE"], gm.evaluate(tu, ti, tv)["RMSE"]
        assert abs(g_tr - o_tr) / o_tr < 0.005, (epoch, g_tr, o_tr)
        assert abs(g_te - o_te) / o_te < 0.005, (epoch, g_te, o_te)


@pytest.mark.parametrize("biased,k", [(True, 32), (True, 128), (False, 10)])
def test_naive_parallelization_tracks_single_threaded_oracle(eng, biased, k):
    """NaiveParallelization (BiasedMatrixFactorization.cs:136-141, :201-204; lists of MultiCore.PartitionIndices, one per
    worker of the GPU, walked with no exclusivity): declared non-reproducible by the reference, so the gate is the statistical
    one -- per-epoch train / test RMSE within 0.5 % of the single-threaded run from the same factors."""
    engine, ctx = eng
    from mymedialite_b200 import synthetic
    d = synthetic.ratings(3000, 800, 300000, "half", 11)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    rng = O.Random(1)
    om = O.Model(u, i, v, biased=biased, num_factors=k)
    om.init(rng)
    r, gm = gpu_model(eng, u, i, v, biased, k, om, schedule=engine._capi.SCHEDULE_NAIVE, max_threads=8)
    ri = O.Random(5).shuffle(np.arange(u.size))
    lr0 = gm.learnrate
    for epoch in range(6):
        om.iterate(rng)
        gm.iterate(random_index=ri)
        o_tr, o_te = om.evaluate(u, i, v)["RMSE"], om.evaluate(tu, ti, tv)["RMSE"]
        g_tr, g_te = gm.evaluate_train()["RMSE"], gm.evaluate(tu, ti, tv)["RMSE"]
        assert abs(g_tr - o_tr) / o_tr < 0.005, (epoch, g_tr, o_tr)
        assert abs(g_te - o_te) / o_te < 0.005, (epoch, g_te, o_te)
    assert gm.learnrate == lr0                      # Decay = 1: UpdateLearnRate (twice, MaxThreads > 1) leaves it alone


@pytest.mark.parametrize("env,k,biased", [({"MMLB200_SGD_OWNED": "1"}, 128, True), ({"MMLB200_SGD_OWNED": "0"}, 32, True),
                                          ({"MMLB200_SGD_OWNED": "1"}, 10, False), ({"MMLB200_SGD_VARIANT": "2"}, 128, True),
                                          ({"MMLB200_SGD_VARIANT": "2"}, 64, False)])
def test_async_loop_forms_track_single_threaded_oracle(eng, monkeypatch, env, k, biased):
    """The forms of the lock-free block schedule the library does not pick by default for this row length -- the owned-users
    loop (users pinned to workers, no block hand-overs) forced on at k = 128 and off at k = 32, and the bulk-reduction form of
    the item step -- against the same gate as the default: every rating visited exactly once, per-epoch train / test RMSE
    within 0.5 % of the oracle's single-threaded run from the same factors."""
    engine, ctx = eng
    from mymedialite_b200 import synthetic
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    d = synthetic.ratings(3000, 800, 300000, "half", 13)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    rng = O.Random(1)
    om = O.Model(u, i, v, biased=biased, num_factors=k)
    om.init(rng)
    r, gm = gpu_model(eng, u, i, v, biased, k, om, num_groups=6, ctas_per_group=3, num_subgroups=4)
    assert np.array_equal(np.sort(gm.schedule()), np.arange(u.size))
    for epoch in range(6):
        om.iterate(rng)
        gm.iterate()
        o_tr, o_te = om.evaluate(u, i, v)["RMSE"], om.evaluate(tu, ti, tv)["RMSE"]
        g_tr, g_te = gm.evaluate_train()["RMSE"], gm.evaluate(tu, ti, tv)["RMSE"]
        assert abs(g_tr - o_tr) / o_tr < 0.005, (epoch, g_tr, o_tr)
        assert abs(g_te - o_te) / o_te < 0.005, (epoch, g_te, o_te)


def test_l2_probe_reports_rates(eng):
    """mml_ctx_probe_l2 (the denominator of bench.py's roofline.binding.l2): positive, and a row atomic is dearer than a row read."""
    engine, ctx = eng
    rd = ctx.probe_l2(0, 4096, 128, 2)
    red = ctx.probe_l2(1, 4096, 128, 2)
    both = ctx.probe_l2(2, 4096, 128, 2)
    assert rd > 1e9 and red > 1e9 and both > 1e9
    assert red < rd and both < rd
    with pytest.raises(Exception):
        ctx.probe_l2(3, 4096, 128, 2)
    with pytest.raises(Exception):
        ctx.probe_l2(0, 4096, 100, 2)


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
