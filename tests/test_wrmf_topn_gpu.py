"""WRMF ALS and Recommend() on the device against the CPU oracle (ItemRecommendation/WRMF.cs, Recommender.cs)."""
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


def events(n_users, n_items, n, seed, dup=0.1):
    rs = np.random.RandomState(seed)
    u = rs.randint(0, n_users, n).astype(np.int32)
    i = (rs.zipf(1.5, n) % n_items).astype(np.int32)
    m = int(n * dup)                       # repeated events must collapse (SparseBooleanMatrix rows are sets)
    u = np.concatenate([u, u[:m]]); i = np.concatenate([i, i[:m]])
    p = rs.permutation(u.size)
    return u[p], i[p]


def test_feedback_known_answers(eng):
    """TestUtils.CreatePosOnlyFeedback (src/Tests/TestUtils.cs:58-72): {(0,0),(0,1),(1,0),(1,2)} + a duplicate."""
    engine, ctx = eng
    u = np.array([0, 0, 1, 1, 0], np.int32); i = np.array([0, 1, 0, 2, 1], np.int32)
    f = engine.DeviceFeedback(ctx, u, i, max_user=2, max_item=3)
    assert f.nnz == 4
    ptr, cols = f.csr(False)
    assert list(ptr) == [0, 2, 4, 4] and list(cols) == [0, 1, 0, 2]
    ptr, rows = f.csr(True)
    assert list(ptr) == [0, 2, 3, 4, 4] and list(rows) == [0, 1, 0, 1]


@pytest.mark.parametrize("nu,ni,n,seed", [(300, 200, 5000, 1), (5000, 900, 120000, 2)])
def test_feedback_csr_matches_oracle(eng, nu, ni, n, seed):
    engine, ctx = eng
    u, i = events(nu, ni, n, seed)
    f = engine.DeviceFeedback(ctx, u, i, max_user=nu, max_item=ni - 1)      # user nu has no events
    optr, ocols = O.feedback_csr(u, i, nu)
    ptr, cols = f.csr(False)
    assert f.nnz == ocols.size and np.array_equal(ptr, optr)
    for r in range(nu + 1):                                                  # rows are sets: compare as sets
        assert np.array_equal(np.sort(ocols[optr[r]:optr[r + 1]]), cols[ptr[r]:ptr[r + 1]])
    optr, ocols = O.feedback_csr(i, u, ni - 1)
    ptr, rows = f.csr(True)
    assert np.array_equal(ptr, optr)
    for r in range(ni):
        assert np.array_equal(np.sort(ocols[optr[r]:optr[r + 1]]), rows[ptr[r]:ptr[r + 1]])


@pytest.mark.parametrize("k,alpha,reg", [(10, 1.0, 0.015), (64, 2.0, 0.1), (128, 1.0, 0.015), (150, 0.5, 0.015)])
def test_wrmf_iterate_matches_oracle(eng, k, alpha, reg):
    """north_star gate: ALS factor rows within 1e-4 relative (fp32) of the reference arithmetic."""
    engine, ctx = eng
    nu, ni = 400, 150
    u, i = events(nu - 10, ni, 9000, k)                                  # the last 10 users have no events
    f = engine.DeviceFeedback(ctx, u, i, max_user=nu - 1, max_item=ni - 1)
    rng = O.Random(3)
    U = rng.init_normal(nu * k).reshape(nu, k); V = rng.init_normal(ni * k).reshape(ni, k)   # MF.cs:56-57 draw order
    m = engine.WrmfModel(ctx, f, k, alpha, reg)
    m.set_model(U, V)
    uptr, ucols = O.feedback_csr(u, i, nu - 1)
    iptr, irows = O.feedback_csr(i, u, ni - 1)
    Uo, Vo = U.copy(), V.copy()
    for _ in range(2):
        m.iterate()
        O.wrmf_optimize(uptr, ucols, Uo, Vo, alpha, reg)
        O.wrmf_optimize(iptr, irows, Vo, Uo, alpha, reg)
    Ug, Vg = m.get_model()
    for got, want in ((Ug, Uo), (Vg, Vo)):
        scale = np.abs(want).max(axis=1, keepdims=True) + 1e-12
        assert (np.abs(got - want) / scale).max() < 1e-4
    empty = np.flatnonzero(np.diff(uptr) == 0)
    assert empty.size > 0 and np.all(Ug[empty] == 0)                     # HCp = 0 => the row is exactly zero


@pytest.mark.parametrize("k", [64, 128, 20])
def test_tensor_core_gram_sum_is_fp32_accurate(eng, k):
    """The tcgen05 kernel's sum_{i in S_u} h_i h_i^T (3 x TF32 split) against a float64 sum: relative error ~1e-6,
    i.e. fp32 level, for the user with the longest item list."""
    engine, ctx = eng
    nu, ni = 300, 500
    rs = np.random.RandomState(k)
    u = rs.randint(0, nu, 20000).astype(np.int32); i = rs.randint(0, ni, 20000).astype(np.int32)     # ~65 distinct items per user
    f = engine.DeviceFeedback(ctx, u, i, max_user=nu - 1, max_item=ni - 1)
    U = (rs.randn(nu, k) * 0.1).astype(np.float32); V = (rs.randn(ni, k) * 0.3).astype(np.float32)
    m = engine.WrmfModel(ctx, f, k)
    m.set_model(U, V)
    user, G = m.debug_gram()
    ptr, cols = f.csr(False)
    items = cols[ptr[user]:ptr[user + 1]]
    assert items.size == np.diff(ptr).max() and items.size > 32          # spans several 32-row stages
    Vd = V[items].astype(np.float64)
    want = Vd.T @ Vd
    assert np.all(G[k:, :] == 0) and np.all(G[:, k:] == 0)
    err = np.abs(G[:k, :k] - want).max() / np.abs(want).max()
    assert err < 5e-6, err
    Ug, Vg = m.get_model()
    assert np.array_equal(Ug, U) and np.array_equal(Vg, V)              # the diagnostic leaves the model alone


@pytest.mark.parametrize("mode", ["fp64", "tensor", "tensor_f64", "tensor_pcg"])
def test_wrmf_paths_agree_on_a_larger_set(eng, mode):
    """The device paths (all-double CUDA cores; tcgen05 Gram sums with the single- or the double-precision Cholesky
    factor, or with the preconditioned-CG row solver) against the oracle on 3000 x 1200, k = 64 (rows up to ~1000
    entries, empty rows, 2 epochs)."""
    engine, ctx = eng
    nu, ni, k = 3000, 1200, 64
    u, i = events(nu - 5, ni, 90000, 77)
    f = engine.DeviceFeedback(ctx, u, i, max_user=nu - 1, max_item=ni - 1)
    rng = O.Random(5)
    U = rng.init_normal(nu * k).reshape(nu, k); V = rng.init_normal(ni * k).reshape(ni, k)
    engine.wrmf_set_mode(dict(fp64=engine._capi.WRMF_FP64, tensor=engine._capi.WRMF_TENSOR,
                              tensor_f64=engine._capi.WRMF_TENSOR_F64, tensor_pcg=engine._capi.WRMF_TENSOR_PCG)[mode])
    try:
        m = engine.WrmfModel(ctx, f, k)
        m.set_model(U, V)
        for _ in range(2):
            m.iterate()
        Ug, Vg = m.get_model()
    finally:
        engine.wrmf_set_mode(engine._capi.WRMF_AUTO)
    uptr, ucols = O.feedback_csr(u, i, nu - 1)
    iptr, irows = O.feedback_csr(i, u, ni - 1)
    Uo, Vo = U.copy(), V.copy()
    for _ in range(2):
        O.wrmf_optimize(uptr, ucols, Uo, Vo, 1.0, 0.015)
        O.wrmf_optimize(iptr, irows, Vo, Uo, 1.0, 0.015)
    for got, want in ((Ug, Uo), (Vg, Vo)):
        scale = np.abs(want).max(axis=1, keepdims=True) + 1e-12
        assert (np.abs(got - want) / scale).max() < 1e-4


def _oracle_lists(U, V, users, n, candidates, ignore_lists):
    out = []
    for b, usr in enumerate(users):
        out.append(O.recommend_mf(U, V, int(usr), n, candidates, None if ignore_lists is None else ignore_lists[b]))
    return out


@pytest.mark.parametrize("n", [-1, 1, 10, 32, 50])
@pytest.mark.parametrize("k", [10, 128])
def test_topn_is_bit_exact(eng, n, k):
    """Item indices, order (score desc, candidate position asc) and fp32 scores equal the reference's."""
    engine, ctx = eng
    rs = np.random.RandomState(n + k)
    nu, ni = 90, 333
    U = (rs.randn(nu, k) * 0.1).astype(np.float32); V = (rs.randn(ni, k) * 0.1).astype(np.float32)
    V[7] = V[3]; V[200] = V[3]                                          # exact score ties -> position decides
    U[5] = 0                                                              # a whole row of ties
    users = np.array([0, 5, 17, 89, 5, 120], np.int32)                    # repeated user, user outside the model
    cand = rs.permutation(ni + 4)[:300].astype(np.int32)                  # shuffled (Eval/Items.cs:94), some ids outside the model
    ign = [rs.choice(ni, 25, replace=False).astype(np.int32) for _ in users]
    got = engine.topn_mf(ctx, U, V, users, n, cand, ign)
    want = _oracle_lists(U, V, users, n, cand, ign)
    for (gi, gs), (wi, ws) in zip(got, want):
        assert np.array_equal(gi, wi)
        assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
    # default candidates = every item, no ignore list
    got = engine.topn_mf(ctx, U, V, users[:3], n)
    want = _oracle_lists(U, V, users[:3], n, None, None)
    for (gi, gs), (wi, ws) in zip(got, want):
        assert np.array_equal(gi, wi) and np.array_equal(gs.view(np.uint32), ws.view(np.uint32))


def test_wrmf_recommend_after_training(eng):
    """The WritePredictions loop (ItemRecommendation/Extensions.cs:75-85): ignore = the user's training items."""
    engine, ctx = eng
    nu, ni, k = 250, 120, 32
    u, i = events(nu, ni, 6000, 9)
    f = engine.DeviceFeedback(ctx, u, i, max_user=nu - 1, max_item=ni - 1)
    m = engine.WrmfModel(ctx, f, k)
    m.init_model(5)
    for _ in range(3):
        m.iterate()
    U, V = m.get_model()
    ptr, cols = f.csr(False)
    users = np.arange(0, nu, 7, dtype=np.int32)
    ign = [cols[ptr[x]:ptr[x + 1]] for x in users]
    got = m.recommend(users, 10, None, ign)
    want = _oracle_lists(U, V, users, 10, None, ign)
    for b, ((gi, gs), (wi, ws)) in enumerate(zip(got, want)):
        assert np.array_equal(gi, wi) and np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
        assert not set(gi) & set(ign[b])


@pytest.mark.parametrize("k", [12, 128])
def test_wrmf_retrain_rows_match_oracle(eng, k):
    """RetrainUser / RetrainItem (WRMF.cs:159-170): only the given rows change, to the reference's optimum given the
    other side (rows are independent, so the oracle's full half-sweep on a copy supplies the expected rows)."""
    engine, ctx = eng
    nu, ni = 3000, 2500                                                   # both sides >= 16 k rows: tensor path at k = 128
    u, i = events(nu, ni, 40000, 7 + k)
    f = engine.DeviceFeedback(ctx, u, i, max_user=nu - 1, max_item=ni - 1)
    rng = O.Random(5)
    U = rng.init_normal(nu * k).reshape(nu, k); V = rng.init_normal(ni * k).reshape(ni, k)
    m = engine.WrmfModel(ctx, f, k, 1.0, 0.015)
    m.set_model(U, V)
    uptr, ucols = O.feedback_csr(u, i, nu - 1)
    iptr, irows = O.feedback_csr(i, u, ni - 1)
    users = np.array([0, 17, 2999, 400], np.int32)
    m.retrain(users, by_item=False)
    Uo = U.copy()
    O.wrmf_optimize(uptr, ucols, Uo, V, 1.0, 0.015)
    Ug, Vg = m.get_model()
    untouched = np.setdiff1d(np.arange(nu), users)
    assert np.array_equal(Ug[untouched], U[untouched]) and np.array_equal(Vg, V)
    scale = np.abs(Uo[users]).max(axis=1, keepdims=True) + 1e-12
    assert (np.abs(Ug[users] - Uo[users]) / scale).max() < 1e-4
    items = np.array([3, 1], np.int32)
    m.retrain(items, by_item=True)
    Vo = V.copy()
    O.wrmf_optimize(iptr, irows, Vo, Ug, 1.0, 0.015)
    _, Vg2 = m.get_model()
    assert np.array_equal(np.delete(Vg2, items, axis=0), np.delete(V, items, axis=0))
    scale = np.abs(Vo[items]).max(axis=1, keepdims=True) + 1e-12
    assert (np.abs(Vg2[items] - Vo[items]) / scale).max() < 1e-4
