"""Recommend() on the tcgen05 path (BF16 / TF32 scoring + fused top-k + exact re-scoring of the finalists) against the CPU
oracle and against the exact CUDA-core path: item ids, order and fp32 scores must be bit-identical
(Recommender.cs:52-103, ItemRecommendation/MF.cs:151-157, DataType/MatrixExtensions.cs:224-241)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["bf16", "tf32"])
def eng(request):
    """Every test runs under both operand precisions of the filter GEMM (mml_topn_set_filter): the results must be the
    same bits either way, only the number of exactly re-scored finalists differs."""
    from mymedialite_b200 import engine
    ctx = engine.Context(0)
    engine.topn_set_filter(engine._capi.TOPN_FILTER_BF16 if request.param == "bf16" else engine._capi.TOPN_FILTER_TF32)
    yield engine, ctx
    engine.topn_set_mode(engine._capi.TOPN_AUTO)
    engine.topn_set_filter(engine._capi.TOPN_FILTER_TF32)
    ctx.close()


def exact_scores(U, V):
    """RowScalarProduct for every pair: fp32 multiply, then fp32 add, f = 0 .. k-1 in order (no FMA)."""
    s = np.zeros((U.shape[0], V.shape[0]), np.float32)
    for f in range(U.shape[1]):
        s = s + U[:, f, None] * V[None, :, f]
    return s


def reference_topn(U, V, users, n, cand, ignore_lists):
    """(items, scores) per user ordered by (score desc, candidate position asc)."""
    ni = V.shape[0]
    cand = np.arange(ni, dtype=np.int32) if cand is None else cand
    ok_c = (cand >= 0) & (cand < ni)
    Vc = V[np.where(ok_c, cand, 0)]
    out = []
    for b, u in enumerate(users):
        if u < 0 or u >= U.shape[0]:
            out.append((np.zeros(0, np.int32), np.zeros(0, np.float32)))
            continue
        s = exact_scores(U[u:u + 1], Vc)[0]
        keep = ok_c.copy()
        if ignore_lists is not None:
            keep &= ~np.isin(cand, ignore_lists[b])
        pos = np.flatnonzero(keep)
        order = pos[np.lexsort((pos, -s[pos].astype(np.float64)))][:n]
        out.append((cand[order].astype(np.int32), s[order]))
    return out


def check(got, want):
    assert len(got) == len(want)
    for b, ((gi, gs), (wi, ws)) in enumerate(zip(got, want)):
        assert np.array_equal(gi, wi), (b, gi, wi)
        assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32)), (b, gs, ws)


@pytest.mark.parametrize("k", [128, 64, 10, 100])
@pytest.mark.parametrize("n", [1, 10, 16])
def test_tensor_path_is_bit_exact(eng, k, n):
    engine, ctx = eng
    rs = np.random.RandomState(1000 * k + n)
    nu, ni = 700, 3000
    U = (rs.randn(nu, k) * 0.1).astype(np.float32); V = (rs.randn(ni, k) * 0.1).astype(np.float32)
    V[7] = V[3]; V[2000] = V[3]; U[5] = 0                                # exact ties: position decides
    users = np.concatenate([rs.permutation(nu)[:520], [5, 5, nu + 3, -1]]).astype(np.int32)   # > 2 row tiles, invalid ids
    ign = [rs.permutation(ni)[:rs.randint(0, 60)].astype(np.int32) for _ in users]          # unsorted ignore lists
    engine.topn_set_mode(engine._capi.TOPN_TENSOR)
    got = engine.topn_mf(ctx, U, V, users, n, None, ign)
    st = engine.topn_last_stats()
    engine.topn_set_mode(engine._capi.TOPN_AUTO)
    check(got, reference_topn(U, V, users, n, None, ign))
    assert st["users_tensor_path"] >= 0.95 * users.size, st             # the filter decides nearly every user itself


def test_tensor_path_candidates_and_small_shapes(eng):
    engine, ctx = eng
    rs = np.random.RandomState(5)
    k, nu, ni = 32, 40, 90                                               # fewer candidates than one MMA tile
    U = (rs.randn(nu, k)).astype(np.float32); V = (rs.randn(ni, k)).astype(np.float32)
    users = np.arange(nu, dtype=np.int32)
    cand = rs.permutation(ni + 6)[:70].astype(np.int32)                  # shuffled, some ids outside the model
    ign = [rs.choice(ni, 30, replace=False).astype(np.int32) for _ in users]
    engine.topn_set_mode(engine._capi.TOPN_TENSOR)
    try:
        for n in (1, 5, 16):
            check(engine.topn_mf(ctx, U, V, users, n, cand, ign), reference_topn(U, V, users, n, cand, ign))
        # more requested than there are candidates left: counts shrink
        ign_all = [np.arange(ni - 3, dtype=np.int32) for _ in users]
        got = engine.topn_mf(ctx, U, V, users, 10, None, ign_all)
        assert all(len(gi) == 3 for gi, _ in got)
        check(got, reference_topn(U, V, users, 10, None, ign_all))
    finally:
        engine.topn_set_mode(engine._capi.TOPN_AUTO)


def test_undecidable_users_fall_back_to_exact_scoring(eng):
    """Scores packed closer together than the TF32 error bound: the filter must notice and hand the users over."""
    engine, ctx = eng
    rs = np.random.RandomState(8)
    k, nu, ni = 128, 300, 40000                                          # > 256 near-ties per candidate split
    base = rs.randn(k).astype(np.float32)
    V = (base[None, :] + 1e-5 * rs.randn(ni, k)).astype(np.float32)     # all items score within ~1e-4 relative
    U = (rs.randn(nu, k) * 0.1).astype(np.float32)
    users = np.arange(nu, dtype=np.int32)
    engine.topn_set_mode(engine._capi.TOPN_AUTO)
    got = engine.topn_mf(ctx, U, V, users, 10)
    st = engine.topn_last_stats()
    check(got, reference_topn(U, V, users, 10, None, None))
    assert st["users_exact_path"] > 0, st
    # non-finite factors: error bound is not finite -> exact path decides (NaN scores never qualify)
    V2 = (rs.randn(ni, k) * 0.1).astype(np.float32); V2[11, 3] = np.nan; V2[12, 0] = np.inf
    got = engine.topn_mf(ctx, U, V2, users[:40], 10)
    engine.topn_set_mode(engine._capi.TOPN_EXACT)
    want = engine.topn_mf(ctx, U, V2, users[:40], 10)
    engine.topn_set_mode(engine._capi.TOPN_AUTO)
    check(got, want)


def test_tensor_and_exact_paths_agree_at_scale(eng):
    """Size-independent property on a shape the oracle would not finish in seconds: both device paths return the
    same bits (20k users x 30k items, k = 128, n = 10, 20 ignored items per user)."""
    engine, ctx = eng
    rs = np.random.RandomState(21)
    k, nu, ni = 128, 20000, 30000
    U = (rs.randn(nu, k) * 0.1).astype(np.float32); V = (rs.randn(ni, k) * 0.1).astype(np.float32)
    users = np.arange(nu, dtype=np.int32)
    ign = list(rs.randint(0, ni, (nu, 20)).astype(np.int32))
    engine.topn_set_mode(engine._capi.TOPN_TENSOR)
    a = engine.topn_mf(ctx, U, V, users, 10, None, ign)
    st = engine.topn_last_stats()
    engine.topn_set_mode(engine._capi.TOPN_EXACT)
    b = engine.topn_mf(ctx, U, V, users, 10, None, ign)
    engine.topn_set_mode(engine._capi.TOPN_AUTO)
    check(a, b)
    assert st["users_tensor_path"] > 0.99 * nu, st
    for bidx in (0, 77, nu - 1):
        assert not set(a[bidx][0]) & set(ign[bidx])


def test_request_outside_the_envelope_is_refused_in_forced_mode(eng):
    engine, ctx = eng
    U = np.ones((4, 8), np.float32); V = np.ones((9, 8), np.float32)
    engine.topn_set_mode(engine._capi.TOPN_TENSOR)
    try:
        with pytest.raises(Exception):
            engine.topn_mf(ctx, U, V, np.arange(4, dtype=np.int32), -1)
    finally:
        engine.topn_set_mode(engine._capi.TOPN_AUTO)
