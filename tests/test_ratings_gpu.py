"""Rating-matrix build on the device (counts, CSR, statistics, block partition) -- bit-exact with the oracle
and with the reference's own known answers (src/Tests/Data/StaticRatingsTest.cs:144-200)."""
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


def test_static_ratings_known_answers(eng):
    """The fixture of StaticRatingsTest.cs (ByUser/ByItem/CountByUser/CountByItem)."""
    engine, ctx = eng
    rows = [(1, 4, .3), (1, 8, .2), (2, 4, .2), (2, 2, .6), (2, 5, .4), (3, 7, .2), (6, 3, .3)]
    u = np.array([r[0] for r in rows], np.int32); i = np.array([r[1] for r in rows], np.int32)
    v = np.array([r[2] for r in rows], np.float32)
    r = engine.DeviceRatings(ctx, u, i, v)
    assert list(r.counts(False)) == [0, 2, 3, 1, 0, 0, 1]
    assert list(r.counts(True)) == [0, 0, 1, 1, 2, 1, 0, 1, 1]
    ptr, idx = r.csr(False)
    by_user = [list(idx[ptr[x]:ptr[x + 1]]) for x in range(7)]
    assert by_user == [[], [0, 1], [2, 3, 4], [5], [], [], [6]]
    ptr, idx = r.csr(True)
    assert list(idx[ptr[4]:ptr[5]]) == [0, 2]
    avg, mn, mx = r.stats()
    assert mn == np.float32(.2) and mx == np.float32(.6)
    assert avg == O.lib().mo_average(v, v.size)


@pytest.mark.parametrize("n,nu,ni", [(1, 1, 1), (1000, 37, 11), (200_000, 5000, 3000), (1_500_000, 70_000, 10_000)])
def test_counts_csr_stats_match_oracle(eng, n, nu, ni):
    engine, ctx = eng
    rs = np.random.RandomState(n % 1000)
    u = rs.randint(0, nu, n).astype(np.int32); i = rs.randint(0, ni, n).astype(np.int32)
    v = (rs.randint(1, 11, n) / 2).astype(np.float32)
    r = engine.DeviceRatings(ctx, u, i, v, max_user=nu + 2, max_item=ni)   # trailing ids without ratings
    assert np.array_equal(r.counts(False), O.count_by(u, nu + 2))
    assert np.array_equal(r.counts(True), O.count_by(i, ni))
    for by_item, ids, mx in ((False, u, nu + 2), (True, i, ni)):
        ptr, idx = r.csr(by_item)
        optr, oidx = O.build_index(ids, mx)
        assert np.array_equal(ptr, optr) and np.array_equal(idx, oidx)
    avg, mn, mxr = r.stats()
    lo, hi = np.zeros(1, np.float32), np.zeros(1, np.float32)
    assert mn == v.min() and mxr == v.max()
    assert avg == pytest.approx(O.lib().mo_average(v, v.size), rel=1e-6)


@pytest.mark.parametrize("g", [1, 3, 8])
def test_partition_blocks_match_oracle(eng, g):
    engine, ctx = eng
    rs = np.random.RandomState(g)
    n, nu, ni = 50_000, 700, 300
    u = rs.randint(0, nu, n).astype(np.int32); i = rs.randint(0, ni, n).astype(np.int32)
    v = np.ones(n, np.float32)
    r = engine.DeviceRatings(ctx, u, i, v, max_user=nu - 1, max_item=ni - 1)
    up = rs.permutation(nu).astype(np.int32); ip = rs.permutation(ni).astype(np.int32)
    ptr, idx = r.partition_blocks(up, ip, g)
    optr, oidx = O.partition_blocks_given(u, i, up, ip, g)
    assert np.array_equal(ptr, optr) and np.array_equal(idx, oidx)


@pytest.mark.parametrize("n,g", [(10, 3), (10, 50), (1, 1), (1000, 7), (300_000, 4736), (1_000_003, 9472)])
def test_partition_indices_match_oracle(eng, n, g):
    """MultiCore.PartitionIndices (MultiCore.cs:79-92): RandomIndex dealt round-robin into min(groups, n) lists --
    bit-exact with the oracle's restatement (itself held by the known answers of src/Tests/MulticoreTest.cs:27-93)."""
    engine, ctx = eng
    ri = O.Random(n).shuffle(np.arange(n))
    g2, optr, oidx = O.partition_indices(ri, g)
    ptr, idx = engine.partition_indices(ctx, ri, g)
    assert np.array_equal(ptr[:g2 + 1], optr[:g2 + 1]) and np.all(ptr[g2:] == n)
    assert np.array_equal(idx, oidx[:n])
    assert np.array_equal(np.sort(idx), np.arange(n))


def test_empty_rating_set(eng):
    engine, ctx = eng
    e = np.zeros(0, np.int32)
    r = engine.DeviceRatings(ctx, e, e, np.zeros(0, np.float32), max_user=3, max_item=2)
    assert list(r.counts(False)) == [0, 0, 0, 0]
    ptr, idx = r.csr(True)
    assert list(ptr) == [0, 0, 0, 0] and idx.size == 0


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 1), (17, 3), (1000, 4), (300_000, 5), (3_000_000, 6)])
def test_shuffle_apply_is_the_reference_fisher_yates(eng, n, seed):
    """Utils.cs:52-64 with the host RNG's swap targets: the device result is the sequential loop's, bit for bit."""
    engine, ctx = eng
    H = O.Random(seed).shuffle_targets(n)
    want = np.arange(n, dtype=np.int32)
    O.lib().mo_shuffle_apply(want, H, n)
    assert np.array_equal(want, O.Random(seed).shuffle(np.arange(n)))     # oracle self-consistency
    got = np.arange(n, dtype=np.int32)
    engine._capi.check(ctx.lib.mml_shuffle_apply(ctx.h, got, H, n))
    assert np.array_equal(got, want)


def test_shuffle_apply_adversarial_targets(eng):
    """All swaps through cell 0 (a dependence chain of length n) and the identity."""
    engine, ctx = eng
    n = 5000
    for H in (np.zeros(n, np.int32), np.arange(n, dtype=np.int32)):
        want = np.arange(n, dtype=np.int32)
        O.lib().mo_shuffle_apply(want, H, n)
        got = np.arange(n, dtype=np.int32)
        engine._capi.check(ctx.lib.mml_shuffle_apply(ctx.h, got, H, n))
        assert np.array_equal(got, want)


def test_ingest_feeds_the_device_from_pinned_memory(eng, tmp_path):
    """Text file -> native parallel reader -> pinned COO -> mml_ratings_create / mml_feedback_create (no copy through the
    host language): same counts, CSR and statistics as uploading the arrays by hand."""
    from mymedialite_b200 import ingest
    engine, ctx = eng
    rng = np.random.default_rng(12)
    n = 300_000
    u = rng.integers(0, 4000, n).astype(np.int32); i = rng.integers(0, 1500, n).astype(np.int32)
    v = (rng.integers(1, 11, n) / 2).astype(np.float32)
    path = tmp_path / "ratings.tsv"
    with open(path, "w") as w:
        w.write("".join("%d\t%d\t%s\n" % (a, b, c) for a, b, c in zip(u, i, v)))
    parsed = ingest.StaticRatingData.Read(str(path))
    assert parsed.pinned and parsed.Count == n
    r = parsed.to_device_ratings(ctx)
    ref = engine.DeviceRatings(ctx, u, i, v)
    assert np.array_equal(r.counts(False), ref.counts(False)) and np.array_equal(r.counts(True), ref.counts(True))
    for by_item in (False, True):
        a, b = r.csr(by_item), ref.csr(by_item)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert r.stats() == ref.stats()
    fb = ingest.ItemData.Read(str(path)).to_device_feedback(ctx)
    want = engine.DeviceFeedback(ctx, u, i)
    assert fb.nnz == want.nnz
    assert all(np.array_equal(x, y) for x, y in zip(fb.csr(), want.csr()))
