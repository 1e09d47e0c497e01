"""Host-side logic (no GPU): MyMediaLite.Random restatement vs the oracle, model text format, synthetic data, sharding."""
import io

import numpy as np
import pytest

from oracle import oracle as O
from mymedialite_b200 import dist, modelio, synthetic
from mymedialite_b200.sysrandom import SystemRandom


def test_system_random_known_answers():
    # the BCL generator's widely published first outputs
    assert [SystemRandom(0).next() for _ in range(1)] == [1559595546]
    r = SystemRandom(1)
    assert [r.next() for _ in range(3)] == [534011718, 237820880, 1002897798]
    r = SystemRandom(42)
    assert [r.next() for _ in range(3)] == [1434747710, 302596119, 269548474]
    r = SystemRandom(1)
    assert [r.next(10) for _ in range(10)] == [2, 1, 4, 7, 6, 4, 3, 9, 1, 6]


@pytest.mark.parametrize("seed", [0, 1, 7, -12345, 2 ** 31 - 1])
def test_host_rng_equals_oracle(seed):
    a, b = SystemRandom(seed), O.Random(seed)
    assert [a.next() for _ in range(300)] == [b.next() for _ in range(300)]
    assert np.array_equal(a.shuffle_targets(500), b.shuffle_targets(500))
    assert np.array_equal(a.shuffle(np.arange(77)), b.shuffle(np.arange(77)))
    assert np.array_equal(a.init_normal(9, 11, 0.0, 0.1).ravel(), b.init_normal(99, 0.0, 0.1))


def test_float_format_is_dotnet_g7():
    cases = {0.1: "0.1", 1234567.0: "1234567", 12345678.0: "1.234568E+07", 1e-5: "1E-05", 1.5e-5: "1.5E-05",
             0.000123456789: "0.0001234568", -3.0: "-3", 1e10: "1E+10", 0.0: "0", 3.5: "3.5", 0.015: "0.015"}
    for x, s in cases.items():
        assert modelio.fmt(x) == s
    for x in np.random.RandomState(0).randn(200).astype(np.float32) * 0.1:
        assert abs(float(modelio.fmt(x)) - x) <= 1e-6 * max(1.0, abs(x))


def test_matrix_and_vector_text_round_trip():
    m = (np.random.RandomState(1).randn(5, 3) * 0.1).astype(np.float32)
    v = np.array([0.5, -1.25, 3.0], np.float32)
    w = io.StringIO()
    modelio.write_header(w, "MyMediaLite.RatingPrediction.CudaBiasedMatrixFactorization")
    modelio.write_vector(w, v)
    modelio.write_matrix(w, m)
    text = w.getvalue()
    lines = text.split("\n")
    assert lines[1] == "2.99" and lines[2] == "3" and lines[6] == "5 3" and lines[7].startswith("0 0 ")
    assert lines[7 + 15] == ""                       # WriteMatrix ends with an empty line (IO/MatrixExtensions.cs:37)
    r = io.StringIO(text)
    assert modelio.read_header(r, "x") == "MyMediaLite.RatingPrediction.CudaBiasedMatrixFactorization"
    np.testing.assert_allclose(modelio.read_vector(r), v)
    np.testing.assert_allclose(modelio.read_matrix(r), m, rtol=1e-6, atol=1e-7)


def test_synthetic_is_deterministic_and_dense():
    a = synthetic.ratings(500, 120, 8000, "half", 3)
    b = synthetic.ratings(500, 120, 8000, "half", 3)
    for x, y in zip(a["train"], b["train"]):
        assert np.array_equal(x, y)
    u, i, v = a["train"]; tu, ti, tv = a["test"]
    allu, alli = np.concatenate([u, tu]), np.concatenate([i, ti])
    assert np.unique(allu).size == 500 and np.unique(alli).size == 120          # ids dense
    assert np.unique(allu.astype(np.int64) * 120 + alli).size == allu.size       # (user, item) pairs distinct
    assert set(np.unique(v)) <= set(np.arange(1, 11) / 2)                       # half-star levels
    assert 0.05 < tu.size / allu.size < 0.15


def test_shards_share_the_item_catalogue():
    a = synthetic.ratings(800, 200, 60000, "int", 11, item_seed=5)
    b = synthetic.ratings(800, 200, 60000, "int", 12, item_seed=5)
    c = synthetic.ratings(800, 200, 60000, "int", 12, item_seed=6)
    ca, cb, cc = (np.bincount(x["train"][1], minlength=200) for x in (a, b, c))
    assert np.corrcoef(ca, cb)[0, 1] > 0.9 > np.corrcoef(ca, cc)[0, 1]          # same popularity law iff same item_seed
    assert not np.array_equal(a["train"][0], b["train"][0])


def test_shard_by_user_partitions_the_ratings():
    rs = np.random.RandomState(2)
    u = rs.randint(0, 1000, 5000); i = rs.randint(0, 50, 5000); v = rs.rand(5000).astype(np.float32)
    perm = rs.permutation(1000)
    seen = np.zeros(5000, int)
    for rank in range(4):
        su, si, sv, idx = dist.shard_by_user(u, i, v, rank, 4, perm)
        assert np.all(perm[su] % 4 == rank) and np.array_equal(su, u[idx]) and np.array_equal(sv, v[idx])
        seen[idx] += 1
    assert np.all(seen == 1)


def test_merge_schedules_orders_by_subepoch_then_rank():
    parts = [(np.array([10, 11, 12, 13]), np.array([0, 0, 1, 1])), (np.array([20, 21, 22]), np.array([0, 1, 1]))]
    assert list(dist.merge_schedules(parts, 2)) == [10, 11, 20, 12, 13, 21, 22]


def test_parallel_options_map_to_the_engine_grid():
    """MaxThreads > 1 selects the DSGD block schedule (BiasedMatrixFactorization.cs:178-184); NaiveParallelization
    (:136-141, :201-204) the list schedule of MultiCore.PartitionIndices (no exclusivity at all)."""
    from mymedialite_b200 import recommenders as R, _capi
    m = R.BiasedMatrixFactorization()
    assert m._params().schedule == _capi.SCHEDULE_SERIAL
    m.MaxThreads = 8
    p = m._params()
    assert p.schedule == _capi.SCHEDULE_DSGD and p.num_groups == 0 and p.ctas_per_group == 0 and p.max_threads == 8
    m.NaiveParallelization = True
    p = m._params()
    assert p.schedule == _capi.SCHEDULE_NAIVE and p.max_threads == 8
    m.NumGpus = 2                                   # several GPUs keep the block schedule
    assert m._params().schedule == _capi.SCHEDULE_DSGD


# ---- Eval.Items host logic (candidate selection, test rows) -- no device needed ---------------------------------------
def test_items_candidates_modes_follow_the_reference():
    """Items.Candidates (Eval/Items.cs:62-95): AllItems are first-seen distinct ids; Intersect / Union keep the order of the
    first sequence; the result is shuffled with MyMediaLite.Random (one Next(i + 1) per element, i = n-1 .. 0)."""
    from mymedialite_b200 import evalitems as E, recommenders as R, sysrandom
    from oracle import oracle as O
    training = R.PosOnlyFeedback([0, 0, 1, 2, 2], [5, 3, 3, 9, 5])
    test = R.PosOnlyFeedback([0, 1, 1, 3], [9, 7, 5, 7])
    expect = {E.TRAINING: [5, 3, 9], E.TEST: [9, 7, 5], E.OVERLAP: [9, 5], E.UNION: [9, 7, 5, 3]}
    for mode, base in expect.items():
        sysrandom.seed(4)
        got = E.Candidates(None, mode, test, training)
        want = O.Random(4).shuffle(np.array(base, np.int32))
        assert got.tolist() == want.tolist(), mode
    sysrandom.seed(4)
    assert sorted(E.Candidates([4, 2, 8], E.EXPLICIT, test, training).tolist()) == [2, 4, 8]
    with pytest.raises(ValueError):
        E.Candidates(None, E.EXPLICIT, test, training)
    with pytest.raises(ValueError):
        E.Candidates(None, "SOMETHING", test, training)


def test_items_rows_are_sets_aligned_with_the_test_users():
    from mymedialite_b200 import evalitems as E
    ptr, idx = E._rows([2, 0, 2, 2, 5], [7, 1, 3, 7, 0], np.array([2, 4, 0, 5], np.int32))
    rows = [idx[ptr[b]:ptr[b + 1]].tolist() for b in range(4)]
    assert rows == [[3, 7], [], [1], [0]]          # duplicates collapse (SparseBooleanMatrix rows), unknown users are empty
    ptr, idx = E._rows([], [], np.array([1], np.int32))
    assert ptr.tolist() == [0, 0] and idx.size == 0
