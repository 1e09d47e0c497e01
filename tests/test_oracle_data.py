"""Oracle vs. the reference's own known answers for the rating-matrix build
(src/Tests/Data/StaticRatingsTest.cs:144-200, RatingsTest.cs, MulticoreTest.cs:27-93,
PosOnlyFeedbackTest.cs:77-154, DataType/MatrixExtensionsTest.cs)."""
import numpy as np

from oracle import oracle as O

U = np.array([1, 1, 2, 2, 2, 3, 6], dtype=np.int32)
I = np.array([4, 8, 4, 2, 5, 7, 3], dtype=np.int32)


def test_count_by_user_item_known_answers():
    cu = O.count_by(U, 6)
    ci = O.count_by(I, 8)
    assert cu.tolist() == [0, 2, 3, 1, 0, 0, 1]
    assert ci.tolist() == [0, 0, 1, 1, 2, 1, 0, 1, 1]


def test_by_user_by_item_known_answers():
    ptr, idx = O.build_index(U, 6)
    assert set(idx[ptr[1]:ptr[2]].tolist()) == {0, 1}
    ptr_i, idx_i = O.build_index(I, 8)
    assert set(idx_i[ptr_i[4]:ptr_i[5]].tolist()) == {0, 2}
    # ascending rating index inside every row (single forward pass, DataSet.cs:171-191)
    for r in range(7):
        row = idx[ptr[r]:ptr[r + 1]]
        assert np.all(np.diff(row) > 0)
    assert ptr[-1] == U.size


def test_random_index_length_and_permutation():
    rng = O.Random(2)
    ri = np.empty(300, dtype=np.int32)
    O.lib().mo_random_index(rng.ref, ri, 300)
    assert np.array_equal(np.sort(ri), np.arange(300))


def _random_ratings(rng, nu, ni, n):
    # src/Tests/TestUtils.cs:27-41
    u = np.empty(n, np.int32); i = np.empty(n, np.int32); v = np.empty(n, np.float32)
    for t in range(n):
        u[t] = rng.next_max(nu); i[t] = rng.next_max(ni); v[t] = 1 + rng.next_max(5)
    return u, i, v


def test_partition_users_and_items_shapes():
    rng = O.Random(11)
    for nu, ni, groups, exp_g in [(15, 30, 3, 3), (15, 30, 20, None), (30, 15, 20, None)]:
        u, i, v = _random_ratings(rng, nu, ni, 300)
        mu, mi = int(u.max()), int(i.max())
        g, ptr, idx, up, ip = O.partition_users_and_items(rng, u, i, mu, mi, groups)
        assert g == (exp_g if exp_g else min(groups, mu + 1, mi + 1))
        assert ptr.size == g * g + 1 and ptr[-1] == 300
        assert np.array_equal(np.sort(idx), np.arange(300))
        for b in range(g * g):
            for t in idx[ptr[b]:ptr[b + 1]]:
                assert up[u[t]] % g == b // g and ip[i[t]] % g == b % g


def test_partition_indices_known_answers():
    rng = O.Random(4)
    ri = np.empty(300, dtype=np.int32)
    O.lib().mo_random_index(rng.ref, ri, 300)
    g, ptr, idx = O.partition_indices(ri, 10)
    assert g == 10 and np.all(np.diff(ptr) == 30)
    assert np.array_equal(idx[ptr[3]:ptr[4]], ri[3::10])
    g, ptr, idx = O.partition_indices(ri[:10], 50)
    assert g == 10


def test_row_scalar_product_and_average():
    a = np.arange(1, 6, dtype=np.float32)
    assert O.lib().mo_row_scalar_product(a, a, 5) == 55.0
    v = np.array([1.0, 1.5, 3.0, 5.0, 3.5, 1.0, 4.0, 2.0, 4.5], dtype=np.float32)
    assert O.lib().mo_average(v, v.size) == np.float32(np.float32(v.astype(np.float64).sum()) / np.float32(9))


def test_feedback_matrix_known_answers():
    fu = np.array([1, 1, 2, 2, 2, 3, 6, 8], dtype=np.int32)
    fi = np.array([4, 8, 4, 2, 5, 7, 3, 1], dtype=np.int32)
    ptr, cols = O.feedback_csr(fu, fi, 8)
    has = lambda r, c: c in cols[ptr[r]:ptr[r + 1]]
    assert has(2, 5) and has(1, 4) and has(6, 3) and has(2, 2)
    assert not has(5, 2) and not has(4, 1) and not has(3, 6)
    ptr_i, cols_i = O.feedback_csr(fi, fu, 8)
    has_i = lambda r, c: c in cols_i[ptr_i[r]:ptr_i[r + 1]]
    assert has_i(5, 2) and has_i(4, 1) and has_i(3, 6) and has_i(2, 2) and not has_i(2, 5)
    # duplicates collapse (rows are HashSets, SparseBooleanMatrix.cs:37-67)
    ptr2, cols2 = O.feedback_csr(np.array([0, 0, 0], np.int32), np.array([3, 3, 1], np.int32), 0)
    assert cols2.tolist() == [3, 1]
