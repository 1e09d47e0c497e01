"""The oracle's ranking measures against the reference's own known answers
(src/Tests/Eval/Measures/AUCTest.cs:33-92, PrecisionAndRecallTest.cs:31-73, src/Tests/Eval/ItemsTest.cs:34-77)."""
import numpy as np

from oracle import oracle as O


def test_auc_known_answers():
    ranking = [1, 2, 3, 4]
    assert O.auc_compute(ranking, [1], 0) == 1.0 and O.auc_compute(ranking, [1, 2], 0) == 1.0
    assert O.auc_compute(ranking, [1, 2, 3], 0) == 1.0
    assert O.auc_compute(ranking, [4], 0) == 0.0
    assert abs(O.auc_compute(ranking, [3], 0) - 0.333) < 0.01 and abs(O.auc_compute(ranking, [2], 0) - 0.666) < 0.01
    assert O.auc_compute(ranking, [1, 3], 0) == 0.75
    assert O.auc_compute(ranking, [1, 2, 3, 4], 0) == 0.5
    assert O.auc_compute(ranking, [2, 4], 0) == 0.25
    for i in range(10):
        assert O.auc_compute(ranking, [1], i) == 1.0 and O.auc_compute(ranking, [1, 2], i) == 1.0
        assert O.auc_compute(ranking, [1, 2, 3], i) == 1.0
        assert O.auc_compute(ranking, [4], i) == i / (i + 3)


def test_precision_recall_known_answers():
    list5, list1, list3, list_last = [1, 2, 3, 4, 5], [1], [1, 2, 3], [5]
    assert O.ap_compute(list5, list1) == 1 and O.ap_compute(list5, list5) == 1
    assert O.ap_compute(list3, list_last) == 0 and O.ap_compute(list5, list_last) == 1 / 5
    for n in (1, 2, 3, 4):
        assert O.hits_at(list3, list1, n) == 1
    assert O.hits_at(list1, list1, 1) == 1
    assert O.precision_at(list1, list1, 1) == 1 and O.precision_at(list3, list1, 1) == 1
    assert O.precision_at(list3, list1, 2) == 1 / 2 and O.precision_at(list3, list1, 3) == 1 / 3
    for n in (1, 2, 3):
        assert O.recall_at(list3, list1, n) == 1
    assert O.recall_at(list1, list1, 1) == 1


def _most_popular(train_items):
    """ItemRecommendation/MostPopular.cs: score = number of training events of the item (items beyond MaxItemID: float.MinValue)."""
    counts = np.bincount(np.asarray(train_items))

    def recommend(user, n, ignore, candidates):
        scored = [(c, float(counts[c])) for c in candidates if c not in ignore and c < counts.size]
        scored.sort(key=lambda t: -t[1])          # stable, as OrderByDescending
        return scored if n < 0 else scored[:n]
    return recommend


def test_items_evaluate_known_answers():
    """ItemsTest.TestEvalDefault: one evaluated list, AUC 0.5, prec@5 0 (OVERLAP candidates = {3})."""
    train_u, train_i = [1, 1, 2, 2, 3, 3], [1, 2, 2, 3, 1, 2]
    test_u, test_i = [2, 2, 4], [3, 4, 4]
    overlap = [x for x in O.first_seen(test_i) if x in set(train_i)]
    assert overlap == [3]
    res, rows = O.items_evaluate(_most_popular(train_i), test_u, test_i, train_u, train_i, candidate_items=overlap)
    assert res["num_lists"] == 1 and res["AUC"] == 0.5 and abs(res["prec@5"]) < 0.01
    res, _ = O.items_evaluate(_most_popular(train_i), test_u, test_i, train_u, train_i, test_users=[1, 2, 3, 4],
                              candidate_items=overlap)
    assert res["num_lists"] == 1 and res["AUC"] == 0.5


def test_measures_on_a_worked_example():
    ranked, correct = [7, 3, 9, 1, 4, 8], {3, 4, 99}
    assert abs(O.ap_compute(ranked, correct) - (1 / 2 + 2 / 5) / 3) < 1e-15
    assert O.reciprocal_rank(ranked, correct) == 0.5
    import math
    want = (1 / math.log2(3) + 1 / math.log2(6)) / (1 / math.log2(2) + 1 / math.log2(3) + 1 / math.log2(4))
    assert abs(O.ndcg_compute(ranked, correct) - want) < 1e-12
    # 6 listed + 2 dropped, one of them the missing relevant item 99
    assert abs(O.auc_compute(ranked, correct, 2) - (1 * 2 + 1 * 1 + 1 * 1 + 2 * 1) / ((8 - 2) * 2)) < 1e-15
