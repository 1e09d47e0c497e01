"""Host-side mirror of Eval.Items (Eval/Items.cs:51-240) for the CUDA-backed item recommenders: candidate selection,
per-user skip rules and averaging as in the reference, the ranking and the measures on the device
(mml_wrmf_evaluate / mml_items_evaluate_mf: one launch for all test users instead of a Recommend() call per user).
"""
import numpy as np

from . import engine, sysrandom

Measures = engine.ITEM_MEASURES                                          # Items.cs:37-49

TRAINING, TEST, OVERLAP, UNION, EXPLICIT = "TRAINING", "TEST", "OVERLAP", "UNION", "EXPLICIT"   # CandidateItems


def _first_seen(ids):
    """HashSet<int> filled in order, then ToArray() (Data/DataSet.cs:112-131): distinct ids in order of first appearance."""
    ids = np.asarray(ids, np.int64)
    if ids.size == 0:
        return np.zeros(0, np.int32)
    _, first = np.unique(ids, return_index=True)
    return ids[np.sort(first)].astype(np.int32)


def Candidates(candidate_items, candidate_item_mode, test, training):
    """Items.Candidates (Items.cs:62-95), including the final Shuffle() (one RNG draw per element)."""
    test_items = _first_seen(test.Items) if test is not None else np.zeros(0, np.int32)
    if candidate_item_mode == TRAINING:
        result = _first_seen(training.Items)
    elif candidate_item_mode == TEST:
        result = test_items
    elif candidate_item_mode == OVERLAP:
        result = test_items[np.isin(test_items, _first_seen(training.Items))]          # Enumerable.Intersect: order of the first
    elif candidate_item_mode == UNION:
        train_items = _first_seen(training.Items)
        result = np.concatenate([test_items, train_items[~np.isin(train_items, test_items)]])
    elif candidate_item_mode == EXPLICIT:
        if candidate_items is None:
            raise ValueError("candidate_items")                                        # ArgumentNullException
        result = np.asarray(candidate_items, np.int32).copy()
    else:
        raise ValueError("Unknown candidate_item_mode: %s" % candidate_item_mode)
    return sysrandom.get_instance().shuffle(np.ascontiguousarray(result, np.int32))


def _rows(users, items, wanted_users):
    """user -> set of items as CSR rows aligned with wanted_users (SparseBooleanMatrix rows: duplicates collapse)."""
    users = np.asarray(users, np.int64); items = np.asarray(items, np.int64)
    if users.size:
        pairs = np.unique(users * (int(items.max()) + 1) + items)
        pu, pi = pairs // (int(items.max()) + 1), pairs % (int(items.max()) + 1)
    else:
        pu, pi = users, items
    lo = np.searchsorted(pu, wanted_users, side="left")
    hi = np.searchsorted(pu, wanted_users, side="right")
    ptr = np.zeros(len(wanted_users) + 1, np.int64)
    ptr[1:] = np.cumsum(hi - lo)
    idx = np.concatenate([pi[a:b] for a, b in zip(lo, hi)]) if ptr[-1] else np.zeros(0, np.int64)
    return ptr, idx.astype(np.int32)


def Evaluate(recommender, test, training, test_users=None, candidate_items=None, candidate_item_mode=OVERLAP,
             repeated_events=False, n=-1):
    """Items.Evaluate (Items.cs:126-209). `recommender` is a trained WRMF of this package (its model stays on the
    device). Returns the reference's result dictionary: the measures averaged over the evaluated users, num_users,
    num_lists, num_items."""
    if test_users is None:
        test_users = _first_seen(test.Users)
    test_users = np.ascontiguousarray(test_users, np.int32)
    cand = Candidates(candidate_items, candidate_item_mode, test, training)
    result = {m: np.float32(0) for m in Measures}
    num_users = 0
    if test_users.size and cand.size:
        test_rows = _rows(test.Users, test.Items, test_users)
        ignore_rows = None if repeated_events else _rows(training.Users, training.Items, test_users)
        rows, used = recommender._model.evaluate(test_users, cand, test_rows, ignore_rows, n)
        for b in np.nonzero(used == 1)[0]:                       # the reference adds (float) values one user at a time
            for j, m in enumerate(Measures):
                result[m] = np.float32(result[m] + rows[b, j])
        num_users = int((used == 1).sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        for m in Measures:
            result[m] = float(np.float32(result[m] / np.float32(num_users)))
    result["num_users"] = num_users
    result["num_lists"] = num_users
    result["num_items"] = int(cand.size)
    return result


def ComputeFit(recommender, test_users=None, candidate_items=None, candidate_item_mode=OVERLAP):
    """Items.ComputeFit (Items.cs:218-229): AUC on the training data, repeated events allowed."""
    fb = recommender.Feedback
    return Evaluate(recommender, fb, fb, test_users, candidate_items, candidate_item_mode, True)["AUC"]
