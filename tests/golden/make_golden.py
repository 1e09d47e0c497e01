#!/usr/bin/env python
"""Generates tests/golden/oracle_config1.json -- outputs of the CPU oracle (oracle/mml_oracle.c) on the reference's
own example files (tests/example.train / example.test, committed verbatim next to this script) and on tiny synthetic
inputs, so that (a) a change in the oracle's arithmetic shows up as a diff of a committed file and (b) the GPU tests can
hold the CUDA path against fixed numbers as well as against the live oracle.

The reference itself (C#) cannot run in this image, so these are ORACLE outputs, not reference outputs: they pin the
restatement against drift, they do not pin it against MyMediaLite (DESIGN.md section 2 lists what does).

    python tests/golden/make_golden.py        # rewrites the fixture; commit the result
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def f32list(a):
    return [float(x) for x in np.asarray(a, np.float32).ravel()]


def hexlist(a):
    """bit patterns: the committed numbers are exact fp32 values, not rounded decimals"""
    return ["%08x" % x for x in np.asarray(a, np.float32).ravel().view(np.uint32)]


def main():
    tr = np.loadtxt(os.path.join(HERE, "example.train"))
    te = np.loadtxt(os.path.join(HERE, "example.test"))
    u, i, v = tr[:, 0].astype(np.int32), tr[:, 1].astype(np.int32), tr[:, 2].astype(np.float32)
    tu, ti, tv = te[:, 0].astype(np.int32), te[:, 1].astype(np.int32), te[:, 2].astype(np.float32)
    out = {"what": "oracle outputs; see make_golden.py", "system_random": {}, "config1": {}}
    for seed in (0, 1, 42):
        r = O.Random(seed)
        out["system_random"][str(seed)] = [r.next() for _ in range(5)]
    r = O.Random(1)
    out["shuffle_targets_seed1_n8"] = [int(x) for x in r.shuffle_targets(8)]
    r = O.Random(1)
    out["init_normal_seed1_first6"] = hexlist(r.init_normal(6))
    # config 1: rating_prediction on example.train / example.test, num_factors=10 num_iter=30, --random-seed=1
    for name, biased in (("BiasedMatrixFactorization", True), ("MatrixFactorization", False)):
        rng = O.Random(1)
        m = O.Model(u, i, v, biased=biased, num_factors=10, num_iter=30)
        m.init(rng)
        per_epoch = []
        for _ in range(30):
            m.iterate(rng)
            e_tr, e_te = m.evaluate(u, i, v), m.evaluate(tu, ti, tv)
            per_epoch.append([e_tr["RMSE"], e_te["RMSE"]])
        out["config1"][name] = {
            "global_bias": hexlist([m.global_bias])[0],
            "rmse_train_test_per_epoch": [[float(np.float32(a)), float(np.float32(b))] for a, b in per_epoch],
            "random_index": [int(x) for x in m.random_index],
            "user_factors_row0": hexlist(m.user_factors[0]),
            "item_factors_row0": hexlist(m.item_factors[0]),
            "user_bias": hexlist(m.user_bias) if biased else None,
            "predict_test": hexlist(m.predict_many(tu, ti)),
            "test_measures": {k: float(np.float32(x)) for k, x in m.evaluate(tu, ti, tv).items()},
        }
    # WRMF + Recommend on the example pairs read as implicit feedback
    rng = O.Random(1)
    nu, ni, k = int(u.max()) + 1, int(i.max()) + 1, 4
    U = rng.init_normal(nu * k).reshape(nu, k); V = rng.init_normal(ni * k).reshape(ni, k)
    uptr, ucols = O.feedback_csr(u, i, nu - 1); iptr, irows = O.feedback_csr(i, u, ni - 1)
    for _ in range(3):
        O.wrmf_optimize(uptr, ucols, U, V, 1.0, 0.015)
        O.wrmf_optimize(iptr, irows, V, U, 1.0, 0.015)
    rec = {}
    for user in range(nu):
        items, scores = O.recommend_mf(U, V, user, 2, None, [int(x) for x in i[u == user]])
        rec[str(user)] = {"items": [int(x) for x in items], "scores": hexlist(scores)}
    out["wrmf_k4_3epochs"] = {"U": hexlist(U), "V": hexlist(V), "top2_ignoring_training_items": rec}
    # ranking measures on a worked list
    ranked, correct = [7, 3, 9, 1, 4, 8], [3, 4, 99]
    out["measures"] = {"ranked": ranked, "correct": correct, "dropped": 2,
                       "AUC": O.auc_compute(ranked, correct, 2), "AP": O.ap_compute(ranked, correct),
                       "NDCG": O.ndcg_compute(ranked, correct), "MRR": O.reciprocal_rank(ranked, correct),
                       "prec@5": O.precision_at(ranked, correct, 5), "recall@5": O.recall_at(ranked, correct, 5)}
    with open(os.path.join(HERE, "oracle_config1.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", os.path.join(HERE, "oracle_config1.json"))


if __name__ == "__main__":
    main()
