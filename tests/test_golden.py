"""Committed golden vectors (tests/golden/oracle_config1.json, written by tests/golden/make_golden.py from the CPU oracle on
the reference's example files). CPU: the oracle still reproduces them bit for bit (drift alarm) and System.Random matches
the public known answers. GPU (-m gpu): the CUDA path reproduces config 1 against the fixed numbers, through the C ABI."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_config1.json")))


def unhex(h):
    return np.array([int(x, 16) for x in h], np.uint32).view(np.float32)


def test_system_random_public_known_answers():
    """First outputs of new System.Random(seed).Next() as published for the .NET Framework algorithm."""
    assert GOLD["system_random"]["0"][:3] == [1559595546, 1755192844, 1649316166]
    assert GOLD["system_random"]["1"][:3] == [534011718, 237820880, 1002897798]
    assert GOLD["system_random"]["42"][:3] == [1434747710, 302596119, 269548474]
    for seed, want in GOLD["system_random"].items():
        r = O.Random(int(seed))
        assert [r.next() for _ in range(len(want))] == want


def test_oracle_reproduces_the_committed_fixture(tmp_path, example_data):
    """make_golden.py run again gives the same file: any change of the oracle's arithmetic must be a reviewed diff."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    u, i, v, tu, ti, tv = example_data
    for name, biased in (("BiasedMatrixFactorization", True), ("MatrixFactorization", False)):
        g = GOLD["config1"][name]
        rng = O.Random(1)
        m = O.Model(u, i, v, biased=biased, num_factors=10, num_iter=30)
        m.init(rng)
        for ep in range(30):
            m.iterate(rng)
            got = [float(np.float32(m.evaluate(u, i, v)["RMSE"])), float(np.float32(m.evaluate(tu, ti, tv)["RMSE"]))]
            assert got == g["rmse_train_test_per_epoch"][ep], (name, ep)
        assert mg.hexlist(m.user_factors[0]) == g["user_factors_row0"] and mg.hexlist(m.item_factors[0]) == g["item_factors_row0"]
        assert mg.hexlist(m.predict_many(tu, ti)) == g["predict_test"]
        assert [int(x) for x in m.random_index] == g["random_index"]
    ms = GOLD["measures"]
    assert O.auc_compute(ms["ranked"], ms["correct"], ms["dropped"]) == ms["AUC"]
    assert O.ndcg_compute(ms["ranked"], ms["correct"]) == ms["NDCG"] and O.ap_compute(ms["ranked"], ms["correct"]) == ms["AP"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,biased", [("BiasedMatrixFactorization", True), ("MatrixFactorization", False)])
def test_cuda_config1_against_the_fixture(example_data, name, biased):
    """Config 1 (example.train / example.test, k=10, 30 epochs, seed 1) on the CUDA path, MaxThreads=1 order: per-epoch
    train / test RMSE, the test predictions and the first factor rows against the committed numbers."""
    from mymedialite_b200 import engine
    u, i, v, tu, ti, tv = example_data
    g = GOLD["config1"][name]
    ctx = engine.Context(0)
    try:
        rng = O.Random(1)
        k = 10
        nu, ni = int(u.max()) + 1, int(i.max()) + 1
        U = rng.init_normal(nu * k).reshape(nu, k); V = rng.init_normal(ni * k).reshape(ni, k)   # InitModel's draw order
        r = engine.DeviceRatings(ctx, u, i, v)
        m = engine.SgdModel(ctx, r, engine.default_params(biased=int(biased), num_factors=k, schedule=engine._capi.SCHEDULE_SERIAL))
        m.set_model(U, V)
        ri = np.array(g["random_index"], np.int32)
        for ep in range(30):
            m.iterate(random_index=ri)
            got = [m.evaluate_train()["RMSE"], m.evaluate(tu, ti, tv)["RMSE"]]
            np.testing.assert_allclose(got, g["rmse_train_test_per_epoch"][ep], rtol=2e-5, atol=2e-5, err_msg="epoch %d" % ep)
        model = m.get_model()
        np.testing.assert_allclose(model["U"][0], unhex(g["user_factors_row0"]), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(model["V"][0], unhex(g["item_factors_row0"]), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(m.predict(tu, ti), unhex(g["predict_test"]), rtol=1e-5, atol=1e-5)
        assert abs(model["global_bias"] - unhex([g["global_bias"]])[0]) < 1e-6
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cuda_wrmf_and_topn_against_the_fixture(example_data):
    from mymedialite_b200 import engine
    u, i, v, tu, ti, tv = example_data
    g = GOLD["wrmf_k4_3epochs"]
    ctx = engine.Context(0)
    try:
        rng = O.Random(1)
        nu, ni, k = int(u.max()) + 1, int(i.max()) + 1, 4
        U = rng.init_normal(nu * k).reshape(nu, k); V = rng.init_normal(ni * k).reshape(ni, k)
        fb = engine.DeviceFeedback(ctx, u, i, max_user=nu - 1, max_item=ni - 1)
        m = engine.WrmfModel(ctx, fb, k, 1.0, 0.015)
        m.set_model(U, V)
        for _ in range(3):
            m.iterate()
        Ug, Vg = m.get_model()
        Uw, Vw = unhex(g["U"]).reshape(nu, k), unhex(g["V"]).reshape(ni, k)
        for got, want in ((Ug, Uw), (Vg, Vw)):
            scale = np.abs(want).max(axis=1, keepdims=True) + 1e-12
            assert (np.abs(got - want) / scale).max() < 1e-4
        # top-2 on the FIXTURE's factors (identical inputs -> identical lists and score bits)
        users = np.arange(nu, dtype=np.int32)
        ign = [i[u == x] for x in users]
        res = engine.topn_mf(ctx, Uw, Vw, users, 2, None, ign)
        for x in users:
            w = g["top2_ignoring_training_items"][str(int(x))]
            assert res[x][0].tolist() == w["items"]
            assert np.array_equal(res[x][1].view(np.uint32), unhex(w["scores"]).view(np.uint32))
    finally:
        ctx.close()
