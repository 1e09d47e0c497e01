"""Concurrent read-only calls on one handle. The reference invokes Predict / Recommend from TPL threads on a single
recommender object (Eval/Items.cs:147-164, Eval/Ratings.cs via RatingsCrossValidation.cs:60-69), so the C ABI must give the
same answers under concurrency as serially: every entry point takes the context's lock ("Threading").
ctypes releases the GIL during the calls, so these threads really overlap inside the library."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from mymedialite_b200 import engine
    ctx = engine.Context(0)
    yield engine, ctx
    ctx.close()


def run_threads(n, fn):
    errs, out = [], [None] * n

    def work(t):
        try:
            out[t] = fn(t)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(n)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    return out


def test_concurrent_predict_and_evaluate(eng):
    from mymedialite_b200 import synthetic
    engine, ctx = eng
    d = synthetic.ratings(2000, 500, 150000, "half", 4)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    r = engine.DeviceRatings(ctx, u, i, v)
    m = engine.SgdModel(ctx, r, engine.default_params(biased=1, num_factors=32))
    m.init_model(7)
    for _ in range(2):
        m.iterate()
    T = 8
    chunks = np.array_split(np.arange(tu.size), T)
    serial = [m.predict(tu[c], ti[c]) for c in chunks]
    serial_eval = m.evaluate(tu, ti, tv)

    def job(t):
        res = None
        for _ in range(20):
            res = m.predict(tu[chunks[t]], ti[chunks[t]])
            e = m.evaluate(tu, ti, tv)
            assert e == serial_eval
        return res
    got = run_threads(T, job)
    for a, b in zip(got, serial):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_concurrent_recommend_on_one_model(eng):
    engine, ctx = eng
    rs = np.random.RandomState(5)
    nu, ni, k = 4000, 3000, 32
    fu = rs.randint(0, nu, 60000).astype(np.int32); fi = rs.randint(0, ni, 60000).astype(np.int32)
    fb = engine.DeviceFeedback(ctx, fu, fi, max_user=nu - 1, max_item=ni - 1)
    wm = engine.WrmfModel(ctx, fb, k)
    wm.set_model((0.1 * rs.randn(nu, k)).astype(np.float32), (0.1 * rs.randn(ni, k)).astype(np.float32))
    wm.iterate()
    T = 8
    blocks = np.array_split(np.arange(nu, dtype=np.int32), T)
    serial = [wm.recommend(b, 10) for b in blocks]

    def job(t):
        res = None
        for _ in range(5):
            res = wm.recommend(blocks[t], 10)
        return res
    got = run_threads(T, job)
    for g, s in zip(got, serial):
        for (gi, gs), (si, ss) in zip(g, s):
            assert np.array_equal(gi, si) and np.array_equal(gs.view(np.uint32), ss.view(np.uint32))
