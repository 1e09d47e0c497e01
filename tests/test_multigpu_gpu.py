"""One process, several GPUs (mml_ctx_create with n_gpus > 1, the NumGpus property of the host classes): every handle
created from such a context is a root over one ordinary handle per GPU, and every entry point fans out to them from one
host thread per GPU (csrc/prims.cu on_ranks). Users are sharded by u % N (MultiCore.cs:64 lifted to GPUs); the item blocks
go round the ring; WRMF solves its rows GPU by GPU and all-gathers them; Recommend() splits the user list.
Skipped on boxes with one GPU (the multi-process path has its own gloo tests in test_dist_cpu.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_one_process_two_gpus_sgd_tracks_one_gpu():
    _two_gpus()
    from mymedialite_b200 import engine, synthetic
    d = synthetic.ratings(3000, 800, 300000, "half", 11)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    rs = np.random.RandomState(5)
    U0 = (0.1 * rs.randn(3000, 32)).astype(np.float32); V0 = (0.1 * rs.randn(800, 32)).astype(np.float32)
    ctx1, ctx2 = engine.Context(0), engine.Context(0, n_gpus=2)
    try:
        runs = []
        for ctx in (ctx1, ctx2):
            r = engine.DeviceRatings(ctx, u, i, v, max_user=2999, max_item=799)
            avg, mn, mx = r.stats()
            assert abs(avg - float(np.float32(v.astype(np.float64).sum()) / np.float32(v.size))) < 1e-6 and (mn, mx) == (v.min(), v.max())
            assert np.array_equal(r.counts(), np.bincount(u, minlength=3000)) and np.array_equal(r.counts(True), np.bincount(i, minlength=800))
            gm = engine.SgdModel(ctx, r, engine.default_params(biased=1, num_factors=32, num_groups=4, ctas_per_group=2, num_subgroups=4))
            gm.set_model(U0, V0)
            g0 = gm.get_model()
            keep_u = np.bincount(u, minlength=3000) > 0
            np.testing.assert_array_equal(g0["U"][keep_u], U0[keep_u])          # the rows come back from the GPU that owns them
            seqs = np.random.RandomState(1)
            rmse = []
            for _ in range(4):
                gm.iterate(seqs.permutation(gm.strata_info()["G"]).astype(np.int32))
                rmse.append((gm.evaluate_train()["RMSE"], gm.evaluate(tu, ti, tv)["RMSE"]))
            g = gm.get_model()
            # Predict / Evaluate agree with the model that get_model returns (every pair answered by the owner of its user)
            pred = gm.predict(tu[:5000], ti[:5000])
            score = g["global_bias"] + g["bu"][tu[:5000]] + g["bi"][ti[:5000]] + np.einsum("nk,nk->n", g["U"][tu[:5000]], g["V"][ti[:5000]])
            want = mn + (mx - mn) / (1.0 + np.exp(-score.astype(np.float64)))
            np.testing.assert_allclose(pred, want, rtol=2e-5, atol=2e-5)
            full = gm.predict(tu, ti)
            assert abs(float(np.sqrt(np.mean((full.astype(np.float64) - tv) ** 2))) - rmse[-1][1]) < 1e-5
            assert np.isfinite(gm.objective()) and gm.stats()[1] > 0
            runs.append(rmse)
            gm.close(); r.close()
        for (a_tr, a_te), (b_tr, b_te) in zip(*runs):
            assert abs(a_tr - b_tr) / a_tr < 0.005 and abs(a_te - b_te) / a_te < 0.005, runs
        assert runs[1][-1][1] < runs[1][0][1]
    finally:
        ctx1.close(); ctx2.close()


def test_one_process_two_gpus_wrmf_and_topn():
    _two_gpus()
    from mymedialite_b200 import engine
    rs = np.random.RandomState(3)
    nu, ni, k = 4000, 900, 32
    fu = rs.randint(0, nu, 120000).astype(np.int32); fi = (rs.zipf(1.3, 120000) % ni).astype(np.int32)
    U0 = (0.1 * rs.randn(nu, k)).astype(np.float32); V0 = (0.1 * rs.randn(ni, k)).astype(np.float32)
    users = np.arange(nu, dtype=np.int32)
    ign = [np.unique(fi[fu == x]) for x in range(nu)]
    ctx1, ctx2 = engine.Context(0), engine.Context(0, n_gpus=2)
    try:
        out = []
        for ctx in (ctx1, ctx2):
            fb = engine.DeviceFeedback(ctx, fu, fi, max_user=nu - 1, max_item=ni - 1)
            wm = engine.WrmfModel(ctx, fb, k)
            wm.set_model(U0, V0)
            wm.iterate(); wm.iterate()
            gU, gV = wm.get_model()
            out.append((gU, gV, wm.recommend(users, 10, None, ign, raw=True), engine.topn_mf(ctx, gU, gV, users[::3], 5, None, ign[::3])))
            wm.close(); fb.close()
        (U1, V1, r1, t1), (U2, V2, r2, t2) = out
        np.testing.assert_allclose(U2, U1, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(V2, V1, rtol=1e-5, atol=1e-6)
        # the lists of the two-GPU context on ITS model: bit-identical to the one-GPU path on the same factors
        ref = engine.topn_mf(ctx1, U2, V2, users, 10, None, ign)
        for b in range(nu):
            kk = int(r2[2][b])
            assert np.array_equal(r2[0][b, :kk], ref[b][0]) and np.array_equal(r2[1][b, :kk].view(np.uint32), ref[b][1].view(np.uint32))
        ref5 = engine.topn_mf(ctx1, U2, V2, users[::3], 5, None, ign[::3])
        for a, b in zip(t2, ref5):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    finally:
        ctx1.close(); ctx2.close()
