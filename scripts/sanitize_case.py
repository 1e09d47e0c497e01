#!/usr/bin/env python
"""The small cases compute-sanitizer runs over (one tool per gpurun call: scripts/gpu_r2.sh <tag> memcheck | racecheck):
config 1 in the reference's serial order, the DSGD epoch kernels (rounds and lock-free, several CTAs per group), the
NaiveParallelization kernel, Evaluate / Predict / objective, fold-in, a WRMF epoch on the tcgen05 Gram path with both row
solvers, the tcgen05 top-N and the exact one, the rating-matrix build primitives."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mymedialite_b200 import engine, synthetic, _capi     # noqa: E402


def main():
    ctx = engine.Context(0)
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    tr = np.loadtxt(os.path.join(g, "example.train"))
    u, i, v = tr[:, 0].astype(np.int32), tr[:, 1].astype(np.int32), tr[:, 2].astype(np.float32)
    r = engine.DeviceRatings(ctx, u, i, v)
    m = engine.SgdModel(ctx, r, engine.default_params(num_factors=10, schedule=_capi.SCHEDULE_SERIAL))
    m.init_model(1)
    for _ in range(3):
        m.iterate(random_index=np.arange(u.size, dtype=np.int32)[::-1].copy())
    print("config 1 serial", m.evaluate_train())
    d = synthetic.ratings(600, 200, 40000, "half", 9)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    r = engine.DeviceRatings(ctx, u, i, v)
    r.csr(); r.csr(True); r.partition_blocks(np.arange(600), np.arange(200), 4)
    engine.partition_indices(ctx, np.arange(u.size, dtype=np.int32), 37)
    for kw in (dict(num_groups=4, num_subgroups=4, intra_block=_capi.INTRA_ROUNDS, hot_item_factor=0.5, hot_copies=4),
               dict(num_groups=4, ctas_per_group=3, num_subgroups=4), dict(num_groups=1, ctas_per_group=6, num_subgroups=2),
               dict(schedule=_capi.SCHEDULE_NAIVE, max_threads=8), dict(num_groups=3, ctas_per_group=2, persistent=0)):
        for k in (10, 64, 128):
            m = engine.SgdModel(ctx, r, engine.default_params(num_factors=k, **kw))
            m.init_model(3)
            for _ in range(2):
                if kw.get("schedule") == _capi.SCHEDULE_NAIVE:
                    m.iterate(random_index=np.random.RandomState(1).permutation(u.size).astype(np.int32))
                else:
                    m.iterate()
            out = (m.evaluate(tu, ti, tv)["RMSE"], m.evaluate_train()["RMSE"], m.objective(), float(m.predict(tu[:7], ti[:7])[0]))
            m.fold_in([[1, 2, 3]], [[3.0, 4.0, 1.5]], np.zeros((1, k), np.float32), 3)
            print(kw, k, out)
            m.close()
    rs = np.random.RandomState(3)
    nu, ni = 900, 400
    fu = rs.randint(0, nu, 30000).astype(np.int32); fi = (rs.zipf(1.4, 30000) % ni).astype(np.int32)
    fb = engine.DeviceFeedback(ctx, fu, fi, max_user=nu - 1, max_item=ni - 1)
    for mode in (_capi.WRMF_TENSOR, _capi.WRMF_TENSOR_PCG, _capi.WRMF_FP64):
        engine.wrmf_set_mode(mode)
        wm = engine.WrmfModel(ctx, fb, 32)
        wm.init_model(2)
        wm.iterate()
        rec = wm.recommend(np.arange(0, nu, 3, dtype=np.int32), 10, None, [[1, 2]] * len(range(0, nu, 3)))
        print("wrmf mode", mode, rec[0][0][:3])
        wm.close()
    engine.wrmf_set_mode(_capi.WRMF_AUTO)
    U, V = (0.1 * rs.randn(nu, 32)).astype(np.float32), (0.1 * rs.randn(ni, 32)).astype(np.float32)
    for mode in (_capi.TOPN_EXACT, _capi.TOPN_TENSOR):
        engine.topn_set_mode(mode)
        print("topn mode", mode, engine.topn_mf(ctx, U, V, np.arange(50, dtype=np.int32), 5)[0][0])
    engine.topn_set_mode(_capi.TOPN_AUTO)
    ctx.close()
    print("OK")


if __name__ == "__main__":
    main()
