"""Diagnostic: RMSE after E epochs for several schedule settings vs. the oracle's serial run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mymedialite_b200 import engine, synthetic
from oracle import oracle as O

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = synthetic.ratings(3000, 800, 300000, "half", 11)
u, i, v = d["train"]; tu, ti, tv = d["test"]
k = 32
rng = O.Random(1)
om = O.Model(u, i, v, biased=True, num_factors=k)
om.init(rng)
U0, V0 = om.user_factors.copy(), om.item_factors.copy()
for _ in range(E):
    om.iterate(rng)
print("oracle serial        train %.5f test %.5f" % (om.evaluate(u, i, v)["RMSE"], om.evaluate(tu, ti, tv)["RMSE"]))
ctx = engine.Context(0)
r = engine.DeviceRatings(ctx, u, i, v)
C = engine._capi
for name, kw in [
    ("rounds nohot G16", dict(intra_block=C.INTRA_ROUNDS, hot_item_factor=0.0, num_groups=16, num_subgroups=4)),
    ("rounds hot8 G16", dict(intra_block=C.INTRA_ROUNDS, hot_item_factor=1.0, num_groups=16, num_subgroups=4)),
    ("async w=16 G16 hotsum", dict(intra_block=C.INTRA_ASYNC, num_groups=16, num_subgroups=4, hot_merge_average=0)),
    ("async w=1 G16", dict(intra_block=C.INTRA_ASYNC, async_workers=1, num_groups=16, num_subgroups=4)),
    ("async w=4 G16", dict(intra_block=C.INTRA_ASYNC, async_workers=4, num_groups=16, num_subgroups=4)),
    ("async w=16 G16", dict(intra_block=C.INTRA_ASYNC, num_groups=16, num_subgroups=4)),
    ("async w=64 G16", dict(intra_block=C.INTRA_ASYNC, num_groups=16, num_subgroups=16)),
    ("async w=1 G64", dict(intra_block=C.INTRA_ASYNC, async_workers=1, num_groups=64, num_subgroups=4)),
    ("async w=16 G64", dict(intra_block=C.INTRA_ASYNC, num_groups=64, num_subgroups=4)),
    ("async w=16 G148", dict(intra_block=C.INTRA_ASYNC, num_groups=148, num_subgroups=4)),
]:
    gm = engine.SgdModel(ctx, r, engine.default_params(num_factors=k, **kw))
    gm.set_model(U0, V0)
    rs = np.random.RandomState(0)
    for _ in range(E):
        gm.iterate(rs.permutation(gm.strata_info()["G"]).astype(np.int32))
    print("%-20s train %.5f test %.5f" % (name, gm.evaluate_train()["RMSE"], gm.evaluate(tu, ti, tv)["RMSE"]))
