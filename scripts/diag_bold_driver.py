"""Per-epoch divergence of the device's serial BiasedMF epoch from the oracle under the bold driver.
Free run from identical factors; prints learn rate, objective and max |delta| of every model part after every epoch.
  python scripts/diag_bold_driver.py [learn_rate ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mymedialite_b200 import engine, synthetic   # noqa: E402
from oracle import oracle as O                   # noqa: E402


def run(ctx, learn_rate, epochs=8):
    d = synthetic.ratings(300, 120, 20000, "half", 13)
    u, i, v = d["train"]
    rng = O.Random(2)
    om = O.Model(u, i, v, biased=True, num_factors=8, bold_driver=1, learn_rate=learn_rate)
    om.init(rng)
    r = engine.DeviceRatings(ctx, u, i, v)
    gm = engine.SgdModel(ctx, r, engine.default_params(biased=1, num_factors=8, bold_driver=1, learn_rate=learn_rate,
                                                        schedule=engine._capi.SCHEDULE_SERIAL))
    gm.set_model(om.user_factors.copy(), om.item_factors.copy())
    ri = None
    print("learn_rate %g" % learn_rate)
    print("epoch  lr_oracle   lr_device   obj_oracle      obj_device      rel_diff   max|dU|    max|dV|    max|dbu|   max|dbi|   max|U|")
    for e in range(epochs):
        om.iterate(rng)
        if ri is None:
            ri = om.random_index.copy()
        gm.iterate(random_index=ri)
        g = gm.get_model()
        oo, og = om.objective(), gm.objective()
        print("%5d  %.7f   %.7f   %-14.6f  %-14.6f  %.2e   %.2e   %.2e   %.2e   %.2e   %.3g" % (
            e + 1, om.learnrate, gm.learnrate, oo, og, abs(oo - og) / abs(oo),
            np.abs(g["U"] - om.user_factors).max(), np.abs(g["V"] - om.item_factors).max(),
            np.abs(g["bu"] - om.user_bias).max(), np.abs(g["bi"] - om.item_bias).max(), np.abs(om.user_factors).max()))
    gm.close(); r.close()


if __name__ == "__main__":
    ctx = engine.Context(0)
    for lr in [float(x) for x in sys.argv[1:]] or [0.01, 0.6]:
        run(ctx, lr)
    ctx.close()
