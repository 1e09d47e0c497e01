"""BoldDriver learn-rate rule on the device (BiasedMatrixFactorization.cs:225-244, 515-552) next to the oracle: the
objective is recomputed after every epoch, the learn rate halves when it grew and gains 5 % when it shrank; the first
comparison is against the loss InitModel computed before the rating scale and the global bias were set (:161-170 vs :186-190).
(Sorted last on purpose: added at the very end of round 1, after the GPU budget was spent -- first run is the driver's.)"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("learn_rate", [0.01, 0.6])
def test_bold_driver_learnrate_sequence_matches_oracle(learn_rate):
    from mymedialite_b200 import engine, synthetic
    d = synthetic.ratings(300, 120, 20000, "half", 13)
    u, i, v = d["train"]
    ctx = engine.Context(0)
    try:
        rng = O.Random(2)
        om = O.Model(u, i, v, biased=True, num_factors=8, bold_driver=1, learn_rate=learn_rate)
        om.init(rng)
        r = engine.DeviceRatings(ctx, u, i, v)
        gm = engine.SgdModel(ctx, r, engine.default_params(biased=1, num_factors=8, bold_driver=1, learn_rate=learn_rate,
                                                            schedule=engine._capi.SCHEDULE_SERIAL))
        gm.set_model(om.user_factors.copy(), om.item_factors.copy())
        om.iterate(rng)
        ri = om.random_index.copy()
        gm.iterate(random_index=ri)
        seq_o, seq_g = [om.learnrate], [gm.learnrate]
        for _ in range(7):
            om.iterate(rng)
            gm.iterate(random_index=ri)
            seq_o.append(om.learnrate); seq_g.append(gm.learnrate)
        assert seq_g == seq_o, (seq_g, seq_o)
        steps = {round(float(b) / float(a), 4) for a, b in zip([float(np.float32(learn_rate))] + seq_o[:-1], seq_o)}
        assert steps <= {0.5, 1.05} and (learn_rate < 0.1 or 0.5 in steps)    # the large rate overshoots at least once
        assert gm.objective() == pytest.approx(om.objective(), rel=1e-4)
    finally:
        ctx.close()
