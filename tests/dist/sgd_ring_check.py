"""Run under torchrun with >= 2 GPUs: multi-GPU DSGD (users sharded, item blocks rotating round the NCCL ring)
against the CPU oracle.

  torchrun --nproc-per-node 2 tests/dist/sgd_ring_check.py

Deterministic part: with the conflict-free `rounds` mode the R-GPU epoch equals a serial pass in the order
(GPU-level sub-epoch S, rank, that rank's dumped schedule of the item block it holds in S); the oracle replays it.
Statistical part: default async mode, per-epoch RMSE within 0.5 % of the oracle's single-threaded run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mymedialite_b200 import engine, synthetic  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = torch.from_numpy(engine.Context.unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
    dist.broadcast(t, 0)
    ctx = engine.Context(local, rank, world, t.cpu().numpy())
    C = engine._capi

    d = synthetic.ratings(1200, 300, 90000, "half", 21)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    mine, tmine = (u % world) == rank, (tu % world) == rank
    k = 32
    om = O.Model(u, i, v, biased=True, num_factors=k)
    om.init(O.Random(1))
    U0, V0 = om.user_factors.copy(), om.item_factors.copy()
    r = engine.DeviceRatings(ctx, u[mine], i[mine], v[mine], max_user=int(u.max()), max_item=int(i.max()))
    local_idx = np.flatnonzero(mine)

    # ---- deterministic replay
    G = 5
    gm = engine.SgdModel(ctx, r, engine.default_params(num_factors=k, num_groups=G, num_subgroups=2,
                                                       intra_block=C.INTRA_ROUNDS, hot_item_factor=0.0))
    gm.set_model(U0, V0)
    assert abs(gm.get_model(False, False)["global_bias"] - om.global_bias) < 1e-6
    for epoch in range(2):
        order, block, _ = gm.schedule(detail=True)
        gm.iterate()
        S_of = block // (G * G)
        parts = [None] * world
        dist.all_gather_object(parts, (local_idx[order], S_of))
        n_sub = max(int(s_of.max()) for _, s_of in parts if s_of.size) + 1   # world x item blocks per rank
        for S in range(n_sub):
            for rk in range(world):
                idx, s_of = parts[rk]
                om.iterate_indices(idx[s_of == S].astype(np.int32))
    g = gm.get_model()
    # user rows live on their rank, item rows are synchronised on every rank
    lu = np.unique(u[mine])
    err_u = np.abs(g["U"][lu] - om.user_factors[lu]).max()
    err_v = np.abs(g["V"] - om.item_factors).max()
    err_b = max(np.abs(g["bu"][lu] - om.user_bias[lu]).max(), np.abs(g["bi"] - om.item_bias).max())
    assert max(err_u, err_v, err_b) < 1e-4, (rank, err_u, err_v, err_b)
    ge = gm.evaluate(tu[tmine], ti[tmine], tv[tmine])["RMSE"]
    oe = om.evaluate(tu, ti, tv)["RMSE"]
    assert abs(ge - oe) < 1e-5, (ge, oe)

    # ---- statistical gate, default async mode
    rng = O.Random(1)
    om2 = O.Model(u, i, v, biased=True, num_factors=k)
    om2.init(rng)
    gm2 = engine.SgdModel(ctx, r, engine.default_params(num_factors=k, num_groups=8, num_subgroups=4))
    gm2.set_model(om2.user_factors.copy(), om2.item_factors.copy())
    worst = 0.0
    for epoch in range(6):
        om2.iterate(rng)
        gm2.iterate()
        o_tr, o_te = om2.evaluate(u, i, v)["RMSE"], om2.evaluate(tu, ti, tv)["RMSE"]
        g_tr, g_te = gm2.evaluate_train()["RMSE"], gm2.evaluate(tu[tmine], ti[tmine], tv[tmine])["RMSE"]
        worst = max(worst, abs(g_tr - o_tr) / o_tr, abs(g_te - o_te) / o_te)
    assert worst < 0.005, worst
    dist.barrier()
    if rank == 0:
        print("ring check ok on %d GPUs: replay max err U %.2e V %.2e bias %.2e; async RMSE deviation %.3f %%"
              % (world, err_u, err_v, err_b, 100 * worst))
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
