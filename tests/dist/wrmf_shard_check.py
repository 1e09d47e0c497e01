"""Run under torchrun with >= 2 GPUs: multi-GPU WRMF (rows of each half-sweep sharded over the ranks, solved rows
all-gathered over NCCL) against the CPU oracle's WRMF.Optimize restatement (1e-4 relative, the north-star gate).

  torchrun --nproc-per-node 2 tests/dist/wrmf_shard_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mymedialite_b200 import engine, synthetic  # noqa: E402
from mymedialite_b200 import dist as mdist  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = mdist.create_context(local)
    for k, mode in ((32, engine._capi.WRMF_AUTO), (128, engine._capi.WRMF_AUTO), (20, engine._capi.WRMF_FP64)):
        engine.wrmf_set_mode(mode)
        n_users, n_items = 3000, 2500
        u, i = synthetic.implicit(n_users, n_items, 120000, 31)
        fb = engine.DeviceFeedback(ctx, u, i, max_user=n_users - 1, max_item=n_items - 1)
        m = engine.WrmfModel(ctx, fb, k)
        rs = np.random.RandomState(2)
        U = (0.1 * rs.randn(n_users, k)).astype(np.float32); V = (0.1 * rs.randn(n_items, k)).astype(np.float32)
        m.set_model(U, V)
        ru, ri = m.shard(False), m.shard(True)
        assert ru[0] == 0 and ru[-1] == n_users and np.all(np.diff(ru) > 0), ru
        assert ri[0] == 0 and ri[-1] == n_items and np.all(np.diff(ri) > 0), ri
        # the ranges are the documented rule (events + 8 k per row, contiguous), restated on the host
        uptr, _ = fb.csr(False)
        iptr_, _ = fb.csr(True)
        assert np.array_equal(ru, mdist.balanced_row_ranges(uptr, world, 8 * k)), (ru, mdist.balanced_row_ranges(uptr, world, 8 * k))
        assert np.array_equal(ri, mdist.balanced_row_ranges(iptr_, world, 8 * k))
        # balance: events per rank within 25 % of the mean on this shape
        per = np.diff(uptr[ru])
        assert per.max() < 1.25 * per.mean(), per
        uptr, ucols = O.feedback_csr(u, i, n_users - 1)
        iptr, irows = O.feedback_csr(i, u, n_items - 1)
        for epoch in range(2):
            m.iterate()
            O.wrmf_optimize(uptr, ucols, U, V)      # in place
            O.wrmf_optimize(iptr, irows, V, U)
        gU, gV = m.get_model()
        eu = np.abs(gU - U).max() / np.abs(U).max(); ev = np.abs(gV - V).max() / np.abs(V).max()
        assert eu < 1e-4 and ev < 1e-4, (k, eu, ev)
        # every rank holds the same model
        t = torch.from_numpy(np.concatenate([gU.ravel(), gV.ravel()])).cuda()
        t0 = t.clone()
        dist.broadcast(t0, 0)
        assert torch.equal(t, t0)
        if rank == 0:
            print("wrmf shard check k=%d world=%d: max rel err U %.2e V %.2e, ranges %s %s" % (k, world, eu, ev, ru.tolist(), ri.tolist()), flush=True)
        m.close(); fb.close()
    engine.wrmf_set_mode(engine._capi.WRMF_AUTO)
    dist.barrier()
    if rank == 0:
        print("OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
