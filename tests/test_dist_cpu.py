"""World-size-2 gloo run of the host-side multi-GPU plumbing (rendezvous bytes, sharding, schedule merge)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as tdist
    from mymedialite_b200 import dist
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        secret = np.arange(128, dtype=np.uint8) * 3 if rank == 0 else np.zeros(128, np.uint8)
        got = dist.broadcast_bytes(secret, 0)                      # the NCCL unique id travels this way
        rs = np.random.RandomState(5)
        u = rs.randint(0, 300, 4000); i = rs.randint(0, 40, 4000); v = rs.rand(4000).astype(np.float32)
        su, si, sv, idx = dist.shard_by_user(u, i, v, rank, world)
        # every rank reports its local "schedule": here simply its shard, split in two GPU-level sub-epochs by item block
        parts = [None] * world
        tdist.all_gather_object(parts, (idx, (si % world - rank) % world))
        order = dist.merge_schedules(parts, world)
        q.put((rank, got.tolist(), int(su.size), sorted(order.tolist()) == list(range(4000)),
               bool(np.all(u[idx] % world == rank))))
    finally:
        tdist.destroy_process_group()


def test_balanced_row_ranges_cover_and_balance():
    """WRMF row sharding rule (mirrors csrc/wrmf.cu): contiguous, covering, cost-balanced ranges."""
    sys.path.insert(0, ROOT)
    from mymedialite_b200 import dist
    rs = np.random.RandomState(3)
    nnz = (rs.zipf(1.6, 5000) % 400).astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(nnz)])
    for world in (1, 2, 3, 8):
        r = dist.balanced_row_ranges(ptr, world, 64)
        assert r[0] == 0 and r[-1] == 5000 and np.all(np.diff(r) >= 0)
        cost = np.array([ptr[r[t + 1]] - ptr[r[t]] + 64 * (r[t + 1] - r[t]) for t in range(world)])
        assert cost.max() <= cost.mean() + nnz.max() + 64          # no rank exceeds its share by more than one row
    assert list(dist.balanced_row_ranges(np.array([0, 0, 0, 0]), 2, 1)) == [0, 2, 3]


@pytest.mark.timeout(180)
def test_two_rank_gloo_plumbing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    want = (np.arange(128, dtype=np.uint8) * 3).tolist()
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, got, n_local, covers, own in res:
        assert got == want and covers and own and n_local > 0
    assert sum(r[2] for r in res) == 4000
