"""Host-side plumbing of the multi-GPU path: one process per GPU, torch.distributed for the rendezvous (NCCL
unique-id hand-over, barriers, scalar reductions). The data path itself -- the item-block ring -- is NCCL calls
inside libmmlb200.so (csrc/dist.cu)."""
import os

import numpy as np


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def shard_by_user(users, items, values, rank, world, user_perm=None):
    """The rank's share of a rating set: user block = user_perm[u] % world (MultiCore.cs:64 lifted to GPUs).
    Returns (users, items, values, index) with index = positions in the full arrays."""
    users = np.asarray(users)
    key = users if user_perm is None else np.asarray(user_perm)[users]
    idx = np.flatnonzero(key % world == rank)
    return users[idx], np.asarray(items)[idx], (None if values is None else np.asarray(values)[idx]), idx


def balanced_row_ranges(row_ptr, world, row_cost):
    """Host restatement of the library's WRMF row sharding (csrc/wrmf.cu shard_rows): contiguous row ranges whose cost
    = events + row_cost per row is balanced; rank r solves rows [ranges[r], ranges[r + 1]) of a half-sweep."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = row_ptr.size - 1
    cum = (row_ptr[1:] - row_ptr[0]) + row_cost * np.arange(1, n + 1, dtype=np.int64)
    total = int(cum[-1]) if n else 0
    ranges = np.full(world + 1, n, dtype=np.int32)
    ranges[0] = 0
    for r in range(1, world):
        # first row count whose cumulative cost reaches r / world of the total
        hit = np.flatnonzero(cum * world >= total * r)
        ranges[r] = (int(hit[0]) + 1) if hit.size else n
    return ranges


def broadcast_bytes(buf, src=0):
    """Hands `buf` (uint8 array, valid on rank `src`) to every rank through torch.distributed; works with gloo and
    nccl process groups."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(buf, np.uint8).copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src)
    return t.cpu().numpy()


def merge_schedules(parts, world):
    """Serial-equivalent order of one multi-GPU DSGD epoch from the ranks' dumped schedules: parts[r] =
    (global rating indices in rank r's execution order, GPU-level sub-epoch S of each entry). Blocks of one
    sub-epoch touch disjoint users and items, so (S, rank, rank-local order) is a valid serialisation."""
    out = []
    for S in range(world):
        for r in range(world):
            idx, s_of = parts[r]
            out.append(np.asarray(idx)[np.asarray(s_of) == S])
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def create_context(device=None):
    """mml context for this rank; with WORLD_SIZE > 1 rank 0 creates the NCCL unique id and broadcasts it."""
    from . import engine
    rank, world, local = env_rank()
    if device is None:
        device = local
    if world == 1:
        return engine.Context(device)
    uid = broadcast_bytes(engine.Context.unique_id() if rank == 0 else np.zeros(128, np.uint8))
    return engine.Context(device, rank, world, uid)
