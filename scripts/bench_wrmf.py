#!/usr/bin/env python
"""Times WRMF.Iterate() (config 3 shape by default: 138k users x 27k items, 20M events, k=128).
usage: python scripts/bench_wrmf.py [--users N] [--items M] [--events E] [--k K] [--epochs R]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=138_000)
    ap.add_argument("--items", type=int, default=27_000)
    ap.add_argument("--events", type=int, default=20_000_000)
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--epochs", type=int, default=3)
    args = ap.parse_args()
    from mymedialite_b200 import engine, synthetic
    from mymedialite_b200 import dist as mdist
    rank, world, local = mdist.env_rank()
    tdist = None
    if world > 1:        # strong scaling: same problem, rows of each half-sweep sharded over the ranks
        import torch
        import torch.distributed as tdist
        torch.cuda.set_device(local)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = mdist.create_context(local)
    t0 = time.time()
    u, i = synthetic.implicit_cuda(args.users, args.items, args.events, 20260103, device="cuda:%d" % local)
    gen_s = time.time() - t0
    t0 = time.time()
    f = engine.DeviceFeedback(ctx, u, i, max_user=args.users - 1, max_item=args.items - 1)
    m = engine.WrmfModel(ctx, f, args.k)
    m.init_model(1)
    ctx.synchronize()
    build_s = time.time() - t0
    ms = []
    for _ in range(args.epochs):
        ctx.synchronize()
        if tdist is not None:
            tdist.barrier()
        m.iterate()
        x = m.stats()[1]
        if tdist is not None:
            import torch
            t = torch.tensor([x], dtype=torch.float64, device="cuda")
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            x = float(t.item())
        ms.append(x)
    nnz = f.nnz
    flop = 2.0 * args.k * args.k * (2.0 * nnz + args.users + args.items)
    if rank == 0:
        print(json.dumps({"shape": {"users": args.users, "items": args.items, "events": int(u.size), "nnz": int(nnz), "k": args.k},
                          "n_gpus": world, "epoch_ms": [round(x, 2) for x in ms], "gen_s": round(gen_s, 1), "build_s": round(build_s, 2),
                          "user_ranges": m.shard(False).tolist(), "item_ranges": m.shard(True).tolist(),
                          "tflops_algorithmic": flop / (min(ms) * 1e-3) / 1e12}), flush=True)
    if tdist is not None:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
