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
    ctx = engine.Context(0)
    t0 = time.time()
    u, i = synthetic.implicit(args.users, args.items, args.events, 20260103)
    gen_s = time.time() - t0
    t0 = time.time()
    f = engine.DeviceFeedback(ctx, u, i, max_user=args.users - 1, max_item=args.items - 1)
    m = engine.WrmfModel(ctx, f, args.k)
    m.init_model(1)
    ctx.synchronize()
    build_s = time.time() - t0
    ms = []
    for _ in range(args.epochs):
        m.iterate()
        ms.append(m.stats()[1])
    nnz = f.nnz
    flop = 2.0 * args.k * args.k * (2.0 * nnz + args.users + args.items)
    print(json.dumps({"shape": {"users": args.users, "items": args.items, "events": int(u.size), "nnz": int(nnz), "k": args.k},
                      "epoch_ms": [round(x, 2) for x in ms], "gen_s": round(gen_s, 1), "build_s": round(build_s, 2),
                      "tflops_algorithmic": flop / (min(ms) * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
