#!/usr/bin/env python
"""Sweep of the DSGD grid shape (worker groups x CTAs per group) on one GPU: epoch time and per-epoch RMSE, next to
the oracle's single-threaded run on the same data and initial model (the north-star RMSE gate, 0.5 %).

  python scripts/sweep_groups.py --workload netflix10 --epochs 8 --oracle 1 --shapes 148x1,37x4,9x16,1x148
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="netflix10")
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("--oracle", type=int, default=0)
    ap.add_argument("--shapes", default="148x1,74x2,37x4,18x8,9x16,4x37,2x74,1x148")
    ap.add_argument("--subgroups", type=int, default=16)
    ap.add_argument("--pop-offset", type=float, default=None)
    ap.add_argument("--variants", default="", help="comma list of MMLB200_SGD_VARIANT values; every shape runs under each")
    ap.add_argument("--biased", type=int, default=1, help="0: plain MatrixFactorization (no bias terms)")
    ap.add_argument("--naive", type=int, default=0, help="1: also time the NaiveParallelization list schedule (shape name 'naive')")
    args = ap.parse_args()
    from mymedialite_b200 import engine
    d, k, desc = bench.make_data(args.workload, 0, 1, args.pop_offset)
    u, i, v = d["train"]; tu, ti, tv = d["test"]
    n = int(u.size)
    ctx = engine.Context(0)
    ratings = engine.DeviceRatings(ctx, u, i, v, max_user=d["n_users"] - 1, max_item=d["n_items"] - 1)

    oracle_out = {}
    U0 = V0 = None
    th = None
    if args.oracle:
        from oracle import oracle as O
        om = O.Model(u, i, v, biased=True, num_factors=k, max_user=d["n_users"] - 1, max_item=d["n_items"] - 1)
        rng = O.Random(1)
        om.init(rng)
        U0, V0 = om.user_factors.copy(), om.item_factors.copy()

        def run_oracle():
            t0 = time.time()
            tr, te = [], []
            for _ in range(args.epochs):
                om.iterate(rng)
                tr.append(om.evaluate(u, i, v)["RMSE"]); te.append(om.evaluate(tu, ti, tv)["RMSE"])
            oracle_out.update(train=tr, test=te, seconds=time.time() - t0)

        th = threading.Thread(target=run_oracle)
        th.start()

    results = []
    variants = [v for v in args.variants.split(",") if v != ""] or [None]
    ri = np.random.RandomState(7).permutation(n).astype(np.int32) if args.naive else None
    for variant, shape in [(v, s) for v in variants for s in args.shapes.split(",")] + ([(None, "naive")] if args.naive else []):
        naive = shape == "naive"
        G, cpg = (0, 0) if naive else (int(x) for x in shape.split("x"))
        if variant is not None:
            os.environ["MMLB200_SGD_VARIANT"] = variant
            shape = "v%s:%s" % (variant, shape)
        t0 = time.time()
        params = engine.default_params(biased=args.biased, num_factors=k, num_groups=G, ctas_per_group=cpg, num_subgroups=args.subgroups)
        if naive:
            params = engine.default_params(biased=1, num_factors=k, schedule=engine._capi.SCHEDULE_NAIVE, max_threads=8)
        try:
            model = engine.SgdModel(ctx, ratings, params)
        except Exception as ex:          # e.g. a grid the variant's register budget cannot keep co-resident
            print(json.dumps({"shape": shape, "error": str(ex)}), flush=True)
            continue
        if U0 is not None:
            model.set_model(U0, V0)
        else:
            model.init_model(1, 0.0, 0.1)
        ctx.synchronize()
        build_s = time.time() - t0
        rs = np.random.RandomState(1)
        ms, tr, te = [], [], []
        for _ in range(args.epochs):
            ctx.flush_l2()
            if naive:
                model.iterate(random_index=ri)
            else:
                model.iterate(rs.permutation(model.strata_info()["G"]).astype(np.int32))
            ms.append(model.stats()[1])
            tr.append(model.evaluate_train()["RMSE"]); te.append(model.evaluate(tu, ti, tv)["RMSE"])
        r = {"shape": shape, "G": G, "cpg": cpg, "build_s": round(build_s, 2), "ms": [round(x, 3) for x in ms],
             "ms_med": float(np.median(ms)), "gratings_s": n / float(np.median(ms)) / 1e6,
             "train": [round(x, 5) for x in tr], "test": [round(x, 5) for x in te]}
        results.append(r)
        print(json.dumps(r), flush=True)
        model.close()
    if th is not None:
        th.join()
        print(json.dumps({"oracle": oracle_out}), flush=True)
        for r in results:
            dtr = max(abs(a - b) / b for a, b in zip(r["train"], oracle_out["train"]))
            dte = max(abs(a - b) / b for a, b in zip(r["test"], oracle_out["test"]))
            print("%-12s %8.3f ms  max rel dev train %.4f test %.4f" % (r["shape"], r["ms_med"], dtr, dte), flush=True)


if __name__ == "__main__":
    main()
