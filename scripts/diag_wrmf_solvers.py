#!/usr/bin/env python
"""WRMF config 3 (138k x 27k, 20M events, k = 128): the preconditioned-CG row solver (mode WRMF_TENSOR_PCG) against the
double-precision Cholesky solver with refinement against the exact operator (mode WRMF_TENSOR_F64, itself within 1e-6 of the
oracle), epoch by epoch: epoch time of both, and the largest relative row error of the CG result
(max over rows of |w_cg - w_ref|_inf / |w_ref|_inf; gate 1e-4) when both start the epoch from the SAME model.
usage: python scripts/diag_wrmf_solvers.py [--epochs E] [--scale S]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    from mymedialite_b200 import engine, synthetic, _capi
    nu, ni, ne, k = int(138_000 * args.scale), int(27_000 * args.scale), int(20_000_000 * args.scale), 128
    u, i = synthetic.implicit_cuda(nu, ni, ne, 20260103)
    ctx = engine.Context(0)
    fb = engine.DeviceFeedback(ctx, u, i, max_user=nu - 1, max_item=ni - 1)
    cg = engine.WrmfModel(ctx, fb, k)
    ref = engine.WrmfModel(ctx, fb, k)
    cg.init_model(1)
    U, V = cg.get_model()
    for e in range(args.epochs):
        ref.set_model(U, V)            # both solvers start every epoch from the CG run's model
        engine.wrmf_set_mode(_capi.WRMF_TENSOR_F64)
        ref.iterate(); ref_ms = ref.stats()[1]
        engine.wrmf_set_mode(_capi.WRMF_TENSOR_PCG)
        cg.iterate(); cg_ms = cg.stats()[1]
        U, V = cg.get_model()
        rU, rV = ref.get_model()
        # the item half-sweep of the reference used ITS user rows; compare the user side exactly and the item side as a bound
        eu = np.abs(U - rU).max(axis=1) / np.maximum(np.abs(rU).max(axis=1), 1e-30)
        ev = np.abs(V - rV).max(axis=1) / np.maximum(np.abs(rV).max(axis=1), 1e-30)
        print(json.dumps({"epoch": e + 1, "cg_ms": round(cg_ms, 2), "cholesky_f64_ms": round(ref_ms, 2),
                          "max_rel_row_err_users": float(eu.max()), "max_rel_row_err_items": float(ev.max()),
                          "p99_users": float(np.quantile(eu, 0.99)), "p99_items": float(np.quantile(ev, 0.99)),
                          "max_abs_U": float(np.abs(U).max())}), flush=True)


if __name__ == "__main__":
    main()
