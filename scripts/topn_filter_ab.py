#!/usr/bin/env python
"""A/B of the top-N filter precision inside ONE process (config 5 shape): alternates TF32 and BF16 and prints the device
time of the tensor path, the users sent to the exact path and the wall time of each call."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mymedialite_b200 import engine  # noqa: E402

users_n, items_n, k, n, ign = 1_000_000, 100_000, 128, 10, 20
ctx = engine.Context(0)
rs = np.random.default_rng(20260105)
V = rs.standard_normal((items_n, k), dtype=np.float32) * np.float32(0.1)
U = rs.standard_normal((users_n, k), dtype=np.float32) * np.float32(0.1)
users = np.arange(users_n, dtype=np.int32)
ign_idx = rs.integers(0, items_n, (users_n, ign), dtype=np.int32).reshape(-1)
ign_ptr = np.arange(users_n + 1, dtype=np.int64) * ign
oi = np.zeros((users_n, n), np.int32); os_ = np.zeros((users_n, n), np.float32); oc = np.zeros(users_n, np.int32)
out = []
ref = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    kind = rep % 2
    engine.topn_set_filter(engine._capi.TOPN_FILTER_TF32 if kind == 0 else engine._capi.TOPN_FILTER_BF16)
    t0 = time.time()
    engine.check(ctx.lib.mml_topn_mf(ctx.h, U, users_n, V, items_n, k, users, users_n, n, None, items_n, ign_ptr, ign_idx, oi, os_, oc))
    wall = time.time() - t0
    st = engine.topn_last_stats()
    if ref is None:
        ref = (oi.copy(), os_.copy())
    same = bool(np.array_equal(ref[0], oi) and np.array_equal(ref[1].view(np.uint32), os_.view(np.uint32)))
    out.append({"filter": "tf32" if kind == 0 else "bf16", "tensor_path_ms": round(st["tensor_path_ms"], 1),
                "exact_users": st["users_exact_path"], "wall_s": round(wall, 3), "same_bits_as_first": same})
    print(json.dumps(out[-1]), flush=True)
