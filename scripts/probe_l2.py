#!/usr/bin/env python
"""L2 service rate of this GPU under the SGD epoch kernel's item-row access pattern (mml_ctx_probe_l2): rows/s and GB/s of
row reads, row atomics, and read + atomic per row, on tables the size of the item matrix of the bench workloads.

  python scripts/probe_l2.py            # one JSON line per (table, mode)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from mymedialite_b200 import engine
    ctx = engine.Context(0)
    for name, rows, k in (("netflix items", 17_770, 128), ("ring sub-epoch items", 2_225, 128), ("ml10m items", 10_681, 64),
                          ("8 x netflix items", 142_160, 128)):
        for mode, what in ((0, "read"), (1, "red.add.v4"), (2, "read + red.add.v4")):
            r = ctx.probe_l2(mode, rows, k, 5)
            touched = k * 4 * (2 if mode == 2 else 1)
            print(json.dumps({"table": name, "rows": rows, "row_bytes": k * 4, "mode": what, "rows_per_s": r,
                              "GB_per_s": r * touched / 1e9}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
