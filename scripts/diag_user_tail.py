#!/usr/bin/env python
"""How heavy is the heaviest user of a bench workload? (a user's ratings are one sequential chain for the SGD kernels)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

d, k, desc = bench.make_data(sys.argv[1] if len(sys.argv) > 1 else "netflix", 0, 1, None)
c = np.bincount(d["train"][0])
cs = np.sort(c)[::-1]
print(json.dumps({"workload": desc, "ratings": int(c.sum()), "users": int(c.size), "mean": float(c.mean()), "max": int(cs[0]),
                  "top10": [int(x) for x in cs[:10]], "users_over_2000": int((c > 2000).sum()),
                  "share_of_ratings_over_2000": float(c[c > 2000].sum() / c.sum()),
                  "act_clip": os.environ.get("MMLB200_SYN_ACT_CLIP")}))
