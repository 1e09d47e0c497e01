#!/bin/bash
# config 5 top-N alone on a fresh box (the run after other GPU processes is slowed by host/driver clean-up)
mkdir -p gpurun_out
timeout 200 python scripts/bench_topn.py --reps 5 --exact-users 1024 > gpurun_out/ss_topn_c5.log 2>&1; echo "rc=$?" >> gpurun_out/ss_topn_c5.log
tail -n 2 gpurun_out/ss_topn_c5.log
