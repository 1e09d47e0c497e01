#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do
MMLB200_TC_DBG=$d MMLB200_TRACE=1 timeout 300 python scripts/bench_topn.py --users 262144 --reps 2 --exact-users 256 > gpurun_out/h_dbg$d.log 2>&1; echo "rc=$?" >> gpurun_out/h_dbg$d.log
done
echo done
