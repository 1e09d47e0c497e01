#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_topn.py --users 151552 --reps 2 --exact-users 256 > gpurun_out/e_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -s 2 -c 2 -f -o gpurun_out/e_prof_topn python scripts/bench_topn.py --users 151552 --reps 2 --exact-users 256 > gpurun_out/e_ncu_full.log 2>&1
echo done
