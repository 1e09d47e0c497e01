#!/bin/bash
# config 5 top-N: BF16 vs TF32 filter, plain runs then the ncu launch list of the BF16 run
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv > gpurun_out/kk_smi.log 2>&1
MMLB200_TC_FILTER=tf32 timeout 200 python scripts/bench_topn.py --reps 3 --exact-users 256 > gpurun_out/kk_topn_tf32.log 2>&1; echo "rc=$?" >> gpurun_out/kk_topn_tf32.log
timeout 200 python scripts/bench_topn.py --reps 3 --exact-users 256 > gpurun_out/kk_topn_bf16.log 2>&1; echo "rc=$?" >> gpurun_out/kk_topn_bf16.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mml -c 120 --csv --log-file gpurun_out/kk_topn_launches.csv python scripts/bench_topn.py --reps 1 --exact-users 256 > gpurun_out/kk_topn_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/kk_topn_ncu.log
cat gpurun_out/kk_smi.log; tail -n 2 gpurun_out/kk_topn_tf32.log; tail -n 2 gpurun_out/kk_topn_bf16.log
