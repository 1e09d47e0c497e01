#!/bin/bash
# why is the BF16 filter slower in a plain run than its kernels under ncu? clocks / power during the run
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 20 > gpurun_out/ll_smi_bf16.log 2>&1 &
SMI=$!
sleep 1
MMLB200_TRACE=1 timeout 200 python scripts/bench_topn.py --reps 3 --exact-users 256 > gpurun_out/ll_topn_bf16.log 2>&1; echo "rc=$?" >> gpurun_out/ll_topn_bf16.log
kill $SMI
grep "batch kernels" gpurun_out/ll_topn_bf16.log | tr '\n' ' '
awk -F, 'NR>1 {print $2}' gpurun_out/ll_smi_bf16.log | sort | uniq -c | sort -k1 -n -r | head -8
awk -F, 'NR>1 && $2+0 > 1000 {n++; p+=$3} END {print n, p/n}' gpurun_out/ll_smi_bf16.log
