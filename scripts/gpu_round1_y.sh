#!/bin/bash
# ncu --set full of the WRMF solve kernel: a mid user batch (launch 4) and the heaviest item batch (launch 9)
mkdir -p gpurun_out
timeout 600 python scripts/bench_wrmf.py --epochs 1 > gpurun_out/y_plain.log 2>&1 || exit 1
for s in 4 9; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wrmf_solve_kernel -s $s -c 1 -f -o gpurun_out/y_prof_solve_$s python scripts/bench_wrmf.py --epochs 1 > gpurun_out/y_ncu_$s.log 2>&1; echo "rc=$?" >> gpurun_out/y_ncu_$s.log
done
echo done
