#!/bin/bash
mkdir -p gpurun_out
for g in 0 74 37; do
timeout 300 python bench.py --no-cpu --workload nf_sub8 --groups $g > gpurun_out/p_sub8_g$g.log 2>&1; echo "rc=$?" >> gpurun_out/p_sub8_g$g.log
done
timeout 300 python bench.py --no-cpu --workload nf_sub8 --steps 2 --warmup 3 > gpurun_out/p_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_epoch_kernel -s 3 -c 1 -f -o gpurun_out/p_prof_sub8 python bench.py --no-cpu --workload nf_sub8 --steps 2 --warmup 3 > gpurun_out/p_ncu.log 2>&1
echo done
