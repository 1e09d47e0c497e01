#!/bin/bash
mkdir -p gpurun_out
S="python scripts/sweep_groups.py --workload nf_sub8 --epochs 4"
timeout 300 $S --shapes 18x8,1x148 --pop-offset 3000 > gpurun_out/v_sub8_flat.log 2>&1; echo "rc=$?" >> gpurun_out/v_sub8_flat.log
timeout 300 $S --shapes 18x8,1x148 --pop-offset 0 > gpurun_out/v_sub8_zipf.log 2>&1; echo "rc=$?" >> gpurun_out/v_sub8_zipf.log
timeout 300 $S --shapes 18x8,37x4,1x148 --subgroups 8 > gpurun_out/v_sub8_w8.log 2>&1; echo "rc=$?" >> gpurun_out/v_sub8_w8.log
timeout 300 python scripts/sweep_groups.py --workload netflix --epochs 4 --shapes 37x4 --subgroups 8 > gpurun_out/v_nf_w8.log 2>&1; echo "rc=$?" >> gpurun_out/v_nf_w8.log
echo done
