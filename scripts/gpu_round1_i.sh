#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,memory.used --format=csv > gpurun_out/i_smi.log 2>&1
nvidia-smi --query-compute-apps=pid,used_memory --format=csv >> gpurun_out/i_smi.log 2>&1
timeout 900 python -m pytest tests/test_wrmf_topn_gpu.py -x -q > gpurun_out/i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/i_pytest.log
for d in 0 3 2 1 0; do
MMLB200_TC_DBG=$d MMLB200_TRACE=1 timeout 300 python scripts/bench_topn.py --users 151552 --reps 4 --exact-users 256 2>&1 | grep -E "batch kernels|shape" | cut -c1-200 > gpurun_out/i_dbg$d.log
nvidia-smi --query-gpu=clocks.sm,power.draw,utilization.gpu --format=csv,noheader >> gpurun_out/i_smi.log 2>&1
done
timeout 900 python scripts/bench_wrmf.py --epochs 2 > gpurun_out/i_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/i_wrmf_c3.log
echo done
