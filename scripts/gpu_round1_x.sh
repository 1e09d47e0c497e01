#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgd_gpu.py -x -q > gpurun_out/x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/x_pytest.log
timeout 600 python scripts/sweep_groups.py --workload netflix --epochs 5 --shapes 37x4,18x8,148x1 > gpurun_out/x_sweep_nf.log 2>&1; echo "rc=$?" >> gpurun_out/x_sweep_nf.log
timeout 600 python scripts/sweep_groups.py --workload nf_sub8 --epochs 4 --shapes 37x4,18x8 > gpurun_out/x_sweep_sub8.log 2>&1; echo "rc=$?" >> gpurun_out/x_sweep_sub8.log
echo done
