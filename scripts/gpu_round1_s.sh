#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgd_gpu.py -x -q > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
timeout 900 python scripts/sweep_groups.py --workload netflix10 --epochs 8 --oracle 1 > gpurun_out/s_sweep_nf10.log 2>&1; echo "rc=$?" >> gpurun_out/s_sweep_nf10.log
timeout 600 python scripts/sweep_groups.py --workload netflix --epochs 6 > gpurun_out/s_sweep_nf.log 2>&1; echo "rc=$?" >> gpurun_out/s_sweep_nf.log
timeout 600 python scripts/sweep_groups.py --workload ml10m --epochs 6 --shapes 148x1,37x4,9x16,1x148 > gpurun_out/s_sweep_ml10m.log 2>&1; echo "rc=$?" >> gpurun_out/s_sweep_ml10m.log
echo done
