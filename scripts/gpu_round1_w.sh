#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgd_gpu.py tests/test_wrmf_topn_gpu.py -x -q > gpurun_out/w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/w_pytest.log
timeout 600 python scripts/sweep_groups.py --workload netflix --epochs 5 --shapes 37x4,18x8,148x1,1x148 > gpurun_out/w_sweep_nf.log 2>&1; echo "rc=$?" >> gpurun_out/w_sweep_nf.log
timeout 600 python scripts/sweep_groups.py --workload nf_sub8 --epochs 4 --shapes 37x4,18x8,1x148 > gpurun_out/w_sweep_sub8.log 2>&1; echo "rc=$?" >> gpurun_out/w_sweep_sub8.log
timeout 600 python scripts/sweep_groups.py --workload ml10m --epochs 5 --shapes 37x4,9x16 > gpurun_out/w_sweep_ml10m.log 2>&1; echo "rc=$?" >> gpurun_out/w_sweep_ml10m.log
timeout 600 python scripts/bench_wrmf.py --epochs 3 > gpurun_out/w_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/w_wrmf_c3.log
echo done
