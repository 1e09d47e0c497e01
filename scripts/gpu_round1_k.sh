#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wrmf_topn_gpu.py -q > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log
timeout 900 python scripts/bench_wrmf.py --epochs 2 > gpurun_out/k_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/k_wrmf_launches.csv python scripts/bench_wrmf.py --epochs 2 > gpurun_out/k_ncu_list.log 2>&1
echo done
