#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_topn_tc_gpu.py tests/test_wrmf_topn_gpu.py -x -q > gpurun_out/c_pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest_tc.log
timeout 600 python scripts/bench_topn.py --reps 2 > gpurun_out/c_topn_c5.log 2>&1; echo "rc=$?" >> gpurun_out/c_topn_c5.log
timeout 300 python scripts/bench_topn.py --users 200000 --reps 1 --exact-users 1024 > gpurun_out/c_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c_topn_launches.csv python scripts/bench_topn.py --users 200000 --reps 1 --exact-users 1024 > gpurun_out/c_ncu_list.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench_ml10m.log 2>&1; echo "rc=$?" >> gpurun_out/c_bench_ml10m.log
echo done
