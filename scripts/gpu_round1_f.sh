#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_topn_tc_gpu.py tests/test_wrmf_topn_gpu.py -x -q > gpurun_out/f_pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest_tc.log
timeout 600 python scripts/bench_topn.py --reps 3 > gpurun_out/f_topn_c5.log 2>&1; echo "rc=$?" >> gpurun_out/f_topn_c5.log
timeout 300 python scripts/bench_topn.py --users 200000 --reps 2 --exact-users 1024 > gpurun_out/f_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_topn_launches.csv python scripts/bench_topn.py --users 200000 --reps 2 --exact-users 1024 > gpurun_out/f_ncu_list.log 2>&1
echo done
