#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wrmf_topn_gpu.py tests/test_topn_tc_gpu.py -x -q > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
timeout 900 python scripts/bench_wrmf.py --epochs 2 > gpurun_out/g_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/g_wrmf_c3.log
MMLB200_TRACE=1 timeout 600 python scripts/bench_topn.py --reps 2 > gpurun_out/g_topn_c5.log 2>&1; echo "rc=$?" >> gpurun_out/g_topn_c5.log
timeout 300 python scripts/bench_topn.py --users 151552 --reps 2 --exact-users 256 > gpurun_out/g_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -s 2 -c 2 -f -o gpurun_out/g_prof_topn python scripts/bench_topn.py --users 151552 --reps 2 --exact-users 256 > gpurun_out/g_ncu_full.log 2>&1
echo done
