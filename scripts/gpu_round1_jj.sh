#!/bin/bash
# BF16 filter for the top-N GEMM: parity tests under both filters, then config 5 with each
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_topn_tc_gpu.py tests/test_wrmf_topn_gpu.py -m gpu -q -x > gpurun_out/jj_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/jj_pytest.log
tail -n 15 gpurun_out/jj_pytest.log
MMLB200_TRACE=1 timeout 200 python scripts/bench_topn.py --reps 2 --exact-users 512 > gpurun_out/jj_topn_bf16.log 2>&1; echo "rc=$?" >> gpurun_out/jj_topn_bf16.log
MMLB200_TC_FILTER=tf32 timeout 200 python scripts/bench_topn.py --reps 2 --exact-users 512 > gpurun_out/jj_topn_tf32.log 2>&1; echo "rc=$?" >> gpurun_out/jj_topn_tf32.log
tail -n 12 gpurun_out/jj_topn_bf16.log; tail -n 2 gpurun_out/jj_topn_tf32.log
