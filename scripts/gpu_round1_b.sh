#!/bin/bash
# GPU call B: tcgen05 top-N parity tests, config-5 / config-3 timings, e2e re-check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_topn_tc_gpu.py -x -q > gpurun_out/b_pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_tc.log
timeout 600 python -m pytest tests/test_wrmf_topn_gpu.py -x -q > gpurun_out/b_pytest_topn.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_topn.log
timeout 300 python scripts/bench_topn.py --users 100000 --items 100000 --reps 2 > gpurun_out/b_topn_100k.log 2>&1; echo "rc=$?" >> gpurun_out/b_topn_100k.log
timeout 600 python scripts/bench_topn.py --reps 2 > gpurun_out/b_topn_c5.log 2>&1; echo "rc=$?" >> gpurun_out/b_topn_c5.log
timeout 900 python scripts/bench_wrmf.py --epochs 2 > gpurun_out/b_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/b_wrmf_c3.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench_ml10m.log 2>&1; echo "rc=$?" >> gpurun_out/b_bench_ml10m.log
echo done
