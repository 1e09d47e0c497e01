#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wrmf_topn_gpu.py tests/test_sgd_gpu.py -x -q > gpurun_out/z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest.log
timeout 600 python scripts/bench_wrmf.py --epochs 3 > gpurun_out/z_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/z_wrmf_c3.log
timeout 600 python bench.py > gpurun_out/z_bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/z_bench_default.log
timeout 600 python scripts/bench_topn.py --reps 2 > gpurun_out/z_topn_c5.log 2>&1; echo "rc=$?" >> gpurun_out/z_topn_c5.log
echo done
