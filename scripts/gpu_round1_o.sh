#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgd_gpu.py tests/test_ratings_gpu.py -x -q > gpurun_out/o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/o_pytest.log
timeout 600 python bench.py > gpurun_out/o_bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/o_bench_default.log
timeout 600 python bench.py --no-cpu --pop-offset 0 > gpurun_out/o_bench_zipf.log 2>&1; echo "rc=$?" >> gpurun_out/o_bench_zipf.log
timeout 600 python bench.py --no-cpu --workload ml10m --steps 10 > gpurun_out/o_bench_ml10m.log 2>&1; echo "rc=$?" >> gpurun_out/o_bench_ml10m.log
echo done
