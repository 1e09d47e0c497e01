#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sgd_gpu.py tests/test_wrmf_topn_gpu.py -x -q > gpurun_out/u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/u_pytest.log
timeout 600 python scripts/sweep_groups.py --workload netflix --epochs 5 --shapes 37x4,148x1 --pf 0,4,8,16 > gpurun_out/u_sweep_pf.log 2>&1; echo "rc=$?" >> gpurun_out/u_sweep_pf.log
timeout 600 python scripts/bench_wrmf.py --epochs 3 > gpurun_out/u_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/u_wrmf_c3.log
timeout 600 python bench.py --cpg 4 --no-cpu > gpurun_out/u_bench_cpg4.log 2>&1; echo "rc=$?" >> gpurun_out/u_bench_cpg4.log
echo done
