#!/bin/bash
# full GPU suite + smoke + default bench (with the rmse_vs_ref check) + reference arm
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/hh_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/hh_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/hh_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/hh_smoke.log
timeout 200 python bench.py --workload tiny --steps 2 --cpu-sample 300000 > gpurun_out/hh_bench_tiny.log 2>&1; echo "rc=$?" >> gpurun_out/hh_bench_tiny.log
( time timeout 500 python bench.py ) > gpurun_out/hh_bench.log 2>&1; echo "rc=$?" >> gpurun_out/hh_bench.log
( time timeout 400 python bench.py --impl reference ) > gpurun_out/hh_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/hh_bench_ref.log
tail -5 gpurun_out/hh_pytest.log gpurun_out/hh_smoke.log gpurun_out/hh_bench_tiny.log
tail -c 1500 gpurun_out/hh_bench.log
