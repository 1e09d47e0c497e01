#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/ee_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ee_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ee_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/ee_smoke.log
timeout 600 python bench.py > gpurun_out/ee_bench.log 2>&1; echo "rc=$?" >> gpurun_out/ee_bench.log
echo done
