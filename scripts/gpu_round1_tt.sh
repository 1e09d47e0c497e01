#!/bin/bash
# per-context locking of the ABI: concurrency tests + the whole GPU suite once more
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/tt_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/tt_pytest.log
tail -n 8 gpurun_out/tt_pytest.log
timeout 200 python bench.py --steps 3 --no-cpu > gpurun_out/tt_bench.log 2>&1; echo "rc=$?" >> gpurun_out/tt_bench.log
tail -c 600 gpurun_out/tt_bench.log
