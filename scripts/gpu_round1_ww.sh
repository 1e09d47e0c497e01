#!/bin/bash
mkdir -p gpurun_out
LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_runtime/lib:$LD_LIBRARY_PATH timeout 20 tests/cpp/build/host_test gpu tests/golden/example.train tests/golden/example.test gpurun_out > gpurun_out/ww_host_test.log 2>&1; echo "rc=$?" >> gpurun_out/ww_host_test.log
tail -n 12 gpurun_out/ww_host_test.log
