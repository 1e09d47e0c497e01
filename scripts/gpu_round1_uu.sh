#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_golden.py -m gpu -q > gpurun_out/uu_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/uu_pytest.log
tail -n 25 gpurun_out/uu_pytest.log
