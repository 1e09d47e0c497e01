#!/bin/bash
# item-ranking evaluation on the device: new GPU tests
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_itemeval_gpu.py -m gpu -q -x > gpurun_out/gg_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/gg_pytest.log
tail -40 gpurun_out/gg_pytest.log
