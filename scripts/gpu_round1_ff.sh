#!/bin/bash
# new GPU tests of this session only: fold-in / retrain / ingest -> device
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_foldin_gpu.py "tests/test_ratings_gpu.py::test_ingest_feeds_the_device_from_pinned_memory" -m gpu -q > gpurun_out/ff_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ff_pytest.log
tail -30 gpurun_out/ff_pytest.log
