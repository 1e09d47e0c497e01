#!/bin/bash
# epilogue hand-back reorder: parity under both filters, then the in-process A/B at config 5
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_topn_tc_gpu.py -m gpu -q -x > gpurun_out/oo_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/oo_pytest.log
tail -n 4 gpurun_out/oo_pytest.log
timeout 200 python scripts/topn_filter_ab.py 8 > gpurun_out/oo_topn_ab.log 2>&1; echo "rc=$?" >> gpurun_out/oo_topn_ab.log
cat gpurun_out/oo_topn_ab.log
