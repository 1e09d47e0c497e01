#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/topn_filter_ab.py 8 > gpurun_out/mm_topn_ab.log 2>&1; echo "rc=$?" >> gpurun_out/mm_topn_ab.log
cat gpurun_out/mm_topn_ab.log
