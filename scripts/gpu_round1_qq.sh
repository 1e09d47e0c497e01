#!/bin/bash
# phase trace of the plain run (no ncu) with the reordered epilogue
mkdir -p gpurun_out
MMLB200_TRACE=1 timeout 200 python scripts/topn_filter_ab.py 4 > gpurun_out/qq_topn_trace.log 2>&1; echo "rc=$?" >> gpurun_out/qq_topn_trace.log
grep -v "batch H2D\|batch D2H" gpurun_out/qq_topn_trace.log | tr '\n' ';' | sed 's/\[mmlb200 topn\] //g; s/  */ /g'
