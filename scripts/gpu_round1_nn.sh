#!/bin/bash
# where does the collecting pass spend its time? knobs: DBG=1 epilogue does not examine, DBG=2 no MMAs issued, 3 = both
mkdir -p gpurun_out
for d in 0 1 2 3; do
  MMLB200_TC_DBG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:score_select -c 4 --csv --log-file gpurun_out/nn_dbg$d.csv python scripts/topn_filter_ab.py 2 > gpurun_out/nn_dbg$d.log 2>&1
  echo "dbg=$d"; grep score_select gpurun_out/nn_dbg$d.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-60,200-
done
