#!/bin/bash
# one ncu --set full capture of the collecting pass (score_select_kernel<1>) after the TMEM hand-back reorder, config 5 batch
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:score_select -s 1 -c 1 -f -o gpurun_out/vv_prof_topn python scripts/topn_filter_ab.py 1 > gpurun_out/vv_ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out/vv_prof_topn.ncu-rep
