#!/bin/bash
# kernel-only durations (ncu, immune to host jitter) of the reordered epilogue: TF32 rep then BF16 rep
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"score_select|finalize|ignore_sorted" -c 32 --csv --log-file gpurun_out/pp_topn.csv python scripts/topn_filter_ab.py 2 > gpurun_out/pp_topn.log 2>&1
cat gpurun_out/pp_topn.log | tail -3
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/pp_topn.csv')):
    if len(r) > 5 and r[0].isdigit():
        print(r[4][:32], r[-1], r[-2])
PY
