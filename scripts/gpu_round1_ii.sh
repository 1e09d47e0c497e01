#!/bin/bash
# smoke after the fix + WRMF retrain test, then a phase trace and an ncu launch list of config 5 top-N
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/ii_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/ii_smoke.log
MMLB200_TRACE=1 timeout 300 python scripts/bench_topn.py --reps 2 --exact-users 1024 > gpurun_out/ii_topn_trace.log 2>&1; echo "rc=$?" >> gpurun_out/ii_topn_trace.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mml -c 200 --csv --log-file gpurun_out/ii_topn_launches.csv python scripts/bench_topn.py --reps 1 --exact-users 256 > gpurun_out/ii_topn_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/ii_topn_ncu.log
tail -n 6 gpurun_out/ii_smoke.log
tail -n 40 gpurun_out/ii_topn_trace.log
