#!/bin/bash
# 1 GPU: whole GPU test suite, smoke, default bench + reference arm, then (after the plain run exited 0) the ncu launch list
# and one --set full capture of the SGD epoch kernel on the same command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/bb_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/bb_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bb_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/bb_smoke.log
timeout 600 python bench.py > gpurun_out/bb_bench.log 2>&1; echo "rc=$?" >> gpurun_out/bb_bench.log
timeout 600 python bench.py --impl reference > gpurun_out/bb_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bb_bench_ref.log
timeout 600 python scripts/bench_wrmf.py --epochs 3 > gpurun_out/bb_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/bb_wrmf_c3.log
timeout 600 python bench.py --workload ml10m --steps 10 --no-cpu > gpurun_out/bb_bench_ml10m.log 2>&1; echo "rc=$?" >> gpurun_out/bb_bench_ml10m.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bb_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mml -c 400 --csv --log-file gpurun_out/bb_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bb_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgd_epoch_kernel -s 3 -c 1 -f -o gpurun_out/bb_prof_sgd python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bb_ncu_full.log 2>&1
echo done
