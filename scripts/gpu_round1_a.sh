#!/bin/bash
# one gpurun call: GPU parity tests, bench on the ML-10M and Netflix shapes, ncu launch list + full capture of the SGD epoch kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_ml10m.log 2>&1; echo "rc=$?" >> gpurun_out/a_bench_ml10m.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/a_bench_ref.log
timeout 900 python bench.py --workload netflix --steps 5 --warmup 3 --no-cpu > gpurun_out/a_bench_netflix.log 2>&1; echo "rc=$?" >> gpurun_out/a_bench_netflix.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/a_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/a_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_epoch_kernel -s 3 -c 1 -f -o gpurun_out/a_prof_sgd python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/a_ncu_full.log 2>&1
echo done
