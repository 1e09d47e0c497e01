#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/m_bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/m_bench_default.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/m_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/m_bench_ref.log
echo done
