#!/bin/bash
# 8-GPU box: SGD ring at 8 GPUs after balancing the GPU-level item blocks (default grid 18 x 8, then 37 x 4)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
MMLB200_TRACE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/dd_bench_8.log 2>&1; echo "rc=$?" >> gpurun_out/dd_bench_8.log
MMLB200_TRACE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --cpg 4 > gpurun_out/dd_bench_8_cpg4.log 2>&1; echo "rc=$?" >> gpurun_out/dd_bench_8_cpg4.log
echo done
