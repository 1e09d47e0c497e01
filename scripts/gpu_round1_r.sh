#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
MMLB200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/r_bench_$N.log 2>&1; echo "rc=$?" >> gpurun_out/r_bench_$N.log
echo done
