#!/bin/bash
mkdir -p gpurun_out
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/n_bench_$N.log 2>&1; echo "rc=$?" >> gpurun_out/n_bench_$N.log
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/n_bench_1.log 2>&1; echo "rc=$?" >> gpurun_out/n_bench_1.log
echo done
