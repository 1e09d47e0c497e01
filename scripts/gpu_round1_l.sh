#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/l_bench_ml10m_$N.log 2>&1; echo "rc=$?" >> gpurun_out/l_bench_ml10m_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --workload netflix10 > gpurun_out/l_bench_nf10_$N.log 2>&1; echo "rc=$?" >> gpurun_out/l_bench_nf10_$N.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload netflix10 > gpurun_out/l_bench_nf10_1.log 2>&1; echo "rc=$?" >> gpurun_out/l_bench_nf10_1.log
echo done
