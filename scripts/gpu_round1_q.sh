#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/q_topo.log 2>&1
df -h /dev/shm >> gpurun_out/q_topo.log 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P,SHM,NET timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --workload netflix10 > gpurun_out/q_bench_2.log 2>&1; echo "rc=$?" >> gpurun_out/q_bench_2.log
echo done
