#!/bin/bash
# 2-GPU box: ring / shard checks against the oracle after the block-balancing change
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist/wrmf_shard_check.py > gpurun_out/cc_wrmf_shard.log 2>&1; echo "rc=$?" >> gpurun_out/cc_wrmf_shard.log
timeout 600 $TR --master-port 29512 tests/dist/sgd_ring_check.py > gpurun_out/cc_sgd_ring.log 2>&1; echo "rc=$?" >> gpurun_out/cc_sgd_ring.log
MMLB200_TRACE=1 timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/cc_bench2.log 2>&1; echo "rc=$?" >> gpurun_out/cc_bench2.log
echo done
