#!/bin/bash
# 8-GPU box: SGD ring at 8 and 4 GPUs (with the per-epoch trace), WRMF row sharding at 8, top-N user sharding at 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
MMLB200_TRACE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/aa_bench_8.log 2>&1; echo "rc=$?" >> gpurun_out/aa_bench_8.log
MMLB200_TRACE=1 timeout 600 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/aa_bench_4.log 2>&1; echo "rc=$?" >> gpurun_out/aa_bench_4.log
timeout 600 $TR --nproc-per-node 8 --master-port 29543 scripts/bench_wrmf.py --epochs 4 > gpurun_out/aa_wrmf_8.log 2>&1; echo "rc=$?" >> gpurun_out/aa_wrmf_8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29544 scripts/bench_topn.py --reps 2 --exact-users 1024 > gpurun_out/aa_topn_8.log 2>&1; echo "rc=$?" >> gpurun_out/aa_topn_8.log
echo done
