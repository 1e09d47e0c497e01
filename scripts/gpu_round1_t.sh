#!/bin/bash
# 2-GPU box: WRMF row sharding + SGD ring checks against the oracle, grid-shape sweep of one 8-GPU sub-epoch, 2-GPU benches
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist/wrmf_shard_check.py > gpurun_out/t_wrmf_shard.log 2>&1; echo "rc=$?" >> gpurun_out/t_wrmf_shard.log
timeout 600 $TR --master-port 29512 tests/dist/sgd_ring_check.py > gpurun_out/t_sgd_ring.log 2>&1; echo "rc=$?" >> gpurun_out/t_sgd_ring.log
timeout 300 python scripts/sweep_groups.py --workload nf_sub8 --epochs 4 --shapes 148x1,74x2,37x4,18x8,9x16,1x148 > gpurun_out/t_sweep_sub8.log 2>&1; echo "rc=$?" >> gpurun_out/t_sweep_sub8.log
for cpg in 1 4 8; do
MMLB200_TRACE=1 timeout 600 $TR --master-port $((29520+cpg)) bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --cpg $cpg > gpurun_out/t_bench2_cpg$cpg.log 2>&1; echo "rc=$?" >> gpurun_out/t_bench2_cpg$cpg.log
done
timeout 600 python scripts/bench_wrmf.py --epochs 3 > gpurun_out/t_wrmf_c3_n1.log 2>&1; echo "rc=$?" >> gpurun_out/t_wrmf_c3_n1.log
timeout 600 $TR --master-port 29531 scripts/bench_wrmf.py --epochs 3 > gpurun_out/t_wrmf_c3_n2.log 2>&1; echo "rc=$?" >> gpurun_out/t_wrmf_c3_n2.log
echo done
