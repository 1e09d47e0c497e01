#!/bin/bash
# final check of the session: full GPU suite, smoke, default bench + reference arm, config 5 top-N, config 3 WRMF
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/rr_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/rr_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/rr_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/rr_smoke.log
( time timeout 500 python bench.py ) > gpurun_out/rr_bench.log 2>&1; echo "rc=$?" >> gpurun_out/rr_bench.log
( time timeout 400 python bench.py --impl reference ) > gpurun_out/rr_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/rr_bench_ref.log
timeout 200 python scripts/bench_topn.py --reps 4 --exact-users 1024 > gpurun_out/rr_topn_c5.log 2>&1; echo "rc=$?" >> gpurun_out/rr_topn_c5.log
timeout 300 python scripts/bench_wrmf.py --epochs 4 > gpurun_out/rr_wrmf_c3.log 2>&1; echo "rc=$?" >> gpurun_out/rr_wrmf_c3.log
tail -n 6 gpurun_out/rr_pytest.log; tail -n 7 gpurun_out/rr_smoke.log; tail -n 2 gpurun_out/rr_topn_c5.log; tail -n 2 gpurun_out/rr_wrmf_c3.log
tail -c 900 gpurun_out/rr_bench.log
